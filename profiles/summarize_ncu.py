"""Turns an .ncu-rep (brought back in gpurun_out/) into the small text summary that is committed under profiles/.
usage: python profiles/summarize_ncu.py gpurun_out/X.ncu-rep profiles/NAME.txt "free-text header" """
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def main():
    rep, out, header = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(out, "w") as f:
        f.write(header.strip() + "\n")
        f.write(f"source: {rep} ({len(data)} launches captured; ncu --set full --clock-control none; cold-cache, serialised replays)\n")
        ki = hdr.index("Kernel Name")
        f.write("kernels: " + " | ".join(sorted(set(r[ki].split("(")[0] for r in data))) + "\n\n")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                f.write(f"{k} [{units[i]}]: {', '.join(r[i] for r in data)}\n")
        f.write("\nwarp stall reasons (warps per issue-active cycle):\n")
        for h in hdr:
            if "average_warps_issue_stalled" in h and "per_issue_active" in h:
                i = hdr.index(h)
                v = [float(r[i]) for r in data]
                if max(v) > 0.3:
                    name = h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")
                    f.write(f"  {name}: {', '.join(f'{x:.2f}' for x in v)}\n")
    print(open(out).read())


if __name__ == "__main__":
    main()
