"""Launch list (ncu --metrics gpu__time_duration.sum --csv) -> per-kernel share table committed under profiles/.
usage: python profiles/summarize_launches.py gpurun_out/X.csv profiles/NAME_summary.txt "free-text header" """
import csv
import re
import sys
from collections import OrderedDict


def main():
    src, out, header = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr, data = rows[0], rows[1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    acc = OrderedDict()
    for r in data:
        name = re.sub(r"^(void\s+)?(<unnamed>::)?", "", r[ki]).split("(")[0]
        t = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
        n, s = acc.get(name, (0, 0.0))
        acc[name] = (n + 1, s + t)
    total = sum(s for _, s in acc.values())
    with open(out, "w") as f:
        f.write(header.strip() + "\n")
        f.write(f"({len(data)} consecutive launches inside the step loop; per-launch times are cold-cache and serialised: compare SHARES)\n\n")
        for name, (n, s) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{name:<28} launches {n:4d}  total {s:10.1f} us  share {100 * s / total:5.1f}%  mean {s / n:8.1f} us\n")


if __name__ == "__main__":
    main()
