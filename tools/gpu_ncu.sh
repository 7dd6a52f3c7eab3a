python tools/ncu_target.py > gpurun_out/ncu_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_spmv_sell|k_update_lagged|k_multidot|k_combine|k_rank_pass' -c 36 -o gpurun_out/r02_hot python tools/ncu_target.py > gpurun_out/ncu_f.log 2>&1
ls -la gpurun_out/r02_hot.ncu-rep; tail -2 gpurun_out/ncu_f.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-f32-detail --no-reorth-detail > gpurun_out/bench_ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 420 --csv --log-file gpurun_out/r02_launches_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-f32-detail --no-reorth-detail > gpurun_out/ncu_b.log 2>&1
tail -2 gpurun_out/ncu_b.log; wc -l gpurun_out/r02_launches_c3.csv
