#!/bin/bash
# round-1 final captures (r01f): launch list of the bench command + full capture of the fused kernel + WBC kernel
set -x
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 70 --csv --log-file gpurun_out/r01f_launches_raw.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python tools/ncu_run.py 4096 > gpurun_out/ncu_run_plain.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:qr_mpc_fused -f -o gpurun_out/fused_r01f \
    python tools/ncu_run.py 4096 > gpurun_out/ncu2.log 2>&1
python tools/ncu_wbc.py > gpurun_out/ncu_wbc_plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qr_wbc_kernel -c 1 -f -o gpurun_out/wbc_r01f \
    python tools/ncu_wbc.py > gpurun_out/ncu3.log 2>&1
python tools/prof.py > gpurun_out/phase_f.txt 2>&1
ls -la gpurun_out/*.ncu-rep
