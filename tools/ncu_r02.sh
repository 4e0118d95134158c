#!/bin/bash
# round-2 captures (r02): launch list of the bench command, full captures of the fused kernel at the bench's launch size
# (65536 instances, the <24,1,0> class), of the h = 30 size class <72,1,0> and of the WBC kernel, the phase-cycle
# profiles, and the sanitizer runs.  The .ncu-rep files are exported to CSV on the box and removed (gpurun_out/ is
# limited to 64 MiB).
set -x
O=gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_bench_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/r02_launches_raw.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_bench.log 2>&1
export_rep() {   # $1 = report stem
    ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_raw.csv 2>/dev/null
    ncu -i $O/$1.ncu-rep --page source --csv --print-source sass > $O/$1_source.csv 2>/dev/null
    gzip -f $O/$1_source.csv
    rm -f $O/$1.ncu-rep
}
python tools/ncu_run2.py a1 10 65536 trot > $O/ncu_run_plain.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:fused_kernelILi24E -c 1 -f \
    -o $O/r02_fused python tools/ncu_run2.py a1 10 65536 trot > $O/ncu2.log 2>&1
export_rep r02_fused
python tools/ncu_run2.py a1 30 1184 trot > $O/ncu_h30_plain.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:fused_kernelILi72E -c 1 -f \
    -o $O/r02_fused_h30 python tools/ncu_run2.py a1 30 1184 trot > $O/ncu4.log 2>&1
export_rep r02_fused_h30
python tools/ncu_wbc.py > $O/ncu_wbc_plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qr_wbc_kernel -c 1 -f -o $O/r02_wbc \
    python tools/ncu_wbc.py > $O/ncu3.log 2>&1
export_rep r02_wbc
python tools/prof.py > $O/r02_phase_cycles.txt 2>&1
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python tools/sanitize.py 0.05 > $O/r02_racecheck.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python tools/sanitize.py 0.05 > $O/r02_memcheck.log 2>&1
timeout 1200 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_concurrency.py -m gpu -q -k "two_streams or different_streams" > $O/r02_racecheck_concurrency.log 2>&1
for f in $O/r02_racecheck.log $O/r02_memcheck.log $O/r02_racecheck_concurrency.log; do echo "== $f"; tail -n 4 $f; done
du -sh $O
