#!/bin/bash
# round-2 follow-up capture (r02b): the h = 30 size class after the tensor-core factorisation (csrc/chol8.h); summaries via
# tools/ncu_summary.py / tools/hot_lines.py -> profiles/r02b_fused_h30_kernel_*
O=gpurun_out
python tools/ncu_run2.py a1 30 1184 trot > $O/ncu_h30_plain.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:fused_kernelILi72E -c 1 -f \
    -o $O/r02b_fused_h30 python tools/ncu_run2.py a1 30 1184 trot > $O/ncu4.log 2>&1
ncu -i $O/r02b_fused_h30.ncu-rep --page raw --csv > $O/r02b_fused_h30_raw.csv 2>/dev/null
ncu -i $O/r02b_fused_h30.ncu-rep --page source --csv --print-source sass > $O/r02b_fused_h30_source.csv 2>/dev/null
ncu -i $O/r02b_fused_h30.ncu-rep --page source --csv --print-source cuda > $O/r02b_fused_h30_cuda.csv 2>/dev/null
gzip -f $O/r02b_fused_h30_source.csv $O/r02b_fused_h30_cuda.csv
rm -f $O/r02b_fused_h30.ncu-rep
tail -2 $O/ncu4.log
