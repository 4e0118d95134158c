#!/bin/bash
# regenerate profiles/r01f_* from gpurun_out/ after tools/ncu_final.sh ran on the GPU box
set -e
cd "$(dirname "$0")/.."
cp gpurun_out/r01f_launches_raw.csv profiles/r01f_launches_raw.csv
python tools/ncu_summary.py gpurun_out/fused_r01f.ncu-rep profiles/r01f_fused_kernel_ncu_metrics.txt "fused_kernel<24"
python tools/ncu_summary.py gpurun_out/wbc_r01f.ncu-rep profiles/r01f_wbc_kernel_ncu_metrics.txt
ncu -i gpurun_out/fused_r01f.ncu-rep --page source --csv --print-source sass > gpurun_out/src_f.csv 2>/dev/null
python tools/hot_lines.py gpurun_out/src_f.csv profiles/r01f_fused_kernel_hot_lines.txt 4096 "ncu --set full, source page, 4096 A1 h=10 trot instances; round-1 final kernel: 96 threads x 6 CTAs/SM, coarse active-set prediction, cycle handling" | head -4
