#!/bin/bash
# ncu evidence of round 2 (run under gpurun, one GPU):
#  1. 4096^2: one-pass metrics only (DRAM bytes + duration) of every launch of ONE solve -- no --set full on a 100 GB working set
#  2. (argument "full") 1024^2: --set full of the dominant kernels over the first two iterations of a solve; the report is exported
#     to CSV pages on the box (raw metrics, source view of the kernels of interest) because gpurun returns at most 64 MiB
set -x
K='regex:stencil_rb_lr|gcr_ortho|gcr_dots|gcr_step|restrict_chiral|prolong_chiral|wilson_rb2|wilson_march'
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --cache-control none \
    --profile-from-start off -c 1500 --csv --log-file gpurun_out/r02_ncu_L4096_launches.csv \
    python bench.py --profile-step --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_L4096.log 2>&1
if [ "$1" == "full" ]; then
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k "$K" -c 40 -o /tmp/prof_r02 -f \
    python bench.py --L 1024 --profile-step --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_L1024.log 2>&1
ncu -i /tmp/prof_r02.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_L1024_raw.csv 2>/dev/null
ncu -i /tmp/prof_r02.ncu-rep --page source --csv -k regex:stencil_rb_lr -c 2 > gpurun_out/r02_ncu_source_rb_lr.csv 2>/dev/null
ncu -i /tmp/prof_r02.ncu-rep --page source --csv -k regex:gcr_ortho -c 2 > gpurun_out/r02_ncu_source_gcr_ortho.csv 2>/dev/null
ncu -i /tmp/prof_r02.ncu-rep --page source --csv -k regex:restrict_chiral -c 1 > gpurun_out/r02_ncu_source_restrict.csv 2>/dev/null
if [ $(stat -c %s /tmp/prof_r02.ncu-rep) -lt 40000000 ]; then cp /tmp/prof_r02.ncu-rep gpurun_out/; fi
fi
du -sh gpurun_out
