#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
CGGI_FL=ntt120 timeout 600 ncu --set full --clock-control none --import-source on -k regex:cggi_fused -s 1 -c 1 -f -o gpurun_out/prof_cggi_ntt120_v5 python scripts/cggi_prof.py > gpurun_out/prof_cggi_v5.log 2>&1
tail -2 gpurun_out/prof_cggi_v5.log
