#!/usr/bin/env bash
# round 2 evidence: ncu launch list of the bench command, --set full capture of the headline kernel (final build) and of the three-prime instance
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_launches_v3.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-aux --no-cggi > gpurun_out/r2_launches_v3.log 2>&1
echo "launch list rc=$?"; tail -2 gpurun_out/r2_launches_v3.log | cut -c1-300
KS_PIN=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gadget_kernel -s 2 -c 1 -f -o gpurun_out/prof_gadget_r2c python scripts/ks_prof.py > gpurun_out/prof_gadget_r2c.log 2>&1
echo "gadget full rc=$?"
KS_PIN=1 KS_KEY3=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gadget_kernel -s 2 -c 1 -f -o gpurun_out/prof_gadget_r2c_p3 python scripts/ks_prof.py > gpurun_out/prof_gadget_r2c_p3.log 2>&1
echo "gadget p3 full rc=$?"
