#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "MB=3"; timeout 300 python scripts/gadget_primes_perf.py 2>&1 | tail -3
echo "MB=4"; PGB_GADGET_MB=4 timeout 300 python scripts/gadget_primes_perf.py 2>&1 | tail -3
