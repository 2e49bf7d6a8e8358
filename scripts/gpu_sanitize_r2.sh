#!/usr/bin/env bash
# compute-sanitizer over the kernels written or rewritten in round 2 (NTT120 gadget kernel incl. the three-prime instance, the NTT120
# whole-rotation CGGI kernel, FFT64 whole-rotation kernel v4): racecheck, then memcheck
set -u
mkdir -p gpurun_out
CS="compute-sanitizer --print-limit 100 --error-exitcode 7"
run() { # name, tool, pytest args...
  local name=$1 tool=$2; shift 2
  timeout 1500 $CS --tool $tool python -m pytest "$@" -m gpu -q -x > gpurun_out/san_${name}.log 2>&1
  echo "$name rc=$?"; grep -E "passed|failed|SUMMARY|ERROR SUMMARY" gpurun_out/san_${name}.log | tail -4
}
run race_gadget racecheck tests/test_gpu_gadget_primes.py tests/test_gpu_core.py -k "three_prime or worst_case or unpinned or external_product_three or automorphism_family or gadget_single_kernel or pinned"
run race_cggi racecheck tests/test_gpu_cggi.py -k "test_blind_rotate_matches_oracle or test_blind_rotate_tall_keys"
run mem_gadget memcheck tests/test_gpu_gadget_primes.py tests/test_gpu_core.py tests/test_gpu_trace.py
run mem_cggi memcheck tests/test_gpu_cggi.py
