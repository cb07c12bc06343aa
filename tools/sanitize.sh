#!/bin/bash
# compute-sanitizer over the small GPU parity cases: memcheck, then racecheck (shared-memory hazards)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== memcheck"
timeout 280 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_parity.py -m gpu -x -q -k "generic or other_block_sizes or bfr_block_sizes or combined or world or x8x16 or one_pixel" > gpurun_out/sanitize_memcheck.log 2>&1
tail -2 gpurun_out/sanitize_memcheck.log | head -1; grep "ERROR SUMMARY" gpurun_out/sanitize_memcheck.log | sort | uniq -c
echo "== racecheck"
timeout 400 compute-sanitizer --tool racecheck --print-limit 40 python -m pytest tests/test_parity.py -m gpu -x -q -k "generic or other_block_sizes or bfr_block_sizes or fixed_swizzle or x8x16 or negative" > gpurun_out/sanitize_racecheck.log 2>&1
grep "passed\|failed" gpurun_out/sanitize_racecheck.log | tail -1; grep "RACECHECK SUMMARY" gpurun_out/sanitize_racecheck.log | sort | uniq -c
grep -A1 "Race reported" gpurun_out/sanitize_racecheck.log | grep -o "in [a-z_]*\.cu[h]*:[0-9]*" | sort | uniq -c | head
