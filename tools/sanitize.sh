#!/bin/bash
# compute-sanitizer over the small-scene GPU parity tests (golden / vanilla / uint8 / training family): memcheck, then racecheck
# on the shared-memory heavy ones.  Usage (on the GPU box): bash tools/sanitize.sh <tag>
tag=${1:-r2}; out=gpurun_out
timeout 420 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
    -k "golden or vanilla or uint8 or statistics_vs_oracle" 2>&1 | grep -E "COMPUTE-SANITIZER|passed|failed|ERROR SUMMARY|Invalid|error" | head -40 > $out/${tag}_sanitizer_memcheck.log
timeout 420 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
    -k "fov_matches_reference_golden or sum_forward_and_backward or vanilla_forward or uint8" 2>&1 | grep -E "COMPUTE-SANITIZER|passed|failed|RACECHECK SUMMARY|hazard|error" | head -40 > $out/${tag}_sanitizer_racecheck.log
cat $out/${tag}_sanitizer_memcheck.log $out/${tag}_sanitizer_racecheck.log
