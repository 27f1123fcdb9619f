#!/bin/bash
# Round 2, 4 GPUs: the 4-rank multi-GPU tests (2x2: diagonal neighbours, corner rules, gathered levels, bmask), now also
# compared with the decomposed vectors of the translated reference.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export MGB_SPIN_TIMEOUT_MS=5000
MGB_TEST_ONLY_WORLD=4 timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -rs 2>&1 | tail -30 > gpurun_out/r02_gputest_multi_n4_refvectors.log; echo "rc=${PIPESTATUS[0]}" >> gpurun_out/r02_gputest_multi_n4_refvectors.log
tail -n 6 gpurun_out/r02_gputest_multi_n4_refvectors.log
