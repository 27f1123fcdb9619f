#!/bin/bash
# Round 2, 2 GPUs: the forked plain-C multi-process driver over the C ABI.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export MGB_SPIN_TIMEOUT_MS=5000
( cd gpurun_out && timeout 120 ../examples/mg_testseamount_mp 2 1 32 32 16 ) > gpurun_out/r02_c_driver_mp_2x1.log 2>&1; echo "rc=$?" >> gpurun_out/r02_c_driver_mp_2x1.log
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "plain_c_multi_process" 2>&1 | tail -5 > gpurun_out/q_pytest.log
tail -n 22 gpurun_out/r02_c_driver_mp_2x1.log; tail -n 3 gpurun_out/q_pytest.log
