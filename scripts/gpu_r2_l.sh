#!/bin/bash
# Round 2, 2-GPU call: the 2-rank multi-GPU tests (now also compared with the decomposed reference vectors) and one bench
# line whose in-run parity check includes those vectors.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export MGB_SPIN_TIMEOUT_MS=5000
MGB_TEST_ONLY_WORLD=2 timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -rs 2>&1 | tail -30 > gpurun_out/r02_gputest_multi_n2_refvectors.log; echo "rc=${PIPESTATUS[0]}" >> gpurun_out/r02_gputest_multi_n2_refvectors.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 3 --warmup 3 --config c2 --no-e2e > gpurun_out/l_n2_bench_c2.json 2> gpurun_out/l_n2_bench_c2.err
tail -n 6 gpurun_out/r02_gputest_multi_n2_refvectors.log
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/l_n2_bench_c2.json").read().strip().splitlines()[-1]); print(round(d["value"],2), d["parity"])
except Exception as e:
    print("ERR", e, open("gpurun_out/l_n2_bench_c2.err").read()[-1500:])
PY
