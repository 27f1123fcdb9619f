#!/bin/bash
# Round 2, 8-GPU call (charged 8x): multi-GPU parity tests of the 8-rank configurations + the two fused 2x2 cases, the
# north-star bench line (c5 = 4096x4096x128 on 8 GPUs) with in-run parity and e2e, c4 (cuc stand-in 2048x2048x128) with
# nsmall = 32 (mg_gather) and nsmall = 8, per-level costs.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export MGB_SPIN_TIMEOUT_MS=5000
N=8
nvidia-smi -L > gpurun_out/n8_gpus.txt; nproc >> gpurun_out/n8_gpus.txt; free -g >> gpurun_out/n8_gpus.txt; nvidia-smi topo -m >> gpurun_out/n8_gpus.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/n8_bench_c5.json 2> gpurun_out/n8_bench_c5.err
timeout 300 $TR bench.py --gpus $N --steps 2 --warmup 3 --config c4 --no-e2e --no-parity > gpurun_out/n8_bench_c4_nsmall32.json 2> gpurun_out/n8_bench_c4_nsmall32.err
timeout 300 $TR bench.py --gpus $N --steps 2 --warmup 3 --config c4 --nsmall 8 --no-e2e --no-parity > gpurun_out/n8_bench_c4_nsmall8.json 2> gpurun_out/n8_bench_c4_nsmall8.err
MGB_TEST_ONLY_WORLD=8 timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -rs 2>&1 | tail -40 > gpurun_out/r02_gputest_multi_n8.log; echo "rc=${PIPESTATUS[0]}" >> gpurun_out/r02_gputest_multi_n8.log
MGB_TEST_ONLY_WORLD=4 timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "cuc-64x64x64-2x2" 2>&1 | tail -5 > gpurun_out/r02_gputest_multi_n4_fused_case.log
timeout 300 $TR scripts/level_costs_multi.py 1024 2048 128 8 > gpurun_out/r02_level_costs_n8_c5_nsmall8.txt 2>&1
tail -n 5 gpurun_out/r02_gputest_multi_n8.log; tail -n 3 gpurun_out/r02_gputest_multi_n4_fused_case.log
for f in gpurun_out/n8_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"],2), d["fcycles_per_solve"], "%.3e"%d["final_rel_residual"], d.get("e2e",{}).get("value"), (d.get("parity") or {}).get("transport"), d["config"].get("host_affinity"), {k:(round(v["ms"],4),round(v["frac"],3)) for k,v in d["roofline"]["kernels"].items() if v}, d["roofline"]["relax_paths_by_level"])
except Exception as e:
    print("ERR", e, open(sys.argv[1].replace(".json",".err")).read()[-800:])
PY
done
grep -E "^lev|^sum|GPUs" gpurun_out/r02_level_costs_n8_c5_nsmall8.txt
