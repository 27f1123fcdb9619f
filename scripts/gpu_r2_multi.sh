#!/bin/bash
# Round 2, N-GPU call: multi-GPU parity tests of the N-rank configurations (both transports), bench lines, per-level costs.
#   gpurun --gpus N -- bash scripts/gpu_r2_multi.sh N
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export MGB_SPIN_TIMEOUT_MS=5000
N=${1:-4}
nvidia-smi -L > gpurun_out/n${N}_gpus.txt
MGB_TEST_ONLY_WORLD=$N timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q -rs 2>&1 | tail -40 > gpurun_out/r02_gputest_multi_n${N}.log; echo "rc=${PIPESTATUS[0]}" >> gpurun_out/r02_gputest_multi_n${N}.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/n${N}_bench_c5.json 2> gpurun_out/n${N}_bench_c5.err
timeout 600 $TR bench.py --gpus $N --steps 2 --warmup 3 --config c4 --no-e2e --no-parity > gpurun_out/n${N}_bench_c4_nsmall32.json 2> gpurun_out/n${N}_bench_c4_nsmall32.err
if [ "$N" = "8" ]; then
  timeout 600 $TR bench.py --gpus $N --steps 2 --warmup 3 --config c4 --nsmall 8 --no-e2e --no-parity > gpurun_out/n${N}_bench_c4_nsmall8.json 2> gpurun_out/n${N}_bench_c4_nsmall8.err
else
  timeout 600 $TR bench.py --gpus $N --steps 2 --warmup 3 --config c3 --no-e2e --no-parity > gpurun_out/n${N}_bench_c3.json 2> gpurun_out/n${N}_bench_c3.err
fi
timeout 600 $TR bench.py --gpus $N --steps 3 --warmup 3 --config c2 --no-e2e --no-parity > gpurun_out/n${N}_bench_c2.json 2> gpurun_out/n${N}_bench_c2.err
timeout 600 $TR scripts/level_costs_multi.py 1024 2048 128 8 > gpurun_out/r02_level_costs_n${N}_c5_nsmall8.txt 2>&1
timeout 600 $TR scripts/level_costs_multi.py 512 512 64 8 > gpurun_out/r02_level_costs_n${N}_c2_nsmall8.txt 2>&1
tail -n 5 gpurun_out/r02_gputest_multi_n${N}.log
for f in gpurun_out/n${N}_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"],2), d["fcycles_per_solve"], "%.3e"%d["final_rel_residual"], d.get("e2e",{}).get("value"), (d.get("parity") or {}).get("transport"), {k:(round(v["ms"],4),round(v["frac"],3)) for k,v in d["roofline"]["kernels"].items() if v}, d["roofline"]["relax_paths_by_level"])
except Exception as e:
    print("ERR", e, open(sys.argv[1].replace(".json",".err")).read()[-800:])
PY
done
grep -E "^lev|^sum|GPUs" gpurun_out/r02_level_costs_n${N}_c5_nsmall8.txt gpurun_out/r02_level_costs_n${N}_c2_nsmall8.txt
