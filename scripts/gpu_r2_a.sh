#!/bin/bash
# Round 2, first GPU call: correctness of the fused relaxation kernel, then quick timings of its warp layouts.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export MGB_SPIN_TIMEOUT_MS=3000
nvidia-smi -L > gpurun_out/a_gpus.txt 2>&1
free -g > gpurun_out/a_host.txt; nproc >> gpurun_out/a_host.txt
timeout 900 python -m pytest tests/test_gpu_variants.py -m gpu -x -q > gpurun_out/a_variants.log 2>&1; echo "variants rc=$?" >> gpurun_out/a_variants.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/a_parity.log 2>&1; echo "parity rc=$?" >> gpurun_out/a_parity.log
for v in 0 1 2 3; do
  MGB_F_VARIANT=$v timeout 300 python bench.py --config c2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/a_c2_v$v.json 2> gpurun_out/a_c2_v$v.err
done
MGB_NO_FUSED_RELAX=1 timeout 300 python bench.py --config c2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/a_c2_old.json 2> gpurun_out/a_c2_old.err
timeout 600 python bench.py --config c5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/a_c5.json 2> gpurun_out/a_c5.err
MGB_NO_FUSED_RELAX=1 timeout 600 python bench.py --config c5 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/a_c5_old.json 2> gpurun_out/a_c5_old.err
tail -3 gpurun_out/a_variants.log gpurun_out/a_parity.log
for f in gpurun_out/a_c2_v*.json gpurun_out/a_c2_old.json gpurun_out/a_c5.json gpurun_out/a_c5_old.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["value"], d["fcycles_per_solve"], d.get("e2e",{}).get("value"), {k:(round(v["ms"],4),round(v["frac"],3)) for k,v in d["roofline"]["kernels"].items() if v}, d["roofline"]["relax_paths_by_level"])
except Exception as e:
    print("ERR", e)
PY
done
