#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export MGB_SPIN_TIMEOUT_MS=3000
timeout 1200 python -m pytest tests/test_gpu_variants.py -m gpu -q > gpurun_out/c_variants.log 2>&1; echo "variants rc=$?" >> gpurun_out/c_variants.log
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e $BARGS > gpurun_out/c_$name.json 2> gpurun_out/c_$name.err; }
BARGS="--config c2"
for v in 0 1 2 3; do run c2_v$v MGB_F_VARIANT=$v; done
run c2_v0_skipS MGB_F_VARIANT=0 MGB_F_SKIP_S=1
run c2_v1_skipS MGB_F_VARIANT=1 MGB_F_SKIP_S=1
run c2_old MGB_NO_FUSED_RELAX=1
BARGS="--config c5"
run c5_v0 MGB_F_VARIANT=0
run c5_v1 MGB_F_VARIANT=1
run c5_v0_skipS MGB_F_VARIANT=0 MGB_F_SKIP_S=1
run c5_old8 MGB_NO_FUSED_RELAX=1 MGB_A4_TCMIN=8
run c5_old16 MGB_NO_FUSED_RELAX=1 MGB_A4_TCMIN=16
timeout 1500 python -m pytest tests/test_gpu_large.py -m gpu -x -q > gpurun_out/c_large.log 2>&1; echo "large rc=$?" >> gpurun_out/c_large.log
tail -n 5 gpurun_out/c_variants.log; tail -n 5 gpurun_out/c_large.log
for f in gpurun_out/c_c*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"],2), d["fcycles_per_solve"], "%.3e"%d["final_rel_residual"], {k:(round(v["ms"],4),round(v["frac"],3)) for k,v in d["roofline"]["kernels"].items() if v}, d["roofline"]["relax_paths_by_level"][:3])
except Exception as e:
    print("ERR", e, open(sys.argv[1].replace(".json",".err")).read()[-500:])
PY
done
