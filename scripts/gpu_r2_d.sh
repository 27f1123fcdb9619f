#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export MGB_SPIN_TIMEOUT_MS=3000
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d_pytest.log
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e $BARGS > gpurun_out/d_$name.json 2> gpurun_out/d_$name.err; }
BARGS="--config c2"
run c2_v0 MGB_F_VARIANT=0
run c2_v1 MGB_F_VARIANT=1
BARGS="--config c5"
run c5_v0 MGB_F_VARIANT=0
run c5_v1 MGB_F_VARIANT=1
run c5_old8 MGB_NO_FUSED_RELAX=1
tail -n 8 gpurun_out/d_pytest.log
for f in gpurun_out/d_c*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"],2), d["fcycles_per_solve"], "%.3e"%d["final_rel_residual"], {k:(round(v["ms"],4),round(v["frac"],3)) for k,v in d["roofline"]["kernels"].items() if v}, d["roofline"]["relax_paths_by_level"][:3])
except Exception as e:
    print("ERR", e, open(sys.argv[1].replace(".json",".err")).read()[-500:])
PY
done
