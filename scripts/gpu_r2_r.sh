#!/bin/bash
# Round 2, single GPU: the configs[1] line (512x512x64) with the translated reference as CPU arm.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py --config c2 --steps 10 > gpurun_out/r_bench_c2.json 2> gpurun_out/r_bench_c2.err
timeout 600 python bench.py --impl reference --config c2 > gpurun_out/r_ref_c2.json 2> gpurun_out/r_ref_c2.err
for f in gpurun_out/r_bench_c2.json gpurun_out/r_ref_c2.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"],2), d.get("fcycles_per_solve"), d.get("e2e",{}).get("value"), d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], d.get("clocks"))
except Exception as e:
    print("ERR", e, open(sys.argv[1].replace(".json",".err")).read()[-800:])
PY
done
