#!/bin/bash
# Round 2, single GPU: the full GPU test-suite (incl. the reference-vector tests), the default bench line (c5) with e2e and
# cpu_baseline, the reference arm, and the c2 line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export MGB_SPIN_TIMEOUT_MS=3000
nvidia-smi -L > gpurun_out/k_gpus.txt; nproc >> gpurun_out/k_gpus.txt; free -g >> gpurun_out/k_gpus.txt
timeout 1500 python -m pytest tests -m gpu -q -rs > gpurun_out/r02_gputest_n1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest_n1.log
timeout 900 python bench.py > gpurun_out/k_bench_c5.json 2> gpurun_out/k_bench_c5.err
( time timeout 900 python bench.py --impl reference ) > gpurun_out/k_ref_c5.json 2> gpurun_out/k_ref_c5.err
timeout 600 python bench.py --config c2 --steps 10 > gpurun_out/k_bench_c2.json 2> gpurun_out/k_bench_c2.err
( time timeout 600 python bench.py --impl reference --config c2 ) > gpurun_out/k_ref_c2.json 2> gpurun_out/k_ref_c2.err
tail -n 8 gpurun_out/r02_gputest_n1.log
for f in gpurun_out/k_bench_c5.json gpurun_out/k_bench_c2.json gpurun_out/k_ref_c5.json gpurun_out/k_ref_c2.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"],2), d.get("fcycles_per_solve"), d.get("e2e"), d.get("cpu_baseline",{}).get("value"), d.get("clocks"), {k:(round(v["ms"],4),round(v["frac"],3)) for k,v in d.get("roofline",{}).get("kernels",{}).items() if v})
except Exception as e:
    print("ERR", e, open(sys.argv[1].replace(".json",".err")).read()[-800:])
PY
done
tail -n 4 gpurun_out/k_ref_c5.err
