#!/bin/bash
# Round 2, single GPU: full GPU suite (incl. live comparisons with the translated reference), default bench line with the
# translated reference as CPU baseline, the reference arm itself, c3 on one GPU, per-level costs.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export MGB_SPIN_TIMEOUT_MS=3000
timeout 1500 python -m pytest tests -m gpu -q -rs > gpurun_out/r02_gputest_n1_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest_n1_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/m_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/m_smoke.log
timeout 900 python bench.py > gpurun_out/m_bench_c5.json 2> gpurun_out/m_bench_c5.err
( time timeout 900 python bench.py --impl reference ) > gpurun_out/m_ref_c5.json 2> gpurun_out/m_ref_c5.err
timeout 600 python bench.py --config c3 --steps 3 --no-e2e > gpurun_out/m_bench_c3.json 2> gpurun_out/m_bench_c3.err
timeout 300 python scripts/level_costs.py 1024 2048 128 > gpurun_out/r02_level_costs_n1_c5.txt 2>&1
timeout 300 python scripts/level_costs.py 512 512 64 > gpurun_out/r02_level_costs_n1_c2.txt 2>&1
tail -n 6 gpurun_out/r02_gputest_n1_final.log; tail -n 3 gpurun_out/m_smoke.log
for f in gpurun_out/m_bench_c5.json gpurun_out/m_ref_c5.json gpurun_out/m_bench_c3.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"],2), d.get("fcycles_per_solve"), d.get("e2e"), d.get("cpu_baseline",{}).get("value"), d.get("cpu_baseline",{}).get("kind"), d.get("clocks"), {k:(round(v["ms"],4),round(v["frac"],3)) for k,v in d.get("roofline",{}).get("kernels",{}).items() if v})
except Exception as e:
    print("ERR", e, open(sys.argv[1].replace(".json",".err")).read()[-800:])
PY
done
tail -n 4 gpurun_out/m_ref_c5.err; grep -E "^lev|^sum" gpurun_out/r02_level_costs_n1_c5.txt gpurun_out/r02_level_costs_n1_c2.txt
