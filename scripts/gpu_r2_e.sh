#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export MGB_SPIN_TIMEOUT_MS=3000
# instrumented build of the fused kernel (clock64 per role), separate library so that the shipped one stays untouched
cp mgroms_b200/libmgb.so /tmp/libmgb_plain.so
MGB_BUILD_DEFS=MGB_F_TIMING python -m mgroms_b200.build > gpurun_out/e_build.log 2>&1
MGB_F_VARIANT=0 timeout 600 python bench.py --config c2 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/e_c2_timing.json 2> gpurun_out/e_c2_timing.err
grep "^F " gpurun_out/e_c2_timing.json | sort | uniq -c | sort -rn | head -40 > gpurun_out/e_c2_timing.txt
MGB_F_VARIANT=0 timeout 600 python bench.py --config c5 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/e_c5_timing.json 2> gpurun_out/e_c5_timing.err
grep "^F " gpurun_out/e_c5_timing.json | sort | uniq -c | sort -rn | head -40 > gpurun_out/e_c5_timing.txt
cp /tmp/libmgb_plain.so mgroms_b200/libmgb.so; touch mgroms_b200/libmgb.so
timeout 900 python -m pytest tests/test_gpu_variants.py -m gpu -q > gpurun_out/e_variants.log 2>&1; echo "variants rc=$?" >> gpurun_out/e_variants.log
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e $BARGS > gpurun_out/e_$name.json 2> gpurun_out/e_$name.err; }
BARGS="--config c2"
run c2_bulk MGB_NO_FUSED_RELAX=1
run c2_nobulk MGB_NO_FUSED_RELAX=1 MGB_NO_BULK_PROLONG=1
BARGS="--config c5"
run c5_bulk MGB_NO_FUSED_RELAX=1
run c5_bulk6 MGB_NO_FUSED_RELAX=1 MGB_BULK_PROLONG_CTAS=6
run c5_bulk2 MGB_NO_FUSED_RELAX=1 MGB_BULK_PROLONG_CTAS=2
head -30 gpurun_out/e_c2_timing.txt; head -30 gpurun_out/e_c5_timing.txt
tail -n 6 gpurun_out/e_variants.log
for f in gpurun_out/e_c?_*bulk*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"],2), d["fcycles_per_solve"], "%.3e"%d["final_rel_residual"], {k:(round(v["ms"],4),round(v["frac"],3)) for k,v in d["roofline"]["kernels"].items() if v}, d["roofline"]["relax_paths_by_level"][:3])
except Exception as e:
    print("ERR", e, open(sys.argv[1].replace(".json",".err")).read()[-500:])
PY
done
