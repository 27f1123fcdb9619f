#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export MGB_SPIN_TIMEOUT_MS=3000
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e $BARGS > gpurun_out/h_$name.json 2> gpurun_out/h_$name.err; }
BARGS="--config c2"
run c2_t0 MGB_THOMAS=0
run c2_t1 MGB_THOMAS=1
run c2_f_t0 MGB_FUSED_RELAX=1 MGB_THOMAS=0
run c2_f_t1 MGB_FUSED_RELAX=1 MGB_THOMAS=1
run c2_f_t0_v1 MGB_FUSED_RELAX=1 MGB_THOMAS=0 MGB_F_VARIANT=1
BARGS="--config c5"
run c5_t0 MGB_THOMAS=0
run c5_t1 MGB_THOMAS=1
run c5_f_t0 MGB_FUSED_RELAX=1 MGB_THOMAS=0
cp mgroms_b200/libmgb.so /tmp/libmgb_plain.so
MGB_BUILD_DEFS=MGB_F_TIMING python -m mgroms_b200.build > gpurun_out/h_build.log 2>&1
MGB_FUSED_RELAX=1 MGB_F_VARIANT=0 timeout 600 python bench.py --config c2 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/h_c2_timing.json 2> gpurun_out/h_c2_timing.err
grep "^F " gpurun_out/h_c2_timing.json | grep -E "tiles 1[89]" | sort | uniq -c | sort -rn | head -30 > gpurun_out/h_c2_timing.txt
cp /tmp/libmgb_plain.so mgroms_b200/libmgb.so; touch mgroms_b200/libmgb.so
timeout 900 python -m pytest tests/test_gpu_variants.py -m gpu -q > gpurun_out/h_variants.log 2>&1; echo "variants rc=$?" >> gpurun_out/h_variants.log
head -24 gpurun_out/h_c2_timing.txt
tail -n 6 gpurun_out/h_variants.log
for f in gpurun_out/h_c?_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"],2), d["fcycles_per_solve"], "%.3e"%d["final_rel_residual"], {k:(round(v["ms"],4),round(v["frac"],3)) for k,v in d["roofline"]["kernels"].items() if v}, d["roofline"]["relax_paths_by_level"][:3])
except Exception as e:
    print("ERR", e, open(sys.argv[1].replace(".json",".err")).read()[-500:])
PY
done
