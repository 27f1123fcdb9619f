#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export MGB_SPIN_TIMEOUT_MS=3000
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e $BARGS > gpurun_out/j_$name.json 2> gpurun_out/j_$name.err; }
BARGS="--config c2"
run c2_f MGB_FUSED_RELAX=1
run c2_f_v1 MGB_FUSED_RELAX=1 MGB_F_VARIANT=1
BARGS="--config c5"
run c5_f MGB_FUSED_RELAX=1
cp mgroms_b200/libmgb.so /tmp/libmgb_plain.so
MGB_BUILD_DEFS=MGB_F_TIMING python -m mgroms_b200.build > gpurun_out/j_build.log 2>&1
MGB_FUSED_RELAX=1 timeout 600 python bench.py --config c2 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/j_c2_timing.json 2> gpurun_out/j_c2_timing.err
grep "^F " gpurun_out/j_c2_timing.json | grep -E "tiles 1[89]" | sort | uniq -c | sort -rn | head -40 > gpurun_out/j_c2_timing.txt
cp /tmp/libmgb_plain.so mgroms_b200/libmgb.so; touch mgroms_b200/libmgb.so
timeout 900 python -m pytest tests/test_gpu_variants.py -m gpu -q -k "fused" > gpurun_out/j_variants.log 2>&1; echo "variants rc=$?" >> gpurun_out/j_variants.log
CMD="python scripts/profile_l1.py seamount 512 1024 128 2"
$CMD > gpurun_out/j_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_relax_A4|k_relax_B|k_relax_C|k_residual|k_restrict8_fused|k_prolong_bulk" -s 19 -c 11 -o gpurun_out/r02_level1_nz128_others $CMD > gpurun_out/j_ncu_full.log 2>&1
head -30 gpurun_out/j_c2_timing.txt
tail -n 4 gpurun_out/j_variants.log; tail -n 3 gpurun_out/j_ncu_full.log
for f in gpurun_out/j_c?_f*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"],2), d["fcycles_per_solve"], "%.3e"%d["final_rel_residual"], {k:(round(v["ms"],4),round(v["frac"],3)) for k,v in d["roofline"]["kernels"].items() if v}, d["roofline"]["relax_paths_by_level"][:3])
except Exception as e:
    print("ERR", e, open(sys.argv[1].replace(".json",".err")).read()[-500:])
PY
done
