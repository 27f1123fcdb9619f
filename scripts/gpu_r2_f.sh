#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export MGB_SPIN_TIMEOUT_MS=3000
cp mgroms_b200/libmgb.so /tmp/libmgb_plain.so
MGB_BUILD_DEFS=MGB_F_TIMING python -m mgroms_b200.build > gpurun_out/f_build.log 2>&1
MGB_F_VARIANT=0 timeout 600 python bench.py --config c2 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/f_c2_timing.json 2> gpurun_out/f_c2_timing.err
grep "^F " gpurun_out/f_c2_timing.json | sort | uniq -c | sort -rn | head -60 > gpurun_out/f_c2_timing.txt
MGB_F_VARIANT=0 timeout 600 python bench.py --config c5 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/f_c5_timing.json 2> gpurun_out/f_c5_timing.err
grep "^F " gpurun_out/f_c5_timing.json | sort | uniq -c | sort -rn | head -60 > gpurun_out/f_c5_timing.txt
cp /tmp/libmgb_plain.so mgroms_b200/libmgb.so; touch mgroms_b200/libmgb.so
head -40 gpurun_out/f_c2_timing.txt; head -40 gpurun_out/f_c5_timing.txt
