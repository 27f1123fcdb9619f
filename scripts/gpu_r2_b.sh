#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
./scripts/micro/thomas > gpurun_out/b_thomas.txt 2>&1
MGB_NO_FUSED_RELAX=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/b_launches_c5_old.csv python bench.py --config c5 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b_ncu_c5.log 2>&1
cat gpurun_out/b_thomas.txt
python scripts/launch_summary.py gpurun_out/b_launches_c5_old.csv 2>&1 | head -40
