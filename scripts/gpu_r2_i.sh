#!/bin/bash
# Round 2: ncu evidence for the four level-1 kernels (one ncu session per gpurun call: launch list + one --set full capture).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python scripts/profile_l1.py seamount 512 1024 128 2"
$CMD > gpurun_out/i_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_relax_A4|k_relax_B|k_relax_C|k_residual|k_restrict8_fused|k_prolong_bulk" -s 3 -c 14 -o gpurun_out/r02_level1_nz128 $CMD > gpurun_out/i_ncu_full.log 2>&1
CMD2="python bench.py --config c5 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/r02_launches_bench_c5.csv $CMD2 > gpurun_out/i_ncu_list.log 2>&1
tail -3 gpurun_out/i_plain.log gpurun_out/i_ncu_full.log gpurun_out/i_ncu_list.log
ls -la gpurun_out/r02_level1_nz128.ncu-rep
python scripts/launch_summary.py gpurun_out/r02_launches_bench_c5.csv | head -30
