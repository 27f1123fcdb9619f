#!/bin/bash
# Round 2: ncu --set full of the four level-1 kernels on the configs[1] block (512x512x64, nz = 64) with the final code.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python scripts/profile_l1.py seamount 512 512 64 2"
$CMD > gpurun_out/s_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_relax_A4|k_relax_B|k_relax_C|k_residual|k_restrict8_fused|k_prolong_bulk|k_prolong8" -c 40 -o gpurun_out/r02_level1_nz64 $CMD > gpurun_out/s_ncu_full.log 2>&1
tail -3 gpurun_out/s_plain.log gpurun_out/s_ncu_full.log
ls -la gpurun_out/r02_level1_nz64.ncu-rep
