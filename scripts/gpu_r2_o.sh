#!/bin/bash
# Round 2, 2 GPUs: the nsmall crossover the VERDICT asked for (per-level costs with the coarse levels gathered early / late).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export MGB_SPIN_TIMEOUT_MS=5000
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for ns in 64 256; do
  timeout 300 $TR scripts/level_costs_multi.py 512 512 64 $ns > gpurun_out/r02_level_costs_n2_c2_nsmall$ns.txt 2>&1
done
timeout 400 $TR scripts/level_costs_multi.py 1024 2048 128 64 > gpurun_out/r02_level_costs_n2_c5_nsmall64.txt 2>&1
grep -E "^lev|^sum|GPUs" gpurun_out/r02_level_costs_n2_c2_nsmall64.txt gpurun_out/r02_level_costs_n2_c2_nsmall256.txt gpurun_out/r02_level_costs_n2_c5_nsmall64.txt
