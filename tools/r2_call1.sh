#!/bin/bash
# round 2, GPU call 1: data for the kernel work (narrow-row shapes, GAT with/without windows, low-degree baselines)
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 500 python tools/width_sweep.py --widths 41,64 --out gpurun_out/r2_width_sweep.json > gpurun_out/r2_width_sweep.log 2>&1 &
P1=$!
wait $P1
tail -30 gpurun_out/r2_width_sweep.log
timeout 400 python tools/shape_bench.py --name reddit --gnn GAT --out gpurun_out/r2_shape_reddit_gat.json > /dev/null 2> gpurun_out/r2_shape_gat.log
timeout 400 python tools/shape_bench.py --name reddit --gnn GAT --opt gat_windows=1 --out gpurun_out/r2_shape_reddit_gat_windows.json > /dev/null 2>> gpurun_out/r2_shape_gat.log
tail -5 gpurun_out/r2_shape_gat.log
ls -la gpurun_out
