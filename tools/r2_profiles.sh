#!/bin/bash
# round 2: the ncu evidence kept under profiles/ (launch lists of the bench command in both schedules, full
# captures of the aggregation kernels of the apply-first schedule, of the staged kernel on the community graph
# and of the low-degree kernel on the Friendster/8 shape).  One GPU, ~4 GPU-minutes.
set -x
mkdir -p gpurun_out
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_reference_order.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-arms > gpurun_out/r2_ncu_bench_ref.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_apply_first.csv \
    python bench.py --apply-first --steps 2 --warmup 3 --no-cpu-baseline --no-arms > gpurun_out/r2_ncu_bench_af.log 2>&1
timeout 250 ncu --set full --clock-control none --import-source on -k regex:spmm -s 8 -c 8 -o gpurun_out/r2_apply_first_spmm_full -f \
    python bench.py --apply-first --steps 1 --warmup 3 --no-cpu-baseline --no-arms > gpurun_out/r2_ncu_full_af.log 2>&1
timeout 250 ncu --set full --clock-control none --import-source on -k regex:spmm_tile -c 3 -o gpurun_out/r2_tile_default_full -f \
    python tools/tile_bench.py --config reddit-communities --generator chunglu --variants "tile=2" --reps 1 > gpurun_out/r2_ncu_full_tile.log 2>&1
timeout 250 ncu --set full --clock-control none --import-source on -k regex:spmm_group -s 2 -c 3 -o gpurun_out/r2_friendster8_spmm_full -f \
    python tools/op_breakdown.py --workload friendster --emulate-parts 8 --sustain 1 --reps 1 > gpurun_out/r2_ncu_full_f8.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r2_launches_*.csv
