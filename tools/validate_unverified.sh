#!/bin/bash
# First GPU call of the next session: run everything that was written without hardware access
# (tests gated by DORY_TEST_UNVERIFIED, the apply-first bench arm, GAT source windows) and leave the
# results under gpurun_out/.  Usage (≈6 GPU-minutes):
#   gpurun --timeout 900 -- 'bash tools/validate_unverified.sh'
set -x
mkdir -p gpurun_out
DORY_TEST_UNVERIFIED=1 timeout 600 python -m pytest tests/test_gpu_zzz_apply_first.py tests/test_gpu_zz_lambda_golden.py \
    -q -m gpu 2>&1 | tail -40 > gpurun_out/unverified_tests.log
cat gpurun_out/unverified_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-apply-first-arm > gpurun_out/bench_reference_order.json 2> gpurun_out/bench_reference_order.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --apply-first > gpurun_out/bench_apply_first.json 2> gpurun_out/bench_apply_first.log
cat gpurun_out/bench_reference_order.json gpurun_out/bench_apply_first.json
# Reddit GAT (configs[2]) with and without source windows
timeout 600 python tools/shape_bench.py --name reddit --gnn GAT --out gpurun_out/shape_reddit_gat.json > /dev/null 2> gpurun_out/shape_gat.log
timeout 600 python tools/shape_bench.py --name reddit --gnn GAT --opt gat_windows=1 --out gpurun_out/shape_reddit_gat_windows.json > /dev/null 2>> gpurun_out/shape_gat.log
python - <<'PY'
import json
for n in ("shape_reddit_gat", "shape_reddit_gat_windows"):
    try:
        d = json.load(open("gpurun_out/%s.json" % n))
        print(n, "epoch_ms", d["epoch_ms"], [(l["layer"], l["dir"], round(l["ms"], 3)) for l in d["layers"]])
    except Exception as ex:
        print(n, "missing:", ex)
PY
# with >= 2 GPUs (gpurun --gpus 2): the apply-first exchanges (t forward, dL/dz backward) on both paths,
# and the C++ driver's multi-partition mode
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
    for ex in p2p nccl; do
        timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
            tools/multi_gpu_check.py --apply-first --exchange $ex 2>&1 | grep -E "rank|MULTI_GPU" | tee -a gpurun_out/multi_apply_first.log
    done
    for ex in p2p nccl; do  # GAT on several GPUs: never run on hardware either
        timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 \
            tools/multi_gpu_check.py --gnn GAT --exchange $ex 2>&1 | grep -E "rank|MULTI_GPU" | tee -a gpurun_out/multi_gat.log
    done
    DORY_TEST_UNVERIFIED=1 timeout 900 python -m pytest tests/test_gpu_zzz_apply_first.py -q -m gpu -k cpp_driver 2>&1 | tail -20 | tee gpurun_out/cpp_multi.log
fi
# narrow-row kernel shapes for the apply-first widths (incl. the new 4 x 4 shape)
timeout 900 python tools/width_sweep.py --widths 41,64 --out gpurun_out/width_sweep.json > gpurun_out/width_sweep.log 2>&1
tail -30 gpurun_out/width_sweep.log
# ncu evidence for the apply-first schedule: launch list of the bench command (kernel shares of the step)
# and one full capture of its aggregation launches (F = 128 and F = 41 rows)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_apply_first.csv \
    python bench.py --apply-first --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_apply_first.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm -s 8 -c 8 -o gpurun_out/apply_first_spmm_full -f \
    python bench.py --apply-first --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_apply_first.log 2>&1
ls -la gpurun_out/
