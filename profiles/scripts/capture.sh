#!/bin/bash
# Capture the ncu evidence for one round (run under gpurun, one GPU):  bash profiles/scripts/capture.sh r01
# 1. launch list of the bench command (gpu__time_duration.sum per launch; the kernel SHARES of the step are what matter)
# 2. one `--set full` capture each of the recurrence kernels, the fused loss, the tensor-core GEMM and the fused sampler
set -u
TAG=${1:-r01}
OUT=gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu --no-sampler"
timeout 300 $BENCH > $OUT/${TAG}_bench_plain.json 2> $OUT/${TAG}_bench_plain.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv $BENCH > $OUT/${TAG}_launches.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k "regex:lstm_(fwd2|bwd2|rec)_kernel" -c 4 -o $OUT/${TAG}_lstm_rec -f $BENCH > $OUT/${TAG}_ncu_rec.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_loss_fused -c 1 -o $OUT/${TAG}_loss -f $BENCH > $OUT/${TAG}_ncu_loss.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:gemm_tc_kernel -c 16 -o $OUT/${TAG}_gemm_tc -f $BENCH > $OUT/${TAG}_ncu_gemm.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:sampler_fused_kernel -c 1 -o $OUT/${TAG}_sampler -f python profiles/scripts/run_sampler.py 18944 128 > $OUT/${TAG}_ncu_sampler.log 2>&1
echo capture done
