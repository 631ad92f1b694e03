#!/bin/bash
# Capture the ncu evidence for one round (run under gpurun, one GPU):  bash profiles/scripts/capture.sh r02
# 1. launch list of the bench command (gpu__time_duration.sum per launch; the kernel SHARES of the step are what matter)
# 2. `--set full` captures: the cluster recurrence kernels, ALL tcgen05 GEMM launches of one training step (gemm_tc_kernel
#    incl. the fused fc_out + cross-entropy instantiation, gemm_ws_kernel), the stand-alone fused loss kernel, the sampler
# ncu serialises kernels, so the serial launch order is used (ARCVAE_ONE_STREAM=1).
set -u
TAG=${1:-r02}
OUT=gpurun_out
export ARCVAE_ONE_STREAM=1
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu --no-sampler --no-extras"
timeout 300 $BENCH > $OUT/${TAG}_bench_plain.json 2> $OUT/${TAG}_bench_plain.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/${TAG}_launches.csv $BENCH > $OUT/${TAG}_launches.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k "regex:lstm_(fwd3|bwd3)_kernel" -c 4 -o $OUT/${TAG}_lstm_rec -f $BENCH > $OUT/${TAG}_ncu_rec.log 2>&1
# one whole step of tensor-core GEMMs: 20 gemm_tc_kernel + 4 gemm_ws_kernel launches per step; skip the 3 warm-up steps
timeout 900 ncu --set full --clock-control none -k "regex:gemm_(tc|ws)_kernel" --launch-skip 72 -c 24 -o $OUT/${TAG}_gemm -f $BENCH > $OUT/${TAG}_ncu_gemm.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_loss_fused -c 1 -o $OUT/${TAG}_loss -f python profiles/scripts/run_loss.py > $OUT/${TAG}_ncu_loss.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:sampler_fused_kernel -c 1 -o $OUT/${TAG}_sampler -f python profiles/scripts/run_sampler.py 18944 128 > $OUT/${TAG}_ncu_sampler.log 2>&1
# summaries on the box (gpurun brings back at most 64 MiB: the 24-launch GEMM report alone is ~48 MB and stays there)
for k in lstm_rec gemm loss sampler; do python profiles/scripts/summarize_ncu.py $OUT/${TAG}_$k.ncu-rep > $OUT/${TAG}_${k}_ncu_summary.txt 2>/dev/null; done
python profiles/scripts/summarize_ncu.py --traffic $OUT/${TAG}_lstm_rec.ncu-rep $OUT/${TAG}_gemm.ncu-rep $OUT/${TAG}_loss.ncu-rep > $OUT/${TAG}_ncu_traffic.json 2>/dev/null
python profiles/summarize_launches.py $OUT/${TAG}_launches.csv > $OUT/${TAG}_launches_summary.txt 2>/dev/null
rm -f $OUT/${TAG}_gemm.ncu-rep
echo capture done
