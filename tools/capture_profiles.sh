#!/bin/bash
# Capture the round's ncu evidence on the GPU box (run under gpurun, one GPU).  Outputs land in gpurun_out/:
#   launches_<tag>.csv       ncu launch list (gpu__time_duration.sum) of `python bench.py --steps 1 --warmup 3 --no-cpu-baseline`
#   prof_<tag>_layer.ncu-rep ncu --set full of one DiT layer's kernels at C2 scale (tools/ncu_kernels.py)
# Each command first runs WITHOUT ncu and must exit 0.
TAG=${1:-r02}
mkdir -p gpurun_out
set -x
timeout 300 python tools/ncu_kernels.py > gpurun_out/ncu_plain_$TAG.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -f -o gpurun_out/prof_${TAG}_layer -k regex:"gemm_tcgen05|attn_d64" -c 5 \
    python tools/ncu_kernels.py > gpurun_out/ncu_layer_$TAG.log 2>&1
timeout 400 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plain_$TAG.json 2> gpurun_out/bench_plain_$TAG.err || exit 1
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 480 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_$TAG.log 2>&1
tail -3 gpurun_out/ncu_layer_$TAG.log
ls -la gpurun_out/*$TAG*
