#!/bin/bash
# Round-2 GPU call 12 (8 GPUs): the C4 batch as BASELINE.json configs[3] names it — 512 utterances sharded over 8 GPUs — as the driver launches it.
mkdir -p gpurun_out/c12
O=gpurun_out/c12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 3 --warmup 2 > $O/bench_8gpu.json 2> $O/bench_8gpu.err; echo "8gpu rc=$?" | tee -a $O/summary.txt
cut -c1-1200 $O/bench_8gpu.json; tail -3 $O/bench_8gpu.err | cut -c1-300
