#!/bin/bash
# Round-2 GPU call 10 (4 GPUs): sharded C4 bench at N = 4 (weak), as the driver launches it.
mkdir -p gpurun_out/c10
O=gpurun_out/c10
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 3 --warmup 2 > $O/bench_4gpu_weak.json 2> $O/bench_4gpu_weak.err; echo "4gpu weak rc=$?" | tee -a $O/summary.txt
cut -c1-1600 $O/bench_4gpu_weak.json; tail -3 $O/bench_4gpu_weak.err | cut -c1-300
