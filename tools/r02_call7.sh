#!/bin/bash
# Round-2 GPU call 7 (2 GPUs): the sharded multi-GPU bench path (weak + strong) and the reference arm.
mkdir -p gpurun_out/c7
O=gpurun_out/c7
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 2 > $O/bench_2gpu_weak.json 2> $O/bench_2gpu_weak.err; echo "2gpu weak rc=$?" | tee -a $O/summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 1 --warmup 1 --scaling strong > $O/bench_2gpu_strong.json 2> $O/bench_2gpu_strong.err; echo "2gpu strong rc=$?" | tee -a $O/summary.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference rc=$?" | tee -a $O/summary.txt
for f in bench_2gpu_weak bench_2gpu_strong bench_reference; do echo "== $f"; cut -c1-1800 $O/$f.json; tail -4 $O/$f.err | cut -c1-300; done
