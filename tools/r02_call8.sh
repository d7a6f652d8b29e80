#!/bin/bash
# Round-2 GPU call 8: wave-cost tile rule (tests + B = 1 latency breakdown) and the round's ncu evidence.
mkdir -p gpurun_out/c8
O=gpurun_out/c8
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
timeout 300 python tools/latency_c1.py > $O/latency.txt 2>&1
timeout 300 python tools/ncu_kernels.py > $O/ncu_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -f -o $O/prof_r02_layer -k regex:"gemm_tcgen05|attn_d64|layernorm_mod" -c 6 \
    python tools/ncu_kernels.py > $O/ncu_layer.log 2>&1; echo "ncu layer rc=$?" | tee -a $O/summary.txt
timeout 400 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $O/bench_plain.json 2> $O/bench_plain.err && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 480 --csv --log-file $O/launches_r02.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $O/ncu_bench.log 2>&1; echo "ncu launches rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest.log | cut -c1-200; cat $O/latency.txt; tail -2 $O/ncu_layer.log; wc -l $O/launches_r02.csv; ls -la $O
