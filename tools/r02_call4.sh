#!/bin/bash
# Round-2 GPU call 4: fp32-mode tests, fp32-mode bench line, ncu launch list of a B = 1 request.
mkdir -p gpurun_out/c4
O=gpurun_out/c4
timeout 900 python -m pytest tests -m gpu -x -q -s > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -s -k "fp32 or manager or checkpoint_file or rejects or programmatic" > $O/pytest_new.log 2>&1; echo "pytest new rc=$?" | tee -a $O/summary.txt
timeout 600 python bench.py --precision fp32 --steps 2 --warmup 1 --no-extras --no-cpu-baseline > $O/bench_fp32.json 2> $O/bench_fp32.err; echo "bench fp32 rc=$?" | tee -a $O/summary.txt
timeout 300 python tools/c1_once.py 3 > $O/c1_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_c1.csv python tools/c1_once.py 3 > $O/ncu_c1.log 2>&1; echo "ncu rc=$?" | tee -a $O/summary.txt
grep -E "passed|failed|error" $O/pytest.log | tail -3; grep -E "fp32|rel-L2|SNR|split|attention_f32|passed|failed|Error|assert" $O/pytest.log $O/pytest_new.log | cut -c1-250 | tail -40
cut -c1-1500 $O/bench_fp32.json; tail -3 $O/bench_fp32.err; tail -2 $O/ncu_c1.log; wc -l $O/launches_c1.csv
