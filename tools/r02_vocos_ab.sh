#!/bin/bash
# Round-2 GPU call 15 (the last 5 GPU-minutes): the new variants of the two Vocos-side memory kernels.
# Priority order: full GPU suite (new defaults + A/B tests), same-box kernel A/B, short bench, C5 sweep, smoke.
mkdir -p gpurun_out/c15
O=gpurun_out/c15
timeout 200 python -m pytest tests -m gpu -q -rP > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
timeout 90 python tools/vocos_kernels_ab.py $O/vocos_ab.json > $O/vocos_ab.txt 2> $O/vocos_ab.err; echo "ab rc=$?" | tee -a $O/summary.txt
timeout 150 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" | tee -a $O/summary.txt
timeout 120 python tools/vocos_sweep.py $O/vocos_sweep.json > $O/vocos_sweep.txt 2> $O/vocos_sweep.err; echo "sweep rc=$?" | tee -a $O/summary.txt
timeout 100 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt
grep -h "fp32 forward\|dwconv7_ln v\|dwconv7_ln planes\|istft_frames v\|istft T=" $O/pytest.log | cut -c1-200 | head -60; grep -h "passed\|failed\|FAILED\|Error" $O/pytest.log | tail -12 | cut -c1-300; cat $O/vocos_ab.txt; cut -c1-300 $O/bench.json; tail -3 $O/bench.err | cut -c1-300; tail -8 $O/vocos_sweep.txt; tail -3 $O/smoke.log | cut -c1-200
