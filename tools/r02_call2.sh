#!/bin/bash
# Round-2 GPU call 2: new tests, attention cost of the wait protocol variants, the reworked bench, allocation-stress soak.
mkdir -p gpurun_out/c2
O=gpurun_out/c2
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
for s in "" _nd _r1; do
  F5_LIB_SUFFIX=$s timeout 200 python tools/attn_bench.py >> $O/attn_bench.txt 2>&1
done
timeout 600 python bench.py --steps 3 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" | tee -a $O/summary.txt
F5_LIB_SUFFIX=_B timeout 200 python tools/soak2.py --seconds 75 --stress hostalloc >> $O/soak2.jsonl 2>> $O/soak2.err; echo "soak2 B hostalloc rc=$?" | tee -a $O/summary.txt
F5_LIB_SUFFIX=_A timeout 200 python tools/soak2.py --seconds 75 --stress hostalloc >> $O/soak2.jsonl 2>> $O/soak2.err; echo "soak2 A hostalloc rc=$?" | tee -a $O/summary.txt
tail -15 $O/pytest.log; cat $O/attn_bench.txt; cut -c1-3000 $O/bench.json; tail -5 $O/bench.err; cat $O/soak2.jsonl | cut -c1-600
