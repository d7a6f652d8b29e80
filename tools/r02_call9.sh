#!/bin/bash
# Round-2 GPU call 9: compute-sanitizer memcheck over the end-to-end smoke (every kernel of the path at tiny size) — ONE tool per call.
mkdir -p gpurun_out/c9
O=gpurun_out/c9
timeout 300 python __graft_entry__.py --smoke > $O/smoke_plain.log 2>&1; echo "smoke plain rc=$?" | tee -a $O/summary.txt
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $O/memcheck_smoke.txt python __graft_entry__.py --smoke > $O/memcheck_smoke.out 2>&1
echo "memcheck smoke rc=$?" | tee -a $O/summary.txt
tail -5 $O/memcheck_smoke.txt; tail -3 $O/memcheck_smoke.out
