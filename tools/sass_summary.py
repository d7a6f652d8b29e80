#!/usr/bin/env python
"""Per-kernel SASS evidence of the in-tree library: instruction count and the Blackwell-native mnemonics (UTCHMMA =
tcgen05.mma, UTMALDG / UTMAREDG = TMA load / reduce, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, SYNCS = mbarrier,
MUFU), plus registers / spills from `cuobjdump -res-usage`.  Runs without a GPU:  python tools/sass_summary.py [lib.so]"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "tts_indic_server_f5_b200/libf5b200.so"
KEYS = ["UTCHMMA", "UTMALDG", "UTMAREDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "SYNCS", "MUFU", "R2UR", "HMMA"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()  # noqa: E731
usage = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
    m = re.search(r"REG:(\d+).*?SHARED:(\d+).*?LOCAL:(\d+)", line)
    if m and cur:
        usage[cur] = (int(m.group(1)), int(m.group(2)), int(m.group(3)))
counts, total = collections.defaultdict(collections.Counter), collections.Counter()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        total[cur] += 1
        op = m.group(1)
        for k in KEYS:
            if op.startswith(k):
                counts[cur][k] += 1
print(f"# {lib}: sm_100a SASS summary (instructions, registers, local bytes = spills, Blackwell mnemonics)")
for fn in sorted(total, key=lambda f: -total[f]):
    reg, sh, loc = usage.get(fn, (0, 0, 0))
    name = demangle(fn)
    name = re.sub(r"\(.*", "", name)[:90]
    ks = " ".join(f"{k}={counts[fn][k]}" for k in KEYS if counts[fn][k])
    print(f"{total[fn]:6d} instr  reg={reg:3d} local={loc:4d}  {name}  {ks}")
