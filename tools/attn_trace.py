"""clock64 event trace of CTA 0 of the attention kernel.  Needs the trace build of the library:
  F5_LIB_SUFFIX=_trace F5_NVCC_EXTRA=-DATT_TRACE=1 python -m tts_indic_server_f5_b200.build --force   (then run this under gpurun)"""
import ctypes, os, sys
os.environ["F5_ATTN_TRACE"] = "1"
os.environ.setdefault("F5_LIB_SUFFIX", "_trace")
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tts_indic_server_f5_b200 import ops, _lib
from tts_indic_server_f5_b200.layout import build_layout
dev = "cuda"; torch.manual_seed(0)
g = torch.Generator().manual_seed(0)
lens = [469 + int(torch.randint(560, 941, (1,), generator=g)) for _ in range(64)]
L = build_layout(lens); D = 1024
qkv = torch.randn(L.rows, 3 * D, device=dev).to(torch.bfloat16)
ab = torch.zeros(L.rows, D, device=dev, dtype=torch.bfloat16)
tiles = L.attn_tiles.to(dev)
for _ in range(3):
    ops.attention(qkv, tiles, ab, 16, 0, D, 2 * D, 0.125)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 2048)()
_lib.lib.f5_attention_trace_dump.argtypes = [ctypes.c_void_p]
print("rc", _lib.lib.f5_attention_trace_dump(buf))
a = np.array(buf[:], dtype=np.int64).reshape(4, 512)
t0 = a[a > 0].min()
prod = a[0][a[0] > 0] - t0
print("producer kv-load issue times (first 24):", prod[:24].tolist())
m = a[1][a[1] > 0] - t0
print("MMA events (per step: top, next QKs issued, PVs issued) first 45:", m[:45].tolist())
for r in (2, 3):
    s = a[r][:512].reshape(-1, 8)[:, :6]
    s = s[(s > 0).all(axis=1)] - t0
    print(f"softmax group {r-2}: per tile [s_full seen, +S in regs, +max/grow/gate, +exp chunks 0-1 and o_full wait/read-out/rescale, +exp chunks 2-3, +st wait & p_full arrive | gap to next s_full] first 14:")
    for k, row in enumerate(s[:14]):
        nxt = int(s[k + 1][0] - row[5]) if k + 1 < len(s) else -1
        print("   ", int(row[0]), np.diff(row).tolist(), "| gap", nxt, " total", int(row[5] - row[0]))
    d = np.diff(s[:, 0])
    print("   tile period (s_full to s_full):", d[:16].tolist())
