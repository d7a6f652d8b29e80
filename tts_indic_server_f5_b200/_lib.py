"""ctypes binding of the C ABI declared in `include/f5_b200.h` (no torch types cross this boundary).

The library must exist: there is NO CPU / PyTorch fallback for any op on the hot path.  Import fails loudly
when `libf5b200.so` is missing, and every call raises on a non-zero return code.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from .build import LIB

F5_EPI_STORE_BF16, F5_EPI_STORE_F32, F5_EPI_RESID_F32 = 0, 1, 2
F5_ACT_NONE, F5_ACT_GELU_TANH, F5_ACT_GELU_ERF, F5_ACT_MISH = 0, 1, 2, 3

EXPORTS = [
    "f5_gemm_bf16", "f5_attention_d64", "f5_attention_f32", "f5_grn_sumsq_f32", "f5_grn_apply_f32", "f5_layernorm_mod", "f5_dwconv7_ln", "f5_grn_sumsq", "f5_grn_apply",
    "f5_text_gather_pos", "f5_pack_bf16", "f5_where_rows", "f5_cfg_euler", "f5_time_sinus", "f5_silu_bf16",
    "f5_istft_frames", "f5_istft_ola", "f5_mel_frames", "f5_randn_rows", "f5_diag_enable", "f5_set_pdl", "f5_set_dwconv7_variant", "f5_set_istft_variant",
    "f5_device_check", "f5_version",
]


class GemmArgs(C.Structure):
    """Mirror of `f5_gemm_args` (include/f5_b200.h)."""
    _fields_ = [
        ("A", C.c_void_p), ("B", C.c_void_p), ("lda", C.c_int64), ("ldb", C.c_int64),
        ("a_rows", C.c_int32), ("a_cols", C.c_int32), ("b_rows", C.c_int32), ("b_cols", C.c_int32),
        ("M", C.c_int32), ("N", C.c_int32), ("block_n", C.c_int32),
        ("num_taps", C.c_int32), ("kc_per_tap", C.c_int32), ("tap_pad", C.c_int32), ("a_grouped", C.c_int32),
        ("b_tap_rows", C.c_int32),
        ("mode", C.c_int32), ("act", C.c_int32),
        ("bias", C.c_void_p), ("gate", C.c_void_p),
        ("out", C.c_void_p), ("ldo", C.c_int64), ("out2", C.c_void_p), ("ldo2", C.c_int64),
        ("addend", C.c_void_p), ("ld_add", C.c_int64), ("resid", C.c_void_p), ("ldr", C.c_int64),
        ("row_pos", C.c_void_p), ("mask_rows", C.c_int32),
        ("rope", C.c_void_p), ("rope_period", C.c_int32), ("rope_tiles", C.c_int32), ("num_sms", C.c_int32),
        ("taps_per_seg", C.c_int32), ("a_lo_off", C.c_int32),
    ]


class F5Error(RuntimeError):
    pass


def _load() -> C.CDLL:
    if not os.path.exists(LIB):
        raise ImportError(f"{LIB} is missing: run `python -m tts_indic_server_f5_b200.build` (nvcc, sm_100a). "
                          "There is no CPU fallback for the F5 hot path.")
    lib = C.CDLL(LIB)
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    sig = {
        "f5_gemm_bf16": [C.POINTER(GemmArgs), vp],
        "f5_attention_d64": [vp, i64, i32, i32, i32, i32, i32, vp, i32, vp, i64, f32, vp],
        "f5_attention_f32": [vp, i64, i32, i32, i32, i32, vp, i32, vp, vp, i64, i32, f32, vp],
        "f5_layernorm_mod": [vp, i64, vp, i64, vp, i64, i32, i32, vp, vp, f32, f32, i32, vp],
        "f5_dwconv7_ln": [vp, i64, vp, i64, i32, i32, vp, vp, vp, vp, vp, f32, i32, vp],
        "f5_grn_sumsq_f32": [vp, i64, i32, vp, i32, vp, vp],
        "f5_grn_apply_f32": [vp, i64, i32, vp, i32, vp, vp, vp, vp],
        "f5_grn_sumsq": [vp, i64, i32, vp, i32, vp, vp],
        "f5_grn_apply": [vp, i64, i32, vp, i32, vp, vp, vp, vp],
        "f5_text_gather_pos": [vp, vp, vp, vp, i32, vp, i64, i32, i32, vp],
        "f5_pack_bf16": [vp, i64, vp, i64, i32, i32, i32, i32, vp, vp, i32, vp],
        "f5_where_rows": [vp, i64, vp, i64, vp, i32, i32, vp],
        "f5_cfg_euler": [vp, i64, vp, i64, i32, i32, vp, vp, i32, f32, vp, i64, i32, vp],
        "f5_time_sinus": [vp, i32, vp, i32, vp, i64, i32, vp],
        "f5_silu_bf16": [vp, vp, i64, i32, vp],
        "f5_istft_frames": [vp, i64, i32, vp, vp, vp],
        "f5_istft_ola": [vp, vp, vp, i32, i32, vp, vp, vp],
        "f5_mel_frames": [vp, vp, i32, i32, vp, vp, vp, i32, vp, i64, vp],
        "f5_randn_rows": [vp, i64, i32, i32, vp, vp, vp, vp],
        "f5_diag_enable": [vp],
        "f5_set_pdl": [i32],
        "f5_set_dwconv7_variant": [i32],
        "f5_set_istft_variant": [i32],
        "f5_device_check": [],
    }
    for name, args in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    lib.f5_version.restype = C.c_char_p
    lib.f5_version.argtypes = []
    return lib


lib = _load()
launch_count = 0  # kernels launched through this binding (bench.py reports it as gpu_launches)


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = {-1: "bad argument", -2: "cuTensorMapEncodeTiled unavailable/failed", -3: "device is not sm_100"}.get(rc)
        if msg is None:
            msg = f"cudaError {rc}"
        raise F5Error(f"{what}: {msg}")


def ptr(t: torch.Tensor | None) -> int | None:
    if t is None:
        return None
    assert t.is_cuda, "device tensor expected"
    return t.data_ptr()


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def call(name: str, *args) -> None:
    global launch_count
    launch_count += 1
    check(getattr(lib, name)(*args), name)


# ------------------------------------------------------------------------------------------ device fault record
_diag_buf: torch.Tensor | None = None


def enable_diag() -> torch.Tensor:
    """Register a pinned host buffer the mbarrier watchdog writes its record into before it traps (`f5_diag_enable`).
    Host memory outlives the CUDA context, so `read_diag()` still works after a launch failure made the context unusable."""
    global _diag_buf
    if _diag_buf is None:
        _diag_buf = torch.zeros(64, dtype=torch.int64).pin_memory()
        check(lib.f5_diag_enable(_diag_buf.data_ptr()), "f5_diag_enable")
    return _diag_buf


def read_diag() -> dict | None:
    """Decode the watchdog record (layout: csrc/f5_common.cuh).  None: no kernel of this library trapped on its watchdog, so a
    launch failure seen by the host was something else (memory fault, Xid)."""
    if _diag_buf is None:
        return None
    w = _diag_buf.numpy().view("uint32")
    if (int(w[0]) & 0xFFF00000) != 0xF5D00000:
        return None
    block_dim = int(w[2]) >> 16
    kernel = {1: f"gemm_tcgen05_kernel ({'8' if block_dim == 384 else '4'} epilogue warps)", 2: "attn_d64_kernel"}.get(int(w[0]) & 0xFFFFF, "?")
    bar = int(w[3]) & 0x7FFFFFFF
    rec = {"kernel": kernel, "block_dim": block_dim, "grid_dim": int(w[1]) >> 16, "block": int(w[1]) & 0xFFFF,
           "thread": int(w[2]) & 0xFFFF, "warp": (int(w[2]) & 0xFFFF) // 32, "barrier_smem_addr": bar,
           "barrier_slot": (bar & 255) // 8, "parity": int(w[3]) >> 31}
    q = _diag_buf.numpy().view("uint64")
    if int(q[2]) != 0:
        rec["waited_ns"] = int(q[3])
        rec["barrier_words"] = [hex(int(x)) for x in q[8:40]]
    return rec
