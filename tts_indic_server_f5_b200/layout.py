"""Packed variable-length row layout shared by all kernels of one batch.

All utterances of a batch live in ONE row-major activation matrix (no per-utterance padding to a common length, so
no padding FLOPs): rows `[start_i, start_i + n_i)` hold utterance i, separated by `GAP` zero rows so that the k=31
convolution's halo (15 rows each side; `model/modules.py:171-176`) reads zeros exactly like the reference's
`padding=15`, and the TMA-fed implicit-conv GEMM needs no per-utterance special case.  The CFG pair is batched into
one pass: rows `[0, R)` are the conditional branch, rows `[R, 2R)` the unconditional one (`model/cfm.py:162-176`).
Per-utterance (batch-1) semantics are preserved: attention tiles never cross utterances, GRN reduces over each
utterance's own rows, convolutions see zeros outside the utterance.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

GAP = 16
ROW_ALIGN = 128


@dataclass
class PackedLayout:
    lengths: list[int]          # n_i (frames per utterance)
    starts: list[int]           # first row of utterance i inside one CFG half
    half_rows: int              # R, multiple of 128
    row_pos: torch.Tensor       # int32 [2R]: position inside the utterance, -1 on gap rows (both halves)
    row_utt: torch.Tensor       # int32 [R]: utterance index, -1 on gap rows
    attn_tiles: torch.Tensor    # int32 [T,4]: q_row0, kv_row0, kv_len, q_rows_valid (<= 256: a PAIR of 128-row tiles), both halves
    seg_rows: torch.Tensor      # int32 [2B,2]: row0, rows (both halves; GRN segments)

    @property
    def rows(self) -> int:
        return 2 * self.half_rows

    @property
    def real_tokens(self) -> int:
        return sum(self.lengths)

    def signature(self) -> tuple:
        return tuple(self.lengths)


def build_layout(lengths: list[int], gap: int = GAP, both_halves: bool = True) -> PackedLayout:
    import numpy as np
    lens = np.asarray(lengths, dtype=np.int64)
    assert lens.ndim == 1 and lens.size >= 1 and (lens >= 1).all()
    starts_np = gap + np.concatenate([[0], np.cumsum(lens[:-1] + gap)])
    r = int(starts_np[-1] + lens[-1] + gap)
    R = (r + ROW_ALIGN - 1) // ROW_ALIGN * ROW_ALIGN
    pos_np = np.full(R, -1, dtype=np.int32)
    utt_np = np.full(R, -1, dtype=np.int32)
    rows = np.repeat(starts_np, lens) + (np.arange(int(lens.sum())) - np.repeat(np.cumsum(lens) - lens, lens))
    pos_np[rows] = (rows - np.repeat(starts_np, lens)).astype(np.int32)
    utt_np[rows] = np.repeat(np.arange(lens.size), lens).astype(np.int32)
    starts = [int(x) for x in starts_np]
    halves = (0, R) if both_halves else (0,)
    tiles, segs = [], []
    for off in halves:
        for s, n in zip(starts, lengths):
            segs.append([off + s, n])
            for q0 in range(0, n, 256):
                tiles.append([off + s + q0, off + s, n, min(256, n - q0)])
    # longest-first ordering keeps the tail of the attention grid short
    tiles.sort(key=lambda t: -t[2])
    pos = torch.from_numpy(pos_np)
    return PackedLayout(list(lengths), starts, R, torch.cat([pos] * len(halves)), torch.from_numpy(utt_np),
                        torch.tensor(tiles, dtype=torch.int32), torch.tensor(segs, dtype=torch.int32))
