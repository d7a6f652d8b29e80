"""ORACLE — test infrastructure only.  numpy restatement of the device noise draw `f5_randn_rows`
(tts_indic_server_f5_b200/csrc/elementwise.cu): Philox4x32-10 (Salmon et al., SC'11; the generator behind curand / torch's CUDA
`randn`, which the reference calls at f5_tts/model/cfm.py:186) keyed by a 64-bit utterance seed, counter
(frame * 32 + lane, 0x4635, 0, 0), Box-Muller on 24-bit uniforms; lane l yields channels 4l .. 4l+3."""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) for c in (c0, c1, c2, c3))
    k0, k1 = np.uint64(k0), np.uint64(k1)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = np.uint64(M0) * c0, np.uint64(M1) * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0, k1 = (k0 + np.uint64(W0)) & mask, (k1 + np.uint64(W1)) & mask
    return c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32)


def box_muller(a, b):
    u1 = ((a >> np.uint32(8)).astype(np.float32) + np.float32(1.0)) * np.float32(2.0 ** -24)
    u2 = (b >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)
    r = np.sqrt(np.float32(-2.0) * np.log(u1))
    ang = np.float32(2.0 * np.pi) * u2.astype(np.float64)          # sincospif(2 u2): evaluate the angle in double
    return (r * np.cos(ang)).astype(np.float32), (r * np.sin(ang)).astype(np.float32)


def randn_rows(seed: int, n_frames: int, channels: int = 100) -> np.ndarray:
    """[n_frames, channels] fp32: what the kernel writes for an utterance with this seed."""
    lanes = (channels + 3) // 4
    pos = np.repeat(np.arange(n_frames, dtype=np.uint64), lanes)
    lane = np.tile(np.arange(lanes, dtype=np.uint64), n_frames)
    seed &= 0xFFFFFFFFFFFFFFFF
    w = philox4x32_10((pos * np.uint64(32) + lane) & np.uint64(0xFFFFFFFF), np.full_like(pos, 0x4635), np.zeros_like(pos), np.zeros_like(pos),
                      seed & 0xFFFFFFFF, seed >> 32)
    a0, a1 = box_muller(w[0], w[1])
    b0, b1 = box_muller(w[2], w[3])
    out = np.stack([a0, a1, b0, b1], axis=1).reshape(n_frames, lanes * 4)
    return out[:, :channels]
