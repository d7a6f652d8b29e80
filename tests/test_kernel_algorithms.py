"""CPU restatements of the two index-heavy algorithms behind the Vocos-side memory kernels, lane by lane / slot by slot as the
CUDA code walks them, checked against numpy / torch references.  They pin the MATH the kernels implement (the factorisation of
the real-input inverse FFT in `csrc/vocos_istft.cu::istft_frames2_kernel`, the register-ring rotation and the key / position test
of `csrc/elementwise.cu::dwconv7_ln_run_kernel` and `dwconv7_ln_cs_kernel`); the kernels themselves are compared with
torch on the GPU in tests/test_kernels_gpu.py.  Replaces: torch.istft inside vocos 0.1.0 ISTFTHead (call site
f5_tts/infer/utils_infer.py:472) and the depthwise Conv1d(k=7, padding=3) of ConvNeXtV2Block (f5_tts/model/modules.py:262)."""
import numpy as np
import torch
import torch.nn.functional as F

W32 = np.exp(2j * np.pi * np.arange(16) / 32)


def _brev(i, bits):
    return int(format(i, f"0{bits}b")[::-1], 2)


def _idft_regs(v, n):
    """`idft_regs<NP>`: radix-2 decimation in frequency over a thread's registers; output index k ends up in v[bitrev(k)]."""
    v = list(v)
    for s in range(int(np.log2(n))):
        half = (n // 2) >> s
        for g in range(0, n, 2 * half):
            for j in range(half):
                a, b = v[g + j], v[g + j + half]
                v[g + j] = a + b
                v[g + j + half] = (a - b) * W32[j * (16 // half)]
    return v


def _irfft1024_like_the_kernel(X):
    """X: 513 bins.  One 512-point complex inverse FFT of Z[k] = (X[k] + conj X[512-k]) + i w^k (X[k] - conj X[512-k]) as
    32 lanes (k1) x 16 registers (k2), then 16 (n2) x 32 (k1) in the second stage; z[m] = x[2m] + i x[2m+1]."""
    X = X.copy()
    X[0], X[512] = X[0].real, X[512].real                    # irfft ignores the imaginary part of DC / Nyquist
    tile = np.zeros((32, 16), complex)
    for lane in range(32):                                    # stage 1: lane = k1
        own = [X[lane + 32 * k2] for k2 in range(16)] + [X[512] if lane == 0 else 0.0]
        src = (32 - lane) & 31
        Z = []
        for k2 in range(16):
            partner = own[16 - k2] if lane == 0 else X[src + 32 * (15 - k2)]     # the shuffle from lane 32 - k1
            A, B = own[k2], np.conj(partner)
            Z.append((A + B) + 1j * np.exp(2j * np.pi * (lane + 32 * k2) / 1024) * (A - B))
        Y = _idft_regs(Z, 16)
        for n2 in range(16):
            tile[lane][n2] = Y[_brev(n2, 4)] * np.exp(2j * np.pi * lane * n2 / 512)
    out = np.zeros(1024)
    for n2 in range(16):                                      # stage 2: lane = (frame, n2)
        z = _idft_regs([tile[k1][n2] for k1 in range(32)], 32)
        for n1 in range(32):
            val = z[_brev(n1, 5)] / 1024
            out[32 * n1 + 2 * n2], out[32 * n1 + 2 * n2 + 1] = val.real, val.imag
    return out


def test_real_input_fft_factorisation_equals_irfft():
    rng = np.random.default_rng(0)
    for _ in range(3):
        mag = np.minimum(np.exp(rng.normal(size=513) * 1.5 + 0.5), 100.0)
        ph = rng.normal(size=513) * 20
        X = mag * (np.cos(ph) + 1j * np.sin(ph))
        ref = np.fft.irfft(X, n=1024)
        got = _irfft1024_like_the_kernel(X)
        assert np.abs(got - ref).max() < 1e-12 * max(1.0, np.abs(ref).max())


def test_cody_waite_phase_reduction_keeps_the_phasor():
    """`polar_fast`: j = rint(p / 2 pi); r = fma(-j, 6.28125, p); r = fma(-j, 1.9353071795864769e-3, r) in fp32 lands in
    [-pi, pi] (+ rounding) with |sin r - sin p| < 3e-7 for |p| up to 100 rad (the MUFU adds ~5e-7)."""
    p = np.linspace(-100, 100, 200001).astype(np.float32)
    j = np.rint(p * np.float32(0.15915494309189535)).astype(np.float32)
    hi = (p.astype(np.float64) - j.astype(np.float64) * 6.28125).astype(np.float32)          # exact product, one rounding: an fma
    r = (hi.astype(np.float64) - j.astype(np.float64) * np.float64(np.float32(1.9353071795864769e-3))).astype(np.float32)
    assert np.abs(r).max() < np.pi + 1e-3
    assert np.abs(np.sin(r.astype(np.float64)) - np.sin(p.astype(np.float64))).max() < 3e-7
    assert np.abs(np.cos(r.astype(np.float64)) - np.cos(p.astype(np.float64))).max() < 3e-7


def _ragged(lens, gaps):
    pos, starts = [], []
    for g, n in zip(gaps, lens):
        pos += [-1] * g
        starts.append(len(pos))
        pos += list(range(n))
    pos += [-1] * gaps[-1]
    return np.array(pos), starts


DEAD = -2 ** 31


def _dwconv_ring(x, pos, w, bias, run_len, ring, keyed):
    """The kernels' control flow: a run of rows per warp (keyed=False: 7-slot ring, position test of form 2) or per CTA
    (keyed=True: 14-slot ring, key test of form 3); the slot of row r - 3 is refilled right after tap 0 has consumed it."""
    M, C = x.shape
    out = np.full((M, C), 7.0)
    for r0 in range(0, M, run_len):
        r1 = min(r0 + run_len, M)
        win = [np.zeros(C) for _ in range(ring)]
        tag = [DEAD if keyed else -1] * ring
        nxt = [r0 - 3]

        def load(slot):
            rr = nxt[0]
            tag[slot] = DEAD if keyed else -1
            if 0 <= rr < M:
                win[slot] = x[rr].copy()                      # gap rows are loaded too; their tag keeps them out
                p = pos[rr]
                tag[slot] = (p - rr if p >= 0 else DEAD) if keyed else p
            nxt[0] += 1

        for s in range(ring):
            load(s)
        base = r0
        while base < r1:
            for j in range(ring):
                row = base + j
                if row >= r1:
                    continue
                c = tag[(j + 3) % ring]
                if c == (DEAD if keyed else -1) or (not keyed and c < 0):
                    out[row] = 0
                    load(j % ring)
                    continue
                acc = bias.copy()
                for k in range(7):
                    sl = (j + k) % ring
                    ok = (tag[sl] == c) if keyed else (tag[sl] >= 0 and tag[sl] == c + k - 3)
                    if ok:
                        acc = acc + win[sl] * w[:, k]
                    if k == 0:
                        load(j % ring)
                out[row] = acc
            base += ring
    return out


def test_dwconv_register_ring_equals_conv1d_on_a_ragged_pack():
    rng = np.random.default_rng(1)
    lens, gaps = [1, 37, 2, 70, 3, 150, 64, 5, 33, 409], [0, 1, 2, 1, 8, 3, 1, 16, 2, 1, 0]
    pos, starts = _ragged(lens, gaps)
    M, C = len(pos), 8
    x = rng.normal(size=(M, C))
    x[pos < 0] = np.nan                                       # a dead row that leaks shows up as NaN
    w, bias = rng.normal(size=(C, 7)), rng.normal(size=C)
    ref = np.zeros((M, C))
    for s0, n in zip(starts, lens):
        ref[s0:s0 + n] = F.conv1d(torch.tensor(x[s0:s0 + n].T[None]), torch.tensor(w[:, None, :]), torch.tensor(bias),
                                  padding=3, groups=C)[0].T.numpy()
    live = pos >= 0
    for ring, keyed, runs in ((7, False, (7, 14, 49, 1400)), (14, True, (14, 28, 70, 1400))):
        for run in runs:
            got = _dwconv_ring(x, pos, w, bias, run, ring, keyed)
            assert np.isfinite(got).all()
            assert np.abs(got[live] - ref[live]).max() < 1e-12 and np.abs(got[~live]).max() == 0


def test_grouped_layernorm_statistics_combine_exactly():
    """Form 3 of dwconv7+LN: per-warp (mean, M2) over 128 channels combined as M2 = sum M2_g + n (mean_g - mean)^2."""
    rng = np.random.default_rng(2)
    v = rng.normal(size=512) * 3 + 1.5
    g = v.reshape(4, 128)
    mg, m2g = g.mean(1), ((g - g.mean(1, keepdims=True)) ** 2).sum(1)
    mean = mg.mean()
    m2 = (m2g + 128 * (mg - mean) ** 2).sum()
    assert abs(mean - v.mean()) < 1e-12 and abs(m2 / 512 - v.var()) < 1e-12
