// Vocos ISTFT head (vocos 0.1.0 `ISTFTHead`, padding="center"; call site infer/utils_infer.py:472), sm_100a.
//   istft_frames: per frame  mag = min(exp(m), 1e2); X = mag (cos p + i sin p); x = irfft_1024(X) * window     (fp32, smem FFT)
//   istft_ola   : overlap-add with hop 256, divide by the squared-window envelope, trim n_fft/2 each side, optional gain
// The inverse real FFT is a 1024-point complex FFT of the Hermitian-extended spectrum, one warp per frame, in registers
// (32 x 32 decomposition, one shared-memory transpose); the stage moves 4104 B in + 4096 B out per frame.
#include "f5_common.cuh"
#include "../../include/f5_b200.h"
#include <cstdlib>

namespace f5 {

constexpr int NFFT = 1024;
constexpr int NBINS = NFFT / 2 + 1;
constexpr int HOP = 256;

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// e^{+2 pi i m / 32}, m < 16 (constant bank: compile-time indices after unrolling cost nothing)
__constant__ float2 kW32[16] = {
    {1.f, 0.f}, {0.98078528040323043f, 0.19509032201612825f}, {0.92387953251128674f, 0.38268343236508978f},
    {0.83146961230254524f, 0.55557023301960218f}, {0.70710678118654757f, 0.70710678118654757f},
    {0.55557023301960229f, 0.83146961230254524f}, {0.38268343236508984f, 0.92387953251128674f},
    {0.19509032201612833f, 0.98078528040323043f}, {0.f, 1.f}, {-0.19509032201612819f, 0.98078528040323043f},
    {-0.38268343236508973f, 0.92387953251128674f}, {-0.55557023301960196f, 0.83146961230254546f},
    {-0.70710678118654746f, 0.70710678118654757f}, {-0.83146961230254535f, 0.55557023301960218f},
    {-0.92387953251128674f, 0.38268343236508989f}, {-0.98078528040323043f, 0.19509032201612861f}};

__host__ __device__ constexpr int brev5(int i) {
  return ((i & 1) << 4) | ((i & 2) << 2) | (i & 4) | ((i & 8) >> 2) | ((i & 16) >> 4);
}

// 32-point inverse DFT (kernel e^{+2 pi i k n / 32}) of the values a thread holds in registers: radix-2 decimation in
// frequency, fully unrolled; output index n ends up in v[brev5(n)].
__device__ __forceinline__ void idft32_regs(float2 (&v)[32]) {
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    const int half = 16 >> s;
#pragma unroll
    for (int g = 0; g < 32; g += 2 * half) {
#pragma unroll
      for (int j = 0; j < half; ++j) {
        const float2 a = v[g + j], b = v[g + j + half];
        v[g + j] = make_float2(a.x + b.x, a.y + b.y);
        const float2 d = make_float2(a.x - b.x, a.y - b.y);
        const int m = j * (16 / half);                       // twiddle e^{+2 pi i j / (2 half)} = W32^m
        if (m == 0) v[g + j + half] = d;
        else if (m == 8) v[g + j + half] = make_float2(-d.y, d.x);
        else v[g + j + half] = cmul(d, kW32[m]);
      }
    }
  }
}

// One WARP per frame, the whole 1024-point inverse FFT in registers as 32 x 32 (k = k1 + 32 k2, n = 32 n1 + n2):
//   lane k1 : builds X[k1 + 32 k2] (mag = min(exp(m), 1e2), phase -> cos/sin; the upper half of the spectrum is the conjugate
//             of a bin another lane built: one shuffle), 32-point inverse DFT over k2 in registers, twiddle e^{2 pi i k1 n2 / 1024}
//   transpose through a per-warp shared-memory tile (conflict-free, 33-float rows)
//   lane n2 : 32-point inverse DFT over k1 in registers, x[32 n1 + n2] = Re(.) / 1024 * window -> coalesced 128-B stores.
// ~2.8 k instructions per lane per frame and no block-wide barrier: the first version (one CTA per frame, ten radix-2
// stages in shared memory, a __syncthreads each, 32-way bank conflicts on the bit-reversed scatter and on the twiddle
// table) ran at 5 % of the HBM bandwidth its 8.2 KB per frame calls for.
constexpr int ISTFT_WARPS = 4;
__global__ void __launch_bounds__(ISTFT_WARPS * 32) istft_frames_kernel(const float* __restrict__ spec, long long lds, int rows,
                                                                       const float* __restrict__ window, float* __restrict__ frames) {
  pdl_wait();
  pdl_launch();
  __shared__ float tre[ISTFT_WARPS][32][33];
  __shared__ float tim[ISTFT_WARPS][32][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float (*re)[33] = tre[warp];
  float (*im)[33] = tim[warp];
  // per-lane step-2 twiddle seeds e^{2 pi i lane n2 / 1024} for n2 = 0, 8, 16, 24 (exact), the rest by recurrence with w1
  float2 w1, wseed[4];
  sincospif(static_cast<float>(lane) / 512.f, &w1.y, &w1.x);
#pragma unroll
  for (int q = 0; q < 4; ++q) sincospif(static_cast<float>(lane * q * 8) / 512.f, &wseed[q].y, &wseed[q].x);
  const int src_lane = (32 - lane) & 31;
  for (int f = blockIdx.x * ISTFT_WARPS + warp; f < rows; f += gridDim.x * ISTFT_WARPS) {
    const float* sp = spec + static_cast<size_t>(f) * lds;
    float2 v[32];
#pragma unroll
    for (int k2 = 0; k2 <= 16; ++k2) {
      const int k = lane + 32 * k2;
      float2 X = make_float2(0.f, 0.f);
      if (k <= NFFT / 2) {
        const float mag = fminf(expf(sp[k]), 100.f);
        float s, c;
        sincosf(sp[NBINS + k], &s, &c);
        X = make_float2(mag * c, (k == 0 || k == NFFT / 2) ? 0.f : mag * s);   // irfft ignores Im of DC / Nyquist
      }
      v[k2] = X;
    }
    // X[k1 + 32 k2], k2 = 16..31 (t = 32 - k2): conj X[32 t - k1] = bin (32 - k1) + 32 (t - 1) of lane 32 - k1; lane 0 owns X[32 t]
#pragma unroll
    for (int t = 16; t >= 1; --t) {
      const float ox = __shfl_sync(0xffffffffu, v[t - 1].x, src_lane);
      const float oy = __shfl_sync(0xffffffffu, v[t - 1].y, src_lane);
      const float2 own = v[t];                               // lane 0: X[32 t] (t = 16: the Nyquist bin, already real)
      v[32 - t] = lane == 0 ? make_float2(own.x, -own.y) : make_float2(ox, -oy);
    }
    idft32_regs(v);                                          // over k2: result for n2 in v[brev5(n2)]
    __syncwarp();                                            // previous frame's tile reads are done
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float2 w = wseed[q];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int n2 = q * 8 + r;
        const float2 bv = cmul(v[brev5(n2)], w);
        re[lane][n2] = bv.x;
        im[lane][n2] = bv.y;
        w = cmul(w, w1);
      }
    }
    __syncwarp();
#pragma unroll
    for (int k1 = 0; k1 < 32; ++k1) v[k1] = make_float2(re[k1][lane], im[k1][lane]);
    idft32_regs(v);                                          // over k1: x[32 n1 + lane] in v[brev5(n1)]
    float* o = frames + static_cast<size_t>(f) * NFFT;
#pragma unroll
    for (int n1 = 0; n1 < 32; ++n1) o[32 * n1 + lane] = v[brev5(n1)].x * (1.f / NFFT) * window[32 * n1 + lane];
  }
}

// ---------------------------------------------------------------------------------------------- v2: real-input form
// The same transform with half the butterflies: x is real, so the 1024-point inverse real FFT is ONE 512-point complex
// inverse FFT of  Z[k] = (X[k] + conj X[512-k]) + i w^k (X[k] - conj X[512-k]),  w = e^{2 pi i / 1024},  k < 512,  with
// x[2m] = Re z[m], x[2m+1] = Im z[m].  512 = 32 (k1, lanes) x 16 (k2, registers):
//   step 1, one frame at a time, lane = k1: X[k1 + 32 k2] from (log-magnitude, phase) with ex2 / sin / cos on the MUFU after an
//     explicit Cody-Waite reduction of the phase to [-pi, pi] (abs. error 5e-7 on a unit phasor; sincosf's slow path and expf
//     were a third of the first version's 4.8 k instructions), the partner bin X[512-k] by one shuffle from lane 32 - k1,
//     16-point inverse DFT over k2 in registers, twiddle e^{2 pi i k1 n2 / 512}, into a per-warp shared-memory tile;
//   step 2, TWO frames at a time, lane = (frame, n2): 32-point inverse DFT over k1 in registers; lane (f, n2) owns
//     z[16 n1 + n2] = (x[32 n1 + 2 n2], x[32 n1 + 2 n2 + 1]): one 8-byte store per n1, 128 B contiguous per half-warp.
// ~1.9 k instructions per frame instead of 4.8 k, 100 registers instead of 168 (16 warps per SM instead of 12).  Measured
// (262 k frames, same box, profiles/r02_vocos_kernels_ab.txt): 1910.9 -> 418.7 us = 5.15 TB/s of algorithmic traffic (4104 B in +
// 4096 B out per frame) = 79 % of the measured HBM peak; rel-L2 2.4e-7 against the first kernel.
template <int NP>
__device__ __forceinline__ void idft_regs(float2 (&v)[NP]) {   // radix-2 DIF like idft32_regs; output n in v[bit-reverse(n)]
  constexpr int STAGES = NP == 32 ? 5 : 4;
#pragma unroll
  for (int s = 0; s < STAGES; ++s) {
    const int half = (NP / 2) >> s;
#pragma unroll
    for (int g = 0; g < NP; g += 2 * half) {
#pragma unroll
      for (int j = 0; j < half; ++j) {
        const float2 a = v[g + j], b = v[g + j + half];
        v[g + j] = make_float2(a.x + b.x, a.y + b.y);
        const float2 d = make_float2(a.x - b.x, a.y - b.y);
        const int m = j * (16 / half);                       // twiddle e^{+2 pi i j / (2 half)} = W32^m
        if (m == 0) v[g + j + half] = d;
        else if (m == 8) v[g + j + half] = make_float2(-d.y, d.x);
        else v[g + j + half] = cmul(d, kW32[m]);
      }
    }
  }
}
__host__ __device__ constexpr int brev4(int i) { return ((i & 1) << 3) | ((i & 2) << 1) | ((i & 4) >> 1) | ((i & 8) >> 3); }

// min(exp(m), 100) (cos p + i sin p)
__device__ __forceinline__ float2 polar_fast(float m, float p) {
  const float mag = fminf(fast_ex2(m * 1.4426950408889634f), 100.f);
  const float j = rintf(p * 0.15915494309189535f);
  float r = fmaf(-j, 6.28125f, p);                           // 2 pi = 6.28125 (exact product for |j| < 2^16) + 1.9353071795864769e-3
  r = fmaf(-j, 1.9353071795864769e-3f, r);
  return make_float2(mag * __cosf(r), mag * __sinf(r));
}

constexpr int ISTFT2_WARPS = 4;
constexpr int ISTFT2_TILE = 32 * 17 + 16;   // words per frame tile: rows of 17 (conflict-free column writes); a warp's two tiles sit 16 banks apart
__global__ void __launch_bounds__(ISTFT2_WARPS * 32, 4) istft_frames2_kernel(const float* __restrict__ spec, long long lds, int rows,
                                                                            const float* __restrict__ window,
                                                                            float* __restrict__ frames) {
  pdl_wait();
  pdl_launch();
  __shared__ float tre[ISTFT2_WARPS][2][ISTFT2_TILE];
  __shared__ float tim[ISTFT2_WARPS][2][ISTFT2_TILE];
  __shared__ float2 wins[NFFT / 2];                          // (window[2 i], window[2 i + 1]) / 1024
  __shared__ float2 twz[512], twt[512];                      // the two twiddle tables, [k2 or n2][lane]: conflict-free LDS.64
  for (int i = threadIdx.x; i < NFFT / 2; i += blockDim.x)
    wins[i] = make_float2(window[2 * i] * (1.f / NFFT), window[2 * i + 1] * (1.f / NFFT));
  for (int i = threadIdx.x; i < 512; i += blockDim.x) {
    float2 t;
    sincospif(static_cast<float>(i) / 512.f, &t.y, &t.x);            // i = k1 + 32 k2 = k: w^k = e^{2 pi i k / 1024}
    twz[i] = t;
    sincospif(static_cast<float>((i & 31) * (i >> 5)) / 256.f, &t.y, &t.x);   // i = k1 + 32 n2: e^{2 pi i k1 n2 / 512}
    twt[i] = t;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int src_lane = (32 - lane) & 31;
  const int h2 = lane >> 4, n2s = lane & 15;                         // step 2: which of the two frames, which n2
  const int pairs = (rows + 1) >> 1;
  for (int pr = blockIdx.x * ISTFT2_WARPS + warp; pr < pairs; pr += gridDim.x * ISTFT2_WARPS) {
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      const int f = 2 * pr + h;
      if (f >= rows) break;                                          // warp-uniform
      const float* sp = spec + static_cast<size_t>(f) * lds;
      float2 X[17];
#pragma unroll
      for (int k2 = 0; k2 < 16; ++k2) {
        const int k = lane + 32 * k2;
        X[k2] = polar_fast(sp[k], sp[NBINS + k]);
      }
      X[16] = make_float2(0.f, 0.f);
      if (lane == 0) {
        X[0].y = 0.f;                                                // irfft ignores Im of DC / Nyquist
        X[16] = make_float2(polar_fast(sp[NFFT / 2], sp[NBINS + NFFT / 2]).x, 0.f);
      }
      float2 Z[16];
#pragma unroll
      for (int k2 = 0; k2 < 16; ++k2) {
        // partner bin 512 - (k1 + 32 k2) = (32 - k1) + 32 (15 - k2): lane 32 - k1; lane 0: 32 (16 - k2), its own
        const float ox = __shfl_sync(0xffffffffu, X[15 - k2].x, src_lane);
        const float oy = __shfl_sync(0xffffffffu, X[15 - k2].y, src_lane);
        const float2 own = X[16 - k2];
        const float2 B = lane == 0 ? make_float2(own.x, -own.y) : make_float2(ox, -oy);
        const float2 A = X[k2];
        const float2 S = make_float2(A.x + B.x, A.y + B.y);
        const float2 D = cmul(twz[32 * k2 + lane], make_float2(A.x - B.x, A.y - B.y));
        Z[k2] = make_float2(S.x - D.y, S.y + D.x);                   // S + i D
      }
      idft_regs<16>(Z);                                              // over k2: n2 in Z[brev4(n2)]
      float* re = tre[warp][h];
      float* im = tim[warp][h];
#pragma unroll
      for (int n2 = 0; n2 < 16; ++n2) {
        const float2 bv = n2 == 0 ? Z[0] : cmul(Z[brev4(n2)], twt[32 * n2 + lane]);
        re[lane * 17 + n2] = bv.x;
        im[lane * 17 + n2] = bv.y;
      }
    }
    __syncwarp();
    {
      const float* re = tre[warp][h2];
      const float* im = tim[warp][h2];
      float2 v[32];
#pragma unroll
      for (int k1 = 0; k1 < 32; ++k1) v[k1] = make_float2(re[k1 * 17 + n2s], im[k1 * 17 + n2s]);
      idft_regs<32>(v);                                              // over k1: z[16 n1 + n2] in v[brev5(n1)]
      const int f = 2 * pr + h2;
      if (f < rows) {
        float2* o = reinterpret_cast<float2*>(frames + static_cast<size_t>(f) * NFFT);
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) {
          const float2 wv = wins[16 * n1 + n2s];
          o[16 * n1 + n2s] = make_float2(v[brev5(n1)].x * wv.x, v[brev5(n1)].y * wv.y);
        }
      }
    }
    __syncwarp();                                                    // the tiles are rewritten by the next pair
  }
}

// grid (ceil(max_wav_len / 256), num_segs); seg = {row0, frames, wav_offset, _}
__global__ void __launch_bounds__(256) istft_ola_kernel(const float* __restrict__ frames, const float* __restrict__ window,
                                                        const int* __restrict__ seg, float* __restrict__ wav,
                                                        const float* __restrict__ gains) {
  pdl_wait();
  pdl_launch();
  const int4 sg = *reinterpret_cast<const int4*>(seg + 4 * blockIdx.y);
  const int T = sg.y;
  const int len = HOP * (T - 1);
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= len) return;
  const int q = s + NFFT / 2;                   // position in the un-trimmed (centre-padded) signal
  int f_hi = q / HOP;
  int f_lo = (q - (NFFT - 1) + HOP - 1) / HOP;  // ceil((q - 1023) / 256)
  if (f_lo < 0) f_lo = 0;
  if (f_hi > T - 1) f_hi = T - 1;
  float acc = 0.f, env = 0.f;
  for (int f = f_lo; f <= f_hi; ++f) {
    const int n = q - f * HOP;
    acc += frames[static_cast<size_t>(sg.x + f) * NFFT + n];
    const float w = window[n];
    env += w * w;
  }
  float y = acc / env;
  if (gains != nullptr) y *= gains[blockIdx.y];
  wav[static_cast<size_t>(sg.z) + s] = y;
}

}  // namespace f5

// 2 = real-input form (512-point complex FFT, two frames per warp; default), 1 = the first kernel (1024-point complex FFT).
// F5_ISTFT_V in the environment / f5_set_istft_variant select one (A/B measurements).
static int f5_istft_variant = [] { const char* e = getenv("F5_ISTFT_V"); return (e != nullptr && e[0] >= '1' && e[0] <= '2') ? e[0] - '0' : 2; }();
extern "C" int f5_set_istft_variant(int v) {
  const int old = f5_istft_variant;
  if (v >= 1 && v <= 2) f5_istft_variant = v;
  return old;
}

extern "C" int f5_istft_frames(const float* spec, int64_t lds, int32_t rows, const float* window, float* frames_out,
                               void* stream) {
  if (!spec || !window || !frames_out || rows <= 0 || lds < 2 * f5::NBINS) return F5_ERR_ARG;
  if (f5_istft_variant == 2) {
    const int blocks2 = ((rows + 1) / 2 + f5::ISTFT2_WARPS - 1) / f5::ISTFT2_WARPS;
    const int grid2 = blocks2 < 148 * 8 ? blocks2 : 148 * 8;
    f5::f5_launch(f5::istft_frames2_kernel, dim3(grid2), dim3(f5::ISTFT2_WARPS * 32), 0, reinterpret_cast<cudaStream_t>(stream), spec, lds, rows, window, frames_out);
    return static_cast<int>(cudaGetLastError());
  }
  const int blocks = (rows + f5::ISTFT_WARPS - 1) / f5::ISTFT_WARPS;
  const int grid = blocks < 148 * 6 ? blocks : 148 * 6;
  f5::f5_launch(f5::istft_frames_kernel, dim3(grid), dim3(f5::ISTFT_WARPS * 32), 0, reinterpret_cast<cudaStream_t>(stream), spec, lds, rows, window, frames_out);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int f5_istft_ola(const float* frames, const float* window, const int32_t* seg, int32_t num_segs,
                            int32_t max_wav_len, float* wav, const float* gains, void* stream) {
  if (!frames || !window || !seg || !wav || num_segs <= 0 || max_wav_len <= 0) return F5_ERR_ARG;
  dim3 grid((max_wav_len + 255) / 256, num_segs);
  f5::f5_launch(f5::istft_ola_kernel, dim3(grid), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), frames, window, seg, wav, gains);
  return static_cast<int>(cudaGetLastError());
}
