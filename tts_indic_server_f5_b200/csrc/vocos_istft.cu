// Vocos ISTFT head (vocos 0.1.0 `ISTFTHead`, padding="center"; call site infer/utils_infer.py:472), sm_100a.
//   istft_frames: per frame  mag = min(exp(m), 1e2); X = mag (cos p + i sin p); x = irfft_1024(X) * window     (fp32, smem FFT)
//   istft_ola   : overlap-add with hop 256, divide by the squared-window envelope, trim n_fft/2 each side, optional gain
// The inverse real FFT is a 1024-point complex FFT of the Hermitian-extended spectrum, one warp per frame, in registers
// (32 x 32 decomposition, one shared-memory transpose); the stage moves 4104 B in + 4096 B out per frame.
#include "f5_common.cuh"
#include "../../include/f5_b200.h"

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

extern "C" int f5_istft_frames(const float* spec, int64_t lds, int32_t rows, const float* window, float* frames_out,
                               void* stream) {
  if (!spec || !window || !frames_out || rows <= 0 || lds < 2 * f5::NBINS) return F5_ERR_ARG;
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
