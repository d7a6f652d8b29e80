// Vocos ISTFT head (vocos 0.1.0 `ISTFTHead`, padding="center"; call site infer/utils_infer.py:472), sm_100a.
//   istft_frames: per frame  mag = min(exp(m), 1e2); X = mag (cos p + i sin p); x = irfft_1024(X) * window     (fp32, smem FFT)
//   istft_ola   : overlap-add with hop 256, divide by the squared-window envelope, trim n_fft/2 each side, optional gain
// The inverse real FFT is done as a 1024-point complex radix-2 FFT of the Hermitian-extended spectrum in shared memory
// (25 kFLOP/frame; the stage is HBM-bound: 4104 B in + 4096 B out per frame).
#include "f5_common.cuh"
#include "../../include/f5_b200.h"

namespace f5 {

constexpr int NFFT = 1024;
constexpr int NBINS = NFFT / 2 + 1;
constexpr int HOP = 256;

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

__global__ void __launch_bounds__(256) istft_frames_kernel(const float* __restrict__ spec, long long lds, int rows,
                                                           const float* __restrict__ window, float* __restrict__ frames) {
  __shared__ float2 buf[NFFT];
  __shared__ float2 tw[NFFT / 2];
  for (int k = threadIdx.x; k < NFFT / 2; k += blockDim.x) {
    float s, c;
    sincospif(static_cast<float>(k) / (NFFT / 2), &s, &c);   // e^{+2 pi i k / N}
    tw[k] = make_float2(c, s);
  }
  for (int f = blockIdx.x; f < rows; f += gridDim.x) {
    __syncthreads();
    const float* sp = spec + static_cast<size_t>(f) * lds;
    for (int k = threadIdx.x; k < NBINS; k += blockDim.x) {
      const float mag = fminf(expf(sp[k]), 100.f);
      float s, c;
      sincosf(sp[NBINS + k], &s, &c);
      float2 X = make_float2(mag * c, mag * s);
      if (k == 0 || k == NFFT / 2) X.y = 0.f;          // irfft ignores the imaginary part of DC / Nyquist
      buf[__brev(static_cast<unsigned>(k)) >> 22] = X;
      if (k > 0 && k < NFFT / 2) buf[__brev(static_cast<unsigned>(NFFT - k)) >> 22] = make_float2(X.x, -X.y);
    }
    __syncthreads();
#pragma unroll 1
    for (int s = 1; s <= 10; ++s) {
      const int half = 1 << (s - 1);
      for (int j = threadIdx.x; j < NFFT / 2; j += blockDim.x) {
        const int pos = j & (half - 1);
        const int i0 = ((j >> (s - 1)) << s) + pos;
        const int i1 = i0 + half;
        const float2 t = cmul(tw[pos << (10 - s)], buf[i1]);
        const float2 u = buf[i0];
        buf[i0] = make_float2(u.x + t.x, u.y + t.y);
        buf[i1] = make_float2(u.x - t.x, u.y - t.y);
      }
      __syncthreads();
    }
    float* o = frames + static_cast<size_t>(f) * NFFT;
    for (int n = threadIdx.x; n < NFFT; n += blockDim.x) o[n] = buf[n].x * (1.f / NFFT) * window[n];
  }
}

// grid (ceil(max_wav_len / 256), num_segs); seg = {row0, frames, wav_offset, _}
__global__ void __launch_bounds__(256) istft_ola_kernel(const float* __restrict__ frames, const float* __restrict__ window,
                                                        const int* __restrict__ seg, float* __restrict__ wav,
                                                        const float* __restrict__ gains) {
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
  const int grid = rows < 148 * 8 ? rows : 148 * 8;
  f5::istft_frames_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(spec, lds, rows, window, frames_out);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int f5_istft_ola(const float* frames, const float* window, const int32_t* seg, int32_t num_segs,
                            int32_t max_wav_len, float* wav, const float* gains, void* stream) {
  if (!frames || !window || !seg || !wav || num_segs <= 0 || max_wav_len <= 0) return F5_ERR_ARG;
  dim3 grid((max_wav_len + 255) / 256, num_segs);
  f5::istft_ola_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(frames, window, seg, wav, gains);
  return static_cast<int>(cudaGetLastError());
}
