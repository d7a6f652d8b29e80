// Prompt log-mel front-end (reference f5_tts/model/modules.py:75-101 `get_vocos_mel_spectrogram` = torchaudio
// MelSpectrogram(n_fft 1024, hop 256, hann, center=True reflect, power=1, 100 HTK mel bands, norm=None) followed by
// log(clamp(1e-5))), sm_100a.  One CTA per frame:
//   reflect-padded 1024-sample frame x window -> 1024-point radix-2 FFT in shared memory (fp32) -> |X_k|, k <= 512
//   -> mel_m = sum_k fb[k][m] |X_k| over the band's non-zero bins only -> log(max(mel, 1e-5))
// Runs once per request on the 5 s prompt (469 frames): launch-latency sized, not roofline sized; it exists so that
// the prompt never leaves the device between the H2D copy and the sampler.
#include "f5_common.cuh"
#include "../../include/f5_b200.h"

namespace f5 {

constexpr int MEL_NFFT = 1024;
constexpr int MEL_HOP = 256;
constexpr int MEL_BINS = MEL_NFFT / 2 + 1;

__global__ void __launch_bounds__(256) mel_frames_kernel(const float* __restrict__ wave, const int* __restrict__ seg,
                                                         const float* __restrict__ window, const float* __restrict__ fbank,
                                                         const int* __restrict__ band, int n_mels, float* __restrict__ mel,
                                                         long long ldm) {
  pdl_wait();
  pdl_launch();
  __shared__ float2 buf[MEL_NFFT];
  __shared__ float2 tw[MEL_NFFT / 2];
  __shared__ float mag[MEL_BINS + 3];
  const int4 sg = *reinterpret_cast<const int4*>(seg + 4 * blockIdx.y);   // wave offset, samples, first output row, frames
  const int f = blockIdx.x;
  if (f >= sg.w) return;
  const float* w = wave + sg.x;
  const int nw = sg.y;
  for (int k = threadIdx.x; k < MEL_NFFT / 2; k += blockDim.x) {
    float s, c;
    sincospif(static_cast<float>(k) / (MEL_NFFT / 2), &s, &c);
    tw[k] = make_float2(c, -s);                        // e^{-2 pi i k / N}: forward transform
  }
  for (int n = threadIdx.x; n < MEL_NFFT; n += blockDim.x) {
    int q = f * MEL_HOP - MEL_NFFT / 2 + n;            // torch.stft center=True, pad_mode="reflect"
    if (q < 0) q = -q;
    if (q >= nw) q = 2 * (nw - 1) - q;
    q = min(max(q, 0), nw - 1);                        // (only reachable when nw < n_fft / 2; torch rejects that case)
    buf[__brev(static_cast<unsigned>(n)) >> 22] = make_float2(w[q] * window[n], 0.f);
  }
  __syncthreads();
#pragma unroll 1
  for (int s = 1; s <= 10; ++s) {
    const int half = 1 << (s - 1);
    for (int j = threadIdx.x; j < MEL_NFFT / 2; j += blockDim.x) {
      const int pos = j & (half - 1);
      const int i0 = ((j >> (s - 1)) << s) + pos;
      const int i1 = i0 + half;
      const float2 t0 = tw[pos << (10 - s)], b1 = buf[i1];
      const float2 t = make_float2(t0.x * b1.x - t0.y * b1.y, t0.x * b1.y + t0.y * b1.x);
      const float2 u = buf[i0];
      buf[i0] = make_float2(u.x + t.x, u.y + t.y);
      buf[i1] = make_float2(u.x - t.x, u.y - t.y);
    }
    __syncthreads();
  }
  for (int k = threadIdx.x; k < MEL_BINS; k += blockDim.x) mag[k] = sqrtf(buf[k].x * buf[k].x + buf[k].y * buf[k].y);
  __syncthreads();
  for (int m = threadIdx.x; m < n_mels; m += blockDim.x) {
    const int k0 = band[2 * m], k1 = band[2 * m + 1];   // non-zero bins of band m: [k0, k1)
    float acc = 0.f;
    for (int k = k0; k < k1; ++k) acc = fmaf(mag[k], fbank[static_cast<size_t>(k) * n_mels + m], acc);
    mel[static_cast<size_t>(sg.z + f) * ldm + m] = logf(fmaxf(acc, 1e-5f));
  }
}

}  // namespace f5

extern "C" int f5_mel_frames(const float* wave, const int32_t* seg, int32_t num_segs, int32_t max_frames, const float* window,
                             const float* fbank, const int32_t* band, int32_t n_mels, float* mel, int64_t ldm, void* stream) {
  using namespace f5;
  if (wave == nullptr || seg == nullptr || window == nullptr || fbank == nullptr || band == nullptr || mel == nullptr) return F5_ERR_ARG;
  if (num_segs <= 0 || max_frames <= 0 || n_mels <= 0 || n_mels > 1024 || ldm < n_mels) return F5_ERR_ARG;
  dim3 grid(max_frames, num_segs);
  f5_launch(mel_frames_kernel, dim3(grid), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), wave, seg, window, fbank, band, n_mels, mel, ldm);
  return static_cast<int>(cudaGetLastError());
}
