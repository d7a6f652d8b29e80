// Microbenchmark: cycles per MUFU.EX2 warp-instruction for the softmax exp loop of attn_tcgen05.cu, with 1 or 2 warps
// per SM sub-partition.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_rate mufu_rate.cu && ./mufu_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../tts_indic_server_f5_b200/csrc/f5_common.cuh"
using namespace f5;

template <int MODE>
__global__ void k(float* out, long long* cyc, float scale, float m, int iters) {
  float r[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) r[i] = (threadIdx.x * 131 + i * 7) * 1e-4f * scale;
  float2 acc = make_float2(0.f, 0.f);
  uint32_t pacc = 0;
  const float2 sc2 = make_float2(scale, scale), nm2 = make_float2(-m, -m);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {           // MUFU only
#pragma unroll
      for (int i = 0; i < 128; ++i) r[i] = fast_ex2(r[i]) - 1.5f;
    } else if (MODE == 1) {    // the kernel's mix: ffma2, 2 ex2, fadd2, pack
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        float2 e = ffma2(make_float2(r[2 * i], r[2 * i + 1]), sc2, nm2);
        e.x = fast_ex2(e.x); e.y = fast_ex2(e.y);
        acc = fadd2(acc, e);
        pacc ^= pack_bf16x2(e.x, e.y);
        r[2 * i] = e.x - 1.5f; r[2 * i + 1] = e.y - 1.5f;
      }
    } else if (MODE == 2) {    // mix with 25 % of the pairs on the polynomial path
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        float2 e = ffma2(make_float2(r[2 * i], r[2 * i + 1]), sc2, nm2);
        if ((i & 3) == 0) e = exp2_poly2(e); else { e.x = fast_ex2(e.x); e.y = fast_ex2(e.y); }
        acc = fadd2(acc, e);
        pacc ^= pack_bf16x2(e.x, e.y);
        r[2 * i] = e.x - 1.5f; r[2 * i + 1] = e.y - 1.5f;
      }
    } else {                   // scalar ffma/fadd instead of the packed forms
#pragma unroll
      for (int i = 0; i < 128; ++i) {
        float e = fast_ex2(fmaf(r[i], scale, -m));
        acc.x += e;
        pacc ^= __float_as_uint(e) >> 16;
        r[i] = e - 1.5f;
      }
    }
  }
  const long long t1 = clock64();
  float s = acc.x + acc.y + __uint_as_float(pacc & 0xff);
#pragma unroll
  for (int i = 0; i < 128; ++i) s += r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int threads) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 200;
  k<MODE><<<148, threads>>>(out, cyc, 0.5f, 0.25f, iters);
  k<MODE><<<148, threads>>>(out, cyc, 0.5f, 0.25f, iters);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  const int warps_per_smsp = threads / 128;
  printf("%-28s %d warp(s)/SMSP: %.2f cycles per MUFU warp-instr per warp, %.2f per SMSP  (%s)\n", name, warps_per_smsp,
         c / (iters * 128.0), c / (iters * 128.0 * warps_per_smsp), cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int threads : {128, 256, 384}) {
    run<0>("mufu only", threads);
    run<1>("ffma2+ex2+fadd2+pack", threads);
    run<2>("same, 25% poly", threads);
    run<3>("scalar ffma+ex2+fadd", threads);
  }
  return 0;
}
