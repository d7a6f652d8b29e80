// Microbenchmark: cycles per tcgen05.mma (kind::f16, M = 128, K = 16, SS operands in shared memory) as a function of N and of
// the number of INDEPENDENT TMEM accumulators the issue stream alternates between.  One CTA per SM, one issuing thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_chain mma_chain.cu && ./mma_chain
#include <cstdio>
#include <cuda_runtime.h>
#include "../../tts_indic_server_f5_b200/csrc/f5_common.cuh"
using namespace f5;

template <int N, int ACCS>
__global__ void k(long long* cyc, int iters) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&tptr, 512); tmem_relinquish(); }
  for (int i = threadIdx.x; i < (128 * 64 * 2 + 256 * 64 * 2) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t tm = tptr;
    const uint64_t ad = umma_desc_k_sw128(smem_u32(smem)), bd = umma_desc_k_sw128(smem_u32(smem + 128 * 64 * 2));
    constexpr uint32_t idesc = umma_idesc_bf16(128, N);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
#pragma unroll
        for (int a = 0; a < ACCS; ++a) umma_f16_ss(tm + a * N, ad + 2 * kk, bd + 2 * kk, idesc, 1);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    cyc[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tptr, 512); }
}

template <int N, int ACCS>
void run() {
  long long* cyc; cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000, smem = 128 * 64 * 2 + 256 * 64 * 2 + 1024;
  cudaFuncSetAttribute(k<N, ACCS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<N, ACCS><<<148, 128, smem>>>(cyc, iters);
  k<N, ACCS><<<148, 128, smem>>>(cyc, iters);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  printf("N=%3d accumulators=%d: %.1f cycles per MMA (math floor %d)   %s\n", N, ACCS, c / (iters * 4.0 * ACCS), N / 2, cudaGetErrorString(cudaGetLastError()));
  cudaFree(cyc);
}

int main() {
  run<64, 1>(); run<64, 2>(); run<64, 4>();
  run<128, 1>(); run<128, 2>();
  run<256, 1>(); run<256, 2>();
  return 0;
}
