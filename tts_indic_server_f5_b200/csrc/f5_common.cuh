// Common device helpers for the sm_100a kernels: mbarrier, TMA, tcgen05/TMEM inline-PTX wrappers,
// UMMA descriptors, small math helpers.  Compile only with -gencode arch=compute_100a,code=sm_100a.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>
#include <cstdio>

#ifndef F5_WATCHDOG_NS
// A stuck pipeline traps instead of hanging the GPU: 4 s of wall time (%globaltimer — per-SM clock64 counters are not
// comparable across SMs).  Before it traps, the waiting thread leaves a record in host-mapped memory (f5_diag_enable) that
// names the kernel shape, the CTA / thread, the barrier and the raw state of every barrier of the CTA, so that the host can
// tell a watchdog trap from a memory fault after the context has died (`_lib.read_diag`).
#define F5_WATCHDOG_NS (4000000000ll)
#endif
// F5_WAIT_ALL_LANES=1 restores the round-1 form of the warp-uniform issue loops, where all 32 lanes of a producer / MMA warp
// poll the mbarrier themselves (A/B builds for the soak test only; see mbar_wait_warp below).
#ifndef F5_WAIT_ALL_LANES
#define F5_WAIT_ALL_LANES 0
#endif

namespace f5 {

constexpr int kNumSMsB200 = 148;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------------ programmatic dependent launch (PDL)
// Every kernel of this library is launched with cudaLaunchAttributeProgrammaticStreamSerialization (f5_launch below) and calls
// pdl_wait() before it touches global memory a predecessor may have written (or writes anything a predecessor may read), so its
// set-up — barrier init, TMEM allocation, tensor-map prefetch, staging of weights — overlaps the tail of the previous kernel.
// pdl_launch() right after it lets the NEXT kernel's CTAs be scheduled as soon as every CTA of this grid is under way.  The
// single-request (B = 1) path is 5120 graph nodes of ~10 us each: the launch gap + prologue is a third of that.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ------------------------------------------------------------------ device-side fault record
// One pointer per translation unit (no -rdc): set by f5_diag_enable() through the F5_DEFINE_DIAG_SETTER each .cu defines.
// The record is ONE 16-byte store (any timed-out thread's record is a valid one; they are all stuck on the same pipeline):
//   u32[0] = 0xF5D00000 | F5_DIAG_TAG (one per .cu: which kernel family)     u32[1] = blockIdx.x | gridDim.x << 16
//   u32[2] = threadIdx.x | blockDim.x << 16                                  u32[3] = barrier smem address | parity << 31
// It is inlined at every wait (a call inside a setmaxnreg region does not register-allocate), so it is kept to a dozen
// instructions: the attention kernel is sensitive to its instruction footprint.  -DF5_DIAG_FULL=1 (soak builds) appends
// u64[2] = claim flag, u64[3] = ns waited, u64[8..40) = the 32 eight-byte words at (barrier & ~255): the raw state of every
// mbarrier of the CTA (each kernel keeps its barriers inside one 256-byte aligned block).
#ifndef F5_DIAG_FULL
#define F5_DIAG_FULL 0
#endif
#ifndef F5_DIAG_TAG
#define F5_DIAG_TAG 0u        // 1: gemm_tcgen05.cu, 2: attn_tcgen05.cu
#endif
#ifndef F5_DIAG
#define F5_DIAG 1             // 0: trap without a record (A/B builds that measure what the record costs)
#endif
static __device__ unsigned long long* f5_diag_ptr = nullptr;
#define F5_DEFINE_DIAG_SETTER(name)                                                                           \
  int name(void* mapped) {                                                                                    \
    unsigned long long* p = reinterpret_cast<unsigned long long*>(mapped);                                    \
    return static_cast<int>(cudaMemcpyToSymbol(f5::f5_diag_ptr, &p, sizeof(p), 0, cudaMemcpyHostToDevice));   \
  }
__device__ __forceinline__ void diag_report_and_trap(uint32_t bar, uint32_t parity, long long waited_ns) {
#if F5_DIAG
  unsigned long long* d = f5_diag_ptr;
  if (d != nullptr) {
    const uint4 rec = make_uint4(0xF5D00000u | F5_DIAG_TAG, blockIdx.x | (gridDim.x << 16), threadIdx.x | (blockDim.x << 16),
                                 bar | (parity << 31));
    asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(d), "r"(rec.x), "r"(rec.y), "r"(rec.z), "r"(rec.w) : "memory");
#if F5_DIAG_FULL
    if (atomicCAS(d + 2, 0ull, 1ull) == 0ull) {
      d[3] = static_cast<unsigned long long>(waited_ns);
      const uint32_t base = bar & ~255u;
#pragma unroll 1
      for (int i = 0; i < 32; ++i) {
        unsigned long long w;
        asm volatile("ld.shared.b64 %0, [%1];" : "=l"(w) : "r"(base + 8u * i));
        d[8 + i] = w;
      }
    }
#endif
    __threadfence_system();
  }
#endif
  __trap();
}

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// try_wait suspends the thread in hardware for a bounded time, so the loop is not a hot spin; the watchdog clock is only
// read every 4096 polls to keep the polling warps (one lane each) off the issue slots the math warps need.  The report is an
// out-of-line call on the cold path only (an inlined printf at every wait bloated the attention kernel past the instruction
// cache; the call sits behind a branch that is never taken in a healthy run).
__device__ __forceinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
  long long t0 = 0;
  for (uint32_t polls = 1;; ++polls) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if ((polls & 4095u) == 0) {
      long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      if (now - t0 > F5_WATCHDOG_NS) diag_report_and_trap(bar, parity, now - t0);
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  mbar_wait_slow(smem_u32(bar), parity);
}
// Slow path of mbar_wait_warp, OUT OF LINE: it is reached only when the first poll fails, and inlining its poll loop + watchdog
// at the ~45 wait sites of the attention kernel's MMA / producer loops inflates a loop whose instruction-cache footprint one
// sub-partition shares with two softmax warps.  (Only the low-register producer / MMA warps call it: a call from the
// setmaxnreg.inc regions does not register-allocate.)  No record from here, and twice the limit: a stuck pipeline stalls every
// role of the CTA within microseconds, so the per-thread waits of the softmax / epilogue warps time out first and leave it.
static __device__ __noinline__ void mbar_wait_warp_slow(uint32_t addr, uint32_t parity) {
  long long t0 = 0;
  for (uint32_t polls = 1;; ++polls) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (__all_sync(0xffffffffu, ok != 0)) return;
    if ((polls & 4095u) == 0) {
      long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      if (now - t0 > 2 * F5_WATCHDOG_NS) __trap();
    }
  }
}

// Wait of a whole producer / MMA warp that walks its loop warp-uniformly with ONE elected lane issuing (TMA, tcgen05.mma,
// commits).  All 32 lanes poll together and the warp leaves only when EVERY lane has seen the phase complete in the same
// poll (one VOTE.ALL on top of the try_wait; the loop condition is warp-uniform, so there is no divergence and the issue
// instructions that follow stay on the uniform datapath).  A parity wait is only correct while the waiter can never be
// two phases behind; with 32 independent pollers (the round-1 form) that would have to hold per LANE, with the vote it has
// to hold per WARP, which the kernels' protocols guarantee: for every barrier such a warp waits on, the barrier's next
// phase cannot start before this warp has issued something AFTER the wait (see the a_gate note in attn_tcgen05.cu for the
// one place where round 1 violated that).
// F5_WAIT_MODE: 0 = vote (default), 1 = round-1 form (every lane for itself; A/B soak builds), 2 = only the elected lane
// polls, the others park at a __syncwarp (measured ~20 % slower on the attention kernel: a divergent region per wait).
#ifndef F5_WAIT_MODE
#define F5_WAIT_MODE (F5_WAIT_ALL_LANES ? 1 : 0)
#endif
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity, bool leader) {
#if F5_WAIT_MODE == 1
  mbar_wait(bar, parity);
#elif F5_WAIT_MODE == 2
  if (leader) mbar_wait(bar, parity);
  __syncwarp();
#else
  const uint32_t addr = smem_u32(bar);
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(addr), "r"(parity)
      : "memory");
  if (__all_sync(0xffffffffu, ok != 0)) return;
  mbar_wait_warp_slow(addr, parity);
#endif
}
// Variants on a 32-bit shared-window address computed ONCE by the caller (the generic-pointer forms re-derive the window
// base from special registers at every call: S2R + LEA on the critical path of each barrier operation).
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  if (ok) return;
  mbar_wait_slow(bar, parity);
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// value the compiler may not rematerialise (keeps special-register reads / address arithmetic out of inner loops)
__device__ __forceinline__ uint32_t pin_u32(uint32_t v) {
  uint32_t r;
  asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
  return r;
}

// generic-proxy writes to smem -> visible to the async proxy (UMMA / TMA store)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ------------------------------------------------------------------ TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// 2-D tile -> the SAME shared-memory offset of every CTA in `cta_mask` of this cluster; each destination CTA's mbarrier at the
// same offset receives the complete_tx for the bytes that land in it
__device__ __forceinline__ void tma_load_2d_multicast(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                      uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// smem tile -> global tensor, elementwise fp32 add performed by the memory system (bulk async-group completion)
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_group_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_group() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

// ------------------------------------------------------------------ tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]; bf16/fp16 inputs, fp32 accumulate. Issued by ONE thread.
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// all previously issued tcgen05.mma of this thread arrive on `bar` when complete (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// same, arriving on the barrier at this offset in EVERY CTA of `cta_mask` (stage release of a multicast-fed ring)
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns; thread i gets lane (row) i.
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// registers -> TMEM: this warp's 32 lanes x 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------ UMMA descriptors
// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle: rows of 128 B (64 bf16), 8-row groups of
// 1024 B (SBO), tile base 1024-B aligned (base_offset 0).  Field layout: cute/arch/mma_sm100_desc.hpp SmemDescriptor.
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);       // start address            [0,14)
  d |= static_cast<uint64_t>(0) << 16;                           // LBO (unused for swizzled K-major) [16,30)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                   // SBO = 1024 B             [32,46)
  d |= static_cast<uint64_t>(1) << 46;                           // descriptor version (Blackwell) [46,48)
  d |= static_cast<uint64_t>(2) << 61;                           // layout = SWIZZLE_128B    [61,64)
  return d;
}
// Instruction descriptor for kind::f16: bf16 A/B (K-major), fp32 D, shape M x N (UMMA InstrDescriptor fields).
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

// ------------------------------------------------------------------ math
__device__ __forceinline__ float gelu_tanh(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float u = k0 * (x + k1 * x * x * x);
  return 0.5f * x * (1.f + tanhf(u));
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float mish(float x) {
  float sp = (x > 20.f) ? x : log1pf(expf(x));
  return x * tanhf(sp);
}
// MUFU-based fast forms used in GEMM epilogues (results are rounded to bf16 or added to an fp32 stream; ex2.approx has
// 2^-22 relative error, rcp.approx 1 ulp — far below bf16's 2^-9).
__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 0.5 x (1 + tanh(u)) == x * sigmoid(2u) == x / (1 + exp(-2u)),  u = sqrt(2/pi) (x + 0.044715 x^3)
__device__ __forceinline__ float gelu_tanh_fast(float x) {
  const float c0 = -2.f * 0.7978845608028654f * 1.4426950408889634f;   // -2 sqrt(2/pi) log2(e)
  const float t = c0 * (x + 0.044715f * x * x * x);
  return x * fast_rcp(1.f + fast_ex2(t));
}
// erf-GELU 0.5 x (1 + erf(x / sqrt 2)) with erf from Abramowitz-Stegun 7.1.26 (|error| < 1.5e-7 — three orders below the
// bf16 rounding of the GEMM output it feeds): erf(z) = 1 - (a1 t + ... + a5 t^5) e^{-z^2}, t = 1 / (1 + p z), z >= 0.
// One ex2 + one rcp instead of erff()'s ~40-instruction branchy polynomial: the K = 512 pointwise GEMMs of Vocos and of
// the text ConvNeXt blocks are epilogue-bound (256 activations per thread per 4096-cycle tile).
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = fast_rcp(fmaf(0.3275911f, z, 1.f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float e = fast_ex2(-z * z * 1.4426950408889634f);
  const float erf_abs = fmaf(-poly * t, e, 1.f);            // erf(|x| / sqrt 2)
  return 0.5f * x + 0.5f * fabsf(x) * erf_abs;              // 0.5 x (1 + sign(x) erf_abs)
}
// x tanh(log(1+e^x)) == x n / (n + 2),  n = e^x (e^x + 2)
__device__ __forceinline__ float mish_fast(float x) {
  const float e = fast_ex2(fminf(x, 20.f) * 1.4426950408889634f);
  const float n = e * (e + 2.f);
  return x * n * fast_rcp(n + 2.f);
}
// packed 2-wide fp32 (sm_100: FFMA2 / FADD2 halve the issue slots of the softmax inner loop)
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(reinterpret_cast<uint64_t&>(d))
      : "l"(reinterpret_cast<const uint64_t&>(a)), "l"(reinterpret_cast<const uint64_t&>(b)), "l"(reinterpret_cast<const uint64_t&>(c)));
  return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("add.f32x2 %0, %1, %2;"
      : "=l"(reinterpret_cast<uint64_t&>(d))
      : "l"(reinterpret_cast<const uint64_t&>(a)), "l"(reinterpret_cast<const uint64_t&>(b)));
  return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  float2 d;
  asm("mul.f32x2 %0, %1, %2;"
      : "=l"(reinterpret_cast<uint64_t&>(d))
      : "l"(reinterpret_cast<const uint64_t&>(a)), "l"(reinterpret_cast<const uint64_t&>(b)));
  return d;
}
// gelu_tanh_fast on a pair: the polynomial and the final products on the packed FP32 forms (5 packed + 4 MUFU instructions per
// two values instead of 10 + 4) — the FF1 epilogue applies it to 256 accumulators per thread per tile
__device__ __forceinline__ float2 gelu_tanh_fast2(float2 x) {
  const float c0 = -2.f * 0.7978845608028654f * 1.4426950408889634f;
  const float2 x2 = fmul2(x, x);
  const float2 p = ffma2(x2, make_float2(0.044715f * c0, 0.044715f * c0), make_float2(c0, c0));   // c0 (1 + 0.044715 x^2)
  const float2 t = fmul2(p, x);
  float2 e = make_float2(fast_ex2(t.x), fast_ex2(t.y));
  e = fadd2(e, make_float2(1.f, 1.f));
  return fmul2(x, make_float2(fast_rcp(e.x), fast_rcp(e.y)));
}
// gelu_erf_fast on a pair (packed polynomial; |x| via sign-bit masks)
__device__ __forceinline__ float2 gelu_erf_fast2(float2 x) {
  const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
  const float2 z = fmul2(ax, make_float2(0.70710678118654752f, 0.70710678118654752f));
  const float2 d = ffma2(z, make_float2(0.3275911f, 0.3275911f), make_float2(1.f, 1.f));
  const float2 t = make_float2(fast_rcp(d.x), fast_rcp(d.y));
  float2 p = ffma2(t, make_float2(1.061405429f, 1.061405429f), make_float2(-1.453152027f, -1.453152027f));
  p = ffma2(p, t, make_float2(1.421413741f, 1.421413741f));
  p = ffma2(p, t, make_float2(-0.284496736f, -0.284496736f));
  p = ffma2(p, t, make_float2(0.254829592f, 0.254829592f));
  const float2 zz = fmul2(z, make_float2(-1.4426950408889634f * z.x, -1.4426950408889634f * z.y));   // -z^2 log2(e)
  const float2 e = make_float2(fast_ex2(zz.x), fast_ex2(zz.y));
  const float2 pt = fmul2(p, t);
  const float2 erf_abs = ffma2(make_float2(-pt.x, -pt.y), e, make_float2(1.f, 1.f));                 // erf(|x| / sqrt 2)
  const float2 hx = fmul2(x, make_float2(0.5f, 0.5f));
  return ffma2(make_float2(0.5f * ax.x, 0.5f * ax.y), erf_abs, hx);                                  // 0.5 x + 0.5 |x| erf_abs
}
// exp2 on the FMA/ALU pipes for a pair of values (Cody-Waite split + degree-3 minimax polynomial, rel. error ~1e-4 — far
// below the bf16 rounding of P).  Used for a fraction of the softmax exponentials so that the MUFU (16 ex2/clk/SM) is not
// the only unit doing them.  Inputs must be <= ~100; they are clamped at -126 (2^-126 stands in for exp2(-inf) = 0).
__device__ __forceinline__ float2 exp2_poly2(float2 x) {
  x.x = fmaxf(x.x, -126.f);
  x.y = fmaxf(x.y, -126.f);
  const float2 magic = make_float2(12582912.f, 12582912.f);   // 1.5 * 2^23: low mantissa bits of (x + magic) hold floor(x)
  float2 r;
  asm("add.rm.ftz.f32x2 %0, %1, %2;"
      : "=l"(reinterpret_cast<uint64_t&>(r))
      : "l"(reinterpret_cast<const uint64_t&>(x)), "l"(reinterpret_cast<const uint64_t&>(magic)));
  const float2 nmagic = make_float2(-12582912.f, -12582912.f);
  const float2 fl = fadd2(r, nmagic);                          // floor(x), exact
  const float2 f = fadd2(x, make_float2(-fl.x, -fl.y));        // fractional part in [0, 1)
  float2 p = ffma2(f, make_float2(0.0771186f, 0.0771186f), make_float2(0.2275957f, 0.2275957f));
  p = ffma2(p, f, make_float2(0.6951786f, 0.6951786f));
  p = ffma2(p, f, make_float2(1.f, 1.f));
  p.x = __uint_as_float(__float_as_uint(p.x) + (__float_as_uint(r.x) << 23));   // multiply by 2^floor(x) through the exponent field
  p.y = __uint_as_float(__float_as_uint(p.y) + (__float_as_uint(r.y) << 23));
  return p;
}
__device__ __forceinline__ float silu(float x) { return x / (1.f + expf(-x)); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// v = hi + lo with hi = bf16(v), lo = bf16(v - hi): the two operand planes of the split-operand ("bf16x3") GEMM mode
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const float2 hf = __bfloat1622float2(h);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = pack_bf16x2(a - hf.x, b - hf.y);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace f5

// ------------------------------------------------------------------ host: launch with the PDL attribute
extern int f5_pdl_enabled;   // gemm_tcgen05.cu; f5_set_pdl()
namespace f5 {
template <typename... KArgs, typename... Args>
inline cudaError_t f5_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = f5_pdl_enabled ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
}  // namespace f5
