// Non-causal variable-length flash attention, head_dim 64, on tcgen05 (sm_100a).
//
// Persistent: one CTA per SM walks a list of work items = (pair of 128-row query tiles of one utterance, head).
//   warp 0     : TMA producer  (Q pair once per item; K_j / V_j tiles of 128 keys through a 2-stage ring)
//   warp 1     : MMA issuer    S_g = Q_g K_j^T (128x128x64, SS) -> TMEM;  O_g += P_g V_j (128x64x128, V as MN-major B operand
//                              read straight from the QKV buffer — no transpose pass) -> TMEM           (g = tile A / tile B)
//   warps 2-5  : softmax group A, warps 6-9: softmax group B — one thread per query row (= TMEM lane): row max, exp2 with
//                packed FFMA2/FADD2, P_g as bf16 into a 128B-swizzled smem tile, lazy rescale of O_g in TMEM, final O/l store.
// The two groups ping-pong on the SAME K/V tiles: while group A runs its exps (MUFU-bound: 16 k exps per tile vs 512 MMA
// cycles) the tensor core works for group B and vice versa, so neither the MUFU nor the tensor pipe waits on the
// softmax -> MMA -> softmax dependency chain of a single tile.  TMEM: S_A | S_B | O_A | O_B = 384 of 512 columns.
#include "f5_common.cuh"
#include "../../include/f5_b200.h"
#include <cstdlib>

namespace f5 {

constexpr int ATT_THREADS = 320;
constexpr int ATT_BM = 128;   // query rows per tile (two tiles per work item)
constexpr int ATT_BN = 128;   // keys per tile
constexpr int ATT_D = 64;
constexpr int ATT_TILE_BYTES = 128 * 64 * 2;  // 16 KB
constexpr int ATT_KV_STAGES = 2;
constexpr int ATT_SMEM = ATT_TILE_BYTES * (2 /*Q*/ + 2 * ATT_KV_STAGES /*K,V*/ + 8 /*P_A,P_B double-buffered*/) + 256;
constexpr int ATT_TMEM_COLS = 512;
#ifndef ATT_EXP_BF16X2
#define ATT_EXP_BF16X2 0    // 1: ex2.approx.ftz.bf16x2 (two exponentials per MUFU op). Measured on B200 (tools/attn_bench.py): C2 1549 us vs 1426 us with fp32 exps — the extra unpack ALU work costs more than the MUFU ops it saves
#endif
#ifndef ATT_POLY_COUNT
#define ATT_POLY_COUNT 0      // of every ATT_POLY_PERIOD score pairs, this many take the polynomial exp2 path (measured on B200: 25-50 % offload was NOT faster — the FMA/ALU pipes are already busy with scale, sum, max and bf16 packing)
#endif
#ifndef ATT_POLY_PERIOD
#define ATT_POLY_PERIOD 4
#endif

struct AttnParams {
  int q_col, k_col, v_col, heads;
  const int* items;   // [num_pairs][4] = q_row0, kv_row0, kv_len, q_rows_valid (1..256)
  int num_work;       // num_pairs * heads
  __nv_bfloat16* out;
  long long ldo;
  float scale_log2;
  long long* trace;   // optional clock64 event trace of CTA 0 (F5_ATTN_TRACE), [4 roles][512]
};

__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

// MN-major B operand (V tile [keys, 64 d], 128-B rows, 128B swizzle): 8-key groups 1024 B apart (SBO); N = 64 fits
// one swizzle row so the leading-dimension offset is never used.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(ATT_TILE_BYTES >> 4) << 16;   // LBO (unused: single 64-wide MN block)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;             // SBO
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

#define F5_TRACE(role, idx) do { if (p.trace != nullptr && blockIdx.x == 0 && (idx) < 512) p.trace[(role) * 512 + (idx)] = clock64(); } while (0)

__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_d64_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();   // 128B-swizzled tiles need 1024-B aligned bases
  uint8_t* sQ = smem;                                         // [2] tiles A, B
  uint8_t* sK = sQ + 2 * ATT_TILE_BYTES;                      // [ATT_KV_STAGES]
  uint8_t* sV = sK + ATT_KV_STAGES * ATT_TILE_BYTES;          // [ATT_KV_STAGES]
  uint8_t* sP = sV + ATT_KV_STAGES * ATT_TILE_BYTES;          // [2 groups][2 buffers][2 K-atom blocks of 16 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 8 * ATT_TILE_BYTES);
  uint64_t* q_full = bars;          // 1
  uint64_t* q_empty = bars + 1;     // 1
  uint64_t* kv_full = bars + 2;     // [ATT_KV_STAGES <= 3]
  uint64_t* kv_empty = bars + 5;    // [ATT_KV_STAGES <= 3]
  uint64_t* s_full = bars + 8;      // [2]
  uint64_t* p_full = bars + 10;     // [2]
  uint64_t* o_full = bars + 12;     // [2]
  uint64_t* o_free = bars + 14;     // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int i = 0; i < ATT_KV_STAGES; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&p_full[g], 128);
      mbar_init(&o_full[g], 1);
      mbar_init(&o_free[g], 128);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, ATT_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t item = 0, g_kv = 0;
      for (int w = blockIdx.x; w < p.num_work; w += gridDim.x, ++item) {
        const int4 it = *reinterpret_cast<const int4*>(p.items + 4 * (w / p.heads));
        const int head = w % p.heads;
        const int nkv = (it.z + ATT_BN - 1) / ATT_BN;
        mbar_wait(q_empty, (item & 1) ^ 1);                    // previous item's last QK has retired
        mbar_expect_tx(q_full, (it.w > ATT_BM ? 2 : 1) * ATT_TILE_BYTES);
        tma_load_2d(sQ, &tmap_qkv, q_full, p.q_col + head * ATT_D, it.x);
        if (it.w > ATT_BM) tma_load_2d(sQ + ATT_TILE_BYTES, &tmap_qkv, q_full, p.q_col + head * ATT_D, it.x + ATT_BM);
        for (int j = 0; j < nkv; ++j, ++g_kv) {
          const uint32_t st = g_kv % ATT_KV_STAGES, ph = (g_kv / ATT_KV_STAGES) & 1;
          mbar_wait(&kv_empty[st], ph ^ 1);
          mbar_expect_tx(&kv_full[st], 2 * ATT_TILE_BYTES);
          tma_load_2d(sK + st * ATT_TILE_BYTES, &tmap_qkv, &kv_full[st], p.k_col + head * ATT_D, it.y + j * ATT_BN);
          tma_load_2d(sV + st * ATT_TILE_BYTES, &tmap_qkv, &kv_full[st], p.v_col + head * ATT_D, it.y + j * ATT_BN);
          F5_TRACE(0, g_kv);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // The WHOLE warp walks the loop (warp-uniform control flow keeps descriptors / addresses in uniform registers, which is
    // what UTCHMMA consumes); only the tcgen05.mma / commit instructions themselves are predicated on one elected lane.
    // Issuing from inside a divergent `if (lane == 0)` region costs a serial R2UR chain per MMA (~85 cycles each, measured),
    // which made 24 MMAs per tile pair slower than the tensor work they describe.
    {
      constexpr uint32_t idesc_qk = umma_idesc_bf16(ATT_BM, ATT_BN);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(ATT_BM, ATT_D) | (1u << 16);   // B (= V) is MN-major
      const bool leader = elect_one();
      const uint64_t qd0 = umma_desc_k_sw128(smem_u32(sQ));
      const uint64_t kd0 = umma_desc_k_sw128(smem_u32(sK));
      const uint64_t pd0 = umma_desc_k_sw128(smem_u32(sP));
      const uint64_t vd0 = umma_desc_mn_sw128(smem_u32(sV));
      constexpr uint64_t TILE16 = ATT_TILE_BYTES >> 4;      // descriptor address field is in 16-B units
      uint32_t item = 0, g_kv = 0, ev = 0;
      uint32_t t[2] = {0, 0};        // tiles processed so far per group (phase of s_full / p_full / o_full)
      uint32_t it_g[2] = {0, 0};     // items processed so far per group (phase of o_free)
      auto issue_qk = [&](int g, uint32_t st) {
        const uint64_t qd = qd0 + g * TILE16, kd = kd0 + st * TILE16;
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < ATT_D / 16; ++kk)
            umma_f16_ss(tmem_base + g * ATT_BN, qd + 2 * kk, kd + 2 * kk, idesc_qk, kk != 0);
          umma_commit(&s_full[g]);
        }
        __syncwarp();
      };
      auto issue_pv = [&](int g, uint32_t st, bool first) {
        const uint64_t pd = pd0 + (g * 4 + (t[g] & 1) * 2) * TILE16, vd = vd0 + st * TILE16;
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < ATT_BN / 16; ++kk)            // 16 keys = 2 KB (128 x 16 B) of the V tile per MMA
            umma_f16_ss(tmem_base + 2 * ATT_BN + g * ATT_D, pd + (kk >> 2) * TILE16 + 2 * (kk & 3), vd + kk * 128, idesc_pv,
                        (!first || kk != 0) ? 1u : 0u);
          umma_commit(&o_full[g]);
        }
        __syncwarp();
      };
      for (int w = blockIdx.x; w < p.num_work; w += gridDim.x, ++item) {
        const int4 it = *reinterpret_cast<const int4*>(p.items + 4 * (w / p.heads));
        const int nkv = (it.z + ATT_BN - 1) / ATT_BN;
        const int ngroups = it.w > ATT_BM ? 2 : 1;
        mbar_wait(q_full, item & 1);
        {
          const uint32_t st = g_kv % ATT_KV_STAGES, ph = (g_kv / ATT_KV_STAGES) & 1;
          mbar_wait(&kv_full[st], ph);
          tc_fence_after();
          for (int g = 0; g < ngroups; ++g) issue_qk(g, st);
          if (nkv == 1 && leader) umma_commit(q_empty);      // every QK of this item is issued: the Q tiles may be refilled
        }
        for (int j = 0; j < nkv; ++j, ++g_kv) {
          const uint32_t st = g_kv % ATT_KV_STAGES;
          const bool more = j + 1 < nkv;
          uint32_t stn = 0;
          if (more) {
            stn = (g_kv + 1) % ATT_KV_STAGES;
            mbar_wait(&kv_full[stn], ((g_kv + 1) / ATT_KV_STAGES) & 1);
          }
          if (lane == 0) { F5_TRACE(1, ev); }
          ++ev;
          for (int g = 0; g < ngroups; ++g) {
            mbar_wait(&p_full[g], t[g] & 1);       // group g consumed S_g(j) and wrote P_g(j)
            tc_fence_after();
            if (lane == 0) { F5_TRACE(1, ev); }
            ++ev;
            if (more) issue_qk(g, stn);            // S_g is free again: next scores first, so group g never waits on its own PV
            if (j == 0) {                          // O_g of the previous item must have been read out
              mbar_wait(&o_free[g], (it_g[g] & 1) ^ 1);
              tc_fence_after();
            }
            issue_pv(g, st, j == 0);
            if (lane == 0) { F5_TRACE(1, ev); }
            ++ev;
            ++t[g];
          }
          if (leader) {
            umma_commit(&kv_empty[st]);              // both groups' PV on (K_j, V_j) retired -> stage reusable
            if (more && j + 2 == nkv) umma_commit(q_empty);   // the last tile's QKs were just issued: next item's Q may load
          }
          __syncwarp();
        }
        for (int g = 0; g < ngroups; ++g) ++it_g[g];
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax groups
    const int g = (warp - 2) >> 2;                // 0: tile A, 1: tile B
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tmem_S = tmem_base + g * ATT_BN;
    const uint32_t tmem_O = tmem_base + 2 * ATT_BN + g * ATT_D;
    uint8_t* sPg = sP + g * 4 * ATT_TILE_BYTES;
    const float2 sc2 = make_float2(p.scale_log2, p.scale_log2);
    uint32_t t = 0;
    for (int w = blockIdx.x; w < p.num_work; w += gridDim.x) {
      const int4 it = *reinterpret_cast<const int4*>(p.items + 4 * (w / p.heads));
      const int head = w % p.heads;
      const int q_valid = it.w - g * ATT_BM;      // rows of this group's tile that exist
      if (q_valid <= 0) continue;                 // tile B absent: the MMA warp skips this group too
      const int kv_len = it.z;
      const int nkv = (kv_len + ATT_BN - 1) / ATT_BN;
      float m_run = -INFINITY, l_run = 0.f;
      for (int j = 0; j < nkv; ++j, ++t) {
        const int kv_valid = min(ATT_BN, kv_len - j * ATT_BN);
        const bool full = kv_valid == ATT_BN;     // only the last tile of an utterance needs key masking
        mbar_wait(&s_full[g], t & 1);
        tc_fence_after();
        if (row == 0) F5_TRACE(2 + g, 4 * t);
        // S row (128 fp32) is read from TMEM ONCE into registers: LDTM bandwidth, not MUFU, limited the two-pass version.
        uint32_t r[128];
        tmem_ld_32x32b_x32(tmem_S + lane_off, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
        tmem_ld_32x32b_x32(tmem_S + lane_off + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
        tmem_ld_32x32b_x32(tmem_S + lane_off + 64, *reinterpret_cast<uint32_t(*)[32]>(&r[64]));
        tmem_ld_32x32b_x32(tmem_S + lane_off + 96, *reinterpret_cast<uint32_t(*)[32]>(&r[96]));
        tmem_ld_wait();
        if (!full) {
#pragma unroll
          for (int i = 0; i < 128; ++i)
            if (i >= kv_valid) r[i] = 0xff800000u;   // -inf: masked keys drop out of the max and give exp2 = 0
        }
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < 64; ++i) mx = fmaxf(mx, fmaxf(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])));
        if (row == 0) F5_TRACE(2 + g, 4 * t + 1);
        // lazy rescale: keep a stale running max while it is within 2^8 of the true one (p <= 256 is harmless in bf16/fp32);
        // O and l are only rescaled when the max really moved.  The first tile always "grows" (m_run = -inf, alpha = 0).
        const float m_new = fmaxf(m_run, mx * p.scale_log2);
        const bool grow = (m_new - m_run) > 8.f;
        const bool any_grow = __any_sync(0xffffffffu, grow);
        float alpha = 1.f;
        if (grow) {
          alpha = fast_ex2(m_run - m_new);
          m_run = m_new;
        }
        // p = exp2(s*scale - m), row sum, bf16 P -> swizzled smem.  P_g is double-buffered: the buffer written for tile t was
        // last read by PV_g(t-2), which retired before QK_g(t) (in-order tensor pipe) — no wait on the previous PV needed.
        if (row == 0) F5_TRACE(2 + g, 4 * t + 2);
        const float2 nm2 = make_float2(-m_run, -m_run);
        float2 ls2 = make_float2(0.f, 0.f);
        uint8_t* sPt = sPg + (t & 1) * 2 * ATT_TILE_BYTES;
#pragma unroll
        for (int c = 0; c < ATT_BN / 32; ++c) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float2 e = ffma2(make_float2(__uint_as_float(r[c * 32 + 2 * i]), __uint_as_float(r[c * 32 + 2 * i + 1])), sc2, nm2);
#if ATT_EXP_BF16X2
            // one MUFU op yields both exponentials, already in the bf16 the PV MMA consumes; the row sum is taken over the
            // SAME rounded values (numerator and denominator stay consistent).
            const uint32_t pb = ex2_bf16x2(pack_bf16x2(e.x, e.y));
            pk[i] = pb;
            ls2 = fadd2(ls2, make_float2(__uint_as_float(pb << 16), __uint_as_float(pb & 0xffff0000u)));
#else
            if ((i % ATT_POLY_PERIOD) < ATT_POLY_COUNT) {   // this pair on the FMA/ALU pipes, the others on the MUFU
              e = exp2_poly2(e);
            } else {
              e.x = fast_ex2(e.x);
              e.y = fast_ex2(e.y);
            }
            ls2 = fadd2(ls2, e);
            pk[i] = pack_bf16x2(e.x, e.y);
#endif
          }
          uint8_t* blk = sPt + (c >> 1) * ATT_TILE_BYTES + row * 128;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int chunk = (c & 1) * 4 + q;
            *reinterpret_cast<uint4*>(blk + ((chunk ^ (row & 7)) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
          }
        }
        l_run = l_run * alpha + (ls2.x + ls2.y);
        if (j > 0 && any_grow) {
          mbar_wait(&o_full[g], (t - 1) & 1);     // the previous PV of this group must have retired before O is rescaled
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < ATT_D / 32; ++c) {
            uint32_t ro[32];
            tmem_ld_32x32b_x32(tmem_O + lane_off + c * 32, ro);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) ro[i] = __float_as_uint(__uint_as_float(ro[i]) * alpha);
            tmem_st_32x32b_x32(tmem_O + lane_off + c * 32, ro);
          }
          tmem_st_wait();
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&p_full[g]);
        if (row == 0) F5_TRACE(2 + g, 4 * t + 3);
      }
      // epilogue of the item: O / l -> bf16 rows of the output
      mbar_wait(&o_full[g], (t - 1) & 1);
      tc_fence_after();
      const float inv_l = 1.f / l_run;
#pragma unroll 1
      for (int c = 0; c < ATT_D / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_O + lane_off + c * 32, r);
        tmem_ld_wait();
        if (row < q_valid) {
          __nv_bfloat16* o = p.out + static_cast<size_t>(it.x + g * ATT_BM + row) * p.ldo + head * ATT_D + c * 32;
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 wv;
            wv.x = pack_bf16x2(__uint_as_float(r[i]) * inv_l, __uint_as_float(r[i + 1]) * inv_l);
            wv.y = pack_bf16x2(__uint_as_float(r[i + 2]) * inv_l, __uint_as_float(r[i + 3]) * inv_l);
            wv.z = pack_bf16x2(__uint_as_float(r[i + 4]) * inv_l, __uint_as_float(r[i + 5]) * inv_l);
            wv.w = pack_bf16x2(__uint_as_float(r[i + 6]) * inv_l, __uint_as_float(r[i + 7]) * inv_l);
            *reinterpret_cast<uint4*>(o + i) = wv;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&o_free[g]);                    // O_g may be overwritten by the next item's first PV
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, ATT_TMEM_COLS);
  }
}

int make_tmap_bf16_2d(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_rows);
static long long* g_trace_buf = nullptr;

}  // namespace f5

extern "C" int f5_attention_trace_dump(long long* host) {
  if (f5::g_trace_buf == nullptr) return F5_ERR_ARG;
  cudaDeviceSynchronize();
  return static_cast<int>(cudaMemcpy(host, f5::g_trace_buf, 4 * 512 * sizeof(long long), cudaMemcpyDeviceToHost));
}

extern "C" int f5_attention_d64(const void* qkv, int64_t ld, int32_t rows, int32_t q_col, int32_t k_col, int32_t v_col,
                                int32_t heads, const int32_t* items, int32_t num_items, void* out, int64_t ldo,
                                float softmax_scale, void* stream) {
  using namespace f5;
  if (qkv == nullptr || items == nullptr || out == nullptr || num_items <= 0 || heads <= 0) return F5_ERR_ARG;
  if ((ldo % 8) != 0) return F5_ERR_ARG;
  const int cols = (q_col > k_col ? (q_col > v_col ? q_col : v_col) : (k_col > v_col ? k_col : v_col)) + heads * ATT_D;
  CUtensorMap tq;
  int rc = make_tmap_bf16_2d(&tq, qkv, rows, cols, ld, 128);
  if (rc != F5_OK) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_d64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set = true;
  }
  AttnParams p;
  p.q_col = q_col; p.k_col = k_col; p.v_col = v_col; p.heads = heads;
  p.items = items;
  p.num_work = num_items * heads;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.scale_log2 = softmax_scale * 1.4426950408889634f;
  static long long* trace_buf = nullptr;
  if (getenv("F5_ATTN_TRACE") != nullptr && trace_buf == nullptr) {
    cudaMalloc(&trace_buf, 4 * 512 * sizeof(long long));
    cudaMemset(trace_buf, 0, 4 * 512 * sizeof(long long));
  }
  p.trace = trace_buf;
  g_trace_buf = trace_buf;
  const int grid = p.num_work < kNumSMsB200 ? p.num_work : kNumSMsB200;
  attn_d64_kernel<<<grid, ATT_THREADS, ATT_SMEM, reinterpret_cast<cudaStream_t>(stream)>>>(tq, p);
  return static_cast<int>(cudaGetLastError());
}
