// Non-causal variable-length flash attention, head_dim 64, on tcgen05 (sm_100a).
//
// One CTA = one (query tile of 128 rows, head).  Keys/values of the query's own utterance are streamed in tiles of
// 128 through a 2-stage TMA ring.
//   warp 0    : TMA producer (Q once, then K_j / V_j)
//   warp 1    : MMA issuer: S = Q K_j^T  (128x128x64, SS) -> TMEM cols [0,128);   O += P_j V_j (128x64x128) -> TMEM cols [128,192)
//   warps 2-5 : online softmax, one thread per query row (= TMEM lane): two passes over S in TMEM (row max; exp2 + row sum),
//               P_j written as bf16 into a 128B-swizzled K-major smem tile, O rescaled in TMEM when the running max moved.
// Two CTAs fit per SM (112 KB smem, 256 TMEM columns each) so one CTA's softmax overlaps the other's MMAs.
//
// V tiles [keys, d] are read straight from the qkv buffer and fed to the PV MMA as an MN-major B operand (no transpose pass).
#include "f5_common.cuh"
#include "../../include/f5_b200.h"

namespace f5 {

constexpr int ATT_THREADS = 192;
constexpr int ATT_BM = 128;   // query rows per CTA
constexpr int ATT_BN = 128;   // keys per tile
constexpr int ATT_D = 64;
constexpr int ATT_TILE_BYTES = 128 * 64 * 2;  // 16 KB
constexpr int ATT_SMEM = ATT_TILE_BYTES * (1 + 2 + 2 + 2) + 128;   // 112 KB + barriers: two CTAs per SM
constexpr int ATT_TMEM_COLS = 256;

struct AttnParams {
  int q_col, k_col, v_col;
  const int* tiles;  // [num_tiles][4] = q_row0, kv_row0, kv_len, q_rows_valid
  __nv_bfloat16* out;
  long long ldo;
  float scale_log2;
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

__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_d64_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();   // 128B-swizzled tiles need 1024-B aligned bases
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + ATT_TILE_BYTES;          // 2 stages
  uint8_t* sV = sK + 2 * ATT_TILE_BYTES;      // 2 stages
  uint8_t* sP = sV + 2 * ATT_TILE_BYTES;      // 2 K-atom blocks of 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * ATT_TILE_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;    // [2]
  uint64_t* kv_empty = bars + 3;   // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* o_full = bars + 7;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int4 tile = *reinterpret_cast<const int4*>(p.tiles + 4 * blockIdx.x);
  const int q_row0 = tile.x, kv_row0 = tile.y, kv_len = tile.z, q_valid = tile.w;
  const int nkv = (kv_len + ATT_BN - 1) / ATT_BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    mbar_init(q_full, 1);
    mbar_init(&kv_full[0], 1); mbar_init(&kv_full[1], 1);
    mbar_init(&kv_empty[0], 1); mbar_init(&kv_empty[1], 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
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
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_O = tmem_base + ATT_BN;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, ATT_TILE_BYTES);
      tma_load_2d(sQ, &tmap_qkv, q_full, p.q_col + head * ATT_D, q_row0);
      for (int j = 0; j < nkv; ++j) {
        const int st = j & 1;
        mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&kv_full[st], 2 * ATT_TILE_BYTES);
        tma_load_2d(sK + st * ATT_TILE_BYTES, &tmap_qkv, &kv_full[st], p.k_col + head * ATT_D, kv_row0 + j * ATT_BN);
        tma_load_2d(sV + st * ATT_TILE_BYTES, &tmap_qkv, &kv_full[st], p.v_col + head * ATT_D, kv_row0 + j * ATT_BN);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_qk = umma_idesc_bf16(ATT_BM, ATT_BN);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(ATT_BM, ATT_D) | (1u << 16);   // B (= V) is MN-major
      const uint64_t qdesc = umma_desc_k_sw128(smem_u32(sQ));
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      {
        const uint64_t kdesc = umma_desc_k_sw128(smem_u32(sK));
#pragma unroll
        for (int kk = 0; kk < ATT_D / 16; ++kk) umma_f16_ss(tmem_S, qdesc + 2 * kk, kdesc + 2 * kk, idesc_qk, kk != 0);
        umma_commit(s_full);
      }
      for (int j = 0; j < nkv; ++j) {
        const int st = j & 1;
        mbar_wait(p_full, j & 1);          // softmax consumed S_j, wrote P_j and rescaled O
        tc_fence_after();
        if (j + 1 < nkv) {
          const int sn = (j + 1) & 1;
          mbar_wait(&kv_full[sn], ((j + 1) >> 1) & 1);
          tc_fence_after();
          const uint64_t kdesc = umma_desc_k_sw128(smem_u32(sK + sn * ATT_TILE_BYTES));
#pragma unroll
          for (int kk = 0; kk < ATT_D / 16; ++kk) umma_f16_ss(tmem_S, qdesc + 2 * kk, kdesc + 2 * kk, idesc_qk, kk != 0);
          umma_commit(s_full);
        }
        const uint32_t sp = smem_u32(sP);
        const uint32_t sv = smem_u32(sV + st * ATT_TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < ATT_BN / 16; ++kk) {
          const uint64_t pdesc = umma_desc_k_sw128(sp + (kk >> 2) * ATT_TILE_BYTES) + 2 * (kk & 3);
          umma_f16_ss(tmem_O, pdesc, umma_desc_mn_sw128(sv + kk * 2048), idesc_pv, (j | kk) != 0);   // 16 keys = 2 KB
        }
        umma_commit(&kv_empty[st]);
        umma_commit(o_full);
      }
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f;
    const float2 sc2 = make_float2(p.scale_log2, p.scale_log2);
    for (int j = 0; j < nkv; ++j) {
      const int kv_valid = min(ATT_BN, kv_len - j * ATT_BN);
      const bool full = kv_valid == ATT_BN;           // only the last tile of an utterance needs key masking
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      // pass 1: row max of the raw scores (S stays in TMEM; re-reading it is cheaper than holding 128 registers)
      float mx = -INFINITY;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t r0[32], r1[32];
        tmem_ld_32x32b_x32(tmem_S + lane_off + h * 64, r0);
        tmem_ld_32x32b_x32(tmem_S + lane_off + h * 64 + 32, r1);
        tmem_ld_wait();
        if (full) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fmaxf(__uint_as_float(r0[i]), __uint_as_float(r1[i])));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (h * 64 + i < kv_valid) mx = fmaxf(mx, __uint_as_float(r0[i]));
            if (h * 64 + 32 + i < kv_valid) mx = fmaxf(mx, __uint_as_float(r1[i]));
          }
        }
      }
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
      // previous PV must be complete before P is overwritten / O is rescaled
      if (j > 0) {
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after();
      }
      // pass 2: p = exp2(s*scale - m), row sum, bf16 P -> swizzled smem
      const float2 nm2 = make_float2(-m_run, -m_run);
      float2 ls2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int c = 0; c < ATT_BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_S + lane_off + c * 32, r);
        tmem_ld_wait();
        uint32_t pk[16];
        if (full) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float2 t = ffma2(make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), sc2, nm2);
            t.x = fast_ex2(t.x);
            t.y = fast_ex2(t.y);
            ls2 = fadd2(ls2, t);
            pk[i] = pack_bf16x2(t.x, t.y);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int k0 = c * 32 + 2 * i;
            const float e0 = k0 < kv_valid ? fast_ex2(fmaf(__uint_as_float(r[2 * i]), p.scale_log2, -m_run)) : 0.f;
            const float e1 = k0 + 1 < kv_valid ? fast_ex2(fmaf(__uint_as_float(r[2 * i + 1]), p.scale_log2, -m_run)) : 0.f;
            ls2.x += e0;
            ls2.y += e1;
            pk[i] = pack_bf16x2(e0, e1);
          }
        }
        uint8_t* blk = sP + (c >> 1) * ATT_TILE_BYTES + row * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = (c & 1) * 4 + q;
          *reinterpret_cast<uint4*>(blk + ((chunk ^ (row & 7)) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        }
      }
      l_run = l_run * alpha + (ls2.x + ls2.y);
      if (j > 0 && any_grow) {
#pragma unroll 1
        for (int c = 0; c < ATT_D / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(tmem_O + lane_off + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
          tmem_st_32x32b_x32(tmem_O + lane_off + c * 32, r);
        }
        tmem_st_wait();
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
    }
    mbar_wait(o_full, (nkv - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.f / l_run;
#pragma unroll 1
    for (int c = 0; c < ATT_D / 32; ++c) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(tmem_O + lane_off + c * 32, r);
      tmem_ld_wait();
      if (row < q_valid) {
        __nv_bfloat16* o = p.out + static_cast<size_t>(q_row0 + row) * p.ldo + head * ATT_D + c * 32;
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(r[i]) * inv_l, __uint_as_float(r[i + 1]) * inv_l);
          w.y = pack_bf16x2(__uint_as_float(r[i + 2]) * inv_l, __uint_as_float(r[i + 3]) * inv_l);
          w.z = pack_bf16x2(__uint_as_float(r[i + 4]) * inv_l, __uint_as_float(r[i + 5]) * inv_l);
          w.w = pack_bf16x2(__uint_as_float(r[i + 6]) * inv_l, __uint_as_float(r[i + 7]) * inv_l);
          *reinterpret_cast<uint4*>(o + i) = w;
        }
      }
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

}  // namespace f5

extern "C" int f5_attention_d64(const void* qkv, int64_t ld, int32_t rows, int32_t q_col, int32_t k_col, int32_t v_col,
                                int32_t heads, const int32_t* tiles, int32_t num_tiles, void* out, int64_t ldo,
                                float softmax_scale, void* stream) {
  using namespace f5;
  if (qkv == nullptr || tiles == nullptr || out == nullptr || num_tiles <= 0 || heads <= 0) return F5_ERR_ARG;
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
  p.q_col = q_col; p.k_col = k_col; p.v_col = v_col;
  p.tiles = tiles;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.scale_log2 = softmax_scale * 1.4426950408889634f;
  dim3 grid(num_tiles, heads);
  attn_d64_kernel<<<grid, ATT_THREADS, ATT_SMEM, reinterpret_cast<cudaStream_t>(stream)>>>(tq, p);
  return static_cast<int>(cudaGetLastError());
}
