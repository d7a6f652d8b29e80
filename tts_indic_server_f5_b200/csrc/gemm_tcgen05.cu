// Persistent warp-specialised tcgen05 GEMM / implicit-conv kernel for sm_100a.
//
//   D[M,N] = A[M,K] * B[N,K]^T     bf16 operands (K-major, TMA 128B-swizzled tiles), fp32 accumulators in TMEM.
//
//   warp 0    : TMA producer    (one lane; ring of STAGES {A 128x64, B BLOCK_Nx64} tiles, full/empty mbarriers)
//   warp 1    : MMA issuer      (one lane; tcgen05.mma 128 x BLOCK_N x 16, commits release smem stages / signal epilogue)
//   warps 2-5 : epilogue        (tcgen05.ld 32 lanes x 32 columns -> registers -> fused epilogue -> global)
//   TMEM      : 2 accumulator buffers of BLOCK_N fp32 columns, so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Tiles are walked n-fastest so that the CTAs running concurrently share A row-blocks through L2 while B (weights,
// a few MB) stays L2-resident.
#define F5_DIAG_TAG 1u
#include "f5_common.cuh"
#include "../../include/f5_b200.h"
#include <cstdlib>

namespace f5 {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
// block size: 6 warps (producer, MMA, 4 epilogue) or, with EW = 8, 12 (producer, MMA, 2 idle, 8 epilogue)

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = BLOCK_N == 256 ? 4 : (BLOCK_N == 128 ? 6 : 8);
  static constexpr int TMEM_COLS = 2 * BLOCK_N;  // 512 / 256 / 128: powers of two >= 32
  static constexpr int EPI_STAGE_BYTES = 4 * 2 * 4096;   // two 32-row x 128-B staging tiles per epilogue warp (double buffer of the TMA reduce path)
  static constexpr int EPI_BIAS_BYTES = 2 * BLOCK_N * 4;  // the tile's bias and gate slices, shared by the four epilogue warps
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_STAGE_BYTES + EPI_BIAS_BYTES + 256 /*barriers*/;
};

struct GemmParams {
  int M, N, num_m_tiles, num_n_tiles, num_k;
  int kc_per_tap, tap_pad, a_grouped, b_tap_rows;
  int taps_per_seg, a_lo_off;   // split-operand ("bf16x3") mode: see f5_gemm_args
  int mode, act;
  const float* bias;
  const float* gate;
  void* out; long long ldo;
  void* out2; long long ldo2;
  const float* addend; long long ld_add;
  float* resid; long long ldr;
  const int* row_pos; int mask_rows;
  const float* rope; int rope_period, rope_tiles;
  int resid_tma;   // RESID mode through TMA reduce-add (N % 32 == 0)
};

// The activation is a compile-time parameter: a runtime switch would inline all three bodies into every unrolled
// element of the epilogue (tens of KB of straight-line code per warp -> instruction-fetch bound with one warp per scheduler).
template <int ACT>
__device__ __forceinline__ float apply_act(float v) {
  if constexpr (ACT == F5_ACT_GELU_TANH) return gelu_tanh_fast(v);
  else if constexpr (ACT == F5_ACT_GELU_ERF) return gelu_erf_fast(v);
  else if constexpr (ACT == F5_ACT_MISH) return mish_fast(v);
  else return v;
}

// named barrier 1: the epilogue warps (EW x 32 threads)
template <int EW>
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory"); }

// CL = 2: the kernel runs as clusters of two CTAs that work on vertically adjacent tiles (M-blocks 2i and 2i+1 of the same
// N-block).  Both need the same B (weight) tile for every k-block: each CTA fetches HALF of it and multicasts that half into
// both CTAs' shared memory, so the L2 -> SM traffic per k-block drops from 16 + 32 KB to 16 + 16 KB per CTA (ncu: 16 TB/s of
// L2 reads for the QKV GEMM, 58 % of the L2's peak and a large slice of the board's power budget).  A stage may be refilled
// only when BOTH CTAs' MMAs have consumed it, so `empty` barriers count two arrivals and every stage release is a multicast
// commit.  The MMAs themselves stay cta_group::1; nothing else in the roles changes.
// EW = 8 (bf16-store GEMMs with 256-wide tiles AND an activation: FF1, the text / Vocos pointwise convs): eight epilogue warps, two per TMEM lane quarter, each taking half
// of the tile's columns.  One epilogue warp per scheduler issues ~36 % of the time and is latency-bound (tmem load ->
// activation -> smem round trip -> stores per 64-column unit): FF1 with its GELU held the tensor pipe at 67 %.  The block is
// three warpgroups (producer / MMA / 2 idle warps, then the two epilogue warpgroups) so that setmaxnreg can move the
// producer warpgroup's registers to the epilogue (56 / 224).
template <int BLOCK_N, int ACT, int CL, int EW>
__global__ void __launch_bounds__((EW == 8 ? 12 : 6) * 32, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_r, const GemmParams p) {
  using Cfg = GemmCfg<BLOCK_N>;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();   // 128B-swizzled tiles need 1024-B aligned bases (no static smem in this kernel)
  uint8_t* stage_base = smem + Cfg::STAGES * Cfg::STAGE_BYTES;
  float* bias_base = reinterpret_cast<float*>(stage_base + Cfg::EPI_STAGE_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stage_base + Cfg::EPI_STAGE_BYTES + Cfg::EPI_BIAS_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tmem_full = empty_bar + Cfg::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // tile walk: unit u = first_unit, first_unit + num_walkers, ... ; CL = 1: unit = tile (m-block = u / num_n, n-fastest);
  // CL = 2: unit = (pair of m-blocks, n-block), this CTA takes m-block 2 * (u / num_n) + rank (a pair's second block may lie
  // beyond M: its loads are zero-filled and nothing is stored).
  const int cta_rank = CL == 1 ? 0 : static_cast<int>(cluster_ctarank());
  const int first_unit = static_cast<int>(blockIdx.x) / CL, num_walkers = static_cast<int>(gridDim.x) / CL;
  const int num_tiles = ((p.num_m_tiles + CL - 1) / CL) * p.num_n_tiles;   // units

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (p.resid_tma) tma_prefetch_desc(&tmap_r);
    for (int i = 0; i < Cfg::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], CL);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], EW * 32);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();            // the peer's barriers are initialised before anything can arrive on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();                                // everything above overlapped the previous kernel's tail; operands are read below
  pdl_launch();

  constexpr int EPI_W0 = EW == 8 ? 4 : 2;    // first epilogue warp
  if (warp < EPI_W0) {
  if constexpr (EW == 8) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    // The WHOLE warp walks the loops (warp-uniform control flow keeps coordinates / addresses in uniform registers, which
    // is what UTMALDG / UTCHMMA consume); only the TMA / MMA / commit instructions are predicated on one elected lane.
    // Issued from inside a divergent `if (lane == 0)` region every such instruction pays a serial R2UR chain (~85 cycles,
    // measured in the attention kernel): hidden under the 512 tensor cycles of a 256-wide k-block, but 2.2x the 192 cycles
    // of a 64-wide one — the grouped conv ran at 145 cycles per MMA against the 48 its shape allows.
    {
      const bool leader = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = first_unit; tile < num_tiles; tile += num_walkers) {
        const int m0 = ((tile / p.num_n_tiles) * CL + cta_rank) * BLOCK_M;
        const int n0 = (tile % p.num_n_tiles) * BLOCK_N;
        // tap = seg * taps_per_seg + t_in: t_in shifts the A rows (implicit conv), seg selects the operand planes of the
        // split-operand mode (A_hi B_hi | A_hi B_lo | A_lo B_hi: B's planes are stacked by rows like taps, A's low plane sits
        // a_lo_off columns to the right); dense / plain conv launches have one segment (taps_per_seg == num_taps)
        int tap = 0, kc = 0, t_in = 0, a_plane = 0;
#pragma unroll 1
        for (int k = 0; k < p.num_k; ++k) {
          mbar_wait_warp(&empty_bar[stage], phase ^ 1, leader);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          if (leader) {
            mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
            tma_load_2d(sa, &tmap_a, &full_bar[stage], (p.a_grouped ? n0 : 0) + kc * BLOCK_K + a_plane, m0 + t_in - p.tap_pad);
            if (CL == 1) {
              tma_load_2d(sb, &tmap_b, &full_bar[stage], kc * BLOCK_K, tap * p.b_tap_rows + n0);
            } else {   // my half of the B tile (BLOCK_N / 2 rows), into both CTAs; the other half arrives from the peer
              tma_load_2d_multicast(sb + cta_rank * (Cfg::B_BYTES / 2), &tmap_b, &full_bar[stage], kc * BLOCK_K,
                                    tap * p.b_tap_rows + n0 + cta_rank * (BLOCK_N / 2), static_cast<uint16_t>(3));
            }
          }
          __syncwarp();
          if (++kc == p.kc_per_tap) {
            kc = 0;
            ++tap;
            if (++t_in == p.taps_per_seg) { t_in = 0; a_plane = (tap == 2 * p.taps_per_seg) ? p.a_lo_off : 0; }
          }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
#if !F5_WAIT_ALL_LANES
      if (CL > 1) {
        // Producer tail: the last STAGES stage releases are multicast tcgen05.commit arrivals from BOTH CTAs of the pair, and
        // nobody waits for them any more.  barrier.cluster orders the executing threads' own operations, not an asynchronous
        // arrive that is still travelling to the peer's shared memory, so drain every stage's final phase before this CTA
        // may reach the cluster barrier and exit (the exited CTA's shared memory may already belong to the next kernel).
        for (int i = 0; i < Cfg::STAGES; ++i) {
          mbar_wait_warp(&empty_bar[stage], phase ^ 1, leader);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
#endif
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (warp-uniform, see the producer)
    {
      const bool leader = elect_one();
      constexpr uint32_t idesc = umma_idesc_bf16(BLOCK_M, BLOCK_N);
      const uint64_t desc0 = umma_desc_k_sw128(smem_u32(smem));
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = first_unit; tile < num_tiles; tile += num_walkers) {
        mbar_wait_warp(&tmem_empty[acc], acc_phase ^ 1, leader);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
#pragma unroll 1
        for (int k = 0; k < p.num_k; ++k) {
          mbar_wait_warp(&full_bar[stage], phase, leader);
          tc_fence_after();
          const uint64_t adesc = desc0 + static_cast<uint64_t>(stage) * (Cfg::STAGE_BYTES >> 4);   // address field is in 16-B units
          const uint64_t bdesc = adesc + (Cfg::A_BYTES >> 4);
          if (leader) {
#pragma unroll
            for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {
              // advance 16 bf16 = 32 B along K inside the 128-B swizzle row: +2 in the (addr >> 4) field
              umma_f16_ss(d_tmem, adesc + 2 * kk, bdesc + 2 * kk, idesc, (k | kk) != 0);
            }
            if (CL == 1) umma_commit(&empty_bar[stage]);   // smem stage reusable once these MMAs retire
            else umma_commit_multicast(&empty_bar[stage], static_cast<uint16_t>(3));   // ... in BOTH CTAs (either may write it)
          }
          __syncwarp();
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        if (leader) umma_commit(&tmem_full[acc]);       // accumulator ready for the epilogue
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  }
  } else {
    if constexpr (EW == 8) asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // ------------------------------------------------------------ epilogue (warps 2..5, or 4..11 when EW == 8)
    // Per 32-column unit: TMEM -> registers (thread = row) -> raw fp32 into a per-warp 4 KB staging tile (32 rows x 128 B,
    // 16-B chunks XOR-swizzled by row) -> read back with 8 lanes per row, so every global access (bf16/fp32 stores, the
    // fp32 residual read-modify-write, addend, bias, gate, RoPE table) is coalesced and each lane needs ONE bias/gate
    // float4 per unit.  All side loads are issued a unit ahead: with ~210 KB of smem carved out the L1 is tiny, so a
    // dependent global load on the critical path costs an L2 round trip.
    const int quarter = warp & 3;           // TMEM lane quarter this warp may access
    const bool epi_leader = elect_one();    // the lane that issues (and later waits for) this warp's TMA reduces
    const int col_half = EW == 8 ? (warp - EPI_W0) >> 2 : 0;      // EW == 8: which half of the tile's columns this warp owns
    uint8_t* stg = stage_base + (warp - EPI_W0) * (EW == 8 ? 4096 : 8192);
    if (EW == 8 && p.mode != F5_EPI_STORE_BF16) __trap();          // the 8-warp form is instantiated for the bf16-store epilogue only
    const int rd_row = lane >> 3, rd_chunk = lane & 7;
    constexpr int UNITS = BLOCK_N / 32;
    const float* side = p.mode == F5_EPI_RESID_F32 ? p.resid : (p.mode == F5_EPI_STORE_F32 ? p.addend : nullptr);
    const long long side_ld = p.mode == F5_EPI_RESID_F32 ? p.ldr : p.ld_add;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = first_unit; tile < num_tiles; tile += num_walkers) {
      const int m0 = ((tile / p.num_n_tiles) * CL + cta_rank) * BLOCK_M;
      const int n0 = (tile % p.num_n_tiles) * BLOCK_N;
      const int mw = m0 + quarter * 32;     // first row of this warp
      const bool rope_tile = p.rope != nullptr && (n0 % p.rope_period) == 0 && (n0 / p.rope_period) < p.rope_tiles;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BLOCK_N;
      uint8_t* wr = stg + lane * 128;
      int pos8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int mr = mw + i * 4 + rd_row;
        pos8[i] = (p.row_pos != nullptr && mr < p.M) ? p.row_pos[mr] : 0;
      }
      float4 nxt[8], nb, ng;
      auto prefetch = [&](int u) {
        const int col = n0 + u * 32 + rd_chunk * 4;
        const bool cok = col < p.N;
        nb = (p.bias != nullptr && cok) ? *reinterpret_cast<const float4*>(p.bias + col) : make_float4(0.f, 0.f, 0.f, 0.f);
        ng = (p.gate != nullptr && cok) ? *reinterpret_cast<const float4*>(p.gate + col) : make_float4(1.f, 1.f, 1.f, 1.f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int mr = mw + i * 4 + rd_row;
          nxt[i] = (side != nullptr && mr < p.M && cok)
                       ? *reinterpret_cast<const float4*>(side + static_cast<size_t>(mr) * side_ld + col)
                       : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      if (p.mode == F5_EPI_STORE_BF16) {
        // bf16 outputs: 64-column units (one full 128-B line per row), math in the thread = row layout, the tile's bias slice
        // served from a per-warp smem copy (fetched before the accumulator wait).
        float* bias_s = bias_base;
        epi_bar_sync<EW>();                                    // every epilogue warp is done with the previous tile's slice
        for (int c = (static_cast<int>(threadIdx.x) - EPI_W0 * 32) * 4; c < BLOCK_N; c += EW * 128) {
          const float4 b = (p.bias != nullptr && n0 + c < p.N) ? *reinterpret_cast<const float4*>(p.bias + n0 + c)
                                                               : make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4*>(bias_s + c) = b;
        }
        const int m = mw + lane;
        const int pos = (p.row_pos != nullptr && m < p.M) ? p.row_pos[m] : 0;
        const bool zero_row = p.mask_rows && pos < 0;
        epi_bar_sync<EW>();
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        constexpr int U_PER_WARP = (BLOCK_N / 64) / (EW == 8 ? 2 : 1);
#pragma unroll 1
        for (int u = col_half * U_PER_WARP; u < (col_half + 1) * U_PER_WARP; ++u) {
          uint32_t r0[32], r1[32];
          tmem_ld_32x32b_x32(taddr + u * 64, r0);
          tmem_ld_32x32b_x32(taddr + u * 64 + 32, r1);
          tmem_ld_wait();
          const int nc = n0 + u * 64;
          if (nc < p.N) {                                  // warp-uniform
            float v[64];
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 ba = *reinterpret_cast<const float4*>(bias_s + u * 64 + j);
              const float4 bb = *reinterpret_cast<const float4*>(bias_s + u * 64 + 32 + j);
              v[j] = __uint_as_float(r0[j]) + ba.x; v[j + 1] = __uint_as_float(r0[j + 1]) + ba.y;
              v[j + 2] = __uint_as_float(r0[j + 2]) + ba.z; v[j + 3] = __uint_as_float(r0[j + 3]) + ba.w;
              v[32 + j] = __uint_as_float(r1[j]) + bb.x; v[33 + j] = __uint_as_float(r1[j + 1]) + bb.y;
              v[34 + j] = __uint_as_float(r1[j + 2]) + bb.z; v[35 + j] = __uint_as_float(r1[j + 3]) + bb.w;
            }
            if constexpr (ACT != F5_ACT_NONE) {
#pragma unroll
              for (int j = 0; j < 64; j += 2) {
                if constexpr (ACT == F5_ACT_GELU_TANH) {
                  const float2 g2 = gelu_tanh_fast2(make_float2(v[j], v[j + 1]));
                  v[j] = g2.x; v[j + 1] = g2.y;
                } else if constexpr (ACT == F5_ACT_GELU_ERF) {
                  const float2 g2 = gelu_erf_fast2(make_float2(v[j], v[j + 1]));
                  v[j] = g2.x; v[j + 1] = g2.y;
                } else {
                  v[j] = apply_act<ACT>(v[j]); v[j + 1] = apply_act<ACT>(v[j + 1]);
                }
              }
            }
            if (rope_tile && u == 0 && pos >= 0) {
              // interleaved-pair rotation of head 0 (x-transformers apply_rotary_pos_emb; model/modules.py:418-419)
              const float4* cs = reinterpret_cast<const float4*>(p.rope + static_cast<size_t>(pos) * 64);
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float4 t = cs[i];
                const float x0 = v[4 * i], x1 = v[4 * i + 1], x2 = v[4 * i + 2], x3 = v[4 * i + 3];
                v[4 * i] = x0 * t.x - x1 * t.y; v[4 * i + 1] = x1 * t.x + x0 * t.y;
                v[4 * i + 2] = x2 * t.z - x3 * t.w; v[4 * i + 3] = x3 * t.z + x2 * t.w;
              }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              uint4 w = make_uint4(0, 0, 0, 0);
              if (!zero_row) {
                w.x = pack_bf16x2(v[8 * q], v[8 * q + 1]); w.y = pack_bf16x2(v[8 * q + 2], v[8 * q + 3]);
                w.z = pack_bf16x2(v[8 * q + 4], v[8 * q + 5]); w.w = pack_bf16x2(v[8 * q + 6], v[8 * q + 7]);
              }
              *reinterpret_cast<uint4*>(wr + ((q ^ (lane & 7)) << 4)) = w;
            }
            __syncwarp();
            const int col = nc + rd_chunk * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int row = i * 4 + rd_row;
              if (mw + row < p.M && col < p.N) {
                const uint4 w = *reinterpret_cast<const uint4*>(stg + row * 128 + ((rd_chunk ^ (row & 7)) << 4));
                *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<size_t>(mw + row) * p.ldo + col) = w;
              }
            }
            __syncwarp();
          }
        }
        tc_fence_before();
        mbar_arrive(&tmem_empty[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        continue;
      }
      if (p.resid_tma) {
        // x += gate * act(acc + bias) WITHOUT reading x into the SM: the update tile goes to smem (thread = row, 128-B rows,
        // 16-B chunks XOR-swizzled by row = the layout of a SWIZZLE_128B box) and one lane hands it to the TMA as a
        // reduce-add (`cp.reduce.async.bulk.tensor ... .add.f32`); the read-modify-write happens at the L2.  The previous
        // form loaded the residual tile through registers one 32-column unit ahead: 16 KB in flight per SM against the
        // ~45 KB the HBM latency-bandwidth product needs, so the K = 1024 out-projection ran at half of either roofline.
        float* bias_s = bias_base;
        float* gate_s = bias_base + BLOCK_N;
        epi_bar_sync<EW>();
        for (int c = (static_cast<int>(threadIdx.x) - EPI_W0 * 32) * 4; c < BLOCK_N; c += EW * 128) {
          const bool cok = n0 + c < p.N;
          *reinterpret_cast<float4*>(bias_s + c) = (p.bias != nullptr && cok) ? *reinterpret_cast<const float4*>(p.bias + n0 + c)
                                                                               : make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4*>(gate_s + c) = (p.gate != nullptr && cok) ? *reinterpret_cast<const float4*>(p.gate + n0 + c)
                                                                               : make_float4(1.f, 1.f, 1.f, 1.f);
        }
        epi_bar_sync<EW>();
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int u = 0; u < UNITS; ++u) {
          uint32_t r0[32];
          tmem_ld_32x32b_x32(taddr + u * 32, r0);
          tmem_ld_wait();
          if (u + 1 == UNITS) {                            // accumulator fully read: the MMA warp may start tile i+2 in it
            tc_fence_before();
            mbar_arrive(&tmem_empty[acc]);
          }
          uint8_t* buf = stg + (u & 1) * 4096;
          if (epi_leader) bulk_wait_group_read<1>();        // the reduce that last read this buffer (two units ago) is done with it
          __syncwarp();
          if (n0 + u * 32 < p.N) {                         // warp-uniform
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 b = *reinterpret_cast<const float4*>(bias_s + u * 32 + 4 * q);
              const float4 g = *reinterpret_cast<const float4*>(gate_s + u * 32 + 4 * q);
              float4 y = make_float4(__uint_as_float(r0[4 * q]) + b.x, __uint_as_float(r0[4 * q + 1]) + b.y,
                                     __uint_as_float(r0[4 * q + 2]) + b.z, __uint_as_float(r0[4 * q + 3]) + b.w);
              if constexpr (ACT != F5_ACT_NONE) {
                y.x = apply_act<ACT>(y.x); y.y = apply_act<ACT>(y.y); y.z = apply_act<ACT>(y.z); y.w = apply_act<ACT>(y.w);
              }
              *reinterpret_cast<float4*>(buf + lane * 128 + ((q ^ (lane & 7)) << 4)) = make_float4(g.x * y.x, g.y * y.y, g.z * y.z, g.w * y.w);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (epi_leader) {
              tma_reduce_add_2d(&tmap_r, buf, n0 + u * 32, mw);   // rows >= M / cols >= N are clipped by the tensor map
              bulk_commit_group();
            }
          }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        continue;
      }
      prefetch(0);                           // in flight while we wait for the accumulator
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int u = 0; u < UNITS; ++u) {
        float4 cur[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) cur[i] = nxt[i];
        const float4 b4 = nb, g4 = ng;
        if (u + 1 < UNITS) prefetch(u + 1);
        uint32_t r0[32];
        tmem_ld_32x32b_x32(taddr + u * 32, r0);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<uint4*>(wr + ((q ^ (lane & 7)) << 4)) = make_uint4(r0[4 * q], r0[4 * q + 1], r0[4 * q + 2], r0[4 * q + 3]);
        __syncwarp();
        const int col = n0 + u * 32 + rd_chunk * 4;
        if (col < p.N) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = i * 4 + rd_row;
            const int mr = mw + row;
            if (mr < p.M) {
              float4 y = *reinterpret_cast<const float4*>(stg + row * 128 + ((rd_chunk ^ (row & 7)) << 4));
              y.x += b4.x; y.y += b4.y; y.z += b4.z; y.w += b4.w;
              if constexpr (ACT != F5_ACT_NONE) {
                y.x = apply_act<ACT>(y.x); y.y = apply_act<ACT>(y.y); y.z = apply_act<ACT>(y.z); y.w = apply_act<ACT>(y.w);
              }
              if (p.mode == F5_EPI_STORE_F32) {
                y.x += cur[i].x; y.y += cur[i].y; y.z += cur[i].z; y.w += cur[i].w;
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + static_cast<size_t>(mr) * p.ldo + col) = y;
                if (p.out2 != nullptr) {
                  const bool zr = p.mask_rows && pos8[i] < 0;
                  const uint2 w = zr ? make_uint2(0, 0) : make_uint2(pack_bf16x2(y.x, y.y), pack_bf16x2(y.z, y.w));
                  *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out2) + static_cast<size_t>(mr) * p.ldo2 + col) = w;
                }
              } else {  // F5_EPI_RESID_F32: x += gate * y
                *reinterpret_cast<float4*>(p.resid + static_cast<size_t>(mr) * p.ldr + col) =
                    make_float4(fmaf(g4.x, y.x, cur[i].x), fmaf(g4.y, y.y, cur[i].y), fmaf(g4.z, y.z, cur[i].z), fmaf(g4.w, y.w, cur[i].w));
              }
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (p.resid_tma && epi_leader) bulk_wait_group<0>();   // every reduce-add of this warp has completed
  }

  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();            // no CTA exits while its peer can still multicast into it or arrive on its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
  }
  return fn;
}

// 2-D bf16 tensor [rows, cols] with row stride ld (elements); box = {64 cols, box_rows}; 128-B swizzle; OOB -> 0.
int make_tmap_bf16_2d(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_rows) {
  EncodeTiledFn enc = get_encode_tiled();
  if (enc == nullptr) return F5_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld % 8) != 0 || rows <= 0 || cols <= 0) return F5_ERR_ARG;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? F5_OK : F5_ERR_DRIVER;
}

// 2-D fp32 tensor [rows, cols] with row stride ld (elements); box = {32 cols, 32 rows} = 32 rows of 128 B; 128-B swizzle.
int make_tmap_f32_2d_box32(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld) {
  EncodeTiledFn enc = get_encode_tiled();
  if (enc == nullptr) return F5_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld % 4) != 0 || rows <= 0 || cols <= 0) return F5_ERR_ARG;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 4};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? F5_OK : F5_ERR_DRIVER;
}

template <int BLOCK_N, int ACT, int CL, int EW = 4>
int launch_gemm(const f5_gemm_args& a, const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N>;
  CUtensorMap ta, tb, tr;
  int rc = make_tmap_bf16_2d(&ta, a.A, a.a_rows, a.a_cols, a.lda, BLOCK_M);
  if (rc != F5_OK) return rc;
  rc = make_tmap_bf16_2d(&tb, a.B, a.b_rows, a.b_cols, a.ldb, BLOCK_N / CL);   // CL = 2: each CTA loads half of the B tile
  if (rc != F5_OK) return rc;
  if (p.resid_tma) {
    rc = make_tmap_f32_2d_box32(&tr, a.resid, a.M, a.N, a.ldr);
    if (rc != F5_OK) return rc;
  } else {
    tr = ta;
  }
  static const cudaError_t attr_rc =       // C++11 magic static (one per template instantiation): set once, thread-safe
      cudaFuncSetAttribute(gemm_tcgen05_kernel<BLOCK_N, ACT, CL, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
  if (attr_rc != cudaSuccess) return static_cast<int>(attr_rc);
  const int sms = a.num_sms > 0 ? a.num_sms : kNumSMsB200;
  const int units = ((p.num_m_tiles + CL - 1) / CL) * p.num_n_tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3((EW == 8 ? 12 : 6) * 32);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CL > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = CL;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (f5_pdl_enabled) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  int max_walkers = sms / CL;
  if (CL > 1) {        // persistent kernel: no more clusters than the device can hold at once (a GPC with an odd SM count strands one)
    static const int max_clusters = [&] {
      cudaLaunchConfig_t q = cfg;
      q.gridDim = dim3(kNumSMsB200);
      q.numAttrs = 1;                        // the cluster attribute only
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, gemm_tcgen05_kernel<BLOCK_N, ACT, CL, EW>, &q) != cudaSuccess || n <= 0) n = kNumSMsB200 / CL;
      return n;
    }();
    if (max_clusters < max_walkers) max_walkers = max_clusters;
  }
  const int walkers = units < max_walkers ? units : max_walkers;
  cfg.gridDim = dim3(walkers * CL);
  return static_cast<int>(cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<BLOCK_N, ACT, CL, EW>, ta, tb, tr, p));
}

}  // namespace f5

int f5_pdl_enabled = [] { const char* e = getenv("F5_PDL"); return (e == nullptr || e[0] != '0') ? 1 : 0; }();

// Programmatic dependent launch for every kernel of the library (default on; F5_PDL=0 in the environment or f5_set_pdl(0)
// switches it off: A/B measurements).  Returns the previous setting.
extern "C" int f5_set_pdl(int enabled) {
  const int old = f5_pdl_enabled;
  f5_pdl_enabled = enabled ? 1 : 0;
  return old;
}

F5_DEFINE_DIAG_SETTER(f5_diag_set_gemm)
int f5_diag_set_attn(void* mapped);   // attn_tcgen05.cu

// Fault record of the mbarrier watchdog: `mapped` is host memory the device can write (pinned, >= 40 x 8 bytes, zeroed);
// NULL switches reporting off.  See f5_common.cuh for the layout.
extern "C" int f5_diag_enable(void* mapped) {
  int rc = f5_diag_set_gemm(mapped);
  return rc != 0 ? rc : f5_diag_set_attn(mapped);
}

extern "C" int f5_gemm_bf16(const f5_gemm_args* a, void* stream) {
  using namespace f5;
  if (a == nullptr || a->A == nullptr || a->B == nullptr) return F5_ERR_ARG;
  if (a->M <= 0 || a->N <= 0 || (a->N % 8) != 0) return F5_ERR_ARG;
  if (a->num_taps < 1 || a->kc_per_tap < 1) return F5_ERR_ARG;
  if (a->block_n != 64 && a->block_n != 128 && a->block_n != 256) return F5_ERR_ARG;
  if (a->a_grouped && a->block_n != 64) return F5_ERR_ARG;
  GemmParams p;
  p.M = a->M; p.N = a->N;
  p.num_m_tiles = (a->M + BLOCK_M - 1) / BLOCK_M;
  p.num_n_tiles = (a->N + a->block_n - 1) / a->block_n;
  p.num_k = a->num_taps * a->kc_per_tap;
  p.kc_per_tap = a->kc_per_tap; p.tap_pad = a->tap_pad; p.a_grouped = a->a_grouped; p.b_tap_rows = a->b_tap_rows;
  p.taps_per_seg = a->taps_per_seg > 0 ? a->taps_per_seg : a->num_taps;
  p.a_lo_off = a->a_lo_off;
  if (a->taps_per_seg > 0 && (a->num_taps != 3 * a->taps_per_seg || a->a_lo_off <= 0 || (a->a_lo_off % 64) != 0)) return F5_ERR_ARG;
  p.mode = a->mode; p.act = a->act;
  p.bias = a->bias; p.gate = a->gate;
  p.out = a->out; p.ldo = a->ldo; p.out2 = a->out2; p.ldo2 = a->ldo2;
  p.addend = a->addend; p.ld_add = a->ld_add;
  p.resid = a->resid; p.ldr = a->ldr;
  p.row_pos = a->row_pos; p.mask_rows = a->mask_rows;
  p.rope = a->rope; p.rope_period = a->rope_period > 0 ? a->rope_period : 1; p.rope_tiles = a->rope_tiles;
  p.resid_tma = (a->mode == F5_EPI_RESID_F32 && (a->N % 32) == 0) ? 1 : 0;
  switch (a->mode) {
    case F5_EPI_STORE_BF16:
      if (a->out == nullptr || (a->ldo % 8) != 0) return F5_ERR_ARG;
      if (a->rope != nullptr && a->row_pos == nullptr) return F5_ERR_ARG;
      break;
    case F5_EPI_STORE_F32:
      if (a->out == nullptr || (a->ldo % 4) != 0) return F5_ERR_ARG;
      if (a->out2 != nullptr && (a->ldo2 % 8) != 0) return F5_ERR_ARG;
      if (a->addend != nullptr && (a->ld_add % 4) != 0) return F5_ERR_ARG;
      break;
    case F5_EPI_RESID_F32:
      if (a->resid == nullptr || (a->ldr % 4) != 0) return F5_ERR_ARG;
      break;
    default:
      return F5_ERR_ARG;
  }
  if (a->mask_rows && a->row_pos == nullptr) return F5_ERR_ARG;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (a->act < 0 || a->act > 3) return F5_ERR_ARG;
#define F5_DISPATCH(BN, CL)                                                        \
  switch (a->act) {                                                                \
    case F5_ACT_GELU_TANH: return launch_gemm<BN, F5_ACT_GELU_TANH, CL>(*a, p, s); \
    case F5_ACT_GELU_ERF: return launch_gemm<BN, F5_ACT_GELU_ERF, CL>(*a, p, s);   \
    case F5_ACT_MISH: return launch_gemm<BN, F5_ACT_MISH, CL>(*a, p, s);           \
    default: return launch_gemm<BN, F5_ACT_NONE, CL>(*a, p, s);                    \
  }
  // clusters of two CTAs sharing the weight tile: only where there are enough M-blocks for every SM pair (the big 256-wide
  // layer GEMMs); F5_GEMM_CLUSTER=0 in the environment forces the single-CTA form (A/B measurements)
  static const bool cluster_ok = [] { const char* e = getenv("F5_GEMM_CLUSTER"); return e == nullptr || e[0] != '0'; }();
  static const bool ew8_ok = [] { const char* e = getenv("F5_GEMM_EW8"); return e == nullptr || e[0] != '0'; }();
  if (a->block_n == 256) {
    const bool cl2 = cluster_ok && !a->a_grouped && p.num_m_tiles >= 2 * kNumSMsB200;
    // eight epilogue warps where the epilogue carries an activation (measured: FF1 + tanh-GELU 557 -> 532 us, but a plain bf16
    // store gets 2-6 % SLOWER with the larger block); F5_GEMM_EW8=0 forces four
    if (ew8_ok && a->mode == F5_EPI_STORE_BF16 && a->act != F5_ACT_NONE) {
      if (cl2) {
        switch (a->act) {
          case F5_ACT_GELU_TANH: return launch_gemm<256, F5_ACT_GELU_TANH, 2, 8>(*a, p, s);
          case F5_ACT_GELU_ERF: return launch_gemm<256, F5_ACT_GELU_ERF, 2, 8>(*a, p, s);
          default: return launch_gemm<256, F5_ACT_MISH, 2, 8>(*a, p, s);
        }
      }
      switch (a->act) {
        case F5_ACT_GELU_TANH: return launch_gemm<256, F5_ACT_GELU_TANH, 1, 8>(*a, p, s);
        case F5_ACT_GELU_ERF: return launch_gemm<256, F5_ACT_GELU_ERF, 1, 8>(*a, p, s);
        default: return launch_gemm<256, F5_ACT_MISH, 1, 8>(*a, p, s);
      }
    }
    if (cl2) { F5_DISPATCH(256, 2) }
    F5_DISPATCH(256, 1)
  }
  if (a->block_n == 128) { F5_DISPATCH(128, 1) }
  F5_DISPATCH(64, 1)
#undef F5_DISPATCH
}
