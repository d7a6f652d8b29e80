// Memory-bound kernels of the F5-TTS hot path (sm_100a): fused LayerNorm+modulation, depthwise conv + LayerNorm, GRN,
// token gather, bf16 packing, CFG+Euler update, time embedding.  All are coalesced / 128-bit vectorised; row-wise
// reductions use warp shuffles (one warp per row), grids are sized from the row count.
#include "f5_common.cuh"
#include "../../include/f5_b200.h"
#include <cstdlib>

namespace f5 {

// ------------------------------------------------------------------------------------------------ LayerNorm * a + b -> bf16
template <int NV>  // float4 per lane: D = NV * 128
__global__ void __launch_bounds__(256) layernorm_mod_kernel(const float* __restrict__ x, long long ldx,
                                                            __nv_bfloat16* __restrict__ y, long long ldy,
                                                            float* __restrict__ y32, long long ldy32, int M,
                                                            const float* __restrict__ a, const float* __restrict__ b,
                                                            float a_off, float eps, int lo_off) {
  pdl_wait();
  pdl_launch();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  constexpr int D = NV * 128;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * ldx);
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = xr[i * 32 + lane];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float d0 = v[i].x - mean, d1 = v[i].y - mean, d2 = v[i].z - mean, d3 = v[i].w - mean;
    q += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / D) + eps);
  uint2* yr = reinterpret_cast<uint2*>(y + static_cast<size_t>(row) * ldy);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 a4 = reinterpret_cast<const float4*>(a)[i * 32 + lane];
    const float4 b4 = reinterpret_cast<const float4*>(b)[i * 32 + lane];
    const float o0 = (v[i].x - mean) * rstd * (a_off + a4.x) + b4.x;
    const float o1 = (v[i].y - mean) * rstd * (a_off + a4.y) + b4.y;
    const float o2 = (v[i].z - mean) * rstd * (a_off + a4.z) + b4.z;
    const float o3 = (v[i].w - mean) * rstd * (a_off + a4.w) + b4.w;
    if (y != nullptr) {
      if (lo_off > 0) {                                        // split-operand mode: high plane here, low plane lo_off columns on
        uint32_t h0, l0, h1, l1;
        split_bf16x2(o0, o1, h0, l0);
        split_bf16x2(o2, o3, h1, l1);
        yr[i * 32 + lane] = make_uint2(h0, h1);
        reinterpret_cast<uint2*>(y + static_cast<size_t>(row) * ldy + lo_off)[i * 32 + lane] = make_uint2(l0, l1);
      } else {
        yr[i * 32 + lane] = make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
      }
    }
    if (y32 != nullptr)
      reinterpret_cast<float4*>(y32 + static_cast<size_t>(row) * ldy32)[i * 32 + lane] = make_float4(o0, o1, o2, o3);
  }
}

// ------------------------------------------------------------------------------------------------ dwconv(k=7) + LN -> bf16
// One warp per row; a CTA (8 warps) walks blocks of 64 consecutive rows, 8 per warp, so the 6 halo rows of a row are the
// rows its neighbour warps read (L1 hits: HBM sees every row once).  The per-channel taps, conv bias and LN affine are
// staged ONCE per CTA in shared memory, transposed to [tap][channel] so that a lane's four channels are one LDS.128 (the
// first version fetched 7 x 4 scalar weights per float4 of x from global memory on every row: 15 % of HBM bandwidth).
constexpr int DW_ROWS_PER_CTA = 64;
template <int NV>  // C = NV * 128
__global__ void __launch_bounds__(256) dwconv7_ln_kernel(const float* __restrict__ x, long long ldx,
                                                         __nv_bfloat16* __restrict__ y, long long ldy, int M,
                                                         const int* __restrict__ row_pos, const float* __restrict__ w,
                                                         const float* __restrict__ bias, const float* __restrict__ ln_w,
                                                         const float* __restrict__ ln_b, float eps, int lo_off) {
  pdl_wait();
  pdl_launch();
  constexpr int C = NV * 128;
  __shared__ float4 wT[7][NV * 32];
  __shared__ float4 cb[NV * 32], lw[NV * 32], lb[NV * 32];
  for (int c4 = threadIdx.x; c4 < NV * 32; c4 += blockDim.x) {
#pragma unroll
    for (int k = 0; k < 7; ++k)
      wT[k][c4] = make_float4(w[(4 * c4 + 0) * 7 + k], w[(4 * c4 + 1) * 7 + k], w[(4 * c4 + 2) * 7 + k], w[(4 * c4 + 3) * 7 + k]);
    cb[c4] = reinterpret_cast<const float4*>(bias)[c4];
    lw[c4] = reinterpret_cast<const float4*>(ln_w)[c4];
    lb[c4] = reinterpret_cast<const float4*>(ln_b)[c4];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int blk = blockIdx.x; blk * DW_ROWS_PER_CTA < M; blk += gridDim.x) {
#pragma unroll 1
    for (int r = 0; r < DW_ROWS_PER_CTA / 8; ++r) {
      const int row = blk * DW_ROWS_PER_CTA + r * 8 + warp;   // the 8 warps work on 8 ADJACENT rows at a time (halo rows shared in L1)
      if (row >= M) break;
      uint2* yr = reinterpret_cast<uint2*>(y + static_cast<size_t>(row) * ldy);
      const int pos = row_pos[row];
      uint2* yl = reinterpret_cast<uint2*>(y + static_cast<size_t>(row) * ldy + lo_off);   // low plane (split-operand mode)
      if (pos < 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          yr[i * 32 + lane] = make_uint2(0, 0);
          if (lo_off > 0) yl[i * 32 + lane] = make_uint2(0, 0);
        }
        continue;
      }
      float4 acc[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) acc[i] = cb[i * 32 + lane];
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        const int rr = row + k - 3;
        if (rr < 0 || rr >= M) continue;
        const int pr = row_pos[rr];
        if (pr < 0 || pr != pos + k - 3) continue;  // gap row or another utterance => zero padding
        const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(rr) * ldx);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const float4 xv = xr[i * 32 + lane];
          const float4 wv = wT[k][i * 32 + lane];
          acc[i].x = fmaf(xv.x, wv.x, acc[i].x);
          acc[i].y = fmaf(xv.y, wv.y, acc[i].y);
          acc[i].z = fmaf(xv.z, wv.z, acc[i].z);
          acc[i].w = fmaf(xv.w, wv.w, acc[i].w);
        }
      }
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) s += (acc[i].x + acc[i].y) + (acc[i].z + acc[i].w);
      const float mean = warp_sum(s) * (1.f / C);
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const float d0 = acc[i].x - mean, d1 = acc[i].y - mean, d2 = acc[i].z - mean, d3 = acc[i].w - mean;
        q += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
      }
      const float rstd = rsqrtf(warp_sum(q) * (1.f / C) + eps);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const float4 a4 = lw[i * 32 + lane], b4 = lb[i * 32 + lane];
        const float o0 = (acc[i].x - mean) * rstd * a4.x + b4.x, o1 = (acc[i].y - mean) * rstd * a4.y + b4.y;
        const float o2 = (acc[i].z - mean) * rstd * a4.z + b4.z, o3 = (acc[i].w - mean) * rstd * a4.w + b4.w;
        if (lo_off > 0) {
          uint32_t h0, l0, h1, l1;
          split_bf16x2(o0, o1, h0, l0);
          split_bf16x2(o2, o3, h1, l1);
          yr[i * 32 + lane] = make_uint2(h0, h1);
          yl[i * 32 + lane] = make_uint2(l0, l1);
        } else {
          yr[i * 32 + lane] = make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
        }
      }
    }
  }
}

// v2 (default): one warp walks a RUN of consecutive rows with the 7-row window in REGISTERS.  The first form reads every input
// row seven times (once per output row that it is a tap of): HBM sees it once, but the L1 / shared-memory datapath (128 B/clk
// per SM) carries 28 LDG.128 of input + 40 LDS.128 of taps / bias / LN affine = 34 KB per output row, 272 clk against the
// 130 clk the row's 3 KB of HBM traffic needs at 6.5 TB/s: it measured 2.0 TB/s = 31 % of the HBM peak.  Here a row is loaded
// once per run (plus 6 halo rows per run): the window is a ring of 7 register slots rotated by unrolling the row loop 7 times
// (slot indices are compile-time), and the load of row r + 4 is issued right after tap 0 of row r has consumed the slot it
// replaces, two iterations before its first use.  Same operation order as the first form (bias, taps 0..6, two-pass LN):
// bit-identical output.  Measured (262 k rows x 512, same box, profiles/r02_vocos_kernels_ab.txt): 400.6 -> 255.2 us =
// 3.16 TB/s = 48 % of the HBM peak; what is left is the 40 LDS.128 per row (20 KB: 160 clk).
template <int NV>  // C = NV * 128
__global__ void __launch_bounds__(128, 3) dwconv7_ln_run_kernel(const float* __restrict__ x, long long ldx,
                                                                __nv_bfloat16* __restrict__ y, long long ldy, int M,
                                                                const int* __restrict__ row_pos, const float* __restrict__ w,
                                                                const float* __restrict__ bias, const float* __restrict__ ln_w,
                                                                const float* __restrict__ ln_b, float eps, int lo_off, int run_len) {
  pdl_wait();
  pdl_launch();
  constexpr int C = NV * 128;
  __shared__ float4 wT[7][NV * 32];
  __shared__ float4 cb[NV * 32], lw[NV * 32], lb[NV * 32];
  for (int c4 = threadIdx.x; c4 < NV * 32; c4 += blockDim.x) {
#pragma unroll
    for (int k = 0; k < 7; ++k)
      wT[k][c4] = make_float4(w[(4 * c4 + 0) * 7 + k], w[(4 * c4 + 1) * 7 + k], w[(4 * c4 + 2) * 7 + k], w[(4 * c4 + 3) * 7 + k]);
    cb[c4] = reinterpret_cast<const float4*>(bias)[c4];
    lw[c4] = reinterpret_cast<const float4*>(ln_w)[c4];
    lb[c4] = reinterpret_cast<const float4*>(ln_b)[c4];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r0 = (blockIdx.x * 4 + warp) * run_len;
  if (r0 >= M) return;
  const int r1 = r0 + run_len < M ? r0 + run_len : M;
  float4 win[7][NV];
  int pw[7];
  // row rr -> a window slot (data + its position; -1 = outside the matrix).  Gap rows are loaded like any other row (their
  // contents never pass the position test below), so the data load does not wait for the position load.
  auto load_row = [&](int rr, float4 (&dst)[NV], int& p) {
    p = -1;
    if (rr >= 0 && rr < M) {
      p = row_pos[rr];
      const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(rr) * ldx);
#pragma unroll
      for (int i = 0; i < NV; ++i) dst[i] = xr[i * 32 + lane];
    }
  };
#pragma unroll
  for (int s = 0; s < 7; ++s) {
#pragma unroll
    for (int i = 0; i < NV; ++i) win[s][i] = make_float4(0.f, 0.f, 0.f, 0.f);
    load_row(r0 - 3 + s, win[s], pw[s]);
  }
#pragma unroll 1
  for (int base = r0; base < r1; base += 7) {
#pragma unroll
    for (int j = 0; j < 7; ++j) {            // at step j the slot (j + k) % 7 holds row + k - 3
      const int row = base + j;
      if (row < r1) {                        // warp-uniform
        const int pos = pw[(j + 3) % 7];
        uint2* yr = reinterpret_cast<uint2*>(y + static_cast<size_t>(row) * ldy);
        uint2* yl = reinterpret_cast<uint2*>(y + static_cast<size_t>(row) * ldy + lo_off);   // low plane (split-operand mode)
        if (pos < 0) {
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            yr[i * 32 + lane] = make_uint2(0, 0);
            if (lo_off > 0) yl[i * 32 + lane] = make_uint2(0, 0);
          }
          load_row(row + 4, win[j % 7], pw[j % 7]);
        } else {
          float4 acc[NV];
#pragma unroll
          for (int i = 0; i < NV; ++i) acc[i] = cb[i * 32 + lane];
#pragma unroll
          for (int k = 0; k < 7; ++k) {
            const int sl = (j + k) % 7;
            if (pw[sl] >= 0 && pw[sl] == pos + k - 3) {   // same utterance, not a gap row (else: the conv's zero padding)
#pragma unroll
              for (int i = 0; i < NV; ++i) {
                const float4 xv = win[sl][i];
                const float4 wv = wT[k][i * 32 + lane];
                acc[i].x = fmaf(xv.x, wv.x, acc[i].x);
                acc[i].y = fmaf(xv.y, wv.y, acc[i].y);
                acc[i].z = fmaf(xv.z, wv.z, acc[i].z);
                acc[i].w = fmaf(xv.w, wv.w, acc[i].w);
              }
            }
            if (k == 0) load_row(row + 4, win[j % 7], pw[j % 7]);   // the oldest slot is free: next step's newest row
          }
          float s = 0.f;
#pragma unroll
          for (int i = 0; i < NV; ++i) s += (acc[i].x + acc[i].y) + (acc[i].z + acc[i].w);
          const float mean = warp_sum(s) * (1.f / C);
          float q = 0.f;
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            const float d0 = acc[i].x - mean, d1 = acc[i].y - mean, d2 = acc[i].z - mean, d3 = acc[i].w - mean;
            q += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
          }
          const float rstd = rsqrtf(warp_sum(q) * (1.f / C) + eps);
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            const float4 a4 = lw[i * 32 + lane], b4 = lb[i * 32 + lane];
            const float o0 = (acc[i].x - mean) * rstd * a4.x + b4.x, o1 = (acc[i].y - mean) * rstd * a4.y + b4.y;
            const float o2 = (acc[i].z - mean) * rstd * a4.z + b4.z, o3 = (acc[i].w - mean) * rstd * a4.w + b4.w;
            if (lo_off > 0) {
              uint32_t h0, l0, h1, l1;
              split_bf16x2(o0, o1, h0, l0);
              split_bf16x2(o2, o3, h1, l1);
              yr[i * 32 + lane] = make_uint2(h0, h1);
              yl[i * 32 + lane] = make_uint2(l0, l1);
            } else {
              yr[i * 32 + lane] = make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
            }
          }
        }
      }
    }
  }
}

// v3: channel-split form (kept for A/B; f5_set_dwconv7_variant(3)).  It removes the LDS traffic that bounds v2: a warp owns
// 128 CHANNELS (one float4 per lane) of a run of rows, so its 7 taps, bias and LN affine are 40 REGISTERS for the whole run,
// the row window is a ring of 14 float4 (rows r-3 .. r+10: every load is issued 8 rows before its first use) and nothing but
// the input row (LDG.128) and the output (STG.64) touches the LSU.  The C / 128 warps of a CTA walk the same run; LayerNorm
// combines their per-warp (mean, M2) exactly (M2 = sum of the groups' M2 + n (mean_g - mean)^2) through a double-buffered
// 8-byte slot per warp and ONE __syncthreads per row.  FMAs are packed (fma.rn.f32x2).  Measured 294.0 us (2.75 TB/s, 42 %):
// ~165 instructions per warp per row x 4 warps and a barrier + two shuffle reductions on every row's critical path make it
// issue / latency bound below v2; results differ from v1 / v2 by at most one bf16 rounding (tested).
template <int NW>  // C = NW * 128, NW warps per CTA
__global__ void __launch_bounds__(NW * 32, 4) dwconv7_ln_cs_kernel(const float* __restrict__ x, long long ldx,
                                                               __nv_bfloat16* __restrict__ y, long long ldy, int M,
                                                               const int* __restrict__ row_pos, const float* __restrict__ w,
                                                               const float* __restrict__ bias, const float* __restrict__ ln_w,
                                                               const float* __restrict__ ln_b, float eps, int lo_off, int run_len) {
  pdl_wait();
  pdl_launch();
  constexpr int C = NW * 128;
  constexpr int RING = 14;
  __shared__ float2 part[2][NW];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c4 = warp * 32 + lane;                       // this lane's float4 of channels
  float4 wk[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) wk[k] = make_float4(w[(4 * c4 + 0) * 7 + k], w[(4 * c4 + 1) * 7 + k], w[(4 * c4 + 2) * 7 + k], w[(4 * c4 + 3) * 7 + k]);
  const float4 cbv = reinterpret_cast<const float4*>(bias)[c4];
  const float4 lwv = reinterpret_cast<const float4*>(ln_w)[c4];
  const float4 lbv = reinterpret_cast<const float4*>(ln_b)[c4];
  const int r0 = blockIdx.x * run_len;                   // CTA-uniform
  if (r0 >= M) return;
  const int r1 = r0 + run_len < M ? r0 + run_len : M;
  // A window slot = the row's float4 + its KEY: row_pos[rr] - rr for a live row (the same value for every row of one
  // utterance), INT_MIN for gap rows and rows outside the matrix.  Tap k of output row `row` reads rr = row + k - 3 and is
  // inside the utterance iff row_pos[rr] == row_pos[row] + k - 3, i.e. iff key(rr) == key(row); everything else is the
  // conv's zero padding.  Gap rows are loaded like any other row (their contents never pass the key test).
  constexpr int DEAD = -2147483647 - 1;
  float4 win[RING];
  int key[RING];
#pragma unroll
  for (int s = 0; s < RING; ++s) {
    const int rr = r0 - 3 + s;
    win[s] = make_float4(0.f, 0.f, 0.f, 0.f);
    key[s] = DEAD;
    if (rr >= 0 && rr < M) {
      const int p = row_pos[rr];
      key[s] = p >= 0 ? p - rr : DEAD;
      win[s] = reinterpret_cast<const float4*>(x + static_cast<size_t>(rr) * ldx)[c4];
    }
  }
  int rn = r0 - 3 + RING;                                 // the next row to load (>= 0), its data and position pointers
  const float* xn = x + static_cast<size_t>(rn) * ldx + 4 * c4;
  const int* pn = row_pos + rn;
  __nv_bfloat16* yo = y + static_cast<size_t>(r0) * ldy + 4 * c4;
  auto load_next = [&](float4& dst, int& kdst) {
    kdst = DEAD;
    if (rn < M) {
      const int p = *pn;
      dst = *reinterpret_cast<const float4*>(xn);
      kdst = p >= 0 ? p - rn : DEAD;
    }
    ++rn;
    ++pn;
    xn += ldx;
  };
  int buf = 0;
#pragma unroll 1
  for (int base = r0; base < r1; base += RING) {
#pragma unroll
    for (int j = 0; j < RING; ++j) {                     // at step j the slot (j + k) % RING holds row + k - 3
      if (base + j < r1) {                               // CTA-uniform
        const int kc = key[(j + 3) % RING];              // CTA-uniform (a property of the row)
        if (kc == DEAD) {
          *reinterpret_cast<uint2*>(yo) = make_uint2(0, 0);
          if (lo_off > 0) *reinterpret_cast<uint2*>(yo + lo_off) = make_uint2(0, 0);
          load_next(win[j % RING], key[j % RING]);
        } else {
          float2 a01 = make_float2(cbv.x, cbv.y), a23 = make_float2(cbv.z, cbv.w);
#pragma unroll
          for (int k = 0; k < 7; ++k) {
            const int sl = (j + k) % RING;
            if (key[sl] == kc) {
              a01 = ffma2(make_float2(win[sl].x, win[sl].y), make_float2(wk[k].x, wk[k].y), a01);
              a23 = ffma2(make_float2(win[sl].z, win[sl].w), make_float2(wk[k].z, wk[k].w), a23);
            }
            if (k == 0) load_next(win[j % RING], key[j % RING]);   // the oldest slot is free: row + RING - 3 goes there
          }
          // this warp's 128 channels: mean and M2 = sum (v - mean)^2; then Chan's combination over the NW warps
          const float mw = warp_sum((a01.x + a01.y) + (a23.x + a23.y)) * (1.f / 128.f);
          const float d0 = a01.x - mw, d1 = a01.y - mw, d2 = a23.x - mw, d3 = a23.y - mw;
          const float m2w = warp_sum((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3));
          float mean = mw, m2 = m2w;
          if (NW > 1) {
            if (lane == 0) part[buf][warp] = make_float2(mw, m2w);
            __syncthreads();                             // one barrier per row: the slots alternate, see above
            float2 pr[NW];
            mean = 0.f;
#pragma unroll
            for (int i = 0; i < NW; ++i) {
              pr[i] = part[buf][i];
              mean += pr[i].x;
            }
            mean *= (1.f / NW);
            m2 = 0.f;
#pragma unroll
            for (int i = 0; i < NW; ++i) {
              const float dm = pr[i].x - mean;
              m2 += pr[i].y + 128.f * dm * dm;
            }
            buf ^= 1;
          }
          const float rstd = rsqrtf(m2 * (1.f / C) + eps);
          const float o0 = (a01.x - mean) * rstd * lwv.x + lbv.x, o1 = (a01.y - mean) * rstd * lwv.y + lbv.y;
          const float o2 = (a23.x - mean) * rstd * lwv.z + lbv.z, o3 = (a23.y - mean) * rstd * lwv.w + lbv.w;
          if (lo_off > 0) {
            uint32_t h0, l0, h1, l1;
            split_bf16x2(o0, o1, h0, l0);
            split_bf16x2(o2, o3, h1, l1);
            *reinterpret_cast<uint2*>(yo) = make_uint2(h0, h1);
            *reinterpret_cast<uint2*>(yo + lo_off) = make_uint2(l0, l1);
          } else {
            *reinterpret_cast<uint2*>(yo) = make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
          }
        }
        yo += ldy;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ GRN
constexpr int GRN_ROWS = 32;  // rows per block
// Deterministic (fixed summation order, no atomics): grid (C/128, num_segs); 256 threads = 64 channel pairs x 4 row
// slices; each thread walks its slice of the utterance's rows in order, the 4 slices are combined in a fixed order.
__global__ void __launch_bounds__(256) grn_sumsq_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, int C,
                                                        const int* __restrict__ seg_rows, float* __restrict__ sumsq) {
  pdl_wait();
  pdl_launch();
  __shared__ float2 part[4][64];
  const int seg = blockIdx.y;
  const int row0 = seg_rows[2 * seg], n = seg_rows[2 * seg + 1];
  const int cp = threadIdx.x & 63, slice = threadIdx.x >> 6;
  const int c2 = blockIdx.x * 64 + cp;                 // bf16x2 index
  float s0 = 0.f, s1 = 0.f;
  if (2 * c2 < C) {
    for (int r = slice; r < n; r += 4) {
      const float2 f = __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(x + static_cast<size_t>(row0 + r) * ldx)[c2]);
      s0 = fmaf(f.x, f.x, s0);
      s1 = fmaf(f.y, f.y, s1);
    }
  }
  part[slice][cp] = make_float2(s0, s1);
  __syncthreads();
  if (slice == 0 && 2 * c2 < C) {
    const float2 a = part[0][cp], b = part[1][cp], c = part[2][cp], d = part[3][cp];
    sumsq[static_cast<size_t>(seg) * C + 2 * c2] = (a.x + b.x) + (c.x + d.x);
    sumsq[static_cast<size_t>(seg) * C + 2 * c2 + 1] = (a.y + b.y) + (c.y + d.y);
  }
}

__global__ void grn_apply_kernel(__nv_bfloat16* __restrict__ x, long long ldx, int C, const int* __restrict__ seg_rows,
                                 const float* __restrict__ sumsq, const float* __restrict__ gamma,
                                 const float* __restrict__ beta) {
  pdl_wait();
  pdl_launch();
  __shared__ float red[32];
  __shared__ float mean_gx;
  const int seg = blockIdx.y;
  const int row0 = seg_rows[2 * seg], n = seg_rows[2 * seg + 1];
  const int r0 = blockIdx.x * GRN_ROWS;
  if (r0 >= n) return;
  const int r1 = min(n, r0 + GRN_ROWS);
  float part = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) part += sqrtf(sumsq[static_cast<size_t>(seg) * C + c]);
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x + 31) / 32 ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) mean_gx = t / C;
  }
  __syncthreads();
  const float inv = 1.f / (mean_gx + 1e-6f);
  for (int c2 = threadIdx.x; c2 < C / 2; c2 += blockDim.x) {
    const float n0 = sqrtf(sumsq[static_cast<size_t>(seg) * C + 2 * c2]) * inv;
    const float n1 = sqrtf(sumsq[static_cast<size_t>(seg) * C + 2 * c2 + 1]) * inv;
    const float g0 = gamma[2 * c2], g1 = gamma[2 * c2 + 1], b0 = beta[2 * c2], b1 = beta[2 * c2 + 1];
    for (int r = r0; r < r1; ++r) {
      __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(x + static_cast<size_t>(row0 + r) * ldx) + c2;
      const float2 f = __bfloat1622float2(*p);
      *p = __floats2bfloat162_rn(g0 * (f.x * n0) + b0 + f.x, g1 * (f.y * n1) + b1 + f.y);
    }
  }
}

// ------------------------------------------------------------------------------------------------ gather / pack / misc
// fp32 forms of the two GRN passes (fp32 precision mode: the pointwise GEMM's GELU output stays fp32 until it is split into
// bf16 planes for the next GEMM); same fixed summation order as the bf16 kernels
__global__ void __launch_bounds__(256) grn_sumsq_f32_kernel(const float* __restrict__ x, long long ldx, int C,
                                                            const int* __restrict__ seg_rows, float* __restrict__ sumsq) {
  pdl_wait();
  pdl_launch();
  __shared__ float2 part[4][64];
  const int seg = blockIdx.y;
  const int row0 = seg_rows[2 * seg], n = seg_rows[2 * seg + 1];
  const int cp = threadIdx.x & 63, slice = threadIdx.x >> 6;
  const int c2 = blockIdx.x * 64 + cp;
  float s0 = 0.f, s1 = 0.f;
  if (2 * c2 < C) {
    for (int r = slice; r < n; r += 4) {
      const float2 f = reinterpret_cast<const float2*>(x + static_cast<size_t>(row0 + r) * ldx)[c2];
      s0 = fmaf(f.x, f.x, s0);
      s1 = fmaf(f.y, f.y, s1);
    }
  }
  part[slice][cp] = make_float2(s0, s1);
  __syncthreads();
  if (slice == 0 && 2 * c2 < C) {
    const float2 a = part[0][cp], b = part[1][cp], c = part[2][cp], d = part[3][cp];
    sumsq[static_cast<size_t>(seg) * C + 2 * c2] = (a.x + b.x) + (c.x + d.x);
    sumsq[static_cast<size_t>(seg) * C + 2 * c2 + 1] = (a.y + b.y) + (c.y + d.y);
  }
}

__global__ void grn_apply_f32_kernel(float* __restrict__ x, long long ldx, int C, const int* __restrict__ seg_rows,
                                     const float* __restrict__ sumsq, const float* __restrict__ gamma,
                                     const float* __restrict__ beta) {
  pdl_wait();
  pdl_launch();
  __shared__ float red[32];
  __shared__ float mean_gx;
  const int seg = blockIdx.y;
  const int row0 = seg_rows[2 * seg], n = seg_rows[2 * seg + 1];
  const int r0 = blockIdx.x * GRN_ROWS;
  if (r0 >= n) return;
  const int r1 = min(n, r0 + GRN_ROWS);
  float part = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) part += sqrtf(sumsq[static_cast<size_t>(seg) * C + c]);
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x + 31) / 32 ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) mean_gx = t / C;
  }
  __syncthreads();
  const float inv = 1.f / (mean_gx + 1e-6f);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float nx = sqrtf(sumsq[static_cast<size_t>(seg) * C + c]) * inv;
    const float g = gamma[c], b = beta[c];
    for (int r = r0; r < r1; ++r) {
      float* p = x + static_cast<size_t>(row0 + r) * ldx + c;
      const float f = *p;
      *p = g * (f * nx) + b + f;
    }
  }
}

__global__ void text_gather_pos_kernel(const int* __restrict__ ids, const int* __restrict__ row_pos,
                                       const float* __restrict__ emb, const float* __restrict__ pos_table, int max_pos,
                                       float* __restrict__ out, long long ldo, int M, int C) {
  pdl_wait();
  pdl_launch();
  const int row = blockIdx.x;
  if (row >= M) return;
  const int pos = row_pos[row];
  float4* o = reinterpret_cast<float4*>(out + static_cast<size_t>(row) * ldo);
  if (pos < 0) {
    for (int c = threadIdx.x; c < C / 4; c += blockDim.x) o[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const float4* e = reinterpret_cast<const float4*>(emb + static_cast<size_t>(ids[row]) * C);
  const float4* pt = reinterpret_cast<const float4*>(pos_table + static_cast<size_t>(min(pos, max_pos - 1)) * C);
  for (int c = threadIdx.x; c < C / 4; c += blockDim.x) {
    const float4 a = e[c], b = pt[c];
    o[c] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  }
}

// one warp per row; C_pad % 8 == 0
__global__ void __launch_bounds__(256) pack_bf16_kernel(const float* __restrict__ src, long long lds,
                                                        __nv_bfloat16* __restrict__ dst, long long ldd, int dst_col,
                                                        int M, int C, int C_pad, const int* __restrict__ src_rows,
                                                        const int* __restrict__ row_pos, int lo_off) {
  pdl_wait();
  pdl_launch();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  int srow = src_rows != nullptr ? src_rows[row] : row;
  if (row_pos != nullptr && row_pos[row] < 0) srow = -1;
  __nv_bfloat16* d = dst + static_cast<size_t>(row) * ldd + dst_col;
  const float* s = src + static_cast<size_t>(srow < 0 ? 0 : srow) * lds;
  for (int c = lane * 2; c < C_pad; c += 64) {
    const float v0 = (srow >= 0 && c < C) ? s[c] : 0.f;
    const float v1 = (srow >= 0 && c + 1 < C) ? s[c + 1] : 0.f;
    if (lo_off > 0) {
      uint32_t h, l;
      split_bf16x2(v0, v1, h, l);
      *reinterpret_cast<uint32_t*>(d + c) = h;
      *reinterpret_cast<uint32_t*>(d + lo_off + c) = l;
    } else {
      *reinterpret_cast<uint32_t*>(d + c) = pack_bf16x2(v0, v1);
    }
  }
}

__global__ void where_rows_kernel(float* __restrict__ x, long long ldx, const float* __restrict__ c, long long ldc,
                                  const int* __restrict__ flag, int M, int C) {
  pdl_wait();
  pdl_launch();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= M || flag[row] == 0) return;
  for (int i = threadIdx.x & 31; i < C; i += 32) x[static_cast<size_t>(row) * ldx + i] = c[static_cast<size_t>(row) * ldc + i];
}

// one warp per token row of the conditional half
__global__ void __launch_bounds__(256) cfg_euler_kernel(float* __restrict__ x, long long ldx, const float* __restrict__ pred,
                                                        long long ldp, int half_rows, int C, const int* __restrict__ row_pos,
                                                        const float* __restrict__ dts, int step, float cfg,
                                                        __nv_bfloat16* __restrict__ xb, long long ldxb, int C_pad) {
  pdl_wait();
  pdl_launch();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= half_rows) return;
  const int lane = threadIdx.x & 31;
  const bool live = row_pos[row] >= 0;
  const float dt = dts[step];
  float* xr = x + static_cast<size_t>(row) * ldx;
  const float* pc = pred + static_cast<size_t>(row) * ldp;
  const float* pu = pred + static_cast<size_t>(row + half_rows) * ldp;
  __nv_bfloat16* b0 = xb + static_cast<size_t>(row) * ldxb;
  __nv_bfloat16* b1 = xb + static_cast<size_t>(row + half_rows) * ldxb;
  for (int c = lane * 2; c < C_pad; c += 64) {
    float v0 = 0.f, v1 = 0.f;
    if (live && c < C) {
      const float a = pc[c], u = pu[c];
      v0 = xr[c] + dt * (a + (a - u) * cfg);
      xr[c] = v0;
    }
    if (live && c + 1 < C) {
      const float a = pc[c + 1], u = pu[c + 1];
      v1 = xr[c + 1] + dt * (a + (a - u) * cfg);
      xr[c + 1] = v1;
    }
    const uint32_t w = pack_bf16x2(v0, v1);
    *reinterpret_cast<uint32_t*>(b0 + c) = w;
    *reinterpret_cast<uint32_t*>(b1 + c) = w;
  }
}

// ------------------------------------------------------------------------------------------------ initial noise
// y0 = randn(n_i, C) per utterance (reference model/cfm.py:181-186 draws it with torch.randn on the model's device).  Counter-
// based so that the value at (utterance seed, frame, channel) does not depend on how the batch is packed: Philox4x32-10 keyed
// by the utterance's 64-bit seed, counter = (frame * 32 + lane, 0x4635, 0, 0); each draw gives four 32-bit words = two Box-
// Muller pairs = channels 4*lane .. 4*lane+3.  One warp per row; gap rows (row_pos < 0) are zero-filled.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  const float u1 = (static_cast<float>(a >> 8) + 1.0f) * 5.9604644775390625e-8f;   // (0, 1]
  const float u2 = static_cast<float>(b >> 8) * 5.9604644775390625e-8f;            // [0, 1)
  const float r = sqrtf(-2.0f * logf(u1));
  float sn, cs;
  sincospif(2.0f * u2, &sn, &cs);
  return make_float2(r * cs, r * sn);
}
__global__ void __launch_bounds__(256) randn_rows_kernel(float* __restrict__ x, long long ldx, int M, int C,
                                                         const int* __restrict__ row_pos, const int* __restrict__ row_utt,
                                                         const unsigned long long* __restrict__ utt_seed) {
  pdl_wait();
  pdl_launch();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const int pos = row_pos[row];
  float* xr = x + static_cast<size_t>(row) * ldx;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (pos >= 0 && 4 * lane < C) {
    const unsigned long long seed = utt_seed[row_utt[row]];
    uint32_t w[4];
    philox4x32_10(static_cast<uint32_t>(pos) * 32u + lane, 0x4635u, 0u, 0u, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), w);
    const float2 a = box_muller(w[0], w[1]), b = box_muller(w[2], w[3]);
    v = make_float4(a.x, a.y, b.x, b.y);
  }
  if (4 * lane + 3 < C) {
    *reinterpret_cast<float4*>(xr + 4 * lane) = v;
  } else {
    const float e[4] = {v.x, v.y, v.z, v.w};
    for (int i = 0; i < 4; ++i)
      if (4 * lane + i < C) xr[4 * lane + i] = e[i];
  }
}

__global__ void time_sinus_kernel(const float* __restrict__ t, int steps, const float* __restrict__ freqs, int dim,
                                  __nv_bfloat16* __restrict__ out, long long ldo, int lo_off) {
  pdl_wait();
  pdl_launch();
  const int s = blockIdx.x;
  const int half = dim / 2;
  for (int k = threadIdx.x; k < half; k += blockDim.x) {
    const float arg = (1000.f * t[s]) * freqs[k];
    const float sn = sinf(arg), cs = cosf(arg);
    const __nv_bfloat16 hs = __float2bfloat16(sn), hc = __float2bfloat16(cs);
    out[static_cast<size_t>(s) * ldo + k] = hs;
    out[static_cast<size_t>(s) * ldo + half + k] = hc;
    if (lo_off > 0) {
      out[static_cast<size_t>(s) * ldo + lo_off + k] = __float2bfloat16(sn - __bfloat162float(hs));
      out[static_cast<size_t>(s) * ldo + lo_off + half + k] = __float2bfloat16(cs - __bfloat162float(hc));
    }
  }
}

// split-operand form: x [rows, cols] contiguous -> y [rows, 2 * cols]: high plane | low plane
__global__ void silu_split_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n, int cols) {
  pdl_wait();
  pdl_launch();
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 2;
  if (i >= n) return;
  const float2 v = *reinterpret_cast<const float2*>(x + i);
  uint32_t h, l;
  split_bf16x2(silu(v.x), silu(v.y), h, l);
  const long long r = i / cols, c = i % cols;
  *reinterpret_cast<uint32_t*>(y + r * 2 * cols + c) = h;
  *reinterpret_cast<uint32_t*>(y + r * 2 * cols + cols + c) = l;
}

__global__ void silu_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n) {
  pdl_wait();
  pdl_launch();
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 2;
  if (i + 1 < n) {
    const float2 v = *reinterpret_cast<const float2*>(x + i);
    *reinterpret_cast<uint32_t*>(y + i) = pack_bf16x2(silu(v.x), silu(v.y));
  } else if (i < n) {
    y[i] = __float2bfloat16(silu(x[i]));
  }
}

}  // namespace f5

using namespace f5;
#define F5_STREAM(s) reinterpret_cast<cudaStream_t>(s)
#define F5_LAUNCH_RC() static_cast<int>(cudaGetLastError())

extern "C" int f5_layernorm_mod(const float* x, int64_t ldx, void* y, int64_t ldy, float* y32, int64_t ldy32, int32_t M,
                                int32_t D, const float* a, const float* b, float a_off, float eps, int32_t lo_off, void* stream) {
  if (!x || (!y && !y32) || !a || !b || M <= 0 || D % 128 != 0 || D > 1024 || ldx % 4 != 0 || ldy % 4 != 0 || ldy32 % 4 != 0 ||
      lo_off < 0 || lo_off % 4 != 0)
    return F5_ERR_ARG;
  const int grid = (M + 7) / 8;
  __nv_bfloat16* yo = reinterpret_cast<__nv_bfloat16*>(y);
  switch (D / 128) {
#define F5_CASE(NV) case NV: f5_launch(layernorm_mod_kernel<NV>, dim3(grid), dim3(256), 0, F5_STREAM(stream), x, ldx, yo, ldy, y32, ldy32, M, a, b, a_off, eps, lo_off); break;
    F5_CASE(1) F5_CASE(2) F5_CASE(3) F5_CASE(4) F5_CASE(5) F5_CASE(6) F5_CASE(7) F5_CASE(8)
#undef F5_CASE
  }
  return F5_LAUNCH_RC();
}

static int f5_dwconv7_variant = [] { const char* e = getenv("F5_DWCONV_V"); return (e != nullptr && e[0] >= '1' && e[0] <= '3') ? e[0] - '0' : 2; }();
extern "C" int f5_set_dwconv7_variant(int v) {
  const int old = f5_dwconv7_variant;
  if (v >= 1 && v <= 3) f5_dwconv7_variant = v;
  return old;
}

extern "C" int f5_dwconv7_ln(const float* x, int64_t ldx, void* y, int64_t ldy, int32_t M, int32_t C, const int32_t* row_pos,
                             const float* w, const float* bias, const float* ln_w, const float* ln_b, float eps,
                             int32_t lo_off, void* stream) {
  if (!x || !y || !row_pos || !w || !bias || !ln_w || !ln_b || M <= 0 || C % 128 != 0 || C > 512 || ldx % 4 != 0 || ldy % 4 != 0 ||
      lo_off < 0 || lo_off % 4 != 0)
    return F5_ERR_ARG;
  __nv_bfloat16* yo = reinterpret_cast<__nv_bfloat16*>(y);
  // Three forms (see the kernels): 2 = one warp per run of rows, window in registers (default: fastest, bit-identical to 1);
  // 3 = channel-split, window and taps in registers; 1 = one warp per row, halo through L1.  F5_DWCONV_V in the environment /
  // f5_set_dwconv7_variant select one (A/B measurements).
  const int variant = f5_dwconv7_variant;
  if (variant == 3) {
    // one CTA (C / 128 warps) per run of rows, 4 CTAs per SM resident; runs are a multiple of the 14-slot window ring
    const int slots = kNumSMsB200 * 4;
    int run = ((M + slots - 1) / slots + 13) / 14 * 14;
    if (run < 14) run = 14;
    const int grid3 = (M + run - 1) / run;
    switch (C / 128) {
#define F5_CASE(NW) case NW: f5_launch(dwconv7_ln_cs_kernel<NW>, dim3(grid3), dim3(NW * 32), 0, F5_STREAM(stream), x, ldx, yo, ldy, M, row_pos, w, bias, ln_w, ln_b, eps, lo_off, run); break;
      F5_CASE(1) F5_CASE(2) F5_CASE(3) F5_CASE(4)
#undef F5_CASE
    }
    return F5_LAUNCH_RC();
  }
  if (variant == 2) {
    // one run per warp where there are enough rows (12 warps per SM resident), a multiple of 7 rows (the window rotation)
    const int slots = kNumSMsB200 * 12;
    int run = ((M + slots - 1) / slots + 6) / 7 * 7;
    if (run < 7) run = 7;
    const int grid2 = ((M + run - 1) / run + 3) / 4;
    switch (C / 128) {
#define F5_CASE(NV) case NV: f5_launch(dwconv7_ln_run_kernel<NV>, dim3(grid2), dim3(128), 0, F5_STREAM(stream), x, ldx, yo, ldy, M, row_pos, w, bias, ln_w, ln_b, eps, lo_off, run); break;
      F5_CASE(1) F5_CASE(2) F5_CASE(3) F5_CASE(4)
#undef F5_CASE
    }
    return F5_LAUNCH_RC();
  }
  const int blocks = (M + DW_ROWS_PER_CTA - 1) / DW_ROWS_PER_CTA;
  const int grid = blocks < kNumSMsB200 * 8 ? blocks : kNumSMsB200 * 8;
  switch (C / 128) {
#define F5_CASE(NV) case NV: f5_launch(dwconv7_ln_kernel<NV>, dim3(grid), dim3(256), 0, F5_STREAM(stream), x, ldx, yo, ldy, M, row_pos, w, bias, ln_w, ln_b, eps, lo_off); break;
    F5_CASE(1) F5_CASE(2) F5_CASE(3) F5_CASE(4)
#undef F5_CASE
  }
  return F5_LAUNCH_RC();
}

static int max_seg_rows_hint = 4096;  // segments never exceed the reference's 4096-frame clamp (model/cfm.py:93,137)

extern "C" int f5_grn_sumsq(const void* x, int64_t ldx, int32_t C, const int32_t* seg_rows, int32_t num_segs, float* sumsq,
                            void* stream) {
  if (!x || !seg_rows || !sumsq || num_segs <= 0 || C % 2 != 0 || ldx % 2 != 0) return F5_ERR_ARG;
  dim3 grid((C + 127) / 128, num_segs);
  f5_launch(grn_sumsq_kernel, dim3(grid), dim3(256), 0, F5_STREAM(stream), reinterpret_cast<const __nv_bfloat16*>(x), ldx, C, seg_rows, sumsq);
  return F5_LAUNCH_RC();
}

extern "C" int f5_grn_apply(void* x, int64_t ldx, int32_t C, const int32_t* seg_rows, int32_t num_segs, const float* sumsq,
                            const float* gamma, const float* beta, void* stream) {
  if (!x || !seg_rows || !sumsq || !gamma || !beta || num_segs <= 0 || C % 2 != 0 || ldx % 2 != 0) return F5_ERR_ARG;
  dim3 grid((max_seg_rows_hint + GRN_ROWS - 1) / GRN_ROWS, num_segs);
  f5_launch(grn_apply_kernel, dim3(grid), dim3(256), 0, F5_STREAM(stream), reinterpret_cast<__nv_bfloat16*>(x), ldx, C, seg_rows, sumsq, gamma, beta);
  return F5_LAUNCH_RC();
}

extern "C" int f5_grn_sumsq_f32(const float* x, int64_t ldx, int32_t C, const int32_t* seg_rows, int32_t num_segs, float* sumsq,
                                void* stream) {
  if (!x || !seg_rows || !sumsq || num_segs <= 0 || C % 2 != 0 || ldx % 2 != 0) return F5_ERR_ARG;
  dim3 grid((C + 127) / 128, num_segs);
  f5_launch(grn_sumsq_f32_kernel, dim3(grid), dim3(256), 0, F5_STREAM(stream), x, ldx, C, seg_rows, sumsq);
  return F5_LAUNCH_RC();
}

extern "C" int f5_grn_apply_f32(float* x, int64_t ldx, int32_t C, const int32_t* seg_rows, int32_t num_segs, const float* sumsq,
                                const float* gamma, const float* beta, void* stream) {
  if (!x || !seg_rows || !sumsq || !gamma || !beta || num_segs <= 0) return F5_ERR_ARG;
  dim3 grid((max_seg_rows_hint + GRN_ROWS - 1) / GRN_ROWS, num_segs);
  f5_launch(grn_apply_f32_kernel, dim3(grid), dim3(256), 0, F5_STREAM(stream), x, ldx, C, seg_rows, sumsq, gamma, beta);
  return F5_LAUNCH_RC();
}

extern "C" int f5_text_gather_pos(const int32_t* ids, const int32_t* row_pos, const float* emb, const float* pos_table,
                                  int32_t max_pos, float* out, int64_t ldo, int32_t M, int32_t C, void* stream) {
  if (!ids || !row_pos || !emb || !pos_table || !out || M <= 0 || C % 4 != 0 || ldo % 4 != 0) return F5_ERR_ARG;
  f5_launch(text_gather_pos_kernel, dim3(M), dim3(128), 0, F5_STREAM(stream), ids, row_pos, emb, pos_table, max_pos, out, ldo, M, C);
  return F5_LAUNCH_RC();
}

extern "C" int f5_pack_bf16(const float* src, int64_t lds, void* dst, int64_t ldd, int32_t dst_col, int32_t M, int32_t C,
                            int32_t C_pad, const int32_t* src_rows, const int32_t* row_pos, int32_t lo_off, void* stream) {
  if (!src || !dst || M <= 0 || C <= 0 || C_pad < C || C_pad % 2 != 0 || dst_col % 2 != 0 || ldd % 2 != 0 || lo_off < 0 || lo_off % 2 != 0)
    return F5_ERR_ARG;
  f5_launch(pack_bf16_kernel, dim3((M + 7) / 8), dim3(256), 0, F5_STREAM(stream), src, lds, reinterpret_cast<__nv_bfloat16*>(dst), ldd, dst_col,
                                                               M, C, C_pad, src_rows, row_pos, lo_off);
  return F5_LAUNCH_RC();
}

extern "C" int f5_where_rows(float* x, int64_t ldx, const float* c, int64_t ldc, const int32_t* flag, int32_t M, int32_t C,
                             void* stream) {
  if (!x || !c || !flag || M <= 0 || C <= 0) return F5_ERR_ARG;
  f5_launch(where_rows_kernel, dim3((M + 7) / 8), dim3(256), 0, F5_STREAM(stream), x, ldx, c, ldc, flag, M, C);
  return F5_LAUNCH_RC();
}

extern "C" int f5_cfg_euler(float* x, int64_t ldx, const float* pred, int64_t ldp, int32_t half_rows, int32_t C,
                            const int32_t* row_pos, const float* dts, int32_t step, float cfg_strength, void* xb,
                            int64_t ldxb, int32_t C_pad, void* stream) {
  if (!x || !pred || !row_pos || !dts || !xb || half_rows <= 0 || C <= 0 || C_pad < C || C_pad % 2 != 0 || ldxb % 2 != 0)
    return F5_ERR_ARG;
  f5_launch(cfg_euler_kernel, dim3((half_rows + 7) / 8), dim3(256), 0, F5_STREAM(stream), x, ldx, pred, ldp, half_rows, C, row_pos, dts, step,
                                                                       cfg_strength, reinterpret_cast<__nv_bfloat16*>(xb),
                                                                       ldxb, C_pad);
  return F5_LAUNCH_RC();
}

extern "C" int f5_randn_rows(float* x, int64_t ldx, int32_t M, int32_t C, const int32_t* row_pos, const int32_t* row_utt,
                             const uint64_t* utt_seed, void* stream) {
  if (!x || !row_pos || !row_utt || !utt_seed || M <= 0 || C <= 0 || C > 128 || ldx % 4 != 0) return F5_ERR_ARG;
  f5_launch(randn_rows_kernel, dim3((M + 7) / 8), dim3(256), 0, F5_STREAM(stream), x, ldx, M, C, row_pos, row_utt,
                                                                reinterpret_cast<const unsigned long long*>(utt_seed));
  return F5_LAUNCH_RC();
}

extern "C" int f5_time_sinus(const float* t, int32_t steps, const float* freqs, int32_t dim, void* out, int64_t ldo,
                             int32_t lo_off, void* stream) {
  if (!t || !freqs || !out || steps <= 0 || dim <= 0 || dim % 2 != 0 || lo_off < 0) return F5_ERR_ARG;
  f5_launch(time_sinus_kernel, dim3(steps), dim3(128), 0, F5_STREAM(stream), t, steps, freqs, dim, reinterpret_cast<__nv_bfloat16*>(out), ldo, lo_off);
  return F5_LAUNCH_RC();
}

extern "C" int f5_silu_bf16(const float* x, void* out, int64_t n, int32_t split_cols, void* stream) {
  if (!x || !out || n <= 0 || split_cols < 0) return F5_ERR_ARG;
  const long long pairs = (n + 1) / 2;
  if (split_cols > 0) {
    if (split_cols % 2 != 0 || n % split_cols != 0) return F5_ERR_ARG;
    f5_launch(silu_split_kernel, dim3(static_cast<unsigned>((pairs + 255) / 256)), dim3(256), 0, F5_STREAM(stream), x,
              reinterpret_cast<__nv_bfloat16*>(out), n, split_cols);
    return F5_LAUNCH_RC();
  }
  f5_launch(silu_bf16_kernel, dim3(static_cast<unsigned>((pairs + 255) / 256)), dim3(256), 0, F5_STREAM(stream), 
      x, reinterpret_cast<__nv_bfloat16*>(out), n);
  return F5_LAUNCH_RC();
}

extern "C" int f5_device_check(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return F5_ERR_ARCH;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return F5_ERR_ARCH;
  return (prop.major == 10 && prop.minor == 0) ? F5_OK : F5_ERR_ARCH;
}

extern "C" const char* f5_version(void) { return "f5_b200 0.2 (sm_100a; tcgen05/TMEM/TMA)"; }
