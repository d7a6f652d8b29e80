// fp32 attention for the fp32 precision mode (reference: F.scaled_dot_product_attention in fp32 at f5_tts/model/modules.py:436
// with x-transformers' rotary embedding on head 0, :414-426).  The served bf16 path runs attn_tcgen05.cu; this kernel exists so
// that the fp32 mode's parity is not limited by bf16 Q/K/V/P operands: every product, the softmax and the accumulation are fp32
// on the CUDA cores.  It is a parity instrument, not a throughput kernel (the fp32 mode is reported, never the headline).
//
// One CTA = 64 query rows of one work item (a work item is the bf16 kernel's: up to 256 query rows of one utterance) x one head.
// Four threads share a query row: each holds the whole (pre-scaled, rotated) q row in registers, scores 16 of the 64 keys of a
// tile (keys c, c+4, ...), and accumulates a 16-wide slice of the output (dims 16 i + 4 c + e), taking the other threads'
// probabilities by shuffle.  K / V tiles of 64 keys are staged in shared memory (K rows padded to 68 floats: conflict-free
// 16-byte reads by the four key owners).  Online softmax in the log2 domain.
#define F5_DIAG_TAG 3u
#include "f5_common.cuh"
#include "../../include/f5_b200.h"

namespace f5 {

constexpr int AF_ROWS = 64, AF_KEYS = 64, AF_D = 64, AF_KPAD = 68;

__device__ __forceinline__ float4 rope_rotate(float4 x, float4 t) {   // t = (cos a, sin a, cos b, sin b) of the two pairs
  return make_float4(x.x * t.x - x.y * t.y, x.y * t.x + x.x * t.y, x.z * t.z - x.w * t.w, x.w * t.z + x.z * t.w);
}

__global__ void __launch_bounds__(256) attn_f32_kernel(const float* __restrict__ qkv, long long ld, int q_col, int k_col, int v_col,
                                                       int heads, const int* __restrict__ items, const float* __restrict__ rope,
                                                       __nv_bfloat16* __restrict__ out, long long ldo, int lo_off, float scale_log2) {
  pdl_wait();
  pdl_launch();
  __shared__ __align__(16) float sK[AF_KEYS][AF_KPAD];
  __shared__ __align__(16) float sV[AF_KEYS][AF_D];
  const int sub = blockIdx.x & 3, head = (blockIdx.x >> 2) % heads, item = (blockIdx.x >> 2) / heads;
  const int4 it = *reinterpret_cast<const int4*>(items + 4 * item);   // q_row0, kv_row0, kv_len, q_rows_valid
  if (it.w <= 0 || it.z <= 0 || sub * AF_ROWS >= it.w) return;         // padding item / no rows in this quarter (uniform per CTA)
  const int t = threadIdx.x, r = t >> 2, c = t & 3, lane = t & 31;
  const bool row_ok = sub * AF_ROWS + r < it.w;
  const int qrow = it.x + sub * AF_ROWS + (row_ok ? r : 0);
  const bool rot = head == 0 && rope != nullptr;                      // RoPE touches head 0 only (modules.py:418-426)

  float q[AF_D];
  {
    const float4* qp = reinterpret_cast<const float4*>(qkv + static_cast<size_t>(qrow) * ld + q_col + head * AF_D);
    const float4* rp = reinterpret_cast<const float4*>(rope + static_cast<size_t>(rot ? qrow - it.y : 0) * AF_D);
#pragma unroll
    for (int i = 0; i < AF_D / 4; ++i) {
      float4 v = qp[i];
      if (rot) v = rope_rotate(v, rp[i]);
      q[4 * i] = v.x * scale_log2; q[4 * i + 1] = v.y * scale_log2; q[4 * i + 2] = v.z * scale_log2; q[4 * i + 3] = v.w * scale_log2;
    }
  }
  float m = -INFINITY, l = 0.f;
  float o[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) o[i] = 0.f;

  for (int j0 = 0; j0 < it.z; j0 += AF_KEYS) {
    __syncthreads();                                                  // the previous tile has been consumed
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = t + 256 * u, key = idx >> 4, f4 = idx & 15;
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (j0 + key < it.z) {
        const float* base = qkv + static_cast<size_t>(it.y + j0 + key) * ld + head * AF_D + 4 * f4;
        kv = *reinterpret_cast<const float4*>(base + k_col);
        vv = *reinterpret_cast<const float4*>(base + v_col);
        if (rot) kv = rope_rotate(kv, reinterpret_cast<const float4*>(rope + static_cast<size_t>(j0 + key) * AF_D)[f4]);
      }
      *reinterpret_cast<float4*>(&sK[key][4 * f4]) = kv;
      *reinterpret_cast<float4*>(&sV[key][4 * f4]) = vv;
    }
    __syncthreads();
    float s[16];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float4* kr = reinterpret_cast<const float4*>(&sK[c + 4 * i][0]);
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
      for (int d = 0; d < AF_D / 4; ++d) {
        const float4 k4 = kr[d];
        a0 = fmaf(q[4 * d], k4.x, a0); a1 = fmaf(q[4 * d + 1], k4.y, a1);
        a2 = fmaf(q[4 * d + 2], k4.z, a2); a3 = fmaf(q[4 * d + 3], k4.w, a3);
      }
      s[i] = (j0 + c + 4 * i < it.z) ? (a0 + a1) + (a2 + a3) : -INFINITY;
      mx = fmaxf(mx, s[i]);
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    const float m_new = fmaxf(m, mx);                                 // finite: every tile holds at least one valid key
    const float alpha = exp2f(m - m_new);
    float ps = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      s[i] = exp2f(s[i] - m_new);
      ps += s[i];
    }
    ps += __shfl_xor_sync(0xffffffffu, ps, 1);
    ps += __shfl_xor_sync(0xffffffffu, ps, 2);
    l = l * alpha + ps;
    m = m_new;
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i] *= alpha;
#pragma unroll
    for (int jj = 0; jj < AF_KEYS; ++jj) {
      const float pj = __shfl_sync(0xffffffffu, s[jj >> 2], (lane & ~3) | (jj & 3));   // key jj is owned by thread (jj & 3) of this row
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 v4 = *reinterpret_cast<const float4*>(&sV[jj][16 * i + 4 * c]);
        o[4 * i] = fmaf(pj, v4.x, o[4 * i]); o[4 * i + 1] = fmaf(pj, v4.y, o[4 * i + 1]);
        o[4 * i + 2] = fmaf(pj, v4.z, o[4 * i + 2]); o[4 * i + 3] = fmaf(pj, v4.w, o[4 * i + 3]);
      }
    }
  }
  if (!row_ok) return;
  const float inv = 1.f / l;
  __nv_bfloat16* op = out + static_cast<size_t>(qrow) * ldo + head * AF_D + 4 * c;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float x0 = o[4 * i] * inv, x1 = o[4 * i + 1] * inv, x2 = o[4 * i + 2] * inv, x3 = o[4 * i + 3] * inv;
    if (lo_off > 0) {
      uint32_t h0, l0, h1, l1;
      split_bf16x2(x0, x1, h0, l0);
      split_bf16x2(x2, x3, h1, l1);
      *reinterpret_cast<uint2*>(op + 16 * i) = make_uint2(h0, h1);
      *reinterpret_cast<uint2*>(op + lo_off + 16 * i) = make_uint2(l0, l1);
    } else {
      *reinterpret_cast<uint2*>(op + 16 * i) = make_uint2(pack_bf16x2(x0, x1), pack_bf16x2(x2, x3));
    }
  }
}

}  // namespace f5

extern "C" int f5_attention_f32(const float* qkv, int64_t ld, int32_t q_col, int32_t k_col, int32_t v_col, int32_t heads,
                                const int32_t* items, int32_t num_items, const float* rope, void* out, int64_t ldo, int32_t lo_off,
                                float softmax_scale, void* stream) {
  using namespace f5;
  if (qkv == nullptr || items == nullptr || out == nullptr || num_items <= 0 || heads <= 0) return F5_ERR_ARG;
  if ((ld % 4) != 0 || (q_col % 4) != 0 || (k_col % 4) != 0 || (v_col % 4) != 0 || (ldo % 4) != 0 || lo_off < 0 || (lo_off % 4) != 0)
    return F5_ERR_ARG;
  f5_launch(attn_f32_kernel, dim3(static_cast<unsigned>(num_items) * heads * 4), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), qkv,
            ld, q_col, k_col, v_col, heads, items, rope, reinterpret_cast<__nv_bfloat16*>(out), ldo, lo_off,
            softmax_scale * 1.4426950408889634f);
  return static_cast<int>(cudaGetLastError());
}
