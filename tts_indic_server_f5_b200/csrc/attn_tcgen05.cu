// Non-causal variable-length flash attention, head_dim 64, on tcgen05 (sm_100a).
//
// Persistent: one CTA per SM walks a list of work items = (pair of 128-row query tiles of one utterance, head).  The
// (item, key tile) pairs form ONE flat sequence of steps s = 0, 1, 2, ... that every role walks in the same order:
//   warp 8     : TMA producer  (Q pair once per item; K_s and V_s tiles of 128 keys through two SEPARATE 3-stage rings:
//                              K_{s+1} is consumed a whole softmax earlier than V_s)
//   warp 9     : MMA issuer    S_g = Q_g K_s^T (128x128x64, SS) -> TMEM;  O_g += P_g V_s (128x64x128, TS: P_g is the A operand
//                              IN TMEM, V the MN-major B operand read straight from the QKV buffer's tile — no transpose)
//   warps 0-3  : softmax group A, warps 4-7: softmax group B (g = query tile A / B) — one thread per query row (= TMEM
//                lane): S row read ONCE into registers, row max, exp2 with packed FFMA2/FADD2, P_g as packed bf16 back into
//                TMEM (tcgen05.st), lazy rescale of O_g in TMEM, final O/l store.
// Why P lives in TMEM: per step the SS form moved 256 KB through shared memory (QK operands 64 KB, PV operands 96 KB,
// P stores 64 KB, TMA fills 32 KB) = 2048 cycles at 128 B/clk — more than the 1024 tensor cycles and as much as the MUFU's
// 2048 (16 k exps per tile at 16/clk).  With P in TMEM the shared-memory traffic halves and the MUFU is the only unit near
// its limit.
// Software pipeline (what keeps the MUFU busy):
//   * a softmax thread releases S_g (`s_free`) as soon as its row sits in registers, so QK_g(s+1) is issued at the START
//     of softmax_g(s) and S_g(s+1) is waiting in TMEM when softmax_g(s) ends — no QK round trip between tiles;
//   * the step sequence is flat across work items: the first QK of the next item is issued during the last softmax of the
//     current one (Q is refilled as soon as the item's last QK retired);
//   * the first two 32-key chunks of exponentials are computed BEFORE the wait for the group's previous PV (their packed P
//     stays in registers): only the tcgen05.st of P, the O rescale and the O read-out need that PV to have retired;
//   * the O/l read-out of item i is deferred into the first softmax of item i+1 (after its row max, before its first P
//     store), so the last PV's latency is hidden behind the S load and the max pass;
//   * the two groups are kept HALF A TILE APART: QK_B(s) is only issued once group A has finished the row max of its
//     tile s (`a_gate`), so B's TMEM load + max pass (no MUFU work) overlaps A's exponentials and vice versa.  Left alone
//     the groups drift into lockstep (measured: both idle the MUFU for ~1100 of 3400 cycles per tile).
//     MMA issue order per step a:  QK_A(a+1) | QK_B(a) | PV_B(a-1) | PV_A(a).
// TMEM: S_A | S_B | P_A | P_B | O_A | O_B = 128+128+64+64+64+64 = 512 columns.
#define F5_DIAG_TAG 2u
#include "f5_common.cuh"
#include "../../include/f5_b200.h"
#include <cstdlib>

namespace f5 {

constexpr int ATT_THREADS = 384;   // warpgroup 0 / 1: softmax groups A / B; warpgroup 2: producer, MMA issuer, 2 idle warps
constexpr int ATT_BM = 128;   // query rows per tile (two tiles per work item)
constexpr int ATT_BN = 128;   // keys per tile
constexpr int ATT_D = 64;
constexpr int ATT_TILE_BYTES = 128 * 64 * 2;  // 16 KB
#ifndef ATT_KV_STAGES_N
#define ATT_KV_STAGES_N 3
#endif
constexpr int ATT_KV_STAGES = ATT_KV_STAGES_N;   // stages of the K ring and of the V ring (<= 4: barrier slots)
constexpr int ATT_SMEM = ATT_TILE_BYTES * (2 /*Q*/ + 2 * ATT_KV_STAGES /*K,V*/) + 256;
constexpr int ATT_TMEM_COLS = 512;
constexpr uint32_t ATT_TM_S = 0, ATT_TM_P = 256, ATT_TM_O = 384;   // column offsets; per group: + g*128 / g*64 / g*64
#ifndef ATT_POLY_COUNT
#define ATT_POLY_COUNT 1      // of every ATT_POLY_PERIOD score pairs, this many take the polynomial exp2 path (FMA pipes) instead of the MUFU
#endif
#ifndef ATT_GATE_POS
#define ATT_GATE_POS 0        // where group A opens the gate for QK_B of the same step: 0 after its row max, 1 / 2 after its first / second exp chunk
#endif
#ifndef ATT_DEFER_CHUNKS
#define ATT_DEFER_CHUNKS 2   // exp chunks computed before waiting for the previous PV (their P stays in registers)
#endif
#ifndef ATT_POLY_PERIOD
#define ATT_POLY_PERIOD 4
#endif

struct AttnParams {
  int q_col, k_col, v_col, heads;
  const int* items;   // [num_pairs][4] = q_row0, kv_row0, kv_len, q_rows_valid (1..256; 0 = padding item, skipped)
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

#ifndef ATT_TRACE
#define ATT_TRACE 0           // 1: clock64 event trace of CTA 0 (tools/attn_trace.py builds this variant); costs instructions, off in the product
#endif
#if ATT_TRACE
#define F5_TRACE(role, idx) do { if (p.trace != nullptr && blockIdx.x == 0 && (idx) < 512) p.trace[(role) * 512 + (idx)] = clock64(); } while (0)
#else
#define F5_TRACE(role, idx) do { } while (0)
#endif

// Walks the CTA's flat step sequence: work items w = blockIdx.x, +gridDim.x, ...; key tiles j = 0..nkv-1 inside each.
struct StepWalker {
  const int* items;
  int num_work, heads;
  int w, j, nkv, ngroups, head;
  int4 it;
  __device__ __forceinline__ explicit StepWalker(const AttnParams& p)
      : items(p.items), num_work(p.num_work), heads(p.heads), w(static_cast<int>(blockIdx.x)), j(0) { load(); }
  __device__ __forceinline__ bool valid() const { return w < num_work; }
  __device__ __forceinline__ void load() {
    while (w < num_work) {                              // items with no query rows are padding of the tile table: skip them
      it = *reinterpret_cast<const int4*>(items + 4 * (w / heads));
      if (it.w > 0 && it.z > 0) break;
      w += static_cast<int>(gridDim.x);
    }
    head = w % heads;
    nkv = (it.z + ATT_BN - 1) / ATT_BN;
    ngroups = it.w > ATT_BM ? 2 : 1;
  }
  __device__ __forceinline__ bool last_tile() const { return j + 1 == nkv; }
  __device__ __forceinline__ void next() {
    if (++j == nkv) { j = 0; w += static_cast<int>(gridDim.x); load(); }
  }
};

// Registers: the file is 16 K per SM sub-partition and every sub-partition hosts one warp of each warpgroup, so a flat
// allocation caps at 168/thread — not enough for the 128-register S row plus the exp pipeline.  setmaxnreg moves
// registers from the producer/MMA warpgroup (72) to the two softmax warpgroups (216): 2 x 216 + 72 = 504 <= 512 = 16 K / 32 (an exact 512 never gets its registers: the inc hangs).
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_d64_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();   // 128B-swizzled tiles need 1024-B aligned bases
  uint8_t* sQ = smem;                                         // [2] tiles A, B
  uint8_t* sK = sQ + 2 * ATT_TILE_BYTES;                      // [ATT_KV_STAGES]
  uint8_t* sV = sK + ATT_KV_STAGES * ATT_TILE_BYTES;          // [ATT_KV_STAGES]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ATT_KV_STAGES * ATT_TILE_BYTES);
  uint64_t* q_full = bars;          // [2 groups]
  uint64_t* q_empty = bars + 2;     // [2 groups]
  uint64_t* k_full = bars + 4;      // [ATT_KV_STAGES <= 4]
  uint64_t* k_empty = bars + 8;
  uint64_t* v_full = bars + 12;
  uint64_t* v_empty = bars + 16;
  uint64_t* s_full = bars + 20;     // [2 groups]
  uint64_t* s_free = bars + 22;
  uint64_t* p_full = bars + 24;
  uint64_t* o_full = bars + 26;
  uint64_t* a_gate = bars + 28;     // group A finished the row max of its tile
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 29);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    mbar_init(a_gate, 128);
    for (int i = 0; i < ATT_KV_STAGES; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&q_full[g], 1);
      mbar_init(&q_empty[g], 1);
      mbar_init(&s_full[g], 1);
      mbar_init(&s_free[g], 128);
      mbar_init(&p_full[g], 128);
      mbar_init(&o_full[g], 1);
    }
    mbar_fence_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_ptr, ATT_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();                                // set-up above overlapped the QKV GEMM's tail; its output is read below
  pdl_launch();

  if (warp >= 8) {
   asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
   if (warp == 8) {
    // ------------------------------------------------------------------ TMA producer
    // Issue order per step s:  Q_A(item) | K(s) | Q_B(item) | V(s-1)  — the order the MMA warp consumes them in.
    // The whole warp walks the loop; only the TMA instructions are predicated on one elected lane (uniform registers).
    {
      const bool leader = elect_one();
      uint32_t item_a = 0, item_b = 0, s = 0;
      int v_col = 0, v_row = 0;
      auto load_v = [&](uint32_t sv_step) {
        const uint32_t sv = sv_step % ATT_KV_STAGES, pv = (sv_step / ATT_KV_STAGES) & 1;
        mbar_wait_warp(&v_empty[sv], pv ^ 1, leader);
        if (leader) {
          mbar_expect_tx(&v_full[sv], ATT_TILE_BYTES);
          tma_load_2d(sV + sv * ATT_TILE_BYTES, &tmap_qkv, &v_full[sv], v_col, v_row);
        }
        __syncwarp();
      };
      for (StepWalker c(p); c.valid(); c.next(), ++s) {
        if (c.j == 0) {
          mbar_wait_warp(&q_empty[0], (item_a & 1) ^ 1, leader);            // previous item's last QK_A has retired
          if (leader) {
            mbar_expect_tx(&q_full[0], ATT_TILE_BYTES);
            tma_load_2d(sQ, &tmap_qkv, &q_full[0], p.q_col + c.head * ATT_D, c.it.x);
          }
          __syncwarp();
          ++item_a;
        }
        const uint32_t st = s % ATT_KV_STAGES, ph = (s / ATT_KV_STAGES) & 1;
        mbar_wait_warp(&k_empty[st], ph ^ 1, leader);
        if (leader) {
          mbar_expect_tx(&k_full[st], ATT_TILE_BYTES);
          tma_load_2d(sK + st * ATT_TILE_BYTES, &tmap_qkv, &k_full[st], p.k_col + c.head * ATT_D, c.it.y + c.j * ATT_BN);
        }
        __syncwarp();
        if (lane == 0) { F5_TRACE(0, s); }
        if (c.j == 0 && c.ngroups == 2) {
          mbar_wait_warp(&q_empty[1], (item_b & 1) ^ 1, leader);            // group B's previous item's last QK_B has retired
          if (leader) {
            mbar_expect_tx(&q_full[1], ATT_TILE_BYTES);
            tma_load_2d(sQ + ATT_TILE_BYTES, &tmap_qkv, &q_full[1], p.q_col + c.head * ATT_D, c.it.x + ATT_BM);
          }
          __syncwarp();
          ++item_b;
        }
        if (s > 0) load_v(s - 1);
        v_col = p.v_col + c.head * ATT_D;
        v_row = c.it.y + c.j * ATT_BN;
      }
      if (s > 0) load_v(s - 1);
    }
   } else if (warp == 9) {
    // ------------------------------------------------------------------ MMA issuer
    // The WHOLE warp walks the loop (warp-uniform control flow keeps descriptors / addresses in uniform registers, which is
    // what UTCHMMA consumes); only the tcgen05.mma / commit instructions themselves are predicated on one elected lane.
    constexpr uint32_t idesc_qk = umma_idesc_bf16(ATT_BM, ATT_BN);
    constexpr uint32_t idesc_pv = umma_idesc_bf16(ATT_BM, ATT_D) | (1u << 16);   // A (= P, TMEM) K-major; B (= V) MN-major
    const bool leader = elect_one();
    const uint64_t qd0 = umma_desc_k_sw128(smem_u32(sQ));
    const uint64_t kd0 = umma_desc_k_sw128(smem_u32(sK));
    const uint64_t vd0 = umma_desc_mn_sw128(smem_u32(sV));
    constexpr uint64_t TILE16 = ATT_TILE_BYTES >> 4;      // descriptor address field is in 16-B units
    uint32_t a = 0, ev = 0;
    uint32_t iq0 = 0, iq1 = 0;     // items started per group (phase of q_full)
    uint32_t tq0 = 0, tq1 = 0;     // QKs issued so far per group (phase of s_free)
    uint32_t tp0 = 0, tp1 = 0;     // PVs issued so far per group (phase of p_full)
    // S_g(step) = Q_g K^T: waits for the group's previous S to be in registers, then 4 MMAs + commit
    auto issue_qk = [&](int g, const StepWalker& c, uint32_t step, uint32_t& iq, uint32_t& tq) {
      const uint32_t st = step % ATT_KV_STAGES;
      if (c.j == 0) { mbar_wait_warp(&q_full[g], iq & 1, leader); ++iq; }
      mbar_wait_warp(&k_full[st], (step / ATT_KV_STAGES) & 1, leader);
      if (tq > 0) mbar_wait_warp(&s_free[g], (tq - 1) & 1, leader);
      tc_fence_after();
      const uint64_t qd = qd0 + g * TILE16, kd = kd0 + st * TILE16;
      if (leader) {
#pragma unroll
        for (int kk = 0; kk < ATT_D / 16; ++kk)
          umma_f16_ss(tmem_base + ATT_TM_S + g * ATT_BN, qd + 2 * kk, kd + 2 * kk, idesc_qk, kk != 0);
        umma_commit(&s_full[g]);
        if (c.last_tile()) umma_commit(&q_empty[g]);   // every QK_g of the item is issued: Q_g may be refilled
      }
      __syncwarp();
      ++tq;
    };
    // O_g (+)= P_g V: P_g from TMEM (8 columns = 16 bf16 keys per MMA), V tile rows kk*16.. (2 KB per MMA)
    auto issue_pv = [&](int g, const StepWalker& c, uint32_t step, uint32_t& tp) {
      const uint32_t st = step % ATT_KV_STAGES;
      mbar_wait_warp(&v_full[st], (step / ATT_KV_STAGES) & 1, leader);
      mbar_wait_warp(&p_full[g], tp & 1, leader);                // group g wrote P_g (and read out the previous item's O_g if j == 0)
      tc_fence_after();
      const uint64_t vd = vd0 + st * TILE16;
      const bool first = c.j == 0;
      if (leader) {
#pragma unroll
        for (int kk = 0; kk < ATT_BN / 16; ++kk)
          umma_f16_ts(tmem_base + ATT_TM_O + g * ATT_D, tmem_base + ATT_TM_P + g * (ATT_BN / 2) + kk * 8, vd + kk * 128,
                      idesc_pv, (!first || kk != 0) ? 1u : 0u);
        umma_commit(&o_full[g]);
      }
      __syncwarp();
      ++tp;
    };
    StepWalker cur(p), prv(p), nxt(p);
    nxt.next();
    if (cur.valid()) issue_qk(0, cur, 0, iq0, tq0);
    while (cur.valid()) {
      if (lane == 0) { F5_TRACE(1, ev); }
      ++ev;
#if F5_WAIT_ALL_LANES
      if (nxt.valid()) issue_qk(0, nxt, a + 1, iq0, tq0);          // QK_A(a+1): at the start of softmax_A(a)
      mbar_wait(a_gate, a & 1);                                    // softmax_A(a) is past its row max
      if (cur.ngroups == 2) issue_qk(1, cur, a, iq1, tq1);         // QK_B(a): group B runs half a tile behind A
      if (leader) umma_commit(&k_empty[a % ATT_KV_STAGES]);        // QK_A(a) (issued a step ago) and QK_B(a) retired -> K stage reusable
      __syncwarp();
#else
      // Round 1 issued QK_A(a+1) BEFORE waiting for gate(a).  That made a_gate the one barrier whose NEXT phase did not depend
      // on its waiter: group A reaches gate(a+1) through S_A(a+1), which was already on its way.  When this warp's try_wait
      // on gate(a) slept through the phase flip (a try_wait may return late — it suspends the thread for a system-dependent
      // time), group A finished tile a, took S_A(a+1), passed its row max and completed gate(a+1): the parity this warp
      // waits for then reads "not complete" for ever.  The soak run's watchdog record shows exactly that (warp 9 stuck on
      // a_gate with every other barrier of the CTA idle; profiles/r02_soak_ab.md).  Now the gate comes first; QK_B(a), which
      // group B needs soonest, goes out right behind it and QK_A(a+1) follows — group A needs S_A(a+1) only ~1500 cycles later.
      mbar_wait_warp(a_gate, a & 1, leader);                       // softmax_A(a) is past its row max
      if (cur.ngroups == 2) issue_qk(1, cur, a, iq1, tq1);         // QK_B(a): group B runs half a tile behind A
      if (leader) umma_commit(&k_empty[a % ATT_KV_STAGES]);        // QK_A(a) (issued a step ago) and QK_B(a) retired -> K stage reusable
      __syncwarp();
      if (nxt.valid()) issue_qk(0, nxt, a + 1, iq0, tq0);          // QK_A(a+1): during softmax_A(a)
#endif
      if (lane == 0) { F5_TRACE(1, ev); }
      ++ev;
      if (a > 0) {
        if (prv.ngroups == 2) issue_pv(1, prv, a - 1, tp1);        // PV_B(a-1)
        if (leader) umma_commit(&v_empty[(a - 1) % ATT_KV_STAGES]);
        __syncwarp();
      }
      issue_pv(0, cur, a, tp0);                                    // PV_A(a)
      if (lane == 0) { F5_TRACE(1, ev); }
      ++ev;
      prv = cur;
      cur = nxt;
      nxt.next();
      ++a;
    }
    if (a > 0) {
      if (prv.ngroups == 2) issue_pv(1, prv, a - 1, tp1);
      if (leader) umma_commit(&v_empty[(a - 1) % ATT_KV_STAGES]);
      __syncwarp();
    }
   }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    // ------------------------------------------------------------------ softmax groups
    const int g = warp >> 2;                      // 0: tile A, 1: tile B
    const int row = static_cast<int>(pin_u32(threadIdx.x & 127u));      // = TMEM lane; warp w may access lanes 32*(w%4)..
    const uint32_t lane_off = static_cast<uint32_t>(row & 96) << 16;
    const uint32_t tmem_S = tmem_base + ATT_TM_S + g * ATT_BN + lane_off;
    const uint32_t tmem_P = tmem_base + ATT_TM_P + g * (ATT_BN / 2) + lane_off;
    const uint32_t tmem_O = tmem_base + ATT_TM_O + g * ATT_D + lane_off;
    // barrier addresses once, as 32-bit shared-window addresses the compiler cannot rematerialise from special registers
    const uint32_t b_s_full = pin_u32(smem_u32(&s_full[g])), b_s_free = pin_u32(smem_u32(&s_free[g]));
    const uint32_t b_p_full = pin_u32(smem_u32(&p_full[g])), b_o_full = pin_u32(smem_u32(&o_full[g]));
    const uint32_t b_gate = pin_u32(smem_u32(a_gate));
    const float2 sc2 = make_float2(p.scale_log2, p.scale_log2);
    uint32_t t = 0;
    // deferred read-out of the previous item's O (runs inside the first softmax of the next item)
    bool pend = false;
    float pend_inv_l = 0.f;
    __nv_bfloat16* pend_out = nullptr;            // nullptr: row beyond q_valid, nothing to store
    auto read_out = [&]() {
#pragma unroll 1
      for (int c = 0; c < ATT_D / 32; ++c) {
        uint32_t ro[32];
        tmem_ld_32x32b_x32(tmem_O + c * 32, ro);
        tmem_ld_wait();
        if (pend_out != nullptr) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 wv;
            wv.x = pack_bf16x2(__uint_as_float(ro[i]) * pend_inv_l, __uint_as_float(ro[i + 1]) * pend_inv_l);
            wv.y = pack_bf16x2(__uint_as_float(ro[i + 2]) * pend_inv_l, __uint_as_float(ro[i + 3]) * pend_inv_l);
            wv.z = pack_bf16x2(__uint_as_float(ro[i + 4]) * pend_inv_l, __uint_as_float(ro[i + 5]) * pend_inv_l);
            wv.w = pack_bf16x2(__uint_as_float(ro[i + 6]) * pend_inv_l, __uint_as_float(ro[i + 7]) * pend_inv_l);
            *reinterpret_cast<uint4*>(pend_out + c * 32 + i) = wv;
          }
        }
      }
      pend = false;
    };
    for (int w = blockIdx.x; w < p.num_work; w += gridDim.x) {
      const int4 it = *reinterpret_cast<const int4*>(p.items + 4 * (w / p.heads));
      const int head = w % p.heads;
      const int q_valid = it.w - g * ATT_BM;      // rows of this group's tile that exist
      if (q_valid <= 0 || it.z <= 0) continue;    // tile B absent / padding item: the MMA warp skips it too
      const int kv_len = it.z;
      const int nkv = (kv_len + ATT_BN - 1) / ATT_BN;
      float m_run = -INFINITY, l_run = 0.f;
      for (int j = 0; j < nkv; ++j, ++t) {
        const int kv_valid = min(ATT_BN, kv_len - j * ATT_BN);
        const bool full = kv_valid == ATT_BN;     // only the last tile of an utterance needs key masking
        mbar_wait_a(b_s_full, t & 1);
        tc_fence_after();
        if (row == 0) F5_TRACE(2 + g, 8 * t);
        // S row (128 fp32) is read from TMEM ONCE into registers, then S_g is handed back to the MMA warp at once.
        uint32_t r[128];
        tmem_ld_32x32b_x32(tmem_S, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
        tmem_ld_32x32b_x32(tmem_S + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
        tmem_ld_32x32b_x32(tmem_S + 64, *reinterpret_cast<uint32_t(*)[32]>(&r[64]));
        tmem_ld_32x32b_x32(tmem_S + 96, *reinterpret_cast<uint32_t(*)[32]>(&r[96]));
        tmem_ld_wait();
        if (row == 0) F5_TRACE(2 + g, 8 * t + 1);
        tc_fence_before();
        mbar_arrive_a(b_s_free);
        if (!full) {
#pragma unroll
          for (int i = 0; i < 128; ++i)
            if (i >= kv_valid) r[i] = 0xff800000u;   // -inf: masked keys drop out of the max and give exp2 = 0
        }
        // row max as 8 independent chains (a single 64-deep FMNMX chain cost ~550 cycles of pure latency per tile)
        float mxs[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) mxs[k] = fmaxf(__uint_as_float(r[2 * k]), __uint_as_float(r[2 * k + 1]));
#pragma unroll
        for (int i = 8; i < 64; ++i) mxs[i & 7] = fmaxf(mxs[i & 7], fmaxf(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])));
        const float mx = fmaxf(fmaxf(fmaxf(mxs[0], mxs[1]), fmaxf(mxs[2], mxs[3])), fmaxf(fmaxf(mxs[4], mxs[5]), fmaxf(mxs[6], mxs[7])));
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
        if (ATT_GATE_POS == 0 && g == 0) mbar_arrive_a(b_gate);   // group B's scores for this step may be issued now (half-tile stagger)
        if (row == 0) F5_TRACE(2 + g, 8 * t + 2);
        // p = exp2(s*scale - m), row sum, P as packed bf16 pairs -> TMEM columns [c*16, c*16+16) of P_g.
        const float2 nm2 = make_float2(-m_run, -m_run);
        float2 ls2 = make_float2(0.f, 0.f), ls2b = make_float2(0.f, 0.f);
        auto exp_chunk = [&](int c, uint32_t (&pk)[16]) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float2 e = ffma2(make_float2(__uint_as_float(r[c * 32 + 2 * i]), __uint_as_float(r[c * 32 + 2 * i + 1])), sc2, nm2);
            if ((i % ATT_POLY_PERIOD) < ATT_POLY_COUNT) {   // this pair on the FMA/ALU pipes, the others on the MUFU
              e = exp2_poly2(e);
            } else {
              e.x = fast_ex2(e.x);
              e.y = fast_ex2(e.y);
            }
            if (i & 1) ls2b = fadd2(ls2b, e); else ls2 = fadd2(ls2, e);
            pk[i] = pack_bf16x2(e.x, e.y);
          }
        };
        // The first two chunks are computed BEFORE waiting for the group's previous PV: that PV (issued at the end of the
        // previous softmax, ~600 cycles of issue + tensor + commit latency) still reads P_g and accumulates into O_g, so
        // only the tcgen05.st of P, the O rescale and the deferred read-out have to wait for it — the exponentials do not
        // (ncu: 60 % of the tiles stalled ~300 cycles here when the wait came first).  Waiting on EVERY tile also keeps
        // this thread exactly one phase behind o_full (a parity wait must never fall two phases behind).
        uint32_t pk0[16], pk1[16];
#if ATT_DEFER_CHUNKS >= 3
        uint32_t pk2[16];
#endif
        exp_chunk(0, pk0);
        if (ATT_GATE_POS == 1 && g == 0) mbar_arrive_a(b_gate);
        exp_chunk(1, pk1);
        if (ATT_GATE_POS == 2 && g == 0) mbar_arrive_a(b_gate);
#if ATT_DEFER_CHUNKS >= 3
        exp_chunk(2, pk2);
#endif
        if (t > 0) {
          mbar_wait_a(b_o_full, (t - 1) & 1);
          tc_fence_after();
        }
        if (pend) read_out();                     // previous item's O / l -> bf16 output rows, before PV(this item, 0) overwrites O_g
        if (j > 0 && any_grow) {
#pragma unroll 1
          for (int c = 0; c < ATT_D / 32; ++c) {
            uint32_t ro[32];
            tmem_ld_32x32b_x32(tmem_O + c * 32, ro);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) ro[i] = __float_as_uint(__uint_as_float(ro[i]) * alpha);
            tmem_st_32x32b_x32(tmem_O + c * 32, ro);
          }
        }
        if (row == 0) F5_TRACE(2 + g, 8 * t + 3);
        tmem_st_32x32b_x16(tmem_P, pk0);
        tmem_st_32x32b_x16(tmem_P + 16, pk1);
#if ATT_DEFER_CHUNKS >= 3
        tmem_st_32x32b_x16(tmem_P + 32, pk2);
#endif
#pragma unroll
        for (int c = ATT_DEFER_CHUNKS; c < ATT_BN / 32; ++c) {
          uint32_t pk[16];
          exp_chunk(c, pk);
          tmem_st_32x32b_x16(tmem_P + c * 16, pk);
        }
        l_run = l_run * alpha + ((ls2.x + ls2b.x) + (ls2.y + ls2b.y));
        if (row == 0) F5_TRACE(2 + g, 8 * t + 4);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive_a(b_p_full);
        if (row == 0) F5_TRACE(2 + g, 8 * t + 5);
      }
      pend = true;
      pend_inv_l = 1.f / l_run;
      pend_out = row < q_valid ? p.out + static_cast<size_t>(it.x + g * ATT_BM + row) * p.ldo + head * ATT_D : nullptr;
    }
    if (pend) {
      mbar_wait_a(b_o_full, (t - 1) & 1);
      tc_fence_after();
      read_out();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, ATT_TMEM_COLS);
  }
}

int make_tmap_bf16_2d(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_rows);
static long long* g_trace_buf = nullptr;

}  // namespace f5

F5_DEFINE_DIAG_SETTER(f5_diag_set_attn)

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
  static const cudaError_t attr_rc =       // C++11 magic static: set once, thread-safe
      cudaFuncSetAttribute(attn_d64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM);
  if (attr_rc != cudaSuccess) return static_cast<int>(attr_rc);
  AttnParams p;
  p.q_col = q_col; p.k_col = k_col; p.v_col = v_col; p.heads = heads;
  p.items = items;
  p.num_work = num_items * heads;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.scale_log2 = softmax_scale * 1.4426950408889634f;
#if ATT_TRACE
  if (g_trace_buf == nullptr) {          // trace builds only (tools/attn_trace.py): the product launcher never allocates
    cudaMalloc(&g_trace_buf, 4 * 512 * sizeof(long long));
    cudaMemset(g_trace_buf, 0, 4 * 512 * sizeof(long long));
  }
#endif
  p.trace = g_trace_buf;
  const int grid = p.num_work < kNumSMsB200 ? p.num_work : kNumSMsB200;
  return static_cast<int>(f5_launch(attn_d64_kernel, dim3(grid), dim3(ATT_THREADS), ATT_SMEM, reinterpret_cast<cudaStream_t>(stream), tq, p));
}
