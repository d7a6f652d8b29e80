/* f5_b200.h — C ABI of the B200-native F5-TTS (IndicF5) inference hot path.
 *
 * One shared library (`tts_indic_server_f5_b200/libf5b200.so`, built for sm_100a only) exports every
 * kernel launcher the path needs.  Signatures carry plain pointers, sizes and a `cudaStream_t` passed as
 * `void*`; device pointers are raw addresses (no torch types).  Every function returns 0 on success or a
 * negative F5_ERR_* / positive cudaError_t code; nothing allocates, nothing synchronises, nothing falls
 * back to the CPU.  Workspaces are owned by the caller (the Python engine allocates them as torch tensors).
 *
 * The reference is pure Python with no FFI: each launcher replaces a torch library-call site of the
 * reference (paths relative to /root/reference/src/server/f5_tts/), cited per function below.  The binding a
 * maintainer adds on the reference side is a ctypes stub — see INTEGRATION.md.
 */
#ifndef F5_B200_H
#define F5_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define F5_OK 0
#define F5_ERR_ARG (-1)       /* bad shape / alignment / null pointer */
#define F5_ERR_DRIVER (-2)    /* cuTensorMapEncodeTiled unavailable or failed */
#define F5_ERR_ARCH (-3)      /* device is not sm_100 */

/* ---- GEMM epilogue modes ------------------------------------------------------------------- */
#define F5_EPI_STORE_BF16 0   /* out_bf16[m,n] = act(acc + bias[n])        (+ optional head-0 RoPE, row mask) */
#define F5_EPI_STORE_F32 1    /* out_f32[m,n]  = act(acc + bias[n]) + addend[m,n]; optional masked bf16 copy out2 */
#define F5_EPI_RESID_F32 2    /* resid[m,n]   += gate[n] * act(acc + bias[n])   (gate NULL => 1) */

#define F5_ACT_NONE 0
#define F5_ACT_GELU_TANH 1    /* DiT FFN, model/modules.py:556 */
#define F5_ACT_GELU_ERF 2     /* ConvNeXtV2 / Vocos, model/modules.py:255 */
#define F5_ACT_MISH 3         /* ConvPositionEmbedding, model/modules.py:173,175 */

/* Dense / implicit-conv GEMM on tcgen05 tensor cores: D[M,N] = A[M,K] * B[N,K]^T, bf16 operands, fp32
 * accumulation in TMEM, fused epilogue.  Replaces every `nn.Linear` / `nn.Conv1d` call site on the path:
 * to_q/to_k/to_v/to_out (model/modules.py:409-411,441), FFN (:324-325), AdaLN linears (:286,:307),
 * input proj (model/backbones/dit.py:85), grouped k=31 conv (model/modules.py:171-176), ConvNeXtV2 pointwise
 * convs (:265,:268), time MLP (:652), proj_out (dit.py:161), and Vocos' embed conv / pointwise / head linears.
 *
 * Implicit conv (num_taps > 1): K loop runs over (tap, kc): A rows are shifted by (tap - tap_pad) (TMA
 * zero-fills rows outside [0, a_rows)), B rows are offset by tap * b_tap_rows.  a_grouped != 0 selects the
 * input-channel block of the output-channel group (grouped conv with group width == block_n).            */
typedef struct f5_gemm_args {
  const void* A;        /* bf16 [a_rows, lda], K-major                                                */
  const void* B;        /* bf16 [b_rows, ldb], K-major ([N,K] like nn.Linear.weight)                  */
  int64_t lda, ldb;     /* row strides in elements (multiples of 8)                                   */
  int32_t a_rows, a_cols; /* extents of the A tensor (rows; valid columns)                            */
  int32_t b_rows, b_cols;
  int32_t M, N;         /* output extents: rows m < M and cols n < N are stored (N % 8 == 0)          */
  int32_t block_n;      /* 64, 128 or 256                                                             */
  int32_t num_taps, kc_per_tap, tap_pad, a_grouped, b_tap_rows;   /* dense GEMM: 1, ceil(K/64), 0, 0, 0 */
  int32_t mode, act;
  const float* bias;    /* [N] or NULL                                                                */
  const float* gate;    /* [N] or NULL (RESID mode)                                                   */
  void* out;  int64_t ldo;      /* STORE modes                                                        */
  void* out2; int64_t ldo2;     /* STORE_F32: optional bf16 copy                                      */
  const float* addend; int64_t ld_add;  /* STORE_F32: optional fp32 [M, ld_add]                       */
  float* resid; int64_t ldr;    /* RESID mode: fp32 [M, ldr], read-modify-write                       */
  const int32_t* row_pos;       /* [M] position of the row inside its utterance, -1 for gap rows, or NULL */
  int32_t mask_rows;            /* != 0: rows with row_pos < 0 are stored as zeros (bf16 outputs)      */
  const float* rope;            /* [max_pos, 32] (cos, sin) pairs or NULL; STORE_BF16 only             */
  int32_t rope_period;          /* RoPE applies to columns [t*period, t*period+64), t < rope_tiles     */
  int32_t rope_tiles;
  int32_t num_sms;              /* persistent grid size (0 => 148)                                     */
  /* Split-operand ("bf16x3") mode — fp32-class products on the bf16 tensor cores for the fp32 precision mode (the reference as
   * deployed is fp32 end to end, core/managers.py:76): every operand is carried as two bf16 planes, v = hi + lo with
   * hi = bf16(v), lo = bf16(v - hi), and D = A_hi B_hi^T + A_hi B_lo^T + A_lo B_hi^T (the lo x lo term is below fp32 round-off of
   * the sum).  taps_per_seg > 0 switches it on: num_taps must be 3 * taps_per_seg; tap = seg * taps_per_seg + t shifts the A rows
   * by (t - tap_pad) as usual; B holds the planes stacked by rows in the order [hi taps | lo taps | hi taps] (b_tap_rows rows per
   * tap); A holds its low plane a_lo_off columns (a multiple of 64) to the right of the high plane and segment 2 reads it.
   * taps_per_seg == 0: ordinary launch.                                                                       */
  int32_t taps_per_seg;
  int32_t a_lo_off;
} f5_gemm_args;

int f5_gemm_bf16(const f5_gemm_args* args, void* stream);

/* Non-causal variable-length attention, head_dim 64, on tcgen05 (S and O accumulate in TMEM, online softmax in
 * fp32).  Replaces F.scaled_dot_product_attention at model/modules.py:436 (+ head split/merge :424-437) with
 * per-utterance (batch-1) semantics: a query tile attends to the keys of its own utterance only.
 * qkv: bf16 [rows, ld] with q at column q_col + h*64, k at k_col + h*64, v at v_col + h*64.
 * tiles: int32 [num_tiles, 4] = {q_row0, kv_row0, kv_len, q_rows_valid <= 256} — one work item is a PAIR of 128-row query
 * tiles (x every head); the kernel is persistent (one CTA per SM).  out: bf16 [rows, ldo], head h at h*64. */
int f5_attention_d64(const void* qkv, int64_t ld, int32_t rows, int32_t q_col, int32_t k_col, int32_t v_col,
                     int32_t heads, const int32_t* tiles, int32_t num_tiles, void* out, int64_t ldo,
                     float softmax_scale, void* stream);

/* fp32 attention for the fp32 precision mode: same work items and per-utterance semantics as f5_attention_d64, but Q / K / V are
 * fp32 (the QKV GEMM's fp32 output), every product, the softmax and the accumulation are fp32 on the CUDA cores, and RoPE
 * (x-transformers rotary on head 0 only, model/modules.py:414-426) is applied while Q and K are loaded: rope fp32 [max_pos, 64] =
 * (cos, sin) per interleaved pair, NULL = none.  out: bf16 [rows, ldo] with head h at h*64; lo_off > 0 writes the split planes.
 * Replaces F.scaled_dot_product_attention at model/modules.py:436 in fp32. */
int f5_attention_f32(const float* qkv, int64_t ld, int32_t q_col, int32_t k_col, int32_t v_col, int32_t heads,
                     const int32_t* tiles, int32_t num_tiles, const float* rope, void* out, int64_t ldo, int32_t lo_off,
                     float softmax_scale, void* stream);

/* y[m,:] = LayerNorm(x_f32[m,:], eps) * (a_off + a[:]) + b[:] (bf16 and/or fp32 output, either may be NULL) — model/modules.py:289,:310,:568 (a_off = 1,
 * a = scale, b = shift) and the affine LayerNorms of ConvNeXtV2 / Vocos (a_off = 0, a = weight, b = bias).  D % 128 == 0, D <= 1024. */
int f5_layernorm_mod(const float* x, int64_t ldx, void* y_bf16, int64_t ldy, float* y_f32, int64_t ldy32, int32_t M,
                     int32_t D, const float* a, const float* b, float a_off, float eps, int32_t lo_off, void* stream);

/* Split-operand outputs (fp32 precision mode).  Kernels that produce a bf16 GEMM operand take `lo_off`: 0 = the ordinary bf16
 * output; > 0 = the value v is written as two planes, hi = bf16(v) at the usual column and lo = bf16(v - hi) lo_off columns to the
 * right (the A operand layout of f5_gemm_bf16's split-operand mode). */

/* Depthwise Conv1d(k=7, pad=3) over the rows of one utterance + affine LayerNorm -> bf16
 * (model/modules.py:262-264; Vocos ConvNeXtBlock).  row_pos marks utterance membership (halo rows with row_pos < 0 or
 * outside [0,M) contribute zero).  w: fp32 [C,7], bias [C]. */
int f5_dwconv7_ln(const float* x, int64_t ldx, void* y, int64_t ldy, int32_t M, int32_t C, const int32_t* row_pos,
                  const float* w, const float* bias, const float* ln_w, const float* ln_b, float eps, int32_t lo_off, void* stream);

/* GRN over each utterance's own rows (model/modules.py:231-234): phase 1 computes the sum of squares per (segment,
 * channel) in a fixed order (deterministic, no atomics); phase 2 applies gamma*(x*Nx)+beta+x in place on the bf16 activations. */
int f5_grn_sumsq(const void* x_bf16, int64_t ldx, int32_t C, const int32_t* seg_rows, int32_t num_segs, float* sumsq,
                 void* stream);
int f5_grn_apply(void* x_bf16, int64_t ldx, int32_t C, const int32_t* seg_rows, int32_t num_segs, const float* sumsq,
                 const float* gamma, const float* beta, void* stream);
/* The same two passes on fp32 activations (fp32 precision mode). */
int f5_grn_sumsq_f32(const float* x, int64_t ldx, int32_t C, const int32_t* seg_rows, int32_t num_segs, float* sumsq, void* stream);
int f5_grn_apply_f32(float* x, int64_t ldx, int32_t C, const int32_t* seg_rows, int32_t num_segs, const float* sumsq,
                     const float* gamma, const float* beta, void* stream);

/* Text token gather + absolute sinusoidal position (model/backbones/dit.py:56-64): out_f32[row,:] = emb[ids[row]] +
 * pos_table[min(row_pos[row], max_pos-1)] for rows with row_pos >= 0, zeros otherwise. */
int f5_text_gather_pos(const int32_t* ids, const int32_t* row_pos, const float* emb, const float* pos_table,
                       int32_t max_pos, float* out, int64_t ldo, int32_t M, int32_t C, void* stream);

/* fp32 -> bf16 row gather/pack into a GEMM A operand: dst[m, dst_col + c] = src[src_rows ? src_rows[m] : m, c] for
 * c < C, zero fill for C <= c < C_pad; rows whose source index (or row_pos[m], when given) is negative are zeros. */
int f5_pack_bf16(const float* src, int64_t lds, void* dst, int64_t ldd, int32_t dst_col, int32_t M, int32_t C,
                 int32_t C_pad, const int32_t* src_rows, const int32_t* row_pos, int32_t lo_off, void* stream);

/* x[m,:] = flag[m] ? c[m,:] : x[m,:]   — prompt re-insert `where(cond_mask, cond, out)` (model/cfm.py:204). */
int f5_where_rows(float* x, int64_t ldx, const float* c, int64_t ldc, const int32_t* flag, int32_t M, int32_t C,
                  void* stream);

/* Fused CFG blend + Euler step (model/cfm.py:176 + torchdiffeq Euler at :200) over the real tokens of the batch:
 * v = pc + (pc - pu) * cfg ; x += dt * v, where pc = pred[row], pu = pred[row + half_rows].  Also refreshes the bf16
 * A operand of the next step's input projection for both CFG halves (xb[row], xb[row + half_rows]).
 * dt is read from dts[step] on the device so the whole sampling loop can live in one CUDA graph. */
int f5_cfg_euler(float* x, int64_t ldx, const float* pred, int64_t ldp, int32_t half_rows, int32_t C,
                 const int32_t* row_pos, const float* dts, int32_t step, float cfg_strength, void* xb, int64_t ldxb,
                 int32_t C_pad, void* stream);

/* Initial noise y0_i = randn(n_i, C) drawn on the device (model/cfm.py:181-186: `torch.randn(dur, num_channels, device=...)`;
 * the server passes no seed, so every request is a fresh draw).  Philox4x32-10 keyed by utt_seed[row_utt[row]], counter
 * (row_pos[row] * 32 + lane, 0x4635, 0, 0), Box-Muller on 24-bit uniforms: the value at (seed, frame, channel) is independent of
 * the packing.  Rows with row_pos < 0 are zero-filled.  x: fp32 [M, ldx] (columns [C, ldx) untouched), C <= 128. */
int f5_randn_rows(float* x, int64_t ldx, int32_t M, int32_t C, const int32_t* row_pos, const int32_t* row_utt,
                  const uint64_t* utt_seed, void* stream);

/* Sinusoidal time embedding (model/modules.py:154-160): out_bf16[s, :] = [sin((1000 t_s) f_k) | cos(...)], k < dim/2;
 * freqs = exp(-k ln(1e4)/(dim/2-1)) is passed in (fp32 [dim/2]). */
int f5_time_sinus(const float* t, int32_t steps, const float* freqs, int32_t dim, void* out_bf16, int64_t ldo,
                  int32_t lo_off, void* stream);

/* out_bf16 = silu(x_f32) elementwise (AdaLN input, model/modules.py:286).  split_cols > 0: x is [n / split_cols, split_cols] and
 * out is [rows, 2 * split_cols] = high plane | low plane (fp32 precision mode). */
int f5_silu_bf16(const float* x, void* out_bf16, int64_t n, int32_t split_cols, void* stream);

/* Vocos ISTFT head (vocos 0.1.0 ISTFTHead, call site infer/utils_infer.py:472): per frame mag = min(exp(m), 1e2),
 * S = mag (cos p + i sin p); irfft(1024) * hann; overlap-add with hop 256; divide by the window envelope; trim
 * n_fft/2 at both ends (torch.istft center=True).  spec: fp32 [rows, lds] with log-magnitudes at columns [0,513) and
 * phases at [513,1026).  frames_out: fp32 [rows, 1024] windowed frames.  seg: int32 [num_segs, 4] = {row0, frames,
 * wav_offset, 0}; wav: fp32, hop*(frames-1) samples per segment, scaled by gains[seg] (RMS un-scaling,
 * infer/utils_infer.py:475-476) when gains != NULL. */
int f5_istft_frames(const float* spec, int64_t lds, int32_t rows, const float* window, float* frames_out, void* stream);
int f5_istft_ola(const float* frames, const float* window, const int32_t* seg, int32_t num_segs, int32_t max_wav_len,
                 float* wav, const float* gains, void* stream);

/* Prompt log-mel front-end (model/modules.py:75-101 `get_vocos_mel_spectrogram`; torchaudio MelSpectrogram n_fft 1024, hop 256,
 * periodic hann, center=True with reflect padding, power 1, HTK mel scale, norm None; then log(clamp(1e-5))).
 * wave: fp32 samples of all prompts back to back; seg: int32 [num_segs, 4] = {first sample, samples, first output row,
 * frames (= 1 + samples / 256)}; window: fp32 [1024]; fbank: fp32 [513, n_mels]; band: int32 [n_mels, 2] = the half-open
 * range of frequency bins where column m of fbank is non-zero; mel: fp32 [rows, ldm], one row per frame. */
int f5_mel_frames(const float* wave, const int32_t* seg, int32_t num_segs, int32_t max_frames, const float* window,
                  const float* fbank, const int32_t* band, int32_t n_mels, float* mel, int64_t ldm, void* stream);

/* Fault record.  Every mbarrier wait in the tcgen05 kernels carries a watchdog: a pipeline that makes no progress for 4 s of
 * wall time traps (the launch fails with cudaErrorLaunchFailure) instead of hanging the GPU.  Before trapping, the waiting
 * thread writes one 16-byte record into `mapped` — pinned host memory the device can address (e.g. cudaHostAlloc; >= 512 bytes,
 * zeroed) — which stays readable on the host after the CUDA context has died: u32[0] = 0xF5D00000 | kernel family (1 GEMM,
 * 2 attention), u32[1] = blockIdx.x | gridDim.x << 16, u32[2] = threadIdx.x | blockDim.x << 16, u32[3] = barrier shared-memory
 * address | parity << 31 (soak builds with -DF5_DIAG_FULL=1 append the raw state of every barrier of the CTA).  NULL switches it
 * off.  The reference has no counterpart (torch raises the CUDA error; core/managers.py:78-80 logs and re-raises). */
int f5_diag_enable(void* mapped);

/* Programmatic dependent launch.  Every kernel of the library is launched with the programmatic-stream-serialization attribute
 * and executes `griddepcontrol.wait` before it touches memory a predecessor may have written, so its set-up (barrier init, TMEM
 * allocation, tensor-map prefetch) overlaps the previous kernel's tail; inside a captured CUDA graph the edges become programmatic
 * dependencies.  Default on (environment F5_PDL=0 turns it off at load).  Returns the previous setting.  No reference counterpart:
 * the reference launches ~40 torch kernels per DiT block with full stream serialization (model/modules.py:558-572). */
int f5_set_pdl(int enabled);

/* Kernel variants of the two Vocos-side memory kernels (A/B measurements; results agree to float round-off).  v outside the
 * valid range only queries.  Both return the previous setting.  f5_dwconv7_ln: 2 = one warp per run of rows, row window in
 * registers (default; bit-identical to 1), 3 = channel-split with taps in registers too, 1 = one warp per row.  f5_istft_frames: 2 = real-input 512-point form (default),
 * 1 = 1024-point complex FFT.  Environment F5_DWCONV_V / F5_ISTFT_V set them at load. */
int f5_set_dwconv7_variant(int v);
int f5_set_istft_variant(int v);

/* Library / device info. */
int f5_device_check(void);      /* 0 if the current device is sm_100 */
const char* f5_version(void);

#ifdef __cplusplus
}
#endif
#endif /* F5_B200_H */
