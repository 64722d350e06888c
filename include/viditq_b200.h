/* viditq_b200 — C ABI of the B200-native low-bit path behind ViDiT-Q's qdiff QuantLayer operator.
 *
 * The reference (thu-nics/ViDiT-Q) has no FFI: its operator boundary is the Python nn.Module contract
 * QuantLayer.forward(input) (qdiff/models/quant_layer.py:99) and its five subclasses. These entry points are what a
 * patched QuantLayer.forward binds (ctypes stub in INTEGRATION.md). Conventions:
 *   - plain device pointers + sizes, no torch types; fp16 tensors are passed as `const void*` (IEEE binary16);
 *   - `stream` is a cudaStream_t; nothing here allocates, synchronises or reads back to the host;
 *   - return 0 (VQ_OK) or a negative VQ_ERR_* code; kernels never fall back to a CPU path.
 */
#ifndef VIDITQ_B200_H_
#define VIDITQ_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  VQ_OK = 0,
  VQ_ERR_ARG = -1,     /* bad shape / alignment / null pointer */
  VQ_ERR_DRIVER = -2,  /* cuTensorMapEncodeTiled entry point not found */
  VQ_ERR_TMAP = -3,    /* tensor-map encode failed */
  VQ_ERR_LAUNCH = -4,  /* kernel launch failed */
  VQ_ERR_UNSUPPORTED = -5
};

/* epilogue selector of vq_gemm_w8a8 / vq_linear_w8a8 */
enum {
  VQ_EPI_BIAS = 0,          /* out = y                         (QuantLayer / attn q,k,v / cross-attn linears) */
  VQ_EPI_GELU_TANH = 1,     /* out = gelu_tanh(y)              (mlp.fc1 -> act, modules.py:52-57)              */
  VQ_EPI_GATE_RESIDUAL = 2  /* out = res + gate * y            (stdit.py:109,118,123,127 gated residual)       */
};

/* sticky device status bits (vq_status_*): set by kernels, polled by the host outside the hot loop */
enum {
  VQ_STATUS_EPS_DEGENERATE = 1 /* a token row had delta < 1e-6: reference quirk Q4 (base_quantizer.py:220-223) */
};

/* Per-output-channel constants of a prepared weight, one 16-byte record per channel n:
 *   c1   = sum_k wq[n,k] - K * zw[n]   (so that sum_k (xq - zx)(wq - zw) = acc - zx*c1 - zw*rowsum_x)
 *   zw   = weight zero point (integer), dw = weight step size, bias = layer bias (0 if none)            */
typedef struct VqColParam {
  int32_t c1;
  int32_t zw;
  float dw;
  float bias;
} VqColParam;

/* Library / device ------------------------------------------------------------------------------------------- */
int vq_version(void);
int vq_num_sms(void);

/* (a2) WeightQuantizer.forward, base_quantizer.py:112-144, hoisted to load time (the reference re-runs it every call,
 * quant_layer.py:185). w: fp16 [N,K] row-major; delta, zp: fp16 [N] (per-output-channel buffers from ckpt.pth);
 * smooth: fp16 [K] channel_wise_scale or NULL (quant_layer.py:178: weight * channel_wise_scale, fp16 product);
 * bias: fp16 [N] or NULL. Outputs: codes u8 [N,K] (n_bits <= 8; values < 2^n_bits), col [N].                     */
int vq_prep_weight(const void* w, const void* delta, const void* zp, const void* smooth, const void* bias, int N,
                   int K, int n_bits, uint8_t* codes, VqColParam* col, void* stream);

/* (a1) DynamicActQuantizer.forward, dynamic_quantizer.py:16-45 + init_quant_params 'token' branch
 * base_quantizer.py:177-228. x: fp16, logically [G, rows, K] with element (g, r, k) at x[g*group_stride + r*ld + k];
 * statistics of token r are pooled over the G batch entries (quirk Q1). smooth: fp16 [K] or NULL (input /
 * channel_wise_scale, quant_layer.py:140). Outputs: codes u8 [G*rows, K] (row g*rows + r), delta/zp fp16 [rows],
 * rowsum i32 [G*rows]. status: device word, VQ_STATUS_EPS_DEGENERATE is OR-ed in if any delta < 1e-6.            */
int vq_act_quant(const void* x, int G, int rows, int K, int64_t group_stride, int64_t ld, const void* smooth,
                 int n_bits, uint8_t* codes, void* delta, void* zp, int32_t* rowsum, uint32_t* status, void* stream);

/* (N3) static activation scales: BaseQuantizer.forward with init_done (base_quantizer.py:112-144) on an ActQuantizer
 * whose delta / zero_point were calibrated by PTQ and loaded from ckpt.pth — per-tensor (`per_group: False`,
 * w8a8_naive.yaml: period = 1) or static per-token (period = rows).  x fp16 [M, K] (row pitch ld); delta, zp: DEVICE
 * fp16 [period], row m uses index m % period; smooth fp16 [K] or NULL.  Outputs codes u8 [M, K], rowsum i32 [M]; feed
 * vq_gemm_w8a8 with the same delta / zp and a_rows_period = period.                                                  */
int vq_act_quant_static(const void* x, int M, int K, int64_t ld, const void* delta, const void* zp, int period,
                        const void* smooth, int n_bits, uint8_t* codes, int32_t* rowsum, void* stream);

/* The same static quantiser with the producer's elementwise step fused in front, for the fused schedule of static
 * checkpoints (the reference applies them as separate torch ops: nn.GELU(approximate="tanh") between fc1 and fc2, timm Mlp;
 * LayerNorm + t2i_modulate of stdit.py:104,122 / blocks.py t2i_modulate): one pass instead of two.
 *   vq_gelu_act_quant_static           codes of h(gelu_tanh(x))
 *   vq_ln_modulate_act_quant_static    codes of h(h(LN(x) * h(1 + scale)) + shift); row m uses modulation vector
 *                                      m / rows_per_mod; K <= 2304 (the row is held in registers)                       */
int vq_gelu_act_quant_static(const void* x, int M, int K, int64_t ld, const void* delta, const void* zp, int period,
                             const void* smooth, int n_bits, uint8_t* codes, int32_t* rowsum, void* stream);
int vq_ln_modulate_act_quant_static(const void* x, const void* shift, const void* scale, int M, int K, int rows_per_mod,
                                    const void* delta, const void* zp, int period, const void* smooth, int n_bits,
                                    uint8_t* codes, int32_t* rowsum, void* stream);


/* (a1) on h(x + addv[(r / rows_per_add) % add_period]): the temporal position embedding added in front of block 0's
 * temporal attention (stdit.py:113-115, `x + self.pos_embed_temporal` on the "(B S) T C" view; in the (T S) token order the
 * embedding of frame t belongs to rows [t S, (t + 1) S): rows_per_add = S, add_period = T), an fp16 add, fused into the
 * quantiser of attn_temp's q|k|v input instead of a separate pass over the hidden tensor.  x fp16 [G*rows, K] contiguous,
 * addv fp16 [add_period, K]; K = 1152.  Outputs as vq_act_quant.                                                     */
int vq_add_act_quant(const void* x, const void* addv, int rows_per_add, int add_period, int G, int rows, int K,
                     const void* smooth, int n_bits, uint8_t* codes, void* delta, void* zp, int32_t* rowsum,
                     uint32_t* status, void* stream);

/* nn.GELU(approximate="tanh") + (a1), one pass: the activation between Mlp.fc1 and Mlp.fc2 (reference
 * opensora/models/stdit/stdit.py:109-111 timm Mlp with approx_gelu; PixArt_blocks / PixArtMS.py:60 likewise) applied to
 * the fp16 fc1 output x on its way into fc2's DynamicActQuantizer (quant_layer.py:137-140: smooth division, then
 * quantise). Same arguments and outputs as vq_act_quant; the statistics and codes are those of gelu(x) (/ smooth).   */
int vq_gelu_act_quant(const void* x, int G, int rows, int K, int64_t group_stride, int64_t ld, const void* smooth,
                      int n_bits, uint8_t* codes, void* delta, void* zp, int32_t* rowsum, uint32_t* status,
                      void* stream);

/* (a1) on a head-major attention output x fp16 [G * rows / S, H, S, head_dim] (what fused attention kernels emit):
 * same statistics and codes as vq_act_quant on the token-major view "(n H S D) -> (n S) (H D)", without the copy the
 * reference pays (blocks.py:189-191 transpose + reshape). head_dim = 72, H * head_dim = 1152.                      */
int vq_act_quant_heads(const void* x, int G, int rows, int H, int S, int head_dim, int n_bits, uint8_t* codes,
                       void* delta, void* zp, int32_t* rowsum, uint32_t* status, void* stream);

/* LayerNorm(eps=1e-6, no affine) + t2i_modulate (blocks.py:51: x*(1+scale)+shift) + (a1), one pass.
 * x: fp16 [G*rows, K]; shift, scale: fp16 [G * rows / rows_per_mod, K] — row (g, r) is modulated by vector
 * (g * rows + r) / rows_per_mod.  rows_per_mod == rows: one vector per pooled batch entry (the reference's batched
 * forward, statistics pooled over G, quirk Q1).  G == 1 with rows_per_mod < rows: several samples stacked along the
 * rows, each with its own modulation and its own un-pooled statistics == the reference's cfg_split=True, which runs
 * the cond / uncond halves as separate batch-1 forwards (qdiff/models/quant_model.py cfg_split branch).
 * smooth: fp16 [K] or NULL (divides the modulated tensor, quant_layer.py:140); y_out (optional, may be NULL): the
 * fp16 tensor the quantiser saw.                                                                                     */
int vq_ln_modulate_act_quant(const void* x, const void* shift, const void* scale, const void* smooth, int G, int rows,
                             int K, int rows_per_mod, int n_bits, void* y_out, uint8_t* codes, void* delta, void* zp,
                             int32_t* rowsum, uint32_t* status, void* stream);

/* (a3-a7) integer GEMM + dequant epilogue on prepared operands. a_codes u8 [M,K]; a_delta/a_zp fp16 [rows] indexed
 * by (m % a_rows_period) — pass a_rows_period = M when every row has its own scale; w_codes u8 [N,K]; out fp16
 * [M, ldo]. res fp16 [M, ldr], gate fp16 [M / rows_per_gate, N] for VQ_EPI_GATE_RESIDUAL.                          */
int vq_gemm_w8a8(const uint8_t* a_codes, const void* a_delta, const void* a_zp, const int32_t* a_rowsum,
                 int a_rows_period, const uint8_t* w_codes, const VqColParam* col, int M, int N, int K, int epi, const void* res,
                 int ldr, const void* gate, int rows_per_gate, void* out, int ldo, void* stream);

/* Per-channel activation maxima for smooth-quant: `input.abs().max(dim=-2)[0]` of quant_layer.py:116,119 (same lines in
 * stdit_quant_layer.py:31,34 / :122,125 / :235,238) — the live statistic of channel_wise_scale_type "dynamic" and of the
 * running-stat EMA that PixArt keeps switched on at inference (quirk Q17, t2i/scripts/quant_txt2img.py:297-300).
 * x fp16 [G, n, K] contiguous; out_bits u32 [G, K], ZERO-FILLED by the caller: on return out_bits[g,k] holds the fp16 bit
 * pattern of max_r |x[g,r,k]| (exact).  gelu != 0: statistics of h(gelu_tanh(x)) instead (fused schedules pass fc2 the
 * pre-activation).                                                                                                     */
int vq_col_absmax(const void* x, int G, int n, int K, int gelu, uint32_t* out_bits, void* stream);

/* (a3-a7, SURVEY.md section 8b entry 3) ONE call per QuantLayer-family forward: dynamic per-token activation quantiser
 * (a1) -> integer GEMM on prepared weight codes -> per-token x per-channel dequant -> bias | GELU | gated residual
 * epilogue; replaces quant_layer.py:185-211 / stdit_quant_layer.py:68-96 / dit_quant_layer.py:18-29.
 * x fp16 [G*rows, K] contiguous (statistics of token r pooled over the G batch entries, quirk Q1); smooth fp16 [K] or NULL;
 * ln_shift / ln_scale: fp16 [G*rows / rows_per_mod, K] or both NULL — when given, LayerNorm(eps 1e-6, no affine) +
 * t2i_modulate (blocks.py:51) run in front of the quantiser (the producers of the q|k|v and fc1 inputs, stdit.py:104,125).
 * Two schedules behind the one call, chosen by vq_linear_set_fused_policy (vq_linear_launch_count says which):
 *  - ONE kernel launch (vq_linear_fused_kernel; K == 1152, G in {1, 2, 4}, rows a multiple of 128 / G when G > 1): producer
 *    warps quantise the 128-row activation panel straight into the swizzled shared-memory operand of tcgen05.mma; no
 *    activation codes ever reach HBM;
 *  - quantise pass + persistent GEMM (vq_act_quant | vq_ln_modulate_act_quant -> vq_gemm_w8a8) through `workspace` — the
 *    DEFAULT: measured faster on B200 at every size inside a CUDA graph (profiles/r02_s3_linear_bench.md).
 * workspace: device scratch of at least vq_linear_workspace_bytes(G, rows, K) bytes (the library never allocates).
 * out_delta / out_zp (optional, may be NULL): fp16 [rows] per-token parameters, the DynamicActQuantizer side state.     */
int64_t vq_linear_workspace_bytes(int G, int rows, int K);
int vq_linear_launch_count(int G, int rows, int K);   /* 1 = fused kernel, 2 = quantise pass + GEMM */
/* mode 0: quantise pass + GEMM (default); 1: panel-resident fused kernel on every supported shape; -1: the same for
 * G * rows <= max_m; 2: overlapped — the persistent GEMM with quantiser warpgroups running ahead of its MMAs (K = 1152, G = 1,
 * >= 1024 rows, bias / gated-residual epilogue; other shapes as mode 0)                                                   */
int vq_linear_set_fused_policy(int mode, int64_t max_m);
int vq_linear_w8a8(const void* x, int G, int rows, int K, const void* smooth, const void* ln_shift, const void* ln_scale,
                   int rows_per_mod, int n_bits, const uint8_t* w_codes, const VqColParam* col, int N, int epi,
                   const void* res, int ldr, const void* gate, int rows_per_gate, void* out, int ldo, void* out_delta,
                   void* out_zp, void* workspace, int64_t workspace_bytes, uint32_t* status, void* stream);

/* (J2: north-star "INT4 x INT8 tensor-core GEMM") W4A8 with PACKED weight codes — the
 * w4a8_timestep_aware_cb.yaml layers (reference base_quantizer.py:129-144 at n_bits = 4).  vq_pack_u4 packs prepared u8 codes
 * (< 16) two per byte (low nibble = even k) into w_packed [N, K/2]; vq_linear_w4a8 is vq_linear_w8a8's fused kernel with the
 * weight operand streamed at half the bytes: the TMA stages packed tiles and two converter warps expand them into the swizzled
 * u8 operand of tcgen05.mma.kind::i8 (there is no INT4 integer MMA kind on sm_100a).  Same arguments otherwise; only the shapes
 * vq_linear_launch_count() reports as 1 are supported (VQ_ERR_UNSUPPORTED otherwise: call vq_linear_w8a8 with the u8 codes). */
int vq_pack_u4(const uint8_t* codes, int N, int K, uint8_t* packed, void* stream);
int vq_linear_w4a8(const void* x, int G, int rows, int K, const void* smooth, const void* ln_shift, const void* ln_scale,
                   int rows_per_mod, int n_bits, const uint8_t* w_packed, const VqColParam* col, int N, int epi,
                   const void* res, int ldr, const void* gate, int rows_per_gate, void* out, int ldo, void* out_delta,
                   void* out_zp, uint32_t* status, void* stream);

/* (section 8e (2), frame sharding) Pack / unpack the rows a frame-sharded forward exchanges around the temporal attention:
 * per-token QUANTISED activations travel as K code bytes + a 16-byte tail {delta fp16, zp fp16, rowsum i32, pad}.  The rows
 * form a 4-D array (d0, d1, d2, d3) in source order; the destination position of row (i0, i1, i2, i3) is
 * i0 s0 + i1 s1 + i2 s2 + i3 s3 (strides in rows) — the "B (T S) <-> rank-major" permutations of the reference's
 * sequence-parallel all-to-all (t2v/opensora/acceleration/communications.py) applied to codes instead of fp16 tensors.
 * unpack == 0: src = codes u8 [rows, K], arrays read, dst = rows with tails [rows, K + 16];
 * unpack != 0: src = rows with tails (source order), dst = codes u8 [rows, K], arrays written, all at the permuted position. */
int vq_row_pack(const uint8_t* src, uint8_t* dst, void* delta, void* zp, int32_t* rowsum, int rows, int K, int d1, int d2,
                int d3, int64_t s0, int64_t s1, int64_t s2, int64_t s3, int unpack, void* stream);

/* (a9) temporal self-attention of STDiT (stdit.py:112-118, blocks.py:151-195 on "(B S) T C"), reading q|k|v in place
 * from the fused GEMM output qkv fp16 [B*T*S, 3*H*head_dim] in the (T S) token layout; out fp16 [B*T*S, H*head_dim].
 * head_dim must be 72, T <= 16. scale = head_dim^-0.5.                                                             */
int vq_attn_temporal(const void* qkv, void* out, int B, int T, int S, int H, int head_dim, float scale, void* stream);

/* (a9 + a1) vq_attn_temporal with the DynamicActQuantizer of attn_temp.proj fused behind it (stdit.py:112-118 ->
 * stdit_quant_layer.py:155-165 on the "(B S) T C" view): one block owns all H = 16 heads of a (batch, position), so the token
 * rows it produces are complete and are quantised in place — u8 codes [B*T*S, 1152] + delta / zp fp16 [B*T*S] + rowsum i32,
 * every token on its own statistics (un-pooled: B == 1, or stacked independent calls).  Bit-identical to vq_attn_temporal
 * followed by vq_act_quant; the fp16 attention output is never written.  head_dim 72, H = 16, T <= 16.                     */
int vq_attn_temporal_quant(const void* qkv, int B, int T, int S, int H, int head_dim, float scale, const void* smooth,
                           int n_bits, uint8_t* codes, void* delta, void* zp, int32_t* rowsum, uint32_t* status,
                           void* stream);

/* (a9) spatial self-attention of STDiT (stdit.py:104-109, blocks.py:151-195 on "(B T) S C"; the reference calls
 * flash-attn / xformers there) and PixArt (PixArt_blocks.py AttentionKVCompress, sr_ratio 1): n_seq independent sequences
 * of S tokens, q|k|v read in place from the fused GEMM output qkv fp16 [n_seq*S, 3*H*head_dim]; out fp16
 * [n_seq*S, H*head_dim] token-major (no transpose copy in front of the projection). tcgen05 flash attention with TMEM
 * accumulators; head_dim must be 72 and S a multiple of 256. scale = head_dim^-0.5.                                  */
int vq_attn_spatial(const void* qkv, void* out, int n_seq, int S, int H, int head_dim, float scale, void* stream);

/* (J3, opt-in) spatial self-attention that consumes INT8 Q / K / V: both matrix products on tcgen05.mma.kind::i8.
 * NO reference counterpart — the reference never quantises Q / K / V or the probabilities (hooks commented out,
 * qdiff/models/quant_block.py:617-623, :630-632; blocks.py:169-188 runs flash-attn on the fp16 linear outputs) — so the
 * default path stays vq_attn_spatial and this entry carries its own tolerance (restated by
 * oracle/attn_i8_oracle.py).  Scheme: Q8 per (token, head), K8 = rint((K - mean_tokens K) / sk) per (64-key block, head),
 * V8 per (sequence, channel), P8 = rint(127 * 2^(x - m)) as u8; integer accumulation is exact.
 *   vq_attn_i8_workspace_bytes  size of the operand workspace (codes + scales) for a shape, -1 if unsupported
 *   vq_attn_i8_quantise         q|k|v fp16 [n_seq*S, 3*H*72] -> workspace (two passes: per-sequence statistics, codes)
 *   vq_attn_i8_attend           workspace -> out fp16 [n_seq*S, H*72] token-major (one persistent tcgen05 kernel)
 *   vq_attn_spatial_i8          both.  workspace: 256-byte aligned device memory; head_dim 72, S a multiple of 256.     */
int64_t vq_attn_i8_workspace_bytes(int n_seq, int S, int H, int head_dim);
int vq_attn_i8_quantise(const void* qkv, void* workspace, int n_seq, int S, int H, int head_dim, void* stream);
int vq_attn_i8_attend(const void* workspace, void* out, int n_seq, int S, int H, int head_dim, float scale, void* stream);
int vq_attn_spatial_i8(const void* qkv, void* out, void* workspace, int n_seq, int S, int H, int head_dim, float scale,
                       void* stream);

/* (a9) cross attention (blocks.py:292-310, xformers BlockDiagonalMask.from_seqlens([N]*B, y_lens)): q fp16
 * [B*N, H*head_dim]; kv fp16 [kv_rows, 2*H*head_dim] (k | v), kv_rows = sum(len) = the rows actually allocated (the
 * TMA tensor map is bounded by it); kv_start / kv_len: device int32 [B]; max_len <= 128.  N a multiple of 256 runs on
 * the tcgen05 flash-attention kernel (keys past a sample's length masked), other N on a small mma.sync kernel.          */
int vq_attn_cross(const void* q, const void* kv, void* out, const int32_t* kv_start, const int32_t* kv_len, int B,
                  int N, int H, int head_dim, int max_len, int64_t kv_rows, float scale, void* stream);

/* (a10 + N1) the sampler update of one denoise step for cfg_split models, fused: forward_with_cfg's combine
 * (iddpm/__init__.py:166-184: model_out / (1 + ptqd_k); eps = u + s (c - u) on channels [:3] (sic), the rest from the
 * conditional branch) + ddim_sample with eta = 0 (gaussian_diffusion.py:289-335, :540-552).  out_cond / out_uncond:
 * fp32 [n, c_out, inner] (c_out = 2 c: eps | learned sigma); x, x_new: fp32 [n, c, inner]; coef: DEVICE fp32 [4] =
 * {sqrt_recip_alphas_cumprod, sqrt_recipm1_alphas_cumprod, sqrt(alpha_bar_prev), sqrt(1 - alpha_bar_prev)} of the
 * step (device-resident so a captured CUDA graph can be replayed for every step).  Bit-identical to the eager CUDA op
 * sequence (every intermediate rounded to fp32 at the same places; `/ (1 + ptqd_k)` as ATen does it on CUDA: times the
 * fp32 reciprocal of the Python scalar).                                              */
int vq_cfg_ddim_step(const float* out_cond, const float* out_uncond, const float* x, const float* coef,
                     float cfg_scale, double ptqd_k, int n, int c_out, int c, int64_t inner, float* x_new, void* stream);

/* (N2) patch embedding fused with the spatial position embedding: STDiT `x_embedder` (PatchEmbed3D, Conv3d kernel = stride =
 * (1, ph, pw), blocks.py:60-110) + rearrange + `x + pos_embed` (stdit.py:255-258); PixArt's Conv2d patchify is the T = 1
 * case.  latent fp32 [B, Cin, T, Hh, Ww] (rounded to fp16 like x.to(dtype)); weight fp16 [C, Cin*ph*pw]; bias fp16 [C] or
 * NULL; pos fp16 [S, C] (S = Hh/ph * Ww/pw) or NULL; out fp16 [B, T*S, C].  Cin*ph*pw <= 16, C % 4 == 0, C <= 1280.      */
int vq_patch_embed(const float* latent, const void* weight, const void* bias, const void* pos, int B, int Cin, int T,
                   int Hh, int Ww, int ph, int pw, int C, void* out, void* stream);

/* status word helpers (host side; the only calls here that synchronise) */
int vq_status_read(const uint32_t* status_dev, uint32_t* host_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VIDITQ_B200_H_ */
