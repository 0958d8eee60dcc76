/* drb200.h — C ABI of libdrb200.so, the B200 (sm_100a) implementation of the DiffusionRenderer denoising hot path.
 *
 * The reference (eggsbenedicto/DiffusionRenderer-ComfyUI) has no FFI of its own: its hot path is Python calling
 * PyTorch library ops.  Each entry point below therefore names the reference Python call site(s) it replaces
 * (file:line into the reference tree).  INTEGRATION.md shows the ctypes binding a maintainer adds on the
 * reference side.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`; bf16 tensors are row-major, 16-byte aligned;
 *   - `stream` is the caller's cudaStream_t passed as void* (e.g. torch.cuda.current_stream().cuda_stream);
 *     calls only enqueue work on it and never synchronise the device;
 *   - return value: 0 = OK, negative = DRB_ERR_*; drb_last_error() returns a thread-local message;
 *   - nothing is retained past a call (no hidden allocations, no global state besides a tensor-map entry point
 *     and the SM count cache); scratch buffers are provided by the caller.
 */
#ifndef DRB200_H_
#define DRB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRB_OK 0
#define DRB_ERR_INVALID (-1)   /* bad shape / alignment / argument   -> Python ValueError */
#define DRB_ERR_CUDA (-2)      /* a CUDA runtime / driver call failed -> Python RuntimeError */
#define DRB_ERR_UNSUPPORTED (-3)

/* Epilogue selectors for drb_gemm_bf16 */
#define DRB_EPI_STORE 0          /* out = bf16(acc)                                        nn.Linear, bias=False            */
#define DRB_EPI_GELU 1           /* out = bf16(gelu_erf(bf16(acc)))                        CleanGeneralDIT.py:454-457       */
#define DRB_EPI_GATED_RESIDUAL 2 /* out = bf16(resid + bf16(gate[n] * bf16(acc)))          CleanGeneralDIT.py:517           */
#define DRB_EPI_QKV_NORM_ROPE 3  /* internal to drb_gemm_qkv_norm_rope                                                      */

#define DRB_CP_MAX_RANKS 8 /* GPUs of one context-parallel group */
#define DRB_CP_FLAG_SLOTS 4   /* flag array of a rank: uint32 [DRB_CP_FLAG_SLOTS][DRB_CP_MAX_RANKS] + status, 256 bytes */
#define DRB_CP_STATUS_WORD 63 /* local flag word set to 1 when an in-kernel wait timed out                          */

/* Cross-GPU ordering folded into a kernel (context parallelism, csrc/cp_sync.cuh).  flag_ptrs[j] = rank j's peer-mapped,
 * zeroed flag array.  wait_epoch != 0: the kernel's operand-loading thread first waits until word [wait_slot][r] of the
 * LOCAL array is >= wait_epoch for every rank r — the peers' P2P stores into this GPU's input buffer have landed.
 * signal_epoch != 0: once every CTA of the kernel has fenced its stores, the last one writes word [signal_slot][rank] =
 * signal_epoch into every rank's array; `counter` is a local, zero-initialised uint32 (reset by the kernel).  Epochs
 * increase per slot and are identical on all ranks.  timeout_ms = 0 means 60 s; on a timeout the kernel sets the local
 * status word and carries on (the host checks it) instead of trapping. */
typedef struct drb_cp_sync {
  void* const* flag_ptrs;
  uint32_t* counter;
  int world, rank;
  int signal_slot, wait_slot;
  uint32_t signal_epoch, wait_epoch;
  uint32_t timeout_ms;
} drb_cp_sync;

const char* drb_last_error(void);
int drb_version(void);
/* 1 when the current device is compute capability 10.x (the only target), else 0. */
int drb_device_supported(void);

/* ---- dense projections --------------------------------------------------------------------------------------
 * out[M,N] = epilogue( A[M,K] @ W[N,K]^T ), bf16 in / fp32 accumulate in TMEM / bf16 out (tcgen05 + TMA).
 * Replaces every nn.Linear on the token stream: to_q/to_k/to_v (CleanGeneralDIT.py:273-276, fused as one
 * [3D,D] weight), to_out (:304), layer1/layer2 (:454,:460), x_embedder.proj['1'] (:417), final_layer.linear (:590).
 * lda/ldw/ldo/ldr are row pitches in elements (multiples of 8).  `resid`/`gate` only for DRB_EPI_GATED_RESIDUAL
 * (`out` may alias `resid`).  cta_group: 1 = one CTA per 128x256 tile, 2 = CTA pair per 256x256 tile, 0 = auto. */
int drb_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* out, int64_t ldo,
                  int M, int N, int K, int epilogue, const void* resid, int64_t ldr, const void* gate,
                  int cta_group, void* stream);
/* The same, with cross-GPU ordering folded in: under context parallelism the out-projection's A operand is the buffer the
 * peers' attention epilogues store into, so its TMA producer waits for their flags (sync->wait_*) before the first load. */
int drb_gemm_bf16_sync(const void* A, int64_t lda, const void* W, int64_t ldw, void* out, int64_t ldo,
                       int M, int N, int K, int epilogue, const void* resid, int64_t ldr, const void* gate,
                       int cta_group, const drb_cp_sync* sync, void* stream);

/* Ring form (context parallelism when heads are not split: every GPU keeps its query rows and visits the K/V blocks of
 * all GPUs in ring order): one launch = one K/V block.  Running state per (row, head): state_ml [q_len, H, 2] fp32 =
 * (reference in log2 units, row sum), state_o [q_len, H*128] fp32 = un-normalised output.  `first` initialises the state,
 * every launch merges its block into it (m' = max(m, M); O' = O 2^(m-m') + O_blk 2^(M-m'); l likewise), `last` writes the
 * normalised bf16 rows to o [q_len, ld_o] instead of the state.  Same kernel, other epilogue. */
int drb_attention_bf16_ring(const void* q, const void* k, const void* v, int64_t ld_qkv, void* o, int64_t ld_o, float* state_o,
                            float* state_ml, int q_len, int kv_len, int num_heads, int first, int last, void* stream);
/* The same with the logit-bound certificate of drb_attention_bf16_bounded (device float, nullable). */
int drb_attention_bf16_ring_bounded(const void* q, const void* k, const void* v, int64_t ld_qkv, void* o, int64_t ld_o,
                                    float* state_o, float* state_ml, int q_len, int kv_len, int num_heads, int first, int last,
                                    const float* max_abs_logit, void* stream);

/* Fused QKV projection: [q | k | v] = A[M,K] @ W[3D,K]^T with, in the epilogue, per-head RMSNorm of q and k (weights
 * wq, wk [128], eps 1e-6) and RoPE from the bf16 tables cos_tab / sin_tab [M,128] — i.e. drb_gemm_bf16 followed by
 * drb_qk_norm_rope without the round trip of q and k through HBM (CleanGeneralDIT.py:268-297).  world == 0: rows are
 * stored to out [M, ldo >= 3D].  world > 0 (context parallelism): head h's q / k / v row of local token s is stored
 * into peer_ptrs[h / (H/world)] at row row0 + s of a [S, peer_ld] buffer laid out q | k | v (each (H/world)*128 wide) —
 * the Ulysses all-to-all is the GEMM's epilogue (P2P stores over NVLink, overlapping the MMA of the next tile). */
int drb_gemm_qkv_norm_rope(const void* A, int64_t lda, const void* W, int64_t ldw, void* out, int64_t ldo, int M, int D,
                           int K, const void* wq, const void* wk, const void* cos_tab, const void* sin_tab,
                           void* const* peer_ptrs, int world, int64_t peer_ld, int row0, void* stream);
/* The same with `batch` independent token sequences stacked along M (row r belongs to sequence r / rows_per_batch): the
 * five G-buffer passes of one clip (nodes.py:187-205) or cond + uncond under CFG (model_diffusion_renderer.py:230-232) run
 * as ONE GEMM, so the weights stream once per batch and the tile grid has no short last wave at M = S/P.  cos_tab /
 * sin_tab are [M,128] (the per-sequence table repeated).  world > 0: sequence b, head h, token s goes to row row0 + s of
 * the owner's [S, peer_ld] buffer at column (sect*batch + b)*(H/world)*128 + (h mod H/world)*128, sect = 0 q, 1 k, 2 v —
 * i.e. (b, h) pairs look like batch*(H/world) heads to drb_attention_bf16_cp_batched.  rows_per_batch >= 32.
 * `sync` (nullable): the last CTA signals "all q / k / v rows of this rank are stored" (sync->signal_*). */
int drb_gemm_qkv_norm_rope_batched(const void* A, int64_t lda, const void* W, int64_t ldw, void* out, int64_t ldo, int M, int D,
                                   int K, const void* wq, const void* wk, const void* cos_tab, const void* sin_tab,
                                   void* const* peer_ptrs, int world, int64_t peer_ld, int row0, int batch, int rows_per_batch,
                                   const drb_cp_sync* sync, void* stream);

/* ---- self-attention ------------------------------------------------------------------------------------------
 * o[s, h*128 + d] = softmax_j(q[s,h,:]·k[j,h,:] / sqrt(128)) v[j,h,d]; no mask, no dropout, head_dim 128.
 * Replaces PytorchDotProductAttention.forward / F.scaled_dot_product_attention (CleanGeneralDIT.py:181-203), with
 * the head-flattened (S, H*128) output the reference's to_out expects (SURVEY.md defect D1).
 * q,k,v: row s at q + s*ld_qkv elements, head h at column h*128 (e.g. three column blocks of one [S,3D] buffer).
 * `kv_len` keys/values, `q_len` queries (equal for the single-GPU path; they differ under context parallelism). */
int drb_attention_bf16(const void* q, const void* k, const void* v, int64_t ld_qkv, void* o, int64_t ld_o,
                       int q_len, int kv_len, int num_heads, void* stream);
/* The same with a caller-supplied certificate: *max_abs_logit is a DEVICE float B with |q.k| / sqrt(128) <= B (natural
 * units) for every query / key pair (read by the kernel, so choosing the flavour costs no host synchronisation).  For
 * B <= 39 the max-free lazy softmax runs (no per-element max; provably overflow-free under the bound whatever the key
 * order); above it, or with NULL (= drb_attention_bf16), the per-tile-max softmax runs, which is exact for unbounded
 * logits.  drb_qk_logit_bound computes B for q / k that went through the DiT's per-head RMSNorm. */
int drb_attention_bf16_bounded(const void* q, const void* k, const void* v, int64_t ld_qkv, void* o, int64_t ld_o,
                               int q_len, int kv_len, int num_heads, const float* max_abs_logit, void* stream);
/* bound[l] = sqrt(128) * max|wq[l,:]| * max|wk[l,:]| * 1.02 for `layers` pairs of per-head RMSNorm weights wq, wk
 * [layers,128] (bf16): after x -> rmsnorm(x) * w every head row has norm <= sqrt(128) max|w| (RoPE is a rotation; 1.02
 * covers the bf16 roundings), so |q.k| / sqrt(128) <= bound (Cauchy-Schwarz).  CleanGeneralDIT.py:23-33,:288-297. */
int drb_qk_logit_bound(const void* wq, const void* wk, float* bound, int layers, void* stream);

/* Context-parallel form (csrc/cp.cu): the same kernel over the H/P heads this GPU owns and all tokens;
 * output row r is stored into o_peers[r / rows_per_rank] (a peer-mapped [rows_per_rank, ld_o] buffer of the GPU that
 * owns token r) at local row r % rows_per_rank, column col0 + h*128 — the inverse Ulysses exchange is the epilogue. */
int drb_attention_bf16_cp(const void* q, const void* k, const void* v, int64_t ld_qkv, void* const* o_peers, int world,
                          int64_t ld_o, int q_len, int kv_len, int num_heads, int rows_per_rank, int col0, void* stream);
/* Batched sequences (see drb_gemm_qkv_norm_rope_batched): num_heads = batch * heads_per_batch attention problems over the
 * same token range; "head" g = b*heads_per_batch + h reads column g*128 of q / k / v and its output row r is stored at
 * local row b*batch_rows + (r mod rows_per_rank), column col0 + h*128, of the owner's [batch*batch_rows, ld_o] buffer.
 * max_abs_logit: as drb_attention_bf16_bounded.  `sync` (nullable): wait for the peers' q / k / v stores before the first
 * load (sync->wait_*) and signal "all output rows of this rank are stored" from the last CTA (sync->signal_*). */
int drb_attention_bf16_cp_batched(const void* q, const void* k, const void* v, int64_t ld_qkv, void* const* o_peers, int world,
                                  int64_t ld_o, int q_len, int kv_len, int num_heads, int rows_per_rank, int col0,
                                  int heads_per_batch, int batch_rows, const float* max_abs_logit, const drb_cp_sync* sync,
                                  void* stream);

/* ---- fused elementwise family ----------------------------------------------------------------------------------
 * AdaLN: out = bf16(bf16(bf16(LN(x)) * bf16(1+scale)) + shift), LN over `D` without affine, eps 1e-6, fp32 stats
 * (CleanGeneralDIT.py:7-11,:481,:506).  If `add_vec` != NULL the residual stream is first updated in place,
 * x <- bf16(x + bf16(add_gate * add_vec)), which is the whole degenerate cross-attention sub-block (:512-517 with a
 * one-token context). */
int drb_adaln_modulate(void* x, void* out, const void* shift, const void* scale, const void* add_gate,
                       const void* add_vec, int rows, int D, void* stream);

/* q,k <- RoPE(RMSNorm_head(q,k)) in place on the [S, 3D] projection buffer (v untouched).
 * Per head of 128: bf16(rmsnorm_fp32(x) * w) (CleanGeneralDIT.py:23-33,:288-289) then
 * bf16(bf16(x*cos) + bf16(rotate_half(x)*sin)) with cos/sin = bf16(cos/sin(angle)) (:67-80).
 * `cos_tab`,`sin_tab`: bf16 [S,128] tables built by the host from the reference's angle formula (:94-159). */
int drb_qk_norm_rope(void* qkv, int64_t ld, const void* wq, const void* wk, const void* cos_tab,
                     const void* sin_tab, int S, int num_heads, void* stream);

/* y[n] = bf16( sum_k W[n,k] * act(x[k]) ) for one bf16 vector x; act: 0 = identity, 1 = SiLU (computed in fp32 on the
 * bf16 input and rounded to bf16 first, as nn.SiLU on a bf16 tensor does).  `add` (nullable) is added afterwards in
 * bf16: y = bf16(y + add).  Replaces the B=1 nn.Linear GEMVs of the timestep / AdaLN-LoRA path
 * (CleanGeneralDIT.py:357-364, :483-488, :500-501, :558-572) and the context projections of cross-attention (:275-276,:304). */
int drb_gemv_bf16(const void* W, int64_t ldw, const void* x, void* y, const void* add, int N, int K, int act,
                  void* stream);
/* Batched form: `count` independent GEMVs sharing x; W_b = W + b*w_batch_stride, y_b = y + b*y_batch_stride,
 * add_b = add + b*add_batch_stride (elements; add may be NULL; add_batch_stride may be 0 to share). */
int drb_gemv_bf16_batched(const void* W, int64_t ldw, int64_t w_batch_stride, const void* x, int64_t x_batch_stride,
                          void* y, int64_t y_batch_stride, const void* add, int64_t add_batch_stride, int count,
                          int N, int K, int act, void* stream);

/* Timestep embedding: sigma (fp32 scalar on device) -> s~ = bf16(sigma); e = bf16([cos(s~ w_i), sin(s~ w_i)]),
 * w_i = exp(-ln(1e4) i / (D/2)) (CleanGeneralDIT.py:321-335, :664); emb = bf16(rmsnorm_fp32(e) * w_aff) (:666).
 * Writes e_out[D] and emb_out[D] (bf16). */
int drb_sigma_embedding(const float* sigma, const void* w_aff, void* e_out, void* emb_out, int D, void* stream);

/* EDM input scaling + patchify of the noisy latent (model_diffusion_renderer.py:30-44; CleanGeneralDIT.py:409-414):
 * x_in = bf16(fp32(x_t) / sqrt(sigma^2 + 0.25)) scattered into token rows of `tokens` [S, ld_tok] at feature
 * c*4 + m*2 + n for channel c in [0,C).  x_t: bf16 [C,T,H,W]. */
int drb_scale_patchify(const void* x_t, const float* sigma, void* tokens, int64_t ld_tok, int C, int T, int H, int W,
                       void* stream);
/* Patchify constant channels (latent condition + ones padding mask, CleanGeneralDIT.py:669-675) once per pass:
 * src bf16 [C,T,H,W] -> tokens[:, (c0+c)*4 + m*2 + n]; if `ones_channel` >= 0 that channel's 4 features are set to 1,
 * and features in [zero_from, ld_tok) are zeroed (K padding). */
int drb_patchify_condition(const void* src, void* tokens, int64_t ld_tok, int c0, int C, int T, int H, int W,
                           int ones_channel, int zero_from, void* stream);

/* Final unpatchify + Euler update (CleanGeneralDIT.py:709-716; model_diffusion_renderer.py:46-82, :232):
 * F[c,t,2h+ph,2w+pw] = y[s, (ph*2+pw)*C + c];  optional CFG  F = bf16(Fc + bf16(g * bf16(Fc - Fu)));
 * den = c_skip x + c_out F;  x_next = bf16(x + (x - den)/sigma * (sigma_next - sigma)), all fp32 inside.
 * y_cond / y_uncond: bf16 [S, ld_y]; x_t in / x_next out: bf16 [C,T,H,W] (may alias); F_out (nullable): bf16 [C,T,H,W].
 * x_t == x_next == NULL skips the Euler update (plain unpatchify into F_out; sigma pointers may then be NULL). */
int drb_unpatchify_euler(const void* y_cond, const void* y_uncond, int64_t ld_y, float guidance, const float* sigma,
                         const float* sigma_next, const void* x_t, void* x_next, void* F_out, int C, int T, int H,
                         int W, void* stream);

/* Stand-alone EDM scheduler ops on a flat bf16 tensor of n elements (the public CleanEDMEulerScheduler methods;
 * the sampler loop itself uses the fused drb_scale_patchify / drb_unpatchify_euler):
 *   scale_model_input: out = bf16(fp32(x) * 1/sqrt(sigma^2 + 0.25))                    model_diffusion_renderer.py:30-44
 *   step:              out = bf16(x + (x - (c_skip x + c_out F)) / sigma * (sigma_next - sigma))         :46-82 */
int drb_edm_scale_input(const void* x, const float* sigma, void* out, int64_t n, void* stream);
int drb_edm_euler_step(const void* model_output, const void* x, const float* sigma, const float* sigma_next, void* out,
                       int64_t n, void* stream);

/* Decode post-process (diffusion_renderer_pipeline.py:300-318): optional normal re-normalisation blend, then
 * (1+v).clamp(0,2)/2*255 -> uint8 (truncating), BCTHW -> BTHWC.  video: bf16 [3,T,H,W]; out: uint8 [T,H,W,3]. */
int drb_postprocess_u8(const void* video, void* out_u8, int T, int H, int W, int normalize_normal, void* stream);

/* ==== CV8x8x8 causal video tokenizer ===========================================================================
 * The reference only wraps diffusers.AutoencoderKLCosmos (CleanVAE.py:3,18,50-51,59-60); the entry points below are the
 * operators of that class.  Inside the tokenizer activations are channels-last per frame,
 * bf16 [T][H][W][C]; pixel-space clips and latents at its boundary are planar bf16 [C][T][H][W] (BCTHW, B = 1). */

/* temporal source-frame mapping of a convolution (kt taps, output frame t, tap dt) */
#define DRB_TMODE_CAUSAL 0 /* src = max(t + dt - (kt-1), 0): first frame replicated in front (CosmosCausalConv3d)        */
#define DRB_TMODE_DOWN2 1  /* src = max(2t + dt - 2, 0): cat[x0, x] then causal conv with temporal stride 2 (Downsample) */
#define DRB_TMODE_UP2 2    /* src = (max(t + dt - (kt-1), 0) + 1) / 2: causal conv of repeat_interleave(x, 2)[1:] (Upsample) */
/* residual term added in the epilogue, indexed by the output position (t, oh, ow) */
#define DRB_RES_NONE 0
#define DRB_RES_SAME 1          /* resid[t, oh, ow]                               resnet / attention skip               */
#define DRB_RES_FRAME_UP2 2     /* resid[(t+1)/2, oh, ow]                         "+ x" of the temporal upsample        */
#define DRB_RES_POOL_HW 3       /* mean of resid[t, 2oh+{0,1}, 2ow+{0,1}] (zero beyond the edge)  "+ avg_pool" spatial  */
#define DRB_RES_POOL_T 4        /* mean of resid[max(2t-1,0)], resid[2t]          "+ avg_pool" temporal                 */
#define DRB_RES_NEAREST_UP_HW 5 /* resid[t, oh/2, ow/2]                           "+ x" of the spatial upsample         */

typedef struct drb_conv3d_args {
  const void* x;     /* bf16 [T_in][H_in][W_in][Cin], Cin % 64 == 0                                                     */
  const void* w;     /* bf16 [Cout][kt][kh][kw][Cin]                                                                    */
  const void* bias;  /* bf16 [Cout]                                                                                     */
  void* out;         /* bf16 [T_out][H_out*out_scale][W_out*out_scale][Cout], Cout % 16 == 0                            */
  const void* resid; /* bf16 [*][resid_H][resid_W][Cout] or NULL (resid_mode DRB_RES_NONE)                              */
  double* stats;     /* nullable: [T_out][2] += (sum, sum of squares) of the stored outputs of each frame               */
  int T_in, H_in, W_in, Cin;
  int T_out, H_out, W_out, Cout; /* positions iterated: out position (h, w) reads source rows stride_hw*h + dy - pad_h  */
  int kt, kh, kw, pad_h, pad_w, stride_hw, tmode;
  int out_scale, out_off_h, out_off_w; /* position (h, w) is stored at (h*out_scale + out_off_h, w*out_scale + out_off_w) */
  int resid_mode, resid_H, resid_W;
} drb_conv3d_args;

/* Implicit-GEMM causal 3-D convolution on tcgen05 tensor cores (TMA box per tap; OOB zero fill = spatial padding).
 * Replaces CosmosCausalConv3d / CosmosConvProjection3d / the convolutions of CosmosDownsample3d and CosmosUpsample3d. */
int drb_conv3d_cl(const drb_conv3d_args* args, void* stream);

/* CosmosPatchEmbed3d ("haar", patch 4): x bf16 [C][T][H][W] -> out bf16 [(T+3)/4][H/4][W/4][64*C]: first frame repeated
 * 4x, two levels of 3-D Haar DWT (sub-bands lll,llh,...,hhh time-first, / sqrt(8) per level).  T = 1 + 4k, H,W % 4 == 0. */
int drb_haar_patch(const void* x, void* out, int C, int T, int H, int W, void* stream);
/* CosmosUnpatcher3d: in bf16 [Tp][Hp][Wp][64*C] -> out bf16 [C][4*Tp-3][4*Hp][4*Wp] (first 3 frames dropped). */
int drb_haar_unpatch(const void* in, void* out, int C, int Tp, int Hp, int Wp, void* stream);
/* CosmosUnpatcher3d of a 3-channel video with the decode post-process of diffusion_renderer_pipeline.py:300-318 fused into
 * its store: in bf16 [Tp][Hp][Wp][192] -> out_u8 uint8 [4*Tp-3][4*Hp][4*Wp][3] = drb_postprocess_u8 of what
 * drb_haar_unpatch would have written (bit-identical), without the planar bf16 video ever reaching HBM. */
int drb_haar_unpatch_u8(const void* in, void* out_u8, int Tp, int Hp, int Wp, int normalize_normal, void* stream);

/* CosmosCausalGroupNorm(num_groups = 1) = one (mean, var) per frame over H*W*C, eps 1e-6, affine.
 * drb_frame_stats_cl zeroes and fills stats[T][2] = (sum, sum of squares); drb_conv3d_cl accumulates the same in its
 * epilogue (the caller zeroes).  drb_groupnorm_apply_cl: out = act(bf16(norm(x)*gamma + beta)), act = SiLU if `silu`. */
int drb_frame_stats_cl(const void* x, double* stats, int T, int64_t per_frame, void* stream);
int drb_groupnorm_apply_cl(const void* x, void* out, const double* stats, const void* gamma, const void* beta, int T,
                           int64_t hw, int C, int silu, void* stream);

/* Spatial attention of the mid block (one head of dim C over the H*W tokens of a frame) is scores GEMM -> row softmax
 * -> P.V GEMM, both GEMMs being drb_gemm_bf16:  s[r, :cols] <- softmax(scale * s[r, :cols]) in place, s[r, cols:ld] <- 0. */
int drb_softmax_rows(void* s, int64_t ld, int rows, int cols, float scale, void* stream);
/* out[c][r] = in[r][c] for r < rows, 0 for rows <= r < ld_out (V^T as the K-major operand of the P.V GEMM). */
int drb_transpose_bf16(const void* in, int64_t ld_in, void* out, int64_t ld_out, int rows, int cols, void* stream);
/* The same attention for C = 512 (the CV8x8x8 mid blocks) as ONE flash-attention kernel: qkv bf16 [frames][n][ld >= 1536] laid
 * out q | k | v (512 columns each), out bf16 [frames][n][ld_o >= 512]; softmax(q k^T / sqrt(512)) v per frame, no score
 * matrix in memory.  A CTA owns 128 query rows and one 256-column half of v / out (TMEM: two score buffers + the half
 * accumulator), so q k^T is computed twice: 1.5x the algorithmic MMA work instead of five passes over 2 n^2 bytes. */
int drb_spatial_attention_d512(const void* qkv, int64_t ld, void* out, int64_t ld_o, int frames, int n, void* stream);
/* Causal temporal attention (one head of dim C over the T frames of a pixel): qkv bf16 [T][hw][3C] -> out [T][hw][C]. */
int drb_temporal_attention_cl(const void* qkv, void* out, int T, int64_t hw, int C, void* stream);

/* Boundary layouts: planar [C][thw] <-> channels-last [thw][Cpad] (extra channels zero), values scaled by `scale`
 * (the sigma_data factor of model_diffusion_renderer.py:146,156 rides along here). */
int drb_planar_to_cl(const void* x, void* out, int C, int Cpad, int64_t thw, float scale, void* stream);
int drb_cl_to_planar(const void* in, void* out, int C, int Cpad, int64_t thw, float scale, void* stream);

/* Per-chunk latent statistics of the upstream chunking tokenizer (pretrained_vae.py:131-152, unused by the reference's
 * live path): x, out bf16 [rows][hw]; mean, std bf16 [rows] (one pair per (batch, channel, latent frame) row).
 * mode 0 (after encode): out = bf16(bf16(x - mean) / std) (:142); mode 1 (before decode): out = bf16(bf16(x * std) + mean) (:150). */
int drb_latent_normalize(const void* x, const void* mean, const void* std, void* out, int rows, int64_t hw, int mode, void* stream);

/* ==== environment-map conditioning of the forward renderer ===================================
 * The reference projects the panorama with nvdiffrast (preprocess_envmap.py); these entry points do it on the device
 * without it.  fp32, channels-last RGB.
 * drb_envmap_latlong_to_cubemap: apply_hdr_preprocessing (:263-286: x brightness, NaN -> 0, clamp [0, 65504], optional
 *   horizontal flip, roll by roll_px) + latlong_to_cubemap_official (:161-206) -> cube [6][R][R][3].
 * drb_envmap_project: render_projection_from_panorama (:408-467) — pixel (h, w) of the (H, W) lat-long grid looks the
 *   cube map up in direction -latlong_vec (linear filter, cube boundary), flipped in both axes — + hdr_mapping_official
 *   (:119-140) -> env_ldr, env_log [H][W][3] in [0, 1].
 * drb_envmap_tonemap: tonemap_image_direct (:469-526) — bilinear resize (align_corners=False) + the same tone mapping. */
int drb_envmap_latlong_to_cubemap(const float* latlong, int He, int We, float brightness, int flip, int roll_px, float* cube,
                                  int R, void* stream);
int drb_envmap_project(const float* cube, int R, float* env_ldr, float* env_log, int H, int W, void* stream);
int drb_envmap_tonemap(const float* src, int Hs, int Ws, float* env_ldr, float* env_log, int H, int W, void* stream);

/* ==== context parallelism over NVLink peer memory ==============================================
 * One video's tokens are split contiguously over `world` GPUs (one process each).  Everything but self-attention is
 * token-local; the Ulysses exchange around it is fused into the producing kernels as P2P stores. */

/* drb_qk_norm_rope for the S_local tokens of this GPU, with head h's q / k / v rows stored into dst_ptrs[h / (H/world)]
 * — the peer that owns the head — at row row0 + s of a [S, dst_ld] buffer laid out q | k | v, each (H/world)*128 wide.
 * cos_tab / sin_tab: the rows of the local tokens.  Replaces CleanGeneralDIT.py:288-297 + the all-to-all. */
int drb_cp_qk_norm_rope_scatter(const void* qkv, int64_t ld, const void* wq, const void* wk, const void* cos_tab,
                                const void* sin_tab, int S_local, int num_heads, void* const* dst_ptrs, int world,
                                int64_t dst_ld, int row0, void* stream);
/* Device-side barrier between the `world` GPUs: flag_ptrs[j] is GPU j's flag array (uint32[world], peer-mapped, zeroed);
 * every GPU passes the same, increasing `epoch`.  All earlier work of this stream (including P2P stores) is visible to
 * the peers once they leave the barrier.  Never blocks the host. */
int drb_cp_barrier(void* const* flag_ptrs, int rank, int world, uint32_t epoch, void* stream);
/* Peer-shareable device memory: cudaMalloc (zero-filled) + CUDA IPC handle export / import (one process per GPU). */
#define DRB_PEER_HANDLE_BYTES 64
int drb_peer_alloc(int64_t bytes, void** ptr);
int drb_peer_free(void* ptr);
int drb_peer_export(const void* ptr, void* handle64);
int drb_peer_import(const void* handle64, void** ptr);
int drb_peer_close(void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* DRB200_H_ */
