/* declip_b200 — C ABI of the B200-native CLIP/DeCLIP dual-encoder training path.
 *
 * The reference (Sense-GVT/DeCLIP) is pure Python/PyTorch and has NO FFI of its own; its
 * de-facto operator interface is the nn.Module forward signatures reached through
 * model_entry() (prototype/model/__init__.py:15-21).  This header is therefore the boundary a
 * maintainer would bind with ctypes from those modules; every entry point
 * cites the reference op site (file:line under /root/reference) whose arithmetic it replaces.
 *
 * Conventions
 *   - plain pointers + sizes, no torch types; all pointers are DEVICE pointers unless noted.
 *   - every call enqueues work on `stream` (a cudaStream_t) and returns immediately.
 *   - return 0 on success, non-zero on failure; dc_last_error() gives the text.  Never throws.
 *   - the library never allocates or frees device memory and keeps no pointer after a call
 *     returns (TMA descriptors are rebuilt/cached by value, keyed on the pointer + shape).
 *   - bf16 = storage dtype of activations and weight shadows; fp32 = master weights, biases,
 *     LayerNorm affine, statistics, gradients of parameters, loss.
 *   - activations are token-major [tokens, width] ("NLD": row = sample*L + position).
 */
#ifndef DECLIP_B200_H_
#define DECLIP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* dc_stream_t; /* cudaStream_t */

/* ------------------------------------------------------------------ library */
int dc_version(void);
const char* dc_last_error(void);
/* Binds the driver entry points, checks the device is compute capability 10.x, raises the
 * dynamic shared-memory limits of every kernel.  Must be called once per process/device. */
int dc_init(int device);
int dc_sm_count(void);
/* Persistent kernels (GEMM, attention, LayerNorm backward) size their grids to dc_sm_count() = device SMs - reserve.  A
 * data-parallel caller reserves the SMs its collective kernels occupy (NCCL: one CTA per channel) while gradient buckets
 * are in flight, so that a collective CTA never waits behind a whole persistent tile loop.  Returns the previous value. */
int dc_set_sm_reserve(int n);

/* ------------------------------------------------------------------ GEMM (tcgen05 / TMEM / TMA)
 * out[M,N] (op)= epilogue(alpha * sum_k A(m,k) * B(n,k))
 *   a_mn_major = 0: A stored [M,K] row-major (K contiguous);  1: A stored [K,M] row-major.
 *   b_mn_major = 0: B stored [N,K] row-major (K contiguous);  1: B stored [K,N] row-major.
 * Replaces every nn.Linear / F.linear / `@` on the hot path:
 *   base_transformer.py:33 (MultiheadAttention in/out proj), :35-41 (c_fc, c_proj),
 *   visual_transformer.py:56 (conv1 as patch GEMM), :72 (x @ proj), text_transformer.py:203,
 *   clip.py:140-141 (logit strips), and their autograd backward (dgrad / wgrad).
 * lda/ldb/ldo/ldo2/ldaux are row strides in ELEMENTS and must be multiples of 8; N % 8 == 0. */
enum {
  DC_EPI_BF16 = 0,       /* out(bf16)  = alpha*acc + bias                                   */
  DC_EPI_BF16_GELU = 1,  /* out2(bf16) = u = alpha*acc + bias ; out(bf16) = u*sigmoid(1.702u) */
  DC_EPI_BF16_RESID = 2, /* out(bf16)  = alpha*acc + bias + aux(bf16)                        */
  DC_EPI_BF16_DGELU = 3, /* out(bf16)  = alpha*acc * quickgelu'(aux(bf16))                   */
  DC_EPI_F32 = 4,        /* out(fp32)  = alpha*acc + bias                                   */
  DC_EPI_F32_ATOMIC = 5, /* out(fp32) += alpha*acc   (split-K allowed; red.global.add.v4.f32) */
  DC_EPI_F32_GROUPMAX16 = 6 /* FILIP late interaction (filip.py:93-104): out(fp32)[m, g] = max over the 16 columns of group
                             * g of alpha*acc, out2(uint8)[m, g] = index of that maximum (first one on ties); out / out2 are
                             * [M, N/16] with row strides ldo / ldo2 — the [M, N] score matrix never leaves TMEM */
};
typedef struct {
  const void* A; int lda; int a_mn_major;
  const void* B; int ldb; int b_mn_major;
  int M, N, K;
  int epilogue;
  float alpha;
  void* out; int ldo;
  void* out2; int ldo2;
  const float* bias;          /* [N] fp32 or NULL */
  const void* aux; int ldaux; /* bf16 [M,N] or NULL */
  int splits;                 /* split-K factor, 0 = auto (only DC_EPI_F32_ATOMIC may split) */
  int block_n;                /* 0 = auto, else 128 or 256 */
  float* colsum;              /* optional fp32 [N]: += sum over rows of the bf16-epilogue output (before rounding);
                                 the bias gradient when `out` is a dY — fused, no extra pass over dY */
  const float* alpha_dev;     /* optional DEVICE scalar multiplied into alpha (e.g. the clamped exp(logit_scale),
                                 clip.py:133-134) so no host sync is needed; NULL = 1 */
} dc_gemm_args;
int dc_gemm_bf16(const dc_gemm_args* args, dc_stream_t stream);
/* Select the 2-CTA (cta_group::2, 256x256 cluster tile) variant for problems with M >= 256 and N >= 256; returns the
 * previous setting.  Default 1 (measured +5-10 % over the 1-CTA 128x256 kernel, which remains the path for small
 * problems and can be forced with 0 or the DC_GEMM_2CTA=0 environment variable). */
int dc_set_gemm_2cta(int enable);
/* Host-side heuristic used for splits = 0: the split-K factor of an fp32-atomic GEMM with `tiles` output tiles and
 * `total_kb` 64-deep k-blocks on `workers` persistent CTAs (clusters): minimises waves x (k-blocks per split + 3). */
int dc_gemm_choose_splits(int tiles, int total_kb, int workers);
/* 1 (default; env DC_ATTN_TC=0 disables) = tcgen05/TMEM attention core for L <= 128, 0 = the mma.sync core (L <= 80). */
void dc_set_attention_tc(int enable);

/* ------------------------------------------------------------------ row kernels (HBM-bound)
 * LayerNorm over the last dim (eps 1e-5): base_transformer.py:10-18, visual_transformer.py:63,69,
 * text_transformer.py:194.  x,y bf16 [rows,width]; gamma,beta fp32; mean,rstd fp32 [rows] saved
 * for backward.  width % 256 == 0, width <= 1024. */
int dc_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                     int rows, int width, float eps, dc_stream_t stream);
/* dx(bf16) = [dres +] LN'(dy); dgamma,dbeta (fp32 [width]) are ACCUMULATED (+=). dres may be NULL.
 * dcol (fp32 [width], may be NULL) += sum_rows dx: the bias gradient of the Linear feeding this LayerNorm's input
 * (out_proj / c_proj), fused here so no separate column-sum pass over dx is needed. */
int dc_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                     const void* dres, void* dx, float* dgamma, float* dbeta, float* dcol, int rows, int width,
                     dc_stream_t stream);
/* out[c] += sum_r x[r,c] (bf16 in, fp32 accumulate): bias gradients of every Linear. */
int dc_colsum_bf16(const void* x, int ldx, float* out, int rows, int cols, dc_stream_t stream);
/* fp32 -> bf16 cast of n elements (weight shadows). */
int dc_cast_f32_bf16(const float* src, void* dst, size_t n, dc_stream_t stream);
/* dst[i] (+)= src[i] * scale, bf16 -> fp32 (gradient buckets travel as bf16 and return to the fp32 .grad buffer) */
int dc_cast_bf16_f32(const void* src, float* dst, size_t n, float scale, int accumulate, dc_stream_t stream);
/* table-driven multi-tensor cast: table (device) holds n_tensors entries {src, dst, numel}. */
typedef struct { const float* src; void* dst; unsigned long long numel; } dc_cast_entry;
int dc_multi_cast_f32_bf16(const dc_cast_entry* table_dev, int n_tensors, unsigned long long max_numel,
                           dc_stream_t stream);

/* Fused multi-tensor AdamW step (torch.optim.AdamW semantics: p *= 1 - lr*wd; Adam moments; bias correction from
 * `step`): the optimiser of the reference configs (config.yaml optimizer.type AdamW).  table (device): one entry per
 * parameter tensor, fp32 master / moments; `group` indexes the host arrays lr[] / weight_decay[] (n_groups <= 32,
 * passed to the kernel by value, so a per-iteration LR schedule never touches the device table) that carry the
 * reference's param groups (utils/misc.py:267-412: no decay on biases / norms / logit_scale).  `shadow` (may be NULL)
 * is the bf16 copy of the parameter that the tensor-core GEMMs read: the kernel rewrites it from the updated master
 * in the same pass, so a step never leaves a stale shadow behind. */
typedef struct {
  float* param; const float* grad; float* exp_avg; float* exp_avg_sq; void* shadow;
  unsigned long long numel; int group; int reserved;
} dc_adamw_entry;
int dc_adamw_multi(const dc_adamw_entry* table_dev, int n_tensors, unsigned long long max_numel, const float* lr_host,
                   const float* wd_host, int n_groups, float beta1, float beta2, float eps, int step,
                   dc_stream_t stream);

/* ------------------------------------------------------------------ attention (L <= 128, head_dim 64)
 * tcgen05/TMEM core (attention_tc.cu) for L <= 128 and even head counts when L <= 64; mma.sync core (L <= 80) otherwise
 * or after dc_set_attention_tc(0).
 * softmax(q k^T / sqrt(64) [+ causal mask]) v per (sample, head): base_transformer.py:44-48 via
 * nn.MultiheadAttention; causal mask text_transformer.py:136-142.
 * qkv bf16 [batch*L, 3*width] (q | k | v, heads contiguous inside each), out bf16 [batch*L, width],
 * lse fp32 [batch*heads*L] saved for backward. */
int dc_attention_fwd(const void* qkv, void* out, float* lse, int batch, int L, int heads, int causal,
                     dc_stream_t stream);
/* dbias (fp32 [3*width], may be NULL) += column sums of dqkv: the in_proj_bias gradient, fused.  The K slice of that
 * gradient is identically zero in exact arithmetic (every row of dS sums to zero); the tcgen05 core leaves it
 * untouched, the mma.sync core adds its rounding noise. */
int dc_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, float* dbias,
                     int batch, int L, int heads, int causal, dc_stream_t stream);

/* ------------------------------------------------------------------ embeddings
 * ViT: conv1 with kernel == stride == patch is a GEMM over non-overlapping patches
 * (visual_transformer.py:56-59).  images fp32 NCHW [B,3,R,R] (channel offset/stride allow the
 * DeCLIP 6-channel two-view layout, declip.py:199) -> patches bf16 [B*G*G, 3*P*P] (c,py,px order). */
int dc_patchify(const float* images, long long sample_stride, void* patches, int batch, int res, int patch,
                dc_stream_t stream);
/* tokens[b,0,:] = cls + pos[0]; tokens[b,1+p,:] = patch_out[b*G2+p,:] + pos[1+p]
 * (visual_transformer.py:60-62).  patch_out bf16 [B*G2,width]; cls fp32 [width]; pos fp32 [G2+1,width]. */
int dc_vit_assemble(const void* patch_out, const float* cls, const float* pos, void* tokens, int batch, int g2,
                    int width, dc_stream_t stream);
/* text_transformer.py:188-190: x[b,l,:] = table[ids[b,l],:] + pos[l,:]; table,pos fp32; ids int64. */
int dc_text_embed(const long long* ids, const float* table, const float* pos, void* x, int batch, int L,
                  int width, dc_stream_t stream);
/* dtable[ids[b,l],:] += dx[b,l,:]  (fp32 atomics) ; dpos[l,:] += sum_b dx[b,l,:].  last_row (int32 [batch], may be
 * NULL) = row index of each sample's EOT token: rows after it have an exactly-zero gradient and are skipped. */
int dc_text_embed_bwd(const long long* ids, const void* dx, float* dtable, float* dpos, const int* last_row, int batch,
                      int L, int width, dc_stream_t stream);
/* rows gather / scatter (cls token rows, EOT rows: visual_transformer.py:69, text_transformer.py:203):
 * dst[i,:] = src[idx[i],:]  /  dst[idx[i],:] = src[i,:]  (bf16 rows, int32 row indices). */
int dc_gather_rows(const void* src, const int* idx, void* dst, int n, int width, dc_stream_t stream);
int dc_scatter_rows(const void* src, const int* idx, void* dst, int n, int width, dc_stream_t stream);
/* eot[b] = b*L + argmax_l ids[b,l]  (text_transformer.py:203) */
int dc_eot_index(const long long* ids, int* eot, int batch, int L, dc_stream_t stream);

/* ------------------------------------------------------------------ contrastive head
 * clip.py:129-130: y = x / (||x|| + eps) row-wise; x fp32 [n,dim] -> y bf16 and/or y_f32 (either may be
 * NULL), inv fp32 [n] = 1/(||x||+eps) (may be NULL). */
int dc_l2norm_fwd(const float* x, void* y, float* y_f32, float* inv, int n, int dim, float eps, dc_stream_t stream);
/* dx(fp32) = inv*dy - x * <dy,x> * inv^2 / ||x||  with inv = 1/(||x|| + eps); dy, x fp32 [n,dim]. */
int dc_l2norm_bwd(const float* dy, const float* x, float* dx, int n, int dim, float eps, dc_stream_t stream);
/* loss.py:40-50 (ClipInfoCELoss) on one logit strip, fused with misc.py:415-428 (accuracy top-1/top-5):
 * logits fp32 [rows,cols] (row stride ld), label[r] = labels ? labels[r] : label0 + r  (the int64 `labels`
 * array serves the DeCLIP masked-language-model head, declip.py:326-334).  skip_col (int32 [rows], may be NULL):
 * one column per row excluded from the softmax — the self-similarity of NT_Xent / NT_Xent_gather (nt_xent.py:19-26,
 * 74-86), whose [positive | negatives] cross-entropy is exactly a row CE with that column masked.
 *   *loss_sum += sum_r (lse_r - logit[r,label_r])      (fp32 atomic; caller zeroes and divides by rows)
 *   *top1/*top5 += #{r : #(logit[r,:] > logit[r,label_r]) < 1 / < 5}   (may be NULL)
 *   lse_out[r] = log-sum-exp of row r (saved for backward). */
int dc_ce_strip_fwd(const float* logits, int ld, int rows, int cols, int label0, const long long* labels,
                    const int* skip_col, float* loss_sum, int* top1, int* top5, float* lse_out, dc_stream_t stream);
/* dlogits([rows,cols], row stride lddl; bf16, or fp32 when out_f32) =
 *   gscale_host * (*gscale_dev) * (softmax(row) - onehot(label));  gscale_dev may be NULL (treated as 1). */
int dc_ce_strip_bwd(const float* logits, int ld, int rows, int cols, int label0, const long long* labels,
                    const int* skip_col, const float* lse, const float* gscale_dev, float gscale_host, void* dlogits,
                    int lddl, int out_f32, dc_stream_t stream);
/* *out += sum_i a[i]*b[i]  (fp32; used for d logit_scale = exp(ls)/s * sum dlogits*logits, clip.py:133-141) */
int dc_dot_f32(const float* a, const float* b, size_t n, float* out, dc_stream_t stream);

/* ------------------------------------------------------------------ DeCLIP heads (all fp32, [batch, <= 1024])
 * nn.BatchNorm1d (+ fused ReLU) of the SimSiam projector / predictor MLPs — declip.py:33-130.  training: batch
 * statistics, running stats updated with momentum (unbiased variance) when non-NULL; eval: running statistics. */
int dc_batchnorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* save_mean,
                     float* save_rstd, float* running_mean, float* running_var, int rows, int channels, float eps,
                     float momentum, int training, int relu, dc_stream_t stream);
/* dgamma/dbeta are ACCUMULATED; `y` is the forward output (its sign is the ReLU mask). */
int dc_batchnorm_bwd(const float* dy, const float* x, const float* y, const float* gamma, const float* save_mean,
                     const float* save_rstd, float* dx, float* dgamma, float* dbeta, int rows, int channels,
                     int training, int relu, dc_stream_t stream);
/* SimsiamLoss D(p, z) = mean_r cos(p_r, stopgrad z_r) — loss_functions/loss.py:52-84.
 * fwd: cosv[r]; bwd: dp[r,:] = gscale * grow[r] * d cos / d p  (grow may be NULL = 1). */
int dc_cosine_rows_fwd(const float* p, const float* z, float* cosv, int n, int dim, dc_stream_t stream);
int dc_cosine_rows_bwd(const float* p, const float* z, const float* grow, float gscale, float* dp, int n, int dim,
                       dc_stream_t stream);
/* Nearest-neighbour memory bank lookup — nnclr_modules/nn_memory_bank.py:42-65: idx[r] = argmax_c sim[r,c]
 * (top-1), then dst[i,:] = bank[idx[i],:] (fp32 rows). */
int dc_argmax_rows(const float* x, int ld, int rows, int cols, int* idx, dc_stream_t stream);
int dc_gather_rows_f32(const float* src, const int* idx, float* dst, int n, int width, dc_stream_t stream);
/* dst[idx[i],:] += src[i,:] (bf16 rows, distinct idx): EOT-row gradient added into the dense ln_final gradient. */
int dc_add_rows(const void* src, const int* idx, void* dst, int n, int width, dc_stream_t stream);

/* ------------------------------------------------------------------ FILIP token-wise late interaction (filip.py:71-106)
 * score1[b,j] = <d1[b,j,:], sum_m d2[b,m,:]>, score2[b,m] = <d2[b,m,:], sum_j d1[b,j,:]> : the row / column sums of
 * the per-pair cross-logit matrix (filip.py:79-81) that rank tokens for the top-16 selection.  d1 [batch,n1,dim],
 * d2 [batch,n2,dim] fp32 (normalised). */
int dc_token_scores(const float* d1, const float* d2, int batch, int n1, int n2, int dim, float* score1, float* score2,
                    dc_stream_t stream);
/* out[i,l] = mean_{j<n} max_{m<group} G[i*n+j, l*group+m]; arg[(i*n+j)*ncand + l] = argmax m   (filip.py:103-104).
 * arg may be NULL; with group = 1 this is the token mean of the [batch*n, ncand] group maxima that the GEMM epilogue
 * DC_EPI_F32_GROUPMAX16 leaves behind (the score matrix G itself is then never materialised). */
int dc_groupmax_mean_fwd(const float* G, int ldg, int batch, int n, int group, int ncand, float* out, int ldo,
                         unsigned char* arg, dc_stream_t stream);
/* dG (bf16 [batch*n, ncand*group]) = one-hot(arg) * dout[i,l] / n */
int dc_groupmax_mean_bwd(const float* dout, int ldd, const unsigned char* arg, int batch, int n, int group, int ncand,
                         void* dG, int ldg, dc_stream_t stream);
/* the same for a block of candidate columns: `arg` and `dout` point at the block's first column, `lda` / `ldd` are the row
 * strides of the full matrices — the backward walks the candidates in blocks so the one-hot operand stays bounded
 * (<= 512 MiB) however many ranks contribute candidates */
int dc_groupmax_mean_bwd_ex(const float* dout, int ldd, const unsigned char* arg, int lda, int batch, int n, int group,
                            int ncand, void* dG, int ldg, dc_stream_t stream);
int dc_add_rows_f32(const float* src, const int* idx, float* dst, int n, int width, dc_stream_t stream);

/* ------------------------------------------------------------------ ModifiedResNet support (modified_resnet.py)
 * Activations are NHWC bf16 = [rows = B*H*W, C]; 1x1 convolutions are dc_gemm_bf16, 3x3 convolutions are
 * im2col (K ordered (ky,kx,c)) + dc_gemm_bf16.
 * stem conv1 (3->32, k3 s2 p1, modified_resnet.py:150): col bf16 [B*H/2*W/2, 32], k = (ky*3+kx)*3 + c, 27..31 zero. */
int dc_im2col_stem(const float* images, long long sample_stride, void* col, int batch, int H, int W, dc_stream_t stream);
/* 3x3 / pad 1 / stride 1: col bf16 [B*H*W, 9*C]; col2im is its transpose (the convolution's dgrad). */
int dc_im2col3x3(const void* in, void* col, int batch, int H, int W, int C, dc_stream_t stream);
int dc_col2im3x3(const void* dcol, void* din, int batch, int H, int W, int C, dc_stream_t stream);
/* nn.AvgPool2d(2) on NHWC, forward (in [B,H,W,C] -> out [B,H/2,W/2,C]) or backward (in = d pooled, out = d input). */
int dc_avgpool2(const void* in, void* out, int batch, int H, int W, int C, int backward, dc_stream_t stream);
/* nn.BatchNorm2d in training mode over the rows of x bf16 [rows,C] (+ optional residual add, + optional ReLU):
 * y = [relu]((x-mean)*rstd*gamma + beta [+ res]).  scratch: fp32 [2*C].  Running stats updated when non-NULL. */
int dc_bn2d_fwd(const void* x, const float* gamma, const float* beta, const void* res, void* y, float* mean, float* rstd,
                float* running_mean, float* running_var, float* scratch, long long rows, int C, float eps, float momentum,
                int training, int relu, dc_stream_t stream);
/* dx = BN'(g), g = dy * [y > 0]; dres (may be NULL) = g (gradient of the residual input); dgamma/dbeta accumulate. */
int dc_bn2d_bwd(const void* dy, const void* x, const void* y, const float* gamma, const float* mean, const float* rstd,
                void* dx, void* dres, float* dgamma, float* dbeta, float* scratch, long long rows, int C, int relu,
                dc_stream_t stream);
int dc_add_bf16(const void* a, const void* b, void* out, size_t n, dc_stream_t stream);
/* AttentionPool2d token assembly (modified_resnet.py:72-74): tokens[b,0] = mean_p x[b,p] + pos[0];
 * tokens[b,1+p] = x[b,p] + pos[1+p]; backward dx[b,p] = dtok[b,1+p] + dtok[b,0]/P. */
int dc_attnpool_assemble(const void* x, const float* pos, void* tokens, int batch, int P, int C, dc_stream_t stream);
int dc_attnpool_assemble_bwd(const void* dtokens, void* dx, int batch, int P, int C, dc_stream_t stream);

/* ------------------------------------------------------------------ implicit-GEMM 3x3 convolution (conv_igemm.cu)
 * out[B*H*W, Cout] (bf16, NHWC) = conv3x3(x[B*H*W, C] NHWC bf16, pad 1, stride 1) with w [Cout, 9*C] bf16 ordered
 * (ky, kx, c) — modified_resnet.py:23,151-154.  The A operand of every k-block is a 4-D TMA box of x shifted by the tap
 * offset (out-of-bounds zero fill = padding): no im2col matrix.  The input gradient is the same call on dy with
 * w' [C, 9*Cout], w'[ci, (ky, kx, co)] = w[co, (2-ky, 2-kx, ci)].  Needs C % 64 == 0, Cout % 64 == 0, W <= 128. */
int dc_conv3x3_igemm_supported(int H, int W, int C, int Cout);
int dc_conv3x3_igemm(const void* x, const void* w, void* out, int batch, int H, int W, int C, int Cout, dc_stream_t stream);
/* dw[Cout, 9*C] (fp32, (ky, kx, c) order) += sum over pixels of dy[p, co] * x[p + tap offset, ci]: the contraction runs over
 * spatial TMA boxes of both NHWC tensors (MN-major operands), split over CTAs with fp32 atomics.  C, Cout multiples of 32
 * (32-channel tensors ride in half-empty 64-channel boxes); rows wider than 64 pixels are cut into equal parts. */
int dc_conv3x3_wgrad_igemm_supported(int H, int W, int C, int Cout);
int dc_conv3x3_wgrad_igemm(const void* dy, const void* x, float* dw, int batch, int H, int W, int C, int Cout,
                           dc_stream_t stream);

/* ------------------------------------------------------------------ fused distributed contrastive head (head.cu)
 * Replaces, for a symmetric image/text pair, the whole chain clip.py:129-146 (normalise, AllGather, two logit
 * strips) + loss_functions/loss.py:40-50 (ClipInfoCELoss) + utils/misc.py:415-428 (accuracy) and its autograd backward
 * (clip.py:43-49 all_reduce of the gathered-tensor gradients) by three launches; no [b, N] strip reaches HBM.
 *   dc_head_prepare : y_f = x_f / (||x_f|| + eps_f) as bf16 into out_rows[b, n_feats * e] (feature f at columns
 *                     [f e, (f+1) e)) — this rank's slice of the gather buffer — and clears the workspace accumulators.
 *   dc_head_forward : per direction d: logits[i, j] = s <X_d[i], Y_d[j]>, i local, j over all n gathered rows; writes
 *                     ws[out..]: sum_i CE_d (2 floats), sum_i (E_softmax[logit] - label logit) (2 floats; times
 *                     exp(logit_scale)/s it is d sum CE / d logit_scale), top-1 / top-5 counts of direction 0, and the
 *                     row log-sum-exps ws[lse..] = [2][b].  strips[d] != NULL additionally stores the fp32 strip
 *                     (compat mode for callers that consume logits).
 *   dc_head_backward: exch = [n_ranks][2 b + 2] — every rank's {lse[0][b], lse[1][b], g[0], g[1]} with g[d] the upstream
 *                     gradient of sum CE_d (an all-gather of ws[lse .. g+2)); writes dx_out[d] = d loss / d raw X_d
 *                     feature (through the normalisation), including — when `cross` — the terms the reference receives
 *                     through AllGather.backward from the other ranks' strips.
 * y_src: n_src == 1 -> one buffer holding all n rows (NCCL all-gather result); n_src == world -> y_src[r] = rank r's
 * b rows (peer-mapped symmetric memory; the kernel reads them over NVLink, no collective).  e in {256,512,768,1024}. */
typedef struct {
  int b, n, e;              /* local rows, gathered rows, feature dim */
  int ld;                   /* row stride of the feature buffers in elements (n_feats * e) */
  int n_src;
  int row0;                 /* global row index of local row 0 (rank * b): labels are row0 + i (loss.py:45) */
  int x_off[2], y_off[2];   /* column offset of the X / Y feature of each direction inside a row */
  int cross;                /* 1: the two directions are transposes of each other (ClipInfoCELoss(li, lt)) */
  int ld_strip;
  const void* x_base;       /* bf16 [b, ld]: this rank's rows */
  const void* y_src[8];
  float* ws;                /* dc_head_workspace_floats(b, e) floats; carries the scale computed by dc_head_prepare */
  float* strips[2];
} dc_head_args;
size_t dc_head_workspace_floats(int b, int e);
/* offsets (floats) into the workspace: {out[16], lse[2][b], g[2], alab[2][b], ea[2][b], counters, dxn[2][b][e], total};
 * out: 0,1 sum CE_d; 2,3 sum (E_softmax[logit] - label logit)_d; 4 / 5 top-1 / top-5 counts of direction 0; 8 s; 9 exp(ls);
 * 10 d loss / d logit_scale. */
int dc_head_layout(int b, int e, long long* offsets8);
/* logit_scale: the raw parameter on the device — s = min(exp(logit_scale), scale_max) in the forward, d s / d logit_scale
 * = exp(logit_scale) (the reference clamps `.data`, clip.py:133-134) — or NULL for the constant scale `scale_max`.
 * ws[out + 8] = s, ws[out + 9] = exp(logit_scale); dc_head_backward leaves d loss / d logit_scale in ws[out + 10]. */
int dc_head_prepare(const float* const* feats, const float* eps, int n_feats, int b, int e, void* out_rows, float* ws,
                    const float* logit_scale, float scale_max, dc_stream_t stream);
/* As dc_head_prepare, and additionally stores this rank's normalised rows at rows [row0, row0 + b) of every peer's gather
 * buffer (peer_buffers: n_peers device pointers into symmetric memory, the OTHER ranks' [n, n_feats * e] buffers; out_rows
 * then points at row row0 of this rank's own buffer): a one-shot all-gather by peer stores over NVLink fused into the
 * normalisation kernel.  After a device barrier every rank holds all n rows locally (dc_head_args.n_src = 1). */
int dc_head_prepare_push(const float* const* feats, const float* eps, int n_feats, int b, int e, void* out_rows,
                         void* const* peer_buffers, int n_peers, long long row0, float* ws, const float* logit_scale,
                         float scale_max, dc_stream_t stream);
int dc_head_forward(const dc_head_args* a, dc_stream_t stream);
/* g_own[2] (device): upstream gradients of this rank's sum CE_d; exch as described above (its own rank's g slots are
 * not read: g_own is used instead, so a single-rank step needs no copy at all). */
int dc_head_backward(const dc_head_args* a, const float* g_own, const float* exch, const float* const* x_raw,
                     const float* eps2, float* const* dx_out, dc_stream_t stream);

/* ------------------------------------------------------------------ host-side text front end (no device work)
 * Byte-pair-encoding tokenizer = prototype/model/utils/text_utils/simple_tokenizer.py:66-134 (OpenAI CLIP BPE plus the
 * extra <|mask|> token: vocabulary = merges + 515) and the truncation / zero padding of TextTransformer.tokenize
 * (text_encoder/text_transformer.py:144-170).  `merges_text` is the DECOMPRESSED content of
 * bpe_simple_vocab_16e6.txt.gz.  Texts are NUL-terminated UTF-8, already cleaned and lower-cased by the caller
 * (simple_tokenizer.py:53-63,126).  Thread-safe; dc_bpe_tokenize splits the batch over `threads` host threads. */
typedef struct dc_bpe dc_bpe_t;
dc_bpe_t* dc_bpe_create(const char* merges_text, long long nbytes);      /* NULL on error (dc_last_error) */
void dc_bpe_destroy(dc_bpe_t* h);
int dc_bpe_vocab_size(const dc_bpe_t* h);
int dc_bpe_token_id(const dc_bpe_t* h, const char* token);              /* e.g. "<|endoftext|>"; -1 if absent */
/* ids of ONE text without SOT / EOT; returns the token count (may exceed `capacity`; only that many are written), < 0 on error */
long long dc_bpe_encode(dc_bpe_t* h, const char* text, int* ids_out, long long capacity);
/* ids[n, context_length] (int64) = SOT + tokens + EOT, truncated to SOT + first context_length-2 tokens + EOT, zero
 * padded; lengths[n] (optional) = tokens written per row. */
int dc_bpe_tokenize(dc_bpe_t* h, const char* const* texts, int n, int context_length, long long* ids, int* lengths,
                    int threads);
/* As above; raw_ascii[i] != 0 marks a caption the caller has NOT cleaned but knows to be printable ASCII without '&'
 * (str.isascii() and str.isprintable()): the library strips it, collapses blanks and lower-cases it itself — the whole of
 * basic_clean / whitespace_clean / lower for such text — so the host spends no per-caption Python time on it. */
int dc_bpe_tokenize_ex(dc_bpe_t* h, const char* const* texts, const unsigned char* raw_ascii, int n, int context_length,
                       long long* ids, int* lengths, int threads);

/* ------------------------------------------------------------------ composite encoders (C++ executors)
 * One call runs a whole tower forward (or backward) as a fixed launch sequence on `stream`.
 * Weight/grad tables are arrays of device pointers in the order documented in encoder.h order
 * (see declip_b200/csrc/encoder.cu: DC_LAYER_* / DC_VIT_* / DC_TEXT_* indices). */
typedef struct {
  int layers, width, heads, seq_len, causal;
  int batch;
  int embed_dim;
  int res, patch; /* ViT only */
  int vocab;      /* text only */
} dc_tower_cfg;
size_t dc_tower_workspace_bytes(const dc_tower_cfg* cfg);
/* dense_out (bf16 [batch*(L-1), width], may be NULL): the patch tokens of the last block, `x[:, 1:, :]` that
 * VisualTransformer.forward returns with return_dense (visual_transformer.py:68); ddense is its gradient (or NULL). */
/* Host-side progress hook of the tower BACKWARD executors: `cb(user, l)` is called from the launching thread right
 * after the kernels of transformer layer l have been enqueued (l counts down from layers-1 to 0), whenever
 * l % every == 0.  Every gradient of layers >= l is then complete in stream order, so a data-parallel caller can record
 * an event and start reducing that slice while the earlier layers still run (utils/dist.py:63-74 overlaps per parameter
 * through autograd hooks).  Per calling thread; cb = NULL clears it. */
typedef void (*dc_progress_cb)(void* user, int layer);
int dc_set_backward_progress_cb(dc_progress_cb cb, void* user, int every);

int dc_vit_forward(const dc_tower_cfg* cfg, const float* images, long long sample_stride,
                   const void* const* w_bf16, const float* const* w_f32, void* workspace, float* features,
                   void* dense_out, dc_stream_t stream);
int dc_vit_backward(const dc_tower_cfg* cfg, const float* dfeatures, const void* ddense, const void* const* w_bf16,
                    const float* const* w_f32, float* const* grads, void* workspace, dc_stream_t stream);
/* The pre-projection feature of the last forward held by `workspace`: ln_post(class token) for the ViT
 * (`return_feature` of VisualTransformer.forward, visual_transformer.py:70,78-79), ln_final(EOT token) for the text
 * tower; bf16 [batch, width].  dc_vit_backward_pre = dc_vit_backward with its gradient `dpre` (bf16 [batch, width],
 * may be NULL) added to the gradient that reaches ln_post (used by SLIP's `predictor_sim`, slip.py:230-238). */
int dc_tower_pre_features(const dc_tower_cfg* cfg, const void* workspace, void* out_bf16, dc_stream_t stream);
int dc_vit_backward_pre(const dc_tower_cfg* cfg, const float* dfeatures, const void* dpre, const void* ddense,
                        const void* const* w_bf16, const float* const* w_f32, float* const* grads, void* workspace,
                        dc_stream_t stream);
/* words_out (bf16 [batch*L, width], may be NULL): ln_final applied to EVERY token — the `words_feat` that
 * TextTransformer.forward returns with mask_type / return_dense (text_transformer.py:194-201).  When it was requested
 * in forward, backward must be called with dense=1 and dwords (bf16 [batch*L, width] or NULL = zero). */
int dc_text_forward(const dc_tower_cfg* cfg, const long long* ids, const void* const* w_bf16,
                    const float* const* w_f32, void* workspace, float* features, void* words_out, dc_stream_t stream);
int dc_text_backward(const dc_tower_cfg* cfg, const long long* ids, const float* dfeatures, int dense,
                     const void* dwords, const void* const* w_bf16, const float* const* w_f32, float* const* grads,
                     void* workspace, dc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DECLIP_B200_H_ */
