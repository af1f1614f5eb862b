// declip_b200 — composite encoder executors: the whole ViT / text-transformer tower forward and
// backward as fixed launch sequences driven from C++ (one C-ABI call per tower per direction), so
// the Python side issues 4 calls per training step instead of ~1200 op dispatches.
//
// Reference call chain being replaced (paths under /root/reference/prototype/model):
//   image_encoder/visual_transformer.py:55-82  VisualTransformer.forward
//   text_encoder/text_transformer.py:183-204   TextTransformer.forward ('Transformer' branch)
//   image_encoder/base_transformer.py:50-53    ResidualAttentionBlock.forward (x12 per tower)
// and their autograd backward.  conv1 is frozen in the reference (visual_transformer.py:12,45-51):
// no wgrad/dgrad is computed for the patch embedding.
//
// Pointer tables (all device pointers):
//   w_bf16 : per layer l -> [4l+0] in_proj_weight [3D,D]  [4l+1] out_proj.weight [D,D]
//                           [4l+2] mlp.c_fc.weight [4D,D] [4l+3] mlp.c_proj.weight [D,4D]
//            then ViT : [4NL+0] conv1.weight viewed [D, 3*P*P]   [4NL+1] proj [D,E]
//                 text: [4NL+0] text_projection.weight [E,D]
//   w_f32  : per layer l -> [8l+0] ln_1.weight [8l+1] ln_1.bias [8l+2] in_proj_bias [8l+3] out_proj.bias
//                           [8l+4] ln_2.weight [8l+5] ln_2.bias [8l+6] c_fc.bias    [8l+7] c_proj.bias
//            then ViT : +0 class_embedding +1 positional_embedding +2 ln_pre.w +3 ln_pre.b +4 ln_post.w +5 ln_post.b
//                 text: +0 token_embedding.weight +1 positional_embedding +2 ln_final.w +3 ln_final.b +4 text_projection.bias
//   grads  : per layer l -> [12l+0] in_proj_weight [12l+1] in_proj_bias [12l+2] out_proj.weight [12l+3] out_proj.bias
//                           [12l+4] ln_1.weight [12l+5] ln_1.bias [12l+6] c_fc.weight [12l+7] c_fc.bias
//                           [12l+8] c_proj.weight [12l+9] c_proj.bias [12l+10] ln_2.weight [12l+11] ln_2.bias
//            then ViT : +0 class_embedding +1 positional_embedding +2 ln_pre.w +3 ln_pre.b +4 ln_post.w +5 ln_post.b +6 proj
//                 text: +0 token_embedding.weight +1 positional_embedding +2 ln_final.w +3 ln_final.b
//                       +4 text_projection.weight +5 text_projection.bias
//   Every gradient is ACCUMULATED (+=) in fp32; the caller owns zeroing.
#include "common.cuh"
#include "internal.h"

namespace dc {

#define DC_TRY(expr)            \
  do {                          \
    int rc__ = (expr);          \
    if (rc__ != 0) return rc__; \
  } while (0)

struct Carver {
  uint8_t* base;
  size_t off;
  explicit Carver(void* b) : base(static_cast<uint8_t*>(b)), off(0) {}
  template <typename T>
  T* take(size_t n) {
    off = (off + 255) & ~static_cast<size_t>(255);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

constexpr int MAX_LAYERS = 48;

struct ProgressHook { dc_progress_cb cb; void* user; int every; };
static thread_local ProgressHook g_progress = {nullptr, nullptr, 1};

struct LayerWs {
  bf16 *ln1, *qkv, *attn, *xmid, *ln2, *u, *h;
  float *mean1, *rstd1, *mean2, *rstd2, *lse;
};

struct TowerWs {
  bf16* xs[MAX_LAYERS + 1];
  LayerWs layer[MAX_LAYERS];
  // tower head / tail
  bf16 *patches, *patch_out, *tokens_pre;
  float *mean_pre, *rstd_pre;
  int* row_idx;
  bf16 *rows, *rows_ln;
  float *mean_post, *rstd_post;
  float *mean_f, *rstd_f;  // ln_final over every token (dense / MLM mode)
  // backward scratch
  bf16 *dxa, *dxb, *dqkv, *du, *dtmp, *dfeat, *drow, *drow2;
  size_t bytes;
};

static void carve(const dc_tower_cfg& c, void* base, TowerWs& w) {
  Carver cv(base);
  const size_t M = static_cast<size_t>(c.batch) * c.seq_len;
  const size_t D = c.width;
  for (int l = 0; l <= c.layers; ++l) w.xs[l] = cv.take<bf16>(M * D);
  for (int l = 0; l < c.layers; ++l) {
    LayerWs& L = w.layer[l];
    L.ln1 = cv.take<bf16>(M * D);
    L.qkv = cv.take<bf16>(M * 3 * D);
    L.attn = cv.take<bf16>(M * D);
    L.xmid = cv.take<bf16>(M * D);
    L.ln2 = cv.take<bf16>(M * D);
    L.u = cv.take<bf16>(M * 4 * D);
    L.h = cv.take<bf16>(M * 4 * D);
    L.mean1 = cv.take<float>(M);
    L.rstd1 = cv.take<float>(M);
    L.mean2 = cv.take<float>(M);
    L.rstd2 = cv.take<float>(M);
    L.lse = cv.take<float>(static_cast<size_t>(c.batch) * c.heads * c.seq_len);
  }
  if (c.patch > 0) {
    const size_t g2 = static_cast<size_t>(c.res / c.patch) * (c.res / c.patch);
    w.patches = cv.take<bf16>(static_cast<size_t>(c.batch) * g2 * 3 * c.patch * c.patch);
    w.patch_out = cv.take<bf16>(static_cast<size_t>(c.batch) * g2 * D);
    w.tokens_pre = cv.take<bf16>(M * D);
    w.mean_pre = cv.take<float>(M);
    w.rstd_pre = cv.take<float>(M);
  } else {
    w.patches = w.patch_out = w.tokens_pre = nullptr;
    w.mean_pre = w.rstd_pre = nullptr;
  }
  w.row_idx = cv.take<int>(c.batch);
  w.rows = cv.take<bf16>(static_cast<size_t>(c.batch) * D);
  w.rows_ln = cv.take<bf16>(static_cast<size_t>(c.batch) * D);
  w.mean_post = cv.take<float>(c.batch);
  w.rstd_post = cv.take<float>(c.batch);
  w.mean_f = cv.take<float>(M);
  w.rstd_f = cv.take<float>(M);
  w.dxa = cv.take<bf16>(M * D);
  w.dxb = cv.take<bf16>(M * D);
  w.dqkv = cv.take<bf16>(M * 3 * D);
  w.du = cv.take<bf16>(M * 4 * D);
  w.dtmp = cv.take<bf16>(M * D);
  w.dfeat = cv.take<bf16>(static_cast<size_t>(c.batch) * c.embed_dim);
  w.drow = cv.take<bf16>(static_cast<size_t>(c.batch) * D);
  w.drow2 = cv.take<bf16>(static_cast<size_t>(c.batch) * D);
  w.bytes = (cv.off + 255) & ~static_cast<size_t>(255);
}

static int check_cfg(const dc_tower_cfg* c) {
  if (c == nullptr) return set_error("tower: null cfg");
  if (c->layers < 0 || c->layers > MAX_LAYERS) return set_error("tower: layers out of range");
  if (c->width % 256 || c->width > 1024) return set_error("tower: width must be a multiple of 256 and <= 1024");
  if (c->heads * 64 != c->width) return set_error("tower: head_dim must be 64");
  if (c->seq_len < 1 || c->seq_len > 80) return set_error("tower: seq_len must be in [1,80]");
  if (c->batch < 1) return set_error("tower: batch must be >= 1");
  if (c->embed_dim % 8) return set_error("tower: embed_dim must be a multiple of 8");
  return 0;
}

__global__ void iota_stride_kernel(int* idx, int n, int stride) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) idx[i] = i * stride;
}

static dc_gemm_args gemm_args(const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, int M, int N, int K,
                              int epi, void* out, int ldo) {
  dc_gemm_args a;
  a.A = A; a.lda = lda; a.a_mn_major = a_mn;
  a.B = B; a.ldb = ldb; a.b_mn_major = b_mn;
  a.M = M; a.N = N; a.K = K;
  a.epilogue = epi; a.alpha = 1.0f;
  a.out = out; a.ldo = ldo;
  a.out2 = nullptr; a.ldo2 = 0;
  a.bias = nullptr; a.aux = nullptr; a.ldaux = 0;
  a.splits = 0; a.block_n = 0;
  a.alpha_dev = nullptr;
  a.colsum = nullptr;
  return a;
}

// y[M,N] = x[M,K] W[N,K]^T (+bias) with the given epilogue
static int linear_fwd(const bf16* x, const bf16* W, const float* bias, int M, int N, int K, int epi, bf16* out,
                      bf16* out2, const bf16* aux, cudaStream_t st) {
  dc_gemm_args a = gemm_args(x, K, 0, W, K, 0, M, N, K, epi, out, N);
  a.bias = bias;
  a.out2 = out2; a.ldo2 = N;
  a.aux = aux; a.ldaux = N;
  return gemm_bf16(a, st);
}
// dx[M,K] = dy[M,N] W[N,K]   (W read MN-major, no transposed copy)
static int linear_dgrad(const bf16* dy, const bf16* W, int M, int N, int K, int epi, bf16* dx, const bf16* aux,
                        cudaStream_t st) {
  dc_gemm_args a = gemm_args(dy, N, 0, W, K, 1, M, K, N, epi, dx, K);
  a.aux = aux; a.ldaux = K;
  return gemm_bf16(a, st);
}
// dW[N,K] += dy[M,N]^T x[M,K] ; db[N] += colsum(dy)
static int linear_wgrad(const bf16* dy, const bf16* x, int M, int N, int K, float* dW, float* db, cudaStream_t st) {
  dc_gemm_args a = gemm_args(dy, N, 1, x, K, 1, N, K, M, DC_EPI_F32_ATOMIC, dW, K);
  DC_TRY(gemm_bf16(a, st));
  if (db != nullptr) DC_TRY(dc_colsum_bf16(dy, N, db, M, N, st));
  return 0;
}

static int layers_forward(const dc_tower_cfg& c, TowerWs& w, const void* const* wb, const float* const* wf,
                          cudaStream_t st) {
  const int M = c.batch * c.seq_len, D = c.width;
  for (int l = 0; l < c.layers; ++l) {
    LayerWs& L = w.layer[l];
    const bf16* w_in = static_cast<const bf16*>(wb[4 * l + 0]);
    const bf16* w_out = static_cast<const bf16*>(wb[4 * l + 1]);
    const bf16* w_fc = static_cast<const bf16*>(wb[4 * l + 2]);
    const bf16* w_pr = static_cast<const bf16*>(wb[4 * l + 3]);
    const float* const* f = wf + 8 * l;
    // x = x + attn(ln_1(x))                                   base_transformer.py:51
    DC_TRY(dc_layernorm_fwd(w.xs[l], f[0], f[1], L.ln1, L.mean1, L.rstd1, M, D, 1e-5f, st));
    DC_TRY(linear_fwd(L.ln1, w_in, f[2], M, 3 * D, D, DC_EPI_BF16, L.qkv, nullptr, nullptr, st));
    DC_TRY(dc_attention_fwd(L.qkv, L.attn, L.lse, c.batch, c.seq_len, c.heads, c.causal, st));
    DC_TRY(linear_fwd(L.attn, w_out, f[3], M, D, D, DC_EPI_BF16_RESID, L.xmid, nullptr, w.xs[l], st));
    // x = x + c_proj(QuickGELU(c_fc(ln_2(x))))                 base_transformer.py:52
    DC_TRY(dc_layernorm_fwd(L.xmid, f[4], f[5], L.ln2, L.mean2, L.rstd2, M, D, 1e-5f, st));
    DC_TRY(linear_fwd(L.ln2, w_fc, f[6], M, 4 * D, D, DC_EPI_BF16_GELU, L.h, L.u, nullptr, st));
    DC_TRY(linear_fwd(L.h, w_pr, f[7], M, D, 4 * D, DC_EPI_BF16_RESID, w.xs[l + 1], nullptr, L.xmid, st));
  }
  return 0;
}

// dx_top: gradient w.r.t. xs[layers] (in w.dxa); on return the gradient w.r.t. xs[0] is in w.dxa.
// Bias gradients are fused into the kernel that PRODUCES the corresponding dY (no separate column-sum passes):
//   c_proj.bias  (colsum of d xs[l+1]) : LN1-backward of layer l+1 (`dcol`), or `top_colsum` rows for the last layer
//   c_fc.bias    (colsum of dU)        : DGELU epilogue of the c_proj dgrad GEMM
//   out_proj.bias(colsum of d xmid)    : LN2-backward of the same layer
//   in_proj_bias (colsum of dqkv)      : attention backward kernel
static int layers_backward(const dc_tower_cfg& c, TowerWs& w, const void* const* wb, const float* const* wf,
                           float* const* grads, cudaStream_t st) {
  const int M = c.batch * c.seq_len, D = c.width;
  for (int l = c.layers - 1; l >= 0; --l) {
    LayerWs& L = w.layer[l];
    const bf16* w_in = static_cast<const bf16*>(wb[4 * l + 0]);
    const bf16* w_out = static_cast<const bf16*>(wb[4 * l + 1]);
    const bf16* w_fc = static_cast<const bf16*>(wb[4 * l + 2]);
    const bf16* w_pr = static_cast<const bf16*>(wb[4 * l + 3]);
    const float* const* f = wf + 8 * l;
    float* const* g = grads + 12 * l;
    bf16* dx = w.dxa;    // grad wrt xs[l+1]
    bf16* dmid = w.dxb;  // grad wrt xmid
    // MLP branch
    {
      dc_gemm_args a = gemm_args(dx, D, 0, w_pr, 4 * D, 1, M, 4 * D, D, DC_EPI_BF16_DGELU, w.du, 4 * D);
      a.aux = L.u; a.ldaux = 4 * D;
      a.colsum = g[7];                                                                // d c_fc.bias
      DC_TRY(gemm_bf16(a, st));                                                       // dU = (dx Wproj) * gelu'(u)
    }
    DC_TRY(linear_wgrad(dx, L.h, M, D, 4 * D, g[8], nullptr, st));
    DC_TRY(linear_dgrad(w.du, w_fc, M, 4 * D, D, DC_EPI_BF16, w.dtmp, nullptr, st));  // dLN2out
    DC_TRY(linear_wgrad(w.du, L.ln2, M, 4 * D, D, g[6], nullptr, st));
    DC_TRY(dc_layernorm_bwd(w.dtmp, L.xmid, f[4], L.mean2, L.rstd2, dx, dmid, g[10], g[11], g[3], M, D, st));
    // attention branch
    // in_proj_bias gradient with the tcgen05 core: Q slice from the attention kernel, V slice = colsum(dAttnOut) fused
    // into the GEMM that produces dAttnOut, K slice identically zero
    const bool tc = attention_tc_enabled() && attention_tc_supported(c.batch, c.seq_len, c.heads);
    {
      dc_gemm_args a = gemm_args(dmid, D, 0, w_out, D, 1, M, D, D, DC_EPI_BF16, w.dtmp, D);   // dAttnOut
      if (tc) a.colsum = g[1] + 2 * D;
      DC_TRY(gemm_bf16(a, st));
    }
    DC_TRY(linear_wgrad(dmid, L.attn, M, D, D, g[2], nullptr, st));
    if (tc)
      DC_TRY(attention_tc_bwd(L.qkv, w.dtmp, L.lse, w.dqkv, g[1], /*dbias_v=*/0, c.batch, c.seq_len, c.heads, c.causal, st));
    else
      DC_TRY(dc_attention_bwd(L.qkv, L.attn, w.dtmp, L.lse, w.dqkv, g[1], c.batch, c.seq_len, c.heads, c.causal, st));
    DC_TRY(linear_dgrad(w.dqkv, w_in, M, 3 * D, D, DC_EPI_BF16, w.dtmp, nullptr, st));  // dLN1out
    DC_TRY(linear_wgrad(w.dqkv, L.ln1, M, 3 * D, D, g[0], nullptr, st));
    float* dcol_prev = (l > 0) ? grads[12 * (l - 1) + 9] : nullptr;                     // c_proj.bias of layer l-1
    DC_TRY(dc_layernorm_bwd(w.dtmp, w.xs[l], f[0], L.mean1, L.rstd1, dmid, dx, g[4], g[5], dcol_prev, M, D, st));
    // gradients of layers >= l are final in stream order (c_proj.bias of layer l-1 belongs to the NEXT slice)
    if (g_progress.cb != nullptr && l > 0 && l % g_progress.every == 0) g_progress.cb(g_progress.user, l);
  }
  return 0;
}

}  // namespace dc

using namespace dc;

extern "C" {

int dc_set_backward_progress_cb(dc_progress_cb cb, void* user, int every) {
  g_progress.cb = cb;
  g_progress.user = user;
  g_progress.every = every < 1 ? 1 : every;
  return 0;
}

size_t dc_tower_workspace_bytes(const dc_tower_cfg* cfg) {
  if (check_cfg(cfg) != 0) return 0;
  TowerWs w;
  carve(*cfg, nullptr, w);
  return w.bytes;
}

// tokens [batch, L, D] <-> patch tokens [batch, L-1, D] (drop / skip the class token row)
__global__ void __launch_bounds__(256) dense_tokens_kernel(const dc::bf16* __restrict__ src, dc::bf16* __restrict__ dst,
                                                           int batch, int L, int D, int to_dense) {
  const int vw = D / 8;
  const size_t total = static_cast<size_t>(batch) * (L - 1) * vw;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t v = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; v < total; v += stride) {
    const int c = static_cast<int>(v % vw) * 8;
    const size_t t = v / vw;
    const size_t b = t / (L - 1), p = t % (L - 1);
    const size_t tok = (b * L + 1 + p) * D + c, den = t * D + c;
    if (to_dense) *reinterpret_cast<uint4*>(dst + den) = *reinterpret_cast<const uint4*>(src + tok);
    else *reinterpret_cast<uint4*>(dst + tok) = *reinterpret_cast<const uint4*>(src + den);
  }
}

int dc_vit_forward(const dc_tower_cfg* cfg, const float* images, long long sample_stride,
                   const void* const* w_bf16, const float* const* w_f32, void* workspace, float* features,
                   void* dense_out, dc_stream_t stream) {
  DC_TRY(check_cfg(cfg));
  const dc_tower_cfg& c = *cfg;
  if (c.patch <= 0 || c.res % c.patch) return set_error("vit: bad res/patch");
  const int g2 = (c.res / c.patch) * (c.res / c.patch);
  if (g2 + 1 != c.seq_len) return set_error("vit: seq_len must be (res/patch)^2 + 1");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TowerWs w;
  carve(c, workspace, w);
  const int M = c.batch * c.seq_len, D = c.width, NL = c.layers;
  const int kdim = 3 * c.patch * c.patch;
  const float* const* xf = w_f32 + 8 * NL;
  // conv1 as a patch GEMM                                      visual_transformer.py:56-59
  DC_TRY(dc_patchify(images, sample_stride, w.patches, c.batch, c.res, c.patch, st));
  DC_TRY(linear_fwd(w.patches, static_cast<const bf16*>(w_bf16[4 * NL + 0]), nullptr, c.batch * g2, D, kdim, DC_EPI_BF16,
                    w.patch_out, nullptr, nullptr, st));
  // class token + positional embedding, ln_pre                 visual_transformer.py:60-63
  DC_TRY(dc_vit_assemble(w.patch_out, xf[0], xf[1], w.tokens_pre, c.batch, g2, D, st));
  DC_TRY(dc_layernorm_fwd(w.tokens_pre, xf[2], xf[3], w.xs[0], w.mean_pre, w.rstd_pre, M, D, 1e-5f, st));
  DC_TRY(layers_forward(c, w, w_bf16, w_f32, st));
  if (dense_out != nullptr) {                                   // dense_feat = x[:, 1:, :]   visual_transformer.py:68
    dense_tokens_kernel<<<sm_count() * 8, 256, 0, st>>>(w.xs[NL], static_cast<bf16*>(dense_out), c.batch, c.seq_len, D, 1);
    DC_CHECK_LAUNCH("dense_tokens");
  }
  // ln_post on the class token, projection                     visual_transformer.py:69-73
  iota_stride_kernel<<<(c.batch + 255) / 256, 256, 0, st>>>(w.row_idx, c.batch, c.seq_len);
  DC_CHECK_LAUNCH("iota_stride");
  DC_TRY(dc_gather_rows(w.xs[NL], w.row_idx, w.rows, c.batch, D, st));
  DC_TRY(dc_layernorm_fwd(w.rows, xf[4], xf[5], w.rows_ln, w.mean_post, w.rstd_post, c.batch, D, 1e-5f, st));
  dc_gemm_args a = gemm_args(w.rows_ln, D, 0, w_bf16[4 * NL + 1], c.embed_dim, 1, c.batch, c.embed_dim, D, DC_EPI_F32,
                             features, c.embed_dim);
  return gemm_bf16(a, st);
}

int dc_tower_pre_features(const dc_tower_cfg* cfg, const void* workspace, void* out_bf16, dc_stream_t stream) {
  DC_TRY(check_cfg(cfg));
  TowerWs w;
  carve(*cfg, const_cast<void*>(workspace), w);
  cudaError_t e = cudaMemcpyAsync(out_bf16, w.rows_ln, static_cast<size_t>(cfg->batch) * cfg->width * sizeof(bf16),
                                  cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return set_error_cuda("tower_pre_features", e);
  return 0;
}

int dc_vit_backward(const dc_tower_cfg* cfg, const float* dfeatures, const void* ddense, const void* const* w_bf16,
                    const float* const* w_f32, float* const* grads, void* workspace, dc_stream_t stream) {
  return dc_vit_backward_pre(cfg, dfeatures, nullptr, ddense, w_bf16, w_f32, grads, workspace, stream);
}

int dc_vit_backward_pre(const dc_tower_cfg* cfg, const float* dfeatures, const void* dpre, const void* ddense,
                        const void* const* w_bf16, const float* const* w_f32, float* const* grads, void* workspace,
                        dc_stream_t stream) {
  DC_TRY(check_cfg(cfg));
  const dc_tower_cfg& c = *cfg;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TowerWs w;
  carve(c, workspace, w);
  const int M = c.batch * c.seq_len, D = c.width, NL = c.layers, E = c.embed_dim;
  const float* const* xf = w_f32 + 8 * NL;
  float* const* xg = grads + 12 * NL;
  const bf16* proj = static_cast<const bf16*>(w_bf16[4 * NL + 1]);
  DC_TRY(dc_cast_f32_bf16(dfeatures, w.dfeat, static_cast<size_t>(c.batch) * E, st));
  // dproj[D,E] += rows_ln^T dfeat
  {
    dc_gemm_args a = gemm_args(w.rows_ln, D, 1, w.dfeat, E, 1, D, E, c.batch, DC_EPI_F32_ATOMIC, xg[6], E);
    DC_TRY(gemm_bf16(a, st));
  }
  // d rows_ln[b,D] = dfeat[b,E] proj[D,E]^T  (+ the gradient of the pre-projection feature when it was handed out)
  {
    dc_gemm_args a = gemm_args(w.dfeat, E, 0, proj, E, 0, c.batch, D, E, dpre != nullptr ? DC_EPI_BF16_RESID : DC_EPI_BF16,
                               w.drow, D);
    a.aux = dpre; a.ldaux = D;
    DC_TRY(gemm_bf16(a, st));
  }
  DC_TRY(dc_layernorm_bwd(w.drow, w.rows, xf[4], w.mean_post, w.rstd_post, nullptr, w.drow2, xg[4], xg[5],
                          NL > 0 ? grads[12 * (NL - 1) + 9] : nullptr, c.batch, D, st));  // + c_proj.bias of the last layer
  cudaError_t e = cudaMemsetAsync(w.dxa, 0, static_cast<size_t>(M) * D * sizeof(bf16), st);
  if (e != cudaSuccess) return set_error_cuda("memset dx", e);
  DC_TRY(dc_scatter_rows(w.drow2, w.row_idx, w.dxa, c.batch, D, st));
  if (ddense != nullptr) {   // gradient of the patch tokens (class-token rows already hold the ln_post path)
    dense_tokens_kernel<<<sm_count() * 8, 256, 0, st>>>(static_cast<const bf16*>(ddense), w.dxa, c.batch, c.seq_len, D, 0);
    DC_CHECK_LAUNCH("dense_tokens_bwd");
    // c_proj.bias of the last layer also sees these rows
    if (NL > 0) DC_TRY(dc_colsum_bf16(ddense, D, grads[12 * (NL - 1) + 9], c.batch * (c.seq_len - 1), D, st));
  }
  DC_TRY(layers_backward(c, w, w_bf16, w_f32, grads, st));
  // ln_pre backward, then positional / class embedding gradients
  DC_TRY(dc_layernorm_bwd(w.dxa, w.tokens_pre, xf[2], w.mean_pre, w.rstd_pre, nullptr, w.dxb, xg[2], xg[3], nullptr, M, D, st));
  DC_TRY(dc_colsum_bf16(w.dxb, c.seq_len * D, xg[1], c.batch, c.seq_len * D, st));
  DC_TRY(dc_colsum_bf16(w.dxb, c.seq_len * D, xg[0], c.batch, D, st));
  return 0;
}

int dc_text_forward(const dc_tower_cfg* cfg, const long long* ids, const void* const* w_bf16,
                    const float* const* w_f32, void* workspace, float* features, void* words_out, dc_stream_t stream) {
  DC_TRY(check_cfg(cfg));
  const dc_tower_cfg& c = *cfg;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TowerWs w;
  carve(c, workspace, w);
  const int M = c.batch * c.seq_len, D = c.width, NL = c.layers, E = c.embed_dim;
  const float* const* xf = w_f32 + 8 * NL;
  DC_TRY(dc_text_embed(ids, xf[0], xf[1], w.xs[0], c.batch, c.seq_len, D, st));        // text_transformer.py:188-190
  DC_TRY(layers_forward(c, w, w_bf16, w_f32, st));
  DC_TRY(dc_eot_index(ids, w.row_idx, c.batch, c.seq_len, st));
  if (words_out != nullptr) {
    // dense mode: ln_final on every token (words_feat), EOT rows gathered from it   text_transformer.py:194-203
    DC_TRY(dc_layernorm_fwd(w.xs[NL], xf[2], xf[3], words_out, w.mean_f, w.rstd_f, M, D, 1e-5f, st));
    DC_TRY(dc_gather_rows(words_out, w.row_idx, w.rows_ln, c.batch, D, st));
  } else {
    // ln_final is row-wise, so it commutes with the EOT row gather: normalise b rows instead of b*77
    DC_TRY(dc_gather_rows(w.xs[NL], w.row_idx, w.rows, c.batch, D, st));
    DC_TRY(dc_layernorm_fwd(w.rows, xf[2], xf[3], w.rows_ln, w.mean_post, w.rstd_post, c.batch, D, 1e-5f, st));
  }
  dc_gemm_args a = gemm_args(w.rows_ln, D, 0, w_bf16[4 * NL + 0], D, 0, c.batch, E, D, DC_EPI_F32, features, E);
  a.bias = xf[4];
  return gemm_bf16(a, st);
}

int dc_text_backward(const dc_tower_cfg* cfg, const long long* ids, const float* dfeatures, int dense,
                     const void* dwords, const void* const* w_bf16, const float* const* w_f32, float* const* grads,
                     void* workspace, dc_stream_t stream) {
  DC_TRY(check_cfg(cfg));
  const dc_tower_cfg& c = *cfg;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TowerWs w;
  carve(c, workspace, w);
  const int M = c.batch * c.seq_len, D = c.width, NL = c.layers, E = c.embed_dim;
  const float* const* xf = w_f32 + 8 * NL;
  float* const* xg = grads + 12 * NL;
  const bf16* wtp = static_cast<const bf16*>(w_bf16[4 * NL + 0]);
  float* dcol_last = NL > 0 ? grads[12 * (NL - 1) + 9] : nullptr;   // c_proj.bias of the last layer
  DC_TRY(dc_cast_f32_bf16(dfeatures, w.dfeat, static_cast<size_t>(c.batch) * E, st));
  // text_projection: dW[E,D] += dfeat^T rows_ln ; db += colsum(dfeat) ; d rows_ln = dfeat W
  DC_TRY(linear_wgrad(w.dfeat, w.rows_ln, c.batch, E, D, xg[4], xg[5], st));
  DC_TRY(linear_dgrad(w.dfeat, wtp, c.batch, E, D, DC_EPI_BF16, w.drow, nullptr, st));
  const size_t act_bytes = static_cast<size_t>(M) * D * sizeof(bf16);
  if (dense) {
    // d ln_final(all tokens) = dwords (+ EOT rows), then one LayerNorm backward over every token
    cudaError_t e = dwords ? cudaMemcpyAsync(w.dtmp, dwords, act_bytes, cudaMemcpyDeviceToDevice, st)
                           : cudaMemsetAsync(w.dtmp, 0, act_bytes, st);
    if (e != cudaSuccess) return set_error_cuda("dense dwords copy", e);
    DC_TRY(dc_add_rows(w.drow, w.row_idx, w.dtmp, c.batch, D, st));
    DC_TRY(dc_layernorm_bwd(w.dtmp, w.xs[NL], xf[2], w.mean_f, w.rstd_f, nullptr, w.dxa, xg[2], xg[3], dcol_last, M, D, st));
  } else {
    DC_TRY(dc_layernorm_bwd(w.drow, w.rows, xf[2], w.mean_post, w.rstd_post, nullptr, w.drow2, xg[2], xg[3], dcol_last,
                            c.batch, D, st));
    cudaError_t e = cudaMemsetAsync(w.dxa, 0, act_bytes, st);
    if (e != cudaSuccess) return set_error_cuda("memset dx", e);
    DC_TRY(dc_scatter_rows(w.drow2, w.row_idx, w.dxa, c.batch, D, st));
  }
  DC_TRY(layers_backward(c, w, w_bf16, w_f32, grads, st));
  return dc_text_embed_bwd(ids, w.dxa, xg[0], xg[1], dense ? nullptr : w.row_idx, c.batch, c.seq_len, D, st);
}

}  // extern "C"
