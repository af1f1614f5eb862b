// declip_b200 — small-head kernels of the DeCLIP objectives (all fp32 I/O, tiny [batch, <=1024] tensors):
// BatchNorm1d (+ReLU) fwd/bwd of the SimSiam projector/predictor MLPs (declip.py:33-130), row-wise cosine
// similarity fwd/bwd (SimsiamLoss, loss_functions/loss.py:52-84), row arg-max + fp32 row gather of the
// nearest-neighbour memory bank (utils/nnclr_modules/nn_memory_bank.py:42-65), bf16 row add (EOT-row gradient
// into the dense ln_final gradient).
#include "common.cuh"
#include "internal.h"

namespace dc {

// ------------------------------------------------------------------------------------------------ BatchNorm1d
// x [rows, C] fp32.  Block (32, 8): 32 consecutive columns (coalesced 128-byte rows) x 8 row lanes.
// training: batch statistics (biased variance for normalisation; running stats get the unbiased one, as
// nn.BatchNorm1d does); eval: running statistics.
__global__ void __launch_bounds__(256) bn_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, float* __restrict__ y,
                                                     float* __restrict__ save_mean, float* __restrict__ save_rstd,
                                                     float* __restrict__ run_mean, float* __restrict__ run_var, int rows,
                                                     int C, float eps, float momentum, int training, int relu) {
  __shared__ float s1[8][33], s2[8][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int c = blockIdx.x * 32 + tx;
  float mean = 0.f, rstd = 0.f;
  if (training) {
    float a = 0.f, b = 0.f;
    if (c < C)
      for (int r = ty; r < rows; r += 8) {
        const float v = x[static_cast<size_t>(r) * C + c];
        a += v;
        b += v * v;
      }
    s1[ty][tx] = a;
    s2[ty][tx] = b;
    __syncthreads();
    a = 0.f; b = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { a += s1[i][tx]; b += s2[i][tx]; }
    mean = a / rows;
    const float var = fmaxf(b / rows - mean * mean, 0.f);
    rstd = rsqrtf(var + eps);
    if (ty == 0 && c < C) {
      save_mean[c] = mean;
      save_rstd[c] = rstd;
      if (run_mean != nullptr) {
        const float unbiased = rows > 1 ? var * rows / (rows - 1) : var;
        run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * mean;
        run_var[c] = (1.f - momentum) * run_var[c] + momentum * unbiased;
      }
    }
  } else if (c < C) {
    mean = run_mean[c];
    rstd = rsqrtf(run_var[c] + eps);
    if (ty == 0) { save_mean[c] = mean; save_rstd[c] = rstd; }
  }
  if (c >= C) return;
  const float g = gamma ? gamma[c] : 1.f, bb = beta ? beta[c] : 0.f;
  for (int r = ty; r < rows; r += 8) {
    float v = (x[static_cast<size_t>(r) * C + c] - mean) * rstd * g + bb;
    if (relu) v = fmaxf(v, 0.f);
    y[static_cast<size_t>(r) * C + c] = v;
  }
}

// dy is the gradient w.r.t. the (post-ReLU) output y; the ReLU mask is y > 0.
__global__ void __launch_bounds__(256) bn_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                     const float* __restrict__ y, const float* __restrict__ gamma,
                                                     const float* __restrict__ save_mean,
                                                     const float* __restrict__ save_rstd, float* __restrict__ dx,
                                                     float* __restrict__ dgamma, float* __restrict__ dbeta, int rows, int C,
                                                     int training, int relu) {
  __shared__ float s1[8][33], s2[8][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int c = blockIdx.x * 32 + tx;
  const float mean = c < C ? save_mean[c] : 0.f, rstd = c < C ? save_rstd[c] : 0.f;
  float a = 0.f, b = 0.f;
  if (c < C)
    for (int r = ty; r < rows; r += 8) {
      const size_t i = static_cast<size_t>(r) * C + c;
      float g = dy[i];
      if (relu && y[i] <= 0.f) g = 0.f;
      a += g;
      b += g * (x[i] - mean) * rstd;
    }
  s1[ty][tx] = a;
  s2[ty][tx] = b;
  __syncthreads();
  a = 0.f; b = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { a += s1[i][tx]; b += s2[i][tx]; }
  if (c >= C) return;
  if (ty == 0) {
    if (dbeta) atomicAdd(&dbeta[c], a);
    if (dgamma) atomicAdd(&dgamma[c], b);
  }
  const float g0 = gamma ? gamma[c] : 1.f;
  const float m1 = a / rows, m2 = b / rows;
  for (int r = ty; r < rows; r += 8) {
    const size_t i = static_cast<size_t>(r) * C + c;
    float g = dy[i];
    if (relu && y[i] <= 0.f) g = 0.f;
    const float xh = (x[i] - mean) * rstd;
    dx[i] = training ? g0 * rstd * (g - m1 - xh * m2) : g0 * rstd * g;
  }
}

// ------------------------------------------------------------------------------------------------ cosine rows
// cos[r] = <p_r, z_r> / (max(||p_r||, eps) max(||z_r||, eps)); one warp per row.
__global__ void __launch_bounds__(256) cosine_fwd_kernel(const float* __restrict__ p, const float* __restrict__ z,
                                                         float* __restrict__ cosv, int n, int dim) {
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  const float* pr = p + static_cast<size_t>(row) * dim;
  const float* zr = z + static_cast<size_t>(row) * dim;
  float pp = 0.f, zz = 0.f, pz = 0.f;
  for (int c = lane; c < dim; c += 32) {
    const float a = pr[c], b = zr[c];
    pp += a * a; zz += b * b; pz += a * b;
  }
  pp = warp_sum(pp); zz = warp_sum(zz); pz = warp_sum(pz);
  if (lane == 0) cosv[row] = pz / (fmaxf(sqrtf(pp), 1e-30f) * fmaxf(sqrtf(zz), 1e-30f));
}
// dp[r] = g[r] * (z_hat - cos * p_hat) / ||p||   (z is a constant: stop-gradient, loss.py:54)
__global__ void __launch_bounds__(256) cosine_bwd_kernel(const float* __restrict__ p, const float* __restrict__ z,
                                                         const float* __restrict__ grow, float gscale,
                                                         float* __restrict__ dp, int n, int dim) {
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  const float* pr = p + static_cast<size_t>(row) * dim;
  const float* zr = z + static_cast<size_t>(row) * dim;
  float pp = 0.f, zz = 0.f, pz = 0.f;
  for (int c = lane; c < dim; c += 32) {
    const float a = pr[c], b = zr[c];
    pp += a * a; zz += b * b; pz += a * b;
  }
  pp = warp_sum(pp); zz = warp_sum(zz); pz = warp_sum(pz);
  const float np = fmaxf(sqrtf(pp), 1e-30f), nz = fmaxf(sqrtf(zz), 1e-30f);
  const float cs = pz / (np * nz);
  const float g = gscale * (grow ? grow[row] : 1.f);
  float* o = dp + static_cast<size_t>(row) * dim;
  for (int c = lane; c < dim; c += 32) o[c] = g * (zr[c] / nz - cs * pr[c] / np) / np;
}

// ------------------------------------------------------------------------------------------------ NN bank
// idx[r] = argmax_c x[r, c]  (first maximum, like torch.topk k=1); one block per row.
__global__ void __launch_bounds__(256) argmax_rows_kernel(const float* __restrict__ x, int ld, int cols,
                                                          int* __restrict__ idx) {
  __shared__ float sv[8];
  __shared__ int si[8];
  const float* xr = x + static_cast<size_t>(blockIdx.x) * ld;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    const float v = xr[c];
    if (v > best) { best = v; bi = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sv[warp] = best; si[warp] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i)
      if (sv[i] > best || (sv[i] == best && si[i] < bi)) { best = sv[i]; bi = si[i]; }
    idx[blockIdx.x] = bi;
  }
}

__global__ void __launch_bounds__(256) gather_rows_f32_kernel(const float* __restrict__ src, const int* __restrict__ idx,
                                                              float* __restrict__ dst, int n, int width) {
  const int vw = width / 4;
  const size_t total = static_cast<size_t>(n) * vw;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t v = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; v < total; v += stride) {
    const int c = static_cast<int>(v % vw) * 4;
    const size_t i = v / vw;
    *reinterpret_cast<float4*>(dst + i * width + c) =
        *reinterpret_cast<const float4*>(src + static_cast<size_t>(idx[i]) * width + c);
  }
}

// dst[idx[i], :] += src[i, :]   (bf16; idx rows distinct)
__global__ void __launch_bounds__(256) add_rows_kernel(const bf16* __restrict__ src, const int* __restrict__ idx,
                                                       bf16* __restrict__ dst, int n, int width) {
  const int vw = width / 8;
  const size_t total = static_cast<size_t>(n) * vw;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t v = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; v < total; v += stride) {
    const int c = static_cast<int>(v % vw) * 8;
    const size_t i = v / vw;
    bf16* d = dst + static_cast<size_t>(idx[i]) * width + c;
    const uint4 a = *reinterpret_cast<const uint4*>(src + i * width + c);
    const uint4 b = *reinterpret_cast<const uint4*>(d);
    float2 fa, fb;
    uint4 w;
    fa = unpack_bf16x2(a.x); fb = unpack_bf16x2(b.x); w.x = pack_bf16x2(fa.x + fb.x, fa.y + fb.y);
    fa = unpack_bf16x2(a.y); fb = unpack_bf16x2(b.y); w.y = pack_bf16x2(fa.x + fb.x, fa.y + fb.y);
    fa = unpack_bf16x2(a.z); fb = unpack_bf16x2(b.z); w.z = pack_bf16x2(fa.x + fb.x, fa.y + fb.y);
    fa = unpack_bf16x2(a.w); fb = unpack_bf16x2(b.w); w.w = pack_bf16x2(fa.x + fb.x, fa.y + fb.y);
    *reinterpret_cast<uint4*>(d) = w;
  }
}

static inline int grid_cap(size_t items, int threads) {
  size_t blocks = (items + threads - 1) / threads;
  const size_t cap = static_cast<size_t>(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace dc

using namespace dc;

extern "C" {

int dc_batchnorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* save_mean,
                     float* save_rstd, float* running_mean, float* running_var, int rows, int channels, float eps,
                     float momentum, int training, int relu, dc_stream_t stream) {
  if (rows <= 0 || channels <= 0) return 0;
  if (!training && (running_mean == nullptr || running_var == nullptr)) return set_error("batchnorm: eval needs running stats");
  bn_fwd_kernel<<<(channels + 31) / 32, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(
      x, gamma, beta, y, save_mean, save_rstd, running_mean, running_var, rows, channels, eps, momentum, training, relu);
  DC_CHECK_LAUNCH("batchnorm_fwd");
  return 0;
}

int dc_batchnorm_bwd(const float* dy, const float* x, const float* y, const float* gamma, const float* save_mean,
                     const float* save_rstd, float* dx, float* dgamma, float* dbeta, int rows, int channels,
                     int training, int relu, dc_stream_t stream) {
  if (rows <= 0 || channels <= 0) return 0;
  bn_bwd_kernel<<<(channels + 31) / 32, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(
      dy, x, y, gamma, save_mean, save_rstd, dx, dgamma, dbeta, rows, channels, training, relu);
  DC_CHECK_LAUNCH("batchnorm_bwd");
  return 0;
}

int dc_cosine_rows_fwd(const float* p, const float* z, float* cosv, int n, int dim, dc_stream_t stream) {
  if (n <= 0) return 0;
  cosine_fwd_kernel<<<(n * 32 + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(p, z, cosv, n, dim);
  DC_CHECK_LAUNCH("cosine_fwd");
  return 0;
}

int dc_cosine_rows_bwd(const float* p, const float* z, const float* grow, float gscale, float* dp, int n, int dim,
                       dc_stream_t stream) {
  if (n <= 0) return 0;
  cosine_bwd_kernel<<<(n * 32 + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(p, z, grow, gscale, dp, n, dim);
  DC_CHECK_LAUNCH("cosine_bwd");
  return 0;
}

int dc_argmax_rows(const float* x, int ld, int rows, int cols, int* idx, dc_stream_t stream) {
  if (rows <= 0) return 0;
  argmax_rows_kernel<<<rows, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, ld, cols, idx);
  DC_CHECK_LAUNCH("argmax_rows");
  return 0;
}

int dc_gather_rows_f32(const float* src, const int* idx, float* dst, int n, int width, dc_stream_t stream) {
  if (n <= 0) return 0;
  if (width & 3) return set_error("gather_rows_f32: width must be a multiple of 4");
  gather_rows_f32_kernel<<<grid_cap(static_cast<size_t>(n) * (width / 4), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, idx, dst, n, width);
  DC_CHECK_LAUNCH("gather_rows_f32");
  return 0;
}

int dc_add_rows(const void* src, const int* idx, void* dst, int n, int width, dc_stream_t stream) {
  if (n <= 0) return 0;
  if (width & 7) return set_error("add_rows: width must be a multiple of 8");
  add_rows_kernel<<<grid_cap(static_cast<size_t>(n) * (width / 8), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(src), idx, static_cast<bf16*>(dst), n, width);
  DC_CHECK_LAUNCH("add_rows");
  return 0;
}

}  // extern "C"

// ================================================================================================ FILIP
// Token-wise late interaction — prototype/model/filip.py:71-106.
namespace dc {

// score1[b,j] = <d1[b,j,:], sum_m d2[b,m,:]>, score2[b,m] = <d2[b,m,:], sum_j d1[b,j,:]>   (filip.py:79-81: the
// row / column sums of the per-pair cross-logit matrix, without forming it).  One block per sample.
__global__ void __launch_bounds__(256) token_scores_kernel(const float* __restrict__ d1, const float* __restrict__ d2,
                                                           int n1, int n2, int dim, float* __restrict__ score1,
                                                           float* __restrict__ score2) {
  extern __shared__ float s_sum[];  // [2][dim]
  const int b = blockIdx.x;
  const float* a = d1 + static_cast<size_t>(b) * n1 * dim;
  const float* c = d2 + static_cast<size_t>(b) * n2 * dim;
  for (int k = threadIdx.x; k < dim; k += blockDim.x) {
    float s1 = 0.f, s2 = 0.f;
    for (int j = 0; j < n1; ++j) s1 += a[static_cast<size_t>(j) * dim + k];
    for (int m = 0; m < n2; ++m) s2 += c[static_cast<size_t>(m) * dim + k];
    s_sum[k] = s1;
    s_sum[dim + k] = s2;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int t = warp; t < n1 + n2; t += nw) {
    const bool first = t < n1;
    const float* row = first ? a + static_cast<size_t>(t) * dim : c + static_cast<size_t>(t - n1) * dim;
    const float* other = first ? s_sum + dim : s_sum;
    float acc = 0.f;
    for (int k = lane; k < dim; k += 32) acc += row[k] * other[k];
    acc = warp_sum(acc);
    if (lane == 0) {
      if (first) score1[static_cast<size_t>(b) * n1 + t] = acc;
      else score2[static_cast<size_t>(b) * n2 + (t - n1)] = acc;
    }
  }
}

// out[i, l] = mean_{j < n} max_{m < group} G[i*n + j, l*group + m]   (filip.py:103-104), arg[i*n+j, l] = argmax m.
// One block per (sample i, 32 candidate samples l); thread (tx = l lane, ty = token lane).
__global__ void __launch_bounds__(256) groupmax_mean_fwd_kernel(const float* __restrict__ G, int ldg, int n, int group,
                                                                int ncand, float* __restrict__ out, int ldo,
                                                                uint8_t* __restrict__ arg) {
  __shared__ float s_part[8][33];
  const int i = blockIdx.y;
  const int l = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (l < ncand) {
    for (int j = threadIdx.y; j < n; j += 8) {
      const float* g = G + (static_cast<size_t>(i) * n + j) * ldg + static_cast<size_t>(l) * group;
      float best = g[0];
      int bi = 0;
      for (int m = 1; m < group; ++m) {
        const float v = g[m];
        if (v > best) { best = v; bi = m; }
      }
      acc += best;
      if (arg != nullptr) arg[(static_cast<size_t>(i) * n + j) * ncand + l] = static_cast<uint8_t>(bi);
    }
  }
  s_part[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && l < ncand) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) s += s_part[r][threadIdx.x];
    out[static_cast<size_t>(i) * ldo + l] = s / n;
  }
}

// dG[i*n + j, l*group + m] = (m == arg) ? dout[i, l] / n : 0   (bf16, feeds the two backward GEMMs)
__global__ void __launch_bounds__(256) groupmax_mean_bwd_kernel(const float* __restrict__ dout, int ldd,
                                                                const uint8_t* __restrict__ arg, int lda, int n,
                                                                int group, int ncand, bf16* __restrict__ dG, int ldg,
                                                                int rows) {
  const size_t total = static_cast<size_t>(rows) * ncand;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  const float invn = 1.0f / n;
  for (size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total; t += stride) {
    const size_t r = t / ncand;
    const int l = static_cast<int>(t - r * ncand);
    const int i = static_cast<int>(r / n);
    const float g = dout[static_cast<size_t>(i) * ldd + l] * invn;
    const int a = arg[r * lda + l];
    bf16* d = dG + r * ldg + static_cast<size_t>(l) * group;
    for (int m = 0; m < group; m += 2)
      *reinterpret_cast<uint32_t*>(d + m) = pack_bf16x2(m == a ? g : 0.f, m + 1 == a ? g : 0.f);
  }
}

// dst[idx[i], :] += src[i, :] (fp32 rows; idx rows distinct)
__global__ void __launch_bounds__(256) add_rows_f32_kernel(const float* __restrict__ src, const int* __restrict__ idx,
                                                           float* __restrict__ dst, int n, int width) {
  const int vw = width / 4;
  const size_t total = static_cast<size_t>(n) * vw;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t v = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; v < total; v += stride) {
    const int c = static_cast<int>(v % vw) * 4;
    const size_t i = v / vw;
    float4* d = reinterpret_cast<float4*>(dst + static_cast<size_t>(idx[i]) * width + c);
    const float4 a = *reinterpret_cast<const float4*>(src + i * width + c);
    float4 b = *d;
    b.x += a.x; b.y += a.y; b.z += a.z; b.w += a.w;
    *d = b;
  }
}

}  // namespace dc

extern "C" {

int dc_token_scores(const float* d1, const float* d2, int batch, int n1, int n2, int dim, float* score1, float* score2,
                    dc_stream_t stream) {
  if (batch <= 0) return 0;
  token_scores_kernel<<<batch, 256, 2 * dim * sizeof(float), static_cast<cudaStream_t>(stream)>>>(d1, d2, n1, n2, dim,
                                                                                               score1, score2);
  DC_CHECK_LAUNCH("token_scores");
  return 0;
}

int dc_groupmax_mean_fwd(const float* G, int ldg, int batch, int n, int group, int ncand, float* out, int ldo,
                         unsigned char* arg, dc_stream_t stream) {
  if (batch <= 0 || ncand <= 0) return 0;
  if (group > 255) return dc::set_error("groupmax: group must be <= 255");
  dim3 grid((ncand + 31) / 32, batch);
  groupmax_mean_fwd_kernel<<<grid, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(G, ldg, n, group, ncand, out, ldo,
                                                                                     arg);
  DC_CHECK_LAUNCH("groupmax_mean_fwd");
  return 0;
}

int dc_groupmax_mean_bwd(const float* dout, int ldd, const unsigned char* arg, int batch, int n, int group, int ncand,
                         void* dG, int ldg, dc_stream_t stream) {
  return dc_groupmax_mean_bwd_ex(dout, ldd, arg, ncand, batch, n, group, ncand, dG, ldg, stream);
}

int dc_groupmax_mean_bwd_ex(const float* dout, int ldd, const unsigned char* arg, int lda, int batch, int n, int group,
                            int ncand, void* dG, int ldg, dc_stream_t stream) {
  if (batch <= 0 || ncand <= 0) return 0;
  if (group & 1) return dc::set_error("groupmax: group must be even");
  const int rows = batch * n;
  groupmax_mean_bwd_kernel<<<dc::grid_cap(static_cast<size_t>(rows) * ncand, 256), 256, 0,
                             static_cast<cudaStream_t>(stream)>>>(dout, ldd, arg, lda, n, group, ncand,
                                                                  static_cast<dc::bf16*>(dG), ldg, rows);
  DC_CHECK_LAUNCH("groupmax_mean_bwd");
  return 0;
}

int dc_add_rows_f32(const float* src, const int* idx, float* dst, int n, int width, dc_stream_t stream) {
  if (n <= 0) return 0;
  if (width & 3) return dc::set_error("add_rows_f32: width must be a multiple of 4");
  add_rows_f32_kernel<<<dc::grid_cap(static_cast<size_t>(n) * (width / 4), 256), 256, 0,
                        static_cast<cudaStream_t>(stream)>>>(src, idx, dst, n, width);
  DC_CHECK_LAUNCH("add_rows_f32");
  return 0;
}

}  // extern "C"
