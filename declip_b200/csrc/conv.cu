// declip_b200 — ModifiedResNet support kernels (prototype/model/image_encoder/modified_resnet.py).
//
// Activations are NHWC bf16, i.e. a [rows = B*H*W, C] matrix, so every 1x1 convolution IS the tcgen05 GEMM of
// gemm.cu and every 3x3 convolution is im2col (below) + the same GEMM with K = 9*C ordered (ky, kx, c) — weights are
// permuted once per step into that order on the host side.  BatchNorm2d in training mode is per-channel statistics
// over the rows (column statistics), ReLU / the residual add are fused into its apply pass, AvgPool2d(2) and the
// AttentionPool2d token assembly are small vector kernels.  All HBM-bound: 16-byte vector accesses, grid-stride.
#include "common.cuh"
#include "internal.h"

namespace dc {

static inline int grid_for_items(size_t items, int threads, int per_sm = 8) {
  size_t blocks = (items + threads - 1) / threads;
  const size_t cap = static_cast<size_t>(sm_count()) * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

// ------------------------------------------------------------------------------------------------ stem conv1
// modified_resnet.py:150 conv1 = Conv2d(3, 32, k=3, s=2, p=1): im2col straight from the fp32 NCHW image.
// col[(b, oy, ox), k] with k = (ky*3 + kx)*3 + c for k < 27, zero for 27 <= k < 32 (K padded for 16-byte rows).
__global__ void __launch_bounds__(256) im2col_stem_kernel(const float* __restrict__ img, long long sample_stride,
                                                          bf16* __restrict__ col, int batch, int H, int W) {
  const int OH = H / 2, OW = W / 2;
  const size_t rows = static_cast<size_t>(batch) * OH * OW;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t r = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; r < rows; r += stride) {
    const int ox = static_cast<int>(r % OW);
    const int oy = static_cast<int>((r / OW) % OH);
    const size_t b = r / (static_cast<size_t>(OW) * OH);
    const float* base = img + b * sample_stride;
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = oy * 2 - 1 + ky;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = ox * 2 - 1 + kx;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
#pragma unroll
          for (int c = 0; c < 3; ++c) v[(ky * 3 + kx) * 3 + c] = __ldg(base + (static_cast<size_t>(c) * H + iy) * W + ix);
        }
      }
    }
    uint4* dst = reinterpret_cast<uint4*>(col + r * 32);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 w;
      w.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]); w.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
      w.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]); w.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
      dst[j] = w;
    }
  }
}

// ------------------------------------------------------------------------------------------------ 3x3, pad 1, stride 1
// col[(b,y,x), (ky*3+kx)*C + c] = in[(b, y-1+ky, x-1+kx), c]  (zero outside).  One block walks image rows (b, y); a
// thread owns ONE (tap, 8-channel vector) slot of the 9 * C / 8 slots of a pixel and steps over x, so the inner loop has no
// integer division (the first version did five 64-bit divisions per 16 bytes and ran at 2 TB/s) and a warp writes one
// contiguous run of the column matrix.  Launch with blockDim.x a multiple of 9 * C / 8 (or 256 when that exceeds 256).
__global__ void __launch_bounds__(256) im2col3x3_kernel(const bf16* __restrict__ in, bf16* __restrict__ col, int batch,
                                                        int H, int W, int C) {
  const int vc = C / 8, R = 9 * vc;
  const int nrows = batch * H;
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  if (blockDim.x % R == 0) {
    const int px = blockDim.x / R;                      // pixels per pass
    const int x0 = threadIdx.x / R, r = threadIdx.x - x0 * R;
    const int tap = r / vc, c = (r - tap * vc) * 8;
    const int ky = tap / 3, kx = tap - ky * 3;
    for (int row = blockIdx.x; row < nrows; row += gridDim.x) {
      const int b = row / H, y = row - b * H;
      const int iy = y - 1 + ky;
      const bool yok = iy >= 0 && iy < H;
      const bf16* src = in + (static_cast<size_t>(b) * H + (yok ? iy : 0)) * W * C + c;
      bf16* dst = col + static_cast<size_t>(row) * W * (9 * static_cast<size_t>(C)) + static_cast<size_t>(r) * 8;
#pragma unroll 4
      for (int x = x0; x < W; x += px) {
        const int ix = x - 1 + kx;
        uint4 v = zero;
        if (yok && ix >= 0 && ix < W) v = *reinterpret_cast<const uint4*>(src + static_cast<size_t>(ix) * C);
        *reinterpret_cast<uint4*>(dst + static_cast<size_t>(x) * (9 * static_cast<size_t>(C))) = v;
      }
    }
    return;
  }
  for (int row = blockIdx.x; row < nrows; row += gridDim.x) {
    const int b = row / H, y = row - b * H;
    bf16* dst = col + static_cast<size_t>(row) * W * (9 * static_cast<size_t>(C));
    for (int idx = threadIdx.x; idx < W * R; idx += blockDim.x) {
      const int x = idx / R, r = idx - x * R;
      const int tap = r / vc, c = (r - tap * vc) * 8;
      const int iy = y - 1 + tap / 3, ix = x - 1 + tap % 3;
      uint4 v = zero;
      if (iy >= 0 && iy < H && ix >= 0 && ix < W)
        v = *reinterpret_cast<const uint4*>(in + ((static_cast<size_t>(b) * H + iy) * W + ix) * C + c);
      *reinterpret_cast<uint4*>(dst + static_cast<size_t>(idx) * 8) = v;
    }
  }
}
// dgrad of the above: din[(b,y,x), c] = sum over taps of dcol[(b, y+1-ky, x+1-kx), tap*C + c]  (a gather: no atomics)
__global__ void __launch_bounds__(256) col2im3x3_kernel(const bf16* __restrict__ dcol, bf16* __restrict__ din, int batch,
                                                        int H, int W, int C) {
  const int vc = C / 8;
  const int nrows = batch * H;
  for (int row = blockIdx.x; row < nrows; row += gridDim.x)
  for (int idx = threadIdx.x; idx < W * vc; idx += blockDim.x) {
    const int x = idx / vc, c = (idx - x * vc) * 8;
    const size_t b = static_cast<size_t>(row / H);
    const int y = row - static_cast<int>(b) * H;
    const size_t q = static_cast<size_t>(row) * W + x;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int oy = y + 1 - tap / 3, ox = x + 1 - tap % 3;
      if (oy >= 0 && oy < H && ox >= 0 && ox < W) {
        const uint4 u = *reinterpret_cast<const uint4*>(dcol + ((b * H + oy) * W + ox) * (9 * static_cast<size_t>(C)) +
                                                        static_cast<size_t>(tap) * C + c);
        float2 f;
        f = unpack_bf16x2(u.x); acc[0] += f.x; acc[1] += f.y;
        f = unpack_bf16x2(u.y); acc[2] += f.x; acc[3] += f.y;
        f = unpack_bf16x2(u.z); acc[4] += f.x; acc[5] += f.y;
        f = unpack_bf16x2(u.w); acc[6] += f.x; acc[7] += f.y;
      }
    }
    uint4 w;
    w.x = pack_bf16x2(acc[0], acc[1]); w.y = pack_bf16x2(acc[2], acc[3]);
    w.z = pack_bf16x2(acc[4], acc[5]); w.w = pack_bf16x2(acc[6], acc[7]);
    *reinterpret_cast<uint4*>(din + q * C + c) = w;
  }
}

// ------------------------------------------------------------------------------------------------ AvgPool2d(2)
__global__ void __launch_bounds__(256) avgpool2_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int batch,
                                                       int H, int W, int C, int backward) {
  const int OH = H / 2, OW = W / 2, vc = C / 8;
  const int nrows = batch * OH;                       // block-per-output-row walk: 32-bit index math only
  for (int row = blockIdx.x; row < nrows; row += gridDim.x)
  for (int idx = threadIdx.x; idx < OW * vc; idx += blockDim.x) {
    const int ox = idx / vc, c = (idx - ox * vc) * 8;
    const size_t b = static_cast<size_t>(row / OH);
    const int oy = row - static_cast<int>(b) * OH;
    const size_t q = static_cast<size_t>(row) * OW + ox;
    const size_t i00 = ((b * H + 2 * oy) * W + 2 * ox) * C + c;
    const size_t offs[4] = {i00, i00 + C, i00 + static_cast<size_t>(W) * C, i00 + static_cast<size_t>(W) * C + C};
    if (!backward) {
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint4 u = *reinterpret_cast<const uint4*>(in + offs[k]);
        float2 f;
        f = unpack_bf16x2(u.x); acc[0] += f.x; acc[1] += f.y;
        f = unpack_bf16x2(u.y); acc[2] += f.x; acc[3] += f.y;
        f = unpack_bf16x2(u.z); acc[4] += f.x; acc[5] += f.y;
        f = unpack_bf16x2(u.w); acc[6] += f.x; acc[7] += f.y;
      }
      uint4 w;
      w.x = pack_bf16x2(0.25f * acc[0], 0.25f * acc[1]); w.y = pack_bf16x2(0.25f * acc[2], 0.25f * acc[3]);
      w.z = pack_bf16x2(0.25f * acc[4], 0.25f * acc[5]); w.w = pack_bf16x2(0.25f * acc[6], 0.25f * acc[7]);
      *reinterpret_cast<uint4*>(out + q * C + c) = w;
    } else {  // in = d(pooled) [OH,OW], out = d(input) [H,W]
      const uint4 u = *reinterpret_cast<const uint4*>(in + q * C + c);
      float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
      uint4 w;
      w.x = pack_bf16x2(0.25f * f0.x, 0.25f * f0.y); w.y = pack_bf16x2(0.25f * f1.x, 0.25f * f1.y);
      w.z = pack_bf16x2(0.25f * f2.x, 0.25f * f2.y); w.w = pack_bf16x2(0.25f * f3.x, 0.25f * f3.y);
#pragma unroll
      for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4*>(out + offs[k]) = w;
    }
  }
}

// ------------------------------------------------------------------------------------------------ BatchNorm2d (NHWC)
// Column statistics of a bf16 [rows, C] matrix: sums[c] += sum_r f(x), sums[C + c] += sum_r g(x).
//   mode 0 (forward):  f = x,        g = x^2
//   mode 1 (backward): f = dy*mask,  g = dy*mask*xhat      (mask = y > 0 when relu; xhat = (x-mean)*rstd)
constexpr int BN_ROWS = 512;
__global__ void __launch_bounds__(256) bn2d_stats_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy,
                                                         const bf16* __restrict__ y, const float* __restrict__ mean,
                                                         const float* __restrict__ rstd, float* __restrict__ sums,
                                                         size_t rows, int C, int mode, int relu) {
  __shared__ float s1[32][65], s2[32][65];
  const int cv = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int c0 = blockIdx.x * 64 + cv * 8;
  const size_t r0 = static_cast<size_t>(blockIdx.y) * BN_ROWS;
  const size_t r1 = r0 + BN_ROWS < rows ? r0 + BN_ROWS : rows;
  float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, b[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (c0 < C) {
    float mu[8], rs[8];
    if (mode == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { mu[i] = mean[c0 + i]; rs[i] = rstd[c0 + i]; }
    }
    for (size_t r = r0 + rl; r < r1; r += 32) {
      float xv[8];
      unpack8(*reinterpret_cast<const uint4*>(x + r * C + c0), xv);
      if (mode == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { a[i] += xv[i]; b[i] += xv[i] * xv[i]; }
      } else {
        float gv[8], yv[8];
        unpack8(*reinterpret_cast<const uint4*>(dy + r * C + c0), gv);
        if (relu) unpack8(*reinterpret_cast<const uint4*>(y + r * C + c0), yv);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float g = (relu && yv[i] <= 0.f) ? 0.f : gv[i];
          a[i] += g;
          b[i] += g * (xv[i] - mu[i]) * rs[i];
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) { s1[rl][cv * 8 + i] = a[i]; s2[rl][cv * 8 + i] = b[i]; }
  __syncthreads();
  if (threadIdx.x < 64) {
    float t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int r = 0; r < 32; ++r) { t1 += s1[r][threadIdx.x]; t2 += s2[r][threadIdx.x]; }
    const int c = blockIdx.x * 64 + threadIdx.x;
    if (c < C) { atomicAdd(&sums[c], t1); atomicAdd(&sums[C + c], t2); }
  }
}

// sums -> mean / rstd (+ running statistics with the unbiased variance, nn.BatchNorm2d semantics)
__global__ void bn2d_finalize_kernel(const float* __restrict__ sums, float* __restrict__ mean, float* __restrict__ rstd,
                                     float* __restrict__ run_mean, float* __restrict__ run_var, double rows, int C,
                                     float eps, float momentum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double m = sums[c] / rows;
  double var = sums[C + c] / rows - m * m;
  if (var < 0) var = 0;
  mean[c] = static_cast<float>(m);
  rstd[c] = static_cast<float>(1.0 / sqrt(var + eps));
  if (run_mean != nullptr) {
    const double unbiased = rows > 1 ? var * rows / (rows - 1) : var;
    run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * static_cast<float>(m);
    run_var[c] = (1.f - momentum) * run_var[c] + momentum * static_cast<float>(unbiased);
  }
}

// eval mode: statistics come from the running buffers (nn.BatchNorm2d.eval())
__global__ void bn2d_eval_stats_kernel(const float* __restrict__ run_mean, const float* __restrict__ run_var,
                                       float* __restrict__ mean, float* __restrict__ rstd, int C, float eps) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  mean[c] = run_mean[c];
  rstd[c] = rsqrtf(run_var[c] + eps);
}

// y = [relu]( (x - mean) * rstd * gamma + beta [+ res] )
// The grid stride is a multiple of C / 8 (the host rounds the grid), so a thread owns ONE group of 8 channels for its
// whole loop: the per-channel scale / shift live in registers and the loop body is loads, 8 FMAs and a store — no
// per-element integer division, no per-element parameter loads (the first version spent its time on both: 1.6 TB/s).
// UNR independent 16-byte loads per operand are issued before the first use to cover the HBM latency.
template <bool RES, bool RELU>
__global__ void __launch_bounds__(256, 4) bn2d_apply_kernel(const bf16* __restrict__ x, const float* __restrict__ mean,
                                                            const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, const bf16* __restrict__ res,
                                                            bf16* __restrict__ y, size_t total, int vc) {
  constexpr int UNR = RES ? 2 : 4;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  const size_t t0 = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int c = static_cast<int>(t0 % static_cast<size_t>(vc)) * 8;
  float sc[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    sc[i] = rstd[c + i] * gamma[c + i];
    sh[i] = beta[c + i] - mean[c + i] * sc[i];
  }
  const uint4* xv = reinterpret_cast<const uint4*>(x);
  const uint4* rv = reinterpret_cast<const uint4*>(res);
  uint4* yv = reinterpret_cast<uint4*>(y);
  for (size_t t = t0; t < total; t += stride * UNR) {
    uint4 xin[UNR], rin[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const size_t tt = t + u * stride;
      if (tt < total) {
        xin[u] = xv[tt];
        if (RES) rin[u] = rv[tt];
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const size_t tt = t + u * stride;
      if (tt >= total) break;
      float f[8], o[8];
      unpack8(xin[u], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = fmaf(f[i], sc[i], sh[i]);
      if (RES) {
        unpack8(rin[u], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] += f[i];
      }
      if (RELU) {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = fmaxf(o[i], 0.f);
      }
      uint4 w;
      w.x = pack_bf16x2(o[0], o[1]); w.y = pack_bf16x2(o[2], o[3]);
      w.z = pack_bf16x2(o[4], o[5]); w.w = pack_bf16x2(o[6], o[7]);
      yv[tt] = w;
    }
  }
}

// dx = gamma * rstd * (g - m1 - xhat * m2), g = dy * relu-mask; optionally dres = g (gradient of the residual input).
// Same thread <-> channel-group ownership as the forward: dx = ka * g + kb * x + kc with per-channel constants in registers.
template <bool RELU, bool DRES>
__global__ void __launch_bounds__(256, 3) bn2d_bwd_apply_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                                                const bf16* __restrict__ y, const float* __restrict__ mean,
                                                                const float* __restrict__ rstd,
                                                                const float* __restrict__ gamma,
                                                                const float* __restrict__ sums, bf16* __restrict__ dx,
                                                                bf16* __restrict__ dres, size_t total, int vc, float invr) {
  constexpr int UNR = 2;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  const size_t t0 = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int C = vc * 8;
  const int c = static_cast<int>(t0 % static_cast<size_t>(vc)) * 8;
  float ka[8], kb[8], kc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float rs = rstd[c + i], m1 = sums[c + i] * invr, m2 = sums[C + c + i] * invr;
    ka[i] = gamma[c + i] * rs;
    kb[i] = -ka[i] * rs * m2;
    kc[i] = -ka[i] * m1 - kb[i] * mean[c + i];
  }
  const uint4* gv4 = reinterpret_cast<const uint4*>(dy);
  const uint4* xv4 = reinterpret_cast<const uint4*>(x);
  const uint4* yv4 = reinterpret_cast<const uint4*>(y);
  uint4* dxv = reinterpret_cast<uint4*>(dx);
  uint4* drv = reinterpret_cast<uint4*>(dres);
  for (size_t t = t0; t < total; t += stride * UNR) {
    uint4 gin[UNR], xin[UNR], yin[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const size_t tt = t + u * stride;
      if (tt < total) {
        gin[u] = gv4[tt];
        xin[u] = xv4[tt];
        if (RELU) yin[u] = yv4[tt];
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const size_t tt = t + u * stride;
      if (tt >= total) break;
      float gv[8], xv[8], o[8];
      unpack8(gin[u], gv);
      unpack8(xin[u], xv);
      if (RELU) {
        float yv[8];
        unpack8(yin[u], yv);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (yv[i] <= 0.f) gv[i] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = fmaf(ka[i], gv[i], fmaf(kb[i], xv[i], kc[i]));
      uint4 w;
      w.x = pack_bf16x2(o[0], o[1]); w.y = pack_bf16x2(o[2], o[3]);
      w.z = pack_bf16x2(o[4], o[5]); w.w = pack_bf16x2(o[6], o[7]);
      dxv[tt] = w;
      if (DRES) {
        uint4 g;
        g.x = pack_bf16x2(gv[0], gv[1]); g.y = pack_bf16x2(gv[2], gv[3]);
        g.z = pack_bf16x2(gv[4], gv[5]); g.w = pack_bf16x2(gv[6], gv[7]);
        drv[tt] = g;
      }
    }
  }
}

// grid for the channel-owning elementwise kernels: blocks * 256 must be a multiple of vc = C / 8
static inline int bn_apply_grid(size_t total, int vc, int per_sm) {
  int g = grid_for_items(total, 256, per_sm);
  int a = vc, b = 256;
  while (b) { const int r = a % b; a = b; b = r; }      // a = gcd(vc, 256)
  const int q = vc / a;                                  // the grid must be a multiple of q
  g = (g + q - 1) / q * q;
  return g;
}

__global__ void bn2d_acc_kernel(const float* __restrict__ s, float* __restrict__ dg, float* __restrict__ db, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    if (db) db[c] += s[c];
    if (dg) dg[c] += s[C + c];
  }
}

// out = a + b (bf16)
__global__ void __launch_bounds__(256) add_bf16_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b,
                                                       bf16* __restrict__ out, size_t nvec) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < nvec; t += stride) {
    float av[8], bv[8];
    unpack8(reinterpret_cast<const uint4*>(a)[t], av);
    unpack8(reinterpret_cast<const uint4*>(b)[t], bv);
    uint4 w;
    w.x = pack_bf16x2(av[0] + bv[0], av[1] + bv[1]); w.y = pack_bf16x2(av[2] + bv[2], av[3] + bv[3]);
    w.z = pack_bf16x2(av[4] + bv[4], av[5] + bv[5]); w.w = pack_bf16x2(av[6] + bv[6], av[7] + bv[7]);
    reinterpret_cast<uint4*>(out)[t] = w;
  }
}

// ------------------------------------------------------------------------------------------------ AttentionPool2d
// modified_resnet.py:72-74: tokens[b,0] = mean_p x[b,p] + pos[0]; tokens[b,1+p] = x[b,p] + pos[1+p]
__global__ void __launch_bounds__(256) attnpool_assemble_kernel(const bf16* __restrict__ x, const float* __restrict__ pos,
                                                                bf16* __restrict__ tok, int batch, int P, int C) {
  const int vc = C / 8;
  const size_t total = static_cast<size_t>(batch) * vc;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int c = static_cast<int>(t % vc) * 8;
    const size_t b = t / vc;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int p = 0; p < P; ++p) {
      float xv[8], o[8];
      unpack8(*reinterpret_cast<const uint4*>(x + (b * P + p) * C + c), xv);
#pragma unroll
      for (int i = 0; i < 8; ++i) { acc[i] += xv[i]; o[i] = xv[i] + pos[static_cast<size_t>(1 + p) * C + c + i]; }
      uint4 w;
      w.x = pack_bf16x2(o[0], o[1]); w.y = pack_bf16x2(o[2], o[3]);
      w.z = pack_bf16x2(o[4], o[5]); w.w = pack_bf16x2(o[6], o[7]);
      *reinterpret_cast<uint4*>(tok + (b * (P + 1) + 1 + p) * C + c) = w;
    }
    uint4 w;
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = acc[i] / P + pos[c + i];
    w.x = pack_bf16x2(o[0], o[1]); w.y = pack_bf16x2(o[2], o[3]);
    w.z = pack_bf16x2(o[4], o[5]); w.w = pack_bf16x2(o[6], o[7]);
    *reinterpret_cast<uint4*>(tok + (b * (P + 1)) * C + c) = w;
  }
}
// dx[b,p] = dtok[b,1+p] + dtok[b,0] / P
__global__ void __launch_bounds__(256) attnpool_assemble_bwd_kernel(const bf16* __restrict__ dtok, bf16* __restrict__ dx,
                                                                    int batch, int P, int C) {
  const int vc = C / 8;
  const size_t total = static_cast<size_t>(batch) * P * vc;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int c = static_cast<int>(t % vc) * 8;
    const size_t q = t / vc;
    const size_t b = q / P, p = q % P;
    float a[8], m[8];
    unpack8(*reinterpret_cast<const uint4*>(dtok + (b * (P + 1) + 1 + p) * C + c), a);
    unpack8(*reinterpret_cast<const uint4*>(dtok + (b * (P + 1)) * C + c), m);
    uint4 w;
    const float ip = 1.0f / P;
    w.x = pack_bf16x2(a[0] + m[0] * ip, a[1] + m[1] * ip); w.y = pack_bf16x2(a[2] + m[2] * ip, a[3] + m[3] * ip);
    w.z = pack_bf16x2(a[4] + m[4] * ip, a[5] + m[5] * ip); w.w = pack_bf16x2(a[6] + m[6] * ip, a[7] + m[7] * ip);
    *reinterpret_cast<uint4*>(dx + q * C + c) = w;
  }
}

}  // namespace dc

using namespace dc;

extern "C" {

int dc_im2col_stem(const float* images, long long sample_stride, void* col, int batch, int H, int W, dc_stream_t stream) {
  if ((H | W) & 1) return set_error("im2col_stem: H, W must be even");
  const size_t rows = static_cast<size_t>(batch) * (H / 2) * (W / 2);
  im2col_stem_kernel<<<grid_for_items(rows, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      images, sample_stride, static_cast<bf16*>(col), batch, H, W);
  DC_CHECK_LAUNCH("im2col_stem");
  return 0;
}

int dc_im2col3x3(const void* in, void* col, int batch, int H, int W, int C, dc_stream_t stream) {
  if (C & 7) return set_error("im2col3x3: C must be a multiple of 8");
  if (static_cast<long long>(batch) * H >= (1ll << 31)) return set_error("im2col3x3: batch * H must fit 31 bits");
  const int R = 9 * (C / 8);
  const int threads = R <= 256 ? (256 / R) * R : 256;
  int blocks = batch * H;
  if (blocks > sm_count() * 16) blocks = sm_count() * 16;
  im2col3x3_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(in), static_cast<bf16*>(col), batch, H, W, C);
  DC_CHECK_LAUNCH("im2col3x3");
  return 0;
}

int dc_col2im3x3(const void* dcol, void* din, int batch, int H, int W, int C, dc_stream_t stream) {
  if (C & 7) return set_error("col2im3x3: C must be a multiple of 8");
  if (static_cast<long long>(batch) * H >= (1ll << 31)) return set_error("col2im3x3: batch * H must fit 31 bits");
  int blocks = batch * H;
  if (blocks > sm_count() * 16) blocks = sm_count() * 16;
  col2im3x3_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(dcol), static_cast<bf16*>(din), batch, H, W, C);
  DC_CHECK_LAUNCH("col2im3x3");
  return 0;
}

int dc_avgpool2(const void* in, void* out, int batch, int H, int W, int C, int backward, dc_stream_t stream) {
  if ((C & 7) || ((H | W) & 1)) return set_error("avgpool2: C % 8 == 0 and even H, W required");
  int blocks = batch * (H / 2);
  if (blocks > sm_count() * 16) blocks = sm_count() * 16;
  avgpool2_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(in), static_cast<bf16*>(out), batch, H, W, C, backward);
  DC_CHECK_LAUNCH("avgpool2");
  return 0;
}

int dc_bn2d_fwd(const void* x, const float* gamma, const float* beta, const void* res, void* y, float* mean, float* rstd,
                float* running_mean, float* running_var, float* scratch, long long rows, int C, float eps, float momentum,
                int training, int relu, dc_stream_t stream) {
  if (rows <= 0) return 0;
  if (C & 7) return set_error("bn2d: C must be a multiple of 8");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (training) {
    cudaError_t e = cudaMemsetAsync(scratch, 0, 2 * static_cast<size_t>(C) * sizeof(float), st);
    if (e != cudaSuccess) return set_error_cuda("bn2d memset", e);
    dim3 grid((C + 63) / 64, static_cast<unsigned>((rows + BN_ROWS - 1) / BN_ROWS));
    bn2d_stats_kernel<<<grid, 256, 0, st>>>(static_cast<const bf16*>(x), nullptr, nullptr, nullptr, nullptr, scratch,
                                            static_cast<size_t>(rows), C, 0, 0);
    DC_CHECK_LAUNCH("bn2d_stats");
    bn2d_finalize_kernel<<<(C + 255) / 256, 256, 0, st>>>(scratch, mean, rstd, running_mean, running_var,
                                                          static_cast<double>(rows), C, eps, momentum);
    DC_CHECK_LAUNCH("bn2d_finalize");
  } else {
    if (running_mean == nullptr || running_var == nullptr) return set_error("bn2d: eval mode needs the running statistics");
    bn2d_eval_stats_kernel<<<(C + 255) / 256, 256, 0, st>>>(running_mean, running_var, mean, rstd, C, eps);
    DC_CHECK_LAUNCH("bn2d_eval_stats");
  }
  {
    const size_t total = static_cast<size_t>(rows) * (C / 8);
    const int g = bn_apply_grid(total, C / 8, 4);
    const bf16* xp = static_cast<const bf16*>(x);
    const bf16* rp = static_cast<const bf16*>(res);
    bf16* yp = static_cast<bf16*>(y);
    if (res != nullptr && relu) bn2d_apply_kernel<true, true><<<g, 256, 0, st>>>(xp, mean, rstd, gamma, beta, rp, yp, total, C / 8);
    else if (res != nullptr) bn2d_apply_kernel<true, false><<<g, 256, 0, st>>>(xp, mean, rstd, gamma, beta, rp, yp, total, C / 8);
    else if (relu) bn2d_apply_kernel<false, true><<<g, 256, 0, st>>>(xp, mean, rstd, gamma, beta, rp, yp, total, C / 8);
    else bn2d_apply_kernel<false, false><<<g, 256, 0, st>>>(xp, mean, rstd, gamma, beta, rp, yp, total, C / 8);
  }
  DC_CHECK_LAUNCH("bn2d_apply");
  return 0;
}

int dc_bn2d_bwd(const void* dy, const void* x, const void* y, const float* gamma, const float* mean, const float* rstd,
                void* dx, void* dres, float* dgamma, float* dbeta, float* scratch, long long rows, int C, int relu,
                dc_stream_t stream) {
  if (rows <= 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(scratch, 0, 2 * static_cast<size_t>(C) * sizeof(float), st);
  if (e != cudaSuccess) return set_error_cuda("bn2d memset", e);
  dim3 grid((C + 63) / 64, static_cast<unsigned>((rows + BN_ROWS - 1) / BN_ROWS));
  bn2d_stats_kernel<<<grid, 256, 0, st>>>(static_cast<const bf16*>(x), static_cast<const bf16*>(dy),
                                          static_cast<const bf16*>(y), mean, rstd, scratch, static_cast<size_t>(rows), C, 1,
                                          relu);
  DC_CHECK_LAUNCH("bn2d_bwd_stats");
  {
    const size_t total = static_cast<size_t>(rows) * (C / 8);
    const int g = bn_apply_grid(total, C / 8, 3);
    const float invr = 1.0f / static_cast<float>(rows);
    const bf16 *gp = static_cast<const bf16*>(dy), *xp = static_cast<const bf16*>(x), *yp = static_cast<const bf16*>(y);
    bf16 *dxp = static_cast<bf16*>(dx), *drp = static_cast<bf16*>(dres);
    if (relu && dres != nullptr)
      bn2d_bwd_apply_kernel<true, true><<<g, 256, 0, st>>>(gp, xp, yp, mean, rstd, gamma, scratch, dxp, drp, total, C / 8, invr);
    else if (relu)
      bn2d_bwd_apply_kernel<true, false><<<g, 256, 0, st>>>(gp, xp, yp, mean, rstd, gamma, scratch, dxp, drp, total, C / 8, invr);
    else if (dres != nullptr)
      bn2d_bwd_apply_kernel<false, true><<<g, 256, 0, st>>>(gp, xp, yp, mean, rstd, gamma, scratch, dxp, drp, total, C / 8, invr);
    else
      bn2d_bwd_apply_kernel<false, false><<<g, 256, 0, st>>>(gp, xp, yp, mean, rstd, gamma, scratch, dxp, drp, total, C / 8, invr);
  }
  DC_CHECK_LAUNCH("bn2d_bwd_apply");
  // dbeta = sum g, dgamma = sum g * xhat  (scratch) -> accumulate
  if (dgamma != nullptr || dbeta != nullptr) {
    bn2d_acc_kernel<<<(C + 255) / 256, 256, 0, st>>>(scratch, dgamma, dbeta, C);
    DC_CHECK_LAUNCH("bn2d_bwd_acc");
  }
  return 0;
}

int dc_add_bf16(const void* a, const void* b, void* out, size_t n, dc_stream_t stream) {
  if (n == 0) return 0;
  if (n & 7) return set_error("add_bf16: n must be a multiple of 8");
  add_bf16_kernel<<<grid_for_items(n / 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(a), static_cast<const bf16*>(b), static_cast<bf16*>(out), n / 8);
  DC_CHECK_LAUNCH("add_bf16");
  return 0;
}

int dc_attnpool_assemble(const void* x, const float* pos, void* tokens, int batch, int P, int C, dc_stream_t stream) {
  if (C & 7) return set_error("attnpool: C must be a multiple of 8");
  attnpool_assemble_kernel<<<grid_for_items(static_cast<size_t>(batch) * (C / 8), 256), 256, 0,
                             static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(x), pos,
                                                                  static_cast<bf16*>(tokens), batch, P, C);
  DC_CHECK_LAUNCH("attnpool_assemble");
  return 0;
}

int dc_attnpool_assemble_bwd(const void* dtokens, void* dx, int batch, int P, int C, dc_stream_t stream) {
  if (C & 7) return set_error("attnpool: C must be a multiple of 8");
  attnpool_assemble_bwd_kernel<<<grid_for_items(static_cast<size_t>(batch) * P * (C / 8), 256), 256, 0,
                                 static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(dtokens),
                                                                      static_cast<bf16*>(dx), batch, P, C);
  DC_CHECK_LAUNCH("attnpool_assemble_bwd");
  return 0;
}

}  // extern "C"
