// declip_b200 — HBM-bound row kernels: LayerNorm fwd/bwd, bias-gradient column sums, casts,
// patch/token embedding, row gather/scatter, L2-normalise, cross-entropy on logit strips.
// All loads/stores are 16-byte vectors, one warp per row where a row reduction is needed
// (no shared memory, shuffle reductions), grids sized in multiples of the SM count.
#include "common.cuh"
#include "internal.h"

namespace dc {

// ------------------------------------------------------------------------------------------------
// LayerNorm forward — reference: image_encoder/base_transformer.py:10-18 (nn.LayerNorm, eps 1e-5).
// One warp per row; the row lives in registers (NV 16-byte vectors per lane), two-pass variance.
template <int NV>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const bf16* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, bf16* __restrict__ y,
                                                     float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                     int rows, float eps) {
  constexpr int W = NV * 256;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  float g[NV][8], b[NV][8];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = (lane + 32 * j) * 8;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + c + 4));
    g[j][0] = g0.x; g[j][1] = g0.y; g[j][2] = g0.z; g[j][3] = g0.w;
    g[j][4] = g1.x; g[j][5] = g1.y; g[j][6] = g1.z; g[j][7] = g1.w;
    b[j][0] = b0.x; b[j][1] = b0.y; b[j][2] = b0.z; b[j][3] = b0.w;
    b[j][4] = b1.x; b[j][5] = b1.y; b[j][6] = b1.z; b[j][7] = b1.w;
  }
  for (int row = warp; row < rows; row += nwarps) {
    const bf16* xr = x + static_cast<size_t>(row) * W;
    float v[NV][8];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const uint4 u = *reinterpret_cast<const uint4*>(xr + (lane + 32 * j) * 8);
      float2 f;
      f = unpack_bf16x2(u.x); v[j][0] = f.x; v[j][1] = f.y;
      f = unpack_bf16x2(u.y); v[j][2] = f.x; v[j][3] = f.y;
      f = unpack_bf16x2(u.z); v[j][4] = f.x; v[j][5] = f.y;
      f = unpack_bf16x2(u.w); v[j][6] = f.x; v[j][7] = f.y;
#pragma unroll
      for (int i = 0; i < 8; ++i) s += v[j][i];
    }
    const float mean = warp_sum(s) * (1.0f / W);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
      for (int i = 0; i < 8; ++i) { const float d = v[j][i] - mean; q += d * d; }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / W) + eps);
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
    bf16* yr = y + static_cast<size_t>(row) * W;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = (v[j][i] - mean) * rstd * g[j][i] + b[j][i];
      uint4 w;
      w.x = pack_bf16x2(o[0], o[1]); w.y = pack_bf16x2(o[2], o[3]);
      w.z = pack_bf16x2(o[4], o[5]); w.w = pack_bf16x2(o[6], o[7]);
      *reinterpret_cast<uint4*>(yr + (lane + 32 * j) * 8) = w;
    }
  }
}

// LayerNorm backward: dx = [dres +] rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma;
// dgamma += sum_rows dy * xhat, dbeta += sum_rows dy, and optionally dcol += sum_rows dx (the bias gradient of
// the Linear that produced this LayerNorm's input — saves a separate column-sum pass over dx).
// One warp per row, register accumulation per lane across the warp's rows, software-pipelined loads (the next
// row's x/dy/dres are in flight while the current row is reduced), then smem + global fp32 atomics per block.
template <int NV>
struct LnRow {
  uint4 x[NV], dy[NV], dr[NV];
};
template <int NV>
__device__ __forceinline__ void ln_bwd_load(LnRow<NV>& r, const bf16* __restrict__ x, const bf16* __restrict__ dy,
                                            const bf16* __restrict__ dres, size_t base, int lane) {
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = (lane + 32 * j) * 8;
    r.x[j] = *reinterpret_cast<const uint4*>(x + base + c);
    r.dy[j] = *reinterpret_cast<const uint4*>(dy + base + c);
    if (dres != nullptr) r.dr[j] = *reinterpret_cast<const uint4*>(dres + base + c);
  }
}
template <int NV>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                                     const float* __restrict__ gamma, const float* __restrict__ mean,
                                                     const float* __restrict__ rstd, const bf16* __restrict__ dres,
                                                     bf16* __restrict__ dx, float* __restrict__ dgamma,
                                                     float* __restrict__ dbeta, float* __restrict__ dcol, int rows) {
  constexpr int W = NV * 256;
  __shared__ float s_acc[3 * W];
  for (int i = threadIdx.x; i < 3 * W; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  float g[NV][8], ag[NV][8], ab[NV][8], ao[NV][8];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = (lane + 32 * j) * 8;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
    g[j][0] = g0.x; g[j][1] = g0.y; g[j][2] = g0.z; g[j][3] = g0.w;
    g[j][4] = g1.x; g[j][5] = g1.y; g[j][6] = g1.z; g[j][7] = g1.w;
#pragma unroll
    for (int i = 0; i < 8; ++i) { ag[j][i] = 0.f; ab[j][i] = 0.f; ao[j][i] = 0.f; }
  }
  LnRow<NV> cur, nxt;
  int row = warp;
  if (row < rows) ln_bwd_load<NV>(cur, x, dy, dres, static_cast<size_t>(row) * W, lane);
  for (; row < rows; row += nwarps) {
    const int nrow = row + nwarps;
    if (nrow < rows) ln_bwd_load<NV>(nxt, x, dy, dres, static_cast<size_t>(nrow) * W, lane);
    const size_t base = static_cast<size_t>(row) * W;
    const float mu = mean[row], rs = rstd[row];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      float xv[8], dv[8];
      unpack8(cur.x[j], xv);
      unpack8(cur.dy[j], dv);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float xh = (xv[i] - mu) * rs;
        const float gy = dv[i] * g[j][i];
        s1 += gy;
        s2 += gy * xh;
        ag[j][i] += dv[i] * xh;
        ab[j][i] += dv[i];
      }
    }
    const float c1 = warp_sum(s1) * (1.0f / W);
    const float c2 = warp_sum(s2) * (1.0f / W);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = (lane + 32 * j) * 8;
      float xv[8], dv[8], o[8];
      unpack8(cur.x[j], xv);
      unpack8(cur.dy[j], dv);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float xh = (xv[i] - mu) * rs;
        o[i] = rs * (dv[i] * g[j][i] - c1 - xh * c2);
      }
      if (dres != nullptr) {
        float rv[8];
        unpack8(cur.dr[j], rv);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] += rv[i];
      }
      if (dcol != nullptr) {
#pragma unroll
        for (int i = 0; i < 8; ++i) ao[j][i] += o[i];
      }
      uint4 w;
      w.x = pack_bf16x2(o[0], o[1]); w.y = pack_bf16x2(o[2], o[3]);
      w.z = pack_bf16x2(o[4], o[5]); w.w = pack_bf16x2(o[6], o[7]);
      *reinterpret_cast<uint4*>(dx + base + c) = w;
    }
    cur = nxt;
  }
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = (lane + 32 * j) * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      atomicAdd(&s_acc[c + i], ag[j][i]);
      atomicAdd(&s_acc[W + c + i], ab[j][i]);
      if (dcol != nullptr) atomicAdd(&s_acc[2 * W + c + i], ao[j][i]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < W; i += blockDim.x) {
    atomicAdd(&dgamma[i], s_acc[i]);
    atomicAdd(&dbeta[i], s_acc[W + i]);
    if (dcol != nullptr) atomicAdd(&dcol[i], s_acc[2 * W + i]);
  }
}

// Round-2 variant: a GROUP of NV warps per row, ONE 16-byte vector per lane and tensor.  The one-warp-per-row kernel above
// needs 188-253 registers per thread (72-96 fp32 column accumulators per lane) and so runs 8 warps per SM: ncu shows the
// issue slots 33 % busy and 12 % of the warp slots occupied — latency-bound at 0.56 of the HBM floor.  Here every lane
// owns 8 columns (24 accumulators, ~60 registers), 24 warps fit an SM, and the two row statistics are combined across the
// NV warps of a group through shared memory with one named barrier per row (double-buffered slots).
template <int NV>
__global__ void __launch_bounds__(NV * 32 * (24 / NV)) ln_bwd_group_kernel(
    const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ mean,
    const float* __restrict__ rstd, const bf16* __restrict__ dres, bf16* __restrict__ dx, float* __restrict__ dgamma,
    float* __restrict__ dbeta, float* __restrict__ dcol, int rows) {
  constexpr int W = NV * 256;
  constexpr int GROUPS = 24 / NV;                 // row groups per block (24 warps)
  __shared__ float s_acc[3 * W];
  __shared__ float s_part[2][GROUPS][NV][2];
  for (int i = threadIdx.x; i < 3 * W; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int grp = warp / NV, wig = warp - grp * NV;            // group inside the block, warp inside the group
  const int col = (wig * 32 + lane) * 8;
  const int row_stride = gridDim.x * GROUPS;
  float g[8], ag[8], ab[8], ao[8];
  {
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + col));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + col + 4));
    g[0] = g0.x; g[1] = g0.y; g[2] = g0.z; g[3] = g0.w; g[4] = g1.x; g[5] = g1.y; g[6] = g1.z; g[7] = g1.w;
#pragma unroll
    for (int i = 0; i < 8; ++i) { ag[i] = 0.f; ab[i] = 0.f; ao[i] = 0.f; }
  }
  int row = blockIdx.x * GROUPS + grp;
  uint4 cx = make_uint4(0u, 0u, 0u, 0u), cdy = cx, cdr = cx;
  if (row < rows) {
    const size_t base = static_cast<size_t>(row) * W + col;
    cx = *reinterpret_cast<const uint4*>(x + base);
    cdy = *reinterpret_cast<const uint4*>(dy + base);
    if (dres != nullptr) cdr = *reinterpret_cast<const uint4*>(dres + base);
  }
  int it = 0;
  for (; row < rows; row += row_stride, ++it) {
    const int nrow = row + row_stride;
    uint4 nx = make_uint4(0u, 0u, 0u, 0u), ndy = nx, ndr = nx;
    if (nrow < rows) {                     // next row in flight while this one is reduced
      const size_t nb = static_cast<size_t>(nrow) * W + col;
      nx = *reinterpret_cast<const uint4*>(x + nb);
      ndy = *reinterpret_cast<const uint4*>(dy + nb);
      if (dres != nullptr) ndr = *reinterpret_cast<const uint4*>(dres + nb);
    }
    const float mu = __ldg(mean + row), rs = __ldg(rstd + row);
    float xv[8], dv[8], xh[8];
    unpack8(cx, xv);
    unpack8(cdy, dv);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      xh[i] = (xv[i] - mu) * rs;
      const float gy = dv[i] * g[i];
      s1 += gy;
      s2 = fmaf(gy, xh[i], s2);
      ag[i] = fmaf(dv[i], xh[i], ag[i]);
      ab[i] += dv[i];
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (NV > 1) {
      float* slot = &s_part[it & 1][grp][0][0];
      if (lane == 0) { slot[wig * 2] = s1; slot[wig * 2 + 1] = s2; }
      asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(NV * 32) : "memory");
      s1 = 0.f; s2 = 0.f;
#pragma unroll
      for (int k = 0; k < NV; ++k) { s1 += slot[k * 2]; s2 += slot[k * 2 + 1]; }
    }
    const float c1 = s1 * (1.0f / W), c2 = s2 * (1.0f / W);
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = rs * (dv[i] * g[i] - c1 - xh[i] * c2);
    if (dres != nullptr) {
      float rv[8];
      unpack8(cdr, rv);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] += rv[i];
    }
    if (dcol != nullptr) {
#pragma unroll
      for (int i = 0; i < 8; ++i) ao[i] += o[i];
    }
    uint4 w;
    w.x = pack_bf16x2(o[0], o[1]); w.y = pack_bf16x2(o[2], o[3]);
    w.z = pack_bf16x2(o[4], o[5]); w.w = pack_bf16x2(o[6], o[7]);
    *reinterpret_cast<uint4*>(dx + static_cast<size_t>(row) * W + col) = w;
    cx = nx; cdy = ndy; cdr = ndr;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    atomicAdd(&s_acc[col + i], ag[i]);
    atomicAdd(&s_acc[W + col + i], ab[i]);
    if (dcol != nullptr) atomicAdd(&s_acc[2 * W + col + i], ao[i]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < W; i += blockDim.x) {
    atomicAdd(&dgamma[i], s_acc[i]);
    atomicAdd(&dbeta[i], s_acc[W + i]);
    if (dcol != nullptr) atomicAdd(&dcol[i], s_acc[2 * W + i]);
  }
}

// Third LayerNorm backward: the same thread <-> 8-column ownership and row groups as ln_bwd_group_kernel, but the rows
// arrive through a 4-stage shared-memory ring filled by a producer warp with cp.async.bulk (a tile of G consecutive rows
// of x / dy / dres is one contiguous run, so each operand is ONE bulk copy of 12 KiB).  Up to 144 KiB per SM are in flight
// independent of the consumers' registers — the register-prefetch kernels keep one row (48 B) per thread in flight and sit
// at 0.4 of the HBM floor.
template <int NV>
__global__ void __launch_bounds__(25 * 32, 1) ln_bwd_pipe_kernel(
    const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ mean,
    const float* __restrict__ rstd, const bf16* __restrict__ dres, bf16* __restrict__ dx, float* __restrict__ dgamma,
    float* __restrict__ dbeta, float* __restrict__ dcol, int rows) {
  constexpr int W = NV * 256;
  constexpr int G = 24 / NV;                      // rows per tile = row groups per block
  constexpr int ST = 4;
  constexpr int ROWB = W * 2, TILE_B = G * ROWB;  // 12288 for every NV
  extern __shared__ uint8_t ln_smem_raw[];
  uint8_t* stg = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ln_smem_raw) + 127) & ~uintptr_t(127));
  float* s_acc = reinterpret_cast<float*>(stg + ST * 3 * TILE_B);
  float* s_part = s_acc + 3 * W;                  // [2][G][NV][2]
  uint64_t* full = reinterpret_cast<uint64_t*>(s_part + 2 * G * NV * 2);
  uint64_t* empty = full + ST;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 3 * W; i += blockDim.x) s_acc[i] = 0.f;
  if (threadIdx.x == 0) {
    for (int s = 0; s < ST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 24); }
    fence_mbar_init();
  }
  __syncthreads();
  const int ntiles = (rows + G - 1) / G;
  const int nop = dres != nullptr ? 3 : 2;
  if (warp == 24) {
    if (lane == 0) {
      int k = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++k) {
        const int s = k % ST;
        const uint32_t ph = (k / ST) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        const int row0 = tile * G;
        const int nvalid = rows - row0 < G ? rows - row0 : G;
        const uint32_t bytes = static_cast<uint32_t>(nvalid) * ROWB;
        mbar_arrive_expect_tx(&full[s], bytes * nop);
        uint8_t* base = stg + s * 3 * TILE_B;
        const size_t off = static_cast<size_t>(row0) * W;
        bulk_load_1d(base, x + off, bytes, &full[s]);
        bulk_load_1d(base + TILE_B, dy + off, bytes, &full[s]);
        if (dres != nullptr) bulk_load_1d(base + 2 * TILE_B, dres + off, bytes, &full[s]);
      }
    }
  } else {
    const int grp = warp / NV, wig = warp - grp * NV;
    const int col = (wig * 32 + lane) * 8;
    float g[8], ag[8], ab[8], ao[8];
    {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + col));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + col + 4));
      g[0] = g0.x; g[1] = g0.y; g[2] = g0.z; g[3] = g0.w; g[4] = g1.x; g[5] = g1.y; g[6] = g1.z; g[7] = g1.w;
#pragma unroll
      for (int i = 0; i < 8; ++i) { ag[i] = 0.f; ab[i] = 0.f; ao[i] = 0.f; }
    }
    int k = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++k) {
      const int s = k % ST;
      const uint32_t ph = (k / ST) & 1;
      const int row = tile * G + grp;
      const bool valid = row < rows;                 // uniform over the group's warps
      float mu = 0.f, rs = 0.f;
      if (valid) { mu = __ldg(mean + row); rs = __ldg(rstd + row); }
      mbar_wait(&full[s], ph);
      if (valid) {
        const uint8_t* base = stg + s * 3 * TILE_B + grp * ROWB + col * 2;
        const uint4 cx = *reinterpret_cast<const uint4*>(base);
        const uint4 cdy = *reinterpret_cast<const uint4*>(base + TILE_B);
        float xv[8], dv[8], xh[8];
        unpack8(cx, xv);
        unpack8(cdy, dv);
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          xh[i] = (xv[i] - mu) * rs;
          const float gy = dv[i] * g[i];
          s1 += gy;
          s2 = fmaf(gy, xh[i], s2);
          ag[i] = fmaf(dv[i], xh[i], ag[i]);
          ab[i] += dv[i];
        }
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        if (NV > 1) {
          float* slot = s_part + ((k & 1) * G + grp) * NV * 2;
          if (lane == 0) { slot[wig * 2] = s1; slot[wig * 2 + 1] = s2; }
          asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(NV * 32) : "memory");
          s1 = 0.f; s2 = 0.f;
#pragma unroll
          for (int q = 0; q < NV; ++q) { s1 += slot[q * 2]; s2 += slot[q * 2 + 1]; }
        }
        const float c1 = s1 * (1.0f / W), c2 = s2 * (1.0f / W);
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = rs * (dv[i] * g[i] - c1 - xh[i] * c2);
        if (dres != nullptr) {
          float rv[8];
          unpack8(*reinterpret_cast<const uint4*>(base + 2 * TILE_B), rv);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] += rv[i];
        }
        if (dcol != nullptr) {
#pragma unroll
          for (int i = 0; i < 8; ++i) ao[i] += o[i];
        }
        uint4 w;
        w.x = pack_bf16x2(o[0], o[1]); w.y = pack_bf16x2(o[2], o[3]);
        w.z = pack_bf16x2(o[4], o[5]); w.w = pack_bf16x2(o[6], o[7]);
        *reinterpret_cast<uint4*>(dx + static_cast<size_t>(row) * W + col) = w;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      atomicAdd(&s_acc[col + i], ag[i]);
      atomicAdd(&s_acc[W + col + i], ab[i]);
      if (dcol != nullptr) atomicAdd(&s_acc[2 * W + col + i], ao[i]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < W; i += blockDim.x) {
    atomicAdd(&dgamma[i], s_acc[i]);
    atomicAdd(&dbeta[i], s_acc[W + i]);
    if (dcol != nullptr) atomicAdd(&dcol[i], s_acc[2 * W + i]);
  }
}

template <int NV>
static int launch_ln_bwd_pipe(const bf16* dy, const bf16* x, const float* gamma, const float* mean, const float* rstd,
                              const bf16* dres, bf16* dx, float* dgamma, float* dbeta, float* dcol, int rows, cudaStream_t st) {
  constexpr int W = NV * 256, G = 24 / NV;
  constexpr int SMEM = 4 * 3 * 12288 + 3 * W * 4 + 2 * 24 * 2 * 4 + 8 * 8 + 128;
  auto kern = ln_bwd_pipe_kernel<NV>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return set_error_cuda("cudaFuncSetAttribute(ln_bwd_pipe)", e);
    attr = true;
  }
  int grid = (rows + G - 1) / G;
  if (grid > sm_count()) grid = sm_count();
  kern<<<grid, 25 * 32, SMEM, st>>>(dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, dcol, rows);
  return 0;
}


// ------------------------------------------------------------------------------------------------
// Column sums (bias gradients): out[c] += sum_r x[r,c].  Block = 8 column-vectors x 32 row lanes,
// 64 columns x ROWS_PER_BLOCK rows per block; warp loads are 4 rows x 128 contiguous bytes.
constexpr int COLSUM_ROWS = 1024;
__global__ void __launch_bounds__(256) colsum_kernel(const bf16* __restrict__ x, int ldx, float* __restrict__ out,
                                                     int rows, int cols) {
  __shared__ float s_part[32][65];
  const int cv = threadIdx.x & 7;
  const int rl = threadIdx.x >> 3;
  const int c0 = blockIdx.x * 64 + cv * 8;
  const int r0 = blockIdx.y * COLSUM_ROWS;
  const int r1 = min(rows, r0 + COLSUM_ROWS);
  float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (c0 < cols) {
    for (int r = r0 + rl; r < r1; r += 32) {
      const uint4 u = *reinterpret_cast<const uint4*>(x + static_cast<size_t>(r) * ldx + c0);
      float2 f;
      f = unpack_bf16x2(u.x); a[0] += f.x; a[1] += f.y;
      f = unpack_bf16x2(u.y); a[2] += f.x; a[3] += f.y;
      f = unpack_bf16x2(u.z); a[4] += f.x; a[5] += f.y;
      f = unpack_bf16x2(u.w); a[6] += f.x; a[7] += f.y;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) s_part[rl][cv * 8 + i] = a[i];
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 32; ++r) s += s_part[r][threadIdx.x];
    const int c = blockIdx.x * 64 + threadIdx.x;
    if (c < cols) atomicAdd(&out[c], s);
  }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cast_kernel(const float* __restrict__ src, bf16* __restrict__ dst, size_t n) {
  const size_t nv = n >> 3;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nv; i += stride) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src) + 2 * i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 2 * i + 1);
    uint4 w;
    w.x = pack_bf16x2(a.x, a.y); w.y = pack_bf16x2(a.z, a.w);
    w.z = pack_bf16x2(b.x, b.y); w.w = pack_bf16x2(b.z, b.w);
    reinterpret_cast<uint4*>(dst)[i] = w;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 7)) {
    const size_t i = (nv << 3) + threadIdx.x;
    dst[i] = __float2bfloat16(src[i]);
  }
}

__global__ void __launch_bounds__(256) uncast_kernel(const bf16* __restrict__ src, float* __restrict__ dst, size_t n,
                                                     float scale, bool accumulate) {
  const size_t nv = n >> 3;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nv; i += stride) {
    float v[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(src) + i), v);
    float4* d = reinterpret_cast<float4*>(dst) + 2 * i;
    float4 a = make_float4(v[0] * scale, v[1] * scale, v[2] * scale, v[3] * scale);
    float4 b = make_float4(v[4] * scale, v[5] * scale, v[6] * scale, v[7] * scale);
    if (accumulate) {
      const float4 pa = d[0], pb = d[1];
      a.x += pa.x; a.y += pa.y; a.z += pa.z; a.w += pa.w;
      b.x += pb.x; b.y += pb.y; b.z += pb.z; b.w += pb.w;
    }
    d[0] = a;
    d[1] = b;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 7)) {
    const size_t i = (nv << 3) + threadIdx.x;
    const float v = __bfloat162float(src[i]) * scale;
    dst[i] = accumulate ? dst[i] + v : v;
  }
}

__global__ void __launch_bounds__(256) multi_cast_kernel(const dc_cast_entry* __restrict__ table) {
  const dc_cast_entry e = table[blockIdx.y];
  const float* src = e.src;
  bf16* dst = static_cast<bf16*>(e.dst);
  const size_t n = e.numel;
  const size_t nv = n >> 3;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nv; i += stride) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src) + 2 * i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 2 * i + 1);
    uint4 w;
    w.x = pack_bf16x2(a.x, a.y); w.y = pack_bf16x2(a.z, a.w);
    w.z = pack_bf16x2(b.x, b.y); w.w = pack_bf16x2(b.z, b.w);
    reinterpret_cast<uint4*>(dst)[i] = w;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 7)) {
    const size_t i = (nv << 3) + threadIdx.x;
    dst[i] = __float2bfloat16(src[i]);
  }
}

// ------------------------------------------------------------------------------------------------
// ViT patchify: conv1 (kernel == stride == patch, no bias) as a GEMM over non-overlapping patches —
// visual_transformer.py:56-59.  dst row = (b, gy, gx), dst col = (c, py, px)  == conv1.weight.view(width, -1).
__global__ void __launch_bounds__(256) patchify_kernel(const float* __restrict__ img, long long sample_stride,
                                                       bf16* __restrict__ dst, int batch, int res, int patch) {
  const int G = res / patch;
  const int kdim = 3 * patch * patch;
  const int vec_per_row = kdim / 8;
  const size_t total = static_cast<size_t>(batch) * G * G * vec_per_row;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t v = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; v < total; v += stride) {
    const int kv = static_cast<int>(v % vec_per_row);
    const size_t row = v / vec_per_row;
    const int gx = static_cast<int>(row % G);
    const int gy = static_cast<int>((row / G) % G);
    const size_t b = row / (static_cast<size_t>(G) * G);
    const int k = kv * 8;
    const int c = k / (patch * patch);
    const int py = (k / patch) % patch;
    const int px = k % patch;
    const float* s = img + b * sample_stride + (static_cast<size_t>(c) * res + (gy * patch + py)) * res + gx * patch + px;
    const float4 a = __ldg(reinterpret_cast<const float4*>(s));
    const float4 bb = __ldg(reinterpret_cast<const float4*>(s) + 1);
    uint4 w;
    w.x = pack_bf16x2(a.x, a.y); w.y = pack_bf16x2(a.z, a.w);
    w.z = pack_bf16x2(bb.x, bb.y); w.w = pack_bf16x2(bb.z, bb.w);
    *reinterpret_cast<uint4*>(dst + row * kdim + k) = w;
  }
}

// visual_transformer.py:60-62: prepend class embedding, add positional embedding.
__global__ void __launch_bounds__(256) vit_assemble_kernel(const bf16* __restrict__ patch_out,
                                                           const float* __restrict__ cls, const float* __restrict__ pos,
                                                           bf16* __restrict__ tokens, int batch, int g2, int width) {
  const int L = g2 + 1;
  const int vw = width / 8;
  const size_t total = static_cast<size_t>(batch) * L * vw;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t v = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; v < total; v += stride) {
    const int c = static_cast<int>(v % vw) * 8;
    const size_t tok = v / vw;
    const int l = static_cast<int>(tok % L);
    const size_t b = tok / L;
    float o[8];
    const float4 p0 = __ldg(reinterpret_cast<const float4*>(pos + static_cast<size_t>(l) * width + c));
    const float4 p1 = __ldg(reinterpret_cast<const float4*>(pos + static_cast<size_t>(l) * width + c + 4));
    o[0] = p0.x; o[1] = p0.y; o[2] = p0.z; o[3] = p0.w; o[4] = p1.x; o[5] = p1.y; o[6] = p1.z; o[7] = p1.w;
    if (l == 0) {
      const float4 c0 = __ldg(reinterpret_cast<const float4*>(cls + c));
      const float4 c1 = __ldg(reinterpret_cast<const float4*>(cls + c + 4));
      o[0] += c0.x; o[1] += c0.y; o[2] += c0.z; o[3] += c0.w; o[4] += c1.x; o[5] += c1.y; o[6] += c1.z; o[7] += c1.w;
    } else {
      const uint4 u = *reinterpret_cast<const uint4*>(patch_out + (b * g2 + (l - 1)) * width + c);
      float2 f;
      f = unpack_bf16x2(u.x); o[0] += f.x; o[1] += f.y;
      f = unpack_bf16x2(u.y); o[2] += f.x; o[3] += f.y;
      f = unpack_bf16x2(u.z); o[4] += f.x; o[5] += f.y;
      f = unpack_bf16x2(u.w); o[6] += f.x; o[7] += f.y;
    }
    uint4 w;
    w.x = pack_bf16x2(o[0], o[1]); w.y = pack_bf16x2(o[2], o[3]);
    w.z = pack_bf16x2(o[4], o[5]); w.w = pack_bf16x2(o[6], o[7]);
    *reinterpret_cast<uint4*>(tokens + tok * width + c) = w;
  }
}

// text_transformer.py:188-190: token embedding gather + positional embedding.
__global__ void __launch_bounds__(256) text_embed_kernel(const long long* __restrict__ ids,
                                                         const float* __restrict__ table, const float* __restrict__ pos,
                                                         bf16* __restrict__ x, int batch, int L, int width) {
  const int vw = width / 8;
  const size_t total = static_cast<size_t>(batch) * L * vw;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t v = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; v < total; v += stride) {
    const int c = static_cast<int>(v % vw) * 8;
    const size_t tok = v / vw;
    const int l = static_cast<int>(tok % L);
    const long long id = ids[tok];
    const float* t = table + static_cast<size_t>(id) * width + c;
    const float* p = pos + static_cast<size_t>(l) * width + c;
    const float4 t0 = __ldg(reinterpret_cast<const float4*>(t));
    const float4 t1 = __ldg(reinterpret_cast<const float4*>(t) + 1);
    const float4 p0 = __ldg(reinterpret_cast<const float4*>(p));
    const float4 p1 = __ldg(reinterpret_cast<const float4*>(p) + 1);
    uint4 w;
    w.x = pack_bf16x2(t0.x + p0.x, t0.y + p0.y); w.y = pack_bf16x2(t0.z + p0.z, t0.w + p0.w);
    w.z = pack_bf16x2(t1.x + p1.x, t1.y + p1.y); w.w = pack_bf16x2(t1.z + p1.z, t1.w + p1.w);
    *reinterpret_cast<uint4*>(x + tok * width + c) = w;
  }
}

// Embedding backward: scatter-add into the dense fp32 table gradient with vector reductions.
// `last_row` (may be NULL): per-sample row index of the EOT token; rows after it carry an exactly-zero gradient
// (causal attention, nothing downstream reads them) and are skipped — this also removes the atomic pile-up of all
// padding positions on embedding row 0.
__global__ void __launch_bounds__(256) text_embed_bwd_kernel(const long long* __restrict__ ids,
                                                             const bf16* __restrict__ dx, float* __restrict__ dtable,
                                                             const int* __restrict__ last_row, int batch, int L,
                                                             int width) {
  const int vw = width / 8;
  const size_t total = static_cast<size_t>(batch) * L * vw;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t v = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; v < total; v += stride) {
    const int c = static_cast<int>(v % vw) * 8;
    const size_t tok = v / vw;
    if (last_row != nullptr && static_cast<int>(tok) > last_row[tok / L]) continue;
    const long long id = ids[tok];
    const uint4 u = *reinterpret_cast<const uint4*>(dx + tok * width + c);
    const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
    float* d = dtable + static_cast<size_t>(id) * width + c;
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d), "f"(f0.x), "f"(f0.y), "f"(f1.x), "f"(f1.y)
                 : "memory");
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d + 4), "f"(f2.x), "f"(f2.y), "f"(f3.x),
                 "f"(f3.y)
                 : "memory");
  }
}

__global__ void __launch_bounds__(256) gather_rows_kernel(const bf16* __restrict__ src, const int* __restrict__ idx,
                                                          bf16* __restrict__ dst, int n, int width, int scatter) {
  const int vw = width / 8;
  const size_t total = static_cast<size_t>(n) * vw;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t v = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; v < total; v += stride) {
    const int c = static_cast<int>(v % vw) * 8;
    const size_t i = v / vw;
    const size_t r = static_cast<size_t>(idx[i]);
    if (!scatter)
      *reinterpret_cast<uint4*>(dst + i * width + c) = *reinterpret_cast<const uint4*>(src + r * width + c);
    else
      *reinterpret_cast<uint4*>(dst + r * width + c) = *reinterpret_cast<const uint4*>(src + i * width + c);
  }
}

// text_transformer.py:203: the EOT token has the highest id in each sequence -> argmax position.
__global__ void eot_index_kernel(const long long* __restrict__ ids, int* __restrict__ eot, int batch, int L) {
  const int lane = threadIdx.x & 31;
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= batch) return;
  long long best = -1;
  int bi = 0;
  for (int l = lane; l < L; l += 32) {
    const long long v = ids[static_cast<size_t>(b) * L + l];
    if (v > best) { best = v; bi = l; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const long long ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if (lane == 0) eot[b] = b * L + bi;
}

// ------------------------------------------------------------------------------------------------
// clip.py:129-130: y = x / (||x|| + eps); one warp per row.
__global__ void __launch_bounds__(256) l2norm_fwd_kernel(const float* __restrict__ x, bf16* __restrict__ y,
                                                         float* __restrict__ y32, float* __restrict__ inv_out, int n,
                                                         int dim, float eps) {
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  const float* xr = x + static_cast<size_t>(row) * dim;
  float s = 0.f;
  for (int c = lane * 4; c < dim; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(xr + c);
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  const float inv = 1.0f / (sqrtf(warp_sum(s)) + eps);
  if (lane == 0 && inv_out) inv_out[row] = inv;
  for (int c = lane * 4; c < dim; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(xr + c);
    if (y != nullptr) {
      uint2 w;
      w.x = pack_bf16x2(v.x * inv, v.y * inv);
      w.y = pack_bf16x2(v.z * inv, v.w * inv);
      *reinterpret_cast<uint2*>(y + static_cast<size_t>(row) * dim + c) = w;
    }
    if (y32 != nullptr)
      *reinterpret_cast<float4*>(y32 + static_cast<size_t>(row) * dim + c) =
          make_float4(v.x * inv, v.y * inv, v.z * inv, v.w * inv);
  }
}

// dx = inv * dy - x * <dy, x> * inv^2 / r,  r = ||x||, inv = 1/(r + eps).
__global__ void __launch_bounds__(256) l2norm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                         float* __restrict__ dx, int n, int dim, float eps) {
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  const float* xr = x + static_cast<size_t>(row) * dim;
  const float* dr = dy + static_cast<size_t>(row) * dim;
  float s = 0.f, d = 0.f;
  for (int c = lane * 4; c < dim; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(xr + c);
    const float4 g = *reinterpret_cast<const float4*>(dr + c);
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    d += v.x * g.x + v.y * g.y + v.z * g.z + v.w * g.w;
  }
  const float r = sqrtf(warp_sum(s));
  d = warp_sum(d);
  const float inv = 1.0f / (r + eps);
  const float k = d * inv * inv / fmaxf(r, 1e-30f);
  float* o = dx + static_cast<size_t>(row) * dim;
  for (int c = lane * 4; c < dim; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(xr + c);
    const float4 g = *reinterpret_cast<const float4*>(dr + c);
    *reinterpret_cast<float4*>(o + c) =
        make_float4(inv * g.x - k * v.x, inv * g.y - k * v.y, inv * g.z - k * v.z, inv * g.w - k * v.w);
  }
}

// ------------------------------------------------------------------------------------------------
// ClipInfoCELoss on one [rows, cols] logit strip (loss_functions/loss.py:40-50) fused with
// accuracy top-1/top-5 (utils/misc.py:415-428).  One block per row; the row (<= 16 KiB) is re-read
// from L1/L2, never re-fetched from HBM.
__device__ __forceinline__ float block_reduce(float v, float* s_red, bool is_max) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  float r = s_red[0];
  for (int i = 1; i < (blockDim.x >> 5); ++i) r = is_max ? fmaxf(r, s_red[i]) : r + s_red[i];
  return r;
}

__global__ void __launch_bounds__(256) ce_fwd_kernel(const float* __restrict__ logits, int ld, int rows, int cols,
                                                     int label0, const long long* __restrict__ labels,
                                                     const int* __restrict__ skip_col, float* __restrict__ loss_sum,
                                                     int* __restrict__ top1, int* __restrict__ top5,
                                                     float* __restrict__ lse_out) {
  __shared__ float s_red[8];
  const int row = blockIdx.x;
  const float* z = logits + static_cast<size_t>(row) * ld;
  const int label = labels ? static_cast<int>(labels[row]) : label0 + row;
  const int skip = skip_col ? skip_col[row] : -1;   // column excluded from the softmax (NT-Xent self-similarity)
  float m = -INFINITY;
  for (int c = threadIdx.x; c < cols; c += blockDim.x)
    if (c != skip) m = fmaxf(m, z[c]);
  m = block_reduce(m, s_red, true);
  const float zl = z[label];
  float s = 0.f, gt = 0.f;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    if (c == skip) continue;
    const float v = z[c];
    s += __expf(v - m);
    gt += (v > zl) ? 1.f : 0.f;
  }
  s = block_reduce(s, s_red, false);
  gt = block_reduce(gt, s_red, false);
  if (threadIdx.x == 0) {
    const float lse = m + __logf(s);
    if (lse_out) lse_out[row] = lse;
    atomicAdd(loss_sum, lse - zl);
    if (top1 && gt < 0.5f) atomicAdd(top1, 1);
    if (top5 && gt < 4.5f) atomicAdd(top5, 1);
  }
}

// dlogits = (gscale_host * *gscale_dev) * (softmax(row) - onehot(label)); bf16 or fp32 output
template <bool OUT_F32>
__global__ void __launch_bounds__(256) ce_bwd_kernel(const float* __restrict__ logits, int ld, int rows, int cols,
                                                     int label0, const long long* __restrict__ labels,
                                                     const int* __restrict__ skip_col, const float* __restrict__ lse,
                                                     const float* __restrict__ gscale_dev, float gscale_host,
                                                     void* __restrict__ dl, int lddl) {
  const int row = blockIdx.x;
  const float* z = logits + static_cast<size_t>(row) * ld;
  const int label = labels ? static_cast<int>(labels[row]) : label0 + row;
  const int skip = skip_col ? skip_col[row] : -1;
  const float l = lse[row];
  const float gs = gscale_host * (gscale_dev ? *gscale_dev : 1.0f);
  const bool vec_ok = ((ld | lddl) & 1) == 0;
  for (int c = threadIdx.x * 2; c < cols; c += blockDim.x * 2) {
    const bool has2 = c + 1 < cols;
    float a = __expf(z[c] - l), b = has2 ? __expf(z[c + 1] - l) : 0.f;
    if (c == skip) a = 0.f;
    if (c + 1 == skip) b = 0.f;
    if (c == label) a -= 1.f;
    if (c + 1 == label) b -= 1.f;
    a *= gs;
    b *= gs;
    if (OUT_F32) {
      float* d = static_cast<float*>(dl) + static_cast<size_t>(row) * lddl + c;
      if (has2 && vec_ok) *reinterpret_cast<float2*>(d) = make_float2(a, b);
      else { d[0] = a; if (has2) d[1] = b; }
    } else {
      bf16* d = static_cast<bf16*>(dl) + static_cast<size_t>(row) * lddl + c;
      if (has2 && vec_ok) *reinterpret_cast<uint32_t*>(d) = pack_bf16x2(a, b);
      else { d[0] = __float2bfloat16(a); if (has2) d[1] = __float2bfloat16(b); }
    }
  }
}

// out += sum_i a[i] * b[i]   (fp32; d logit_scale = sum dlogits * logits / s)
__global__ void __launch_bounds__(256) dot_kernel(const float* __restrict__ a, const float* __restrict__ b, size_t n,
                                                  float* __restrict__ out) {
  __shared__ float s_red[8];
  float acc = 0.f;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) acc += a[i] * b[i];
  acc = block_reduce(acc, s_red, false);
  if (threadIdx.x == 0) atomicAdd(out, acc);
}

static inline int grid_for(size_t work_items, int threads, int per_sm) {
  size_t blocks = (work_items + threads - 1) / threads;
  const size_t cap = static_cast<size_t>(sm_count()) * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace dc

using namespace dc;

extern "C" {

int dc_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                     int rows, int width, float eps, dc_stream_t stream) {
  if (rows <= 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = grid_for(static_cast<size_t>(rows) * 32, 256, 8);
  const bf16* xp = static_cast<const bf16*>(x);
  bf16* yp = static_cast<bf16*>(y);
  switch (width) {
    case 256: ln_fwd_kernel<1><<<grid, 256, 0, st>>>(xp, gamma, beta, yp, mean, rstd, rows, eps); break;
    case 512: ln_fwd_kernel<2><<<grid, 256, 0, st>>>(xp, gamma, beta, yp, mean, rstd, rows, eps); break;
    case 768: ln_fwd_kernel<3><<<grid, 256, 0, st>>>(xp, gamma, beta, yp, mean, rstd, rows, eps); break;
    case 1024: ln_fwd_kernel<4><<<grid, 256, 0, st>>>(xp, gamma, beta, yp, mean, rstd, rows, eps); break;
    default: return set_error("layernorm: width must be 256, 512, 768 or 1024");
  }
  DC_CHECK_LAUNCH("layernorm_fwd");
  return 0;
}

int dc_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                     const void* dres, void* dx, float* dgamma, float* dbeta, float* dcol, int rows, int width,
                     dc_stream_t stream) {
  if (rows <= 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bf16* dyp = static_cast<const bf16*>(dy);
  const bf16* xp = static_cast<const bf16*>(x);
  const bf16* rp = static_cast<const bf16*>(dres);
  bf16* dxp = static_cast<bf16*>(dx);
  static const bool v1 = [] { const char* e = getenv("DC_LN_BWD_V1"); return e != nullptr && e[0] == '1'; }();
  static const bool v2 = [] { const char* e = getenv("DC_LN_BWD_GROUP"); return e != nullptr && e[0] == '1'; }();
  if (!v1 && !v2 && (width == 256 || width == 512 || width == 768 || width == 1024)) {
    int rc;
    switch (width / 256) {
      case 1: rc = launch_ln_bwd_pipe<1>(dyp, xp, gamma, mean, rstd, rp, dxp, dgamma, dbeta, dcol, rows, st); break;
      case 2: rc = launch_ln_bwd_pipe<2>(dyp, xp, gamma, mean, rstd, rp, dxp, dgamma, dbeta, dcol, rows, st); break;
      case 3: rc = launch_ln_bwd_pipe<3>(dyp, xp, gamma, mean, rstd, rp, dxp, dgamma, dbeta, dcol, rows, st); break;
      default: rc = launch_ln_bwd_pipe<4>(dyp, xp, gamma, mean, rstd, rp, dxp, dgamma, dbeta, dcol, rows, st); break;
    }
    if (rc) return rc;
    DC_CHECK_LAUNCH("layernorm_bwd");
    return 0;
  }
  if (!v1 && (width == 256 || width == 512 || width == 768 || width == 1024)) {
    // warp-group-per-row kernel: one 768-thread block per SM (24 warps), blocks walk the rows with stride grid x groups
    const int nv = width / 256;
    const int groups = 24 / nv;
    int grid = (rows + groups - 1) / groups;
    if (grid > sm_count()) grid = sm_count();
    const int threads = nv * 32 * groups;
    switch (nv) {
      case 1: ln_bwd_group_kernel<1><<<grid, threads, 0, st>>>(dyp, xp, gamma, mean, rstd, rp, dxp, dgamma, dbeta, dcol, rows); break;
      case 2: ln_bwd_group_kernel<2><<<grid, threads, 0, st>>>(dyp, xp, gamma, mean, rstd, rp, dxp, dgamma, dbeta, dcol, rows); break;
      case 3: ln_bwd_group_kernel<3><<<grid, threads, 0, st>>>(dyp, xp, gamma, mean, rstd, rp, dxp, dgamma, dbeta, dcol, rows); break;
      default: ln_bwd_group_kernel<4><<<grid, threads, 0, st>>>(dyp, xp, gamma, mean, rstd, rp, dxp, dgamma, dbeta, dcol, rows); break;
    }
    DC_CHECK_LAUNCH("layernorm_bwd");
    return 0;
  }
  // round-1 kernel: one resident block per SM (188-253 registers/thread)
  const int grid = grid_for(static_cast<size_t>(rows) * 32, 256, 1);
  switch (width) {
    case 256: ln_bwd_kernel<1><<<grid, 256, 0, st>>>(dyp, xp, gamma, mean, rstd, rp, dxp, dgamma, dbeta, dcol, rows); break;
    case 512: ln_bwd_kernel<2><<<grid, 256, 0, st>>>(dyp, xp, gamma, mean, rstd, rp, dxp, dgamma, dbeta, dcol, rows); break;
    case 768: ln_bwd_kernel<3><<<grid, 256, 0, st>>>(dyp, xp, gamma, mean, rstd, rp, dxp, dgamma, dbeta, dcol, rows); break;
    case 1024: ln_bwd_kernel<4><<<grid, 256, 0, st>>>(dyp, xp, gamma, mean, rstd, rp, dxp, dgamma, dbeta, dcol, rows); break;
    default: return set_error("layernorm: width must be 256, 512, 768 or 1024");
  }
  DC_CHECK_LAUNCH("layernorm_bwd");
  return 0;
}

int dc_colsum_bf16(const void* x, int ldx, float* out, int rows, int cols, dc_stream_t stream) {
  if (rows <= 0 || cols <= 0) return 0;
  if ((cols & 7) || (ldx & 7)) return set_error("colsum: cols and ldx must be multiples of 8");
  dim3 grid((cols + 63) / 64, (rows + COLSUM_ROWS - 1) / COLSUM_ROWS);
  colsum_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(x), ldx, out, rows, cols);
  DC_CHECK_LAUNCH("colsum");
  return 0;
}

int dc_cast_f32_bf16(const float* src, void* dst, size_t n, dc_stream_t stream) {
  if (n == 0) return 0;
  cast_kernel<<<grid_for(n / 8 + 1, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, static_cast<bf16*>(dst), n);
  DC_CHECK_LAUNCH("cast");
  return 0;
}

int dc_cast_bf16_f32(const void* src, float* dst, size_t n, float scale, int accumulate, dc_stream_t stream) {
  if (n == 0) return 0;
  if ((reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(dst) & 15)) return set_error("cast_bf16_f32: 16-byte aligned buffers");
  uncast_kernel<<<grid_for(n / 8 + 1, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(src), dst, n, scale,
                                                                                        accumulate != 0);
  DC_CHECK_LAUNCH("uncast");
  return 0;
}

int dc_multi_cast_f32_bf16(const dc_cast_entry* table_dev, int n_tensors, unsigned long long max_numel,
                           dc_stream_t stream) {
  if (n_tensors <= 0) return 0;
  int bx = static_cast<int>((max_numel / 8 + 255) / 256);
  if (bx > 64) bx = 64;
  if (bx < 1) bx = 1;
  dim3 grid(bx, n_tensors);
  multi_cast_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(table_dev);
  DC_CHECK_LAUNCH("multi_cast");
  return 0;
}

int dc_patchify(const float* images, long long sample_stride, void* patches, int batch, int res, int patch,
                dc_stream_t stream) {
  if (patch % 8 || res % patch) return set_error("patchify: patch must be a multiple of 8 and divide res");
  const size_t total = static_cast<size_t>(batch) * (res / patch) * (res / patch) * (3 * patch * patch / 8);
  patchify_kernel<<<grid_for(total, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      images, sample_stride, static_cast<bf16*>(patches), batch, res, patch);
  DC_CHECK_LAUNCH("patchify");
  return 0;
}

int dc_vit_assemble(const void* patch_out, const float* cls, const float* pos, void* tokens, int batch, int g2,
                    int width, dc_stream_t stream) {
  const size_t total = static_cast<size_t>(batch) * (g2 + 1) * (width / 8);
  vit_assemble_kernel<<<grid_for(total, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(patch_out), cls, pos, static_cast<bf16*>(tokens), batch, g2, width);
  DC_CHECK_LAUNCH("vit_assemble");
  return 0;
}

int dc_text_embed(const long long* ids, const float* table, const float* pos, void* x, int batch, int L, int width,
                  dc_stream_t stream) {
  const size_t total = static_cast<size_t>(batch) * L * (width / 8);
  text_embed_kernel<<<grid_for(total, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      ids, table, pos, static_cast<bf16*>(x), batch, L, width);
  DC_CHECK_LAUNCH("text_embed");
  return 0;
}

int dc_text_embed_bwd(const long long* ids, const void* dx, float* dtable, float* dpos, const int* last_row, int batch,
                      int L, int width, dc_stream_t stream) {
  const size_t total = static_cast<size_t>(batch) * L * (width / 8);
  if (dtable != nullptr) {
    text_embed_bwd_kernel<<<grid_for(total, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        ids, static_cast<const bf16*>(dx), dtable, last_row, batch, L, width);
    DC_CHECK_LAUNCH("text_embed_bwd");
  }
  if (dpos != nullptr) return dc_colsum_bf16(dx, L * width, dpos, batch, L * width, stream);
  return 0;
}

int dc_gather_rows(const void* src, const int* idx, void* dst, int n, int width, dc_stream_t stream) {
  if (n <= 0) return 0;
  const size_t total = static_cast<size_t>(n) * (width / 8);
  gather_rows_kernel<<<grid_for(total, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(src), idx, static_cast<bf16*>(dst), n, width, 0);
  DC_CHECK_LAUNCH("gather_rows");
  return 0;
}

int dc_scatter_rows(const void* src, const int* idx, void* dst, int n, int width, dc_stream_t stream) {
  if (n <= 0) return 0;
  const size_t total = static_cast<size_t>(n) * (width / 8);
  gather_rows_kernel<<<grid_for(total, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(src), idx, static_cast<bf16*>(dst), n, width, 1);
  DC_CHECK_LAUNCH("scatter_rows");
  return 0;
}

int dc_eot_index(const long long* ids, int* eot, int batch, int L, dc_stream_t stream) {
  if (batch <= 0) return 0;
  eot_index_kernel<<<(batch * 32 + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(ids, eot, batch, L);
  DC_CHECK_LAUNCH("eot_index");
  return 0;
}

int dc_l2norm_fwd(const float* x, void* y, float* y_f32, float* inv, int n, int dim, float eps, dc_stream_t stream) {
  if (n <= 0) return 0;
  if (dim & 3) return set_error("l2norm: dim must be a multiple of 4");
  l2norm_fwd_kernel<<<(n * 32 + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, static_cast<bf16*>(y), y_f32,
                                                                                     inv, n, dim, eps);
  DC_CHECK_LAUNCH("l2norm_fwd");
  return 0;
}

int dc_l2norm_bwd(const float* dy, const float* x, float* dx, int n, int dim, float eps, dc_stream_t stream) {
  if (n <= 0) return 0;
  if (dim & 3) return set_error("l2norm: dim must be a multiple of 4");
  l2norm_bwd_kernel<<<(n * 32 + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(dy, x, dx, n, dim, eps);
  DC_CHECK_LAUNCH("l2norm_bwd");
  return 0;
}

int dc_ce_strip_fwd(const float* logits, int ld, int rows, int cols, int label0, const long long* labels,
                    const int* skip_col, float* loss_sum, int* top1, int* top5, float* lse_out, dc_stream_t stream) {
  if (rows <= 0) return 0;
  if (labels == nullptr && (label0 < 0 || label0 + rows > cols)) return set_error("ce_strip: labels out of range");
  ce_fwd_kernel<<<rows, 256, 0, static_cast<cudaStream_t>(stream)>>>(logits, ld, rows, cols, label0, labels, skip_col,
                                                                 loss_sum, top1, top5, lse_out);
  DC_CHECK_LAUNCH("ce_strip_fwd");
  return 0;
}

int dc_ce_strip_bwd(const float* logits, int ld, int rows, int cols, int label0, const long long* labels,
                    const int* skip_col, const float* lse, const float* gscale_dev, float gscale_host, void* dlogits,
                    int lddl, int out_f32, dc_stream_t stream) {
  if (rows <= 0) return 0;
  if (out_f32)
    ce_bwd_kernel<true><<<rows, 256, 0, static_cast<cudaStream_t>(stream)>>>(logits, ld, rows, cols, label0, labels,
                                                                         skip_col, lse, gscale_dev, gscale_host, dlogits,
                                                                         lddl);
  else
    ce_bwd_kernel<false><<<rows, 256, 0, static_cast<cudaStream_t>(stream)>>>(logits, ld, rows, cols, label0, labels,
                                                                          skip_col, lse, gscale_dev, gscale_host, dlogits,
                                                                          lddl);
  DC_CHECK_LAUNCH("ce_strip_bwd");
  return 0;
}

int dc_dot_f32(const float* a, const float* b, size_t n, float* out, dc_stream_t stream) {
  if (n == 0) return 0;
  dot_kernel<<<grid_for(n, 256, 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, b, n, out);
  DC_CHECK_LAUNCH("dot");
  return 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// Fused multi-tensor AdamW (decoupled weight decay, torch.optim.AdamW semantics) — the optimiser step of the
// reference configs (experiments/*/config.yaml: optimizer type AdamW; prototype/optimizer/__init__.py:18-26).
// One launch for every parameter tensor: blockIdx.y selects the tensor from a device table.
namespace dc {
struct AdamWGroups { float lr[32]; float wd[32]; };

__global__ void __launch_bounds__(256) adamw_multi_kernel(const dc_adamw_entry* __restrict__ table, const AdamWGroups hp,
                                                          float beta1, float beta2, float eps, float bc1, float bc2) {
  const dc_adamw_entry e = table[blockIdx.y];
  float* __restrict__ p = e.param;
  const float* __restrict__ g = e.grad;
  float* __restrict__ m = e.exp_avg;
  float* __restrict__ v = e.exp_avg_sq;
  bf16* __restrict__ sh = static_cast<bf16*>(e.shadow);
  const size_t n = e.numel;
  const float lr = hp.lr[e.group], wd = hp.wd[e.group];
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  const size_t nv = n >> 2;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nv; i += stride) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = reinterpret_cast<const float4*>(g)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* pp = &pv.x; const float* gp = &gv.x; float* mp = &mv.x; float* vp = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      pp[k] *= 1.0f - lr * wd;
      mp[k] = beta1 * mp[k] + (1.0f - beta1) * gp[k];
      vp[k] = beta2 * vp[k] + (1.0f - beta2) * gp[k] * gp[k];
      const float denom = sqrtf(vp[k]) * inv_sqrt_bc2 + eps;
      pp[k] -= step_size * mp[k] / denom;
    }
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (sh != nullptr) {   // shadows are 256-byte aligned slices of the tower's flat bf16 buffer
      uint2 w;
      w.x = pack_bf16x2(pv.x, pv.y);
      w.y = pack_bf16x2(pv.z, pv.w);
      reinterpret_cast<uint2*>(sh)[i] = w;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const size_t i = (nv << 2) + threadIdx.x;
    float pk = p[i] * (1.0f - lr * wd);
    const float mk = beta1 * m[i] + (1.0f - beta1) * g[i];
    const float vk = beta2 * v[i] + (1.0f - beta2) * g[i] * g[i];
    pk -= step_size * mk / (sqrtf(vk) * inv_sqrt_bc2 + eps);
    p[i] = pk; m[i] = mk; v[i] = vk;
    if (sh != nullptr) sh[i] = __float2bfloat16(pk);
  }
}
}  // namespace dc

extern "C" int dc_adamw_multi(const dc_adamw_entry* table_dev, int n_tensors, unsigned long long max_numel,
                              const float* lr_host, const float* wd_host, int n_groups, float beta1, float beta2,
                              float eps, int step, dc_stream_t stream) {
  if (n_tensors <= 0) return 0;
  if (step < 1) return dc::set_error("adamw: step must be >= 1");
  if (n_groups < 1 || n_groups > 32 || lr_host == nullptr || wd_host == nullptr)
    return dc::set_error("adamw: 1..32 param groups with host lr / weight_decay arrays");
  dc::AdamWGroups hp;
  for (int i = 0; i < 32; ++i) {
    hp.lr[i] = i < n_groups ? lr_host[i] : 0.f;
    hp.wd[i] = i < n_groups ? wd_host[i] : 0.f;
  }
  const float bc1 = 1.0f - powf(beta1, static_cast<float>(step));
  const float bc2 = 1.0f - powf(beta2, static_cast<float>(step));
  int bx = static_cast<int>((max_numel / 4 + 255) / 256);
  if (bx > 128) bx = 128;
  if (bx < 1) bx = 1;
  dim3 grid(bx, n_tensors);
  dc::adamw_multi_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(table_dev, hp, beta1, beta2, eps, bc1, bc2);
  DC_CHECK_LAUNCH("adamw_multi");
  return 0;
}
