// declip_b200 — implicit-GEMM 3x3 / pad 1 / stride 1 convolution on NHWC bf16 (ModifiedResNet Bottleneck.conv2 and the
// stem, prototype/model/image_encoder/modified_resnet.py:23,151-154): out[p, co] = sum_{tap, ci} x[p + off(tap), ci] *
// w[co, tap, ci], accumulated in TMEM by tcgen05.mma.
//
// No im2col matrix exists.  An output tile is a spatial box of <= 128 pixels — full-width rows {W, bh, nb} of one image
// (or nb whole images when an image has <= 64 pixels) — and the A operand of k-block (tap, channel block) is ONE 4-D TMA
// box {64 channels, W, bh, nb} of the activation tensor shifted by the tap offset (dx - 1, dy - 1): the TMA unit's
// out-of-bounds zero fill IS the padding (negative / past-the-edge coordinates), and the box lands in shared memory
// as [pixel][64 channels] 128-byte-swizzled rows — exactly the K-major A tile of the GEMM kernels.  B is the weight
// matrix [Cout, 9 * Cin] (tap-major K) through an ordinary 2-D map.  The same kernel computes the input gradient: dx =
// conv3x3(dy, flip / transpose(w)) — the host passes the weights rearranged — so neither the [rows, 9C] column matrix
// of the forward nor the `dcol` matrix + col2im gather of the backward is ever written (24 + 5 ms of a 128 ms step).
//
// Launch: one CTA per (pixel tile, 64..256-wide Cout block), 192 threads: warp 0 TMA producer, warp 1 tcgen05.mma issuer
// + TMEM allocation, warps 2-5 epilogue (one output pixel per thread, bf16 stores of 64-byte channel segments).  Two
// CTAs fit an SM (<= 96 KiB of stages, <= 256 TMEM columns each; three for the 64-wide tile), so one CTA's epilogue
// overlaps the others' main loops and the 9-k-block pipeline drain of a small-channel tile is hidden behind its neighbours.
#include <string.h>
#include "common.cuh"
#include "internal.h"

namespace dc {

constexpr int CV_BM = 128;
constexpr int CV_BK = 64;
constexpr int CV_THREADS = 192;

struct ConvParams {
  int B, H, W, C, Cout;
  int bh, nb;            // rows of an image / images per pixel tile
  int tiles_y;           // pixel tiles per image (nb == 1) — otherwise 1
  int n_ptiles, n_ctiles;
  int kblocks;           // 9 * cblocks
  int cblocks;           // ceil(C / 64): a 32-channel tensor rides in half-empty boxes (TMA zero fill past the channel dim)
  bf16* out;
};

template <int BN>
__global__ void __launch_bounds__(CV_THREADS, BN == 64 ? 3 : 2)      // 64-wide tiles: 3 x 73 KiB of stages fit, one more tile in flight
conv3x3_igemm_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const ConvParams p) {
  constexpr int STAGES = BN == 256 ? 2 : 3;
  constexpr int A_BYTES = CV_BM * CV_BK * 2;
  constexpr int B_BYTES = BN * CV_BK * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ct = blockIdx.x % p.n_ctiles;            // Cout block fastest: CTAs running together share the activation tile
  const int pt = blockIdx.x / p.n_ctiles;
  int n0, y0;
  if (p.nb > 1) { n0 = pt * p.nb; y0 = 0; }
  else { n0 = pt / p.tiles_y; y0 = (pt - n0 * p.tiles_y) * p.bh; }
  const int valid_rows = p.W * p.bh * p.nb;
  const uint32_t a_tx = static_cast<uint32_t>(valid_rows) * 128u;     // bytes one activation box delivers

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tfull, 1);
    fence_mbar_init();
    prefetch_tensormap(&tmX);
    prefetch_tensormap(&tmW);
  }
  if (warp == 1) { tmem_alloc(tmem_slot, BN < 32 ? 32 : BN); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < p.kblocks; ++kb) {
        const int tap = kb / p.cblocks, cb = kb - tap * p.cblocks;
        const int dy = tap / 3 - 1, dx = tap % 3 - 1;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full_bar[stage], a_tx + B_BYTES);
        uint8_t* sa = smem + stage * STAGE_BYTES;
        tma_load_4d(sa, &tmX, &full_bar[stage], cb * CV_BK, dx, y0 + dy, n0);
        tma_load_2d(sa + A_BYTES, &tmW, &full_bar[stage], tap * p.C + cb * CV_BK, ct * BN);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(CV_BM, BN, false, false);
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < p.kblocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
        const uint32_t sb = sa + A_BYTES;
#pragma unroll
        for (int k = 0; k < CV_BK / 16; ++k)
          umma_bf16(tmem_d, umma_smem_desc(sa + k * 32, 16, 1024), umma_smem_desc(sb + k * 32, 16, 1024), idesc,
                    (kb > 0 || k > 0) ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(tfull);
    }
    __syncwarp();
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + lane;                    // row of the tile = pixel in box order (x fastest, then y, then image)
    const int x = r % p.W;
    const int yy = (r / p.W) % p.bh;
    const int nn = r / (p.W * p.bh);
    const bool ok = r < valid_rows && (y0 + yy) < p.H && (n0 + nn) < p.B;
    bf16* orow = p.out + ((static_cast<size_t>(n0 + nn) * p.H + (y0 + yy)) * p.W + x) * p.Cout + ct * BN;
    mbar_wait(tfull, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t acc[32];
      tmem_ld32(tmem_d + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(c * 32), acc);
      tmem_ld_wait();
      if (ok && ct * BN + c * 32 < p.Cout) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(acc[8 * j + 0]), __uint_as_float(acc[8 * j + 1]));
          w.y = pack_bf16x2(__uint_as_float(acc[8 * j + 2]), __uint_as_float(acc[8 * j + 3]));
          w.z = pack_bf16x2(__uint_as_float(acc[8 * j + 4]), __uint_as_float(acc[8 * j + 5]));
          w.w = pack_bf16x2(__uint_as_float(acc[8 * j + 6]), __uint_as_float(acc[8 * j + 7]));
          *reinterpret_cast<uint4*>(orow + c * 32 + 8 * j) = w;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_d, BN < 32 ? 32 : BN); }
}

template <int BN>
static int launch_conv(const CUtensorMap& tmX, const CUtensorMap& tmW, const ConvParams& p, cudaStream_t st) {
  constexpr int STAGES = BN == 256 ? 2 : 3;
  constexpr int SMEM = STAGES * (CV_BM * CV_BK * 2 + BN * CV_BK * 2) + 128 + 1024;
  auto kern = conv3x3_igemm_kernel<BN>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return set_error_cuda("cudaFuncSetAttribute(conv3x3_igemm)", e);
    attr = true;
  }
  kern<<<p.n_ptiles * p.n_ctiles, CV_THREADS, SMEM, st>>>(tmX, tmW, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error_cuda("conv3x3_igemm launch", e);
  count_launch();
  return 0;
}

// ------------------------------------------------------------------------------------------------ weight gradient
// dW[co, tap, ci] += sum_p dy[p, co] * x[p + off(tap), ci]: a GEMM whose contraction runs over the PIXELS.  Both operands are
// read MN-major straight from the NHWC tensors: a k-block is one spatial box of <= 64 pixels {W, bh} of one image — dy
// unshifted (A, 2 boxes of 64 output channels), x shifted by the tap offset with out-of-bounds zero fill (B, BN / 64 boxes of
// 64 input channels).  Boxes shorter than 64 pixels leave the tail rows of the 64-row k-block untouched: the stages are
// zeroed once, so those rows contribute nothing.  Split over the pixel k-blocks across CTAs, fp32 red.global.add into
// dW [Cout, 9 * C] (zeroed by the caller).
struct WgradParams {
  int B, H, W, C, Cout;
  int bw, bh, tiles_x, tiles_y, rows_per_kb, total_kb, ksplits;
  int co_tiles, ci_tiles;
  float* dw;
};

template <int BN>
__global__ void __launch_bounds__(CV_THREADS, 2)
conv3x3_wgrad_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX, const WgradParams p) {
  constexpr int STAGES = BN == 256 ? 2 : 3;
  constexpr int A_BYTES = 2 * 64 * 128;            // two [64 pixels][64 co] boxes
  constexpr int B_BYTES = (BN / 64) * 64 * 128;    // BN / 64 [64 pixels][64 ci] boxes
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int u = blockIdx.x;
  const int split = u % p.ksplits; u /= p.ksplits;
  const int cit = u % p.ci_tiles; u /= p.ci_tiles;
  const int tap = u % 9;
  const int cot = u / 9;
  const int dy = tap / 3 - 1, dx = tap % 3 - 1;
  const uint32_t tx = static_cast<uint32_t>(2 + BN / 64) * static_cast<uint32_t>(p.rows_per_kb) * 128u;

  // zero the operand stages once: pixel rows a box does not cover must read as zeros in every k-block
  for (int i = threadIdx.x; i < STAGES * STAGE_BYTES / 16; i += CV_THREADS)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tfull, 1);
    fence_mbar_init();
    prefetch_tensormap(&tmDY);
    prefetch_tensormap(&tmX);
  }
  if (warp == 1) { tmem_alloc(tmem_slot, BN < 32 ? 32 : BN); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = split; kb < p.total_kb; kb += p.ksplits) {
        const int txi = kb % p.tiles_x;
        const int q = kb / p.tiles_x;
        const int n = q / p.tiles_y;
        const int y0 = (q - n * p.tiles_y) * p.bh;
        const int x0 = txi * p.bw;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full_bar[stage], tx);
        uint8_t* sa = smem + stage * STAGE_BYTES;
#pragma unroll
        for (int a = 0; a < 2; ++a) tma_load_4d(sa + a * 8192, &tmDY, &full_bar[stage], cot * CV_BM + a * 64, x0, y0, n);
#pragma unroll
        for (int b = 0; b < BN / 64; ++b)
          tma_load_4d(sa + A_BYTES + b * 8192, &tmX, &full_bar[stage], cit * BN + b * 64, x0 + dx, y0 + dy, n);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(CV_BM, BN, true, true);
      int stage = 0; uint32_t phase = 0;
      bool first = true;
      for (int kb = split; kb < p.total_kb; kb += p.ksplits) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
        const uint32_t sb = sa + A_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_d, umma_smem_desc(sa + k * 2048, 8192, 1024), umma_smem_desc(sb + k * 2048, 8192, 1024), idesc,
                    (!first || k > 0) ? 1u : 0u);
        first = false;
        umma_commit(&empty_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(tfull);
    }
    __syncwarp();
  } else {
    const int quad = warp & 3;
    const int co = cot * CV_BM + quad * 32 + lane;
    const bool any = split < p.total_kb;              // this CTA had at least one k-block
    if (any) {
      mbar_wait(tfull, 0);
      tc_fence_after();
      float* drow = p.dw + static_cast<size_t>(co) * (9 * p.C) + tap * p.C + cit * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t acc[32];
        tmem_ld32(tmem_d + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(c * 32), acc);
        tmem_ld_wait();
        if (co < p.Cout && cit * BN + c * 32 < p.C) {
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(drow + c * 32 + i), "f"(__uint_as_float(acc[i])),
                         "f"(__uint_as_float(acc[i + 1])), "f"(__uint_as_float(acc[i + 2])), "f"(__uint_as_float(acc[i + 3]))
                         : "memory");
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_d, BN < 32 ? 32 : BN); }
}

template <int BN>
static int launch_wgrad(const CUtensorMap& tmDY, const CUtensorMap& tmX, const WgradParams& p, cudaStream_t st) {
  constexpr int STAGES = BN == 256 ? 2 : 3;
  constexpr int SMEM = STAGES * (2 * 8192 + (BN / 64) * 8192) + 128 + 1024;
  auto kern = conv3x3_wgrad_kernel<BN>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return set_error_cuda("cudaFuncSetAttribute(conv3x3_wgrad)", e);
    attr = true;
  }
  kern<<<p.co_tiles * 9 * p.ci_tiles * p.ksplits, CV_THREADS, SMEM, st>>>(tmDY, tmX, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error_cuda("conv3x3_wgrad launch", e);
  count_launch();
  return 0;
}

}  // namespace dc

using namespace dc;

extern "C" int dc_conv3x3_igemm_supported(int H, int W, int C, int Cout) {
  return (C % 32 == 0 && Cout % 32 == 0 && W >= 1 && W <= 128 && H >= 1) ? 1 : 0;
}

extern "C" int dc_conv3x3_igemm(const void* x, const void* w, void* out, int batch, int H, int W, int C, int Cout,
                                dc_stream_t stream) {
  if (!dc_conv3x3_igemm_supported(H, W, C, Cout)) return set_error("conv3x3_igemm: needs C % 32 == 0, Cout % 32 == 0, W <= 128");
  ConvParams p;
  memset(&p, 0, sizeof(p));
  p.B = batch; p.H = H; p.W = W; p.C = C; p.Cout = Cout;
  if (W * H <= CV_BM / 2) {              // small images: several whole images per tile
    p.bh = H; p.nb = CV_BM / (W * H); p.tiles_y = 1;
    p.n_ptiles = (batch + p.nb - 1) / p.nb;
  } else {
    p.bh = CV_BM / W; if (p.bh > H) p.bh = H;
    p.nb = 1;
    p.tiles_y = (H + p.bh - 1) / p.bh;
    p.n_ptiles = batch * p.tiles_y;
  }
  const int BN = Cout >= 256 ? 256 : (Cout >= 128 ? 128 : 64);
  p.n_ctiles = (Cout + BN - 1) / BN;
  p.cblocks = (C + CV_BK - 1) / CV_BK;
  p.kblocks = 9 * p.cblocks;
  p.out = static_cast<bf16*>(out);
  CUtensorMap tmX, tmW;
  const long long dims[4] = {C, W, H, batch};
  const long long strides[3] = {static_cast<long long>(C) * 2, static_cast<long long>(W) * C * 2,
                                static_cast<long long>(H) * W * C * 2};
  const int box[4] = {CV_BK, W, p.bh, p.nb};
  int rc = make_tmap_4d(&tmX, x, dims, strides, box);
  if (rc) return rc;
  rc = make_tmap_2d(&tmW, w, 9ll * C, Cout, 9ll * C, CV_BK, BN);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (BN == 256) return launch_conv<256>(tmX, tmW, p, st);
  if (BN == 128) return launch_conv<128>(tmX, tmW, p, st);
  return launch_conv<64>(tmX, tmW, p, st);
}

extern "C" int dc_conv3x3_wgrad_igemm_supported(int H, int W, int C, int Cout) {
  if (C % 32 != 0 || Cout % 32 != 0 || W < 1 || H < 1) return 0;
  const int tiles_x = (W + 63) / 64;
  return (W % tiles_x == 0 && W / tiles_x <= 64) ? 1 : 0;
}

extern "C" int dc_conv3x3_wgrad_igemm(const void* dy, const void* x, float* dw, int batch, int H, int W, int C, int Cout,
                                      dc_stream_t stream) {
  if (!dc_conv3x3_wgrad_igemm_supported(H, W, C, Cout))
    return set_error("conv3x3_wgrad_igemm: needs C % 32 == 0, Cout % 32 == 0, W <= 64 or W a multiple of ceil(W / 64)");
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.B = batch; p.H = H; p.W = W; p.C = C; p.Cout = Cout;
  p.tiles_x = (W + 63) / 64;
  p.bw = W / p.tiles_x;
  p.bh = 64 / p.bw; if (p.bh > H) p.bh = H; if (p.bh < 1) p.bh = 1;
  if (p.tiles_x > 1) p.bh = 1;
  p.tiles_y = (H + p.bh - 1) / p.bh;
  p.rows_per_kb = p.bw * p.bh;
  p.total_kb = batch * p.tiles_y * p.tiles_x;
  const int BN = C >= 256 ? 256 : (C >= 128 ? 128 : 64);
  p.co_tiles = (Cout + CV_BM - 1) / CV_BM;
  p.ci_tiles = (C + BN - 1) / BN;
  const int tiles = p.co_tiles * 9 * p.ci_tiles;
  int ks = (2 * sm_count() + tiles - 1) / tiles;        // ~two CTAs per SM
  if (ks < 1) ks = 1;
  if (ks > p.total_kb) ks = p.total_kb;
  p.ksplits = ks;
  p.dw = dw;
  CUtensorMap tmDY, tmX;
  const int box[4] = {64, p.bw, p.bh, 1};
  {
    const long long dims[4] = {Cout, W, H, batch};
    const long long strides[3] = {static_cast<long long>(Cout) * 2, static_cast<long long>(W) * Cout * 2,
                                  static_cast<long long>(H) * W * Cout * 2};
    int rc = make_tmap_4d(&tmDY, dy, dims, strides, box);
    if (rc) return rc;
  }
  {
    const long long dims[4] = {C, W, H, batch};
    const long long strides[3] = {static_cast<long long>(C) * 2, static_cast<long long>(W) * C * 2,
                                  static_cast<long long>(H) * W * C * 2};
    int rc = make_tmap_4d(&tmX, x, dims, strides, box);
    if (rc) return rc;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (BN == 256) return launch_wgrad<256>(tmDY, tmX, p, st);
  if (BN == 128) return launch_wgrad<128>(tmDY, tmX, p, st);
  return launch_wgrad<64>(tmDY, tmX, p, st);
}
