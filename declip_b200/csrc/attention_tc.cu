// declip_b200 — tcgen05 / TMEM attention core for the short sequences of the CLIP towers (L = 50 ViT-B/32 and the
// ResNet attention pool, L = 77 causal text; head_dim 64): softmax(q k^T / 8 [+ causal]) v, forward and backward
// (image_encoder/base_transformer.py:44-48 nn.MultiheadAttention, text_encoder/text_transformer.py:136-142 mask).
//
// One persistent CTA per SM walks 128-row tiles.  A tile is ONE (sample, head) pair when L > 64 (rows >= L are TMA
// zero fill) or TWO pairs when L <= 64 (rows 0-63 / 64-127; the cross-pair blocks of the 128 x 128 score tile are
// masked to zero, which makes every product below block-diagonal-correct).  All products are M = 128 UMMAs:
//
//   forward   S  = Q K^T                 (A = Q  K-major,  B = K  K-major,  N = 128, K = 64)
//             O  = P V                   (A = P  K-major,  B = V  MN-major, N = 64,  K = 128)
//   backward  S, dP = dO V^T             (as S)
//             dQ = dS K                  (A = dS K-major,  B = K  MN-major)
//             dK = dS^T Q                (A = dS MN-major, B = Q  MN-major)
//             dV = P^T dO                (A = P  MN-major, B = dO MN-major)
//
// P / dS live in shared memory once, as two [128 q][64 keys] 128B-swizzled blocks — the same bytes are a K-major A
// (K = keys) and an MN-major A (M = keys).  Eight warps (256 threads, so the softmax code may use up to 255 registers):
// two per TMEM lane quadrant, each owning 32-column chunks of its rows, one row per thread — row statistics need one
// 64-thread named-barrier exchange and no shuffles.  Lane 0 of warp 0 is also the UMMA issuer and lane 0 of warp 1 the
// TMA producer (Q, K, V, dO of the tile after next into the buffer the finished tile just released); program order of
// the eight warps (softmax -> p_ready -> o_full -> epilogue) makes separate "empty" barriers unnecessary.  delta = rowsum(dO * O) is computed as rowsum(P * dP), so O is never read.
// CTA c only ever sees head group c % (heads / pairs-per-tile), so the in_proj bias gradient needs no per-tile work:
// its Q slice is the column sum of a dQ that the tensor core keeps accumulating over the CTA's tiles in 64 spare TMEM
// columns, its K slice is identically zero (rows of dS sum to zero) and its V slice equals colsum(dO), which the
// caller gets for free from the epilogue of the GEMM that produces dO (or from this kernel with db_v = 1).
// Outputs leave through per-warp staged TMA stores (scattered 16-byte global stores made the LSU the bottleneck).
#include <stdlib.h>
#include "common.cuh"
#include "gemm_common.cuh"
#include "internal.h"

namespace dc {

constexpr int AT_THREADS = 256;
constexpr uint32_t AT_TILE_BYTES = 128 * 128;  // one [128 rows][64 bf16] operand tile
constexpr int AT_MAX_HEADS = 32;

struct AttnTcParams {
  int L, heads, batch, causal;
  int pp;         // pairs per tile (1 or 2)
  int rp;         // rows per pair slot = 128 / pp
  int groups;     // heads / pp: head groups; CTA c works on group c % groups only (one fixed head per tile slot)
  int per_group;  // CTAs per group; CTA c takes samples b = c / groups, + per_group, ...
  int db_v;       // backward: also produce the V slice of dbias in this kernel (external callers)
  int D;          // heads * 64
  float* lse;     // forward: out (may be null), backward: in; [batch * heads, L], natural log
  bf16* out;      // forward [batch * L, D]
  bf16* dqkv;     // backward [batch * L, 3 D]
  float* dbias;   // backward, optional [3 D]: Q slice always, V slice when db_v, K slice is identically zero (untouched)
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
// 32 consecutive bf16 of row `row` (columns [32 * chunk, +32)) of a [2 blocks][128 rows][64] 128B-swizzled operand.
__device__ __forceinline__ void store_chunk_sw128(uint8_t* base, int row, int chunk, const uint32_t (&pk)[16]) {
  uint8_t* rowp = base + (chunk >> 1) * AT_TILE_BYTES + row * 128;
  const int j0 = (chunk & 1) * 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int phys = (j0 + i) ^ (row & 7);
    *reinterpret_cast<uint4*>(rowp + phys * 16) = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
  }
}

// One warp's 32 rows x 32 bf16 columns leave through a 2 KiB staging buffer (64-byte swizzle, conflict-free 16-byte
// st.shared) and ONE 3-D TMA store {32 columns, 32 sequence positions, 1 sample}: positions >= L lie outside the tensor
// and are not written, the LSU sees no scattered 16-byte global stores.  Two buffers per warp: the store issued two
// chunks ago must have been read out, the previous one may still be in flight.
__device__ __forceinline__ void stage_store_rows(const uint32_t (&pk)[16], uint8_t* stage, uint32_t& sidx,
                                                 const CUtensorMap* tm, int col0, int l0, int b) {
  const int lane = lane_id();
  if (lane == 0) bulk_wait_read1();
  __syncwarp();
  stage += (sidx & 1u) * 2048;
  sidx ^= 1u;
  uint8_t* rowp = stage + lane * 64;
  const int sw = (lane >> 1) & 3;    // CU_TENSOR_MAP_SWIZZLE_64B: 16-byte chunk index ^= address bits [7,9)
#pragma unroll
  for (int j = 0; j < 4; ++j)
    *reinterpret_cast<uint4*>(rowp + ((j ^ sw) << 4)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    tma_store_3d(tm, stage, col0, l0, b);
    bulk_commit();
  }
}

template <bool BWD>
struct AttnSmem {
  static constexpr int NIN = BWD ? 4 : 3;                        // Q, K, V (, dO)
  static constexpr uint32_t IN_BYTES = NIN * AT_TILE_BYTES;
  static constexpr uint32_t P_OFF = 2 * IN_BYTES;                // P: 2 blocks
  static constexpr uint32_t DS_OFF = P_OFF + 2 * AT_TILE_BYTES;  // dS: 2 blocks (backward)
  static constexpr uint32_t STAGE_OFF = P_OFF + (BWD ? 4 : 2) * AT_TILE_BYTES;   // 8 warps x 2 x 2 KiB output staging
  static constexpr uint32_t TAIL_OFF = STAGE_OFF + 8 * 4096;
  static constexpr uint32_t XCHG_BYTES = (BWD ? 1 : 2) * 2 * 128 * 4;   // [which][half][row]; backward needs one
  static constexpr uint32_t DB_BYTES = BWD ? 2 * 64 * 4 : 0;     // V-slice bias gradient per tile slot
  static constexpr uint32_t BAR_OFF = TAIL_OFF + XCHG_BYTES + DB_BYTES;
  static constexpr uint32_t TOTAL = BAR_OFF + 16 * 8 + 1024;     // + alignment slack
};

template <bool BWD>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
               const __grid_constant__ CUtensorMap tmOUT, const AttnTcParams p) {
  using SM = AttnSmem<BWD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* s_in = smem;
  uint8_t* sP = smem + SM::P_OFF;
  uint8_t* sdS = smem + SM::DS_OFF;
  float* xchg = reinterpret_cast<float*>(smem + SM::TAIL_OFF);  // [2][2][128]
  float* s_db = xchg + 256;                                      // [2 slots][64] (backward, db_v)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFF);
  uint64_t* ld_full = bars;        // [2]  TMA bytes of one input buffer landed
  uint64_t* s_full = bars + 2;     //      first-stage UMMAs (S [, dP]) complete
  uint64_t* p_ready = bars + 3;    //      all 8 warps wrote P [, dS] and are done with S / dP and the previous outputs
  uint64_t* o_full = bars + 4;     //      second-stage UMMAs complete (also: input buffer and P / dS are free again)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t TMEM_COLS = BWD ? 512 : 256;

  if (threadIdx.x == 0) {
    mbar_init(&ld_full[0], 1); mbar_init(&ld_full[1], 1);
    mbar_init(s_full, 1); mbar_init(p_ready, 8); mbar_init(o_full, BWD ? 2 : 1);
    fence_mbar_init();
    prefetch_tensormap(&tmQKV);
    prefetch_tensormap(&tmOUT);
    if (BWD) prefetch_tensormap(&tmDO);
  }
  // P / dS start as zeros: regions no warp ever writes (cross-pair blocks, fully masked causal chunks, rows >= L of
  // quadrants without work) must read as exact zeros in every tile
  for (uint32_t i = threadIdx.x; i < (BWD ? 4u : 2u) * AT_TILE_BYTES / 16; i += AT_THREADS)
    reinterpret_cast<uint4*>(sP)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (BWD && threadIdx.x < 128) s_db[threadIdx.x] = 0.f;
  fence_proxy_async_smem();
  if (warp == 0) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tdP = tmem_base + 128;
  const uint32_t tO = tmem_base + 128;                                      // forward
  const uint32_t tdQ = tmem_base + 256, tdK = tmem_base + 320, tdV = tmem_base + 384;
  const uint32_t tdQs = tmem_base + 448;   // sum over this CTA's tiles of dQ (same head) -> Q slice of dbias
  const int grp = blockIdx.x % p.groups;   // head group of this CTA
  const int b0 = blockIdx.x / p.groups;    // first sample; then += per_group

  // ---- control helpers, each executed by ONE elected lane (warp 1: TMA producer, warp 0: UMMA issuer)
  auto issue_loads = [&](int it, int b) {
    const int buf = it & 1;
    mbar_arrive_expect_tx(&ld_full[buf], SM::IN_BYTES);
    uint8_t* dst = s_in + buf * SM::IN_BYTES;
    for (int slot = 0; slot < p.pp; ++slot) {
      const int h = grp * p.pp + slot;
      const uint32_t off = slot * p.rp * 128;
      tma_load_4d(dst + off, &tmQKV, &ld_full[buf], 0, h, 0, b);
      tma_load_4d(dst + AT_TILE_BYTES + off, &tmQKV, &ld_full[buf], 0, p.heads + h, 0, b);
      tma_load_4d(dst + 2 * AT_TILE_BYTES + off, &tmQKV, &ld_full[buf], 0, 2 * p.heads + h, 0, b);
      if (BWD) tma_load_4d(dst + 3 * AT_TILE_BYTES + off, &tmDO, &ld_full[buf], 0, h, 0, b);
    }
  };
  auto issue_stage1 = [&](int it) {   // S = Q K^T [, dP = dO V^T]
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, false, false);
    const int buf = it & 1;
    mbar_wait(&ld_full[buf], (it >> 1) & 1);
    tc_fence_after();
    const uint32_t sQ = smem_u32(s_in + buf * SM::IN_BYTES);
    const uint32_t sK = sQ + AT_TILE_BYTES, sV = sK + AT_TILE_BYTES, sdO = sV + AT_TILE_BYTES;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      umma_bf16(tS, umma_smem_desc(sQ + k * 32, 16, 1024), umma_smem_desc(sK + k * 32, 16, 1024), idesc_s, k > 0);
    if (BWD) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16(tdP, umma_smem_desc(sdO + k * 32, 16, 1024), umma_smem_desc(sV + k * 32, 16, 1024), idesc_s, k > 0);
    }
    umma_commit(s_full);
  };
  // second stage, part A (warp 0): O = P V   |   dQ = dS K and its running sum over this CTA's tiles
  auto issue_stage2a = [&](int it) {
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, false, true);
    const uint32_t aP = smem_u32(sP), adS = smem_u32(sdS);
    const uint32_t sQ = smem_u32(s_in + (it & 1) * SM::IN_BYTES);
    const uint32_t sK = sQ + AT_TILE_BYTES, sV = sK + AT_TILE_BYTES;
    if (!BWD) {
#pragma unroll
      for (int s = 0; s < 8; ++s)
        umma_bf16(tO, umma_smem_desc(aP + (s >> 2) * AT_TILE_BYTES + (s & 3) * 32, 16, 1024),
                  umma_smem_desc(sV + s * 2048, AT_TILE_BYTES, 1024), idesc_o, s > 0);
    } else {
#pragma unroll
      for (int s = 0; s < 8; ++s)
        umma_bf16(tdQ, umma_smem_desc(adS + (s >> 2) * AT_TILE_BYTES + (s & 3) * 32, 16, 1024),
                  umma_smem_desc(sK + s * 2048, AT_TILE_BYTES, 1024), idesc_o, s > 0);
      if (p.dbias != nullptr) {
#pragma unroll
        for (int s = 0; s < 8; ++s)
          umma_bf16(tdQs, umma_smem_desc(adS + (s >> 2) * AT_TILE_BYTES + (s & 3) * 32, 16, 1024),
                    umma_smem_desc(sK + s * 2048, AT_TILE_BYTES, 1024), idesc_o, (it > 0 || s > 0) ? 1u : 0u);
      }
    }
    umma_commit(o_full);
  };
  // second stage, part B (warp 2, backward only): dK = dS^T Q, dV = P^T dO
  auto issue_stage2b = [&](int it) {
    constexpr uint32_t idesc_t = umma_idesc_bf16(128, 64, true, true);
    const uint32_t aP = smem_u32(sP), adS = smem_u32(sdS);
    const uint32_t sQ = smem_u32(s_in + (it & 1) * SM::IN_BYTES);
    const uint32_t sdO = sQ + 3 * AT_TILE_BYTES;
#pragma unroll
    for (int s = 0; s < 8; ++s)
      umma_bf16(tdK, umma_smem_desc(adS + s * 2048, AT_TILE_BYTES, 1024),
                umma_smem_desc(sQ + s * 2048, AT_TILE_BYTES, 1024), idesc_t, s > 0);
#pragma unroll
    for (int s = 0; s < 8; ++s)
      umma_bf16(tdV, umma_smem_desc(aP + s * 2048, AT_TILE_BYTES, 1024),
                umma_smem_desc(sdO + s * 2048, AT_TILE_BYTES, 1024), idesc_t, s > 0);
    umma_commit(o_full);
  };

  const int q = warp & 3;     // TMEM lane quadrant of this warp
  const int hf = warp >> 2;   // which of the two warps of the quadrant
  const int row = q * 32 + lane;
  const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
  const int slot = (p.pp == 2) ? (row >> 6) : 0;
  const int l = (p.pp == 2) ? (row & 63) : row;
  const bool row_valid = l < p.L;
  // 32-column chunks of the score tile owned by this warp (warp-uniform)
  int nch = 0, ch0 = 0, ch1 = 0;
  if (p.pp == 2) {
    if ((q & 1) * 32 < p.L) { nch = 1; ch0 = slot * 2 + hf; }
  } else if (q * 32 < p.L) {
    const int need = (p.L + 31) >> 5;
    const int ntot = p.causal ? min(q + 1, need) : need;
    if (hf < ntot) { ch0 = hf; nch = 1; }
    if (hf + 2 < ntot) { ch1 = hf + 2; nch = 2; }
  }
  const float kScaleLog2 = 0.125f * 1.4426950408889634f;
  const float kLn2 = 0.6931471805599453f;
  float* xa = xchg;         // [2][128]
  float* xb = xchg + 256;   // [2][128]

  const int h = grp * p.pp + slot;   // this thread's head (fixed for the whole kernel)
  const int l0 = (p.pp == 2) ? (q & 1) * 32 : q * 32;   // sequence position of this warp's first row
  const bool warp_rows = l0 < p.L;                     // any valid row in this warp?
  uint8_t* wstage = smem + SM::STAGE_OFF + warp * 4096;
  uint32_t sidx = 0;
  // validity bit mask of the 32 columns of a chunk for this thread's row (bit j = column j participates)
  auto chunk_mask = [&](int c) -> uint32_t {
    if (!row_valid) return 0u;
    const int key0 = (p.pp == 2) ? (c & 1) * 32 : c * 32;
    int lim = p.L;
    if (p.causal) lim = min(lim, l + 1);
    const int n = lim - key0;                    // valid columns are j < n
    return n <= 0 ? 0u : (n >= 32 ? 0xffffffffu : ((1u << n) - 1u));
  };
  const uint32_t mask0 = nch > 0 ? chunk_mask(ch0) : 0u;
  const uint32_t mask1 = nch > 1 ? chunk_mask(ch1) : 0u;

  // ---- prologue: the first two tiles' loads, the first tile's first-stage UMMAs
  const int stride = p.per_group;
  if (warp == 1) {
    if (elect_one()) {
      if (b0 < p.batch) issue_loads(0, b0);
      if (b0 + stride < p.batch) issue_loads(1, b0 + stride);
    }
    __syncwarp();
  }
  if (warp == 0) {
    if (elect_one()) {
      if (b0 < p.batch) issue_stage1(0);
    }
    __syncwarp();
  }
  float lse_next = 0.f;
  if (BWD && b0 < p.batch && row_valid) lse_next = p.lse[(static_cast<size_t>(b0) * p.heads + h) * p.L + l];

  int it = 0;
  for (int b = b0; b < p.batch; b += stride, ++it) {
    mbar_wait(s_full, it & 1);
    tc_fence_after();
    float m = 0.f, tot = 0.f;   // forward row statistics
    if (!BWD) {
      // ---------------------------------------------------------------- forward softmax
      uint32_t raw0[32], raw1[32];
      if (nch > 0) tmem_ld32(tS + lane_base + ch0 * 32, raw0);
      if (nch > 1) tmem_ld32(tS + lane_base + ch1 * 32, raw1);
      tmem_ld_wait();
      float mx = -INFINITY;
      auto scale_mask = [&](uint32_t msk, uint32_t (&raw)[32]) {
        asm volatile("" : "+r"(msk));   // keep the per-column tests out of the loop-invariant hoister (register pressure)
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float v = ((msk >> j) & 1u) ? __uint_as_float(raw[j]) * kScaleLog2 : -INFINITY;
          raw[j] = __float_as_uint(v);
          mx = fmaxf(mx, v);
        }
      };
      if (nch > 0) scale_mask(mask0, raw0);
      if (nch > 1) scale_mask(mask1, raw1);
      xa[hf * 128 + row] = mx;
      named_bar_sync(1 + q, 64);
      m = fmaxf(xa[row], xa[128 + row]);
      if (m == -INFINITY) m = 0.f;
      float sum = 0.f;
      auto exp_store = [&](int c, const uint32_t (&raw)[32]) {
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float p0 = ex2_approx(__uint_as_float(raw[j]) - m);
          const float p1 = ex2_approx(__uint_as_float(raw[j + 1]) - m);
          sum += p0 + p1;
          pk[j >> 1] = pack_bf16x2(p0, p1);
        }
        store_chunk_sw128(sP, row, c, pk);
      };
      if (nch > 0) exp_store(ch0, raw0);
      if (nch > 1) exp_store(ch1, raw1);
      xb[hf * 128 + row] = sum;
      named_bar_sync(1 + q, 64);
      tot = xb[row] + xb[128 + row];
    } else {
      // ---------------------------------------------------------------- backward: P, delta, dS
      const float lse_l2 = lse_next * 1.4426950408889634f;
      uint32_t pk0[16], pk1[16];
      float dsum = 0.f;
      auto pass1 = [&](int c, uint32_t msk, uint32_t (&pk)[16]) {
        asm volatile("" : "+r"(msk));   // see scale_mask
        {
          uint32_t sr[32];
          tmem_ld32(tS + lane_base + c * 32, sr);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const bool ok0 = (msk >> j) & 1u, ok1 = (msk >> (j + 1)) & 1u;
            const float p0 = ok0 ? ex2_approx(fmaf(__uint_as_float(sr[j]), kScaleLog2, -lse_l2)) : 0.f;
            const float p1 = ok1 ? ex2_approx(fmaf(__uint_as_float(sr[j + 1]), kScaleLog2, -lse_l2)) : 0.f;
            pk[j >> 1] = pack_bf16x2(p0, p1);
          }
        }
        {
          // delta uses the bf16-rounded P — the same values the dV / dS products consume
          uint32_t dr[32];
          tmem_ld32(tdP + lane_base + c * 32, dr);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float2 pf = unpack_bf16x2(pk[j >> 1]);
            dsum = fmaf(pf.x, __uint_as_float(dr[j]), dsum);
            dsum = fmaf(pf.y, __uint_as_float(dr[j + 1]), dsum);
          }
        }
        store_chunk_sw128(sP, row, c, pk);
      };
      if (nch > 0) pass1(ch0, mask0, pk0);
      if (nch > 1) pass1(ch1, mask1, pk1);
      xa[hf * 128 + row] = dsum;
      named_bar_sync(1 + q, 64);
      const float delta = xa[row] + xa[128 + row];
      auto pass2 = [&](int c, const uint32_t (&pk)[16]) {
        uint32_t dr[32], dk[16];
        tmem_ld32(tdP + lane_base + c * 32, dr);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float2 pf = unpack_bf16x2(pk[j >> 1]);
          const float d0 = pf.x * (__uint_as_float(dr[j]) - delta) * 0.125f;
          const float d1 = pf.y * (__uint_as_float(dr[j + 1]) - delta) * 0.125f;
          dk[j >> 1] = pack_bf16x2(d0, d1);
        }
        store_chunk_sw128(sdS, row, c, dk);
      };
      if (nch > 0) pass2(ch0, pk0);
      if (nch > 1) pass2(ch1, pk1);
      // (xa is rewritten only after the next s_full, which is signalled after BOTH warps of the quadrant arrived on
      // p_ready below — no second exchange barrier is needed)
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(p_ready);
    // ---- UMMA issuers: second stage of this tile (split over two elected lanes in the backward), then the first stage
    // of the next tile (its S / dP columns are free: every warp arrived on p_ready, i.e. finished reading them, and
    // finished the previous tile's epilogue)
    if (warp == 0) {
      if (elect_one()) {
        mbar_wait(p_ready, it & 1);
        tc_fence_after();
        issue_stage2a(it);
        if (b + stride < p.batch) issue_stage1(it + 1);
      }
      __syncwarp();
    }
    if (BWD && warp == 2) {
      if (elect_one()) {
        mbar_wait(p_ready, it & 1);
        tc_fence_after();
        issue_stage2b(it);
      }
      __syncwarp();
    }
    if (BWD) {
      // prefetch the next tile's log-sum-exp while the second-stage UMMAs run
      const int nb = b + stride;
      lse_next = 0.f;
      if (nb < p.batch && row_valid) lse_next = p.lse[(static_cast<size_t>(nb) * p.heads + h) * p.L + l];
    }
    mbar_wait(o_full, it & 1);
    tc_fence_after();
    // ---- TMA producer: this tile's input buffer is free again -> fetch the tile after next into it
    if (warp == 1) {
      if (elect_one()) {
        if (b + 2 * stride < p.batch) issue_loads(it + 2, b + 2 * stride);
      }
      __syncwarp();
    }
    if (!BWD) {
      // ---------------------------------------------------------------- forward epilogue
      if (warp_rows) {
        uint32_t o[32];
        tmem_ld32(tO + lane_base + hf * 32, o);
        tmem_ld_wait();
        const float inv = 1.0f / fmaxf(tot, 1e-30f);
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j)
          pk[j] = pack_bf16x2(__uint_as_float(o[2 * j]) * inv, __uint_as_float(o[2 * j + 1]) * inv);
        stage_store_rows(pk, wstage, sidx, &tmOUT, h * 64 + hf * 32, l0, b);
        if (row_valid && hf == 0 && p.lse != nullptr)
          p.lse[(static_cast<size_t>(b) * p.heads + h) * p.L + l] = (m + __log2f(fmaxf(tot, 1e-30f))) * kLn2;
      }
    } else {
      // ---------------------------------------------------------------- backward epilogue: dQ | dK | dV
      if (warp_rows) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        uint32_t r[32];
        tmem_ld32((a == 0 ? tdQ : (a == 1 ? tdK : tdV)) + lane_base + hf * 32, r);
        tmem_ld_wait();
        {
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
          stage_store_rows(pk, wstage, sidx, &tmOUT, a * p.D + h * 64 + hf * 32, l0, b);
        }
        if (a == 2 && p.dbias != nullptr && p.db_v) {
          // rows >= L hold exact zeros (their P / dS rows and columns are zero), so no masking is needed
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          const float cs = chunk_colsum(v);
          atomicAdd(&s_db[slot * 64 + hf * 32 + lane], cs);
        }
      }
      }
    }
    tc_fence_before();   // orders this tile's tcgen05.ld before the p_ready arrive of the next tile
  }
  if (BWD && p.dbias != nullptr && it > 0 && warp_rows) {
    // Q slice of the in_proj bias gradient: column sums of the dQ accumulated over this CTA's tiles (all of head h);
    // the last o_full wait above covers the accumulating UMMAs.  The K slice is identically zero (rows of dS sum to
    // zero), the V slice is colsum(dO) — produced here only on request (db_v), else by the caller.
    uint32_t r[32];
    tmem_ld32(tdQs + lane_base + hf * 32, r);
    tmem_ld_wait();
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    const float cs = chunk_colsum(v);
    atomicAdd(&p.dbias[h * 64 + hf * 32 + lane], cs);
  }
  if (lane == 0) bulk_wait_read0();   // the staging buffers must stay valid until the last stores have read them
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
  if (BWD && p.dbias != nullptr && p.db_v && threadIdx.x < 64 * p.pp) {
    const int sl = threadIdx.x >> 6;
    atomicAdd(&p.dbias[2 * p.D + (grp * p.pp + sl) * 64 + (threadIdx.x & 63)], s_db[threadIdx.x]);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Forward, two CTAs per SM [r2] — the default forward (ViT b = 512: 42.7 -> 31.0 us under ncu, text 59 -> 45 us with CUDA
// events; profiles/r02_attention_fwd2.md).  The kernel above keeps ONE tile in flight per CTA: load -> S -> softmax -> P V
// -> epilogue is a serial chain (2.3 us per ViT tile; ncu: 21 % of the warp samples sit in the o_full wait, the rest is
// spread thinly over the chain, tensor pipe 16 % busy, DRAM at half of its peak) and its 160 KiB of shared memory
// (double-buffered Q / K / V) admit one CTA per SM.  This variant is the same tile algorithm with a footprint of 99 KiB
// and <= 128 registers, so that TWO CTAs are resident per SM and one CTA's softmax / epilogue runs under the other's UMMA
// and TMA phases:
//   * Q, K, V are single-buffered, but each is re-filled the moment its last reader finished: Q and K of the next tile are
//     requested right after `s_full` (the S UMMAs have read them), V right after `o_full` — separate mbarriers `ld_qk` /
//     `ld_v`, so a load is in flight during the whole softmax + P V + epilogue span of the current tile;
//   * the score chunk is read from TMEM twice (row maximum, then exponentials) instead of being parked in 64 registers;
//   * one 2 KiB output staging buffer per warp (a warp stores one 32 x 32 chunk per tile);
//   * the next tile's S UMMAs are issued right after this tile's P V if its Q / K have landed (non-blocking
//     `mbarrier.test_wait`), else after the epilogue — the issuing warp never parks in front of its own epilogue.
struct Attn2Smem {
  static constexpr uint32_t P_OFF = 3 * AT_TILE_BYTES;                  // Q | K | V | P (2 blocks)
  static constexpr uint32_t STAGE_OFF = P_OFF + 2 * AT_TILE_BYTES;      // 8 warps x 2 KiB output staging
  static constexpr uint32_t TAIL_OFF = STAGE_OFF + 8 * 2048;
  static constexpr uint32_t XCHG_BYTES = 2 * 2 * 128 * 4;               // [which][half][row]
  static constexpr uint32_t BAR_OFF = TAIL_OFF + XCHG_BYTES;
  static constexpr uint32_t TOTAL = BAR_OFF + 16 * 8 + 1024;            // + alignment slack
};
static_assert(2 * (Attn2Smem::TOTAL + 1024) <= 233472, "two forward CTAs must fit one SM's shared memory");

__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {   // non-blocking poll
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

__global__ void __launch_bounds__(AT_THREADS, 2)
attn_tc_fwd2_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmOUT, const AttnTcParams p) {
  using SM = Attn2Smem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + AT_TILE_BYTES;
  uint8_t* sV = smem + 2 * AT_TILE_BYTES;
  uint8_t* sP = smem + SM::P_OFF;
  float* xa = reinterpret_cast<float*>(smem + SM::TAIL_OFF);   // [2][128] row maxima of the two warps of a quadrant
  float* xb = xa + 256;                                         // [2][128] row sums
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFF);
  uint64_t* ld_qk = bars;          // Q and K of a tile landed
  uint64_t* ld_v = bars + 1;       // V of a tile landed
  uint64_t* s_full = bars + 2;     // S = Q K^T complete (Q, K buffers free)
  uint64_t* p_ready = bars + 3;    // all 8 warps wrote P and are done with S and the previous O
  uint64_t* o_full = bars + 4;     // O = P V complete (V buffer and P free)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t TMEM_COLS = 256;
  if (threadIdx.x == 0) {
    mbar_init(ld_qk, 1); mbar_init(ld_v, 1);
    mbar_init(s_full, 1); mbar_init(p_ready, 8); mbar_init(o_full, 1);
    fence_mbar_init();
    prefetch_tensormap(&tmQKV);
    prefetch_tensormap(&tmOUT);
  }
  // P starts as zeros: regions no warp ever writes (cross-pair blocks, fully masked causal chunks, rows >= L of quadrants
  // without work) must read as exact zeros in every tile
  for (uint32_t i = threadIdx.x; i < 2u * AT_TILE_BYTES / 16; i += AT_THREADS)
    reinterpret_cast<uint4*>(sP)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  if (warp == 0) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tO = tmem_base + 128;
  const int grp = blockIdx.x % p.groups;   // head group of this CTA
  const int b0 = blockIdx.x / p.groups;    // first sample; then += per_group
  const int stride = p.per_group;

  auto issue_qk = [&](int b) {
    mbar_arrive_expect_tx(ld_qk, 2 * AT_TILE_BYTES);
    for (int slot = 0; slot < p.pp; ++slot) {
      const int h = grp * p.pp + slot;
      const uint32_t off = slot * p.rp * 128;
      tma_load_4d(sQ + off, &tmQKV, ld_qk, 0, h, 0, b);
      tma_load_4d(sK + off, &tmQKV, ld_qk, 0, p.heads + h, 0, b);
    }
  };
  auto issue_v = [&](int b) {
    mbar_arrive_expect_tx(ld_v, AT_TILE_BYTES);
    for (int slot = 0; slot < p.pp; ++slot)
      tma_load_4d(sV + slot * p.rp * 128, &tmQKV, ld_v, 0, 2 * p.heads + grp * p.pp + slot, 0, b);
  };
  auto issue_s = [&]() {     // S = Q K^T
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, false, false);
    const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      umma_bf16(tS, umma_smem_desc(aQ + k * 32, 16, 1024), umma_smem_desc(aK + k * 32, 16, 1024), idesc_s, k > 0);
    umma_commit(s_full);
  };
  auto issue_o = [&]() {     // O = P V
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, false, true);
    const uint32_t aP = smem_u32(sP), aV = smem_u32(sV);
#pragma unroll
    for (int s = 0; s < 8; ++s)
      umma_bf16(tO, umma_smem_desc(aP + (s >> 2) * AT_TILE_BYTES + (s & 3) * 32, 16, 1024),
                umma_smem_desc(aV + s * 2048, AT_TILE_BYTES, 1024), idesc_o, s > 0);
    umma_commit(o_full);
  };

  const int q = warp & 3;     // TMEM lane quadrant of this warp
  const int hf = warp >> 2;   // which of the two warps of the quadrant
  const int row = q * 32 + lane;
  const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
  const int slot = (p.pp == 2) ? (row >> 6) : 0;
  const int l = (p.pp == 2) ? (row & 63) : row;
  const bool row_valid = l < p.L;
  int nch = 0, ch0 = 0, ch1 = 0;     // 32-column chunks of the score tile owned by this warp (warp-uniform)
  if (p.pp == 2) {
    if ((q & 1) * 32 < p.L) { nch = 1; ch0 = slot * 2 + hf; }
  } else if (q * 32 < p.L) {
    const int need = (p.L + 31) >> 5;
    const int ntot = p.causal ? min(q + 1, need) : need;
    if (hf < ntot) { ch0 = hf; nch = 1; }
    if (hf + 2 < ntot) { ch1 = hf + 2; nch = 2; }
  }
  const float kScaleLog2 = 0.125f * 1.4426950408889634f;
  const float kLn2 = 0.6931471805599453f;
  const int h = grp * p.pp + slot;   // this thread's head (fixed for the whole kernel)
  const int l0 = (p.pp == 2) ? (q & 1) * 32 : q * 32;   // sequence position of this warp's first row
  const bool warp_rows = l0 < p.L;
  uint8_t* wstage = smem + SM::STAGE_OFF + warp * 2048;
  auto chunk_mask = [&](int c) -> uint32_t {   // bit j = column j of chunk c participates in this thread's row
    if (!row_valid) return 0u;
    const int key0 = (p.pp == 2) ? (c & 1) * 32 : c * 32;
    int lim = p.L;
    if (p.causal) lim = min(lim, l + 1);
    const int n = lim - key0;
    return n <= 0 ? 0u : (n >= 32 ? 0xffffffffu : ((1u << n) - 1u));
  };
  const uint32_t mask0 = nch > 0 ? chunk_mask(ch0) : 0u;
  const uint32_t mask1 = nch > 1 ? chunk_mask(ch1) : 0u;

  if (warp == 1) {
    if (elect_one()) {
      if (b0 < p.batch) { issue_qk(b0); issue_v(b0); }
    }
    __syncwarp();
  }
  if (warp == 0) {
    if (elect_one()) {
      if (b0 < p.batch) {
        mbar_wait(ld_qk, 0);
        tc_fence_after();
        issue_s();
      }
    }
    __syncwarp();
  }

  int it = 0;
  for (int b = b0; b < p.batch; b += stride, ++it) {
    const bool has_next = b + stride < p.batch;
    const uint32_t par = it & 1;
    mbar_wait(s_full, par);
    tc_fence_after();
    if (warp == 1) {           // Q / K are free: fetch the next tile's
      if (elect_one()) {
        if (has_next) issue_qk(b + stride);
      }
      __syncwarp();
    }
    // ---------------------------------------------------------------- softmax: row maximum, then exponentials -> P
    float mx = -INFINITY;
    auto max_chunk = [&](int c, uint32_t msk) {
      uint32_t raw[32];
      tmem_ld32(tS + lane_base + c * 32, raw);
      tmem_ld_wait();
      asm volatile("" : "+r"(msk));   // keep the per-column tests out of the loop-invariant hoister (register pressure)
#pragma unroll
      for (int j = 0; j < 32; ++j) mx = fmaxf(mx, ((msk >> j) & 1u) ? __uint_as_float(raw[j]) : -INFINITY);
    };
    if (nch > 0) max_chunk(ch0, mask0);
    if (nch > 1) max_chunk(ch1, mask1);
    xa[hf * 128 + row] = mx;
    named_bar_sync(1 + q, 64);
    float m = fmaxf(xa[row], xa[128 + row]) * kScaleLog2;   // the scale is positive: max commutes with it
    if (!(m > -INFINITY)) m = 0.f;
    float sum = 0.f;
    auto exp_chunk = [&](int c, uint32_t msk) {
      uint32_t raw[32], pk[16];
      tmem_ld32(tS + lane_base + c * 32, raw);
      tmem_ld_wait();
      asm volatile("" : "+r"(msk));
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float p0 = ((msk >> j) & 1u) ? ex2_approx(fmaf(__uint_as_float(raw[j]), kScaleLog2, -m)) : 0.f;
        const float p1 = ((msk >> (j + 1)) & 1u) ? ex2_approx(fmaf(__uint_as_float(raw[j + 1]), kScaleLog2, -m)) : 0.f;
        sum += p0 + p1;
        pk[j >> 1] = pack_bf16x2(p0, p1);
      }
      store_chunk_sw128(sP, row, c, pk);
    };
    if (nch > 0) exp_chunk(ch0, mask0);
    if (nch > 1) exp_chunk(ch1, mask1);
    xb[hf * 128 + row] = sum;
    named_bar_sync(1 + q, 64);
    const float tot = xb[row] + xb[128 + row];
    fence_proxy_async_smem();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(p_ready);
    // ---- UMMA issuer: O = P V; then the next tile's S if its Q / K are already here
    bool s_issued = false;
    if (warp == 0) {
      if (elect_one()) {
        mbar_wait(p_ready, par);
        mbar_wait(ld_v, par);
        tc_fence_after();
        issue_o();
        if (has_next && mbar_test_wait(ld_qk, par ^ 1u)) {
          tc_fence_after();
          issue_s();
          s_issued = true;
        }
      }
      __syncwarp();
    }
    mbar_wait(o_full, par);
    tc_fence_after();
    if (warp == 1) {           // V is free: fetch the next tile's
      if (elect_one()) {
        if (has_next) issue_v(b + stride);
      }
      __syncwarp();
    }
    // ---------------------------------------------------------------- epilogue: O / rowsum -> out, log-sum-exp
    if (warp_rows) {
      uint32_t o[32];
      tmem_ld32(tO + lane_base + hf * 32, o);
      tmem_ld_wait();
      const float inv = 1.0f / fmaxf(tot, 1e-30f);
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 16; ++j)
        pk[j] = pack_bf16x2(__uint_as_float(o[2 * j]) * inv, __uint_as_float(o[2 * j + 1]) * inv);
      if (lane == 0) bulk_wait_read0();      // this warp's previous store has read the staging buffer
      __syncwarp();
      uint8_t* rowp = wstage + lane * 64;
      const int sw = (lane >> 1) & 3;        // CU_TENSOR_MAP_SWIZZLE_64B
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(rowp + ((j ^ sw) << 4)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_3d(&tmOUT, wstage, h * 64 + hf * 32, l0, b);
        bulk_commit();
      }
      if (row_valid && hf == 0 && p.lse != nullptr)
        p.lse[(static_cast<size_t>(b) * p.heads + h) * p.L + l] = (m + __log2f(fmaxf(tot, 1e-30f))) * kLn2;
    }
    tc_fence_before();   // orders this tile's tcgen05.ld before the p_ready arrive of the next tile
    if (warp == 0) {     // the next tile's S, if its operands were still in flight above
      if (elect_one()) {
        if (has_next && !s_issued) {
          mbar_wait(ld_qk, par ^ 1u);
          tc_fence_after();
          issue_s();
        }
      }
      __syncwarp();
    }
  }
  if (lane == 0) bulk_wait_read0();   // the staging buffers must stay valid until the last stores have read them
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <bool BWD>
static int launch_attn_tc(const void* qkv, const void* dout, AttnTcParams p, cudaStream_t st) {
  if (p.L <= 0 || p.L > 128 || p.heads > AT_MAX_HEADS) return DC_ATTN_TC_UNSUPPORTED;
  p.pp = p.L <= 64 ? 2 : 1;
  p.rp = 128 / p.pp;
  if (p.heads % p.pp != 0) return DC_ATTN_TC_UNSUPPORTED;
  p.groups = p.heads / p.pp;
  if (p.groups > sm_count()) return DC_ATTN_TC_UNSUPPORTED;
  p.per_group = sm_count() / p.groups;
  if (p.per_group > p.batch) p.per_group = p.batch;
  p.D = p.heads * 64;
  CUtensorMap tmQKV, tmDO;
  {
    const long long dims[4] = {64, 3LL * p.heads, p.L, p.batch};
    const long long strides[3] = {128, 3LL * p.D * 2, static_cast<long long>(p.L) * 3 * p.D * 2};
    const int box[4] = {64, 1, p.rp, 1};
    int rc = make_tmap_4d(&tmQKV, qkv, dims, strides, box);
    if (rc) return rc;
  }
  if (BWD) {
    const long long dims[4] = {64, p.heads, p.L, p.batch};
    const long long strides[3] = {128, 1LL * p.D * 2, static_cast<long long>(p.L) * p.D * 2};
    const int box[4] = {64, 1, p.rp, 1};
    int rc = make_tmap_4d(&tmDO, dout, dims, strides, box);
    if (rc) return rc;
  } else {
    tmDO = tmQKV;
  }
  CUtensorMap tmOUT;
  {
    // forward: out [batch * L, D]; backward: dqkv [batch * L, 3 D], as {columns, position, sample} so that a 32-row box
    // starting inside a sample never spills into the next one
    const long long ncol = (BWD ? 3LL : 1LL) * p.D;
    const long long dims[3] = {ncol, p.L, p.batch};
    const long long strides[2] = {ncol * 2, static_cast<long long>(p.L) * ncol * 2};
    const int box[3] = {32, 32, 1};
    int rc = make_tmap_nd(&tmOUT, BWD ? static_cast<const void*>(p.dqkv) : static_cast<const void*>(p.out), 3, dims,
                          strides, box, 64);
    if (rc) return rc;
  }
  auto kern = attn_tc_kernel<BWD>;
  constexpr size_t smem = AttnSmem<BWD>::TOTAL;
  static bool set = false;
  if (!set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return set_error_cuda("cudaFuncSetAttribute(attn_tc)", e);
    set = true;
  }
  const int grid = p.groups * p.per_group;
  kern<<<grid, AT_THREADS, smem, st>>>(tmQKV, tmDO, tmOUT, p);
  DC_CHECK_LAUNCH(BWD ? "attention_tc_bwd" : "attention_tc_fwd");
  return 0;
}

// Forward through the two-CTAs-per-SM kernel (the default; DC_ATTN_FWD_V1=1 selects the one-tile-per-SM kernel above).
static int launch_attn_fwd2(const void* qkv, AttnTcParams p, cudaStream_t st) {
  if (p.L <= 0 || p.L > 128 || p.heads > AT_MAX_HEADS) return DC_ATTN_TC_UNSUPPORTED;
  p.pp = p.L <= 64 ? 2 : 1;
  p.rp = 128 / p.pp;
  if (p.heads % p.pp != 0) return DC_ATTN_TC_UNSUPPORTED;
  p.groups = p.heads / p.pp;
  if (p.groups > sm_count()) return DC_ATTN_TC_UNSUPPORTED;
  p.D = p.heads * 64;
  auto kern = attn_tc_fwd2_kernel;
  constexpr size_t smem = Attn2Smem::TOTAL;
  static bool set = false;
  if (!set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return set_error_cuda("cudaFuncSetAttribute(attn_tc_fwd2)", e);
    // two CTAs only fit with the full shared-memory carve-out (with the default preference the occupancy query — sized for
    // one block — answered 1 and the first version of this launcher ran one CTA per SM)
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return set_error_cuda("cudaFuncSetAttribute(attn_tc_fwd2 carveout)", e);
    set = true;
  }
  // 2 CTAs per SM by construction: __launch_bounds__(256, 2) caps the registers, the static_assert above the shared memory,
  // 2 x 256 TMEM columns; if a device admitted only one, the second half of the grid would simply run as a second wave
  p.per_group = 2 * sm_count() / p.groups;
  if (p.per_group > p.batch) p.per_group = p.batch;
  CUtensorMap tmQKV, tmOUT;
  {
    const long long dims[4] = {64, 3LL * p.heads, p.L, p.batch};
    const long long strides[3] = {128, 3LL * p.D * 2, static_cast<long long>(p.L) * 3 * p.D * 2};
    const int box[4] = {64, 1, p.rp, 1};
    int rc = make_tmap_4d(&tmQKV, qkv, dims, strides, box);
    if (rc) return rc;
  }
  {
    const long long dims[3] = {p.D, p.L, p.batch};
    const long long strides[2] = {static_cast<long long>(p.D) * 2, static_cast<long long>(p.L) * p.D * 2};
    const int box[3] = {32, 32, 1};
    int rc = make_tmap_nd(&tmOUT, p.out, 3, dims, strides, box, 64);
    if (rc) return rc;
  }
  kern<<<p.groups * p.per_group, AT_THREADS, smem, st>>>(tmQKV, tmOUT, p);
  DC_CHECK_LAUNCH("attention_tc_fwd2");
  return 0;
}

int attention_tc_fwd(const void* qkv, void* out, float* lse, int batch, int L, int heads, int causal, cudaStream_t st) {
  AttnTcParams p{};
  p.L = L; p.heads = heads; p.batch = batch; p.causal = causal;
  p.lse = lse; p.out = static_cast<bf16*>(out);
  static const bool v1 = [] { const char* e = getenv("DC_ATTN_FWD_V1"); return e != nullptr && e[0] == '1'; }();
  if (v1) return launch_attn_tc<false>(qkv, nullptr, p, st);      // the one-tile-per-SM forward (A/B, fallback)
  return launch_attn_fwd2(qkv, p, st);
}

bool attention_tc_supported(int batch, int L, int heads) {
  if (batch <= 0 || L <= 0 || L > 128 || heads > AT_MAX_HEADS) return false;
  const int pp = L <= 64 ? 2 : 1;
  return heads % pp == 0 && heads / pp <= sm_count();
}

int attention_tc_bwd(const void* qkv, const void* dout, const float* lse, void* dqkv, float* dbias, int dbias_v,
                     int batch, int L, int heads, int causal, cudaStream_t st) {
  AttnTcParams p{};
  p.db_v = dbias_v;
  p.L = L; p.heads = heads; p.batch = batch; p.causal = causal;
  p.lse = const_cast<float*>(lse); p.dqkv = static_cast<bf16*>(dqkv); p.dbias = dbias;
  return launch_attn_tc<true>(qkv, dout, p, st);
}

}  // namespace dc
