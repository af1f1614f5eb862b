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
// The in_proj bias gradient (column sums of dQ | dK | dV) accumulates in shared memory per head and is flushed once.
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
  int num_tiles;  // batch * heads / pp
  int D;          // heads * 64
  float* lse;     // forward: out (may be null), backward: in; [batch * heads, L], natural log
  bf16* out;      // forward [batch * L, D]
  bf16* dqkv;     // backward [batch * L, 3 D]
  float* dbias;   // backward, optional [3 D]
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

template <bool BWD>
struct AttnSmem {
  static constexpr int NIN = BWD ? 4 : 3;                        // Q, K, V (, dO)
  static constexpr uint32_t IN_BYTES = NIN * AT_TILE_BYTES;
  static constexpr uint32_t P_OFF = 2 * IN_BYTES;                // P: 2 blocks
  static constexpr uint32_t DS_OFF = P_OFF + 2 * AT_TILE_BYTES;  // dS: 2 blocks (backward)
  static constexpr uint32_t TAIL_OFF = P_OFF + (BWD ? 4 : 2) * AT_TILE_BYTES;
  static constexpr uint32_t XCHG_BYTES = 2 * 2 * 128 * 4;        // [which][half][row]
  static constexpr uint32_t DB_BYTES = BWD ? AT_MAX_HEADS * 192 * 4 : 0;
  static constexpr uint32_t BAR_OFF = TAIL_OFF + XCHG_BYTES + DB_BYTES;
  static constexpr uint32_t TOTAL = BAR_OFF + 16 * 8 + 1024;     // + alignment slack
};

template <bool BWD>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO, const AttnTcParams p) {
  using SM = AttnSmem<BWD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* s_in = smem;
  uint8_t* sP = smem + SM::P_OFF;
  uint8_t* sdS = smem + SM::DS_OFF;
  float* xchg = reinterpret_cast<float*>(smem + SM::TAIL_OFF);  // [2][2][128]
  float* s_db = xchg + 512;                                      // [heads][192] (backward)
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
    mbar_init(s_full, 1); mbar_init(p_ready, 8); mbar_init(o_full, 1);
    fence_mbar_init();
    prefetch_tensormap(&tmQKV);
    if (BWD) prefetch_tensormap(&tmDO);
  }
  // P / dS start as zeros: regions no warp ever writes (cross-pair blocks, fully masked causal chunks, rows >= L of
  // quadrants without work) must read as exact zeros in every tile
  for (uint32_t i = threadIdx.x; i < (BWD ? 4u : 2u) * AT_TILE_BYTES / 16; i += AT_THREADS)
    reinterpret_cast<uint4*>(sP)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (BWD)
    for (int i = threadIdx.x; i < p.heads * 192; i += AT_THREADS) s_db[i] = 0.f;
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

  // ---- control helpers, each executed by ONE elected lane (warp 1: TMA producer, warp 0: UMMA issuer)
  auto issue_loads = [&](int it, int tile) {
    const int buf = it & 1;
    mbar_arrive_expect_tx(&ld_full[buf], SM::IN_BYTES);
    uint8_t* dst = s_in + buf * SM::IN_BYTES;
    for (int slot = 0; slot < p.pp; ++slot) {
      const int pair = tile * p.pp + slot;
      const int b = pair / p.heads, h = pair - b * p.heads;
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
  auto issue_stage2 = [&](int it) {   // O = P V   |   dQ = dS K, dK = dS^T Q, dV = P^T dO
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, false, true);
    constexpr uint32_t idesc_t = umma_idesc_bf16(128, 64, true, true);
    const uint32_t aP = smem_u32(sP), adS = smem_u32(sdS);
    const uint32_t sQ = smem_u32(s_in + (it & 1) * SM::IN_BYTES);
    const uint32_t sK = sQ + AT_TILE_BYTES, sV = sK + AT_TILE_BYTES, sdO = sV + AT_TILE_BYTES;
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
#pragma unroll
      for (int s = 0; s < 8; ++s)
        umma_bf16(tdK, umma_smem_desc(adS + s * 2048, AT_TILE_BYTES, 1024),
                  umma_smem_desc(sQ + s * 2048, AT_TILE_BYTES, 1024), idesc_t, s > 0);
#pragma unroll
      for (int s = 0; s < 8; ++s)
        umma_bf16(tdV, umma_smem_desc(aP + s * 2048, AT_TILE_BYTES, 1024),
                  umma_smem_desc(sdO + s * 2048, AT_TILE_BYTES, 1024), idesc_t, s > 0);
    }
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

  auto tile_pair = [&](int tile, int& b, int& h) {
    const int pair = tile * p.pp + slot;
    b = pair / p.heads;
    h = pair - b * p.heads;
  };
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
  const int stride = gridDim.x;
  if (warp == 1) {
    if (lane == 0) {
      if (blockIdx.x < p.num_tiles) issue_loads(0, blockIdx.x);
      if (blockIdx.x + stride < p.num_tiles) issue_loads(1, blockIdx.x + stride);
    }
    __syncwarp();
  }
  if (warp == 0) {
    if (lane == 0 && blockIdx.x < p.num_tiles) issue_stage1(0);
    __syncwarp();
  }
  float lse_next = 0.f;
  if (BWD && blockIdx.x < p.num_tiles && row_valid) {
    int b, h;
    tile_pair(blockIdx.x, b, h);
    lse_next = p.lse[(static_cast<size_t>(b) * p.heads + h) * p.L + l];
  }

  int it = 0;
  for (int tile = blockIdx.x; tile < p.num_tiles; tile += stride, ++it) {
    int b, h;
    tile_pair(tile, b, h);
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
    // ---- UMMA issuer: second stage of this tile, then the first stage of the next one (its S / dP columns are free:
    // every warp arrived on p_ready, i.e. finished reading them, and finished the previous tile's epilogue)
    if (warp == 0) {
      if (lane == 0) {
        mbar_wait(p_ready, it & 1);
        tc_fence_after();
        issue_stage2(it);
        if (tile + stride < p.num_tiles) issue_stage1(it + 1);
      }
      __syncwarp();
    }
    if (BWD) {
      // prefetch the next tile's log-sum-exp while the second-stage UMMAs run
      const int nt = tile + stride;
      lse_next = 0.f;
      if (nt < p.num_tiles && row_valid) {
        int nb, nh;
        tile_pair(nt, nb, nh);
        lse_next = p.lse[(static_cast<size_t>(nb) * p.heads + nh) * p.L + l];
      }
    }
    mbar_wait(o_full, it & 1);
    tc_fence_after();
    // ---- TMA producer: this tile's input buffer is free again -> fetch the tile after next into it
    if (warp == 1) {
      if (lane == 0 && tile + 2 * stride < p.num_tiles) issue_loads(it + 2, tile + 2 * stride);
      __syncwarp();
    }
    if (!BWD) {
      // ---------------------------------------------------------------- forward epilogue
      uint32_t o[32];
      tmem_ld32(tO + lane_base + hf * 32, o);
      tmem_ld_wait();
      if (row_valid) {
        const float inv = 1.0f / fmaxf(tot, 1e-30f);
        bf16* dst = p.out + (static_cast<size_t>(b) * p.L + l) * p.D + h * 64 + hf * 32;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(o[8 * g]) * inv, __uint_as_float(o[8 * g + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(o[8 * g + 2]) * inv, __uint_as_float(o[8 * g + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(o[8 * g + 4]) * inv, __uint_as_float(o[8 * g + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(o[8 * g + 6]) * inv, __uint_as_float(o[8 * g + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + 8 * g) = u;
        }
        if (hf == 0 && p.lse != nullptr)
          p.lse[(static_cast<size_t>(b) * p.heads + h) * p.L + l] = (m + __log2f(fmaxf(tot, 1e-30f))) * kLn2;
      }
    } else {
      // ---------------------------------------------------------------- backward epilogue: dQ | dK | dV
      const size_t ld = static_cast<size_t>(3) * p.D;
      bf16* dst0 = p.dqkv + (static_cast<size_t>(b) * p.L + l) * ld + h * 64 + hf * 32;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        uint32_t r[32];
        tmem_ld32((a == 0 ? tdQ : (a == 1 ? tdK : tdV)) + lane_base + hf * 32, r);
        tmem_ld_wait();
        if (row_valid) {
          bf16* dst = dst0 + a * p.D;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(r[8 * g]), __uint_as_float(r[8 * g + 1]));
            u.y = pack_bf16x2(__uint_as_float(r[8 * g + 2]), __uint_as_float(r[8 * g + 3]));
            u.z = pack_bf16x2(__uint_as_float(r[8 * g + 4]), __uint_as_float(r[8 * g + 5]));
            u.w = pack_bf16x2(__uint_as_float(r[8 * g + 6]), __uint_as_float(r[8 * g + 7]));
            *reinterpret_cast<uint4*>(dst + 8 * g) = u;
          }
        }
        if (p.dbias != nullptr) {
          // rows >= L hold exact zeros (their P / dS rows and columns are zero), so no masking is needed
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          const float cs = chunk_colsum(v);
          atomicAdd(&s_db[h * 192 + a * 64 + hf * 32 + lane], cs);
        }
      }
    }
    tc_fence_before();   // orders this tile's tcgen05.ld before the p_ready arrive of the next tile
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
  if (BWD && p.dbias != nullptr) {
    for (int i = threadIdx.x; i < p.heads * 192; i += AT_THREADS) {
      const float v = s_db[i];
      const int h = i / 192, r = i - h * 192;
      if (v != 0.f) atomicAdd(&p.dbias[(r >> 6) * p.D + h * 64 + (r & 63)], v);
    }
  }
}

template <bool BWD>
static int launch_attn_tc(const void* qkv, const void* dout, AttnTcParams p, cudaStream_t st) {
  if (p.L <= 0 || p.L > 128 || p.heads > AT_MAX_HEADS) return DC_ATTN_TC_UNSUPPORTED;
  p.pp = p.L <= 64 ? 2 : 1;
  p.rp = 128 / p.pp;
  const long long pairs = static_cast<long long>(p.batch) * p.heads;
  if (pairs % p.pp != 0) return DC_ATTN_TC_UNSUPPORTED;
  p.num_tiles = static_cast<int>(pairs / p.pp);
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
  auto kern = attn_tc_kernel<BWD>;
  constexpr size_t smem = AttnSmem<BWD>::TOTAL;
  static bool set = false;
  if (!set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return set_error_cuda("cudaFuncSetAttribute(attn_tc)", e);
    set = true;
  }
  const int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
  kern<<<grid, AT_THREADS, smem, st>>>(tmQKV, tmDO, p);
  DC_CHECK_LAUNCH(BWD ? "attention_tc_bwd" : "attention_tc_fwd");
  return 0;
}

int attention_tc_fwd(const void* qkv, void* out, float* lse, int batch, int L, int heads, int causal, cudaStream_t st) {
  AttnTcParams p{};
  p.L = L; p.heads = heads; p.batch = batch; p.causal = causal;
  p.lse = lse; p.out = static_cast<bf16*>(out);
  return launch_attn_tc<false>(qkv, nullptr, p, st);
}

int attention_tc_bwd(const void* qkv, const void* dout, const float* lse, void* dqkv, float* dbias, int batch, int L,
                     int heads, int causal, cudaStream_t st) {
  AttnTcParams p{};
  p.L = L; p.heads = heads; p.batch = batch; p.causal = causal;
  p.lse = const_cast<float*>(lse); p.dqkv = static_cast<bf16*>(dqkv); p.dbias = dbias;
  return launch_attn_tc<true>(qkv, dout, p, st);
}

}  // namespace dc
