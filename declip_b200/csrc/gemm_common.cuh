// declip_b200 — pieces shared by the 1-CTA (gemm.cu) and 2-CTA (gemm2.cu) tcgen05 GEMM kernels: kernel parameters
// and the fused epilogue (TMEM -> registers -> math -> staged TMA store / direct fp32 store).
#pragma once
#include "common.cuh"
#include "internal.h"

namespace dc {

struct GemmKParams {
  int M, N, K;
  int num_m, num_n, splits, kb_per_split, total_kb, num_tiles;
  int epi;
  float alpha;
  void* out;
  int ldo;
  void* out2;
  int ldo2;
  const float* bias;
  const bf16* aux;
  int ldaux;
  const float* alpha_dev;
  float* colsum;
};

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int GEMM_THREADS = 384;  // 4 control warps + 8 epilogue warps
constexpr int EPI_WARPS = 8;

// ---------------------------------------------------------------------------------------------- epilogue
// One epilogue warp owns 32 accumulator rows (its TMEM lane quadrant) x BN/2 columns of a tile, processed in
// 32-column chunks.  Everything the math needs from global memory is fetched BEFORE it is needed: the bias slice
// is loaded to registers before the warp blocks on the accumulator barrier and broadcast by warp shuffles;
// the aux operand (residual / pre-activation) of chunk c+1 is in flight while chunk c is processed.
// bf16 outputs leave through two per-warp 2 KiB staging buffers (64-byte swizzle, conflict-free st.shared.v4) and one
// TMA store per 32x32 chunk: fully coalesced 64-byte row segments, OOB rows/cols clipped by the TMA unit, and the
// LSU is free for the next chunk.  fp32 outputs (logit strips, features, split-K wgrad atomics) are written
// directly.  The epilogue mode is a compile-time parameter (one branch per tile).
template <int BUFS>
__device__ __forceinline__ void stage_store_chunk(const float (&v)[32], uint8_t* stage, uint32_t& sidx,
                                                  const CUtensorMap* tm, int col0, int row0) {
  const int lane = lane_id();
  // BUFS rotating 2 KiB staging buffers per warp: the store issued BUFS chunks ago (same buffer) must have been read
  // out, the BUFS - 1 younger ones may still be in flight
  if (lane == 0) bulk_wait_read<BUFS - 1>();
  __syncwarp();
  stage += (sidx & (BUFS - 1)) * 2048;
  sidx = (sidx + 1) & (BUFS - 1);
  uint8_t* rowp = stage + lane * 64;
  const int sw = (lane >> 1) & 3;    // CU_TENSOR_MAP_SWIZZLE_64B: 16-byte chunk index ^= address bits [7,9)
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 w;
    w.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]); w.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
    w.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]); w.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
    *reinterpret_cast<uint4*>(rowp + ((j ^ sw) << 4)) = w;
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    tma_store_2d(tm, stage, col0, row0);
    bulk_commit();
  }
}

// Residual / pre-activation operand of a 32 x 32 chunk.  One row per lane is how the accumulator arrives, but reading the
// bf16 aux operand that way (16 bytes per lane at a row stride) costs 32 LSU wavefronts per instruction, 128 per chunk: the
// +residual and xGELU' epilogues ran 30-50 % behind the plain one (out_proj+residual 763 vs 1167 TFLOP/s).  Instead the
// warp fetches the chunk COALESCED — four lanes per 64-byte row segment, eight rows per instruction — one chunk ahead
// into registers, then transposes it through the staging buffer the output of this chunk will use next
// (64-byte swizzle: conflict-free on both sides), so that every lane ends up with its own row again.
__device__ __forceinline__ void aux_fetch(const GemmKParams& p, int row0, int col0, uint4 (&a)[4]) {
  const int lane = lane_id();
  const int col = col0 + (lane & 3) * 8;
#pragma unroll
  for (int r8 = 0; r8 < 4; ++r8) {
    const int row = row0 + (lane >> 2) + 8 * r8;
    a[r8] = make_uint4(0u, 0u, 0u, 0u);
    if (row < p.M && col < p.N) a[r8] = *reinterpret_cast<const uint4*>(p.aux + static_cast<size_t>(row) * p.ldaux + col);
  }
}
template <int BUFS>
__device__ __forceinline__ void aux_transpose(uint4 (&a)[4], uint8_t* stage, uint32_t sidx) {
  const int lane = lane_id();
  if (lane == 0) bulk_wait_read<BUFS - 1>();        // the TMA store issued BUFS chunks ago from this buffer has been read out
  __syncwarp();
  uint8_t* buf = stage + (sidx & (BUFS - 1)) * 2048;
#pragma unroll
  for (int r8 = 0; r8 < 4; ++r8) {
    const int rr = (lane >> 2) + 8 * r8;
    *reinterpret_cast<uint4*>(buf + rr * 64 + (((lane & 3) ^ ((rr >> 1) & 3)) << 4)) = a[r8];
  }
  __syncwarp();
  const int sw = (lane >> 1) & 3;
#pragma unroll
  for (int j = 0; j < 4; ++j) a[j] = *reinterpret_cast<const uint4*>(buf + lane * 64 + ((j ^ sw) << 4));
  __syncwarp();                              // all rows read before any lane's output overwrites the buffer
}

// Column sums of a 32x32 chunk held one row per lane: butterfly reduce-scatter (31 shuffles); lane j ends with
// the sum of column j.
__device__ __forceinline__ float chunk_colsum(float (&v)[32]) {
  const int lane = lane_id();
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = upper ? v[i] : v[i + off];
      const float keep = upper ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// Waits for the accumulator, then drains this warp's 32 rows x (NCH * 32) columns starting at column `colbase`.
template <int EPI, int NCH, int BUFS = 2>
__device__ __forceinline__ void epilogue_tile(const GemmKParams& p, const CUtensorMap* tm_out, const CUtensorMap* tm_out2,
                                              float alpha, uint32_t taddr, int row0, int colbase, uint32_t& sidx,
                                              uint8_t* stage, uint64_t* tfull, uint32_t parity) {
  constexpr bool HAS_AUX = (EPI == DC_EPI_BF16_RESID || EPI == DC_EPI_BF16_DGELU);
  constexpr bool HAS_BIAS = (EPI != DC_EPI_F32_ATOMIC && EPI != DC_EPI_BF16_DGELU && EPI != DC_EPI_F32_GROUPMAX16);
  constexpr bool GMAX = (EPI == DC_EPI_F32_GROUPMAX16);
  float gbest[GMAX ? 2 * NCH : 1];
  int garg[GMAX ? 2 * NCH : 1];
  constexpr bool OUT_BF16 = (EPI <= DC_EPI_BF16_DGELU);
  const int lane = lane_id();
  const int row = row0 + lane;
  const bool row_ok = row < p.M;
  const bool use_bias = HAS_BIAS && p.bias != nullptr;
  // (1) bias slice -> registers (parked in smem after the barrier wait)
  float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
  if (use_bias && lane < NCH * 8) {
    const int c = colbase + lane * 4;
    if (c < p.N) bv = __ldg(reinterpret_cast<const float4*>(p.bias + c));
  }
  // (2) aux of chunk 0 in flight
  uint4 ax[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) ax[g] = make_uint4(0u, 0u, 0u, 0u);
  if (HAS_AUX) aux_fetch(p, row0, colbase, ax);
  mbar_wait(tfull, parity);
  tc_fence_after();
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int col0 = colbase + c * 32;
    uint4 axn[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) axn[g] = make_uint4(0u, 0u, 0u, 0u);
    if (HAS_AUX && c + 1 < NCH) aux_fetch(p, row0, col0 + 32, axn);
    if (col0 < p.N) {  // warp-uniform
      if (HAS_AUX) aux_transpose<BUFS>(ax, stage, sidx);     // coalesced fetch -> one row per lane
      uint32_t r[32];
      tmem_ld32(taddr + static_cast<uint32_t>(c * 32), r);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * alpha;
      if (use_bias) {   // lane (c*8 + g) holds the bias of columns c*32 + 4g .. 4g+3: broadcast by shuffle, no smem
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const int src = c * 8 + g;
          v[4 * g + 0] += __shfl_sync(0xffffffffu, bv.x, src);
          v[4 * g + 1] += __shfl_sync(0xffffffffu, bv.y, src);
          v[4 * g + 2] += __shfl_sync(0xffffffffu, bv.z, src);
          v[4 * g + 3] += __shfl_sync(0xffffffffu, bv.w, src);
        }
      }
      if (EPI == DC_EPI_BF16_RESID || EPI == DC_EPI_BF16_DGELU) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float a8[8];
          float2 f;
          f = unpack_bf16x2(ax[g].x); a8[0] = f.x; a8[1] = f.y;
          f = unpack_bf16x2(ax[g].y); a8[2] = f.x; a8[3] = f.y;
          f = unpack_bf16x2(ax[g].z); a8[4] = f.x; a8[5] = f.y;
          f = unpack_bf16x2(ax[g].w); a8[6] = f.x; a8[7] = f.y;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (EPI == DC_EPI_BF16_RESID) v[8 * g + i] += a8[i];
            else v[8 * g + i] *= quick_gelu_grad(a8[i]);
          }
        }
      }
      if (OUT_BF16) {
        if (EPI == DC_EPI_BF16_GELU) {
          stage_store_chunk<BUFS>(v, stage, sidx, tm_out2, col0, row0);   // pre-activation u (saved for backward)
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = quick_gelu(v[i]);
        }
        stage_store_chunk<BUFS>(v, stage, sidx, tm_out, col0, row0);
        if (p.colsum != nullptr) {                            // fused bias gradient: colsum += sum_rows out
          if (!row_ok) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0.f;
          }
          const float cs = chunk_colsum(v);
          if (col0 + lane < p.N) atomicAdd(p.colsum + col0 + lane, cs);
        }
      } else if (GMAX) {
        // max / arg-max over each group of 16 consecutive columns (two groups per 32-column chunk), kept in registers
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float best = v[16 * h];
          int bi = 0;
#pragma unroll
          for (int m = 1; m < 16; ++m)
            if (v[16 * h + m] > best) { best = v[16 * h + m]; bi = m; }
          gbest[GMAX ? 2 * c + h : 0] = best;
          garg[GMAX ? 2 * c + h : 0] = bi;
        }
      } else if (row_ok) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int col = col0 + g * 8;
          if (col < p.N) {
            float* dst = static_cast<float*>(p.out) + static_cast<size_t>(row) * p.ldo + col;
            if (EPI == DC_EPI_F32) {
              *reinterpret_cast<float4*>(dst) = make_float4(v[8 * g], v[8 * g + 1], v[8 * g + 2], v[8 * g + 3]);
              *reinterpret_cast<float4*>(dst + 4) = make_float4(v[8 * g + 4], v[8 * g + 5], v[8 * g + 6], v[8 * g + 7]);
            } else {
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v[8 * g]), "f"(v[8 * g + 1]),
                           "f"(v[8 * g + 2]), "f"(v[8 * g + 3])
                           : "memory");
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4), "f"(v[8 * g + 4]),
                           "f"(v[8 * g + 5]), "f"(v[8 * g + 6]), "f"(v[8 * g + 7])
                           : "memory");
            }
          }
        }
      }
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) ax[g] = axn[g];
  }
  if (GMAX && row_ok) {
    const int groups = p.N >> 4;
    float* o = static_cast<float*>(p.out) + static_cast<size_t>(row) * p.ldo;
    uint8_t* a = static_cast<uint8_t*>(p.out2) + static_cast<size_t>(row) * p.ldo2;
#pragma unroll
    for (int g = 0; g < 2 * NCH; ++g) {
      const int gi = (colbase >> 4) + g;
      if (gi < groups) {
        o[gi] = gbest[GMAX ? g : 0];
        a[gi] = static_cast<uint8_t>(garg[GMAX ? g : 0]);
      }
    }
  }
}

}  // namespace dc
