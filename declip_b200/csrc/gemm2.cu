// declip_b200 — 2-CTA (cta_group::2) variant of the persistent tcgen05 GEMM: a cluster of two CTAs on the two SMs of
// a TPC computes one 256 x 256 output tile.  Each CTA stages its own 128 rows of A and HALF of the B tile
// (128 of the 256 N rows) per k-block — 32 KiB per stage instead of 48 KiB — and the leader CTA issues
// tcgen05.mma.cta_group::2 (M = 256): the tensor cores of both SMs read the B halves from both shared memories, so
// per-SM shared-memory traffic (TMA writes + UMMA reads) drops from ~192 to ~128 B/clk, the limit that holds the
// 1-CTA 128x256 kernel at ~0.7 of peak.  Accumulators: each CTA's TMEM holds its own 128 rows x 256 columns
// (2 stages), drained by that CTA's 8 epilogue warps with the same fused epilogue as gemm.cu.
//
// Cross-CTA synchronisation:
//   full[s]   (leader smem, count 1): leader arrive.expect_tx(2 x 32 KiB); both CTAs' TMA loads complete_tx on the
//             LEADER's barrier (.cta_group::2 TMA with a cluster address) — the peer never arrives explicitly.
//   empty[s], tfull[a] (each CTA, count 1): tcgen05.commit.cta_group::2 ... multicast::cluster, mask 0b11.
//   tempty[a] (leader smem, count 16): one arrive per epilogue warp of BOTH CTAs (remote arrive via mapa).
#include <string.h>
#include "gemm_common.cuh"

namespace dc {

constexpr int BN2 = 256;
#ifndef DC_G2_EPI_BUFS
#define DC_G2_EPI_BUFS 2
#endif
constexpr int G2_EPI_BUFS = DC_G2_EPI_BUFS;      // rotating 2 KiB output staging buffers per epilogue warp (4 buffers + 5 stages
                                                 // measured 1 % SLOWER than 2 + 6: profiles/r02_gemm_perf_v3_bufs_ab.md)
constexpr int G2_STAGES = G2_EPI_BUFS == 4 ? 5 : 6;   // the deeper store staging costs one of the six 32 KiB operand stages
constexpr int G2_A_BYTES = BM * BK * 2;          // 16 KiB: this CTA's 128 rows
constexpr int G2_B_BYTES = (BN2 / 2) * BK * 2;   // 16 KiB: this CTA's half of the 256 N rows
constexpr int G2_STAGE_BYTES = G2_A_BYTES + G2_B_BYTES;
constexpr int G2_TMEM_COLS = 2 * BN2;
constexpr int G2_STAGING_BYTES = EPI_WARPS * G2_EPI_BUFS * 2048;   // G2_EPI_BUFS x (32x32 bf16) per epilogue warp
constexpr int G2_BIAS_BYTES = 0;                      // bias is broadcast by warp shuffles
constexpr int G2_SMEM_BYTES = G2_STAGES * G2_STAGE_BYTES + G2_STAGING_BYTES + 256 + G2_BIAS_BYTES + 1024;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address) in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // relaxed: the TMEM reads this arrival publishes are already complete (tcgen05.wait::ld) and fenced with
  // tcgen05.fence::before_thread_sync; a release at cluster scope costs an ERRBAR + MEMBAR.ALL (22 % of the stall
  // samples of the first 2-CTA version) and delays the MMA warp's re-use of the accumulator stage.
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const void* tm, uint32_t mbar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(mbar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <bool A_MN, bool B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmOut2,
                  const GemmKParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + G2_STAGES * G2_STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + G2_STAGING_BYTES);
  uint64_t* empty_bar = full_bar + G2_STAGES;
  uint64_t* tfull_bar = empty_bar + G2_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
    if (p.epi <= DC_EPI_BF16_DGELU) prefetch_tensormap(&tmOut);
    if (p.epi == DC_EPI_BF16_GELU) prefetch_tensormap(&tmOut2);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < G2_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);   // the leader's expect_tx arrive; both CTAs' TMA loads complete_tx on the leader's copy
      mbar_init(&empty_bar[s], 1);  // multicast commit from the leader's MMA thread
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);                // multicast commit
      mbar_init(&tempty_bar[s], 2 * EPI_WARPS);   // epilogue warps of both CTAs (only the leader's copy is used)
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_2sm(tmem_slot, G2_TMEM_COLS);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();   // barriers of both CTAs initialised and TMEM allocated before any remote traffic
  tc_fence_after();
  pdl_wait();                 // everything above overlaps the tail of the previous kernel in the stream
  pdl_launch_dependents();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_mn = p.num_m * p.num_n;   // num_m counts 256-row tiles here

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = cluster_id; t < p.num_tiles; t += num_clusters) {
        const int split = t / tiles_mn;
        const int mn = t - split * tiles_mn;
        const int m_blk = mn / p.num_n;
        const int n_blk = mn - m_blk * p.num_n;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.total_kb, kb0 + p.kb_per_split);
        const int row_a = m_blk * 2 * BM + static_cast<int>(rank) * BM;           // this CTA's 128 rows of A
        const int row_b = n_blk * BN2 + static_cast<int>(rank) * (BN2 / 2);       // this CTA's half of the N rows
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          const uint32_t full_leader = mapa_u32(smem_u32(&full_bar[stage]), 0);
          // Only the leader arrives (expecting the bytes of BOTH CTAs); the peer's loads just complete_tx on the
          // leader's barrier.  A per-stage remote arrive.release.cluster from the peer costs a cluster-scope fence
          // (ERRBAR/MEMBAR, 7 % of all stall samples in the first version) and halved the tensor-pipe utilisation.
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * G2_STAGE_BYTES);
          uint8_t* sa = smem + stage * G2_STAGE_BYTES;
          uint8_t* sb = sa + G2_A_BYTES;
          if (!A_MN) {
            tma_load_2d_2sm(sa, &tmA, full_leader, kb * BK, row_a);
          } else {
#pragma unroll
            for (int a = 0; a < BM / 64; ++a) tma_load_2d_2sm(sa + a * (BK * 128), &tmA, full_leader, row_a + a * 64, kb * BK);
          }
          if (!B_MN) {
            tma_load_2d_2sm(sb, &tmB, full_leader, kb * BK, row_b);
          } else {
#pragma unroll
            for (int a = 0; a < BN2 / 2 / 64; ++a)
              tma_load_2d_2sm(sb + a * (BK * 128), &tmB, full_leader, row_b + a * 64, kb * BK);
          }
          if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ UMMA issuer (leader CTA only)
    if (leader && elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN2, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int t = cluster_id; t < p.num_tiles; t += num_clusters) {
        const int split = t / tiles_mn;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.total_kb, kb0 + p.kb_per_split);
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * BN2);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * G2_STAGE_BYTES);
          const uint32_t sb = sa + G2_A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adesc = A_MN ? umma_smem_desc(sa + k * 2048, BK * 128, 1024)
                                        : umma_smem_desc(sa + k * 32, 16, 1024);
            const uint64_t bdesc = B_MN ? umma_smem_desc(sb + k * 2048, BK * 128, 1024)
                                        : umma_smem_desc(sb + k * 32, 16, 1024);
            umma_bf16_2sm(tmem_d, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit_2sm(&empty_bar[stage]);   // frees the stage in BOTH CTAs
          if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_2sm(&tfull_bar[as]);        // accumulator ready in BOTH CTAs
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (both CTAs, own 128 rows)
    const int quad = warp & 3;
    const int half = (warp - 4) >> 2;
    constexpr int NCH = BN2 / 64;
    uint8_t* stage_buf = staging + (warp - 4) * (G2_EPI_BUFS * 2048);
    uint32_t sidx = 0;
    const float alpha = p.alpha * (p.alpha_dev != nullptr ? __ldg(p.alpha_dev) : 1.0f);
    int as = 0;
    uint32_t aphase = 0;
    for (int t = cluster_id; t < p.num_tiles; t += num_clusters) {
      const int split = t / tiles_mn;
      const int mn = t - split * tiles_mn;
      const int m_blk = mn / p.num_n;
      const int n_blk = mn - m_blk * p.num_n;
      const int row0 = m_blk * 2 * BM + static_cast<int>(rank) * BM + quad * 32;
      const int colbase = n_blk * BN2 + half * (BN2 / 2);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                             static_cast<uint32_t>(as * BN2 + half * (BN2 / 2));
#define DC_EPI_CASE(E) \
  case E: epilogue_tile<E, NCH, G2_EPI_BUFS>(p, &tmOut, &tmOut2, alpha, taddr, row0, colbase, sidx, stage_buf, &tfull_bar[as], aphase); break
      switch (p.epi) {
        DC_EPI_CASE(DC_EPI_BF16);
        DC_EPI_CASE(DC_EPI_BF16_GELU);
        DC_EPI_CASE(DC_EPI_BF16_RESID);
        DC_EPI_CASE(DC_EPI_BF16_DGELU);
        DC_EPI_CASE(DC_EPI_F32);
        DC_EPI_CASE(DC_EPI_F32_GROUPMAX16);
        default: epilogue_tile<DC_EPI_F32_ATOMIC, NCH, G2_EPI_BUFS>(p, &tmOut, &tmOut2, alpha, taddr, row0, colbase, sidx, stage_buf,
                                                       &tfull_bar[as], aphase); break;
      }
#undef DC_EPI_CASE
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[as]), 0));   // the leader's barrier
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if (lane == 0) bulk_wait_read0();
    __syncwarp();
  }

  tc_fence_before();
  cluster_sync_all();   // no CTA leaves (or frees TMEM) while its peer can still signal it or read its smem
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, G2_TMEM_COLS);
  }
}

template <bool A_MN, bool B_MN>
static int launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const CUtensorMap& tmO2,
                        const GemmKParams& p, int clusters, cudaStream_t stream) {
  auto kern = gemm2_bf16_kernel<A_MN, B_MN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM_BYTES);
    if (e != cudaSuccess) return set_error_cuda("cudaFuncSetAttribute(gemm2)", e);
    attr_set = true;
  }
  cudaError_t e = launch_pdl(kern, dim3(2 * clusters), dim3(GEMM_THREADS), G2_SMEM_BYTES, stream, tmA, tmB, tmO, tmO2, p);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return set_error_cuda("gemm2 launch", e);
  count_launch();
  return 0;
}

// Same contract as gemm_bf16(); returns 1 when the problem does not qualify (caller falls back to the 1-CTA kernel).
int gemm2_bf16(const dc_gemm_args& a, cudaStream_t stream) {
  if (a.M < 2 * BM || a.N < BN2) return 1;
  const int sms = sm_count();
  GemmKParams p;
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.num_m = (a.M + 2 * BM - 1) / (2 * BM);
  p.num_n = (a.N + BN2 - 1) / BN2;
  p.total_kb = (a.K + BK - 1) / BK;
  const int clusters_max = sms / 2;
  int splits = a.splits;
  if (a.epilogue != DC_EPI_F32_ATOMIC) splits = 1;
  if (splits <= 0) splits = choose_splits(p.num_m * p.num_n, p.total_kb, clusters_max);
  if (splits > p.total_kb) splits = p.total_kb;
  p.kb_per_split = (p.total_kb + splits - 1) / splits;
  p.splits = (p.total_kb + p.kb_per_split - 1) / p.kb_per_split;
  p.num_tiles = p.num_m * p.num_n * p.splits;
  p.epi = a.epilogue;
  p.alpha = a.alpha;
  p.out = a.out; p.ldo = a.ldo;
  p.out2 = a.out2; p.ldo2 = a.ldo2;
  p.bias = a.bias;
  p.aux = static_cast<const bf16*>(a.aux); p.ldaux = a.ldaux;
  p.alpha_dev = a.alpha_dev;

  CUtensorMap tmA, tmB, tmO, tmO2;
  int rc;
  if (!a.a_mn_major) rc = make_tmap_2d(&tmA, a.A, a.K, a.M, a.lda, 64, BM);
  else               rc = make_tmap_2d(&tmA, a.A, a.M, a.K, a.lda, 64, BK);
  if (rc) return rc;
  if (!a.b_mn_major) rc = make_tmap_2d(&tmB, a.B, a.K, a.N, a.ldb, 64, BN2 / 2);
  else               rc = make_tmap_2d(&tmB, a.B, a.N, a.K, a.ldb, 64, BK);
  if (rc) return rc;
  memset(&tmO, 0, sizeof(tmO));
  memset(&tmO2, 0, sizeof(tmO2));
  if (a.epilogue <= DC_EPI_BF16_DGELU) {
    if (a.ldo & 7) return set_error("gemm: ldo must be a multiple of 8 for bf16 outputs");
    rc = make_tmap_2d(&tmO, a.out, a.N, a.M, a.ldo, 32, 32, 64);
    if (rc) return rc;
    if (a.epilogue == DC_EPI_BF16_GELU) {
      rc = make_tmap_2d(&tmO2, a.out2, a.N, a.M, a.ldo2, 32, 32, 64);
      if (rc) return rc;
    }
  }
  p.colsum = (a.epilogue <= DC_EPI_BF16_DGELU) ? a.colsum : nullptr;
  const int clusters = p.num_tiles < clusters_max ? p.num_tiles : clusters_max;
  if (!a.a_mn_major && !a.b_mn_major) return launch_gemm2<false, false>(tmA, tmB, tmO, tmO2, p, clusters, stream);
  if (!a.a_mn_major && a.b_mn_major) return launch_gemm2<false, true>(tmA, tmB, tmO, tmO2, p, clusters, stream);
  if (a.a_mn_major && !a.b_mn_major) return launch_gemm2<true, false>(tmA, tmB, tmO, tmO2, p, clusters, stream);
  return launch_gemm2<true, true>(tmA, tmB, tmO, tmO2, p, clusters, stream);
}

}  // namespace dc
