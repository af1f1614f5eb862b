// declip_b200 — persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   out[M,N] (op)= epilogue(alpha * sum_k A(m,k) * B(n,k)),  bf16 operands, fp32 accumulate in TMEM.
//
// One CTA per SM (persistent, static tile schedule, n-fastest so concurrently running CTAs share
// the same A rows / the whole of B through L2).  384 threads:
//   warp 0   : TMA producer  (cp.async.bulk.tensor 2-D boxes -> 128B-swizzled smem stages)
//   warp 1   : UMMA issuer   (one elected lane, tcgen05.mma.cta_group::1.kind::f16, 128 x BN x 16)
//   warp 2   : TMEM allocator (2 accumulator stages x BN fp32 columns)
//   warps 4-11: epilogue     (tcgen05.ld 32x32b.x32 -> registers -> fused epilogue -> global; two warps per
//                             TMEM lane quadrant, each draining half of the tile's columns)
// Three pipelines: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue; the epilogue of
// tile i overlaps the main loop of tile i+1), and the tile loop itself.
//
// Operand "major-ness" is a template parameter so forward (K-major x K-major), dgrad
// (K-major x MN-major) and wgrad (MN-major x MN-major) all read the tensors where they lie in
// HBM — no transposed copies are ever written.
//
// Replaces (reference, /root/reference): every nn.Linear/F.linear/`@` on the training hot path —
// prototype/model/image_encoder/base_transformer.py:33-41, visual_transformer.py:56,72,
// text_encoder/text_transformer.py:203, model/clip.py:140-141 — and their autograd backward.
#include <string.h>
#include <atomic>
#include "gemm_common.cuh"

namespace dc {

int gemm2_bf16(const dc_gemm_args& a, cudaStream_t stream);   // gemm2.cu
static std::atomic<int> g_gemm_2cta{1};   // cta_group::2 by default where the problem is >= one 256x256 cluster tile

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int STAGING_BYTES = EPI_WARPS * 4096;       // per-epilogue-warp 2 x (32x32 bf16) TMA-store staging
  static constexpr int BIAS_BYTES = 0;                          // bias is broadcast by warp shuffles
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 256 /*barriers*/ + BIAS_BYTES +
                                    1024 /*align slack*/;
};

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmOut2,
                 const GemmKParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + Cfg::STAGING_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
    if (p.epi <= DC_EPI_BF16_DGELU) prefetch_tensormap(&tmOut);
    if (p.epi == DC_EPI_BF16_GELU) prefetch_tensormap(&tmOut2);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], EPI_WARPS);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                 // everything above overlaps the tail of the previous kernel in the stream
  pdl_launch_dependents();

  const int tiles_mn = p.num_m * p.num_n;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        const int split = t / tiles_mn;
        const int mn = t - split * tiles_mn;
        const int m_blk = mn / p.num_n;
        const int n_blk = mn - m_blk * p.num_n;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.total_kb, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          if (!A_MN) {
            tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, m_blk * BM);
          } else {
#pragma unroll
            for (int a = 0; a < BM / 64; ++a)
              tma_load_2d(sa + a * (BK * 128), &tmA, &full_bar[stage], m_blk * BM + a * 64, kb * BK);
          }
          if (!B_MN) {
            tma_load_2d(sb, &tmB, &full_bar[stage], kb * BK, n_blk * BN);
          } else {
#pragma unroll
            for (int a = 0; a < BN / 64; ++a)
              tma_load_2d(sb + a * (BK * 128), &tmB, &full_bar[stage], n_blk * BN + a * 64, kb * BK);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ UMMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        const int split = t / tiles_mn;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.total_kb, kb0 + p.kb_per_split);
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // K-major: advance 16 elements (32 B) inside the 128 B swizzle row.
            // MN-major: advance 16 k-rows (16 * 128 B).
            const uint64_t adesc = A_MN ? umma_smem_desc(sa + k * 2048, BK * 128, 1024)
                                        : umma_smem_desc(sa + k * 32, 16, 1024);
            const uint64_t bdesc = B_MN ? umma_smem_desc(sb + k * 2048, BK * 128, 1024)
                                        : umma_smem_desc(sb + k * 32, 16, 1024);
            umma_bf16(tmem_d, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem stage once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[as]);  // accumulator complete -> epilogue
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (TMEM -> regs -> global)
    const int quad = warp & 3;            // TMEM lane quadrant this warp may access
    const int half = (warp - 4) >> 2;     // which half of the tile's columns
    constexpr int NCH = BN / 64;          // 32-column chunks per warp
    uint8_t* stage = staging + (warp - 4) * 4096;
    uint32_t sidx = 0;
    const float alpha = p.alpha * (p.alpha_dev != nullptr ? __ldg(p.alpha_dev) : 1.0f);
    int as = 0;
    uint32_t aphase = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
      const int split = t / tiles_mn;
      const int mn = t - split * tiles_mn;
      const int m_blk = mn / p.num_n;
      const int n_blk = mn - m_blk * p.num_n;
      const int row0 = m_blk * BM + quad * 32;
      const int colbase = n_blk * BN + half * (BN / 2);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                             static_cast<uint32_t>(as * BN + half * (BN / 2));
#define DC_EPI_CASE(E) \
  case E: epilogue_tile<E, NCH>(p, &tmOut, &tmOut2, alpha, taddr, row0, colbase, sidx, stage, &tfull_bar[as], aphase); break
      switch (p.epi) {
        DC_EPI_CASE(DC_EPI_BF16);
        DC_EPI_CASE(DC_EPI_BF16_GELU);
        DC_EPI_CASE(DC_EPI_BF16_RESID);
        DC_EPI_CASE(DC_EPI_BF16_DGELU);
        DC_EPI_CASE(DC_EPI_F32);
        DC_EPI_CASE(DC_EPI_F32_GROUPMAX16);
        default: epilogue_tile<DC_EPI_F32_ATOMIC, NCH>(p, &tmOut, &tmOut2, alpha, taddr, row0, colbase, sidx, stage,
                                                       &tfull_bar[as], aphase); break;
      }
#undef DC_EPI_CASE
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if (lane == 0) bulk_wait_read0();  // staging smem must outlive the last TMA store's read
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------ host side
template <int BN, bool A_MN, bool B_MN>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const CUtensorMap& tmO2,
                       const GemmKParams& p, int grid, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  auto kern = gemm_bf16_kernel<BN, A_MN, B_MN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return set_error_cuda("cudaFuncSetAttribute(gemm)", e);
    attr_set = true;
  }
  cudaError_t e = launch_pdl(kern, dim3(grid), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, stream, tmA, tmB, tmO, tmO2, p);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return set_error_cuda("gemm launch", e);
  count_launch();
  return 0;
}

int gemm_bf16(const dc_gemm_args& a, cudaStream_t stream) {
  if (a.M <= 0 || a.N <= 0 || a.K <= 0) return set_error("gemm: empty problem");
  const bool gmax = a.epilogue == DC_EPI_F32_GROUPMAX16;      // scalar stores of [M, N/16] values: any row stride
  if ((a.N & 7) || (a.lda & 7) || (a.ldb & 7) || (!gmax && (a.ldo & 3)))
    return set_error("gemm: N, lda, ldb must be multiples of 8 (ldo of 4)");
  if (a.epilogue < 0 || a.epilogue > DC_EPI_F32_GROUPMAX16) return set_error("gemm: bad epilogue");
  if (a.epilogue == DC_EPI_F32_GROUPMAX16 && (a.out2 == nullptr || (a.N & 15)))
    return set_error("gemm: the group-max epilogue needs out2 (arg-max bytes) and N % 16 == 0");
  if ((a.epilogue == DC_EPI_BF16_RESID || a.epilogue == DC_EPI_BF16_DGELU) && a.aux == nullptr)
    return set_error("gemm: epilogue needs aux");
  if (a.epilogue == DC_EPI_BF16_GELU && a.out2 == nullptr) return set_error("gemm: GELU epilogue needs out2");

  if (g_gemm_2cta.load(std::memory_order_relaxed) && (a.block_n == 0 || a.block_n == 256)) {
    const int rc2 = gemm2_bf16(a, stream);   // 1 = problem too small for 256x256 cluster tiles -> 1-CTA kernel below
    if (rc2 != 1) return rc2;
  }
  int BN = a.block_n;
  const int sms = sm_count();
  if (BN == 0) BN = (a.N <= 128) ? 128 : 256;   // 128x256 tiles feed the tensor pipe ~25 % better (measured)
  if (BN != 128 && BN != 256) return set_error("gemm: block_n must be 128 or 256");

  GemmKParams p;
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.num_m = (a.M + BM - 1) / BM;
  p.num_n = (a.N + BN - 1) / BN;
  p.total_kb = (a.K + BK - 1) / BK;
  int splits = a.splits;
  if (a.epilogue != DC_EPI_F32_ATOMIC) splits = 1;
  if (splits <= 0) splits = choose_splits(p.num_m * p.num_n, p.total_kb, sms);
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

  CUtensorMap tmA, tmB;
  int rc;
  // K-major operand [rows, K]: box {64 k, rows_per_tile}.  MN-major operand [K, mn]: box {64 mn, 64 k}.
  if (!a.a_mn_major) rc = make_tmap_2d(&tmA, a.A, a.K, a.M, a.lda, 64, BM);
  else               rc = make_tmap_2d(&tmA, a.A, a.M, a.K, a.lda, 64, BK);
  if (rc) return rc;
  if (!a.b_mn_major) rc = make_tmap_2d(&tmB, a.B, a.K, a.N, a.ldb, 64, BN);
  else               rc = make_tmap_2d(&tmB, a.B, a.N, a.K, a.ldb, 64, BK);
  if (rc) return rc;

  // bf16 outputs leave through TMA stores of 32x32 boxes (64-byte swizzle)
  CUtensorMap tmO, tmO2;
  memset(&tmO, 0, sizeof(tmO));
  memset(&tmO2, 0, sizeof(tmO2));
  if (a.epilogue <= DC_EPI_BF16_DGELU) {
    if (a.ldo & 7) return set_error("gemm: ldo must be a multiple of 8 for bf16 outputs");
    rc = make_tmap_2d(&tmO, a.out, a.N, a.M, a.ldo, 32, 32, 64);
    if (rc) return rc;
    if (a.epilogue == DC_EPI_BF16_GELU) {
      if (a.ldo2 & 7) return set_error("gemm: ldo2 must be a multiple of 8");
      rc = make_tmap_2d(&tmO2, a.out2, a.N, a.M, a.ldo2, 32, 32, 64);
      if (rc) return rc;
    }
  }
  p.colsum = (a.epilogue <= DC_EPI_BF16_DGELU) ? a.colsum : nullptr;

  const int grid = p.num_tiles < sms ? p.num_tiles : sms;
#define DC_LAUNCH(BN_, AM_, BM_) return launch_gemm<BN_, AM_, BM_>(tmA, tmB, tmO, tmO2, p, grid, stream)
  if (BN == 256) {
    if (!a.a_mn_major && !a.b_mn_major) DC_LAUNCH(256, false, false);
    if (!a.a_mn_major && a.b_mn_major) DC_LAUNCH(256, false, true);
    if (a.a_mn_major && !a.b_mn_major) DC_LAUNCH(256, true, false);
    DC_LAUNCH(256, true, true);
  } else {
    if (!a.a_mn_major && !a.b_mn_major) DC_LAUNCH(128, false, false);
    if (!a.a_mn_major && a.b_mn_major) DC_LAUNCH(128, false, true);
    if (a.a_mn_major && !a.b_mn_major) DC_LAUNCH(128, true, false);
    DC_LAUNCH(128, true, true);
  }
#undef DC_LAUNCH
}

}  // namespace dc

extern "C" int dc_set_gemm_2cta(int enable) {
  const int old = dc::g_gemm_2cta.exchange(enable ? 1 : 0);
  return old;
}

extern "C" int dc_gemm_bf16(const dc_gemm_args* args, dc_stream_t stream) {
  if (args == nullptr) return dc::set_error("dc_gemm_bf16: null args");
  return dc::gemm_bf16(*args, static_cast<cudaStream_t>(stream));
}
