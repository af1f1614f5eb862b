// declip_b200 — fused distributed contrastive head (sm_100a): the logit strips of CLIP.forward (clip.py:129-141),
// ClipInfoCELoss (loss_functions/loss.py:40-50), accuracy top-1/top-5 (utils/misc.py:415-428) and their backward in
// three launches, with no [b, N] strip and no N x N matrix in HBM:
//
//   head_prep_kernel : L2-normalise the tower outputs (clip.py:129-130) straight into this rank's rows of the
//                      gather buffer (bf16 [b, F*E]) and zero the accumulators of the other two kernels.
//   head_fwd_kernel  : per direction d (image->text, text->image) and 128-row block: S = X_loc Y_all^T on the tensor
//                      cores (TMA -> smem ring -> tcgen05.mma, accumulator in TMEM), and straight out of TMEM the
//                      ONLINE row softmax statistics (max, sum, softmax-weighted mean logit), the label logit and the
//                      rank of the label.  Column ranges are split over CTAs; the last CTA of a row block merges the
//                      partials and adds the row losses / accuracy counts / d(logit_scale) terms to 6 scalars.
//   head_bwd_kernel  : flash-style recomputation.  Per (direction, row block, 256-wide slice of E, column range):
//                      S tile -> W = g_own (softmax_row - Y) + g_peer (softmax_col - Y) as bf16 in shared memory
//                      (the column softmax only needs the OTHER direction's row LSEs of all ranks — an
//                      all-gather of 2b+2 floats per rank replaces the reference's all-reduce of two [N, E] gradients,
//                      clip.py:43-49) -> dX += W Y_all, a second tcgen05.mma whose A operand is that W tile and whose B
//                      operand is the SAME Y tile read MN-major.  dX accumulates in TMEM over the CTA's column tiles,
//                      leaves through fp32 red.global.add, and the last CTA of a row block applies the L2-norm
//                      backward and writes d(image_features) / d(text_features).
//
// Y_all is read through one TMA tensor map per source: one source = the NCCL-gathered buffer [N, F*E]; W sources =
// every rank's own [b, F*E] rows mapped into this process (symmetric memory over NVLink): the kernel then pulls peer
// features tile by tile while it computes, and no all-gather collective runs at all.
#include <string.h>
#include "common.cuh"
#include "internal.h"

namespace dc {

constexpr int HD_BM = 128;            // rows per block = TMEM lanes
constexpr int HD_BK = 64;
constexpr int HD_THREADS = 192;       // warp 0 TMA, warp 1 UMMA + TMEM alloc, warps 2-5 softmax / drain (one row per thread)
constexpr int HD_MAX_SRC = 8;
constexpr int HD_MAX_JS = 32;         // column-range splits (partials per row)
// out[16]: 0,1 sum CE per direction; 2,3 sum (E_softmax[logit] - label logit); 4,5 top-1 / top-5 counts (direction 0);
//          8 s_used, 9 s_raw (head_prep); 10 d loss / d logit_scale (head_bwd)
constexpr int HD_OUT_SUSED = 8, HD_OUT_SRAW = 9, HD_OUT_DLS = 10;

// ---- workspace layout (floats), shared by the three kernels and the host wrapper
struct HeadWs {
  long long out, lse, g, alab, ea, cnt, part, dxn, total;
  __host__ __device__ HeadWs(int b, int e) {
    const int rb = (b + HD_BM - 1) / HD_BM;
    out = 0;                                   // [16] sum CE_A, sum CE_B, sum(ea - alab)_A, _B, top1, top5
    lse = 16;                                  // [2][b]  row log-sum-exp, then
    g = lse + 2ll * b;                         // [2]     upstream gradients of the two CE sums (exchange vector tail)
    alab = (g + 2 + 3) / 4 * 4;                // [2][b]  label logit
    ea = alab + 2ll * b;                       // [2][b]  softmax-weighted mean logit (for d logit_scale)
    cnt = ea + 2ll * b;                        // [4*rb]  int counters (fwd, bwd last-arriver per direction / row block)
    part = (cnt + 4ll * rb + 3) / 4 * 4;       // [2][b][HD_MAX_JS][4]
    dxn = part + 2ll * b * HD_MAX_JS * 4;      // [2][b][e]
    total = dxn + 2ll * b * e;
  }
};

struct HeadParams {
  int b, n, e, ld, n_src, rows_per_src, row0, cross;
  int x_off[2], y_off[2];
  int rb, js, es, ntiles, jt;                  // row blocks, column splits, E slices (bwd), column tiles, tile width (bwd)
  const bf16* x_base;
  float* ws;
  const float* g_own;                          // [2] upstream gradients of this rank's two CE sums (device)
  float* strips[2];
  int ld_strip;
  const float* exch;                           // [n_ranks][2b + 2]
  const float* x_raw[2];
  float eps[2];
  float* dx_out[2];
};

struct HeadMaps {
  CUtensorMap x[2];                            // X operand (this rank's rows), per direction
  CUtensorMap y[2][HD_MAX_SRC];                // Y operand per direction and source
};

__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

// ------------------------------------------------------------------------------------------------ prepare
struct PrepArgs {
  const float* x[8];
  float eps[8];
  int n_feats, b, e;
  bf16* out;
  float* ws;
  bf16* push[HD_MAX_SRC];              // peers' gather buffers (symmetric memory): this rank's rows are ALSO stored there
  int n_push;
  long long push_row0;                 // first row of this rank inside a gather buffer
  const float* logit_scale;            // raw parameter (device) or NULL
  float scale_max;                     // clamp of exp(logit_scale) (clip.py:133-134; +inf: none) / the constant scale when NULL
  long long zero_a0, zero_a1, zero_c0, zero_c1, zero_b0, zero_b1;      // float ranges to clear (b: 16-byte aligned)
};

__global__ void __launch_bounds__(256) head_prep_kernel(const PrepArgs a) {
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  for (int r = gw; r < a.b * a.n_feats; r += nw) {
    const int f = r / a.b, row = r - f * a.b;
    const float* xr = a.x[f] + static_cast<size_t>(row) * a.e;
    float s = 0.f;
    for (int c = lane * 4; c < a.e; c += 128) {
      const float4 v = *reinterpret_cast<const float4*>(xr + c);
      s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    const float inv = 1.0f / (sqrtf(warp_sum(s)) + a.eps[f]);
    const size_t off = (static_cast<size_t>(row) * a.n_feats + f) * a.e;
    bf16* o = a.out + off;
    const size_t poff = (static_cast<size_t>(a.push_row0 + row) * a.n_feats + f) * a.e;
    for (int c = lane * 4; c < a.e; c += 128) {
      const float4 v = *reinterpret_cast<const float4*>(xr + c);
      uint2 w;
      w.x = pack_bf16x2(v.x * inv, v.y * inv);
      w.y = pack_bf16x2(v.z * inv, v.w * inv);
      *reinterpret_cast<uint2*>(o + c) = w;
      // one-shot all-gather by peer stores over NVLink: every rank ends up with all rows in its OWN buffer, which the
      // head kernels then read through the local L2 (pulling peer tiles instead re-reads them once per row block:
      // measured +1.7 ms/step at 2 GPUs)
      for (int k = 0; k < a.n_push; ++k) *reinterpret_cast<uint2*>(a.push[k] + poff + c) = w;
    }
  }
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long nt = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = a.zero_a0 + tid; i < a.zero_a1; i += nt) a.ws[i] = 0.f;
  if (tid == 0) {      // s = min(exp(logit_scale), max) is used in the forward; d s / d logit_scale = exp(logit_scale)
    const float s_raw = a.logit_scale != nullptr ? __expf(*a.logit_scale) : a.scale_max;
    a.ws[a.zero_a0 + HD_OUT_SUSED] = fminf(s_raw, a.scale_max);
    a.ws[a.zero_a0 + HD_OUT_SRAW] = s_raw;
    for (int i = HD_OUT_DLS; i < 16; ++i) a.ws[a.zero_a0 + i] = 0.f;
  }
  for (long long i = a.zero_c0 + tid; i < a.zero_c1; i += nt) a.ws[i] = 0.f;      // int counters: 0 == 0.0f bit pattern
  for (long long i = a.zero_b0 + tid * 4; i < a.zero_b1; i += nt * 4) *reinterpret_cast<float4*>(a.ws + i) = make_float4(0.f, 0.f, 0.f, 0.f);
}

// label logit of one row: s * <x_row[x_off : x_off+e], x_row[y_off : y_off+e]> (the positive pair sits in the same row
// of the feature buffer: label = global index of the row itself, loss.py:42-45)
__device__ __forceinline__ float label_dot(const bf16* xrow, int x_off, int y_off, int e) {
  float acc = 0.f;
  for (int c = 0; c < e; c += 8) {
    const uint4 a = *reinterpret_cast<const uint4*>(xrow + x_off + c);
    const uint4 b = *reinterpret_cast<const uint4*>(xrow + y_off + c);
    float fa[8], fb[8];
    unpack8(a, fa);
    unpack8(b, fb);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc = fmaf(fa[i], fb[i], acc);
  }
  return acc;
}

// ------------------------------------------------------------------------------------------------ forward
constexpr int HF_BN = 256;
constexpr int HF_STAGES = 4;
constexpr int HF_STAGE_BYTES = (HD_BM + HF_BN) * HD_BK * 2;      // 48 KiB
constexpr int HF_SMEM = HF_STAGES * HF_STAGE_BYTES + 256 + 1024;

__global__ void __launch_bounds__(HD_THREADS, 1) head_fwd_kernel(const __grid_constant__ HeadMaps maps, const HeadParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + HF_STAGES * HF_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + HF_STAGES;
  uint64_t* tfull_bar = empty_bar + HF_STAGES;      // [2]
  uint64_t* tempty_bar = tfull_bar + 2;             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  int* s_flag = reinterpret_cast<int*>(tmem_slot + 1);
  float* s_red = reinterpret_cast<float*>(tmem_slot + 4);       // [4][4]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int u = blockIdx.x;
  const int js = u % p.js; u /= p.js;
  const int rbk = u % p.rb;
  const int d = u / p.rb;
  const int kblocks = p.e / HD_BK;
  const HeadWs L(p.b, p.e);

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < HF_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 128); }
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 2 * HF_BN); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int t = js; t < p.ntiles; t += p.js) {
        const int col0 = t * HF_BN;
        const int src = col0 / p.rows_per_src;
        const int srow = col0 - src * p.rows_per_src;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], HF_STAGE_BYTES);
          uint8_t* sa = smem + stage * HF_STAGE_BYTES;
          tma_load_2d(sa, &maps.x[d], &full_bar[stage], kb * HD_BK, rbk * HD_BM);
          tma_load_2d(sa + HD_BM * HD_BK * 2, &maps.y[d][src], &full_bar[stage], kb * HD_BK, srow);
          if (++stage == HF_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(HD_BM, HF_BN, false, false);
      int stage = 0; uint32_t phase = 0; int as = 0; uint32_t aphase = 0;
      for (int t = js; t < p.ntiles; t += p.js) {
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * HF_BN);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * HF_STAGE_BYTES);
          const uint32_t sb = sa + HD_BM * HD_BK * 2;
#pragma unroll
          for (int k = 0; k < HD_BK / 16; ++k)
            umma_bf16(tmem_d, umma_smem_desc(sa + k * 32, 16, 1024), umma_smem_desc(sb + k * 32, 16, 1024), idesc,
                      (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (++stage == HF_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[as]);
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // -------------------------------------------------------------- online softmax statistics, one row per thread
    const int quad = warp & 3;
    const int row = rbk * HD_BM + quad * 32 + lane;        // local row
    const bool valid = row < p.b;
    const float s = __ldcg(p.ws + L.out + HD_OUT_SUSED);
    const int label = p.row0 + row;                         // global column of the positive (loss.py:45)
    float alab = 0.f;
    if (valid) alab = s * label_dot(p.x_base + static_cast<size_t>(row) * p.ld, p.x_off[d], p.y_off[d], p.e);
    const float s2 = s * LOG2E, alab2 = alab * LOG2E;       // statistics in the log2 domain (one FMUL folded into the scale)
    float m = -INFINITY, l = 0.f, ea = 0.f;
    int cnt = 0;
    float* strip = (p.strips[d] != nullptr && valid) ? p.strips[d] + static_cast<size_t>(row) * p.ld_strip : nullptr;
    int as = 0; uint32_t aphase = 0;
    for (int t = js; t < p.ntiles; t += p.js) {
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(as * HF_BN);
#pragma unroll 1
      for (int c = 0; c < HF_BN / 32; ++c) {
        const int col0 = t * HF_BN + c * 32;
        if (col0 >= p.n) break;                             // uniform
        uint32_t r[32];
        tmem_ld32(taddr + static_cast<uint32_t>(c * 32), r);
        tmem_ld_wait();
        float v[32];
        float cmax = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          v[i] = (col0 + i < p.n) ? __uint_as_float(r[i]) * s2 : -INFINITY;
          cmax = fmaxf(cmax, v[i]);
        }
        if (strip != nullptr) {
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            if (col0 + i < p.n)      // n is a multiple of 4 on this path (checked on the host)
              *reinterpret_cast<float4*>(strip + col0 + i) =
                  make_float4(__uint_as_float(r[i]) * s, __uint_as_float(r[i + 1]) * s, __uint_as_float(r[i + 2]) * s,
                              __uint_as_float(r[i + 3]) * s);
        }
        const float mn = fmaxf(m, cmax);
        const float corr = ex2f(m - mn);                    // exp2(-inf) = 0 on the first chunk
        float ls = 0.f, es = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float pexp = ex2f(v[i] - mn);
          ls += pexp;
          es = fmaf(pexp, v[i], es);                        // -inf columns: pexp = 0 but 0 * -inf = nan -> guarded below
          cnt += (v[i] > alab2 && col0 + i != label) ? 1 : 0;
        }
        if (col0 + 32 > p.n) {                              // ragged last chunk: redo the weighted sum without the masked columns
          es = 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (col0 + i < p.n) es = fmaf(ex2f(v[i] - mn), v[i], es);
        }
        l = l * corr + ls;
        ea = ea * corr + es;
        m = mn;
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    // partial of this column range -> global; last CTA of the (direction, row block) merges
    float* part = p.ws + L.part;
    if (valid) {
      float4 o = make_float4(m, l, ea, __int_as_float(cnt));
      *reinterpret_cast<float4*>(part + ((static_cast<size_t>(d) * p.b + row) * HD_MAX_JS + js) * 4) = o;
    }
    __threadfence();
    named_bar_sync(1, 128);
    if (threadIdx.x == 64) {
      int* counter = reinterpret_cast<int*>(p.ws + L.cnt) + d * p.rb + rbk;
      *s_flag = (atomicAdd(counter, 1) == p.js - 1) ? 1 : 0;
    }
    named_bar_sync(1, 128);
    if (*s_flag) {
      __threadfence();
      float ce = 0.f, dl = 0.f, t1 = 0.f, t5 = 0.f;
      if (valid) {
        float M = -INFINITY;
        const float4* pp = reinterpret_cast<const float4*>(part + (static_cast<size_t>(d) * p.b + row) * HD_MAX_JS * 4);
        for (int j = 0; j < p.js; ++j) M = fmaxf(M, __ldcg(pp + j).x);
        float Ls = 0.f, Es = 0.f;
        int C = 0;
        for (int j = 0; j < p.js; ++j) {
          const float4 q = __ldcg(pp + j);
          const float w = ex2f(q.x - M);
          Ls = fmaf(q.y, w, Ls);
          Es = fmaf(q.z, w, Es);
          C += __float_as_int(q.w);
        }
        const float lse = (M + log2f(Ls)) * LN2;            // natural log-sum-exp of the row
        const float eav = Es / Ls * LN2;                    // softmax-weighted mean logit
        p.ws[L.lse + static_cast<size_t>(d) * p.b + row] = lse;
        p.ws[L.alab + static_cast<size_t>(d) * p.b + row] = alab;
        p.ws[L.ea + static_cast<size_t>(d) * p.b + row] = eav;
        ce = lse - alab;
        dl = eav - alab;
        t1 = C == 0 ? 1.f : 0.f;
        t5 = C < 5 ? 1.f : 0.f;
      }
      ce = warp_sum(ce); dl = warp_sum(dl); t1 = warp_sum(t1); t5 = warp_sum(t5);
      if (lane == 0) { s_red[quad * 4 + 0] = ce; s_red[quad * 4 + 1] = dl; s_red[quad * 4 + 2] = t1; s_red[quad * 4 + 3] = t5; }
      named_bar_sync(1, 128);
      if (threadIdx.x == 64) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        for (int q = 0; q < 4; ++q) { a0 += s_red[q * 4]; a1 += s_red[q * 4 + 1]; a2 += s_red[q * 4 + 2]; a3 += s_red[q * 4 + 3]; }
        atomicAdd(p.ws + L.out + d, a0);
        atomicAdd(p.ws + L.out + 2 + d, a1);
        if (d == 0) { atomicAdd(p.ws + L.out + 4, a2); atomicAdd(p.ws + L.out + 5, a3); }   // accuracy is logged on logits_per_image
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 2 * HF_BN); }
}

// ------------------------------------------------------------------------------------------------ backward
constexpr int HB_XSTAGES = 3;
constexpr int HB_XBYTES = HD_BM * HD_BK * 2;                       // 16 KiB
constexpr int HB_YBYTES_MAX = 128 * 1024;                          // jt * e * 2
constexpr int HB_WBYTES_MAX = HD_BM * 128 * 2;                     // 32 KiB (jt = 128)
constexpr int HB_SMEM = HB_YBYTES_MAX + HB_XSTAGES * HB_XBYTES + HB_WBYTES_MAX + 256 + 1024;
constexpr int HB_OCOLS = 256;

template <int JT>
__global__ void __launch_bounds__(HD_THREADS, 1) head_bwd_kernel(const __grid_constant__ HeadMaps maps, const HeadParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ybuf = smem;
  uint8_t* xbuf = ybuf + HB_YBYTES_MAX;
  uint8_t* wbuf = xbuf + HB_XSTAGES * HB_XBYTES;
  uint64_t* xfull = reinterpret_cast<uint64_t*>(wbuf + HB_WBYTES_MAX);
  uint64_t* xempty = xfull + HB_XSTAGES;
  uint64_t* yfull = xempty + HB_XSTAGES;
  uint64_t* yfree = yfull + 1;       // O MMAs of the tile retired: Y tile and W tile may be overwritten
  uint64_t* wfree = yfree + 1;
  uint64_t* sfull = wfree + 1;
  uint64_t* sfree = sfull + 1;
  uint64_t* wfull = sfree + 1;
  uint64_t* ofull = wfull + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ofull + 1);
  int* s_flag = reinterpret_cast<int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int u = blockIdx.x;
  const int js = u % p.js; u /= p.js;
  const int es = u % p.es; u /= p.es;
  const int rbk = u % p.rb;
  const int d = u / p.rb;
  const int kblocks = p.e / HD_BK;
  const HeadWs L(p.b, p.e);
  constexpr int YBOX = JT * HD_BK * 2;                             // bytes of one [JT x 64] box

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < HB_XSTAGES; ++s) { mbar_init(&xfull[s], 1); mbar_init(&xempty[s], 1); }
    mbar_init(yfull, 1); mbar_init(yfree, 1); mbar_init(wfree, 1); mbar_init(sfull, 1);
    mbar_init(sfree, 128); mbar_init(wfull, 128); mbar_init(ofull, 1);
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base;                                // [0, JT)
  const uint32_t tmem_o = tmem_base + 256;                          // [256, 512)

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0, it = 0;
      for (int t = js; t < p.ntiles; t += p.js, ++it) {
        const int col0 = t * JT;
        const int src = col0 / p.rows_per_src;
        const int srow = col0 - src * p.rows_per_src;
        if (it > 0) mbar_wait(yfree, (it - 1) & 1);
        mbar_arrive_expect_tx(yfull, static_cast<uint32_t>(kblocks) * YBOX);
        for (int kb = 0; kb < kblocks; ++kb) tma_load_2d(ybuf + kb * YBOX, &maps.y[d][src], yfull, kb * HD_BK, srow);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&xempty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&xfull[stage], HB_XBYTES);
          tma_load_2d(xbuf + stage * HB_XBYTES, &maps.x[d], &xfull[stage], kb * HD_BK, rbk * HD_BM);
          if (++stage == HB_XSTAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(HD_BM, JT, false, false);
      constexpr uint32_t idesc_o = umma_idesc_bf16(HD_BM, HB_OCOLS, false, true);
      int stage = 0; uint32_t phase = 0, it = 0;
      for (int t = js; t < p.ntiles; t += p.js, ++it) {
        mbar_wait(yfull, it & 1);
        if (it > 0) mbar_wait(sfree, (it - 1) & 1);                 // softmax threads have read the previous S tile
        tc_fence_after();
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&xfull[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(xbuf + stage * HB_XBYTES);
          const uint32_t sb = smem_u32(ybuf + kb * YBOX);
#pragma unroll
          for (int k = 0; k < HD_BK / 16; ++k)
            umma_bf16(tmem_s, umma_smem_desc(sa + k * 32, 16, 1024), umma_smem_desc(sb + k * 32, 16, 1024), idesc_s,
                      (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&xempty[stage]);
          if (++stage == HB_XSTAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(sfull);
        mbar_wait(wfull, it & 1);                                   // W tile written (generic proxy -> fenced for the async proxy)
        tc_fence_after();
        const uint32_t sw = smem_u32(wbuf);
        const uint32_t sy = smem_u32(ybuf + (es * (HB_OCOLS / 64)) * YBOX);   // the 4 boxes of this CTA's E slice
#pragma unroll
        for (int k = 0; k < JT / 16; ++k)
          umma_bf16(tmem_o, umma_smem_desc(sw + (k >> 2) * (HD_BM * 128) + (k & 3) * 32, 16, 1024),
                    umma_smem_desc(sy + k * 2048, YBOX, 1024), idesc_o, (it > 0 || k > 0) ? 1u : 0u);
        umma_commit(yfree);
        umma_commit(wfree);
      }
      umma_commit(ofull);
    }
    __syncwarp();
  } else {
    const int quad = warp & 3;
    const int row = rbk * HD_BM + quad * 32 + lane;
    const int rloc = quad * 32 + lane;                              // row inside the tile
    const bool valid = row < p.b;
    const float s = __ldcg(p.ws + L.out + HD_OUT_SUSED);
    const float s2 = s * LOG2E;
    const int xw = 2 * p.b + 2;                                     // exchange vector length per rank
    const int my_rank = p.row0 / p.b;
    const float g_own = __ldg(p.g_own + d);
    if (blockIdx.x == 0 && threadIdx.x == 64) {
      // d loss / d logit_scale = exp(logit_scale) / s * sum_d g_d * sum_i (E_softmax[logit] - label logit): the own-loss
      // part only — every rank adds its own, the gradient all-reduce sums them
      const float* o = p.ws + L.out;
      p.ws[L.out + HD_OUT_DLS] = __ldcg(o + HD_OUT_SRAW) / __ldcg(o + HD_OUT_SUSED) *
                                 (__ldg(p.g_own) * __ldcg(o + 2) + __ldg(p.g_own + 1) * __ldcg(o + 3));
    }
    const float lse_own2 = valid ? __ldg(p.exch + static_cast<size_t>(my_rank) * xw + d * p.b + row) * LOG2E : 0.f;
    const int label = p.row0 + row;
    uint32_t it = 0;
    for (int t = js; t < p.ntiles; t += p.js, ++it) {
      mbar_wait(sfull, it & 1);
      tc_fence_after();
      if (it > 0) mbar_wait(wfree, (it - 1) & 1);                   // the previous W tile has been consumed by its O MMAs
      const uint32_t taddr = tmem_s + (static_cast<uint32_t>(quad * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < JT / 32; ++c) {
        const int col0 = t * JT + c * 32;
        uint32_t r[32];
        tmem_ld32(taddr + static_cast<uint32_t>(c * 32), r);
        tmem_ld_wait();
        float w[32];
        // peer terms: column j belongs to rank j / b; its loss in the OTHER direction has row LSE exch[rank][(1-d) b + j % b]
        const int jr = min(col0, p.n - 1) / p.b;                    // 32 | b is checked on the host when cross is on
        const float g_peer = !p.cross ? 0.f : (jr == my_rank ? __ldg(p.g_own + (1 - d))
                                                            : __ldg(p.exch + static_cast<size_t>(jr) * xw + 2 * p.b + (1 - d)));
        const float* lse_peer = p.exch + static_cast<size_t>(jr) * xw + (1 - d) * p.b + (col0 - jr * p.b);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int col = col0 + i;
          float wi = 0.f;
          if (valid && col < p.n) {
            const float a2 = __uint_as_float(r[i]) * s2;
            const float y = (col == label) ? 1.f : 0.f;
            wi = g_own * (ex2f(a2 - lse_own2) - y);
            if (p.cross) wi = fmaf(g_peer, ex2f(a2 - __ldg(lse_peer + i) * LOG2E) - y, wi);
          }
          w[i] = wi;
        }
        // K-major A tile of the second MMA, 128-byte swizzle: box (c / 2) of 64 columns, row = 128 B, chunk16 ^= row & 7
        uint8_t* wrow = wbuf + (c >> 1) * (HD_BM * 128) + rloc * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 o;
          o.x = pack_bf16x2(w[8 * q + 0], w[8 * q + 1]); o.y = pack_bf16x2(w[8 * q + 2], w[8 * q + 3]);
          o.z = pack_bf16x2(w[8 * q + 4], w[8 * q + 5]); o.w = pack_bf16x2(w[8 * q + 6], w[8 * q + 7]);
          const int chunk = (c & 1) * 4 + q;
          *reinterpret_cast<uint4*>(wrow + ((chunk ^ (rloc & 7)) << 4)) = o;
        }
      }
      tc_fence_before();
      mbar_arrive(sfree);
      fence_proxy_async_smem();
      mbar_arrive(wfull);
    }
    // ---- drain dX slice: TMEM -> s * acc -> fp32 red.add
    mbar_wait(ofull, 0);
    tc_fence_after();
    float* dxn = p.ws + L.dxn + (static_cast<size_t>(d) * p.b + row) * p.e + es * HB_OCOLS;
#pragma unroll 1
    for (int c = 0; c < HB_OCOLS / 32; ++c) {
      uint32_t r[32];
      tmem_ld32(tmem_o + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(c * 32), r);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dxn + c * 32 + i),
                       "f"(__uint_as_float(r[i]) * s), "f"(__uint_as_float(r[i + 1]) * s), "f"(__uint_as_float(r[i + 2]) * s),
                       "f"(__uint_as_float(r[i + 3]) * s)
                       : "memory");
      }
    }
    // ---- last CTA of the (direction, row block): L2-norm backward on the finished rows (clip.py:129-130)
    __threadfence();
    named_bar_sync(1, 128);
    if (threadIdx.x == 64) {
      int* counter = reinterpret_cast<int*>(p.ws + L.cnt) + 2 * p.rb + d * p.rb + rbk;
      *s_flag = (atomicAdd(counter, 1) == p.js * p.es - 1) ? 1 : 0;
    }
    named_bar_sync(1, 128);
    if (*s_flag && p.dx_out[d] != nullptr) {
      __threadfence();
      const float eps = p.eps[d];
      for (int rr = quad; rr < HD_BM; rr += 4) {                    // one warp per row
        const int grow = rbk * HD_BM + rr;
        if (grow >= p.b) break;
        const float* xr = p.x_raw[d] + static_cast<size_t>(grow) * p.e;
        const float* gr = p.ws + L.dxn + (static_cast<size_t>(d) * p.b + grow) * p.e;
        float ss = 0.f, dd = 0.f;
        for (int c = lane * 4; c < p.e; c += 128) {
          const float4 v = *reinterpret_cast<const float4*>(xr + c);
          const float4 g = __ldcg(reinterpret_cast<const float4*>(gr + c));
          ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
          dd += v.x * g.x + v.y * g.y + v.z * g.z + v.w * g.w;
        }
        const float rn = sqrtf(warp_sum(ss));
        dd = warp_sum(dd);
        const float inv = 1.0f / (rn + eps);
        const float k = dd * inv * inv / fmaxf(rn, 1e-30f);
        float* o = p.dx_out[d] + static_cast<size_t>(grow) * p.e;
        for (int c = lane * 4; c < p.e; c += 128) {
          const float4 v = *reinterpret_cast<const float4*>(xr + c);
          const float4 g = __ldcg(reinterpret_cast<const float4*>(gr + c));
          *reinterpret_cast<float4*>(o + c) = make_float4(inv * g.x - k * v.x, inv * g.y - k * v.y, inv * g.z - k * v.z, inv * g.w - k * v.w);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ------------------------------------------------------------------------------------------------ host side
static int fill_common(const dc_head_args* a, HeadParams& p, HeadMaps& maps, int box_rows_y) {
  if (a == nullptr) return set_error("head: null args");
  if (a->b <= 0 || a->n <= 0 || a->e <= 0) return set_error("head: empty problem");
  if (a->e % 256 != 0 || a->e > 1024) return set_error("head: feature dim must be 256, 512, 768 or 1024");
  if (a->n_src < 1 || a->n_src > HD_MAX_SRC) return set_error("head: 1..8 feature sources");
  if (a->n_src > 1 && a->n != a->n_src * a->b) return set_error("head: with per-rank sources n must equal n_src * b");
  if (a->n_src > 1 && a->b % 256 != 0) return set_error("head: per-rank sources need b to be a multiple of 256");
  if (a->cross && a->n > a->b && a->b % 32 != 0) return set_error("head: gathered symmetric pairs need b to be a multiple of 32");
  if (a->row0 < 0 || a->row0 + a->b > a->n) return set_error("head: local rows outside the gathered range");
  if (a->ld % 8 != 0) return set_error("head: feature row stride must be a multiple of 8 elements");
  memset(&p, 0, sizeof(p));
  p.b = a->b; p.n = a->n; p.e = a->e; p.ld = a->ld; p.n_src = a->n_src; p.row0 = a->row0; p.cross = a->cross;
  p.rows_per_src = a->n_src == 1 ? ((a->n + 255) / 256 * 256 + 256) : a->b;     // one source: never crosses over
  for (int d = 0; d < 2; ++d) { p.x_off[d] = a->x_off[d]; p.y_off[d] = a->y_off[d]; }
  p.rb = (a->b + HD_BM - 1) / HD_BM;
  p.x_base = static_cast<const bf16*>(a->x_base);
  p.ws = a->ws;
  for (int d = 0; d < 2; ++d) {
    int rc = make_tmap_2d(&maps.x[d], static_cast<const bf16*>(a->x_base) + a->x_off[d], a->e, a->b, a->ld, 64, HD_BM);
    if (rc) return rc;
    for (int s = 0; s < a->n_src; ++s) {
      const long long rows = a->n_src == 1 ? a->n : a->b;
      rc = make_tmap_2d(&maps.y[d][s], static_cast<const bf16*>(a->y_src[s]) + a->y_off[d], a->e, rows, a->ld, 64, box_rows_y);
      if (rc) return rc;
    }
  }
  return 0;
}

}  // namespace dc

using namespace dc;

extern "C" size_t dc_head_workspace_floats(int b, int e) {
  if (b <= 0 || e <= 0) return 0;
  return static_cast<size_t>(HeadWs(b, e).total);
}

extern "C" int dc_head_layout(int b, int e, long long* offsets8) {
  if (b <= 0 || e <= 0 || offsets8 == nullptr) return set_error("head: bad layout query");
  const HeadWs L(b, e);
  offsets8[0] = L.out; offsets8[1] = L.lse; offsets8[2] = L.g; offsets8[3] = L.alab; offsets8[4] = L.ea;
  offsets8[5] = L.cnt; offsets8[6] = L.dxn; offsets8[7] = L.total;
  return 0;
}

static int head_prepare_impl(const float* const* feats, const float* eps, int n_feats, int b, int e, void* out_rows,
                             void* const* push, int n_push, long long push_row0, float* ws, const float* logit_scale,
                             float scale_max, dc_stream_t stream);

extern "C" int dc_head_prepare(const float* const* feats, const float* eps, int n_feats, int b, int e, void* out_rows,
                               float* ws, const float* logit_scale, float scale_max, dc_stream_t stream) {
  return head_prepare_impl(feats, eps, n_feats, b, e, out_rows, nullptr, 0, 0, ws, logit_scale, scale_max, stream);
}

extern "C" int dc_head_prepare_push(const float* const* feats, const float* eps, int n_feats, int b, int e, void* out_rows,
                                    void* const* peer_buffers, int n_peers, long long row0, float* ws,
                                    const float* logit_scale, float scale_max, dc_stream_t stream) {
  if (n_peers < 0 || n_peers > HD_MAX_SRC || (n_peers > 0 && peer_buffers == nullptr))
    return set_error("head_prepare_push: 0..8 peer buffers");
  return head_prepare_impl(feats, eps, n_feats, b, e, out_rows, peer_buffers, n_peers, row0, ws, logit_scale, scale_max, stream);
}

static int head_prepare_impl(const float* const* feats, const float* eps, int n_feats, int b, int e, void* out_rows,
                             void* const* push, int n_push, long long push_row0, float* ws, const float* logit_scale,
                             float scale_max, dc_stream_t stream) {
  if (feats == nullptr || eps == nullptr || out_rows == nullptr || ws == nullptr) return set_error("head_prepare: null argument");
  if (n_feats < 1 || n_feats > 8) return set_error("head_prepare: 1..8 features");
  if (e % 4 != 0) return set_error("head_prepare: feature dim must be a multiple of 4");
  PrepArgs a;
  memset(&a, 0, sizeof(a));
  for (int f = 0; f < n_feats; ++f) { a.x[f] = feats[f]; a.eps[f] = eps[f]; }
  a.n_feats = n_feats; a.b = b; a.e = e;
  a.out = static_cast<bf16*>(out_rows);
  a.ws = ws;
  for (int k = 0; k < n_push; ++k) a.push[k] = static_cast<bf16*>(push[k]);
  a.n_push = n_push;
  a.push_row0 = push_row0;
  a.logit_scale = logit_scale;
  a.scale_max = scale_max;
  const HeadWs L(b, e);
  const int rb = (b + HD_BM - 1) / HD_BM;
  a.zero_a0 = L.out; a.zero_a1 = L.out + 8;                 // the accumulated scalars (8..15 are written by one thread)
  a.zero_c0 = L.cnt; a.zero_c1 = L.cnt + 4 * rb;            // last-arriver counters (ints share the float workspace)
  a.zero_b0 = L.dxn; a.zero_b1 = L.total;                   // dX accumulators of the backward kernel
  int blocks = (b * n_feats * 32 + 255) / 256;
  if (blocks < 64) blocks = 64;
  head_prep_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  DC_CHECK_LAUNCH("head_prep");
  return 0;
}

extern "C" int dc_head_forward(const dc_head_args* a, dc_stream_t stream) {
  HeadParams p;
  HeadMaps maps;
  int rc = fill_common(a, p, maps, HF_BN);
  if (rc) return rc;
  for (int d = 0; d < 2; ++d) p.strips[d] = a->strips[d];
  p.ld_strip = a->ld_strip;
  if ((a->strips[0] || a->strips[1]) && (a->n % 4 != 0 || a->ld_strip % 4 != 0)) return set_error("head: strips need n and ld_strip multiples of 4");
  p.ntiles = (a->n + HF_BN - 1) / HF_BN;
  int js = sm_count() / (2 * p.rb);
  if (js < 1) js = 1;
  if (js > p.ntiles) js = p.ntiles;
  if (js > HD_MAX_JS) js = HD_MAX_JS;
  p.js = js;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HF_SMEM);
    if (e != cudaSuccess) return set_error_cuda("cudaFuncSetAttribute(head_fwd)", e);
    attr = true;
  }
  head_fwd_kernel<<<2 * p.rb * p.js, HD_THREADS, HF_SMEM, static_cast<cudaStream_t>(stream)>>>(maps, p);
  DC_CHECK_LAUNCH("head_fwd");
  return 0;
}

extern "C" int dc_head_backward(const dc_head_args* a, const float* g_own, const float* exch, const float* const* x_raw,
                                const float* eps2, float* const* dx_out, dc_stream_t stream) {
  if (g_own == nullptr || exch == nullptr || x_raw == nullptr || eps2 == nullptr || dx_out == nullptr)
    return set_error("head_backward: null argument");
  const int jt = a != nullptr && a->e <= 512 ? 128 : 64;
  HeadParams p;
  HeadMaps maps;
  int rc = fill_common(a, p, maps, jt);
  if (rc) return rc;
  p.exch = exch;
  p.g_own = g_own;
  for (int d = 0; d < 2; ++d) { p.x_raw[d] = x_raw[d]; p.eps[d] = eps2[d]; p.dx_out[d] = dx_out[d]; }
  p.jt = jt;
  p.es = a->e / HB_OCOLS;
  p.ntiles = (a->n + jt - 1) / jt;
  int js = sm_count() / (2 * p.rb * p.es);
  if (js < 1) js = 1;
  if (js > p.ntiles) js = p.ntiles;
  if (js > HD_MAX_JS) js = HD_MAX_JS;
  p.js = js;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(head_bwd_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, HB_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(head_bwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, HB_SMEM);
    if (e != cudaSuccess) return set_error_cuda("cudaFuncSetAttribute(head_bwd)", e);
    attr = true;
  }
  const int grid = 2 * p.rb * p.es * p.js;
  if (jt == 128) head_bwd_kernel<128><<<grid, HD_THREADS, HB_SMEM, static_cast<cudaStream_t>(stream)>>>(maps, p);
  else head_bwd_kernel<64><<<grid, HD_THREADS, HB_SMEM, static_cast<cudaStream_t>(stream)>>>(maps, p);
  DC_CHECK_LAUNCH("head_bwd");
  return 0;
}
