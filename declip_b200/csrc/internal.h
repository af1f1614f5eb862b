// declip_b200 — host-side internals shared by the translation units (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <utility>
#include "../../include/declip_b200.h"

namespace dc {

int set_error(const char* msg);
int set_error_cuda(const char* what, cudaError_t e);
void count_launch();
int sm_count();            // SMs the persistent kernels may fill: device SMs minus dc_set_sm_reserve()
int set_sm_reserve(int n);

// 2-D bf16 tensor map, 128-byte swizzle.  `inner` = contiguous extent (elements), `outer` = rows,
// `ld` = row stride (elements); box = {box_inner (must be 64 -> 128 B), box_outer <= 256}.
// OOB elements are zero-filled by the TMA unit, so ragged M/N/K tails need no special casing.
int make_tmap_2d(CUtensorMap* tm, const void* ptr, long long inner, long long outer, long long ld, int box_inner,
                 int box_outer, int swizzle_bytes = 128);

// 4-D bf16 tensor map, 128-byte swizzle: dims[0] is the contiguous extent (elements), strides_bytes[i] is the byte stride
// of dims[i + 1]; box[0] must be 64.
int make_tmap_4d(CUtensorMap* tm, const void* ptr, const long long dims[4], const long long strides_bytes[3],
                 const int box[4]);

// rank-N (2..5) bf16 tensor map; swizzle_bytes 128 (box[0] = 64 elements) or 64 (box[0] = 32 elements)
int make_tmap_nd(CUtensorMap* tm, const void* ptr, int rank, const long long* dims, const long long* strides_bytes,
                 const int* box, int swizzle_bytes);

// tcgen05 attention core (attention_tc.cu); return DC_ATTN_TC_UNSUPPORTED when the shape is outside its envelope
constexpr int DC_ATTN_TC_UNSUPPORTED = -100;
int attention_tc_fwd(const void* qkv, void* out, float* lse, int batch, int L, int heads, int causal, cudaStream_t st);
// dbias (optional, [3 D]): Q slice always; V slice only when dbias_v != 0 (it equals colsum(dout), which a caller that
// produces dout with a GEMM gets from that GEMM's fused column sum); the K slice is identically zero and left untouched.
int attention_tc_bwd(const void* qkv, const void* dout, const float* lse, void* dqkv, float* dbias, int dbias_v,
                     int batch, int L, int heads, int causal, cudaStream_t st);
bool attention_tc_supported(int batch, int L, int heads);
bool attention_tc_enabled();

// Split-K factor for an fp32-atomic GEMM of `tiles` output tiles and `total_kb` k-blocks on `workers` persistent CTAs
// (clusters): minimises waves x (k-blocks per split + per-item overhead) — `tiles * splits` just above a multiple of
// `workers` costs a whole extra wave, which a fill-the-machine-twice rule hits for most weight-gradient shapes.
inline int choose_splits(int tiles, int total_kb, int workers) {
  if (tiles >= workers) return 1;
  int max_splits = (total_kb + 3) / 4;          // at least 4 k-blocks per split
  if (max_splits > 64) max_splits = 64;
  int best = 1;
  long long best_cost = -1;
  for (int s = 1; s <= max_splits; ++s) {
    const int kbps = (total_kb + s - 1) / s;
    const int eff = (total_kb + kbps - 1) / kbps;                 // splits that actually get work
    const long long waves = (static_cast<long long>(tiles) * eff + workers - 1) / workers;
    const long long cost = waves * (kbps + 3);                   // +3: pipeline fill / accumulator hand-off per item
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = s; }
  }
  return best;
}

// Launches `kern` with programmatic stream serialization (see common.cuh: pdl_wait) unless DC_PDL=0.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// launchers implemented in the .cu files, used by the composite encoders
int gemm_bf16(const dc_gemm_args& a, cudaStream_t stream);

}  // namespace dc

#define DC_CHECK_LAUNCH(what)                                         \
  do {                                                                \
    cudaError_t e__ = cudaGetLastError();                             \
    if (e__ != cudaSuccess) return dc::set_error_cuda(what, e__);     \
    dc::count_launch();                                               \
  } while (0)
