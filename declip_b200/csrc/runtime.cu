// declip_b200 — library runtime: error slot, device check, driver entry points, TMA descriptors.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <mutex>
#include <stdlib.h>
#include "internal.h"

namespace dc {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
static int g_sm_count = 0;
static int g_device = -1;

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;

int set_error(const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return -1;
}
int set_error_cuda(const char* what, cudaError_t e) {
  snprintf(g_err, sizeof(g_err), "%s: %s (%d)", what, cudaGetErrorString(e), static_cast<int>(e));
  return static_cast<int>(e) ? static_cast<int>(e) : -1;
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
static std::atomic<int> g_sm_reserve{0};
int sm_count() {
  const int n = (g_sm_count > 0 ? g_sm_count : 148) - g_sm_reserve.load(std::memory_order_relaxed);
  return n < 2 ? 2 : n;
}
int set_sm_reserve(int n) { return g_sm_reserve.exchange(n < 0 ? 0 : n); }

int make_tmap_2d(CUtensorMap* tm, const void* ptr, long long inner, long long outer, long long ld, int box_inner,
                 int box_outer, int swizzle_bytes) {
  if (g_encode == nullptr) return set_error("dc_init() was not called (no cuTensorMapEncodeTiled entry point)");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return set_error("tensor map: base pointer must be 16-byte aligned");
  if ((ld * 2) % 16 != 0) return set_error("tensor map: row stride must be a multiple of 16 bytes");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(outer)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_inner), static_cast<cuuint32_t>(box_outer)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_err, sizeof(g_err),
             "cuTensorMapEncodeTiled failed (%d): ptr=%p inner=%lld outer=%lld ld=%lld box=%dx%d", static_cast<int>(r),
             ptr, inner, outer, ld, box_inner, box_outer);
    return -2;
  }
  return 0;
}

bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("DC_PDL"); return !(e != nullptr && e[0] == '0'); }();
  return on;
}

int make_tmap_nd(CUtensorMap* tm, const void* ptr, int rank, const long long* dims, const long long* strides_bytes,
                 const int* box, int swizzle_bytes) {
  if (g_encode == nullptr) return set_error("dc_init() was not called (no cuTensorMapEncodeTiled entry point)");
  if (rank < 2 || rank > 5) return set_error("tensor map: rank must be in [2, 5]");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return set_error("tensor map: base pointer must be 16-byte aligned");
  cuuint64_t d[5];
  cuuint64_t s[4];
  cuuint32_t b[5];
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  for (int i = 0; i < rank; ++i) { d[i] = static_cast<cuuint64_t>(dims[i]); b[i] = static_cast<cuuint32_t>(box[i]); }
  for (int i = 0; i + 1 < rank; ++i) {
    if (strides_bytes[i] % 16 != 0) return set_error("tensor map: strides must be multiples of 16 bytes");
    s[i] = static_cast<cuuint64_t>(strides_bytes[i]);
  }
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(ptr), d, s,
                        b, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_err, sizeof(g_err), "cuTensorMapEncodeTiled(rank %d) failed (%d): dims0=%lld box0=%d", rank,
             static_cast<int>(r), dims[0], box[0]);
    return -2;
  }
  return 0;
}

int make_tmap_4d(CUtensorMap* tm, const void* ptr, const long long dims[4], const long long strides_bytes[3],
                 const int box[4]) {
  return make_tmap_nd(tm, ptr, 4, dims, strides_bytes, box, 128);
}

}  // namespace dc

extern "C" {

int dc_version(void) { return 100; }

int dc_gemm_choose_splits(int tiles, int total_kb, int workers) { return dc::choose_splits(tiles, total_kb, workers); }

const char* dc_last_error(void) { return dc::g_err; }

long long dc_launch_count(void) { return dc::g_launches.load(); }

int dc_sm_count(void) { return dc::sm_count(); }

int dc_set_sm_reserve(int n) { return dc::set_sm_reserve(n); }

int dc_init(int device) {
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return dc::set_error_cuda("cudaSetDevice", e);
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return dc::set_error_cuda("cudaGetDeviceProperties", e);
  if (prop.major != 10) {
    char buf[256];
    snprintf(buf, sizeof(buf), "declip_b200 requires an sm_100a (compute capability 10.x) device, found %d.%d (%s); "
             "there is no fallback path", prop.major, prop.minor, prop.name);
    return dc::set_error(buf);
  }
  dc::g_sm_count = prop.multiProcessorCount;
  dc::g_device = device;
  if (dc::g_encode == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || fn == nullptr) return dc::set_error_cuda("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled)", e);
    dc::g_encode = reinterpret_cast<dc::PFN_encodeTiled>(fn);
  }
  return 0;
}

}  // extern "C"
