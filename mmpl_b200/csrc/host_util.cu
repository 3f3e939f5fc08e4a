#include "host_util.h"

#include <cstdlib>

#include <cstring>
#include <mutex>
#include <unordered_map>

namespace mmpl {

static thread_local char g_err[512] = "";

char* last_error_buf() { return g_err; }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

namespace {

struct MapKey {
  const void* ptr;
  uint64_t rows, cols, ld;
  uint32_t box_rows;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    uint64_t h = reinterpret_cast<uint64_t>(k.ptr);
    h = h * 0x9E3779B97F4A7C15ull + k.rows;
    h = h * 0x9E3779B97F4A7C15ull + k.cols;
    h = h * 0x9E3779B97F4A7C15ull + k.ld;
    h = h * 0x9E3779B97F4A7C15ull + k.box_rows;
    return static_cast<size_t>(h ^ (h >> 29));
  }
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

std::mutex g_mu;
std::unordered_map<MapKey, CUtensorMap*, MapKeyHash> g_maps;
EncodeTiledFn g_encode = nullptr;

// The driver entry point is resolved at run time so that the library has no link-time
// dependency on libcuda (it must load on a machine without a GPU driver).
EncodeTiledFn resolve_encode() {
  if (g_encode) return g_encode;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) {
    set_error("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: %s",
              cudaGetErrorString(e));
    return nullptr;
  }
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  return g_encode;
}

}  // namespace

const CUtensorMap* get_tensor_map_bf16(const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld,
                                       uint32_t box_rows) {
  std::lock_guard<std::mutex> lock(g_mu);
  MapKey key{ptr, rows, cols, ld, box_rows};
  auto it = g_maps.find(key);
  if (it != g_maps.end()) return it->second;

  EncodeTiledFn encode = resolve_encode();
  if (!encode) return nullptr;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld * 2) % 16 != 0) {
    set_error("tensor map: base pointer and row pitch must be 16-byte aligned (ptr=%p ld=%llu)",
              ptr, (unsigned long long)ld);
    return nullptr;
  }
  if (box_rows == 0 || box_rows > 256) {
    set_error("tensor map: box_rows=%u out of range", box_rows);
    return nullptr;
  }
  CUtensorMap* m = new CUtensorMap;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estride[2] = {1, 1};
  CUresult r = encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride,
                      box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu ld=%llu box=%u)",
              (int)r, (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld,
              box_rows);
    delete m;
    return nullptr;
  }
  g_maps.emplace(key, m);
  return m;
}

void clear_tensor_map_cache() {
  std::lock_guard<std::mutex> lock(g_mu);
  for (auto& kv : g_maps) delete kv.second;
  g_maps.clear();
}

static int64_t g_workspace_generation = 0;
int64_t workspace_generation() { return g_workspace_generation; }
void bump_workspace_generation() { ++g_workspace_generation; }

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

bool pdl_enabled() {
  static const bool on = std::getenv("MMPL_B200_NO_PDL") == nullptr;
  return on;
}

bool device_is_sm100() {
  static int cached = -1;
  if (cached < 0) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess)
      return false;
    cached = (major == 10) ? 1 : 0;
  }
  return cached == 1;
}

}  // namespace mmpl
