// Host-side helpers: error reporting across the C ABI and TMA tensor-map creation.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "mmpl_b200.h"

#include <cstdarg>
#include <cstdio>
#include <utility>

namespace mmpl {

// Last error message, readable through mmpl_last_error().
char* last_error_buf();
void set_error(const char* fmt, ...);

#define MMPL_CHECK(cond, code, ...) \
  do {                              \
    if (!(cond)) {                  \
      ::mmpl::set_error(__VA_ARGS__); \
      return (code);                \
    }                               \
  } while (0)

#define MMPL_CUDA(expr)                                                              \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      ::mmpl::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                        __LINE__);                                                   \
      return MMPL_ERR_CUDA;                                                          \
    }                                                                                \
  } while (0)

// 2-D bf16 tensor map over a row-major [rows, cols] view with leading dimension `ld` (elements),
// box = [box_rows, 64 cols] with 128-byte swizzle. Out-of-bounds elements read as zero.
// Maps are cached by (ptr, rows, cols, ld, box_rows); returns nullptr after set_error on failure.
const CUtensorMap* get_tensor_map_bf16(const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld,
                                       uint32_t box_rows);
void clear_tensor_map_cache();

// A failed launch is picked up by the MMPL_CUDA(cudaGetLastError()) that follows every launch site.
#define MMPL_CUDA_LAUNCH(expr) (void)(expr)

// Launch with programmatic stream serialization (PDL, see ptx.cuh) unless MMPL_B200_NO_PDL is set in the environment.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Bumped whenever a process-level device workspace (attention partials, merge counters, stream-K slots) is reallocated:
// anything that recorded device pointers of the library - a captured CUDA graph of a launch sequence - is stale when
// the value it saw differs from the current one (mmpl_workspace_generation()).
int64_t workspace_generation();
void bump_workspace_generation();

// Number of SMs on the current device (cached) and sm_100 check.
int sm_count();
bool device_is_sm100();

}  // namespace mmpl
