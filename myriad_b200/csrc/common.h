// Host-side helpers shared by all translation units of libmyriad_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/myriad_b200.h"

namespace myr {

void set_error(const char* fmt, ...);
// Profiling aid (myr_gemm_set_trace): next 148 x 6 int64 slot of the trace buffer, or null when tracing is off.
long long* next_trace_slot();
void set_trace_buffer(void* buf);
void count_launch();
int sm_count();

// cuTensorMapEncodeTiled obtained through the runtime (no link-time libcuda dependency).
// dims/strides innermost first; strides in BYTES for dims 1..rank-1; box in elements. fp16, SWIZZLE_128B.
int make_tmap_f16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box);

#define MYR_CHECK_ARG(cond, ...)          \
  do {                                    \
    if (!(cond)) {                        \
      myr::set_error(__VA_ARGS__);        \
      return MYR_ERR_INVALID;             \
    }                                     \
  } while (0)

#define MYR_CHECK_CUDA(expr)                                                                   \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      myr::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return MYR_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

// every kernel launch in the library is followed by this macro: it also feeds myr_launch_count()
#define MYR_CHECK_LAUNCH()                \
  do {                                    \
    myr::count_launch();                  \
    MYR_CHECK_CUDA(cudaGetLastError());   \
  } while (0)

// Programmatic dependent launch switch for the whole library (myr_set_pdl / env MYR_PDL=0).
bool pdl_enabled();
void set_pdl(int v);

// Launch through cudaLaunchKernelEx; with `pdl` (and the library switch on) the kernel carries the programmatic stream
// serialization attribute. Only kernels that execute pdl_wait() before their first dependent access may pass pdl = true.
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_kernel_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                                bool pdl, int cluster_x, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (pdl && pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = (unsigned)cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl,
                                        Args... args) {
  return launch_kernel_cluster(kernel, grid, block, smem, stream, pdl, 1, args...);
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace myr
