// Programmatic dependent launch (sm_90+): a kernel launched with the attribute may become resident while its predecessor in the stream is
// still draining; it must execute pdl_wait() before touching anything the predecessor writes (or writing anything it reads). Kernels on the
// LLM decode chain use this so that launch latency, block scheduling and constant prefetch overlap the previous kernel's tail.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>

namespace vg {

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline bool pdl_enabled() {
  static const int v = [] {
    const char* e = getenv("VG_PDL");
    return e ? atoi(e) : 1;
  }();
  return v != 0;
}

inline bool pdl_all_enabled() {  // measured on B200 (profiles/r01_pdl.md): PDL on every launch gains nothing outside the decode chain -> opt-in
  static const int v = [] {
    const char* e = getenv("VG_PDL_ALL");
    return e ? atoi(e) : 0;
  }();
  return v != 0 && pdl_enabled();
}

template <bool kAlways = true, typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = (kAlways ? pdl_enabled() : pdl_all_enabled()) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Ordinary launch site (attribute only with VG_PDL_ALL=1); the kernels still execute pdl_wait(), which is a no-op without the attribute.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  return launch_pdl<false>(kernel, grid, block, smem, stream, args...);
}

}  // namespace vg

