// Programmatic dependent launch (PDL): a kernel launched through launch_pdl may start while its predecessor in the stream
// is still running; it runs its prologue (barrier init, TMEM allocation, tensor-map prefetch) and then blocks in
// ptx::pdl_wait() until the predecessor has completed and flushed its memory.  Only kernels that CALL ptx::pdl_wait()
// before their first access to global memory may be launched this way.  Inside a CUDA graph capture the attribute becomes
// a programmatic edge.  GLC_NO_PDL=1 falls back to fully serialised launches.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <utility>

namespace glc {

inline bool pdl_enabled() {
  static const bool on = getenv("GLC_NO_PDL") == nullptr;
  return on;
}

template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
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
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

}  // namespace glc
