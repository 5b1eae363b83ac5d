// Kernel launch helper: every launch carries the programmatic-stream-serialization attribute (PDL) so the next kernel's
// prologue overlaps this kernel's body (see ptx.cuh pdl_*). Works under stream capture (the edge becomes a programmatic
// graph dependency). B200SEG_PDL=0 in the environment falls back to plain serialized launches (A/B measurements).
#pragma once
#include <cuda_runtime.h>
#include <cstdlib>

namespace b200seg {

inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200SEG_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                            Args... args) {
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

}  // namespace b200seg
