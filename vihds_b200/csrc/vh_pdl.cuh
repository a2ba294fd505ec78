// Programmatic dependent launch (sm_90+): a kernel launched with the "programmatic stream serialization" attribute may
// start while its predecessor on the stream is still running; it runs whatever does not depend on the predecessor, then
// `pdl_wait()` blocks until the predecessor grid has completed and its writes are visible.  A predecessor lets the
// dependent grid start early with `pdl_trigger()` (otherwise: when it exits).  In a latency-bound chain of small kernels
// this hides the launch latency and the dependent kernel's prologue (weight staging, optimiser-state loads).
//
// Rules kept by every user in this library:
//   * only kernels that CONTAIN pdl_wait() are launched with the attribute (a kernel without it could finish -- and
//     release ITS dependents -- before its own predecessor has completed);
//   * a kernel triggers only once every one of its CTAs is running (the call sits at the top of the kernel), so the
//     dependent grid can never occupy resources a not-yet-scheduled CTA of the primary is waiting for;
//   * nothing is written to global memory before pdl_wait().
// Both instructions are no-ops in a launch without the attribute / without dependents.  VIHDS_PDL=0 (read per call) turns
// the attribute off.  Under stream capture the attribute becomes a programmatic edge of the graph (CUDA >= 12.3).
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <utility>

namespace vh {

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline bool pdl_enabled() {
  const char* m = getenv("VIHDS_PDL");
  return !(m && *m == '0');
}

template <class... KArgs, class... Args>
inline cudaError_t launch_maybe_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl,
                                    Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = (pdl && pdl_enabled()) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace vh
