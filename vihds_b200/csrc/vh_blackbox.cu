// dr_blackbox (models/dr_blackbox.py) kernels -- placeholder translation unit until the MLP right-hand side lands.
#include <cuda_runtime.h>

#include "vh_dispatch.cuh"

namespace vh {
void set_error(const char* fmt, ...);
int launch_bb_fwd(const vh_problem*, const vh_fwd_io*, cudaStream_t) {
  set_error("dr_blackbox kernels are not built into this library yet");
  return VH_ERR_UNSUPPORTED;
}
int launch_bb_bwd(const vh_problem*, const vh_bwd_io*, cudaStream_t) {
  set_error("dr_blackbox kernels are not built into this library yet");
  return VH_ERR_UNSUPPORTED;
}
}  // namespace vh
