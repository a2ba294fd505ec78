// Instantiates the bwd kernels of the double-receiver family for double (one translation unit per dtype x direction so
// the 30 (model x solver) instantiations of each compile in parallel).
#include "vh_launch.cuh"
namespace vh {
int launch_bwd_f64(const vh_problem* p, const vh_bwd_io* io, cudaStream_t stream) { return launch_bwd<double>(p, io, stream); }
}  // namespace vh
