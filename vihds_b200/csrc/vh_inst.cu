// One translation unit per (direction, dtype, model): compiled many times from this single source with
//   -DVH_DIR=0|1 (forward | reverse)  -DVH_REAL=float|double  -DVH_VER=1|2 -DVH_EXT=0|1|2 (plain | relay | degrader) -DVH_DYN=0|1
//   -DVH_GROWTH=0|4|5|6 (growth-only family: species count; 4 auto, 5 inducer, 6 prpr)  -DVH_FN=<symbol>
// so that the 5 solver instantiations of each of the 24 combinations build in parallel (see Makefile).
#include "vh_launch.cuh"

namespace vh {
#if defined(VH_GROWTH) && VH_GROWTH
typedef GrowthModel<VH_REAL, VH_GROWTH, (VH_DYN != 0)> ModelT;  // VH_GROWTH = number of species (4: auto, 5: inducer, 6: prpr)
#else
typedef DrModel<VH_REAL, VH_VER, VH_EXT, (VH_DYN != 0)> ModelT;
#endif
#if VH_DIR == 0
int VH_FN(const vh_problem* p, const vh_fwd_io* io, cudaStream_t stream) { return launch_fwd_model<ModelT>(p, io, stream); }
#else
int VH_FN(const vh_problem* p, const vh_bwd_io* io, cudaStream_t stream) { return launch_bwd_model<ModelT>(p, io, stream); }
#endif
}  // namespace vh
