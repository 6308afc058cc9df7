// placeholder until the DMMA kernels land
#include "common.cuh"
#include "kernels.h"
namespace pyvb {
bool dmma_supported(int, int) { return false; }
cudaError_t launch_zstep_dmma(long long, int, int, const double *, long long, const double *, int, const double *,
                              const double *, double *, double *, double *, double *, double *, cudaStream_t) {
    return cudaErrorNotSupported;
}
int stats_dmma_nchunks(long long, int, int) { return 1; }
cudaError_t launch_stats_dmma(long long, int, int, const double *, long long, const double *, const double *,
                              double *, int, cudaStream_t) {
    return cudaErrorNotSupported;
}
}  // namespace pyvb
