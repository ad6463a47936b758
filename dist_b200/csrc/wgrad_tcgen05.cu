// distb200_gemm_wgrad on tcgen05 tensor cores - placeholder routing to the FFMA kernel until the MN-major kernel lands.
#include "common.cuh"

namespace distb200 {

int wgrad_tcgen05_launch(const distb200_wgrad_desc& d, cudaStream_t stream) { return wgrad_simt_launch(d, stream); }

}  // namespace distb200
