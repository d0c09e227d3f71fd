// Tensor-core (tcgen05 / TMEM / TMA) GEMM for the species-grouped MLP.  Placeholder until the kernel lands: selecting
// MlpImpl::Tcgen05 fails loudly rather than silently falling back.
#include "species_mlp.cuh"

namespace nnpops {

void launch_gemm_tcgen05(const GemmArgs&, cudaStream_t) {
    throw std::runtime_error("nnpops_b200: the tcgen05 MLP GEMM is not available in this build");
}

}  // namespace nnpops
