/*
 * TEST / MEASUREMENT INFRASTRUCTURE ONLY.  extern "C" shim around the UNMODIFIED reference CUDA classes
 * (CudaANISymmetryFunctions), compiled together with the reference sources where they lie
 * under /root/reference (oracle/Makefile, target refcuda) for sm_100a into oracle/_ref/libnnpops_ref_cuda.so.  No reference
 * source is copied into this repository.  Used by bench.py's "gpu_comparator" key (the reference's own kernels on the same B200,
 * same inputs, timed with CUDA events) and by tests/test_reference_cuda_gpu.py (our kernels against the reference's on the GPU).
 */
#include <vector>
#include "CudaANISymmetryFunctions.h"   // /root/reference/src/ani


extern "C" {

void* refcuda_ani_create(int n_atoms, int n_species, float rcr, float rca, int periodic, const int* species, int n_radial,
                         const float* radial_fn, int n_angular, const float* angular_fn, int torchani) {
    try {
        std::vector<int> sp(species, species + n_atoms);
        std::vector<RadialFunction> rf;
        for (int k = 0; k < n_radial; k++) rf.push_back({radial_fn[2 * k], radial_fn[2 * k + 1]});
        std::vector<AngularFunction> af;
        for (int m = 0; m < n_angular; m++) af.push_back({angular_fn[4 * m], angular_fn[4 * m + 1], angular_fn[4 * m + 2], angular_fn[4 * m + 3]});
        return new CudaANISymmetryFunctions(n_atoms, n_species, rcr, rca, periodic != 0, sp, rf, af, torchani != 0);
    } catch (...) { return nullptr; }
}
void refcuda_ani_destroy(void* h) { delete static_cast<CudaANISymmetryFunctions*>(h); }
// all pointers may be host or device memory (CudaANISymmetryFunctions.h:62-64)
int refcuda_ani_forward(void* h, const float* pos, const float* box, float* radial, float* angular) {
    try { static_cast<CudaANISymmetryFunctions*>(h)->computeSymmetryFunctions(pos, box, radial, angular); return 0; } catch (...) { return 1; }
}
int refcuda_ani_backward(void* h, const float* radial_grad, const float* angular_grad, float* pos_grad) {
    try { static_cast<CudaANISymmetryFunctions*>(h)->backprop(radial_grad, angular_grad, pos_grad); return 0; } catch (...) { return 1; }
}

}
