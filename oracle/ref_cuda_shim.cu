/*
 * TEST / MEASUREMENT INFRASTRUCTURE ONLY.  extern "C" shim around the UNMODIFIED reference CUDA classes
 * (CudaANISymmetryFunctions, CudaCFConvNeighbors, CudaCFConv), compiled together with the reference sources where they lie
 * under /root/reference (oracle/Makefile, target refcuda) for sm_100a into oracle/_ref/libnnpops_ref_cuda.so.  No reference
 * source is copied into this repository.  Used by bench.py's "gpu_comparator" key (the reference's own kernels on the same B200,
 * same inputs, timed with CUDA events) and by tests/test_reference_cuda_gpu.py (our kernels against the reference's on the GPU).
 */
#include <vector>
#include "CudaANISymmetryFunctions.h"   // /root/reference/src/ani
#include "CudaCFConv.h"                 // /root/reference/src/schnet


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

// SchNet continuous-filter convolution (src/schnet/CudaCFConv.h:37-115 neighbour list, :138-190 layer).  The reference's list is an
// N x N table in managed memory (24 N^2 bytes, int arithmetic): N <= 16 000.
void* refcuda_cfconv_neighbors_create(int n_atoms, float cutoff, int periodic) {
    try { return new CudaCFConvNeighbors(n_atoms, cutoff, periodic != 0); } catch (...) { return nullptr; }
}
void refcuda_cfconv_neighbors_destroy(void* h) { delete static_cast<CudaCFConvNeighbors*>(h); }
int refcuda_cfconv_neighbors_build(void* h, const float* pos, const float* box) {
    try { static_cast<CudaCFConvNeighbors*>(h)->build(pos, box); return 0; } catch (...) { return 1; }
}
// number of (undirected) pairs of the last build; synchronises the device
int refcuda_cfconv_neighbors_count(void* h) {
    cudaDeviceSynchronize();
    return *static_cast<CudaCFConvNeighbors*>(h)->getNeighborCount();
}
void* refcuda_cfconv_create(int n_atoms, int width, int n_gaussians, float cutoff, int periodic, float gaussian_width, int activation,
                            const float* w1, const float* b1, const float* w2, const float* b2) {
    try {
        return new CudaCFConv(n_atoms, width, n_gaussians, cutoff, periodic != 0, gaussian_width,
                              activation == 0 ? CFConv::ShiftedSoftplus : CFConv::Tanh, w1, b1, w2, b2);
    } catch (...) { return nullptr; }
}
void refcuda_cfconv_destroy(void* h) { delete static_cast<CudaCFConv*>(h); }
int refcuda_cfconv_compute(void* h, void* nb, const float* pos, const float* box, const float* input, float* output) {
    try { static_cast<CudaCFConv*>(h)->compute(*static_cast<CudaCFConvNeighbors*>(nb), pos, box, input, output); return 0; } catch (...) { return 1; }
}
int refcuda_cfconv_backprop(void* h, void* nb, const float* pos, const float* box, const float* input, const float* output_deriv,
                            float* input_deriv, float* position_deriv) {
    try {
        static_cast<CudaCFConv*>(h)->backprop(*static_cast<CudaCFConvNeighbors*>(nb), pos, box, input, output_deriv, input_deriv, position_deriv);
        return 0;
    } catch (...) { return 1; }
}

}
