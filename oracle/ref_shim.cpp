/*
 * TEST INFRASTRUCTURE ONLY.  extern "C" shim around the UNMODIFIED reference CPU classes
 * (CpuANISymmetryFunctions, CpuCFConv, CpuCFConvNeighbors).  It is compiled together with the
 * reference sources where they lie under /root/reference (see oracle/Makefile) into
 * oracle/_ref/libnnpops_ref.so; no reference source is copied into this repository.
 * Used by tests/ to pin the oracle restatements and by bench.py's cpu_baseline / --impl reference leg.
 */
#include <vector>
#include "CpuANISymmetryFunctions.h"   // /root/reference/src/ani
#include "CpuCFConv.h"                 // /root/reference/src/schnet

extern "C" {

// AEV forward (+ optional backward when radial_grad != NULL) through the reference class.
int ref_ani(int n_atoms, int n_species, float rcr, float rca, int torchani, const int* species,
            int n_radial, const float* radial_fn, int n_angular, const float* angular_fn,
            const float* pos, const float* box, float* radial, float* angular,
            const float* radial_grad, const float* angular_grad, float* pos_grad) {
    std::vector<int> sp(species, species + n_atoms);
    std::vector<RadialFunction> rf;
    for (int k = 0; k < n_radial; k++) rf.push_back({radial_fn[2 * k], radial_fn[2 * k + 1]});
    std::vector<AngularFunction> af;
    for (int m = 0; m < n_angular; m++) af.push_back({angular_fn[4 * m], angular_fn[4 * m + 1], angular_fn[4 * m + 2], angular_fn[4 * m + 3]});
    CpuANISymmetryFunctions ani(n_atoms, n_species, rcr, rca, box != nullptr, sp, rf, af, torchani != 0);
    ani.computeSymmetryFunctions(pos, box, radial, angular);
    if (radial_grad) ani.backprop(radial_grad, angular_grad, pos_grad);
    return 0;
}

// CFConv forward (+ optional backward) through the reference classes.  n_pairs_out (optional) gets the
// size of the half neighbour list.
int ref_cfconv(int n_atoms, int width, int n_gauss, float cutoff, float gauss_width, int activation,
               const float* w1, const float* b1, const float* w2, const float* b2,
               const float* pos, const float* box, const float* input, float* output,
               const float* output_grad, float* input_grad, float* pos_grad, long long* n_pairs_out) {
    bool periodic = box != nullptr;
    CpuCFConvNeighbors nb(n_atoms, cutoff, periodic);
    nb.build(pos, box);
    if (n_pairs_out) {
        long long n = 0;
        for (auto& row : nb.getNeighbors()) n += (long long)row.size();
        *n_pairs_out = n;
    }
    CpuCFConv conv(n_atoms, width, n_gauss, cutoff, periodic, gauss_width,
                   activation == 0 ? CFConv::ShiftedSoftplus : CFConv::Tanh, w1, b1, w2, b2);
    conv.compute(nb, pos, box, input, output);
    if (output_grad) conv.backprop(nb, pos, box, input, output_grad, input_grad, pos_grad);
    return 0;
}

}
