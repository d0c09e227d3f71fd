// The whole per-atom network of one 128-atom tile in ONE kernel: forward layers, energy, backward seed, backward layers and
// dE/dAEV, with every activation and gradient staying on the SM (tensor memory / shared memory).  Replaces, for networks with
// three hidden layers whose (compacted) input has at most 128 columns, the twelve per-layer GEMM launches of mlp_tcgen05.cu and
// their hi/lo activation round trips through HBM (3.5 GB per 50 000-atom water evaluation, 13.6 x the algorithmic bytes).
// Arithmetic: src/pytorch/BatchedNN.py:97-109 + BatchedNN.cpp:30-42 of the reference (see species_mlp.cuh).
#pragma once
#include <cuda_fp16.h>
#include <vector>
#include "common.cuh"

namespace nnpops {

class MlpChain {
public:
    struct SpeciesDesc {
        int d[4];               // padded widths: input, hidden 1..3 (multiples of 16; input the same for all species)
        int rowStart, rows;     // rows of this species in the species-sorted feature matrix
        const float* W[3];      // HOST, layer l: [M * d[l+1]][d[l]] (member-major rows, K-major)
        const float* bias[3];   // DEVICE, layer l: [M][d[l+1]]
        const float* w3;        // DEVICE, output layer: [M][d[3]]
    };
    // true when the fused kernel can run this network (see the limits in mlp_chain.cu)
    static bool eligible(int numSpecies, const SpeciesDesc* sp, int featureStride);
    MlpChain(int ensemble, int numSpecies, const SpeciesDesc* sp, const __half* featHi, const __half* featLo, int featureStride);
    ~MlpChain();
    MlpChain(const MlpChain&) = delete;
    MlpChain& operator=(const MlpChain&) = delete;
    // energyAcc += sum over rows and members of the network output without the last bias (double, device);
    // dX[rows][featureStride] = d(sum)/dX * outScale / seedScale-lift, i.e. the caller passes seedScale = lift / M and outScale = 1 / lift
    void launch(double* energyAcc, float* dX, float seedScale, float outScale, cudaStream_t stream);

private:
    struct Impl;
    Impl* impl_;
};

// Variant of the same kernel (mlp_chain2.cu, behind NNPOPS_CHAIN_V2=1): one accumulator per output column (the hi.lo cross terms are
// folded in by the tensor core's input scaling, tcgen05.mma scale-input-d), chunks of 128 output columns, one in-order MMA issuer, all
// epilogue warps on every chunk.  Same interface, limits and results; measured slower on B200 (0.65 vs 0.54 ms on the bench box: 1.8 x
// the mbarrier waits per chain on the issuing thread), kept as the starting point for a cta_group::2 version.
class MlpChain2 {
public:
    static bool eligible(int numSpecies, const MlpChain::SpeciesDesc* sp, int featureStride);
    MlpChain2(int ensemble, int numSpecies, const MlpChain::SpeciesDesc* sp, const __half* featHi, const __half* featLo, int featureStride);
    ~MlpChain2();
    MlpChain2(const MlpChain2&) = delete;
    MlpChain2& operator=(const MlpChain2&) = delete;
    void launch(double* energyAcc, float* dX, float seedScale, float outScale, cudaStream_t stream);

private:
    struct Impl;
    Impl* impl_;
};

}  // namespace nnpops
