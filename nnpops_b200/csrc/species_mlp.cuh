// Species-grouped per-atom MLP ("BatchedNN" of the reference, src/pytorch/BatchedNN.py:41-111 + BatchedNN.cpp:30-42).
//
// The reference replicates the network weights PER ATOM (weights[1, N, M, out, in]) and streams them through a batched
// mat-vec; here atoms are grouped by species once (the species of a Holder never change) and every layer becomes a real GEMM
//   Z[n_s x (M*out)] = A[n_s x in] . W^T     (layer 0: all M ensemble members side by side; deeper layers: one GEMM per member)
// with bias + CELU(0.1) fused in the epilogue.  Backward is the transposed chain with celu' recovered from the saved
// activations (celu'(z) = 1 for a > 0 else a/alpha + 1), giving dE/dAEV; no parameter gradients (BatchedNN.cpp:40).
// Zero padding of layer widths to multiples of 32 is exact (celu(0) = 0), exactly like the reference's own padding to
// the per-layer maximum (BatchedNN.py:71-83).
#pragma once
#include <cuda_fp16.h>
#include <memory>
#include <vector>
#include "common.cuh"
#include "mlp_chain.cuh"

namespace nnpops {

constexpr int kMlpPad = 32;   // layer widths are zero-padded to multiples of 32 (ANI-2x widths already are)
constexpr float kCeluAlpha = 0.1f;

enum class MlpImpl : int { Simt = 0, Tcgen05 = 1 };

class SpeciesMlp {
public:
    // dims[s][l], l = 0..numLayers: layer l of species s maps dims[s][l] -> dims[s][l+1]; dims[s][0] = numFeatures for all s and
    // dims[s][numLayers] = 1.  params: for s, for member e, for layer l: W (out x in, row-major) followed by b (out).
    // rowStart[s]..rowStart[s+1]: rows of species s in the species-sorted feature matrix.
    SpeciesMlp(int numSpecies, int ensemble, int numLayers, const int* dims, const float* params, const int* rowStart,
               int featureStride);
    ~SpeciesMlp();
    SpeciesMlp(const SpeciesMlp&) = delete;
    SpeciesMlp& operator=(const SpeciesMlp&) = delete;

    void setImpl(MlpImpl impl);
    // features: device [rows][featureStride] fp32 (species-sorted rows, columns >= numFeatures must be zero);
    // energy: device float[1] = (1/M) sum_atoms sum_members E.
    void forward(const float* features, float* energy, cudaStream_t stream);
    // featureGrad: device [rows][featureStride] <- dE/dfeatures (same row order); uses the activations of the last forward
    void backward(float* featureGrad, cudaStream_t stream);
    // Tensor-core path with the fused layer-chain kernel (mlp_chain.cu): energy and dE/dfeatures of one evaluation in one launch.
    // fused() says whether the network shape allows it (else call forward + backward); NNPOPS_NO_CHAIN=1 switches it off.
    bool fused() const { return chain_ != nullptr || chain2_ != nullptr; }
    void forwardBackward(float* energy, float* featureGrad, cudaStream_t stream);

    // tensor-core path: the feature matrix as fp16 hi/lo pairs [rows][featureStride]; the AEV kernels write it directly and
    // forward() is then called with features == nullptr
    __half* featHi() { return featHi_; }
    __half* featLo() { return featLo_; }
    bool tensorCore() const { return impl_ == MlpImpl::Tcgen05; }
    int numRows() const { return rows_; }
    double flopsForward() const { return flopsFwd_; }   // algorithmic (un-padded) flops of one forward pass

private:
    struct Layer {
        int in, out;         // true sizes
        int inP, outP;       // padded sizes
        float* W = nullptr;  // [M*outP][inP]          (K-major for the forward GEMM)
        float* Wt = nullptr; // layer 0: [inP][M*outP]; deeper: [M][inP][outP]   (K-major for the backward GEMM)
        float* b = nullptr;  // [M*outP]
        __half *Whi = nullptr, *Wlo = nullptr, *Wthi = nullptr, *Wtlo = nullptr;   // fp16 hi/lo pairs of W and Wt (tensor-core path)
    };
    int S_, M_, L_, rows_, featStride_, featP_;
    std::vector<std::vector<Layer>> layers_;   // [S][L]
    std::vector<int> rowStart_;
    std::vector<float*> act_;    // [L-1] activation buffers  [rows][M*maxOutP(l)]
    std::vector<float*> dz_;     // [L-1] gradient buffers of the same shapes
    std::vector<int> width_;     // [L-1] M*maxOutP(l)
    // tensor-core path: every matrix lives as an fp16 hi/lo pair
    std::vector<__half*> actHi_, actLo_, dzHi_, dzLo_;
    __half *featHi_ = nullptr, *featLo_ = nullptr;
    void forwardTc(const float* features, float* energy, cudaStream_t stream);
    void backwardTc(float* featureGrad, cudaStream_t stream);
    void forwardRowsTc(int s, int r0, int nr, int w0, cudaStream_t stream);
    void backwardRowsTc(int s, int r0, int nr, int w0, float* featureGrad, cudaStream_t stream);
    std::unique_ptr<MlpChain> chain_;     // the fused layer-chain kernel (default)
    std::unique_ptr<MlpChain2> chain2_;   // its single-accumulator variant (NNPOPS_CHAIN_V2=1)
    std::vector<std::vector<std::vector<float>>> hW_;   // host copy of the padded weights [S][L] until setImpl has built its operands
    double* energyAcc_ = nullptr;
    double energyBias_ = 0;      // sum over atoms and members of the last-layer bias
    MlpImpl impl_ = MlpImpl::Simt;
    double flopsFwd_ = 0;
    bool haveForward_ = false;
};

// C[m, n] (+epilogue) = sum_k A[m, k] * B[n, k];  batched over blockIdx.z with column/row offsets.  fp32 SIMT version.
struct GemmArgs {
    const float* A; int lda; int aBatchCols;     // A + batch*aBatchCols
    const float* B; int ldb; long long bBatch;   // B + batch*bBatch (elements)
    float* C; int ldc; int cBatchCols;
    const float* bias; int biasBatch;            // epilogue 1: bias[batch*biasBatch + n]
    const float* act; int ldact; int actBatchCols;   // epilogue 2: multiply by celu'(act[m, batch*actBatchCols + n])
    int M, N, K, batch;
    int epilogue;                                // 0 plain, 1 bias + celu, 2 times celu'(act)
};
void launch_gemm_simt(const GemmArgs& a, cudaStream_t stream);

// Same contraction on the tensor cores with operands stored as fp16 hi/lo pairs (see mlp_tcgen05.cu).
struct GemmArgsH {
    const __half* Ahi; const __half* Alo; int lda; int aCols; int aBatchCols;      // A: [M][lda], batch b reads columns b*aBatchCols + k
    const __half* Bhi; const __half* Blo; int ldb; int bRows; int bBatchRows;      // B: [bRows][ldb], batch b reads rows b*bBatchRows + n
    __half* Chi; __half* Clo; float* C32; int ldc; int cBatchCols;
    const float* bias; int biasBatch;
    const __half* actHi; const __half* actLo; int ldact; int actBatchCols;
    float outScale;
    const float* w3; double* energyAcc; float seedScale;   // epilogue 3 (last hidden layer): w3[batch*biasBatch + n]
    int M, N, K, batch;
    // 0 fp32 out * outScale, 1 bias + celu -> hi/lo, 2 times celu'(act) -> hi/lo,
    // 3 last hidden layer: a = celu(z + bias); energy += a*w3; out = seedScale*w3*celu'(a) -> hi/lo (the backward seed)
    int epilogue;
};
void launch_gemm_tcgen05(const GemmArgsH& a, cudaStream_t stream);

}  // namespace nnpops
