// ANI atomic-environment-vector (AEV) forward/backward on sm_100a.  Replaces the reference's
// CudaANISymmetryFunctions (src/ani/CudaANISymmetryFunctions.cu:186-670, kernels K1-K5 of SURVEY.md section 2.3) behind the same
// contract as the abstract class ANISymmetryFunctions (src/ani/ANISymmetryFunctions.h:29-154).
#pragma once
#include <cuda_fp16.h>
#include <vector>
#include "cell_list.cuh"

namespace nnpops {

constexpr int kAniMaxRadial = 64;
constexpr int kAniMaxAngular = 128;
constexpr int kAniMaxSpecies = 32;
constexpr int kAniMaxShf = 16;

// Parameter tables, resident in HBM and staged to shared memory by each kernel.
struct AniTables {
    int nSpecies, nRadial, nAngular, nPairs;
    float rcr, rca;
    float rcr2, rca2;            // cutoff^2 formed in fp32 exactly like the reference (radialCutoff*radialCutoff)
    int torchani;
    float radialScale;           // 0.25 (TorchANI) or 1            CpuANISymmetryFunctions.cpp:99-103
    float cosScale;              // 0.95 (TorchANI) or 1            CpuANISymmetryFunctions.cpp:391-392
    float rEta[kAniMaxRadial], rEtaL2[kAniMaxRadial], rShf[kAniMaxRadial];
    float aEta[kAniMaxAngular], aEtaL2[kAniMaxAngular], aShf[kAniMaxAngular], aZeta[kAniMaxAngular];
    float aCos[kAniMaxAngular], aSin[kAniMaxAngular], aScale[kAniMaxAngular];   // cos/sin(thetas), 2^(1-zeta)
    // factorised form  m = a * nShfZ + z  (single EtaA, single Zeta), used by the fast kernels when fast != 0
    int fast, nShfA, nShfZ;
    float fEta, fEtaL2, fZeta, fScale;
    float fShfA[kAniMaxShf], fCos[kAniMaxShf], fSin[kAniMaxShf];
    float invRcr, invRca, sqrtCosScale;   // 1 / rcr, 1 / rca, sqrt(cosScale): host-computed for the geometry kernels
};

class AniAev {
public:
    // radialFn: nRadial x {eta, rs}; angularFn: nAngular x {eta, rs, zeta, thetas}  (RadialFunction / AngularFunction of
    // ANISymmetryFunctions.h:29-39); atomSpecies: host array [numAtoms].
    AniAev(int numAtoms, int numSpecies, float rcr, float rca, const int* atomSpecies, int nRadial, const float* radialFn,
           int nAngular, const float* angularFn, bool torchani, int maxRadialNeighbors, int maxAngularNeighbors);
    ~AniAev();
    AniAev(const AniAev&) = delete;
    AniAev& operator=(const AniAev&) = delete;

    // Optional map atom -> output row (device int[numAtoms]); nullptr = identity.  Used by the fused ANI model to write the AEV
    // matrix in species-sorted row order.
    void setRowMap(const int* deviceRowMap) { rowMap_ = deviceRowMap; }
    // Optional ownership mask (device unsigned char[numAtoms], indexed by atom): when one box is sharded over several GPUs a rank
    // evaluates only the centres it owns (all atoms stay neighbour candidates) and backward() switches to the centre-owned radial
    // form, so that positionGrad holds this rank's PARTIAL dE/dx over all atoms (sum over ranks = total).  nullptr = all owned.
    void setOwned(const unsigned char* deviceMask) { owned_ = deviceMask; }
    // Verlet skin (0 = off, the default): the neighbour search keeps candidate rows within cutoff + skin and reuses them until some
    // atom has moved more than skin / 2 (or the box has changed); the decision is taken on the device every call.  The rows handed to
    // the AEV kernels are exact at every step.  The steady state of the reference's callers is a time-stepping loop that rebuilds its
    // N x N table every iteration (BenchmarkCudaCFConv.cu:105-112).
    void setSkin(float skin);
    float skin() const { return skin_; }
    void skinStats(unsigned long long* rebuilds, unsigned long long* reuses);   // synchronises

    // positions [n][3], box [3][3] or nullptr (all device, fp32).  radial/angular: device, row strides in floats.
    // ev (optional): forward records ev[0] after the neighbour rows and ev[1] after the radial kernel; backward records ev[0]
    // after the radial kernel -- used by the benchmark to time each kernel on the launching stream.
    // splitHi/splitLo (optional): write the AEV as fp16 hi/lo pairs into one [n][radialStride] matrix pair (radial block first,
    // angular block at column radialWidth()) instead of fp32 -- the operand format of the tensor-core MLP.
    void forward(const float* positions, const float* box, float* radial, int radialStride, float* angular, int angularStride,
                 cudaStream_t stream, cudaEvent_t* ev = nullptr, __half* splitHi = nullptr, __half* splitLo = nullptr);
    // uses the positions/box of the most recent forward (ANISymmetryFunctions.h:83-84)
    void backward(const float* radialGrad, int radialStride, const float* angularGrad, int angularStride, float* positionGrad,
                  cudaStream_t stream, cudaEvent_t* ev = nullptr);
    // synchronises; returns nonzero when a neighbour row overflowed its capacity during any call so far
    int overflowed();
    // does not synchronise: the flag as of the last forward whose device work has completed (every forward copies it to pinned host
    // memory behind the row kernel).  Lets callers that must not block -- CUDA-graph replay, the fused model -- report an overflow
    // on their next call instead of never.
    int overflowPoll() const { return flagHost_ ? *(volatile int*)flagHost_ : 0; }
    int maxRadialNeighbors() const { return capR_; }
    int maxAngularNeighbors() const { return capA_; }

    int numAtoms() const { return n_; }
    int numSpecies() const { return tabHost_.nSpecies; }
    int radialWidth() const { return tabHost_.nSpecies * tabHost_.nRadial; }
    int angularWidth() const { return tabHost_.nPairs * tabHost_.nAngular; }
    long long countTriples(cudaStream_t stream);   // sum_i n_i (n_i - 1) / 2 over the last forward (synchronises)
    long long countRadialPairs(cudaStream_t stream);   // undirected pairs within Rcr over the last forward (synchronises)

private:
    int n_;
    int capR_, capA_;
    AniTables tabHost_;
    AniTables* tab_ = nullptr;
    int* species_ = nullptr;     // device [n]
    CellList cells_;
    int* rowRad_ = nullptr;      // [n][capR] sorted indices, grouped by species
    int* rowAng_ = nullptr;      // [n][capA]
    int* offRad_ = nullptr;      // [n][S+1]
    int* offAng_ = nullptr;      // [n][S+1]
    int* flag_ = nullptr;        // overflow flag
    int* flagHost_ = nullptr;    // pinned host mirror of the flag
    float skin_ = 0.0f;
    int capC_ = 0;
    int* candRow_ = nullptr;     // [n][capC] sorted indices within cutoff + skin at the last rebuild
    int* candCnt_ = nullptr;     // [n]
    float* skinRefPos_ = nullptr;   // [n][3] positions at the last rebuild
    float* skinRefBox_ = nullptr;   // [9]
    int* skinRebuild_ = nullptr;    // device flag of the current call
    unsigned long long* skinStats_ = nullptr;   // {rebuild steps, reuse steps}
    unsigned long long* counters_ = nullptr;   // [2] scratch for countTriples / countRadialPairs
    // second-generation angular kernels (ani_angular_v2.cu)
    float4* geoA_ = nullptr;     // [n][capA] {sqrt(cosScale) * unit vector, r / 2}, species-grouped like rowAng
    float4* geoB_ = nullptr;     // [n][capA] {fc, fc', 1 / r, species << 24 | atom index}
    int* segHist_ = nullptr;     // [2][512] size histogram of the (centre, species pair) blocks + fill cursors
    int4* segs_ = nullptr;       // [n * nPairs] non-empty blocks, largest first
    int* nSeg_ = nullptr;
    bool lastForwardV2_ = false;
    float4* radGeoA_ = nullptr;  // [n][capR] {unit vector, r} of every radial pair, species-grouped like rowRad
    float4* radGeoB_ = nullptr;  // [n][capR] {fc, fc', species << 24 | atom index, -}
    bool lastForwardRadV2_ = false;
    float4* gradAcc_ = nullptr;  // [n] padded force accumulator of the second-generation backward kernels
    bool useV2(const float* angular, int angularStride, const __half* splitHi, const __half* splitLo) const;
    const int* rowMap_ = nullptr;
    const unsigned char* owned_ = nullptr;
    bool haveForward_ = false;
    cudaStream_t aux_ = nullptr;            // radial kernels run here, concurrently with the angular kernels on the caller's stream
    cudaEvent_t evFork_ = nullptr, evJoin_ = nullptr;
};

}  // namespace nnpops
