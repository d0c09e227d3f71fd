// Fused ANI energy + force evaluation: AEV forward -> species-grouped MLP forward/backward -> AEV backward, one object per
// system, everything on one stream, no host synchronisation.  This is the B200-native counterpart of the reference's
// OptimizedTorchANI module chain (src/pytorch/OptimizedTorchANI.py:33-54) for the BASELINE.json headline metric.
#pragma once
#include <memory>
#include "ani_aev.cuh"
#include "species_mlp.cuh"

namespace nnpops {

class AniModel {
public:
    AniModel(int numAtoms, int numSpecies, float rcr, float rca, const int* atomSpecies, int nRadial, const float* radialFn,
             int nAngular, const float* angularFn, int ensemble, int numLayers, const int* dims, const float* params,
             int maxRadialNeighbors, int maxAngularNeighbors, bool compact = true, int shardRank = 0, int shardCount = 1,
             const unsigned char* ownedMask = nullptr);
    ~AniModel();
    // energy: device float[1]; positionGrad: device [n][3] = dE/dx (forces = -positionGrad)
    void energyAndGradient(const float* positions, const float* box, float* energy, float* positionGrad, cudaStream_t stream);
    // Per-stage CUDA-event timing on the launching stream.  Stages: 0 cell list + neighbour rows, 1 radial fwd, 2 angular fwd,
    // 3 MLP fwd, 4 MLP bwd, 5 radial bwd, 6 angular bwd.  timingBegin allocates event sets for up to maxSteps evaluations;
    // timingEnd synchronises and returns the summed milliseconds per stage and the number of evaluations recorded.
    static constexpr int kStages = 7;
    void timingBegin(int maxSteps);
    int timingEnd(float* stageMs);
    bool timingActive() const { return timingUsed_ < timingCap_; }
    void readFeatures(int which, float* out, cudaStream_t stream);   // atom order, [n][aevLength]
    int aevLength() const { return nFeatFull_; }        // the model's full AEV length (readFeatures layout)
    int activeFeatures() const { return nFeat_; }      // columns that can be non-zero for this system's species
    double denseFlopsForward() const { return denseFlopsFwd_; }   // MLP forward flops counted on the full AEV length
    AniAev& aev() { return *aev_; }
    SpeciesMlp& mlp() { return *mlp_; }
    float* features() { return feat_; }          // [n][featureStride], species-sorted rows
    float* featureGrad() { return featGrad_; }
    int featureStride() const { return stride_; }
    const std::vector<int>& rowOfAtom() const { return rowOfAtom_; }

private:
    std::unique_ptr<AniAev> aev_;
    std::unique_ptr<SpeciesMlp> mlp_;
    int n_, stride_, nFeat_, nFeatFull_;
    double denseFlopsFwd_ = 0;
    unsigned char* owned_ = nullptr;   // device [numAtoms] when the box is sharded over several ranks (else nullptr)
    int* colOfFull_ = nullptr;   // device [nFeatFull]: compact column or -1
    float* feat_ = nullptr;
    float* featGrad_ = nullptr;
    int* rowMap_ = nullptr;
    std::vector<int> rowOfAtom_;
    std::vector<cudaEvent_t> events_;   // (kStages + 1) per recorded step
    int timingCap_ = 0, timingUsed_ = 0;
};

}  // namespace nnpops
