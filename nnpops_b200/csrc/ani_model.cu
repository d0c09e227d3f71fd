#include "ani_model.cuh"
#include <algorithm>
#include <cstdlib>

namespace nnpops {

namespace {
// out[i][f] (atom order, full AEV layout) <- row rowMap[i], column colOfFull[f] of the compact species-sorted matrix (0 for a
// column of an absent species)
__global__ void gather_rows_kernel(const float* __restrict__ src, const __half* __restrict__ hi, const __half* __restrict__ lo, int stride,
                                   const int* __restrict__ rowMap, const int* __restrict__ colOfFull, int n, int width,
                                   float* __restrict__ out) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * width) return;
    const int i = (int)(idx / width), c = colOfFull[(int)(idx % width)];
    if (c < 0) { out[idx] = 0.0f; return; }
    const size_t o = (size_t)rowMap[i] * stride + c;
    out[idx] = hi ? fmaf(__half2float(lo[o]), 1.0f / 2048.0f, __half2float(hi[o])) : src[o];
}
}  // namespace

void AniModel::readFeatures(int which, float* out, cudaStream_t stream) {
    if (n_ == 0) return;
    const size_t tot = (size_t)n_ * nFeatFull_;
    const bool split = which == 0 && mlp_->tensorCore();
    gather_rows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(which == 0 ? feat_ : featGrad_, split ? mlp_->featHi() : nullptr,
                                                                          split ? mlp_->featLo() : nullptr, stride_, rowMap_, colOfFull_, n_,
                                                                          nFeatFull_, out);
    NNP_CUDA_CHECK(cudaGetLastError());
}

// Active-feature compaction.  An AEV column belongs to a neighbour species (radial block) or a pair of neighbour species
// (angular block); a block whose species do not occur in the system is identically zero for every atom and every geometry
// (the species of a Holder never change, SymmetryFunctions.cpp:52-92).  The fused model therefore runs the AEV kernels and the
// first MLP layer in the species space of the system: S' present species, AEV length S' nR + S'(S'+1)/2 nA, first-layer weight
// columns gathered once on the host.  The products that disappear are all 0 * w, so energies and forces are unchanged; the
// gradient with respect to the dropped (position-independent) columns is not formed.  Water (H, O) keeps 128 of ANI-2x's 1008
// columns, a 5-element protein 560.  NNPOPS_DENSE_AEV=1 (or compact = false) disables it.
AniModel::AniModel(int numAtoms, int numSpecies, float rcr, float rca, const int* atomSpecies, int nRadial, const float* radialFn,
                   int nAngular, const float* angularFn, int ensemble, int numLayers, const int* dims, const float* params,
                   int maxRadialNeighbors, int maxAngularNeighbors, bool compact, int shardRank, int shardCount, const unsigned char* ownedMask)
    : n_(numAtoms) {
    NNP_REQUIRE(numSpecies >= 1 && numSpecies <= kAniMaxSpecies, "unsupported number of species");
    NNP_REQUIRE(shardCount >= 1 && shardRank >= 0 && shardRank < shardCount, "shard rank must be in [0, shard count)");
    // One box sharded over several GPUs: atom i is a CENTRE of rank i mod shardCount (interleaved ownership balances species and
    // density without any knowledge of the geometry); every atom stays a neighbour candidate on every rank.
    // Spatial decomposition with ghost halos passes the ownership explicitly: the atoms of this rank's brick are centres, the ghosts
    // around it are neighbour candidates only (ownedMask, host [numAtoms]).
    std::vector<unsigned char> owned(numAtoms > 0 ? numAtoms : 1, 1);
    if (ownedMask != nullptr)
        for (int i = 0; i < numAtoms; i++) owned[i] = ownedMask[i] ? 1 : 0;
    else if (shardCount > 1)
        for (int i = 0; i < numAtoms; i++) owned[i] = (i % shardCount == shardRank) ? 1 : 0;
    const int S = numSpecies, L = numLayers;
    const int nFeatFull = S * nRadial + S * (S + 1) / 2 * nAngular;
    NNP_REQUIRE(dims[0] == nFeatFull, "network input size must equal the AEV length");
    nFeatFull_ = nFeatFull;
    if (std::getenv("NNPOPS_DENSE_AEV") != nullptr) compact = false;
    std::vector<int> compactOf(S, -1), present;
    {
        std::vector<char> seen(S, 0);
        for (int i = 0; i < numAtoms; i++) {
            NNP_REQUIRE(atomSpecies[i] >= 0 && atomSpecies[i] < S, "atom species out of range");
            seen[atomSpecies[i]] = 1;
        }
        for (int s = 0; s < S; s++)
            if (seen[s] || !compact || numAtoms == 0) { compactOf[s] = (int)present.size(); present.push_back(s); }
    }
    const int Sc = (int)present.size();
    auto pairIndex = [](int nS, int a, int b) { const int lo = a < b ? a : b, hi = a < b ? b : a; return lo * nS - (lo * (lo - 1)) / 2 + (hi - lo); };
    // compact column -> full column
    std::vector<int> fullOf;
    for (int a = 0; a < Sc; a++)
        for (int k = 0; k < nRadial; k++) fullOf.push_back(present[a] * nRadial + k);
    for (int a = 0; a < Sc; a++)
        for (int b = a; b < Sc; b++)
            for (int m = 0; m < nAngular; m++) fullOf.push_back(S * nRadial + pairIndex(S, present[a], present[b]) * nAngular + m);
    const int nFeat = (int)fullOf.size();
    nFeat_ = nFeat;
    std::vector<int> colOfFull(nFeatFull, -1);
    for (int c = 0; c < nFeat; c++) colOfFull[fullOf[c]] = c;
    // species, layer sizes and parameters in the compact species space
    std::vector<int> speciesC(numAtoms), dimsC((size_t)Sc * (L + 1));
    for (int i = 0; i < numAtoms; i++) speciesC[i] = compactOf[atomSpecies[i]];
    std::vector<float> paramsC;
    std::vector<long long> atomsOf(S, 0);   // centres of this rank per species
    for (int i = 0; i < numAtoms; i++) atomsOf[atomSpecies[i]] += owned[i];
    {
        const float* p = params;
        for (int s = 0; s < S; s++) {
            const int* d = dims + (size_t)s * (L + 1);
            NNP_REQUIRE(d[0] == nFeatFull, "all species must take the same number of input features");
            const bool keep = compactOf[s] >= 0;
            if (keep) {
                int* dc = dimsC.data() + (size_t)compactOf[s] * (L + 1);
                for (int l = 0; l <= L; l++) dc[l] = d[l];
                dc[0] = nFeat;
            }
            for (int e = 0; e < ensemble; e++)
                for (int l = 0; l < L; l++) {
                    const int in = d[l], out = d[l + 1];
                    if (keep) {
                        if (l == 0) {
                            for (int o = 0; o < out; o++)
                                for (int c = 0; c < nFeat; c++) paramsC.push_back(p[(size_t)o * in + fullOf[c]]);
                        } else {
                            paramsC.insert(paramsC.end(), p, p + (size_t)out * in);
                        }
                        paramsC.insert(paramsC.end(), p + (size_t)out * in, p + (size_t)out * in + out);
                        denseFlopsFwd_ += 2.0 * in * out * (double)atomsOf[s];
                    }
                    p += (size_t)out * in + out;
                }
        }
    }
    aev_.reset(new AniAev(numAtoms, Sc, rcr, rca, speciesC.data(), nRadial, radialFn, nAngular, angularFn, true, maxRadialNeighbors,
                          maxAngularNeighbors));
    NNP_REQUIRE(aev_->radialWidth() + aev_->angularWidth() == nFeat, "internal: compact AEV length mismatch");
    stride_ = (nFeat + kMlpPad - 1) / kMlpPad * kMlpPad;
    // species-sorted row order (stable in the atom index): the species of a system never change, so this is done once
    std::vector<int> rowStart(Sc + 1, 0);
    for (int i = 0; i < numAtoms; i++) rowStart[speciesC[i] + 1] += owned[i];
    for (int s = 0; s < Sc; s++) rowStart[s + 1] += rowStart[s];
    const int nOwned = rowStart[Sc];
    std::vector<int> cursor(rowStart.begin(), rowStart.end() - 1);
    rowOfAtom_.resize(numAtoms);
    // centres of other ranks have empty neighbour rows; whatever the AEV kernels store for them lands in the spare row nOwned
    for (int i = 0; i < numAtoms; i++) rowOfAtom_[i] = owned[i] ? cursor[speciesC[i]]++ : nOwned;
    const size_t na = (size_t)nOwned + 1;
    if (shardCount > 1 || ownedMask != nullptr) {
        NNP_CUDA_CHECK(cudaMalloc(&owned_, owned.size()));
        NNP_CUDA_CHECK(cudaMemcpy(owned_, owned.data(), owned.size(), cudaMemcpyHostToDevice));
    }
    NNP_CUDA_CHECK(cudaMalloc(&rowMap_, sizeof(int) * (size_t)(numAtoms > 0 ? numAtoms : 1)));
    if (numAtoms > 0) NNP_CUDA_CHECK(cudaMemcpy(rowMap_, rowOfAtom_.data(), sizeof(int) * numAtoms, cudaMemcpyHostToDevice));
    NNP_CUDA_CHECK(cudaMalloc(&colOfFull_, sizeof(int) * nFeatFull));
    NNP_CUDA_CHECK(cudaMemcpy(colOfFull_, colOfFull.data(), sizeof(int) * nFeatFull, cudaMemcpyHostToDevice));
    NNP_CUDA_CHECK(cudaMalloc(&feat_, sizeof(float) * na * stride_));
    NNP_CUDA_CHECK(cudaMalloc(&featGrad_, sizeof(float) * na * stride_));
    NNP_CUDA_CHECK(cudaMemset(feat_, 0, sizeof(float) * na * stride_));   // padding columns stay zero forever
    NNP_CUDA_CHECK(cudaMemset(featGrad_, 0, sizeof(float) * na * stride_));
    aev_->setRowMap(rowMap_);
    aev_->setOwned(owned_);
    mlp_.reset(new SpeciesMlp(Sc, ensemble, numLayers, dimsC.data(), paramsC.data(), rowStart.data(), stride_));
}

AniModel::~AniModel() {
    cudaFree(rowMap_); cudaFree(owned_); cudaFree(colOfFull_); cudaFree(feat_); cudaFree(featGrad_);
    for (cudaEvent_t e : events_) cudaEventDestroy(e);
}

void AniModel::energyAndGradient(const float* positions, const float* box, float* energy, float* positionGrad, cudaStream_t stream) {
    const int rw = aev_->radialWidth();
    cudaEvent_t* ev = nullptr;
    if (timingUsed_ < timingCap_) ev = events_.data() + (size_t)(timingUsed_++) * (kStages + 1);
    if (ev) cudaEventRecord(ev[0], stream);
    const bool tc = mlp_->tensorCore();   // the AEV kernels then write the fp16 hi/lo operand pair of the MLP directly
    aev_->forward(positions, box, feat_, stride_, feat_ + rw, stride_, stream, ev ? ev + 1 : nullptr,   // ev[1], ev[2]
                  tc ? mlp_->featHi() : nullptr, tc ? mlp_->featLo() : nullptr);
    if (ev) cudaEventRecord(ev[3], stream);
    if (mlp_->fused()) {   // forward + backward of the network in one kernel: the whole MLP time is booked on stage 3
        mlp_->forwardBackward(energy, featGrad_, stream);
        if (ev) cudaEventRecord(ev[4], stream);
    } else {
        mlp_->forward(tc ? nullptr : feat_, energy, stream);
        if (ev) cudaEventRecord(ev[4], stream);
        mlp_->backward(featGrad_, stream);
    }
    if (ev) cudaEventRecord(ev[5], stream);
    aev_->backward(featGrad_, stride_, featGrad_ + rw, stride_, positionGrad, stream, ev ? ev + 6 : nullptr);   // ev[6]
    if (ev) cudaEventRecord(ev[7], stream);
}

void AniModel::timingBegin(int maxSteps) {
    for (cudaEvent_t e : events_) cudaEventDestroy(e);
    events_.assign((size_t)maxSteps * (kStages + 1), nullptr);
    for (auto& e : events_) NNP_CUDA_CHECK(cudaEventCreate(&e));
    timingCap_ = maxSteps;
    timingUsed_ = 0;
}

int AniModel::timingEnd(float* stageMs) {
    NNP_CUDA_CHECK(cudaDeviceSynchronize());
    for (int k = 0; k < kStages; k++) stageMs[k] = 0.0f;
    for (int i = 0; i < timingUsed_; i++) {
        cudaEvent_t* ev = events_.data() + (size_t)i * (kStages + 1);
        for (int k = 0; k < kStages; k++) {
            float ms = 0.0f;
            NNP_CUDA_CHECK(cudaEventElapsedTime(&ms, ev[k], ev[k + 1]));
            stageMs[k] += ms;
        }
    }
    const int used = timingUsed_;
    for (cudaEvent_t e : events_) cudaEventDestroy(e);
    events_.clear();
    timingCap_ = timingUsed_ = 0;
    return used;
}

}  // namespace nnpops
