#include "ani_model.cuh"

namespace nnpops {

namespace {
__global__ void gather_rows_kernel(const float* __restrict__ src, const __half* __restrict__ hi, const __half* __restrict__ lo, int stride,
                                   const int* __restrict__ rowMap, int n, int width, float* __restrict__ out) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * width) return;
    const int i = (int)(idx / width), c = (int)(idx % width);
    const size_t o = (size_t)rowMap[i] * stride + c;
    out[idx] = hi ? fmaf(__half2float(lo[o]), 1.0f / 2048.0f, __half2float(hi[o])) : src[o];
}
}  // namespace

void AniModel::readFeatures(int which, float* out, cudaStream_t stream) {
    if (n_ == 0) return;
    const size_t tot = (size_t)n_ * nFeat_;
    const bool split = which == 0 && mlp_->tensorCore();
    gather_rows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(which == 0 ? feat_ : featGrad_, split ? mlp_->featHi() : nullptr,
                                                                          split ? mlp_->featLo() : nullptr, stride_, rowMap_, n_, nFeat_, out);
    NNP_CUDA_CHECK(cudaGetLastError());
}

AniModel::AniModel(int numAtoms, int numSpecies, float rcr, float rca, const int* atomSpecies, int nRadial, const float* radialFn,
                   int nAngular, const float* angularFn, int ensemble, int numLayers, const int* dims, const float* params,
                   int maxRadialNeighbors, int maxAngularNeighbors)
    : n_(numAtoms) {
    aev_.reset(new AniAev(numAtoms, numSpecies, rcr, rca, atomSpecies, nRadial, radialFn, nAngular, angularFn, true, maxRadialNeighbors,
                          maxAngularNeighbors));
    const int nFeat = aev_->radialWidth() + aev_->angularWidth();
    NNP_REQUIRE(dims[0] == nFeat, "network input size must equal the AEV length");
    nFeat_ = nFeat;
    stride_ = (nFeat + kMlpPad - 1) / kMlpPad * kMlpPad;
    // species-sorted row order (stable in the atom index): the species of a system never change, so this is done once
    std::vector<int> rowStart(numSpecies + 1, 0);
    for (int i = 0; i < numAtoms; i++) rowStart[atomSpecies[i] + 1]++;
    for (int s = 0; s < numSpecies; s++) rowStart[s + 1] += rowStart[s];
    std::vector<int> cursor(rowStart.begin(), rowStart.end() - 1);
    rowOfAtom_.resize(numAtoms);
    for (int i = 0; i < numAtoms; i++) rowOfAtom_[i] = cursor[atomSpecies[i]]++;
    const size_t na = (size_t)(numAtoms > 0 ? numAtoms : 1);
    NNP_CUDA_CHECK(cudaMalloc(&rowMap_, sizeof(int) * na));
    if (numAtoms > 0) NNP_CUDA_CHECK(cudaMemcpy(rowMap_, rowOfAtom_.data(), sizeof(int) * numAtoms, cudaMemcpyHostToDevice));
    NNP_CUDA_CHECK(cudaMalloc(&feat_, sizeof(float) * na * stride_));
    NNP_CUDA_CHECK(cudaMalloc(&featGrad_, sizeof(float) * na * stride_));
    NNP_CUDA_CHECK(cudaMemset(feat_, 0, sizeof(float) * na * stride_));   // padding columns stay zero forever
    NNP_CUDA_CHECK(cudaMemset(featGrad_, 0, sizeof(float) * na * stride_));
    aev_->setRowMap(rowMap_);
    mlp_.reset(new SpeciesMlp(numSpecies, ensemble, numLayers, dims, params, rowStart.data(), stride_));
}

AniModel::~AniModel() {
    cudaFree(rowMap_); cudaFree(feat_); cudaFree(featGrad_);
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
    mlp_->forward(tc ? nullptr : feat_, energy, stream);
    if (ev) cudaEventRecord(ev[4], stream);
    mlp_->backward(featGrad_, stream);
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
