// Species-grouped MLP: host orchestration + the fp32 SIMT GEMM (validation path).  The tensor-core path lives in
// mlp_tcgen05.cu and plugs in through the same GemmArgs.
#include "species_mlp.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace nnpops {


namespace {

constexpr int BM = 128, BN = 64, BK = 16;

__device__ __forceinline__ float celu(float x) { return x > 0.0f ? x : kCeluAlpha * (__expf(x * (1.0f / kCeluAlpha)) - 1.0f); }
// derivative recovered from the activation value a = celu(z): 1 for a > 0, exp(z/alpha) = a/alpha + 1 otherwise
__device__ __forceinline__ float celu_grad_from_act(float a) { return a > 0.0f ? 1.0f : a * (1.0f / kCeluAlpha) + 1.0f; }

__global__ void __launch_bounds__(256) gemm_tn_simt_kernel(GemmArgs g) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int z = blockIdx.z;
    const float* A = g.A + (size_t)z * g.aBatchCols;
    const float* B = g.B + (size_t)z * g.bBatch;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.0f;
    for (int k0 = 0; k0 < g.K; k0 += BK) {
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const int idx = tid + i * 256, row = idx >> 2, kq = idx & 3;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m0 + row < g.M) v = *reinterpret_cast<const float4*>(A + (size_t)(m0 + row) * g.lda + k0 + kq * 4);
            As[kq * 4 + 0][row] = v.x; As[kq * 4 + 1][row] = v.y; As[kq * 4 + 2][row] = v.z; As[kq * 4 + 3][row] = v.w;
        }
        {
            const int row = tid >> 2, kq = tid & 3;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n0 + row < g.N) v = *reinterpret_cast<const float4*>(B + (size_t)(n0 + row) * g.ldb + k0 + kq * 4);
            Bs[kq * 4 + 0][row] = v.x; Bs[kq * 4 + 1][row] = v.y; Bs[kq * 4 + 2][row] = v.z; Bs[kq * 4 + 3][row] = v.w;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; k++) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* C = g.C + (size_t)z * g.cBatchCols;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int m = m0 + ty * 8 + i;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int n = n0 + tx * 4 + j;
            if (n >= g.N) continue;
            float v = acc[i][j];
            if (g.epilogue == 1) v = celu(v + g.bias[(size_t)z * g.biasBatch + n]);
            else if (g.epilogue == 2) v *= celu_grad_from_act(g.act[(size_t)m * g.ldact + (size_t)z * g.actBatchCols + n]);
            C[(size_t)m * g.ldc + n] = v;
        }
    }
}

// last layer (out = 1): E[r, e] = sum_c A[r, e*hP + c] * w[e][c] + b[e]; accumulate sum over rows and members in double
__global__ void final_energy_kernel(const float* __restrict__ act, int ld, int rows, int M, int hP, const float* __restrict__ w,
                                    const float* __restrict__ b, double* __restrict__ acc) {
    // one warp per (row, member)
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    float s = 0.0f;
    if (gw < rows * M) {
        const int r = gw / M, e = gw % M;
        const float* a = act + (size_t)r * ld + (size_t)e * hP;
        for (int c = lane; c < hP; c += 32) s = fmaf(a[c], w[e * hP + c], s);
        s = warp_sum(s);
        if (lane == 0) s += b[e];
    }
    __shared__ double part[32];
    double d = (gw < rows * M && lane == 0) ? (double)s : 0.0;
    if (lane == 0) part[threadIdx.x >> 5] = d;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) t += part[i];
        atomicAdd(acc, t);
    }
}

__global__ void finish_energy_kernel(const double* __restrict__ acc, double bias, int M, float* __restrict__ energy) {
    *energy = (float)((*acc + bias) / M);
}

// dZ[r, e*hP + c] = (w[e][c] / M) * celu'(A[r, e*hP + c])
__global__ void backward_seed_kernel(const float* __restrict__ act, int ld, int rows, int M, int hP, const float* __restrict__ w,
                                     float* __restrict__ dz, int lddz) {
    const int width = M * hP;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)rows * width) return;
    const int r = (int)(idx / width), c = (int)(idx % width);
    dz[(size_t)r * lddz + c] = (w[c] * (1.0f / M)) * celu_grad_from_act(act[(size_t)r * ld + c]);
}

constexpr float kLoScale = 2048.0f, kLoInv = 1.0f / 2048.0f;
constexpr float kGradScale = 1024.0f;   // power-of-two lift of the backward chain into fp16's comfortable range (exact)

__device__ __forceinline__ void split_h(float v, __half& hi, __half& lo) {
    hi = __float2half_rn(v);
    lo = __float2half_rn((v - __half2float(hi)) * kLoScale);
}
__device__ __forceinline__ float join_h(__half hi, __half lo) { return fmaf(__half2float(lo), kLoInv, __half2float(hi)); }

// fp32 [rows][ld] -> hi/lo fp16 pairs, 8 elements per thread
__global__ void split_rows_kernel(const float* __restrict__ src, size_t n8, __half* __restrict__ hi, __half* __restrict__ lo) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n8) return;
    const float4 a = reinterpret_cast<const float4*>(src)[2 * i], b = reinterpret_cast<const float4*>(src)[2 * i + 1];
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint4 ph, pl;
    __half* hh = reinterpret_cast<__half*>(&ph);
    __half* hl = reinterpret_cast<__half*>(&pl);
#pragma unroll
    for (int k = 0; k < 8; k++) split_h(v[k], hh[k], hl[k]);
    reinterpret_cast<uint4*>(hi)[i] = ph;
    reinterpret_cast<uint4*>(lo)[i] = pl;
}

__global__ void final_energy_h_kernel(const __half* __restrict__ actHi, const __half* __restrict__ actLo, int ld, int rows, int M, int hP,
                                      const float* __restrict__ w, const float* __restrict__ b, double* __restrict__ acc) {
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    float s = 0.0f;
    if (gw < rows * M) {
        const int r = gw / M, e = gw % M;
        const size_t o = (size_t)r * ld + (size_t)e * hP;
        for (int c = lane; c < hP; c += 32) s = fmaf(join_h(actHi[o + c], actLo[o + c]), w[e * hP + c], s);
        s = warp_sum(s);
        if (lane == 0) s += b[e];
    }
    __shared__ double part[32];
    if (lane == 0) part[threadIdx.x >> 5] = (gw < rows * M) ? (double)s : 0.0;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) t += part[i];
        atomicAdd(acc, t);
    }
}

__global__ void backward_seed_h_kernel(const __half* __restrict__ actHi, const __half* __restrict__ actLo, int ld, int rows, int M, int hP,
                                       const float* __restrict__ w, __half* __restrict__ dzHi, __half* __restrict__ dzLo, int lddz) {
    const int width = M * hP;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)rows * width) return;
    const int r = (int)(idx / width), c = (int)(idx % width);
    const float a = join_h(actHi[(size_t)r * ld + c], actLo[(size_t)r * ld + c]);
    const float v = (w[c] * (kGradScale / M)) * celu_grad_from_act(a);
    split_h(v, dzHi[(size_t)r * lddz + c], dzLo[(size_t)r * lddz + c]);
}

}  // namespace

void launch_gemm_simt(const GemmArgs& a, cudaStream_t stream) {
    if (a.M <= 0 || a.N <= 0) return;
    NNP_REQUIRE(a.K % BK == 0 && a.lda % 4 == 0 && a.ldb % 4 == 0 && a.aBatchCols % 4 == 0 && a.bBatch % 4 == 0,
                "gemm_simt: K must be a multiple of 16 and leading dimensions multiples of 4");
    dim3 grid((a.N + BN - 1) / BN, (a.M + BM - 1) / BM, a.batch);
    gemm_tn_simt_kernel<<<grid, 256, 0, stream>>>(a);
    count_launch();
}

static int pad_to(int x, int p) { return (x + p - 1) / p * p; }

SpeciesMlp::SpeciesMlp(int numSpecies, int ensemble, int numLayers, const int* dims, const float* params, const int* rowStart,
                       int featureStride)
    : S_(numSpecies), M_(ensemble), L_(numLayers), featStride_(featureStride) {
    NNP_REQUIRE(numLayers >= 2, "the per-atom network needs at least two layers");
    NNP_REQUIRE(ensemble >= 1, "ensemble size must be >= 1");
    rowStart_.assign(rowStart, rowStart + numSpecies + 1);
    rows_ = rowStart_[numSpecies];
    const int nFeat = dims[0];
    featP_ = pad_to(nFeat, kMlpPad);
    NNP_REQUIRE(featureStride >= featP_ && featureStride % 4 == 0, "featureStride must be >= numFeatures padded to 32");
    layers_.resize(S_);
    width_.assign(L_ - 1, 0);
    const float* p = params;
    for (int s = 0; s < S_; s++) {
        const int* d = dims + (size_t)s * (L_ + 1);
        NNP_REQUIRE(d[0] == nFeat, "all species must take the same number of input features");
        NNP_REQUIRE(d[L_] == 1, "the last layer must have one output");
        layers_[s].resize(L_);
        for (int l = 0; l < L_; l++) {
            Layer& ly = layers_[s][l];
            ly.in = d[l]; ly.out = d[l + 1];
            ly.inP = pad_to(ly.in, kMlpPad);
            ly.outP = (l == L_ - 1) ? 1 : pad_to(ly.out, kMlpPad);
            if (l < L_ - 1) width_[l] = std::max(width_[l], M_ * ly.outP);
        }
    }
    // params are laid out species -> member -> layer, so walk them in that order while filling the per-layer staging
    std::vector<std::vector<std::vector<float>>> hW(S_), hWt(S_), hb(S_);
    for (int s = 0; s < S_; s++) {
        hW[s].resize(L_); hWt[s].resize(L_); hb[s].resize(L_);
        for (int l = 0; l < L_; l++) {
            const Layer& ly = layers_[s][l];
            hW[s][l].assign((size_t)M_ * ly.outP * ly.inP, 0.0f);
            hWt[s][l].assign((size_t)M_ * ly.outP * ly.inP, 0.0f);
            hb[s][l].assign((size_t)M_ * ly.outP, 0.0f);
        }
        for (int e = 0; e < M_; e++)
            for (int l = 0; l < L_; l++) {
                const Layer& ly = layers_[s][l];
                for (int o = 0; o < ly.out; o++)
                    for (int i = 0; i < ly.in; i++) {
                        const float v = p[(size_t)o * ly.in + i];
                        hW[s][l][((size_t)e * ly.outP + o) * ly.inP + i] = v;
                        if (l == 0) hWt[s][l][(size_t)i * (M_ * ly.outP) + (size_t)e * ly.outP + o] = v;   // [inP][M*outP]
                        else hWt[s][l][((size_t)e * ly.inP + i) * ly.outP + o] = v;                          // [M][inP][outP]
                    }
                p += (size_t)ly.out * ly.in;
                for (int o = 0; o < ly.out; o++) hb[s][l][(size_t)e * ly.outP + o] = p[o];
                if (l == L_ - 1) energyBias_ += (double)p[0] * (double)(rowStart_[s + 1] - rowStart_[s]);
                p += ly.out;
                if (l < L_ - 1) flopsFwd_ += 2.0 * ly.in * ly.out * (double)(rowStart_[s + 1] - rowStart_[s]);
                else flopsFwd_ += 2.0 * ly.in * (double)(rowStart_[s + 1] - rowStart_[s]);
            }
        for (int l = 0; l < L_; l++) {
            Layer& ly = layers_[s][l];
            const size_t nw = hW[s][l].size(), nb = hb[s][l].size();
            NNP_CUDA_CHECK(cudaMalloc(&ly.W, sizeof(float) * nw));
            NNP_CUDA_CHECK(cudaMalloc(&ly.Wt, sizeof(float) * nw));
            NNP_CUDA_CHECK(cudaMalloc(&ly.b, sizeof(float) * nb));
            NNP_CUDA_CHECK(cudaMemcpy(ly.W, hW[s][l].data(), sizeof(float) * nw, cudaMemcpyHostToDevice));
            NNP_CUDA_CHECK(cudaMemcpy(ly.Wt, hWt[s][l].data(), sizeof(float) * nw, cudaMemcpyHostToDevice));
            NNP_CUDA_CHECK(cudaMemcpy(ly.b, hb[s][l].data(), sizeof(float) * nb, cudaMemcpyHostToDevice));
        }
    }
    hW_ = std::move(hW);
    act_.assign(L_ - 1, nullptr);
    dz_.assign(L_ - 1, nullptr);
    const size_t nr = (size_t)rows_ + 1;   // one spare row (target of the AEV stores of centres this rank does not own)
    for (int l = 0; l < L_ - 1; l++) {
        NNP_CUDA_CHECK(cudaMalloc(&act_[l], sizeof(float) * nr * width_[l]));
        NNP_CUDA_CHECK(cudaMalloc(&dz_[l], sizeof(float) * nr * width_[l]));
    }
    NNP_CUDA_CHECK(cudaMalloc(&energyAcc_, sizeof(double)));
}

SpeciesMlp::~SpeciesMlp() {
    for (auto& sp : layers_)
        for (auto& ly : sp) {
            cudaFree(ly.W); cudaFree(ly.Wt); cudaFree(ly.b);
            cudaFree(ly.Whi); cudaFree(ly.Wlo); cudaFree(ly.Wthi); cudaFree(ly.Wtlo);
        }
    for (auto* v : {&actHi_, &actLo_, &dzHi_, &dzLo_})
        for (__half* p : *v) cudaFree(p);
    cudaFree(featHi_); cudaFree(featLo_);
    for (float* p : act_) cudaFree(p);
    for (float* p : dz_) cudaFree(p);
    cudaFree(energyAcc_);
}

static void to_split_device(const float* src, size_t count, __half** hi, __half** lo) {
    NNP_REQUIRE(count % 8 == 0, "split: element count must be a multiple of 8");
    NNP_CUDA_CHECK(cudaMalloc(hi, sizeof(__half) * count));
    NNP_CUDA_CHECK(cudaMalloc(lo, sizeof(__half) * count));
    split_rows_kernel<<<(unsigned)((count / 8 + 255) / 256), 256>>>(src, count / 8, *hi, *lo);
    NNP_CUDA_CHECK(cudaGetLastError());
}

void SpeciesMlp::setImpl(MlpImpl impl) {
    impl_ = impl;
    if (impl != MlpImpl::Tcgen05 || featHi_ != nullptr) return;
    // one-time preparation of the tensor-core representation: weights and work buffers as fp16 hi/lo pairs
    for (int s = 0; s < S_; s++)
        for (int l = 0; l < L_ - 1; l++) {
            Layer& ly = layers_[s][l];
            const size_t cnt = (size_t)M_ * ly.outP * ly.inP;
            to_split_device(ly.W, cnt, &ly.Whi, &ly.Wlo);
            to_split_device(ly.Wt, cnt, &ly.Wthi, &ly.Wtlo);
        }
    const size_t nr = (size_t)rows_ + 1;   // one spare row (target of the AEV stores of centres this rank does not own)
    actHi_.assign(L_ - 1, nullptr); actLo_.assign(L_ - 1, nullptr); dzHi_.assign(L_ - 1, nullptr); dzLo_.assign(L_ - 1, nullptr);
    for (int l = 0; l < L_ - 1; l++) {
        for (__half** p : {&actHi_[l], &actLo_[l], &dzHi_[l], &dzLo_[l]}) {
            NNP_CUDA_CHECK(cudaMalloc(p, sizeof(__half) * nr * width_[l]));
            NNP_CUDA_CHECK(cudaMemset(*p, 0, sizeof(__half) * nr * width_[l]));
        }
    }
    NNP_CUDA_CHECK(cudaMalloc(&featHi_, sizeof(__half) * nr * featStride_));
    NNP_CUDA_CHECK(cudaMalloc(&featLo_, sizeof(__half) * nr * featStride_));
    NNP_CUDA_CHECK(cudaMemset(featHi_, 0, sizeof(__half) * nr * featStride_));   // padding columns stay zero
    NNP_CUDA_CHECK(cudaMemset(featLo_, 0, sizeof(__half) * nr * featStride_));
    NNP_CUDA_CHECK(cudaDeviceSynchronize());
    // the fp32 activation buffers of the validation path are not needed any more
    for (float*& p : act_) { cudaFree(p); p = nullptr; }
    for (float*& p : dz_) { cudaFree(p); p = nullptr; }
    // three hidden layers on a compact input: the whole chain runs in one kernel with the activations on chip (mlp_chain.cu)
    if (L_ == 4 && std::getenv("NNPOPS_NO_CHAIN") == nullptr && !hW_.empty()) {
        std::vector<MlpChain::SpeciesDesc> sd(S_);
        for (int s = 0; s < S_; s++) {
            MlpChain::SpeciesDesc& d = sd[s];
            d.d[0] = layers_[s][0].inP;
            for (int l = 0; l < 3; l++) { d.d[l + 1] = layers_[s][l].outP; d.W[l] = hW_[s][l].data(); d.bias[l] = layers_[s][l].b; }
            d.rowStart = rowStart_[s]; d.rows = rowStart_[s + 1] - rowStart_[s];
            d.w3 = layers_[s][3].W;
        }
        bool padOk = true;   // the chain reads featureStride columns of X: they must all be real (padded) input columns
        for (int s = 0; s < S_; s++) padOk = padOk && layers_[s][0].inP == featStride_;
        // NNPOPS_CHAIN_V2=1: the single-accumulator variant (mlp_chain2.cu; correct, measured slower: DESIGN.md section 4)
        if (padOk && std::getenv("NNPOPS_CHAIN_V2") != nullptr && MlpChain2::eligible(S_, sd.data(), featStride_))
            chain2_.reset(new MlpChain2(M_, S_, sd.data(), featHi_, featLo_, featStride_));
        else if (padOk && MlpChain::eligible(S_, sd.data(), featStride_))
            chain_.reset(new MlpChain(M_, S_, sd.data(), featHi_, featLo_, featStride_));
    }
    hW_.clear();
    hW_.shrink_to_fit();
}

void SpeciesMlp::forwardBackward(float* energy, float* featureGrad, cudaStream_t stream) {
    NNP_REQUIRE(fused(), "SpeciesMlp::forwardBackward needs the fused chain kernel");
    NNP_CUDA_CHECK(cudaMemsetAsync(energyAcc_, 0, sizeof(double), stream));
    if (chain2_) chain2_->launch(energyAcc_, featureGrad, kGradScale / M_, 1.0f / kGradScale, stream);
    else chain_->launch(energyAcc_, featureGrad, kGradScale / M_, 1.0f / kGradScale, stream);
    finish_energy_kernel<<<1, 1, 0, stream>>>(energyAcc_, energyBias_, M_, energy);
    count_launch();
    NNP_CUDA_CHECK(cudaGetLastError());
    haveForward_ = true;
}

void SpeciesMlp::forwardTc(const float* features, float* energy, cudaStream_t stream) {
    NNP_REQUIRE(featStride_ % 8 == 0, "feature stride must be a multiple of 8");
    if (rows_ > 0 && features != nullptr) {
        const size_t n8 = (size_t)rows_ * featStride_ / 8;
        split_rows_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, stream>>>(features, n8, featHi_, featLo_);
        count_launch();
    }
    for (int s = 0; s < S_; s++) {
        const int r0 = rowStart_[s], nr = rowStart_[s + 1] - r0;
        if (nr > 0) forwardRowsTc(s, r0, nr, r0, stream);
    }
}

// rows [r0, r0 + nr) of species s; activations go to rows [w0, w0 + nr) of the work buffers
void SpeciesMlp::forwardRowsTc(int s, int r0, int nr, int w0, cudaStream_t stream) {
    for (int l = 0; l < L_ - 1; l++) {
        const Layer& ly = layers_[s][l];
        GemmArgsH g;
        std::memset(&g, 0, sizeof(g));
        g.M = nr; g.epilogue = 1; g.bias = ly.b; g.K = ly.inP; g.outScale = 1.0f;
        g.Bhi = ly.Whi; g.Blo = ly.Wlo; g.ldb = ly.inP; g.bRows = M_ * ly.outP;
        g.Chi = actHi_[l] + (size_t)w0 * width_[l]; g.Clo = actLo_[l] + (size_t)w0 * width_[l]; g.ldc = width_[l];
        if (l == L_ - 2) {   // last hidden layer: energy and the backward seed come straight out of the epilogue
            g.epilogue = 3; g.w3 = layers_[s][L_ - 1].W; g.energyAcc = energyAcc_; g.seedScale = kGradScale / M_;
            g.Chi = dzHi_[l] + (size_t)w0 * width_[l]; g.Clo = dzLo_[l] + (size_t)w0 * width_[l];
        }
        if (l == 0) {
            g.Ahi = featHi_ + (size_t)r0 * featStride_; g.Alo = featLo_ + (size_t)r0 * featStride_; g.lda = featStride_;
            g.aCols = featStride_; g.aBatchCols = 0; g.bBatchRows = 0; g.N = M_ * ly.outP; g.batch = 1; g.cBatchCols = 0; g.biasBatch = 0;
        } else {
            g.Ahi = actHi_[l - 1] + (size_t)w0 * width_[l - 1]; g.Alo = actLo_[l - 1] + (size_t)w0 * width_[l - 1];
            g.lda = width_[l - 1]; g.aCols = width_[l - 1]; g.aBatchCols = ly.inP; g.bBatchRows = ly.outP; g.N = ly.outP; g.batch = M_;
            g.cBatchCols = ly.outP; g.biasBatch = ly.outP;
        }
        launch_gemm_tcgen05(g, stream);
    }
}

void SpeciesMlp::backwardTc(float* featureGrad, cudaStream_t stream) {
    for (int s = 0; s < S_; s++) {
        const int r0 = rowStart_[s], nr = rowStart_[s + 1] - r0;
        if (nr > 0) backwardRowsTc(s, r0, nr, r0, featureGrad, stream);
    }
}

void SpeciesMlp::backwardRowsTc(int s, int r0, int nr, int w0, float* featureGrad, cudaStream_t stream) {
    for (int l = L_ - 2; l >= 1; l--) {
        const Layer& ly = layers_[s][l];
        GemmArgsH g;
        std::memset(&g, 0, sizeof(g));
        g.M = nr; g.N = ly.inP; g.K = ly.outP; g.batch = M_; g.epilogue = 2; g.outScale = 1.0f;
        g.Ahi = dzHi_[l] + (size_t)w0 * width_[l]; g.Alo = dzLo_[l] + (size_t)w0 * width_[l]; g.lda = width_[l]; g.aCols = width_[l];
        g.aBatchCols = ly.outP;
        g.Bhi = ly.Wthi; g.Blo = ly.Wtlo; g.ldb = ly.outP; g.bRows = M_ * ly.inP; g.bBatchRows = ly.inP;
        g.Chi = dzHi_[l - 1] + (size_t)w0 * width_[l - 1]; g.Clo = dzLo_[l - 1] + (size_t)w0 * width_[l - 1]; g.ldc = width_[l - 1];
        g.cBatchCols = ly.inP;
        g.actHi = actHi_[l - 1] + (size_t)w0 * width_[l - 1]; g.actLo = actLo_[l - 1] + (size_t)w0 * width_[l - 1];
        g.ldact = width_[l - 1]; g.actBatchCols = ly.inP;
        launch_gemm_tcgen05(g, stream);
    }
    {
        const Layer& ly = layers_[s][0];
        GemmArgsH g;
        std::memset(&g, 0, sizeof(g));
        g.M = nr; g.N = ly.inP; g.K = M_ * ly.outP; g.batch = 1; g.epilogue = 0; g.outScale = 1.0f / kGradScale;
        g.Ahi = dzHi_[0] + (size_t)w0 * width_[0]; g.Alo = dzLo_[0] + (size_t)w0 * width_[0]; g.lda = width_[0]; g.aCols = width_[0];
        g.Bhi = ly.Wthi; g.Blo = ly.Wtlo; g.ldb = M_ * ly.outP; g.bRows = ly.inP;
        g.C32 = featureGrad + (size_t)r0 * featStride_; g.ldc = featStride_;
        launch_gemm_tcgen05(g, stream);
    }
}

void SpeciesMlp::forward(const float* features, float* energy, cudaStream_t stream) {
    NNP_CUDA_CHECK(cudaMemsetAsync(energyAcc_, 0, sizeof(double), stream));
    if (impl_ == MlpImpl::Tcgen05) {
        forwardTc(features, energy, stream);
        finish_energy_kernel<<<1, 1, 0, stream>>>(energyAcc_, energyBias_, M_, energy);
        count_launch();
        NNP_CUDA_CHECK(cudaGetLastError());
        haveForward_ = true;
        return;
    }
    for (int s = 0; s < S_; s++) {
        const int r0 = rowStart_[s], nr = rowStart_[s + 1] - r0;
        if (nr == 0) continue;
        for (int l = 0; l < L_ - 1; l++) {
            const Layer& ly = layers_[s][l];
            GemmArgs g;
            std::memset(&g, 0, sizeof(g));
            g.M = nr; g.epilogue = 1; g.bias = ly.b; g.ldb = ly.inP; g.K = ly.inP;
            g.C = act_[l] + (size_t)r0 * width_[l]; g.ldc = width_[l];
            if (l == 0) {   // all ensemble members side by side: N = M * outP
                g.A = features + (size_t)r0 * featStride_; g.lda = featStride_; g.aBatchCols = 0;
                g.B = ly.W; g.bBatch = 0; g.N = M_ * ly.outP; g.batch = 1; g.cBatchCols = 0; g.biasBatch = 0;
            } else {
                g.A = act_[l - 1] + (size_t)r0 * width_[l - 1]; g.lda = width_[l - 1]; g.aBatchCols = ly.inP;
                g.B = ly.W; g.bBatch = (long long)ly.outP * ly.inP; g.N = ly.outP; g.batch = M_; g.cBatchCols = ly.outP;
                g.biasBatch = ly.outP;
            }
            launch_gemm_simt(g, stream);
        }
        const Layer& last = layers_[s][L_ - 1];
        const int warps = nr * M_;
        final_energy_kernel<<<(warps * 32 + 255) / 256, 256, 0, stream>>>(act_[L_ - 2] + (size_t)r0 * width_[L_ - 2], width_[L_ - 2], nr, M_,
                                                                           last.inP, last.W, last.b, energyAcc_);
        count_launch();
    }
    finish_energy_kernel<<<1, 1, 0, stream>>>(energyAcc_, 0.0, M_, energy);
    count_launch();
    NNP_CUDA_CHECK(cudaGetLastError());
    haveForward_ = true;
}

void SpeciesMlp::backward(float* featureGrad, cudaStream_t stream) {
    NNP_REQUIRE(haveForward_, "SpeciesMlp::backward called before forward");
    if (impl_ == MlpImpl::Tcgen05) {
        backwardTc(featureGrad, stream);
        NNP_CUDA_CHECK(cudaGetLastError());
        return;
    }
    for (int s = 0; s < S_; s++) {
        const int r0 = rowStart_[s], nr = rowStart_[s + 1] - r0;
        if (nr == 0) continue;
        {   // seed: gradient w.r.t. the pre-activation of the last hidden layer
            const Layer& last = layers_[s][L_ - 1];
            const int l = L_ - 2;
            const size_t tot = (size_t)nr * M_ * last.inP;
            backward_seed_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(act_[l] + (size_t)r0 * width_[l], width_[l], nr, M_, last.inP,
                                                                                    last.W, dz_[l] + (size_t)r0 * width_[l], width_[l]);
            count_launch();
        }
        for (int l = L_ - 2; l >= 1; l--) {   // dZ_{l-1} = (dZ_l . W_l) * celu'(A_{l-1}),  per member
            const Layer& ly = layers_[s][l];
            GemmArgs g;
            std::memset(&g, 0, sizeof(g));
            g.M = nr; g.N = ly.inP; g.K = ly.outP; g.batch = M_; g.epilogue = 2;
            g.A = dz_[l] + (size_t)r0 * width_[l]; g.lda = width_[l]; g.aBatchCols = ly.outP;
            g.B = ly.Wt; g.ldb = ly.outP; g.bBatch = (long long)ly.inP * ly.outP;
            g.C = dz_[l - 1] + (size_t)r0 * width_[l - 1]; g.ldc = width_[l - 1]; g.cBatchCols = ly.inP;
            g.act = act_[l - 1] + (size_t)r0 * width_[l - 1]; g.ldact = width_[l - 1]; g.actBatchCols = ly.inP;
            launch_gemm_simt(g, stream);
        }
        {   // dX = dZ_0 . W_0  (sum over ensemble members is part of the contraction: K = M * outP)
            const Layer& ly = layers_[s][0];
            GemmArgs g;
            std::memset(&g, 0, sizeof(g));
            g.M = nr; g.N = ly.inP; g.K = M_ * ly.outP; g.batch = 1; g.epilogue = 0;
            g.A = dz_[0] + (size_t)r0 * width_[0]; g.lda = width_[0];
            g.B = ly.Wt; g.ldb = M_ * ly.outP;
            g.C = featureGrad + (size_t)r0 * featStride_; g.ldc = featStride_;
            launch_gemm_simt(g, stream);
        }
    }
    NNP_CUDA_CHECK(cudaGetLastError());
}

}  // namespace nnpops
