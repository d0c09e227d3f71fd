// Literal BatchedLinear (per-atom weights), kept for drop-in compatibility with NNPOpsBatchedNN::BatchedLinear
// (src/pytorch/BatchedNN.cpp:30-50).  HBM-bound: every weight is read exactly once per call, 16-byte vector loads, one warp per
// output row.  The scalable path is the species-grouped MLP (species_mlp.cu).
#include "common.cuh"

namespace nnpops {

namespace {

// out[a][m][o] = sum_i W[a][m][o][i] * v[a][m or 0][i] + b[a][m][o]
__global__ void batched_linear_fwd_kernel(const float* __restrict__ v, const float* __restrict__ W, const float* __restrict__ b,
                                          float* __restrict__ out, long long rows, int numModels, int vecModels, int nOut, int nIn) {
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // row = (a*numModels + m)*nOut + o
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const long long am = row / nOut;
    const long long a = am / numModels;
    const int m = (int)(am % numModels);
    const float* w = W + row * nIn;
    const float* x = v + (a * vecModels + (vecModels == 1 ? 0 : m)) * nIn;
    float s = 0.0f;
    if ((nIn & 3) == 0 && ((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(x)) & 15) == 0) {
        for (int i = lane * 4; i < nIn; i += 128) {
            const float4 wv = __ldcs(reinterpret_cast<const float4*>(w + i));
            const float4 xv = *reinterpret_cast<const float4*>(x + i);
            s = fmaf(wv.x, xv.x, s); s = fmaf(wv.y, xv.y, s); s = fmaf(wv.z, xv.z, s); s = fmaf(wv.w, xv.w, s);
        }
    } else {
        for (int i = lane; i < nIn; i += 32) s = fmaf(w[i], x[i], s);
    }
    s = warp_sum(s);
    if (lane == 0) out[row] = s + b[row];
}

// g[a][m][i] = sum_o go[a][m][o] * W[a][m][o][i]; one CTA per (a, m), threads over i
__global__ void batched_linear_bwd_kernel(const float* __restrict__ go, const float* __restrict__ W, float* __restrict__ g, int nOut,
                                          int nIn) {
    const long long am = blockIdx.x;
    const float* w = W + am * nOut * (long long)nIn;
    const float* gv = go + am * nOut;
    for (int i = threadIdx.x; i < nIn; i += blockDim.x) {
        float s = 0.0f;
        for (int o = 0; o < nOut; o++) s = fmaf(gv[o], __ldcs(w + (long long)o * nIn + i), s);
        g[am * nIn + i] = s;
    }
}

}  // namespace

void batched_linear_forward(const float* v, const float* W, const float* b, float* out, int numAtoms, int numModels, int vecModels,
                            int nOut, int nIn, cudaStream_t stream) {
    NNP_REQUIRE(vecModels == 1 || vecModels == numModels, "vectors must have 1 or num_models entries on the ensemble axis");
    const long long rows = (long long)numAtoms * numModels * nOut;
    if (rows == 0) return;
    const long long blocks = (rows * 32 + 255) / 256;
    batched_linear_fwd_kernel<<<(unsigned)blocks, 256, 0, stream>>>(v, W, b, out, rows, numModels, vecModels, nOut, nIn);
    count_launch();
    NNP_CUDA_CHECK(cudaGetLastError());
}

void batched_linear_backward(const float* go, const float* W, float* g, int numAtoms, int numModels, int nOut, int nIn,
                             cudaStream_t stream) {
    const long long am = (long long)numAtoms * numModels;
    if (am == 0 || nIn == 0) return;
    batched_linear_bwd_kernel<<<(unsigned)am, 256, 0, stream>>>(go, W, g, nOut, nIn);
    count_launch();
    NNP_CUDA_CHECK(cudaGetLastError());
}

}  // namespace nnpops
