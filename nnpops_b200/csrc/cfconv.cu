// SchNet continuous-filter convolution (CFConv) and its neighbour list on sm_100a.  Replaces the reference classes
// CudaCFConvNeighbors / CudaCFConv (src/schnet/CudaCFConv.cu:94-188 and :283-528, kernels K6-K8 of SURVEY.md section 2.3) behind the
// abstract interfaces of src/schnet/CFConv.h:37-85,109-217.
//
// Design.  The filter  f_c(r) = fc(r) * (b2 + W2 act(b1 + W1 gauss(r)))_c  depends on the pair only through the scalar distance,
// so it is evaluated ONCE per object on a fine radial grid (fp64 setup kernel, 32 points per Gaussian width) and each pair
// interpolates it with a cubic Hermite spline (value and derivative tables).  That replaces the 46.9 kflop of dense layers per
// pair (CpuCFConv.cpp:150-178) by 4 table rows and ~40 flop per feature quadruple, with an interpolation error far below the
// 1e-5 parity tolerance (tests/test_cfconv_gpu.py).  The neighbour list is a FULL (both directions) CSR list built from the
// cell list, so forward and backward are gather-only: one warp per centre atom, no atomics, deterministic.
#include <cmath>
#include <vector>
#include "cell_list.cuh"

namespace nnpops {

namespace {

constexpr int kWPB = 8;

// ------------------------------------------------------------------------------------------------------------------ neighbours
// MODE 0: count neighbours of each sorted atom (strict r2 < rc2, CpuCFConv.cpp:105-113); MODE 1: fill the CSR row
template <int MODE>
__global__ void __launch_bounds__(kWPB * 32)
cf_neighbors_kernel(int n, const float4* __restrict__ sorted, const int* __restrict__ sortedCell, const Geom* __restrict__ geom,
                    const int* __restrict__ cellStart, float cutoff2, int* __restrict__ counts, const int* __restrict__ rowPtr,
                    int* __restrict__ nbr, long long capacity) {
    __shared__ Geom g;
    if (threadIdx.x == 0) g = *geom;
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = blockIdx.x * kWPB + w;
    if (p >= n) return;
    const float4 ci = sorted[p];
    long long cursor = MODE == 1 ? rowPtr[p] : 0;
    int mine = 0;
    for_each_candidate_run(g, cellStart, sortedCell[p], [&](int b, int e) {
        for (int q0 = b; q0 < e; q0 += 32) {
            const int q = q0 + lane;
            bool ok = false;
            if (q < e && q != p) {
                const float4 cj = sorted[q];
                float dx = __fsub_rn(cj.x, ci.x), dy = __fsub_rn(cj.y, ci.y), dz = __fsub_rn(cj.z, ci.z);
                ok = min_image_mul(g, dx, dy, dz) < cutoff2;
            }
            const unsigned m = __ballot_sync(kFull, ok);
            if (MODE == 1 && ok) {
                const long long slot = cursor + __popc(m & ((1u << lane) - 1u));
                if (slot < capacity) nbr[slot] = q;
            }
            cursor += __popc(m);
            mine += __popc(m);
        }
    });
    if (MODE == 0 && lane == 0) counts[p] = mine;
}

__global__ void cf_scan_kernel(const int* __restrict__ counts, int n, int* __restrict__ rowPtr, long long* __restrict__ total) {
    __shared__ long long warpTot[32];
    __shared__ long long carry;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    constexpr int IPT = 8;
    for (int base = 0; base < n; base += blockDim.x * IPT) {
        const int i0 = base + threadIdx.x * IPT;
        int v[IPT];
        long long s = 0;
#pragma unroll
        for (int k = 0; k < IPT; k++) { v[k] = (i0 + k < n) ? counts[i0 + k] : 0; s += v[k]; }
        long long x = s;
        for (int o = 1; o < 32; o <<= 1) { long long y = __shfl_up_sync(kFull, x, o); if (lane >= o) x += y; }
        if (lane == 31) warpTot[w] = x;
        __syncthreads();
        if (w == 0) {
            long long t = lane < nw ? warpTot[lane] : 0;
            for (int o = 1; o < 32; o <<= 1) { long long y = __shfl_up_sync(kFull, t, o); if (lane >= o) t += y; }
            warpTot[lane] = t;
        }
        __syncthreads();
        long long run = carry + (w > 0 ? warpTot[w - 1] : 0) + x - s;
#pragma unroll
        for (int k = 0; k < IPT; k++) { if (i0 + k < n) rowPtr[i0 + k] = (int)run; run += v[k]; }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = run;
        __syncthreads();
    }
    if (threadIdx.x == 0) { rowPtr[n] = (int)carry; *total = carry; }
}

// ------------------------------------------------------------------------------------------------------------------ filter table
// One CTA per table point r_p = p*h: Gaussians -> dense1 -> activation -> dense2 -> cutoff, value and d/dr, in fp64.
__global__ void cf_table_kernel(int P, int W, int G, double rc, double sigma, int activation, const float* __restrict__ w1,
                                const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2, double h,
                                double* __restrict__ tabF64, float* __restrict__ tabD) {
    extern __shared__ double sh[];
    double *gs = sh, *dgs = gs + G, *y1 = dgs + G, *dy1 = y1 + W;
    const int p = blockIdx.x;
    const double r = p * h;
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        const double mu = (double)((float)g * (float)rc / (float)(G - 1));   // centres formed in fp32 like CpuCFConv.cpp:121-122
        const double x = (r - mu) / sigma;
        gs[g] = exp(-0.5 * x * x);
        dgs[g] = -x * gs[g] / sigma;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < W; i += blockDim.x) {
        double s = b1[i], ds = 0;
        for (int g = 0; g < G; g++) { s += gs[g] * (double)w1[(size_t)i * G + g]; ds += dgs[g] * (double)w1[(size_t)i * G + g]; }
        if (activation == 0) { const double e = exp(s); y1[i] = log(0.5 * e + 0.5); dy1[i] = ds * e / (e + 1.0); }
        else { const double t = tanh(s); y1[i] = t; dy1[i] = ds * (1.0 - t * t); }
    }
    __syncthreads();
    const double kPiD = 3.14159265358979323846;
    const double fc = 0.5 * cos(kPiD * r / rc) + 0.5, dfc = -(0.5 * kPiD / rc) * sin(kPiD * r / rc);
    for (int i = threadIdx.x; i < W; i += blockDim.x) {
        double s = b2[i], ds = 0;
        for (int j = 0; j < W; j++) { s += y1[j] * (double)w2[(size_t)i * W + j]; ds += dy1[j] * (double)w2[(size_t)i * W + j]; }
        tabF64[(size_t)p * W + i] = fc * s;
        tabD[(size_t)p * W + i] = (float)((dfc * s + fc * ds) * h);   // derivative pre-multiplied by the grid spacing
    }
}

// fp32 value table and the table of forward differences G[p] = f[p + 1] - f[p], formed in fp64.  The interpolant is evaluated as
// f0 + h01 * G + ... and its derivative as g01 * G + ...: with f1 - f0 taken from two rounded fp32 entries instead, the derivative
// (the force) loses |f| * eps / h to cancellation -- 1.5e-5 of the position gradient at h = sigma / 32, 6e-5 at sigma / 128.
__global__ void cf_table_finish_kernel(int P, int W, const double* __restrict__ tabF64, float* __restrict__ tabF, float* __restrict__ tabG) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)P * W) return;
    tabF[i] = (float)tabF64[i];
    tabG[i] = i + W < (size_t)P * W ? (float)(tabF64[i + W] - tabF64[i]) : 0.0f;
}

// ------------------------------------------------------------------------------------------------------------------ compute
struct Hermite {
    int idx;
    float h00, h10, h01, h11;      // value basis
    float g00, g10, g01, g11;      // derivative basis (divided by h)
};

__device__ __forceinline__ Hermite hermite(float r, float invH, int P) {
    Hermite b;
    const float t = r * invH;
    b.idx = min((int)t, P - 2);
    const float u = t - (float)b.idx, u2 = u * u, om = 1.0f - u;
    // h00 = 1 - h01 and g00 = -g01: the value is f0 + h01 (f1 - f0) + h10 d0 + h11 d1 with the difference read from its own table
    b.h00 = 1.0f; b.h10 = u * om * om; b.h01 = u2 * (3.0f - 2.0f * u); b.h11 = u2 * (u - 1.0f);
    b.g00 = 0.0f; b.g10 = (3.0f * u2 - 4.0f * u + 1.0f) * invH;
    b.g01 = (6.0f * u - 6.0f * u2) * invH; b.g11 = (3.0f * u2 - 2.0f * u) * invH;
    return b;
}

// out[i][c] = sum_j f_c(r_ij) x[j][c]  over the full neighbour row of i (covers both scatter directions of CpuCFConv.cpp:182-185).
// lane owns features {c0 + lane + 32 k}; the distance of 32 neighbours is computed in parallel and broadcast by shuffle.
template <int KF>
__global__ void __launch_bounds__(kWPB * 32)
cf_forward_kernel(int n, int W, const float4* __restrict__ sorted, const int* __restrict__ sortedOrig, const Geom* __restrict__ geom,
                  const int* __restrict__ rowPtr, const int* __restrict__ nbr, const float* __restrict__ tabF, const float* __restrict__ tabD,
                  const float* __restrict__ tabG, float invH, int P, const float* __restrict__ x, float* __restrict__ out) {
    __shared__ Geom g;
    if (threadIdx.x == 0) g = *geom;
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = blockIdx.x * kWPB + w;
    if (p >= n) return;
    const float4 ci = sorted[p];
    const int b = rowPtr[p], e = rowPtr[p + 1];
    const int orig = sortedOrig[p];
    for (int c0 = 0; c0 < W; c0 += 32 * KF) {
        float acc[KF];
#pragma unroll
        for (int k = 0; k < KF; k++) acc[k] = 0.0f;
        for (int q0 = b; q0 < e; q0 += 32) {
            const int q = q0 + lane;
            float r = 0.0f;
            int oj = 0;
            if (q < e) {
                const int j = nbr[q];
                const float4 cj = sorted[j];
                float dx = __fsub_rn(cj.x, ci.x), dy = __fsub_rn(cj.y, ci.y), dz = __fsub_rn(cj.z, ci.z);
                r = sqrtf(min_image_mul(g, dx, dy, dz));
                oj = sortedOrig[j];
            }
            const int cnt = min(32, e - q0);
            for (int t = 0; t < cnt; t++) {
                const float rt = __shfl_sync(kFull, r, t);
                const int ot = __shfl_sync(kFull, oj, t);
                const Hermite hb = hermite(rt, invH, P);
                const float* f0 = tabF + (size_t)hb.idx * W;
                const float* d0 = tabD + (size_t)hb.idx * W;
                const float* g0 = tabG + (size_t)hb.idx * W;
                const float* xr = x + (size_t)ot * W;
#pragma unroll
                for (int k = 0; k < KF; k++) {
                    const int c = c0 + lane + 32 * k;
                    if (c < W) {
                        const float f = f0[c] + hb.h10 * d0[c] + hb.h01 * g0[c] + hb.h11 * d0[W + c];
                        acc[k] = fmaf(f, xr[c], acc[k]);
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < KF; k++) {
            const int c = c0 + lane + 32 * k;
            if (c < W) out[(size_t)orig * W + c] = acc[k];
        }
    }
}

// inputGrad[i][c] = sum_j f_c(r) go[j][c];   posGrad[i] = - sum_j (1/r) sum_c f'_c(r) (x[j][c] go[i][c] + x[i][c] go[j][c]) delta_ij
// (CpuCFConv.cpp:283-296 in gather form: every directed pair is evaluated by its centre, so no atomics)
template <int KF>
__global__ void __launch_bounds__(kWPB * 32)
cf_backward_kernel(int n, int W, const float4* __restrict__ sorted, const int* __restrict__ sortedOrig, const Geom* __restrict__ geom,
                   const int* __restrict__ rowPtr, const int* __restrict__ nbr, const float* __restrict__ tabF, const float* __restrict__ tabD,
                   const float* __restrict__ tabG, float invH, int P, const float* __restrict__ x, const float* __restrict__ go, float* __restrict__ inputGrad,
                   float* __restrict__ posGrad) {
    __shared__ Geom g;
    if (threadIdx.x == 0) g = *geom;
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = blockIdx.x * kWPB + w;
    if (p >= n) return;
    const float4 ci = sorted[p];
    const int b = rowPtr[p], e = rowPtr[p + 1];
    const int orig = sortedOrig[p];
    float fx = 0.0f, fy = 0.0f, fz = 0.0f;
    for (int c0 = 0; c0 < W; c0 += 32 * KF) {
        float acc[KF], xi[KF], gi[KF];
#pragma unroll
        for (int k = 0; k < KF; k++) {
            const int c = c0 + lane + 32 * k;
            acc[k] = 0.0f;
            xi[k] = c < W ? x[(size_t)orig * W + c] : 0.0f;
            gi[k] = c < W ? go[(size_t)orig * W + c] : 0.0f;
        }
        for (int q0 = b; q0 < e; q0 += 32) {
            const int q = q0 + lane;
            float r = 1.0f, ux = 0.0f, uy = 0.0f, uz = 0.0f;
            int oj = 0;
            if (q < e) {
                const int j = nbr[q];
                const float4 cj = sorted[j];
                float dx = __fsub_rn(cj.x, ci.x), dy = __fsub_rn(cj.y, ci.y), dz = __fsub_rn(cj.z, ci.z);
                r = sqrtf(min_image_mul(g, dx, dy, dz));
                const float ir = 1.0f / r;
                ux = dx * ir; uy = dy * ir; uz = dz * ir;
                oj = sortedOrig[j];
            }
            const int cnt = min(32, e - q0);
            float wmine = 0.0f;   // lane t ends up holding the pair weight of neighbour q0 + t
            for (int t = 0; t < cnt; t++) {
                const float rt = __shfl_sync(kFull, r, t);
                const int ot = __shfl_sync(kFull, oj, t);
                const Hermite hb = hermite(rt, invH, P);
                const float* f0 = tabF + (size_t)hb.idx * W;
                const float* d0 = tabD + (size_t)hb.idx * W;
                const float* g0 = tabG + (size_t)hb.idx * W;
                const float* xr = x + (size_t)ot * W;
                const float* gr = go + (size_t)ot * W;
                float wsum = 0.0f;
#pragma unroll
                for (int k = 0; k < KF; k++) {
                    const int c = c0 + lane + 32 * k;
                    if (c < W) {
                        const float a0 = f0[c], a1 = d0[c], a2 = g0[c], a3 = d0[W + c];   // f[p], h f'[p], f[p + 1] - f[p], h f'[p + 1]
                        const float f = a0 + hb.h10 * a1 + hb.h01 * a2 + hb.h11 * a3;
                        const float df = hb.g10 * a1 + hb.g01 * a2 + hb.g11 * a3;
                        const float gj = gr[c];
                        acc[k] = fmaf(f, gj, acc[k]);
                        wsum = fmaf(df, xr[c] * gi[k] + xi[k] * gj, wsum);
                    }
                }
                wsum = warp_sum(wsum);
                if (lane == t) wmine = wsum;
            }
            fx = fmaf(wmine, ux, fx); fy = fmaf(wmine, uy, fy); fz = fmaf(wmine, uz, fz);
        }
#pragma unroll
        for (int k = 0; k < KF; k++) {
            const int c = c0 + lane + 32 * k;
            if (c < W) inputGrad[(size_t)orig * W + c] = acc[k];
        }
    }
    fx = warp_sum(fx); fy = warp_sum(fy); fz = warp_sum(fz);
    if (lane == 0) { posGrad[3 * (size_t)orig] = -fx; posGrad[3 * (size_t)orig + 1] = -fy; posGrad[3 * (size_t)orig + 2] = -fz; }
}

}  // namespace

class CFConvNeighborList {
public:
    CFConvNeighborList(int numAtoms, float cutoff) : n_(numAtoms), cutoff_(cutoff) {
        NNP_REQUIRE(numAtoms >= 0, "numAtoms must be non-negative");
        NNP_REQUIRE(cutoff > 0, "cutoff must be positive");
        cells_.init(n_);
        const size_t na = (size_t)(n_ > 0 ? n_ : 1);
        NNP_CUDA_CHECK(cudaMalloc(&counts_, sizeof(int) * na));
        NNP_CUDA_CHECK(cudaMalloc(&rowPtr_, sizeof(int) * (na + 1)));
        NNP_CUDA_CHECK(cudaMalloc(&total_, sizeof(long long)));
        NNP_CUDA_CHECK(cudaMallocHost(&totalHost_, sizeof(long long)));
    }
    ~CFConvNeighborList() {
        cells_.release();
        cudaFree(counts_); cudaFree(rowPtr_); cudaFree(nbr_); cudaFree(total_); cudaFreeHost(totalHost_);
    }
    // Builds the full neighbour list.  Reads back one integer (the pair count) to size the list: the only host sync of the path.
    void build(const float* positions, const float* box, cudaStream_t stream) {
        periodic_ = box != nullptr;
        built_ = true;
        pairs_ = 0;
        if (n_ == 0) return;
        cells_.build<float>(positions, box, nullptr, cutoff_, stream);
        const int grid = (n_ + kWPB - 1) / kWPB;
        const float c2 = cutoff_ * cutoff_;
        cf_neighbors_kernel<0><<<grid, kWPB * 32, 0, stream>>>(n_, cells_.sorted, cells_.sortedCell, cells_.geom, cells_.cellStart, c2, counts_,
                                                               rowPtr_, nbr_, 0);
        cf_scan_kernel<<<1, 1024, 0, stream>>>(counts_, n_, rowPtr_, total_);
        NNP_CUDA_CHECK(cudaMemcpyAsync(totalHost_, total_, sizeof(long long), cudaMemcpyDeviceToHost, stream));
        NNP_CUDA_CHECK(cudaStreamSynchronize(stream));
        const long long total = *totalHost_;
        NNP_REQUIRE(total < 0x7fffffffLL, "neighbour list too long for 32-bit row offsets");
        if ((size_t)total > cap_) {
            cudaFree(nbr_);
            cap_ = (size_t)(total * 1.25) + 1024;
            NNP_CUDA_CHECK(cudaMalloc(&nbr_, sizeof(int) * cap_));
        }
        cf_neighbors_kernel<1><<<grid, kWPB * 32, 0, stream>>>(n_, cells_.sorted, cells_.sortedCell, cells_.geom, cells_.cellStart, c2, counts_,
                                                               rowPtr_, nbr_, (long long)cap_);
        count_launch(3);
        NNP_CUDA_CHECK(cudaGetLastError());
        pairs_ = total / 2;
    }
    int n_;
    float cutoff_;
    bool periodic_ = false, built_ = false;
    long long pairs_ = 0;
    CellList cells_;
    int* counts_ = nullptr;
    int* rowPtr_ = nullptr;
    int* nbr_ = nullptr;
    size_t cap_ = 0;
    long long* total_ = nullptr;
    long long* totalHost_ = nullptr;
};

class CFConvFilter {
public:
    CFConvFilter(int width, int numGaussians, float cutoff, float gaussianWidth, int activation, const float* w1, const float* b1,
                 const float* w2, const float* b2 /* device pointers */, int pointsPerSigma)
        : W_(width), G_(numGaussians), cutoff_(cutoff) {
        NNP_REQUIRE(width > 0 && numGaussians > 1, "width must be positive and numGaussians > 1");
        NNP_REQUIRE(cutoff > 0 && gaussianWidth > 0, "cutoff and gaussianWidth must be positive");
        NNP_REQUIRE(activation == 0 || activation == 1, "activation must be 0 (shifted softplus) or 1 (tanh)");
        // table step = sigma / 64: the cubic Hermite value error goes as h^4, the error of its derivative (the force) as h^3
        // (3 201 rows x W x 3 tables for sigma 0.2 / cutoff 10: L2-resident)
        const int pps = pointsPerSigma > 0 ? pointsPerSigma : 64;
        long long P = (long long)std::ceil((double)pps * cutoff / gaussianWidth) + 1;
        if (P < 256) P = 256;
        if (P > 65536) P = 65536;
        P_ = (int)P;
        const double h = (double)cutoff / (P_ - 1);
        invH_ = (float)(1.0 / h);
        NNP_CUDA_CHECK(cudaMalloc(&tabF_, sizeof(float) * (size_t)P_ * W_));
        NNP_CUDA_CHECK(cudaMalloc(&tabD_, sizeof(float) * (size_t)P_ * W_));
        NNP_CUDA_CHECK(cudaMalloc(&tabG_, sizeof(float) * (size_t)P_ * W_));
        double* f64 = nullptr;
        NNP_CUDA_CHECK(cudaMalloc(&f64, sizeof(double) * (size_t)P_ * W_));
        const size_t smem = sizeof(double) * (2 * (size_t)G_ + 2 * (size_t)W_);
        NNP_REQUIRE(smem <= 48 * 1024, "width/numGaussians too large for the filter-table kernel");
        // one-time set-up on the legacy default stream: the weights may have been staged on a non-blocking stream of the caller, which
        // the default stream does not order against, so wait for the whole device first
        NNP_CUDA_CHECK(cudaDeviceSynchronize());
        cf_table_kernel<<<P_, 128, smem>>>(P_, W_, G_, (double)cutoff, (double)gaussianWidth, activation, w1, b1, w2, b2, h, f64, tabD_);
        cf_table_finish_kernel<<<(unsigned)(((size_t)P_ * W_ + 255) / 256), 256>>>(P_, W_, f64, tabF_, tabG_);
        NNP_CUDA_CHECK(cudaGetLastError());
        NNP_CUDA_CHECK(cudaDeviceSynchronize());
        cudaFree(f64);
    }
    ~CFConvFilter() { cudaFree(tabF_); cudaFree(tabD_); cudaFree(tabG_); }

    void compute(const CFConvNeighborList& nb, const float* input, float* output, cudaStream_t stream) const {
        check(nb);
        if (nb.n_ == 0) return;
        const int grid = (nb.n_ + kWPB - 1) / kWPB;
        if (W_ <= 32) cf_forward_kernel<1><<<grid, kWPB * 32, 0, stream>>>(nb.n_, W_, nb.cells_.sorted, nb.cells_.sortedOrig, nb.cells_.geom, nb.rowPtr_, nb.nbr_, tabF_, tabD_, tabG_, invH_, P_, input, output);
        else if (W_ <= 64) cf_forward_kernel<2><<<grid, kWPB * 32, 0, stream>>>(nb.n_, W_, nb.cells_.sorted, nb.cells_.sortedOrig, nb.cells_.geom, nb.rowPtr_, nb.nbr_, tabF_, tabD_, tabG_, invH_, P_, input, output);
        else cf_forward_kernel<4><<<grid, kWPB * 32, 0, stream>>>(nb.n_, W_, nb.cells_.sorted, nb.cells_.sortedOrig, nb.cells_.geom, nb.rowPtr_, nb.nbr_, tabF_, tabD_, tabG_, invH_, P_, input, output);
        count_launch();
        NNP_CUDA_CHECK(cudaGetLastError());
    }
    void backprop(const CFConvNeighborList& nb, const float* input, const float* outputGrad, float* inputGrad, float* posGrad,
                  cudaStream_t stream) const {
        check(nb);
        if (nb.n_ == 0) return;
        const int grid = (nb.n_ + kWPB - 1) / kWPB;
        if (W_ <= 32) cf_backward_kernel<1><<<grid, kWPB * 32, 0, stream>>>(nb.n_, W_, nb.cells_.sorted, nb.cells_.sortedOrig, nb.cells_.geom, nb.rowPtr_, nb.nbr_, tabF_, tabD_, tabG_, invH_, P_, input, outputGrad, inputGrad, posGrad);
        else if (W_ <= 64) cf_backward_kernel<2><<<grid, kWPB * 32, 0, stream>>>(nb.n_, W_, nb.cells_.sorted, nb.cells_.sortedOrig, nb.cells_.geom, nb.rowPtr_, nb.nbr_, tabF_, tabD_, tabG_, invH_, P_, input, outputGrad, inputGrad, posGrad);
        else cf_backward_kernel<4><<<grid, kWPB * 32, 0, stream>>>(nb.n_, W_, nb.cells_.sorted, nb.cells_.sortedOrig, nb.cells_.geom, nb.rowPtr_, nb.nbr_, tabF_, tabD_, tabG_, invH_, P_, input, outputGrad, inputGrad, posGrad);
        count_launch();
        NNP_CUDA_CHECK(cudaGetLastError());
    }
    int width() const { return W_; }
    float cutoff() const { return cutoff_; }

private:
    void check(const CFConvNeighborList& nb) const {
        NNP_REQUIRE(nb.built_, "CFConvNeighbors.build() must be called before CFConv");
        NNP_REQUIRE(nb.cutoff_ == cutoff_, "CFConv and its neighbour list must use the same cutoff");
    }
    int W_, G_, P_;
    float cutoff_, invH_;
    float* tabF_ = nullptr;
    float* tabD_ = nullptr;
    float* tabG_ = nullptr;   // forward differences of the value table
};

// plain functions for the C ABI (c_api.cu)
CFConvNeighborList* cfconv_neighbors_create(int numAtoms, float cutoff) { return new CFConvNeighborList(numAtoms, cutoff); }
void cfconv_neighbors_destroy(CFConvNeighborList* p) { delete p; }
void cfconv_neighbors_build(CFConvNeighborList* p, const float* positions, const float* box, cudaStream_t s) { p->build(positions, box, s); }
long long cfconv_neighbors_pairs(const CFConvNeighborList* p) { return p->pairs_; }
CFConvFilter* cfconv_create(int width, int numGaussians, float cutoff, float gaussianWidth, int activation, const float* w1, const float* b1,
                            const float* w2, const float* b2, int pointsPerSigma) {
    return new CFConvFilter(width, numGaussians, cutoff, gaussianWidth, activation, w1, b1, w2, b2, pointsPerSigma);
}
void cfconv_destroy(CFConvFilter* p) { delete p; }
void cfconv_compute(const CFConvFilter* f, const CFConvNeighborList* nb, const float* input, float* output, cudaStream_t s) {
    f->compute(*nb, input, output, s);
}
void cfconv_backprop(const CFConvFilter* f, const CFConvNeighborList* nb, const float* input, const float* outputGrad, float* inputGrad,
                     float* posGrad, cudaStream_t s) {
    f->backprop(*nb, input, outputGrad, inputGrad, posGrad, s);
}

}  // namespace nnpops
