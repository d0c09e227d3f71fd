// Particle-mesh Ewald: direct-space sum over a pair list and the reciprocal-space spread / FFT / convolve / gather.
// Replaces the reference kernels K11-K14 (src/pytorch/pme/pmeCUDA.cu:30-276) behind the ops pme::pme_direct and
// pme::pme_reciprocal (pmeCUDA.cu:278-418); arithmetic restated from SURVEY.md appendix A.4 and pmeCPU.cpp:74-353.
// The FFTs are cuFFT plans cached per grid shape (the reference calls cuFFT through torch::fft::rfftn/irfftn, pmeCUDA.cu:354,395).
#include <cufft.h>
#include <map>
#include <memory>
#include <mutex>
#include "workspace_cache.cuh"
#include <tuple>
#include "common.cuh"

namespace nnpops {

namespace {

constexpr float kTwoOverSqrtPi = 1.1283791670955126f;   // M_2_SQRTPI

// ------------------------------------------------------------------------------------------------------------------
// Direct space (pmeCPU.cpp:105-157).  One thread per pair; energy is accumulated per thread in double, reduced per CTA
// and added with one atomic per CTA.  dE/dx and dE/dq are accumulated with float reductions (red.global.add.f32).
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pme_direct_pairs_kernel(long long numPairs, const int* __restrict__ neighbors, const float* __restrict__ deltas,
                        const float* __restrict__ distances, const float* __restrict__ charges, const int* __restrict__ exclusions,
                        int maxExcl, float alpha, float coulomb, float* __restrict__ posDeriv, float* __restrict__ chargeDeriv,
                        double* __restrict__ energyAcc) {
    double energy = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < numPairs; i += (long long)gridDim.x * blockDim.x) {
        const int a1 = neighbors[i], a2 = neighbors[numPairs + i];
        bool include = a1 > -1;
        // exclusion rows are sorted descending (pme.py:92), so the scan stops at the first entry below a2
        for (int j = 0; include && j < maxExcl; j++) {
            const int x = exclusions[(size_t)a1 * maxExcl + j];
            if (x < a2) break;
            if (x == a2) include = false;
        }
        if (!include) continue;
        const float r = distances[i];
        const float invR = 1.0f / r;
        const float alphaR = alpha * r;
        const float expTerm = expf(-alphaR * alphaR);
        const float erfcTerm = erfcf(alphaR);
        const float pref = coulomb * invR;
        const float c1 = charges[a1], c2 = charges[a2];
        energy += (double)(pref * erfcTerm * c1 * c2);
        atomicAdd(&chargeDeriv[a1], pref * erfcTerm * c2);
        atomicAdd(&chargeDeriv[a2], pref * erfcTerm * c1);
        const float dEdR = pref * c1 * c2 * (erfcTerm + alphaR * expTerm * kTwoOverSqrtPi) * invR * invR;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float f = dEdR * deltas[3 * i + k];
            atomicAdd(&posDeriv[3 * (size_t)a1 + k], -f);
            atomicAdd(&posDeriv[3 * (size_t)a2 + k], f);
        }
    }
    __shared__ double part[8];
    for (int o = 16; o > 0; o >>= 1) energy += __shfl_xor_sync(kFull, energy, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = energy;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += part[w];
        if (t != 0.0) atomicAdd(energyAcc, t);
    }
}

// exclusion correction: subtract the erf part that reciprocal space adds for excluded pairs, using the NON-periodic
// displacement (pmeCPU.cpp:133-157, by design: pme.py:25-28)
__global__ void __launch_bounds__(256)
pme_direct_exclusions_kernel(int numAtoms, const float* __restrict__ pos, const float* __restrict__ charges,
                             const int* __restrict__ exclusions, int maxExcl, float alpha, float coulomb,
                             float* __restrict__ posDeriv, float* __restrict__ chargeDeriv, double* __restrict__ energyAcc) {
    double energy = 0.0;
    const long long total = (long long)numAtoms * maxExcl;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int a1 = (int)(idx / maxExcl);
        const int a2 = exclusions[idx];
        if (a2 <= a1) continue;   // rows are sorted descending: entries > a1 come first; each pair once
        float dr[3];
        for (int k = 0; k < 3; k++) dr[k] = pos[3 * (size_t)a1 + k] - pos[3 * (size_t)a2 + k];
        const float rr = sqrtf(dr[0] * dr[0] + dr[1] * dr[1] + dr[2] * dr[2]);
        const float invR = 1.0f / rr;
        const float alphaR = alpha * rr;
        const float expTerm = expf(-alphaR * alphaR);
        const float erfTerm = erff(alphaR);
        const float pref = coulomb * invR;
        const float c1 = charges[a1], c2 = charges[a2];
        energy -= (double)(pref * erfTerm * c1 * c2);
        atomicAdd(&chargeDeriv[a1], -pref * erfTerm * c2);
        atomicAdd(&chargeDeriv[a2], -pref * erfTerm * c1);
        const float dEdR = pref * c1 * c2 * (erfTerm - alphaR * expTerm * kTwoOverSqrtPi) * invR * invR;
        for (int k = 0; k < 3; k++) {
            atomicAdd(&posDeriv[3 * (size_t)a1 + k], dEdR * dr[k]);
            atomicAdd(&posDeriv[3 * (size_t)a2 + k], -dEdR * dr[k]);
        }
    }
    __shared__ double part[8];
    for (int o = 16; o > 0; o >>= 1) energy += __shfl_xor_sync(kFull, energy, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = energy;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += part[w];
        if (t != 0.0) atomicAdd(energyAcc, t);
    }
}

__global__ void publish_energy_kernel(const double* __restrict__ acc, double scale, float* __restrict__ out) { *out = (float)(*acc * scale); }

// ------------------------------------------------------------------------------------------------------------------
// Reciprocal space
// ------------------------------------------------------------------------------------------------------------------
struct RecipBox {
    float box[9];
    float recip[9];   // lower-triangular inverse (pmeCPU.cpp:11-25)
};

__device__ __forceinline__ void invert_box(const float* __restrict__ b, RecipBox& rb) {
    for (int i = 0; i < 9; i++) rb.box[i] = b[i];
    const float det = b[0] * b[4] * b[8];
    const float scale = 1.0f / det;
    rb.recip[0] = b[4] * b[8] * scale; rb.recip[1] = 0; rb.recip[2] = 0;
    rb.recip[3] = -b[3] * b[8] * scale; rb.recip[4] = b[0] * b[8] * scale; rb.recip[5] = 0;
    rb.recip[6] = (b[3] * b[7] - b[4] * b[6]) * scale; rb.recip[7] = -b[0] * b[7] * scale; rb.recip[8] = b[0] * b[4] * scale;
}

// Cardinal B-spline weights (and derivatives) of one atom along the three axes (pmeCPU.cpp:27-72): order-2 hat function raised
// to ORDER by the standard recursion; the derivative is the difference of the order-(ORDER-1) weights.
template <int ORDER, bool DERIV>
__device__ __forceinline__ void spline(const float* __restrict__ pos, int atom, const RecipBox& rb, const int* gridSize, int* gridIndex,
                                       float (&data)[3][ORDER], float (&ddata)[3][ORDER]) {
    float p[3] = {pos[3 * (size_t)atom], pos[3 * (size_t)atom + 1], pos[3 * (size_t)atom + 2]};
    for (int i = 2; i >= 0; i--) {
        const float s = floorf(p[i] * rb.recip[4 * i]);
        for (int j = 0; j < 3; j++) p[j] -= s * rb.box[3 * i + j];
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
        float t = p[0] * rb.recip[i] + p[1] * rb.recip[3 + i] + p[2] * rb.recip[6 + i];
        t = (t - floorf(t)) * gridSize[i];
        const int ti = (int)t;
        const float dr = t - ti;
        gridIndex[i] = ti % gridSize[i];
        float* d = data[i];
#pragma unroll
        for (int k = 0; k < ORDER; k++) d[k] = 0.0f;
        d[0] = 1.0f - dr; d[1] = dr;
#pragma unroll
        for (int j = 3; j <= ORDER; j++) {
            if (DERIV && j == ORDER) {
                ddata[i][0] = -d[0];
#pragma unroll
                for (int k = 1; k < ORDER; k++) ddata[i][k] = d[k - 1] - d[k];
            }
            const float div = 1.0f / (j - 1);
#pragma unroll
            for (int m = j - 1; m >= 0; m--) {
                const float lo = m > 0 ? d[m - 1] : 0.0f, hi = m < j - 1 ? d[m] : 0.0f;
                d[m] = div * ((dr + (float)(j - 1 - m)) * lo + ((float)(m + 1) - dr) * hi);
            }
        }
    }
}

template <int ORDER>
__global__ void __launch_bounds__(128)
pme_spread_kernel(int numAtoms, const float* __restrict__ pos, const float* __restrict__ charges, const float* __restrict__ box,
                  int gx, int gy, int gz, float sqrtCoulomb, float* __restrict__ grid) {
    __shared__ RecipBox rb;
    if (threadIdx.x == 0) invert_box(box, rb);
    __syncthreads();
    const int gridSize[3] = {gx, gy, gz};
    for (int atom = blockIdx.x * blockDim.x + threadIdx.x; atom < numAtoms; atom += gridDim.x * blockDim.x) {
        int gi[3];
        float data[3][ORDER], ddata[3][ORDER];
        spline<ORDER, false>(pos, atom, rb, gridSize, gi, data, ddata);
        const float q = charges[atom] * sqrtCoulomb;
#pragma unroll
        for (int ix = 0; ix < ORDER; ix++) {
            const int xi = (gi[0] + ix) % gx;
            const float dx = q * data[0][ix];
#pragma unroll
            for (int iy = 0; iy < ORDER; iy++) {
                const int yi = (gi[1] + iy) % gy;
                const float dxdy = dx * data[1][iy];
                float* row = grid + ((size_t)xi * gy + yi) * gz;
#pragma unroll
                for (int iz = 0; iz < ORDER; iz++) atomicAdd(row + (gi[2] + iz) % gz, dxdy * data[2][iz]);
            }
        }
    }
}

// multiply the half-complex grid by the Ewald kernel in place and accumulate the energy (pmeCPU.cpp:234-266)
__global__ void __launch_bounds__(256)
pme_convolve_kernel(float2* __restrict__ recip, const float* __restrict__ box, int gx, int gy, int gz, float alpha,
                    const float* __restrict__ xmod, const float* __restrict__ ymod, const float* __restrict__ zmod,
                    double* __restrict__ energyAcc) {
    __shared__ RecipBox rb;
    if (threadIdx.x == 0) invert_box(box, rb);
    __syncthreads();
    const int zsize = gz / 2 + 1, yzsize = gy * zsize;
    const long long total = (long long)gx * yzsize;
    const float scaleFactor = (float)(3.14159265358979323846 * (double)rb.box[0] * (double)rb.box[4] * (double)rb.box[8]);
    const float recipExpFactor = (float)(3.14159265358979323846 * 3.14159265358979323846 / ((double)alpha * (double)alpha));
    double energy = 0.0;
    for (long long index = (long long)blockIdx.x * blockDim.x + threadIdx.x; index < total; index += (long long)gridDim.x * blockDim.x) {
        const int kx = (int)(index / yzsize), rem = (int)(index - (long long)kx * yzsize);
        const int ky = rem / zsize, kz = rem - ky * zsize;
        const int mx = (kx < (gx + 1) / 2) ? kx : kx - gx;
        const int my = (ky < (gy + 1) / 2) ? ky : ky - gy;
        const int mz = (kz < (gz + 1) / 2) ? kz : kz - gz;
        const float mhx = mx * rb.recip[0];
        const float mhy = mx * rb.recip[3] + my * rb.recip[4];
        const float mhz = mx * rb.recip[6] + my * rb.recip[7] + mz * rb.recip[8];
        const float bx = scaleFactor * xmod[kx];
        const float bxby = bx * ymod[ky];
        const float m2 = (mhx * mhx + mhy * mhy) + mhz * mhz;
        const float denom = m2 * bxby * zmod[kz];
        const float eterm = (index == 0) ? 0.0f : expf(-recipExpFactor * m2) / denom;
        const float scale = (kz > 0 && kz <= (gz - 1) / 2) ? 2.0f : 1.0f;
        float2 g = recip[index];
        energy += (double)(scale * eterm * (g.x * g.x + g.y * g.y));
        g.x *= eterm; g.y *= eterm;
        recip[index] = g;
    }
    __shared__ double part[8];
    for (int o = 16; o > 0; o >>= 1) energy += __shfl_xor_sync(kFull, energy, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = energy;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += part[w];
        atomicAdd(energyAcc, t);
    }
}

template <int ORDER>
__global__ void __launch_bounds__(128)
pme_interpolate_kernel(int numAtoms, const float* __restrict__ pos, const float* __restrict__ charges, const float* __restrict__ box,
                       int gx, int gy, int gz, float sqrtCoulomb, const float* __restrict__ grid, float* __restrict__ posDeriv,
                       float* __restrict__ chargeDeriv) {
    __shared__ RecipBox rb;
    if (threadIdx.x == 0) invert_box(box, rb);
    __syncthreads();
    const int gridSize[3] = {gx, gy, gz};
    for (int atom = blockIdx.x * blockDim.x + threadIdx.x; atom < numAtoms; atom += gridDim.x * blockDim.x) {
        int gi[3];
        float data[3][ORDER], ddata[3][ORDER];
        spline<ORDER, true>(pos, atom, rb, gridSize, gi, data, ddata);
        float dpx = 0, dpy = 0, dpz = 0, dq = 0;
#pragma unroll
        for (int ix = 0; ix < ORDER; ix++) {
            const int xi = (gi[0] + ix) % gx;
            const float dx = data[0][ix], ddx = ddata[0][ix];
#pragma unroll
            for (int iy = 0; iy < ORDER; iy++) {
                const int yi = (gi[1] + iy) % gy;
                const float dy = data[1][iy], ddy = ddata[1][iy];
                const float* row = grid + ((size_t)xi * gy + yi) * gz;
#pragma unroll
                for (int iz = 0; iz < ORDER; iz++) {
                    const float dz = data[2][iz], ddz = ddata[2][iz];
                    const float g = __ldg(row + (gi[2] + iz) % gz);
                    dpx += ddx * dy * dz * g; dpy += dx * ddy * dz * g; dpz += dx * dy * ddz * g; dq += dx * dy * dz * g;
                }
            }
        }
        const float scale = charges[atom] * sqrtCoulomb;
        posDeriv[3 * (size_t)atom] = scale * (dpx * gx * rb.recip[0]);
        posDeriv[3 * (size_t)atom + 1] = scale * (dpx * gx * rb.recip[3] + dpy * gy * rb.recip[4]);
        posDeriv[3 * (size_t)atom + 2] = scale * (dpx * gx * rb.recip[6] + dpy * gy * rb.recip[7] + dpz * gz * rb.recip[8]);
        chargeDeriv[atom] = dq * sqrtCoulomb;
    }
}

#define NNP_CUFFT_CHECK(expr)                                                                                 \
    do {                                                                                                      \
        cufftResult r__ = (expr);                                                                             \
        if (r__ != CUFFT_SUCCESS) throw std::runtime_error("cuFFT error " + std::to_string((int)r__) + " at " __FILE__ ":" + std::to_string(__LINE__)); \
    } while (0)

// scratch of the reciprocal part, one per (device, grid, stream): the cuFFT plans are bound to the stream once, at creation
struct PmeWorkspace : WorkspaceBase {
    int gx, gy, gz;
    float* realGrid = nullptr;
    float2* scratch = nullptr;     // copy of the half-complex grid for the C2R transform (cuFFT may overwrite its input)
    double* energyAcc = nullptr;
    cufftHandle r2c = 0, c2r = 0;
    cudaStream_t r2cStream = nullptr, c2rStream = nullptr;   // the stream each plan is currently bound to
    ~PmeWorkspace() override {
        cudaFree(realGrid); cudaFree(scratch); cudaFree(energyAcc);
        if (r2c) cufftDestroy(r2c);
        if (c2r) cufftDestroy(c2r);
    }
};
struct DirectWorkspace : WorkspaceBase {
    double* acc = nullptr;
    ~DirectWorkspace() override { cudaFree(acc); }
};

WorkspaceCache<PmeWorkspace, std::tuple<int, int, int, int>> g_pmeWs(16);
WorkspaceCache<DirectWorkspace, int> g_directWs(64);

std::shared_ptr<PmeWorkspace> pme_workspace(int gx, int gy, int gz, cudaStream_t stream) {
    int dev = 0;
    NNP_CUDA_CHECK(cudaGetDevice(&dev));
    return g_pmeWs.get(std::make_tuple(dev, gx, gy, gz), stream, [=]() {
        std::unique_ptr<PmeWorkspace> ws(new PmeWorkspace);
        ws->gx = gx; ws->gy = gy; ws->gz = gz;
        const size_t nReal = (size_t)gx * gy * gz, nCplx = (size_t)gx * gy * (gz / 2 + 1);
        NNP_CUDA_CHECK(cudaMalloc(&ws->realGrid, sizeof(float) * nReal));
        NNP_CUDA_CHECK(cudaMalloc(&ws->scratch, sizeof(float2) * nCplx));
        NNP_CUDA_CHECK(cudaMalloc(&ws->energyAcc, sizeof(double)));
        NNP_CUFFT_CHECK(cufftPlan3d(&ws->r2c, gx, gy, gz, CUFFT_R2C));
        NNP_CUFFT_CHECK(cufftPlan3d(&ws->c2r, gx, gy, gz, CUFFT_C2R));
        NNP_CUFFT_CHECK(cufftSetStream(ws->r2c, stream));
        NNP_CUFFT_CHECK(cufftSetStream(ws->c2r, stream));
        ws->r2cStream = ws->c2rStream = stream;
        return ws.release();
    });
}

std::shared_ptr<DirectWorkspace> direct_workspace(cudaStream_t stream) {
    int dev = 0;
    NNP_CUDA_CHECK(cudaGetDevice(&dev));
    return g_directWs.get(dev, stream, []() {
        std::unique_ptr<DirectWorkspace> ws(new DirectWorkspace);
        NNP_CUDA_CHECK(cudaMalloc(&ws->acc, sizeof(double)));
        return ws.release();
    });
}

int sm_count() { return current_sm_count(); }

}  // namespace

// energy: device float[1]; posDeriv [n][3] and chargeDeriv [n] are overwritten (they are what the reference saves for backward)
void pme_direct(const float* positions, const float* charges, const int* neighbors, const float* deltas, const float* distances,
                const int* exclusions, int numAtoms, long long numPairs, int maxExcl, float alpha, float coulomb, float* energy,
                float* posDeriv, float* chargeDeriv, cudaStream_t stream) {
    const std::shared_ptr<DirectWorkspace> accHold = direct_workspace(stream);
    double* acc = accHold->acc;
    NNP_CUDA_CHECK(cudaMemsetAsync(acc, 0, sizeof(double), stream));
    NNP_CUDA_CHECK(cudaMemsetAsync(posDeriv, 0, sizeof(float) * 3 * (size_t)numAtoms, stream));
    NNP_CUDA_CHECK(cudaMemsetAsync(chargeDeriv, 0, sizeof(float) * (size_t)numAtoms, stream));
    if (numPairs > 0) {
        const int grid = (int)std::min<long long>((numPairs + 255) / 256, (long long)sm_count() * 8);
        pme_direct_pairs_kernel<<<grid, 256, 0, stream>>>(numPairs, neighbors, deltas, distances, charges, exclusions, maxExcl, alpha, coulomb,
                                                          posDeriv, chargeDeriv, acc);
        count_launch();
    }
    if (maxExcl > 0 && numAtoms > 0) {
        const long long total = (long long)numAtoms * maxExcl;
        const int grid = (int)std::min<long long>((total + 255) / 256, (long long)sm_count() * 8);
        pme_direct_exclusions_kernel<<<grid, 256, 0, stream>>>(numAtoms, positions, charges, exclusions, maxExcl, alpha, coulomb, posDeriv,
                                                               chargeDeriv, acc);
        count_launch();
    }
    publish_energy_kernel<<<1, 1, 0, stream>>>(acc, 1.0, energy);
    count_launch();
    NNP_CUDA_CHECK(cudaGetLastError());
}

// The reciprocal-space forward pass in two stages, so that a system sharded over several GPUs can sum its charge grids between
// them (one all-reduce of gx * gy * gz floats; SURVEY.md section 8e, PME row):
//   pme_spread: realGrid[gx][gy][gz] <- B-spline charge spreading of the atoms given (zeroes the grid first)   pmeCUDA.cu:30-93
//   pme_solve : R2C FFT of realGrid, convolution with the Ewald kernel, energy                                  pmeCUDA.cu:95-168
void pme_spread(const float* positions, const float* charges, const float* box, int numAtoms, int gx, int gy, int gz, int order,
                float coulomb, float* realGrid, cudaStream_t stream) {
    NNP_REQUIRE(order == 4 || order == 5, "Only pmeOrder 4 or 5 is supported with CUDA");
    NNP_REQUIRE(gx > 0 && gy > 0 && gz > 0, "The grid dimensions must be positive");
    const size_t nReal = (size_t)gx * gy * gz;
    NNP_CUDA_CHECK(cudaMemsetAsync(realGrid, 0, sizeof(float) * nReal, stream));
    const float sqrtCoulomb = (float)std::sqrt((double)coulomb);
    if (numAtoms > 0) {
        const int grid = std::min((numAtoms + 127) / 128, sm_count() * 8);
        if (order == 4) pme_spread_kernel<4><<<grid, 128, 0, stream>>>(numAtoms, positions, charges, box, gx, gy, gz, sqrtCoulomb, realGrid);
        else pme_spread_kernel<5><<<grid, 128, 0, stream>>>(numAtoms, positions, charges, box, gx, gy, gz, sqrtCoulomb, realGrid);
        count_launch();
    }
    NNP_CUDA_CHECK(cudaGetLastError());
}

// recipGrid: device float2 [gx][gy][gz/2+1], receives the convolved half-complex grid (saved by the caller for backward)
void pme_solve(float* realGrid, const float* box, int gx, int gy, int gz, float alpha, const float* xmod, const float* ymod,
               const float* zmod, float* energy, float* recipGrid, cudaStream_t stream) {
    NNP_REQUIRE(gx > 0 && gy > 0 && gz > 0, "The grid dimensions must be positive");
    const std::shared_ptr<PmeWorkspace> wsHold = pme_workspace(gx, gy, gz, stream);
    PmeWorkspace& ws = *wsHold;
    NNP_CUDA_CHECK(cudaMemsetAsync(ws.energyAcc, 0, sizeof(double), stream));
    if (ws.r2cStream != stream) {   // only when the workspace is shared with a capturing stream (workspace_cache.cuh)
        NNP_CUFFT_CHECK(cufftSetStream(ws.r2c, stream));
        ws.r2cStream = stream;
    }
    NNP_CUFFT_CHECK(cufftExecR2C(ws.r2c, realGrid, reinterpret_cast<cufftComplex*>(recipGrid)));
    const long long total = (long long)gx * gy * (gz / 2 + 1);
    const int grid = (int)std::min<long long>((total + 255) / 256, (long long)sm_count() * 8);
    pme_convolve_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<float2*>(recipGrid), box, gx, gy, gz, alpha, xmod, ymod, zmod, ws.energyAcc);
    publish_energy_kernel<<<1, 1, 0, stream>>>(ws.energyAcc, 0.5, energy);
    count_launch(2);
    NNP_CUDA_CHECK(cudaGetLastError());
}

void pme_reciprocal_forward(const float* positions, const float* charges, const float* box, int numAtoms, int gx, int gy, int gz, int order,
                            float alpha, float coulomb, const float* xmod, const float* ymod, const float* zmod, float* energy,
                            float* recipGrid, cudaStream_t stream) {
    NNP_REQUIRE(order == 4 || order == 5, "Only pmeOrder 4 or 5 is supported with CUDA");
    NNP_REQUIRE(gx > 0 && gy > 0 && gz > 0, "The grid dimensions must be positive");
    const std::shared_ptr<PmeWorkspace> wsHold = pme_workspace(gx, gy, gz, stream);
    PmeWorkspace& ws = *wsHold;
    pme_spread(positions, charges, box, numAtoms, gx, gy, gz, order, coulomb, ws.realGrid, stream);
    pme_solve(ws.realGrid, box, gx, gy, gz, alpha, xmod, ymod, zmod, energy, recipGrid, stream);
}

void pme_reciprocal_backward(const float* positions, const float* charges, const float* box, int numAtoms, int gx, int gy, int gz, int order,
                             float coulomb, const float* recipGrid, float* posDeriv, float* chargeDeriv, cudaStream_t stream) {
    NNP_REQUIRE(order == 4 || order == 5, "Only pmeOrder 4 or 5 is supported with CUDA");
    const std::shared_ptr<PmeWorkspace> wsHold = pme_workspace(gx, gy, gz, stream);
    PmeWorkspace& ws = *wsHold;
    const size_t nCplx = (size_t)gx * gy * (gz / 2 + 1);
    NNP_CUDA_CHECK(cudaMemcpyAsync(ws.scratch, recipGrid, sizeof(float2) * nCplx, cudaMemcpyDeviceToDevice, stream));
    if (ws.c2rStream != stream) {
        NNP_CUFFT_CHECK(cufftSetStream(ws.c2r, stream));
        ws.c2rStream = stream;
    }
    NNP_CUFFT_CHECK(cufftExecC2R(ws.c2r, reinterpret_cast<cufftComplex*>(ws.scratch), ws.realGrid));
    const float sqrtCoulomb = (float)std::sqrt((double)coulomb);
    if (numAtoms > 0) {
        const int grid = std::min((numAtoms + 127) / 128, sm_count() * 8);
        if (order == 4)
            pme_interpolate_kernel<4><<<grid, 128, 0, stream>>>(numAtoms, positions, charges, box, gx, gy, gz, sqrtCoulomb, ws.realGrid, posDeriv,
                                                                chargeDeriv);
        else
            pme_interpolate_kernel<5><<<grid, 128, 0, stream>>>(numAtoms, positions, charges, box, gx, gy, gz, sqrtCoulomb, ws.realGrid, posDeriv,
                                                                chargeDeriv);
        count_launch();
    }
    NNP_CUDA_CHECK(cudaGetLastError());
}

}  // namespace nnpops
