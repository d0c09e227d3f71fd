// Cell-list build: geometry -> histogram -> scan -> scatter -> deterministic in-cell ordering.  See cell_list.cuh.
#include "cell_list.cuh"
#include <algorithm>

namespace nnpops {

unsigned long long g_launches = 0;

namespace {

constexpr float kCellMargin = 1.001f;   // cells are at least this many cutoffs wide: absorbs fp32 rounding of the cell index

template <typename T>
__global__ void geom_kernel(const T* __restrict__ pos, int n, const T* __restrict__ box, float cutoff, int maxCells, Geom* out,
                            const int* __restrict__ run) {
    if (run != nullptr && *run == 0) return;   // Verlet-skin reuse step: the previous cell list stands
    __shared__ float smin[3][32], smax[3][32];
    Geom g;
    const float cw = cutoff * kCellMargin;
    if (box != nullptr) {
        for (int i = 0; i < 9; i++) g.box[i] = (float)box[i];
        g.periodic = 1;
        g.triclinic = (g.box[1] != 0 || g.box[2] != 0 || g.box[3] != 0 || g.box[5] != 0 || g.box[6] != 0 || g.box[7] != 0) ? 1 : 0;
        g.inv[0] = 1.0f / g.box[0]; g.inv[1] = 1.0f / g.box[4]; g.inv[2] = 1.0f / g.box[8];
        // perpendicular widths of the (reduced, lower-triangular) cell: V / |b x c|, V / |c x a|, V / |a x b|
        const float ax = g.box[0], bx = g.box[3], by = g.box[4], cx = g.box[6], cy = g.box[7], cz = g.box[8];
        const float vol = fabsf(ax * by * cz);
        const float bc = sqrtf((by * cz) * (by * cz) + (bx * cz) * (bx * cz) + (bx * cy - by * cx) * (bx * cy - by * cx));
        const float ca = sqrtf((cz * ax) * (cz * ax) + (cy * ax) * (cy * ax));
        const float w[3] = {vol / bc, vol / ca, fabsf(cz)};
        for (int d = 0; d < 3; d++) {
            g.nc[d] = max(1, (int)floorf(w[d] / cw));
            g.origin[d] = 0; g.cellInv[d] = 0;
        }
    } else {
        for (int i = 0; i < 9; i++) g.box[i] = 0;
        g.inv[0] = g.inv[1] = g.inv[2] = 0;
        g.periodic = 0; g.triclinic = 0;
        float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int i = threadIdx.x; i < n; i += blockDim.x)
            for (int d = 0; d < 3; d++) {
                float v = (float)pos[3 * (size_t)i + d];
                if (v == v) { lo[d] = fminf(lo[d], v); hi[d] = fmaxf(hi[d], v); }
            }
        for (int d = 0; d < 3; d++) {
            for (int o = 16; o > 0; o >>= 1) {
                lo[d] = fminf(lo[d], __shfl_xor_sync(kFull, lo[d], o));
                hi[d] = fmaxf(hi[d], __shfl_xor_sync(kFull, hi[d], o));
            }
            if ((threadIdx.x & 31) == 0) { smin[d][threadIdx.x >> 5] = lo[d]; smax[d][threadIdx.x >> 5] = hi[d]; }
        }
        __syncthreads();
        const int nw = blockDim.x >> 5;
        for (int d = 0; d < 3; d++) {
            float l = INFINITY, h = -INFINITY;
            for (int w = 0; w < nw; w++) { l = fminf(l, smin[d][w]); h = fmaxf(h, smax[d][w]); }
            if (!(h >= l)) { l = 0; h = 0; }
            float ext = h - l;
            g.origin[d] = l;
            g.nc[d] = max(1, (int)floorf(ext / cw));
            g.cellInv[d] = ext > 0 ? (float)g.nc[d] / ext : 0.0f;
        }
    }
    // keep the cell count inside the allocation (larger cells are always safe)
    double prod = (double)g.nc[0] * g.nc[1] * g.nc[2];
    if (prod > (double)maxCells) {
        double f = cbrt((double)maxCells / prod);
        for (int d = 0; d < 3; d++) {
            int old = g.nc[d];
            g.nc[d] = max(1, (int)floor(g.nc[d] * f));
            if (!g.periodic) g.cellInv[d] *= (float)g.nc[d] / (float)old;
        }
    }
    g.ncells = g.nc[0] * g.nc[1] * g.nc[2];
    g.anyOutside = 0;
    if (threadIdx.x == 0) *out = g;
}

template <typename T>
__device__ __forceinline__ int cell_of(const Geom& g, T px, T py, T pz, bool* outside = nullptr) {
    int ix, iy, iz;
    if (g.periodic) {
        // fractional coordinates in the lower-triangular box, wrapped into [0, 1)
        float fz = (float)pz * g.inv[2];
        float fy = ((float)py - fz * g.box[7]) * g.inv[1];
        float fx = ((float)px - fy * g.box[3] - fz * g.box[6]) * g.inv[0];
        if (outside) *outside = !(fx >= 0.0f && fx < 1.0f && fy >= 0.0f && fy < 1.0f && fz >= 0.0f && fz < 1.0f);
        fx -= floorf(fx); fy -= floorf(fy); fz -= floorf(fz);
        ix = (int)(fx * g.nc[0]); iy = (int)(fy * g.nc[1]); iz = (int)(fz * g.nc[2]);
    } else {
        ix = (int)(((float)px - g.origin[0]) * g.cellInv[0]);
        iy = (int)(((float)py - g.origin[1]) * g.cellInv[1]);
        iz = (int)(((float)pz - g.origin[2]) * g.cellInv[2]);
    }
    ix = min(max(ix, 0), g.nc[0] - 1); iy = min(max(iy, 0), g.nc[1] - 1); iz = min(max(iz, 0), g.nc[2] - 1);
    return (ix * g.nc[1] + iy) * g.nc[2] + iz;
}

template <typename T>
__global__ void count_kernel(const T* __restrict__ pos, int n, Geom* __restrict__ geom, int* __restrict__ cellCount,
                             int* __restrict__ cellOf, int* __restrict__ slot, const int* __restrict__ run) {
    if (run != nullptr && *run == 0) return;
    __shared__ Geom g;
    if (threadIdx.x == 0) g = *geom;
    __syncthreads();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool outside = false;
    int c = cell_of<T>(g, pos[3 * (size_t)i], pos[3 * (size_t)i + 1], pos[3 * (size_t)i + 2], &outside);
    if (outside) geom->anyOutside = 1;   // benign race: every writer stores the same value
    cellOf[i] = c;
    slot[i] = atomicAdd(&cellCount[c], 1);
}

// exclusive scan of cellCount[0..ncells) into cellStart[0..ncells] by ONE CTA, through shared memory: a tile of counts is loaded with
// coalesced accesses, every thread scans its own contiguous chunk of the tile in shared memory (odd chunk length: no bank
// conflicts), the chunk totals are scanned across the CTA, and the result goes back coalesced.  The counts are zeroed on the way
// (no per-build memset).  The ANI path uses cells of half the cutoff: 27 000 cells for the 50 000-atom box, one tile.
constexpr int kScanTile = 47 * 1024;   // ints of shared memory per tile (188 KB)

__global__ void __launch_bounds__(1024)
scan_kernel(int* __restrict__ cellCount, int* __restrict__ cellStart, const Geom* __restrict__ geom, int maxCells, const int* __restrict__ run) {
    if (run != nullptr && *run == 0) return;
    extern __shared__ int tile[];
    __shared__ int warpTot[32];
    __shared__ int carry;
    const int ncells = geom->ncells;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    for (int base = 0; base < ncells; base += kScanTile) {
        const int cnt = min(kScanTile, ncells - base);
        int chunk = (cnt + blockDim.x - 1) / blockDim.x;
        chunk |= 1;                                           // odd stride between threads: conflict-free shared-memory walks
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) { tile[i] = cellCount[base + i]; cellCount[base + i] = 0; }
        __syncthreads();
        const int b = min(threadIdx.x * chunk, cnt), e = min(b + chunk, cnt);
        int sum = 0;
        for (int i = b; i < e; i++) sum += tile[i];
        int x = sum;
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(kFull, x, o); if (lane >= o) x += y; }
        if (lane == 31) warpTot[w] = x;
        __syncthreads();
        if (w == 0) {
            int t = lane < nw ? warpTot[lane] : 0;
            for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(kFull, t, o); if (lane >= o) t += y; }
            warpTot[lane] = t;   // inclusive
        }
        __syncthreads();
        int run0 = carry + (w > 0 ? warpTot[w - 1] : 0) + x - sum;
        for (int i = b; i < e; i++) { const int v = tile[i]; tile[i] = run0; run0 += v; }
        __syncthreads();
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) cellStart[base + i] = tile[i];
        if (threadIdx.x == 0) carry += warpTot[nw - 1];
        __syncthreads();
    }
    if (threadIdx.x == 0) cellStart[ncells] = carry;
}

__global__ void scatter_kernel(int n, const int* __restrict__ cellOf, const int* __restrict__ slot, const int* __restrict__ cellStart,
                               int* __restrict__ tmpIdx, const int* __restrict__ run) {
    if (run != nullptr && *run == 0) return;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) tmpIdx[cellStart[cellOf[i]] + slot[i]] = i;
}

template <typename T>
__global__ void order_kernel(const T* __restrict__ pos, const int* __restrict__ tags, int n, const int* __restrict__ cellOf,
                             const int* __restrict__ cellStart, const int* __restrict__ tmpIdx, float4* __restrict__ sorted,
                             int* __restrict__ sortedOrig, int* __restrict__ sortedCell, const int* __restrict__ run) {
    if (run != nullptr && *run == 0) return;
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int i = tmpIdx[p];
    const int c = cellOf[i];
    const int b = cellStart[c], e = cellStart[c + 1];
    int rank = 0;
    for (int q = b; q < e; q++) rank += (tmpIdx[q] < i) ? 1 : 0;
    const int dst = b + rank;
    float4 v;
    v.x = (float)pos[3 * (size_t)i]; v.y = (float)pos[3 * (size_t)i + 1]; v.z = (float)pos[3 * (size_t)i + 2];
    v.w = __int_as_float(tags ? tags[i] : 0);
    sorted[dst] = v;
    sortedOrig[dst] = i;
    sortedCell[dst] = c;
}

}  // namespace

void CellList::init(int numAtoms) {
    release();
    n = numAtoms;
    long long want = 4LL * (long long)numAtoms + 64;
    maxCells = (int)(want > (1LL << 24) ? (1LL << 24) : want);
    size_t na = (size_t)(n > 0 ? n : 1);
    NNP_CUDA_CHECK(cudaMalloc(&geom, sizeof(Geom)));
    NNP_CUDA_CHECK(cudaMalloc(&cellCount, sizeof(int) * (maxCells + 1)));
    NNP_CUDA_CHECK(cudaMemset(cellCount, 0, sizeof(int) * (maxCells + 1)));   // kept zero between builds by scan_kernel
    NNP_CUDA_CHECK(cudaMalloc(&cellStart, sizeof(int) * (maxCells + 1)));
    NNP_CUDA_CHECK(cudaMalloc(&cellOf, sizeof(int) * na));
    NNP_CUDA_CHECK(cudaMalloc(&slot, sizeof(int) * na));
    NNP_CUDA_CHECK(cudaMalloc(&tmpIdx, sizeof(int) * na));
    NNP_CUDA_CHECK(cudaMalloc(&sorted, sizeof(float4) * na));
    NNP_CUDA_CHECK(cudaMalloc(&sortedOrig, sizeof(int) * na));
    NNP_CUDA_CHECK(cudaMalloc(&sortedCell, sizeof(int) * na));
}

void CellList::release() {
    cudaFree(geom); cudaFree(cellCount); cudaFree(cellStart); cudaFree(cellOf); cudaFree(slot); cudaFree(tmpIdx);
    cudaFree(sorted); cudaFree(sortedOrig); cudaFree(sortedCell);
    geom = nullptr; cellCount = cellStart = cellOf = slot = tmpIdx = sortedOrig = sortedCell = nullptr; sorted = nullptr;
    n = 0; maxCells = 0;
}

template <typename T>
void CellList::build(const T* positions, const T* box, const int* tags, float cutoff, cudaStream_t stream, const int* run) {
    if (n == 0) return;
    const int tb = 256, nb = (n + tb - 1) / tb;
    geom_kernel<T><<<1, 1024, 0, stream>>>(positions, n, box, cutoff, maxCells, geom, run);
    count_kernel<T><<<nb, tb, 0, stream>>>(positions, n, geom, cellCount, cellOf, slot, run);
    static bool attr[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr[dev]) {
        NNP_CUDA_CHECK(cudaFuncSetAttribute(scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kScanTile * (int)sizeof(int)));
        attr[dev] = true;
    }
    scan_kernel<<<1, 1024, kScanTile * sizeof(int), stream>>>(cellCount, cellStart, geom, maxCells, run);
    scatter_kernel<<<nb, tb, 0, stream>>>(n, cellOf, slot, cellStart, tmpIdx, run);
    order_kernel<T><<<nb, tb, 0, stream>>>(positions, tags, n, cellOf, cellStart, tmpIdx, sorted, sortedOrig, sortedCell, run);
    count_launch(5);
    NNP_CUDA_CHECK(cudaGetLastError());
}

template void CellList::build<float>(const float*, const float*, const int*, float, cudaStream_t, const int*);
template void CellList::build<double>(const double*, const double*, const int*, float, cudaStream_t, const int*);

}  // namespace nnpops
