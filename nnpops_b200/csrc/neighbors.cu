// neighbors::getNeighborPairs on the cell list.  Replaces the reference's one-thread-per-candidate-pair kernels
// (src/pytorch/neighbors/getNeighborPairsCUDA.cu:31-101, K9/K10 of SURVEY.md section 2.3), which launch N(N-1)/2 threads and
// index them in int32.
//
// Output conventions reproduced exactly (getNeighborPairs.py:60-96, getNeighborPairsCPU.cpp:56-98):
//   * pair (row, col) with row > col, deltas = pos[row] - pos[col], minimum image by DIVISION round(delta_z / box[2][2]) ... in the
//     order z, y, x (getNeighborPairsCPU.cpp:65-69), inclusive cutoff  distance <= cutoff (:73,81);
//   * max_num_pairs == -1: output length N(N-1)/2, slot = row(row-1)/2 + col, misses are -1 / NaN;
//   * max_num_pairs  >  0: compacted list padded with -1 / NaN, pairs beyond the capacity dropped.
// Unlike the reference CUDA path the compacted list is deterministic: pass 1 counts the pairs each atom owns (the atom with the
// smaller cell-sorted index owns the pair), an exclusive scan gives its offset, pass 2 writes them (ballot compaction inside the
// warp), so the order depends only on the input.
// num_found always holds the number of pairs inside the cutoff (the reference CUDA path leaves it 0 in all-pairs mode).
#include <map>
#include <mutex>
#include "cell_list.cuh"
#include "pair_delta.cuh"
#include "workspace_cache.cuh"

namespace nnpops {

namespace {

constexpr int kWPB = 8;

// MODE 0: count pairs owned by each sorted atom; MODE 1: write the compacted list at offsets[p]; MODE 2: all-pairs slots
template <typename T, int MODE>
__global__ void __launch_bounds__(kWPB * 32)
pairs_kernel(int n, const T* __restrict__ pos, const T* __restrict__ boxPtr, const float4* __restrict__ sorted,
             const int* __restrict__ sortedOrig, const int* __restrict__ sortedCell, const Geom* __restrict__ geom,
             const int* __restrict__ cellStart, T cutoff, int* __restrict__ counts, const long long* __restrict__ offsets,
             long long capacity, int* __restrict__ neighbors, T* __restrict__ deltas, T* __restrict__ distances,
             unsigned long long* __restrict__ found) {
    __shared__ Geom g;
    __shared__ Box<T> bx;
    if (threadIdx.x == 0) {
        g = *geom;
        bx.periodic = boxPtr != nullptr;
        for (int i = 0; i < 9; i++) bx.b[i] = boxPtr ? boxPtr[i] : (T)0;
    }
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = blockIdx.x * kWPB + w;
    if (p >= n) return;
    // Half shell in sorted order: the unordered pair {p, q} is handled by the atom with the smaller sorted index, so every pair is
    // tested once.  A cheap fp32 pre-test on the sorted coordinates (multiply-by-reciprocal minimum image, 1 % margin) discards
    // the ~85 % of candidates that are clearly outside the cutoff before the exact reference arithmetic (divisions) runs.
    const int op = sortedOrig[p];
    const float4 cp = sorted[p];
    const T pp[3] = {pos[3 * (size_t)op], pos[3 * (size_t)op + 1], pos[3 * (size_t)op + 2]};
    const float pre2 = (float)cutoff * (float)cutoff * 1.0201f;
    long long cursor = (MODE == 1) ? offsets[p] : 0;
    int mine = 0;
    // Survivors of the pre-test are queued (ballot compaction, scan order preserved) and the exact test runs on full batches of
    // 32, so the divisions of the reference arithmetic execute on dense lanes instead of the ~15 % that survive each sweep.
    __shared__ int queueAll[kWPB][64];
    // fp32 only: the survivor's minimum-image displacement (candidate - centre) travels with it.  The sorted coordinates ARE the fp32
    // positions, and the multiply-by-reciprocal image of the pre-test equals the reference's division image step for step unless a
    // quotient sits on a rounding tie (|component| of half a box edge): only those survivors are re-derived with the division form.
    // A negated displacement is the displacement of the swapped pair exactly (rint is odd), so orientation row > col costs a sign.
    constexpr bool kCarry = sizeof(T) == 4;
    __shared__ float4 deltaAll[kCarry ? kWPB : 1][64];
    int* queue = queueAll[w];
    float4* qdelta = deltaAll[kCarry ? w : 0];
    const float tieX = 0.499f * (float)bx.b[0], tieY = 0.499f * (float)bx.b[4], tieZ = 0.499f * (float)bx.b[8];
    int queued = 0;
    auto drain = [&](int count) {   // exact test + output for queue[0 .. count)
        bool ok = false;
        int row = -1, col = -1;
        T dx = 0, dy = 0, dz = 0, d = 0;
        if (lane < count) {
            const int oq = sortedOrig[queue[lane]];
            bool exact = !kCarry;
            if (kCarry) {
                const float4 e = qdelta[lane];
                exact = bx.periodic && (fabsf(e.x) >= tieX || fabsf(e.y) >= tieY || fabsf(e.z) >= tieZ);
                if (!exact) {
                    const float sgn = oq > op ? 1.0f : -1.0f;   // delta = pos[row] - pos[col], row the larger original index
                    dx = (T)(sgn * e.x); dy = (T)(sgn * e.y); dz = (T)(sgn * e.z);
                    d = (T)__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(e.x, e.x), __fmul_rn(e.y, e.y)), __fmul_rn(e.z, e.z)));
                    row = max(op, oq); col = min(op, oq);
                }
            }
            if (exact) {
                const T pq[3] = {pos[3 * (size_t)oq], pos[3 * (size_t)oq + 1], pos[3 * (size_t)oq + 2]};
                if (oq > op) { row = oq; col = op; d = pair_delta<T>(bx, pq, pp, dx, dy, dz); }
                else         { row = op; col = oq; d = pair_delta<T>(bx, pp, pq, dx, dy, dz); }
            }
            ok = d <= cutoff;
        }
        const unsigned m = __ballot_sync(kFull, ok);
        if (MODE != 0 && ok) {
            long long slot;
            if (MODE == 1) slot = cursor + __popc(m & ((1u << lane) - 1u));
            else slot = (long long)row * (row - 1) / 2 + col;
            if (slot < capacity) {
                neighbors[slot] = row;
                neighbors[capacity + slot] = col;
                deltas[3 * slot] = dx; deltas[3 * slot + 1] = dy; deltas[3 * slot + 2] = dz;
                distances[slot] = d;
            }
        }
        cursor += __popc(m);
        mine += __popc(m);
    };
    // runs that do not cross a periodic face need no minimum-image step in the pre-test (it would subtract zero; cell_list.cuh)
    const bool alwaysImage = g.periodic && (g.triclinic || g.anyOutside);
    for_each_candidate_run_w(g, cellStart, sortedCell[p], [&](int b, int e, bool wrapped) {
        const bool image = alwaysImage || wrapped;
        for (int q0 = max(b, p + 1); q0 < e; q0 += 32) {
            const int q = q0 + lane;
            bool keep = false;
            float ax = 0.0f, ay = 0.0f, az = 0.0f;
            if (q < e) {
                const float4 cq = sorted[q];
                ax = __fsub_rn(cq.x, cp.x); ay = __fsub_rn(cq.y, cp.y); az = __fsub_rn(cq.z, cp.z);
                keep = (image ? min_image_mul(g, ax, ay, az) : ax * ax + ay * ay + az * az) <= pre2;
            }
            const unsigned m = __ballot_sync(kFull, keep);
            if (keep) {
                const int slot = queued + __popc(m & ((1u << lane) - 1u));
                queue[slot] = q;
                if (kCarry) qdelta[slot] = make_float4(ax, ay, az, 0.0f);
            }
            queued += __popc(m);
            __syncwarp();
            if (queued >= 32) {
                drain(32);
                __syncwarp();
                const int carry = (lane < queued - 32) ? queue[32 + lane] : 0;
                float4 carryD = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                if (kCarry && lane < queued - 32) carryD = qdelta[32 + lane];
                __syncwarp();
                if (lane < queued - 32) { queue[lane] = carry; if (kCarry) qdelta[lane] = carryD; }
                queued -= 32;
                __syncwarp();
            }
        }
    });
    if (queued > 0) drain(queued);
    if (lane == 0) {
        if (MODE == 0) counts[p] = mine;
        if (MODE == 2 && mine) atomicAdd(found, (unsigned long long)mine);
    }
}

// exclusive scan of int counts into 64-bit offsets, single CTA, 8 items per thread per sweep; total -> *found
__global__ void scan_counts_kernel(const int* __restrict__ counts, int n, long long* __restrict__ offsets,
                                   unsigned long long* __restrict__ found, int* __restrict__ numFoundOut) {
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
        for (int k = 0; k < IPT; k++) { if (i0 + k < n) offsets[i0 + k] = run; run += v[k]; }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = run;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        *found = (unsigned long long)carry;
        *numFoundOut = (int)(carry > 0x7fffffffLL ? 0x7fffffffLL : carry);
    }
}

// fill slots [start, capacity) with -1 / NaN; start = 0 (all-pairs mode, before the pair kernel) or *found (compact mode, after)
template <typename T>
__global__ void pad_kernel(long long capacity, const unsigned long long* __restrict__ found, bool fromFound, int* __restrict__ neighbors,
                           T* __restrict__ deltas, T* __restrict__ distances) {
    const long long start = fromFound ? (long long)*found : 0;
    const T nanv = Arith<T>::nan();
    for (long long i = start + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < capacity; i += (long long)gridDim.x * blockDim.x) {
        neighbors[i] = -1; neighbors[capacity + i] = -1;
        deltas[3 * i] = nanv; deltas[3 * i + 1] = nanv; deltas[3 * i + 2] = nanv;
        distances[i] = nanv;
    }
}

__global__ void publish_found_kernel(const unsigned long long* __restrict__ found, int* __restrict__ out) {
    const unsigned long long f = *found;
    *out = (int)(f > 0x7fffffffULL ? 0x7fffffffULL : f);
}

// backward (getNeighborPairsCUDA.cu:80-101): grad_pos[row] += g, grad_pos[col] -= g, g = grad_delta + delta / dist * grad_dist
template <typename T>
__global__ void pairs_backward_kernel(long long numPairs, const int* __restrict__ neighbors, const T* __restrict__ deltas,
                                      const T* __restrict__ distances, const T* __restrict__ gradDeltas, const T* __restrict__ gradDistances,
                                      T* __restrict__ gradPos) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= numPairs) return;
    const int row = neighbors[i], col = neighbors[numPairs + i];
    if (row < 0) return;
    const T d = distances[i], gd = gradDistances[i];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const T gk = gradDeltas[3 * i + k] + deltas[3 * i + k] / d * gd;
        atomicAdd(&gradPos[3 * (size_t)row + k], gk);
        atomicAdd(&gradPos[3 * (size_t)col + k], -gk);
    }
}

struct Workspace : WorkspaceBase {
    CellList cells;
    int* counts = nullptr;
    long long* offsets = nullptr;
    unsigned long long* found = nullptr;
    ~Workspace() override {
        cells.release();
        cudaFree(counts); cudaFree(offsets); cudaFree(found);
    }
};

// (device, numAtoms, stream) -> workspace (workspace_cache.cuh: stream-keyed, graph-pinned, LRU)
WorkspaceCache<Workspace, std::pair<int, int>> g_ws(32);

std::shared_ptr<Workspace> workspace(int n, cudaStream_t stream) {
    int dev = 0;
    NNP_CUDA_CHECK(cudaGetDevice(&dev));
    return g_ws.get(std::make_pair(dev, n), stream, [n]() {
        Workspace* ws = new Workspace;
        ws->cells.init(n);
        NNP_CUDA_CHECK(cudaMalloc(&ws->counts, sizeof(int) * (size_t)(n > 0 ? n : 1)));
        NNP_CUDA_CHECK(cudaMalloc(&ws->offsets, sizeof(long long) * (size_t)(n > 0 ? n : 1)));
        NNP_CUDA_CHECK(cudaMalloc(&ws->found, sizeof(unsigned long long)));
        return ws;
    });
}

}  // namespace

// neighbors: int [2][P]; deltas: T [P][3]; distances: T [P]; numFound: int [1]; P = maxNumPairs, or N(N-1)/2 when maxNumPairs == -1
template <typename T>
void neighbor_pairs(const T* positions, const T* box, int n, T cutoff, long long maxNumPairs, int* neighbors, T* deltas, T* distances,
                    int* numFound, cudaStream_t stream) {
    NNP_REQUIRE(n > 0, "Expected the 1nd dimension size of \"positions\" to be more than 0");
    NNP_REQUIRE(cutoff > 0, "Expected \"cutoff\" to be positive");
    NNP_REQUIRE(maxNumPairs > 0 || maxNumPairs == -1, "Expected \"max_num_pairs\" to be positive or equal to -1");
    const std::shared_ptr<Workspace> wsHold = workspace(n, stream);
    Workspace& ws = *wsHold;
    ws.cells.build<T>(positions, box, nullptr, (float)cutoff * 1.0001f + 1e-30f, stream);
    const int grid = (n + kWPB - 1) / kWPB;
    const bool allPairs = maxNumPairs == -1;
    const long long capacity = allPairs ? (long long)n * (n - 1) / 2 : maxNumPairs;
    const int padGrid = (int)std::min<long long>((capacity + 255) / 256, (long long)current_sm_count() * 16);
    if (allPairs) {
        NNP_CUDA_CHECK(cudaMemsetAsync(ws.found, 0, sizeof(unsigned long long), stream));
        if (capacity > 0) pad_kernel<T><<<padGrid, 256, 0, stream>>>(capacity, ws.found, false, neighbors, deltas, distances);
        pairs_kernel<T, 2><<<grid, kWPB * 32, 0, stream>>>(n, positions, box, ws.cells.sorted, ws.cells.sortedOrig, ws.cells.sortedCell,
                                                            ws.cells.geom, ws.cells.cellStart, cutoff, ws.counts, ws.offsets, capacity,
                                                            neighbors, deltas, distances, ws.found);
        publish_found_kernel<<<1, 1, 0, stream>>>(ws.found, numFound);
        count_launch(3);
    } else {
        pairs_kernel<T, 0><<<grid, kWPB * 32, 0, stream>>>(n, positions, box, ws.cells.sorted, ws.cells.sortedOrig, ws.cells.sortedCell,
                                                            ws.cells.geom, ws.cells.cellStart, cutoff, ws.counts, ws.offsets, capacity,
                                                            neighbors, deltas, distances, ws.found);
        scan_counts_kernel<<<1, 1024, 0, stream>>>(ws.counts, n, ws.offsets, ws.found, numFound);
        pairs_kernel<T, 1><<<grid, kWPB * 32, 0, stream>>>(n, positions, box, ws.cells.sorted, ws.cells.sortedOrig, ws.cells.sortedCell,
                                                            ws.cells.geom, ws.cells.cellStart, cutoff, ws.counts, ws.offsets, capacity,
                                                            neighbors, deltas, distances, ws.found);
        pad_kernel<T><<<padGrid, 256, 0, stream>>>(capacity, ws.found, true, neighbors, deltas, distances);
        count_launch(4);
    }
    NNP_CUDA_CHECK(cudaGetLastError());
}

template <typename T>
void neighbor_pairs_backward(const int* neighbors, const T* deltas, const T* distances, const T* gradDeltas, const T* gradDistances,
                             long long numPairs, int n, T* gradPositions, cudaStream_t stream) {
    NNP_CUDA_CHECK(cudaMemsetAsync(gradPositions, 0, sizeof(T) * 3 * (size_t)n, stream));
    if (numPairs <= 0) return;
    pairs_backward_kernel<T><<<(unsigned)((numPairs + 255) / 256), 256, 0, stream>>>(numPairs, neighbors, deltas, distances, gradDeltas,
                                                                                     gradDistances, gradPositions);
    count_launch();
    NNP_CUDA_CHECK(cudaGetLastError());
}

template void neighbor_pairs<float>(const float*, const float*, int, float, long long, int*, float*, float*, int*, cudaStream_t);
template void neighbor_pairs<double>(const double*, const double*, int, double, long long, int*, double*, double*, int*, cudaStream_t);
template void neighbor_pairs_backward<float>(const int*, const float*, const float*, const float*, const float*, long long, int, float*,
                                             cudaStream_t);
template void neighbor_pairs_backward<double>(const int*, const double*, const double*, const double*, const double*, long long, int,
                                              double*, cudaStream_t);

}  // namespace nnpops
