// PME direct space fused into the cell-list traversal: no pair list, no atomics.
//
// The reference evaluates the direct-space sum from a materialised pair list (pme.py:131-165: getNeighborPairs, then the kernel
// computeDirect, pmeCUDA.cu:30-95, one thread per pair with eight float atomics per pair).  At BASELINE config 5 (200 000 charges,
// 0.9 nm) that list is 29.5 M pairs = 730 MB written and read back, and the atomics bound the kernel.  Here ONE kernel does both
// steps, centre-owned: a warp takes a centre atom, walks the 27 cells around it, accepts a neighbour with exactly the arithmetic of
// getNeighborPairs (pair_delta.cuh: delta = pos[row] - pos[col] with row > col, sequential minimum image by division, distance <=
// cutoff), applies the exclusion test of computeDirect (the row atom's sorted exclusion list is scanned for the column atom) and
// accumulates the centre's own dE/dx, dE/dq in registers; every pair is visited from both ends, its energy is counted at the row end.
// The exclusion correction (pmeCPU.cpp:133-157: minus the erf part, NON-periodic displacement) is centre-owned too.  Outputs are
// plain stores -- nothing is zeroed or reduced across warps -- so a range of centres is a complete, independent piece of work:
// `shardIndex / shardCount` select a contiguous range of the CELL-SORTED atoms (cells are ordered x-major: a slab of the box) for
// one GPU of several (pme.py: PME.compute_direct_sharded); atoms outside the range get zeros.
#include <memory>
#include "cell_list.cuh"
#include "pair_delta.cuh"
#include "workspace_cache.cuh"

namespace nnpops {

namespace {

constexpr int kWPB = 8;
constexpr float kTwoOverSqrtPi = 1.1283791670955126f;   // M_2_SQRTPI

// the terms of one pair as computeDirect forms them (c1 = charge of the row atom, c2 = of the column atom)
struct PairTerms {
    float energy, dq1, dq2, dEdR;
};
__device__ __forceinline__ PairTerms erfc_terms(float r, float c1, float c2, float alpha, float coulomb) {
    const float invR = 1.0f / r;
    const float alphaR = alpha * r;
    const float expTerm = expf(-alphaR * alphaR);
    const float erfcTerm = erfcf(alphaR);
    const float pref = coulomb * invR;
    PairTerms t;
    t.energy = pref * erfcTerm * c1 * c2;
    t.dq1 = pref * erfcTerm * c2;
    t.dq2 = pref * erfcTerm * c1;
    t.dEdR = pref * c1 * c2 * (erfcTerm + alphaR * expTerm * kTwoOverSqrtPi) * invR * invR;
    return t;
}

// is `target` in the descending row `row` of the exclusion table?
__device__ __forceinline__ bool excluded(const int* __restrict__ exclusions, int maxExcl, int row, int target) {
    for (int j = 0; j < maxExcl; j++) {
        const int x = exclusions[(size_t)row * maxExcl + j];
        if (x < target) return false;
        if (x == target) return true;
    }
    return false;
}

#ifndef PME_MIN_BLOCKS
#define PME_MIN_BLOCKS 3   // 75 registers, no spills (ptxas settles on 64 registers with spills when left alone)
#endif
__global__ void __launch_bounds__(kWPB * 32, PME_MIN_BLOCKS)
pme_direct_fused_kernel(int n, int pLo, int pHi, const float* __restrict__ boxPtr, const float4* __restrict__ sorted,
                        const int* __restrict__ sortedOrig, const int* __restrict__ sortedCell, const Geom* __restrict__ geom,
                        const int* __restrict__ cellStart, const float* __restrict__ pos, const float* __restrict__ charges,
                        const int* __restrict__ exclusions, int maxExcl, float cutoff, float alpha, float coulomb,
                        float* __restrict__ posDeriv, float* __restrict__ chargeDeriv, double* __restrict__ energyAcc) {
    __shared__ Geom g;
    __shared__ Box<float> bx;
    __shared__ float4 queueAll[kWPB][96];   // up to 31 waiting survivors + 64 new ones
    __shared__ int queueIdxAll[kWPB][96];
    __shared__ double part[kWPB];
    if (threadIdx.x == 0) {
        g = *geom;
        bx.periodic = boxPtr != nullptr;
        for (int i = 0; i < 9; i++) bx.b[i] = boxPtr ? boxPtr[i] : 0.0f;
    }
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = pLo + blockIdx.x * kWPB + w;
    double energy = 0.0;
    if (p < pHi) {
        const int op = sortedOrig[p];
        const float4 cp = sorted[p];
        const float qp = cp.w;   // the centre's own charge (tag word)
        const float pre2 = cutoff * cutoff * 1.0201f;
        // |component| of a minimum-image displacement at which the multiply-by-reciprocal image used below could pick another image
        // than the reference's division (a rounding tie at half a box edge): such a survivor is re-derived with the reference form
        const float tieX = 0.499f * bx.b[0], tieY = 0.499f * bx.b[4], tieZ = 0.499f * bx.b[8];
        float fx = 0.0f, fy = 0.0f, fz = 0.0f, dq = 0.0f;
        float4* queue = queueAll[w];      // {delta = candidate - centre (minimum image), charge}
        int* queueIdx = queueIdxAll[w];   // sorted position of the candidate
        int queued = 0;
        auto drain = [&](int count) {   // exact test + the pair's terms for queue[0 .. count)
            if (lane < count) {
                float4 e = queue[lane];
                const int q = queueIdx[lane];
                if (bx.periodic && (fabsf(e.x) >= tieX || fabsf(e.y) >= tieY || fabsf(e.z) >= tieZ)) {
                    const float4 cq = sorted[q];
                    const float pq[3] = {cq.x, cq.y, cq.z}, pp[3] = {cp.x, cp.y, cp.z};
                    pair_delta<float>(bx, pq, pp, e.x, e.y, e.z);
                }
                // the reference's distance: sqrt of the sum of squares, every step rounded (getNeighborPairsCPU.cpp:71-73)
                const float d = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(e.x, e.x), __fmul_rn(e.y, e.y)), __fmul_rn(e.z, e.z)));
                if (d <= cutoff) {
                    bool include = true;
                    if (maxExcl > 0) {   // computeDirect scans the row atom's list (the larger index) for the column atom
                        const int oq = sortedOrig[q];
                        include = !excluded(exclusions, maxExcl, max(op, oq), min(op, oq));
                    }
                    if (include) {
                        const PairTerms t = erfc_terms(d, qp, e.w, alpha, coulomb);
                        // posDeriv[row] -= dEdR (pos[row] - pos[col]), posDeriv[col] += the same: for the centre, either way, dEdR * (other - centre)
                        fx += t.dEdR * e.x; fy += t.dEdR * e.y; fz += t.dEdR * e.z;
                        dq += t.dq1;
                        if (q < p) energy += (double)t.energy;   // every pair is seen from both ends: counted once
                    }
                }
            }
        };
        // runs that do not cross a periodic face need no minimum-image step (it would subtract zero; cell_list.cuh)
        const bool alwaysImage = g.periodic && (g.triclinic || g.anyOutside);
        auto flush = [&]() {   // a full batch of survivors: evaluate it and move the rest of the queue down
            drain(32);
            __syncwarp();
            float4 carry0 = make_float4(0.0f, 0.0f, 0.0f, 0.0f), carry1 = carry0;   // up to 63 entries stay behind
            int carryIdx0 = 0, carryIdx1 = 0;
            if (lane < queued - 32) { carry0 = queue[32 + lane]; carryIdx0 = queueIdx[32 + lane]; }
            if (lane < queued - 64) { carry1 = queue[64 + lane]; carryIdx1 = queueIdx[64 + lane]; }
            __syncwarp();
            if (lane < queued - 32) { queue[lane] = carry0; queueIdx[lane] = carryIdx0; }
            if (lane < queued - 64) { queue[32 + lane] = carry1; queueIdx[32 + lane] = carryIdx1; }
            queued -= 32;
            __syncwarp();
        };
        for_each_candidate_run_w(g, cellStart, sortedCell[p], [&](int b, int e, bool wrapped) {
            const bool image = alwaysImage || wrapped;
            // two sweeps of 32 candidates per iteration: two independent loads and tests in flight per lane
            for (int q0 = b; q0 < e; q0 += 64) {
                const int qa = q0 + lane, qb = qa + 32;
                const float4 ca = qa < e ? sorted[qa] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                const float4 cb = qb < e ? sorted[qb] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                float ax = ca.x - cp.x, ay = ca.y - cp.y, az = ca.z - cp.z;
                float bx2 = cb.x - cp.x, by2 = cb.y - cp.y, bz2 = cb.z - cp.z;
                bool keepA = false, keepB = false;
                if (image) {
                    const float ra = min_image_mul(g, ax, ay, az), rb = min_image_mul(g, bx2, by2, bz2);
                    keepA = qa < e && qa != p && ra <= pre2;
                    keepB = qb < e && qb != p && rb <= pre2;
                } else {
                    keepA = qa < e && qa != p && ax * ax + ay * ay + az * az <= pre2;
                    keepB = qb < e && qb != p && bx2 * bx2 + by2 * by2 + bz2 * bz2 <= pre2;
                }
                const unsigned ma = __ballot_sync(kFull, keepA), mb = __ballot_sync(kFull, keepB);
                const unsigned below = (1u << lane) - 1u;
                if (keepA) {
                    const int slot = queued + __popc(ma & below);
                    queue[slot] = make_float4(ax, ay, az, ca.w);
                    queueIdx[slot] = qa;
                }
                if (keepB) {
                    const int slot = queued + __popc(ma) + __popc(mb & below);
                    queue[slot] = make_float4(bx2, by2, bz2, cb.w);
                    queueIdx[slot] = qb;
                }
                queued += __popc(ma) + __popc(mb);
                __syncwarp();
                while (queued >= 32) flush();
            }
        });
        if (queued > 0) drain(queued);
        // exclusion correction, centre-owned: the reference handles the pair (a1, a2), a2 > a1, when a2 is in a1's list
        for (int j = lane; j < maxExcl; j += 32) {
            const int x = exclusions[(size_t)op * maxExcl + j];
            if (x < 0 || x == op) continue;
            const int a1 = min(op, x), a2 = max(op, x);
            if (x < op && !excluded(exclusions, maxExcl, a1, a2)) continue;   // the pair is only corrected when a1 lists a2
            float dr[3];
            for (int k = 0; k < 3; k++) dr[k] = pos[3 * (size_t)a1 + k] - pos[3 * (size_t)a2 + k];
            const float rr = sqrtf(dr[0] * dr[0] + dr[1] * dr[1] + dr[2] * dr[2]);
            const float invR = 1.0f / rr;
            const float alphaR = alpha * rr;
            const float expTerm = expf(-alphaR * alphaR);
            const float erfTerm = erff(alphaR);
            const float pref = coulomb * invR;
            const float c1 = charges[a1], c2 = charges[a2];
            const float dEdR = pref * c1 * c2 * (erfTerm - alphaR * expTerm * kTwoOverSqrtPi) * invR * invR;
            const float s = op == a1 ? dEdR : -dEdR;   // posDeriv[a1] += dEdR * dr, posDeriv[a2] -= dEdR * dr
            fx += s * dr[0]; fy += s * dr[1]; fz += s * dr[2];
            dq -= pref * erfTerm * (op == a1 ? c2 : c1);
            if (op == a1) energy -= (double)(pref * erfTerm * c1 * c2);
        }
        fx = warp_sum(fx); fy = warp_sum(fy); fz = warp_sum(fz); dq = warp_sum(dq);
        if (lane == 0) {
            posDeriv[3 * (size_t)op] = fx; posDeriv[3 * (size_t)op + 1] = fy; posDeriv[3 * (size_t)op + 2] = fz;
            chargeDeriv[op] = dq;
        }
    }
    for (int o = 16; o > 0; o >>= 1) energy += __shfl_xor_sync(kFull, energy, o);
    if (lane == 0) part[w] = energy;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int i = 0; i < kWPB; i++) t += part[i];
        if (t != 0.0) atomicAdd(energyAcc, t);
    }
}

// atoms outside the shard's range of sorted positions receive zeros
__global__ void pme_zero_outside_kernel(int n, int pLo, int pHi, const int* __restrict__ sortedOrig, float* __restrict__ posDeriv,
                                        float* __restrict__ chargeDeriv) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n || (p >= pLo && p < pHi)) return;
    const int o = sortedOrig[p];
    posDeriv[3 * (size_t)o] = 0.0f; posDeriv[3 * (size_t)o + 1] = 0.0f; posDeriv[3 * (size_t)o + 2] = 0.0f;
    chargeDeriv[o] = 0.0f;
}

__global__ void pme_publish_kernel(double* __restrict__ acc, float* __restrict__ out) { *out = (float)*acc; *acc = 0.0; }

struct FusedWorkspace : WorkspaceBase {
    CellList cells;
    double* acc = nullptr;   // zero between calls (the publish kernel resets it)
    ~FusedWorkspace() override {
        cells.release();
        cudaFree(acc);
    }
};

WorkspaceCache<FusedWorkspace, std::pair<int, int>> g_fusedWs(16);

}  // namespace

// energy: device float [1]; posDeriv [n][3], chargeDeriv [n]: overwritten.  box: device float [3][3] or nullptr (non-periodic).
void pme_direct_fused(const float* positions, const float* charges, const float* box, const int* exclusions, int numAtoms, int maxExcl,
                      float cutoff, float alpha, float coulomb, int shardIndex, int shardCount, float* energy, float* posDeriv,
                      float* chargeDeriv, cudaStream_t stream) {
    NNP_REQUIRE(numAtoms > 0, "Expected the 1nd dimension size of \"positions\" to be more than 0");
    NNP_REQUIRE(cutoff > 0, "Expected \"cutoff\" to be positive");
    NNP_REQUIRE(shardCount >= 1 && shardIndex >= 0 && shardIndex < shardCount, "pme_direct_fused: invalid shard");
    int dev = 0;
    NNP_CUDA_CHECK(cudaGetDevice(&dev));
    const int n = numAtoms;
    const std::shared_ptr<FusedWorkspace> hold = g_fusedWs.get(std::make_pair(dev, n), stream, [n]() {
        std::unique_ptr<FusedWorkspace> ws(new FusedWorkspace);
        ws->cells.init(n);
        NNP_CUDA_CHECK(cudaMalloc(&ws->acc, sizeof(double)));
        NNP_CUDA_CHECK(cudaMemset(ws->acc, 0, sizeof(double)));
        return ws.release();
    });
    FusedWorkspace& ws = *hold;
    NNP_CUDA_CHECK(cudaMemsetAsync(ws.acc, 0, sizeof(double), stream));   // (also reset by the publish kernel; a failed call must not leak into the next)
    // the charge rides in the tag word of the sorted coordinates: one 16-byte load brings a candidate's position and charge
    ws.cells.build<float>(positions, box, reinterpret_cast<const int*>(charges), cutoff * 1.0001f + 1e-30f, stream);
    const int per = (n + shardCount - 1) / shardCount;
    const int pLo = std::min(n, shardIndex * per), pHi = std::min(n, pLo + per);
    if (shardCount > 1) {
        pme_zero_outside_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, pLo, pHi, ws.cells.sortedOrig, posDeriv, chargeDeriv);
        count_launch();
    }
    if (pHi > pLo) {
        pme_direct_fused_kernel<<<(pHi - pLo + kWPB - 1) / kWPB, kWPB * 32, 0, stream>>>(
            n, pLo, pHi, box, ws.cells.sorted, ws.cells.sortedOrig, ws.cells.sortedCell, ws.cells.geom, ws.cells.cellStart, positions, charges,
            exclusions, exclusions ? maxExcl : 0, cutoff, alpha, coulomb, posDeriv, chargeDeriv, ws.acc);
        count_launch();
    }
    pme_publish_kernel<<<1, 1, 0, stream>>>(ws.acc, energy);
    count_launch();
    NNP_CUDA_CHECK(cudaGetLastError());
}

}  // namespace nnpops
