// ANI AEV forward/backward kernels for sm_100a.  See ani_aev.cuh for the contract and DESIGN.md for the layout.
//
// Pipeline per forward:  cell list  ->  ani_rows_kernel (per-atom neighbour rows, grouped by species, written once and
// reused by the four compute kernels)  ->  ani_radial_fwd_kernel + ani_angular_fwd_kernel.
// Backward: ani_radial_bwd_kernel (gather-only, plain store) -> ani_angular_bwd_kernel (centre-owned triples, neighbour
// forces exchanged by warp shuffles, one red.add per neighbour component).
//
// Mathematics: SURVEY.md appendix A.1/A.2; the behaviour being matched is CpuANISymmetryFunctions.cpp:112-194 (forward)
// and :228-353 (backward).  cos(theta - thetas) is evaluated as c*cos(thetas) + sqrt(1-c^2)*sin(thetas) instead of through
// acosf/cosf, and a^zeta as ex2(zeta*lg2(a)).
#include "ani_aev.cuh"
#include "ani_angular_v2.cuh"
#include <cuda_fp16.h>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <type_traits>

namespace nnpops {

namespace {

constexpr int kWPB = 8;   // warps (= centre atoms) per CTA
constexpr int kCellReach = 2;   // the ANI cell list has cells of half the cutoff: neighbours lie within +-2 cells

__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2a(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rsqrta(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// One AEV element: fp32, or the fp16 hi/lo pair consumed by the tensor-core MLP (hi = fp16(v), lo = fp16((v - hi) * 2^11))
struct AevOut {
    float* f32;
    __half* hi;
    __half* lo;
};
__device__ __forceinline__ void store_aev(const AevOut& o, size_t idx, float v) {
    if (o.hi) {
        const __half h = __float2half_rn(v);
        o.hi[idx] = h;
        o.lo[idx] = __float2half_rn((v - __half2float(h)) * 2048.0f);
    } else {
        o.f32[idx] = v;
    }
}

// zero `count` consecutive AEV elements starting at idx (warp-cooperative); 16-byte stores when the range allows it
__device__ __forceinline__ void zero_aev_range(const AevOut& o, size_t idx, int count, int lane) {
    if (o.hi) {
        __half* hi = o.hi + idx;
        __half* lo = o.lo + idx;
        if ((count & 7) == 0 && ((reinterpret_cast<uintptr_t>(hi) | reinterpret_cast<uintptr_t>(lo)) & 15) == 0) {
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
            const int units = count >> 3;                      // 16-byte units per array
            for (int u = lane; u < units; u += 32) { reinterpret_cast<uint4*>(hi)[u] = z; reinterpret_cast<uint4*>(lo)[u] = z; }
        } else {
            for (int i = lane; i < count; i += 32) { hi[i] = __float2half_rn(0.0f); lo[i] = __float2half_rn(0.0f); }
        }
    } else {
        float* f = o.f32 + idx;
        if ((count & 3) == 0 && (reinterpret_cast<uintptr_t>(f) & 15) == 0) {
            const int units = count >> 2;
            for (int u = lane; u < units; u += 32) reinterpret_cast<float4*>(f)[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            for (int i = lane; i < count; i += 32) f[i] = 0.0f;
        }
    }
}

// zero the blocks [p*nA, (p+1)*nA) of one AEV row for every pair p whose bit is set in emptyMask (nPairs <= 32): one pass of
// 16-byte stores over the row instead of one call per empty block
__device__ __forceinline__ void zero_aev_blocks(const AevOut& o, size_t rowIdx, int nA, int nPairs, unsigned emptyMask, int lane) {
    const int total = nPairs * nA;
    if (o.hi) {
        __half* hi = o.hi + rowIdx;
        __half* lo = o.lo + rowIdx;
        if ((nA & 7) == 0 && ((reinterpret_cast<uintptr_t>(hi) | reinterpret_cast<uintptr_t>(lo)) & 15) == 0) {
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
            for (int u = lane; u < (total >> 3); u += 32)
                if ((emptyMask >> ((u << 3) / nA)) & 1u) { reinterpret_cast<uint4*>(hi)[u] = z; reinterpret_cast<uint4*>(lo)[u] = z; }
        } else {
            for (int i = lane; i < total; i += 32)
                if ((emptyMask >> (i / nA)) & 1u) { hi[i] = __float2half_rn(0.0f); lo[i] = __float2half_rn(0.0f); }
        }
    } else {
        float* f = o.f32 + rowIdx;
        if ((nA & 3) == 0 && (reinterpret_cast<uintptr_t>(f) & 15) == 0) {
            for (int u = lane; u < (total >> 2); u += 32)
                if ((emptyMask >> ((u << 2) / nA)) & 1u) reinterpret_cast<float4*>(f)[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            for (int i = lane; i < total; i += 32)
                if ((emptyMask >> (i / nA)) & 1u) f[i] = 0.0f;
        }
    }
}

__device__ __forceinline__ int pair_index(int S, int s, int t) {   // CpuANISymmetryFunctions.cpp:39-43
    int lo = min(s, t), hi = max(s, t);
    return lo * S - (lo * (lo - 1)) / 2 + (hi - lo);
}

// ------------------------------------------------------------------------------------------------------------------
// Neighbour rows.  One warp per centre atom (sorted order).  Candidates come from the <= 18 contiguous runs of the cell
// list; the accept test is the reference's fp32 expression (strict r2 < Rc^2, CpuANISymmetryFunctions.cpp:129-135).
// ------------------------------------------------------------------------------------------------------------------
#ifndef NNP_ROWS_MINB
#define NNP_ROWS_MINB 5
#endif
template <bool SKIN>
__global__ void __launch_bounds__(kWPB * 32, NNP_ROWS_MINB)
ani_rows_kernel(int n, const float4* __restrict__ sorted, const int* __restrict__ sortedCell, const Geom* __restrict__ geom,
                const int* __restrict__ cellStart, const AniTables* __restrict__ tab, int capR, int capA,
                int* __restrict__ rowRad, int* __restrict__ rowAng, int* __restrict__ offRad, int* __restrict__ offAng,
                int* __restrict__ flag, const int* __restrict__ sortedOrig, const unsigned char* __restrict__ owned,
                const int* __restrict__ rebuild, float skinCut2, int capC, int* __restrict__ candRow, int* __restrict__ candCnt) {
    // Verlet skin (candRow != nullptr): on a rebuild step (*rebuild != 0) the candidates of the cell-list scan that lie within
    // (Rcr + skin)^2 are also written to candRow; on the other steps the candidates come from candRow instead of the cells -- about
    // 75 distance tests per centre instead of 380 -- and every one takes the minimum-image step.  Either way the rows hold exactly
    // the atoms with r2 < Rcr^2 at the CURRENT positions.
    extern __shared__ unsigned char smemRaw[];
    __shared__ Geom g;
    if (threadIdx.x == 0) g = *geom;
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = blockIdx.x * kWPB + w;
    if (p >= n) return;
    constexpr bool useSkin = SKIN;
    const bool scanCells = !useSkin || *rebuild != 0;
    if (owned != nullptr && !owned[sortedOrig[p]]) {
        // a centre of another rank (one box sharded over several GPUs): empty rows, so every downstream kernel skips it
        const int S1 = tab->nSpecies + 1;
        for (int i = lane; i < S1; i += 32) { offRad[(size_t)p * S1 + i] = 0; offAng[(size_t)p * S1 + i] = 0; }
        return;
    }
    uint32_t* list = reinterpret_cast<uint32_t*>(smemRaw) + (size_t)w * capR;
    int* cnt = reinterpret_cast<int*>(reinterpret_cast<uint32_t*>(smemRaw) + (size_t)kWPB * capR) + w * 128;
    int* cntR = cnt, *cntA = cnt + 32, *curR = cnt + 64, *curA = cnt + 96;
    const int S = tab->nSpecies;
    const float rcr2 = tab->rcr2, rca2 = tab->rca2;
    const float4 ci = sorted[p];
    int count = 0;
    // the minimum-image step is skipped for runs that do not cross a periodic face (it subtracts exactly zero there, see
    // for_each_candidate_run_w): 3 FRND on the XU pipe and 9 more instructions per candidate
    const bool alwaysImage = g.periodic && (g.triclinic || g.anyOutside);
    int nCand = 0;
    // IMAGE is a compile-time tag: as a run-time flag the compiler predicates the minimum-image step, and its 14 predicated-off
    // instructions still take issue slots in the hottest loop of the kernel
    auto visit = [&](int q, bool valid, const float4& cj, auto imageTag) {
        constexpr bool image = decltype(imageTag)::value;
        bool ok = false, cand = false;
        uint32_t packed = 0;
        if (valid && q != p) {
            float dx = __fsub_rn(cj.x, ci.x), dy = __fsub_rn(cj.y, ci.y), dz = __fsub_rn(cj.z, ci.z);
            float r2;
            if (image) r2 = min_image_mul(g, dx, dy, dz);
            else r2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            cand = r2 < skinCut2;
            if (r2 < rcr2) {
                ok = true;
                packed = (uint32_t)q | ((uint32_t)__float_as_int(cj.w) & 0x7f000000u) | (r2 < rca2 ? 0x80000000u : 0u);
            }
        }
        const unsigned m = __ballot_sync(kFull, ok);
        if (ok) {
            const int idx = count + __popc(m & ((1u << lane) - 1u));
            if (idx < capR) list[idx] = packed;
        }
        count += __popc(m);
        if (useSkin && scanCells) {   // remember every atom within the skin cutoff
            const unsigned mc = __ballot_sync(kFull, cand);
            if (cand) {
                const int idx = nCand + __popc(mc & ((1u << lane) - 1u));
                if (idx < capC) candRow[(size_t)p * capC + idx] = q;
            }
            nCand += __popc(mc);
        }
    };
    if (scanCells) {
        // Candidates: the cell list of this class has HALF-size cells (>= 1.001 * cutoff / 2 wide), so the neighbourhood is the
        // 5 x 5 x 5 block of cells around the centre's: (2.5 Rc)^3 of volume instead of (3 Rc)^3 -- 1.7 x fewer distance tests.  The
        // block is 25 z-columns of 5 contiguous cells.  Lane l < 25 derives the bounds of column l (two runs when the column crosses the
        // periodic z face), a warp scan lays all runs end to end, and the candidates are then walked 32 at a time over the
        // CONCATENATION -- every lane holds a candidate, where a walk run by run leaves most of the last chunk of every run empty.
        constexpr int R = kCellReach;
        const int c = sortedCell[p];
        const int nx = g.nc[0], ny = g.nc[1], nz = g.nc[2];
        const int cz = c % nz, cy = (c / nz) % ny, cx = c / (nz * ny);
        const bool per = g.periodic != 0;
        // un-wrapped candidates of a periodic box lie within (R + 1) cells: the minimum-image step subtracts exactly zero when that is
        // less than half the box in every dimension (orthorhombic box, all atoms inside the primary cell)
        const bool coarse = per && (nx < 2 * R + 3 || ny < 2 * R + 3 || nz < 2 * R + 3);
        int x0, NX, y0, NY;
        if (per) { NX = min(nx, 2 * R + 1); x0 = nx >= 2 * R + 1 ? cx - R : 0; NY = min(ny, 2 * R + 1); y0 = ny >= 2 * R + 1 ? cy - R : 0; }
        else { x0 = max(cx - R, 0); NX = min(cx + R, nx - 1) - x0 + 1; y0 = max(cy - R, 0); NY = min(cy + R, ny - 1) - y0 + 1; }
        int b0 = 0, e0 = 0, b1 = 0, e1 = 0;
        bool w0 = coarse || alwaysImage;
        if (lane < NX * NY) {
            int ix = x0 + lane / NY, iy = y0 + lane % NY;
            if (ix < 0) { ix += nx; w0 = true; } else if (ix >= nx) { ix -= nx; w0 = true; }
            if (iy < 0) { iy += ny; w0 = true; } else if (iy >= ny) { iy -= ny; w0 = true; }
            const int base = (ix * ny + iy) * nz;
            int za, zb, za1 = 0, zb1 = -1;
            if (!per) { za = max(cz - R, 0); zb = min(cz + R, nz - 1); }
            else if (nz < 2 * R + 1) { za = 0; zb = nz - 1; }
            else {
                za = cz - R; zb = cz + R;
                if (za < 0) { za1 = za + nz; zb1 = nz - 1; za = 0; }
                else if (zb >= nz) { za1 = 0; zb1 = zb - nz; zb = nz - 1; }
            }
            b0 = cellStart[base + za]; e0 = cellStart[base + zb + 1];
            if (zb1 >= za1) { b1 = cellStart[base + za1]; e1 = cellStart[base + zb1 + 1]; }
        }
        const int len0 = e0 - b0, len1 = e1 - b1;
        int incl = len0 + len1;
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(kFull, incl, o); if (lane >= o) incl += y; }
        const int T = __shfl_sync(kFull, incl, 31);
        int* sStart = cnt;          // [64] first flat index of run slot 2 l (first run of column l) and 2 l + 1 (its wrapped run)
        int* sRunB = cnt + 64;      // [64] first sorted index of the run, bit 31 = reached across a periodic face
        sStart[2 * lane] = incl - len0 - len1; sStart[2 * lane + 1] = incl - len1;
        sRunB[2 * lane] = b0 | (w0 ? (int)0x80000000 : 0); sRunB[2 * lane + 1] = b1 | (int)0x80000000;
        __syncwarp();
        auto locate = [&](int f, int& q, bool& wr) {      // flat candidate index -> sorted index: the last slot that starts at or before f
            int sl = 0;
#pragma unroll
            for (int step = 32; step >= 1; step >>= 1) sl += (sStart[sl + step] <= f) ? step : 0;
            const int rb = sRunB[sl];
            q = (rb & 0x7fffffff) + f - sStart[sl];
            wr = rb < 0;
        };
        int q = p;
        bool wr = false;
        float4 cur = ci;
        if (lane < T) { locate(lane, q, wr); cur = sorted[q]; }
        for (int f0 = 0; f0 < T; f0 += 32) {
            // the coordinates of the next 32 candidates are requested before this chunk is tested (the loop is L2-latency bound)
            int qn = p;
            bool wn = false;
            float4 nxt = ci;
            if (f0 + 32 + lane < T) { locate(f0 + 32 + lane, qn, wn); nxt = sorted[qn]; }
            const bool valid = f0 + lane < T;
            if (__any_sync(kFull, valid && wr)) visit(q, valid, cur, std::true_type{});
            else visit(q, valid, cur, std::false_type{});
            q = qn; wr = wn; cur = nxt;
        }
        __syncwarp();               // the run tables share their shared memory with the species counters used below
        if (useSkin) {
            if (nCand > capC) { if (lane == 0) atomicOr(flag, 1); nCand = capC; }
            if (lane == 0) candCnt[p] = nCand;
        }
    } else {
        const int nc = candCnt[p];
        if (g.periodic) {
            for (int q0 = 0; q0 < nc; q0 += 32) {
                const bool valid = q0 + lane < nc;
                const int q = valid ? candRow[(size_t)p * capC + q0 + lane] : p;
                visit(q, valid, sorted[q], std::true_type{});
            }
        } else {
            for (int q0 = 0; q0 < nc; q0 += 32) {
                const bool valid = q0 + lane < nc;
                const int q = valid ? candRow[(size_t)p * capC + q0 + lane] : p;
                visit(q, valid, sorted[q], std::false_type{});
            }
        }
    }
    if (count > capR) { if (lane == 0) atomicOr(flag, 1); count = capR; }
    cntR[lane] = 0; cntA[lane] = 0;
    __syncwarp();
    for (int i = lane; i < count; i += 32) {
        const uint32_t e = list[i];
        const int s = (e >> 24) & 0x7f;
        atomicAdd(&cntR[s], 1);
        if (e & 0x80000000u) atomicAdd(&cntA[s], 1);
    }
    __syncwarp();
    {   // exclusive scans over species (lane = species)
        int vr = lane < S ? cntR[lane] : 0, va = lane < S ? cntA[lane] : 0;
        int xr = vr, xa = va;
        for (int o = 1; o < 32; o <<= 1) {
            int yr = __shfl_up_sync(kFull, xr, o), ya = __shfl_up_sync(kFull, xa, o);
            if (lane >= o) { xr += yr; xa += ya; }
        }
        const int totR = __shfl_sync(kFull, xr, 31), totA = __shfl_sync(kFull, xa, 31);
        if (totA > capA && lane == 0) atomicOr(flag, 2);
        curR[lane] = xr - vr; curA[lane] = xa - va;
        if (lane < S) {
            offRad[(size_t)p * (S + 1) + lane] = xr - vr;
            offAng[(size_t)p * (S + 1) + lane] = min(xa - va, capA);
        }
        if (lane == 0) {
            offRad[(size_t)p * (S + 1) + S] = totR;
            offAng[(size_t)p * (S + 1) + S] = min(totA, capA);
        }
    }
    __syncwarp();
    // ordered placement: chunks of 32 entries in list order, rank inside a chunk from match_any -> deterministic rows
    for (int base = 0; base < count; base += 32) {
        const int i = base + lane;
        const bool valid = i < count;
        const uint32_t e = valid ? list[i] : 0u;
        const int s = valid ? (int)((e >> 24) & 0x7f) : 0xff;
        const bool ang = valid && (e & 0x80000000u);
        const unsigned lt = (1u << lane) - 1u;
        const unsigned mr = __match_any_sync(kFull, s);
        const unsigned ma = __match_any_sync(kFull, ang ? s : 0xff);
        int dstR = 0, dstA = 0;
        if (valid) dstR = curR[s] + __popc(mr & lt);
        if (ang) dstA = curA[s] + __popc(ma & lt);
        __syncwarp();
        if (valid && (mr & lt) == 0) curR[s] += __popc(mr);
        if (ang && (ma & lt) == 0) curA[s] += __popc(ma);
        __syncwarp();
        const int j = (int)(e & 0x00ffffffu);
        if (valid) rowRad[(size_t)p * capR + dstR] = j;
        if (ang && dstA < capA) rowAng[(size_t)p * capA + dstA] = j;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Radial forward:  radial[i][s][k] = scale * sum_{j in species s} fc(r_ij) exp(-eta_k (r_ij - Rs_k)^2)
// lane = (h, k): k = radial function, h = neighbour sub-stream; accumulators live in registers, no atomics, one store/element.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWPB * 32)
ani_radial_fwd_kernel(int n, const float4* __restrict__ sorted, const int* __restrict__ sortedOrig, const Geom* __restrict__ geom,
                      const AniTables* __restrict__ tab, const int* __restrict__ rowRad, const int* __restrict__ offRad, int capR,
                      const int* __restrict__ rowMap, AevOut out, int stride) {
    extern __shared__ unsigned char smemRaw[];
    __shared__ Geom g;
    __shared__ float sEtaL2[kAniMaxRadial], sShf[kAniMaxRadial];
    const int nR = tab->nRadial, S = tab->nSpecies;
    if (threadIdx.x == 0) g = *geom;
    for (int i = threadIdx.x; i < nR; i += blockDim.x) { sEtaL2[i] = tab->rEtaL2[i]; sShf[i] = tab->rShf[i]; }
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = blockIdx.x * kWPB + w;
    if (p >= n) return;
    float* sr = reinterpret_cast<float*>(smemRaw) + (size_t)w * 2 * capR;
    float* sfc = sr + capR;
    const int* off = offRad + (size_t)p * (S + 1);
    const int cnt = min(off[S], capR);
    const float4 ci = sorted[p];
    const float rcr = tab->rcr;
    const float kf = kPi / rcr;
    for (int q = lane; q < cnt; q += 32) {
        const float4 cj = sorted[rowRad[(size_t)p * capR + q]];
        float dx = __fsub_rn(cj.x, ci.x), dy = __fsub_rn(cj.y, ci.y), dz = __fsub_rn(cj.z, ci.z);
        const float r = sqrtf(min_image_mul(g, dx, dy, dz));
        sr[q] = r;
        sfc[q] = 0.5f * cosf(r * kf) + 0.5f;
    }
    __syncwarp();
    const int orig = sortedOrig[p];
    const size_t orow = (size_t)(rowMap ? rowMap[orig] : orig) * stride;
    const float scale = tab->radialScale;
    for (int k0 = 0; k0 < nR; k0 += 32) {
        const int kk = min(nR - k0, 32);
        int KP = 1;
        while (KP < kk) KP <<= 1;
        const int H = 32 / KP, k = lane % KP, h = lane / KP;
        const bool kval = k < kk;
        const float eta2 = kval ? sEtaL2[k0 + k] : 0.0f, shf = kval ? sShf[k0 + k] : 0.0f;
        for (int s = 0; s < S; s++) {
            const int b = off[s], e = min(off[s + 1], capR);
            float acc = 0.0f;
            for (int q = b + h; q < e; q += H) {
                const float t = sr[q] - shf;
                acc = fmaf(sfc[q], ex2a(-eta2 * t * t), acc);
            }
            for (int o = KP; o < 32; o <<= 1) acc += __shfl_xor_sync(kFull, acc, o);
            if (h == 0 && kval) store_aev(out, orow + s * nR + k0 + k, acc * scale);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Angular forward.  Neighbours of the centre are staged in shared memory grouped by species, so every species-pair block
// (s, t) is a dense index range of triples.  lane <-> triple; each lane keeps 32 channel accumulators in registers; a
// butterfly transpose-reduce leaves lane m with channel m, written as one coalesced 128-byte store with the 2^(1-zeta)
// scale folded in (the reference's separate scale pass, K3, disappears).
// MODE 0: generic function table (any list of angular functions, literal evaluation per channel)
// MODE 1: factorised  m = a*NSZ + z  with a single EtaA and a single Zeta (ANI-1x/2x shapes): NSZ pow + NSA exp per triple
// ------------------------------------------------------------------------------------------------------------------
struct AngTablesSmem {
    float etaL2[kAniMaxAngular], eta[kAniMaxAngular], shf[kAniMaxAngular], zeta[kAniMaxAngular], cosT[kAniMaxAngular],
        sinT[kAniMaxAngular], scale[kAniMaxAngular];
};

__device__ __forceinline__ void load_ang_tables(AngTablesSmem& t, const AniTables* __restrict__ tab) {
    const int nA = tab->nAngular;
    for (int i = threadIdx.x; i < nA; i += blockDim.x) {
        t.etaL2[i] = tab->aEtaL2[i]; t.eta[i] = tab->aEta[i]; t.shf[i] = tab->aShf[i]; t.zeta[i] = tab->aZeta[i];
        t.cosT[i] = tab->aCos[i]; t.sinT[i] = tab->aSin[i]; t.scale[i] = tab->aScale[i];
    }
}

// lane L ends with sum over lanes of v[L]
__device__ __forceinline__ float transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
    for (int off = 16, cnt = 16; off >= 1; off >>= 1, cnt >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < cnt; i++) {
            const float send = up ? v[i] : v[i + cnt];
            const float keep = up ? v[i + cnt] : v[i];
            v[i] = keep + __shfl_xor_sync(kFull, send, off);
        }
    }
    return v[0];
}

template <int MODE, int NSA, int NSZ, bool TORCHANI>
__global__ void __launch_bounds__(kWPB * 32, 3)
ani_angular_fwd_kernel(int n, const float4* __restrict__ sorted, const int* __restrict__ sortedOrig, const Geom* __restrict__ geom,
                       const AniTables* __restrict__ tab, const int* __restrict__ rowAng, const int* __restrict__ offAng, int capA,
                       const int* __restrict__ rowMap, AevOut out, int stride) {
    extern __shared__ unsigned char smemRaw[];
    __shared__ Geom g;
    __shared__ AngTablesSmem T;
    __shared__ float fShfA[kAniMaxShf], fCos[kAniMaxShf], fSin[kAniMaxShf];
    if (threadIdx.x == 0) g = *geom;
    load_ang_tables(T, tab);
    if (threadIdx.x < kAniMaxShf) { fShfA[threadIdx.x] = tab->fShfA[threadIdx.x]; fCos[threadIdx.x] = tab->fCos[threadIdx.x]; fSin[threadIdx.x] = tab->fSin[threadIdx.x]; }
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = blockIdx.x * kWPB + w;
    if (p >= n) return;
    const int S = tab->nSpecies, nA = tab->nAngular;
    float* sdx = reinterpret_cast<float*>(smemRaw) + (size_t)w * 6 * capA;
    float *sdy = sdx + capA, *sdz = sdy + capA, *sr = sdz + capA, *sir = sr + capA, *sfc = sir + capA;
    const int* off = offAng + (size_t)p * (S + 1);
    const int cnt = min(off[S], capA);
    const float4 ci = sorted[p];
    const float kf = kPi / tab->rca;
    for (int q = lane; q < cnt; q += 32) {
        const float4 cj = sorted[rowAng[(size_t)p * capA + q]];
        float dx = __fsub_rn(cj.x, ci.x), dy = __fsub_rn(cj.y, ci.y), dz = __fsub_rn(cj.z, ci.z);
        const float r = sqrtf(min_image_mul(g, dx, dy, dz));
        sdx[q] = dx; sdy[q] = dy; sdz[q] = dz; sr[q] = r; sir[q] = 1.0f / r;
        sfc[q] = 0.5f * cosf(r * kf) + 0.5f;
    }
    __syncwarp();
    const int orig = sortedOrig[p];
    const size_t orow = (size_t)(rowMap ? rowMap[orig] : orig) * stride;
    const float cosScale = tab->cosScale;
    const float fEtaL2 = tab->fEtaL2, fZeta = tab->fZeta, fScale = tab->fScale;

    // species offsets of this centre live in registers (lane s holds segment s) and are broadcast by shuffle; species without
    // angular neighbours are skipped wholesale and their output blocks zero-filled with wide stores
    const int myB = lane < S ? min(off[lane], capA) : 0;
    const int myN = (lane < S ? min(off[lane + 1], capA) : 0) - myB;
    const int nPairs = tab->nPairs;
    const bool maskZero = nPairs <= 32;      // empty blocks are collected in a bit mask and zero-filled in one pass at the end
    unsigned nonEmpty = 0u;
    for (int m0 = 0; m0 < nA; m0 += 32) {
        int pIdx = 0;
        for (int s = 0; s < S; s++) {
            const int bs = __shfl_sync(kFull, myB, s), ns = __shfl_sync(kFull, myN, s);
            if (ns == 0) {   // every pair (s, t >= s) is empty
                if (m0 == 0 && !maskZero) zero_aev_range(out, orow + (size_t)pIdx * nA, (S - s) * nA, lane);
                pIdx += S - s;
                continue;
            }
            for (int t = s; t < S; t++, pIdx++) {
                const int bt = __shfl_sync(kFull, myB, t), nt = __shfl_sync(kFull, myN, t);
                const int ntrip = (s == t) ? (ns * (ns - 1)) / 2 : ns * nt;
                const size_t dst = orow + (size_t)pIdx * nA + m0;
                if (ntrip <= 0) {
                    if (m0 == 0 && !maskZero) zero_aev_range(out, orow + (size_t)pIdx * nA, nA, lane);
                    continue;
                }
                if (maskZero) nonEmpty |= 1u << pIdx;
                const float invNt = 1.0f / (float)nt;
                float acc[32];
#pragma unroll
                for (int i = 0; i < 32; i++) acc[i] = 0.0f;
                for (int q0 = 0; q0 < ntrip; q0 += 32) {
                    const int q = q0 + lane;
                    const bool valid = q < ntrip;
                    const int qq = valid ? q : 0;
                    int ia, ib;
                    if (s == t) {   // unordered pairs a < b inside one segment: qq = b(b-1)/2 + a
                        // b = floor((1 + sqrt(1 + 8 q)) / 2) from an approximate root, fixed by one exact integer comparison each way
                        int b = (int)fmaf(__fsqrt_rn(fmaf(8.0f, (float)qq, 1.0f)), 0.5f, 0.5f);
                        b -= ((b * (b - 1)) >> 1) > qq ? 1 : 0;
                        b += ((b * (b + 1)) >> 1) <= qq ? 1 : 0;
                        ia = bs + qq - ((b * (b - 1)) >> 1); ib = bs + b;
                    } else {
                        const int a = (int)(((float)qq + 0.5f) * invNt);   // exact: the quotient is >= 0.5/nt away from an integer
                        ia = bs + a; ib = bt + qq - a * nt;
                    }
                    const float ax = sdx[ia], ay = sdy[ia], az = sdz[ia], bx = sdx[ib], by = sdy[ib], bz = sdz[ib];
                    const float dot = ax * bx + ay * by + az * bz;
                    const float ipr = sir[ia] * sir[ib];
                    const float c = cosScale * dot * ipr;
                    float sn;
                    if (TORCHANI) {
                        sn = sqrtf(fmaxf(1.0f - c * c, 0.0f));
                    } else {   // the reference switches to asin of the cross product near |cos| = 1 (CpuANISymmetryFunctions.cpp:396-404)
                        const float cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx;
                        sn = sqrtf(cx * cx + cy * cy + cz * cz) * ipr;
                    }
                    const float rm = 0.5f * (sr[ia] + sr[ib]);
                    const float F = valid ? sfc[ia] * sfc[ib] : 0.0f;
                    if (MODE == 1) {
                        float P[NSZ];
#pragma unroll
                        for (int z = 0; z < NSZ; z++) {
                            const float base = fmaxf(1.0f + c * fCos[z] + sn * fSin[z], 0.0f);
                            P[z] = F * ex2a(fZeta * lg2a(base));
                        }
#pragma unroll
                        for (int a = 0; a < NSA; a++) {
                            const float tt = rm - fShfA[a];
                            const float E = ex2a(-fEtaL2 * tt * tt);
#pragma unroll
                            for (int z = 0; z < NSZ; z++) acc[a * NSZ + z] = fmaf(P[z], E, acc[a * NSZ + z]);
                        }
                    } else {
#pragma unroll
                        for (int mm = 0; mm < 32; mm++) {
                            const int m = m0 + mm;
                            if (m < nA) {
                                const float base = fmaxf(1.0f + c * T.cosT[m] + sn * T.sinT[m], 0.0f);
                                const float tt = rm - T.shf[m];
                                acc[mm] = fmaf(F * ex2a(T.zeta[m] * lg2a(base)), ex2a(-T.etaL2[m] * tt * tt), acc[mm]);
                            }
                        }
                    }
                }
                const float v = transpose_reduce32(acc, lane);
                if (m0 + lane < nA) store_aev(out, dst + lane, v * (MODE == 1 ? fScale : T.scale[m0 + lane]));
            }
        }
    }
    if (maskZero) zero_aev_blocks(out, orow, nA, nPairs, ~nonEmpty, lane);
}

// ------------------------------------------------------------------------------------------------------------------
// Angular forward, grouped form (factorised tables, <= 32 species pairs): G lanes per centre, 32 / G centres per warp.
// A liquid-density centre has ~140 triples spread over a few species-pair blocks of 15-70 triples; with a whole warp per
// centre the last iteration of every block runs mostly empty lanes and the 32-lane transpose-reduce is paid per block.
// Groups of G = 8 lanes fill their iterations (ceil(ntrip / 8) * 8 slots) and run the reduction of 4 centres at once
// (3 butterfly levels instead of 5); lane gl of a group ends with channels [gl * 32 / G, (gl + 1) * 32 / G) and stores them
// as one vector.  The block loop is warp-uniform (all groups visit the same species pair; the trip count is the warp maximum).
// ------------------------------------------------------------------------------------------------------------------
template <int G, int NSA, int NSZ, bool TORCHANI>
__global__ void __launch_bounds__(kWPB * 32)
ani_angular_fwd_grouped_kernel(int n, const float4* __restrict__ sorted, const int* __restrict__ sortedOrig, const Geom* __restrict__ geom,
                               const AniTables* __restrict__ tab, const int* __restrict__ rowAng, const int* __restrict__ offAng, int capA,
                               const int* __restrict__ rowMap, AevOut out, int stride) {
    static_assert(NSA * NSZ == 32, "the grouped kernel assumes 32 angular channels");
    constexpr int CPW = 32 / G;   // centres per warp
    constexpr int R = 32 / G;     // channels per lane after the in-group reduction
    extern __shared__ unsigned char smemRaw[];
    __shared__ Geom g;
    __shared__ float fShfA[kAniMaxShf], fCos[kAniMaxShf], fSin[kAniMaxShf];
    if (threadIdx.x == 0) g = *geom;
    if (threadIdx.x < kAniMaxShf) { fShfA[threadIdx.x] = tab->fShfA[threadIdx.x]; fCos[threadIdx.x] = tab->fCos[threadIdx.x]; fSin[threadIdx.x] = tab->fSin[threadIdx.x]; }
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, sub = lane / G, gl = lane % G;
    const int p0 = (blockIdx.x * kWPB + w) * CPW;
    if (p0 >= n) return;                       // warp-uniform
    const int p = p0 + sub;
    const bool centre = p < n;
    const int S = tab->nSpecies, nA = 32, nPairs = tab->nPairs;
    // per centre: float4 {dx, dy, dz, r} and float2 {1 / r, fc} per neighbour, then the species offsets
    const size_t perCentre = (size_t)6 * capA + (kAniMaxSpecies + 1 + 3) / 4 * 4;   // in floats, a multiple of 4
    float* cbase = reinterpret_cast<float*>(smemRaw) + ((size_t)w * CPW + sub) * perCentre;
    float4* sv = reinterpret_cast<float4*>(cbase);
    float2* sq = reinterpret_cast<float2*>(cbase + 4 * capA);
    int* sOff = reinterpret_cast<int*>(cbase + 6 * capA);
    const int pc = centre ? p : p0;            // lanes of a missing centre mirror the first one with zero neighbours
    const int* off = offAng + (size_t)pc * (S + 1);
    const int cnt = centre ? min(off[S], capA) : 0;
    for (int i = gl; i <= S; i += G) sOff[i] = centre ? min(off[i], capA) : 0;
    const float4 ci = sorted[pc];
    const float kf = kPi / tab->rca;
    for (int q = gl; q < cnt; q += G) {
        const float4 cj = sorted[rowAng[(size_t)p * capA + q]];
        float dx = __fsub_rn(cj.x, ci.x), dy = __fsub_rn(cj.y, ci.y), dz = __fsub_rn(cj.z, ci.z);
        const float r = sqrtf(min_image_mul(g, dx, dy, dz));
        sv[q] = make_float4(dx, dy, dz, r);
        sq[q] = make_float2(1.0f / r, 0.5f * cosf(r * kf) + 0.5f);
    }
    __syncwarp();
    const int orig = sortedOrig[pc];
    const size_t orow = (size_t)(rowMap ? rowMap[orig] : orig) * stride;
    const float cosScale = tab->cosScale;
    const float fEtaL2 = tab->fEtaL2, fZeta = tab->fZeta, fScale = tab->fScale;
    unsigned written = 0u;                     // warp-uniform: species pairs stored by every centre of the warp
    int pIdx = 0;
    for (int s = 0; s < S; s++) {
        const int bs = sOff[s], ns = sOff[s + 1] - bs;
        for (int t = s; t < S; t++, pIdx++) {
            const int bt = sOff[t], nt = sOff[t + 1] - bt;
            const int ntrip = (s == t) ? (ns * (ns - 1)) / 2 : ns * nt;
            const int maxTrip = __reduce_max_sync(kFull, ntrip);
            if (maxTrip <= 0) continue;
            written |= 1u << pIdx;
            // lane gl starts at triple q = gl and advances by G: (ia, ib) are kept incrementally.  Same species: unordered pairs
            // a < b with q = b (b - 1) / 2 + a; different species: q = a * nt + b.
            int ia, ib;
            if (s == t) {
                int bq = (int)fmaf(__fsqrt_rn(fmaf(8.0f, (float)gl, 1.0f)), 0.5f, 0.5f);
                bq -= ((bq * (bq - 1)) >> 1) > gl ? 1 : 0;
                bq += ((bq * (bq + 1)) >> 1) <= gl ? 1 : 0;
                ia = gl - ((bq * (bq - 1)) >> 1); ib = bq;
            } else {
                const int nt1 = max(nt, 1);
                ia = gl / nt1; ib = gl - ia * nt1;
            }
            float acc[32];
#pragma unroll
            for (int i = 0; i < 32; i++) acc[i] = 0.0f;
            for (int q = gl; q < maxTrip; q += G) {
                const bool valid = q < ntrip;
                const int ja = valid ? bs + ia : 0, jb = valid ? (s == t ? bs : bt) + ib : 0;
                const float4 va = sv[ja], vb = sv[jb];
                const float2 qa = sq[ja], qb = sq[jb];
                const float dot = va.x * vb.x + va.y * vb.y + va.z * vb.z;
                const float ipr = qa.x * qb.x;
                const float c = cosScale * dot * ipr;
                float sn;
                if (TORCHANI) {
                    const float x = fmaxf(fmaf(-c, c, 1.0f), 1e-30f);
                    sn = x * rsqrta(x);
                } else {
                    const float cx = va.y * vb.z - va.z * vb.y, cy = va.z * vb.x - va.x * vb.z, cz = va.x * vb.y - va.y * vb.x;
                    sn = sqrtf(cx * cx + cy * cy + cz * cz) * ipr;
                }
                const float rm = 0.5f * (va.w + vb.w);
                const float F = valid ? qa.y * qb.y : 0.0f;
                float P[NSZ];
#pragma unroll
                for (int z = 0; z < NSZ; z++) {
                    const float base = fmaxf(1.0f + c * fCos[z] + sn * fSin[z], 0.0f);
                    P[z] = F * ex2a(fZeta * lg2a(base));
                }
#pragma unroll
                for (int a = 0; a < NSA; a++) {
                    const float tt = rm - fShfA[a];
                    const float E = ex2a(-fEtaL2 * tt * tt);
#pragma unroll
                    for (int z = 0; z < NSZ; z++) acc[a * NSZ + z] = fmaf(P[z], E, acc[a * NSZ + z]);
                }
                // advance the pair by G triples
                if (s == t) {
                    ia += G;
                    while (ia >= ib) { ia -= ib; ib++; }
                } else {
                    ib += G;
                    while (ib >= nt && nt > 0) { ib -= nt; ia++; }
                }
            }
            // in-group transpose-reduce: log2(G) levels; lane gl keeps channels [gl * R, gl * R + R)
#pragma unroll
            for (int o = G / 2, c2 = 16; o >= 1; o >>= 1, c2 >>= 1) {
                const bool up = (gl & o) != 0;
#pragma unroll
                for (int i = 0; i < c2; i++) {
                    const float send = up ? acc[i] : acc[i + c2];
                    const float keep = up ? acc[i + c2] : acc[i];
                    acc[i] = keep + __shfl_xor_sync(kFull, send, o);
                }
            }
            if (centre) {
                const size_t dst = orow + (size_t)pIdx * nA + gl * R;
                if (out.hi) {
                    uint32_t wh[(R + 1) / 2], wl[(R + 1) / 2];
#pragma unroll
                    for (int i = 0; i < R; i += 2) {
                        const float v0 = acc[i] * fScale, v1 = (i + 1 < R) ? acc[i + 1] * fScale : 0.0f;
                        const __half2 h2 = __floats2half2_rn(v0, v1);
                        const float2 f2 = __half22float2(h2);
                        const __half2 l2 = __floats2half2_rn((v0 - f2.x) * 2048.0f, (v1 - f2.y) * 2048.0f);
                        wh[i / 2] = *reinterpret_cast<const uint32_t*>(&h2);
                        wl[i / 2] = *reinterpret_cast<const uint32_t*>(&l2);
                    }
                    if (R == 4) {
                        *reinterpret_cast<uint2*>(out.hi + dst) = make_uint2(wh[0], wh[R > 2 ? 1 : 0]);
                        *reinterpret_cast<uint2*>(out.lo + dst) = make_uint2(wl[0], wl[R > 2 ? 1 : 0]);
                    } else if (R == 2) {
                        *reinterpret_cast<uint32_t*>(out.hi + dst) = wh[0];
                        *reinterpret_cast<uint32_t*>(out.lo + dst) = wl[0];
                    } else {
                        out.hi[dst] = __ushort_as_half((unsigned short)(wh[0] & 0xffffu));
                        out.lo[dst] = __ushort_as_half((unsigned short)(wl[0] & 0xffffu));
                    }
                } else {
                    if (R == 4) *reinterpret_cast<float4*>(out.f32 + dst) = make_float4(acc[0] * fScale, acc[1] * fScale, acc[2] * fScale, acc[3] * fScale);
                    else if (R == 2) *reinterpret_cast<float2*>(out.f32 + dst) = make_float2(acc[0] * fScale, acc[1] * fScale);
                    else out.f32[dst] = acc[0] * fScale;
                }
            }
        }
    }
    // species pairs that no centre of this warp populated: zero-fill, one centre at a time with the whole warp
    if (written != (nPairs >= 32 ? 0xffffffffu : (1u << nPairs) - 1u)) {
        for (int cidx = 0; cidx < CPW; cidx++) {
            if (p0 + cidx >= n) break;
            const size_t rowC = __shfl_sync(kFull, (unsigned long long)orow, cidx * G);
            zero_aev_blocks(out, rowC, nA, nPairs, ~written, lane);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Radial backward, gather-only: dE/dx_i = -sum_j w_ij delta_ij / r_ij with
// w_ij = scale * sum_k (G[i][s_j][k] + G[j][s_i][k]) * exp(..)(fc' - 2 eta (r - Rs_k) fc)   (CpuANISymmetryFunctions.cpp:228-263).
// Every directed pair is evaluated by its centre: no atomics inside the pair loop, one accumulate per centre at the end.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWPB * 32)
ani_radial_bwd_kernel(int n, const float4* __restrict__ sorted, const int* __restrict__ sortedOrig, const Geom* __restrict__ geom,
                      const AniTables* __restrict__ tab, const int* __restrict__ rowRad, const int* __restrict__ offRad, int capR,
                      const int* __restrict__ rowMap, const float* __restrict__ grad, int stride, float* __restrict__ posGrad) {
    // lane <-> neighbour: every lane evaluates all radial functions of its own pair, with the neighbour's gradient row segment
    // (nR consecutive floats) fetched by 16-byte loads that are all in flight before the first use; the centre's own gradient
    // row is staged once per warp in shared memory (indexed by the neighbour's species).  One warp reduction per centre.
    extern __shared__ unsigned char smemRaw[];
    __shared__ Geom g;
    __shared__ float sEtaL2[kAniMaxRadial], sEta[kAniMaxRadial], sShf[kAniMaxRadial];
    const int nR = tab->nRadial, S = tab->nSpecies;
    if (threadIdx.x == 0) g = *geom;
    for (int i = threadIdx.x; i < nR; i += blockDim.x) { sEtaL2[i] = tab->rEtaL2[i]; sEta[i] = tab->rEta[i]; sShf[i] = tab->rShf[i]; }
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = blockIdx.x * kWPB + w;
    if (p >= n) return;
    float* sGi = reinterpret_cast<float*>(smemRaw) + (size_t)w * S * nR;
    const int cnt = min(offRad[(size_t)p * (S + 1) + S], capR);
    const float4 ci = sorted[p];
    const int sp = __float_as_int(ci.w) >> 24;
    const int orig = sortedOrig[p];
    {
        const float* gi = grad + (size_t)(rowMap ? rowMap[orig] : orig) * stride;
        for (int i = lane; i < S * nR; i += 32) sGi[i] = gi[i];
    }
    __syncwarp();
    const float rcr = tab->rcr, kf = kPi / rcr;
    const bool vec = (nR & 3) == 0 && (stride & 3) == 0 && ((reinterpret_cast<uintptr_t>(grad) & 15) == 0);
    float fx = 0.0f, fy = 0.0f, fz = 0.0f;
    for (int q = lane; q < cnt; q += 32) {
        const int j = rowRad[(size_t)p * capR + q];
        const float4 cj = sorted[j];
        const int oj = sortedOrig[j];
        const float* gj = grad + (size_t)(rowMap ? rowMap[oj] : oj) * stride + sp * nR;
        float dx = __fsub_rn(cj.x, ci.x), dy = __fsub_rn(cj.y, ci.y), dz = __fsub_rn(cj.z, ci.z);
        const float r = sqrtf(min_image_mul(g, dx, dy, dz));
        const float ir = 1.0f / r;
        float sn, cs;
        sincosf(r * kf, &sn, &cs);
        const float fc = 0.5f * cs + 0.5f, dfc = -0.5f * kf * sn;
        const float* gi = sGi + (__float_as_int(cj.w) >> 24) * nR;
        float wsum = 0.0f;
        if (vec) {
            for (int k0 = 0; k0 < nR; k0 += 16) {
                float4 gv[4];
#pragma unroll
                for (int u = 0; u < 4; u++)
                    gv[u] = (k0 + 4 * u < nR) ? __ldg(reinterpret_cast<const float4*>(gj + k0 + 4 * u)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const float gq[4] = {gv[u].x, gv[u].y, gv[u].z, gv[u].w};
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const int k = k0 + 4 * u + e;
                        if (k < nR) {
                            const float t = r - sShf[k];
                            const float ex = ex2a(-sEtaL2[k] * t * t);
                            wsum = fmaf(gi[k] + gq[e], ex * (dfc - 2.0f * sEta[k] * t * fc), wsum);
                        }
                    }
                }
            }
        } else {
            for (int k = 0; k < nR; k++) {
                const float t = r - sShf[k];
                const float ex = ex2a(-sEtaL2[k] * t * t);
                wsum = fmaf(gi[k] + gj[k], ex * (dfc - 2.0f * sEta[k] * t * fc), wsum);
            }
        }
        const float wr = wsum * ir;
        fx = fmaf(wr, dx, fx); fy = fmaf(wr, dy, fy); fz = fmaf(wr, dz, fz);
    }
    fx = warp_sum(fx); fy = warp_sum(fy); fz = warp_sum(fz);
    if (lane == 0) {   // accumulated (not stored): the angular kernel adds into the same zero-initialised array concurrently
        const float sc = -tab->radialScale;
        atomicAdd(posGrad + 3 * (size_t)orig, sc * fx); atomicAdd(posGrad + 3 * (size_t)orig + 1, sc * fy);
        atomicAdd(posGrad + 3 * (size_t)orig + 2, sc * fz);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Angular backward (CpuANISymmetryFunctions.cpp:265-353).  One warp per centre, lane <-> triple: the triples of the centre are
// enumerated flat in rotation order (see the loop), the forces on the two neighbours of a triple are accumulated in per-warp
// shared-memory arrays (collision-poor by construction) and flushed with one red.add per neighbour component at the end.
// ------------------------------------------------------------------------------------------------------------------
struct TripleForce {
    float ax, ay, az, bx, by, bz;
};

template <int MODE, int NSA, int NSZ, bool TORCHANI>
__device__ __forceinline__ TripleForce
triple_backward(float dax, float day, float daz, float ra, float ira, float fa, float dfa, float dbx, float dby, float dbz, float rb,
                float irb, float fb, float dfb, const float* __restrict__ gp, int m0, int nA, const AngTablesSmem& T,
                const float* __restrict__ fShfA, const float* __restrict__ fCos, const float* __restrict__ fSin, float cosScale,
                float fEta, float fEtaL2, float fZeta, float fScale) {
    const float dot = dax * dbx + day * dby + daz * dbz;
    const float ipr = ira * irb;
    const float c = cosScale * dot * ipr;
    float sn, isn;
    if (TORCHANI) {
        sn = sqrtf(fmaxf(1.0f - c * c, 1e-30f));
        isn = 1.0f / sn;
    } else {
        const float cx = day * dbz - daz * dby, cy = daz * dbx - dax * dbz, cz = dax * dby - day * dbx;
        sn = sqrtf(cx * cx + cy * cy + cz * cz) * ipr;
        isn = 1.0f / sn;
    }
    const float rm = 0.5f * (ra + rb);
    float A = 0.0f, B = 0.0f, C = 0.0f;
    if (MODE == 1) {
        float P[NSZ], Q[NSZ];
#pragma unroll
        for (int z = 0; z < NSZ; z++) {
            const float base = fmaxf(1.0f + c * fCos[z] + sn * fSin[z], 0.0f);
            const float pm1 = ex2a((fZeta - 1.0f) * lg2a(base));
            P[z] = pm1 * base;
            Q[z] = -fZeta * pm1 * (sn * fCos[z] - c * fSin[z]);
        }
#pragma unroll
        for (int a = 0; a < NSA; a++) {
            const float tt = rm - fShfA[a];
            const float E = ex2a(-fEtaL2 * tt * tt);
            const float dE = -fEta * tt * E;
            float Ta = 0.0f, Ua = 0.0f;
#pragma unroll
            for (int z = 0; z < NSZ; z++) {
                const float gv = gp[a * NSZ + z];
                Ta = fmaf(gv, P[z], Ta);
                Ua = fmaf(gv, Q[z], Ua);
            }
            A = fmaf(Ta, E, A); B = fmaf(Ta, dE, B); C = fmaf(Ua, E, C);
        }
        A *= fScale; B *= fScale; C *= fScale;
    } else {
        for (int m = m0; m < nA; m++) {
            const float base = fmaxf(1.0f + c * T.cosT[m] + sn * T.sinT[m], 0.0f);
            const float pm1 = ex2a((T.zeta[m] - 1.0f) * lg2a(base));
            const float Pz = pm1 * base;
            const float Qz = -T.zeta[m] * pm1 * (sn * T.cosT[m] - c * T.sinT[m]);
            const float tt = rm - T.shf[m];
            const float E = ex2a(-T.etaL2[m] * tt * tt);
            const float dE = -T.eta[m] * tt * E;
            const float gv = gp[m] * T.scale[m];
            A = fmaf(gv * Pz, E, A); B = fmaf(gv * Pz, dE, B); C = fmaf(gv * Qz, E, C);
        }
    }
    const float F = fa * fb;
    const float wa = (dfa * fb * A + F * B) * ira;
    const float wb = (fa * dfb * A + F * B) * irb;
    const float kk = -cosScale * isn * ipr * (F * C);
    const float pa = dot * ira * ira, pb = dot * irb * irb;
    TripleForce f;
    f.ax = wa * dax + kk * (dbx - pa * dax); f.ay = wa * day + kk * (dby - pa * day); f.az = wa * daz + kk * (dbz - pa * daz);
    f.bx = wb * dbx + kk * (dax - pb * dbx); f.by = wb * dby + kk * (day - pb * dby); f.bz = wb * dbz + kk * (daz - pb * dbz);
    return f;
}

// ------------------------------------------------------------------------------------------------------------------
// Radial backward, centre-owned (scatter) form, used when one box is sharded over several GPUs: a rank only holds dE/dAEV of ITS
// centres, so the term of pair (i, j) that comes from AEV_i is evaluated by centre i and pushed to both atoms:
//   w_ij = scale * sum_k G[i][s_j][k] * exp(..)(fc' - 2 eta (r - Rs_k) fc),   dE/dx_i -= w_ij delta / r,   dE/dx_j += w_ij delta / r
// (CpuANISymmetryFunctions.cpp:228-263 restricted to the centre's own row).  Summed over the ranks this equals the gather form.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWPB * 32)
ani_radial_bwd_scatter_kernel(int n, const float4* __restrict__ sorted, const int* __restrict__ sortedOrig, const Geom* __restrict__ geom,
                              const AniTables* __restrict__ tab, const int* __restrict__ rowRad, const int* __restrict__ offRad, int capR,
                              const int* __restrict__ rowMap, const float* __restrict__ grad, int stride, float* __restrict__ posGrad) {
    extern __shared__ unsigned char smemRaw[];
    __shared__ Geom g;
    __shared__ float sEtaL2[kAniMaxRadial], sEta[kAniMaxRadial], sShf[kAniMaxRadial];
    const int nR = tab->nRadial, S = tab->nSpecies;
    if (threadIdx.x == 0) g = *geom;
    for (int i = threadIdx.x; i < nR; i += blockDim.x) { sEtaL2[i] = tab->rEtaL2[i]; sEta[i] = tab->rEta[i]; sShf[i] = tab->rShf[i]; }
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = blockIdx.x * kWPB + w;
    if (p >= n) return;
    const int cnt = min(offRad[(size_t)p * (S + 1) + S], capR);
    if (cnt == 0) return;                      // also every centre owned by another rank
    float* sGi = reinterpret_cast<float*>(smemRaw) + (size_t)w * S * nR;
    const float4 ci = sorted[p];
    const int orig = sortedOrig[p];
    {
        const float* gi = grad + (size_t)(rowMap ? rowMap[orig] : orig) * stride;
        for (int i = lane; i < S * nR; i += 32) sGi[i] = gi[i];
    }
    __syncwarp();
    const float rcr = tab->rcr, kf = kPi / rcr, sc = tab->radialScale;
    float fx = 0.0f, fy = 0.0f, fz = 0.0f;
    for (int q = lane; q < cnt; q += 32) {
        const int j = rowRad[(size_t)p * capR + q];
        const float4 cj = sorted[j];
        float dx = __fsub_rn(cj.x, ci.x), dy = __fsub_rn(cj.y, ci.y), dz = __fsub_rn(cj.z, ci.z);
        const float r = sqrtf(min_image_mul(g, dx, dy, dz));
        const float ir = 1.0f / r;
        float sn, cs;
        sincosf(r * kf, &sn, &cs);
        const float fc = 0.5f * cs + 0.5f, dfc = -0.5f * kf * sn;
        const float* gi = sGi + (__float_as_int(cj.w) >> 24) * nR;
        float wsum = 0.0f;
        for (int k = 0; k < nR; k++) {
            const float t = r - sShf[k];
            const float ex = ex2a(-sEtaL2[k] * t * t);
            wsum = fmaf(gi[k], ex * (dfc - 2.0f * sEta[k] * t * fc), wsum);
        }
        const float wr = sc * wsum * ir;
        const float gx = wr * dx, gy = wr * dy, gz = wr * dz;
        fx -= gx; fy -= gy; fz -= gz;
        float* dst = posGrad + 3 * (size_t)sortedOrig[j];
        atomicAdd(dst, gx); atomicAdd(dst + 1, gy); atomicAdd(dst + 2, gz);
    }
    fx = warp_sum(fx); fy = warp_sum(fy); fz = warp_sum(fz);
    if (lane == 0) {
        float* dst = posGrad + 3 * (size_t)orig;
        atomicAdd(dst, fx); atomicAdd(dst + 1, fy); atomicAdd(dst + 2, fz);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Angular backward, fast form for the factorised TorchANI tables (same mathematics and enumeration as ani_angular_bwd_kernel
// below): neighbour data packed as two float4 per neighbour, the centre's gradient row read as float4 (pitch 36), the species
// pair looked up in a shared table, sin(theta) and its reciprocal from one rsqrt.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kBwdPitch = 36;

template <int NSA, int NSZ>
__global__ void __launch_bounds__(kWPB * 32)
ani_angular_bwd_fast_kernel(int n, const float4* __restrict__ sorted, const int* __restrict__ sortedOrig, const Geom* __restrict__ geom,
                            const AniTables* __restrict__ tab, const int* __restrict__ rowAng, const int* __restrict__ offAng, int capA,
                            const int* __restrict__ rowMap, const float* __restrict__ grad, int stride, float* __restrict__ posGrad) {
    static_assert(NSA * NSZ == 32, "32 angular channels");
    extern __shared__ unsigned char smemRaw[];
    __shared__ Geom g;
    __shared__ float fShfA[kAniMaxShf], fCos[kAniMaxShf], fSin[kAniMaxShf];
    __shared__ unsigned char pairTab[kAniMaxSpecies * kAniMaxSpecies];
    const int S = tab->nSpecies, nPairs = tab->nPairs;
    if (threadIdx.x == 0) g = *geom;
    if (threadIdx.x < kAniMaxShf) { fShfA[threadIdx.x] = tab->fShfA[threadIdx.x]; fCos[threadIdx.x] = tab->fCos[threadIdx.x]; fSin[threadIdx.x] = tab->fSin[threadIdx.x]; }
    for (int i = threadIdx.x; i < S * S; i += blockDim.x) pairTab[i] = (unsigned char)pair_index(S, i / S, i % S);
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = blockIdx.x * kWPB + w;
    if (p >= n) return;
    // per warp: float4 sv[capA] {dx, dy, dz, r}, float4 sw[capA] {1/r, fc, dfc, species}, float sf[3][capA] forces, int sorig[capA],
    // float sG[nPairs][36]
    const size_t perWarp = (size_t)12 * capA + (size_t)nPairs * kBwdPitch;
    float* wb = reinterpret_cast<float*>(smemRaw) + (size_t)w * perWarp;
    float4* sv = reinterpret_cast<float4*>(wb);
    float4* sw = reinterpret_cast<float4*>(wb + 4 * capA);
    float *sfx = wb + 8 * capA, *sfy = sfx + capA, *sfz = sfy + capA;
    int* sorig = reinterpret_cast<int*>(sfz + capA);
    float* sG = reinterpret_cast<float*>(sorig + capA);
    const int* off = offAng + (size_t)p * (S + 1);
    const int cnt = min(off[S], capA);
    const int orig = sortedOrig[p];
    if (cnt < 2) return;
    const float4 ci = sorted[p];
    const float kf = kPi / tab->rca;
    for (int q = lane; q < cnt; q += 32) {
        const int j = rowAng[(size_t)p * capA + q];
        const float4 cj = sorted[j];
        float dx = __fsub_rn(cj.x, ci.x), dy = __fsub_rn(cj.y, ci.y), dz = __fsub_rn(cj.z, ci.z);
        const float r = sqrtf(min_image_mul(g, dx, dy, dz));
        float sn, cs;
        sincosf(r * kf, &sn, &cs);
        sv[q] = make_float4(dx, dy, dz, r);
        sw[q] = make_float4(1.0f / r, 0.5f * cs + 0.5f, -0.5f * kf * sn, __int_as_float(__float_as_int(cj.w) >> 24));
        sorig[q] = sortedOrig[j];
        sfx[q] = 0.0f; sfy[q] = 0.0f; sfz[q] = 0.0f;
    }
    {
        const float* gi = grad + (size_t)(rowMap ? rowMap[orig] : orig) * stride;
        const int tot = nPairs * 32;
        for (int i = lane; i < tot; i += 32) sG[(i >> 5) * kBwdPitch + (i & 31)] = gi[i];
    }
    __syncwarp();
    const float cosScale = tab->cosScale, fEta = tab->fEta, fEtaL2 = tab->fEtaL2, fZeta = tab->fZeta, fScale = tab->fScale;
    float cxs = 0.0f, cys = 0.0f, czs = 0.0f;
    const int total = (cnt * (cnt - 1)) >> 1;
    int a = lane % cnt, d1 = lane / cnt;          // q = d1 * cnt + a, partner b = a + d1 + 1 (mod cnt)
    for (int q = lane; q < total; q += 32) {
        int b = a + d1 + 1;
        if (b >= cnt) b -= cnt;
        const float4 va = sv[a], vb = sv[b], wa4 = sw[a], wb4 = sw[b];
        const float* gp = sG + pairTab[__float_as_int(wa4.w) * S + __float_as_int(wb4.w)] * kBwdPitch;
        const float dot = va.x * vb.x + va.y * vb.y + va.z * vb.z;
        const float ipr = wa4.x * wb4.x;
        const float c = cosScale * dot * ipr;
        const float x = fmaxf(fmaf(-c, c, 1.0f), 1e-30f);
        const float isn = rsqrta(x), sn = x * isn;
        const float rm = 0.5f * (va.w + vb.w);
        float P[NSZ], Q[NSZ];
#pragma unroll
        for (int z = 0; z < NSZ; z++) {
            const float base = fmaxf(1.0f + c * fCos[z] + sn * fSin[z], 0.0f);
            const float pm1 = ex2a((fZeta - 1.0f) * lg2a(base));
            P[z] = pm1 * base;
            Q[z] = pm1 * (c * fSin[z] - sn * fCos[z]);        // times zeta below
        }
        float A = 0.0f, Bp = 0.0f, C = 0.0f;
#pragma unroll
        for (int k = 0; k < NSA; k++) {
            const float tt = rm - fShfA[k];
            const float E = ex2a(-fEtaL2 * tt * tt);
            float gv[NSZ];
            if (NSZ == 4) {
                const float4 t4 = *reinterpret_cast<const float4*>(gp + k * NSZ);
                gv[0] = t4.x; gv[1] = t4.y; gv[2] = t4.z; gv[3] = t4.w;
            } else {
#pragma unroll
                for (int z = 0; z < NSZ; z += 4) {
                    const float4 t4 = *reinterpret_cast<const float4*>(gp + k * NSZ + z);
                    gv[z] = t4.x; gv[z + 1] = t4.y; gv[z + 2] = t4.z; gv[z + 3] = t4.w;
                }
            }
            float Ta = 0.0f, Ua = 0.0f;
#pragma unroll
            for (int z = 0; z < NSZ; z++) { Ta = fmaf(gv[z], P[z], Ta); Ua = fmaf(gv[z], Q[z], Ua); }
            const float TE = Ta * E;
            A += TE; Bp = fmaf(TE, tt, Bp); C = fmaf(Ua, E, C);
        }
        A *= fScale; const float B = -fEta * fScale * Bp; C *= fZeta * fScale;
        const float F = wa4.y * wb4.y;
        const float FB = F * B;
        const float wa = (wa4.z * wb4.y * A + FB) * wa4.x;
        const float wbb = (wa4.y * wb4.z * A + FB) * wb4.x;
        const float kk = -cosScale * isn * ipr * (F * C);
        const float pa = dot * wa4.x * wa4.x, pb = dot * wb4.x * wb4.x;
        const float fax = wa * va.x + kk * (vb.x - pa * va.x), fay = wa * va.y + kk * (vb.y - pa * va.y), faz = wa * va.z + kk * (vb.z - pa * va.z);
        const float fbx = wbb * vb.x + kk * (va.x - pb * vb.x), fby = wbb * vb.y + kk * (va.y - pb * vb.y), fbz = wbb * vb.z + kk * (va.z - pb * vb.z);
        atomicAdd(&sfx[a], fax); atomicAdd(&sfy[a], fay); atomicAdd(&sfz[a], faz);
        atomicAdd(&sfx[b], fbx); atomicAdd(&sfy[b], fby); atomicAdd(&sfz[b], fbz);
        cxs -= fax + fbx; cys -= fay + fby; czs -= faz + fbz;
        a += 32;
        while (a >= cnt) { a -= cnt; d1++; }
    }
    __syncwarp();
    for (int q = lane; q < cnt; q += 32) {
        float* dst = posGrad + 3 * (size_t)sorig[q];
        atomicAdd(dst, sfx[q]); atomicAdd(dst + 1, sfy[q]); atomicAdd(dst + 2, sfz[q]);
    }
    cxs = warp_sum(cxs); cys = warp_sum(cys); czs = warp_sum(czs);
    if (lane == 0) {
        float* dst = posGrad + 3 * (size_t)orig;
        atomicAdd(dst, cxs); atomicAdd(dst + 1, cys); atomicAdd(dst + 2, czs);
    }
}

template <int MODE, int NSA, int NSZ, bool TORCHANI>
__global__ void __launch_bounds__(kWPB * 32)
ani_angular_bwd_kernel(int n, const float4* __restrict__ sorted, const int* __restrict__ sortedOrig, const Geom* __restrict__ geom,
                       const AniTables* __restrict__ tab, const int* __restrict__ rowAng, const int* __restrict__ offAng, int capA,
                       const int* __restrict__ rowMap, const float* __restrict__ grad, int stride, float* __restrict__ posGrad,
                       int gPitch) {
    extern __shared__ unsigned char smemRaw[];
    __shared__ Geom g;
    __shared__ AngTablesSmem T;
    __shared__ float fShfA[kAniMaxShf], fCos[kAniMaxShf], fSin[kAniMaxShf];
    if (threadIdx.x == 0) g = *geom;
    load_ang_tables(T, tab);
    if (threadIdx.x < kAniMaxShf) { fShfA[threadIdx.x] = tab->fShfA[threadIdx.x]; fCos[threadIdx.x] = tab->fCos[threadIdx.x]; fSin[threadIdx.x] = tab->fSin[threadIdx.x]; }
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = blockIdx.x * kWPB + w;
    if (p >= n) return;
    const int S = tab->nSpecies, nA = tab->nAngular, nPairs = tab->nPairs;
    const size_t perWarp = (size_t)12 * capA + (size_t)nPairs * gPitch;
    float* sdx = reinterpret_cast<float*>(smemRaw) + (size_t)w * perWarp;
    float *sdy = sdx + capA, *sdz = sdy + capA, *sr = sdz + capA, *sir = sr + capA, *sfc = sir + capA, *sdfc = sfc + capA;
    float *sfx = sdfc + capA, *sfy = sfx + capA, *sfz = sfy + capA;   // force accumulators of the neighbours
    int* ssp = reinterpret_cast<int*>(sfz + capA);
    int* sorig = ssp + capA;
    float* sG = reinterpret_cast<float*>(sorig + capA);
    const int* off = offAng + (size_t)p * (S + 1);
    const int cnt = min(off[S], capA);
    const int orig = sortedOrig[p];
    if (cnt < 2) return;   // no triples -> no contribution (positionGrad was initialised by the radial kernel)
    const float4 ci = sorted[p];
    const float kf = kPi / tab->rca;
    for (int q = lane; q < cnt; q += 32) {
        const int j = rowAng[(size_t)p * capA + q];
        const float4 cj = sorted[j];
        float dx = __fsub_rn(cj.x, ci.x), dy = __fsub_rn(cj.y, ci.y), dz = __fsub_rn(cj.z, ci.z);
        const float r = sqrtf(min_image_mul(g, dx, dy, dz));
        float sn, cs;
        sincosf(r * kf, &sn, &cs);
        sdx[q] = dx; sdy[q] = dy; sdz[q] = dz; sr[q] = r; sir[q] = 1.0f / r;
        sfc[q] = 0.5f * cs + 0.5f; sdfc[q] = -0.5f * kf * sn;
        ssp[q] = __float_as_int(cj.w) >> 24;
        sorig[q] = sortedOrig[j];
        sfx[q] = 0.0f; sfy[q] = 0.0f; sfz[q] = 0.0f;
    }
    {   // stage the centre's gradient row [nPairs][nA] with a pitch that spreads species pairs over banks
        const float* gi = grad + (size_t)(rowMap ? rowMap[orig] : orig) * stride;
        const int tot = nPairs * nA;
        for (int i = lane; i < tot; i += 32) sG[(i / nA) * gPitch + (i % nA)] = gi[i];
    }
    __syncwarp();
    const float cosScale = tab->cosScale, fEta = tab->fEta, fEtaL2 = tab->fEtaL2, fZeta = tab->fZeta, fScale = tab->fScale;
    float cxs = 0.0f, cys = 0.0f, czs = 0.0f;   // force on the centre (negated sum)
    auto eval = [&](int a, int b) {
        const float* gp = sG + pair_index(S, ssp[a], ssp[b]) * gPitch;
        return triple_backward<MODE, NSA, NSZ, TORCHANI>(sdx[a], sdy[a], sdz[a], sr[a], sir[a], sfc[a], sdfc[a], sdx[b], sdy[b], sdz[b],
                                                         sr[b], sir[b], sfc[b], sdfc[b], gp, 0, nA, T, fShfA, fCos, fSin, cosScale, fEta,
                                                         fEtaL2, fZeta, fScale);
    };
    // Flat enumeration of the cnt (cnt - 1) / 2 triples in rotation order: entry q is the pair (a, a + d mod cnt) with d = q / cnt + 1,
    // a = q mod cnt (the antipodal step of an even row lists only a < cnt / 2, which is exactly where q runs out).  Lanes of one
    // iteration hold consecutive a (all distinct, at most two values of d), so the shared-memory force accumulators see at most
    // two-way address collisions; every iteration has 32 busy lanes whatever the row length.
    {
        const int total = (cnt * (cnt - 1)) >> 1;
        const float invCnt = 1.0f / (float)cnt;
        for (int q0 = 0; q0 < total; q0 += 32) {
            const int q = q0 + lane;
            if (q < total) {
                const int d1 = (int)(((float)q + 0.5f) * invCnt);
                const int a = q - d1 * cnt;
                int b = a + d1 + 1;
                if (b >= cnt) b -= cnt;
                const TripleForce f = eval(a, b);
                atomicAdd(&sfx[a], f.ax); atomicAdd(&sfy[a], f.ay); atomicAdd(&sfz[a], f.az);
                atomicAdd(&sfx[b], f.bx); atomicAdd(&sfy[b], f.by); atomicAdd(&sfz[b], f.bz);
                cxs -= f.ax + f.bx; cys -= f.ay + f.by; czs -= f.az + f.bz;
            }
        }
        __syncwarp();
        for (int a = lane; a < cnt; a += 32) {
            float* dst = posGrad + 3 * (size_t)sorig[a];
            atomicAdd(dst, sfx[a]); atomicAdd(dst + 1, sfy[a]); atomicAdd(dst + 2, sfz[a]);
        }
    }
    cxs = warp_sum(cxs); cys = warp_sum(cys); czs = warp_sum(czs);
    if (lane == 0) {
        float* dst = posGrad + 3 * (size_t)orig;
        atomicAdd(dst, cxs); atomicAdd(dst + 1, cys); atomicAdd(dst + 2, czs);
    }
}

__global__ void count_kernel_triples(int n, int S, const int* __restrict__ offRad, const int* __restrict__ offAng, int capR, int capA,
                                     unsigned long long* __restrict__ counters) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long tr = 0, pr = 0;
    if (p < n) {
        const long long na = min(offAng[(size_t)p * (S + 1) + S], capA);
        tr = (unsigned long long)(na * (na - 1) / 2);
        pr = (unsigned long long)min(offRad[(size_t)p * (S + 1) + S], capR);
    }
    for (int o = 16; o > 0; o >>= 1) {
        tr += __shfl_xor_sync(kFull, tr, o);
        pr += __shfl_xor_sync(kFull, pr, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&counters[0], tr); atomicAdd(&counters[1], pr); }
}

// Verlet skin, step 1: does any atom sit further than skin / 2 from where it was when the candidate rows were built (or has the
// box changed, or is this the first call)?  The answer stays on the device: the kernels behind it read the flag.
__global__ void skin_check_kernel(int n, const float* __restrict__ pos, const float* __restrict__ box, const float* __restrict__ refPos,
                                  const float* __restrict__ refBox, float lim2, int force, int* __restrict__ rebuild) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && force) *rebuild = 1;
    if (i < 9 && box != nullptr && box[i] != refBox[i]) *rebuild = 1;
    if (i >= n) return;
    const float dx = pos[3 * (size_t)i] - refPos[3 * (size_t)i], dy = pos[3 * (size_t)i + 1] - refPos[3 * (size_t)i + 1],
                dz = pos[3 * (size_t)i + 2] - refPos[3 * (size_t)i + 2];
    if (dx * dx + dy * dy + dz * dz > lim2) *rebuild = 1;   // benign race: every writer stores the same value
}
// step 2 (after the conditional cell-list build): a rebuild step records the reference positions / box; a reuse step refreshes the
// coordinates of the sorted copy (the atoms keep the sorted slots of the last build) and counts itself
__global__ void skin_refresh_kernel(int n, const float* __restrict__ pos, const float* __restrict__ box, const int* __restrict__ sortedOrig,
                                    float4* __restrict__ sorted, float* __restrict__ refPos, float* __restrict__ refBox,
                                    const int* __restrict__ rebuild, unsigned long long* __restrict__ stats) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const bool rb = *rebuild != 0;
    if (p == 0) atomicAdd(&stats[rb ? 0 : 1], 1ull);
    if (rb) {
        if (p < 9 && box != nullptr) refBox[p] = box[p];
        if (p < n) { refPos[3 * (size_t)p] = pos[3 * (size_t)p]; refPos[3 * (size_t)p + 1] = pos[3 * (size_t)p + 1]; refPos[3 * (size_t)p + 2] = pos[3 * (size_t)p + 2]; }
    } else if (p < n) {
        const int i = sortedOrig[p];
        float4 v = sorted[p];
        v.x = pos[3 * (size_t)i]; v.y = pos[3 * (size_t)i + 1]; v.z = pos[3 * (size_t)i + 2];
        sorted[p] = v;
    }
}

template <typename K>
void set_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) NNP_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------------

AniAev::AniAev(int numAtoms, int numSpecies, float rcr, float rca, const int* atomSpecies, int nRadial, const float* radialFn,
               int nAngular, const float* angularFn, bool torchani, int maxRadialNeighbors, int maxAngularNeighbors)
    : n_(numAtoms) {
    NNP_REQUIRE(numAtoms >= 0 && numAtoms < (1 << 24), "numAtoms must be in [0, 2^24)");
    NNP_REQUIRE(numSpecies >= 1 && numSpecies <= kAniMaxSpecies, "numSpecies must be in [1, 32]");
    NNP_REQUIRE(nRadial >= 0 && nRadial <= kAniMaxRadial, "at most 64 radial functions are supported");
    NNP_REQUIRE(nAngular >= 0 && nAngular <= kAniMaxAngular, "at most 128 angular functions are supported");
    for (int i = 0; i < numAtoms; i++)
        NNP_REQUIRE(atomSpecies[i] >= 0 && atomSpecies[i] < numSpecies, "atomSpecies entries must be in [0, numSpecies)");
    capR_ = maxRadialNeighbors > 0 ? maxRadialNeighbors : 256;
    capA_ = maxAngularNeighbors > 0 ? maxAngularNeighbors : 64;   // 3.5 A at liquid density: 17 on average, 24 at most; shared-memory rows scale with it
    capR_ = (capR_ + 31) / 32 * 32;
    capA_ = (capA_ + 31) / 32 * 32;
    AniTables& t = tabHost_;
    std::memset(&t, 0, sizeof(t));
    t.nSpecies = numSpecies; t.nRadial = nRadial; t.nAngular = nAngular; t.nPairs = numSpecies * (numSpecies + 1) / 2;
    t.rcr = rcr; t.rca = rca; t.rcr2 = rcr * rcr; t.rca2 = rca * rca;
    t.torchani = torchani ? 1 : 0;
    t.radialScale = torchani ? 0.25f : 1.0f;
    t.cosScale = torchani ? 0.95f : 1.0f;
    t.invRcr = 1.0f / rcr; t.invRca = 1.0f / rca; t.sqrtCosScale = std::sqrt(t.cosScale);
    const double log2e = 1.4426950408889634;
    for (int k = 0; k < nRadial; k++) {
        t.rEta[k] = radialFn[2 * k]; t.rShf[k] = radialFn[2 * k + 1];
        t.rEtaL2[k] = (float)((double)radialFn[2 * k] * log2e);
    }
    for (int m = 0; m < nAngular; m++) {
        const float eta = angularFn[4 * m], rs = angularFn[4 * m + 1], zeta = angularFn[4 * m + 2], th = angularFn[4 * m + 3];
        t.aEta[m] = eta; t.aEtaL2[m] = (float)((double)eta * log2e); t.aShf[m] = rs; t.aZeta[m] = zeta;
        t.aCos[m] = (float)std::cos((double)th); t.aSin[m] = (float)std::sin((double)th);
        t.aScale[m] = (float)std::pow(2.0, 1.0 - (double)zeta);
    }
    // detect the factorised TorchANI layout m = a * nShfZ + z with one EtaA and one Zeta
    t.fast = 0;
    if (nAngular == 32) {
        for (int nz : {4, 8}) {
            const int na = 32 / nz;
            bool ok = true;
            for (int m = 0; m < 32 && ok; m++) {
                const int a = m / nz, z = m % nz;
                ok = angularFn[4 * m] == angularFn[0] && angularFn[4 * m + 2] == angularFn[2] &&
                     angularFn[4 * m + 1] == angularFn[4 * (a * nz) + 1] && angularFn[4 * m + 3] == angularFn[4 * z + 3];
            }
            if (ok) {
                t.fast = 1; t.nShfA = na; t.nShfZ = nz;
                t.fEta = angularFn[0]; t.fEtaL2 = (float)((double)angularFn[0] * log2e); t.fZeta = angularFn[2];
                t.fScale = (float)std::pow(2.0, 1.0 - (double)angularFn[2]);
                for (int a = 0; a < na; a++) t.fShfA[a] = angularFn[4 * (a * nz) + 1];
                for (int z = 0; z < nz; z++) {
                    t.fCos[z] = (float)std::cos((double)angularFn[4 * z + 3]);
                    t.fSin[z] = (float)std::sin((double)angularFn[4 * z + 3]);
                }
                break;
            }
        }
    }
    const size_t na = (size_t)(n_ > 0 ? n_ : 1);
    NNP_CUDA_CHECK(cudaMalloc(&tab_, sizeof(AniTables)));
    NNP_CUDA_CHECK(cudaMemcpy(tab_, &t, sizeof(AniTables), cudaMemcpyHostToDevice));
    NNP_CUDA_CHECK(cudaMalloc(&species_, sizeof(int) * na));
    if (n_ > 0) {
        // the tag every atom carries through the cell list (float4.w of the sorted copy): species << 24 | atom index, so that one
        // gather of a neighbour's coordinates also brings its species and its index (numAtoms < 2^24, numSpecies <= 32)
        std::vector<int> tagged(n_);
        for (int i = 0; i < n_; i++) tagged[i] = (atomSpecies[i] << 24) | i;
        NNP_CUDA_CHECK(cudaMemcpy(species_, tagged.data(), sizeof(int) * n_, cudaMemcpyHostToDevice));
    }
    cells_.init(n_);
    NNP_CUDA_CHECK(cudaMalloc(&rowRad_, sizeof(int) * na * capR_));
    NNP_CUDA_CHECK(cudaMalloc(&rowAng_, sizeof(int) * na * capA_));
    NNP_CUDA_CHECK(cudaMalloc(&offRad_, sizeof(int) * na * (numSpecies + 1)));
    NNP_CUDA_CHECK(cudaMalloc(&offAng_, sizeof(int) * na * (numSpecies + 1)));
    if (angular_v2_supported(t) && 2 * capA_ <= 256 && std::getenv("NNPOPS_ANGULAR_V1") == nullptr) {
        NNP_CUDA_CHECK(cudaMalloc(&geoA_, sizeof(float4) * na * capA_));
        NNP_CUDA_CHECK(cudaMalloc(&geoB_, sizeof(float4) * na * capA_));
        NNP_CUDA_CHECK(cudaMalloc(&segHist_, 2 * kSegBins * sizeof(int)));
        NNP_CUDA_CHECK(cudaMalloc(&segs_, sizeof(int4) * na * t.nPairs));
        NNP_CUDA_CHECK(cudaMalloc(&nSeg_, sizeof(int)));
        NNP_CUDA_CHECK(cudaMemset(nSeg_, 0, sizeof(int)));
    }
    NNP_CUDA_CHECK(cudaMalloc(&gradAcc_, sizeof(float4) * na));
    if (radial_v2_supported(t) && std::getenv("NNPOPS_RADIAL_V1") == nullptr) {
        NNP_CUDA_CHECK(cudaMalloc(&radGeoA_, sizeof(float4) * na * capR_));
        NNP_CUDA_CHECK(cudaMalloc(&radGeoB_, sizeof(float4) * na * capR_));
    }
    NNP_CUDA_CHECK(cudaMalloc(&flag_, sizeof(int)));
    NNP_CUDA_CHECK(cudaMemset(flag_, 0, sizeof(int)));
    NNP_CUDA_CHECK(cudaMallocHost(&flagHost_, sizeof(int)));
    *flagHost_ = 0;
    NNP_CUDA_CHECK(cudaMalloc(&counters_, 2 * sizeof(unsigned long long)));
    NNP_CUDA_CHECK(cudaStreamCreateWithFlags(&aux_, cudaStreamNonBlocking));
    NNP_CUDA_CHECK(cudaEventCreateWithFlags(&evFork_, cudaEventDisableTiming));
    NNP_CUDA_CHECK(cudaEventCreateWithFlags(&evJoin_, cudaEventDisableTiming));
}

void AniAev::setSkin(float skin) {
    NNP_REQUIRE(skin >= 0.0f, "the Verlet skin must be >= 0");
    cudaFree(candRow_); cudaFree(candCnt_); cudaFree(skinRefPos_); cudaFree(skinRefBox_); cudaFree(skinRebuild_); cudaFree(skinStats_);
    candRow_ = candCnt_ = skinRebuild_ = nullptr; skinRefPos_ = skinRefBox_ = nullptr; skinStats_ = nullptr;
    skin_ = skin;
    if (skin <= 0.0f) return;
    const size_t na = (size_t)(n_ > 0 ? n_ : 1);
    const float rc = tabHost_.rcr > tabHost_.rca ? tabHost_.rcr : tabHost_.rca;
    const double grow = std::pow((rc + skin) / rc, 3.0);
    capC_ = ((int)std::ceil(capR_ * grow * 1.05) + 31) / 32 * 32;
    NNP_CUDA_CHECK(cudaMalloc(&candRow_, sizeof(int) * na * capC_));
    NNP_CUDA_CHECK(cudaMalloc(&candCnt_, sizeof(int) * na));
    NNP_CUDA_CHECK(cudaMalloc(&skinRefPos_, sizeof(float) * 3 * na));
    NNP_CUDA_CHECK(cudaMalloc(&skinRefBox_, sizeof(float) * 9));
    NNP_CUDA_CHECK(cudaMalloc(&skinRebuild_, sizeof(int)));
    NNP_CUDA_CHECK(cudaMalloc(&skinStats_, 2 * sizeof(unsigned long long)));
    // reference positions start at 3.4e38 (bytes 0x7f): the first call -- eager or a graph replay -- sees every atom "moved" and builds
    NNP_CUDA_CHECK(cudaMemset(skinRefPos_, 0x7f, sizeof(float) * 3 * na));
    NNP_CUDA_CHECK(cudaMemset(skinRefBox_, 0, sizeof(float) * 9));
    NNP_CUDA_CHECK(cudaMemset(skinStats_, 0, 2 * sizeof(unsigned long long)));
}

void AniAev::skinStats(unsigned long long* rebuilds, unsigned long long* reuses) {
    unsigned long long h[2] = {0, 0};
    if (skinStats_) {
        NNP_CUDA_CHECK(cudaDeviceSynchronize());
        NNP_CUDA_CHECK(cudaMemcpy(h, skinStats_, sizeof(h), cudaMemcpyDeviceToHost));
    }
    if (rebuilds) *rebuilds = h[0];
    if (reuses) *reuses = h[1];
}

AniAev::~AniAev() {
    cudaFree(candRow_); cudaFree(candCnt_); cudaFree(skinRefPos_); cudaFree(skinRefBox_); cudaFree(skinRebuild_); cudaFree(skinStats_);
    cudaFree(tab_); cudaFree(species_); cudaFree(rowRad_); cudaFree(rowAng_); cudaFree(offRad_); cudaFree(offAng_);
    cudaFree(flag_); cudaFree(counters_);
    cudaFree(radGeoA_); cudaFree(radGeoB_); cudaFree(gradAcc_);
    cudaFree(geoA_); cudaFree(geoB_); cudaFree(segHist_); cudaFree(segs_); cudaFree(nSeg_);
    if (flagHost_) cudaFreeHost(flagHost_);
    if (aux_) cudaStreamDestroy(aux_);
    if (evFork_) cudaEventDestroy(evFork_);
    if (evJoin_) cudaEventDestroy(evJoin_);
    cells_.release();
}

#define ANI_DISPATCH(KERNEL, ...)                                                                                      \
    do {                                                                                                               \
        const bool ta = tabHost_.torchani != 0;                                                                        \
        if (tabHost_.fast && tabHost_.nShfA == 8 && tabHost_.nShfZ == 4) {                                             \
            if (ta) { auto k = KERNEL<1, 8, 4, true>; set_smem(k, smem); k<<<grid, kWPB * 32, smem, stream>>>(__VA_ARGS__); } \
            else    { auto k = KERNEL<1, 8, 4, false>; set_smem(k, smem); k<<<grid, kWPB * 32, smem, stream>>>(__VA_ARGS__); } \
        } else if (tabHost_.fast && tabHost_.nShfA == 4 && tabHost_.nShfZ == 8) {                                      \
            if (ta) { auto k = KERNEL<1, 4, 8, true>; set_smem(k, smem); k<<<grid, kWPB * 32, smem, stream>>>(__VA_ARGS__); } \
            else    { auto k = KERNEL<1, 4, 8, false>; set_smem(k, smem); k<<<grid, kWPB * 32, smem, stream>>>(__VA_ARGS__); } \
        } else {                                                                                                       \
            if (ta) { auto k = KERNEL<0, 1, 1, true>; set_smem(k, smem); k<<<grid, kWPB * 32, smem, stream>>>(__VA_ARGS__); } \
            else    { auto k = KERNEL<0, 1, 1, false>; set_smem(k, smem); k<<<grid, kWPB * 32, smem, stream>>>(__VA_ARGS__); } \
        }                                                                                                              \
    } while (0)

bool AniAev::useV2(const float* angular, int angularStride, const __half* splitHi, const __half* splitLo) const {
    if (geoA_ == nullptr || tabHost_.nAngular == 0) return false;
    // 16-byte stores of 8 channels: the angular block of every output row must start on a multiple of 8 elements
    const bool aligned = ((angularStride | radialWidth()) & 7) == 0 &&
                         ((reinterpret_cast<uintptr_t>(angular) | reinterpret_cast<uintptr_t>(splitHi) | reinterpret_cast<uintptr_t>(splitLo)) & 15) == 0;
    return aligned;
}

void AniAev::forward(const float* positions, const float* box, float* radial, int radialStride, float* angular, int angularStride,
                     cudaStream_t stream, cudaEvent_t* ev, __half* splitHi, __half* splitLo) {
    if (n_ == 0) return;
    // split output: one [n][radialStride] matrix pair, radial block first (radial/angular then only carry the column offsets)
    const AevOut radialOut = {radial, splitHi, splitLo};
    const AevOut angularOut = {angular, splitHi ? splitHi + radialWidth() : nullptr, splitLo ? splitLo + radialWidth() : nullptr};
    const float cut = tabHost_.rcr > tabHost_.rca ? tabHost_.rcr : tabHost_.rca;
    const int grid = (n_ + kWPB - 1) / kWPB;
    if (skin_ > 0.0f) {
        // Verlet skin: candidate rows within cut + skin are kept while no atom has moved more than skin / 2 since they were built.
        // Whether this step rebuilds is decided ON THE DEVICE (skinRebuild_): the cell-list kernels return at once on a reuse step,
        // the row kernel takes its candidates from the kept rows instead of the cells.  No host round trip, CUDA-graph friendly.
        const int tb = 256, nb = (std::max(n_, 9) + tb - 1) / tb;
        NNP_CUDA_CHECK(cudaMemsetAsync(skinRebuild_, 0, sizeof(int), stream));
        skin_check_kernel<<<nb, tb, 0, stream>>>(n_, positions, box, skinRefPos_, skinRefBox_, 0.25f * skin_ * skin_, 0, skinRebuild_);
        cells_.build<float>(positions, box, species_, (cut + skin_) / kCellReach, stream, skinRebuild_);
        skin_refresh_kernel<<<nb, tb, 0, stream>>>(n_, positions, box, cells_.sortedOrig, cells_.sorted, skinRefPos_, skinRefBox_, skinRebuild_, skinStats_);
        count_launch(2);
    } else {
        cells_.build<float>(positions, box, species_, cut / kCellReach, stream);
    }
    // second-generation angular kernels (ani_angular_v2.cu): geometry rows of the angular neighbours, size-sorted (centre, species
    // pair) segments, TMA-staged forward and backward kernels
    const bool v2 = useV2(angular, angularStride, splitHi, splitLo);
    lastForwardV2_ = v2;
    // radial path, second generation: pair geometry (kept for the backward kernel) and the radial AEV in one kernel
    const bool rv2 = radGeoA_ != nullptr && (radialStride & 1) == 0 &&
                     ((reinterpret_cast<uintptr_t>(radial) | reinterpret_cast<uintptr_t>(splitHi) | reinterpret_cast<uintptr_t>(splitLo)) & 7) == 0;
    lastForwardRadV2_ = rv2;
    {
        const size_t smem = (size_t)kWPB * capR_ * sizeof(uint32_t) + (size_t)kWPB * 128 * sizeof(int);
        const float sc = cut + skin_;
        if (skin_ > 0.0f) {
            set_smem(ani_rows_kernel<true>, smem);
            ani_rows_kernel<true><<<grid, kWPB * 32, smem, stream>>>(n_, cells_.sorted, cells_.sortedCell, cells_.geom, cells_.cellStart, tab_,
                                                                   capR_, capA_, rowRad_, rowAng_, offRad_, offAng_, flag_, cells_.sortedOrig, owned_,
                                                                   skinRebuild_, sc * sc, capC_, candRow_, candCnt_);
        } else {
            set_smem(ani_rows_kernel<false>, smem);
            ani_rows_kernel<false><<<grid, kWPB * 32, smem, stream>>>(n_, cells_.sorted, cells_.sortedCell, cells_.geom, cells_.cellStart, tab_,
                                                                    capR_, capA_, rowRad_, rowAng_, offRad_, offAng_, flag_, cells_.sortedOrig, owned_,
                                                                    nullptr, 0.0f, 0, nullptr, nullptr);
        }
        count_launch();
    }
    if (ev) cudaEventRecord(ev[0], stream);
    // The radial and angular kernels are independent (disjoint outputs) and individually latency-bound: fork the radial kernel
    // onto the auxiliary stream so the two run concurrently, join before returning.
    const bool fork = tabHost_.nRadial > 0 && tabHost_.nAngular > 0;
    cudaStream_t rs = fork ? aux_ : stream;
    if (fork) {
        NNP_CUDA_CHECK(cudaEventRecord(evFork_, stream));
        NNP_CUDA_CHECK(cudaStreamWaitEvent(aux_, evFork_, 0));
    }
    // 4 bytes to pinned memory behind the row kernel (on the side stream when there is one, off the critical path): overflowPoll()
    // then sees a truncated row without anybody blocking
    NNP_CUDA_CHECK(cudaMemcpyAsync(flagHost_, flag_, sizeof(int), cudaMemcpyDeviceToHost, rs));
    if (tabHost_.nRadial > 0 && rv2) {
        const AevOutPtr o = {radialOut.f32, radialOut.hi, radialOut.lo};
        radial_v2_forward(n_, tabHost_, tab_, cells_.sorted, cells_.sortedOrig, cells_.geom, rowRad_, offRad_, capR_, radGeoA_, radGeoB_, rowMap_, o,
                          radialStride, rs);
    } else if (tabHost_.nRadial > 0) {
        const size_t smem = (size_t)kWPB * 2 * capR_ * sizeof(float);
        set_smem(ani_radial_fwd_kernel, smem);
        ani_radial_fwd_kernel<<<grid, kWPB * 32, smem, rs>>>(n_, cells_.sorted, cells_.sortedOrig, cells_.geom, tab_, rowRad_, offRad_,
                                                             capR_, rowMap_, radialOut, radialStride);
        count_launch();
    }
    if (fork) NNP_CUDA_CHECK(cudaEventRecord(evJoin_, aux_));
    if (ev) cudaEventRecord(ev[1], stream);
    if (tabHost_.nAngular > 0 && v2) {
        const AevOutPtr o = {angularOut.f32, angularOut.hi, angularOut.lo};
        angular_v2_geometry(n_, tabHost_, tab_, cells_.sorted, cells_.sortedOrig, cells_.geom, rowAng_, offAng_, capA_, geoA_, geoB_, stream);
        angular_v2_build_segments(n_, tabHost_, offAng_, segHist_, segHist_ + kSegBins, segs_, nSeg_, cells_.sortedOrig, rowMap_, o, angularStride, stream);
        angular_v2_forward(n_, tabHost_, tab_, offAng_, capA_, geoA_, geoB_, segs_, nSeg_, cells_.sortedOrig, rowMap_, o, angularStride, stream);
    } else if (tabHost_.nAngular > 0) {
        static const int groupLanes = std::getenv("NNPOPS_ANGULAR_GROUP") ? std::atoi(std::getenv("NNPOPS_ANGULAR_GROUP")) : 8;
        const bool aligned = ((angularStride | radialWidth()) & 3) == 0 &&
                             ((reinterpret_cast<uintptr_t>(angular) | reinterpret_cast<uintptr_t>(splitHi) | reinterpret_cast<uintptr_t>(splitLo)) & 15) == 0;
        if (tabHost_.fast && tabHost_.nShfA == 8 && tabHost_.nShfZ == 4 && tabHost_.torchani && tabHost_.nPairs <= 32 && aligned &&
            (groupLanes == 8 || groupLanes == 16)) {
            const int cpw = 32 / groupLanes;
            const size_t smem = (size_t)kWPB * cpw * ((size_t)6 * capA_ + (kAniMaxSpecies + 1 + 3) / 4 * 4) * sizeof(float);
            const int gridG = (n_ + kWPB * cpw - 1) / (kWPB * cpw);
            if (groupLanes == 8) {
                auto k = ani_angular_fwd_grouped_kernel<8, 8, 4, true>;
                set_smem(k, smem);
                k<<<gridG, kWPB * 32, smem, stream>>>(n_, cells_.sorted, cells_.sortedOrig, cells_.geom, tab_, rowAng_, offAng_, capA_, rowMap_,
                                                     angularOut, angularStride);
            } else {
                auto k = ani_angular_fwd_grouped_kernel<16, 8, 4, true>;
                set_smem(k, smem);
                k<<<gridG, kWPB * 32, smem, stream>>>(n_, cells_.sorted, cells_.sortedOrig, cells_.geom, tab_, rowAng_, offAng_, capA_, rowMap_,
                                                     angularOut, angularStride);
            }
        } else {
            const size_t smem = (size_t)kWPB * 6 * capA_ * sizeof(float);
            ANI_DISPATCH(ani_angular_fwd_kernel, n_, cells_.sorted, cells_.sortedOrig, cells_.geom, tab_, rowAng_, offAng_, capA_, rowMap_,
                         angularOut, angularStride);
        }
        count_launch();
    }
    if (fork) NNP_CUDA_CHECK(cudaStreamWaitEvent(stream, evJoin_, 0));
    NNP_CUDA_CHECK(cudaGetLastError());
    haveForward_ = true;
}

void AniAev::backward(const float* radialGrad, int radialStride, const float* angularGrad, int angularStride, float* positionGrad,
                      cudaStream_t stream, cudaEvent_t* ev) {
    if (n_ == 0) return;
    NNP_REQUIRE(haveForward_, "backward() called before forward()");
    const int grid = (n_ + kWPB - 1) / kWPB;
    const bool angV2 = tabHost_.nAngular > 0 && lastForwardV2_ && (angularStride & 3) == 0 && (reinterpret_cast<uintptr_t>(angularGrad) & 15) == 0;
    const bool radV2 = tabHost_.nRadial > 0 && lastForwardRadV2_ && (radialStride & 3) == 0 && (reinterpret_cast<uintptr_t>(radialGrad) & 15) == 0;
    // both second-generation kernels: forces go into a padded [n][4] buffer by 16-byte vector reductions, copied out at the end
    const bool padded = (angV2 || tabHost_.nAngular == 0) && (radV2 || tabHost_.nRadial == 0) && gradAcc_ != nullptr;
    float* acc = padded ? reinterpret_cast<float*>(gradAcc_) : positionGrad;
    NNP_CUDA_CHECK(cudaMemsetAsync(acc, 0, sizeof(float) * (padded ? 4 : 3) * n_, stream));
    const bool fork = tabHost_.nRadial > 0 && tabHost_.nAngular > 0;
    cudaStream_t rs = fork ? aux_ : stream;
    if (fork) {
        NNP_CUDA_CHECK(cudaEventRecord(evFork_, stream));
        NNP_CUDA_CHECK(cudaStreamWaitEvent(aux_, evFork_, 0));
    }
    if (radV2) {
        radial_v2_backward(n_, tabHost_, tab_, offRad_, capR_, radGeoA_, radGeoB_, cells_.sortedOrig, rowMap_, radialGrad, radialStride, acc, padded, rs);
    } else if (tabHost_.nRadial > 0) {
        const size_t smem = (size_t)kWPB * tabHost_.nSpecies * tabHost_.nRadial * sizeof(float);
        // the centre-owned (scatter) form is required when the box is sharded and is also the faster one on a single GPU (no
        // gathers of neighbour gradient rows: 0.224 vs 0.255 ms for the backward stage); NNPOPS_RADIAL_GATHER selects the gather form
        static const bool gather = std::getenv("NNPOPS_RADIAL_GATHER") != nullptr;
        auto kernel = (owned_ != nullptr || !gather) ? ani_radial_bwd_scatter_kernel : ani_radial_bwd_kernel;
        set_smem(kernel, smem);
        kernel<<<grid, kWPB * 32, smem, rs>>>(n_, cells_.sorted, cells_.sortedOrig, cells_.geom, tab_, rowRad_, offRad_, capR_, rowMap_,
                                              radialGrad, radialStride, positionGrad);
        count_launch();
    }
    if (fork) NNP_CUDA_CHECK(cudaEventRecord(evJoin_, aux_));
    if (ev) cudaEventRecord(ev[0], stream);
    if (angV2) {
        angular_v2_backward(n_, tabHost_, tab_, offAng_, capA_, geoA_, geoB_, cells_.sortedOrig, rowMap_, angularGrad, angularStride, acc, padded, stream);
    } else if (tabHost_.nAngular > 0) {
        const int gPitch = tabHost_.nAngular + 1;
        const size_t smem = (size_t)kWPB * ((size_t)12 * capA_ + (size_t)tabHost_.nPairs * gPitch) * sizeof(float);
        NNP_REQUIRE(smem <= 200 * 1024, "angular gradient row does not fit in shared memory (numSpecies^2 * numAngular too large)");
        static const bool slowBwd = std::getenv("NNPOPS_ANGULAR_BWD_GENERIC") != nullptr;
        if (!slowBwd && tabHost_.fast && tabHost_.nShfA == 8 && tabHost_.nShfZ == 4 && tabHost_.torchani && tabHost_.nSpecies <= 15 &&
            (capA_ & 3) == 0) {
            const size_t smemF = (size_t)kWPB * ((size_t)12 * capA_ + (size_t)tabHost_.nPairs * kBwdPitch) * sizeof(float);
            NNP_REQUIRE(smemF <= 200 * 1024, "angular gradient row does not fit in shared memory (numSpecies^2 * numAngular too large)");
            auto k = ani_angular_bwd_fast_kernel<8, 4>;
            set_smem(k, smemF);
            k<<<grid, kWPB * 32, smemF, stream>>>(n_, cells_.sorted, cells_.sortedOrig, cells_.geom, tab_, rowAng_, offAng_, capA_, rowMap_,
                                                 angularGrad, angularStride, positionGrad);
        } else {
            ANI_DISPATCH(ani_angular_bwd_kernel, n_, cells_.sorted, cells_.sortedOrig, cells_.geom, tab_, rowAng_, offAng_, capA_, rowMap_,
                         angularGrad, angularStride, positionGrad, gPitch);
        }
        count_launch();
    }
    if (fork) NNP_CUDA_CHECK(cudaStreamWaitEvent(stream, evJoin_, 0));
    if (padded) grad_compact(n_, gradAcc_, positionGrad, stream);
    NNP_CUDA_CHECK(cudaGetLastError());
}

int AniAev::overflowed() {
    int h = 0;
    NNP_CUDA_CHECK(cudaDeviceSynchronize());
    NNP_CUDA_CHECK(cudaMemcpy(&h, flag_, sizeof(int), cudaMemcpyDeviceToHost));
    return h;
}

long long AniAev::countTriples(cudaStream_t stream) {
    if (n_ == 0 || !haveForward_) return 0;
    unsigned long long h[2] = {0, 0};
    NNP_CUDA_CHECK(cudaMemsetAsync(counters_, 0, 2 * sizeof(unsigned long long), stream));
    count_kernel_triples<<<(n_ + 255) / 256, 256, 0, stream>>>(n_, tabHost_.nSpecies, offRad_, offAng_, capR_, capA_, counters_);
    NNP_CUDA_CHECK(cudaMemcpyAsync(h, counters_, sizeof(h), cudaMemcpyDeviceToHost, stream));
    NNP_CUDA_CHECK(cudaStreamSynchronize(stream));
    return (long long)h[0];
}

long long AniAev::countRadialPairs(cudaStream_t stream) {
    if (n_ == 0 || !haveForward_) return 0;
    unsigned long long h[2] = {0, 0};
    NNP_CUDA_CHECK(cudaMemsetAsync(counters_, 0, 2 * sizeof(unsigned long long), stream));
    count_kernel_triples<<<(n_ + 255) / 256, 256, 0, stream>>>(n_, tabHost_.nSpecies, offRad_, offAng_, capR_, capA_, counters_);
    NNP_CUDA_CHECK(cudaMemcpyAsync(h, counters_, sizeof(h), cudaMemcpyDeviceToHost, stream));
    NNP_CUDA_CHECK(cudaStreamSynchronize(stream));
    return (long long)(h[1] / 2);
}

}  // namespace nnpops
