// Angular AEV kernels, second generation (sm_100a).  Behaviour matched: CpuANISymmetryFunctions.cpp:139-194 (forward) and
// :265-353 (backward); reference CUDA counterparts CudaANISymmetryFunctions.cu:242-290 and :473-596.
//
// What changed against the first generation (ani_aev.cu: ani_angular_fwd_grouped_kernel / ani_angular_bwd_fast_kernel):
//
//  * Neighbour geometry is computed ONCE, by the row kernel, and stored next to the index rows (species-grouped order):
//      geoA[p][q] = { sqrt(cs) * delta / r  (3 floats),  r / 2 }          cs = 0.95 (TorchANI) so that  uA . uB = cs * cos(theta)
//      geoB[p][q] = { fc(r), fc'(r), 1 / r, (species << 24) | atom index }
//    The compute kernels pull a centre's rows into shared memory with cp.async.bulk (TMA, mbarrier completion) instead of
//    gathering coordinates and redoing min-image / sqrt / cos per kernel.
//
//  * Forward: the unit of work is a SEGMENT = one (centre, species pair) block of triples.  All non-empty segments of the system are
//    binned by size and laid out largest first (angular_v2_build_segments); a warp takes 8 consecutive segments -- practically equal
//    sizes -- and gives each 4 lanes, so every lane slot of every iteration holds a real triple (the per-centre kernels padded each
//    block to the largest of the centres sharing a warp: 70 % lane fill on liquid water, 51 % on a protein).  lane <-> triple, 32
//    channel accumulators in registers, a 2-level butterfly leaves lane gl with channels [8 gl, 8 gl + 8): one 16-byte store per
//    hi / lo half.  Triples of a same-species segment are enumerated in rotation order (a, a + d mod n), those of a mixed segment
//    as (a, d): both advance by "a += 4, wrap", no division in the loop.
//
//  * Backward: warp per centre, lane <-> triple in flat rotation order, step = min(32, 2 n - 2) triples per iteration.  With that
//    step the lanes of one iteration that share the parity of d hold pairwise different a and pairwise different b (checked
//    exhaustively for n <= 64 in tests/test_abi_and_host.py), so the forces on the two neighbours are accumulated with plain
//    16-byte read-modify-writes into [parity][slot] arrays -- the first generation used 6 shared-memory float atomics per triple,
//    which sm_100 executes as compare-and-swap loops.  The channel contraction runs as G[z] = sum_a g[a][z] E_a first (the radial
//    factor is shared by the 4 angular channels), and the centre's force is minus the sum of its neighbours' at the end.
#include "ani_angular_v2.cuh"
#include <algorithm>
#include <cmath>
#include <map>
#include <mutex>

namespace nnpops {

namespace {

constexpr int kWarpsPerCta = 8;
constexpr int kFwdBudget = 256;    // neighbour entries of shared memory per warp in the forward kernel (>= 2 * capA required)
constexpr int kBwdPitch = 36;      // floats per species-pair block of the staged gradient row (32 + 4: conflict-free float4 reads)

__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2a(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rsqrta(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

typedef unsigned long long u64;   // packed fp32x2 (FADD2 / FMUL2 / FFMA2: one issue slot for two lanes of work)
__device__ __forceinline__ u64 pk2(float x, float y) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y)); return r; }
__device__ __forceinline__ void unpk2(u64 v, float& x, float& y) { asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v)); }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
// 1-D bulk copy global -> shared (TMA): 16-byte aligned addresses, size a multiple of 16
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// accumulate a force on an atom: one 16-byte vector reduction into a padded [n][4] buffer (VEC), or three scalar ones into [n][3]
template <bool VEC>
__device__ __forceinline__ void grad_add(float* __restrict__ dst, size_t atom, float x, float y, float z) {
    if (VEC) {
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * atom), "f"(x), "f"(y), "f"(z), "f"(0.0f) : "memory");
    } else {
        atomicAdd(dst + 3 * atom, x); atomicAdd(dst + 3 * atom + 1, y); atomicAdd(dst + 3 * atom + 2, z);
    }
}

__device__ __forceinline__ int pair_index(int S, int s, int t) {   // CpuANISymmetryFunctions.cpp:39-43, s <= t
    return s * S - (s * (s - 1)) / 2 + (t - s);
}

__device__ __forceinline__ void zero_block32(const AevOutPtr& o, size_t idx) {
    if (o.hi) {
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int i = 0; i < 4; i++) { reinterpret_cast<uint4*>(o.hi + idx)[i] = z; reinterpret_cast<uint4*>(o.lo + idx)[i] = z; }
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) reinterpret_cast<float4*>(o.f32 + idx)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Segment list: counting sort of the non-empty (centre, species pair) blocks by size, largest first.  Two small kernels, one thread
// per centre, 512 centres per CTA.  All counting goes through shared-memory histograms, so a CTA touches each global bin once
// (atomics of many threads on the few populated bins of a global histogram serialise: 44 us for 150 000 segments).
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int seg_bin(int ntrip) { return min((ntrip + kSegGroup - 1) / kSegGroup, kSegBins - 1); }

__global__ void __launch_bounds__(kSegBins)
ani_seg_hist_kernel(int n, int S, const int* __restrict__ offAng, int* __restrict__ hist) {
    __shared__ int local[kSegBins];
    const int tid = threadIdx.x;
    local[tid] = 0;
    __syncthreads();
    const int p = blockIdx.x * blockDim.x + tid;
    if (p < n) {
        const int* off = offAng + (size_t)p * (S + 1);
        for (int s = 0; s < S; s++) {
            const int ns = off[s + 1] - off[s];
            if (ns == 0) continue;
            for (int t = s; t < S; t++) {
                const int ntrip = (s == t) ? (ns * (ns - 1)) / 2 : ns * (off[t + 1] - off[t]);
                if (ntrip > 0) atomicAdd(&local[seg_bin(ntrip)], 1);
            }
        }
    }
    __syncthreads();
    if (local[tid] > 0) atomicAdd(&hist[tid], local[tid]);
}

__global__ void __launch_bounds__(kSegBins)
ani_seg_scatter_kernel(int n, int S, const int* __restrict__ offAng, const int* __restrict__ hist, int* __restrict__ cursor,
                       int4* __restrict__ segs, int* __restrict__ nSeg, const int* __restrict__ sortedOrig, const int* __restrict__ rowMap,
                       AevOutPtr out, int stride) {
    __shared__ int base[kSegBins];       // first slot of this CTA's entries of a bin
    __shared__ int local[kSegBins];      // entries of this CTA per bin, then the running cursor
    __shared__ int warpTot[kSegBins / 32];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    local[tid] = 0;
    {   // thread t owns bin (kSegBins - 1 - t): inclusive scan in descending bin order
        const int v = hist[kSegBins - 1 - tid];
        int x = v;
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(kFull, x, o); if (lane >= o) x += y; }
        if (lane == 31) warpTot[w] = x;
        __syncthreads();
        int add = 0;
        for (int i = 0; i < w; i++) add += warpTot[i];
        base[kSegBins - 1 - tid] = add + x - v;
        if (blockIdx.x == 0 && tid == kSegBins - 1) *nSeg = add + x;
    }
    __syncthreads();
    const int p = blockIdx.x * blockDim.x + tid;
    const int* off = offAng + (size_t)min(p, n - 1) * (S + 1);
    if (p < n) {
        for (int s = 0; s < S; s++) {
            const int ns = off[s + 1] - off[s];
            if (ns == 0) continue;
            for (int t = s; t < S; t++) {
                const int ntrip = (s == t) ? (ns * (ns - 1)) / 2 : ns * (off[t + 1] - off[t]);
                if (ntrip > 0) atomicAdd(&local[seg_bin(ntrip)], 1);
            }
        }
    }
    __syncthreads();
    {   // reserve this CTA's range of every populated bin with ONE global atomic
        const int c = local[tid];
        if (c > 0) base[tid] += atomicAdd(&cursor[tid], c);
        local[tid] = 0;
    }
    __syncthreads();
    if (p >= n) return;
    const int orig = sortedOrig[p];
    const int outRow = rowMap ? rowMap[orig] : orig;
    const size_t orow = (size_t)outRow * stride;
    int pIdx = 0;
    for (int s = 0; s < S; s++) {
        const int ns = off[s + 1] - off[s];
        for (int t = s; t < S; t++, pIdx++) {
            const int nt = off[t + 1] - off[t];
            const int ntrip = (s == t) ? (ns * (ns - 1)) / 2 : ns * nt;
            if (ntrip <= 0) { zero_block32(out, orow + (size_t)pIdx * 32); continue; }
            const int bin = seg_bin(ntrip);
            // everything the forward kernel needs to know about the segment: centre, species pair block, the two sub-ranges of the
            // centre's angular row (start and length, 8 bits each: capA <= 128), output row
            segs[base[bin] + atomicAdd(&local[bin], 1)] =
                make_int4(p, (pIdx << 1) | (s == t ? 1 : 0), off[s] | (ns << 8) | (off[t] << 16) | (nt << 24), outRow);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Forward
// ------------------------------------------------------------------------------------------------------------------
#ifndef NNP_FWD_MINB
#define NNP_FWD_MINB 3
#endif
#ifndef NNP_FWD_UNROLL
#define NNP_FWD_UNROLL 1
#endif
template <int NSA, int NSZ>
__global__ void __launch_bounds__(kWarpsPerCta * 32, NNP_FWD_MINB)
ani_angular_fwd_seg_kernel(const AniTables* __restrict__ tab, const int* __restrict__ offAng, int capA, const float4* __restrict__ geoA,
                           const float4* __restrict__ geoB, const int4* __restrict__ segs, const int* __restrict__ nSegPtr,
                           AevOutPtr out, int stride) {
    static_assert(NSA * NSZ == 32, "32 angular channels");
    constexpr int G = kSegGroup, GPW = 32 / G;
    extern __shared__ __align__(128) unsigned char smemRaw[];
    __shared__ __align__(8) unsigned long long bars[kWarpsPerCta];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, sub = lane / G, gl = lane % G;
    if (lane == 0) mbar_init(smem_u32(&bars[w]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const uint32_t bar = smem_u32(&bars[w]);
    float4* sA = reinterpret_cast<float4*>(smemRaw + (size_t)w * kFwdBudget * 20);
    float* sFc = reinterpret_cast<float*>(sA + kFwdBudget);
    const uint32_t sAu = smem_u32(sA);
    const int S = tab->nSpecies;
    const float fZeta = tab->fZeta, nEtaL2 = -tab->fEtaL2, fScale = tab->fScale;
    u64 nShf2[NSA / 2], cz2[NSZ / 2], sz2[NSZ / 2];   // packed fp32x2 constants: two radial / angular shifts per register pair
#pragma unroll
    for (int a = 0; a < NSA / 2; a++) nShf2[a] = pk2(-tab->fShfA[2 * a], -tab->fShfA[2 * a + 1]);
#pragma unroll
    for (int z = 0; z < NSZ / 2; z++) { cz2[z] = pk2(tab->fCos[2 * z], tab->fCos[2 * z + 1]); sz2[z] = pk2(tab->fSin[2 * z], tab->fSin[2 * z + 1]); }
    const u64 nEta2 = pk2(nEtaL2, nEtaL2), one2 = pk2(1.0f, 1.0f);
    const int nSeg = *nSegPtr;
    const int nChunks = (nSeg + GPW - 1) / GPW;
    uint32_t phase = 0;
    for (int chunk = blockIdx.x * kWarpsPerCta + w; chunk < nChunks; chunk += gridDim.x * kWarpsPerCta) {
        const int segIdx = chunk * GPW + sub;
        const bool valid = segIdx < nSeg;
        int p = 0, bs = 0, ns = 0, bt = 0, nt = 0, ntrip = 0, e = 0, pIdx = 0, outRow = 0;
        bool same = true;
        if (valid) {
            const int4 sg = segs[segIdx];              // one 16-byte load: no dependent look-ups in the offset tables
            p = sg.x; pIdx = sg.y >> 1; same = (sg.y & 1) != 0; outRow = sg.w;
            bs = sg.z & 0xff; ns = (sg.z >> 8) & 0xff; bt = (sg.z >> 16) & 0xff; nt = (sg.z >> 24) & 0xff;
            ntrip = same ? (ns * (ns - 1)) / 2 : ns * nt;
            e = same ? ns : ns + nt;
        }
        // shared-memory offsets of the 8 segments: a scan over the groups (every lane of a group holds the same e)
        int incl = e;
#pragma unroll
        for (int o = G; o < 32; o <<= 1) { const int y = __shfl_up_sync(kFull, incl, o); if (lane >= o) incl += y; }
        const int allEntries = __shfl_sync(kFull, incl, 31);
        int first = 0;
        while (first < GPW) {   // ONE pass unless the 8 segments need more than kFwdBudget entries (then as many groups as fit per pass)
            int run = allEntries, lim = GPW, myBase = incl - e;
            if (allEntries > kFwdBudget) {
                run = 0; myBase = 0;
                for (int g = first; g < GPW; g++) {
                    const int eg = __shfl_sync(kFull, e, g * G);
                    if (run + eg > kFwdBudget) { lim = g; break; }
                    if (g == sub) myBase = run;
                    run += eg;
                }
            }
            if (run == 0) break;                      // only invalid segments left
            const bool act = valid && sub >= first && sub < lim;
            if (lane == 0) mbar_expect_tx(bar, (uint32_t)run * 16u);
            __syncwarp();
            if (act) {
                const size_t rowBase = (size_t)p * capA;
                if (gl == 0) {
                    bulk_g2s(sAu + (uint32_t)myBase * 16u, geoA + rowBase + bs, (uint32_t)ns * 16u, bar);
                    if (!same) bulk_g2s(sAu + (uint32_t)(myBase + ns) * 16u, geoA + rowBase + bt, (uint32_t)nt * 16u, bar);
                }
                for (int i = gl; i < e; i += G) sFc[myBase + i] = geoB[rowBase + (i < ns ? bs + i : bt + i - ns)].x;
            }
            __syncwarp();
            mbar_wait(bar, phase);
            phase ^= 1u;
            const int trip = act ? ntrip : 0;
            const int maxTrip = __reduce_max_sync(kFull, trip);
            // lane gl starts at triple q = gl and advances by G.  Same species: q = d * ns + a  <->  pair (a, a + d + 1 mod ns);
            // mixed: q = d * ns + a  <->  (a of species s, d of species t).
            const int nsl = act ? ns : (1 << 28);
            int a = gl, d = 0;
            while (a >= nsl) { a -= nsl; d++; }
            const int mb = act ? myBase : 0;
            const int tOff = same ? 0 : ns;
            u64 acc2[NSA / 2][NSZ];                 // acc2[k][z] = channels ((2 k) * NSZ + z, (2 k + 1) * NSZ + z)
#pragma unroll
            for (int k = 0; k < NSA / 2; k++)
#pragma unroll
                for (int z = 0; z < NSZ; z++) acc2[k][z] = pk2(0.f, 0.f);
            // one triple: lane-local state (a, d) -> neighbour slots, angular and radial factors, 32 channel updates
            auto triple = [&](int ta, int td, bool v) {
                int x = ta + td + 1;
                x -= (x >= nsl) ? nsl : 0;
                const int ja = v ? mb + ta : 0, jb = v ? mb + (same ? x : tOff + td) : 0;
                const float4 va = sA[ja], vb = sA[jb];
                const float fa = sFc[ja], fb = sFc[jb];
                const float c = fmaf(va.z, vb.z, fmaf(va.y, vb.y, va.x * vb.x));
                const float xx = fmaxf(fmaf(-c, c, 1.0f), 1e-30f);
                const float sn = xx * rsqrta(xx);
                const float rm = va.w + vb.w;
                const float F = v ? fa * fb : 0.0f;
                const u64 c2 = pk2(c, c), sn2 = pk2(sn, sn);
                u64 Pd[NSZ];                            // (P_z, P_z)
#pragma unroll
                for (int z = 0; z < NSZ / 2; z++) {
                    const u64 base2 = ffma2(sn2, sz2[z], ffma2(c2, cz2[z], one2));
                    float b0, b1;
                    unpk2(base2, b0, b1);
                    const float p0 = F * ex2a(fZeta * lg2a(fabsf(b0))), p1 = F * ex2a(fZeta * lg2a(fabsf(b1)));
                    Pd[2 * z] = pk2(p0, p0); Pd[2 * z + 1] = pk2(p1, p1);
                }
                const u64 rm2 = pk2(rm, rm);
#pragma unroll
                for (int k = 0; k < NSA / 2; k++) {
                    const u64 tt2 = fadd2(rm2, nShf2[k]);
                    const u64 ar2 = fmul2(fmul2(tt2, nEta2), tt2);
                    float a0, a1;
                    unpk2(ar2, a0, a1);
                    const u64 E2 = pk2(ex2a(a0), ex2a(a1));
#pragma unroll
                    for (int z = 0; z < NSZ; z++) acc2[k][z] = ffma2(Pd[z], E2, acc2[k][z]);
                }
            };
#if NNP_FWD_UNROLL == 2
            // two triples per lane and trip: two independent MUFU / FMA chains in flight per warp
            for (int q = gl; q < maxTrip; q += 2 * G) {
                int a1 = a + G, d1 = d;
                while (a1 >= nsl) { a1 -= nsl; d1++; }
                triple(a, d, q < trip);
                triple(a1, d1, q + G < trip);
                a = a1 + G; d = d1;
                while (a >= nsl) { a -= nsl; d++; }
            }
#else
            for (int q = gl; q < maxTrip; q += G) {
                triple(a, d, q < trip);
                a += G;
                while (a >= nsl) { a -= nsl; d++; }
            }
#endif
            float acc[32];
#pragma unroll
            for (int k = 0; k < NSA / 2; k++)
#pragma unroll
                for (int z = 0; z < NSZ; z++) unpk2(acc2[k][z], acc[(2 * k) * NSZ + z], acc[(2 * k + 1) * NSZ + z]);
            // in-group transpose-reduce: lane gl ends with channels [8 gl, 8 gl + 8)
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
            if (act) {
                const size_t dst = (size_t)outRow * stride + (size_t)pIdx * 32 + gl * 8;
                if (out.hi) {
                    uint32_t wh[4], wl[4];
#pragma unroll
                    for (int i = 0; i < 8; i += 2) {
                        const float v0 = acc[i] * fScale, v1 = acc[i + 1] * fScale;
                        const __half2 h2 = __floats2half2_rn(v0, v1);
                        const float2 f2 = __half22float2(h2);
                        const __half2 l2 = __floats2half2_rn((v0 - f2.x) * 2048.0f, (v1 - f2.y) * 2048.0f);
                        wh[i / 2] = *reinterpret_cast<const uint32_t*>(&h2);
                        wl[i / 2] = *reinterpret_cast<const uint32_t*>(&l2);
                    }
                    *reinterpret_cast<uint4*>(out.hi + dst) = make_uint4(wh[0], wh[1], wh[2], wh[3]);
                    *reinterpret_cast<uint4*>(out.lo + dst) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
                } else {
                    reinterpret_cast<float4*>(out.f32 + dst)[0] = make_float4(acc[0] * fScale, acc[1] * fScale, acc[2] * fScale, acc[3] * fScale);
                    reinterpret_cast<float4*>(out.f32 + dst)[1] = make_float4(acc[4] * fScale, acc[5] * fScale, acc[6] * fScale, acc[7] * fScale);
                }
            }
            __syncwarp();      // every lane is done with the staged rows before the next bulk copy lands on them
            first = lim;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Backward
// ------------------------------------------------------------------------------------------------------------------
template <int NSA, int NSZ, bool VEC>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 4)
ani_angular_bwd_v2_kernel(int n, const AniTables* __restrict__ tab, const int* __restrict__ offAng, int capA,
                          const float4* __restrict__ geoA, const float4* __restrict__ geoB, const int* __restrict__ sortedOrig,
                          const int* __restrict__ rowMap, const float* __restrict__ grad, int stride, float* __restrict__ posGrad) {
    static_assert(NSZ == 4 && NSA * NSZ == 32, "8 x 4 channels");
    extern __shared__ __align__(128) unsigned char smemRaw[];
    __shared__ __align__(8) unsigned long long bars[kWarpsPerCta];
    __shared__ unsigned char pairTab[kAniMaxSpecies * kAniMaxSpecies];
    const int S = tab->nSpecies, nPairs = tab->nPairs;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) mbar_init(smem_u32(&bars[w]), 1);
    for (int i = threadIdx.x; i < S * S; i += blockDim.x) {
        const int s = i / S, t = i % S;
        pairTab[i] = (unsigned char)pair_index(S, min(s, t), max(s, t));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    // per warp: float4 sA[capA], sB[capA], acc[4][capA] (force accumulators X0, X1 for the first, Y0, Y1 for the second neighbour of a
    // triple, indexed by the parity of d), float sG[nPairs][36]
    const size_t perWarp = (size_t)capA * 96 + (size_t)nPairs * kBwdPitch * 4;
    unsigned char* wbase = smemRaw + (size_t)w * perWarp;
    float4* sA = reinterpret_cast<float4*>(wbase);
    float4* sB = sA + capA;
    float4* sAcc = sB + capA;
    float* sG = reinterpret_cast<float*>(sAcc + 4 * capA);
    const uint32_t bar = smem_u32(&bars[w]);
    const float cs = tab->cosScale;
    const float rcs = 1.0f / cs, sqcs = sqrtf(cs), k1 = 1.0f / sqcs;
    const float nEtaL2 = -tab->fEtaL2, zm1 = tab->fZeta - 1.0f;
    const float kA = k1 * tab->fScale, kB = -k1 * tab->fEta * tab->fScale, kC = sqcs * tab->fZeta * tab->fScale;
    u64 nShf2[NSA / 2], cz2[NSZ / 2], sz2[NSZ / 2];
#pragma unroll
    for (int a = 0; a < NSA / 2; a++) nShf2[a] = pk2(-tab->fShfA[2 * a], -tab->fShfA[2 * a + 1]);
#pragma unroll
    for (int z = 0; z < NSZ / 2; z++) { cz2[z] = pk2(tab->fCos[2 * z], tab->fCos[2 * z + 1]); sz2[z] = pk2(tab->fSin[2 * z], tab->fSin[2 * z + 1]); }
    const u64 nEta2 = pk2(nEtaL2, nEtaL2), one2 = pk2(1.0f, 1.0f);
    uint32_t phase = 0;
    // persistent: a warp walks over centres, so the table set-up above and the CTA launch are paid once per warp
    for (int p = blockIdx.x * kWarpsPerCta + w; p < n; p += gridDim.x * kWarpsPerCta) {
    const int cnt = offAng[(size_t)p * (S + 1) + S];
    if (cnt < 2) continue;
    const int orig = sortedOrig[p];
    const float* gi = grad + (size_t)(rowMap ? rowMap[orig] : orig) * stride;
    if (lane == 0) mbar_expect_tx(bar, (uint32_t)cnt * 32u + (uint32_t)nPairs * 128u);
    __syncwarp();
    if (lane == 0) {
        bulk_g2s(smem_u32(sA), geoA + (size_t)p * capA, (uint32_t)cnt * 16u, bar);
        bulk_g2s(smem_u32(sB), geoB + (size_t)p * capA, (uint32_t)cnt * 16u, bar);
    }
    for (int i = lane; i < nPairs; i += 32) bulk_g2s(smem_u32(sG + i * kBwdPitch), gi + i * 32, 128u, bar);
    for (int i = lane; i < cnt; i += 32) {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        sAcc[i] = z; sAcc[capA + i] = z; sAcc[2 * capA + i] = z; sAcc[3 * capA + i] = z;
    }
    __syncwarp();
    mbar_wait(bar, phase);
    phase ^= 1u;
    // re-lay every 32-float gradient block from [a][z] to [a / 2][z][a % 2] (in place; lane <-> destination element)
    {
        const int src = (2 * (lane >> 3) + (lane & 1)) * 4 + ((lane >> 1) & 3);
        for (int blk = 0; blk < nPairs; blk++) {
            const float v = sG[blk * kBwdPitch + src];
            __syncwarp();
            sG[blk * kBwdPitch + lane] = v;
        }
        __syncwarp();
    }
    const int total = (cnt * (cnt - 1)) >> 1;
    const int step = min(32, 2 * cnt - 2);
    int a = lane, d1 = 0;
    if (a >= cnt) { a -= cnt; d1 = 1; }
    const bool laneOn = lane < step;
    for (int q0 = 0; q0 < total; q0 += step) {
        if (laneOn && q0 + lane < total) {
            int b = a + d1 + 1;
            b -= (b >= cnt) ? cnt : 0;
            const float4 va = sA[a], vb = sA[b], wa = sB[a], wb = sB[b];
            const int spa = __float_as_int(wa.w) >> 24, spb = __float_as_int(wb.w) >> 24;
            const float* gp = sG + pairTab[spa * S + spb] * kBwdPitch;
            const float c = fmaf(va.z, vb.z, fmaf(va.y, vb.y, va.x * vb.x));
            const float xx = fmaxf(fmaf(-c, c, 1.0f), 1e-30f);
            const float isn = rsqrta(xx), sn = xx * isn;
            const float rm = va.w + vb.w;
            // angular factors, two shifts at a time (packed fp32x2): P_z = base^zeta, Q_z = base^(zeta - 1) (c sin_z - sn cos_z)
            const u64 c2 = pk2(c, c), sn2 = pk2(sn, sn), nsn2 = pk2(-sn, -sn);
            u64 P2[NSZ / 2], Q2[NSZ / 2];
#pragma unroll
            for (int z = 0; z < NSZ / 2; z++) {
                const u64 base2 = ffma2(sn2, sz2[z], ffma2(c2, cz2[z], one2));
                float b0, b1;
                unpk2(base2, b0, b1);
                const u64 pm2 = pk2(ex2a(zm1 * lg2a(fabsf(b0))), ex2a(zm1 * lg2a(fabsf(b1))));
                P2[z] = fmul2(pm2, base2);
                Q2[z] = fmul2(pm2, ffma2(c2, sz2[z], fmul2(nsn2, cz2[z])));
            }
            // radial factors two shifts at a time; the staged gradient block is laid out [a / 2][z][a % 2], so that one 16-byte load
            // brings the (a, a + 1) pairs of two angular channels:  GE_z = sum_a g[a][z] E_a,  GT_z = sum_a g[a][z] E_a (rm - Rs_a)
            const u64 rm2 = pk2(rm, rm);
            u64 GE2[NSZ], GT2[NSZ];
#pragma unroll
            for (int z = 0; z < NSZ; z++) { GE2[z] = pk2(0.f, 0.f); GT2[z] = pk2(0.f, 0.f); }
#pragma unroll
            for (int k = 0; k < NSA / 2; k++) {
                const u64 tt2 = fadd2(rm2, nShf2[k]);
                const u64 ar2 = fmul2(fmul2(tt2, nEta2), tt2);
                float a0, a1;
                unpk2(ar2, a0, a1);
                const u64 E2 = pk2(ex2a(a0), ex2a(a1));
                const u64 Et2 = fmul2(E2, tt2);
                const float4 ga = *reinterpret_cast<const float4*>(gp + k * 8), gb = *reinterpret_cast<const float4*>(gp + k * 8 + 4);
                const u64 g0 = pk2(ga.x, ga.y), g1 = pk2(ga.z, ga.w), g2 = pk2(gb.x, gb.y), g3 = pk2(gb.z, gb.w);
                GE2[0] = ffma2(g0, E2, GE2[0]); GE2[1] = ffma2(g1, E2, GE2[1]); GE2[2] = ffma2(g2, E2, GE2[2]); GE2[3] = ffma2(g3, E2, GE2[3]);
                GT2[0] = ffma2(g0, Et2, GT2[0]); GT2[1] = ffma2(g1, Et2, GT2[1]); GT2[2] = ffma2(g2, Et2, GT2[2]); GT2[3] = ffma2(g3, Et2, GT2[3]);
            }
            float P[NSZ], Q[NSZ], GE[NSZ], GT[NSZ];
#pragma unroll
            for (int z = 0; z < NSZ / 2; z++) { unpk2(P2[z], P[2 * z], P[2 * z + 1]); unpk2(Q2[z], Q[2 * z], Q[2 * z + 1]); }
#pragma unroll
            for (int z = 0; z < NSZ; z++) {
                float lo, hi;
                unpk2(GE2[z], lo, hi); GE[z] = lo + hi;
                unpk2(GT2[z], lo, hi); GT[z] = lo + hi;
            }
            float A = 0.0f, Bp = 0.0f, C = 0.0f;
#pragma unroll
            for (int z = 0; z < NSZ; z++) { A = fmaf(P[z], GE[z], A); Bp = fmaf(P[z], GT[z], Bp); C = fmaf(Q[z], GE[z], C); }
            // W_a = k1 (fc'_a fc_b A + F B),  Kc = sqrt(cs) / sin(theta') * F * C  (see the derivation in DESIGN.md section 3)
            const float F = wa.x * wb.x;
            const float FB = F * (kB * Bp);
            const float Ak = kA * A;
            const float Wa = fmaf(wa.y * wb.x, Ak, FB), Wb = fmaf(wa.x * wb.y, Ak, FB);
            const float Kc = (kC * isn) * (F * C);
            const float ga = Kc * wa.z, gb = Kc * wb.z;
            const float cth = c * rcs;
            const float al = fmaf(ga, cth, Wa), be = fmaf(gb, cth, Wb);
            const float fax = fmaf(al, va.x, -ga * vb.x), fay = fmaf(al, va.y, -ga * vb.y), faz = fmaf(al, va.z, -ga * vb.z);
            const float fbx = fmaf(be, vb.x, -gb * va.x), fby = fmaf(be, vb.y, -gb * va.y), fbz = fmaf(be, vb.z, -gb * va.z);
            float4* xa = sAcc + (d1 & 1) * capA + a;
            float4* yb = sAcc + (2 + (d1 & 1)) * capA + b;
            float4 ax = *xa, by = *yb;
            ax.x += fax; ax.y += fay; ax.z += faz;
            by.x += fbx; by.y += fby; by.z += fbz;
            *xa = ax; *yb = by;
        }
        __syncwarp();
        a += step;
        if (a >= cnt) { a -= cnt; d1++; }
        if (a >= cnt) { a -= cnt; d1++; }
    }
    float cx = 0.0f, cy = 0.0f, czc = 0.0f;
    for (int q = lane; q < cnt; q += 32) {
        const float4 x0 = sAcc[q], x1 = sAcc[capA + q], y0 = sAcc[2 * capA + q], y1 = sAcc[3 * capA + q];
        const float fx = (x0.x + x1.x) + (y0.x + y1.x), fy = (x0.y + x1.y) + (y0.y + y1.y), fz = (x0.z + x1.z) + (y0.z + y1.z);
        grad_add<VEC>(posGrad, (size_t)(__float_as_int(sB[q].w) & 0x00ffffff), fx, fy, fz);
        cx -= fx; cy -= fy; czc -= fz;
    }
    cx = warp_sum(cx); cy = warp_sum(cy); czc = warp_sum(czc);
    if (lane == 0) grad_add<VEC>(posGrad, (size_t)orig, cx, cy, czc);
    __syncwarp();                              // every lane is done with the staged rows before the next bulk copies land on them
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Geometry of the angular neighbours (geoA / geoB, see the top of the file): 8 lanes per centre, 4 centres per warp.
// delta and r2 are the reference's fp32 expressions (min_image_mul), as in the row kernel that accepted the pair.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ani_angular_geo_kernel(int n, const AniTables* __restrict__ tab, const float4* __restrict__ sorted, const int* __restrict__ sortedOrig,
                       const Geom* __restrict__ geom, const int* __restrict__ rowAng, const int* __restrict__ offAng, int capA,
                       float4* __restrict__ geoA, float4* __restrict__ geoB) {
    __shared__ Geom g;
    if (threadIdx.x == 0) g = *geom;
    __syncthreads();
    const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, gl = threadIdx.x & 7;
    if (p >= n) return;
    const int S = tab->nSpecies;
    const int cnt = offAng[(size_t)p * (S + 1) + S];
    const float4 ci = sorted[p];
    const float invRca = tab->invRca, sq = tab->sqrtCosScale;
    for (int q0 = 0; q0 < cnt; q0 += 32) {       // 4 entries per lane per trip: all index loads, then all coordinate gathers, in flight together
        int j[4];
        float4 cj[4];
#pragma unroll
        for (int u = 0; u < 4; u++) { const int q = q0 + gl + 8 * u; j[u] = q < cnt ? rowAng[(size_t)p * capA + q] : -1; }
#pragma unroll
        for (int u = 0; u < 4; u++) cj[u] = j[u] >= 0 ? sorted[j[u]] : ci;
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (j[u] < 0) continue;
            const int q = q0 + gl + 8 * u;
            float dx = __fsub_rn(cj[u].x, ci.x), dy = __fsub_rn(cj[u].y, ci.y), dz = __fsub_rn(cj[u].z, ci.z);
            const float r = sqrtf(min_image_mul(g, dx, dy, dz));
            const float ir = 1.0f / r;
            float sn, cs;
            sincospif(r * invRca, &sn, &cs);
            const float k = ir * sq;
            geoA[(size_t)p * capA + q] = make_float4(dx * k, dy * k, dz * k, 0.5f * r);
            geoB[(size_t)p * capA + q] = make_float4(0.5f * cs + 0.5f, -0.5f * kPi * invRca * sn, ir, cj[u].w);   // .w: species << 24 | atom index
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Radial forward (CpuANISymmetryFunctions.cpp:139-160), warp per centre.  Pass 1, lane <-> neighbour: pair geometry, written to
// radGeoA = {unit vector, r} / radGeoB = {fc, fc', species << 24 | atom index} for the backward kernel and staged as (r, fc) in
// shared memory.  Pass 2, lane = (h, kq): kq owns the channel pair (2 kq, 2 kq + 1) -- packed fp32x2 arithmetic -- and h is one
// of 32 / KQ neighbour sub-streams; one shuffle reduction over h per species block.
// ------------------------------------------------------------------------------------------------------------------
template <int NR>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
ani_radial_fwd_v2_kernel(int n, const AniTables* __restrict__ tab, const float4* __restrict__ sorted, const int* __restrict__ sortedOrig,
                         const Geom* __restrict__ geom, const int* __restrict__ rowRad, const int* __restrict__ offRad, int capR,
                         float4* __restrict__ radGeoA, float4* __restrict__ radGeoB, const int* __restrict__ rowMap, AevOutPtr out,
                         int stride) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    __shared__ Geom g;
    if (threadIdx.x == 0) g = *geom;
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float2* sRF = reinterpret_cast<float2*>(smemRaw) + (size_t)w * capR;
    constexpr int nR = NR;
    const int S = tab->nSpecies;
    const float invRcr = tab->invRcr;
    static_assert(NR == 4 || NR == 8 || NR == 16 || NR == 32 || NR == 64, "channel pairs must tile a warp");
    constexpr int KQ = NR / 2, H = 32 / KQ;
    const int kq = lane % KQ, h = lane / KQ;
    const float nEta = -tab->rEtaL2[0], scale = tab->radialScale;
    const u64 nEta2 = pk2(nEta, nEta);
    const u64 nShf2 = pk2(-tab->rShf[2 * kq], -tab->rShf[2 * kq + 1]);
    // persistent: a warp walks over centres (a centre is a few hundred warp instructions)
    for (int p = blockIdx.x * kWarpsPerCta + w; p < n; p += gridDim.x * kWarpsPerCta) {
    const int* off = offRad + (size_t)p * (S + 1);
    const int cnt = min(off[S], capR);
    const float4 ci = sorted[p];
    for (int q0 = 0; q0 < cnt; q0 += 64) {       // two entries per lane per trip, both gathers in flight together
        int j[2];
        float4 cj[2];
#pragma unroll
        for (int u = 0; u < 2; u++) { const int q = q0 + lane + 32 * u; j[u] = q < cnt ? rowRad[(size_t)p * capR + q] : -1; }
#pragma unroll
        for (int u = 0; u < 2; u++) cj[u] = j[u] >= 0 ? sorted[j[u]] : ci;
#pragma unroll
        for (int u = 0; u < 2; u++) {
            if (j[u] < 0) continue;
            const int q = q0 + lane + 32 * u;
            float dx = __fsub_rn(cj[u].x, ci.x), dy = __fsub_rn(cj[u].y, ci.y), dz = __fsub_rn(cj[u].z, ci.z);
            const float r = sqrtf(min_image_mul(g, dx, dy, dz));
            const float ir = 1.0f / r;
            float sn, cs;
            sincospif(r * invRcr, &sn, &cs);
            const float fc = 0.5f * cs + 0.5f;
            radGeoA[(size_t)p * capR + q] = make_float4(dx * ir, dy * ir, dz * ir, r);
            radGeoB[(size_t)p * capR + q] = make_float4(fc, -0.5f * kPi * invRcr * sn, cj[u].w, 0.0f);   // .z: species << 24 | atom index
            sRF[q] = make_float2(r, fc);
        }
    }
    __syncwarp();
    const int orig = sortedOrig[p];
    const size_t orow = (size_t)(rowMap ? rowMap[orig] : orig) * stride;
    for (int sp = 0; sp < S; sp++) {
        const int b = min(off[sp], cnt), e2 = min(off[sp + 1], cnt);
        u64 acc = pk2(0.f, 0.f);
        for (int q = b + h; q < e2; q += H) {
            const float2 rf = sRF[q];
            const u64 t2 = fadd2(pk2(rf.x, rf.x), nShf2);
            const u64 a2 = fmul2(fmul2(t2, nEta2), t2);
            float a0, a1;
            unpk2(a2, a0, a1);
            acc = ffma2(pk2(rf.y, rf.y), pk2(ex2a(a0), ex2a(a1)), acc);
        }
        float v0, v1;
        unpk2(acc, v0, v1);
        for (int o = KQ; o < 32; o <<= 1) { v0 += __shfl_xor_sync(kFull, v0, o); v1 += __shfl_xor_sync(kFull, v1, o); }
        if (h == 0) {
            const size_t idx = orow + (size_t)sp * nR + 2 * kq;
            v0 *= scale; v1 *= scale;
            if (out.hi) {
                const __half2 h2 = __floats2half2_rn(v0, v1);
                const float2 f2 = __half22float2(h2);
                *reinterpret_cast<__half2*>(out.hi + idx) = h2;
                *reinterpret_cast<__half2*>(out.lo + idx) = __floats2half2_rn((v0 - f2.x) * 2048.0f, (v1 - f2.y) * 2048.0f);
            } else {
                *reinterpret_cast<float2*>(out.f32 + idx) = make_float2(v0, v1);
            }
        }
    }
    __syncwarp();                              // the staged (r, fc) pairs are re-used by the next centre
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Radial backward, centre-owned (CpuANISymmetryFunctions.cpp:228-263 restricted to the centre's own AEV row; summed over all
// centres -- or ranks -- it equals the reference's gather form).  Warp per centre, lane <-> neighbour; the pair geometry comes from
// the rows the row kernel wrote (coalesced 16-byte loads, nothing is re-derived), the centre's gradient row is staged by one bulk copy:
//   w_ij = scale * sum_k G[i][s_j][k] e_k (fc' - 2 eta (r - Rs_k) fc) = scale * (fc' S0 - 2 eta fc S1),
//   S0 = sum_k g_k e_k,  S1 = sum_k g_k e_k (r - Rs_k)
// evaluated two channels at a time in packed fp32x2.
// ------------------------------------------------------------------------------------------------------------------
template <int NR, bool VEC>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
ani_radial_bwd_v2_kernel(int n, const AniTables* __restrict__ tab, const int* __restrict__ offRad, int capR,
                         const float4* __restrict__ radGeoA, const float4* __restrict__ radGeoB, const int* __restrict__ sortedOrig,
                         const int* __restrict__ rowMap, const float* __restrict__ grad, int stride, float* __restrict__ posGrad) {
    static_assert(NR % 4 == 0, "channel quads");
    extern __shared__ __align__(128) unsigned char smemRaw[];
    __shared__ __align__(8) unsigned long long bars[kWarpsPerCta];
    const int S = tab->nSpecies;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) mbar_init(smem_u32(&bars[w]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    float* sG = reinterpret_cast<float*>(smemRaw) + (size_t)w * S * NR;
    const uint32_t bar = smem_u32(&bars[w]);
    const float nEta = -tab->rEtaL2[0];
    const u64 nEta2 = pk2(nEta, nEta);
    u64 nShf2[NR / 2];
#pragma unroll
    for (int k = 0; k < NR / 2; k++) nShf2[k] = pk2(-tab->rShf[2 * k], -tab->rShf[2 * k + 1]);
    const float sc = tab->radialScale, m2eta = -2.0f * tab->rEta[0];
    uint32_t phase = 0;
    // persistent: a warp walks over centres (a centre is ~400 warp instructions -- one CTA per 8 centres spent as long being launched)
    for (int p = blockIdx.x * kWarpsPerCta + w; p < n; p += gridDim.x * kWarpsPerCta) {
    const int cnt = min(offRad[(size_t)p * (S + 1) + S], capR);
    if (cnt == 0) continue;                    // also every centre owned by another rank
    const int orig = sortedOrig[p];
    const float* gi = grad + (size_t)(rowMap ? rowMap[orig] : orig) * stride;
    if (lane == 0) {
        mbar_expect_tx(bar, (uint32_t)(S * NR * 4));
        bulk_g2s(smem_u32(sG), gi, (uint32_t)(S * NR * 4), bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1u;
    float fx = 0.0f, fy = 0.0f, fz = 0.0f;
    for (int q0 = 0; q0 < cnt; q0 += 64) {       // two pairs per lane per trip: the four 16-byte loads are in flight together
        float4 gaa[2], gbb[2];
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int q = q0 + lane + 32 * u;
            if (q < cnt) { gaa[u] = radGeoA[(size_t)p * capR + q]; gbb[u] = radGeoB[(size_t)p * capR + q]; }
        }
#pragma unroll
        for (int u = 0; u < 2; u++) {
            if (q0 + lane + 32 * u >= cnt) continue;
            const float4 ga = gaa[u], gb = gbb[u];
            const int tagged = __float_as_int(gb.z);
            const float* gp = sG + (tagged >> 24) * NR;
            const u64 r2 = pk2(ga.w, ga.w);
            u64 S0 = pk2(0.f, 0.f), S1 = pk2(0.f, 0.f);
#pragma unroll
            for (int k = 0; k < NR / 2; k += 2) {
                const float4 g4 = *reinterpret_cast<const float4*>(gp + 2 * k);
                const u64 t0 = fadd2(r2, nShf2[k]), t1 = fadd2(r2, nShf2[k + 1]);
                const u64 a0 = fmul2(fmul2(t0, nEta2), t0), a1 = fmul2(fmul2(t1, nEta2), t1);
                float x0, x1, x2, x3;
                unpk2(a0, x0, x1); unpk2(a1, x2, x3);
                const u64 e0 = pk2(ex2a(x0), ex2a(x1)), e1 = pk2(ex2a(x2), ex2a(x3));
                const u64 g0 = pk2(g4.x, g4.y), g1 = pk2(g4.z, g4.w);
                S0 = ffma2(g0, e0, S0); S0 = ffma2(g1, e1, S0);
                S1 = ffma2(g0, fmul2(e0, t0), S1); S1 = ffma2(g1, fmul2(e1, t1), S1);
            }
            float s00, s01, s10, s11;
            unpk2(S0, s00, s01); unpk2(S1, s10, s11);
            const float wr = sc * fmaf(gb.y, s00 + s01, m2eta * gb.x * (s10 + s11));
            const float gx = wr * ga.x, gy = wr * ga.y, gz = wr * ga.z;
            fx -= gx; fy -= gy; fz -= gz;
            grad_add<VEC>(posGrad, (size_t)(tagged & 0x00ffffff), gx, gy, gz);
        }
    }
    fx = warp_sum(fx); fy = warp_sum(fy); fz = warp_sum(fz);
    if (lane == 0) grad_add<VEC>(posGrad, (size_t)orig, fx, fy, fz);
    __syncwarp();                              // every lane is done with the staged gradient row before the next bulk copy
    }
}

__global__ void grad_compact_kernel(int n, const float4* __restrict__ acc, float* __restrict__ posGrad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 v = acc[i];
    posGrad[3 * (size_t)i] = v.x; posGrad[3 * (size_t)i + 1] = v.y; posGrad[3 * (size_t)i + 2] = v.z;
}

template <typename K>
void set_smem_v2(K kernel, size_t bytes) {
    // the attribute is per device and per kernel instantiation
    int dev = 0;
    cudaGetDevice(&dev);
    static std::mutex mtx;
    static std::map<std::pair<const void*, int>, size_t> have;
    std::lock_guard<std::mutex> lock(mtx);
    size_t& h = have[{reinterpret_cast<const void*>(kernel), dev}];
    if (h >= bytes && h != 0) return;
    NNP_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(bytes, 1)));
    NNP_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    h = std::max<size_t>(bytes, 1);
}

int sm_count_v2() {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    static int cache[64] = {0};
    if (dev >= 0 && dev < 64 && cache[dev]) return cache[dev];
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (dev >= 0 && dev < 64) cache[dev] = sms;
    return sms;
}

}  // namespace

bool radial_v2_supported(const AniTables& t) {
    if (t.nRadial != 16 || !t.torchani) return false;
    for (int k = 1; k < t.nRadial; k++)
        if (t.rEta[k] != t.rEta[0]) return false;
    return true;
}

void radial_v2_forward(int n, const AniTables& tabHost, const AniTables* tab, const float4* sorted, const int* sortedOrig, const Geom* geom,
                       const int* rowRad, const int* offRad, int capR, float4* radGeoA, float4* radGeoB, const int* rowMap, AevOutPtr out,
                       int stride, cudaStream_t stream) {
    const size_t smem = (size_t)kWarpsPerCta * capR * sizeof(float2);
    auto k = ani_radial_fwd_v2_kernel<16>;
    set_smem_v2(k, smem);
    const int grid = (n + kWarpsPerCta - 1) / kWarpsPerCta;   // latency-bound (gathers): one centre per warp keeps the most warps in flight
    k<<<grid, kWarpsPerCta * 32, smem, stream>>>(n, tab, sorted, sortedOrig, geom, rowRad, offRad, capR, radGeoA, radGeoB, rowMap, out, stride);
    count_launch();
}

void angular_v2_geometry(int n, const AniTables& tabHost, const AniTables* tab, const float4* sorted, const int* sortedOrig, const Geom* geom,
                         const int* rowAng, const int* offAng, int capA, float4* geoA, float4* geoB, cudaStream_t stream) {
    const long long threads = (long long)n * 8;
    ani_angular_geo_kernel<<<(int)((threads + 255) / 256), 256, 0, stream>>>(n, tab, sorted, sortedOrig, geom, rowAng, offAng, capA, geoA, geoB);
    count_launch();
}

void radial_v2_backward(int n, const AniTables& tabHost, const AniTables* tab, const int* offRad, int capR, const float4* radGeoA,
                        const float4* radGeoB, const int* sortedOrig, const int* rowMap, const float* grad, int stride, float* posGrad,
                        bool padded, cudaStream_t stream) {
    const size_t smem = (size_t)kWarpsPerCta * tabHost.nSpecies * 16 * sizeof(float);
    const int grid = (n + kWarpsPerCta - 1) / kWarpsPerCta;   // latency-bound: one centre per warp keeps the most warps in flight
    if (padded) {
        auto k = ani_radial_bwd_v2_kernel<16, true>;
        set_smem_v2(k, smem);
        k<<<grid, kWarpsPerCta * 32, smem, stream>>>(n, tab, offRad, capR, radGeoA, radGeoB, sortedOrig, rowMap, grad, stride, posGrad);
    } else {
        auto k = ani_radial_bwd_v2_kernel<16, false>;
        set_smem_v2(k, smem);
        k<<<grid, kWarpsPerCta * 32, smem, stream>>>(n, tab, offRad, capR, radGeoA, radGeoB, sortedOrig, rowMap, grad, stride, posGrad);
    }
    count_launch();
}

void grad_compact(int n, const float4* acc, float* posGrad, cudaStream_t stream) {
    grad_compact_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, acc, posGrad);
    count_launch();
}

bool angular_v2_supported(const AniTables& t) {
    return t.fast && t.nShfA == 8 && t.nShfZ == 4 && t.torchani && t.nAngular == 32 && t.nSpecies <= 15;
}

void angular_v2_build_segments(int n, const AniTables& tabHost, const int* offAng, const int* hist, int* cursor, int4* segs, int* nSeg,
                               const int* sortedOrig, const int* rowMap, AevOutPtr out, int stride, cudaStream_t stream) {
    NNP_CUDA_CHECK(cudaMemsetAsync(const_cast<int*>(hist), 0, 2 * kSegBins * sizeof(int), stream));   // hist and cursor are adjacent
    const int grid = (n + kSegBins - 1) / kSegBins;
    ani_seg_hist_kernel<<<grid, kSegBins, 0, stream>>>(n, tabHost.nSpecies, offAng, const_cast<int*>(hist));
    ani_seg_scatter_kernel<<<grid, kSegBins, 0, stream>>>(n, tabHost.nSpecies, offAng, hist, cursor, segs, nSeg, sortedOrig, rowMap, out, stride);
    count_launch(2);
}

void angular_v2_forward(int n, const AniTables& tabHost, const AniTables* tab, const int* offAng, int capA, const float4* geoA,
                        const float4* geoB, const int4* segs, const int* nSeg, const int* sortedOrig, const int* rowMap, AevOutPtr out,
                        int stride, cudaStream_t stream) {
    NNP_REQUIRE(2 * capA <= kFwdBudget, "angular neighbour capacity above 128 is not supported by the segment kernel");
    const size_t smem = (size_t)kWarpsPerCta * kFwdBudget * 20;
    auto k = ani_angular_fwd_seg_kernel<8, 4>;
    set_smem_v2(k, smem);
    // persistent: at most 3 CTAs per SM, never more CTAs than chunks of 8 segments could exist
    const long long maxChunks = ((long long)n * tabHost.nPairs + 7) / 8;
    const int grid = (int)std::min<long long>((long long)sm_count_v2() * NNP_FWD_MINB, (maxChunks + kWarpsPerCta - 1) / kWarpsPerCta);
    if (grid <= 0) return;
    k<<<grid, kWarpsPerCta * 32, smem, stream>>>(tab, offAng, capA, geoA, geoB, segs, nSeg, out, stride);
    count_launch();
}

void angular_v2_backward(int n, const AniTables& tabHost, const AniTables* tab, const int* offAng, int capA, const float4* geoA,
                         const float4* geoB, const int* sortedOrig, const int* rowMap, const float* grad, int stride, float* posGrad,
                         bool padded, cudaStream_t stream) {
    const size_t smem = (size_t)kWarpsPerCta * ((size_t)capA * 96 + (size_t)tabHost.nPairs * kBwdPitch * 4);
    NNP_REQUIRE(smem <= 200 * 1024, "angular gradient row does not fit in shared memory (numSpecies^2 * numAngular too large)");
    const int perSm = std::max(1, std::min(4, (int)((220 * 1024) / (smem + 1024))));
    const int grid = std::min((n + kWarpsPerCta - 1) / kWarpsPerCta, sm_count_v2() * perSm);
    if (padded) {
        auto k = ani_angular_bwd_v2_kernel<8, 4, true>;
        set_smem_v2(k, smem);
        k<<<grid, kWarpsPerCta * 32, smem, stream>>>(n, tab, offAng, capA, geoA, geoB, sortedOrig, rowMap, grad, stride, posGrad);
    } else {
        auto k = ani_angular_bwd_v2_kernel<8, 4, false>;
        set_smem_v2(k, smem);
        k<<<grid, kWarpsPerCta * 32, smem, stream>>>(n, tab, offAng, capA, geoA, geoB, sortedOrig, rowMap, grad, stride, posGrad);
    }
    count_launch();
}

}  // namespace nnpops
