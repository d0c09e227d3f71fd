// Fused per-tile layer chain of the species MLP on tcgen05, single-accumulator variant (NNPOPS_CHAIN_V2=1; default: mlp_chain.cu).
//
// Same chain, same operands and same results as mlp_chain.cu -- six GEMMs F1 F2 F3 G3 G2 G1 per (128-atom tile, ensemble member), fp16
// hi/lo operand pairs (x = hi + 2^-11 lo), activations alternating between tensor memory and shared memory -- restructured after the
// event traces of that kernel (profiles/r08_summary.md: its MMA side alone, with the epilogue arithmetic removed, sustains 60 % of the
// tensor pipe, bounded by accumulator hand-overs and by the issue rate of one thread):
//   * ONE accumulator per output column instead of a (hi.hi | cross terms) pair.  The cross terms are accumulated first,
//         D  = sum_k Ahi_k . Blo_k + Alo_k . Bhi_k                (scaled by 2^11, like the lo operands)
//     and the first hi.hi MMA folds them in with the tensor core's own input scaling (tcgen05.mma ... scale-input-d = 11):
//         D  = D * 2^-11 + Ahi_0 . Bhi_0 ,   D += Ahi_k . Bhi_k   (exact: a power of two)
//     so an accumulator stage of 128 columns holds a chunk of 128 OUTPUT columns: half as many accumulator hand-overs per chain
//     (11 instead of 19), MMAs of N = 128 throughout (12 instead of 16 instructions per 128 x 64 block of a weight matrix), half the
//     tensor-memory loads in the epilogue and no combine step.  The price: the hi half of a weight block is fetched twice (L2), i.e.
//     102 instead of 57 block waits per chain on the issuing thread.
//   * all sixteen epilogue warps work on every chunk (32 columns per thread).
//   * ONE MMA issuer hands the chunks to the tensor pipe in order, so that the epilogue of a chunk runs under the MMAs of the next;
//     16 KB weight blocks are [128 output rows][64 k] of one part (hi or lo), ring of eight, one producer thread for the weights and
//     one for X.
// Measured on B200 (50 000-atom bench box): 0.65 ms per evaluation against 0.54 ms of mlp_chain.cu (0.61 ms with two concurrent
// issuers): the block waits and commits of the issuing thread cost more than the saved MMAs and hand-overs.  Not the default; kept,
// tested (tests/test_ani_gpu.py), as the single-CTA half of a cta_group::2 kernel, where one instruction drives two SMs.
#include "mlp_chain.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include "tcgen05_util.cuh"

namespace nnpops {

namespace {
using namespace tc;

constexpr int kMaxSp = 7;
constexpr int kRows = 128;
constexpr uint32_t kTile = 128 * 128;        // 16 KB: [128 rows][64 halves]
constexpr int kMaxRing = 8;
constexpr int kFirstEpiWarp = 4, kEpiWarps = 16;   // warps 0..3: weight producer, MMA issuer, idle, X producer
constexpr int kThreads = (kFirstEpiWarp + kEpiWarps) * 32;
constexpr uint32_t kAccCols = 128;           // one accumulator stage: a chunk of 128 output columns
constexpr uint32_t kOpaHi = 256, kOpaLo = 384;   // A operand in TMEM: packed fp16 pairs, hi part and lo part (128 columns each: K <= 256)
constexpr int kStashCols = 256;
constexpr uint32_t kMaxSmem = 232448;
constexpr uint32_t kLoShift = 11;            // log2(kLoScale)

struct ChainSpecies {
    int d0, d1, d2, d3;
    int rows, rowStart, tileBegin, tileEnd;
    const float* bias[3];
    const float* w3;
};

struct ChainParams {
    CUtensorMap maps[kMaxSp][8];   // 0..5: packed weights of F1 F2 F3 G3 G2 G1; 6, 7: X hi / lo of the species' rows
    ChainSpecies sp[kMaxSp];
    int numSpecies, numTiles, M;
    int mpu, numUnits;             // work unit = (tile, mpu consecutive ensemble members)
    int sChunks, ring;             // 64-column tiles of the shared-memory activation buffer (per part); weight ring depth
    float* dX;
    int ldx;
    float* stash;                  // [grid][256 columns][128 rows] fp32
    double* energyAcc;
    float seedScale, outScale;
};

__host__ __device__ constexpr int chunks64(int n) { return (n + 63) >> 6; }
__host__ __device__ constexpr int chunks128(int n) { return (n + 127) >> 7; }

__device__ __forceinline__ void layer_dims(const ChainSpecies& s, int j, int& N, int& K) {
    switch (j) {
        case 0: N = s.d1; K = s.d0; break;
        case 1: N = s.d2; K = s.d1; break;
        case 2: N = s.d3; K = s.d2; break;
        case 3: N = s.d2; K = s.d3; break;
        case 4: N = s.d1; K = s.d2; break;
        default: N = s.d0; K = s.d1; break;
    }
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// tcgen05.mma with the shared-memory descriptors given as low words (start address | LBO) + one common high word.  SCALED: the
// accumulator is multiplied by 2^-11 before the product is added (scale-input-d, an immediate).
template <bool SCALED>
__device__ __forceinline__ void umma_ss(uint32_t tmemD, uint32_t aLo, uint32_t bLo, uint32_t descHi, uint32_t idesc, uint32_t accumulate) {
    if (SCALED)
        asm volatile(
            "{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %5, 0;\nmov.b64 da, {%1, %3};\nmov.b64 db, {%2, %3};\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p, 11;\n}\n" ::"r"(tmemD), "r"(aLo), "r"(bLo), "r"(descHi), "r"(idesc), "r"(accumulate) : "memory");
    else
        asm volatile(
            "{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %5, 0;\nmov.b64 da, {%1, %3};\nmov.b64 db, {%2, %3};\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n}\n" ::"r"(tmemD), "r"(aLo), "r"(bLo), "r"(descHi), "r"(idesc), "r"(accumulate) : "memory");
}
template <bool SCALED>
__device__ __forceinline__ void umma_ts(uint32_t tmemD, uint32_t tmemA, uint32_t bLo, uint32_t descHi, uint32_t idesc, uint32_t accumulate) {
    if (SCALED)
        asm volatile(
            "{\n.reg .pred p;\n.reg .b64 db;\nsetp.ne.b32 p, %5, 0;\nmov.b64 db, {%2, %3};\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p, 11;\n}\n" ::"r"(tmemD), "r"(tmemA), "r"(bLo), "r"(descHi), "r"(idesc), "r"(accumulate) : "memory");
    else
        asm volatile(
            "{\n.reg .pred p;\n.reg .b64 db;\nsetp.ne.b32 p, %5, 0;\nmov.b64 db, {%2, %3};\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n}\n" ::"r"(tmemD), "r"(tmemA), "r"(bLo), "r"(descHi), "r"(idesc), "r"(accumulate) : "memory");
}
static_assert(kLoShift == 11, "the scale-input-d immediate in umma_ss / umma_ts is written out as 11");

// packed fp32x2 arithmetic (FFMA2 / FMUL2 / FADD2: one issue slot for two columns)
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 pk2u(uint32_t a, uint32_t b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ void unpk2(u64 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float ex2_f(float x) { float e; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x)); return e; }
// (a, b) as a packed pair -> packed fp16 pair of the high parts and packed pair of the scaled low parts
__device__ __forceinline__ void split_pack2(u64 v, uint32_t& hi, uint32_t& lo) {
    float a, b;
    unpk2(v, a, b);
    const __half2 h2 = __floats2half2_rn(a, b);
    const float2 f2 = __half22float2(h2);
    hi = *reinterpret_cast<const uint32_t*>(&h2);
    const u64 d = fmul2(ffma2(pk2(f2.x, f2.y), pk2(-1.0f, -1.0f), v), pk2(kLoScale, kLoScale));   // (v - hi) * 2^11, exact
    unpk2(d, a, b);
    lo = pack_h2(a, b);
}

// what an epilogue thread needs to know about its 32 columns of a chunk
struct EpiCtx {
    uint32_t laneBase;      // TMEM address of the thread's lane, column 0
    uint32_t sbufRow;       // shared-memory address of the thread's row in tile 0 of the activation buffer (hi part)
    uint32_t sLoOff;        // byte offset of the lo tiles of that buffer
    int r7;                 // row & 7 (swizzle phase)
    int cg;                 // column group of the warp inside a chunk: columns [32 cg, 32 cg + 32)
    int n0;                 // first column (of the layer's output) of the thread's 32
    uint32_t accCol;        // first TMEM column of those 32 accumulators
    bool rowOk;
};

// TYPE: 0 F1, 1 F2, 2 F3, 3 G3, 4 G2, 5 G1.  The accumulators are pulled into registers as 16 packed column pairs and the stage is
// handed back to the MMA warp BEFORE any arithmetic.
template <int TYPE>
__device__ __forceinline__ void epi_chunk(const ChainParams& P, const EpiCtx& x, const float* __restrict__ colA, const float* __restrict__ colB,
                                          float* stash, float* dxRow, bool firstMember, uint32_t accFullBar, uint32_t fullPhase, uint32_t accEmptyBar, int lane,
                                          float& esum) {
    // fetched before the accumulator is needed: the biases of the thread's columns (every lane reads the same 16-byte words:
    // broadcast), or -- G2 -- A1 of these columns, written to the stash by F1's epilogue of this chain
    u64 pre[(TYPE <= 2 || TYPE == 4) ? 16 : 1];
    if (TYPE <= 2) {
        const ulonglong2* src = reinterpret_cast<const ulonglong2*>(colA + x.n0);
#pragma unroll
        for (int i = 0; i < 8; i++) { const ulonglong2 t = __ldg(src + i); pre[2 * i] = t.x; pre[2 * i + 1] = t.y; }
    }
    if (TYPE == 4) {
#pragma unroll
        for (int i = 0; i < 16; i++) pre[i] = pk2(__ldcg(stash + (size_t)(x.n0 + 2 * i) * kRows), __ldcg(stash + (size_t)(x.n0 + 2 * i + 1) * kRows));
    }
    mbar_wait(accFullBar, fullPhase);
    tc_fence_after();
    u64 vv[16];
    {
        uint32_t r[32];
        tmem_ld32(x.laneBase + x.accCol, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; i++) vv[i] = pk2u(r[2 * i], r[2 * i + 1]);
    }
    // the thread's part of the accumulator stage has been read: hand it back to the MMA warp
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(accEmptyBar);

    if (TYPE == 5) {
        if (x.rowOk) {
            const u64 os2 = pk2(P.outScale, P.outScale);
            float* dst = dxRow + x.n0;
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                float a, b, c, d;
                unpk2(fmul2(vv[i], os2), a, b);
                unpk2(fmul2(vv[i + 1], os2), c, d);
                if (firstMember) *reinterpret_cast<float4*>(dst + 2 * i) = make_float4(a, b, c, d);
                else red_add_v4(dst + 2 * i, a, b, c, d);
            }
        }
        return;
    }
    const u64 celuScale2 = pk2(1.4426950408889634f / kCeluAlpha, 1.4426950408889634f / kCeluAlpha);
    const u64 alpha2 = pk2(kCeluAlpha, kCeluAlpha), negAlpha2 = pk2(-kCeluAlpha, -kCeluAlpha);
    const u64 invAlpha2 = pk2(1.0f / kCeluAlpha, 1.0f / kCeluAlpha), one2 = pk2(1.0f, 1.0f);
    const u64 loInv2 = pk2(kLoInv, kLoInv);
    u64 esum2 = pk2(0.0f, 0.0f);
    // the thread's 32 columns lie in 64-column tile (n0 >> 6) of the swizzled shared-memory buffer, 16-byte units 4 (cg & 1) ...
    const uint32_t tileAddr = x.sbufRow + (uint32_t)(x.n0 >> 6) * kTile;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int col = x.n0 + 16 * h;
        const int u0 = (x.cg & 1) * 4 + 2 * h;
        const uint32_t sA = tileAddr + (uint32_t)((u0 ^ x.r7) << 4);
        const uint32_t sB = tileAddr + (uint32_t)(((u0 + 1) ^ x.r7) << 4);
        u64 w3v[TYPE == 2 ? 8 : 1];
        uint32_t hh[TYPE == 3 ? 8 : 1], ll[TYPE == 3 ? 8 : 1];
        if (TYPE == 2) {   // output-layer weights of these columns
            const ulonglong2* src = reinterpret_cast<const ulonglong2*>(colB + col);
#pragma unroll
            for (int i = 0; i < 4; i++) { const ulonglong2 t = __ldg(src + i); w3v[2 * i] = t.x; w3v[2 * i + 1] = t.y; }
        }
        if (TYPE == 3) {   // celu'(A2) from the activation tile this thread is about to overwrite
            const uint4 h0 = ld_shared_v4(sA), h1 = ld_shared_v4(sB), l0 = ld_shared_v4(sA + x.sLoOff), l1 = ld_shared_v4(sB + x.sLoOff);
            hh[0] = h0.x; hh[1] = h0.y; hh[2] = h0.z; hh[3] = h0.w; hh[4] = h1.x; hh[5] = h1.y; hh[6] = h1.z; hh[7] = h1.w;
            ll[0] = l0.x; ll[1] = l0.y; ll[2] = l0.z; ll[3] = l0.w; ll[4] = l1.x; ll[5] = l1.y; ll[6] = l1.z; ll[7] = l1.w;
        }
        uint32_t ph[8], pl[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            u64 v = vv[8 * h + i];
            if (TYPE <= 2) {
                const u64 z2 = fadd2(v, pre[8 * h + i]);
                float z0, z1, t0, t1, n0, n1;
                unpk2(z2, z0, z1);
                unpk2(fmul2(z2, celuScale2), t0, t1);
                const float e0 = ex2_f(t0), e1 = ex2_f(t1);
                unpk2(ffma2(pk2(e0, e1), alpha2, negAlpha2), n0, n1);   // for large positive z the exponential overflows to +inf, which the select discards
                const bool p0 = z0 > 0.0f, p1 = z1 > 0.0f;
                const u64 a2 = pk2(p0 ? z0 : n0, p1 ? z1 : n1);
                if (TYPE == 2) {
                    esum2 = ffma2(a2, w3v[i], esum2);
                    const u64 seed2 = pk2(P.seedScale, P.seedScale);
                    v = fmul2(fmul2(w3v[i], seed2), pk2(p0 ? 1.0f : e0, p1 ? 1.0f : e1));
                } else {
                    v = a2;
                    if (TYPE == 0) {
                        __stcg(stash + (size_t)(col + 2 * i) * kRows, p0 ? z0 : n0);
                        __stcg(stash + (size_t)(col + 2 * i + 1) * kRows, p1 ? z1 : n1);
                    }
                }
            } else {
                u64 act2;
                if (TYPE == 4) act2 = pre[8 * h + i];
                else {
                    const float2 fh = unpack_h2(hh[i]), fl = unpack_h2(ll[i]);
                    act2 = ffma2(pk2(fl.x, fl.y), loInv2, pk2(fh.x, fh.y));
                }
                float a0, a1, g0, g1;
                unpk2(act2, a0, a1);
                unpk2(ffma2(act2, invAlpha2, one2), g0, g1);   // celu'(z) from celu(z): 1 for a > 0, a / alpha + 1 otherwise
                v = fmul2(v, pk2(a0 > 0.0f ? 1.0f : g0, a1 > 0.0f ? 1.0f : g1));
            }
            split_pack2(v, ph[i], pl[i]);
        }
        if (TYPE == 1 || TYPE == 3) {
            st_shared_v4(sA, ph[0], ph[1], ph[2], ph[3]);
            st_shared_v4(sB, ph[4], ph[5], ph[6], ph[7]);
            st_shared_v4(sA + x.sLoOff, pl[0], pl[1], pl[2], pl[3]);
            st_shared_v4(sB + x.sLoOff, pl[4], pl[5], pl[6], pl[7]);
        } else {
            tmem_st8(x.laneBase + kOpaHi + (uint32_t)(col >> 1), ph);
            tmem_st8(x.laneBase + kOpaLo + (uint32_t)(col >> 1), pl);
        }
    }
    if (TYPE == 2 && x.rowOk) {   // rows beyond the species' last atom hold zeros in X (TMA out-of-bounds fill), finite everywhere
        float s0, s1;
        unpk2(esum2, s0, s1);
        esum += s0 + s1;
    }
}

__global__ void __launch_bounds__(kThreads, 1) mlp_chain2_kernel(const __grid_constant__ ChainParams P) {
    extern __shared__ unsigned char smemRaw[];
    const uint32_t rawAddr = smem_u32(smemRaw);
    const uint32_t base = (rawAddr + 1023u) & ~1023u;
    const uint32_t sbuf = base;   // X of the chain (F1), then A2 (F2 -> F3, G3), then dZ1 (G3 -> G2): hi tiles [0, sChunks), lo tiles behind
    const uint32_t ring = sbuf + 2u * P.sChunks * kTile;
    const uint32_t barBase = ring + (uint32_t)P.ring * kTile;
    auto bFull = [&](int s) { return barBase + 8u * s; };
    auto bEmpty = [&](int s) { return barBase + 8u * (kMaxRing + s); };
    auto accFull = [&](int s) { return barBase + 8u * (2 * kMaxRing + s); };
    auto accEmpty = [&](int s) { return barBase + 8u * (2 * kMaxRing + 2 + s); };
    auto opReady = [&](int s) { return barBase + 8u * (2 * kMaxRing + 4 + s); };
    const uint32_t xFull = barBase + 8u * (2 * kMaxRing + 8), xEmpty = xFull + 8u;
    const uint32_t chainDone = xEmpty + 8u;   // the MMAs of a chain have completed (its last layer reads the TMEM operand)
    const uint32_t tmemSlot = chainDone + 8u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < P.numSpecies; s++)
            for (int m = 0; m < 8; m++) asm volatile("prefetch.tensormap [%0];" ::"l"(&P.maps[s][m]) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < P.ring; s++) { mbar_init(bFull(s), 1); mbar_init(bEmpty(s), 1); }
        for (int s = 0; s < 2; s++) { mbar_init(accFull(s), 1); mbar_init(accEmpty(s), kEpiWarps); }
        for (int s = 0; s < 4; s++) mbar_init(opReady(s), kEpiWarps / 2);   // a 64-column slice is written by 2 column groups x 4 lane quarters
        mbar_init(xFull, 1);
        mbar_init(xEmpty, 1);
        mbar_init(chainDone, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // flat chunk schedule of one chain per species (entry = layer << 3 | chunk of 128 columns)
    unsigned char* const chunkTab = smemRaw + (barBase - rawAddr) + 256;
    if (warp == 2 && lane < P.numSpecies) {
        int n = 0;
        for (int j = 0; j < 6; j++) {
            int N, K;
            layer_dims(P.sp[lane], j, N, K);
            for (int c = 0; c < chunks128(N); c++) chunkTab[lane * 32 + n++] = (unsigned char)(j << 3 | c);
        }
        chunkTab[lane * 32 + 31] = (unsigned char)n;
    }
    __syncwarp();
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmemSlot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmemBase;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmemBase) : "r"(tmemSlot));

    auto species_of = [&](int t) {
        int s = 0;
        while (s + 1 < P.numSpecies && t >= P.sp[s].tileEnd) s++;
        return s;
    };

    if (warp == 0) {
        // ---- weight producer: per chunk the 16 KB blocks arrive in the order the issuer consumes them -- pass 1 (cross terms) hi(kc),
        // lo(kc) for every k-chunk, pass 2 hi(kc) again
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            auto load_block = [&](const CUtensorMap* map, int k0, int row0) {
                mbar_wait(bEmpty(stage), phase ^ 1u);
                mbar_expect_tx(bFull(stage), kTile);
                tma_load_2d(ring + stage * kTile, map, bFull(stage), k0, row0);
                if (++stage == P.ring) { stage = 0; phase ^= 1u; }
            };
            for (int u = blockIdx.x; u < P.numUnits; u += gridDim.x) {
                const int upt = P.M / P.mpu, t = u / upt, e0 = (u - t * upt) * P.mpu, e1 = e0 + P.mpu;
                const int si = species_of(t);
                const ChainSpecies& sp = P.sp[si];
                for (int e = e0; e < e1; e++)
                    for (int j = 0; j < 6; j++) {
                        int N, K;
                        layer_dims(sp, j, N, K);
                        const int cn = chunks128(N), ck = chunks64(K);
                        for (int c = 0; c < cn; c++) {
                            const int row0 = (e * cn + c) * 256;   // 128 hi rows, then 128 lo rows
                            for (int kc = 0; kc < ck; kc++) {
                                load_block(&P.maps[si][j], kc * 64, row0);
                                load_block(&P.maps[si][j], kc * 64, row0 + 128);
                            }
                            for (int kc = 0; kc < ck; kc++) load_block(&P.maps[si][j], kc * 64, row0);
                        }
                    }
            }
        }
        __syncwarp();
    } else if (warp == 3) {
        // ---- X producer.  X shares the activation buffer: it is (re)loaded for every chain, as soon as G2 of the previous chain -- the
        // last reader of that buffer -- has completed (the issuer arrives on xEmpty when it has seen G1's operand)
        if (elect_one()) {
            uint32_t xPhase = 0;
            for (int u = blockIdx.x; u < P.numUnits; u += gridDim.x) {
                const int upt = P.M / P.mpu, t = u / upt, e0 = (u - t * upt) * P.mpu, e1 = e0 + P.mpu;
                const int si = species_of(t);
                const ChainSpecies& sp = P.sp[si];
                const int row0 = (t - sp.tileBegin) * kRows;
                const int cx = chunks64(sp.d0);
                for (int e = e0; e < e1; e++) {
                    mbar_wait(xEmpty, xPhase ^ 1u);
                    xPhase ^= 1u;
                    mbar_expect_tx(xFull, 2u * cx * kTile);
                    for (int kc = 0; kc < cx; kc++) {
                        tma_load_2d(sbuf + kc * kTile, &P.maps[si][6], xFull, kc * 64, row0);
                        tma_load_2d(sbuf + (P.sChunks + kc) * kTile, &P.maps[si][7], xFull, kc * 64, row0);
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ---- MMA issuer: ONE elected thread hands the chunks to the tensor pipe in order, alternating between the two accumulator
        // stages, so that the epilogue of a chunk runs under the MMAs of the next.  A chunk of layer j is issued only after every
        // 64-column slice of its A operand has been published (opReady), i.e. after all MMAs of layer j - 1 have completed and been
        // read by the epilogue.
        if (elect_one()) {
            const uint32_t descHi = (uint32_t)(make_desc(0) >> 32);
            auto desc_lo = [](uint32_t addr) { return ((addr >> 4) & 0x3fffu) | (1u << 16); };   // start address | LBO = 1
            constexpr uint32_t kTile16 = kTile >> 4;
            const uint32_t ringLo = desc_lo(ring), sLo = desc_lo(sbuf);
            int stage = 0, chunkIdx = 0;
            uint32_t phase = 0, accBits = 0, opBits = 0, xPhase = 0;
            // wait for the next weight block of the ring; returns the low word of its descriptor
            auto next_block = [&](int& slot) {
                mbar_wait(bFull(stage), phase);
                tc_fence_after();
                slot = stage;
                const uint32_t lo = ringLo + stage * kTile16;
                if (++stage == P.ring) { stage = 0; phase ^= 1u; }
                return lo;
            };
            for (int u = blockIdx.x; u < P.numUnits; u += gridDim.x) {
                const int upt = P.M / P.mpu, t = u / upt, e0 = (u - t * upt) * P.mpu, e1 = e0 + P.mpu;
                const int si = species_of(t);
                const ChainSpecies& sp = P.sp[si];
                for (int e = e0; e < e1; e++) {
                    mbar_wait(xFull, xPhase);
                    xPhase ^= 1u;
                    tc_fence_after();
                    for (int j = 0; j < 6; j++) {
                        int N, K;
                        layer_dims(sp, j, N, K);
                        const int cn = chunks128(N), ck = chunks64(K);
                        const int lastSteps = (K - (ck - 1) * 64) >> 4;
                        const bool aTmem = (j & 1) != 0;                 // F2, G3, G1 read their A operand from tensor memory
                        // A operand: address of k-chunk 0 (TMEM column, or descriptor low word), stride per k-chunk, offset of the lo part
                        const uint32_t aBase = aTmem ? tmemBase + kOpaHi : sLo;
                        const uint32_t aChunk = aTmem ? 32u : kTile16;
                        const uint32_t aLoOff = aTmem ? (kOpaLo - kOpaHi) : (uint32_t)P.sChunks * kTile16;
                        const uint32_t aStep = aTmem ? 8u : 2u;          // per k16 step
                        bool needOp = j > 0;   // the A operand of this layer has not been waited for yet
                        for (int c = 0; c < cn; c++) {
                            const int st = chunkIdx++ & 1;
                            const uint32_t d = tmemBase + st * kAccCols;
                            const int nValid = min(128, N - 128 * c);
                            const uint32_t idesc = idesc_f16(128, nValid);
                            mbar_wait(accEmpty(st), ((accBits >> st) & 1u) ^ 1u);
                            accBits ^= 1u << st;
                            tc_fence_after();
                            // pass 1: the cross terms, D = sum_k Ahi_k . Blo_k + Alo_k . Bhi_k
                            for (int kc = 0; kc < ck; kc++) {
                                if (needOp) {   // columns [64 kc, 64 kc + 64) of the A operand come from the previous layer's epilogue
                                    mbar_wait(opReady(kc), (opBits >> kc) & 1u);
                                    opBits ^= 1u << kc;
                                    tc_fence_after();
                                }
                                int slotH, slotL;
                                const uint32_t bHi = next_block(slotH);
                                const uint32_t bLo = next_block(slotL);
                                const uint32_t a = aBase + kc * aChunk;
                                const int steps = kc == ck - 1 ? lastSteps : 4;
#pragma unroll
                                for (int k = 0; k < 4; k++)
                                    if (k < steps) {
                                        if (aTmem) {
                                            umma_ts<false>(d, a + aStep * k, bLo + 2 * k, descHi, idesc, (kc | k) != 0 ? 1u : 0u);
                                            umma_ts<false>(d, a + aLoOff + aStep * k, bHi + 2 * k, descHi, idesc, 1u);
                                        } else {
                                            umma_ss<false>(d, a + aStep * k, bLo + 2 * k, descHi, idesc, (kc | k) != 0 ? 1u : 0u);
                                            umma_ss<false>(d, a + aLoOff + aStep * k, bHi + 2 * k, descHi, idesc, 1u);
                                        }
                                    }
                                umma_commit(bEmpty(slotH));
                                umma_commit(bEmpty(slotL));
                            }
                            if (needOp) {
                                needOp = false;
                                // every slice of this layer's A operand is published, so all MMAs of the previous layer are complete: when
                                // that layer was G2, the activation buffer is free for the next chain's X
                                if (j == 5) mbar_arrive(xEmpty);
                            }
                            // pass 2: D = D * 2^-11 + sum_k Ahi_k . Bhi_k
                            for (int kc = 0; kc < ck; kc++) {
                                int slotH;
                                const uint32_t bHi = next_block(slotH);
                                const uint32_t a = aBase + kc * aChunk;
                                const int steps = kc == ck - 1 ? lastSteps : 4;
                                if (aTmem) {
                                    if (kc == 0) umma_ts<true>(d, a, bHi, descHi, idesc, 1u);
                                    else umma_ts<false>(d, a, bHi, descHi, idesc, 1u);
#pragma unroll
                                    for (int k = 1; k < 4; k++)
                                        if (k < steps) umma_ts<false>(d, a + aStep * k, bHi + 2 * k, descHi, idesc, 1u);
                                } else {
                                    if (kc == 0) umma_ss<true>(d, a, bHi, descHi, idesc, 1u);
                                    else umma_ss<false>(d, a, bHi, descHi, idesc, 1u);
#pragma unroll
                                    for (int k = 1; k < 4; k++)
                                        if (k < steps) umma_ss<false>(d, a + aStep * k, bHi + 2 * k, descHi, idesc, 1u);
                                }
                                umma_commit(bEmpty(slotH));
                                if (kc == ck - 1) {
                                    umma_commit(accFull(st));
                                    // G1 reads dZ0 from the TMEM operand columns the next chain's first epilogue overwrites: that epilogue waits
                                    // until the last G1 chunk has completed
                                    if (j == 5 && c == cn - 1) umma_commit(chainDone);
                                }
                            }
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp >= kFirstEpiWarp) {
        // ---- epilogue: all sixteen warps on every chunk; warp = (column group cg of 32 columns) x (lane quarter q) ----
        const int ew = warp - kFirstEpiWarp, q = warp & 3, cg = ew >> 2;
        const int r = q * 32 + lane;
        EpiCtx x;
        x.laneBase = tmemBase + ((uint32_t)(q * 32) << 16);
        x.sbufRow = sbuf + (uint32_t)r * 128;
        x.sLoOff = (uint32_t)P.sChunks * kTile;
        x.r7 = r & 7;
        x.cg = cg;
        float* const stash = P.stash + (size_t)blockIdx.x * (kStashCols * kRows) + r;
        uint32_t fullBits = 0, chainPhase = 0;
        int chunkIdx = 0;   // running chunk index: accumulator stage = its parity (issuer of the chunk)
        for (int u = blockIdx.x; u < P.numUnits; u += gridDim.x) {
            const int upt = P.M / P.mpu, t = u / upt, e0 = (u - t * upt) * P.mpu, e1 = e0 + P.mpu;
            const int si = species_of(t);
            const ChainSpecies& sp = P.sp[si];
            const int row0 = (t - sp.tileBegin) * kRows;
            x.rowOk = row0 + r < sp.rows;
            float* const dxRow = P.dX + (size_t)(sp.rowStart + row0 + r) * P.ldx;
            float esum = 0.0f;
            const unsigned char* const tab = chunkTab + si * 32;
            const int nChunks = tab[31];
            for (int e = e0; e < e1; e++) {
                // the previous chain's G1 must be done with the TMEM operand before F1 rewrites it
                mbar_wait(chainDone, chainPhase ^ 1u);
                chainPhase ^= 1u;
                for (int ci = 0; ci < nChunks; ci++) {          // chunks alternate between the two accumulator stages
                    const int j = tab[ci] >> 3, c = tab[ci] & 7, s = chunkIdx++ & 1;
                    int N, K;
                    layer_dims(sp, j, N, K);
                    x.n0 = c * 128 + cg * 32;
                    x.accCol = (uint32_t)s * kAccCols + (uint32_t)cg * 32;
                    const bool valid = x.n0 < N;             // warp-uniform: widths are multiples of 32
                    const uint32_t fullPhase = (fullBits >> s) & 1u;
                    fullBits ^= 1u << s;
                    if (valid) {
                        switch (j) {
                            case 0: epi_chunk<0>(P, x, sp.bias[0] + (size_t)e * N, nullptr, stash, dxRow, false, accFull(s), fullPhase, accEmpty(s), lane, esum); break;
                            case 1: epi_chunk<1>(P, x, sp.bias[1] + (size_t)e * N, nullptr, stash, dxRow, false, accFull(s), fullPhase, accEmpty(s), lane, esum); break;
                            case 2: epi_chunk<2>(P, x, sp.bias[2] + (size_t)e * N, sp.w3 + (size_t)e * N, stash, dxRow, false, accFull(s), fullPhase, accEmpty(s), lane, esum); break;
                            case 3: epi_chunk<3>(P, x, nullptr, nullptr, stash, dxRow, false, accFull(s), fullPhase, accEmpty(s), lane, esum); break;
                            case 4: epi_chunk<4>(P, x, nullptr, nullptr, stash, dxRow, false, accFull(s), fullPhase, accEmpty(s), lane, esum); break;
                            default: epi_chunk<5>(P, x, nullptr, nullptr, stash, dxRow, P.mpu == P.M && e == 0, accFull(s), fullPhase, accEmpty(s), lane, esum); break;
                        }
                    } else {
                        mbar_wait(accFull(s), fullPhase);
                        __syncwarp();
                        if (lane == 0) mbar_arrive(accEmpty(s));
                    }
                    // publish the 64-column slice of the next layer's A operand this warp has (or has not) contributed to -- when the
                    // next layer has such a slice at all (its K is this layer's N)
                    const int slice = x.n0 >> 6;
                    if (j < 5 && slice * 64 < N) {
                        if (j == 1 || j == 3) fence_proxy_async_smem();
                        else { tmem_st_wait(); tc_fence_before(); }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(opReady(slice));
                    }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) esum += __shfl_xor_sync(0xffffffffu, esum, o);
            if (lane == 0 && esum != 0.0f) atomicAdd(P.energyAcc, (double)esum);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmemBase), "r"(512u) : "memory");
    }
}

// weights of one GEMM as the TMA-friendly block matrix: for member e and 128-row chunk c, rows (e*cn + c)*256 + [0, 128) hold the
// fp16 high parts of output rows 128 c .. 128 c + 127 (zero beyond N) and rows + [128, 256) the scaled low parts; K columns, row pitch K
void pack_weights(std::vector<__half>& out, int M, int N, int K, const float* W, bool transposed, int ldw, int memberStride) {
    const int cn = chunks128(N);
    out.assign((size_t)M * cn * 256 * K, __float2half_rn(0.0f));
    for (int e = 0; e < M; e++)
        for (int n = 0; n < N; n++) {
            const int c = n >> 7, rr = n & 127;
            __half* hi = out.data() + ((size_t)(e * cn + c) * 256 + rr) * K;
            __half* lo = hi + (size_t)128 * K;
            for (int k = 0; k < K; k++) {
                // forward: W[e][n][k]; transposed (backward): the same layer's W[e][k][n]
                const float v = transposed ? W[(size_t)e * memberStride + (size_t)k * ldw + n] : W[(size_t)e * memberStride + (size_t)n * ldw + k];
                const __half h = __float2half_rn(v);
                hi[k] = h;
                lo[k] = __float2half_rn((v - __half2float(h)) * kLoScale);
            }
        }
}

}  // namespace

struct MlpChain2::Impl {
    ChainParams P;
    std::vector<__half*> dev;
    float* stash = nullptr;
    int grid = 0;
    long long rows = 0;
    uint32_t smem = 0;
};

bool MlpChain2::eligible(int numSpecies, const MlpChain::SpeciesDesc* sp, int featureStride) {
    if (numSpecies < 1 || numSpecies > kMaxSp) return false;
    int sMax = 0;
    for (int s = 0; s < numSpecies; s++) {
        const int* d = sp[s].d;
        if (d[0] != featureStride || d[0] > 128 || d[1] > 256 || d[2] > 256 || d[3] > 256) return false;
        for (int l = 0; l < 4; l++)
            if (d[l] <= 0 || d[l] % 32 != 0) return false;
        sMax = std::max(sMax, std::max(chunks64(d[2]), chunks64(d[0])));
    }
    const uint32_t fixed = 2u * sMax * kTile + 1024 + 512;
    return fixed + 4 * kTile <= kMaxSmem;   // a pass-1 pair of blocks in use plus a pair in flight
}

MlpChain2::MlpChain2(int ensemble, int numSpecies, const MlpChain::SpeciesDesc* sp, const __half* featHi, const __half* featLo, int featureStride)
    : impl_(new Impl) {
    NNP_REQUIRE(eligible(numSpecies, sp, featureStride), "MlpChain2: network shape not supported by the fused kernel");
    ChainParams& P = impl_->P;
    std::memset(&P, 0, sizeof(P));
    P.numSpecies = numSpecies; P.M = ensemble; P.ldx = featureStride;
    int tiles = 0;
    for (int s = 0; s < numSpecies; s++) {
        ChainSpecies& c = P.sp[s];
        c.d0 = sp[s].d[0]; c.d1 = sp[s].d[1]; c.d2 = sp[s].d[2]; c.d3 = sp[s].d[3];
        c.rows = sp[s].rows; c.rowStart = sp[s].rowStart;
        c.tileBegin = tiles;
        tiles += (c.rows + kRows - 1) / kRows;
        c.tileEnd = tiles;
        impl_->rows = std::max<long long>(impl_->rows, (long long)c.rowStart + c.rows);
        for (int l = 0; l < 3; l++) c.bias[l] = sp[s].bias[l];
        c.w3 = sp[s].w3;
        P.sChunks = std::max(P.sChunks, std::max(chunks64(c.d2), chunks64(c.d0)));
        // GEMM j: N, K, layer, transposed
        const int gN[6] = {c.d1, c.d2, c.d3, c.d2, c.d1, c.d0}, gK[6] = {c.d0, c.d1, c.d2, c.d3, c.d2, c.d1};
        const int gL[6] = {0, 1, 2, 2, 1, 0};
        for (int j = 0; j < 6; j++) {
            const int l = gL[j], in = sp[s].d[l], out = sp[s].d[l + 1];
            std::vector<__half> host;
            pack_weights(host, ensemble, gN[j], gK[j], sp[s].W[l], j >= 3, in, out * in);
            __half* d = nullptr;
            NNP_CUDA_CHECK(cudaMalloc(&d, sizeof(__half) * host.size()));
            NNP_CUDA_CHECK(cudaMemcpy(d, host.data(), sizeof(__half) * host.size(), cudaMemcpyHostToDevice));
            impl_->dev.push_back(d);
            P.maps[s][j] = tc_make_map(d, (long long)ensemble * chunks128(gN[j]) * 256, gK[j], gK[j]);
        }
        if (c.rows > 0) {
            P.maps[s][6] = tc_make_map(featHi + (size_t)c.rowStart * featureStride, c.rows, featureStride, featureStride);
            P.maps[s][7] = tc_make_map(featLo + (size_t)c.rowStart * featureStride, c.rows, featureStride, featureStride);
        } else {
            P.maps[s][6] = P.maps[s][0]; P.maps[s][7] = P.maps[s][0];
        }
    }
    P.numTiles = tiles;
    const uint32_t fixed = 2u * P.sChunks * kTile + 1024 + 512;
    P.ring = (int)std::min<uint32_t>(kMaxRing, (kMaxSmem - fixed) / kTile);
    impl_->smem = fixed + P.ring * kTile;
    int dev = 0, sms = 0;
    NNP_CUDA_CHECK(cudaGetDevice(&dev));
    NNP_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // work unit = (tile, mpu consecutive members): members of a tile are dealt to different CTAs until there are at least eight units
    // per SM (see mlp_chain.cu)
    int mpu = ensemble;
    while (mpu > 1 && (long long)tiles * (ensemble / mpu) < 8LL * sms) {
        int next = mpu - 1;
        while (next > 1 && ensemble % next != 0) next--;
        mpu = next;
    }
    if (const char* e = std::getenv("NNPOPS_CHAIN_MPU")) {   // development: force the members per unit (must divide the ensemble size)
        const int v = std::atoi(e);
        if (v >= 1 && v <= ensemble && ensemble % v == 0) mpu = v;
    }
    P.mpu = mpu;
    P.numUnits = tiles * (ensemble / mpu);
    impl_->grid = std::max(1, std::min(P.numUnits, sms));
    if (const char* e = std::getenv("NNPOPS_CHAIN_GRID")) impl_->grid = std::max(1, std::min(impl_->grid, std::atoi(e)));   // development: several units per CTA on small systems
    NNP_CUDA_CHECK(cudaMalloc(&impl_->stash, sizeof(float) * (size_t)impl_->grid * kStashCols * kRows));
    P.stash = impl_->stash;
}

MlpChain2::~MlpChain2() {
    for (__half* p : impl_->dev) cudaFree(p);
    cudaFree(impl_->stash);
    delete impl_;
}

void MlpChain2::launch(double* energyAcc, float* dX, float seedScale, float outScale, cudaStream_t stream) {
    ChainParams& P = impl_->P;
    if (P.numTiles == 0) return;
    P.energyAcc = energyAcc; P.dX = dX; P.seedScale = seedScale; P.outScale = outScale;
    if (P.mpu != P.M) NNP_CUDA_CHECK(cudaMemsetAsync(dX, 0, sizeof(float) * (size_t)impl_->rows * P.ldx, stream));   // all members accumulate with red.add
    // per device, every time: the attribute is cheap to set and a process may drive several GPUs
    NNP_CUDA_CHECK(cudaFuncSetAttribute(mlp_chain2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)impl_->smem));
    mlp_chain2_kernel<<<impl_->grid, kThreads, impl_->smem, stream>>>(P);
    NNP_CUDA_CHECK(cudaGetLastError());
    count_launch();
}

}  // namespace nnpops
