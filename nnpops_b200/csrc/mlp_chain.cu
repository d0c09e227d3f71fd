// Fused per-tile layer chain of the species MLP on tcgen05 (see mlp_chain.cuh).
//
// One persistent CTA per SM walks over tiles of 128 atoms of one species and, per tile, over the M ensemble members; for each
// (tile, member) it runs the six GEMMs of the network back to back ("chain"):
//     F1  A1  = celu(X   W0^T + b0)        N = d1, K = d0     A operand: X    (shared memory, TMA, re-fetched per chain)
//     F2  A2  = celu(A1  W1^T + b1)        N = d2, K = d1     A operand: A1   (tensor memory)
//     F3  a3  = celu(A2  W2^T + b2)        N = d3, K = d2     A operand: A2   (shared memory); E += a3 . w3; dZ2 = w3/M celu'
//     G3  dZ1 = (dZ2 W2) * celu'(A2)       N = d2, K = d3     A operand: dZ2  (tensor memory); overwrites A2 in place
//     G2  dZ0 = (dZ1 W1) * celu'(A1)       N = d1, K = d2     A operand: dZ1  (shared memory)
//     G1  dX += dZ0 W0                     N = d0, K = d1     A operand: dZ0  (tensor memory); st / red.add to global
// Every operand is an fp16 hi/lo pair (mlp_tcgen05.cu: x = hi + 2^-11 lo, three MMAs per product, fp32 accumulation), so the
// result carries ~22 mantissa bits.  Activations alternate between tensor memory and shared memory because neither holds two
// consecutive ones: TMEM = 2 accumulator stages x 128 columns + 256 columns of packed fp16 A operand (K <= 256); shared memory
// = one activation buffer (<= 96 KB: X during F1, A2 from F2 to G3, dZ1 from G3 to G2; X is re-fetched per chain from L2 under G1 of the
// previous chain, which buys 64 KB for the weight ring) + a ring of eight 16 KB weight blocks.  The only activation that does not
// fit, A1 (needed again for celu' in G2), goes through a per-CTA fp32 scratch that stays in L2 (128 KB per CTA, rewritten every chain).
//
// Work split inside the CTA (20 warps, 640 threads):
//   warps 0, 3    TMA producers (one elected lane each): warp 0 feeds MMA issuer 0 and loads X once per chain, warp 3 feeds issuer 1;
//                 one 16 KB weight block [64 hi rows + 64 lo rows][64 k] per (n-chunk, k-chunk) into the issuer's half of the ring
//   warps 1, 2    MMA issuers (one elected lane each; warp 1 also allocates the tensor memory): issuer g owns accumulator stage g and
//                 the chunks of a chain with parity g; per k16 step  [D1 | D2] (+)= Ahi . [Bhi; Blo]  (N = 128) and D2 += Alo . Bhi
//                 (N = 64); tcgen05.commit frees ring slots and publishes accumulator stages
//   warps 4-11    epilogue group 0 (accumulator stage 0),  warps 12-19 epilogue group 1 (stage 1): 64-column chunks alternate
//                 between the groups, so the epilogue of chunk c overlaps the MMAs of chunk c + 1.  A thread owns one row (TMEM
//                 lane) and 32 columns; it pulls its accumulators into registers, hands the stage back, and writes the next layer's
//                 A operand straight into tensor memory (tcgen05.st) or into the swizzled K-major shared-memory tiles, then arrives
//                 on the per-64-column "operand ready" barrier the MMA warps wait on before they issue the first k-chunk that needs
//                 those columns.
// What bounds it (measured with the CHAIN_EXP_* / CHAIN_TRACE switches below): DESIGN.md section 4, profiles/r08_summary.md.
#include "mlp_chain.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include "tcgen05_util.cuh"

namespace nnpops {

namespace {
using namespace tc;

constexpr int kMaxSp = 7;
constexpr int kRows = 128;
constexpr uint32_t kTile = 128 * 128;        // 16 KB: [128 rows][64 halves] A tile, or [64 hi + 64 lo rows][64 halves] B block
constexpr int kMaxRing = 8;
constexpr int kFirstEpiWarp = 4, kGroupWarps = 8;   // warps 0..3: TMA producer, two MMA issuers, one idle (keeps warp % 4 = TMEM lane quarter simple)
constexpr int kThreads = (kFirstEpiWarp + 2 * kGroupWarps) * 32;
constexpr uint32_t kAccCols = 128;           // one accumulator stage: D1 (64 columns) | D2 (64 columns)
constexpr uint32_t kOpaHi = 256, kOpaLo = 384;   // A operand in TMEM: packed fp16 pairs, hi part and lo part (128 columns each: K <= 256)
constexpr int kStashCols = 256;
constexpr uint32_t kMaxSmem = 232448;

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
    int mpu, numUnits;             // work unit = (tile, mpu consecutive ensemble members); numUnits = numTiles * (M / mpu)
    int xChunks, sChunks, ring;    // 64-column chunks of X and of the shared-memory activation buffer; weight ring depth
    float* dX;
    int ldx;
    float* stash;                  // [grid][256 columns][128 rows] fp32
    double* energyAcc;
    float seedScale, outScale;
};

__host__ __device__ constexpr int chunks_of(int n) { return (n + 63) >> 6; }

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

#ifdef CHAIN_TRACE
// development: per-role event log of CTA 0 (role 0/1 = MMA issuers, 2/3 = lane 0 of the first warp of each epilogue group); entry =
// tag << 40 | clock.  Printed by the kernel itself when it ends.
constexpr int kTraceLen = 1024;
__device__ long long gTrace[4][kTraceLen];
__device__ int gTraceN[4];
__device__ __forceinline__ void trace_ev(int role, int tag) {
    if (blockIdx.x != 0) return;
    const int n = gTraceN[role];
    if (n < kTraceLen) { gTrace[role][n] = ((long long)tag << 40) | (clock64() & 0xffffffffffLL); gTraceN[role] = n + 1; }
}
#define TRACE(role, tag) trace_ev(role, tag)
#else
#define TRACE(role, tag)
#endif

// tcgen05.mma with the shared-memory descriptors given as low words (start address | LBO) + one common high word: the issue loop
// then runs on 32-bit adds only
__device__ __forceinline__ void umma_ss_lo(uint32_t tmemD, uint32_t aLo, uint32_t bLo, uint32_t descHi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 da, db;\n"
        "setp.ne.b32 p, %5, 0;\n"
        "mov.b64 da, {%1, %3};\n"
        "mov.b64 db, {%2, %3};\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n"
        "}\n" ::"r"(tmemD), "r"(aLo), "r"(bLo), "r"(descHi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_ts_lo(uint32_t tmemD, uint32_t tmemA, uint32_t bLo, uint32_t descHi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 db;\n"
        "setp.ne.b32 p, %5, 0;\n"
        "mov.b64 db, {%2, %3};\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n"
        "}\n" ::"r"(tmemD), "r"(tmemA), "r"(bLo), "r"(descHi), "r"(idesc), "r"(accumulate) : "memory");
}

// what an epilogue thread needs to know about its chunk
struct EpiCtx {
    uint32_t laneBase;      // TMEM address of the thread's lane, column 0
    uint32_t accCol;        // first column of the accumulator stage (+ the warp's 32-column half)
    uint32_t sbufRow;       // shared-memory address of the thread's row in chunk 0 of the activation buffer (hi part)
    uint32_t sLoOff;        // byte offset of the lo tiles of that buffer
    int r7;                 // row & 7 (swizzle phase)
    int hsel, n0, c;
    int trace;              // development (CHAIN_TRACE): event-log role of this thread's warp, or -1
    bool rowOk;
};

// packed fp32x2 arithmetic (FFMA2 / FMUL2 / FADD2: one issue slot for two columns).  The epilogue warps share their schedulers with the
// MMA issuers and are issue-bound, so every instruction saved here shortens the phases in which the tensor pipe waits for an operand.
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

// TYPE: 0 F1, 1 F2, 2 F3, 3 G3, 4 G2, 5 G1.  Processes the thread's 32 columns [n0, n0 + 32) of chunk c: the accumulator pair (D1, D2) is
// pulled into registers as 16 packed column pairs and the stage is handed back to the MMA warp BEFORE any arithmetic.
template <int TYPE>
__device__ __forceinline__ void epi_chunk(const ChainParams& P, const EpiCtx& x, const float* __restrict__ colA, const float* __restrict__ colB,
                                          float* stash, float* dxRow, bool firstMember, uint32_t accFullBar, uint32_t fullPhase, uint32_t accEmptyBar, int lane,
                                          float& esum) {
    // fetched before the accumulator is needed (with 224 KB of shared memory in use the L1 is small: a load after the wait may cost an
    // L2 trip): the biases of the thread's columns (every lane reads the same 16-byte words: broadcast), or -- G2 -- A1 of these columns,
    // written to the stash by F1's epilogue of this chain
    u64 pre[(TYPE <= 2 || TYPE == 4) ? 16 : 1];
#ifndef CHAIN_EXP_EPI_LITE
    if (TYPE <= 2) {
        const ulonglong2* src = reinterpret_cast<const ulonglong2*>(colA + x.n0);
#pragma unroll
        for (int i = 0; i < 8; i++) { const ulonglong2 t = __ldg(src + i); pre[2 * i] = t.x; pre[2 * i + 1] = t.y; }
    }
    if (TYPE == 4) {
#pragma unroll
        for (int i = 0; i < 16; i++) pre[i] = pk2(__ldcg(stash + (size_t)(x.n0 + 2 * i) * kRows), __ldcg(stash + (size_t)(x.n0 + 2 * i + 1) * kRows));
    }
#endif
    if (x.trace >= 0 && lane == 0) TRACE(x.trace, 0x600 | TYPE << 4 | x.c);
    mbar_wait(accFullBar, fullPhase);
    tc_fence_after();
    if (x.trace >= 0 && lane == 0) TRACE(x.trace, 0x700 | TYPE << 4 | x.c);
#ifdef CHAIN_EXP_EPI_NONE
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(accEmptyBar);
    return;
#endif
    u64 vv[16];
    const u64 loInv2 = pk2(kLoInv, kLoInv);
#pragma unroll
    for (int h = 0; h < 2; h++) {
        uint32_t r1[16], r2[16];
        tmem_ld16(x.laneBase + x.accCol + 16 * h, r1);
        tmem_ld16(x.laneBase + x.accCol + 64 + 16 * h, r2);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; i++) vv[8 * h + i] = ffma2(pk2u(r2[2 * i], r2[2 * i + 1]), loInv2, pk2u(r1[2 * i], r1[2 * i + 1]));
    }
    // the accumulator stage has been read completely: hand it back to the MMA warp
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(accEmptyBar);
    if (x.trace >= 0 && lane == 0) TRACE(x.trace, 0x800 | TYPE << 4 | x.c);

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
#ifdef CHAIN_EXP_EPI_LITE
#pragma unroll
    for (int h = 0; h < 2; h++) {
        uint32_t ph[8], pl[8];
#pragma unroll
        for (int i = 0; i < 8; i++) { float a, b; unpk2(vv[8 * h + i], a, b); ph[i] = pack_h2(a, b); pl[i] = 0; }
        const int u0 = x.hsel * 4 + 2 * h;
        const uint32_t sA = x.sbufRow + (uint32_t)x.c * kTile + (uint32_t)((u0 ^ x.r7) << 4);
        const uint32_t sB = x.sbufRow + (uint32_t)x.c * kTile + (uint32_t)(((u0 + 1) ^ x.r7) << 4);
        if (TYPE == 1 || TYPE == 3) {
            st_shared_v4(sA, ph[0], ph[1], ph[2], ph[3]);
            st_shared_v4(sB, ph[4], ph[5], ph[6], ph[7]);
            st_shared_v4(sA + x.sLoOff, pl[0], pl[1], pl[2], pl[3]);
            st_shared_v4(sB + x.sLoOff, pl[4], pl[5], pl[6], pl[7]);
        } else {
            tmem_st8(x.laneBase + kOpaHi + (uint32_t)((x.n0 + 16 * h) >> 1), ph);
            tmem_st8(x.laneBase + kOpaLo + (uint32_t)((x.n0 + 16 * h) >> 1), pl);
        }
    }
#else
    const u64 celuScale2 = pk2(1.4426950408889634f / kCeluAlpha, 1.4426950408889634f / kCeluAlpha);
    const u64 alpha2 = pk2(kCeluAlpha, kCeluAlpha), negAlpha2 = pk2(-kCeluAlpha, -kCeluAlpha);
    const u64 invAlpha2 = pk2(1.0f / kCeluAlpha, 1.0f / kCeluAlpha), one2 = pk2(1.0f, 1.0f);
    u64 esum2 = pk2(0.0f, 0.0f);
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int col = x.n0 + 16 * h;
        // where the two 16-byte units of this half live in the thread's row of the swizzled shared-memory tile
        const int u0 = x.hsel * 4 + 2 * h;
        const uint32_t sA = x.sbufRow + (uint32_t)x.c * kTile + (uint32_t)((u0 ^ x.r7) << 4);
        const uint32_t sB = x.sbufRow + (uint32_t)x.c * kTile + (uint32_t)(((u0 + 1) ^ x.r7) << 4);
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
                float g0, g1;
                unpk2(ffma2(act2, invAlpha2, one2), g0, g1);   // celu'(z) from celu(z): 1 for a > 0, a / alpha + 1 (<= 1) otherwise
                v = fmul2(v, pk2(fminf(g0, 1.0f), fminf(g1, 1.0f)));
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
#endif
}

__global__ void __launch_bounds__(kThreads, 1) mlp_chain_kernel(const __grid_constant__ ChainParams P) {
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
    const uint32_t chainDone = xEmpty + 8u;   // both issuers' MMAs of a chain have completed (its last layer reads the TMEM operand)
    const uint32_t tmemSlot = chainDone + 8u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < P.numSpecies; s++)
            for (int m = 0; m < 8; m++) asm volatile("prefetch.tensormap [%0];" ::"l"(&P.maps[s][m]) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < P.ring; s++) { mbar_init(bFull(s), 1); mbar_init(bEmpty(s), 1); }
        for (int s = 0; s < 2; s++) { mbar_init(accFull(s), 1); mbar_init(accEmpty(s), kGroupWarps); }
        for (int s = 0; s < 4; s++) mbar_init(opReady(s), kGroupWarps);
        mbar_init(xFull, 1);
        mbar_init(xEmpty, 2);   // both MMA issuers
        mbar_init(chainDone, 2);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // flat chunk schedule of one chain per species (entry = layer << 3 | chunk): keeps the epilogue's loop state down to a counter
    unsigned char* const chunkTab = smemRaw + (barBase - rawAddr) + 256;
    if (warp == 2 && lane < P.numSpecies) {
        int n = 0;
        for (int j = 0; j < 6; j++) {
            int N, K;
            layer_dims(P.sp[lane], j, N, K);
            for (int c = 0; c < chunks_of(N); c++) chunkTab[lane * 32 + n++] = (unsigned char)(j << 3 | c);
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

    if (warp == 0 || warp == 3) {
        // ---- TMA producers: warp 0 feeds issuer 0 (even chunks), warp 3 issuer 1 (odd chunks), each through its own half of the weight
        // ring.  One ring per issuer, because an mbarrier wait only knows the PARITY of a phase: a consumer that is allowed to wait
        // for fill m + 1 of a slot before the other consumer has seen fill m reads "done" from the stale parity.  Warp 0 also loads X.
        if (elect_one()) {
            const int g = warp == 0 ? 0 : 1;
            const int half = P.ring >> 1, s0 = g * half;
            int stage = 0;
            uint32_t phase = 0, xPhase = 0;
            for (int u = blockIdx.x; u < P.numUnits; u += gridDim.x) {
                const int upt = P.M / P.mpu, t = u / upt, e0 = (u - t * upt) * P.mpu, e1 = e0 + P.mpu;
                const int si = species_of(t);
                const ChainSpecies& sp = P.sp[si];
                for (int e = e0; e < e1; e++) {
                    if (g == 0) {
                        // X shares the activation buffer: it is (re)loaded for every chain, as soon as G2 of the previous chain -- the
                        // last reader of that buffer -- has completed (both issuers arrive on xEmpty when they have seen G1's operand)
                        const int row0 = (t - sp.tileBegin) * kRows;
                        const int cx = chunks_of(sp.d0);
                        mbar_wait(xEmpty, xPhase ^ 1u);
                        xPhase ^= 1u;
                        mbar_expect_tx(xFull, 2u * cx * kTile);
                        for (int kc = 0; kc < cx; kc++) {
                            tma_load_2d(sbuf + kc * kTile, &P.maps[si][6], xFull, kc * 64, row0);
                            tma_load_2d(sbuf + (P.sChunks + kc) * kTile, &P.maps[si][7], xFull, kc * 64, row0);
                        }
                    }
                    int chunkIdx = 0;
                    for (int j = 0; j < 6; j++) {
                        int N, K;
                        layer_dims(sp, j, N, K);
                        const int cn = chunks_of(N), ck = chunks_of(K);
                        for (int c = 0; c < cn; c++) {
                            if ((chunkIdx++ & 1) != g) continue;
                            for (int kc = 0; kc < ck; kc++) {
                                mbar_wait(bEmpty(s0 + stage), phase ^ 1u);
                                mbar_expect_tx(bFull(s0 + stage), kTile);
                                tma_load_2d(ring + (s0 + stage) * kTile, &P.maps[si][j], bFull(s0 + stage), kc * 64, (e * cn + c) * 128);
                                if (++stage == half) { stage = 0; phase ^= 1u; }
                            }
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1 || warp == 2) {
        // ---- MMA issuers: warp 1 owns accumulator stage 0 (even chunks of a chain), warp 2 stage 1 (odd chunks); in each ONE elected
        // thread runs the whole loop.  Two issuers because a single thread cannot get the ~60 instructions of a 16 KB weight block
        // (8 small MMAs + barrier traffic) out in the 384 cycles the tensor pipe needs for it.  Ordering between the two is carried
        // by the data: a chunk of layer j is issued only after every 64-column slice of its A operand has been published (opReady),
        // i.e. after all MMAs of layer j - 1 -- from both issuers -- have completed and been read by the epilogue.
        if (elect_one()) {
            const int g = warp - 1;
            constexpr uint32_t idescWide = idesc_f16(128, 128), idescNarrow = idesc_f16(128, 64);
            const uint32_t descHi = (uint32_t)(make_desc(0) >> 32);
            auto desc_lo = [](uint32_t addr) { return ((addr >> 4) & 0x3fffu) | (1u << 16); };   // start address | LBO = 1
            constexpr uint32_t kTile16 = kTile >> 4;
            const int half = P.ring >> 1, s0 = g * half;   // this issuer's half of the weight ring
            const uint32_t ringLo = desc_lo(ring) + s0 * kTile16, sLo = desc_lo(sbuf);
            const uint32_t d = tmemBase + g * kAccCols;
            const uint32_t accFullBar = accFull(g), accEmptyBar = accEmpty(g);
            int stage = 0;
            uint32_t phase = 0, accPhase = 0, opBits = 0, xPhase = 0;
            for (int u = blockIdx.x; u < P.numUnits; u += gridDim.x) {
                const int upt = P.M / P.mpu, t = u / upt, e0 = (u - t * upt) * P.mpu, e1 = e0 + P.mpu;
                const int si = species_of(t);
                const ChainSpecies& sp = P.sp[si];
                for (int e = e0; e < e1; e++) {
                    mbar_wait(xFull, xPhase);
                    xPhase ^= 1u;
                    tc_fence_after();
                    int chunkIdx = 0;
                    for (int j = 0; j < 6; j++) {
                        int N, K;
                        layer_dims(sp, j, N, K);
                        const int cn = chunks_of(N), ck = chunks_of(K);
                        const int lastSteps = (K - (ck - 1) * 64) >> 4;
                        const bool aTmem = (j & 1) != 0;                 // F2, G3, G1 read their A operand from tensor memory
                        // A operand: address of k-chunk 0 (TMEM column, or descriptor low word), stride per k-chunk, offset of the lo part
                        const uint32_t aBase = aTmem ? tmemBase + kOpaHi : sLo;
                        const uint32_t aChunk = aTmem ? 32u : kTile16;
                        const uint32_t aLoOff = aTmem ? (kOpaLo - kOpaHi) : (uint32_t)P.sChunks * kTile16;
                        bool needOp = j > 0;   // the A operand of this layer has not been waited for yet (by this issuer)
                        bool chainCommitted = false;
                        for (int c = 0; c < cn; c++) {
                            const int acc = chunkIdx & 1;
                            chunkIdx++;
                            if (acc != g) continue;   // the other issuer's chunk
                            TRACE(g, 0x100 | j << 4 | c);
                            mbar_wait(accEmptyBar, accPhase ^ 1u);
                            accPhase ^= 1u;
                            tc_fence_after();
                            TRACE(g, 0x200 | j << 4 | c);
                            for (int kc = 0; kc < ck; kc++) {
                                if (needOp) {   // columns [64 kc, 64 kc + 64) of the A operand come from the previous layer's epilogue
                                    mbar_wait(opReady(kc), (opBits >> kc) & 1u);
                                    opBits ^= 1u << kc;
                                    TRACE(g, 0x300 | j << 4 | kc);
                                }
                                mbar_wait(bFull(s0 + stage), phase);
                                tc_fence_after();
                                TRACE(g, 0x400 | j << 4 | kc);
                                const uint32_t bLo = ringLo + stage * kTile16;
                                const uint32_t a = aBase + kc * aChunk;
#ifdef CHAIN_EXP_MMA_LITE
                                const int steps = 1;
#else
                                const int steps = kc == ck - 1 ? lastSteps : 4;
#endif
                                if (aTmem) {
#pragma unroll
                                    for (int k = 0; k < 4; k++)
                                        if (k < steps) {
                                            umma_ts_lo(d, a + 8 * k, bLo + 2 * k, descHi, idescWide, (kc | k) != 0 ? 1u : 0u);
                                            umma_ts_lo(d + 64, a + aLoOff + 8 * k, bLo + 2 * k, descHi, idescNarrow, 1u);
                                        }
                                } else {
#pragma unroll
                                    for (int k = 0; k < 4; k++)
                                        if (k < steps) {
                                            umma_ss_lo(d, a + 2 * k, bLo + 2 * k, descHi, idescWide, (kc | k) != 0 ? 1u : 0u);
                                            umma_ss_lo(d + 64, a + aLoOff + 2 * k, bLo + 2 * k, descHi, idescNarrow, 1u);
                                        }
                                }
                                umma_commit(bEmpty(s0 + stage));
                                if (kc == ck - 1) {
                                    TRACE(g, 0x500 | j << 4 | c);
                                    umma_commit(accFullBar);
                                    // G1 reads dZ0 from the TMEM operand columns the next chain's first epilogue overwrites: that epilogue waits
                                    // until the last G1 chunk of BOTH issuers has completed
                                    if (j == 5 && c >= cn - 2) { umma_commit(chainDone); chainCommitted = true; }
                                }
                                if (++stage == half) { stage = 0; phase ^= 1u; }
                            }
                            if (needOp) {
                                needOp = false;
                                // every slice of this layer's A operand is published, so all MMAs of the previous layer are complete: when
                                // that layer was G2, the activation buffer is free for the next chain's X
                                if (j == 5) mbar_arrive(xEmpty);
                            }
                        }
                        if (needOp) {
                            // no chunk of this layer was ours.  The operand barriers must still be OBSERVED phase by phase: a wait only
                            // knows the parity of a phase, so an issuer that skipped one would take "two phases back" for "done" at its
                            // next real wait and run ahead of the epilogue.  (xEmpty counts both issuers.)
                            for (int kc = 0; kc < ck; kc++) {
                                mbar_wait(opReady(kc), (opBits >> kc) & 1u);
                                opBits ^= 1u << kc;
                            }
                            if (j == 5) mbar_arrive(xEmpty);
                        }
                        if (j == 5 && !chainCommitted) mbar_arrive(chainDone);   // G1 had a single chunk and it was the other issuer's
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp >= kFirstEpiWarp) {
        // ---- epilogue groups ----
        const int ew = warp - kFirstEpiWarp, g = ew >> 3, q = warp & 3, hsel = (ew >> 2) & 1;
        const int r = q * 32 + lane;
        EpiCtx x;
        x.laneBase = tmemBase + ((uint32_t)(q * 32) << 16);
        x.accCol = (uint32_t)g * kAccCols + (uint32_t)hsel * 32;
        x.sbufRow = sbuf + (uint32_t)r * 128;
        x.sLoOff = (uint32_t)P.sChunks * kTile;
        x.r7 = r & 7;
        x.hsel = hsel;
        x.trace = (ew & 7) == 0 ? 2 + g : -1;
        float* const stash = P.stash + (size_t)blockIdx.x * (kStashCols * kRows) + r;
        uint32_t fullPhase = 0, chainPhase = 0;
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
                bool chainWaited = false;
                for (int ci = g; ci < nChunks; ci += 2) {          // chunks alternate between the two accumulator stages / groups
                    const int j = tab[ci] >> 3, c = tab[ci] & 7;
                    int N, K;
                    layer_dims(sp, j, N, K);
                    {
                        x.c = c;
                        x.n0 = c * 64 + hsel * 32;
                        const bool valid = x.n0 < N;             // warp-uniform: widths are multiples of 32
                        if (!chainWaited) {   // the previous chain's G1 (both issuers) must be done with the TMEM operand before F1 rewrites it
                            mbar_wait(chainDone, chainPhase ^ 1u);
                            chainPhase ^= 1u;
                            chainWaited = true;
                        }
                        if (valid) {
                            switch (j) {
                                case 0: epi_chunk<0>(P, x, sp.bias[0] + (size_t)e * N, nullptr, stash, dxRow, false, accFull(g), fullPhase, accEmpty(g), lane, esum); break;
                                case 1: epi_chunk<1>(P, x, sp.bias[1] + (size_t)e * N, nullptr, stash, dxRow, false, accFull(g), fullPhase, accEmpty(g), lane, esum); break;
                                case 2: epi_chunk<2>(P, x, sp.bias[2] + (size_t)e * N, sp.w3 + (size_t)e * N, stash, dxRow, false, accFull(g), fullPhase, accEmpty(g), lane, esum); break;
                                case 3: epi_chunk<3>(P, x, nullptr, nullptr, stash, dxRow, false, accFull(g), fullPhase, accEmpty(g), lane, esum); break;
                                case 4: epi_chunk<4>(P, x, nullptr, nullptr, stash, dxRow, false, accFull(g), fullPhase, accEmpty(g), lane, esum); break;
                                default: epi_chunk<5>(P, x, nullptr, nullptr, stash, dxRow, P.mpu == P.M && e == 0, accFull(g), fullPhase, accEmpty(g), lane, esum); break;
                            }
                        } else {
                            mbar_wait(accFull(g), fullPhase);
                            __syncwarp();
                            if (lane == 0) mbar_arrive(accEmpty(g));
                        }
                        fullPhase ^= 1u;
                        if (j < 5) {
                            // publish the 64-column slice of the next layer's A operand this warp has (or has not) contributed to
                            if (j == 1 || j == 3) fence_proxy_async_smem();
                            else { tmem_st_wait(); tc_fence_before(); }
                            __syncwarp();
                            if (lane == 0) mbar_arrive(opReady(c));
                            if (x.trace >= 0 && lane == 0) TRACE(x.trace, 0x900 | j << 4 | c);
                        }
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
#ifdef CHAIN_TRACE
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int r = 0; r < 4; r++) {
            for (int i = 0; i < gTraceN[r]; i++) printf("T %d %x %lld\n", r, (int)(gTrace[r][i] >> 40), gTrace[r][i] & 0xffffffffffLL);
            gTraceN[r] = 0;
        }
    }
#endif
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmemBase), "r"(512u) : "memory");
    }
}

// weights of one GEMM as the TMA-friendly block matrix: for member e and 64-row chunk c, rows (e*cn + c)*128 + [0, 64) hold the fp16
// high parts of output rows 64 c .. 64 c + 63 (zero beyond N) and rows + [64, 128) the scaled low parts; K columns, row pitch K
void pack_weights(std::vector<__half>& out, int M, int N, int K, const float* W, bool transposed, int ldw, int memberStride) {
    const int cn = chunks_of(N);
    out.assign((size_t)M * cn * 128 * K, __float2half_rn(0.0f));
    for (int e = 0; e < M; e++)
        for (int n = 0; n < N; n++) {
            const int c = n >> 6, rr = n & 63;
            __half* hi = out.data() + ((size_t)(e * cn + c) * 128 + rr) * K;
            __half* lo = hi + (size_t)64 * K;
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

struct MlpChain::Impl {
    ChainParams P;
    std::vector<__half*> dev;
    float* stash = nullptr;
    int grid = 0;
    long long rows = 0;
    uint32_t smem = 0;
};

bool MlpChain::eligible(int numSpecies, const SpeciesDesc* sp, int featureStride) {
    if (numSpecies < 1 || numSpecies > kMaxSp) return false;
    int sMax = 0;
    for (int s = 0; s < numSpecies; s++) {
        const int* d = sp[s].d;
        if (d[0] != featureStride || d[0] > 128 || d[1] > 256 || d[2] > 256 || d[3] > 256) return false;
        for (int l = 0; l < 4; l++)
            if (d[l] <= 0 || d[l] % 32 != 0) return false;
        sMax = std::max(sMax, std::max(chunks_of(d[2]), chunks_of(d[0])));
    }
    const uint32_t fixed = 2u * sMax * kTile + 1024 + 512;
    return fixed + 4 * kTile <= kMaxSmem;
}

MlpChain::MlpChain(int ensemble, int numSpecies, const SpeciesDesc* sp, const __half* featHi, const __half* featLo, int featureStride)
    : impl_(new Impl) {
    NNP_REQUIRE(eligible(numSpecies, sp, featureStride), "MlpChain: network shape not supported by the fused kernel");
    ChainParams& P = impl_->P;
    std::memset(&P, 0, sizeof(P));
    P.numSpecies = numSpecies; P.M = ensemble; P.ldx = featureStride;
    P.xChunks = chunks_of(featureStride);
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
        P.sChunks = std::max(P.sChunks, std::max(chunks_of(c.d2), chunks_of(c.d0)));
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
            P.maps[s][j] = tc_make_map(d, (long long)ensemble * chunks_of(gN[j]) * 128, gK[j], gK[j]);
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
    P.ring = (int)std::min<uint32_t>(kMaxRing, (kMaxSmem - fixed) / kTile) & ~1;   // an even number of slots: one half per MMA issuer
    if (const char* e = std::getenv("NNPOPS_CHAIN_RING")) P.ring = std::max(2, std::min(P.ring, std::atoi(e))) & ~1;   // development: ring depth sensitivity
    impl_->smem = fixed + P.ring * kTile;
    int dev = 0, sms = 0;
    NNP_CUDA_CHECK(cudaGetDevice(&dev));
    NNP_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // Work unit = (tile, mpu consecutive members).  Whole tiles (mpu = M: X loaded once, dX written without atomics by the first member)
    // leave the SMs unevenly loaded when there are only a few tiles per SM -- 392 tiles of two costs on 148 SMs: 13 % -- so the members
    // of a tile are dealt to different CTAs until there are at least eight units per SM (each unit then reloads its 64 KB X tile from
    // L2 and adds its dX contribution with red.global.add on a zeroed matrix).
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

MlpChain::~MlpChain() {
    for (__half* p : impl_->dev) cudaFree(p);
    cudaFree(impl_->stash);
    delete impl_;
}

void MlpChain::launch(double* energyAcc, float* dX, float seedScale, float outScale, cudaStream_t stream) {
    ChainParams& P = impl_->P;
    if (P.numTiles == 0) return;
    P.energyAcc = energyAcc; P.dX = dX; P.seedScale = seedScale; P.outScale = outScale;
    if (P.mpu != P.M) NNP_CUDA_CHECK(cudaMemsetAsync(dX, 0, sizeof(float) * (size_t)impl_->rows * P.ldx, stream));   // all members accumulate with red.add
    // per device, every time: the attribute is cheap to set and a process may drive several GPUs
    NNP_CUDA_CHECK(cudaFuncSetAttribute(mlp_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)impl_->smem));
    mlp_chain_kernel<<<impl_->grid, kThreads, impl_->smem, stream>>>(P);
    NNP_CUDA_CHECK(cudaGetLastError());
    count_launch();
}

}  // namespace nnpops
