// fp32-accurate GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA) for the
// species-grouped MLP.
//
// Precision scheme ("fp16 x 3, scaled low part"): every fp32 operand x is stored as the pair
//     hi = fp16(x),   lo = fp16((x - hi) * 2^11)                      (x - hi is exact in fp32)
// and   sum_k a_k b_k  ~=  D1 + 2^-11 * D2,   D1 = sum ah*bh,   D2 = sum (ah*bl + al*bh)
// with D1 and D2 accumulated in fp32 in two TMEM accumulators.  fp16 x fp16 products are exact in fp32, the dropped al*bl term
// is 2^-22 relative, so the result carries ~22 mantissa bits -- the accuracy of an fp32 FMA chain -- at 3 fp16 MMAs per
// product (twice the rate of a 3xTF32 scheme, and the same bytes per element as fp32).
//
// Kernel: persistent CTAs (one per SM), tile 128 x 128, K chunks of 64 halves (= one 128-byte swizzle row).
//   warp 0      TMA producer   : 4 boxes per chunk (A hi/lo, B hi/lo) into a 3-stage ring, mbarrier complete_tx
//   warp 1      MMA issuer     : one elected lane, 12 tcgen05.mma (M128 N128|64 K16) per chunk, tcgen05.commit frees the stage
//   warp 2      TMEM allocator : 512 columns = 2 accumulator stages x {D1, D2} x 128 columns
//   warps 4..7  epilogue       : tcgen05.ld -> D1 + 2^-11 D2 -> bias/CELU or celu' mask -> re-split to hi/lo fp16 (or plain fp32)
// so the epilogue of tile i overlaps the main loop of tile i+1.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>
#include "species_mlp.cuh"
#include "tcgen05_util.cuh"
#include "workspace_cache.cuh"

namespace nnpops {

namespace {

// Development switches (scripts/gemm_dbg.py): parts of the kernel can be turned off through NNPOPS_GEMM_DBG when the library is built with
// -DNNPOPS_GEMM_DEBUG (NNPOPS_BUILD_DEFINES=-DNNPOPS_GEMM_DEBUG python -m nnpops_b200.build --force); compiled out otherwise.
#ifdef NNPOPS_GEMM_DEBUG
#define GEMM_DBG(bit) ((g.dbg & (bit)) != 0)
#else
#define GEMM_DBG(bit) false
#endif

constexpr int TBM = 128, TBN = 128, TBK = 64, kAccStages = 2;
// Operand ring depth: 3 stages; the celu'-mask epilogue (mode 2) gives one stage up for per-warp activation prefetch buffers.
__host__ __device__ constexpr int stages_of(int mode) { return mode == 2 ? 2 : 3; }
constexpr uint32_t kTileBytes = TBM * TBK * 2;        // 16 KB: 128 rows x 128 bytes
constexpr uint32_t kStageBytes = 4 * kTileBytes;      // streaming mode: A hi, A lo, B hi, B lo
constexpr uint32_t kStageWarpBytes = 32 * 64;         // per epilogue warp: 32 rows x 32 halves (XOR-swizzled), transposes to coalesced rows
// mode 2 adds two more such blocks per warp: the activation hi / lo block of the warp's NEXT tile, filled by cp.async
__host__ __device__ constexpr uint32_t warp_bytes_of(int mode) { return mode == 2 ? 3 * kStageWarpBytes : kStageWarpBytes; }
constexpr int kEpiWarps = 16;          // 4 TMEM lane quarters x 4 column blocks of 32: one 32 x 32 block per warp per tile
constexpr int kFirstEpiWarp = 2;     // warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer
constexpr int kThreads = (kFirstEpiWarp + kEpiWarps) * 32;   // 576 threads -> up to 112 registers each
__host__ __device__ constexpr uint32_t smem_bytes_of(int mode) {
    return stages_of(mode) * kStageBytes + kEpiWarps * warp_bytes_of(mode) + 1024 /*alignment slack*/ + 256 /*barriers*/;
}
using namespace tc;   // inline-PTX wrappers (tcgen05_util.cuh)

struct TcArgs {
    int M, N, K, batch;
    int aBatchCols, bBatchRows;
    int mode;                      // 0: fp32 out * outScale; 1: bias + celu -> split; 2: * celu'(act) -> split
    __half* Chi; __half* Clo; float* C32; int ldc; int cBatchCols;
    const float* bias; int biasBatch;
    const __half* actHi; const __half* actLo; int ldact; int actBatchCols;
    float outScale;
    const float* w3; double* energyAcc; float seedScale;
    int wide;                      // 1: issue Ahi.[Bhi;Blo] as one N = 256 MMA
    int dbg;                       // development switches (NNPOPS_GEMM_DBG bit mask), 0 in production
};

// ---- epilogue of one 128 x 64 slice: TMEM -> registers -> fused element-wise op -> shared-memory transpose -> coalesced global ----
// A thread owns one row of the accumulator (TMEM lane), so direct stores would touch 32 different lines per instruction.  Every
// 32 x 32 block is therefore transposed through a per-warp staging buffer: the thread writes its 64-byte row segment, then the
// warp stores two full row segments per instruction (lanes 0-15 -> row r, lanes 16-31 -> row r + 1, 4 bytes per lane).
// Staging layout: 32 rows x 64 bytes, no padding; 16-byte unit u of row r lives at unit (u ^ ((r >> 1) & 3)).  With that XOR both
// access patterns are bank-conflict free: a thread writing/reading its own row (lanes = rows) and a quarter-warp moving two
// complete rows (lanes = 16-byte units of consecutive rows).  Global accesses are 16 bytes per lane, 8 full rows per instruction.
__device__ __forceinline__ uint32_t stage_off(int r, int u) { return (uint32_t)(r * 64 + ((u ^ ((r >> 1) & 3)) << 4)); }

__device__ __forceinline__ void staged_store_half(unsigned char* stg, const uint32_t (&pk)[16], __half* gbase, size_t ld, int rowsValid, int lane) {
#pragma unroll
    for (int i = 0; i < 4; i++)
        *reinterpret_cast<uint4*>(stg + stage_off(lane, i)) = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
    __syncwarp();
    const int r0 = lane >> 2, u = lane & 3;
    __half* dst = gbase + (size_t)r0 * ld + u * 8;
    const size_t step = 8 * ld;
#pragma unroll
    for (int it = 0; it < 4; it++) {
        const int r = it * 8 + r0;
        const uint4 v = *reinterpret_cast<const uint4*>(stg + stage_off(r, u));
        if (r < rowsValid) *reinterpret_cast<uint4*>(dst) = v;
        dst += step;
    }
    __syncwarp();
}

// the reverse: coalesced load of a 32 x 32 block of halves, each thread ends up with its own row (16 packed registers)
__device__ __forceinline__ void staged_load_half(unsigned char* stg, uint32_t (&pk)[16], const __half* gbase, size_t ld, int rowsValid, int lane) {
    const int r0 = lane >> 2, u = lane & 3;
    const __half* src = gbase + (size_t)r0 * ld + u * 8;
    const size_t step = 8 * ld;
    uint4 tmp[4];
#pragma unroll
    for (int it = 0; it < 4; it++) {      // all loads in flight before the first shared-memory store
        tmp[it] = (it * 8 + r0 < rowsValid) ? __ldg(reinterpret_cast<const uint4*>(src)) : make_uint4(0u, 0u, 0u, 0u);
        src += step;
    }
#pragma unroll
    for (int it = 0; it < 4; it++) *reinterpret_cast<uint4*>(stg + stage_off(it * 8 + r0, u)) = tmp[it];
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint4 t = *reinterpret_cast<const uint4*>(stg + stage_off(lane, i));
        pk[4 * i] = t.x; pk[4 * i + 1] = t.y; pk[4 * i + 2] = t.z; pk[4 * i + 3] = t.w;
    }
    __syncwarp();
}

// warp (q = TMEM lane quarter, hsel = column block) handles rows q*32 + lane and columns [hsel*32, hsel*32 + 32) of the tile.  Four
// epilogue warps per scheduler: the latencies of one warp's chain (bias / activation loads, TMEM load, staging) hide behind the others
template <int MODE, typename WAIT, typename ACTDONE>
__device__ __forceinline__ void epilogue_tile(const TcArgs& g, uint32_t tmemBase, int acc, int q, int hsel, int lane, int mt, int nt, int z,
                                              unsigned char* stg, float& esum, const unsigned char* actBuf, float biasLane, float w3Lane,
                                              WAIT&& waitAccumulator, ACTDONE&& actConsumed) {
    const int mw = mt * TBM + q * 32;              // first row of this warp
    const int m = mw + lane, n0 = nt * TBN;
    const int rowsValid = min(32, g.M - mw);       // warp-uniform, may be <= 0
    const uint32_t tbase = tmemBase + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 2 * TBN);
    bool waited = false;
    const int c = hsel;
    const int n = n0 + c * 32;
    if (n < g.N) {                                  // warp-uniform
        if (MODE == 1 || MODE == 3) {
            // bias (and last-layer weights) of the warp's 32 columns were fetched one tile ahead, one column per lane; they pass
            // through the (idle) staging block so that every lane can read them as broadcast float4
            reinterpret_cast<float*>(stg)[lane] = biasLane;
            if (MODE == 3) { reinterpret_cast<float*>(stg)[32 + lane] = w3Lane; reinterpret_cast<float*>(stg)[64 + lane] = w3Lane * g.seedScale; }
            __syncwarp();
        }
        waitAccumulator(); waited = true;
        uint32_t ph[16], pl[16];
        const bool rowOk = m < g.M;
#pragma unroll
        for (int h = 0; h < 2; h++) {               // two 16-column halves keep the register footprint small
            float4 bv[4], wv[4], sv[4];
            if (MODE == 1 || MODE == 3) {
                const float4* bp = reinterpret_cast<const float4*>(stg) + 4 * h;
#pragma unroll
                for (int j = 0; j < 4; j++) bv[j] = bp[j];
                if (MODE == 3) {
#pragma unroll
                    for (int j = 0; j < 4; j++) { wv[j] = bp[8 + j]; sv[j] = bp[16 + j]; }
                }
            }
            uint32_t r1[16], r2[16];
            tmem_ld16(tbase + c * 32 + 16 * h, r1);
            tmem_ld16(tbase + TBN + c * 32 + 16 * h, r2);
            tmem_ld_wait();
            if (rowsValid <= 0 || GEMM_DBG(4)) continue;
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; j++) v[j] = fmaf(__uint_as_float(r2[j]), kLoInv, __uint_as_float(r1[j]));
            if (MODE == 0) {                      // fp32 result (dE/dAEV): the main loop is long, direct stores stay hidden
                if (rowOk) {
                    float* dst = g.C32 + (size_t)m * g.ldc + (size_t)z * g.cBatchCols + n + 16 * h;
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        *reinterpret_cast<float4*>(dst + j) =
                            make_float4(v[j] * g.outScale, v[j + 1] * g.outScale, v[j + 2] * g.outScale, v[j + 3] * g.outScale);
                }
                continue;
            }
            if (MODE == 1) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    v[4 * j] = celu_f(v[4 * j] + bv[j].x); v[4 * j + 1] = celu_f(v[4 * j + 1] + bv[j].y);
                    v[4 * j + 2] = celu_f(v[4 * j + 2] + bv[j].z); v[4 * j + 3] = celu_f(v[4 * j + 3] + bv[j].w);
                }
            } else if (MODE == 3) {
                // last hidden layer: a = celu(z + b) feeds the energy (a . w3) and the backward seed w3 / M * celu'(z); celu' is the
                // exponential celu already computed (x <= 0) or 1, and w3 arrives pre-multiplied by the seed scale in a second copy
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float b4[4] = {bv[j].x, bv[j].y, bv[j].z, bv[j].w};
                    const float w4[4] = {wv[j].x, wv[j].y, wv[j].z, wv[j].w};
                    const float s4[4] = {sv[j].x, sv[j].y, sv[j].z, sv[j].w};
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const float x = v[4 * j + i] + b4[i];
                        float e;
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * (1.4426950408889634f / kCeluAlpha)));
                        const bool pos = x > 0.0f;
                        const float a = pos ? x : fmaf(kCeluAlpha, e, -kCeluAlpha);
                        esum = fmaf(a, w4[i], esum);
                        v[4 * j + i] = s4[i] * (pos ? 1.0f : e);
                    }
                }
                if (!rowOk) esum = 0.0f;          // rows beyond M carry zero activations but a non-zero bias: keep them out of the energy
            } else {
                // the activation block of this tile sits in the warp's cp.async buffers: read this half's 16 columns of the lane's row
                // (hi and lo), and hand the buffers back for the next tile's prefetch once the second half has been read
                uint32_t actH[8], actL[8];
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    const uint4 th = *reinterpret_cast<const uint4*>(actBuf + stage_off(lane, 2 * h + i));
                    const uint4 tl = *reinterpret_cast<const uint4*>(actBuf + kStageWarpBytes + stage_off(lane, 2 * h + i));
                    actH[4 * i] = th.x; actH[4 * i + 1] = th.y; actH[4 * i + 2] = th.z; actH[4 * i + 3] = th.w;
                    actL[4 * i] = tl.x; actL[4 * i + 1] = tl.y; actL[4 * i + 2] = tl.z; actL[4 * i + 3] = tl.w;
                }
                if (h == 1) actConsumed();
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const float2 fh = unpack_h2(actH[i]), fl = unpack_h2(actL[i]);
                    v[2 * i] *= celu_grad_from_act_f(fmaf(fl.x, kLoInv, fh.x));
                    v[2 * i + 1] *= celu_grad_from_act_f(fmaf(fl.y, kLoInv, fh.y));
                }
            }
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const __half2 h2 = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
                const float2 f2 = __half22float2(h2);
                ph[8 * h + i] = *reinterpret_cast<const uint32_t*>(&h2);
                pl[8 * h + i] = pack_h2((v[2 * i] - f2.x) * kLoScale, (v[2 * i + 1] - f2.y) * kLoScale);
            }
        }
        if (MODE == 1 || MODE == 3) __syncwarp();   // all lanes are done reading the bias block before it becomes the store staging
        if (MODE != 0 && rowsValid > 0 && !GEMM_DBG(5)) {
            const size_t co = (size_t)mw * g.ldc + (size_t)z * g.cBatchCols + n;
            staged_store_half(stg, ph, g.Chi + co, g.ldc, rowsValid, lane);
            staged_store_half(stg, pl, g.Clo + co, g.ldc, rowsValid, lane);
        }
    }
    if (!waited) waitAccumulator();                // tail tile with no columns for this warp: still consume the barrier phase
}

// One 64-wide K chunk: D1 += Ahi.Bhi, D2 += Ahi.Blo + Alo.Bhi.  The B hi and lo tiles are adjacent in shared memory (256 rows of
// 128 bytes) and D1 | D2 are adjacent in TMEM, so for a full-width tile Ahi . [Bhi; Blo] is ONE N = 256 instruction producing D1
// and the first half of D2 together: 8 instead of 12 tcgen05.mma per chunk.  descBits = all descriptor fields except the start
// address (constant for the kernel); the start-address field counts 16-byte units, so stepping K by 16 halves adds 2.
// KSTEPS < 4: the last chunk of a K that is not a multiple of 64 (the columns beyond K in the staged tiles are never multiplied).
template <bool FULL, int KSTEPS = TBK / 16>
__device__ __forceinline__ void issue_chunk(uint32_t d1, uint64_t descBits, uint32_t stageAddr, uint32_t idescTile, bool firstChunk) {
    const uint64_t aHi0 = descBits + (stageAddr >> 4);
    constexpr uint32_t kTile16 = kTileBytes >> 4;
    constexpr uint32_t idescWide = (1u << 4) | ((uint32_t)(2 * TBN >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
    const uint32_t d2 = d1 + TBN;
#pragma unroll
    for (int k4 = 0; k4 < KSTEPS; k4++) {
        const uint64_t aHi = aHi0 + 2 * k4, aLo = aHi + kTile16, bHi = aHi + 2 * kTile16, bLo = aHi + 3 * kTile16;
        const uint32_t accum = (firstChunk && k4 == 0) ? 0u : 1u;
        if (FULL) {
            umma_f16(d1, aHi, bHi, idescWide, accum);          // N = 256: [D1 | D2] (+)= Ahi . [Bhi; Blo]
        } else {
            umma_f16(d1, aHi, bHi, idescTile, accum);
            umma_f16(d2, aHi, bLo, idescTile, accum);
        }
        umma_f16(d2, aLo, bHi, idescTile, 1u);
    }
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap mapAhi, const __grid_constant__ CUtensorMap mapAlo,
                    const __grid_constant__ CUtensorMap mapBhi, const __grid_constant__ CUtensorMap mapBlo, const TcArgs g) {
    extern __shared__ unsigned char smemRaw[];
    const uint32_t rawAddr = smem_u32(smemRaw);
    const uint32_t base = (rawAddr + 1023u) & ~1023u;
    unsigned char* const basePtr = smemRaw + (base - rawAddr);
    constexpr int kStages = stages_of(MODE);
    constexpr uint32_t kDataBytes = kStages * kStageBytes;
    const uint32_t barBase = base + kDataBytes + kEpiWarps * warp_bytes_of(MODE);
    // barriers: full[kStages], empty[kStages], accFull[kAccStages], accEmpty[kAccStages]; then the TMEM address slot
    auto fullBar = [&](int s) { return barBase + 8u * s; };
    auto emptyBar = [&](int s) { return barBase + 8u * (kStages + s); };
    auto accFullBar = [&](int s) { return barBase + 8u * (2 * kStages + s); };
    auto accEmptyBar = [&](int s) { return barBase + 8u * (2 * kStages + kAccStages + s); };
    const uint32_t tmemSlot = barBase + 8u * (2 * kStages + 2 * kAccStages);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // Programmatic dependent launch: the next GEMM of the chain may start its CTAs (barrier init, TMEM allocation, descriptor
    // prefetch) on SMs this grid has already left; it blocks in griddepcontrol.wait below until this grid has completed.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapAhi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapAlo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBhi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBlo) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; s++) { mbar_init(fullBar(s), 1); mbar_init(emptyBar(s), 1); }
        for (int s = 0; s < kAccStages; s++) { mbar_init(accFullBar(s), 1); mbar_init(accEmptyBar(s), kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmemSlot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");   // everything this kernel reads from global memory was written before this point
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmemBase;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmemBase) : "r"(tmemSlot));

    const int tilesM = (g.M + TBM - 1) / TBM, tilesN = (g.N + TBN - 1) / TBN;
    const int kChunks = (g.K + TBK - 1) / TBK;
    const int lastSteps = (g.K - (kChunks - 1) * TBK) / 16;   // k16 steps of the last chunk (4 unless K % 64 != 0)
    const int numTiles = tilesM * tilesN * g.batch;
    // Tile order: batch member fastest, then n-tile, then m-tile.  Concurrent CTAs then work on the same rows of A for all
    // ensemble members at once (each row of the activation matrix is read as one contiguous run from HBM) and on few m-tiles.
    auto decode = [&](int t, int& mt, int& nt, int& z) {
        z = t % g.batch;
        nt = (t / g.batch) % tilesN;
        mt = t / (g.batch * tilesN);
    };

    // Producer and MMA warps stay converged: every lane runs the loops and the barrier waits (so all control values are
    // warp-uniform), and the one lane picked by elect.sync issues the TMA / tcgen05 instructions.
    if (warp == 0) {
        const bool leader = elect_one();
        int stage = 0;
        uint32_t phase = 0;
        for (int t = blockIdx.x; t < numTiles; t += gridDim.x) {
            int mt, nt, z;
            decode(t, mt, nt, z);
            const int m0 = mt * TBM, yb = z * g.bBatchRows + nt * TBN;
            for (int kc = 0; kc < kChunks; kc++) {
                mbar_wait(emptyBar(stage), phase ^ 1u);
                if (leader) {
                    const uint32_t sb = base + stage * kStageBytes;
                    mbar_expect_tx(fullBar(stage), GEMM_DBG(8) ? kStageBytes - 2 * kTileBytes : kStageBytes);
                    const int xa = z * g.aBatchCols + kc * TBK;
                    if (!GEMM_DBG(8)) {
                    tma_load_2d(sb, &mapAhi, fullBar(stage), xa, m0);
                    tma_load_2d(sb + kTileBytes, &mapAlo, fullBar(stage), xa, m0);
                    }
                    tma_load_2d(sb + 2 * kTileBytes, &mapBhi, fullBar(stage), kc * TBK, yb);
                    tma_load_2d(sb + 3 * kTileBytes, &mapBlo, fullBar(stage), kc * TBK, yb);
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        const bool leader = elect_one();
        const uint64_t descBits = make_desc(0);
        int stage = 0, acc = 0;
        uint32_t phase = 0, accPhase = 0;
        for (int t = blockIdx.x; t < numTiles; t += gridDim.x) {
            int mt, nt, z;
            decode(t, mt, nt, z);
            const int nTile = min(TBN, g.N - nt * TBN);   // a multiple of 32
            const uint32_t idescTile = (1u << 4) | ((uint32_t)(nTile >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
            mbar_wait(accEmptyBar(acc), accPhase ^ 1u);
            tc_fence_after();
            const uint32_t d1 = tmemBase + (uint32_t)(acc * 2 * TBN);
            for (int kc = 0; kc < kChunks; kc++) {
                mbar_wait(fullBar(stage), phase);
                tc_fence_after();
                if (leader) {
                    const uint32_t sb = base + stage * kStageBytes;
                    if (GEMM_DBG(16)) {}
                    else if (kc == kChunks - 1 && lastSteps != TBK / 16) {
                        if (nTile == TBN) {
                            if (lastSteps == 2) issue_chunk<true, 2>(d1, descBits, sb, idescTile, kc == 0);
                            else if (lastSteps == 1) issue_chunk<true, 1>(d1, descBits, sb, idescTile, kc == 0);
                            else issue_chunk<true, 3>(d1, descBits, sb, idescTile, kc == 0);
                        } else {
                            if (lastSteps == 2) issue_chunk<false, 2>(d1, descBits, sb, idescTile, kc == 0);
                            else if (lastSteps == 1) issue_chunk<false, 1>(d1, descBits, sb, idescTile, kc == 0);
                            else issue_chunk<false, 3>(d1, descBits, sb, idescTile, kc == 0);
                        }
                    }
                    else if (nTile == TBN) issue_chunk<true>(d1, descBits, sb, idescTile, kc == 0);
                    else issue_chunk<false>(d1, descBits, sb, idescTile, kc == 0);
                    umma_commit(emptyBar(stage));          // frees the smem stage once these MMAs have read it
                    if (kc == kChunks - 1) umma_commit(accFullBar(acc));
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
            if (++acc == kAccStages) { acc = 0; accPhase ^= 1u; }
        }
    } else if (warp >= kFirstEpiWarp) {
        const int q = warp & 3;                          // TMEM lane quarter this warp may access (warp id mod 4)
        const int hsel = (warp - kFirstEpiWarp) >> 2;    // which 32-column block of the tile (distinct among the 4 warps of a quarter)
        unsigned char* const stg = basePtr + kDataBytes + (warp - kFirstEpiWarp) * warp_bytes_of(MODE);
        int acc = 0;
        uint32_t accPhase = 0;
        // mode 2: the activation block (hi, lo) this warp needs for celu' is fetched with cp.async one tile ahead into the warp's
        // own shared-memory blocks (same swizzled layout as the store staging), so its HBM latency is off the critical path
        unsigned char* const actBuf = stg + kStageWarpBytes;
        auto prefetch_act = [&](int t) {
            int mt, nt, z;
            decode(t, mt, nt, z);
            const int mw = mt * TBM + q * 32, n = nt * TBN + hsel * 32;
            const int rowsValid = (n < g.N) ? min(32, g.M - mw) : 0;
            const int r0 = lane >> 2, u = lane & 3;
            const size_t ao = (size_t)max(min(mw, g.M - 1), 0) * g.ldact + (size_t)z * g.actBatchCols + min(n, g.N - 32) + u * 8;
#pragma unroll
            for (int it = 0; it < 4; it++) {
                const int r = it * 8 + r0;
                const uint32_t bytes = r < rowsValid ? 16u : 0u;
                const size_t o = ao + (bytes ? (size_t)r * g.ldact : 0);
                const uint32_t dst = smem_u32(actBuf + stage_off(r, u));
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(g.actHi + o), "r"(bytes) : "memory");
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + kStageWarpBytes), "l"(g.actLo + o), "r"(bytes) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        if (MODE == 2 && !GEMM_DBG(2) && (int)blockIdx.x < numTiles) prefetch_act(blockIdx.x);
        float biasNext = 0.0f, w3Next = 0.0f;
        auto prefetch_cols = [&](int t) {
            int mt, nt, z;
            decode(t, mt, nt, z);
            const int col = nt * TBN + hsel * 32 + lane;
            if (col < g.N) {
                biasNext = __ldg(g.bias + (size_t)z * g.biasBatch + col);
                if (MODE == 3) w3Next = __ldg(g.w3 + (size_t)z * g.biasBatch + col);
            }
        };
        if ((MODE == 1 || MODE == 3) && (int)blockIdx.x < numTiles) prefetch_cols(blockIdx.x);
        for (int t = blockIdx.x; t < numTiles; t += gridDim.x) {
            int mt, nt, z;
            decode(t, mt, nt, z);
            float esum = 0.0f;
            const float biasCur = biasNext, w3Cur = w3Next;
            if ((MODE == 1 || MODE == 3) && t + (int)gridDim.x < numTiles) prefetch_cols(t + gridDim.x);
            if (MODE == 2 && !GEMM_DBG(2)) {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                __syncwarp();
            }
            bool actHanded = false;
            auto actConsumed = [&]() {
                __syncwarp();
                if (MODE == 2 && !GEMM_DBG(2) && t + (int)gridDim.x < numTiles) prefetch_act(t + gridDim.x);
                actHanded = true;
            };
            epilogue_tile<MODE>(g, tmemBase, acc, q, hsel, lane, mt, nt, z, stg, esum, actBuf, biasCur, w3Cur, [&]() {
                mbar_wait(accFullBar(acc), accPhase);
                tc_fence_after();
            }, actConsumed);
            if (MODE == 2 && !actHanded) actConsumed();   // tiles this warp skipped (no rows / columns): keep the prefetch chain going
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(accEmptyBar(acc));
            if (MODE == 3) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) esum += __shfl_xor_sync(0xffffffffu, esum, o);
                if (lane == 0 && esum != 0.0f) atomicAdd(g.energyAcc, (double)esum);
            }
            if (++acc == kAccStages) { acc = 0; accPhase ^= 1u; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmemBase), "r"(512u) : "memory");
    }
}

// ---- host side ----------------------------------------------------------------------------------------------------

}  // namespace

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        NNP_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        NNP_REQUIRE(p != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available in this driver");
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D fp16 tensor [rows][cols] with row pitch ld (elements); box = 64 columns x 128 rows, 128-byte swizzle, OOB reads give zeros.
// The encodes are cached: a model launches the same dozen GEMMs on the same buffers every evaluation, and 48 driver calls per
// evaluation are a visible share of the host time of small systems and of the sharded-box mode.
CUtensorMap tc_make_map(const __half* ptr, long long rows, long long cols, long long ld) {
    NNP_REQUIRE(((uintptr_t)ptr & 15) == 0 && (ld * 2) % 16 == 0, "tcgen05 GEMM operands must be 16-byte aligned");
    static std::mutex mu;
    static std::map<std::tuple<const void*, long long, long long, long long>, CUtensorMap> cache;
    const auto key = std::make_tuple((const void*)ptr, rows, cols, ld);
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    CUtensorMap m;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    const cuuint32_t box[2] = {(cuuint32_t)TBK, (cuuint32_t)TBM};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(ptr), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    NNP_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed");
    if (cache.size() > 4096) cache.clear();   // long-lived processes creating many models: bound the table
    cache.emplace(key, m);
    return m;
}

namespace {

bool g_forceStreaming = std::getenv("NNPOPS_GEMM_STREAMING") != nullptr;   // A/B switch for measurements

int num_sms() { return current_sm_count(); }

}  // namespace

void gemm_tcgen05_set_streaming(bool on) { g_forceStreaming = on; }

void launch_gemm_tcgen05(const GemmArgsH& a, cudaStream_t stream) {
    if (a.M <= 0 || a.N <= 0) return;
    NNP_REQUIRE(a.K % 16 == 0 && a.K > 0, "tcgen05 GEMM: K must be a positive multiple of 16");
    NNP_REQUIRE(a.N % 32 == 0, "tcgen05 GEMM: N must be a multiple of 32");
    NNP_REQUIRE(a.ldc % 8 == 0 && a.cBatchCols % 8 == 0 && a.ldact % 8 == 0 && a.actBatchCols % 8 == 0,
                "tcgen05 GEMM: output leading dimensions must be multiples of 8");
    // the attribute belongs to the (function, device) pair: a process may build models on several GPUs
    static bool attrSetOn[64] = {false};
    int dev = 0;
    NNP_CUDA_CHECK(cudaGetDevice(&dev));
    bool& attrSet = attrSetOn[dev & 63];
    if (!attrSet) {
        NNP_CUDA_CHECK(cudaFuncSetAttribute(gemm_tcgen05_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes_of(0)));
        NNP_CUDA_CHECK(cudaFuncSetAttribute(gemm_tcgen05_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes_of(1)));
        NNP_CUDA_CHECK(cudaFuncSetAttribute(gemm_tcgen05_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes_of(2)));
        NNP_CUDA_CHECK(cudaFuncSetAttribute(gemm_tcgen05_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes_of(3)));
        attrSet = true;
    }
    const CUtensorMap mAhi = tc_make_map(a.Ahi, a.M, a.aCols, a.lda), mAlo = tc_make_map(a.Alo, a.M, a.aCols, a.lda);
    const CUtensorMap mBhi = tc_make_map(a.Bhi, a.bRows, a.K, a.ldb), mBlo = tc_make_map(a.Blo, a.bRows, a.K, a.ldb);
    TcArgs g;
    std::memset(&g, 0, sizeof(g));
    g.M = a.M; g.N = a.N; g.K = a.K; g.batch = a.batch; g.aBatchCols = a.aBatchCols; g.bBatchRows = a.bBatchRows; g.mode = a.epilogue;
    g.Chi = a.Chi; g.Clo = a.Clo; g.C32 = a.C32; g.ldc = a.ldc; g.cBatchCols = a.cBatchCols; g.bias = a.bias; g.biasBatch = a.biasBatch;
    g.actHi = a.actHi; g.actLo = a.actLo; g.ldact = a.ldact; g.actBatchCols = a.actBatchCols; g.outScale = a.outScale;
    g.w3 = a.w3; g.energyAcc = a.energyAcc; g.seedScale = a.seedScale;
    g.wide = 1;
    { const char* e = std::getenv("NNPOPS_GEMM_DBG"); g.dbg = e ? std::atoi(e) : 0; }
    const int tiles = ((a.M + TBM - 1) / TBM) * ((a.N + TBN - 1) / TBN) * a.batch;
    const int grid = tiles < num_sms() ? tiles : num_sms();
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    static const bool noPdl = std::getenv("NNPOPS_NO_PDL") != nullptr;
    cfg.attrs = attr; cfg.numAttrs = noPdl ? 0 : 1;
    switch (g.mode) {
        case 0: cfg.dynamicSmemBytes = smem_bytes_of(0); NNP_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<0>, mAhi, mAlo, mBhi, mBlo, g)); break;
        case 1: cfg.dynamicSmemBytes = smem_bytes_of(1); NNP_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<1>, mAhi, mAlo, mBhi, mBlo, g)); break;
        case 2: cfg.dynamicSmemBytes = smem_bytes_of(2); NNP_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<2>, mAhi, mAlo, mBhi, mBlo, g)); break;
        case 3: cfg.dynamicSmemBytes = smem_bytes_of(3); NNP_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<3>, mAhi, mAlo, mBhi, mBlo, g)); break;
        default: NNP_REQUIRE(false, "tcgen05 GEMM: unknown epilogue mode");
    }
    count_launch();
}

// Micro-benchmark of one GEMM shape on synthetic operands (development aid behind nnpops_debug_gemm_bench): mean milliseconds.
double gemm_tcgen05_bench(int M, int N, int K, int batch, int mode, int iters) {
    const int lda = batch > 1 ? batch * K : K, ldc = batch * N;
    __half *Ahi, *Alo, *Bhi, *Blo, *Chi, *Clo, *actHi, *actLo;
    float *C32, *bias, *w3;
    double* eacc;
    const size_t na = (size_t)M * lda, nb = (size_t)batch * N * K, nc = (size_t)M * ldc;
    NNP_CUDA_CHECK(cudaMalloc(&Ahi, 2 * na)); NNP_CUDA_CHECK(cudaMalloc(&Alo, 2 * na));
    NNP_CUDA_CHECK(cudaMalloc(&Bhi, 2 * nb)); NNP_CUDA_CHECK(cudaMalloc(&Blo, 2 * nb));
    NNP_CUDA_CHECK(cudaMalloc(&Chi, 2 * nc)); NNP_CUDA_CHECK(cudaMalloc(&Clo, 2 * nc));
    NNP_CUDA_CHECK(cudaMalloc(&actHi, 2 * nc)); NNP_CUDA_CHECK(cudaMalloc(&actLo, 2 * nc));
    NNP_CUDA_CHECK(cudaMalloc(&C32, 4 * nc)); NNP_CUDA_CHECK(cudaMalloc(&bias, 4 * (size_t)ldc)); NNP_CUDA_CHECK(cudaMalloc(&w3, 4 * (size_t)ldc));
    NNP_CUDA_CHECK(cudaMalloc(&eacc, 8));
    cudaMemset(Ahi, 0x2c, 2 * na); cudaMemset(Alo, 0x2c, 2 * na); cudaMemset(Bhi, 0x2c, 2 * nb); cudaMemset(Blo, 0x2c, 2 * nb);
    cudaMemset(actHi, 0x2c, 2 * nc); cudaMemset(actLo, 0x2c, 2 * nc); cudaMemset(bias, 0, 4 * (size_t)ldc); cudaMemset(w3, 0, 4 * (size_t)ldc);
    cudaMemset(eacc, 0, 8);
    GemmArgsH g;
    std::memset(&g, 0, sizeof(g));
    g.Ahi = Ahi; g.Alo = Alo; g.lda = lda; g.aCols = lda; g.aBatchCols = batch > 1 ? K : 0;
    g.Bhi = Bhi; g.Blo = Blo; g.ldb = K; g.bRows = batch * N; g.bBatchRows = batch > 1 ? N : 0;
    g.Chi = Chi; g.Clo = Clo; g.C32 = C32; g.ldc = ldc; g.cBatchCols = batch > 1 ? N : 0; g.bias = bias; g.biasBatch = batch > 1 ? N : 0;
    g.actHi = actHi; g.actLo = actLo; g.ldact = ldc; g.actBatchCols = batch > 1 ? N : 0; g.outScale = 1.0f; g.w3 = w3; g.energyAcc = eacc;
    g.seedScale = 1.0f; g.M = M; g.N = N; g.K = K; g.batch = batch; g.epilogue = mode;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 2; i++) launch_gemm_tcgen05(g, nullptr);
    cudaEventRecord(e0);
    for (int i = 0; i < iters; i++) launch_gemm_tcgen05(g, nullptr);
    cudaEventRecord(e1);
    NNP_CUDA_CHECK(cudaEventSynchronize(e1));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    for (void* p : {(void*)Ahi, (void*)Alo, (void*)Bhi, (void*)Blo, (void*)Chi, (void*)Clo, (void*)actHi, (void*)actLo, (void*)C32, (void*)bias,
                    (void*)w3, (void*)eacc})
        cudaFree(p);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return ms / iters;
}

}  // namespace nnpops
