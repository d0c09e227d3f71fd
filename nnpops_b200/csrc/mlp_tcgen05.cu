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
#include <cstring>
#include <map>
#include <tuple>
#include "species_mlp.cuh"

namespace nnpops {

namespace {

constexpr int TBM = 128, TBN = 128, TBK = 64, kStages = 3, kAccStages = 2;
constexpr uint32_t kTileBytes = TBM * TBK * 2;        // 16 KB: 128 rows x 128 bytes
constexpr uint32_t kStageBytes = 4 * kTileBytes;      // A hi, A lo, B hi, B lo
constexpr uint32_t kSmemBytes = kStages * kStageBytes + 1024 /*alignment slack*/ + 256 /*barriers*/;
constexpr float kLoScale = 2048.0f, kLoInv = 1.0f / 2048.0f;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmemD, uint64_t descA, uint64_t descB, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmemD), "l"(descA), "l"(descB), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major operand tile [rows][64 halves] written by TMA with SWIZZLE_128B: 8-row groups are 1024 bytes apart (SBO), the
// leading-dimension offset is unused for swizzled K-major layouts (encoded 1), descriptor version 1 (Blackwell), layout type 2.
__device__ __forceinline__ uint64_t make_desc(uint32_t smemAddr) {
    uint64_t d = 0;
    d |= (uint64_t)((smemAddr >> 4) & 0x3fff);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, "
        "%27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float celu_f(float x) { return x > 0.0f ? x : kCeluAlpha * (__expf(x * (1.0f / kCeluAlpha)) - 1.0f); }
__device__ __forceinline__ float celu_grad_from_act_f(float a) { return a > 0.0f ? 1.0f : a * (1.0f / kCeluAlpha) + 1.0f; }

__device__ __forceinline__ void split_f(float v, __half& hi, __half& lo) {
    hi = __float2half_rn(v);
    lo = __float2half_rn((v - __half2float(hi)) * kLoScale);
}

struct TcArgs {
    int M, N, K, batch;
    int aBatchCols, bBatchRows;
    int mode;                      // 0: fp32 out * outScale; 1: bias + celu -> split; 2: * celu'(act) -> split
    __half* Chi; __half* Clo; float* C32; int ldc; int cBatchCols;
    const float* bias; int biasBatch;
    const __half* actHi; const __half* actLo; int ldact; int actBatchCols;
    float outScale;
    const float* w3; double* energyAcc; float seedScale;
};

__global__ void __launch_bounds__(256, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap mapAhi, const __grid_constant__ CUtensorMap mapAlo,
                    const __grid_constant__ CUtensorMap mapBhi, const __grid_constant__ CUtensorMap mapBlo, const TcArgs g) {
    extern __shared__ unsigned char smemRaw[];
    const uint32_t base = (smem_u32(smemRaw) + 1023u) & ~1023u;
    const uint32_t barBase = base + kStages * kStageBytes;
    // barriers: full[kStages], empty[kStages], accFull[kAccStages], accEmpty[kAccStages]; then the TMEM base address slot
    auto fullBar = [&](int s) { return barBase + 8u * s; };
    auto emptyBar = [&](int s) { return barBase + 8u * (kStages + s); };
    auto accFullBar = [&](int s) { return barBase + 8u * (2 * kStages + s); };
    auto accEmptyBar = [&](int s) { return barBase + 8u * (2 * kStages + kAccStages + s); };
    const uint32_t tmemSlot = barBase + 8u * (2 * kStages + 2 * kAccStages);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapAhi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapAlo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBhi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBlo) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; s++) { mbar_init(fullBar(s), 1); mbar_init(emptyBar(s), 1); }
        for (int s = 0; s < kAccStages; s++) { mbar_init(accFullBar(s), 1); mbar_init(accEmptyBar(s), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmemSlot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmemBase;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmemBase) : "r"(tmemSlot));

    const int tilesM = (g.M + TBM - 1) / TBM, tilesN = (g.N + TBN - 1) / TBN;
    const int numTiles = tilesM * tilesN * g.batch;
    const int kChunks = g.K / TBK;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < numTiles; t += gridDim.x) {
                const int nt = t % tilesN, mt = (t / tilesN) % tilesM, z = t / (tilesN * tilesM);
                const int m0 = mt * TBM, n0 = nt * TBN;
                for (int kc = 0; kc < kChunks; kc++) {
                    mbar_wait(emptyBar(stage), phase ^ 1u);
                    const uint32_t sb = base + stage * kStageBytes;
                    mbar_expect_tx(fullBar(stage), kStageBytes);
                    const int xa = z * g.aBatchCols + kc * TBK, xb = kc * TBK, yb = z * g.bBatchRows + n0;
                    tma_load_2d(sb, &mapAhi, fullBar(stage), xa, m0);
                    tma_load_2d(sb + kTileBytes, &mapAlo, fullBar(stage), xa, m0);
                    tma_load_2d(sb + 2 * kTileBytes, &mapBhi, fullBar(stage), xb, yb);
                    tma_load_2d(sb + 3 * kTileBytes, &mapBlo, fullBar(stage), xb, yb);
                    if (++stage == kStages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int stage = 0, acc = 0;
            uint32_t phase = 0, accPhase = 0;
            for (int t = blockIdx.x; t < numTiles; t += gridDim.x) {
                const int nt = t % tilesN;
                const int nTile = min(TBN, g.N - nt * TBN);   // 64 or 128 (N is a multiple of 64)
                const uint32_t idesc = (1u << 4) | ((uint32_t)(nTile >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
                mbar_wait(accEmptyBar(acc), accPhase ^ 1u);
                tc_fence_after();
                const uint32_t d1 = tmemBase + (uint32_t)(acc * 2 * TBN), d2 = d1 + TBN;
                for (int kc = 0; kc < kChunks; kc++) {
                    mbar_wait(fullBar(stage), phase);
                    tc_fence_after();
                    const uint32_t sb = base + stage * kStageBytes;
#pragma unroll
                    for (int k4 = 0; k4 < TBK / 16; k4++) {
                        const uint64_t aHi = make_desc(sb + k4 * 32), aLo = make_desc(sb + kTileBytes + k4 * 32);
                        const uint64_t bHi = make_desc(sb + 2 * kTileBytes + k4 * 32), bLo = make_desc(sb + 3 * kTileBytes + k4 * 32);
                        const uint32_t first = (kc | k4) ? 1u : 0u;
                        umma_f16(d1, aHi, bHi, idesc, first);
                        umma_f16(d2, aHi, bLo, idesc, first);
                        umma_f16(d2, aLo, bHi, idesc, 1u);
                    }
                    umma_commit(emptyBar(stage));          // frees the smem stage once these MMAs have read it
                    if (kc == kChunks - 1) umma_commit(accFullBar(acc));
                    if (++stage == kStages) { stage = 0; phase ^= 1u; }
                }
                if (++acc == kAccStages) { acc = 0; accPhase ^= 1u; }
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3;   // TMEM lane quarter owned by this warp
        int acc = 0;
        uint32_t accPhase = 0;
        for (int t = blockIdx.x; t < numTiles; t += gridDim.x) {
            const int nt = t % tilesN, mt = (t / tilesN) % tilesM, z = t / (tilesN * tilesM);
            const int m = mt * TBM + q * 32 + lane, n0 = nt * TBN;
            float esum = 0.0f;
            mbar_wait(accFullBar(acc), accPhase);
            tc_fence_after();
            const uint32_t tbase = tmemBase + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 2 * TBN);
#pragma unroll 1
            for (int c = 0; c < TBN / 32; c++) {
                const int n = n0 + c * 32;
                if (n >= g.N) break;                       // warp-uniform
                uint32_t r1[32], r2[32];
                tmem_ld32(tbase + c * 32, r1);
                tmem_ld32(tbase + TBN + c * 32, r2);
                tmem_ld_wait();
                if (m < g.M) {
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; j++) v[j] = fmaf(__uint_as_float(r2[j]), kLoInv, __uint_as_float(r1[j]));
                    if (g.mode == 0) {
                        float* dst = g.C32 + (size_t)m * g.ldc + (size_t)z * g.cBatchCols + n;
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            *reinterpret_cast<float4*>(dst + j) =
                                make_float4(v[j] * g.outScale, v[j + 1] * g.outScale, v[j + 2] * g.outScale, v[j + 3] * g.outScale);
                    } else {
                        if (g.mode == 1) {
                            const float* bp = g.bias + (size_t)z * g.biasBatch + n;
#pragma unroll
                            for (int j = 0; j < 32; j++) v[j] = celu_f(v[j] + __ldg(bp + j));
                        } else if (g.mode == 3) {
                            const float* bp = g.bias + (size_t)z * g.biasBatch + n;
                            const float* wp = g.w3 + (size_t)z * g.biasBatch + n;
#pragma unroll
                            for (int j = 0; j < 32; j++) {
                                const float a = celu_f(v[j] + __ldg(bp + j));
                                const float wv = __ldg(wp + j);
                                esum = fmaf(a, wv, esum);
                                v[j] = g.seedScale * wv * celu_grad_from_act_f(a);
                            }
                        } else {
                            const size_t ao = (size_t)m * g.ldact + (size_t)z * g.actBatchCols + n;
#pragma unroll
                            for (int j = 0; j < 32; j += 8) {
                                const uint4 ah = *reinterpret_cast<const uint4*>(g.actHi + ao + j);
                                const uint4 al = *reinterpret_cast<const uint4*>(g.actLo + ao + j);
                                const __half* hh = reinterpret_cast<const __half*>(&ah);
                                const __half* hl = reinterpret_cast<const __half*>(&al);
#pragma unroll
                                for (int i = 0; i < 8; i++) {
                                    const float a = fmaf(__half2float(hl[i]), kLoInv, __half2float(hh[i]));
                                    v[j + i] *= celu_grad_from_act_f(a);
                                }
                            }
                        }
                        const size_t co = (size_t)m * g.ldc + (size_t)z * g.cBatchCols + n;
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            uint4 ph, pl;
                            __half* hh = reinterpret_cast<__half*>(&ph);
                            __half* hl = reinterpret_cast<__half*>(&pl);
#pragma unroll
                            for (int i = 0; i < 8; i++) split_f(v[j + i], hh[i], hl[i]);
                            *reinterpret_cast<uint4*>(g.Chi + co + j) = ph;
                            *reinterpret_cast<uint4*>(g.Clo + co + j) = pl;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(accEmptyBar(acc));
            if (g.mode == 3) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) esum += __shfl_xor_sync(0xffffffffu, esum, o);
                if (lane == 0) atomicAdd(g.energyAcc, (double)esum);
            }
            if (++acc == kAccStages) { acc = 0; accPhase ^= 1u; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmemBase), "r"(512u) : "memory");
    }
}

// ---- host side ----------------------------------------------------------------------------------------------------

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
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

// 2-D fp16 tensor [rows][cols] with row pitch ld (elements); box = 64 columns x 128 rows, 128-byte swizzle, OOB reads give zeros
CUtensorMap make_map(const __half* ptr, long long rows, long long cols, long long ld) {
    NNP_REQUIRE(((uintptr_t)ptr & 15) == 0 && (ld * 2) % 16 == 0, "tcgen05 GEMM operands must be 16-byte aligned");
    CUtensorMap m;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    const cuuint32_t box[2] = {(cuuint32_t)TBK, (cuuint32_t)TBM};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(ptr), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    NNP_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed");
    return m;
}

int num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        NNP_CUDA_CHECK(cudaGetDevice(&dev));
        NNP_CUDA_CHECK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    }
    return n;
}

}  // namespace

void launch_gemm_tcgen05(const GemmArgsH& a, cudaStream_t stream) {
    if (a.M <= 0 || a.N <= 0) return;
    NNP_REQUIRE(a.K % TBK == 0 && a.K > 0, "tcgen05 GEMM: K must be a positive multiple of 64");
    NNP_REQUIRE(a.N % 64 == 0, "tcgen05 GEMM: N must be a multiple of 64");
    NNP_REQUIRE(a.ldc % 8 == 0 && a.cBatchCols % 8 == 0 && a.ldact % 8 == 0 && a.actBatchCols % 8 == 0,
                "tcgen05 GEMM: output leading dimensions must be multiples of 8");
    static bool attrSet = false;
    if (!attrSet) {
        NNP_CUDA_CHECK(cudaFuncSetAttribute(gemm_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
        attrSet = true;
    }
    const CUtensorMap mAhi = make_map(a.Ahi, a.M, a.aCols, a.lda), mAlo = make_map(a.Alo, a.M, a.aCols, a.lda);
    const CUtensorMap mBhi = make_map(a.Bhi, a.bRows, a.K, a.ldb), mBlo = make_map(a.Blo, a.bRows, a.K, a.ldb);
    TcArgs g;
    std::memset(&g, 0, sizeof(g));
    g.M = a.M; g.N = a.N; g.K = a.K; g.batch = a.batch; g.aBatchCols = a.aBatchCols; g.bBatchRows = a.bBatchRows; g.mode = a.epilogue;
    g.Chi = a.Chi; g.Clo = a.Clo; g.C32 = a.C32; g.ldc = a.ldc; g.cBatchCols = a.cBatchCols; g.bias = a.bias; g.biasBatch = a.biasBatch;
    g.actHi = a.actHi; g.actLo = a.actLo; g.ldact = a.ldact; g.actBatchCols = a.actBatchCols; g.outScale = a.outScale;
    g.w3 = a.w3; g.energyAcc = a.energyAcc; g.seedScale = a.seedScale;
    const int tiles = ((a.M + TBM - 1) / TBM) * ((a.N + TBN - 1) / TBN) * a.batch;
    const int grid = tiles < num_sms() ? tiles : num_sms();
    gemm_tcgen05_kernel<<<grid, 256, kSmemBytes, stream>>>(mAhi, mAlo, mBhi, mBlo, g);
    count_launch();
}

}  // namespace nnpops
