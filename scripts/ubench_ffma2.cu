// Micro-benchmark: issue cost of packed fp32 FMA (FFMA2, sm_100) against scalar FFMA, alone and mixed with MUFU.EX2.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench_ffma2 scripts/ubench_ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                       rc = *reinterpret_cast<unsigned long long*>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MODE>
__global__ void k(float* out, int iters, float seed) {
    float2 a[8], b = make_float2(seed, seed * 0.5f);
    for (int i = 0; i < 8; i++) a[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.25f);
    float m = seed;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) { a[i].x = fmaf(a[i].x, b.x, b.y); a[i].y = fmaf(a[i].y, b.x, b.y); }          // 16 FFMA
            if (MODE == 1) { a[i] = ffma2(a[i], b, b); }                                                      // 8 FFMA2 (same flops)
            if (MODE == 2) { a[i].x = fmaf(a[i].x, b.x, b.y); a[i].y = fmaf(a[i].y, b.x, b.y); if (i < 2) m = ex2a(m); }   // + 2 MUFU
            if (MODE == 3) { a[i] = ffma2(a[i], b, b); if (i < 2) m = ex2a(m); }
        }
    }
    float s = m;
    for (int i = 0; i < 8; i++) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name) {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(out, 100, 1.0001f);
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(out, iters, 1.0001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fma = 16.0 * iters * 148.0 * 8 * 256;
    printf("%-28s %.3f ms  %.1f TFLOP/s (fp32 FMA flops)  %.2f clk per 16 FMA per SMSP-warp-slot\n", name, ms, 2 * fma / ms / 1e9,
           ms * 1e-3 * 1.965e9 / (iters * 16.0));   // 16 warps per SMSP share the slot
    cudaFree(out);
}
// a loop shaped like the angular kernels: NF independent FFMA (or NF / 2 FFMA2), NA integer adds, NM independent MUFU.EX2
template <int NF, int NA, int NM, bool PACKED>
__global__ void mix(float* out, int iters, float seed) {
    float2 a[16];
    float m[NM > 0 ? NM : 1];
    int c[8];
    for (int i = 0; i < 16; i++) a[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.25f);
    for (int i = 0; i < NM; i++) m[i] = seed * (i + 1) * 1e-3f;
    for (int i = 0; i < 8; i++) c[i] = threadIdx.x + i;
    const float2 b = make_float2(seed, seed * 0.5f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < NF / 32; r++)
#pragma unroll
            for (int i = 0; i < 16; i++) {
                if (PACKED) a[i] = ffma2(a[i], b, b);
                else { a[i].x = fmaf(a[i].x, b.x, b.y); a[i].y = fmaf(a[i].y, b.x, b.y); }
            }
#pragma unroll
        for (int i = 0; i < NM; i++) m[i] = ex2a(m[i]);
#pragma unroll
        for (int i = 0; i < NA; i++) c[i & 7] += c[(i + 1) & 7] ^ it;
    }
    float s = 0;
    for (int i = 0; i < 16; i++) s += a[i].x + a[i].y;
    for (int i = 0; i < NM; i++) s += m[i];
    for (int i = 0; i < 8; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NF, int NA, int NM, bool PACKED>
void runmix(const char* name, int cpsm) {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    const int iters = 4000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    mix<NF, NA, NM, PACKED><<<148 * cpsm, 256>>>(out, 100, 1.0001f);
    cudaEventRecord(e0);
    mix<NF, NA, NM, PACKED><<<148 * cpsm, 256>>>(out, iters, 1.0001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // warps per SMSP = cpsm * 8 / 4; clocks per warp-iteration per SMSP slot
    printf("%-44s %d CTA/SM: %.3f ms  %.1f clk per warp-iteration per SMSP\n", name, cpsm, ms, ms * 1e-3 * 1.965e9 / (iters * (cpsm * 2.0)));
    cudaFree(out);
}
int main() {
    for (int cpsm : {3, 8}) {
        runmix<96, 0, 0, false>("96 FFMA", cpsm);
        runmix<96, 0, 17, false>("96 FFMA + 17 MUFU", cpsm);
        runmix<96, 0, 17, true>("48 FFMA2 + 17 MUFU", cpsm);
        runmix<64, 32, 17, false>("64 FFMA + 32 IADD/LOP + 17 MUFU", cpsm);
        runmix<64, 32, 0, false>("64 FFMA + 32 IADD/LOP", cpsm);
        runmix<32, 0, 17, false>("32 FFMA + 17 MUFU", cpsm);
        runmix<32, 0, 8, false>("32 FFMA + 8 MUFU", cpsm);
        runmix<0, 0, 17, false>("17 MUFU", cpsm);
    }
    run<0>("16 FFMA"); run<1>("8 FFMA2"); run<2>("16 FFMA + 2 MUFU"); run<3>("8 FFMA2 + 2 MUFU");
    return 0;
}
