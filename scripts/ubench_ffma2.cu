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
int main() {
    run<0>("16 FFMA"); run<1>("8 FFMA2"); run<2>("16 FFMA + 2 MUFU"); run<3>("8 FFMA2 + 2 MUFU");
    return 0;
}
