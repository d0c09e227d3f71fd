// Shared device/host helpers for the nnpops_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

namespace nnpops {

#define NNP_CUDA_CHECK(expr)                                                                      \
    do {                                                                                          \
        cudaError_t err__ = (expr);                                                               \
        if (err__ != cudaSuccess)                                                                 \
            throw std::runtime_error(std::string("CUDA error ") + cudaGetErrorString(err__) +     \
                                     " at " __FILE__ ":" + std::to_string(__LINE__));             \
    } while (0)

#define NNP_REQUIRE(cond, msg)                                                                    \
    do {                                                                                          \
        if (!(cond)) throw std::runtime_error(std::string(msg));                                  \
    } while (0)

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;
constexpr float kPi = 3.14159265358979323846f;

// Periodic box + cell grid, computed on the device by geom_kernel so that no host read of the box is needed
// (the reference probes the box on the host: CudaANISymmetryFunctions.cu:341-359).
struct Geom {
    float box[9];      // row-major box vectors a, b, c (reduced form: a=(ax,0,0), b=(bx,by,0), c=(cx,cy,cz))
    float inv[3];      // 1/ax, 1/by, 1/cz in fp32 -- the reciprocal the reference multiplies by
    int periodic;      // 0/1
    int triclinic;     // any off-diagonal != 0
    int nc[3];         // cells per dimension (x, y, z); z is the fastest-varying index of a cell id
    int ncells;
    float origin[3];   // non-periodic: lower corner of the bounding box
    float cellInv[3];  // non-periodic: cells per unit length
    int anyOutside;    // periodic: set by the cell-list build when some atom lies outside the primary cell [0, 1)^3 (fractional)
};

// Minimum-image displacement, "multiply by reciprocal" flavour used by ANI and CFConv
// (CpuANISymmetryFunctions.cpp:355-379, CpuCFConv.cpp:30-55).  Explicit round-to-nearest intrinsics keep nvcc from
// contracting into FMAs, so r2 is bit-identical to the reference CPU arithmetic and cutoff decisions cannot differ.
// rintf (one FRND) replaces the reference's round(): the two differ only on exact ties |delta * inv| = k + 0.5, where both
// images have the same |delta| (= half the box), hence the same r2; such a pair can only be inside the cutoff when the box is
// narrower than two cutoffs, which the reference excludes.
__device__ __forceinline__ float min_image_mul(const Geom& g, float& dx, float& dy, float& dz) {
    if (g.periodic) {
        if (g.triclinic) {
            float s3 = rintf(__fmul_rn(dz, g.inv[2]));
            dx = __fsub_rn(dx, __fmul_rn(s3, g.box[6]));
            dy = __fsub_rn(dy, __fmul_rn(s3, g.box[7]));
            dz = __fsub_rn(dz, __fmul_rn(s3, g.box[8]));
            float s2 = rintf(__fmul_rn(dy, g.inv[1]));
            dx = __fsub_rn(dx, __fmul_rn(s2, g.box[3]));
            dy = __fsub_rn(dy, __fmul_rn(s2, g.box[4]));
            float s1 = rintf(__fmul_rn(dx, g.inv[0]));
            dx = __fsub_rn(dx, __fmul_rn(s1, g.box[0]));
        } else {
            dx = __fsub_rn(dx, __fmul_rn(rintf(__fmul_rn(dx, g.inv[0])), g.box[0]));
            dy = __fsub_rn(dy, __fmul_rn(rintf(__fmul_rn(dy, g.inv[1])), g.box[4]));
            dz = __fsub_rn(dz, __fmul_rn(rintf(__fmul_rn(dz, g.inv[2])), g.box[8]));
        }
    }
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// ex2.approx / lg2.approx / sin / cos fast paths, wrapped so call sites stay readable
__device__ __forceinline__ float fast_exp2(float x) { return exp2f(x); }   // with -use_fast_math -> ex2.approx.ftz
__device__ __forceinline__ float fast_log2(float x) { return __log2f(x); }

constexpr float kLog2e = 1.4426950408889634f;

// number of kernels this library has launched in this process (reported by bench.py as gpu_launches)
extern unsigned long long g_launches;
inline void count_launch(int n = 1) { g_launches += (unsigned long long)n; }

}  // namespace nnpops
