// Reference arithmetic of a pair displacement as neighbors::getNeighborPairs defines it (getNeighborPairsCPU.cpp:56-98): delta =
// pos[row] - pos[col], minimum image by DIVISION, sequentially z, y, x, every step rounded to nearest (no FMA contraction), so that
// distances and cutoff decisions are bit-identical to the reference.  Shared by getNeighborPairs (neighbors.cu) and the fused
// PME direct-space kernel (pme_direct_fused.cu), which must accept exactly the pairs of that list.
#pragma once
#include "common.cuh"

namespace nnpops {

template <typename T> struct Arith;
template <> struct Arith<float> {
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float rnd(float a) { return rintf(a); }     // torch.round: half to even
    static __device__ __forceinline__ float sqrt(float a) { return __fsqrt_rn(a); }
    static __device__ __forceinline__ float nan() { return __int_as_float(0x7fc00000); }
};
template <> struct Arith<double> {
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ double rnd(double a) { return rint(a); }
    static __device__ __forceinline__ double sqrt(double a) { return __dsqrt_rn(a); }
    static __device__ __forceinline__ double nan() { return __longlong_as_double(0x7ff8000000000000LL); }
};

template <typename T>
struct Box {
    T b[9];
    int periodic;
};

// delta = pos[row] - pos[col] with the reference's sequential minimum image; returns the distance
template <typename T>
__device__ __forceinline__ T pair_delta(const Box<T>& bx, const T* __restrict__ pr, const T* __restrict__ pc, T& dx, T& dy, T& dz) {
    using A = Arith<T>;
    dx = A::sub(pr[0], pc[0]); dy = A::sub(pr[1], pc[1]); dz = A::sub(pr[2], pc[2]);
    if (bx.periodic) {
        const T s3 = A::rnd(A::div(dz, bx.b[8]));
        dx = A::sub(dx, A::mul(s3, bx.b[6])); dy = A::sub(dy, A::mul(s3, bx.b[7])); dz = A::sub(dz, A::mul(s3, bx.b[8]));
        const T s2 = A::rnd(A::div(dy, bx.b[4]));
        dx = A::sub(dx, A::mul(s2, bx.b[3])); dy = A::sub(dy, A::mul(s2, bx.b[4])); dz = A::sub(dz, A::mul(s2, bx.b[5]));
        const T s1 = A::rnd(A::div(dx, bx.b[0]));
        dx = A::sub(dx, A::mul(s1, bx.b[0])); dy = A::sub(dy, A::mul(s1, bx.b[1])); dz = A::sub(dz, A::mul(s1, bx.b[2]));
    }
    return A::sqrt(A::add(A::add(A::mul(dx, dx), A::mul(dy, dy)), A::mul(dz, dz)));
}

}  // namespace nnpops
