// Angular AEV kernels, second generation (factorised TorchANI tables: 8 radial shifts x 4 angular shifts, ANI-2x; 4 x 8, ANI-1x).
// Replaces the inner loops of CudaANISymmetryFunctions.cu:242-290 (forward) and :473-596 (backward); the mathematics is
// CpuANISymmetryFunctions.cpp:139-194 / :265-353.  See ani_angular_v2.cu for the design.
#pragma once
#include "ani_aev.cuh"

namespace nnpops {

struct AevOutPtr {      // one AEV element is written as fp32, or as the fp16 hi/lo pair of the tensor-core MLP
    float* f32;
    __half* hi;
    __half* lo;
};

constexpr int kSegBins = 512;      // segments are binned by their number of G-lane chunks (descending order in the list)
constexpr int kSegGroup = 4;       // lanes per segment in the forward kernel

// true when the tables / layout are the ones the v2 kernels are compiled for
bool angular_v2_supported(const AniTables& t);

// Segment list of one forward: every non-empty (centre, species pair) block, sorted by size (largest first).  Empty blocks are
// zero-filled here.  hist [kSegBins] and cursor [kSegBins] are scratch, adjacent in memory (hist first).
void angular_v2_build_segments(int n, const AniTables& tabHost, const int* offAng, const int* hist, int* cursor, int4* segs, int* nSeg,
                               const int* sortedOrig, const int* rowMap, AevOutPtr out, int stride, cudaStream_t stream);

void angular_v2_forward(int n, const AniTables& tabHost, const AniTables* tab, const int* offAng, int capA, const float4* geoA,
                        const float4* geoB, const int4* segs, const int* nSeg, const int* sortedOrig, const int* rowMap, AevOutPtr out,
                        int stride, cudaStream_t stream);

void angular_v2_backward(int n, const AniTables& tabHost, const AniTables* tab, const int* offAng, int capA, const float4* geoA,
                         const float4* geoB, const int* sortedOrig, const int* rowMap, const float* grad, int stride, float* posGrad,
                         bool padded, cudaStream_t stream);

// Radial path, second generation: the forward kernel evaluates the pair geometry once and the radial AEV with it; the backward
// kernel reads the geometry rows  radGeoA = {unit vector, r},  radGeoB = {fc, fc', species << 24 | atom index, -}.
// Supported for 16 radial functions with one EtaR (ANI-1x / ANI-2x).
bool radial_v2_supported(const AniTables& t);
void radial_v2_forward(int n, const AniTables& tabHost, const AniTables* tab, const float4* sorted, const int* sortedOrig, const Geom* geom,
                       const int* rowRad, const int* offRad, int capR, float4* radGeoA, float4* radGeoB, const int* rowMap, AevOutPtr out,
                       int stride, cudaStream_t stream);
// geometry rows of the angular neighbours (geoA / geoB) from the index rows of the row kernel
void angular_v2_geometry(int n, const AniTables& tabHost, const AniTables* tab, const float4* sorted, const int* sortedOrig, const Geom* geom,
                         const int* rowAng, const int* offAng, int capA, float4* geoA, float4* geoB, cudaStream_t stream);
void radial_v2_backward(int n, const AniTables& tabHost, const AniTables* tab, const int* offRad, int capR, const float4* radGeoA,
                        const float4* radGeoB, const int* sortedOrig, const int* rowMap, const float* grad, int stride, float* posGrad,
                        bool padded, cudaStream_t stream);
// posGrad of the backward kernels: [n][3] accumulated with scalar reductions, or (padded) a zeroed [n][4] buffer accumulated with one
// 16-byte vector reduction per force (red.global.add.v4.f32: a third of the L2 atomic operations) and copied out by grad_compact
void grad_compact(int n, const float4* acc, float* posGrad, cudaStream_t stream);

}  // namespace nnpops
