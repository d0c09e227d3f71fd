// Cell-list neighbour search primitive shared by the ANI, CFConv, getNeighborPairs and PME-direct paths.
//
// The reference has no cell list: every search is an O(N^2) scan (CpuANISymmetryFunctions.cpp:114-125,
// CudaANISymmetryFunctions.cu:203-209, CpuCFConv.cpp:100-106, getNeighborPairsCUDA.cu:129-133).  This primitive must
// therefore reproduce the RESULT of those scans: every pair the reference accepts is visited (cells are >= 1.001 x cutoff wide
// measured perpendicular to the box faces), and the accept test itself is evaluated by the caller with the reference's own
// fp32 expression on the original coordinates.
//
// Layout in HBM: atoms are counting-sorted by cell id (z fastest) into `sorted` = float4{x, y, z, tag-bits}; inside a cell
// atoms are ordered by original index, so the order -- and every downstream floating-point sum -- is deterministic.
#pragma once
#include "common.cuh"

namespace nnpops {

struct CellList {
    int n = 0;
    int maxCells = 0;
    Geom* geom = nullptr;        // [1]
    int* cellCount = nullptr;    // [maxCells + 1]
    int* cellStart = nullptr;    // [maxCells + 1]
    int* cellOf = nullptr;       // [n] cell of original atom i
    int* slot = nullptr;         // [n] arrival slot inside the cell (unordered)
    int* tmpIdx = nullptr;       // [n] unordered cell-sorted original indices
    float4* sorted = nullptr;    // [n] x, y, z, tag
    int* sortedOrig = nullptr;   // [n] original index of sorted position p
    int* sortedCell = nullptr;   // [n] cell id of sorted position p

    void init(int numAtoms);
    void release();
    // positions: device [n][3] (float or double); box: device [3][3] of the same type or nullptr (non-periodic);
    // tags: device int [n] or nullptr.  All launches go to `stream`; nothing synchronises.
    // run (optional): device flag; when it reads 0 at execution time every kernel returns at once and the previous list stands
    // (Verlet-skin reuse decided on the device, no host round trip).
    template <typename T>
    void build(const T* positions, const T* box, const int* tags, float cutoff, cudaStream_t stream, const int* run = nullptr);
};

// Visit the candidate ranges [begin, end) of the sorted array that can hold neighbours of an atom in cell `c`.
// Cells adjacent along z are contiguous in memory, so a 27-cell neighbourhood is at most 18 runs (9 away from the z faces).
// f(begin, end, wrapped): `wrapped` tells whether the run was reached across a periodic face (or the grid is too coarse to tell);
// for an un-wrapped run of an orthorhombic box with >= 5 cells per dimension and all atoms inside the primary cell the
// minimum-image step of a pair (centre, candidate) subtracts exactly zero and may be skipped by the caller.
template <typename F>
__device__ __forceinline__ void for_each_candidate_run_w(const Geom& g, const int* __restrict__ cellStart, int c, F&& f) {
    const int nx = g.nc[0], ny = g.nc[1], nz = g.nc[2];
    const int cz = c % nz;
    const int cy = (c / nz) % ny;
    const int cx = c / (nz * ny);
    const bool per = g.periodic != 0;
    // z runs
    int za[2], zb[2], nrun = 1;
    bool zw[2] = {false, false};
    if (!per) { za[0] = max(cz - 1, 0); zb[0] = min(cz + 1, nz - 1); }
    else if (nz <= 3) { za[0] = 0; zb[0] = nz - 1; zw[0] = true; }
    else if (cz == 0) { za[0] = 0; zb[0] = 1; za[1] = nz - 1; zb[1] = nz - 1; zw[1] = true; nrun = 2; }
    else if (cz == nz - 1) { za[0] = nz - 2; zb[0] = nz - 1; za[1] = 0; zb[1] = 0; zw[1] = true; nrun = 2; }
    else { za[0] = cz - 1; zb[0] = cz + 1; }
    const bool coarse = per && (nx < 5 || ny < 5 || nz < 5);
    const int x0 = (per && nx <= 2) ? 0 : -1, x1 = (per && nx == 1) ? 0 : 1;
    const int y0 = (per && ny <= 2) ? 0 : -1, y1 = (per && ny == 1) ? 0 : 1;
    for (int ox = x0; ox <= x1; ox++) {
        int ix = cx + ox;
        bool xw = false;
        if (per) { if (ix < 0) { ix += nx; xw = true; } else if (ix >= nx) { ix -= nx; xw = true; } } else if (ix < 0 || ix >= nx) continue;
        for (int oy = y0; oy <= y1; oy++) {
            int iy = cy + oy;
            bool yw = false;
            if (per) { if (iy < 0) { iy += ny; yw = true; } else if (iy >= ny) { iy -= ny; yw = true; } } else if (iy < 0 || iy >= ny) continue;
            const int base = (ix * ny + iy) * nz;
            for (int r = 0; r < nrun; r++) f(cellStart[base + za[r]], cellStart[base + zb[r] + 1], coarse || xw || yw || zw[r]);
        }
    }
}

template <typename F>
__device__ __forceinline__ void for_each_candidate_run(const Geom& g, const int* __restrict__ cellStart, int c, F&& f) {
    for_each_candidate_run_w(g, cellStart, c, [&](int b, int e, bool) { f(b, e); });
}

}  // namespace nnpops
