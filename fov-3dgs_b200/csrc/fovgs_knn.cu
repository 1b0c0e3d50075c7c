// fovgs_knn.cu — mean squared distance to the 3 nearest neighbours of every point.
//
// Replaces simple_knn's distCUDA2 (fov3dgs/submodules/simple-knn/simple_knn.cu:63-218, spatial.cu), which the reference
// calls once per model to initialise the Gaussian scales (scene/gaussian_model.py:20,256) — and imports at module import
// time, so no reference script starts without it.  The reference Morton-sorts the points and prunes 1024-point boxes; this
// is a different structure with the same (exact) answer: a RECTILINEAR hash grid — per-axis cell boundaries at the quantiles
// of the coordinate (4096-bin histograms), so reconstructed point clouds with a dense core and far outliers still put a few
// points in every cell — with per-cell linked lists built by atomic exchange (no sort, no scan, no host synchronisation),
// searched in growing cubes of cells until the third-best distance is provably final.  Init-time only, not on the frame path.
// Measured on a B200: 1 M uniform points 2.5 ms; the 6 M-point multi-scale bench scene (thin disk + dense blob + sparse shell
// out to r = 30) 4.4 s — the cube of CELLS grows slowly in real distance where another cluster's quantiles cut thin slabs
// through a sparse region.  Exact either way; a Morton/box structure like the reference's would be the fix if multi-million
// point initialisations mattered (COLMAP clouds are 10^5-10^6 points).
#include <float.h>
#include "fovgs_internal.cuh"

namespace fovgs {

struct KnnHeader {
    int mn[3], mx[3];     // bounding box as order-preserving ints
    int G;                // cells per axis
};

__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void k_knn_init(KnnHeader* h, int G, int* head, size_t ncell) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        for (int k = 0; k < 3; k++) { h->mn[k] = 0x7fffffff; h->mx[k] = (int)0x80000000; }
        h->G = G;
    }
    for (size_t c = i; c < ncell; c += (size_t)gridDim.x * blockDim.x) head[c] = -1;
}

__global__ void k_knn_bbox(int P, const float* __restrict__ pts, KnnHeader* h) {
    int mn[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, mx[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x)
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float v = pts[3 * (size_t)i + k];
            if (v == v) { const int o = f2ord(v); mn[k] = min(mn[k], o); mx[k] = max(mx[k], o); }
        }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        mn[k] = __reduce_min_sync(0xffffffffu, mn[k]);
        mx[k] = __reduce_max_sync(0xffffffffu, mx[k]);
    }
    if ((threadIdx.x & 31) == 0)
#pragma unroll
        for (int k = 0; k < 3; k++) { atomicMin(&h->mn[k], mn[k]); atomicMax(&h->mx[k], mx[k]); }
}

constexpr int KNN_BINS = 4096;
constexpr int KNN_GMAX = 256;

// per-axis histograms of the coordinates over the bounding box
__global__ void k_knn_hist(int P, const float* __restrict__ pts, const KnnHeader* __restrict__ h, unsigned* __restrict__ hist) {
    float lo[3], inv[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        lo[k] = ord2f(h->mn[k]);
        inv[k] = (float)KNN_BINS / fmaxf(ord2f(h->mx[k]) - lo[k], 1e-20f);
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x)
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int b = min(max((int)((pts[3 * (size_t)i + k] - lo[k]) * inv[k]), 0), KNN_BINS - 1);
            atomicAdd(&hist[k * KNN_BINS + b], 1u);
        }
}

// cell boundaries bnd[axis][0..G]: bin edges at which the cumulative count crosses multiples of P/G (one block, 3 warps used)
__global__ void k_knn_bounds(int P, const KnnHeader* __restrict__ h, const unsigned* __restrict__ hist, float* __restrict__ bnd) {
    const int axis = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (axis >= 3) return;
    const int G = h->G;
    const float lo = ord2f(h->mn[axis]), hi = ord2f(h->mx[axis]);
    const float bw = fmaxf(hi - lo, 1e-20f) / (float)KNN_BINS;
    float* out = bnd + axis * (KNN_GMAX + 1);
    if (lane == 0) {
        unsigned long long cum = 0;
        int g = 1;
        out[0] = lo;
        for (int b = 0; b < KNN_BINS && g < G; b++) {
            cum += hist[axis * KNN_BINS + b];
            // boundaries are strictly increasing bin edges: a bin heavier than P/G simply becomes one cell
            while (g < G && cum * (unsigned long long)G >= (unsigned long long)P * (unsigned long long)g) {
                const float e = lo + bw * (float)(b + 1);
                if (e > out[g - 1]) { out[g] = e; g++; } else break;
            }
        }
        for (; g < G; g++) out[g] = fmaxf(hi, out[g - 1]);   // unused tail cells collapse onto the upper bound
        out[G] = hi + fabsf(hi) * 1e-6f + 1e-30f;
    }
}

struct KnnGrid {
    const float* bnd;   // [3][KNN_GMAX + 1]
    int G;
};
__device__ __forceinline__ KnnGrid knn_grid(const KnnHeader* h, const float* bnd) {
    KnnGrid g;
    g.bnd = bnd;
    g.G = h->G;
    return g;
}
// largest c in [0, G-1] with bnd[c] <= v
__device__ __forceinline__ int knn_cell_coord(const KnnGrid& g, int k, float v) {
    const float* b = g.bnd + k * (KNN_GMAX + 1);
    int lo = 0, hi = g.G - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (b[mid] <= v) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__global__ void k_knn_insert(int P, const float* __restrict__ pts, const KnnHeader* __restrict__ h, const float* __restrict__ bnd,
                             int* head, int* next) {
    const KnnGrid g = knn_grid(h, bnd);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
        const int cx = knn_cell_coord(g, 0, pts[3 * (size_t)i]), cy = knn_cell_coord(g, 1, pts[3 * (size_t)i + 1]),
                  cz = knn_cell_coord(g, 2, pts[3 * (size_t)i + 2]);
        const size_t cell = ((size_t)cz * g.G + cy) * g.G + cx;
        next[i] = atomicExch(&head[cell], i);
    }
}

__global__ void k_knn_search(int P, const float* __restrict__ pts, const KnnHeader* __restrict__ h, const float* __restrict__ bnd,
                             const int* __restrict__ head, const int* __restrict__ next, float* __restrict__ out) {
    const KnnGrid g = knn_grid(h, bnd);
    const int G = g.G;
    const float* bx = bnd, *by = bnd + (KNN_GMAX + 1), *bz = bnd + 2 * (KNN_GMAX + 1);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
        const float px = pts[3 * (size_t)i], py = pts[3 * (size_t)i + 1], pz = pts[3 * (size_t)i + 2];
        const int cx = knn_cell_coord(g, 0, px), cy = knn_cell_coord(g, 1, py), cz = knn_cell_coord(g, 2, pz);
        float b0 = FLT_MAX, b1 = FLT_MAX, b2 = FLT_MAX;
        for (int r = 0; r < G; r++) {
            const int x0 = max(cx - r, 0), x1 = min(cx + r, G - 1), y0 = max(cy - r, 0), y1 = min(cy + r, G - 1),
                      z0 = max(cz - r, 0), z1 = min(cz + r, G - 1);
            for (int z = z0; z <= z1; z++)
                for (int y = y0; y <= y1; y++) {
                    const bool inner_row = (abs(z - cz) < r) && (abs(y - cy) < r);
                    for (int x = x0; x <= x1; x++) {
                        if (inner_row && abs(x - cx) < r) { x = cx + r - 1; continue; }   // interior of the cube: seen at smaller r
                        for (int j = head[((size_t)z * G + y) * G + x]; j >= 0; j = next[j]) {
                            if (j == i) continue;
                            const float dx = pts[3 * (size_t)j] - px, dy = pts[3 * (size_t)j + 1] - py, dz = pts[3 * (size_t)j + 2] - pz;
                            float d = dx * dx + dy * dy + dz * dz;
                            if (b0 > d) { const float t = b0; b0 = d; d = t; }
                            if (b1 > d) { const float t = b1; b1 = d; d = t; }
                            if (b2 > d) { b2 = d; }
                        }
                    }
                }
            // everything closer than the cube's nearest face has been seen; a face clamped to the grid boundary has no points beyond
            float dface = FLT_MAX;
            if (cx - r > 0) dface = fminf(dface, px - bx[cx - r]);
            if (cx + r < G - 1) dface = fminf(dface, bx[cx + r + 1] - px);
            if (cy - r > 0) dface = fminf(dface, py - by[cy - r]);
            if (cy + r < G - 1) dface = fminf(dface, by[cy + r + 1] - py);
            if (cz - r > 0) dface = fminf(dface, pz - bz[cz - r]);
            if (cz + r < G - 1) dface = fminf(dface, bz[cz + r + 1] - pz);
            if (dface == FLT_MAX) break;                               // the cube covers the whole grid
            if (b2 != FLT_MAX && dface > 0.0f && b2 <= dface * dface * 0.999f) break;  // third neighbour is final
        }
        out[i] = (b0 + b1 + b2) / 3.0f;
    }
}

static inline int knn_grid_res(int P) {
    // ~8 points per cell on average: the third neighbour is then usually final after the 3x3x3 cube around the point's cell
    // (with 2 per cell the search needed cubes of 5-7 cells: 4.5 s for 6 M points instead of the time below); <= 256^3 cells
    int G = (int)floor(cbrt((double)P / 8.0) + 0.5);
    if (G < 1) G = 1;
    if (G > KNN_GMAX) G = KNN_GMAX;
    return G;
}

}  // namespace fovgs

using namespace fovgs;

extern "C" size_t fovgs_knn_workspace_bytes(int32_t P) {
    if (P < 0) return 0;
    const int G = knn_grid_res(P);
    return 256 + 3 * KNN_BINS * 4 + 3 * (KNN_GMAX + 1) * 4 + 256 + (size_t)G * G * G * 4 + (size_t)P * 4 + 256;
}

// mean_dist2[i] = mean of the squared distances from point i to its 3 nearest neighbours (FLT_MAX-based when P < 4, like the
// reference).  All pointers are device memory; nothing synchronises.
extern "C" int fovgs_knn_mean_dist2(int32_t P, const float* points, float* mean_dist2, void* workspace, size_t workspace_bytes,
                                    void* stream) {
    if (P < 0 || (P > 0 && (!points || !mean_dist2 || !workspace))) return FOVGS_ERR_INVALID_ARG;
    if (P == 0) return 0;
    if (workspace_bytes < fovgs_knn_workspace_bytes(P)) return FOVGS_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int G = knn_grid_res(P);
    const size_t ncell = (size_t)G * G * G;
    KnnHeader* h = (KnnHeader*)workspace;
    unsigned* hist = (unsigned*)((char*)workspace + 256);
    float* bnd = (float*)(hist + 3 * KNN_BINS);
    int* head = (int*)((char*)workspace + 256 + 3 * KNN_BINS * 4 + ((3 * (KNN_GMAX + 1) * 4 + 255) / 256) * 256);
    int* next = head + ncell;
    const int blocks = 148 * 8;
    cudaMemsetAsync(hist, 0, 3 * KNN_BINS * 4, st);
    k_knn_init<<<blocks, 256, 0, st>>>(h, G, head, ncell);
    k_knn_bbox<<<blocks, 256, 0, st>>>(P, points, h);
    k_knn_hist<<<blocks, 256, 0, st>>>(P, points, h, hist);
    k_knn_bounds<<<1, 96, 0, st>>>(P, h, hist, bnd);
    k_knn_insert<<<blocks, 256, 0, st>>>(P, points, h, bnd, head, next);
    k_knn_search<<<(P + 127) / 128, 128, 0, st>>>(P, points, h, bnd, head, next, mean_dist2);
    return cudaGetLastError() == cudaSuccess ? 0 : FOVGS_ERR_CUDA;
}
