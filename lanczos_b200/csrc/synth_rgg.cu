// Device-side generator of the synthetic random-geometric-graph Laplacian of BASELINE config 4
// (include/lz_synth.h has the definition).  Built into its own shared object, liblz_synth.so:
// benchmark/test input, not part of the Lanczos drop-in boundary.
//
// Every vertex re-derives the coordinates of its candidate neighbours (the points of the 27 cells
// around its own) from the counter-based hash, so no coordinate array is kept and a rank needs
// nothing from other ranks but the per-cell prefix of the point counts, which it recomputes.
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/lz_synth.h"

namespace {

constexpr int kThreads = 256;

__host__ __device__ inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
__host__ __device__ inline uint64_t cell_hash(uint64_t seed, int64_t c) {
    return mix64(seed + 0x9e3779b97f4a7c15ull * (uint64_t)(c + 1));
}
__device__ inline double unit53(uint64_t h) { return (double)(h >> 11) * 0x1.0p-53; }
__device__ inline double coord(uint64_t hc, int k, int a, int cell_coord) {
    const uint64_t h = mix64(hc ^ (0xd1342543de82ef95ull * (uint64_t)(4 * k + a + 1)));
    return __dadd_rn((double)cell_coord, unit53(h));
}

struct Params {
    int ncx, ncy, ncz;
    uint64_t seed;
    double r2;
    double cdf[32];
};

__global__ void __launch_bounds__(kThreads)
cell_counts_kernel(Params p, int64_t c0, int64_t c1, int32_t* __restrict__ count) {
    const int64_t c = c0 + (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (c >= c1) return;
    const double u = unit53(cell_hash(p.seed, c));
    int n = 0;
#pragma unroll
    for (int k = 0; k < 31; ++k) n += (p.cdf[k] <= u);
    count[c - c0] = n;
}

// cell that owns vertex v: largest c with prefix[c] <= v
__device__ inline int64_t cell_of(const int64_t* __restrict__ prefix, int64_t ncells, int64_t v) {
    int64_t lo = 0, hi = ncells;          // prefix[lo] <= v < prefix[hi]
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(prefix + mid) <= v) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(kThreads)
positions_kernel(Params p, const int64_t* __restrict__ prefix, int64_t row0, int64_t row1, double* __restrict__ xyz) {
    const int64_t v = row0 + (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (v >= row1) return;
    const int64_t ncells = (int64_t)p.ncx * p.ncy * p.ncz;
    const int64_t c = cell_of(prefix, ncells, v);
    const int k = (int)(v - __ldg(prefix + c));
    const int cx = (int)(c % p.ncx), cy = (int)((c / p.ncx) % p.ncy), cz = (int)(c / ((int64_t)p.ncx * p.ncy));
    const uint64_t hc = cell_hash(p.seed, c);
    xyz[3 * (v - row0) + 0] = coord(hc, k, 0, cx);
    xyz[3 * (v - row0) + 1] = coord(hc, k, 1, cy);
    xyz[3 * (v - row0) + 2] = coord(hc, k, 2, cz);
}

// FILL == false: nnz_row[v - row0] = neighbours + 1.   FILL == true: write the row.
template <bool FILL>
__global__ void __launch_bounds__(kThreads)
rows_kernel(Params p, const int64_t* __restrict__ prefix, int64_t row0, int64_t row1,
            int32_t* __restrict__ nnz_row, const int32_t* __restrict__ indptr,
            int32_t* __restrict__ indices, double* __restrict__ data) {
    const int64_t v = row0 + (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (v >= row1) return;
    const int64_t ncells = (int64_t)p.ncx * p.ncy * p.ncz;
    const int64_t c = cell_of(prefix, ncells, v);
    const int k = (int)(v - __ldg(prefix + c));
    const int cx = (int)(c % p.ncx), cy = (int)((c / p.ncx) % p.ncy), cz = (int)(c / ((int64_t)p.ncx * p.ncy));
    const uint64_t hc = cell_hash(p.seed, c);
    const double x = coord(hc, k, 0, cx), y = coord(hc, k, 1, cy), z = coord(hc, k, 2, cz);
    int deg = 0;
    int64_t at = FILL ? (int64_t)indptr[v - row0] : 0;
    int64_t diag_at = -1;
    // neighbour cells in ascending cell id => columns come out sorted
    for (int dz = -1; dz <= 1; ++dz) {
        const int qz = cz + dz;
        if (qz < 0 || qz >= p.ncz) continue;
        for (int dy = -1; dy <= 1; ++dy) {
            const int qy = cy + dy;
            if (qy < 0 || qy >= p.ncy) continue;
            for (int dx = -1; dx <= 1; ++dx) {
                const int qx = cx + dx;
                if (qx < 0 || qx >= p.ncx) continue;
                const int64_t q = qx + (int64_t)p.ncx * (qy + (int64_t)p.ncy * qz);
                const int64_t q0 = __ldg(prefix + q);
                const int nq = (int)(__ldg(prefix + q + 1) - q0);
                const uint64_t hq = cell_hash(p.seed, q);
                for (int m = 0; m < nq; ++m) {
                    if (q == c && m == k) {
                        if (FILL) { diag_at = at; indices[at] = (int32_t)v; ++at; }
                        continue;
                    }
                    const double ex = __dadd_rn(x, -coord(hq, m, 0, qx));
                    const double ey = __dadd_rn(y, -coord(hq, m, 1, qy));
                    const double ez = __dadd_rn(z, -coord(hq, m, 2, qz));
                    const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez));
                    if (d2 <= p.r2) {
                        ++deg;
                        if (FILL) { indices[at] = (int32_t)(q0 + m); data[at] = -1.0; ++at; }
                    }
                }
            }
        }
    }
    if (FILL) data[diag_at] = (double)deg;
    else nnz_row[v - row0] = deg + 1;
}

Params make_params(const lzs_rgg* g) {
    Params p;
    p.ncx = g->ncx; p.ncy = g->ncy; p.ncz = g->ncz;
    p.seed = g->seed;
    p.r2 = g->r2;
    for (int k = 0; k < 32; ++k) p.cdf[k] = g->cdf[k];
    return p;
}

inline unsigned blocks_for(int64_t n) { return (unsigned)((n + kThreads - 1) / kThreads); }

}  // namespace

extern "C" {

int lzs_rgg_cell_counts(const lzs_rgg* g, int64_t c0, int64_t c1, int32_t* count_dev, void* stream) {
    if (!g || !count_dev || c1 < c0) return (int)cudaErrorInvalidValue;
    if (c1 == c0) return 0;
    cell_counts_kernel<<<blocks_for(c1 - c0), kThreads, 0, (cudaStream_t)stream>>>(make_params(g), c0, c1, count_dev);
    return (int)cudaGetLastError();
}

int lzs_rgg_positions(const lzs_rgg* g, const int64_t* prefix_dev, int64_t row0, int64_t row1,
                      double* xyz_dev, void* stream) {
    if (!g || !prefix_dev || !xyz_dev || row1 < row0) return (int)cudaErrorInvalidValue;
    if (row1 == row0) return 0;
    positions_kernel<<<blocks_for(row1 - row0), kThreads, 0, (cudaStream_t)stream>>>(make_params(g), prefix_dev, row0, row1, xyz_dev);
    return (int)cudaGetLastError();
}

int lzs_rgg_row_degrees(const lzs_rgg* g, const int64_t* prefix_dev, int64_t row0, int64_t row1,
                        int32_t* nnz_row_dev, void* stream) {
    if (!g || !prefix_dev || !nnz_row_dev || row1 < row0) return (int)cudaErrorInvalidValue;
    if (row1 == row0) return 0;
    rows_kernel<false><<<blocks_for(row1 - row0), kThreads, 0, (cudaStream_t)stream>>>(
        make_params(g), prefix_dev, row0, row1, nnz_row_dev, nullptr, nullptr, nullptr);
    return (int)cudaGetLastError();
}

int lzs_rgg_fill(const lzs_rgg* g, const int64_t* prefix_dev, int64_t row0, int64_t row1,
                 const int32_t* indptr_dev, int32_t* indices_dev, double* data_dev, void* stream) {
    if (!g || !prefix_dev || !indptr_dev || !indices_dev || !data_dev || row1 < row0) return (int)cudaErrorInvalidValue;
    if (row1 == row0) return 0;
    rows_kernel<true><<<blocks_for(row1 - row0), kThreads, 0, (cudaStream_t)stream>>>(
        make_params(g), prefix_dev, row0, row1, nullptr, indptr_dev, indices_dev, data_dev);
    return (int)cudaGetLastError();
}

}  // extern "C"
