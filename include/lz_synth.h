/*
 * lz_synth — synthetic benchmark inputs generated on the device (NOT part of the drop-in
 * boundary; used by bench.py and the tests only).
 *
 * BASELINE config 4 names an "irregular 3D random-geometric-graph Laplacian, 50M vertices,
 * ~14 nnz/row".  The reference has no generator for it (its irregular operators come from
 * IrrGrid/IrrLap point clouds of a few thousand points), and building 7e8 entries with
 * NumPy/SciPy on the host would take longer than the whole benchmark, so the graph is defined
 * by a counter-based hash and produced by kernels, one row block per GPU:
 *
 *   - space is a box of ncx x ncy x ncz unit cells; cell c = cx + ncx*(cy + ncy*cz);
 *   - cell c holds count(c) points, count = inverse-CDF of Poisson(lambda) at u(c) (capped at 31):
 *     a Poisson point process, i.e. uniformly distributed points; point k of cell c sits at
 *     (cx + u0, cy + u1, cz + u2), u_a = hash(seed, c, k, a) in [0,1);
 *   - vertices are numbered cell by cell (cell order = locality order), vertex = prefix(c) + k;
 *   - i ~ j  iff  (dx*dx + dy*dy) + dz*dz <= r2 (fp64, each operation rounded), r <= 1 cell;
 *   - operator: graph Laplacian L = D - A, rows sorted by column, diagonal = degree.
 * oracle/rgg_oracle.py restates the same definition in NumPy; tests compare patterns bit for bit.
 *
 * All pointers are device pointers on the current device; `stream` is a cudaStream_t.
 * Every function returns 0 on success, a cudaError_t value otherwise.
 */
#ifndef LZ_SYNTH_H
#define LZ_SYNTH_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct lzs_rgg {
    int32_t ncx, ncy, ncz;     /* cells per axis */
    int32_t reserved;
    uint64_t seed;
    double r2;                 /* squared connection radius in cell units (<= 1) */
    double cdf[32];            /* cdf[k] = P(count <= k) of the per-cell count; count = #{k : cdf[k] <= u} */
} lzs_rgg;

/* count_dev[c - c0] = points in cell c, c in [c0, c1) */
int lzs_rgg_cell_counts(const lzs_rgg* g, int64_t c0, int64_t c1, int32_t* count_dev, void* stream);
/* xyz_dev[3*(v - row0) + a] = coordinate a of vertex v in [row0, row1); prefix_dev[ncells + 1] */
int lzs_rgg_positions(const lzs_rgg* g, const int64_t* prefix_dev, int64_t row0, int64_t row1,
                      double* xyz_dev, void* stream);
/* nnz_row_dev[v - row0] = entries of row v of L (neighbours + the diagonal) */
int lzs_rgg_row_degrees(const lzs_rgg* g, const int64_t* prefix_dev, int64_t row0, int64_t row1,
                        int32_t* nnz_row_dev, void* stream);
/* rows [row0, row1) of L as CSR with GLOBAL column indices: indptr_dev[rows + 1] is the local
 * exclusive prefix of nnz_row, indices_dev / data_dev have indptr[rows] entries */
int lzs_rgg_fill(const lzs_rgg* g, const int64_t* prefix_dev, int64_t row0, int64_t row1,
                 const int32_t* indptr_dev, int32_t* indices_dev, double* data_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LZ_SYNTH_H */
