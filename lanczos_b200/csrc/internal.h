// Internal data structures shared by the translation units of liblanczos_b200.
#pragma once
#include <vector>
#include "common.cuh"
#include "fin.cuh"

enum lz_op_kind { LZ_OP_STENCIL = 0, LZ_OP_CSR = 1, LZ_OP_SELL = 2 };

struct lz_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    int sms = 148;
    double* partials = nullptr;    // 2 * kMaxPartials doubles: CTA partial sums of streaming kernels
    double* scratch = nullptr;     // 64 doubles of device scratch (lz_dot, lz_reorthogonalize, ...)
    int* kba_done = nullptr;           // per-z-chunk completion counters of the KBA kernel (kba.cu); zero between launches
    unsigned int* tickets = nullptr;   // last-CTA tickets of the fin tails (fin.cuh); zero between launches
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    // second stream: the interior part of a row shard's SpMV runs here while the ghost entries and the beta
    // sum travel over NVLink on `stream` (created on first use; lower priority than nothing - see lanczos.cu)
    cudaStream_t gstream = nullptr;            // capture stream of the CUDA-graph path (lanczos.cu)
    std::vector<struct lz_graph_slot*> graphs; // captured solves, keyed on every pointer / scalar they bake in
    cudaStream_t stream2 = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    void* arena = nullptr;         // grow-only device workspace of lz_lanczos_run
    size_t arena_bytes = 0;
    std::vector<cudaEvent_t> event_pool;   // profile mode only
};

// Geometry of a structured-grid operator as the kernels see it.  `nz` is the number of
// z-planes held locally; the planes below/above the local slab come from `zlo`/`zhi`
// offsets (single shard: periodic wrap inside the same vector, or absent for Dirichlet)
// or from ghost buffers filled by the neighbouring shard.
struct lz_stencil {
    int dim = 3;
    int64_t nx = 1, ny = 1, nz = 1;
    int bc = LZ_BC_PERIODIC;
    double center = 0.0, offx = 0.0, offy = 0.0, offz = 0.0;
    int points = 7;                    // 7: axis neighbours only; 27: the full 3x3x3 box (3-D only)
    double w27[4] = {0, 0, 0, 0};      // 27-point weights by number of non-zero offsets: centre, face, edge, corner
    const double* diag = nullptr;
    // sharded execution: ghost planes (nx*ny doubles each) written by the z-neighbours.
    // When null, the kernel wraps (periodic) or drops (Dirichlet) the out-of-slab plane.
    const double* ghost_lo = nullptr;
    const double* ghost_hi = nullptr;
    int sharded = 0;
};

struct lz_csr {
    int64_t nnz = 0;
    int32_t* indptr = nullptr;     // device, M+1
    int32_t* indices = nullptr;    // device, nnz
    double* data = nullptr;        // device, nnz
    int lanes_per_row = 8;
};

struct lz_sell {
    int64_t nnz_true = 0, nnz_stored = 0;
    int64_t nchunks = 0;           // chunks of 32 rows
    int sigma = 0;
    int64_t* chunk_off = nullptr;  // device, nchunks+1: start of each chunk in col/val (elements)
    int32_t* col = nullptr;        // device, nnz_stored, column-major inside a chunk
    double* val = nullptr;         // device, nnz_stored
    int32_t* row_of = nullptr;     // device, nchunks*32: original row handled by (chunk, lane); -1 = padding row
    // value-free form (every off-diagonal entry equals uni_a, e.g. an unweighted graph Laplacian): the kernel
    // reads no values, deff[row] corrects the slots with col == row (diagonal entry, padding)
    int uniform = 0;
    double uni_a = 0.0;
    double* deff = nullptr;        // device, M
    // row shards: spans (runs of `split_span` chunks = one sorting window) whose rows touch no ghost column
    // ("interior": can be applied before the ghost exchange has completed) and the others ("boundary")
    int split_span = 0;            // chunks per interior work item (0: not classified)
    int bnd_span = 0;              // chunks per boundary work item
    int32_t* spans_int = nullptr;  // device, n_int span indices, ascending
    int32_t* spans_bnd = nullptr;  // device, n_bnd span indices, ascending
    int n_int = 0, n_bnd = 0;
    int np_int = 0;                // CTAs (= partials) of the last interior launch
    // windowed form (sellw.cu): per sorting window the sorted list of 32-entry granules of x it refers to, and
    // 16-bit local indices (granule rank * 32 + col % 32) for every stored entry
    int windowed = 0;              // 1: every window; 2: row shard, only the interior windows have lists
    int win_span = 0;              // chunks per window (= sigma / 32)
    int64_t win_count = 0;         // windows
    int win_maxg = 0;              // largest granule list
    int64_t win_total = 0;         // granules over all windows
    size_t win_smem = 0;           // stage size of the kernel (win_maxg * 256 B)
    int32_t* win_gran_off = nullptr;   // device, win_count + 1
    int32_t* win_gran = nullptr;       // device, win_total
    void* win_lcol = nullptr;          // device, 16-bit stage indices in blocks of [32 lanes][8 entries]
    int64_t* win_off8 = nullptr;       // device, nchunks + 1: first block of each chunk
    int64_t win_blocks = 0;
    int win_banked = 0;                // entries of a row stored in the bank-aware order (value-free form)
    void* win_lrow = nullptr;          // device, nchunks*32 x uint32: row inside its window | stage index of x[row] << 16
    double* win_deff = nullptr;        // device, nchunks*32: deff in (chunk, lane) order (value-free form)
};

struct lz_op {
    lz_ctx* ctx = nullptr;
    lz_op_kind kind = LZ_OP_STENCIL;
    int64_t M = 0;
    lz_stencil st;
    lz_csr csr;
    lz_sell sell;
    unsigned long long serial = 0; // unique per operator ever created (keys the captured-solve cache)
    int fused_per_sm = 0;          // cached occupancy of the fused step kernel
    int kb_zc = 0;                 // z-chunk length of the last KB launch that accumulated alpha (border kernel)
    int64_t ncols = 0;             // sparse: number of columns (> M for a row shard: M owned + ghosts)
    const double* xghost = nullptr;   // sparse row shard: where the ghost entries of x live (set per launch)
    // host copy of the CSR arrays is NOT kept; export reads them back from the device.
};

namespace lz {

// Destination of the halo planes of a vector being produced (sharded structured grids): the
// first `plane` elements go to lo_dst, the last `plane` elements to hi_dst (peer memory).
struct HaloPush {
    double* lo_dst = nullptr;
    double* hi_dst = nullptr;
    int64_t plane = 0;
};

// ---- operator apply with fused dot:  y = s * (H x),  partials[cta] = sum y * (s*x) -----
// `scale_dev` (nullable => 1) points at a device double.  `partials` has room for
// kMaxPartials doubles; *nparts receives the number written.
// `fin` (nullable): bookkeeping the last CTA to finish runs on the summed partials (fin.cuh).
int launch_apply_dot(lz_op* op, const double* x, const double* scale_dev, double* y,
                     double* partials, int* nparts, int* launches, const int* flag_dev = nullptr,
                     const FinTail* fin = nullptr);
// Row shard in SELL form, split by spans (lz_sell::spans_int / spans_bnd): part 1 = interior spans
// (partials[0 .. np_int)), part 2 = boundary spans (partials[np_int ..), *nparts = np_int + its CTAs; a fin
// tail sums both ranges).  `stream`: where to launch (the context's second stream for the interior part).
bool spmv_split_supported(const lz_op* op);
int sell_classify_spans(lz_op* op);
int sell_detect_uniform(lz_op* op);
int sellw_build(lz_op* op);
int scan_i64(int64_t* a_dev, int64_t n, cudaStream_t q);     // exclusive prefix in place (set-up only)
int launch_spmv_windowed(lz_op* op, const int32_t* win_list, int nlist, const double* x, const double* scale_dev,
                         double* y, double* partials, int* grid_out, const int* flag_dev, const FinTail& ft,
                         cudaStream_t stream);
int launch_spmv_part(lz_op* op, int part, const double* x, const double* scale_dev, double* y, double* partials,
                     int* nparts, const int* flag_dev, const FinTail* fin, cudaStream_t stream);

// ---- "recompute" step of matrix-free operators (stencil.cu, stencil27.cu) -------------------
// KA: launch_apply_dot with y == nullptr reduces alpha without writing w.
// KB: out = s * (H x) - (ca*sa) * x - (cb*sb) * b, partials[cta] = sum out^2 (b nullable).
struct StencilUpdate {
    const double* b = nullptr;
    const double* ca = nullptr;
    const double* sa = nullptr;
    const double* cb = nullptr;
    const double* sb = nullptr;
    HaloPush halo;             // sharded structured grids: boundary planes of `out` -> the neighbours' ghost buffers
    // alpha of the vector being produced, accumulated while it is still in registers (stencil.cu, KB + border
    // kernel): partials of out . H out over the edges inside a CTA tile go to alpha_partials[cta] (nullable)
    double* alpha_partials = nullptr;
};
bool recompute_step_supported(const lz_op* op);
bool recompute_step_preferred(const lz_op* op);
int launch_apply_update_norm(lz_op* op, const double* x, const double* scale_dev, const StencilUpdate* upd,
                             double* out, double* partials, int* nparts, int* launches,
                             const FinTail* fin = nullptr);
// alpha inside KB: supported for 7-point operators on grids made of whole 64 x 8 tiles
bool update_alpha_supported(const lz_op* op, const double* x, const double* out);
// the edges KB could not reach (across tile borders and z-chunks, slab top): partials[cta] of out . H out over them
int launch_alpha_border(lz_op* op, const double* v, double* partials, int* nparts, const FinTail* fin);

// ---- KB with the alpha reduction of its output chasing it through L2 (kba.cu) ------------------------
bool kba_step_supported(const lz_op* op, const double* x, const double* b, const double* out);
int launch_kba_step(lz_op* op, const double* x, const double* scale_dev, const StencilUpdate* upd, double* out,
                    const FinTail* fin_beta, const FinOp* fin_alpha, int* nparts);

// ---- the whole solve of a small matrix-free problem in one persistent cooperative kernel (small.cu) ----
bool small_solve_supported(const lz_op* op, const lz_run_opts* opts, int32_t n, const double* V_dev, int64_t ldv);
int launch_small_solve(lz_ctx* ctx, lz_op* op, const double* v0_dev, int32_t n, const lz_run_opts* opts,
                       double* alpha_host, double* beta_host, double* V_dev, int64_t ldv, double* row_scale_host,
                       lz_run_info* info);

// ---- single-pass fused step (fused.cu) ---------------------------------------------------
bool fused_step_supported(const lz_op* op);
int launch_fused_step(lz_op* op, const double* u, const double* rj, const double* rjm1,
                      const double* s_j, const double* alpha_j, const double* beta_j, const double* s_jm1,
                      double* out_r, double* out_u, double* partials, int* nparts);

// ---- streaming vector kernels (vecops.cu) ------------------------------------------------
// partials[cta] = sum x*y
int launch_dot(lz_ctx* ctx, const double* x, const double* y, int64_t M, double* partials, int* nparts,
               const FinTail* fin = nullptr);
// out = w - ca*sa*a - cb*sb*b with ca/sa/cb/sb read from device memory (nullable b),
// partials[cta] = sum out^2.  In-place (out == w) is allowed.
int launch_update_norm(lz_ctx* ctx, const double* w, const double* a, const double* b,
                       const double* ca_dev, const double* sa_dev, const double* cb_dev,
                       const double* sb_dev, double* out, int64_t M, double* partials, int* nparts,
                       const HaloPush* halo = nullptr, const FinTail* fin = nullptr,
                       const double* sw_dev = nullptr /* w is scaled by *sw_dev first (nullable: 1) */);
int launch_halo_push(lz_ctx* ctx, const double* x, int64_t M, const HaloPush* halo);
int launch_ghost_push(lz_ctx* ctx, const double* x, const int32_t* send_idx, int nsend, int world,
                      const int* seg_start, double* const* dst, const int* flag_dev, cudaStream_t stream = nullptr);
int launch_scale(lz_ctx* ctx, double* x, int64_t M, double s);

// ---- Gram-Schmidt block GEMV pair (reorth.cu) --------------------------------------------
// dots: part[(r * ncg) + g] = partial of V[r,:] . V[j,:]   for r in [0, nrows)   (nrows <= j+1)
// `tail` (nullable): the last CTA turns the partials into the sweep's coefficients (fin.cuh IpTail)
int launch_cgs_dots(lz_ctx* ctx, const double* V, int64_t ldv, int nrows, const double* target,
                    int64_t M, double* part, int* ncg_out, const int* flag_dev, const IpTail* tail = nullptr);
// update: out = cself * target - sum_{r<nrows} coef[r] * V[r,:]   (coef, cself on the device)
int launch_cgs_update(lz_ctx* ctx, const double* V, int64_t ldv, int nrows, const double* target,
                      const double* coef_dev, const double* cself_dev, double* out, int64_t M,
                      const int* flag_dev, const HaloPush* halo = nullptr);
// fused middle of CGS2 (K4c): target <- cself * target - V_k coef in place, and the partials of
// V_k^T target_new, from one read of the k basis rows staged in shared memory
bool cgs_update_dots_supported(const double* V, int64_t ldv, int k, const double* target);
int launch_cgs_update_dots(lz_ctx* ctx, const double* V, int64_t ldv, int k, double* target,
                           const double* coef_dev, const double* cself_dev, int64_t M, double* part,
                           int* ncg_out, const int* flag_dev, const IpTail* tail = nullptr);
// Y[c,:] = sum_r S[r + c*n] * V[r,:]  (S on the device, n x k column-major)
int launch_ritz_lift(lz_ctx* ctx, const double* V, int64_t ldv, int n, int64_t M,
                     const double* S_dev, int k, double* Y, int64_t ldy);

}  // namespace lz
