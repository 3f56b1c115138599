/*
 * lanczos_b200 — C ABI of the B200-native Lanczos tridiagonalization path.
 *
 * This is the drop-in boundary for the hot path of jgslunde/Lanczos.  The reference
 * is pure Python and has no FFI of its own: its seam is the `use_cuda=True` backend
 * switch inside `Lanczos.execute_Lanczos` (Python/Regular/Lanczos.py:85-91) and
 * `IrrLanczos.execute_LanczosOld` (Python/Irregular/IrrLanczos.py:203-208), where
 * NumPy/SciPy objects are swapped for CuPy ones.  Each entry point below names the
 * reference lines whose work it takes over.  INTEGRATION.md shows the ctypes stub a
 * reference maintainer would add at that seam.
 *
 * Conventions
 *  - every function returns an int status (LZ_OK == 0); the message of the last
 *    failure on the calling thread is available from lz_last_error();
 *  - no exceptions, no longjmp, no torch types: plain pointers and sizes only;
 *  - pointers named *_dev are CUDA device pointers on the context's device, pointers
 *    named *_host are ordinary host pointers; the caller owns every buffer it passes;
 *  - all GPU work is enqueued on the context's stream; functions documented as
 *    "synchronises" wait for that stream before returning;
 *  - a context is not thread-safe; different contexts are independent;
 *  - all arithmetic is IEEE fp64; indices are int32 (as scipy.sparse CSR), sizes int64.
 */
#ifndef LANCZOS_B200_H
#define LANCZOS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LZ_ABI_VERSION 2

/* status codes */
#define LZ_OK               0
#define LZ_ERR_INVALID      1   /* bad argument (the Python mirror raises ValueError) */
#define LZ_ERR_CUDA         2   /* a CUDA runtime call failed */
#define LZ_ERR_NOMEM        3   /* device allocation failed */
#define LZ_ERR_BREAKDOWN    4   /* beta == 0 or non-finite: Krylov space exhausted */
#define LZ_ERR_UNSUPPORTED  5
#define LZ_ERR_PEER         6   /* multi-GPU exchange timed out / peer failure */

/* boundary conditions of a structured-grid operator */
#define LZ_BC_PERIODIC   0      /* reference convention, Hamiltonian.py:92-97 */
#define LZ_BC_DIRICHLET  1      /* 1Dbox.py:15-22 */

/* sparse storage used on the device */
#define LZ_FMT_CSR   0          /* CSR, sub-warp per row */
#define LZ_FMT_SELL  1          /* SELL-C-sigma, C = 32; operators whose off-diagonal entries are all equal (an
                                   unweighted graph Laplacian) are applied without reading their values */
#define LZ_FMT_SELL_VALUES 2    /* SELL-C-sigma, always with the stored values (comparison runs, tests) */

/* re-orthogonalisation policy */
#define LZ_REORTH_NONE       0
#define LZ_REORTH_FULL       1  /* every step, like Lanczos.reorthogonalize (Lanczos.py:233-251) */
#define LZ_REORTH_SELECTIVE  2  /* omega-recurrence monitor, fires at sqrt(eps) */

typedef struct lz_ctx lz_ctx;   /* one per (device, stream) */
typedef struct lz_op lz_op;     /* an operator H living on a context */
typedef struct lz_team lz_team; /* the set of row shards that run one distributed solve */

/* ---- library ---------------------------------------------------------------- */
int lz_abi_version(void);
const char* lz_last_error(void);
int lz_device_count(int* count);

/* ---- context ----------------------------------------------------------------
 * Replaces: the implicit CuPy device/stream that `import cupy as np` binds at
 * Lanczos.py:85-88.  `cuda_stream` is a cudaStream_t (NULL = legacy default stream). */
int lz_ctx_create(int device, void* cuda_stream, lz_ctx** out);
int lz_ctx_destroy(lz_ctx* ctx);
int lz_ctx_sync(lz_ctx* ctx);

/* ---- operators --------------------------------------------------------------
 * lz_op_stencil_create: matrix-free structured-grid operator
 *     (H x)_i = (center + diag_i) x_i + sum_axis offdiag[axis] (x_{i+e} + x_{i-e}),
 * index map i = x + nx*(y + ny*z) (Hamiltonian.py:73-84).  Replaces the CSR that
 * Hamiltonian.create_sparse_T("7") + create_sparse_V build (Hamiltonian.py:35-69) and
 * that Lanczos.py:88 uploads.  dim in {1,2,3}; shape[dim]; offdiag[dim]; diag_dev may
 * be NULL, otherwise M doubles on the device (kept by reference, not copied).
 * An axis with offdiag == 0 is skipped. */
int lz_op_stencil_create(lz_ctx* ctx, int dim, const int64_t* shape, int bc,
                         double center, const double* offdiag,
                         const double* diag_dev, lz_op** out);

/* lz_op_stencil27_create: matrix-free 27-point operator on a 3-D grid, the reference's default
 * Laplacian (Hamiltonian.create_sparse_T(points="27"), Hamiltonian.py:102-128):
 * weights[4] = coefficient of the centre / the 6 face / the 12 edge / the 8 corner neighbours
 * (reference T: T_factor * 3/13 * {-44/3, 1, 1/2, 1/3}; H = -T + V, 3Ddeuteron.py:80). */
int lz_op_stencil27_create(lz_ctx* ctx, const int64_t* shape, int bc, const double* weights,
                           const double* diag_dev, lz_op** out);

/* lz_op_csr_create: general sparse operator from host CSR arrays (scipy layout:
 * indptr[M+1], indices[nnz], data[nnz]).  Replaces cupyx.scipy.sparse.csr_matrix(H)
 * at Lanczos.py:88 / csc_matrix(H) at IrrLanczos.py:205.  The arrays are copied
 * (and, for LZ_FMT_SELL, converted by the library); the caller keeps its own.
 * sigma = sorting window in rows (multiple of 32; 0 -> library default). */
int lz_op_csr_create(lz_ctx* ctx, int64_t M, int64_t nnz, const int32_t* indptr_host,
                     const int32_t* indices_host, const double* data_host,
                     int fmt, int sigma, lz_op** out);

/* Row shard of a sparse operator for a team run: M_local rows, columns renumbered so that
 * [0, M_local) are the owned entries of x and [M_local, ncols) index the shard's ghost list. */
int lz_op_csr_shard_create(lz_ctx* ctx, int64_t M_local, int64_t ncols, int64_t nnz,
                           const int32_t* indptr_host, const int32_t* indices_host,
                           const double* data_host, int fmt, int sigma, lz_op** out);

/* As lz_op_csr_create / lz_op_csr_shard_create for CSR arrays that already live on the context's
 * device - the reference's GPU mode holds H as a cupyx matrix in device memory (Lanczos.py:88,
 * read back with H.get() at :137).  The arrays are copied / converted on the device (SELL: same
 * layout, bit for bit, as the host conversion); ncols == M for a whole operator, > M for a row
 * shard with renumbered ghost columns.  Synchronises. */
int lz_op_csr_create_dev(lz_ctx* ctx, int64_t M, int64_t ncols, int64_t nnz, const int32_t* indptr_dev,
                         const int32_t* indices_dev, const double* data_dev, int fmt, int sigma,
                         lz_op** out);

/* Diagonal potential evaluated on the device: out_dev[i + nx*(j + ny*k)] = f(x[i], y[j], z[k]) for
 * a postfix program f over the coordinates and constants (op codes below; CONST carries the index
 * of its constant in bits 8..).  Replaces the three nested Python loops of
 * Hamiltonian.create_sparse_V (Hamiltonian.py:35-46); the result is what lz_op_stencil*_create
 * take as diag_dev.  Arithmetic ops round once each (no fma contraction), like NumPy's ufuncs.
 * shape[3]; x/y/z_host: the coordinate axes (Hamiltonian.py:15-17).  Synchronises. */
#define LZ_POT_X       0
#define LZ_POT_Y       1
#define LZ_POT_Z       2
#define LZ_POT_CONST   3
#define LZ_POT_ADD     4    /* binary: ADD SUB MUL DIV POW MIN MAX */
#define LZ_POT_SUB     5
#define LZ_POT_MUL     6
#define LZ_POT_DIV     7
#define LZ_POT_POW     8
#define LZ_POT_MIN     9
#define LZ_POT_MAX     10
#define LZ_POT_NEG     11   /* unary: NEG SQRT EXP LOG ABS SIN COS TANH SQUARE */
#define LZ_POT_SQRT    12
#define LZ_POT_EXP     13
#define LZ_POT_LOG     14
#define LZ_POT_ABS     15
#define LZ_POT_SIN     16
#define LZ_POT_COS     17
#define LZ_POT_TANH    18
#define LZ_POT_SQUARE  19
int lz_potential_eval(lz_ctx* ctx, const int64_t* shape, const double* x_host, const double* y_host,
                      const double* z_host, int32_t nops, const int32_t* ops_host, int32_t nconsts,
                      const double* consts_host, double* out_dev);

int lz_op_rows(const lz_op* op, int64_t* M);
int lz_op_nnz(const lz_op* op, int64_t* nnz_true, int64_t* nnz_stored);
/* *value_free = 1 when the operator is applied from its column indices alone (LZ_FMT_SELL, all off-diagonal
 * entries equal): 4 instead of 12 bytes of HBM traffic per stored entry. */
int lz_op_value_free(const lz_op* op, int32_t* value_free);
/* *granules = 0, or - when the operator is applied in the windowed SELL form (per sorting window the entries of x
 * it refers to are staged in shared memory, the stored column indices are 16-bit offsets into that stage) - the
 * largest number of 32-entry granules of x one window stages.  LZ_SELL_WINDOW=0 in the environment disables the form. */
int lz_op_windowed(const lz_op* op, int32_t* granules);

/* y = H x on the device (`H*V[j]`, Lanczos.py:108,116).  Enqueues only. */
int lz_op_apply(lz_op* op, const double* x_dev, double* y_dev);

/* Export the operator as sorted CSR with duplicates summed (what scipy holds after
 * `H.sort_indices()`, 3Ddeuteron.py:81) into host arrays, for bit-exact pattern
 * checks.  Call with indptr_host == NULL to query nnz only.  Synchronises. */
int lz_op_export_csr(lz_op* op, int64_t* nnz, int32_t* indptr_host,
                     int32_t* indices_host, double* data_host);
int lz_op_destroy(lz_op* op);

/* ---- the Lanczos loop ---------------------------------------------------------
 * Replaces Lanczos.py:100-119 (== IrrLanczos.py:217-238) including reorthogonalize.
 */
typedef struct lz_run_opts {
    int32_t reorth;        /* LZ_REORTH_*                                            */
    int32_t cgs_passes;    /* 1 or 2 classical Gram-Schmidt sweeps per reorth        */
    int32_t ref_compat;    /* 1: reproduce the reference loop exactly (pre-step that
                              discards v0, (2-|v|^2) form of the sweep, beta taken
                              before the sweep); 0: v0/|v0| is the first basis vector */
    int32_t profile;       /* 1: bracket the bandwidth kernels with CUDA events and report
                              their summed device times in lz_run_info (bench.py roofline) */
    int32_t step_kernel;   /* 0 auto (3 for matrix-free 3/5/7-point operators, else 1); 1 two-pass step (K1
                              apply+dot, K3 update+norm: 48*M B + the operator's own bytes);
                              2 single-pass fused step KF (40*M B; 3-D structured grids with
                              nx % 64 == 0, ny % 8 == 0, one GPU, reorth != full);
                              3 recompute step (matrix-free operators: KA reduces alpha without
                              writing H v, KB applies H again inside the update: 32*M B)       */
    int32_t flags;         /* 0 = defaults.  bit 0: do not fuse the middle of CGS2 (K4c: update of sweep 1 + dots of
                              sweep 2 from one read of the basis); bit 1 (with ref_compat): the sweep takes the
                              LZ_SWEEP_GPU form of Regular/Lanczos.py:236-238 instead of the (2 - |v|^2) form;
                              bit 2: recompute step with alpha accumulated inside KB + a border kernel instead
                              of a KA pass (measured slower); bit 3: sparse row shards
                              without the interior/boundary overlap on a second stream; bit 4: recompute step
                              as the single KBA kernel (measured slower); bit 5: small matrix-free problems
                              (<= 148*256*8 unknowns, reorth full / none) in ONE persistent cooperative kernel
                              (measured slower than the replayed CUDA graph: 30 vs 21.5 us/step at 200 x 200)  */
    double  breakdown_tol; /* stop when beta <= breakdown_tol * |alpha_0| (0: only 0/NaN) */
    double  select_tol;    /* selective: orthogonality level that triggers (0 -> sqrt(eps)) */
} lz_run_opts;

typedef struct lz_run_info {
    int32_t steps_done;    /* Lanczos steps completed (== n unless breakdown)        */
    int32_t reorth_count;  /* steps in which the Gram-Schmidt sweeps ran             */
    int32_t launches;      /* kernels launched by this call                          */
    int32_t apply_launches;   /* profile == 1: launches and summed device ms of ...  */
    int32_t update_launches;
    int32_t dots_launches;
    int32_t gsupd_launches;
    int32_t fused_launches;
    float   gpu_ms;        /* device time of the loop (CUDA events on the stream)    */
    float   apply_ms;      /* ... K1/K2 operator apply + alpha dot (KA when step_kernel == 3) */
    float   update_ms;     /* ... K3 three-term update + norm (KB when step_kernel == 3)      */
    float   dots_ms;       /* ... K4a Gram-Schmidt dots                              */
    float   gsupd_ms;      /* ... K4b Gram-Schmidt update                            */
    float   fused_ms;      /* ... KF single-pass fused step                          */
    int32_t step_kernel;   /* the step kernel that ran (1, 2 or 3 as in lz_run_opts; 4: the whole solve ran in
                              the persistent cooperative kernel for small matrix-free problems) */
    int32_t gsfused_launches; /* ... K4c fused Gram-Schmidt update + dots              */
    float   gsfused_ms;
    int32_t border_launches;  /* ... the border kernel that completes alpha when KB accumulates it */
    float   border_ms;
    int32_t alpha_in_update;  /* 1: alpha of the next vector was accumulated inside KB (+ border kernel)
                                 instead of a KA pass over the vector; 2: one kernel per step (KBA) - the
                                 alpha reduction chases KB's output through L2 (24*M B of HBM per step) */
    int32_t overlap;          /* 1: sparse row shards - the interior rows of the next apply ran while the ghost
                                 entries and the beta sum travelled on a second stream  */
    int32_t graph;            /* launch-bound solves (<= 4 M unknowns, one GPU): 1 = this call captured the whole
                                 solve into a CUDA graph and launched it, 2 = it replayed a cached graph, 0 = plain
                                 launches (first call of a solve, large problems, profile mode, LZ_GRAPH=0) */
} lz_run_info;

/* Runs n steps from v0_dev (M doubles).  Outputs: alpha_host[n], beta_host[n-1]
 * (reference numbering, Lanczos.py:112: beta[k] couples rows k and k+1);
 * V_dev (nullable unless reorth != NONE): n rows of ldv >= M doubles, row-major like
 * the reference's in-loop layout (Lanczos.py:104).  Row j holds the j-th Lanczos
 * vector up to the factor row_scale_host[j] (nullable): q_j = row_scale[j] * V[j,:].
 * Rows that went through a Gram-Schmidt sweep are stored normalised (scale 1); the
 * others are stored un-normalised to keep the step at two passes over HBM.
 * Use lz_basis_normalize to fold the factors in.  Synchronises. */
int lz_lanczos_run(lz_ctx* ctx, lz_op* op, const double* v0_dev, int32_t n,
                   const lz_run_opts* opts, double* alpha_host, double* beta_host,
                   double* V_dev, int64_t ldv, double* row_scale_host, lz_run_info* info);

/* V[j,:] *= row_scale[j] for every j (skips factors equal to 1).  Enqueues only. */
int lz_basis_normalize(lz_ctx* ctx, double* V_dev, int64_t ldv, int32_t n, int64_t M,
                       const double* row_scale_host);

/* sweep forms of the reference's reorthogonalize */
#define LZ_SWEEP_CPU  0   /* V[j] = 2 V[j] - sum_i (V[j].V[i]) V[i], i over all rows incl. j
                             (Lanczos.py:247-249 with use_cuda=False; IrrLanczos.py:453-460 both modes) */
#define LZ_SWEEP_GPU  1   /* V[j] = V[j] - sum_{i != j} (V[j].V[i]) V[i]
                             (Lanczos.py:236-238, Regular with use_cuda=True)                        */

/* One Gram-Schmidt sweep of row j of V against all rows, in place: the staticmethod
 * Lanczos.reorthogonalize(V, j, use_cuda) (Lanczos.py:233-251; IrrLanczos.py:448-466) in the
 * given form.  V is (n x ldv) row-major on the device.  Synchronises. */
int lz_reorthogonalize(lz_ctx* ctx, double* V_dev, int64_t ldv, int32_t n, int64_t M, int32_t j,
                       int32_t form);

/* Ritz vectors: Y[c,:] = sum_j S[j,c] * row_scale[j] * V[j,:], c < k.  The lift loop of
 * get_H_eigs (Lanczos.py:154-156).  S_host is n x k column-major (column c = c-th
 * eigenvector of H_eff), row_scale_host nullable (all ones), Y_dev is k rows of ldy.
 * Synchronises. */
int lz_ritz_vectors(lz_ctx* ctx, const double* V_dev, int64_t ldv, int32_t n, int64_t M,
                    const double* row_scale_host, const double* S_host, int32_t k,
                    double* Y_dev, int64_t ldy);

/* ---- row-sharded (multi-GPU) solves ---------------------------------------------
 * The reference has no distributed path; this is the multi-GPU form of the same loop
 * (BASELINE.json north_star): vectors and the Krylov basis are split into contiguous row
 * blocks, one per GPU (z-slabs of a structured grid).  A "team" is the set of `world` shards;
 * `nlocal` of them are driven by the calling process (1 under torchrun, one process per GPU;
 * all of them when a single process drives several shards, e.g. in tests).
 *
 * Each rank owns one exchange buffer (lz_comm_alloc) that every other rank maps
 * (lz_comm_open of its cudaIpc handle, or the same pointer inside one process).  The fused
 * kernels store halo planes and partial sums straight into the peers' buffers over NVLink;
 * sums are taken in rank order, so alpha/beta are bit-identical on every rank.
 */
int lz_comm_bytes(int world, int32_t max_steps, int64_t plane_elems, int64_t nghost, int64_t* bytes);
int lz_comm_alloc(lz_ctx* ctx, int64_t bytes, void** dev_ptr, unsigned char* ipc_handle64 /*nullable*/);
int lz_comm_open(lz_ctx* ctx, const unsigned char* ipc_handle64, void** dev_ptr);
int lz_comm_close(lz_ctx* ctx, void* dev_ptr);
int lz_comm_free(lz_ctx* ctx, void* dev_ptr);

/* local_ranks[nlocal], ctxs[nlocal]; global_rows = M of the whole operator; max_steps = largest
 * n a run will use; plane_elems = nx*ny of a structured grid (0: none); nghost = largest ghost
 * list of any rank of a sparse operator (0: none). */
int lz_team_create(int world, int nlocal, const int* local_ranks, lz_ctx* const* ctxs,
                   int64_t global_rows, int32_t max_steps, int64_t plane_elems, int64_t nghost,
                   lz_team** out);
/* comm_ptrs[world]: exchange buffers of all ranks as mapped in this process; lower/upper = ranks
 * owning the slab below/above this shard (-1: domain boundary, Dirichlet). */
int lz_team_attach(lz_team* team, int local_index, void* const* comm_ptrs, int lower_rank, int upper_rank);
/* Ghost-index exchange of a sparse row shard: send_idx_host[nsend] = local rows whose x entries
 * other ranks need, grouped by destination rank (seg_start[world+1]); the segment for rank q
 * lands at offset dst_off[q] of q's gather buffer (= position in q's ghost list). */
int lz_team_set_ghosts(lz_team* team, int local_index, int32_t nsend, const int32_t* send_idx_host,
                       const int32_t* seg_start, const int64_t* dst_off);
/* As lz_lanczos_run, with one operator / start vector / basis buffer per local shard (the local
 * slab: an lz_op_stencil_create with the local extents).  alpha/beta are the global values. */
int lz_team_lanczos_run(lz_team* team, lz_op* const* ops, const double* const* v0_dev, int32_t n,
                        const lz_run_opts* opts, double* alpha_host, double* beta_host,
                        double* const* V_dev, const int64_t* ldv, double* row_scale_host,
                        lz_run_info* info);
/* y = H x over the shards of a team (x_dev / y_dev: one local block per local shard) plus the two
 * global sums of the residual diagnostics (print_good_eigs, Lanczos.py:171-176):
 * dots_host[0] = x . H x, dots_host[1] = H x . H x, identical on every rank.  Synchronises. */
int lz_team_apply_dots(lz_team* team, lz_op* const* ops, const double* const* x_dev,
                       double* const* y_dev, double* dots_host);
/* A team whose run failed part-way (LZ_ERR_CUDA / LZ_ERR_PEER) refuses further runs: destroy it and
 * create a new one on every rank. */
int lz_team_destroy(lz_team* team);

/* Deterministic device reductions used by the diagnostics (test_is_normalized,
 * print_good_eigs: Lanczos.py:166-185, 288-304): result_host[0] = x.y.  Synchronises. */
int lz_dot(lz_ctx* ctx, const double* x_dev, const double* y_dev, int64_t M, double* result_host);

#ifdef __cplusplus
}
#endif
#endif /* LANCZOS_B200_H */
