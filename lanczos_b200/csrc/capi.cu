// extern "C" surface of liblanczos_b200 (see include/lanczos_b200.h): contexts, operators,
// and the small entry points that are thin wrappers over one kernel.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "internal.h"

namespace lz {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }

// Off unless LZ_PDL=1.  Measured on B200 at 512^3 it is worth 0.7 % of a step (0.8472 -> 0.8415 ms), and it is
// not safe with this library's loads: a dependent grid's CTAs - and with them the L1 invalidation that normally
// separates two kernels - may start while the primary is still writing, and data the primary produced is then
// read through the non-coherent path (ld.global.nc / __ldg), which griddepcontrol.wait does not cover.  Seen as
// wrong alpha from step 2-3 on in row-sharded two-pass runs (tools/pdl_race_check.py with LZ_PDL=1), never with the attribute off.
}  // namespace lz
void lz_ctx_drop_graphs(lz_ctx* c);     // lanczos.cu
namespace lz {
bool pdl_enabled() {
    static const bool on = []() {
        const char* e = getenv("LZ_PDL");
        return e && e[0] == '1';
    }();
    return on;
}

int launch_stencil_apply_dot(lz_op* op, const double* x, const double* scale_dev, double* y,
                             double* partials, int* nparts, const int* flag_dev, const FinTail* fin);
int launch_spmv_dot(lz_op* op, const double* x, const double* scale_dev, double* y, double* partials,
                    int* nparts, const int* flag_dev, const FinTail* fin);
int launch_stencil27_apply_dot(lz_op* op, const double* x, const double* scale_dev, double* y,
                               double* partials, int* nparts, const int* flag_dev, const FinTail* fin);
int launch_stencil_update_norm(lz_op* op, const double* x, const double* scale_dev, const StencilUpdate* upd,
                               double* out, double* partials, int* nparts, const FinTail* fin);
int launch_stencil27_update_norm(lz_op* op, const double* x, const double* scale_dev, const StencilUpdate* upd,
                                 double* out, double* partials, int* nparts, const FinTail* fin);
int build_csr(lz_op* op, int64_t M, int64_t nnz, const int32_t* indptr, const int32_t* indices,
              const double* data);
int build_sell(lz_op* op, int64_t M, int64_t nnz, const int32_t* indptr, const int32_t* indices,
               const double* data, int sigma);
int build_from_device(lz_op* op, int64_t M, int64_t ncols, int64_t nnz, const int32_t* indptr,
                      const int32_t* indices, const double* data, int fmt, int sigma);

int launch_apply_dot(lz_op* op, const double* x, const double* scale_dev, double* y, double* partials,
                     int* nparts, int* launches, const int* flag_dev, const FinTail* fin) {
    if (launches) *launches = 1;
    if (op->kind == LZ_OP_STENCIL && op->st.points == 27)
        return launch_stencil27_apply_dot(op, x, scale_dev, y, partials, nparts, flag_dev, fin);
    if (op->kind == LZ_OP_STENCIL) return launch_stencil_apply_dot(op, x, scale_dev, y, partials, nparts, flag_dev, fin);
    return launch_spmv_dot(op, x, scale_dev, y, partials, nparts, flag_dev, fin);
}

// Matrix-free operators can re-evaluate H x inside the update instead of storing it (KA + KB).
bool recompute_step_supported(const lz_op* op) { return op->kind == LZ_OP_STENCIL; }
// ... and it pays where applying H is cheap next to the 16*M bytes it saves: the 7-point family.  The
// 27-point kernel is latency/issue-bound (0.8 ms per apply at 512^3): two applies measured 2.07 ms per
// step against 1.62 ms for K1b + K3, so `auto` keeps the two-pass step there.
bool recompute_step_preferred(const lz_op* op) { return op->kind == LZ_OP_STENCIL && op->st.points == 7; }

int launch_apply_update_norm(lz_op* op, const double* x, const double* scale_dev, const StencilUpdate* upd,
                             double* out, double* partials, int* nparts, int* launches, const FinTail* fin) {
    if (launches) *launches = 1;
    if (op->kind != LZ_OP_STENCIL) {
        set_error("the recompute step needs a matrix-free operator");
        return LZ_ERR_UNSUPPORTED;
    }
    if (op->st.points == 27) return launch_stencil27_update_norm(op, x, scale_dev, upd, out, partials, nparts, fin);
    return launch_stencil_update_norm(op, x, scale_dev, upd, out, partials, nparts, fin);
}

// sum of np partials -> out[0] (one CTA, fixed order)
__global__ void __launch_bounds__(kThreads)
sum_partials_kernel(const double* __restrict__ p, int np, double* __restrict__ out) {
    __shared__ double red[kWarps];
    double a = 0.0;
    for (int i = threadIdx.x; i < np; i += kThreads) a += p[i];
    const double t = block_sum(a, red);
    if (threadIdx.x == 0) out[0] = t;
}

// coefficients of the stand-alone Gram-Schmidt sweep (Lanczos.reorthogonalize on arbitrary V):
// coef[r] = ip_r (r != j), coef[j] = 0, cself = 2 - ip_j (LZ_SWEEP_CPU) or 1 (LZ_SWEEP_GPU)
__global__ void __launch_bounds__(kThreads)
sweep_coef_kernel(const double* __restrict__ part, int ncg, int n, int j, double* __restrict__ coef,
                  double* __restrict__ cself, int form) {
    for (int r = threadIdx.x; r < n; r += kThreads) {
        const double* p = part + (int64_t)r * ncg;
        double a = 0.0;
        for (int g = 0; g < ncg; ++g) a += p[g];
        if (r == j) { cself[0] = (form == LZ_SWEEP_GPU) ? 1.0 : 2.0 - a; coef[r] = 0.0; }
        else coef[r] = a;
    }
}
__global__ void set_one_kernel(double* p) { if (threadIdx.x == 0 && blockIdx.x == 0) p[0] = 1.0; }

}  // namespace lz

using namespace lz;

extern "C" {

int lz_abi_version(void) { return LZ_ABI_VERSION; }
const char* lz_last_error(void) { return get_error(); }

int lz_device_count(int* count) {
    LZ_REQUIRE(count, "lz_device_count: null argument");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { n = 0; (void)cudaGetLastError(); }
    *count = n;
    return LZ_OK;
}

int lz_ctx_create(int device, void* cuda_stream, lz_ctx** out) {
    LZ_REQUIRE(out, "lz_ctx_create: null output");
    *out = nullptr;
    int ndev = 0;
    LZ_CUDA(cudaGetDeviceCount(&ndev));
    LZ_REQUIRE(device >= 0 && device < ndev, "lz_ctx_create: device %d out of range (0..%d)", device, ndev - 1);
    LZ_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    LZ_CUDA(cudaGetDeviceProperties(&prop, device));
    LZ_REQUIRE(prop.major >= 10, "lz_ctx_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only",
               device, prop.major, prop.minor);
    lz_ctx* c = new lz_ctx();
    c->device = device;
    c->stream = (cudaStream_t)cuda_stream;
    c->sms = prop.multiProcessorCount;
    cudaError_t e = cudaMalloc((void**)&c->partials, (size_t)2 * kMaxPartials * 8);
    if (e == cudaSuccess) e = cudaMalloc((void**)&c->scratch, 64 * 8);
    if (e == cudaSuccess) e = cudaMalloc((void**)&c->kba_done, 4096 * sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(c->kba_done, 0, 4096 * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc((void**)&c->tickets, 64 * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(c->tickets, 0, 64 * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev_begin);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev_end);
    if (e != cudaSuccess) {
        set_error("lz_ctx_create: %s", cudaGetErrorString(e));
        lz_ctx_destroy(c);
        return LZ_ERR_CUDA;
    }
    *out = c;
    return LZ_OK;
}

int lz_ctx_destroy(lz_ctx* c) {
    if (!c) return LZ_OK;
    cudaSetDevice(c->device);
    lz_ctx_drop_graphs(c);
    if (c->gstream) cudaStreamDestroy(c->gstream);
    if (c->stream2) cudaStreamDestroy(c->stream2);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->partials) cudaFree(c->partials);
    if (c->scratch) cudaFree(c->scratch);
    if (c->tickets) cudaFree(c->tickets);
    if (c->kba_done) cudaFree(c->kba_done);
    if (c->arena) cudaFree(c->arena);
    for (cudaEvent_t e : c->event_pool) cudaEventDestroy(e);
    if (c->ev_begin) cudaEventDestroy(c->ev_begin);
    if (c->ev_end) cudaEventDestroy(c->ev_end);
    delete c;
    return LZ_OK;
}

int lz_ctx_sync(lz_ctx* c) {
    LZ_REQUIRE(c, "lz_ctx_sync: null context");
    LZ_CUDA(cudaSetDevice(c->device));
    LZ_CUDA(cudaStreamSynchronize(c->stream));
    return LZ_OK;
}

int lz_op_stencil_create(lz_ctx* ctx, int dim, const int64_t* shape, int bc, double center,
                         const double* offdiag, const double* diag_dev, lz_op** out) {
    LZ_REQUIRE(ctx && shape && offdiag && out, "lz_op_stencil_create: null argument");
    LZ_REQUIRE(dim >= 1 && dim <= 3, "lz_op_stencil_create: dim must be 1, 2 or 3 (got %d)", dim);
    LZ_REQUIRE(bc == LZ_BC_PERIODIC || bc == LZ_BC_DIRICHLET, "lz_op_stencil_create: unknown boundary condition %d", bc);
    lz_op* op = new lz_op();
    op->ctx = ctx;
    op->kind = LZ_OP_STENCIL;
    lz_stencil& st = op->st;
    st.dim = dim;
    st.bc = bc;
    st.center = center;
    st.diag = diag_dev;
    int64_t ext[3] = {1, 1, 1};
    double off[3] = {0.0, 0.0, 0.0};
    for (int a = 0; a < dim; ++a) {
        if (shape[a] < 1 || shape[a] > 0x7fffffff) {
            set_error("lz_op_stencil_create: extent %lld of axis %d out of range", (long long)shape[a], a);
            delete op;
            return LZ_ERR_INVALID;
        }
        ext[a] = shape[a];
        off[a] = offdiag[a];
    }
    st.nx = ext[0]; st.ny = ext[1]; st.nz = ext[2];
    st.offx = off[0]; st.offy = off[1]; st.offz = off[2];
    op->M = ext[0] * ext[1] * ext[2];
    *out = op;
    return LZ_OK;
}

static int csr_create_impl(lz_ctx* ctx, int64_t M, int64_t ncols, int64_t nnz, const int32_t* indptr,
                           const int32_t* indices, const double* data, int fmt, int sigma, lz_op** out);

int lz_op_stencil27_create(lz_ctx* ctx, const int64_t* shape, int bc, const double* weights,
                           const double* diag_dev, lz_op** out) {
    LZ_REQUIRE(ctx && shape && weights && out, "lz_op_stencil27_create: null argument");
    const double off[3] = {weights[1], weights[1], weights[1]};
    LZ_CHECK(lz_op_stencil_create(ctx, 3, shape, bc, weights[0], off, diag_dev, out));
    lz_stencil& st = (*out)->st;
    st.points = 27;
    for (int k = 0; k < 4; ++k) st.w27[k] = weights[k];
    return LZ_OK;
}

int lz_op_csr_create(lz_ctx* ctx, int64_t M, int64_t nnz, const int32_t* indptr, const int32_t* indices,
                     const double* data, int fmt, int sigma, lz_op** out) {
    return csr_create_impl(ctx, M, M, nnz, indptr, indices, data, fmt, sigma, out);
}

int lz_op_csr_shard_create(lz_ctx* ctx, int64_t M_local, int64_t ncols, int64_t nnz, const int32_t* indptr,
                           const int32_t* indices, const double* data, int fmt, int sigma, lz_op** out) {
    LZ_REQUIRE(ncols >= M_local, "lz_op_csr_shard_create: ncols < M_local");
    return csr_create_impl(ctx, M_local, ncols, nnz, indptr, indices, data, fmt, sigma, out);
}

int lz_op_csr_create_dev(lz_ctx* ctx, int64_t M, int64_t ncols, int64_t nnz, const int32_t* indptr_dev,
                         const int32_t* indices_dev, const double* data_dev, int fmt, int sigma, lz_op** out) {
    LZ_REQUIRE(ctx && indptr_dev && out && (nnz == 0 || (indices_dev && data_dev)), "lz_op_csr_create_dev: null argument");
    LZ_REQUIRE(M >= 1 && M <= 0x7fffffff && ncols >= M && ncols <= 0x7fffffff, "lz_op_csr_create_dev: M / ncols out of range");
    LZ_REQUIRE(nnz >= 0 && nnz <= 0x7fffffff, "lz_op_csr_create_dev: nnz must fit int32 indptr");
    LZ_REQUIRE(fmt == LZ_FMT_CSR || fmt == LZ_FMT_SELL || fmt == LZ_FMT_SELL_VALUES, "lz_op_csr_create_dev: unknown format %d", fmt);
    LZ_CUDA(cudaSetDevice(ctx->device));
    lz_op* op = new lz_op();
    op->ctx = ctx;
    int st = build_from_device(op, M, ncols, nnz, indptr_dev, indices_dev, data_dev, fmt, sigma);
    if (st != LZ_OK) { lz_op_destroy(op); return st; }
    op->ncols = ncols;
    st = sell_classify_spans(op);
    if (st == LZ_OK && fmt == LZ_FMT_SELL) st = sell_detect_uniform(op);
    if (st == LZ_OK) st = sellw_build(op);
    if (st != LZ_OK) { lz_op_destroy(op); return st; }
    *out = op;
    return LZ_OK;
}

static int csr_create_impl(lz_ctx* ctx, int64_t M, int64_t ncols, int64_t nnz, const int32_t* indptr,
                           const int32_t* indices, const double* data, int fmt, int sigma, lz_op** out) {
    LZ_REQUIRE(ctx && indptr && out && (nnz == 0 || (indices && data)), "lz_op_csr_create: null argument");
    LZ_REQUIRE(M >= 1 && M <= 0x7fffffff && ncols <= 0x7fffffff, "lz_op_csr_create: M out of range");
    LZ_REQUIRE(nnz >= 0 && nnz <= 0x7fffffff, "lz_op_csr_create: nnz must fit int32 indptr");
    LZ_REQUIRE(fmt == LZ_FMT_CSR || fmt == LZ_FMT_SELL || fmt == LZ_FMT_SELL_VALUES, "lz_op_csr_create: unknown format %d", fmt);
    LZ_REQUIRE(indptr[0] == 0 && indptr[M] == nnz, "lz_op_csr_create: indptr does not span [0, nnz]");
    for (int64_t i = 0; i < M; ++i)
        LZ_REQUIRE(indptr[i + 1] >= indptr[i], "lz_op_csr_create: indptr not monotone at row %lld", (long long)i);
    for (int64_t k = 0; k < nnz; ++k)
        LZ_REQUIRE(indices[k] >= 0 && indices[k] < ncols, "lz_op_csr_create: column index %d out of range at entry %lld",
                   indices[k], (long long)k);
    LZ_CUDA(cudaSetDevice(ctx->device));
    lz_op* op = new lz_op();
    op->ctx = ctx;
    int st = (fmt == LZ_FMT_CSR) ? build_csr(op, M, nnz, indptr, indices, data)
                                 : build_sell(op, M, nnz, indptr, indices, data, sigma);
    if (st != LZ_OK) { lz_op_destroy(op); return st; }
    op->ncols = ncols;
    st = sell_classify_spans(op);              // row shards: interior / boundary spans for the overlapped apply
    if (st == LZ_OK && fmt == LZ_FMT_SELL) st = sell_detect_uniform(op);   // unweighted graph Laplacians: value-free kernel
    if (st == LZ_OK) st = sellw_build(op);     // locality-ordered matrices: x staged in shared memory, 16-bit indices
    if (st != LZ_OK) { lz_op_destroy(op); return st; }
    *out = op;
    return LZ_OK;
}

int lz_op_rows(const lz_op* op, int64_t* M) {
    LZ_REQUIRE(op && M, "lz_op_rows: null argument");
    *M = op->M;
    return LZ_OK;
}

int lz_op_nnz(const lz_op* op, int64_t* nnz_true, int64_t* nnz_stored) {
    LZ_REQUIRE(op, "lz_op_nnz: null argument");
    int64_t t = 0, s = 0;
    if (op->kind == LZ_OP_CSR) { t = s = op->csr.nnz; }
    else if (op->kind == LZ_OP_SELL) { t = op->sell.nnz_true; s = op->sell.nnz_stored; }
    else {
        const lz_stencil& g = op->st;
        int per = 1 + 2 * ((g.offx != 0.0) + (g.offy != 0.0) + (g.offz != 0.0));
        if (g.points == 27) per = 27;
        t = (int64_t)per * op->M;   // upper bound (boundaries/duplicates merge), matrix-free: nothing stored
        s = 0;
    }
    if (nnz_true) *nnz_true = t;
    if (nnz_stored) *nnz_stored = s;
    return LZ_OK;
}

int lz_op_value_free(const lz_op* op, int32_t* value_free) {
    LZ_REQUIRE(op && value_free, "lz_op_value_free: null argument");
    *value_free = (op->kind == LZ_OP_SELL && op->sell.uniform) ? 1 : 0;
    return LZ_OK;
}

int lz_op_windowed(const lz_op* op, int32_t* granules) {
    LZ_REQUIRE(op && granules, "lz_op_windowed: null argument");
    *granules = (op->kind == LZ_OP_SELL && op->sell.windowed) ? op->sell.win_maxg : 0;
    return LZ_OK;
}

int lz_op_apply(lz_op* op, const double* x_dev, double* y_dev) {
    LZ_REQUIRE(op && x_dev && y_dev, "lz_op_apply: null argument");
    LZ_REQUIRE(x_dev != y_dev, "lz_op_apply: in-place apply is not supported");
    LZ_REQUIRE(op->kind == LZ_OP_STENCIL || op->ncols <= op->M, "lz_op_apply: row shards are applied by the team run");
    LZ_CUDA(cudaSetDevice(op->ctx->device));
    int np = 0, l = 0;
    return launch_apply_dot(op, x_dev, nullptr, y_dev, op->ctx->partials + kMaxPartials, &np, &l);
}

int lz_op_destroy(lz_op* op) {
    if (!op) return LZ_OK;
    if (op->ctx) cudaSetDevice(op->ctx->device);
    if (op->csr.indptr) cudaFree(op->csr.indptr);
    if (op->csr.indices) cudaFree(op->csr.indices);
    if (op->csr.data) cudaFree(op->csr.data);
    if (op->sell.chunk_off) cudaFree(op->sell.chunk_off);
    if (op->sell.col) cudaFree(op->sell.col);
    if (op->sell.val) cudaFree(op->sell.val);
    if (op->sell.row_of) cudaFree(op->sell.row_of);
    if (op->sell.deff) cudaFree(op->sell.deff);
    if (op->sell.spans_int) cudaFree(op->sell.spans_int);
    if (op->sell.spans_bnd) cudaFree(op->sell.spans_bnd);
    if (op->sell.win_gran_off) cudaFree(op->sell.win_gran_off);
    if (op->sell.win_gran) cudaFree(op->sell.win_gran);
    if (op->sell.win_lcol) cudaFree(op->sell.win_lcol);
    if (op->sell.win_lrow) cudaFree(op->sell.win_lrow);
    if (op->sell.win_off8) cudaFree(op->sell.win_off8);
    if (op->sell.win_deff) cudaFree(op->sell.win_deff);
    delete op;
    return LZ_OK;
}

// Export as canonical CSR: rows ascending, columns ascending inside a row, duplicates summed.
int lz_op_export_csr(lz_op* op, int64_t* nnz_out, int32_t* indptr_host, int32_t* indices_host,
                     double* data_host) {
    LZ_REQUIRE(op && nnz_out, "lz_op_export_csr: null argument");
    lz_ctx* ctx = op->ctx;
    LZ_CUDA(cudaSetDevice(ctx->device));
    const int64_t M = op->M;
    // gather (row, col, val) triplets row by row in a host-side canonical form
    auto canonical_emit = [&](auto&& row_entries) -> int {
        // row_entries(i, vec) fills the raw entries of row i
        int64_t total = 0;
        std::vector<std::pair<int32_t, double>> e, merged;
        std::vector<int32_t> ip((size_t)M + 1, 0);
        std::vector<int32_t> idx;
        std::vector<double> dat;
        for (int64_t i = 0; i < M; ++i) {
            e.clear();
            row_entries(i, e);
            std::stable_sort(e.begin(), e.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
            merged.clear();
            for (auto& kv : e) {
                if (!merged.empty() && merged.back().first == kv.first) merged.back().second += kv.second;
                else merged.push_back(kv);
            }
            total += (int64_t)merged.size();
            if (indptr_host) {
                for (auto& kv : merged) { idx.push_back(kv.first); dat.push_back(kv.second); }
                ip[(size_t)i + 1] = (int32_t)total;
            }
        }
        *nnz_out = total;
        if (indptr_host) {
            LZ_REQUIRE(indices_host && data_host, "lz_op_export_csr: indices/data output is null");
            memcpy(indptr_host, ip.data(), ((size_t)M + 1) * 4);
            memcpy(indices_host, idx.data(), idx.size() * 4);
            memcpy(data_host, dat.data(), dat.size() * 8);
        }
        return LZ_OK;
    };

    if (op->kind == LZ_OP_STENCIL) {
        const lz_stencil& g = op->st;
        std::vector<double> diag;
        if (g.diag) {
            diag.resize((size_t)M);
            LZ_CUDA(cudaMemcpyAsync(diag.data(), g.diag, (size_t)M * 8, cudaMemcpyDeviceToHost, ctx->stream));
            LZ_CUDA(cudaStreamSynchronize(ctx->stream));
        }
        const int64_t ext[3] = {g.nx, g.ny, g.nz};
        const double off[3] = {g.offx, g.offy, g.offz};
        const int64_t stride[3] = {1, g.nx, g.nx * g.ny};
        return canonical_emit([&](int64_t i, std::vector<std::pair<int32_t, double>>& e) {
            const int64_t c[3] = {i % g.nx, (i / g.nx) % g.ny, i / (g.nx * g.ny)};
            if (g.points == 27) {
                for (int dz = -1; dz <= 1; ++dz)
                    for (int dy = -1; dy <= 1; ++dy)
                        for (int dx = -1; dx <= 1; ++dx) {
                            int64_t k[3] = {c[0] + dx, c[1] + dy, c[2] + dz};
                            bool inside = true;
                            for (int a = 0; a < 3; ++a) {
                                if (k[a] < 0 || k[a] >= ext[a]) {
                                    if (g.bc != LZ_BC_PERIODIC) { inside = false; break; }
                                    k[a] = (k[a] + ext[a]) % ext[a];
                                }
                            }
                            if (!inside) continue;
                            const int nz_off = (dx != 0) + (dy != 0) + (dz != 0);
                            double w = g.w27[nz_off];
                            if (nz_off == 0 && g.diag) w += diag[(size_t)i];
                            e.push_back({(int32_t)(k[0] + k[1] * stride[1] + k[2] * stride[2]), w});
                        }
                return;
            }
            e.push_back({(int32_t)i, g.center + (g.diag ? diag[(size_t)i] : 0.0)});
            for (int a = 0; a < 3; ++a) {
                if (off[a] == 0.0) continue;
                for (int d = -1; d <= 1; d += 2) {
                    int64_t k = c[a] + d;
                    if (k < 0 || k >= ext[a]) {
                        if (g.bc != LZ_BC_PERIODIC) continue;
                        k = (k + ext[a]) % ext[a];
                    }
                    e.push_back({(int32_t)(i + (k - c[a]) * stride[a]), off[a]});
                }
            }
        });
    }
    if (op->kind == LZ_OP_CSR) {
        const lz_csr& c = op->csr;
        std::vector<int32_t> ip((size_t)M + 1), idx((size_t)c.nnz);
        std::vector<double> dat((size_t)c.nnz);
        LZ_CUDA(cudaMemcpyAsync(ip.data(), c.indptr, ip.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
        if (c.nnz) {
            LZ_CUDA(cudaMemcpyAsync(idx.data(), c.indices, idx.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
            LZ_CUDA(cudaMemcpyAsync(dat.data(), c.data, dat.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
        }
        LZ_CUDA(cudaStreamSynchronize(ctx->stream));
        return canonical_emit([&](int64_t i, std::vector<std::pair<int32_t, double>>& e) {
            for (int32_t k = ip[(size_t)i]; k < ip[(size_t)i + 1]; ++k) e.push_back({idx[(size_t)k], dat[(size_t)k]});
        });
    }
    // SELL: walk the chunks, drop padding (stored beyond the row length cannot be told from a
    // genuine explicit zero, so explicit zeros of the input are dropped as well)
    const lz_sell& sl = op->sell;
    std::vector<int64_t> off((size_t)sl.nchunks + 1);
    std::vector<int32_t> col((size_t)sl.nnz_stored), row_of((size_t)sl.nchunks * 32);
    std::vector<double> val((size_t)sl.nnz_stored);
    LZ_CUDA(cudaMemcpyAsync(off.data(), sl.chunk_off, off.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    LZ_CUDA(cudaMemcpyAsync(row_of.data(), sl.row_of, row_of.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (sl.nnz_stored) {
        LZ_CUDA(cudaMemcpyAsync(col.data(), sl.col, col.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
        LZ_CUDA(cudaMemcpyAsync(val.data(), sl.val, val.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    LZ_CUDA(cudaStreamSynchronize(ctx->stream));
    std::vector<int64_t> slot_of((size_t)M, -1);
    for (int64_t q = 0; q < (int64_t)row_of.size(); ++q)
        if (row_of[(size_t)q] >= 0) slot_of[(size_t)row_of[(size_t)q]] = q;
    return canonical_emit([&](int64_t i, std::vector<std::pair<int32_t, double>>& e) {
        const int64_t q = slot_of[(size_t)i];
        if (q < 0) return;
        const int64_t c = q / 32, l = q % 32;
        const int width = (int)((off[(size_t)c + 1] - off[(size_t)c]) / 32);
        for (int k = 0; k < width; ++k) {
            const size_t at = (size_t)(off[(size_t)c] + (int64_t)k * 32 + l);
            if (val[at] != 0.0) e.push_back({col[at], val[at]});
        }
    });
}

int lz_reorthogonalize(lz_ctx* ctx, double* V_dev, int64_t ldv, int32_t n, int64_t M, int32_t j, int32_t form) {
    LZ_REQUIRE(ctx && V_dev, "lz_reorthogonalize: null argument");
    LZ_REQUIRE(form == LZ_SWEEP_CPU || form == LZ_SWEEP_GPU, "lz_reorthogonalize: unknown sweep form %d", form);
    LZ_REQUIRE(n >= 1 && j >= 0 && j < n && ldv >= M && M >= 1, "lz_reorthogonalize: bad shape (n=%d, j=%d)", n, j);
    LZ_CUDA(cudaSetDevice(ctx->device));
    double* part = nullptr;
    double* coef = nullptr;
    LZ_CUDA(cudaMalloc((void**)&part, (size_t)n * kMaxPartials * 8));
    cudaError_t e = cudaMalloc((void**)&coef, ((size_t)n + 2) * 8);
    if (e != cudaSuccess) { cudaFree(part); set_error("lz_reorthogonalize: %s", cudaGetErrorString(e)); return LZ_ERR_NOMEM; }
    double* cself = coef + n;
    double* one = coef + n + 1;
    double* target = V_dev + (int64_t)j * ldv;
    int ncg = 0;
    int st = launch_cgs_dots(ctx, V_dev, ldv, n, target, M, part, &ncg, nullptr);
    if (st == LZ_OK) {
        sweep_coef_kernel<<<1, kThreads, 0, ctx->stream>>>(part, ncg, n, j, coef, cself, form);
        set_one_kernel<<<1, 32, 0, ctx->stream>>>(one);
        // rows before j, then rows after j (row j itself is the in-place target)
        st = launch_cgs_update(ctx, V_dev, ldv, j, target, coef, cself, target, M, nullptr);
        if (st == LZ_OK && j + 1 < n)
            st = launch_cgs_update(ctx, V_dev + (int64_t)(j + 1) * ldv, ldv, n - j - 1, target, coef + j + 1,
                                   one, target, M, nullptr);
    }
    cudaStreamSynchronize(ctx->stream);
    cudaFree(part);
    cudaFree(coef);
    if (st == LZ_OK) LZ_CUDA(cudaGetLastError());
    return st;
}

int lz_ritz_vectors(lz_ctx* ctx, const double* V_dev, int64_t ldv, int32_t n, int64_t M,
                    const double* row_scale_host, const double* S_host, int32_t k, double* Y_dev,
                    int64_t ldy) {
    LZ_REQUIRE(ctx && V_dev && S_host && Y_dev, "lz_ritz_vectors: null argument");
    LZ_REQUIRE(n >= 1 && k >= 1 && ldv >= M && ldy >= M, "lz_ritz_vectors: bad shape");
    LZ_CUDA(cudaSetDevice(ctx->device));
    std::vector<double> S((size_t)n * k);
    for (int c = 0; c < k; ++c)
        for (int r = 0; r < n; ++r)
            S[(size_t)c * n + r] = S_host[(size_t)c * n + r] * (row_scale_host ? row_scale_host[r] : 1.0);
    double* S_dev = nullptr;
    LZ_CUDA(cudaMalloc((void**)&S_dev, S.size() * 8));
    cudaError_t e = cudaMemcpyAsync(S_dev, S.data(), S.size() * 8, cudaMemcpyHostToDevice, ctx->stream);
    int st = LZ_OK;
    if (e != cudaSuccess) { set_error("lz_ritz_vectors: %s", cudaGetErrorString(e)); st = LZ_ERR_CUDA; }
    if (st == LZ_OK) st = launch_ritz_lift(ctx, V_dev, ldv, n, M, S_dev, k, Y_dev, ldy);
    cudaStreamSynchronize(ctx->stream);   // S (pageable) and S_dev are released here
    cudaFree(S_dev);
    return st;
}

int lz_dot(lz_ctx* ctx, const double* x_dev, const double* y_dev, int64_t M, double* result_host) {
    LZ_REQUIRE(ctx && x_dev && y_dev && result_host && M >= 1, "lz_dot: bad argument");
    LZ_CUDA(cudaSetDevice(ctx->device));
    int np = 0;
    LZ_CHECK(launch_dot(ctx, x_dev, y_dev, M, ctx->partials, &np));
    sum_partials_kernel<<<1, kThreads, 0, ctx->stream>>>(ctx->partials, np, ctx->scratch);
    LZ_CUDA(cudaMemcpyAsync(result_host, ctx->scratch, 8, cudaMemcpyDeviceToHost, ctx->stream));
    LZ_CUDA(cudaStreamSynchronize(ctx->stream));
    return LZ_OK;
}

}  // extern "C"
