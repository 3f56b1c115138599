// K2: sparse operator apply with the Lanczos alpha dot fused in, for irregular meshes.
//
//   y = s * (H x),   partial[cta] = sum_i y_i * (s * x_i)
//
// replaces `r = H*V[j]` + `np.dot(V[j], r)` of IrrLanczos.execute_LanczosOld
// (IrrLanczos.py:234,237; SciPy csr_matvec / csc_matvec on the CPU path, cuSPARSE behind
// cupyx on the reference's GPU path).  Two device layouts:
//   * CSR (as given by scipy): a sub-warp of T lanes per row, T chosen from the mean row
//     length, so that a warp reads a contiguous run of indices/data (warp-per-row-group);
//   * SELL-C-sigma, C = 32: rows sorted by length inside windows of sigma rows, chunks of 32
//     rows stored column-major, one warp per chunk: every indices/data load is a fully
//     coalesced 128 B / 256 B transaction and the x gathers of a warp hit nearby sectors
//     when the vertex numbering has locality.
// Both are HBM/L2-gather bound: 12 B per stored entry + 16 B per row.
#include <stdlib.h>
#include <algorithm>
#include <numeric>
#include <vector>
#include "internal.h"

namespace lz {

template <int T>
__global__ void __launch_bounds__(kThreads)
spmv_csr_dot_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                    const double* __restrict__ data, const double* __restrict__ x,
                    const double* __restrict__ scale, double* __restrict__ y, int64_t M,
                    double* __restrict__ partials, const double* __restrict__ xg, const FinTail fin) {
    pdl_prologue();
    __shared__ double red[kWarps];
    constexpr int RPW = 32 / T;                       // rows per warp
    // sharded operator: columns >= M are ghost entries that live in the exchange buffer
    const double* const xgs = xg ? xg - M : x;
    const double s = scale ? __ldg(scale) : 1.0;
    const int lane = threadIdx.x & 31;
    const int sub = lane % T;
    const int64_t warp_global = ((int64_t)blockIdx.x * kThreads + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * kThreads) >> 5;
    double acc = 0.0;
    for (int64_t base = warp_global * RPW; base < M; base += nwarps * RPW) {
        const int64_t row = base + lane / T;
        const bool valid = row < M;
        double sum = 0.0;
        if (valid) {
            const int32_t k0 = __ldg(indptr + row), k1 = __ldg(indptr + row + 1);
            for (int32_t k = k0 + sub; k < k1; k += T) {
                const int32_t c = __ldg(indices + k);
                sum = fma(__ldg(data + k), __ldg((c < M ? x : xgs) + c), sum);
            }
        }
#pragma unroll
        for (int o = T / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (valid && sub == 0) {
            const double yi = s * sum;
            y[row] = yi;
            acc = fma(yi, s * __ldg(x + row), acc);
        }
    }
    const double tot = block_sum(acc, red);
    if (threadIdx.x == 0 && partials) partials[blockIdx.x] = tot;
    fin_tail(fin, partials, red);
}

__device__ __forceinline__ int32_t ld_stream_i32(const int32_t* p) {
    int32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// One warp per chunk of 32 rows.  The CTA takes "spans" of `span` consecutive chunks (whole sorting
// windows), dealt round-robin over the grid: all resident CTAs work on neighbouring spans (the x
// entries they gather stay in L2), a CTA stays inside one span for span/8 iterations (its x
// footprint - the window's rows and their neighbours - is reused out of L1), and the static deal
// keeps the partial sums deterministic.  Column indices and values are streamed past L1, which is
// left to the gathers.  Measured on the 50M-vertex config-4 graph (tools/tune_sell.py, B200):
// 3.27 ms with a plain grid-stride over chunks and sigma = 1024, 1.93 ms with spans, sigma = 2048;
// L2 evict-first hints on the streams made it slower and are not used.
// UNI: every off-diagonal entry of the operator has the same value `uni_a` (an unweighted graph Laplacian,
// BASELINE configs 2 and 4: L = D - A).  The values are then not read at all - 4 instead of 12 bytes per
// stored entry: y_i = a * sum_slots x[col] + deff_i * x_i, where deff_i collects what the slots with col == i
// (the diagonal entry and the padding, which gathers x_i itself) should have contributed beyond a * x_i.
template <bool UNI>
__global__ void __launch_bounds__(kThreads)
spmv_sell_dot_kernel(const int64_t* __restrict__ chunk_off, const int32_t* __restrict__ col,
                     const double* __restrict__ val, const int32_t* __restrict__ row_of,
                     const double* __restrict__ x, const double* __restrict__ scale,
                     double* __restrict__ y, int64_t nchunks, double* __restrict__ partials,
                     const double* __restrict__ xg, int32_t M, int span, const FinTail fin,
                     const int32_t* __restrict__ span_list, int nlist, const int* __restrict__ flag,
                     const double* __restrict__ deff, double uni_a) {
    pdl_prologue();
    if (flag && *flag == 0) return;
    __shared__ double red[kWarps];
    const double s = scale ? __ldg(scale) : 1.0;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const double* const xgs = xg ? xg - M : x;        // ghost columns (>= M) of a sharded operator
    // all spans, or the listed ones (interior / boundary part of a row shard)
    const int64_t nspans = span_list ? (int64_t)nlist : (nchunks + span - 1) / span;
    double acc = 0.0;
    for (int64_t si = blockIdx.x; si < nspans; si += gridDim.x) {
        const int64_t sp = span_list ? (int64_t)__ldg(span_list + si) : si;
        const int64_t c_end = min(nchunks, (sp + 1) * span);
        for (int64_t c = sp * span + warp; c < c_end; c += kWarps) {
            const int64_t o0 = __ldg(chunk_off + c), o1 = __ldg(chunk_off + c + 1);
            const int width = (int)((o1 - o0) >> 5);
            const int32_t* pc = col + o0 + lane;
            const double* pv = val + o0 + lane;
            double sum = 0.0;
            // blocks of 8 entries with predication (no serial remainder loop): all column indices of
            // a block are in flight together, then all gathers of x - rows of 7..14 entries are one
            // or two blocks deep instead of a chain of dependent load pairs
            constexpr int U = 8;
            for (int k = 0; k < width; k += U) {
                int32_t cc[U];
                double vv[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const bool on = k + u < width;
                    cc[u] = on ? ld_stream_i32(pc + (k + u) * 32) : 0;
                    if (!UNI) vv[u] = on ? ld_stream1(pv + (k + u) * 32) : 0.0;
                }
                double xx[U];
#pragma unroll
                for (int u = 0; u < U; ++u) xx[u] = (k + u < width) ? __ldg((cc[u] < M ? x : xgs) + cc[u]) : 0.0;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (UNI) sum += xx[u];
                    else sum = fma(vv[u], xx[u], sum);
                }
            }
            const int32_t row = __ldg(row_of + c * 32 + lane);
            if (row >= 0) {
                const double xr = __ldg(x + row);
                if (UNI) sum = fma(uni_a, sum, __ldg(deff + row) * xr);
                const double yi = s * sum;
                y[row] = yi;
                acc = fma(yi, s * xr, acc);
            }
        }
    }
    const double tot = block_sum(acc, red);
    if (threadIdx.x == 0 && partials) partials[blockIdx.x] = tot;
    fin_tail(fin, partials, red);
}

// Chunks per CTA work item: one sorting window (a CTA that stays inside a window re-uses its gathers out of L1),
// halved while the operator is too small to give every CTA of the grid one.  (Sizing the items so that every CTA
// gets >= 4 was measured at 8 GPUs x 6.3 M rows: 0.448 instead of 0.436 ms/step - the block scheduler already
// balances the 1.3 windows per CTA, and shorter items lose L1 reuse.)
static int balanced_span(const lz_ctx* ctx, int64_t nchunks, int sigma) {
    int span = std::max(kWarps, sigma / 32);
    const int64_t want = (int64_t)ctx->sms * 16;
    while (span >= 2 * kWarps && span % 2 == 0 && (nchunks + span - 1) / span < want) span /= 2;   // stays a divisor of the window
    return span;
}

int launch_spmv_dot(lz_op* op, const double* x, const double* scale_dev, double* y,
                    double* partials, int* nparts, const int* flag_dev, const FinTail* fin) {
    (void)flag_dev;   // predicated applies exist only on the structured-grid fused path
    lz_ctx* ctx = op->ctx;
    const FinTail ft = fin ? *fin : FinTail{};
    const int64_t cap = std::min<int64_t>((int64_t)ctx->sms * 8, kMaxPartials);
    if (op->kind == LZ_OP_CSR) {
        const lz_csr& c = op->csr;
        const int T = c.lanes_per_row;
        const int64_t rows_per_cta = (int64_t)kWarps * (32 / T);
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((op->M + rows_per_cta - 1) / rows_per_cta, cap));
#define LZ_CSR_LAUNCH(TT)                                                                          \
    LZ_CUDA(launch_k(spmv_csr_dot_kernel<TT>, dim3(grid), dim3(kThreads), 0, ctx->stream, c.indptr,  \
                     c.indices, c.data, x, scale_dev, y, op->M, partials, op->xghost, ft))
        switch (T) {
            case 2: LZ_CSR_LAUNCH(2); break;
            case 4: LZ_CSR_LAUNCH(4); break;
            case 8: LZ_CSR_LAUNCH(8); break;
            case 16: LZ_CSR_LAUNCH(16); break;
            default: LZ_CSR_LAUNCH(32); break;
        }
#undef LZ_CSR_LAUNCH
        LZ_CUDA(cudaGetLastError());
        if (nparts) *nparts = grid;
        return LZ_OK;
    }
    const lz_sell& sl = op->sell;
    if (sl.windowed == 1) {
        int grid = 0;
        LZ_CHECK(launch_spmv_windowed(op, nullptr, 0, x, scale_dev, y, partials, &grid, flag_dev, ft, ctx->stream));
        if (nparts) *nparts = grid;
        return LZ_OK;
    }
    // spans: one sorting window each; shorter (down to one chunk per warp) when the operator is too
    // small to give every resident CTA a few spans
    const int span = balanced_span(ctx, sl.nchunks, sl.sigma);
    const int64_t nspans = (sl.nchunks + span - 1) / span;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(nspans, std::min<int64_t>((int64_t)ctx->sms * 16, kMaxPartials)));
    if (sl.uniform)
        LZ_CUDA(launch_k(spmv_sell_dot_kernel<true>, dim3(grid), dim3(kThreads), 0, ctx->stream, sl.chunk_off, sl.col,
                         sl.val, sl.row_of, x, scale_dev, y, sl.nchunks, partials, op->xghost, (int32_t)op->M, span, ft,
                         (const int32_t*)nullptr, 0, flag_dev, (const double*)sl.deff, sl.uni_a));
    else
        LZ_CUDA(launch_k(spmv_sell_dot_kernel<false>, dim3(grid), dim3(kThreads), 0, ctx->stream, sl.chunk_off, sl.col,
                         sl.val, sl.row_of, x, scale_dev, y, sl.nchunks, partials, op->xghost, (int32_t)op->M, span, ft,
                         (const int32_t*)nullptr, 0, flag_dev, (const double*)nullptr, 0.0));
    if (nparts) *nparts = grid;
    return LZ_OK;
}

static int upload(void** dev, const void* host, size_t bytes, cudaStream_t s);

// ---- value-free form for operators whose off-diagonal entries are all equal ----------------------------
// One warp per chunk: deff[row] = sum over the slots with col == row of (val - a); any other slot whose value
// is not exactly `a` clears the flag.
__global__ void __launch_bounds__(kThreads)
sell_uniform_kernel(const int64_t* __restrict__ chunk_off, const int32_t* __restrict__ col,
                    const double* __restrict__ val, const int32_t* __restrict__ row_of, int64_t nchunks, double a,
                    double* __restrict__ deff, int* __restrict__ mismatch) {
    const int64_t c = ((int64_t)blockIdx.x * kThreads + threadIdx.x) >> 5;
    if (c >= nchunks) return;
    const int lane = threadIdx.x & 31;
    const int32_t row = row_of[c * 32 + lane];
    const int64_t o0 = chunk_off[c];
    const int width = (int)((chunk_off[c + 1] - o0) >> 5);
    double d = 0.0;
    int bad = 0;
    for (int k = 0; k < width; ++k) {
        const int64_t at = o0 + (int64_t)k * 32 + lane;
        const double v = val[at];
        if (row >= 0 && col[at] == row) d += v - a;
        else if (row >= 0 && v != a) bad = 1;
    }
    if (row >= 0) deff[row] = d;
    if (bad) *mismatch = 1;
}

// Decide whether the operator qualifies (LZ_SELL_UNIFORM=0 turns the form off) and build deff.
int sell_detect_uniform(lz_op* op) {
    static const bool off = []() { const char* e = getenv("LZ_SELL_UNIFORM"); return e && e[0] == '0'; }();
    lz_ctx* ctx = op->ctx;
    lz_sell& sl = op->sell;
    if (off || op->kind != LZ_OP_SELL || sl.nchunks == 0 || sl.nnz_stored == 0) return LZ_OK;
    // candidate value: the first off-diagonal entry among the first chunks
    const int64_t probe_chunks = std::min<int64_t>(sl.nchunks, 4);
    std::vector<int64_t> off4((size_t)probe_chunks + 1);
    LZ_CUDA(cudaMemcpyAsync(off4.data(), sl.chunk_off, off4.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    LZ_CUDA(cudaStreamSynchronize(ctx->stream));
    const int64_t cnt = off4[(size_t)probe_chunks];
    if (cnt <= 0) return LZ_OK;
    std::vector<int32_t> hc((size_t)cnt), hr((size_t)probe_chunks * 32);
    std::vector<double> hv((size_t)cnt);
    LZ_CUDA(cudaMemcpyAsync(hc.data(), sl.col, (size_t)cnt * 4, cudaMemcpyDeviceToHost, ctx->stream));
    LZ_CUDA(cudaMemcpyAsync(hv.data(), sl.val, (size_t)cnt * 8, cudaMemcpyDeviceToHost, ctx->stream));
    LZ_CUDA(cudaMemcpyAsync(hr.data(), sl.row_of, hr.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    LZ_CUDA(cudaStreamSynchronize(ctx->stream));
    bool found = false;
    double a = 0.0;
    for (int64_t c = 0; c < probe_chunks && !found; ++c)
        for (int64_t at = off4[(size_t)c]; at < off4[(size_t)c + 1] && !found; ++at) {
            const int32_t row = hr[(size_t)(c * 32 + (at - off4[(size_t)c]) % 32)];
            if (row >= 0 && hc[(size_t)at] != row) { a = hv[(size_t)at]; found = true; }
        }
    if (!found || a == 0.0) return LZ_OK;
    double* deff = nullptr;
    LZ_CUDA(cudaMalloc((void**)&deff, (size_t)op->M * 8));
    int* flag = reinterpret_cast<int*>(ctx->scratch + 24);
    LZ_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
    const unsigned cgrid = (unsigned)((sl.nchunks + kWarps - 1) / kWarps);
    sell_uniform_kernel<<<cgrid, kThreads, 0, ctx->stream>>>(sl.chunk_off, sl.col, sl.val, sl.row_of, sl.nchunks, a, deff, flag);
    int bad = 1;
    cudaError_t e = cudaMemcpyAsync(&bad, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess || bad) {
        cudaFree(deff);
        if (e != cudaSuccess) { set_error("sell_detect_uniform: %s", cudaGetErrorString(e)); return LZ_ERR_CUDA; }
        return LZ_OK;
    }
    sl.deff = deff;
    sl.uni_a = a;
    sl.uniform = 1;
    return LZ_OK;
}

// ---- row shards: interior / boundary spans ---------------------------------------------------------
// A span (one sorting window of sigma rows = sigma/32 chunks) is "boundary" when any entry stored in it
// refers to a ghost column (>= M).  With a locality-preserving row order the boundary spans are the few
// windows next to the block ends; everything else can be applied while the ghost entries are in flight.
__global__ void __launch_bounds__(kThreads)
sell_span_ghost_kernel(const int64_t* __restrict__ chunk_off, const int32_t* __restrict__ col, int64_t nchunks,
                       int32_t M, int span, int* __restrict__ span_has_ghost) {
    const int64_t c = ((int64_t)blockIdx.x * kThreads + threadIdx.x) >> 5;
    if (c >= nchunks) return;
    const int lane = threadIdx.x & 31;
    const int64_t o0 = chunk_off[c], o1 = chunk_off[c + 1];
    int any = 0;
    for (int64_t k = o0 + lane; k < o1; k += 32) any |= (col[k] >= M);
    any = __any_sync(0xffffffffu, any);
    if (lane == 0 && any) span_has_ghost[c / span] = 1;
}

bool spmv_split_supported(const lz_op* op) {
    return op->kind == LZ_OP_SELL && op->sell.split_span > 0;     // either list may be empty
}

int sell_classify_spans(lz_op* op) {
    lz_ctx* ctx = op->ctx;
    lz_sell& sl = op->sell;
    if (op->kind != LZ_OP_SELL || op->ncols <= op->M || sl.nchunks == 0) return LZ_OK;
    const int span = std::max(kWarps, sl.sigma / 32);
    const int64_t nspans = (sl.nchunks + span - 1) / span;
    int* flags = nullptr;
    LZ_CUDA(cudaMalloc((void**)&flags, (size_t)nspans * sizeof(int)));
    LZ_CUDA(cudaMemsetAsync(flags, 0, (size_t)nspans * sizeof(int), ctx->stream));
    const unsigned cgrid = (unsigned)((sl.nchunks + kWarps - 1) / kWarps);
    sell_span_ghost_kernel<<<cgrid, kThreads, 0, ctx->stream>>>(sl.chunk_off, sl.col, sl.nchunks, (int32_t)op->M, span, flags);
    std::vector<int> h((size_t)nspans);
    cudaError_t e = cudaMemcpyAsync(h.data(), flags, (size_t)nspans * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(flags);
    if (e != cudaSuccess) { set_error("sell_classify_spans: %s", cudaGetErrorString(e)); return LZ_ERR_CUDA; }
    // Windows are classified; the lists hold work items.  Interior windows: items of balanced_span() chunks
    // (a whole window when the shard is large enough to keep every CTA busy with several).  Boundary windows:
    // items of kWarps chunks (one chunk per warp), so that the few of them spread over many CTAs - one CTA per
    // window would take ~90 us at 50 M rows, a tenth of the whole interior part.
    std::vector<int32_t> in, bd;
    const int ipiece = balanced_span(ctx, sl.nchunks, sl.sigma);           // divides the window
    const int bpiece = (span % kWarps == 0) ? kWarps : span;
    const int ni = span / ipiece, nb = span / bpiece;
    for (int64_t s = 0; s < nspans; ++s) {
        if (!h[(size_t)s]) {
            for (int k = 0; k < ni; ++k)
                if ((s * ni + k) * (int64_t)ipiece < sl.nchunks) in.push_back((int32_t)(s * ni + k));
            continue;
        }
        for (int k = 0; k < nb; ++k)
            if ((s * nb + k) * (int64_t)bpiece < sl.nchunks) bd.push_back((int32_t)(s * nb + k));
    }
    sl.split_span = ipiece;
    sl.bnd_span = bpiece;
    sl.n_int = (int)in.size();
    sl.n_bnd = (int)bd.size();
    LZ_CHECK(upload((void**)&sl.spans_int, in.data(), in.size() * 4, ctx->stream));
    LZ_CHECK(upload((void**)&sl.spans_bnd, bd.data(), bd.size() * 4, ctx->stream));
    LZ_CUDA(cudaStreamSynchronize(ctx->stream));
    return LZ_OK;
}

int launch_spmv_part(lz_op* op, int part, const double* x, const double* scale_dev, double* y, double* partials,
                     int* nparts, const int* flag_dev, const FinTail* fin, cudaStream_t stream) {
    lz_ctx* ctx = op->ctx;
    lz_sell& sl = op->sell;
    LZ_REQUIRE(spmv_split_supported(op) && (part == 1 || part == 2), "launch_spmv_part: operator is not split");
    const int32_t* list = part == 1 ? sl.spans_int : sl.spans_bnd;
    const int nlist = part == 1 ? sl.n_int : sl.n_bnd;
    const int span = part == 1 ? sl.split_span : sl.bnd_span;    // chunks per entry of the list
    // finer grid than the plain launch: CTAs leave the SMs often, so that the one-CTA exchange kernels of the
    // main stream find a slot while the interior part is running
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(nlist, std::min<int64_t>((int64_t)ctx->sms * 16, kMaxPartials / 2)));
    FinTail ft = fin ? *fin : FinTail{};
    double* pout = partials;
    if (part == 1) {
        sl.np_int = grid;
    } else {
        pout = partials + sl.np_int;            // boundary partials follow the interior ones
        if (ft.op.kind != FIN_NONE) { ft.op.extra = partials; ft.op.nextra = sl.np_int; }
    }
    // an empty list still launches one CTA: its partial is 0 and its tail runs the bookkeeping / the exchange
    if (part == 1 && sl.windowed) {
        int wgrid = 0;
        LZ_CHECK(launch_spmv_windowed(op, list, nlist, x, scale_dev, y, pout, &wgrid, flag_dev, ft, stream));
        sl.np_int = wgrid;
        if (nparts) *nparts = wgrid;
        return LZ_OK;
    }
    if (sl.uniform)
        LZ_CUDA(launch_k(spmv_sell_dot_kernel<true>, dim3(grid), dim3(kThreads), 0, stream, sl.chunk_off, sl.col, sl.val,
                         sl.row_of, x, scale_dev, y, sl.nchunks, pout, op->xghost, (int32_t)op->M, span, ft,
                         list, nlist, flag_dev, (const double*)sl.deff, sl.uni_a));
    else
        LZ_CUDA(launch_k(spmv_sell_dot_kernel<false>, dim3(grid), dim3(kThreads), 0, stream, sl.chunk_off, sl.col, sl.val,
                         sl.row_of, x, scale_dev, y, sl.nchunks, pout, op->xghost, (int32_t)op->M, span, ft,
                         list, nlist, flag_dev, (const double*)nullptr, 0.0));
    if (nparts) *nparts = (part == 1) ? grid : sl.np_int + grid;
    return LZ_OK;
}

// Ghost-index exchange of a row-sharded sparse operator: entry e of my send list goes to the
// gather buffer of the rank that owns segment q (seg_start[q] <= e < seg_start[q+1]) at offset
// dst_off[q] + (e - seg_start[q]) - a gather from my vector followed by NVLink peer stores.
struct GhostPushArgs {
    const int32_t* send_idx;      // device, nsend local row indices
    int nsend;
    int world;
    int seg_start[17];            // prefix over peers
    double* dst[16];              // peer gather buffer (parity already applied) + dst_off
};

__global__ void __launch_bounds__(kThreads)
ghost_push_kernel(const double* __restrict__ x, const GhostPushArgs g, const int* __restrict__ flag) {
    pdl_prologue();
    if (flag && *flag == 0) return;
    const int tid = blockIdx.x * kThreads + threadIdx.x;
    const int nthr = gridDim.x * kThreads;
    for (int e = tid; e < g.nsend; e += nthr) {
        int q = 0;
        while (e >= g.seg_start[q + 1]) ++q;
        g.dst[q][e - g.seg_start[q]] = __ldg(x + __ldg(g.send_idx + e));
    }
}

int launch_ghost_push(lz_ctx* ctx, const double* x, const int32_t* send_idx, int nsend, int world,
                      const int* seg_start, double* const* dst, const int* flag_dev, cudaStream_t stream) {
    if (nsend <= 0) return LZ_OK;
    GhostPushArgs g;
    g.send_idx = send_idx;
    g.nsend = nsend;
    g.world = world;
    for (int q = 0; q <= world; ++q) g.seg_start[q] = seg_start[q];
    for (int q = world + 1; q < 17; ++q) g.seg_start[q] = nsend;
    for (int q = 0; q < 16; ++q) g.dst[q] = q < world ? dst[q] : nullptr;
    const int grid = std::max(1, std::min((nsend + kThreads - 1) / kThreads, ctx->sms * 4));
    LZ_CUDA(launch_k(ghost_push_kernel, dim3(grid), dim3(kThreads), 0, stream ? stream : ctx->stream, x, g, flag_dev));
    return LZ_OK;
}

// ---- host-side construction -------------------------------------------------------------------

static int upload(void** dev, const void* host, size_t bytes, cudaStream_t s) {
    LZ_CUDA(cudaMalloc(dev, bytes ? bytes : 16));
    if (bytes) LZ_CUDA(cudaMemcpyAsync(*dev, host, bytes, cudaMemcpyHostToDevice, s));
    return LZ_OK;
}

int build_csr(lz_op* op, int64_t M, int64_t nnz, const int32_t* indptr, const int32_t* indices,
              const double* data) {
    lz_ctx* ctx = op->ctx;
    op->kind = LZ_OP_CSR;
    op->M = M;
    op->csr.nnz = nnz;
    LZ_CHECK(upload((void**)&op->csr.indptr, indptr, (size_t)(M + 1) * 4, ctx->stream));
    LZ_CHECK(upload((void**)&op->csr.indices, indices, (size_t)nnz * 4, ctx->stream));
    LZ_CHECK(upload((void**)&op->csr.data, data, (size_t)nnz * 8, ctx->stream));
    const double mean = M > 0 ? (double)nnz / (double)M : 1.0;
    op->csr.lanes_per_row = mean <= 2.5 ? 2 : mean <= 5.0 ? 4 : mean <= 12.0 ? 8 : mean <= 24.0 ? 16 : 32;
    LZ_CUDA(cudaStreamSynchronize(ctx->stream));
    return LZ_OK;
}

int build_sell(lz_op* op, int64_t M, int64_t nnz, const int32_t* indptr, const int32_t* indices,
               const double* data, int sigma) {
    lz_ctx* ctx = op->ctx;
    if (sigma <= 0) sigma = 2048;
    sigma = (sigma + 31) / 32 * 32;
    const int64_t nchunks = (M + 31) / 32;
    std::vector<int32_t> row_of((size_t)nchunks * 32, -1);
    // sort rows by descending length inside each window of sigma rows (stable)
    {
        std::vector<int32_t> idx;
        for (int64_t w0 = 0; w0 < M; w0 += sigma) {
            const int64_t w1 = std::min<int64_t>(M, w0 + sigma);
            idx.resize((size_t)(w1 - w0));
            std::iota(idx.begin(), idx.end(), (int32_t)w0);
            std::stable_sort(idx.begin(), idx.end(), [&](int32_t a, int32_t b) {
                return (indptr[a + 1] - indptr[a]) > (indptr[b + 1] - indptr[b]);
            });
            std::copy(idx.begin(), idx.end(), row_of.begin() + w0);
        }
    }
    std::vector<int64_t> chunk_off((size_t)nchunks + 1, 0);
    for (int64_t c = 0; c < nchunks; ++c) {
        int width = 0;
        for (int l = 0; l < 32; ++l) {
            const int32_t r = row_of[(size_t)c * 32 + l];
            if (r >= 0) width = std::max(width, (int)(indptr[r + 1] - indptr[r]));
        }
        chunk_off[(size_t)c + 1] = chunk_off[(size_t)c] + (int64_t)width * 32;
    }
    const int64_t stored = chunk_off[(size_t)nchunks];
    std::vector<int32_t> col((size_t)stored);
    std::vector<double> val((size_t)stored);
    for (int64_t c = 0; c < nchunks; ++c) {
        const int64_t o0 = chunk_off[(size_t)c];
        const int width = (int)((chunk_off[(size_t)c + 1] - o0) / 32);
        for (int l = 0; l < 32; ++l) {
            const int32_t r = row_of[(size_t)c * 32 + l];
            const int32_t k0 = r >= 0 ? indptr[r] : 0;
            const int len = r >= 0 ? (int)(indptr[r + 1] - indptr[r]) : 0;
            // padding entries multiply the row's own x by 0: the gather stays local
            const int32_t pad_col = r >= 0 ? r : 0;
            for (int k = 0; k < width; ++k) {
                const size_t at = (size_t)(o0 + (int64_t)k * 32 + l);
                if (k < len) { col[at] = indices[k0 + k]; val[at] = data[k0 + k]; }
                else { col[at] = pad_col; val[at] = 0.0; }
            }
        }
    }
    op->kind = LZ_OP_SELL;
    op->M = M;
    op->sell.nnz_true = nnz;
    op->sell.nnz_stored = stored;
    op->sell.nchunks = nchunks;
    op->sell.sigma = sigma;
    LZ_CHECK(upload((void**)&op->sell.chunk_off, chunk_off.data(), chunk_off.size() * 8, ctx->stream));
    LZ_CHECK(upload((void**)&op->sell.col, col.data(), col.size() * 4, ctx->stream));
    LZ_CHECK(upload((void**)&op->sell.val, val.data(), val.size() * 8, ctx->stream));
    LZ_CHECK(upload((void**)&op->sell.row_of, row_of.data(), row_of.size() * 4, ctx->stream));
    LZ_CUDA(cudaStreamSynchronize(ctx->stream));
    return LZ_OK;
}

// ---- device-side construction (operator already resident in HBM) --------------------------------
// The reference's GPU mode holds H as a cupyx matrix on the device (Lanczos.py:88,137); at the sizes
// of BASELINE config 4 (5e7 rows, 7e8 entries) a round trip through host CSR would dominate set-up,
// so the SELL-32-sigma conversion also exists as kernels.  It reproduces build_sell bit for bit
// (stable descending-length order inside each window), which tests/test_gpu_parity.py checks.

// error bits: 1 indptr not monotone / does not span [0, nnz], 2 column out of range
__global__ void __launch_bounds__(kThreads)
csr_validate_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices, int64_t M,
                    int64_t ncols, int64_t nnz, int* __restrict__ err) {
    const int64_t tid = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * kThreads;
    int bad = 0;
    if (tid == 0 && (indptr[0] != 0 || (int64_t)indptr[M] != nnz)) bad |= 1;
    for (int64_t i = tid; i < M; i += nthr)
        if (indptr[i + 1] < indptr[i]) bad |= 1;
    for (int64_t k = tid; k < nnz; k += nthr) {
        const int32_t c = indices[k];
        if (c < 0 || c >= ncols) bad |= 2;
    }
    if (bad) atomicOr(err, bad);
}

// One CTA per window of sigma rows: slot of row t = #(longer rows) + #(equally long rows before it).
__global__ void __launch_bounds__(kThreads)
sell_order_kernel(const int32_t* __restrict__ indptr, int64_t M, int sigma, int32_t* __restrict__ row_of) {
    extern __shared__ int32_t s_len[];
    const int64_t w0 = (int64_t)blockIdx.x * sigma;
    const int cnt = (int)min((int64_t)sigma, M - w0);
    for (int t = threadIdx.x; t < cnt; t += kThreads) s_len[t] = indptr[w0 + t + 1] - indptr[w0 + t];
    __syncthreads();
    for (int t = threadIdx.x; t < cnt; t += kThreads) {
        const int32_t mine = s_len[t];
        int slot = 0;
        for (int u = 0; u < cnt; ++u) {
            const int32_t l = s_len[u];
            slot += (l > mine) || (l == mine && u < t);
        }
        row_of[w0 + slot] = (int32_t)(w0 + t);
    }
}

// One warp per chunk: stored entries of the chunk = 32 * (longest of its rows).
__global__ void __launch_bounds__(kThreads)
sell_width_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ row_of, int64_t nchunks,
                  int64_t* __restrict__ stored) {
    const int64_t c = ((int64_t)blockIdx.x * kThreads + threadIdx.x) >> 5;
    if (c >= nchunks) return;
    const int lane = threadIdx.x & 31;
    const int32_t r = row_of[c * 32 + lane];
    int w = r >= 0 ? indptr[r + 1] - indptr[r] : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w = max(w, __shfl_xor_sync(0xffffffffu, w, o));
    if (lane == 0) stored[c] = (int64_t)w * 32;
}

__global__ void __launch_bounds__(kThreads)
sell_fill_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                 const double* __restrict__ data, const int32_t* __restrict__ row_of,
                 const int64_t* __restrict__ chunk_off, int64_t nchunks, int32_t* __restrict__ col,
                 double* __restrict__ val) {
    const int64_t c = ((int64_t)blockIdx.x * kThreads + threadIdx.x) >> 5;
    if (c >= nchunks) return;
    const int lane = threadIdx.x & 31;
    const int64_t o0 = chunk_off[c];
    const int width = (int)((chunk_off[c + 1] - o0) >> 5);
    const int32_t r = row_of[c * 32 + lane];
    const int32_t k0 = r >= 0 ? indptr[r] : 0;
    const int len = r >= 0 ? indptr[r + 1] - k0 : 0;
    const int32_t pad_col = r >= 0 ? r : 0;
    for (int k = 0; k < width; ++k) {
        const int64_t at = o0 + (int64_t)k * 32 + lane;
        if (k < len) { col[at] = indices[k0 + k]; val[at] = data[k0 + k]; }
        else { col[at] = pad_col; val[at] = 0.0; }
    }
}

// exclusive prefix of n int64 values by one CTA (set-up only; n = number of chunks + 1)
__global__ void __launch_bounds__(1024)
scan_i64_kernel(int64_t* __restrict__ a, int64_t n) {
    __shared__ int64_t s_tot[1024];
    const int t = threadIdx.x;
    const int64_t per = (n + 1023) / 1024;
    const int64_t b = min(n, t * per), e = min(n, b + per);
    int64_t sum = 0;
    for (int64_t i = b; i < e; ++i) sum += a[i];
    s_tot[t] = sum;
    __syncthreads();
    if (t == 0) {
        int64_t run = 0;
        for (int i = 0; i < 1024; ++i) { const int64_t v = s_tot[i]; s_tot[i] = run; run += v; }
    }
    __syncthreads();
    int64_t run = s_tot[t];
    for (int64_t i = b; i < e; ++i) { const int64_t v = a[i]; a[i] = run; run += v; }
}

int scan_i64(int64_t* a_dev, int64_t n, cudaStream_t q) {
    scan_i64_kernel<<<1, 1024, 0, q>>>(a_dev, n);
    LZ_CUDA(cudaGetLastError());
    return LZ_OK;
}

int build_from_device(lz_op* op, int64_t M, int64_t ncols, int64_t nnz, const int32_t* indptr,
                      const int32_t* indices, const double* data, int fmt, int sigma) {
    lz_ctx* ctx = op->ctx;
    cudaStream_t q = ctx->stream;
    const int wide = ctx->sms * 8;
    // validate on the device (same conditions as the host path)
    int* err = reinterpret_cast<int*>(ctx->scratch);
    LZ_CUDA(cudaMemsetAsync(err, 0, sizeof(int), q));
    csr_validate_kernel<<<wide, kThreads, 0, q>>>(indptr, indices, M, ncols, nnz, err);
    int h_err = 0;
    LZ_CUDA(cudaMemcpyAsync(&h_err, err, sizeof(int), cudaMemcpyDeviceToHost, q));
    LZ_CUDA(cudaStreamSynchronize(q));
    LZ_REQUIRE(!(h_err & 1), "lz_op_csr_create_dev: indptr is not monotone or does not span [0, nnz]");
    LZ_REQUIRE(!(h_err & 2), "lz_op_csr_create_dev: column index out of range");
    op->M = M;
    if (fmt == LZ_FMT_CSR) {
        op->kind = LZ_OP_CSR;
        op->csr.nnz = nnz;
        LZ_CUDA(cudaMalloc((void**)&op->csr.indptr, (size_t)(M + 1) * 4));
        LZ_CUDA(cudaMalloc((void**)&op->csr.indices, nnz ? (size_t)nnz * 4 : 16));
        LZ_CUDA(cudaMalloc((void**)&op->csr.data, nnz ? (size_t)nnz * 8 : 16));
        LZ_CUDA(cudaMemcpyAsync(op->csr.indptr, indptr, (size_t)(M + 1) * 4, cudaMemcpyDeviceToDevice, q));
        if (nnz) {
            LZ_CUDA(cudaMemcpyAsync(op->csr.indices, indices, (size_t)nnz * 4, cudaMemcpyDeviceToDevice, q));
            LZ_CUDA(cudaMemcpyAsync(op->csr.data, data, (size_t)nnz * 8, cudaMemcpyDeviceToDevice, q));
        }
        const double mean = (double)nnz / (double)M;
        op->csr.lanes_per_row = mean <= 2.5 ? 2 : mean <= 5.0 ? 4 : mean <= 12.0 ? 8 : mean <= 24.0 ? 16 : 32;
        LZ_CUDA(cudaStreamSynchronize(q));
        return LZ_OK;
    }
    if (sigma <= 0) sigma = 2048;
    sigma = (sigma + 31) / 32 * 32;
    LZ_REQUIRE(sigma <= 8192, "lz_op_csr_create_dev: sigma must be <= 8192 rows");
    const int64_t nchunks = (M + 31) / 32;
    lz_sell& sl = op->sell;
    op->kind = LZ_OP_SELL;
    sl.nnz_true = nnz;
    sl.nchunks = nchunks;
    sl.sigma = sigma;
    LZ_CUDA(cudaMalloc((void**)&sl.row_of, (size_t)nchunks * 32 * 4));
    LZ_CUDA(cudaMalloc((void**)&sl.chunk_off, (size_t)(nchunks + 1) * 8));
    LZ_CUDA(cudaMemsetAsync(sl.row_of, 0xff, (size_t)nchunks * 32 * 4, q));          // -1: padding rows
    LZ_CUDA(cudaMemsetAsync(sl.chunk_off, 0, (size_t)(nchunks + 1) * 8, q));
    const int64_t nwin = (M + sigma - 1) / sigma;
    sell_order_kernel<<<(unsigned)nwin, kThreads, (size_t)sigma * 4, q>>>(indptr, M, sigma, sl.row_of);
    const unsigned cgrid = (unsigned)((nchunks + kWarps - 1) / kWarps);
    sell_width_kernel<<<cgrid, kThreads, 0, q>>>(indptr, sl.row_of, nchunks, sl.chunk_off);
    scan_i64_kernel<<<1, 1024, 0, q>>>(sl.chunk_off, nchunks + 1);
    int64_t stored = 0;
    LZ_CUDA(cudaMemcpyAsync(&stored, sl.chunk_off + nchunks, 8, cudaMemcpyDeviceToHost, q));
    LZ_CUDA(cudaStreamSynchronize(q));
    LZ_CUDA(cudaGetLastError());
    sl.nnz_stored = stored;
    LZ_CUDA(cudaMalloc((void**)&sl.col, stored ? (size_t)stored * 4 : 16));
    LZ_CUDA(cudaMalloc((void**)&sl.val, stored ? (size_t)stored * 8 : 16));
    sell_fill_kernel<<<cgrid, kThreads, 0, q>>>(indptr, indices, data, sl.row_of, sl.chunk_off, nchunks, sl.col, sl.val);
    LZ_CUDA(cudaGetLastError());
    LZ_CUDA(cudaStreamSynchronize(q));
    return LZ_OK;
}

}  // namespace lz
