// KP: the whole Lanczos solve of a SMALL matrix-free problem in one persistent cooperative kernel.
//
// Up to a few hundred thousand unknowns a Lanczos step is a handful of kernels of a few microseconds
// each, and what it costs is the chain of launches, drains and one-CTA reductions between them (config 1,
// 200 x 200, n = 100: 39 us/step in round 1, 22 us/step as a replayed CUDA graph, of which ~2 us is
// arithmetic).  Here one grid of <= 148 CTAs stays resident for the whole solve - the pre-step, n steps and
// the Gram-Schmidt sweeps of Lanczos.py:100-119 / :233-251 - and the phases of a step are separated by
// grid-wide barriers (one atomic + one spin on a counter in L2, ~1 us) instead of kernel boundaries:
//
//   beta = |r|                               partial per CTA -> barrier -> every CTA adds the partials in CTA order
//   V[j] = r / beta                          (rows are stored normalised here: no lazy factor)
//   sweep:  ip_i = V[j].V[i]                 warp w of a CTA takes rows w, w+8, ..: CTA-slice dot -> partial[cta][i]
//           barrier; coef_i = sum_cta        thread i adds the partials of row i in CTA order
//           V[j] = c V[j] - sum coef_i V[i]  own elements; barrier (the neighbours read the new row)
//   w = H V[j], alpha = V[j].w               neighbours from global memory (L2); partial -> barrier -> sum
//   r = w - alpha V[j] - beta V[j-1]         registers
//
// A thread owns elements g, g + T, g + 2T, ... (g = global thread id, T = threads of the grid, at most 8 per
// thread), so every vector operation is coalesced and r, V[j], V[j-1] live in registers.  All sums are taken
// in a fixed order (thread -> warp tree -> CTA -> CTA order), so results are bit-reproducible and identical in
// every CTA.  Selected by lz_lanczos_run for structured-grid operators with M <= 148 * 256 * 8 on one GPU,
// reorth full or none; everything else takes the kernel-per-phase path (lanczos.cu).
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "internal.h"

namespace lz {

int arena_reserve(lz_ctx* ctx, size_t bytes);            // lanczos.cu

constexpr int kSmallEpt = 8;                             // elements per thread, at most

struct SmallArgs {
    // operator
    int nx, ny, nz, periodic;
    double c, ox, oy, oz;
    const double* diag;
    int64_t M;
    // run
    int n, ref, reorth_full, passes, gpu_sweep;
    double tol_rel;
    const double* v0;
    double* V;
    int64_t ldv;
    // workspace
    double* alpha;        // [n]
    double* beta;         // [n + 1]
    double* red;          // [2][grid]  scalar partials, double-buffered by reduction parity
    double* dpart;        // [grid][ldp] dot partials of a sweep
    int ldp;
    unsigned int* bar;    // grid barrier counter (zero at launch)
    int* flags;           // [0] first breakdown step (-1: none)
    double* tmp;          // [M] start vector normalised (pre-step neighbours)
};

__device__ __forceinline__ unsigned int ld_acq_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_cg(const double* p) {       // coherent at L2: written by other CTAs of this grid
    double v;
    asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// Grid barrier: arrivals are counted with one atomic per CTA; the last arriver publishes the epoch on a
// SEPARATE cache line that everybody else polls - pollers and atomics never fight over a line (with all
// CTAs spinning on the counter itself a barrier cost ~7 us on 148 SMs).
struct GridSync {
    unsigned int* ctr;              // ctr[0]: arrivals (monotonic), ctr[32]: epoch released (128 B further on)
    unsigned int target = 0;
    unsigned int epoch = 0;
    __device__ __forceinline__ void sync() {
        __syncthreads();
        if (threadIdx.x == 0) {
            target += gridDim.x;
            epoch += 1;
            // release + acquire at gpu scope: the writes of this CTA (ordered before this thread by the barrier above) are visible to
            // whoever acquires the epoch; cheaper than a full fence on either side of a relaxed atomic
            unsigned int old;
            asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(ctr) : "memory");
            if (old == target - 1) {
                asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(ctr + 32), "r"(epoch) : "memory");
            } else {
                while (ld_acq_u32(ctr + 32) < epoch) {}
            }
        }
        __syncthreads();
    }
};

// sum over the grid of one value per thread; every thread of every CTA gets the same bits
__device__ __forceinline__ double grid_sum(double v, const SmallArgs& a, GridSync& gs, int& parity, double* sred,
                                           double* sbc) {
    const double t = block_sum(v, sred);
    double* buf = a.red + (size_t)parity * gridDim.x;
    parity ^= 1;
    if (threadIdx.x == 0) buf[blockIdx.x] = t;
    gs.sync();
    double s = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += kThreads) s += ld_cg(buf + i);
    const double tot = block_sum(s, sred);
    if (threadIdx.x == 0) sbc[0] = tot;
    __syncthreads();
    const double out = sbc[0];
    __syncthreads();
    return out;
}

// (H x)_e with x read from global memory; the same operation order as stencil_apply_dot_kernel
__device__ __forceinline__ double apply_at(const SmallArgs& a, const double* x, int64_t e) {
    const int ix = (int)(e % a.nx);
    const int64_t t = e / a.nx;
    const int iy = (int)(t % a.ny);
    const int iz = (int)(t / a.ny);
    auto nb = [&](int i, int n, int d, bool& ok) -> int {
        int k = i + d;
        ok = true;
        if (k < 0) { if (a.periodic) k = n - 1; else ok = false; }
        else if (k >= n) { if (a.periodic) k = 0; else ok = false; }
        return k;
    };
    bool ok;
    const int64_t plane = (int64_t)a.nx * a.ny;
    double vm = 0.0, vp = 0.0, ym = 0.0, yp = 0.0, xl = 0.0, xr = 0.0;
    int k;
    if (a.oz != 0.0) {
        k = nb(iz, a.nz, -1, ok); if (ok) vm = ld_cg(x + (int64_t)k * plane + (int64_t)iy * a.nx + ix);
        k = nb(iz, a.nz, +1, ok); if (ok) vp = ld_cg(x + (int64_t)k * plane + (int64_t)iy * a.nx + ix);
    }
    if (a.oy != 0.0) {
        k = nb(iy, a.ny, -1, ok); if (ok) ym = ld_cg(x + (int64_t)iz * plane + (int64_t)k * a.nx + ix);
        k = nb(iy, a.ny, +1, ok); if (ok) yp = ld_cg(x + (int64_t)iz * plane + (int64_t)k * a.nx + ix);
    }
    k = nb(ix, a.nx, -1, ok); if (ok) xl = ld_cg(x + (int64_t)iz * plane + (int64_t)iy * a.nx + k);
    k = nb(ix, a.nx, +1, ok); if (ok) xr = ld_cg(x + (int64_t)iz * plane + (int64_t)iy * a.nx + k);
    const double vc = ld_cg(x + e);
    const double dg = a.diag ? __ldg(a.diag + e) : 0.0;
    double r = a.oz * vm;
    r = fma(a.oy, ym, r);
    r = fma(a.ox, xl, r);
    r = fma(a.c + dg, vc, r);
    r = fma(a.ox, xr, r);
    r = fma(a.oy, yp, r);
    r = fma(a.oz, vp, r);
    return r;
}

__global__ void __launch_bounds__(kThreads, 1)
small_lanczos_kernel(const SmallArgs a) {
    __shared__ double sred[kWarps];
    __shared__ double sbc[2];
    __shared__ double vs[kSmallEpt * kThreads];          // the CTA's slice of V[j] (for the row dots)
    extern __shared__ double scoef[];                    // [n + 1] sweep coefficients
    __shared__ double sseg[8 * kThreads];                // partial coefficient sums of the CTA runs
    GridSync gs{a.bar};
    int parity = 0;
    const int64_t T = (int64_t)gridDim.x * kThreads;
    const int64_t g = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ept = (int)((a.M + T - 1) / T);             // elements per thread actually in use
    int64_t el[kSmallEpt];
    bool on[kSmallEpt];
#pragma unroll
    for (int k = 0; k < kSmallEpt; ++k) { el[k] = g + k * T; on[k] = el[k] < a.M; }
    double r[kSmallEpt], v[kSmallEpt], vprev[kSmallEpt];
#pragma unroll
    for (int k = 0; k < kSmallEpt; ++k) { r[k] = 0.0; v[k] = 0.0; vprev[k] = 0.0; }

    // ---- start vector: q = v0 / |v0| ------------------------------------------------------------
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < kSmallEpt; ++k) if (on[k]) { v[k] = __ldg(a.v0 + el[k]); acc = fma(v[k], v[k], acc); }
    const double nrm0 = sqrt(grid_sum(acc, a, gs, parity, sred, sbc));
    if (!(nrm0 > 0.0) && g == 0 && a.flags[0] < 0) a.flags[0] = 0;
    const double s0 = nrm0 > 0.0 ? 1.0 / nrm0 : 0.0;
    double beta_prev = 0.0;                               // beta_j of the coming step
    if (a.ref) {
        // pre-step (Lanczos.py:108-110): r = H q - (q.Hq) q with q = v0/|v0|; q is discarded
#pragma unroll
        for (int k = 0; k < kSmallEpt; ++k) if (on[k]) { v[k] *= s0; a.tmp[el[k]] = v[k]; }
        gs.sync();
        acc = 0.0;
#pragma unroll
        for (int k = 0; k < kSmallEpt; ++k) if (on[k]) { r[k] = apply_at(a, a.tmp, el[k]); acc = fma(r[k], v[k], acc); }
        const double apre = grid_sum(acc, a, gs, parity, sred, sbc);
#pragma unroll
        for (int k = 0; k < kSmallEpt; ++k) { r[k] = fma(-apre, v[k], r[k]); v[k] = 0.0; }
    } else {
        // clean start: q_0 = v0/|v0| is row 0
#pragma unroll
        for (int k = 0; k < kSmallEpt; ++k) r[k] = v[k];
    }

    // slice dots of the vector held in `vs` against rows 0 .. nrows-1 -> dpart[cta][i]; four rows in flight per warp:
    // the loads (L2, ~0.5 us away) of a group are issued before any is used
    auto slice_dots = [&](int nrows) {
        for (int i0 = warp; i0 < nrows; i0 += 4 * kWarps) {
            double d[4] = {0.0, 0.0, 0.0, 0.0};
            for (int k = 0; k < ept; ++k) {
                double x[4][kWarps];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + u * kWarps;
                    const double* vi = a.V + (int64_t)min(i, nrows - 1) * a.ldv;
#pragma unroll
                    for (int m = 0; m < kWarps; ++m) {
                        const int64_t e = (int64_t)blockIdx.x * kThreads + lane + 32 * m + k * T;
                        x[u][m] = (e < a.M) ? ld_cg(vi + e) : 0.0;
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int m = 0; m < kWarps; ++m) d[u] = fma(vs[k * kThreads + lane + 32 * m], x[u][m], d[u]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * kWarps;
                const double t = warp_sum(d[u]);
                if (lane == 0 && i < nrows) a.dpart[(size_t)blockIdx.x * a.ldp + i] = t;
            }
        }
    };
    // scoef[i] = inv * sum over the CTAs of dpart[cta][i], CTAs in order.  The CTAs are cut into nseg runs so that all
    // 256 threads load (thread = (run, row)); a run is summed in order with up to 32 loads in flight, then the runs
    // are added in order.
    auto coefficients = [&](int nrows, double inv) {
        const int G = (int)gridDim.x;
        for (int i0 = 0; i0 < nrows; i0 += kThreads) {
            const int R = min(nrows - i0, kThreads);
            const int nseg = max(1, min(kThreads / R, 8));
            const int per = (G + nseg - 1) / nseg;
            const int seg = threadIdx.x / R, i = i0 + threadIdx.x % R;
            double sacc = 0.0;
            if (seg < nseg) {
                const int c0 = seg * per, c1 = min(G, c0 + per);
                int cta = c0;
                for (; cta + 32 <= c1; cta += 32) {
                    double x[32];
#pragma unroll
                    for (int u = 0; u < 32; ++u) x[u] = ld_cg(a.dpart + (size_t)(cta + u) * a.ldp + i);
#pragma unroll
                    for (int u = 0; u < 32; ++u) sacc += x[u];
                }
                for (; cta + 8 <= c1; cta += 8) {
                    double x[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) x[u] = ld_cg(a.dpart + (size_t)(cta + u) * a.ldp + i);
#pragma unroll
                    for (int u = 0; u < 8; ++u) sacc += x[u];
                }
                for (; cta < c1; ++cta) sacc += ld_cg(a.dpart + (size_t)cta * a.ldp + i);
                sseg[seg * kThreads + (i - i0)] = sacc;
            }
            __syncthreads();
            if ((int)threadIdx.x < R) {
                double t = 0.0;
                for (int q = 0; q < nseg; ++q) t += sseg[q * kThreads + threadIdx.x];
                scoef[i0 + threadIdx.x] = inv * t;
            }
            __syncthreads();
        }
    };
    // v = cself * v - sum_i scoef[i] V[i] on the thread's own elements; 32 rows of every element in flight,
    // subtracted in row order
    auto subtract_rows = [&](int nrows, double cself) {
        double t[kSmallEpt];
#pragma unroll
        for (int k = 0; k < kSmallEpt; ++k) t[k] = cself * v[k];
        int i = 0;
        for (; i + 32 <= nrows; i += 32) {
#pragma unroll
            for (int k = 0; k < kSmallEpt; ++k) {
                if (!on[k]) continue;
                const double* col = a.V + el[k];
                double x[32];
#pragma unroll
                for (int u = 0; u < 32; ++u) x[u] = ld_cg(col + (int64_t)(i + u) * a.ldv);
#pragma unroll
                for (int u = 0; u < 32; ++u) t[k] = fma(-scoef[i + u], x[u], t[k]);
            }
        }
        for (; i + 8 <= nrows; i += 8) {
#pragma unroll
            for (int k = 0; k < kSmallEpt; ++k) {
                if (!on[k]) continue;
                const double* col = a.V + el[k];
                double x[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) x[u] = ld_cg(col + (int64_t)(i + u) * a.ldv);
#pragma unroll
                for (int u = 0; u < 8; ++u) t[k] = fma(-scoef[i + u], x[u], t[k]);
            }
        }
        for (; i < nrows; ++i) {
#pragma unroll
            for (int k = 0; k < kSmallEpt; ++k)
                if (on[k]) t[k] = fma(-scoef[i], ld_cg(a.V + el[k] + (int64_t)i * a.ldv), t[k]);
        }
#pragma unroll
        for (int k = 0; k < kSmallEpt; ++k) if (on[k]) v[k] = t[k];
    };

    for (int j = 0; j < a.n; ++j) {
        double* row = a.V + (int64_t)j * a.ldv;
        const bool sweep = a.reorth_full && (j > 0 || a.ref);
        const int nrows = j;                                  // rows before j; the self term is handled separately
        // ---- beta_j = |r|, V[j] = r / beta_j; with a sweep, its first pass shares the barrier: the slice dots are
        //      taken with r (V[i].r, divided by beta afterwards) and published together with the partial of |r|^2
        double beta, sumsq = 0.0;
        const bool fused_first = sweep && !(j == 0 && !a.ref);
        if (j == 0 && !a.ref) beta = nrm0;
        else {
            acc = 0.0;
#pragma unroll
            for (int k = 0; k < kSmallEpt; ++k) acc = fma(r[k], r[k], acc);
            if (fused_first && nrows > 0) {
#pragma unroll
                for (int k = 0; k < kSmallEpt; ++k) vs[k * kThreads + threadIdx.x] = r[k];
                __syncthreads();
                slice_dots(nrows);
            }
            sumsq = grid_sum(acc, a, gs, parity, sred, sbc);   // (its barrier also publishes dpart)
            beta = sqrt(sumsq);
        }
        if (g == 0) {
            a.beta[j] = beta;
            const double mag = (j == 0) ? 0.0 : fabs(a.alpha[0]);
            const bool ok = isfinite(beta) && beta > a.tol_rel * mag && beta > 0.0;
            if (!ok && a.flags[0] < 0) a.flags[0] = j;
        }
        const double inv_beta = (beta > 0.0) ? 1.0 / beta : 0.0;
#pragma unroll
        for (int k = 0; k < kSmallEpt; ++k) { vprev[k] = v[k]; v[k] = (beta > 0.0) ? r[k] / beta : 0.0; }
        // ---- Gram-Schmidt sweeps against the rows before (and, in the reference's form, including) row j --
        if (sweep) {
            for (int p = 0; p < a.passes; ++p) {
                const int ref_form = (a.ref && p == 0 && !a.gpu_sweep) ? 1 : 0;
                if (nrows == 0 && !ref_form) continue;
                double self = 0.0, inv = 1.0;
                if (p == 0 && fused_first) {
                    // |V[j]|^2 = |r|^2 / beta^2 (1 to rounding; the reference sums the squares of the normalised row)
                    self = (beta > 0.0) ? sumsq / (beta * beta) : 0.0;
                    inv = inv_beta;
                } else {
                    // the CTA's slice of V[j] into shared memory, then warp w takes rows w, w + 8, ...
#pragma unroll
                    for (int k = 0; k < kSmallEpt; ++k) vs[k * kThreads + threadIdx.x] = v[k];
                    __syncthreads();
                    slice_dots(nrows);
                    if (ref_form) {                           // (2 - |v|^2): the reference's sum includes row j itself
                        double q = 0.0;
#pragma unroll
                        for (int k = 0; k < kSmallEpt; ++k) q = fma(v[k], v[k], q);
                        self = grid_sum(q, a, gs, parity, sred, sbc);      // (its barrier also publishes dpart)
                    } else {
                        gs.sync();
                    }
                }
                coefficients(nrows, inv);
                subtract_rows(nrows, ref_form ? 2.0 - self : 1.0);
                __syncthreads();                          // scoef / vs are reused by the next pass
                if (p + 1 < a.passes) gs.sync();          // dpart is rewritten by the next pass
            }
        }
#pragma unroll
        for (int k = 0; k < kSmallEpt; ++k) if (on[k]) row[el[k]] = v[k];
        gs.sync();                                        // the neighbours read the new row
        // ---- w = H V[j], alpha_j = V[j].w; r = w - alpha_j V[j] - beta_j V[j-1] -----------------------
        acc = 0.0;
#pragma unroll
        for (int k = 0; k < kSmallEpt; ++k) if (on[k]) { r[k] = apply_at(a, row, el[k]); acc = fma(r[k], v[k], acc); }
        const double alpha = grid_sum(acc, a, gs, parity, sred, sbc);
        if (g == 0) a.alpha[j] = alpha;
#pragma unroll
        for (int k = 0; k < kSmallEpt; ++k) {
            double t = fma(-alpha, v[k], r[k]);
            if (j > 0) t = fma(-beta, vprev[k], t);
            r[k] = t;
        }
        beta_prev = beta;
    }
    (void)beta_prev;
}

bool small_solve_supported(const lz_op* op, const lz_run_opts* opts, int32_t n, const double* V_dev, int64_t ldv) {
    static const bool off = []() { const char* e = getenv("LZ_SMALL"); return e && e[0] == '0'; }();
    if (off || !(opts->flags & 32) || (opts->flags & (4 | 16)) || op->kind != LZ_OP_STENCIL || op->st.points != 7 || op->st.sharded) return false;
    if (!V_dev || ldv < op->M || opts->profile) return false;
    if (opts->reorth == LZ_REORTH_SELECTIVE || opts->step_kernel != 0) return false;
    const int64_t cap = (int64_t)op->ctx->sms * kThreads * kSmallEpt;
    return op->M <= cap && n >= 1 && n <= 4096;
}

int launch_small_solve(lz_ctx* ctx, lz_op* op, const double* v0_dev, int32_t n, const lz_run_opts* opts,
                       double* alpha_host, double* beta_host, double* V_dev, int64_t ldv, double* row_scale_host,
                       lz_run_info* info) {
    const lz_stencil& st = op->st;
    const int passes = opts->cgs_passes <= 0 ? 1 : opts->cgs_passes;
    int grid = (int)std::min<int64_t>(ctx->sms, (op->M + kThreads - 1) / kThreads);
    // LZ_SMALL_GRID: fewer CTAs make the grid barriers cheaper and every phase longer (tuning knob)
    if (const char* e = getenv("LZ_SMALL_GRID")) {
        const int64_t least = (op->M + (int64_t)kThreads * kSmallEpt - 1) / ((int64_t)kThreads * kSmallEpt);
        grid = (int)std::max<int64_t>(least, std::min<int64_t>(grid, atoi(e)));
    }
    grid = std::max(grid, 1);
    const size_t nd = (size_t)n + 2;
    const int ldp = (int)((nd + 7) & ~(size_t)7);
    auto up = [](size_t b) { return (b + 511) & ~(size_t)511; };
    const size_t need = up(nd * 8) * 2 + up((size_t)2 * grid * 8) + up((size_t)grid * ldp * 8) + up(512) + up(64) +
                        up((size_t)op->M * 8) + 4096;
    LZ_CHECK(arena_reserve(ctx, need));
    char* base = (char*)ctx->arena;
    size_t off = 0;
    auto take = [&](size_t b) { char* p = base + off; off += up(b); return p; };
    SmallArgs a{};
    a.nx = (int)st.nx; a.ny = (int)st.ny; a.nz = (int)st.nz;
    a.periodic = (st.bc == LZ_BC_PERIODIC);
    a.c = st.center; a.ox = st.offx; a.oy = st.offy; a.oz = st.offz;
    a.diag = st.diag;
    a.M = op->M;
    a.n = n;
    a.ref = opts->ref_compat ? 1 : 0;
    a.reorth_full = opts->reorth == LZ_REORTH_FULL;
    a.passes = passes;
    a.gpu_sweep = (opts->flags & 2) ? 1 : 0;
    a.tol_rel = opts->breakdown_tol;
    a.v0 = v0_dev;
    a.V = V_dev;
    a.ldv = ldv;
    a.alpha = (double*)take(nd * 8);
    a.beta = (double*)take(nd * 8);
    a.red = (double*)take((size_t)2 * grid * 8);
    a.dpart = (double*)take((size_t)grid * ldp * 8);
    a.ldp = ldp;
    a.bar = (unsigned int*)take(512);
    a.flags = (int*)take(64);
    a.tmp = (double*)take((size_t)op->M * 8);
    cudaStream_t q = ctx->stream;
    LZ_CUDA(cudaMemsetAsync(a.alpha, 0, (char*)a.red - (char*)a.alpha, q));
    LZ_CUDA(cudaMemsetAsync(a.bar, 0, 512, q));
    const int h_flags[8] = {-1, 0, 0, 0, 0, 0, 0, 0};
    LZ_CUDA(cudaMemcpyAsync(a.flags, h_flags, sizeof(h_flags), cudaMemcpyHostToDevice, q));
    const size_t smem = nd * 8;
    LZ_CUDA(cudaFuncSetAttribute((const void*)small_lanczos_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    LZ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)small_lanczos_kernel, kThreads, smem));
    LZ_REQUIRE(per_sm >= 1, "launch_small_solve: the kernel does not fit an SM");
    LZ_CUDA(cudaEventRecord(ctx->ev_begin, q));
    void* args[] = {(void*)&a};
    LZ_CUDA(cudaLaunchCooperativeKernel((const void*)small_lanczos_kernel, dim3(grid), dim3(kThreads), args, smem, q));
    LZ_CUDA(cudaEventRecord(ctx->ev_end, q));
    std::vector<double> h_a(nd), h_b(nd);
    int hf[8];
    LZ_CUDA(cudaMemcpyAsync(h_a.data(), a.alpha, nd * 8, cudaMemcpyDeviceToHost, q));
    LZ_CUDA(cudaMemcpyAsync(h_b.data(), a.beta, nd * 8, cudaMemcpyDeviceToHost, q));
    LZ_CUDA(cudaMemcpyAsync(hf, a.flags, sizeof(hf), cudaMemcpyDeviceToHost, q));
    LZ_CUDA(cudaStreamSynchronize(q));
    for (int j = 0; j < n; ++j) alpha_host[j] = h_a[(size_t)j];
    for (int k = 0; k + 1 < n; ++k) beta_host[k] = h_b[(size_t)k + 1];         // Lanczos.py:112 numbering
    if (row_scale_host) for (int j = 0; j < n; ++j) row_scale_host[j] = 1.0;   // rows are stored normalised
    float ms = 0.f;
    LZ_CUDA(cudaEventElapsedTime(&ms, ctx->ev_begin, ctx->ev_end));
    int status = LZ_OK, steps_done = n;
    if (hf[0] >= 0 && hf[0] < n) {
        steps_done = hf[0];
        set_error("Lanczos breakdown: beta[%d] = %.3e (Krylov space exhausted after %d steps)", hf[0], h_b[(size_t)hf[0]], steps_done);
        status = LZ_ERR_BREAKDOWN;
    }
    if (info) {
        memset(info, 0, sizeof(*info));
        info->steps_done = steps_done;
        info->reorth_count = a.reorth_full ? (a.ref ? n : n - 1) : 0;
        info->launches = 1;
        info->gpu_ms = ms;
        info->step_kernel = 4;                            // persistent small-problem kernel
    }
    return status;
}

}  // namespace lz
