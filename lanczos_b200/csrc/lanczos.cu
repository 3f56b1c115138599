// The Lanczos tridiagonalisation loop: host-side launch sequence plus the single-CTA
// "scalar" kernels that finish every reduction deterministically on the device and keep
// alpha, beta, the lazy normalisation factors and the Gram-Schmidt coefficients in HBM, so
// that the host never synchronises inside the loop.
//
// Replaces Lanczos.execute_Lanczos lines 100-119 (Python/Regular/Lanczos.py) and
// IrrLanczos.execute_LanczosOld lines 217-238 (Python/Irregular/IrrLanczos.py).
//
// Per step j (device kernels, one stream):
//   [K4a cgs_dots -> fin_ip -> K4b cgs_update] x passes   (full: every step; selective: predicated)
//   K1 apply_dot(row j, s_j) -> w, partials            fin_alpha -> alpha[j]
//   K3 update_norm(w, row j, row j-1) -> row j+1       fin_beta  -> beta[j+1], s_{j+1} = 1/beta
//   [omega recurrence -> flag for step j+1]             (selective only)
// Basis rows are stored un-normalised (row j = r_j, q_j = s_j * row j) unless a Gram-Schmidt
// sweep rewrote them (then s_j = 1): the plain step moves 48*M bytes.
#include <math.h>
#include <vector>
#include "internal.h"

namespace lz {

// Fixed-order sum of n partials by one CTA: thread t adds p[t], p[t+256], ... then the
// block tree.  Independent of everything but n => bit-reproducible run to run.
__device__ __forceinline__ double cta_sum_partials(const double* __restrict__ p, int n, double* red) {
    double a = 0.0;
    for (int i = threadIdx.x; i < n; i += kThreads) a += p[i];
    return block_sum(a, red);
}

struct RunState {          // all device pointers
    double* alpha;         // [n]
    double* beta;          // [n+1]  beta[j] = |r_j|  (r_j is what row j stores before any sweep)
    double* scale;         // [n+1]  q_j = scale[j] * row_j
    double* coef;          // [n+1]  Gram-Schmidt coefficients for K4b (already times scale[r])
    double* cself;         // [1]
    double* v0scale;       // [1]    1/|v0|
    double* alpha_pre;     // [1]    alpha of the pre-step (discarded by the reference)
    double* omega_a;       // [n+2]  selective monitor, omega_{j,k}
    double* omega_b;       // [n+2]
    double* anorm;         // [1]    running estimate of |H|
    int* flags;            // [0] first breakdown step (-1: none), [1] reorth flag of the coming step,
                           // [2] reorth count, [3] force-next flag
};

__global__ void __launch_bounds__(kThreads)
fin_v0norm_kernel(const double* __restrict__ partials, int np, RunState st) {
    __shared__ double red[kWarps];
    const double s = cta_sum_partials(partials, np, red);
    if (threadIdx.x == 0) {
        const double nrm = sqrt(s);
        st.v0scale[0] = (nrm > 0.0) ? 1.0 / nrm : 0.0;
        if (!(nrm > 0.0) && st.flags[0] < 0) st.flags[0] = 0;
    }
}

__global__ void __launch_bounds__(kThreads)
fin_alpha_kernel(const double* __restrict__ partials, int np, double* __restrict__ out) {
    __shared__ double red[kWarps];
    const double s = cta_sum_partials(partials, np, red);
    if (threadIdx.x == 0) out[0] = s;
}

// beta[jn] = sqrt(sum), scale[jn] = 1/beta; breakdown bookkeeping.
__global__ void __launch_bounds__(kThreads)
fin_beta_kernel(const double* __restrict__ partials, int np, RunState st, int jn, double tol_rel,
                const double* __restrict__ magnitude) {
    __shared__ double red[kWarps];
    const double s = cta_sum_partials(partials, np, red);
    if (threadIdx.x == 0) {
        const double b = sqrt(s);
        st.beta[jn] = b;
        const double thresh = tol_rel * fabs(magnitude[0]);
        const bool ok = isfinite(b) && (b > thresh) && (b > 0.0);
        st.scale[jn] = (isfinite(b) && b > 0.0) ? 1.0 / b : 0.0;
        if (!ok && st.flags[0] < 0) st.flags[0] = jn;
    }
}

// Copy for the non-ref start: beta[0] = |v0|, scale[0] = 1/|v0|.
__global__ void init_first_row_kernel(RunState st) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const double s = st.v0scale[0];
        st.scale[0] = s;
        st.beta[0] = (s > 0.0) ? 1.0 / s : 0.0;
    }
}

// Gram-Schmidt coefficients from the dots partials:
//   ip_r = scale[r] * scale[j] * sum_g part[r*ncg + g]            r < j
//   coef[r] = ip_r * scale[r]
//   cself = (ref_form ? 2 - scale[j]^2 * (row_j . row_j) : 1) * scale[j];  scale[j] <- 1
__global__ void __launch_bounds__(kThreads)
fin_ip_kernel(const double* __restrict__ part, int ncg, int j, int self_included, int ref_form,
              RunState st, const int* __restrict__ flag, int count) {
    if (flag && *flag == 0) return;
    const double sj = st.scale[j];
    for (int r = threadIdx.x; r < j; r += kThreads) {
        const double* p = part + (int64_t)r * ncg;
        double a = 0.0;
        for (int g = 0; g < ncg; ++g) a += p[g];
        const double sr = st.scale[r];
        st.coef[r] = (a * sr * sj) * sr;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double c = 1.0;
        if (self_included && ref_form) {
            const double* p = part + (int64_t)j * ncg;
            double a = 0.0;
            for (int g = 0; g < ncg; ++g) a += p[g];
            c = 2.0 - a * sj * sj;
        }
        st.cself[0] = c * sj;
        st.scale[j] = 1.0;           // K4b stores the row normalised
        if (count) st.flags[2] += 1;
    }
}

// Selective re-orthogonalisation monitor (Simon's omega recurrence in the PROPACK form).
// Called after step j finished (alpha[j], beta[j+1] known); estimates
// omega_{j+1,k} ~ q_{j+1} . q_k for k <= j, and raises flags[1] for step j+1 when the
// largest estimate exceeds `delta` (and for the step after it: vectors are re-orthogonalised
// in pairs).  om_cur = omega_{j,.}, om_prev = omega_{j-1,.}; result overwrites om_prev.
__global__ void __launch_bounds__(kThreads)
omega_kernel(RunState st, int j, double* om_cur, double* om_prev, double delta, double eps1, double psi) {
    __shared__ double red[kWarps];
    const double bj1 = st.beta[j + 1];
    const double bj = (j > 0) ? st.beta[j] : 0.0;
    const double aj = st.alpha[j];
    // the reorth of the coming step was decided by the previous call (flags[1]); if this step's
    // vector was itself re-orthogonalised, its omegas are at round-off level.
    if (st.flags[1]) {
        for (int k = threadIdx.x; k < j; k += kThreads) om_cur[k] = eps1;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double an = st.anorm[0];
        an = fmax(an, fabs(aj) + bj + bj1);
        st.anorm[0] = an;
    }
    __syncthreads();
    const double anorm = st.anorm[0];
    double mx = 0.0;
    for (int k = threadIdx.x; k < j; k += kThreads) {
        const double bk1 = st.beta[k + 1];
        const double bk = (k > 0) ? st.beta[k] : 0.0;
        const double ok1 = (k + 1 < j) ? om_cur[k + 1] : ((k + 1 == j) ? 1.0 : 0.0);
        double t = bk1 * ok1 + (st.alpha[k] - aj) * om_cur[k] - bj * om_prev[k];
        if (k > 0) t += bk * om_cur[k - 1];
        const double d = eps1 * (fabs(aj) + bj1 + fabs(st.alpha[k]) + bk1) + eps1 * anorm;
        t = (t + copysign(d, t)) / bj1;
        om_prev[k] = t;                      // becomes omega_{j+1,k} after the swap on the host side
        mx = fmax(mx, fabs(t));
    }
    // block max
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        double m = 0.0;
        for (int w = 0; w < kWarps; ++w) m = fmax(m, red[w]);
        om_prev[j] = psi;                    // omega_{j+1,j}
        om_prev[j + 1] = 1.0;
        int fire = 0;
        if (st.flags[3]) { fire = 1; st.flags[3] = 0; }          // second vector of a pair
        else if (m > delta) { fire = 1; st.flags[3] = 1; }
        st.flags[1] = fire;
    }
}

// Carve the run's device workspace out of the context's grow-only arena (no cudaMalloc /
// cudaFree - and hence no implicit device synchronisation - in steady state).
struct Carver {
    char* base;
    size_t off = 0;
    template <typename T> T* take(size_t count) {
        T* p = reinterpret_cast<T*>(base + off);
        off += (count * sizeof(T) + 511) & ~(size_t)511;
        return p;
    }
};

int arena_reserve(lz_ctx* ctx, size_t bytes) {
    if (ctx->arena_bytes >= bytes) return LZ_OK;
    if (ctx->arena) { LZ_CUDA(cudaStreamSynchronize(ctx->stream)); cudaFree(ctx->arena); ctx->arena = nullptr; ctx->arena_bytes = 0; }
    LZ_CUDA(cudaMalloc(&ctx->arena, bytes));
    ctx->arena_bytes = bytes;
    return LZ_OK;
}

// CUDA-event pairs around selected launches, for the per-kernel roofline of bench.py.  Events
// come from a pool owned by the context (created once, reused by every run).
struct KernelTimer {
    lz_ctx* ctx = nullptr;
    std::vector<int> kind;
    size_t used = 0;
    bool on = false;
    cudaEvent_t next() {
        if (used == ctx->event_pool.size()) {
            cudaEvent_t e = nullptr;
            cudaEventCreate(&e);
            ctx->event_pool.push_back(e);
        }
        return ctx->event_pool[used++];
    }
    void begin(int k) {
        if (!on) return;
        cudaEventRecord(next(), ctx->stream);
        kind.push_back(k);
    }
    void end() {
        if (!on) return;
        cudaEventRecord(next(), ctx->stream);
        kind.push_back(-1);
    }
    void collect(float* ms_by_kind, int* count_by_kind, int nk) {
        for (int k = 0; k < nk; ++k) { ms_by_kind[k] = 0.f; count_by_kind[k] = 0; }
        for (size_t i = 0; i + 1 < used; i += 2) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, ctx->event_pool[i], ctx->event_pool[i + 1]) == cudaSuccess &&
                kind[i] >= 0 && kind[i] < nk) {
                ms_by_kind[kind[i]] += ms;
                count_by_kind[kind[i]] += 1;
            }
        }
        used = 0;
        kind.clear();
    }
};

}  // namespace lz

using namespace lz;

extern "C" int lz_lanczos_run(lz_ctx* ctx, lz_op* op, const double* v0_dev, int32_t n,
                              const lz_run_opts* opts, double* alpha_host, double* beta_host,
                              double* V_dev, int64_t ldv, double* row_scale_host, lz_run_info* info) {
    LZ_REQUIRE(ctx && op && v0_dev && opts && alpha_host, "lz_lanczos_run: null argument");
    LZ_REQUIRE(op->ctx == ctx, "lz_lanczos_run: operator belongs to another context");
    const int64_t M = op->M;
    LZ_REQUIRE(n >= 1, "lz_lanczos_run: n must be >= 1");
    LZ_REQUIRE(n <= M, "n cannot be larger than M!");                 // Lanczos.py:76-77
    LZ_REQUIRE(!(opts->ref_compat && n < 2), "lz_lanczos_run: ref_compat needs n >= 2 (the reference raises IndexError for n == 1)");
    LZ_REQUIRE(n < 2 || beta_host, "lz_lanczos_run: beta_host is null");
    const int reorth = opts->reorth;
    LZ_REQUIRE(reorth >= LZ_REORTH_NONE && reorth <= LZ_REORTH_SELECTIVE, "lz_lanczos_run: bad reorth mode %d", reorth);
    const int passes = opts->cgs_passes <= 0 ? 1 : opts->cgs_passes;
    LZ_REQUIRE(passes <= 2, "lz_lanczos_run: cgs_passes must be 1 or 2");
    LZ_REQUIRE(reorth == LZ_REORTH_NONE || V_dev, "lz_lanczos_run: re-orthogonalisation needs the basis buffer V_dev");
    LZ_REQUIRE(!V_dev || ldv >= M, "lz_lanczos_run: ldv < M");
    LZ_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;

    // ---- workspace ---------------------------------------------------------------------------
    const size_t nd = (size_t)n + 2;
    const int64_t ld_int = (M + 63) & ~(int64_t)63;
    const size_t vec_bytes = (size_t)ld_int * 8 + 512;
    size_t need = 16 * 512 + 6 * (nd * 8 + 512) + vec_bytes;
    if (!V_dev) need += 3 * vec_bytes;
    if (reorth != LZ_REORTH_NONE) need += (size_t)(n + 1) * kMaxPartials * 8 + 512;
    LZ_CHECK(arena_reserve(ctx, need));
    Carver cv{(char*)ctx->arena};
    RunState st{};
    st.alpha = cv.take<double>(nd);
    st.beta = cv.take<double>(nd);
    st.scale = cv.take<double>(nd);
    st.coef = cv.take<double>(nd);
    st.omega_a = cv.take<double>(nd);
    st.omega_b = cv.take<double>(nd);
    st.cself = cv.take<double>(1);
    st.v0scale = cv.take<double>(1);
    st.alpha_pre = cv.take<double>(1);
    st.anorm = cv.take<double>(1);
    st.flags = cv.take<int>(8);
    double* w = cv.take<double>((size_t)ld_int);
    double* ring = V_dev ? nullptr : cv.take<double>((size_t)ld_int * 3);
    double* gs_part = (reorth != LZ_REORTH_NONE) ? cv.take<double>((size_t)(n + 1) * kMaxPartials) : nullptr;

    // alpha .. omega_b are contiguous 512-byte-rounded blocks: one memset clears them all
    LZ_CUDA(cudaMemsetAsync(st.alpha, 0, (char*)st.cself - (char*)st.alpha, s));
    LZ_CUDA(cudaMemsetAsync(st.anorm, 0, 8, s));
    {
        const double one = 1.0;                       // omega_{0,0} = 1
        LZ_CUDA(cudaMemcpyAsync(st.omega_a, &one, 8, cudaMemcpyHostToDevice, s));
        const int h_flags[8] = {-1, reorth == LZ_REORTH_FULL ? 1 : 0, 0, 0, 0, 0, 0, 0};
        LZ_CUDA(cudaMemcpyAsync(st.flags, h_flags, sizeof(h_flags), cudaMemcpyHostToDevice, s));
    }
    auto row = [&](int j) -> double* {
        return V_dev ? V_dev + (int64_t)j * ldv : ring + (int64_t)(j % 3) * ld_int;
    };
    double* part = ctx->partials;
    int launches = 0, np = 0;
    const int ref = opts->ref_compat ? 1 : 0;
    const double eps = 2.220446049250313e-16;
    const double delta = opts->select_tol > 0.0 ? opts->select_tol : sqrt(eps);
    const double eps1 = eps * 1.5;      // noise floor of the omega recurrence
    const double psi = eps * sqrt((double)M);
    KernelTimer kt;
    kt.ctx = ctx;
    kt.on = (opts->profile != 0);
    if (kt.on) {   // create the pool outside the timed loop
        while (ctx->event_pool.size() < (size_t)(2 * (2 + 2 * passes) * n + 8)) {
            cudaEvent_t e = nullptr;
            LZ_CUDA(cudaEventCreate(&e));
            ctx->event_pool.push_back(e);
        }
    }
    enum { K_APPLY = 0, K_UPDATE = 1, K_DOTS = 2, K_GSUPD = 3, K_NKINDS = 4 };

    LZ_CUDA(cudaEventRecord(ctx->ev_begin, s));

    // ---- start: |v0| ---------------------------------------------------------------------
    LZ_CHECK(launch_dot(ctx, v0_dev, v0_dev, M, part, &np)); ++launches;
    fin_v0norm_kernel<<<1, kThreads, 0, s>>>(part, np, st); ++launches;
    if (ref) {
        // pre-step (Lanczos.py:108-110): r = H q - (q.Hq) q with q = v0/|v0|; r becomes row 0
        int l2 = 0;
        LZ_CHECK(launch_apply_dot(op, v0_dev, st.v0scale, w, part, &np, &l2)); launches += l2;
        fin_alpha_kernel<<<1, kThreads, 0, s>>>(part, np, st.alpha_pre); ++launches;
        LZ_CHECK(launch_update_norm(ctx, w, v0_dev, nullptr, st.alpha_pre, st.v0scale, nullptr, nullptr,
                                    row(0), M, part, &np)); ++launches;
        fin_beta_kernel<<<1, kThreads, 0, s>>>(part, np, st, 0, opts->breakdown_tol, st.alpha_pre); ++launches;
    } else {
        LZ_CUDA(cudaMemcpyAsync(row(0), v0_dev, (size_t)M * 8, cudaMemcpyDeviceToDevice, s));
        init_first_row_kernel<<<1, 32, 0, s>>>(st); ++launches;
    }

    double* om_cur = st.omega_a;
    double* om_prev = st.omega_b;
    for (int j = 0; j < n; ++j) {
        double* rj = row(j);
        // ---- Gram-Schmidt sweeps of q_j against the rows before it ---------------------------
        const bool maybe_reorth = (reorth == LZ_REORTH_FULL) || (reorth == LZ_REORTH_SELECTIVE && j > 0);
        if (maybe_reorth && (j > 0 || ref)) {
            const int* flag = (reorth == LZ_REORTH_SELECTIVE) ? (st.flags + 1) : nullptr;
            for (int p = 0; p < passes; ++p) {
                const int ref_form = (ref && p == 0) ? 1 : 0;
                const int nrows = ref_form ? j + 1 : j;      // the reference's sum includes row j itself
                if (nrows == 0) continue;
                int ncg = 0;
                kt.begin(K_DOTS);
                LZ_CHECK(launch_cgs_dots(ctx, V_dev, ldv, nrows, rj, M, gs_part, &ncg, flag)); ++launches;
                kt.end();
                fin_ip_kernel<<<1, kThreads, 0, s>>>(gs_part, ncg, j, ref_form, ref_form, st, flag,
                                                     (p == 0) ? 1 : 0);
                ++launches;
                kt.begin(K_GSUPD);
                LZ_CHECK(launch_cgs_update(ctx, V_dev, ldv, j, rj, st.coef, st.cself, rj, M, flag)); ++launches;
                kt.end();
            }
        }
        // ---- w = H q_j, alpha_j = q_j . w ------------------------------------------------------
        int l2 = 0;
        kt.begin(K_APPLY);
        LZ_CHECK(launch_apply_dot(op, rj, st.scale + j, w, part, &np, &l2)); launches += l2;
        kt.end();
        fin_alpha_kernel<<<1, kThreads, 0, s>>>(part, np, st.alpha + j); ++launches;
        // ---- r = w - alpha_j q_j - beta_j q_{j-1}; beta_{j+1} = |r| --------------------------
        double* out = (j + 1 < n) ? row(j + 1) : w;
        kt.begin(K_UPDATE);
        LZ_CHECK(launch_update_norm(ctx, w, rj, j > 0 ? row(j - 1) : nullptr, st.alpha + j, st.scale + j,
                                    st.beta + j, j > 0 ? st.scale + j - 1 : nullptr, out, M, part, &np));
        ++launches;
        kt.end();
        fin_beta_kernel<<<1, kThreads, 0, s>>>(part, np, st, j + 1, opts->breakdown_tol, st.alpha); ++launches;
        if (reorth == LZ_REORTH_SELECTIVE && j + 1 < n) {
            omega_kernel<<<1, kThreads, 0, s>>>(st, j, om_cur, om_prev, delta, eps1, psi); ++launches;
            std::swap(om_cur, om_prev);
        }
    }
    LZ_CUDA(cudaGetLastError());
    LZ_CUDA(cudaEventRecord(ctx->ev_end, s));

    // ---- results ---------------------------------------------------------------------------
    // alpha, beta, scale are adjacent in the arena: one D2H copy
    const size_t blk = (nd * 8 + 511) & ~(size_t)511;
    std::vector<char> h_blk(3 * blk);
    int h_flags[8];
    LZ_CUDA(cudaMemcpyAsync(h_blk.data(), st.alpha, 3 * blk, cudaMemcpyDeviceToHost, s));
    LZ_CUDA(cudaMemcpyAsync(h_flags, st.flags, sizeof(h_flags), cudaMemcpyDeviceToHost, s));
    LZ_CUDA(cudaStreamSynchronize(s));
    const double* h_alpha = reinterpret_cast<const double*>(h_blk.data());
    const double* h_beta = reinterpret_cast<const double*>(h_blk.data() + blk);
    const double* h_scale = reinterpret_cast<const double*>(h_blk.data() + 2 * blk);
    for (int j = 0; j < n; ++j) alpha_host[j] = h_alpha[j];
    for (int k = 0; k + 1 < n; ++k) beta_host[k] = h_beta[k + 1];     // Lanczos.py:112 numbering
    if (row_scale_host)
        for (int j = 0; j < n; ++j) row_scale_host[j] = h_scale[j];
    float ms = 0.f;
    LZ_CUDA(cudaEventElapsedTime(&ms, ctx->ev_begin, ctx->ev_end));
    float kms[K_NKINDS];
    int kcnt[K_NKINDS];
    kt.collect(kms, kcnt, K_NKINDS);
    int steps_done = n;
    int status = LZ_OK;
    // a breakdown at index jn means row jn could not be normalised: steps 0..jn-1 are valid.
    // beta[n] (after the last step) is not part of the output and is ignored.
    if (h_flags[0] >= 0 && h_flags[0] < n) {
        steps_done = h_flags[0];
        set_error("Lanczos breakdown: beta[%d] = %.3e (Krylov space exhausted after %d steps)",
                  h_flags[0], h_beta[h_flags[0]], steps_done);
        status = LZ_ERR_BREAKDOWN;
    }
    if (info) {
        info->steps_done = steps_done;
        info->reorth_count = (reorth == LZ_REORTH_FULL) ? (ref ? n : n - 1) : h_flags[2];
        info->launches = launches;
        info->gpu_ms = ms;
        info->apply_ms = kms[K_APPLY];   info->apply_launches = kcnt[K_APPLY];
        info->update_ms = kms[K_UPDATE]; info->update_launches = kcnt[K_UPDATE];
        info->dots_ms = kms[K_DOTS];     info->dots_launches = kcnt[K_DOTS];
        info->gsupd_ms = kms[K_GSUPD];   info->gsupd_launches = kcnt[K_GSUPD];
    }
    return status;
}

extern "C" int lz_basis_normalize(lz_ctx* ctx, double* V_dev, int64_t ldv, int32_t n, int64_t M,
                                  const double* row_scale_host) {
    LZ_REQUIRE(ctx && V_dev && row_scale_host, "lz_basis_normalize: null argument");
    LZ_CUDA(cudaSetDevice(ctx->device));
    for (int j = 0; j < n; ++j) {
        if (row_scale_host[j] == 1.0) continue;
        LZ_CHECK(launch_scale(ctx, V_dev + (int64_t)j * ldv, M, row_scale_host[j]));
    }
    return LZ_OK;
}
