// Scalar bookkeeping of the Lanczos loop on the device, and the "tail" that folds it into the
// kernel that produced the partial sums.
//
// Every streaming kernel of the loop ends with one partial sum per CTA.  What follows - the
// fixed-order sum of the partials, the cross-rank sum over NVLink peer memory, alpha / beta / the
// lazy normalisation factor / the breakdown flag, and the omega recurrence of selective
// re-orthogonalisation - is a few hundred flops.  Run as kernels of their own (round 1:
// fin_scalar_kernel x 2 + omega_kernel per step) each costs a full grid drain plus a launch, ~10 %
// of a 512^3 step and most of a small one.  Here the last CTA of the producing kernel to finish
// does it ("last-CTA ticket": every CTA publishes its partial, fences, takes a ticket from a
// device counter; the CTA that draws gridDim.x - 1 knows all partials are visible).  The order
// of the sum does not depend on which CTA is last, so results stay bit-reproducible.
//
// replaces the scalar lines of Lanczos.py:112-113,118 (beta = norm(r), V[j] = r / beta,
// alpha = dot(V[j], r)).
#pragma once
#include "common.cuh"
#include "peer.cuh"

namespace lz {

struct RunState {          // all device pointers
    double* alpha;         // [n]
    double* beta;          // [n+1]  beta[j] = |r_j|  (r_j is what row j stores before any sweep)
    double* scale;         // [n+1]  q_j = scale[j] * row_j
    double* coef;          // [n+1]  Gram-Schmidt coefficients for K4b (already times scale[r])
    double* omega_a;       // [n+2]  selective monitor, omega_{j,k}
    double* omega_b;       // [n+2]
    double* cself;         // [1]
    double* v0scale;       // [1]    1/|v0|
    double* alpha_pre;     // [1]    alpha of the pre-step (discarded by the reference)
    double* anorm;         // [1]    running estimate of |H|
    int* flags;            // [0] first breakdown step (-1: none), [1] reorth flag of the coming step,
                           // [2] reorth count, [3] force-next flag, [4] peer timeout,
                           // [5 + (r & 1)] speculate: the interior part of H row_r may be applied before the monitor
                           //               has decided about row r (sparse row shards, lanczos.cu),
                           // [7] the interior part of the coming row must (still / again) be applied after the join
};

enum { FIN_NONE = -1, FIN_V0NORM = 0, FIN_ALPHA = 1, FIN_BETA = 2, FIN_ALPHA_S2 = 3 };

// What to do with the sum `s` of one reduction.
struct FinOp {
    int kind = FIN_NONE;
    int jn = 0;                        // FIN_BETA: index of beta / scale; FIN_ALPHA_S2: index of scale
    double* out = nullptr;             // FIN_ALPHA / FIN_ALPHA_S2: where alpha goes
    double tol_rel = 0.0;              // FIN_BETA: breakdown when beta <= tol_rel * |magnitude[0]|
    const double* magnitude = nullptr;
    // FIN_ALPHA_S2 fed by two kernels (KB + border kernel): partials of the first one
    const double* extra = nullptr;
    int nextra = 0;
    // selective re-orthogonalisation: the omega recurrence of step `omega_j` runs right after FIN_BETA
    int omega_j = -1;
    double* om_cur = nullptr;
    double* om_prev = nullptr;
    double delta = 0.0, eps1 = 0.0, psi = 0.0;
};

// The tail a producing kernel carries (kind == FIN_NONE: it only writes its partials).
struct FinTail {
    FinOp op;
    unsigned int* ticket = nullptr;    // device counter, zero between launches
    RunState st{};
    PeerComm pc;                       // world == 1: no exchange
    unsigned long long seq = 0;
    int mode = LZ_XCHG_FUSED;          // LZ_XCHG_PUSH: several shards driven by one process - the combine
                                       // phase runs as fin_scalar_kernel once every shard has pushed
};

// Tail of a kernel that produced Gram-Schmidt dot partials part[r * ncg + g] (K4a, K4c): the coefficients of
// the sweep of row j, see fin_ip_body.
struct IpTail {
    int on = 0;                        // 0: the kernel only writes its partials
    int j = 0, self_included = 0, ref_form = 0, count = 0;
    unsigned int* ticket = nullptr;
    RunState st{};
    PeerComm pc;
    unsigned long long seq = 0;
    int mode = LZ_XCHG_FUSED;
};

#ifdef __CUDACC__

// Fixed-order sum of n partials by one CTA: thread t adds p[t], p[t+256], ... then the
// block tree.  Independent of everything but n => bit-reproducible run to run.  The partials were
// written by other CTAs of the same grid (or an earlier one): read them past L1.
__device__ __forceinline__ double cta_sum_partials(const double* p, int n, double* red) {
    double a = 0.0;
    for (int i = threadIdx.x; i < n; i += kThreads) a += __ldcg(p + i);
    return block_sum(a, red);
}

// thread 0 only: the bookkeeping for the (globally summed) value s
__device__ __forceinline__ void fin_apply(const FinOp& f, const RunState& st, double s) {
    if (f.kind == FIN_V0NORM) {
        const double nrm = sqrt(s);
        st.v0scale[0] = (nrm > 0.0) ? 1.0 / nrm : 0.0;
        if (!(nrm > 0.0) && st.flags[0] < 0) st.flags[0] = 0;
    } else if (f.kind == FIN_ALPHA) {
        f.out[0] = s;
    } else if (f.kind == FIN_ALPHA_S2) {      // alpha_j = q_j.H q_j from the un-normalised r_j
        const double sc = st.scale[f.jn];
        f.out[0] = s * sc * sc;
    } else if (f.kind == FIN_BETA) {
        const double b = sqrt(s);
        st.beta[f.jn] = b;
        const double thresh = f.tol_rel * fabs(f.magnitude[0]);
        const bool ok = isfinite(b) && (b > thresh) && (b > 0.0);
        st.scale[f.jn] = (isfinite(b) && b > 0.0) ? 1.0 / b : 0.0;
        if (!ok && st.flags[0] < 0) st.flags[0] = f.jn;
    }
}

// Selective re-orthogonalisation monitor (Simon's omega recurrence in the PROPACK form).
// Called after step j finished (alpha[j], beta[j+1] known); estimates
// omega_{j+1,k} ~ q_{j+1} . q_k for k <= j, and raises flags[1] for step j+1 when the
// largest estimate exceeds `delta` (and for the step after it: vectors are re-orthogonalised
// in pairs).  om_cur = omega_{j,.}, om_prev = omega_{j-1,.}; result overwrites om_prev.
// Whole CTA (kThreads threads).
__device__ __forceinline__ void omega_body(const RunState& st, int j, double* om_cur, double* om_prev,
                                           double delta, double eps1, double psi, double* red) {
    __syncthreads();                         // alpha[j] / beta[j+1] written by thread 0 just before
    const double bj1 = st.beta[j + 1];
    const double bj = (j > 0) ? st.beta[j] : 0.0;
    const double aj = st.alpha[j];
    // if this step's vector was itself re-orthogonalised, its omegas are at round-off level
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
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    __syncthreads();
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
        // Overlapped apply of sparse row shards: the interior part of H row_{j+1} was started (or not) on the
        // strength of the guess made one step ago; it has to run after all if that guess was "do not start" or
        // if row j+1 is about to be swept.  The guess for row j+2: start early unless a sweep is due or the
        // estimate is within a factor 64 of the threshold (it grows by a few |H|/beta per step at most).
        st.flags[7] = (fire || !st.flags[5 + ((j + 1) & 1)]) ? 1 : 0;
        st.flags[5 + (j & 1)] = (!fire && m * 64.0 < delta) ? 1 : 0;          // (j + 2) & 1 == j & 1
    }
}

// One scalar: CTA partials -> local sum -> (sharded) sum over ranks -> bookkeeping.  Whole CTA of
// kThreads threads; `red` holds kWarps doubles.
__device__ __forceinline__ void fin_scalar_body(const FinOp& f, const RunState& st, const PeerComm& pc,
                                                unsigned long long seq, int mode, const double* partials,
                                                int np, double* red) {
    double s = 0.0;
    if (mode != LZ_XCHG_COMBINE) {
        s = cta_sum_partials(partials, np, red);
        if (f.nextra > 0) {
            const double e = cta_sum_partials(f.extra, f.nextra, red);
            s += e;                              // thread 0 holds both sums
        }
    }
    if (pc.world > 1) {
        if (mode != LZ_XCHG_COMBINE) {
            if (threadIdx.x == 0) peer_store(pc, seq, 0, s);
            peer_publish(pc, seq);
            if (mode == LZ_XCHG_PUSH) return;
        }
        peer_wait(pc, seq);
        if (threadIdx.x == 0) s = peer_sum(pc, seq, 0);
    }
    if (threadIdx.x == 0) fin_apply(f, st, s);
    if (f.kind == FIN_BETA && f.omega_j >= 0)
        omega_body(st, f.omega_j, f.om_cur, f.om_prev, f.delta, f.eps1, f.psi, red);
}

// Gram-Schmidt coefficients from the dots partials (whole CTA, any block size):
//   ip_r = scale[r] * scale[j] * sum_ranks sum_g part[r*ncg + g]            r < j
//   coef[r] = ip_r * scale[r]
//   cself = (ref_form ? 2 - scale[j]^2 * (row_j . row_j) : 1) * scale[j];  scale[j] <- 1
__device__ __forceinline__ void fin_ip_body(const double* part, int ncg, int j, int self_included, int ref_form,
                                            const RunState& st, const PeerComm& pc, unsigned long long seq, int mode,
                                            int count) {
    const int nrows = self_included ? j + 1 : j;
    const bool sharded = pc.world > 1;
    const int nthr = blockDim.x;
    __shared__ double s_self;
    if (threadIdx.x == 0) s_self = 0.0;
    __syncthreads();
    if (mode != LZ_XCHG_COMBINE) {
        const double sj0 = st.scale[j];
        for (int r = threadIdx.x; r < nrows; r += nthr) {
            const double* p = part + (int64_t)r * ncg;
            double a = 0.0;
            for (int g = 0; g < ncg; ++g) a += __ldcg(p + g);
            if (sharded) peer_store(pc, seq, r, a);
            else if (r < j) { const double sr = st.scale[r]; st.coef[r] = (a * sr * sj0) * sr; }
            else s_self = a;
        }
        if (sharded) {
            peer_publish(pc, seq);
            if (mode == LZ_XCHG_PUSH) return;
        }
    }
    if (sharded) {
        peer_wait(pc, seq);
        const double sj0 = st.scale[j];
        for (int r = threadIdx.x; r < nrows; r += nthr) {
            const double a = peer_sum(pc, seq, r);
            if (r < j) { const double sr = st.scale[r]; st.coef[r] = (a * sr * sj0) * sr; }
            else s_self = a;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const double sj = st.scale[j];
        double c = 1.0;
        if (self_included && ref_form) c = 2.0 - s_self * sj * sj;
        st.cself[0] = c * sj;
        st.scale[j] = 1.0;           // K4b stores the row normalised
        if (count) st.flags[2] += 1;
    }
}

// Tail of a kernel that wrote dot partials: the last CTA to finish turns them into coefficients.  Call with
// the whole CTA after its partials have been stored.
__device__ __forceinline__ void ip_tail(const IpTail& t, const double* part, int ncg) {
    if (!t.on) return;
    __shared__ int s_last_ip;
    __syncthreads();                                       // every thread's partial stores are issued
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int k = atomicAdd(t.ticket, 1u);
        s_last_ip = (k == gridDim.x - 1);
        if (s_last_ip) *t.ticket = 0;
    }
    __syncthreads();
    if (!s_last_ip) return;
    __threadfence();
    fin_ip_body(part, ncg, t.j, t.self_included, t.ref_form, t.st, t.pc, t.seq, t.mode, t.count);
}

// Tail of a producing kernel.  Call with the whole CTA after thread 0 has stored
// partials[blockIdx.x]; the CTA that finishes last runs the bookkeeping.  `red`: kWarps doubles.
__device__ __forceinline__ void fin_tail(const FinTail& t, const double* partials, double* red) {
    if (t.op.kind == FIN_NONE) return;
    __shared__ int s_last;
    if (threadIdx.x == 0) {
        __threadfence();                                   // my partial is visible before my ticket
        const unsigned int k = atomicAdd(t.ticket, 1u);
        s_last = (k == gridDim.x - 1);
        if (s_last) *t.ticket = 0;                         // ready for the next launch
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    fin_scalar_body(t.op, t.st, t.pc, t.seq, t.mode, partials, (int)gridDim.x, red);
}

#endif  // __CUDACC__

}  // namespace lz
