// The Lanczos tridiagonalisation loop: the host-side launch sequence.  Every reduction is finished
// deterministically on the device - across CTAs and, in a row-sharded run, across GPUs through NVLink
// peer memory - by the tail of the kernel that produced it (fin.cuh); alpha, beta, the lazy
// normalisation factors and the Gram-Schmidt coefficients stay in HBM, so that the host never
// synchronises inside the loop.
//
// Replaces Lanczos.execute_Lanczos lines 100-119 (Python/Regular/Lanczos.py) and
// IrrLanczos.execute_LanczosOld lines 217-238 (Python/Irregular/IrrLanczos.py).
//
// Per step j (device kernels, one stream per shard):
//   [K4a cgs_dots -> fin_ip -> K4b cgs_update] x passes   (full: every step; selective: predicated)
//   matrix-free operators:  KA2(row j) -> alpha[j] in its tail;  KB(row j, row j-1) -> row j+1, beta[j+1],
//                           s_{j+1} = 1/beta and the omega recurrence in its tail              (32*M bytes)
//   stored operators:       K2 apply_dot(row j) -> w, alpha[j];  K3 update_norm(w, row j, row j-1) -> row j+1,
//                           beta[j+1], ...                                  (48*M bytes + the operator's own)
// Basis rows are stored un-normalised (row j = r_j, q_j = s_j * row j) unless a Gram-Schmidt
// sweep rewrote them (then s_j = 1).
//
// Sharded runs (lz_team): every vector and basis row is split into contiguous row blocks, one
// per GPU (z-slabs of a structured grid).  KB / K3 / K4b store the boundary planes of the vector they
// produce straight into the neighbours' ghost buffers; the tails push their partial sums
// to every peer and add the P contributions in rank order (peer.cuh) - that flag also
// publishes the halo.  No NCCL call, no extra pass over HBM, no extra launch per step.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "internal.h"
#include "peer.cuh"

namespace lz {

// One scalar: CTA partials -> local sum -> (sharded) sum over ranks -> bookkeeping.  The stand-alone
// form of the fin tail (fin.cuh): the combine phase when one process drives several shards, and the
// opt-in single-pass step.
__global__ void __launch_bounds__(kThreads)
fin_scalar_kernel(const double* __restrict__ partials, int np, const FinOp f, const RunState st, const PeerComm pc,
                  unsigned long long seq, int mode, const int* __restrict__ flag) {
    pdl_prologue();
    if (flag && *flag == 0) return;
    __shared__ double red[kWarps];
    fin_scalar_body(f, st, pc, seq, mode, partials, np, red);
}

// Fin of the single-pass fused step: partials[0..g) = sum r^2, partials[g..2g) = sum r.u with
// r = r_{jn}, u = H r_{jn}:  beta[jn] = |r|, scale[jn] = 1/beta, alpha[jn] = scale^2 * r.u
__global__ void __launch_bounds__(kThreads)
fin_fused_kernel(const double* __restrict__ partials, int g, RunState st, int jn, double tol_rel,
                 const double* __restrict__ magnitude) {
    __shared__ double red[kWarps];
    const double srr = cta_sum_partials(partials, g, red);
    const double sru = cta_sum_partials(partials + g, g, red);
    if (threadIdx.x != 0) return;
    const double b = sqrt(srr);
    st.beta[jn] = b;
    const double thresh = tol_rel * fabs(magnitude[0]);
    const bool ok = isfinite(b) && (b > thresh) && (b > 0.0);
    const double sc = (isfinite(b) && b > 0.0) ? 1.0 / b : 0.0;
    st.scale[jn] = sc;
    st.alpha[jn] = sru * sc * sc;
    if (!ok && st.flags[0] < 0) st.flags[0] = jn;
}

// Non-ref start: beta[0] = |v0|, scale[0] = 1/|v0|.
__global__ void init_first_row_kernel(RunState st) {
    pdl_prologue();
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const double s = st.v0scale[0];
        st.scale[0] = s;
        st.beta[0] = (s > 0.0) ? 1.0 / s : 0.0;
    }
}

// Gram-Schmidt coefficients from the dots partials as a kernel of its own (fin_ip_body, fin.cuh): the
// combine phase when one process drives several shards, and the opt-in single-pass step; elsewhere it is
// the tail of the kernel that produced the partials.
__global__ void __launch_bounds__(kThreads)
fin_ip_kernel(const double* __restrict__ part, int ncg, int j, int self_included, int ref_form,
              RunState st, PeerComm pc, unsigned long long seq, int mode,
              const int* __restrict__ flag, int count) {
    pdl_prologue();
    if (flag && *flag == 0) return;
    fin_ip_body(part, ncg, j, self_included, ref_form, st, pc, seq, mode, count);
}

// Flag-only exchange: publishes the halo planes that the preceding kernel of this stream stored
// into the neighbours' ghost buffers, and waits for theirs.
__global__ void __launch_bounds__(32)
peer_sync_kernel(PeerComm pc, unsigned long long seq, int mode, const int* __restrict__ flag) {
    pdl_prologue();
    if (flag && *flag == 0) return;
    if (pc.world <= 1) return;
    if (mode != LZ_XCHG_COMBINE) {
        peer_publish(pc, seq);
        if (mode == LZ_XCHG_PUSH) return;
    }
    peer_wait(pc, seq);
}

// the omega recurrence as a kernel of its own (single-pass fused step only; elsewhere it rides in
// the tail of the kernel that produced beta)
__global__ void __launch_bounds__(kThreads)
omega_kernel(RunState st, int j, double* om_cur, double* om_prev, double delta, double eps1, double psi) {
    __shared__ double red[kWarps];
    omega_body(st, j, om_cur, om_prev, delta, eps1, psi, red);
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
    if (ctx->arena) {
        LZ_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->arena);
        ctx->arena = nullptr;
        ctx->arena_bytes = 0;
    }
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

// One row shard of a distributed solve, as this process sees it.
struct lz_shard {
    lz_ctx* ctx = nullptr;
    int rank = 0;
    PeerComm pc;
    void* comm[kMaxWorld] = {};          // exchange buffers of all ranks, mapped in this process
    // structured-grid halo: where my boundary planes go / where my neighbours' planes arrive
    int lower = -1, upper = -1;          // neighbour ranks (-1: none, domain boundary)
    int64_t plane = 0;
    // sparse row shard: ghost-index exchange (what I send to whom, where it lands)
    int32_t* send_idx = nullptr;         // device, nsend local row indices, grouped by destination rank
    int nsend = 0;
    int seg_start[kMaxWorld + 1] = {};
    int64_t dst_off[kMaxWorld] = {};
};

struct lz_team {
    int world = 1;
    int nlocal = 1;
    int kmax = 0;
    int64_t plane = 0, nghost = 0;
    int64_t global_rows = 0;             // M of the whole operator (same number on every rank)
    CommLayout layout{};
    std::vector<lz_shard> shards;
    unsigned long long seq = 0;          // sequence number of the last cross-rank exchange
    bool poisoned = false;               // a run failed part-way: the ranks' sequence numbers may differ for good
};

namespace {

struct ShardRun {           // per-shard state of one run
    lz_ctx* ctx = nullptr;
    lz_op* op = nullptr;
    lz_shard* sh = nullptr; // null: single shard, no exchange
    RunState st{};
    PeerComm pc;
    const double* v0 = nullptr;
    double* V = nullptr;
    int64_t ldv = 0;
    int64_t M = 0;
    double* w = nullptr;
    double* w2 = nullptr;   // second work vector of the fused single-pass step
    double* ring = nullptr;
    int64_t ld_int = 0;
    double* gs_part = nullptr;
    double* om_cur = nullptr;
    double* om_prev = nullptr;
    KernelTimer kt;
    int np = 0;
    int np_beta = 0;        // sparse row shards: CTAs of the last update kernel (its partials: first half of ctx->partials;
                            // the apply kernels of that mode write theirs to the second half)
    int np_kb = 0;          // CTAs of the last KB launch (its alpha partials sit in the second half of ctx->partials)
    int ncg = 0;
    double* row(int j) const { return V ? V + (int64_t)j * ldv : ring + (int64_t)(j % 3) * ld_int; }
};

HaloPush halo_for(const lz_team* team, const ShardRun& r, int parity) {
    HaloPush h{};
    if (!team || !r.sh || r.sh->plane == 0) return h;
    const lz_shard* sh = r.sh;
    h.plane = sh->plane;
    const size_t pb = (size_t)sh->plane * 8;
    // my first plane is the plane ABOVE the lower neighbour's slab -> its ghost_hi
    if (sh->lower >= 0)
        h.lo_dst = reinterpret_cast<double*>((char*)sh->comm[sh->lower] + team->layout.ghost_hi_off + parity * pb);
    if (sh->upper >= 0)
        h.hi_dst = reinterpret_cast<double*>((char*)sh->comm[sh->upper] + team->layout.ghost_lo_off + parity * pb);
    return h;
}

void bind_ghosts(const lz_team* team, ShardRun& r, int parity) {
    if (!team || !r.sh || r.sh->plane == 0 || r.op->kind != LZ_OP_STENCIL) return;
    const lz_shard* sh = r.sh;
    const size_t pb = (size_t)sh->plane * 8;
    char* mine = (char*)sh->comm[sh->rank];
    r.op->st.sharded = 1;
    r.op->st.ghost_lo = (sh->lower >= 0) ? reinterpret_cast<const double*>(mine + team->layout.ghost_lo_off + parity * pb) : nullptr;
    r.op->st.ghost_hi = (sh->upper >= 0) ? reinterpret_cast<const double*>(mine + team->layout.ghost_hi_off + parity * pb) : nullptr;
}

// sparse row shard: ghost entries of x arrive in my gather buffer of the given parity
void bind_gather(const lz_team* team, ShardRun& r, int parity) {
    if (!team || !r.sh || r.op->kind == LZ_OP_STENCIL) return;
    if (team->nghost == 0 || r.op->ncols <= r.op->M) { r.op->xghost = nullptr; return; }
    char* mine = (char*)r.sh->comm[r.sh->rank];
    r.op->xghost = reinterpret_cast<const double*>(mine + team->layout.gather_off) + (size_t)parity * team->nghost;
}

// sparse row shard: send the entries of `x` that other ranks need (gather + NVLink peer stores)
int push_ghosts(const lz_team* team, ShardRun& r, const double* x, int parity, const int* flag, int* launches,
                cudaStream_t stream = nullptr) {
    if (!team || !r.sh || r.op->kind == LZ_OP_STENCIL || r.sh->nsend == 0) return LZ_OK;
    double* dst[kMaxWorld];
    for (int q = 0; q < team->world; ++q)
        dst[q] = reinterpret_cast<double*>((char*)r.sh->comm[q] + team->layout.gather_off) +
                 (size_t)parity * team->nghost + r.sh->dst_off[q];
    ++*launches;
    return launch_ghost_push(r.ctx, x, r.sh->send_idx, r.sh->nsend, team->world, r.sh->seg_start, dst, flag, stream);
}

// ---- cache of captured solves (see run_loop) -----------------------------------------------------------
struct GraphKey {
    unsigned long long op_serial;
    const void* op; const void* v0; const void* V; const void* arena; const void* diag;
    int64_t ldv, M;
    int32_t n;
    lz_run_opts opts;
};
}  // namespace
struct lz_graph_slot {
    GraphKey key;
    cudaGraphExec_t exec = nullptr;
    int seen = 0;
    int launches = 0;
    unsigned long long stamp = 0;
};
namespace {
using GraphSlot = lz_graph_slot;

bool same_key(const GraphKey& a, const GraphKey& b) {
    return a.op_serial == b.op_serial && a.op == b.op && a.v0 == b.v0 && a.V == b.V && a.arena == b.arena && a.diag == b.diag && a.ldv == b.ldv &&
           a.M == b.M && a.n == b.n && memcmp(&a.opts, &b.opts, sizeof(lz_run_opts)) == 0;
}

// the slot for `key` in the context's small cache (least recently used slot recycled)
GraphSlot* graph_lookup(lz_ctx* ctx, const GraphKey& key) {
    static unsigned long long clock = 0;
    constexpr size_t kSlots = 4;
    auto& slots = ctx->graphs;
    for (GraphSlot* s : slots)
        if (same_key(s->key, key)) { s->stamp = ++clock; return s; }
    GraphSlot* s = nullptr;
    if (slots.size() < kSlots) {
        s = new GraphSlot();
        slots.push_back(s);
    } else {
        s = slots[0];
        for (GraphSlot* t : slots) if (t->stamp < s->stamp) s = t;
        if (s->exec) cudaGraphExecDestroy(s->exec);
        *s = GraphSlot();
    }
    s->key = key;
    s->stamp = ++clock;
    return s;
}

// The loop over `nl` local shards (nl == 1 and team == nullptr: the plain single-GPU solve).
int run_loop(lz_team* team, int nl, lz_op* const* ops, const double* const* v0s, int32_t n,
             const lz_run_opts* opts, double* alpha_host, double* beta_host, double* const* Vs,
             const int64_t* ldvs, double* row_scale_host, lz_run_info* info) {
    LZ_REQUIRE(ops && v0s && opts && alpha_host, "lz_lanczos_run: null argument");
    LZ_REQUIRE(n >= 1, "lz_lanczos_run: n must be >= 1");
    LZ_REQUIRE(!(opts->ref_compat && n < 2), "lz_lanczos_run: ref_compat needs n >= 2 (the reference raises IndexError for n == 1)");
    LZ_REQUIRE(n < 2 || beta_host, "lz_lanczos_run: beta_host is null");
    const int reorth = opts->reorth;
    LZ_REQUIRE(reorth >= LZ_REORTH_NONE && reorth <= LZ_REORTH_SELECTIVE, "lz_lanczos_run: bad reorth mode %d", reorth);
    const int passes = opts->cgs_passes <= 0 ? 1 : opts->cgs_passes;
    LZ_REQUIRE(passes <= 2, "lz_lanczos_run: cgs_passes must be 1 or 2");
    const int world = team ? team->world : 1;
    LZ_REQUIRE(!team || n + 2 <= team->kmax, "lz_team_lanczos_run: n = %d exceeds the team's max_steps", n);
    const bool split = nl > 1;            // several shards driven by this process: push / combine phases
    const size_t nd = (size_t)n + 2;
    // single-pass fused step KF: structured grid, one shard, steps that (normally) need no sweep
    // (opt-in: measured on par with the two-pass step on B200, see fused.cu)
    const bool fused = !team && nl == 1 && ops[0] && opts->step_kernel == 2 && reorth != LZ_REORTH_FULL &&
                       n >= 2 && fused_step_supported(ops[0]);
    // "recompute" step KA + KB (32*M B): matrix-free operators only; the default for them
    bool recompute = (opts->step_kernel == 0 || opts->step_kernel == 3);
    for (int s = 0; s < nl && recompute; ++s)
        recompute = ops[s] && (opts->step_kernel == 3 ? recompute_step_supported(ops[s]) : recompute_step_preferred(ops[s]));
    if (opts->step_kernel == 3 && !recompute) {
        set_error("lz_lanczos_run: the recompute step needs matrix-free (stencil) operators");
        return LZ_ERR_UNSUPPORTED;
    }
    LZ_REQUIRE(opts->step_kernel >= 0 && opts->step_kernel <= 3, "lz_lanczos_run: unknown step_kernel %d", opts->step_kernel);
    if (opts->step_kernel == 2 && !fused) {
        set_error("lz_lanczos_run: the fused single-pass step needs a 3-D structured grid with nx %% 64 == 0, "
                  "ny %% 8 == 0 on one GPU and reorth != full");
        return LZ_ERR_UNSUPPORTED;
    }

    std::vector<ShardRun> R(nl);
    int64_t M_local = 0;
    for (int s = 0; s < nl; ++s) {
        ShardRun& r = R[s];
        r.op = ops[s];
        LZ_REQUIRE(r.op && v0s[s], "lz_lanczos_run: null operator or start vector");
        r.ctx = r.op->ctx;
        r.sh = team ? &team->shards[s] : nullptr;
        LZ_REQUIRE(!team || r.sh->ctx == r.ctx, "lz_team_lanczos_run: operator %d is not on its shard's context", s);
        r.v0 = v0s[s];
        r.V = Vs ? Vs[s] : nullptr;
        r.ldv = ldvs ? ldvs[s] : 0;
        r.M = r.op->M;
        M_local += r.M;
        LZ_REQUIRE(reorth == LZ_REORTH_NONE || r.V, "lz_lanczos_run: re-orthogonalisation needs the basis buffer V_dev");
        LZ_REQUIRE(!r.V || r.ldv >= r.M, "lz_lanczos_run: ldv < M");
        if (team) r.pc = r.sh->pc;
    }
    if (!team) LZ_REQUIRE(n <= M_local, "n cannot be larger than M!");             // Lanczos.py:76-77
    const double M_global = team ? (double)team->global_rows : (double)M_local;
    LZ_REQUIRE(!team || n <= team->global_rows, "n cannot be larger than M!");
    (void)world;

    // ---- workspace ---------------------------------------------------------------------------
    for (int s = 0; s < nl; ++s) {
        ShardRun& r = R[s];
        LZ_CUDA(cudaSetDevice(r.ctx->device));
        r.ld_int = (r.M + 63) & ~(int64_t)63;
        const size_t vec_bytes = (size_t)r.ld_int * 8 + 512;
        size_t need = 16 * 512 + 6 * (nd * 8 + 512) + vec_bytes;
        if (fused) need += vec_bytes;
        if (!r.V) need += 3 * vec_bytes;
        if (reorth != LZ_REORTH_NONE) need += (size_t)(n + 1) * kMaxPartials * 8 + 512;
        LZ_CHECK(arena_reserve(r.ctx, need));
        Carver cv{(char*)r.ctx->arena};
        RunState& st = r.st;
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
        r.w = cv.take<double>((size_t)r.ld_int);
        r.w2 = fused ? cv.take<double>((size_t)r.ld_int) : nullptr;
        r.ring = r.V ? nullptr : cv.take<double>((size_t)r.ld_int * 3);
        r.gs_part = (reorth != LZ_REORTH_NONE) ? cv.take<double>((size_t)(n + 1) * kMaxPartials) : nullptr;
        r.om_cur = st.omega_a;
        r.om_prev = st.omega_b;
        r.pc.err = st.flags + 4;
        cudaStream_t q = r.ctx->stream;
        // alpha .. omega_b are contiguous 512-byte-rounded blocks: one memset clears them all
        LZ_CUDA(cudaMemsetAsync(st.alpha, 0, (char*)st.cself - (char*)st.alpha, q));
        LZ_CUDA(cudaMemsetAsync(st.anorm, 0, 8, q));
        const double one = 1.0;                       // omega_{0,0} = 1
        LZ_CUDA(cudaMemcpyAsync(st.omega_a, &one, 8, cudaMemcpyHostToDevice, q));
        const int h_flags[8] = {-1, reorth == LZ_REORTH_FULL ? 1 : 0, 0, 0, 0, 0, 0, 1};   // [5],[6]: no speculation yet, [7]: apply late
        LZ_CUDA(cudaMemcpyAsync(st.flags, h_flags, sizeof(h_flags), cudaMemcpyHostToDevice, q));
        r.kt.ctx = r.ctx;
        r.kt.on = (opts->profile != 0);
        if (r.kt.on) {   // create the pool outside the timed loop
            while (r.ctx->event_pool.size() < (size_t)(2 * (5 + 2 * passes) * n + 16)) {
                cudaEvent_t e = nullptr;
                LZ_CUDA(cudaEventCreate(&e));
                r.ctx->event_pool.push_back(e);
            }
        }
    }

    int launches = 0;
    const int ref = opts->ref_compat ? 1 : 0;
    const double eps = 2.220446049250313e-16;
    const double delta = opts->select_tol > 0.0 ? opts->select_tol : sqrt(eps);
    const double eps1 = eps * 1.5;      // noise floor of the omega recurrence
    const double psi = eps * sqrt(M_global);
    enum { K_APPLY = 0, K_UPDATE = 1, K_DOTS = 2, K_GSUPD = 3, K_FUSED = 4, K_GSFUSED = 5, K_BORDER = 6, K_NKINDS = 7 };
    const bool allow_gs_fusion = passes == 2 && !(opts->flags & 1);
    const bool gpu_sweep = (opts->flags & 2) != 0;       // Regular/Lanczos.py:236-238: self term dropped
    const bool sel = (reorth == LZ_REORTH_SELECTIVE);
    // alpha of the NEXT vector accumulated inside KB while that vector is in registers, plus a small border
    // kernel for the edges that cross CTA tiles (stencil.cu): replaces KA2's full pass over the vector.  Not
    // with full re-orthogonalisation (the sweep rewrites the vector, alpha must be taken afterwards anyway).
    // KBA (kba.cu): one kernel per step - KB plus the alpha reduction of the vector it writes, read back through
    // L2 a few z-planes behind: 24*M bytes of HBM per step instead of 32*M.  One GPU, whole tiles.
    bool kba = recompute && !team && nl == 1 && reorth != LZ_REORTH_FULL && (opts->flags & 16);
    if (kba) {
        const double* any_row = R[0].V ? R[0].V : R[0].ring;
        kba = kba_step_supported(R[0].op, any_row, any_row, any_row) && ((reinterpret_cast<uintptr_t>(R[0].v0) & 15) == 0) &&
              (((R[0].V ? R[0].ldv : R[0].ld_int) & 1) == 0);
    }
    bool kb_alpha = !kba && recompute && reorth != LZ_REORTH_FULL && (opts->flags & 4);
    for (int s = 0; s < nl && kb_alpha; ++s)
        kb_alpha = update_alpha_supported(R[s].op, R[s].V ? R[s].V : R[s].ring, R[s].V ? R[s].V : R[s].ring);

    // Row shards of a stored (sparse) operator.  (1) The ghost entries of a new vector are pushed by a kernel
    // of their own AFTER the update kernel, and the flag of the beta exchange is what publishes them - so that
    // exchange runs as a kernel after the push, not in the update kernel's tail.  (2) H is applied to the
    // un-normalised row and 1/beta is folded in afterwards (alpha = s^2 r.Hr; K3 scales w on load), so the
    // apply needs nothing from the beta exchange.  (3) Overlap: the interior spans of the next apply (no ghost
    // column) run on the main stream while push + exchange run on a second, high-priority stream; only the
    // boundary spans wait for the neighbours.  Not with full re-orthogonalisation (the sweep rewrites the row
    // the early apply would read; Gram-Schmidt dominates there anyway).
    const bool sparse_team = team && !recompute && !fused && ops[0]->kind != LZ_OP_STENCIL;
    bool overlap = sparse_team && reorth != LZ_REORTH_FULL && !(opts->flags & 8);
    for (int s = 0; s < nl && overlap; ++s) overlap = spmv_split_supported(R[s].op);
    if (overlap) {
        for (int s = 0; s < nl; ++s) {
            lz_ctx* c = R[s].ctx;
            LZ_CUDA(cudaSetDevice(c->device));
            if (!c->stream2) {
                int lo = 0, hi = 0;
                LZ_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
                LZ_CUDA(cudaStreamCreateWithPriority(&c->stream2, cudaStreamNonBlocking, hi));
                LZ_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
                LZ_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
            }
        }
    }
    bool interior_done = false;          // the interior spans of H row_j were applied during the previous step

    // run `fn(shard)` on every local shard, on its device
    auto each = [&](auto&& fn) -> int {
        for (int s = 0; s < nl; ++s) {
            if (nl > 1) LZ_CUDA(cudaSetDevice(R[s].ctx->device));
            LZ_CHECK(fn(R[s]));
        }
        return LZ_OK;
    };
    // a cross-rank exchange: one fused kernel per shard, or (several local shards) a push phase
    // over all shards followed by a combine phase
    auto exchange = [&](auto&& launch) -> int {
        unsigned long long seq = 0;
        if (team) seq = ++team->seq;
        if (!split) return each([&](ShardRun& r) { return launch(r, seq, (int)LZ_XCHG_FUSED); });
        LZ_CHECK(each([&](ShardRun& r) { return launch(r, seq, (int)LZ_XCHG_PUSH); }));
        return each([&](ShardRun& r) { return launch(r, seq, (int)LZ_XCHG_COMBINE); });
    };
    // A kernel that ends in a reduction, with the scalar bookkeeping `op_of(shard)` folded into its tail
    // (fin.cuh): produce(shard, tail) launches it.  One shard per process: the tail also does the cross-rank
    // sum.  Several local shards: the tails push, then one combine kernel per shard.  `pred`: the kernel
    // (and its combine) is predicated on the re-orthogonalisation flag of the step.
    auto produce_fin = [&](auto&& op_of, auto&& produce, bool pred, int part_off = 0) -> int {
        unsigned long long seq = 0;
        if (team) seq = ++team->seq;
        LZ_CHECK(each([&](ShardRun& r) {
            FinTail t;
            t.op = op_of(r);
            t.ticket = r.ctx->tickets;
            t.st = r.st;
            t.pc = r.pc;
            t.seq = seq;
            t.mode = split ? (int)LZ_XCHG_PUSH : (int)LZ_XCHG_FUSED;
            return produce(r, &t);
        }));
        if (!split) return LZ_OK;
        return each([&](ShardRun& r) {
            LZ_CUDA(launch_k(fin_scalar_kernel, dim3(1), dim3(kThreads), 0, r.ctx->stream,
                             (const double*)(r.ctx->partials + part_off), r.np, op_of(r), r.st, r.pc, seq, (int)LZ_XCHG_COMBINE,
                             pred ? (const int*)(r.st.flags + 1) : (const int*)nullptr));
            ++launches;
            return LZ_OK;
        });
    };
    auto peer_sync = [&](bool predicated) -> int {
        if (!team) return LZ_OK;
        return exchange([&](ShardRun& r, unsigned long long seq, int mode) -> int {
            LZ_CUDA(launch_k(peer_sync_kernel, dim3(1), dim3(32), 0, r.ctx->stream, r.pc, seq, mode,
                             predicated ? (const int*)(r.st.flags + 1) : (const int*)nullptr));
            ++launches;
            return LZ_OK;
        });
    };
    auto op_alpha = [](double* out) { FinOp f; f.kind = FIN_ALPHA; f.out = out; return f; };
    auto op_beta = [&](ShardRun& r, int jn, const double* mag, int omega_j) {
        FinOp f;
        f.kind = FIN_BETA;
        f.jn = jn;
        f.tol_rel = opts->breakdown_tol;
        f.magnitude = mag;
        if (omega_j >= 0) {
            f.omega_j = omega_j;
            f.om_cur = r.om_cur;
            f.om_prev = r.om_prev;
            f.delta = delta; f.eps1 = eps1; f.psi = psi;
        }
        return f;
    };
    // the beta exchange of row jn (+ omega) as a kernel of its own (sparse row shards: after the ghost push);
    // `side`: on the shard's second stream
    auto fin_beta_kernel = [&](int jn, auto&& mag_of, int omega_j, bool side) -> int {
        return exchange([&](ShardRun& r, unsigned long long seq, int mode) -> int {
            LZ_CUDA(launch_k(fin_scalar_kernel, dim3(1), dim3(kThreads), 0, side ? r.ctx->stream2 : r.ctx->stream,
                             (const double*)r.ctx->partials, r.np_beta, op_beta(r, jn, mag_of(r), omega_j), r.st, r.pc, seq,
                             mode, (const int*)nullptr));
            ++launches;
            return LZ_OK;
        });
    };
    // alpha of row jn from the in-tile partials KB left in the second half of the partials buffer (r.np of
    // them) plus the border kernel's own
    auto op_alpha_s2 = [](ShardRun& r, int jn) {
        FinOp f;
        f.kind = FIN_ALPHA_S2;
        f.jn = jn;
        f.out = r.st.alpha + jn;
        f.extra = r.ctx->partials + kMaxPartials;
        f.nextra = r.np_kb;
        return f;
    };
    // border kernel of row jn (the vector KB just produced): the edges that cross CTA tiles, z-chunks and the
    // slab top, whose upper ghost plane arrived with the beta exchange
    auto alpha_border = [&](int jn) -> int {
        return produce_fin([&](ShardRun& r) { return op_alpha_s2(r, jn); },
                           [&](ShardRun& r, const FinTail* t) {
                               bind_ghosts(team, r, jn & 1);
                               r.kt.begin(K_BORDER);
                               const int rc = launch_alpha_border(r.op, r.row(jn), r.ctx->partials, &r.np, t);
                               r.kt.end();
                               ++launches;
                               return rc;
                           }, false);
    };

    // ---- CUDA graph for launch-bound solves ---------------------------------------------------------
    // Up to a few million unknowns a step is 4-6 kernels of a few microseconds each and the loop is bound by
    // launch latency (config 1, 200 x 200: 28 us/step of which ~5 us is arithmetic).  The whole enqueue below
    // - start vector, pre-step, n steps - is then captured once into a CUDA graph and replayed: inside a graph
    // the kernels follow each other without a trip through the launch queue.  The graph bakes in every
    // pointer and scalar, so it is keyed on all of them; it is captured the second time the same solve is
    // asked for (a one-off solve would pay more for capture + instantiation than it saves) and replayed from
    // then on.  Single shard only, not in profile mode; LZ_GRAPH=0 turns it off, LZ_GRAPH=1 forces it for
    // any size and captures on the first call.
    GraphKey gkey{};
    GraphSlot* gslot = nullptr;
    bool capturing = false;
    cudaStream_t user_stream = nullptr;
    {
        static const int gmode = []() { const char* e = getenv("LZ_GRAPH"); return e ? atoi(e) : -1; }();
        const bool small = M_local <= ((int64_t)1 << 22);
        if (!team && nl == 1 && !opts->profile && gmode != 0 && (small || gmode == 1)) {
            ShardRun& r = R[0];
            if (r.op->serial == 0) { static unsigned long long next_serial = 0; r.op->serial = ++next_serial; }
            gkey.op_serial = r.op->serial;
            gkey.op = r.op; gkey.v0 = r.v0; gkey.V = r.V; gkey.arena = r.ctx->arena; gkey.ldv = r.ldv; gkey.M = r.M;
            gkey.n = n; gkey.opts = *opts; gkey.diag = r.op->st.diag;
            gslot = graph_lookup(r.ctx, gkey);
            if (gslot->exec == nullptr && (gslot->seen >= 1 || gmode == 1)) {
                if (!r.ctx->gstream) LZ_CUDA(cudaStreamCreateWithFlags(&r.ctx->gstream, cudaStreamNonBlocking));
                LZ_CUDA(cudaStreamBeginCapture(r.ctx->gstream, cudaStreamCaptureModeRelaxed));
                user_stream = r.ctx->stream;
                r.ctx->stream = r.ctx->gstream;              // every launch below goes into the capture
                capturing = true;
            }
            gslot->seen += 1;
        }
    }
    const bool replay = gslot && gslot->exec && !capturing;

    if (!capturing) LZ_CHECK(each([&](ShardRun& r) { LZ_CUDA(cudaEventRecord(r.ctx->ev_begin, r.ctx->stream)); return LZ_OK; }));

    auto enqueue_all = [&]() -> int {
    // ---- start: |v0| ---------------------------------------------------------------------
    LZ_CHECK(produce_fin([](ShardRun&) { FinOp f; f.kind = FIN_V0NORM; return f; },
                         [&](ShardRun& r, const FinTail* t) {
                             ++launches;
                             return launch_dot(r.ctx, r.v0, r.v0, r.M, r.ctx->partials, &r.np, t);
                         }, false));
    bool alpha_known = false;            // alpha of the coming row already produced by KB + border kernel
    if (ref) {
        // pre-step (Lanczos.py:108-110): r = H q - (q.Hq) q with q = v0/|v0|; r becomes row 0.
        // Sharded: the ghost planes of v0 travel through the parity-1 buffers.
        if (team) {
            LZ_CHECK(each([&](ShardRun& r) {
                HaloPush h = halo_for(team, r, 1);
                if (h.lo_dst || h.hi_dst) ++launches;
                LZ_CHECK(launch_halo_push(r.ctx, r.v0, r.M, &h));
                return push_ghosts(team, r, r.v0, 1, nullptr, &launches);
            }));
            LZ_CHECK(peer_sync(false));
        }
        LZ_CHECK(produce_fin([&](ShardRun& r) { return op_alpha(r.st.alpha_pre); },
                             [&](ShardRun& r, const FinTail* t) {
                                 bind_ghosts(team, r, 1);
                                 bind_gather(team, r, 1);
                                 int l2 = 0;
                                 const int rc = launch_apply_dot(r.op, r.v0, r.st.v0scale, recompute ? nullptr : r.w,
                                                                 r.ctx->partials, &r.np, &l2, nullptr, t);
                                 launches += l2;
                                 return rc;
                             }, false));
        if (sparse_team) {
            LZ_CHECK(each([&](ShardRun& r) {
                ++launches;
                LZ_CHECK(launch_update_norm(r.ctx, r.w, r.v0, nullptr, r.st.alpha_pre, r.st.v0scale, nullptr, nullptr,
                                            r.row(0), r.M, r.ctx->partials, &r.np_beta, nullptr, nullptr));
                return push_ghosts(team, r, r.row(0), 0, nullptr, &launches);
            }));
            LZ_CHECK(fin_beta_kernel(0, [](ShardRun& r) { return (const double*)r.st.alpha_pre; }, -1, false));
        } else
        LZ_CHECK(produce_fin([&](ShardRun& r) { return op_beta(r, 0, r.st.alpha_pre, -1); },
                             [&](ShardRun& r, const FinTail* t) {
                                 HaloPush h = halo_for(team, r, 0);
                                 if (kba) {
                                     StencilUpdate u;
                                     u.ca = r.st.alpha_pre;
                                     u.sa = r.st.v0scale;
                                     FinOp fa;
                                     fa.kind = FIN_ALPHA_S2; fa.jn = 0; fa.out = r.st.alpha;
                                     ++launches;
                                     return launch_kba_step(r.op, r.v0, r.st.v0scale, &u, r.row(0), t, &fa, &r.np);
                                 }
                                 if (recompute) {
                                     StencilUpdate u;
                                     u.ca = r.st.alpha_pre;
                                     u.sa = r.st.v0scale;
                                     u.halo = h;        // KB stores the boundary planes of row 0 to the neighbours
                                     if (kb_alpha) u.alpha_partials = r.ctx->partials + kMaxPartials;
                                     int l2 = 0;
                                     LZ_CHECK(launch_apply_update_norm(r.op, r.v0, r.st.v0scale, &u, r.row(0), r.ctx->partials,
                                                                       &r.np, &l2, t));
                                     r.np_kb = r.np;
                                     launches += l2;
                                     return LZ_OK;
                                 }
                                 ++launches;
                                 LZ_CHECK(launch_update_norm(r.ctx, r.w, r.v0, nullptr, r.st.alpha_pre, r.st.v0scale, nullptr,
                                                             nullptr, r.row(0), r.M, r.ctx->partials, &r.np, &h, t));
                                 return push_ghosts(team, r, r.row(0), 0, nullptr, &launches);
                             }, false));
        if (kb_alpha) {
            LZ_CHECK(alpha_border(0));
            alpha_known = true;
        }
        if (kba) alpha_known = true;
    } else {
        LZ_CHECK(each([&](ShardRun& r) {
            LZ_CUDA(cudaMemcpyAsync(r.row(0), r.v0, (size_t)r.M * 8, cudaMemcpyDeviceToDevice, r.ctx->stream));
            init_first_row_kernel<<<1, 32, 0, r.ctx->stream>>>(r.st);
            ++launches;
            HaloPush h = halo_for(team, r, 0);
            if (h.lo_dst || h.hi_dst) ++launches;
            LZ_CHECK(launch_halo_push(r.ctx, r.v0, r.M, &h));
            return push_ghosts(team, r, r.v0, 0, nullptr, &launches);
        }));
        LZ_CHECK(peer_sync(false));
    }

    if (fused) {
        ShardRun& r = R[0];
        cudaStream_t q = r.ctx->stream;
        const int* flag = sel ? r.st.flags + 1 : nullptr;
        PeerComm solo;                       // world = 1
        double* ucur = r.w;
        double* unxt = r.w2;
        int l2 = 0;
        auto alpha_s2 = [&](int jn, const int* fl) {
            FinOp f;
            f.kind = FIN_ALPHA_S2;
            f.jn = jn;
            f.out = r.st.alpha + jn;
            fin_scalar_kernel<<<1, kThreads, 0, q>>>(r.ctx->partials, r.np, f, r.st, solo, 0, LZ_XCHG_FUSED, fl);
            ++launches;
        };
        // u_0 = H r_0 (un-normalised), alpha_0 = s_0^2 r_0.u_0
        r.kt.begin(K_APPLY);
        LZ_CHECK(launch_apply_dot(r.op, r.row(0), nullptr, ucur, r.ctx->partials, &r.np, &l2, nullptr));
        r.kt.end();
        launches += l2;
        alpha_s2(0, nullptr);
        for (int j = 0; j < n; ++j) {
            if (sel && j > 0) {
                // predicated on the device flag: sweep q_j, then recompute u_j = H q_j and alpha_j
                for (int p = 0; p < passes; ++p) {
                    r.kt.begin(K_DOTS);
                    LZ_CHECK(launch_cgs_dots(r.ctx, r.V, r.ldv, j, r.row(j), r.M, r.gs_part, &r.ncg, flag));
                    r.kt.end();
                    fin_ip_kernel<<<1, kThreads, 0, q>>>(r.gs_part, r.ncg, j, 0, 0, r.st, solo, 0, LZ_XCHG_FUSED, flag,
                                                         p == 0 ? 1 : 0);
                    r.kt.begin(K_GSUPD);
                    LZ_CHECK(launch_cgs_update(r.ctx, r.V, r.ldv, j, r.row(j), r.st.coef, r.st.cself, r.row(j), r.M,
                                               flag, nullptr));
                    r.kt.end();
                    launches += 3;
                }
                LZ_CHECK(launch_apply_dot(r.op, r.row(j), nullptr, ucur, r.ctx->partials, &r.np, &l2, flag));
                launches += l2;
                alpha_s2(j, flag);
            }
            if (j + 1 < n) {
                r.kt.begin(K_FUSED);
                LZ_CHECK(launch_fused_step(r.op, ucur, r.row(j), j > 0 ? r.row(j - 1) : nullptr, r.st.scale + j,
                                           r.st.alpha + j, r.st.beta + j, j > 0 ? r.st.scale + j - 1 : nullptr,
                                           r.row(j + 1), unxt, r.ctx->partials, &r.np));
                r.kt.end();
                fin_fused_kernel<<<1, kThreads, 0, q>>>(r.ctx->partials, r.np, r.st, j + 1, opts->breakdown_tol,
                                                        r.st.alpha);
                launches += 2;
                std::swap(ucur, unxt);
                if (sel) {
                    omega_kernel<<<1, kThreads, 0, q>>>(r.st, j, r.om_cur, r.om_prev, delta, eps1, psi);
                    ++launches;
                    std::swap(r.om_cur, r.om_prev);
                }
            }
        }
    }

    for (int j = 0; !fused && j < n; ++j) {
        const int par = j & 1;
        // ---- Gram-Schmidt sweeps of q_j against the rows before it ---------------------------
        const bool maybe_reorth = ((reorth == LZ_REORTH_FULL) || (sel && j > 0)) && (j > 0 || ref);
        if (maybe_reorth) {
            bool pushed = false;
            // CGS2: the update of the first sweep and the dots of the second read the same j rows -
            // one fused kernel (K4c) when the staged tile fits shared memory
            bool fuse = allow_gs_fusion && j >= 1;
            for (int s = 0; s < nl && fuse; ++s) fuse = cgs_update_dots_supported(R[s].V, R[s].ldv, j, R[s].row(j));
            // the coefficients of a sweep are formed in the tail of the kernel that produced the dots (fin.cuh
            // IpTail); several local shards: the tails push, one combine kernel per shard follows
            auto dots_with_coefficients = [&](int form, int count_it, auto&& produce) -> int {
                unsigned long long seq = 0;
                if (team) seq = ++team->seq;
                LZ_CHECK(each([&](ShardRun& r) {
                    IpTail t;
                    t.on = 1;
                    t.j = j;
                    t.self_included = form;
                    t.ref_form = form;
                    t.count = (count_it && !split) ? 1 : 0;
                    t.ticket = r.ctx->tickets;
                    t.st = r.st;
                    t.pc = r.pc;
                    t.seq = seq;
                    t.mode = split ? (int)LZ_XCHG_PUSH : (int)LZ_XCHG_FUSED;
                    return produce(r, &t);
                }));
                if (!split) return LZ_OK;
                return each([&](ShardRun& r) {
                    LZ_CUDA(launch_k(fin_ip_kernel, dim3(1), dim3(kThreads), 0, r.ctx->stream, (const double*)r.gs_part,
                                     r.ncg, j, form, form, r.st, r.pc, seq, (int)LZ_XCHG_COMBINE,
                                     sel ? (const int*)(r.st.flags + 1) : (const int*)nullptr, count_it));
                    ++launches;
                    return LZ_OK;
                });
            };
            for (int p = 0; p < passes; ++p) {
                const int ref_form = (ref && p == 0 && !gpu_sweep) ? 1 : 0;   // (2 - |v|^2) form incl. the self term
                const int nrows = ref_form ? j + 1 : j;      // the reference's sum includes row j itself
                if (nrows == 0) continue;
                if (!(fuse && p == 1)) {                     // (fused CGS2: K4c already left the second sweep's coefficients)
                    LZ_CHECK(dots_with_coefficients(ref_form, p == 0 ? 1 : 0, [&](ShardRun& r, const IpTail* t) {
                        r.kt.begin(K_DOTS);
                        const int rc = launch_cgs_dots(r.ctx, r.V, r.ldv, nrows, r.row(j), r.M, r.gs_part, &r.ncg,
                                                       sel ? r.st.flags + 1 : nullptr, t);
                        r.kt.end();
                        ++launches;
                        return rc;
                    }));
                }
                if (fuse && p == 0) {
                    LZ_CHECK(dots_with_coefficients(0, 0, [&](ShardRun& r, const IpTail* t) {
                        r.kt.begin(K_GSFUSED);
                        const int rc = launch_cgs_update_dots(r.ctx, r.V, r.ldv, j, r.row(j), r.st.coef, r.st.cself, r.M,
                                                              r.gs_part, &r.ncg, sel ? r.st.flags + 1 : nullptr, t);
                        r.kt.end();
                        ++launches;
                        return rc;
                    }));
                    continue;                 // its halo planes / ghost entries travel after the second sweep
                }
                LZ_CHECK(each([&](ShardRun& r) {
                    HaloPush h = halo_for(team, r, par);
                    r.kt.begin(K_GSUPD);
                    const int rc = launch_cgs_update(r.ctx, r.V, r.ldv, j, r.row(j), r.st.coef, r.st.cself, r.row(j),
                                                     r.M, sel ? r.st.flags + 1 : nullptr, &h);
                    r.kt.end();
                    ++launches;
                    LZ_CHECK(rc);
                    return push_ghosts(team, r, r.row(j), par, sel ? r.st.flags + 1 : nullptr, &launches);
                }));
                pushed = true;
            }
            // the re-written row's halo planes must reach the neighbours before K1 reads them
            if (pushed) LZ_CHECK(peer_sync(sel));
        }
        // ---- w = H q_j, alpha_j = q_j . w ------------------------------------------------------
        // (alpha_j known from KB + border kernel of the previous step: only if a sweep just rewrote row j)
        if (sparse_team) {
            // alpha_j = s_j^2 (r_j . H r_j), H applied to the un-normalised row; partials in the second half of the
            // buffer (the beta exchange of the previous step may still be reading the first half)
            if (interior_done && sel) {
                // the early interior apply was speculative (flags[5 + (j & 1)]): it runs now if it was not started,
                // or again if a sweep just rewrote row j (flags[7], set by the monitor of the previous step)
                LZ_CHECK(each([&](ShardRun& r) {
                    ++launches;
                    return launch_spmv_part(r.op, 1, r.row(j), nullptr, r.w, r.ctx->partials + kMaxPartials, &r.np,
                                            r.st.flags + 7, nullptr, r.ctx->stream);
                }));
            }
            const bool early = interior_done;
            LZ_CHECK(produce_fin([&](ShardRun& r) { FinOp f; f.kind = FIN_ALPHA_S2; f.jn = j; f.out = r.st.alpha + j; return f; },
                                 [&](ShardRun& r, const FinTail* t) {
                                     bind_gather(team, r, par);
                                     int l2 = 1;
                                     int rc;
                                     r.kt.begin(K_APPLY);
                                     if (early) rc = launch_spmv_part(r.op, 2, r.row(j), nullptr, r.w, r.ctx->partials + kMaxPartials,
                                                                      &r.np, nullptr, t, r.ctx->stream);
                                     else rc = launch_apply_dot(r.op, r.row(j), nullptr, r.w, r.ctx->partials + kMaxPartials, &r.np,
                                                                &l2, nullptr, t);
                                     r.kt.end();
                                     launches += l2;
                                     return rc;
                                 }, false, kMaxPartials));
            interior_done = false;
        } else if (!alpha_known || maybe_reorth) {
            const bool pred = alpha_known;
            LZ_CHECK(produce_fin([&](ShardRun& r) { return op_alpha(r.st.alpha + j); },
                                 [&](ShardRun& r, const FinTail* t) {
                                     bind_ghosts(team, r, par);
                                     bind_gather(team, r, par);
                                     int l2 = 0;
                                     r.kt.begin(K_APPLY);
                                     const int rc = launch_apply_dot(r.op, r.row(j), r.st.scale + j, recompute ? nullptr : r.w,
                                                                     r.ctx->partials, &r.np, &l2,
                                                                     pred ? r.st.flags + 1 : nullptr, t);
                                     r.kt.end();
                                     launches += l2;
                                     return rc;
                                 }, pred));
        } else {
            LZ_CHECK(each([&](ShardRun& r) { bind_ghosts(team, r, par); bind_gather(team, r, par); return LZ_OK; }));
        }
        // ---- r = w - alpha_j q_j - beta_j q_{j-1}; beta_{j+1} = |r| --------------------------
        const bool want_alpha = kb_alpha && (j + 1 < n);
        if (sparse_team) {
            const bool ahead = overlap && (j + 1 < n);
            LZ_CHECK(each([&](ShardRun& r) {
                double* out = (j + 1 < n) ? r.row(j + 1) : r.w;
                r.kt.begin(K_UPDATE);
                const int rc = launch_update_norm(r.ctx, r.w, r.row(j), j > 0 ? r.row(j - 1) : nullptr, r.st.alpha + j,
                                                  r.st.scale + j, r.st.beta + j, j > 0 ? r.st.scale + j - 1 : nullptr, out,
                                                  r.M, r.ctx->partials, &r.np_beta, nullptr, nullptr, r.st.scale + j);
                r.kt.end();
                ++launches;
                LZ_CHECK(rc);
                if (ahead) {
                    // fork: ghost push + beta exchange on the second stream, interior spans of H row_{j+1} here
                    LZ_CUDA(cudaEventRecord(r.ctx->ev_fork, r.ctx->stream));
                    LZ_CUDA(cudaStreamWaitEvent(r.ctx->stream2, r.ctx->ev_fork, 0));
                }
                if (j + 1 < n) return push_ghosts(team, r, out, (j + 1) & 1, nullptr, &launches, ahead ? r.ctx->stream2 : nullptr);
                return LZ_OK;
            }));
            LZ_CHECK(fin_beta_kernel(j + 1, [](ShardRun& r) { return (const double*)r.st.alpha; }, (sel && j + 1 < n) ? j : -1, ahead));
            if (ahead) {
                LZ_CHECK(each([&](ShardRun& r) {
                    LZ_CUDA(cudaEventRecord(r.ctx->ev_join, r.ctx->stream2));
                    ++launches;
                    r.kt.begin(K_APPLY);
                    const int rc = launch_spmv_part(r.op, 1, r.row(j + 1), nullptr, r.w, r.ctx->partials + kMaxPartials, &r.np,
                                                    sel ? r.st.flags + 5 + ((j + 1) & 1) : nullptr, nullptr, r.ctx->stream);
                    r.kt.end();
                    LZ_CHECK(rc);
                    LZ_CUDA(cudaStreamWaitEvent(r.ctx->stream, r.ctx->ev_join, 0));      // join: beta, scale, flags, ghosts
                    return LZ_OK;
                }));
                interior_done = true;
            }
        } else
        LZ_CHECK(produce_fin([&](ShardRun& r) { return op_beta(r, j + 1, r.st.alpha, (sel && j + 1 < n) ? j : -1); },
                             [&](ShardRun& r, const FinTail* t) {
                                 double* out = (j + 1 < n) ? r.row(j + 1) : r.w;
                                 HaloPush h = (j + 1 < n) ? halo_for(team, r, (j + 1) & 1) : HaloPush{};
                                 if (kba) {
                                     StencilUpdate u;
                                     u.b = j > 0 ? r.row(j - 1) : nullptr;
                                     u.ca = r.st.alpha + j;
                                     u.sa = r.st.scale + j;
                                     u.cb = r.st.beta + j;
                                     u.sb = j > 0 ? r.st.scale + j - 1 : nullptr;
                                     FinOp fa;
                                     fa.kind = FIN_ALPHA_S2; fa.jn = j + 1; fa.out = r.st.alpha + j + 1;
                                     r.kt.begin(K_UPDATE);
                                     const int rc3 = launch_kba_step(r.op, r.row(j), r.st.scale + j, &u, out, t,
                                                                     (j + 1 < n) ? &fa : nullptr, &r.np);
                                     r.kt.end();
                                     ++launches;
                                     return rc3;
                                 }
                                 if (recompute) {
                                     StencilUpdate u;
                                     u.b = j > 0 ? r.row(j - 1) : nullptr;
                                     u.ca = r.st.alpha + j;
                                     u.sa = r.st.scale + j;
                                     u.cb = r.st.beta + j;
                                     u.sb = j > 0 ? r.st.scale + j - 1 : nullptr;
                                     u.halo = h;
                                     if (want_alpha) u.alpha_partials = r.ctx->partials + kMaxPartials;
                                     int l2 = 0;
                                     r.kt.begin(K_UPDATE);
                                     const int rc2 = launch_apply_update_norm(r.op, r.row(j), r.st.scale + j, &u, out,
                                                                              r.ctx->partials, &r.np, &l2, t);
                                     r.kt.end();
                                     r.np_kb = r.np;
                                     launches += l2;
                                     return rc2;
                                 }
                                 r.kt.begin(K_UPDATE);
                                 const int rc = launch_update_norm(r.ctx, r.w, r.row(j), j > 0 ? r.row(j - 1) : nullptr,
                                                                   r.st.alpha + j, r.st.scale + j, r.st.beta + j,
                                                                   j > 0 ? r.st.scale + j - 1 : nullptr, out, r.M,
                                                                   r.ctx->partials, &r.np, &h, t);
                                 r.kt.end();
                                 ++launches;
                                 LZ_CHECK(rc);
                                 if (j + 1 < n) return push_ghosts(team, r, out, (j + 1) & 1, nullptr, &launches);
                                 return LZ_OK;
                             }, false));
        if (sel && j + 1 < n)
            for (int s = 0; s < nl; ++s) std::swap(R[s].om_cur, R[s].om_prev);
        alpha_known = false;
        if (want_alpha) {
            LZ_CHECK(alpha_border(j + 1));
            alpha_known = true;
        }
        if (kba && j + 1 < n) alpha_known = true;
    }
    return LZ_OK;
    };   // enqueue_all

    if (replay) {
        launches = gslot->launches;
        LZ_CUDA(cudaGraphLaunch(gslot->exec, R[0].ctx->stream));
    } else {
        const int rc_enq = enqueue_all();
        if (capturing) {
            ShardRun& r = R[0];
            cudaGraph_t graph = nullptr;
            const cudaError_t ce = cudaStreamEndCapture(r.ctx->gstream, &graph);
            r.ctx->stream = user_stream;
            if (rc_enq != LZ_OK) { if (graph) cudaGraphDestroy(graph); return rc_enq; }
            if (ce != cudaSuccess || !graph) {
                set_error("lz_lanczos_run: graph capture failed: %s", cudaGetErrorString(ce));
                (void)cudaGetLastError();
                return LZ_ERR_CUDA;
            }
            cudaGraphExec_t exec = nullptr;
            const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ie != cudaSuccess) { set_error("lz_lanczos_run: cudaGraphInstantiate: %s", cudaGetErrorString(ie)); return LZ_ERR_CUDA; }
            gslot->exec = exec;
            gslot->launches = launches;
            LZ_CUDA(cudaEventRecord(r.ctx->ev_begin, r.ctx->stream));
            LZ_CUDA(cudaGraphLaunch(exec, r.ctx->stream));
        } else {
            LZ_CHECK(rc_enq);
        }
    }
    LZ_CHECK(each([&](ShardRun& r) {
        LZ_CUDA(cudaGetLastError());
        LZ_CUDA(cudaEventRecord(r.ctx->ev_end, r.ctx->stream));
        return LZ_OK;
    }));

    // ---- results (identical on every shard: read them from shard 0) -----------------------------
    const size_t blk = (nd * 8 + 511) & ~(size_t)511;
    std::vector<char> h_blk(3 * blk);
    std::vector<int> h_flags(8 * nl);
    for (int s = 0; s < nl; ++s) {
        ShardRun& r = R[s];
        if (nl > 1) LZ_CUDA(cudaSetDevice(r.ctx->device));
        if (s == 0) LZ_CUDA(cudaMemcpyAsync(h_blk.data(), r.st.alpha, 3 * blk, cudaMemcpyDeviceToHost, r.ctx->stream));
        LZ_CUDA(cudaMemcpyAsync(&h_flags[8 * s], r.st.flags, 8 * sizeof(int), cudaMemcpyDeviceToHost, r.ctx->stream));
    }
    for (int s = 0; s < nl; ++s) {
        if (nl > 1) LZ_CUDA(cudaSetDevice(R[s].ctx->device));
        LZ_CUDA(cudaStreamSynchronize(R[s].ctx->stream));
    }
    const double* h_alpha = reinterpret_cast<const double*>(h_blk.data());
    const double* h_beta = reinterpret_cast<const double*>(h_blk.data() + blk);
    const double* h_scale = reinterpret_cast<const double*>(h_blk.data() + 2 * blk);
    for (int j = 0; j < n; ++j) alpha_host[j] = h_alpha[j];
    for (int k = 0; k + 1 < n; ++k) beta_host[k] = h_beta[k + 1];     // Lanczos.py:112 numbering
    if (row_scale_host)
        for (int j = 0; j < n; ++j) row_scale_host[j] = h_scale[j];
    float ms = 0.f;
    float kms[K_NKINDS] = {};
    int kcnt[K_NKINDS] = {};
    for (int s = 0; s < nl; ++s) {
        ShardRun& r = R[s];
        if (nl > 1) LZ_CUDA(cudaSetDevice(r.ctx->device));
        float m1 = 0.f;
        LZ_CUDA(cudaEventElapsedTime(&m1, r.ctx->ev_begin, r.ctx->ev_end));
        ms = std::max(ms, m1);
        float km[K_NKINDS];
        int kc[K_NKINDS];
        r.kt.collect(km, kc, K_NKINDS);
        if (s == 0) for (int k = 0; k < K_NKINDS; ++k) { kms[k] = km[k]; kcnt[k] = kc[k]; }
    }
    int steps_done = n;
    int status = LZ_OK;
    for (int s = 0; s < nl; ++s) {
        if (h_flags[8 * s + 4]) {
            set_error("multi-GPU exchange timed out on rank %d (a peer did not arrive)", team ? team->shards[s].rank : 0);
            status = LZ_ERR_PEER;
        }
    }
    // a breakdown at index jn means row jn could not be normalised: steps 0..jn-1 are valid.
    // beta[n] (after the last step) is not part of the output and is ignored.
    if (status == LZ_OK && h_flags[0] >= 0 && h_flags[0] < n) {
        steps_done = h_flags[0];
        set_error("Lanczos breakdown: beta[%d] = %.3e (Krylov space exhausted after %d steps)",
                  h_flags[0], h_beta[h_flags[0]], steps_done);
        status = LZ_ERR_BREAKDOWN;
    }
    if (info) {
        memset(info, 0, sizeof(*info));
        info->steps_done = steps_done;
        info->reorth_count = (reorth == LZ_REORTH_FULL) ? (ref ? n : n - 1) : h_flags[2];
        info->launches = launches;
        info->gpu_ms = ms;
        info->apply_ms = kms[K_APPLY];   info->apply_launches = kcnt[K_APPLY];
        info->update_ms = kms[K_UPDATE]; info->update_launches = kcnt[K_UPDATE];
        info->dots_ms = kms[K_DOTS];     info->dots_launches = kcnt[K_DOTS];
        info->gsupd_ms = kms[K_GSUPD];   info->gsupd_launches = kcnt[K_GSUPD];
        info->fused_ms = kms[K_FUSED];   info->fused_launches = kcnt[K_FUSED];
        info->step_kernel = fused ? 2 : (recompute ? 3 : 1);
        info->gsfused_ms = kms[K_GSFUSED]; info->gsfused_launches = kcnt[K_GSFUSED];
        info->border_ms = kms[K_BORDER]; info->border_launches = kcnt[K_BORDER];
        info->alpha_in_update = kb_alpha ? 1 : (kba ? 2 : 0);
        info->overlap = overlap ? 1 : 0;
        info->graph = replay ? 2 : (capturing ? 1 : 0);
    }
    return status;
}

}  // namespace

void lz_ctx_drop_graphs(lz_ctx* c) {
    for (lz_graph_slot* s : c->graphs) {
        if (s->exec) cudaGraphExecDestroy(s->exec);
        delete s;
    }
    c->graphs.clear();
}

extern "C" int lz_lanczos_run(lz_ctx* ctx, lz_op* op, const double* v0_dev, int32_t n,
                              const lz_run_opts* opts, double* alpha_host, double* beta_host,
                              double* V_dev, int64_t ldv, double* row_scale_host, lz_run_info* info) {
    LZ_REQUIRE(ctx && op && v0_dev && opts && alpha_host, "lz_lanczos_run: null argument");
    LZ_REQUIRE(op->ctx == ctx, "lz_lanczos_run: operator belongs to another context");
    LZ_REQUIRE(!op->st.sharded, "lz_lanczos_run: this operator is a shard of a team; use lz_team_lanczos_run");
    LZ_CUDA(cudaSetDevice(ctx->device));
    // launch-bound sizes: the whole solve in one persistent kernel (small.cu)
    if (n >= 1 && !(opts->ref_compat && n < 2) && n <= op->M && (n < 2 || beta_host) && opts->cgs_passes <= 2 &&
        (opts->reorth == LZ_REORTH_NONE || opts->reorth == LZ_REORTH_FULL) &&
        small_solve_supported(op, opts, n, V_dev, ldv))
        return launch_small_solve(ctx, op, v0_dev, n, opts, alpha_host, beta_host, V_dev, ldv, row_scale_host, info);
    lz_op* ops[1] = {op};
    const double* v0s[1] = {v0_dev};
    double* Vs[1] = {V_dev};
    const int64_t ldvs[1] = {ldv};
    return run_loop(nullptr, 1, ops, v0s, n, opts, alpha_host, beta_host, Vs, ldvs, row_scale_host, info);
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

// ================================ multi-GPU team ==============================================

extern "C" int lz_comm_bytes(int world, int32_t max_steps, int64_t plane, int64_t nghost, int64_t* bytes) {
    LZ_REQUIRE(bytes && world >= 1 && world <= kMaxWorld && max_steps >= 1 && plane >= 0 && nghost >= 0,
               "lz_comm_bytes: bad argument (world <= %d)", kMaxWorld);
    *bytes = (int64_t)CommLayout::make(world, max_steps + 2, plane, nghost).total;
    return LZ_OK;
}

extern "C" int lz_comm_alloc(lz_ctx* ctx, int64_t bytes, void** dev_ptr, unsigned char* ipc_handle64) {
    LZ_REQUIRE(ctx && dev_ptr && bytes > 0, "lz_comm_alloc: bad argument");
    LZ_CUDA(cudaSetDevice(ctx->device));
    void* p = nullptr;
    LZ_CUDA(cudaMalloc(&p, (size_t)bytes));
    LZ_CUDA(cudaMemset(p, 0, (size_t)bytes));
    LZ_CUDA(cudaDeviceSynchronize());
    if (ipc_handle64) {
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        cudaIpcMemHandle_t h;
        cudaError_t e = cudaIpcGetMemHandle(&h, p);
        if (e != cudaSuccess) {
            cudaFree(p);
            set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
            return LZ_ERR_CUDA;
        }
        memcpy(ipc_handle64, &h, 64);
    }
    *dev_ptr = p;
    return LZ_OK;
}

extern "C" int lz_comm_open(lz_ctx* ctx, const unsigned char* ipc_handle64, void** dev_ptr) {
    LZ_REQUIRE(ctx && ipc_handle64 && dev_ptr, "lz_comm_open: bad argument");
    LZ_CUDA(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle64, 64);
    LZ_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return LZ_OK;
}

extern "C" int lz_comm_close(lz_ctx* ctx, void* dev_ptr) {
    LZ_REQUIRE(ctx, "lz_comm_close: bad argument");
    if (!dev_ptr) return LZ_OK;
    LZ_CUDA(cudaSetDevice(ctx->device));
    LZ_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return LZ_OK;
}

extern "C" int lz_comm_free(lz_ctx* ctx, void* dev_ptr) {
    LZ_REQUIRE(ctx, "lz_comm_free: bad argument");
    if (!dev_ptr) return LZ_OK;
    LZ_CUDA(cudaSetDevice(ctx->device));
    LZ_CUDA(cudaFree(dev_ptr));
    return LZ_OK;
}

extern "C" int lz_team_create(int world, int nlocal, const int* local_ranks, lz_ctx* const* ctxs,
                              int64_t global_rows, int32_t max_steps, int64_t plane, int64_t nghost,
                              lz_team** out) {
    LZ_REQUIRE(out && local_ranks && ctxs, "lz_team_create: null argument");
    LZ_REQUIRE(world >= 1 && world <= kMaxWorld, "lz_team_create: world must be in 1..%d", kMaxWorld);
    LZ_REQUIRE(nlocal >= 1 && nlocal <= world, "lz_team_create: bad nlocal");
    LZ_REQUIRE(max_steps >= 1 && plane >= 0 && nghost >= 0, "lz_team_create: bad sizes");
    lz_team* t = new lz_team();
    t->world = world;
    t->nlocal = nlocal;
    t->kmax = max_steps + 2;
    t->plane = plane;
    t->nghost = nghost;
    t->global_rows = global_rows;
    t->layout = CommLayout::make(world, t->kmax, plane, nghost);
    t->shards.resize(nlocal);
    for (int s = 0; s < nlocal; ++s) {
        if (local_ranks[s] < 0 || local_ranks[s] >= world || !ctxs[s]) {
            delete t;
            set_error("lz_team_create: bad rank or context for local shard %d", s);
            return LZ_ERR_INVALID;
        }
        t->shards[s].ctx = ctxs[s];
        t->shards[s].rank = local_ranks[s];
        t->shards[s].plane = plane;
    }
    // shards of one process on different devices talk through direct peer access
    for (int a = 0; a < nlocal; ++a)
        for (int b = 0; b < nlocal; ++b) {
            const int da = ctxs[a]->device, db = ctxs[b]->device;
            if (da == db) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, da, db);
            if (can) {
                cudaSetDevice(da);
                cudaError_t e = cudaDeviceEnablePeerAccess(db, 0);
                if (e != cudaSuccess) (void)cudaGetLastError();     // already enabled
            }
        }
    *out = t;
    return LZ_OK;
}

extern "C" int lz_team_attach(lz_team* team, int local_index, void* const* comm_ptrs, int lower_rank,
                              int upper_rank) {
    LZ_REQUIRE(team && comm_ptrs, "lz_team_attach: null argument");
    LZ_REQUIRE(local_index >= 0 && local_index < team->nlocal, "lz_team_attach: bad shard index");
    LZ_REQUIRE(lower_rank >= -1 && lower_rank < team->world && upper_rank >= -1 && upper_rank < team->world,
               "lz_team_attach: bad neighbour rank");
    lz_shard& sh = team->shards[local_index];
    sh.lower = lower_rank;
    sh.upper = upper_rank;
    sh.pc.world = team->world;
    sh.pc.rank = sh.rank;
    sh.pc.kmax = team->kmax;
    for (int q = 0; q < team->world; ++q) {
        LZ_REQUIRE(comm_ptrs[q], "lz_team_attach: exchange buffer of rank %d is null", q);
        sh.comm[q] = comm_ptrs[q];
        sh.pc.flags[q] = reinterpret_cast<unsigned long long*>((char*)comm_ptrs[q] + team->layout.flags_off);
        sh.pc.slots[q] = reinterpret_cast<double*>((char*)comm_ptrs[q] + team->layout.slots_off);
    }
    return LZ_OK;
}

extern "C" int lz_team_set_ghosts(lz_team* team, int local_index, int32_t nsend, const int32_t* send_idx_host,
                                  const int32_t* seg_start, const int64_t* dst_off) {
    LZ_REQUIRE(team && seg_start && dst_off, "lz_team_set_ghosts: null argument");
    LZ_REQUIRE(local_index >= 0 && local_index < team->nlocal, "lz_team_set_ghosts: bad shard index");
    LZ_REQUIRE(nsend >= 0 && (nsend == 0 || send_idx_host), "lz_team_set_ghosts: bad send list");
    lz_shard& sh = team->shards[local_index];
    LZ_REQUIRE(seg_start[0] == 0 && seg_start[team->world] == nsend, "lz_team_set_ghosts: segments do not span the send list");
    LZ_CUDA(cudaSetDevice(sh.ctx->device));
    if (sh.send_idx) { cudaFree(sh.send_idx); sh.send_idx = nullptr; }
    sh.nsend = nsend;
    for (int q = 0; q <= team->world; ++q) sh.seg_start[q] = seg_start[q];
    for (int q = 0; q < team->world; ++q) {
        LZ_REQUIRE(seg_start[q + 1] >= seg_start[q], "lz_team_set_ghosts: segments not monotone");
        LZ_REQUIRE(dst_off[q] >= 0 && dst_off[q] + (seg_start[q + 1] - seg_start[q]) <= team->nghost,
                   "lz_team_set_ghosts: destination range of rank %d exceeds the gather buffer", q);
        sh.dst_off[q] = dst_off[q];
    }
    if (nsend > 0) {
        LZ_CUDA(cudaMalloc((void**)&sh.send_idx, (size_t)nsend * 4));
        LZ_CUDA(cudaMemcpy(sh.send_idx, send_idx_host, (size_t)nsend * 4, cudaMemcpyHostToDevice));
    }
    return LZ_OK;
}

extern "C" int lz_team_lanczos_run(lz_team* team, lz_op* const* ops, const double* const* v0_dev,
                                   int32_t n, const lz_run_opts* opts, double* alpha_host,
                                   double* beta_host, double* const* V_dev, const int64_t* ldv,
                                   double* row_scale_host, lz_run_info* info) {
    LZ_REQUIRE(team && ops && v0_dev && opts, "lz_team_lanczos_run: null argument");
    for (int s = 0; s < team->nlocal; ++s)
        LZ_REQUIRE(team->shards[s].comm[team->shards[s].rank], "lz_team_lanczos_run: shard %d is not attached", s);
    if (team->poisoned) {
        set_error("lz_team_lanczos_run: an earlier run on this team failed part-way; its exchange sequence may "
                  "differ between the ranks - destroy the team and create a new one on every rank");
        return LZ_ERR_PEER;
    }
    const int rc = run_loop(team, team->nlocal, ops, v0_dev, n, opts, alpha_host, beta_host, V_dev, ldv,
                            row_scale_host, info);
    // a breakdown is detected identically on every rank; anything else may have left the ranks out of step
    if (rc != LZ_OK && rc != LZ_ERR_BREAKDOWN && rc != LZ_ERR_INVALID) team->poisoned = true;
    return rc;
}

// y = H x over the shards of a team, with the two global sums the residual diagnostics need.
// The same exchange as one loop step: halo planes / ghost entries of x pushed to the neighbours,
// a flag round, the operator kernel, and the rank-ordered sums of the CTA partials.
extern "C" int lz_team_apply_dots(lz_team* team, lz_op* const* ops, const double* const* x_dev,
                                  double* const* y_dev, double* dots_host) {
    LZ_REQUIRE(team && ops && x_dev && y_dev && dots_host, "lz_team_apply_dots: null argument");
    const int nl = team->nlocal;
    const bool split = nl > 1;
    std::vector<ShardRun> R(nl);
    for (int s = 0; s < nl; ++s) {
        ShardRun& r = R[s];
        r.op = ops[s];
        LZ_REQUIRE(r.op && x_dev[s] && y_dev[s], "lz_team_apply_dots: null operator or vector of shard %d", s);
        r.ctx = r.op->ctx;
        r.sh = &team->shards[s];
        LZ_REQUIRE(r.sh->ctx == r.ctx, "lz_team_apply_dots: operator %d is not on its shard's context", s);
        LZ_REQUIRE(r.sh->comm[r.sh->rank], "lz_team_apply_dots: shard %d is not attached", s);
        r.M = r.op->M;
        r.pc = r.sh->pc;
        LZ_CUDA(cudaSetDevice(r.ctx->device));
        LZ_CUDA(cudaMemsetAsync(r.ctx->scratch, 0, 64 * 8, r.ctx->stream));
        r.pc.err = reinterpret_cast<int*>(r.ctx->scratch + 8);
        r.st.flags = reinterpret_cast<int*>(r.ctx->scratch + 16);      // fin_apply(FIN_ALPHA) touches no flag
    }
    int launches = 0;
    auto each = [&](auto&& fn) -> int {
        for (int s = 0; s < nl; ++s) {
            if (nl > 1) LZ_CUDA(cudaSetDevice(R[s].ctx->device));
            LZ_CHECK(fn(R[s], s));
        }
        return LZ_OK;
    };
    auto exchange = [&](auto&& launch) -> int {
        const unsigned long long seq = ++team->seq;
        if (!split) return each([&](ShardRun& r, int) { return launch(r, seq, (int)LZ_XCHG_FUSED); });
        LZ_CHECK(each([&](ShardRun& r, int) { return launch(r, seq, (int)LZ_XCHG_PUSH); }));
        return each([&](ShardRun& r, int) { return launch(r, seq, (int)LZ_XCHG_COMBINE); });
    };
    auto sum_to = [&](int slot) -> int {
        return exchange([&](ShardRun& r, unsigned long long seq, int mode) -> int {
            FinOp f;
            f.kind = FIN_ALPHA;
            f.out = r.ctx->scratch + slot;
            LZ_CUDA(launch_k(fin_scalar_kernel, dim3(1), dim3(kThreads), 0, r.ctx->stream, (const double*)r.ctx->partials,
                             r.np, f, r.st, r.pc, seq, mode, (const int*)nullptr));
            return LZ_OK;
        });
    };
    LZ_CHECK(each([&](ShardRun& r, int s) {
        HaloPush h = halo_for(team, r, 0);
        LZ_CHECK(launch_halo_push(r.ctx, x_dev[s], r.M, &h));
        return push_ghosts(team, r, x_dev[s], 0, nullptr, &launches);
    }));
    LZ_CHECK(exchange([&](ShardRun& r, unsigned long long seq, int mode) -> int {
        LZ_CUDA(launch_k(peer_sync_kernel, dim3(1), dim3(32), 0, r.ctx->stream, r.pc, seq, mode, (const int*)nullptr));
        return LZ_OK;
    }));
    LZ_CHECK(each([&](ShardRun& r, int s) {
        bind_ghosts(team, r, 0);
        bind_gather(team, r, 0);
        int l2 = 0;
        return launch_apply_dot(r.op, x_dev[s], nullptr, y_dev[s], r.ctx->partials, &r.np, &l2, nullptr, nullptr);
    }));
    LZ_CHECK(sum_to(0));                                               // x . H x
    LZ_CHECK(each([&](ShardRun& r, int s) { return launch_dot(r.ctx, y_dev[s], y_dev[s], r.M, r.ctx->partials, &r.np); }));
    LZ_CHECK(sum_to(1));                                               // H x . H x
    std::vector<int> err(nl, 0);
    for (int s = 0; s < nl; ++s) {
        ShardRun& r = R[s];
        if (nl > 1) LZ_CUDA(cudaSetDevice(r.ctx->device));
        if (s == 0) LZ_CUDA(cudaMemcpyAsync(dots_host, r.ctx->scratch, 16, cudaMemcpyDeviceToHost, r.ctx->stream));
        LZ_CUDA(cudaMemcpyAsync(&err[s], r.pc.err, sizeof(int), cudaMemcpyDeviceToHost, r.ctx->stream));
    }
    for (int s = 0; s < nl; ++s) {
        if (nl > 1) LZ_CUDA(cudaSetDevice(R[s].ctx->device));
        LZ_CUDA(cudaStreamSynchronize(R[s].ctx->stream));
    }
    for (int s = 0; s < nl; ++s)
        if (err[s]) {
            set_error("multi-GPU exchange timed out on rank %d (a peer did not arrive)", team->shards[s].rank);
            team->poisoned = true;
            return LZ_ERR_PEER;
        }
    return LZ_OK;
}

extern "C" int lz_team_destroy(lz_team* team) {
    if (team)
        for (lz_shard& sh : team->shards)
            if (sh.send_idx) { cudaSetDevice(sh.ctx->device); cudaFree(sh.send_idx); }
    delete team;
    return LZ_OK;
}
