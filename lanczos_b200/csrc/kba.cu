// KBA: the recompute step as ONE kernel per Lanczos step - KB of step j with the alpha reduction of the
// vector it produces chasing it through L2.
//
//   KB  (stencil.cu, MODE 2):  r_{j+1} = s (H r_j) - alpha_j s r_j - beta_j s' r_{j-1},  partial of |r_{j+1}|^2   24 N bytes of HBM
//   KA2 (stencil.cu):          partial of r_{j+1} . H r_{j+1}  (alpha_{j+1} = s_{j+1}^2 times it)                   8 N bytes of HBM
//
// KA2 re-reads from HBM what KB wrote a moment ago, because a 1 GB vector does not survive in a 126 MB L2
// until the next kernel.  Here both run in one persistent grid over one queue of work items ordered by
// z-chunk:   KB(0) KB(1) KA(0) KB(2) KA(1) ... KB(C-1) KA(C-2) KA(C-1).
// A KA item of chunk c needs every plane of chunk c and the first plane of chunk c + 1 of the NEW vector:
// it waits (acquire) on two counters that the KB items bump (release) when their stores are done.  Items are
// dealt round-robin (item q -> CTA q mod grid), so an item only ever waits for items earlier in the queue,
// which resident CTAs are already working on: no deadlock, and the partial sums stay bit-reproducible.  With
// chunks of 8 planes the new vector is read back ~50 MB of traffic after it was written: from L2, not HBM.
// The step then moves 24 N bytes over HBM instead of 32 N, in one launch instead of two
// (replaces `r = H*V[j]`, `alpha[j] = np.dot(V[j], r)`, `r = r - V[j]*alpha[j] - V[j-1]*beta[j-1]`,
// `beta = norm(r)` of Lanczos.py:112-119 for the step after the first).
//
// Whole tiles only (nx % 64 == 0, ny % 16 == 0), all three couplings present, one GPU; everything else keeps
// KA2 + KB.  The bodies below are the MODE 2 body of stencil_apply_dot_kernel and the body of
// stencil_alpha_fast_kernel for one work item each, minus the predicates whole tiles make redundant; the new
// vector is read with coherent loads (it was written by this very kernel: ld.global.nc is off limits).
#include <stdlib.h>
#include "internal.h"

namespace lz {

struct KbaArgs {
    int nx, ny, nz, periodic;
    int64_t plane;
    double c, ox, oy, oz;
    const double* x;        // r_j (row j)
    const double* b;        // r_{j-1} (nullable)
    double* y;              // r_{j+1}
    const double* diag;
    const double* scale;    // s_j
    const double* ca; const double* sa; const double* cb; const double* sb;
    int zc, chunks;         // planes per chunk, number of chunks
    int tb_x, tb_y;         // KB tiles (64 x 8) per plane
    int ta_x, ta_y;         // KA tiles (64 x 16) per plane
    int* done;              // [chunks] KB items finished per chunk (zero at launch; reset by the tail)
    double* partials;       // [grid] |r_{j+1}|^2
    double* alpha_partials; // [grid] r_{j+1} . H r_{j+1}
    FinTail fin_beta;       // FIN_BETA (+ omega) of row j+1
    FinOp fin_alpha;        // FIN_ALPHA_S2 of row j+1 (runs after fin_beta: needs scale[j+1])
};

__device__ __forceinline__ double2 ld_coh2(const void* p) {
    double2 v;
    asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_coh1(const void* p) {
    double v;
    asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// ---- one KB item: tile (tx, ty) of 64 x 8 points, planes [z0, z1) ------------------------------------
template <bool HAS_DIAG>
__device__ __forceinline__ void kb_item(const KbaArgs& a, int tx, int ty, int z0, int z1, double s, double fa,
                                        double fb, double& acc) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ix = tx * 64 + 2 * lane;
    const int iy = ty * kWarps + warp;
    int iym = iy - 1, iyp = iy + 1;
    bool hym = true, hyp = true;
    if (iym < 0) { if (a.periodic) iym = a.ny - 1; else hym = false; }
    if (iyp >= a.ny) { if (a.periodic) iyp = 0; else hyp = false; }
    const int64_t off_c = (int64_t)iy * a.nx + ix;
    const int64_t off_m = (int64_t)iym * a.nx + ix;
    const int64_t off_p = (int64_t)iyp * a.nx + ix;
    const bool edge_l = (lane == 0), edge_r = (lane == 31);
    int ixl = ix - 1, ixr = ix + 2;
    bool hxl = true, hxr = true;
    if (ixl < 0) { if (a.periodic) ixl = a.nx - 1; else hxl = false; }
    if (ixr >= a.nx) { if (a.periodic) ixr -= a.nx; else hxr = false; }
    const int64_t off_l = (int64_t)iy * a.nx + ixl;
    const int64_t off_r = (int64_t)iy * a.nx + ixr;

    const double* pc = a.x + (int64_t)z0 * a.plane;
    double2 vm = make_double2(0.0, 0.0), vc, vp;
    vc = ld_cached2(pc + off_c);
    {
        const double* pm = (z0 > 0) ? (pc - a.plane) : (a.periodic ? a.x + (int64_t)(a.nz - 1) * a.plane : nullptr);
        if (pm) vm = ld_cached2(pm + off_c);
    }
#pragma unroll 1
    for (int z = z0; z < z1; ++z) {
        vp = make_double2(0.0, 0.0);
        {
            const double* pp = (z + 1 < a.nz) ? (pc + a.plane) : (a.periodic ? a.x : nullptr);
            if (pp) vp = ld_cached2(pp + off_c);
        }
        double2 bv = make_double2(0.0, 0.0);
        if (a.b) bv = ld_stream2(a.b + (int64_t)z * a.plane + off_c);
        double2 ym = make_double2(0.0, 0.0), yp = ym;
        if (hym) ym = ld_cached2(pc + off_m);
        if (hyp) yp = ld_cached2(pc + off_p);
        double left = __shfl_up_sync(0xffffffffu, vc.y, 1);
        double right = __shfl_down_sync(0xffffffffu, vc.x, 1);
        if (edge_l) left = hxl ? __ldg(pc + off_l) : 0.0;
        if (edge_r) right = hxr ? __ldg(pc + off_r) : 0.0;
        double2 dg = make_double2(0.0, 0.0);
        if (HAS_DIAG) dg = ld_cached2(a.diag + (int64_t)z * a.plane + off_c);
        double2 out;
        {   // ascending column order of the sorted CSR row (interior points), as in stencil.cu
            double r = a.oz * vm.x;
            r = fma(a.oy, ym.x, r);
            r = fma(a.ox, left, r);
            r = fma(a.c + dg.x, vc.x, r);
            r = fma(a.ox, vc.y, r);
            r = fma(a.oy, yp.x, r);
            r = fma(a.oz, vp.x, r);
            r *= s;
            r = fma(-fa, vc.x, r);
            if (a.b) r = fma(-fb, bv.x, r);
            acc = fma(r, r, acc);
            out.x = r;
        }
        {
            double r = a.oz * vm.y;
            r = fma(a.oy, ym.y, r);
            r = fma(a.ox, vc.x, r);
            r = fma(a.c + dg.y, vc.y, r);
            r = fma(a.ox, right, r);
            r = fma(a.oy, yp.y, r);
            r = fma(a.oz, vp.y, r);
            r *= s;
            r = fma(-fa, vc.y, r);
            if (a.b) r = fma(-fb, bv.y, r);
            acc = fma(r, r, acc);
            out.y = r;
        }
        // default store policy: the line should still be in L2 when the KA item of this chunk reads it
        *reinterpret_cast<double2*>(a.y + (int64_t)z * a.plane + off_c) = out;
        vm = vc;
        vc = vp;
        pc += a.plane;
    }
}

// ---- one KA item: tile (tx, ty) of 64 x 16 points of the NEW vector, planes [z0, z1) -------------------
// (symmetric form with forward neighbours, three planes in registers: see stencil_alpha_fast_kernel)
template <bool HAS_DIAG>
__device__ __forceinline__ void ka_item(const KbaArgs& a, int tx, int ty, int z0, int z1, double& acc) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double cc = a.c, ox2 = 2.0 * a.ox, oy2 = 2.0 * a.oy, oz2 = 2.0 * a.oz;
    const uint32_t plane_b = (uint32_t)(a.plane * 8);
    const int ix = tx * 64 + 2 * lane;
    const int iy0 = ty * 16 + 2 * warp;
    int iyu = iy0 + 2;
    double oy2u = oy2;
    if (iyu >= a.ny) { if (a.periodic) iyu -= a.ny; else { iyu = iy0 + 1; oy2u = 0.0; } }
    const bool edge_r = (lane == 31);
    int ixr = ix + 2;
    double ox2r = ox2;
    if (ixr >= a.nx) { if (a.periodic) ixr -= a.nx; else { ixr = ix; ox2r = 0.0; } }
    const uint32_t o0 = (uint32_t)(iy0 * a.nx + ix) * 8u;
    const uint32_t o1 = o0 + (uint32_t)a.nx * 8u;
    const uint32_t ou = (uint32_t)(iyu * a.nx + ix) * 8u;
    const uint32_t oe0 = (uint32_t)(iy0 * a.nx + ixr) * 8u;
    const uint32_t oe1 = oe0 + (uint32_t)a.nx * 8u;
    const char* base = reinterpret_cast<const char*>(a.y);
    const char* pl = base + (int64_t)z0 * plane_b;
    const char* pd = HAS_DIAG ? reinterpret_cast<const char*>(a.diag) + (int64_t)z0 * plane_b : nullptr;
    const double2 zz = make_double2(0.0, 0.0);
    double2 c0 = zz, c1 = zz, cu = zz, n0 = zz, n1 = zz, nu = zz, m0 = zz, m1 = zz, mu = zz;
    double cr0 = 0.0, cr1 = 0.0, nr0 = 0.0, nr1 = 0.0, mr0 = 0.0, mr1 = 0.0;
    auto loadp = [&](const char* p, double2& q0, double2& q1, double2& qu, double& e0, double& e1) {
        q0 = ld_coh2(p + o0);
        q1 = ld_coh2(p + o1);
        qu = ld_coh2(p + ou);
        if (edge_r) {
            e0 = ld_coh1(p + oe0);
            e1 = ld_coh1(p + oe1);
        }
    };
    auto reduce_plane = [&](double ozn) {
        double d00 = cc, d01 = cc, d10 = cc, d11 = cc;
        if (HAS_DIAG) {
            const double2 e0 = ld_cached2(reinterpret_cast<const double*>(pd + o0));
            const double2 e1 = ld_cached2(reinterpret_cast<const double*>(pd + o1));
            d00 += e0.x; d01 += e0.y; d10 += e1.x; d11 += e1.y;
            pd += plane_b;
        }
        double r0 = __shfl_down_sync(0xffffffffu, c0.x, 1);
        double r1 = __shfl_down_sync(0xffffffffu, c1.x, 1);
        if (edge_r) { r0 = cr0; r1 = cr1; }
        const double t00 = fma(ozn, n0.x, fma(oy2, c1.x, fma(ox2, c0.y, d00 * c0.x)));
        const double t01 = fma(ozn, n0.y, fma(oy2, c1.y, fma(ox2r, r0, d01 * c0.y)));
        const double t10 = fma(ozn, n1.x, fma(oy2u, cu.x, fma(ox2, c1.y, d10 * c1.x)));
        const double t11 = fma(ozn, n1.y, fma(oy2u, cu.y, fma(ox2r, r1, d11 * c1.y)));
        acc = fma(c0.x, t00, acc);
        acc = fma(c0.y, t01, acc);
        acc = fma(c1.x, t10, acc);
        acc = fma(c1.y, t11, acc);
    };
    auto rotate = [&]() {
        c0 = n0; c1 = n1; cu = nu; cr0 = nr0; cr1 = nr1;
        n0 = m0; n1 = m1; nu = mu; nr0 = mr0; nr1 = mr1;
    };
    const char* pz1 = (z1 < a.nz) ? base + (int64_t)z1 * plane_b : (a.periodic ? base : nullptr);
    const double oz_last = pz1 ? oz2 : 0.0;
    if (!pz1) pz1 = pl;
    loadp(pl, c0, c1, cu, cr0, cr1);
    loadp((z0 + 1 < z1) ? pl + plane_b : pz1, n0, n1, nu, nr0, nr1);
    const char* pm = pl + 2 * (int64_t)plane_b;
    int z = z0;
#pragma unroll 2
    for (; z < z1 - 2; ++z) {
        loadp(pm, m0, m1, mu, mr0, mr1);
        pm += plane_b;
        reduce_plane(oz2);
        rotate();
    }
    if (z < z1 - 1) {
        loadp(pz1, m0, m1, mu, mr0, mr1);
        reduce_plane(oz2);
        rotate();
    }
    reduce_plane(oz_last);
}

#ifndef LZ_KBA_MINBLOCKS
#define LZ_KBA_MINBLOCKS 4
#endif
template <bool HAS_DIAG>
__global__ void __launch_bounds__(kThreads, LZ_KBA_MINBLOCKS)
stencil_kba_kernel(const KbaArgs a) {
    pdl_prologue();
    __shared__ double red[kWarps];
    const double s = a.scale ? __ldg(a.scale) : 1.0;
    const double fa = (a.ca ? __ldg(a.ca) : 1.0) * (a.sa ? __ldg(a.sa) : 1.0);
    const double fb = a.b ? (a.cb ? __ldg(a.cb) : 1.0) * (a.sb ? __ldg(a.sb) : 1.0) : 0.0;
    const int TB = a.tb_x * a.tb_y, TA = a.ta_x * a.ta_y;
    const int64_t nq = (int64_t)a.chunks * (TB + TA);
    double acc_b = 0.0, acc_a = 0.0;
    for (int64_t q = blockIdx.x; q < nq; q += gridDim.x) {
        // queue position -> (kind, chunk, tile):  KB(0) | KB(c) KA(c-1), c = 1 .. C-1 | KA(C-1)
        bool is_kb;
        int cz, t;
        if (q < TB) { is_kb = true; cz = 0; t = (int)q; }
        else {
            const int64_t q1 = q - TB;
            const int pair = (int)(q1 / (TB + TA));
            const int rem = (int)(q1 % (TB + TA));
            if (pair < a.chunks - 1) {
                if (rem < TB) { is_kb = true; cz = pair + 1; t = rem; }
                else { is_kb = false; cz = pair; t = rem - TB; }
            } else { is_kb = false; cz = a.chunks - 1; t = rem; }
        }
        const int z0 = cz * a.zc, z1 = min(z0 + a.zc, a.nz);
        if (is_kb) {
            kb_item<HAS_DIAG>(a, t % a.tb_x, t / a.tb_x, z0, z1, s, fa, fb, acc_b);
            __syncthreads();                              // every store of this item has been issued
            if (threadIdx.x == 0) {
                __threadfence();                          // ... and is visible before the count goes up
                atomicAdd(a.done + cz, 1);
            }
        } else {
            if (threadIdx.x == 0) {
                const int cn = (cz + 1 < a.chunks) ? cz + 1 : 0;      // plane above the chunk (wrap: chunk 0, long done)
                while (ld_acquire_gpu(a.done + cz) < TB) __nanosleep(32);
                while (ld_acquire_gpu(a.done + cn) < TB) __nanosleep(32);
            }
            __syncthreads();
            ka_item<HAS_DIAG>(a, t % a.ta_x, t / a.ta_x, z0, z1, acc_a);
        }
    }
    const double ta = block_sum(acc_a, red);
    if (threadIdx.x == 0) a.alpha_partials[blockIdx.x] = ta;
    const double tb = block_sum(acc_b, red);
    if (threadIdx.x == 0) a.partials[blockIdx.x] = tb;
    // tail: the last CTA to finish takes beta (+ omega), then alpha (needs the scale beta just produced)
    __shared__ int s_last;
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int k = atomicAdd(a.fin_beta.ticket, 1u);
        s_last = (k == gridDim.x - 1);
        if (s_last) *a.fin_beta.ticket = 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int c = threadIdx.x; c < a.chunks; c += kThreads) a.done[c] = 0;          // ready for the next launch
    fin_scalar_body(a.fin_beta.op, a.fin_beta.st, a.fin_beta.pc, 0, LZ_XCHG_FUSED, a.partials, (int)gridDim.x, red);
    __syncthreads();
    if (a.fin_alpha.kind != FIN_NONE)
        fin_scalar_body(a.fin_alpha, a.fin_beta.st, a.fin_beta.pc, 0, LZ_XCHG_FUSED, a.alpha_partials, (int)gridDim.x, red);
}

bool kba_step_supported(const lz_op* op, const double* x, const double* b, const double* out) {
    static const bool off = []() { const char* e = getenv("LZ_KBA"); return e && e[0] == '0'; }();
    if (off || op->kind != LZ_OP_STENCIL || op->st.points != 7 || op->st.sharded) return false;
    const lz_stencil& st = op->st;
    return st.dim == 3 && st.offx != 0.0 && st.offy != 0.0 && st.offz != 0.0 && st.nx % 64 == 0 && st.ny % 16 == 0 &&
           st.nz >= 2 && st.nx * st.ny * 8 < ((int64_t)1 << 32) &&
           ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(out) |
             reinterpret_cast<uintptr_t>(st.diag)) & 15) == 0;
}

// KB of the step that turns x = row j into out = row j+1, with beta_{j+1} (fin_beta) and alpha_{j+1} (fin_alpha,
// FIN_NONE to skip) taken in the tail.  `want_alpha` false: the KA items are left out (last step of a run).
int launch_kba_step(lz_op* op, const double* x, const double* scale_dev, const StencilUpdate* upd, double* out,
                    const FinTail* fin_beta, const FinOp* fin_alpha, int* nparts) {
    const lz_stencil& st = op->st;
    lz_ctx* ctx = op->ctx;
    KbaArgs a{};
    a.nx = (int)st.nx; a.ny = (int)st.ny; a.nz = (int)st.nz;
    a.periodic = (st.bc == LZ_BC_PERIODIC);
    a.plane = st.nx * st.ny;
    a.c = st.center; a.ox = st.offx; a.oy = st.offy; a.oz = st.offz;
    a.x = x; a.b = upd->b; a.y = out; a.diag = st.diag; a.scale = scale_dev;
    a.ca = upd->ca; a.sa = upd->sa; a.cb = upd->cb; a.sb = upd->sb;
    a.tb_x = (int)(st.nx / 64); a.tb_y = (int)(st.ny / kWarps);
    a.ta_x = (int)(st.nx / 64); a.ta_y = (int)(st.ny / 16);
    // chunk length: the new vector must still be in L2 when its KA items come round - about two chunks of
    // output plus the two input streams of one chunk pass through in between: keep that under ~48 MB
    static const int zc_env = []() { const char* e = getenv("LZ_KBA_ZC"); return e ? atoi(e) : 0; }();
    const double plane_mb = (double)a.plane * 8.0 / 1048576.0;
    int zc = zc_env > 0 ? zc_env : (int)(12.0 / plane_mb);
    zc = std::max(2, std::min(zc, (int)st.nz));
    a.zc = zc;
    a.chunks = (int)((st.nz + zc - 1) / zc);
    LZ_REQUIRE(a.chunks <= 4096, "launch_kba_step: too many z-chunks");
    a.done = ctx->kba_done;
    a.partials = ctx->partials;
    a.alpha_partials = ctx->partials + kMaxPartials;
    a.fin_beta = *fin_beta;
    if (fin_alpha) a.fin_alpha = *fin_alpha;
    if (!fin_alpha) a.ta_x = a.ta_y = 0;                  // no KA items
    const void* fn = st.diag ? (const void*)stencil_kba_kernel<true> : (const void*)stencil_kba_kernel<false>;
    int per_sm = 0;
    LZ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kThreads, 0));
    if (per_sm < 1) per_sm = 1;
    const int64_t nq = (int64_t)a.chunks * ((int64_t)a.tb_x * a.tb_y + (int64_t)a.ta_x * a.ta_y);
    // every CTA must be resident (items wait for items of other CTAs)
    const int grid = (int)std::min<int64_t>(nq, std::min<int64_t>((int64_t)ctx->sms * per_sm, kMaxPartials));
    void* args[] = {(void*)&a};
    LZ_CUDA(launch_fn(fn, dim3(grid), dim3(kThreads), 0, ctx->stream, args));
    if (nparts) *nparts = grid;
    return LZ_OK;
}

}  // namespace lz
