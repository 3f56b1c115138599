// K1b: matrix-free 27-point stencil apply with the Lanczos alpha dot fused in.
//
// The reference's default Laplacian is the 27-point one (Hamiltonian.create_sparse_T(points="27"),
// Hamiltonian.py:24-25,102-128: weights by the number of non-zero offsets - centre, face, edge,
// corner - periodic wrap :106-111); 3Ddeuteron.py:76 uses it.  This kernel applies
//     (H x)_i = (w0 + diag_i) x_i + w1 * sum_faces x + w2 * sum_edges x + w3 * sum_corners x
// with the same index map and boundary handling as K1 (stencil.cu).
//
// Marching in z like K1, but each xy-plane is first condensed, per point, into two in-plane sums
//     P0 = w0 b_c + w1 a_c + w1 (b_m + b_p) + w2 (a_m + a_p)     (weight pattern of the plane dz = 0)
//     P1 = w1 b_c + w2 a_c + w2 (b_m + b_p) + w3 (a_m + a_p)     (weight pattern of the planes dz = +-1)
// where b_r is the value on row r in {y-1, y, y+1} and a_r the sum of its two x neighbours; then
//     out(z) = P0(z) + P1(z-1) + P1(z+1).
// Every plane is loaded once per CTA row-triple (three 128-bit loads per thread, two of them L1
// hits) and its two sums ride in registers for the next two steps: 16 B of HBM traffic per point,
// the same as the 7-point kernel.
#include "internal.h"

namespace lz {

struct Stencil27Args {
    int nx, ny, nz;
    int periodic;
    int64_t plane;
    double w0, w1, w2, w3;
    const double* x;
    double* y;
    const double* diag;
    const double* zlo;
    const double* zhi;
    const double* scale;
    const double* b;      // MODE 2 (KB, see stencil.cu): v_{j-1} and the update coefficients
    const double* ca;
    const double* sa;
    const double* cb;
    const double* sb;
    double* halo_lo;      // MODE 2, sharded: plane 0 of `out` also goes to the lower neighbour's ghost buffer,
    double* halo_hi;      // plane nz-1 to the upper neighbour's (NVLink peer stores, as K3 does)
    const int* skip;
    double* partials;
    int tiles_x, tiles_y, chunks_z, zc;
    int64_t nitems;
    FinTail fin;          // what the last CTA does with the summed partials (fin.cuh)
};

template <int VEC>
__device__ __forceinline__ void load_row(const double* p, double (&v)[VEC]) {
    if constexpr (VEC == 2) {
        const double2 t = __ldg(reinterpret_cast<const double2*>(p));
        v[0] = t.x;
        v[1] = t.y;
    } else {
        v[0] = __ldg(p);
    }
}

// 5 CTAs/SM (48 registers, a few spilled words) measured 0.81 ms at 512^3 against 1.11 ms for the
// compiler's default 72 registers / 3 CTAs: like K1 this kernel is latency-bound.
// MODE as in stencil.cu: 0 apply + alpha partial, 1 alpha partial only (KA), 2 apply fused with the
// three-term update and the norm partial (KB).
template <int VEC, bool HAS_DIAG, int MODE>
__global__ void __launch_bounds__(kThreads, MODE == 2 ? 4 : 5)
stencil27_apply_dot_kernel(const Stencil27Args a) {
    pdl_prologue();
    if (a.skip && *a.skip == 0) return;
    __shared__ double red[kWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double s = a.scale ? __ldg(a.scale) : 1.0;
    const double fa = MODE == 2 ? (a.ca ? __ldg(a.ca) : 1.0) * (a.sa ? __ldg(a.sa) : 1.0) : 0.0;
    const double fb = (MODE == 2 && a.b) ? (a.cb ? __ldg(a.cb) : 1.0) * (a.sb ? __ldg(a.sb) : 1.0) : 0.0;
    constexpr int TX = 32 * VEC;
    double acc_alpha = 0.0;

    for (int64_t item = blockIdx.x; item < a.nitems; item += gridDim.x) {
        const int tx = (int)(item % a.tiles_x);
        const int64_t t = item / a.tiles_x;
        const int ty = (int)(t % a.tiles_y);
        const int cz = (int)(t / a.tiles_y);
        const int ix = tx * TX + VEC * lane;
        const int iy = ty * kWarps + warp;
        const bool act = (ix < a.nx) && (iy < a.ny);
        const int z0 = cz * a.zc;
        const int z1 = min(z0 + a.zc, a.nz);

        // the three rows y-1, y, y+1 (offsets inside a plane) and the two x neighbours off the warp
        int rows[3] = {iy - 1, iy, iy + 1};
        bool rok[3] = {act, act, act};
        if (rows[0] < 0) { if (a.periodic) rows[0] = a.ny - 1; else rok[0] = false; }
        if (rows[2] >= a.ny) { if (a.periodic) rows[2] = 0; else rok[2] = false; }
        const bool edge_l = (lane == 0);
        const bool edge_r = (lane == 31) || (ix + VEC >= a.nx);
        int ixl = ix - 1, ixr = ix + VEC;
        bool hxl = true, hxr = true;
        if (ixl < 0) { if (a.periodic) ixl = a.nx - 1; else hxl = false; }
        if (ixr >= a.nx) { if (a.periodic) ixr -= a.nx; else hxr = false; }

        // condense plane `pl` (nullptr: outside a Dirichlet wall) into P0, P1 and the centre values
        auto condense = [&](const double* pl, double (&P0)[VEC], double (&P1)[VEC], double (&ctr)[VEC]) {
            double am[VEC], ac[VEC], ap[VEC], bm[VEC], bc[VEC], bp[VEC];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                double v[VEC];
#pragma unroll
                for (int e = 0; e < VEC; ++e) v[e] = 0.0;
                const bool ok = rok[r] && (pl != nullptr);
                const double* prow = pl ? pl + (int64_t)rows[r] * a.nx : nullptr;
                if (ok) load_row<VEC>(prow + ix, v);
                double left = __shfl_up_sync(0xffffffffu, v[VEC - 1], 1);
                double right = __shfl_down_sync(0xffffffffu, v[0], 1);
                if (edge_l) left = (ok && hxl) ? __ldg(prow + ixl) : 0.0;
                if (edge_r) right = (ok && hxr) ? __ldg(prow + ixr) : 0.0;
                double asum[VEC];
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    const double xl = (e == 0) ? left : v[e - 1];
                    const double xr = (e == VEC - 1) ? right : v[e + 1];
                    asum[e] = xl + xr;
                }
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    if (r == 0) { am[e] = asum[e]; bm[e] = v[e]; }
                    else if (r == 1) { ac[e] = asum[e]; bc[e] = v[e]; }
                    else { ap[e] = asum[e]; bp[e] = v[e]; }
                }
            }
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                const double bs = bm[e] + bp[e], as = am[e] + ap[e];
                P0[e] = fma(a.w2, as, fma(a.w1, bs, fma(a.w1, ac[e], a.w0 * bc[e])));
                P1[e] = fma(a.w3, as, fma(a.w2, bs, fma(a.w2, ac[e], a.w1 * bc[e])));
                ctr[e] = bc[e];
            }
        };
        auto plane_ptr = [&](int z) -> const double* {
            if (z < 0) return a.zlo;
            if (z >= a.nz) return a.zhi;
            return a.x + (int64_t)z * a.plane;
        };

        double P1m[VEC], P0c[VEC], P1c[VEC], cc[VEC], P0p[VEC], P1p[VEC], cp[VEC];
        condense(plane_ptr(z0 - 1), P0p, P1m, cp);      // only P1 of plane z0-1 is needed
        condense(plane_ptr(z0), P0c, P1c, cc);
        for (int z = z0; z < z1; ++z) {
            condense(plane_ptr(z + 1), P0p, P1p, cp);
            if (act) {
                const int64_t at = (int64_t)z * a.plane + (int64_t)iy * a.nx + ix;
                double dg[VEC];
#pragma unroll
                for (int e = 0; e < VEC; ++e) dg[e] = 0.0;
                if (HAS_DIAG) load_row<VEC>(a.diag + at, dg);
                double bv[VEC];
#pragma unroll
                for (int e = 0; e < VEC; ++e) bv[e] = 0.0;
                if (MODE == 2) {
                    if (a.b) {
                        if constexpr (VEC == 2) { const double2 t = ld_stream2(a.b + at); bv[0] = t.x; bv[1] = t.y; }
                        else bv[0] = ld_stream1(a.b + at);
                    }
                }
                double out[VEC];
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    double r = (P1m[e] + P1p[e]) + P0c[e];
                    r = fma(dg[e], cc[e], r);
                    r *= s;
                    if (MODE == 2) {
                        r = fma(-fa, cc[e], r);
                        if (a.b) r = fma(-fb, bv[e], r);
                        acc_alpha = fma(r, r, acc_alpha);
                    } else {
                        acc_alpha = fma(r, s * cc[e], acc_alpha);
                    }
                    out[e] = r;
                }
                if (MODE != 1) {
                    if constexpr (VEC == 2) st_stream2(a.y + at, make_double2(out[0], out[1]));
                    else st_stream1(a.y + at, out[0]);
                }
                if (MODE == 2) {
                    const int64_t inpl = (int64_t)iy * a.nx + ix;
                    if (a.halo_lo && z == 0) {
                        if constexpr (VEC == 2) *reinterpret_cast<double2*>(a.halo_lo + inpl) = make_double2(out[0], out[1]);
                        else a.halo_lo[inpl] = out[0];
                    }
                    if (a.halo_hi && z == a.nz - 1) {
                        if constexpr (VEC == 2) *reinterpret_cast<double2*>(a.halo_hi + inpl) = make_double2(out[0], out[1]);
                        else a.halo_hi[inpl] = out[0];
                    }
                }
            }
#pragma unroll
            for (int e = 0; e < VEC; ++e) { P1m[e] = P1c[e]; P0c[e] = P0p[e]; P1c[e] = P1p[e]; cc[e] = cp[e]; }
        }
    }
    const double tot = block_sum(acc_alpha, red);
    if (threadIdx.x == 0 && a.partials) a.partials[blockIdx.x] = tot;
    fin_tail(a.fin, a.partials, red);
}

template <int VEC, int MODE>
static const void* pick27(bool has_diag) {
    return has_diag ? (const void*)stencil27_apply_dot_kernel<VEC, true, MODE>
                    : (const void*)stencil27_apply_dot_kernel<VEC, false, MODE>;
}

static int launch_stencil27(lz_op* op, int mode, const double* x, const double* scale_dev, double* y,
                            const StencilUpdate* upd, double* partials, int* nparts, const int* flag_dev,
                            const FinTail* fin);

int launch_stencil27_apply_dot(lz_op* op, const double* x, const double* scale_dev, double* y,
                               double* partials, int* nparts, const int* flag_dev, const FinTail* fin) {
    return launch_stencil27(op, y ? 0 : 1, x, scale_dev, y, nullptr, partials, nparts, flag_dev, fin);
}

int launch_stencil27_update_norm(lz_op* op, const double* x, const double* scale_dev, const StencilUpdate* upd,
                                 double* out, double* partials, int* nparts, const FinTail* fin) {
    return launch_stencil27(op, 2, x, scale_dev, out, upd, partials, nparts, nullptr, fin);
}

static int launch_stencil27(lz_op* op, int mode, const double* x, const double* scale_dev, double* y,
                            const StencilUpdate* upd, double* partials, int* nparts, const int* flag_dev,
                            const FinTail* fin) {
    const lz_stencil& st = op->st;
    lz_ctx* ctx = op->ctx;
    Stencil27Args a;
    a.nx = (int)st.nx; a.ny = (int)st.ny; a.nz = (int)st.nz;
    a.periodic = (st.bc == LZ_BC_PERIODIC);
    a.plane = st.nx * st.ny;
    a.w0 = st.w27[0]; a.w1 = st.w27[1]; a.w2 = st.w27[2]; a.w3 = st.w27[3];
    a.x = x; a.y = y; a.diag = st.diag;
    a.scale = scale_dev; a.partials = partials; a.skip = flag_dev;
    a.b = upd ? upd->b : nullptr;
    a.ca = upd ? upd->ca : nullptr; a.sa = upd ? upd->sa : nullptr;
    a.cb = upd ? upd->cb : nullptr; a.sb = upd ? upd->sb : nullptr;
    a.halo_lo = upd ? upd->halo.lo_dst : nullptr;
    a.halo_hi = upd ? upd->halo.hi_dst : nullptr;
    if (fin) a.fin = *fin;
    if (st.sharded) { a.zlo = st.ghost_lo; a.zhi = st.ghost_hi; }
    else if (a.periodic) { a.zlo = x + (st.nz - 1) * a.plane; a.zhi = x; }
    else { a.zlo = nullptr; a.zhi = nullptr; }
    const bool aligned = ((st.nx & 1) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                         ((reinterpret_cast<uintptr_t>(y) & 15) == 0) && ((reinterpret_cast<uintptr_t>(a.b) & 15) == 0) &&
                         (((reinterpret_cast<uintptr_t>(a.halo_lo) | reinterpret_cast<uintptr_t>(a.halo_hi)) & 15) == 0) &&
                         ((reinterpret_cast<uintptr_t>(st.diag) & 15) == 0) &&
                         (!st.sharded || (((reinterpret_cast<uintptr_t>(st.ghost_lo) |
                                            reinterpret_cast<uintptr_t>(st.ghost_hi)) & 15) == 0));
    const int vec = aligned ? 2 : 1;
    const int TX = 32 * vec;
    a.tiles_x = (int)((st.nx + TX - 1) / TX);
    a.tiles_y = (int)((st.ny + kWarps - 1) / kWarps);
    const void* fn;
    const bool hd = st.diag != nullptr;
    if (vec == 2) fn = mode == 2 ? pick27<2, 2>(hd) : mode == 1 ? pick27<2, 1>(hd) : pick27<2, 0>(hd);
    else fn = mode == 2 ? pick27<1, 2>(hd) : mode == 1 ? pick27<1, 1>(hd) : pick27<1, 0>(hd);
    int per_sm = 0;
    LZ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kThreads, 0));
    if (per_sm < 1) per_sm = 1;
    const int64_t gmax = std::min<int64_t>((int64_t)ctx->sms * per_sm, kMaxPartials);
    const int64_t tiles = (int64_t)a.tiles_x * a.tiles_y;
    int best_chunks = 1;
    double best_cost = 1e300;
    const int max_chunks = (int)std::min<int64_t>(st.nz, 4096);
    for (int ch = 1; ch <= max_chunks; ++ch) {
        const int zc = (int)((st.nz + ch - 1) / ch);
        const int chunks = (int)((st.nz + zc - 1) / zc);
        const int64_t items = tiles * chunks;
        const int64_t g = std::min<int64_t>(items, gmax);
        const int64_t rounds = (items + g - 1) / g;
        const double cost = (double)rounds * (zc + 2.0) / ((double)st.nz * tiles / gmax);
        if (cost < best_cost - 1e-12) { best_cost = cost; best_chunks = chunks; }
        if (zc <= 8) break;
    }
    a.zc = (int)((st.nz + best_chunks - 1) / best_chunks);
    a.chunks_z = (int)((st.nz + a.zc - 1) / a.zc);
    a.nitems = tiles * a.chunks_z;
    const int grid = (int)std::min<int64_t>(a.nitems, gmax);
    void* args[] = {(void*)&a};
    LZ_CUDA(launch_fn(fn, dim3(grid), dim3(kThreads), 0, ctx->stream, args));
    if (nparts) *nparts = grid;
    return LZ_OK;
}

}  // namespace lz
