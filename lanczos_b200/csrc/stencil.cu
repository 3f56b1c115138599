// K1: matrix-free 3-/5-/7-point stencil apply with the Lanczos alpha dot fused in.
//
//   MODE 0 (K1):  y = s * (H x),   partial[cta] = sum_i y_i * (s * x_i)
//   MODE 1 (KA):  the same partial, y is NOT written                         (8*N bytes)
//   MODE 2 (KB):  out = s * (H x) - fa * x - fb * b,  partial[cta] = sum out^2   (24*N bytes)
//
// KA + KB are the "recompute" Lanczos step: a matrix-free operator costs no HBM traffic of its own,
// so w = H v_j is evaluated twice - once to reduce alpha_j, once inside the three-term update -
// instead of being written (8*N) and read back (8*N): 32*N bytes per step instead of 48*N
// (Lanczos.py:116-119 folded into two passes over v_j).  The arithmetic of w and of the update is
// operation for operation that of K1 followed by K3 (vecops.cu).
//
// replaces `r = H*V[j]` + `alpha[j] = np.dot(V[j], r)` (Lanczos.py:116,118) for operators
// with the reference's structured-grid pattern (Hamiltonian.py:73-99).  `s` is the lazy
// normalisation factor 1/beta_j of the un-normalised Lanczos vector stored in the basis.
//
// Layout/algorithm: a CTA of 8 warps owns an xy-tile of (32*VEC) x 8 points and marches
// through a chunk of z-planes keeping the z-1 / z / z+1 values of its own column in
// registers (2.5-D register pipeline).  Per plane a thread issues one 128-bit load for
// the new z+1 values and two for the y-1 / y+1 rows (L1 hits except on the two tile-edge
// rows); x neighbours travel by warp shuffle, only the two edge lanes load them.  HBM
// traffic is the compulsory 8 B read + 8 B write per point (+ 2/zc for the chunk halo).
#include <stdlib.h>
#include "internal.h"

namespace lz {

struct StencilArgs {
    int nx, ny, nz;
    int periodic;
    int64_t plane;
    double c, ox, oy, oz;
    const double* x;
    double* y;
    const double* diag;
    const double* zlo;    // plane below local plane 0 (nullptr: zero)
    const double* zhi;    // plane above local plane nz-1
    const double* scale;  // device scalar, nullptr: 1
    const double* b;      // MODE 2: v_{j-1} (nullable)
    const double* ca;     // MODE 2: device scalars, fa = ca * sa, fb = cb * sb (nullable => 1)
    const double* sa;
    const double* cb;
    const double* sb;
    double* halo_lo;      // MODE 2, sharded: plane 0 of `out` also goes to the lower neighbour's ghost buffer,
    double* halo_hi;      // plane nz-1 to the upper neighbour's (NVLink peer stores, as K3 does)
    const int* skip;      // device flag, nullptr or *skip != 0: run
    double* partials;
    double* alpha_partials;   // KB with ALPHA: partial of out . H out over the edges inside the CTA's tiles
    int tiles_x, tiles_y, chunks_z, zc;
    int64_t nitems;
    FinTail fin;          // what the last CTA does with the summed partials (fin.cuh)
};

template <int VEC>
__device__ __forceinline__ void load_vec(const double* p, double (&v)[VEC]) {
    if constexpr (VEC == 2) {
        double2 t = __ldg(reinterpret_cast<const double2*>(p));
        v[0] = t.x;
        v[1] = t.y;
    } else {
        v[0] = __ldg(p);
    }
}

template <int VEC>
__device__ __forceinline__ void zero_vec(double (&v)[VEC]) {
#pragma unroll
    for (int e = 0; e < VEC; ++e) v[e] = 0.0;
}

// Resident CTAs per SM the register budget is cut for (the kernels are latency-bound: occupancy matters
// more than unrolling).  K1 / KA: 6 x 40 registers; KB: 4 x 64 registers, no spills.
#ifndef LZ_K1_MINBLOCKS
#define LZ_K1_MINBLOCKS 6
#endif
#ifndef LZ_KA_MINBLOCKS
#define LZ_KA_MINBLOCKS 6
#endif
#ifndef LZ_KB_MINBLOCKS
#define LZ_KB_MINBLOCKS 4
#endif
// ALPHA (MODE 2, grids made of whole 64 x 8 tiles): alpha of the vector being produced,
//     out . H out = sum_i (c + d_i) out_i^2 + 2 sum_edges o out_i out_j,
// is accumulated while `out` is in registers, for every edge whose two ends this CTA produces: the edge
// inside a thread's pair and to the next lane (shuffle), to the row of the next warp (one shared-memory
// row exchange per plane), to the previous plane of the z-chunk (two registers).  The edges that leave the
// tile (lane 31 -> next tile, warp 7 -> next tile, last plane of a chunk -> next chunk / wrap / ghost plane)
// are left to stencil_alpha_border_kernel: ~3.5 B per point instead of KA2's 8.
template <int VEC, bool HAS_Y, bool HAS_Z, bool HAS_DIAG, int MODE, bool ALPHA = false>
__global__ void __launch_bounds__(kThreads, MODE == 2 ? LZ_KB_MINBLOCKS : (MODE == 1 ? LZ_KA_MINBLOCKS : LZ_K1_MINBLOCKS))
stencil_apply_dot_kernel(const StencilArgs a) {
    pdl_prologue();
    if (a.skip && *a.skip == 0) return;
    __shared__ double red[kWarps];
    __shared__ double2 xrow[ALPHA ? 2 : 1][ALPHA ? kWarps : 1][ALPHA ? 32 : 1];   // rows of `out`, double-buffered
    double acc2 = 0.0;
    double po[VEC];                      // `out` of the previous plane of this chunk
    zero_vec<VEC>(po);
    int it = 0;                          // running plane count of this CTA: parity of the exchange buffer
    const double ox2 = 2.0 * a.ox, oy2 = 2.0 * a.oy, oz2 = 2.0 * a.oz;
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
        const bool act = (ix < a.nx) && (iy < a.ny);   // nx % VEC == 0 => whole vector in range
        const int z0 = cz * a.zc;
        const int z1 = min(z0 + a.zc, a.nz);

        // y neighbours (row offsets inside a plane)
        int iym = iy - 1, iyp = iy + 1;
        bool hym = true, hyp = true;
        if (iym < 0) { if (a.periodic) iym = a.ny - 1; else hym = false; }
        if (iyp >= a.ny) { if (a.periodic) iyp = 0; else hyp = false; }
        const int64_t off_c = (int64_t)iy * a.nx + ix;
        const int64_t off_m = (int64_t)iym * a.nx + ix;
        const int64_t off_p = (int64_t)iyp * a.nx + ix;
        // x neighbours that cannot come from a shuffle
        const bool edge_l = (lane == 0);
        const bool edge_r = (lane == 31) || (ix + VEC >= a.nx);
        int ixl = ix - 1, ixr = ix + VEC;
        bool hxl = true, hxr = true;
        if (ixl < 0) { if (a.periodic) ixl = a.nx - 1; else hxl = false; }
        if (ixr >= a.nx) { if (a.periodic) ixr -= a.nx; else hxr = false; }
        const int64_t off_l = (int64_t)iy * a.nx + ixl;
        const int64_t off_r = (int64_t)iy * a.nx + ixr;

        const double* pc = a.x + (int64_t)z0 * a.plane;
        double vm[VEC], vc[VEC], vp[VEC];
        zero_vec<VEC>(vm);
        zero_vec<VEC>(vc);
        if (act) load_vec<VEC>(pc + off_c, vc);
        if (HAS_Z) {
            const double* pm = (z0 > 0) ? (pc - a.plane) : a.zlo;
            if (act && pm) load_vec<VEC>(pm + off_c, vm);
        }

#pragma unroll 1
        for (int z = z0; z < z1; ++z) {
            zero_vec<VEC>(vp);
            if (HAS_Z) {
                const double* pp = (z + 1 < a.nz) ? (pc + a.plane) : a.zhi;
                if (act && pp) load_vec<VEC>(pp + off_c, vp);
            }
            double bv[VEC];
            zero_vec<VEC>(bv);
            if (MODE == 2) {
                if (act && a.b) {
                    const double* pb = a.b + (int64_t)z * a.plane + off_c;
                    if constexpr (VEC == 2) { const double2 t = ld_stream2(pb); bv[0] = t.x; bv[1] = t.y; }
                    else bv[0] = ld_stream1(pb);
                }
            }
            double ym[VEC], yp[VEC];
            zero_vec<VEC>(ym);
            zero_vec<VEC>(yp);
            if (HAS_Y) {
                if (act && hym) load_vec<VEC>(pc + off_m, ym);
                if (act && hyp) load_vec<VEC>(pc + off_p, yp);
            }
            double left = __shfl_up_sync(0xffffffffu, vc[VEC - 1], 1);
            double right = __shfl_down_sync(0xffffffffu, vc[0], 1);
            if (edge_l) left = (act && hxl) ? __ldg(pc + off_l) : 0.0;
            if (edge_r) right = (act && hxr) ? __ldg(pc + off_r) : 0.0;
            double dg[VEC];
            zero_vec<VEC>(dg);
            if (HAS_DIAG) {
                if (act) load_vec<VEC>(a.diag + (int64_t)z * a.plane + off_c, dg);
            }
            if (act) {
                double out[VEC];
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    const double xl = (e == 0) ? left : vc[e - 1];
                    const double xr = (e == VEC - 1) ? right : vc[e + 1];
                    // ascending column order of the sorted CSR row (interior points)
                    double r = a.oz * vm[e];
                    r = fma(a.oy, ym[e], r);
                    r = fma(a.ox, xl, r);
                    r = fma(a.c + dg[e], vc[e], r);
                    r = fma(a.ox, xr, r);
                    r = fma(a.oy, yp[e], r);
                    r = fma(a.oz, vp[e], r);
                    r *= s;
                    if (MODE == 2) {
                        r = fma(-fa, vc[e], r);             // K3: w - fa * v_j - fb * v_{j-1}
                        if (a.b) r = fma(-fb, bv[e], r);
                        acc_alpha = fma(r, r, acc_alpha);
                    } else {
                        acc_alpha = fma(r, s * vc[e], acc_alpha);
                    }
                    out[e] = r;
                }
                if (MODE != 1) {
                    double* py = a.y + (int64_t)z * a.plane + off_c;
                    if constexpr (VEC == 2) st_stream2(py, make_double2(out[0], out[1]));
                    else st_stream1(py, out[0]);
                }
                if constexpr (ALPHA && VEC == 2) {
                    // every thread of the CTA is active here (whole tiles): barriers are uniform
                    acc2 = fma((a.c + dg[0]) * out[0], out[0], acc2);
                    acc2 = fma((a.c + dg[1]) * out[1], out[1], acc2);
                    acc2 = fma(ox2 * out[0], out[1], acc2);
                    const double nxt = __shfl_down_sync(0xffffffffu, out[0], 1);
                    if (lane < 31) acc2 = fma(ox2 * out[1], nxt, acc2);
                    if (z > z0) {
                        acc2 = fma(oz2 * po[0], out[0], acc2);
                        acc2 = fma(oz2 * po[1], out[1], acc2);
                    }
                    po[0] = out[0];
                    po[1] = out[1];
                    const int buf = it & 1;
                    ++it;
                    xrow[buf][warp][lane] = make_double2(out[0], out[1]);
                    __syncthreads();
                    if (warp < kWarps - 1) {
                        const double2 up = xrow[buf][warp + 1][lane];
                        acc2 = fma(oy2 * out[0], up.x, acc2);
                        acc2 = fma(oy2 * out[1], up.y, acc2);
                    }
                }
                if (MODE == 2) {
                    if (a.halo_lo && z == 0) {
                        if constexpr (VEC == 2) *reinterpret_cast<double2*>(a.halo_lo + off_c) = make_double2(out[0], out[1]);
                        else a.halo_lo[off_c] = out[0];
                    }
                    if (a.halo_hi && z == a.nz - 1) {
                        if constexpr (VEC == 2) *reinterpret_cast<double2*>(a.halo_hi + off_c) = make_double2(out[0], out[1]);
                        else a.halo_hi[off_c] = out[0];
                    }
                }
            }
#pragma unroll
            for (int e = 0; e < VEC; ++e) { vm[e] = vc[e]; vc[e] = vp[e]; }
            pc += a.plane;
        }
    }
    if constexpr (ALPHA) {
        const double t2 = block_sum(acc2, red);
        if (threadIdx.x == 0) a.alpha_partials[blockIdx.x] = t2;
    }
    const double tot = block_sum(acc_alpha, red);
    if (threadIdx.x == 0 && a.partials) a.partials[blockIdx.x] = tot;
    fin_tail(a.fin, a.partials, red);
}

// The edges of out . H out that stencil_apply_dot_kernel<ALPHA> could not reach (see there), for a vector v
// on a grid of whole 64 x 8 tiles walked in z-chunks of zc planes:
//   Y: rows 8t+7 -> 8t+8 (wrap: periodic only)        nz * tiles_y * nx products, two full rows per pair
//   Z: plane (chunk end - 1) -> next plane (or the plane above the slab: wrap / ghost / none)
//   X: columns 64t+63 -> 64t+64 (wrap: periodic only)  nz * ny * tiles_x products
// One flat index space, grid-stride, double2 loads for Y and Z; partials[cta] = 2 * sum o v_i v_j.
struct BorderArgs {
    int nx, ny, nz, periodic;
    int tiles_x, tiles_y, zc, chunks_z;
    int64_t plane;
    double ox2, oy2, oz2;
    const double* v;
    const double* zhi;    // plane above the slab (nullptr: none)
    double* partials;
    FinTail fin;
};

__global__ void __launch_bounds__(kThreads)
stencil_alpha_border_kernel(const BorderArgs a) {
    pdl_prologue();
    __shared__ double red[kWarps];
    const int64_t tid = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * kThreads;
    const int hx = a.nx >> 1;                                    // double2 per row
    const int ty_pairs = (a.periodic || a.tiles_y == 0) ? a.tiles_y : a.tiles_y - 1;   // Dirichlet: no wrap pair
    const int tx_pairs = a.periodic ? a.tiles_x : a.tiles_x - 1;
    const int64_t WY = (int64_t)a.nz * ty_pairs * hx;
    const int zpairs = (a.zhi != nullptr) ? a.chunks_z : a.chunks_z - 1;
    const int64_t WZ = (int64_t)zpairs * (a.plane >> 1);
    const int64_t WX = (int64_t)a.nz * a.ny * tx_pairs;
    double accy = 0.0, accz = 0.0, accx = 0.0;
#pragma unroll 2
    for (int64_t i = tid; i < WY; i += nthr) {
        const int x2 = (int)(i % hx);
        const int64_t t = i / hx;
        const int ty = (int)(t % ty_pairs);
        const int z = (int)(t / ty_pairs);
        const int ya = ty * kWarps + kWarps - 1;
        const int yb = (ya + 1 == a.ny) ? 0 : ya + 1;
        const double* pz = a.v + (int64_t)z * a.plane + 2 * x2;
        const double2 p = ld_stream2(pz + (int64_t)ya * a.nx);
        const double2 q = ld_stream2(pz + (int64_t)yb * a.nx);
        accy = fma(p.x, q.x, accy);
        accy = fma(p.y, q.y, accy);
    }
#pragma unroll 2
    for (int64_t i = tid; i < WZ; i += nthr) {
        const int64_t hp = a.plane >> 1;
        const int64_t e2 = i % hp;
        const int cz = (int)(i / hp);
        const int za = min((cz + 1) * a.zc, a.nz) - 1;
        const double* pa = a.v + (int64_t)za * a.plane + 2 * e2;
        const double* pb = (za + 1 < a.nz) ? pa + a.plane : a.zhi + 2 * e2;
        const double2 p = ld_stream2(pa);
        const double2 q = ld_stream2(pb);
        accz = fma(p.x, q.x, accz);
        accz = fma(p.y, q.y, accz);
    }
#pragma unroll 2
    for (int64_t i = tid; i < WX; i += nthr) {
        const int tx = (int)(i % tx_pairs);
        const int64_t row = i / tx_pairs;                        // y + ny * z
        const int xa = tx * 64 + 63;
        const int xb = (xa + 1 == a.nx) ? 0 : xa + 1;
        const double* pr = a.v + row * a.nx;
        accx = fma(ld_stream1(pr + xa), ld_stream1(pr + xb), accx);
    }
    const double tot = block_sum(fma(a.oy2, accy, fma(a.oz2, accz, a.ox2 * accx)), red);
    if (threadIdx.x == 0) a.partials[blockIdx.x] = tot;
    fin_tail(a.fin, a.partials, red);
}

// KA2: alpha_j alone for a full 3-D 7-point operator, from the symmetric form
//     x.Hx = sum_i (c + d_i) x_i^2 + 2 sum_i x_i (ox x_{i+ex} + oy x_{i+ey} + oz x_{i+ez})
// (every edge is counted once, from its lower end; a Dirichlet wall drops the edge, a periodic one
// wraps it, a slab boundary takes the upper ghost plane - the edge below the slab belongs to the
// lower neighbour's sum).  Only "forward" neighbours are needed: a thread owns 2 x-points on 2
// consecutive rows, so per plane it issues two 128-bit loads of its own points and ONE of the row
// above its pair (K1/KA issue one own load and two neighbour-row loads per 2 points), all for the
// plane after the one being reduced - a full iteration of latency hiding with twice the HBM bytes
// in flight per thread.  Used for KA when nx is even, all three axes couple and the vector is
// 16-byte aligned; otherwise MODE 1 of the general kernel runs.  The value differs from K1's
// sum_i y_i x_i only in rounding.
#ifndef LZ_KA2_MINBLOCKS
#define LZ_KA2_MINBLOCKS 5
#endif
template <bool HAS_DIAG>
__global__ void __launch_bounds__(kThreads, LZ_KA2_MINBLOCKS)
stencil_alpha_kernel(const StencilArgs a) {
    pdl_prologue();
    if (a.skip && *a.skip == 0) return;
    __shared__ double red[kWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double s = a.scale ? __ldg(a.scale) : 1.0;
    constexpr int TX = 64, TY = 2 * kWarps;
    const double ox2 = 2.0 * a.ox, oy2 = 2.0 * a.oy, oz2 = 2.0 * a.oz;
    double acc = 0.0;

    for (int64_t item = blockIdx.x; item < a.nitems; item += gridDim.x) {
        const int tx = (int)(item % a.tiles_x);
        const int64_t t = item / a.tiles_x;
        const int ty = (int)(t % a.tiles_y);
        const int cz = (int)(t / a.tiles_y);
        const int ix = tx * TX + 2 * lane;
        const int iy0 = ty * TY + 2 * warp;                 // rows iy0, iy0 + 1; the row above: iy0 + 2
        const bool act0 = (ix < a.nx) && (iy0 < a.ny);
        const bool act1 = (ix < a.nx) && (iy0 + 1 < a.ny);
        // forward y neighbours: row iy0 + 1 is the thread's own second row (or the wrap of a last odd row)
        int iy1 = iy0 + 1, iyu = iy0 + 2;
        bool h1 = act0, hu = act1;
        if (iy1 >= a.ny) { if (a.periodic) iy1 -= a.ny; else h1 = false; }
        if (iyu >= a.ny) { if (a.periodic) iyu -= a.ny; else hu = false; }
        const bool own1 = act1;                             // row iy0 + 1 exists and is this thread's
        const int64_t off0 = (int64_t)iy0 * a.nx + ix;
        const int64_t off1 = (int64_t)iy1 * a.nx + ix;
        const int64_t offu = (int64_t)iyu * a.nx + ix;
        // forward x neighbour off the warp / off the row
        const bool edge_r = (lane == 31) || (ix + 2 >= a.nx);
        int ixr = ix + 2;
        bool hxr = true;
        if (ixr >= a.nx) { if (a.periodic) ixr -= a.nx; else hxr = false; }
        const int z0 = cz * a.zc;
        const int z1 = min(z0 + a.zc, a.nz);

        auto plane_of = [&](int p) -> const double* {       // p in [z0, z1]; p == nz: the plane above the slab
            return (p < a.nz) ? a.x + (int64_t)p * a.plane : a.zhi;
        };
        double2 c0 = make_double2(0.0, 0.0), c1 = c0, cu = c0;   // plane z: rows iy0, iy0+1 (or its wrap), iy0+2
        double cr0 = 0.0, cr1 = 0.0;                              // right neighbours of the edge lane
        auto fetch = [&](const double* pl, double2& r0, double2& r1, double2& ru, double& e0, double& e1) {
            r0 = r1 = ru = make_double2(0.0, 0.0);
            e0 = e1 = 0.0;
            if (!pl) return;
            if (act0) r0 = ld_cached2(pl + off0);
            if (h1) r1 = ld_cached2(pl + off1);
            if (hu) ru = ld_cached2(pl + offu);
            if (edge_r && hxr) {
                if (act0) e0 = __ldg(pl + (int64_t)iy0 * a.nx + ixr);
                if (own1) e1 = __ldg(pl + (int64_t)iy1 * a.nx + ixr);
            }
        };
        fetch(plane_of(z0), c0, c1, cu, cr0, cr1);
#pragma unroll 1
        for (int z = z0; z < z1; ++z) {
            double2 n0, n1, nu;
            double nr0, nr1;
            fetch(plane_of(z + 1), n0, n1, nu, nr0, nr1);   // consumed next iteration (and as the z+1 term now)
            double2 d0 = make_double2(0.0, 0.0), d1 = d0;
            if (HAS_DIAG) {
                const double* pd = a.diag + (int64_t)z * a.plane;
                if (act0) d0 = ld_stream2(pd + off0);
                if (own1) d1 = ld_stream2(pd + off1);
            }
            double r0 = __shfl_down_sync(0xffffffffu, c0.x, 1);
            double r1 = __shfl_down_sync(0xffffffffu, c1.x, 1);
            if (edge_r) { r0 = cr0; r1 = cr1; }
            // row iy0: up neighbour is c1 (own second row, or the periodic wrap when ny is odd)
            {
                double tx0 = fma(oz2, n0.x, fma(oy2, c1.x, fma(ox2, c0.y, (a.c + d0.x) * c0.x)));
                double tx1 = fma(oz2, n0.y, fma(oy2, c1.y, fma(ox2, r0, (a.c + d0.y) * c0.y)));
                acc = fma(c0.x, tx0, acc);
                acc = fma(c0.y, tx1, acc);
            }
            if (own1) {
                double tx0 = fma(oz2, n1.x, fma(oy2, cu.x, fma(ox2, c1.y, (a.c + d1.x) * c1.x)));
                double tx1 = fma(oz2, n1.y, fma(oy2, cu.y, fma(ox2, r1, (a.c + d1.y) * c1.y)));
                acc = fma(c1.x, tx0, acc);
                acc = fma(c1.y, tx1, acc);
            }
            c0 = n0; c1 = n1; cu = nu; cr0 = nr0; cr1 = nr1;
        }
    }
    const double tot = block_sum(acc * (s * s), red);
    if (threadIdx.x == 0 && a.partials) a.partials[blockIdx.x] = tot;
    fin_tail(a.fin, a.partials, red);
}

// KA2, lean form for grids made of whole tiles (nx % 64 == 0, ny % 16 == 0, plane < 4 GB): every
// thread is active, so nothing is predicated.  Boundary cases are folded into data instead of control
// flow: an absent neighbour (Dirichlet wall, missing plane) is read from a clamped valid address and
// its coupling coefficient is zero; the plane base is a uniform pointer bumped once per plane and the
// per-thread offsets are 32-bit.  With three planes in registers (the loads of plane z + 2 in flight
// while plane z is reduced) it runs at 0.181 ms at 512^3 against 0.215 ms for the general form.
#ifndef LZ_KA2_WX
#define LZ_KA2_WX 1          // warps side by side in x: the CTA tile is (64 * WX) x (16 / WX) points
#endif
#ifndef LZ_KA2F_MINBLOCKS
#define LZ_KA2F_MINBLOCKS 4
#endif
template <bool HAS_DIAG>
__global__ void __launch_bounds__(kThreads, LZ_KA2F_MINBLOCKS)
stencil_alpha_fast_kernel(const StencilArgs a) {
    constexpr int WX = LZ_KA2_WX;
    pdl_prologue();
    if (a.skip && *a.skip == 0) return;
    __shared__ double red[kWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double s = a.scale ? __ldg(a.scale) : 1.0;
    constexpr int TX = 64 * WX, TY = 2 * kWarps / WX;
    const double cc = a.c, ox2 = 2.0 * a.ox, oy2 = 2.0 * a.oy, oz2 = 2.0 * a.oz;
    const uint32_t plane_b = (uint32_t)(a.plane * 8);
    double acc = 0.0;

    for (int64_t item = blockIdx.x; item < a.nitems; item += gridDim.x) {
        const int tx = (int)(item % a.tiles_x);
        const int64_t t = item / a.tiles_x;
        const int ty = (int)(t % a.tiles_y);
        const int cz = (int)(t / a.tiles_y);
        const int ix = tx * TX + 64 * (warp % WX) + 2 * lane;
        const int iy0 = ty * TY + 2 * (warp / WX);
        int iyu = iy0 + 2;
        double oy2u = oy2;                                   // coupling of row iy0 + 1 to the row above the pair
        if (iyu >= a.ny) { if (a.periodic) iyu -= a.ny; else { iyu = iy0 + 1; oy2u = 0.0; } }
        const bool edge_r = (lane == 31);
        int ixr = ix + 2;
        double ox2r = ox2;                                   // coupling of the pair's second point to its right neighbour
        if (ixr >= a.nx) { if (a.periodic) ixr -= a.nx; else { ixr = ix; ox2r = 0.0; } }
        const uint32_t o0 = (uint32_t)(iy0 * a.nx + ix) * 8u;
        const uint32_t o1 = o0 + (uint32_t)a.nx * 8u;
        const uint32_t ou = (uint32_t)(iyu * a.nx + ix) * 8u;
        const uint32_t oe0 = (uint32_t)(iy0 * a.nx + ixr) * 8u;
        const uint32_t oe1 = oe0 + (uint32_t)a.nx * 8u;
        const int z0 = cz * a.zc;
        const int z1 = min(z0 + a.zc, a.nz);
        const char* pl = reinterpret_cast<const char*>(a.x) + (int64_t)z0 * plane_b;
        const char* pd = HAS_DIAG ? reinterpret_cast<const char*>(a.diag) + (int64_t)z0 * plane_b : nullptr;

        // three planes ride in registers: c = plane z (being reduced), n = plane z + 1 (its z-coupling;
        // loaded one iteration ago), m = plane z + 2 (in flight while plane z is reduced)
        const double2 zz = make_double2(0.0, 0.0);
        double2 c0 = zz, c1 = zz, cu = zz, n0 = zz, n1 = zz, nu = zz, m0 = zz, m1 = zz, mu = zz;
        double cr0 = 0.0, cr1 = 0.0, nr0 = 0.0, nr1 = 0.0, mr0 = 0.0, mr1 = 0.0;
        auto loadp = [&](const char* p, double2& q0, double2& q1, double2& qu, double& e0, double& e1) {
            q0 = ld_cached2(reinterpret_cast<const double*>(p + o0));
            q1 = ld_cached2(reinterpret_cast<const double*>(p + o1));
            qu = ld_cached2(reinterpret_cast<const double*>(p + ou));
            if (edge_r) {
                e0 = __ldg(reinterpret_cast<const double*>(p + oe0));
                e1 = __ldg(reinterpret_cast<const double*>(p + oe1));
            }
        };
        auto reduce_plane = [&](double ozn) {
            double d00 = cc, d01 = cc, d10 = cc, d11 = cc;
            if (HAS_DIAG) {
                const double2 e0 = ld_stream2(reinterpret_cast<const double*>(pd + o0));
                const double2 e1 = ld_stream2(reinterpret_cast<const double*>(pd + o1));
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
        // the plane above the chunk: the next chunk's first plane, the plane above the slab (periodic
        // wrap / ghost plane), or absent (then a valid plane is read and its coupling is zero)
        const char* pz1 = (z1 < a.nz) ? reinterpret_cast<const char*>(a.x) + (int64_t)z1 * plane_b
                                      : reinterpret_cast<const char*>(a.zhi);
        const double oz_last = pz1 ? oz2 : 0.0;
        if (!pz1) pz1 = pl;
        loadp(pl, c0, c1, cu, cr0, cr1);
        loadp((z0 + 1 < z1) ? pl + plane_b : pz1, n0, n1, nu, nr0, nr1);
        const char* pm = pl + 2 * (int64_t)plane_b;              // plane z + 2
        int z = z0;
#pragma unroll 3
        for (; z < z1 - 2; ++z) {
            loadp(pm, m0, m1, mu, mr0, mr1);
            pm += plane_b;
            reduce_plane(oz2);
            rotate();
        }
        if (z < z1 - 1) {                                        // z == z1 - 2: plane z + 2 is the one above the chunk
            loadp(pz1, m0, m1, mu, mr0, mr1);
            reduce_plane(oz2);
            rotate();
        }
        reduce_plane(oz_last);                                   // z == z1 - 1
    }
    const double tot = block_sum(acc * (s * s), red);
    if (threadIdx.x == 0 && a.partials) a.partials[blockIdx.x] = tot;
    fin_tail(a.fin, a.partials, red);
}

// (KB variants that were built and measured - cp.async rings, a lean two-row form, TMA-staged tiles - are
// recorded in DESIGN.md section 8; the general kernel above is what runs.)

template <int VEC, bool HAS_Y, bool HAS_Z, int MODE>
static const void* pick_diag(bool has_diag) {
    if (has_diag) return (const void*)stencil_apply_dot_kernel<VEC, HAS_Y, HAS_Z, true, MODE>;
    return (const void*)stencil_apply_dot_kernel<VEC, HAS_Y, HAS_Z, false, MODE>;
}
template <int VEC, int MODE>
static const void* pick_yz(bool has_y, bool has_z, bool has_diag) {
    if (has_y && has_z) return pick_diag<VEC, true, true, MODE>(has_diag);
    if (has_y) return pick_diag<VEC, true, false, MODE>(has_diag);
    if (has_z) return pick_diag<VEC, false, true, MODE>(has_diag);
    return pick_diag<VEC, false, false, MODE>(has_diag);
}
template <int VEC>
static const void* pick_kernel(bool has_y, bool has_z, bool has_diag, int mode) {
    if (mode == 1) return pick_yz<VEC, 1>(has_y, has_z, has_diag);
    if (mode == 2) return pick_yz<VEC, 2>(has_y, has_z, has_diag);
    return pick_yz<VEC, 0>(has_y, has_z, has_diag);
}

static int launch_stencil(lz_op* op, int mode, const double* x, const double* scale_dev, double* y,
                          const StencilUpdate* upd, double* partials, int* nparts, const int* flag_dev,
                          const FinTail* fin);

int launch_stencil_apply_dot(lz_op* op, const double* x, const double* scale_dev, double* y,
                             double* partials, int* nparts, const int* flag_dev, const FinTail* fin) {
    return launch_stencil(op, y ? 0 : 1, x, scale_dev, y, nullptr, partials, nparts, flag_dev, fin);
}

int launch_stencil_update_norm(lz_op* op, const double* x, const double* scale_dev, const StencilUpdate* upd,
                               double* out, double* partials, int* nparts, const FinTail* fin) {
    return launch_stencil(op, 2, x, scale_dev, out, upd, partials, nparts, nullptr, fin);
}

static int launch_stencil(lz_op* op, int mode, const double* x, const double* scale_dev, double* y,
                          const StencilUpdate* upd, double* partials, int* nparts, const int* flag_dev,
                          const FinTail* fin) {
    const lz_stencil& st = op->st;
    lz_ctx* ctx = op->ctx;
    StencilArgs a;
    a.nx = (int)st.nx; a.ny = (int)st.ny; a.nz = (int)st.nz;
    a.periodic = (st.bc == LZ_BC_PERIODIC);
    a.plane = st.nx * st.ny;
    a.c = st.center; a.ox = st.offx; a.oy = st.offy; a.oz = st.offz;
    a.x = x; a.y = y; a.diag = st.diag;
    a.scale = scale_dev; a.partials = partials; a.skip = flag_dev;
    a.b = upd ? upd->b : nullptr;
    a.ca = upd ? upd->ca : nullptr; a.sa = upd ? upd->sa : nullptr;
    a.cb = upd ? upd->cb : nullptr; a.sb = upd ? upd->sb : nullptr;
    a.halo_lo = upd ? upd->halo.lo_dst : nullptr;
    a.halo_hi = upd ? upd->halo.hi_dst : nullptr;
    if (fin) a.fin = *fin;
    a.alpha_partials = (upd && mode == 2) ? upd->alpha_partials : nullptr;
    const bool has_y = (st.offy != 0.0);
    const bool has_z = (st.offz != 0.0);
    if (st.sharded) {
        a.zlo = st.ghost_lo;
        a.zhi = st.ghost_hi;
    } else if (a.periodic) {
        a.zlo = x + (st.nz - 1) * a.plane;
        a.zhi = x;
    } else {
        a.zlo = nullptr;
        a.zhi = nullptr;
    }
    const bool aligned = ((st.nx & 1) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                         ((reinterpret_cast<uintptr_t>(y) & 15) == 0) && ((reinterpret_cast<uintptr_t>(a.b) & 15) == 0) &&
                         (((reinterpret_cast<uintptr_t>(a.halo_lo) | reinterpret_cast<uintptr_t>(a.halo_hi)) & 15) == 0) &&
                         ((reinterpret_cast<uintptr_t>(st.diag) & 15) == 0) &&
                         (!st.sharded || (((reinterpret_cast<uintptr_t>(st.ghost_lo) |
                                            reinterpret_cast<uintptr_t>(st.ghost_hi)) & 15) == 0));
    const int vec = aligned ? 2 : 1;
    const int TX = 32 * vec;
    // KA2: the symmetric-form alpha kernel for full 3-D couplings (2 rows per thread)
    const bool alpha_only = (mode == 1) && aligned && has_y && has_z && (st.offx != 0.0);
    a.tiles_x = (int)((st.nx + TX - 1) / TX);
    a.tiles_y = (int)((st.ny + (alpha_only ? 2 * kWarps : kWarps) - 1) / (alpha_only ? 2 * kWarps : kWarps));
    const void* fn = (vec == 2) ? pick_kernel<2>(has_y, has_z, st.diag != nullptr, mode)
                                : pick_kernel<1>(has_y, has_z, st.diag != nullptr, mode);
    if (alpha_only) {
        const bool whole_tiles = (st.nx % (64 * LZ_KA2_WX) == 0) && (st.ny % (2 * kWarps / LZ_KA2_WX) == 0) &&
                                 (st.nx * st.ny * 8 < (int64_t)1 << 32);
        if (whole_tiles) {
            a.tiles_x = (int)(st.nx / (64 * LZ_KA2_WX));
            a.tiles_y = (int)(st.ny / (2 * kWarps / LZ_KA2_WX));
        }
        if (whole_tiles) fn = st.diag ? (const void*)stencil_alpha_fast_kernel<true> : (const void*)stencil_alpha_fast_kernel<false>;
        else fn = st.diag ? (const void*)stencil_alpha_kernel<true> : (const void*)stencil_alpha_kernel<false>;
    }
    if (a.alpha_partials) {
        if (!(vec == 2 && update_alpha_supported(op, x, y))) {
            set_error("alpha inside KB needs a 7-point operator on whole 64 x 8 tiles and 16-byte aligned vectors");
            return LZ_ERR_UNSUPPORTED;
        }
        fn = st.diag ? (const void*)stencil_apply_dot_kernel<2, true, true, true, 2, true>
                     : (const void*)stencil_apply_dot_kernel<2, true, true, false, 2, true>;
    }
    int per_sm = 0;
    LZ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kThreads, 0));
    if (per_sm < 1) per_sm = 1;
    const int64_t gmax = std::min<int64_t>((int64_t)ctx->sms * per_sm, kMaxPartials);
    // choose the z-chunking: few chunks (small halo overhead 2/zc) but a CTA count that
    // fills the persistent grid evenly.
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
        // time ~ rounds * zc * (1 + halo), normalised by the ideal items*zc/gmax
        const double cost = (double)rounds * (zc + (has_z ? (alpha_only ? 1.0 : 2.0) : 0.0)) / ((double)st.nz * tiles / gmax);
        if (cost < best_cost - 1e-12) { best_cost = cost; best_chunks = chunks; }
        if (zc <= 8) break;
    }
    a.zc = (int)((st.nz + best_chunks - 1) / best_chunks);
    a.chunks_z = (int)((st.nz + a.zc - 1) / a.zc);
    a.nitems = tiles * a.chunks_z;
    if (a.alpha_partials) op->kb_zc = a.zc;          // the border kernel walks the same chunks
    const int grid = (int)std::min<int64_t>(a.nitems, gmax);
    void* args[] = {(void*)&a};
    LZ_CUDA(launch_fn(fn, dim3(grid), dim3(kThreads), 0, ctx->stream, args));
    if (nparts) *nparts = grid;
    return LZ_OK;
}

bool update_alpha_supported(const lz_op* op, const double* x, const double* out) {
    static const bool off = []() { const char* e = getenv("LZ_KB_ALPHA"); return e && e[0] == '0'; }();
    if (off || op->kind != LZ_OP_STENCIL || op->st.points != 7) return false;
    const lz_stencil& st = op->st;
    return st.offx != 0.0 && st.offy != 0.0 && st.offz != 0.0 && st.nx % 64 == 0 && st.ny % kWarps == 0 &&
           st.nx * st.ny * 8 < ((int64_t)1 << 32) &&
           ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(st.diag)) & 15) == 0;
}

int launch_alpha_border(lz_op* op, const double* v, double* partials, int* nparts, const FinTail* fin) {
    const lz_stencil& st = op->st;
    lz_ctx* ctx = op->ctx;
    LZ_REQUIRE(op->kb_zc > 0, "launch_alpha_border: no KB launch with alpha partials preceded");
    BorderArgs a;
    a.nx = (int)st.nx; a.ny = (int)st.ny; a.nz = (int)st.nz;
    a.periodic = (st.bc == LZ_BC_PERIODIC);
    a.tiles_x = (int)(st.nx / 64);
    a.tiles_y = (int)(st.ny / kWarps);
    a.zc = op->kb_zc;
    a.chunks_z = (int)((st.nz + a.zc - 1) / a.zc);
    a.plane = st.nx * st.ny;
    a.ox2 = 2.0 * st.offx; a.oy2 = 2.0 * st.offy; a.oz2 = 2.0 * st.offz;
    a.v = v;
    a.zhi = st.sharded ? st.ghost_hi : (a.periodic ? v : nullptr);
    a.partials = partials;
    if (fin) a.fin = *fin;
    const int64_t work = (int64_t)a.nz * a.tiles_y * (a.nx / 2) + (int64_t)a.chunks_z * (a.plane / 2) +
                         (int64_t)a.nz * a.ny * a.tiles_x;
    const int64_t want = (work + (int64_t)kThreads * 4 - 1) / ((int64_t)kThreads * 4);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, std::min<int64_t>((int64_t)ctx->sms * 8, kMaxPartials)));
    LZ_CUDA(launch_k(stencil_alpha_border_kernel, dim3(grid), dim3(kThreads), 0, ctx->stream, a));
    if (nparts) *nparts = grid;
    return LZ_OK;
}

}  // namespace lz
