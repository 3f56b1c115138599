// K4a / K4b: full or selective re-orthogonalisation against the Krylov basis held in HBM as
// a tall-skinny block GEMV pair (classical Gram-Schmidt), and K5: the Ritz-vector lift.
//
// Replaces Lanczos.reorthogonalize (Lanczos.py:233-251; IrrLanczos.py:448-466):
//     ip = sum(V[j]*V, axis=1)                     -> cgs_dots    (h = V_k^T v, one sweep)
//     V[j] = 2 V[j] - sum(ip[:,None]*V, axis=0)    -> cgs_update  (v = c v - V_k h, one sweep)
// without the two n x M temporaries of the reference, and only over the rows that exist.
// Both kernels are HBM-bound (0.25 flop/B): they run on CUDA cores with 128-bit streaming
// loads; tensor cores have nothing to offer here.
//
// Rows of the basis may be stored un-normalised (lazy 1/beta); the scale factors are
// folded into the coefficients by the scalar kernels in lanczos.cu, never into the data.
#include "internal.h"

namespace lz {

constexpr int kRowsPerCta = 8;     // register-blocked rows of the dots kernel

// part[r * ncg + g] = sum_{i in column group g} V[r, i] * target[i]
// grid = nrb * ncg CTAs, row-block index fastest so that the CTAs sharing a column group
// (and hence the same `target` strip) are co-resident and the strip is served from L2.
__global__ void __launch_bounds__(kThreads)
cgs_dots_kernel(const double* __restrict__ V, int64_t ldv, int nrows,
                const double* __restrict__ target, int64_t M, int64_t chunk, int nrb, int ncg,
                int vec_ok, double* __restrict__ part, const int* __restrict__ flag) {
    if (flag && *flag == 0) return;
    __shared__ double red[kWarps];
    const int rb = blockIdx.x % nrb;
    const int g = blockIdx.x / nrb;
    const int r0 = rb * kRowsPerCta;
    const int64_t c0 = (int64_t)g * chunk;
    const int64_t c1 = min(M, c0 + chunk);
    double acc[kRowsPerCta];
#pragma unroll
    for (int r = 0; r < kRowsPerCta; ++r) acc[r] = 0.0;
    const double* base = V + (int64_t)r0 * ldv;
    const int nr = min(kRowsPerCta, nrows - r0);

    if (vec_ok) {
        // c0 and chunk are even; handle an odd M tail below
        const int64_t e1 = c1 & ~(int64_t)1;
        if (nr == kRowsPerCta) {
            for (int64_t i = c0 + 2 * threadIdx.x; i < e1; i += 2 * kThreads) {
                const double2 t = ld_cached2(target + i);
                double2 v[kRowsPerCta];
#pragma unroll
                for (int r = 0; r < kRowsPerCta; ++r) v[r] = ld_stream2(base + (int64_t)r * ldv + i);
#pragma unroll
                for (int r = 0; r < kRowsPerCta; ++r) acc[r] = fma(t.y, v[r].y, fma(t.x, v[r].x, acc[r]));
            }
        } else {
            for (int64_t i = c0 + 2 * threadIdx.x; i < e1; i += 2 * kThreads) {
                const double2 t = ld_cached2(target + i);
#pragma unroll
                for (int r = 0; r < kRowsPerCta; ++r) {
                    if (r < nr) {
                        const double2 v = ld_stream2(base + (int64_t)r * ldv + i);
                        acc[r] = fma(t.y, v.y, fma(t.x, v.x, acc[r]));
                    }
                }
            }
        }
        if (e1 < c1 && threadIdx.x == 0) {   // odd last element of the vector
#pragma unroll
            for (int r = 0; r < kRowsPerCta; ++r)
                if (r < nr) acc[r] = fma(target[e1], base[(int64_t)r * ldv + e1], acc[r]);
        }
    } else {
        for (int64_t i = c0 + threadIdx.x; i < c1; i += kThreads) {
            const double t = __ldg(target + i);
#pragma unroll
            for (int r = 0; r < kRowsPerCta; ++r)
                if (r < nr) acc[r] = fma(t, ld_stream1(base + (int64_t)r * ldv + i), acc[r]);
        }
    }
#pragma unroll
    for (int r = 0; r < kRowsPerCta; ++r) {
        const double tot = block_sum(acc[r], red);
        if (threadIdx.x == 0 && r < nr) part[(int64_t)(r0 + r) * ncg + g] = tot;
    }
}

int launch_cgs_dots(lz_ctx* ctx, const double* V, int64_t ldv, int nrows, const double* target,
                    int64_t M, double* part, int* ncg_out, const int* flag_dev) {
    const int nrb = (nrows + kRowsPerCta - 1) / kRowsPerCta;
    const int64_t target_ctas = (int64_t)ctx->sms * 8;
    int64_t ncg = std::max<int64_t>(1, (target_ctas + nrb - 1) / nrb);
    const int64_t min_chunk = 2 * kThreads * 2;          // at least two iterations per CTA
    ncg = std::min<int64_t>(ncg, std::max<int64_t>(1, M / min_chunk));
    ncg = std::min<int64_t>(ncg, kMaxPartials);
    int64_t chunk = (M + ncg - 1) / ncg;
    chunk = (chunk + 2 * kThreads - 1) / (2 * kThreads) * (2 * kThreads);   // multiple of 512 columns
    ncg = (M + chunk - 1) / chunk;
    const int vec_ok = ((((uintptr_t)V | (uintptr_t)target) & 15) == 0) && ((ldv & 1) == 0);
    cgs_dots_kernel<<<(unsigned)(nrb * ncg), kThreads, 0, ctx->stream>>>(
        V, ldv, nrows, target, M, chunk, nrb, (int)ncg, vec_ok, part, flag_dev);
    LZ_CUDA(cudaGetLastError());
    if (ncg_out) *ncg_out = (int)ncg;
    return LZ_OK;
}

// out = cself * target - sum_{r < nrows} coef[r] * V[r, :]      (out may alias target)
constexpr int kCoefSmem = 2048;
__global__ void __launch_bounds__(kThreads)
cgs_update_kernel(const double* __restrict__ V, int64_t ldv, int nrows, const double* target,
                  const double* __restrict__ coef, const double* __restrict__ cself_p, double* out,
                  int64_t M, int vec_ok, const int* __restrict__ flag, const HaloPush halo) {
    if (flag && *flag == 0) return;
    __shared__ double sc[kCoefSmem];
    const int ns = min(nrows, kCoefSmem);
    for (int r = threadIdx.x; r < ns; r += kThreads) sc[r] = coef[r];
    __syncthreads();
    const double cself = __ldg(cself_p);
    const int64_t tid = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * kThreads;
    if (vec_ok) {
        const int64_t M2 = M >> 1;
        for (int64_t i = tid; i < M2; i += nthr) {
            const double2 t = ld_stream2_rw(target + 2 * i);
            double ax = cself * t.x, ay = cself * t.y;
            const double* p = V + 2 * i;
            int r = 0;
            for (; r + 8 <= ns; r += 8) {
                double2 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = ld_stream2(p + (int64_t)(r + u) * ldv);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    ax = fma(-sc[r + u], v[u].x, ax);
                    ay = fma(-sc[r + u], v[u].y, ay);
                }
            }
            for (; r < ns; ++r) {
                const double2 v = ld_stream2(p + (int64_t)r * ldv);
                ax = fma(-sc[r], v.x, ax);
                ay = fma(-sc[r], v.y, ay);
            }
            for (; r < nrows; ++r) {       // beyond the smem window (nrows > 4096)
                const double c = __ldg(coef + r);
                const double2 v = ld_stream2(p + (int64_t)r * ldv);
                ax = fma(-c, v.x, ax);
                ay = fma(-c, v.y, ay);
            }
            st_stream2(out + 2 * i, make_double2(ax, ay));
            if (halo.lo_dst && 2 * i < halo.plane) st_stream2(halo.lo_dst + 2 * i, make_double2(ax, ay));
            if (halo.hi_dst && 2 * i >= M - halo.plane)
                st_stream2(halo.hi_dst + (2 * i - (M - halo.plane)), make_double2(ax, ay));
        }
        if (tid == 0 && (M & 1)) {
            const int64_t i = M - 1;
            double a = cself * target[i];
            for (int r = 0; r < nrows; ++r) a = fma(-__ldg(coef + r), V[(int64_t)r * ldv + i], a);
            out[i] = a;
            if (halo.lo_dst && i < halo.plane) halo.lo_dst[i] = a;
            if (halo.hi_dst && i >= M - halo.plane) halo.hi_dst[i - (M - halo.plane)] = a;
        }
    } else {
        for (int64_t i = tid; i < M; i += nthr) {
            double a = cself * target[i];
            for (int r = 0; r < nrows; ++r) {
                const double c = (r < ns) ? sc[r] : __ldg(coef + r);
                a = fma(-c, ld_stream1(V + (int64_t)r * ldv + i), a);
            }
            out[i] = a;
            if (halo.lo_dst && i < halo.plane) halo.lo_dst[i] = a;
            if (halo.hi_dst && i >= M - halo.plane) halo.hi_dst[i - (M - halo.plane)] = a;
        }
    }
}

int launch_cgs_update(lz_ctx* ctx, const double* V, int64_t ldv, int nrows, const double* target,
                      const double* coef_dev, const double* cself_dev, double* out, int64_t M,
                      const int* flag_dev, const HaloPush* halo) {
    const int64_t want = (M / 2 + kThreads - 1) / kThreads;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)ctx->sms * 8));
    int vec_ok = ((((uintptr_t)V | (uintptr_t)target | (uintptr_t)out) & 15) == 0) && ((ldv & 1) == 0);
    HaloPush h{};
    if (halo && (halo->lo_dst || halo->hi_dst)) {
        h = *halo;
        vec_ok = vec_ok && (((uintptr_t)h.lo_dst | (uintptr_t)h.hi_dst) & 15) == 0 && ((h.plane & 1) == 0) &&
                 ((M & 1) == 0);
    }
    cgs_update_kernel<<<grid, kThreads, 0, ctx->stream>>>(V, ldv, nrows, target, coef_dev, cself_dev, out,
                                                          M, vec_ok, flag_dev, h);
    LZ_CUDA(cudaGetLastError());
    return LZ_OK;
}

// K5: Y[c, :] (+)= sum_{r < n} S[r + c*lds] * V[r, :]  for a block of kLiftCols columns per
// sweep; the basis is read once per block of columns.  n <= kLiftRowsSmem per launch.
constexpr int kLiftCols = 4;
constexpr int kLiftRowsSmem = 1024;
__global__ void __launch_bounds__(kThreads)
ritz_lift_kernel(const double* __restrict__ V, int64_t ldv, int n, int64_t M,
                 const double* __restrict__ S, int lds, int ncols, double* Y, int64_t ldy,
                 int accumulate, int vec_ok) {
    __shared__ double ss[kLiftRowsSmem * kLiftCols];
    for (int q = threadIdx.x; q < n * kLiftCols; q += kThreads) {
        const int r = q / kLiftCols, c = q % kLiftCols;
        ss[q] = (c < ncols) ? S[(int64_t)c * lds + r] : 0.0;
    }
    __syncthreads();
    const int64_t tid = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * kThreads;
    const int64_t cnt = vec_ok ? (M >> 1) : M;
    for (int64_t i = tid; i < cnt; i += nthr) {
        double ax[kLiftCols], ay[kLiftCols];
#pragma unroll
        for (int c = 0; c < kLiftCols; ++c) { ax[c] = 0.0; ay[c] = 0.0; }
        if (accumulate) {
#pragma unroll
            for (int c = 0; c < kLiftCols; ++c) {
                if (c < ncols) {
                    if (vec_ok) {
                        const double2 y = ld_stream2_rw(Y + (int64_t)c * ldy + 2 * i);
                        ax[c] = y.x; ay[c] = y.y;
                    } else {
                        ax[c] = Y[(int64_t)c * ldy + i];
                    }
                }
            }
        }
        for (int r = 0; r < n; ++r) {
            double vx, vy = 0.0;
            if (vec_ok) {
                const double2 v = ld_stream2(V + (int64_t)r * ldv + 2 * i);
                vx = v.x; vy = v.y;
            } else {
                vx = ld_stream1(V + (int64_t)r * ldv + i);
            }
#pragma unroll
            for (int c = 0; c < kLiftCols; ++c) {
                ax[c] = fma(ss[r * kLiftCols + c], vx, ax[c]);
                ay[c] = fma(ss[r * kLiftCols + c], vy, ay[c]);
            }
        }
#pragma unroll
        for (int c = 0; c < kLiftCols; ++c) {
            if (c < ncols) {
                if (vec_ok) st_stream2(Y + (int64_t)c * ldy + 2 * i, make_double2(ax[c], ay[c]));
                else Y[(int64_t)c * ldy + i] = ax[c];
            }
        }
    }
    if (vec_ok && (M & 1) && tid == 0) {
        const int64_t i = M - 1;
        for (int c = 0; c < ncols; ++c) {
            double a = accumulate ? Y[(int64_t)c * ldy + i] : 0.0;
            for (int r = 0; r < n; ++r) a = fma(ss[r * kLiftCols + c], V[(int64_t)r * ldv + i], a);
            Y[(int64_t)c * ldy + i] = a;
        }
    }
}

int launch_ritz_lift(lz_ctx* ctx, const double* V, int64_t ldv, int n, int64_t M,
                     const double* S_dev, int k, double* Y, int64_t ldy) {
    const int vec_ok = ((((uintptr_t)V | (uintptr_t)Y) & 15) == 0) && ((ldv & 1) == 0) && ((ldy & 1) == 0);
    const int64_t cnt = vec_ok ? (M >> 1) : M;
    const int64_t want = (cnt + kThreads - 1) / kThreads;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)ctx->sms * 8));
    for (int c0 = 0; c0 < k; c0 += kLiftCols) {
        const int nc = std::min(kLiftCols, k - c0);
        for (int w0 = 0; w0 < n; w0 += kLiftRowsSmem) {
            const int wn = std::min(kLiftRowsSmem, n - w0);
            ritz_lift_kernel<<<grid, kThreads, 0, ctx->stream>>>(
                V + (int64_t)w0 * ldv, ldv, wn, M, S_dev + (int64_t)c0 * n + w0, n, nc,
                Y + (int64_t)c0 * ldy, ldy, w0 > 0, vec_ok);
            LZ_CUDA(cudaGetLastError());
        }
    }
    return LZ_OK;
}

}  // namespace lz
