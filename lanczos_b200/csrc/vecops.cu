// K3 and friends: streaming fp64 vector kernels with fused, deterministic reductions.
//
//   update_norm:  r = w - (alpha*s_j) * a - (beta*s_jm1) * b ,  partial[cta] = sum r^2
//
// replaces `r = r - V[j]*alpha[j] - V[j-1]*beta[j-1]` and the following
// `np.linalg.norm(r)` (Lanczos.py:119,112).  `a`/`b` are basis rows stored un-normalised;
// s_j, s_jm1 fold the lazy 1/beta factors in.  4 doubles per thread per iteration as two
// independent 128-bit loads per operand; grid = resident CTAs, grid-stride.
#include "internal.h"

namespace lz {

static int stream_grid(lz_ctx* ctx, int64_t M, int per_thread) {
    const int64_t want = (M + (int64_t)kThreads * per_thread - 1) / ((int64_t)kThreads * per_thread);
    const int64_t cap = std::min<int64_t>((int64_t)ctx->sms * 8, kMaxPartials);
    return (int)std::max<int64_t>(1, std::min<int64_t>(want, cap));
}

__global__ void __launch_bounds__(kThreads)
dot_kernel(const double* __restrict__ x, const double* __restrict__ y, int64_t M, int vec_ok,
           double* __restrict__ partials) {
    __shared__ double red[kWarps];
    double acc0 = 0.0, acc1 = 0.0;
    const int64_t tid = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * kThreads;
    if (vec_ok) {
        const int64_t M2 = M >> 1;
        for (int64_t i = tid; i < M2; i += nthr) {
            const double2 a = ld_stream2(x + 2 * i), b = ld_stream2(y + 2 * i);
            acc0 = fma(a.x, b.x, acc0);
            acc1 = fma(a.y, b.y, acc1);
        }
        if (tid == 0 && (M & 1)) acc0 = fma(x[M - 1], y[M - 1], acc0);
    } else {
        for (int64_t i = tid; i < M; i += nthr) acc0 = fma(ld_stream1(x + i), ld_stream1(y + i), acc0);
    }
    const double tot = block_sum(acc0 + acc1, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = tot;
}

int launch_dot(lz_ctx* ctx, const double* x, const double* y, int64_t M, double* partials, int* nparts) {
    const int grid = stream_grid(ctx, M, 4);
    const int vec_ok = (((uintptr_t)x | (uintptr_t)y) & 15) == 0;
    dot_kernel<<<grid, kThreads, 0, ctx->stream>>>(x, y, M, vec_ok, partials);
    LZ_CUDA(cudaGetLastError());
    if (nparts) *nparts = grid;
    return LZ_OK;
}

// out = w - ca*sa*a - cb*sb*b ;  b nullable.  ca/sa/cb/sb are device scalars (nullable => 1).
template <bool HAS_B>
__global__ void __launch_bounds__(kThreads)
update_norm_kernel(const double* w, const double* __restrict__ a, const double* __restrict__ b,
                   const double* __restrict__ ca, const double* __restrict__ sa,
                   const double* __restrict__ cb, const double* __restrict__ sb,
                   double* out, int64_t M, int vec_ok, double* __restrict__ partials) {
    __shared__ double red[kWarps];
    const double fa = (ca ? __ldg(ca) : 1.0) * (sa ? __ldg(sa) : 1.0);
    const double fb = HAS_B ? (cb ? __ldg(cb) : 1.0) * (sb ? __ldg(sb) : 1.0) : 0.0;
    double acc0 = 0.0, acc1 = 0.0;
    const int64_t tid = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * kThreads;
    if (vec_ok) {
        const int64_t M2 = M >> 1;
        int64_t i = tid;
        // two independent 128-bit transactions per operand per iteration
        for (; i + nthr < M2; i += 2 * nthr) {
            const int64_t i1 = i + nthr;
            const double2 w0 = ld_stream2_rw(w + 2 * i), w1 = ld_stream2_rw(w + 2 * i1);
            const double2 a0 = ld_stream2(a + 2 * i), a1 = ld_stream2(a + 2 * i1);
            double2 r0, r1;
            r0.x = fma(-fa, a0.x, w0.x); r0.y = fma(-fa, a0.y, w0.y);
            r1.x = fma(-fa, a1.x, w1.x); r1.y = fma(-fa, a1.y, w1.y);
            if (HAS_B) {
                const double2 b0 = ld_stream2(b + 2 * i), b1 = ld_stream2(b + 2 * i1);
                r0.x = fma(-fb, b0.x, r0.x); r0.y = fma(-fb, b0.y, r0.y);
                r1.x = fma(-fb, b1.x, r1.x); r1.y = fma(-fb, b1.y, r1.y);
            }
            st_stream2(out + 2 * i, r0);
            st_stream2(out + 2 * i1, r1);
            acc0 = fma(r0.x, r0.x, acc0); acc1 = fma(r0.y, r0.y, acc1);
            acc0 = fma(r1.x, r1.x, acc0); acc1 = fma(r1.y, r1.y, acc1);
        }
        for (; i < M2; i += nthr) {
            const double2 w0 = ld_stream2_rw(w + 2 * i);
            const double2 a0 = ld_stream2(a + 2 * i);
            double2 r0;
            r0.x = fma(-fa, a0.x, w0.x); r0.y = fma(-fa, a0.y, w0.y);
            if (HAS_B) {
                const double2 b0 = ld_stream2(b + 2 * i);
                r0.x = fma(-fb, b0.x, r0.x); r0.y = fma(-fb, b0.y, r0.y);
            }
            st_stream2(out + 2 * i, r0);
            acc0 = fma(r0.x, r0.x, acc0); acc1 = fma(r0.y, r0.y, acc1);
        }
        if (tid == 0 && (M & 1)) {
            double r = fma(-fa, a[M - 1], w[M - 1]);
            if (HAS_B) r = fma(-fb, b[M - 1], r);
            out[M - 1] = r;
            acc0 = fma(r, r, acc0);
        }
    } else {
        for (int64_t i = tid; i < M; i += nthr) {
            double r = fma(-fa, ld_stream1(a + i), w[i]);
            if (HAS_B) r = fma(-fb, ld_stream1(b + i), r);
            out[i] = r;
            acc0 = fma(r, r, acc0);
        }
    }
    const double tot = block_sum(acc0 + acc1, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = tot;
}

int launch_update_norm(lz_ctx* ctx, const double* w, const double* a, const double* b,
                       const double* ca_dev, const double* sa_dev, const double* cb_dev,
                       const double* sb_dev, double* out, int64_t M, double* partials, int* nparts) {
    const int grid = stream_grid(ctx, M, 4);
    const int vec_ok = (((uintptr_t)w | (uintptr_t)a | (uintptr_t)b | (uintptr_t)out) & 15) == 0;
    if (b)
        update_norm_kernel<true><<<grid, kThreads, 0, ctx->stream>>>(w, a, b, ca_dev, sa_dev, cb_dev, sb_dev,
                                                                    out, M, vec_ok, partials);
    else
        update_norm_kernel<false><<<grid, kThreads, 0, ctx->stream>>>(w, a, nullptr, ca_dev, sa_dev, nullptr,
                                                                     nullptr, out, M, vec_ok, partials);
    LZ_CUDA(cudaGetLastError());
    if (nparts) *nparts = grid;
    return LZ_OK;
}

__global__ void __launch_bounds__(kThreads) scale_kernel(double* x, int64_t M, double s) {
    const int64_t tid = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * kThreads;
    for (int64_t i = tid; i < M; i += nthr) x[i] *= s;
}

int launch_scale(lz_ctx* ctx, double* x, int64_t M, double s) {
    const int grid = stream_grid(ctx, M, 1);
    scale_kernel<<<grid, kThreads, 0, ctx->stream>>>(x, M, s);
    LZ_CUDA(cudaGetLastError());
    return LZ_OK;
}

}  // namespace lz
