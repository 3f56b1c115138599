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
           double* __restrict__ partials, const FinTail fin) {
    pdl_prologue();
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
    fin_tail(fin, partials, red);
}

int launch_dot(lz_ctx* ctx, const double* x, const double* y, int64_t M, double* partials, int* nparts,
               const FinTail* fin) {
    const int grid = stream_grid(ctx, M, 4);
    const int vec_ok = (((uintptr_t)x | (uintptr_t)y) & 15) == 0;
    LZ_CUDA(launch_k(dot_kernel, dim3(grid), dim3(kThreads), 0, ctx->stream, x, y, M, vec_ok, partials,
                     fin ? *fin : FinTail{}));
    if (nparts) *nparts = grid;
    return LZ_OK;
}

// Sharded runs: the first / last z-plane of the vector being produced is also stored straight
// into the neighbouring GPUs' ghost buffers (NVLink peer stores), so that the halo exchange
// costs no extra pass and no extra launch.
template <bool HALO>
__device__ __forceinline__ void halo_store2(const HaloPush& h, int64_t e, int64_t M, double2 v) {
    if (HALO) {
        if (h.lo_dst && e < h.plane) st_stream2(h.lo_dst + e, v);
        if (h.hi_dst && e >= M - h.plane) st_stream2(h.hi_dst + (e - (M - h.plane)), v);
    }
}
template <bool HALO>
__device__ __forceinline__ void halo_store1(const HaloPush& h, int64_t e, int64_t M, double v) {
    if (HALO) {
        if (h.lo_dst && e < h.plane) h.lo_dst[e] = v;
        if (h.hi_dst && e >= M - h.plane) h.hi_dst[e - (M - h.plane)] = v;
    }
}

// out = w - ca*sa*a - cb*sb*b ;  b nullable.  ca/sa/cb/sb are device scalars (nullable => 1).
template <bool HAS_B, bool HALO>
__global__ void __launch_bounds__(kThreads)
update_norm_kernel(const double* w, const double* __restrict__ a, const double* __restrict__ b,
                   const double* __restrict__ ca, const double* __restrict__ sa,
                   const double* __restrict__ cb, const double* __restrict__ sb,
                   double* out, int64_t M, int vec_ok, double* __restrict__ partials,
                   const HaloPush halo, const FinTail fin, const double* __restrict__ sw_p) {
    pdl_prologue();
    __shared__ double red[kWarps];
    // w may arrive un-normalised (H applied to the un-normalised row before beta was known): r = sw*w - ...
    const double sw = sw_p ? __ldg(sw_p) : 1.0;
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
            r0.x = fma(-fa, a0.x, sw * w0.x); r0.y = fma(-fa, a0.y, sw * w0.y);
            r1.x = fma(-fa, a1.x, sw * w1.x); r1.y = fma(-fa, a1.y, sw * w1.y);
            if (HAS_B) {
                const double2 b0 = ld_stream2(b + 2 * i), b1 = ld_stream2(b + 2 * i1);
                r0.x = fma(-fb, b0.x, r0.x); r0.y = fma(-fb, b0.y, r0.y);
                r1.x = fma(-fb, b1.x, r1.x); r1.y = fma(-fb, b1.y, r1.y);
            }
            st_stream2(out + 2 * i, r0);
            st_stream2(out + 2 * i1, r1);
            halo_store2<HALO>(halo, 2 * i, M, r0);
            halo_store2<HALO>(halo, 2 * i1, M, r1);
            acc0 = fma(r0.x, r0.x, acc0); acc1 = fma(r0.y, r0.y, acc1);
            acc0 = fma(r1.x, r1.x, acc0); acc1 = fma(r1.y, r1.y, acc1);
        }
        for (; i < M2; i += nthr) {
            const double2 w0 = ld_stream2_rw(w + 2 * i);
            const double2 a0 = ld_stream2(a + 2 * i);
            double2 r0;
            r0.x = fma(-fa, a0.x, sw * w0.x); r0.y = fma(-fa, a0.y, sw * w0.y);
            if (HAS_B) {
                const double2 b0 = ld_stream2(b + 2 * i);
                r0.x = fma(-fb, b0.x, r0.x); r0.y = fma(-fb, b0.y, r0.y);
            }
            st_stream2(out + 2 * i, r0);
            halo_store2<HALO>(halo, 2 * i, M, r0);
            acc0 = fma(r0.x, r0.x, acc0); acc1 = fma(r0.y, r0.y, acc1);
        }
        if (tid == 0 && (M & 1)) {
            double r = fma(-fa, a[M - 1], sw * w[M - 1]);
            if (HAS_B) r = fma(-fb, b[M - 1], r);
            out[M - 1] = r;
            halo_store1<HALO>(halo, M - 1, M, r);
            acc0 = fma(r, r, acc0);
        }
    } else {
        for (int64_t i = tid; i < M; i += nthr) {
            double r = fma(-fa, ld_stream1(a + i), sw * w[i]);
            if (HAS_B) r = fma(-fb, ld_stream1(b + i), r);
            out[i] = r;
            halo_store1<HALO>(halo, i, M, r);
            acc0 = fma(r, r, acc0);
        }
    }
    const double tot = block_sum(acc0 + acc1, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = tot;
    fin_tail(fin, partials, red);
}

int launch_update_norm(lz_ctx* ctx, const double* w, const double* a, const double* b,
                       const double* ca_dev, const double* sa_dev, const double* cb_dev,
                       const double* sb_dev, double* out, int64_t M, double* partials, int* nparts,
                       const HaloPush* halo, const FinTail* fin, const double* sw_dev) {
    const int grid = stream_grid(ctx, M, 4);
    const FinTail ft = fin ? *fin : FinTail{};
    HaloPush h{};
    const bool push = halo && (halo->lo_dst || halo->hi_dst);
    if (push) h = *halo;
    int vec_ok = (((uintptr_t)w | (uintptr_t)a | (uintptr_t)b | (uintptr_t)out) & 15) == 0;
    if (push) vec_ok = vec_ok && (((uintptr_t)h.lo_dst | (uintptr_t)h.hi_dst) & 15) == 0 &&
                       ((h.plane & 1) == 0) && ((M & 1) == 0);
#define LZ_UPD(HB, HL, bb, cbb, sbb)                                                               \
    LZ_CUDA(launch_k(update_norm_kernel<HB, HL>, dim3(grid), dim3(kThreads), 0, ctx->stream, w, a,       \
                     (const double*)bb, ca_dev, sa_dev, (const double*)cbb, (const double*)sbb, out, M,  \
                     vec_ok, partials, h, ft, sw_dev))
    if (b) { if (push) LZ_UPD(true, true, b, cb_dev, sb_dev); else LZ_UPD(true, false, b, cb_dev, sb_dev); }
    else { if (push) LZ_UPD(false, true, nullptr, nullptr, nullptr); else LZ_UPD(false, false, nullptr, nullptr, nullptr); }
#undef LZ_UPD
    LZ_CUDA(cudaGetLastError());
    if (nparts) *nparts = grid;
    return LZ_OK;
}

// plain halo publication of an existing vector (start vector / first row)
__global__ void __launch_bounds__(kThreads)
halo_push_kernel(const double* __restrict__ x, int64_t M, const HaloPush h) {
    pdl_prologue();
    const int64_t tid = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * kThreads;
    for (int64_t e = tid; e < h.plane; e += nthr) {
        if (h.lo_dst) h.lo_dst[e] = x[e];
        if (h.hi_dst) h.hi_dst[e] = x[M - h.plane + e];
    }
}

int launch_halo_push(lz_ctx* ctx, const double* x, int64_t M, const HaloPush* halo) {
    if (!halo || !(halo->lo_dst || halo->hi_dst)) return LZ_OK;
    const int grid = stream_grid(ctx, halo->plane, 1);
    LZ_CUDA(launch_k(halo_push_kernel, dim3(grid), dim3(kThreads), 0, ctx->stream, x, M, *halo));
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
