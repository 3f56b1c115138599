// K4a / K4b: full or selective re-orthogonalisation against the Krylov basis held in HBM as
// a tall-skinny block GEMV pair (classical Gram-Schmidt), K4c: the fused middle of CGS2 (tile of the
// basis staged in shared memory by a TMA tensor copy), and K5: the Ritz-vector lift.
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
#include <stdlib.h>
#include <stdlib.h>
#include "internal.h"
#include "tma.cuh"

namespace lz {

constexpr int kRowsPerCta = 8;     // register-blocked rows of the dots kernel

// part[r * ncg + g] = sum_{i in column group g} V[r, i] * target[i]
// grid = nrb * ncg CTAs, row-block index fastest so that the CTAs sharing a column group
// (and hence the same `target` strip) are co-resident and the strip is served from L2.
// (3 CTAs/SM: with the tail in the kernel the compiler's own choice was 64 registers, which serialises part of the
// 8-row load batch - 5.17 instead of 4.84 ms at 512^3, 60 rows)
__global__ void __launch_bounds__(kThreads, 3)
cgs_dots_kernel(const double* __restrict__ V, int64_t ldv, int nrows,
                const double* __restrict__ target, int64_t M, int64_t chunk, int nrb, int ncg,
                int vec_ok, double* __restrict__ part, const int* __restrict__ flag, const IpTail tail) {
    pdl_prologue();
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
    ip_tail(tail, part, ncg);
}

int launch_cgs_dots(lz_ctx* ctx, const double* V, int64_t ldv, int nrows, const double* target,
                    int64_t M, double* part, int* ncg_out, const int* flag_dev, const IpTail* tail) {
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
    LZ_CUDA(launch_k(cgs_dots_kernel, dim3((unsigned)(nrb * ncg)), dim3(kThreads), 0, ctx->stream,
                     V, ldv, nrows, target, M, chunk, nrb, (int)ncg, vec_ok, part, flag_dev,
                     tail ? *tail : IpTail{}));
    if (ncg_out) *ncg_out = (int)ncg;
    return LZ_OK;
}

// out = cself * target - sum_{r < nrows} coef[r] * V[r, :]      (out may alias target)
constexpr int kCoefSmem = 2048;
__global__ void __launch_bounds__(kThreads)
cgs_update_kernel(const double* __restrict__ V, int64_t ldv, int nrows, const double* target,
                  const double* __restrict__ coef, const double* __restrict__ cself_p, double* out,
                  int64_t M, int vec_ok, const int* __restrict__ flag, const HaloPush halo) {
    pdl_prologue();
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
    LZ_CUDA(launch_k(cgs_update_kernel, dim3(grid), dim3(kThreads), 0, ctx->stream, V, ldv, nrows, target,
                     coef_dev, cself_dev, out, M, vec_ok, flag_dev, h));
    return LZ_OK;
}

// K4c: the middle of CGS2 from ONE read of the basis.  Classical Gram-Schmidt applied twice is
//     h1 = V^T v;  v' = c v - V h1;  h2 = V^T v';  v'' = v' - V h2
// i.e. four sweeps over the k basis rows, (4k + 6) * 8 * M bytes.  The update of the first pass and the
// dots of the second touch the same rows: here a CTA stages a tile of TC columns of all k rows (and of
// v) in shared memory with 16-byte cp.async copies, forms v' for the tile (thread <-> column, rows in
// order: the same bits as K4b), stores it, and reduces V_tile . v'_tile for every row (warp <-> rows,
// lanes <-> columns) into per-lane accumulators that live across the CTA's tiles.  CGS2 then moves
// (3k + 5) * 8 * M bytes.  Two or three CTAs per SM overlap one CTA's copy with another's arithmetic;
// the bytes in flight are the staged tiles, not registers.
template <int TC, int RMAX>
__global__ void __launch_bounds__(TC)
cgs_update_dots_kernel(const double* __restrict__ V, int64_t ldv, int k, double* target,
                       const double* __restrict__ coef, const double* __restrict__ cself_p, int64_t M,
                       int64_t ntiles, double* __restrict__ part, const int* __restrict__ flag, const IpTail tail) {
    pdl_prologue();
    if (flag && *flag == 0) return;
    extern __shared__ __align__(128) double sm[];
    double* S = sm;                          // [k + 1][TC]: rows 0..k-1 of the basis, row k = v
    double* vp = sm + (size_t)(k + 1) * TC;  // [TC]: v' of the tile
    double* sh = vp + TC;                    // [k]: h1
    constexpr int NW = TC / 32;
    constexpr int CPL = TC / 32;             // columns per lane in the reduction
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int r = t; r < k; r += TC) sh[r] = coef[r];
    const double cself = __ldg(cself_p);
    double acc[RMAX];
#pragma unroll
    for (int q = 0; q < RMAX; ++q) acc[q] = 0.0;
    constexpr int UPR = TC / 2;              // 16-byte units per row of the tile
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t c0 = tile * TC;
        __syncthreads();                     // the previous tile's reduction has finished reading S / vp
        for (int u = t; u < (k + 1) * UPR; u += TC) {
            const int r = u / UPR, cu = u - r * UPR;
            const int64_t col = c0 + 2 * cu;
            const double* src = (r < k ? V + (int64_t)r * ldv : target) + col;
            const int valid = col + 1 < M ? 16 : (col < M ? 8 : 0);      // zero-fill beyond M
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(S + (size_t)r * TC + 2 * cu);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(valid ? src : V), "r"(valid) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        {   // v' for column t of the tile (rows in order, like cgs_update_kernel)
            double a = cself * S[(size_t)k * TC + t];
            for (int r = 0; r < k; ++r) a = fma(-sh[r], S[(size_t)r * TC + t], a);
            vp[t] = a;
            if (c0 + t < M) st_stream1(target + c0 + t, a);
        }
        __syncthreads();
        double vl[CPL];
#pragma unroll
        for (int c = 0; c < CPL; ++c) vl[c] = vp[lane + 32 * c];
#pragma unroll
        for (int q = 0; q < RMAX; ++q) {
            const int r = warp + q * NW;
            if (r < k) {
                const double* row = S + (size_t)r * TC + lane;
                double a = acc[q];
#pragma unroll
                for (int c = 0; c < CPL; ++c) a = fma(row[32 * c], vl[c], a);
                acc[q] = a;
            }
        }
    }
#pragma unroll
    for (int q = 0; q < RMAX; ++q) {
        const int r = warp + q * NW;
        const double tot = warp_sum(acc[q]);
        if (lane == 0 && r < k) part[(int64_t)r * gridDim.x + blockIdx.x] = tot;
    }
    ip_tail(tail, part, (int)gridDim.x);
}

// ---- K4c with TMA: the same computation, the tile fetched by ONE tensor copy ----------------------
// The k basis rows and the target (row k of the same array) are a 2-D tensor (M columns x (k + 1) rows,
// row stride ldv); a tile is the box TC x (k + 1).  One elected thread arms an mbarrier with the box's
// byte count and issues a single cp.async.bulk.tensor.2d; the TMA unit walks the rows, zero-fills the
// columns past M and lands the box densely in shared memory.  Two stages: the box of the CTA's next
// tile is in flight while the current one is reduced, so one CTA keeps 60-120 KB of HBM reads in flight
// with no registers or issue slots spent on them.
template <int TC, int RMAX>
__global__ void __launch_bounds__(TC)
cgs_update_dots_tma_kernel(const __grid_constant__ CUtensorMap tmap, int k, double* target,
                           const double* __restrict__ coef, const double* __restrict__ cself_p, int64_t M,
                           int64_t ntiles, double* __restrict__ part, const int* __restrict__ flag, const IpTail tail) {
    pdl_prologue();
    if (flag && *flag == 0) return;
    extern __shared__ __align__(128) double sm[];
    const size_t stage_d = (size_t)(k + 1) * TC;     // TC * 8 is a multiple of 128 B: every stage is 128-byte aligned
    double* vp = sm + 2 * stage_d;                   // [TC]
    double* sh = vp + TC;                            // [k]
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(sh + ((k + 1) & ~1) + 2);
    constexpr int NW = TC / 32;
    constexpr int CPL = TC / 32;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const uint32_t bar0 = smem_addr(bars), bar1 = smem_addr(bars + 1);
    const uint32_t box_bytes = (uint32_t)stage_d * 8u;
    if (t == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int r = t; r < k; r += TC) sh[r] = coef[r];
    const double cself = __ldg(cself_p);
    double acc[RMAX];
#pragma unroll
    for (int q = 0; q < RMAX; ++q) acc[q] = 0.0;
    __syncthreads();

    int64_t tile = blockIdx.x;
    if (t == 0 && tile < ntiles) {
        mbar_arrive_expect_tx(bar0, box_bytes);
        tma_load_2d(smem_addr(sm), &tmap, (int)(tile * TC), 0, bar0);
    }
    for (int it = 0; tile < ntiles; ++it, tile += gridDim.x) {
        const int s = it & 1;
        const int64_t c0 = tile * TC;
        const int64_t next = tile + gridDim.x;
        if (t == 0 && next < ntiles) {
            // stage s^1 was last read (generic proxy) before the barrier that ended the previous iteration
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            const uint32_t bn = s ? bar0 : bar1;
            mbar_arrive_expect_tx(bn, box_bytes);
            tma_load_2d(smem_addr(sm + (size_t)(s ^ 1) * stage_d), &tmap, (int)(next * TC), 0, bn);
        }
        mbar_wait(s ? bar1 : bar0, (uint32_t)(it >> 1) & 1u);
        const double* S = sm + (size_t)s * stage_d;
        {   // columns past M arrive as zeros: v' = 0 there and nothing is stored
            double a = cself * S[(size_t)k * TC + t];
            for (int r = 0; r < k; ++r) a = fma(-sh[r], S[(size_t)r * TC + t], a);
            vp[t] = a;
            if (c0 + t < M) st_stream1(target + c0 + t, a);
        }
        __syncthreads();
        double vl[CPL];
#pragma unroll
        for (int c = 0; c < CPL; ++c) vl[c] = vp[lane + 32 * c];
#pragma unroll
        for (int q = 0; q < RMAX; ++q) {
            const int r = warp + q * NW;
            if (r < k) {
                const double* row = S + (size_t)r * TC + lane;
                double a = acc[q];
#pragma unroll
                for (int c = 0; c < CPL; ++c) a = fma(row[32 * c], vl[c], a);
                acc[q] = a;
            }
        }
        __syncthreads();                     // stage s and vp are free again
    }
#pragma unroll
    for (int q = 0; q < RMAX; ++q) {
        const int r = warp + q * NW;
        const double tot = warp_sum(acc[q]);
        if (lane == 0 && r < k) part[(int64_t)r * gridDim.x + blockIdx.x] = tot;
    }
    ip_tail(tail, part, (int)gridDim.x);
}

struct UpdDotsCfg { int tc, rmax; size_t smem; };
// TMA form: two stages of (k + 1) * TC doubles in one CTA; TC = 256 up to 48 rows, else 128
static bool update_dots_tma_config(int k, UpdDotsCfg* cfg) {
    const size_t cap = 200 * 1024;
    for (int tc : {256, 128}) {
        if (k + 1 > 256) continue;                                 // box rows
        if (tc == 256 && k > 48) continue;
        const size_t bytes = (2 * (size_t)(k + 1) * tc + tc + k + 8) * 8;
        const int nw = tc / 32;
        const int need = (k + nw - 1) / nw;
        if (bytes <= cap && need <= 26) {
            cfg->tc = tc;
            cfg->rmax = need <= 8 ? 8 : (need <= 16 ? 16 : 26);
            cfg->smem = bytes;
            return true;
        }
    }
    return false;
}
static int update_dots_use_tma() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("LZ_K4C_TMA");      // 0: force the cp.async form (comparison runs)
        v = e ? atoi(e) : 1;
    }
    return v;
}
static bool update_dots_config(int k, UpdDotsCfg* cfg) {
    // TC = 256 while two CTAs of (k + 2) * 2 KB fit an SM, else TC = 128; rows per warp <= RMAX
    const size_t cap = 100 * 1024;
    for (int tc : {256, 128}) {
        const size_t bytes = ((size_t)(k + 1) * tc + tc + k + 2) * 8;
        const int nw = tc / 32;
        const int need = (k + nw - 1) / nw;
        if (bytes <= cap && need <= 26) {
            cfg->tc = tc;
            cfg->rmax = need <= 8 ? 8 : (need <= 16 ? 16 : 26);
            cfg->smem = bytes;
            return true;
        }
    }
    return false;
}

bool cgs_update_dots_supported(const double* V, int64_t ldv, int k, const double* target) {
    UpdDotsCfg c;
    return k >= 1 && ((((uintptr_t)V | (uintptr_t)target) & 15) == 0) && ((ldv & 1) == 0) && update_dots_config(k, &c);
}

// target <- cself * target - V_k coef (in place), part[r * ncg + g] = partial of V[r, :] . target_new
int launch_cgs_update_dots(lz_ctx* ctx, const double* V, int64_t ldv, int k, double* target,
                           const double* coef_dev, const double* cself_dev, int64_t M, double* part,
                           int* ncg_out, const int* flag_dev, const IpTail* tail) {
    IpTail tl = tail ? *tail : IpTail{};
    UpdDotsCfg c;
    LZ_REQUIRE(cgs_update_dots_supported(V, ldv, k, target), "fused Gram-Schmidt update+dots: unsupported shape (k = %d)", k);
    update_dots_config(k, &c);
    const void* fn = nullptr;
    UpdDotsCfg ct;
    // TMA form: the target must be row k of the same array (it is, in the Lanczos loop)
    bool tma = update_dots_use_tma() && target == V + (int64_t)k * ldv && M <= 0x7fffffff && encode_tiled_fn() &&
               update_dots_tma_config(k, &ct);
    CUtensorMap tmap;
    if (tma) {
        const cuuint64_t gdim[2] = {(cuuint64_t)M, (cuuint64_t)(k + 1)};
        const cuuint64_t gstride[1] = {(cuuint64_t)ldv * 8};
        const cuuint32_t box[2] = {(cuuint32_t)ct.tc, (cuuint32_t)(k + 1)};
        const cuuint32_t estr[2] = {1, 1};
        tma = encode_tiled_fn()(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(V), gdim, gstride, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    }
    if (tma) {
        c = ct;
#define LZ_UDT(TCV, RM) if (c.tc == TCV && c.rmax == RM) fn = (const void*)cgs_update_dots_tma_kernel<TCV, RM>
        LZ_UDT(256, 8); LZ_UDT(256, 16); LZ_UDT(256, 26); LZ_UDT(128, 8); LZ_UDT(128, 16); LZ_UDT(128, 26);
#undef LZ_UDT
    } else {
#define LZ_UD(TCV, RM) if (c.tc == TCV && c.rmax == RM) fn = (const void*)cgs_update_dots_kernel<TCV, RM>
        LZ_UD(256, 8); LZ_UD(256, 16); LZ_UD(256, 26); LZ_UD(128, 8); LZ_UD(128, 16); LZ_UD(128, 26);
#undef LZ_UD
    }
    LZ_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem));
    int per_sm = 0;
    LZ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, c.tc, c.smem));
    if (per_sm < 1) per_sm = 1;
    const int64_t ntiles = (M + c.tc - 1) / c.tc;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ntiles, std::min<int64_t>((int64_t)ctx->sms * per_sm, kMaxPartials)));
    int kk = k;
    if (tma) {
        void* targs[] = {(void*)&tmap, (void*)&kk, (void*)&target, (void*)&coef_dev, (void*)&cself_dev,
                         (void*)&M, (void*)&ntiles, (void*)&part, (void*)&flag_dev, (void*)&tl};
        LZ_CUDA(launch_fn(fn, dim3(grid), dim3(c.tc), c.smem, ctx->stream, targs));
    } else {
        void* args[] = {(void*)&V, (void*)&ldv, (void*)&kk, (void*)&target, (void*)&coef_dev, (void*)&cself_dev,
                        (void*)&M, (void*)&ntiles, (void*)&part, (void*)&flag_dev, (void*)&tl};
        LZ_CUDA(launch_fn(fn, dim3(grid), dim3(c.tc), c.smem, ctx->stream, args));
    }
    if (ncg_out) *ncg_out = grid;
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

// K5 as a tall-skinny GEMM: Y (k x M) = S^T (k x n) . V (n x M), the whole lift of get_H_eigs
// (Lanczos.py:154-156) from ONE read of the basis per 64 Ritz vectors instead of one per 4.
//
// A CTA of 16 warps owns a tile of 256 columns of M and 64 output rows (Ritz vectors): warp w works on
// the 128 columns of half (w & 1) and the 8 Ritz vectors of group (w >> 1); a thread holds an 8 x 4 block
// of accumulators (columns 2*lane, 2*lane+1 of each 64-column quarter).  The rows of the basis tile stream
// through a shared-memory ring filled by 16-byte cp.async copies, four rows per step (one copy per thread)
// and four steps ahead (32 KB in flight per SM); the ring runs on across tiles, so the pipeline never
// drains.  A tile's rows are padded to a multiple of four with zero-filled copies, so a step never straddles
// two tiles and its four rows are straight-line code: per basis row a thread issues 2 + 4 shared-memory
// loads (its 4 values, 8 broadcast coefficients) for 32 fp64 FMAs.  S sits in shared memory (<= 128 rows per
// launch, more rows accumulate into Y in further launches).  At n = k = 60, 512^3: 2 n 8 M = 129 GB of
// compulsory traffic and 0.97 TFLOP - the two roofs (6.5 TB/s, 37 TFLOP/s) are 20 ms and 26 ms.
constexpr int kGemmThreads = 512;
constexpr int kGemmTile = 256;          // columns of M per CTA tile
constexpr int kGemmOut = 64;            // Ritz vectors per pass
constexpr int kGemmRowsPerStep = 4;     // basis rows per pipeline step
constexpr int kGemmDepth = 4;           // steps in flight
constexpr int kGemmRing = kGemmRowsPerStep * kGemmDepth;
constexpr int kGemmRowsSmem = 128;      // rows of S per launch
constexpr int kGemmAcc = 8;             // Ritz vectors per thread

__device__ __forceinline__ void cp_async16_zfill(void* smem, const void* gmem, int src_bytes) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}

__global__ void __launch_bounds__(kGemmThreads, 1)
ritz_lift_gemm_kernel(const double* __restrict__ V, int64_t ldv, int n, int64_t M, const double* __restrict__ S,
                      int lds, int ncols, double* Y, int64_t ldy, int accumulate, int64_t ntiles) {
    extern __shared__ __align__(128) double gsm[];
    double* ss = gsm;                                        // [nr4][kGemmOut], zero beyond n rows / ncols columns
    double* ring = gsm + (size_t)kGemmRowsSmem * kGemmOut;   // [kGemmRing][kGemmTile]
    const int nr4 = (n + kGemmRowsPerStep - 1) / kGemmRowsPerStep * kGemmRowsPerStep;
    for (int q = threadIdx.x; q < nr4 * kGemmOut; q += kGemmThreads) {
        const int r = q / kGemmOut, c = q % kGemmOut;
        ss[q] = (r < n && c < ncols) ? S[(int64_t)c * lds + r] : 0.0;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int half = warp & 1, grp = warp >> 1;
    const int steps_per_tile = nr4 / kGemmRowsPerStep;
    const int my_tiles = (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
    const int64_t nsteps = (int64_t)my_tiles * steps_per_tile;
    // producer: thread t copies 16-byte chunk (t & 127) of row (t >> 7) of every step (own (tile, step) cursor)
    const int chunk = threadIdx.x & 127, rsub = threadIdx.x >> 7;
    int64_t pstep = 0;
    int pst = 0;                                             // step inside the tile
    int64_t ptile = blockIdx.x;
    auto issue_step = [&]() {
        if (pstep < nsteps) {
            const int r = pst * kGemmRowsPerStep + rsub;
            const int64_t col = ptile * kGemmTile + 2 * chunk;
            const int64_t left = M - col;
            const int bytes = (r < n) ? (left >= 2 ? 16 : (left == 1 ? 8 : 0)) : 0;
            const double* src = bytes ? V + (int64_t)r * ldv + col : V;
            cp_async16_zfill(ring + (size_t)((pstep & (kGemmDepth - 1)) * kGemmRowsPerStep + rsub) * kGemmTile + 2 * chunk, src, bytes);
        }
        ++pstep;
        if (++pst == steps_per_tile) { pst = 0; ptile += gridDim.x; }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    static_assert((kGemmDepth & (kGemmDepth - 1)) == 0, "ring depth must be a power of two");
#pragma unroll 1
    for (int s = 0; s < kGemmDepth - 1; ++s) issue_step();
    double acc[kGemmAcc][4];
    const int cbase = half * 128 + 2 * lane;                 // this thread's columns: cbase, +1, cbase + 64, +65
    int st = 0;                                              // consumer: step inside the tile
    int64_t tile = blockIdx.x;
#pragma unroll 1
    for (int64_t step = 0; step < nsteps; ++step) {
        issue_step();
        asm volatile("cp.async.wait_group %0;" :: "n"(kGemmDepth - 1) : "memory");
        __syncthreads();                                     // the rows of `step` have landed (and ss on the first pass)
        const int64_t col = tile * kGemmTile + cbase;
        if (st == 0) {
#pragma unroll
            for (int c = 0; c < kGemmAcc; ++c) {
                acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.0;
                if (accumulate && grp * kGemmAcc + c < ncols) {
                    double* py = Y + (int64_t)(grp * kGemmAcc + c) * ldy + col;
                    if (col + 1 < M) { const double2 y = ld_stream2_rw(py); acc[c][0] = y.x; acc[c][1] = y.y; }
                    else if (col < M) acc[c][0] = py[0];
                    if (col + 65 < M) { const double2 y = ld_stream2_rw(py + 64); acc[c][2] = y.x; acc[c][3] = y.y; }
                    else if (col + 64 < M) acc[c][2] = py[64];
                }
            }
        }
        const double* rows = ring + (size_t)((step & (kGemmDepth - 1)) * kGemmRowsPerStep) * kGemmTile + cbase;
        const double* sr = ss + (size_t)(st * kGemmRowsPerStep) * kGemmOut + grp * kGemmAcc;
#pragma unroll
        for (int u = 0; u < kGemmRowsPerStep; ++u) {
            const double2 va = *reinterpret_cast<const double2*>(rows + (size_t)u * kGemmTile);
            const double2 vb = *reinterpret_cast<const double2*>(rows + (size_t)u * kGemmTile + 64);
#pragma unroll
            for (int c = 0; c < kGemmAcc; c += 2) {
                const double2 sc = *reinterpret_cast<const double2*>(sr + (size_t)u * kGemmOut + c);
                acc[c][0] = fma(sc.x, va.x, acc[c][0]);
                acc[c][1] = fma(sc.x, va.y, acc[c][1]);
                acc[c][2] = fma(sc.x, vb.x, acc[c][2]);
                acc[c][3] = fma(sc.x, vb.y, acc[c][3]);
                acc[c + 1][0] = fma(sc.y, va.x, acc[c + 1][0]);
                acc[c + 1][1] = fma(sc.y, va.y, acc[c + 1][1]);
                acc[c + 1][2] = fma(sc.y, vb.x, acc[c + 1][2]);
                acc[c + 1][3] = fma(sc.y, vb.y, acc[c + 1][3]);
            }
        }
        if (++st == steps_per_tile) {
            st = 0;
            tile += gridDim.x;
#pragma unroll
            for (int c = 0; c < kGemmAcc; ++c) {
                if (grp * kGemmAcc + c < ncols) {
                    double* py = Y + (int64_t)(grp * kGemmAcc + c) * ldy + col;
                    if (col + 1 < M) st_stream2(py, make_double2(acc[c][0], acc[c][1]));
                    else if (col < M) py[0] = acc[c][0];
                    if (col + 65 < M) st_stream2(py + 64, make_double2(acc[c][2], acc[c][3]));
                    else if (col + 64 < M) py[64] = acc[c][2];
                }
            }
        }
        __syncthreads();                                     // everybody is done with the slots the next step refills
    }
}

// K5 on the fp64 tensor cores: the same tiling and cp.async ring as ritz_lift_gemm_kernel, the inner product as
// mma.sync.m8n8k4.f64 - D (8 Ritz vectors x 8 columns) += A (8 x 4: S^T) . B (4 x 8: basis rows x columns).  A warp
// owns 32 Ritz vectors x 64 columns = 4 x 8 accumulator tiles; per pipeline step (four basis rows = one k-step) it
// loads 4 A and 8 B fragments (one double per lane each, from rows padded by 4 doubles so that the 16 lanes of a
// half-warp hit 16 different 8-byte banks) and issues 32 DMMAs: 44 instructions where the FMA form needs 152.
constexpr int kMmaSsLd = kGemmOut + 4;        // padded row strides (doubles)
constexpr int kMmaRingLd = kGemmTile + 4;

__device__ __forceinline__ void dmma8x8x4(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(kThreads, 1)
ritz_lift_mma_kernel(const double* __restrict__ V, int64_t ldv, int n, int64_t M, const double* __restrict__ S,
                     int lds, int ncols, double* Y, int64_t ldy, int accumulate, int64_t ntiles) {
    extern __shared__ __align__(128) double gsm[];
    double* ss = gsm;                                            // [nr4][kMmaSsLd], zero beyond n rows / ncols columns
    double* ring = gsm + (size_t)kGemmRowsSmem * kMmaSsLd;       // [kGemmRing][kMmaRingLd]
    const int nr4 = (n + kGemmRowsPerStep - 1) / kGemmRowsPerStep * kGemmRowsPerStep;
    for (int q = threadIdx.x; q < nr4 * kGemmOut; q += kThreads) {
        const int r = q / kGemmOut, c = q % kGemmOut;
        ss[r * kMmaSsLd + c] = (r < n && c < ncols) ? S[(int64_t)c * lds + r] : 0.0;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = 32 * (warp & 1), m0 = 64 * (warp >> 1);
    const int gid = lane >> 2, tig = lane & 3;
    const int steps_per_tile = nr4 / kGemmRowsPerStep;
    const int my_tiles = (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
    const int64_t nsteps = (int64_t)my_tiles * steps_per_tile;
    // producer: thread t copies 16-byte chunk (t & 127) of rows (t >> 7) and (t >> 7) + 2 of every step
    const int chunk = threadIdx.x & 127, rsub = threadIdx.x >> 7;
    int64_t pstep = 0;
    int pst = 0;
    int64_t ptile = blockIdx.x;
    auto issue_step = [&]() {
        if (pstep < nsteps) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int rr = rsub + 2 * u;
                const int r = pst * kGemmRowsPerStep + rr;
                const int64_t col = ptile * kGemmTile + 2 * chunk;
                const int64_t left = M - col;
                const int bytes = (r < n) ? (left >= 2 ? 16 : (left == 1 ? 8 : 0)) : 0;
                const double* src = bytes ? V + (int64_t)r * ldv + col : V;
                cp_async16_zfill(ring + (size_t)((pstep & (kGemmDepth - 1)) * kGemmRowsPerStep + rr) * kMmaRingLd + 2 * chunk, src, bytes);
            }
        }
        ++pstep;
        if (++pst == steps_per_tile) { pst = 0; ptile += gridDim.x; }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll 1
    for (int s = 0; s < kGemmDepth - 1; ++s) issue_step();
    double acc[4][8][2];
    int st = 0;
    int64_t tile = blockIdx.x;
#pragma unroll 1
    for (int64_t step = 0; step < nsteps; ++step) {
        issue_step();
        asm volatile("cp.async.wait_group %0;" :: "n"(kGemmDepth - 1) : "memory");
        __syncthreads();
        const int64_t colbase = tile * kGemmTile + m0 + 2 * tig;         // + 8 * mj: this lane's two output columns
        if (st == 0) {
#pragma unroll
            for (int ci = 0; ci < 4; ++ci)
#pragma unroll
                for (int mj = 0; mj < 8; ++mj) {
                    acc[ci][mj][0] = acc[ci][mj][1] = 0.0;
                    const int c = c0 + 8 * ci + gid;
                    const int64_t col = colbase + 8 * mj;
                    if (accumulate && c < ncols) {
                        double* py = Y + (int64_t)c * ldy + col;
                        if (col + 1 < M) { const double2 y = ld_stream2_rw(py); acc[ci][mj][0] = y.x; acc[ci][mj][1] = y.y; }
                        else if (col < M) acc[ci][mj][0] = py[0];
                    }
                }
        }
        const double* rows = ring + (size_t)((step & (kGemmDepth - 1)) * kGemmRowsPerStep + tig) * kMmaRingLd + m0 + gid;
        const double* sr = ss + (size_t)(st * kGemmRowsPerStep + tig) * kMmaSsLd + c0 + gid;
        double a[4], b[8];
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) a[ci] = sr[8 * ci];
#pragma unroll
        for (int mj = 0; mj < 8; ++mj) b[mj] = rows[8 * mj];
#pragma unroll
        for (int ci = 0; ci < 4; ++ci)
#pragma unroll
            for (int mj = 0; mj < 8; ++mj) dmma8x8x4(acc[ci][mj], a[ci], b[mj]);
        if (++st == steps_per_tile) {
            st = 0;
            tile += gridDim.x;
#pragma unroll
            for (int ci = 0; ci < 4; ++ci)
#pragma unroll
                for (int mj = 0; mj < 8; ++mj) {
                    const int c = c0 + 8 * ci + gid;
                    const int64_t col = colbase + 8 * mj;
                    if (c < ncols) {
                        double* py = Y + (int64_t)c * ldy + col;
                        if (col + 1 < M) st_stream2(py, make_double2(acc[ci][mj][0], acc[ci][mj][1]));
                        else if (col < M) py[0] = acc[ci][mj][0];
                    }
                }
        }
        __syncthreads();
    }
}

int launch_ritz_lift(lz_ctx* ctx, const double* V, int64_t ldv, int n, int64_t M,
                     const double* S_dev, int k, double* Y, int64_t ldy) {
    const int vec_ok = ((((uintptr_t)V | (uintptr_t)Y) & 15) == 0) && ((ldv & 1) == 0) && ((ldy & 1) == 0);
    static const bool gemm_off = []() { const char* e = getenv("LZ_K5_GEMM"); return e && e[0] == '0'; }();
    // LZ_K5_MMA=0: the CUDA-core FMA form below (61.8 ms at 512^3, n = k = 60, against 49.5 ms)
    static const bool mma_off = []() { const char* e = getenv("LZ_K5_MMA"); return e && e[0] == '0'; }();
    if (vec_ok && !gemm_off && !mma_off) {
        const size_t smem = ((size_t)kGemmRowsSmem * kMmaSsLd + (size_t)kGemmRing * kMmaRingLd) * 8;
        LZ_CUDA(cudaFuncSetAttribute((const void*)ritz_lift_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int64_t ntiles = (M + kGemmTile - 1) / kGemmTile;
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ntiles, (int64_t)ctx->sms));
        for (int c0 = 0; c0 < k; c0 += kGemmOut) {
            const int nc = std::min(kGemmOut, k - c0);
            for (int w0 = 0; w0 < n; w0 += kGemmRowsSmem) {
                const int wn = std::min(kGemmRowsSmem, n - w0);
                ritz_lift_mma_kernel<<<grid, kThreads, smem, ctx->stream>>>(
                    V + (int64_t)w0 * ldv, ldv, wn, M, S_dev + (int64_t)c0 * n + w0, n, nc,
                    Y + (int64_t)c0 * ldy, ldy, w0 > 0, ntiles);
                LZ_CUDA(cudaGetLastError());
            }
        }
        return LZ_OK;
    }
    if (vec_ok && !gemm_off) {
        const size_t smem = ((size_t)kGemmRowsSmem * kGemmOut + (size_t)kGemmRing * kGemmTile) * 8;
        LZ_CUDA(cudaFuncSetAttribute((const void*)ritz_lift_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int64_t ntiles = (M + kGemmTile - 1) / kGemmTile;
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ntiles, (int64_t)ctx->sms));
        for (int c0 = 0; c0 < k; c0 += kGemmOut) {
            const int nc = std::min(kGemmOut, k - c0);
            for (int w0 = 0; w0 < n; w0 += kGemmRowsSmem) {
                const int wn = std::min(kGemmRowsSmem, n - w0);
                ritz_lift_gemm_kernel<<<grid, kGemmThreads, smem, ctx->stream>>>(
                    V + (int64_t)w0 * ldv, ldv, wn, M, S_dev + (int64_t)c0 * n + w0, n, nc,
                    Y + (int64_t)c0 * ldy, ldy, w0 > 0, ntiles);
                LZ_CUDA(cudaGetLastError());
            }
        }
        return LZ_OK;
    }
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
