// Windowed SELL: the x entries a sorting window needs are staged in shared memory, the column indices become
// 16-bit offsets into that stage.
//
// Why: on a locality-ordered matrix (BASELINE config 4: vertices in cell order, ~14 entries per row) the plain
// SELL kernel of spmv.cu is bound by the L1 tag stage, not by HBM - ncu on the 50 M-vertex graph: 15.8 sectors per
// gather request, l1tex 85 % busy, DRAM 36 %.  The rows of one window (sigma = 2048 consecutive rows) refer to a few
// contiguous runs of columns (their own neighbourhood in the row above/below and the plane above/below), ~11 K
// distinct entries.  So per window:
//   (1) the distinct 32-entry granules of x it refers to are copied to shared memory, one 256-byte bulk copy
//       (cp.async.bulk, completion on an mbarrier) per granule, two stages deep: the copies of window i + 1 are in
//       flight while window i is computed;
//   (2) the entries are gathered from shared memory through 16-bit stage indices (granule rank * 32 + col % 32):
//       2 instead of 4 bytes per stored entry from HBM, and no gather reaches L1.  The indices are stored in blocks
//       of [32 lanes][8 entries], so a lane fetches 8 of them with one 16-byte load; a chunk's width is padded to a
//       multiple of 8 with indices of a slot that holds 0.0;
//   (3) x[row] comes out of the stage too (the set-up adds the rows' own granules), the diagonal coefficients
//       are stored in (chunk, lane) order, and the window's y (consecutive rows) is collected
//       in shared memory and written out coalesced.
// Value-free operators only (every off-diagonal entry equal, spmv.cu: sell_detect_uniform): the kernel reads no
// values.  With LZ_SELLW_BANKS=0 the per-row summation order is the one of spmv_sell_dot_kernel and y is
// bit-identical to it; by default a row's entries are stored in a bank-aware order (sellw_bank_kernel).
//
// Built at operator creation when every window's granule set fits (sellw_build); otherwise the operator keeps the
// plain kernel.  LZ_SELL_WINDOW=0 turns the form off.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "fin.cuh"
#include "internal.h"
#include "tma.cuh"

namespace lz {

constexpr int kWinThreads = 1024;
constexpr int kWinWarps = kWinThreads / 32;
constexpr int kWinMaxGran = 448;        // largest granule set the set-up kernels collect (the stage sizes decide below)
constexpr int kWinHash = 2048;          // open-addressing table of the set-up kernels (power of two, > 2 * kWinMaxGran)
constexpr size_t kWinSmemMax = 232448 - 1024;   // opt-in dynamic shared memory of one CTA, less the static part

__device__ __forceinline__ uint4 ld_stream_u4(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}

// Stage the `ng` granules of a window into `buf` and arrive on `bar` (every thread of the CTA arrives once; the
// barrier's phase completes when all have and every byte has landed).  x 16-byte aligned: thread t issues the bulk
// copy of granule t.  Otherwise (a basis row of odd length): warp q copies granules q, q + 32, ... with 8-byte
// asynchronous copies.  The granule that straddles M, and ghost columns of a row shard, are copied by hand.
template <int NT>
__device__ __forceinline__ void stage_window(double* buf, uint32_t bar, const double* __restrict__ x,
                                             const double* __restrict__ xg, const int32_t* __restrict__ gran,
                                             int32_t g0, int ng, int32_t M, int32_t ncols, bool al16) {
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) buf[ng * 32] = 0.0;             // the slot the padding indices point at
    if (al16) {
        for (int gi = threadIdx.x; gi < ng; gi += NT) {
            const int32_t base = __ldg(gran + g0 + gi) * 32;
            double* dst = buf + gi * 32;
            if (base + 32 <= M) {
                mbar_expect_tx(bar, 256);
                bulk_load_1d(smem_addr(dst), x + base, 256, bar);
            } else {
                for (int l = 0; l < 32; ++l) {
                    const int32_t idx = base + l;
                    dst[l] = idx < M ? __ldg(x + idx) : ((idx < ncols && xg) ? __ldg(xg + (idx - M)) : 0.0);
                }
            }
        }
        mbar_arrive(bar);
        return;
    }
    const int mine_at = warp + (NT / 32) * lane;
    const int32_t mine = mine_at < ng ? __ldg(gran + g0 + mine_at) : 0;
    for (int it = 0, gi = warp; gi < ng; ++it, gi += NT / 32) {
        const int32_t base = __shfl_sync(0xffffffffu, mine, it) * 32;
        double* dst = buf + gi * 32;
        if (base + 32 <= M) {
            cp_async8(dst + lane, x + base + lane);
        } else {
            const int32_t idx = base + lane;
            dst[lane] = idx < M ? __ldg(x + idx) : ((idx < ncols && xg) ? __ldg(xg + (idx - M)) : 0.0);
        }
    }
    mbar_arrive_after_cp_async(bar);
}

// One chunk of 32 rows: what it reads from global memory is requested first (chunk_issue), consumed later
// (chunk_finish), so that a warp has the loads of two chunks in flight.
struct ChunkRegs {
    int c;                                                // chunk, -1: none
    uint4 q0, q1;                                         // the first 16 stage indices of the lane's row
    int b0;                                               // first block
    int nb;                                               // blocks of 8 entries
    uint32_t rr;                                          // row inside the window | stage index of x[row] << 16
    double dd;
};

// where a chunk's index blocks are (fetched a pair of chunks ahead of the indices themselves, which depend on it)
struct ChunkMeta {
    int c;                                                // chunk, -1: none
    int b0, nb;                                           // its blocks [b0, b0 + nb)
};

__device__ __forceinline__ ChunkMeta chunk_meta(int64_t c, int64_t c_end, const int64_t* __restrict__ off8) {
    ChunkMeta m;
    m.c = c < c_end ? (int)c : -1;
    m.b0 = m.c >= 0 ? (int)__ldg(off8 + c) : 0;
    m.nb = m.c >= 0 ? (int)__ldg(off8 + c + 1) - m.b0 : 0;
    return m;
}

__device__ __forceinline__ void chunk_issue(ChunkRegs& r, const ChunkMeta& m, const uint4* __restrict__ lc8,
                                            const uint32_t* __restrict__ lrow, const double* __restrict__ deff_p,
                                            int lane) {
    r.c = m.c;
    if (m.c < 0) return;
    const int64_t c = m.c;
    r.b0 = m.b0;
    r.nb = m.nb;
    const uint4* p = lc8 + (int64_t)m.b0 * 32 + lane;
    r.rr = __ldg(lrow + c * 32 + lane);
    r.dd = __ldg(deff_p + c * 32 + lane);
    r.q0 = r.nb > 0 ? ld_stream_u4(p) : make_uint4(0, 0, 0, 0);
    r.q1 = r.nb > 1 ? ld_stream_u4(p + 32) : make_uint4(0, 0, 0, 0);
}

// 8 entries: sum += stage[index], in entry order (the low half of a word is the earlier entry)
__device__ __forceinline__ void add8(double& sum, const uint4 q, const double* sx) {
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        sum += sx[w[i] & 0xffffu];
        sum += sx[w[i] >> 16];
    }
}

__device__ __forceinline__ void chunk_finish(const ChunkRegs& r, const double* sx, double* sy,
                                             const uint4* __restrict__ lc8, double s, double uni_a, int lane,
                                             double& acc) {
    if (r.c < 0) return;
    double sum = 0.0;
    if (r.nb > 0) add8(sum, r.q0, sx);
    if (r.nb > 1) add8(sum, r.q1, sx);
    for (int b = 2; b < r.nb; ++b) add8(sum, ld_stream_u4(lc8 + ((int64_t)r.b0 + b) * 32 + lane), sx);
    if (r.rr != 0xffffffffu) {
        const double xr = sx[r.rr >> 16];
        sum = fma(uni_a, sum, r.dd * xr);
        const double yi = s * sum;
        sy[r.rr & 0xffffu] = yi;
        acc = fma(yi, s * xr, acc);
    }
}

// Fallback for granule sets too large for two CTAs per SM (spmv_sellw2_dot_kernel below, the one that normally
// runs): one CTA of 32 warps per SM, two stages of x (each the largest granule set + the zero slot) and two buffers
// for the window's y; the copies of window i + 1 are in flight while window i is computed.  Measured on the
// config-4 graph: 1.11 ms per apply against 0.98 ms for the two-CTA form (plain SELL kernel: 1.55 ms) - with one
// CTA nothing fills the waits between the phases of a window.  LZ_SELLW_VARIANT=1 forces this form.
__global__ void __launch_bounds__(kWinThreads, 1)
spmv_sellw_dot_kernel(const uint4* __restrict__ lc8, const int64_t* __restrict__ off8,
                      const uint32_t* __restrict__ lrow, const double* __restrict__ x,
                      const double* __restrict__ scale, double* __restrict__ y, int64_t nchunks,
                      double* __restrict__ partials, const double* __restrict__ xg, int32_t M, int32_t ncols, int span,
                      const FinTail fin, const int32_t* __restrict__ win_list, int nlist, int64_t nwin,
                      const int* __restrict__ flag, const double* __restrict__ deff_p, double uni_a,
                      const int32_t* __restrict__ gran_off, const int32_t* __restrict__ gran, int stage_doubles) {
    pdl_prologue();
    if (flag && *flag == 0) return;
    extern __shared__ __align__(128) double sx[];
    __shared__ double red[kWinWarps];
    __shared__ __align__(8) unsigned long long bars[2];
    const int sigma = span * 32;
    double* const sy0 = sx + 2 * stage_doubles;
    const double s = scale ? __ldg(scale) : 1.0;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const bool al16 = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
    const int64_t nitems = win_list ? (int64_t)nlist : nwin;
    const uint32_t bar0 = smem_addr(&bars[0]), bar1 = smem_addr(&bars[1]);
    if (threadIdx.x == 0) {
        mbar_init(bar0, kWinThreads);
        mbar_init(bar1, kWinThreads);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    double acc = 0.0;
    auto window_of = [&](int64_t si) { return win_list ? (int64_t)__ldg(win_list + si) : si; };
    int64_t si = blockIdx.x;
    if (si < nitems) {
        const int64_t w = window_of(si);
        const int32_t g0 = __ldg(gran_off + w);
        stage_window<kWinThreads>(sx, bar0, x, xg, gran, g0, __ldg(gran_off + w + 1) - g0, M, ncols, al16);
    }
    for (int it = 0; si < nitems; si += gridDim.x, ++it) {
        const int64_t w = window_of(si);
        const int64_t sn = si + gridDim.x;
        if (sn < nitems) {                                // next window into the other stage (free since the barrier below)
            const int64_t wn = window_of(sn);
            const int32_t g0 = __ldg(gran_off + wn);
            stage_window<kWinThreads>(sx + ((it + 1) & 1) * stage_doubles, (it & 1) ? bar0 : bar1, x, xg, gran, g0,
                                      __ldg(gran_off + wn + 1) - g0, M, ncols, al16);
        }
        const int64_t c_end = min(nchunks, (w + 1) * span);
        int64_t c = w * span + warp;
        ChunkRegs a, b;
        chunk_issue(a, chunk_meta(c, c_end, off8), lc8, lrow, deff_p, lane);
        chunk_issue(b, chunk_meta(c + kWinWarps, c_end, off8), lc8, lrow, deff_p, lane);
        mbar_wait((it & 1) ? bar1 : bar0, (uint32_t)(it >> 1) & 1u);
        const double* cur = sx + (it & 1) * stage_doubles;
        double* sy = sy0 + (it & 1) * sigma;
        for (;;) {
            c += 2 * kWinWarps;
            const ChunkMeta ma = chunk_meta(c, c_end, off8), mb = chunk_meta(c + kWinWarps, c_end, off8);
            chunk_finish(a, cur, sy, lc8, s, uni_a, lane, acc);
            chunk_finish(b, cur, sy, lc8, s, uni_a, lane, acc);
            if (ma.c < 0) break;
            chunk_issue(a, ma, lc8, lrow, deff_p, lane);
            chunk_issue(b, mb, lc8, lrow, deff_p, lane);
        }
        __syncthreads();                                  // sy is complete; this stage may be overwritten from now on
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // ... by bulk copies (async proxy) as well
        const int64_t r0 = w * sigma;
        const int rows = (int)min((int64_t)sigma, (int64_t)M - r0);
        for (int t = threadIdx.x; t < rows; t += kWinThreads) y[r0 + t] = sy[t];
    }
    // CTA sum in warp order; the bookkeeping tail is written for kThreads threads, so the upper warps leave first
    acc = warp_sum(acc);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (warp >= kWarps) return;
    if (threadIdx.x == 0 && partials) {
        double t = 0.0;
#pragma unroll
        for (int q = 0; q < kWinWarps; ++q) t += red[q];
        partials[blockIdx.x] = t;
    }
    fin_tail(fin, partials, red);
}

// Two CTAs of 16 warps per SM, each with ONE stage of x and one buffer for y: a CTA's phases (stage the window,
// fetch the indices, gather out of shared memory, write y) follow each other, and the two CTAs of an SM fill each
// other's waits.  Windows are dealt round-robin over the CTAs (neighbouring CTAs work on neighbouring windows: what
// they stage overlaps and comes out of L2).  Needs (granules * 256 + sigma * 8) bytes twice per SM.
constexpr int kWin2Threads = 512;
constexpr int kWin2Warps = kWin2Threads / 32;

__global__ void __launch_bounds__(kWin2Threads, 2)
spmv_sellw2_dot_kernel(const uint4* __restrict__ lc8, const int64_t* __restrict__ off8,
                       const uint32_t* __restrict__ lrow, const double* __restrict__ x,
                       const double* __restrict__ scale, double* __restrict__ y, int64_t nchunks,
                       double* __restrict__ partials, const double* __restrict__ xg, int32_t M, int32_t ncols, int span,
                       const FinTail fin, const int32_t* __restrict__ win_list, int nlist, int64_t nwin,
                       const int* __restrict__ flag, const double* __restrict__ deff_p, double uni_a,
                       const int32_t* __restrict__ gran_off, const int32_t* __restrict__ gran, int stage_doubles) {
    pdl_prologue();
    if (flag && *flag == 0) return;
    extern __shared__ __align__(128) double sx[];
    __shared__ double red[kWin2Warps];
    __shared__ __align__(8) unsigned long long bars[1];
    const int sigma = span * 32;
    double* const sy = sx + stage_doubles;
    const double s = scale ? __ldg(scale) : 1.0;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const bool al16 = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
    const int64_t nitems = win_list ? (int64_t)nlist : nwin;
    const uint32_t bar = smem_addr(&bars[0]);
    if (threadIdx.x == 0) {
        mbar_init(bar, kWin2Threads);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    double acc = 0.0;
    auto window_of = [&](int64_t si) { return win_list ? (int64_t)__ldg(win_list + si) : si; };
    // the window after this one: its id and granule range are fetched an iteration ahead
    int64_t si = blockIdx.x;
    int64_t w = si < nitems ? window_of(si) : 0;
    int32_t g0 = si < nitems ? __ldg(gran_off + w) : 0;
    int32_t g1 = si < nitems ? __ldg(gran_off + w + 1) : 0;
    for (int it = 0; si < nitems; si += gridDim.x, ++it) {
        stage_window<kWin2Threads>(sx, bar, x, xg, gran, g0, g1 - g0, M, ncols, al16);
        const int64_t c_end = min(nchunks, (w + 1) * span);
        int64_t c = w * span + warp;
        ChunkRegs a, b;
        chunk_issue(a, chunk_meta(c, c_end, off8), lc8, lrow, deff_p, lane);
        chunk_issue(b, chunk_meta(c + kWin2Warps, c_end, off8), lc8, lrow, deff_p, lane);
        const int64_t w_cur = w;
        const int64_t sn = si + gridDim.x;
        if (sn < nitems) {
            w = window_of(sn);
            g0 = __ldg(gran_off + w);
            g1 = __ldg(gran_off + w + 1);
        }
        mbar_wait(bar, (uint32_t)it & 1u);
        for (;;) {
            c += 2 * kWin2Warps;
            const ChunkMeta ma = chunk_meta(c, c_end, off8), mb = chunk_meta(c + kWin2Warps, c_end, off8);
            chunk_finish(a, sx, sy, lc8, s, uni_a, lane, acc);
            chunk_finish(b, sx, sy, lc8, s, uni_a, lane, acc);
            if (ma.c < 0) break;
            chunk_issue(a, ma, lc8, lrow, deff_p, lane);
            chunk_issue(b, mb, lc8, lrow, deff_p, lane);
        }
        __syncthreads();                                  // sy is complete, the stage is free
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        const int64_t r0 = w_cur * sigma;
        const int rows = (int)min((int64_t)sigma, (int64_t)M - r0);
        for (int t = threadIdx.x; t < rows; t += kWin2Threads) y[r0 + t] = sy[t];
        // the next write to sy follows the wait for the next stage, whose barrier needs every thread's arrival -
        // made after this write-out
    }
    acc = warp_sum(acc);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (warp >= kWarps) return;
    if (threadIdx.x == 0 && partials) {
        double t = 0.0;
#pragma unroll
        for (int q = 0; q < kWin2Warps; ++q) t += red[q];
        partials[blockIdx.x] = t;
    }
    fin_tail(fin, partials, red);
}

// ---- set-up ---------------------------------------------------------------------------------------------
// blocks of 8 entries per chunk (to be scanned)
__global__ void __launch_bounds__(kThreads)
sellw_blocks_kernel(const int64_t* __restrict__ chunk_off, int64_t nchunks, int64_t* __restrict__ off8) {
    const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (c > nchunks) return;
    off8[c] = c < nchunks ? (((chunk_off[c + 1] - chunk_off[c]) >> 5) + 7) / 8 : 0;
}

// One CTA per window.  Collects the distinct granules (index / 32) of the window's stored columns and of its own rows
// in a shared hash set; COUNT: writes how many (kWinMaxGran + 1 = too many); otherwise writes the sorted granule
// list at gran[gran_off[w] ..], the 16-bit stage index of every stored entry in blocks of [32 lanes][8 entries]
// (padding: the zero slot behind the last granule), and per (chunk, lane) the row's place in the window, the stage
// index of x[row] and its diagonal coefficient.
template <bool COUNT>
__global__ void __launch_bounds__(kThreads)
sellw_setup_kernel(const int64_t* __restrict__ chunk_off, const int32_t* __restrict__ col,
                   const int32_t* __restrict__ row_of, const double* __restrict__ deff, int64_t nchunks, int span,
                   int32_t* __restrict__ gcount, const int32_t* __restrict__ gran_off, int32_t* __restrict__ gran,
                   const int64_t* __restrict__ off8, uint16_t* __restrict__ lc8, uint32_t* __restrict__ lrow,
                   double* __restrict__ deff_p) {
    __shared__ int32_t table[kWinHash];
    __shared__ int32_t keys[kWinMaxGran + 1];
    __shared__ int32_t sorted[kWinMaxGran + 1];
    __shared__ int count;
    const int64_t w = blockIdx.x;
    const int64_t c0 = w * span, c1 = min(nchunks, c0 + span);
    const int64_t e0 = chunk_off[c0], e1 = chunk_off[c1];
    bool skipped = false;                                  // a window that stays with the plain kernel
    if constexpr (!COUNT) skipped = gran_off[w + 1] == gran_off[w];
    for (int i = threadIdx.x; i < kWinHash; i += kThreads) table[i] = -1;
    if (threadIdx.x == 0) count = 0;
    __syncthreads();
    auto insert = [&](int32_t g) {
        uint32_t h = ((uint32_t)g * 2654435761u) & (kWinHash - 1);
        while (*(volatile int*)&count <= kWinMaxGran) {    // too many: stop collecting, the table must not fill up
            const int32_t prev = atomicCAS(&table[h], -1, g);
            if (prev == g) break;
            if (prev == -1) {
                const int at = atomicAdd(&count, 1);
                if (at <= kWinMaxGran) keys[at] = g;
                break;
            }
            h = (h + 1) & (kWinHash - 1);
        }
    };
    if (!skipped) {
        for (int64_t e = e0 + threadIdx.x; e < e1; e += kThreads) insert(col[e] >> 5);
        for (int64_t t = c0 * 32 + threadIdx.x; t < c1 * 32; t += kThreads) {
            const int32_t row = row_of[t];
            if (row >= 0) insert(row >> 5);
        }
    }
    __syncthreads();
    const int n = min(count, kWinMaxGran + 1);
    if constexpr (COUNT) {
        if (threadIdx.x == 0) gcount[w] = n;
        return;
    } else {
        // rank sort (n <= kWinMaxGran distinct keys)
        for (int i = threadIdx.x; i < n; i += kThreads) {
            const int32_t mine = keys[i];
            int r = 0;
            for (int j = 0; j < n; ++j) r += keys[j] < mine;
            sorted[r] = mine;
        }
        __syncthreads();
        auto stage_index = [&](int32_t idx) {
            const int32_t g = idx >> 5;
            int lo = 0, hi = n - 1;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (sorted[mid] < g) lo = mid + 1;
                else hi = mid;
            }
            return (uint32_t)((lo << 5) | (idx & 31));
        };
        const int32_t base = gran_off[w];
        for (int i = threadIdx.x; i < n; i += kThreads) gran[base + i] = sorted[i];
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const uint16_t zero_slot = (uint16_t)(n * 32);
        const int64_t r0 = w * (int64_t)span * 32;
        for (int64_t c = c0 + warp; c < c1; c += kWarps) {
            const int64_t o0 = chunk_off[c];
            const int width = (int)((chunk_off[c + 1] - o0) >> 5);
            const int64_t b0 = off8[c];
            const int nb = (int)(off8[c + 1] - b0);
            for (int k = 0; k < nb * 8; ++k) {
                // a skipped window's indices are never used: all zero
                const uint16_t v = skipped ? (uint16_t)0
                                           : (k < width ? (uint16_t)stage_index(col[o0 + (int64_t)k * 32 + lane]) : zero_slot);
                lc8[((b0 + (k >> 3)) * 32 + lane) * 8 + (k & 7)] = v;
            }
            const int32_t row = row_of[c * 32 + lane];
            lrow[c * 32 + lane] = (row >= 0 && !skipped) ? ((uint32_t)(row - r0) | (stage_index(row) << 16)) : 0xffffffffu;
            if (deff_p) deff_p[c * 32 + lane] = (row >= 0 && deff) ? deff[row] : 0.0;
        }
    }
}

// Bank-aware order of a row's entries (value-free form only: the stored values of a weighted operator stay in the
// SELL order).  A gather of 16 lanes x 8 bytes is one pass through shared memory when the 16 addresses fall into 16
// different 8-byte banks (stage index mod 16); with the entries in column order it takes ~3.1 passes (random
// banks).  Per half-chunk (16 rows, S = 8 or 16 slots): slot by slot, a maximum bipartite matching lanes -> banks
// (augmenting paths) over the entries not placed yet picks one entry per lane with pairwise different banks; a
// lane that stays unmatched takes the zero slot if it can still afford to (S minus its row length spare slots),
// else its entry from the least loaded bank.  Simulated on random banks: 1.6 passes per slot.  One thread per
// half-chunk; rows longer than 16 entries keep their order.  The sums of a row are then added in a different
// (fixed) order than in the plain kernel: y differs from it by rounding.
__global__ void __launch_bounds__(kThreads)
sellw_bank_kernel(const int64_t* __restrict__ off8, const int32_t* __restrict__ gran_off, int64_t nchunks, int span,
                  uint16_t* __restrict__ lc8) {
    const int64_t h = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int64_t c = h >> 1;
    if (c >= nchunks) return;
    const int half = (int)(h & 1);
    const int64_t w = c / span;
    const int ng = gran_off[w + 1] - gran_off[w];
    if (ng == 0) return;                                   // a window that stays with the plain kernel
    const uint16_t zero_slot = (uint16_t)(ng * 32);
    const int64_t b0 = off8[c];
    const int nb = (int)(off8[c + 1] - b0);
    if (nb < 1 || nb > 2) return;
    const int S = nb * 8;
    uint16_t ent[16][16];
    int cnt[16];
    auto at = [&](int l, int k) -> uint16_t& { return lc8[((b0 + (k >> 3)) * 32 + (half * 16 + l)) * 8 + (k & 7)]; };
    for (int l = 0; l < 16; ++l) {
        int n = 0;
        for (int k = 0; k < S; ++k) {
            const uint16_t v = at(l, k);
            if (v != zero_slot) ent[l][n++] = v;
        }
        cnt[l] = n;
    }
    for (int k = 0; k < S; ++k) {
        uint32_t avail[16];
        int owner[16];
        for (int l = 0; l < 16; ++l) {
            uint32_t m = 0;
            for (int i = 0; i < cnt[l]; ++i) m |= 1u << (ent[l][i] & 15);
            avail[l] = m;
            owner[l] = -1;
        }
        for (int l0 = 0; l0 < 16; ++l0) {
            if (cnt[l0] == 0) continue;
            int st_l[17], st_b[17];
            uint32_t st_m[17];
            uint32_t visited = 0;
            int depth = 0;
            st_l[0] = l0;
            st_m[0] = avail[l0];
            while (depth >= 0) {
                const uint32_t m = st_m[depth] & ~visited;
                if (!m) { --depth; continue; }
                const int b = __ffs(m) - 1;
                visited |= 1u << b;
                st_m[depth] = m & ~(1u << b);
                st_b[depth] = b;
                if (owner[b] < 0) {
                    for (int d = depth; d >= 0; --d) owner[st_b[d]] = st_l[d];
                    break;
                }
                if (depth + 1 < 17) {
                    ++depth;
                    st_l[depth] = owner[b];
                    st_m[depth] = avail[owner[b]];
                }
            }
        }
        int bank_of[16], load[16];
        for (int l = 0; l < 16; ++l) { bank_of[l] = -1; load[l] = 0; }
        for (int b = 0; b < 16; ++b)
            if (owner[b] >= 0) { bank_of[owner[b]] = b; load[b] = 1; }
        for (int l = 0; l < 16; ++l) {
            int pick = -1;
            if (bank_of[l] >= 0) {
                for (int i = 0; i < cnt[l]; ++i)
                    if ((ent[l][i] & 15) == bank_of[l]) { pick = i; break; }
            } else if (cnt[l] >= S - k) {                  // no spare slot left: the entry from the least loaded bank
                int best = 1 << 30;
                for (int i = 0; i < cnt[l]; ++i)
                    if (load[ent[l][i] & 15] < best) { best = load[ent[l][i] & 15]; pick = i; }
                if (pick >= 0) ++load[ent[l][pick] & 15];
            }
            uint16_t v = zero_slot;
            if (pick >= 0) {
                v = ent[l][pick];
                ent[l][pick] = ent[l][cnt[l] - 1];
                --cnt[l];
            }
            at(l, k) = v;
        }
    }
}

int sellw_build(lz_op* op) {
    // read per call: LZ_SELL_WINDOW=0 keeps the plain kernel, LZ_SELL_WINDOW_MIN overrides the smallest number of
    // windows the form is built for (tests build small operators)
    const char* env = getenv("LZ_SELL_WINDOW");
    const bool off = env && env[0] == '0';
    const char* env_min = getenv("LZ_SELL_WINDOW_MIN");
    lz_ctx* ctx = op->ctx;
    lz_sell& sl = op->sell;
    if (off || op->kind != LZ_OP_SELL || sl.nchunks == 0 || sl.nnz_stored == 0 || sl.sigma % 32 != 0 || sl.sigma > 8192) return LZ_OK;
    // value-free operators only: with stored values the plain kernel already runs at 73 % of the copy peak on the
    // config-4 graph (1.92 ms); this form with the values fetched in SELL order measured 2.08 ms
    if (!sl.uniform) return LZ_OK;
    const int span = sl.sigma / 32;
    const int64_t nwin = (sl.nchunks + span - 1) / span;
    // a window per CTA only pays when there are enough of them to fill the machine a few times over
    if (nwin < (env_min ? (int64_t)atoll(env_min) : (int64_t)ctx->sms * 4)) return LZ_OK;
    // row shards: the windows with ghost columns stay with the plain kernel in pieces (launch_spmv_part); the
    // interior list must consist of whole windows
    if (op->ncols > op->M && !(sl.split_span == span)) return LZ_OK;
    cudaStream_t q = ctx->stream;
    int32_t* gcount = nullptr;
    LZ_CUDA(cudaMalloc((void**)&gcount, (size_t)(nwin + 1) * 4));
    sellw_setup_kernel<true><<<(unsigned)nwin, kThreads, 0, q>>>(sl.chunk_off, sl.col, sl.row_of, nullptr, sl.nchunks, span,
                                                               gcount, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    std::vector<int32_t> h((size_t)nwin + 1, 0);
    cudaError_t e = cudaMemcpyAsync(h.data(), gcount, (size_t)nwin * 4, cudaMemcpyDeviceToHost, q);
    if (e == cudaSuccess) e = cudaStreamSynchronize(q);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { cudaFree(gcount); set_error("sellw_build: %s", cudaGetErrorString(e)); return LZ_ERR_CUDA; }
    // row shards: the boundary windows are never run in this form, so they need not fit
    if (op->ncols > op->M && sl.n_bnd > 0) {
        std::vector<int32_t> bd((size_t)sl.n_bnd);
        e = cudaMemcpyAsync(bd.data(), sl.spans_bnd, bd.size() * 4, cudaMemcpyDeviceToHost, q);
        if (e == cudaSuccess) e = cudaStreamSynchronize(q);
        if (e != cudaSuccess) { cudaFree(gcount); set_error("sellw_build: %s", cudaGetErrorString(e)); return LZ_ERR_CUDA; }
        for (int32_t piece : bd) h[(size_t)((int64_t)piece * sl.bnd_span / span)] = 0;
    }
    int maxg = 0;
    int64_t total = 0;
    for (int64_t w = 0; w < nwin; ++w) { maxg = std::max(maxg, h[(size_t)w]); total += h[(size_t)w]; }
    // two stages of x (granules + the zero slot, kept 16-byte aligned) and two buffers for the window's y must fit
    // the shared memory of one CTA; otherwise: plain kernel
    const int stage_doubles = maxg * 32 + 2;
    const size_t smem = (2 * (size_t)stage_doubles + 2 * (size_t)sl.sigma) * sizeof(double);
    if (maxg > kWinMaxGran || smem > kWinSmemMax || total > (int64_t)INT32_MAX / 2) { cudaFree(gcount); return LZ_OK; }
    int32_t run = 0;
    for (int64_t w = 0; w <= nwin; ++w) { const int32_t v = h[(size_t)w]; h[(size_t)w] = run; run += v; }
    int32_t* gran = nullptr;
    uint16_t* lc8 = nullptr;
    uint32_t* lrow = nullptr;
    int64_t* off8 = nullptr;
    double* deff_p = nullptr;
    int64_t blocks = 0;
    cudaError_t e2 = cudaMalloc((void**)&off8, (size_t)(sl.nchunks + 1) * 8);
    if (e2 == cudaSuccess) {
        sellw_blocks_kernel<<<(unsigned)((sl.nchunks + 1 + kThreads - 1) / kThreads), kThreads, 0, q>>>(sl.chunk_off, sl.nchunks, off8);
        if (scan_i64(off8, sl.nchunks + 1, q) != LZ_OK) e2 = cudaErrorUnknown;
    }
    if (e2 == cudaSuccess) e2 = cudaMemcpyAsync(&blocks, off8 + sl.nchunks, 8, cudaMemcpyDeviceToHost, q);
    if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(q);
    if (e2 == cudaSuccess) e2 = cudaMalloc((void**)&gran, (size_t)std::max<int64_t>(total, 4) * 4);
    if (e2 == cudaSuccess) e2 = cudaMalloc((void**)&lc8, (size_t)std::max<int64_t>(blocks, 1) * 256 * 2);
    if (e2 == cudaSuccess) e2 = cudaMalloc((void**)&lrow, (size_t)sl.nchunks * 32 * 4);
    if (e2 == cudaSuccess) e2 = cudaMalloc((void**)&deff_p, (size_t)sl.nchunks * 32 * 8);
    if (e2 == cudaSuccess) e2 = cudaMemcpyAsync(gcount, h.data(), (size_t)(nwin + 1) * 4, cudaMemcpyHostToDevice, q);
    if (e2 == cudaSuccess) {
        sellw_setup_kernel<false><<<(unsigned)nwin, kThreads, 0, q>>>(sl.chunk_off, sl.col, sl.row_of, sl.deff, sl.nchunks, span,
                                                                    nullptr, gcount, gran, off8, lc8, lrow, deff_p);
        const char* env_banks = getenv("LZ_SELLW_BANKS");
        if (!(env_banks && env_banks[0] == '0')) {
            sellw_bank_kernel<<<(unsigned)((2 * sl.nchunks + kThreads - 1) / kThreads), kThreads, 0, q>>>(off8, gcount, sl.nchunks, span, lc8);
            sl.win_banked = 1;
        }
        e2 = cudaStreamSynchronize(q);
        if (e2 == cudaSuccess) e2 = cudaGetLastError();
    }
    if (e2 == cudaSuccess) e2 = cudaFuncSetAttribute(spmv_sellw_dot_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWinSmemMax);
    if (e2 == cudaSuccess) e2 = cudaFuncSetAttribute(spmv_sellw2_dot_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kWinSmemMax / 2));
    if (e2 != cudaSuccess) {
        cudaGetLastError();
        if (gran) cudaFree(gran);
        if (lc8) cudaFree(lc8);
        if (lrow) cudaFree(lrow);
        if (off8) cudaFree(off8);
        if (deff_p) cudaFree(deff_p);
        cudaFree(gcount);
        if (e2 == cudaErrorMemoryAllocation) return LZ_OK;   // no room for the second index array: plain kernel
        set_error("sellw_build: %s", cudaGetErrorString(e2));
        return LZ_ERR_CUDA;
    }
    sl.win_gran_off = gcount;
    sl.win_gran = gran;
    sl.win_lcol = lc8;
    sl.win_off8 = off8;
    sl.win_blocks = blocks;
    sl.win_lrow = lrow;
    sl.win_deff = deff_p;
    sl.win_span = span;
    sl.win_count = nwin;
    sl.win_maxg = maxg;
    sl.win_total = total;
    sl.win_smem = smem;
    sl.windowed = (op->ncols > op->M) ? 2 : 1;         // 2: only the interior list of a row shard (launch_spmv_part)
    return LZ_OK;
}

// All windows (win_list == nullptr) or the listed ones.  Same contract as the plain launches of spmv.cu.
int launch_spmv_windowed(lz_op* op, const int32_t* win_list, int nlist, const double* x, const double* scale_dev,
                         double* y, double* partials, int* grid_out, const int* flag_dev, const FinTail& ft,
                         cudaStream_t stream) {
    lz_ctx* ctx = op->ctx;
    const lz_sell& sl = op->sell;
    const int64_t nitems = win_list ? (int64_t)nlist : sl.win_count;
    int grid = (int)std::max<int64_t>(1, std::min<int64_t>(nitems, (int64_t)ctx->sms));
    const int stage_doubles = sl.win_maxg * 32 + 2;
    const char* var_env = getenv("LZ_SELLW_VARIANT");
    const size_t smem2 = ((size_t)stage_doubles + (size_t)sl.sigma) * sizeof(double);
    const bool two_ctas = smem2 <= kWinSmemMax / 2 - 1024 && !(var_env && var_env[0] == '1');
    if (two_ctas) {
        grid = (int)std::max<int64_t>(1, std::min<int64_t>(nitems, (int64_t)ctx->sms * 2));
        LZ_CUDA(launch_k(spmv_sellw2_dot_kernel, dim3(grid), dim3(kWin2Threads), smem2, stream, (const uint4*)sl.win_lcol,
                         (const int64_t*)sl.win_off8, (const uint32_t*)sl.win_lrow, x, scale_dev, y, sl.nchunks, partials,
                         op->xghost, (int32_t)op->M, (int32_t)op->ncols, sl.win_span, ft, win_list, nlist, sl.win_count,
                         flag_dev, (const double*)sl.win_deff, sl.uni_a, (const int32_t*)sl.win_gran_off,
                         (const int32_t*)sl.win_gran, stage_doubles));
    } else {
        LZ_CUDA(launch_k(spmv_sellw_dot_kernel, dim3(grid), dim3(kWinThreads), sl.win_smem, stream, (const uint4*)sl.win_lcol,
                         (const int64_t*)sl.win_off8, (const uint32_t*)sl.win_lrow, x, scale_dev, y, sl.nchunks, partials,
                         op->xghost, (int32_t)op->M, (int32_t)op->ncols, sl.win_span, ft, win_list, nlist, sl.win_count,
                         flag_dev, (const double*)sl.win_deff, sl.uni_a, (const int32_t*)sl.win_gran_off,
                         (const int32_t*)sl.win_gran, stage_doubles));
    }
    if (grid_out) *grid_out = grid;
    return LZ_OK;
}

}  // namespace lz
