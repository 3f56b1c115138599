// KF: the whole plain Lanczos step in ONE pass over HBM (structured grids, no sweep this step).
//
// The two-pass step (K1 then K3) moves 48*N bytes because w = H v_j has to be complete before
// r = w - alpha v_j - beta v_{j-1} can be applied with alpha = v_j.w.  H is linear, so the
// operator can be applied to the UN-normalised residual as soon as it is formed:
//
//     r_{j+1} = s_j u_j - (alpha_j s_j) r_j - (beta_j s_{j-1}) r_{j-1}        (s_j = 1/beta_j)
//     u_{j+1} = H r_{j+1}
//     partials:  sum r_{j+1}^2  (-> beta_{j+1}),   sum r_{j+1} . u_{j+1}  (-> alpha_{j+1} = s^2 . )
//
// with alpha_j, beta_j already known from the partials of the previous launch.  One kernel reads
// u_j, r_j, r_{j-1} and writes r_{j+1}, u_{j+1}: 40*N bytes per step instead of 48*N, one launch
// and one fin kernel per step.  (Reference lines folded: Lanczos.py:112-113,116,118,119.)
//
// Structure (sm_100a): a CTA of 10 warps, one per tile row (8 rows + 2 halo rows).
//   * Input: the three vectors are streamed plane by plane into a 3-stage shared-memory ring with
//     16-byte cp.async copies (LDGSTS), every thread copying exactly the operands it reads back
//     itself, two planes ahead: completion is a per-thread cp.async.wait_group, no barrier, and
//     the bytes in flight per SM are set by the ring depth, not by registers or occupancy.
//     (The alternatives that were measured - plain loads, 1-D bulk-TMA row copies, register tiles - are in
//     DESIGN.md section 8.)
//   * Compute: warps 1..8 own the tile rows, warps 0 and 9 recompute r on the two halo rows, lanes
//     0 / 31 on the two halo columns, the first / last iteration on the two halo planes - the
//     stencil needs r_{j+1} on the tile's one-point halo and that is cheaper to recompute (extra
//     L2 reads, no extra HBM traffic) than to exchange.  r planes live in a 3-slot shared-memory
//     ring (one named barrier per plane, which also frees the input stage); the z-1 / z / z+1
//     values of a thread's own column stay in registers, x neighbours travel by shuffle.
// Tile: 64 x 8 points in xy, marching through a chunk of z-planes.
#include "internal.h"

namespace lz {

constexpr int kFTileX = 64;
constexpr int kFTileY = 8;
constexpr int kFRows = kFTileY + 2;          // tile rows + two halo rows
constexpr int kFConsumers = kFRows * 32;     // 320 consumer threads
constexpr int kFThreads = kFConsumers;       // no dedicated producer: lane 0 of each warp streams its own row
constexpr int kFRowStride = kFTileX + 4;     // r ring: [pad][left halo][64 values][right halo][pad]
constexpr int kFPitch = 72;                  // staged row pitch in doubles (576 B); idx 0 <-> x0-2
constexpr int kFStages = 4;
constexpr int kFStageDoubles = 3 * kFRows * kFPitch;

struct FusedArgs {
    int nx, ny, nz;
    int periodic;
    int64_t plane;
    double c, ox, oy, oz;
    const double* src[3];  // u = H r_j, r_j, r_{j-1} (null at j = 0)
    const double* diag;
    double* out_r;
    double* out_u;
    const double* s_j;     // device scalars
    const double* alpha_j;
    const double* beta_j;
    const double* s_jm1;
    double* partials;      // [2][gridDim.x]: sum r^2, sum r.u
    int tiles_x, tiles_y, chunks_z, zc;
    int64_t nitems;
};

// ---- cp.async primitives (PTX; LDGSTS in SASS) ------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ void consumer_barrier() {
    asm volatile("bar.sync 1, %0;" :: "n"(kFConsumers) : "memory");
}

template <bool HAS_PREV, bool HAS_DIAG>
__global__ void __launch_bounds__(kFThreads, 2)
fused_step_kernel(const FusedArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* stage_buf = reinterpret_cast<double*>(smem_raw);                              // [S][3][rows][pitch]
    double (*ring)[kFRows][kFRowStride] =
        reinterpret_cast<double (*)[kFRows][kFRowStride]>(stage_buf + kFStages * kFStageDoubles);   // [3][rows][stride]
    double* red = &ring[3][0][0];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NARR = HAS_PREV ? 3 : 2;
    constexpr int ARRS = kFRows * kFPitch;            // doubles between the staged vectors of one stage
    const bool interior = (warp >= 1) && (warp <= kFTileY);
    const double sj = __ldg(a.s_j);
    const double cu = sj;
    const double ca = __ldg(a.alpha_j) * sj;
    const double cb = HAS_PREV ? __ldg(a.beta_j) * __ldg(a.s_jm1) : 0.0;
    const int64_t plane = a.plane;
    const int64_t wrap_back = (int64_t)a.nz * plane;
    double acc_rr = 0.0, acc_ru = 0.0;
    // this thread's slots: staged operands (per stage) and r ring (per slot)
    double* const my_stage = stage_buf + (size_t)warp * kFPitch + 2 + 2 * lane;
    const int hidx = (lane == 0) ? -1 - 2 * lane : kFTileX - 2 * lane;   // halo column relative to my_stage
    int si = 0, sc = 0;                   // stage cursors: next to issue / next to consume (mod S)

    for (int64_t item = blockIdx.x; item < a.nitems; item += gridDim.x) {
        const int tx = (int)(item % a.tiles_x);
        const int64_t t = item / a.tiles_x;
        const int ty = (int)(t % a.tiles_y);
        const int cz = (int)(t / a.tiles_y);
        const int x0 = tx * kFTileX;
        const int iy_raw = ty * kFTileY + warp - 1;                // -1 and +8: halo rows
        const bool row_ok = a.periodic || (iy_raw >= 0 && iy_raw < a.ny);
        int iy = iy_raw;
        if (iy < 0) iy += a.ny;
        if (iy >= a.ny) iy -= a.ny;
        // halo column of this row, fetched by the edge lanes of the interior warps
        int ixh = (lane == 0) ? x0 - 1 : x0 + kFTileX;
        bool hcol = interior && (lane == 0 || lane == 31);
        if (ixh < 0) { if (a.periodic) ixh += a.nx; else hcol = false; }
        if (ixh >= a.nx) { if (a.periodic) ixh -= a.nx; else hcol = false; }
        const int z0 = cz * a.zc;
        const int z1 = min(z0 + a.zc, a.nz);

        // issue cursor: plane zi (un-wrapped) and the element offsets of this thread in that plane
        int zi = z0 - 1;
        int64_t goff = (int64_t)zi * plane + (int64_t)iy * a.nx + x0 + 2 * lane;   // un-wrapped (may be < 0)
        const int64_t hdelta = (int64_t)ixh - (x0 + 2 * lane);     // halo column relative to goff
        // every thread copies exactly the operands it will read itself: no cross-thread hazard,
        // completion by cp.async.wait_group.  One commit group per plane, empty groups included.
        auto issue = [&]() {
            if (zi <= z1) {
                const bool pl_ok = a.periodic || (zi >= 0 && zi < a.nz);
                if (pl_ok && row_ok) {
                    double* d = my_stage + si * kFStageDoubles;
                    int64_t g = goff;                               // periodic wrap of planes -1 and nz
                    if (zi < 0) g += wrap_back;
                    else if (zi >= a.nz) g -= wrap_back;
#pragma unroll
                    for (int arr = 0; arr < NARR; ++arr) {
                        cp_async16(d + arr * ARRS, a.src[arr] + g);
                        if (hcol) cp_async8(d + arr * ARRS + hidx, a.src[arr] + g + hdelta);
                    }
                }
                ++zi;
                goff += plane;
                si = (si + 1 == kFStages) ? 0 : si + 1;
            }
            cp_async_commit();
        };

        double2 rm = make_double2(0.0, 0.0), rc = rm, rp = rm;
        consumer_barrier();                 // r ring reuse across items
        // the stage cursors of issue and consume coincide at an item boundary
        si = sc;
#pragma unroll
        for (int k = 0; k < kFStages - 1; ++k) issue();

        int zp = z0 - 1;                    // plane produced next (un-wrapped)
        int rs = (zp + 3) % 3;              // its r-ring slot
        double* out_r_p = a.out_r + (int64_t)z0 * plane + (int64_t)iy_raw * a.nx + x0 + 2 * lane;   // plane z0, my point
        // r_{j+1} on plane zp -> r ring, registers, HBM
        auto produce = [&]() {
            issue();                        // refill the stage this thread read one call ago
            cp_async_wait<kFStages - 1>();  // plane zp has landed
            const bool pl_ok = a.periodic || (zp >= 0 && zp < a.nz);
            const double* sb = my_stage + sc * kFStageDoubles;
            double2 r = make_double2(0.0, 0.0);
            double rh = 0.0;
            if (pl_ok && row_ok) {
                const double2 uu = *reinterpret_cast<const double2*>(sb);
                const double2 vj = *reinterpret_cast<const double2*>(sb + ARRS);
                r.x = fma(-ca, vj.x, cu * uu.x);
                r.y = fma(-ca, vj.y, cu * uu.y);
                if (HAS_PREV) {
                    const double2 vm = *reinterpret_cast<const double2*>(sb + 2 * ARRS);
                    r.x = fma(-cb, vm.x, r.x);
                    r.y = fma(-cb, vm.y, r.y);
                }
                if (hcol) {
                    rh = fma(-ca, sb[ARRS + hidx], cu * sb[hidx]);
                    if (HAS_PREV) rh = fma(-cb, sb[2 * ARRS + hidx], rh);
                }
                if (interior && zp >= z0 && zp < z1) { st_stream2(out_r_p, r); out_r_p += plane; }
            }
            double* rr = &ring[rs][warp][2 + 2 * lane];
            *reinterpret_cast<double2*>(rr) = r;
            if (lane == 0 || lane == 31) rr[hidx] = rh;
            rm = rc; rc = rp; rp = r;
            ++zp;
            rs = (rs == 2) ? 0 : rs + 1;
            sc = (sc + 1 == kFStages) ? 0 : sc + 1;
        };

        produce();
        produce();
        int cs = z0 % 3;                    // r-ring slot of the plane being consumed
        double* out_u_p = a.out_u + (int64_t)z0 * plane + (int64_t)iy_raw * a.nx + x0 + 2 * lane;
        const double* diag_p = HAS_DIAG ? a.diag + (int64_t)z0 * plane + (int64_t)iy_raw * a.nx + x0 + 2 * lane : nullptr;
        for (int z = z0; z < z1; ++z) {
            produce();
            consumer_barrier();
            if (interior) {
                const double* rr = &ring[cs][warp][2 + 2 * lane];
                const double2 ym = *reinterpret_cast<const double2*>(rr - kFRowStride);
                const double2 yp = *reinterpret_cast<const double2*>(rr + kFRowStride);
                double xl = __shfl_up_sync(0xffffffffu, rc.y, 1);
                double xr = __shfl_down_sync(0xffffffffu, rc.x, 1);
                if (lane == 0) xl = rr[-1];
                if (lane == 31) xr = rr[2];
                double c0 = a.c, c1 = a.c;
                if (HAS_DIAG) {
                    const double2 d = ld_stream2(diag_p);
                    diag_p += plane;
                    c0 += d.x; c1 += d.y;
                }
                double u0 = a.oz * rm.x;
                u0 = fma(a.oy, ym.x, u0);
                u0 = fma(a.ox, xl, u0);
                u0 = fma(c0, rc.x, u0);
                u0 = fma(a.ox, rc.y, u0);
                u0 = fma(a.oy, yp.x, u0);
                u0 = fma(a.oz, rp.x, u0);
                double u1 = a.oz * rm.y;
                u1 = fma(a.oy, ym.y, u1);
                u1 = fma(a.ox, rc.x, u1);
                u1 = fma(c1, rc.y, u1);
                u1 = fma(a.ox, xr, u1);
                u1 = fma(a.oy, yp.y, u1);
                u1 = fma(a.oz, rp.y, u1);
                st_stream2(out_u_p, make_double2(u0, u1));
                out_u_p += plane;
                acc_rr = fma(rc.x, rc.x, acc_rr);
                acc_rr = fma(rc.y, rc.y, acc_rr);
                acc_ru = fma(rc.x, u0, acc_ru);
                acc_ru = fma(rc.y, u1, acc_ru);
            }
            cs = (cs == 2) ? 0 : cs + 1;
        }
        cp_async_wait<0>();
    }
    // deterministic CTA reduction of the two partials (10 warps, fixed order)
    acc_rr = warp_sum(acc_rr);
    acc_ru = warp_sum(acc_ru);
    consumer_barrier();
    if (lane == 0) { red[warp] = acc_rr; red[kFRows + warp] = acc_ru; }
    consumer_barrier();
    if (threadIdx.x == 0) {
        double t0 = 0.0, t1 = 0.0;
        for (int w = 0; w < kFRows; ++w) { t0 += red[w]; t1 += red[kFRows + w]; }
        a.partials[blockIdx.x] = t0;
        a.partials[gridDim.x + blockIdx.x] = t1;
    }
}

constexpr size_t kFSmemBytes = (size_t)kFStages * kFStageDoubles * 8 + (size_t)3 * kFRows * kFRowStride * 8 +
                               kFStages * 8 + 2 * kFRows * 8 + 128;

bool fused_step_supported(const lz_op* op) {
    if (op->kind != LZ_OP_STENCIL) return false;
    const lz_stencil& st = op->st;
    if (st.sharded || st.points != 7) return false;
    if (st.offx == 0.0 || st.offy == 0.0 || st.offz == 0.0) return false;
    if (st.nx % kFTileX != 0 || st.ny % kFTileY != 0) return false;
    if (st.nz < 2) return false;
    if (reinterpret_cast<uintptr_t>(st.diag) & 15) return false;
    return true;
}

// out_r = s_j u - alpha_j s_j rj - beta_j s_jm1 rjm1 ; out_u = H out_r ; partials[0..g) = sum r^2,
// partials[g..2g) = sum r.u ; *nparts = g.
int launch_fused_step(lz_op* op, const double* u, const double* rj, const double* rjm1,
                      const double* s_j, const double* alpha_j, const double* beta_j, const double* s_jm1,
                      double* out_r, double* out_u, double* partials, int* nparts) {
    const lz_stencil& st = op->st;
    lz_ctx* ctx = op->ctx;
    LZ_REQUIRE(((reinterpret_cast<uintptr_t>(u) | reinterpret_cast<uintptr_t>(rj) | reinterpret_cast<uintptr_t>(rjm1) |
                 reinterpret_cast<uintptr_t>(out_r) | reinterpret_cast<uintptr_t>(out_u)) & 15) == 0,
               "fused step: vectors must be 16-byte aligned");
    FusedArgs a;
    a.nx = (int)st.nx; a.ny = (int)st.ny; a.nz = (int)st.nz;
    a.periodic = (st.bc == LZ_BC_PERIODIC);
    a.plane = st.nx * st.ny;
    a.c = st.center; a.ox = st.offx; a.oy = st.offy; a.oz = st.offz;
    a.src[0] = u; a.src[1] = rj; a.src[2] = rjm1; a.diag = st.diag;
    a.out_r = out_r; a.out_u = out_u;
    a.s_j = s_j; a.alpha_j = alpha_j; a.beta_j = beta_j; a.s_jm1 = s_jm1;
    a.partials = partials;
    a.tiles_x = (int)(st.nx / kFTileX);
    a.tiles_y = (int)(st.ny / kFTileY);
    const void* fn;
    if (rjm1) fn = st.diag ? (const void*)fused_step_kernel<true, true> : (const void*)fused_step_kernel<true, false>;
    else fn = st.diag ? (const void*)fused_step_kernel<false, true> : (const void*)fused_step_kernel<false, false>;
    if (op->fused_per_sm == 0) {
        const void* all[4] = {(const void*)fused_step_kernel<true, true>, (const void*)fused_step_kernel<true, false>,
                              (const void*)fused_step_kernel<false, true>, (const void*)fused_step_kernel<false, false>};
        for (const void* f : all)
            LZ_CUDA(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFSmemBytes));
        int per_sm = 0;
        LZ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)fused_step_kernel<true, true>,
                                                              kFThreads, kFSmemBytes));
        op->fused_per_sm = per_sm < 1 ? 1 : per_sm;
    }
    const int64_t gmax = std::min<int64_t>((int64_t)ctx->sms * op->fused_per_sm, kMaxPartials / 2);
    const int64_t tiles = (int64_t)a.tiles_x * a.tiles_y;
    // z-chunking: halo planes cost 2/zc extra reads of three vectors; balance against an even fill
    int best_chunks = 1;
    double best_cost = 1e300;
    for (int ch = 1; ch <= (int)st.nz; ++ch) {
        const int zc = (int)((st.nz + ch - 1) / ch);
        const int chunks = (int)((st.nz + zc - 1) / zc);
        const int64_t items = tiles * chunks;
        const int64_t g = std::min<int64_t>(items, gmax);
        const int64_t rounds = (items + g - 1) / g;
        const double cost = (double)rounds * (zc + 2.0 * 0.6) / ((double)st.nz * tiles / gmax);
        if (cost < best_cost - 1e-12) { best_cost = cost; best_chunks = chunks; }
        if (zc <= 8) break;
    }
    a.zc = (int)((st.nz + best_chunks - 1) / best_chunks);
    a.chunks_z = (int)((st.nz + a.zc - 1) / a.zc);
    a.nitems = tiles * a.chunks_z;
    const int grid = (int)std::min<int64_t>(a.nitems, gmax);
    void* args[] = {(void*)&a};
    LZ_CUDA(cudaLaunchKernel(fn, dim3(grid), dim3(kFThreads), args, kFSmemBytes, ctx->stream));
    if (nparts) *nparts = grid;
    return LZ_OK;
}

}  // namespace lz
