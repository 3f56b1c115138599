// Cross-GPU exchange over NVLink peer memory (cudaIpc-mapped buffers), used by the row-sharded
// Lanczos loop: a push-based all-to-all of a few doubles ("scalar allreduce" of alpha, beta and
// the Gram-Schmidt coefficients) whose flag also publishes the halo planes / ghost values that
// the producing kernel stored straight into the neighbour's memory.
//
// Every rank owns one exchange buffer with the same layout:
//   flags [nslots][world]        u64   sequence number of the last payload from rank q
//   slots [nslots][world][kmax]  f64   payload of rank q
//   ghost_lo[2][plane], ghost_hi[2][plane]  halo planes (double-buffered by row parity)
//   gather[2][nghost]                     ghost entries of a sparse operator
// A reduction with sequence number `seq` uses slot seq % nslots.  Rank r writes its payload
// into slots[slot][r] of EVERY rank (itself included), then stores flags[slot][r] = seq with
// release semantics at system scope; a consumer spins (acquire, system scope) until all world
// flags of its own buffer equal seq and adds the payloads in rank order - so the result is
// bit-identical on every rank and independent of arrival order.  A rank can run at most one
// EXECUTED reduction ahead of a peer (it needs the peer's next payload to go further).  Sequence
// numbers are handed out on the host for every exchange that is enqueued, including the ones whose
// kernels are predicated off on the device (selective re-orthogonalisation: the Gram-Schmidt
// exchanges of a step that does not fire return without publishing).  Between two executed
// exchanges at most kMaxSkippedRun numbers are skipped, so the payload a fast rank writes while a
// slow peer still reads exchange `a` carries a number in (a, a + kMaxSkippedRun + 1]: the ring must
// be longer than that span for the two never to share a slot.
#pragma once
#include "common.cuh"

namespace lz {

constexpr int kMaxWorld = 16;
constexpr int kMaxSkippedRun = 6;   // lanczos.cu: per step <= 2 fin_ip + peer_sync + the alpha of the re-swept row
                                    // (all predicated on the same flag), rounded up
constexpr int kPeerSlots = 8;
static_assert(kPeerSlots >= kMaxSkippedRun + 2, "a skipped run of exchanges must not wrap the slot ring");

struct PeerComm {
    int world = 1;
    int rank = 0;
    int kmax = 0;                                 // doubles per payload
    int* err = nullptr;                           // local device int, set to 1 on a spin timeout
    unsigned long long* flags[kMaxWorld] = {};    // flags region of rank q's buffer, as mapped here
    double* slots[kMaxWorld] = {};                // slots region of rank q's buffer, as mapped here
};

// byte offsets inside an exchange buffer
struct CommLayout {
    size_t flags_off, slots_off, ghost_lo_off, ghost_hi_off, gather_off, total;
    static CommLayout make(int world, int kmax, int64_t plane, int64_t nghost) {
        auto up = [](size_t v) { return (v + 511) & ~(size_t)511; };
        CommLayout L;
        L.flags_off = 0;
        L.slots_off = up((size_t)kPeerSlots * world * 8);
        L.ghost_lo_off = L.slots_off + up((size_t)kPeerSlots * world * kmax * 8);
        L.ghost_hi_off = L.ghost_lo_off + up((size_t)2 * plane * 8);
        L.gather_off = L.ghost_hi_off + up((size_t)2 * plane * 8);
        L.total = L.gather_off + up((size_t)2 * nghost * 8) + 512;
        return L;
    }
};

#ifdef __CUDACC__

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_volatile_f64(const double* p) {
    double v;
    asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// Whole CTA: after every thread has stored its share of the payload, publish it.
__device__ __forceinline__ void peer_publish(const PeerComm& pc, unsigned long long seq) {
    __threadfence_system();
    __syncthreads();
    const int slot = (int)(seq % kPeerSlots);
    if (threadIdx.x < pc.world)
        st_release_sys(pc.flags[threadIdx.x] + (size_t)slot * pc.world + pc.rank, seq);
}

// payload element i of this rank for reduction seq -> every rank's buffer
__device__ __forceinline__ void peer_store(const PeerComm& pc, unsigned long long seq, int i, double v) {
    const int slot = (int)(seq % kPeerSlots);
    const size_t at = ((size_t)slot * pc.world + pc.rank) * pc.kmax + i;
    for (int q = 0; q < pc.world; ++q) pc.slots[q][at] = v;
}

// Whole CTA: wait until the payloads of all ranks for `seq` have landed in my buffer.
__device__ __forceinline__ void peer_wait(const PeerComm& pc, unsigned long long seq) {
    const int slot = (int)(seq % kPeerSlots);
    // after one timeout the run is lost anyway: do not wait again (a dead peer must not turn
    // into minutes of spinning)
    if (threadIdx.x < pc.world && *reinterpret_cast<volatile int*>(pc.err) == 0) {
        const unsigned long long* f = pc.flags[pc.rank] + (size_t)slot * pc.world + threadIdx.x;
        const long long t0 = clock64();
        while (ld_acquire_sys(f) != seq) {
            if (clock64() - t0 > 40000000000LL) {     // ~20 s at 2 GHz: a peer died
                *pc.err = 1;
                break;
            }
            __nanosleep(64);
        }
    }
    __syncthreads();
}

// sum over ranks, in rank order, of payload element i
__device__ __forceinline__ double peer_sum(const PeerComm& pc, unsigned long long seq, int i) {
    const int slot = (int)(seq % kPeerSlots);
    const double* base = pc.slots[pc.rank] + (size_t)slot * pc.world * pc.kmax + i;
    double t = 0.0;
    for (int q = 0; q < pc.world; ++q) t += ld_volatile_f64(base + (size_t)q * pc.kmax);
    return t;
}

#endif  // __CUDACC__

// exchange modes of the single-CTA "fin" kernels
enum { LZ_XCHG_FUSED = 0,      // push my payload, wait for everybody, combine (one process per GPU)
       LZ_XCHG_PUSH = 1,       // push only      } single-process emulation of several shards on one
       LZ_XCHG_COMBINE = 2 };  // wait + combine } GPU: all pushes are enqueued before any combine

}  // namespace lz
