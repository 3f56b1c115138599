// Shared device/host helpers for the lanczos_b200 kernels (sm_100a, fp64 on CUDA cores).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <utility>
#include "../../include/lanczos_b200.h"

namespace lz {

// ------------------------------------------------------------------ error plumbing
void set_error(const char* fmt, ...);
const char* get_error();

#define LZ_CUDA(call)                                                                 \
    do {                                                                              \
        cudaError_t e__ = (call);                                                     \
        if (e__ != cudaSuccess) {                                                     \
            lz::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),    \
                          __FILE__, __LINE__);                                        \
            return (e__ == cudaErrorMemoryAllocation) ? LZ_ERR_NOMEM : LZ_ERR_CUDA;   \
        }                                                                             \
    } while (0)

#define LZ_CHECK(expr)                                                                \
    do {                                                                              \
        int s__ = (expr);                                                             \
        if (s__ != LZ_OK) return s__;                                                 \
    } while (0)

#define LZ_REQUIRE(cond, ...)                                                         \
    do {                                                                              \
        if (!(cond)) {                                                                \
            lz::set_error(__VA_ARGS__);                                               \
            return LZ_ERR_INVALID;                                                    \
        }                                                                             \
    } while (0)

// ------------------------------------------------------------------ launch geometry
constexpr int kThreads = 256;          // every streaming kernel uses 256-thread CTAs
constexpr int kWarps = kThreads / 32;
constexpr int kMaxPartials = 4096;     // upper bound on CTAs that write a partial sum

// ------------------------------------------------------------------ launches
// Programmatic dependent launch: every kernel of the loop starts with pdl_wait() (blocks until the
// preceding kernel of the stream has completed and its writes are visible) right after
// pdl_trigger() (lets the NEXT kernel of the stream be scheduled as soon as this one's CTAs leave
// the SMs).  Launched with the programmatic-stream-serialization attribute the launch latency and
// the CTA ramp of kernel k+1 overlap the tail of kernel k instead of following its drain; without
// the attribute both instructions are no-ops.  The attribute is OFF unless LZ_PDL=1 is set: see pdl_enabled()
// in capi.cu for the measured gain (0.7 %) and the hazard with non-coherent loads that keeps it opt-in.
bool pdl_enabled();
#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                            Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
// untyped form (kernel chosen at run time from a table of template instances)
inline cudaError_t launch_fn(const void* fn, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, void** args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelExC(&cfg, fn, args);
}
#endif

// ------------------------------------------------------------------ device helpers
#ifdef __CUDACC__

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// first statement of every loop kernel
__device__ __forceinline__ void pdl_prologue() { pdl_trigger(); pdl_wait(); }

// Streaming 128-bit fp64 loads/stores.  The Krylov vectors are far larger than L2
// (1.07 GB each at 512^3), so everything that is touched once per kernel bypasses L1
// allocation; neighbour-reused loads (stencil rows) use the default policy.
__device__ __forceinline__ double2 ld_stream2(const double* p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];"
                 : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ double ld_stream1(const double* p) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
// Same streaming policy but through the coherent path: for operands that the same kernel
// also writes (in-place updates), where ld.global.nc would be undefined.
__device__ __forceinline__ double2 ld_stream2_rw(const double* p) {
    double2 v;
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];"
                 : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double2 ld_cached2(const double* p) {
    return __ldg(reinterpret_cast<const double2*>(p));
}
__device__ __forceinline__ void st_stream2(double* p, double2 v) {
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};"
                 :: "l"(p), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void st_stream1(double* p, double v) {
    asm volatile("st.global.L1::no_allocate.f64 [%0], %1;" :: "l"(p), "d"(v) : "memory");
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic CTA reduction: fixed shuffle tree inside each warp, then warp 0 adds the
// per-warp sums in warp order.  Result valid in thread 0.  `smem` holds >= kWarps doubles.
__device__ __forceinline__ double block_sum(double v, double* smem) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();                       // protect smem reuse across calls
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 0; w < kWarps; ++w) t += smem[w];
    }
    return t;
}

#endif  // __CUDACC__

}  // namespace lz
