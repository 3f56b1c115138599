// Diagonal potential of a structured-grid Hamiltonian, evaluated on the device.
//
// replaces Hamiltonian.create_sparse_V (Python/Regular/Hamiltonian.py:35-46): three nested Python
// loops that call potential(x[i], y[j], z[k]) N^3 times and build a diagonal CSR matrix.  Here the
// caller's potential - traced once on the host into a short postfix program (lanczos_b200/hamiltonian.py:
// the unchanged NumPy function is called with symbolic arguments) - is interpreted per grid point by
// one kernel that writes the diagonal straight into HBM, index map i + N*(j + N*k) (Hamiltonian.py:42,73-76).
// 8 B written per point, nothing read but the 3 coordinate axes (L1-resident).
#include "internal.h"

namespace lz {

constexpr int kPotMaxOps = 96;
constexpr int kPotMaxConsts = 32;
constexpr int kPotStack = 12;

struct PotProgram {
    int nops;
    int ops[kPotMaxOps];           // code | (const index << 8)
    double consts[kPotMaxConsts];
};

__global__ void __launch_bounds__(kThreads)
potential_eval_kernel(const PotProgram p, int nx, int ny, int nz, const double* __restrict__ xs,
                      const double* __restrict__ ys, const double* __restrict__ zs, double* __restrict__ out) {
    const int64_t total = (int64_t)nx * ny * nz;
    const int64_t nthr = (int64_t)gridDim.x * kThreads;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += nthr) {
        const int ix = (int)(i % nx);
        const int64_t t = i / nx;
        const int iy = (int)(t % ny);
        const int iz = (int)(t / ny);
        const double x = __ldg(xs + ix), y = __ldg(ys + iy), z = __ldg(zs + iz);
        double st[kPotStack];
        int sp = 0;
        for (int k = 0; k < p.nops; ++k) {
            const int code = p.ops[k] & 0xff;
            switch (code) {
                case LZ_POT_X: st[sp++] = x; break;
                case LZ_POT_Y: st[sp++] = y; break;
                case LZ_POT_Z: st[sp++] = z; break;
                case LZ_POT_CONST: st[sp++] = p.consts[p.ops[k] >> 8]; break;
                // two roundings, never contracted into an fma: NumPy evaluates a*b + c as two ufunc calls
                case LZ_POT_ADD: --sp; st[sp - 1] = __dadd_rn(st[sp - 1], st[sp]); break;
                case LZ_POT_SUB: --sp; st[sp - 1] = __dsub_rn(st[sp - 1], st[sp]); break;
                case LZ_POT_MUL: --sp; st[sp - 1] = __dmul_rn(st[sp - 1], st[sp]); break;
                case LZ_POT_DIV: --sp; st[sp - 1] = __ddiv_rn(st[sp - 1], st[sp]); break;
                case LZ_POT_POW: --sp; st[sp - 1] = pow(st[sp - 1], st[sp]); break;
                case LZ_POT_MIN: --sp; st[sp - 1] = fmin(st[sp - 1], st[sp]); break;
                case LZ_POT_MAX: --sp; st[sp - 1] = fmax(st[sp - 1], st[sp]); break;
                case LZ_POT_NEG: st[sp - 1] = -st[sp - 1]; break;
                case LZ_POT_SQRT: st[sp - 1] = __dsqrt_rn(st[sp - 1]); break;
                case LZ_POT_EXP: st[sp - 1] = exp(st[sp - 1]); break;
                case LZ_POT_LOG: st[sp - 1] = log(st[sp - 1]); break;
                case LZ_POT_ABS: st[sp - 1] = fabs(st[sp - 1]); break;
                case LZ_POT_SIN: st[sp - 1] = sin(st[sp - 1]); break;
                case LZ_POT_COS: st[sp - 1] = cos(st[sp - 1]); break;
                case LZ_POT_TANH: st[sp - 1] = tanh(st[sp - 1]); break;
                case LZ_POT_SQUARE: st[sp - 1] = __dmul_rn(st[sp - 1], st[sp - 1]); break;
                default: break;
            }
        }
        out[i] = st[0];
    }
}

}  // namespace lz

using namespace lz;

extern "C" int lz_potential_eval(lz_ctx* ctx, const int64_t* shape, const double* x_host, const double* y_host,
                                 const double* z_host, int32_t nops, const int32_t* ops_host, int32_t nconsts,
                                 const double* consts_host, double* out_dev) {
    LZ_REQUIRE(ctx && shape && x_host && y_host && z_host && ops_host && out_dev, "lz_potential_eval: null argument");
    LZ_REQUIRE(nops >= 1 && nops <= kPotMaxOps, "lz_potential_eval: program of %d ops (1..%d supported)", nops, kPotMaxOps);
    LZ_REQUIRE(nconsts >= 0 && nconsts <= kPotMaxConsts && (nconsts == 0 || consts_host),
               "lz_potential_eval: %d constants (at most %d)", nconsts, kPotMaxConsts);
    const int64_t nx = shape[0], ny = shape[1], nz = shape[2];
    LZ_REQUIRE(nx >= 1 && ny >= 1 && nz >= 1 && nx < (1 << 30) && ny < (1 << 30) && nz < (1 << 30), "lz_potential_eval: bad shape");
    PotProgram p{};
    p.nops = nops;
    // validate on the host: operand counts, stack depth, constant indices - the kernel trusts the program
    int depth = 0;
    for (int k = 0; k < nops; ++k) {
        const int code = ops_host[k] & 0xff;
        int pop = 0, push = 1;
        if (code <= LZ_POT_CONST) { pop = 0; }
        else if (code >= LZ_POT_ADD && code <= LZ_POT_MAX) { pop = 2; }
        else if (code >= LZ_POT_NEG && code <= LZ_POT_SQUARE) { pop = 1; }
        else { set_error("lz_potential_eval: unknown op code %d at %d", code, k); return LZ_ERR_INVALID; }
        if (code == LZ_POT_CONST) LZ_REQUIRE((ops_host[k] >> 8) >= 0 && (ops_host[k] >> 8) < nconsts, "lz_potential_eval: constant index out of range at op %d", k);
        LZ_REQUIRE(depth >= pop, "lz_potential_eval: stack underflow at op %d", k);
        depth += push - pop;
        LZ_REQUIRE(depth <= kPotStack, "lz_potential_eval: expression needs a stack deeper than %d", kPotStack);
        p.ops[k] = ops_host[k];
    }
    LZ_REQUIRE(depth == 1, "lz_potential_eval: the program leaves %d values on the stack", depth);
    for (int k = 0; k < nconsts; ++k) p.consts[k] = consts_host[k];
    LZ_CUDA(cudaSetDevice(ctx->device));
    double* axes = nullptr;
    LZ_CUDA(cudaMalloc((void**)&axes, (size_t)(nx + ny + nz) * 8));
    cudaError_t e = cudaMemcpyAsync(axes, x_host, (size_t)nx * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(axes + nx, y_host, (size_t)ny * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(axes + nx + ny, z_host, (size_t)nz * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        const int64_t total = nx * ny * nz;
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((total + kThreads - 1) / kThreads, (int64_t)ctx->sms * 16));
        potential_eval_kernel<<<grid, kThreads, 0, ctx->stream>>>(p, (int)nx, (int)ny, (int)nz, axes, axes + nx, axes + nx + ny, out_dev);
        e = cudaGetLastError();
    }
    cudaStreamSynchronize(ctx->stream);
    cudaFree(axes);
    if (e != cudaSuccess) { set_error("lz_potential_eval: %s", cudaGetErrorString(e)); return LZ_ERR_CUDA; }
    return LZ_OK;
}
