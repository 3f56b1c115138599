/* Minimal C99 client of the drop-in boundary (include/lanczos_b200.h): proves that the header is plain C
 * and that the shared library links and runs from C.  With a B200 present it runs a 16^3 Laplacian through
 * lz_lanczos_run; without one it stops after the device query (the library has no CPU path).
 *
 *   gcc -std=c99 -Wall -Werror -Iinclude examples/c_abi_demo.c -Llanczos_b200 -llanczos_b200 \
 *       -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/lanczos_b200 -o /tmp/c_abi_demo && /tmp/c_abi_demo
 */
#include <stdio.h>
#include <stdlib.h>
#include "lanczos_b200.h"

/* device memory through the CUDA runtime, declared by hand to keep this file free of CUDA headers */
extern int cudaMalloc(void** p, size_t bytes);
extern int cudaMemcpy(void* dst, const void* src, size_t bytes, int kind);
extern int cudaFree(void* p);

int main(void) {
    int ndev = -1;
    printf("lz_abi_version = %d\n", lz_abi_version());
    if (lz_device_count(&ndev) != LZ_OK) { printf("lz_device_count failed: %s\n", lz_last_error()); return 1; }
    printf("devices = %d\n", ndev);
    if (lz_op_rows(NULL, NULL) != LZ_ERR_INVALID) { printf("argument validation is broken\n"); return 1; }
    printf("error string: %s\n", lz_last_error());
    if (ndev < 1) { printf("no CUDA device: stopping after the boundary checks\n"); return 0; }

    lz_ctx* ctx = NULL;
    lz_op* op = NULL;
    if (lz_ctx_create(0, NULL, &ctx) != LZ_OK) { printf("lz_ctx_create: %s\n", lz_last_error()); return 1; }
    const int64_t shape[3] = {16, 16, 16};
    const double off[3] = {-1.0, -1.0, -1.0};
    if (lz_op_stencil_create(ctx, 3, shape, LZ_BC_PERIODIC, 6.0, off, NULL, &op) != LZ_OK) {
        printf("lz_op_stencil_create: %s\n", lz_last_error());
        return 1;
    }
    enum { M = 16 * 16 * 16, N = 12, LD = 4096 };
    double* v0_host = (double*)malloc(sizeof(double) * M);
    for (int i = 0; i < M; ++i) v0_host[i] = (double)((i * 2654435761u) % 1000u) / 500.0 - 1.0;
    void *v0 = NULL, *V = NULL;
    if (cudaMalloc(&v0, sizeof(double) * M) || cudaMalloc(&V, sizeof(double) * N * LD)) { printf("cudaMalloc failed\n"); return 1; }
    cudaMemcpy(v0, v0_host, sizeof(double) * M, 1 /* cudaMemcpyHostToDevice */);
    lz_run_opts opts = {LZ_REORTH_FULL, 1, 1, 0, 0, 0, 0.0, 0.0};
    lz_run_info info;
    double alpha[N], beta[N - 1], scale[N];
    if (lz_lanczos_run(ctx, op, (const double*)v0, N, &opts, alpha, beta, (double*)V, LD, scale, &info) != LZ_OK) {
        printf("lz_lanczos_run: %s\n", lz_last_error());
        return 1;
    }
    printf("steps_done = %d, launches = %d, step kernel = %d\n", info.steps_done, info.launches, info.step_kernel);
    for (int j = 0; j < 3; ++j) printf("alpha[%d] = %.15f  beta[%d] = %.15f\n", j, alpha[j], j, beta[j]);
    cudaFree(v0);
    cudaFree(V);
    free(v0_host);
    lz_op_destroy(op);
    lz_ctx_destroy(ctx);
    return 0;
}
