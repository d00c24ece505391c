// TEST INFRASTRUCTURE: proves that the emulator's strictness checks fire.  Each case is a tiny "CUDA" program with one deliberate
// host-side mistake that a GPU punishes and a naive CPU emulation would forgive; run as  strictness <case>.
#include "cuda_runtime.h"
#include <cstdio>
#include <cstring>

__global__ void fill(int *p, int n, int v) {
    unsigned char *dyn = emu::dyn_smem(); (void)dyn;     // (what build_emu.py rewrites `extern __shared__` into)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = v;
}
#define GO(kern, g, b, sm, ...) emu::launch_kernel((const void *)(kern), #kern, g, b, sm, [&] { kern(__VA_ARGS__); })

int main(int argc, char **argv) {
    const char *c = argc > 1 ? argv[1] : "";
    int *d = nullptr, h[64];
    cudaMalloc(&d, sizeof(h));
    cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    if (!strcmp(c, "ok")) {
        cudaFuncSetAttribute(fill, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        GO(fill, 2, 32, 100 * 1024, d, 64, 7);
        cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, s);
        cudaStreamSynchronize(s);
        return h[63] == 7 ? 0 : 1;
    }
    if (!strcmp(c, "empty_grid")) { GO(fill, 0, 32, 0, d, 64, 7); return 0; }
    if (!strcmp(c, "smem_no_optin")) { GO(fill, 1, 32, 100 * 1024, d, 64, 7); return 0; }
    if (!strcmp(c, "smem_optin_too_small")) {
        cudaFuncSetAttribute(fill, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        GO(fill, 1, 32, 100 * 1024, d, 64, 7); return 0;
    }
    if (!strcmp(c, "read_before_sync")) {       // exit code 3 = the host saw poison, as it may on a GPU
        GO(fill, 1, 32, 0, d, 64, 7);
        cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, s);
        const bool early = h[63] != 7;
        cudaStreamSynchronize(s);
        return early && h[63] == 7 ? 3 : 1;
    }
    if (!strcmp(c, "event_before_sync")) {      // exit code 3 = cudaErrorNotReady
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a, s); cudaEventRecord(b, s);
        float ms;
        const bool early = cudaEventElapsedTime(&ms, a, b) == cudaErrorNotReady;
        cudaStreamSynchronize(s);
        return early && cudaEventElapsedTime(&ms, a, b) == cudaSuccess ? 3 : 1;
    }
    if (!strcmp(c, "uninitialised")) {          // exit code 3 = fresh device memory is not zero
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        return h[0] != 0 ? 3 : 1;
    }
    fprintf(stderr, "unknown case %s\n", c);
    return 2;
}
