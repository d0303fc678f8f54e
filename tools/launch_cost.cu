// launch_cost.cu -- fixed per-launch costs of a persistent 148-CTA kernel on B200: what does an (almost) empty kernel cost
// with the fused kernel's launch shape (480 threads, ~217 kB dynamic smem, 512 TMEM columns, 25 kB weight staging)?
#include <cstdio>
#include <cstdlib>
#include "../bitnetmcu_b200/csrc/sm100_ptx.cuh"
using namespace bnm;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__global__ void __launch_bounds__(480, 1) k(int mode, const uint4 *w, uint32_t w16, int *sink) {
    extern __shared__ uint8_t smem[];
    __shared__ uint32_t tb;
    const uint32_t warp = threadIdx.x >> 5;
    if (mode >= 2 && warp == 1) tmem_alloc<512>(&tb);
    if (mode >= 3) {
        uint4 *dst = reinterpret_cast<uint4 *>(smem);
        for (uint32_t i = threadIdx.x; i < w16; i += blockDim.x) dst[i] = __ldg(w + i);
    }
    __syncthreads();
    if (mode >= 2) { tc_fence_before(); __syncthreads(); if (warp == 1) tmem_dealloc<512>(tb); }
    if (mode >= 3 && smem[threadIdx.x] == 123 && sink) *sink = 1;
}

int main() {
    CK(cudaSetDevice(0));
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 222208));
    uint4 *w; int *sink;
    CK(cudaMalloc(&w, 65536)); CK(cudaMemset(w, 1, 65536)); CK(cudaMalloc(&sink, 4));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    struct C { const char *name; int mode; size_t smem; int threads; } cases[] = {
        {"empty, 32 threads, no smem, 148 CTAs", 0, 0, 32}, {"empty, 480 threads, no smem", 0, 0, 480}, {"empty, 480 threads, 217 kB smem", 1, 222208, 480},
        {"+ tmem alloc/dealloc 512 cols", 2, 222208, 480}, {"+ 25.6 kB weight staging", 3, 222208, 480}};
    for (auto c : cases) {
        for (int i = 0; i < 5; i++) k<<<148, c.threads, c.smem>>>(c.mode, w, 1600, sink);
        CK(cudaDeviceSynchronize());
        const int N = 200;
        cudaEventRecord(a);
        for (int i = 0; i < N; i++) k<<<148, c.threads, c.smem>>>(c.mode, w, 1600, sink);
        cudaEventRecord(b);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("%-45s %7.2f us per launch (back-to-back, %d launches)\n", c.name, ms * 1e3 / N, N);
    }
    return 0;
}
