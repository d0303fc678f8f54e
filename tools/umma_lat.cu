// umma_lat.cu -- latency microbenchmarks for the fused FC kernel's building blocks (single CTA, 128 threads).
// Prints cycles (clock64, SM clock) for: MMA batches issue->mbarrier completion, tcgen05.ld/st round trips.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../bitnetmcu_b200/csrc/sm100_ptx.cuh"
using namespace bnm;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__global__ void __launch_bounds__(128, 1) lat_kernel(long long *out, int n_mma, int N, int ts, int reps) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc<512>(&tmem_base_s);
    for (int i = tid; i < 98304 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x01010101u * (i & 3);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tmem_base_s;
    uint32_t ph = 0;
    long long t_issue = 0, t_done = 0, t_ld = 0, t_st = 0;
    const uint32_t idesc = make_idesc_i8(128, N);
    for (int r = 0; r < reps; r++) {
        long long t0 = clock64();
        if (tid == 0) {
            for (int s = 0; s < n_mma; s++) {
                uint64_t bd = make_smem_desc(smem_u32(smem) + 32768 + (s & 7) * (N * 32), 128, 256, UMMA_LAYOUT_NONE);
                if (ts) umma_i8_ts(tb, tb + 256 + (s & 7) * 8, bd, idesc, s > 0);
                else {
                    uint64_t ad = make_smem_desc(smem_u32(smem) + ((s & 7) >> 2) * 16384 + (s & 3) * 32, 0, 1024, UMMA_LAYOUT_SW128);
                    umma_i8_ss(tb, ad, bd, idesc, s > 0);
                }
            }
            umma_commit(&bar);
        }
        long long t1 = clock64();
        __syncwarp();
        mbar_wait(&bar, ph);
        ph ^= 1;
        tc_fence_after();
        long long t2 = clock64();
        uint32_t v[4][16];
        for (int c = 0; c < 4; c++) tmem_ld_x16(tb + ((uint32_t)(warp * 32) << 16) + c * 16, v[c]);
        tmem_ld_wait();
        long long t3 = clock64();
        uint32_t w[4];
        for (int c = 0; c < 4; c++) { w[0] = v[c][0]; w[1] = v[c][5]; w[2] = v[c][9]; w[3] = v[c][13]; tmem_st_x4(tb + 256 + ((uint32_t)(warp * 32) << 16) + c * 4, w); }
        tmem_st_wait();
        long long t4 = clock64();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (r > 0) { t_issue += t1 - t0; t_done += t2 - t0; t_ld += t3 - t2; t_st += t4 - t3; }
    }
    if (tid == 0) { out[0] = t_issue / (reps - 1); out[1] = t_done / (reps - 1); out[2] = t_ld / (reps - 1); out[3] = t_st / (reps - 1); }
    if (tid == 64) { out[4] = t_done / (reps - 1); }
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tb);
}

int main() {
    CK(cudaSetDevice(0));
    CK(cudaFuncSetAttribute(lat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304 + 1024));
    long long *d, h[8];
    CK(cudaMalloc(&d, 64));
    struct C { int n_mma, N, ts; } cases[] = {{1, 64, 0}, {2, 64, 0}, {8, 64, 0}, {16, 64, 0}, {1, 64, 1}, {2, 64, 1}, {8, 64, 1}, {2, 16, 1}, {8, 256, 0}, {2, 256, 1}, {5, 160, 1}, {8, 160, 0}};
    for (auto c : cases) {
        lat_kernel<<<1, 128, 98304 + 1024>>>(d, c.n_mma, c.N, c.ts, 50);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost));
        printf("n_mma=%2d N=%3d %s: issue %5lld cyc, issue->done(thread0) %5lld, (thread64) %5lld | tmem ld 64col %4lld, st 16col %4lld\n",
               c.n_mma, c.N, c.ts ? "TS" : "SS", h[0], h[1], h[4], h[2], h[3]);
    }
    return 0;
}
