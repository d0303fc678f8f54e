// umma_tput.cu -- tcgen05.mma kind::i8 dispatch throughput for small tiles: does the cost per instruction come from the
// issuing thread or from the (shared) tensor pipe front end?  One CTA, W issuing warps (one elected lane each), each issues
// R back-to-back MMAs (same descriptors, accumulate) into its own TMEM columns, then commits; we time until all complete.
#include <cstdio>
#include <cstdlib>
#include "../bitnetmcu_b200/csrc/sm100_ptx.cuh"
using namespace bnm;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__global__ void __launch_bounds__(512, 1) tput_kernel(long long *out, int W, int R, int N, int ts) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[8];
    __shared__ uint32_t tmem_base_s;
    const uint32_t tid = threadIdx.x;
    const uint32_t warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    if (tid == 0) { for (int i = 0; i < 8; i++) mbar_init(&bar[i], 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc<512>(&tmem_base_s);
    for (int i = tid; i < 98304 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x01020304u;
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tmem_base_s;
    const uint32_t idesc = make_idesc_i8(128, N);
    long long t0 = clock64(), t1 = 0, t2 = 0;
    if ((warp & 3) == 0 && (warp >> 2) < (uint32_t)W) {
        const uint32_t g = warp >> 2;
        const uint64_t bd = make_smem_desc(smem_u32(smem) + 32768, 128, 256, UMMA_LAYOUT_NONE);
        const uint64_t ad = make_smem_desc(smem_u32(smem), 0, 1024, UMMA_LAYOUT_SW128);
        const uint32_t d = tb + g * 64, a = tb + 448;
        const bool leader = elect_one();
        for (int r = 0; r < R; r++) {
            if (leader) { if (ts) umma_i8_ts(d, a, bd, idesc, 1); else umma_i8_ss(d, ad, bd, idesc, 1); }
        }
        if (leader) umma_commit(&bar[g]);
        __syncwarp();
        t1 = clock64();
        mbar_wait(&bar[g], 0);
        t2 = clock64();
        if ((tid & 31) == 0) { out[g * 2] = t1 - t0; out[g * 2 + 1] = t2 - t0; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tb);
}

int main() {
    CK(cudaSetDevice(0));
    CK(cudaFuncSetAttribute(tput_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304 + 1024));
    long long *d, h[16];
    CK(cudaMalloc(&d, 128));
    for (int ts = 0; ts < 2; ts++)
        for (int N : {16, 64, 256})
            for (int W : {1, 2, 4})
                for (int R : {1, 8, 64}) {
                    if (N == 256 && W > 1) continue;
                    tput_kernel<<<1, 512, 98304 + 1024>>>(d, W, R, N, ts);
                    CK(cudaDeviceSynchronize());
                    CK(cudaMemcpy(h, d, 128, cudaMemcpyDeviceToHost));
                    long long issue = 0, done = 0;
                    for (int g = 0; g < W; g++) { issue = issue > h[2 * g] ? issue : h[2 * g]; done = done > h[2 * g + 1] ? done : h[2 * g + 1]; }
                    printf("%s N=%3d issuers=%d MMAs/issuer=%2d : issue %6lld cyc, all done %6lld cyc  -> %6.1f cyc per MMA (chip-wide per SM)\n",
                           ts ? "TS" : "SS", N, W, R, issue, done, (double)done / (W * R));
                }
    return 0;
}
