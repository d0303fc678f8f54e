// tmem_ld_bench.cu -- tcgen05.ld / tcgen05.st throughput: how many bytes per cycle does TMEM deliver to the register file, per warp,
// per lane quarter (SM sub-partition) and per SM, for the x16 / x32 / x64 shapes the kernels use?  One CTA; W warps, warp w reads
// lane quarter w % 4; R loads each, back to back with one wait at the end of every batch of 4 (enough independent loads in
// flight) and a dependent variant (wait after every load).
#include <cstdio>
#include <cstdlib>
#include "../bitnetmcu_b200/csrc/sm100_ptx.cuh"
using namespace bnm;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

template <int X, bool DEP, bool STORE>
__global__ void __launch_bounds__(512, 1) k(long long *out, int W, int R) {
    __shared__ uint32_t tmem_base_s;
    const uint32_t tid = threadIdx.x;
    const uint32_t warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    if (warp == 0) tmem_alloc<512>(&tmem_base_s);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tmem_base_s + (((warp & 3) * 32) << 16) + (warp >> 2) * 64;
    uint32_t acc = 0;
    long long t0 = clock64();
    if (warp < (uint32_t)W) {
        for (int r = 0; r < R; r += 4) {
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (X == 16) {
                    uint32_t v[16];
                    if (STORE) { for (int i = 0; i < 16; i++) v[i] = acc + i; tmem_st_x16(tb + (u & 3) * 16, v); if (DEP) tmem_st_wait(); }
                    else { tmem_ld_x16(tb + (u & 3) * 16, v); if (DEP) tmem_ld_wait(); acc += v[0] + v[15]; }
                } else if (X == 32) {
                    uint32_t v[32];
                    tmem_ld_x32(tb + (u & 1) * 32, v); if (DEP) tmem_ld_wait(); acc += v[0] + v[31];
                } else {
                    uint32_t v[64];
                    tmem_ld_x64(tb, v); if (DEP) tmem_ld_wait(); acc += v[0] + v[63];
                }
            }
            if (!DEP) { if (STORE) tmem_st_wait(); else tmem_ld_wait(); }
        }
    }
    long long t1 = clock64();
    if (warp < (uint32_t)W && (tid & 31) == 0) { out[warp * 2] = t1 - t0; out[warp * 2 + 1] = acc; }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem_base_s);
}

template <int X, bool DEP, bool STORE> void run(const char *name, long long *d) {
    long long h[32];
    for (int W : {1, 2, 4, 8, 12}) {
        const int R = 4096;
        k<X, DEP, STORE><<<1, 512>>>(d, W, R);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
        long long mx = 0;
        for (int w = 0; w < W; w++) mx = h[2 * w] > mx ? h[2 * w] : mx;
        const double bytes = (double)R * 32 * X * 4;
        printf("%-28s W=%2d  %7.1f cyc/instr/warp  %6.1f B/cyc/warp  %7.1f B/cyc/SM\n", name, W, (double)mx / R, bytes / mx, bytes * W / mx);
    }
}

int main() {
    CK(cudaSetDevice(0));
    long long *d;
    CK(cudaMalloc(&d, 512));
    run<16, false, false>("ld.32x32b.x16 (4 in flight)", d);
    run<32, false, false>("ld.32x32b.x32 (4 in flight)", d);
    run<64, false, false>("ld.32x32b.x64 (4 in flight)", d);
    run<16, true, false>("ld.32x32b.x16 (dependent)", d);
    run<64, true, false>("ld.32x32b.x64 (dependent)", d);
    run<16, false, true>("st.32x32b.x16 (4 in flight)", d);
    return 0;
}
