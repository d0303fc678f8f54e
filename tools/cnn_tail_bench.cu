// cnn_tail_bench.cu -- the depthwise tail of the CNN front-end (cnn_tcgen05.cu: cnn_tile_tail) timed in isolation: conv1 sums come
// from shared memory instead of TMEM, no barriers, W warps per SM sub-partition each running R tiles.  Answers: how many cycles
// does one (image, channel) tile cost a warp alone, and how well do 2 / 3 / 4 warps per sub-partition share the FMA and ALU pipes?
#include <cstdio>
#include <cstdlib>
#include "../bitnetmcu_b200/csrc/cnn_tcgen05.cu"
using namespace bnm;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

template <bool kPacked>
__global__ void __launch_bounds__(512, 1) k(long long *out, int *sink, int R) {
    __shared__ __align__(16) int s_rows[512 / 32][14 * 16];   // one block of 14 x 16 sums per warp (all lanes read the same values)
    const uint32_t warp = threadIdx.x >> 5;
    for (int i = threadIdx.x & 31; i < 14 * 16; i += 32) s_rows[warp][i] = (int)((i * 2654435761u + warp * 97u) % 60000u) - 20000;
    CnnTailW W;
    W.w01[0] = 0x0305; W.w01[1] = 0xfe07; W.w01[2] = 0x0109; W.wva = 0x02fd; W.wvb = 0x0402; W.wsa = 0x0006; W.wsb = 0x00fb;
    for (int i = 0; i < 9; i++) W.k3[i] = 0x0203 + i;
    __syncthreads();
    int acc = 0;
    const uint32_t tm = smem_u32(&s_rows[warp][0]);
    long long t0 = clock64();
    for (int r = 0; r < R; r++) {
        int f[4] = {0, 0, 0, 0};
        cnn_tile_tail<kPacked, true>(tm, 0, 0, 0, 0, 0, W, f, nullptr);
        acc += f[0] + f[1] + f[2] + f[3];
        W.w01[0] ^= acc & 1;   // keep the compiler from hoisting the tile out of the loop
    }
    long long t1 = clock64();
    if (acc == 0x12345678) *sink = acc;
    if ((threadIdx.x & 31) == 0) out[warp] = t1 - t0;
}

int main() {
    CK(cudaSetDevice(0));
    long long *d, h[16]; int *sink;
    CK(cudaMalloc(&d, 128)); CK(cudaMalloc(&sink, 4));
    const int R = 200;
    for (int packed = 0; packed < 2; packed++)
        for (int w : {1, 2, 3, 4}) {
            if (packed) k<true><<<1, 128 * w>>>(d, sink, R); else k<false><<<1, 128 * w>>>(d, sink, R);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h, d, 128, cudaMemcpyDeviceToHost));
            long long mx = 0;
            for (int i = 0; i < 4 * w; i++) mx = h[i] > mx ? h[i] : mx;
            printf("conv3 %-6s %d warps/SMSP: %8.1f cycles per tile per warp, %8.1f cycles per tile per SMSP\n", packed ? "IDP.2A" : "IMAD", w, (double)mx / R, (double)mx / R / w);
        }
    return 0;
}
