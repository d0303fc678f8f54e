// umma_probe.cu -- standalone hardware probe for the tcgen05 kind::i8 building blocks of the fused FC kernel.
// One 128-thread CTA computes D[128 x N] = A[128 x K] * B[N x K]^T (int8 x int8 -> int32) with
//   A source : 0 = smem SWIZZLE_128B tile filled by TMA (layer 1 of the chain)
//              1 = TMEM, written by the threads with tcgen05.st (".ts" MMA; inter-layer activations)
//              2 = smem, written by the threads in the same layout variant as B
//   B layout : 0 = no-swizzle K-major (8x16B core matrices), one [N x 32B] tile per K-step
//              1 = SWIZZLE_32B K-major,                      one [N x 32B] tile per K-step
//              2 = SWIZZLE_128B K-major atoms [N x 128B] with in-atom K advance (needs K % 128 == 0)
// and checks the result against a CPU GEMM.  Build: see tools/Makefile.  Run under `timeout`.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../bitnetmcu_b200/csrc/sm100_ptx.cuh"

using namespace bnm;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

struct ProbeParams {
    const int8_t *A;      // [128][K] row-major (global)
    const int8_t *B;      // [N][K] row-major (global)
    int32_t *D;           // [128][N]
    int K, N, a_src, b_layout;
    int *err;
};

__device__ __forceinline__ uint32_t b_tile_offset(int layout, int n, int kb /*byte in 32B step*/) {
    // offset inside one [N x 32B] K-step tile
    if (layout == 0) return (n >> 3) * 256 + (kb >> 4) * 128 + (n & 7) * 16 + (kb & 15);        // core matrix 8 rows x 16B
    return (n >> 3) * 256 + (n & 7) * 32 + ((((kb >> 4) ^ ((n & 7) >> 2)) & 1) << 4) + (kb & 15);  // SW32: bit4 ^= bit7
}

__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ CUtensorMap tmapA, ProbeParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sA = smem;                       // up to 128 x 256 = 32 KB (SW128: 2 atoms of 16 KB)
    uint8_t *sB = smem + 32768;               // up to 256 x 256 = 64 KB
    __shared__ uint64_t bar_full, bar_mma;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int K = p.K, N = p.N, ksteps = K / 32;

    if (tid == 0) { mbar_init(&bar_full, 1); mbar_init(&bar_mma, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc<512>(&tmem_base_s);
    // ---- stage B (all threads; generic proxy writes)
    for (int idx = tid; idx < N * K; idx += 128) {
        int n = idx / K, k = idx % K;
        uint32_t off;
        if (p.b_layout == 2) off = (k >> 7) * (N * 128) + (n >> 3) * 1024 + (n & 7) * 128 + ((((k & 127) >> 4) ^ (n & 7)) << 4) + (k & 15);
        else off = (k >> 5) * (N * 32) + b_tile_offset(p.b_layout, n, k & 31);
        sB[off] = (uint8_t)p.B[idx];
    }
    if (p.a_src == 2) {
        for (int idx = tid; idx < 128 * K; idx += 128) {
            int m = idx / K, k = idx % K;
            uint32_t off = (k >> 5) * (128 * 32) + b_tile_offset(p.b_layout == 2 ? 0 : p.b_layout, m, k & 31);
            sA[off] = (uint8_t)p.A[idx];
        }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t d_tmem = tmem_base;            // columns [0, N)
    const uint32_t a_tmem = tmem_base + 256;      // columns [256, 256 + K/4)

    if (p.a_src == 0 && tid == 0) {
        mbar_arrive_expect_tx(&bar_full, 128 * ((K + 127) / 128) * 128);
        for (int a = 0; a < (K + 127) / 128; a++) tma_load_2d(sA + a * 16384, &tmapA, a * 128, 0, &bar_full);
    }
    if (p.a_src == 1) {
        // thread = row; pack its K int8 into K/4 words, little-endian (element 4j in the low byte), store to TMEM
        const uint32_t *row = reinterpret_cast<const uint32_t *>(p.A + (size_t)tid * K);
        for (int c = 0; c < K / 4; c += 8) {
            uint32_t v[8];
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] = row[c + j];
            tmem_st_x8(a_tmem + ((uint32_t)(warp * 32) << 16) + c, v);
        }
        tmem_st_wait();
        tc_fence_before();
    }
    __syncthreads();
    if (tid == 0) {
        tc_fence_after();
        if (p.a_src == 0) mbar_wait(&bar_full, 0, p.err, 11);
        tc_fence_after();
        const uint32_t idesc = make_idesc_i8(128, N);
        for (int s = 0; s < ksteps; s++) {
            uint64_t bdesc;
            if (p.b_layout == 2) bdesc = make_smem_desc(smem_u32(sB) + (s >> 2) * (N * 128) + (s & 3) * 32, 0, 1024, UMMA_LAYOUT_SW128);
            else if (p.b_layout == 0) bdesc = make_smem_desc(smem_u32(sB) + s * (N * 32), 128, 256, UMMA_LAYOUT_NONE);
            else bdesc = make_smem_desc(smem_u32(sB) + s * (N * 32), 0, 256, UMMA_LAYOUT_SW32);
            if (p.a_src == 1) {
                umma_i8_ts(d_tmem, a_tmem + s * 8, bdesc, idesc, s > 0);
            } else {
                uint64_t adesc;
                if (p.a_src == 0) adesc = make_smem_desc(smem_u32(sA) + (s >> 2) * 16384 + (s & 3) * 32, 0, 1024, UMMA_LAYOUT_SW128);
                else if (p.b_layout == 1) adesc = make_smem_desc(smem_u32(sA) + s * (128 * 32), 0, 256, UMMA_LAYOUT_SW32);
                else adesc = make_smem_desc(smem_u32(sA) + s * (128 * 32), 128, 256, UMMA_LAYOUT_NONE);
                umma_i8_ss(d_tmem, adesc, bdesc, idesc, s > 0);
            }
        }
        umma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, 0, p.err, 12);
    tc_fence_after();
    for (int c = 0; c < N; c += 16) {
        uint32_t v[16];
        tmem_ld_x16(d_tmem + ((uint32_t)(warp * 32) << 16) + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; j++) p.D[(size_t)tid * N + c + j] = (int32_t)v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem_base);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    int dev_count = 0;
    CK(cudaGetDeviceCount(&dev_count));
    CK(cudaSetDevice(0));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s sm_%d%d SMs=%d\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    EncodeTiledFn encode = (EncodeTiledFn)fn;

    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + 65536));
    int *d_err;
    CK(cudaMalloc(&d_err, 4));
    struct Case { int K, N, a_src, b_layout; };
    std::vector<Case> cases;
    for (int bl = 0; bl < 3; bl++) {
        cases.push_back({256, 64, 0, bl});     // layer 1: TMA SW128 A
        cases.push_back({64, 64, 1, bl == 2 ? 0 : bl});   // TS
        cases.push_back({64, 64, 2, bl == 2 ? 0 : bl});   // A by threads
    }
    cases.push_back({256, 160, 0, 0});
    cases.push_back({160, 160, 1, 0});
    cases.push_back({64, 16, 1, 0});
    cases.push_back({64, 16, 1, 1});
    cases.push_back({192, 96, 0, 0});          // K=192: second TMA box half out of bounds (zero fill)
    cases.push_back({32, 16, 1, 0});
    cases.push_back({96, 64, 1, 1});
    int n_fail = 0;
    for (auto c : cases) {
        std::vector<int8_t> A(128 * c.K), B(c.N * c.K);
        srand(c.K * 131 + c.N * 7 + c.a_src * 3 + c.b_layout);
        for (auto &v : A) v = (int8_t)(rand() % 256 - 128);
        for (auto &v : B) v = (int8_t)(rand() % 256 - 128);
        int8_t *dA, *dB;
        int32_t *dD;
        CK(cudaMalloc(&dA, A.size()));
        CK(cudaMalloc(&dB, B.size()));
        CK(cudaMalloc(&dD, 128 * c.N * 4));
        CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
        CK(cudaMemset(dD, 0xff, 128 * c.N * 4));
        CK(cudaMemset(d_err, 0, 4));
        CUtensorMap tmap;
        memset(&tmap, 0, sizeof(tmap));
        cuuint64_t gdim[2] = {(cuuint64_t)c.K, 128};
        cuuint64_t gstride[1] = {(cuuint64_t)c.K};
        cuuint32_t box[2] = {128, 128};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, dA, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed %d\n", (int)r); return 3; }
        ProbeParams p{dA, dB, dD, c.K, c.N, c.a_src, c.b_layout, d_err};
        probe_kernel<<<1, 128, 32768 + 65536>>>(tmap, p);
        cudaError_t e = cudaDeviceSynchronize();
        int err = 0;
        if (e != cudaSuccess) {
            printf("case K=%d N=%d a_src=%d b_layout=%d: KERNEL ERROR %s\n", c.K, c.N, c.a_src, c.b_layout, cudaGetErrorString(e));
            return 4;  // context is dead after a trap
        }
        CK(cudaMemcpy(&err, d_err, 4, cudaMemcpyDeviceToHost));
        std::vector<int32_t> D(128 * c.N);
        CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
        long bad = 0;
        for (int m = 0; m < 128; m++)
            for (int n = 0; n < c.N; n++) {
                int32_t s = 0;
                for (int k = 0; k < c.K; k++) s += (int32_t)A[m * c.K + k] * (int32_t)B[n * c.K + k];
                if (s != D[m * c.N + n]) bad++;
            }
        printf("case K=%3d N=%3d a_src=%d b_layout=%d: %s (%ld/%d mismatches, err=%d) D[0][0..3]=%d %d %d %d\n", c.K, c.N, c.a_src,
               c.b_layout, bad ? "FAIL" : "ok", bad, 128 * c.N, err, D[0], D[1], D[2], D[3]);
        if (bad) n_fail++;
        cudaFree(dA); cudaFree(dB); cudaFree(dD);
    }
    printf("probe done: %d failing cases\n", n_fail);
    return 0;
}
