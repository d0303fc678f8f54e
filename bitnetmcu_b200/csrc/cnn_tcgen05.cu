// cnn_tcgen05.cu -- CNN front-end (BitNetMCU_MNIST_dll.c:64-80) with conv1 on the tensor cores.
//
// conv1 is the one convolution of the chain with a real reduction to share: 1 input channel -> C output channels,
// out1[c][p] = sum_k w1[c][k] * patch[p][k].  im2col'd it is a GEMM with M = 196 positions x images, N = C, K = 9 (padded
// to one K = 32 tcgen05 step): two MMA instructions per image instead of 588 dp4a per (image, channel) thread.  conv2 and
// conv3 are depthwise (groups = C, models.py:111-118): as GEMMs they would be block diagonal (< 2 % utilisation), so they
// stay on the CUDA cores exactly as in k_cnn_frontend16 and read conv1's output from shared memory.
//
// Per group of ipb = 256 / C images (one CTA, 256 threads, persistent over groups):
//   1. images -> smem, 4-byte sliding windows per image row (as k_cnn_frontend16)
//   2. A operand: thread = output position, 9 patch bytes + 7 zeros -> one 16-byte store into the no-swizzle K-major
//      core-matrix layout (8 rows x 16 B; the second K half stays zero); B operand = w1 in the same layout, built once
//   3. one elected thread: tcgen05.mma kind::i8 M=128 N=round_up(C,16) K=32 per 128 positions, D int32 in TMEM
//   4. epilogue: thread = position row, tcgen05.ld, ReLU >> 4 (inference.c:261-272) -> int16 planes [image][channel][196]
//   5. thread = (image, channel): conv2 (dp2a) + pool + conv3 (IMAD) + pool + ReLUNorm over the C*4 features (dll.c:80)
// All integer: bit-exact with the reference (27 CNN parity tests pass with BNM_CNN_TC=1, memcheck clean).
// STATUS: experimental, off by default.  This first version runs the five phases one after the other in a single CTA per SM
// (137 kB of shared memory for the int16 planes) and measures 0.093 G images/s (CNN-64) against 0.142 for k_cnn_frontend16,
// whose two CTAs per SM keep 16 warps on the FMA-bound conv2.  What it establishes is the building block: conv1 as an
// im2col'd tcgen05 GEMM, bit-exact.  To win it has to overlap the phases (two smaller CTAs per SM, or producer/consumer
// warps with double-buffered planes) -- DESIGN.md section 7.
#include <cstdio>
#include <cstdlib>

#include "kernels.h"
#include "sm100_ptx.cuh"

namespace bnm {

constexpr int kTcThreads = 256;
constexpr uint32_t kPlaneStride = 198;   // int16 per (image, channel) plane: 196 values + 2 pad = 99 words -> conflict-free over channels

// kVer 1: 256 threads, all M tiles of a group in one MMA batch (up to 512 TMEM columns), one CTA per SM (validated, slow).
// kVer 2: 128 threads, ipb = 128 / C images per group, one M tile at a time through a single 64-column accumulator, so three
//         CTAs fit an SM (70 kB of shared memory, 64 TMEM columns each) and overlap each other's phases.  BNM_CNN_TC=2:
//         bit-exact on hardware for both CNN fixtures (tools/cnn_tc_debug.py); throughput not measured yet.
template <int kVer>
__global__ void __launch_bounds__(kVer == 1 ? kTcThreads : 128, kVer == 1 ? 1 : 3)
k_cnn_frontend16_tc(const int8_t *__restrict__ images, const int8_t *__restrict__ w1, const int8_t *__restrict__ w2,
                    const int8_t *__restrict__ w3, uint32_t C, uint32_t n_pad, uint32_t ipb, uint32_t n_mtiles,
                    int8_t *__restrict__ feats, uint32_t feat_stride, size_t n, int *err) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_mma;
    __shared__ uint32_t tmem_base_s;
    __shared__ int s_max[16];
    const uint32_t t = threadIdx.x, lane = t & 31, n_thr = blockDim.x;
    const uint32_t warp = __shfl_sync(0xffffffffu, t >> 5, 0);
    constexpr uint32_t kTmemCols = kVer == 1 ? 512 : 64;

    uint8_t *base = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    uint32_t *s_img = reinterpret_cast<uint32_t *>(base);             // [ipb][64] words = 16 rows x 16 bytes
    uint32_t *s_win = s_img + ipb * 64;                               // [ipb][16][14] sliding 4-byte windows
    uint8_t *s_a = reinterpret_cast<uint8_t *>(s_win + ipb * 224);    // [n_mtiles][128 rows x 32 B], canonical no-swizzle K-major
    s_a += (128u - (smem_u32(s_a) & 127u)) & 127u;
    uint8_t *s_b = s_a + n_mtiles * 4096;                             // [n_pad rows x 32 B], same layout
    uint16_t *s_c1 = reinterpret_cast<uint16_t *>(s_b + n_pad * 32);  // [ipb][C][kPlaneStride] conv1 outputs (ReLU >> 4 < 2^15)

    // ---- one-time setup
    if (t == 0) { mbar_init(&bar_mma, 1); fence_mbar_init(); }
    if (warp == 1) tmem_alloc<kTmemCols>(&tmem_base_s);
    for (uint32_t i = t; i < n_mtiles * 256; i += n_thr) reinterpret_cast<uint4 *>(s_a)[i] = make_uint4(0, 0, 0, 0);
    for (uint32_t i = t; i < n_pad * 32; i += n_thr) {
        const uint32_t nn = i >> 5, k = i & 31;
        const int8_t v = (nn < C && k < 9) ? w1[nn * 9 + k] : (int8_t)0;
        s_b[(nn >> 3) * 256 + (k >> 4) * 128 + (nn & 7) * 16 + (k & 15)] = (uint8_t)v;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const uint64_t b_desc = make_smem_desc(smem_u32(s_b), 128, 256, UMMA_LAYOUT_NONE);
    const uint32_t idesc = make_idesc_i8(128, n_pad);
    uint32_t mma_phase = 0;

    const bool active = t < ipb * C;
    const uint32_t il = active ? t / C : 0, ch = active ? t % C : 0;
    int w2p[3], k3[9], w2v, w2s;   // conv2: (w0, w1) per kernel row + the third column; conv3: scalars
    {
        const int8_t *b = w2 + ch * 9, *c = w3 + ch * 9;
#pragma unroll
        for (int r = 0; r < 3; r++) w2p[r] = (int)((uint32_t)(uint8_t)b[3 * r] | ((uint32_t)(uint8_t)b[3 * r + 1] << 8));
#pragma unroll
        for (int i = 0; i < 9; i++) k3[i] = c[i];
        w2v = (int)((uint32_t)(uint8_t)b[2] | ((uint32_t)(uint8_t)b[5] << 8));
        w2s = (int)(uint32_t)(uint8_t)b[8];
    }

    const uint32_t n_pos = ipb * 196;
    const size_t n_groups = (n + ipb - 1) / ipb;
    for (size_t g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const size_t img_base = g * ipb;
        // ---- 1. images and their sliding windows
        for (uint32_t idx = t; idx < ipb * 64; idx += n_thr) {
            const size_t img = img_base + idx / 64;
            s_img[idx] = img < n ? reinterpret_cast<const uint32_t *>(images)[img * 64 + (idx & 63)] : 0u;
        }
        if (t < ipb) s_max[t] = 0;
        __syncthreads();
        for (uint32_t idx = t; idx < ipb * 224; idx += n_thr) {
            const uint32_t im = idx / 224, r = (idx % 224) / 14, x = idx % 14;
            const uint32_t lo = s_img[im * 64 + r * 4 + (x >> 2)];
            const uint32_t hi = (x >> 2) < 3 ? s_img[im * 64 + r * 4 + (x >> 2) + 1] : 0u;
            s_win[idx] = __funnelshift_r(lo, hi, 8 * (x & 3));
        }
        __syncthreads();
        // ---- 2. im2col rows: (a0 a1 a2 b0 | b1 b2 c0 c1 | c2 0 0 0 | 0 0 0 0) = taps in the order of w1[c][0..8]
        for (uint32_t R = t; R < n_pos; R += n_thr) {
            const uint32_t im = R / 196, pos = R % 196, y = pos / 14, x = pos % 14;
            const uint32_t *wrow = s_win + im * 224 + y * 14 + x;
            const uint32_t wa = wrow[0], wb = wrow[14], wc = wrow[28];
            uint4 q;
            q.x = __byte_perm(wa, wb, 0x4210);
            q.y = __byte_perm(wb, wc, 0x5421);
            q.z = (wc >> 16) & 0xffu;
            q.w = 0;
            const uint32_t tile = R >> 7, r = R & 127;
            *reinterpret_cast<uint4 *>(s_a + tile * 4096 + (r >> 3) * 256 + (r & 7) * 16) = q;
        }
        fence_proxy_async_smem();   // the tensor core (async proxy) reads what these generic stores wrote
        __syncthreads();
        // ---- 3 + 4. conv1 for all channels of all positions (one MMA per 128 positions), then ReLU >> 4 -> int16 planes
        //             [image][channel][position]; epilogue thread = position row (TMEM lane)
        auto drain_tile = [&](uint32_t tile, uint32_t d_col) {
            const uint32_t R = tile * 128 + (warp & 3) * 32 + lane;
            const bool valid = R < n_pos;
            const uint32_t im = valid ? R / 196 : 0, pos = valid ? R % 196 : 0;
            uint16_t *dst = s_c1 + (size_t)im * C * kPlaneStride + pos;
            for (uint32_t c0 = 0; c0 < n_pad; c0 += 16) {
                uint32_t x[16];
                tmem_ld_x16(tmem_base + (((warp & 3) * 32) << 16) + d_col + c0, x);
                tmem_ld_wait();
                if (valid) {
#pragma unroll
                    for (int j = 0; j < 16; j++)
                        if (c0 + j < C) dst[(c0 + j) * kPlaneStride] = (uint16_t)(max((int)x[j], 0) >> 4);
                }
            }
        };
        if (kVer == 1) {
            if (warp == 0) {
                tc_fence_after();
                if (elect_one()) {
                    for (uint32_t tile = 0; tile < n_mtiles; tile++)
                        umma_i8_ss(tmem_base + tile * n_pad, make_smem_desc(smem_u32(s_a) + tile * 4096, 128, 256, UMMA_LAYOUT_NONE), b_desc, idesc, 0);
                    umma_commit(&bar_mma);
                }
                __syncwarp();
            }
            mbar_wait(&bar_mma, mma_phase, err, 7);
            mma_phase ^= 1;
            tc_fence_after();
            // warps w and w+4 share a lane quarter and take alternate tiles
            for (uint32_t tile = warp >> 2; tile < n_mtiles; tile += 2) drain_tile(tile, tile * n_pad);
            tc_fence_before();
            __syncthreads();
        } else {
            for (uint32_t tile = 0; tile < n_mtiles; tile++) {   // one accumulator, tile by tile; the other CTAs of the SM fill the gaps
                if (warp == 0) {
                    tc_fence_after();
                    if (elect_one()) {
                        umma_i8_ss(tmem_base, make_smem_desc(smem_u32(s_a) + tile * 4096, 128, 256, UMMA_LAYOUT_NONE), b_desc, idesc, 0);
                        umma_commit(&bar_mma);
                    }
                    __syncwarp();
                }
                mbar_wait(&bar_mma, mma_phase, err, 7);
                mma_phase ^= 1;
                tc_fence_after();
                drain_tile(tile, 0);
                tc_fence_before();
                __syncthreads();   // every warp has read D before the next MMA overwrites it
            }
        }
        // ---- 5. depthwise tail on the CUDA cores: thread = (image, channel)
        int f[4] = {0, 0, 0, 0};
        if (active) {
            const uint32_t *pw = reinterpret_cast<const uint32_t *>(s_c1 + (size_t)(il * C + ch) * kPlaneStride);   // 7 words per row
            uint32_t c1[3][14];   // rolling conv1 rows as int16 pairs: c1[.][x] = (v[x], v[x+1]), c1[.][13] = (v[13], 0)
            int c2e[12], pl[3][6], c3e[4];
#pragma unroll
            for (int y = 0; y < 14; y++) {
                uint32_t W[7];
#pragma unroll
                for (int j = 0; j < 7; j++) W[j] = pw[y * 7 + j];
#pragma unroll
                for (int x = 0; x < 14; x++)
                    c1[y % 3][x] = (x & 1) ? __byte_perm(W[x >> 1], x < 13 ? W[(x >> 1) + 1] : 0u, 0x5432) : W[x >> 1];
                if (y >= 2) {
                    const int r = y - 2;  // conv2 output row
                    int v[12];
#pragma unroll
                    for (int x = 0; x < 12; x++) {
                        int s = __dp2a_lo((int)c1[r % 3][x], w2p[0], 0);
                        s = __dp2a_lo((int)c1[(r + 1) % 3][x], w2p[1], s);
                        s = __dp2a_lo((int)c1[(r + 2) % 3][x], w2p[2], s);
                        s = __dp2a_lo((int)__byte_perm(c1[r % 3][x + 2], c1[(r + 1) % 3][x + 2], 0x5410), w2v, s);
                        s = __dp2a_lo((int)c1[(r + 2) % 3][x + 2], w2s, s);
                        v[x] = max(s, 0) >> 4;
                    }
                    if ((r & 1) == 0) {
#pragma unroll
                        for (int x = 0; x < 12; x++) c2e[x] = v[x];
                    } else {
                        const int p = r >> 1;  // pooled row 0..5
#pragma unroll
                        for (int j = 0; j < 6; j++) pl[p % 3][j] = max(max(c2e[2 * j], c2e[2 * j + 1]), max(v[2 * j], v[2 * j + 1]));
                        if (p >= 2) {
                            const int q = p - 2;  // conv3 output row 0..3
                            int u[4];
#pragma unroll
                            for (int x = 0; x < 4; x++) {
                                int s = 0;
#pragma unroll
                                for (int dr = 0; dr < 3; dr++)
#pragma unroll
                                    for (int dc = 0; dc < 3; dc++) s += k3[3 * dr + dc] * pl[(q + dr) % 3][x + dc];
                                u[x] = max(s, 0) >> 4;
                            }
                            if ((q & 1) == 0) {
#pragma unroll
                                for (int x = 0; x < 4; x++) c3e[x] = u[x];
                            } else {
#pragma unroll
                                for (int j = 0; j < 2; j++)
                                    f[(q >> 1) * 2 + j] = max(max(c3e[2 * j], c3e[2 * j + 1]), max(u[2 * j], u[2 * j + 1]));
                            }
                        }
                    }
                }
            }
            atomicMax(&s_max[il], max(max(f[0], f[1]), max(f[2], f[3])));
        }
        __syncthreads();
        if (active && img_base + il < n) {
            // ReLUNorm over the C*4 features of this image (dll.c:80); all features are >= 0 here
            const uint32_t shift = 32u - (uint32_t)__clz(s_max[il] >> 7);   // bit length of max >> 7 (inference.c:41-47)
            const int rounding = (int)((1u << shift) >> 1);
            uint32_t packed = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) packed |= (uint32_t)min(127, (f[j] + rounding) >> shift) << (8 * j);
            *reinterpret_cast<uint32_t *>(feats + (img_base + il) * feat_stride + ch * 4) = packed;
        }
        __syncthreads();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<kTmemCols>(tmem_base);
}

// returns false when the shape is not covered (the caller then uses k_cnn_frontend16)
bool launch_cnn_frontend_tc(const int8_t *images, const int8_t *w1, const int8_t *w2, const int8_t *w3, uint32_t channels,
                            uint32_t xy, int8_t *features, uint32_t feat_stride, size_t n, int sm_count, int *d_err, int version,
                            cudaStream_t st) {
    if (xy != 16 || channels < 16 || channels > 64) return false;
    const uint32_t threads = version == 2 ? 128 : kTcThreads;
    const uint32_t ipb = threads / channels;
    if (ipb == 0 || ipb > 16) return false;
    const uint32_t n_pad = (channels + 15) / 16 * 16;
    const uint32_t n_mtiles = (ipb * 196 + 127) / 128;
    if (version != 2 && n_mtiles * n_pad > 512) return false;
    const size_t smem = 256 + (size_t)ipb * (64 + 224) * 4 + (size_t)n_mtiles * 4096 + (size_t)n_pad * 32 +
                        (size_t)ipb * channels * kPlaneStride * 2 + 64;
    if (smem > 226 * 1024) return false;
    static size_t attr_bytes[2] = {0, 0};   // opt-in dynamic shared memory granted so far (static + dynamic must stay <= 227 kB)
    const int vi = version == 2 ? 1 : 0;
    if (smem > attr_bytes[vi]) {
        const cudaError_t e = vi ? cudaFuncSetAttribute(k_cnn_frontend16_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                                 : cudaFuncSetAttribute(k_cnn_frontend16_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            cudaGetLastError();   // not sticky: leave no stale error behind for the caller's launch check
            return false;
        }
        attr_bytes[vi] = smem;
    }
    const size_t n_groups = (n + ipb - 1) / ipb;
    const size_t max_ctas = (size_t)sm_count * (vi ? 3 : 1);
    const unsigned grid = (unsigned)(n_groups < max_ctas ? n_groups : max_ctas);
    if (vi) k_cnn_frontend16_tc<2><<<grid, threads, smem, st>>>(images, w1, w2, w3, channels, n_pad, ipb, n_mtiles, features, feat_stride, n, d_err);
    else k_cnn_frontend16_tc<1><<<grid, threads, smem, st>>>(images, w1, w2, w3, channels, n_pad, ipb, n_mtiles, features, feat_stride, n, d_err);
    return true;
}

}  // namespace bnm
