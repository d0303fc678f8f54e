// cnn_tcgen05.cu -- CNN front-end (BitNetMCU_MNIST_dll.c:64-80) with conv1 on the tensor cores and its result consumed
// straight out of TMEM by the depthwise tail.
//
// conv1 (processconv33ReLU #1, inference.c:238-277 called at dll.c:71) is the one convolution of the chain with a reduction to
// share: 1 input channel -> C output channels, out1[c][p] = sum_t w1[c][t] * patch[p][t].  conv2 / conv3 are depthwise
// (groups = C, models.py:111-118): as GEMMs they are block diagonal (< 2 % tensor utilisation) and stay on the CUDA cores.
// The GEMM is oriented so that the accumulator already has the layout the depthwise tail wants:
//
//     D[M = (image, channel)][N = position]  =  A[(image, channel)][K = (image slot, tap)]  x  B[position][K]^T
//
//   * a TILE is 128 consecutive (image, channel) items = the 128 TMEM lanes = the 128 threads of a consumer warpgroup:
//     thread r owns item r and reads ITS channel-image's 14 x 14 conv1 sums from its own TMEM lane with tcgen05.ld -- no
//     shared-memory round trip, no transposition;
//   * B is the im2col matrix of the images a tile touches: one 16-byte row (9 taps + padding) per (image, position),
//     positions laid out with a row stride of 16 (N = 224 = 14 rows x 16, two accumulator halves of 112 columns), images side
//     by side along K (16 bytes each), built by one producer warp per warpgroup with one 128-bit store per position;
//   * A holds w1: row (image slot s, channel c) carries w1[c][0..8] in K-chunk s and zeros elsewhere, so one MMA computes
//     conv1 of up to 128 / C images at once.  Tiles ignore image boundaries (item = image * C + channel), so every lane is
//     used for any C that is a multiple of 16; a GROUP = 128 / gcd(C, 128) images is a whole number of tiles.
//   * both operands use the no-swizzle K-major canonical layout (8 rows x 16 bytes core matrices) with K-chunks a whole
//     plane apart: [K chunk][row][16 B], LBO = plane size, SBO = 128.
//
// Per (image, channel) thread, after tcgen05.ld of a conv1 row (14 int32 sums):
//   ReLU >> 4 (inference.c:261-272)          : SHF.R.S32 + I2IP.U16.S32.SAT (packs two values, clamps negatives to 0)
//   conv2, 12 x 12 outputs x 9 taps (dll.c:72): IDP.2A over int16 pairs, 5 per output: rows are processed in pairs (y, y+1) that
//                                               share the vertical pair of their third column, so 6 PRMT per 24 outputs
//   maxpool (dll.c:73)                        : taken on the raw sums BEFORE ReLU >> 4 (both are monotone): VIMNMX3.RELU + VIMNMX +
//                                               SHF per pooled value instead of ReLU >> 4 on all 144 sums
//   conv3 (dll.c:74), maxpool (dll.c:76)      : IMAD (int32 x int8), pooled the same way
//   ReLUNorm over the C*4 features (dll.c:80)  : per-image maximum through shared memory once the group's tiles are done
// The FMA pipe (IDP / IMAD issue, 2 cycles per warp instruction) bounds it: 720 IDP.2A + 144 IMAD per item, against
// 588 IDP.4A + 720 IDP.2A + 196 IMAD for the CUDA-core kernel (generic_kernels.cu k_cnn_frontend16).
// All integer: bit-exact with the reference.
#include <cstdio>
#include <cstdlib>

#include "kernels.h"
#include "sm100_ptx.cuh"

namespace bnm {

constexpr uint32_t kCnnTcMaxThreads = 480;       // 3 consumer warpgroups (12 warps) + 3 producer / MMA-issuer warps
constexpr uint32_t kPlaneBytes = 224 * 16;       // im2col plane of one image: 14 rows x 16 positions x 16 bytes
constexpr uint32_t kPieceCols = 64;               // accumulator piece: four conv1 rows of 16 columns
constexpr uint32_t kMaxCnnWG = 3;                 // consumer warpgroups: 3 x (2 buffers x 64 columns) = 384 of the 512 TMEM columns

struct CnnTcParams {
    const int8_t *images;
    const int8_t *w1, *w2, *w3;
    int8_t *feats;
    uint32_t feat_stride;
    size_t n;
    int *err;
    uint32_t C;
    uint32_t G, T;                 // images / tiles per group
    uint32_t a_plane_bytes;        // 128 rows x 16 B = 2048 (K-chunk stride of A)
    uint32_t a_phase_bytes;        // one tile phase of A: n_chunks_max * 2048
    uint32_t n_chunks_max;         // K-chunks (image slots) per tile, rounded up to even
    // shared memory carve-up (byte offsets from the 128-aligned base)
    uint32_t off_a, off_b, off_img, off_raw, off_wc, off_imax;
    uint32_t b_buf_bytes;          // one im2col buffer: the images of one tile + 1 spare plane + 32 rows
    uint32_t raw_buf_bytes;        // T * 128 * 16
    size_t n_groups;
    uint32_t n_wg;                 // consumer warpgroups (3 when the im2col buffers of three fit in shared memory, else 2)
};

__device__ __forceinline__ uint32_t pack_relu_u16(int hi, int lo) {
    uint32_t d;   // (clamp(hi, 0, 65535) << 16) | clamp(lo, 0, 65535): the ReLU comes with the conversion
    asm("cvt.pack.sat.u16.s32 %0, %1, %2;" : "=r"(d) : "r"(hi), "r"(lo));
    return d;
}

// (uint16, uint16) . (int8, int8) + c: pooled conv2 outputs reach 65535, beyond the int16 range of __dp2a_lo(int, int, int)
__device__ __forceinline__ int dp2a_u16_s8(uint32_t a, int b, int c) {
    int d;
    asm("dp2a.lo.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// per-channel weights of the depthwise tail, 16 words
struct CnnTailW {
    int w01[3];        // conv2 kernel row r: (w[r][0], w[r][1]) as int8 pair in the low two bytes
    int wva, wvb;      // third column, vertical pairs: (w[1][2], w[2][2]) for the upper row of a pair, (w[0][2], w[1][2]) for the lower
    int wsa, wsb;      // third column, single taps: (w[0][2], 0) and (w[2][2], 0)
    int k3[9];         // conv3: scalars, or -- kConv3Packed -- k3[0..6] = the same seven pair words for conv3's kernel
};

// One tile of one consumer thread: conv1 sums from TMEM -> 4 raw features (before ReLUNorm) of this (image, channel).
// tm = TMEM address of this thread's lane quarter at the warpgroup's first accumulator column.
// kConv3Packed: conv3's inputs (pooled conv2 outputs) are known to fit 16 bits for this model's weights (static bound, checked on
// the host), so conv3 runs on IDP.2A like conv2 (80 instead of 144 FMA-pipe instructions); otherwise int32 x int8 IMADs.
// kFromSmem (tools/cnn_tail_bench.cu only): the conv1 sums are read from shared memory at byte address tm + 64 * row instead of
// TMEM and no barrier is touched, so the arithmetic of the tail can be timed in isolation.
template <bool kConv3Packed, bool kFromSmem = false>
__device__ __forceinline__ void cnn_tile_tail(uint32_t tm, uint32_t bar_full0, uint32_t bar_full1, uint32_t bar_free0,
                                              uint32_t bar_free1, uint32_t parity, const CnnTailW &W, int (&f)[4], int *err) {
    auto load_row = [&](int y, uint32_t (&r)[16]) {
        if (kFromSmem) {
#pragma unroll
            for (int q = 0; q < 4; q++)
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r[4 * q]), "=r"(r[4 * q + 1]), "=r"(r[4 * q + 2]), "=r"(r[4 * q + 3])
                             : "r"(tm + 64 * y + 16 * q));
        } else {
            tmem_ld_x16(tm + ((y >> 2) & 1) * 64 + (y & 3) * 16, r);   // piece y / 4 lives in buffer (y / 4) & 1
        }
    };
    auto wait_row = [&](uint32_t (&r)[16]) { if (!kFromSmem) tmem_ld_wait_x16(r); };
    uint32_t E[14][7], O[14][7];   // conv1 rows as int16 pairs: E[y][j] = (v[2j], v[2j+1]), O[y][j] = (v[2j+1], v[2j+2]); 4 rows live
    int pl[6][6];                  // pooled conv2 rows (3 live)
    int c3e[4];
    uint32_t PE[6][3], PO[6][3];   // kConv3Packed: pooled rows as uint16 pairs
    uint32_t cur[16], nxt[16];
    // Accumulator pieces: conv1 rows 4k .. 4k+3 of a tile are piece k (64 columns), pieces alternate between the warpgroup's two
    // TMEM buffers.  Buffer b is filled twice per tile (pieces b and b + 2): its barrier phases are 2 t and 2 t + 1 for tile t, so
    // the wait parity of piece k is (k >> 1) whatever the tile -- compile-time constants in this unrolled code.
    if (!kFromSmem) {
        mbar_wait_a(bar_full0, 0, err, 11);
        tc_fence_after();
    }
    load_row(0, cur);
    wait_row(cur);
#pragma unroll
    for (int y = 0; y < 14; y++) {
        if (y + 1 < 14) {
            if (((y + 1) & 3) == 0 && !kFromSmem) {   // first row of the next piece
                mbar_wait_a((((y + 1) >> 2) & 1) ? bar_full1 : bar_full0, ((y + 1) >> 3) & 1, err, 12);
                tc_fence_after();
            }
            load_row(y + 1, nxt);   // prefetch the next conv1 row while this one is processed
        }
        {   // ReLU >> 4 of conv1 (inference.c:261-272) and int16 pairing
            int t[14];
#pragma unroll
            for (int x = 0; x < 14; x++) t[x] = (int)cur[x] >> 4;   // floor(s / 16); negatives stay negative and clamp to 0 below
#pragma unroll
            for (int j = 0; j < 7; j++) E[y][j] = pack_relu_u16(t[2 * j + 1], t[2 * j]);
#pragma unroll
            for (int j = 0; j < 6; j++) O[y][j] = __byte_perm(E[y][j], E[y][j + 1], 0x5432);
            O[y][6] = E[y][6] >> 16;
        }
        if (y >= 3 && (y & 1)) {
            // conv2 output rows (y-3, y-2) from conv1 rows R0..R3 = y-3..y, pooled into row p
            const int p = (y - 3) >> 1, R0 = y - 3, R1 = y - 2, R2 = y - 1, R3 = y;
#define BNM_P(r, x) (((x) & 1) ? O[r][(x) >> 1] : E[r][(x) >> 1])
#pragma unroll
            for (int j = 0; j < 6; j++) {
                int a[2], b[2];
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int x = 2 * j + e;
                    const int v = (int)__byte_perm(BNM_P(R1, x + 2), BNM_P(R2, x + 2), 0x5410);   // (R1[x+2], R2[x+2])
                    int s = __dp2a_lo((int)BNM_P(R0, x), W.w01[0], 0);
                    s = __dp2a_lo((int)BNM_P(R1, x), W.w01[1], s);
                    s = __dp2a_lo((int)BNM_P(R2, x), W.w01[2], s);
                    s = __dp2a_lo(v, W.wva, s);
                    a[e] = __dp2a_lo((int)BNM_P(R0, x + 2), W.wsa, s);
                    int u = __dp2a_lo((int)BNM_P(R1, x), W.w01[0], 0);
                    u = __dp2a_lo((int)BNM_P(R2, x), W.w01[1], u);
                    u = __dp2a_lo((int)BNM_P(R3, x), W.w01[2], u);
                    u = __dp2a_lo(v, W.wvb, u);
                    b[e] = __dp2a_lo((int)BNM_P(R3, x + 2), W.wsb, u);
                }
                pl[p][j] = max(__vimax3_s32_relu(a[0], a[1], b[0]), b[1]) >> 4;   // maxpool, then ReLU >> 4 (monotone: same result)
            }
#undef BNM_P
            if (!kConv3Packed) {
                if (p >= 2) {
                    const int q = p - 2;   // conv3 output row
                    int u[4];
#pragma unroll
                    for (int x = 0; x < 4; x++) {
                        int s = 0;
#pragma unroll
                        for (int dr = 0; dr < 3; dr++)
#pragma unroll
                            for (int dc = 0; dc < 3; dc++) s += W.k3[3 * dr + dc] * pl[q + dr][x + dc];
                        u[x] = s;
                    }
                    if ((q & 1) == 0) {
#pragma unroll
                        for (int x = 0; x < 4; x++) c3e[x] = u[x];
                    } else {
#pragma unroll
                        for (int j = 0; j < 2; j++)
                            f[(q >> 1) * 2 + j] = max(__vimax3_s32_relu(c3e[2 * j], c3e[2 * j + 1], u[2 * j]), u[2 * j + 1]) >> 4;
                    }
                }
            } else {
                // pooled row p as uint16 pairs, then conv3 exactly like conv2: output rows (0,1) after pooled row 3, (2,3) after row 5
#pragma unroll
                for (int j = 0; j < 3; j++) PE[p][j] = __byte_perm((uint32_t)pl[p][2 * j], (uint32_t)pl[p][2 * j + 1], 0x5410);
#pragma unroll
                for (int j = 0; j < 2; j++) PO[p][j] = __byte_perm((uint32_t)pl[p][2 * j + 1], (uint32_t)pl[p][2 * j + 2], 0x5410);
                PO[p][2] = (uint32_t)pl[p][5];
                if (p == 3 || p == 5) {
                    const int Q0 = p - 3, Q1 = p - 2, Q2 = p - 1, Q3 = p;
#define BNM_Q(r, x) (((x) & 1) ? PO[r][(x) >> 1] : PE[r][(x) >> 1])
#pragma unroll
                    for (int j = 0; j < 2; j++) {
                        int a[2], b[2];
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            const int x = 2 * j + e;
                            const int v = (int)__byte_perm(BNM_Q(Q1, x + 2), BNM_Q(Q2, x + 2), 0x5410);
                            int s = dp2a_u16_s8(BNM_Q(Q0, x), W.k3[0], 0);
                            s = dp2a_u16_s8(BNM_Q(Q1, x), W.k3[1], s);
                            s = dp2a_u16_s8(BNM_Q(Q2, x), W.k3[2], s);
                            s = dp2a_u16_s8((uint32_t)v, W.k3[3], s);
                            a[e] = dp2a_u16_s8(BNM_Q(Q0, x + 2), W.k3[5], s);
                            int u = dp2a_u16_s8(BNM_Q(Q1, x), W.k3[0], 0);
                            u = dp2a_u16_s8(BNM_Q(Q2, x), W.k3[1], u);
                            u = dp2a_u16_s8(BNM_Q(Q3, x), W.k3[2], u);
                            u = dp2a_u16_s8((uint32_t)v, W.k3[4], u);
                            b[e] = dp2a_u16_s8(BNM_Q(Q3, x + 2), W.k3[6], u);
                        }
                        f[(p == 5 ? 2 : 0) + j] = max(__vimax3_s32_relu(a[0], a[1], b[0]), b[1]) >> 4;
                    }
#undef BNM_Q
                }
            }
        }
        if (y + 1 < 14) {
            wait_row(nxt);
            if (!kFromSmem && ((((y + 1) & 3) == 3) || y + 1 == 13)) {   // the last row of a piece is in registers: its buffer may be overwritten
                tc_fence_before();
                mbar_arrive_a((((y + 1) >> 2) & 1) ? bar_free1 : bar_free0);
            }
#pragma unroll
            for (int x = 0; x < 16; x++) cur[x] = nxt[x];
        }
    }
}

template <bool kConv3Packed>
__global__ void __launch_bounds__(kCnnTcMaxThreads, 1) k_cnn_frontend16_tc(const __grid_constant__ CnnTcParams P) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[kMaxCnnWG][2], bar_free[kMaxCnnWG][2];   // [warpgroup][accumulator buffer]
    __shared__ __align__(8) uint64_t bar_group[kMaxCnnWG];                              // [warpgroup]: a group's per-image maxima are published
    __shared__ uint32_t tmem_base_s;
    const uint32_t t = threadIdx.x, lane = t & 31;
    const uint32_t warp = __shfl_sync(0xffffffffu, t >> 5, 0);
    uint8_t *base = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    const uint32_t C = P.C, G = P.G, T = P.T, n_wg = P.n_wg, n_thr = blockDim.x;

    // ---- one-time setup: barriers, TMEM, the A operand (w1 per tile phase), per-channel tail weights
    if (t == 0) {
#pragma unroll
        for (int g = 0; g < (int)kMaxCnnWG; g++) {
#pragma unroll
            for (int h = 0; h < 2; h++) { mbar_init(&bar_full[g][h], 1); mbar_init(&bar_free[g][h], 128); }
            mbar_init(&bar_group[g], 128);
        }
        fence_mbar_init();
    }
    if (warp == n_wg * 4) tmem_alloc<512>(&tmem_base_s);
    {
        uint4 *a4 = reinterpret_cast<uint4 *>(base + P.off_a);
        const uint32_t total = T * P.n_chunks_max * 128;   // 16-byte rows
        for (uint32_t i = t; i < total; i += n_thr) {
            const uint32_t phase = i / (P.n_chunks_max * 128), rem = i % (P.n_chunks_max * 128), chunk = rem >> 7, r = rem & 127;
            const uint32_t item = phase * 128 + r, il = item / C, ch = item % C, i0 = (phase * 128) / C;
            uint32_t q[4] = {0, 0, 0, 0};
            if (il - i0 == chunk) {
                const int8_t *w = P.w1 + ch * 9;
#pragma unroll
                for (int k = 0; k < 9; k++) q[k >> 2] |= (uint32_t)(uint8_t)w[k] << (8 * (k & 3));
            }
            a4[i] = make_uint4(q[0], q[1], q[2], q[3]);   // [phase][chunk][row][16 B]
        }
        int *wc = reinterpret_cast<int *>(base + P.off_wc);
        for (uint32_t ch = t; ch < C; ch += n_thr) {
            const int8_t *b = P.w2 + ch * 9, *c = P.w3 + ch * 9;
            auto pair = [](int8_t lo, int8_t hi) { return (int)((uint32_t)(uint8_t)lo | ((uint32_t)(uint8_t)hi << 8)); };
            int *o = wc + ch * 16;
            o[0] = pair(b[0], b[1]); o[1] = pair(b[3], b[4]); o[2] = pair(b[6], b[7]);
            o[3] = pair(b[5], b[8]); o[4] = pair(b[2], b[5]);
            o[5] = pair(b[2], 0); o[6] = pair(b[8], 0);
            if (kConv3Packed) {
                o[7] = pair(c[0], c[1]); o[8] = pair(c[3], c[4]); o[9] = pair(c[6], c[7]);
                o[10] = pair(c[5], c[8]); o[11] = pair(c[2], c[5]);
                o[12] = pair(c[2], 0); o[13] = pair(c[8], 0);
                o[14] = o[15] = 0;
            } else {
#pragma unroll
                for (int k = 0; k < 9; k++) o[7 + k] = c[k];
            }
        }
        int *imax = reinterpret_cast<int *>(base + P.off_imax);
        for (uint32_t i = t; i < kMaxCnnWG * 3 * 8; i += n_thr) imax[i] = 0;
    }
    fence_proxy_async_smem();   // the tensor core (async proxy) reads A
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp >= n_wg * 4) {
        // ======================= producer + MMA issuer of warpgroup g (one warp) =======================
        const uint32_t g = warp - n_wg * 4;
        const bool leader = elect_one();
        uint32_t *s_img = reinterpret_cast<uint32_t *>(base + P.off_img + g * (8 * 256 + 64));
        const uint32_t n_ld = (G * 16 + 31) / 32;   // 16-byte loads per lane and group (<= 4)
        const uint32_t idesc = make_idesc_i8(128, kPieceCols);
        const uint32_t a_base = smem_u32(base + P.off_a);
        // The im2col operand is built per TILE: the images a tile touches (up to 128 / C + 1 of them; tiles of one group overlap in
        // at most one image, which is then built twice), double-buffered, so that the buffers of three warpgroups fit in shared
        // memory for every channel count.
        uint4 img_regs[4];
        auto unit_images = [&](size_t grp, uint32_t tau, size_t &first, uint32_t &count) {
            const uint32_t i0 = (tau * 128) / C, i1 = (tau * 128 + 127) / C;
            first = grp * G + i0;
            count = i1 - i0 + 1;
        };
        auto load_unit = [&](size_t grp, uint32_t tau) {   // global -> registers (latency hidden behind the previous tile's work)
            size_t first; uint32_t count;
            unit_images(grp, tau, first, count);
#pragma unroll
            for (uint32_t k = 0; k < 4; k++) {
                img_regs[k] = make_uint4(0, 0, 0, 0);
                const uint32_t idx = k * 32 + lane;   // uint4 index inside the unit: image = idx / 16
                if (idx < count * 16 && first + idx / 16 < P.n)
                    img_regs[k] = reinterpret_cast<const uint4 *>(P.images)[first * 16 + idx];
            }
        };
        auto build_unit = [&](uint32_t buf, uint32_t count) {   // registers -> s_img -> im2col planes of buffer buf
            __syncwarp();
#pragma unroll
            for (uint32_t k = 0; k < 4; k++) {
                const uint32_t idx = k * 32 + lane;
                if (idx < count * 16) reinterpret_cast<uint4 *>(s_img)[idx] = img_regs[k];
            }
            __syncwarp();
            uint8_t *bbuf = base + P.off_b + (g * 2 + buf) * P.b_buf_bytes;
            // one task = four consecutive positions (y, 4 xg .. 4 xg + 3) of one image: six shared-memory words (two per image row)
            // feed four 16-byte im2col rows; 56 tasks per image (14 rows x 4 column groups, the last group holds x = 12, 13 only)
#pragma unroll 1
            for (uint32_t task = lane; task < count * 56; task += 32) {
                const uint32_t im = task / 56, rem = task - im * 56, y = rem >> 2, xg = rem & 3;
                const uint32_t *row = s_img + im * 64 + y * 4 + xg;
                const uint32_t a0 = row[0], a1 = row[1], b0 = row[4], b1 = row[5], c0 = row[8], c1 = row[9];
                uint8_t *dst = bbuf + im * kPlaneBytes + (y * 16 + xg * 4) * 16;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    if (i < 2 || xg < 3) {
                        const uint32_t wa = i ? __funnelshift_r(a0, a1, 8 * i) : a0, wb = i ? __funnelshift_r(b0, b1, 8 * i) : b0,
                                       wc = i ? __funnelshift_r(c0, c1, 8 * i) : c0;
                        // taps in the order of w1[c][0..8]: (a0 a1 a2 b0 | b1 b2 c0 c1 | c2 . . . | . . . .); bytes 9..15 meet zeros in A
                        uint4 q;
                        q.x = __byte_perm(wa, wb, 0x4210);
                        q.y = __byte_perm(wb, wc, 0x5421);
                        q.z = wc >> 16;
                        q.w = 0;
                        *reinterpret_cast<uint4 *>(dst + 16 * i) = q;
                    }
                }
            }
            fence_proxy_async_smem();   // the MMAs (async proxy) read what these generic-proxy stores wrote
            __syncwarp();
        };
        // piece k of a tile (conv1 rows 4k .. 4k+3 = im2col rows 64k .. 64k+63; the last piece runs 32 rows past the plane into
        // whatever follows -- rows nobody reads) goes to accumulator buffer k & 1 once its previous occupant (piece k - 2 of this
        // tile, or piece k + 2 of the tile before) has been drained: the (2 t + (k >> 1))-th fill waits for drain number one less,
        // i.e. parity (k >> 1) ^ 1 -- a fresh barrier passes a wait on parity 1
        auto issue_piece = [&](uint32_t buf, uint32_t tau, uint32_t count, uint32_t k) {
            const uint32_t n_k = (count + 1) / 2;   // K-steps of 32 bytes = two image slots
            const uint32_t b_base = smem_u32(base + P.off_b + (g * 2 + buf) * P.b_buf_bytes);
            const uint32_t h = k & 1;
            mbar_wait_relaxed(&bar_free[g][h], (k >> 1) ^ 1, P.err, 13);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + g * 2 * kPieceCols + h * kPieceCols;
            for (uint32_t s = 0; s < n_k; s++) {
                const uint64_t ad = make_smem_desc(a_base + tau * P.a_phase_bytes + s * 2 * P.a_plane_bytes, P.a_plane_bytes, 128, UMMA_LAYOUT_NONE);
                const uint64_t bd = make_smem_desc(b_base + s * 2 * kPlaneBytes + k * kPieceCols * 16, kPlaneBytes, 128, UMMA_LAYOUT_NONE);
                if (leader) umma_i8_ss(d_tmem, ad, bd, idesc, s != 0);
            }
            if (leader) umma_commit(&bar_full[g][h]);
            __syncwarp();
        };
        // tiles of this warpgroup, in order: groups j = (k * gridDim + blockIdx) * n_wg + g, tiles tau = 0 .. T-1 of each
        const size_t stride = (size_t)gridDim.x * n_wg;
        size_t j = (size_t)blockIdx.x * n_wg + g;
        uint32_t tau = 0, buf = 1;
        auto advance = [&](size_t &jj, uint32_t &tt) { if (++tt == T) { tt = 0; jj += stride; } };
        // Software pipeline over this warpgroup's tiles: while tile u is being issued (four pieces), tile u + 1 is built into the
        // other im2col buffer and the images of tile u + 2 are fetched into registers.  The first pass has no current tile and only
        // builds (one copy of every stage in the code: the kernel's hot instructions have to fit the 32 kB instruction cache).
        bool cur_valid = false;
        uint32_t count = 0;
        size_t jn = j; uint32_t tn = 0;
        if (jn < P.n_groups) load_unit(jn, tn);
        while (cur_valid || jn < P.n_groups) {
            bool built_next = false;
#pragma unroll 1
            for (uint32_t k = 0; k < 4; k++) {
                if ((k == 2 || !cur_valid) && !built_next && jn < P.n_groups) {
                    // k == 2: every MMA of the tile before the current one has completed (its pieces 2 and 3 were drained before
                    // the waits of this tile's pieces 0 and 1 passed), so the other im2col buffer is free
                    size_t nfirst; uint32_t ncount;
                    unit_images(jn, tn, nfirst, ncount);
                    build_unit(buf ^ 1, ncount);
                    size_t j2 = jn; uint32_t t2 = tn;
                    advance(j2, t2);
                    if (j2 < P.n_groups) load_unit(j2, t2);
                    built_next = true;
                }
                if (cur_valid) issue_piece(buf, tau, count, k);
            }
            // the tile just built becomes the current one
            cur_valid = built_next;
            if (built_next) {
                size_t f0;
                unit_images(jn, tn, f0, count);
                j = jn; tau = tn; buf ^= 1;
                advance(jn, tn);
            }
        }
    } else {
        // ======================= consumer warpgroup g: thread r = TMEM lane r =======================
        const uint32_t g = warp >> 2, r = t & 127;
        const uint32_t tm = tmem_base + (((warp & 3) * 32) << 16) + g * 2 * kPieceCols;
        const uint32_t bf0 = smem_u32(&bar_full[g][0]), bf1 = smem_u32(&bar_full[g][1]);
        const uint32_t be0 = smem_u32(&bar_free[g][0]), be1 = smem_u32(&bar_free[g][1]);
        const uint32_t bgrp = smem_u32(&bar_group[g]);
        const int *wc = reinterpret_cast<const int *>(base + P.off_wc);
        int *imax = reinterpret_cast<int *>(base + P.off_imax) + g * 24;   // [3][8]
        uint32_t parity = 0, gcount = 0;
        const size_t stride = (size_t)gridDim.x * n_wg;
        // ReLUNorm over the C*4 features of each image (dll.c:80) needs the maximum over ALL channels of the image, i.e. over other
        // threads' results.  It runs one group late: a group's threads publish their maxima (shared-memory atomicMax) and ARRIVE on
        // the warpgroup's mbarrier without waiting; the wait comes a whole group of work later, when the phase has long completed,
        // and only then are the (thread-private) raw features of the previous group normalised and stored.  No warp ever idles
        // at a barrier for its slower siblings.
        const uint32_t il_one = r / C, ch_one = r - il_one * C;   // T == 1 (C divides 128): the thread's image slot and channel never change
        auto normalise_group = [&](size_t jg, uint32_t gc) {
            const int4 *raw = reinterpret_cast<const int4 *>(base + P.off_raw + (g * 2 + (gc & 1)) * P.raw_buf_bytes);
            const int *imx = imax + (gc % 3) * 8;
            for (uint32_t it = 0; it < T; it++) {
                const uint32_t item = it * 128 + r, il = T == 1 ? il_one : item / C, ch = T == 1 ? ch_one : item - il * C;
                const size_t img = jg * G + il;
                const int m = imx[il];
                const uint32_t shift = 32u - (uint32_t)__clz(m >> 7);   // bit length of max >> 7 (inference.c:41-47); features are >= 0
                const int rounding = (int)((1u << shift) >> 1);
                const int4 v = raw[item];
                const uint32_t packed = (uint32_t)min(127, (v.x + rounding) >> shift) | ((uint32_t)min(127, (v.y + rounding) >> shift) << 8) |
                                        ((uint32_t)min(127, (v.z + rounding) >> shift) << 16) | ((uint32_t)min(127, (v.w + rounding) >> shift) << 24);
                if (img < P.n) *reinterpret_cast<uint32_t *>(P.feats + img * P.feat_stride + ch * 4) = packed;
            }
        };
        CnnTailW W;
        auto load_weights = [&](uint32_t ch) {
            const int4 *w4 = reinterpret_cast<const int4 *>(wc + ch * 16);
            const int4 q0 = w4[0], q1 = w4[1], q2 = w4[2], q3 = w4[3];
            W.w01[0] = q0.x; W.w01[1] = q0.y; W.w01[2] = q0.z; W.wva = q0.w;
            W.wvb = q1.x; W.wsa = q1.y; W.wsb = q1.z; W.k3[0] = q1.w;
            W.k3[1] = q2.x; W.k3[2] = q2.y; W.k3[3] = q2.z; W.k3[4] = q2.w;
            W.k3[5] = q3.x; W.k3[6] = q3.y; W.k3[7] = q3.z; W.k3[8] = q3.w;
        };
        load_weights(ch_one);
        size_t j_prev = 0;
        for (size_t j = (size_t)blockIdx.x * n_wg + g;; j_prev = j, j += stride, gcount++) {
            const bool have = j < P.n_groups;   // one extra pass drains the last group
            int4 *raw = reinterpret_cast<int4 *>(base + P.off_raw + (g * 2 + (gcount & 1)) * P.raw_buf_bytes);
            if (have) {
                for (uint32_t tau = 0; tau < T; tau++, parity ^= 1) {
                    const uint32_t item = tau * 128 + r;
                    if (T > 1) load_weights(item % C);   // T == 1: the thread's channel never changes, loaded once above
                    int f[4] = {0, 0, 0, 0};
                    cnn_tile_tail<kConv3Packed>(tm, bf0, bf1, be0, be1, parity, W, f, P.err);
                    raw[item] = make_int4(f[0], f[1], f[2], f[3]);   // thread-private slot: read back only by this thread
                }
            }
            // end of group gcount: (1) finish the previous group, (2) clear the maxima of the next one, (3) publish ours, (4) arrive
            if (gcount > 0) {
                mbar_wait_a(bgrp, (gcount - 1) & 1, P.err, 14);
                normalise_group(j_prev, gcount - 1);
            }
            if (!have) break;
            if (r < 8) imax[((gcount + 1) % 3) * 8 + r] = 0;
            {
                int *imx = imax + (gcount % 3) * 8;
                for (uint32_t it = 0; it < T; it++) {
                    const uint32_t item = it * 128 + r;
                    const int4 v = raw[item];
                    atomicMax(&imx[T == 1 ? il_one : item / C], max(max(v.x, v.y), max(v.z, v.w)));
                }
            }
            mbar_arrive_a(bgrp);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == n_wg * 4) tmem_dealloc<512>(tmem_base);
}

static uint32_t gcd_u32(uint32_t a, uint32_t b) { return b ? gcd_u32(b, a % b) : a; }

// geometry + shared-memory carve-up for a channel count; returns the dynamic shared memory needed, 0 when the shape is not covered
static size_t cnn_tc_plan(uint32_t channels, uint32_t xy, CnnTcParams &p) {
    if (xy != 16 || channels < 16 || channels > 128 || channels % 16) return 0;
    p.C = channels;
    p.G = 128 / gcd_u32(channels, 128);
    p.T = p.G * channels / 128;
    uint32_t max_imgs = 1;
    for (uint32_t tau = 0; tau < p.T; tau++) max_imgs = std::max(max_imgs, (tau * 128 + 127) / channels - (tau * 128) / channels + 1);
    p.n_chunks_max = (max_imgs + 1) / 2 * 2;
    p.a_plane_bytes = 128 * 16;
    p.a_phase_bytes = p.n_chunks_max * p.a_plane_bytes;
    // one tile's images + one spare plane (an odd image count reads one zero-weighted chunk past the last image) + 32 rows (the
    // last accumulator piece of a tile covers im2col rows 192..255 of a 224-row plane)
    p.b_buf_bytes = (max_imgs + 1) * kPlaneBytes + 32 * 16;
    p.raw_buf_bytes = p.T * 128 * 16;
    for (p.n_wg = kMaxCnnWG; p.n_wg >= 2; p.n_wg--) {   // three consumer warpgroups when their im2col buffers fit, else two
        uint32_t off = 0;
        auto take = [&](uint32_t bytes) { uint32_t o = off; off += (bytes + 127) / 128 * 128; return o; };
        p.off_a = take(p.T * p.a_phase_bytes);
        p.off_b = take(p.n_wg * 2 * p.b_buf_bytes);
        p.off_img = take(p.n_wg * (8 * 256 + 64));
        p.off_raw = take(p.n_wg * 2 * p.raw_buf_bytes);
        p.off_wc = take(channels * 64);
        p.off_imax = take(kMaxCnnWG * 3 * 8 * 4);
        const size_t smem = (size_t)off + 128;
        if (smem <= 226 * 1024) return smem;
    }
    return 0;
}

bool cnn_frontend_tc_supported(uint32_t channels, uint32_t xy) {
    CnnTcParams p{};
    return cnn_tc_plan(channels, xy, p) != 0;
}

// returns false when the shape is not covered (the caller then uses k_cnn_frontend16)
bool launch_cnn_frontend_tc(const int8_t *images, const int8_t *w1, const int8_t *w2, const int8_t *w3, uint32_t channels,
                            uint32_t xy, int8_t *features, uint32_t feat_stride, size_t n, int sm_count, int *d_err,
                            bool conv3_fits_u16, cudaStream_t st) {
    CnnTcParams p{};
    size_t smem = cnn_tc_plan(channels, xy, p);
    if (!smem) return false;
    smem = std::max<size_t>(smem, 116 * 1024);   // one CTA per SM (each allocates all 512 TMEM columns): more than half of the shared memory
    p.images = images; p.w1 = w1; p.w2 = w2; p.w3 = w3; p.feats = features; p.feat_stride = feat_stride; p.n = n; p.err = d_err;
    p.n_groups = (n + p.G - 1) / p.G;
    static size_t granted = 0;
    if (smem > granted) {
        if (cudaFuncSetAttribute(k_cnn_frontend16_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
            cudaFuncSetAttribute(k_cnn_frontend16_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            cudaGetLastError();   // not sticky: leave no stale error behind for the caller's launch check
            return false;
        }
        granted = smem;
    }
    const size_t want = (p.n_groups + p.n_wg - 1) / p.n_wg;
    const unsigned grid = (unsigned)std::min<size_t>(want, (size_t)sm_count);
    const unsigned threads = p.n_wg * 160;   // 4 consumer warps + 1 producer warp per warpgroup
    if (conv3_fits_u16) k_cnn_frontend16_tc<true><<<grid, threads, smem, st>>>(p);
    else k_cnn_frontend16_tc<false><<<grid, threads, smem, st>>>(p);
    return true;
}

}  // namespace bnm
