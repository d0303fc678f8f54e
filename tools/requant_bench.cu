// requant_bench.cu -- which instruction mix requantises 64 int32 accumulators -> 16 packed int8 words fastest?
//   A : VIADDMNMX.RELU + IMAD(2^k) + PRMT byte picks                       (current fc_chain_kernel epilogue)
//   B : mad.wide (x*2^(33-s) + C) hi word = 2v-128 | I2IP.S8.SAT | (w>>1 ^ 0x40..) & 0x7f..   (shift+round on the FMA pipe)
//   M : 3 words B + 1 word A per 16 accumulators
// Each variant is checked against the plain formula clamp((x+r)>>s,0,127) and timed with W warps per SMSP.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

struct Coef { int rounding, cap; uint32_t mult; uint32_t shift; int mw; long long cw; };
__device__ __forceinline__ Coef coef(int m) {
    Coef c;
    c.shift = 32u - (uint32_t)__clz(m >> 7);
    c.rounding = (int)((1u << c.shift) >> 1);
    c.cap = (int)((128u << c.shift) - 1u);
    c.mult = 1u << (24u - c.shift);
    // B: t = floor((x + r - 2^(s+6)) / 2^(s-1)) = hi32(x * 2^(33-s) + (r - 2^(s+6)) * 2^(33-s)),  valid for s >= 3
    c.mw = (int)(1u << (33u - c.shift));
    c.cw = ((long long)c.rounding - (1ll << (c.shift + 6))) * (long long)c.mw;
    return c;
}
__device__ __forceinline__ uint32_t packA(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, const Coef &c) {
    uint32_t z0, z1, z2, z3;
    asm("mul.lo.u32 %0, %1, %2;" : "=r"(z0) : "r"((uint32_t)__viaddmin_s32_relu((int)x0, c.rounding, c.cap)), "r"(c.mult));
    asm("mul.lo.u32 %0, %1, %2;" : "=r"(z1) : "r"((uint32_t)__viaddmin_s32_relu((int)x1, c.rounding, c.cap)), "r"(c.mult));
    asm("mul.lo.u32 %0, %1, %2;" : "=r"(z2) : "r"((uint32_t)__viaddmin_s32_relu((int)x2, c.rounding, c.cap)), "r"(c.mult));
    asm("mul.lo.u32 %0, %1, %2;" : "=r"(z3) : "r"((uint32_t)__viaddmin_s32_relu((int)x3, c.rounding, c.cap)), "r"(c.mult));
    uint32_t lo = __byte_perm(z0, z1, 0x0073), hi = __byte_perm(z2, z3, 0x0073);
    return __byte_perm(lo, hi, 0x5410);
}
// C : clamp(x + r, 0, cap) * mult  ==  min(max(x, 0), cap - r) * mult + r * mult : the rounding add moves into the IMAD's
//     addend and the clamp becomes a 2-input VIMNMX.RELU (full rate) instead of the 3-input VIADDMNMX.RELU (half rate)
__device__ __forceinline__ uint32_t packC(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, int capr, uint32_t mult, uint32_t radd) {
    uint32_t z0, z1, z2, z3;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(z0) : "r"((uint32_t)__vimin_s32_relu((int)x0, capr)), "r"(mult), "r"(radd));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(z1) : "r"((uint32_t)__vimin_s32_relu((int)x1, capr)), "r"(mult), "r"(radd));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(z2) : "r"((uint32_t)__vimin_s32_relu((int)x2, capr)), "r"(mult), "r"(radd));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(z3) : "r"((uint32_t)__vimin_s32_relu((int)x3, capr)), "r"(mult), "r"(radd));
    uint32_t lo = __byte_perm(z0, z1, 0x0073), hi = __byte_perm(z2, z3, 0x0073);
    return __byte_perm(lo, hi, 0x5410);
}
__device__ __forceinline__ int max16_2in(const uint32_t *v, int m) {   // 2-input VIMNMX chains (full rate)
    int m2 = 0, m3 = 0, m4 = 0;
#pragma unroll
    for (int j = 0; j < 16; j += 4) { m = max(m, (int)v[j]); m2 = max(m2, (int)v[j + 1]); m3 = max(m3, (int)v[j + 2]); m4 = max(m4, (int)v[j + 3]); }
    return max(max(m, m2), max(m3, m4));
}
__device__ __forceinline__ int hiB(uint32_t x, const Coef &c) {
    int d;   // t = (x >> (s-1)) - 127 = hi32(x * 2^(33-s)) - 127: ONE IMAD.HI (FMA pipe)
    asm("mad.hi.s32 %0, %1, %2, %3;" : "=r"(d) : "r"((int)x), "r"(c.mw), "r"(-127));
    return d;
}
__device__ __forceinline__ uint32_t packB(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, const Coef &c) {
    int t0 = hiB(x0, c), t1 = hiB(x1, c), t2 = hiB(x2, c), t3 = hiB(x3, c);
    uint32_t lo, w;
    asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(lo) : "r"(t3), "r"(t2), "r"(0));
    asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(w) : "r"(t1), "r"(t0), "r"(lo));
    return ((w >> 1) ^ 0x40404040u) & 0x7f7f7f7fu;
}
__device__ __forceinline__ int max16(const uint32_t *v, int m) {
    int m2 = 0;
#pragma unroll
    for (int j = 0; j < 16; j += 4) { m = __vimax3_s32(m, (int)v[j], (int)v[j + 1]); m2 = __vimax3_s32(m2, (int)v[j + 2], (int)v[j + 3]); }
    return max(m, m2);
}

template <int NB> __global__ void k(const int *in, uint32_t *out, long long *cyc, int iters) {
    uint32_t v[64];
    for (int i = 0; i < 64; i++) v[i] = in[(threadIdx.x % 256) * 64 + i];
    uint32_t acc[16] = {0};
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        int mc[4];
#pragma unroll
        for (int c = 0; c < 4; c++) mc[c] = NB >= 6 ? max16_2in(v + 16 * c, 0) : max16(v + 16 * c, 0);
        Coef k = coef(NB >= 6 ? max(max(mc[0], mc[1]), max(mc[2], mc[3])) : max(__vimax3_s32(mc[0], mc[1], mc[2]), mc[3]));
        if (NB >= 5) {
            const int capr = k.cap - k.rounding;
            const uint32_t radd = (uint32_t)k.rounding * k.mult;
#pragma unroll
            for (int w = 0; w < 16; w++) acc[w] ^= packC(v[4 * w], v[4 * w + 1], v[4 * w + 2], v[4 * w + 3], capr, k.mult, radd);
        } else if (NB > 0 && __all_sync(0xffffffffu, k.shift >= 3)) {
#pragma unroll
            for (int w = 0; w < 16; w++) {
                uint32_t r = (w & 3) < NB ? packB(v[4 * w], v[4 * w + 1], v[4 * w + 2], v[4 * w + 3], k)
                                          : packA(v[4 * w], v[4 * w + 1], v[4 * w + 2], v[4 * w + 3], k);
                acc[w] ^= r;
            }
        } else {
#pragma unroll
            for (int w = 0; w < 16; w++) acc[w] ^= packA(v[4 * w], v[4 * w + 1], v[4 * w + 2], v[4 * w + 3], k);
        }
        v[it & 63] += (acc[it & 15] & 1);   // light loop-carried dependency
    }
    long long t1 = clock64();
    for (int w = 0; w < 16; w++) out[threadIdx.x * 16 + w] = acc[w];
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int NB> void launch(const int *din, uint32_t *dout, long long *dc, int nt, int iters) { k<NB><<<1, nt>>>(din, dout, dc, iters); }

int main() {
    const int NT = 512;
    int *h = (int *)malloc(256 * 64 * 4);
    srand(1);
    for (int r = 0; r < 256; r++) {
        int scale = 1 << (5 + r % 16);
        if (getenv("BIGROWS")) scale = 1 << (12 + r % 8);
        for (int i = 0; i < 64; i++) h[r * 64 + i] = (rand() % (2 * scale)) - scale;
        if (r % 7 == 0) h[r * 64 + 5] = (128 << (r % 12)) - 1;           // top-of-bucket values: the (x+r)>>s == 128 clip case
        if (r % 11 == 0) for (int i = 0; i < 64; i++) h[r * 64 + i] = -abs(h[r * 64 + i]) - 1;
    }
    int *din; uint32_t *dout; long long *dc;
    CK(cudaMalloc(&din, 256 * 64 * 4)); CK(cudaMalloc(&dout, NT * 16 * 4)); CK(cudaMalloc(&dc, 8));
    CK(cudaMemcpy(din, h, 256 * 64 * 4, cudaMemcpyHostToDevice));
    uint32_t *ho = (uint32_t *)malloc(NT * 16 * 4);
    const char *names[7] = {"4A+0B (VIADDMNMX+IMAD+PRMT)", "3A+1B", "2A+2B", "1A+3B", "0A+4B (IMAD.HI+I2IP.S8+fix)",
                            "C (VIMNMX.RELU+IMAD(+r*mult)+PRMT)", "C2 (C + 2-input max chains)"};
    for (int var = 0; var < 7; var++) {
        // correctness: one iteration
        if (var == 0) launch<0>(din, dout, dc, 256, 1); else if (var == 1) launch<1>(din, dout, dc, 256, 1); else if (var == 2) launch<2>(din, dout, dc, 256, 1); else if (var == 3) launch<3>(din, dout, dc, 256, 1); else if (var == 4) launch<4>(din, dout, dc, 256, 1); else if (var == 5) launch<5>(din, dout, dc, 256, 1); else launch<6>(din, dout, dc, 256, 1);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(ho, dout, 256 * 16 * 4, cudaMemcpyDeviceToHost));
        long bad = 0;
        for (int r = 0; r < 256; r++) {
            int m = 0; for (int i = 0; i < 64; i++) m = h[r * 64 + i] > m ? h[r * 64 + i] : m;
            int s = 0; for (unsigned sc = (unsigned)(m >> 7); sc; sc >>= 1) s++;
            int rd = (1 << s) >> 1;
            for (int i = 0; i < 64; i++) {
                int x = h[r * 64 + i]; int y = x < 0 ? 0 : ((x + rd) >> s); if (y > 127) y = 127;
                int got = (ho[r * 16 + i / 4] >> (8 * (i % 4))) & 255;
                if (got != y) bad++;
            }
        }
        printf("%-28s correctness: %ld mismatches of %d\n", names[var], bad, 256 * 64);
        for (int wps : {1, 2, 3, 4}) {
            int nt = 128 * wps;
            if (var == 0) launch<0>(din, dout, dc, nt, 200); else if (var == 1) launch<1>(din, dout, dc, nt, 200); else if (var == 2) launch<2>(din, dout, dc, nt, 200); else if (var == 3) launch<3>(din, dout, dc, nt, 200); else if (var == 4) launch<4>(din, dout, dc, nt, 200); else if (var == 5) launch<5>(din, dout, dc, nt, 200); else launch<6>(din, dout, dc, nt, 200);
            CK(cudaDeviceSynchronize());
            long long c; CK(cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost));
            printf("    %d warps/SMSP: %7.1f cycles per 64-accumulator epilogue per warp  -> %6.1f per SMSP-epilogue\n", wps, c / 200.0, c / 200.0 / wps);
        }
    }
    return 0;
}
