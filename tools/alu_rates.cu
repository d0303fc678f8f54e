// alu_rates.cu -- per-SMSP issue rate of the integer instructions the ReLUNorm epilogue can be built from.
// One CTA; W warps per SMSP; each thread runs ITER x 16 independent ops.  Prints cycles per warp-instruction per SMSP.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

template <int OP> __device__ __forceinline__ int op(int a, int b, int c) {
    int d;
    if (OP == 0) asm volatile("add.s32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));                       // IADD / VIADD
    else if (OP == 1) d = __viaddmin_s32_relu(a, b, c);                                                  // VIADDMNMX.RELU
    else if (OP == 2) d = __vimax3_s32(a, b, c);                                                         // VIMNMX3
    else if (OP == 3) d = (int)__byte_perm((unsigned)a, (unsigned)b, 0x0073);                            // PRMT
    else if (OP == 4) asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));                // IMAD
    else if (OP == 5) asm volatile("shr.s32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));                   // SHF.R
    else if (OP == 6) asm volatile("lop3.b32 %0, %1, %2, %3, 0xFE;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); // LOP3 (a|b|c)
    else if (OP == 7) asm volatile("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); // I2IP
    else if (OP == 8) asm volatile("mul.hi.s32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));                // IMAD.HI
    else if (OP == 9) asm volatile("max.s32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));                   // VIMNMX (2-input)
    else if (OP == 10) asm volatile("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));   // IMAD with add
    else if (OP == 11) d = __vimax3_s16x2(a, b, c);
    else if (OP == 12) d = __viaddmin_s16x2_relu(a, b, c);
    else if (OP == 13) asm volatile("shf.l.wrap.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); // SHF funnel
    else if (OP == 14) asm volatile("bfe.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));       // BFE
    else if (OP == 15) asm volatile("dp4a.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));  // IDP.4A
    else if (OP == 16) asm volatile("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));  // IDP.2A.LO
    else if (OP == 17) asm volatile("dp2a.hi.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));  // IDP.2A.HI
    else if (OP == 18) asm volatile("dp2a.lo.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));  // IDP.2A.LO.U16.S8
    else if (OP == 19) asm volatile("cvt.pack.sat.u16.s32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));        // I2IP.U16.S32.SAT
    else if (OP == 20) asm volatile("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));    // IDP.4A.U8.S8
    else if (OP == 21) d = __vimax3_s32_relu(a, b, c);                                                     // VIMNMX3.RELU
    else if (OP == 22) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));      // FFMA (3 registers)
    else d = a;
    return d;
}

template <int OP> __global__ void k(long long *out, int *sink, int iters, int b, int c) {
    int v[16];
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = threadIdx.x * 17 + i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) v[i] = op<OP>(v[i], b, c);
    }
    long long t1 = clock64();
    int s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s ^= v[i];
    if (s == 0x12345678) *sink = s;
    if (threadIdx.x == 0) out[0] = t1 - t0;
}

// two independent instruction streams in one warp: 8 accumulators on OPA (FMA pipe), 8 on OPB (ALU pipe) -- do the pipes overlap?
template <int OPA, int OPB> __global__ void kmix(long long *out, int *sink, int iters, int b, int c) {
    int v[16];
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = threadIdx.x * 17 + i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) { v[2 * i] = op<OPA>(v[2 * i], b, c); v[2 * i + 1] = op<OPB>(v[2 * i + 1], b, c); }
    }
    long long t1 = clock64();
    int s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s ^= v[i];
    if (s == 0x12345678) *sink = s;
    if (threadIdx.x == 0) out[0] = t1 - t0;
}
template <int OPA, int OPB> void runmix(const char *name, long long *d, int *sink) {
    const int iters = 2000;
    printf("%-28s", name);
    for (int warps_per_smsp : {1, 2, 4}) {
        kmix<OPA, OPB><<<1, 128 * warps_per_smsp>>>(d, sink, iters, 3, 1000);
        CK(cudaDeviceSynchronize());
        long long h;
        CK(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost));
        printf("  %dw/SMSP: %5.2f cyc per (A+B) pair/SMSP", warps_per_smsp, (double)h / (iters * 8.0 * warps_per_smsp));
    }
    printf("\n");
}

template <int OP> void run(const char *name, long long *d, int *sink) {
    const int iters = 2000;
    printf("%-16s", name);
    for (int warps_per_smsp : {1, 2, 4}) {
        k<OP><<<1, 128 * warps_per_smsp>>>(d, sink, iters, 3, 1000);
        CK(cudaDeviceSynchronize());
        long long h;
        CK(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost));
        printf("  %dw/SMSP: %5.2f cyc/warp-instr/SMSP", warps_per_smsp, (double)h / (iters * 16.0 * warps_per_smsp));
    }
    printf("\n");
}

int main() {
    long long *d; int *sink;
    CK(cudaMalloc(&d, 64)); CK(cudaMalloc(&sink, 4));
    run<0>("IADD", d, sink); run<1>("VIADDMNMX.RELU", d, sink); run<2>("VIMNMX3", d, sink); run<9>("VIMNMX(2in)", d, sink);
    run<3>("PRMT", d, sink); run<4>("IMAD(mul.lo)", d, sink); run<10>("IMAD(mad.lo)", d, sink); run<8>("IMAD.HI", d, sink);
    run<5>("SHF.R", d, sink); run<13>("SHF funnel", d, sink); run<6>("LOP3", d, sink); run<7>("I2IP.sat", d, sink); run<14>("BFE", d, sink);
    run<11>("VIMNMX3.16x2", d, sink); run<12>("VIADDMNMX.16x2", d, sink); run<15>("IDP.4A", d, sink);
    run<20>("IDP.4A.U8.S8", d, sink); run<16>("IDP.2A.LO", d, sink); run<17>("IDP.2A.HI", d, sink); run<18>("IDP.2A.LO.U16", d, sink);
    run<19>("I2IP.U16.SAT", d, sink); run<21>("VIMNMX3.RELU", d, sink); run<22>("FFMA", d, sink);
    runmix<16, 5>("IDP.2A + SHF.R", d, sink); runmix<16, 3>("IDP.2A + PRMT", d, sink); runmix<16, 19>("IDP.2A + I2IP.U16", d, sink);
    runmix<15, 3>("IDP.4A + PRMT", d, sink); runmix<10, 3>("IMAD + PRMT", d, sink); runmix<16, 10>("IDP.2A + IMAD", d, sink);
    runmix<10, 5>("IMAD + SHF.R", d, sink); runmix<16, 9>("IDP.2A + VIMNMX", d, sink);
    return 0;
}
