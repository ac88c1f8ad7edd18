// Micro-benchmark: issue rate of the integer instructions the PEE kernels are built from, per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ops ops.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>

#define CHAINS 8
#define UNROLL 16
#define ITERS 256

template <int OP>
__device__ __forceinline__ void op(unsigned& x, unsigned a, unsigned b, unsigned long long& w) {
    if (OP == 0) asm volatile("dp2a.lo.u32.s32 %0, %1, %2, %0;" : "+r"(x) : "r"(a), "r"(b));
    if (OP == 1) asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(x) : "r"(a), "r"(b));
    if (OP == 2) asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(w) : "r"(a), "r"(b));
    if (OP == 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(a), "r"(b));
    if (OP == 4) asm volatile("prmt.b32 %0, %0, %1, 0x3254;" : "+r"(x) : "r"(a));
    if (OP == 5) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));
    if (OP == 6) asm volatile("min.s32 %0, %0, %1;" : "+r"(x) : "r"(a));
    if (OP == 7) asm volatile("min.s32.relu %0, %0, %1;" : "+r"(x) : "r"(a));
    if (OP == 8) asm volatile("popc.b32 %0, %0;" : "+r"(x));
    if (OP == 9) asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; vote.sync.ballot.b32 %0, p, 0xffffffff;}" : "+r"(x) : "r"(a));
    if (OP == 10) asm volatile("shfl.sync.up.b32 %0, %0, 1, 0, 0xffffffff;" : "+r"(x));
    if (OP == 11) asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; selp.b32 %0, %1, %2, p;}" : "+r"(x) : "r"(a), "r"(b));
    if (OP == 12) asm volatile("add.s32 %0, %0, %1;" : "+r"(x) : "r"(a));
    if (OP == 13) {  // pair: one fma-pipe op and one alu-pipe op
        asm volatile("dp2a.lo.u32.s32 %0, %1, %2, %0;" : "+r"(x) : "r"(a), "r"(b));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(a), "r"(b));
    }
    if (OP == 14) {
        asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(x) : "r"(a), "r"(b));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(a), "r"(b));
    }
    if (OP == 15) asm volatile("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(x) : "r"(a), "r"(b));
    if (OP == 16) asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; @p add.s32 %0, %0, %2;}" : "+r"(x) : "r"(a), "r"(b));
    if (OP == 17) asm volatile("mad.lo.s32 %0, %0, 5, %1;" : "+r"(x) : "r"(a));          // imm multiplier
    if (OP == 18) asm volatile("max.s32 %0, %0, %1; min.s32 %0, %0, %2;" : "+r"(x) : "r"(a), "r"(b));
    if (OP == 19) asm volatile("shl.b32 %0, %0, 1; add.s32 %0, %0, %1;" : "+r"(x) : "r"(a));  // LEA?
    if (OP == 20) asm volatile("bfe.u32 %0, %0, 16, 16;" : "+r"(x));
    if (OP == 21) asm volatile("vadd2.u32.u32.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));
    // packed 16-bit pairs (sm_90+): VIADD.16x2, VIMNMX.U16x2 / .S16x2[.RELU]
    if (OP == 22) asm volatile("add.u16x2 %0, %0, %1;" : "+r"(x) : "r"(a));
    if (OP == 23) asm volatile("min.u16x2 %0, %0, %1;" : "+r"(x) : "r"(a));
    if (OP == 24) asm volatile("max.s16x2.relu %0, %0, %1;" : "+r"(x) : "r"(a));
    if (OP == 25) {  // unsigned clamp of two pixels: max then min
        asm volatile("max.u16x2 %0, %0, %1;" : "+r"(x) : "r"(a));
        asm volatile("min.u16x2 %0, %0, %1;" : "+r"(x) : "r"(b));
    }
    if (OP == 26) {  // packed op next to an fma-pipe op
        asm volatile("dp2a.lo.u32.s32 %0, %1, %2, %0;" : "+r"(x) : "r"(a), "r"(b));
        asm volatile("add.u16x2 %0, %0, %1;" : "+r"(x) : "r"(a));
    }
    if (OP == 27) asm volatile("mad.hi.u32 %0, %1, 2, %0;" : "+r"(x) : "r"(a));            // IMAD.HI: x + (a >> 31)
    if (OP == 28) asm volatile("min.s32 %0, %0, %1; max.s32 %0, %0, %2;" : "+r"(x) : "r"(a), "r"(b));  // VIMNMX3?
}

template <int OP>
__global__ void __launch_bounds__(512) k(unsigned* out, unsigned a, unsigned b, long long* cyc) {
    unsigned x[CHAINS];
    unsigned long long w[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) { x[c] = threadIdx.x * 7 + c; w[c] = x[c]; }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) op<OP>(x[c], a + c, b, w[c]);
    }
    __syncthreads();
    const long long t1 = clock64();
    unsigned s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c] + (unsigned)w[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int per_op, unsigned* out, long long* cyc, int warps) {
    k<OP><<<148, warps * 32>>>(out, 3, 0x00ff0001u, cyc);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP><<<148, warps * 32>>>(out, 3, 0x00ff0001u, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    const double ninstr = (double)warps * ITERS * UNROLL * CHAINS * per_op;
    printf("%-28s warps/SM %2d : %6.3f warp-instr/clk/SM  (%.3f per SMSP)  %.3f ms\n", name, warps, ninstr / avg, ninstr / avg / 4, ms);
}

int main() {
    unsigned* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    for (int warps : {8, 16}) {
        run<0>("IDP.2A (dp2a.lo.u32.s32)", 1, out, cyc, warps);
        run<15>("IDP.4A (dp4a.u32.s32)", 1, out, cyc, warps);
        run<1>("IMAD", 1, out, cyc, warps);
        run<17>("IMAD imm", 1, out, cyc, warps);
        run<2>("IMAD.WIDE", 1, out, cyc, warps);
        run<3>("LOP3", 1, out, cyc, warps);
        run<4>("PRMT", 1, out, cyc, warps);
        run<5>("SHF", 1, out, cyc, warps);
        run<20>("BFE", 1, out, cyc, warps);
        run<6>("VIMNMX min", 1, out, cyc, warps);
        run<7>("VIMNMX min.relu", 1, out, cyc, warps);
        run<18>("max+min (clamp)", 2, out, cyc, warps);
        run<12>("add.s32", 1, out, cyc, warps);
        run<19>("shl+add (LEA?)", 1, out, cyc, warps);
        run<8>("POPC", 1, out, cyc, warps);
        run<9>("ISETP+VOTE", 2, out, cyc, warps);
        run<10>("SHFL.UP", 1, out, cyc, warps);
        run<11>("ISETP+SEL", 2, out, cyc, warps);
        run<16>("ISETP+@p IADD", 2, out, cyc, warps);
        run<13>("IDP + LOP3 pair", 2, out, cyc, warps);
        run<14>("IMAD + LOP3 pair", 2, out, cyc, warps);
        run<21>("vadd2", 1, out, cyc, warps);
        run<22>("VIADD.16x2 (add.u16x2)", 1, out, cyc, warps);
        run<23>("VIMNMX.U16x2 (min.u16x2)", 1, out, cyc, warps);
        run<24>("VIMNMX.S16x2.RELU", 1, out, cyc, warps);
        run<25>("max.u16x2 + min.u16x2", 2, out, cyc, warps);
        run<26>("IDP + VIADD.16x2 pair", 2, out, cyc, warps);
        run<27>("IMAD.HI.U32", 1, out, cyc, warps);
        run<28>("min.s32 + max.s32", 2, out, cyc, warps);
    }
    return 0;
}
