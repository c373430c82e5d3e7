// Issue-rate microbenchmark for the instruction mix of the SSIMULACRA2 kernels on sm_100a (development aid).
// Prints warp-instructions per cycle per SM for each op; 8 independent chains per thread.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define CH 8
template <int OP>
__global__ void __launch_bounds__(512) k(float* out, int iters, float a0, float b0)
{
    float x[2 * CH];
    double d[CH];
    u64 p[CH];
    for (int i = 0; i < 2 * CH; i++) x[i] = a0 + threadIdx.x * 1e-3f + i;
    for (int i = 0; i < CH; i++) { d[i] = (double)x[i]; p[i] = ((u64)__float_as_uint(x[2 * i]) << 32) | __float_as_uint(x[2 * i + 1]); }
    const float a = a0, b = b0;
    u64 pa = ((u64)__float_as_uint(a) << 32) | __float_as_uint(a), pb = ((u64)__float_as_uint(b) << 32) | __float_as_uint(b);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < CH; i++) {
                if (OP == 0) x[i] = fmaf(x[i], a, b);                                 // FFMA 8 chains
                if (OP == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));
                if (OP == 2) x[i] = __fadd_rn(x[i], a);
                if (OP == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pa));
                if (OP == 4) x[i] = __fmul_rn(x[i], a);
                if (OP == 5) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pa));
                if (OP == 6) d[i] = fma(d[i], (double)a, (double)b);
                if (OP == 7) { x[i] = fmaf(x[i], x[i + CH], b); }                      // FFMA reg-reg-imm? (3 regs distinct)
                if (OP == 8) { x[i] = fmaf(x[i], x[i + CH], x[(i + 1) % CH + CH]); }  // 3 distinct regs
                if (OP == 9) { x[i] = fmaf(x[i], a, b); x[i + CH] = __fadd_rn(x[i + CH], a); }   // 2 per
                if (OP == 10) { x[i] = fmaf(x[i], a, b); x[i + CH] = fmaxf(x[i + CH], a + i); }  // fma + alu
                if (OP == 11) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
                if (OP == 12) d[i] = (double)(float)d[i] + 1.0;                         // F2F pair + DADD
                if (OP == 13) asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(d[i]));
                if (OP == 14) asm volatile("{ .reg .b64 c; mov.b64 c, {0f3F8000A8, 0f3F8000A8}; fma.rn.f32x2 %0, %0, c, %1; }" : "+l"(p[i]) : "l"(pb));
                if (OP == 15) asm volatile("{ .reg .b64 c, e; mov.b64 c, {0f3F8000A8, 0f3F8000A8}; mov.b64 e, {0f3F000000, 0f3F000000}; fma.rn.f32x2 %0, %0, c, e; }" : "+l"(p[i]));
                if (OP == 16) asm volatile("{ .reg .b64 c; mov.b64 c, {0f3F8000A8, 0f3F8000A8}; mul.rn.f32x2 %0, %0, c; }" : "+l"(p[i]));
                if (OP == 17) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb)); x[i] = fmaxf(x[i], x[i + CH]) ; x[i+CH] = fminf(x[i + CH], b + i);}
            }
        }
    }
    float s = 0;
    for (int i = 0; i < 2 * CH; i++) s += x[i];
    for (int i = 0; i < CH; i++) s += (float)d[i] + (float)(p[i] & 0xffff);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP>
void run(const char* name, int per_iter)
{
    float* out; cudaMalloc(&out, 148 * 4 * 512 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 20000;
    k<OP><<<148 * 2, 512>>>(out, 100, 1.0001f, 0.5f);
    cudaEventRecord(e0);
    k<OP><<<148 * 2, 512>>>(out, iters, 1.0001f, 0.5f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double warp_inst = (double)148 * 2 * 16 * iters * 4 * CH * per_iter;
    printf("%-28s %8.3f ms  %6.2f warp-inst/ns/GPU  = %5.2f warp-inst/clk/SM at %d MHz nominal\n", name, ms, warp_inst / (ms * 1e6),
           warp_inst / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1000);
    cudaFree(out);
}
int main()
{
    run<0>("FFMA r,imm? (x*a+b)", 1);
    run<1>("FFMA2", 1);
    run<2>("FADD", 1);
    run<3>("FADD2", 1);
    run<4>("FMUL", 1);
    run<5>("FMUL2", 1);
    run<6>("DFMA", 1);
    run<7>("FFMA x*y+b", 1);
    run<8>("FFMA x*y+z", 1);
    run<9>("FFMA+FADD", 2);
    run<10>("FFMA+FMNMX", 2);
    run<11>("MUFU.RCP", 1);
    run<12>("F2F.F32.F64+F2F.F64.F32+DADD", 3);
    run<13>("rcp.f64 seq", 1);
    run<14>("FFMA2 r,imm,r", 1);
    run<15>("FFMA2 r,imm,imm", 1);
    run<16>("FMUL2 r,imm", 1);
    run<17>("FFMA2 + 2 FMNMX", 3);
    return 0;
}
