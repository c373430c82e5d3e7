// FP64-pipe questions behind the colour front-end: rates of DMUL / DADD / DFMA, the effect of a
// constant-bank operand, whether F2F.F32.F64 / MUFU.RCP64H share the pipe with DFMA, and dependent-issue latencies.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64 fp64.cu && ./fp64
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double kC[4] = {1.0000001, 0.9999999, 1.25, 0.75};
constexpr int ITERS = 2048;
template <int OP, int ILP>
__global__ void k(float* out, float seed, long long* cyc)
{
    double d[ILP];
    float f[ILP];
    for (int i = 0; i < ILP; i++) { d[i] = seed + i + threadIdx.x * 1e-3; f[i] = (float)d[i]; }
    const double c0 = kC[0] * seed, c1 = kC[1] * seed;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (OP == 0) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(c0), "d"(c1));
            if (OP == 1) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(c0));
            if (OP == 2) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(c0));
            if (OP == 3) d[i] = fma(d[i], kC[2], kC[3]);                       // constant-bank operands
            if (OP == 4) { asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(c0), "d"(c1)); asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f[i]) : "d"(d[i])); }
            if (OP == 5) { asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(c0), "d"(c1)); asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(d[i]) : "d"(d[i])); }
            if (OP == 6) { asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(c0), "d"(c1)); asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[i]) : "f"(seed)); asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[i]) : "f"(seed)); asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[i]) : "f"(seed)); }
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < ILP; i++) s += f[i] + (float)d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP, int ILP>
void run(const char* name, int per, float* out, long long* cyc)
{
    for (int warps : {1, 4, 8, 16}) {
        k<OP, ILP><<<148, warps * 32>>>(out, 1.0f, cyc);
        cudaDeviceSynchronize();
        double inst = (double)ITERS * ILP * per * warps;
        printf("%-44s ILP %d warps/SM %2d: %.3f warp-instr/clk/SM, %.1f clk per instr per warp\n", name, ILP, warps, inst / *cyc,
               (double)*cyc / (ITERS * ILP * per));
    }
}
int main()
{
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMallocManaged(&cyc, 8);
    run<0, 1>("DFMA (latency)", 1, out, cyc);
    run<1, 1>("DMUL (latency)", 1, out, cyc);
    run<2, 1>("DADD (latency)", 1, out, cyc);
    run<0, 8>("DFMA", 1, out, cyc);
    run<1, 8>("DMUL", 1, out, cyc);
    run<2, 8>("DADD", 1, out, cyc);
    run<3, 8>("DFMA const-bank operands", 1, out, cyc);
    run<4, 8>("DFMA + F2F.F32.F64 (2 instr)", 2, out, cyc);
    run<5, 4>("DFMA + MUFU.RCP64H (2 instr)", 2, out, cyc);
    run<6, 8>("DFMA + 3 FFMA (4 instr)", 4, out, cyc);
    return 0;
}
