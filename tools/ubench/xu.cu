// Throughput of the conversion / special-function instructions the front-end leans on, per SM per clock.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o xu xu.cu && ./xu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 4096, ILP = 8;
template <int OP>
__global__ void k(float* out, float seed, long long* cyc)
{
    float f[ILP];
    double d[ILP];
    for (int i = 0; i < ILP; i++) { f[i] = seed + i + threadIdx.x; d[i] = f[i]; }
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (OP == 0) { asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d[i]) : "f"(f[i])); asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f[i]) : "d"(d[i])); }
            if (OP == 1) { asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d[i]) : "f"(f[i])); f[i] = __double_as_longlong(d[i]) >> 40; }
            if (OP == 2) { asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f[i]) : "d"(d[i])); d[i] = __longlong_as_double(((long long)__float_as_int(f[i]) << 29) | 0x3ff0000000000000ll); }
            if (OP == 3) { asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(d[i]) : "d"(d[i])); }
            if (OP == 4) { asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[i]) : "d"(1.0000001)); }
            if (OP == 5) { asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(f[i])); }
            if (OP == 6) { int v; asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(v) : "f"(f[i])); asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f[i]) : "r"(v)); }
            if (OP == 7) { int v = __float_as_int(f[i]) & 0xff; asm volatile("cvt.rn.f64.s32 %0, %1;" : "=d"(d[i]) : "r"(v)); f[i] = __double_as_longlong(d[i]) >> 40; }
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < ILP; i++) s += f[i] + (float)d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main()
{
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMallocManaged(&cyc, 8);
    const char* names[] = {"cvt f32->f64 + cvt f64->f32 (2 instr)", "cvt f32->f64", "cvt f64->f32", "rcp.approx.f64 (MUFU.RCP64H)", "fma.f64", "rcp.approx.f32 (MUFU)", "cvt f32->s32 + s32->f32 (2 instr)", "cvt s32->f64"};
    const int per[] = {2, 1, 1, 1, 1, 1, 2, 1};
    for (int op = 0; op < 8; op++) {
        for (int warps : {4, 8, 16, 32}) {
            auto run = [&](auto kern) { kern<<<148, warps * 32>>>(out, 1.5f, cyc); cudaDeviceSynchronize(); };
            switch (op) { case 0: run(k<0>); break; case 1: run(k<1>); break; case 2: run(k<2>); break; case 3: run(k<3>); break; case 4: run(k<4>); break; case 5: run(k<5>); break; case 6: run(k<6>); break; case 7: run(k<7>); break; }
            double inst = (double)ITERS * ILP * per[op] * warps;   // warp-instructions per SM
            printf("%-42s warps/SM %2d: %.3f warp-instr/clk/SM  (%.1f lanes/clk/SM)\n", names[op], warps, inst / *cyc, 32 * inst / *cyc);
        }
    }
    return 0;
}
