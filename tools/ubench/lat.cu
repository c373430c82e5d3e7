// Dependent-issue latency of the f32 / packed f32x2 instructions the k_hv warps chain (development aid, sm_100a).
// One warp per SM sub-partition slot under test; ILP independent chains per thread; prints cycles per instruction of a chain.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
template <int OP, int ILP>
__global__ void k(float* out, long long* cyc, int iters, float a0, float b0)
{
    float x[ILP];
    u64 p[ILP];
    for (int i = 0; i < ILP; i++) { x[i] = a0 + threadIdx.x * 1e-3f + i; p[i] = ((u64)__float_as_uint(x[i]) << 32) | __float_as_uint(x[i] + 1.0f); }
    const u64 pa = ((u64)__float_as_uint(a0) << 32) | __float_as_uint(a0), pb = ((u64)__float_as_uint(b0) << 32) | __float_as_uint(b0);
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 16; r++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) {
                if (OP == 0) x[i] = fmaf(x[i], a0, b0);
                if (OP == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));
                if (OP == 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pa));
                if (OP == 3) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pa));
                if (OP == 4) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
                if (OP == 5) x[i] = __fadd_rn(x[i], a0);
                if (OP == 6) { float lo, hi; asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i])); lo = fmaxf(lo, 0.f); hi = fmaxf(hi, 0.f);
                               asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(lo), "f"(hi)); asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pa)); }
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < ILP; i++) s += x[i] + (float)(p[i] & 0xffff);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP, int ILP>
void run(const char* name, int warps)
{
    float* out; long long* cyc; cudaMalloc(&out, 1024 * 4 * 148); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    k<OP, ILP><<<1, 32 * warps>>>(out, cyc, 10, 1.0001f, 0.5f);
    k<OP, ILP><<<1, 32 * warps>>>(out, cyc, iters, 1.0001f, 0.5f);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s ILP %d warps %2d: %6.2f cycles per chain step (%.2f per instruction issued by the warp)\n", name, ILP, warps, (double)c / (iters * 16.0), (double)c / (iters * 16.0 * ILP));
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    run<0, 1>("FFMA", 1); run<0, 2>("FFMA", 1); run<0, 4>("FFMA", 1);
    run<1, 1>("FFMA2", 1); run<1, 2>("FFMA2", 1); run<1, 3>("FFMA2", 1); run<1, 4>("FFMA2", 1); run<1, 8>("FFMA2", 1);
    run<2, 1>("FADD2", 1); run<3, 1>("FMUL2", 1); run<5, 1>("FADD", 1);
    run<4, 1>("MUFU.RCP", 1); run<4, 4>("MUFU.RCP", 1);
    run<6, 1>("unpack+FMNMX x2+pack+FADD2", 1);
    // several warps on one SM (4 warps = one per sub-partition, 8 = two per sub-partition, ...)
    run<1, 1>("FFMA2", 4); run<1, 1>("FFMA2", 8); run<1, 1>("FFMA2", 12); run<1, 2>("FFMA2", 8); run<1, 4>("FFMA2", 8);
    return 0;
}
