// Which warps of a CTA share a scheduler sub-partition, and how fast one warp alone can issue FFMA2 / LDS.64
// (development aid for the warp-role layout of k_hv).  One CTA, 12 warps; the warps in `mask` run the loop.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
template <int op>
__global__ void __launch_bounds__(384) k(unsigned mask, int iters, u64* cyc, float* out)
{
    __shared__ float sm[4096];
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 4096; i += 384) sm[i] = i;
    __syncthreads();
    if (!((mask >> warp) & 1)) return;
    u64 p[8];
    for (int i = 0; i < 8; i++) p[i] = 0x3f8000003f800000ull + i;
    const u64 pa = 0x3f8000a83f8000a8ull, pb = 0x3f0000003f000000ull;
    const unsigned sa = (unsigned)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 8;
    u64 t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (op == 0) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));
                if (op == 1) { u64 v; asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(sa + (i * 4 + r) * 256)); p[i] ^= v; }
                if (op == 2) { float lo = __uint_as_float((unsigned)p[i]); lo = fmaf(lo, 1.0001f, 0.5f); p[i] = (p[i] & 0xffffffff00000000ull) | __float_as_uint(lo); }
            }
    }
    u64 t1 = clock64();
    if ((threadIdx.x & 31) == 0) cyc[warp] = t1 - t0;
    float s = 0;
    for (int i = 0; i < 8; i++) s += (float)(p[i] & 0xffff);
    out[threadIdx.x] = s;
}
int main()
{
    u64* cyc; float* out;
    cudaMallocManaged(&cyc, 12 * 8); cudaMalloc(&out, 384 * 4);
    const int iters = 4000;
    unsigned masks[] = {1u, 3u, 0x11u, 0x5u, 0x9u, 0x101u, 0x111u, 0xfu, 0x888u, 0xfffu};
    for (int op = 0; op < 3; op++)
        for (unsigned m : masks) {
            for (int i = 0; i < 12; i++) cyc[i] = 0;
            if (op == 0) k<0><<<1, 384>>>(m, iters, cyc, out);
            if (op == 1) k<1><<<1, 384>>>(m, iters, cyc, out);
            if (op == 2) k<2><<<1, 384>>>(m, iters, cyc, out);
            cudaDeviceSynchronize();
            u64 mx = 0; for (int i = 0; i < 12; i++) mx = cyc[i] > mx ? cyc[i] : mx;
            printf("op %d (%s) mask %03x: %.2f cycles per instruction per warp\n", op, op == 0 ? "FFMA2" : op == 1 ? "LDS.64" : "FFMA", m, (double)mx / (iters * 32.0));
        }
    return 0;
}
