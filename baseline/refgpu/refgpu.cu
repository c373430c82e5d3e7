// refgpu: the REFERENCE'S GPU DESIGN for SSIMULACRA2, restated in CUDA C++ so that it can be timed on the same B200
// next to the product.  MEASUREMENT TOOLING ONLY: nothing in turbo_metrics_b200/ links,
// imports or calls this; it is built by `make -C baseline/refgpu` and driven by bench.py (--impl refgpu / the
// "gpu_reference_design" block) and tests/test_refgpu.py.
//
// What is restated (from the description in SURVEY.md 2.2 / 3.3, not from the reference's source text):
//   * images are packed f32 C3 with the pitch nppiMalloc_32f_C3 gives them        (ssimulacra2-cuda/src/lib.rs:48-107)
//   * K1  NV12 / P016 limited-range BT.709 -> linear RGB, one thread per 2x2 quad, 16x8 blocks, hardware powf
//                                                                                  (cuda-colorspace-kernel/src/biplanar.rs:7-70)
//   * K3  2x2 box downscale of linear RGB, one thread per output pixel, 32x8      (ssimulacra2-cuda-kernel/src/downscale.rs:4-35)
//   * K4  linear RGB -> XYB per pixel, libdevice cbrtf                             (xyb.rs:3-102)
//   * N1  nppiMul_32f_C3R x3 (ref^2, dis^2, ref*dis)                               (lib.rs:300-317)
//   * K5  ONE VERTICAL recursive-Gaussian pass over 5 images per launch, one thread per sample column of the packed row
//         (3W columns), 96-thread blocks, the last 11 inputs in a shared-memory ring (blur.rs:33-137)
//   * N2  nppiTranspose_32f_C3R x7 so that the second pass is "vertical" again      (lib.rs:342-361, 383-390)
//   * K6  error maps per sample in f32                                             (error_maps.rs:4-60)
//   * N3/N4 nppiSum / nppiSqr / nppiSqr_I per map, 24-byte D2H copies              (lib.rs:417-447)
//   * one CUDA graph per instance, recorded with stream fork/join, launched per pair, host sync per pair, the 108 sums
//     post-processed on the host                                                   (lib.rs:140-229, 271-291, 449-623)
// Arithmetic follows the reference's GPU path (vertical pass first, f32 tails, x^4 by two f32 squarings, hardware powf),
// which is NOT the CPU oracle's: scores agree with the product to a few hundredths, not to the parity bar.
#include <cuda_runtime.h>
#include <npp.h>
#include <nppi.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <vector>

namespace {

constexpr int kScales = 6;
const double kWeight[108] = {
#include "weights.inc"
};

#define CK(x)                                                                                     \
    do {                                                                                          \
        cudaError_t e_ = (x);                                                                     \
        if (e_ != cudaSuccess) {                                                                  \
            fprintf(stderr, "refgpu: %s failed: %s (%s:%d)\n", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return -1;                                                                            \
        }                                                                                         \
    } while (0)
#define NK(x)                                                                                     \
    do {                                                                                          \
        NppStatus s_ = (x);                                                                       \
        if (s_ != NPP_SUCCESS) {                                                                  \
            fprintf(stderr, "refgpu: %s failed: %d (%s:%d)\n", #x, (int)s_, __FILE__, __LINE__);  \
            return -1;                                                                            \
        }                                                                                         \
    } while (0)

struct Img {
    float* p = nullptr;
    int step = 0;   // bytes
    int w = 0, h = 0;
};

// ---------------- K1: biplanar YUV 4:2:0 -> linear RGB (packed f32) ----------------
__device__ __forceinline__ float bt709_eotf_fast(float v)
{
    return v < 0.081f ? v * (1.0f / 4.5f) : __powf((v + 0.099f) * (1.0f / 1.099f), 1.0f / 0.45f);
}
template <typename T, int K>   // K = container bits - 8
__global__ void k_yuv420_to_linear(const uint8_t* __restrict__ src, int pitch, int coded_h, float* __restrict__ dst, int dstep, int w,
                                   int h, float rc, float g1, float g2, float bc)
{
    const int qx = blockIdx.x * blockDim.x + threadIdx.x, qy = blockIdx.y * blockDim.y + threadIdx.y;
    if (2 * qx >= w || 2 * qy >= h) return;
    const T* uvrow = reinterpret_cast<const T*>(src + (size_t)pitch * coded_h + (size_t)qy * pitch);
    const int cbi = uvrow[2 * qx], cri = uvrow[2 * qx + 1];
    const float cb = (float)(cbi - (128 << K)) * (1.0f / (float)(224 << K));
    const float cr = (float)(cri - (128 << K)) * (1.0f / (float)(224 << K));
    for (int dy = 0; dy < 2; dy++) {
        const int y = 2 * qy + dy;
        if (y >= h) break;
        const T* yrow = reinterpret_cast<const T*>(src + (size_t)y * pitch);
        float* drow = reinterpret_cast<float*>(reinterpret_cast<char*>(dst) + (size_t)y * dstep);
        for (int dx = 0; dx < 2; dx++) {
            const int x = 2 * qx + dx;
            if (x >= w) break;
            const int yi = max((int)yrow[x], 16 << K) - (16 << K);
            const float luma = (float)yi * (1.0f / (float)(219 << K));
            const float r = luma + rc * cr, g = luma + g1 * cb + g2 * cr, b = luma + bc * cb;
            drow[3 * x + 0] = fminf(fmaxf(bt709_eotf_fast(r), 0.0f), 1.0f);
            drow[3 * x + 1] = fminf(fmaxf(bt709_eotf_fast(g), 0.0f), 1.0f);
            drow[3 * x + 2] = fminf(fmaxf(bt709_eotf_fast(b), 0.0f), 1.0f);
        }
    }
}

// ---------------- K3: 2x2 box downscale ----------------
__global__ void k_downscale(const float* __restrict__ src, int sstep, int sw, int sh, float* __restrict__ dst, int dstep, int dw, int dh)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh) return;
    float* d = reinterpret_cast<float*>(reinterpret_cast<char*>(dst) + (size_t)y * dstep) + 3 * x;
    for (int c = 0; c < 3; c++) {
        float acc = 0.0f;
        for (int iy = 0; iy < 2; iy++)
            for (int ix = 0; ix < 2; ix++) {
                const int sx = min(2 * x + ix, sw - 1), sy = min(2 * y + iy, sh - 1);
                acc += reinterpret_cast<const float*>(reinterpret_cast<const char*>(src) + (size_t)sy * sstep)[3 * sx + c];
            }
        d[c] = acc * 0.25f;
    }
}

// ---------------- K4: linear RGB -> XYB ----------------
__global__ void k_xyb(const float* __restrict__ src, int sstep, float* __restrict__ dst, int dstep, int w, int h)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const float* s = reinterpret_cast<const float*>(reinterpret_cast<const char*>(src) + (size_t)y * sstep) + 3 * x;
    float* d = reinterpret_cast<float*>(reinterpret_cast<char*>(dst) + (size_t)y * dstep) + 3 * x;
    const float r = s[0], g = s[1], b = s[2];
    const float bias = 0.0037930734f, broot = 0.15595420054f;
    float l = fmaf(0.30f, r, fmaf(0.622f, g, fmaf(0.078f, b, bias)));
    float m = fmaf(0.23f, r, fmaf(0.692f, g, fmaf(0.078f, b, bias)));
    float t = fmaf(0.24342269f, r, fmaf(0.20476745f, g, fmaf(0.55180986f, b, bias)));
    l = cbrtf(fmaxf(l, 0.0f)) - broot;
    m = cbrtf(fmaxf(m, 0.0f)) - broot;
    t = cbrtf(fmaxf(t, 0.0f)) - broot;
    const float X = 0.5f * (l - m), Y = 0.5f * (l + m);
    d[0] = fmaf(X, 14.0f, 0.42f);
    d[1] = Y + 0.01f;
    d[2] = (t - Y) + 0.55f;
}

// ---------------- K5: one vertical recursive-Gaussian pass over 5 images ----------------
struct Blur5 {
    const float* src[5];
    float* dst[5];
    int sstep, dstep;
};
constexpr int kBlurThreads = 96, kRing = 11;
__global__ void __launch_bounds__(kBlurThreads) k_blur_vertical5(Blur5 a, int cols /* 3 * w */, int h)
{
    __shared__ float ring[kRing][kBlurThreads];
    const int col = blockIdx.x * kBlurThreads + threadIdx.x;
    if (col >= cols) return;
    const float* src = a.src[blockIdx.y];
    float* dst = a.dst[blockIdx.y];
    const float in1 = 0.055295236f, in3 = -0.058836687f, in5 = 0.012955819f;
    const float pv1 = 1.9021131f, pv3 = 1.1755705f, pv5 = 1.2246469e-16f;
    float p1 = 0, p3 = 0, p5 = 0, q1 = 0, q3 = 0, q5 = 0;
    for (int i = 0; i < kRing; i++) ring[i][threadIdx.x] = 0.0f;
    // preload x[0..3] (the taps x[n+4] of n = -4..-1 arrive inside the loop)
    for (int n = -4; n < h; n++) {
        const int r = n + 4;                                    // newest row
        const float right = r < h ? reinterpret_cast<const float*>(reinterpret_cast<const char*>(src) + (size_t)r * a.sstep)[col] : 0.0f;
        const int l = n - 6;                                    // oldest row still needed
        const float left = l >= 0 ? ring[l % kRing][threadIdx.x] : 0.0f;
        if (r < h) ring[r % kRing][threadIdx.x] = right;
        const float sum = left + right;
        const float o1 = fmaf(pv1, p1, fmaf(-1.0f, q1, sum * in1));
        const float o3 = fmaf(pv3, p3, fmaf(-1.0f, q3, sum * in3));
        const float o5 = fmaf(pv5, p5, fmaf(-1.0f, q5, sum * in5));
        q1 = p1; q3 = p3; q5 = p5;
        p1 = o1; p3 = o3; p5 = o5;
        if (n >= 0) reinterpret_cast<float*>(reinterpret_cast<char*>(dst) + (size_t)n * a.dstep)[col] = o1 + o3 + o5;
    }
}

// ---------------- K6: error maps ----------------
struct ErrArgs {
    const float *ref, *dis, *mu1, *mu2, *s11, *s22, *s12;
    float *ssim, *art, *det;
    int step;
};
__global__ void k_error_maps(ErrArgs a, int cols, int h)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cols || y >= h) return;
    const size_t o = (size_t)y * a.step / 4 + x;
    const float mu1 = a.mu1[o], mu2 = a.mu2[o];
    const float mu11 = mu1 * mu1, mu22 = mu2 * mu2, mu12 = mu1 * mu2;
    const float md = mu1 - mu2;
    const float num_m = fmaf(md, -md, 1.0f);
    const float num_s = fmaf(2.0f, a.s12[o] - mu12, 0.0009f);
    const float den_s = (a.s11[o] - mu11) + (a.s22[o] - mu22) + 0.0009f;
    a.ssim[o] = fmaxf(1.0f - (num_m * num_s) / den_s, 0.0f);
    const float d1 = (1.0f + fabsf(a.dis[o] - mu2)) / (1.0f + fabsf(a.ref[o] - mu1)) - 1.0f;
    a.art[o] = fmaxf(d1, 0.0f);
    a.det[o] = fmaxf(-d1, 0.0f);
}

dim3 grid2d(int w, int h, dim3 b) { return dim3((w + b.x - 1) / b.x, (h + b.y - 1) / b.y); }

}  // namespace

struct refgpu_t {
    int w = 0, h = 0, bits = 8;
    int sw[kScales], sh[kScales];
    Img lin[kScales][2];                 // linear RGB pyramid (scale 0 = converted input)
    Img img[kScales][10], imgt[kScales][10];
    std::vector<cudaStream_t> streams;   // 0 = main, 1 = alt, 2.. = forks, then 36 reduction streams
    std::vector<cudaEvent_t> events;
    size_t ev_next = 0;
    cudaEvent_t join_ev = nullptr;       // converts on stream 1 -> graph launch on stream 0
    Npp8u* sum_scratch[kScales * 6] = {};
    Npp64f* dsums = nullptr;             // [scale][map][L1,L4][channel]
    double* hsums = nullptr;             // pinned
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    size_t graph_nodes = 0, kernel_nodes = 0;
    size_t bytes = 0;
    NppStreamContext ctx0;
};

namespace {

int alloc_img(refgpu_t* r, Img& im, int w, int h)
{
    im.w = w; im.h = h;
    im.p = nppiMalloc_32f_C3(w, h, &im.step);
    if (!im.p) { fprintf(stderr, "refgpu: nppiMalloc_32f_C3(%d,%d) failed\n", w, h); return -1; }
    r->bytes += (size_t)im.step * h;
    return 0;
}
NppStreamContext ctx_on(const refgpu_t* r, cudaStream_t s)
{
    NppStreamContext c = r->ctx0;
    c.hStream = s;
    return c;
}
cudaEvent_t next_event(refgpu_t* r)
{
    if (r->ev_next == r->events.size()) {
        cudaEvent_t e;
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        r->events.push_back(e);
    }
    return r->events[r->ev_next++];
}
void wait_for(refgpu_t* r, cudaStream_t waiter, cudaStream_t on)
{
    cudaEvent_t e = next_event(r);
    cudaEventRecord(e, on);
    cudaStreamWaitEvent(waiter, e, 0);
}

int record(refgpu_t* r)
{
    cudaStream_t main_s = r->streams[0], alt = r->streams[1];
    cudaStream_t* fork = &r->streams[2];        // 5 fork streams
    cudaStream_t* red = &r->streams[7];         // 36 reduction streams
    CK(cudaStreamBeginCapture(main_s, cudaStreamCaptureModeRelaxed));
    const dim3 b328(32, 8);
    for (int s = 0; s < kScales; s++) {
        const int w = r->sw[s], h = r->sh[s];
        Img* im = r->img[s];
        Img* it = r->imgt[s];
        wait_for(r, alt, main_s);
        if (s >= 1) {
            k_downscale<<<grid2d(w, h, b328), b328, 0, main_s>>>(r->lin[s - 1][0].p, r->lin[s - 1][0].step, r->sw[s - 1], r->sh[s - 1],
                                                                 r->lin[s][0].p, r->lin[s][0].step, w, h);
            k_downscale<<<grid2d(w, h, b328), b328, 0, alt>>>(r->lin[s - 1][1].p, r->lin[s - 1][1].step, r->sw[s - 1], r->sh[s - 1],
                                                              r->lin[s][1].p, r->lin[s][1].step, w, h);
        }
        k_xyb<<<grid2d(w, h, b328), b328, 0, main_s>>>(r->lin[s][0].p, r->lin[s][0].step, im[8].p, im[8].step, w, h);
        k_xyb<<<grid2d(w, h, b328), b328, 0, alt>>>(r->lin[s][1].p, r->lin[s][1].step, im[9].p, im[9].step, w, h);
        wait_for(r, main_s, alt);
        // products on three streams
        const NppiSize roi = {w, h}, roit = {h, w};
        const int ma[3] = {8, 9, 8}, mb[3] = {8, 9, 9};
        for (int i = 0; i < 3; i++) {
            wait_for(r, fork[i], main_s);
            NK(nppiMul_32f_C3R_Ctx(im[ma[i]].p, im[ma[i]].step, im[mb[i]].p, im[mb[i]].step, im[i].p, im[i].step, roi, ctx_on(r, fork[i])));
        }
        for (int i = 0; i < 3; i++) wait_for(r, main_s, fork[i]);
        // first (vertical) pass: 0->3, 1->4, 2->5, 8->6, 9->7
        {
            Blur5 a;
            const int srcs[5] = {0, 1, 2, 8, 9}, dsts[5] = {3, 4, 5, 6, 7};
            for (int i = 0; i < 5; i++) { a.src[i] = im[srcs[i]].p; a.dst[i] = im[dsts[i]].p; }
            a.sstep = im[0].step; a.dstep = im[3].step;
            k_blur_vertical5<<<dim3((3 * w + kBlurThreads - 1) / kBlurThreads, 5), kBlurThreads, 0, main_s>>>(a, 3 * w, h);
        }
        // transposes on five streams: img 3..7 -> imgt 0..4
        for (int i = 0; i < 5; i++) {
            wait_for(r, fork[i], main_s);
            NK(nppiTranspose_32f_C3R_Ctx(im[3 + i].p, im[3 + i].step, it[i].p, it[i].step, roi, ctx_on(r, fork[i])));
        }
        for (int i = 0; i < 5; i++) wait_for(r, main_s, fork[i]);
        // second pass in the transposed domain: imgt 0..4 -> 5..9 (s11, s22, s12, mu1, mu2)
        {
            Blur5 a;
            for (int i = 0; i < 5; i++) { a.src[i] = it[i].p; a.dst[i] = it[5 + i].p; }
            a.sstep = it[0].step; a.dstep = it[5].step;
            k_blur_vertical5<<<dim3((3 * h + kBlurThreads - 1) / kBlurThreads, 5), kBlurThreads, 0, main_s>>>(a, 3 * h, w);
        }
        // the XYB images themselves into the transposed domain
        for (int i = 0; i < 2; i++) {
            wait_for(r, fork[i], main_s);
            NK(nppiTranspose_32f_C3R_Ctx(im[8 + i].p, im[8 + i].step, it[i].p, it[i].step, roi, ctx_on(r, fork[i])));
        }
        for (int i = 0; i < 2; i++) wait_for(r, main_s, fork[i]);
        {
            ErrArgs a{it[0].p, it[1].p, it[8].p, it[9].p, it[5].p, it[6].p, it[7].p, it[2].p, it[3].p, it[4].p, it[0].step};
            k_error_maps<<<grid2d(3 * h, w, b328), b328, 0, main_s>>>(a, 3 * h, w);
        }
        // reductions: per map one stream for the plain sum, one for the 4th-power sum
        for (int m = 0; m < 3; m++) {
            cudaStream_t s1 = red[s * 6 + m * 2], s4 = red[s * 6 + m * 2 + 1];
            Npp64f* d1 = r->dsums + ((s * 3 + m) * 2 + 0) * 3;
            Npp64f* d4 = r->dsums + ((s * 3 + m) * 2 + 1) * 3;
            wait_for(r, s1, main_s);
            wait_for(r, s4, main_s);
            NK(nppiSum_32f_C3R_Ctx(it[2 + m].p, it[2 + m].step, roit, r->sum_scratch[s * 6 + m * 2], d1, ctx_on(r, s1)));
            CK(cudaMemcpyAsync(r->hsums + ((s * 3 + m) * 2 + 0) * 3, d1, 24, cudaMemcpyDeviceToHost, s1));
            NK(nppiSqr_32f_C3R_Ctx(it[2 + m].p, it[2 + m].step, it[5 + m].p, it[5 + m].step, roit, ctx_on(r, s4)));
            NK(nppiSqr_32f_C3IR_Ctx(it[5 + m].p, it[5 + m].step, roit, ctx_on(r, s4)));
            NK(nppiSum_32f_C3R_Ctx(it[5 + m].p, it[5 + m].step, roit, r->sum_scratch[s * 6 + m * 2 + 1], d4, ctx_on(r, s4)));
            CK(cudaMemcpyAsync(r->hsums + ((s * 3 + m) * 2 + 1) * 3, d4, 24, cudaMemcpyDeviceToHost, s4));
        }
    }
    for (int i = 0; i < 36; i++) wait_for(r, main_s, red[i]);
    wait_for(r, main_s, alt);
    CK(cudaStreamEndCapture(main_s, &r->graph));
    CK(cudaGraphInstantiate(&r->exec, r->graph, 0));
    size_t n = 0;
    CK(cudaGraphGetNodes(r->graph, nullptr, &n));
    std::vector<cudaGraphNode_t> nodes(n);
    CK(cudaGraphGetNodes(r->graph, nodes.data(), &n));
    r->graph_nodes = n;
    for (size_t i = 0; i < n; i++) {
        cudaGraphNodeType t;
        cudaGraphNodeGetType(nodes[i], &t);
        if (t == cudaGraphNodeTypeKernel) r->kernel_nodes++;
    }
    return 0;
}

double post_process(const refgpu_t* r, double* norms108)
{
    double score = 0.0;
    for (int c = 0; c < 3; c++)
        for (int s = 0; s < kScales; s++) {
            const double inv = 1.0 / ((double)r->sw[s] * r->sh[s]);
            for (int n = 0; n < 2; n++)
                for (int m = 0; m < 3; m++) {
                    double v = r->hsums[((s * 3 + m) * 2 + n) * 3 + c] * inv;
                    if (n == 1) v = std::sqrt(std::sqrt(v));
                    const int idx = c * 36 + s * 6 + n * 3 + m;
                    if (norms108) norms108[idx] = v;
                    score += kWeight[idx] * std::fabs(v);
                }
        }
    score *= 0.9562382616834844;
    score = 6.248496625763138e-5 * score * score * score + 2.326765642916932 * score - 0.020884521182843837 * score * score;
    return score > 0.0 ? 100.0 - 10.0 * std::pow(score, 0.6276336467831387) : 100.0;
}

}  // namespace

extern "C" {

int refgpu_create(refgpu_t** out, int w, int h, int bits)
{
    if (!out || w < 8 || h < 8 || (bits != 8 && bits != 16)) return -1;
    refgpu_t* r = new refgpu_t();
    r->w = w; r->h = h; r->bits = bits;
    NK(nppGetStreamContext(&r->ctx0));
    r->streams.resize(2 + 5 + 36);
    for (auto& s : r->streams) CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    int cw = w, ch = h;
    for (int s = 0; s < kScales; s++) {
        r->sw[s] = cw; r->sh[s] = ch;
        for (int i = 0; i < 2; i++)
            if (alloc_img(r, r->lin[s][i], cw, ch)) return -1;
        for (int i = 0; i < 10; i++) {
            if (alloc_img(r, r->img[s][i], cw, ch)) return -1;
            if (alloc_img(r, r->imgt[s][i], ch, cw)) return -1;
        }
        for (int i = 0; i < 6; i++) {
            size_t sz = 0;
            const NppiSize roit = {ch, cw};
            NK(nppiSumGetBufferHostSize_32f_C3R_Ctx(roit, &sz, r->ctx0));
            CK(cudaMalloc(&r->sum_scratch[s * 6 + i], sz));
            r->bytes += sz;
        }
        cw = (cw + 1) / 2; ch = (ch + 1) / 2;
    }
    CK(cudaMalloc(&r->dsums, kScales * 3 * 2 * 3 * sizeof(Npp64f)));
    CK(cudaMallocHost(&r->hsums, kScales * 3 * 2 * 3 * sizeof(double)));
    CK(cudaEventCreateWithFlags(&r->join_ev, cudaEventDisableTiming));
    if (record(r)) return -1;
    *out = r;
    return 0;
}

// One pair, like TurboMetrics::compute_one (turbo-metrics/src/lib.rs:268-360): convert both frames on two streams, join, launch
// the graph, synchronise, post-process on the host.
int refgpu_compute(refgpu_t* r, const void* ref_yuv, const void* dis_yuv, int pitch, int coded_h, double* score, double* norms108)
{
    if (!r || !ref_yuv || !dis_yuv || !score) return -1;
    const double kr = 0.21264, kb = 0.07219, kg = 1.0 - kr - kb;     // BT.709 from its primaries
    const float rc = (float)(2.0 * (1.0 - kr)), bc = (float)(2.0 * (1.0 - kb));
    const float g1 = (float)(-2.0 * kb * (1.0 - kb) / kg), g2 = (float)(-2.0 * kr * (1.0 - kr) / kg);
    const dim3 b(16, 8);
    const dim3 g = grid2d((r->w + 1) / 2, (r->h + 1) / 2, b);
    const void* src[2] = {ref_yuv, dis_yuv};
    for (int i = 0; i < 2; i++) {
        cudaStream_t st = r->streams[i];
        if (r->bits == 8)
            k_yuv420_to_linear<uint8_t, 0><<<g, b, 0, st>>>((const uint8_t*)src[i], pitch, coded_h, r->lin[0][i].p, r->lin[0][i].step, r->w, r->h,
                                                            rc, g1, g2, bc);
        else
            k_yuv420_to_linear<uint16_t, 8><<<g, b, 0, st>>>((const uint8_t*)src[i], pitch, coded_h, r->lin[0][i].p, r->lin[0][i].step, r->w,
                                                             r->h, rc, g1, g2, bc);
    }
    CK(cudaEventRecord(r->join_ev, r->streams[1]));
    CK(cudaStreamWaitEvent(r->streams[0], r->join_ev, 0));
    CK(cudaGraphLaunch(r->exec, r->streams[0]));
    CK(cudaStreamSynchronize(r->streams[0]));
    *score = post_process(r, norms108);
    return 0;
}

int refgpu_info(const refgpu_t* r, size_t* graph_nodes, size_t* kernel_nodes, size_t* bytes)
{
    if (!r) return -1;
    if (graph_nodes) *graph_nodes = r->graph_nodes;
    if (kernel_nodes) *kernel_nodes = r->kernel_nodes;
    if (bytes) *bytes = r->bytes;
    return 0;
}

void refgpu_destroy(refgpu_t* r)
{
    if (!r) return;
    cudaDeviceSynchronize();
    if (r->exec) cudaGraphExecDestroy(r->exec);
    if (r->graph) cudaGraphDestroy(r->graph);
    for (int s = 0; s < kScales; s++) {
        for (int i = 0; i < 2; i++) nppiFree(r->lin[s][i].p);
        for (int i = 0; i < 10; i++) { nppiFree(r->img[s][i].p); nppiFree(r->imgt[s][i].p); }
    }
    for (auto p : r->sum_scratch) cudaFree(p);
    cudaFree(r->dsums);
    cudaFreeHost(r->hsums);
    for (auto e : r->events) cudaEventDestroy(e);
    if (r->join_ev) cudaEventDestroy(r->join_ev);
    for (auto s : r->streams) cudaStreamDestroy(s);
    delete r;
}

}  // extern "C"
