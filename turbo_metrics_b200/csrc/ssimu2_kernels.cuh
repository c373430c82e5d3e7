// ssimu2_kernels.cuh -- device code of the B200-native SSIMULACRA2 frame-pair scorer (sm_100a).
//
// Pipeline per batch of frame pairs (3 launches, every launch covers all frames and all 6 scales):
//   k_frontend2 : source pair -> linear RGB -> 2x box pyramid (kept in registers) -> XYB planes of all scales
//   k_hv        : per scale: products -> HORIZONTAL recursive Gaussian of {ref^2, dis^2, ref*dis, ref, dis} x 3 channels
//                 -> VERTICAL recursive Gaussian -> SSIM / artifact / detail-loss maps -> L1 / L4 partial sums (f64);
//                 the 15 H-pass planes never leave the SM
//   k_finalize  : partial sums -> 108 norms -> weighted sum -> score (f64)
// The development pipeline "split" runs the two passes as k_hpass + k_vpass with the H-pass planes in HBM (what the
// bit-identity tests read back).
//
// Arithmetic contract: everything that feeds the recursive filters replicates, operation for
// operation, the reference's CPU implementation (crates/ssimulacra2-cuda/examples/cpu.rs), including
// its libm cbrtf / powf (exact_math.cuh), so the filter inputs and outputs are bit-identical to the
// CPU path; compile with -fmad=false, fused ops are explicit fmaf()/fma().  The GPU reference
// kernels each function replaces are cited at the function.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "exact_math.cuh"

#include <type_traits>

namespace ssimu2 {

constexpr int kMaxScales = 6;
constexpr int kMaxBatch = 1024;   // frame pairs per launch group (the frame table lives in device memory)

enum Fmt : int { kNV12 = 0, kP016 = 1, kSRGB8 = 2, kSRGB16 = 3, kSRGBF32 = 4, kLINEARF32 = 5,
                 // internal: P016 whose samples carry more than 10 significant bits (SSIMU2_FLAG_P016_DEEP: 12-bit sources).
                 // Same layout and arithmetic as kP016; no 10-bit memo tables, all three transfer functions are evaluated
                 kP016A = 6 };
template <int FMT>
struct FmtIs {
    static constexpr bool p016 = FMT == kP016 || FMT == kP016A;
    static constexpr bool yuv = p016 || FMT == kNV12;
};

struct ScaleDesc {
    int w, h, pitch;       // pitch in floats (multiple of 32)
    int n_bands;           // H-pass work items (32-row bands)
    int n_strips;          // V-pass work items (column strips)
    int strip0;            // index of this scale's first strip in the partial-sum table
    long long xyb_off;     // float offset of scale s inside a slot's XYB buffer: [2 img][3 ch][h][pitch]
    long long hb_off;      // float offset of scale s inside a slot's H-pass buffer: [15][h][pitch]
    int nb;                // k_hv: 12-row bands, ceil((h + 4) / 12)
    int rec0;              // k_hv: index of this scale's first hand-off record (records: [strip][band])
    int item0;             // k_hv: strips of the larger scales before this one
};

struct YuvCoef {           // MatrixCoefficients::coefficients, cuda-colorspace-kernel/src/lib.rs:183-201
    float y, r, b, g1, g2;
    int luma_min, neutral;
};

struct Geo {
    ScaleDesc sc[kMaxScales];
    int nscales;
    int items_h, items_v;        // work items per frame (all scales)
    int total_strips;            // partial-sum rows per frame
    int total_recs;              // k_hv hand-off records per frame
    long long xyb_stride;        // floats per slot
    long long hb_stride;         // floats per slot
    YuvCoef coef;
    // Exact memo of the R and B transfer for YUV sources: R' = luma(Y) + r*Cr and B' = luma(Y) + b*Cb depend on two
    // integer codes only, so linear R / B are looked up in two N x N tables ([0]: R by (Cr, Y), [1]: B by (Cb, Y),
    // index c * N + y) built at create time BY THE SAME device code (bit-identical by construction).  G' depends on
    // three codes and keeps the arithmetic path.  Codes are sample >> lut_shift (P016: the 10 significant bits);
    // samples with non-zero low bits take the arithmetic path.
    const float* eotf_lut;
    int lut_n, lut_shift;
};

struct FrameIn {
    const uint8_t* p0;
    const uint8_t* p1;
    uint32_t pitch;
    uint32_t pad;
};

struct FramePair {         // one entry of a batch slot's frame table (device memory, uploaded with the batch)
    FrameIn ref, dis;
};

// ------------------------------------------------------------------------------------------
// constants
// ------------------------------------------------------------------------------------------
__device__ const float kSrgb8Lut[256] = {
#include "srgb8_lut.inc"
};

// Recursive Gaussian sigma = 1.5, radius 5 (cpu.rs:931-948; ssimulacra2-cuda-kernel/build.rs:28-145).
#define RG_IN_1 0.055295236f
#define RG_IN_3 (-0.058836687f)
#define RG_IN_5 0.012955819f
#define RG_PREV_1 1.9021131f
#define RG_PREV_3 1.1755705f
#define RG_PREV_5 0.00000000000000012246469f

// ---- packed f32x2 arithmetic (sm_100 FFMA2 / FADD2 / FMUL2) ----------------------------------------
// One instruction = two independent IEEE round-to-nearest f32 operations on a 64-bit register pair, so the
// results are bit-identical to the scalar forms while the filter costs half the issue slots.  Negations are
// written as unpack / negate / pack: ptxas folds them into the source modifiers of the consuming instruction.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
struct f2 {
    unsigned long long v;
};
__device__ __forceinline__ f2 f2_pack(float lo, float hi)
{
    f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void f2_unpack(f2 a, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); }
__device__ __forceinline__ f2 f2_splat(float c) { return f2_pack(c, c); }
__device__ __forceinline__ f2 f2_add(f2 a, f2 b)
{
    f2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ f2 f2_sub(f2 a, f2 b)
{
    f2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ f2 f2_mul(f2 a, f2 b)
{
    f2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ f2 f2_fma(f2 a, f2 b, f2 c)
{
    f2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return r;
}
// a - m and a + m where m is the result of f2_mul: ptxas contracts mul.rn.f32x2 + add/sub.rn.f32x2 into one FFMA2
// (even with --fmad=false), which would change the rounding; fma(m, -+1, a) is the same single-rounded a -+ m and
// cannot be contracted.
__device__ __forceinline__ f2 f2_sub_prod(f2 a, f2 m) { return f2_fma(m, f2_pack(-1.0f, -1.0f), a); }
__device__ __forceinline__ f2 f2_add_prod(f2 a, f2 m) { return f2_fma(m, f2_pack(1.0f, 1.0f), a); }
__device__ __forceinline__ f2 f2_neg(f2 a)
{
    float lo, hi;
    f2_unpack(a, lo, hi);
    return f2_pack(-lo, -hi);
}
__device__ __forceinline__ f2 f2_abs(f2 a)
{
    float lo, hi;
    f2_unpack(a, lo, hi);
    return f2_pack(fabsf(lo), fabsf(hi));
}
__device__ __forceinline__ f2 f2_max0(f2 a)
{
    float lo, hi;
    f2_unpack(a, lo, hi);
    return f2_pack(fmaxf(lo, 0.0f), fmaxf(hi, 0.0f));
}
__device__ __forceinline__ f2 f2_rcp_approx(f2 a)
{
    float lo, hi, rl, rh;
    f2_unpack(a, lo, hi);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rl) : "f"(lo));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rh) : "f"(hi));
    return f2_pack(rl, rh);
}
__device__ __forceinline__ float f2_hsum(f2 a)
{
    float lo, hi;
    f2_unpack(a, lo, hi);
    return lo + hi;
}
__device__ __forceinline__ f2 lds64(uint32_t addr)
{
    f2 r;
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(r.v) : "r"(addr));
    return r;
}

// ------------------------------------------------------------------------------------------
// colour front-end
// ------------------------------------------------------------------------------------------
__device__ const exact_math::PowfTables kPowfTablesInit = {{EM_POWF_LOG2_TAB}, {EM_EXP2F_TAB}};
__constant__ exact_math::Consts kEM = EM_CONSTS_INIT;
// cbrt scale table in the constant bank (filled once per device by ssimu2_create): neighbouring pixels have similar
// magnitudes, so the per-lane indexed LDC sees 1-3 distinct addresses per warp and the look-up stays off the LSU pipe,
// which the front-end's table gathers keep busy
__constant__ exact_math::CbrtScale kCbrtC;

// BT709::eotf, cuda-colorspace-kernel/src/lib.rs:220-236 (same body for the BT601 variants).
// The reference uses __nv_fast_powf (not reproducible on a CPU); the oracle and this kernel both use
// glibc's powf, bit for bit (exact_math.cuh).
// The argument comes from integer samples and f32 coefficients: finite, and > 0.16 on the power branch,
// so the unchecked powf applies.
__device__ __forceinline__ float bt709_eotf(float v, const exact_math::PowfTables& T)
{
    const float BETA = 0.018053968510807f;
    const float ALPHA = 1.0f + 5.5f * BETA;
    const float THRESHOLD = 0.08124285829863521110029445797874f;
    if (v >= THRESHOLD)
        return exact_math::powf_glibc<false>(exact_math::fdiv_normal(v + (ALPHA - 1.0f), ALPHA), 1.0f / 0.45f, kEM, T);
    return v / 4.5f;
}

// srgb_inverse_oetf, cuda-colorspace-kernel/src/srgb.rs:40-48.
// CHECKED: the f32 pixel format can carry anything (nan, inf, huge); u16 cannot.
template <bool CHECKED>
__device__ __forceinline__ float srgb_inverse_oetf(float x, const exact_math::PowfTables& T)
{
    const float SRGB_ALPHA = 1.0550107f;
    const float SRGB_BETA = 0.0030412825f;
    if (x < 12.92f * SRGB_BETA)
        return x / 12.92f;
    if (CHECKED) return exact_math::powf_glibc<true>((x + (SRGB_ALPHA - 1.0f)) / SRGB_ALPHA, 2.4f, kEM, T);
    return exact_math::powf_glibc<false>(exact_math::fdiv_normal(x + (SRGB_ALPHA - 1.0f), SRGB_ALPHA), 2.4f, kEM, T);
}

__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }

// One source pixel -> linear RGB.  Replaces biplanaryuv420_to_linearrgb_generic
// (cuda-colorspace-kernel/src/biplanar.rs:7-70), srgb_to_linear_u8_lookup / srgb_to_linear::<16> /
// srgb_to_linear_f32 (srgb.rs:50-127).  x, y must be inside the frame.
template <int FMT>
__device__ __forceinline__ void load_px(const FrameIn& f, int x, int y, const YuvCoef& k, const exact_math::PowfTables& T,
                                        const float* __restrict__ lut, int lut_n, int lut_shift, float& r, float& g, float& b)
{
    if constexpr (FmtIs<FMT>::yuv) {
        int Y, cbi, cri;
        if constexpr (FMT == kNV12) {
            Y = __ldg(f.p0 + (size_t)y * f.pitch + x);
            const uint8_t* uv = f.p1 + (size_t)(y >> 1) * f.pitch + 2 * (x >> 1);
            uchar2 c = __ldg(reinterpret_cast<const uchar2*>(uv));
            cbi = c.x; cri = c.y;
        } else {
            Y = __ldg(reinterpret_cast<const uint16_t*>(f.p0 + (size_t)y * f.pitch) + x);
            const uint16_t* uv = reinterpret_cast<const uint16_t*>(f.p1 + (size_t)(y >> 1) * f.pitch) + 2 * (x >> 1);
            ushort2 c = __ldg(reinterpret_cast<const ushort2*>(uv));
            cbi = c.x; cri = c.y;
        }
        float cb = (float)(cbi - k.neutral);
        float cr = (float)(cri - k.neutral);
        float g_ = fmaf(k.g1, cb, k.g2 * cr);
        float luma = (float)(max(Y, k.luma_min) - k.luma_min) * k.y;
        g = clamp01(bt709_eotf(luma + g_, T));
        if (lut != nullptr && ((Y | cbi | cri) & ((1 << lut_shift) - 1)) == 0) {
            const int yc = Y >> lut_shift;
            r = __ldg(lut + (size_t)(cri >> lut_shift) * lut_n + yc);
            b = __ldg(lut + (size_t)lut_n * lut_n + (size_t)(cbi >> lut_shift) * lut_n + yc);
        } else {
            float r_ = k.r * cr;
            float b_ = k.b * cb;
            r = clamp01(bt709_eotf(luma + r_, T));
            b = clamp01(bt709_eotf(luma + b_, T));
        }
    } else if constexpr (FMT == kSRGB8) {
        const uint8_t* p = f.p0 + (size_t)y * f.pitch + 3 * x;
        r = kSrgb8Lut[__ldg(p)];
        g = kSrgb8Lut[__ldg(p + 1)];
        b = kSrgb8Lut[__ldg(p + 2)];
    } else if constexpr (FMT == kSRGB16) {
        const uint16_t* p = reinterpret_cast<const uint16_t*>(f.p0 + (size_t)y * f.pitch) + 3 * x;
        if (lut != nullptr) {
            // exact memo by code (k_build_srgb16_lut: the same expression, evaluated once per code at create time)
            r = __ldg(lut + __ldg(p)); g = __ldg(lut + __ldg(p + 1)); b = __ldg(lut + __ldg(p + 2));
        } else {
            r = srgb_inverse_oetf<false>((float)__ldg(p) / 65535.0f, T);
            g = srgb_inverse_oetf<false>((float)__ldg(p + 1) / 65535.0f, T);
            b = srgb_inverse_oetf<false>((float)__ldg(p + 2) / 65535.0f, T);
        }
    } else {
        const float* p = reinterpret_cast<const float*>(f.p0 + (size_t)y * f.pitch) + 3 * x;
        r = __ldg(p); g = __ldg(p + 1); b = __ldg(p + 2);
        if constexpr (FMT == kSRGBF32) {
            r = srgb_inverse_oetf<true>(r, T); g = srgb_inverse_oetf<true>(g, T); b = srgb_inverse_oetf<true>(b, T);
        }
    }
}

// linear RGB -> rescaled XYB.  cpu.rs:421-496 (== ssimulacra2-cuda-kernel/src/xyb.rs:3-102),
// with the CPU path's libm cbrtf reproduced exactly.
__device__ __forceinline__ void linear_to_xyb(float r, float g, float b, const exact_math::CbrtScale& S, float& X, float& Y,
                                              float& B)
{
    const float K_M02 = 0.078f, K_M00 = 0.30f, K_M01 = 1.0f - K_M02 - K_M00;
    const float K_M12 = 0.078f, K_M10 = 0.23f, K_M11 = 1.0f - K_M12 - K_M10;
    const float K_M20 = 0.24342269f, K_M21 = 0.20476745f, K_M22 = 1.0f - K_M20 - K_M21;
    const float K_B0 = 0.0037930734f;
    const float K_B0_ROOT = 0.1559542025327239180319220163705f;
    float rg = fmaf(K_M00, r, fmaf(K_M01, g, fmaf(K_M02, b, K_B0)));
    float gr = fmaf(K_M10, r, fmaf(K_M11, g, fmaf(K_M12, b, K_B0)));
    float bb = fmaf(K_M20, r, fmaf(K_M21, g, fmaf(K_M22, b, K_B0)));
    rg = fmaxf(rg, 0.0f); gr = fmaxf(gr, 0.0f); bb = fmaxf(bb, 0.0f);
    // in-range frames give arguments in [0.0037, 1.01]; anything else (only possible with the f32 pixel
    // formats) takes the fully checked routine
    const float lo = fminf(fminf(rg, gr), bb), hi = fmaxf(fmaxf(rg, gr), bb);
    if (lo >= 1.17549435e-38f && hi <= 3.0e38f) {
        rg = exact_math::cbrtf_glibc<false>(rg, kEM, &S);
        gr = exact_math::cbrtf_glibc<false>(gr, kEM, &S);
        bb = exact_math::cbrtf_glibc<false>(bb, kEM, &S);
    } else {
        rg = exact_math::cbrtf_glibc<true>(rg, kEM, &S);
        gr = exact_math::cbrtf_glibc<true>(gr, kEM, &S);
        bb = exact_math::cbrtf_glibc<true>(bb, kEM, &S);
    }
    rg -= K_B0_ROOT; gr -= K_B0_ROOT; bb -= K_B0_ROOT;
    float x = 0.5f * (rg - gr);
    float y = 0.5f * (rg + gr);
    X = fmaf(x, 14.0f, 0.42f);
    Y = y + 0.01f;
    B = (bb - y) + 0.55f;
}

// ------------------------------------------------------------------------------------------
// Front-end: colour conversion + linear-RGB pyramid + XYB of every scale.
// Replaces the colour conversion kernels, downscale_by_2 (ssimulacra2-cuda-kernel/src/downscale.rs:4-35,
// host loop ssimulacra2-cuda/src/lib.rs:162-183) and linear_to_xyb_packed (xyb.rs:82-102, lib.rs:188-210);
// follows cpu.rs:545-579 (box sum order (0,0),(1,0),(0,1),(1,1), edge clamp min(src-1), x0.25) and
// cpu.rs:363-377 (downscale in LINEAR RGB, XYB recomputed per scale).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float box4(float a, float b, float c, float d) { return ((((0.0f + a) + b) + c) + d) * 0.25f; }

// ------------------------------------------------------------------------------------------
// k_frontend2: colour conversion, linear-RGB pyramid and XYB of every scale; a WARP owns a 32x32 source region and needs no
// shared-memory pyramid.  Two region routines:
//   frontend_region_fast (further down): interior regions of NV12 / P016 / sRGB8 -- the hot path, see its header;
//   frontend_region (general): every format, frame edges (clamped coordinates), P016 samples with non-zero low bits.
//     lane = a 4x4 pixel patch (8 x 4 patches per pass, two passes): levels 0, 1 and 2 of the pyramid are formed in
//     registers; levels 3 and 4 by warp shuffles; level 5 from the two passes; the 16 + 4 + 1 pixels of levels 3-5 of a
//     region share ONE XYB evaluation.
// Both follow cpu.rs:545-579 operation for operation (box sum order (0,0),(1,0),(0,1),(1,1), edge clamp at every level).
// ------------------------------------------------------------------------------------------
struct YuvChroma {
    float g_, r_, b_;
    const float* rrow;     // memo rows (null: arithmetic path)
    const float* brow;
};

template <int FMT>
__device__ __forceinline__ YuvChroma yuv_chroma(int cbi, int cri, const YuvCoef& k, const float* __restrict__ lut, int lut_n,
                                                int lut_shift)
{
    YuvChroma c;
    const float cb = (float)(cbi - k.neutral), cr = (float)(cri - k.neutral);
    c.g_ = fmaf(k.g1, cb, k.g2 * cr);
    c.r_ = k.r * cr;
    c.b_ = k.b * cb;
    const bool ok = lut != nullptr && ((cbi | cri) & ((1 << lut_shift) - 1)) == 0;
    c.rrow = ok ? lut + (size_t)(cri >> lut_shift) * lut_n : nullptr;
    c.brow = ok ? lut + (size_t)lut_n * lut_n + (size_t)(cbi >> lut_shift) * lut_n : nullptr;
    return c;
}

// out-of-line copy of the transfer function for the rarely taken arithmetic R / B path (keeps the hot loop small)
__device__ __noinline__ float bt709_eotf_cold(float v, const exact_math::PowfTables* T) { return clamp01(bt709_eotf(v, *T)); }

// one luma sample + its block's chroma -> linear RGB (same expressions as load_px)
__device__ __forceinline__ void yuv_px(int Y, const YuvChroma& c, const YuvCoef& k, const exact_math::PowfTables& T, int lut_shift,
                                       float& r, float& g, float& b)
{
    const float luma = (float)(max(Y, k.luma_min) - k.luma_min) * k.y;
    g = clamp01(bt709_eotf(luma + c.g_, T));
    if (c.rrow != nullptr && (Y & ((1 << lut_shift) - 1)) == 0) {
        const int yc = Y >> lut_shift;
        r = __ldg(c.rrow + yc);
        b = __ldg(c.brow + yc);
    } else {
        r = bt709_eotf_cold(luma + c.r_, &T);
        b = bt709_eotf_cold(luma + c.b_, &T);
    }
}

// XYB for the pixel formats whose linear values are guaranteed in [0, 1] (every integer format): no range check
template <int FMT>
__device__ __forceinline__ void xyb_of(float r, float g, float b, const exact_math::CbrtScale& S, float& X, float& Y, float& B)
{
    if constexpr (FMT == kSRGBF32 || FMT == kLINEARF32) {
        linear_to_xyb(r, g, b, S, X, Y, B);
    } else {
        const float K_M02 = 0.078f, K_M00 = 0.30f, K_M01 = 1.0f - K_M02 - K_M00;
        const float K_M12 = 0.078f, K_M10 = 0.23f, K_M11 = 1.0f - K_M12 - K_M10;
        const float K_M20 = 0.24342269f, K_M21 = 0.20476745f, K_M22 = 1.0f - K_M20 - K_M21;
        const float K_B0 = 0.0037930734f;
        const float K_B0_ROOT = 0.1559542025327239180319220163705f;
        float rg = fmaf(K_M00, r, fmaf(K_M01, g, fmaf(K_M02, b, K_B0)));
        float gr = fmaf(K_M10, r, fmaf(K_M11, g, fmaf(K_M12, b, K_B0)));
        float bb = fmaf(K_M20, r, fmaf(K_M21, g, fmaf(K_M22, b, K_B0)));
        rg = fmaxf(rg, 0.0f); gr = fmaxf(gr, 0.0f); bb = fmaxf(bb, 0.0f);
        rg = exact_math::cbrtf_glibc<false>(rg, kEM, &S) - K_B0_ROOT;
        gr = exact_math::cbrtf_glibc<false>(gr, kEM, &S) - K_B0_ROOT;
        bb = exact_math::cbrtf_glibc<false>(bb, kEM, &S) - K_B0_ROOT;
        const float x = 0.5f * (rg - gr), y = 0.5f * (rg + gr);
        X = fmaf(x, 14.0f, 0.42f);
        Y = y + 0.01f;
        B = (bb - y) + 0.55f;
    }
}

struct Rgb {
    float r, g, b;
};
// 2x2 box of a (x,y), b (x+1,y), c (x,y+1), d (x+1,y+1) with the edge clamp of downscale_by_2: a neighbour past the
// edge is replaced by the clamped one.
__device__ __forceinline__ Rgb box_clamped(Rgb a, Rgb b, Rgb c, Rgb d, bool xin, bool yin)
{
    if (!xin) { b = a; d = c; }
    if (!yin) { c = a; d = b; }
    return Rgb{box4(a.r, b.r, c.r, d.r), box4(a.g, b.g, c.g, d.g), box4(a.b, b.b, c.b, d.b)};
}
__device__ __forceinline__ Rgb shfl_rgb_xor(Rgb v, int m)
{
    return Rgb{__shfl_xor_sync(0xffffffffu, v.r, m), __shfl_xor_sync(0xffffffffu, v.g, m), __shfl_xor_sync(0xffffffffu, v.b, m)};
}
__device__ __forceinline__ Rgb shfl_rgb(Rgb v, int src)
{
    return Rgb{__shfl_sync(0xffffffffu, v.r, src), __shfl_sync(0xffffffffu, v.g, src), __shfl_sync(0xffffffffu, v.b, src)};
}

#ifndef KF2_MINB
#define KF2_MINB 4     // x 128 threads = 16 warps per SM at 128 registers: the lock-step FP64 chains need the registers (no spills)
#endif
constexpr int kF2Region = 32;
#ifndef KF2_THREADS
#define KF2_THREADS 128
#endif

constexpr int kF2Threads = KF2_THREADS;
constexpr int kF2RegionsPerWarp = 2;    // consecutive regions along x

// ---- fast path of the front-end: regions that lie completely inside the frame (all but the last row / column of
// regions), NV12 / P016 / sRGB8.  Same arithmetic as frontend_region below, operation for operation; what changes is the
// schedule:
//   lane = a 4 (wide) x 8 (tall) pixel patch of the 32x32 region, walked one ROW of four pixels at a time, so every XYB row
//   leaves as one 128-bit store per plane (a warp instruction writes whole 128-byte lines) and every level-1 row as one
//   64-bit store; the pyramid needs no parking: the 2x2 box of level 1 is accumulated in the reference's order while the
//   two rows go by ((0,0)+(1,0) from the first row, then +(0,1), +(1,1)), level 2 likewise over two row pairs, levels 3-5
//   by shuffles at the end of the region;
//   integer -> float conversions and the per-pixel luma arithmetic come from two small shared-memory tables
//   (exact: built with the same expressions), the divisions by the transfer-function constants use a compile-time
//   reciprocal + Markstein's correction (IEEE-exact quotient, tests/test_gpu_parity.py::test_device_divisions_are_ieee),
//   the transfer function is evaluated branch-free (both segments, select), XYB runs on pixel PAIRS in packed f32x2.
struct FeTables {
    exact_math::PowfTables T;
    exact_math::CbrtScale S;
    float luma[1024];     // YUV: (float)(max(Y, luma_min) - luma_min) * k.y by code (Y >> lut_shift); sRGB8: the 256-entry table
    float chroma[1024];   // YUV: (float)(C - neutral) by code
};

// n / d for a compile-time d: r = RN(1 / d); q = RN(n r); the residual n - d q is exact in one FMA and the correction
// lands on the correctly rounded quotient (Markstein).  n = 0 gives 0; n must not be so small that the quotient is subnormal.
__device__ __forceinline__ float fdiv_rcp(float n, float d, float r)
{
    const float q = n * r;
    const float rem = fmaf(-d, q, n);
    return fmaf(rem, r, q);
}

// bt709_eotf + clamp01, branch-free (the power segment is evaluated for every argument and discarded below the threshold;
// powf_glibc<false> is straight-line code with masked table indices, so a non-positive argument costs nothing but garbage)
__device__ __forceinline__ float bt709_eotf_clamped(float v, const exact_math::PowfTables& T)
{
    const float BETA = 0.018053968510807f;
    const float ALPHA = 1.0f + 5.5f * BETA;
    const float THRESHOLD = 0.08124285829863521110029445797874f;
    const float p = exact_math::powf_glibc<false>(fdiv_rcp(v + (ALPHA - 1.0f), ALPHA, 1.0f / ALPHA), 1.0f / 0.45f, kEM, T);
    const float l = fdiv_rcp(v, 4.5f, 1.0f / 4.5f);
    return clamp01(v >= THRESHOLD ? p : l);
}

// bt709_eotf_clamped on N values with the power segments advanced in lock-step
template <int N>
__device__ __forceinline__ void bt709_eotf_clamped_n(float (&v)[N], const exact_math::PowfTables& T)
{
    const float BETA = 0.018053968510807f;
    const float ALPHA = 1.0f + 5.5f * BETA;
    const float THRESHOLD = 0.08124285829863521110029445797874f;
    float p[N];
#pragma unroll
    for (int i = 0; i < N; i++) p[i] = fdiv_rcp(v[i] + (ALPHA - 1.0f), ALPHA, 1.0f / ALPHA);
    exact_math::powf_glibc_n<N>(p, 1.0f / 0.45f, kEM, T);
#pragma unroll
    for (int i = 0; i < N; i++) {
        const float l = fdiv_rcp(v[i], 4.5f, 1.0f / 4.5f);
        v[i] = clamp01(v[i] >= THRESHOLD ? p[i] : l);
    }
}

// srgb_inverse_oetf on N values that are 0 or in [1e-30, 1e6] (checked by the caller), branch-free with the power segments in
// lock-step: the same quotients (fdiv_normal is the IEEE quotient on normal operands) and the same powf as the checked routine
template <int N>
__device__ __forceinline__ void srgb_inverse_oetf_n(float (&v)[N], const exact_math::PowfTables& T)
{
    const float SRGB_ALPHA = 1.0550107f;
    const float SRGB_BETA = 0.0030412825f;
    float p[N];
#pragma unroll
    for (int i = 0; i < N; i++) p[i] = exact_math::fdiv_normal(v[i] + (SRGB_ALPHA - 1.0f), SRGB_ALPHA);
    exact_math::powf_glibc_n<N>(p, 2.4f, kEM, T);
#pragma unroll
    for (int i = 0; i < N; i++) {
        const float l = exact_math::fdiv_normal(v[i], 12.92f);
        v[i] = v[i] < 12.92f * SRGB_BETA ? l : p[i];
    }
}

// not inlined: five call sites per row pair; as a function the hot loop is 0.7 k instead of 1.7 k instructions (11 KB instead of
// 27 KB of code per format), which the instruction cache prefers (NV12 front-end 1.47 -> 1.37 ms per 32 1080p pairs)
#ifndef XYB_PAIR_INLINE
#define XYB_PAIR_INLINE __noinline__
#endif
// XYB of two pixels whose linear values are in [0, 1] (the integer formats): xyb_of on both lanes of a packed pair.
// The fmaxf(mixed, 0) of the reference is dropped: mixed >= bias > 0 for non-negative inputs, so it never changes a bit.
struct Xyb2 {
    f2 X, Y, B;
};
// (arguments and results by value: references would travel through local memory at the call)
__device__ XYB_PAIR_INLINE Xyb2 xyb_pair_v(f2 r, f2 g, f2 b)
{
    f2 X, Y, B;
    const float K_M02 = 0.078f, K_M00 = 0.30f, K_M01 = 1.0f - K_M02 - K_M00;
    const float K_M12 = 0.078f, K_M10 = 0.23f, K_M11 = 1.0f - K_M12 - K_M10;
    const float K_M20 = 0.24342269f, K_M21 = 0.20476745f, K_M22 = 1.0f - K_M20 - K_M21;
    const float K_B0 = 0.0037930734f;
    const float K_B0_ROOT = 0.1559542025327239180319220163705f;
    const f2 bias = f2_splat(K_B0);
    const f2 mrg = f2_fma(f2_splat(K_M00), r, f2_fma(f2_splat(K_M01), g, f2_fma(f2_splat(K_M02), b, bias)));
    const f2 mgr = f2_fma(f2_splat(K_M10), r, f2_fma(f2_splat(K_M11), g, f2_fma(f2_splat(K_M12), b, bias)));
    const f2 mbb = f2_fma(f2_splat(K_M20), r, f2_fma(f2_splat(K_M21), g, f2_fma(f2_splat(K_M22), b, bias)));
    float a0, a1, b0, b1, c0, c1;
    f2_unpack(mrg, a0, a1);
    f2_unpack(mgr, b0, b1);
    f2_unpack(mbb, c0, c1);
    {
        float v[6] = {a0, a1, b0, b1, c0, c1};
        exact_math::cbrtf_glibc_n<6>(v, kEM, &kCbrtC);   // six chains in lock-step: see exact_math.cuh
        a0 = v[0]; a1 = v[1]; b0 = v[2]; b1 = v[3]; c0 = v[4]; c1 = v[5];
    }
    const f2 nroot = f2_splat(-K_B0_ROOT);
    const f2 rg = f2_add(f2_pack(a0, a1), nroot), gr = f2_add(f2_pack(b0, b1), nroot), bb = f2_add(f2_pack(c0, c1), nroot);
    const f2 half = f2_splat(0.5f);
    const f2 x = f2_mul(half, f2_sub(rg, gr)), y = f2_mul(half, f2_add(rg, gr));   // 0.5 * s is exact: a later contraction into an FMA cannot change the result
    X = f2_fma(x, f2_splat(14.0f), f2_splat(0.42f));
    Y = f2_add(y, f2_splat(0.01f));
    B = f2_add(f2_sub(bb, y), f2_splat(0.55f));
    return Xyb2{X, Y, B};
}
__device__ __forceinline__ void xyb_pair(f2 r, f2 g, f2 b, const exact_math::CbrtScale&, f2& X, f2& Y, f2& B)
{
    const Xyb2 o = xyb_pair_v(r, g, b);
    X = o.X; Y = o.Y; B = o.B;
}


__device__ __forceinline__ float lds_f32(uint32_t addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

template <int FMT>
struct FastFmt {
    static constexpr bool ok = true;   // every format has the fast schedule (edge regions / odd alignments use the general path)
    static constexpr bool yuv = FmtIs<FMT>::yuv;
    // bytes per sample of plane 0 (YUV: luma sample; packed formats: pixel), alignment the row loads need (mask)
    static constexpr int bpp = FmtIs<FMT>::p016 ? 2 : (FMT == kNV12 ? 1 : (FMT == kSRGB8 ? 3 : (FMT == kSRGB16 ? 6 : 12)));
    static constexpr uint32_t align = FmtIs<FMT>::p016 || FMT == kSRGB16 ? 7u : (FMT == kLINEARF32 || FMT == kSRGBF32 ? 15u : 3u);
    static constexpr bool needs_lut = FMT == kNV12 || FMT == kP016 || FMT == kSRGB16;
};

// the raw words of one row of four pixels (P016: 2 words, NV12: 1, sRGB8: 3) / of the two chroma samples under them
struct RowWords {
    uint32_t w0, w1, w2, w3, w4, w5, w6, w7, w8, w9, w10, w11;   // sRGB16 uses six, linear f32 all twelve
};
// `asm volatile` on purpose: the compiler sinks an ordinary load to just above its first use, which puts the whole DRAM
// latency in front of the warp; a volatile asm keeps its place among the stores of the rows, i.e. where it was written --
// half an iteration ahead of the consumer.
template <int FMT>
__device__ __forceinline__ RowWords load_row_words(const uint8_t* p)
{
    RowWords r{0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    if constexpr (FMT == kSRGB16) {
        asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(r.w0), "=r"(r.w1) : "l"(p));
        asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2+8];" : "=r"(r.w2), "=r"(r.w3) : "l"(p));
        asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2+16];" : "=r"(r.w4), "=r"(r.w5) : "l"(p));
    } else if constexpr (FMT == kLINEARF32 || FMT == kSRGBF32) {
        asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.w0), "=r"(r.w1), "=r"(r.w2), "=r"(r.w3) : "l"(p));
        asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4+16];" : "=r"(r.w4), "=r"(r.w5), "=r"(r.w6), "=r"(r.w7) : "l"(p));
        asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4+32];" : "=r"(r.w8), "=r"(r.w9), "=r"(r.w10), "=r"(r.w11) : "l"(p));
    } else if constexpr (FmtIs<FMT>::p016) {
        asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(r.w0), "=r"(r.w1) : "l"(p));
    } else if constexpr (FMT == kNV12) {
        asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(r.w0) : "l"(p));
    } else {
        asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(r.w0) : "l"(p));
        asm volatile("ld.global.nc.u32 %0, [%1+4];" : "=r"(r.w1) : "l"(p));
        asm volatile("ld.global.nc.u32 %0, [%1+8];" : "=r"(r.w2) : "l"(p));
    }
    return r;
}

// L2 prefetch of the raw samples of one interior region (this lane's 8 luma rows + 4 chroma rows).  A prefetch has no
// destination register, so unlike a load it cannot be sunk to its consumer: issued one work item ahead, it turns the DRAM
// latency in front of every first use into an L2 hit.
template <int FMT>
__device__ __forceinline__ void prefetch_region(const FrameIn& f, int X0, int Y0)
{
    const int lane = threadIdx.x & 31;
    const int x = X0 + 4 * (lane & 7), y0 = Y0 + 8 * (lane >> 3);
    constexpr int kBpp = FastFmt<FMT>::bpp;
    const uint8_t* pY = f.p0 + (size_t)y0 * f.pitch + (size_t)(x * kBpp);
#pragma unroll
    for (int r = 0; r < 8; r++) asm volatile("prefetch.global.L2 [%0];" ::"l"(pY + (size_t)r * f.pitch));
    if constexpr (FastFmt<FMT>::yuv) {
        const uint8_t* pC = f.p1 + (size_t)(y0 >> 1) * f.pitch + (size_t)(x * kBpp);
#pragma unroll
        for (int r = 0; r < 4; r++) asm volatile("prefetch.global.L2 [%0];" ::"l"(pC + (size_t)r * f.pitch));
    }
}

// One interior 32x32 region of one image; executed by a full warp.  Returns false (warp-uniform) if the region holds P016
// samples with non-zero low bits (not 10-bit content): the caller then redoes the region with the general path.
// Register budget: the six cube roots of a pixel pair advance in lock-step (exact_math::cbrtf_glibc_n: the FP64 pipe has a
// 23-cycle dependent-issue latency, so a warp must carry several independent chains) and want ~60 registers; the kernel
// runs at 128 registers x 16 warps per SM.  The state carried across rows is kept small -- 32-bit lane offsets against
// warp-uniform base pointers, the first of the lane's two level-2 pixels parked in shared memory (`park`, three floats per
// thread) -- and the two pixel pairs of a row are evaluated one after the other (an empty asm ties the second pair's inputs
// to the first pair's results): a value spilled to local memory comes back through an L1 that the table gathers keep
// evicting, and every spill reload showed up as a long-scoreboard stall (profiles/r2_frontend_notes.md).
template <int FMT>
__device__ __forceinline__ bool frontend_region_fast(const Geo& g, const FrameIn& f, float* __restrict__ ximg, int img, int X0, int Y0,
                                                     const FeTables& tb, float* __restrict__ park, int park_stride)
{
    const int lane = threadIdx.x & 31;
    const int pxi = lane & 7, pyi = lane >> 3;
    const int x = X0 + 4 * pxi, y0 = Y0 + 8 * pyi;
    const uint32_t s_luma = smem_u32(tb.luma), s_chroma = smem_u32(tb.chroma);
    const uint32_t pitch0 = (uint32_t)g.sc[0].pitch, pitch1 = (uint32_t)g.sc[1].pitch, pitch2 = (uint32_t)g.sc[2].pitch;
    const size_t plane0 = (size_t)g.sc[0].h * pitch0, plane1 = (size_t)g.sc[1].h * pitch1, plane2 = (size_t)g.sc[2].h * pitch2;
    // warp-uniform plane bases of levels 0-2, per-lane element offsets
    float* const b0 = ximg + g.sc[0].xyb_off + (size_t)img * 3 * plane0;
    float* const b1 = ximg + g.sc[1].xyb_off + (size_t)img * 3 * plane1;
    float* const b2 = ximg + g.sc[2].xyb_off + (size_t)img * 3 * plane2;
    uint32_t o0 = (uint32_t)y0 * pitch0 + (uint32_t)x;
    uint32_t o1 = (uint32_t)(y0 >> 1) * pitch1 + (uint32_t)(x >> 1);
    const uint32_t o2 = (uint32_t)(y0 >> 2) * pitch2 + (uint32_t)(x >> 2);
    constexpr int kBpp = FastFmt<FMT>::bpp;
    constexpr bool kYuv = FastFmt<FMT>::yuv;
    uint32_t oY = (uint32_t)y0 * f.pitch + (uint32_t)(x * kBpp);
    uint32_t oC = (uint32_t)(y0 >> 1) * f.pitch + (uint32_t)(x * kBpp);
    const char* lut = reinterpret_cast<const char*>(g.eotf_lut);
    const uint32_t lut_b = (uint32_t)g.lut_n * (uint32_t)g.lut_n * 4u;   // byte offset of the B table
    uint32_t bad = 0;
    uint32_t tiny = 0xffffffffu;   // sRGB f32: smallest (bits - 1) seen, i.e. the smallest NON-ZERO sample
    float s1[2][3];
    float s2[3] = {0.f, 0.f, 0.f};
    Rgb l2b = Rgb{0.f, 0.f, 0.f};

    // The raw samples are fetched one step ahead (the second row of a pair at the top of the pair, the next pair's chroma
    // and first row in its middle): a load that is consumed right away exposes the full DRAM latency to the warp.
    RowWords wy_next = load_row_words<FMT>(f.p0 + oY), wc_next{0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    oY += f.pitch;
    if constexpr (kYuv) {
        wc_next = load_row_words<FMT>(f.p1 + oC);
        oC += f.pitch;
    }
#pragma unroll 1
    for (int rp = 0; rp < 4; rp++) {
        const RowWords wc = wc_next, wy0 = wy_next;
        const RowWords wy1 = load_row_words<FMT>(f.p0 + oY);
        oY += f.pitch;
        // ---- chroma of the row pair: two samples, each shared by a 2x2 block
        float g_[2] = {0.f, 0.f};
        float r_[2] = {0.f, 0.f}, b_[2] = {0.f, 0.f};   // kP016A: the chroma terms of R' and B' (yuv_chroma)
        uint32_t roff[2] = {0, 0}, boff[2] = {0, 0};
        if constexpr (FMT == kP016A) {
            const uint32_t w[2] = {wc.w0, wc.w1};
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const float cb = (float)((int)(w[j] & 0xFFFFu) - g.coef.neutral), cr = (float)((int)(w[j] >> 16) - g.coef.neutral);
                g_[j] = fmaf(g.coef.g1, cb, g.coef.g2 * cr);
                r_[j] = g.coef.r * cr;
                b_[j] = g.coef.b * cb;
            }
        } else if constexpr (FMT == kP016) {
            bad |= wc.w0 | wc.w1;
            const uint32_t w[2] = {wc.w0, wc.w1};
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const float cb = lds_f32(s_chroma + ((w[j] >> 4) & 0xFFCu)), cr = lds_f32(s_chroma + ((w[j] >> 20) & 0xFFCu));
                g_[j] = fmaf(g.coef.g1, cb, g.coef.g2 * cr);
                boff[j] = lut_b + ((w[j] << 6) & 0x3FF000u);
                roff[j] = (w[j] >> 10) & 0x3FF000u;
            }
        } else if constexpr (FMT == kNV12) {
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const uint32_t cbi = (wc.w0 >> (16 * j)) & 0xFFu, cri = (wc.w0 >> (16 * j + 8)) & 0xFFu;
                const float cb = lds_f32(s_chroma + cbi * 4u), cr = lds_f32(s_chroma + cri * 4u);
                g_[j] = fmaf(g.coef.g1, cb, g.coef.g2 * cr);
                boff[j] = lut_b + (cbi << 10);
                roff[j] = cri << 10;
            }
        }
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const RowWords wy = r == 0 ? wy0 : wy1;
            if (r == 1 && rp < 3) {
                // next pair: chroma + first row
                wy_next = load_row_words<FMT>(f.p0 + oY);
                oY += f.pitch;
                if constexpr (kYuv) {
                    wc_next = load_row_words<FMT>(f.p1 + oC);
                    oC += f.pitch;
                }
            }
            // ---- one row of four pixels -> linear RGB -> XYB, as two pixel pairs
            float lr[4], lb[4], lg[4];
            uint32_t yc[4] = {0, 0, 0, 0};
            if constexpr (FMT == kP016) {
                bad |= wy.w0 | wy.w1;
                yc[0] = (wy.w0 >> 4) & 0xFFCu; yc[1] = (wy.w0 >> 20) & 0xFFCu; yc[2] = (wy.w1 >> 4) & 0xFFCu; yc[3] = (wy.w1 >> 20) & 0xFFCu;
            } else if constexpr (FMT == kNV12) {
                yc[0] = (wy.w0 << 2) & 0x3FCu; yc[1] = (wy.w0 >> 6) & 0x3FCu; yc[2] = (wy.w0 >> 14) & 0x3FCu; yc[3] = (wy.w0 >> 22) & 0x3FCu;
            }
            if constexpr (FMT == kP016A) {
                // the same expressions as yuv_px on the full 16-bit samples; three transfer functions per pixel, as three
                // lock-step groups of four
                const int Ys[4] = {(int)(wy.w0 & 0xFFFFu), (int)(wy.w0 >> 16), (int)(wy.w1 & 0xFFFFu), (int)(wy.w1 >> 16)};
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float luma = (float)(max(Ys[i], g.coef.luma_min) - g.coef.luma_min) * g.coef.y;
                    lr[i] = luma + r_[i >> 1];
                    lg[i] = luma + g_[i >> 1];
                    lb[i] = luma + b_[i >> 1];
                }
                bt709_eotf_clamped_n<4>(lr, tb.T);
                bt709_eotf_clamped_n<4>(lg, tb.T);
                bt709_eotf_clamped_n<4>(lb, tb.T);
            } else if constexpr (kYuv) {
                // the exact R / B memo: all eight gathers of the row are issued before the first transfer-function chain
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    lr[i] = __ldg(reinterpret_cast<const float*>(lut + (roff[i >> 1] + yc[i])));
                    lb[i] = __ldg(reinterpret_cast<const float*>(lut + (boff[i >> 1] + yc[i])));
                }
            } else if constexpr (FMT == kSRGB16) {
                // u16 codes R0 G0 | B0 R1 | G1 B1 | R2 G2 | B2 R3 | G3 B3 -> the exact transfer memo (one gather per code)
                const float* t16 = g.eotf_lut;
                auto lo = [&](uint32_t w) { return __ldg(t16 + (w & 0xFFFFu)); };
                auto hi = [&](uint32_t w) { return __ldg(t16 + (w >> 16)); };
                lr[0] = lo(wy.w0); lg[0] = hi(wy.w0); lb[0] = lo(wy.w1);
                lr[1] = hi(wy.w1); lg[1] = lo(wy.w2); lb[1] = hi(wy.w2);
                lr[2] = lo(wy.w3); lg[2] = hi(wy.w3); lb[2] = lo(wy.w4);
                lr[3] = hi(wy.w4); lg[3] = lo(wy.w5); lb[3] = hi(wy.w5);
            } else if constexpr (FMT == kLINEARF32) {
                // the values themselves.  The unchecked cube root wants non-negative, finite-sized arguments: the sign bit is
                // dropped here (keeps every table index in range) and any value that was negative, -0, nan or above 1e30
                // marks the region for the general path, which has the fully checked routines
                const uint32_t ww[12] = {wy.w0, wy.w1, wy.w2, wy.w3, wy.w4, wy.w5, wy.w6, wy.w7, wy.w8, wy.w9, wy.w10, wy.w11};
#pragma unroll
                for (int i = 0; i < 12; i++) bad = max(bad, ww[i]);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    lr[i] = __uint_as_float(ww[3 * i] & 0x7fffffffu);
                    lg[i] = __uint_as_float(ww[3 * i + 1] & 0x7fffffffu);
                    lb[i] = __uint_as_float(ww[3 * i + 2] & 0x7fffffffu);
                }
            } else if constexpr (FMT == kSRGBF32) {
                // non-linear f32 samples: the checked transfer function of the general path handles anything; here a sample must
                // be 0 or in [1e-30, 1e6] (normal operands and quotients for the hand-rolled divisions, a normal positive
                // argument for the unchecked powf) -- otherwise the region is redone by the general path
                const uint32_t ww[12] = {wy.w0, wy.w1, wy.w2, wy.w3, wy.w4, wy.w5, wy.w6, wy.w7, wy.w8, wy.w9, wy.w10, wy.w11};
#pragma unroll
                for (int i = 0; i < 12; i++) {
                    bad = max(bad, ww[i]);
                    tiny = min(tiny, ww[i] - 1u);
                }
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    lr[i] = __uint_as_float(ww[3 * i] & 0x7fffffffu);
                    lg[i] = __uint_as_float(ww[3 * i + 1] & 0x7fffffffu);
                    lb[i] = __uint_as_float(ww[3 * i + 2] & 0x7fffffffu);
                }
                srgb_inverse_oetf_n<4>(lr, tb.T);
                srgb_inverse_oetf_n<4>(lg, tb.T);
                srgb_inverse_oetf_n<4>(lb, tb.T);
            } else {
                // R0 G0 B0 R1 | G1 B1 R2 G2 | B2 R3 G3 B3
                auto tab = [&](uint32_t w, int k) { return lds_f32(s_luma + (k == 0 ? (w << 2) & 0x3FCu : (w >> (8 * k - 2)) & 0x3FCu)); };
                lr[0] = tab(wy.w0, 0); lg[0] = tab(wy.w0, 1); lb[0] = tab(wy.w0, 2);
                lr[1] = tab(wy.w0, 3); lg[1] = tab(wy.w1, 0); lb[1] = tab(wy.w1, 1);
                lr[2] = tab(wy.w1, 2); lg[2] = tab(wy.w1, 3); lb[2] = tab(wy.w2, 0);
                lr[3] = tab(wy.w2, 1); lg[3] = tab(wy.w2, 2); lb[3] = tab(wy.w2, 3);
            }
            if constexpr (kYuv && FMT != kP016A) {
#pragma unroll
                for (int i = 0; i < 4; i++) lg[i] = lds_f32(s_luma + yc[i]) + g_[i >> 1];
                bt709_eotf_clamped_n<4>(lg, tb.T);
            }
            f2 Xp[2], Yp[2], Bp[2];
#pragma unroll
            for (int j = 0; j < 2; j++) {
                // level 1: box sums in the reference's order (0,0), (1,0), (0,1), (1,1); every value is >= +0, so the leading
                // "0 +" of box4 is the identity
                if (r == 0) {
                    s1[j][0] = lr[2 * j] + lr[2 * j + 1];
                    s1[j][1] = lg[2 * j] + lg[2 * j + 1];
                    s1[j][2] = lb[2 * j] + lb[2 * j + 1];
                } else {
                    s1[j][0] = (s1[j][0] + lr[2 * j]) + lr[2 * j + 1];
                    s1[j][1] = (s1[j][1] + lg[2 * j]) + lg[2 * j + 1];
                    s1[j][2] = (s1[j][2] + lb[2 * j]) + lb[2 * j + 1];
                }
                xyb_pair(f2_pack(lr[2 * j], lr[2 * j + 1]), f2_pack(lg[2 * j], lg[2 * j + 1]), f2_pack(lb[2 * j], lb[2 * j + 1]), tb.S, Xp[j],
                         Yp[j], Bp[j]);
                // evaluate the second pair AFTER the first (see the header comment)
                if (j == 0) asm volatile("" : "+f"(lr[2]), "+f"(lr[3]) : "l"(Bp[0].v));
            }
            {
                float e0, e1, e2, e3;
                f2_unpack(Xp[0], e0, e1); f2_unpack(Xp[1], e2, e3);
                *reinterpret_cast<float4*>(b0 + o0) = make_float4(e0, e1, e2, e3);
                f2_unpack(Yp[0], e0, e1); f2_unpack(Yp[1], e2, e3);
                *reinterpret_cast<float4*>(b0 + plane0 + o0) = make_float4(e0, e1, e2, e3);
                f2_unpack(Bp[0], e0, e1); f2_unpack(Bp[1], e2, e3);
                *reinterpret_cast<float4*>(b0 + 2 * plane0 + o0) = make_float4(e0, e1, e2, e3);
            }
            o0 += pitch0;
        }
        // ---- level 1: two pixels, one 64-bit store per plane
        const Rgb v10 = Rgb{s1[0][0] * 0.25f, s1[0][1] * 0.25f, s1[0][2] * 0.25f};
        const Rgb v11 = Rgb{s1[1][0] * 0.25f, s1[1][1] * 0.25f, s1[1][2] * 0.25f};
        {
            f2 X, Y, B;
            xyb_pair(f2_pack(v10.r, v11.r), f2_pack(v10.g, v11.g), f2_pack(v10.b, v11.b), tb.S, X, Y, B);
            float e0, e1;
            f2_unpack(X, e0, e1); *reinterpret_cast<float2*>(b1 + o1) = make_float2(e0, e1);
            f2_unpack(Y, e0, e1); *reinterpret_cast<float2*>(b1 + plane1 + o1) = make_float2(e0, e1);
            f2_unpack(B, e0, e1); *reinterpret_cast<float2*>(b1 + 2 * plane1 + o1) = make_float2(e0, e1);
            o1 += pitch1;
        }
        // ---- level 2: box of the 2x2 level-1 pixels of two consecutive row pairs
        if ((rp & 1) == 0) {
            s2[0] = v10.r + v11.r; s2[1] = v10.g + v11.g; s2[2] = v10.b + v11.b;
        } else {
            const Rgb v2 = Rgb{((s2[0] + v10.r) + v11.r) * 0.25f, ((s2[1] + v10.g) + v11.g) * 0.25f, ((s2[2] + v10.b) + v11.b) * 0.25f};
            if (rp == 1) {
                park[0] = v2.r; park[park_stride] = v2.g; park[2 * park_stride] = v2.b;
            } else {
                l2b = v2;
            }
        }
    }
    const Rgb l2a = Rgb{park[0], park[park_stride], park[2 * park_stride]};
    {
        // the lane's two level-2 pixels (one above the other)
        f2 X, Y, B;
        xyb_pair(f2_pack(l2a.r, l2b.r), f2_pack(l2a.g, l2b.g), f2_pack(l2a.b, l2b.b), tb.S, X, Y, B);
        float e0, e1;
        float* q2 = b2 + o2;
        f2_unpack(X, e0, e1); q2[0] = e0; q2[pitch2] = e1;
        f2_unpack(Y, e0, e1); q2[plane2] = e0; q2[plane2 + pitch2] = e1;
        f2_unpack(B, e0, e1); q2[2 * plane2] = e0; q2[2 * plane2 + pitch2] = e1;
    }
    // ---- levels 3, 4, 5: 16 + 4 + 1 pixels per region
    //   level 3: lanes with even pxi (with the lane to the right); level 4: lanes 0, 4, 16, 20; level 5: lane 0
    const Rgb p0 = shfl_rgb_xor(l2a, 1), p1 = shfl_rgb_xor(l2b, 1);
    const Rgb v3 = Rgb{box4(l2a.r, p0.r, l2b.r, p1.r), box4(l2a.g, p0.g, l2b.g, p1.g), box4(l2a.b, p0.b, l2b.b, p1.b)};
    const Rgb a3 = shfl_rgb_xor(v3, 2), b3 = shfl_rgb_xor(v3, 8), c3 = shfl_rgb_xor(v3, 10);
    const Rgb v4 = Rgb{box4(v3.r, a3.r, b3.r, c3.r), box4(v3.g, a3.g, b3.g, c3.g), box4(v3.b, a3.b, b3.b, c3.b)};
    const Rgb a4 = shfl_rgb_xor(v4, 4), b4 = shfl_rgb_xor(v4, 16), c4 = shfl_rgb_xor(v4, 20);
    const Rgb v5 = Rgb{box4(v4.r, a4.r, b4.r, c4.r), box4(v4.g, a4.g, b4.g, c4.g), box4(v4.b, a4.b, b4.b, c4.b)};
    // one XYB evaluation for all of them: level 3 stays on its lanes, level 4 moves one lane to the right (1, 5, 17, 21),
    // level 5 to lane 3
    const Rgb m4 = shfl_rgb(v4, (lane - 1) & 31), m5 = shfl_rgb(v5, 0);
    Rgb v = v3;
    int s = 3, ox = (X0 >> 3) + (pxi >> 1), oy = (Y0 >> 3) + pyi;
    if (pxi & 1) {
        s = -1;
        if ((lane & 0x0B) == 1) {   // lanes 1, 5, 17, 21
            v = m4; s = 4; ox = (X0 >> 4) + (pxi >> 2); oy = (Y0 >> 4) + (pyi >> 1);
        } else if (lane == 3) {
            v = m5; s = 5; ox = X0 >> 5; oy = Y0 >> 5;
        }
    }
    float X, Yv, B;
    xyb_of<FMT>(v.r, v.g, v.b, tb.S, X, Yv, B);
    if (s >= 3 && s < g.nscales) {
        const ScaleDesc& sd = g.sc[s];
        const size_t plane = (size_t)sd.h * sd.pitch;
        float* q = ximg + sd.xyb_off + (size_t)img * 3 * plane + (size_t)oy * sd.pitch + ox;
        q[0] = X; q[plane] = Yv; q[2 * plane] = B;
    }
    if constexpr (FMT == kP016) return !__any_sync(0xffffffffu, (bad & 0x003F003Fu) != 0);
    if constexpr (FMT == kLINEARF32) return !__any_sync(0xffffffffu, bad > 0x7149f2cau);   // bits of 1e30f; negatives and nan are larger
    if constexpr (FMT == kSRGBF32)   // above 1e6f (0x49742400), or non-zero below 1e-30f (0x0da24260)
        return !__any_sync(0xffffffffu, bad > 0x49742400u || tiny < 0x0da24260u - 1u);
    return true;
}

// One 32x32 region of one image of one frame; executed by a full warp.
template <int FMT>
__device__ __forceinline__ void frontend_region(const Geo& g, const FrameIn& f, float* __restrict__ ximg /* slot base + image */,
                                                int img, int X0, int Y0, const exact_math::PowfTables& T,
                                                const exact_math::CbrtScale& S, float* __restrict__ scr, int scr_stride)
{
    // scr: this thread's column of a shared-memory scratch ([15][threads] floats).  Values that are produced long before
    // they are consumed (the level-1 pixels of a patch, pass 0's level-3 / level-4 pixels) are parked there so the
    // pixel loop fits 64 registers without spilling.
    auto park = [&](int slot, const Rgb& v) {
        scr[(3 * slot + 0) * scr_stride] = v.r; scr[(3 * slot + 1) * scr_stride] = v.g; scr[(3 * slot + 2) * scr_stride] = v.b;
    };
    auto unpark = [&](int slot) {
        return Rgb{scr[(3 * slot + 0) * scr_stride], scr[(3 * slot + 1) * scr_stride], scr[(3 * slot + 2) * scr_stride]};
    };
    const int lane = threadIdx.x & 31;
    const int pxi = lane & 7, pyi = lane >> 3;
    const int ns = g.nscales;
    const int W0 = g.sc[0].w, H0 = g.sc[0].h;
    auto plane_of = [&](int s) { return (size_t)g.sc[s].h * g.sc[s].pitch; };
    auto base_of = [&](int s) { return ximg + g.sc[s].xyb_off + (size_t)img * 3 * plane_of(s); };
    const bool vec_ok = FmtIs<FMT>::yuv &&
                        ((((uintptr_t)f.p0 | (uintptr_t)f.p1 | (uintptr_t)f.pitch) & 3u) == 0);
    Rgb l3b = Rgb{0.f, 0.f, 0.f}, l4b = l3b;   // level 3 / 4 pixels of pass 1 (pass 0's are parked in slots 3, 4)
    park(3, l3b);
    park(4, l3b);
    float* const gd0 = base_of(0);
    const size_t plane0 = plane_of(0);
    const int pitch0 = g.sc[0].pitch;

#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
        const int px0 = X0 + 4 * pxi, py0 = Y0 + 4 * (pyi + 4 * pass);
        Rgb l2 = Rgb{0.f, 0.f, 0.f};
        // the patch is walked as four 2x2 blocks (one copy of the colour / XYB code, 4 pixels = 12 cube roots in flight)
#pragma unroll 1
        for (int blk = 0; blk < 4; blk++) {
            const int x = px0 + 2 * (blk & 1), y = py0 + (blk & 2);
            Rgb lin[2][2];
            bool done = false;
            if constexpr (FmtIs<FMT>::yuv) {
                if (vec_ok && x + 2 <= W0 && y + 2 <= H0) {
                    // interior block: two luma samples per row in one load, one chroma pair for all four pixels
                    int Y00, Y01, Y10, Y11, cbi, cri;
                    if constexpr (FmtIs<FMT>::p016) {
                        const uint32_t a = __ldg(reinterpret_cast<const uint32_t*>(f.p0 + (size_t)y * f.pitch + 2 * x));
                        const uint32_t b = __ldg(reinterpret_cast<const uint32_t*>(f.p0 + (size_t)(y + 1) * f.pitch + 2 * x));
                        const uint32_t c = __ldg(reinterpret_cast<const uint32_t*>(f.p1 + (size_t)(y >> 1) * f.pitch + 2 * x));
                        Y00 = a & 0xffff; Y01 = a >> 16; Y10 = b & 0xffff; Y11 = b >> 16; cbi = c & 0xffff; cri = c >> 16;
                    } else {
                        const uint32_t a = __ldg(reinterpret_cast<const uint16_t*>(f.p0 + (size_t)y * f.pitch + x));
                        const uint32_t b = __ldg(reinterpret_cast<const uint16_t*>(f.p0 + (size_t)(y + 1) * f.pitch + x));
                        const uint32_t c = __ldg(reinterpret_cast<const uint16_t*>(f.p1 + (size_t)(y >> 1) * f.pitch + x));
                        Y00 = a & 0xff; Y01 = a >> 8; Y10 = b & 0xff; Y11 = b >> 8; cbi = c & 0xff; cri = c >> 8;
                    }
                    const YuvChroma ch = yuv_chroma<FMT>(cbi, cri, g.coef, g.eotf_lut, g.lut_n, g.lut_shift);
                    yuv_px(Y00, ch, g.coef, T, g.lut_shift, lin[0][0].r, lin[0][0].g, lin[0][0].b);
                    yuv_px(Y01, ch, g.coef, T, g.lut_shift, lin[0][1].r, lin[0][1].g, lin[0][1].b);
                    yuv_px(Y10, ch, g.coef, T, g.lut_shift, lin[1][0].r, lin[1][0].g, lin[1][0].b);
                    yuv_px(Y11, ch, g.coef, T, g.lut_shift, lin[1][1].r, lin[1][1].g, lin[1][1].b);
                    done = true;
                }
            }
            if (!done) {
                // generic path: per-pixel loads at clamped coordinates (edge blocks, packed RGB formats); rolled for the
                // YUV formats, where it only serves the frame edges
#pragma unroll(FmtIs<FMT>::yuv ? 1 : 4)
                for (int p = 0; p < 4; p++) {
                    Rgb o;
                    load_px<FMT>(f, min(x + (p & 1), W0 - 1), min(y + (p >> 1), H0 - 1), g.coef, T, g.eotf_lut, g.lut_n, g.lut_shift, o.r,
                                 o.g, o.b);
                    if (p == 0) lin[0][0] = o;
                    else if (p == 1) lin[0][1] = o;
                    else if (p == 2) lin[1][0] = o;
                    else lin[1][1] = o;
                }
            }
            // ---- scale 0: XYB of the four pixels
            {
                float X[2][2], Yv[2][2], B[2][2];
#pragma unroll
                for (int j = 0; j < 2; j++)
#pragma unroll
                    for (int i = 0; i < 2; i++) xyb_of<FMT>(lin[j][i].r, lin[j][i].g, lin[j][i].b, S, X[j][i], Yv[j][i], B[j][i]);
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    if (y + j < H0) {
                        float* q = gd0 + (size_t)(y + j) * pitch0 + x;
                        if (x + 2 <= W0) {
                            *reinterpret_cast<float2*>(q) = make_float2(X[j][0], X[j][1]);
                            *reinterpret_cast<float2*>(q + plane0) = make_float2(Yv[j][0], Yv[j][1]);
                            *reinterpret_cast<float2*>(q + 2 * plane0) = make_float2(B[j][0], B[j][1]);
                        } else if (x < W0) {
                            q[0] = X[j][0]; q[plane0] = Yv[j][0]; q[2 * plane0] = B[j][0];
                        }
                    }
                }
            }
            // ---- level 1 (clamped loads make the out-of-frame level-0 neighbours equal to the clamped ones already)
            const Rgb v1 = box_clamped(lin[0][0], lin[0][1], lin[1][0], lin[1][1], true, true);
            if (ns > 1) {
                const ScaleDesc& sd = g.sc[1];
                const int ox = x >> 1, oy = y >> 1;
                float X, Yv, B;
                xyb_of<FMT>(v1.r, v1.g, v1.b, S, X, Yv, B);
                if (ox < sd.w && oy < sd.h) {
                    const size_t plane = plane_of(1);
                    float* q = base_of(1) + (size_t)oy * sd.pitch + ox;
                    q[0] = X; q[plane] = Yv; q[2 * plane] = B;
                }
                // level 2 = box of the four level-1 pixels of the patch, in block order (0,0),(1,0),(0,1),(1,1)
                if (blk < 3) park(blk, v1);
                else l2 = box_clamped(unpark(0), unpark(1), unpark(2), v1, (px0 >> 1) + 1 < sd.w, (py0 >> 1) + 1 < sd.h);
            }
        }
        if (ns > 2) {
            const ScaleDesc& sd = g.sc[2];
            const int ox = px0 >> 2, oy = py0 >> 2;
            float X, Yv, B;
            xyb_of<FMT>(l2.r, l2.g, l2.b, S, X, Yv, B);
            if (ox < sd.w && oy < sd.h) {
                const size_t plane = plane_of(2);
                float* q = base_of(2) + (size_t)oy * sd.pitch + ox;
                q[0] = X; q[plane] = Yv; q[2 * plane] = B;
            }
            // level 3: lanes with even (pxi, pyi) own a pixel; neighbours by shuffle
            const Rgb b = shfl_rgb_xor(l2, 1), c = shfl_rgb_xor(l2, 8), d = shfl_rgb_xor(l2, 9);
            const Rgb v3 = box_clamped(l2, b, c, d, ox + 1 < sd.w, oy + 1 < sd.h);
            if (pass == 0) park(3, v3);
            l3b = v3;
        }
        if (ns > 3) {
            // level 4: lanes with pxi % 4 == 0, pyi == 0
            const ScaleDesc& sd = g.sc[3];
            const int ox = px0 >> 3, oy = py0 >> 3;
            const Rgb v = l3b;
            const Rgb b = shfl_rgb_xor(v, 2), c = shfl_rgb_xor(v, 16), d = shfl_rgb_xor(v, 18);
            const Rgb v4 = box_clamped(v, b, c, d, ox + 1 < sd.w, oy + 1 < sd.h);
            if (pass == 0) park(4, v4);
            l4b = v4;
        }
    }

    if (ns <= 3) return;
    const Rgb l3a = unpark(3), l4a = unpark(4);
    // ---- levels 3, 4, 5 of the region: 16 + 4 + 1 pixels, one XYB evaluation.
    //   lanes 0-7 and 16-23 (pyi 0 / 2): level 3; even pxi = pass 0's pixel of that lane, odd pxi = pass 1's of lane - 1
    //   lanes 8-11: level 4 (pass = bit 1, source lane = 4 * bit 0);  lane 12: level 5
    Rgb v = Rgb{0.f, 0.f, 0.f};
    int s = -1, ox = 0, oy = 0;
    {
        const Rgb up1 = shfl_rgb(l3b, lane - 1 < 0 ? 0 : lane - 1);
        if ((pyi & 1) == 0) {
            const int pass = pxi & 1;
            v = pass ? up1 : l3a;
            s = 3;
            ox = (X0 >> 3) + (pxi >> 1);
            oy = (Y0 >> 3) + (pyi >> 1) + 2 * pass;
        }
        const int src4 = 4 * (lane & 1);
        const Rgb a0 = shfl_rgb(l4a, src4), a1 = shfl_rgb(l4b, src4);
        if (lane >= 8 && lane < 12) {
            const int pass = (lane >> 1) & 1;
            v = pass ? a1 : a0;
            s = 4;
            ox = (X0 >> 4) + (lane & 1);
            oy = (Y0 >> 4) + pass;
        }
        if (ns > 5) {
            // level 5 from the four level-4 pixels: (0,0) pass 0 lane 0, (1,0) pass 0 lane 4, (0,1) pass 1 lane 0, (1,1) pass 1 lane 4
            const ScaleDesc& sd = g.sc[4];
            const Rgb p00 = shfl_rgb(l4a, 0), p10 = shfl_rgb(l4a, 4), p01 = shfl_rgb(l4b, 0), p11 = shfl_rgb(l4b, 4);
            if (lane == 12) {
                const int x4 = X0 >> 4, y4 = Y0 >> 4;
                v = box_clamped(p00, p10, p01, p11, x4 + 1 < sd.w, y4 + 1 < sd.h);
                s = 5;
                ox = X0 >> 5;
                oy = Y0 >> 5;
            }
        }
    }
    float X, Yv, B;
    xyb_of<FMT>(v.r, v.g, v.b, S, X, Yv, B);
    if (s >= 3 && s < ns) {
        const ScaleDesc& sd = g.sc[s];
        if (ox < sd.w && oy < sd.h) {
            const size_t plane = plane_of(s);
            float* q = base_of(s) + (size_t)oy * sd.pitch + ox;
            q[0] = X; q[plane] = Yv; q[2 * plane] = B;
        }
    }
}

// grid: (ceil(regions_x / (8 * kF2RegionsPerWarp)), regions_y, frames); warp w of a CTA owns kF2RegionsPerWarp regions along x
template <int FMT>
__global__ void __launch_bounds__(kF2Threads, KF2_MINB) k_frontend2(const __grid_constant__ Geo g, const FramePair* __restrict__ in,
                                                             float* __restrict__ xyb_base, int frame0)
{
    __shared__ FeTables tb;
    __shared__ float scratch[15 * kF2Threads];
    {
        const uint64_t* src = reinterpret_cast<const uint64_t*>(&kPowfTablesInit);
        uint64_t* dst = reinterpret_cast<uint64_t*>(&tb.T);
        for (int i = threadIdx.x; i < (int)(sizeof(exact_math::PowfTables) / 8); i += kF2Threads) dst[i] = src[i];
        for (int i = threadIdx.x; i < 256; i += kF2Threads) tb.S.tab[i] = exact_math::cbrt_scale_entry(i);
        if constexpr (FMT == kNV12 || FMT == kP016) {
            // the per-sample integer -> float work of yuv_px / yuv_chroma, once per code (same expressions => same bits)
            for (int c = threadIdx.x; c < g.lut_n; c += kF2Threads) {
                const int v = c << g.lut_shift;
                tb.luma[c] = (float)(max(v, g.coef.luma_min) - g.coef.luma_min) * g.coef.y;
                tb.chroma[c] = (float)(v - g.coef.neutral);
            }
        } else if constexpr (FMT == kSRGB8) {
            for (int i = threadIdx.x; i < 256; i += kF2Threads) tb.luma[i] = kSrgb8Lut[i];
        }
    }
    __syncthreads();
    const int frame = frame0 + blockIdx.z, warp = threadIdx.x >> 5;
    const int rx_n = (g.sc[0].w + kF2Region - 1) / kF2Region;
    const FramePair fp = in[frame];
    float* xyb_slot = xyb_base + (size_t)frame * g.xyb_stride;
    const int Y0 = blockIdx.y * kF2Region;
    const int rx0 = (blockIdx.x * (kF2Threads / 32) + warp) * kF2RegionsPerWarp;
    const bool rows_inside = Y0 + kF2Region <= g.sc[0].h;
    if constexpr (FastFmt<FMT>::ok) {
        if (rows_inside && (rx0 + 1) * kF2Region <= g.sc[0].w) prefetch_region<FMT>(fp.ref, rx0 * kF2Region, Y0);
    }
#pragma unroll 1
    for (int i = 0; i < kF2RegionsPerWarp; i++) {
        const int rx = rx0 + i;
        if (rx >= rx_n) break;
        const int X0 = rx * kF2Region;
#pragma unroll 1
        for (int img = 0; img < 2; img++) {
            const FrameIn& f = img ? fp.dis : fp.ref;
            bool done = false;
            if constexpr (FastFmt<FMT>::ok) {
                // the next work item of this warp (the other image of this region, then the next region) one step ahead
                if (img == 0) {
                    if (rows_inside && X0 + kF2Region <= g.sc[0].w) prefetch_region<FMT>(fp.dis, X0, Y0);
                } else if (i + 1 < kF2RegionsPerWarp && rows_inside && X0 + 2 * kF2Region <= g.sc[0].w) {
                    prefetch_region<FMT>(fp.ref, X0 + kF2Region, Y0);
                }
                // interior region, aligned rows, (YUV) the exact R / B memo present
                const uint32_t align = FastFmt<FMT>::align;
                const bool fast = X0 + kF2Region <= g.sc[0].w && Y0 + kF2Region <= g.sc[0].h &&
                                  ((((uintptr_t)f.p0 | (uintptr_t)f.p1 | (uintptr_t)f.pitch) & align) == 0) &&
                                  (!FastFmt<FMT>::needs_lut || g.eotf_lut != nullptr);
                if (fast) done = frontend_region_fast<FMT>(g, f, xyb_slot, img, X0, Y0, tb, scratch + threadIdx.x, kF2Threads);
            }
            if (!done) frontend_region<FMT>(g, f, xyb_slot, img, X0, Y0, tb.T, tb.S, scratch + threadIdx.x, kF2Threads);
        }
    }
}

// ---- TMA / mbarrier plumbing (PTX; sm_90+ instructions, SASS: UTMALDG / SYNCS) -------------------
// Tensor maps of one batch slot, per scale.  All are 4-D {x, y, plane, frame} views of [frame][plane][h][pitch] f32.
struct alignas(64) TmaMaps {
    CUtensorMap hb[kMaxScales];      // V pass load : 15 planes, box {64, 2, 15, 1}, no swizzle
    CUtensorMap xyb[kMaxScales];     // V pass load :  6 planes, box {64, 2, 6, 1},  no swizzle
};
struct alignas(64) TmaMapsX {
    CUtensorMap xyb_in[kMaxScales];  // k_hv load   :  6 planes, box {76, 12, 6, 1},  no swizzle
};
struct alignas(64) TmaMapsH {
    CUtensorMap xyb_in[kMaxScales];  // H pass load :  6 planes, box {32, 32, 6, 1},  128B swizzle
    CUtensorMap hb_out[kMaxScales];  // H pass store: 15 planes, box {32, 32, 15, 1}, 128B swizzle
};

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// try_wait suspends the thread in hardware for up to the hinted time before it reports failure, so a waiting
// warp does not burn issue slots of the warps that have work.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(20000u)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3)
{
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------------------------
// k_hpass: products + horizontal recursive Gaussian, TMA in / TMA out.
// Replaces nppiMul x3 (ssimulacra2-cuda/src/lib.rs:300-317) and one of the two
// blur_plane_pass_fused launches + its nppiTranspose x5 (lib.rs:328-361).
// Follows image_multiply cpu.rs:537-543 and RecursiveGaussian::horizontal_row cpu.rs:967-1022.
//
// CTA = one 32-row band of one scale of one frame; 15 scan warps + 1 control warp, 2 CTAs per SM.
//   control : one lane streams 6 x 32 x 32 XYB chunks into a 2-stage shared-memory ring with
//             cp.async.bulk.tensor (128B swizzle; columns < 0 or >= W arrive as zeros = the filter's zero
//             padding) and writes each finished 15 x 32 x 32 output tile back with a TMA store (clipped at the
//             image edge by the hardware)
//   scan    : warp p (0..14) = plane (quantity q = p/3, channel c = p%3), lane = row; walks the chunk 4 columns
//             per 128-bit shared-memory access (swizzle => conflict-free), filter state in registers across chunks
// Step t consumes x[t] (right tap, index n+4) and x[t-10] (left tap, n-6) and emits y[t-4]
// (n = t-4, cpu.rs:976-984).  Chunk k (k = -1, 0, ...) covers steps 32k+4 .. 32k+35, i.e. outputs
// 32k .. 32k+31; chunk -1 only warms the state with x[0..3] (its other inputs are the zero padding).
// ------------------------------------------------------------------------------------------
constexpr int kHRows = 32;
constexpr int kHCols = 32;
constexpr int kHScanWarps = 15;
constexpr int kHThreads = (kHScanWarps + 1) * 32;
constexpr int kHInStages = 2;
constexpr uint32_t kHInBytes = 6 * kHRows * kHCols * sizeof(float);    // 24576
constexpr uint32_t kHOutBytes = 15 * kHRows * kHCols * sizeof(float);  // 61440
constexpr size_t kHSmemBytes = (size_t)kHInStages * kHInBytes + kHOutBytes + 128 + 64;  // tiles + ones row + mbarriers

struct HState {
    float p1, p3, p5, pp1, pp3, pp5;
};

__device__ __forceinline__ float hstep(HState& s, float left, float right)
{
    float sum = left + right;
    float o1 = sum * RG_IN_1, o3 = sum * RG_IN_3, o5 = sum * RG_IN_5;
    o1 = o1 - s.pp1;  // == fmaf(MUL_PREV2 = -1, prev2, out): the product is exact
    o3 = o3 - s.pp3;
    o5 = o5 - s.pp5;
    s.pp1 = s.p1; s.pp3 = s.p3; s.pp5 = s.p5;
    o1 = fmaf(RG_PREV_1, s.p1, o1);
    o3 = fmaf(RG_PREV_3, s.p3, o3);
    o5 = fmaf(RG_PREV_5, s.p5, o5);
    s.p1 = o1; s.p3 = o3; s.p5 = o5;
    return (o1 + o3) + o5;
}

// One code path for all five quantities (keeps the kernel inside the instruction cache): the filter input is
// x[j] * y[j] with (x, y) = (ref, ref), (dis, dis), (ref, dis), (ref, 1), (dis, 1); multiplying by 1.0f is exact, so
// the mu planes are filtered on exactly the values the reference filters (cpu.rs:396-399).
// s_x / s_y / s_out are shared-space byte addresses of this lane's 128-byte row; xs = (lane & 7) << 4 is the
// 128B-swizzle term.
__device__ __forceinline__ float4 lds128(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v)
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ void hscan_chunk(uint32_t s_x, uint32_t s_y, uint32_t s_out, uint32_t xs, HState& st,
                                            float (&hist)[12])
{
    float win[12 + kHCols];
#pragma unroll
    for (int i = 0; i < 12; i++) win[i] = hist[i];
#pragma unroll
    for (int gI = 0; gI < kHCols / 4; gI++) {
        const uint32_t off = ((uint32_t)gI << 4) ^ xs;
        const float4 a = lds128(s_x + off), d = lds128(s_y + off);
        const float pr[4] = {a.x * d.x, a.y * d.y, a.z * d.z, a.w * d.w};
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            win[12 + 4 * gI + j] = pr[j];
            o[j] = hstep(st, win[4 * gI + j + 2], pr[j]);  // left tap = 10 columns back
        }
        sts128(s_out + off, make_float4(o[0], o[1], o[2], o[3]));
    }
#pragma unroll
    for (int i = 0; i < 12; i++) hist[i] = win[kHCols + i];
}

// dynamic shared memory: [2 in stages][out tile][ones row 128 B][mbarriers]
constexpr uint32_t kHOffOut = kHInStages * kHInBytes;
constexpr uint32_t kHOffOnes = kHOffOut + kHOutBytes;
constexpr uint32_t kHOffBars = kHOffOnes + 128;

// One 32-row band (work item `item` of the scale-ordered list) of frame `frame`; 512 threads; hs = 1024-aligned smem.
__device__ __forceinline__ void hpass_band(const Geo& g, const TmaMapsH& maps, int item, int frame, char* hs)
{
    const uint32_t sbase = smem_u32(hs);
    if ((sbase & 1023u) != 0) __trap();  // the 128B swizzle pattern is anchored at 1024-byte boundaries
    uint64_t* bars = reinterpret_cast<uint64_t*>(hs + kHOffBars);
    uint64_t* full_in = bars;            // [2]
    uint64_t* empty_in = bars + 2;       // [2]
    uint64_t* scan_done = bars + 4;
    uint64_t* out_free = bars + 5;

    int s = 0;
    while (item >= g.sc[s].n_bands) { item -= g.sc[s].n_bands; s++; }
    const int W = g.sc[s].w;
    const int row0 = item * kHRows;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nchunks = (W + kHCols - 1) / kHCols + 1;

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < kHInStages; i++) {
            mbar_init(&full_in[i], 1);
            mbar_init(&empty_in[i], kHScanWarps);
        }
        mbar_init(scan_done, kHScanWarps);
        mbar_init(out_free, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) reinterpret_cast<float*>(hs + kHOffOnes)[tid] = 1.0f;
    __syncthreads();

    if (warp == kHScanWarps) {
        // ===== control warp =====
        if (lane == 0) {
            const CUtensorMap* min = &maps.xyb_in[s];
            const CUtensorMap* mout = &maps.hb_out[s];
            auto issue_load = [&](int kk) {  // chunk k = kk - 1 into stage kk % 2
                const int st = kk % kHInStages;
                if (kk >= kHInStages) mbar_wait(&empty_in[st], (uint32_t)(((kk / kHInStages) - 1) & 1));
                mbar_expect_tx(&full_in[st], kHInBytes);
                tma_load_4d(hs + st * kHInBytes, min, &full_in[st], kHCols * (kk - 1) + 4, row0, 0, frame);
            };
            issue_load(0);
            for (int kk = 0; kk < nchunks; kk++) {
                if (kk + 1 < nchunks) issue_load(kk + 1);
                mbar_wait(scan_done, (uint32_t)(kk & 1));  // the scan warps have written tile kk (generic proxy, fenced)
                if (kk >= 1) {
                    tma_store_4d(mout, hs + kHOffOut, kHCols * (kk - 1), row0, 0, frame);
                    tma_store_commit();
                    tma_store_wait_read();                 // shared memory may be overwritten again
                }
                mbar_arrive(out_free);
            }
            tma_store_wait_all();
        }
        return;
    }

    // ===== scan warps =====
    const int q = warp / 3, ch = warp - 3 * q;
    const uint32_t xs = (uint32_t)(lane & 7) << 4;
    HState st = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float hist[12];
#pragma unroll
    for (int i = 0; i < 12; i++) hist[i] = 0.f;
    // plane of the first / second factor inside an input stage (0..2 = ref X,Y,B; 3..5 = dis X,Y,B); second < 0: ones
    const int px = (q == 1 || q == 4) ? 3 + ch : ch;
    const int py = (q == 0) ? ch : ((q == 1 || q == 2) ? 3 + ch : -1);
    const uint32_t row_x = (uint32_t)((px * kHRows + lane) * 128);
    const uint32_t row_y = (uint32_t)(((py < 0 ? 0 : py) * kHRows + lane) * 128);
    const uint32_t so = sbase + kHOffOut + (uint32_t)((warp * kHRows + lane) * 128);

    for (int kk = 0; kk < nchunks; kk++) {
        const int stg = kk % kHInStages;
        mbar_wait(&full_in[stg], (uint32_t)((kk / kHInStages) & 1));
        if (kk >= 1) mbar_wait(out_free, (uint32_t)((kk - 1) & 1));  // tile kk-1 has left shared memory
        const uint32_t sin = sbase + stg * kHInBytes;
        // the ones row is not swizzled data, any 16-byte chunk of it will do: cancel the swizzle term
        hscan_chunk(sin + row_x, py < 0 ? (sbase + kHOffOnes) : (sin + row_y), so, xs, st, hist);
        fence_proxy_async();  // make this lane's tile writes visible to the TMA store
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(scan_done);
            mbar_arrive(&empty_in[stg]);
        }
    }
}

__global__ void __launch_bounds__(kHThreads, 2) k_hpass(const __grid_constant__ Geo g, const __grid_constant__ TmaMapsH maps)
{
    extern __shared__ __align__(1024) char hsm[];
    hpass_band(g, maps, blockIdx.x, blockIdx.y, hsm);
}

// ------------------------------------------------------------------------------------------
// k_vpass: vertical recursive Gaussian + error maps + partial sums.
// Replaces the second blur_plane_pass_fused launch, 2 nppiTranspose, compute_error_maps
// (ssimulacra2-cuda-kernel/src/error_maps.rs:4-60) and the 6 x {nppiSum, nppiSqr, nppiSqr_I, nppiSum}
// reductions (ssimulacra2-cuda/src/lib.rs:372-447).  Follows vertical_pass cpu.rs:1054-1115,
// ssim_map cpu.rs:581-638 and edge_diff_map cpu.rs:640-683 (f64 tails included).
//
// CTA = one 64-column strip of one scale of one frame; 192 threads = (channel c, column x);
// each thread runs the 5 filters of its (c, x) down the column, evaluates the three maps for its
// channel at every output row and accumulates 6 f64 sums.  Rows are processed kVUnroll per
// iteration so that 7*kVUnroll independent loads are in flight per thread; the 10-row delay line
// of each filter is a per-thread ring in shared memory (no barriers in the main loop).
// ------------------------------------------------------------------------------------------
constexpr int kVCols = 64;
constexpr int kVRing = 10;

// Correctly rounded f32 quotient for operands in the normal range (no zero / inf / nan / subnormal
// handling): reciprocal seed + one Newton step + Markstein's residual correction.  Checked against
// IEEE division on the device by tests/test_gpu_parity.py::test_device_division.
__device__ __forceinline__ float div_rn_normal(float n, float d)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
    float e = fmaf(-d, r, 1.0f);
    r = fmaf(r, e, r);
    float q = n * r;
    float rem = fmaf(-d, q, n);
    return fmaf(rem, r, q);
}

// vstep / div_rn_normal / error_maps on two adjacent columns at once (same operations, same order).
struct VState2 {
    f2 p1, p3, p5, pp1, pp3, pp5;
};

__device__ __forceinline__ f2 vstep2(VState2& s, f2 top, f2 bottom)
{
    const f2 sum = f2_add(top, bottom);
    const f2 a1 = f2_fma(s.p1, f2_splat(-RG_PREV_1), s.pp1);
    const f2 a3 = f2_fma(s.p3, f2_splat(-RG_PREV_3), s.pp3);
    const f2 a5 = f2_fma(s.p5, f2_splat(-RG_PREV_5), s.pp5);
    const f2 o1 = f2_fma(sum, f2_splat(RG_IN_1), f2_neg(a1));
    const f2 o3 = f2_fma(sum, f2_splat(RG_IN_3), f2_neg(a3));
    const f2 o5 = f2_fma(sum, f2_splat(RG_IN_5), f2_neg(a5));
    s.pp1 = s.p1; s.pp3 = s.p3; s.pp5 = s.p5;
    s.p1 = o1; s.p3 = o3; s.p5 = o5;
    return f2_add(f2_add(o1, o3), o5);
}

__device__ __forceinline__ f2 div_rn_normal2(f2 n, f2 d)
{
    const f2 one = f2_splat(1.0f);
    f2 r = f2_rcp_approx(d);
    const f2 nd = f2_neg(d);
    const f2 e = f2_fma(nd, r, one);
    r = f2_fma(r, e, r);
    const f2 q = f2_mul(n, r);
    const f2 rem = f2_fma(nd, q, n);
    return f2_fma(rem, r, q);
}

// The three error maps of two adjacent pixels of one channel and their contributions to the six sums.
// SSIM' term: cpu.rs:604-631 -- identical f32 arithmetic up to the quotient q; d = 1 - q is then
// formed in f32 (exact whenever q is in [0.5, 2], i.e. wherever d is small) instead of f64.
// Edge terms: cpu.rs:658-674 computes (1+|dis-mu2|)/(1+|ref-mu1|) - 1 in f64; here the algebraically
// equal (a-b)/(1+b) in f32, which keeps a RELATIVE error of ~2e-7 on every term (the f32 form of the
// original expression would not).  Nothing downstream amplifies these errors: they enter the sums
// directly, 3 orders of magnitude under the 1e-4 bar.
// art = max(d1, 0), detail = max(-d1, 0); their 4th powers are formed from the clamped values (identical to
// selecting d1^4 by the sign of d1: one of the two is exactly zero).
__device__ __forceinline__ void error_maps2(const f2 (&o)[5], f2 ref, f2 dis, f2 (&part)[6])
{
    const f2 C2 = f2_splat(0.0009f), one = f2_splat(1.0f);
    const f2 s11 = o[0], s22 = o[1], s12 = o[2], mu1 = o[3], mu2 = o[4];
    const f2 mu11 = f2_mul(mu1, mu1), mu22 = f2_mul(mu2, mu2), mu12 = f2_mul(mu1, mu2);
    const f2 mu_diff = f2_sub(mu1, mu2);
    const f2 num_m = f2_fma(mu_diff, f2_neg(mu_diff), one);
    const f2 num_s = f2_fma(f2_splat(2.0f), f2_sub_prod(s12, mu12), C2);
    const f2 denom_s = f2_add(f2_add(f2_sub_prod(s11, mu11), f2_sub_prod(s22, mu22)), C2);
    const f2 q = div_rn_normal2(f2_mul(num_m, num_s), denom_s);
    const f2 d = f2_max0(f2_sub(one, q));
    part[0] = f2_add(part[0], d);
    const f2 d2 = f2_mul(d, d);
    part[1] = f2_fma(d2, d2, part[1]);

    const f2 a = f2_abs(f2_sub(dis, mu2)), b = f2_abs(f2_sub(ref, mu1));
    const f2 den = f2_add(one, b);
    f2 r = f2_rcp_approx(den);
    r = f2_fma(r, f2_fma(f2_neg(den), r, one), r);
    const f2 d1 = f2_mul(f2_sub(a, b), r);
    const f2 art = f2_max0(d1), det = f2_max0(f2_neg(d1));
    const f2 art2 = f2_mul(art, art), det2 = f2_mul(det, det);
    part[2] = f2_add(part[2], art);
    part[3] = f2_fma(art2, art2, part[3]);
    part[4] = f2_add(part[4], det);
    part[5] = f2_fma(det2, det2, part[5]);
}

// The two halves of error_maps2 for k_hv, where the SSIM map and the edge maps of a channel run in different warps:
// the same operations in the same order on the same operands.
__device__ __forceinline__ void ssim_map2(f2 s11, f2 s22, f2 s12, f2 mu1, f2 mu2, f2 (&part)[2])
{
    const f2 C2 = f2_splat(0.0009f), one = f2_splat(1.0f);
    const f2 mu11 = f2_mul(mu1, mu1), mu22 = f2_mul(mu2, mu2), mu12 = f2_mul(mu1, mu2);
    const f2 mu_diff = f2_sub(mu1, mu2);
    const f2 num_m = f2_fma(mu_diff, f2_neg(mu_diff), one);
    const f2 num_s = f2_fma(f2_splat(2.0f), f2_sub_prod(s12, mu12), C2);
    const f2 denom_s = f2_add(f2_add(f2_sub_prod(s11, mu11), f2_sub_prod(s22, mu22)), C2);
    const f2 q = div_rn_normal2(f2_mul(num_m, num_s), denom_s);
    const f2 d = f2_max0(f2_sub(one, q));
    part[0] = f2_add(part[0], d);
    const f2 d2 = f2_mul(d, d);
    part[1] = f2_fma(d2, d2, part[1]);
}
__device__ __forceinline__ void edge_maps2(f2 mu1, f2 mu2, f2 ref, f2 dis, f2 (&part)[4])
{
    const f2 one = f2_splat(1.0f);
    const f2 a = f2_abs(f2_sub(dis, mu2)), b = f2_abs(f2_sub(ref, mu1));
    const f2 den = f2_add(one, b);
    f2 r = f2_rcp_approx(den);
    r = f2_fma(r, f2_fma(f2_neg(den), r, one), r);
    const f2 d1 = f2_mul(f2_sub(a, b), r);
    const f2 art = f2_max0(d1), det = f2_max0(f2_neg(d1));
    const f2 art2 = f2_mul(art, art), det2 = f2_mul(det, det);
    part[0] = f2_add(part[0], art);
    part[1] = f2_fma(art2, art2, part[1]);
    part[2] = f2_add(part[2], det);
    part[3] = f2_fma(det2, det2, part[3]);
}

// k_vpass, TMA-fed.  CTA = one 64-column strip of one scale of one frame; 3 consumer warps (warp = channel, lane =
// a PAIR of adjacent columns, all arithmetic packed f32x2) + 1 producer warp.  The producer streams
// {hb rows t, t+1 ; xyb rows t-4, t-3} boxes into a 5-stage shared-memory ring with cp.async.bulk.tensor (out-of-range
// rows / columns arrive as zeros = the filter's zero padding, so the loop has no edge cases; an all-zero column
// contributes exactly 0 to every sum); full/empty mbarriers per stage.  Each consumer thread runs its 5 filters down
// its two columns: the 10-row delay line lives in REGISTERS (the row loop is unrolled by 10 = one turn of the ring),
// shared memory is read once per value with 64-bit loads, there are no block-wide barriers and no global-memory
// instructions in the loop.
constexpr int kVRowsPerStage = 2;
constexpr int kVStages = kVRing / kVRowsPerStage;                  // 5 stages = one 10-row group
constexpr int kVBoxHb = 15 * kVRowsPerStage * kVCols;              // floats
constexpr int kVBoxXyb = 6 * kVRowsPerStage * kVCols;
constexpr int kVStageFloats = kVBoxHb + kVBoxXyb;                  // 2688 floats = 10752 B
constexpr uint32_t kVStageBytes = kVStageFloats * sizeof(float);
constexpr int kVConsumers = 3 * (kVCols / 2);                      // 96
constexpr int kVTmaThreads = kVConsumers + 32;                     // + producer warp
constexpr size_t kVSmemBytes = (size_t)kVStages * kVStageBytes + 128;

__global__ void __launch_bounds__(kVTmaThreads, 2) k_vpass(const __grid_constant__ Geo g, const __grid_constant__ TmaMaps maps,
                                                           double* __restrict__ partials)
{
    extern __shared__ __align__(128) float vs[];
    __shared__ uint64_t full_bar[kVStages], empty_bar[kVStages];

    const int frame = blockIdx.y;
    int item = blockIdx.x, s = 0;
    while (item >= g.sc[s].n_strips) { item -= g.sc[s].n_strips; s++; }
    const ScaleDesc sd = g.sc[s];
    const int H = sd.h;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ngroups = (H + 4 + kVRing - 1) / kVRing;

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < kVStages; i++) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], kVConsumers / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == kVConsumers / 32) {
        // ===== producer warp: one elected lane issues the TMA loads =====
        if (lane == 0) {
            const CUtensorMap* mhb = &maps.hb[s];
            const CUtensorMap* mxyb = &maps.xyb[s];
            const int x0 = item * kVCols;
            for (int gi = 0; gi < ngroups; gi++) {
#pragma unroll 1
                for (int st = 0; st < kVStages; st++) {
                    if (gi > 0) mbar_wait(&empty_bar[st], (uint32_t)((gi - 1) & 1));
                    float* dst = vs + st * kVStageFloats;
                    const int row = gi * kVRing + st * kVRowsPerStage;
                    mbar_expect_tx(&full_bar[st], kVStageBytes);
                    tma_load_4d(dst, mhb, &full_bar[st], x0, row, 0, frame);
                    tma_load_4d(dst + kVBoxHb, mxyb, &full_bar[st], x0, row - 4, 0, frame);
                }
            }
        }
        return;
    }

    // ===== consumers: warp = channel, lane = column pair =====
    const int c = warp;
    const uint32_t sbase = smem_u32(vs) + (uint32_t)lane * 8u;
    VState2 stq[5];
    f2 dl[kVRing][5];  // delay line: the input of 10 rows ago, per quantity
    f2 zero2 = f2_splat(0.0f);
#pragma unroll
    for (int qi = 0; qi < 5; qi++) stq[qi] = VState2{zero2, zero2, zero2, zero2, zero2, zero2};
#pragma unroll
    for (int i = 0; i < kVRing; i++)
#pragma unroll
        for (int qi = 0; qi < 5; qi++) dl[i][qi] = zero2;
    double acc[6] = {0, 0, 0, 0, 0, 0};

    for (int gi = 0; gi < ngroups; gi++) {
        const uint32_t parity = (uint32_t)(gi & 1);
        f2 part[6] = {zero2, zero2, zero2, zero2, zero2, zero2};
#pragma unroll
        for (int st = 0; st < kVStages; st++) {
            mbar_wait(&full_bar[st], parity);
            const uint32_t sb = sbase + (uint32_t)st * kVStageBytes;
#pragma unroll
            for (int r = 0; r < kVRowsPerStage; r++) {
                const int slot = st * kVRowsPerStage + r;
                const int t = gi * kVRing + slot;
                f2 o[5];
#pragma unroll
                for (int qi = 0; qi < 5; qi++) {
                    const f2 v = lds64(sb + (uint32_t)(((qi * 3 + c) * kVRowsPerStage + r) * kVCols) * 4u);
                    o[qi] = vstep2(stq[qi], dl[slot][qi], v);
                    dl[slot][qi] = v;
                }
                const f2 fr = lds64(sb + (uint32_t)(kVBoxHb + ((0 + c) * kVRowsPerStage + r) * kVCols) * 4u);
                const f2 fd = lds64(sb + (uint32_t)(kVBoxHb + ((3 + c) * kVRowsPerStage + r) * kVCols) * 4u);
                if (t >= 4 && t < H + 4) error_maps2(o, fr, fd, part);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[st]);
            if (st == kVStages / 2) {
#pragma unroll
                for (int k = 0; k < 6; k++) { acc[k] += (double)f2_hsum(part[k]); part[k] = zero2; }
            }
        }
#pragma unroll
        for (int k = 0; k < 6; k++) acc[k] += (double)f2_hsum(part[k]);
    }

    // one warp = one channel: reduce over the 32 column pairs and write the strip's partial sums
#pragma unroll
    for (int k = 0; k < 6; k++) {
        double vsum = acc[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) vsum += __shfl_down_sync(0xffffffffu, vsum, off);
        if (lane == 0) partials[((size_t)frame * g.total_strips + sd.strip0 + item) * 18 + c * 6 + k] = vsum;
    }
}

// ------------------------------------------------------------------------------------------
// k_hv: horizontal pass, vertical pass, error maps and partial sums in ONE kernel -- the 15 H-pass planes
// (60 B per pyramid pixel, written and read back by k_hpass / k_vpass) never leave the SM.
// Same arithmetic, operation for operation, as k_hpass + k_vpass (cpu.rs:967-1022, :1054-1115, :581-683).
//
// CTA = one 64-column strip of one scale of one frame, walked top to bottom in 12-row bands; 7 warps:
//   P  (1 warp) : TMA-loads the XYB tile of band j (6 planes x 12 rows x 76 columns: 8 columns of left halo for the
//                 10-tap history, 4 of right halo for the look-ahead tap) into a 3-deep ring, and fetches the
//                 horizontal filter state the strip to the LEFT left behind for band j (see below)
//   H  (3 warps): warp = channel, lane = (quantity, row pair): 5 x 6 = 30 lanes, two rows per lane in packed f32x2;
//                 scans the 64 columns of the band, writes the 15-plane tile of the band into a 3-deep ring
//   V  (3 warps): warp = channel, lane = column pair: the five vertical filters + error maps + sums of k_vpass;
//                 the 10-row delay line is read back from the tile ring (this band's tile and the previous one)
// The horizontal recursion runs along the whole row, so strip k of a band continues from the filter state strip
// k-1 reached at its right edge: 6 floats per (plane, row), published through global memory ([6][96] f2 per
// band, 4.6 KB) with a release flag per (strip, band); the halo columns supply the 10 products of history.
// Strips of one chain (frame, scale) therefore run as a systolic wavefront, one band apart.  Work items are
// handed out by an atomic ticket in dependency order (strip-major), so a CTA only ever waits for CTAs that are
// already running: no deadlock regardless of how many CTAs are resident.
// ------------------------------------------------------------------------------------------
constexpr int kXR = 12;                                  // rows per band
constexpr int kXC = kVCols;                              // columns per strip (the strip list is the V pass's)
constexpr int kXInLead = 8;                              // tile starts at column x0 - 8
constexpr int kXInW = 76;                                // x0 - 8 .. x0 + 67
constexpr int kXInPlane = kXR * kXInW;                   // 912 floats
constexpr int kXInFloats = 6 * kXInPlane;
constexpr uint32_t kXInBytes = kXInFloats * 4;           // 21888
constexpr int kXHbPitch = 68;                            // floats; with the plane pad: conflict-free 128-bit stores
constexpr int kXHbPlane = kXR * kXHbPitch + 8;           // 824 floats
constexpr int kXHbFloats = 15 * kXHbPlane;
constexpr uint32_t kXHbBytes = kXHbFloats * 4;           // 49440
constexpr int kXHThreads = 96;
constexpr int kXWarps = 16;                              // roles by warp id, see k_hv
constexpr int kXNIn = 3;                                 // depth of the XYB tile ring
constexpr int kXThreads = kXWarps * 32;
constexpr int kXHsF2 = 6 * kXHThreads;                   // hand-off record: [6 state words][96 H threads] f2
constexpr uint32_t kXHsBytes = kXHsF2 * 8;               // 4608
constexpr int kXSub = 4;                                 // rows per mu hand-off between the Vb and Va warps of a channel
constexpr int kXMuSlotF = kXSub * 2 * kXC;               // [4 rows][mu1, mu2][64 columns] floats
constexpr uint32_t kXMuSlotBytes = kXMuSlotF * 4;        // 2048
constexpr uint32_t kXOffIn = 0;
constexpr uint32_t kXOffHb = kXOffIn + kXNIn * kXInBytes;
constexpr uint32_t kXOffMu = kXOffHb + 3 * kXHbBytes;    // [3 channels][2 slots] mu hand-off
constexpr uint32_t kXOffOnes = (kXOffMu + 6 * kXMuSlotBytes + 127) / 128 * 128;  // rows of ones (second factor of the mu planes)
constexpr uint32_t kXOnesBytes = 384;
constexpr uint32_t kXOffBars = kXOffOnes + kXOnesBytes;
constexpr int kXNumBars = 17 + 12 + 3;
constexpr size_t kXSmemBytes = kXOffBars + kXNumBars * 8;
static_assert(kXOffHb % 16 == 0 && kXHbBytes % 16 == 0 && kXInBytes % 128 == 0 && kXOffIn % 128 == 0 && kXOffOnes % 128 == 0 &&
                  kXOffBars % 8 == 0, "k_hv smem layout");
static_assert(kXSmemBytes <= 232448, "k_hv shared memory");
static_assert(kXR % kXSub == 0, "k_hv sub-bands");

// mbarrier wait with a watchdog: a protocol bug must end in a trap, not in a hung GPU
__device__ __forceinline__ void mbar_wait_wd(uint64_t* bar, uint32_t parity)
{
    const uint32_t a = smem_u32(bar);
    for (uint32_t it = 0;; it++) {
        uint32_t ok;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(a), "r"(parity), "r"(20000u)
            : "memory");
        if (ok) return;
        if (it > 400000u) __trap();
    }
}
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ f2 ldcg64(const void* p)
{
    f2 r;
    asm volatile("ld.global.cg.b64 %0, [%1];" : "=l"(r.v) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_cta_shared(uint32_t addr, uint32_t v)
{
    asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_cta_shared(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
// named barriers (ids 1..15): a hardware wait costs no issue slots, unlike a try_wait loop -- used for the hand-off between
// the two V warps of a channel, which share their sub-partition with the H warps
__device__ __forceinline__ void nbar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void nbar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void sts64(uint32_t addr, f2 v) { asm volatile("st.shared.b64 [%0], %1;" ::"r"(addr), "l"(v.v) : "memory"); }

// hstep on two rows at once (see f2_sub_prod for the form of the subtraction).
struct HState2 {
    f2 p1, p3, p5, pp1, pp3, pp5;
};
__device__ __forceinline__ f2 hstep2(HState2& s, f2 left, f2 right)
{
    const f2 sum = f2_add(left, right);
    f2 o1 = f2_mul(sum, f2_splat(RG_IN_1)), o3 = f2_mul(sum, f2_splat(RG_IN_3)), o5 = f2_mul(sum, f2_splat(RG_IN_5));
    // t = pp - o (as fma(o, -1, pp)); the wanted o - pp is its exact negation, folded into the next FMA's operand
    const f2 t1 = f2_sub_prod(s.pp1, o1), t3 = f2_sub_prod(s.pp3, o3), t5 = f2_sub_prod(s.pp5, o5);
    s.pp1 = s.p1; s.pp3 = s.p3; s.pp5 = s.p5;
    o1 = f2_fma(f2_splat(RG_PREV_1), s.p1, f2_neg(t1));
    o3 = f2_fma(f2_splat(RG_PREV_3), s.p3, f2_neg(t3));
    o5 = f2_fma(f2_splat(RG_PREV_5), s.p5, f2_neg(t5));
    s.p1 = o1; s.p3 = o3; s.p5 = o5;
    return f2_add(f2_add(o1, o3), o5);
}

struct HvArgs {
    f2* hstate;                 // [frames][total_recs][6][96]
    uint32_t* flags;            // [frames][total_recs], == epoch once the record is published
    uint32_t* ticket;           // work-item counter (reset by k_finalize)
    double* partials;
    uint32_t epoch;
    int nframes;
    // SSIMU2_FLAG_SCORE_ONLY: per scale, the channels whose two SSIM' weights are both zero (bit c).  Their s11 / s22 /
    // s12 filters and SSIM' map cannot change the score.  The mask 0b101 (X and B: what the weights give at scale 0, 75 % of
    // all pixels) has its own warp-role map ("lite" strips, hv_role); any other mask runs the full kernel.
    // (packed 4 bits per scale: a dynamically indexed array in the parameter block would be copied to local memory)
    uint32_t lite_bits;
    __device__ __forceinline__ uint32_t lite(int s) const { return (lite_bits >> (4 * s)) & 7u; }
};

// the four column groups (16 tile columns) of one H-scan iteration: products of both rows, 16 filter steps, stores
struct HGroupLoad {
    float4 xa, xb, ya, yb;
};
__device__ __forceinline__ HGroupLoad h_load(uint32_t axA, uint32_t axB, uint32_t ayA, uint32_t ayB, uint32_t off)
{
    HGroupLoad L;
    L.xa = lds128(axA + off); L.xb = lds128(axB + off); L.ya = lds128(ayA + off); L.yb = lds128(ayB + off);
    return L;
}
__device__ __forceinline__ void h_products(const HGroupLoad& L, f2 (&w)[16], int at)
{
    w[(at + 0) & 15] = f2_pack(L.xa.x * L.ya.x, L.xb.x * L.yb.x);
    w[(at + 1) & 15] = f2_pack(L.xa.y * L.ya.y, L.xb.y * L.yb.y);
    w[(at + 2) & 15] = f2_pack(L.xa.z * L.ya.z, L.xb.z * L.yb.z);
    w[(at + 3) & 15] = f2_pack(L.xa.w * L.ya.w, L.xb.w * L.yb.w);
}

#ifdef KX_TRACE
// Timing trace of ONE work item (development aid, compiled out of the shipped build; results are unaffected):
// [warp][band - KX_TRACE_J0][event] SM clocks, read back by tools/hv_trace.py through ssimu2_debug_hv_trace.
#ifndef KX_TRACE_ITEM
#define KX_TRACE_ITEM 485
#endif
#ifndef KX_TRACE_J0
#define KX_TRACE_J0 60
#endif
__device__ unsigned long long g_hv_trace[16][32][12];
#define KX_TR(ev) do { if (tr_on && lane == 0 && j >= KX_TRACE_J0 && j < KX_TRACE_J0 + 32) g_hv_trace[warp][j - KX_TRACE_J0][ev] = (unsigned long long)clock64(); } while (0)
#else
#define KX_TR(ev) do { } while (0)
#endif
constexpr int kXMaxNReg = 128;   // 16 warps x 128 registers = the whole register file
constexpr int kXHUnroll = 2;     // unroll of the 16-column body of the H scan: 2 halves the loop-carried register moves (-1.3 %); 4 overflows the instruction cache (+1.7 %)
// Warp roles (16 warps).  The scheduler sub-partition of a warp is (warp id % 4).  Per 12-row band an H warp needs
// ~1660 FP32-pipe cycles, a Va warp ~1150, a Vb warp ~750.  Every channel has TWO H warps that take alternate bands,
// so the waits / state fetch / first loads of band j+1 overlap the scan of band j:
//   SP0: H0a H0b Va0 P_out    SP1: H1a H1b Va1 (idle)    SP2: H2a H2b Va2 P_state    SP3: Vb0 Vb1 Vb2 P_tma
// plane slot of quantity q (s11, s22, s12, mu1, mu2) in the H-pass tile: planes [slot * 3 + channel]
__host__ __device__ constexpr int hv_slot(int q) { return q == 1 ? 3 : (q == 3 ? 1 : q); }
// Lite strips (score-only mode, X and B without SSIM'): 9 instead of 15 filtered quantities.  The H work is re-dealt over two
// warp sets -- HA = the five quantities of Y (as in the full map), HB = mu1 / mu2 of X and of B (24 lanes) -- and the V warps
// are placed so that no sub-partition carries more than ~1900 FP32-pipe cycles per band (full map: 2800 on three of them):
//   SP0: HAa HAb (idle) P_out    SP1: HBa HBb (idle) (idle)    SP2: Va_Y Vb_Y (idle) P_state    SP3: Vb_X Vb_B (idle) P_tma
__device__ __forceinline__ int hv_role(int warp, bool lite, int& ch, int& par)
{
    // 0 = H, 1 = Va, 2 = Vb, 3 = P_tma, 4 = P_out, 5 = idle, 6 = P_state, 7 = HB (lite strips)
    const int sp = warp & 3, row = warp >> 2;
    ch = sp; par = row;
    if (lite) {
        if (row == 2) return 5;
        if (row == 3) return sp == 0 ? 4 : (sp == 1 ? 5 : (sp == 2 ? 6 : 3));
        if (sp == 0) { ch = 1; return 0; }
        if (sp == 1) { ch = 0; return 7; }
        if (sp == 2) { ch = 1; par = 0; return row == 0 ? 1 : 2; }
        ch = row == 0 ? 0 : 2; par = 0;
        return 2;
    }
    if (sp < 3) {
        if (row < 2) return 0;
        if (row == 2) return 1;
        return sp == 0 ? 4 : (sp == 2 ? 6 : 5);
    }
    ch = row;
    return row < 3 ? 2 : 3;
}

__global__ void __maxnreg__(kXMaxNReg) k_hv(const __grid_constant__ Geo g, const __grid_constant__ TmaMapsX maps,
                                                     const HvArgs a)
{
    extern __shared__ __align__(1024) char xs[];
    __shared__ int s_item;
    const uint32_t sbase = smem_u32(xs);
    uint64_t* bars = reinterpret_cast<uint64_t*>(xs + kXOffBars);
    uint64_t* in_full = bars;        // [3] TMA
    uint64_t* in_free = bars + 3;    // [3] 3 Vb warps + every H warp (6, lite strips 4)
    uint64_t* hb_full = bars + 6;    // [3] 3 H warps
    uint64_t* hb_free = bars + 9;    // [3] 3 Va + 3 Vb warps
    uint64_t* hs_ready = bars + 12;  // [2] P_state: the left strip has published the state record of band j (slot j & 1)
    uint64_t* hs_free = bars + 15;   // [2] the 3 H warps of that band parity have read it
    // bars[29..31]: six plain words, bands completed by each H warp [parity][channel] (P_out may lag by any number of bands)
    const uint32_t hso_done = sbase + kXOffBars + 29 * 8;
    uint64_t* mu_full = bars + 17;   // [3 channels][2 slots] Vb -> Va
    uint64_t* mu_free = bars + 23;   // [3 channels][2 slots] Va -> Vb
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        const int it = (int)atomicAdd(a.ticket, 1u);
        s_item = it;
        int sc = 0;
        while (sc + 1 < g.nscales && it >= (g.sc[sc].item0 + g.sc[sc].n_strips) * a.nframes) sc++;
        const bool lt = a.lite(sc) == 5u;
        // arrivals per phase: the H warps that write a tile (one per channel / per set), the V warps that read it
        for (int i = 0; i < 3; i++) {
            mbar_init(&in_full[i], 1);
            // a tile slot is handed back by the three Vb warps AND by every H warp that waits on in_full: a parity wait is only
            // sound if no waiter can fall a whole ring cycle behind, and an H warp is a mere observer of the bands of the other
            // parity (see the H loop) -- without its arrival here the tile of band j + 3 could land before it has seen band j
            mbar_init(&in_free[i], lt ? 3 + 4 : 3 + 6);
            mbar_init(&hb_full[i], lt ? 2 : 3);
            mbar_init(&hb_free[i], lt ? 4 : 6);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(&hs_ready[i], 1);
            mbar_init(&hs_free[i], lt ? 2 : 3);
        }
        bars[29] = bars[30] = bars[31] = 0;
        for (int i = 0; i < 6; i++) {
            mbar_init(&mu_full[i], 1);
            mbar_init(&mu_free[i], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // tile slot 2 plays "the band above band 0": zeros (the vertical filter's zero padding); the ones row
    {
        float4* z = reinterpret_cast<float4*>(xs + kXOffHb + 2 * kXHbBytes);
        for (int i = tid; i < (int)(kXHbBytes / 16); i += kXThreads) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        float* ones = reinterpret_cast<float*>(xs + kXOffOnes);
        if (tid < (int)(kXOnesBytes / 4)) ones[tid] = 1.0f;
    }
    __syncthreads();

    // work item -> (scale, strip, frame), strip-major inside a scale so that the left neighbour has a smaller ticket
    int item = s_item, s = 0;
    {
        const int nf = a.nframes;
        while (s + 1 < g.nscales && item >= (g.sc[s].item0 + g.sc[s].n_strips) * nf) s++;
        item -= g.sc[s].item0 * nf;
    }
    const int k = item / a.nframes, frame = item - k * a.nframes;
#ifdef KX_TRACE
    const bool tr_on = s_item == KX_TRACE_ITEM;
#endif
    const ScaleDesc sd = g.sc[s];
    const int W = sd.w, H = sd.h, nb = sd.nb;
    const int x0 = k * kXC;
    const size_t rec_base = (size_t)frame * g.total_recs + sd.rec0;   // + strip * nb + band
    const bool last_strip = (k == sd.n_strips - 1);
    const bool lite = a.lite(s) == 5u;
    int c, hpar;
    const int role = hv_role(warp, lite, c, hpar);

    if (role == 3) {
        // ===== P_tma: tile loads =====
        if (lane != 0) return;
        const CUtensorMap* map = &maps.xyb_in[s];
        for (int j = 0; j < nb; j++) {
            const int si = j % kXNIn;
            if (j >= kXNIn) mbar_wait_wd(&in_free[si], (uint32_t)((j / kXNIn - 1) & 1));
            mbar_expect_tx(&in_full[si], kXInBytes);
            tma_load_4d(xs + kXOffIn + si * kXInBytes, map, &in_full[si], x0 - kXInLead, j * kXR, 0, frame);
            KX_TR(0);
        }
        return;
    }
    if (role == 6) {
        // ===== P_state: watches the flags of the strip to the left; the H warps then read the record from L2 =====
        if (k == 0 || lane != 0) return;
        for (int j = 0; j < nb; j++) {
            if (j >= 2) mbar_wait_wd(&hs_free[j & 1], (uint32_t)(((j >> 1) - 1) & 1));
            const uint32_t* fl = a.flags + rec_base + (size_t)(k - 1) * nb + j;
            for (uint32_t it = 0; ld_relaxed_u32(fl) != a.epoch; it++) {
                __nanosleep(40);
                if (it > 40000000u) __trap();
            }
            asm volatile("fence.acq_rel.gpu;" ::: "memory");   // acquire: the record was written before the flag
            mbar_arrive(&hs_ready[j & 1]);
        }
        return;
    }
    if (role == 4) {
        // ===== P_out: releases the right-edge state record of every band (written to global memory by the H warps) =====
        if (last_strip || lane != 0) return;
        for (int j = 0; j < nb; j++) {
            const uint32_t w0 = hso_done + (uint32_t)(j & 1) * 12u, need = (uint32_t)(j >> 1) + 1u;
            for (uint32_t it = 0; ld_acquire_cta_shared(w0) < need || ld_acquire_cta_shared(w0 + 4) < need ||
                                  ld_acquire_cta_shared(w0 + 8) < need; it++) {
                __nanosleep(300);
                if (it > 40000000u) __trap();
            }
            __threadfence();
            st_release_u32(a.flags + rec_base + (size_t)k * nb + j, a.epoch);
        }
        return;
    }
    if (role == 5) return;

    if (role == 0 || role == 7) {
        // ===== H: warp = channel, lane = (quantity, row pair); lanes 30, 31 shadow lane 29 =====
        // (HB of a lite strip: lanes 0-11 = mu1 / mu2 of X, 12-23 = mu1 / mu2 of B, lanes 24-31 shadow lane 23)
        // Lane -> (quantity q, row pair rp), chosen with the tile layouts so that every 128-bit shared-memory access of
        // the scan is conflict-free (4 wavefronts; tools/banksim.py): lanes are grouped by quantity in the order
        // s11, s12, mu1, s22, mu2; s12 and mu1 walk their row pairs in a rotated order; the planes of the output tile are
        // stored in the order s11, mu1, s12, s22, mu2 (kXSlot); the rows of ones sit at a per-channel offset.
        int ch = c, q, rp;
        if (role == 0) {
            const int l = lane < 30 ? lane : 29;
            const int qidx = l / 6, jj = l - 6 * qidx;
            q = (0x41320 >> (4 * qidx)) & 7;
            rp = q == 2 ? ((0x541032 >> (4 * jj)) & 7) : (q == 3 ? ((0x325410 >> (4 * jj)) & 7) : jj);
        } else {
            const int l = lane < 24 ? lane : 23;
            ch = l < 12 ? 0 : 2;
            const int l12 = l < 12 ? l : l - 12;
            q = 3 + l12 / 6;
            rp = l12 % 6;
        }
        const int px = (q == 1 || q == 4) ? 3 + ch : ch;
        const int py = (q == 0) ? ch : ((q == 1 || q == 2) ? 3 + ch : -1);
        const uint32_t offxA = (uint32_t)((px * kXInPlane + rp * kXInW) * 4), offxB = offxA + 6 * kXInW * 4;
        const uint32_t offyA = py < 0 ? 0u : (uint32_t)((py * kXInPlane + rp * kXInW) * 4), offyB = offyA + 6 * kXInW * 4;
        const uint32_t onesA = sbase + kXOffOnes + (ch == 1 ? 32u : 16u), onesB = sbase + kXOffOnes + (ch == 1 ? 64u : 0u);
        const uint32_t offo = (uint32_t)(((hv_slot(q) * 3 + ch) * kXHbPlane + rp * kXHbPitch) * 4);
        const int hidx = role == 0 ? ch * 32 + lane : lane;   // slot of this lane in the hand-off record (HB: 0-31, the X slots)
        for (int j = 0; j < nb; j++) {
            // every H warp walks ALL the phases of the ring barriers in order (a parity wait must never skip a phase),
            // but only scans the bands of its own parity
            const int si = j % 3, sin = j % kXNIn;
            const bool mine = (j & 1) == hpar;
            mbar_wait_wd(&in_full[sin], (uint32_t)((j / kXNIn) & 1));
            if (mine) KX_TR(0);
            HState2 st;
            if (mine) {
                if (k > 0) {
                    mbar_wait_wd(&hs_ready[j & 1], (uint32_t)((j >> 1) & 1));
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&hs_free[j & 1]);   // P_state may go on to band j + 2
                    const f2* rec = a.hstate + (rec_base + (size_t)(k - 1) * nb + j) * kXHsF2 + hidx;
                    st.p1 = ldcg64(rec); st.p3 = ldcg64(rec + 96); st.p5 = ldcg64(rec + 2 * 96);
                    st.pp1 = ldcg64(rec + 3 * 96); st.pp3 = ldcg64(rec + 4 * 96); st.pp5 = ldcg64(rec + 5 * 96);
                } else {
                    const f2 z = f2_splat(0.0f);
                    st = HState2{z, z, z, z, z, z};
                }
            }
            if (!mine) {
                // observed, not consumed: hand the slot back at once (found by tools/soak.py: in a fast lite strip the tile of
                // band j + 3 could land while this warp was still suspended in the wait for band j; its parity wait then saw
                // "not complete" for ever and the strip chain dead-locked)
                __syncwarp();
                if (lane == 0) mbar_arrive(&in_free[sin]);
                if (j >= 2) mbar_wait_wd(&hb_free[si], (uint32_t)(((j - 2) / 3) & 1));
                continue;
            }
            const uint32_t inb = sbase + kXOffIn + sin * kXInBytes;
            const uint32_t axA = inb + offxA, axB = inb + offxB;
            const uint32_t ayA = py < 0 ? onesA : inb + offyA, ayB = py < 0 ? onesB : inb + offyB;
            uint32_t ao = sbase + kXOffHb + si * kXHbBytes + offo;
            // w[c & 15] = product of tile column c (both rows).  Tile column c is x[n + 4] of output column n = c - 12;
            // the left tap x[n - 6] is tile column c - 10.
            f2 w[16];
            {
                const HGroupLoad L0 = h_load(axA, axB, ayA, ayB, 0), L1 = h_load(axA, axB, ayA, ayB, 16),
                                 L2 = h_load(axA, axB, ayA, ayB, 32);
                h_products(L0, w, 0);
                h_products(L1, w, 4);
                h_products(L2, w, 8);
                if (k == 0) {
                    // the recursion starts at n = -4 (cpu.rs:976): four warm-up steps on x[0..3], no output
#pragma unroll
                    for (int e = 0; e < 4; e++) (void)hstep2(st, f2_splat(0.0f), w[8 + e]);
                }
            }
            HGroupLoad nxt = h_load(axA, axB, ayA, ayB, 48);
            KX_TR(1);
            if (j >= 2) mbar_wait_wd(&hb_free[si], (uint32_t)(((j - 2) / 3) & 1));
            KX_TR(2);
            uint32_t axI = axA, ayI = ayA, ayIB = ayB;
#pragma unroll kXHUnroll
            for (int it = 0; it < 4; it++, axI += 64, ayI += 64, ayIB += 64) {
#pragma unroll
                for (int gg = 0; gg < 4; gg++) {
                    const HGroupLoad cur = nxt;
                    // prefetch the next group's operands before this group's 4 steps.  The very last prefetch (group 19) lies
                    // one group past the tile row: it stays inside shared memory, is never used, and leaving it unclamped keeps
                    // every offset of the loop body a compile-time immediate
                    nxt = h_load(axI, axI + 6 * kXInW * 4, ayI, ayIB, (uint32_t)(gg + 4) * 16u);
                    h_products(cur, w, 12 + 4 * gg);
                    float oa[4], ob[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const f2 o = hstep2(st, w[(4 * gg + e + 2) & 15], w[(4 * gg + e + 12) & 15]);
                        f2_unpack(o, oa[e], ob[e]);
                    }
                    sts128(ao + (uint32_t)gg * 16u, make_float4(oa[0], oa[1], oa[2], oa[3]));
                    sts128(ao + (uint32_t)gg * 16u + 6 * kXHbPitch * 4, make_float4(ob[0], ob[1], ob[2], ob[3]));
                }
                ao += 64;
            }
            if (last_strip && x0 + kXC > W) {
                // columns past the right edge hold the filter's ring-out: the V pass must see zeros there
                float* t0 = reinterpret_cast<float*>(xs + kXOffHb + si * kXHbBytes) + (hv_slot(q) * 3 + ch) * kXHbPlane + rp * kXHbPitch;
                for (int cc = W - x0; cc < kXC; cc++) { t0[cc] = 0.0f; t0[6 * kXHbPitch + cc] = 0.0f; }
            }
            if (!last_strip) {
                // the state at the right edge of the band goes straight to the record of (strip, band): 256 B per store
                unsigned long long* rec = reinterpret_cast<unsigned long long*>(a.hstate + (rec_base + (size_t)k * nb + j) * kXHsF2) + hidx;
                __stcg(rec, st.p1.v); __stcg(rec + 96, st.p3.v); __stcg(rec + 2 * 96, st.p5.v);
                __stcg(rec + 3 * 96, st.pp1.v); __stcg(rec + 4 * 96, st.pp3.v); __stcg(rec + 5 * 96, st.pp5.v);
            }
            __syncwarp();
            KX_TR(3);
            if (lane == 0) {
                mbar_arrive(&in_free[sin]);    // this warp's last read of the XYB tile was in the scan
                mbar_arrive(&hb_full[si]);
                if (!last_strip) {
                    // progress words of the publisher: one per channel; HB stands for X and B
                    const uint32_t w0 = hso_done + (uint32_t)((j & 1) * 3) * 4u, done = (uint32_t)(j >> 1) + 1u;
                    if (role == 0) {
                        st_release_cta_shared(w0 + (uint32_t)c * 4u, done);
                    } else {
                        st_release_cta_shared(w0, done);
                        st_release_cta_shared(w0 + 8u, done);
                    }
                }
            }
        }
        return;
    }

    // ===== V: lane = column pair.  Vb runs the mu1 / mu2 filters and the edge maps and hands the blurred mu rows to
    // Va (4 rows at a time through a 2-slot ring); Va runs the s11 / s22 / s12 filters and the SSIM map.
    // A filter step for input row t = 12 j + i reads x[t] (row i of this band's tile) and x[t - 10] (10 rows up: row i + 2
    // of the previous band's tile for i < 10, else row i - 10 of this one) and yields output row t - 4. =====
    const f2 zero2 = f2_splat(0.0f);
    const uint32_t lane8 = (uint32_t)lane * 8u;
    constexpr uint32_t kXMu2Off = (uint32_t)((hv_slot(4) - hv_slot(3)) * 3 * kXHbPlane * 4);   // mu2 plane - mu1 plane
    const uint32_t mub = sbase + kXOffMu + (uint32_t)c * 2u * kXMuSlotBytes + lane8;
    const int nsub = nb * (kXR / kXSub);
    if (role == 2) {
        VState2 stq[2];
#pragma unroll
        for (int qi = 0; qi < 2; qi++) stq[qi] = VState2{zero2, zero2, zero2, zero2, zero2, zero2};
        f2 fifo_r[4] = {zero2, zero2, zero2, zero2}, fifo_d[4] = {zero2, zero2, zero2, zero2};  // XYB rows t-4 .. t-1
        double acc[4] = {0, 0, 0, 0};
        int n = 0;   // sub-band counter
        for (int j = 0; j < nb; j++) {
            const int si = j % 3, sp = (j + 2) % 3;
            mbar_wait_wd(&hb_full[si], (uint32_t)((j / 3) & 1));
            mbar_wait_wd(&in_full[si], (uint32_t)((j / 3) & 1));
            KX_TR(0);
            const uint32_t cur = sbase + kXOffHb + si * kXHbBytes + lane8 + (uint32_t)((hv_slot(3) * 3 + c) * kXHbPlane * 4);
            const uint32_t prv = sbase + kXOffHb + sp * kXHbBytes + lane8 + (uint32_t)((hv_slot(3) * 3 + c) * kXHbPlane * 4);
            const uint32_t inb = sbase + kXOffIn + si * kXInBytes + kXInLead * 4 + lane8 + (uint32_t)(c * kXInPlane * 4);
            f2 part[4] = {zero2, zero2, zero2, zero2};   // f32 sums of one band (12 rows x 2 columns), then f64
#pragma unroll 1
            for (int i4 = 0; i4 < kXR; i4 += kXSub, n++) {
                const int p = n & 1;
                const bool hand = !(lite && c != 1);        // lite strips: nobody consumes the blurred mu rows of X and B
                KX_TR(2 + i4 / 4 * 3);
                if (hand && n >= 2) nbar_sync(7 + 2 * c + p, 64);   // Va has read this slot's previous rows
                KX_TR(3 + i4 / 4 * 3);
                const uint32_t mus = mub + (uint32_t)p * kXMuSlotBytes;
                // delayed-tap rows of the sub-band: previous tile rows i4 + 2 .. i4 + 5, except that for i4 = 8 the last two
                // (band rows 10, 11) are rows 0, 1 of this band's tile.  Formed once per sub-band: the V warps are the
                // busiest of the kernel and every integer instruction per row shows (-3 %).
                const uint32_t cur_i4 = cur + (uint32_t)(i4 * kXHbPitch * 4);
                const uint32_t d_lo = prv + (uint32_t)((i4 + 2) * kXHbPitch * 4);
                const uint32_t d_hi = i4 == 8 ? cur - (uint32_t)(2 * kXHbPitch * 4) : d_lo;
                // rows of one sub-band; CHECKED only for the sub-bands that straddle output rows -4..-1 or H..: the
                // common path has no branch per row, so the four rows' map chains interleave
                auto rows = [&](auto checked) {
#pragma unroll
                    for (int r = 0; r < kXSub; r++) {
                        const int i = i4 + r, t = j * kXR + i;
                        const uint32_t a_t = cur_i4 + (uint32_t)(r * kXHbPitch * 4);
                        const uint32_t a_d = (r < 2 ? d_lo : d_hi) + (uint32_t)(r * kXHbPitch * 4);
                        const f2 m1 = vstep2(stq[0], lds64(a_d), lds64(a_t));
                        const f2 m2 = vstep2(stq[1], lds64(a_d + kXMu2Off), lds64(a_t + kXMu2Off));
                        if (hand) {
                            sts64(mus + (uint32_t)(r * 2 * kXC * 4), m1);
                            sts64(mus + (uint32_t)((r * 2 + 1) * kXC * 4), m2);
                        }
                        const f2 fr = fifo_r[r], fd = fifo_d[r];   // XYB of output row t - 4
                        fifo_r[r] = lds64(inb + (uint32_t)(i * kXInW * 4));
                        fifo_d[r] = lds64(inb + (uint32_t)((3 * kXInPlane + i * kXInW) * 4));
                        if (!decltype(checked)::value || (t >= 4 && t < H + 4)) edge_maps2(m1, m2, fr, fd, part);
                    }
                };
                const int t0 = j * kXR + i4;
                if (t0 >= 4 && t0 + kXSub <= H + 4) rows(std::false_type{}); else rows(std::true_type{});
                KX_TR(4 + i4 / 4 * 3);
                if (hand) nbar_arrive(1 + 2 * c + p, 64);   // rows ready for Va
            }
#pragma unroll
            for (int kk = 0; kk < 4; kk++) acc[kk] += (double)f2_hsum(part[kk]);
            __syncwarp();
            KX_TR(1);
            if (lane == 0) {
                mbar_arrive(&in_free[si]);
                mbar_arrive(&hb_free[sp]);   // the previous band's tile goes back to the H warps
            }
        }
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
            double vsum = acc[kk];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) vsum += __shfl_down_sync(0xffffffffu, vsum, off);
            if (lane == 0) a.partials[((size_t)frame * g.total_strips + sd.strip0 + k) * 18 + c * 6 + 2 + kk] = vsum;
        }
        return;
    }

    // ----- Va -----
    VState2 stq[3];
#pragma unroll
    for (int qi = 0; qi < 3; qi++) stq[qi] = VState2{zero2, zero2, zero2, zero2, zero2, zero2};
    double acc[2] = {0, 0};
    int n = 0;
    for (int j = 0; j < nb; j++) {
        const int si = j % 3, sp = (j + 2) % 3;
        mbar_wait_wd(&hb_full[si], (uint32_t)((j / 3) & 1));
        KX_TR(0);
        const uint32_t cur = sbase + kXOffHb + si * kXHbBytes + lane8 + (uint32_t)(c * kXHbPlane * 4);
        const uint32_t prv = sbase + kXOffHb + sp * kXHbBytes + lane8 + (uint32_t)(c * kXHbPlane * 4);
        f2 part[2] = {zero2, zero2};
#pragma unroll 1
        for (int i4 = 0; i4 < kXR; i4 += kXSub, n++) {
            const int p = n & 1;
            f2 o[kXSub][3];
#pragma unroll
            for (int r = 0; r < kXSub; r++)
                for (int qi = 0; qi < 3; qi++) o[r][qi] = zero2;
            const uint32_t cur_i4 = cur + (uint32_t)(i4 * kXHbPitch * 4);
            const uint32_t d_lo = prv + (uint32_t)((i4 + 2) * kXHbPitch * 4);
            const uint32_t d_hi = i4 == 8 ? cur - (uint32_t)(2 * kXHbPitch * 4) : d_lo;
#pragma unroll
            for (int r = 0; r < kXSub; r++) {
                const uint32_t a_t = cur_i4 + (uint32_t)(r * kXHbPitch * 4);
                const uint32_t a_d = (r < 2 ? d_lo : d_hi) + (uint32_t)(r * kXHbPitch * 4);
#pragma unroll
                for (int qi = 0; qi < 3; qi++) {
                    const uint32_t pl = (uint32_t)(hv_slot(qi) * 3 * kXHbPlane) * 4u;
                    o[r][qi] = vstep2(stq[qi], lds64(a_d + pl), lds64(a_t + pl));
                }
            }
            if (i4 == kXR - kXSub) {   // last read of the previous band's tile
                __syncwarp();
                if (lane == 0) mbar_arrive(&hb_free[sp]);
            }
            KX_TR(2 + i4 / 4 * 3);
            nbar_sync(1 + 2 * c + p, 64);
            KX_TR(3 + i4 / 4 * 3);
            const uint32_t mus = mub + (uint32_t)p * kXMuSlotBytes;
            auto maps = [&](auto checked) {
#pragma unroll
                for (int r = 0; r < kXSub; r++) {
                    const int t = j * kXR + i4 + r;
                    const f2 m1 = lds64(mus + (uint32_t)(r * 2 * kXC * 4)), m2 = lds64(mus + (uint32_t)((r * 2 + 1) * kXC * 4));
                    if (!decltype(checked)::value || (t >= 4 && t < H + 4)) ssim_map2(o[r][0], o[r][1], o[r][2], m1, m2, part);
                }
            };
            const int t0 = j * kXR + i4;
            if (t0 >= 4 && t0 + kXSub <= H + 4) maps(std::false_type{}); else maps(std::true_type{});
            KX_TR(4 + i4 / 4 * 3);
            if (n + 2 < nsub) nbar_arrive(7 + 2 * c + p, 64);   // Vb waits for it before it reuses the slot (never after the last use)
        }
        acc[0] += (double)f2_hsum(part[0]);
        acc[1] += (double)f2_hsum(part[1]);
        KX_TR(1);
    }
#pragma unroll
    for (int kk = 0; kk < 2; kk++) {
        double vsum = acc[kk];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) vsum += __shfl_down_sync(0xffffffffu, vsum, off);
        if (lane == 0) a.partials[((size_t)frame * g.total_strips + sd.strip0 + k) * 18 + c * 6 + kk] = vsum;
    }
}

// ------------------------------------------------------------------------------------------
// k_finalize: partial sums -> 108 norms -> score.  Replaces the host-side
// Ssimulacra2::post_process_scores (ssimulacra2-cuda/src/lib.rs:449-623); follows
// Msssim::score cpu.rs:728-871 (weight order [channel][scale][L1,L4][ssim,artifact,detail]).
// One CTA per frame.
// ------------------------------------------------------------------------------------------
#define SSIMU2_WEIGHTS108     0.0, 0.0007376606707406586, 0.0, 0.0, 0.0007793481682867309, 0.0, \
    0.0, 0.0004371155730107379, 0.0, 1.1041726426657346, 0.00066284834129271, 0.00015231632783718752, \
    0.0, 0.0016406437456599754, 0.0, 1.8422455520539298, 11.441172603757666, 0.0, \
    0.0007989109436015163, 0.000176816438078653, 0.0, 1.8787594979546387, 10.94906990605142, 0.0, \
    0.0007289346991508072, 0.9677937080626833, 0.0, 0.00014003424285435884, 0.9981766977854967, 0.00031949755934435053, \
    0.0004550992113792063, 0.0, 0.0, 0.0013648766163243398, 0.0, 0.0, \
    0.0, 0.0, 0.0, 7.466890328078848, 0.0, 17.445833984131262, \
    0.0006235601634041466, 0.0, 0.0, 6.683678146179332, 0.00037724407979611296, 1.027889937768264, \
    225.20515300849274, 0.0, 0.0, 19.213238186143016, 0.0011401524586618361, 0.001237755635509985, \
    176.39317598450694, 0.0, 0.0, 24.43300999870476, 0.28520802612117757, 0.0004485436923833408, \
    0.0, 0.0, 0.0, 34.77906344483772, 44.835625328877896, 0.0, \
    0.0, 0.0, 0.0, 0.0, 0.0, 0.0, \
    0.0, 0.0008680556573291698, 0.0, 0.0, 0.0, 0.0, \
    0.0, 0.0005313191874358747, 0.0, 0.00016533814161379112, 0.0, 0.0, \
    0.0, 0.0, 0.0, 0.0004179171803251336, 0.0017290828234722833, 0.0, \
    0.0020827005846636437, 0.0, 0.0, 8.826982764996862, 23.19243343998926, 0.0, \
    95.1080498811086, 0.9863978034400682, 0.9834382792465353, 0.0012286405048278493, 171.2667255897307, 0.9807858872435379, \
    0.0, 0.0, 0.0, 0.0005130064588990679, 0.0, 0.00010854057858411537,
__device__ const double kWeight[108] = {SSIMU2_WEIGHTS108};


__global__ void __launch_bounds__(128) k_finalize(const __grid_constant__ Geo g, const double* __restrict__ partials,
                                                  double* __restrict__ norms_out, double* __restrict__ scores_ring,
                                                  unsigned long long first_ticket, unsigned long long ring_cap,
                                                  double* __restrict__ scores_out, uint32_t* __restrict__ hv_ticket)
{
    if (hv_ticket != nullptr && blockIdx.x == 0 && threadIdx.x == 0) { hv_ticket[0] = 0u; hv_ticket[1] = 0u; }  // work counter of k_hv, for the next batch
    __shared__ double norms[108];
    const int frame = blockIdx.x, tid = threadIdx.x;
    if (tid < 108) norms[tid] = 0.0;
    __syncthreads();
    if (tid < 108) {
        // tid = s*18 + c*6 + k, k: 0 ssim L1, 1 ssim L4, 2 art L1, 3 art L4, 4 det L1, 5 det L4
        int s = tid / 18, c = (tid / 6) % 3, k = tid % 6;
        if (s < g.nscales) {
            const ScaleDesc& sd = g.sc[s];
            const double* p = partials + ((size_t)frame * g.total_strips + sd.strip0) * 18 + c * 6 + k;
            double sum = 0.0;
            for (int i = 0; i < sd.n_strips; i++) sum += p[(size_t)i * 18];
            double one_per_pixels = 1.0 / (double)((size_t)sd.w * sd.h);
            double v = one_per_pixels * sum;
            int m = k >> 1, n = k & 1;
            if (n) v = sqrt(sqrt(v));
            norms[c * 36 + s * 6 + n * 3 + m] = v;
        }
    }
    __syncthreads();
    if (tid < 108) norms_out[(size_t)frame * 108 + tid] = norms[tid];
    if (tid == 0) {
        // cpu.rs:840-868; with fewer than 6 scales the weight cursor is dense, as in the reference.
        double ssim = 0.0;
        int i = 0;
        for (int c = 0; c < 3; c++)
            for (int s = 0; s < g.nscales; s++)
                for (int n = 0; n < 2; n++)
                    for (int m = 0; m < 3; m++) {
                        ssim = fma(kWeight[i], fabs(norms[c * 36 + s * 6 + n * 3 + m]), ssim);
                        i++;
                    }
        ssim *= 0.9562382616834844;
        ssim = fma(6.248496625763138e-5 * ssim * ssim, ssim,
                   fma(2.326765642916932, ssim, -0.020884521182843837 * ssim * ssim));
        if (ssim > 0.0)
            ssim = fma(pow(ssim, 0.6276336467831387), -10.0, 100.0);
        else
            ssim = 100.0;
        scores_out[frame] = ssim;
        scores_ring[(first_ticket + frame) % ring_cap] = ssim;
    }
}

// Builds Geo::eotf_lut with the arithmetic path of load_px (same expressions, same device routines).
__global__ void k_build_eotf_lut(const YuvCoef k, int n, int shift, float* __restrict__ out)
{
    __shared__ exact_math::PowfTables T;
    {
        const uint64_t* src = reinterpret_cast<const uint64_t*>(&kPowfTablesInit);
        uint64_t* dst = reinterpret_cast<uint64_t*>(&T);
        for (int i = threadIdx.x; i < (int)(sizeof(exact_math::PowfTables) / 8); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * n) return;
    const int c = idx / n, y = idx - c * n;
    const int Y = y << shift, C = c << shift;
    const float cc = (float)(C - k.neutral);
    const float luma = (float)(max(Y, k.luma_min) - k.luma_min) * k.y;
    const float r_ = k.r * cc, b_ = k.b * cc;
    out[idx] = clamp01(bt709_eotf(luma + r_, T));
    out[(size_t)n * n + idx] = clamp01(bt709_eotf(luma + b_, T));
}

// Builds the 65536-entry memo of the sRGB16 transfer (Geo::eotf_lut of an SSIMU2_FMT_SRGB16 handle) with the arithmetic path of
// load_px: same expression, same device routine, so a looked-up value has the bits of the computed one.
__global__ void k_build_srgb16_lut(float* __restrict__ out)
{
    __shared__ exact_math::PowfTables T;
    {
        const uint64_t* src = reinterpret_cast<const uint64_t*>(&kPowfTablesInit);
        uint64_t* dst = reinterpret_cast<uint64_t*>(&T);
        for (int i = threadIdx.x; i < (int)(sizeof(exact_math::PowfTables) / 8); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < 65536) out[c] = srgb_inverse_oetf<false>((float)c / 65535.0f, T);
}

// Test hook: the device build of exact_math.cuh over an array.
//   op 0: out[i] = cbrtf(in[i])            op 1: out[i] = powf(in[i], y)
//   op 2: out[i] = fdiv_normal(in[i], y)   op 3: in = n pairs of doubles (num, den), out = n doubles ddiv_normal
//   op 4: out[i] = div_rn_normal(in[2i], in[2i+1]) (the V-pass quotient)
//   op 5 / 6: the unchecked hot-path forms of op 0 / 1 (positive normal arguments only)
__global__ void k_debug_math(int op, const float* __restrict__ in, float y, float* __restrict__ out, size_t n)
{
    __shared__ exact_math::CbrtScale S;
    S.tab[threadIdx.x] = exact_math::cbrt_scale_entry(threadIdx.x);
    __syncthreads();
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (op == 0) out[i] = exact_math::cbrtf_glibc<true>(in[i], kEM, &S);
    else if (op == 1) out[i] = exact_math::powf_glibc<true>(in[i], y, kEM, kPowfTablesInit);
    else if (op == 5) out[i] = exact_math::cbrtf_glibc<false>(in[i], kEM, &S);
    else if (op == 6) out[i] = exact_math::powf_glibc<false>(in[i], y, kEM, kPowfTablesInit);
    else if (op == 2) out[i] = exact_math::fdiv_normal(in[i], y);
    else if (op == 3) {
        const double* din = reinterpret_cast<const double*>(in);
        reinterpret_cast<double*>(out)[i] = exact_math::ddiv_normal(din[2 * i], din[2 * i + 1]);
    } else out[i] = div_rn_normal(in[2 * i], in[2 * i + 1]);
}

}  // namespace ssimu2
