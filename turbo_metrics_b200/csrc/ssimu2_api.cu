// ssimu2_api.cu -- host side of libssimu2_b200.so: the C ABI declared in include/ssimu2_b200.h.
//
// Mirrors the reference's metric op (crates/ssimulacra2-cuda/src/lib.rs:27-291) and the per-pair
// driver that calls it (crates/turbo-metrics/src/lib.rs:268-360), re-designed for throughput:
//   * a handle owns `ring` batch slots; each slot has its own stream, workspace and result buffers;
//   * pairs are collected into batches of `batch` and every batch is 4 kernel launches that cover
//     all frames and all 6 scales (the reference records a 305-node graph per pair and syncs the
//     host after every pair, lib.rs:342-352);
//   * only the f64 scores (and, for parity tests, the 108 norms) travel back to the host.
// There is no CPU fallback: without a usable CUDA device every call fails.
#include "../../include/ssimu2_b200.h"
#include "ssimu2_kernels.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

using namespace ssimu2;

namespace {

constexpr uint64_t kResultCap = 16384;  // tickets kept in the host/device result rings
constexpr uint32_t kDefaultBatch = 8;
constexpr uint32_t kDefaultRing = 3;

#define CU_TRY(expr)                                 \
    do {                                             \
        cudaError_t _e = (expr);                     \
        if (_e != cudaSuccess) return (int)_e;       \
    } while (0)

struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_in = nullptr;      // dependency on the submitter's stream
    cudaEvent_t ev_done = nullptr;    // batch complete (results on host)
    cudaEvent_t ev_k[5] = {};         // per-kernel timing marks
    float* xyb = nullptr;             // [batch] XYB planes of all scales
    float* hb = nullptr;              // [batch] H-pass planes
    double* partials = nullptr;       // [batch][total_strips][18]
    double* norms_d = nullptr;        // [batch][108]
    double* scores_d = nullptr;       // [batch]
    double* norms_h = nullptr;        // pinned
    double* scores_h = nullptr;       // pinned
    uint8_t* staging = nullptr;       // device staging for host frames: [batch][2][staging_frame_bytes]
    BatchIn in{};
    TmaMaps maps{};                   // tensor maps of this slot's H-pass / XYB buffers, per scale (V pass)
    TmaMapsH maps_h{};                // (H pass)
    TmaMapsX maps_x{};                // (fused H+V kernel)
    f2* hstate = nullptr;             // k_hv hand-off records: [batch][total_recs][6][96]
    uint32_t* hvflags = nullptr;      // [batch][total_recs] + 1 ticket counter at the end
    uint32_t epoch = 0;               // launch counter of k_hv on this slot (flag value)
    uint32_t count = 0;               // pairs recorded
    uint64_t first_ticket = 0;
    bool inflight = false;            // fully launched, results not harvested yet
    bool awaiting = false;            // front-end launched, H pass waiting for the next batch (fused mode)
    bool staged = false;              // holds host frames copied on the slot stream
    bool was_timed = false;
    cudaEvent_t ev_mid = nullptr;     // main stream -> slot stream hand-off
    cudaEvent_t ev_f = nullptr;       // front-end of this slot done (two-stream pipeline)
    bool timed = false;
    void* last_stream = nullptr;
    bool have_dep = false;
};

}  // namespace

struct ssimu2_handle {
    ssimu2_config cfg{};
    Geo geo{};
    uint32_t batch = 0, ring = 0;
    std::vector<Slot> slots;
    uint32_t cur = 0;
    int awaiting = -1;                // slot index in the `awaiting` state, or -1
    bool fuse = false;                // cross-batch fusion of front-end and H pass (pipeline "fh", ring >= 2)
    bool frontend2 = true;            // warp-per-region front-end (SSIMU2_FRONTEND=1 selects the shared-memory tile version)
    int pipeline = 0;                 // 0 = "hv": front-end, fused H+V kernel, finalize (default)
                                      // 1 = "fh": k_fused_fh + k_vpass (ring >= 2) / 2 = "split": four kernels
    cudaStream_t main_stream = nullptr;
    cudaStream_t f_stream = nullptr, hv_stream = nullptr;  // pipeline "hv", ring >= 2: front-end stream / H+V stream
    bool two_stream = false;
    int num_sms = 0, f2p_ctas = 1;
    uint64_t next_ticket = 0;
    double* scores_ring_d = nullptr;  // [kResultCap] device score stream
    float* eotf_lut = nullptr;        // exact R / B transfer memo for YUV sources (see Geo)
    std::vector<double> res_scores;   // host result ring
    std::vector<double> res_norms;    // [kResultCap][108]
    std::vector<int32_t> res_slot;    // slot that served the ticket (for debug_read)
    size_t device_bytes = 0;
    size_t staging_frame_bytes = 0;
    uint64_t launches = 0;
    float last_ms[4] = {0, 0, 0, 0};
    double total_ms[4] = {0, 0, 0, 0};  // per-kernel device time summed over harvested batches
    uint64_t timed_batches = 0, timed_pairs = 0;
    uint64_t alg_bytes = 0;
};

namespace {

// ---- colour coefficients (host, f32 arithmetic as in cuda-colorspace-kernel/src/lib.rs:183-218) ----
struct V3 { float x, y, z; };
static V3 xyz_of(float x, float y) { return V3{x / y, 1.0f, (1.0f - x - y) / y}; }
static float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static V3 cross(V3 a, V3 b) { return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

static void luma_constants(int matrix, float& kr, float& kb)
{
    // primaries: constants.rs:3-18
    float p[3][6] = {{0.640f, 0.330f, 0.300f, 0.600f, 0.150f, 0.060f},
                     {0.630f, 0.340f, 0.310f, 0.595f, 0.155f, 0.070f},
                     {0.640f, 0.330f, 0.290f, 0.600f, 0.150f, 0.060f}};
    const float* m = p[matrix];
    V3 r = xyz_of(m[0], m[1]), g = xyz_of(m[2], m[3]), b = xyz_of(m[4], m[5]), w = xyz_of(0.3127f, 0.3290f);
    V3 xr{r.x, g.x, b.x}, yr{r.y, g.y, b.y}, zr{r.z, g.z, b.z};
    float mul = 1.0f / dot(xr, cross(yr, zr));
    kr = dot(w, cross(g, b)) * mul;
    kb = dot(w, cross(r, g)) * mul;
}

static YuvCoef make_coef(int fmt, int matrix, int full_range)
{
    YuvCoef c{};
    if (fmt != kNV12 && fmt != kP016) return c;
    int bits = fmt == kNV12 ? 8 : 16;
    float kr, kb;
    luma_constants(matrix, kr, kb);
    uint32_t lmin, lrange, crange;
    if (full_range) {
        lmin = 0; lrange = (1u << bits) - 1; crange = (1u << bits) - 1;
    } else {
        lmin = 16u << (bits - 8);
        lrange = (235u << (bits - 8)) - lmin;
        crange = (240u << (bits - 8)) - lmin;
    }
    float kg = 1.0f - kr - kb;
    c.y = 1.0f / (float)lrange;
    c.r = 2.0f * (1.0f - kr) * 1.0f / (float)crange;
    c.b = 2.0f * (1.0f - kb) * 1.0f / (float)crange;
    c.g1 = -2.0f * (1.0f - kb) * kb / kg * 1.0f / (float)crange;
    c.g2 = -2.0f * (1.0f - kr) * kr / kg * 1.0f / (float)crange;
    c.luma_min = (int)lmin;
    c.neutral = 1 << (bits - 1);
    return c;
}

static int in_bytes_per_px(int fmt)
{
    switch (fmt) {
    case kNV12: return 3;       // 1.5 B x 2 images
    case kP016: return 6;
    case kSRGB8: return 6;
    case kSRGB16: return 12;
    default: return 24;
    }
}

static void build_geo(ssimu2_handle* h)
{
    Geo& g = h->geo;
    int w = (int)h->cfg.width, hh = (int)h->cfg.height;
    long long xyb_off = 0, hb_off = 0;
    int strips = 0, ns = 0, recs = 0;
    unsigned long long sum_px = 0, sum_px_ge1 = 0;
    for (int s = 0; s < kMaxScales; s++) {
        if (w < 8 || hh < 8) break;  // cpu.rs:359: tested on the size BEFORE this scale's downscale
        if (s > 0) { w = (w + 1) / 2; hh = (hh + 1) / 2; }
        ScaleDesc& d = g.sc[s];
        d.w = w; d.h = hh;
        d.pitch = (w + 31) / 32 * 32;
        d.n_bands = (hh + kHRows - 1) / kHRows;
        d.n_strips = (w + kVCols - 1) / kVCols;
        d.strip0 = strips;
        strips += d.n_strips;
        d.xyb_off = xyb_off;
        xyb_off += 6LL * hh * d.pitch;
        d.hb_off = hb_off;
        hb_off += 15LL * hh * d.pitch;
        d.nb = (hh + 4 + kXR - 1) / kXR;
        d.rec0 = recs;
        d.item0 = strips - d.n_strips;
        recs += d.n_strips * d.nb;
        sum_px += (unsigned long long)w * hh;
        if (s >= 1) sum_px_ge1 += (unsigned long long)w * hh;
        ns++;
    }
    g.nscales = ns;
    g.total_strips = strips;
    g.total_recs = recs;
    g.xyb_stride = xyb_off;
    g.hb_stride = hb_off;
    g.items_h = 0; g.items_v = 0;
    for (int s = 0; s < ns; s++) { g.items_h += g.sc[s].n_bands; g.items_v += g.sc[s].n_strips; }
    g.coef = make_coef(h->cfg.format, h->cfg.matrix, h->cfg.full_range);
    // B_alg = 120*sum(P_s) + 2*in_0*P_0 + 72*sum_{s>=1}(P_s)
    h->alg_bytes = 120ULL * sum_px + 2ULL * in_bytes_per_px(h->cfg.format) * h->cfg.width * h->cfg.height + 72ULL * sum_px_ge1;
}

// ---- TMA tensor maps (driver entry point fetched through the runtime; libcuda is not linked) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_plane_map(EncodeTiledFn enc, CUtensorMap* out, float* base, const ScaleDesc& d, int planes, long long slot_stride,
                          uint32_t batch, uint32_t box_cols, uint32_t box_rows, bool swizzle128)
{
    // 4-D view {x, y, plane, frame} of a [frame][plane][h][pitch] f32 buffer; out-of-range elements read as 0
    cuuint64_t dims[4] = {(cuuint64_t)d.w, (cuuint64_t)d.h, (cuuint64_t)planes, (cuuint64_t)batch};
    cuuint64_t strides[3] = {(cuuint64_t)d.pitch * 4, (cuuint64_t)d.h * d.pitch * 4, (cuuint64_t)slot_stride * 4};
    cuuint32_t box[4] = {box_cols, box_rows, (cuuint32_t)planes, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : SSIMU2_E_INTERNAL;
}

static int build_tma_maps(ssimu2_handle* h, Slot& sl)
{
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
        qres != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return SSIMU2_E_INTERNAL;
    }
    EncodeTiledFn enc = (EncodeTiledFn)fn;
    const Geo& g = h->geo;
    for (int s = 0; s < g.nscales; s++) {
        float* xyb = sl.xyb + g.sc[s].xyb_off;
        int r;
        if (h->pipeline == 0) {
            r = make_plane_map(enc, &sl.maps_x.xyb_in[s], xyb, g.sc[s], 6, g.xyb_stride, h->batch, kXInW, kXR, false);
            if (r) return r;
            continue;
        }
        float* hb = sl.hb + g.sc[s].hb_off;
        r = make_plane_map(enc, &sl.maps.hb[s], hb, g.sc[s], 15, g.hb_stride, h->batch, kVCols, kVRowsPerStage, false);
        if (r) return r;
        r = make_plane_map(enc, &sl.maps.xyb[s], xyb, g.sc[s], 6, g.xyb_stride, h->batch, kVCols, kVRowsPerStage, false);
        if (r) return r;
        r = make_plane_map(enc, &sl.maps_h.xyb_in[s], xyb, g.sc[s], 6, g.xyb_stride, h->batch, kHCols, kHRows, true);
        if (r) return r;
        r = make_plane_map(enc, &sl.maps_h.hb_out[s], hb, g.sc[s], 15, g.hb_stride, h->batch, kHCols, kHRows, true);
        if (r) return r;
    }
    return 0;
}

// ---- launch logic -------------------------------------------------------------------------------
// ring == 1 (or SSIMU2_NO_FUSE): the four kernels of a batch run back to back on the slot's stream, with
//   CUDA events between them (this is the mode bench.py uses to time each kernel alone).
// ring >= 2: software pipeline across batches.  The front-end of batch k is launched in the SAME kernel as the
//   H pass of batch k-1 (k_fused_fh: compute-bound and memory-bound CTAs share the SMs) on the handle's main
//   stream; the V pass + finalize + result copy of k-1 follow on that slot's own stream.  The last batch of a
//   burst is completed by flush / get_score with an H-only launch.
template <int FMT>
static int launch_unfused(ssimu2_handle* h, Slot& sl)
{
    const Geo& g = h->geo;
    const uint32_t n = sl.count;
    cudaStream_t st = sl.stream;
    if (sl.timed) cudaEventRecord(sl.ev_k[0], st);
    if (h->frontend2) {
        const int rx = (g.sc[0].w + kF2Region - 1) / kF2Region, ry = (g.sc[0].h + kF2Region - 1) / kF2Region;
        const int per_cta = (kF2Threads / 32) * kF2RegionsPerWarp;
        k_frontend2<FMT><<<dim3((rx + per_cta - 1) / per_cta, ry, n), kF2Threads, 0, st>>>(g, sl.in, sl.xyb);
    } else {
        dim3 grid((g.sc[0].w + 63) / 64, (g.sc[0].h + 63) / 64, n);
        k_frontend<FMT><<<grid, kFThreads, kFSmemTotal, st>>>(g, sl.in, sl.xyb);
    }
    if (sl.timed) cudaEventRecord(sl.ev_k[1], st);
    k_hpass<<<dim3(g.items_h, n), kHThreads, kHSmemBytes, st>>>(g, sl.maps_h);
    if (sl.timed) cudaEventRecord(sl.ev_k[2], st);
    k_vpass<<<dim3(g.items_v, n), kVTmaThreads, kVSmemBytes, st>>>(g, sl.maps, sl.partials);
    if (sl.timed) cudaEventRecord(sl.ev_k[3], st);
    k_finalize<<<n, 128, 0, st>>>(g, sl.partials, sl.norms_d, h->scores_ring_d, sl.first_ticket, kResultCap, sl.scores_d, nullptr);
    if (sl.timed) cudaEventRecord(sl.ev_k[4], st);
    h->launches += 4;
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(sl.scores_h, sl.scores_d, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(sl.norms_h, sl.norms_d, (size_t)n * 108 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaEventRecord(sl.ev_done, st));
    sl.inflight = true;
    sl.was_timed = sl.timed;
    return 0;
}

// pipeline "hv": front-end, fused H+V kernel, finalize.
//   ring == 1 : back to back on the slot's stream, with timing events between the kernels.
//   ring >= 2 : two handle-wide streams.  The persistent front-end of batch k+1 (f_stream) runs WHILE the H+V kernel
//               of batch k (hv_stream) does: both are sized to share every SM (see k_frontend2p).
template <int FMT>
static int launch_hv(ssimu2_handle* h, Slot& sl)
{
    const Geo& g = h->geo;
    const uint32_t n = sl.count;
    HvArgs a{};
    a.hstate = sl.hstate;
    a.flags = sl.hvflags;
    a.ticket = sl.hvflags + (size_t)h->batch * g.total_recs;
    a.partials = sl.partials;
    a.epoch = ++sl.epoch;
    a.nframes = (int)n;
    cudaStream_t fs = h->two_stream ? h->f_stream : sl.stream;
    cudaStream_t st = h->two_stream ? h->hv_stream : sl.stream;
    // timing marks: [0,1] around the front-end (on its stream), [2,3] around k_hv, [3,4] around finalize
    if (sl.timed) cudaEventRecord(sl.ev_k[0], fs);
    if (h->two_stream) {
        k_frontend2p<FMT><<<h->num_sms * h->f2p_ctas, kF2PThreads, 0, fs>>>(g, sl.in, sl.xyb, a.ticket + 1, (int)n);
    } else if (h->frontend2) {
        const int rx = (g.sc[0].w + kF2Region - 1) / kF2Region, ry = (g.sc[0].h + kF2Region - 1) / kF2Region;
        const int per_cta = (kF2Threads / 32) * kF2RegionsPerWarp;
        k_frontend2<FMT><<<dim3((rx + per_cta - 1) / per_cta, ry, n), kF2Threads, 0, fs>>>(g, sl.in, sl.xyb);
    } else {
        dim3 grid((g.sc[0].w + 63) / 64, (g.sc[0].h + 63) / 64, n);
        k_frontend<FMT><<<grid, kFThreads, kFSmemTotal, fs>>>(g, sl.in, sl.xyb);
    }
    if (sl.timed) cudaEventRecord(sl.ev_k[1], fs);
    if (h->two_stream) {
        CU_TRY(cudaEventRecord(sl.ev_f, fs));
        CU_TRY(cudaStreamWaitEvent(st, sl.ev_f, 0));
    }
    if (sl.timed) cudaEventRecord(sl.ev_k[2], st);
    k_hv<<<(unsigned)(g.items_v * n), kXThreads, kXSmemBytes, st>>>(g, sl.maps_x, a);
    if (sl.timed) cudaEventRecord(sl.ev_k[3], st);
    k_finalize<<<n, 128, 0, st>>>(g, sl.partials, sl.norms_d, h->scores_ring_d, sl.first_ticket, kResultCap, sl.scores_d, a.ticket);
    if (sl.timed) cudaEventRecord(sl.ev_k[4], st);
    h->launches += 3;
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(sl.scores_h, sl.scores_d, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(sl.norms_h, sl.norms_d, (size_t)n * 108 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaEventRecord(sl.ev_done, st));
    sl.inflight = true;
    sl.was_timed = sl.timed;
    return 0;
}

// [H pass of slot hp] + [front-end of slot fk] in one launch on the main stream; either may be null
template <int FMT>
static int launch_fused(ssimu2_handle* h, Slot* hp, Slot* fk)
{
    const Geo& g = h->geo;
    FuseArgs fa{};
    fa.frames_h = hp ? (int)hp->count : 1;
    fa.n_h_blocks = hp ? g.items_h * (int)hp->count : 0;
    fa.tiles_x = (g.sc[0].w + 63) / 64;
    fa.tiles_y = (g.sc[0].h + 63) / 64;
    fa.n_f_blocks = fk ? fa.tiles_x * fa.tiles_y * (int)fk->count : 0;
    const unsigned total = (unsigned)(fa.n_h_blocks + fa.n_f_blocks);
    if (total == 0) return 0;
    Slot& any = hp ? *hp : *fk;
    k_fused_fh<FMT><<<total, kHThreads, kHSmemBytes, h->main_stream>>>(g, hp ? hp->maps_h : any.maps_h, fk ? fk->in : any.in,
                                                                    fk ? fk->xyb : any.xyb, fa);
    h->launches += 1;
    CU_TRY(cudaGetLastError());
    return 0;
}

// V pass + finalize + result copy of a slot whose H pass has been enqueued on the main stream
static int launch_tail(ssimu2_handle* h, Slot& sl)
{
    const Geo& g = h->geo;
    const uint32_t n = sl.count;
    CU_TRY(cudaEventRecord(sl.ev_mid, h->main_stream));
    CU_TRY(cudaStreamWaitEvent(sl.stream, sl.ev_mid, 0));
    cudaStream_t st = sl.stream;
    k_vpass<<<dim3(g.items_v, n), kVTmaThreads, kVSmemBytes, st>>>(g, sl.maps, sl.partials);
    k_finalize<<<n, 128, 0, st>>>(g, sl.partials, sl.norms_d, h->scores_ring_d, sl.first_ticket, kResultCap, sl.scores_d, nullptr);
    h->launches += 2;
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(sl.scores_h, sl.scores_d, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(sl.norms_h, sl.norms_d, (size_t)n * 108 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaEventRecord(sl.ev_done, st));
    sl.inflight = true;
    sl.was_timed = false;
    return 0;
}

template <int FMT>
static int launch_batch_fmt(ssimu2_handle* h, Slot& sl, int si)
{
    if (h->pipeline == 0) return launch_hv<FMT>(h, sl);
    if (!h->fuse) return launch_unfused<FMT>(h, sl);
    Slot* prev = h->awaiting >= 0 ? &h->slots[h->awaiting] : nullptr;
    int r = launch_fused<FMT>(h, prev, &sl);
    if (r) return r;
    if (prev) {
        r = launch_tail(h, *prev);
        if (r) return r;
        prev->awaiting = false;
    }
    sl.awaiting = true;
    h->awaiting = si;
    return 0;
}

// the slot whose front-end ran but whose H pass is still waiting for a partner: finish it alone
template <int FMT>
static int complete_awaiting_fmt(ssimu2_handle* h)
{
    if (h->awaiting < 0) return 0;
    Slot& sl = h->slots[h->awaiting];
    int r = launch_fused<FMT>(h, &sl, nullptr);
    if (r) return r;
    r = launch_tail(h, sl);
    if (r) return r;
    sl.awaiting = false;
    h->awaiting = -1;
    return 0;
}

static int complete_awaiting(ssimu2_handle* h)
{
#define CALL_(F) complete_awaiting_fmt<F>(h)
    switch (h->cfg.format) {
    case kNV12: return CALL_(kNV12);
    case kP016: return CALL_(kP016);
    case kSRGB8: return CALL_(kSRGB8);
    case kSRGB16: return CALL_(kSRGB16);
    case kSRGBF32: return CALL_(kSRGBF32);
    case kLINEARF32: return CALL_(kLINEARF32);
    }
#undef CALL_
    return SSIMU2_E_UNSUPPORTED;
}

// launch the batch recorded in slot si (it must hold at least one pair and not be launched yet)
static int launch_batch(ssimu2_handle* h, int si)
{
    Slot& sl = h->slots[si];
    if (sl.count == 0 || sl.inflight || sl.awaiting) return 0;
    cudaStream_t first = h->fuse ? h->main_stream : (h->two_stream ? h->f_stream : sl.stream);
    if (sl.have_dep && sl.last_stream != (void*)first) {
        // order the batch after everything the submitter enqueued so far on its stream
        CU_TRY(cudaEventRecord(sl.ev_in, (cudaStream_t)sl.last_stream));
        CU_TRY(cudaStreamWaitEvent(first, sl.ev_in, 0));
    }
    if ((h->fuse || h->two_stream) && sl.staged) {
        // host frames were copied on the slot's stream: the front-end on the shared stream must see them
        CU_TRY(cudaEventRecord(sl.ev_in, sl.stream));
        CU_TRY(cudaStreamWaitEvent(first, sl.ev_in, 0));
    }
    sl.have_dep = false;
    sl.staged = false;
#define CALL_(F) launch_batch_fmt<F>(h, sl, si)
    switch (h->cfg.format) {
    case kNV12: return CALL_(kNV12);
    case kP016: return CALL_(kP016);
    case kSRGB8: return CALL_(kSRGB8);
    case kSRGB16: return CALL_(kSRGB16);
    case kSRGBF32: return CALL_(kSRGBF32);
    case kLINEARF32: return CALL_(kLINEARF32);
    }
#undef CALL_
    return SSIMU2_E_UNSUPPORTED;
}

// wait for a launched slot and move its results into the ticket-indexed host rings
static int harvest(ssimu2_handle* h, uint32_t si)
{
    Slot& sl = h->slots[si];
    if (!sl.inflight) return 0;
    CU_TRY(cudaEventSynchronize(sl.ev_done));
    for (uint32_t i = 0; i < sl.count; i++) {
        uint64_t t = sl.first_ticket + i, r = t % kResultCap;
        h->res_scores[r] = sl.scores_h[i];
        memcpy(&h->res_norms[r * 108], &sl.norms_h[(size_t)i * 108], 108 * sizeof(double));
        h->res_slot[r] = (int32_t)si;
    }
    if (sl.was_timed) {
        for (int k = 0; k < 4; k++) {
            if (h->pipeline == 0) {
                // front-end [0,1], k_hv [2,3], (no separate V pass), finalize [3,4]
                static const int a[4] = {0, 2, 3, 3}, b[4] = {1, 3, 3, 4};
                h->last_ms[k] = 0.f;
                if (a[k] != b[k]) cudaEventElapsedTime(&h->last_ms[k], sl.ev_k[a[k]], sl.ev_k[b[k]]);
            } else {
                cudaEventElapsedTime(&h->last_ms[k], sl.ev_k[k], sl.ev_k[k + 1]);
            }
            h->total_ms[k] += h->last_ms[k];
        }
        h->timed_batches++;
        h->timed_pairs += sl.count;
    }
    sl.inflight = false;
    sl.count = 0;
    return 0;
}

// make slot `cur` ready to record a pair
static int prepare_cur(ssimu2_handle* h)
{
    Slot& sl = h->slots[h->cur];
    if (sl.awaiting) {  // the ring wrapped onto the batch whose H pass is still parked: finish it first
        int r = complete_awaiting(h);
        if (r) return r;
    }
    if (sl.inflight) {
        int r = harvest(h, h->cur);
        if (r) return r;
    }
    if (sl.count == 0) {
        sl.first_ticket = h->next_ticket;
        sl.in.first_ticket = h->next_ticket;
        sl.have_dep = false;
        sl.last_stream = nullptr;
    }
    return 0;
}

static int finish_pair(ssimu2_handle* h)
{
    Slot& sl = h->slots[h->cur];
    sl.count++;
    h->next_ticket++;
    if (sl.count == h->batch) {
        int r = launch_batch(h, (int)h->cur);
        if (r) return r;
        h->cur = (h->cur + 1) % h->ring;
    }
    return 0;
}

static bool frame_ok(const ssimu2_handle* h, const ssimu2_frame* f)
{
    if (!f || !f->plane[0] || f->pitch == 0) return false;
    if ((h->cfg.format == kNV12 || h->cfg.format == kP016) && !f->plane[1]) return false;
    return true;
}

// locate a ticket: 0 = results on host, 1 = in a launched slot, 2 = in the slot being filled
static int locate(ssimu2_handle* h, uint64_t ticket, uint32_t* slot_out)
{
    if (ticket >= h->next_ticket) return -1;
    if (h->next_ticket - ticket > kResultCap) return -1;
    for (uint32_t i = 0; i < h->ring; i++) {
        Slot& sl = h->slots[i];
        if (sl.count && ticket >= sl.first_ticket && ticket < sl.first_ticket + sl.count) {
            *slot_out = i;
            return sl.inflight ? 1 : 2;
        }
    }
    return 0;
}

}  // namespace

extern "C" {

uint32_t ssimu2_version(void) { return (1u << 16) | 0u; }

const char* ssimu2_strerror(int status)
{
    switch (status) {
    case SSIMU2_OK: return "ok";
    case SSIMU2_E_INVALID: return "invalid argument";
    case SSIMU2_E_UNSUPPORTED: return "unsupported format or size";
    case SSIMU2_E_NOMEM: return "out of memory";
    case SSIMU2_E_NODEVICE: return "no usable CUDA device (sm_100 required)";
    case SSIMU2_E_TICKET: return "unknown or expired ticket";
    case SSIMU2_E_INTERNAL: return "internal error";
    }
    if (status > 0) return cudaGetErrorString((cudaError_t)status);
    return "unknown status";
}

int ssimu2_create(ssimu2_t** out, const ssimu2_config* cfg)
{
    if (!out || !cfg) return SSIMU2_E_INVALID;
    *out = nullptr;
    if (cfg->width < 8 || cfg->height < 8 || cfg->width > 32768 || cfg->height > 32768) return SSIMU2_E_UNSUPPORTED;
    if (cfg->format < 0 || cfg->format > kLINEARF32) return SSIMU2_E_UNSUPPORTED;
    if (cfg->matrix < 0 || cfg->matrix > 2) return SSIMU2_E_UNSUPPORTED;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        return SSIMU2_E_NODEVICE;
    }
    if (cfg->device < 0 || cfg->device >= ndev) return SSIMU2_E_INVALID;
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10) return SSIMU2_E_NODEVICE;  // the only SASS in this library is sm_100a

    ssimu2_handle* h = new (std::nothrow) ssimu2_handle();
    if (!h) return SSIMU2_E_NOMEM;
    h->cfg = *cfg;
    h->batch = cfg->batch ? cfg->batch : kDefaultBatch;
    if (h->batch > (uint32_t)kMaxBatch) h->batch = kMaxBatch;
    h->ring = cfg->ring ? cfg->ring : kDefaultRing;
    if (h->ring > 16) h->ring = 16;
    build_geo(h);
    int rc = 0;
#define CR(expr)                                  \
    do {                                          \
        cudaError_t _e = (expr);                  \
        if (_e != cudaSuccess) {                  \
            rc = (_e == cudaErrorMemoryAllocation) ? SSIMU2_E_NOMEM : (int)_e; \
            goto fail;                            \
        }                                         \
    } while (0)
    CR(cudaSetDevice(cfg->device));
    try {
        h->slots.resize(h->ring);
        h->res_scores.assign(kResultCap, 0.0);
        h->res_norms.assign(kResultCap * 108, 0.0);
        h->res_slot.assign(kResultCap, -1);
    } catch (...) {
        rc = SSIMU2_E_NOMEM;
        goto fail;
    }
    CR(cudaMalloc(&h->scores_ring_d, kResultCap * sizeof(double)));
    CR(cudaMemset(h->scores_ring_d, 0, kResultCap * sizeof(double)));
    h->device_bytes += kResultCap * sizeof(double);
    CR(cudaFuncSetAttribute((const void*)k_hpass, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHSmemBytes));
    CR(cudaFuncSetAttribute((const void*)k_vpass, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kVSmemBytes));
    static_assert(kHSmemBytes >= kFSmemTotal, "the fused kernel runs front-end tiles inside the H pass allocation");
    {
        // SSIMU2_PIPELINE = hv (default) | fh | split; "split" keeps the H-pass planes in HBM (ssimu2_debug_read what = 1)
        const char* pm = getenv("SSIMU2_PIPELINE");
        h->pipeline = (pm && !strcmp(pm, "fh")) ? 1 : ((pm && !strcmp(pm, "split")) ? 2 : 0);
    }
    {
        const char* fe = getenv("SSIMU2_FRONTEND");
        h->frontend2 = !(fe && !strcmp(fe, "1"));
    }
    CR(cudaFuncSetAttribute((const void*)k_hv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kXSmemBytes));
    h->fuse = h->pipeline == 1 && h->ring >= 2 && getenv("SSIMU2_NO_FUSE") == nullptr;
    h->two_stream = h->pipeline == 0 && h->ring >= 2 && getenv("SSIMU2_TWO_STREAM") != nullptr;  // experimental, off by default
    h->num_sms = prop.multiProcessorCount;
    if (const char* e = getenv("SSIMU2_F2P_CTAS")) h->f2p_ctas = atoi(e) > 0 ? atoi(e) : 1;
    if (h->two_stream) {
        int lo = 0, hi = 0;
        CR(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CR(cudaStreamCreateWithPriority(&h->f_stream, cudaStreamNonBlocking, lo));
        CR(cudaStreamCreateWithPriority(&h->hv_stream, cudaStreamNonBlocking, hi));  // k_hv CTAs are placed first
    }
    CR(cudaStreamCreateWithFlags(&h->main_stream, cudaStreamNonBlocking));
    {
        static const void* ffn[6] = {(const void*)k_frontend<kNV12>,    (const void*)k_frontend<kP016>,
                                     (const void*)k_frontend<kSRGB8>,   (const void*)k_frontend<kSRGB16>,
                                     (const void*)k_frontend<kSRGBF32>, (const void*)k_frontend<kLINEARF32>};
        CR(cudaFuncSetAttribute(ffn[cfg->format], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFSmemTotal));
        static const void* gfn[6] = {(const void*)k_fused_fh<kNV12>,    (const void*)k_fused_fh<kP016>,
                                     (const void*)k_fused_fh<kSRGB8>,   (const void*)k_fused_fh<kSRGB16>,
                                     (const void*)k_fused_fh<kSRGBF32>, (const void*)k_fused_fh<kLINEARF32>};
        CR(cudaFuncSetAttribute(gfn[cfg->format], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHSmemBytes));
    }
    if ((cfg->format == kNV12 || cfg->format == kP016) && getenv("SSIMU2_NO_LUT") == nullptr) {
        const int n = cfg->format == kNV12 ? 256 : 1024, shift = cfg->format == kNV12 ? 0 : 6;
        const size_t bytes = (size_t)2 * n * n * sizeof(float);
        CR(cudaMalloc(&h->eotf_lut, bytes));
        k_build_eotf_lut<<<(n * n + 255) / 256, 256, 0, h->main_stream>>>(h->geo.coef, n, shift, h->eotf_lut);
        CR(cudaGetLastError());
        CR(cudaStreamSynchronize(h->main_stream));
        h->geo.eotf_lut = h->eotf_lut;
        h->geo.lut_n = n;
        h->geo.lut_shift = shift;
        h->device_bytes += bytes;
    }
    for (uint32_t i = 0; i < h->ring; i++) {
        Slot& sl = h->slots[i];
        const Geo& g = h->geo;
        size_t xyb_b = (size_t)g.xyb_stride * h->batch * sizeof(float);
        size_t hb_b = h->pipeline == 0 ? 0 : (size_t)g.hb_stride * h->batch * sizeof(float);
        size_t hs_b = h->pipeline == 0 ? (size_t)g.total_recs * h->batch * kXHsBytes : 0;
        size_t fl_b = h->pipeline == 0 ? ((size_t)g.total_recs * h->batch + 2) * sizeof(uint32_t) : 0;
        size_t part_b = (size_t)g.total_strips * 18 * h->batch * sizeof(double);
        CR(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
        CR(cudaEventCreateWithFlags(&sl.ev_in, cudaEventDisableTiming));
        CR(cudaEventCreateWithFlags(&sl.ev_done, cudaEventDisableTiming));
        CR(cudaEventCreateWithFlags(&sl.ev_mid, cudaEventDisableTiming));
        CR(cudaEventCreateWithFlags(&sl.ev_f, cudaEventDisableTiming));
        for (int k = 0; k < 5; k++) CR(cudaEventCreate(&sl.ev_k[k]));
        CR(cudaMalloc(&sl.xyb, xyb_b));
        if (hb_b) CR(cudaMalloc(&sl.hb, hb_b));
        if (hs_b) {
            CR(cudaMalloc(&sl.hstate, hs_b));
            CR(cudaMalloc(&sl.hvflags, fl_b));
            CR(cudaMemset(sl.hvflags, 0, fl_b));
        }
        CR(cudaMalloc(&sl.partials, part_b));
        CR(cudaMalloc(&sl.norms_d, (size_t)h->batch * 108 * sizeof(double)));
        CR(cudaMalloc(&sl.scores_d, (size_t)h->batch * sizeof(double)));
        CR(cudaMallocHost(&sl.norms_h, (size_t)h->batch * 108 * sizeof(double)));
        CR(cudaMallocHost(&sl.scores_h, (size_t)h->batch * sizeof(double)));
        h->device_bytes += xyb_b + hb_b + hs_b + fl_b + part_b + (size_t)h->batch * 109 * sizeof(double);
        rc = build_tma_maps(h, sl);
        if (rc) goto fail;
        sl.timed = !h->fuse && getenv("SSIMU2_NO_TIMING") == nullptr;
    }
    *out = h;
    return SSIMU2_OK;
fail:
    ssimu2_destroy(h);
    return rc;
#undef CR
}

int ssimu2_destroy(ssimu2_t* h)
{
    if (!h) return SSIMU2_OK;
    cudaSetDevice(h->cfg.device);
    if (h->main_stream) cudaStreamSynchronize(h->main_stream);
    if (h->f_stream) cudaStreamSynchronize(h->f_stream);
    if (h->hv_stream) cudaStreamSynchronize(h->hv_stream);
    for (auto& sl : h->slots) {
        if (sl.stream) cudaStreamSynchronize(sl.stream);
        if (sl.ev_mid) cudaEventDestroy(sl.ev_mid);
        if (sl.ev_f) cudaEventDestroy(sl.ev_f);
        if (sl.ev_in) cudaEventDestroy(sl.ev_in);
        if (sl.ev_done) cudaEventDestroy(sl.ev_done);
        for (int k = 0; k < 5; k++)
            if (sl.ev_k[k]) cudaEventDestroy(sl.ev_k[k]);
        cudaFree(sl.xyb); cudaFree(sl.hb); cudaFree(sl.hstate); cudaFree(sl.hvflags); cudaFree(sl.partials); cudaFree(sl.norms_d); cudaFree(sl.scores_d);
        cudaFree(sl.staging);
        if (sl.norms_h) cudaFreeHost(sl.norms_h);
        if (sl.scores_h) cudaFreeHost(sl.scores_h);
        if (sl.stream) cudaStreamDestroy(sl.stream);
    }
    cudaFree(h->scores_ring_d);
    cudaFree(h->eotf_lut);
    if (h->main_stream) cudaStreamDestroy(h->main_stream);
    if (h->f_stream) cudaStreamDestroy(h->f_stream);
    if (h->hv_stream) cudaStreamDestroy(h->hv_stream);
    delete h;
    return SSIMU2_OK;
}

int ssimu2_mem_usage(const ssimu2_t* h, size_t* bytes)
{
    if (!h || !bytes) return SSIMU2_E_INVALID;
    *bytes = h->device_bytes;
    return SSIMU2_OK;
}

int ssimu2_submit(ssimu2_t* h, const ssimu2_frame* ref, const ssimu2_frame* dis, void* stream, uint64_t* ticket)
{
    if (!h || !frame_ok(h, ref) || !frame_ok(h, dis)) return SSIMU2_E_INVALID;
    CU_TRY(cudaSetDevice(h->cfg.device));
    int r = prepare_cur(h);
    if (r) return r;
    Slot& sl = h->slots[h->cur];
    cudaStream_t first = h->fuse ? h->main_stream : (h->two_stream ? h->f_stream : sl.stream);
    if (sl.have_dep && sl.last_stream != stream && sl.last_stream != (void*)first) {
        // submitter switched streams inside one batch: pin the dependency on the previous one now
        CU_TRY(cudaEventRecord(sl.ev_in, (cudaStream_t)sl.last_stream));
        CU_TRY(cudaStreamWaitEvent(first, sl.ev_in, 0));
    }
    sl.last_stream = stream;
    sl.have_dep = true;
    FrameIn& a = sl.in.ref[sl.count];
    FrameIn& b = sl.in.dis[sl.count];
    a.p0 = (const uint8_t*)ref->plane[0]; a.p1 = (const uint8_t*)ref->plane[1]; a.pitch = ref->pitch; a.pad = 0;
    b.p0 = (const uint8_t*)dis->plane[0]; b.p1 = (const uint8_t*)dis->plane[1]; b.pitch = dis->pitch; b.pad = 0;
    if (ticket) *ticket = h->next_ticket;
    return finish_pair(h);
}

int ssimu2_submit_batch(ssimu2_t* h, uint32_t n, const ssimu2_frame* refs, const ssimu2_frame* diss, void* stream,
                        uint64_t* first_ticket)
{
    if (!h || (n && (!refs || !diss))) return SSIMU2_E_INVALID;
    if (first_ticket) *first_ticket = h->next_ticket;
    for (uint32_t i = 0; i < n; i++) {
        int r = ssimu2_submit(h, &refs[i], &diss[i], stream, nullptr);
        if (r) return r;
    }
    return SSIMU2_OK;
}

int ssimu2_submit_host(ssimu2_t* h, const ssimu2_frame* ref, const ssimu2_frame* dis, size_t frame_bytes, uint64_t* ticket)
{
    if (!h || !frame_ok(h, ref) || !frame_ok(h, dis) || frame_bytes == 0) return SSIMU2_E_INVALID;
    const bool yuv = h->cfg.format == kNV12 || h->cfg.format == kP016;
    if (yuv && (ref->plane[1] <= ref->plane[0] || dis->plane[1] <= dis->plane[0] ||
                ref->plane[1] - ref->plane[0] >= frame_bytes || dis->plane[1] - dis->plane[0] >= frame_bytes))
        return SSIMU2_E_INVALID;
    CU_TRY(cudaSetDevice(h->cfg.device));
    if (frame_bytes > h->staging_frame_bytes) {
        // (re)allocate the staging rings; rare (first call), so a full drain is acceptable
        int fr = ssimu2_flush(h);
        if (fr) return fr;
        for (uint32_t i = 0; i < h->ring; i++) {
            int r = harvest(h, i);
            if (r) return r;
        }
        size_t fb = (frame_bytes + 255) / 256 * 256;
        for (auto& sl : h->slots) {
            cudaFree(sl.staging);
            sl.staging = nullptr;
            if (cudaMalloc(&sl.staging, fb * 2 * h->batch) != cudaSuccess) { cudaGetLastError(); return SSIMU2_E_NOMEM; }
        }
        h->device_bytes += (fb - h->staging_frame_bytes) * 2 * h->batch * h->ring;
        h->staging_frame_bytes = fb;
    }
    int r = prepare_cur(h);
    if (r) return r;
    Slot& sl = h->slots[h->cur];
    uint8_t* da = sl.staging + (size_t)(2 * sl.count) * h->staging_frame_bytes;
    uint8_t* db = da + h->staging_frame_bytes;
    CU_TRY(cudaMemcpyAsync(da, (const void*)ref->plane[0], frame_bytes, cudaMemcpyHostToDevice, sl.stream));
    CU_TRY(cudaMemcpyAsync(db, (const void*)dis->plane[0], frame_bytes, cudaMemcpyHostToDevice, sl.stream));
    FrameIn& a = sl.in.ref[sl.count];
    FrameIn& b = sl.in.dis[sl.count];
    a.p0 = da; a.p1 = yuv ? da + (ref->plane[1] - ref->plane[0]) : nullptr; a.pitch = ref->pitch; a.pad = 0;
    b.p0 = db; b.p1 = yuv ? db + (dis->plane[1] - dis->plane[0]) : nullptr; b.pitch = dis->pitch; b.pad = 0;
    sl.staged = true;
    if (ticket) *ticket = h->next_ticket;
    return finish_pair(h);
}

int ssimu2_submit_host_batch(ssimu2_t* h, uint32_t n, const ssimu2_frame* refs, const ssimu2_frame* diss, size_t frame_bytes,
                             uint64_t* first_ticket)
{
    if (!h || (n && (!refs || !diss))) return SSIMU2_E_INVALID;
    if (first_ticket) *first_ticket = h->next_ticket;
    for (uint32_t i = 0; i < n; i++) {
        int r = ssimu2_submit_host(h, &refs[i], &diss[i], frame_bytes, nullptr);
        if (r) return r;
    }
    return SSIMU2_OK;
}

int ssimu2_flush(ssimu2_t* h)
{
    if (!h) return SSIMU2_E_INVALID;
    CU_TRY(cudaSetDevice(h->cfg.device));
    Slot& sl = h->slots[h->cur];
    if (sl.count && !sl.inflight && !sl.awaiting) {
        int r = launch_batch(h, (int)h->cur);
        if (r) return r;
        h->cur = (h->cur + 1) % h->ring;
    }
    // nothing may stay parked after a flush: the last batch gets its H pass without a partner
    return h->fuse ? complete_awaiting(h) : SSIMU2_OK;
}

int ssimu2_wait(ssimu2_t* h, uint64_t ticket)
{
    if (!h) return SSIMU2_E_INVALID;
    CU_TRY(cudaSetDevice(h->cfg.device));
    uint32_t si = 0;
    int where = locate(h, ticket, &si);
    if (where < 0) return SSIMU2_E_TICKET;
    if (where == 2) {
        // not fully launched yet: a partial batch still being filled, or a batch parked between its front-end and
        // its H pass.  Launch what is missing for THIS ticket only.
        Slot& sl = h->slots[si];
        if (!sl.awaiting) {
            int r = launch_batch(h, (int)si);
            if (r) return r;
            if (si == h->cur) h->cur = (h->cur + 1) % h->ring;
        }
        if (sl.awaiting) {
            int r = complete_awaiting(h);
            if (r) return r;
        }
        where = 1;
    }
    if (where == 1) return harvest(h, si);
    return SSIMU2_OK;
}

int ssimu2_get_score(ssimu2_t* h, uint64_t ticket, double* score)
{
    if (!h || !score) return SSIMU2_E_INVALID;
    int r = ssimu2_wait(h, ticket);
    if (r) return r;
    *score = h->res_scores[ticket % kResultCap];
    return SSIMU2_OK;
}

int ssimu2_get_scores(ssimu2_t* h, uint64_t first_ticket, uint32_t n, double* scores)
{
    if (!h || (n && !scores)) return SSIMU2_E_INVALID;
    for (uint32_t i = 0; i < n; i++) {
        int r = ssimu2_wait(h, first_ticket + i);
        if (r) return r;
        scores[i] = h->res_scores[(first_ticket + i) % kResultCap];
    }
    return SSIMU2_OK;
}

int ssimu2_get_norms(ssimu2_t* h, uint64_t ticket, double* norms108)
{
    if (!h || !norms108) return SSIMU2_E_INVALID;
    int r = ssimu2_wait(h, ticket);
    if (r) return r;
    memcpy(norms108, &h->res_norms[(ticket % kResultCap) * 108], 108 * sizeof(double));
    return SSIMU2_OK;
}

int ssimu2_compute_sync(ssimu2_t* h, const ssimu2_frame* ref, const ssimu2_frame* dis, void* stream, double* score)
{
    uint64_t t = 0;
    int r = ssimu2_submit(h, ref, dis, stream, &t);
    if (r) return r;
    return ssimu2_get_score(h, t, score);
}

int ssimu2_stream_wait(ssimu2_t* h, uint64_t ticket, void* stream)
{
    if (!h) return SSIMU2_E_INVALID;
    CU_TRY(cudaSetDevice(h->cfg.device));
    uint32_t si = 0;
    int where = locate(h, ticket, &si);
    if (where < 0) return SSIMU2_E_TICKET;
    if (where == 2) {
        Slot& sl = h->slots[si];
        if (!sl.awaiting) {
            int r = launch_batch(h, (int)si);
            if (r) return r;
            if (si == h->cur) h->cur = (h->cur + 1) % h->ring;
        }
        if (sl.awaiting) {
            int r = complete_awaiting(h);
            if (r) return r;
        }
        where = 1;
    }
    if (where == 1) CU_TRY(cudaStreamWaitEvent((cudaStream_t)stream, h->slots[si].ev_done, 0));
    return SSIMU2_OK;
}

int ssimu2_scores_device(ssimu2_t* h, uint64_t* dptr, uint64_t* capacity)
{
    if (!h || !dptr || !capacity) return SSIMU2_E_INVALID;
    *dptr = (uint64_t)(uintptr_t)h->scores_ring_d;
    *capacity = kResultCap;
    return SSIMU2_OK;
}

int ssimu2_get_info(const ssimu2_t* h, ssimu2_info* info)
{
    if (!h || !info) return SSIMU2_E_INVALID;
    memset(info, 0, sizeof(*info));
    info->nscales = (uint32_t)h->geo.nscales;
    for (int s = 0; s < h->geo.nscales; s++) {
        info->width[s] = (uint32_t)h->geo.sc[s].w;
        info->height[s] = (uint32_t)h->geo.sc[s].h;
        info->pitch[s] = (uint32_t)h->geo.sc[s].pitch;
    }
    info->batch = h->batch;
    info->ring = h->ring;
    info->alg_bytes_per_pair = h->alg_bytes;
    info->kernel_launches = h->launches;
    return SSIMU2_OK;
}

int ssimu2_debug_read(ssimu2_t* h, uint64_t ticket, int what, int scale, float* out, size_t out_floats)
{
    if (!h || !out || scale < 0 || scale >= h->geo.nscales) return SSIMU2_E_INVALID;
    int r = ssimu2_wait(h, ticket);
    if (r) return r;
    int si = h->res_slot[ticket % kResultCap];
    if (si < 0) return SSIMU2_E_TICKET;
    Slot& sl = h->slots[si];
    if (ticket < sl.first_ticket) return SSIMU2_E_TICKET;
    size_t idx = (size_t)(ticket - sl.first_ticket);
    if (idx >= h->batch) return SSIMU2_E_TICKET;
    const ScaleDesc& sd = h->geo.sc[scale];
    int planes;
    const float* src;
    if (what == 0) {
        planes = 6;
        src = sl.xyb + idx * h->geo.xyb_stride + sd.xyb_off;
    } else if (what == 1) {
        if (!sl.hb) return SSIMU2_E_UNSUPPORTED;  // the fused H+V pipeline never materialises these planes
        planes = 15;
        src = sl.hb + idx * h->geo.hb_stride + sd.hb_off;
    } else {
        return SSIMU2_E_INVALID;
    }
    if (out_floats < (size_t)planes * sd.w * sd.h) return SSIMU2_E_INVALID;
    CU_TRY(cudaMemcpy2D(out, (size_t)sd.w * sizeof(float), src, (size_t)sd.pitch * sizeof(float),
                        (size_t)sd.w * sizeof(float), (size_t)planes * sd.h, cudaMemcpyDeviceToHost));
    return SSIMU2_OK;
}

int ssimu2_debug_math(int op, const float* in, float y, float* out, size_t n)
{
    if (!in || !out || op < 0 || op > 6) return SSIMU2_E_INVALID;
    if (n == 0) return SSIMU2_OK;
    const size_t in_b = op == 3 ? n * 16 : (op == 4 ? n * 8 : n * 4), out_b = op == 3 ? n * 8 : n * 4;
    float *din = nullptr, *dout = nullptr;
    if (cudaMalloc(&din, in_b) != cudaSuccess || cudaMalloc(&dout, out_b) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(din);
        return SSIMU2_E_NOMEM;
    }
    int rc = (int)cudaMemcpy(din, in, in_b, cudaMemcpyHostToDevice);
    if (!rc) {
        k_debug_math<<<(unsigned)((n + 255) / 256), 256>>>(op, din, y, dout, n);
        rc = (int)cudaGetLastError();
    }
    if (!rc) rc = (int)cudaMemcpy(out, dout, out_b, cudaMemcpyDeviceToHost);
    cudaFree(din);
    cudaFree(dout);
    return rc;
}

int ssimu2_last_batch_ms(ssimu2_t* h, float ms[4])
{
    if (!h || !ms) return SSIMU2_E_INVALID;
    memcpy(ms, h->last_ms, sizeof(h->last_ms));
    return SSIMU2_OK;
}

int ssimu2_kernel_ms(ssimu2_t* h, double ms_total[4], uint64_t* batches, uint64_t* pairs, int reset)
{
    if (!h || !ms_total) return SSIMU2_E_INVALID;
    for (int k = 0; k < 4; k++) ms_total[k] = h->total_ms[k];
    if (batches) *batches = h->timed_batches;
    if (pairs) *pairs = h->timed_pairs;
    if (reset) {
        for (int k = 0; k < 4; k++) h->total_ms[k] = 0;
        h->timed_batches = h->timed_pairs = 0;
    }
    return SSIMU2_OK;
}

}  // extern "C"
