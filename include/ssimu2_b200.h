/*
 * ssimu2_b200.h -- C ABI of the B200-native SSIMULACRA2 frame-pair scorer.
 *
 * Drop-in boundary for the reference's `Ssimulacra2` metric op and the colour front-end
 * that feeds it (paths relative to /root/reference/crates):
 *
 *   ssimulacra2-cuda/src/lib.rs:27-107   struct Ssimulacra2 / Ssimulacra2::new      -> ssimu2_create
 *   ssimulacra2-cuda/src/lib.rs:110      Ssimulacra2::mem_usage                     -> ssimu2_mem_usage
 *   ssimulacra2-cuda/src/lib.rs:283-287  Ssimulacra2::compute (async)               -> ssimu2_submit
 *   ssimulacra2-cuda/src/lib.rs:271-279  Ssimulacra2::compute_sync                  -> ssimu2_compute_sync
 *   ssimulacra2-cuda/src/lib.rs:253-266  Ssimulacra2::compute_srgb_sync             -> ssimu2_compute_sync (SSIMU2_FMT_SRGB8)
 *   ssimulacra2-cuda/src/lib.rs:232-250  Ssimulacra2::compute_from_cpu_srgb_sync    -> ssimu2_submit_host + ssimu2_get_score
 *   ssimulacra2-cuda/src/lib.rs:289-291  Ssimulacra2::get_score                     -> ssimu2_get_score
 *   cuda-colorspace/src/lib.rs:33-123    ColorspaceConversion::biplanaryuv420_to_linearrgb_{8,16}
 *                                        (folded into the scorer: SSIMU2_FMT_NV12 / SSIMU2_FMT_P016)
 *   cuda-colorspace/src/lib.rs:144-169   srgb_to_linear_{u8,u16,f32}                (SSIMU2_FMT_SRGB8/16/F32)
 *   turbo-metrics/src/lib.rs:268-360     TurboMetrics::compute_one (the caller)     -> ssimu2_submit / ssimu2_get_score
 *   turbo-metrics/src/lib.rs:362-433     TurboMetrics::compute_all frame loop       -> ssimu2_submit_batch + tickets
 *
 * Conventions
 *   - Plain C types only.  Device pointers are CUdeviceptr-compatible 64-bit integers,
 *     streams are CUstream / cudaStream_t handles passed as void*.
 *   - Every function returns 0 (SSIMU2_OK) or a negative ssimu2_status; a positive value is
 *     a CUDA runtime error code passed through.  Nothing throws or aborts across the ABI.
 *   - Input frames are BORROWED: read-only, never freed, and must stay valid until the
 *     ticket's score has been fetched (or ssimu2_wait has returned for it).
 *   - A handle is bound to one device and one (width, height, format).  Calls on one handle
 *     are not re-entrant; different handles may be driven from different threads.
 *   - There is no CPU fallback: if no CUDA device is usable every entry point fails.
 */
#ifndef SSIMU2_B200_H
#define SSIMU2_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ssimu2_handle ssimu2_t;

typedef enum {
    SSIMU2_OK = 0,
    SSIMU2_E_INVALID = -1,     /* bad argument */
    SSIMU2_E_UNSUPPORTED = -2, /* format / size not supported (reference: todo!() panics) */
    SSIMU2_E_NOMEM = -3,
    SSIMU2_E_NODEVICE = -4, /* no CUDA device / not an sm_100 part */
    SSIMU2_E_TICKET = -5,   /* ticket unknown or already overwritten in the result ring */
    SSIMU2_E_INTERNAL = -6
} ssimu2_status;

/* Pixel formats crossing the boundary (turbo-metrics/src/lib.rs:125-130 `HwFrame`). */
typedef enum {
    SSIMU2_FMT_NV12 = 0,      /* NvDecFrame::NV12: u8 Y plane + interleaved u8 CbCr plane, 4:2:0 */
    SSIMU2_FMT_P016 = 1,      /* NvDecFrame::P016: u16 MSB-aligned Y + interleaved CbCr, 4:2:0 */
    SSIMU2_FMT_SRGB8 = 2,     /* Npp8 : packed RGB u8, sRGB transfer (256-entry table) */
    SSIMU2_FMT_SRGB16 = 3,    /* Npp16: packed RGB u16, sRGB transfer (analytic) */
    SSIMU2_FMT_SRGBF32 = 4,   /* Npp32: packed RGB f32 in [0,1], sRGB transfer (analytic) */
    SSIMU2_FMT_LINEARF32 = 5  /* packed linear RGB f32: the input of Ssimulacra2::new itself */
} ssimu2_format;

/* cuda-colorspace/src/lib.rs:8-13 `ColorMatrix` (only used by the YUV formats). */
typedef enum { SSIMU2_MATRIX_BT709 = 0, SSIMU2_MATRIX_BT601_525 = 1, SSIMU2_MATRIX_BT601_625 = 2 } ssimu2_matrix;

/* One device frame.  YUV 4:2:0: plane[0] = Y, plane[1] = interleaved CbCr
 * (NVDEC: plane[1] = plane[0] + pitch * coded_height, cudarse-video/src/dec.rs:299-366),
 * both with the same pitch.  Packed RGB formats use plane[0] only.  pitch is in bytes. */
typedef struct {
    uint64_t plane[2];
    uint32_t pitch;
    uint32_t reserved;
} ssimu2_frame;

typedef struct {
    uint32_t width, height;
    int32_t format;     /* ssimu2_format */
    int32_t matrix;     /* ssimu2_matrix */
    int32_t full_range; /* 0 = limited (what the reference implements), 1 = full */
    int32_t device;     /* CUDA device ordinal */
    uint32_t batch;     /* frame pairs per kernel launch group (0 = default) */
    uint32_t ring;      /* batches in flight, one stream each (0 = default) */
} ssimu2_config;

/* ---- lifetime ------------------------------------------------------------------------ */
int ssimu2_create(ssimu2_t **out, const ssimu2_config *cfg);
int ssimu2_destroy(ssimu2_t *h);
/* Device bytes owned by the handle (Ssimulacra2::mem_usage). */
int ssimu2_mem_usage(const ssimu2_t *h, size_t *bytes);
const char *ssimu2_strerror(int status);
/* Library / ABI version: (major << 16) | minor. */
uint32_t ssimu2_version(void);

/* ---- scoring: device frames ---------------------------------------------------------- */
/* Enqueue one pair.  Work is ordered after everything already enqueued on `stream` (may be
 * NULL = legacy default stream).  Never blocks on the GPU unless the ring is full.  Pairs are
 * grouped into batches of cfg.batch; a partial batch is launched by ssimu2_flush or by
 * fetching one of its tickets.  Tickets increase by 1 per pair in submission order. */
int ssimu2_submit(ssimu2_t *h, const ssimu2_frame *ref, const ssimu2_frame *dis, void *stream, uint64_t *ticket);
/* n pairs at once; *first_ticket receives the ticket of pair 0 (pair i has first_ticket+i). */
int ssimu2_submit_batch(ssimu2_t *h, uint32_t n, const ssimu2_frame *refs, const ssimu2_frame *diss, void *stream,
                        uint64_t *first_ticket);
/* Launch whatever is pending. */
int ssimu2_flush(ssimu2_t *h);
/* Block until the ticket's batch has completed on the device. */
int ssimu2_wait(ssimu2_t *h, uint64_t ticket);
/* Score of a ticket (flushes and waits as needed).  100 = identical, unbounded below. */
int ssimu2_get_score(ssimu2_t *h, uint64_t ticket, double *score);
/* The scores of n consecutive tickets in submission order: the per-frame score stream of the CLI loop
 * (turbo-metrics-cli/src/main.rs:307-322) in one call. */
int ssimu2_get_scores(ssimu2_t *h, uint64_t first_ticket, uint32_t n, double *scores);
/* The 108 per-scale / per-channel norms in WEIGHT order:
 * index = channel*36 + scale*6 + norm*3 + map, norm in {L1,L4}, map in {ssim, artifact, detail}. */
int ssimu2_get_norms(ssimu2_t *h, uint64_t ticket, double *norms108);
/* Submit + wait + score in one call (Ssimulacra2::compute_sync). */
int ssimu2_compute_sync(ssimu2_t *h, const ssimu2_frame *ref, const ssimu2_frame *dis, void *stream, double *score);
/* Make `stream` wait (device side) until the ticket's batch is done, so the caller may
 * recycle the input frames in stream order without a host sync. */
int ssimu2_stream_wait(ssimu2_t *h, uint64_t ticket, void *stream);

/* ---- scoring: host frames ------------------------------------------------------------ */
/* Ssimulacra2::compute_from_cpu_srgb_sync generalised to every format: the frames live in
 * HOST memory with the same layout an ssimu2_frame describes (plane[] are host addresses;
 * for YUV plane[1] must lie inside the same allocation as plane[0], after it).  The library
 * copies them to its own device staging ring (pinned memory is faster but not required) and
 * enqueues the pair.  frame_bytes = bytes to copy starting at plane[0]. */
int ssimu2_submit_host(ssimu2_t *h, const ssimu2_frame *ref, const ssimu2_frame *dis, size_t frame_bytes,
                       uint64_t *ticket);
/* n host pairs at once (all frames share frame_bytes); pair i has *first_ticket + i.  The frame loop of a caller that
 * decodes on the CPU (turbo-metrics/src/input_image.rs:206-228 copies each frame itself) becomes one call per chunk. */
int ssimu2_submit_host_batch(ssimu2_t *h, uint32_t n, const ssimu2_frame *refs, const ssimu2_frame *diss,
                             size_t frame_bytes, uint64_t *first_ticket);

/* ---- results on the device ----------------------------------------------------------- */
/* Device address of the f64 score ring (one entry per ticket, index ticket % capacity). */
int ssimu2_scores_device(ssimu2_t *h, uint64_t *dptr, uint64_t *capacity);

/* ---- introspection (used by the parity tests and bench.py) ---------------------------- */
typedef struct {
    uint32_t nscales;
    uint32_t width[6], height[6], pitch[6]; /* pitch in floats */
    uint32_t batch, ring;
    uint64_t alg_bytes_per_pair; /* B_alg of SURVEY.md section 8(d) for this geometry */
    uint64_t kernel_launches;    /* kernels launched so far by this handle */
} ssimu2_info;
int ssimu2_get_info(const ssimu2_t *h, ssimu2_info *info);
/* Copy an intermediate plane set of the batch slot that served `ticket` to host memory.
 * what = 0: XYB planes of `scale`:                float[2][3][h][w]  (ref X,Y,B then dis X,Y,B)
 * what = 1: H-pass output of `scale`:            float[15][h][w]    (s11,s22,s12,mu1,mu2) x 3 channels
 * Only valid until that slot is reused (i.e. right after ssimu2_get_score). */
int ssimu2_debug_read(ssimu2_t *h, uint64_t ticket, int what, int scale, float *out, size_t out_floats);
/* Run the device arithmetic helpers over an array (host pointers) so the tests can compare them bit for
 * bit with libm / IEEE division:  op 0: out[i] = cbrtf(in[i]);  op 1: powf(in[i], y);  op 2: in[i] / y (f32);
 * op 3: `in` holds n (num, den) pairs of DOUBLES, `out` n doubles: num / den;
 * op 4: `in` holds n (num, den) pairs of floats: the V-pass quotient;
 * op 5 / 6: the unchecked hot-path forms of op 0 / 1 (positive normal arguments only). */
int ssimu2_debug_math(int op, const float *in, float y, float *out, size_t n);
/* Average device time (ms) of the last completed batch per kernel: pyramid, hpass, vpass, finalize. */
int ssimu2_last_batch_ms(ssimu2_t *h, float ms[4]);
/* Device time (ms, CUDA events on the batch's own stream) summed per kernel over every batch
 * completed so far: frontend, hpass, vpass, finalize; optional counters; reset != 0 clears them. */
int ssimu2_kernel_ms(ssimu2_t *h, double ms_total[4], uint64_t *batches, uint64_t *pairs, int reset);

#ifdef __cplusplus
}
#endif
#endif /* SSIMU2_B200_H */
