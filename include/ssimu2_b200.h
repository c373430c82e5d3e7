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
 *   cudarse-video/src/dec.rs:277-287     FrameMapping drop (surface reuse)          -> ssimu2_stream_wait_input
 *   cuda-colorspace/src/lib.rs:33-123    ColorspaceConversion::biplanaryuv420_to_linearrgb_{8,16}
 *                                        (folded into the scorer: SSIMU2_FMT_NV12 / SSIMU2_FMT_P016)
 *   cuda-colorspace/src/lib.rs:144-169   srgb_to_linear_{u8,u16,f32}                (SSIMU2_FMT_SRGB8/16/F32)
 *   turbo-metrics/src/lib.rs:268-360     TurboMetrics::compute_one (the caller)     -> ssimu2_submit / ssimu2_get_score
 *   turbo-metrics/src/lib.rs:362-433     TurboMetrics::compute_all frame loop       -> ssimu2_submit_batch + tickets;
 *                                        across the GPUs of a box                   -> ssimu2_shard_*
 *
 * Conventions
 *   - Plain C types only.  Device pointers are CUdeviceptr-compatible 64-bit integers,
 *     streams are CUstream / cudaStream_t handles passed as void*.
 *   - Every function returns 0 (SSIMU2_OK) or a negative ssimu2_status; a positive value is
 *     a CUDA runtime error code passed through.  Nothing throws or aborts across the ABI.
 *   - Input frames are BORROWED: read-only, never freed.  LIFETIME RULE: the scorer reads a frame when its batch (or,
 *     with cfg.input_group, its input group) is LAUNCHED, which can be later than ssimu2_submit returns.  A frame must
 *     therefore stay valid AND UNMODIFIED until one of: ssimu2_wait / ssimu2_get_score / ssimu2_get_scores has
 *     returned for its ticket; ssimu2_completed reports a watermark above its ticket; or the stream that will
 *     overwrite it has been made to wait with ssimu2_stream_wait_input (inputs consumed) or ssimu2_stream_wait (batch
 *     done).  The reference enqueues its graph on the caller's stream at once and syncs per pair
 *     (turbo-metrics/src/lib.rs:342-358); here the dependency on the caller's stream is recorded at submit time, so
 *     everything enqueued on `stream` BEFORE the submit is seen, and nothing enqueued after it is waited for.
 *   - Decoder-style surface pools (cudarse-video/src/dec.rs:277-287: a surface is valid until unmapped): set
 *     cfg.input_group (pairs per front-end launch, e.g. 1-4) so that surfaces are consumed shortly after submit, and
 *     recycle each one behind ssimu2_stream_wait_input(ticket, decoder_stream) -- no host synchronisation.
 *   - Scales: like the reference's CPU implementation (cpu.rs:359) a frame whose scale-s size is below 8x8 stops the
 *     pyramid at s+1 scales (fewer than 6 for min(width, height) < 113) and the 108 weights are consumed densely over
 *     the scales that exist.  The reference's GPU op always runs 6 scales with fixed weight offsets
 *     (ssimulacra2-cuda/src/lib.rs:61-65, 586-603) and gives a different score for such small frames; the parity
 *     target named by BASELINE.json is the CPU implementation.  ssimu2_info.nscales tells which case applies.
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
 * both with the same pitch.  Packed RGB formats use plane[0] only.  pitch is in bytes and must hold one row
 * (width x 1 / 2 bytes for NV12 / P016, width x 3 / 6 / 12 bytes for the packed formats): SSIMU2_E_INVALID otherwise. */
typedef struct {
    uint64_t plane[2];
    uint32_t pitch;
    uint32_t reserved;
} ssimu2_frame;

typedef enum {
    SSIMU2_PIPELINE_DEFAULT = 0, /* front-end, fused H+V filter/map/sum kernel, finalize */
    SSIMU2_PIPELINE_SPLIT = 1    /* development: separate H and V passes, H-pass planes kept in HBM (ssimu2_debug_read) */
} ssimu2_pipeline;

/* ssimu2_config.flags */
#define SSIMU2_FLAG_SCORE_ONLY 1u /* compute only what the score depends on: the filters and SSIM map of every (scale, channel) \
                                     whose two SSIM weights are both zero are skipped (the reference computes them and multiplies \
                                     by 0.0, ssimulacra2-cuda/src/lib.rs:586-603).  Scores are bit-identical to the full mode;    \
                                     ssimu2_get_norms returns SSIMU2_E_UNSUPPORTED. */
#define SSIMU2_FLAG_NO_TIMING 2u  /* do not record the per-kernel timing events (ssimu2_b200_debug.h) */
#define SSIMU2_FLAG_P016_DEEP 4u  /* SSIMU2_FMT_P016 only: the samples carry more than 10 significant bits (12-bit sources, e.g. HEVC \
                                     Main12 through NVDEC).  Results never depend on this flag; it selects a front-end that      \
                                     evaluates all three transfer functions instead of the exact 10-bit memo tables, which such  \
                                     content cannot use (without the flag it is scored correctly, at a third of the speed). */

typedef struct {
    uint32_t width, height;
    int32_t format;       /* ssimu2_format */
    int32_t matrix;       /* ssimu2_matrix */
    int32_t full_range;   /* 0 = limited (what the reference implements), 1 = full */
    int32_t device;       /* CUDA device ordinal */
    uint32_t batch;       /* frame pairs per kernel launch group, at most 1024.  0 = chosen from the frame size: 8 at 4K, 32 at
                             1080p, 128 for 512x512 (enough strips for ~8 waves, workspace of all ring slots <= 8 GiB) */
    uint32_t ring;        /* batches in flight, one stream each (0 = default 3) */
    uint32_t pipeline;    /* ssimu2_pipeline */
    uint32_t flags;       /* SSIMU2_FLAG_* */
    uint32_t input_group; /* pairs per front-end launch; 0 = whole batch.  Smaller groups consume the input frames sooner */
    uint32_t reserved[5]; /* must be 0 */
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
/* *watermark = the lowest ticket whose results have not reached the host yet: every pair below it has completed and its
 * input frames are no longer referenced (batches complete out of order across ring slots; this is the in-order bound). */
int ssimu2_completed(ssimu2_t *h, uint64_t *watermark);
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
/* Same, but only until the ticket's INPUT FRAMES have been consumed (its front-end launch is done); launches the pending
 * input group if the ticket is still waiting in one.  This is the call that lets a decoder reuse a surface. */
int ssimu2_stream_wait_input(ssimu2_t *h, uint64_t ticket, void *stream);

/* ---- scoring: host frames ------------------------------------------------------------ */
/* Ssimulacra2::compute_from_cpu_srgb_sync generalised to every format: the frames live in
 * HOST memory with the same layout an ssimu2_frame describes (plane[] are host addresses;
 * for YUV plane[1] must lie inside the same allocation as plane[0], after it).  The library
 * copies them to its own device staging ring (pinned memory is faster but not required) and
 * enqueues the pair.  frame_bytes = bytes to copy starting at plane[0].
 * Lifetime: a PAGEABLE buffer has been read when the call returns (the CUDA runtime stages it) and may be reused at once,
 * like the slices of the reference's call; a PINNED buffer is read asynchronously by the copy engine: keep it unmodified
 * until the pair's score has been fetched or ssimu2_completed() has passed its ticket. */
int ssimu2_submit_host(ssimu2_t *h, const ssimu2_frame *ref, const ssimu2_frame *dis, size_t frame_bytes,
                       uint64_t *ticket);
/* n host pairs at once (all frames share frame_bytes); pair i has *first_ticket + i.  The frame loop of a caller that
 * decodes on the CPU (turbo-metrics/src/input_image.rs:206-228 copies each frame itself) becomes one call per chunk. */
int ssimu2_submit_host_batch(ssimu2_t *h, uint32_t n, const ssimu2_frame *refs, const ssimu2_frame *diss,
                             size_t frame_bytes, uint64_t *first_ticket);

/* ---- results on the device ----------------------------------------------------------- */
/* Device address of the f64 score ring (one entry per ticket, index ticket % capacity). */
int ssimu2_scores_device(ssimu2_t *h, uint64_t *dptr, uint64_t *capacity);

/* ---- introspection ----------------------------------------------------------------- */
typedef struct {
    uint32_t nscales;
    uint32_t width[6], height[6], pitch[6]; /* pitch in floats */
    uint32_t batch, ring;
    uint64_t alg_bytes_per_pair; /* B_alg of SURVEY.md section 8(d) for this geometry */
    uint64_t kernel_launches;    /* kernels launched so far by this handle */
    uint32_t pipeline, flags, input_group;
    uint32_t strips_per_pair;    /* k_hv work items (64-column strips over all scales) per pair */
    uint64_t io_bytes_per_pair;  /* B_io: compulsory input bytes of a pair + 8 */
} ssimu2_info;
int ssimu2_get_info(const ssimu2_t *h, ssimu2_info *info);

/* ---- frame-sharded scoring across the GPUs of one box (ssimu2_shard.cu) -------------------------------------------
 * The reference drives one GPU (device 0 is hard-coded, turbo-metrics/src/lib.rs:438-456) from one thread
 * (frame loop turbo-metrics/src/lib.rs:362-433).  SSIMULACRA2 pairs are independent (ssimulacra2-cuda/README.md:26-27),
 * so N GPUs = N handles, each driven by its own host thread INSIDE the library; the caller keeps the single-threaded
 * submit / fetch loop.  Global tickets increase by one per pair in submission order; pair g goes to device
 * (g / cfg.batch) % n_devices (whole batches stay on one GPU).  No collective, no peer access: only f64 scores leave a GPU.
 * Frames are HOST frames (any device would need them anyway): each worker copies its pairs to its own GPU. */
typedef struct ssimu2_shard ssimu2_shard_t;
/* cfg->device is ignored; devices[i] are CUDA ordinals (n_devices >= 1). */
int ssimu2_shard_create(ssimu2_shard_t **out, const ssimu2_config *cfg, const int32_t *devices, uint32_t n_devices);
int ssimu2_shard_destroy(ssimu2_shard_t *s);
/* n host pairs (layout as ssimu2_submit_host); returns at once, the copies and launches run on the worker threads.
 * The frames must stay valid until their scores have been fetched.  *first_ticket = global ticket of pair 0. */
int ssimu2_shard_submit_host(ssimu2_shard_t *s, uint32_t n, const ssimu2_frame *refs, const ssimu2_frame *diss,
                             size_t frame_bytes, uint64_t *first_ticket);
/* n DEVICE pairs that already live on the GPU the sharding rule assigns them to (ssimu2_shard_device_of); `streams[d]`
 * (may be NULL = no dependency) is the stream of device d the frames were produced on. */
int ssimu2_shard_submit_device(ssimu2_shard_t *s, uint32_t n, const ssimu2_frame *refs, const ssimu2_frame *diss,
                               void *const *streams, uint64_t *first_ticket);
/* CUDA ordinal that global ticket `ticket` is (or will be) scored on. */
int ssimu2_shard_device_of(const ssimu2_shard_t *s, uint64_t ticket, int32_t *device);
/* Scores of n consecutive global tickets, in submission order (blocks until they are done). */
int ssimu2_shard_get_scores(ssimu2_shard_t *s, uint64_t first_ticket, uint32_t n, double *scores);
int ssimu2_shard_flush(ssimu2_shard_t *s);

#ifdef __cplusplus
}
#endif
#endif /* SSIMU2_B200_H */
