/* oss_b200.h — C ABI of liboss_b200.so, the B200-native engine behind
 * OpenSkyStacker's ImageStacker for ONE path: libstacker's per-frame pipeline
 * (calibrate -> detect stars -> FOCAS triangle align -> warp -> mean stack).
 *
 * The reference has no FFI/plugin layer: its "operator API" is the C++ class
 * openskystacker::ImageStacker (libstacker/include/libstacker/imagestacker.h:20-114)
 * plus the free functions of libstacker/include/libstacker/util.h:57-76 and
 * StarDetector::getStars (libstacker/include/libstacker/stardetector.h:28).  This
 * header is what a libstacker build would bind INSTEAD of those bodies; each entry
 * point cites the reference code it replaces.  INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - plain C types only; every function returns an oss_status (0 = OK, < 0 = error)
 *     unless stated; oss_last_error(ctx) gives the text.  Nothing throws.
 *   - frames are interleaved HWC in the order the decoder produced them (OpenCV: BGR),
 *     width*height*channels samples, no row padding.
 *   - `mem` says where a frame pointer lives: OSS_MEM_HOST (pageable or pinned; the
 *     library stages it through pinned buffers and overlaps the copy with compute)
 *     or OSS_MEM_DEVICE (already in this context's GPU memory; used in place).
 *   - one oss_ctx drives one GPU; calls on one ctx are not re-entrant.
 *   - there is NO CPU fallback: without a CUDA device every call fails with
 *     OSS_E_CUDA.
 *
 * All citations are relative to the reference tree (BenJuan26/OpenSkyStacker v0.3.0).
 */
#ifndef OSS_B200_H
#define OSS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define OSS_API __attribute__((visibility("default")))
#else
#define OSS_API
#endif

typedef struct oss_ctx oss_ctx;

typedef enum {
    OSS_OK = 0,
    OSS_E_CUDA = -1,       /* CUDA / NCCL failure, or no device: never falls back to the CPU */
    OSS_E_ARG = -2,        /* bad argument / call order */
    OSS_E_TYPE = -3,       /* "Images must be the same type."             imagestacker.cpp:236-246 */
    OSS_E_SIZE = -4,       /* "Images must all be the same size."         imagestacker.cpp:254-258 */
    OSS_E_THREADS = -5,    /* "Number of threads must be at least 1."     imagestacker.cpp:289-292 */
    OSS_E_NO_ALIGNED = -6, /* "No images could be aligned to the reference image. Try using a lower tolerance."
                              imagestacker.cpp:355-359 */
    OSS_E_CAPACITY = -7,   /* a frame produced more candidate pixels / stars than the workspace holds */
    OSS_E_NOMEM = -8
} oss_status;

typedef enum { OSS_U8 = 0, OSS_U16 = 1, OSS_F32 = 2 } oss_sample; /* util.cpp:441-456 (8U, 16U, 32F branches) */
typedef enum { OSS_MEM_HOST = 0, OSS_MEM_DEVICE = 1 } oss_mem;
/* ImageRecord::FrameType minus LIGHT, model.h:16-22 */
typedef enum { OSS_MASTER_BIAS = 0, OSS_MASTER_DARK = 1, OSS_MASTER_DARKFLAT = 2, OSS_MASTER_FLAT = 3 } oss_master;

/* openskystacker::Star, model.h:34-62 (radius is never written by the reference and is omitted) */
typedef struct {
    int32_t x, y, area;
    float peak, value;
} oss_star;

/* What StarDetectorImpl::getStars computes on the way, stardetector.cpp:121-129 */
typedef struct {
    double mean, sigma;
    float threshold, minpeak;
    int32_t ncandidates; /* pixels > threshold */
    int32_t ncomponents; /* connected components with peak > minPeak */
    int32_t nstars;
    int32_t overflow;    /* != 0: workspace capacity exceeded, result invalid */
} oss_detect_info;

/* Per-light outcome of generateAlignedImage, util.cpp:557-582 */
typedef struct {
    int32_t nstars;  /* stars detected in the light */
    int32_t k;       /* matched pairs handed to findTransform (focas.cpp:86) */
    int32_t err;     /* 0 = aligned and stacked; -1 = skipped like util.cpp:746-747 */
    float M[6];      /* row-major 2x3, maps reference -> light (what WARP_INVERSE_MAP consumes) */
    float threshold; /* detection threshold of this light */
    int32_t overflow;
} oss_frame_report;

/* ---- lifetime ------------------------------------------------------------------------ */
OSS_API int oss_create(oss_ctx **ctx, int device);
OSS_API void oss_destroy(oss_ctx *ctx);
OSS_API const char *oss_last_error(const oss_ctx *ctx);
OSS_API const char *oss_status_string(int status); /* the reference's processingError texts */
OSS_API const char *oss_version(void);
OSS_API int oss_device_count(void); /* CUDA devices visible to this process (0: none, every other call fails with OSS_E_CUDA) */

/* Pinned host memory for decoders to write into (optional; any host pointer works). */
OSS_API void *oss_host_alloc(size_t bytes);
OSS_API void oss_host_free(void *p);
/* Device memory owned by the ctx, for callers that stage frames themselves (bench, tests). */
OSS_API void *oss_device_alloc(oss_ctx *ctx, size_t bytes);
OSS_API void oss_device_free(oss_ctx *ctx, void *p);
OSS_API int oss_memcpy_h2d(oss_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes);
OSS_API int oss_memcpy_d2h(oss_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes);

/* ---- geometry: replaces validateImageSizes' contract, imagestacker.cpp:402-557 -------- */
/* All frames of one stack share (width, height, channels, sample).  channels is 1 or 3.
 * frames_in_flight (>= 1) sizes the batch the engine processes per kernel launch. */
OSS_API int oss_set_geometry(oss_ctx *ctx, int width, int height, int channels, int sample,
                             int frames_in_flight);
/* Number of independent pipelines (streams + work spaces) that batches of lights alternate
 * between: the latency-bound detection tail of one batch overlaps the bulk kernels of the next
 * n - 1 batches; 1..4, default 3.  The accumulation order stays the frame order.  Takes effect at
 * the next oss_set_geometry. */
OSS_API int oss_set_pipelines(oss_ctx *ctx, int n);
/* SM partition: the FP32-pipe-bound sky kernel (k_sky_sub_stats) gets a green context of `fp_sms` SMs (rounded up to the
 * device's granularity, 8 on sm_100), the HBM-bound bulk kernels (masters, calibration, warp + accumulate) the remaining
 * SMs, so the two classes run side by side on disjoint SMs instead of back to back; results are bit-identical (only the
 * order in time changes, never the accumulation order).  0 = no partition.  Takes effect at the next oss_set_geometry;
 * oss_get_sm_split returns what the driver granted (0, 0 when off).  Environment OSS_B200_SPLIT sets the default. */
OSS_API int oss_set_sm_split(oss_ctx *ctx, int fp_sms);
OSS_API int oss_get_sm_split(oss_ctx *ctx, int *fp_sms, int *hbm_sms);
/* Detection work space: room for width*height / one_in candidate pixels (pixels above the detection threshold) per
 * frame; 1..64, default 16 (6.25 %).  The reference has no such limit (stardetector.cpp:192-246 grows vectors); here a
 * frame that exceeds it is NOT silently dropped: its report carries overflow != 0 and oss_finish / oss_sync return
 * OSS_E_CAPACITY, upon which the caller raises the capacity and runs the stack again.  Threshold 1 (1.5 sigma: 6.7 % of
 * pure-noise pixels) needs one_in <= 8; thresholds >= 2 fit the default on any star field.  Takes effect at the next
 * oss_set_geometry. */
OSS_API int oss_set_candidate_capacity(oss_ctx *ctx, int one_in);

/* ---- master frames: stackBias / stackDarks / stackDarkFlats / stackFlats, util.cpp:584-722
 * Build order follows ImageStackerPrivate::process (imagestacker.cpp:262-281): bias first,
 * then dark and dark-flat (each minus bias when present), then flat (minus bias, minus
 * dark-flat, divided by the mean of its channel 0).  A master that was never built is
 * "absent" exactly like an empty cv::Mat (util.cpp:269-271). */
OSS_API int oss_build_master(oss_ctx *ctx, int kind, const void *const *frames, int n, int mem);
/* The same build in pieces, for callers that decode calibration frames a few at a time (stackDarks' loop, util.cpp:596-607,
 * reads one file per iteration) and for multi-GPU builds: oss_master_begin names the master, the total number of frames
 * and the rows [row0, row0 + rows) of it that THIS context accumulates (rows <= 0: the whole frame; a slice starts on a
 * multiple of 16 rows); oss_master_add hands in the next n FULL frames, in order (only the slice's bytes are copied /
 * read; host frames may be reused when it returns); oss_master_end closes it (a whole flat is normalised here, a sliced
 * one in oss_comm_allgather_master).  oss_build_master = begin(whole) + add + end. */
OSS_API int oss_master_begin(oss_ctx *ctx, int kind, int n_total, int row0, int rows);
OSS_API int oss_master_add(oss_ctx *ctx, const void *const *frames, int n, int mem);
OSS_API int oss_master_end(oss_ctx *ctx);
OSS_API int oss_set_master(oss_ctx *ctx, int kind, const float *data, int mem);
OSS_API int oss_get_master(oss_ctx *ctx, int kind, float *out_host); /* OSS_E_ARG if absent */
OSS_API int oss_clear_masters(oss_ctx *ctx);

/* ---- stacking: ImageStackerPrivate::process + processConcurrent ------------------------
 * oss_set_reference: getCalibratedImage(ref) (util.cpp:266-274), workingImage = ref.clone()
 *   and totalValidImages = 1 (imagestacker.cpp:286-287,304); the reference stars are
 *   detected ONCE here and cached (the reference re-detects them for every light,
 *   util.cpp:559 — same values, no need to repeat).  Asynchronous like oss_add_lights: it
 *   returns once the work is queued (a host frame has been copied by then and may be reused);
 *   the detection runs while the first lights are being calibrated.  Its results and a
 *   workspace overflow (OSS_E_CAPACITY) surface at oss_get_reference_stars / oss_finish.
 * oss_add_lights: the body of processConcurrent's loop (util.cpp:738-751) for n lights:
 *   calibrate, getStars, triangles, vote, HFTI fit, warpAffine, accumulate; lights whose
 *   alignment fails are skipped silently.  Asynchronous: it returns once the copies and
 *   kernels are queued; OSS_MEM_HOST frames must stay valid and unmodified until
 *   oss_sync / oss_finish / oss_get_frame_report returns (pinned buffers from
 *   oss_host_alloc are copied by DMA while earlier batches compute).
 * oss_finish: workingImage /= totalValidImages (imagestacker.cpp:361), copies the float32
 *   HWC result to out_host (may be NULL to leave it on the device) and returns the count.
 *   OSS_E_NO_ALIGNED when totalValidImages < 2 (imagestacker.cpp:355-359). */
OSS_API int oss_set_reference(oss_ctx *ctx, const void *frame, int mem, int threshold);
OSS_API int oss_add_lights(oss_ctx *ctx, const void *const *frames, int n, int mem);
OSS_API int oss_sync(oss_ctx *ctx);
/* Host pipelining: returns when the OSS_MEM_HOST frames of all oss_add_lights calls except the `keep_recent` most recent
 * ones have been copied to the device, i.e. their buffers may be refilled (the worker threads of processConcurrent,
 * util.cpp:738-751, decode the next frames meanwhile) — without waiting for the GPU work on them like oss_sync does. */
OSS_API int oss_wait_host_frames(oss_ctx *ctx, int keep_recent);
OSS_API int oss_finish(oss_ctx *ctx, float *out_host, int *total_valid);
OSS_API int oss_reset_stack(oss_ctx *ctx); /* forget reference + lights, keep masters */

OSS_API int oss_num_reports(oss_ctx *ctx);
OSS_API int oss_get_frame_report(oss_ctx *ctx, int i, oss_frame_report *out);
OSS_API int oss_get_reference_stars(oss_ctx *ctx, oss_star *out, int cap, int *n, oss_detect_info *info);
/* sum image so far (before the divide), float32 HWC, and the running valid count */
OSS_API int oss_get_sum(oss_ctx *ctx, float *out_host, int *total_valid);
OSS_API const float *oss_result_device_ptr(oss_ctx *ctx);

/* ---- detect-only: ImageStackerPrivate::detectStars (imagestacker.cpp:392-400) and
 * StarDetector::getStars (stardetector.cpp:108-147).  No calibration is applied, like the
 * reference's -s mode.  out may be NULL (count only). */
OSS_API int oss_detect_stars(oss_ctx *ctx, const void *frame, int mem, int threshold, oss_star *out,
                             int cap, int *n, oss_detect_info *info);
/* Same detector on n frames at once (config 5: batched -s sweeps); counts[i] receives the
 * star count of frame i. */
OSS_API int oss_detect_stars_batch(oss_ctx *ctx, const void *const *frames, int n, int mem,
                                   int threshold, int32_t *counts);

/* Threshold sweep of the -s mode on one frame (config: "star-detect-only threshold sweep 1-100"): the sky
 * background, subtraction and sigma are computed once, thresholding + components + deblending per
 * threshold.  counts[i] = number of stars at threshold t_first + i (-1 = workspace overflow). */
OSS_API int oss_detect_sweep(oss_ctx *ctx, const void *frame, int mem, int t_first, int t_last, int32_t *counts);

/* ---- alignment only: generateTriangleList + findMatches + findTransform (focas.cpp) -----
 * xy are interleaved int32 (x0,y0,x1,y1,...) star lists in detection order; only the first
 * 40 of each are used (focas.cpp:7-8).  matches_out (may be NULL) receives 3 rows of 120
 * ints like the reference's `matches`; votes_out (may be NULL) the 40x40 vote table. */
OSS_API int oss_align_lists(oss_ctx *ctx, const int32_t *xy_ref, int n_ref, const int32_t *xy_tgt, int n_tgt,
                            int32_t *k, int32_t *err, float *M, int32_t *matches_out, int32_t *votes_out);

/* ---- warp only: cv::warpAffine(INTER_LINEAR | WARP_INVERSE_MAP), util.cpp:579 ----------- */
OSS_API int oss_warp_affine(oss_ctx *ctx, const float *src_host, float *dst_host, const float *M);
/* The stacking path's warp + accumulate in isolation (util.cpp:579 + :749 for n frames): dst = sum over i, in order, of
 * warpAffine(src[i], M + 6 i) for the frames whose err[i] == 0 (err may be NULL; util.cpp:746-747 skips the others),
 * n <= frames_in_flight.  `which` selects the kernel — OSS_WARP_AUTO is what oss_add_lights runs for this geometry,
 * OSS_WARP_TMA / OSS_WARP_TILED force the TMA-staged / cp.async-staged variant (OSS_E_ARG when the geometry cannot use
 * it) — and `bands` >= 1 splits the launch into that many bands of tile rows like oss_comm_reduce does. */
typedef enum { OSS_WARP_AUTO = 0, OSS_WARP_TMA = 1, OSS_WARP_TILED = 2 } oss_warp_kernel;
OSS_API int oss_warp_accumulate(oss_ctx *ctx, const float *const *src_host, const float *M, const int32_t *err, int n,
                                float *dst_host, int which, int bands);

/* ---- multi-GPU: frames shard across ranks (the reference's -j decomposition, util.cpp:738;
 * combine = imagestacker.cpp:348-351).  One process per GPU; the caller distributes the
 * 128-byte id from rank 0 to all ranks (e.g. over torch.distributed / MPI / a file). */
OSS_API int oss_comm_unique_id(void *id128);
OSS_API int oss_comm_init(oss_ctx *ctx, const void *id128, int rank, int nranks);
OSS_API int oss_comm_broadcast_masters(oss_ctx *ctx, int root);
/* one master (oss_master kind), asynchronously on a stream of its own: on the root it overlaps the building of the next
 * master and the reference frame; calibration waits for it on the device.  Every rank calls it for the same kinds in the
 * same order (the image list tells each rank which masters exist). */
OSS_API int oss_comm_broadcast_master(oss_ctx *ctx, int kind, int root);
/* Masters built cooperatively: rank r accumulates the rows oss_comm_master_slice returns (1/N of the frame) of EVERY
 * calibration frame (oss_master_begin/add/end), then oss_comm_allgather_master exchanges the slices in place over NVLink on
 * the comm stream (asynchronously: it overlaps the next master's slice and the reference; calibration waits for it on the
 * device) and normalises a flat.  Per-pixel arithmetic is unchanged, so the masters are bit-identical to a 1-GPU build;
 * each rank reads (and, from the host, uploads) 1/N of the calibration bytes. */
OSS_API int oss_comm_master_slice(oss_ctx *ctx, int *row0, int *rows);
OSS_API int oss_comm_allgather_master(oss_ctx *ctx, int kind);
/* ships the calibrated reference's triangle and star lists; `threshold` is the detection threshold of the stack (the same
 * on every rank: it is a command-line value); only `root` keeps the reference frame inside its sum
 * (imagestacker.cpp:287,304).  No rank waits on the host: alignment kernels wait for the data on the device. */
OSS_API int oss_comm_broadcast_reference(oss_ctx *ctx, int root, int threshold);
/* ncclReduce(sum) + count to root over NVLink, queued behind this rank's last warp + accumulate (asynchronous); afterwards
 * oss_finish on root gives the stack, oss_sync on the others drains their part */
OSS_API int oss_comm_reduce(oss_ctx *ctx, int root);

/* ---- bench utilities (no reference counterpart) -------------------------------------------
 * oss_mark(0) / oss_mark(1) record CUDA events on the engine's compute stream; oss_elapsed_ms
 * waits for mark 1 and returns the device time between them.
 * oss_synth_*: seeded synthetic u16 frames rendered directly in device memory (star field
 * through a per-frame motion, or dark/bias/flat-like calibration frames), so that device-
 * resident benchmarks need no host rendering.  stars_xyas = nstars * (x, y, amplitude, sigma). */
OSS_API int oss_mark(oss_ctx *ctx, int which);
OSS_API int oss_elapsed_ms(oss_ctx *ctx, double *ms);
OSS_API int oss_synth_light(oss_ctx *ctx, void *dst_dev, const float *stars_xyas, int nstars, float background,
                            float noise, float pedestal_adu, int vignette, unsigned seed, const float *gains3);
OSS_API int oss_synth_calib(oss_ctx *ctx, void *dst_dev, float level_adu, float sigma_adu, int flat_profile,
                            float pedestal_adu, unsigned seed);

/* ---- instrumentation ------------------------------------------------------------------- */
/* CUDA-event timing of every kernel launch, on the stream it is launched on. */
OSS_API int oss_profile_enable(oss_ctx *ctx, int on);
OSS_API int oss_profile_reset(oss_ctx *ctx);
OSS_API int oss_profile_count(oss_ctx *ctx);
OSS_API int oss_profile_get(oss_ctx *ctx, int i, const char **name, double *total_ms, int *launches);
/* kernels launched by this ctx since creation (the bench's gpu_launches claim) */
OSS_API long long oss_launch_count(oss_ctx *ctx);
/* test hook: copy an intermediate plane of batch slot `slot` to the host.
 * what: 0 calibrated (HWC f32), 1 gray (HW f32), 2 sky-subtracted "stars" (HW f32),
 *       3 low-res resize (f32), 4 low-res median (f32) */
OSS_API int oss_debug_copy(oss_ctx *ctx, int what, int slot, float *out_host);

#ifdef __cplusplus
}
#endif
#endif /* OSS_B200_H */
