// liboss_b200.so — host side of the C ABI declared in include/oss_b200.h.
// Orchestrates the sm_100a kernels of k_*.cuh on one GPU: stream-ordered, batched over
// frames, no host round trip inside a batch.  There is no CPU fallback anywhere in this
// file: without a CUDA device every entry point fails with OSS_E_CUDA.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "k_calibrate.cuh"
#include "k_detect.cuh"
#include "k_focas.cuh"
#include "k_sky.cuh"
#include "k_synth.cuh"
#include "k_warp.cuh"

using namespace oss;

// ----------------------------------------------------------------------------------------
struct ProfEntry { std::string name; double ms; int launches; };
struct ProfPending { int entry; cudaEvent_t a, b; int pipe; };

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;

static bool nccl_load(std::string *why)
{
    if (g_nccl.handle) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) { *why = std::string("cannot load NCCL: ") + dlerror(); return false; }
#define SYM(field, name) *(void **)(&g_nccl.field) = dlsym(h, name); if (!g_nccl.field) { *why = "NCCL symbol missing: " name; return false; }
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(Broadcast, "ncclBroadcast")
    SYM(Reduce, "ncclReduce")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.handle = h;
    return true;
}

// Scheduling.  Kernels fall in two classes: BULK kernels that fill the GPU (masters, calibration, sky, warp) and the
// latency-bound detection TAIL of a batch (thresholding, scans, union-find, component walk, alignment: ~35 dependent
// launches of a few CTAs each).  All bulk kernels go to ONE low-priority stream in a fixed order; the tail of each batch
// runs on a high-priority stream of its own pipeline (= work space) and overlaps the bulk kernels of the following
// batches.  The warp + accumulate of a batch is queued on the bulk stream only after the bulk kernels of the next
// npipes - 1 batches, by which time its tail has finished, so the bulk stream never waits:
//     cal(0) cal(1) sky(0) cal(2) sky(1) warp(0) cal(3) sky(2) warp(1) ...   (bulk stream, 3 pipelines)
//            lowres(0)     lowres(1) tail(0)    lowres(2) tail(1) ...        (chain streams)
// (calibration runs one batch ahead; lowres = the resize / median kernels, 1/64 of the pixels)
// Accumulation order = batch order = frame order.  Co-running two bulk kernels from different streams was measured
// slower than running them back to back (they evict each other's CTAs from the SMs), descending stream priorities
// per pipeline likewise; the block scheduler serves equal-priority streams in launch order anyway.
#define OSS_MAX_PIPES 4
struct Pipe {
    cudaStream_t copy_stream = nullptr;  // H2D staging copies of this pipeline
    cudaStream_t chain_stream = nullptr; // high priority: the latency-bound detection tail + alignment of a batch
    cudaStream_t side_stream = nullptr;  // high priority: the large-component kernels of a batch, beside the small-component one
    Workspace ws{};
    int *d_xy = nullptr, *d_nxy = nullptr;
    unsigned char *raw_stage[2] = {nullptr, nullptr};
    cudaEvent_t copy_done[2] = {nullptr, nullptr}, raw_free[2] = {nullptr, nullptr};
    cudaEvent_t warp_done = nullptr, detect_done = nullptr, sky_done = nullptr, cal_done = nullptr, med_done = nullptr;
    cudaEvent_t side_fork = nullptr, side_join = nullptr;
    CUtensorMap cal_map;       // [slot][row][column * channel] view of ws.cal for the TMA warp kernel
    bool has_cal_map = false;
    long long batch_counter = 0;
};

struct oss_ctx {
    int device = 0;
    // the fields below alias the CURRENT pipeline (select_pipe); all launch helpers use them
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    Pipe pipes[OSS_MAX_PIPES];
    int npipes = 3, want_pipes = 3, cur = 0;
    // host-frame tickets: one event per oss_add_lights call with OSS_MEM_HOST frames, completed when the call's frames have
    // been copied to the device (oss_wait_host_frames)
    cudaStream_t ticket_stream = nullptr;
    cudaEvent_t tickets[16] = {};
    long long ticket_count = 0;
    struct MasterBuild { int kind = -1, n_total = 0, added = 0, row0 = 0, rows = 0; bool sliced = false; } mb;
    int slice_row0[4] = {0, 0, 0, 0}, slice_rows[4] = {0, 0, 0, 0}; // rows of each master this rank built last (0 rows: whole)
    int cap_one_in = 16;                 // detection work space: P / cap_one_in candidate pixels per frame
    bool cal_ring = true;                // batch calibration by k_calibrate_gray_ring (bulk copies); OSS_B200_CAL_RING=0: k_calibrate_gray_batch
    unsigned calr_attr = 0;              // k_calibrate_gray_ring instantiations whose attributes are set on this device
    cudaStream_t bulk = nullptr;         // the one low-priority stream of all bulk kernels
    // SM partition (green contexts): the FP32-bound sky kernel runs on `fp_stream`, which lives in a green context of
    // split_fp SMs, while the HBM-bound bulk kernels (masters, calibration, warp) run on `bulk`, re-created in the green
    // context of the remaining SMs — the two classes co-run on disjoint SMs at their own full residency.  0 = no split.
    int split_fp = 0, split_fp_got = 0, split_hbm_got = 0;
    cudaStream_t fp_stream = nullptr;
    CUgreenCtx green_fp = nullptr, green_hbm = nullptr;
    struct PendingWarp { int pipe, nb; };
    std::vector<PendingWarp> pending_warps; // detected batches whose warp + accumulate is not queued yet, oldest first
    int sky_blocks = 0; // CTAs per frame of the last sky_sub_stats launch (= partial sums finalize_stats adds up)
    long long light_batches = 0;
    std::string err;
    Geom g{};
    bool has_geom = false;
    int slots = 0;
    size_t frame_bytes = 0, stage_stride = 0; // stage_stride: frame_bytes rounded up to 256
    std::vector<void *> allocs; // geometry-scoped device allocations
    std::vector<void *> user_allocs;
    float *master[4] = {nullptr, nullptr, nullptr, nullptr};
    float *sum = nullptr, *ref_cal = nullptr, *result = nullptr;
    bool has_ref = false, ref_in_sum = true;
    bool sum_unset = false;              // a new stack whose sum plane has not been written yet: the first warp launch starts from 0
    bool slot0_is_ref = false;           // oss_debug_copy: the calibrated plane of "slot 0" is the reference's (it lives in ref_cal)
    bool dirty = true;                   // something was queued on a chain / side stream (or a warp deferred) since the last sync_all
    int threshold = 20;
    int *d_valid = nullptr;
    int extra_valid = 0; // lights counted on other ranks (after oss_comm_reduce)
    Workspace ws{};
    ResizeTap *dx = nullptr, *dy = nullptr, *ux = nullptr, *uy = nullptr;
    RefTriangles *d_ref = nullptr, *d_ref_tmp = nullptr;
    int *d_xy = nullptr, *d_nxy = nullptr; // [slots + 1] star lists for alignment (last = reference)
    int *d_dbg_matches = nullptr, *d_dbg_table = nullptr;
    long long *d_dbg_clk = nullptr;
    float *d_M = nullptr, *d_flat_ratio = nullptr;
    double *d_flat_partial = nullptr;
    unsigned char *raw_stage[2] = {nullptr, nullptr};
    cudaEvent_t copy_done[2] = {nullptr, nullptr}, raw_free[2] = {nullptr, nullptr};
    long long batch_counter = 0;
    // host-side mirrors (pinned), one entry per light handed in so far
    std::vector<DetectState *> h_ds_chunks;
    std::vector<AlignOut *> h_al_chunks;
    int nlights = 0;
    std::vector<oss_star> ref_stars;
    oss_detect_info ref_info{};
    // oss_set_reference is asynchronous: its results are kept on the device / in pinned memory and pulled on demand
    oss_star *d_ref_stars = nullptr;   // [cap] copy of the reference's star list (the work space slot is reused by lights)
    DetectState *h_ref_ds = nullptr;   // pinned
    cudaEvent_t ref_ready = nullptr;   // reference triangles built (alignment kernels of every pipeline wait for it)
    cudaEvent_t ref_copied = nullptr;  // d_ref_stars / h_ref_ds written
    bool ref_host_valid = true;
    // instrumentation
    bool profiling = false;
    std::vector<ProfEntry> prof;
    std::vector<ProfPending> prof_pending;
    std::vector<cudaEvent_t> event_pool;
    long long launches = 0;
    cudaEvent_t mark[2] = {nullptr, nullptr};
    float *synth_plane = nullptr;
    float4 *synth_stars = nullptr;
    // multi-GPU
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    cudaStream_t comm_stream = nullptr;   // master broadcasts run here, beside the bulk stream
    cudaEvent_t comm_fork = nullptr, masters_ready = nullptr;
    int reduce_root = 0;
};

static const int kChunk = 1024;

static int fail(oss_ctx *c, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    return code;
}

// the stream a pipeline's bulk kernels (calibration, sky, copies; also the masters) go to
static cudaStream_t bulk_of(oss_ctx *c, int) { return c->bulk; }

static void select_pipe(oss_ctx *c, int p)
{
    Pipe &o = c->pipes[c->cur];
    o.batch_counter = c->batch_counter;
    Pipe &n = c->pipes[p];
    c->cur = p;
    c->stream = bulk_of(c, p); c->copy_stream = n.copy_stream;
    c->ws = n.ws; c->d_xy = n.d_xy; c->d_nxy = n.d_nxy;
    for (int i = 0; i < 2; i++) { c->raw_stage[i] = n.raw_stage[i]; c->copy_done[i] = n.copy_done[i]; c->raw_free[i] = n.raw_free[i]; }
    c->batch_counter = n.batch_counter;
}

static int flush_warps(oss_ctx *c, size_t keep, int bands, int (*after_band)(oss_ctx *, int, int));

static cudaError_t sync_all(oss_ctx *c, bool with_comm = true)
{
    cudaError_t e = cudaSuccess;
    if (flush_warps(c, 0, 1, nullptr) != OSS_OK) e = cudaErrorLaunchFailure;
    for (int p = 0; p < OSS_MAX_PIPES; p++) {
        if (c->pipes[p].copy_stream) { cudaError_t r = cudaStreamSynchronize(c->pipes[p].copy_stream); if (r != cudaSuccess) e = r; }
        if (c->pipes[p].chain_stream) { cudaError_t r = cudaStreamSynchronize(c->pipes[p].chain_stream); if (r != cudaSuccess) e = r; }
    }
    if (c->bulk) { cudaError_t r = cudaStreamSynchronize(c->bulk); if (r != cudaSuccess) e = r; }
    if (c->fp_stream) { cudaError_t r = cudaStreamSynchronize(c->fp_stream); if (r != cudaSuccess) e = r; }
    if (with_comm && c->comm_stream) { cudaError_t r = cudaStreamSynchronize(c->comm_stream); if (r != cudaSuccess) e = r; }
    c->dirty = false;
    return e;
}

#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return fail(c, OSS_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

#define NK(call)                                                                                      \
    do {                                                                                              \
        ncclResult_t r_ = (call);                                                                     \
        if (r_ != ncclSuccess)                                                                        \
            return fail(c, OSS_E_CUDA, "%s failed: %s", #call, g_nccl.GetErrorString(r_));           \
    } while (0)

// ---- profiling-aware launch ------------------------------------------------------------
static int prof_entry(oss_ctx *c, const char *name)
{
    for (size_t i = 0; i < c->prof.size(); i++)
        if (c->prof[i].name == name) return (int)i;
    c->prof.push_back({name, 0.0, 0});
    return (int)c->prof.size() - 1;
}

static cudaEvent_t get_event(oss_ctx *c)
{
    if (!c->event_pool.empty()) { cudaEvent_t e = c->event_pool.back(); c->event_pool.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

static void prof_flush(oss_ctx *c)
{
    // dev aid: OSS_B200_TIMELINE=<file> appends "name pipe start_ms end_ms" per launch, relative to the first launch of the flush
    if (const char *path = getenv("OSS_B200_TIMELINE")) {
        if (!c->prof_pending.empty()) {
            if (FILE *fp = fopen(path, "a")) {
                fprintf(fp, "# flush of %zu launches\n", c->prof_pending.size());
                for (auto &p : c->prof_pending) {
                    float t0 = 0.f, t1 = 0.f;
                    cudaEventElapsedTime(&t0, c->prof_pending[0].a, p.a);
                    cudaEventElapsedTime(&t1, c->prof_pending[0].a, p.b);
                    fprintf(fp, "%s %d %.4f %.4f\n", c->prof[p.entry].name.c_str(), p.pipe, t0, t1);
                }
                fclose(fp);
            }
        }
    }
    for (auto &p : c->prof_pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) { c->prof[p.entry].ms += ms; c->prof[p.entry].launches++; }
        c->event_pool.push_back(p.a);
        c->event_pool.push_back(p.b);
    }
    c->prof_pending.clear();
}

#define LAUNCH(name, kern, grid, block, smem, ...)                                   \
    do {                                                                             \
        ProfPending pp_{-1, nullptr, nullptr, c->cur};                               \
        if (c->profiling) {                                                          \
            pp_.entry = prof_entry(c, name);                                         \
            pp_.a = get_event(c); pp_.b = get_event(c);                              \
            cudaEventRecord(pp_.a, c->stream);                                       \
        }                                                                            \
        if (c->stream != c->bulk) c->dirty = true;                                   \
        kern<<<grid, block, smem, c->stream>>>(__VA_ARGS__);                         \
        c->launches++;                                                               \
        if (c->profiling) { cudaEventRecord(pp_.b, c->stream); c->prof_pending.push_back(pp_); } \
        CK(cudaGetLastError());                                                      \
    } while (0)

template <typename T>
static int dalloc(oss_ctx *c, T **p, size_t count, bool geom_scoped = true)
{
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, count * sizeof(T) + 256);
    if (e != cudaSuccess) return fail(c, OSS_E_NOMEM, "cudaMalloc(%zu bytes) failed: %s", count * sizeof(T), cudaGetErrorString(e));
    if (geom_scoped) c->allocs.push_back(q);
    *p = (T *)q;
    return OSS_OK;
}
#define DA(ptr, count) do { int r_ = dalloc(c, &(ptr), (size_t)(count)); if (r_) return r_; } while (0)

static void free_geometry(oss_ctx *c)
{
    cudaDeviceSynchronize();
    c->pending_warps.clear(); // batches detected but never finished belong to the work spaces freed below
    for (void *p : c->allocs) cudaFree(p);
    c->allocs.clear();
    for (int i = 0; i < 4; i++) c->master[i] = nullptr;
    for (int p = 0; p < OSS_MAX_PIPES; p++) {
        Pipe &pp = c->pipes[p];
        for (int i = 0; i < 2; i++) {
            pp.raw_stage[i] = nullptr;
            if (pp.copy_done[i]) { cudaEventDestroy(pp.copy_done[i]); pp.copy_done[i] = nullptr; }
            if (pp.raw_free[i]) { cudaEventDestroy(pp.raw_free[i]); pp.raw_free[i] = nullptr; }
        }
        if (pp.warp_done) { cudaEventDestroy(pp.warp_done); pp.warp_done = nullptr; }
        if (pp.detect_done) { cudaEventDestroy(pp.detect_done); pp.detect_done = nullptr; }
        if (pp.sky_done) { cudaEventDestroy(pp.sky_done); pp.sky_done = nullptr; }
        if (pp.cal_done) { cudaEventDestroy(pp.cal_done); pp.cal_done = nullptr; }
        if (pp.med_done) { cudaEventDestroy(pp.med_done); pp.med_done = nullptr; }
        if (pp.side_fork) { cudaEventDestroy(pp.side_fork); pp.side_fork = nullptr; }
        if (pp.side_join) { cudaEventDestroy(pp.side_join); pp.side_join = nullptr; }
        pp.ws = Workspace{};
        pp.d_xy = pp.d_nxy = nullptr;
        pp.batch_counter = 0;
    }
    c->light_batches = 0;
    c->has_geom = false;
    c->has_ref = false;
    c->synth_plane = nullptr;
    c->synth_stars = nullptr;
}

// ---- SM partition ----------------------------------------------------------------------
struct GreenApi {
    bool looked = false, ok = false;
    CUresult (*DeviceGet)(CUdevice *, int) = nullptr;
    CUresult (*DeviceGetDevResource)(CUdevice, CUdevResource *, CUdevResourceType) = nullptr;
    CUresult (*DevSmResourceSplitByCount)(CUdevResource *, unsigned int *, const CUdevResource *, CUdevResource *, unsigned int, unsigned int) = nullptr;
    CUresult (*DevResourceGenerateDesc)(CUdevResourceDesc *, CUdevResource *, unsigned int) = nullptr;
    CUresult (*GreenCtxCreate)(CUgreenCtx *, CUdevResourceDesc, CUdevice, unsigned int) = nullptr;
    CUresult (*GreenCtxDestroy)(CUgreenCtx) = nullptr;
    CUresult (*GreenCtxStreamCreate)(CUstream *, CUgreenCtx, unsigned int, int) = nullptr;
};
static GreenApi g_green;

static bool green_load()
{
    if (g_green.looked) return g_green.ok;
    g_green.looked = true;
    bool ok = true;
    auto get = [&](const char *name, void **fn) {
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !*fn) ok = false;
    };
    get("cuDeviceGet", (void **)&g_green.DeviceGet);
    get("cuDeviceGetDevResource", (void **)&g_green.DeviceGetDevResource);
    get("cuDevSmResourceSplitByCount", (void **)&g_green.DevSmResourceSplitByCount);
    get("cuDevResourceGenerateDesc", (void **)&g_green.DevResourceGenerateDesc);
    get("cuGreenCtxCreate", (void **)&g_green.GreenCtxCreate);
    get("cuGreenCtxDestroy", (void **)&g_green.GreenCtxDestroy);
    get("cuGreenCtxStreamCreate", (void **)&g_green.GreenCtxStreamCreate);
    g_green.ok = ok;
    return ok;
}

static void split_teardown(oss_ctx *c)
{
    if (c->fp_stream) { cudaStreamDestroy(c->fp_stream); c->fp_stream = nullptr; }
    if (c->green_fp || c->green_hbm) {
        // the bulk stream lives in green_hbm: replace it by an ordinary low-priority stream first
        if (c->bulk) cudaStreamDestroy(c->bulk);
        c->bulk = nullptr;
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        cudaStreamCreateWithPriority(&c->bulk, cudaStreamNonBlocking, lo);
        if (c->green_fp) g_green.GreenCtxDestroy(c->green_fp);
        if (c->green_hbm) g_green.GreenCtxDestroy(c->green_hbm);
        c->green_fp = c->green_hbm = nullptr;
    }
    c->split_fp_got = c->split_hbm_got = 0;
}

// (re)build the partition: `fp_sms` SMs (rounded up to the device's granularity, 8 on sm_100) for the sky kernel, the rest
// for the HBM-bound bulk kernels.  Called with every stream idle.
static int split_setup(oss_ctx *c, int fp_sms)
{
    split_teardown(c);
    if (fp_sms <= 0) return OSS_OK;
    if (!green_load()) return fail(c, OSS_E_CUDA, "this driver has no green-context entry points (SM partition unavailable)");
    CUdevice dev;
    CUdevResource all, grp, rest;
    unsigned int n = 1;
    CUdevResourceDesc d_fp, d_hbm;
#define GK(call) do { CUresult r_ = (call); if (r_ != CUDA_SUCCESS) { split_teardown(c); return fail(c, OSS_E_CUDA, "%s failed (CUresult %d)", #call, (int)r_); } } while (0)
    GK(g_green.DeviceGet(&dev, c->device));
    GK(g_green.DeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM));
    if (fp_sms >= (int)all.sm.smCount) return fail(c, OSS_E_ARG, "SM split %d leaves nothing of %u SMs for the HBM-bound kernels", fp_sms, all.sm.smCount);
    GK(g_green.DevSmResourceSplitByCount(&grp, &n, &all, &rest, 0, (unsigned)fp_sms));
    if (n < 1 || rest.sm.smCount == 0) return fail(c, OSS_E_ARG, "SM split %d not realisable on %u SMs", fp_sms, all.sm.smCount);
    GK(g_green.DevResourceGenerateDesc(&d_fp, &grp, 1));
    GK(g_green.DevResourceGenerateDesc(&d_hbm, &rest, 1));
    GK(g_green.GreenCtxCreate(&c->green_fp, d_fp, dev, CU_GREEN_CTX_DEFAULT_STREAM));
    GK(g_green.GreenCtxCreate(&c->green_hbm, d_hbm, dev, CU_GREEN_CTX_DEFAULT_STREAM));
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    CUstream s_fp = nullptr, s_hbm = nullptr;
    GK(g_green.GreenCtxStreamCreate(&s_fp, c->green_fp, CU_STREAM_NON_BLOCKING, lo));
    GK(g_green.GreenCtxStreamCreate(&s_hbm, c->green_hbm, CU_STREAM_NON_BLOCKING, lo));
#undef GK
    c->fp_stream = (cudaStream_t)s_fp;
    if (c->bulk) cudaStreamDestroy(c->bulk);
    c->bulk = (cudaStream_t)s_hbm;
    c->split_fp_got = (int)grp.sm.smCount;
    c->split_hbm_got = (int)rest.sm.smCount;
    return OSS_OK;
}

// ---- exported: lifetime ----------------------------------------------------------------
extern "C" {

const char *oss_version(void) { return "oss-b200 0.2 (sm_100a)"; }

int oss_device_count(void)
{
    int n = 0;
    return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

const char *oss_status_string(int s)
{
    switch (s) {
    case OSS_OK: return "ok";
    case OSS_E_CUDA: return "CUDA/NCCL failure or no CUDA device (this library has no CPU path)";
    case OSS_E_ARG: return "bad argument or call order";
    case OSS_E_TYPE: return "Images must be the same type.";
    case OSS_E_SIZE: return "Images must all be the same size.";
    case OSS_E_THREADS: return "Number of threads must be at least 1.";
    case OSS_E_NO_ALIGNED: return "No images could be aligned to the reference image. Try using a lower tolerance.";
    case OSS_E_CAPACITY: return "detection workspace capacity exceeded";
    case OSS_E_NOMEM: return "out of device memory";
    }
    return "unknown";
}

int oss_create(oss_ctx **out, int device)
{
    if (!out) return OSS_E_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return OSS_E_CUDA;
    if (device < 0 || device >= ndev) return OSS_E_ARG;
    oss_ctx *c = new oss_ctx();
    c->device = device;
    bool ok = cudaSetDevice(device) == cudaSuccess;
    int prio_least = 0, prio_greatest = 0; // numerically greatest <= least
    ok = ok && cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest) == cudaSuccess;
    for (int p = 0; ok && p < OSS_MAX_PIPES; p++)
        ok = cudaStreamCreateWithPriority(&c->pipes[p].chain_stream, cudaStreamNonBlocking, prio_greatest) == cudaSuccess &&
             cudaStreamCreateWithPriority(&c->pipes[p].side_stream, cudaStreamNonBlocking, prio_greatest) == cudaSuccess &&
             cudaStreamCreateWithPriority(&c->pipes[p].copy_stream, cudaStreamNonBlocking, prio_greatest) == cudaSuccess;
    ok = ok && cudaStreamCreateWithPriority(&c->bulk, cudaStreamNonBlocking, prio_least) == cudaSuccess;
    if (!ok) { oss_destroy(c); return OSS_E_CUDA; } // oss_destroy null-checks every member: fine on a partly built ctx
    if (cudaEventCreateWithFlags(&c->ref_ready, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ref_copied, cudaEventDisableTiming) != cudaSuccess ||
        cudaMallocHost((void **)&c->h_ref_ds, sizeof(DetectState)) != cudaSuccess) {
        oss_destroy(c);
        return OSS_E_CUDA;
    }
    if (const char *e = getenv("OSS_B200_PIPES")) c->want_pipes = atoi(e);
    if (const char *e = getenv("OSS_B200_SPLIT")) c->split_fp = atoi(e);
    if (const char *e = getenv("OSS_B200_CAL_RING")) c->cal_ring = atoi(e) != 0;
    if (const char *e = getenv("OSS_B200_CAP_DIV")) { int v = atoi(e); if (v >= 1 && v <= 64) c->cap_one_in = v; } // default only
    c->stream = c->bulk;
    c->copy_stream = c->pipes[0].copy_stream;
    // constant tables
    float taps[31];
    for (int i = 0; i < 16; i++) { memcpy(&taps[i], &kGaussBits[i], 4); taps[30 - i] = taps[i]; }
    std::vector<uint32_t> ijk;
    for (int i = 0; i < OSS_NOBJS - 2; i++)
        for (int j = i + 1; j < OSS_NOBJS - 1; j++)
            for (int k = j + 1; k < OSS_NOBJS; k++) ijk.push_back((uint32_t)i | ((uint32_t)j << 8) | ((uint32_t)k << 16));
    if (cudaMemcpyToSymbol(c_gauss, taps, sizeof taps) != cudaSuccess ||
        cudaMemcpyToSymbol(d_tri_ijk, ijk.data(), ijk.size() * 4) != cudaSuccess ||
        cudaFuncSetAttribute(k_sky_sub_stats, cudaFuncAttributeMaxDynamicSharedMemorySize, kSkySmemBytes) != cudaSuccess ||
        cudaFuncSetAttribute(k_focas, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FocasSmem)) != cudaSuccess ||
        cudaFuncSetAttribute(k_components<CW_CAP, CW_WARPS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCompSmemBytes) != cudaSuccess ||
        cudaFuncSetAttribute(k_components<CW_CAP_SMALL, CW_WARPS_SMALL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCompSmemBytesSmall) != cudaSuccess ||
        cudaFuncSetAttribute(k_components_big<CBIG_SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CompBigSmem<CBIG_SMALL>)) != cudaSuccess ||
        cudaFuncSetAttribute(k_components_big<OSS_DEBLEND_MAX - 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CompBigSmem<OSS_DEBLEND_MAX - 1>)) != cudaSuccess ||
        cudaFuncSetAttribute(k_warp_accumulate_tiled<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, warp_tiled_smem<3>()) != cudaSuccess ||
        cudaFuncSetAttribute(k_warp_accumulate_tiled<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, warp_tiled_smem<1>()) != cudaSuccess ||
        cudaFuncSetAttribute(k_warp_accumulate_tma<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WarpTmaSmem<3>) + 128) != cudaSuccess ||
        cudaFuncSetAttribute(k_warp_accumulate_tma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WarpTmaSmem<1>) + 128) != cudaSuccess) {
        oss_destroy(c);
        return OSS_E_CUDA;
    }
    *out = c;
    return OSS_OK;
}

void oss_destroy(oss_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    free_geometry(c);
    for (void *p : c->user_allocs) cudaFree(p);
    for (auto p : c->h_ds_chunks) cudaFreeHost(p);
    for (auto p : c->h_al_chunks) cudaFreeHost(p);
    prof_flush(c);
    for (auto e : c->event_pool) cudaEventDestroy(e);
    for (int p = 0; p < OSS_MAX_PIPES; p++) { if (c->pipes[p].copy_stream) cudaStreamDestroy(c->pipes[p].copy_stream); if (c->pipes[p].chain_stream) cudaStreamDestroy(c->pipes[p].chain_stream); if (c->pipes[p].side_stream) cudaStreamDestroy(c->pipes[p].side_stream); }
    split_teardown(c);
    for (auto &t : c->tickets) if (t) cudaEventDestroy(t);
    if (c->ticket_stream) cudaStreamDestroy(c->ticket_stream);
    if (c->bulk) cudaStreamDestroy(c->bulk);
    if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
    if (c->comm_fork) cudaEventDestroy(c->comm_fork);
    if (c->masters_ready) cudaEventDestroy(c->masters_ready);
    if (c->ref_ready) cudaEventDestroy(c->ref_ready);
    if (c->ref_copied) cudaEventDestroy(c->ref_copied);
    if (c->h_ref_ds) cudaFreeHost(c->h_ref_ds);
    delete c;
}

const char *oss_last_error(const oss_ctx *c) { return c ? c->err.c_str() : "no context"; }

void *oss_host_alloc(size_t bytes)
{
    void *p = nullptr;
    return cudaMallocHost(&p, bytes) == cudaSuccess ? p : nullptr;
}
void oss_host_free(void *p) { if (p) cudaFreeHost(p); }

void *oss_device_alloc(oss_ctx *c, size_t bytes)
{
    if (!c) return nullptr;
    cudaSetDevice(c->device);
    void *p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) { c->err = "cudaMalloc failed"; return nullptr; }
    c->user_allocs.push_back(p);
    return p;
}
void oss_device_free(oss_ctx *c, void *p)
{
    if (!c || !p) return;
    for (size_t i = 0; i < c->user_allocs.size(); i++)
        if (c->user_allocs[i] == p) { c->user_allocs.erase(c->user_allocs.begin() + i); break; }
    cudaFree(p);
}
int oss_memcpy_h2d(oss_ctx *c, void *d, const void *s, size_t n)
{
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return OSS_OK;
}
int oss_memcpy_d2h(oss_ctx *c, void *d, const void *s, size_t n)
{
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return OSS_OK;
}

// ---- geometry --------------------------------------------------------------------------
// 3-D tensor map over a pipeline's calibrated planes for k_warp_accumulate_tma; false (the cp.async kernel is used) when the
// rows are not 16-byte multiples or the driver entry point is missing
static bool make_cal_map(oss_ctx *c, Pipe &pp, int slots)
{
    typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeTiled encode = nullptr;
    static bool looked = false;
    if (!looked) {
        looked = true;
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            encode = (EncodeTiled)fn;
    }
    if (getenv("OSS_B200_NO_TMA")) return false;
    const Geom &g = c->g;
    const long long rowf = (long long)g.W * g.C;
    if (!encode || (rowf * 4) % 16 != 0 || (pp.ws.stridePC * 4) % 16 != 0 || ((uintptr_t)pp.ws.cal & 15)) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)rowf, (cuuint64_t)g.H, (cuuint64_t)slots};
    const cuuint64_t strides[2] = {(cuuint64_t)rowf * 4, (cuuint64_t)pp.ws.stridePC * 4};
    const cuuint32_t box[3] = {(cuuint32_t)(WT_COLS * g.C), (cuuint32_t)WT_ROWS, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return encode(&pp.cal_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, pp.ws.cal, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int oss_set_geometry(oss_ctx *c, int W, int H, int C, int sample, int slots)
{
    if (!c) return OSS_E_ARG;
    if (W < 8 || H < 8 || (C != 1 && C != 3) || sample < OSS_U8 || sample > OSS_F32)
        return fail(c, OSS_E_ARG, "unsupported geometry %dx%dx%d sample %d", W, H, C, sample);
    if ((long long)W * H >= (1ll << 31)) return fail(c, OSS_E_ARG, "frame too large");
    if (slots < 1) slots = 1;
    if (slots > OSS_MAX_BATCH) slots = OSS_MAX_BATCH;
    CK(cudaSetDevice(c->device));
    free_geometry(c);
    if ((c->split_fp > 0) != (c->fp_stream != nullptr) || (c->split_fp > 0 && c->split_fp_got < c->split_fp)) {
        int r_ = split_setup(c, c->split_fp);
        if (r_) return r_;
        c->stream = c->bulk;
    }
    Geom g{};
    g.W = W; g.H = H; g.C = C; g.sample = sample;
    g.lw = W / 8; g.lh = H / 8;
    g.nseg = (W + OSS_SEG - 1) / OSS_SEG;
    g.roi_x = W / 20; g.roi_y = H / 20;
    g.roi_w = (int)(W * 0.9); g.roi_h = (int)(H * 0.9);
    g.P = (long long)W * H;
    long long div = c->cap_one_in;
    long long cap = g.P / div;
    if (cap < 65536) cap = g.P < 65536 ? g.P : 65536; // small frames: every pixel may be a candidate
    g.cap = (int)cap;
    c->g = g;
    c->slots = slots;
    const size_t ssz = sample == OSS_U8 ? 1 : (sample == OSS_U16 ? 2 : 4);
    c->frame_bytes = (size_t)g.P * C * ssz;
    c->stage_stride = (c->frame_bytes + 255) / 256 * 256;
    const long long PC = g.P * C, nsegs = (long long)H * g.nseg, lo = (long long)g.lw * g.lh;
    c->npipes = c->want_pipes < 1 ? 1 : (c->want_pipes > OSS_MAX_PIPES ? OSS_MAX_PIPES : c->want_pipes);
    auto pad64 = [](long long v) { return (v + 63) / 64 * 64; }; // keep every slot 256-byte aligned
    for (int pi = 0; pi < c->npipes; pi++) {
        Pipe &pp = c->pipes[pi];
        Workspace &w = pp.ws;
        w = Workspace{};
        w.strideP = pad64(g.P); w.stridePC = pad64(PC); w.strideLo = lo; w.strideSeg = nsegs;
        w.stat_blocks = sky_max_blocks(W, H);
        w.stridePart = 2ll * w.stat_blocks;
        long long scan_max = nsegs + 1 > cap ? nsegs + 1 : cap;
        w.strideScan = (scan_max + SCAN_ITEMS - 1) / SCAN_ITEMS + 1;
        DA(w.cal, slots * w.stridePC + 1024); // + slack: staged warp rows may run past the last frame's end
        if (C == 3) { DA(w.gray, slots * w.strideP); } else w.gray = w.cal;
        DA(w.stars, slots * w.strideP);
        DA(w.lo, slots * lo); DA(w.med, slots * lo);
        DA(w.segmax, slots * nsegs); DA(w.segmask, slots * nsegs); DA(w.segoff, slots * (nsegs + 1));
        DA(w.partials, slots * w.stridePart);
        DA(w.ds, slots);
        DA(w.cand_idx, slots * cap); DA(w.cand_val, slots * cap); DA(w.nbr, slots * cap); DA(w.parent, slots * cap);
        DA(w.owner, slots * cap); DA(w.comp_size, slots * cap); DA(w.comp_peak, slots * cap);
        DA(w.flag, slots * cap); DA(w.flagscan, slots * cap);
        DA(w.comp_root, slots * cap); DA(w.comp_off, slots * cap); DA(w.comp_nstar, slots * cap);
        DA(w.comp_staroff, slots * cap);
        DA(w.scratch, 12ll * slots * cap);
        DA(w.star_tmp, slots * cap); DA(w.star_out, slots * cap);
        DA(w.blocksum, slots * w.strideScan);
        DA(w.align, slots);
        DA(w.warp_ab, (long long)slots * W); DA(w.warp_xy, (long long)slots * H);
        DA(pp.d_xy, (slots + 1) * 2 * OSS_NOBJS); DA(pp.d_nxy, slots + 1);
        for (int i = 0; i < 2; i++) {
            DA(pp.raw_stage[i], (size_t)slots * c->stage_stride);
            CK(cudaEventCreateWithFlags(&pp.copy_done[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&pp.raw_free[i], cudaEventDisableTiming));
        }
        CK(cudaEventCreateWithFlags(&pp.warp_done, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&pp.detect_done, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&pp.sky_done, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&pp.cal_done, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&pp.med_done, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&pp.side_fork, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&pp.side_join, cudaEventDisableTiming));
        pp.has_cal_map = make_cal_map(c, pp, slots);
        pp.batch_counter = 0;
    }
    c->cur = 0;
    c->batch_counter = 0;
    select_pipe(c, 0);
    DA(c->dx, g.lw); DA(c->dy, g.lh); DA(c->ux, W); DA(c->uy, H);
    DA(c->sum, PC); DA(c->ref_cal, PC); DA(c->result, PC);
    DA(c->d_valid, 4);
    DA(c->d_ref, 1); DA(c->d_ref_tmp, 1);
    DA(c->d_ref_stars, g.cap);
    DA(c->d_dbg_matches, 3 * OSS_MAX_MATCH); DA(c->d_dbg_table, OSS_NOBJS * OSS_NOBJS); DA(c->d_dbg_clk, 8);
    DA(c->d_M, 8); DA(c->d_flat_ratio, 4); DA(c->d_flat_partial, 1024);
    LAUNCH("resize_taps", k_resize_taps, (g.lw + 255) / 256, 256, 0, c->dx, W, g.lw, 1);
    LAUNCH("resize_taps", k_resize_taps, (g.lh + 255) / 256, 256, 0, c->dy, H, g.lh, 0);
    LAUNCH("resize_taps", k_resize_taps, (W + 255) / 256, 256, 0, c->ux, g.lw, W, 1);
    LAUNCH("resize_taps", k_resize_taps, (H + 255) / 256, 256, 0, c->uy, g.lh, H, 0);
    CK(cudaMemsetAsync(c->sum, 0, PC * sizeof(float), c->stream));
    CK(cudaMemsetAsync(c->d_valid, 0, 16, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->has_geom = true;
    c->nlights = 0;
    c->extra_valid = 0;
    c->batch_counter = 0;
    return OSS_OK;
}

}  // extern "C"

// ---- helpers ---------------------------------------------------------------------------
static const int kScanGrid = 24;    // CTAs per frame of the scan phases
static const int kSparseGrid = 74;  // CTAs (of 256 threads) per frame of the grid-stride kernels over candidate lists

static int run_scan(oss_ctx *c, const int *in, long long strideIn, int *out, long long strideOut, const int *n_dev,
                    int n_dev_stride, int n_host_max, int nframes, int *total_out, int total_stride, int write_total_at_n = 0)
{
    Workspace &w = c->ws;
    int nb = (n_host_max + SCAN_ITEMS - 1) / SCAN_ITEMS;
    if (nb < 1) nb = 1;
    int n_host = n_dev ? 0 : n_host_max;
    // few CTAs per frame striding over the blocks: sized for the work usually there, not for the capacity (k_scan_phase1)
    const int gx = nb < kScanGrid ? nb : kScanGrid;
    LAUNCH("scan", k_scan_phase1, dim3(gx, nframes), 1024, 0, in, strideIn, out, strideOut, n_dev, n_dev_stride, n_host,
           w.blocksum, w.strideScan);
    LAUNCH("scan", k_scan_phase2, nframes, 1024, 0, w.blocksum, w.strideScan, n_dev, n_dev_stride, n_host, out, strideOut,
           write_total_at_n, total_out, total_stride);
    if (nb > 1)
        LAUNCH("scan", k_scan_phase3, dim3(gx, nframes), 1024, 0, out, strideOut, n_dev, n_dev_stride, n_host, w.blocksum,
               w.strideScan);
    return OSS_OK;
}

template <typename T, int C>
static int launch_calibrate(oss_ctx *c, const FramePtrs &fp, int n, const float *b, const float *d, const float *f)
{
    Workspace &w = c->ws;
    const Geom &g = c->g;
    const int key = (b ? 4 : 0) | (d ? 2 : 0) | (f ? 1 : 0);
    // batches of u16 frames that share masters: masters stay in registers across the batch
    const bool batched = std::is_same<T, uint16_t>::value && key != 0 && n > 1;
    long long P_tail_from = 0; // pixels [P_tail_from, P) still need the generic kernel
    if (batched) {
        constexpr int PXB = C == 3 ? 4 : 8;
        unsigned gridb = (unsigned)((g.P / PXB + 255) / 256);
        const bool ring = c->cal_ring && g.P >= CALR_NT * PXB;
        const unsigned gridr = (unsigned)(g.P / (CALR_NT * PXB)); // whole chunks; the rest goes to the generic kernel below
        const int smemr = calr_smem_bytes<C>(n);
#define CALB(HB, HD, HF) { if (ring) { auto kr = k_calibrate_gray_ring<C, HB, HD, HF>; \
            const unsigned abit = 1u << (key + (C == 3 ? 8 : 0)); /* function attributes are per device: remembered per context */ \
            if (!(c->calr_attr & abit)) { CK(cudaFuncSetAttribute(kr, cudaFuncAttributeMaxDynamicSharedMemorySize, calr_smem_bytes<C>(OSS_MAX_BATCH))); c->calr_attr |= abit; } \
            LAUNCH("calibrate_gray", kr, gridr, CALR_NT, smemr, fp, n, b, d, f, w.cal, w.stridePC, w.gray, w.strideP); } \
          else { auto kf = k_calibrate_gray_batch<C, HB, HD, HF>; LAUNCH("calibrate_gray", kf, gridb, 256, 0, fp, n, b, d, f, w.cal, w.stridePC, w.gray, w.strideP, g.P); } }
        switch (key) {
        case 1: CALB(false, false, true); break;
        case 2: CALB(false, true, false); break;
        case 3: CALB(false, true, true); break;
        case 4: CALB(true, false, false); break;
        case 5: CALB(true, false, true); break;
        case 6: CALB(true, true, false); break;
        case 7: CALB(true, true, true); break;
        }
#undef CALB
        P_tail_from = ring ? (long long)gridr * CALR_NT * PXB : g.P / PXB * PXB;
        if (P_tail_from == g.P) return OSS_OK;
    }
    // generic kernel: whole frames, or only the ragged tail the batched kernel left
    const long long first_group = P_tail_from / Sample<T>::PX; // tail starts inside this PX-group
    const long long groups = (g.P + Sample<T>::PX - 1) / Sample<T>::PX - first_group;
    dim3 grid((unsigned)((groups + 255) / 256), n);
#define CAL(HB, HD, HF) { auto kf = k_calibrate_gray<T, C, HB, HD, HF>; LAUNCH(batched ? "calibrate_tail" : "calibrate_gray", kf, grid, 256, 0, fp, b, d, f, w.cal, w.stridePC, w.gray, w.strideP, g.P, first_group); }
    switch (key) {
    case 0: CAL(false, false, false); break;
    case 1: CAL(false, false, true); break;
    case 2: CAL(false, true, false); break;
    case 3: CAL(false, true, true); break;
    case 4: CAL(true, false, false); break;
    case 5: CAL(true, false, true); break;
    case 6: CAL(true, true, false); break;
    case 7: CAL(true, true, true); break;
    }
#undef CAL
    return OSS_OK;
}

static int dispatch_calibrate(oss_ctx *c, const FramePtrs &fp, int n, bool use_masters)
{
    const float *b = use_masters ? c->master[OSS_MASTER_BIAS] : nullptr;
    const float *d = use_masters ? c->master[OSS_MASTER_DARK] : nullptr;
    const float *f = use_masters ? c->master[OSS_MASTER_FLAT] : nullptr;
    const Geom &g = c->g;
    if (g.sample == OSS_U16) return g.C == 3 ? launch_calibrate<uint16_t, 3>(c, fp, n, b, d, f) : launch_calibrate<uint16_t, 1>(c, fp, n, b, d, f);
    if (g.sample == OSS_U8) return g.C == 3 ? launch_calibrate<uint8_t, 3>(c, fp, n, b, d, f) : launch_calibrate<uint8_t, 1>(c, fp, n, b, d, f);
    return g.C == 3 ? launch_calibrate<float, 3>(c, fp, n, b, d, f) : launch_calibrate<float, 1>(c, fp, n, b, d, f);
}

// stage n host frames into raw_stage[set] (async on the copy stream) or pass device pointers through
// wait == false: only queue the copies (prefetch of a later batch); the consumer calls stage_wait before its first kernel
static int stage_frames(oss_ctx *c, const void *const *frames, int n, int mem, FramePtrs *fp, int *set_out, bool wait = true,
                        size_t byte0 = 0, size_t nbytes = 0)
{
    if (!nbytes) nbytes = c->frame_bytes;
    for (int i = 0; i < n; i++) {
        if (!frames[i]) return fail(c, OSS_E_ARG, "null frame pointer");
        if (((uintptr_t)frames[i] & 15) && mem == OSS_MEM_DEVICE) return fail(c, OSS_E_ARG, "device frames must be 16-byte aligned");
    }
    *set_out = -1;
    if (mem == OSS_MEM_DEVICE) {
        for (int i = 0; i < n; i++) fp->p[i] = (const unsigned char *)frames[i] + byte0;
        return OSS_OK;
    }
    int set = (int)(c->batch_counter++ & 1);
    CK(cudaStreamWaitEvent(c->copy_stream, c->raw_free[set], 0));
    for (int i = 0; i < n; i++) {
        unsigned char *dst = c->raw_stage[set] + (size_t)i * c->stage_stride;
        CK(cudaMemcpyAsync(dst, (const unsigned char *)frames[i] + byte0, nbytes, cudaMemcpyHostToDevice, c->copy_stream));
        fp->p[i] = dst;
    }
    CK(cudaEventRecord(c->copy_done[set], c->copy_stream));
    if (wait) CK(cudaStreamWaitEvent(c->stream, c->copy_done[set], 0));
    *set_out = set;
    return OSS_OK;
}

static int stage_wait(oss_ctx *c, int set)
{
    if (set >= 0) CK(cudaStreamWaitEvent(c->stream, c->copy_done[set], 0));
    return OSS_OK;
}

static int run_detect_tail(oss_ctx *c, int n, int threshold);

// The detection of a batch in four pieces, so that oss_add_lights can interleave two batches (all on the selected pipeline):
//   detect_calibrate  BULK stream: calibration + gray of the batch's frames
//   detect_lowres     CHAIN stream: resize / 8 and median 5x5 of the gray planes — 1/64 of the pixels, two small kernels that
//                     would leave the GPU mostly idle on the bulk stream; here they run beside the NEXT batch's calibration
//   detect_sky        BULK stream (behind the low-res grid): sky background, subtraction, statistics
//   run_detect_tail   CHAIN stream: thresholding ... star lists
static int detect_calibrate(oss_ctx *c, const FramePtrs &fp, int n, bool use_masters, int stage_set)
{
    c->slot0_is_ref = false;
    if (use_masters && c->masters_ready) CK(cudaStreamWaitEvent(c->stream, c->masters_ready, 0)); // broadcasts in flight
    int r = dispatch_calibrate(c, fp, n, use_masters);
    if (r) return r;
    if (stage_set >= 0) CK(cudaEventRecord(c->raw_free[stage_set], c->stream));
    CK(cudaEventRecord(c->pipes[c->cur].cal_done, c->stream));
    return OSS_OK;
}

static int detect_lowres(oss_ctx *c, int n)
{
    Workspace &w = c->ws;
    const Geom &g = c->g;
    cudaStream_t bulk = c->stream;
    c->stream = c->pipes[c->cur].chain_stream;
    CK(cudaStreamWaitEvent(c->stream, c->pipes[c->cur].cal_done, 0));
    LAUNCH("resize_down", k_resize_down, dim3((g.lw + 31) / 32, (g.lh + 7) / 8, n), 256, 0, w.gray, g.C == 3 ? w.strideP : w.stridePC, g.W,
           c->dx, c->dy, w.lo, w.strideLo, g.lw, g.lh);
    LAUNCH("median5", k_median5, dim3((g.lw + 31) / 32, (g.lh + 7) / 8, n), 256, 0, w.lo, w.med, w.strideLo, g.lw, g.lh);
    CK(cudaEventRecord(c->pipes[c->cur].med_done, c->stream));
    c->stream = bulk;
    return OSS_OK;
}

static int detect_sky(oss_ctx *c, int n)
{
    Workspace &w = c->ws;
    const Geom &g = c->g;
    if (c->fp_stream) { // SM partition: the sky kernel runs beside the HBM-bound kernels, ordered by events only
        c->stream = c->fp_stream;
        CK(cudaStreamWaitEvent(c->stream, c->pipes[c->cur].cal_done, 0));
    }
    CK(cudaStreamWaitEvent(c->stream, c->pipes[c->cur].med_done, 0));
    const long long strideGray = g.C == 3 ? w.strideP : w.stridePC;
    const int seg_h = sky_segment_rows(g.W, g.H, n);
    const dim3 grid((g.W + SKY_W - 1) / SKY_W, (g.H + seg_h - 1) / seg_h, n);
    c->sky_blocks = grid.x * grid.y;
    LAUNCH("sky_sub_stats", k_sky_sub_stats, grid, SKY_NT, kSkySmemBytes, w.gray, strideGray, w.med, w.strideLo, g.lw, c->ux, c->uy,
           w.stars, w.strideP, w.segmax, w.strideSeg, g.nseg, w.partials, w.stridePart, g, seg_h);
    // the rest of the batch's detection (and whatever the caller queues next) goes to the pipeline's high-priority stream
    CK(cudaEventRecord(c->pipes[c->cur].sky_done, c->stream));
    c->stream = c->pipes[c->cur].chain_stream;
    CK(cudaStreamWaitEvent(c->stream, c->pipes[c->cur].sky_done, 0));
    return OSS_OK;
}

// calibrate (optional) + full star detection for n frames sitting in fp; results in ws slots 0..n-1
static int run_detect(oss_ctx *c, const FramePtrs &fp, int n, bool use_masters, int threshold, int stage_set)
{
    int r = detect_calibrate(c, fp, n, use_masters, stage_set);
    if (!r) r = detect_lowres(c, n);
    if (!r) r = detect_sky(c, n);
    return r ? r : run_detect_tail(c, n, threshold);
}

// threshold -> candidates -> components -> stars, from the sky-subtracted planes already in the work space
static int run_detect_tail(oss_ctx *c, int n, int threshold)
{
    Workspace &w = c->ws;
    const Geom &g = c->g;
    int r;
    LAUNCH("finalize_stats", k_finalize_stats, n, 256, 0, w.partials, w.stridePart, c->sky_blocks, w.ds, threshold, g);
    const long long nsegs = w.strideSeg;
    const int segblocks = (int)((nsegs + 255) / 256);
    LAUNCH("segment_masks", k_segment_masks, dim3(segblocks, n), 256, 0, w.stars, w.strideP, w.segmax, w.strideSeg, w.segmask,
           w.segoff, w.ds, g);
    r = run_scan(c, w.segoff, nsegs + 1, w.segoff, nsegs + 1, nullptr, 0, (int)nsegs, n, nullptr, 0, 1);
    if (r) return r;
    LAUNCH("compact_candidates", k_compact_candidates, dim3(segblocks, n), 256, 0, w.stars, w.strideP, w.segmask, w.strideSeg,
           w.segoff, w.cand_idx, w.cand_val, w.owner, w.ds, g);
    const int DS = (int)(sizeof(DetectState) / sizeof(int));
    const int gs = kSparseGrid;
    LAUNCH("links", k_links, dim3(gs, n), 256, 0, w.cand_idx, w.segmask, w.strideSeg, w.segoff, w.nbr, w.parent, w.comp_size,
           w.comp_peak, w.ds, g);
    LAUNCH("ccl_union", k_ccl_union, dim3(gs, n), 256, 0, w.nbr, w.parent, w.ds, g);
    LAUNCH("ccl_flatten", k_ccl_flatten, dim3(gs, n), 256, 0, w.parent, w.cand_val, w.comp_size, w.comp_peak, w.ds, g);
    // flag = kept roots, comp_off (pre-scan) = their sizes
    LAUNCH("root_flags", k_root_flags, dim3(gs, n), 256, 0, w.parent, w.comp_size, w.comp_peak, w.flag, w.comp_staroff, w.ds, g);
    int *ncand = &w.ds[0].ncand, *ncomp = &w.ds[0].ncomp, *nstars = &w.ds[0].nstars;
    r = run_scan(c, w.flag, g.cap, w.flagscan, g.cap, ncand, DS, g.cap, n, ncomp, DS);
    if (r) return r;
    r = run_scan(c, w.comp_staroff, g.cap, w.comp_staroff, g.cap, ncand, DS, g.cap, n, nullptr, 0);
    if (r) return r;
    LAUNCH("gather_roots", k_gather_roots, dim3(gs, n), 256, 0, w.flag, w.flagscan, w.comp_staroff, w.comp_size, w.comp_root, w.comp_off, w.comp_nstar,
           w.scratch, w.ds, g);
    LAUNCH("comp_members", k_comp_members, dim3(gs, n), 256, 0, w.parent, w.flag, w.flagscan, w.comp_off, w.comp_nstar, w.scratch, w.ds, g);
    // One warp per component up to 64 px (all components walked, larger ones skipped) on this stream; beside it, on the pipeline's
    // side stream, the size classes k_gather_roots listed: 65..256 px (one warp each), ..2048 px and larger (one CTA each, the
    // last one incl. the >= 10000-px bypass).  The big grids are what a star-heavy batch needs (61 MP at threshold 5: ~50 components
    // above 256 px per frame, ~0.1 ms each on its CTA's thread 0); CTAs without work leave at once, but each of the last instance's
    // needs an SM to itself first (210 KB), hence only 16.
    {
        Pipe &pp = c->pipes[c->cur];
        cudaStream_t main = c->stream;
        CK(cudaEventRecord(pp.side_fork, main));
        LAUNCH("components", (k_components<CW_CAP_SMALL, CW_WARPS_SMALL, false>), dim3(kSparseGrid, n), 32 * CW_WARPS_SMALL, kCompSmemBytesSmall,
               w.cand_idx, w.cand_val, w.nbr, w.owner, w.comp_size, w.comp_root, w.comp_off, w.comp_nstar, w.scratch, w.star_tmp, w.ds, g);
        c->stream = pp.side_stream;
        CK(cudaStreamWaitEvent(c->stream, pp.side_fork, 0));
        LAUNCH("components_mid", (k_components<CW_CAP, CW_WARPS, true>), dim3(kSparseGrid, n), 32 * CW_WARPS, kCompSmemBytes, w.cand_idx, w.cand_val,
               w.nbr, w.owner, w.comp_size, w.comp_root, w.comp_off, w.comp_nstar, w.scratch, w.star_tmp, w.ds, g);
        LAUNCH("components_big", k_components_big<CBIG_SMALL>, 296, CBIG_NT, sizeof(CompBigSmem<CBIG_SMALL>), w.cand_idx, w.cand_val, w.nbr,
               w.comp_size, w.comp_root, w.comp_off, w.comp_nstar, w.scratch, w.star_tmp, w.owner, w.ds, n, g);
        LAUNCH("components_big", k_components_big<OSS_DEBLEND_MAX - 1>, 16, CBIG_NT, sizeof(CompBigSmem<OSS_DEBLEND_MAX - 1>), w.cand_idx,
               w.cand_val, w.nbr, w.comp_size, w.comp_root, w.comp_off, w.comp_nstar, w.scratch, w.star_tmp, w.owner, w.ds, n, g);
        CK(cudaEventRecord(pp.side_join, c->stream));
        c->stream = main;
        CK(cudaStreamWaitEvent(c->stream, pp.side_join, 0));
    }
    r = run_scan(c, w.comp_nstar, g.cap, w.comp_staroff, g.cap, ncomp, DS, g.cap, n, nstars, DS);
    if (r) return r;
    LAUNCH("gather_stars", k_gather_stars, dim3(gs, n), 256, 0, w.star_tmp, w.comp_off, w.comp_nstar, w.comp_staroff, w.star_out,
           w.ds, g);
    return OSS_OK;
}

static void fill_info(const DetectState &d, oss_detect_info *info)
{
    info->mean = d.mean; info->sigma = d.sigma; info->threshold = d.threshold; info->minpeak = d.minpeak;
    info->ncandidates = d.ncand; info->ncomponents = d.ncomp; info->nstars = d.nstars; info->overflow = d.overflow;
}

static int need_geom(oss_ctx *c)
{
    if (!c) return OSS_E_ARG;
    if (!c->has_geom) return fail(c, OSS_E_ARG, "oss_set_geometry has not been called");
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) return fail(c, OSS_E_CUDA, "cudaSetDevice failed: %s", cudaGetErrorString(e));
    return OSS_OK;
}
#define NEED_GEOM() do { int r_ = need_geom(c); if (r_) return r_; } while (0)
// single-frame / bookkeeping entry points: drain every pipeline, then work on pipeline 0
#define USE_PIPE0()                                                                            \
    do {                                                                                       \
        int r_ = need_geom(c);                                                                 \
        if (r_) return r_;                                                                     \
        cudaError_t e_ = sync_all(c);                                                          \
        if (e_ != cudaSuccess) return fail(c, OSS_E_CUDA, "stream sync failed: %s", cudaGetErrorString(e_)); \
        select_pipe(c, 0);                                                                     \
    } while (0)

// same, but master broadcasts in flight on the comm stream (oss_comm_broadcast_master) keep running: for the calls that
// only touch this rank's own bulk work (building masters, the reference)
#define USE_PIPE0_LOCAL()                                                                      \
    do {                                                                                       \
        int r_ = need_geom(c);                                                                 \
        if (r_) return r_;                                                                     \
        if (c->dirty || !c->pending_warps.empty()) {                                           \
            cudaError_t e_ = sync_all(c, false);                                               \
            if (e_ != cudaSuccess) return fail(c, OSS_E_CUDA, "stream sync failed: %s", cudaGetErrorString(e_)); \
        }                                                                                      \
        select_pipe(c, 0);                                                                     \
    } while (0)
// (not dirty: whatever is still in flight — masters being built, a finalize — sits on the bulk stream, and so does everything
// these calls queue first: stream order is enough, and the host can run ahead of the GPU instead of draining ~12 streams)

extern "C" {

// flat /= mean(channel 0) of the WHOLE master (util.cpp:686-688), on c->stream
static int normalise_flat(oss_ctx *c)
{
    const Geom &g = c->g;
    const long long count = g.P * g.C;
    float *acc = c->master[OSS_MASTER_FLAT];
    const int nblk = 592;
    LAUNCH("flat_mean", k_channel0_partial, nblk, 256, 0, acc, g.P, g.C, c->d_flat_partial);
    LAUNCH("flat_mean", k_flat_ratio, 1, 32, 0, c->d_flat_partial, nblk, g.P, c->d_flat_ratio);
    LAUNCH("flat_scale", k_scale_by, (unsigned)((count + 1023) / 1024), 256, 0, acc, count, c->d_flat_ratio, 0.f);
    return OSS_OK;
}

int oss_master_begin(oss_ctx *c, int kind, int n_total, int row0, int rows)
{
    USE_PIPE0_LOCAL();
    if (kind < 0 || kind > 3 || n_total < 1) return fail(c, OSS_E_ARG, "bad master arguments");
    const Geom &g = c->g;
    const bool whole = rows <= 0 || (row0 == 0 && rows >= g.H);
    if (whole) { row0 = 0; rows = g.H; }
    // a slice starts on a multiple of 16 rows: every sample type then starts on a 16-byte boundary whatever the width
    if (!whole && (row0 < 0 || row0 + rows > g.H || (row0 & 15))) return fail(c, OSS_E_ARG, "master slice [%d, %d) of %d rows: start must be a multiple of 16", row0, row0 + rows, g.H);
    if (!c->master[kind]) DA(c->master[kind], g.P * g.C);
    c->mb = oss_ctx::MasterBuild{};
    c->mb.kind = kind; c->mb.n_total = n_total; c->mb.row0 = row0; c->mb.rows = rows; c->mb.sliced = !whole;
    c->slice_row0[kind] = row0; c->slice_rows[kind] = whole ? 0 : rows;
    return OSS_OK;
}

int oss_master_add(oss_ctx *c, const void *const *frames, int n, int mem)
{
    NEED_GEOM();
    oss_ctx::MasterBuild &mb = c->mb;
    if (mb.kind < 0) return fail(c, OSS_E_ARG, "oss_master_begin has not been called");
    if (n < 1 || !frames || mb.added + n > mb.n_total) return fail(c, OSS_E_ARG, "master frames: %d given after %d of %d", n, mb.added, mb.n_total);
    select_pipe(c, 0);
    const Geom &g = c->g;
    const int kind = mb.kind;
    const long long off = (long long)mb.row0 * g.W * g.C, count = (long long)mb.rows * g.W * g.C; // this slice, in samples
    const size_t ssz = g.sample == OSS_U8 ? 1 : (g.sample == OSS_U16 ? 2 : 4);
    float *acc = c->master[kind] + off;
    const float *sub1 = nullptr, *sub2 = nullptr;
    if (kind != OSS_MASTER_BIAS && c->master[OSS_MASTER_BIAS]) sub1 = c->master[OSS_MASTER_BIAS] + off;     // util.cpp:593,627,662
    if (kind == OSS_MASTER_FLAT && c->master[OSS_MASTER_DARKFLAT]) sub2 = c->master[OSS_MASTER_DARKFLAT] + off; // util.cpp:663
    const float r = (float)(1.0 / (double)mb.n_total);
    const int per = c->slots;
    int last_set = -1;
    for (int b0 = 0; b0 < n; b0 += per) {
        int nb = n - b0 < per ? n - b0 : per;
        FramePtrs fp{};
        int set;
        int rc = stage_frames(c, frames + b0, nb, mem, &fp, &set, true, (size_t)off * ssz, (size_t)count * ssz);
        if (rc) return rc;
        int first = mb.added == 0, last = mb.added + nb == mb.n_total;
#define ACC(T) { auto kf = k_master_accumulate<T>; LAUNCH("master_accumulate", kf, (unsigned)((count + 256ll * Sample<T>::PX - 1) / (256ll * Sample<T>::PX)), 256, 0, fp, nb, sub1, sub2, acc, first, last, r, count); }
        if (g.sample == OSS_U16) ACC(uint16_t)
        else if (g.sample == OSS_U8) ACC(uint8_t)
        else ACC(float)
#undef ACC
        if (set >= 0) { CK(cudaEventRecord(c->raw_free[set], c->stream)); last_set = set; }
        mb.added += nb;
    }
    // host frames may be reused as soon as this call returns (the copies are waited for; the kernels keep running)
    if (last_set >= 0) CK(cudaStreamSynchronize(c->copy_stream));
    return OSS_OK;
}

int oss_master_end(oss_ctx *c)
{
    NEED_GEOM();
    oss_ctx::MasterBuild &mb = c->mb;
    if (mb.kind < 0) return fail(c, OSS_E_ARG, "oss_master_begin has not been called");
    if (mb.added != mb.n_total) return fail(c, OSS_E_ARG, "master %d: %d of %d frames were added", mb.kind, mb.added, mb.n_total);
    select_pipe(c, 0);
    // a sliced flat is normalised once all slices are in place (oss_comm_allgather_master): the mean is over the whole frame
    if (mb.kind == OSS_MASTER_FLAT && !mb.sliced) { int rc = normalise_flat(c); if (rc) return rc; }
    mb.kind = -1;
    return OSS_OK;
}

// ---- masters ---------------------------------------------------------------------------
int oss_build_master(oss_ctx *c, int kind, const void *const *frames, int n, int mem)
{
    if (!c) return OSS_E_ARG;
    if (kind < 0 || kind > 3 || n < 1 || !frames) return fail(c, OSS_E_ARG, "bad master arguments");
    int rc = oss_master_begin(c, kind, n, 0, 0);
    if (!rc) rc = oss_master_add(c, frames, n, mem);
    if (!rc) rc = oss_master_end(c);
    return rc;
}

int oss_set_master(oss_ctx *c, int kind, const float *data, int mem)
{
    USE_PIPE0();
    if (kind < 0 || kind > 3 || !data) return fail(c, OSS_E_ARG, "bad master arguments");
    const long long count = c->g.P * c->g.C;
    if (!c->master[kind]) DA(c->master[kind], count);
    CK(cudaMemcpyAsync(c->master[kind], data, count * sizeof(float),
                       mem == OSS_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return OSS_OK;
}

int oss_get_master(oss_ctx *c, int kind, float *out)
{
    USE_PIPE0();
    if (kind < 0 || kind > 3 || !out) return fail(c, OSS_E_ARG, "bad master arguments");
    if (!c->master[kind]) return fail(c, OSS_E_ARG, "master %d is absent", kind);
    return oss_memcpy_d2h(c, out, c->master[kind], (size_t)c->g.P * c->g.C * sizeof(float));
}

int oss_clear_masters(oss_ctx *c)
{
    USE_PIPE0();
    CK(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < 4; k++)
        if (c->master[k]) {
            for (size_t i = 0; i < c->allocs.size(); i++)
                if (c->allocs[i] == c->master[k]) { c->allocs.erase(c->allocs.begin() + i); break; }
            cudaFree(c->master[k]);
            c->master[k] = nullptr;
        }
    return OSS_OK;
}

// ---- stacking --------------------------------------------------------------------------
// the sum plane of a new stack is not cleared: its first warp + accumulate launch starts from zero (zero_init) and writes every
// tile; whoever reads the plane before that (no lights at all) clears it first
static int ensure_sum(oss_ctx *c)
{
    if (!c->sum_unset) return OSS_OK;
    CK(cudaMemsetAsync(c->sum, 0, (size_t)c->g.P * c->g.C * sizeof(float), c->bulk));
    c->sum_unset = false;
    return OSS_OK;
}

static int reset_stack_async(oss_ctx *c)
{
    c->sum_unset = true;
    CK(cudaMemsetAsync(c->d_valid, 0, 16, c->stream));
    c->has_ref = false;
    c->ref_in_sum = true;
    c->nlights = 0;
    c->extra_valid = 0;
    c->ref_stars.clear();
    c->ref_info = oss_detect_info{};
    c->ref_host_valid = true;
    return OSS_OK;
}

int oss_reset_stack(oss_ctx *c)
{
    USE_PIPE0_LOCAL();
    int rc = reset_stack_async(c);
    if (rc) return rc;
    CK(cudaStreamSynchronize(c->stream));
    return OSS_OK;
}

int oss_set_reference(oss_ctx *c, const void *frame, int mem, int threshold)
{
    USE_PIPE0_LOCAL();
    if (!frame) return fail(c, OSS_E_ARG, "null reference frame");
    int rc = reset_stack_async(c); // on the bulk stream, in front of everything the new stack queues
    if (rc) return rc;
    const Geom &g = c->g;
    Workspace &w = c->ws;
    FramePtrs fp{};
    int set;
    rc = stage_frames(c, &frame, 1, mem, &fp, &set);
    if (rc) return rc;
    // the reference's calibrated plane is written where the final combine reads it (imagestacker.cpp:287: workingImage = the
    // calibrated reference), not into slot 0 of the work space and copied (0.1 ms at 24 MP RGB)
    float *const ws_cal = w.cal, *const ws_gray = w.gray;
    w.cal = c->ref_cal;
    if (g.C == 1) w.gray = c->ref_cal;
    rc = run_detect(c, fp, 1, true, threshold, set);
    w.cal = ws_cal; w.gray = ws_gray;
    if (rc) return rc;
    c->slot0_is_ref = true;
    if (set >= 0) CK(cudaEventSynchronize(c->copy_done[set])); // a host frame may be reused as soon as this call returns
    int *xy_ref = c->d_xy + c->slots * 2 * OSS_NOBJS, *nxy_ref = c->d_nxy + c->slots;
    LAUNCH("pack_xy", k_pack_xy, 1, OSS_NOBJS, 0, w.star_out, w.ds, g.cap, xy_ref, nxy_ref);
    LAUNCH("build_ref_triangles", k_build_ref_triangles, 1, FOCAS_NT, 0, xy_ref, nxy_ref, c->d_ref);
    CK(cudaEventRecord(c->ref_ready, c->stream));
    // results stay on the device / go to pinned memory without blocking the caller (ref_host() pulls them on demand),
    // so the first batch of lights can start while the reference's latency-bound detection tail is still running
    CK(cudaMemcpyAsync(c->h_ref_ds, w.ds, sizeof(DetectState), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(c->d_ref_stars, w.star_out, (size_t)g.cap * sizeof(oss_star), cudaMemcpyDeviceToDevice, c->stream));
    CK(cudaEventRecord(c->ref_copied, c->stream));
    c->ref_host_valid = false;
    c->threshold = threshold;
    c->has_ref = true;
    c->ref_in_sum = true;
    if (c->npipes > 1) c->light_batches = 1; // the first batch of lights goes to another pipeline than the reference's
    return OSS_OK;
}

// host copies of the reference's detection results (and its deferred capacity error)
static int ref_host(oss_ctx *c)
{
    if (c->ref_host_valid) return OSS_OK;
    CK(cudaEventSynchronize(c->ref_copied));
    const DetectState d = *c->h_ref_ds;
    fill_info(d, &c->ref_info);
    if (d.overflow) return fail(c, OSS_E_CAPACITY, "reference frame: %d candidate pixels exceed the workspace (%d)", d.ncand, c->g.cap);
    c->ref_stars.resize(d.nstars);
    if (d.nstars) CK(cudaMemcpy(c->ref_stars.data(), c->d_ref_stars, (size_t)d.nstars * sizeof(oss_star), cudaMemcpyDeviceToHost));
    c->ref_host_valid = true;
    return OSS_OK;
}

static int ensure_report_room(oss_ctx *c, int upto)
{
    while ((int)c->h_ds_chunks.size() * kChunk < upto) {
        DetectState *a = nullptr;
        AlignOut *b = nullptr;
        CK(cudaMallocHost((void **)&a, sizeof(DetectState) * kChunk));
        CK(cudaMallocHost((void **)&b, sizeof(AlignOut) * kChunk));
        c->h_ds_chunks.push_back(a);
        c->h_al_chunks.push_back(b);
    }
    return OSS_OK;
}

} // extern "C"

// a light whose candidate pixels exceeded the work space was left out of the stack (k_focas sees 0 stars): the reference
// would have stacked it, so the stack is reported as failed instead of silently differing (streams must be drained)
static int check_light_overflow(oss_ctx *c)
{
    for (int i = 0; i < c->nlights; i++) {
        const DetectState &d = c->h_ds_chunks[i / kChunk][i % kChunk];
        if (d.overflow)
            return fail(c, OSS_E_CAPACITY, "light %d: %d candidate pixels exceed the detection work space (%d); raise it with "
                        "oss_set_candidate_capacity and run the stack again", i, d.ncand, c->g.cap);
    }
    return OSS_OK;
}

// warp + accumulate of the first nb slots of a pipeline's calibrated planes into `dst`, tile rows [t0, t1).
// which: OSS_WARP_AUTO = the TMA kernel where the pipeline has a tensor map (rows are 16-byte multiples), else the cp.async
// one; OSS_WARP_TMA / OSS_WARP_TILED force one of them (tests).  Launches on c->stream.
static int launch_warp(oss_ctx *c, Pipe &pp, int nb, int t0, int t1, float *dst, int which, int zero_init = 0)
{
    const Geom &g = c->g;
    Workspace &w = pp.ws;
    if (which == OSS_WARP_TMA && !pp.has_cal_map)
        return fail(c, OSS_E_ARG, "the TMA warp kernel needs rows of a multiple of 16 bytes (width * channels * 4)");
    const bool tma = which == OSS_WARP_TMA || (which == OSS_WARP_AUTO && pp.has_cal_map);
    const dim3 wgrid((g.W + WT_W - 1) / WT_W, t1 - t0);
    if (tma && g.C == 3) {
        LAUNCH("warp_accumulate", k_warp_accumulate_tma<3>, wgrid, 256, sizeof(WarpTmaSmem<3>) + 128, pp.cal_map, w.cal, w.stridePC, w.align,
               nb, w.warp_ab, (long long)g.W, w.warp_xy, (long long)g.H, dst, g.W, g.H, t0, zero_init);
    } else if (tma) {
        LAUNCH("warp_accumulate", k_warp_accumulate_tma<1>, wgrid, 256, sizeof(WarpTmaSmem<1>) + 128, pp.cal_map, w.cal, w.stridePC, w.align,
               nb, w.warp_ab, (long long)g.W, w.warp_xy, (long long)g.H, dst, g.W, g.H, t0, zero_init);
    } else if (g.C == 3) {
        LAUNCH("warp_accumulate", k_warp_accumulate_tiled<3>, wgrid, 256, warp_tiled_smem<3>(), w.cal, w.stridePC, w.align, nb,
               w.warp_ab, (long long)g.W, w.warp_xy, (long long)g.H, dst, g.W, g.H, t0, zero_init);
    } else {
        LAUNCH("warp_accumulate", k_warp_accumulate_tiled<1>, wgrid, 256, warp_tiled_smem<1>(), w.cal, w.stridePC, w.align, nb,
               w.warp_ab, (long long)g.W, w.warp_xy, (long long)g.H, dst, g.W, g.H, t0, zero_init);
    }
    return OSS_OK;
}

// queue the warp + accumulate kernels of the oldest pending batches until at most `keep` remain.
// bands > 1 (multi-GPU, the very last batch of a stack): that batch's warp is launched as `bands` kernels over horizontal
// bands of tile rows, and after each one `after_band(c, first_row, rows)` runs (it queues the NCCL reduce of those rows of
// the sum on the comm stream), so that the reduce of band k overlaps the warp of band k + 1.
static int flush_warps(oss_ctx *c, size_t keep, int bands = 1, int (*after_band)(oss_ctx *, int, int) = nullptr)
{
    const Geom &g = c->g;
    while (c->pending_warps.size() > keep) {
        const oss_ctx::PendingWarp pw = c->pending_warps.front();
        c->pending_warps.erase(c->pending_warps.begin());
        const bool last = c->pending_warps.empty();
        Pipe &pp = c->pipes[pw.pipe];
        Workspace &w = pp.ws;
        cudaStream_t saved = c->stream;
        c->stream = c->bulk;
        CK(cudaStreamWaitEvent(c->stream, pp.detect_done, 0));
        LAUNCH("warp_tables", k_warp_tables, dim3((g.W + g.H + 255) / 256, pw.nb), 256, 0, w.align, w.warp_ab, (long long)g.W, w.warp_xy,
               (long long)g.H, g.W, g.H);
        const int tiles_y = (g.H + WT_H - 1) / WT_H, nb = (last && bands > 1) ? bands : 1;
        for (int b = 0; b < nb; b++) {
            const int t0 = (int)((long long)tiles_y * b / nb), t1 = (int)((long long)tiles_y * (b + 1) / nb);
            if (t1 > t0) { int rc = launch_warp(c, pp, pw.nb, t0, t1, c->sum, OSS_WARP_AUTO, c->sum_unset ? 1 : 0); if (rc) return rc; }
            if (nb > 1 && after_band) {
                const int r0 = t0 * WT_H, r1 = t1 * WT_H < g.H ? t1 * WT_H : g.H;
                int rc = after_band(c, r0, r1 - r0);
                if (rc) return rc;
            }
        }
        c->sum_unset = false; // every tile of the plane has been written by this batch's launch(es)
        CK(cudaEventRecord(pp.warp_done, c->stream));
        c->stream = saved;
    }
    return OSS_OK;
}

extern "C" {

int oss_add_lights(oss_ctx *c, const void *const *frames, int n, int mem)
{
    NEED_GEOM();
    if (!c->has_ref) return fail(c, OSS_E_ARG, "oss_set_reference (or oss_comm_broadcast_reference) must come first");
    if (n < 0 || (n && !frames)) return fail(c, OSS_E_ARG, "bad light list");
    const Geom &g = c->g;
    int rc = ensure_report_room(c, c->nlights + n);
    if (rc) return rc;
    // Take the next pipeline for the frames [b0, b0 + nb): wait until its previous batch has been warped, stage the frames,
    // calibrate them (bulk stream) and queue the low-res grid behind that (chain stream).
    struct LightBatch { int pipe = -1, nb = 0; };
    bool host_frames = false;
    auto start_batch = [&](int b0, int nb, LightBatch *out) -> int {
        select_pipe(c, (int)(c->light_batches++ % c->npipes));
        for (bool busy = true; busy;) { // its pending warp has to be queued before the calibrated planes are overwritten
            busy = false;
            for (const auto &pw : c->pending_warps) busy |= pw.pipe == c->cur;
            if (busy) { int r = flush_warps(c, c->pending_warps.size() - 1, 1, nullptr); if (r) return r; }
        }
        CK(cudaStreamWaitEvent(c->stream, c->pipes[c->cur].warp_done, 0));
        if (c->cur == 0) CK(cudaStreamWaitEvent(c->stream, c->ref_copied, 0)); // ... and the reference's detection tail has left pipeline 0
        FramePtrs fp{};
        int set;
        int r = stage_frames(c, frames + b0, nb, mem, &fp, &set);
        if (!r && set >= 0) {
            if (!c->ticket_stream) CK(cudaStreamCreateWithFlags(&c->ticket_stream, cudaStreamNonBlocking));
            CK(cudaStreamWaitEvent(c->ticket_stream, c->copy_done[set], 0));
            host_frames = true;
        }
        if (!r) r = detect_calibrate(c, fp, nb, true, set);
        if (!r) r = detect_lowres(c, nb);
        out->pipe = c->cur;
        out->nb = nb;
        return r;
    };
    // Bulk stream order within one call: cal(0) cal(1) sky(0) cal(2) sky(1) ... — the calibration runs one batch ahead, so
    // that the small resize / median kernels of a batch (chain stream) overlap the next batch's calibration instead of
    // leaving the GPU idle between calibrate and sky.
    LightBatch cur, nxt;
    for (int b0 = 0; b0 < n; b0 += c->slots) {
        const int nb = n - b0 < c->slots ? n - b0 : c->slots;
        if (cur.pipe < 0) { rc = start_batch(b0, nb, &cur); if (rc) return rc; }
        nxt = LightBatch{};
        const int b1 = b0 + c->slots;
        if (b1 < n && c->npipes > 2) { rc = start_batch(b1, n - b1 < c->slots ? n - b1 : c->slots, &nxt); if (rc) return rc; }
        select_pipe(c, cur.pipe);
        Workspace &w = c->ws;
        rc = detect_sky(c, nb);
        if (!rc) rc = run_detect_tail(c, nb, c->threshold);
        if (rc) return rc;
        LAUNCH("pack_xy", k_pack_xy, nb, OSS_NOBJS, 0, w.star_out, w.ds, g.cap, c->d_xy, c->d_nxy);
        CK(cudaStreamWaitEvent(c->stream, c->ref_ready, 0)); // the reference may still be in flight on pipeline 0
        LAUNCH("focas_match_fit", k_focas, nb, FOCAS_NT, sizeof(FocasSmem), c->d_ref, c->d_xy, c->d_nxy, w.align, c->d_valid,
               (int *)nullptr, (int *)nullptr, (long long *)nullptr);
        // reports back to pinned host memory, asynchronously; a batch never straddles a chunk
        // boundary more than once
        int done = 0;
        while (done < nb) {
            int gi = c->nlights + done, ch = gi / kChunk, off = gi % kChunk;
            int m = nb - done < kChunk - off ? nb - done : kChunk - off;
            CK(cudaMemcpyAsync(c->h_ds_chunks[ch] + off, w.ds + done, sizeof(DetectState) * m, cudaMemcpyDeviceToHost, c->stream));
            CK(cudaMemcpyAsync(c->h_al_chunks[ch] + off, w.align + done, sizeof(AlignOut) * m, cudaMemcpyDeviceToHost, c->stream));
            done += m;
        }
        CK(cudaEventRecord(c->pipes[c->cur].detect_done, c->stream));
        // warp + accumulate: queued in batch order, but only after the bulk kernels of the next npipes - 1
        // batches, so that this batch's latency-bound tail has run by the time the bulk stream reaches its warp
        c->pending_warps.push_back({c->cur, nb});
        rc = flush_warps(c, (size_t)(c->npipes - 1));
        if (rc) return rc;
        c->nlights += nb;
        cur = nxt; // pipe < 0: the next batch (if any) has not been started yet
    }
    if (host_frames) {
        cudaEvent_t &t = c->tickets[c->ticket_count % 16];
        if (!t) CK(cudaEventCreateWithFlags(&t, cudaEventDisableTiming));
        CK(cudaEventRecord(t, c->ticket_stream));
        c->ticket_count++;
    }
    return OSS_OK;
}

// Blocks until the OSS_MEM_HOST frames of every oss_add_lights call except the `keep_recent` most recent ones have been copied
// to the device: those host buffers may be overwritten (the decode of the next batch) while the GPU is still working on
// them and the copies of the most recent calls are still in flight.
int oss_wait_host_frames(oss_ctx *c, int keep_recent)
{
    if (!c || keep_recent < 0) return OSS_E_ARG;
    const long long idx = c->ticket_count - 1 - keep_recent;
    if (idx < 0) return OSS_OK;
    CK(cudaSetDevice(c->device));
    // (a slot of the ring that has been re-used holds a LATER ticket of the same in-order stream: waiting for it is safe)
    CK(cudaEventSynchronize(c->tickets[idx % 16]));
    return OSS_OK;
}

int oss_sync(oss_ctx *c)
{
    USE_PIPE0();
    CK(sync_all(c));
    return check_light_overflow(c);
}

int oss_num_reports(oss_ctx *c) { return c ? c->nlights : 0; }

int oss_get_frame_report(oss_ctx *c, int i, oss_frame_report *out)
{
    USE_PIPE0();
    if (i < 0 || i >= c->nlights || !out) return fail(c, OSS_E_ARG, "report index out of range");
    CK(cudaStreamSynchronize(c->stream));
    const DetectState &d = c->h_ds_chunks[i / kChunk][i % kChunk];
    const AlignOut &a = c->h_al_chunks[i / kChunk][i % kChunk];
    out->nstars = d.nstars; out->k = a.k; out->err = a.err;
    memcpy(out->M, a.M, sizeof a.M);
    out->threshold = d.threshold; out->overflow = d.overflow;
    return OSS_OK;
}

int oss_get_reference_stars(oss_ctx *c, oss_star *out, int cap, int *n, oss_detect_info *info)
{
    USE_PIPE0();
    if (!c->has_ref) return fail(c, OSS_E_ARG, "no reference");
    { int rc = ref_host(c); if (rc) return rc; }
    int m = (int)c->ref_stars.size();
    if (n) *n = m;
    if (out) memcpy(out, c->ref_stars.data(), sizeof(oss_star) * (size_t)(m < cap ? m : cap));
    if (info) *info = c->ref_info;
    return OSS_OK;
}

static int total_valid(oss_ctx *c, int *out)
{
    int v = 0;
    CK(cudaMemcpyAsync(&v, c->d_valid, sizeof v, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    *out = v + (c->ref_in_sum ? 1 : 0); // imagestacker.cpp:304: the reference counts as one
    return OSS_OK;
}

int oss_get_sum(oss_ctx *c, float *out, int *tv)
{
    USE_PIPE0();
    { int rc = ensure_sum(c); if (rc) return rc; }
    if (tv) { int rc = total_valid(c, tv); if (rc) return rc; }
    if (out) return oss_memcpy_d2h(c, out, c->sum, (size_t)c->g.P * c->g.C * sizeof(float));
    return OSS_OK;
}

int oss_finish(oss_ctx *c, float *out, int *tv)
{
    if (!c) return OSS_E_ARG;
    { int r_ = need_geom(c); if (r_) return r_; }
    if (!c->has_ref) return fail(c, OSS_E_ARG, "no reference");
    const long long PC = c->g.P * c->g.C;
    // (sum [+ reference]) * float(1 / totalValidImages), the count taken on the device.  Single GPU: queued straight behind the
    // last warp, BEFORE the host drains the streams for its checks (an error found there is returned whatever landed in the
    // result plane); with a communicator the reduce on the comm stream has to be over first.
    bool queued = false;
    if (!c->comm_stream) {
        if (flush_warps(c, 0, 1, nullptr) != OSS_OK) return OSS_E_CUDA;
        select_pipe(c, 0);
        { int rc = ensure_sum(c); if (rc) return rc; }
        LAUNCH("finalize", k_finalize, (unsigned)((PC + 1023) / 1024), 256, 0, c->sum, c->ref_in_sum ? c->ref_cal : (const float *)nullptr,
               c->result, PC, c->d_valid);
        queued = true;
    }
    USE_PIPE0();
    { int rc = ref_host(c); if (rc) return rc; } // a reference that overflowed the work space invalidates the stack
    CK(cudaStreamSynchronize(c->copy_stream));
    { int rc = check_light_overflow(c); if (rc) return rc; }
    { int rc = ensure_sum(c); if (rc) return rc; }
    if (!queued)
        LAUNCH("finalize", k_finalize, (unsigned)((PC + 1023) / 1024), 256, 0, c->sum, c->ref_in_sum ? c->ref_cal : (const float *)nullptr,
               c->result, PC, c->d_valid);
    int v;
    int rc = total_valid(c, &v);
    if (rc) return rc;
    if (tv) *tv = v;
    if (v < 2) return fail(c, OSS_E_NO_ALIGNED, "%s", oss_status_string(OSS_E_NO_ALIGNED));
    if (out) return oss_memcpy_d2h(c, out, c->result, (size_t)PC * sizeof(float));
    CK(cudaStreamSynchronize(c->stream));
    return OSS_OK;
}

const float *oss_result_device_ptr(oss_ctx *c) { return c ? c->result : nullptr; }

// ---- detect only -----------------------------------------------------------------------
int oss_detect_stars(oss_ctx *c, const void *frame, int mem, int threshold, oss_star *out, int cap, int *n, oss_detect_info *info)
{
    USE_PIPE0();
    if (!frame) return fail(c, OSS_E_ARG, "null frame");
    FramePtrs fp{};
    int set;
    int rc = stage_frames(c, &frame, 1, mem, &fp, &set);
    if (rc) return rc;
    rc = run_detect(c, fp, 1, false, threshold, set); // imagestacker.cpp:395: readImage only, no calibration
    if (rc) return rc;
    DetectState d;
    CK(cudaMemcpyAsync(&d, c->ws.ds, sizeof d, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (info) fill_info(d, info);
    if (n) *n = d.nstars;
    if (d.overflow) return fail(c, OSS_E_CAPACITY, "%d candidate pixels exceed the workspace (%d)", d.ncand, c->g.cap);
    int m = d.nstars < cap ? d.nstars : cap;
    if (out && m > 0) CK(cudaMemcpy(out, c->ws.star_out, sizeof(oss_star) * (size_t)m, cudaMemcpyDeviceToHost));
    return OSS_OK;
}

int oss_detect_stars_batch(oss_ctx *c, const void *const *frames, int n, int mem, int threshold, int32_t *counts)
{
    USE_PIPE0();
    if (n < 0 || (n && (!frames || !counts))) return fail(c, OSS_E_ARG, "bad arguments");
    std::vector<DetectState> d(c->slots);
    // host frames: the copies of batch b + 1 (other staging set) are queued before batch b is waited for, so the H2D
    // transfer of one batch overlaps the detection of the previous one
    FramePtrs fp{}, fp_next{};
    int set = -1, set_next = -1;
    if (n > 0) {
        int rc = stage_frames(c, frames, n < c->slots ? n : c->slots, mem, &fp, &set, false);
        if (rc) return rc;
    }
    for (int b0 = 0; b0 < n; b0 += c->slots) {
        const int nb = n - b0 < c->slots ? n - b0 : c->slots;
        c->stream = bulk_of(c, c->cur); // run_detect leaves the pipeline on its chain stream; the previous batch is complete
        int rc = stage_wait(c, set);
        if (!rc) rc = run_detect(c, fp, nb, false, threshold, set);
        if (rc) return rc;
        const int b1 = b0 + c->slots;
        if (b1 < n) {
            rc = stage_frames(c, frames + b1, n - b1 < c->slots ? n - b1 : c->slots, mem, &fp_next, &set_next, false);
            if (rc) return rc;
        }
        CK(cudaMemcpyAsync(d.data(), c->ws.ds, sizeof(DetectState) * nb, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        for (int i = 0; i < nb; i++) counts[b0 + i] = d[i].overflow ? -1 : d[i].nstars;
        fp = fp_next;
        set = set_next;
    }
    return OSS_OK;
}

// -s threshold sweep (BASELINE config 5): sky background, subtraction and sigma once, then
// thresholding + components + deblending per threshold.  counts[i] = stars at t_first + i.
int oss_detect_sweep(oss_ctx *c, const void *frame, int mem, int t_first, int t_last, int32_t *counts)
{
    USE_PIPE0();
    if (!frame || !counts || t_first < 1 || t_last < t_first) return fail(c, OSS_E_ARG, "bad sweep arguments");
    FramePtrs fp{};
    int set;
    int rc = stage_frames(c, &frame, 1, mem, &fp, &set);
    if (rc) return rc;
    rc = run_detect(c, fp, 1, false, t_first, set);
    if (rc) return rc;
    for (int t = t_first; t <= t_last; t++) {
        if (t > t_first) { rc = run_detect_tail(c, 1, t); if (rc) return rc; }
        DetectState d;
        CK(cudaMemcpyAsync(&d, c->ws.ds, sizeof d, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        counts[t - t_first] = d.overflow ? -1 : d.nstars;
    }
    return OSS_OK;
}

// ---- alignment only --------------------------------------------------------------------
int oss_align_lists(oss_ctx *c, const int32_t *xy1, int n1, const int32_t *xy2, int n2, int32_t *k, int32_t *err, float *M,
                    int32_t *matches_out, int32_t *votes_out)
{
    USE_PIPE0();
    if (n1 < 0 || n2 < 0 || (n1 && !xy1) || (n2 && !xy2)) return fail(c, OSS_E_ARG, "bad star lists");
    int m1 = n1 < OSS_NOBJS ? n1 : OSS_NOBJS, m2 = n2 < OSS_NOBJS ? n2 : OSS_NOBJS;
    int *xy_ref = c->d_xy + c->slots * 2 * OSS_NOBJS, *nxy_ref = c->d_nxy + c->slots;
    // the reference slot of d_xy is shared with the stack's reference list: restore it afterwards
    std::vector<int> keep(2 * OSS_NOBJS + 1);
    CK(cudaMemcpyAsync(keep.data(), xy_ref, sizeof(int) * 2 * OSS_NOBJS, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(keep.data() + 2 * OSS_NOBJS, nxy_ref, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (m1) CK(cudaMemcpyAsync(xy_ref, xy1, sizeof(int) * 2 * m1, cudaMemcpyHostToDevice, c->stream));
    if (m2) CK(cudaMemcpyAsync(c->d_xy, xy2, sizeof(int) * 2 * m2, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(nxy_ref, &m1, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d_nxy, &m2, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    LAUNCH("build_ref_triangles", k_build_ref_triangles, 1, FOCAS_NT, 0, xy_ref, nxy_ref, c->d_ref_tmp);
    LAUNCH("focas_match_fit", k_focas, 1, FOCAS_NT, sizeof(FocasSmem), c->d_ref_tmp, c->d_xy, c->d_nxy, c->ws.align, (int *)nullptr,
           c->d_dbg_matches, c->d_dbg_table, c->d_dbg_clk);
    AlignOut o;
    CK(cudaMemcpyAsync(&o, c->ws.align, sizeof o, cudaMemcpyDeviceToHost, c->stream));
    if (matches_out) CK(cudaMemcpyAsync(matches_out, c->d_dbg_matches, sizeof(int) * 3 * OSS_MAX_MATCH, cudaMemcpyDeviceToHost, c->stream));
    if (votes_out) CK(cudaMemcpyAsync(votes_out, c->d_dbg_table, sizeof(int) * OSS_NOBJS * OSS_NOBJS, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(xy_ref, keep.data(), sizeof(int) * 2 * OSS_NOBJS, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(nxy_ref, keep.data() + 2 * OSS_NOBJS, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (getenv("OSS_B200_FOCAS_CLK")) {
        long long clk[8];
        cudaMemcpy(clk, c->d_dbg_clk, sizeof clk, cudaMemcpyDeviceToHost);
        fprintf(stderr, "focas clocks: tri %lld chain %lld rebuild+sort %lld vote %lld select %lld fit %lld (nB %lld, vote bisections %lld)\n",
                clk[0], clk[1], clk[2], clk[3], clk[4], clk[5], clk[6], clk[7]);
    }
    if (k) *k = o.k;
    if (err) *err = o.err;
    if (M) memcpy(M, o.M, sizeof o.M);
    return OSS_OK;
}

// ---- warp only -------------------------------------------------------------------------
// n frames, each through its own matrix, summed in frame order onto a zeroed accumulator (0.f + v == v): the stacking
// path's own coordinate tables and warp kernels, with the kernel and the band split selectable so that tests reach every
// variant.  err (may be NULL): per-frame alignment status; frames with err != 0 are skipped like util.cpp:746-747.
int oss_warp_accumulate(oss_ctx *c, const float *const *src, const float *M, const int32_t *err, int n, float *dst, int which, int bands)
{
    USE_PIPE0();
    if (!src || !dst || !M || n < 1 || n > c->slots) return fail(c, OSS_E_ARG, "bad warp arguments (1 <= n <= frames_in_flight)");
    if (which < OSS_WARP_AUTO || which > OSS_WARP_TILED || bands < 1) return fail(c, OSS_E_ARG, "bad warp kernel selection");
    const Geom &g = c->g;
    const size_t bytes = (size_t)g.P * g.C * sizeof(float);
    Pipe &pp = c->pipes[0];
    Workspace &w = pp.ws;
    std::vector<AlignOut> ao(n);
    for (int i = 0; i < n; i++) {
        if (!src[i]) return fail(c, OSS_E_ARG, "null frame");
        ao[i] = AlignOut{};
        ao[i].err = err ? err[i] : 0;
        memcpy(ao[i].M, M + 6 * i, sizeof ao[i].M);
        CK(cudaMemcpyAsync(w.cal + (size_t)i * w.stridePC, src[i], bytes, cudaMemcpyHostToDevice, c->stream));
    }
    CK(cudaMemcpyAsync(w.align, ao.data(), sizeof(AlignOut) * n, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemsetAsync(c->result, 0, bytes, c->stream));
    LAUNCH("warp_tables", k_warp_tables, dim3((g.W + g.H + 255) / 256, n), 256, 0, w.align, w.warp_ab, (long long)g.W, w.warp_xy,
           (long long)g.H, g.W, g.H);
    const int tiles_y = (g.H + WT_H - 1) / WT_H;
    for (int b = 0; b < bands; b++) { // same band boundaries as flush_warps / oss_comm_reduce
        const int t0 = (int)((long long)tiles_y * b / bands), t1 = (int)((long long)tiles_y * (b + 1) / bands);
        if (t1 > t0) { int rc = launch_warp(c, pp, n, t0, t1, c->result, which); if (rc) return rc; }
    }
    CK(cudaMemcpyAsync(dst, c->result, bytes, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return OSS_OK;
}

int oss_warp_affine(oss_ctx *c, const float *src, float *dst, const float *M)
{
    return oss_warp_accumulate(c, &src, M, nullptr, 1, dst, OSS_WARP_AUTO, 1);
}

// ---- multi-GPU -------------------------------------------------------------------------
int oss_comm_unique_id(void *id128)
{
    std::string why;
    if (!id128 || !nccl_load(&why)) return OSS_E_CUDA;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return OSS_E_CUDA;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id128, &id, 128);
    return OSS_OK;
}

int oss_comm_init(oss_ctx *c, const void *id128, int rank, int nranks)
{
    if (!c || !id128 || rank < 0 || rank >= nranks) return fail(c, OSS_E_ARG, "bad communicator arguments");
    std::string why;
    if (!nccl_load(&why)) return fail(c, OSS_E_CUDA, "%s", why.c_str());
    CK(cudaSetDevice(c->device));
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    NK(g_nccl.CommInitRank(&c->comm, nranks, id, rank));
    c->rank = rank;
    c->nranks = nranks;
    return OSS_OK;
}

int oss_comm_broadcast_masters(oss_ctx *c, int root)
{
    USE_PIPE0();
    if (!c->comm) return fail(c, OSS_E_ARG, "oss_comm_init has not been called");
    const long long count = c->g.P * c->g.C;
    int present[4];
    for (int k = 0; k < 4; k++) present[k] = c->master[k] != nullptr;
    int *d_present = nullptr;
    { int r_ = dalloc(c, &d_present, 4); if (r_) return r_; }
    CK(cudaMemcpyAsync(d_present, present, sizeof present, cudaMemcpyHostToDevice, c->stream));
    NK(g_nccl.Broadcast(d_present, d_present, 4, ncclInt32, root, c->comm, c->stream));
    CK(cudaMemcpyAsync(present, d_present, sizeof present, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < 4; k++) {
        if (!present[k]) continue;
        if (!c->master[k]) DA(c->master[k], count);
        NK(g_nccl.Broadcast(c->master[k], c->master[k], (size_t)count, ncclFloat32, root, c->comm, c->stream));
    }
    CK(cudaStreamSynchronize(c->stream));
    return OSS_OK;
}

static int ensure_comm_stream(oss_ctx *c)
{
    if (c->comm_stream) return OSS_OK;
    int lo = 0, hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CK(cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, hi));
    CK(cudaEventCreateWithFlags(&c->comm_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->masters_ready, cudaEventDisableTiming));
    return OSS_OK;
}

// One master, asynchronously: the broadcast runs on its own stream behind whatever produced the master, so on the root
// it overlaps the building of the next master and the reference frame; calibration kernels wait for it on the device.
// Every rank calls it for the same kinds in the same order.
int oss_comm_broadcast_master(oss_ctx *c, int kind, int root)
{
    NEED_GEOM();
    if (!c->comm) return fail(c, OSS_E_ARG, "oss_comm_init has not been called");
    if (kind < 0 || kind > 3) return fail(c, OSS_E_ARG, "bad master kind");
    if (c->rank == root && !c->master[kind]) return fail(c, OSS_E_ARG, "the root has not built master %d", kind);
    const long long count = c->g.P * c->g.C;
    if (!c->master[kind]) DA(c->master[kind], count);
    { int rc = ensure_comm_stream(c); if (rc) return rc; }
    CK(cudaEventRecord(c->comm_fork, c->bulk)); // after the kernels that wrote (root) or still read (others) this master
    CK(cudaStreamWaitEvent(c->comm_stream, c->comm_fork, 0));
    // (a scatter of 1/N pieces followed by an all-gather was measured slower than NCCL's own broadcast at 4 ranks: 21.06 vs
    // 20.10 ms per step)
    NK(g_nccl.Broadcast(c->master[kind], c->master[kind], (size_t)count, ncclFloat32, root, c->comm, c->comm_stream));
    CK(cudaEventRecord(c->masters_ready, c->comm_stream));
    return OSS_OK;
}

// canonical row slice of rank r of n: starts on multiples of 16 rows
static void master_slice(int H, int r, int n, int *row0, int *rows)
{
    auto start = [&](int k) { return k <= 0 ? 0 : (k >= n ? H : (int)(((long long)H * k / n) & ~15ll)); };
    *row0 = start(r);
    *rows = start(r + 1) - start(r);
}

int oss_comm_master_slice(oss_ctx *c, int *row0, int *rows)
{
    NEED_GEOM();
    int a = 0, b = c->g.H;
    if (c->comm && c->nranks > 1) master_slice(c->g.H, c->rank, c->nranks, &a, &b);
    if (row0) *row0 = a;
    if (rows) *rows = b;
    return OSS_OK;
}

// Every rank has built ITS row slice of master `kind` (oss_master_begin with oss_comm_master_slice's rows ... oss_master_end):
// the slices are exchanged in place (one broadcast per slice, grouped) on the comm stream, beside whatever the bulk stream
// does next (the next master's slice, the reference); a flat is then normalised by the mean of its channel 0 over the
// whole frame, by every rank alike.  Per-pixel arithmetic and its order are those of the single-GPU build: the masters are
// bit-identical on every rank and to the 1-GPU ones.  Calibration kernels wait for it on the device.
int oss_comm_allgather_master(oss_ctx *c, int kind)
{
    NEED_GEOM();
    if (!c->comm) return fail(c, OSS_E_ARG, "oss_comm_init has not been called");
    if (kind < 0 || kind > 3 || !c->master[kind]) return fail(c, OSS_E_ARG, "master %d has not been built on this rank", kind);
    int row0, rows;
    master_slice(c->g.H, c->rank, c->nranks, &row0, &rows);
    if (c->nranks > 1 && (c->slice_row0[kind] != row0 || c->slice_rows[kind] != rows))
        return fail(c, OSS_E_ARG, "master %d was not built over this rank's slice [%d, %d)", kind, row0, row0 + rows);
    { int rc = ensure_comm_stream(c); if (rc) return rc; }
    CK(cudaEventRecord(c->comm_fork, c->bulk)); // behind the kernels that wrote this rank's slice
    CK(cudaStreamWaitEvent(c->comm_stream, c->comm_fork, 0));
    const size_t row_floats = (size_t)c->g.W * c->g.C;
    if (c->nranks > 1) {
        NK(g_nccl.GroupStart());
        for (int r = 0; r < c->nranks; r++) {
            int a, n;
            master_slice(c->g.H, r, c->nranks, &a, &n);
            if (n <= 0) continue;
            float *p = c->master[kind] + (size_t)a * row_floats;
            NK(g_nccl.Broadcast(p, p, (size_t)n * row_floats, ncclFloat32, r, c->comm, c->comm_stream));
        }
        NK(g_nccl.GroupEnd());
    }
    if (kind == OSS_MASTER_FLAT) {
        cudaStream_t saved = c->stream;
        c->stream = c->comm_stream;
        int rc = normalise_flat(c);
        c->stream = saved;
        if (rc) return rc;
    }
    c->slice_rows[kind] = 0; // whole again
    CK(cudaEventRecord(c->masters_ready, c->comm_stream));
    // later slices of other masters read this one's OWN slice only (bias / dark-flat rows of the same range), which the
    // exchange does not write; kernels that need the whole master wait for masters_ready
    return OSS_OK;
}

int oss_comm_broadcast_reference(oss_ctx *c, int root, int threshold)
{
    NEED_GEOM();
    if (!c->comm) return fail(c, OSS_E_ARG, "oss_comm_init has not been called");
    { int rc = ensure_comm_stream(c); if (rc) return rc; }
    const bool is_root = c->rank == root;
    if (is_root) {
        // the root does not wait on the host: the broadcasts are ordered on the device behind the reference's alignment
        // data (its detection tail may still be running), and its own lights can be queued right away
        if (!c->has_ref) return fail(c, OSS_E_ARG, "root has no reference");
        if (c->threshold != threshold) return fail(c, OSS_E_ARG, "threshold %d differs from the reference's (%d)", threshold, c->threshold);
        CK(cudaStreamWaitEvent(c->comm_stream, c->ref_ready, 0));
    } else {
        // neither do the others: they reset their stack (master exchanges in flight keep running), queue the receive and
        // go on to their lights; the alignment kernels wait for ref_ready on the device.  The threshold comes from the
        // caller (every rank has the same command line), so nothing has to be read back.
        cudaError_t e = sync_all(c, false);
        if (e != cudaSuccess) return fail(c, OSS_E_CUDA, "stream sync failed: %s", cudaGetErrorString(e));
        select_pipe(c, 0);
        int rc = oss_reset_stack(c);
        if (rc) return rc;
    }
    int *xy_ref = c->pipes[0].d_xy + c->slots * 2 * OSS_NOBJS, *nxy_ref = c->pipes[0].d_nxy + c->slots;
    NK(g_nccl.GroupStart());
    NK(g_nccl.Broadcast(c->d_ref, c->d_ref, sizeof(RefTriangles), ncclUint8, root, c->comm, c->comm_stream));
    NK(g_nccl.Broadcast(xy_ref, xy_ref, 2 * OSS_NOBJS, ncclInt32, root, c->comm, c->comm_stream));
    NK(g_nccl.Broadcast(nxy_ref, nxy_ref, 1, ncclInt32, root, c->comm, c->comm_stream));
    NK(g_nccl.GroupEnd());
    if (!is_root) {
        CK(cudaEventRecord(c->ref_ready, c->comm_stream));
        c->threshold = threshold;
        c->has_ref = true;
    }
    c->ref_in_sum = is_root; // imagestacker.cpp:287,304: the reference is counted once
    return OSS_OK;
}

// NCCL reduce of rows [r0, r0 + rows) of the sum image on the comm stream, behind what the bulk stream has queued so far
static int reduce_band(oss_ctx *c, int r0, int rows)
{
    if (rows <= 0) return OSS_OK;
    const size_t row_floats = (size_t)c->g.W * c->g.C;
    CK(cudaEventRecord(c->comm_fork, c->bulk));
    CK(cudaStreamWaitEvent(c->comm_stream, c->comm_fork, 0));
    float *p = c->sum + (size_t)r0 * row_floats;
    NK(g_nccl.Reduce(p, p, (size_t)rows * row_floats, ncclFloat32, ncclSum, c->reduce_root, c->comm, c->comm_stream));
    return OSS_OK;
}

// One reduce of the float32 sum image (+ the valid count) to `root`.  It is issued in OSS_REDUCE_BANDS row bands; when the
// last batch's warp + accumulate is still pending (oss_add_lights defers it, no oss_sync in between) that warp is launched
// band by band as well, and the reduce of a band overlaps the warp of the next one.  Every rank issues the same bands.
#define OSS_REDUCE_BANDS 4
int oss_comm_reduce(oss_ctx *c, int root)
{
    NEED_GEOM();
    if (!c->comm) return fail(c, OSS_E_ARG, "oss_comm_init has not been called");
    { int rc = ensure_comm_stream(c); if (rc) return rc; }
    c->reduce_root = root;
    const int tiles_y = (c->g.H + WT_H - 1) / WT_H;
    if (!c->pending_warps.empty()) {
        int rc = flush_warps(c, 0, OSS_REDUCE_BANDS, reduce_band);
        if (rc) return rc;
    } else {
        { int rc = ensure_sum(c); if (rc) return rc; } // a rank without any light contributes zeros
        for (int b = 0; b < OSS_REDUCE_BANDS; b++) { // same band boundaries as flush_warps
            const int t0 = (int)((long long)tiles_y * b / OSS_REDUCE_BANDS), t1 = (int)((long long)tiles_y * (b + 1) / OSS_REDUCE_BANDS);
            const int r0 = t0 * WT_H, r1 = t1 * WT_H < c->g.H ? t1 * WT_H : c->g.H;
            int rc = reduce_band(c, r0, r1 - r0);
            if (rc) return rc;
        }
    }
    // the valid count rides behind the last band: every alignment kernel (they bump d_valid) has finished before the last
    // batch's warp starts (it waits for detect_done), and the last band's reduce is ordered behind that warp — no host wait
    CK(cudaEventRecord(c->comm_fork, c->bulk));
    CK(cudaStreamWaitEvent(c->comm_stream, c->comm_fork, 0));
    for (int p = 0; p < c->npipes; p++) // (a stack without any light: the chain streams may still hold the reference)
        if (c->pipes[p].detect_done) CK(cudaStreamWaitEvent(c->comm_stream, c->pipes[p].detect_done, 0));
    NK(g_nccl.Reduce(c->d_valid, c->d_valid, 1, ncclInt32, ncclSum, root, c->comm, c->comm_stream));
    return OSS_OK; // oss_finish (root) / oss_sync (others) drain the comm stream
}

// ---- device timer + synthetic frames (bench utilities) ---------------------------------
int oss_set_pipelines(oss_ctx *c, int n)
{
    if (!c || n < 1 || n > OSS_MAX_PIPES) return fail(c, OSS_E_ARG, "pipelines must be 1..%d", OSS_MAX_PIPES);
    c->want_pipes = n; // takes effect at the next oss_set_geometry
    return OSS_OK;
}

int oss_set_sm_split(oss_ctx *c, int fp_sms)
{
    if (!c || fp_sms < 0) return fail(c, OSS_E_ARG, "SM split must be >= 0");
    c->split_fp = fp_sms; // takes effect at the next oss_set_geometry
    return OSS_OK;
}

int oss_get_sm_split(oss_ctx *c, int *fp_sms, int *hbm_sms)
{
    if (!c) return OSS_E_ARG;
    if (fp_sms) *fp_sms = c->split_fp_got;
    if (hbm_sms) *hbm_sms = c->split_hbm_got;
    return OSS_OK;
}

int oss_set_candidate_capacity(oss_ctx *c, int one_in)
{
    if (!c || one_in < 1 || one_in > 64) return fail(c, OSS_E_ARG, "candidate capacity is width*height / (1..64)");
    c->cap_one_in = one_in; // takes effect at the next oss_set_geometry
    return OSS_OK;
}

int oss_mark(oss_ctx *c, int which)
{
    if (!c || which < 0 || which > 1) return OSS_E_ARG;
    CK(cudaSetDevice(c->device));
    CK(sync_all(c));
    if (c->has_geom) select_pipe(c, 0);
    if (!c->mark[which]) CK(cudaEventCreate(&c->mark[which]));
    CK(cudaEventRecord(c->mark[which], c->stream));
    return OSS_OK;
}

int oss_elapsed_ms(oss_ctx *c, double *ms)
{
    if (!c || !ms || !c->mark[0] || !c->mark[1]) return OSS_E_ARG;
    CK(cudaEventSynchronize(c->mark[1]));
    float f = 0.f;
    CK(cudaEventElapsedTime(&f, c->mark[0], c->mark[1]));
    *ms = f;
    return OSS_OK;
}

int oss_synth_light(oss_ctx *c, void *dst_dev, const float *stars_xyas, int nstars, float background, float noise,
                    float pedestal_adu, int vignette, unsigned seed, const float *gains3)
{
    USE_PIPE0();
    const Geom &g = c->g;
    if (g.sample != OSS_U16 || !dst_dev || nstars < 0 || nstars > 65536) return fail(c, OSS_E_ARG, "synth needs a u16 geometry");
    if (!c->synth_plane) { DA(c->synth_plane, g.P); DA(c->synth_stars, 65536); }
    if (nstars) CK(cudaMemcpyAsync(c->synth_stars, stars_xyas, sizeof(float4) * nstars, cudaMemcpyHostToDevice, c->stream));
    LAUNCH("synth", k_synth_fill, 1184, 256, 0, c->synth_plane, g.P, background);
    if (nstars) LAUNCH("synth", k_synth_splat, nstars, 128, 0, c->synth_plane, g.W, g.H, c->synth_stars, nstars);
    const float g0 = gains3 ? gains3[0] : 1.f, g1 = gains3 ? gains3[1] : 1.f, g2 = gains3 ? gains3[2] : 1.f;
    LAUNCH("synth", k_synth_quantize, 1184, 256, 0, c->synth_plane, (uint16_t *)dst_dev, g.W, g.H, g.C, noise, pedestal_adu, vignette,
           seed, g0, g1, g2);
    CK(cudaStreamSynchronize(c->stream));
    return OSS_OK;
}

int oss_synth_calib(oss_ctx *c, void *dst_dev, float level_adu, float sigma_adu, int flat_profile, float pedestal_adu, unsigned seed)
{
    USE_PIPE0();
    const Geom &g = c->g;
    if (g.sample != OSS_U16 || !dst_dev) return fail(c, OSS_E_ARG, "synth needs a u16 geometry");
    LAUNCH("synth", k_synth_calib, 1184, 256, 0, (uint16_t *)dst_dev, g.W, g.H, g.C, level_adu, sigma_adu, flat_profile, pedestal_adu, seed);
    CK(cudaStreamSynchronize(c->stream));
    return OSS_OK;
}

// ---- instrumentation -------------------------------------------------------------------
int oss_profile_enable(oss_ctx *c, int on)
{
    if (!c) return OSS_E_ARG;
    cudaSetDevice(c->device);
    sync_all(c);
    prof_flush(c);
    c->profiling = on != 0;
    return OSS_OK;
}
int oss_profile_reset(oss_ctx *c)
{
    if (!c) return OSS_E_ARG;
    cudaSetDevice(c->device);
    sync_all(c);
    prof_flush(c);
    c->prof.clear();
    return OSS_OK;
}
int oss_profile_count(oss_ctx *c)
{
    if (!c) return 0;
    cudaSetDevice(c->device);
    sync_all(c);
    prof_flush(c);
    return (int)c->prof.size();
}
int oss_profile_get(oss_ctx *c, int i, const char **name, double *ms, int *launches)
{
    if (!c || i < 0 || i >= (int)c->prof.size()) return OSS_E_ARG;
    if (name) *name = c->prof[i].name.c_str();
    if (ms) *ms = c->prof[i].ms;
    if (launches) *launches = c->prof[i].launches;
    return OSS_OK;
}
long long oss_launch_count(oss_ctx *c) { return c ? c->launches : 0; }

int oss_debug_copy(oss_ctx *c, int what, int slot, float *out)
{
    USE_PIPE0();
    if (slot < 0 || slot >= c->slots || !out) return fail(c, OSS_E_ARG, "bad slot");
    const Geom &g = c->g;
    const Workspace &w = c->ws;
    const float *src = nullptr;
    size_t count = 0;
    switch (what) {
    case 0: src = (slot == 0 && c->slot0_is_ref) ? c->ref_cal : w.cal + slot * w.stridePC; count = (size_t)g.P * g.C; break;
    case 1: src = g.C == 3 ? w.gray + slot * w.strideP : ((slot == 0 && c->slot0_is_ref) ? c->ref_cal : w.cal + slot * w.stridePC); count = (size_t)g.P; break;
    case 2: src = w.stars + slot * w.strideP; count = (size_t)g.P; break;
    case 3: src = w.lo + slot * w.strideLo; count = (size_t)g.lw * g.lh; break;
    case 4: src = w.med + slot * w.strideLo; count = (size_t)g.lw * g.lh; break;
    default: return fail(c, OSS_E_ARG, "unknown plane %d", what);
    }
    return oss_memcpy_d2h(c, out, src, count * sizeof(float));
}

}  // extern "C"
