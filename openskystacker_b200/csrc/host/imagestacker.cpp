// ImageStackerPrivate::process / detectStars (libstacker/src/imagestacker.cpp:213-400) on the
// oss_b200 C ABI.  The per-frame loop of processConcurrent (util.cpp:724-758) becomes:
// `threads` host workers decode files into pinned buffers while the GPU works through the
// previous batch; everything numeric happens in liboss_b200.so.
#include "imagestacker.hpp"

#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cmath>
#include <condition_variable>
#include <cstring>
#include <thread>

#include "../../../include/oss_b200.h"

namespace openskystacker {

class ImageStackerPrivate {
public:
    oss_ctx *ctx = nullptr;              // detect-only context (getStars / detectStars)
    int device = 0;
    std::vector<int> devices;            // GPUs process() shards the lights over (empty: `device`)
    std::vector<oss_ctx *> ctxs;         // one per entry of `devices`
    int ctx_device[64] = {};
    bool comm_ready = false;             // the contexts share an NCCL communicator
    bool dynamic_sharding = false;       // lights from one shared queue instead of light k -> GPU k % N
    std::atomic<int> lightsDone{0};
    std::string lastCapacityError;
    mutable std::mutex mutex; // the reference guards every accessor with a QMutex (imagestacker.cpp:591-786)
    std::string refImageFileName, saveFilePath;
    std::vector<std::string> targetImageFileNames, darkFrameFileNames, darkFlatFrameFileNames, flatFrameFileNames,
        biasFrameFileNames;
    Image workingImage, refImage, finalImage;
    bool useDarks = false, useDarkFlats = false, useFlats = false, useBias = false;
    int currentOperation = 0, totalOperations = 0;
    std::vector<ImageStacker::LightReport> reports;
    std::vector<void *> pinned;

    int getTotalOperations() const // imagestacker.cpp:218-228
    {
        int ops = (int)targetImageFileNames.size() + 4;
        if (useBias) ops += 1;
        if (useDarks) ops += 1;
        if (useDarkFlats) ops += 1;
        if (useFlats) ops += 1;
        return ops;
    }
    void releasePinned()
    {
        for (void *p : pinned) oss_host_free(p);
        pinned.clear();
    }
};

namespace {

// decode files[i] into dst[i] with `threads` workers, strided like util.cpp:738.  More workers than files: a file is shared by
// several of them (readImageBandInto: a TIFF's strips are independent), so that a batch of 5 large frames still uses -j 16
bool decodeMany(const std::vector<std::string> &files, size_t first, size_t count, const ImageInfo &info, void *const *dst,
                int threads, std::string *err)
{
    std::atomic<bool> ok(true);
    std::mutex emu;
    const int bands = count ? std::max(1, std::min(16, threads / (int)count)) : 1;
    const size_t tasks = count * (size_t)bands;
    int nt = std::max(1, std::min<int>(threads, (int)tasks));
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; t++)
        pool.emplace_back([&, t] {
            for (size_t k = (size_t)t; k < tasks && ok; k += (size_t)nt) {
                const size_t file = k / (size_t)bands;
                std::string e;
                if (!readImageBandInto(files[first + file], info, dst[file], (int)(k % (size_t)bands), bands, &e)) {
                    std::lock_guard<std::mutex> lk(emu);
                    if (ok.exchange(false) && err) *err = e;
                }
            }
        });
    for (auto &th : pool) th.join();
    return ok;
}

}  // namespace

bool saveImageLikeGui(const std::string &path, const Image &image, std::string *err)
{
    if (image.channels != 1) return writeTiffF32(path, image.data.data(), image.width, image.height, image.channels, err);
    std::vector<uint16_t> px(image.data.size());
    for (size_t i = 0; i < px.size(); i++) {
        // cv::Mat::convertTo(CV_16U, 65535) on CV_32F: saturate_cast<ushort>(v * 65535.f) = cvRound (round half to even), clamped
        const float v = image.data[i] * 65535.0f;
        long q = v != v ? 0 : (v <= 0.f ? 0 : (v >= 65535.f ? 65535 : lrintf(v)));
        px[i] = (uint16_t)q;
    }
    return writeTiffU16(path, px.data(), image.width, image.height, 1, err);
}

int visibleDeviceCount()
{
    return oss_device_count();
}

ImageStacker::ImageStacker(int device) : d_ptr(new ImageStackerPrivate)
{
    d_ptr->device = device;
}

ImageStacker::~ImageStacker()
{
    d_ptr->releasePinned();
    if (d_ptr->ctx) oss_destroy(d_ptr->ctx);
    for (oss_ctx *c : d_ptr->ctxs) if (c) oss_destroy(c);
}

void ImageStacker::setDevices(const std::vector<int> &devices)
{
    std::lock_guard<std::mutex> lk(d_ptr->mutex);
    d_ptr->devices = devices;
    if (d_ptr->devices.size() > 64) d_ptr->devices.resize(64);
}
void ImageStacker::setDynamicSharding(bool on) { std::lock_guard<std::mutex> lk(d_ptr->mutex); d_ptr->dynamic_sharding = on; }
bool ImageStacker::getDynamicSharding() const { std::lock_guard<std::mutex> lk(d_ptr->mutex); return d_ptr->dynamic_sharding; }
std::vector<int> ImageStacker::getDevices() const
{
    std::lock_guard<std::mutex> lk(d_ptr->mutex);
    return d_ptr->devices.empty() ? std::vector<int>{d_ptr->device} : d_ptr->devices;
}

#define ACCESSOR(Type, Get, Set, field)                                                                         \
    Type ImageStacker::Get() const { std::lock_guard<std::mutex> lk(d_ptr->mutex); return d_ptr->field; }      \
    void ImageStacker::Set(const Type &value) { std::lock_guard<std::mutex> lk(d_ptr->mutex); d_ptr->field = value; }
ACCESSOR(std::string, getRefImageFileName, setRefImageFileName, refImageFileName)
ACCESSOR(std::vector<std::string>, getTargetImageFileNames, setTargetImageFileNames, targetImageFileNames)
ACCESSOR(std::vector<std::string>, getDarkFrameFileNames, setDarkFrameFileNames, darkFrameFileNames)
ACCESSOR(std::vector<std::string>, getDarkFlatFrameFileNames, setDarkFlatFrameFileNames, darkFlatFrameFileNames)
ACCESSOR(std::vector<std::string>, getFlatFrameFileNames, setFlatFrameFileNames, flatFrameFileNames)
ACCESSOR(std::vector<std::string>, getBiasFrameFileNames, setBiasFrameFileNames, biasFrameFileNames)
ACCESSOR(std::string, getSaveFilePath, setSaveFilePath, saveFilePath)
Image ImageStacker::getWorkingImage() const
{
    std::lock_guard<std::mutex> lk(d_ptr->mutex);
    return d_ptr->workingImage.data.empty() ? d_ptr->finalImage : d_ptr->workingImage;
}
void ImageStacker::setWorkingImage(const Image &value) { std::lock_guard<std::mutex> lk(d_ptr->mutex); d_ptr->workingImage = value; }
ACCESSOR(Image, getRefImage, setRefImage, refImage)
ACCESSOR(Image, getFinalImage, setFinalImage, finalImage)
#undef ACCESSOR
#define FLAG(Get, Set, field)                                                                                    \
    bool ImageStacker::Get() const { std::lock_guard<std::mutex> lk(d_ptr->mutex); return d_ptr->field; }       \
    void ImageStacker::Set(bool value) { std::lock_guard<std::mutex> lk(d_ptr->mutex); d_ptr->field = value; }
FLAG(getUseDarks, setUseDarks, useDarks)
FLAG(getUseDarkFlats, setUseDarkFlats, useDarkFlats)
FLAG(getUseFlats, setUseFlats, useFlats)
FLAG(getUseBias, setUseBias, useBias)
#undef FLAG

std::vector<ImageStacker::LightReport> ImageStacker::getLightReports() const
{
    std::lock_guard<std::mutex> lk(d_ptr->mutex);
    return d_ptr->reports;
}

// ---------------------------------------------------------------------------------------------------------------------
// process(): ImageStackerPrivate::process (imagestacker.cpp:230-376) + processConcurrent (util.cpp:724-758).
//
// One host thread per GPU (`devices`, default: the constructor's device).  Light k goes to GPU k % N — the reference's own
// split over its worker threads (util.cpp:738) — and `threads` keeps its meaning as the number of DECODE workers, divided
// among the GPU threads.  Per GPU thread:
//   masters    every calibration frame is decoded ONCE (by a producer thread, two chunks of pinned buffers ahead) and each GPU
//              uploads and accumulates only its 1/N row slice of it (oss_master_begin/add/end), then the slices are exchanged
//              over NVLink (oss_comm_allgather_master): bit-identical masters, 1/N of the PCIe traffic per GPU;
//   reference  GPU 0 calibrates + detects it, its triangle / star lists are broadcast;
//   lights     two sets of pinned frames: while the GPU copies and works on one batch, the workers decode the next one into
//              the other set; the set is handed back by oss_wait_host_frames (copies done), not by a full oss_sync;
//   combine    one NCCL reduce of the float32 sums + the valid count to GPU 0, then /= N (imagestacker.cpp:348-361).
namespace {

struct Barrier { // std::barrier without C++20
    std::mutex m;
    std::condition_variable cv;
    int n, waiting = 0, phase = 0;
    explicit Barrier(int n_) : n(n_) {}
    void wait()
    {
        std::unique_lock<std::mutex> lk(m);
        const int p = phase;
        if (++waiting == n) { waiting = 0; phase++; cv.notify_all(); }
        else cv.wait(lk, [&] { return phase != p; });
    }
};

// calibration frames decoded once, consumed by every GPU thread: a ring of two chunks of pinned buffers
struct ChunkRing {
    struct Chunk { std::vector<void *> buf; int kind = -1, first = 0, n = 0, total = 0, consumers_left = 0; bool ready = false; };
    Chunk chunk[2];
    std::mutex m;
    std::condition_variable cv;
    bool failed = false;
    std::string why;
};

}  // namespace

void ImageStacker::process(int tolerance, int threads)
{
    ImageStackerPrivate *d = d_ptr.get();
    std::mutex sig; // the signals are emitted from several GPU threads
    auto error = [&](const std::string &m) { std::lock_guard<std::mutex> lk(sig); if (processingError) processingError(m); };
    auto progress = [&](const std::string &m, int pct) { std::lock_guard<std::mutex> lk(sig); if (updateProgress) updateProgress(m, pct); };
    const auto started = std::chrono::steady_clock::now();
    // dev aid: OSS_STACKER_TIMING=1 prints where the wall clock of process() goes (stderr)
    const bool timing = getenv("OSS_STACKER_TIMING") != nullptr;
    auto mark = [&](const char *what) {
        if (timing) fprintf(stderr, "[oss timing] %-28s %8.3f s\n", what, std::chrono::duration<double>(std::chrono::steady_clock::now() - started).count());
    };

    // imagestacker.cpp:236-246
    ImageType refType = getImageType(d->refImageFileName);
    for (const std::string &f : d->targetImageFileNames)
        if (getImageType(f) != refType) { error(oss_status_string(OSS_E_TYPE)); return; }
    if (refType == RAW_IMAGE) { error("RAW decoding (LibRaw) is not part of this build; convert the frames to 16-bit TIFF or FITS."); return; }

    cancel = false;
    d->currentOperation = 1;
    d->totalOperations = d->getTotalOperations();
    progress("Checking image sizes...", 100 * d->currentOperation++ / d->totalOperations);

    // validateImageSizes (imagestacker.cpp:402-557), from the file headers instead of a full decode
    ImageInfo ref;
    std::string why;
    if (!probeImage(d->refImageFileName, &ref, &why)) { error(why); return; }
    auto sameSize = [&](const std::vector<std::string> &files) {
        for (const std::string &f : files) {
            ImageInfo i;
            if (!probeImage(f, &i, &why)) { error(why); return false; }
            if (i.width != ref.width || i.height != ref.height) { error(oss_status_string(OSS_E_SIZE)); return false; }
            if (i.channels != ref.channels || i.sample != ref.sample) { error("Images must have the same channel count and bit depth."); return false; }
        }
        return true;
    };
    if (!sameSize(d->targetImageFileNames)) return;
    if (d->useBias && !sameSize(d->biasFrameFileNames)) return;
    if (d->useDarks && !sameSize(d->darkFrameFileNames)) return;
    if (d->useDarkFlats && !sameSize(d->darkFlatFrameFileNames)) return;
    if (d->useFlats && !sameSize(d->flatFrameFileNames)) return;
    mark("headers checked");

    struct MasterJob { bool use; const std::vector<std::string> *files; int kind; const char *message; };
    const MasterJob masters[4] = { // the reference's order (imagestacker.cpp:262-281)
        {d->useBias, &d->biasFrameFileNames, OSS_MASTER_BIAS, "Stacking bias frames..."},
        {d->useDarks, &d->darkFrameFileNames, OSS_MASTER_DARK, "Stacking dark frames..."},
        {d->useDarkFlats, &d->darkFlatFrameFileNames, OSS_MASTER_DARKFLAT, "Stacking dark flat frames..."},
        {d->useFlats, &d->flatFrameFileNames, OSS_MASTER_FLAT, "Stacking flat frames..."}};
    for (const MasterJob &m : masters)
        if (m.use && m.files->empty()) { error("no frames given for an enabled calibration step"); return; }

    // ---- devices, contexts
    std::vector<int> devs = d->devices.empty() ? std::vector<int>{d->device} : d->devices;
    const int N = (int)devs.size();
    while ((int)d->ctxs.size() < N) d->ctxs.push_back(nullptr);
    for (int r = 0; r < N; r++)
        if (!d->ctxs[r] || d->ctx_device[r] != devs[r]) {
            if (d->ctxs[r]) oss_destroy(d->ctxs[r]);
            d->ctxs[r] = nullptr;
            int st = oss_create(&d->ctxs[r], devs[r]);
            if (st) { error(oss_status_string(st)); return; }
            d->ctx_device[r] = devs[r];
            d->comm_ready = false;
        }
    mark("contexts");
    const size_t frame_bytes = ref.bytes();
    // frames in flight per pipeline: about 0.75 GB of raw frames per batch (24 MP RGB16: 5, the host-fed optimum measured by
    // bench.py), at least 2, at most 16; 3 pipelines
    const int slots = (int)std::max<size_t>(2, std::min<size_t>(16, (768u << 20) / std::max<size_t>(frame_bytes, 1)));
    const int hw = (int)std::max(1u, std::thread::hardware_concurrency());
    const int workers = std::max(1, std::min(threads, hw));
    const int workers_per_dev = std::max(1, workers / N);
    const std::vector<std::string> &lights = d->targetImageFileNames;

    // ---- one attempt at the whole stack with a given detection work space; OSS_E_CAPACITY asks for a larger one
    const bool dynamic = d->dynamic_sharding && N > 1;
    std::vector<std::vector<size_t>> takenIdx(N); // list positions of the lights each GPU processed, in its own order
    auto attempt = [&](int capacity_one_in, Image *out, int *valid_out, std::vector<std::vector<LightReport>> *reps) -> int {
        std::vector<std::vector<size_t>> *taken = &takenIdx;
        std::atomic<size_t> next_light{0};
        std::atomic<int> status(0);
        std::string first_error;
        std::mutex emu;
        auto fail = [&](int st, const std::string &m) {
            std::lock_guard<std::mutex> lk(emu);
            int expect = 0;
            if (status.compare_exchange_strong(expect, st)) first_error = m;
        };
        Barrier bar(N);
        unsigned char nccl_id[128];
        if (N > 1 && !d->comm_ready && oss_comm_unique_id(nccl_id)) { error("NCCL is not available: multi-GPU stacking needs it"); return OSS_E_CUDA; }

        // pinned staging: per GPU two sets of `slots` frames for the lights; shared: two chunks of `slots` calibration frames
        d->releasePinned();
        auto pinned = [&](size_t bytes) { void *p = oss_host_alloc(bytes); if (p) d->pinned.push_back(p); return p; };
        std::vector<std::array<std::vector<void *>, 2>> stage(N);
        ChunkRing ring;
        bool any_master = false;
        for (const MasterJob &m : masters) any_master |= m.use;
        // (pinning memory costs ~0.5 s per GB: the calibration chunks, idle once the masters are built — every GPU has passed the
        // barrier behind the last master by then —, double as GPU 0's two sets for the lights)
        if (any_master)
            for (int s2 = 0; s2 < 2; s2++)
                for (int i = 0; i < slots; i++) {
                    void *p = pinned(frame_bytes);
                    if (!p) { error("out of pinned host memory"); return OSS_E_NOMEM; }
                    ring.chunk[s2].buf.push_back(p);
                }
        for (int r = 0; r < N; r++)
            for (int s2 = 0; s2 < 2; s2++) {
                if (r == 0 && any_master) { stage[r][s2] = ring.chunk[s2].buf; continue; }
                for (int i = 0; i < slots; i++) {
                    void *p = pinned(frame_bytes);
                    if (!p) { error("out of pinned host memory"); return OSS_E_NOMEM; }
                    stage[r][s2].push_back(p);
                }
            }
        mark("pinned staging allocated");
        // producer: decodes the calibration frames of every enabled master, chunk by chunk, two chunks ahead of the GPUs
        std::thread producer([&] {
            int slot = 0;
            for (const MasterJob &m : masters) {
                if (!m.use) continue;
                const int total = (int)m.files->size();
                for (int first = 0; first < total; first += slots) {
                    ChunkRing::Chunk &c = ring.chunk[slot];
                    {
                        std::unique_lock<std::mutex> lk(ring.m);
                        ring.cv.wait(lk, [&] { return (!c.ready && c.consumers_left == 0) || ring.failed; });
                        if (ring.failed) return;
                    }
                    const int n = std::min(slots, total - first);
                    std::string e;
                    const bool ok = decodeMany(*m.files, (size_t)first, (size_t)n, ref, c.buf.data(), workers, &e);
                    {
                        std::lock_guard<std::mutex> lk(ring.m);
                        if (!ok) { ring.failed = true; ring.why = e; }
                        c.kind = m.kind; c.first = first; c.n = n; c.total = total; c.consumers_left = N; c.ready = true;
                    }
                    ring.cv.notify_all();
                    if (!ok) return;
                    slot ^= 1;
                }
            }
        });

        auto gpu_thread = [&](int r) {
            oss_ctx *ctx = d->ctxs[r];
            auto ck = [&](int st) { if (st) fail(st, st == OSS_E_NO_ALIGNED ? oss_status_string(st) : oss_last_error(ctx)); return st == 0; };
            // every rank runs the same sequence of collectives: a rank that fails keeps walking through the barriers and
            // skips the work; collectives are only entered when nobody has failed at the preceding barrier
            bool ok = ck(oss_set_candidate_capacity(ctx, capacity_one_in)) &&
                      ck(oss_set_geometry(ctx, ref.width, ref.height, ref.channels, (int)ref.sample, slots)) && ck(oss_clear_masters(ctx));
            if (N > 1 && !d->comm_ready && ok) ok = ck(oss_comm_init(ctx, nccl_id, r, N));
            if (r == 0) mark("geometry (device work space)");
            bar.wait();
            if (status) return;
            int row0 = 0, rows = 0;
            if (N > 1) ok = ck(oss_comm_master_slice(ctx, &row0, &rows));
            // ---- masters
            int slot = 0;
            for (const MasterJob &m : masters) {
                if (!m.use) continue;
                if (r == 0) progress(m.message, 100 * d->currentOperation++ / d->totalOperations);
                const int total = (int)m.files->size();
                if (ok) ok = ck(oss_master_begin(ctx, m.kind, total, row0, N > 1 ? rows : 0));
                for (int first = 0; first < total; first += slots) {
                    ChunkRing::Chunk &c = ring.chunk[slot];
                    {
                        std::unique_lock<std::mutex> lk(ring.m);
                        ring.cv.wait(lk, [&] { return (c.ready && c.kind == m.kind && c.first == first) || ring.failed; });
                        if (ring.failed) { fail(OSS_E_ARG, ring.why); ok = false; }
                    }
                    if (ok && !status) ok = ck(oss_master_add(ctx, c.buf.data(), c.n, OSS_MEM_HOST)); // returns when this GPU's slice is copied
                    {
                        std::lock_guard<std::mutex> lk(ring.m);
                        if (c.ready && --c.consumers_left == 0) c.ready = false;
                    }
                    ring.cv.notify_all();
                    slot ^= 1;
                    if (ring.failed) break;
                }
                if (ok && !status) ok = ck(oss_master_end(ctx));
                bar.wait(); // every GPU has its slice (or someone failed): enter the exchange together
                if (status) return;
                if (N > 1) ok = ck(oss_comm_allgather_master(ctx, m.kind));
            }
            // ---- reference (imagestacker.cpp:286-287)
            if (r == 0) mark("masters queued");
            if (r == 0) {
                progress("Stacking light frames...", 100 * d->currentOperation++ / d->totalOperations);
                std::string e;
                if (!readImageInto(d->refImageFileName, ref, stage[0][0][0], &e)) { fail(OSS_E_ARG, e); ok = false; }
                if (ok) ok = ck(oss_set_reference(ctx, stage[0][0][0], OSS_MEM_HOST, tolerance));
            }
            bar.wait();
            if (status) return;
            if (N > 1) ok = ck(oss_comm_broadcast_reference(ctx, 0, tolerance));
            // ---- lights r, r + N, r + 2N, ... (util.cpp:738) or, with dynamic sharding, the next `slots` lights of the shared
            // queue whenever this GPU is ready for more; two pinned sets: decode(b + 1) overlaps copy + compute of b
            std::vector<std::string> mine;
            std::vector<size_t> &mineIdx = (*taken)[r];
            mineIdx.clear();
            if (!dynamic)
                for (size_t k = (size_t)r; k < lights.size(); k += (size_t)N) { mine.push_back(lights[k]); mineIdx.push_back(k); }
            size_t done = 0;
            int set = 0;
            while (ok && !status) {
                if (dynamic) {
                    const size_t first = next_light.fetch_add((size_t)slots);
                    for (size_t k = first; k < std::min(lights.size(), first + (size_t)slots); k++) { mine.push_back(lights[k]); mineIdx.push_back(k); }
                }
                if (done >= mine.size()) break;
                const size_t nb = std::min<size_t>((size_t)slots, mine.size() - done);
                std::string e;
                if (!decodeMany(mine, done, nb, ref, stage[r][set].data(), workers_per_dev, &e)) { fail(OSS_E_ARG, e); ok = false; break; }
                ok = ck(oss_add_lights(ctx, stage[r][set].data(), (int)nb, OSS_MEM_HOST));
                done += nb;
                set ^= 1;
                // the set we decode into next was handed to the call before this one: wait for ITS copies only
                if (ok) ok = ck(oss_wait_host_frames(ctx, 1));
                d->lightsDone += (int)nb;
                progress("Stacking light frames...", 100 * (d->currentOperation + d->lightsDone.load()) / d->totalOperations);
            }
            if (r == 0) mark("lights queued (this GPU)");
            bar.wait();
            if (status) return;
            if (N > 1) ok = ck(oss_comm_reduce(ctx, 0));
            if (r == 0) {
                out->width = ref.width; out->height = ref.height; out->channels = ref.channels;
                out->data.resize((size_t)ref.width * ref.height * ref.channels);
                if (ok) ck(oss_finish(ctx, out->data.data(), valid_out));
                mark("finished (result on the host)");
            } else if (ok) ck(oss_sync(ctx));
            // per-light outcomes of this GPU (also after a failed finish: they say why)
            std::vector<LightReport> &mineRep = (*reps)[r];
            mineRep.clear();
            const int nrep = oss_num_reports(ctx);
            for (int i = 0; i < nrep; i++) {
                oss_frame_report fr;
                if (oss_get_frame_report(ctx, i, &fr)) break;
                LightReport lr{fr.nstars, fr.k, fr.err, {}, fr.overflow};
                memcpy(lr.M, fr.M, sizeof lr.M);
                mineRep.push_back(lr);
            }
        };

        d->lightsDone = 0;
        std::vector<std::thread> pool;
        for (int r = 1; r < N; r++) pool.emplace_back(gpu_thread, r);
        gpu_thread(0);
        for (auto &t : pool) t.join();
        { std::lock_guard<std::mutex> lk(ring.m); ring.failed = ring.failed || status != 0; }
        ring.cv.notify_all();
        producer.join();
        if (N > 1 && !status) d->comm_ready = true;
        if (status) {
            if (N > 1) { // a failed multi-GPU attempt may leave collectives half queued: start from fresh contexts next time
                for (int r = 0; r < N; r++) { oss_destroy(d->ctxs[r]); d->ctxs[r] = nullptr; }
                d->comm_ready = false;
            }
            if (status != OSS_E_CAPACITY) error(first_error);
            else d->lastCapacityError = first_error;
        }
        return status;
    };

    if (threads < 1) { error(oss_status_string(OSS_E_THREADS)); return; } // imagestacker.cpp:289-292

    Image out;
    int valid = 0;
    std::vector<std::vector<LightReport>> reps(N);
    // threshold 1 is 1.5 sigma: 6.7 % of the pixels of a sparse field are candidates, above the default work space of 1/16
    int capacity = tolerance <= 1 ? 4 : 16;
    if (const char *e = getenv("OSS_STACKER_CAPACITY")) { int v = atoi(e); if (v >= 1 && v <= 64) capacity = v; } // test hook: start smaller
    int st = attempt(capacity, &out, &valid, &reps);
    while (st == OSS_E_CAPACITY && capacity > 1) { // the reference has no such limit: enlarge the work space and run again
        capacity = capacity > 4 ? 4 : 1;
        for (int r = 0; r < N; r++)
            if (!d->ctxs[r]) { int c2 = oss_create(&d->ctxs[r], devs[r]); if (c2) { error(oss_status_string(c2)); return; } }
        d->currentOperation = 2;
        st = attempt(capacity, &out, &valid, &reps);
    }
    if (st == OSS_E_CAPACITY) error(d->lastCapacityError);
    {
        std::lock_guard<std::mutex> lk(d->mutex);
        d->reports.assign(lights.size(), LightReport{0, 0, -1, {}, 0});
        for (int r = 0; r < N; r++)
            for (size_t i = 0; i < reps[r].size(); i++)
                if (i < takenIdx[r].size() && takenIdx[r][i] < lights.size()) d->reports[takenIdx[r][i]] = reps[r][i]; // back to list order
    }
    d->releasePinned();
    mark("pinned staging released");
    if (st) return; // OSS_E_NO_ALIGNED carries the reference's text (imagestacker.cpp:355-359)
    {
        std::lock_guard<std::mutex> lk(d->mutex);
        d->finalImage = out;          // one copy (288 MB at 24 MP RGB): getWorkingImage() hands out the final image as long as
        d->workingImage = Image();    // nobody has set another working image (the reference's two cv::Mat share their pixels)
    }
    int seconds = (int)std::chrono::duration_cast<std::chrono::seconds>(std::chrono::steady_clock::now() - started).count();
    mark("images kept");
    if (finished) finished(out, "Stacking completed in " + getTimeString(seconds) + ".");
    mark("finished() returned");
}

std::vector<Star> ImageStacker::getStars(const std::string &filename, int threshold)
{
    ImageStackerPrivate *d = d_ptr.get();
    std::vector<Star> out;
    auto error = [&](const std::string &m) { if (processingError) processingError(m); };
    ImageInfo info;
    std::vector<uint8_t> px;
    std::string why;
    if (!readImage(filename, &info, &px, &why)) { error(why); return out; }
    if (!d->ctx) {
        int st = oss_create(&d->ctx, d->device);
        if (st) { error(oss_status_string(st)); return out; }
    }
    if (oss_set_geometry(d->ctx, info.width, info.height, info.channels, (int)info.sample, 1)) { error(oss_last_error(d->ctx)); return out; }
    int n = 0;
    if (oss_detect_stars(d->ctx, px.data(), OSS_MEM_HOST, threshold, nullptr, 0, &n, nullptr)) { error(oss_last_error(d->ctx)); return out; }
    std::vector<oss_star> tmp((size_t)n);
    if (n && oss_detect_stars(d->ctx, px.data(), OSS_MEM_HOST, threshold, tmp.data(), n, &n, nullptr)) { error(oss_last_error(d->ctx)); return out; }
    out.resize(tmp.size());
    for (size_t i = 0; i < tmp.size(); i++) { out[i].x = tmp[i].x; out[i].y = tmp[i].y; out[i].area = tmp[i].area; out[i].peak = tmp[i].peak; out[i].value = tmp[i].value; }
    return out;
}

// imagestacker.cpp:392-400: readImage (no calibration) -> getStars -> count
void ImageStacker::detectStars(const std::string &filename, int threshold)
{
    bool failed = false;
    auto saved = processingError;
    processingError = [&](const std::string &m) { failed = true; if (saved) saved(m); };
    std::vector<Star> list = getStars(filename, threshold);
    processingError = saved;
    if (!failed && doneDetectingStars) doneDetectingStars((int)list.size());
}

}  // namespace openskystacker
