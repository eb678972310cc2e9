// ImageStackerPrivate::process / detectStars (libstacker/src/imagestacker.cpp:213-400) on the
// oss_b200 C ABI.  The per-frame loop of processConcurrent (util.cpp:724-758) becomes:
// `threads` host workers decode files into pinned buffers while the GPU works through the
// previous batch; everything numeric happens in liboss_b200.so.
#include "imagestacker.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstring>
#include <thread>

#include "../../../include/oss_b200.h"

namespace openskystacker {

class ImageStackerPrivate {
public:
    oss_ctx *ctx = nullptr;
    int device = 0;
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

// decode files[i] into dst[i] with `threads` workers, strided like util.cpp:738
bool decodeMany(const std::vector<std::string> &files, size_t first, size_t count, const ImageInfo &info, void *const *dst,
                int threads, std::string *err)
{
    std::atomic<bool> ok(true);
    std::mutex emu;
    int nt = std::max(1, std::min<int>(threads, (int)count));
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; t++)
        pool.emplace_back([&, t] {
            for (size_t k = (size_t)t; k < count && ok; k += (size_t)nt) {
                std::string e;
                if (!readImageInto(files[first + k], info, dst[k], &e)) {
                    std::lock_guard<std::mutex> lk(emu);
                    if (ok.exchange(false) && err) *err = e;
                }
            }
        });
    for (auto &th : pool) th.join();
    return ok;
}

}  // namespace

ImageStacker::ImageStacker(int device) : d_ptr(new ImageStackerPrivate)
{
    d_ptr->device = device;
}

ImageStacker::~ImageStacker()
{
    d_ptr->releasePinned();
    if (d_ptr->ctx) oss_destroy(d_ptr->ctx);
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
ACCESSOR(Image, getWorkingImage, setWorkingImage, workingImage)
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

void ImageStacker::process(int tolerance, int threads)
{
    ImageStackerPrivate *d = d_ptr.get();
    auto error = [&](const std::string &m) { if (processingError) processingError(m); };
    auto progress = [&](const std::string &m, int pct) { if (updateProgress) updateProgress(m, pct); };
    const auto started = std::chrono::steady_clock::now();

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

    if (!d->ctx) {
        int st = oss_create(&d->ctx, d->device);
        if (st) { error(oss_status_string(st)); return; }
    }
    oss_ctx *ctx = d->ctx;
    auto check = [&](int st) {
        if (st) error(st == OSS_E_NO_ALIGNED ? oss_status_string(st) : oss_last_error(ctx));
        return st == 0;
    };
    const int slots = 10;
    if (!check(oss_set_geometry(ctx, ref.width, ref.height, ref.channels, (int)ref.sample, slots))) return;
    const size_t frame_bytes = ref.bytes();
    const int hw = (int)std::max(1u, std::thread::hardware_concurrency());
    const int workers = std::max(1, std::min(threads, hw));

    // two sets of pinned staging frames: one being decoded while the other feeds the GPU
    d->releasePinned();
    void *stage[2][10];
    for (int s = 0; s < 2; s++)
        for (int i = 0; i < slots; i++) {
            stage[s][i] = oss_host_alloc(frame_bytes);
            if (!stage[s][i]) { error("out of pinned host memory"); return; }
            d->pinned.push_back(stage[s][i]);
        }

    // masters, in the reference's order (imagestacker.cpp:262-281)
    auto master = [&](bool use, const std::vector<std::string> &files, int kind, const char *message) {
        if (!use) return true;
        progress(message, 100 * d->currentOperation++ / d->totalOperations);
        if (files.empty()) { error("no frames given for an enabled calibration step"); return false; }
        // oss_build_master wants all frames of one master at once: decode them in chunks of `slots`
        // and let the library accumulate chunk by chunk is not expressible, so decode all of them.
        std::vector<void *> bufs(files.size());
        std::vector<std::vector<uint8_t>> heap(files.size());
        for (size_t i = 0; i < files.size(); i++) { heap[i].resize(frame_bytes); bufs[i] = heap[i].data(); }
        if (!decodeMany(files, 0, files.size(), ref, bufs.data(), workers, &why)) { error(why); return false; }
        return check(oss_build_master(ctx, kind, bufs.data(), (int)files.size(), OSS_MEM_HOST));
    };
    if (!check(oss_clear_masters(ctx))) return;
    if (!master(d->useBias, d->biasFrameFileNames, OSS_MASTER_BIAS, "Stacking bias frames...")) return;
    if (!master(d->useDarks, d->darkFrameFileNames, OSS_MASTER_DARK, "Stacking dark frames...")) return;
    if (!master(d->useDarkFlats, d->darkFlatFrameFileNames, OSS_MASTER_DARKFLAT, "Stacking dark flat frames...")) return;
    if (!master(d->useFlats, d->flatFrameFileNames, OSS_MASTER_FLAT, "Stacking flat frames...")) return;

    progress("Stacking light frames...", 100 * d->currentOperation++ / d->totalOperations);
    if (!readImageInto(d->refImageFileName, ref, stage[0][0], &why)) { error(why); return; }
    if (!check(oss_set_reference(ctx, stage[0][0], OSS_MEM_HOST, tolerance))) return; // imagestacker.cpp:286-287

    if (threads < 1) { error(oss_status_string(OSS_E_THREADS)); return; } // imagestacker.cpp:289-292

    const std::vector<std::string> &lights = d->targetImageFileNames;
    size_t done = 0;
    int set = 0;
    while (done < lights.size()) {
        size_t nb = std::min<size_t>(slots, lights.size() - done);
        if (!decodeMany(lights, done, nb, ref, stage[set], workers, &why)) { error(why); return; }
        // the other set may still be in flight: it was queued one round ago, wait for it before reusing it next round
        if (!check(oss_add_lights(ctx, stage[set], (int)nb, OSS_MEM_HOST))) return;
        done += nb;
        set ^= 1;
        if (done < lights.size() && !check(oss_sync(ctx))) return; // frees the set we are about to decode into
        progress("Stacking light frames...", 100 * (d->currentOperation + (int)done) / d->totalOperations);
    }

    Image out;
    out.width = ref.width; out.height = ref.height; out.channels = ref.channels;
    out.data.resize((size_t)ref.width * ref.height * ref.channels);
    int valid = 0;
    int st = oss_finish(ctx, out.data.data(), &valid);
    {
        std::lock_guard<std::mutex> lk(d->mutex);
        d->reports.clear();
        for (int i = 0; i < oss_num_reports(ctx); i++) {
            oss_frame_report r;
            if (oss_get_frame_report(ctx, i, &r)) break;
            LightReport lr{r.nstars, r.k, r.err, {}};
            memcpy(lr.M, r.M, sizeof lr.M);
            d->reports.push_back(lr);
        }
    }
    if (!check(st)) return; // OSS_E_NO_ALIGNED carries the reference's text (imagestacker.cpp:355-359)
    {
        std::lock_guard<std::mutex> lk(d->mutex);
        d->workingImage = out;
        d->finalImage = out;
    }
    int seconds = (int)std::chrono::duration_cast<std::chrono::seconds>(std::chrono::steady_clock::now() - started).count();
    if (finished) finished(out, "Stacking completed in " + getTimeString(seconds) + ".");
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
