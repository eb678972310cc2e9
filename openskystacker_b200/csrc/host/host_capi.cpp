// Plain-C doors into the host layer (image codecs + image list) so that the Python tests can
// exercise them with ctypes.  Not part of include/oss_b200.h: nothing numeric happens here.
#include <cstring>
#include <string>

#include "image_io.hpp"
#include "imagelist.hpp"
#include "imagestacker.hpp"

using namespace openskystacker;

extern "C" {

// whcs = {width, height, channels, sample}; returns 0 on success
__attribute__((visibility("default"))) int oss_host_probe_image(const char *path, int *whcs, char *err, int errcap)
{
    ImageInfo i;
    std::string e;
    if (!probeImage(path, &i, &e)) { if (err) { strncpy(err, e.c_str(), errcap - 1); err[errcap - 1] = 0; } return -1; }
    whcs[0] = i.width; whcs[1] = i.height; whcs[2] = i.channels; whcs[3] = (int)i.sample;
    return 0;
}

__attribute__((visibility("default"))) int oss_host_read_image(const char *path, const int *whcs, void *dst, char *err, int errcap)
{
    ImageInfo i;
    i.width = whcs[0]; i.height = whcs[1]; i.channels = whcs[2]; i.sample = (SampleKind)whcs[3];
    std::string e;
    if (!readImageInto(path, i, dst, &e)) { if (err) { strncpy(err, e.c_str(), errcap - 1); err[errcap - 1] = 0; } return -1; }
    return 0;
}

__attribute__((visibility("default"))) int oss_host_write_tiff_f32(const char *path, const float *px, int w, int h, int c)
{
    std::string e;
    return writeTiffF32(path, px, w, h, c, &e) ? 0 : -1;
}

// records as lines "type reference checked width height filename\n"; returns the loadImageList error code
__attribute__((visibility("default"))) int oss_host_load_image_list(const char *path, char *out, int cap)
{
    int err = 0;
    std::vector<ImageRecord *> recs = loadImageList(path, &err);
    std::string text;
    for (ImageRecord *r : recs) {
        text += std::to_string((int)r->type) + " " + std::to_string((int)r->reference) + " " + std::to_string((int)r->checked) + " " +
                std::to_string(r->width) + " " + std::to_string(r->height) + " " + r->filename + "\n";
        delete r;
    }
    if (out && cap > 0) { strncpy(out, text.c_str(), cap - 1); out[cap - 1] = 0; }
    return err;
}

// the GUI's output path (mainwindow.cpp:91-101): single-channel float -> 16-bit TIFF, otherwise float32 TIFF
__attribute__((visibility("default"))) int oss_host_save_like_gui(const char *path, const float *px, int w, int h, int c)
{
    Image img;
    img.width = w; img.height = h; img.channels = c;
    img.data.assign(px, px + (size_t)w * h * c);
    std::string e;
    return saveImageLikeGui(path, img, &e) ? 0 : -1;
}

__attribute__((visibility("default"))) void oss_host_time_string(int seconds, char *out, int cap)
{
    std::string s = getTimeString(seconds);
    strncpy(out, s.c_str(), cap - 1);
    out[cap - 1] = 0;
}

}  // extern "C"
