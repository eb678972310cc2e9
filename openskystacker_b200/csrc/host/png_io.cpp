// PNG input for the ImageStacker facade: the reference's RGB_IMAGE branch reads whatever cv::imread(ANYDEPTH | ANYCOLOR)
// reads (libstacker/src/util.cpp:291-294), and 16-bit PNG is the other lossless container astro cameras / converters
// produce.  Chunks are parsed here, the deflate stream goes through zlib's inflate (the only codec library in the build
// image), rows are un-filtered and handed over in cv::imread's layout:
//   grey -> 1 channel;  RGB / RGBA / palette -> 3 channels in B, G, R order (alpha dropped);  grey + alpha -> 3 equal channels
//   (what cv2 4.13, the test pin, returns for it; OpenCV 3.3's PngDecoder typed it CV_8UC1 — rare enough in astro data)
//   bit depth 16 -> uint16 (PNG samples are big-endian), everything else -> uint8 (1/2/4-bit grey scaled like libpng's
//   png_set_expand_gray_1_2_4_to_8, palette indices expanded through PLTE);  Adam7-interlaced files are de-interlaced.
#include <zlib.h>

#include <cstring>
#include <fstream>

#include "image_io.hpp"

namespace openskystacker {
namespace {

bool pfail(std::string *err, const std::string &m)
{
    if (err) *err = m;
    return false;
}

struct PngHeader {
    uint32_t width = 0, height = 0;
    int depth = 0, color = 0, interlace = 0;
    int samples() const { return color == 0 ? 1 : color == 2 ? 3 : color == 3 ? 1 : color == 4 ? 2 : 4; }
    int bits_per_pixel() const { return samples() * depth; }
};

uint32_t be32(const uint8_t *p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }

const uint8_t kSig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};

bool parseHeader(const uint8_t *b, size_t n, PngHeader *h, std::string *err)
{
    if (n < 33 || memcmp(b, kSig, 8) || be32(b + 8) != 13 || memcmp(b + 12, "IHDR", 4)) return pfail(err, "not a PNG file");
    h->width = be32(b + 16); h->height = be32(b + 20);
    h->depth = b[24]; h->color = b[25]; h->interlace = b[28];
    if (b[26] != 0 || b[27] != 0 || h->interlace > 1) return pfail(err, "unsupported PNG compression / filter / interlace method");
    const bool ok = (h->color == 0 && (h->depth == 1 || h->depth == 2 || h->depth == 4 || h->depth == 8 || h->depth == 16)) ||
                    (h->color == 3 && (h->depth == 1 || h->depth == 2 || h->depth == 4 || h->depth == 8)) ||
                    ((h->color == 2 || h->color == 4 || h->color == 6) && (h->depth == 8 || h->depth == 16));
    if (!ok || !h->width || !h->height || h->width > (1u << 30) || h->height > (1u << 30)) return pfail(err, "invalid PNG header");
    return true;
}

void describe(const PngHeader &h, ImageInfo *info)
{
    info->width = (int)h.width;
    info->height = (int)h.height;
    info->channels = h.color == 0 ? 1 : 3;
    info->sample = h.depth == 16 ? SAMPLE_U16 : SAMPLE_U8;
}

int paeth(int a, int b, int c)
{
    int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// un-filter `rows` scanlines of `stride` bytes each (every line preceded by its filter byte) in place; returns false on a bad filter
bool unfilter(uint8_t *data, size_t rows, size_t stride, int bpp)
{
    std::vector<uint8_t> zero(stride, 0);
    const uint8_t *prev = zero.data();
    for (size_t y = 0; y < rows; y++) {
        uint8_t *line = data + y * (stride + 1);
        const int f = line[0];
        uint8_t *cur = line + 1;
        switch (f) {
        case 0: break;
        case 1: for (size_t i = (size_t)bpp; i < stride; i++) cur[i] = (uint8_t)(cur[i] + cur[i - bpp]); break;
        case 2: for (size_t i = 0; i < stride; i++) cur[i] = (uint8_t)(cur[i] + prev[i]); break;
        case 3:
            for (size_t i = 0; i < stride; i++) cur[i] = (uint8_t)(cur[i] + (((i >= (size_t)bpp ? cur[i - bpp] : 0) + prev[i]) >> 1));
            break;
        case 4:
            for (size_t i = 0; i < stride; i++)
                cur[i] = (uint8_t)(cur[i] + paeth(i >= (size_t)bpp ? cur[i - bpp] : 0, prev[i], i >= (size_t)bpp ? prev[i - bpp] : 0));
            break;
        default: return false;
        }
        prev = cur;
    }
    return true;
}

// one un-filtered scanline of `w` pixels -> destination pixels x0, x0 + dx, ... of the output row (cv::imread's layout)
template <typename T>
void storePixels(const PngHeader &h, const uint8_t *line, uint32_t w, const std::vector<uint8_t> &plte, T *dst_row, uint32_t x0, uint32_t dx,
                 int channels)
{
    const int ns = h.samples();
    for (uint32_t i = 0; i < w; i++) {
        T *d = dst_row + (size_t)(x0 + i * dx) * channels;
        if (h.depth == 16) {
            const uint8_t *s = line + (size_t)i * ns * 2;
            auto at = [&](int k) { return (T)(s[2 * k] << 8 | s[2 * k + 1]); };
            if (channels == 1) d[0] = at(0);
            else if (ns <= 2) d[0] = d[1] = d[2] = at(0);      // grey + alpha
            else { d[0] = at(2); d[1] = at(1); d[2] = at(0); } // RGB(A) -> BGR
        } else if (h.depth == 8) {
            const uint8_t *s = line + (size_t)i * ns;
            if (h.color == 3) { const uint8_t *p = plte.data() + 3 * (size_t)s[0]; d[0] = p[2]; d[1] = p[1]; d[2] = p[0]; }
            else if (channels == 1) d[0] = s[0];
            else if (ns <= 2) d[0] = d[1] = d[2] = s[0];
            else { d[0] = s[2]; d[1] = s[1]; d[2] = s[0]; }
        } else { // 1, 2, 4 bits: grey (scaled to 0..255) or palette index
            const int per = 8 / h.depth, shift = (per - 1 - (int)(i % per)) * h.depth, mask = (1 << h.depth) - 1;
            const int v = (line[i / per] >> shift) & mask;
            if (h.color == 3) { const uint8_t *p = plte.data() + 3 * (size_t)v; d[0] = p[2]; d[1] = p[1]; d[2] = p[0]; }
            else d[0] = (T)(v * (255 / mask));
        }
    }
}

bool decodePng(const std::vector<uint8_t> &file, const PngHeader &h, const ImageInfo &info, void *dst, std::string *err)
{
    // gather IDAT, PLTE
    std::vector<uint8_t> plte(768, 0);
    z_stream z;
    memset(&z, 0, sizeof z);
    if (inflateInit(&z) != Z_OK) return pfail(err, "zlib initialisation failed");
    const int bpp_bits = h.bits_per_pixel(), bpp = bpp_bits >= 8 ? bpp_bits / 8 : 1;
    // pass geometry (Adam7) or the whole image as a single pass
    static const int xs[7] = {0, 4, 0, 2, 0, 1, 0}, ys[7] = {0, 0, 4, 0, 2, 0, 1}, dxs[7] = {8, 8, 4, 4, 2, 2, 1}, dys[7] = {8, 8, 8, 4, 4, 2, 2};
    struct Pass { uint32_t w, hgt, x0, y0, dx, dy; size_t stride, offset; };
    std::vector<Pass> passes;
    size_t total = 0;
    const int npass = h.interlace ? 7 : 1;
    for (int p = 0; p < npass; p++) {
        Pass q;
        if (h.interlace) {
            q.x0 = xs[p]; q.y0 = ys[p]; q.dx = dxs[p]; q.dy = dys[p];
            q.w = (h.width + q.dx - 1 - q.x0) / q.dx; q.hgt = (h.height + q.dy - 1 - q.y0) / q.dy;
        } else { q.x0 = q.y0 = 0; q.dx = q.dy = 1; q.w = h.width; q.hgt = h.height; }
        if (q.w == 0 || q.hgt == 0) continue;
        q.stride = ((size_t)q.w * bpp_bits + 7) / 8;
        q.offset = total;
        total += (q.stride + 1) * q.hgt;
        passes.push_back(q);
    }
    std::vector<uint8_t> raw(total);
    z.next_out = raw.data();
    z.avail_out = (uInt)std::min<size_t>(total, 0x7fffffffu);
    size_t produced_before = 0;
    bool done = false, have_plte = false;
    size_t pos = 8;
    while (pos + 12 <= file.size()) {
        const uint32_t len = be32(file.data() + pos);
        const uint8_t *type = file.data() + pos + 4, *body = file.data() + pos + 8;
        if (pos + 12 + (size_t)len > file.size()) { inflateEnd(&z); return pfail(err, "truncated PNG chunk"); }
        if (!memcmp(type, "PLTE", 4)) { memcpy(plte.data(), body, std::min<size_t>(len, 768)); have_plte = true; }
        else if (!memcmp(type, "IDAT", 4) && !done) {
            z.next_in = const_cast<Bytef *>(body);
            z.avail_in = len;
            while (z.avail_in && !done) {
                if (z.avail_out == 0) { // more than 2 GiB of scanlines: re-arm the output window
                    produced_before = (size_t)(z.next_out - raw.data());
                    if (produced_before >= total) { done = true; break; }
                    z.avail_out = (uInt)std::min<size_t>(total - produced_before, 0x7fffffffu);
                }
                const int rc = inflate(&z, Z_NO_FLUSH);
                if (rc == Z_STREAM_END) done = true;
                else if (rc != Z_OK) { inflateEnd(&z); return pfail(err, "corrupt PNG image data"); }
            }
        } else if (!memcmp(type, "IEND", 4)) break;
        pos += 12 + (size_t)len;
    }
    const size_t produced = (size_t)(z.next_out - raw.data());
    inflateEnd(&z);
    if (produced < total) return pfail(err, "PNG image data ends early");
    if (h.color == 3 && !have_plte) return pfail(err, "palette PNG without PLTE");
    for (const Pass &q : passes) {
        uint8_t *base = raw.data() + q.offset;
        if (!unfilter(base, q.hgt, q.stride, bpp)) return pfail(err, "bad PNG filter type");
        for (uint32_t y = 0; y < q.hgt; y++) {
            const uint8_t *line = base + (size_t)y * (q.stride + 1) + 1;
            const size_t row = (size_t)(q.y0 + y * q.dy) * h.width * info.channels;
            if (info.sample == SAMPLE_U16) storePixels<uint16_t>(h, line, q.w, plte, (uint16_t *)dst + row, q.x0, q.dx, info.channels);
            else storePixels<uint8_t>(h, line, q.w, plte, (uint8_t *)dst + row, q.x0, q.dx, info.channels);
        }
    }
    return true;
}

bool loadAll(const std::string &path, std::vector<uint8_t> *out, size_t limit, std::string *err)
{
    std::ifstream f(path, std::ios::binary);
    if (!f) return pfail(err, "cannot open " + path);
    f.seekg(0, std::ios::end);
    size_t n = (size_t)f.tellg();
    f.seekg(0);
    if (limit && n > limit) n = limit;
    out->resize(n);
    f.read((char *)out->data(), (std::streamsize)n);
    return true;
}

}  // namespace

bool isPngFile(const std::string &path)
{
    std::ifstream f(path, std::ios::binary);
    uint8_t sig[8] = {0};
    return f && f.read((char *)sig, 8) && !memcmp(sig, kSig, 8);
}

bool probePng(const std::string &path, ImageInfo *info, std::string *err)
{
    std::vector<uint8_t> head;
    PngHeader h;
    if (!loadAll(path, &head, 64, err) || !parseHeader(head.data(), head.size(), &h, err)) return false;
    describe(h, info);
    return true;
}

bool readPngInto(const std::string &path, const ImageInfo *expect, ImageInfo *info_out, void *dst, std::vector<uint8_t> *alloc, std::string *err)
{
    std::vector<uint8_t> file;
    PngHeader h;
    if (!loadAll(path, &file, 0, err) || !parseHeader(file.data(), file.size(), &h, err)) return false;
    ImageInfo info;
    describe(h, &info);
    if (expect && (info.width != expect->width || info.height != expect->height || info.channels != expect->channels || info.sample != expect->sample))
        return pfail(err, path + ": geometry differs from the reference frame");
    if (info_out) *info_out = info;
    if (alloc) { alloc->resize(info.bytes()); dst = alloc->data(); }
    return decodePng(file, h, info, dst, err);
}

}  // namespace openskystacker
