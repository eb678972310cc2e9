#include "image_io.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <map>

namespace openskystacker {
namespace {

// A window [base, base + n) of the file: the whole file mapped read-only (decode: no copy of the pixel data through a
// buffer) or a few KB read around the header / the directory (probe: libtiff writes the directory AFTER the pixel data, a
// header-only probe of a 144 MB frame must not read 144 MB).  All offsets are file offsets.
struct Reader {
    const uint8_t *p = nullptr;
    size_t n = 0, base = 0, file_size = 0;
    std::vector<uint8_t> own;
    void *map = nullptr;
    size_t map_len = 0;
    bool big = false;
    Reader() = default;
    Reader(const Reader &) = delete;
    Reader &operator=(const Reader &) = delete;
    ~Reader() { release(); }
    void release() { if (map) munmap(map, map_len); map = nullptr; p = nullptr; n = 0; own.clear(); }
    bool has(size_t o, size_t len) const { return o >= base && len <= n && o - base <= n - len; }
    const uint8_t *at(size_t o) const { return p + (o - base); }
    uint16_t u16(size_t o) const { const uint8_t *b = at(o); return big ? (uint16_t)(b[0] << 8 | b[1]) : (uint16_t)(b[0] | b[1] << 8); }
    uint32_t u32(size_t o) const
    {
        const uint8_t *b = at(o);
        return big ? ((uint32_t)b[0] << 24 | (uint32_t)b[1] << 16 | (uint32_t)b[2] << 8 | b[3])
                   : ((uint32_t)b[0] | (uint32_t)b[1] << 8 | (uint32_t)b[2] << 16 | (uint32_t)b[3] << 24);
    }
};

struct Ifd {
    int width = 0, height = 0, spp = 1, bps = 8, compression = 1, photometric = 1, planar = 1, predictor = 1, format = 1;
    long rows_per_strip = -1;
    std::vector<uint32_t> offsets, counts;
    bool tiled = false;
};

bool fail(std::string *err, const std::string &m)
{
    if (err) *err = m;
    return false;
}

// bytes [lo, lo + len) of the file into the reader's own buffer (clipped to the file)
bool loadWindow(const std::string &path, size_t lo, size_t len, Reader *r, std::string *err)
{
    r->release();
    int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) return fail(err, "cannot open " + path);
    struct stat st;
    if (fstat(fd, &st) != 0) { ::close(fd); return fail(err, "cannot stat " + path); }
    r->file_size = (size_t)st.st_size;
    if (lo > r->file_size) lo = r->file_size;
    if (len > r->file_size - lo) len = r->file_size - lo;
    r->own.resize(len);
    size_t got = 0;
    while (got < len) {
        ssize_t k = pread(fd, r->own.data() + got, len - got, (off_t)(lo + got));
        if (k <= 0) break;
        got += (size_t)k;
    }
    ::close(fd);
    if (got < len) return fail(err, "cannot read " + path);
    r->p = r->own.data(); r->n = len; r->base = lo;
    return true;
}

// the whole file, mapped (read into memory where mapping is not possible)
bool loadWhole(const std::string &path, Reader *r, std::string *err)
{
    r->release();
    int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) return fail(err, "cannot open " + path);
    struct stat st;
    if (fstat(fd, &st) != 0) { ::close(fd); return fail(err, "cannot stat " + path); }
    r->file_size = (size_t)st.st_size;
    void *m = r->file_size ? mmap(nullptr, r->file_size, PROT_READ, MAP_PRIVATE, fd, 0) : MAP_FAILED;
    ::close(fd);
    if (m == MAP_FAILED) return loadWindow(path, 0, (size_t)-1, r, err);
    madvise(m, r->file_size, MADV_SEQUENTIAL);
    r->map = m; r->map_len = r->file_size;
    r->p = (const uint8_t *)m; r->n = r->file_size; r->base = 0;
    return true;
}

bool parseIfd(const Reader &r, size_t off, Ifd *ifd, std::string *err)
{
    if (!r.has(off, 2)) return fail(err, "TIFF directory lies beyond the probed window");
    int n = r.u16(off);
    if (!r.has(off + 2, (size_t)n * 12)) return fail(err, "truncated TIFF directory");
    auto values = [&](size_t e, std::vector<uint32_t> *v) {
        int type = r.u16(e + 2);
        uint32_t cnt = r.u32(e + 4);
        size_t sz = type == 3 ? 2 : (type == 4 ? 4 : 1);
        size_t p = (size_t)cnt * sz <= 4 ? e + 8 : r.u32(e + 8);
        if (!r.has(p, (size_t)cnt * sz)) return false;
        v->resize(cnt);
        for (uint32_t i = 0; i < cnt; i++) (*v)[i] = type == 3 ? r.u16(p + 2 * i) : (type == 4 ? r.u32(p + 4 * i) : *r.at(p + i));
        return true;
    };
    for (int i = 0; i < n; i++) {
        size_t e = off + 2 + (size_t)i * 12;
        int tag = r.u16(e);
        std::vector<uint32_t> v;
        bool bulk = tag == 273 || tag == 279;
        if (!values(e, &v)) {
            if (bulk) continue; // strip tables may lie beyond a header-only probe
            return fail(err, "truncated TIFF tag");
        }
        if (v.empty()) continue;
        switch (tag) {
        case 256: ifd->width = (int)v[0]; break;
        case 257: ifd->height = (int)v[0]; break;
        case 258: ifd->bps = (int)v[0]; break;
        case 259: ifd->compression = (int)v[0]; break;
        case 262: ifd->photometric = (int)v[0]; break;
        case 273: ifd->offsets = v; break;
        case 277: ifd->spp = (int)v[0]; break;
        case 278: ifd->rows_per_strip = (long)v[0]; break;
        case 279: ifd->counts = v; break;
        case 284: ifd->planar = (int)v[0]; break;
        case 317: ifd->predictor = (int)v[0]; break;
        case 339: ifd->format = (int)v[0]; break;
        case 322: case 323: case 324: case 325: ifd->tiled = true; break;
        default: break;
        }
    }
    return true;
}

bool describe(const Ifd &d, ImageInfo *info, std::string *err)
{
    if (d.width <= 0 || d.height <= 0) return fail(err, "TIFF without dimensions");
    if (d.tiled) return fail(err, "tiled TIFFs are not supported");
    if (d.planar != 1 && d.spp > 1) return fail(err, "planar TIFFs are not supported");
    if (d.spp != 1 && d.spp != 3 && d.spp != 4) return fail(err, "unsupported samples per pixel");
    info->width = d.width;
    info->height = d.height;
    info->channels = d.spp == 1 ? 1 : 3; // an alpha channel is dropped, like IMREAD_ANYCOLOR does
    if (d.bps == 8 && d.format == 1) info->sample = SAMPLE_U8;
    else if (d.bps == 16 && d.format == 1) info->sample = SAMPLE_U16;
    else if (d.bps == 32 && d.format == 3) info->sample = SAMPLE_F32;
    else return fail(err, "unsupported TIFF sample format");
    if (d.compression != 1 && d.compression != 5 && d.compression != 8 && d.compression != 32946)
        return fail(err, "unsupported TIFF compression (only none / LZW / Deflate)");
    return true;
}

// TIFF-flavoured LZW (MSB-first codes, 9..12 bits, early change)
bool lzwDecode(const uint8_t *src, size_t n, std::vector<uint8_t> *out, size_t expect)
{
    // A table entry is (where its string was last written in the OUTPUT, its length): the string of a new entry is the previous
    // code's string plus the first byte of the current one, and those bytes sit next to each other in the output — decoding a
    // code is a forward copy inside the output buffer (no chain walk, no per-code allocation; ~5x the speed of the textbook form).
    struct Entry { uint32_t pos, len; };
    Entry table[4096];
    const size_t slack = 8192; // one string may run past `expect` before the loop notices
    out->resize(expect + slack);
    uint8_t *o = out->data();
    size_t op = 0, bitpos = 0;
    int width = 9, next = 258;
    bool have_prev = false;
    size_t prev_pos = 0, prev_len = 0;
    const size_t nbits = n * 8;
    while (bitpos + width <= nbits && op < expect) {
        const size_t byte = bitpos >> 3;
        uint32_t v;
        if (byte + 4 <= n) { memcpy(&v, src + byte, 4); v = __builtin_bswap32(v); }
        else { v = 0; for (int i = 0; i < 4; i++) v = v << 8 | (byte + i < n ? src[byte + i] : 0); }
        const int code = (int)((v >> (32 - width - (bitpos & 7))) & ((1u << width) - 1));
        bitpos += width;
        if (code == 257) break;
        if (code == 256) { width = 9; next = 258; have_prev = false; continue; }
        size_t cur_pos = op, cur_len;
        if (code < 256) {
            o[op] = (uint8_t)code;
            cur_len = 1;
        } else if (have_prev && code < next) {
            const Entry e = table[code];
            cur_len = e.len;
            if (e.pos + cur_len <= op) memcpy(o + op, o + e.pos, cur_len);
            else for (size_t i = 0; i < cur_len; i++) o[op + i] = o[e.pos + i];
        } else if (have_prev && code == next) { // KwKwK: the previous string plus its own first byte
            for (size_t i = 0; i < prev_len; i++) o[op + i] = o[prev_pos + i];
            o[op + prev_len] = o[op];
            cur_len = prev_len + 1;
        } else return false;
        if (cur_len > slack) return false; // cannot happen with 12-bit codes
        op += cur_len;
        if (have_prev && next < 4096) {
            table[next] = {(uint32_t)prev_pos, (uint32_t)(prev_len + 1)};
            next++;
            if (next + 1 >= (1 << width) && width < 12) width++;
        }
        prev_pos = cur_pos; prev_len = cur_len; have_prev = true;
    }
    out->resize(std::min(op, expect + slack));
    return true;
}

template <typename T>
void storeRow(const uint8_t *row, int width, int spp, int channels, bool swap_bytes, T *dst)
{
    // the common cases without the per-sample generic path: native byte order, grey or RGB without alpha (a 24 MP RGB16 frame:
    // 0.35 s -> 0.04 s per thread)
    if (!swap_bytes && spp == 1 && channels == 1) { memcpy(dst, row, (size_t)width * sizeof(T)); return; }
    if (!swap_bytes && spp == 3 && channels == 3) {
        for (int x = 0; x < width; x++) { // RGB in the file -> BGR like imread
            T px[3];
            memcpy(px, row + (size_t)x * 3 * sizeof(T), 3 * sizeof(T));
            dst[3 * x] = px[2]; dst[3 * x + 1] = px[1]; dst[3 * x + 2] = px[0];
        }
        return;
    }
    for (int x = 0; x < width; x++) {
        T px[4];
        for (int c = 0; c < spp; c++) {
            T v;
            memcpy(&v, row + ((size_t)x * spp + c) * sizeof(T), sizeof(T));
            if (swap_bytes && sizeof(T) == 2) { uint16_t u; memcpy(&u, &v, 2); u = (uint16_t)(u << 8 | u >> 8); memcpy(&v, &u, 2); }
            if (swap_bytes && sizeof(T) == 4) { uint32_t u; memcpy(&u, &v, 4); u = __builtin_bswap32(u); memcpy(&v, &u, 4); }
            px[c] = v;
        }
        if (channels == 1) dst[x] = px[0];
        else { dst[3 * x] = px[2]; dst[3 * x + 1] = px[1]; dst[3 * x + 2] = px[0]; } // RGB in file -> BGR like imread
    }
}

// strips [nstrips * band / nbands, nstrips * (band + 1) / nbands): strips are independent (compression, predictor), so several
// threads can share one frame
bool decode(const Reader &r, const Ifd &d, const ImageInfo &info, void *dst, std::string *err, int band = 0, int nbands = 1)
{
    const size_t ssz = info.sample == SAMPLE_U8 ? 1 : info.sample == SAMPLE_U16 ? 2 : 4;
    const size_t row_bytes = (size_t)d.width * d.spp * ssz;
    long rps = d.rows_per_strip <= 0 || d.rows_per_strip > d.height ? d.height : d.rows_per_strip;
    size_t nstrips = ((size_t)d.height + rps - 1) / rps;
    if (d.offsets.size() < nstrips || d.counts.size() < nstrips) return fail(err, "TIFF strip table is incomplete");
    const bool host_little = true;
    const bool swap = r.big == host_little && ssz > 1;
    std::vector<uint8_t> strip;
    const size_t s_lo = nstrips * (size_t)band / (size_t)nbands, s_hi = nstrips * (size_t)(band + 1) / (size_t)nbands;
    for (size_t s = s_lo; s < s_hi; s++) {
        int y0 = (int)(s * rps), rows = (int)std::min<long>(rps, d.height - y0);
        size_t want = row_bytes * rows;
        if (!r.has(d.offsets[s], d.counts[s])) return fail(err, "TIFF strip lies beyond the file");
        const uint8_t *data = r.at(d.offsets[s]);
        if (d.compression == 5) {
            if (!lzwDecode(data, d.counts[s], &strip, want) || strip.size() < want) return fail(err, "corrupt LZW strip");
            data = strip.data();
        } else if (d.compression == 8 || d.compression == 32946) { // Adobe / legacy Deflate: one zlib stream per strip
            strip.resize(want);
            uLongf got = (uLongf)want;
            if (uncompress(strip.data(), &got, data, (uLong)d.counts[s]) != Z_OK || got < want) return fail(err, "corrupt Deflate strip");
            data = strip.data();
        } else if (d.counts[s] < want) return fail(err, "short TIFF strip");
        if (d.predictor == 2) { // horizontal differencing, per sample, in the file's byte order
            if (data != strip.data()) { strip.assign(data, data + want); data = strip.data(); }
            uint8_t *p = strip.data();
            for (int y = 0; y < rows; y++) {
                uint8_t *row = p + (size_t)y * row_bytes;
                for (size_t i = (size_t)d.spp; i < (size_t)d.width * d.spp; i++) {
                    if (ssz == 1) row[i] = (uint8_t)(row[i] + row[i - d.spp]);
                    else if (ssz == 2) {
                        auto get = [&](size_t k) { return r.big ? (uint16_t)(row[2 * k] << 8 | row[2 * k + 1]) : (uint16_t)(row[2 * k] | row[2 * k + 1] << 8); };
                        uint16_t v = (uint16_t)(get(i) + get(i - d.spp));
                        if (r.big) { row[2 * i] = (uint8_t)(v >> 8); row[2 * i + 1] = (uint8_t)v; }
                        else { row[2 * i] = (uint8_t)v; row[2 * i + 1] = (uint8_t)(v >> 8); }
                    } else return fail(err, "predictor on 32-bit samples is not supported");
                }
            }
        }
        for (int y = 0; y < rows; y++) {
            const uint8_t *row = data + (size_t)y * row_bytes;
            size_t o = (size_t)(y0 + y) * d.width * info.channels;
            if (ssz == 1) storeRow<uint8_t>(row, d.width, d.spp, info.channels, false, (uint8_t *)dst + o);
            else if (ssz == 2) storeRow<uint16_t>(row, d.width, d.spp, info.channels, swap, (uint16_t *)dst + o);
            else storeRow<float>(row, d.width, d.spp, info.channels, swap, (float *)dst + o);
        }
    }
    return true;
}

// limit == 0: the whole file (mapped), for decoding; otherwise a header probe that reads `limit` bytes at the start and, when the
// directory lies elsewhere (libtiff puts it behind the pixel data), `limit` bytes either side of it
bool open(const std::string &path, size_t limit, Reader *r, Ifd *d, ImageInfo *info, std::string *err)
{
    if (!(limit ? loadWindow(path, 0, limit, r, err) : loadWhole(path, r, err))) return false;
    if (!r->has(0, 8)) return fail(err, path + ": not a TIFF file");
    if (r->p[0] == 'I' && r->p[1] == 'I') r->big = false;
    else if (r->p[0] == 'M' && r->p[1] == 'M') r->big = true;
    else return fail(err, path + ": not a TIFF, PNG, JPEG or FITS file (camera RAW decoding is not part of this build)");
    if (r->u16(2) != 42) return fail(err, "not a classic TIFF (BigTIFF is not supported)");
    const size_t off = r->u32(4);
    if (limit && !r->has(off, 2 + 12 * 64)) { // directory outside the header window: fetch its neighbourhood
        const bool big = r->big;
        if (!loadWindow(path, off > limit ? off - limit : 0, (off > limit ? limit : off) + limit, r, err)) return false;
        r->big = big;
    }
    if (!parseIfd(*r, off, d, err)) {
        if (!limit) return false;
        const bool big = r->big; // out-of-line values far from the directory: read it all
        if (!loadWhole(path, r, err)) return false;
        r->big = big;
        *d = Ifd();
        if (!parseIfd(*r, off, d, err)) return false;
    }
    return describe(*d, info, err);
}

void putU16(std::vector<uint8_t> &b, uint16_t v) { b.push_back((uint8_t)v); b.push_back((uint8_t)(v >> 8)); }
void putU32(std::vector<uint8_t> &b, uint32_t v) { for (int i = 0; i < 4; i++) b.push_back((uint8_t)(v >> (8 * i))); }

template <typename T>
bool writeTiff(const std::string &path, const T *px, int w, int h, int c, int format, std::string *err)
{
    const size_t ssz = sizeof(T), data_bytes = (size_t)w * h * c * ssz;
    if (data_bytes + 1024 > 0xffffffffull) return fail(err, "image too large for a classic TIFF");
    std::vector<uint8_t> head;
    head.push_back('I'); head.push_back('I');
    putU16(head, 42);
    putU32(head, 8);
    struct Tag { uint16_t id, type; uint32_t count, value; };
    const uint32_t ifd_size = 2 + 12 * 12 + 4, bps_off = 8 + ifd_size, fmt_off = bps_off + 8, data_off = fmt_off + 8;
    std::vector<Tag> tags = {
        {256, 4, 1, (uint32_t)w}, {257, 4, 1, (uint32_t)h},
        {258, 3, (uint32_t)c, c == 1 ? (uint32_t)(8 * ssz) : bps_off},
        {259, 3, 1, 1}, {262, 3, 1, c == 1 ? 1u : 2u}, {273, 4, 1, data_off}, {277, 3, 1, (uint32_t)c},
        {278, 4, 1, (uint32_t)h}, {279, 4, 1, (uint32_t)data_bytes}, {284, 3, 1, 1},
        {305, 2, 4, 0x0073736fu /* "oss" */}, {339, 3, (uint32_t)c, c == 1 ? (uint32_t)format : fmt_off},
    }; // ascending tag order, as the TIFF specification asks (libtiff warns otherwise)
    putU16(head, (uint16_t)tags.size());
    for (auto &t : tags) { putU16(head, t.id); putU16(head, t.type); putU32(head, t.count); putU32(head, t.value); }
    putU32(head, 0);
    for (int i = 0; i < 4; i++) putU16(head, (uint16_t)(8 * ssz));
    for (int i = 0; i < 4; i++) putU16(head, (uint16_t)format);
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) return fail(err, "cannot create " + path);
    bool ok = fwrite(head.data(), 1, head.size(), f) == head.size();
    if (c == 1) ok = ok && fwrite(px, ssz, (size_t)w * h, f) == (size_t)w * h;
    else {
        std::vector<T> row((size_t)w * 3);
        for (int y = 0; ok && y < h; y++) {
            const T *s = px + (size_t)y * w * 3;
            for (int x = 0; x < w; x++) { row[3 * x] = s[3 * x + 2]; row[3 * x + 1] = s[3 * x + 1]; row[3 * x + 2] = s[3 * x]; }
            ok = fwrite(row.data(), ssz, row.size(), f) == row.size();
        }
    }
    ok = fclose(f) == 0 && ok;
    return ok ? true : fail(err, "write to " + path + " failed");
}

}  // namespace

bool probeImage(const std::string &path, ImageInfo *info, std::string *err)
{
    if (isFitsFile(path)) return probeFits(path, info, err);
    if (isPngFile(path)) return probePng(path, info, err);
    if (isJpegFile(path)) return probeJpeg(path, info, err);
    Reader r;
    Ifd d;
    return open(path, 1 << 16, &r, &d, info, err);
}

bool readImageInto(const std::string &path, const ImageInfo &expect, void *dst, std::string *err)
{
    if (isFitsFile(path)) return readFitsInto(path, &expect, nullptr, dst, nullptr, err);
    if (isPngFile(path)) return readPngInto(path, &expect, nullptr, dst, nullptr, err);
    if (isJpegFile(path)) return readJpegInto(path, &expect, nullptr, dst, nullptr, err);
    Reader r;
    Ifd d;
    ImageInfo info;
    if (!open(path, 0, &r, &d, &info, err)) return false;
    if (info.width != expect.width || info.height != expect.height || info.channels != expect.channels || info.sample != expect.sample)
        return fail(err, path + ": geometry differs from the reference frame");
    return decode(r, d, info, dst, err);
}

bool readImageBandInto(const std::string &path, const ImageInfo &expect, void *dst, int band, int nbands, std::string *err)
{
    if (nbands <= 1) return readImageInto(path, expect, dst, err);
    if (isFitsFile(path) || isPngFile(path) || isJpegFile(path)) // one stream per file: band 0 decodes it all
        return band == 0 ? readImageInto(path, expect, dst, err) : true;
    Reader r;
    Ifd d;
    ImageInfo info;
    if (!open(path, 0, &r, &d, &info, err)) return false;
    if (info.width != expect.width || info.height != expect.height || info.channels != expect.channels || info.sample != expect.sample)
        return fail(err, path + ": geometry differs from the reference frame");
    return decode(r, d, info, dst, err, band, nbands);
}

bool readImage(const std::string &path, ImageInfo *info, std::vector<uint8_t> *pixels, std::string *err)
{
    if (isFitsFile(path)) return readFitsInto(path, nullptr, info, nullptr, pixels, err);
    if (isPngFile(path)) return readPngInto(path, nullptr, info, nullptr, pixels, err);
    if (isJpegFile(path)) return readJpegInto(path, nullptr, info, nullptr, pixels, err);
    Reader r;
    Ifd d;
    if (!open(path, 0, &r, &d, info, err)) return false;
    pixels->resize(info->bytes());
    return decode(r, d, *info, pixels->data(), err);
}

bool writeTiffF32(const std::string &path, const float *px, int w, int h, int c, std::string *err) { return writeTiff<float>(path, px, w, h, c, 3, err); }
bool writeTiffU16(const std::string &path, const uint16_t *px, int w, int h, int c, std::string *err) { return writeTiff<uint16_t>(path, px, w, h, c, 1, err); }

}  // namespace openskystacker
