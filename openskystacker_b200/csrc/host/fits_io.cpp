// FITS primary-HDU reader for the ImageStacker facade: what fitsToMat (libstacker/src/util.cpp:300-346) gets out of
// CCfits, without CCfits.  The reference reads the primary image as ONE plane of axis(1) rows x axis(0) columns
// (no flip: FITS row 0 becomes Mat row 0) into a native array and scales it to float:
//   BITPIX   8  valarray<unsigned char>   -> x * 1/255            -> SAMPLE_U8  (the engine applies the same factor)
//   BITPIX  16  valarray<unsigned short>  -> x * 1/65535          -> SAMPLE_U16
//   BITPIX -32  valarray<float>           -> clone                -> SAMPLE_F32
//   BITPIX -64  valarray<double>          -> convertTo(CV_32F)    -> SAMPLE_F32 (cast here, on the host)
//   BITPIX  32  valarray<unsigned long> wrapped as CV_32SC1: 8-byte elements read as 4-byte ones (util.cpp:323-326),
//               i.e. garbage on LP64 — reported as unsupported instead of reproduced.
// CFITSIO applies BZERO / BSCALE while converting to the native type (BITPIX 16 + BZERO 32768 is how FITS stores
// unsigned 16-bit data) and saturates values that do not fit; so does this reader (round half up like CFITSIO's
// (unsigned short)(v + .5)).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "image_io.hpp"

namespace openskystacker {
namespace {

bool ffail(std::string *err, const std::string &why)
{
    if (err) *err = why;
    return false;
}

struct FitsHeader {
    int bitpix = 0, naxis = 0;
    long axis[3] = {1, 1, 1};
    double bzero = 0.0, bscale = 1.0;
    long data_offset = 0;
};

bool parseHeader(FILE *fp, const std::string &path, FitsHeader *h, std::string *err)
{
    char card[81];
    bool first = true, end = false;
    long cards = 0;
    while (!end) {
        if (fread(card, 1, 80, fp) != 80) return ffail(err, path + ": truncated FITS header");
        card[80] = 0;
        cards++;
        if (first) {
            if (strncmp(card, "SIMPLE  =", 9) != 0) return ffail(err, path + ": not a FITS file");
            first = false;
            continue;
        }
        if (strncmp(card, "END     ", 8) == 0) { end = true; break; }
        if (card[8] != '=') continue; // COMMENT, HISTORY, blank
        char key[9];
        memcpy(key, card, 8);
        key[8] = 0;
        for (int i = 7; i >= 0 && key[i] == ' '; i--) key[i] = 0;
        const char *val = card + 10;
        if (!strcmp(key, "BITPIX")) h->bitpix = atoi(val);
        else if (!strcmp(key, "NAXIS")) h->naxis = atoi(val);
        else if (!strcmp(key, "NAXIS1")) h->axis[0] = atol(val);
        else if (!strcmp(key, "NAXIS2")) h->axis[1] = atol(val);
        else if (!strcmp(key, "NAXIS3")) h->axis[2] = atol(val);
        else if (!strcmp(key, "BZERO")) h->bzero = atof(val);
        else if (!strcmp(key, "BSCALE")) h->bscale = atof(val);
        if (cards > 36 * 1000) return ffail(err, path + ": FITS header without END");
    }
    h->data_offset = (cards * 80 + 2879) / 2880 * 2880;
    return true;
}

bool describeFits(const FitsHeader &h, const std::string &path, ImageInfo *info, std::string *err)
{
    if (h.naxis < 2 || h.axis[0] <= 0 || h.axis[1] <= 0) return ffail(err, path + ": FITS primary HDU holds no 2-D image");
    info->width = (int)h.axis[0];  // util.cpp:309  cols = axis(0)
    info->height = (int)h.axis[1]; // util.cpp:308  rows = axis(1)
    info->channels = 1;
    switch (h.bitpix) {
    case 8: info->sample = SAMPLE_U8; break;
    case 16: info->sample = SAMPLE_U16; break;
    case -32: case -64: info->sample = SAMPLE_F32; break;
    case 32: return ffail(err, path + ": BITPIX 32 FITS images are not supported (the reference mis-reads them, util.cpp:323-326)");
    default: return ffail(err, path + ": unsupported FITS BITPIX");
    }
    return true;
}

inline uint16_t be16(const uint8_t *p) { return (uint16_t)((p[0] << 8) | p[1]); }
inline uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline uint64_t be64(const uint8_t *p) { return ((uint64_t)be32(p) << 32) | be32(p + 4); }

template <typename T>
inline T saturate(double v, double lo, double hi)
{
    if (!(v >= lo)) return (T)lo; // also NaN
    if (v > hi) return (T)hi;
    return (T)(v + 0.5);
}

bool decodeFits(FILE *fp, const FitsHeader &h, const ImageInfo &info, const std::string &path, void *dst, std::string *err)
{
    const size_t n = (size_t)info.width * info.height; // first plane only, like Mat(rows, cols, type, &contents[0])
    const size_t esz = (size_t)(h.bitpix < 0 ? -h.bitpix : h.bitpix) / 8;
    std::vector<uint8_t> raw(n * esz);
    if (fseek(fp, h.data_offset, SEEK_SET) != 0 || fread(raw.data(), 1, raw.size(), fp) != raw.size())
        return ffail(err, path + ": truncated FITS data");
    const bool plain = h.bzero == 0.0 && h.bscale == 1.0;
    switch (h.bitpix) {
    case 8: {
        uint8_t *o = (uint8_t *)dst;
        if (plain) memcpy(o, raw.data(), n);
        else for (size_t i = 0; i < n; i++) o[i] = saturate<uint8_t>(h.bzero + h.bscale * raw[i], 0.0, 255.0);
        break;
    }
    case 16: {
        uint16_t *o = (uint16_t *)dst;
        if (h.bzero == 32768.0 && h.bscale == 1.0) // unsigned 16-bit data: exact
            for (size_t i = 0; i < n; i++) o[i] = (uint16_t)(be16(&raw[2 * i]) ^ 0x8000u);
        else
            for (size_t i = 0; i < n; i++) o[i] = saturate<uint16_t>(h.bzero + h.bscale * (double)(int16_t)be16(&raw[2 * i]), 0.0, 65535.0);
        break;
    }
    case -32: {
        float *o = (float *)dst;
        for (size_t i = 0; i < n; i++) {
            uint32_t u = be32(&raw[4 * i]);
            float f;
            memcpy(&f, &u, 4);
            o[i] = plain ? f : (float)(h.bzero + h.bscale * (double)f);
        }
        break;
    }
    case -64: {
        float *o = (float *)dst;
        for (size_t i = 0; i < n; i++) {
            uint64_t u = be64(&raw[8 * i]);
            double d;
            memcpy(&d, &u, 8);
            o[i] = (float)(plain ? d : h.bzero + h.bscale * d); // convertTo(CV_32F): saturate_cast<float>(double)
        }
        break;
    }
    default: return ffail(err, path + ": unsupported FITS BITPIX");
    }
    return true;
}

}  // namespace

bool isFitsFile(const std::string &path)
{
    FILE *fp = fopen(path.c_str(), "rb");
    if (!fp) return false;
    char magic[9] = {0};
    size_t got = fread(magic, 1, 9, fp);
    fclose(fp);
    return got == 9 && strncmp(magic, "SIMPLE  =", 9) == 0;
}

bool probeFits(const std::string &path, ImageInfo *info, std::string *err)
{
    FILE *fp = fopen(path.c_str(), "rb");
    if (!fp) return ffail(err, "cannot open " + path);
    FitsHeader h;
    bool ok = parseHeader(fp, path, &h, err) && describeFits(h, path, info, err);
    fclose(fp);
    return ok;
}

bool readFitsInto(const std::string &path, const ImageInfo *expect, ImageInfo *info_out, void *dst, std::vector<uint8_t> *alloc,
                  std::string *err)
{
    FILE *fp = fopen(path.c_str(), "rb");
    if (!fp) return ffail(err, "cannot open " + path);
    FitsHeader h;
    ImageInfo info;
    bool ok = parseHeader(fp, path, &h, err) && describeFits(h, path, &info, err);
    if (ok && expect &&
        (info.width != expect->width || info.height != expect->height || info.channels != expect->channels || info.sample != expect->sample))
        ok = ffail(err, path + ": geometry differs from the reference frame");
    if (ok && alloc) { alloc->resize(info.bytes()); dst = alloc->data(); }
    ok = ok && decodeFits(fp, h, info, path, dst, err);
    fclose(fp);
    if (ok && info_out) *info_out = info;
    return ok;
}

}  // namespace openskystacker
