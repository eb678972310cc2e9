// Host-side image decode / encode for the ImageStacker facade (file codecs stay on the host,
// as in the reference: readImage -> cv::imread, libstacker/src/util.cpp:277-298; cl/cl.cpp:205
// cv::imwrite).  TIFF: baseline strips, 8/16-bit unsigned or 32-bit float samples, 1 or 3
// channels, uncompressed, LZW or Deflate (+ horizontal predictor), either byte order.  PNG: png_io.cpp, JPEG: jpeg_io.cpp.  Frames come back
// interleaved HWC in cv::imread's channel order (B, G, R).  FITS (fits_io.cpp): the primary HDU's first plane, BITPIX 8 / 16 /
// -32 / -64, like fitsToMat (util.cpp:300-346); recognised by content ("SIMPLE  ="), so all three entry points take either.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace openskystacker {

enum SampleKind { SAMPLE_U8 = 0, SAMPLE_U16 = 1, SAMPLE_F32 = 2 }; // == oss_sample

struct ImageInfo {
    int width = -1, height = -1, channels = 0;
    SampleKind sample = SAMPLE_U8;
    size_t bytes() const { return (size_t)width * height * channels * (sample == SAMPLE_U8 ? 1 : sample == SAMPLE_U16 ? 2 : 4); }
};

// header only: what validateImageSizes (imagestacker.cpp:402-557) needs, without decoding pixels
bool probeImage(const std::string &path, ImageInfo *info, std::string *err);
// decode into dst (info->bytes() bytes, e.g. pinned memory from oss_host_alloc)
bool readImageInto(const std::string &path, const ImageInfo &expect, void *dst, std::string *err);
// band `band` of `nbands` of a frame (TIFF: a share of its strips, which are independent; other formats: band 0 decodes the whole
// file, the other bands return at once): lets several decode workers share one large frame
bool readImageBandInto(const std::string &path, const ImageInfo &expect, void *dst, int band, int nbands, std::string *err);
bool readImage(const std::string &path, ImageInfo *info, std::vector<uint8_t> *pixels, std::string *err);
// float32 / uint16 TIFF, channel order flipped back to RGB like cv::imwrite does
bool writeTiffF32(const std::string &path, const float *hwc_bgr, int width, int height, int channels, std::string *err);
// FITS primary HDU (fits_io.cpp)
bool isFitsFile(const std::string &path);
bool probeFits(const std::string &path, ImageInfo *info, std::string *err);
bool readFitsInto(const std::string &path, const ImageInfo *expect, ImageInfo *info_out, void *dst, std::vector<uint8_t> *alloc,
                  std::string *err);
// PNG (png_io.cpp): grey / RGB / palette, 1-16 bits, alpha dropped, Adam7; recognised by content like FITS
bool isPngFile(const std::string &path);
bool probePng(const std::string &path, ImageInfo *info, std::string *err);
bool readPngInto(const std::string &path, const ImageInfo *expect, ImageInfo *info_out, void *dst, std::vector<uint8_t> *alloc,
                 std::string *err);
// JPEG (jpeg_io.cpp): 8-bit baseline / progressive Huffman files, grey or YCbCr / RGB, EXIF orientation applied like cv::imread
bool isJpegFile(const std::string &path);
bool probeJpeg(const std::string &path, ImageInfo *info, std::string *err);
bool readJpegInto(const std::string &path, const ImageInfo *expect, ImageInfo *info_out, void *dst, std::vector<uint8_t> *alloc,
                  std::string *err);
bool writeTiffU16(const std::string &path, const uint16_t *hwc_bgr, int width, int height, int channels, std::string *err);

}  // namespace openskystacker
