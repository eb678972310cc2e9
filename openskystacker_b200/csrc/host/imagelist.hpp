// Image-list JSON + per-file records: loadImageList / getImageRecord / getImageType of
// libstacker/src/util.cpp:66-149,213-225,465-530, without Qt.
#pragma once
#include <ctime>
#include <string>
#include <vector>

namespace openskystacker {

enum ImageType { RAW_IMAGE, FITS_IMAGE, RGB_IMAGE }; // util.h

// model.h:13-32
struct ImageRecord {
    enum FrameType { LIGHT, DARK, DARK_FLAT, FLAT, BIAS };
    std::string filename;
    FrameType type = LIGHT;
    float shutter = -1;
    float iso = -1;
    time_t timestamp = -1;
    bool reference = false;
    bool checked = true;
    int width = -1;
    int height = -1;
};

ImageType getImageType(const std::string &filename);
ImageRecord *getImageRecord(const std::string &filename);
// err: -1 not a JSON array, -2 element is not an object, -3 no filename, -4 unknown type (util.cpp:480-520)
std::vector<ImageRecord *> loadImageList(const std::string &filename, int *err);
std::string getTimeString(int seconds); // util.cpp:30-64

}  // namespace openskystacker
