// openskystacker::ImageStacker — the reference's public class
// (libstacker/include/libstacker/imagestacker.h:20-114) re-hosted on liboss_b200.so.
// Same member names, argument meaning, call order, progress / error texts.  Qt is absent
// here, so QString -> std::string, QStringList -> std::vector<std::string>, cv::Mat -> Image
// and the Qt signals are std::function callbacks with the signals' names; imagestacker_qt.hpp
// shows the 20-line QObject adapter for a Qt build.
#pragma once
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "image_io.hpp"
#include "imagelist.hpp"

struct oss_ctx;

namespace openskystacker {

// float32 interleaved HWC image in OpenCV channel order: what the reference hands around as cv::Mat(CV_32FC1/3)
struct Image {
    int width = 0, height = 0, channels = 0;
    std::vector<float> data;
    bool empty() const { return data.empty(); }
};

// model.h:34-62
struct Star {
    int x = 0, y = 0, area = 0;
    float peak = 0, radius = 0, value = 0;
};

// MainWindow::finishedStacking (openskystacker/ui/mainwindow.cpp:91-101): a single-channel result is converted to 16 bits
// (`image.convertTo(converted, CV_16U, 65535)`: round-half-even of v * 65535, saturated) before imwrite, anything else is
// written as float32 — the GUI's output path (the CLI always writes float32, cl/cl.cpp:205).
bool saveImageLikeGui(const std::string &path, const Image &image, std::string *err);
// CUDA devices visible to this process (0 without a driver)
int visibleDeviceCount();

class ImageStackerPrivate;

class ImageStacker {
public:
    explicit ImageStacker(int device = 0);
    ~ImageStacker();
    ImageStacker(const ImageStacker &) = delete;
    ImageStacker &operator=(const ImageStacker &) = delete;

    bool cancel = false; // never read by process(), like the reference (imagestacker.cpp:248)

    // not in the reference: the GPUs process() shards the lights over (light k -> devices[k % N], the reference's own
    // split over its -j worker threads, util.cpp:738); default: the constructor's device alone
    void setDevices(const std::vector<int> &devices);
    std::vector<int> getDevices() const;
    // not in the reference: with several GPUs, hand the lights out from ONE shared queue (each GPU thread takes the next
    // batch when its pinned set is free) instead of the static stride: the host links of a multi-GPU box are not equal (GPUs
    // on the far socket copy at ~2/3 of the rate of the near ones when all copy at once) and a static split waits for the
    // slowest.  The per-GPU order of additions then depends on timing: the stack is reproducible to float reassociation
    // (1e-6 of full scale) instead of bit for bit; reports stay in list order.  Default: off (the reference's split).
    void setDynamicSharding(bool on);
    bool getDynamicSharding() const;

    // get/set — imagestacker.h:38-78
    std::string getRefImageFileName() const;
    void setRefImageFileName(const std::string &value);
    std::vector<std::string> getTargetImageFileNames() const;
    void setTargetImageFileNames(const std::vector<std::string> &value);
    std::vector<std::string> getDarkFrameFileNames() const;
    void setDarkFrameFileNames(const std::vector<std::string> &value);
    std::vector<std::string> getDarkFlatFrameFileNames() const;
    void setDarkFlatFrameFileNames(const std::vector<std::string> &value);
    std::vector<std::string> getFlatFrameFileNames() const;
    void setFlatFrameFileNames(const std::vector<std::string> &value);
    std::vector<std::string> getBiasFrameFileNames() const;
    void setBiasFrameFileNames(const std::vector<std::string> &value);
    std::string getSaveFilePath() const;
    void setSaveFilePath(const std::string &value);
    Image getWorkingImage() const;
    void setWorkingImage(const Image &value);
    Image getRefImage() const;
    void setRefImage(const Image &value);
    Image getFinalImage() const;
    void setFinalImage(const Image &value);
    bool getUseDarks() const;
    void setUseDarks(bool value);
    bool getUseDarkFlats() const;
    void setUseDarkFlats(bool value);
    bool getUseFlats() const;
    void setUseFlats(bool value);
    bool getUseBias() const;
    void setUseBias(bool value);

    // signals — imagestacker.h:86-101
    std::function<void(const Image &image, const std::string &message)> finished;
    std::function<void(const std::string &message, int percentComplete)> updateProgress;
    std::function<void(const std::string &message)> processingError;
    std::function<void(int)> doneDetectingStars;

    // slots — imagestacker.h:104-112
    void process(int tolerance, int threads);
    void detectStars(const std::string &filename, int threshold);

    // not in the reference: the star list behind doneDetectingStars, and per-light outcomes
    std::vector<Star> getStars(const std::string &filename, int threshold);
    struct LightReport { int nstars, matches, err; float M[6]; int overflow; };
    std::vector<LightReport> getLightReports() const;

private:
    std::unique_ptr<ImageStackerPrivate> d_ptr;
};

}  // namespace openskystacker
