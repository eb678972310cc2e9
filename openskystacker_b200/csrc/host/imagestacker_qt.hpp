// Qt adapter (Qt is absent from the build image: tests/test_host_layer.py type-checks this header against minimal stand-ins of the
// Qt / OpenCV classes it touches, tests/qt_stub; it is not linked into anything here).
// In a Qt build this header turns the callback-style openskystacker::ImageStacker of
// imagestacker.hpp back into the QObject the reference's CLI, GUI and tests connect to
// (cl/cl.cpp:18-23, openskystacker/ui/mainwindow.cpp:71-88, test/testoss.cpp:20-23), so that
// those callers compile unchanged against liboss_stacker.so.
#pragma once
#ifdef OSS_WITH_QT
#include <QObject>
#include <QString>
#include <QStringList>
#include <algorithm>
#include <opencv2/core/core.hpp>

#include "imagestacker.hpp"

namespace openskystacker {

class QImageStacker : public QObject {
    Q_OBJECT
    ImageStacker impl;
    static std::vector<std::string> list(const QStringList &l) { std::vector<std::string> v; for (const QString &s : l) v.push_back(s.toStdString()); return v; }
    static QStringList qlist(const std::vector<std::string> &v) { QStringList l; for (const std::string &s : v) l.push_back(QString::fromStdString(s)); return l; }
    static cv::Mat mat(const Image &i)
    {
        if (i.data.empty()) return cv::Mat();
        return cv::Mat(i.height, i.width, i.channels == 1 ? CV_32FC1 : CV_32FC3, (void *)i.data.data()).clone();
    }
    static Image image(const cv::Mat &m) // CV_32FC1 / CV_32FC3, continuous or not
    {
        Image i;
        i.width = m.cols; i.height = m.rows; i.channels = m.channels();
        i.data.resize((size_t)m.rows * m.cols * m.channels());
        for (int y = 0; y < m.rows; y++) {
            const float *row = m.ptr<float>(y);
            std::copy(row, row + (size_t)m.cols * m.channels(), i.data.begin() + (size_t)y * m.cols * m.channels());
        }
        return i;
    }
public:
    explicit QImageStacker(QObject *parent = 0, int device = 0) : QObject(parent), impl(device), cancel(impl.cancel)
    {
        impl.updateProgress = [this](const std::string &m, int p) { emit updateProgress(QString::fromStdString(m), p); };
        impl.processingError = [this](const std::string &m) { emit processingError(QString::fromStdString(m)); };
        impl.doneDetectingStars = [this](int n) { emit doneDetectingStars(n); };
        impl.finished = [this](const Image &i, const std::string &m) { emit finished(mat(i), QString::fromStdString(m)); };
    }
    bool &cancel; // imagestacker.h:36 (never read by process(), like the reference)
    ImageStacker &engine() { return impl; } // setDevices / setDynamicSharding / getLightReports: not in the reference

    // get/set — imagestacker.h:38-78
    QString getRefImageFileName() const { return QString::fromStdString(impl.getRefImageFileName()); }
    void setRefImageFileName(const QString &v) { impl.setRefImageFileName(v.toStdString()); }
    QStringList getTargetImageFileNames() const { return qlist(impl.getTargetImageFileNames()); }
    void setTargetImageFileNames(const QStringList &v) { impl.setTargetImageFileNames(list(v)); }
    QStringList getDarkFrameFileNames() const { return qlist(impl.getDarkFrameFileNames()); }
    void setDarkFrameFileNames(const QStringList &v) { impl.setDarkFrameFileNames(list(v)); }
    QStringList getDarkFlatFrameFileNames() const { return qlist(impl.getDarkFlatFrameFileNames()); }
    void setDarkFlatFrameFileNames(const QStringList &v) { impl.setDarkFlatFrameFileNames(list(v)); }
    QStringList getFlatFrameFileNames() const { return qlist(impl.getFlatFrameFileNames()); }
    void setFlatFrameFileNames(const QStringList &v) { impl.setFlatFrameFileNames(list(v)); }
    QStringList getBiasFrameFileNames() const { return qlist(impl.getBiasFrameFileNames()); }
    void setBiasFrameFileNames(const QStringList &v) { impl.setBiasFrameFileNames(list(v)); }
    QString getSaveFilePath() const { return QString::fromStdString(impl.getSaveFilePath()); }
    void setSaveFilePath(const QString &v) { impl.setSaveFilePath(v.toStdString()); }
    cv::Mat getWorkingImage() const { return mat(impl.getWorkingImage()); }
    void setWorkingImage(const cv::Mat &v) { impl.setWorkingImage(image(v)); }
    cv::Mat getRefImage() const { return mat(impl.getRefImage()); }
    void setRefImage(const cv::Mat &v) { impl.setRefImage(image(v)); }
    cv::Mat getFinalImage() const { return mat(impl.getFinalImage()); }
    void setFinalImage(const cv::Mat &v) { impl.setFinalImage(image(v)); }
    bool getUseDarks() const { return impl.getUseDarks(); }
    void setUseDarks(bool v) { impl.setUseDarks(v); }
    bool getUseDarkFlats() const { return impl.getUseDarkFlats(); }
    void setUseDarkFlats(bool v) { impl.setUseDarkFlats(v); }
    bool getUseFlats() const { return impl.getUseFlats(); }
    void setUseFlats(bool v) { impl.setUseFlats(v); }
    bool getUseBias() const { return impl.getUseBias(); }
    void setUseBias(bool v) { impl.setUseBias(v); }
signals:
    void finished(cv::Mat image, QString message);
    void updateProgress(QString message, int percentComplete);
    void processingError(QString message);
    void doneDetectingStars(int);
public slots:
    void process(int tolerance, int threads) { impl.process(tolerance, threads); }
    void detectStars(QString filename, int threshold) { impl.detectStars(filename.toStdString(), threshold); }
    // readQImage / qImageReady (imagestacker.h:95,109: the GUI's thumbnail loader) stay with the GUI: they are file decode +
    // QImage conversion, nothing of the stacking path
};

}  // namespace openskystacker
#endif
