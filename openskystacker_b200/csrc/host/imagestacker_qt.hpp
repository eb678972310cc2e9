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
#include <opencv2/core/core.hpp>

#include "imagestacker.hpp"

namespace openskystacker {

class QImageStacker : public QObject {
    Q_OBJECT
    ImageStacker impl;
    static std::vector<std::string> list(const QStringList &l) { std::vector<std::string> v; for (const QString &s : l) v.push_back(s.toStdString()); return v; }
public:
    explicit QImageStacker(QObject *parent = 0) : QObject(parent)
    {
        impl.updateProgress = [this](const std::string &m, int p) { emit updateProgress(QString::fromStdString(m), p); };
        impl.processingError = [this](const std::string &m) { emit processingError(QString::fromStdString(m)); };
        impl.doneDetectingStars = [this](int n) { emit doneDetectingStars(n); };
        impl.finished = [this](const Image &i, const std::string &m) {
            cv::Mat mat(i.height, i.width, i.channels == 1 ? CV_32FC1 : CV_32FC3, (void *)i.data.data());
            emit finished(mat.clone(), QString::fromStdString(m));
        };
    }
    void setRefImageFileName(const QString &v) { impl.setRefImageFileName(v.toStdString()); }
    void setTargetImageFileNames(const QStringList &v) { impl.setTargetImageFileNames(list(v)); }
    void setDarkFrameFileNames(const QStringList &v) { impl.setDarkFrameFileNames(list(v)); }
    void setDarkFlatFrameFileNames(const QStringList &v) { impl.setDarkFlatFrameFileNames(list(v)); }
    void setFlatFrameFileNames(const QStringList &v) { impl.setFlatFrameFileNames(list(v)); }
    void setBiasFrameFileNames(const QStringList &v) { impl.setBiasFrameFileNames(list(v)); }
    void setUseDarks(bool v) { impl.setUseDarks(v); }
    void setUseDarkFlats(bool v) { impl.setUseDarkFlats(v); }
    void setUseFlats(bool v) { impl.setUseFlats(v); }
    void setUseBias(bool v) { impl.setUseBias(v); }
signals:
    void finished(cv::Mat image, QString message);
    void updateProgress(QString message, int percentComplete);
    void processingError(QString message);
    void doneDetectingStars(int);
public slots:
    void process(int tolerance, int threads) { impl.process(tolerance, threads); }
    void detectStars(QString filename, int threshold) { impl.detectStars(filename.toStdString(), threshold); }
};

}  // namespace openskystacker
#endif
