#pragma once
#define CV_32FC1 5
#define CV_32FC3 21
namespace cv {
class Mat {
public:
    Mat() {}
    Mat(int rows, int cols, int type, void *data) { (void)rows; (void)cols; (void)type; (void)data; }
    Mat clone() const { return Mat(); }
};
}
