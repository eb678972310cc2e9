#pragma once
#define CV_32FC1 5
#define CV_32FC3 21
namespace cv {
class Mat {
public:
    int rows = 0, cols = 0;
    Mat() {}
    Mat(int r, int c, int type, void *data) : rows(r), cols(c) { (void)type; (void)data; }
    Mat clone() const { return Mat(); }
    int channels() const { return 1; }
    template <typename T> const T *ptr(int y) const { (void)y; return nullptr; }
};
}
