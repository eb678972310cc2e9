// Build shim (ours, not reference code): the minimum of OpenCV that the
// reference's focas.cpp / adjoiningpixel.cpp touch, so that they compile
// UNMODIFIED from /root/reference into oracle/_ref/ (see oracle/Makefile).
#pragma once
#include <vector>
#include <algorithm>
#include <cmath>
#include <string>
namespace cv {
struct Point {
    int x, y;
    Point() : x(0), y(0) {}
    // cv::Point_<int>(float,float) truncates toward zero via the implicit
    // float->int conversion; adjoiningpixel.cpp:117 relies on that.
    Point(int x_, int y_) : x(x_), y(y_) {}
};
}
