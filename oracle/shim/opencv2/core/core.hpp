// oracle/shim/opencv2/core/core.hpp -- TEST INFRASTRUCTURE ONLY.
// The handful of OpenCV names the reference's PnPsolver / Sim3Solver sources mention (cv::KeyPoint with pt and
// octave; a cv::Mat that PnPsolver::Refine constructs and never reads, PnPsolver.cpp:228-231).  Not OpenCV.
#pragma once
#define CV_32F 5
namespace cv {
struct Point2f {
    float x, y;
    Point2f() : x(0.f), y(0.f) {}
    Point2f(float x_, float y_) : x(x_), y(y_) {}
};
struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : size(0.f), angle(-1.f), response(0.f), octave(0), class_id(-1) {}
};
struct Mat {
    int rows, cols, type;
    Mat() : rows(0), cols(0), type(0) {}
    Mat(int r, int c, int t) : rows(r), cols(c), type(t) {}
};
}  // namespace cv
