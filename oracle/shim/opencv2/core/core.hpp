// oracle/shim/opencv2/core/core.hpp -- TEST INFRASTRUCTURE ONLY.
// The handful of OpenCV names the reference's PnPsolver / Sim3Solver sources mention (cv::KeyPoint with pt and
// octave; a cv::Mat that PnPsolver::Refine constructs and never reads, PnPsolver.cpp:228-231).  Not OpenCV.
#pragma once
#include <cstring>
#include <memory>
#include <vector>
#define CV_8U 0
#define CV_32F 5
#define CV_64F 6
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
struct Point3f {
    float x, y, z;
    Point3f() : x(0.f), y(0.f), z(0.f) {}
    Point3f(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
};
// a dense row-major CV_32F / CV_64F matrix: owns its storage or wraps caller memory (MLPnPsolver.cpp:133-136)
struct Mat {
    int rows, cols, type;
    void *data;
    std::shared_ptr<std::vector<double>> own;      // storage (doubles are large and aligned enough for floats too)
    Mat() : rows(0), cols(0), type(0), data(nullptr) {}
    Mat(int r, int c, int t) : rows(r), cols(c), type(t), own(std::make_shared<std::vector<double>>((size_t)r * c + 1, 0.0)) { data = own->data(); }
    Mat(int r, int c, int t, void *p) : rows(r), cols(c), type(t), data(p) {}
    size_t elem_size() const { return type == CV_64F ? 8 : (type == CV_32F ? 4 : 1); }
    // a view of row i (shares the storage), the typed row pointer, a copy (ORBmatcher.cpp: descriptor rows)
    Mat row(int i) const
    {
        Mat r(1, cols, type, static_cast<unsigned char *>(data) + (size_t)i * cols * elem_size());
        r.own = own;
        return r;
    }
    template <class T> T *ptr(int i = 0) { return reinterpret_cast<T *>(static_cast<unsigned char *>(data) + (size_t)i * cols * elem_size()); }
    template <class T> const T *ptr(int i = 0) const { return reinterpret_cast<const T *>(static_cast<const unsigned char *>(data) + (size_t)i * cols * elem_size()); }
    Mat clone() const
    {
        Mat c(rows, cols, type);
        std::memcpy(c.data, data, (size_t)rows * cols * elem_size());
        return c;
    }
    double get(int i, int j) const
    {
        return type == CV_64F ? static_cast<const double *>(data)[i * cols + j] : (double)static_cast<const float *>(data)[i * cols + j];
    }
    void convertTo(Mat &dst, int t) const
    {
        Mat out(rows, cols, t);
        for (int i = 0; i < rows; ++i)
            for (int j = 0; j < cols; ++j) {
                if (t == CV_64F) static_cast<double *>(out.data)[i * cols + j] = get(i, j);
                else static_cast<float *>(out.data)[i * cols + j] = (float)get(i, j);      // saturate_cast<float>(double): a plain conversion
            }
        dst = out;
    }
};
}  // namespace cv
