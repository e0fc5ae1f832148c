// Minimal STAND-IN for <opencv2/core/core.hpp> (TEST INFRASTRUCTURE: this image has no OpenCV C++ headers).  Only the
// members r-vio_b200/host/rvio_ref_api.hpp and the System.cc transcript touch, with OpenCV's names and types, so that
// the literal-signature adaptor is type-checked by a real compiler.  Not OpenCV; never shipped.
#pragma once
#include <cstddef>
#include <cstdint>
#include <map>
#include <string>
#include <vector>
#define CV_8UC1 0
#define CV_64F 6
namespace cv {
struct Point2f { float x, y; Point2f() : x(0), y(0) {} Point2f(float a, float b) : x(a), y(b) {} };
class Mat {
public:
    Mat() : data(nullptr), cols(0), rows(0), step(0), ch_(1) {}
    Mat(int r, int c, int type, void* d, size_t s = 0) : data((unsigned char*)d), cols(c), rows(r), step(s ? s : (size_t)c * (type == CV_64F ? 8 : 1)), ch_(1) {}
    unsigned char* data; int cols, rows; size_t step;
    int channels() const { return ch_; }
    template <class T> T& at(int i, int j) { return ((T*)(data + (size_t)i * step))[j]; }
    template <class T> const T& at(int i, int j) const { return ((const T*)(data + (size_t)i * step))[j]; }
private:
    int ch_;
};
class FileNode {
public:
    FileNode() : v_(0), m_(nullptr) {}
    FileNode(double v, const std::vector<double>* m) : v_(v), m_(m) {}
    operator int() const { return (int)v_; }
    operator float() const { return (float)v_; }
    operator double() const { return v_; }
    double v_; const std::vector<double>* m_;
};
inline void operator>>(const FileNode& n, Mat& m)
{
    static std::vector<double> store;
    store = n.m_ ? *n.m_ : std::vector<double>(16, 0.0);
    m = Mat(4, 4, CV_64F, store.data());
}
class FileStorage {
public:
    FileNode operator[](const char* key) const
    {
        auto it = scalars.find(key);
        auto im = mats.find(key);
        return FileNode(it == scalars.end() ? 0.0 : it->second, im == mats.end() ? nullptr : &im->second);
    }
    std::map<std::string, double> scalars;
    std::map<std::string, std::vector<double> > mats;
};
}  // namespace cv
