// Minimal stand-in for <opencv2/core.hpp> -- ONLY so that include/spx_ximgproc.hpp can be compiled and linked in an
// image without OpenCV headers (tests/cpp/Makefile).  It models the handful of cv:: names the adapter uses with the
// same spelling and semantics (Mat header/refcount-free owning buffer, _InputArray/_OutputArray wrappers, Ptr, Error).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_32S 4
#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)

namespace cv {
typedef std::string String;
class Exception : public std::runtime_error {
public:
    Exception(int c, const String& m) : std::runtime_error(m), code(c) {}
    int code;
};
#define CV_Error(code, msg) throw cv::Exception((code), (msg))

class Mat {
public:
    Mat() : rows(0), cols(0), data(nullptr), step(0), type_(0) {}
    Mat(int r, int c, int t) { create(r, c, t); }
    void create(int r, int c, int t)
    {
        if (r == rows && c == cols && t == type_ && buf_) return;
        rows = r; cols = c; type_ = t;
        step = (size_t)c * elemSize();
        buf_ = std::shared_ptr<std::vector<unsigned char>>(new std::vector<unsigned char>(step * (size_t)r));
        data = buf_->data();
    }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return type_; }
    int depth() const { return type_ & 7; }
    int channels() const { return (type_ >> CV_CN_SHIFT) + 1; }
    size_t elemSize() const { return (size_t)channels() * (depth() == CV_8U ? 1 : 4); }
    template <typename T> T& at(int y, int x) { return *reinterpret_cast<T*>(data + (size_t)y * step + (size_t)x * sizeof(T)); }
    int rows, cols;
    unsigned char* data;
    size_t step;
private:
    int type_;
    std::shared_ptr<std::vector<unsigned char>> buf_;
};
class _InputArray {
public:
    _InputArray(const Mat& m) : m_(const_cast<Mat*>(&m)) {}
    Mat getMat() const { return *m_; }
protected:
    Mat* m_;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray(Mat& m) : _InputArray(m) {}
    void create(int r, int c, int t) const { m_->create(r, c, t); }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
class Algorithm {
public:
    virtual ~Algorithm() {}
};
template <typename T> using Ptr = std::shared_ptr<T>;
template <typename T, typename... A> Ptr<T> makePtr(A&&... a) { return std::make_shared<T>(std::forward<A>(a)...); }
}  // namespace cv
