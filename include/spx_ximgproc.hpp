// spx_ximgproc.hpp -- header-only C++ adapter: the cv::ximgproc superpixel surface the reference GUI
// calls (mainwindow.h:170-172, mainwindow.cpp:2096-2127), implemented on the C ABI of libspx_b200.so.
//
// Usage in the reference:  replace  #include <opencv2/ximgproc.hpp>  (mainwindow.h:19) by
//     #include <opencv2/core.hpp>
//     #include "spx_ximgproc.hpp"
// and link -lspx_b200 instead of opencv_ximgproc.  The classes live in namespace spx::ximgproc and are
// re-exported into cv::ximgproc when SPX_XIMGPROC_EXPORT_TO_CV is defined (the default), so that
// mainwindow.cpp compiles unchanged.  Every non-zero status of the C ABI becomes a cv::Exception with the
// same code (CV_Error), exactly the failure mode of upstream's CV_Assert.
//
// Requires only cv::Mat, cv::Ptr/makePtr, cv::InputArray/OutputArray, cv::Algorithm, CV_Error (opencv2/core.hpp).
#pragma once
#include "spx_b200.h"

#ifndef SPX_XIMGPROC_NO_OPENCV_INCLUDE
#include <opencv2/core.hpp>
#endif

namespace spx { namespace ximgproc {

enum SLICType { SLIC = SPX_SLIC, SLICO = SPX_SLICO, MSLIC = SPX_MSLIC };   // cv::ximgproc::SLICType (slic.hpp)

namespace detail {
inline void check(int rc, const char* where)
{
    if (rc != SPX_OK) CV_Error(rc, cv::String(where) + ": " + spx_last_error());
}
// the GUI passes a continuous CV_8UC3 Mat (mainwindow.cpp:2093-2097); any 8-bit 1..4-channel Mat is accepted
inline cv::Mat image8u(cv::InputArray a, const char* where)
{
    cv::Mat m = a.getMat();
    if (m.empty() || m.depth() != CV_8U || m.channels() < 1 || m.channels() > 4)
        CV_Error(SPX_ERR_BADARG, cv::String(where) + ": 8-bit image with 1..4 channels required");
    return m;
}
inline int device()
{
    return 0;   // one image stays on one GPU
}
}  // namespace detail

// ---- cv::ximgproc::SuperpixelSLIC (mainwindow.h:170) ------------------------------------------------
class SuperpixelSLIC : public cv::Algorithm {
public:
    // colour_code < 0: `image` is what the GUI passes today (already Lab / HSV).  COLOR_BGR2Lab (44) / COLOR_BGR2HSV (40): `image`
    // is BGR and the cvtColor of mainwindow.cpp:2096-2097 runs on the device after the upload.
    SuperpixelSLIC(cv::InputArray image, int algorithm, int region_size, float ruler, int colour_code = -1) : h_(nullptr), w_(0), hgt_(0)
    {
        cv::Mat m = detail::image8u(image, "createSuperpixelSLIC");
        w_ = m.cols; hgt_ = m.rows;
        if (colour_code < 0)
            detail::check(spx_slic_create(m.data, m.step, m.cols, m.rows, m.channels(), algorithm, region_size, ruler,
                                          detail::device(), &h_), "createSuperpixelSLIC");
        else
            detail::check(spx_slic_create_bgr(m.data, m.step, m.cols, m.rows, colour_code, algorithm, region_size, ruler,
                                              detail::device(), &h_), "createSuperpixelSLIC");
    }
    ~SuperpixelSLIC() { spx_slic_destroy(h_); }
    SuperpixelSLIC(const SuperpixelSLIC&) = delete;
    SuperpixelSLIC& operator=(const SuperpixelSLIC&) = delete;

    virtual int getNumberOfSuperpixels() const                                   // mainwindow.cpp:2109
    {
        int n = 0;
        detail::check(spx_slic_get_number_of_superpixels(h_, &n), "SuperpixelSLIC::getNumberOfSuperpixels");
        return n;
    }
    virtual void iterate(int num_iterations = 10)                                // mainwindow.cpp:2106
    {
        detail::check(spx_slic_iterate(h_, num_iterations), "SuperpixelSLIC::iterate");
    }
    virtual void getLabels(cv::OutputArray labels_out) const                     // mainwindow.cpp:2110
    {
        labels_out.create(hgt_, w_, CV_32SC1);
        cv::Mat m = labels_out.getMat();
        detail::check(spx_slic_get_labels(h_, reinterpret_cast<int32_t*>(m.data), m.step), "SuperpixelSLIC::getLabels");
    }
    virtual void getLabelContourMask(cv::OutputArray image, bool thick_line = true) const   // mainwindow.cpp:2111
    {
        image.create(hgt_, w_, CV_8UC1);
        cv::Mat m = image.getMat();
        detail::check(spx_slic_get_label_contour_mask(h_, m.data, m.step, thick_line ? 1 : 0),
                      "SuperpixelSLIC::getLabelContourMask");
    }
    virtual void enforceLabelConnectivity(int min_element_size = 25)             // mainwindow.cpp:2108
    {
        detail::check(spx_slic_enforce_label_connectivity(h_, min_element_size), "SuperpixelSLIC::enforceLabelConnectivity");
    }
private:
    spx_slic_t* h_;
    int w_, hgt_;
};
inline cv::Ptr<SuperpixelSLIC> createSuperpixelSLIC(cv::InputArray image, int algorithm = SLICO, int region_size = 10,
                                                    float ruler = 10.0f)       // mainwindow.cpp:2105
{
    return cv::makePtr<SuperpixelSLIC>(image, algorithm, region_size, ruler);
}
// cvtColor(image, converted, code) + createSuperpixelSLIC(converted, ...) of mainwindow.cpp:2096-2105 in one call
inline cv::Ptr<SuperpixelSLIC> createSuperpixelSLIC_BGR(cv::InputArray bgr, int code, int algorithm = SLICO, int region_size = 10,
                                                        float ruler = 10.0f)
{
    return cv::makePtr<SuperpixelSLIC>(bgr, algorithm, region_size, ruler, code);
}

// ---- cv::ximgproc::SuperpixelLSC (mainwindow.h:171) -------------------------------------------------
class SuperpixelLSC : public cv::Algorithm {
public:
    SuperpixelLSC(cv::InputArray image, int region_size, float ratio) : h_(nullptr), w_(0), hgt_(0)
    {
        cv::Mat m = detail::image8u(image, "createSuperpixelLSC");
        w_ = m.cols; hgt_ = m.rows;
        detail::check(spx_lsc_create(m.data, m.step, m.cols, m.rows, m.channels(), region_size, ratio, detail::device(), &h_),
                      "createSuperpixelLSC");
    }
    ~SuperpixelLSC() { spx_lsc_destroy(h_); }
    SuperpixelLSC(const SuperpixelLSC&) = delete;
    SuperpixelLSC& operator=(const SuperpixelLSC&) = delete;

    virtual int getNumberOfSuperpixels() const                                   // mainwindow.cpp:2118
    {
        int n = 0;
        detail::check(spx_lsc_get_number_of_superpixels(h_, &n), "SuperpixelLSC::getNumberOfSuperpixels");
        return n;
    }
    virtual void iterate(int num_iterations = 10)                                // mainwindow.cpp:2115
    {
        detail::check(spx_lsc_iterate(h_, num_iterations), "SuperpixelLSC::iterate");
    }
    virtual void getLabels(cv::OutputArray labels_out) const                     // mainwindow.cpp:2119
    {
        labels_out.create(hgt_, w_, CV_32SC1);
        cv::Mat m = labels_out.getMat();
        detail::check(spx_lsc_get_labels(h_, reinterpret_cast<int32_t*>(m.data), m.step), "SuperpixelLSC::getLabels");
    }
    virtual void getLabelContourMask(cv::OutputArray image, bool thick_line = true) const   // mainwindow.cpp:2120
    {
        image.create(hgt_, w_, CV_8UC1);
        cv::Mat m = image.getMat();
        detail::check(spx_lsc_get_label_contour_mask(h_, m.data, m.step, thick_line ? 1 : 0),
                      "SuperpixelLSC::getLabelContourMask");
    }
    virtual void enforceLabelConnectivity(int min_element_size = 25)             // mainwindow.cpp:2117
    {
        detail::check(spx_lsc_enforce_label_connectivity(h_, min_element_size), "SuperpixelLSC::enforceLabelConnectivity");
    }
private:
    spx_lsc_t* h_;
    int w_, hgt_;
};
inline cv::Ptr<SuperpixelLSC> createSuperpixelLSC(cv::InputArray image, int region_size = 10, float ratio = 0.075f)
{                                                                                 // mainwindow.cpp:2114
    return cv::makePtr<SuperpixelLSC>(image, region_size, ratio);
}

// ---- cv::ximgproc::SuperpixelSEEDS (mainwindow.h:172) -----------------------------------------------
class SuperpixelSEEDS : public cv::Algorithm {
public:
    SuperpixelSEEDS(int image_width, int image_height, int image_channels, int num_superpixels, int num_levels, int prior,
                    int histogram_bins, bool double_step)
        : h_(nullptr), w_(image_width), hgt_(image_height), c_(image_channels)
    {
        detail::check(spx_seeds_create(image_width, image_height, image_channels, num_superpixels, num_levels, prior,
                                       histogram_bins, double_step ? 1 : 0, detail::device(), &h_), "createSuperpixelSEEDS");
    }
    ~SuperpixelSEEDS() { spx_seeds_destroy(h_); }
    SuperpixelSEEDS(const SuperpixelSEEDS&) = delete;
    SuperpixelSEEDS& operator=(const SuperpixelSEEDS&) = delete;

    virtual int getNumberOfSuperpixels()                                          // mainwindow.cpp:2125
    {
        int n = 0;
        detail::check(spx_seeds_get_number_of_superpixels(h_, &n), "SuperpixelSEEDS::getNumberOfSuperpixels");
        return n;
    }
    virtual void iterate(cv::InputArray img, int num_iterations = 4)              // mainwindow.cpp:2124
    {
        cv::Mat m = detail::image8u(img, "SuperpixelSEEDS::iterate");
        if (m.cols != w_ || m.rows != hgt_ || m.channels() != c_)
            CV_Error(SPX_ERR_ASSERT, "SuperpixelSEEDS::iterate: image size/channels differ from createSuperpixelSEEDS");
        detail::check(spx_seeds_iterate(h_, m.data, m.step, num_iterations), "SuperpixelSEEDS::iterate");
    }
    virtual void getLabels(cv::OutputArray labels_out)                            // mainwindow.cpp:2126
    {
        labels_out.create(hgt_, w_, CV_32SC1);
        cv::Mat m = labels_out.getMat();
        detail::check(spx_seeds_get_labels(h_, reinterpret_cast<int32_t*>(m.data), m.step), "SuperpixelSEEDS::getLabels");
    }
    virtual void getLabelContourMask(cv::OutputArray image, bool thick_line = false)   // mainwindow.cpp:2127
    {
        image.create(hgt_, w_, CV_8UC1);
        cv::Mat m = image.getMat();
        detail::check(spx_seeds_get_label_contour_mask(h_, m.data, m.step, thick_line ? 1 : 0),
                      "SuperpixelSEEDS::getLabelContourMask");
    }
private:
    spx_seeds_t* h_;
    int w_, hgt_, c_;
};
inline cv::Ptr<SuperpixelSEEDS> createSuperpixelSEEDS(int image_width, int image_height, int image_channels,
                                                      int num_superpixels, int num_levels, int prior = 2,
                                                      int histogram_bins = 5, bool double_step = false)
{                                                                                 // mainwindow.cpp:2123
    return cv::makePtr<SuperpixelSEEDS>(image_width, image_height, image_channels, num_superpixels, num_levels, prior,
                                        histogram_bins, double_step);
}

// ---- cv::cvtColor(src, dst, COLOR_BGR2Lab / COLOR_BGR2HSV) for CV_8UC3 (mainwindow.cpp:2096-2097) ----
// The GUI calls cv::cvtColor from imgproc, which it keeps linking; this helper is the GPU replacement for
// exactly the two codes on the path (44 = COLOR_BGR2Lab, 40 = COLOR_BGR2HSV). In place is allowed.
inline void cvtColor8u(cv::InputArray src, cv::OutputArray dst, int code)
{
    cv::Mat s = src.getMat();
    if (s.empty() || s.type() != CV_8UC3) CV_Error(SPX_ERR_BADARG, "cvtColor8u: CV_8UC3 required");
    dst.create(s.rows, s.cols, CV_8UC3);
    cv::Mat d = dst.getMat();
    int rc;
    if (code == 44) rc = spx_bgr2lab8(s.data, s.step, d.data, d.step, s.cols, s.rows, detail::device());
    else if (code == 40) rc = spx_bgr2hsv8(s.data, s.step, d.data, d.step, s.cols, s.rows, detail::device());
    else { CV_Error(SPX_ERR_BADARG, "cvtColor8u: only COLOR_BGR2Lab and COLOR_BGR2HSV are on the path"); return; }
    detail::check(rc, "cvtColor8u");
}

}}  // namespace spx::ximgproc

#ifndef SPX_XIMGPROC_NO_EXPORT_TO_CV
namespace cv { namespace ximgproc {
using spx::ximgproc::SLICType;
using spx::ximgproc::SLIC;
using spx::ximgproc::SLICO;
using spx::ximgproc::MSLIC;
using spx::ximgproc::SuperpixelSLIC;
using spx::ximgproc::SuperpixelLSC;
using spx::ximgproc::SuperpixelSEEDS;
using spx::ximgproc::createSuperpixelSLIC;
using spx::ximgproc::createSuperpixelLSC;
using spx::ximgproc::createSuperpixelSEEDS;
}}
#endif
