// compute_click.cpp -- the reference GUI's COMPUTE click (mainwindow.cpp:2062-2151), verbatim in structure, compiled
// against include/spx_ximgproc.hpp instead of <opencv2/ximgproc.hpp>.  Shows that the adapter is a drop-in at source
// level: the body below uses only cv::ximgproc names and cv::Mat.  With a GPU it prints the cell counts; without one
// the constructors throw cv::Exception(-217) exactly like a failing CV_Assert would, and the program reports that.
#include <opencv2/core.hpp>
#include "spx_ximgproc.hpp"
#include <cstdio>
#include <cstdlib>

using namespace cv;
using namespace cv::ximgproc;

static Mat synthetic_bgr(int w, int h)
{
    Mat img(h, w, CV_8UC3);
    unsigned s = 12345u;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int cell = ((x / 37) * 7 + (y / 29) * 13) & 7;
            for (int c = 0; c < 3; c++) {
                s = s * 1664525u + 1013904223u;
                img.data[(size_t)y * img.step + 3 * x + c] = (unsigned char)(40 + cell * 24 + c * 10 + (s >> 29));
            }
        }
    return img;
}

static int compute(const Mat& image, int algorithm, bool lab)
{
    // mainwindow.cpp:2070-2081 -- parameters as read from the widgets (defaults of mainwindow.ui)
    int region_size = 30, ruler = 30, min_element_size = 25, num_iterations = 10;
    int num_superpixels = 1000, num_levels = 4, prior = 2, num_histogram_bins = 5;
    bool double_step = false, thick_lines = true;
    double ratio = 0.075;
    Mat converted, labels, maskTemp;
    int nb_cells = 0;
    spx::ximgproc::cvtColor8u(image, converted, lab ? 44 : 40);             // :2096-2097 (cv::cvtColor in the GUI)
    if (algorithm <= 2) {                                                    // :2104-2112
        Ptr<SuperpixelSLIC> slic = createSuperpixelSLIC(converted, algorithm + SLIC, region_size, float(ruler));
        slic->iterate(num_iterations);
        if (min_element_size > 0) slic->enforceLabelConnectivity(min_element_size);
        nb_cells = slic->getNumberOfSuperpixels();
        slic->getLabels(labels);
        slic->getLabelContourMask(maskTemp, !thick_lines);
    } else if (algorithm == 3) {                                             // :2113-2121
        Ptr<SuperpixelLSC> lsc = createSuperpixelLSC(converted, region_size, float(ratio));
        lsc->iterate(num_iterations);
        if (min_element_size > 0) lsc->enforceLabelConnectivity(min_element_size);
        nb_cells = lsc->getNumberOfSuperpixels();
        lsc->getLabels(labels);
        lsc->getLabelContourMask(maskTemp, !thick_lines);
    } else {                                                                 // :2122-2128
        Ptr<SuperpixelSEEDS> seeds = createSuperpixelSEEDS(converted.cols, converted.rows, converted.channels(),
                                                           num_superpixels, num_levels, prior, num_histogram_bins, double_step);
        seeds->iterate(converted, num_iterations);
        nb_cells = seeds->getNumberOfSuperpixels();
        seeds->getLabels(labels);
        seeds->getLabelContourMask(maskTemp, !thick_lines);
    }
    long contour = 0;
    int maxlabel = 0;
    for (int y = 0; y < labels.rows; y++)
        for (int x = 0; x < labels.cols; x++) {
            contour += maskTemp.data[(size_t)y * maskTemp.step + x] != 0;
            int l = labels.at<int>(y, x);
            if (l > maxlabel) maxlabel = l;
        }
    std::printf("algorithm %d lab %d: nb_cells %d max_label %d contour_px %ld\n", algorithm, (int)lab, nb_cells, maxlabel, contour);
    return (maxlabel < nb_cells && nb_cells > 0) ? 0 : 1;
}

int main(int argc, char** argv)
{
    Mat image = synthetic_bgr(640, 480);
    int bad = 0;
    try {
        for (int algorithm = 0; algorithm <= 4; algorithm++) {
            if (argc > 1 && std::atoi(argv[1]) != algorithm) continue;
            bad += compute(image, algorithm, true);
        }
        bad += compute(image, 0, false);
    } catch (const cv::Exception& e) {
        std::printf("cv::Exception %d: %s\n", e.code, e.what());
        return e.code == SPX_ERR_GPU ? 3 : 2;      // 3 = no CUDA device (expected on a CPU-only box)
    }
    return bad ? 1 : 0;
}
