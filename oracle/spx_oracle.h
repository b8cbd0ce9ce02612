/* spx_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the algorithms the reference GUI reaches through
 * cv::cvtColor and cv::ximgproc::SuperpixelSLIC / LSC / SEEDS
 * (reference call sites: mainwindow.cpp:2096-2097, 2105-2111, 2114-2120, 2123-2127).
 * The arithmetic lives in OpenCV + opencv_contrib (module ximgproc, "openCV 4.1",
 * README.md:46), which is NOT vendored under /root/reference and not installed in
 * this image, so this file restates the published algorithms.
 *
 * Pinning status (see DESIGN.md "Oracle"):
 *   - BGR->Lab / BGR->HSV 8-bit          : pinned exhaustively (2^24 colours) against cv2 4.13.
 *   - SLIC(100) create/iterate, enforceLabelConnectivity, getLabelContourMask:
 *                                          pinned by the reference's shipped fixture
 *                                          example/logo-absurdephoton-segmentation-{image,grid,mask}.png.
 *   - SLICO, MSLIC, LSC, SEEDS            : PARITY UNPINNED (no fixture, no ximgproc binary).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may call into this library.
 */
#ifndef SPX_ORACLE_H
#define SPX_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* number of OpenMP threads the oracle may use (1 = scalar, faithful accumulation order) */
void spxo_set_threads(int n);
int  spxo_get_threads(void);

/* cv::cvtColor(..., COLOR_BGR2Lab / COLOR_BGR2HSV) on 8UC3 (mainwindow.cpp:2096-2097) */
void spxo_bgr2lab8(const uint8_t* src, size_t sstride, uint8_t* dst, size_t dstride, int w, int h);
void spxo_bgr2hsv8(const uint8_t* src, size_t sstride, uint8_t* dst, size_t dstride, int w, int h);

/* ---- SuperpixelSLIC (SLIC=100, SLICO=101, MSLIC=102), mainwindow.cpp:2105-2111 ---- */
typedef struct spxo_slic spxo_slic;
spxo_slic* spxo_slic_create(const uint8_t* img, size_t stride, int w, int h, int channels,
                            int algorithm, int region_size, float ruler);
void spxo_slic_destroy(spxo_slic*);
void spxo_slic_iterate(spxo_slic*, int num_iterations);
void spxo_slic_enforce_label_connectivity(spxo_slic*, int min_element_size);
int  spxo_slic_get_number_of_superpixels(const spxo_slic*);
void spxo_slic_get_labels(const spxo_slic*, int32_t* dst);           /* dense w*h */
void spxo_slic_set_labels(spxo_slic*, const int32_t* src, int numlabels);
/* centres: K rows of (c0..c[C-1], x, y) float32 */
void spxo_slic_get_centres(const spxo_slic*, float* dst);
void spxo_slic_get_label_contour_mask(const spxo_slic*, uint8_t* dst, int thick_line);

/* stand-alone stages on an arbitrary label map (used for bit-exact injected-label tests) */
int  spxo_enforce_label_connectivity(int32_t* labels, int w, int h, int numlabels, int min_element_size);
void spxo_label_contour_mask(const int32_t* labels, int w, int h, int thick_line, uint8_t* dst);       /* SLIC/LSC flavour */
void spxo_label_contour_mask_seeds(const int32_t* labels, int w, int h, int thick_line, uint8_t* dst); /* SEEDS flavour */

/* ---- SuperpixelLSC, mainwindow.cpp:2114-2120 (PARITY UNPINNED) ---- */
typedef struct spxo_lsc spxo_lsc;
spxo_lsc* spxo_lsc_create(const uint8_t* img, size_t stride, int w, int h, int channels,
                          int region_size, float ratio);
void spxo_lsc_destroy(spxo_lsc*);
void spxo_lsc_iterate(spxo_lsc*, int num_iterations);
void spxo_lsc_enforce_label_connectivity(spxo_lsc*, int min_element_size);
int  spxo_lsc_get_number_of_superpixels(const spxo_lsc*);
void spxo_lsc_get_labels(const spxo_lsc*, int32_t* dst);
void spxo_lsc_set_labels(spxo_lsc*, const int32_t* src, int numlabels);
/* centres: K rows of (2C+4 feature means, seed x, seed y) float32 */
void spxo_lsc_get_centres(const spxo_lsc*, float* dst);
/* post-pass visiting order: 0 (default) = ascending region index (authors' order), 1 = hashed index (CUDA default) */
void spxo_lsc_set_merge_order(spxo_lsc*, int order);
void spxo_lsc_set_merge_algo(spxo_lsc*, int algo);      /* 1: fixed-point evaluation of the sequential rule (same output) */
int  spxo_lsc_merge_iterations(const spxo_lsc*);
void spxo_lsc_get_label_contour_mask(const spxo_lsc*, uint8_t* dst, int thick_line);

/* ---- SuperpixelSEEDS, mainwindow.cpp:2123-2127 (PARITY UNPINNED; sizing pinned by screenshot) ---- */
typedef struct spxo_seeds spxo_seeds;
spxo_seeds* spxo_seeds_create(int w, int h, int channels, int num_superpixels, int num_levels,
                              int prior, int histogram_bins, int double_step);
void spxo_seeds_destroy(spxo_seeds*);
/* visiting order: 0 = upstream's sequential raster order, 1 = the phase-parallel order of the CUDA path */
void spxo_seeds_set_order(spxo_seeds*, int order);
/* colour-homogeneity energy of a labelling of the image last given to iterate() (higher is better) */
double spxo_seeds_energy(const spxo_seeds*, const int32_t* labels);
void spxo_seeds_iterate(spxo_seeds*, const uint8_t* img, size_t stride, int num_iterations);
int  spxo_seeds_get_number_of_superpixels(const spxo_seeds*);
void spxo_seeds_get_labels(const spxo_seeds*, int32_t* dst);
void spxo_seeds_get_label_contour_mask(const spxo_seeds*, uint8_t* dst, int thick_line);

#ifdef __cplusplus
}
#endif
#endif
