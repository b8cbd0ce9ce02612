/* spx_b200.h -- C ABI of the B200-native superpixel engine (libspx_b200.so).
 *
 * Drop-in boundary for the ONE hot path of AbsurdePhoton/superpixels-segmentation-gui-opencv:
 * MainWindow::on_button_compute_clicked (mainwindow.cpp:2062-2151), i.e. cv::cvtColor(BGR2Lab|BGR2HSV)
 * followed by cv::ximgproc::SuperpixelSLIC / SuperpixelLSC / SuperpixelSEEDS.
 * Each entry point names the reference interface it replaces.  The cv::ximgproc-shaped C++ adapter
 * over this ABI is include/spx_ximgproc.hpp; the binding a maintainer adds is in INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; strides are in BYTES; images are 8-bit interleaved (CV_8UC<n>),
 *     label maps are int32 (CV_32SC1), masks are uint8 {0,255} (CV_8UC1).
 *   - every function returns an int status: 0 = ok, negative = error (OpenCV-style codes below);
 *     nothing throws or aborts across the boundary; spx_last_error() gives the message (thread-local).
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     SPX_ERR_GPU.
 *   - "_dev" variants take/return DEVICE pointers and enqueue on the object's stream without
 *     synchronising (used by the batch path, where frames are resident in HBM).
 *   - an object is bound to one device and one stream; calls on one object must be serialised by
 *     the caller (same contract as the reference: single GUI thread, create -> iterate ->
 *     [enforceLabelConnectivity] -> getters).
 */
#ifndef SPX_B200_H
#define SPX_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPX_OK            0
#define SPX_ERR_BADARG   (-5)    /* cv::Error::StsBadArg   */
#define SPX_ERR_ASSERT   (-215)  /* cv::Error::StsAssert   */
#define SPX_ERR_NOMEM    (-4)    /* cv::Error::StsNoMem    */
#define SPX_ERR_GPU      (-217)  /* cv::Error::GpuApiCallError */
#define SPX_ERR_INTERNAL (-3)    /* cv::Error::StsInternal */

/* cv::ximgproc::SLICType (slic.hpp), passed by the GUI as algorithm+SLIC (mainwindow.cpp:2105) */
#define SPX_SLIC   100
#define SPX_SLICO  101
#define SPX_MSLIC  102

int         spx_version(void);
const char* spx_last_error(void);
int         spx_device_count(int* count);
/* device memory of destroyed objects stays in the device's stream-ordered pool for the next object (the GUI creates a
 * new Superpixel object on every COMPUTE click); this returns it to the driver */
int         spx_release_cached_memory(void);

/* ---- cv::cvtColor(converted, converted, COLOR_BGR2Lab / COLOR_BGR2HSV), mainwindow.cpp:2096-2097 ----
 * 8UC3 -> 8UC3, bit-exact with OpenCV's 8-bit integer pipelines.  In-place (src == dst) is allowed. */
int spx_bgr2lab8(const uint8_t* src, size_t src_stride, uint8_t* dst, size_t dst_stride, int width, int height, int device);
int spx_bgr2hsv8(const uint8_t* src, size_t src_stride, uint8_t* dst, size_t dst_stride, int width, int height, int device);
int spx_bgr2lab8_dev(const void* d_src, size_t src_stride, void* d_dst, size_t dst_stride, int width, int height, void* cuda_stream);
int spx_bgr2hsv8_dev(const void* d_src, size_t src_stride, void* d_dst, size_t dst_stride, int width, int height, void* cuda_stream);

/* ---- cv::ximgproc::SuperpixelSLIC (mainwindow.h:170, mainwindow.cpp:2105-2111) ---- */
typedef struct spx_slic spx_slic_t;

/* createSuperpixelSLIC(image, algorithm, region_size, ruler) -- mainwindow.cpp:2105.
 * `image` is the already colour-converted 8-bit image (1..4 channels; the GUI passes 3). */
int spx_slic_create(const uint8_t* image, size_t stride, int width, int height, int channels,
                    int algorithm, int region_size, float ruler, int device, spx_slic_t** out);
/* same, image already resident on the current CUDA device; work is enqueued on `cuda_stream`
 * (NULL = a private non-blocking stream owned by the object).  ORDERING: the object reads d_image on ITS stream.  With
 * cuda_stream = NULL nothing orders that read against the stream that produced the image: the caller must have synchronised
 * the producer (stream / event / device synchronise) before the call; with a caller-supplied stream, stream order does it.
 * The same holds for spx_slic_reset_dev and for every frame handed to spx_slic_batch_dev. */
int spx_slic_create_dev(const void* d_image, size_t stride, int width, int height, int channels,
                        int algorithm, int region_size, float ruler, void* cuda_stream, spx_slic_t** out);
/* re-run the constructor on a new frame of identical geometry/parameters, reusing all device
 * allocations (what `slic = createSuperpixelSLIC(...)` does on every COMPUTE click, minus malloc). */
/* the same with the GUI's cvtColor(converted, converted, COLOR_BGR2Lab | COLOR_BGR2HSV) (mainwindow.cpp:2096-2097) folded in: the
 * BGR image is uploaded once and converted on the device (colour_code = 44 / 40), saving a host round trip of the image */
int spx_slic_create_bgr(const uint8_t* bgr, size_t stride, int width, int height, int colour_code, int algorithm, int region_size,
                        float ruler, int device, spx_slic_t** out);
int spx_slic_reset(spx_slic_t*, const uint8_t* image, size_t stride);
int spx_slic_reset_dev(spx_slic_t*, const void* d_image, size_t stride);
/* SuperpixelSLIC::iterate(num_iterations) -- mainwindow.cpp:2106 */
int spx_slic_iterate(spx_slic_t*, int num_iterations);
/* SuperpixelSLIC::enforceLabelConnectivity(min_element_size) -- mainwindow.cpp:2108 (0 = no-op) */
int spx_slic_enforce_label_connectivity(spx_slic_t*, int min_element_size);
/* SuperpixelSLIC::getNumberOfSuperpixels() -- mainwindow.cpp:2109 (synchronises) */
int spx_slic_get_number_of_superpixels(spx_slic_t*, int* out);
/* SuperpixelSLIC::getLabels(labels) -- mainwindow.cpp:2110; CV_32SC1 */
int spx_slic_get_labels(spx_slic_t*, int32_t* dst, size_t dst_stride);
int spx_slic_get_labels_dev(spx_slic_t*, void* d_dst, size_t dst_stride);
/* SuperpixelSLIC::getLabelContourMask(mask, thick_line) -- mainwindow.cpp:2111; CV_8UC1 {0,255} */
int spx_slic_get_label_contour_mask(spx_slic_t*, uint8_t* dst, size_t dst_stride, int thick_line);
int spx_slic_get_label_contour_mask_dev(spx_slic_t*, void* d_dst, size_t dst_stride, int thick_line);
/* test hooks: inject a label map (bit-exact checks of connectivity / contour), read the centres.
 * centres: rows of (c0..c[channels-1], x, y) float32; `rows` = capacity of dst in rows. */
int spx_slic_set_labels(spx_slic_t*, const int32_t* src, size_t src_stride, int numlabels);
/* on = 1: the object is one of several working concurrently on the device (other streams): use the centre-update kernel with the
 * fewest instructions instead of the one with the shortest latency (default, best for a single COMPUTE click).  Results are
 * identical.  spx_slic_batch_dev sets it for its pooled objects. */
int spx_slic_set_throughput_mode(spx_slic_t*, int on);
int spx_slic_get_centres(spx_slic_t*, float* dst, int rows);
int spx_slic_synchronize(spx_slic_t*);
/* number of kernels this object has launched so far (bench.py's gpu_launches) */
int spx_slic_launch_count(spx_slic_t*, long long* out);
/* measurement hook (bench.py roofline): average device time in ms of ONE launch of the dominant kernel (the fused
 * assignment + accumulation pass) over `reps` launches at the object's current state, timed with CUDA events on the
 * object's stream; when flush_l2 != 0 a 256 MB buffer is overwritten before every launch.  State is left unchanged. */
int spx_slic_time_assign(spx_slic_t*, int reps, int flush_l2, float* avg_ms);
/* measurement hook (bench.py roofline): re-seeds the object and runs iterate(iters) launch by launch (no graph), every
 * assignment launch between its own pair of CUDA events on the object's stream; assign_ms[i] = duration of the i-th
 * assignment launch averaged over `reps` runs (one extra warm-up run first).  Leaves the state iterate(iters) gives. */
int spx_slic_time_iterations(spx_slic_t*, int iters, int reps, int flush_l2, float* assign_ms);
/* cv::Ptr<SuperpixelSLIC> release (the GUI overwrites `slic` on every COMPUTE, mainwindow.cpp:2105) */
int spx_slic_destroy(spx_slic_t*);

/* Host-side waiting policy of this process on `device`: on = 1 makes threads that wait for the device sleep
 * (cudaDeviceScheduleBlockingSync) instead of spinning; for hosts with more pipeline threads than cores. */
int spx_set_blocking_sync(int device, int on);

/* ---- row-strip mode: ONE image over several GPUs (north-star: gigapixel images) ----
 * One process per GPU.  Every rank creates the object with the geometry of the WHOLE image but uploads only the rows it
 * needs: rows [row0,row1) of the image, of which it owns (assigns) [own_row0,own_row1) -- multiples of 32 -- plus two halo
 * rows on every inner side.  Cluster state is replicated; per iteration the host exchanges the integer accumulators:
 *     spx_slic_create_strip(...)            seeds whose grid position lies in the owned rows, zeros elsewhere
 *     [sum kx,ky,kc,ikx,iky over ranks]     -> every rank has all seeds     (NCCL all-reduce on the pointers below)
 *     spx_slic_strip_rebin()
 *     repeat: spx_slic_strip_assign()       assignment + accumulation over the owned tile rows
 *             [sum acc over ranks]          exact: integer sums (only clusters near a cut receive pixels from two ranks)
 *             spx_slic_strip_update()       centre update (identical on every rank) + re-binning
 *     spx_slic_get_label_rows()             the owned rows of the label map
 * The result is bit-identical to the single-GPU object.  SLIC (algorithm 100), 3 channels, region_size >= 7. */
typedef struct {
    void* kx; void* ky;      /* float[Kcap]            */
    void* kc;                /* float[channels][Kcap]  */
    void* ikx; void* iky;    /* int[Kcap]              */
    void* acc;               /* uint64[channels+3][Kcap]: sums of c0..c2, x, y, count */
    int K, Kcap, channels;
    void* labels;            /* int32 label map of the WHOLE image on this device (rows outside the owned band: unspecified) */
    int label_pitch;         /* elements per label row */
} spx_slic_state_t;
int spx_slic_create_strip(const uint8_t* rows, size_t stride, int width, int height, int channels, int algorithm,
                          int region_size, float ruler, int row0, int row1, int own_row0, int own_row1,
                          void* cuda_stream, spx_slic_t** out);
int spx_slic_device_state(spx_slic_t*, spx_slic_state_t* out);
int spx_slic_strip_rebin(spx_slic_t*);
int spx_slic_strip_assign(spx_slic_t*);
int spx_slic_strip_update(spx_slic_t*);
/* The exchange step over NVLink peer memory instead of an NCCL all-reduce (one process per GPU, same node):
 *     spx_slic_strip_peer_alloc()   allocates this rank's exchange buffer, returns its 64-byte CUDA IPC handle
 *     [all-gather the handles]      world x 64 bytes, any transport (torch.distributed in strips.py)
 *     spx_slic_strip_peer_open()    maps every peer's buffer
 *     repeat: spx_slic_strip_assign(); spx_slic_strip_exchange(); spx_slic_strip_update();
 * strip_exchange = the rank's kernels store its partial sums (only the index range of clusters its band touched) into every
 * peer's buffer, publish the range and a flag, wait for the peers' flags and add their sums (integers: bit-identical to the
 * all-reduce).  Every rank must destroy its object only
 * after all ranks have finished iterating (a barrier on the host side). */
/* Neighbour exchange (the north-star's "NVLink peer exchange of boundary cluster sums"): call spx_slic_strip_partition() BEFORE
 * spx_slic_strip_peer_alloc().  own_rows = {own_row0, own_row1} of every rank, in rank order.  The cluster state is then
 * PARTITIONED: a rank updates only the clusters whose seed row lies within 5 * region_size rows of its band and exchanges partial
 * sums only with rank - 1 / rank + 1, over the index range of the clusters both maintain (a few seed rows).  Exact while no
 * centre moves more than 4 * region_size rows from its seed row; the centre update checks that and a violation surfaces as
 * SPX_ERR_ASSERT at the next synchronising call (strips.py then reruns with the all-to-all exchange).  Returns SPX_ERR_BADARG
 * when the bands are too thin (three ranks would maintain one cluster).  info_out[4] = the maintained index range [k_lo, k_hi)
 * and the core range [core_lo, core_hi) of clusters whose seed row lies in this rank's own band (for gathering centres). */
int spx_slic_strip_partition(spx_slic_t*, const int* own_rows /* world x 2 */, int world, int rank, int* info_out /* 4 ints or NULL */);
/* the same arithmetic without an object or a device (host logic only): plan_out[4q..4q+3] = k_lo, k_hi, core_lo, core_hi of rank q */
int spx_slic_strip_plan(int width, int height, int region_size, const int* own_rows /* world x 2 */, int world, int* plan_out /* world x 4 */);
int spx_slic_strip_peer_alloc(spx_slic_t*, int world, int rank, void* ipc_handle_out /* 64 bytes */, size_t* bytes_out);
int spx_slic_strip_peer_open(spx_slic_t*, const void* ipc_handles /* world x 64 bytes, rank order */);
int spx_slic_strip_exchange(spx_slic_t*);
/* assign + exchange + update in one call (peer modes). */
int spx_slic_strip_iteration(spx_slic_t*);
int spx_slic_strip_exchanged_bytes(spx_slic_t*, unsigned long long* out);   /* bytes this rank has stored into its peers so far */
int spx_slic_get_label_rows(spx_slic_t*, int32_t* dst, size_t dst_stride, int row0, int row1);

/* ---- batch of independent frames on one device (BASELINE.json configs[4]) ----
 * For every frame i: createSuperpixelSLIC(frame_i) ; iterate(num_iterations) ;
 * [enforceLabelConnectivity(min_element_size) if > 0] ; getLabels -> d_labels[i]
 * ; [getLabelContourMask(thick_line) -> d_masks[i] if d_masks != NULL].
 * All pointers in the arrays are DEVICE pointers (dense: stride = width*channels, width*4, width).
 * `n_streams` engines work concurrently on private streams.  The frames must be complete in device memory when the call is
 * made (see spx_slic_create_dev: ORDERING).  Blocks until the batch is done. */
int spx_slic_batch_dev(const void* const* d_frames, int n_frames, int width, int height, int channels,
                       int algorithm, int region_size, float ruler, int num_iterations,
                       int min_element_size, void* const* d_labels, void* const* d_masks, int thick_line,
                       int n_streams, long long* launches_out);

/* frees the engines (device buffers, captured graphs) spx_slic_batch_dev keeps pooled between calls */
int spx_slic_batch_release(void);

/* ---- stand-alone stages on a caller-supplied label map (host buffers) ----
 * same arithmetic as the member functions above; used for bit-exact stage tests. */
int spx_enforce_label_connectivity(int32_t* labels, size_t stride, int width, int height,
                                   int numlabels, int min_element_size, int device, int* numlabels_out);
/* flavour 0 = SuperpixelSLIC/LSC::getLabelContourMask, 1 = SuperpixelSEEDS::getLabelContourMask */
int spx_label_contour_mask(const int32_t* labels, size_t stride, int width, int height, int thick_line,
                           int flavour, uint8_t* dst, size_t dst_stride, int device);

/* test hook: the tiled SLIC kernel replaces the IEEE division by the constant xywt=(S/ruler)^2 with a correctly rounded
 * multiply/FMA sequence; this counts inputs (out of n pseudo-random ones) where it differs from the division. */
int spx_selftest_division(float w, int n, int device, int* mismatches);

/* ---- cv::ximgproc::SuperpixelLSC (mainwindow.h:171, mainwindow.cpp:2114-2120) ---- */
typedef struct spx_lsc spx_lsc_t;
int spx_lsc_create(const uint8_t* image, size_t stride, int width, int height, int channels,
                   int region_size, float ratio, int device, spx_lsc_t** out);
int spx_lsc_create_bgr(const uint8_t* bgr, size_t stride, int width, int height, int colour_code, int region_size, float ratio, int device,
                       spx_lsc_t** out);
int spx_lsc_iterate(spx_lsc_t*, int num_iterations);
int spx_lsc_enforce_label_connectivity(spx_lsc_t*, int min_element_size);
int spx_lsc_get_number_of_superpixels(spx_lsc_t*, int* out);
int spx_lsc_get_labels(spx_lsc_t*, int32_t* dst, size_t dst_stride);
int spx_lsc_get_label_contour_mask(spx_lsc_t*, uint8_t* dst, size_t dst_stride, int thick_line);
int spx_lsc_set_labels(spx_lsc_t*, const int32_t* src, size_t src_stride, int numlabels);
/* test hooks: centres = rows of (2*channels+4 feature means, seed x, seed y) float32; kernel launch count */
int spx_lsc_get_centres(spx_lsc_t*, float* dst, int rows);
int spx_lsc_launch_count(spx_lsc_t*, long long* out);
/* visiting order of the small regions in enforceLabelConnectivity's merge pass: 0 (DEFAULT) = ascending region index, the
 * authors' order -- what upstream computes, and what the cv::ximgproc adapter therefore runs; 1 (opt-in) = hashed region
 * index: the same rule in a different order (final superpixel count differs by about 1 %), NOT upstream's result. */
int spx_lsc_set_merge_order(spx_lsc_t*, int order);
int spx_lsc_destroy(spx_lsc_t*);

/* ---- cv::ximgproc::SuperpixelSEEDS (mainwindow.h:172, mainwindow.cpp:2123-2127) ---- */
typedef struct spx_seeds spx_seeds_t;
/* histogram_bins: any value >= 1 (the GUI slider goes to 50); one histogram of bins^channels int counters is kept per block of
 * every level, as upstream does: SPX_ERR_NOMEM when that does not fit the free device memory. */
int spx_seeds_create(int width, int height, int channels, int num_superpixels, int num_levels,
                     int prior, int histogram_bins, int double_step, int device, spx_seeds_t** out);
/* SuperpixelSEEDS::iterate(img, num_iterations) -- mainwindow.cpp:2124 */
int spx_seeds_iterate(spx_seeds_t*, const uint8_t* image, size_t stride, int num_iterations);
int spx_seeds_get_number_of_superpixels(spx_seeds_t*, int* out);
int spx_seeds_get_labels(spx_seeds_t*, int32_t* dst, size_t dst_stride);
int spx_seeds_get_label_contour_mask(spx_seeds_t*, uint8_t* dst, size_t dst_stride, int thick_line);
/* number of kernels launched so far (test / bench hook) */
int spx_seeds_launch_count(spx_seeds_t*, long long* out);
int spx_seeds_destroy(spx_seeds_t*);

/* ---- the Segmentation tab's working set on the device (the callers either side of the path) ----
 * Replaces the full-size Mats the GUI keeps next to the image (mainwindow.h:176-185: labels, labels_mask CV_32SC1; mask, grid,
 * selection CV_8UC3) and the full-image passes it runs over them.  Plane ids: 0 image, 1 mask, 2 grid, 3 selection (8UC3, BGR),
 * 4 labels, 5 labels_mask (32SC1).  Every call is synchronous; results are bit-identical to the cv:: calls they replace
 * (addWeighted / dilate / compare / setTo / inRange / countNonZero / minMaxIdx of OpenCV 4.13). */
typedef struct spx_session spx_session_t;
int spx_session_create(int width, int height, int device, spx_session_t** out);
int spx_session_destroy(spx_session_t*);
int spx_session_upload(spx_session_t*, int plane, const void* src, size_t stride);      /* uploading plane 4 rebuilds the cell index */
int spx_session_download(spx_session_t*, int plane, void* dst, size_t stride);
/* the GUI's post-step after COMPUTE (mainwindow.cpp:2130-2138): labels <- getLabels(), labels_mask = 0, mask = 0,
 * selection = 0, grid = cvtColor(contour, GRAY2RGB) with setTo(colour, grid); builds the per-cell index */
int spx_session_begin(spx_session_t*, const int32_t* labels, size_t labels_stride, const uint8_t* contour_mask, size_t contour_stride,
                      const uint8_t colour[3]);
/* grid.setTo(colour, gray(grid) > 0) -- on_comboBox_grid_color_currentIndexChanged, mainwindow.cpp:1085-1104 */
int spx_session_recolour_grid(spx_session_t*, const uint8_t colour[3]);
/* Mat1b m = plane == value (plane 4 or 5) -- mainwindow.cpp:2159, 228, 258, 1334 */
int spx_session_eq_mask(spx_session_t*, int plane, int value, uint8_t* dst, size_t stride);
/* WhichCell(posX, posY), mainwindow.cpp:2153-2171: the cell under (x, y) is erased from mask and labels_mask (paint == 0) or
 * painted with `colour` and claimed for `label_id`; one pass over the cell's bounding box.  cell_out (optional): its label */
int spx_session_which_cell(spx_session_t*, int x, int y, int paint, const uint8_t colour[3], int label_id, int* cell_out);
/* where plane `key_plane` (4 / 5) == value: colour plane p3 (1..3, or -1) <- colour, int plane p1 (4 / 5, or -1) <- newval:
 * label delete (mainwindow.cpp:228-229), join (258-260), colour change (1334-1335) */
int spx_session_set_where_eq(spx_session_t*, int key_plane, int value, int p3, const uint8_t colour[3], int p1, int newval);
/* "update the cells from the white mask" (mainwindow.cpp:313-331): pure-white mask pixels become cell max(labels)+1 */
int spx_session_new_cell_from_white(spx_session_t*, const uint8_t colour[3], int label_id, int* new_cell, int* hits);
/* minMaxIdx(labels) maximum -- mainwindow.cpp:313 */
int spx_session_max_label(spx_session_t*, int* out);
/* per-cell index: boxes[k] = x0, y0, x1, y1 (inclusive; x1 < x0: unused id), counts[k] = pixels; rows = capacity of both */
int spx_session_cell_index(spx_session_t*, int32_t* boxes, int32_t* counts, int rows, int* cells);
/* ShowSegmentation(), mainwindow.cpp:1936-1969, as one pass: disp = image*a, + mask*b, + grid*c (addWeighted, blend = slider
 * value 0..100), selection dilated by `dilation` copied over where non-zero, holes (pure-black mask pixels) in white;
 * dst = vh rows of vw BGR pixels; holes (optional) = countNonZero over the whole mask */
int spx_session_compose(spx_session_t*, int vx, int vy, int vw, int vh, int show_image, int blend_image, int show_mask, int blend_mask,
                        int show_grid, int blend_grid, int show_selection, int dilation, int show_holes, uint8_t* dst, size_t stride,
                        int* holes);

/* ---- the reference's session data file <base>-segmentation-data.xml (OpenCV FileStorage XML; mainwindow.cpp:673-706 writes it,
 * 775-820 reads it): GridColor, LabelsCount, LabelId<i> / LabelName<i> / LabelColor<i>, and the CV_32SC1 matrices Labels and
 * LabelsMask.  Host side only (no GPU, no OpenCV); files round-trip with cv::FileStorage. */
typedef struct { int id; char name[128]; uint8_t colour[3]; } spx_label_t;
int spx_session_xml_write(const char* path, int grid_colour, const spx_label_t* list, int n_labels, const int32_t* labels,
                          size_t labels_stride, const int32_t* labels_mask, size_t mask_stride, int width, int height);
/* labels / labels_mask may be NULL (size query: rows, cols, n_labels are still returned); list holds list_capacity entries */
int spx_session_xml_read(const char* path, int* grid_colour, spx_label_t* list, int list_capacity, int* n_labels, int32_t* labels,
                         size_t labels_stride, int32_t* labels_mask, size_t mask_stride, int* rows, int* cols);

/* ---- the pre-processing that feeds the path (MainWindow::on_button_apply_clicked, mainwindow.cpp:1979-2047) ----
 * flags: 1 equalize (EqualizeHistogram, mat-image-tools.cpp:184-202), 2 normalize (cv::normalize(1, 255, NORM_MINMAX),
 * mainwindow.cpp:2029), 4 colour balance (SimplestColorBalance(image, percent), mat-image-tools.cpp:204-229), 8 gaussian
 * blur (cv::GaussianBlur 3x3, mainwindow.cpp:2031); applied in that order, bit-identical to OpenCV 4.13's 8-bit arithmetic.
 * src / dst: 8UC3 (BGR) host images, may alias. */
int spx_preprocess_8uc3(const uint8_t* src, size_t src_stride, uint8_t* dst, size_t dst_stride, int width, int height, int flags,
                        int percent, int device);

#ifdef __cplusplus
}
#endif
#endif /* SPX_B200_H */
