// session.cu -- the Segmentation tab's working set kept on the GPU, and the passes the GUI runs over it after COMPUTE
// (SURVEY.md section 8f, ranks 1 and 2).
//
// The reference keeps five full-size Mats next to the image (mainwindow.h:176-185): labels, labels_mask (CV_32SC1), mask, grid,
// selection (CV_8UC3), and recomputes a full-image `labels == label` compare plus two masked setTo passes on every click or
// drag step (WhichCell, mainwindow.cpp:2153-2171), and up to five full-viewport addWeighted / dilate / inRange passes on
// every refresh (ShowSegmentation, mainwindow.cpp:1936-1969).  Here the planes live in HBM; a click is ONE pass restricted
// to the cell's bounding box (per-cell index built once per label map), a refresh is ONE fused pass over the viewport.
//
// Arithmetic pinned against cv2 4.13 (tests/test_session.py): cv::addWeighted on 8-bit data evaluates
// fma(src1, alpha, fl(src2*beta + gamma)) in float32 and rounds half to even (cvRound) with saturation; cv::dilate with a
// (2d+1)^2 rectangle is a max filter that ignores pixels outside the (viewport) image; compare gives 0 / 255.
#include "common.cuh"
#include <climits>
#include <vector>

namespace spx {

struct Session {
    int device = 0, w = 0, h = 0;
    cudaStream_t st = nullptr;
    int* labels = nullptr;           // [h][w]
    int* labels_mask = nullptr;      // [h][w]
    unsigned char* planes[4] = {nullptr, nullptr, nullptr, nullptr};   // image, mask, grid, selection: 8UC3, pitch 3w
    unsigned char* scratch = nullptr;                                   // h*w bytes (compare results, contour upload)
    unsigned char* out = nullptr;                                       // composed viewport (at most h*w*3)
    int* box = nullptr;              // per cell: x0, y0, x1, y1 (inclusive), [K][4]
    int* count = nullptr;            // per cell: pixels
    int* d_flag = nullptr;           // [4]
    int* h_flag = nullptr;           // pinned [4]
    int K = 0;                       // cells indexed (max label + 1), 0: no index
    std::vector<int> h_box;          // host copy of the boxes (a click needs one box to size its launch)
    ~Session()
    {
        cudaSetDevice(device);
        if (st) cudaStreamSynchronize(st);
        pool_free(labels, st); pool_free(labels_mask, st);
        for (auto p : planes) pool_free(p, st);
        pool_free(scratch, st); pool_free(out, st); pool_free(box, st); pool_free(count, st); pool_free(d_flag, st);
        if (h_flag) cudaFreeHost(h_flag);
        if (st) cudaStreamDestroy(st);
    }
};

// ---- kernels ---------------------------------------------------------------------------------------------------
// plane == value -> 0 / 255   (Mat1b m = labels == label, mainwindow.cpp:2159, 228, 258, 1334)
__global__ void eq_mask_kernel(const int* __restrict__ src, int n, int value, unsigned char* __restrict__ dst)
{
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < n) {
        const int4 v = *reinterpret_cast<const int4*>(src + i);
        uchar4 o;
        o.x = v.x == value ? 255 : 0; o.y = v.y == value ? 255 : 0; o.z = v.z == value ? 255 : 0; o.w = v.w == value ? 255 : 0;
        *reinterpret_cast<uchar4*>(dst + i) = o;
    } else
        for (int k = i; k < n; k++) dst[k] = src[k] == value ? 255 : 0;
}
// inside the rectangle [x0,x1] x [y0,y1]: where key == value: colour plane <- colour (when do3), int plane <- newval (when do1)
__global__ void set_where_eq_kernel(const int* __restrict__ key, int w, int x0, int y0, int x1, int y1, int value,
                                    unsigned char* __restrict__ p3, int do3, uchar3 colour, int* __restrict__ p1, int do1, int newval)
{
    const int x = x0 + blockIdx.x * blockDim.x + threadIdx.x, y = y0 + blockIdx.y;
    if (x > x1 || y > y1) return;
    const size_t i = (size_t)y * w + x;
    if (key[i] != value) return;
    if (do3) { p3[3 * i] = colour.x; p3[3 * i + 1] = colour.y; p3[3 * i + 2] = colour.z; }
    if (do1) p1[i] = newval;
}
// mask-driven setTo (labels.setTo(maxLabel, white_mask), mainwindow.cpp:328-330): where the 8UC3 plane equals `match`
__global__ void set_where_colour_kernel(const unsigned char* __restrict__ key3, int n, uchar3 match, unsigned char* __restrict__ p3a, int doa,
                                        uchar3 ca, unsigned char* __restrict__ p3b, int dob, uchar3 cb, int* __restrict__ p1a, int do1a, int va,
                                        int* __restrict__ p1b, int do1b, int vb, int* __restrict__ hits)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool hit = false;
    if (i < n) hit = key3[3 * i] == match.x && key3[3 * i + 1] == match.y && key3[3 * i + 2] == match.z;
    if (hit) {
        if (do1a) p1a[i] = va;
        if (do1b) p1b[i] = vb;
        if (dob) { p3b[3 * i] = cb.x; p3b[3 * i + 1] = cb.y; p3b[3 * i + 2] = cb.z; }
        if (doa) { p3a[3 * i] = ca.x; p3a[3 * i + 1] = ca.y; p3a[3 * i + 2] = ca.z; }      // may be the key plane itself: written last
    }
    const unsigned b = __ballot_sync(0xffffffffu, hit);
    if (hits && (threadIdx.x & 31) == 0 && b) atomicAdd(hits, __popc(b));
}
// max of a plane (minMaxIdx(labels), mainwindow.cpp:313)
__global__ void max_kernel(const int* __restrict__ src, int n, int* __restrict__ out)
{
    int m = INT_MIN;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = max(m, src[i]);
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}
// per-cell bounding boxes and pixel counts; a warp covers 32 consecutive pixels of a row (1-3 cells), so the atomics are
// issued once per run of equal labels
__global__ void cell_index_kernel(const int* __restrict__ lab, int w, int h, int K, int* __restrict__ box, int* __restrict__ count)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const bool in = x < w;
    const int l = in ? lab[(size_t)y * w + x] : -1;
    const int prev = __shfl_up_sync(0xffffffffu, l, 1);
    const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || prev != l);
    const bool head = (heads >> lane) & 1u;
    if (!head || l < 0 || l >= K) return;
    const unsigned after = lane == 31 ? 0u : heads >> (lane + 1);
    const int len = after ? __ffs(after) : 32 - lane;             // run length within the warp
    const int xe = min(x + len - 1, w - 1);
    atomicMin(&box[4 * l], x); atomicMin(&box[4 * l + 1], y); atomicMax(&box[4 * l + 2], xe); atomicMax(&box[4 * l + 3], y);
    atomicAdd(&count[l], xe - x + 1);
}
__global__ void box_init_kernel(int* box, int* count, int K, int w, int h)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    box[4 * k] = w; box[4 * k + 1] = h; box[4 * k + 2] = -1; box[4 * k + 3] = -1;
    count[k] = 0;
}
// contour mask (CV_8UC1 0/255) -> grid: cvtColor(GRAY2RGB) then setTo(colour, grid)  (mainwindow.cpp:2133-2136)
__global__ void grid_from_contour_kernel(const unsigned char* __restrict__ m, int n, uchar3 colour, unsigned char* __restrict__ grid)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool on = m[i] != 0;
    grid[3 * i] = on ? colour.x : 0; grid[3 * i + 1] = on ? colour.y : 0; grid[3 * i + 2] = on ? colour.z : 0;
}
// grid.setTo(colour, gray(grid) > 0)  (mainwindow.cpp:1099-1103): RGB2GRAY of a non-black pixel can only round to 0 when the
// weighted sum is below 0.5, so the exact 8-bit conversion is evaluated: (R*9798 + G*19235 + B*3735 + 16384) >> 15 with the
// plane read as RGB (the GUI passes COLOR_RGB2GRAY on its BGR-ordered plane: channel 0 takes the R weight)
__global__ void grid_recolour_kernel(unsigned char* __restrict__ grid, int n, uchar3 colour)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c0 = grid[3 * i], c1 = grid[3 * i + 1], c2 = grid[3 * i + 2];
    const int gray = (c0 * 9798 + c1 * 19235 + c2 * 3735 + 16384) >> 15;
    if (gray > 0) { grid[3 * i] = colour.x; grid[3 * i + 1] = colour.y; grid[3 * i + 2] = colour.z; }
}

// ---- ShowSegmentation (mainwindow.cpp:1936-1969) as one pass over the viewport -------------------------------
struct ComposeArgs {
    const unsigned char *image, *mask, *grid, *selection;
    int w, h, vx, vy, vw, vh;
    int show_image, show_mask, show_grid, show_selection, show_holes, dilation;
    float a_image, a_mask, a_grid;
};
// cv::addWeighted(d, alpha, s, beta, 0) on 8-bit data: fma(d, alpha, fl(s*beta)) in float32, cvRound, saturate
__device__ __forceinline__ int aw8(int d, float alpha, int s, float beta)
{
    const float v = __fmaf_rn((float)d, alpha, __fmul_rn((float)s, beta));
    return min(max(__float2int_rn(v), 0), 255);
}
__global__ void compose_kernel(ComposeArgs a, unsigned char* __restrict__ out)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= a.vw || y >= a.vh) return;
    const size_t src = ((size_t)(a.vy + y) * a.w + a.vx + x) * 3;
    int d[3] = {0, 0, 0};
    if (a.show_image)
#pragma unroll
        for (int c = 0; c < 3; c++) d[c] = aw8(0, 1.0f - a.a_image, a.image[src + c], a.a_image);
    int m[3] = {0, 0, 0};
    if (a.show_mask || a.show_holes)
#pragma unroll
        for (int c = 0; c < 3; c++) m[c] = a.mask[src + c];
    if (a.show_mask)
#pragma unroll
        for (int c = 0; c < 3; c++) d[c] = aw8(d[c], 1.0f, m[c], a.a_mask);
    if (a.show_grid)
#pragma unroll
        for (int c = 0; c < 3; c++) d[c] = aw8(d[c], 1.0f, a.grid[src + c], a.a_grid);
    if (a.show_selection) {
        // DilatePixels(selection(viewport), dilation) then copyTo(disp, selection): per channel, non-zero values overwrite
        int s[3] = {0, 0, 0};
        const int r = a.dilation > 0 ? a.dilation : 0;
        for (int yy = max(y - r, 0); yy <= min(y + r, a.vh - 1); yy++)
            for (int xx = max(x - r, 0); xx <= min(x + r, a.vw - 1); xx++) {
                const size_t q = ((size_t)(a.vy + yy) * a.w + a.vx + xx) * 3;
#pragma unroll
                for (int c = 0; c < 3; c++) s[c] = max(s[c], (int)a.selection[q + c]);
            }
#pragma unroll
        for (int c = 0; c < 3; c++) if (s[c]) d[c] = s[c];
    }
    if (a.show_holes && m[0] == 0 && m[1] == 0 && m[2] == 0)        // inRange(mask, 0, 0) drawn white, added with weight 1
#pragma unroll
        for (int c = 0; c < 3; c++) d[c] = aw8(d[c], 1.0f, 255, 1.0f);
    const size_t o = ((size_t)y * a.vw + x) * 3;
    out[o] = (unsigned char)d[0]; out[o + 1] = (unsigned char)d[1]; out[o + 2] = (unsigned char)d[2];
}
// countNonZero(inRange(mask, 0, 0)) over the WHOLE mask (mainwindow.cpp:1957-1966)
__global__ void holes_kernel(const unsigned char* __restrict__ mask, int n, int* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool hole = i < n && mask[3 * i] == 0 && mask[3 * i + 1] == 0 && mask[3 * i + 2] == 0;
    const unsigned b = __ballot_sync(0xffffffffu, hole);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(out, __popc(b));
}

static int build_index(Session* s)
{
    const int n = s->w * s->h;
    SPX_CUDA(cudaMemsetAsync(s->d_flag, 0x80, sizeof(int), s->st));                   // very negative
    max_kernel<<<148 * 4, 256, 0, s->st>>>(s->labels, n, s->d_flag);
    SPX_CUDA(cudaMemcpyAsync(s->h_flag, s->d_flag, sizeof(int), cudaMemcpyDeviceToHost, s->st));
    SPX_CUDA(cudaStreamSynchronize(s->st));
    const int K = s->h_flag[0] + 1;
    SPX_REQUIRE(K >= 1 && K <= n, SPX_ERR_BADARG, "session: labels must lie in [0, width*height)");
    pool_free(s->box, s->st); pool_free(s->count, s->st);
    s->box = nullptr; s->count = nullptr;
    SPX_CHECK(pool_alloc((void**)&s->box, (size_t)K * 4 * sizeof(int), s->st));
    SPX_CHECK(pool_alloc((void**)&s->count, (size_t)K * sizeof(int), s->st));
    box_init_kernel<<<cdiv(K, 256), 256, 0, s->st>>>(s->box, s->count, K, s->w, s->h);
    cell_index_kernel<<<dim3(cdiv(s->w, 256), s->h), 256, 0, s->st>>>(s->labels, s->w, s->h, K, s->box, s->count);
    SPX_CUDA(cudaGetLastError());
    s->h_box.resize((size_t)K * 4);
    SPX_CUDA(cudaMemcpyAsync(s->h_box.data(), s->box, (size_t)K * 4 * sizeof(int), cudaMemcpyDeviceToHost, s->st));
    SPX_CUDA(cudaStreamSynchronize(s->st));
    s->K = K;
    return SPX_OK;
}

}  // namespace spx

using spx::Session;
struct spx_session { Session* s; };
#define SES_ENTER(h)                                                                  \
    SPX_REQUIRE((h) && (h)->s, SPX_ERR_BADARG, "NULL session handle");                \
    Session* s = (h)->s;                                                              \
    SPX_CUDA(cudaSetDevice(s->device))

extern "C" int spx_session_create(int w, int h, int device, spx_session_t** out) try
{
    using namespace spx;
    SPX_REQUIRE(out, SPX_ERR_BADARG, "session_create: out is NULL");
    *out = nullptr;
    SPX_REQUIRE(w > 0 && h > 0 && h <= SPX_MAX_HEIGHT && (long long)w * h < (1ll << 29), SPX_ERR_BADARG, "session_create: bad size %dx%d", w, h);
    SPX_CHECK(require_device(device));
    SPX_CUDA(cudaSetDevice(device));
    Session* s = new Session();
    s->device = device; s->w = w; s->h = h;
    const size_t n = (size_t)w * h;
    int rc = SPX_OK;
    if (cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking) != cudaSuccess) { set_error("session_create: stream"); rc = SPX_ERR_GPU; }
    if (rc == SPX_OK) rc = pool_alloc((void**)&s->labels, n * sizeof(int), s->st);
    if (rc == SPX_OK) rc = pool_alloc((void**)&s->labels_mask, n * sizeof(int), s->st);
    for (int p = 0; p < 4 && rc == SPX_OK; p++) rc = pool_alloc((void**)&s->planes[p], n * 3, s->st);
    if (rc == SPX_OK) rc = pool_alloc((void**)&s->scratch, n, s->st);
    if (rc == SPX_OK) rc = pool_alloc((void**)&s->out, n * 3, s->st);
    if (rc == SPX_OK) rc = pool_alloc((void**)&s->d_flag, 4 * sizeof(int), s->st);
    if (rc == SPX_OK && cudaMallocHost(&s->h_flag, 4 * sizeof(int)) != cudaSuccess) { set_error("session_create: pinned flags"); rc = SPX_ERR_GPU; }
    if (rc == SPX_OK) {
        cudaMemsetAsync(s->labels, 0, n * sizeof(int), s->st); cudaMemsetAsync(s->labels_mask, 0, n * sizeof(int), s->st);
        for (int p = 0; p < 4; p++) cudaMemsetAsync(s->planes[p], 0, n * 3, s->st);
        if (cudaStreamSynchronize(s->st) != cudaSuccess) { set_error("session_create: clearing the planes failed"); rc = SPX_ERR_GPU; }
    }
    if (rc != SPX_OK) { delete s; return rc; }
    *out = new spx_session{s};
    return SPX_OK;
} SPX_API_CATCH
extern "C" int spx_session_destroy(spx_session_t* h) try
{
    if (!h) return SPX_OK;
    delete h->s;
    delete h;
    return SPX_OK;
} SPX_API_CATCH
// plane ids: 0 image, 1 mask, 2 grid, 3 selection (8UC3); 4 labels, 5 labels_mask (32SC1)
extern "C" int spx_session_upload(spx_session_t* h, int plane, const void* src, size_t stride) try
{
    SES_ENTER(h);
    SPX_REQUIRE(src && plane >= 0 && plane <= 5, SPX_ERR_BADARG, "session_upload: bad plane / source");
    const size_t rowb = plane < 4 ? (size_t)s->w * 3 : (size_t)s->w * 4;
    SPX_REQUIRE(stride >= rowb, SPX_ERR_BADARG, "session_upload: stride too small");
    void* dst = plane < 4 ? (void*)s->planes[plane] : (plane == 4 ? (void*)s->labels : (void*)s->labels_mask);
    SPX_CUDA(cudaMemcpy2DAsync(dst, rowb, src, stride, rowb, s->h, cudaMemcpyHostToDevice, s->st));
    SPX_CUDA(cudaStreamSynchronize(s->st));
    if (plane == 4) SPX_CHECK(spx::build_index(s));
    return SPX_OK;
} SPX_API_CATCH
extern "C" int spx_session_download(spx_session_t* h, int plane, void* dst, size_t stride) try
{
    SES_ENTER(h);
    SPX_REQUIRE(dst && plane >= 0 && plane <= 5, SPX_ERR_BADARG, "session_download: bad plane / destination");
    const size_t rowb = plane < 4 ? (size_t)s->w * 3 : (size_t)s->w * 4;
    SPX_REQUIRE(stride >= rowb, SPX_ERR_BADARG, "session_download: stride too small");
    const void* src = plane < 4 ? (const void*)s->planes[plane] : (plane == 4 ? (const void*)s->labels : (const void*)s->labels_mask);
    SPX_CUDA(cudaMemcpy2DAsync(dst, stride, src, rowb, rowb, s->h, cudaMemcpyDeviceToHost, s->st));
    SPX_CUDA(cudaStreamSynchronize(s->st));
    return SPX_OK;
} SPX_API_CATCH
// the GUI's post-step after COMPUTE (mainwindow.cpp:2130-2138): new labels, labels_mask = 0, mask = 0, selection = 0,
// grid = contour mask in `colour`
extern "C" int spx_session_begin(spx_session_t* h, const int32_t* labels, size_t lstride, const uint8_t* contour, size_t cstride,
                                 const uint8_t colour[3]) try
{
    using namespace spx;
    SES_ENTER(h);
    SPX_REQUIRE(labels && contour && colour && lstride >= (size_t)s->w * 4 && cstride >= (size_t)s->w, SPX_ERR_BADARG, "session_begin: bad argument");
    const size_t n = (size_t)s->w * s->h;
    SPX_CUDA(cudaMemcpy2DAsync(s->labels, (size_t)s->w * 4, labels, lstride, (size_t)s->w * 4, s->h, cudaMemcpyHostToDevice, s->st));
    SPX_CUDA(cudaMemcpy2DAsync(s->scratch, s->w, contour, cstride, s->w, s->h, cudaMemcpyHostToDevice, s->st));
    SPX_CUDA(cudaMemsetAsync(s->labels_mask, 0, n * sizeof(int), s->st));
    SPX_CUDA(cudaMemsetAsync(s->planes[1], 0, n * 3, s->st));
    SPX_CUDA(cudaMemsetAsync(s->planes[3], 0, n * 3, s->st));
    grid_from_contour_kernel<<<cdiv((int)n, 256), 256, 0, s->st>>>(s->scratch, (int)n, make_uchar3(colour[0], colour[1], colour[2]), s->planes[2]);
    SPX_CUDA(cudaGetLastError());
    return build_index(s);
} SPX_API_CATCH
extern "C" int spx_session_recolour_grid(spx_session_t* h, const uint8_t colour[3]) try
{
    using namespace spx;
    SES_ENTER(h);
    SPX_REQUIRE(colour, SPX_ERR_BADARG, "session_recolour_grid: colour is NULL");
    const int n = s->w * s->h;
    grid_recolour_kernel<<<cdiv(n, 256), 256, 0, s->st>>>(s->planes[2], n, make_uchar3(colour[0], colour[1], colour[2]));
    SPX_CUDA(cudaGetLastError());
    SPX_CUDA(cudaStreamSynchronize(s->st));
    return SPX_OK;
} SPX_API_CATCH
// Mat1b m = plane == value  (plane: 4 labels, 5 labels_mask)
extern "C" int spx_session_eq_mask(spx_session_t* h, int plane, int value, uint8_t* dst, size_t stride) try
{
    using namespace spx;
    SES_ENTER(h);
    SPX_REQUIRE(dst && (plane == 4 || plane == 5) && stride >= (size_t)s->w, SPX_ERR_BADARG, "session_eq_mask: bad argument");
    const int n = s->w * s->h;
    eq_mask_kernel<<<cdiv(cdiv(n, 4), 256), 256, 0, s->st>>>(plane == 4 ? s->labels : s->labels_mask, n, value, s->scratch);
    SPX_CUDA(cudaGetLastError());
    SPX_CUDA(cudaMemcpy2DAsync(dst, stride, s->scratch, s->w, s->w, s->h, cudaMemcpyDeviceToHost, s->st));
    SPX_CUDA(cudaStreamSynchronize(s->st));
    return SPX_OK;
} SPX_API_CATCH
// WhichCell (mainwindow.cpp:2153-2171): the cell under (x, y) is erased from mask and labels_mask, or (paint != 0) painted
// with `colour` and claimed for `label_id`.  One launch over the cell's bounding box.
extern "C" int spx_session_which_cell(spx_session_t* h, int x, int y, int paint, const uint8_t colour[3], int label_id, int* cell_out) try
{
    using namespace spx;
    SES_ENTER(h);
    SPX_REQUIRE(x >= 0 && x < s->w && y >= 0 && y < s->h, SPX_ERR_BADARG, "which_cell: (%d, %d) outside the image", x, y);
    SPX_REQUIRE(s->K > 0, SPX_ERR_BADARG, "which_cell: no labels (call session_begin first)");
    SPX_REQUIRE(!paint || colour, SPX_ERR_BADARG, "which_cell: colour is NULL");
    SPX_CUDA(cudaMemcpyAsync(s->h_flag, s->labels + (size_t)y * s->w + x, sizeof(int), cudaMemcpyDeviceToHost, s->st));
    SPX_CUDA(cudaStreamSynchronize(s->st));
    const int cell = s->h_flag[0];
    SPX_REQUIRE(cell >= 0 && cell < s->K, SPX_ERR_INTERNAL, "which_cell: label %d outside the index", cell);
    const int* b = &s->h_box[(size_t)cell * 4];
    const uchar3 c = paint ? make_uchar3(colour[0], colour[1], colour[2]) : make_uchar3(0, 0, 0);
    dim3 grid(cdiv(b[2] - b[0] + 1, 128), b[3] - b[1] + 1);
    set_where_eq_kernel<<<grid, 128, 0, s->st>>>(s->labels, s->w, b[0], b[1], b[2], b[3], cell, s->planes[1], 1, c, s->labels_mask, 1,
                                                 paint ? label_id : 0);
    SPX_CUDA(cudaGetLastError());
    SPX_CUDA(cudaStreamSynchronize(s->st));
    if (cell_out) *cell_out = cell;
    return SPX_OK;
} SPX_API_CATCH
// where key plane (4 labels / 5 labels_mask) == value: colour plane `p3` (1 mask, 2 grid, 3 selection, -1 none) <- colour,
// int plane `p1` (4, 5, -1 none) <- newval.  Covers the label list actions: delete (mainwindow.cpp:228-229), join
// (258-260), colour change (1334-1335).
extern "C" int spx_session_set_where_eq(spx_session_t* h, int key_plane, int value, int p3, const uint8_t colour[3], int p1, int newval) try
{
    using namespace spx;
    SES_ENTER(h);
    SPX_REQUIRE((key_plane == 4 || key_plane == 5) && (p3 == -1 || (p3 >= 1 && p3 <= 3)) && (p1 == -1 || p1 == 4 || p1 == 5) && (p3 == -1 || colour),
                SPX_ERR_BADARG, "set_where_eq: bad plane selection");
    const int* key = key_plane == 4 ? s->labels : s->labels_mask;
    int x0 = 0, y0 = 0, x1 = s->w - 1, y1 = s->h - 1;
    if (key_plane == 4 && s->K > 0 && value >= 0 && value < s->K) { const int* b = &s->h_box[(size_t)value * 4]; x0 = b[0]; y0 = b[1]; x1 = b[2]; y1 = b[3]; }
    if (x1 >= x0 && y1 >= y0) {
        const uchar3 c = p3 >= 0 ? make_uchar3(colour[0], colour[1], colour[2]) : make_uchar3(0, 0, 0);
        dim3 grid(cdiv(x1 - x0 + 1, 128), y1 - y0 + 1);
        set_where_eq_kernel<<<grid, 128, 0, s->st>>>(key, s->w, x0, y0, x1, y1, value, p3 >= 0 ? s->planes[p3] : nullptr, p3 >= 0 ? 1 : 0, c,
                                                     p1 == 4 ? s->labels : s->labels_mask, p1 >= 0 ? 1 : 0, newval);
        SPX_CUDA(cudaGetLastError());
    }
    SPX_CUDA(cudaStreamSynchronize(s->st));
    if (p1 == 4) SPX_CHECK(build_index(s));
    return SPX_OK;
} SPX_API_CATCH
// the "update cells from the white mask" action (mainwindow.cpp:313-331): every pixel of `mask` that is pure white becomes
// a new cell: labels <- max(labels)+1, labels_mask <- label_id, mask <- colour, grid <- 0 there (the new cell's outline
// is drawn by the caller: findContours/drawContours stay on the host).  hits = pixels updated.
extern "C" int spx_session_new_cell_from_white(spx_session_t* h, const uint8_t colour[3], int label_id, int* new_cell, int* hits) try
{
    using namespace spx;
    SES_ENTER(h);
    SPX_REQUIRE(colour && s->K > 0, SPX_ERR_BADARG, "new_cell_from_white: bad argument / no labels");
    const int n = s->w * s->h;
    const int cell = s->K;                       // max(labels) + 1
    SPX_CUDA(cudaMemsetAsync(s->d_flag + 1, 0, sizeof(int), s->st));
    set_where_colour_kernel<<<cdiv(n, 256), 256, 0, s->st>>>(s->planes[1], n, make_uchar3(255, 255, 255), s->planes[1], 1,
                                                             make_uchar3(colour[0], colour[1], colour[2]), s->planes[2], 1, make_uchar3(0, 0, 0),
                                                             s->labels, 1, cell, s->labels_mask, 1, label_id, s->d_flag + 1);
    SPX_CUDA(cudaGetLastError());
    SPX_CUDA(cudaMemcpyAsync(s->h_flag + 1, s->d_flag + 1, sizeof(int), cudaMemcpyDeviceToHost, s->st));
    SPX_CUDA(cudaStreamSynchronize(s->st));
    if (hits) *hits = s->h_flag[1];
    if (new_cell) *new_cell = cell;
    if (s->h_flag[1] > 0) SPX_CHECK(build_index(s));
    return SPX_OK;
} SPX_API_CATCH
extern "C" int spx_session_max_label(spx_session_t* h, int* out) try
{
    SES_ENTER(h);
    SPX_REQUIRE(out && s->K > 0, SPX_ERR_BADARG, "max_label: bad argument / no labels");
    *out = s->K - 1;
    return SPX_OK;
} SPX_API_CATCH
// per-cell index: boxes[k] = x0, y0, x1, y1 (inclusive; x1 < x0 for an unused id), counts[k] = pixels; rows = capacity
extern "C" int spx_session_cell_index(spx_session_t* h, int32_t* boxes, int32_t* counts, int rows, int* cells) try
{
    SES_ENTER(h);
    SPX_REQUIRE(s->K > 0, SPX_ERR_BADARG, "cell_index: no labels");
    if (cells) *cells = s->K;
    const int r = rows < s->K ? rows : s->K;
    if (boxes && r > 0) memcpy(boxes, s->h_box.data(), (size_t)r * 4 * sizeof(int));
    if (counts && r > 0) {
        SPX_CUDA(cudaMemcpyAsync(counts, s->count, (size_t)r * sizeof(int), cudaMemcpyDeviceToHost, s->st));
        SPX_CUDA(cudaStreamSynchronize(s->st));
    }
    return SPX_OK;
} SPX_API_CATCH
// ShowSegmentation (mainwindow.cpp:1936-1969): the viewport composed in one pass; blend values are the sliders' 0..100.
// dst: vh rows of vw BGR pixels.  holes (optional): countNonZero over the whole mask of pure-black pixels.
extern "C" int spx_session_compose(spx_session_t* h, int vx, int vy, int vw, int vh, int show_image, int blend_image, int show_mask,
                                   int blend_mask, int show_grid, int blend_grid, int show_selection, int dilation, int show_holes,
                                   uint8_t* dst, size_t stride, int* holes) try
{
    using namespace spx;
    SES_ENTER(h);
    SPX_REQUIRE(dst && vx >= 0 && vy >= 0 && vw > 0 && vh > 0 && vx + vw <= s->w && vy + vh <= s->h && stride >= (size_t)vw * 3, SPX_ERR_BADARG,
                "compose: viewport %d,%d %dx%d outside the %dx%d image (or bad destination)", vx, vy, vw, vh, s->w, s->h);
    ComposeArgs a;
    a.image = s->planes[0]; a.mask = s->planes[1]; a.grid = s->planes[2]; a.selection = s->planes[3];
    a.w = s->w; a.h = s->h; a.vx = vx; a.vy = vy; a.vw = vw; a.vh = vh;
    a.show_image = show_image; a.show_mask = show_mask; a.show_grid = show_grid; a.show_selection = show_selection; a.show_holes = show_holes;
    a.dilation = dilation;
    // the GUI passes double(slider)/100 to addWeighted, which narrows its scalars to float
    a.a_image = (float)((double)blend_image / 100); a.a_mask = (float)((double)blend_mask / 100); a.a_grid = (float)((double)blend_grid / 100);
    compose_kernel<<<dim3(cdiv(vw, 128), vh), 128, 0, s->st>>>(a, s->out);
    SPX_CUDA(cudaGetLastError());
    if (show_holes && holes) {
        const int n = s->w * s->h;
        SPX_CUDA(cudaMemsetAsync(s->d_flag + 2, 0, sizeof(int), s->st));
        holes_kernel<<<cdiv(n, 256), 256, 0, s->st>>>(s->planes[1], n, s->d_flag + 2);
        SPX_CUDA(cudaMemcpyAsync(s->h_flag + 2, s->d_flag + 2, sizeof(int), cudaMemcpyDeviceToHost, s->st));
    }
    SPX_CUDA(cudaMemcpy2DAsync(dst, stride, s->out, (size_t)vw * 3, (size_t)vw * 3, vh, cudaMemcpyDeviceToHost, s->st));
    SPX_CUDA(cudaStreamSynchronize(s->st));
    if (show_holes && holes) *holes = s->h_flag[2];
    return SPX_OK;
} SPX_API_CATCH
