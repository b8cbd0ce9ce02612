// prefilter.cu -- the image pre-processing of the Segmentation tab (MainWindow::on_button_apply_clicked,
// mainwindow.cpp:1979-2047; mat-image-tools.cpp:184-229), i.e. the step that feeds the hot path:
//   equalize        EqualizeHistogram: BGR2YCrCb, equalizeHist on Y, YCrCb2BGR
//   normalize       cv::normalize(image, image, 1, 255, NORM_MINMAX)
//   colour balance  SimplestColorBalance(image, percent): per channel clamp to the percent/2 percentiles, stretch to 0..255
//   gaussian blur   cv::GaussianBlur(image, image, Size(3,3), 0, 0)
// applied in that order on the device in one upload / download.  Every stage is either a 256-entry look-up table derived
// from a histogram (computed on the device, the table on the host: 256 entries) or a 3x3 stencil; the arithmetic is OpenCV
// 4.13's 8-bit integer / float32 pipeline, pinned against cv2 in tests/test_prefilter.py.  Non-local-means denoising and the
// Canny contour overlay of the same dialog are not data-parallel stencils of this kind and stay with OpenCV.
#include "common.cuh"
#include <algorithm>
#include <cmath>
#include <vector>

namespace spx {

// per-channel histograms of an interleaved 8UC3 image (hist[c][v]); MODE 1: channel 0 holds Y of BGR2YCrCb instead of B
template <int MODE>
__global__ void __launch_bounds__(256) hist3_kernel(const unsigned char* __restrict__ img, size_t pitch, int w, int h, unsigned* __restrict__ hist)
{
    __shared__ unsigned sh[3 * 256];
    for (int i = threadIdx.x; i < 768; i += 256) sh[i] = 0;
    __syncthreads();
    const int y = blockIdx.y;
    for (int x = blockIdx.x * 256 + threadIdx.x; x < w; x += gridDim.x * 256) {
        const unsigned char* p = img + (size_t)y * pitch + (size_t)x * 3;
        const int b = p[0], g = p[1], r = p[2];
        if (MODE == 1) atomicAdd(&sh[(b * 1868 + g * 9617 + r * 4899 + 8192) >> 14], 1u);
        else { atomicAdd(&sh[b], 1u); atomicAdd(&sh[256 + g], 1u); atomicAdd(&sh[512 + r], 1u); }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 768; i += 256)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}
// per channel look-up (lut[c][v]), in place
__global__ void __launch_bounds__(256) lut3_kernel(unsigned char* __restrict__ img, size_t pitch, int w, int h, const unsigned char* __restrict__ lut)
{
    __shared__ unsigned char sl[768];
    for (int i = threadIdx.x; i < 768; i += 256) sl[i] = lut[i];
    __syncthreads();
    const int y = blockIdx.y;
    for (int x = blockIdx.x * 256 + threadIdx.x; x < w; x += gridDim.x * 256) {
        unsigned char* p = img + (size_t)y * pitch + (size_t)x * 3;
        p[0] = sl[p[0]]; p[1] = sl[256 + p[1]]; p[2] = sl[512 + p[2]];
    }
}
__device__ __forceinline__ int sat8(int v) { return min(max(v, 0), 255); }
// EqualizeHistogram: BGR -> YCrCb (8-bit, shift 14), Y <- lut[Y], YCrCb -> BGR; in place
__global__ void __launch_bounds__(256) equalize_kernel(unsigned char* __restrict__ img, size_t pitch, int w, int h, const unsigned char* __restrict__ lut)
{
    __shared__ unsigned char sl[256];
    sl[threadIdx.x] = lut[threadIdx.x];
    __syncthreads();
    const int y = blockIdx.y;
    for (int x = blockIdx.x * 256 + threadIdx.x; x < w; x += gridDim.x * 256) {
        unsigned char* p = img + (size_t)y * pitch + (size_t)x * 3;
        const int b = p[0], g = p[1], r = p[2];
        const int Y = (b * 1868 + g * 9617 + r * 4899 + 8192) >> 14;
        const int Cr = sat8(((r - Y) * 11682 + 128 * 16384 + 8192) >> 14) - 128;
        const int Cb = sat8(((b - Y) * 9241 + 128 * 16384 + 8192) >> 14) - 128;
        const int Ye = sl[Y];
        p[0] = (unsigned char)sat8(Ye + ((Cb * 29049 + 8192) >> 14));
        p[1] = (unsigned char)sat8(Ye + ((Cb * -5636 + Cr * -11698 + 8192) >> 14));
        p[2] = (unsigned char)sat8(Ye + ((Cr * 22987 + 8192) >> 14));
    }
}
// GaussianBlur 3x3, sigma 0: taps 1 2 1 / 2 4 2 / 1 2 1, (sum + 8) >> 4 (exact in OpenCV's 8.8 fixed point), BORDER_REFLECT_101
__global__ void __launch_bounds__(256) blur3_kernel(const unsigned char* __restrict__ src, size_t sp, unsigned char* __restrict__ dst, size_t dp, int w, int h)
{
    const int i = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;          // i: byte column 0 .. 3w-1
    if (i >= 3 * w) return;
    const int x = i / 3;
    const int il = x > 0 ? i - 3 : (w > 1 ? i + 3 : i), ir = x < w - 1 ? i + 3 : (w > 1 ? i - 3 : i);
    const int yu = y > 0 ? y - 1 : (h > 1 ? 1 : 0), yd = y < h - 1 ? y + 1 : (h > 1 ? h - 2 : 0);
    const unsigned char *r0 = src + (size_t)yu * sp, *r1 = src + (size_t)y * sp, *r2 = src + (size_t)yd * sp;
    const int s = (r0[il] + 2 * r0[i] + r0[ir]) + 2 * (r1[il] + 2 * r1[i] + r1[ir]) + (r2[il] + 2 * r2[i] + r2[ir]);
    dst[(size_t)y * dp + i] = (unsigned char)((s + 8) >> 4);
}

// ---- host-side tables (256 entries each; float32 arithmetic as OpenCV's) ----------------------------------------
static inline unsigned char sat_round(float v)
{
    const float r = nearbyintf(v);                      // round half to even, as cvRound
    return (unsigned char)(r < 0.f ? 0.f : (r > 255.f ? 255.f : r));
}
// cv::equalizeHist
static void equalize_lut(const unsigned* hist, long long total, unsigned char* lut)
{
    int i0 = 0;
    while (i0 < 256 && !hist[i0]) i0++;
    if (i0 == 256) { for (int i = 0; i < 256; i++) lut[i] = (unsigned char)i; return; }
    if ((long long)hist[i0] == total) { for (int i = 0; i < 256; i++) lut[i] = (unsigned char)i0; return; }
    const float scale = 255.0f / (float)(total - hist[i0]);
    long long sum = 0;
    for (int i = 0; i < 256; i++) lut[i] = 0;
    for (int i = i0 + 1; i < 256; i++) { sum += hist[i]; lut[i] = sat_round((float)sum * scale); }
}
// cv::normalize(NORM_MINMAX) to [lo, hi] over the value range [mn, mx]: convertTo(scale, shift) = float32 fma(v, scale, shift)
// in OpenCV 4.13's dispatched build (pinned over every (min, max) pair in tests/test_prefilter.py)
static void minmax_lut(int mn, int mx, double lo, double hi, unsigned char* lut)
{
    const double scale = (mx - mn) > 2.220446049250313e-16 ? (hi - lo) / (double)(mx - mn) : 0.0;
    const double shift = lo - (double)mn * scale;
    const float fs = (float)scale, fh = (float)shift;
    for (int v = 0; v < 256; v++) lut[v] = sat_round(fmaf((float)v, fs, fh));
}

struct PrefilterBuffers {
    unsigned char* img = nullptr; unsigned char* tmp = nullptr; unsigned* hist = nullptr; unsigned char* lut = nullptr;
    size_t pitch = 0;
    cudaStream_t st = nullptr;
    ~PrefilterBuffers()
    {
        pool_free(img, st); pool_free(tmp, st); pool_free(hist, st); pool_free(lut, st);
        if (st) cudaStreamDestroy(st);
    }
};

}  // namespace spx

// flags: bit 0 equalize, bit 1 normalize, bit 2 colour balance (with `percent`), bit 3 gaussian blur; applied in the order of
// on_button_apply_clicked (mainwindow.cpp:2028-2032).  src / dst: 8UC3 host images (may be the same buffer).
extern "C" int spx_preprocess_8uc3(const uint8_t* src, size_t src_stride, uint8_t* dst, size_t dst_stride, int width, int height, int flags,
                                   int percent, int device) try
{
    using namespace spx;
    SPX_REQUIRE(src && dst && width > 0 && height > 0 && height <= SPX_MAX_HEIGHT && src_stride >= (size_t)width * 3 && dst_stride >= (size_t)width * 3, SPX_ERR_BADARG,
                "preprocess: bad image / stride");
    SPX_REQUIRE(percent >= 0 && percent <= 100, SPX_ERR_BADARG, "preprocess: percent must lie in 0..100");
    SPX_CHECK(require_device(device));
    SPX_CUDA(cudaSetDevice(device));
    const int w = width, h = height;
    const long long npx = (long long)w * h;
    PrefilterBuffers B;
    SPX_CUDA(cudaStreamCreateWithFlags(&B.st, cudaStreamNonBlocking));
    B.pitch = round_up((size_t)w * 3, 16);
    SPX_CHECK(pool_alloc((void**)&B.img, B.pitch * h, B.st));
    SPX_CHECK(pool_alloc((void**)&B.tmp, B.pitch * h, B.st));
    SPX_CHECK(pool_alloc((void**)&B.hist, 768 * sizeof(unsigned), B.st));
    SPX_CHECK(pool_alloc((void**)&B.lut, 768, B.st));
    SPX_CUDA(cudaMemcpy2DAsync(B.img, B.pitch, src, src_stride, (size_t)w * 3, h, cudaMemcpyHostToDevice, B.st));
    const dim3 grid(std::min(cdiv(w, 256), 8), h);
    unsigned hh[768];
    unsigned char lut[768];
    auto histogram = [&](int mode) -> int {
        SPX_CUDA(cudaMemsetAsync(B.hist, 0, 768 * sizeof(unsigned), B.st));
        if (mode) hist3_kernel<1><<<grid, 256, 0, B.st>>>(B.img, B.pitch, w, h, B.hist);
        else hist3_kernel<0><<<grid, 256, 0, B.st>>>(B.img, B.pitch, w, h, B.hist);
        SPX_CUDA(cudaGetLastError());
        SPX_CUDA(cudaMemcpyAsync(hh, B.hist, sizeof(hh), cudaMemcpyDeviceToHost, B.st));
        SPX_CUDA(cudaStreamSynchronize(B.st));
        return SPX_OK;
    };
    if (flags & 1) {                                        // EqualizeHistogram (mat-image-tools.cpp:184-202)
        SPX_CHECK(histogram(1));
        equalize_lut(hh, npx, lut);
        SPX_CUDA(cudaMemcpyAsync(B.lut, lut, 256, cudaMemcpyHostToDevice, B.st));
        equalize_kernel<<<grid, 256, 0, B.st>>>(B.img, B.pitch, w, h, B.lut);
        SPX_CUDA(cudaGetLastError());
    }
    if (flags & 2) {                                        // cv::normalize(image, image, 1, 255, NORM_MINMAX): one range for all channels
        SPX_CHECK(histogram(0));
        int mn = 255, mx = 0;
        for (int i = 0; i < 768; i++) if (hh[i]) { mn = std::min(mn, i & 255); mx = std::max(mx, i & 255); }
        minmax_lut(mn, mx, 1.0, 255.0, lut);
        memcpy(lut + 256, lut, 256); memcpy(lut + 512, lut, 256);
        SPX_CUDA(cudaMemcpyAsync(B.lut, lut, 768, cudaMemcpyHostToDevice, B.st));
        lut3_kernel<<<grid, 256, 0, B.st>>>(B.img, B.pitch, w, h, B.lut);
        SPX_CUDA(cudaGetLastError());
    }
    if (flags & 4) {                                        // SimplestColorBalance (mat-image-tools.cpp:204-229)
        SPX_CHECK(histogram(0));
        const float percent_f = (float)percent, half_percent = percent_f / 200.0f;
        const float fn = (float)npx;
        long long ilo = (long long)std::floor((double)(fn * half_percent));                     // cvFloor(float(cols) * half_percent)
        long long ihi = (long long)std::ceil((double)fn * (1.0 - (double)half_percent));        // cvCeil(float(cols) * (1.0 - half_percent))
        if (ilo > npx - 1) ilo = npx - 1;
        if (ihi > npx - 1) ihi = npx - 1;                   // upstream reads one past the end when percent == 0; clamped here
        for (int c = 0; c < 3; c++) {
            const unsigned* hc = hh + 256 * c;
            int lowval = 0, highval = 255;
            long long cum = 0;
            bool gotlo = false;
            for (int v = 0; v < 256; v++) {                 // sorted[i] = the value whose cumulative count first exceeds i
                cum += hc[v];
                if (!gotlo && cum > ilo) { lowval = v; gotlo = true; }
                if (cum > ihi) { highval = v; break; }
            }
            // clamp, then normalize(0, 255, MINMAX) over the clamped channel's actual range
            int mn = 255, mx = 0;
            for (int v = 0; v < 256; v++) if (hc[v]) { const int cv = std::min(std::max(v, lowval), highval); mn = std::min(mn, cv); mx = std::max(mx, cv); }
            unsigned char l2[256];
            minmax_lut(mn, mx, 0.0, 255.0, l2);
            for (int v = 0; v < 256; v++) lut[256 * c + v] = l2[std::min(std::max(v, lowval), highval)];
        }
        SPX_CUDA(cudaMemcpyAsync(B.lut, lut, 768, cudaMemcpyHostToDevice, B.st));
        lut3_kernel<<<grid, 256, 0, B.st>>>(B.img, B.pitch, w, h, B.lut);
        SPX_CUDA(cudaGetLastError());
    }
    unsigned char* result = B.img;
    if (flags & 8) {                                        // cv::GaussianBlur(image, image, Size(3,3), 0, 0)
        blur3_kernel<<<dim3(cdiv(3 * w, 256), h), 256, 0, B.st>>>(B.img, B.pitch, B.tmp, B.pitch, w, h);
        SPX_CUDA(cudaGetLastError());
        result = B.tmp;
    }
    SPX_CUDA(cudaMemcpy2DAsync(dst, dst_stride, result, B.pitch, (size_t)w * 3, h, cudaMemcpyDeviceToHost, B.st));
    SPX_CUDA(cudaStreamSynchronize(B.st));
    return SPX_OK;
} SPX_API_CATCH
