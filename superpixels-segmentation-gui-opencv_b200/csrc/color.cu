// color.cu -- 8-bit BGR->Lab / BGR->HSV (replaces cv::cvtColor at mainwindow.cpp:2096-2097).
//
// Pure integer LUT pipeline (bit-exact with OpenCV's RGB2Lab_b / RGB2HSV_b).  HBM-bound: 3 B read +
// 3 B written per pixel.  Each thread converts 4 pixels = three aligned 32-bit words in, three out;
// the gamma (512 B) and cube-root (4 KB) tables are staged in shared memory once per CTA.
#include "common.cuh"
#include "color_tables.inc"

namespace spx {

// the tables live in ordinary global memory: every lane of a warp stages a DIFFERENT entry into shared memory, which
// the constant cache would serialise 32-fold
__device__ unsigned short c_gamma[256];
__device__ unsigned short c_cbrt[2048];
__device__ int c_sdiv[256];
__device__ int c_hdiv[256];
static bool g_tables_loaded[64] = {false};

static int load_tables()
{
    int dev = 0;
    SPX_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && g_tables_loaded[dev]) return SPX_OK;
    SPX_CUDA(cudaMemcpyToSymbol(c_gamma, SPX_TAB_GAMMA, sizeof(SPX_TAB_GAMMA)));
    SPX_CUDA(cudaMemcpyToSymbol(c_cbrt, SPX_TAB_CBRT, sizeof(SPX_TAB_CBRT)));
    SPX_CUDA(cudaMemcpyToSymbol(c_sdiv, SPX_TAB_SDIV, sizeof(SPX_TAB_SDIV)));
    SPX_CUDA(cudaMemcpyToSymbol(c_hdiv, SPX_TAB_HDIV, sizeof(SPX_TAB_HDIV)));
    if (dev < 64) g_tables_loaded[dev] = true;
    return SPX_OK;
}

__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }
__device__ __forceinline__ int sat8(int v) { return min(max(v, 0), 255); }

struct LabOp {
    const unsigned short* gam;
    const unsigned short* cb;
    __device__ __forceinline__ void operator()(int b, int g, int r, int& o0, int& o1, int& o2) const
    {
        int B = gam[b], G = gam[g], R = gam[r];
        int fX = cb[descale(R * 1777 + G * 1541 + B * 778, 12)];
        int fY = cb[descale(R * 871 + G * 2929 + B * 296, 12)];
        int fZ = cb[descale(R * 73 + G * 448 + B * 3575, 12)];
        // no saturation needed: over all 2^24 colours L stays in 0..255, a in 42..226, b in 20..223 (the
        // exhaustive GPU test compares every colour with the oracle)
        o0 = (296 * fY + (16384 - 1336934)) >> 15;
        o1 = (500 * (fX - fY) + (128 * 32768 + 16384)) >> 15;
        o2 = (200 * (fY - fZ) + (128 * 32768 + 16384)) >> 15;
    }
};
struct HsvOp {
    const int* sdiv;
    const int* hdiv;
    __device__ __forceinline__ void operator()(int b, int g, int r, int& o0, int& o1, int& o2) const
    {
        int v = max(max(b, g), r), mn = min(min(b, g), r);
        int df = v - mn;
        int s = (df * sdiv[v] + (1 << 11)) >> 12;
        int hh = (v == r) ? (g - b) : (v == g) ? (b - r + 2 * df) : (r - g + 4 * df);
        hh = (hh * hdiv[df] + (1 << 11)) >> 12;
        if (hh < 0) hh += 180;
        o0 = hh; o1 = s; o2 = v;
    }
};

// MODE 0 = Lab, 1 = HSV.  VEC: rows and pointers are 4-byte aligned -> 4 pixels per thread as 3 words.
template <int MODE, bool VEC>
__global__ void __launch_bounds__(256) cvt8_kernel(const unsigned char* __restrict__ src, size_t sstride,
                                                   unsigned char* __restrict__ dst, size_t dstride, int w, int h)
{
    __shared__ unsigned short s_gam[256];
    __shared__ unsigned short s_cb[2048];
    __shared__ int s_sd[256];
    __shared__ int s_hd[256];
    if (MODE == 0) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) s_gam[i] = c_gamma[i];
        for (int i = threadIdx.x; i < 2048; i += blockDim.x) s_cb[i] = c_cbrt[i];
    } else {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) { s_sd[i] = c_sdiv[i]; s_hd[i] = c_hdiv[i]; }
    }
    __syncthreads();
    LabOp lab{s_gam, s_cb};
    HsvOp hsv{s_sd, s_hd};
    const int groups = (w + 3) >> 2;   // 4-pixel groups per row
    const long long total = (long long)groups * h;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        int y = (int)(t / groups), gx = (int)(t - (long long)y * groups);
        int x0 = gx * 4;
        const unsigned char* sp = src + (size_t)y * sstride + (size_t)x0 * 3;
        unsigned char* dp = dst + (size_t)y * dstride + (size_t)x0 * 3;
        if (VEC && x0 + 4 <= w) {
            const unsigned int* s4 = reinterpret_cast<const unsigned int*>(sp);
            unsigned int a = s4[0], b = s4[1], c = s4[2];
            int p[12];
            p[0] = a & 255; p[1] = (a >> 8) & 255; p[2] = (a >> 16) & 255; p[3] = a >> 24;
            p[4] = b & 255; p[5] = (b >> 8) & 255; p[6] = (b >> 16) & 255; p[7] = b >> 24;
            p[8] = c & 255; p[9] = (c >> 8) & 255; p[10] = (c >> 16) & 255; p[11] = c >> 24;
            int o[12];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (MODE == 0) lab(p[3 * k], p[3 * k + 1], p[3 * k + 2], o[3 * k], o[3 * k + 1], o[3 * k + 2]);
                else hsv(p[3 * k], p[3 * k + 1], p[3 * k + 2], o[3 * k], o[3 * k + 1], o[3 * k + 2]);
            }
            unsigned int* d4 = reinterpret_cast<unsigned int*>(dp);
            d4[0] = o[0] | (o[1] << 8) | (o[2] << 16) | (o[3] << 24);
            d4[1] = o[4] | (o[5] << 8) | (o[6] << 16) | (o[7] << 24);
            d4[2] = o[8] | (o[9] << 8) | (o[10] << 16) | (o[11] << 24);
        } else {
            int n = min(4, w - x0);
            for (int k = 0; k < n; k++) {
                int o0, o1, o2;
                if (MODE == 0) lab(sp[3 * k], sp[3 * k + 1], sp[3 * k + 2], o0, o1, o2);
                else hsv(sp[3 * k], sp[3 * k + 1], sp[3 * k + 2], o0, o1, o2);
                dp[3 * k] = (unsigned char)o0; dp[3 * k + 1] = (unsigned char)o1; dp[3 * k + 2] = (unsigned char)o2;
            }
        }
    }
}

int cvt8_dev(int mode, const void* d_src, size_t sstride, void* d_dst, size_t dstride, int w, int h, cudaStream_t st)
{
    SPX_REQUIRE(d_src && d_dst && w > 0 && h > 0, SPX_ERR_BADARG, "cvt8: bad argument");
    SPX_REQUIRE(sstride >= (size_t)w * 3 && dstride >= (size_t)w * 3, SPX_ERR_BADARG, "cvt8: stride smaller than a row");
    SPX_CHECK(load_tables());
    bool vec = ((uintptr_t)d_src % 4 == 0) && ((uintptr_t)d_dst % 4 == 0) && sstride % 4 == 0 && dstride % 4 == 0;
    long long total = (long long)((w + 3) / 4) * h;
    int blocks = (int)((total + 255) / 256);
    int maxb = 148 * 8;      // grid-stride: one wave of 8 CTAs per SM, the tables are staged once per CTA
    if (blocks > maxb) blocks = maxb;
    const unsigned char* s = (const unsigned char*)d_src;
    unsigned char* d = (unsigned char*)d_dst;
    if (mode == 0) {
        if (vec) cvt8_kernel<0, true><<<blocks, 256, 0, st>>>(s, sstride, d, dstride, w, h);
        else cvt8_kernel<0, false><<<blocks, 256, 0, st>>>(s, sstride, d, dstride, w, h);
    } else {
        if (vec) cvt8_kernel<1, true><<<blocks, 256, 0, st>>>(s, sstride, d, dstride, w, h);
        else cvt8_kernel<1, false><<<blocks, 256, 0, st>>>(s, sstride, d, dstride, w, h);
    }
    SPX_CUDA(cudaGetLastError());
    return SPX_OK;
}

static int cvt8_host(int mode, const uint8_t* src, size_t sstride, uint8_t* dst, size_t dstride, int w, int h, int device)
{
    SPX_REQUIRE(src && dst && w > 0 && h > 0, SPX_ERR_BADARG, "cvtColor: bad argument");
    SPX_REQUIRE(sstride >= (size_t)w * 3 && dstride >= (size_t)w * 3, SPX_ERR_BADARG, "cvtColor: stride smaller than a row");
    SPX_CHECK(require_device(device));
    size_t pitch = round_up((size_t)w * 3, 16);
    unsigned char* d = nullptr;
    SPX_CHECK(pool_alloc((void**)&d, pitch * h, 0));
    int rc = load_tables();
    if (rc == SPX_OK) rc = (cudaStreamSynchronize(0) == cudaSuccess) ? SPX_OK : SPX_ERR_GPU;      // (the allocation above is stream-0 ordered)
    // pageable images: upload, conversion and download pipelined chunk by chunk on the staging lanes
    struct Ctx { int mode; size_t pitch; int w; } ctx{mode, pitch, w};
    if (rc == SPX_OK) {
        rc = staged_transform_2d(dst, dstride, src, sstride, d, pitch, (size_t)w * 3, h,
                                 [](void* c, unsigned char* rows, size_t n, cudaStream_t st) -> int {
                                     const Ctx* x = static_cast<const Ctx*>(c);
                                     return cvt8_dev(x->mode, rows, x->pitch, rows, x->pitch, x->w, (int)n, st);
                                 }, &ctx);
        if (rc == SPX_NOT_STAGED) {
            rc = upload_2d(d, pitch, src, sstride, (size_t)w * 3, h, 0);
            if (rc == SPX_OK) rc = cvt8_dev(mode, d, pitch, d, pitch, w, h, 0);
            if (rc == SPX_OK) rc = download_2d(dst, dstride, d, pitch, (size_t)w * 3, h, 0);
        }
    }
    pool_free(d, 0);
    return rc;
}
}  // namespace spx

extern "C" int spx_bgr2lab8(const uint8_t* s, size_t ss, uint8_t* d, size_t ds, int w, int h, int dev) { return spx::cvt8_host(0, s, ss, d, ds, w, h, dev); }
extern "C" int spx_bgr2hsv8(const uint8_t* s, size_t ss, uint8_t* d, size_t ds, int w, int h, int dev) { return spx::cvt8_host(1, s, ss, d, ds, w, h, dev); }
extern "C" int spx_bgr2lab8_dev(const void* s, size_t ss, void* d, size_t ds, int w, int h, void* st) { return spx::cvt8_dev(0, s, ss, d, ds, w, h, (cudaStream_t)st); }
extern "C" int spx_bgr2hsv8_dev(const void* s, size_t ss, void* d, size_t ds, int w, int h, void* st) { return spx::cvt8_dev(1, s, ss, d, ds, w, h, (cudaStream_t)st); }
