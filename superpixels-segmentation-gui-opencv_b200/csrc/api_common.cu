// api_common.cu -- version / error reporting / device probing of libspx_b200.so
#include "common.cuh"
#include <cstdlib>
#include <algorithm>
#include <mutex>
#include <thread>
#include <vector>

namespace spx {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
namespace {
std::mutex g_pin_mu;
std::vector<void*> g_pin_free;
constexpr size_t PIN_BLOCK = 256;
}
int pinned_scratch_alloc(void** p, size_t bytes)
{
    SPX_REQUIRE(p && bytes <= PIN_BLOCK, SPX_ERR_INTERNAL, "pinned_scratch_alloc: at most %zu bytes", PIN_BLOCK);
    {
        std::lock_guard<std::mutex> lk(g_pin_mu);
        if (!g_pin_free.empty()) { *p = g_pin_free.back(); g_pin_free.pop_back(); memset(*p, 0, PIN_BLOCK); return SPX_OK; }
    }
    SPX_CUDA(cudaMallocHost(p, PIN_BLOCK));
    memset(*p, 0, PIN_BLOCK);
    return SPX_OK;
}
void pinned_scratch_free(void* p)
{
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_pin_mu);
    if (g_pin_free.size() < 256) g_pin_free.push_back(p);      // (portable pinned memory: usable from any device of the process)
    else cudaFreeHost(p);
}
// ---- staged copies between pageable host memory and the device (see common.cuh) -----------------------------------------
namespace {
constexpr size_t STAGE_BYTES = 2u << 20;             // per pinned buffer
constexpr int STAGE_MAX_LANES = 8;
struct StageLane {
    int device = -1;
    cudaStream_t st = nullptr;
    unsigned char* buf[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    bool ok = false;
    bool init(int dev)
    {
        device = dev;
        ok = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess;
        for (int k = 0; k < 2 && ok; k++)
            ok = cudaMallocHost((void**)&buf[k], STAGE_BYTES) == cudaSuccess && cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming) == cudaSuccess;
        if (!ok) (void)cudaGetLastError();
        return ok;
    }
};
std::mutex g_stage_mu;
std::vector<StageLane*> g_stage_free;
StageLane* stage_acquire(int dev)
{
    {
        std::lock_guard<std::mutex> lk(g_stage_mu);
        for (size_t i = 0; i < g_stage_free.size(); i++)
            if (g_stage_free[i]->device == dev) { StageLane* l = g_stage_free[i]; g_stage_free.erase(g_stage_free.begin() + i); return l; }
    }
    StageLane* l = new StageLane();
    if (!l->init(dev)) { delete l; return nullptr; }      // (leaks nothing that matters: creation failed half way only when the device is gone)
    return l;
}
void stage_release(StageLane* l)
{
    std::lock_guard<std::mutex> lk(g_stage_mu);
    g_stage_free.push_back(l);
}
bool is_pageable(const void* p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { (void)cudaGetLastError(); return true; }
    return at.type == cudaMemoryTypeUnregistered;
}
int stage_lanes(size_t bytes)
{
    static const int hw = (int)std::thread::hardware_concurrency();
    static const int forced = getenv("SPX_STAGE_LANES") ? atoi(getenv("SPX_STAGE_LANES")) : -1;     // 0 = never stage
    if (forced == 0 || bytes < STAGE_BYTES) return 0;
    int n = forced > 0 ? forced : std::max(1, std::min(4, hw / 4));          // measured on a 16-core host: 3 / 4 / 6 / 8 lanes -> 6.97 / 5.71 / 5.93 / 6.24 ms per 4K click
    static const size_t per_lane = getenv("SPX_STAGE_PER_LANE_MB") ? (size_t)atoi(getenv("SPX_STAGE_PER_LANE_MB")) << 20 : STAGE_BYTES;
    n = (int)std::min<size_t>((size_t)n, std::max<size_t>(1, bytes / per_lane));
    return std::min(n, STAGE_MAX_LANES);
}
// one lane: rows [r0, r1).  up: host -> device, otherwise device -> host.  Returns a cudaError_t as int.
int stage_lane_run(StageLane* L, bool up, unsigned char* dev, size_t dpitch, unsigned char* host, size_t hpitch, size_t wb, size_t r0, size_t r1)
{
    if (cudaSetDevice(L->device) != cudaSuccess) return (int)cudaGetLastError();
    if (wb > STAGE_BYTES) return (int)cudaErrorInvalidValue;
    // chunks of at most one staging buffer, about four per lane but not below 512 KB (measured on one box, tools/stage_ab.sh:
    // 4K click 8.34 ms with the runtime's own staging, 6.52 ms with this policy; 2 lanes 8.49 ms; 256 KB chunks 6.94 ms)
    const size_t lane_bytes = (r1 - r0) * wb;
    static const size_t min_chunk = getenv("SPX_STAGE_MIN_CHUNK_KB") ? (size_t)atoi(getenv("SPX_STAGE_MIN_CHUNK_KB")) << 10 : (512u << 10);
    const size_t chunk = std::min(STAGE_BYTES, std::max<size_t>(min_chunk, lane_bytes / 4));
    const size_t rows_per = std::max<size_t>(1, chunk / wb);
    cudaError_t e = cudaSuccess;
    int k = 0;
    if (up) {
        for (size_t r = r0; r < r1 && e == cudaSuccess; r += rows_per, k ^= 1) {
            const size_t n = std::min(rows_per, r1 - r);
            e = cudaEventSynchronize(L->ev[k]);                    // the DMA that last read this buffer
            if (e != cudaSuccess) break;
            for (size_t i = 0; i < n; i++) memcpy(L->buf[k] + i * wb, host + (r + i) * hpitch, wb);
            e = cudaMemcpy2DAsync(dev + r * dpitch, dpitch, L->buf[k], wb, wb, n, cudaMemcpyHostToDevice, L->st);
            if (e == cudaSuccess) e = cudaEventRecord(L->ev[k], L->st);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(L->st);
    } else {
        // software pipeline: chunk c+1 is in flight while chunk c is copied out of its pinned buffer
        size_t r = r0;
        size_t pend_r = 0, pend_n = 0;
        int pend_k = -1;
        while ((r < r1 || pend_k >= 0) && e == cudaSuccess) {
            int cur_k = -1;
            size_t cur_r = 0, cur_n = 0;
            if (r < r1) {
                cur_k = k; cur_r = r; cur_n = std::min(rows_per, r1 - r);
                e = cudaMemcpy2DAsync(L->buf[k], wb, dev + r * dpitch, dpitch, wb, cur_n, cudaMemcpyDeviceToHost, L->st);
                if (e == cudaSuccess) e = cudaEventRecord(L->ev[k], L->st);
                r += cur_n; k ^= 1;
            }
            if (pend_k >= 0 && e == cudaSuccess) {
                e = cudaEventSynchronize(L->ev[pend_k]);
                if (e == cudaSuccess)
                    for (size_t i = 0; i < pend_n; i++) memcpy(host + (pend_r + i) * hpitch, L->buf[pend_k] + i * wb, wb);
            }
            pend_k = cur_k; pend_r = cur_r; pend_n = cur_n;
        }
    }
    return (int)e;
}
int staged_copy(bool up, void* dev, size_t dpitch, void* host, size_t hpitch, size_t wb, size_t rows, int lanes)
{
    int device = 0;
    SPX_CUDA(cudaGetDevice(&device));
    std::vector<StageLane*> L;
    for (int i = 0; i < lanes; i++) {
        StageLane* l = stage_acquire(device);
        if (!l) break;
        L.push_back(l);
    }
    SPX_REQUIRE(!L.empty(), SPX_ERR_GPU, "staged copy: cannot create a staging lane (pinned memory / stream)");
    const int n = (int)L.size();
    std::vector<int> rc((size_t)n, 0);
    std::vector<std::thread> th;
    auto part = [&](int i) { return rows * (size_t)i / (size_t)n; };
    auto run = [&](int i) { rc[i] = stage_lane_run(L[i], up, (unsigned char*)dev, dpitch, (unsigned char*)host, hpitch, wb, part(i), part(i + 1)); };
    th.reserve((size_t)n);
    int started = 1;                                               // lanes 1 .. started-1 have a thread; a failed thread creation
    for (; started < n; started++) {                               // (resource limits) leaves the rest to this thread
        try { th.emplace_back(run, started); } catch (...) { break; }
    }
    run(0);
    for (int i = started; i < n; i++) run(i);
    for (auto& t : th) t.join();
    for (StageLane* l : L) stage_release(l);
    for (int i = 0; i < n; i++)
        SPX_REQUIRE(rc[i] == 0, SPX_ERR_GPU, "staged copy: %s", cudaGetErrorString((cudaError_t)rc[i]));
    return SPX_OK;
}
}  // namespace
int upload_2d(void* d_dst, size_t dpitch, const void* h_src, size_t hpitch, size_t width_bytes, size_t rows, cudaStream_t st)
{
    if (!rows || !width_bytes) return SPX_OK;
    const int lanes = (width_bytes <= STAGE_BYTES && is_pageable(h_src)) ? stage_lanes(width_bytes * rows) : 0;
    if (lanes == 0) {
        // pinned source: asynchronous on st, as cudaMemcpy2DAsync always was here (the buffer must stay valid until st reaches it);
        // small pageable source: the runtime stages it and returns when the host buffer is free
        SPX_CUDA(cudaMemcpy2DAsync(d_dst, dpitch, h_src, hpitch, width_bytes, rows, cudaMemcpyHostToDevice, st));
        return SPX_OK;
    }
    SPX_CUDA(cudaStreamSynchronize(st));
    return staged_copy(true, d_dst, dpitch, const_cast<void*>(h_src), hpitch, width_bytes, rows, lanes);
}
int download_2d(void* h_dst, size_t hpitch, const void* d_src, size_t dpitch, size_t width_bytes, size_t rows, cudaStream_t st)
{
    if (!rows || !width_bytes) { SPX_CUDA(cudaStreamSynchronize(st)); return SPX_OK; }
    const int lanes = (width_bytes <= STAGE_BYTES && is_pageable(h_dst)) ? stage_lanes(width_bytes * rows) : 0;
    if (lanes == 0) {
        SPX_CUDA(cudaMemcpy2DAsync(h_dst, hpitch, d_src, dpitch, width_bytes, rows, cudaMemcpyDeviceToHost, st));
        SPX_CUDA(cudaStreamSynchronize(st));
        return SPX_OK;
    }
    SPX_CUDA(cudaStreamSynchronize(st));
    return staged_copy(false, const_cast<void*>(d_src), dpitch, h_dst, hpitch, width_bytes, rows, lanes);
}
namespace {
int stage_lane_transform(StageLane* L, unsigned char* dev, size_t dpitch, const unsigned char* src, size_t spitch, unsigned char* dst, size_t hdpitch,
                         size_t wb, size_t r0, size_t r1, staged_kernel_fn kernel, void* ctx, int* krc)
{
    if (cudaSetDevice(L->device) != cudaSuccess) return (int)cudaGetLastError();
    const size_t lane_bytes = (r1 - r0) * wb;
    const size_t chunk = std::min(STAGE_BYTES, std::max<size_t>(512u << 10, lane_bytes / 4));
    const size_t rows_per = std::max<size_t>(1, chunk / wb);
    cudaError_t e = cudaSuccess;
    int k = 0, pend_k = -1;
    size_t r = r0, pend_r = 0, pend_n = 0;
    while ((r < r1 || pend_k >= 0) && e == cudaSuccess && *krc == SPX_OK) {
        int cur_k = -1;
        size_t cur_r = 0, cur_n = 0;
        if (r < r1) {
            cur_k = k; cur_r = r; cur_n = std::min(rows_per, r1 - r);
            // buf[k] was drained two steps ago (the pending chunk sits in the other buffer)
            for (size_t i = 0; i < cur_n; i++) memcpy(L->buf[k] + i * wb, src + (r + i) * spitch, wb);
            e = cudaMemcpy2DAsync(dev + r * dpitch, dpitch, L->buf[k], wb, wb, cur_n, cudaMemcpyHostToDevice, L->st);
            if (e == cudaSuccess) { const int rc = kernel(ctx, dev + r * dpitch, cur_n, L->st); if (rc != SPX_OK) *krc = rc; }
            if (e == cudaSuccess) e = cudaMemcpy2DAsync(L->buf[k], wb, dev + r * dpitch, dpitch, wb, cur_n, cudaMemcpyDeviceToHost, L->st);
            if (e == cudaSuccess) e = cudaEventRecord(L->ev[k], L->st);
            r += cur_n; k ^= 1;
        }
        if (pend_k >= 0 && e == cudaSuccess) {
            e = cudaEventSynchronize(L->ev[pend_k]);
            if (e == cudaSuccess)
                for (size_t i = 0; i < pend_n; i++) memcpy(dst + (pend_r + i) * hdpitch, L->buf[pend_k] + i * wb, wb);
        }
        pend_k = cur_k; pend_r = cur_r; pend_n = cur_n;
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(L->st);
    return (int)e;
}
}  // namespace
int staged_transform_2d(void* h_dst, size_t hdpitch, const void* h_src, size_t hspitch, unsigned char* d_scratch, size_t dpitch,
                        size_t width_bytes, size_t rows, staged_kernel_fn kernel, void* ctx)
{
    const int lanes = (width_bytes <= STAGE_BYTES && is_pageable(h_src) && is_pageable(h_dst)) ? stage_lanes(width_bytes * rows) : 0;
    if (lanes == 0) return SPX_NOT_STAGED;
    int device = 0;
    SPX_CUDA(cudaGetDevice(&device));
    std::vector<StageLane*> L;
    for (int i = 0; i < lanes; i++) {
        StageLane* l = stage_acquire(device);
        if (!l) break;
        L.push_back(l);
    }
    if (L.empty()) return SPX_NOT_STAGED;
    const int n = (int)L.size();
    std::vector<int> rc((size_t)n, 0), krc((size_t)n, SPX_OK);
    std::vector<std::thread> th;
    auto part = [&](int i) { return rows * (size_t)i / (size_t)n; };
    auto run = [&](int i) {
        rc[i] = stage_lane_transform(L[i], d_scratch, dpitch, (const unsigned char*)h_src, hspitch, (unsigned char*)h_dst, hdpitch, width_bytes,
                                     part(i), part(i + 1), kernel, ctx, &krc[i]);
    };
    th.reserve((size_t)n);
    int started = 1;
    for (; started < n; started++) {
        try { th.emplace_back(run, started); } catch (...) { break; }
    }
    run(0);
    for (int i = started; i < n; i++) run(i);
    for (auto& t : th) t.join();
    for (StageLane* l : L) stage_release(l);
    for (int i = 0; i < n; i++) {
        SPX_REQUIRE(rc[i] == 0, SPX_ERR_GPU, "staged transform: %s", cudaGetErrorString((cudaError_t)rc[i]));
        if (krc[i] != SPX_OK) return krc[i];
    }
    return SPX_OK;
}
bool pdl_enabled()
{
    static const bool on = !getenv("SPX_NO_PDL");
    return on;
}
int require_device(int device)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        set_error("no CUDA device available (%s); this library has no CPU fallback",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        (void)cudaGetLastError();
        return SPX_ERR_GPU;
    }
    if (device < 0 || device >= n) {
        set_error("device %d out of range [0,%d)", device, n);
        return SPX_ERR_BADARG;
    }
    SPX_CUDA(cudaSetDevice(device));
    return SPX_OK;
}
int pool_alloc(void** p, size_t bytes, cudaStream_t st)
{
    static bool tuned[64] = {false};
    int dev = 0;
    SPX_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && !tuned[dev]) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        tuned[dev] = true;
    }
    cudaError_t e = cudaMallocAsync(p, bytes > 0 ? bytes : 4, st);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        set_error("device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        return e == cudaErrorMemoryAllocation ? SPX_ERR_NOMEM : SPX_ERR_GPU;
    }
    return SPX_OK;
}
void pool_free(void* p, cudaStream_t st)
{
    if (p) cudaFreeAsync(p, st);
}
}  // namespace spx

extern "C" int spx_release_cached_memory(void) try
{
    int dev = 0;
    SPX_CUDA(cudaGetDevice(&dev));
    cudaMemPool_t pool;
    SPX_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
    SPX_CUDA(cudaDeviceSynchronize());
    SPX_CUDA(cudaMemPoolTrimTo(pool, 0));
    return SPX_OK;
} SPX_API_CATCH
extern "C" int spx_version(void) { return 100; }
extern "C" const char* spx_last_error(void) { return spx::g_err; }
extern "C" int spx_device_count(int* count) try
{
    SPX_REQUIRE(count, SPX_ERR_BADARG, "count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { (void)cudaGetLastError(); n = 0; }
    *count = n;
    return SPX_OK;
} SPX_API_CATCH

// Host threads waiting for the device sleep instead of spinning (cudaDeviceScheduleBlockingSync) -- for hosts that run more
// worker threads than they have cores (8 ranks x 4 pipeline threads on a 32-core box); 0 restores the driver's default.
extern "C" int spx_set_blocking_sync(int device, int on) try
{
    using namespace spx;
    SPX_CHECK(require_device(device));
    SPX_CUDA(cudaSetDeviceFlags(on ? cudaDeviceScheduleBlockingSync : cudaDeviceScheduleAuto));
    return SPX_OK;
} SPX_API_CATCH
