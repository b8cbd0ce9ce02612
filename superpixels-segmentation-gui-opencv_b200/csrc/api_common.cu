// api_common.cu -- version / error reporting / device probing of libspx_b200.so
#include "common.cuh"
#include <cstdlib>
#include <mutex>
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
