// api_common.cu -- version / error reporting / device probing of libspx_b200.so
#include "common.cuh"

namespace spx {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
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
}  // namespace spx

extern "C" int spx_version(void) { return 100; }
extern "C" const char* spx_last_error(void) { return spx::g_err; }
extern "C" int spx_device_count(int* count)
{
    SPX_REQUIRE(count, SPX_ERR_BADARG, "count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { (void)cudaGetLastError(); n = 0; }
    *count = n;
    return SPX_OK;
}
