// common.cuh -- shared helpers of the B200 superpixel engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <exception>
#include <new>
#include "../../include/spx_b200.h"

namespace spx {

void set_error(const char* fmt, ...);

#define SPX_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            spx::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return SPX_ERR_GPU;                                                                \
        }                                                                                      \
    } while (0)

#define SPX_CHECK(st)                         \
    do {                                      \
        int s__ = (st);                       \
        if (s__ != SPX_OK) return s__;        \
    } while (0)

#define SPX_REQUIRE(cond, code, ...)          \
    do {                                      \
        if (!(cond)) {                        \
            spx::set_error(__VA_ARGS__);      \
            return (code);                    \
        }                                     \
    } while (0)

// Every extern "C" entry point is a function-try-block ending in SPX_API_CATCH: nothing may throw across the C ABI
// (std::bad_alloc from new / std::vector / std::map on the host side becomes SPX_ERR_NOMEM).
#define SPX_API_CATCH                                                                          \
    catch (const std::bad_alloc&) { spx::set_error("out of host memory"); return SPX_ERR_NOMEM; } \
    catch (const std::exception& ex__) { spx::set_error("internal error: %s", ex__.what()); return SPX_ERR_INTERNAL; } \
    catch (...) { spx::set_error("internal error (unknown exception)"); return SPX_ERR_INTERNAL; }

// several kernels launch one grid row per image row (gridDim.y <= 65535): taller images are rejected up front
#define SPX_MAX_HEIGHT 65535
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline size_t round_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

// Programmatic dependent launch: the kernel may become resident while its predecessor in the stream drains; it must execute
// pdl_wait() before touching anything the predecessor writes.  pdl_trigger() in the predecessor lets the dependent in early.
// Both are no-ops in launches without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, Args&&... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// fails loudly when there is no usable CUDA device (no CPU fallback anywhere in this library)
int require_device(int device);

// Device memory comes from the device's stream-ordered pool (cudaMallocAsync) with the release threshold raised, so that
// the reference's usage pattern -- a new Superpixel object on every COMPUTE click -- reuses the previous object's memory
// instead of paying cudaMalloc/cudaFree (milliseconds for the 100+ MB buffers of a 4K frame) each time.
int pool_alloc(void** p, size_t bytes, cudaStream_t st);
void pool_free(void* p, cudaStream_t st);
// Small pinned host scratch (a few flags / counters per object).  cudaMallocHost / cudaFreeHost cost a good fraction of a
// millisecond each -- per COMPUTE click, since the GUI makes a new Superpixel object every time -- so freed blocks of 256 bytes
// are kept and handed out again.
int pinned_scratch_alloc(void** p, size_t bytes);
void pinned_scratch_free(void* p);
// Host <-> device copies of whole images for the host-buffer entry points.  The GUI's cv::Mat buffers are PAGEABLE: a plain
// cudaMemcpy2D of a 4K frame from pageable memory runs at ~11 GB/s (the driver stages it on one thread) and such copies are 70 %
// of a COMPUTE click.  These helpers check the pointer: pinned / registered memory goes straight to cudaMemcpy2DAsync; pageable
// memory of 2 MB or more is cut into row ranges, one per lane (up to 4 host threads), each lane staging through its own pair of
// pinned buffers on its own stream, so the host-side memcpy of one chunk overlaps the DMA of the previous one and the lanes
// overlap each other.  For pageable memory both return only when the host buffer is free / filled (like the pageable cudaMemcpy2D
// they replace) and `st` is synchronised first, so the copy is ordered after everything enqueued on it; a PINNED source is
// uploaded asynchronously on `st` exactly as before (download_2d always returns with the data in place).
int upload_2d(void* d_dst, size_t dpitch, const void* h_src, size_t hpitch, size_t width_bytes, size_t rows, cudaStream_t st);
// host image -> device -> per-row-range kernel -> host image, pipelined on the same lanes (cvtColor on host buffers): a lane
// uploads a chunk of rows, runs `kernel(ctx, d_rows, n_rows, lane_stream)` on it and downloads it while it fills the next chunk,
// so the whole call takes about max(upload, download) instead of their sum.  d_scratch: device image of `rows` rows, pitch
// dpitch.  Returns SPX_ERR_INTERNAL + 1000 when the buffers are not pageable / too small to stage (the caller then copies whole).
typedef int (*staged_kernel_fn)(void* ctx, unsigned char* d_rows, size_t n_rows, cudaStream_t st);
constexpr int SPX_NOT_STAGED = 1000;
int staged_transform_2d(void* h_dst, size_t hdpitch, const void* h_src, size_t hspitch, unsigned char* d_scratch, size_t dpitch,
                        size_t width_bytes, size_t rows, staged_kernel_fn kernel, void* ctx);
int download_2d(void* h_dst, size_t hpitch, const void* d_src, size_t dpitch, size_t width_bytes, size_t rows, cudaStream_t st);

// 8-bit BGR -> Lab (mode 0) / HSV (mode 1) on device buffers, in place allowed (color.cu)
int cvt8_dev(int mode, const void* d_src, size_t sstride, void* d_dst, size_t dstride, int w, int h, cudaStream_t st);

// ---- post-processing stages shared by SLIC / LSC / SEEDS engines (post.cu) ----
struct PostWorkspace {
    int w = 0, h = 0;
    int* parent = nullptr;      // union-find parents / component roots        [N]
    int* csize = nullptr;       // component sizes (indexed by root pixel)      [N]
    int* newid = nullptr;       // new label per component root                [N]
    int* link = nullptr;        // small component -> component supplying `adjlabel` [N]
    int* blocksum = nullptr;    // scan scratch
    int* flag = nullptr;        // device flags / counters [8]
    unsigned char* diff = nullptr;   // contour: differing-neighbour bit masks [N]
    int* h_flag = nullptr;      // pinned host mirror [8]
    long long launches = 0;
    cudaStream_t st_alloc = nullptr;
    int alloc(int w_, int h_, cudaStream_t st = nullptr);
    void release();
};
// labels: dense int32 with pitch `lpitch` (elements). In place. numlabels_out written to ws.h_flag[0]
// asynchronously is NOT done: the function synchronises the stream once to read the new label count.
int enforce_connectivity_dev(PostWorkspace& ws, int* d_labels, int lpitch, int numlabels, int min_element_size,
                             cudaStream_t st, int* numlabels_out);
// flavour 0: SLIC/LSC istaken rule; 1: SEEDS rule.  d_mask pitch in bytes.
int contour_mask_dev(PostWorkspace& ws, const int* d_labels, int lpitch, int thick_line, int flavour,
                     unsigned char* d_mask, size_t mpitch, cudaStream_t st);

}  // namespace spx
