// post.cu -- enforceLabelConnectivity and getLabelContourMask on the GPU.
//
// Replaces SuperpixelSLIC/LSC::enforceLabelConnectivity (mainwindow.cpp:2108,2117) and
// Superpixel{SLIC,LSC,SEEDS}::getLabelContourMask (mainwindow.cpp:2111,2120,2127).
// Both upstream routines are raster-order sequential; the data-parallel forms below give
// bit-identical results:
//
//  connectivity:  4-connected components by union-find with root = smallest raster index
//                 (= the pixel at which upstream's scan starts the segment), component sizes,
//                 for every small component a pointer to the component that supplies upstream's
//                 `adjlabel` (last of the L,U,R,D neighbours of the root pixel whose component
//                 starts earlier), pointer chasing to a kept component, new ids = exclusive scan
//                 of kept roots in raster order.
//  contour mask:  upstream's `istaken` recurrence only looks at causal neighbours (W,NW,N,NE), so
//                 it has a unique fixed point; Jacobi sweeps from the all-false state reach it.
#include "common.cuh"

namespace spx {

// ------------------------------------------------------------------------------------------------
// workspace
// ------------------------------------------------------------------------------------------------
int PostWorkspace::alloc(int w_, int h_)
{
    if (w_ == w && h_ == h && parent) return SPX_OK;
    release();
    w = w_; h = h_;
    size_t N = (size_t)w * h;
    SPX_CUDA(cudaMalloc(&parent, N * sizeof(int)));
    SPX_CUDA(cudaMalloc(&csize, N * sizeof(int)));
    SPX_CUDA(cudaMalloc(&newid, N * sizeof(int)));
    SPX_CUDA(cudaMalloc(&link, N * sizeof(int)));
    SPX_CUDA(cudaMalloc(&blocksum, (N / 1024 + 2) * sizeof(int)));
    SPX_CUDA(cudaMalloc(&flag, 8 * sizeof(int)));
    SPX_CUDA(cudaMalloc(&diff, N));
    SPX_CUDA(cudaMalloc(&takenA, N));
    SPX_CUDA(cudaMalloc(&takenB, N));
    SPX_CUDA(cudaMallocHost(&h_flag, 8 * sizeof(int)));
    return SPX_OK;
}
void PostWorkspace::release()
{
    cudaFree(parent); cudaFree(csize); cudaFree(newid); cudaFree(link); cudaFree(blocksum); cudaFree(flag);
    cudaFree(diff); cudaFree(takenA); cudaFree(takenB);
    if (h_flag) cudaFreeHost(h_flag);
    parent = csize = newid = link = blocksum = flag = nullptr;
    diff = takenA = takenB = nullptr; h_flag = nullptr;
    w = h = 0;
}

// ------------------------------------------------------------------------------------------------
// connected components (union-find, root = min raster index)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int uf_find(const int* __restrict__ P, int a)
{
    int p = P[a];
    while (p != a) { a = p; p = P[a]; }
    return a;
}
__device__ __forceinline__ void uf_unite(int* P, int a, int b)
{
    bool done;
    do {
        a = uf_find(P, a);
        b = uf_find(P, b);
        if (a < b) { int old = atomicMin(&P[b], a); done = (old == b); b = old; }
        else if (b < a) { int old = atomicMin(&P[a], b); done = (old == a); a = old; }
        else done = true;
    } while (!done);
}

// parent[i] = i-1 when the left neighbour has the same label (rows become chains), else i
__global__ void ccl_init_kernel(const int* __restrict__ lab, int lp, int w, int h, int* __restrict__ P)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const int* row = lab + (size_t)y * lp;
    int i = y * w + x;
    P[i] = (x > 0 && row[x - 1] == row[x]) ? i - 1 : i;
}
// vertical unions; skipped when the left neighbour already carries the same union
__global__ void ccl_merge_kernel(const int* __restrict__ lab, int lp, int w, int h, int* __restrict__ P)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y + 1;
    if (x >= w || y >= h) return;
    const int* row = lab + (size_t)y * lp;
    const int* up = row - lp;
    int me = row[x];
    if (up[x] != me) return;
    if (x > 0 && row[x - 1] == me && up[x - 1] == me) return;
    uf_unite(P, y * w + x, (y - 1) * w + x);
}
__global__ void ccl_flatten_kernel(int* __restrict__ P, int* __restrict__ csize, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    P[i] = uf_find(P, i);
    csize[i] = 0;
}
// component sizes: warp-aggregated atomics (a warp covers 32 consecutive pixels, i.e. 1-3 roots)
__global__ void ccl_size_kernel(const int* __restrict__ P, int* __restrict__ csize, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool act = i < n;
    int r = act ? P[i] : -1 - (int)(threadIdx.x & 31);
    unsigned m = __match_any_sync(0xffffffffu, r);
    if (act && (int)(threadIdx.x & 31) == __ffs(m) - 1) atomicAdd(&csize[r], __popc(m));
}
// per root pixel: kept flag or link to the component supplying upstream's `adjlabel`
// link encoding in newid[] for small roots: adjacent root index, or -1 (no neighbour labelled: pixel 0)
__global__ void ccl_link_kernel(const int* __restrict__ P, const int* __restrict__ csize, int* __restrict__ link,
                                int w, int h, int min_sp)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    int i = y * w + x;
    if (P[i] != i) return;
    if (csize[i] > min_sp) return;   // kept: newid written by the scan
    int adj = -1;
    if (x > 0) adj = P[i - 1];
    if (y > 0) adj = P[i - w];
    if (x < w - 1) { int r = P[i + 1]; if (r < i) adj = r; }
    if (y < h - 1) { int r = P[i + w]; if (r < i) adj = r; }
    link[i] = adj;
}

// exclusive scan of kept-root flags over raster order: block counts -> block offsets -> ids
#define SCAN_CHUNK 2048
__device__ __forceinline__ int kept_flag(const int* P, const int* csize, int i, int n, int min_sp)
{
    return (i < n && P[i] == i && csize[i] > min_sp) ? 1 : 0;
}
__global__ void __launch_bounds__(256) scan_count_kernel(const int* __restrict__ P, const int* __restrict__ csize, int n,
                                                          int min_sp, int* __restrict__ blocksum)
{
    __shared__ int s[8];
    int base = blockIdx.x * SCAN_CHUNK;
    int c = 0;
    for (int k = threadIdx.x; k < SCAN_CHUNK; k += 256) c += kept_flag(P, csize, base + k, n, min_sp);
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int k = 0; k < 8; k++) t += s[k]; blocksum[blockIdx.x] = t; }
}
__global__ void __launch_bounds__(1024) scan_blocks_kernel(int* __restrict__ blocksum, int nb, int* __restrict__ total)
{
    __shared__ int s_w[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        int i = base + threadIdx.x;
        int v = i < nb ? blocksum[i] : 0;
        int inc = v;
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if ((int)(threadIdx.x & 31) >= o) inc += t; }
        if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) {
            int t = s_w[threadIdx.x], ti = t;
            for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, ti, o); if ((int)threadIdx.x >= o) ti += u; }
            s_w[threadIdx.x] = ti - t;   // exclusive warp offsets
        }
        __syncthreads();
        int excl = s_carry + s_w[threadIdx.x >> 5] + inc - v;
        if (i < nb) blocksum[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = s_carry;
}
__global__ void __launch_bounds__(256) scan_apply_kernel(const int* __restrict__ P, const int* __restrict__ csize, int n,
                                                          int min_sp, const int* __restrict__ blocksum, int* __restrict__ newid)
{
    __shared__ int s_w[8];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = blocksum[blockIdx.x];
    __syncthreads();
    int base = blockIdx.x * SCAN_CHUNK;
    for (int k = 0; k < SCAN_CHUNK; k += 256) {
        int i = base + k + threadIdx.x;
        int f = kept_flag(P, csize, i, n, min_sp);
        unsigned b = __ballot_sync(0xffffffffu, f);
        int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        int pre = __popc(b & ((1u << lane) - 1));
        if (lane == 0) s_w[wid] = __popc(b);
        __syncthreads();
        int off = s_carry;
        for (int q = 0; q < wid; q++) off += s_w[q];
        if (f) newid[i] = off + pre;
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int q = 0; q < 8; q++) t += s_w[q]; s_carry += t; }
        __syncthreads();
    }
}
// small roots: chase links to a kept root (links strictly decrease, so the walk terminates)
__global__ void ccl_resolve_kernel(const int* __restrict__ P, const int* __restrict__ csize, const int* __restrict__ link,
                                   int* __restrict__ newid, int n, int min_sp)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || P[i] != i || csize[i] > min_sp) return;
    int r = link[i];
    while (r >= 0 && csize[r] <= min_sp) r = link[r];
    // r < 0: the chain ends at a small component containing pixel 0 -> upstream's adjlabel is still 0
    newid[i] = r >= 0 ? newid[r] : 0;
}
__global__ void ccl_relabel_kernel(const int* __restrict__ P, const int* __restrict__ newid, int* __restrict__ lab, int lp,
                                   int w, int h)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    lab[(size_t)y * lp + x] = newid[P[y * w + x]];
}

int enforce_connectivity_dev(PostWorkspace& ws, int* d_labels, int lp, int numlabels, int min_element_size,
                             cudaStream_t st, int* numlabels_out)
{
    SPX_REQUIRE(min_element_size >= 0 && min_element_size <= 100, SPX_ERR_ASSERT,
                "enforceLabelConnectivity: min_element_size >= 0 && min_element_size <= 100");
    if (min_element_size == 0) { if (numlabels_out) *numlabels_out = numlabels; return SPX_OK; }
    const int w = ws.w, h = ws.h, n = w * h;
    const int supsz = n / (numlabels > 0 ? numlabels : 1);
    int div = (int)(100.0f / (float)min_element_size + 0.5f);
    if (div < 1) div = 1;
    int min_sp = supsz / div;
    if (min_sp < 3) min_sp = 3;
    int* d_link = ws.link;
    dim3 b2(256), g2(cdiv(w, 256), h);
    int g1 = cdiv(n, 256);
    ccl_init_kernel<<<g2, b2, 0, st>>>(d_labels, lp, w, h, ws.parent);
    if (h > 1) ccl_merge_kernel<<<dim3(cdiv(w, 256), h - 1), b2, 0, st>>>(d_labels, lp, w, h, ws.parent);
    ccl_flatten_kernel<<<g1, 256, 0, st>>>(ws.parent, ws.csize, n);
    ccl_size_kernel<<<g1, 256, 0, st>>>(ws.parent, ws.csize, n);
    ccl_link_kernel<<<g2, b2, 0, st>>>(ws.parent, ws.csize, d_link, w, h, min_sp);
    int nb = cdiv(n, SCAN_CHUNK);
    scan_count_kernel<<<nb, 256, 0, st>>>(ws.parent, ws.csize, n, min_sp, ws.blocksum);
    scan_blocks_kernel<<<1, 1024, 0, st>>>(ws.blocksum, nb, ws.flag);
    scan_apply_kernel<<<nb, 256, 0, st>>>(ws.parent, ws.csize, n, min_sp, ws.blocksum, ws.newid);
    ccl_resolve_kernel<<<g1, 256, 0, st>>>(ws.parent, ws.csize, d_link, ws.newid, n, min_sp);
    ccl_relabel_kernel<<<g2, b2, 0, st>>>(ws.parent, ws.newid, d_labels, lp, w, h);
    ws.launches += (h > 1) ? 10 : 9;
    SPX_CUDA(cudaGetLastError());
    SPX_CUDA(cudaMemcpyAsync(ws.h_flag, ws.flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    SPX_CUDA(cudaStreamSynchronize(st));
    if (numlabels_out) *numlabels_out = ws.h_flag[0];
    return SPX_OK;
}

// ------------------------------------------------------------------------------------------------
// contour mask
// ------------------------------------------------------------------------------------------------
// bit k of diff[p] = neighbour k is in bounds and has a different label; k follows upstream's
// dx8/dy8 order: W, NW, N, NE, E, SE, S, SW.  Bits 0..3 are the causal (already visited) ones.
__global__ void contour_diff_kernel(const int* __restrict__ lab, int lp, int w, int h, unsigned char* __restrict__ diff,
                                    unsigned char* __restrict__ taken)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const int dx8[8] = {-1, -1, 0, 1, 1, 1, 0, -1};
    const int dy8[8] = {0, -1, -1, -1, 0, 1, 1, 1};
    int me = lab[(size_t)y * lp + x];
    unsigned d = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        int xx = x + dx8[k], yy = y + dy8[k];
        if (xx >= 0 && xx < w && yy >= 0 && yy < h && lab[(size_t)yy * lp + xx] != me) d |= 1u << k;
    }
    diff[(size_t)y * w + x] = (unsigned char)d;
    taken[(size_t)y * w + x] = 0;
}
__global__ void contour_sweep_kernel(const unsigned char* __restrict__ diff, const unsigned char* __restrict__ tin,
                                     unsigned char* __restrict__ tout, int w, int h, int lw, int* __restrict__ changed)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    size_t i = (size_t)y * w + x;
    unsigned d = diff[i];
    int np = __popc(d);
    unsigned char t = 0;
    if (np > lw) {
        if ((d & 1u) && tin[i - 1]) np--;
        if ((d & 2u) && tin[i - w - 1]) np--;
        if ((d & 4u) && tin[i - w]) np--;
        if ((d & 8u) && tin[i - w + 1]) np--;
        t = np > lw ? 1 : 0;
    }
    tout[i] = t;
    if (t != tin[i]) *changed = 1;
}
// MODE 0: from the converged `taken` flags; MODE 1: pure stencil popc(diff) > thr
template <int MODE>
__global__ void contour_emit_kernel(const unsigned char* __restrict__ diff, const unsigned char* __restrict__ taken,
                                    unsigned char* __restrict__ mask, size_t mp, int w, int h, int thr)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    size_t i = (size_t)y * w + x;
    bool on = MODE == 0 ? taken[i] != 0 : __popc((unsigned)diff[i]) > thr;
    mask[(size_t)y * mp + x] = on ? 255 : 0;
}

int contour_mask_dev(PostWorkspace& ws, const int* d_labels, int lp, int thick_line, int flavour,
                     unsigned char* d_mask, size_t mp, cudaStream_t st)
{
    const int w = ws.w, h = ws.h;
    dim3 b2(256), g2(cdiv(w, 256), h);
    contour_diff_kernel<<<g2, b2, 0, st>>>(d_labels, lp, w, h, ws.diff, ws.takenA);
    ws.launches++;
    if (flavour == 1 && thick_line) {
        contour_emit_kernel<1><<<g2, b2, 0, st>>>(ws.diff, ws.takenA, d_mask, mp, w, h, 1);
        ws.launches++;
        SPX_CUDA(cudaGetLastError());
        return SPX_OK;
    }
    const int lw = flavour == 1 ? 1 : (thick_line ? 2 : 1);
    unsigned char* a = ws.takenA;
    unsigned char* b = ws.takenB;
    const int group = lw == 1 ? 4 : 16;
    for (int round = 0; round < 100000; round++) {
        SPX_CUDA(cudaMemsetAsync(ws.flag + 1, 0, sizeof(int), st));
        for (int k = 0; k < group; k++) {
            // only the last sweep of a group reports changes: if it changed nothing we are at the fixed point
            contour_sweep_kernel<<<g2, b2, 0, st>>>(ws.diff, a, b, w, h, lw, k == group - 1 ? ws.flag + 1 : ws.flag + 2);
            unsigned char* t = a; a = b; b = t;
        }
        ws.launches += group;
        SPX_CUDA(cudaMemcpyAsync(ws.h_flag + 1, ws.flag + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        SPX_CUDA(cudaStreamSynchronize(st));
        if (!ws.h_flag[1]) break;
    }
    contour_emit_kernel<0><<<g2, b2, 0, st>>>(ws.diff, a, d_mask, mp, w, h, 0);
    ws.launches++;
    SPX_CUDA(cudaGetLastError());
    return SPX_OK;
}

}  // namespace spx

// ------------------------------------------------------------------------------------------------
// stand-alone C ABI (host buffers)
// ------------------------------------------------------------------------------------------------
extern "C" int spx_enforce_label_connectivity(int32_t* labels, size_t stride, int width, int height,
                                              int numlabels, int min_element_size, int device, int* numlabels_out)
{
    using namespace spx;
    SPX_REQUIRE(labels && width > 0 && height > 0 && stride >= (size_t)width * 4 && stride % 4 == 0, SPX_ERR_BADARG,
                "enforceLabelConnectivity: bad argument");
    SPX_CHECK(require_device(device));
    PostWorkspace ws;
    int rc = ws.alloc(width, height);
    int* d = nullptr;
    if (rc == SPX_OK && cudaMalloc(&d, (size_t)width * height * 4) != cudaSuccess) { set_error("cudaMalloc failed"); rc = SPX_ERR_NOMEM; }
    if (rc == SPX_OK && cudaMemcpy2D(d, (size_t)width * 4, labels, stride, (size_t)width * 4, height, cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("H2D copy failed"); rc = SPX_ERR_GPU;
    }
    int k = numlabels;
    if (rc == SPX_OK) rc = enforce_connectivity_dev(ws, d, width, numlabels, min_element_size, 0, &k);
    if (rc == SPX_OK && cudaMemcpy2D(labels, stride, d, (size_t)width * 4, (size_t)width * 4, height, cudaMemcpyDeviceToHost) != cudaSuccess) {
        set_error("D2H copy failed"); rc = SPX_ERR_GPU;
    }
    if (numlabels_out) *numlabels_out = k;
    cudaFree(d);
    ws.release();
    return rc;
}

extern "C" int spx_label_contour_mask(const int32_t* labels, size_t stride, int width, int height, int thick_line,
                                      int flavour, uint8_t* dst, size_t dst_stride, int device)
{
    using namespace spx;
    SPX_REQUIRE(labels && dst && width > 0 && height > 0 && stride >= (size_t)width * 4 && stride % 4 == 0 &&
                    dst_stride >= (size_t)width && (flavour == 0 || flavour == 1),
                SPX_ERR_BADARG, "getLabelContourMask: bad argument");
    SPX_CHECK(require_device(device));
    PostWorkspace ws;
    int rc = ws.alloc(width, height);
    int* d = nullptr;
    unsigned char* m = nullptr;
    if (rc == SPX_OK && (cudaMalloc(&d, (size_t)width * height * 4) != cudaSuccess || cudaMalloc(&m, (size_t)width * height) != cudaSuccess)) {
        set_error("cudaMalloc failed"); rc = SPX_ERR_NOMEM;
    }
    if (rc == SPX_OK && cudaMemcpy2D(d, (size_t)width * 4, labels, stride, (size_t)width * 4, height, cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("H2D copy failed"); rc = SPX_ERR_GPU;
    }
    if (rc == SPX_OK) rc = contour_mask_dev(ws, d, width, thick_line, flavour, m, width, 0);
    if (rc == SPX_OK && cudaMemcpy2D(dst, dst_stride, m, width, width, height, cudaMemcpyDeviceToHost) != cudaSuccess) {
        set_error("D2H copy failed"); rc = SPX_ERR_GPU;
    }
    cudaFree(d); cudaFree(m);
    ws.release();
    return rc;
}
