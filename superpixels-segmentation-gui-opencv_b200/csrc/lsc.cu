// lsc.cu -- SuperpixelLSC engine and its C ABI (replaces createSuperpixelLSC / iterate / enforceLabelConnectivity /
// getLabels / getNumberOfSuperpixels / getLabelContourMask as called from mainwindow.cpp:2114-2120).
//
// PARITY UNPINNED (no lsc.cpp, no ximgproc binary, no fixture): implements, operation for operation, the restatement
// of Li & Chen's Linear Spectral Clustering that the test oracle under oracle/ (LSC section) documents.  Design:
//   * the 2C+4 sin/cos features are never materialised (44 B/px upstream): every kernel recomputes them in registers
//     from the 8-bit pixel and its coordinates through small tables (256 entries per channel, W and H entries for x/y)
//     that the host builds once per image in double precision;
//   * feature means come from exact per-channel histograms (GPU) weighted by the tables (host, 3*256 + W + H terms);
//   * the assignment is a gather over the 3x3 bins of current seeds (a seed's window is +-step, inclusive), arg-min
//     in (distance, seed index) order = upstream's ascending scan with strict '<';
//   * every sum over pixels is an exact fixed-point int64 (order free): warp-level segmented reduction over runs of
//     equal labels, one 64-bit global atomic per run and field;
//   * connectivity: pre-pass = 8-connected components (union-find, root = first pixel in raster order) + the authors'
//     `adj` rule resolved by relaxation over component roots; post-pass = 4-connected regions, region statistics and
//     pixel lists in parallel, then the inherently sequential nearest-neighbour merge of small regions on the region
//     graph by a single CTA (one step per small region), then a scan for the final ids.
#include "common.cuh"
#include "f32x2.cuh"
#include <cub/device/device_scan.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <vector>

namespace spx {

constexpr int LSC_MAXC = 4;
constexpr int LSC_MAXD = 2 * LSC_MAXC + 4;
constexpr int LSC_UNK = -2;

struct LscDev {
    int w, h, C, D, K, stepx, stepy, nbx, nby;
    size_t ipitch;
    const unsigned char* img;
    const float *colcos, *colsin, *wcol;          // [C][256]
    const float *xcos, *xsin, *wx, *ycos, *ysin, *wy;
    int *sx, *sy;                                 // seeds
    float* cen;                                   // K x D
    long long* sums;                              // K x (D+4): D features, W, x, y, count
    int *head, *next;                             // bins of stepx x stepy pixels
    // tiled assignment (lsc_tile.cuh): colour tables interleaved per value (cos, sin, w, 0) [3][256]; per 32x32 tile the seeds
    // whose window meets it, appended by the kernels that move the seeds (null: not tiled)
    const float4* ctab;
    int *tl_count, *tl_k;
    int tiles_x, tiles_y;
};

template <int C>
__device__ __forceinline__ float lsc_features(const LscDev& L, int x, int y, float (&f)[2 * C + 4])
{
    const unsigned char* p = L.img + (size_t)y * L.ipitch + (size_t)x * C;
    int v[C];
#pragma unroll
    for (int c = 0; c < C; c++) v[c] = p[c];
    float wc = L.wcol[v[0]];
#pragma unroll
    for (int c = 1; c < C; c++) wc = __fadd_rn(wc, L.wcol[c * 256 + v[c]]);
    const float Wt = __fadd_rn(wc, __fadd_rn(L.wx[x], L.wy[y]));
    const float inv = __fdiv_rn(1.0f, Wt);
#pragma unroll
    for (int c = 0; c < C; c++) {
        f[2 * c] = __fmul_rn(L.colcos[c * 256 + v[c]], inv);
        f[2 * c + 1] = __fmul_rn(L.colsin[c * 256 + v[c]], inv);
    }
    f[2 * C] = __fmul_rn(L.xcos[x], inv); f[2 * C + 1] = __fmul_rn(L.xsin[x], inv);
    f[2 * C + 2] = __fmul_rn(L.ycos[y], inv); f[2 * C + 3] = __fmul_rn(L.ysin[y], inv);
    return Wt;
}

// ---- warp-level segmented sums over runs of equal keys in consecutive lanes; the last lane of a run adds to global
struct RunInfo { int start; bool tail; };
__device__ __forceinline__ RunInfo run_info(int key, int lane)
{
    const int prev = __shfl_up_sync(0xffffffffu, key, 1);
    const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || prev != key);
    RunInfo r;
    r.start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
    r.tail = lane == 31 || ((heads >> (lane + 1)) & 1u);
    return r;
}
__device__ __forceinline__ long long run_sum(long long v, const RunInfo& r, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const long long u = __shfl_up_sync(0xffffffffu, v, d);
        if (lane - d >= r.start) v += u;
    }
    return v;
}

// ---- histograms per channel (exact feature means) ------------------------------------------------------------
__global__ void __launch_bounds__(256) lsc_hist_kernel(const unsigned char* __restrict__ img, size_t ipitch, int w, int h, int C,
                                                        unsigned long long* __restrict__ hist)
{
    __shared__ unsigned sh[LSC_MAXC * 256];
    for (int i = threadIdx.x; i < C * 256; i += 256) sh[i] = 0;
    __syncthreads();
    const int y = blockIdx.y;
    for (int x = blockIdx.x * 256 + threadIdx.x; x < w; x += gridDim.x * 256)
        for (int c = 0; c < C; c++) atomicAdd(&sh[c * 256 + img[(size_t)y * ipitch + (size_t)x * C + c]], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < C * 256; i += 256)
        if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
}

// ---- initial centres: plain mean of the normalised features over +-step/4, one warp per seed ----------------------
template <int C>
__global__ void __launch_bounds__(128) lsc_init_centres_kernel(LscDev L)
{
    constexpr int D = 2 * C + 4;
    const int k = (blockIdx.x * 128 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (k >= L.K) return;
    const int sx = L.sx[k], sy = L.sy[k];
    const int x0 = max(sx - L.stepx / 4, 0), x1 = min(sx + L.stepx / 4, L.w - 1);
    const int y0 = max(sy - L.stepy / 4, 0), y1 = min(sy + L.stepy / 4, L.h - 1);
    const int ww = x1 - x0 + 1, n = ww * (y1 - y0 + 1);
    long long s[D];
#pragma unroll
    for (int i = 0; i < D; i++) s[i] = 0;
    float f[D];
    for (int e = lane; e < n; e += 32) {
        lsc_features<C>(L, x0 + e % ww, y0 + e / ww, f);
#pragma unroll
        for (int i = 0; i < D; i++) s[i] += __float2ll_rz(__fmul_rn(f[i], 1099511627776.0f));
    }
#pragma unroll
    for (int i = 0; i < D; i++) {
        long long v = s[i];
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) L.cen[(size_t)k * D + i] = (float)__ddiv_rn(__ddiv_rn((double)v, (double)n), 1099511627776.0);
    }
}

constexpr int LSC_CAP = 32;                        // candidates per tile of the tiled assignment
constexpr int LSC_TH = 32;                         // tile height (32 x 64 tiles measured: same speed at S = 30, lists overflow below S = 16)
// seed k into the list of every 32 x LSC_TH tile its window [s - step, s + step] meets
__device__ __forceinline__ void lsc_tile_append(const LscDev& L, int k, int sx, int sy)
{
    if (!L.tl_count) return;
    const int tx0 = max(sx - L.stepx, 0) >> 5, tx1 = min(sx + L.stepx, L.w - 1) >> 5;
    const int ty0 = max(sy - L.stepy, 0) / LSC_TH, ty1 = min(sy + L.stepy, L.h - 1) / LSC_TH;
    for (int ty = ty0; ty <= ty1; ty++)
        for (int tx = tx0; tx <= tx1; tx++) {
            const int t = ty * L.tiles_x + tx;
            const int pos = atomicAdd(&L.tl_count[t], 1);
            if (pos < LSC_CAP) L.tl_k[(size_t)t * LSC_CAP + pos] = k;
        }
}

__global__ void lsc_bin_kernel(LscDev L)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= L.K) return;
    const int bx = min(max(L.sx[k], 0) / L.stepx, L.nbx - 1), by = min(max(L.sy[k], 0) / L.stepy, L.nby - 1);
    L.next[k] = atomicExch(&L.head[by * L.nbx + bx], k);
    lsc_tile_append(L, k, L.sx[k], L.sy[k]);
}

// ---- assignment + exact accumulation -----------------------------------------------------------------------------
// The CTA's 32x8 pixels search the bins (bx-1..bx+1) x (by-1..by+1) of their own bin; the union over the CTA is a small
// block of bins, so its seeds and centres are staged in shared memory once (a thread per bin walks the list) and every pixel
// loops over the staged set with the window test -- instead of each pixel chasing nine linked lists through L2.  The explicit
// (d, k) tie-break makes the result independent of the staging order.  More than LSC_STAGE seeds: the per-pixel walk.
constexpr int LSC_STAGE = 96;
template <int C>
__global__ void __launch_bounds__(256) lsc_assign_kernel(LscDev L, int* __restrict__ labels, int do_assign)
{
    constexpr int D = 2 * C + 4;
    __shared__ int s_n;
    __shared__ int s_k[LSC_STAGE];
    __shared__ int2 s_xy[LSC_STAGE];
    __shared__ float s_cen[LSC_STAGE][D];
    const int lane = threadIdx.x & 31;
    const int x = blockIdx.x * 32 + lane, y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const bool inside = x < L.w && y < L.h;
    int staged = -1;                                              // -1: per-pixel walk
    if (do_assign) {
        if (threadIdx.x == 0) s_n = 0;
        __syncthreads();
        const int X0 = blockIdx.x * 32, Y0 = blockIdx.y * 8;
        const int bx0 = max(X0 / L.stepx - 1, 0), bx1 = min(min(X0 + 31, L.w - 1) / L.stepx + 1, L.nbx - 1);
        const int by0 = max(Y0 / L.stepy - 1, 0), by1 = min(min(Y0 + 7, L.h - 1) / L.stepy + 1, L.nby - 1);
        const int nbw = bx1 - bx0 + 1, nb = nbw * (by1 - by0 + 1);
        for (int b = threadIdx.x; b < nb; b += 256)
            for (int k = L.head[(by0 + b / nbw) * L.nbx + bx0 + b % nbw]; k >= 0; k = L.next[k]) {
                const int pos = atomicAdd(&s_n, 1);
                if (pos < LSC_STAGE) s_k[pos] = k;
            }
        __syncthreads();
        const int n = s_n;
        if (n <= LSC_STAGE) {
            staged = n;
            for (int i = threadIdx.x; i < n * (D + 2); i += 256) {
                const int j = i / (D + 2), f = i % (D + 2), k = s_k[j];
                if (f < D) s_cen[j][f] = L.cen[(size_t)k * D + f];
                else if (f == D) s_xy[j].x = L.sx[k];
                else s_xy[j].y = L.sy[k];
            }
        }
        __syncthreads();
    }
    float f[D];
    float Wt = 0.f;
    int label = -1 - lane;
    if (inside) {
        Wt = lsc_features<C>(L, x, y, f);
        const size_t li = (size_t)y * L.w + x;
        if (do_assign) {
            float best = FLT_MAX;
            int bestk = -1;
            if (staged >= 0) {
                for (int j = 0; j < staged; j++) {
                    const int2 sxy = s_xy[j];
                    if (abs(x - sxy.x) > L.stepx || abs(y - sxy.y) > L.stepy) continue;
                    float d = 0.f;
#pragma unroll
                    for (int i = 0; i < D; i++) { const float t = __fsub_rn(f[i], s_cen[j][i]); d = __fadd_rn(d, __fmul_rn(t, t)); }
                    const int k = s_k[j];
                    if (d < best || (d == best && k < bestk)) { best = d; bestk = k; }
                }
            } else {
                const int bx = x / L.stepx, by = y / L.stepy;
                for (int yy = max(by - 1, 0); yy <= min(by + 1, L.nby - 1); yy++)
                    for (int xx = max(bx - 1, 0); xx <= min(bx + 1, L.nbx - 1); xx++)
                        for (int k = L.head[yy * L.nbx + xx]; k >= 0; k = L.next[k]) {
                            if (abs(x - L.sx[k]) > L.stepx || abs(y - L.sy[k]) > L.stepy) continue;
                            const float* c = L.cen + (size_t)k * D;
                            float d = 0.f;
#pragma unroll
                            for (int i = 0; i < D; i++) { const float t = __fsub_rn(f[i], c[i]); d = __fadd_rn(d, __fmul_rn(t, t)); }
                            if (d < best || (d == best && k < bestk)) { best = d; bestk = k; }
                        }
            }
            if (bestk >= 0) { labels[li] = bestk; label = bestk; }
            else label = labels[li];                       // covered by no window: keeps its label
        } else label = labels[li];
    }
    const RunInfo r = run_info(label, lane);
    long long* dst = L.sums + (size_t)(label >= 0 ? label : 0) * (D + 4);
    const bool emit = inside && r.tail;
#pragma unroll
    for (int i = 0; i < D; i++) {
        const long long q = inside ? __float2ll_rz(__fmul_rn(__fmul_rn(Wt, f[i]), 4294967296.0f)) : 0ll;
        const long long s = run_sum(q, r, lane);
        if (emit) atomicAdd((unsigned long long*)&dst[i], (unsigned long long)s);
    }
    {
        long long s = run_sum(inside ? __float2ll_rz(__fmul_rn(Wt, 16777216.0f)) : 0ll, r, lane);
        if (emit) atomicAdd((unsigned long long*)&dst[D], (unsigned long long)s);
        s = run_sum(inside ? (long long)x : 0ll, r, lane);
        if (emit) atomicAdd((unsigned long long*)&dst[D + 1], (unsigned long long)s);
        if (emit) {
            const long long cnt = lane - r.start + 1;
            atomicAdd((unsigned long long*)&dst[D + 2], (unsigned long long)(cnt * y));
            atomicAdd((unsigned long long*)&dst[D + 3], (unsigned long long)cnt);
        }
    }
}

#include "lsc_tile.cuh"

// centres = W-weighted means, seeds = integer mean position; sums cleared; seeds re-binned
__global__ void lsc_finalize_kernel(LscDev L)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= L.K) return;
    long long* a = L.sums + (size_t)k * (L.D + 4);
    const long long sw = a[L.D], cnt = a[L.D + 3];
    for (int i = 0; i < L.D; i++) {
        L.cen[(size_t)k * L.D + i] = (cnt > 0 && sw != 0) ? (float)__ddiv_rn(__ddiv_rn((double)a[i], (double)sw), 256.0) : 0.f;
        a[i] = 0;
    }
    const int sx = cnt > 0 ? (int)(a[L.D + 1] / cnt) : 0, sy = cnt > 0 ? (int)(a[L.D + 2] / cnt) : 0;
    a[L.D] = 0; a[L.D + 1] = 0; a[L.D + 2] = 0; a[L.D + 3] = 0;
    L.sx[k] = sx; L.sy[k] = sy;
    const int bx = min(sx / L.stepx, L.nbx - 1), by = min(sy / L.stepy, L.nby - 1);
    L.next[k] = atomicExch(&L.head[by * L.nbx + bx], k);
    lsc_tile_append(L, k, sx, sy);
}

// ---- connected components (union-find, root = smallest raster index) ----------------------------------------------
__device__ __forceinline__ int lsc_find(const int* P, int a)
{
    int p = P[a];
    while (p != a) { a = p; p = P[a]; }
    return a;
}
__device__ __forceinline__ void lsc_unite(int* P, int a, int b)
{
    bool done;
    do {
        a = lsc_find(P, a); b = lsc_find(P, b);
        if (a < b) { int old = atomicMin(&P[b], a); done = (old == b); b = old; }
        else if (b < a) { int old = atomicMin(&P[a], b); done = (old == a); a = old; }
        else done = true;
    } while (!done);
}
// parent[i] = first pixel of the horizontal run of equal labels that holds i (one ballot per warp; a run that began in an
// earlier warp chains through the pixel left of the warp's first lane): finds start one or two hops from a run head
__global__ void lsc_ccl_init_kernel(const int* __restrict__ lab, int w, int h, int* __restrict__ P)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const bool in = x < w;
    const int i = y * w + x;
    const bool start = !in || x == 0 || lab[i - 1] != lab[i];
    const unsigned heads = __ballot_sync(0xffffffffu, start) & (0xffffffffu >> (31 - lane));
    if (!in) return;
    P[i] = heads ? i - (lane - (31 - __clz(heads))) : i - lane - 1;
}
// unions with the row above; a union that a horizontal neighbour already implies is skipped:
//   up:        implied when the left neighbour and its upper neighbour carry the same label (left ~ up-left ~ up by the run above)
//   up-left:   (only when up differs) implied when the left neighbour has the label (its own "up" union reaches up-left)
//   up-right:  (only when up differs) implied when the right neighbour has the label
template <int CONN>
__global__ void lsc_ccl_merge_kernel(const int* __restrict__ lab, int w, int h, int* __restrict__ P)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y + 1;
    if (x >= w || y >= h) return;
    const int i = y * w + x, me = lab[i];
    const bool left = x > 0 && lab[i - 1] == me;
    if (lab[i - w] == me) {
        if (!(left && lab[i - w - 1] == me)) lsc_unite(P, i, i - w);
        return;
    }
    if (CONN == 8) {
        if (x > 0 && lab[i - w - 1] == me && !left) lsc_unite(P, i, i - w - 1);
        if (x < w - 1 && lab[i - w + 1] == me && lab[i + 1] != me) lsc_unite(P, i, i - w + 1);
    }
}
// flatten; rootidx = own index for roots, -1 otherwise; flag = 1 for roots
__global__ void lsc_ccl_flatten_kernel(int* __restrict__ P, int* __restrict__ csize, int* __restrict__ rootidx, int* __restrict__ flag, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int r = lsc_find(P, i);
    P[i] = r;
    csize[i] = 0;
    rootidx[i] = r == i ? i : -1;
    flag[i] = r == i ? 1 : 0;
}
__global__ void lsc_ccl_size_kernel(const int* __restrict__ P, int* __restrict__ csize, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int key = i < n ? P[i] : -1 - lane;
    const RunInfo r = run_info(key, lane);
    if (i < n && r.tail) atomicAdd(&csize[key], lane - r.start + 1);
}

// ---- pre-pass: the authors' `adj` rule on the component roots (see the test oracle, LSC pre-pass) --------
__global__ void lsc_pre_init_kernel(const int* __restrict__ P, const int* __restrict__ lab, const int* __restrict__ csize,
                                    int* __restrict__ FL, int* __restrict__ ADJ, int n, int bound)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || P[i] != i) return;
    FL[i] = csize[i] <= bound ? LSC_UNK : lab[i];
    ADJ[i] = LSC_UNK;
}
// adj_k = (final label of the component left of / above the root differs from L_k) ? that label : adj_{k-1},
// relaxed over the compacted list of component roots (1.1 M roots instead of 33 M pixels at 8K)
__global__ void __launch_bounds__(256) lsc_pre_relax_list_kernel(const int* __restrict__ roots, const int* __restrict__ d_nroots,
                                                                  const int* __restrict__ P, const int* __restrict__ lab,
                                                                  const int* __restrict__ csize, const int* __restrict__ prevroot,
                                                                  volatile int* FL, volatile int* ADJ, int w, int bound, int* __restrict__ pending)
{
    const int nr = *d_nroots;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nr; t += gridDim.x * blockDim.x) {
        const int i = roots[t];
        if (ADJ[i] != LSC_UNK) continue;
        const int x = i % w;
        const int nb = x > 0 ? i - 1 : (i >= w ? i - w : -1);
        int adj;
        if (nb < 0) adj = 0;                       // pixel 0: nothing is visited yet, adj keeps its initial value
        else {
            const int fl = FL[P[nb]];
            if (fl == LSC_UNK) { *pending = 1; continue; }
            if (fl != lab[i]) adj = fl;
            else {
                const int pr = prevroot[i];
                if (pr < 0) adj = 0;
                else {
                    const int a = ADJ[pr];
                    if (a == LSC_UNK) { *pending = 1; continue; }
                    adj = a;
                }
            }
        }
        if (csize[i] <= bound) FL[i] = adj;
        __threadfence();
        ADJ[i] = adj;
    }
}
__global__ void lsc_pre_relabel_kernel(const int* __restrict__ P, const int* __restrict__ FL, int* __restrict__ lab, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) lab[i] = FL[P[i]];
}

// ---- post-pass ---------------------------------------------------------------------------------------------------
// reg[px] = region id (raster order of the roots); region size; pixel lists
__global__ void lsc_post_region_kernel(const int* __restrict__ P, const int* __restrict__ rid, const int* __restrict__ csize,
                                       int* __restrict__ reg, int* __restrict__ size, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int root = P[i], r = rid[root];
    reg[i] = r;
    if (root == i) size[r] = csize[i];
}
template <int C>
__global__ void __launch_bounds__(256) lsc_post_sums_kernel(LscDev L, const int* __restrict__ reg, long long* __restrict__ rsum)
{
    constexpr int D = 2 * C + 4;
    const int lane = threadIdx.x & 31;
    const int x = blockIdx.x * 32 + lane, y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const bool inside = x < L.w && y < L.h;
    float f[D];
    float Wt = 0.f;
    int key = -1 - lane;
    if (inside) { Wt = lsc_features<C>(L, x, y, f); key = reg[(size_t)y * L.w + x]; }
    const RunInfo r = run_info(key, lane);
    long long* dst = rsum + (size_t)(key >= 0 ? key : 0) * (D + 1);
    const bool emit = inside && r.tail;
#pragma unroll
    for (int i = 0; i < D; i++) {
        const long long s = run_sum(inside ? __float2ll_rz(__fmul_rn(__fmul_rn(Wt, f[i]), 4294967296.0f)) : 0ll, r, lane);
        if (emit) atomicAdd((unsigned long long*)&dst[i], (unsigned long long)s);
    }
    const long long s = run_sum(inside ? __float2ll_rz(__fmul_rn(Wt, 16777216.0f)) : 0ll, r, lane);
    if (emit) atomicAdd((unsigned long long*)&dst[D], (unsigned long long)s);
}
__global__ void lsc_post_stats_kernel(const long long* __restrict__ rsum, const int* __restrict__ size, int D, int R, int threshold,
                                      float* __restrict__ cen, float* __restrict__ Wr, int* __restrict__ alive,
                                      int* __restrict__ smallflag, int* __restrict__ done, int* __restrict__ cur)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const long long* a = rsum + (size_t)r * (D + 1);
    const long long sw = a[D];
    for (int j = 0; j < D; j++) cen[(size_t)r * D + j] = sw != 0 ? (float)__ddiv_rn(__ddiv_rn((double)a[j], (double)sw), 256.0) : 0.f;
    Wr[r] = (float)__ddiv_rn((double)sw, 16777216.0);
    alive[r] = 1; done[r] = 0; cur[r] = r;
    smallflag[r] = size[r] < threshold ? 1 : 0;
}

// ---- the merge of small regions -------------------------------------------------------------------------------
// Upstream's rule is sequential: for i = 0, 1, 2, ...: if region i is (still) smaller than the threshold it merges into
// the adjacent region with the nearest centre (ties -> smaller id) and that region's centre / weight / size are updated.
// A merge of region j can change the decision of a later region i only if i and j are adjacent or share a neighbour
// (the merge target).  So a pending region that is the smallest pending index within graph distance 2 takes exactly
// the decision it would take in the sequential order, and all such regions can merge in the same round: they are
// pairwise more than 2 apart, hence their targets are distinct and nobody reads what another one writes.
// One round = a few passes over the pixel map (min-propagation of pending indices over region adjacencies, nearest
// neighbour search for the safe regions, relabel); rounds repeat until nothing is pending.
// visiting priority of a small region: order 0 = its index (upstream's raster order; the dependency chains then span the
// image, thousands of rounds), order 1 = a bijective hash of the index (chains of O(100) rounds; the default).
__device__ __forceinline__ unsigned lsc_prio(int r, int order)
{
    unsigned x = (unsigned)r;
    if (order) {
        x *= 2654435761u; x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
    }
    return x;
}
// region adjacency: one 64-bit key (min<<32 | max) per 4-neighbour pixel pair with different regions; sorted and
// de-duplicated once, the rounds then work on the (much smaller) edge list.  Merged regions are followed through `cur`.
__global__ void lsc_edge_count_kernel(const int* __restrict__ reg, int w, int h, unsigned long long* __restrict__ keys,
                                      unsigned long long* __restrict__ count, unsigned long long cap)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const int i = y * w + x, rp = reg[i];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        if (k == 0 ? x + 1 >= w : y + 1 >= h) continue;
        const int rq = reg[k == 0 ? i + 1 : i + w];
        if (rq == rp) continue;
        // skip the pair when the pixel before it on the same boundary run already emitted it (cheap partial dedupe)
        if (k == 0 && y > 0 && reg[i - w] == rp && reg[i - w + 1] == rq) continue;
        if (k == 1 && x > 0 && reg[i - 1] == rp && reg[i + w - 1] == rq) continue;
        const unsigned lo = (unsigned)min(rp, rq), hi = (unsigned)max(rp, rq);
        const unsigned long long pos = atomicAdd(count, 1ull);
        if (pos < cap) keys[pos] = ((unsigned long long)lo << 32) | hi;
    }
}
// (no __restrict__ / read-only path: the persistent rounds kernel reads what other CTAs wrote before a grid barrier)
__device__ __forceinline__ int lsc_cur(const int* cur, int r)
{
    int c = cur[r];
    while (c != r) { r = c; c = cur[r]; }
    return r;
}
__device__ __forceinline__ unsigned long long lsc_pair_key(const float* cen, int D, int i, int r)
{
    float d = 0.f;
    for (int j = 0; j < D; j++) { const float u = __fsub_rn(cen[(size_t)i * D + j], cen[(size_t)r * D + j]); d = __fadd_rn(d, __fmul_rn(u, u)); }
    return ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)r;
}
// ---- all merge rounds in ONE persistent cooperative kernel, on shrinking work lists.
// A round only concerns the pending regions and the adjacency edges with a pending end: priorities are finite at pending
// regions only, so min-propagation over distance 1 and 2, the choice of the best neighbour and the merge never read anything
// else.  A region that stops being pending never becomes pending again (sizes only grow, only pending regions move), so both
// lists shrink monotonically: the kernel keeps a pending list and a live-edge list (edges rewritten to their current
// representatives every round), compacts them with warp-aggregated appends, and separates the passes by grid-wide barriers.
// 8K frame: 1.05 M regions (1.03 M of them small), 1.7 M edges, 195 rounds; the pending list shrinks by 2-4 % per round (the
// fragments around one superpixel are all within distance 2 of each other, so one of them merges per round): 17.8 ms, against
// 23.2 ms for the kernel-per-pass form that swept every region and edge in every round.
struct LscRounds {
    unsigned long long* edges[2]; int E;          // live edges, double buffered
    int* plist[2];                                // pending regions, double buffered
    const int* smallflag; int* done; int* size; int threshold; int R; int order;
    uint4* rr;                                    // per region (ra, rb, pending, -): one 16-byte record, one sector per visit
    unsigned long long* best;
    float* cen; int D; float* Wr; int* alive; int* cur;
    int* cnt;                                     // [0..3] pending-list lengths by round & 3, [4..7] edge-list lengths, [8] barrier, [9] rounds
    int max_rounds;
};
__device__ __forceinline__ void lsc_grid_barrier(unsigned* counter, unsigned target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        while (*(volatile unsigned*)counter < target) { }
        __threadfence();
    }
    __syncthreads();
}
// all 32 lanes call; returns the slot of a kept item
__device__ __forceinline__ int lsc_warp_append(bool keep, int* counter, int lane)
{
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    int base = 0;
    if (lane == 0 && m) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    return base + __popc(m & ((1u << lane) - 1));
}
__global__ void __launch_bounds__(1024) lsc_rounds_kernel(LscRounds a)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x, lane = threadIdx.x & 31;
    const unsigned nblk = gridDim.x;
    unsigned* bar = reinterpret_cast<unsigned*>(a.cnt + 8);
    unsigned epoch = 0;
    int round = 0;
    int nP = a.R, nE = a.E;                         // round 0 scans every region / every edge
    for (; round < a.max_rounds; round++) {
        const int* pin = a.plist[round & 1]; int* pout = a.plist[(round + 1) & 1];
        const unsigned long long* ein = a.edges[round & 1]; unsigned long long* eout = a.edges[(round + 1) & 1];
        int* cP = a.cnt + (round & 3); int* cE = a.cnt + 4 + (round & 3);
        if (tid == 0) { a.cnt[(round + 2) & 3] = 0; a.cnt[4 + ((round + 2) & 3)] = 0; }     // slots of round+2: idle now
        // ---- pending test over the previous pending list (round 0: all regions)
        for (int i0 = tid - lane; i0 < nP; i0 += nth) {
            const int i = i0 + lane;
            int r = -1;
            bool p = false;
            if (i < nP) {
                r = round ? pin[i] : i;
                p = a.smallflag[r] && !a.done[r] && a.size[r] < a.threshold;
                const unsigned pr = p ? lsc_prio(r, a.order) : 0xffffffffu;
                a.rr[r] = make_uint4(pr, pr, p ? 1u : 0u, 0u);
                if (p) a.best[r] = ~0ull;
            }
            const int slot = lsc_warp_append(p, cP, lane);
            if (p) pout[slot] = r;
        }
        lsc_grid_barrier(bar, ++epoch * nblk);
        nP = *(volatile int*)cP;
        if (nP == 0) break;                                                      // uniform
        // ---- live edges: current representatives, at least one pending end; non-pending ends start at +inf
        for (int e0 = tid - lane; e0 < nE; e0 += nth) {
            const int e = e0 + lane;
            bool keep = false;
            int rp = 0, rq = 0;
            if (e < nE) {
                const unsigned long long k = ein[e];
                rp = lsc_cur(a.cur, (int)(k >> 32)); rq = lsc_cur(a.cur, (int)(k & 0xffffffffull));
                const bool pp = a.rr[rp].z, pq = a.rr[rq].z;
                keep = rp != rq && (pp || pq);
                if (keep) { if (!pp) a.rr[rp].x = 0xffffffffu; if (!pq) a.rr[rq].x = 0xffffffffu; }
            }
            const int slot = lsc_warp_append(keep, cE, lane);
            if (keep) eout[slot] = ((unsigned long long)(unsigned)rp << 32) | (unsigned)rq;
        }
        lsc_grid_barrier(bar, ++epoch * nblk);
        nE = *(volatile int*)cE;
        // ---- ra = min pending priority over the closed neighbourhood
        for (int e = tid; e < nE; e += nth) {
            const unsigned long long k = eout[e];
            const int rp = (int)(k >> 32), rq = (int)(k & 0xffffffffull);
            const uint4 P = a.rr[rp], Q = a.rr[rq];
            if (Q.z) { const unsigned v = lsc_prio(rq, a.order); if (v < P.x) atomicMin(&a.rr[rp].x, v); }
            if (P.z) { const unsigned v = lsc_prio(rp, a.order); if (v < Q.x) atomicMin(&a.rr[rq].x, v); }
        }
        lsc_grid_barrier(bar, ++epoch * nblk);
        // ---- rb (pending regions only) = the same over distance 2
        for (int e = tid; e < nE; e += nth) {
            const unsigned long long k = eout[e];
            const int rp = (int)(k >> 32), rq = (int)(k & 0xffffffffull);
            const uint4 P = a.rr[rp], Q = a.rr[rq];
            const unsigned m = min(P.x, Q.x);
            if (P.z && m < P.y) atomicMin(&a.rr[rp].y, m);
            if (Q.z && m < Q.y) atomicMin(&a.rr[rq].y, m);
        }
        lsc_grid_barrier(bar, ++epoch * nblk);
        // ---- best neighbour of every region that merges in this round
        for (int e = tid; e < nE; e += nth) {
            const unsigned long long k = eout[e];
            const int rp = (int)(k >> 32), rq = (int)(k & 0xffffffffull);
            const uint4 P = a.rr[rp], Q = a.rr[rq];
            if (P.z && P.y == lsc_prio(rp, a.order)) { const unsigned long long key = lsc_pair_key(a.cen, a.D, rp, rq); if (key < a.best[rp]) atomicMin(&a.best[rp], key); }
            if (Q.z && Q.y == lsc_prio(rq, a.order)) { const unsigned long long key = lsc_pair_key(a.cen, a.D, rq, rp); if (key < a.best[rq]) atomicMin(&a.best[rq], key); }
        }
        lsc_grid_barrier(bar, ++epoch * nblk);
        // ---- merge
        for (int q = tid; q < nP; q += nth) {
            const int i = pout[q];
            if (a.rr[i].y != lsc_prio(i, a.order)) continue;
            a.done[i] = 1;
            const unsigned long long k = a.best[i];
            if (k == ~0ull) continue;                                            // no neighbour at all (one region covers the image)
            const int nb = (int)(unsigned)(k & 0xffffffffull);
            const float wn = a.Wr[nb], wi = a.Wr[i];
            for (int j = 0; j < a.D; j++) {
                const float u = __fmul_rn(a.cen[(size_t)nb * a.D + j], wn), v = __fmul_rn(a.cen[(size_t)i * a.D + j], wi);
                a.cen[(size_t)nb * a.D + j] = __fdiv_rn(__fadd_rn(u, v), __fadd_rn(wn, wi));
            }
            a.Wr[nb] = __fadd_rn(wn, wi);
            a.size[nb] += a.size[i];
            a.alive[i] = 0;
            a.cur[i] = nb;
        }
        lsc_grid_barrier(bar, ++epoch * nblk);
    }
    // chains of merges are as long as the rounds they span: flatten once for the relabelling pass
    for (int r = tid; r < a.R; r += nth) a.cur[r] = lsc_cur(a.cur, r);
    if (tid == 0) a.cnt[9] = round;
}
// ---- the authors' visiting order (order 0) without its serial dependency chains ------------------------------------
// In region order every fragment around a superpixel is within graph distance 2 of the next one, so the rounds above
// degenerate into a wavefront (thousands of rounds, a few dozen merges each).  The sequential loop is evaluated as a
// FIXED POINT instead (normative CPU form: the test oracle under oracle/, LSC section, lsc_merge_fixedpoint): region i is visited at
// "time" i and its decision (merge or not, into whom) depends on the decisions of regions j < i only.  Take any table
// of tentative decisions tg[], derive from it the state every region has at every time, re-take the decisions against
// that state, repeat until the table reproduces itself.  After k rounds the first k regions (at least) hold their
// sequential decision, and a table that a FULL round reproduces IS the sequential result (induction over i).
//   children of x = { c : tg[c] == x }, ascending c (c hops into x at time c): counted, scanned, filled, ranked
//   T_x(t)  = x's original (centre, weight, size) folded with M_c for its children c < t in ascending order -- the
//             sequential float update; prefix states are kept per child slot.  A thread walks a region's list and
//             descends (explicit stack) into children whose own state is not there yet
//   M_x     = T_x(x): the state x has when it is visited; depends on smaller indices only
//   rep_t(p)= follow p -> tg[p] while the node merged at a time < t and AFTER it was entered (in a consistent table
//             that always holds; in a tentative one it cuts cycles)
//   decision of x = M_x.size < threshold and arg min over adjacency edges (p,q) with rep_x(p) == x, y = rep_x(q) != x
//             of (|M_x.centre - T_y(x).centre|^2, y)
// Rounds after the first are INCREMENTAL: a changed decision of c (old target o, new target t) moves the trajectories
// of o and t from time c on, and of their owners from the time the changed state is folded in (eff[] = earliest
// affected time, pushed up the chains).  A region whose chain meets such a node is a dirty member; the small regions
// x > that time which own a neighbour of a dirty member are re-decided, everybody else keeps its decision.  The
// marking is a heuristic for speed only: the loop ends when a FULL round changes nothing.
// 8K frame: 1.05 M regions, 1.7 M edges, 87 rounds.
struct LscFix {
    const unsigned long long* edges; int E; int R; int D; int threshold;
    const float* cen; const float* Wr; const int* size; const int* smallflag;
    int* tg;                 // tentative decision: target region, -1 = does not merge
    int* off;                // [R+1] children lists (CSR)
    int* nch;                // [R] counts / fill cursors
    int* tmp; int* ch;       // [R] unsorted / ascending children (tmp doubles as "previous target" between rounds)
    int* pre;                // [R] number of children with a smaller index than the parent
    int* prog;               // [R] child slots folded so far in this round
    float* S;                // [R][D+2] S[c] = state of c's owner right after c is folded in: centre, weight, size (int bits).
                             // Indexed by the child, so a slot keeps its place from round to round and an incremental round
                             // refolds a list from the first child >= eff[owner] only
    unsigned long long* best;
    int* inc_off; int* inc;  // [R+1], [2E] static incidence lists of the region adjacency graph
    int* eff;                // [R] earliest time from which the region's trajectory / decision differs from the last round
    int* need;               // [R] re-decide in this round
    int* dq; int* dtau;      // [R] dirty members and their times (between rounds dtau[x] = 1: x changed its decision)
    int* bsum;               // per-CTA partial sums of the scan
    int* cnt;                // [0..3] changed decisions by round & 3, [4..7] "incomplete" by pass & 3, [8] barrier, [9] rounds,
                             // [10] passes, [11] full rounds, [12..15] dirty members by round & 3, [16..] us per phase
    int* cur; int* alive; int max_rounds;
};
struct LscState { float c[LSC_MAXD]; float W; int size; };
__device__ __forceinline__ LscState lsc_fix_load(const LscFix& a, int y, int k)
{
    LscState s;
    if (k == 0) {
#pragma unroll
        for (int j = 0; j < LSC_MAXD; j++) s.c[j] = j < a.D ? a.cen[(size_t)y * a.D + j] : 0.f;
        s.W = a.Wr[y]; s.size = a.size[y];
    } else {
        const float* p = a.S + (size_t)a.ch[a.off[y] + k - 1] * (a.D + 2);
#pragma unroll
        for (int j = 0; j < LSC_MAXD; j++) s.c[j] = j < a.D ? __ldcg(p + j) : 0.f;
        s.W = __ldcg(p + a.D); s.size = __float_as_int(__ldcg(p + a.D + 1));
    }
    return s;
}
__device__ __forceinline__ int lsc_fix_size(const LscFix& a, int y, int k)
{
    return k == 0 ? a.size[y] : __float_as_int(__ldcg(a.S + (size_t)a.ch[a.off[y] + k - 1] * (a.D + 2) + a.D + 1));
}
// grid-wide exclusive scan of cntv[0..R) into offv[0..R] (cntv is zeroed); CTA b owns a contiguous chunk, thread t of it a run
__device__ __forceinline__ void lsc_grid_scan(int* cntv, int* offv, int R, int* bsum, int* s_warp, unsigned* bar, unsigned& epoch)
{
    const unsigned nblk = gridDim.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chunk = (R + (int)nblk - 1) / (int)nblk, per = (chunk + (int)blockDim.x - 1) / (int)blockDim.x;
    const int lo = min((int)blockIdx.x * chunk + (int)threadIdx.x * per, R), hi = min(min(lo + per, ((int)blockIdx.x + 1) * chunk), R);
    int sum = 0;
    for (int x = lo; x < hi; x++) sum += cntv[x];
    int incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const int w = s_warp[lane];
        int wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xffffffffu, wi, d); if (lane >= d) wi += v; }
        s_warp[lane] = wi - w;
        if (lane == 31) bsum[blockIdx.x] = wi;
    }
    __syncthreads();
    const int texcl = s_warp[warp] + incl - sum;
    lsc_grid_barrier(bar, ++epoch * nblk);
    int base = 0;
    for (int b = lane; b < (int)blockIdx.x; b += 32) base += bsum[b];
#pragma unroll
    for (int d = 16; d; d >>= 1) base += __shfl_xor_sync(0xffffffffu, base, d);
    int run = base + texcl;
    for (int x = lo; x < hi; x++) { offv[x] = run; run += cntv[x]; cntv[x] = 0; }
    if (hi == R && lo < R) offv[R] = run;
    if (R == 0 && blockIdx.x == 0 && threadIdx.x == 0) offv[0] = 0;
    lsc_grid_barrier(bar, ++epoch * nblk);
}
constexpr int LSC_FIX_STACK = 24;
constexpr int LSC_INF = 0x7fffffff;
__global__ void __launch_bounds__(1024) lsc_fixpoint_kernel(LscFix a)
{
    __shared__ int s_warp[32];
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x, lane = threadIdx.x & 31;
    const int gwarp = tid >> 5, nwarps = nth >> 5;
    const unsigned nblk = gridDim.x;
    unsigned* bar = reinterpret_cast<unsigned*>(a.cnt + 8);
    unsigned epoch = 0;
    const int R = a.R, D = a.D;
    int round = 0, passes = 0, fullrounds = 0;
    unsigned long long tlast = 0, tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    auto tick = [&](int ph) {
        if (tid == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); if (ph >= 0) tacc[ph] += t - tlast; tlast = t; }
    };
    tick(-1);
    // ---- once: incidence lists of the adjacency graph; initial table = nobody merges
    for (int x = tid; x < R; x += nth) { a.tg[x] = -1; a.nch[x] = 0; a.prog[x] = 0; a.best[x] = ~0ull; a.eff[x] = LSC_INF; a.need[x] = 0; }
    lsc_grid_barrier(bar, ++epoch * nblk);
    for (int e = tid; e < a.E; e += nth) {
        const unsigned long long key = a.edges[e];
        atomicAdd(&a.nch[(int)(key >> 32)], 1); atomicAdd(&a.nch[(int)(key & 0xffffffffull)], 1);
    }
    lsc_grid_barrier(bar, ++epoch * nblk);
    lsc_grid_scan(a.nch, a.inc_off, R, a.bsum, s_warp, bar, epoch);
    for (int e = tid; e < a.E; e += nth) {
        const unsigned long long key = a.edges[e];
        const int p = (int)(key >> 32), q = (int)(key & 0xffffffffull);
        a.inc[a.inc_off[p] + atomicAdd(&a.nch[p], 1)] = q; a.inc[a.inc_off[q] + atomicAdd(&a.nch[q], 1)] = p;
    }
    lsc_grid_barrier(bar, ++epoch * nblk);
    for (int x = tid; x < R; x += nth) a.nch[x] = 0;
    lsc_grid_barrier(bar, ++epoch * nblk);
    tick(7);
    bool full = true;
    for (; round < a.max_rounds; round++) {
        int* ndq = a.cnt + 12 + (round & 3);
        if (tid == 0) a.cnt[12 + ((round + 2) & 3)] = 0;
        // ---- children lists: count.  Incremental round: push the times of the changed decisions up the chains
        for (int c = tid; c < R; c += nth) {
            const int p = a.tg[c];
            if (p >= 0) atomicAdd(&a.nch[p], 1);
            if (!full && a.dtau[c]) {                                            // c changed its decision in the last round
                for (int side = 0; side < 2; side++) {
                    int n = side ? p : a.tmp[c], tm = c;
                    while (n >= 0) {
                        if (tm < a.eff[n]) atomicMin(&a.eff[n], tm);
                        if (tm >= n) break;                                      // folded in after n's own visit: n's owner never sees it
                        tm = n; n = a.tg[n];
                    }
                }
            }
        }
        lsc_grid_barrier(bar, ++epoch * nblk);
        tick(0);
        // ---- exclusive scan of the counts -> off[]
        lsc_grid_scan(a.nch, a.off, R, a.bsum, s_warp, bar, epoch);
        tick(1);
        // ---- fill (any order).  Dirty members: regions whose chain meets a moved trajectory / changed decision
        for (int c0 = tid - lane; c0 < R; c0 += nth) {
            const int c = c0 + lane;
            bool dirty = false;
            int tau = LSC_INF;
            if (c < R) {
                const int p = a.tg[c];
                if (p >= 0) a.tmp[a.off[p] + atomicAdd(&a.nch[p], 1)] = c;
                if (full) { if (a.smallflag[c]) a.need[c] = 1; }
                else {
                    int n = c, tl = -1;
                    for (;;) { tau = min(tau, a.eff[n]); const int t = a.tg[n]; if (t < 0 || n <= tl) break; tl = n; n = t; }
                    dirty = tau != LSC_INF;
                    if (dirty && a.smallflag[c] && tau < c) a.need[c] = 1;
                }
            }
            if (!full) {
                const int slot = lsc_warp_append(dirty, ndq, lane);
                if (dirty) { a.dq[slot] = c; a.dtau[slot] = tau; }
            }
        }
        lsc_grid_barrier(bar, ++epoch * nblk);
        tick(2);
        // ---- ascending order by rank counting; pre[x] = children that arrive before x's own visit
        for (int x = tid; x < R; x += nth) {
            const int p = a.tg[x];
            if (p >= 0) {
                const int lo = a.off[p], hi = a.off[p + 1];
                int rank = 0;
                for (int j = lo; j < hi; j++) rank += a.tmp[j] < x ? 1 : 0;
                a.ch[lo + rank] = x;
            }
            const int lo = a.off[x], hi = a.off[x + 1];
            const int ex = full ? -1 : a.eff[x];                                 // slots of children < eff[x] are still valid
            int pr = 0, k0 = 0;
            for (int j = lo; j < hi; j++) { const int c = a.tmp[j]; pr += c < x ? 1 : 0; k0 += c < ex ? 1 : 0; }
            a.pre[x] = pr; a.prog[x] = k0;
        }
        // ---- (incremental) a warp per dirty member: the small regions that own one of its neighbours later than tau
        if (!full) {
            const int n = *(volatile int*)ndq;
            for (int i = gwarp; i < n; i += nwarps) {
                const int q = a.dq[i], tau = a.dtau[i];
                for (int j = a.inc_off[q] + lane; j < a.inc_off[q + 1]; j += 32) {
                    int x = a.inc[j], tl = -1;
                    for (;;) {
                        if (x > tau && a.smallflag[x]) a.need[x] = 1;
                        const int t = a.tg[x];
                        if (t < 0 || x <= tl) break;
                        tl = x; x = t;
                    }
                }
            }
        }
        lsc_grid_barrier(bar, ++epoch * nblk);
        tick(3);
        // ---- trajectories: a slot can be folded once the child's own state M_c is complete (prog[c] >= pre[c]);
        //      a thread descends into children that are not, to a bounded depth (then another pass)
        for (;;) {
            int* inc = a.cnt + 4 + (passes & 3);
            if (tid == 0) a.cnt[4 + ((passes + 2) & 3)] = 0;
            bool incomplete = false;
            for (int x = tid; x < R; x += nth) {
                if (*(volatile int*)&a.prog[x] >= a.off[x + 1] - a.off[x]) continue;
                int stack[LSC_FIX_STACK];
                int sp = 1;
                stack[0] = x;
                while (sp > 0) {
                    const int n = stack[sp - 1];
                    const int lo = a.off[n], limit = sp == 1 ? a.off[n + 1] - lo : a.pre[n];
                    int k = *(volatile int*)&a.prog[n];
                    if (k >= limit) { sp--; continue; }
                    __threadfence();
                    LscState st = lsc_fix_load(a, n, k);
                    bool descend = false;
                    const int pn = a.pre[n];
                    const bool owned = a.tg[n] >= 0;
                    for (; k < limit; k++) {
                        const int c = a.ch[lo + k];
                        const int pc = a.pre[c];
                        if (pc > 0) {                                             // (a child without earlier children brings its original state)
                            if (*(volatile int*)&a.prog[c] < pc) {
                                if (sp == LSC_FIX_STACK) { incomplete = true; sp = 0; }
                                else stack[sp++] = c;
                                descend = true;
                                break;
                            }
                            __threadfence();
                        }
                        const LscState sc = lsc_fix_load(a, c, pc);
                        const float wsum = __fadd_rn(st.W, sc.W);
#pragma unroll
                        for (int j = 0; j < LSC_MAXD; j++)
                            if (j < D) st.c[j] = __fdiv_rn(__fadd_rn(__fmul_rn(st.c[j], st.W), __fmul_rn(sc.c[j], sc.W)), wsum);
                        st.W = wsum; st.size += sc.size;
                        float* p = a.S + (size_t)c * (D + 2);
#pragma unroll
                        for (int j = 0; j < LSC_MAXD; j++) if (j < D) __stcg(p + j, st.c[j]);
                        __stcg(p + D, st.W); __stcg(p + D + 1, __int_as_float(st.size));
                        if (owned && k + 1 == pn) { __threadfence(); atomicMax(&a.prog[n], k + 1); }      // M_n is complete: its owner may be waiting
                    }
                    __threadfence();
                    atomicMax(&a.prog[n], k);
                    if (!descend) sp--;
                }
            }
            if (__syncthreads_or(incomplete) && threadIdx.x == 0) *(volatile int*)inc = 1;
            lsc_grid_barrier(bar, ++epoch * nblk);
            passes++;
            if (!*(volatile int*)inc) break;                                     // uniform
        }
        tick(4);
        // ---- candidates: member p of a small region x that is re-decided, every neighbour q of p, seen at time x
        for (int p = tid; p < R; p += nth) {
            int x = p, tarr = -1;
            for (;;) {
                const int tx = a.tg[x];
                if (x > tarr && a.need[x]) {
                    const LscState sx = lsc_fix_load(a, x, a.pre[x]);
                    for (int j = a.inc_off[p]; j < a.inc_off[p + 1]; j++) {
                        int y = a.inc[j], ty = -1;
                        for (;;) { const int t = y < x ? a.tg[y] : -1; if (t < 0 || y <= ty) break; ty = y; y = t; }
                        if (y == x) continue;
                        int lo = a.off[y], hi = a.off[y + 1];
                        const int base = lo;
                        while (lo < hi) { const int mid = (lo + hi) >> 1; if (a.ch[mid] < x) lo = mid + 1; else hi = mid; }
                        const LscState sy = lsc_fix_load(a, y, lo - base);
                        float d = 0.f;
#pragma unroll
                        for (int jj = 0; jj < LSC_MAXD; jj++) if (jj < D) { const float u = __fsub_rn(sx.c[jj], sy.c[jj]); d = __fadd_rn(d, __fmul_rn(u, u)); }
                        const unsigned long long k2 = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)y;
                        if (k2 < a.best[x]) atomicMin(&a.best[x], k2);
                    }
                }
                if (tx >= 0 && x > tarr) { tarr = x; x = tx; } else break;
            }
        }
        lsc_grid_barrier(bar, ++epoch * nblk);
        tick(5);
        // ---- decisions (and the reset of the per-round arrays)
        int* chg = a.cnt + (round & 3);
        if (tid == 0) a.cnt[(round + 2) & 3] = 0;
        int nchanged = 0;
        for (int x = tid; x < R; x += nth) {
            int e = LSC_INF;
            bool ch = false;
            if (a.need[x]) {
                const unsigned long long b = a.best[x];
                const bool m = b != ~0ull && lsc_fix_size(a, x, a.pre[x]) < a.threshold;
                const int t = m ? (int)(unsigned)(b & 0xffffffffull) : -1;
                const int old = a.tg[x];
                if (t != old) { nchanged++; a.tg[x] = t; a.tmp[x] = old; e = x; ch = true; }
                a.best[x] = ~0ull; a.need[x] = 0;
            }
            a.eff[x] = e; a.dtau[x] = ch ? 1 : 0; a.nch[x] = 0;
        }
#pragma unroll
        for (int d = 16; d; d >>= 1) nchanged += __shfl_xor_sync(0xffffffffu, nchanged, d);
        if (lane == 0 && nchanged) atomicAdd(chg, nchanged);
        lsc_grid_barrier(bar, ++epoch * nblk);
        tick(6);
        const int nc = *(volatile int*)chg;                                      // uniform
        if (full) fullrounds++;
        if (nc == 0) { if (full) { round++; break; } full = true; }              // a full round must confirm the table
        else full = nc > (R >> 4);
    }
    // the table is consistent now: owners are plain chains
    for (int r = tid; r < R; r += nth) {
        int y = r;
        for (int t = a.tg[y]; t >= 0; t = a.tg[y]) y = t;
        a.cur[r] = y; a.alive[r] = a.tg[r] < 0 ? 1 : 0;
    }
    if (tid == 0) { a.cnt[9] = round; a.cnt[10] = passes; a.cnt[11] = fullrounds; for (int i = 0; i < 8; i++) a.cnt[16 + i] = (int)(tacc[i] / 1000ull); }
}
__global__ void lsc_post_relabel_kernel(const int* __restrict__ reg, const int* __restrict__ cur, const int* __restrict__ newid,
                                        int* __restrict__ lab, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) lab[i] = newid[lsc_cur(cur, reg[i])];
}

struct MaxOp { __device__ __forceinline__ int operator()(int a, int b) const { return a > b ? a : b; } };

// ---- engine ----------------------------------------------------------------------------------------------------
struct LscEngine {
    int device = 0;
    cudaStream_t st = nullptr;
    LscDev L{};
    int S = 0;
    float ratio = 0.f;
    int numlabels = 0;
    int merge_order = 0;                    // post-pass visiting order: 0 = the authors' raster order (default, what upstream computes), 1 = hashed (opt-in)
    unsigned char* d_img = nullptr;
    int* d_labels = nullptr;
    unsigned char* d_mask = nullptr;
    float* d_tab = nullptr;                 // all float tables in one allocation
    unsigned long long* d_hist = nullptr;
    void* d_scan_tmp = nullptr;
    size_t scan_tmp_bytes = 0;
    int* d_flag = nullptr;
    int* d_tile_overflow = nullptr;         // tiles of the tiled assignment whose candidate list overflowed (diagnostic)
    bool tiled = false;                     // 3 channels and steps >= LSC_TILE_MIN_STEP: lsc_tile_kernel
    int* h_pinned = nullptr;
    long long launches = 0;
    PostWorkspace post;
    std::vector<void*> allocs;

    template <typename T> int dalloc(T** p, size_t n)
    {
        SPX_CHECK(pool_alloc((void**)p, n * sizeof(T), st));
        allocs.push_back((void*)*p);
        return SPX_OK;
    }
    ~LscEngine()
    {
        cudaSetDevice(device);
        if (st) cudaStreamSynchronize(st);
        for (void* p : allocs) pool_free(p, st);
        if (d_scan_tmp) pool_free(d_scan_tmp, st);
        post.release();
        pinned_scratch_free(h_pinned);
        if (st) cudaStreamDestroy(st);
    }
    int scan_tmp(size_t bytes)
    {
        if (bytes <= scan_tmp_bytes) return SPX_OK;
        if (d_scan_tmp) pool_free(d_scan_tmp, st);
        d_scan_tmp = nullptr; scan_tmp_bytes = 0;
        SPX_CHECK(pool_alloc(&d_scan_tmp, bytes, st));
        scan_tmp_bytes = bytes;
        return SPX_OK;
    }
    int exclusive_sum(const int* in, int* out, int n)
    {
        size_t bytes = 0;
        SPX_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, st));
        SPX_CHECK(scan_tmp(bytes));
        SPX_CUDA(cub::DeviceScan::ExclusiveSum(d_scan_tmp, bytes, in, out, n, st));
        launches += 2;
        return SPX_OK;
    }
    int exclusive_max(const int* in, int* out, int n)
    {
        size_t bytes = 0;
        SPX_CUDA(cub::DeviceScan::ExclusiveScan(nullptr, bytes, in, out, MaxOp(), -1, n, st));
        SPX_CHECK(scan_tmp(bytes));
        SPX_CUDA(cub::DeviceScan::ExclusiveScan(d_scan_tmp, bytes, in, out, MaxOp(), -1, n, st));
        launches += 2;
        return SPX_OK;
    }

    int create(const uint8_t* image, size_t stride, int w, int h, int C, int S_, float ratio_, int cvt_mode = -1)
    {
        SPX_REQUIRE(image && w > 0 && h > 0 && h <= SPX_MAX_HEIGHT && (long long)w * h < (1ll << 30) && stride >= (size_t)w * C, SPX_ERR_BADARG,
                    "createSuperpixelLSC: bad image");
        SPX_REQUIRE(C >= 1 && C <= LSC_MAXC, SPX_ERR_BADARG, "createSuperpixelLSC: 1..4 channels supported, got %d", C);
        SPX_REQUIRE(S_ >= 1, SPX_ERR_BADARG, "createSuperpixelLSC: region_size must be >= 1");
        S = S_; ratio = ratio_;
        const int K0 = (int)((float)(w * h) / (float)(S * S));
        const int ColNum = (int)std::sqrt((double)K0 * w / h);
        SPX_REQUIRE(ColNum >= 1 && K0 / ColNum >= 1, SPX_ERR_BADARG, "createSuperpixelLSC: region_size %d leaves no seed in a %dx%d image", S, w, h);
        const int RowNum = K0 / ColNum;
        L.w = w; L.h = h; L.C = C; L.D = 2 * C + 4; L.K = ColNum * RowNum;
        L.stepx = w / ColNum; L.stepy = h / RowNum;
        L.nbx = cdiv(w, L.stepx); L.nby = cdiv(h, L.stepy);
        L.ipitch = round_up((size_t)w * C, 16);
        numlabels = L.K;
        const size_t N = (size_t)w * h;
        SPX_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        SPX_CHECK(dalloc(&d_img, L.ipitch * h));
        SPX_CHECK(dalloc(&d_labels, N));
        SPX_CHECK(dalloc(&d_mask, N));
        const size_t ntab = (size_t)3 * C * 256 + 3 * (size_t)w + 3 * (size_t)h;
        SPX_CHECK(dalloc(&d_tab, ntab));
        SPX_CHECK(dalloc(&d_hist, (size_t)C * 256));
        SPX_CHECK(dalloc(&L.sx, (size_t)L.K)); SPX_CHECK(dalloc(&L.sy, (size_t)L.K));
        SPX_CHECK(dalloc(&L.cen, (size_t)L.K * L.D));
        SPX_CHECK(dalloc(&L.sums, (size_t)L.K * (L.D + 4)));
        SPX_CHECK(dalloc(&L.head, (size_t)L.nbx * L.nby)); SPX_CHECK(dalloc(&L.next, (size_t)L.K));
        SPX_CHECK(dalloc(&d_flag, 8));
        SPX_CHECK(dalloc(&d_tile_overflow, 1));
        SPX_CUDA(cudaMemsetAsync(d_tile_overflow, 0, sizeof(int), st));
        static const int min_step = getenv("SPX_LSC_TILE_MIN_STEP") ? atoi(getenv("SPX_LSC_TILE_MIN_STEP")) : LSC_TILE_MIN_STEP;   // (experiments)
        tiled = C == 3 && L.stepx >= min_step && L.stepy >= min_step && !getenv("SPX_LSC_NO_TILE") &&
                20.0f * ratio_ < 32768.0f;          // (lsc_sadd64's split assumes |feature| * 2^32 * 4 < 2^55)
        L.tiles_x = cdiv(w, 32); L.tiles_y = cdiv(h, LSC_TH);
        L.ctab = nullptr; L.tl_count = nullptr; L.tl_k = nullptr;
        float4* d_ctab = nullptr;
        if (tiled) {
            SPX_CHECK(dalloc(&d_ctab, (size_t)3 * 256));
            SPX_CHECK(dalloc(&L.tl_count, (size_t)L.tiles_x * L.tiles_y));
            SPX_CHECK(dalloc(&L.tl_k, (size_t)L.tiles_x * L.tiles_y * LSC_CAP));
            SPX_CUDA(cudaMemsetAsync(L.tl_count, 0, (size_t)L.tiles_x * L.tiles_y * sizeof(int), st));
            L.ctab = d_ctab;
        }
        SPX_CHECK(pinned_scratch_alloc((void**)&h_pinned, 32 * sizeof(int)));
        L.img = d_img;
        SPX_CHECK(upload_2d(d_img, L.ipitch, image, stride, (size_t)w * C, h, st));
        if (cvt_mode >= 0) {                  // cvtColor of mainwindow.cpp:2096-2097 on the uploaded image, in place
            SPX_REQUIRE(C == 3, SPX_ERR_BADARG, "createSuperpixelLSC: the fused colour conversion needs a 3-channel BGR image");
            SPX_CHECK(cvt8_dev(cvt_mode, d_img, L.ipitch, d_img, L.ipitch, w, h, st));
            launches++;
        }
        SPX_CUDA(cudaMemsetAsync(d_labels, 0, N * sizeof(int), st));
        SPX_CUDA(cudaMemsetAsync(L.sums, 0, (size_t)L.K * (L.D + 4) * sizeof(long long), st));
        SPX_CUDA(cudaMemsetAsync(d_hist, 0, (size_t)C * 256 * sizeof(unsigned long long), st));
        SPX_CUDA(cudaMemsetAsync(L.head, 0xFF, (size_t)L.nbx * L.nby * sizeof(int), st));
        lsc_hist_kernel<<<dim3(min(cdiv(w, 256), 8), h), 256, 0, st>>>(d_img, L.ipitch, w, h, C, d_hist);
        launches++;
        std::vector<unsigned long long> hist((size_t)C * 256);
        SPX_CUDA(cudaMemcpyAsync(hist.data(), d_hist, hist.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        SPX_CUDA(cudaStreamSynchronize(st));

        // ---- tables (double trigonometry, rounded once) and seeds: O(256 C + W + H + K) host work
        std::vector<float> tab(ntab);
        float* colcos = tab.data(); float* colsin = colcos + C * 256; float* wcol = colsin + C * 256;
        float* xcos = wcol + C * 256; float* xsin = xcos + w; float* wx = xsin + w;
        float* ycos = wx + w; float* ysin = ycos + h; float* wy = ysin + h;
        const double PI = 3.1415926, color = 20.0;
        const float dist_coeff = 20.0f * ratio;
        const double dist = (double)dist_coeff;
        for (int c = 0; c < C; c++) {
            const double scale = c == 0 ? 1.0 : 2.55;
            for (int v = 0; v < 256; v++) {
                const double th = ((double)v / 255.0) * PI / 2.0;
                colcos[c * 256 + v] = (float)(color * std::cos(th) * scale);
                colsin[c * 256 + v] = (float)(color * std::sin(th) * scale);
            }
        }
        for (int x = 0; x < w; x++) {
            const double th = ((double)x / (double)L.stepx) * PI / 2.0;
            xcos[x] = (float)(dist * std::cos(th)); xsin[x] = (float)(dist * std::sin(th));
        }
        for (int y = 0; y < h; y++) {
            const double th = ((double)y / (double)L.stepy) * PI / 2.0;
            ycos[y] = (float)(dist * std::cos(th)); ysin[y] = (float)(dist * std::sin(th));
        }
        for (int c = 0; c < C; c++) {
            double a = 0.0, b = 0.0;
            for (int v = 0; v < 256; v++) { a += (double)hist[c * 256 + v] * (double)colcos[c * 256 + v]; b += (double)hist[c * 256 + v] * (double)colsin[c * 256 + v]; }
            const float sc = (float)(a / (double)N), ss = (float)(b / (double)N);
            for (int v = 0; v < 256; v++) wcol[c * 256 + v] = colcos[c * 256 + v] * sc + colsin[c * 256 + v] * ss;
        }
        {
            double a = 0.0, b = 0.0;
            for (int x = 0; x < w; x++) { a += (double)xcos[x]; b += (double)xsin[x]; }
            float sc = (float)(a * (double)h / (double)N), ss = (float)(b * (double)h / (double)N);
            for (int x = 0; x < w; x++) wx[x] = xcos[x] * sc + xsin[x] * ss;
            a = 0.0; b = 0.0;
            for (int y = 0; y < h; y++) { a += (double)ycos[y]; b += (double)ysin[y]; }
            sc = (float)(a * (double)w / (double)N); ss = (float)(b * (double)w / (double)N);
            for (int y = 0; y < h; y++) wy[y] = ycos[y] * sc + ysin[y] * ss;
        }
        std::vector<int> seeds((size_t)2 * L.K);
        {
            const int row_remain = h - L.stepy * RowNum, col_remain = w - L.stepx * ColNum;
            int t1 = 1, n = 0;
            for (int i = 0; i < RowNum; i++) {
                int t2 = 1;
                int cy = (int)(i * L.stepy + 0.5 * L.stepy + t1);
                if (cy >= h - 1) cy = h - 1;
                if (t1 < row_remain) t1++;
                for (int j = 0; j < ColNum; j++) {
                    int cx = (int)(j * L.stepx + 0.5 * L.stepx + t2);
                    if (t2 < col_remain) t2++;
                    if (cx >= w - 1) cx = w - 1;
                    seeds[n] = cx; seeds[L.K + n] = cy; n++;
                }
            }
        }
        std::vector<float4> ctab((size_t)3 * 256);
        if (tiled) {
            for (int i = 0; i < 3 * 256; i++) ctab[i] = make_float4(colcos[i], colsin[i], wcol[i], 0.f);
            SPX_CUDA(cudaMemcpyAsync(d_ctab, ctab.data(), ctab.size() * sizeof(float4), cudaMemcpyHostToDevice, st));
        }
        SPX_CUDA(cudaMemcpyAsync(d_tab, tab.data(), ntab * sizeof(float), cudaMemcpyHostToDevice, st));
        SPX_CUDA(cudaMemcpyAsync(L.sx, seeds.data(), (size_t)L.K * sizeof(int), cudaMemcpyHostToDevice, st));
        SPX_CUDA(cudaMemcpyAsync(L.sy, seeds.data() + L.K, (size_t)L.K * sizeof(int), cudaMemcpyHostToDevice, st));
        L.colcos = d_tab; L.colsin = L.colcos + C * 256; L.wcol = L.colsin + C * 256;
        L.xcos = L.wcol + C * 256; L.xsin = L.xcos + w; L.wx = L.xsin + w;
        L.ycos = L.wx + w; L.ysin = L.ycos + h; L.wy = L.ysin + h;
        const int nb = cdiv(L.K * 32, 128);
        switch (C) {
            case 1: lsc_init_centres_kernel<1><<<nb, 128, 0, st>>>(L); break;
            case 2: lsc_init_centres_kernel<2><<<nb, 128, 0, st>>>(L); break;
            case 3: lsc_init_centres_kernel<3><<<nb, 128, 0, st>>>(L); break;
            default: lsc_init_centres_kernel<4><<<nb, 128, 0, st>>>(L); break;
        }
        lsc_bin_kernel<<<cdiv(L.K, 256), 256, 0, st>>>(L);
        launches += 2;
        SPX_CUDA(cudaGetLastError());
        SPX_CUDA(cudaStreamSynchronize(st));       // the host vectors above go out of scope
        return SPX_OK;
    }

    int iterate(int n)
    {
        SPX_REQUIRE(n >= 0, SPX_ERR_BADARG, "iterate: num_iterations must be >= 0");
        const dim3 grid(cdiv(L.w, 32), cdiv(L.h, 8));
        for (int it = 0; it < n; it++) {
            if (tiled) {
                const dim3 tg(L.tiles_x, L.tiles_y);
                lsc_tile_kernel<<<tg, 128, 0, st>>>(L, d_labels, d_tile_overflow);
                SPX_CUDA(cudaMemsetAsync(L.tl_count, 0, (size_t)L.tiles_x * L.tiles_y * sizeof(int), st));
            }
            else switch (L.C) {
                case 1: lsc_assign_kernel<1><<<grid, 256, 0, st>>>(L, d_labels, 1); break;
                case 2: lsc_assign_kernel<2><<<grid, 256, 0, st>>>(L, d_labels, 1); break;
                case 3: lsc_assign_kernel<3><<<grid, 256, 0, st>>>(L, d_labels, 1); break;
                default: lsc_assign_kernel<4><<<grid, 256, 0, st>>>(L, d_labels, 1); break;
            }
            SPX_CUDA(cudaMemsetAsync(L.head, 0xFF, (size_t)L.nbx * L.nby * sizeof(int), st));
            lsc_finalize_kernel<<<cdiv(L.K, 256), 256, 0, st>>>(L);
            launches += 2;
        }
        SPX_CUDA(cudaGetLastError());
        return SPX_OK;
    }

    int enforce(int min_element_size)
    {
        SPX_REQUIRE(min_element_size >= 0 && min_element_size <= 100, SPX_ERR_ASSERT,
                    "enforceLabelConnectivity: min_element_size >= 0 && min_element_size <= 100");
        if (min_element_size == 0) return SPX_OK;
        const int w = L.w, h = L.h, n = w * h, D = L.D;
        const int threshold = n / (numlabels * 4);
        const int g1 = cdiv(n, 256);
        const dim3 g2(cdiv(w, 256), h), g2m(cdiv(w, 256), max(h - 1, 1));
        int *P, *csize, *rootidx, *flag, *A, *B, *Cc;
        std::vector<void*> tmp;
        auto talloc = [&](int** p, size_t cnt) -> int {
            SPX_CHECK(pool_alloc((void**)p, (cnt > 0 ? cnt : 1) * sizeof(int), st));
            tmp.push_back(*p);
            return SPX_OK;
        };
        auto tfree = [&]() { for (void* p : tmp) pool_free(p, st); tmp.clear(); };
        int rc = SPX_OK;
#define LSC_TRY(expr) do { rc = (expr); if (rc != SPX_OK) { tfree(); return rc; } } while (0)
#define LSC_TRYCUDA(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { set_error("%s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); tfree(); return SPX_ERR_GPU; } } while (0)
        LSC_TRY(talloc(&P, n)); LSC_TRY(talloc(&csize, n)); LSC_TRY(talloc(&rootidx, n)); LSC_TRY(talloc(&flag, n));
        LSC_TRY(talloc(&A, n)); LSC_TRY(talloc(&B, n)); LSC_TRY(talloc(&Cc, n));

        // ---- pre-pass: 8-connected components, `adj` rule
        lsc_ccl_init_kernel<<<g2, 256, 0, st>>>(d_labels, w, h, P);
        if (h > 1) lsc_ccl_merge_kernel<8><<<g2m, 256, 0, st>>>(d_labels, w, h, P);
        lsc_ccl_flatten_kernel<<<g1, 256, 0, st>>>(P, csize, rootidx, flag, n);
        lsc_ccl_size_kernel<<<g1, 256, 0, st>>>(P, csize, n);
        launches += 4;
        int* prevroot = Cc;
        LSC_TRY(exclusive_max(rootidx, prevroot, n));
        int* FL = A; int* ADJ = B;
        lsc_pre_init_kernel<<<g1, 256, 0, st>>>(P, d_labels, csize, FL, ADJ, n, min_element_size);
        launches++;
        // the relaxation only concerns component roots: compact them once (rootidx is free after the scan above)
        int* roots = rootidx;
        int* d_nroots = d_flag + 1;
        {
            size_t bytes = 0;
            cub::CountingInputIterator<int> idx(0);
            LSC_TRYCUDA(cub::DeviceSelect::Flagged(nullptr, bytes, idx, flag, roots, d_nroots, n, st));
            LSC_TRY(scan_tmp(bytes));
            LSC_TRYCUDA(cub::DeviceSelect::Flagged(d_scan_tmp, bytes, idx, flag, roots, d_nroots, n, st));
            launches += 2;
        }
        for (int round = 0; round < 1000000; round++) {
            LSC_TRYCUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), st));
            for (int k = 0; k < 4; k++)
                lsc_pre_relax_list_kernel<<<148 * 8, 256, 0, st>>>(roots, d_nroots, P, d_labels, csize, prevroot, FL, ADJ, w, min_element_size, d_flag);
            launches += 4;
            LSC_TRYCUDA(cudaMemcpyAsync(h_pinned, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
            LSC_TRYCUDA(cudaStreamSynchronize(st));
            if (!h_pinned[0]) break;
        }
        lsc_pre_relabel_kernel<<<g1, 256, 0, st>>>(P, FL, d_labels, n);
        launches++;
        if (getenv("SPX_LSC_DEBUG")) { cudaStreamSynchronize(st); fprintf(stderr, "[lsc] pre-pass done, launches so far %lld\n", launches); }

        // ---- post-pass: 4-connected regions, statistics, pixel lists
        lsc_ccl_init_kernel<<<g2, 256, 0, st>>>(d_labels, w, h, P);
        if (h > 1) lsc_ccl_merge_kernel<4><<<g2m, 256, 0, st>>>(d_labels, w, h, P);
        lsc_ccl_flatten_kernel<<<g1, 256, 0, st>>>(P, csize, rootidx, flag, n);
        lsc_ccl_size_kernel<<<g1, 256, 0, st>>>(P, csize, n);
        launches += 4;
        int* rid = A;
        LSC_TRY(exclusive_sum(flag, rid, n));
        LSC_TRYCUDA(cudaMemcpyAsync(h_pinned, rid + n - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        LSC_TRYCUDA(cudaMemcpyAsync(h_pinned + 1, flag + n - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        LSC_TRYCUDA(cudaStreamSynchronize(st));
        const int R = h_pinned[0] + h_pinned[1];
        int *reg = B, *size, *alive, *smallflag, *done, *cur, *newid;
        uint4* rr;
        unsigned long long* best;
        float *cen, *Wr;
        long long* rsum;
        LSC_TRY(talloc(&size, R + 1)); LSC_TRY(talloc(&alive, R + 1)); LSC_TRY(talloc(&smallflag, R)); LSC_TRY(talloc(&done, R));
        LSC_TRY(talloc(&cur, R)); LSC_TRY(talloc((int**)&rr, (size_t)R * 4));
        LSC_TRY(talloc(&newid, R + 1)); LSC_TRY(talloc((int**)&best, (size_t)R * 2));
        LSC_TRY(talloc((int**)&cen, (size_t)R * D)); LSC_TRY(talloc((int**)&Wr, R));
        LSC_TRY(talloc((int**)&rsum, (size_t)R * (D + 1) * 2));
        LSC_TRYCUDA(cudaMemsetAsync(size, 0, (size_t)(R + 1) * sizeof(int), st));
        LSC_TRYCUDA(cudaMemsetAsync(rsum, 0, (size_t)R * (D + 1) * sizeof(long long), st));
        LSC_TRYCUDA(cudaMemsetAsync(alive, 0, (size_t)(R + 1) * sizeof(int), st));
        lsc_post_region_kernel<<<g1, 256, 0, st>>>(P, rid, csize, reg, size, n);
        const dim3 gp(cdiv(w, 32), cdiv(h, 8));
        switch (L.C) {
            case 1: lsc_post_sums_kernel<1><<<gp, 256, 0, st>>>(L, reg, rsum); break;
            case 2: lsc_post_sums_kernel<2><<<gp, 256, 0, st>>>(L, reg, rsum); break;
            case 3: lsc_post_sums_kernel<3><<<gp, 256, 0, st>>>(L, reg, rsum); break;
            default: lsc_post_sums_kernel<4><<<gp, 256, 0, st>>>(L, reg, rsum); break;
        }
        const int gr = cdiv(R, 256);
        lsc_post_stats_kernel<<<gr, 256, 0, st>>>(rsum, size, D, R, threshold, cen, Wr, alive, smallflag, done, cur);
        launches += 3;
        // ---- region adjacency edges, sorted and de-duplicated once
        // (key buffers sized for the worst case: every horizontal and vertical pixel pair)
        const unsigned long long ecap = (unsigned long long)2 * n;
        unsigned long long *kbuf0, *kbuf1, *ecount;
        LSC_TRY(talloc((int**)&kbuf0, (size_t)ecap * 2)); LSC_TRY(talloc((int**)&kbuf1, (size_t)ecap * 2)); LSC_TRY(talloc((int**)&ecount, 4));
        LSC_TRYCUDA(cudaMemsetAsync(ecount, 0, 2 * sizeof(unsigned long long), st));
        lsc_edge_count_kernel<<<g2, 256, 0, st>>>(reg, w, h, kbuf0, ecount, ecap);
        launches++;
        unsigned long long h_ec = 0;
        LSC_TRYCUDA(cudaMemcpyAsync(&h_ec, ecount, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        LSC_TRYCUDA(cudaStreamSynchronize(st));
        int E = 0;
        unsigned long long* edges = kbuf1;
        if (h_ec > 0) {
            const int nraw = (int)h_ec;
            size_t b1 = 0, b2 = 0;
            LSC_TRYCUDA(cub::DeviceRadixSort::SortKeys(nullptr, b1, kbuf0, kbuf1, nraw, 0, 64, st));
            int* d_nuniq = reinterpret_cast<int*>(ecount + 1);
            LSC_TRYCUDA(cub::DeviceSelect::Unique(nullptr, b2, kbuf1, kbuf0, d_nuniq, nraw, st));
            LSC_TRY(scan_tmp(std::max(b1, b2)));
            LSC_TRYCUDA(cub::DeviceRadixSort::SortKeys(d_scan_tmp, b1, kbuf0, kbuf1, nraw, 0, 64, st));
            LSC_TRYCUDA(cub::DeviceSelect::Unique(d_scan_tmp, b2, kbuf1, kbuf0, d_nuniq, nraw, st));
            launches += 4;
            LSC_TRYCUDA(cudaMemcpyAsync(h_pinned, d_nuniq, sizeof(int), cudaMemcpyDeviceToHost, st));
            LSC_TRYCUDA(cudaStreamSynchronize(st));
            E = h_pinned[0];
            edges = kbuf0;
        }
        // ---- merge rounds (see the comment above lsc_rounds_kernel)
        int rounds = 0;
        {
            static int s_sms = 0;
            if (!s_sms) {
                int dev = 0;
                LSC_TRYCUDA(cudaGetDevice(&dev));
                LSC_TRYCUDA(cudaDeviceGetAttribute(&s_sms, cudaDevAttrMultiProcessorCount, dev));
            }
            int *plist0, *plist1, *rcnt;
            LSC_TRY(talloc(&plist0, (size_t)R)); LSC_TRY(talloc(&plist1, (size_t)R)); LSC_TRY(talloc(&rcnt, 32));
            LSC_TRYCUDA(cudaMemsetAsync(rcnt, 0, 32 * sizeof(int), st));
            LscRounds ra_{};
            ra_.edges[0] = edges; ra_.edges[1] = edges == kbuf0 ? kbuf1 : kbuf0; ra_.E = E;
            ra_.plist[0] = plist0; ra_.plist[1] = plist1;
            ra_.smallflag = smallflag; ra_.done = done; ra_.size = size; ra_.threshold = threshold; ra_.R = R;
            ra_.order = merge_order; ra_.rr = rr; ra_.best = best;
            ra_.cen = cen; ra_.D = D; ra_.Wr = Wr; ra_.alive = alive; ra_.cur = cur;
            ra_.cnt = rcnt; ra_.max_rounds = R + 2;
            int blocks = std::min(std::max(cdiv(std::max(R, E), 1024), 1), s_sms);   // one CTA per SM at most: all co-resident
            if (merge_order == 0 && !getenv("SPX_LSC_ROUNDS")) {
                // the authors' order: fixed-point evaluation of the sequential rule (see lsc_fixpoint_kernel)
                LscFix fx{};
                fx.edges = edges; fx.E = E; fx.R = R; fx.D = D; fx.threshold = threshold;
                fx.cen = cen; fx.Wr = Wr; fx.size = size; fx.smallflag = smallflag;
                fx.tg = done; fx.best = best; fx.cur = cur; fx.alive = alive; fx.cnt = rcnt; fx.max_rounds = R + 2;
                fx.tmp = plist0; fx.ch = plist1;
                LSC_TRY(talloc(&fx.off, (size_t)R + 1)); LSC_TRY(talloc(&fx.nch, (size_t)R)); LSC_TRY(talloc(&fx.pre, (size_t)R));
                LSC_TRY(talloc(&fx.prog, (size_t)R)); LSC_TRY(talloc((int**)&fx.S, (size_t)R * (D + 2))); LSC_TRY(talloc(&fx.bsum, (size_t)s_sms + 1));
                LSC_TRY(talloc(&fx.inc_off, (size_t)R + 1)); LSC_TRY(talloc(&fx.inc, (size_t)2 * E + 1)); LSC_TRY(talloc(&fx.eff, (size_t)R));
                LSC_TRY(talloc(&fx.need, (size_t)R)); LSC_TRY(talloc(&fx.dq, (size_t)R)); LSC_TRY(talloc(&fx.dtau, (size_t)R));
                void* args[] = {(void*)&fx};
                LSC_TRYCUDA(cudaLaunchCooperativeKernel((const void*)lsc_fixpoint_kernel, dim3(blocks), dim3(1024), args, 0, st));
            } else {
                void* args[] = {(void*)&ra_};
                LSC_TRYCUDA(cudaLaunchCooperativeKernel((const void*)lsc_rounds_kernel, dim3(blocks), dim3(1024), args, 0, st));
            }
            launches++;
            LSC_TRYCUDA(cudaMemcpyAsync(h_pinned, rcnt + 9, 16 * sizeof(int), cudaMemcpyDeviceToHost, st));
            LSC_TRYCUDA(cudaStreamSynchronize(st));
            rounds = h_pinned[0];
            if (getenv("SPX_LSC_DEBUG") && merge_order == 0)
                fprintf(stderr, "[lsc] fixed point: %d rounds (%d full), %d trajectory passes; us per phase: setup %d count+push %d scan %d fill+dirty %d rank+mark %d trajectories %d candidates %d decisions %d\n",
                        h_pinned[0], h_pinned[2], h_pinned[1], h_pinned[14], h_pinned[7], h_pinned[8], h_pinned[9], h_pinned[10], h_pinned[11], h_pinned[12], h_pinned[13]);
        }
        if (getenv("SPX_LSC_DEBUG")) fprintf(stderr, "[lsc] post-pass: %d regions, %d adjacency edges (%llu raw), threshold %d, %d merge rounds, launches %lld\n", R, E, h_ec, threshold, rounds, launches);
        LSC_TRY(exclusive_sum(alive, newid, R + 1));
        lsc_post_relabel_kernel<<<g1, 256, 0, st>>>(reg, cur, newid, d_labels, n);
        launches++;
        LSC_TRYCUDA(cudaMemcpyAsync(h_pinned, newid + R, sizeof(int), cudaMemcpyDeviceToHost, st));
        LSC_TRYCUDA(cudaStreamSynchronize(st));
        LSC_TRYCUDA(cudaGetLastError());
        numlabels = h_pinned[0];
        tfree();
#undef LSC_TRY
#undef LSC_TRYCUDA
        return SPX_OK;
    }
};

}  // namespace spx

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
using spx::LscEngine;
struct spx_lsc { LscEngine* e; };

#define LSC_ENTER(h)                                                                      \
    SPX_REQUIRE((h) && (h)->e, SPX_ERR_BADARG, "NULL SuperpixelLSC handle");              \
    LscEngine* e = (h)->e;                                                                \
    SPX_CUDA(cudaSetDevice(e->device))

extern "C" int spx_lsc_create(const uint8_t* image, size_t stride, int w, int h, int C, int S, float ratio, int device, spx_lsc_t** out) try
{
    using namespace spx;
    SPX_REQUIRE(out, SPX_ERR_BADARG, "createSuperpixelLSC: out is NULL");
    *out = nullptr;
    SPX_CHECK(require_device(device));
    LscEngine* e = new LscEngine();
    e->device = device;
    int rc = e->create(image, stride, w, h, C, S, ratio);
    if (rc != SPX_OK) { delete e; return rc; }
    *out = new spx_lsc{e};
    return SPX_OK;
} SPX_API_CATCH
extern "C" int spx_lsc_create_bgr(const uint8_t* bgr, size_t stride, int w, int h, int colour_code, int S, float ratio, int device, spx_lsc_t** out) try
{
    using namespace spx;
    SPX_REQUIRE(out, SPX_ERR_BADARG, "createSuperpixelLSC: out is NULL");
    *out = nullptr;
    SPX_REQUIRE(colour_code == 44 || colour_code == 40, SPX_ERR_BADARG, "createSuperpixelLSC: colour code must be COLOR_BGR2Lab (44) or COLOR_BGR2HSV (40)");
    SPX_CHECK(require_device(device));
    LscEngine* e = new LscEngine();
    e->device = device;
    int rc = e->create(bgr, stride, w, h, 3, S, ratio, colour_code == 44 ? 0 : 1);
    if (rc != SPX_OK) { delete e; return rc; }
    *out = new spx_lsc{e};
    return SPX_OK;
} SPX_API_CATCH
extern "C" int spx_lsc_iterate(spx_lsc_t* h, int n) try
{
    LSC_ENTER(h);
    return e->iterate(n);
} SPX_API_CATCH
extern "C" int spx_lsc_enforce_label_connectivity(spx_lsc_t* h, int min_element_size) try
{
    LSC_ENTER(h);
    return e->enforce(min_element_size);
} SPX_API_CATCH
extern "C" int spx_lsc_get_number_of_superpixels(spx_lsc_t* h, int* out) try
{
    LSC_ENTER(h);
    SPX_REQUIRE(out, SPX_ERR_BADARG, "getNumberOfSuperpixels: out is NULL");
    SPX_CUDA(cudaStreamSynchronize(e->st));
    *out = e->numlabels;
    return SPX_OK;
} SPX_API_CATCH
extern "C" int spx_lsc_get_labels(spx_lsc_t* h, int32_t* dst, size_t dst_stride) try
{
    LSC_ENTER(h);
    SPX_REQUIRE(dst && dst_stride >= (size_t)e->L.w * 4, SPX_ERR_BADARG, "getLabels: bad destination");
    SPX_CHECK(spx::download_2d(dst, dst_stride, e->d_labels, (size_t)e->L.w * 4, (size_t)e->L.w * 4, e->L.h, e->st));
    SPX_CUDA(cudaStreamSynchronize(e->st));
    return SPX_OK;
} SPX_API_CATCH
extern "C" int spx_lsc_get_label_contour_mask(spx_lsc_t* h, uint8_t* dst, size_t dst_stride, int thick_line) try
{
    LSC_ENTER(h);
    SPX_REQUIRE(dst && dst_stride >= (size_t)e->L.w, SPX_ERR_BADARG, "getLabelContourMask: bad destination");
    SPX_CHECK(e->post.alloc(e->L.w, e->L.h, e->st));
    long long l0 = e->post.launches;
    SPX_CHECK(spx::contour_mask_dev(e->post, e->d_labels, e->L.w, thick_line, 0, e->d_mask, e->L.w, e->st));
    e->launches += e->post.launches - l0;
    SPX_CHECK(spx::download_2d(dst, dst_stride, e->d_mask, e->L.w, e->L.w, e->L.h, e->st));
    SPX_CUDA(cudaStreamSynchronize(e->st));
    return SPX_OK;
} SPX_API_CATCH
extern "C" int spx_lsc_set_labels(spx_lsc_t* h, const int32_t* src, size_t src_stride, int numlabels) try
{
    LSC_ENTER(h);
    SPX_REQUIRE(src && src_stride >= (size_t)e->L.w * 4, SPX_ERR_BADARG, "setLabels: bad source");
    SPX_CUDA(cudaMemcpy2DAsync(e->d_labels, (size_t)e->L.w * 4, src, src_stride, (size_t)e->L.w * 4, e->L.h, cudaMemcpyHostToDevice, e->st));
    SPX_CUDA(cudaStreamSynchronize(e->st));
    if (numlabels > 0) e->numlabels = numlabels;
    return SPX_OK;
} SPX_API_CATCH
extern "C" int spx_lsc_get_centres(spx_lsc_t* h, float* dst, int rows) try
{
    LSC_ENTER(h);
    const int K = e->L.K, D = e->L.D;
    SPX_REQUIRE(dst && rows >= K, SPX_ERR_BADARG, "getCentres: need room for %d rows", K);
    std::vector<float> cen((size_t)K * D);
    std::vector<int> sx(K), sy(K);
    SPX_CUDA(cudaStreamSynchronize(e->st));
    SPX_CUDA(cudaMemcpy(cen.data(), e->L.cen, cen.size() * sizeof(float), cudaMemcpyDeviceToHost));
    SPX_CUDA(cudaMemcpy(sx.data(), e->L.sx, (size_t)K * sizeof(int), cudaMemcpyDeviceToHost));
    SPX_CUDA(cudaMemcpy(sy.data(), e->L.sy, (size_t)K * sizeof(int), cudaMemcpyDeviceToHost));
    for (int k = 0; k < K; k++) {
        for (int i = 0; i < D; i++) dst[(size_t)k * (D + 2) + i] = cen[(size_t)k * D + i];
        dst[(size_t)k * (D + 2) + D] = (float)sx[k];
        dst[(size_t)k * (D + 2) + D + 1] = (float)sy[k];
    }
    return SPX_OK;
} SPX_API_CATCH
extern "C" int spx_lsc_set_merge_order(spx_lsc_t* h, int order) try
{
    SPX_REQUIRE(h && h->e && (order == 0 || order == 1), SPX_ERR_BADARG, "set_merge_order: bad argument");
    h->e->merge_order = order;
    return SPX_OK;
} SPX_API_CATCH
extern "C" int spx_lsc_launch_count(spx_lsc_t* h, long long* out) try
{
    SPX_REQUIRE(h && h->e && out, SPX_ERR_BADARG, "launch_count: bad argument");
    *out = h->e->launches;
    return SPX_OK;
} SPX_API_CATCH
extern "C" int spx_lsc_destroy(spx_lsc_t* h) try
{
    if (!h) return SPX_OK;
    delete h->e;
    delete h;
    return SPX_OK;
} SPX_API_CATCH
