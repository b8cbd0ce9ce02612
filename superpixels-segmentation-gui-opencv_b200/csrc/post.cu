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
#include <algorithm>

namespace spx {

// ------------------------------------------------------------------------------------------------
// workspace
// ------------------------------------------------------------------------------------------------
int PostWorkspace::alloc(int w_, int h_, cudaStream_t st)
{
    if (w_ == w && h_ == h && parent) return SPX_OK;
    release();
    st_alloc = st;
    w = w_; h = h_;
    size_t N = (size_t)w * h;
    SPX_CHECK(pool_alloc((void**)&parent, N * sizeof(int), st));
    SPX_CHECK(pool_alloc((void**)&csize, N * sizeof(int), st));
    SPX_CHECK(pool_alloc((void**)&newid, N * sizeof(int), st));
    SPX_CHECK(pool_alloc((void**)&link, N * sizeof(int), st));
    SPX_CHECK(pool_alloc((void**)&blocksum, (N / 1024 + 2) * sizeof(int), st));
    SPX_CHECK(pool_alloc((void**)&flag, 16 * sizeof(int), st));
    SPX_CHECK(pool_alloc((void**)&diff, N + 64, st));       // (+64: the bitmap and prefix words of a tiny image)
    SPX_CHECK(pinned_scratch_alloc((void**)&h_flag, 16 * sizeof(int)));
    return SPX_OK;
}
void PostWorkspace::release()
{
    pool_free(parent, st_alloc); pool_free(csize, st_alloc); pool_free(newid, st_alloc); pool_free(link, st_alloc);
    pool_free(blocksum, st_alloc); pool_free(flag, st_alloc); pool_free(diff, st_alloc);
    pinned_scratch_free(h_flag);
    parent = csize = newid = link = blocksum = flag = nullptr;
    diff = nullptr; h_flag = nullptr;
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

// Block-local union-find first (shared memory), global unions only along tile borders, and after that everything works on the
// (few) LOCAL ROOTS instead of on pixels:
//   ccl_local_kernel   32x32 tile per CTA of 256 threads (a warp = one tile row at a time, 4 rows per thread): parent = head of the
//                      horizontal run (one ballot), vertical unions inside the tile with shared-memory atomicMin (skipped when the
//                      left neighbour already carries the same union), pointer jumping until every pixel points at its local root
//                      (= the tile's smallest raster index of the component, so global roots stay "smallest raster index"),
//                      pixel counts per local root (warp-aggregated), P[i] = global index of the local root; the local roots go
//                      to a compact list with their counts.
//   ccl_border_kernel  unions across tile borders (top row with the row above, left column with the column to the left).
//   ccl_roots_kernel   list: local root -> global root (flattened), its count added to the global root's size.
//   ccl_keep_kernel    list: global roots: kept (size > min) -> one bit in a raster-order bitmap; small -> its link.
//   bitmap scan        new id of a kept root = number of kept roots before it in raster order = prefix popcount (a 1 MB scan at 4K
//                      instead of two passes over the 33 MB parent array).
//   ccl_relabel_kernel pixel -> local root -> global root -> (links while small) -> prefix popcount.
constexpr int CT = 32;
__device__ __forceinline__ int sm_find(const int* P, int a)
{
    int p = P[a];
    while (p != a) { a = p; p = P[a]; }
    return a;
}
// The pixel counts come from the runs of equal local root along a tile row (the head of a run adds the run's length: one shared
// atomic per run, not per pixel); root pixels append themselves to the list, one global atomic per warp row that holds any.
__global__ void __launch_bounds__(256) ccl_local_kernel(const int* __restrict__ lab, int lp, int w, int h, int* __restrict__ P,
                                                        int* __restrict__ csize, int* __restrict__ roots, int* __restrict__ nroots)
{
    __shared__ int sl[CT][CT + 1];
    __shared__ int sp[CT * CT];
    __shared__ int sc[CT * CT];
    __shared__ int s_roots[CT * CT];
    __shared__ int s_n, s_base;
    if (threadIdx.x == 0) s_n = 0;
    // 8 warps; warp q owns the SLAB of tile rows 4q .. 4q+3 and sweeps it top-down, lane = column, everything in registers:
    // a run of equal labels that continues a run of the row above takes that run's root (one ballot + one shuffle), otherwise it
    // opens a root at its head; only a run that joins two different roots (U shapes, rare) touches the parent table with atomics.
    // Then the 7 slab borders are united (shared-memory union-find on slab roots: chains at most 8 deep), then one short find per
    // pixel.  (The form before this one united every row with the row above concurrently: parent chains as deep as a component
    // is tall, 223 thread-instructions per pixel, 75 us at 4K.)
    const int lx = threadIdx.x & 31, wq = threadIdx.x >> 5;
    const int x = blockIdx.x * CT + lx;
    int me[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int y = blockIdx.y * CT + 4 * wq + k;
        me[k] = (x < w && y < h) ? lab[(size_t)y * lp + x] : -1 - lx;    // outside: never equal to a neighbour
    }
    const unsigned le = 0xffffffffu >> (31 - lx);                        // lanes <= lx
    int above_root = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int ly = 4 * wq + k, t = ly * CT + lx;
        sl[ly][lx] = me[k];
        sc[t] = 0;
        const int prevl = __shfl_up_sync(0xffffffffu, me[k], 1);
        const unsigned hb = __ballot_sync(0xffffffffu, lx == 0 || prevl != me[k]);
        const int headpos = 31 - __clz(hb & le);
        const unsigned after = lx == 31 ? 0u : hb >> (lx + 1);           // heads to the right of this lane
        const int endpos = after ? lx + __ffs(after) : 32;               // the run is lanes [headpos, endpos)
        const unsigned runmask = (endpos == 32 ? 0xffffffffu : (1u << endpos) - 1u) & ~((1u << headpos) - 1u);
        const bool conn = k > 0 && me[k] == me[k > 0 ? k - 1 : 0];
        const unsigned cb = __ballot_sync(0xffffffffu, conn) & runmask;
        const int src = cb ? __ffs(cb) - 1 : lx;
        const int taken = __shfl_sync(0xffffffffu, above_root, src);
        const int root = cb ? taken : ly * CT + headpos;
        sp[t] = root;                                                    // (a new root's head writes sp[root] = root)
        if (__any_sync(0xffffffffu, conn && above_root != root)) {       // a run joining two roots: unite them (warp-uniform branch)
            __syncwarp();
            if (conn && above_root != root) {
                int a = root, b = above_root;
                bool done;
                do {
                    a = sm_find(sp, a); b = sm_find(sp, b);
                    if (a < b) { const int old = atomicMin(&sp[b], a); done = old == b; b = old; }
                    else if (b < a) { const int old = atomicMin(&sp[a], b); done = old == a; a = old; }
                    else done = true;
                } while (!done);
            }
            __syncwarp();
        }
        above_root = root;
    }
    __syncthreads();
    // slab borders: rows 4, 8, .., 28 with the row above (the left neighbour carries the union when it has the same pair)
    if (threadIdx.x < 7 * CT) {
        const int ly = 4 * (1 + (threadIdx.x >> 5)), t = ly * CT + lx;
        const int m = sl[ly][lx];
        if (m >= 0 && sl[ly - 1][lx] == m && !(lx > 0 && sl[ly][lx - 1] == m && sl[ly - 1][lx - 1] == m)) {
            int a = t, b = t - CT;
            bool done;
            do {
                a = sm_find(sp, a); b = sm_find(sp, b);
                if (a < b) { const int old = atomicMin(&sp[b], a); done = old == b; b = old; }
                else if (b < a) { const int old = atomicMin(&sp[a], b); done = old == a; a = old; }
                else done = true;
            } while (!done);
        }
    }
    __syncthreads();
    // local roots of the tile -> a shared list (typically 3-6 entries)
    int r[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int ly = 4 * wq + k, y = blockIdx.y * CT + ly;
        r[k] = sm_find(sp, ly * CT + lx);
        if (x < w && y < h && r[k] == ly * CT + lx) s_roots[atomicAdd(&s_n, 1)] = r[k];
    }
    __syncthreads();
    const int nr = s_n;
    if (nr <= 16) {
        // few roots: every warp counts its pixels of root j (one REDUX), one shared atomic per warp and root
        for (int j = 0; j < nr; j++) {
            const int rj = s_roots[j];
            int c = (r[0] == rj) + (r[1] == rj) + (r[2] == rj) + (r[3] == rj);
            c = __reduce_add_sync(0xffffffffu, c);
            if (lx == 0 && c) atomicAdd(&sc[rj], c);
        }
    } else {
        // many roots (noise): the head of every run of equal root along a row adds the run's length
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int prevr = __shfl_up_sync(0xffffffffu, r[k], 1);
            const unsigned hb = __ballot_sync(0xffffffffu, lx == 0 || prevr != r[k]);
            if ((hb >> lx) & 1u) {
                const unsigned above = lx == 31 ? 0u : hb >> (lx + 1);
                atomicAdd(&sc[r[k]], above ? __ffs(above) : 32 - lx);    // (pixels outside the image are their own roots, never listed)
            }
        }
    }
    if (threadIdx.x == 0 && nr) s_base = atomicAdd(nroots, nr);          // one global atomic per tile
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int y = blockIdx.y * CT + 4 * wq + k;
        if (x < w && y < h) P[y * w + x] = (blockIdx.y * CT + (r[k] >> 5)) * w + blockIdx.x * CT + (r[k] & 31);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < nr; j += 256) {
        const int rj = s_roots[j];
        const int gi = (blockIdx.y * CT + (rj >> 5)) * w + blockIdx.x * CT + (rj & 31);
        csize[gi] = sc[rj];
        roots[s_base + j] = gi;
    }
}
// one thread per pixel of a tile's top row / left column
__global__ void ccl_border_kernel(const int* __restrict__ lab, int lp, int w, int h, int* __restrict__ P)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int tiles_y = (h + CT - 1) / CT, tiles_x = (w + CT - 1) / CT;
    const int nh = (tiles_y - 1) * w;                                    // horizontal borders: rows CT, 2CT, ...
    if (i < nh) {
        const int y = (i / w + 1) * CT, x = i % w;
        const int* row = lab + (size_t)y * lp;
        const int me = row[x];
        if (row[x - lp] != me) return;
        if (x % CT != 0 && row[x - 1] == me && row[x - lp - 1] == me) return;   // the left neighbour (same tile pair) carries it
        uf_unite(P, y * w + x, (y - 1) * w + x);
        return;
    }
    const int j = i - nh;
    if (j >= (tiles_x - 1) * h) return;
    const int x = (j / h + 1) * CT, y = j % h;
    const int* row = lab + (size_t)y * lp;
    if (row[x] == row[x - 1]) uf_unite(P, y * w + x, y * w + x - 1);
}
// local root -> global root; its pixels are added to the global root's size
__global__ void ccl_roots_kernel(int* __restrict__ P, int* __restrict__ csize, const int* __restrict__ roots, const int* __restrict__ nroots)
{
    const int n = *nroots;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {      // (the host does not know n)
        const int lr = roots[j];
        const int g = uf_find(P, lr);
        if (g != lr) { P[lr] = g; atomicAdd(&csize[g], csize[lr]); }
    }
}
// global root of the pixel next to a root pixel: pixel -> local root -> global root (the list pass flattened the local roots)
__device__ __forceinline__ int root2(const int* __restrict__ P, int i) { return P[P[i]]; }
// a small root's link: the component supplying upstream's `adjlabel` = last of the L,U,R,D neighbours of the root pixel whose
// component starts earlier (left and up always do); -1: no neighbour labelled yet (pixel 0)
__device__ __forceinline__ int small_root_link(const int* __restrict__ P, int i, int w, int h)
{
    const int x = i % w, y = i / w;
    int adj = -1;
    if (x > 0) adj = root2(P, i - 1);
    if (y > 0) adj = root2(P, i - w);
    if (x < w - 1) { int r = root2(P, i + 1); if (r < i) adj = r; }
    if (y < h - 1) { int r = root2(P, i + w); if (r < i) adj = r; }
    return adj;
}
__global__ void ccl_keep_kernel(const int* __restrict__ P, const int* __restrict__ csize, const int* __restrict__ roots, const int* __restrict__ nroots,
                                int min_sp, unsigned* __restrict__ bitmap, int* __restrict__ link, int w, int h)
{
    const int n = *nroots;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const int r = roots[j];
        if (P[r] != r) continue;
        if (csize[r] > min_sp) atomicOr(&bitmap[r >> 5], 1u << (r & 31));
        else link[r] = small_root_link(P, r, w, h);
    }
}
// exclusive prefix popcount of the bitmap words: block sums -> block offsets (scan_blocks_kernel) -> per-word prefix
#define SCANW_CHUNK 1024
__global__ void __launch_bounds__(256) scanw_count_kernel(const unsigned* __restrict__ bitmap, int nwords, int* __restrict__ blocksum)
{
    __shared__ int s[8];
    const int base = blockIdx.x * SCANW_CHUNK;
    int c = 0;
    for (int k = threadIdx.x; k < SCANW_CHUNK; k += 256) { const int i = base + k; if (i < nwords) c += __popc(bitmap[i]); }
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
__global__ void __launch_bounds__(256) scanw_apply_kernel(const unsigned* __restrict__ bitmap, int nwords, const int* __restrict__ blocksum,
                                                           int* __restrict__ prefix)
{
    __shared__ int s_w[8];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = blocksum[blockIdx.x];
    __syncthreads();
    const int base = blockIdx.x * SCANW_CHUNK;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int k = 0; k < SCANW_CHUNK; k += 256) {
        const int i = base + k + threadIdx.x;
        const int v = i < nwords ? __popc(bitmap[i]) : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) s_w[wid] = inc;
        __syncthreads();
        int off = s_carry;
        for (int q = 0; q < wid; q++) off += s_w[q];
        if (i < nwords) prefix[i] = off + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int q = 0; q < 8; q++) t += s_w[q]; s_carry += t; }
        __syncthreads();
    }
}
// new label of a pixel = id of its component's root if kept, otherwise of the kept root its chain of links ends at (links
// strictly decrease, so the walk terminates; a chain that ends at a small component holding pixel 0 keeps upstream's adjlabel 0)
__device__ __forceinline__ int ccl_new_id(const int* __restrict__ P, const int* __restrict__ csize, const int* __restrict__ link,
                                          const unsigned* __restrict__ bitmap, const int* __restrict__ prefix, int lr, int min_sp)
{
    int r = P[lr];
    while (r >= 0 && csize[r] <= min_sp) r = link[r];
    return r >= 0 ? prefix[r >> 5] + __popc(bitmap[r >> 5] & ((1u << (r & 31)) - 1u)) : 0;
}
// list: the new id of every local root, written over its own parent entry as -1 - id.  Only that thread reads the entry (the
// walk to a kept root goes through csize / link, indexed by GLOBAL roots), and only the relabel pass reads P afterwards.
__global__ void ccl_final_kernel(int* __restrict__ P, const int* __restrict__ csize, const int* __restrict__ link,
                                 const unsigned* __restrict__ bitmap, const int* __restrict__ prefix, const int* __restrict__ roots,
                                 const int* __restrict__ nroots, int min_sp)
{
    const int n = *nroots;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const int lr = roots[j];
        P[lr] = -1 - ccl_new_id(P, csize, link, bitmap, prefix, lr, min_sp);
    }
}
// 4 pixels per thread: pixel -> local root -> -1 - id (a local root pixel holds its id itself); neighbours mostly share the root
__device__ __forceinline__ int ccl_id_of(const int* __restrict__ P, int q) { return -1 - (q >= 0 ? P[q] : q); }
__global__ void __launch_bounds__(256) ccl_relabel_kernel(const int* __restrict__ P, int* __restrict__ lab, int lp, int w, int h, int vec)
{
    const int x0 = 4 * (blockIdx.x * blockDim.x + threadIdx.x), y = blockIdx.y;
    if (x0 >= w) return;
    const int* prow = P + (size_t)y * w + x0;
    int* lrow = lab + (size_t)y * lp + x0;
    if (vec && x0 + 4 <= w) {
        const int4 q = *reinterpret_cast<const int4*>(prow);
        int4 o;
        o.x = ccl_id_of(P, q.x);
        o.y = q.y == q.x ? o.x : ccl_id_of(P, q.y);
        o.z = q.z == q.y ? o.y : ccl_id_of(P, q.z);
        o.w = q.w == q.z ? o.z : ccl_id_of(P, q.w);
        *reinterpret_cast<int4*>(lrow) = o;
    } else {
        const int n = min(4, w - x0);
        for (int j = 0; j < n; j++) lrow[j] = ccl_id_of(P, prow[j]);
    }
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
    dim3 b2(256), g4(cdiv(cdiv(w, 4), 256), h);
    const int vec4 = (w % 4 == 0 && lp % 4 == 0 && ((uintptr_t)d_labels & 15u) == 0 && ((uintptr_t)ws.parent & 15u) == 0) ? 1 : 0;
    // scratch: the list of local roots in newid [N ints]; bitmap + per-word prefix in diff [N bytes >= 2 * (N/32 + 1) ints]
    int* roots = ws.newid;
    const int nwords = cdiv(n, 32);
    unsigned* bitmap = reinterpret_cast<unsigned*>(ws.diff);
    int* prefix = reinterpret_cast<int*>(ws.diff) + nwords;
    int* nroots = ws.flag + 1;
    SPX_CUDA(cudaMemsetAsync(nroots, 0, sizeof(int), st));
    SPX_CUDA(cudaMemsetAsync(bitmap, 0, (size_t)nwords * sizeof(unsigned), st));
    ccl_local_kernel<<<dim3(cdiv(w, CT), cdiv(h, CT)), 256, 0, st>>>(d_labels, lp, w, h, ws.parent, ws.csize, roots, nroots);
    const int nborder = (cdiv(h, CT) - 1) * w + (cdiv(w, CT) - 1) * h;
    if (nborder > 0) ccl_border_kernel<<<cdiv(nborder, 256), 256, 0, st>>>(d_labels, lp, w, h, ws.parent);
    // the list kernels walk *nroots entries with a fixed grid (typical: 3-5 roots per tile; the host never learns the count)
    const int lb = std::max(1, std::min(cdiv(16 * cdiv(w, CT) * cdiv(h, CT), 256), 148 * 8));
    ccl_roots_kernel<<<lb, 256, 0, st>>>(ws.parent, ws.csize, roots, nroots);
    ccl_keep_kernel<<<lb, 256, 0, st>>>(ws.parent, ws.csize, roots, nroots, min_sp, bitmap, d_link, w, h);
    const int nb = cdiv(nwords, SCANW_CHUNK);
    scanw_count_kernel<<<nb, 256, 0, st>>>(bitmap, nwords, ws.blocksum);
    scan_blocks_kernel<<<1, 1024, 0, st>>>(ws.blocksum, nb, ws.flag);
    scanw_apply_kernel<<<nb, 256, 0, st>>>(bitmap, nwords, ws.blocksum, prefix);
    ccl_final_kernel<<<lb, 256, 0, st>>>(ws.parent, ws.csize, d_link, bitmap, prefix, roots, nroots, min_sp);
    ccl_relabel_kernel<<<g4, b2, 0, st>>>(ws.parent, d_labels, lp, w, h, vec4);
    ws.launches += nborder > 0 ? 9 : 8;
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
// MODE 0: write diff and clear the mask (the flags of the recurrence live in the mask buffer itself, 0 / 255);
// MODE 1: pure stencil (SEEDS flavour with thick_line): mask = popc(diff) > thr.
template <int MODE>
__global__ void contour_diff_kernel(const int* __restrict__ lab, int lp, int w, int h, unsigned char* __restrict__ diff,
                                    unsigned char* __restrict__ mask, size_t mp, int thr)
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
    if (MODE == 0) { diff[(size_t)y * w + x] = (unsigned char)d; mask[(size_t)y * mp + x] = 0; }
    else mask[(size_t)y * mp + x] = __popc(d) > thr ? 255 : 0;
}

// The recurrence  taken(p) = popc(diff(p)) - #{causal differing neighbours already taken} > lw  only looks backwards in
// raster order, so it has a unique fixed point and ANY evaluation order that keeps re-evaluating until nothing changes
// reaches it.  One persistent grid (all CTAs co-resident, cooperative launch) sweeps the image; a thread owns a run of
// 64 pixels of one row and walks it left to right in place, so dependencies along a row resolve inside the sweep and
// every sweep moves the settled front at least one row down; a grid-wide barrier separates the sweeps; the kernel ends
// after the first sweep that changed nothing.  No host round trip.
constexpr int CT_RUN = 64;
__device__ __forceinline__ void contour_grid_barrier(unsigned* counter, unsigned target)
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
__global__ void __launch_bounds__(1024) contour_sweeps_kernel(const unsigned char* __restrict__ diff, unsigned char* mask, size_t mp,
                                                               int w, int h, int lw, int vec, unsigned* bar, int* count /* [3] */,
                                                               int* dirty, int* list0, int* list1)
{
    const int rpr = (w + CT_RUN - 1) / CT_RUN;
    const int total = rpr * h;
    const unsigned nblk = gridDim.x;
    // `full`: this sweep visits every run (the first one, and the VERIFICATION sweep after the lists have run dry: the kernel ends
    // only when a full sweep changes nothing, i.e. at the fixed point whatever the list bookkeeping missed -- at 16384 x 16384 the
    // list-only form left a handful of pixels unsettled in most runs)
    bool full = true;
    for (unsigned sweep = 0;; sweep++) {
        // a run is re-evaluated only if one of its inputs (the run to its left, the three runs above) changed in the
        // previous sweep: a changed run appends its dependents to the next sweep's work list (once: `dirty` flag).
        // count[s%3] = length of sweep s's list, filled during sweep s-1, cleared during sweep s+1.
        const int* cur = (sweep & 1) ? list1 : list0;
        int* nxt = (sweep & 1) ? list0 : list1;
        int* ncount = &count[(sweep + 1) % 3];
        if (blockIdx.x == 0 && threadIdx.x == 0) count[(sweep + 2) % 3] = 0;
        const int nwork = full ? total : *(volatile int*)&count[sweep % 3];
        for (int wi = blockIdx.x * blockDim.x + threadIdx.x; wi < nwork; wi += nblk * blockDim.x) {
            const int run = full ? wi : __ldcg(cur + wi);
            if (sweep > 0) dirty[run] = 0;
            bool touched = false;
            const int y = run / rpr, rx = run - y * rpr, x0 = rx * CT_RUN;
            const int n = min(CT_RUN, w - x0);
            const unsigned char* drow = diff + (size_t)y * w + x0;
            unsigned char* mrow = mask + (size_t)y * mp + x0;
            const unsigned char* arow = mask + (size_t)(y - 1) * mp + x0;                // only read when y > 0
            // flags are read through L2 (other CTAs wrote them in the previous sweep)
            unsigned prev = x0 > 0 ? (__ldcg(mrow - 1) != 0) : 0;                         // W of the first pixel
            unsigned aL = (y > 0 && x0 > 0) ? (__ldcg(arow - 1) != 0) : 0;
            if (vec && n == CT_RUN) {
                // whole run in registers: 4 x 16 bytes of diff, own flags and the flags of the row above
                uint4 dq[4], mq[4], aq[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    dq[k] = __ldg(reinterpret_cast<const uint4*>(drow) + k);
                    mq[k] = __ldcg(reinterpret_cast<const uint4*>(mrow) + k);
                    aq[k] = y > 0 ? __ldcg(reinterpret_cast<const uint4*>(arow) + k) : make_uint4(0u, 0u, 0u, 0u);
                }
                const unsigned aEnd = (y > 0 && x0 + CT_RUN < w) ? (__ldcg(arow + CT_RUN) != 0) : 0;
                const unsigned* dw = reinterpret_cast<const unsigned*>(dq);
                unsigned* mw = reinterpret_cast<unsigned*>(mq);
                const unsigned* aw = reinterpret_cast<const unsigned*>(aq);
                bool rc = false;
                unsigned aC = aw[0] & 0xffu ? 1u : 0u;
#pragma unroll
                for (int j = 0; j < CT_RUN; j++) {
                    const unsigned aR = j + 1 < CT_RUN ? ((aw[(j + 1) >> 2] >> (8 * ((j + 1) & 3))) & 0xffu ? 1u : 0u) : aEnd;
                    const unsigned d = (dw[j >> 2] >> (8 * (j & 3))) & 0xffu;
                    int np = __popc(d);
                    unsigned t = 0;
                    if (np > lw) {
                        np -= (int)(((d & 1u) ? prev : 0u) + ((d & 2u) ? aL : 0u) + ((d & 4u) ? aC : 0u) + ((d & 8u) ? aR : 0u));
                        t = np > lw ? 1u : 0u;
                        const unsigned old = (mw[j >> 2] >> (8 * (j & 3))) & 0xffu ? 1u : 0u;
                        if (old != t) { mw[j >> 2] ^= 0xffu << (8 * (j & 3)); rc = true; }
                    }
                    prev = t; aL = aC; aC = aR;
                }
                if (rc) {
#pragma unroll
                    for (int k = 0; k < 4; k++) reinterpret_cast<uint4*>(mrow)[k] = mq[k];
                    touched = true;
                }
            } else {
                unsigned aC = y > 0 ? (__ldcg(arow) != 0) : 0;
                for (int j = 0; j < n; j++) {
                    const unsigned aR = (y > 0 && x0 + j + 1 < w) ? (__ldcg(arow + j + 1) != 0) : 0;
                    const unsigned d = drow[j];
                    int np = __popc(d);
                    unsigned t = 0;
                    if (np > lw) {
                        np -= (int)(((d & 1u) ? prev : 0u) + ((d & 2u) ? aL : 0u) + ((d & 4u) ? aC : 0u) + ((d & 8u) ? aR : 0u));
                        t = np > lw ? 1u : 0u;
                        const unsigned old = __ldcg(mrow + j) != 0;
                        if (old != t) { mrow[j] = t ? 255 : 0; touched = true; }
                    }
                    prev = t; aL = aC; aC = aR;
                }
            }
            if (touched) {
                auto push = [&](int r) { if (atomicExch(&dirty[r], 1) == 0) nxt[atomicAdd(ncount, 1)] = r; };
                if (rx + 1 < rpr) push(run + 1);
                if (y + 1 < h) {
                    if (rx > 0) push(run + rpr - 1);
                    push(run + rpr);
                    if (rx + 1 < rpr) push(run + rpr + 1);
                }
            }
        }
        contour_grid_barrier(bar, (sweep + 1) * nblk);
        if (*(volatile int*)ncount == 0) {
            if (full) break;
            full = true;
        } else full = false;
    }
}

// ---- the same recurrence, row by row, as a PREFIX COMPOSITION (the shipped form for widths up to 16384) ----------------
// Step 1, fully parallel (contour_pre_kernel): per pixel one byte = popc(diff) << 4 | causal difference bits (W, NW, N, NE).
// Step 2 (contour_band_kernel): once the row above is known, pixel x of a row is one of three functions of its west flag:
//     c = popc(diff) - #{NW, N, NE differ and are taken};   no west difference: t = [c > lw]   (constant)
//     west differs: c > lw+1 -> 1, c <= lw -> 0 (constants), c == lw+1 -> t = NOT t(x-1)
// Functions {0, 1, NOT, ID} are closed under composition (2 bits: f(0) | f(1) << 1), and a composition is "the last constant,
// flipped once per NOT after it": a thread composes its P pixels, a warp resolves its 32 threads with three ballots, the CTA
// its warps the same way from one word per warp in shared memory (total function and the function of the warp's first
// pixel) -- ONE barrier per row: the flag left of a thread is its carry-in, the flag right of a warp follows from the
// neighbour's first-pixel function and the warp's own carry-out.  The pre bytes arrive through a private cp.async ring.
// Rows stay sequential, so the image is cut into horizontal BANDS, one CTA each, all co-resident (cooperative launch).  A
// band needs the flags of the row above it: round 0 starts `warm` rows early from all-zero flags (errors of the guess die
// out within a few rows) and remembers which flags of the row above it used; after a grid-wide barrier every band compares
// them with what the band above really produced and, if they differ, recomputes from the true row -- repeated until a round
// recomputes nothing.  The recurrence only looks backwards, so that fixed point is unique: the result is upstream's
// sequential scan, bit for bit.
constexpr int CB_THREADS = 1024;
constexpr int CB_WARM = 8;
constexpr int CB_MAXW = 16384;             // one CTA walks whole rows: 1024 threads x 16 pixels
constexpr int CB_AHEAD = 4;                 // rows of pre bytes in flight (cp.async); ring of CB_AHEAD + 1 rows in shared memory
constexpr int CB_RING = CB_AHEAD + 1;
struct ContourBand {
    const unsigned char* pre; size_t pp;    // packed per-pixel byte, pitch pp (multiple of 16)
    int w, h, lw;
    unsigned char* mask; size_t mp;
    unsigned char* halo;          // [bands][w] flags of the row above each band that its current rows were computed from
    int rows; int bands; int warm;
    unsigned* bar; int* cnt;      // cnt[0..2]: bands recomputed in round r, by r % 3; cnt[3..4]: diagnostics
};
// 4 pixels per thread: three label rows (one 16-byte load each where aligned, the two outer neighbours as scalar loads that
// hit L1) -> 4 packed bytes.  Measured alternatives: outer neighbours by shuffle 29 us, 4 rows per thread rolling 29 us, this 23 us.
__global__ void __launch_bounds__(256) contour_pre_kernel(const int* __restrict__ lab, int lp, int w, int h, unsigned char* __restrict__ pre, size_t pp, int vec)
{
    const int x0 = 4 * (blockIdx.x * blockDim.x + threadIdx.x), y = blockIdx.y;
    if (x0 >= w) return;
    int R[3][6];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        const int yy = min(max(y - 1 + r, 0), h - 1);
        const int* row = lab + (size_t)yy * lp;
        if (vec && x0 + 4 <= w) {
            const int4 v = __ldg(reinterpret_cast<const int4*>(row + x0));
            R[r][1] = v.x; R[r][2] = v.y; R[r][3] = v.z; R[r][4] = v.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) R[r][1 + j] = __ldg(row + min(x0 + j, w - 1));
        }
        R[r][0] = __ldg(row + max(x0 - 1, 0));
        R[r][5] = __ldg(row + min(x0 + 4, w - 1));
    }
    unsigned rowmask = 0xffu;
    if (y == 0) rowmask &= ~(2u | 4u | 8u);
    if (y == h - 1) rowmask &= ~(32u | 64u | 128u);
    unsigned out = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int x = x0 + j;
        const int me = R[1][j + 1];
        unsigned d = (R[1][j] != me ? 1u : 0u) | (R[0][j] != me ? 2u : 0u) | (R[0][j + 1] != me ? 4u : 0u) | (R[0][j + 2] != me ? 8u : 0u) |
                     (R[1][j + 2] != me ? 16u : 0u) | (R[2][j + 2] != me ? 32u : 0u) | (R[2][j + 1] != me ? 64u : 0u) | (R[2][j] != me ? 128u : 0u);
        unsigned vm = rowmask;
        if (x == 0) vm &= ~(1u | 2u | 128u);
        if (x == w - 1) vm &= ~(8u | 16u | 32u);
        d &= vm;
        if (x >= w) d = 0;
        out |= (((unsigned)__popc(d) << 4) | (d & 15u)) << (8 * j);
    }
    *reinterpret_cast<unsigned*>(pre + (size_t)y * pp + x0) = out;
}
__device__ __forceinline__ unsigned cb_compose(unsigned later, unsigned earlier)
{
    // table over (later << 2 | earlier): constants absorb, ID passes, NOT swaps the two bits of `earlier`
    //   later 0 -> 0, 3 -> 3, 2 (ID) -> earlier, 1 (NOT) -> {0->3, 1->2, 2->1, 3->0}
    return (0xFFE41B00u >> (2u * ((later << 2) | earlier))) & 3u;
}
__device__ __forceinline__ void cb_cp4(void* smem_dst, const void* gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cb_cp8(void* smem_dst, const void* gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cb_cp16(void* smem_dst, const void* gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cb_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cb_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// "the last constant, flipped once per NOT after it" over the entries strictly below position `pos`; none: cin flipped
__device__ __forceinline__ unsigned cb_resolve(unsigned mc, unsigned mv, unsigned mn, int pos, unsigned cin)
{
    const unsigned below = pos >= 32 ? 0xffffffffu : (1u << pos) - 1u;
    const unsigned lc = mc & below;
    if (lc) { const int p = 31 - __clz(lc); return ((mv >> p) & 1u) ^ (__popc(mn & below & ~((2u << p) - 1u)) & 1u); }
    return cin ^ (__popc(mn & below) & 1u);
}
// rows [ya, yb) of the band; rows < ystore are warm-up (computed, not stored).  On entry sflag (w + 2 bytes, x shifted by
// one, zero at both ends) holds the flags of row ya-1.  All threads of the CTA call; warps without pixels leave at once.
template <int P, int T>
__device__ void cb_rows(const ContourBand& a, int ya, int yb, int ystore, int yhalo, unsigned char* halo_out,
                        const unsigned char* sflag, unsigned* ring, unsigned* pub)
{
    constexpr int WPT = P / 4;                                      // words per thread and row
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = tid * P;
    const int nw = (a.w + 32 * P - 1) / (32 * P);                   // warps that own pixels
    if (warp >= nw) return;
    const int nthr = nw * 32;
    const bool active = x0 < a.w;
    auto fetch = [&](int y) {
        if (active && y < a.h) {
            unsigned* dst = ring + ((size_t)(y % CB_RING) * T + tid) * WPT;
            const unsigned char* src = a.pre + (size_t)y * a.pp + x0;
            if (WPT == 1) cb_cp4(dst, src); else if (WPT == 2) cb_cp8(dst, src); else cb_cp16(dst, src);
        }
        cb_commit();
    };
#pragma unroll
    for (int k = 0; k < CB_AHEAD; k++) fetch(ya + k);
    // flags of row ya-1 at x0-1 .. x0+P
    unsigned fl = 0;
#pragma unroll
    for (int j = 0; j < P + 2; j++) fl |= (x0 + j < a.w + 2 && sflag[x0 + j] ? 1u : 0u) << j;
    for (int y = ya; y < yb; y++) {
        fetch(y + CB_AHEAD);                                        // into the slot of row y-1
        cb_wait<CB_AHEAD>();                                        // row y has landed
        unsigned pw[WPT];
        {
            const unsigned* src = ring + ((size_t)(y % CB_RING) * T + tid) * WPT;
#pragma unroll
            for (int q = 0; q < WPT; q++) pw[q] = active ? src[q] : 0u;
        }
        unsigned codes = 0, tot = 2u;                               // ID
#pragma unroll
        for (int j = 0; j < P; j++) {
            const unsigned byte = (pw[j >> 2] >> (8 * (j & 3))) & 0xffu;
            // k = clamp(c - lw, 0, 2) with c = popc(diff) - taken causal-above neighbours; (k, west bit) -> function:
            // k = 0 -> 0; k = 1 -> west ? NOT : 1; k = 2 -> 1   (popc(diff) <= lw implies k = 0)
            const int k = min(max((int)(byte >> 4) - a.lw - __popc((byte >> 1) & (fl >> j) & 7u), 0), 2);
            unsigned f = (0xF70u >> (2u * (2u * (unsigned)k + (byte & 1u)))) & 3u;
            if (x0 + j >= a.w) f = 2u;                              // beyond the row: transparent
            codes |= f << (2 * j);
            tot = cb_compose(f, tot);
        }
        const unsigned mc = __ballot_sync(0xffffffffu, tot == 0u || tot == 3u);
        const unsigned mv = __ballot_sync(0xffffffffu, tot == 3u);
        const unsigned mn = __ballot_sync(0xffffffffu, tot == 1u);
        unsigned* prow = pub + (y & 1) * 32;
        if (lane == 0) {
            unsigned wt;                                            // the warp's total function
            if (mc) { const int p = 31 - __clz(mc); wt = (((mv >> p) & 1u) ^ (__popc(mn & ~((2u << p) - 1u)) & 1u)) ? 3u : 0u; }
            else wt = (__popc(mn) & 1) ? 1u : 2u;
            prow[warp] = wt | ((codes & 3u) << 2);
        }
        asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");
        const unsigned rec = lane < nw ? prow[lane] : 2u;
        const unsigned wt_ = rec & 3u;
        const unsigned wc = __ballot_sync(0xffffffffu, wt_ == 0u || wt_ == 3u);
        const unsigned wvv = __ballot_sync(0xffffffffu, wt_ == 3u);
        const unsigned wn = __ballot_sync(0xffffffffu, wt_ == 1u);
        const unsigned cw = cb_resolve(wc, wvv, wn, warp, 0u);      // carry into this warp = flag left of its first pixel
        unsigned t = cb_resolve(mc, mv, mn, lane, cw);              // carry into this thread = flag of pixel x0-1
        const unsigned left = t;
        unsigned bits = 0;
#pragma unroll
        for (int j = 0; j < P; j++) { t = (codes >> (2 * j + t)) & 1u; bits |= t << j; }
        if (!active) bits = 0;
        else if (a.w - x0 < P) bits &= (1u << (a.w - x0)) - 1u;
        if (active) {
            const int n = min(P, a.w - x0);
            if (y >= ystore) {
                unsigned char* mo = a.mask + (size_t)y * a.mp + x0;
                if (n == P && (((uintptr_t)mo) & 3u) == 0) {
#pragma unroll
                    for (int q = 0; q < WPT; q++) {
                        const unsigned b4 = bits >> (4 * q);
                        reinterpret_cast<unsigned*>(mo)[q] = ((b4 & 1u) ? 0xffu : 0u) | ((b4 & 2u) ? 0xff00u : 0u) | ((b4 & 4u) ? 0xff0000u : 0u) | ((b4 & 8u) ? 0xff000000u : 0u);
                    }
                } else {
                    for (int j = 0; j < n; j++) mo[j] = (bits >> j) & 1u ? 255 : 0;
                }
            }
            if (y == yhalo) for (int j = 0; j < n; j++) halo_out[x0 + j] = (unsigned char)((bits >> j) & 1u);
        }
        // flags of this row at x0-1 .. x0+P for the next one
        unsigned right = __shfl_down_sync(0xffffffffu, bits & 1u, 1);
        const unsigned recn = __shfl_sync(0xffffffffu, rec, (warp + 1) & 31);
        if (lane == 31) {
            const unsigned cnext = cb_resolve(wc, wvv, wn, warp + 1, 0u);                  // carry out of this warp
            right = warp + 1 < nw ? ((recn >> 2) >> cnext) & 1u : 0u;
        }
        fl = left | (bits << 1) | (right << (P + 1));
    }
    cb_wait<0>();
}
template <int P, int T>
__global__ void __launch_bounds__(T, T == 512 ? 2 : 1) contour_band_kernel(ContourBand a)
{
    extern __shared__ unsigned char cb_smem[];
    __shared__ unsigned pub[2 * 32];
    __shared__ int s_diff;
    unsigned char* sflag = cb_smem;
    const int fw = a.w + 2;
    unsigned* ring = reinterpret_cast<unsigned*>(cb_smem + ((fw + 15) & ~15));
    const int b = blockIdx.x;
    const int y0 = b * a.rows, y1 = min(y0 + a.rows, a.h);
    const unsigned nblk = gridDim.x;
    unsigned char* myhalo = a.halo + (size_t)b * a.w;
    // ---- round 0: start `warm` rows early from all-zero flags
    {
        const int ya = max(y0 - a.warm, 0);
        for (int i = threadIdx.x; i < fw; i += T) sflag[i] = 0;
        __syncthreads();
        cb_rows<P, T>(a, ya, y1, y0, y0 - 1, myhalo, sflag, ring, pub);
    }
    contour_grid_barrier(a.bar, nblk);
    for (unsigned round = 1;; round++) {
        int* cnt = a.cnt + round % 3;
        if (b == 0 && threadIdx.x == 0) a.cnt[(round + 1) % 3] = 0;
        if (b > 0) {
            // the flags this band's rows were computed from against the row above as it stands now (one snapshot, through L2)
            if (threadIdx.x == 0) { s_diff = 0; sflag[0] = 0; sflag[fw - 1] = 0; }
            __syncthreads();
            const unsigned char* trow = a.mask + (size_t)(y0 - 1) * a.mp;
            bool differ = false;
            for (int x = threadIdx.x; x < a.w; x += T) {
                const unsigned char v = __ldcg(trow + x) ? 1 : 0;
                sflag[x + 1] = v;
                differ |= v != __ldcg(myhalo + x);
            }
            if (differ) s_diff = 1;
            __syncthreads();
            if (s_diff) {                                           // uniform
                for (int x = threadIdx.x; x < a.w; x += T) myhalo[x] = sflag[x + 1];
                if (threadIdx.x == 0) atomicAdd(cnt, 1);
                cb_rows<P, T>(a, y0, y1, y0, -2, myhalo, sflag, ring, pub);
            }
        }
        contour_grid_barrier(a.bar, (round + 1) * nblk);
        const int redone = *(volatile int*)cnt;
        if (b == 0 && threadIdx.x == 0) { a.cnt[3] = (int)round; a.cnt[4] += redone; }     // diagnostics: rounds, bands recomputed
        if (redone == 0) break;
    }
}

int contour_mask_dev(PostWorkspace& ws, const int* d_labels, int lp, int thick_line, int flavour,
                     unsigned char* d_mask, size_t mp, cudaStream_t st)
{
    const int w = ws.w, h = ws.h;
    dim3 b2(256), g2(cdiv(w, 256), h);
    if (flavour == 1 && thick_line) {
        contour_diff_kernel<1><<<g2, b2, 0, st>>>(d_labels, lp, w, h, ws.diff, d_mask, mp, 1);
        ws.launches++;
        SPX_CUDA(cudaGetLastError());
        return SPX_OK;
    }
    int lw = flavour == 1 ? 1 : (thick_line ? 2 : 1);
    static int s_sms = 0;
    if (!s_sms) {
        int dev = 0;
        SPX_CUDA(cudaGetDevice(&dev));
        SPX_CUDA(cudaDeviceGetAttribute(&s_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    if (w <= CB_MAXW && (((size_t)w + 15) & ~(size_t)15) * (size_t)h <= (size_t)w * h * sizeof(int) && !getenv("SPX_CONTOUR_SWEEPS")) {
        // row recurrence by prefix composition, one CTA per band of rows (see contour_band_kernel)
        ContourBand cb{};
        const size_t pp = ((size_t)w + 15) & ~(size_t)15;
        unsigned char* pre = reinterpret_cast<unsigned char*>(ws.newid);          // pp * h bytes <= 4 N (scratch of the connectivity pass)
        const int vecl = (lp % 4 == 0 && ((uintptr_t)d_labels & 15u) == 0) ? 1 : 0;
        contour_pre_kernel<<<dim3(cdiv(cdiv(w, 4), 256), h), 256, 0, st>>>(d_labels, lp, w, h, pre, pp, vecl);
        cb.pre = pre; cb.pp = pp; cb.w = w; cb.h = h; cb.lw = lw; cb.mask = d_mask; cb.mp = mp;
        cb.halo = ws.diff;                                   // bands * w bytes <= N
        cb.warm = lw >= 2 ? 12 : CB_WARM;
        int T = w > 8192 ? 1024 : 512;                       // 512: two CTAs per SM, their row phases interleave
        if (getenv("SPX_CB_T") && w <= 8192) T = atoi(getenv("SPX_CB_T")) == 1024 ? 1024 : 512;
        const int ctas = s_sms * (T == 512 ? 2 : 1);
        cb.rows = std::max(16, cdiv(h, ctas));                // (8 rows + 8 warm-up rows per band measured slower: 48 vs 41 us at 4K)
        if (getenv("SPX_CB_WARM")) cb.warm = atoi(getenv("SPX_CB_WARM"));
        if (getenv("SPX_CB_ROWS")) cb.rows = std::max(cdiv(h, ctas), atoi(getenv("SPX_CB_ROWS")));
        cb.bands = cdiv(h, cb.rows);
        SPX_CUDA(cudaMemsetAsync(ws.flag + 4, 0, 6 * sizeof(int), st));    // barrier counter + 3 round counters + 2 diagnostics
        cb.bar = reinterpret_cast<unsigned*>(ws.flag + 4); cb.cnt = ws.flag + 5;
        const int P = w > 8 * T ? 16 : (w > 4 * T ? 8 : 4);
        const void* fn = T == 1024 ? (P == 16 ? (const void*)contour_band_kernel<16, 1024> : P == 8 ? (const void*)contour_band_kernel<8, 1024> : (const void*)contour_band_kernel<4, 1024>)
                                   : (P == 16 ? (const void*)contour_band_kernel<16, 512> : P == 8 ? (const void*)contour_band_kernel<8, 512>
                                                                                                     : (const void*)contour_band_kernel<4, 512>);
        void* args[] = {(void*)&cb};
        const size_t smem = (((size_t)(w + 2) + 15) & ~(size_t)15) + (size_t)CB_RING * T * P;
        if (smem > 48 * 1024) SPX_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SPX_CUDA(cudaLaunchCooperativeKernel(fn, dim3(cb.bands), dim3(T), args, smem, st));
        ws.launches += 2;
        SPX_CUDA(cudaGetLastError());
        if (getenv("SPX_CONTOUR_DEBUG")) {
            SPX_CUDA(cudaMemcpyAsync(ws.h_flag + 8, ws.flag + 8, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
            SPX_CUDA(cudaStreamSynchronize(st));
            fprintf(stderr, "[contour] %dx%d lw %d: %d bands of %d rows, %d rounds, %d bands recomputed\n", w, h, lw, cb.bands, cb.rows, ws.h_flag[8], ws.h_flag[9]);
        }
        return SPX_OK;
    }
    contour_diff_kernel<0><<<g2, b2, 0, st>>>(d_labels, lp, w, h, ws.diff, d_mask, mp, 0);
    SPX_CUDA(cudaMemsetAsync(ws.flag + 4, 0, 4 * sizeof(int), st));        // barrier counter + 3 change flags
    // persistent grid: one CTA of 1024 threads per SM (the in-kernel barrier needs all of them running: cooperative launch)
    const int runs = cdiv(w, CT_RUN) * h;
    int blocks = std::min(cdiv(runs, 1024), s_sms);
    if (blocks < 1) blocks = 1;
    const unsigned char* diffp = ws.diff;
    // dirty flags and the two work lists live in the connectivity scratch, which is free here (3 * runs ints <= N ints)
    int* dirty = ws.csize;
    int* list0 = dirty + runs;
    int* list1 = list0 + runs;
    SPX_CUDA(cudaMemsetAsync(dirty, 0, (size_t)runs * sizeof(int), st));
    unsigned* bar = reinterpret_cast<unsigned*>(ws.flag + 4);
    int* cnt = ws.flag + 5;
    // 16-byte vector access to diff / mask rows when every run starts aligned
    int vec = (w % 16 == 0 && mp % 16 == 0 && (uintptr_t)d_mask % 16 == 0 && (uintptr_t)ws.diff % 16 == 0) ? 1 : 0;
    void* args[] = {(void*)&diffp, (void*)&d_mask, (void*)&mp, (void*)&w, (void*)&h, (void*)&lw, (void*)&vec, (void*)&bar, (void*)&cnt,
                    (void*)&dirty, (void*)&list0, (void*)&list1};
    SPX_CUDA(cudaLaunchCooperativeKernel((const void*)contour_sweeps_kernel, dim3(blocks), dim3(1024), args, 0, st));
    ws.launches += 2;
    SPX_CUDA(cudaGetLastError());
    return SPX_OK;
}

}  // namespace spx

// ------------------------------------------------------------------------------------------------
// stand-alone C ABI (host buffers)
// ------------------------------------------------------------------------------------------------
extern "C" int spx_enforce_label_connectivity(int32_t* labels, size_t stride, int width, int height,
                                              int numlabels, int min_element_size, int device, int* numlabels_out) try
{
    using namespace spx;
    SPX_REQUIRE(labels && width > 0 && height > 0 && height <= SPX_MAX_HEIGHT && (long long)width * height < (1ll << 31) && stride >= (size_t)width * 4 && stride % 4 == 0, SPX_ERR_BADARG,
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
} SPX_API_CATCH

extern "C" int spx_label_contour_mask(const int32_t* labels, size_t stride, int width, int height, int thick_line,
                                      int flavour, uint8_t* dst, size_t dst_stride, int device) try
{
    using namespace spx;
    SPX_REQUIRE(labels && dst && width > 0 && height > 0 && height <= SPX_MAX_HEIGHT && (long long)width * height < (1ll << 31) && stride >= (size_t)width * 4 && stride % 4 == 0 &&
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
} SPX_API_CATCH
