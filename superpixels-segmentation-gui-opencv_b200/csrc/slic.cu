// slic.cu -- SuperpixelSLIC (SLIC / SLICO / MSLIC) engine and its C ABI.
//
// Replaces cv::ximgproc::createSuperpixelSLIC / iterate / enforceLabelConnectivity / getLabels /
// getNumberOfSuperpixels / getLabelContourMask as called from mainwindow.cpp:2105-2111.
//
// Upstream's assignment is a cluster-ordered scatter ("for n in 0..K-1: for pixels in the window of
// n: if d < dist ..."), i.e. for every pixel the arg-min over the clusters whose window contains it,
// ties to the lowest n.  Here it is a gather: centres are binned by (int(kx)/S, int(ky)/S); a pixel
// in bin (bx,by) only has to look at bins bx-r..bx+r x by-r..by+r (r = 1; 2 for MSLIC's enlarged
// windows).  Same float32 operation order as upstream, no FMA contraction -> identical labels.
//
// This file holds the engine, the seeding / binning / centre-update kernels and the GENERIC
// per-pixel assignment kernel (any channel count, any region size).  The tiled fast path for
// 3-channel images lives in slic_tile.cu.
#include "slic.cuh"
#include "slic_generic.cuh"
#include <cfloat>
#include <chrono>
#include <map>
#include <set>
#include <vector>
#include <cstdlib>

namespace spx {

// implemented in slic_tile.cu; returns false when the configuration is outside the fast path
bool slic_tile_supported(const SlicGeom& g);
int slic_tile_prepare(const SlicGeom& g, const unsigned char* d_img, unsigned* d_img4, cudaStream_t st);
int slic_tile_assign(const SlicGeom& g, const SlicArrays& a, const unsigned char* d_img, const unsigned* d_img4, int* d_labels,
                     float* d_dcplane, const float* d_area_px, const int* d_K, int parity, int* d_overflow, cudaStream_t st);

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int refl101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
    return p;
}
// DetectChEdges at one pixel: sum_c (I(x+1)-I(x-1))^2 + (I(y+1)-I(y-1))^2 (Sobel ksize=1, REFLECT_101).
// All terms are integers < 2^24, so float32 evaluation order is irrelevant.
__device__ float edge_mag(const unsigned char* __restrict__ img, size_t pitch, int w, int h, int C, int x, int y)
{
    int xm = refl101(x - 1, w), xp = refl101(x + 1, w), ym = refl101(y - 1, h), yp = refl101(y + 1, h);
    int e = 0;
    for (int c = 0; c < C; c++) {
        int dx = (int)img[(size_t)y * pitch + (size_t)xp * C + c] - (int)img[(size_t)y * pitch + (size_t)xm * C + c];
        int dy = (int)img[(size_t)yp * pitch + (size_t)x * C + c] - (int)img[(size_t)ym * pitch + (size_t)x * C + c];
        e += dx * dx + dy * dy;
    }
    return (float)e;
}

struct SeedParams {
    int hex;                 // 0: GetChSeedsS strip grid, 1: GetChSeedsK hexagonal grid
    int xstrips, xerr_i, yerr_i;   // strip grid
    int ne, no;              // hex grid: seeds per even / odd row
};

// GetChSeedsS / GetChSeedsK + PerturbSeeds, one thread per seed
__global__ void slic_seed_kernel(SlicGeom g, SlicArrays a, const unsigned char* __restrict__ img, SeedParams sp)
{
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= g.K) return;
    const int S = g.S;
    int X, Y;
    if (!sp.hex) {
        int gy = n / sp.xstrips, gx = n - gy * sp.xstrips;
        X = gx * S + S / 2 + gx * sp.xerr_i;
        Y = gy * S + S / 2 + gy * sp.yerr_i;
        X = min(X, g.w - 1); Y = min(Y, g.h - 1);
    } else {
        int per2 = sp.ne + sp.no;
        int pr = n / per2, rem = n - pr * per2;
        int r = 2 * pr, gx = rem;
        if (rem >= sp.ne) { r++; gx = rem - sp.ne; }
        Y = r * S + S / 2;
        X = gx * S + ((S / 2) << (r & 1));
    }
    // PerturbSeeds: strict-min edge magnitude over the 8 neighbours, move only if BOTH coordinates change
    const int dx8[8] = {-1, -1, 0, 1, 1, 1, 0, -1};
    const int dy8[8] = {0, -1, -1, -1, 0, 1, 1, 1};
    int sx = X, sy = Y;
    float best = edge_mag(img, g.ipitch, g.w, g.h, g.C, X, Y);
#pragma unroll 1
    for (int i = 0; i < 8; i++) {
        int nx = X + dx8[i], ny = Y + dy8[i];
        if (nx >= 0 && nx < g.w && ny >= 0 && ny < g.h) {
            float e = edge_mag(img, g.ipitch, g.w, g.h, g.C, nx, ny);
            if (e < best) { best = e; sx = nx; sy = ny; }
        }
    }
    // strip mode: a seed belongs to the rank whose rows hold its grid position; the others contribute zeros to the sum
    // that assembles the seeds (spx_slic_create_strip)
    const bool mine = Y >= g.own_y0 && Y < g.own_y1;
    if (sx != X && sy != Y) { X = sx; Y = sy; }
    a.ioff[n] = S;
    a.adaptk[n] = 1.0f;
    if (!mine) {
        a.kx[n] = 0.f; a.ky[n] = 0.f; a.ikx[n] = 0; a.iky[n] = 0;
        for (int c = 0; c < g.C; c++) a.kc[(size_t)c * g.Kcap + n] = 0.f;
        return;
    }
    a.kx[n] = (float)X; a.ky[n] = (float)Y;
    a.ikx[n] = X; a.iky[n] = Y;
    for (int c = 0; c < g.C; c++) a.kc[(size_t)c * g.Kcap + n] = (float)img[(size_t)Y * g.ipitch + (size_t)X * g.C + c];
}

// ---- binning: centre n goes into its S x S bin (linked list, used by the generic kernel and the tile fallback) and,
// for the tiled path, its record is appended to the list of every tile its window [ix-off, ix+off) x [iy-off, iy+off)
// touches.  The scatter is done by the whole CTA over (centre, tile) pairs, so the global atomics of one centre do not
// queue behind each other.
constexpr int BIN_CTA = 128;
struct BinStage {                       // what the scatter needs about the CTA's centres
    float4 a[BIN_CTA];                  // kx, ky, k0, k1
    float4 b[BIN_CTA];                  // k2, n, ix, iy (ints as bits)
    int4 span[BIN_CTA];                 // tx0, ty0, tiles across, tiles down  (across = 0: nothing to scatter)
};
__device__ __forceinline__ void bin_stage(const SlicGeom& g, const SlicArrays& a, int par, BinStage& sh, int slot, int n, bool valid)
{
    int4 span = make_int4(0, 0, 0, 0);
    if (valid) {
        const int ix = a.ikx[n], iy = a.iky[n];
        int* head = par ? a.head[1] : a.head[0];
        int bx = min(max(ix, 0) / g.S, g.nbx - 1), by = min(max(iy, 0) / g.S, g.nby - 1);
        a.next[n] = atomicExch(&head[by * g.nbx + bx], n);
        if (g.tile_on) {
            const int off = a.ioff[n];
            const int tx0 = max(ix - off, 0) / TILE_W, tx1 = min(ix + off - 1, g.w - 1) / TILE_W;
            const int ty0 = max(max(iy - off, 0) / TILE_H, g.ty_lo), ty1 = min(min(iy + off - 1, g.h - 1) / TILE_H, g.ty_hi - 1);
            if (tx1 >= tx0 && ty1 >= ty0) span = make_int4(tx0, ty0, tx1 - tx0 + 1, ty1 - ty0 + 1);
            sh.a[slot] = make_float4(a.kx[n], a.ky[n], a.kc[n], a.kc[(size_t)g.Kcap + n]);
            sh.b[slot] = make_float4(a.kc[2 * (size_t)g.Kcap + n], __int_as_float(n), __int_as_float(ix), __int_as_float(iy));
        }
    }
    sh.span[slot] = span;
}
// all threads of the CTA; call after __syncthreads()
__device__ __forceinline__ void bin_scatter(const SlicGeom& g, const SlicArrays& a, int par, const BinStage& sh)
{
    if (!g.tile_on) return;
    int* cnt = par ? a.tl_count[1] : a.tl_count[0];
    TileRec* recs = par ? a.tl_rec[1] : a.tl_rec[0];
    // work item = (centre slot, j) with j in 0..15 walking the span in a 4x4 pattern (bigger spans: strided repeats)
#pragma unroll 4
    for (int i = threadIdx.x; i < BIN_CTA * 16; i += BIN_CTA) {
        const int slot = i >> 4, j = i & 15;
        const int4 sp = sh.span[slot];
        for (int ty = j >> 2; ty < sp.w; ty += 4)
            for (int tx = j & 3; tx < sp.z; tx += 4) {
                const int tile = (sp.y + ty) * g.tiles_x + sp.x + tx;
                const int pos = atomicAdd(&cnt[tile], 1);
                if (pos < TILE_CAP) {
                    float4* dst = reinterpret_cast<float4*>(&recs[(size_t)tile * TILE_CAP + pos]);
                    dst[0] = sh.a[slot];
                    dst[1] = sh.b[slot];
                }
            }
    }
}
__global__ void __launch_bounds__(BIN_CTA) slic_bin_kernel(SlicGeom g, SlicArrays a, const int* __restrict__ d_K, int parity)
{
    __shared__ BinStage sh;
    const int n = blockIdx.x * BIN_CTA + threadIdx.x;
    bin_stage(g, a, parity, sh, threadIdx.x, n, n < *d_K);
    __syncthreads();
    bin_scatter(g, a, parity, sh);
}
__global__ void fill_i32_kernel(int* p, int v, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void slico_begin_kernel(SlicGeom g, SlicArrays a, float* __restrict__ dcplane, size_t nplane)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nplane) dcplane[i] = FLT_MAX;
    if (i < (size_t)g.K) { a.maxc[i] = 10.0f * 10.0f; a.maxc_acc[i] = __float_as_uint(FLT_MIN); }
}

// ------------------------------------------------------------------------------------------------
// GENERIC assignment + accumulation: one thread per pixel, walks the bin lists in global memory.
// Exact for every configuration; used when the tiled kernel does not apply.
// ------------------------------------------------------------------------------------------------
template <int ALG>   // 100 SLIC, 101 SLICO, 102 MSLIC
__global__ void __launch_bounds__(256) slic_assign_generic_kernel(SlicGeom g, SlicArrays a, const unsigned char* __restrict__ img,
                                                                  int* __restrict__ labels, float* __restrict__ dcplane,
                                                                  int parity, int tile_x0, int tile_y0, int tile_w, int tile_h)
{
    int x = tile_x0 + blockIdx.x * 32 + (threadIdx.x & 31);
    int y = tile_y0 + blockIdx.y * 8 + (threadIdx.x >> 5);
    bool inside = x < min(g.w, tile_x0 + tile_w) && y < min(g.h, tile_y0 + tile_h);
    slic_pixel_generic<ALG>(g, a, img, labels, dcplane, parity, x, y, inside);
}

// ------------------------------------------------------------------------------------------------
// centre update (SeedNormInvoker): mean = float(sum)/float(count), count<=0 -> 1; then re-bin.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BIN_CTA) slic_finalize_kernel(SlicGeom g, SlicArrays a, const int* __restrict__ d_K, int parity, int do_bin)
{
    __shared__ BinStage sh;
    int t = blockIdx.x * BIN_CTA + threadIdx.x;
    const int nb = g.nbx * g.nby;
    pdl_trigger();
    pdl_wait();
    if (t < nb) (parity ? a.head[1] : a.head[0])[t] = -1;
    if (g.tile_on && t < g.tiles_x * g.tiles_y) (parity ? a.tl_count[1] : a.tl_count[0])[t] = 0;
    //              // the lists just consumed; reused two iterations later
    const int K = *d_K;
    if (t < K) {
        unsigned long long cnt = a.acc[(size_t)(g.C + 2) * g.Kcap + t];
        float m = (float)(cnt == 0 ? 1ull : cnt);
        for (int c = 0; c < g.C; c++) {
            a.kc[(size_t)c * g.Kcap + t] = __fdiv_rn((float)a.acc[(size_t)c * g.Kcap + t], m);
            a.acc[(size_t)c * g.Kcap + t] = 0;
        }
        float kx = __fdiv_rn((float)a.acc[(size_t)g.C * g.Kcap + t], m);
        float ky = __fdiv_rn((float)a.acc[(size_t)(g.C + 1) * g.Kcap + t], m);
        a.acc[(size_t)g.C * g.Kcap + t] = 0;
        a.acc[(size_t)(g.C + 1) * g.Kcap + t] = 0;
        a.acc[(size_t)(g.C + 2) * g.Kcap + t] = 0;
        a.kx[t] = kx; a.ky[t] = ky;
        a.ikx[t] = (int)kx; a.iky[t] = (int)ky;
        if (g.alg == 101) {                               // maxchans[k] <- accumulated maximum; seeds the next one
            unsigned mb = a.maxc_acc[t];
            a.maxc[t] = __uint_as_float(mb);
        }
    }
    if (!do_bin) return;                                  // uniform
    bin_stage(g, a, parity ^ 1, sh, threadIdx.x, t, t < K);
    __syncthreads();
    bin_scatter(g, a, parity ^ 1, sh);
}

// The same update for SLIC / SLICO with the scatter spread over threads: 16 threads per centre.  All 16 read the six sums (one
// broadcast load each) and compute the means redundantly; thread j = 0 writes the centre, empties its sums and re-bins it; thread
// j takes tile (j & 3, j >> 2) of the centre's span (larger spans: strided repeats).  No shared staging, no barrier: two dependent
// memory round trips (sums, list position) instead of five.
constexpr int FIN16_CTA = 128;
// Strip partition: a rank updates only the clusters [k_lo, k_hi) it maintains; a cluster whose centre has
// moved more than 4S - 2 rows away from its seed row could reach a band whose rank does not maintain it: *err = 100.
struct FinRange { int k_lo, k_hi; int guard, xstrips, yerr_i; int* err; };
// a rank maintains the clusters whose seed row lies within STRIP_MARGIN_S * S rows of its band; centres may then drift
// (STRIP_MARGIN_S - 1) * S rows from their seed row (measured on the synthetic frames: up to 39 rows = 2 S in 10 iterations)
constexpr int STRIP_MARGIN_S = 5;
__global__ void __launch_bounds__(FIN16_CTA) slic_finalize16_kernel(SlicGeom g, SlicArrays a, int parity, FinRange fr)
{
    const int t = blockIdx.x * FIN16_CTA + threadIdx.x;
    pdl_trigger();                                        // the next assignment may start loading its pixels
    pdl_wait();                                           // the assignment's sums
    const int nb = g.nbx * g.nby;
    if (t < nb) (parity ? a.head[1] : a.head[0])[t] = -1;                            // the lists just consumed; reused two iterations later
    if (g.tile_on && t < g.tiles_x * g.tiles_y) (parity ? a.tl_count[1] : a.tl_count[0])[t] = 0;
    const int n = fr.k_lo + (t >> 4), j = t & 15;
    if (n >= fr.k_hi) return;
    constexpr int C = 3;                                  // (launched for 3-channel images only)
    unsigned long long sum[C + 3];
    #pragma unroll
    for (int f = 0; f < C + 3; f++) sum[f] = a.acc[(size_t)f * g.Kcap + n];
    const int off = a.ioff[n];
    const float m = (float)(sum[C + 2] == 0 ? 1ull : sum[C + 2]);
    float kc[C];
    for (int c = 0; c < C; c++) kc[c] = __fdiv_rn((float)sum[c], m);
    const float kx = __fdiv_rn((float)sum[C], m), ky = __fdiv_rn((float)sum[C + 1], m);
    const int ix = (int)kx, iy = (int)ky;
    const int par = parity ^ 1;
    if (j == 0) {
        for (int f = 0; f < C + 3; f++) a.acc[(size_t)f * g.Kcap + n] = 0;          // (the group's loads are the same instruction, one step earlier)
        for (int c = 0; c < C; c++) a.kc[(size_t)c * g.Kcap + n] = kc[c];
        a.kx[n] = kx; a.ky[n] = ky; a.ikx[n] = ix; a.iky[n] = iy;
        if (g.alg == 101) a.maxc[n] = __uint_as_float(a.maxc_acc[n]);                 // maxchans[k] <- accumulated maximum; seeds the next one
        if (fr.guard) {
            const int gy = n / fr.xstrips;
            const int seed_y = min(gy * g.S + g.S / 2 + gy * fr.yerr_i, g.h - 1);
            if (abs(iy - seed_y) > (STRIP_MARGIN_S - 1) * g.S - 2 || fr.guard == 2) atomicExch(fr.err, 100);
        }
        int* head = par ? a.head[1] : a.head[0];
        const int bx = min(max(ix, 0) / g.S, g.nbx - 1), by = min(max(iy, 0) / g.S, g.nby - 1);
        a.next[n] = atomicExch(&head[by * g.nbx + bx], n);
    }
    if (!g.tile_on) return;
    const int tx0 = max(ix - off, 0) / TILE_W, tx1 = min(ix + off - 1, g.w - 1) / TILE_W;
    const int ty0 = max(max(iy - off, 0) / TILE_H, g.ty_lo), ty1 = min(min(iy + off - 1, g.h - 1) / TILE_H, g.ty_hi - 1);
    int* cnt = par ? a.tl_count[1] : a.tl_count[0];
    TileRec* recs = par ? a.tl_rec[1] : a.tl_rec[0];
    const float4 ra = make_float4(kx, ky, kc[0], kc[1]);
    const float4 rb = make_float4(kc[2], __int_as_float(n), __int_as_float(ix), __int_as_float(iy));
    for (int ty = ty0 + (j >> 2); ty <= ty1; ty += 4)
        for (int tx = tx0 + (j & 3); tx <= tx1; tx += 4) {
            const int tile = ty * g.tiles_x + tx;
            const int pos = atomicAdd(&cnt[tile], 1);
            if (pos < TILE_CAP) {
                float4* dst = reinterpret_cast<float4*>(&recs[(size_t)tile * TILE_CAP + pos]);
                dst[0] = ra;
                dst[1] = rb;
            }
        }
}

// The instruction-lean form of the same update: one WARP per 32 centres.  Phase 1: lane = centre (sums, means, centre arrays, bin
// list: the arithmetic is done once per centre).  Phase 2: the (centre, tile) pairs of the warp's 32 spans, 16 per centre and 32
// per step; a lane fetches its centre's record and span from the owning lane with shuffles and appends it to tile (j & 3, j >> 2)
// of the span (larger spans: strided repeats).  0.44 M warp-instructions per 20 k centres against 3.0 M of the 16-threads-per-
// centre kernel, at the price of a longer dependent chain (4K single frame: 0.63 vs 0.56 ms per iterate(10)): used where many
// centres are updated at once (8K frames, row strips of gigapixel images) and for objects in throughput mode.
__global__ void __launch_bounds__(FIN16_CTA) slic_finalize_warp_kernel(SlicGeom g, SlicArrays a, int parity, FinRange fr)
{
    const int t = blockIdx.x * FIN16_CTA + threadIdx.x;
    pdl_trigger();
    pdl_wait();
    const int nb = g.nbx * g.nby;
    for (int i = t; i < nb; i += gridDim.x * FIN16_CTA) (parity ? a.head[1] : a.head[0])[i] = -1;
    if (g.tile_on)
        for (int i = t; i < g.tiles_x * g.tiles_y; i += gridDim.x * FIN16_CTA) (parity ? a.tl_count[1] : a.tl_count[0])[i] = 0;
    const int lane = threadIdx.x & 31;
    const int n = fr.k_lo + t;
    if (n - lane >= fr.k_hi) return;                      // (whole warp)
    const bool valid = n < fr.k_hi;
    constexpr int C = 3;                                  // (launched for 3-channel images only)
    const int par = parity ^ 1;
    float kc0 = 0.f, kc1 = 0.f, kc2 = 0.f, kx = 0.f, ky = 0.f;
    int ix = 0, iy = 0, tx0 = 0, ty0 = 0, nx = 0, ny = 0;
    if (valid) {
        unsigned long long sum[C + 3];
#pragma unroll
        for (int f = 0; f < C + 3; f++) sum[f] = a.acc[(size_t)f * g.Kcap + n];
        const int off = a.ioff[n];
        const float m = (float)(sum[C + 2] == 0 ? 1ull : sum[C + 2]);
        kc0 = __fdiv_rn((float)sum[0], m); kc1 = __fdiv_rn((float)sum[1], m); kc2 = __fdiv_rn((float)sum[2], m);
        kx = __fdiv_rn((float)sum[C], m); ky = __fdiv_rn((float)sum[C + 1], m);
        ix = (int)kx; iy = (int)ky;
#pragma unroll
        for (int f = 0; f < C + 3; f++) a.acc[(size_t)f * g.Kcap + n] = 0;
        a.kc[n] = kc0; a.kc[(size_t)g.Kcap + n] = kc1; a.kc[2 * (size_t)g.Kcap + n] = kc2;
        a.kx[n] = kx; a.ky[n] = ky; a.ikx[n] = ix; a.iky[n] = iy;
        if (g.alg == 101) a.maxc[n] = __uint_as_float(a.maxc_acc[n]);
        if (fr.guard) {
            const int gy = n / fr.xstrips;
            const int seed_y = min(gy * g.S + g.S / 2 + gy * fr.yerr_i, g.h - 1);
            if (abs(iy - seed_y) > (STRIP_MARGIN_S - 1) * g.S - 2 || fr.guard == 2) atomicExch(fr.err, 100);
        }
        int* head = par ? a.head[1] : a.head[0];
        const int bx = min(max(ix, 0) / g.S, g.nbx - 1), by = min(max(iy, 0) / g.S, g.nby - 1);
        a.next[n] = atomicExch(&head[by * g.nbx + bx], n);
        if (g.tile_on) {
            tx0 = max(ix - off, 0) / TILE_W;
            const int tx1 = min(ix + off - 1, g.w - 1) / TILE_W;
            ty0 = max(max(iy - off, 0) / TILE_H, g.ty_lo);
            const int ty1 = min(min(iy + off - 1, g.h - 1) / TILE_H, g.ty_hi - 1);
            nx = max(tx1 - tx0 + 1, 0); ny = max(ty1 - ty0 + 1, 0);
        }
    }
    if (!g.tile_on) return;
    int* cnt = par ? a.tl_count[1] : a.tl_count[0];
    TileRec* recs = par ? a.tl_rec[1] : a.tl_rec[0];
    const int span = tx0 | (ty0 << 16), ext = nx | (ny << 16);
    const int j = lane & 15;
#pragma unroll 4
    for (int i = 0; i < 16; i++) {
        const int src = 2 * i + (lane >> 4);              // the centre (lane of phase 1) this lane works for in this step
        const float skx = __shfl_sync(0xffffffffu, kx, src), sky = __shfl_sync(0xffffffffu, ky, src);
        const float s0 = __shfl_sync(0xffffffffu, kc0, src), s1 = __shfl_sync(0xffffffffu, kc1, src), s2 = __shfl_sync(0xffffffffu, kc2, src);
        const int six = __shfl_sync(0xffffffffu, ix, src), siy = __shfl_sync(0xffffffffu, iy, src);
        const int sspan = __shfl_sync(0xffffffffu, span, src), sext = __shfl_sync(0xffffffffu, ext, src);
        const int sx0 = sspan & 0xffff, sy0 = sspan >> 16, snx = sext & 0xffff, sny = sext >> 16;
        const int sn = n - lane + src;
        const float4 ra = make_float4(skx, sky, s0, s1);
        const float4 rb = make_float4(s2, __int_as_float(sn), __int_as_float(six), __int_as_float(siy));
        for (int ty = j >> 2; ty < sny; ty += 4)
            for (int tx = j & 3; tx < snx; tx += 4) {
                const int tile = (sy0 + ty) * g.tiles_x + sx0 + tx;
                const int pos = atomicAdd(&cnt[tile], 1);
                if (pos < TILE_CAP) {
                    float4* dst = reinterpret_cast<float4*>(&recs[(size_t)tile * TILE_CAP + pos]);
                    dst[0] = ra;
                    dst[1] = rb;
                }
            }
    }
}

// ------------------------------------------------------------------------------------------------
// MSLIC content-adaptive step (normative statement: DESIGN.md, "MSLIC restatement")
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float tri_area5(const float* a, const float* b, const float* c)
{
    float uu = 0.f, vv = 0.f, uv = 0.f;
#pragma unroll
    for (int i = 0; i < 5; i++) {
        float u = __fsub_rn(b[i], a[i]), v = __fsub_rn(c[i], a[i]);
        uu = __fadd_rn(uu, __fmul_rn(u, u));
        vv = __fadd_rn(vv, __fmul_rn(v, v));
        uv = __fadd_rn(uv, __fmul_rn(u, v));
    }
    float gg = __fsub_rn(__fmul_rn(uu, vv), __fmul_rn(uv, uv));
    if (gg < 0.f) gg = 0.f;
    return __fmul_rn(0.5f, __fsqrt_rn(gg));
}
// The manifold area of the pixel quad at (x, y) depends on the image only: it is computed once per image into a float
// plane (label pitch) ...
__global__ void __launch_bounds__(256) mslic_area_px_kernel(SlicGeom g, const unsigned char* __restrict__ img, float* __restrict__ plane, float q)
{
    int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= g.w - 1 || y >= g.h - 1) return;
    float P[4][5];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        int xx = x + (j & 1), yy = y + (j >> 1);
        P[j][0] = (float)xx; P[j][1] = (float)yy;
        const unsigned char* p = img + (size_t)yy * g.ipitch + (size_t)xx * g.C;
#pragma unroll
        for (int c = 0; c < 3; c++) P[j][2 + c] = c < g.C ? __fmul_rn(q, (float)p[c]) : 0.f;
    }
    plane[(size_t)y * g.lp + x] = __fadd_rn(tri_area5(P[0], P[1], P[2]), tri_area5(P[3], P[1], P[2]));
}
// ... and every iteration only sums the plane per label (fixed point x256, exact 64-bit integer sums)
__global__ void __launch_bounds__(256) mslic_area_sum_kernel(SlicGeom g, SlicArrays a, const float* __restrict__ plane,
                                                              const int* __restrict__ labels)
{
    int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    bool inside = x < g.w - 1 && y < g.h - 1;
    long long fx = 0;
    int label = -1;
    if (inside) {
        fx = (long long)__fmul_rn(plane[(size_t)y * g.lp + x], 256.0f);
        label = labels[(size_t)y * g.lp + x];
    }
    // warp sum per label: the lanes of a row mostly share one label, so one ballot against lane 0's label settles the
    // common case; the 64-bit sum goes through three 32-bit REDUX (21-bit limbs: 32 lanes x 2^21 < 2^32)
    int key = inside ? label : -1 - (int)(threadIdx.x & 31);
    const int key0 = __shfl_sync(0xffffffffu, key, 0);
    unsigned m = __ballot_sync(0xffffffffu, key == key0);
    if (m != 0xffffffffu) m = __match_any_sync(0xffffffffu, key);
    const unsigned long long u = (unsigned long long)fx;                 // areas are >= 0 and < 2^63
    const unsigned l0 = (unsigned)(u & 0x1fffffu), l1 = (unsigned)((u >> 21) & 0x1fffffu), l2 = (unsigned)(u >> 42);
    const unsigned s0 = __reduce_add_sync(m, l0), s1 = __reduce_add_sync(m, l1), s2 = __reduce_add_sync(m, l2);
    if (inside && (int)(threadIdx.x & 31) == __ffs(m) - 1)
        atomicAdd((unsigned long long*)&a.area[label], (unsigned long long)s0 + ((unsigned long long)s1 << 21) + ((unsigned long long)s2 << 42));
}
// The content-adaptive step as four small grid-wide kernels (a single CTA walking 37 k clusters took 86 us at 8K):
//   total   sum of the cluster areas (exact 64-bit integer)
//   adaptk  mean area -> adaptk[k]; clusters to split counted per block and in total
//   split   (when the growth cap allows) new centres appended in ascending k: n = K + 3 * (flagged clusters before k)
//   finish  window half-sizes for all (old and new) clusters, areas cleared, K published
constexpr int MA = 256;
__device__ __forceinline__ bool mslic_flag(const SlicArrays& a, int k, float abar)
{
    const float ak = __fdiv_rn((float)a.area[k], 256.0f);
    return ak > __fmul_rn(4.0f, abar);                       // split = 4
}
__device__ __forceinline__ float mslic_abar(const SlicArrays& a, int K)
{
    const long long tot = *reinterpret_cast<const long long*>(a.ms_scratch);
    return __fdiv_rn((float)tot, __fmul_rn((float)K, 256.0f));
}
__global__ void __launch_bounds__(MA) mslic_total_kernel(SlicArrays a, const int* __restrict__ d_K)
{
    __shared__ long long s_red[MA / 32];
    const int k = blockIdx.x * MA + threadIdx.x, lane = threadIdx.x & 31;
    long long part = k < *d_K ? a.area[k] : 0;
    for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) s_red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t = 0;
        for (int i = 0; i < MA / 32; i++) t += s_red[i];
        if (t) atomicAdd(reinterpret_cast<unsigned long long*>(a.ms_scratch), (unsigned long long)t);
    }
}
__global__ void __launch_bounds__(MA) mslic_adaptk_kernel(SlicArrays a, const int* __restrict__ d_K)
{
    __shared__ int s_cnt[MA / 32];
    const int K = *d_K;
    const int k = blockIdx.x * MA + threadIdx.x, lane = threadIdx.x & 31;
    const float target = 2.0f;
    const float abar = mslic_abar(a, K);
    bool f = false;
    if (k < K) {
        const float ak = __fdiv_rn((float)a.area[k], 256.0f);
        float r = ak > 0.f ? __fsqrt_rn(__fdiv_rn(abar, ak)) : target;
        r = fminf(fmaxf(r, __fdiv_rn(1.0f, target)), target);
        a.adaptk[k] = r;
        f = mslic_flag(a, k, abar);
    }
    const unsigned m = __ballot_sync(0xffffffffu, f);
    if (lane == 0) s_cnt[threadIdx.x >> 5] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        int c = 0;
        for (int i = 0; i < MA / 32; i++) c += s_cnt[i];
        a.ms_scratch[4 + blockIdx.x] = c;
        if (c) atomicAdd(&a.ms_scratch[2], c);
    }
}
__global__ void __launch_bounds__(MA) mslic_split_kernel(SlicGeom g, SlicArrays a, const unsigned char* __restrict__ img,
                                                         const int* __restrict__ d_K, int K0)
{
    __shared__ int s_w[MA / 32];
    __shared__ int s_base;
    const int K = *d_K, nsplit = a.ms_scratch[2];
    const bool do_split = nsplit > 0 && K + 3 * nsplit <= 4 * K0 && K + 3 * nsplit <= g.Kcap;
    if (blockIdx.x == 0 && threadIdx.x == 0) a.ms_scratch[3] = do_split ? K + 3 * nsplit : K;
    if (!do_split) return;                                    // uniform
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    // flagged clusters in the blocks before this one
    int part = 0;
    for (int b = tid; b < (int)blockIdx.x; b += MA) part += a.ms_scratch[4 + b];
    for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) s_w[wid] = part;
    __syncthreads();
    if (tid == 0) { int t = 0; for (int i = 0; i < MA / 32; i++) t += s_w[i]; s_base = t; }
    __syncthreads();
    const int base = s_base;
    const int k = blockIdx.x * MA + tid;
    const float abar = mslic_abar(a, K);
    const bool f = k < K && mslic_flag(a, k, abar);
    const unsigned m = __ballot_sync(0xffffffffu, f);
    __syncthreads();
    if (lane == 0) s_w[wid] = __popc(m);
    __syncthreads();
    if (!f) return;
    int off = base + __popc(m & ((1u << lane) - 1));
    for (int q = 0; q < wid; q++) off += s_w[q];
    int n = K + 3 * off;
    const float o = __fmul_rn(__fmul_rn(0.5f, (float)g.S), a.adaptk[k]);
    const float cx = a.kx[k], cy = a.ky[k];
    const float ox[4] = {-o, o, -o, o}, oy[4] = {-o, -o, o, o};
    const float ad = a.adaptk[k];
    for (int j = 0; j < 4; j++) {
        float nx = __fadd_rn(cx, ox[j]), ny = __fadd_rn(cy, oy[j]);
        nx = fminf(fmaxf(nx, 0.f), (float)(g.w - 1));
        ny = fminf(fmaxf(ny, 0.f), (float)(g.h - 1));
        const int dst = j == 0 ? k : n++;
        a.kx[dst] = nx; a.ky[dst] = ny;
        a.ikx[dst] = (int)nx; a.iky[dst] = (int)ny;
        for (int c = 0; c < g.C; c++)
            a.kc[(size_t)c * g.Kcap + dst] = (float)img[(size_t)(int)ny * g.ipitch + (size_t)(int)nx * g.C + c];
        a.adaptk[dst] = ad;
    }
}
__global__ void __launch_bounds__(MA) mslic_finish_kernel(SlicGeom g, SlicArrays a, int* __restrict__ d_K)
{
    const int newK = a.ms_scratch[3];
    const int k = blockIdx.x * MA + threadIdx.x;
    if (k < newK) {
        a.ioff[k] = (int)__fmul_rn((float)g.S, a.adaptk[k]);
        a.area[k] = 0;
    }
    if (k == 0) *d_K = newK;                                  // nobody in this kernel reads *d_K
}
// scratch cleared for the next iteration (after finish: total and split count are read by the kernels before it)
__global__ void mslic_scratch_clear_kernel(SlicArrays a)
{
    if (threadIdx.x < 3) a.ms_scratch[threadIdx.x] = 0;
}

// ------------------------------------------------------------------------------------------------
// engine
// ------------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------
// Row-strip mode, peer exchange: the ONE exchange step of an iteration -- every rank needs the sums of
// the integer accumulators over all ranks -- done by the ranks' own kernels over NVLink peer memory instead of an NCCL
// all-reduce.  Every rank owns an exchange buffer (cudaMalloc, exported with cudaIpcGetMemHandle, mapped by all peers):
//     recv[2][world][n6] u64   slot [iteration parity][source rank]: the source's partial sums, written by the source
//     flag[world]        int   flag[source] = number of the last iteration whose sums the source has delivered
//     range[2][world]    int2  [lo, hi): the cluster index range the source delivered in that slot
// push: only the index range of clusters that received pixels from this rank's band travels (~K/world entries per field):
//       plain coalesced stores into the slot of every peer, a system-scope fence, and -- by the last CTA to finish -- the
//       range and then the iteration number into every peer's buffer.
// wait+sum: spins on the local flags (the peers write them), then adds the peers' slots to the local accumulators; the
//       ordinary centre update follows.  Integer sums: the result is bit-identical to the single-GPU object.
// Slots are double-buffered by iteration parity: a peer can deliver iteration i+2 only after it has seen this rank's flag
// for i+1, which is written after this rank's wait+sum of iteration i has finished reading slot i (stream order).
// ------------------------------------------------------------------------------------------------
constexpr int STRIP_MAX_WORLD = 16;
struct StripPeers {
    unsigned long long* base[STRIP_MAX_WORLD];     // exchange buffer of every rank in THIS process's address space
    int world, rank, Kcap, nf;                     // nf = channels + 3 fields of Kcap entries each
    size_t n6;                                     // nf * Kcap
    // neighbour mode (strip partition): only the two adjacent ranks exchange, and only the index range of the clusters both maintain
    int mode;                                      // 0: every rank to every rank (replicated cluster state); 1: neighbours only
    int nbr[2], nlo[2], nhi[2];                    // rank - 1 / rank + 1 (or -1) and the shared cluster range [nlo, nhi) with each
};
// layout of an exchange buffer: recv[2][world][n6] u64 | flag[MAX_WORLD] int | range[2][MAX_WORLD] int2
__device__ __forceinline__ unsigned long long* strip_slot(unsigned long long* base, const StripPeers& P, int par, int src)
{
    return base + ((size_t)par * P.world + src) * P.n6;
}
__device__ __forceinline__ int* strip_flags(unsigned long long* base, const StripPeers& P)
{
    return reinterpret_cast<int*>(base + (size_t)2 * P.world * P.n6);
}
__device__ __forceinline__ int2* strip_range(unsigned long long* base, const StripPeers& P, int par, int src)
{
    return reinterpret_cast<int2*>(strip_flags(base, P) + STRIP_MAX_WORLD) + par * STRIP_MAX_WORLD + src;
}
// the index range [lo, hi) of the clusters that received pixels from this rank's band (cluster indices are row-major in
// the seed grid, so a band of rows is nearly a contiguous index range: ~K/world entries instead of K)
__global__ void __launch_bounds__(256) strip_range_kernel(const unsigned long long* __restrict__ count, int K, int* __restrict__ range)
{
    int lo = 0x7fffffff, hi = -1;
    for (int k = blockIdx.x * 256 + threadIdx.x; k < K; k += gridDim.x * 256)
        if (count[k]) { lo = min(lo, k); hi = max(hi, k); }
    for (int o = 16; o; o >>= 1) { lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
    if ((threadIdx.x & 31) == 0 && hi >= 0) { atomicMin(&range[0], lo); atomicMax(&range[1], hi); }
}
__global__ void __launch_bounds__(256) strip_push_kernel(const __grid_constant__ StripPeers P, const unsigned long long* __restrict__ acc,
                                                          int iter, int* __restrict__ range, unsigned* __restrict__ done,
                                                          unsigned long long* __restrict__ bytes_sent)
{
    const int par = iter & 1;
    if (P.mode == 1) {
        for (int q = 0; q < 2; q++) {
            const int p = P.nbr[q], lo = P.nlo[q], len = P.nhi[q] - P.nlo[q];
            if (p < 0 || len <= 0) continue;
            unsigned long long* slot = strip_slot(P.base[p], P, par, P.rank);
            for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < (size_t)P.nf * len; i += (size_t)gridDim.x * 256) {
                const int f = (int)(i / len);
                const size_t j = (size_t)f * P.Kcap + lo + (i - (size_t)f * len);
                slot[j] = acc[j];
            }
        }
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0 && atomicAdd(done, 1u) == gridDim.x - 1) {
            *done = 0;
            for (int q = 0; q < 2; q++) {
                const int p = P.nbr[q];
                if (p < 0) continue;
                *bytes_sent += (unsigned long long)P.nf * max(P.nhi[q] - P.nlo[q], 0) * 8ull;
                volatile int* r = reinterpret_cast<volatile int*>(strip_range(P.base[p], P, par, P.rank));
                r[0] = P.nlo[q]; r[1] = max(P.nhi[q], P.nlo[q]);
            }
            __threadfence_system();
            for (int q = 0; q < 2; q++)
                if (P.nbr[q] >= 0) *reinterpret_cast<volatile int*>(&strip_flags(P.base[P.nbr[q]], P)[P.rank]) = iter;
        }
        return;
    }
    int lo = range[0], hi = range[1] + 1;
    if (hi <= lo) lo = hi = 0;                     // the band received no pixel at all
    const int len = hi - lo;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < (size_t)P.nf * len; i += (size_t)gridDim.x * 256) {
        const int f = (int)(i / len);
        const size_t j = (size_t)f * P.Kcap + lo + (i - (size_t)f * len);
        const unsigned long long v = acc[j];
        for (int p = 0; p < P.world; p++)
            if (p != P.rank) strip_slot(P.base[p], P, par, P.rank)[j] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned ticket = atomicAdd(done, 1u);
        if (ticket == gridDim.x - 1) {             // every CTA's stores are fenced: publish the range, then the flag
            *done = 0;
            range[0] = 0x7fffffff; range[1] = -1;
            *bytes_sent += (unsigned long long)P.nf * len * 8ull * (P.world - 1);
            for (int p = 0; p < P.world; p++)
                if (p != P.rank) {
                    volatile int* r = reinterpret_cast<volatile int*>(strip_range(P.base[p], P, par, P.rank));
                    r[0] = lo; r[1] = hi;
                }
            __threadfence_system();
            for (int p = 0; p < P.world; p++)
                if (p != P.rank) *reinterpret_cast<volatile int*>(&strip_flags(P.base[p], P)[P.rank]) = iter;
        }
    }
}
// The wait is bounded: a peer that died or skipped the exchange must not hang this rank in a kernel nobody can cancel.  After
// ~4 s (STRIP_SPIN_BUDGET sleeps of 200 ns, generous against any real skew between ranks) the kernel gives up, raises *err and
// adds nothing; the host reports SPX_ERR_GPU at its next synchronising call (spx_slic_get_label_rows / strip_exchanged_bytes).
constexpr unsigned STRIP_SPIN_BUDGET = 20u * 1000u * 1000u;
__global__ void __launch_bounds__(256) strip_wait_sum_kernel(const __grid_constant__ StripPeers P, unsigned long long* __restrict__ acc, int iter,
                                                              int* __restrict__ err)
{
    __shared__ int2 rg[STRIP_MAX_WORLD];
    __shared__ int s_fail;
    const int par = iter & 1;
    if (threadIdx.x == 0) s_fail = 0;
    __syncthreads();
    const bool is_src = (int)threadIdx.x < P.world && (int)threadIdx.x != P.rank &&
                        (P.mode == 0 || (int)threadIdx.x == P.nbr[0] || (int)threadIdx.x == P.nbr[1]);
    if (is_src) {
        const volatile int* f = strip_flags(P.base[P.rank], P) + threadIdx.x;
        unsigned spins = 0;
        while (*f < iter && *(volatile int*)err == 0) {
            __nanosleep(200);
            if (++spins > STRIP_SPIN_BUDGET) { atomicExch(err, 1 + (int)threadIdx.x); break; }
        }
        if (*f < iter) s_fail = 1;
        __threadfence_system();
        const volatile int2* r = strip_range(P.base[P.rank], P, par, threadIdx.x);
        rg[threadIdx.x] = make_int2(r->x, r->y);
    }
    __syncthreads();
    if (s_fail) return;
    for (int p = 0; p < P.world; p++) {
        if (p == P.rank || (P.mode == 1 && p != P.nbr[0] && p != P.nbr[1])) continue;
        const int lo = rg[p].x, len = rg[p].y - rg[p].x;
        const unsigned long long* slot = strip_slot(P.base[P.rank], P, par, p);
        for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < (size_t)P.nf * len; i += (size_t)gridDim.x * 256) {
            const int f = (int)(i / len);
            const size_t j = (size_t)f * P.Kcap + lo + (i - (size_t)f * len);
            const unsigned long long v = __ldcv(&slot[j]);       // (not through a stale L1 line)
            if (v) atomicAdd(&acc[j], v);                        // ranges of different sources may overlap: another thread may own j there
        }
    }
}

// The strip partition's index ranges (host arithmetic only; also exported as spx_slic_strip_plan for tests without a device).
// Seeds of the strip grid (GetChSeedsS): row gy sits at y = gy*S + S/2 + gy*yerr_i; cluster index = gy * xstrips + gx.
static int strip_plan(int w, int h, int S, const int* own_rows, int world, std::vector<int>& lo, std::vector<int>& hi,
                      std::vector<int>& clo, std::vector<int>& chi)
{
    SPX_REQUIRE(w > 0 && h > 0 && S >= 1 && own_rows && world >= 1, SPX_ERR_BADARG, "strip_plan: bad argument");
    const int xstrips = (int)(0.5f + (float)w / (float)S), ystrips = (int)(0.5f + (float)h / (float)S);
    SPX_REQUIRE(xstrips >= 1 && ystrips >= 1, SPX_ERR_BADARG, "strip_plan: region_size leaves no seed");
    const int yerr_i = (int)((float)(h - S * ystrips) / (float)ystrips);
    const int M = STRIP_MARGIN_S * S;
    auto seed_y = [&](int gy) { return std::min(gy * S + S / 2 + gy * yerr_i, h - 1); };
    lo.assign(world, 0); hi.assign(world, 0); clo.assign(world, 0); chi.assign(world, 0);
    for (int q = 0; q < world; q++) {
        const int y0 = own_rows[2 * q], y1 = own_rows[2 * q + 1];
        SPX_REQUIRE(y1 > y0, SPX_ERR_BADARG, "strip_partition: rank %d owns no rows", q);
        int a0 = ystrips, a1 = 0, c0 = ystrips, c1 = 0;
        for (int gy = 0; gy < ystrips; gy++) {
            const int y = seed_y(gy);
            if (y >= y0 - M && y < y1 + M) { a0 = std::min(a0, gy); a1 = std::max(a1, gy + 1); }
            if (y >= y0 && y < y1) { c0 = std::min(c0, gy); c1 = std::max(c1, gy + 1); }
        }
        if (a1 < a0) a0 = a1 = 0;
        if (c1 < c0) c0 = c1 = a0;
        lo[q] = a0 * xstrips; hi[q] = a1 * xstrips; clo[q] = c0 * xstrips; chi[q] = c1 * xstrips;
    }
    for (int q = 0; q + 2 < world; q++)
        SPX_REQUIRE(hi[q] <= lo[q + 2], SPX_ERR_BADARG, "strip_partition: bands are too thin for the neighbour exchange (need more than %d rows per rank)", (2 * STRIP_MARGIN_S + 2) * S);
    return SPX_OK;
}

struct SlicEngine {
    int device = 0;
    cudaStream_t st = nullptr;
    bool own_stream = false;
    SlicGeom g{};
    SlicArrays a{};
    float ruler = 10.f;
    unsigned char* d_img = nullptr;
    unsigned* d_img4 = nullptr;      // tiled path: padded Lab1 words (c0,c1,c2,1)
    int* d_labels = nullptr;
    float* d_dcplane = nullptr;      // SLICO only
    float* d_area_px = nullptr;      // MSLIC only: manifold area of every pixel quad (label pitch)
    unsigned char* d_mask = nullptr;
    int* d_K = nullptr;
    int* d_overflow = nullptr;
    int* h_pinned = nullptr;         // [4]
    int K0 = 0;
    int numlabels = 0;
    int parity = 0;
    bool use_tile = false;
    bool use_graph = true;
    // Centre update: slic_finalize16_kernel (16 threads per centre) has the shortest dependent chain -- best for one frame on an
    // otherwise idle device (4K iterate(10) 0.575 vs 0.649 ms) -- but executes 3.0 M warp-instructions per launch, 7 % of an
    // assignment launch; with several frames in flight on other streams (spx_slic_batch_dev) that costs 5 % of the throughput
    // (14.2 vs 15.0 GP/s), so the batch engines use slic_finalize_warp_kernel (lane = centre, shuffles for the scatter), as does
    // any update over more than 32768 centres (8K frames, gigapixel strips).
    bool lean_update = false;
    long long launches = 0;
    PostWorkspace post;
    std::map<long long, cudaGraphExec_t> graphs;
    std::vector<void*> allocs;
    // row-strip mode, peer exchange
    StripPeers peers{};
    void* xchg = nullptr;            // this rank's exchange buffer (cudaMalloc: exportable)
    size_t xchg_bytes = 0;
    unsigned* d_push_done = nullptr;
    int* d_range = nullptr;
    unsigned long long* d_bytes_sent = nullptr;
    int* d_strip_err = nullptr;      // raised by strip_wait_sum_kernel when a peer never delivered (1 + its rank)
    int xchg_iter = 0;
    bool peers_open = false;
    // strip partition (neighbour exchange): the clusters this rank maintains and the ranges it shares with rank - 1 / rank + 1
    bool part_on = false;
    int part_k_lo = 0, part_k_hi = 0, part_core_lo = 0, part_core_hi = 0;
    int part_nbr[2] = {-1, -1}, part_nlo[2] = {0, 0}, part_nhi[2] = {0, 0};

    template <typename T> int dalloc(T** p, size_t n)
    {
        SPX_CHECK(pool_alloc((void**)p, n * sizeof(T), st));
        allocs.push_back((void*)*p);
        return SPX_OK;
    }
    ~SlicEngine()
    {
        cudaSetDevice(device);
        if (st) cudaStreamSynchronize(st);
        for (auto& kv : graphs) cudaGraphExecDestroy(kv.second);
        for (void* p : allocs) pool_free(p, st);
        post.release();
        pinned_scratch_free(h_pinned);
        strip_peer_close();
        if (own_stream && st) cudaStreamDestroy(st);
    }
    void strip_peer_close()
    {
        if (peers_open)
            for (int p = 0; p < peers.world; p++)
                if (p != peers.rank && peers.base[p]) cudaIpcCloseMemHandle(peers.base[p]);
        peers_open = false;
        if (xchg) cudaFree(xchg);
        xchg = nullptr;
    }
    // Strip partition.  own_rows[2q], own_rows[2q+1] = the rows rank q owns.  Rank q MAINTAINS the clusters whose seed row lies within
    // M = 5S rows of its band (cluster indices are row-major in the seed grid: a contiguous index range).  A centre may drift up to
    // 4S rows from its seed row (checked every update), so its window -- and every pixel carrying its label -- stays within 5S rows of
    // the seed row: exactly the bands of the ranks that maintain it.  Bands taller than 10S + 2 seed pitches make those at most two
    // ADJACENT ranks, which exchange their partial sums over the shared index range; nobody else needs them.
    int strip_partition(const int* own_rows, int world, int rank, int* info)
    {
        SPX_REQUIRE(own_rows && world >= 2 && world <= STRIP_MAX_WORLD && rank >= 0 && rank < world, SPX_ERR_BADARG, "strip_partition: bad argument");
        SPX_REQUIRE(g.alg == SPX_SLIC && !xchg, SPX_ERR_BADARG, "strip_partition: SLIC only, before strip_peer_alloc");
        std::vector<int> lo, hi, clo, chi;
        SPX_CHECK(strip_plan(g.w, g.h, g.S, own_rows, world, lo, hi, clo, chi));
        part_on = true;
        part_k_lo = lo[rank]; part_k_hi = hi[rank]; part_core_lo = clo[rank]; part_core_hi = chi[rank];
        for (int q = 0; q < 2; q++) {
            const int p = q == 0 ? rank - 1 : rank + 1;
            part_nbr[q] = (p >= 0 && p < world) ? p : -1;
            part_nlo[q] = part_nhi[q] = 0;
            if (part_nbr[q] >= 0) { part_nlo[q] = std::max(lo[rank], lo[p]); part_nhi[q] = std::max(part_nlo[q], std::min(hi[rank], hi[p])); }
        }
        if (info) { info[0] = part_k_lo; info[1] = part_k_hi; info[2] = part_core_lo; info[3] = part_core_hi; }
        return SPX_OK;
    }
    int strip_peer_alloc(int world, int rank, void* handle_out, size_t* bytes_out)
    {
        SPX_REQUIRE(world >= 1 && world <= STRIP_MAX_WORLD && rank >= 0 && rank < world && handle_out, SPX_ERR_BADARG,
                    "strip_peer_alloc: need 1 <= world <= %d, 0 <= rank < world", STRIP_MAX_WORLD);
        SPX_REQUIRE(!xchg, SPX_ERR_BADARG, "strip_peer_alloc: called twice");
        peers = StripPeers{};
        peers.world = world; peers.rank = rank; peers.Kcap = g.Kcap; peers.nf = g.C + 3; peers.n6 = (size_t)(g.C + 3) * g.Kcap;
        peers.mode = part_on ? 1 : 0;
        for (int q = 0; q < 2; q++) { peers.nbr[q] = part_nbr[q]; peers.nlo[q] = part_nlo[q]; peers.nhi[q] = part_nhi[q]; }
        xchg_bytes = (size_t)2 * world * peers.n6 * sizeof(unsigned long long) + (size_t)STRIP_MAX_WORLD * sizeof(int) +
                     (size_t)2 * STRIP_MAX_WORLD * sizeof(int2);
        SPX_CUDA(cudaMalloc(&xchg, xchg_bytes));
        SPX_CUDA(cudaMemsetAsync(xchg, 0, xchg_bytes, st));
        SPX_CHECK(dalloc(&d_push_done, 1));
        SPX_CUDA(cudaMemsetAsync(d_push_done, 0, sizeof(unsigned), st));
        SPX_CHECK(dalloc(&d_range, 2));
        const int r0[2] = {0x7fffffff, -1};
        SPX_CUDA(cudaMemcpyAsync(d_range, r0, sizeof(r0), cudaMemcpyHostToDevice, st));
        SPX_CHECK(dalloc(&d_bytes_sent, 1));
        SPX_CUDA(cudaMemsetAsync(d_bytes_sent, 0, sizeof(unsigned long long), st));
        SPX_CHECK(dalloc(&d_strip_err, 1));
        SPX_CUDA(cudaMemsetAsync(d_strip_err, 0, sizeof(int), st));
        SPX_CUDA(cudaStreamSynchronize(st));                      // zeroed before any peer can map and write it
        cudaIpcMemHandle_t hd;
        SPX_CUDA(cudaIpcGetMemHandle(&hd, xchg));
        static_assert(sizeof(hd) == 64, "cudaIpcMemHandle_t is 64 bytes");
        memcpy(handle_out, &hd, sizeof(hd));
        if (bytes_out) *bytes_out = xchg_bytes;
        peers.base[rank] = (unsigned long long*)xchg;
        xchg_iter = 0;
        return SPX_OK;
    }
    // after a stream synchronisation: did an exchange give up waiting for a peer?
    int strip_check_err()
    {
        if (!d_strip_err) return SPX_OK;
        SPX_CUDA(cudaMemcpyAsync(h_pinned + 1, d_strip_err, sizeof(int), cudaMemcpyDeviceToHost, st));
        SPX_CUDA(cudaStreamSynchronize(st));
        SPX_REQUIRE(h_pinned[1] != 100, SPX_ERR_ASSERT, "strip partition: a cluster moved more than 4 * region_size rows away from its seed row; "
                    "the neighbour exchange does not cover that -- rerun with the all-to-all exchange");
        SPX_REQUIRE(h_pinned[1] == 0, SPX_ERR_GPU, "strip exchange: rank %d never delivered its sums (peer gone or exchange skipped)", h_pinned[1] - 1);
        return SPX_OK;
    }
    int strip_peer_open(const void* handles)
    {
        SPX_REQUIRE(xchg && handles && !peers_open, SPX_ERR_BADARG, "strip_peer_open: call strip_peer_alloc first (once)");
        for (int p = 0; p < peers.world; p++) {
            if (p == peers.rank) continue;
            cudaIpcMemHandle_t hd;
            memcpy(&hd, (const char*)handles + (size_t)p * sizeof(hd), sizeof(hd));
            void* ptr = nullptr;
            SPX_CUDA(cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess));
            peers.base[p] = (unsigned long long*)ptr;
        }
        peers_open = true;
        return SPX_OK;
    }
    // the exchange step: deliver the local partial sums to every peer, wait for theirs, add them up
    // neighbour mode: the two halves of the exchange (the shared ranges are a few seed rows of clusters: tens to hundreds of KB)
    int strip_nb_blocks() const
    {
        const int len = std::max(std::max(peers.nhi[0] - peers.nlo[0], peers.nhi[1] - peers.nlo[1]), 1);
        return (int)std::min<size_t>(((size_t)peers.nf * len + 255) / 256, 148);
    }
    int strip_push_neighbours()
    {
        xchg_iter++;
        strip_push_kernel<<<strip_nb_blocks(), 256, 0, st>>>(peers, a.acc, xchg_iter, d_range, d_push_done, d_bytes_sent);
        launches++;
        return SPX_OK;
    }
    int strip_wait_neighbours()
    {
        strip_wait_sum_kernel<<<strip_nb_blocks(), 256, 0, st>>>(peers, a.acc, xchg_iter, d_strip_err);
        launches++;
        return SPX_OK;
    }
    // One whole iteration of a strip object: assign, exchange, update.  SPX_STRIP_SPLIT=1 (partitioned mode) assigns the tile rows
    // next to the cuts FIRST -- only they can hold pixels of the clusters shared with a neighbour (within 2 * 5S rows of the cut)
    // --, pushes their sums and assigns the interior of the band while the neighbours' sums travel.  Measured (16384^2, r2): 2 GPUs
    // 4096^2 1.11 vs 1.19 ms, but 4 GPUs 4.51 vs 4.46 ms and 8 GPUs 2.72 vs 2.59 ms per iterate(10): two more launches per
    // iteration cost more than the hidden flag latency saves, so it is opt-in.
    int strip_iteration()
    {
        if (!(peers_open && peers.mode == 1) || !getenv("SPX_STRIP_SPLIT")) {
            launches++;
            SPX_CHECK(enqueue_assign());
            SPX_CHECK(strip_exchange());
            return strip_update();
        }
        const int nbt = cdiv(2 * STRIP_MARGIN_S * g.S, TILE_H);
        const int top_end = peers.nbr[0] >= 0 ? std::min(g.ty_lo + nbt, g.ty_hi) : g.ty_lo;
        const int bot_beg = peers.nbr[1] >= 0 ? std::max(g.ty_hi - nbt, top_end) : g.ty_hi;
        auto assign_rows = [&](int r0, int r1) -> int {
            if (r1 <= r0) return SPX_OK;
            SlicGeom g2 = g;
            g2.ty_lo = r0; g2.ty_hi = r1;
            launches++;
            return slic_tile_assign(g2, a, d_img, d_img4, d_labels, d_dcplane, d_area_px, d_K, parity, d_overflow, st);
        };
        SPX_CHECK(assign_rows(g.ty_lo, top_end));
        SPX_CHECK(assign_rows(bot_beg, g.ty_hi));
        SPX_CHECK(strip_push_neighbours());
        SPX_CHECK(assign_rows(top_end, bot_beg));
        SPX_CHECK(strip_wait_neighbours());
        SPX_CUDA(cudaGetLastError());
        return strip_update();
    }
    int strip_exchange()
    {
        SPX_REQUIRE(peers_open || peers.world == 1, SPX_ERR_BADARG, "strip_exchange: peers are not mapped");
        if (peers.world == 1) return SPX_OK;
        if (peers.mode == 1) {
            SPX_CHECK(strip_push_neighbours());
            SPX_CHECK(strip_wait_neighbours());
        } else {
            xchg_iter++;
            const int nb = (int)std::min<size_t>((peers.n6 + 255) / 256, 4 * 148);
            strip_range_kernel<<<std::min(cdiv(g.Kcap, 256), 148), 256, 0, st>>>(a.acc + (size_t)(g.C + 2) * g.Kcap, g.Kcap, d_range);
            strip_push_kernel<<<nb, 256, 0, st>>>(peers, a.acc, xchg_iter, d_range, d_push_done, d_bytes_sent);
            strip_wait_sum_kernel<<<nb, 256, 0, st>>>(peers, a.acc, xchg_iter, d_strip_err);
            launches += 3;
        }
        SPX_CUDA(cudaGetLastError());
        return SPX_OK;
    }

    SeedParams seed_params(int* K_out) const
    {
        SeedParams sp{};
        const int S = g.S, w = g.w, h = g.h;
        if (g.alg == SPX_SLICO) {
            sp.hex = 1;
            int rows = 0;
            while (rows * S + S / 2 <= h - 1) rows++;
            auto cnt = [&](int off) { int c = 0; while (c * S + off <= w - 1) c++; return c; };
            sp.ne = cnt(S / 2);
            sp.no = cnt((S / 2) << 1);
            *K_out = (rows / 2) * (sp.ne + sp.no) + (rows & 1) * sp.ne;
        } else {
            int xstrips = (int)(0.5f + (float)w / (float)S);
            int ystrips = (int)(0.5f + (float)h / (float)S);
            if (xstrips < 1 || ystrips < 1) { *K_out = 0; return sp; }
            float xe = (float)(w - S * xstrips) / (float)xstrips;
            float ye = (float)(h - S * ystrips) / (float)ystrips;
            sp.xstrips = xstrips; sp.xerr_i = (int)xe; sp.yerr_i = (int)ye;
            *K_out = xstrips * ystrips;
        }
        return sp;
    }

    int init(int w, int h, int C, int alg, int S, float ruler_)
    {
        SPX_REQUIRE(w > 0 && h > 0 && h <= SPX_MAX_HEIGHT && (long long)w * h < (1ll << 31), SPX_ERR_BADARG,
                    "createSuperpixelSLIC: bad image size %dx%d (need w*h < 2^31 and h <= 65535)", w, h);
        SPX_REQUIRE(C >= 1 && C <= SPX_MAXC, SPX_ERR_BADARG, "createSuperpixelSLIC: 1..4 channels supported, got %d", C);
        SPX_REQUIRE(alg == SPX_SLIC || alg == SPX_SLICO || alg == SPX_MSLIC, SPX_ERR_INTERNAL, "No such algorithm");
        SPX_REQUIRE(S >= 1, SPX_ERR_BADARG, "createSuperpixelSLIC: region_size must be >= 1");
        SPX_REQUIRE(ruler_ > 0.f || alg == SPX_SLICO, SPX_ERR_BADARG, "createSuperpixelSLIC: ruler must be > 0");
        ruler = ruler_;
        g.w = w; g.h = h; g.C = C; g.S = S; g.alg = alg;
        g.ipitch = round_up((size_t)w * C, 16) + 16;
        g.tiles_x = cdiv(w, TILE_W); g.tiles_y = cdiv(h, TILE_H);
        g.lp = g.tiles_x * TILE_W;
        g.p4 = g.tiles_x * TILE_W;
        g.ty_lo = 0; g.ty_hi = g.tiles_y; g.own_y0 = 0; g.own_y1 = h;
        g.nbx = cdiv(w, S); g.nby = cdiv(h, S);
        g.radius = alg == SPX_MSLIC ? 2 : 1;
        if (alg == SPX_SLICO) g.xywt = (float)(S * S);
        else { float q = (float)S / ruler; g.xywt = q * q; }
        int K = 0;
        seed_params(&K);
        SPX_REQUIRE(K >= 1, SPX_ERR_BADARG, "createSuperpixelSLIC: region_size %d leaves no seed in a %dx%d image", S, w, h);
        g.K = K; K0 = K;
        g.Kcap = alg == SPX_MSLIC ? 4 * K : K;
        const size_t kc = (size_t)g.Kcap;
        SPX_CHECK(dalloc(&d_img, g.ipitch * (size_t)(h + 1)));
        SPX_CHECK(dalloc(&d_labels, (size_t)g.lp * g.tiles_y * TILE_H));
        SPX_CHECK(dalloc(&d_mask, (size_t)w * h));
        SPX_CHECK(dalloc(&a.kx, kc)); SPX_CHECK(dalloc(&a.ky, kc)); SPX_CHECK(dalloc(&a.kc, kc * C));
        SPX_CHECK(dalloc(&a.ikx, kc)); SPX_CHECK(dalloc(&a.iky, kc)); SPX_CHECK(dalloc(&a.ioff, kc));
        SPX_CHECK(dalloc(&a.adaptk, kc));
        SPX_CHECK(dalloc(&a.maxc, kc)); SPX_CHECK(dalloc(&a.maxxy, kc));
        SPX_CHECK(dalloc(&a.maxc_acc, kc)); SPX_CHECK(dalloc(&a.maxxy_acc, kc));
        SPX_CHECK(dalloc(&a.acc, kc * (C + 3)));
        SPX_CHECK(dalloc(&a.next, kc));
        SPX_CHECK(dalloc(&a.head[0], (size_t)g.nbx * g.nby)); SPX_CHECK(dalloc(&a.head[1], (size_t)g.nbx * g.nby));
        SPX_CHECK(dalloc(&a.area, kc));
        SPX_CHECK(dalloc(&a.ms_scratch, (size_t)8 + kc / 256 + 8));
        SPX_CHECK(dalloc(&d_K, 4)); SPX_CHECK(dalloc(&d_overflow, 4));
        if (alg == SPX_SLICO) SPX_CHECK(dalloc(&d_dcplane, (size_t)g.lp * g.tiles_y * TILE_H));
        if (alg == SPX_MSLIC) {
            SPX_CHECK(dalloc(&d_area_px, (size_t)g.lp * g.tiles_y * TILE_H));
            SPX_CUDA(cudaMemsetAsync(d_area_px, 0, (size_t)g.lp * g.tiles_y * TILE_H * sizeof(float), st));   // last row / column and padding stay 0
        }
        SPX_CHECK(pinned_scratch_alloc((void**)&h_pinned, 4 * sizeof(int)));
        SPX_CUDA(cudaMemsetAsync(d_img, 0, g.ipitch * (size_t)(h + 1), st));
        use_tile = slic_tile_supported(g) && !getenv("SPX_SLIC_GENERIC");
        g.tile_on = use_tile ? 1 : 0;
        if (use_tile) {
            const size_t nt = (size_t)g.tiles_x * g.tiles_y;
            const size_t n4 = (size_t)g.p4 * g.tiles_y * TILE_H;
            SPX_CHECK(dalloc(&d_img4, n4));
            SPX_CUDA(cudaMemsetAsync(d_img4, 0, n4 * sizeof(unsigned), st));
            for (int p = 0; p < 2; p++) {
                SPX_CHECK(dalloc(&a.tl_count[p], nt));
                SPX_CHECK(dalloc(&a.tl_rec[p], nt * TILE_CAP));
            }
        }
        use_graph = !getenv("SPX_NO_GRAPH");
        return SPX_OK;
    }

    // the part of the constructor that depends on the pixels: seeds, perturbation, first binning
    int seed()
    {
        const size_t kc = (size_t)g.Kcap;
        g.K = K0; numlabels = K0; parity = 0;
        SPX_CUDA(cudaMemsetAsync(d_labels, 0, (size_t)g.lp * g.tiles_y * TILE_H * sizeof(int), st));
        SPX_CUDA(cudaMemsetAsync(a.acc, 0, kc * (g.C + 3) * sizeof(unsigned long long), st));
        SPX_CUDA(cudaMemsetAsync(a.area, 0, kc * sizeof(long long), st));
        SPX_CUDA(cudaMemsetAsync(a.ms_scratch, 0, 4 * sizeof(int), st));
        SPX_CUDA(cudaMemsetAsync(a.head[0], 0xFF, (size_t)g.nbx * g.nby * sizeof(int), st));
        SPX_CUDA(cudaMemsetAsync(a.head[1], 0xFF, (size_t)g.nbx * g.nby * sizeof(int), st));
        SPX_CUDA(cudaMemsetAsync(d_overflow, 0, 4 * sizeof(int), st));
        if (use_tile)
            for (int p = 0; p < 2; p++) SPX_CUDA(cudaMemsetAsync(a.tl_count[p], 0, (size_t)g.tiles_x * g.tiles_y * sizeof(int), st));
        h_pinned[0] = K0;
        SPX_CUDA(cudaMemcpyAsync(d_K, h_pinned, sizeof(int), cudaMemcpyHostToDevice, st));
        int K = 0;
        SeedParams sp = seed_params(&K);
        if (use_tile) { SPX_CHECK(slic_tile_prepare(g, d_img, d_img4, st)); launches++; }
        if (g.alg == SPX_MSLIC) {
            mslic_area_px_kernel<<<dim3(cdiv(g.w, 32), cdiv(g.h, 8)), 256, 0, st>>>(g, d_img, d_area_px, (float)g.S / ruler);
            launches++;
        }
        slic_seed_kernel<<<cdiv(K0, 128), 128, 0, st>>>(g, a, d_img, sp);
        slic_bin_kernel<<<cdiv(K0, BIN_CTA), BIN_CTA, 0, st>>>(g, a, d_K, 0);
        launches += 2;
        SPX_CUDA(cudaGetLastError());
        return SPX_OK;
    }

    int upload(const uint8_t* img, size_t stride, bool from_device)
    {
        SPX_REQUIRE(img && stride >= (size_t)g.w * g.C, SPX_ERR_BADARG, "createSuperpixelSLIC: bad image pointer/stride");
        if (from_device) SPX_CUDA(cudaMemcpy2DAsync(d_img, g.ipitch, img, stride, (size_t)g.w * g.C, g.h, cudaMemcpyDeviceToDevice, st));
        else SPX_CHECK(upload_2d(d_img, g.ipitch, img, stride, (size_t)g.w * g.C, g.h, st));      // (pageable cv::Mat: staged on several lanes)
        return SPX_OK;
    }

    // ---- row-strip mode (one image over several GPUs): explicit steps instead of iterate()
    int strip_rebin()
    {
        // the seeds have been assembled across ranks: rebuild the bins / tile lists of parity 0 from them
        SPX_CUDA(cudaMemsetAsync(a.head[0], 0xFF, (size_t)g.nbx * g.nby * sizeof(int), st));
        if (use_tile) SPX_CUDA(cudaMemsetAsync(a.tl_count[0], 0, (size_t)g.tiles_x * g.tiles_y * sizeof(int), st));
        slic_bin_kernel<<<cdiv(K0, BIN_CTA), BIN_CTA, 0, st>>>(g, a, d_K, 0);
        launches++;
        SPX_CUDA(cudaGetLastError());
        return SPX_OK;
    }
    // SLIC / SLICO centre update over the clusters [fr.k_lo, fr.k_hi): the latency-lean kernel for a single frame with few
    // centres, the instruction-lean one for many centres or objects in throughput mode (see the kernels)
    int launch_centre_update(const FinRange& fr)
    {
        const int n = max(fr.k_hi - fr.k_lo, 1);
        if (lean_update || n > 32768) {
            SPX_CUDA(launch_pdl(slic_finalize_warp_kernel, dim3(cdiv(n, FIN16_CTA)), dim3(FIN16_CTA), st, g, a, parity, fr));
        } else {
            const int th16 = max(max(16 * n, g.nbx * g.nby), g.tiles_x * g.tiles_y);
            SPX_CUDA(launch_pdl(slic_finalize16_kernel, dim3(cdiv(th16, FIN16_CTA)), dim3(FIN16_CTA), st, g, a, parity, fr));
        }
        return SPX_OK;
    }
    int strip_update()
    {
        FinRange fr{0, g.K, 0, 1, 0, d_strip_err};
        if (part_on) {
            int K = 0;
            const SeedParams sp = seed_params(&K);
            // test hook: SPX_STRIP_FORCE_DRIFT=i raises the drift flag in the i-th update, so that the host's rerun path gets exercised
            static const int force_at = getenv("SPX_STRIP_FORCE_DRIFT") ? atoi(getenv("SPX_STRIP_FORCE_DRIFT")) : 0;
            fr = FinRange{part_k_lo, part_k_hi, (force_at > 0 && xchg_iter == force_at) ? 2 : 1, sp.xstrips, sp.yerr_i, d_strip_err};
        }
        SPX_CHECK(launch_centre_update(fr));
        launches++;
        parity ^= 1;
        SPX_CUDA(cudaGetLastError());
        return SPX_OK;
    }

    // one iteration: assignment (+ events around it when asked) and the centre update
    int enqueue_iteration(cudaEvent_t e0 = nullptr, cudaEvent_t e1 = nullptr)
    {
        const int threads_fin = max(max(g.Kcap, g.nbx * g.nby), g.tiles_x * g.tiles_y);
        if (e0) SPX_CUDA(cudaEventRecord(e0, st));
        if (use_tile) {
            SPX_CHECK(slic_tile_assign(g, a, d_img, d_img4, d_labels, d_dcplane, d_area_px, d_K, parity, d_overflow, st));
        } else {
            dim3 grid(cdiv(g.w, 32), cdiv(g.h, 8));
            if (g.alg == SPX_SLIC) slic_assign_generic_kernel<100><<<grid, 256, 0, st>>>(g, a, d_img, d_labels, d_dcplane, parity, 0, 0, g.w, g.h);
            else if (g.alg == SPX_SLICO) slic_assign_generic_kernel<101><<<grid, 256, 0, st>>>(g, a, d_img, d_labels, d_dcplane, parity, 0, 0, g.w, g.h);
            else slic_assign_generic_kernel<102><<<grid, 256, 0, st>>>(g, a, d_img, d_labels, d_dcplane, parity, 0, 0, g.w, g.h);
        }
        launches++;
        if (e1) SPX_CUDA(cudaEventRecord(e1, st));
        const int do_bin = g.alg == SPX_MSLIC ? 0 : 1;
        if (do_bin && g.C == 3) {
            SPX_CHECK(launch_centre_update(FinRange{0, g.K, 0, 1, 0, nullptr}));
        } else
            SPX_CUDA(launch_pdl(slic_finalize_kernel, dim3(cdiv(threads_fin, BIN_CTA)), dim3(BIN_CTA), st, g, a, (const int*)d_K, parity, do_bin));
        launches++;
        if (g.alg == SPX_MSLIC) {
            // the tiled assignment kernel already summed the manifold areas per cluster; the generic path does it here
            if (!use_tile) { mslic_area_sum_kernel<<<dim3(cdiv(g.w, 32), cdiv(g.h, 8)), 256, 0, st>>>(g, a, d_area_px, d_labels); launches++; }
            const int nbk = cdiv(g.Kcap, MA);
            mslic_total_kernel<<<nbk, MA, 0, st>>>(a, d_K);
            mslic_adaptk_kernel<<<nbk, MA, 0, st>>>(a, d_K);
            mslic_split_kernel<<<nbk, MA, 0, st>>>(g, a, d_img, d_K, K0);
            mslic_finish_kernel<<<nbk, MA, 0, st>>>(g, a, d_K);
            mslic_scratch_clear_kernel<<<1, 32, 0, st>>>(a);
            launches += 4;                     // + the two counted below
            slic_bin_kernel<<<cdiv(g.Kcap, BIN_CTA), BIN_CTA, 0, st>>>(g, a, d_K, parity ^ 1);
            launches += 2;
        }
        parity ^= 1;
        return SPX_OK;
    }
    int enqueue_begin()
    {
        if (g.alg == SPX_SLICO) {
            size_t np = (size_t)g.lp * g.tiles_y * TILE_H;       // the plane has the label pitch
            slico_begin_kernel<<<(unsigned)((max(np, (size_t)g.K) + 255) / 256), 256, 0, st>>>(g, a, d_dcplane, np);
            launches++;
        }
        return SPX_OK;
    }
    int enqueue_iterations(int n)
    {
        SPX_CHECK(enqueue_begin());
        for (int it = 0; it < n; it++) SPX_CHECK(enqueue_iteration());
        SPX_CUDA(cudaGetLastError());
        return SPX_OK;
    }

    int enqueue_assign()
    {
        if (use_tile) return slic_tile_assign(g, a, d_img, d_img4, d_labels, d_dcplane, d_area_px, d_K, parity, d_overflow, st);
        dim3 grid(cdiv(g.w, 32), cdiv(g.h, 8));
        if (g.alg == SPX_SLIC) slic_assign_generic_kernel<100><<<grid, 256, 0, st>>>(g, a, d_img, d_labels, d_dcplane, parity, 0, 0, g.w, g.h);
        else if (g.alg == SPX_SLICO) slic_assign_generic_kernel<101><<<grid, 256, 0, st>>>(g, a, d_img, d_labels, d_dcplane, parity, 0, 0, g.w, g.h);
        else slic_assign_generic_kernel<102><<<grid, 256, 0, st>>>(g, a, d_img, d_labels, d_dcplane, parity, 0, 0, g.w, g.h);
        SPX_CUDA(cudaGetLastError());
        return SPX_OK;
    }
    int time_assign(int reps, int flush, float* avg_ms)
    {
        void* fl = nullptr;
        const size_t flbytes = 256u << 20;
        if (flush) SPX_CHECK(pool_alloc(&fl, flbytes, st));
        cudaEvent_t e0, e1;
        SPX_CUDA(cudaEventCreate(&e0)); SPX_CUDA(cudaEventCreate(&e1));
        double tot = 0;
        int rc = SPX_OK;
        for (int r = 0; r < reps + 1 && rc == SPX_OK; r++) {      // launch 0 is a warm-up
            if (fl) cudaMemsetAsync(fl, r, flbytes, st);
            cudaEventRecord(e0, st);
            rc = enqueue_assign();
            cudaEventRecord(e1, st);
            if (cudaEventSynchronize(e1) != cudaSuccess) { set_error("time_assign: kernel failed"); rc = SPX_ERR_GPU; }
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            if (r > 0) tot += ms;
            launches++;
        }
        // the accumulators were fed `reps+1` times: clear them (centres / labels are unchanged)
        cudaMemsetAsync(a.acc, 0, (size_t)g.Kcap * (g.C + 3) * sizeof(unsigned long long), st);
        cudaMemsetAsync(a.area, 0, (size_t)g.Kcap * sizeof(long long), st);
        cudaStreamSynchronize(st);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        if (fl) pool_free(fl, st);
        *avg_ms = (float)(tot / (reps > 0 ? reps : 1));
        return rc;
    }
    // The assignment launches of a whole iterate(iters) from the seeds, each between its own pair of events on the
    // object's stream (no graph): ms[i] = duration of the i-th assignment launch averaged over `reps` runs -- what
    // bench.py's roofline is quoted on.  Leaves the object in the state iterate(iters) on a fresh object gives.
    int time_iterations(int iters, int reps, int flush, float* ms)
    {
        void* fl = nullptr;
        const size_t flbytes = 256u << 20;
        if (flush) SPX_CHECK(pool_alloc(&fl, flbytes, st));
        std::vector<cudaEvent_t> ev(2 * (size_t)iters);
        for (auto& e : ev) SPX_CUDA(cudaEventCreate(&e));
        std::vector<double> tot((size_t)iters, 0.0);
        int rc = SPX_OK;
        for (int r = 0; r < reps + 1 && rc == SPX_OK; r++) {      // run 0 is a warm-up
            rc = seed();
            if (rc == SPX_OK) rc = enqueue_begin();
            for (int it = 0; it < iters && rc == SPX_OK; it++) {
                if (fl) cudaMemsetAsync(fl, r + it, flbytes, st);
                rc = enqueue_iteration(ev[2 * it], ev[2 * it + 1]);
            }
            if (rc == SPX_OK && cudaStreamSynchronize(st) != cudaSuccess) { set_error("time_iterations: kernel failed"); rc = SPX_ERR_GPU; }
            for (int it = 0; it < iters && rc == SPX_OK && r > 0; it++) {
                float t = 0;
                cudaEventElapsedTime(&t, ev[2 * it], ev[2 * it + 1]);
                tot[it] += t;
            }
        }
        for (auto& e : ev) cudaEventDestroy(e);
        if (fl) pool_free(fl, st);
        for (int it = 0; it < iters; it++) ms[it] = (float)(tot[it] / (reps > 0 ? reps : 1));
        if (rc == SPX_OK && g.alg == SPX_MSLIC) {
            SPX_CUDA(cudaMemcpyAsync(h_pinned, d_K, sizeof(int), cudaMemcpyDeviceToHost, st));
            SPX_CUDA(cudaStreamSynchronize(st));
            g.K = h_pinned[0];
        }
        numlabels = g.K;
        return rc;
    }

    int iterate(int n)
    {
        SPX_REQUIRE(n >= 0, SPX_ERR_BADARG, "iterate: num_iterations must be >= 0");
        if (n == 0) return SPX_OK;
        long long key = ((long long)n << 1) | parity;
        // A graph pays off from the second use on: capturing + instantiating it costs more than launching the ~20 kernels once,
        // and the GUI makes a new object for every COMPUTE click.  The first iterate() with a given (n, parity) launches directly.
        const bool first_use = use_graph && graphs.find(key) == graphs.end() && !seen_keys.count(key);
        if (first_use) seen_keys.insert(key);
        if (!use_graph || first_use) { SPX_CHECK(enqueue_iterations(n)); }
        else {
            auto it = graphs.find(key);
            if (it == graphs.end()) {
                cudaGraph_t graph = nullptr;
                long long l0 = launches;
                int p0 = parity;
                SPX_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
                int rc = enqueue_iterations(n);
                cudaError_t e = cudaStreamEndCapture(st, &graph);
                if (rc != SPX_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
                SPX_CUDA(e);
                cudaGraphExec_t exec = nullptr;
                SPX_CUDA(cudaGraphInstantiate(&exec, graph, 0));
                cudaGraphDestroy(graph);
                graphs[key] = exec;
                graph_launches[key] = launches - l0;
                launches = l0; parity = p0;
                it = graphs.find(key);
            }
            SPX_CUDA(cudaGraphLaunch(it->second, st));
            launches += graph_launches[key];
            parity ^= (n & 1);
        }
        if (g.alg == SPX_MSLIC) {
            SPX_CUDA(cudaMemcpyAsync(h_pinned, d_K, sizeof(int), cudaMemcpyDeviceToHost, st));
            SPX_CUDA(cudaStreamSynchronize(st));
            g.K = h_pinned[0];
        }
        numlabels = g.K;
        return SPX_OK;
    }
    std::map<long long, long long> graph_launches;
    std::set<long long> seen_keys;

    int ensure_post()
    {
        return post.alloc(g.w, g.h, st);
    }
    int enforce(int min_element_size)
    {
        SPX_CHECK(ensure_post());
        long long l0 = post.launches;
        int k = numlabels;
        SPX_CHECK(enforce_connectivity_dev(post, d_labels, g.lp, numlabels, min_element_size, st, &k));
        launches += post.launches - l0;
        numlabels = k;
        return SPX_OK;
    }
    int contour(unsigned char* d_dst, size_t pitch, int thick_line)
    {
        SPX_CHECK(ensure_post());
        long long l0 = post.launches;
        SPX_CHECK(contour_mask_dev(post, d_labels, g.lp, thick_line, 0, d_dst, pitch, st));
        launches += post.launches - l0;
        return SPX_OK;
    }
};

static int make_engine(int w, int h, int C, int alg, int S, float ruler, cudaStream_t user_stream, SlicEngine** out)
{
    SlicEngine* e = new SlicEngine();
    cudaGetDevice(&e->device);
    if (user_stream) { e->st = user_stream; e->own_stream = false; }
    else {
        cudaError_t ce = cudaStreamCreateWithFlags(&e->st, cudaStreamNonBlocking);
        if (ce != cudaSuccess) { set_error("cudaStreamCreate: %s", cudaGetErrorString(ce)); delete e; return SPX_ERR_GPU; }
        e->own_stream = true;
    }
    int rc = e->init(w, h, C, alg, S, ruler);
    if (rc != SPX_OK) { delete e; return rc; }
    *out = e;
    return SPX_OK;
}

}  // namespace spx

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
using spx::SlicEngine;
struct spx_slic { SlicEngine* e; };

#define SLIC_ENTER(h)                                                                     \
    SPX_REQUIRE((h) && (h)->e, SPX_ERR_BADARG, "NULL SuperpixelSLIC handle");             \
    SlicEngine* e = (h)->e;                                                               \
    SPX_CUDA(cudaSetDevice(e->device))

static int slic_create_common(const void* image, size_t stride, int w, int h, int C, int alg, int S, float ruler,
                              cudaStream_t stream, bool from_device, spx_slic_t** out, int cvt_mode = -1)
{
    using namespace spx;
    SPX_REQUIRE(out, SPX_ERR_BADARG, "createSuperpixelSLIC: out is NULL");
    *out = nullptr;
    SPX_REQUIRE(image, SPX_ERR_BADARG, "createSuperpixelSLIC: image is NULL");
    SPX_REQUIRE(cvt_mode < 0 || C == 3, SPX_ERR_BADARG, "createSuperpixelSLIC: the fused colour conversion needs a 3-channel BGR image");
    SlicEngine* e = nullptr;
    static const bool trace = getenv("SPX_TRACE") != nullptr;          // host-side phases of the constructor, for tools/click_stages.py
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = trace ? now() : 0;
    SPX_CHECK(make_engine(w, h, C, alg, S, ruler, stream, &e));
    const double t1 = trace ? now() : 0;
    int rc = e->upload((const uint8_t*)image, stride, from_device);
    const double t2 = trace ? now() : 0;
    // cvtColor(BGR2Lab / BGR2HSV) of mainwindow.cpp:2096-2097 on the freshly uploaded image, in place: no extra host round trip
    if (rc == SPX_OK && cvt_mode >= 0) { rc = cvt8_dev(cvt_mode, e->d_img, e->g.ipitch, e->d_img, e->g.ipitch, w, h, e->st); e->launches++; }
    if (rc == SPX_OK) rc = e->seed();
    if (trace) fprintf(stderr, "[spx] create %dx%d: engine (stream, allocations, memsets) %.3f ms, upload %.3f ms, seeding enqueued %.3f ms\n", w, h,
                       t1 - t0, t2 - t1, now() - t2);
    if (rc != SPX_OK) { delete e; return rc; }
    *out = new spx_slic{e};
    return SPX_OK;
}

extern "C" int spx_slic_create(const uint8_t* image, size_t stride, int w, int h, int C, int alg, int S, float ruler,
                               int device, spx_slic_t** out) try
{
    SPX_CHECK(spx::require_device(device));
    return slic_create_common(image, stride, w, h, C, alg, S, ruler, nullptr, false, out);
} SPX_API_CATCH
extern "C" int spx_slic_create_bgr(const uint8_t* bgr, size_t stride, int w, int h, int colour_code, int alg, int S, float ruler,
                                   int device, spx_slic_t** out) try
{
    SPX_REQUIRE(colour_code == 44 || colour_code == 40, SPX_ERR_BADARG, "createSuperpixelSLIC: colour code must be COLOR_BGR2Lab (44) or COLOR_BGR2HSV (40)");
    SPX_CHECK(spx::require_device(device));
    SPX_CUDA(cudaSetDevice(device));
    return slic_create_common(bgr, stride, w, h, 3, alg, S, ruler, nullptr, false, out, colour_code == 44 ? 0 : 1);
} SPX_API_CATCH
extern "C" int spx_slic_create_dev(const void* d_image, size_t stride, int w, int h, int C, int alg, int S, float ruler,
                                   void* stream, spx_slic_t** out) try
{
    int dev = 0;
    SPX_CUDA(cudaGetDevice(&dev));
    SPX_CHECK(spx::require_device(dev));
    return slic_create_common(d_image, stride, w, h, C, alg, S, ruler, (cudaStream_t)stream, true, out);
} SPX_API_CATCH
// ---- row-strip mode ------------------------------------------------------------------------------------------
extern "C" int spx_slic_create_strip(const uint8_t* rows, size_t stride, int w, int h, int C, int alg, int S, float ruler,
                                     int row0, int row1, int own_row0, int own_row1, void* stream, spx_slic_t** out) try
{
    using namespace spx;
    SPX_REQUIRE(out, SPX_ERR_BADARG, "create_strip: out is NULL");
    *out = nullptr;
    SPX_REQUIRE(rows && alg == SPX_SLIC && C == 3, SPX_ERR_BADARG, "create_strip: 3-channel SLIC only");
    SPX_REQUIRE(0 <= row0 && row0 <= own_row0 && own_row0 < own_row1 && own_row1 <= row1 && row1 <= h, SPX_ERR_BADARG,
                "create_strip: need 0 <= row0 <= own_row0 < own_row1 <= row1 <= height");
    SPX_REQUIRE((own_row0 % TILE_H == 0) && (own_row1 % TILE_H == 0 || own_row1 == h), SPX_ERR_BADARG,
                "create_strip: owned rows must start and end on multiples of %d (or at the image end)", TILE_H);
    SPX_REQUIRE((own_row0 == 0 || own_row0 - row0 >= 2) && (own_row1 == h || row1 - own_row1 >= 2), SPX_ERR_BADARG,
                "create_strip: two halo rows are needed on every inner side (seed perturbation reads them)");
    int dev = 0;
    SPX_CUDA(cudaGetDevice(&dev));
    SPX_CHECK(require_device(dev));
    SlicEngine* e = nullptr;
    SPX_CHECK(make_engine(w, h, C, alg, S, ruler, (cudaStream_t)stream, &e));
    int rc = SPX_OK;
    if (!e->use_tile) { set_error("create_strip: configuration outside the tiled path (region_size >= 7 required)"); rc = SPX_ERR_BADARG; }
    if (rc == SPX_OK) {
        e->g.ty_lo = own_row0 / TILE_H; e->g.ty_hi = cdiv(own_row1, TILE_H);
        e->g.own_y0 = own_row0; e->g.own_y1 = own_row1;
        e->use_graph = false;
        cudaError_t ce = cudaMemcpy2DAsync(e->d_img + (size_t)row0 * e->g.ipitch, e->g.ipitch, rows, stride, (size_t)w * C, row1 - row0,
                                           cudaMemcpyHostToDevice, e->st);
        if (ce != cudaSuccess) { set_error("create_strip: upload failed: %s", cudaGetErrorString(ce)); rc = SPX_ERR_GPU; }
    }
    if (rc == SPX_OK) rc = e->seed();
    if (rc != SPX_OK) { delete e; return rc; }
    *out = new spx_slic{e};
    return SPX_OK;
} SPX_API_CATCH
extern "C" int spx_slic_device_state(spx_slic_t* h, spx_slic_state_t* out) try
{
    SLIC_ENTER(h);
    SPX_REQUIRE(out, SPX_ERR_BADARG, "device_state: out is NULL");
    out->kx = e->a.kx; out->ky = e->a.ky; out->kc = e->a.kc; out->ikx = e->a.ikx; out->iky = e->a.iky; out->acc = e->a.acc;
    out->K = e->g.K; out->Kcap = e->g.Kcap; out->channels = e->g.C;
    out->labels = e->d_labels; out->label_pitch = e->g.lp;
    return SPX_OK;
} SPX_API_CATCH
extern "C" int spx_slic_strip_rebin(spx_slic_t* h) try
{
    SLIC_ENTER(h);
    return e->strip_rebin();
} SPX_API_CATCH
extern "C" int spx_slic_strip_assign(spx_slic_t* h) try
{
    SLIC_ENTER(h);
    e->launches++;
    return e->enqueue_assign();
} SPX_API_CATCH
extern "C" int spx_slic_strip_update(spx_slic_t* h) try
{
    SLIC_ENTER(h);
    return e->strip_update();
} SPX_API_CATCH
extern "C" int spx_slic_strip_plan(int w, int h, int region_size, const int* own_rows, int world, int* plan_out) try
{
    SPX_REQUIRE(plan_out, SPX_ERR_BADARG, "strip_plan: plan_out is NULL");
    std::vector<int> lo, hi, clo, chi;
    SPX_CHECK(spx::strip_plan(w, h, region_size, own_rows, world, lo, hi, clo, chi));
    for (int q = 0; q < world; q++) { plan_out[4 * q] = lo[q]; plan_out[4 * q + 1] = hi[q]; plan_out[4 * q + 2] = clo[q]; plan_out[4 * q + 3] = chi[q]; }
    return SPX_OK;
} SPX_API_CATCH
extern "C" int spx_slic_strip_partition(spx_slic_t* h, const int* own_rows, int world, int rank, int* info_out) try
{
    SLIC_ENTER(h);
    return e->strip_partition(own_rows, world, rank, info_out);
} SPX_API_CATCH
extern "C" int spx_slic_strip_peer_alloc(spx_slic_t* h, int world, int rank, void* ipc_handle_out, size_t* bytes_out) try
{
    SLIC_ENTER(h);
    return e->strip_peer_alloc(world, rank, ipc_handle_out, bytes_out);
} SPX_API_CATCH
extern "C" int spx_slic_strip_peer_open(spx_slic_t* h, const void* ipc_handles) try
{
    SLIC_ENTER(h);
    return e->strip_peer_open(ipc_handles);
} SPX_API_CATCH
extern "C" int spx_slic_strip_exchange(spx_slic_t* h) try
{
    SLIC_ENTER(h);
    return e->strip_exchange();
} SPX_API_CATCH
extern "C" int spx_slic_strip_iteration(spx_slic_t* h) try
{
    SLIC_ENTER(h);
    return e->strip_iteration();
} SPX_API_CATCH
extern "C" int spx_slic_strip_exchanged_bytes(spx_slic_t* h, unsigned long long* out) try
{
    SLIC_ENTER(h);
    SPX_REQUIRE(out, SPX_ERR_BADARG, "strip_exchanged_bytes: out is NULL");
    *out = 0;
    if (!e->d_bytes_sent) return SPX_OK;
    SPX_CUDA(cudaMemcpyAsync(out, e->d_bytes_sent, sizeof(*out), cudaMemcpyDeviceToHost, e->st));
    SPX_CUDA(cudaStreamSynchronize(e->st));
    return e->strip_check_err();
} SPX_API_CATCH
extern "C" int spx_slic_get_label_rows(spx_slic_t* h, int32_t* dst, size_t dst_stride, int row0, int row1) try
{
    SLIC_ENTER(h);
    SPX_REQUIRE(dst && dst_stride >= (size_t)e->g.w * 4 && 0 <= row0 && row0 <= row1 && row1 <= e->g.h, SPX_ERR_BADARG, "getLabelRows: bad argument");
    if (row1 > row0)
        SPX_CUDA(cudaMemcpy2DAsync(dst, dst_stride, e->d_labels + (size_t)row0 * e->g.lp, (size_t)e->g.lp * 4, (size_t)e->g.w * 4, row1 - row0,
                                   cudaMemcpyDeviceToHost, e->st));
    SPX_CUDA(cudaStreamSynchronize(e->st));
    return e->strip_check_err();
} SPX_API_CATCH

extern "C" int spx_slic_reset(spx_slic_t* h, const uint8_t* image, size_t stride) try
{
    SLIC_ENTER(h);
    SPX_CHECK(e->upload(image, stride, false));
    return e->seed();
} SPX_API_CATCH
extern "C" int spx_slic_reset_dev(spx_slic_t* h, const void* d_image, size_t stride) try
{
    SLIC_ENTER(h);
    SPX_CHECK(e->upload((const uint8_t*)d_image, stride, true));
    return e->seed();
} SPX_API_CATCH
extern "C" int spx_slic_iterate(spx_slic_t* h, int n) try
{
    SLIC_ENTER(h);
    return e->iterate(n);
} SPX_API_CATCH
extern "C" int spx_slic_enforce_label_connectivity(spx_slic_t* h, int min_element_size) try
{
    SLIC_ENTER(h);
    return e->enforce(min_element_size);
} SPX_API_CATCH
extern "C" int spx_slic_get_number_of_superpixels(spx_slic_t* h, int* out) try
{
    SLIC_ENTER(h);
    SPX_REQUIRE(out, SPX_ERR_BADARG, "getNumberOfSuperpixels: out is NULL");
    SPX_CUDA(cudaStreamSynchronize(e->st));
    *out = e->numlabels;
    return SPX_OK;
} SPX_API_CATCH
extern "C" int spx_slic_get_labels(spx_slic_t* h, int32_t* dst, size_t dst_stride) try
{
    SLIC_ENTER(h);
    SPX_REQUIRE(dst && dst_stride >= (size_t)e->g.w * 4, SPX_ERR_BADARG, "getLabels: bad destination");
    SPX_CHECK(spx::download_2d(dst, dst_stride, e->d_labels, (size_t)e->g.lp * 4, (size_t)e->g.w * 4, e->g.h, e->st));
    SPX_CUDA(cudaStreamSynchronize(e->st));
    return SPX_OK;
} SPX_API_CATCH
extern "C" int spx_slic_get_labels_dev(spx_slic_t* h, void* d_dst, size_t dst_stride) try
{
    SLIC_ENTER(h);
    SPX_REQUIRE(d_dst && dst_stride >= (size_t)e->g.w * 4, SPX_ERR_BADARG, "getLabels: bad destination");
    SPX_CUDA(cudaMemcpy2DAsync(d_dst, dst_stride, e->d_labels, (size_t)e->g.lp * 4, (size_t)e->g.w * 4, e->g.h, cudaMemcpyDeviceToDevice, e->st));
    return SPX_OK;
} SPX_API_CATCH
extern "C" int spx_slic_get_label_contour_mask(spx_slic_t* h, uint8_t* dst, size_t dst_stride, int thick_line) try
{
    SLIC_ENTER(h);
    SPX_REQUIRE(dst && dst_stride >= (size_t)e->g.w, SPX_ERR_BADARG, "getLabelContourMask: bad destination");
    SPX_CHECK(e->contour(e->d_mask, e->g.w, thick_line));
    SPX_CHECK(spx::download_2d(dst, dst_stride, e->d_mask, e->g.w, e->g.w, e->g.h, e->st));
    SPX_CUDA(cudaStreamSynchronize(e->st));
    return SPX_OK;
} SPX_API_CATCH
extern "C" int spx_slic_get_label_contour_mask_dev(spx_slic_t* h, void* d_dst, size_t dst_stride, int thick_line) try
{
    SLIC_ENTER(h);
    SPX_REQUIRE(d_dst && dst_stride >= (size_t)e->g.w, SPX_ERR_BADARG, "getLabelContourMask: bad destination");
    return e->contour((unsigned char*)d_dst, dst_stride, thick_line);
} SPX_API_CATCH
extern "C" int spx_slic_set_labels(spx_slic_t* h, const int32_t* src, size_t src_stride, int numlabels) try
{
    SLIC_ENTER(h);
    SPX_REQUIRE(src && src_stride >= (size_t)e->g.w * 4, SPX_ERR_BADARG, "setLabels: bad source");
    SPX_CUDA(cudaMemcpy2DAsync(e->d_labels, (size_t)e->g.lp * 4, src, src_stride, (size_t)e->g.w * 4, e->g.h, cudaMemcpyHostToDevice, e->st));
    SPX_CUDA(cudaStreamSynchronize(e->st));
    if (numlabels > 0) e->numlabels = numlabels;
    return SPX_OK;
} SPX_API_CATCH
extern "C" int spx_slic_get_centres(spx_slic_t* h, float* dst, int rows) try
{
    SLIC_ENTER(h);
    const int K = e->g.K, C = e->g.C, Kcap = e->g.Kcap;
    SPX_REQUIRE(dst && rows >= K, SPX_ERR_BADARG, "getCentres: need room for %d rows", K);
    std::vector<float> tmp((size_t)Kcap * (C + 2));
    SPX_CUDA(cudaStreamSynchronize(e->st));
    SPX_CUDA(cudaMemcpy(tmp.data(), e->a.kc, sizeof(float) * (size_t)Kcap * C, cudaMemcpyDeviceToHost));
    SPX_CUDA(cudaMemcpy(tmp.data() + (size_t)Kcap * C, e->a.kx, sizeof(float) * Kcap, cudaMemcpyDeviceToHost));
    SPX_CUDA(cudaMemcpy(tmp.data() + (size_t)Kcap * (C + 1), e->a.ky, sizeof(float) * Kcap, cudaMemcpyDeviceToHost));
    for (int k = 0; k < K; k++)
        for (int f = 0; f < C + 2; f++) dst[(size_t)k * (C + 2) + f] = tmp[(size_t)f * Kcap + k];
    return SPX_OK;
} SPX_API_CATCH
extern "C" int spx_slic_synchronize(spx_slic_t* h) try
{
    SLIC_ENTER(h);
    SPX_CUDA(cudaStreamSynchronize(e->st));
    return SPX_OK;
} SPX_API_CATCH
extern "C" int spx_slic_set_throughput_mode(spx_slic_t* h, int on) try
{
    SLIC_ENTER(h);
    if (e->lean_update != (on != 0)) {
        SPX_CUDA(cudaStreamSynchronize(e->st));
        for (auto& kv : e->graphs) cudaGraphExecDestroy(kv.second);   // captured with the other centre-update kernel
        e->graphs.clear();
        e->graph_launches.clear();
        e->lean_update = on != 0;
    }
    return SPX_OK;
} SPX_API_CATCH
extern "C" int spx_slic_launch_count(spx_slic_t* h, long long* out) try
{
    SPX_REQUIRE(h && h->e && out, SPX_ERR_BADARG, "launch_count: bad argument");
    *out = h->e->launches;
    return SPX_OK;
} SPX_API_CATCH
extern "C" int spx_slic_time_assign(spx_slic_t* h, int reps, int flush_l2, float* avg_ms) try
{
    SLIC_ENTER(h);
    SPX_REQUIRE(avg_ms && reps >= 1, SPX_ERR_BADARG, "time_assign: bad argument");
    return e->time_assign(reps, flush_l2, avg_ms);
} SPX_API_CATCH
extern "C" int spx_slic_time_iterations(spx_slic_t* h, int iters, int reps, int flush_l2, float* assign_ms) try
{
    SLIC_ENTER(h);
    SPX_REQUIRE(assign_ms && reps >= 1 && iters >= 1, SPX_ERR_BADARG, "time_iterations: bad argument");
    SPX_REQUIRE(e->g.ty_lo == 0 && e->g.ty_hi == e->g.tiles_y, SPX_ERR_BADARG, "time_iterations: not for strip objects");
    return e->time_iterations(iters, reps, flush_l2, assign_ms);
} SPX_API_CATCH
extern "C" int spx_slic_destroy(spx_slic_t* h) try
{
    if (!h) return SPX_OK;
    delete h->e;
    delete h;
    return SPX_OK;
} SPX_API_CATCH
