// seeds.cu -- SuperpixelSEEDS engine and its C ABI (replaces createSuperpixelSEEDS / iterate / getLabels /
// getNumberOfSuperpixels / getLabelContourMask as called from mainwindow.cpp:2123-2127).
//
// PARITY UNPINNED upstream (no seeds.cpp, no ximgproc binary; only the grid sizing is pinned by the reference's
// screenshot: 3888x5184 / 1000 / 4 levels -> 972).  Upstream's hill climbing is Gauss-Seidel in raster order; the GPU
// form keeps its decision rules (conditional histogram intersection for blocks, posterior x neighbourhood prior for
// pixels, checkSplit_* connectivity guards) and replaces the visiting order by a PHASE-PARALLEL one: a sweep is cut
// into 8 phases by (x mod 4, y mod 2) [(y mod 4, x mod 2) for vertical sweeps]; a `decide` kernel takes all
// decisions of a phase against the state at the start of the phase, an `apply` kernel performs the moves
// (integer atomics on histograms / sizes: exact and order free).  Moves of one phase are >= 4 columns or >= 2 rows
// apart, so no decision reads a block / pixel another one moves.  the test oracle under oracle/ (SEEDS section) states the same order
// (order 1) and the results match it bit for bit; the relation to upstream's order is statistical (tests/).
//
// Data: bin index per pixel (u16), integer histograms per level [block][bin], block sizes T, current superpixel of
// every block per level (`parent`), blocks per superpixel (`nrp`), int32 label map.
#include "common.cuh"
#include <cmath>
#include <cstdlib>
#include <map>
#include <set>
#include <vector>

namespace spx {

constexpr int SEEDS_MAXLEV = 24;
constexpr float SEEDS_REQ_CONF = 0.1f;

struct SeedsDev {
    int w, h, C, bins, prior, nlev, top, sw, sh, hsize, K;
    int nr_w[SEEDS_MAXLEV], nr_h[SEEDS_MAXLEV];
    unsigned* bin;                   // histogram bin of every pixel (bins^channels can exceed 16 bits: the GUI slider goes to 50)
    int* hist[SEEDS_MAXLEV];
    int* T[SEEDS_MAXLEV];
    int* parent[SEEDS_MAXLEV];
    int* nrp;
    int* labels;
    unsigned char* dec;
};

__device__ __forceinline__ int seeds_blk0(const SeedsDev& S, int x, int y)
{
    return min(y / S.sh, S.nr_h[0] - 1) * S.nr_w[0] + min(x / S.sw, S.nr_w[0] - 1);
}
__device__ __forceinline__ int seeds_parent0(const SeedsDev& S, int l, int b)   // immediate parent of block b of level l
{
    const int y = b / S.nr_w[l], x = b - y * S.nr_w[l];
    return min(y / 2, S.nr_h[l + 1] - 1) * S.nr_w[l + 1] + min(x / 2, S.nr_w[l + 1] - 1);
}

// ---- initImage -------------------------------------------------------------------------------------------------
__global__ void seeds_bins_kernel(SeedsDev S, const unsigned char* __restrict__ img, size_t ipitch)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= S.w) return;
    const unsigned char* p = img + (size_t)y * ipitch + (size_t)x * S.C;
    unsigned b = 0;
    for (int c = 0; c < S.C; c++) b = b * S.bins + (unsigned)(p[c] * S.bins / 256);
    S.bin[(size_t)y * S.w + x] = b;
    const int blk = seeds_blk0(S, x, y);
    atomicAdd(&S.hist[0][(size_t)blk * S.hsize + b], 1);
    atomicAdd(&S.T[0][blk], 1);
}
__global__ void seeds_levelup_kernel(SeedsDev S, int l)      // level l -> l+1, thread per (block, bin); bin == hsize: T
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int nb = S.nr_w[l] * S.nr_h[l];
    if (i >= (long long)nb * (S.hsize + 1)) return;
    const int b = (int)(i / (S.hsize + 1)), n = (int)(i - (long long)b * (S.hsize + 1));
    const int p = seeds_parent0(S, l, b);
    if (n < S.hsize) {
        const int v = S.hist[l][(size_t)b * S.hsize + n];
        if (v) atomicAdd(&S.hist[l + 1][(size_t)p * S.hsize + n], v);
    } else atomicAdd(&S.T[l + 1][p], S.T[l][b]);
}
// parent[l][b] = immediate parent, for every level (parents are re-pointed by go_down later)
__global__ void seeds_parent_init_kernel(SeedsDev S, int l, int count_nrp)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= S.nr_w[l] * S.nr_h[l]) return;
    const int p = l < S.top ? seeds_parent0(S, l, b) : b;
    S.parent[l][b] = p;
    if (count_nrp) atomicAdd(&S.nrp[p], 1);
}
__global__ void seeds_go_down_kernel(SeedsDev S, int cur)
{
    const int nl = cur - 1;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= S.nr_w[nl] * S.nr_h[nl]) return;
    const int p = S.parent[cur][seeds_parent0(S, nl, b)];
    S.parent[nl][b] = p;
    atomicAdd(&S.nrp[p], 1);
}
__global__ void seeds_update_labels_kernel(SeedsDev S)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= S.w) return;
    const int b = seeds_blk0(S, x, y);
    S.labels[(size_t)y * S.w + x] = S.top > 0 ? S.parent[0][b] : b;
}
// labels of a fresh object (before the first iterate): the top-level grid
__global__ void seeds_grid_labels_kernel(SeedsDev S)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= S.w) return;
    int b = seeds_blk0(S, x, y);
    for (int l = 0; l < S.top; l++) b = seeds_parent0(S, l, b);
    S.labels[(size_t)y * S.w + x] = b;
}

// ---- split checks (a22 is the block / pixel that would move) ------------------------------------------------------
__device__ __forceinline__ bool split_hf(int a11, int a12, int a21, int a22, int a31, int a32)
{
    if (a22 != a21 && a22 == a12 && a22 == a32) return false;
    if (a22 != a11 && a22 == a12 && a22 == a21) return false;
    if (a22 != a31 && a22 == a32 && a22 == a21) return false;
    return true;
}
__device__ __forceinline__ bool split_hb(int a12, int a13, int a22, int a23, int a32, int a33)
{
    if (a22 != a23 && a22 == a12 && a22 == a32) return false;
    if (a22 != a13 && a22 == a12 && a22 == a23) return false;
    if (a22 != a33 && a22 == a32 && a22 == a23) return false;
    return true;
}
__device__ __forceinline__ bool split_vf(int a11, int a12, int a13, int a21, int a22, int a23)
{
    if (a22 != a12 && a22 == a21 && a22 == a23) return false;
    if (a22 != a11 && a22 == a21 && a22 == a12) return false;
    if (a22 != a13 && a22 == a23 && a22 == a12) return false;
    return true;
}
__device__ __forceinline__ bool split_vb(int a21, int a22, int a23, int a31, int a32, int a33)
{
    if (a22 != a32 && a22 == a21 && a22 == a23) return false;
    if (a22 != a31 && a22 == a21 && a22 == a32) return false;
    if (a22 != a33 && a22 == a23 && a22 == a32) return false;
    return true;
}

// ---- block level --------------------------------------------------------------------------------------------------
// gain of giving block (level, sub), currently part of superpixel `own`, to superpixel `tgt`.  One warp per call: the lanes
// split the histogram bins (coalesced reads of the three histograms), integer partial sums, butterfly reduction -- every
// lane ends with the same exact totals and evaluates the same float expression.
__device__ __forceinline__ float seeds_intersect_conditional(const SeedsDev& S, int tgt, int own, int level, int sub, int lane)
{
    const int* hA = S.hist[S.top] + (size_t)tgt * S.hsize;
    const int* hB = S.hist[S.top] + (size_t)own * S.hsize;
    const int* h2 = S.hist[level] + (size_t)sub * S.hsize;
    const long long cA = S.T[S.top][tgt], c2 = S.T[level][sub], cB = (long long)S.T[S.top][own] - c2;
    long long sumA = 0, sumB = 0;
    for (int n = lane; n < S.hsize; n += 32) {
        const long long v2 = h2[n];
        const long long a = hA[n] * c2, b = v2 * cA;
        sumA += a < b ? a : b;
        const long long c = (hB[n] - v2) * c2, d = v2 * cB;
        sumB += c < d ? c : d;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) { sumA += __shfl_xor_sync(0xffffffffu, sumA, o); sumB += __shfl_xor_sync(0xffffffffu, sumB, o); }
    const float iA = __fdiv_rn(__ll2float_rn(sumA), __fmul_rn(__ll2float_rn(cA), __ll2float_rn(c2)));
    const float iB = cB > 0 ? __fdiv_rn(__ll2float_rn(sumB), __fmul_rn(__ll2float_rn(cB), __ll2float_rn(c2))) : 0.f;
    return __fsub_rn(iA, iB);
}
// phase (c4, c2): horizontal: x mod 4 == c4, y mod 2 == c2; vertical: y mod 4 == c4, x mod 2 == c2
__device__ __forceinline__ bool seeds_phase_coords(int i, int j, int vertical, int c4, int c2, int nw, int nh, int& x, int& y)
{
    if (!vertical) { x = c4 + 4 * i; y = c2 + 2 * j; return x >= 1 && x < nw - 2 && y >= 1 && y < nh - 1; }
    x = c2 + 2 * i; y = c4 + 4 * j;
    return x >= 1 && x < nw - 1 && y >= 1 && y < nh - 2;
}
// one warp per pair (all lanes take the same branches: the pair, its labels and the split checks are warp-uniform)
__global__ void seeds_block_decide_kernel(SeedsDev S, int level, int vertical, int c4, int c2, float req, int ni)
{
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    int x, y;
    const int nw = S.nr_w[level], nh = S.nr_h[level];
    if (!seeds_phase_coords(gw % ni, gw / ni, vertical, c4, c2, nw, nh, x, y)) return;
    const int* P = S.parent[level];
    const int st = nw;
#define PP(yy, xx) P[(yy) * st + (xx)]
    const int A = PP(y, x), B = vertical ? PP(y + 1, x) : PP(y, x + 1);
    int d = 0;
    if (A != B) {
        const int nA = S.nrp[A], nB = S.nrp[B];
        if (!vertical) {
            if (nA == 2 || (nA > 2 && split_hf(PP(y - 1, x - 1), PP(y - 1, x), PP(y, x - 1), A, PP(y + 1, x - 1), PP(y + 1, x))))
                if (seeds_intersect_conditional(S, B, A, level, y * st + x, lane) > req) d = 1;
            if (!d && (nB == 2 || (nB > 2 && split_hb(PP(y - 1, x + 1), PP(y - 1, x + 2), B, PP(y, x + 2), PP(y + 1, x + 1), PP(y + 1, x + 2)))))
                if (seeds_intersect_conditional(S, A, B, level, y * st + x + 1, lane) > req) d = 2;
        } else {
            if (nA == 2 || (nA > 2 && split_vf(PP(y - 1, x - 1), PP(y - 1, x), PP(y - 1, x + 1), PP(y, x - 1), A, PP(y, x + 1))))
                if (seeds_intersect_conditional(S, B, A, level, y * st + x, lane) > req) d = 1;
            if (!d && (nB == 2 || (nB > 2 && split_vb(PP(y + 1, x - 1), B, PP(y + 1, x + 1), PP(y + 2, x - 1), PP(y + 2, x), PP(y + 2, x + 1)))))
                if (seeds_intersect_conditional(S, A, B, level, (y + 1) * st + x, lane) > req) d = 2;
        }
    }
#undef PP
    if (lane == 0) S.dec[y * nw + x] = (unsigned char)d;
}
// one warp per pair: the lanes split the histogram bins
__global__ void seeds_block_apply_kernel(SeedsDev S, int level, int vertical, int c4, int c2, int ni)
{
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    int x, y;
    const int nw = S.nr_w[level], nh = S.nr_h[level];
    if (!seeds_phase_coords(gw % ni, gw / ni, vertical, c4, c2, nw, nh, x, y)) return;
    const int d = S.dec[y * nw + x];
    if (!d) return;
    const int i0 = y * nw + x, i1 = vertical ? i0 + nw : i0 + 1;
    const int A = S.parent[level][i0], B = S.parent[level][i1];
    const int sub = d == 1 ? i0 : i1, from = d == 1 ? A : B, to = d == 1 ? B : A;
    const int* h2 = S.hist[level] + (size_t)sub * S.hsize;
    int* hf = S.hist[S.top] + (size_t)from * S.hsize;
    int* ht = S.hist[S.top] + (size_t)to * S.hsize;
    for (int n = lane; n < S.hsize; n += 32) {
        const int v = h2[n];
        if (v) { atomicSub(&hf[n], v); atomicAdd(&ht[n], v); }
    }
    __syncwarp();
    if (lane == 0) {
        const int t = S.T[level][sub];
        atomicSub(&S.T[S.top][from], t); atomicAdd(&S.T[S.top][to], t);
        atomicSub(&S.nrp[from], 1); atomicAdd(&S.nrp[to], 1);
        S.parent[level][sub] = to;
    }
}

// ---- pixel level --------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool seeds_probability(const SeedsDev& S, int idx, int l1, int l2, int prior1, int prior2)
{
    const unsigned color = S.bin[idx];
    const int* H = S.hist[S.top];
    const int* T = S.T[S.top];
    const float t1 = (float)T[l1], t2 = (float)T[l2];
    float P1 = __fmul_rn((float)H[(size_t)l1 * S.hsize + color], t2);
    float P2 = __fmul_rn((float)H[(size_t)l2 * S.hsize + color], t1);
    if (S.prior) {
        float p = __fdiv_rn((float)prior1, (float)prior2);
        switch (S.prior) {
        case 5: p = __fmul_rn(p, p);
        case 4: p = __fmul_rn(p, p);
        case 3: p = __fmul_rn(p, p);
        case 2:
            p = __fmul_rn(p, p);
            P1 = __fmul_rn(P1, t2);
            P2 = __fmul_rn(P2, t1);
        case 1:
            P1 = __fmul_rn(P1, p);
            break;
        }
    }
    return P2 > P1;
}
__global__ void seeds_pixel_decide_kernel(SeedsDev S, int vertical, int c4, int c2, int fb)
{
    int x, y;
    const int w = S.w, h = S.h;
    if (!seeds_phase_coords(blockIdx.x * blockDim.x + threadIdx.x, blockIdx.y, vertical, c4, c2, w, h, x, y)) return;
    const int* L = S.labels;
#define LL(yy, xx) L[(size_t)(yy) * w + (xx)]
    const int A = LL(y, x), B = vertical ? LL(y + 1, x) : LL(y, x + 1);
    int d = 0;
    if (A != B) {
        int pA = 0, pB = 0;
        bool okf, okb;
        const int i0 = y * w + x;
        int i1;
        if (!vertical) {
            // 3 rows x 4 columns around the pair
            const int a11 = LL(y - 1, x - 1), a12 = LL(y - 1, x), a13 = LL(y - 1, x + 1), a14 = LL(y - 1, x + 2);
            const int a21 = LL(y, x - 1), a24 = LL(y, x + 2);
            const int a31 = LL(y + 1, x - 1), a32 = LL(y + 1, x), a33 = LL(y + 1, x + 1), a34 = LL(y + 1, x + 2);
            if (S.prior) {
                pA = (a11 == A) + (a12 == A) + (a13 == A) + (a14 == A) + (a21 == A) + (a24 == A) + (a31 == A) + (a32 == A) + (a33 == A) + (a34 == A);
                pB = (a11 == B) + (a12 == B) + (a13 == B) + (a14 == B) + (a21 == B) + (a24 == B) + (a31 == B) + (a32 == B) + (a33 == B) + (a34 == B);
            }
            okf = split_hf(a11, a12, a21, A, a31, a32);
            okb = split_hb(a13, a14, B, a24, a33, a34);
            i1 = i0 + 1;
        } else {
            // 4 rows x 3 columns around the pair
            const int a11 = LL(y - 1, x - 1), a12 = LL(y - 1, x), a13 = LL(y - 1, x + 1);
            const int a21 = LL(y, x - 1), a23 = LL(y, x + 1);
            const int a31 = LL(y + 1, x - 1), a33 = LL(y + 1, x + 1);
            const int a41 = LL(y + 2, x - 1), a42 = LL(y + 2, x), a43 = LL(y + 2, x + 1);
            if (S.prior) {
                pA = (a11 == A) + (a12 == A) + (a13 == A) + (a21 == A) + (a23 == A) + (a31 == A) + (a33 == A) + (a41 == A) + (a42 == A) + (a43 == A);
                pB = (a11 == B) + (a12 == B) + (a13 == B) + (a21 == B) + (a23 == B) + (a31 == B) + (a33 == B) + (a41 == B) + (a42 == B) + (a43 == B);
            }
            okf = split_vf(a11, a12, a13, a21, A, a23);
            okb = split_vb(a31, B, a33, a41, a42, a43);
            i1 = i0 + w;
        }
        if (fb) {
            if (okf) {
                if (seeds_probability(S, i0, A, B, pA, pB)) d = 1;
                else if (okb && seeds_probability(S, i1, B, A, pB, pA)) d = 2;
            }
        } else {
            if (okb) {
                if (seeds_probability(S, i1, B, A, pB, pA)) d = 2;
                else if (okf && seeds_probability(S, i0, A, B, pA, pB)) d = 1;
            }
        }
    }
#undef LL
    S.dec[(size_t)y * w + x] = (unsigned char)d;
}
__global__ void seeds_pixel_apply_kernel(SeedsDev S, int vertical, int c4, int c2)
{
    int x, y;
    const int w = S.w, h = S.h;
    if (!seeds_phase_coords(blockIdx.x * blockDim.x + threadIdx.x, blockIdx.y, vertical, c4, c2, w, h, x, y)) return;
    const int d = S.dec[(size_t)y * w + x];
    if (!d) return;
    const int i0 = y * w + x, i1 = vertical ? i0 + w : i0 + 1;
    const int A = S.labels[i0], B = S.labels[i1];
    const int idx = d == 1 ? i0 : i1, from = d == 1 ? A : B, to = d == 1 ? B : A;
    const unsigned color = S.bin[idx];
    atomicSub(&S.hist[S.top][(size_t)from * S.hsize + color], 1);
    atomicAdd(&S.hist[S.top][(size_t)to * S.hsize + color], 1);
    atomicSub(&S.T[S.top][from], 1); atomicAdd(&S.T[S.top][to], 1);
    S.labels[idx] = to;
}

// ---- engine ----------------------------------------------------------------------------------------------------
struct SeedsEngine {
    int device = 0;
    cudaStream_t st = nullptr;
    SeedsDev S{};
    int double_step = 0;
    size_t ipitch = 0;
    unsigned char* d_img = nullptr;
    unsigned char* d_mask = nullptr;
    long long launches = 0;
    PostWorkspace post;
    std::vector<void*> allocs;

    template <typename T> int dalloc(T** p, size_t n)
    {
        SPX_CHECK(pool_alloc((void**)p, n * sizeof(T), st));
        allocs.push_back((void*)*p);
        return SPX_OK;
    }
    ~SeedsEngine()
    {
        cudaSetDevice(device);
        if (st) cudaStreamSynchronize(st);
        for (auto& kv : graphs) cudaGraphExecDestroy(kv.second);
        for (void* p : allocs) pool_free(p, st);
        post.release();
        if (st) cudaStreamDestroy(st);
    }

    int create(int w, int h, int C, int num_superpixels, int num_levels, int prior, int bins, int dstep)
    {
        SPX_REQUIRE(w > 0 && h > 0 && h <= SPX_MAX_HEIGHT && (long long)w * h < (1ll << 30), SPX_ERR_BADARG, "createSuperpixelSEEDS: bad image size %dx%d", w, h);
        SPX_REQUIRE(C >= 1 && C <= 4, SPX_ERR_BADARG, "createSuperpixelSEEDS: 1..4 channels supported, got %d", C);
        SPX_REQUIRE(num_superpixels >= 1 && num_levels >= 1 && bins >= 1, SPX_ERR_BADARG, "createSuperpixelSEEDS: bad parameter");
        double hs = 1.0;
        for (int c = 0; c < C; c++) hs *= bins;
        SPX_REQUIRE(hs < 2147483648.0, SPX_ERR_BADARG, "createSuperpixelSEEDS: histogram_bins^channels = %.0f does not fit an int", hs);
        const int nsp_h = (int)sqrtf((float)num_superpixels * (float)h / (float)w);
        SPX_REQUIRE(nsp_h >= 1, SPX_ERR_BADARG, "createSuperpixelSEEDS: too few superpixels for this aspect ratio");
        const int nsp_w = (int)floorf((float)nsp_h * (float)w / (float)h);
        SPX_REQUIRE(nsp_w >= 1, SPX_ERR_BADARG, "createSuperpixelSEEDS: too few superpixels for this aspect ratio");
        int L = (num_levels > SEEDS_MAXLEV ? SEEDS_MAXLEV : num_levels) + 1;
        float wf, hf;
        do {
            L--;
            wf = (float)w / (float)nsp_w / (float)(1 << (L - 1));
            hf = (float)h / (float)nsp_h / (float)(1 << (L - 1));
        } while ((wf < 1.f || hf < 1.f) && L > 1);
        SPX_REQUIRE(wf >= 1.f && hf >= 1.f, SPX_ERR_BADARG, "createSuperpixelSEEDS: more superpixels than pixels");
        S.w = w; S.h = h; S.C = C; S.bins = bins; S.prior = prior > 5 ? 5 : (prior < 0 ? 0 : prior);
        double_step = dstep ? 1 : 0;
        S.nlev = L; S.top = L - 1; S.hsize = (int)hs;
        S.sw = (int)ceilf(wf); S.sh = (int)ceilf(hf);
        S.nr_w[0] = w / S.sw; S.nr_h[0] = h / S.sh;
        for (int l = 1; l < L; l++) { S.nr_w[l] = S.nr_w[l - 1] / 2; S.nr_h[l] = S.nr_h[l - 1] / 2; }
        SPX_REQUIRE(S.nr_w[S.top] >= 1 && S.nr_h[S.top] >= 1, SPX_ERR_BADARG, "createSuperpixelSEEDS: too many levels for this image");
        S.K = S.nr_w[S.top] * S.nr_h[S.top];
        const size_t N = (size_t)w * h;
        ipitch = round_up((size_t)w * C, 16);
        {   // upstream allocates one histogram of bins^channels entries per block of every level, too: a request that cannot
            // fit fails there with cv::Exception (OutOfMemory); here with SPX_ERR_NOMEM before anything is allocated
            double need = 0;
            for (int l = 0; l < L; l++) need += (double)S.nr_w[l] * S.nr_h[l] * (hs + 2.0) * sizeof(int);
            size_t fr = ~(size_t)0, tot = 0;
            if (need > 1e9) SPX_CUDA(cudaMemGetInfo(&fr, &tot));   // (the query itself costs a good part of a millisecond: only when it can matter)
            SPX_REQUIRE(need < 0.9 * (double)fr, SPX_ERR_NOMEM,
                        "createSuperpixelSEEDS: the histograms of %d levels with %d^%d bins need %.1f GB, %.1f GB are free", L, bins, C, need / 1e9, (double)fr / 1e9);
        }
        SPX_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        SPX_CHECK(dalloc(&d_img, ipitch * h));
        SPX_CHECK(dalloc(&d_mask, N));
        SPX_CHECK(dalloc(&S.bin, N));
        SPX_CHECK(dalloc(&S.labels, N));
        SPX_CHECK(dalloc(&S.dec, N));
        SPX_CHECK(dalloc(&S.nrp, (size_t)S.K));
        for (int l = 0; l < L; l++) {
            const size_t nb = (size_t)S.nr_w[l] * S.nr_h[l];
            SPX_CHECK(dalloc(&S.hist[l], nb * S.hsize));
            SPX_CHECK(dalloc(&S.T[l], nb));
            SPX_CHECK(dalloc(&S.parent[l], nb));
        }
        seeds_grid_labels_kernel<<<dim3(cdiv(w, 256), h), 256, 0, st>>>(S);
        launches++;
        SPX_CUDA(cudaGetLastError());
        return SPX_OK;
    }

    void phase_grid(int vertical, int nw, int nh, dim3& grid, int& ni)
    {
        // horizontal: i walks x in steps of 4, j walks y in steps of 2 (vertical: the other way round)
        ni = vertical ? cdiv(nw, 2) : cdiv(nw, 4);
        const int nj = vertical ? cdiv(nh, 4) : cdiv(nh, 2);
        grid = dim3(cdiv(ni, 128), nj);
    }
    void update_blocks(int level, float req)
    {
        const int nw = S.nr_w[level], nh = S.nr_h[level];
        for (int vertical = 0; vertical < 2; vertical++) {
            dim3 grid; int ni;
            phase_grid(vertical, nw, nh, grid, ni);
            const int nj = grid.y;
            for (int c2 = 0; c2 < 2; c2++)
                for (int c4 = 0; c4 < 4; c4++) {
                    seeds_block_decide_kernel<<<cdiv(ni * nj * 32, 128), 128, 0, st>>>(S, level, vertical, c4, c2, req, ni);
                    seeds_block_apply_kernel<<<cdiv(ni * nj * 32, 128), 128, 0, st>>>(S, level, vertical, c4, c2, ni);
                    launches += 2;
                }
        }
    }
    void update_pixels(int fb)
    {
        for (int vertical = 0; vertical < 2; vertical++) {
            dim3 grid; int ni;
            phase_grid(vertical, S.w, S.h, grid, ni);
            for (int c2 = 0; c2 < 2; c2++)
                for (int c4 = 0; c4 < 4; c4++) {
                    seeds_pixel_decide_kernel<<<grid, 128, 0, st>>>(S, vertical, c4, c2, fb);
                    seeds_pixel_apply_kernel<<<grid, 128, 0, st>>>(S, vertical, c4, c2);
                    launches += 2;
                }
        }
    }

    // everything after the upload: ~1700 small launches for 10 iterations at 1080p, so the sequence is captured once per
    // iteration count into a CUDA graph and replayed
    std::map<int, cudaGraphExec_t> graphs;
    std::map<int, long long> graph_launches;
    std::set<int> seen_n;
    int iterate(const uint8_t* img, size_t stride, int n)
    {
        SPX_REQUIRE(img && stride >= (size_t)S.w * S.C, SPX_ERR_BADARG, "SuperpixelSEEDS::iterate: bad image pointer/stride");
        SPX_REQUIRE(n >= 0, SPX_ERR_BADARG, "SuperpixelSEEDS::iterate: num_iterations must be >= 0");
        SPX_CHECK(upload_2d(d_img, ipitch, img, stride, (size_t)S.w * S.C, S.h, st));
        // (as for SLIC: a graph only from the second iterate(n) of an object on; the GUI's one-shot objects launch directly)
        const bool first_use = graphs.find(n) == graphs.end() && !seen_n.count(n);
        if (first_use) seen_n.insert(n);
        if (getenv("SPX_NO_GRAPH") || first_use) { SPX_CHECK(enqueue(n)); }
        else {
            auto it = graphs.find(n);
            if (it == graphs.end()) {
                cudaGraph_t graph = nullptr;
                const long long l0 = launches;
                SPX_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
                const int rc = enqueue(n);
                const cudaError_t e = cudaStreamEndCapture(st, &graph);
                if (rc != SPX_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
                SPX_CUDA(e);
                cudaGraphExec_t exec = nullptr;
                SPX_CUDA(cudaGraphInstantiate(&exec, graph, 0));
                cudaGraphDestroy(graph);
                graphs[n] = exec;
                graph_launches[n] = launches - l0;
                launches = l0;
                it = graphs.find(n);
            }
            SPX_CUDA(cudaGraphLaunch(it->second, st));
            launches += graph_launches[n];
        }
        SPX_CUDA(cudaStreamSynchronize(st));        // the caller's image buffer may go away after iterate() returns
        return SPX_OK;
    }
    int enqueue(int n)
    {
        const int w = S.w, h = S.h;
        for (int l = 0; l < S.nlev; l++) {
            const size_t nb = (size_t)S.nr_w[l] * S.nr_h[l];
            SPX_CUDA(cudaMemsetAsync(S.hist[l], 0, nb * S.hsize * sizeof(int), st));
            SPX_CUDA(cudaMemsetAsync(S.T[l], 0, nb * sizeof(int), st));
        }
        SPX_CUDA(cudaMemsetAsync(S.nrp, 0, (size_t)S.K * sizeof(int), st));
        seeds_bins_kernel<<<dim3(cdiv(w, 256), h), 256, 0, st>>>(S, d_img, ipitch);
        launches++;
        for (int l = 0; l < S.top; l++) {
            const long long items = (long long)S.nr_w[l] * S.nr_h[l] * (S.hsize + 1);
            seeds_levelup_kernel<<<(unsigned)((items + 255) / 256), 256, 0, st>>>(S, l);
            launches++;
        }
        int cur = S.top - 1;
        for (int l = 0; l < S.nlev; l++) {
            seeds_parent_init_kernel<<<cdiv(S.nr_w[l] * S.nr_h[l], 256), 256, 0, st>>>(S, l, l == cur ? 1 : 0);
            launches++;
        }
        while (cur >= 0) {
            if (double_step) update_blocks(cur, SEEDS_REQ_CONF);
            update_blocks(cur, 0.f);
            if (cur > 0) {
                SPX_CUDA(cudaMemsetAsync(S.nrp, 0, (size_t)S.K * sizeof(int), st));
                seeds_go_down_kernel<<<cdiv(S.nr_w[cur - 1] * S.nr_h[cur - 1], 256), 256, 0, st>>>(S, cur);
                launches++;
            }
            cur--;
        }
        seeds_update_labels_kernel<<<dim3(cdiv(w, 256), h), 256, 0, st>>>(S);
        launches++;
        int fb = 1;
        for (int it = 0; it < n; it++) { update_pixels(fb); fb = !fb; }
        SPX_CUDA(cudaGetLastError());
        return SPX_OK;
    }
};

}  // namespace spx

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
using spx::SeedsEngine;
struct spx_seeds { SeedsEngine* e; };

#define SEEDS_ENTER(h)                                                                    \
    SPX_REQUIRE((h) && (h)->e, SPX_ERR_BADARG, "NULL SuperpixelSEEDS handle");            \
    SeedsEngine* e = (h)->e;                                                              \
    SPX_CUDA(cudaSetDevice(e->device))

extern "C" int spx_seeds_create(int w, int h, int C, int num_superpixels, int num_levels, int prior, int bins, int double_step,
                                int device, spx_seeds_t** out) try
{
    using namespace spx;
    SPX_REQUIRE(out, SPX_ERR_BADARG, "createSuperpixelSEEDS: out is NULL");
    *out = nullptr;
    SPX_CHECK(require_device(device));
    SeedsEngine* e = new SeedsEngine();
    e->device = device;
    int rc = e->create(w, h, C, num_superpixels, num_levels, prior, bins, double_step);
    if (rc != SPX_OK) { delete e; return rc; }
    *out = new spx_seeds{e};
    return SPX_OK;
} SPX_API_CATCH
extern "C" int spx_seeds_iterate(spx_seeds_t* h, const uint8_t* image, size_t stride, int n) try
{
    SEEDS_ENTER(h);
    return e->iterate(image, stride, n);
} SPX_API_CATCH
extern "C" int spx_seeds_get_number_of_superpixels(spx_seeds_t* h, int* out) try
{
    SEEDS_ENTER(h);
    SPX_REQUIRE(out, SPX_ERR_BADARG, "getNumberOfSuperpixels: out is NULL");
    *out = e->S.K;
    return SPX_OK;
} SPX_API_CATCH
extern "C" int spx_seeds_get_labels(spx_seeds_t* h, int32_t* dst, size_t dst_stride) try
{
    SEEDS_ENTER(h);
    SPX_REQUIRE(dst && dst_stride >= (size_t)e->S.w * 4, SPX_ERR_BADARG, "getLabels: bad destination");
    SPX_CHECK(spx::download_2d(dst, dst_stride, e->S.labels, (size_t)e->S.w * 4, (size_t)e->S.w * 4, e->S.h, e->st));
    SPX_CUDA(cudaStreamSynchronize(e->st));
    return SPX_OK;
} SPX_API_CATCH
extern "C" int spx_seeds_get_label_contour_mask(spx_seeds_t* h, uint8_t* dst, size_t dst_stride, int thick_line) try
{
    SEEDS_ENTER(h);
    SPX_REQUIRE(dst && dst_stride >= (size_t)e->S.w, SPX_ERR_BADARG, "getLabelContourMask: bad destination");
    SPX_CHECK(e->post.alloc(e->S.w, e->S.h, e->st));
    long long l0 = e->post.launches;
    SPX_CHECK(spx::contour_mask_dev(e->post, e->S.labels, e->S.w, thick_line, 1, e->d_mask, e->S.w, e->st));
    e->launches += e->post.launches - l0;
    SPX_CHECK(spx::download_2d(dst, dst_stride, e->d_mask, e->S.w, e->S.w, e->S.h, e->st));
    SPX_CUDA(cudaStreamSynchronize(e->st));
    return SPX_OK;
} SPX_API_CATCH
extern "C" int spx_seeds_launch_count(spx_seeds_t* h, long long* out) try
{
    SPX_REQUIRE(h && h->e && out, SPX_ERR_BADARG, "launch_count: bad argument");
    *out = h->e->launches;
    return SPX_OK;
} SPX_API_CATCH
extern "C" int spx_seeds_destroy(spx_seeds_t* h) try
{
    if (!h) return SPX_OK;
    delete h->e;
    delete h;
    return SPX_OK;
} SPX_API_CATCH
