/* spx_oracle_seeds.c -- CPU ORACLE for cv::ximgproc::SuperpixelSEEDS (test infrastructure, NOT product code).
 *
 * PARITY UNPINNED: opencv_contrib's seeds.cpp is not available here; the only artefact of the reference that pins
 * anything is screenshots/screenshot-segmentation.jpg (3888x5184, 1000 requested, 4 levels -> 972 superpixels), which
 * fixes the grid sizing below.  Everything else restates Van den Bergh et al., "SEEDS: Superpixels Extracted via
 * Energy-Driven Sampling" in the structure of the authors' code that upstream ports (SURVEY.md Appendix C.4;
 * reference call sites mainwindow.cpp:2123-2127):
 *
 *   sizing     nsp_h = (int)sqrtf(num*h/w), nsp_w = (int)(nsp_h*w/h); levels lowered until a level-0 block is >= 1 px;
 *              seeds_w = ceil(w/nsp_w/2^(L-1)), blocks per level halve (last block absorbs the remainder)
 *   bins       bin = sum_c (v_c*bins/256) * bins^(C-1-c)   (first channel most significant)
 *   iterate    initImage; for level = top-1 .. 0: [double_step: updateBlocks(level, 0.1)], updateBlocks(level, 0),
 *              goDownOneLevel; updateLabels; num_iterations x updatePixels
 *   blocks     a block on a superpixel boundary moves to the neighbouring superpixel when the conditional histogram
 *              intersection gain exceeds the required confidence and the move cannot split / empty its superpixel
 *   pixels     a boundary pixel moves when hist[B][bin]*T[A] > hist[A][bin]*T[B], modulated by the 3x4 / 4x3
 *              neighbourhood prior raised to 2^(prior-1); checkSplit_* keeps superpixels connected
 *
 * Upstream is Gauss-Seidel: every accepted move immediately edits the histograms the next decision reads, in raster
 * order.  No parallel machine can reproduce that order, so this oracle has TWO visiting orders over the same decision
 * rules:
 *   order 0 (sequential)      the upstream raster order, x++/y++ skip after a backward move;
 *   order 1 (phase-parallel)  the order the CUDA path uses: a sweep is cut into 8 phases by (x mod 4, y mod 2)
 *                             [(y mod 4, x mod 2) for vertical sweeps]; all decisions of a phase are taken against the
 *                             state at the start of the phase, then applied together.  Moves of one phase are >= 4
 *                             columns or >= 2 rows apart, so no decision reads a block / pixel another one moves.
 * The CUDA path must match order 1 bit for bit; order 1 is compared with order 0 statistically (boundary recall,
 * undersegmentation error, energy) in tests/.
 * Histogram sums are integers, the intersection is float(sum)/(float(countA)*float(count2)) -- order free.
 */
#include "spx_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define SEEDS_MAXLEV 24
#define SEEDS_REQ_CONF 0.1f

struct spxo_seeds {
    int w, h, C, bins, prior, double_step;
    int nlev, top, sw, sh;
    int nr_w[SEEDS_MAXLEV], nr_h[SEEDS_MAXLEV];
    int hsize, K, order, fb;
    uint32_t* bin;              /* per pixel */
    int* blk0;                  /* per pixel: level-0 block */
    int* parent0[SEEDS_MAXLEV]; /* immediate parent (level l -> l+1) */
    int* parent[SEEDS_MAXLEV];  /* current superpixel of a block */
    int* hist[SEEDS_MAXLEV];    /* [block*hsize + bin] pixel counts */
    int* T[SEEDS_MAXLEV];       /* pixels per block */
    int* nrp;                   /* blocks of the current level per superpixel */
    int32_t* labels;            /* per pixel superpixel */
};

spxo_seeds* spxo_seeds_create(int w, int h, int channels, int num_superpixels, int num_levels, int prior,
                              int histogram_bins, int double_step)
{
    if (w <= 0 || h <= 0 || channels < 1 || channels > 4 || num_superpixels < 1 || num_levels < 1 || histogram_bins < 1) return NULL;
    double hs = 1.0;
    for (int c = 0; c < channels; c++) hs *= histogram_bins;
    if (hs >= 2147483648.0) return NULL;        /* (memory is the only other limit, as upstream) */
    const int nsp_h = (int)sqrtf((float)num_superpixels * (float)h / (float)w);
    if (nsp_h < 1) return NULL;
    const int nsp_w = (int)floorf((float)nsp_h * (float)w / (float)h);
    if (nsp_w < 1) return NULL;
    int L = (num_levels > SEEDS_MAXLEV ? SEEDS_MAXLEV : num_levels) + 1;
    float wf, hf;
    do {
        L--;
        wf = (float)w / (float)nsp_w / (float)(1 << (L - 1));
        hf = (float)h / (float)nsp_h / (float)(1 << (L - 1));
    } while ((wf < 1.f || hf < 1.f) && L > 1);
    if (wf < 1.f || hf < 1.f) return NULL;
    spxo_seeds* s = (spxo_seeds*)calloc(1, sizeof(*s));
    s->w = w; s->h = h; s->C = channels; s->bins = histogram_bins; s->prior = prior > 5 ? 5 : (prior < 0 ? 0 : prior);
    s->double_step = double_step ? 1 : 0;
    s->nlev = L; s->top = L - 1; s->hsize = (int)hs;
    s->sw = (int)ceilf(wf); s->sh = (int)ceilf(hf);
    s->nr_w[0] = w / s->sw; s->nr_h[0] = h / s->sh;
    for (int l = 1; l < L; l++) { s->nr_w[l] = s->nr_w[l - 1] / 2; s->nr_h[l] = s->nr_h[l - 1] / 2; }
    if (s->nr_w[s->top] < 1 || s->nr_h[s->top] < 1) { free(s); return NULL; }
    s->K = s->nr_w[s->top] * s->nr_h[s->top];
    const size_t N = (size_t)w * h;
    s->bin = (uint32_t*)malloc(N * sizeof(uint32_t));
    s->blk0 = (int*)malloc(N * sizeof(int));
    s->labels = (int32_t*)calloc(N, sizeof(int32_t));
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int bx = x / s->sw, by = y / s->sh;
            if (bx > s->nr_w[0] - 1) bx = s->nr_w[0] - 1;
            if (by > s->nr_h[0] - 1) by = s->nr_h[0] - 1;
            s->blk0[(size_t)y * w + x] = by * s->nr_w[0] + bx;
        }
    for (int l = 0; l < L; l++) {
        const int nb = s->nr_w[l] * s->nr_h[l];
        s->parent0[l] = (int*)malloc(sizeof(int) * nb);
        s->parent[l] = (int*)malloc(sizeof(int) * nb);
        s->hist[l] = (int*)malloc(sizeof(int) * (size_t)nb * s->hsize);
        s->T[l] = (int*)malloc(sizeof(int) * nb);
        for (int y = 0; y < s->nr_h[l]; y++)
            for (int x = 0; x < s->nr_w[l]; x++) {
                int p = y * s->nr_w[l] + x;
                if (l < s->top) {
                    int px = x / 2, py = y / 2;
                    if (px > s->nr_w[l + 1] - 1) px = s->nr_w[l + 1] - 1;
                    if (py > s->nr_h[l + 1] - 1) py = s->nr_h[l + 1] - 1;
                    s->parent0[l][p] = py * s->nr_w[l + 1] + px;
                } else s->parent0[l][p] = p;
            }
    }
    s->nrp = (int*)malloc(sizeof(int) * s->K);
    /* labels before the first iterate(): the top-level grid */
    for (size_t i = 0; i < N; i++) {
        int b = s->blk0[i];
        for (int l = 0; l < s->top; l++) b = s->parent0[l][b];
        s->labels[i] = b;
    }
    return s;
}

void spxo_seeds_destroy(spxo_seeds* s)
{
    if (!s) return;
    free(s->bin); free(s->blk0); free(s->labels); free(s->nrp);
    for (int l = 0; l < s->nlev; l++) { free(s->parent0[l]); free(s->parent[l]); free(s->hist[l]); free(s->T[l]); }
    free(s);
}
void spxo_seeds_set_order(spxo_seeds* s, int order) { s->order = order ? 1 : 0; }

/* ---- split checks (a22 is the block / pixel that would move) ------------------------------------------------ */
static inline int split_hf(int a11, int a12, int a21, int a22, int a31, int a32)
{
    if (a22 != a21 && a22 == a12 && a22 == a32) return 0;
    if (a22 != a11 && a22 == a12 && a22 == a21) return 0;
    if (a22 != a31 && a22 == a32 && a22 == a21) return 0;
    return 1;
}
static inline int split_hb(int a12, int a13, int a22, int a23, int a32, int a33)
{
    if (a22 != a23 && a22 == a12 && a22 == a32) return 0;
    if (a22 != a13 && a22 == a12 && a22 == a23) return 0;
    if (a22 != a33 && a22 == a32 && a22 == a23) return 0;
    return 1;
}
static inline int split_vf(int a11, int a12, int a13, int a21, int a22, int a23)
{
    if (a22 != a12 && a22 == a21 && a22 == a23) return 0;
    if (a22 != a11 && a22 == a21 && a22 == a12) return 0;
    if (a22 != a13 && a22 == a23 && a22 == a12) return 0;
    return 1;
}
static inline int split_vb(int a21, int a22, int a23, int a31, int a32, int a33)
{
    if (a22 != a32 && a22 == a21 && a22 == a23) return 0;
    if (a22 != a31 && a22 == a21 && a22 == a32) return 0;
    if (a22 != a33 && a22 == a23 && a22 == a32) return 0;
    return 1;
}

/* gain of giving block (level, sub), currently part of superpixel `own`, to superpixel `tgt` */
static float intersect_conditional(const spxo_seeds* s, int tgt, int own, int level, int sub)
{
    const int* hA = s->hist[s->top] + (size_t)tgt * s->hsize;
    const int* hB = s->hist[s->top] + (size_t)own * s->hsize;
    const int* h2 = s->hist[level] + (size_t)sub * s->hsize;
    const long long cA = s->T[s->top][tgt], c2 = s->T[level][sub], cB = (long long)s->T[s->top][own] - c2;
    long long sumA = 0, sumB = 0;
    for (int n = 0; n < s->hsize; n++) {
        const long long a = hA[n] * c2, b = h2[n] * cA;
        sumA += a < b ? a : b;
        const long long c = (hB[n] - h2[n]) * c2, d = h2[n] * cB;
        sumB += c < d ? c : d;
    }
    const float iA = (float)sumA / ((float)cA * (float)c2);
    const float iB = cB > 0 ? (float)sumB / ((float)cB * (float)c2) : 0.f;
    return iA - iB;
}

static void block_move(spxo_seeds* s, int level, int sub, int from, int to)
{
    int* hf_ = s->hist[s->top] + (size_t)from * s->hsize;
    int* ht = s->hist[s->top] + (size_t)to * s->hsize;
    const int* h2 = s->hist[level] + (size_t)sub * s->hsize;
    for (int n = 0; n < s->hsize; n++) { hf_[n] -= h2[n]; ht[n] += h2[n]; }
    s->T[s->top][from] -= s->T[level][sub]; s->T[s->top][to] += s->T[level][sub];
    s->nrp[from]--; s->nrp[to]++;
    s->parent[level][sub] = to;
}

/* 0 none, 1 forward (block/pixel at (x,y) goes to B), 2 backward (the one at (x+1,y) / (x,y+1) goes to A) */
static int block_decide(const spxo_seeds* s, int level, int x, int y, int vertical, float req)
{
    const int* P = s->parent[level];
    const int st = s->nr_w[level];
    const int A = P[y * st + x], B = vertical ? P[(y + 1) * st + x] : P[y * st + x + 1];
    if (A == B) return 0;
#define PP(yy, xx) P[(yy) * st + (xx)]
    if (!vertical) {
        if (s->nrp[A] == 2 || (s->nrp[A] > 2 && split_hf(PP(y - 1, x - 1), PP(y - 1, x), PP(y, x - 1), PP(y, x), PP(y + 1, x - 1), PP(y + 1, x))))
            if (intersect_conditional(s, B, A, level, y * st + x) > req) return 1;
        if (s->nrp[B] == 2 || (s->nrp[B] > 2 && split_hb(PP(y - 1, x + 1), PP(y - 1, x + 2), PP(y, x + 1), PP(y, x + 2), PP(y + 1, x + 1), PP(y + 1, x + 2))))
            if (intersect_conditional(s, A, B, level, y * st + x + 1) > req) return 2;
    } else {
        if (s->nrp[A] == 2 || (s->nrp[A] > 2 && split_vf(PP(y - 1, x - 1), PP(y - 1, x), PP(y - 1, x + 1), PP(y, x - 1), PP(y, x), PP(y, x + 1))))
            if (intersect_conditional(s, B, A, level, y * st + x) > req) return 1;
        if (s->nrp[B] == 2 || (s->nrp[B] > 2 && split_vb(PP(y + 1, x - 1), PP(y + 1, x), PP(y + 1, x + 1), PP(y + 2, x - 1), PP(y + 2, x), PP(y + 2, x + 1))))
            if (intersect_conditional(s, A, B, level, (y + 1) * st + x) > req) return 2;
    }
#undef PP
    return 0;
}
static void block_apply(spxo_seeds* s, int level, int x, int y, int vertical, int dec)
{
    const int st = s->nr_w[level];
    const int i0 = y * st + x, i1 = vertical ? (y + 1) * st + x : i0 + 1;
    const int A = s->parent[level][i0], B = s->parent[level][i1];
    if (dec == 1) block_move(s, level, i0, A, B);
    else if (dec == 2) block_move(s, level, i1, B, A);
}

static void update_blocks(spxo_seeds* s, int level, float req)
{
    const int nw = s->nr_w[level], nh = s->nr_h[level];
    if (s->order == 0) {
        for (int y = 1; y < nh - 1; y++)
            for (int x = 1; x < nw - 2; x++) {
                const int d = block_decide(s, level, x, y, 0, req);
                block_apply(s, level, x, y, 0, d);
                if (d == 2) x++;
            }
        for (int x = 1; x < nw - 1; x++)
            for (int y = 1; y < nh - 2; y++) {
                const int d = block_decide(s, level, x, y, 1, req);
                block_apply(s, level, x, y, 1, d);
                if (d == 2) y++;
            }
    } else {
        unsigned char* dec = (unsigned char*)malloc((size_t)nw * nh);
        for (int cy = 0; cy < 2; cy++)
            for (int cx = 0; cx < 4; cx++) {
                for (int y = 1; y < nh - 1; y++)
                    for (int x = 1; x < nw - 2; x++)
                        if ((x & 3) == cx && (y & 1) == cy) dec[y * nw + x] = (unsigned char)block_decide(s, level, x, y, 0, req);
                for (int y = 1; y < nh - 1; y++)
                    for (int x = 1; x < nw - 2; x++)
                        if ((x & 3) == cx && (y & 1) == cy) block_apply(s, level, x, y, 0, dec[y * nw + x]);
            }
        for (int cx = 0; cx < 2; cx++)
            for (int cy = 0; cy < 4; cy++) {
                for (int x = 1; x < nw - 1; x++)
                    for (int y = 1; y < nh - 2; y++)
                        if ((y & 3) == cy && (x & 1) == cx) dec[y * nw + x] = (unsigned char)block_decide(s, level, x, y, 1, req);
                for (int x = 1; x < nw - 1; x++)
                    for (int y = 1; y < nh - 2; y++)
                        if ((y & 3) == cy && (x & 1) == cx) block_apply(s, level, x, y, 1, dec[y * nw + x]);
            }
        free(dec);
    }
}

static void go_down(spxo_seeds* s, int cur)
{
    const int nl = cur - 1;
    memset(s->nrp, 0, sizeof(int) * s->K);
    const int nb = s->nr_w[nl] * s->nr_h[nl];
    for (int b = 0; b < nb; b++) {
        const int p = s->parent[cur][s->parent0[nl][b]];
        s->parent[nl][b] = p;
        s->nrp[p]++;
    }
}

/* ---- pixel level -------------------------------------------------------------------------------------------- */
static inline int probability(const spxo_seeds* s, int idx, int l1, int l2, int prior1, int prior2)
{
    const uint32_t color = s->bin[idx];
    const int* H = s->hist[s->top];
    const int* T = s->T[s->top];
    float P1 = (float)H[(size_t)l1 * s->hsize + color] * (float)T[l2];
    float P2 = (float)H[(size_t)l2 * s->hsize + color] * (float)T[l1];
    if (s->prior) {
        float p = (float)prior1 / (float)prior2;
        switch (s->prior) {
        case 5: p *= p;  /* fall through */
        case 4: p *= p;  /* fall through */
        case 3: p *= p;  /* fall through */
        case 2:
            p *= p;
            P1 *= (float)T[l2];
            P2 *= (float)T[l1];
            /* fall through */
        case 1:
            P1 *= p;
            break;
        }
    }
    return P2 > P1;
}
static inline int count3x4(const spxo_seeds* s, int x, int y, int label)
{
    const int32_t* L = s->labels; const int w = s->w;
    int c = 0;
    for (int dy = -1; dy <= 1; dy++)
        for (int dx = -1; dx <= 2; dx++) {
            if (dy == 0 && (dx == 0 || dx == 1)) continue;
            c += L[(size_t)(y + dy) * w + x + dx] == label;
        }
    return c;
}
static inline int count4x3(const spxo_seeds* s, int x, int y, int label)
{
    const int32_t* L = s->labels; const int w = s->w;
    int c = 0;
    for (int dy = -1; dy <= 2; dy++)
        for (int dx = -1; dx <= 1; dx++) {
            if (dx == 0 && (dy == 0 || dy == 1)) continue;
            c += L[(size_t)(y + dy) * w + x + dx] == label;
        }
    return c;
}
static int pixel_decide(const spxo_seeds* s, int x, int y, int vertical)
{
    const int32_t* L = s->labels; const int w = s->w;
#define LL(yy, xx) L[(size_t)(yy) * w + (xx)]
    const int A = LL(y, x), B = vertical ? LL(y + 1, x) : LL(y, x + 1);
    if (A == B) return 0;
    int pA = 0, pB = 0;
    if (s->prior) {
        pA = vertical ? count4x3(s, x, y, A) : count3x4(s, x, y, A);
        pB = vertical ? count4x3(s, x, y, B) : count3x4(s, x, y, B);
    }
    int okf, okb, i0 = y * w + x, i1;
    if (!vertical) {
        okf = split_hf(LL(y - 1, x - 1), LL(y - 1, x), LL(y, x - 1), A, LL(y + 1, x - 1), LL(y + 1, x));
        okb = split_hb(LL(y - 1, x + 1), LL(y - 1, x + 2), B, LL(y, x + 2), LL(y + 1, x + 1), LL(y + 1, x + 2));
        i1 = i0 + 1;
    } else {
        okf = split_vf(LL(y - 1, x - 1), LL(y - 1, x), LL(y - 1, x + 1), LL(y, x - 1), A, LL(y, x + 1));
        okb = split_vb(LL(y + 1, x - 1), B, LL(y + 1, x + 1), LL(y + 2, x - 1), LL(y + 2, x), LL(y + 2, x + 1));
        i1 = i0 + w;
    }
#undef LL
    if (s->fb) {          /* forward first; the backward move is only tried when the forward split check passed */
        if (okf) {
            if (probability(s, i0, A, B, pA, pB)) return 1;
            if (okb && probability(s, i1, B, A, pB, pA)) return 2;
        }
    } else {
        if (okb) {
            if (probability(s, i1, B, A, pB, pA)) return 2;
            if (okf && probability(s, i0, A, B, pA, pB)) return 1;
        }
    }
    return 0;
}
static void pixel_apply(spxo_seeds* s, int x, int y, int vertical, int dec)
{
    if (!dec) return;
    const int w = s->w;
    const int i0 = y * w + x, i1 = vertical ? i0 + w : i0 + 1;
    const int A = s->labels[i0], B = s->labels[i1];
    const int idx = dec == 1 ? i0 : i1, from = dec == 1 ? A : B, to = dec == 1 ? B : A;
    int* H = s->hist[s->top];
    H[(size_t)from * s->hsize + s->bin[idx]]--; H[(size_t)to * s->hsize + s->bin[idx]]++;
    s->T[s->top][from]--; s->T[s->top][to]++;
    s->labels[idx] = to;
}
static void update_pixels(spxo_seeds* s)
{
    const int w = s->w, h = s->h;
    if (s->order == 0) {
        for (int y = 1; y < h - 1; y++)
            for (int x = 1; x < w - 2; x++) {
                const int d = pixel_decide(s, x, y, 0);
                pixel_apply(s, x, y, 0, d);
                if (d == 2) x++;
            }
        for (int x = 1; x < w - 1; x++)
            for (int y = 1; y < h - 2; y++) {
                const int d = pixel_decide(s, x, y, 1);
                pixel_apply(s, x, y, 1, d);
                if (d == 2) y++;
            }
    } else {
        unsigned char* dec = (unsigned char*)malloc((size_t)w * h);
        for (int cy = 0; cy < 2; cy++)
            for (int cx = 0; cx < 4; cx++) {
                for (int y = 1 ; y < h - 1; y++)
                    for (int x = 1; x < w - 2; x++)
                        if ((x & 3) == cx && (y & 1) == cy) dec[(size_t)y * w + x] = (unsigned char)pixel_decide(s, x, y, 0);
                for (int y = 1; y < h - 1; y++)
                    for (int x = 1; x < w - 2; x++)
                        if ((x & 3) == cx && (y & 1) == cy) pixel_apply(s, x, y, 0, dec[(size_t)y * w + x]);
            }
        for (int cx = 0; cx < 2; cx++)
            for (int cy = 0; cy < 4; cy++) {
                for (int y = 1; y < h - 2; y++)
                    for (int x = 1; x < w - 1; x++)
                        if ((y & 3) == cy && (x & 1) == cx) dec[(size_t)y * w + x] = (unsigned char)pixel_decide(s, x, y, 1);
                for (int y = 1; y < h - 2; y++)
                    for (int x = 1; x < w - 1; x++)
                        if ((y & 3) == cy && (x & 1) == cx) pixel_apply(s, x, y, 1, dec[(size_t)y * w + x]);
            }
        free(dec);
    }
    s->fb = !s->fb;
}

void spxo_seeds_iterate(spxo_seeds* s, const uint8_t* img, size_t stride, int num_iterations)
{
    const int w = s->w, h = s->h, C = s->C;
    const size_t N = (size_t)w * h;
    /* initImage: bins, level-0 histograms, upper levels by summing children, parents reset */
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const uint8_t* p = img + (size_t)y * stride + (size_t)x * C;
            uint32_t b = 0;
            for (int c = 0; c < C; c++) b = b * s->bins + (uint32_t)(p[c] * s->bins / 256);
            s->bin[(size_t)y * w + x] = b;
        }
    for (int l = 0; l < s->nlev; l++) {
        const int nb = s->nr_w[l] * s->nr_h[l];
        memset(s->hist[l], 0, sizeof(int) * (size_t)nb * s->hsize);
        memset(s->T[l], 0, sizeof(int) * nb);
    }
    for (size_t i = 0; i < N; i++) { s->hist[0][(size_t)s->blk0[i] * s->hsize + s->bin[i]]++; s->T[0][s->blk0[i]]++; }
    for (int l = 0; l < s->top; l++) {
        const int nb = s->nr_w[l] * s->nr_h[l];
        for (int b = 0; b < nb; b++) {
            const int p = s->parent0[l][b];
            for (int n = 0; n < s->hsize; n++) s->hist[l + 1][(size_t)p * s->hsize + n] += s->hist[l][(size_t)b * s->hsize + n];
            s->T[l + 1][p] += s->T[l][b];
        }
    }
    for (int l = 0; l < s->nlev; l++) {
        const int nb = s->nr_w[l] * s->nr_h[l];
        for (int b = 0; b < nb; b++) s->parent[l][b] = s->parent0[l][b];
    }
    s->fb = 1;
    int cur = s->top - 1;
    if (cur >= 0) {
        memset(s->nrp, 0, sizeof(int) * s->K);
        const int nb = s->nr_w[cur] * s->nr_h[cur];
        for (int b = 0; b < nb; b++) s->nrp[s->parent[cur][b]]++;
    }
    while (cur >= 0) {
        if (s->double_step) update_blocks(s, cur, SEEDS_REQ_CONF);
        update_blocks(s, cur, 0.f);
        if (cur > 0) go_down(s, cur);
        cur--;
    }
    /* updateLabels */
    if (s->top > 0) for (size_t i = 0; i < N; i++) s->labels[i] = s->parent[0][s->blk0[i]];
    else for (size_t i = 0; i < N; i++) s->labels[i] = s->blk0[i];
    for (int it = 0; it < num_iterations; it++) update_pixels(s);
}

int spxo_seeds_get_number_of_superpixels(const spxo_seeds* s) { return s->K; }
void spxo_seeds_get_labels(const spxo_seeds* s, int32_t* dst) { memcpy(dst, s->labels, sizeof(int32_t) * (size_t)s->w * s->h); }
void spxo_seeds_get_label_contour_mask(const spxo_seeds* s, uint8_t* dst, int thick_line)
{
    spxo_label_contour_mask_seeds(s->labels, s->w, s->h, thick_line, dst);
}
/* energy of a labelling: sum over superpixels and bins of (hist/T)^2 * T  (higher = more homogeneous colour) */
double spxo_seeds_energy(const spxo_seeds* s, const int32_t* labels)
{
    const size_t N = (size_t)s->w * s->h;
    long long* H = (long long*)calloc((size_t)s->K * s->hsize, sizeof(long long));
    long long* T = (long long*)calloc(s->K, sizeof(long long));
    for (size_t i = 0; i < N; i++) { H[(size_t)labels[i] * s->hsize + s->bin[i]]++; T[labels[i]]++; }
    double e = 0.0;
    for (int k = 0; k < s->K; k++)
        if (T[k])
            for (int n = 0; n < s->hsize; n++) e += (double)H[(size_t)k * s->hsize + n] * (double)H[(size_t)k * s->hsize + n] / (double)T[k];
    free(H); free(T);
    return e / (double)N;
}
