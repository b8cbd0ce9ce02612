/* spx_oracle_lsc.c -- CPU ORACLE for cv::ximgproc::SuperpixelLSC (test infrastructure, NOT product code).
 *
 * PARITY UNPINNED: opencv_contrib's lsc.cpp is not available here (not vendored under /root/reference, no ximgproc
 * binary), and the reference ships no LSC fixture.  This file restates Li & Chen, "Superpixel Segmentation using
 * Linear Spectral Clustering" (CVPR 2015) in the structure of the authors' code that upstream ports
 * (reference call sites mainwindow.cpp:2114-2120):
 *
 *   coefficients   color = 20, dist = color*ratio, threshold = 4, PI = 3.1415926
 *   seeds          K0 = int(float(W*H)/float(S*S)); ColNum = int(sqrt(double(K0)*W/H)); RowNum = K0/ColNum;
 *                  stepx = W/ColNum, stepy = H/RowNum; remainders spread one extra pixel per strip (t1/t2 counters)
 *   features       per pixel 2C+4 values: color*cos/sin(v/255*PI/2) (x2.55 for channels >= 1), dist*cos/sin(x/stepx*PI/2),
 *                  dist*cos/sin(y/stepy*PI/2); weight W(p) = f(p).mean(f); f(p) /= W(p)
 *   initial centre mean of f over the window +-step/4 around each seed (inclusive, clamped)
 *   iterate        windowed (+-step, inclusive, clamped) arg-min of the squared feature distance, clusters in ascending
 *                  order with strict '<'; then centres = W-weighted means, seed = integer mean position
 *   connectivity   pre-pass: 8-connected regions of <= min_element_size pixels take the label of an already visited
 *                  neighbour (the authors' `adj` rule, including its unconditional break and its carry-over);
 *                  post-pass: 4-connected regions smaller than W*H/(K*threshold) are merged, in region order, into the
 *                  adjacent region with the nearest feature centre (centres/weights/sizes updated incrementally).
 *
 * Where the published code leaves the float evaluation order to the compiler or accumulates in an order no parallel
 * machine can reproduce, this restatement FIXES an order-free definition (the CUDA path implements the same one):
 *   - trigonometry is evaluated in double and rounded once to float (tables of 256 / W / H entries);
 *   - feature means: sum over table entries weighted by exact integer histograms, in double, ascending index;
 *   - W(p) = ((wc0[v0] + wc1[v1]) + wc2[v2] ...) + (wx[x] + wy[y]) with per-entry weights w = cos*s_cos + sin*s_sin;
 *   - f_i(p) = raw_i * (1.0f / W(p));
 *   - all sums over pixels are exact fixed-point integers: (int64)(f_i * 2^40) for plain means, (int64)(W*f_i * 2^32)
 *     and (int64)(W * 2^24) for weighted means; a mean is float(double(sum)/double(denominator)/scale);
 *   - ties between equally near neighbour regions in the post-pass go to the smaller region id.
 */
#include "spx_oracle.h"
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>

#define LSC_MAXC 4
#define LSC_MAXD (2 * LSC_MAXC + 4)
#define LSC_PI 3.1415926

struct spxo_lsc {
    int w, h, C, D, S;
    float ratio;
    int K, ColNum, RowNum, stepx, stepy;
    int merge_algo;          /* 0 = the sequential loop (normative); 1 = fixed-point evaluation of the SAME sequential rule (the form the
                                CUDA path runs; must give identical output -- tests/test_oracle.py) */
    int merge_iters;         /* merge_algo 1: iterations the last call needed */
    int merge_order;         /* post-pass visiting order: 0 = ascending region index (authors; default here and in the CUDA path), 1 = hashed index (CUDA opt-in) */
    uint8_t* img;            /* dense interleaved copy */
    int32_t* labels;
    int* sx; int* sy;        /* seed positions */
    float* cen;              /* K x D */
    float colcos[LSC_MAXC][256], colsin[LSC_MAXC][256], wcol[LSC_MAXC][256];
    float *xcos, *xsin, *wx, *ycos, *ysin, *wy;
};

static inline float lsc_features(const spxo_lsc* s, int x, int y, float* f)
{
    const uint8_t* p = s->img + ((size_t)y * s->w + x) * s->C;
    float wc = s->wcol[0][p[0]];
    for (int c = 1; c < s->C; c++) wc = wc + s->wcol[c][p[c]];
    const float Wt = wc + (s->wx[x] + s->wy[y]);
    const float inv = 1.0f / Wt;
    for (int c = 0; c < s->C; c++) { f[2 * c] = s->colcos[c][p[c]] * inv; f[2 * c + 1] = s->colsin[c][p[c]] * inv; }
    f[2 * s->C] = s->xcos[x] * inv; f[2 * s->C + 1] = s->xsin[x] * inv;
    f[2 * s->C + 2] = s->ycos[y] * inv; f[2 * s->C + 3] = s->ysin[y] * inv;
    return Wt;
}

spxo_lsc* spxo_lsc_create(const uint8_t* img, size_t stride, int w, int h, int channels, int region_size, float ratio)
{
    if (!img || w <= 0 || h <= 0 || channels < 1 || channels > LSC_MAXC || region_size < 1) return NULL;
    const int K0 = (int)((float)(w * h) / (float)(region_size * region_size));
    const int ColNum = (int)sqrt((double)K0 * w / h);
    if (ColNum < 1) return NULL;
    const int RowNum = K0 / ColNum;
    if (RowNum < 1) return NULL;
    spxo_lsc* s = (spxo_lsc*)calloc(1, sizeof(*s));
    s->w = w; s->h = h; s->C = channels; s->D = 2 * channels + 4; s->S = region_size; s->ratio = ratio;
    s->ColNum = ColNum; s->RowNum = RowNum; s->K = ColNum * RowNum;
    s->stepx = w / ColNum; s->stepy = h / RowNum;
    const size_t N = (size_t)w * h;
    s->img = (uint8_t*)malloc(N * channels);
    for (int y = 0; y < h; y++) memcpy(s->img + (size_t)y * w * channels, img + (size_t)y * stride, (size_t)w * channels);
    s->labels = (int32_t*)calloc(N, sizeof(int32_t));
    s->sx = (int*)malloc(sizeof(int) * s->K); s->sy = (int*)malloc(sizeof(int) * s->K);
    s->cen = (float*)calloc((size_t)s->K * s->D, sizeof(float));
    s->xcos = (float*)malloc(sizeof(float) * w); s->xsin = (float*)malloc(sizeof(float) * w); s->wx = (float*)malloc(sizeof(float) * w);
    s->ycos = (float*)malloc(sizeof(float) * h); s->ysin = (float*)malloc(sizeof(float) * h); s->wy = (float*)malloc(sizeof(float) * h);

    /* seeds (Seeds() of the authors' code; x is the column here) */
    {
        const int row_remain = h - s->stepy * RowNum, col_remain = w - s->stepx * ColNum;
        int t1 = 1, n = 0;
        for (int i = 0; i < RowNum; i++) {
            int t2 = 1;
            int cy = (int)(i * s->stepy + 0.5 * s->stepy + t1);
            if (cy >= h - 1) cy = h - 1;
            if (t1 < row_remain) t1++;
            for (int j = 0; j < ColNum; j++) {
                int cx = (int)(j * s->stepx + 0.5 * s->stepx + t2);
                if (t2 < col_remain) t2++;
                if (cx >= w - 1) cx = w - 1;
                s->sx[n] = cx; s->sy[n] = cy; n++;
            }
        }
    }
    /* feature tables */
    const double color = 20.0;
    const float dist_coeff = 20.0f * ratio;
    const double dist = (double)dist_coeff;
    for (int c = 0; c < channels; c++) {
        const double scale = c == 0 ? 1.0 : 2.55;
        for (int v = 0; v < 256; v++) {
            const double th = ((double)v / 255.0) * LSC_PI / 2.0;
            s->colcos[c][v] = (float)(color * cos(th) * scale);
            s->colsin[c][v] = (float)(color * sin(th) * scale);
        }
    }
    for (int x = 0; x < w; x++) {
        const double th = ((double)x / (double)s->stepx) * LSC_PI / 2.0;
        s->xcos[x] = (float)(dist * cos(th)); s->xsin[x] = (float)(dist * sin(th));
    }
    for (int y = 0; y < h; y++) {
        const double th = ((double)y / (double)s->stepy) * LSC_PI / 2.0;
        s->ycos[y] = (float)(dist * cos(th)); s->ysin[y] = (float)(dist * sin(th));
    }
    /* feature means from exact histograms, then per-entry weights */
    {
        long long* hist = (long long*)calloc((size_t)channels * 256, sizeof(long long));
        for (size_t i = 0; i < N; i++)
            for (int c = 0; c < channels; c++) hist[c * 256 + s->img[i * channels + c]]++;
        for (int c = 0; c < channels; c++) {
            double a = 0.0, b = 0.0;
            for (int v = 0; v < 256; v++) { a += (double)hist[c * 256 + v] * (double)s->colcos[c][v]; b += (double)hist[c * 256 + v] * (double)s->colsin[c][v]; }
            const float sc = (float)(a / (double)N), ss = (float)(b / (double)N);
            for (int v = 0; v < 256; v++) s->wcol[c][v] = s->colcos[c][v] * sc + s->colsin[c][v] * ss;
        }
        free(hist);
        double a = 0.0, b = 0.0;
        for (int x = 0; x < w; x++) { a += (double)s->xcos[x]; b += (double)s->xsin[x]; }
        float sc = (float)(a * (double)h / (double)N), ss = (float)(b * (double)h / (double)N);
        for (int x = 0; x < w; x++) s->wx[x] = s->xcos[x] * sc + s->xsin[x] * ss;
        a = 0.0; b = 0.0;
        for (int y = 0; y < h; y++) { a += (double)s->ycos[y]; b += (double)s->ysin[y]; }
        sc = (float)(a * (double)w / (double)N); ss = (float)(b * (double)w / (double)N);
        for (int y = 0; y < h; y++) s->wy[y] = s->ycos[y] * sc + s->ysin[y] * ss;
    }
    /* initial centres: plain mean of the normalised features over +-step/4 */
    for (int k = 0; k < s->K; k++) {
        const int x0 = s->sx[k] - s->stepx / 4 <= 0 ? 0 : s->sx[k] - s->stepx / 4;
        const int y0 = s->sy[k] - s->stepy / 4 <= 0 ? 0 : s->sy[k] - s->stepy / 4;
        const int x1 = s->sx[k] + s->stepx / 4 >= w - 1 ? w - 1 : s->sx[k] + s->stepx / 4;
        const int y1 = s->sy[k] + s->stepy / 4 >= h - 1 ? h - 1 : s->sy[k] + s->stepy / 4;
        long long sum[LSC_MAXD] = {0};
        long long count = 0;
        float f[LSC_MAXD];
        for (int y = y0; y <= y1; y++)
            for (int x = x0; x <= x1; x++) {
                lsc_features(s, x, y, f);
                for (int i = 0; i < s->D; i++) sum[i] += (long long)(f[i] * 1099511627776.0f);
                count++;
            }
        for (int i = 0; i < s->D; i++) s->cen[(size_t)k * s->D + i] = (float)((double)sum[i] / (double)count / 1099511627776.0);
    }
    return s;
}

void spxo_lsc_destroy(spxo_lsc* s)
{
    if (!s) return;
    free(s->img); free(s->labels); free(s->sx); free(s->sy); free(s->cen);
    free(s->xcos); free(s->xsin); free(s->wx); free(s->ycos); free(s->ysin); free(s->wy);
    free(s);
}

void spxo_lsc_iterate(spxo_lsc* s, int num_iterations)
{
    const int w = s->w, h = s->h, D = s->D, K = s->K;
    const size_t N = (size_t)w * h;
    float* dist = (float*)malloc(N * sizeof(float));
    long long* sum = (long long*)malloc(sizeof(long long) * (size_t)K * (D + 4));
    for (int itr = 0; itr < num_iterations; itr++) {
        for (size_t i = 0; i < N; i++) dist[i] = FLT_MAX;
        /* assignment: row bands in parallel, clusters in ascending order inside a band (same result as the serial loop) */
        const int nthr = spxo_get_threads();
#pragma omp parallel for schedule(static) num_threads(nthr)
        for (int band = 0; band < nthr; band++) {
            const int r0 = (int)((long long)h * band / nthr), r1 = (int)((long long)h * (band + 1) / nthr);
            float f[LSC_MAXD];
            for (int k = 0; k < K; k++) {
                const int x0 = s->sx[k] - s->stepx <= 0 ? 0 : s->sx[k] - s->stepx;
                int y0 = s->sy[k] - s->stepy <= 0 ? 0 : s->sy[k] - s->stepy;
                const int x1 = s->sx[k] + s->stepx >= w - 1 ? w - 1 : s->sx[k] + s->stepx;
                int y1 = s->sy[k] + s->stepy >= h - 1 ? h - 1 : s->sy[k] + s->stepy;
                if (y0 < r0) y0 = r0;
                if (y1 > r1 - 1) y1 = r1 - 1;
                const float* c = s->cen + (size_t)k * D;
                for (int y = y0; y <= y1; y++)
                    for (int x = x0; x <= x1; x++) {
                        lsc_features(s, x, y, f);
                        float d = 0.f;
                        for (int i = 0; i < D; i++) { float t = f[i] - c[i]; d = d + t * t; }
                        const size_t pi = (size_t)y * w + x;
                        if (d < dist[pi]) { dist[pi] = d; s->labels[pi] = k; }
                    }
            }
        }
        /* update: exact fixed-point sums (order free) */
        memset(sum, 0, sizeof(long long) * (size_t)K * (D + 4));
        float f[LSC_MAXD];
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                const float Wt = lsc_features(s, x, y, f);
                long long* a = sum + (size_t)s->labels[(size_t)y * w + x] * (D + 4);
                for (int i = 0; i < D; i++) a[i] += (long long)((Wt * f[i]) * 4294967296.0f);
                a[D] += (long long)(Wt * 16777216.0f);
                a[D + 1] += x; a[D + 2] += y; a[D + 3] += 1;
            }
        for (int k = 0; k < K; k++) {
            const long long* a = sum + (size_t)k * (D + 4);
            const long long cnt = a[D + 3];
            for (int i = 0; i < D; i++)
                s->cen[(size_t)k * D + i] = (cnt > 0 && a[D] != 0) ? (float)((double)a[i] / (double)a[D] / 256.0) : 0.f;
            s->sx[k] = cnt > 0 ? (int)(a[D + 1] / cnt) : 0;
            s->sy[k] = cnt > 0 ? (int)(a[D + 2] / cnt) : 0;
        }
    }
    free(dist); free(sum);
}

int spxo_lsc_get_number_of_superpixels(const spxo_lsc* s) { return s->K; }
void spxo_lsc_get_labels(const spxo_lsc* s, int32_t* dst) { memcpy(dst, s->labels, sizeof(int32_t) * (size_t)s->w * s->h); }
void spxo_lsc_set_labels(spxo_lsc* s, const int32_t* src, int numlabels)
{
    memcpy(s->labels, src, sizeof(int32_t) * (size_t)s->w * s->h);
    if (numlabels > 0) s->K = numlabels;
}
void spxo_lsc_get_centres(const spxo_lsc* s, float* dst)   /* K rows of (D features, seed x, seed y) */
{
    for (int k = 0; k < s->K; k++) {
        for (int i = 0; i < s->D; i++) dst[(size_t)k * (s->D + 2) + i] = s->cen[(size_t)k * s->D + i];
        dst[(size_t)k * (s->D + 2) + s->D] = (float)s->sx[k];
        dst[(size_t)k * (s->D + 2) + s->D + 1] = (float)s->sy[k];
    }
}
void spxo_lsc_get_label_contour_mask(const spxo_lsc* s, uint8_t* dst, int thick_line)
{
    spxo_label_contour_mask(s->labels, s->w, s->h, thick_line, dst);   /* same routine as SuperpixelSLIC */
}

/* ---- connectivity ------------------------------------------------------------------------------------------- */
static const int dx8[8] = { -1, -1, 0, 1, 1, 1, 0, -1 };
static const int dy8[8] = { 0, -1, -1, -1, 0, 1, 1, 1 };
static const int dx4[4] = { -1, 0, 1, 0 };
static const int dy4[4] = { 0, -1, 0, 1 };

static void lsc_pre_enforce(spxo_lsc* s, int bound)
{
    const int w = s->w, h = s->h;
    const size_t N = (size_t)w * h;
    uint8_t* mask = (uint8_t*)calloc(N, 1);
    int* qx = (int*)malloc(sizeof(int) * N); int* qy = (int*)malloc(sizeof(int) * N);
    int adj = 0;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            if (mask[(size_t)y * w + x]) continue;
            const int L = s->labels[(size_t)y * w + x];
            for (int k = 0; k < 8; k++) {
                const int nx = x + dx8[k], ny = y + dy8[k];
                if (nx >= 0 && nx < w && ny >= 0 && ny < h) {
                    if (mask[(size_t)ny * w + nx] && s->labels[(size_t)ny * w + nx] != L) adj = s->labels[(size_t)ny * w + nx];
                    break;                         /* unconditional, as in the authors' code */
                }
            }
            mask[(size_t)y * w + x] = 1;
            int head = 0, tail = 0;
            qx[tail] = x; qy[tail] = y; tail++;
            while (head < tail) {
                const int cx = qx[head], cy = qy[head]; head++;
                for (int k = 0; k < 8; k++) {
                    const int nx = cx + dx8[k], ny = cy + dy8[k];
                    if (nx >= 0 && nx < w && ny >= 0 && ny < h && !mask[(size_t)ny * w + nx] && s->labels[(size_t)ny * w + nx] == L) {
                        mask[(size_t)ny * w + nx] = 1; qx[tail] = nx; qy[tail] = ny; tail++;
                    }
                }
            }
            if (tail <= bound)
                for (int i = 0; i < tail; i++) s->labels[(size_t)qy[i] * w + qx[i]] = adj;
        }
    free(mask); free(qx); free(qy);
}

static unsigned lsc_prio(int r)
{
    unsigned x = (unsigned)r;
    x *= 2654435761u; x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
    return x;
}
static int g_prio_order = 0;
static int cmp_prio(const void* a, const void* b)
{
    const unsigned pa = lsc_prio(*(const int*)a), pb = lsc_prio(*(const int*)b);
    return pa < pb ? -1 : (pa > pb ? 1 : 0);
}
void spxo_lsc_set_merge_order(spxo_lsc* s, int order) { s->merge_order = order ? 1 : 0; }
void spxo_lsc_set_merge_algo(spxo_lsc* s, int algo) { s->merge_algo = algo ? 1 : 0; }
int spxo_lsc_merge_iterations(const spxo_lsc* s) { return s->merge_iters; }

/* ---- fixed-point evaluation of the sequential merge (visiting order 0 only) ------------------------------------------
 * The sequential loop visits region i at "time" i.  Its decision (merge or not, and into whom) is a function of the
 * decisions of the regions j < i only.  So: take ANY table of tentative decisions mg[]/tg[], derive from it -- for every
 * region and every time -- representative, centre, weight and size, re-take every decision against that state, and
 * repeat until no decision changes.  After k rounds the first k regions (at least) hold their sequential decision, and a
 * table that reproduces itself IS the sequential result (induction over i).  Every step of a round is data parallel.
 *   children of x  = { c : mg[c] && tg[c] == x }, ascending c; c hops into x at time c
 *   T_x(t)         = orig_x folded with M_c for the children c < t in ascending order (the sequential float update)
 *   M_x            = T_x(x), the state x has when it is visited (depends on indices < x only: well founded)
 *   rep_t(p)       = follow p -> tg[p] while the node was visited (index < t), merged, and AFTER it was entered
 *                    (index > arrival time; in a consistent table that always holds, in a tentative one it cuts cycles)
 *   decision of x  = small at time x (M_x.size < threshold) and argmin over edges (p,q), rep_x(p) == x, y = rep_x(q) != x
 *                    of (|M_x.cen - T_y(x).cen|^2, y)
 */
typedef struct { float c[LSC_MAXD]; float W; long long size; } lsc_rstate;
static void lsc_fold(lsc_rstate* a, const lsc_rstate* b, int D)
{
    for (int j = 0; j < D; j++) { const float u = a->c[j] * a->W, v = b->c[j] * b->W; a->c[j] = (u + v) / (a->W + b->W); }
    a->W = a->W + b->W;
    a->size += b->size;
}
static int cmp_u64(const void* a, const void* b)
{
    const unsigned long long x = *(const unsigned long long*)a, y = *(const unsigned long long*)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}
/* returns the number of rounds; fills cur[] (final owner) and alive[] */
static int lsc_merge_fixedpoint(int R, int D, const float* cen0, const float* W0, const long long* size0, int threshold,
                                const unsigned long long* edges, size_t E, int* cur, uint8_t* alive)
{
    uint8_t* mg = (uint8_t*)calloc(R > 0 ? R : 1, 1);
    int* tg = (int*)malloc(sizeof(int) * (R > 0 ? R : 1));
    int* off = (int*)malloc(sizeof(int) * (R + 1));
    int* ch = (int*)malloc(sizeof(int) * (R > 0 ? R : 1));
    int* pre = (int*)malloc(sizeof(int) * (R > 0 ? R : 1));
    lsc_rstate* S = (lsc_rstate*)malloc(sizeof(lsc_rstate) * (R > 0 ? R : 1));     /* S[off[x]+k] = T_x after k+1 children */
    lsc_rstate* orig = (lsc_rstate*)malloc(sizeof(lsc_rstate) * (R > 0 ? R : 1));
    unsigned long long* best = (unsigned long long*)malloc(sizeof(unsigned long long) * (R > 0 ? R : 1));
    for (int r = 0; r < R; r++) {
        for (int j = 0; j < D; j++) orig[r].c[j] = cen0[(size_t)r * D + j];
        orig[r].W = W0[r]; orig[r].size = size0[r]; tg[r] = -1;
    }
#define T_STATE(y, k) ((k) == 0 ? &orig[y] : &S[off[y] + (k) - 1])
    int rounds = 0;
    for (;;) {
        rounds++;
        /* children lists (ascending child index by construction) */
        for (int r = 0; r <= R; r++) off[r] = 0;
        for (int c = 0; c < R; c++) if (mg[c]) off[tg[c] + 1]++;
        for (int r = 0; r < R; r++) off[r + 1] += off[r];
        {
            int* fillp = (int*)malloc(sizeof(int) * (R > 0 ? R : 1));
            for (int r = 0; r < R; r++) fillp[r] = off[r];
            for (int c = 0; c < R; c++) if (mg[c]) ch[fillp[tg[c]]++] = c;
            free(fillp);
        }
        /* trajectories: children < x first (ascending x: M of smaller indices is complete), the rest afterwards */
        for (int x = 0; x < R; x++) {
            lsc_rstate st = orig[x];
            int k = 0;
            for (int s = off[x]; s < off[x + 1] && ch[s] < x; s++, k++) { const int c = ch[s]; lsc_fold(&st, T_STATE(c, pre[c]), D); S[s] = st; }
            pre[x] = k;
        }
        for (int x = 0; x < R; x++) {
            if (off[x] + pre[x] == off[x + 1]) continue;
            lsc_rstate st = *T_STATE(x, pre[x]);
            for (int s = off[x] + pre[x]; s < off[x + 1]; s++) { const int c = ch[s]; lsc_fold(&st, T_STATE(c, pre[c]), D); S[s] = st; }
        }
        /* candidates */
        for (int r = 0; r < R; r++) best[r] = ~0ull;
        for (size_t e = 0; e < E; e++)
            for (int dir = 0; dir < 2; dir++) {
                const int a = dir ? (int)(edges[e] & 0xffffffffu) : (int)(edges[e] >> 32);
                const int b = dir ? (int)(edges[e] >> 32) : (int)(edges[e] & 0xffffffffu);
                int x = a, tarr = -1;
                for (;;) {
                    if (x > tarr && size0[x] < threshold) {            /* x is visited at time x with a among its members */
                        int y = b, ty = -1;
                        while (y < x && mg[y] && y > ty) { ty = y; y = tg[y]; }
                        if (y != x) {
                            /* children of y with index < x */
                            int lo = off[y], hi = off[y + 1];
                            while (lo < hi) { const int mid = (lo + hi) >> 1; if (ch[mid] < x) lo = mid + 1; else hi = mid; }
                            const lsc_rstate* sy = T_STATE(y, lo - off[y]);
                            const lsc_rstate* sx = T_STATE(x, pre[x]);
                            float d = 0.f;
                            for (int j = 0; j < D; j++) { const float t = sx->c[j] - sy->c[j]; d = d + t * t; }
                            unsigned db; memcpy(&db, &d, 4);
                            const unsigned long long key = ((unsigned long long)db << 32) | (unsigned)y;
                            if (key < best[x]) best[x] = key;
                        }
                    }
                    if (mg[x] && x > tarr) { tarr = x; x = tg[x]; } else break;
                }
            }
        /* decisions */
        int changed = 0;
        for (int x = 0; x < R; x++) {
            if (size0[x] >= threshold) continue;
            const int m = T_STATE(x, pre[x])->size < threshold && best[x] != ~0ull;
            const int t = m ? (int)(unsigned)(best[x] & 0xffffffffu) : -1;
            if (m != mg[x] || t != tg[x]) changed++;
            mg[x] = (uint8_t)m; tg[x] = t;
        }
        if (getenv("SPXO_FP_TRACE")) fprintf(stderr, "round %d changed %d\n", rounds, changed);
        if (!changed) break;
    }
#undef T_STATE
    for (int r = 0; r < R; r++) {
        int y = r, ty = -1;
        while (mg[y] && y > ty) { ty = y; y = tg[y]; }
        cur[r] = y; alive[r] = !mg[r];
    }
    free(mg); free(tg); free(off); free(ch); free(pre); free(S); free(orig); free(best);
    return rounds;
}

static void lsc_post_enforce(spxo_lsc* s, int threshold)
{
    const int w = s->w, h = s->h, D = s->D;
    const size_t N = (size_t)w * h;
    int32_t* reg = (int32_t*)malloc(sizeof(int32_t) * N);
    for (size_t i = 0; i < N; i++) reg[i] = -1;
    int* qx = (int*)malloc(sizeof(int) * N); int* qy = (int*)malloc(sizeof(int) * N);
    /* 4-connected regions in raster order, with their pixel lists (qx/qy are filled region after region) */
    int R = 0;
    size_t cap = 1024, fill = 0;
    size_t* off = (size_t*)malloc(sizeof(size_t) * (cap + 1));
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            if (reg[(size_t)y * w + x] >= 0) continue;
            if ((size_t)R + 1 >= cap) { cap *= 2; off = (size_t*)realloc(off, sizeof(size_t) * (cap + 1)); }
            const int L = s->labels[(size_t)y * w + x];
            size_t head = fill;
            off[R] = fill;
            reg[(size_t)y * w + x] = R; qx[fill] = x; qy[fill] = y; fill++;
            while (head < fill) {
                const int cx = qx[head], cy = qy[head]; head++;
                for (int k = 0; k < 4; k++) {
                    const int nx = cx + dx4[k], ny = cy + dy4[k];
                    if (nx >= 0 && nx < w && ny >= 0 && ny < h && reg[(size_t)ny * w + nx] < 0 && s->labels[(size_t)ny * w + nx] == L) {
                        reg[(size_t)ny * w + nx] = R; qx[fill] = nx; qy[fill] = ny; fill++;
                    }
                }
            }
            R++;
        }
    off[R] = fill;
    /* region statistics */
    float* cen = (float*)malloc(sizeof(float) * (size_t)R * D);
    float* Wr = (float*)malloc(sizeof(float) * R);
    long long* size = (long long*)malloc(sizeof(long long) * R);
    int* cur = (int*)malloc(sizeof(int) * R);       /* current owner of an original region */
    int* next = (int*)malloc(sizeof(int) * R);      /* member lists of the current regions */
    int* tail = (int*)malloc(sizeof(int) * R);
    uint8_t* alive = (uint8_t*)malloc(R);
    float f[LSC_MAXD];
    for (int r = 0; r < R; r++) {
        long long sum[LSC_MAXD] = {0}, sw = 0;
        for (size_t i = off[r]; i < off[r + 1]; i++) {
            const float Wt = lsc_features(s, qx[i], qy[i], f);
            for (int j = 0; j < D; j++) sum[j] += (long long)((Wt * f[j]) * 4294967296.0f);
            sw += (long long)(Wt * 16777216.0f);
        }
        for (int j = 0; j < D; j++) cen[(size_t)r * D + j] = sw != 0 ? (float)((double)sum[j] / (double)sw / 256.0) : 0.f;
        Wr[r] = (float)((double)sw / 16777216.0);
        size[r] = (long long)(off[r + 1] - off[r]);
        cur[r] = r; next[r] = -1; tail[r] = r; alive[r] = 1;
    }
    if (s->merge_algo && !s->merge_order) {
        /* region adjacency: unique (min, max) pairs of 4-neighbour pixels */
        size_t ecap = 4 * N + 4, E = 0;
        unsigned long long* ek = (unsigned long long*)malloc(sizeof(unsigned long long) * ecap);
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                const int a = reg[(size_t)y * w + x];
                if (x + 1 < w) { const int b = reg[(size_t)y * w + x + 1]; if (a != b) ek[E++] = ((unsigned long long)(a < b ? a : b) << 32) | (unsigned)(a < b ? b : a); }
                if (y + 1 < h) { const int b = reg[(size_t)(y + 1) * w + x]; if (a != b) ek[E++] = ((unsigned long long)(a < b ? a : b) << 32) | (unsigned)(a < b ? b : a); }
            }
        qsort(ek, E, sizeof(unsigned long long), cmp_u64);
        size_t U = 0;
        for (size_t e = 0; e < E; e++) if (e == 0 || ek[e] != ek[e - 1]) ek[U++] = ek[e];
        s->merge_iters = lsc_merge_fixedpoint(R, D, cen, Wr, size, threshold, ek, U, cur, alive);
        free(ek);
        int* newid2 = (int*)malloc(sizeof(int) * R);
        int K2 = 0;
        for (int r = 0; r < R; r++) newid2[r] = alive[r] ? K2++ : -1;
        for (size_t i = 0; i < N; i++) s->labels[i] = newid2[cur[reg[i]]];
        s->K = K2;
        free(newid2);
        free(reg); free(qx); free(qy); free(off); free(cen); free(Wr); free(size); free(cur); free(next); free(tail); free(alive);
        return;
    }
    /* sequential merge of small regions: in region order (order 0), or in the order of a bijective hash of the region
       index (order 1: same rule, but the dependency chains between merges are short, which is what a GPU needs) */
    int* nb = (int*)malloc(sizeof(int) * 64);
    int nbcap = 64;
    int* visit = (int*)malloc(sizeof(int) * (R > 0 ? R : 1));
    for (int r = 0; r < R; r++) visit[r] = r;
    if (s->merge_order) {
        g_prio_order = 1;
        qsort(visit, R, sizeof(int), cmp_prio);
    }
    for (int vi = 0; vi < R; vi++) {
        const int i = visit[vi];
        if (size[i] >= threshold) continue;
        int nnb = 0;
        for (int m = i; m >= 0; m = next[m])
            for (size_t p = off[m]; p < off[m + 1]; p++)
                for (int k = 0; k < 4; k++) {
                    const int nx = qx[p] + dx4[k], ny = qy[p] + dy4[k];
                    if (nx < 0 || nx >= w || ny < 0 || ny >= h) continue;
                    const int r = cur[reg[(size_t)ny * w + nx]];
                    if (r == i) continue;
                    int seen = 0;
                    for (int q = 0; q < nnb; q++) if (nb[q] == r) { seen = 1; break; }
                    if (!seen) {
                        if (nnb == nbcap) { nbcap *= 2; nb = (int*)realloc(nb, sizeof(int) * nbcap); }
                        nb[nnb++] = r;
                    }
                }
        if (nnb == 0) continue;
        int best = -1;
        float bestd = 0.f;
        for (int q = 0; q < nnb; q++) {
            float d = 0.f;
            for (int j = 0; j < D; j++) { float t = cen[(size_t)i * D + j] - cen[(size_t)nb[q] * D + j]; d = d + t * t; }
            if (best < 0 || d < bestd || (d == bestd && nb[q] < best)) { best = nb[q]; bestd = d; }
        }
        for (int j = 0; j < D; j++) {
            const float a = cen[(size_t)best * D + j] * Wr[best], b = cen[(size_t)i * D + j] * Wr[i];
            cen[(size_t)best * D + j] = (a + b) / (Wr[best] + Wr[i]);
        }
        Wr[best] = Wr[best] + Wr[i];
        size[best] += size[i];
        for (int m = i; m >= 0; m = next[m]) cur[m] = best;
        next[tail[best]] = i; tail[best] = tail[i];
        alive[i] = 0;
    }
    /* compact ids in ascending region order */
    int* newid = (int*)malloc(sizeof(int) * R);
    int K = 0;
    for (int r = 0; r < R; r++) newid[r] = alive[r] ? K++ : -1;
    for (size_t i = 0; i < N; i++) s->labels[i] = newid[cur[reg[i]]];
    s->K = K;
    free(reg); free(qx); free(qy); free(off); free(cen); free(Wr); free(size); free(cur); free(next); free(tail);
    free(alive); free(nb); free(newid); free(visit);
}

void spxo_lsc_enforce_label_connectivity(spxo_lsc* s, int min_element_size)
{
    if (min_element_size <= 0) return;
    int threshold = (s->w * s->h) / (s->K * 4);
    lsc_pre_enforce(s, min_element_size);
    lsc_post_enforce(s, threshold);
}
