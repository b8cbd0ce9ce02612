/* spx_oracle.c -- CPU ORACLE (test infrastructure, NOT product code). See spx_oracle.h.
 *
 * Colour conversion, the SLIC family (SLIC=100, SLICO=101, MSLIC=102), label-connectivity
 * enforcement and the two contour-mask flavours.  Restates the published algorithm of
 * OpenCV imgproc (cvtColor 8-bit) and opencv_contrib ximgproc/src/slic.cpp (version "4.1",
 * README.md:46; not vendored).  Normative pseudo-code: SURVEY.md Appendix D; every step there
 * is fixture-validated for SLIC=100.  Compile with -ffp-contract=off (no FMA contraction):
 * the label maps are sensitive to single-ulp changes in the distance arithmetic.
 */
#include "spx_oracle.h"
#include "spx_color_tables.h"
#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static int g_threads = 1;
void spxo_set_threads(int n) { g_threads = n < 1 ? 1 : n; }
int  spxo_get_threads(void) { return g_threads; }

/* ------------------------------------------------------------------------------------------
 * cv::cvtColor 8-bit BGR->Lab (mainwindow.cpp:2096): integer LUT pipeline of OpenCV's RGB2Lab_b.
 * ------------------------------------------------------------------------------------------ */
static inline int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }
static inline uint8_t sat8(int v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }

void spxo_bgr2lab8(const uint8_t* src, size_t sstride, uint8_t* dst, size_t dstride, int w, int h)
{
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (int y = 0; y < h; y++) {
        const uint8_t* s = src + (size_t)y * sstride;
        uint8_t* d = dst + (size_t)y * dstride;
        for (int x = 0; x < w; x++) {
            int B = SPX_TAB_GAMMA[s[3 * x]], G = SPX_TAB_GAMMA[s[3 * x + 1]], R = SPX_TAB_GAMMA[s[3 * x + 2]];
            int fX = SPX_TAB_CBRT[descale(R * 1777 + G * 1541 + B * 778, 12)];
            int fY = SPX_TAB_CBRT[descale(R * 871 + G * 2929 + B * 296, 12)];
            int fZ = SPX_TAB_CBRT[descale(R * 73 + G * 448 + B * 3575, 12)];
            d[3 * x]     = sat8(descale(296 * fY - 1336934, 15));
            d[3 * x + 1] = sat8(descale(500 * (fX - fY) + 128 * 32768, 15));
            d[3 * x + 2] = sat8(descale(200 * (fY - fZ) + 128 * 32768, 15));
        }
    }
}

/* cv::cvtColor 8-bit BGR->HSV (mainwindow.cpp:2097): H in 0..179 */
void spxo_bgr2hsv8(const uint8_t* src, size_t sstride, uint8_t* dst, size_t dstride, int w, int h)
{
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (int y = 0; y < h; y++) {
        const uint8_t* s = src + (size_t)y * sstride;
        uint8_t* d = dst + (size_t)y * dstride;
        for (int x = 0; x < w; x++) {
            int b = s[3 * x], g = s[3 * x + 1], r = s[3 * x + 2];
            int v = b > g ? b : g; if (r > v) v = r;
            int mn = b < g ? b : g; if (r < mn) mn = r;
            int df = v - mn;
            int sat = (df * SPX_TAB_SDIV[v] + (1 << 11)) >> 12;
            int hh = (v == r) ? (g - b) : (v == g) ? (b - r + 2 * df) : (r - g + 4 * df);
            hh = (hh * SPX_TAB_HDIV[df] + (1 << 11)) >> 12;
            if (hh < 0) hh += 180;
            d[3 * x] = (uint8_t)hh; d[3 * x + 1] = (uint8_t)sat; d[3 * x + 2] = (uint8_t)v;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * SuperpixelSLIC
 * ------------------------------------------------------------------------------------------ */
#define SPXO_MAXC 4
struct spxo_slic {
    int w, h, C, algorithm, S;
    float ruler;
    int K;
    int K0;                  /* seeds at creation (MSLIC growth cap = 4*K0) */
    uint8_t* ch[SPXO_MAXC];  /* planes (cv::split) */
    int32_t* labels;
    float* kc[SPXO_MAXC];    /* centre colour per channel */
    float* kx; float* ky;
    float* adaptk;           /* MSLIC */
    int cap;                 /* capacity of the centre arrays */
};

static void slic_reserve(spxo_slic* s, int cap)
{
    if (cap <= s->cap) return;
    for (int c = 0; c < s->C; c++) s->kc[c] = (float*)realloc(s->kc[c], sizeof(float) * cap);
    s->kx = (float*)realloc(s->kx, sizeof(float) * cap);
    s->ky = (float*)realloc(s->ky, sizeof(float) * cap);
    s->adaptk = (float*)realloc(s->adaptk, sizeof(float) * cap);
    s->cap = cap;
}

static inline int refl101(int p, int n) /* BORDER_REFLECT_101 */
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) { if (p < 0) p = -p; else p = 2 * (n - 1) - p; }
    return p;
}

/* DetectChEdges at one pixel: sum over channels of (Sobel_x,k=1)^2 + (Sobel_y,k=1)^2, float32 */
static float edge_mag(const spxo_slic* s, int x, int y)
{
    int xm = refl101(x - 1, s->w), xp = refl101(x + 1, s->w);
    int ym = refl101(y - 1, s->h), yp = refl101(y + 1, s->h);
    float e = 0.f;
    for (int c = 0; c < s->C; c++) {
        const uint8_t* p = s->ch[c];
        float dx = (float)p[(size_t)y * s->w + xp] - (float)p[(size_t)y * s->w + xm];
        float dy = (float)p[(size_t)yp * s->w + x] - (float)p[(size_t)ym * s->w + x];
        e = e + (dx * dx + dy * dy);
    }
    return e;
}

static void slic_seed_color(spxo_slic* s, int n, int X, int Y)
{
    for (int c = 0; c < s->C; c++) s->kc[c][n] = (float)s->ch[c][(size_t)Y * s->w + X];
}

spxo_slic* spxo_slic_create(const uint8_t* img, size_t stride, int w, int h, int channels,
                            int algorithm, int region_size, float ruler)
{
    if (!img || w <= 0 || h <= 0 || channels < 1 || channels > SPXO_MAXC || region_size < 1) return NULL;
    if (algorithm < 100 || algorithm > 102) return NULL;
    spxo_slic* s = (spxo_slic*)calloc(1, sizeof(*s));
    s->w = w; s->h = h; s->C = channels; s->algorithm = algorithm; s->S = region_size; s->ruler = ruler;
    size_t N = (size_t)w * h;
    for (int c = 0; c < channels; c++) {
        s->ch[c] = (uint8_t*)malloc(N);
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) s->ch[c][(size_t)y * w + x] = img[(size_t)y * stride + (size_t)x * channels + c];
    }
    s->labels = (int32_t*)calloc(N, sizeof(int32_t));
    const int S = region_size;
    if (algorithm == 101) {
        /* GetChSeedsK: hexagonal grid, rows Y=r*S+S/2, X=x*S+((S/2)<<(r&1)) */
        int n = 0, r = 0;
        for (int pass = 0; pass < 2; pass++) {
            n = 0; r = 0;
            for (int y = 0; ; y++) {
                int Y = y * S + S / 2;
                if (Y > h - 1) break;
                for (int x = 0; ; x++) {
                    int X = x * S + ((S / 2) << (r & 1));
                    if (X > w - 1) break;
                    if (pass) { s->kx[n] = (float)X; s->ky[n] = (float)Y; slic_seed_color(s, n, X, Y); }
                    n++;
                }
                r++;
            }
            if (!pass) slic_reserve(s, n > 0 ? n : 1);
        }
        s->K = n;
    } else {
        /* GetChSeedsS: strip grid with truncated per-strip error */
        int xstrips = (int)(0.5f + (float)w / (float)S);
        int ystrips = (int)(0.5f + (float)h / (float)S);
        int xerr = w - S * xstrips, yerr = h - S * ystrips;
        float xerrperstrip = (float)xerr / (float)xstrips;
        float yerrperstrip = (float)yerr / (float)ystrips;
        int xoff = S / 2, yoff = S / 2;
        int K = xstrips * ystrips;
        slic_reserve(s, K > 0 ? K : 1);
        int n = 0;
        for (int y = 0; y < ystrips; y++) {
            int ye = y * (int)yerrperstrip;
            int Y = y * S + yoff + ye;
            if (Y > h - 1) Y = h - 1;
            for (int x = 0; x < xstrips; x++) {
                int xe = x * (int)xerrperstrip;
                int X = x * S + xoff + xe;
                if (X > w - 1) X = w - 1;
                s->kx[n] = (float)X; s->ky[n] = (float)Y; slic_seed_color(s, n, X, Y);
                n++;
            }
        }
        s->K = K;
    }
    /* PerturbSeeds: move to the strict-min edge magnitude of the 8 neighbours, only when both
       coordinates changed (upstream quirk, fixture-validated) */
    static const int dx8[8] = { -1, -1, 0, 1, 1, 1, 0, -1 };
    static const int dy8[8] = { 0, -1, -1, -1, 0, 1, 1, 1 };
    for (int n = 0; n < s->K; n++) {
        int ox = (int)s->kx[n], oy = (int)s->ky[n];
        int sx = ox, sy = oy;
        float best = edge_mag(s, ox, oy);
        for (int i = 0; i < 8; i++) {
            int nx = ox + dx8[i], ny = oy + dy8[i];
            if (nx >= 0 && nx < w && ny >= 0 && ny < h) {
                float e = edge_mag(s, nx, ny);
                if (e < best) { best = e; sx = nx; sy = ny; }
            }
        }
        if (sx != ox && sy != oy) {
            s->kx[n] = (float)sx; s->ky[n] = (float)sy; slic_seed_color(s, n, sx, sy);
        }
    }
    for (int n = 0; n < s->K; n++) s->adaptk[n] = 1.0f;
    s->K0 = s->K;
    return s;
}

void spxo_slic_destroy(spxo_slic* s)
{
    if (!s) return;
    for (int c = 0; c < SPXO_MAXC; c++) { free(s->ch[c]); free(s->kc[c]); }
    free(s->labels); free(s->kx); free(s->ky); free(s->adaptk); free(s);
}

/* centre update: per-label sums of channels, x, y, count -> means (true division, float32).
 * threads==1: float32 accumulators in column-major order as upstream (exact below 2^24);
 * threads>1 : exact int64 partial sums per thread, converted once (identical below 2^24). */
static void slic_update_centres(spxo_slic* s)
{
    const int K = s->K, C = s->C, w = s->w, h = s->h;
    if (g_threads <= 1) {
        float* sig = (float*)calloc((size_t)K * (C + 2), sizeof(float));
        int* cnt = (int*)calloc(K, sizeof(int));
        for (int x = 0; x < w; x++)
            for (int y = 0; y < h; y++) {
                int l = s->labels[(size_t)y * w + x];
                for (int c = 0; c < C; c++) sig[(size_t)c * K + l] += (float)s->ch[c][(size_t)y * w + x];
                sig[(size_t)C * K + l] += (float)x;
                sig[(size_t)(C + 1) * K + l] += (float)y;
                cnt[l]++;
            }
        for (int k = 0; k < K; k++) {
            int m = cnt[k] <= 0 ? 1 : cnt[k];
            for (int c = 0; c < C; c++) s->kc[c][k] = sig[(size_t)c * K + k] / (float)m;
            s->kx[k] = sig[(size_t)C * K + k] / (float)m;
            s->ky[k] = sig[(size_t)(C + 1) * K + k] / (float)m;
        }
        free(sig); free(cnt);
    } else {
        int T = g_threads;
        int64_t* part = (int64_t*)calloc((size_t)T * K * (C + 3), sizeof(int64_t));
#pragma omp parallel num_threads(T)
        {
#ifdef _OPENMP
            int t = omp_get_thread_num();
#else
            int t = 0;
#endif
            int64_t* p = part + (size_t)t * K * (C + 3);
#pragma omp for schedule(static)
            for (int y = 0; y < h; y++)
                for (int x = 0; x < w; x++) {
                    int l = s->labels[(size_t)y * w + x];
                    for (int c = 0; c < C; c++) p[(size_t)c * K + l] += s->ch[c][(size_t)y * w + x];
                    p[(size_t)C * K + l] += x;
                    p[(size_t)(C + 1) * K + l] += y;
                    p[(size_t)(C + 2) * K + l] += 1;
                }
        }
        for (int k = 0; k < K; k++) {
            int64_t acc[SPXO_MAXC + 3] = { 0 };
            for (int t = 0; t < T; t++)
                for (int f = 0; f < C + 3; f++) acc[f] += part[(size_t)t * K * (C + 3) + (size_t)f * K + k];
            int64_t m = acc[C + 2] <= 0 ? 1 : acc[C + 2];
            for (int c = 0; c < C; c++) s->kc[c][k] = (float)acc[c] / (float)m;
            s->kx[k] = (float)acc[C] / (float)m;
            s->ky[k] = (float)acc[C + 1] / (float)m;
        }
        free(part);
    }
}

/* One assignment sweep in upstream's cluster-ordered scatter form, restricted to rows [r0,r1)
 * so that row bands can run on different threads with identical results (per pixel the
 * candidate clusters are still visited in ascending n, strict '<').
 * mode 0 = SLIC / MSLIC (adaptk==NULL for SLIC), 1 = SLICO. */
static void slic_assign_band(spxo_slic* s, int r0, int r1, float* distvec, float xywt,
                             float* distchans, float* distxy, const float* maxchans)
{
    const int K = s->K, C = s->C, w = s->w, h = s->h, S = s->S;
    const int slico = s->algorithm == 101, mslic = s->algorithm == 102;
    for (int n = 0; n < K; n++) {
        int off = S;
        if (mslic) off = (int)((float)S * s->adaptk[n]);
        int iy = (int)s->ky[n], ix = (int)s->kx[n];
        int y1 = iy - off; if (y1 < 0) y1 = 0;
        int y2 = iy + off; if (y2 > h) y2 = h;
        int x1 = ix - off; if (x1 < 0) x1 = 0;
        int x2 = ix + off; if (x2 > w) x2 = w;
        if (y1 < r0) y1 = r0;
        if (y2 > r1) y2 = r1;
        const float kxn = s->kx[n], kyn = s->ky[n];
        for (int y = y1; y < y2; y++) {
            for (int x = x1; x < x2; x++) {
                size_t i = (size_t)y * w + x;
                float dc = 0.f;
                for (int c = 0; c < C; c++) {
                    float t = (float)s->ch[c][i] - s->kc[c][n];
                    dc = dc + t * t;
                }
                float ex = (float)x - kxn, ey = (float)y - kyn;
                float dxy = ex * ex + ey * ey;
                float d;
                if (slico) {
                    distchans[i] = dc; distxy[i] = dxy;     /* last visiting cluster's values stay */
                    d = dc / maxchans[n] + dxy / xywt;
                } else {
                    d = dc + dxy / xywt;
                    if (mslic) d = d / s->adaptk[n];
                }
                if (d < distvec[i]) { distvec[i] = d; s->labels[i] = n; }
            }
        }
    }
}

static void slic_assign(spxo_slic* s, float* distvec, float xywt, float* distchans, float* distxy,
                        const float* maxchans)
{
    size_t N = (size_t)s->w * s->h;
    for (size_t i = 0; i < N; i++) distvec[i] = FLT_MAX;
    int T = g_threads;
    if (T <= 1) { slic_assign_band(s, 0, s->h, distvec, xywt, distchans, distxy, maxchans); return; }
    int bands = T * 4;
#pragma omp parallel for num_threads(T) schedule(dynamic, 1)
    for (int b = 0; b < bands; b++) {
        int r0 = (int)((int64_t)s->h * b / bands), r1 = (int)((int64_t)s->h * (b + 1) / bands);
        slic_assign_band(s, r0, r1, distvec, xywt, distchans, distxy, maxchans);
    }
}

static void mslic_adapt(spxo_slic* s);   /* spx_oracle.c, below */

void spxo_slic_iterate(spxo_slic* s, int num_iterations)
{
    size_t N = (size_t)s->w * s->h;
    float* distvec = (float*)malloc(N * sizeof(float));
    if (s->algorithm == 101) {
        /* PerformSLICO: dist = distchans/maxchans[n] + distxy/(S*S); maxchans/maxxy start at 100 / S*S,
           are reset to FLT_MIN after the first assignment, then track the per-label maxima of the
           distchans / distxy planes (which hold the LAST visiting cluster's value per pixel). */
        float* distchans = (float*)malloc(N * sizeof(float));
        float* distxy = (float*)malloc(N * sizeof(float));
        for (size_t i = 0; i < N; i++) { distchans[i] = FLT_MAX; distxy[i] = FLT_MAX; }
        float* maxchans = (float*)malloc(sizeof(float) * s->K);
        float* maxxy = (float*)malloc(sizeof(float) * s->K);
        /* start values as in Achanta's SLICO ("just start with 10"): 10*10 and S*S [recalled, UNPINNED] */
        for (int k = 0; k < s->K; k++) { maxchans[k] = 10.0f * 10.0f; maxxy[k] = (float)(s->S * s->S); }
        const float xywt = (float)(s->S * s->S);
        for (int itr = 0; itr < num_iterations; itr++) {
            slic_assign(s, distvec, xywt, distchans, distxy, maxchans);
            if (itr == 0) for (int k = 0; k < s->K; k++) { maxchans[k] = FLT_MIN; maxxy[k] = FLT_MIN; }
            for (size_t i = 0; i < N; i++) {
                int l = s->labels[i];
                if (maxchans[l] < distchans[i]) maxchans[l] = distchans[i];
                if (maxxy[l] < distxy[i]) maxxy[l] = distxy[i];
            }
            slic_update_centres(s);
        }
        free(distchans); free(distxy); free(maxchans); free(maxxy);
    } else {
        /* PerformSLIC / PerformMSLIC */
        const float q = (float)s->S / s->ruler;
        const float xywt = q * q;
        for (int itr = 0; itr < num_iterations; itr++) {
            slic_assign(s, distvec, xywt, NULL, NULL, NULL);
            slic_update_centres(s);
            if (s->algorithm == 102) mslic_adapt(s);
        }
    }
    free(distvec);
}

/* MSLIC content-adaptive step (PARITY UNPINNED: restates Liu et al., "Manifold SLIC", CVPR 2016,
 * in the structure SURVEY.md App. C.2 recalls for upstream's PerformMSLIC; constants split=4,
 * target=2 "from paper").  After each centre update:
 *   area(p)  = area of the unit square at p on the 2-manifold (x, y, q*c0, q*c1, q*c2), q = S/ruler,
 *              as two triangles (p,p+x,p+y) and (p+x+y,p+x,p+y); float32, Gram determinant;
 *   A[k]     = sum over pixels of label k of (int64)(area(p)*256) (fixed point, so the sum is
 *              order independent), Abar = float(sum A)/(K*256);
 *   adaptk[k]= clamp(sqrt(Abar / A[k]), 1/target, target)      (small windows where content is dense)
 *   split    : clusters with A[k] > split*Abar spawn 3 new centres at the quarter offsets of their
 *              window (K grows, appended in ascending k order); the parent moves to the 1st quarter.
 *              An iteration whose splits would take K above 4*K0 performs no split at all.
 */
static inline float tri_area5(const float* a, const float* b, const float* c)
{
    float uu = 0.f, vv = 0.f, uv = 0.f;
    for (int i = 0; i < 5; i++) {
        float u = b[i] - a[i], v = c[i] - a[i];
        uu = uu + u * u; vv = vv + v * v; uv = uv + u * v;
    }
    float g = uu * vv - uv * uv;
    if (g < 0.f) g = 0.f;
    return 0.5f * sqrtf(g);
}

static void mslic_adapt(spxo_slic* s)
{
    const int w = s->w, h = s->h, C = s->C, K = s->K, S = s->S;
    const float q = (float)S / s->ruler;
    int64_t* A = (int64_t*)calloc(K, sizeof(int64_t));
    int64_t tot = 0;
    for (int y = 0; y < h - 1; y++)
        for (int x = 0; x < w - 1; x++) {
            float P[4][5];
            for (int j = 0; j < 4; j++) {
                int xx = x + (j & 1), yy = y + (j >> 1);
                P[j][0] = (float)xx; P[j][1] = (float)yy;
                for (int c = 0; c < 3; c++)
                    P[j][2 + c] = c < C ? q * (float)s->ch[c][(size_t)yy * w + xx] : 0.f;
            }
            float a = tri_area5(P[0], P[1], P[2]) + tri_area5(P[3], P[1], P[2]);
            int64_t fx = (int64_t)(a * 256.0f);      /* fixed point: order-independent sums */
            A[s->labels[(size_t)y * w + x]] += fx;
            tot += fx;
        }
    const float target = 2.0f, split = 4.0f;
    float abar = (float)tot / ((float)K * 256.0f);
    int newK = K;
    for (int k = 0; k < K; k++) {
        float ak = (float)A[k] / 256.0f;
        float r = ak > 0.f ? sqrtf(abar / ak) : target;
        if (r < 1.0f / target) r = 1.0f / target;
        if (r > target) r = target;
        s->adaptk[k] = r;
        if (ak > split * abar) newK += 3;
    }
    if (newK > K && newK <= 4 * s->K0) {   /* growth cap: no splits once K would exceed 4*K0 */
        slic_reserve(s, newK);
        int n = K;
        for (int k = 0; k < K; k++) {
            if (!((float)A[k] / 256.0f > split * abar)) continue;
            float o = 0.5f * (float)S * s->adaptk[k];
            float cx = s->kx[k], cy = s->ky[k];
            const float ox[4] = { -o, o, -o, o }, oy[4] = { -o, -o, o, o };
            for (int j = 0; j < 4; j++) {
                float nx = cx + ox[j], ny = cy + oy[j];
                if (nx < 0.f) nx = 0.f; if (nx > (float)(w - 1)) nx = (float)(w - 1);
                if (ny < 0.f) ny = 0.f; if (ny > (float)(h - 1)) ny = (float)(h - 1);
                int dst = j == 0 ? k : n++;
                s->kx[dst] = nx; s->ky[dst] = ny;
                for (int c = 0; c < C; c++) s->kc[c][dst] = (float)s->ch[c][(size_t)(int)ny * w + (int)nx];
                s->adaptk[dst] = s->adaptk[k];
            }
        }
        s->K = newK;
    }
    free(A);
}

int spxo_slic_get_number_of_superpixels(const spxo_slic* s) { return s->K; }
void spxo_slic_get_labels(const spxo_slic* s, int32_t* dst) { memcpy(dst, s->labels, sizeof(int32_t) * (size_t)s->w * s->h); }
void spxo_slic_set_labels(spxo_slic* s, const int32_t* src, int numlabels)
{
    memcpy(s->labels, src, sizeof(int32_t) * (size_t)s->w * s->h);
    if (numlabels > 0) s->K = numlabels;   /* only meaningful before enforce / contour */
}
void spxo_slic_get_centres(const spxo_slic* s, float* dst)
{
    for (int k = 0; k < s->K; k++) {
        for (int c = 0; c < s->C; c++) dst[(size_t)k * (s->C + 2) + c] = s->kc[c][k];
        dst[(size_t)k * (s->C + 2) + s->C] = s->kx[k];
        dst[(size_t)k * (s->C + 2) + s->C + 1] = s->ky[k];
    }
}
void spxo_slic_enforce_label_connectivity(spxo_slic* s, int min_element_size)
{
    if (min_element_size == 0) return;
    s->K = spxo_enforce_label_connectivity(s->labels, s->w, s->h, s->K, min_element_size);
}
void spxo_slic_get_label_contour_mask(const spxo_slic* s, uint8_t* dst, int thick_line)
{
    spxo_label_contour_mask(s->labels, s->w, s->h, thick_line, dst);
}

/* ------------------------------------------------------------------------------------------
 * enforceLabelConnectivity (mainwindow.cpp:2108): raster scan, 4-connected flood fill per
 * component; components of <= min_sp_sz pixels take the label of the adjacent component found
 * at the seed pixel (neighbour order L,U,R,D, last already-relabelled one wins).
 * Returns the new number of labels.
 * ------------------------------------------------------------------------------------------ */
int spxo_enforce_label_connectivity(int32_t* labels, int w, int h, int numlabels, int min_element_size)
{
    if (min_element_size <= 0) return numlabels;
    static const int dx4[4] = { -1, 0, 1, 0 };
    static const int dy4[4] = { 0, -1, 0, 1 };
    const int sz = w * h;
    const int supsz = sz / (numlabels > 0 ? numlabels : 1);
    int div = (int)(100.0f / (float)min_element_size + 0.5f);
    if (div < 1) div = 1;
    int min_sp_sz = supsz / div; if (min_sp_sz < 3) min_sp_sz = 3;
    int32_t* nl = (int32_t*)malloc(sizeof(int32_t) * (size_t)sz);
    for (int i = 0; i < sz; i++) nl[i] = INT_MAX;
    int* xv = (int*)malloc(sizeof(int) * (size_t)sz);
    int* yv = (int*)malloc(sizeof(int) * (size_t)sz);
    int label = 0, adj = 0;
    for (int j = 0; j < h; j++)
        for (int k = 0; k < w; k++) {
            if (nl[(size_t)j * w + k] != INT_MAX) continue;
            nl[(size_t)j * w + k] = label;
            xv[0] = k; yv[0] = j;
            for (int n = 0; n < 4; n++) {
                int x = k + dx4[n], y = j + dy4[n];
                if (x >= 0 && x < w && y >= 0 && y < h && nl[(size_t)y * w + x] != INT_MAX) adj = nl[(size_t)y * w + x];
            }
            int count = 1;
            const int32_t old = labels[(size_t)j * w + k];
            for (int c = 0; c < count; c++)
                for (int n = 0; n < 4; n++) {
                    int x = xv[c] + dx4[n], y = yv[c] + dy4[n];
                    if (x >= 0 && x < w && y >= 0 && y < h) {
                        size_t i = (size_t)y * w + x;
                        if (nl[i] == INT_MAX && labels[i] == old) { xv[count] = x; yv[count] = y; nl[i] = label; count++; }
                    }
                }
            if (count <= min_sp_sz) {
                for (int c = 0; c < count; c++) nl[(size_t)yv[c] * w + xv[c]] = adj;
                label--;
            }
            label++;
        }
    memcpy(labels, nl, sizeof(int32_t) * (size_t)sz);
    free(nl); free(xv); free(yv);
    return label;
}

/* ------------------------------------------------------------------------------------------
 * getLabelContourMask, SLIC/LSC flavour (mainwindow.cpp:2111,2120): raster scan; a pixel is a
 * contour when more than line_width of its in-bounds 8-neighbours are not yet marked and carry
 * a different label.  line_width = thick_line ? 2 : 1.
 * ------------------------------------------------------------------------------------------ */
void spxo_label_contour_mask(const int32_t* labels, int w, int h, int thick_line, uint8_t* dst)
{
    static const int dx8[8] = { -1, -1, 0, 1, 1, 1, 0, -1 };
    static const int dy8[8] = { 0, -1, -1, -1, 0, 1, 1, 1 };
    const int lw = thick_line ? 2 : 1;
    memset(dst, 0, (size_t)w * h);
    for (int j = 0; j < h; j++)
        for (int k = 0; k < w; k++) {
            int np = 0;
            const int32_t me = labels[(size_t)j * w + k];
            for (int i = 0; i < 8; i++) {
                int x = k + dx8[i], y = j + dy8[i];
                if (x >= 0 && x < w && y >= 0 && y < h) {
                    size_t idx = (size_t)y * w + x;
                    if (!dst[idx] && labels[idx] != me) np++;
                }
            }
            if (np > lw) dst[(size_t)j * w + k] = 255;
        }
}

/* getLabelContourMask, SEEDS flavour (mainwindow.cpp:2127) [recalled; PARITY UNPINNED]:
 * count differing in-bounds 8-neighbours, a neighbour counts only if thick_line or it is not
 * yet marked; mark when the count exceeds 1. */
void spxo_label_contour_mask_seeds(const int32_t* labels, int w, int h, int thick_line, uint8_t* dst)
{
    static const int dx8[8] = { -1, -1, 0, 1, 1, 1, 0, -1 };
    static const int dy8[8] = { 0, -1, -1, -1, 0, 1, 1, 1 };
    memset(dst, 0, (size_t)w * h);
    for (int j = 0; j < h; j++)
        for (int k = 0; k < w; k++) {
            int np = 0;
            const int32_t me = labels[(size_t)j * w + k];
            for (int i = 0; i < 8; i++) {
                int x = k + dx8[i], y = j + dy8[i];
                if (x >= 0 && x < w && y >= 0 && y < h) {
                    size_t idx = (size_t)y * w + x;
                    if (labels[idx] != me && (thick_line || !dst[idx])) np++;
                }
            }
            if (np > 1) dst[(size_t)j * w + k] = 255;
        }
}
