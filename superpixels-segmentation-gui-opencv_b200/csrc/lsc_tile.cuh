// lsc_tile.cuh -- tiled form of the LSC assignment + exact accumulation (3 channels, steps >= LSC_TILE_MIN_STEP).
// Included by lsc.cu after LscDev / lsc_features.  Same results as lsc_assign_kernel, bit for bit (the sums are exact
// fixed-point integers and the arg-min is taken in (distance, seed index) order either way); ~5x fewer instructions.
//
// One CTA (128 threads) per 32 x LSC_TH (32) tile.
//   setup      : the kernels that move the seeds (lsc_bin_kernel, lsc_finalize_kernel) append every seed to the list of each
//                tile its window [s - step, s + step] meets (13 per tile at S = 20, 10 at S = 30); they are ranked by seed index (ascending
//                index + strict '<' = the oracle's (distance, index) order) and staged as one record each: the negated
//                centre, the key (k << 6) | slot, and two 32-entry penalty rows (0 inside the window, +inf outside; adding
//                0 to a distance >= 0 is exact, +inf can never win).
//   assignment : a warp owns a 16x8 patch per step, 2x2 pixels per thread; the 10 features are rebuilt in registers from the
//                tables; candidates whose window misses the patch are culled by one ballot; the distance runs on the
//                packed f32x2 pipe in the oracle's order ((((t0^2 + t1^2) + t2^2) + ...) + t9^2), never contracted.
//   sums       : the thread's 2x2 pixels are split by at most two winners and added, field by field, to the CTA's
//                per-candidate 64-bit accumulators (two fire-and-forget 32-bit shared-memory atomics: low 24 bits / signed
//                rest, no carry to wait for; 8 privatised copies picked by lane, so a warp sitting inside one superpixel is a
//                4-way, not a 32-way, conflict); one 64-bit global atomic per field per candidate at the end.
// Tiles whose list overflows LSC_CAP and pixels covered by no window take exact per-pixel paths.
// (no include guard needed: included once, inside namespace spx, by lsc.cu, which includes f32x2.cuh at file scope)

constexpr int LSC_NCOPY = 8;
constexpr int LSC_TILE_MIN_STEP = 10;   // measured at 4K (r2): step 10 tiled 3.2 ms vs per-pixel 4.6 ms per iterate(10), step 12: 2.2 vs 4.3; step 8: 15 vs 4.7 (lists overflow)
constexpr int LSC_SLOTW = 24;                      // words per slot: 10 feature sums + W as {lo, hi}; x | y << 16; count
constexpr int LSC_COPYW = LSC_CAP * LSC_SLOTW + 4; // the +4 words rotate the copies over the banks

struct __align__(16) LscRec {
    float4 c0, c1, c2;                             // -cen[0..3], -cen[4..7], (-cen[8], -cen[9], key as int bits, 0)
    float px[32], py[LSC_TH];
};
struct LscTileSmem {
    LscRec rec[LSC_CAP];                           // ascending seed index
    int4 win[LSC_CAP];                             // window, inclusive: x_lo, x_hi, y_lo, y_hi
    __align__(16) unsigned acc[LSC_NCOPY][LSC_COPYW];
    int k[LSC_CAP];
};

// a 64-bit value into a {lo, hi} pair of 32-bit shared-memory words WITHOUT waiting for a carry: lo takes the low 24 bits,
// hi the (signed) rest.  A word receives at most 64 adds per tile (a copy serves 4 lanes of 4 warps for 2 steps, two winners
// each), so lo stays below 2^30; |v| < 2^45 here (fixed point x 2^32 of features <= 51, four pixels), so hi fits easily.
// Both atomics are fire-and-forget (no returned value, no dependent instruction).  lsc_fold64 undoes the split.
__device__ __forceinline__ void lsc_sadd64(unsigned* p, unsigned long long v)
{
    atomicAdd(p, (unsigned)v & 0xffffffu);
    atomicAdd(p + 1, (unsigned)(int)((long long)v >> 24));
}
__device__ __forceinline__ unsigned long long lsc_fold64(unsigned lo, unsigned hi)
{
    return (unsigned long long)((long long)lo + ((long long)(int)hi << 24));
}
__device__ __forceinline__ unsigned long long lsc_q32(float Wt, float f)
{
    return (unsigned long long)__float2ll_rz(__fmul_rn(__fmul_rn(Wt, f), 4294967296.0f));
}
__device__ __forceinline__ unsigned long long lsc_q24(float Wt) { return (unsigned long long)__float2ll_rz(__fmul_rn(Wt, 16777216.0f)); }

// exact per-pixel form (bin walk, global atomics) for overflowing tiles; keep = true: the pixel keeps its label
__device__ __noinline__ void lsc_pixel_exact(const LscDev& L, int* __restrict__ labels, int x, int y, bool keep)
{
    constexpr int D = 10;
    float f[D];
    const float Wt = lsc_features<3>(L, x, y, f);
    const size_t li = (size_t)y * L.w + x;
    int bestk = -1;
    if (!keep) {
        float best = FLT_MAX;
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
    if (bestk >= 0) labels[li] = bestk;
    else bestk = labels[li];
    unsigned long long* dst = (unsigned long long*)(L.sums + (size_t)bestk * (D + 4));
#pragma unroll
    for (int i = 0; i < D; i++) atomicAdd(&dst[i], lsc_q32(Wt, f[i]));
    atomicAdd(&dst[D], lsc_q24(Wt));
    atomicAdd(&dst[D + 1], (unsigned long long)x); atomicAdd(&dst[D + 2], (unsigned long long)y); atomicAdd(&dst[D + 3], 1ull);
}

// the thread's 2x2 block into the per-candidate accumulators.  SPLIT = false: all four pixels belong to slot A.
template <bool SPLIT>
__device__ __forceinline__ void lsc_acc_block(unsigned* cp, const u64 (&P)[2][10], const float (&Wt)[2][2], unsigned A, unsigned B,
                                              bool e1, bool e2, bool e3, unsigned xo, unsigned yo)
{
    unsigned* const da = cp + LSC_SLOTW * A;
    unsigned* const db = cp + LSC_SLOTW * B;
    const bool hasB = SPLIT && A != B;
#pragma unroll
    for (int i = 0; i <= 10; i++) {
        unsigned long long q0, q1, q2, q3;
        if (i < 10) {
            float f00, f01, f10, f11;
            unpack2(P[0][i], f00, f01); unpack2(P[1][i], f10, f11);
            q0 = lsc_q32(Wt[0][0], f00); q1 = lsc_q32(Wt[0][1], f01); q2 = lsc_q32(Wt[1][0], f10); q3 = lsc_q32(Wt[1][1], f11);
        } else {
            q0 = lsc_q24(Wt[0][0]); q1 = lsc_q24(Wt[0][1]); q2 = lsc_q24(Wt[1][0]); q3 = lsc_q24(Wt[1][1]);
        }
        const unsigned long long T = (q0 + q1) + (q2 + q3);
        if (SPLIT) {
            const unsigned long long SA = q0 + (e1 ? q1 : 0ull) + (e2 ? q2 : 0ull) + (e3 ? q3 : 0ull);
            lsc_sadd64(da + 2 * i, SA);
            if (hasB) lsc_sadd64(db + 2 * i, T - SA);
        } else lsc_sadd64(da + 2 * i, T);
    }
    if (SPLIT) {
        const unsigned cntA = 1u + (e1 ? 1u : 0u) + (e2 ? 1u : 0u) + (e3 ? 1u : 0u);
        const unsigned dxA = (e1 ? 1u : 0u) + (e3 ? 1u : 0u), dyA = (e2 ? 1u : 0u) + (e3 ? 1u : 0u);
        atomicAdd(da + 22, (cntA * xo + dxA) | ((cntA * yo + dyA) << 16));
        atomicAdd(da + 23, cntA);
        if (hasB) {
            const unsigned cntB = 4u - cntA;
            atomicAdd(db + 22, (cntB * xo + (2u - dxA)) | ((cntB * yo + (2u - dyA)) << 16));
            atomicAdd(db + 23, cntB);
        }
    } else {
        atomicAdd(da + 22, (4u * xo + 2u) | ((4u * yo + 2u) << 16));
        atomicAdd(da + 23, 4u);
    }
}
// one pixel into slot s
__device__ __forceinline__ void lsc_acc_pixel(unsigned* cp, const u64 (&P)[2][10], const float (&Wt)[2][2], int r, int k, unsigned s,
                                              unsigned xo, unsigned yo)
{
    unsigned* const d = cp + LSC_SLOTW * s;
    const float W = Wt[r][k];
#pragma unroll
    for (int i = 0; i < 10; i++) {
        float f0, f1;
        unpack2(P[r][i], f0, f1);
        lsc_sadd64(d + 2 * i, lsc_q32(W, k ? f1 : f0));
    }
    lsc_sadd64(d + 20, lsc_q24(W));
    atomicAdd(d + 22, (xo + k) | ((yo + r) << 16));
    atomicAdd(d + 23, 1u);
}

__global__ void __launch_bounds__(128, 5) lsc_tile_kernel(const __grid_constant__ LscDev L, int* __restrict__ labels, int* __restrict__ d_overflow)
{
    constexpr int D = 10;
    __shared__ LscTileSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx0 = blockIdx.x * 32, ty0 = blockIdx.y * LSC_TH;
    // ---- candidates: the kernels that move the seeds appended every seed whose window meets this tile
    const int tile = blockIdx.y * L.tiles_x + blockIdx.x;
    const int cnt = __ldg(&L.tl_count[tile]);
    if (cnt > LSC_CAP) {
        if (tid == 0) atomicAdd(d_overflow, 1);
        for (int r = warp; r < LSC_TH; r += 4) {
            const int x = tx0 + lane, y = ty0 + r;
            if (x < L.w && y < L.h) lsc_pixel_exact(L, labels, x, y, false);
        }
        return;
    }
    int my_k = 0;
    if (tid < cnt) { my_k = L.tl_k[(size_t)tile * LSC_CAP + tid]; sm.k[tid] = my_k; }
    {   // zero the accumulators of the slots in use: copy tid >> 4, 16-byte stores
        unsigned* cpz = sm.acc[tid >> 4];
        for (int i = tid & 15; i < cnt * (LSC_SLOTW / 4); i += 16) *reinterpret_cast<uint4*>(cpz + 4 * i) = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    // ---- rank by seed index, stage the records; clear the accumulators
    if (tid < cnt) {
        const int k = my_k;
        int rank = 0;
        for (int j = 0; j < cnt; j++) rank += sm.k[j] < k ? 1 : 0;
        const float2* c = reinterpret_cast<const float2*>(L.cen + (size_t)k * D);
        const float2 a0 = c[0], a1 = c[1], a2 = c[2], a3 = c[3], a4 = c[4];
        LscRec& R = sm.rec[rank];
        R.c0 = make_float4(-a0.x, -a0.y, -a1.x, -a1.y);
        R.c1 = make_float4(-a2.x, -a2.y, -a3.x, -a3.y);
        R.c2 = make_float4(-a4.x, -a4.y, __int_as_float((k << 6) | rank), 0.f);
        const int2 s = make_int2(L.sx[k], L.sy[k]);
        sm.win[rank] = make_int4(s.x - L.stepx, s.x + L.stepx, s.y - L.stepy, s.y + L.stepy);
    }
    __syncthreads();
    {
        const float inf = __int_as_float(0x7f800000);
        for (int i = tid; i < cnt * (32 + LSC_TH); i += 128) {
            const int c = i / (32 + LSC_TH), e = i - c * (32 + LSC_TH);
            const bool isy = e >= 32;
            const int j = isy ? e - 32 : e;
            const int4 wn = sm.win[c];
            const int p = (isy ? ty0 : tx0) + j;
            const int lo = isy ? wn.z : wn.x, hi = isy ? wn.w : wn.y;
            (isy ? sm.rec[c].py : sm.rec[c].px)[j] = (p >= lo && p <= hi) ? 0.f : inf;
        }
    }
    __syncthreads();

    // ---- assignment: lanes 8 x 4, 2x2 pixels per thread, warps 2 x 2, two steps of 16 rows
    const int lx = lane & 7, ly = lane >> 3, wx = warp & 1, wy = warp >> 1;
    const int xo = 16 * wx + 2 * lx;
    const int x = tx0 + xo;
    const int xc0 = min(x, L.w - 1), xc1 = min(x + 1, L.w - 1);
    const float xcs[2] = {__ldg(&L.xcos[xc0]), __ldg(&L.xcos[xc1])}, xsn[2] = {__ldg(&L.xsin[xc0]), __ldg(&L.xsin[xc1])};
    const float wxv[2] = {__ldg(&L.wx[xc0]), __ldg(&L.wx[xc1])};
    const bool interior = tx0 + 32 <= L.w && ty0 + LSC_TH <= L.h;
    const bool pair_store = (L.w & 1) == 0;
    const int X0 = tx0 + 16 * wx, X1 = X0 + 15;
    unsigned* const cp = sm.acc[lane & (LSC_NCOPY - 1)];
    const unsigned char* rec0 = reinterpret_cast<const unsigned char*>(&sm.rec[0]);
    const unsigned char* thr_px = reinterpret_cast<const unsigned char*>(&sm.rec[0].px[xo]);
#pragma unroll 1
    for (int step = 0; step < LSC_TH / 16; step++) {
        const int yo = 16 * step + 8 * wy + 2 * ly;
        const int y = ty0 + yo;
        // features of the 2x2 block, packed by row: P[r][i] = (f_i(x, y + r), f_i(x + 1, y + r))
        u64 P[2][D];
        float Wt[2][2];
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const int yc = min(y + r, L.h - 1);
            const float ycs = __ldg(&L.ycos[yc]), ysn = __ldg(&L.ysin[yc]), wyv = __ldg(&L.wy[yc]);
            float f[2][D];
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const unsigned char* p = L.img + (size_t)yc * L.ipitch + (size_t)(k ? xc1 : xc0) * 3;
                const int v0 = __ldg(p), v1 = __ldg(p + 1), v2 = __ldg(p + 2);
                const float4 t0 = __ldg(&L.ctab[v0]), t1 = __ldg(&L.ctab[256 + v1]), t2 = __ldg(&L.ctab[512 + v2]);   // (cos, sin, w, -)
                const float wc = __fadd_rn(__fadd_rn(t0.z, t1.z), t2.z);
                const float W = __fadd_rn(wc, __fadd_rn(wxv[k], wyv));
                const float inv = __fdiv_rn(1.0f, W);
                Wt[r][k] = W;
                f[k][0] = __fmul_rn(t0.x, inv); f[k][1] = __fmul_rn(t0.y, inv);
                f[k][2] = __fmul_rn(t1.x, inv); f[k][3] = __fmul_rn(t1.y, inv);
                f[k][4] = __fmul_rn(t2.x, inv); f[k][5] = __fmul_rn(t2.y, inv);
                f[k][6] = __fmul_rn(xcs[k], inv); f[k][7] = __fmul_rn(xsn[k], inv);
                f[k][8] = __fmul_rn(ycs, inv); f[k][9] = __fmul_rn(ysn, inv);
            }
#pragma unroll
            for (int i = 0; i < D; i++) P[r][i] = pack2(f[0][i], f[1][i]);
        }
        // candidates whose window meets the warp's patch
        const int Y0 = ty0 + 16 * step + 8 * wy, Y1 = Y0 + 7;
        unsigned m;
        {
            const int4 wn = sm.win[lane];
            m = __ballot_sync(0xffffffffu, lane < cnt && wn.x <= X1 && wn.y >= X0 && wn.z <= Y1 && wn.w >= Y0);
        }
        float best[2][2] = {{FLT_MAX, FLT_MAX}, {FLT_MAX, FLT_MAX}};
        int key[2][2] = {{-1, -1}, {-1, -1}};
        const unsigned char* thr_py = reinterpret_cast<const unsigned char*>(&sm.rec[0].py[yo]);
        while (m) {
            const int c = __ffs(m) - 1;
            m &= m - 1;
            const int ro = c * (int)sizeof(LscRec);
            const float4 c0 = *reinterpret_cast<const float4*>(rec0 + ro);
            const float4 c1 = *reinterpret_cast<const float4*>(rec0 + ro + 16);
            const float4 c2 = *reinterpret_cast<const float4*>(rec0 + ro + 32);
            const int kw = __float_as_int(c2.z);
            const u64 pxv = *reinterpret_cast<const u64*>(thr_px + ro);
            const float2 pyv = *reinterpret_cast<const float2*>(thr_py + ro);
            const float cc[D] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w, c2.x, c2.y};
#pragma unroll
            for (int r = 0; r < 2; r++) {
                u64 t = add2(P[r][0], bcast2(cc[0]));
                u64 u = add2(P[r][1], bcast2(cc[1]));
                u64 d = add2_products(mul2(t, t), mul2(u, u));
#pragma unroll
                for (int i = 2; i < D; i++) {
                    t = add2(P[r][i], bcast2(cc[i]));
                    d = add2_products(d, mul2(t, t));
                }
                d = add2(d, add2(pxv, bcast2(r ? pyv.y : pyv.x)));
                float d0, d1;
                unpack2(d, d0, d1);
                if (d0 < best[r][0]) { best[r][0] = d0; key[r][0] = kw; }
                if (d1 < best[r][1]) { best[r][1] = d1; key[r][1] = kw; }
            }
        }
        // ---- labels + sums.  Fast path: interior tile and every pixel of the patch found a candidate.
        const int any_neg = key[0][0] | key[0][1] | key[1][0] | key[1][1];
        if (interior && !__any_sync(0xffffffffu, any_neg < 0)) {
            int* lrow = labels + (size_t)y * L.w + x;
            const int n00 = key[0][0] >> 6, n01 = key[0][1] >> 6, n10 = key[1][0] >> 6, n11 = key[1][1] >> 6;
            if (pair_store) {
                *reinterpret_cast<int2*>(lrow) = make_int2(n00, n01);
                *reinterpret_cast<int2*>(lrow + L.w) = make_int2(n10, n11);
            } else {
                lrow[0] = n00; lrow[1] = n01; lrow[L.w] = n10; lrow[L.w + 1] = n11;
            }
            const unsigned s00 = key[0][0] & 63, s01 = key[0][1] & 63, s10 = key[1][0] & 63, s11 = key[1][1] & 63;
            const unsigned A = s00;
            const bool e1 = s01 == A, e2 = s10 == A, e3 = s11 == A;
            if (__all_sync(0xffffffffu, e1 && e2 && e3)) {
                lsc_acc_block<false>(cp, P, Wt, A, A, true, true, true, (unsigned)xo, (unsigned)yo);
            } else {
                const unsigned B = !e1 ? s01 : (!e2 ? s10 : s11);
                const bool two = (e1 || s01 == B) && (e2 || s10 == B) && (e3 || s11 == B);
                if (two) lsc_acc_block<true>(cp, P, Wt, A, B, e1, e2, e3, (unsigned)xo, (unsigned)yo);
                else {
                    lsc_acc_pixel(cp, P, Wt, 0, 0, s00, (unsigned)xo, (unsigned)yo); lsc_acc_pixel(cp, P, Wt, 0, 1, s01, (unsigned)xo, (unsigned)yo);
                    lsc_acc_pixel(cp, P, Wt, 1, 0, s10, (unsigned)xo, (unsigned)yo); lsc_acc_pixel(cp, P, Wt, 1, 1, s11, (unsigned)xo, (unsigned)yo);
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) {                      // (unrolled: r, k must be compile-time indices of the register arrays)
                const int r = i >> 1, k = i & 1;
                const int xx = x + k, yy = y + r;
                if (xx >= L.w || yy >= L.h) continue;
                const int kk = key[r][k];
                if (kk >= 0) {
                    labels[(size_t)yy * L.w + xx] = kk >> 6;
                    lsc_acc_pixel(cp, P, Wt, r, k, (unsigned)(kk & 63), (unsigned)xo, (unsigned)yo);
                } else lsc_pixel_exact(L, labels, xx, yy, true);     // covered by no window: keeps its label, sums straight to global
            }
        }
    }

    __syncthreads();
    // ---- flush: fold the copies, one global atomic per field per candidate
    for (int i = tid; i < cnt * 13; i += 128) {
        const int c = i / 13, f = i - 13 * c;
        const int k = __float_as_int(sm.rec[c].c2.z) >> 6;
        unsigned long long* dst = (unsigned long long*)(L.sums + (size_t)k * (D + 4));
        if (f < 11) {
            unsigned long long v = 0;
#pragma unroll
            for (int p = 0; p < LSC_NCOPY; p++) {
                const unsigned* a = sm.acc[p] + LSC_SLOTW * c + 2 * f;
                v += lsc_fold64(a[0], a[1]);
            }
            if (v) atomicAdd(&dst[f], v);
        } else {
            unsigned xs = 0, ys = 0, n = 0;
#pragma unroll
            for (int p = 0; p < LSC_NCOPY; p++) {
                const unsigned* a = sm.acc[p] + LSC_SLOTW * c;
                xs += a[22] & 0xffffu; ys += a[22] >> 16; n += a[23];
            }
            if (!n) continue;
            if (f == 11) {
                atomicAdd(&dst[D + 1], (unsigned long long)xs + (unsigned long long)n * (unsigned long long)tx0);
                atomicAdd(&dst[D + 2], (unsigned long long)ys + (unsigned long long)n * (unsigned long long)ty0);
            } else atomicAdd(&dst[D + 3], (unsigned long long)n);
        }
    }
}
