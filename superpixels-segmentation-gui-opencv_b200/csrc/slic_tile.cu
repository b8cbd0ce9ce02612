// slic_tile.cu -- tiled fast path of the SLIC assignment + centre accumulation (3-channel, SLIC / SLICO / MSLIC).   [v7]
//
// One CTA (128 threads) owns a 32x32 pixel tile of the "Lab1" image (one 32-bit word per pixel: c0,c1,c2,1; rows and
// columns padded to the tile grid with zero words, so no load or store in the hot loop needs a bounds check).
//   setup   : the centre update appended to the tile's list every cluster whose window [int(k)-S, int(k)+S) meets the
//             tile.  The CTA reads the list (one 32-byte record per cluster), ranks it by cluster index (upstream visits
//             clusters in ascending n with a strict '<', so ascending order reproduces its tie-breaking), stages
//             (-k0,-k1,-k2,n) per candidate and tabulates the separable spatial terms ex2[c][x], ey2[c][y] with packed
//             f32x2 arithmetic (+inf outside the window: an out-of-window candidate can never win).
//   phase 1 : a warp handles a 16x8 patch per step, 2x2 pixels per thread, and visits only the candidates whose window
//             meets the patch (6.4 per patch at S=20; 4.0 per pixel is the minimum).  Upstream's float32 sequence
//                 d = ((t0*t0 + t1*t1) + t2*t2) + (ex2+ey2)/xywt
//             is evaluated with the sm_100 packed instructions FADD2/FMUL2 (same rounding per element, half the issue
//             slots; scalar operands are broadcast by the instruction), never contracted.  The IEEE division by the
//             constant xywt is a multiply when xywt is a power of two (folded into the tables), otherwise Markstein's
//             correctly rounded q = a*r; q += fma(-q, w, a)*r with r = RN(1/w).  The winner travels as one key
//             (n << 6) | slot; labels go out as 8-byte stores.
//   accumulate : right after the assignment a thread adds its 2x2 pixels (split by at most two winners with
//             byte-SIMD masks) to the CTA's per-candidate accumulators -- packed words (c0 | count, c1, c2, x | y) in 8
//             privatised copies; MSLIC adds the manifold-area limbs the same way, SLICO the maximum of the pixels'
//             colour distance to their last visitor.  Then one 64-bit global atomic per field per candidate that received
//             pixels.
// Algorithmic traffic: 3 B/px read + 4 B/px written per iteration (the padded word layout moves 4 + 4).
#include "slic.cuh"
#include "slic_generic.cuh"
#include "f32x2.cuh"
#include <cfloat>
#include <cstdlib>
#include <cmath>
#include <type_traits>

namespace spx {

constexpr int TW = TILE_W;   // 32: two warps side by side, 16 px each
constexpr int TH = TILE_H;   // 32: two warp rows x 8 rows x two steps
constexpr int CAP = TILE_CAP;
constexpr int NT = 128;      // threads per CTA
constexpr int NCOPY = 8;
constexpr int MIN_S = 7;     // measured at 4K (r2): S = 7 tiled 2.1 ms vs generic 5.3 ms per iterate(10) although many tiles overflow into the
                             // exact per-pixel path; S = 6: 12.1 vs 5.5 ms (every tile overflows CAP = 48); S = 8..13: 4-6x faster tiled

// ---- shared memory layout -----------------------------------------------------------------------
// one record per candidate, so that a single multiply addresses everything the inner loop reads
struct __align__(16) CandRec {
    float4 k;                    // (-k0, -k1, -k2, key = (n << 6) | slot as int bits)
    float4 m;                    // SLICO: (maxchans[n], RN(1/maxchans[n]), -maxchans[n], -); MSLIC: the same for adaptk[n]
    float ex2[TW];               // (x-kx)^2 [times 1/xywt when it is a power of two], +inf outside the window
    float ey2[TH];               // (y-ky)^2 likewise
};
struct SlicoSmem {               // SLICO only
    unsigned mx[NCOPY][CAP];     // per candidate: running maximum, over its pixels, of the colour distance to the pixel's LAST
};                               // visiting cluster (upstream's distchans plane; float bits, values are >= 0)
struct MslicSmem {               // MSLIC only: manifold area (fixed point x256) per candidate, as 16-bit limbs
    unsigned ar[NCOPY][CAP][2];  // [0]: sum of (v & 0xffff), [1]: sum of (v >> 16); a tile has 1024 pixels, v < 2^32
};
struct NoSmem {};
// rec[].k.w carries key = (n << 6) | slot: the inner loop selects one register and both the label and the slot follow
__device__ __forceinline__ int key_n(int key) { return key >> 6; }
__device__ __forceinline__ int key_slot(int key) { return key & 63; }
static_assert(CAP <= 64, "slot must fit the 6 low bits of the key");
template <int ALG>
struct __align__(16) TileSmemT {
    CandRec rec[CAP];            // ascending n
    int4 win[CAP];               // window: x_lo, x_hi, y_lo, y_hi (half open)
    // accumulators: NCOPY privatised copies (a lane uses copy lane&7, so lanes that sit in the same cluster rarely hit
    // the same address); per candidate 4 packed words: c0 | count<<18, c1, c2, x | y<<16 (tile-relative; a tile has
    // 1024 pixels, so c <= 261120 < 2^18, count <= 2^10, x,y sums <= 31744 < 2^16)
    __align__(16) unsigned acc[NCOPY][CAP * 4 + 4];
    float kxy[CAP][2];           // centre position per slot
    int list[CAP];
    typename std::conditional<ALG == SPX_SLICO, SlicoSmem, NoSmem>::type o;
    typename std::conditional<ALG == SPX_MSLIC, MslicSmem, NoSmem>::type ms;
    int bad;                     // SLICO / MSLIC: a divisor the reciprocal sequence cannot handle -> exact generic path for the tile
};

struct TileArgs {
    const int* tl_count; const TileRec* tl_rec;
    unsigned long long* acc;
    const unsigned* img4;        // Lab1 words, pitch p4 (pixels)
    int w, h, S, Kcap, tiles_x, lp, p4;
    int ty_off;                  // first tile row of the launch (strip mode)
    float xywt, rcp_xywt;
    const float* maxc; unsigned* maxc_acc; float* dcplane;      // SLICO (dcplane has the label pitch)
    const float* area_px; unsigned long long* area;              // MSLIC: per-pixel manifold area (label pitch), per-cluster sums
};

__device__ __forceinline__ void acc_global_pixel(const TileArgs& t, int n, unsigned px, int x, int y)
{
    atomicAdd(&t.acc[n], (u64)(px & 0xffu)); atomicAdd(&t.acc[(size_t)t.Kcap + n], (u64)((px >> 8) & 0xffu));
    atomicAdd(&t.acc[2 * (size_t)t.Kcap + n], (u64)((px >> 16) & 0xffu)); atomicAdd(&t.acc[3 * (size_t)t.Kcap + n], (u64)x);
    atomicAdd(&t.acc[4 * (size_t)t.Kcap + n], (u64)y); atomicAdd(&t.acc[5 * (size_t)t.Kcap + n], 1ull);
}
// MSLIC: fixed-point area of a pixel quad, as mslic_adapt reads it (x256, truncated)
__device__ __forceinline__ unsigned long long area_fx(float a) { return (unsigned long long)(long long)__fmul_rn(a, 256.0f); }

// exact per-pixel path for a tile whose candidate list overflowed (kept out of line: it is cold)
template <int ALG>
__device__ __noinline__ void tile_fallback(const SlicGeom& g, const SlicArrays& a, const unsigned char* img, int* labels, float* dcplane,
                                           const float* area_px, unsigned long long* area, int parity, int tx0, int ty0, int warp, int lane)
{
    for (int r = warp; r < TH; r += NT / 32) {
        const int x = tx0 + lane, y = ty0 + r;
        slic_pixel_generic<ALG>(g, a, img, labels, dcplane, parity, x, y, x < g.w && y < g.h);
        if constexpr (ALG == SPX_MSLIC) {                        // the tile's share of the per-cluster manifold areas
            if (x < g.w - 1 && y < g.h - 1) {
                const size_t li = (size_t)y * g.lp + x;
                atomicAdd(&area[labels[li]], area_fx(area_px[li]));
            }
        }
    }
}

// The thread's 2x2 pixels go into the CTA's per-candidate accumulators right after the assignment (no slot map, no second
// pass).  The block belongs to at most two candidates A (the first pixel's) and B in all but ~1 % of the cases:
// byte-SIMD sums (even bytes c0,c2 / odd bytes c1,1 as 16-bit lanes) split by masks, 4 packed atomics per candidate.
template <int ALG>
__device__ __forceinline__ void acc_block(TileSmemT<ALG>& sm, int lane, unsigned s00, unsigned s01, unsigned s10, unsigned s11,
                                          unsigned p00, unsigned p01, unsigned p10, unsigned p11, unsigned xo, unsigned yo,
                                          float l00 = 0.f, float l01 = 0.f, float l10 = 0.f, float l11 = 0.f)
{
    const unsigned A = s00;
    const bool e1 = s01 == A, e2 = s10 == A, e3 = s11 == A;
    const unsigned B = !e1 ? s01 : (!e2 ? s10 : s11);
    const bool two = (e1 || s01 == B) && (e2 || s10 == B) && (e3 || s11 == B);
    unsigned* const cp = sm.acc[lane & (NCOPY - 1)];
    if (two) {
        const unsigned E0 = p00 & 0x00ff00ffu, E1 = p01 & 0x00ff00ffu, E2 = p10 & 0x00ff00ffu, E3 = p11 & 0x00ff00ffu;
        const unsigned O0 = (p00 >> 8) & 0x00ff00ffu, O1 = (p01 >> 8) & 0x00ff00ffu, O2 = (p10 >> 8) & 0x00ff00ffu, O3 = (p11 >> 8) & 0x00ff00ffu;
        const unsigned m1 = e1 ? 0xffffffffu : 0u, m2 = e2 ? 0xffffffffu : 0u, m3 = e3 ? 0xffffffffu : 0u;
        const unsigned AE = E0 + (E1 & m1) + (E2 & m2) + (E3 & m3);
        const unsigned AO = O0 + (O1 & m1) + (O2 & m2) + (O3 & m3);
        const unsigned cntA = AO >> 16;
        const unsigned dxA = (m1 & 1u) + (m3 & 1u), dyA = (m2 & 1u) + (m3 & 1u);       // offsets of A's pixels inside the block
        unsigned* da = cp + 4 * A;
        atomicAdd(da + 0, (AE & 0xffffu) | (cntA << 18)); atomicAdd(da + 1, AO & 0xffffu); atomicAdd(da + 2, AE >> 16);
        atomicAdd(da + 3, (cntA * xo + dxA) | ((cntA * yo + dyA) << 16));
        if constexpr (ALG == SPX_SLICO) {
            const unsigned d0 = __float_as_uint(l00), d1 = __float_as_uint(l01), d2 = __float_as_uint(l10), d3 = __float_as_uint(l11);
            atomicMax(&sm.o.mx[lane & (NCOPY - 1)][A], max(max(d0, d1 & m1), max(d2 & m2, d3 & m3)));
            if (A != B) atomicMax(&sm.o.mx[lane & (NCOPY - 1)][B], max(max(d1 & ~m1, d2 & ~m2), d3 & ~m3));
        }
        if (A != B) {
            const unsigned BE = (E0 + E1 + E2 + E3) - AE, BO = (O0 + O1 + O2 + O3) - AO;
            const unsigned cntB = 4u - cntA;
            unsigned* db = cp + 4 * B;
            atomicAdd(db + 0, (BE & 0xffffu) | (cntB << 18)); atomicAdd(db + 1, BO & 0xffffu); atomicAdd(db + 2, BE >> 16);
            atomicAdd(db + 3, (cntB * xo + (2u - dxA)) | ((cntB * yo + (2u - dyA)) << 16));
        }
    } else {
        const unsigned ss[4] = {s00, s01, s10, s11}, pw[4] = {p00, p01, p10, p11};
        const float ll[4] = {l00, l01, l10, l11};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            unsigned* d = cp + 4 * ss[j];
            atomicAdd(d + 0, (pw[j] & 0xffu) | (1u << 18)); atomicAdd(d + 1, (pw[j] >> 8) & 0xffu);
            atomicAdd(d + 2, (pw[j] >> 16) & 0xffu); atomicAdd(d + 3, (xo + (j & 1)) | ((yo + (j >> 1)) << 16));
            if constexpr (ALG == SPX_SLICO) atomicMax(&sm.o.mx[lane & (NCOPY - 1)][ss[j]], __float_as_uint(ll[j]));
        }
    }
}
// MSLIC: manifold area (fixed point x256) of one pixel quad into the winner's limbs; beyond 32 bits straight to the global sum
template <int ALG>
__device__ __forceinline__ void acc_area(const TileArgs& t, TileSmemT<ALG>& sm, int lane, unsigned s, float a)
{
    if constexpr (ALG == SPX_MSLIC) {
        const unsigned long long v = area_fx(a);
        if (!v) return;
        if (v >> 32) atomicAdd(&t.area[key_n(__float_as_int(sm.rec[s].k.w))], v);
        else {
            atomicAdd(&sm.ms.ar[lane & (NCOPY - 1)][s][0], (unsigned)v & 0xffffu);
            atomicAdd(&sm.ms.ar[lane & (NCOPY - 1)][s][1], (unsigned)v >> 16);
        }
    }
}

// phase-1 cold path: some pixel of the warp's patch is covered by no window (or lies outside the image).
// Everything is passed by value so that the hot loop's state never has to live in local memory.
template <int ALG>
__device__ __noinline__ void step_slow(const TileArgs& t, TileSmemT<ALG>& sm, int* labels, int k00, int k01, int k10, int k11,
                                       unsigned p00, unsigned p01, unsigned p10, unsigned p11, float l00, float l01, float l10, float l11,
                                       int x, int y, int xo, int yo)
{
    const unsigned px[4] = {p00, p01, p10, p11};
    const float ldc[4] = {l00, l01, l10, l11};
    const int keys[4] = {k00, k01, k10, k11};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int r = i >> 1, k = i & 1;
        const int xx = x + k, yy = y + r;
        const int key = keys[i];                                // < 0: no candidate
        if (xx < t.w && yy < t.h) {
            const size_t li = (size_t)yy * t.lp + xx;
            if (key >= 0) {
                labels[li] = key_n(key);
                if constexpr (ALG == SPX_SLICO) {
                    t.dcplane[li] = ldc[i];
                    atomicMax(&sm.o.mx[threadIdx.x & (NCOPY - 1)][key_slot(key)], __float_as_uint(ldc[i]));
                }
                {
                    unsigned* d = sm.acc[threadIdx.x & (NCOPY - 1)] + 4 * key_slot(key);
                    atomicAdd(d + 0, (px[i] & 0xffu) | (1u << 18)); atomicAdd(d + 1, (px[i] >> 8) & 0xffu);
                    atomicAdd(d + 2, (px[i] >> 16) & 0xffu); atomicAdd(d + 3, (unsigned)(xo + k) | ((unsigned)(yo + r) << 16));
                    if constexpr (ALG == SPX_MSLIC) { if (xx < t.w - 1 && yy < t.h - 1) acc_area<ALG>(t, sm, threadIdx.x & 31, (unsigned)key_slot(key), t.area_px[li]); }
                }
            } else {
                const int n = labels[li];
                acc_global_pixel(t, n, px[i], xx, yy);          // keeps its label; sums go straight to global
                if constexpr (ALG == SPX_MSLIC) { if (xx < t.w - 1 && yy < t.h - 1) atomicAdd(&t.area[n], area_fx(t.area_px[li])); }
                if constexpr (ALG == SPX_SLICO) atomicMax(&t.maxc_acc[n], __float_as_uint(t.dcplane[li]));   // plane keeps its value
            }
        }
    }
}

template <bool POW2, int ALG>
__global__ void __launch_bounds__(NT, ALG == SPX_SLIC ? 9 : 8) slic_tile_kernel(const __grid_constant__ TileArgs t, const __grid_constant__ SlicGeom g,
                                                           const __grid_constant__ SlicArrays a, const unsigned char* __restrict__ img,
                                                           int* __restrict__ labels, int parity, int* __restrict__ d_overflow)
{
    __shared__ TileSmemT<ALG> sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = t.S;
    const int tyi = blockIdx.y + t.ty_off;
    const int tile = tyi * t.tiles_x + blockIdx.x;
    const int tx0 = blockIdx.x * TW, ty0 = tyi * TH;

    // ---- first pixels in flight before anything else (lanes 8 x 4, 2x2 pixels per thread, warps 2 x 2)
    const int lx = lane & 7, ly = lane >> 3;
    const int wx = warp & 1, wy = warp >> 1;
    const int xo = 16 * wx + 2 * lx;           // tile-relative x of the thread's left pixels
    const int x = tx0 + xo;
    const unsigned* pimg = t.img4 + (size_t)(ty0 + 8 * wy + 2 * ly) * t.p4 + x;    // this thread's first pixel pair
    uint2 nxt0 = __ldg(reinterpret_cast<const uint2*>(pimg));
    uint2 nxt1 = __ldg(reinterpret_cast<const uint2*>(pimg + t.p4));
    const int p4_16 = 16 * t.p4, p4_17 = 17 * t.p4;

    // ---- candidates: the centre-update kernel appended every cluster whose window touches this tile
    // (the record loads do not wait for the count: slots beyond it hold stale records that are never used)
    pdl_wait();                                                   // the centre update that wrote the lists (the pixel loads above are already in flight)
    float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0;
    if (tid < CAP) {
        const float4* src = reinterpret_cast<const float4*>(&t.tl_rec[(size_t)tile * CAP + tid]);
        r0 = __ldg(src); r1 = __ldg(src + 1);
    }
    const int cnt = __ldg(&t.tl_count[tile]);
    if (cnt > CAP) {
        if (tid == 0) atomicAdd(d_overflow, 1);
        tile_fallback<ALG>(g, a, img, labels, t.dcplane, t.area_px, t.area, parity, tx0, ty0, warp, lane);
        return;
    }
    float mxv = 1.0f;
    int woff = S;                                                  // window half-size (MSLIC: int(S * adaptk[n]))
    if (tid < cnt) {
        sm.list[tid] = __float_as_int(r1.y);
        if constexpr (ALG == SPX_SLICO) mxv = __ldg(&t.maxc[__float_as_int(r1.y)]);
        if constexpr (ALG == SPX_MSLIC) { mxv = a.adaptk[__float_as_int(r1.y)]; woff = a.ioff[__float_as_int(r1.y)]; }
    }
    if (ALG != SPX_SLIC && tid == 0) sm.bad = 0;
    __syncthreads();
    // ---- rank by cluster index (upstream's visiting order), stage centres
    if (tid < cnt) {
        const int n = __float_as_int(r1.y);
        int rank = 0;
        for (int j = 0; j < cnt; j++) rank += sm.list[j] < n ? 1 : 0;
        sm.rec[rank].k = make_float4(-r0.z, -r0.w, -r1.x, __int_as_float((n << 6) | rank));
        const int ix = __float_as_int(r1.z), iy = __float_as_int(r1.w);
        sm.win[rank] = make_int4(ix - woff, ix + woff, iy - woff, iy + woff);
        sm.kxy[rank][0] = r0.x; sm.kxy[rank][1] = r0.y;
        if constexpr (ALG != SPX_SLIC) {
            // v / m (m = maxchans[n] or adaptk[n]) is evaluated as q = v*r; q += fma(-q, m, v)*r with r = RN(1/m): correctly
            // rounded unless the significand of m is all ones
            sm.rec[rank].m = make_float4(mxv, __fdiv_rn(1.0f, mxv), -mxv, 0.f);
            if ((__float_as_uint(mxv) & 0x7fffffu) == 0x7fffffu) sm.bad = 1;
        }
    }
    if constexpr (ALG == SPX_SLICO) {
        for (int i = tid; i < NCOPY * CAP; i += NT) (&sm.o.mx[0][0])[i] = 0u;
    }
    if constexpr (ALG == SPX_MSLIC) {
        for (int i = tid; i < NCOPY * CAP * 2; i += NT) (&sm.ms.ar[0][0][0])[i] = 0u;
    }
    {   // zero the accumulators of the candidates in use: copy tid>>4, candidates (tid&15), +16, ... as 16-byte stores
        unsigned* cp = sm.acc[tid >> 4];
        for (int c = tid & 15; c < cnt; c += 16) *reinterpret_cast<uint4*>(cp + 4 * c) = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    if constexpr (ALG != SPX_SLIC) {
        if (sm.bad) {
            tile_fallback<ALG>(g, a, img, labels, t.dcplane, t.area_px, t.area, parity, tx0, ty0, warp, lane);
            return;
        }
    }
    // ---- separable spatial tables, two entries per job: job (c, r): r < 16 -> ex2[c][2r..2r+1], else ey2[c][2(r-16)..]
    {
        const float inf = __int_as_float(0x7f800000);
        for (int i = tid; i < cnt * 32; i += NT) {
            const int c = i >> 5, r = i & 31;
            const bool isx = r < 16;
            const int j = 2 * (r & 15);
            const int4 wn = sm.win[c];
            const int p = (isx ? tx0 : ty0) + j;
            const int lo = isx ? wn.x : wn.z, hi = isx ? wn.y : wn.w;
            const float k = sm.kxy[c][isx ? 0 : 1];
            // int -> float without the conversion unit: (2^23 + p) - 2^23, exact for p < 2^23
            const float f0 = __fsub_rn(__int_as_float(0x4B000000 | p), 8388608.0f);
            const u64 e = add2(pack2(f0, __fadd_rn(f0, 1.0f)), bcast2(-k));
            u64 v = mul2(e, e);
            if (POW2) v = mul2(v, bcast2(t.rcp_xywt));
            float v0, v1;
            unpack2(v, v0, v1);
            v0 = (p >= lo && p < hi) ? v0 : inf;
            v1 = (p + 1 >= lo && p + 1 < hi) ? v1 : inf;
            float* dst = isx ? &sm.rec[c].ex2[j] : &sm.rec[c].ey2[j];
            *reinterpret_cast<float2*>(dst) = make_float2(v0, v1);
        }
    }
    __syncthreads();

    // ---- phase 1: a warp owns a 16x8 patch per step; the CTA's 4 warps cover 32x16, two steps cover the tile
    const int X0 = tx0 + 16 * wx, X1 = X0 + 16;
    int* lrow = labels + (size_t)(ty0 + 8 * wy + 2 * ly) * t.lp + x;
    const bool interior = tx0 + TW <= t.w && ty0 + TH <= t.h;
#pragma unroll 1
    for (int step = 0; step < TH / 16; step++) {
        const int yo = 16 * step + 8 * wy + 2 * ly;
        const int y = ty0 + yo;
        unsigned px[2][2] = {{nxt0.x, nxt0.y}, {nxt1.x, nxt1.y}};
        if (step + 1 < TH / 16) {                                 // prefetch the next step's pixels
            nxt0 = __ldg(reinterpret_cast<const uint2*>(pimg + p4_16));
            nxt1 = __ldg(reinterpret_cast<const uint2*>(pimg + p4_17));
        }
        // unpack without the conversion unit: byte -> float as (0x4B0000bb as float) - 2^23, built with PRMT
        u64 P[2][3];
        {
            const unsigned MAGIC = 0x4B000000u;
            const u64 M2 = bcast2(-8388608.0f);
#pragma unroll
            for (int r = 0; r < 2; r++) {
                const float a0 = __uint_as_float(__byte_perm(px[r][0], MAGIC, 0x7440)), b0 = __uint_as_float(__byte_perm(px[r][1], MAGIC, 0x7440));
                const float a1 = __uint_as_float(__byte_perm(px[r][0], MAGIC, 0x7441)), b1 = __uint_as_float(__byte_perm(px[r][1], MAGIC, 0x7441));
                const float a2 = __uint_as_float(__byte_perm(px[r][0], MAGIC, 0x7442)), b2 = __uint_as_float(__byte_perm(px[r][1], MAGIC, 0x7442));
                P[r][0] = add2(pack2(a0, b0), M2); P[r][1] = add2(pack2(a1, b1), M2); P[r][2] = add2(pack2(a2, b2), M2);
            }
        }
        // candidates whose window meets this warp's patch (CAP <= 64: two ballots)
        const int Y0 = ty0 + 16 * step + 8 * wy, Y1 = Y0 + 8;
        unsigned m0, m1 = 0;
        {
            int4 wn = sm.win[lane];
            m0 = __ballot_sync(0xffffffffu, lane < cnt && wn.x < X1 && wn.y > X0 && wn.z < Y1 && wn.w > Y0);
            if (cnt > 32) {
                wn = sm.win[min(lane + 32, CAP - 1)];
                m1 = __ballot_sync(0xffffffffu, lane + 32 < cnt && wn.x < X1 && wn.y > X0 && wn.z < Y1 && wn.w > Y0);
            }
        }
        float best[2][2] = {{FLT_MAX, FLT_MAX}, {FLT_MAX, FLT_MAX}};
        int key[2][2] = {{-1, -1}, {-1, -1}};           // (n << 6) | slot of the best candidate so far
        float ldc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};     // SLICO: colour distance to the last cluster whose window holds the pixel
        const unsigned char* rec0 = reinterpret_cast<const unsigned char*>(&sm.rec[0]);
        const unsigned char* thr_ex = reinterpret_cast<const unsigned char*>(&sm.rec[0].ex2[xo]);
        const unsigned char* thr_ey = reinterpret_cast<const unsigned char*>(&sm.rec[0].ey2[yo]);
#pragma unroll 1
        for (int half = 0; half < 2; half++) {
            unsigned m = half ? m1 : m0;
            while (m) {
                const int c = (__ffs(m) - 1) + 32 * half;
                m &= m - 1;
                const int ro = c * (int)sizeof(CandRec);
                const float4 kc = *reinterpret_cast<const float4*>(rec0 + ro);
                const int kw = __float_as_int(kc.w);
                float4 km = make_float4(0.f, 0.f, 0.f, 0.f);
                if constexpr (ALG != SPX_SLIC) km = *reinterpret_cast<const float4*>(rec0 + ro + 16);
                const u64 exv = *reinterpret_cast<const u64*>(thr_ex + ro);
                const float2 eyv = *reinterpret_cast<const float2*>(thr_ey + ro);
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    const u64 t0 = add2(P[r][0], bcast2(kc.x)), t1 = add2(P[r][1], bcast2(kc.y)), t2 = add2(P[r][2], bcast2(kc.z));
                    u64 d = add2_products(mul2(t0, t0), mul2(t1, t1));
                    d = add2_products(d, mul2(t2, t2));
                    u64 sxy = add2(exv, bcast2(r ? eyv.y : eyv.x));
                    if constexpr (ALG == SPX_SLICO) {
                        float s0, s1, c0, c1;
                        unpack2(sxy, s0, s1);
                        unpack2(d, c0, c1);
                        const float inf = __int_as_float(0x7f800000);
                        ldc[r][0] = s0 < inf ? c0 : ldc[r][0];            // candidates come in ascending n: the last one stays
                        ldc[r][1] = s1 < inf ? c1 : ldc[r][1];
                        const u64 q0 = mul2(d, bcast2(km.y));
                        const u64 e = fma2(q0, bcast2(km.z), d);
                        d = fma2(e, bcast2(km.y), q0);
                    }
                    if (!POW2) {
                        const u64 q0 = mul2(sxy, bcast2(t.rcp_xywt));
                        const u64 e = fma2(q0, bcast2(-t.xywt), sxy);
                        sxy = fma2(e, bcast2(t.rcp_xywt), q0);
                    }
                    d = add2(d, sxy);
                    if constexpr (ALG == SPX_MSLIC) {                      // the whole distance is divided by adaptk[n]
                        const u64 q0 = mul2(d, bcast2(km.y));
                        const u64 e = fma2(q0, bcast2(km.z), d);
                        d = fma2(e, bcast2(km.y), q0);
                    }
                    float d0, d1;
                    unpack2(d, d0, d1);
                    if (d0 < best[r][0]) { best[r][0] = d0; key[r][0] = kw; }
                    if (d1 < best[r][1]) { best[r][1] = d1; key[r][1] = kw; }
                }
            }
        }
        // ---- labels + accumulation.  Fast path: every pixel of the patch found a candidate and the tile is interior.
        const int any_neg = key[0][0] | key[0][1] | key[1][0] | key[1][1];
        if (interior && !__any_sync(0xffffffffu, any_neg < 0)) {
            const int n00 = key_n(key[0][0]), n01 = key_n(key[0][1]), n10 = key_n(key[1][0]), n11 = key_n(key[1][1]);
            *reinterpret_cast<int2*>(lrow) = make_int2(n00, n01);
            *reinterpret_cast<int2*>(lrow + t.lp) = make_int2(n10, n11);
            acc_block<ALG>(sm, lane, key_slot(key[0][0]), key_slot(key[0][1]), key_slot(key[1][0]), key_slot(key[1][1]), px[0][0], px[0][1],
                           px[1][0], px[1][1], (unsigned)xo, (unsigned)yo, ldc[0][0], ldc[0][1], ldc[1][0], ldc[1][1]);
            if constexpr (ALG == SPX_MSLIC) {                      // (zero in the image's last row / column: the plane holds 0 there)
                const float2 a0 = __ldg(reinterpret_cast<const float2*>(t.area_px + (lrow - labels)));
                const float2 a1 = __ldg(reinterpret_cast<const float2*>(t.area_px + (lrow - labels) + t.lp));
                const unsigned s0 = key_slot(key[0][0]), s1 = key_slot(key[0][1]), s2 = key_slot(key[1][0]), s3 = key_slot(key[1][1]);
                const unsigned long long v0 = area_fx(a0.x), v1 = area_fx(a0.y), v2 = area_fx(a1.x), v3 = area_fx(a1.y);
                if (s0 == s1 && s0 == s2 && s0 == s3 && !((v0 | v1 | v2 | v3) >> 32)) {      // one candidate, ordinary values: one pair of atomics
                    const unsigned lo = ((unsigned)v0 & 0xffffu) + ((unsigned)v1 & 0xffffu) + ((unsigned)v2 & 0xffffu) + ((unsigned)v3 & 0xffffu);
                    const unsigned hi = ((unsigned)v0 >> 16) + ((unsigned)v1 >> 16) + ((unsigned)v2 >> 16) + ((unsigned)v3 >> 16);
                    if (lo | hi) {
                        atomicAdd(&sm.ms.ar[lane & (NCOPY - 1)][s0][0], lo);
                        atomicAdd(&sm.ms.ar[lane & (NCOPY - 1)][s0][1], hi);
                    }
                } else {
                    acc_area<ALG>(t, sm, lane, s0, a0.x); acc_area<ALG>(t, sm, lane, s1, a0.y);
                    acc_area<ALG>(t, sm, lane, s2, a1.x); acc_area<ALG>(t, sm, lane, s3, a1.y);
                }
            }
            if constexpr (ALG == SPX_SLICO) {
                float* prow = t.dcplane + (lrow - labels);
                *reinterpret_cast<float2*>(prow) = make_float2(ldc[0][0], ldc[0][1]);
                *reinterpret_cast<float2*>(prow + t.lp) = make_float2(ldc[1][0], ldc[1][1]);
            }
        } else {
            step_slow<ALG>(t, sm, labels, key[0][0], key[0][1], key[1][0], key[1][1], px[0][0], px[0][1], px[1][0], px[1][1],
                                  ldc[0][0], ldc[0][1], ldc[1][0], ldc[1][1], x, y, xo, yo);
        }
        pimg += p4_16;
        lrow += 16 * t.lp;
    }

    pdl_trigger();                                                // the centre update may become resident while the last tiles flush
    __syncthreads();
    // ---- flush: fold the copies, then one global atomic per field per candidate that received pixels
    for (int i = tid; i < cnt * 4; i += NT) {
        const int c = i >> 2, f = i & 3;
        unsigned v = 0, v0 = 0;
#pragma unroll
        for (int p = 0; p < NCOPY; p++) { v += sm.acc[p][i]; v0 += sm.acc[p][4 * c]; }
        const unsigned count = v0 >> 18;
        if (!count) continue;
        const int n = key_n(__float_as_int(sm.rec[c].k.w));
        if (f == 0) {
            atomicAdd(&t.acc[n], (u64)(v & 0x3ffffu));
            atomicAdd(&t.acc[(size_t)5 * t.Kcap + n], (u64)count);
            if constexpr (ALG == SPX_MSLIC) {
                u64 lo = 0, hi = 0;
#pragma unroll
                for (int p = 0; p < NCOPY; p++) { lo += sm.ms.ar[p][c][0]; hi += sm.ms.ar[p][c][1]; }
                if (lo | hi) atomicAdd(&t.area[n], lo + (hi << 16));
            }
            if constexpr (ALG == SPX_SLICO) {
                unsigned mxb = 0;
#pragma unroll
                for (int p = 0; p < NCOPY; p++) mxb = max(mxb, sm.o.mx[p][c]);
                atomicMax(&t.maxc_acc[n], mxb);
            }
        } else if (f < 3) {
            if (v) atomicAdd(&t.acc[(size_t)f * t.Kcap + n], (u64)v);
        } else {
            atomicAdd(&t.acc[(size_t)3 * t.Kcap + n], (u64)(v & 0xffffu) + (u64)count * (u64)tx0);
            atomicAdd(&t.acc[(size_t)4 * t.Kcap + n], (u64)(v >> 16) + (u64)count * (u64)ty0);
        }
    }
}

// xywt = 2^k ?
static bool is_pow2(float v)
{
    int e;
    return v > 0.f && std::frexp(v, &e) == 0.5f;
}
// Markstein's correctly rounded division needs RN(1/w) to be a good reciprocal: true for every binary32 w whose
// significand is not all ones (and away from over/underflow, which (S/ruler)^2 never approaches).
static bool markstein_ok(float w)
{
    unsigned bits;
    memcpy(&bits, &w, 4);
    return w > 1e-18f && w < 1e18f && (bits & 0x7fffffu) != 0x7fffffu;
}

bool slic_tile_supported(const SlicGeom& g)
{
    static const int min_s = getenv("SPX_TILE_MIN_S") ? atoi(getenv("SPX_TILE_MIN_S")) : MIN_S;    // (experiments)
    return (g.alg == SPX_SLIC || g.alg == SPX_SLICO || g.alg == SPX_MSLIC) && g.C == 3 && g.S >= min_s && (is_pow2(g.xywt) || markstein_ok(g.xywt));
}

// 8UC3 rows -> Lab1 words (c0 | c1<<8 | c2<<16 | 1<<24); the padding of the tile grid stays zero
__global__ void __launch_bounds__(256) slic_expand_lab1_kernel(const unsigned char* __restrict__ img, size_t ipitch,
                                                                unsigned* __restrict__ img4, int p4, int w, int h)
{
    const int gx = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    const int x = gx * 4;
    if (x >= w || y >= h) return;
    const unsigned* src = reinterpret_cast<const unsigned*>(img + (size_t)y * ipitch + (size_t)x * 3);   // 12-byte groups, 4-aligned
    const unsigned a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2);    // the image buffer is padded by 16 bytes per row
    uint4 o;
    o.x = (a & 0x00ffffffu) | 0x01000000u;
    o.y = (__byte_perm(a, b, 0x0543) & 0x00ffffffu) | 0x01000000u;
    o.z = (__byte_perm(b, c, 0x0432) & 0x00ffffffu) | 0x01000000u;
    o.w = (c >> 8) | 0x01000000u;
    if (x + 1 >= w) o.y = 0;
    if (x + 2 >= w) o.z = 0;
    if (x + 3 >= w) o.w = 0;
    *reinterpret_cast<uint4*>(img4 + (size_t)y * p4 + x) = o;
}

int slic_tile_prepare(const SlicGeom& g, const unsigned char* d_img, unsigned* d_img4, cudaStream_t st)
{
    dim3 grid(cdiv(cdiv(g.w, 4), 256), g.h);
    slic_expand_lab1_kernel<<<grid, 256, 0, st>>>(d_img, g.ipitch, d_img4, g.p4, g.w, g.h);
    SPX_CUDA(cudaGetLastError());
    return SPX_OK;
}

int slic_tile_assign(const SlicGeom& g, const SlicArrays& a, const unsigned char* d_img, const unsigned* d_img4, int* d_labels,
                     float* d_dcplane, const float* d_area_px, const int* d_K, int parity, int* d_overflow, cudaStream_t st)
{
    (void)d_K;
    dim3 grid(g.tiles_x, g.ty_hi - g.ty_lo);
    TileArgs t;
    t.ty_off = g.ty_lo;
    t.tl_count = a.tl_count[parity]; t.tl_rec = a.tl_rec[parity]; t.acc = a.acc;
    t.img4 = d_img4;
    t.w = g.w; t.h = g.h; t.S = g.S; t.Kcap = g.Kcap; t.tiles_x = g.tiles_x; t.lp = g.lp; t.p4 = g.p4;
    t.xywt = g.xywt; t.rcp_xywt = 1.0f / g.xywt;
    t.maxc = a.maxc; t.maxc_acc = a.maxc_acc; t.dcplane = d_dcplane;
    t.area_px = d_area_px; t.area = reinterpret_cast<unsigned long long*>(a.area);
    const bool p2 = is_pow2(g.xywt);
    if (g.alg == SPX_SLICO) {
        if (p2) SPX_CUDA(launch_pdl(slic_tile_kernel<true, SPX_SLICO>, grid, dim3(NT), st, t, g, a, d_img, d_labels, parity, d_overflow));
        else SPX_CUDA(launch_pdl(slic_tile_kernel<false, SPX_SLICO>, grid, dim3(NT), st, t, g, a, d_img, d_labels, parity, d_overflow));
    } else if (g.alg == SPX_MSLIC) {
        if (p2) SPX_CUDA(launch_pdl(slic_tile_kernel<true, SPX_MSLIC>, grid, dim3(NT), st, t, g, a, d_img, d_labels, parity, d_overflow));
        else SPX_CUDA(launch_pdl(slic_tile_kernel<false, SPX_MSLIC>, grid, dim3(NT), st, t, g, a, d_img, d_labels, parity, d_overflow));
    } else {
        if (p2) SPX_CUDA(launch_pdl(slic_tile_kernel<true, SPX_SLIC>, grid, dim3(NT), st, t, g, a, d_img, d_labels, parity, d_overflow));
        else SPX_CUDA(launch_pdl(slic_tile_kernel<false, SPX_SLIC>, grid, dim3(NT), st, t, g, a, d_img, d_labels, parity, d_overflow));
    }
    SPX_CUDA(cudaGetLastError());
    return SPX_OK;
}

// test hook: count inputs for which the Markstein sequence differs from the IEEE division
__global__ void div_selftest_kernel(float w, float r, unsigned seed, int n, int* bad)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned h = seed + (unsigned)i * 2654435761u;
    h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16;
    // spatial terms are sums of two squares of offsets up to a few thousand pixels: sample [0, 2^26) densely in mantissa
    float aa = __uint_as_float((h & 0x007fffffu) | ((100u + (h >> 23) % 53u) << 23));
    const u64 s = pack2(aa, aa * 0.37f);
    const u64 R2 = pack2(r, r), NW2 = pack2(-w, -w);
    const u64 q0 = mul2(s, R2);
    const u64 e = fma2(q0, NW2, s);
    const u64 q = fma2(e, R2, q0);
    float q_lo, q_hi;
    unpack2(q, q_lo, q_hi);
    if (q_lo != __fdiv_rn(aa, w) || q_hi != __fdiv_rn(aa * 0.37f, w)) atomicAdd(bad, 1);
}

}  // namespace spx

extern "C" int spx_selftest_division(float w, int n, int device, int* mismatches) try
{
    using namespace spx;
    SPX_REQUIRE(mismatches && n > 0, SPX_ERR_BADARG, "selftest_division: bad argument");
    SPX_CHECK(require_device(device));
    int* d = nullptr;
    SPX_CUDA(cudaMalloc(&d, sizeof(int)));
    SPX_CUDA(cudaMemset(d, 0, sizeof(int)));
    div_selftest_kernel<<<cdiv(n, 256), 256>>>(w, 1.0f / w, 12345u, n, d);
    SPX_CUDA(cudaMemcpy(mismatches, d, sizeof(int), cudaMemcpyDeviceToHost));
    cudaFree(d);
    return SPX_OK;
} SPX_API_CATCH
