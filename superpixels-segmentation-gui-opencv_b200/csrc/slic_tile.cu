// slic_tile.cu -- tiled fast path of the SLIC assignment + centre accumulation (3-channel, SLIC=100).
//
// One CTA (128 threads) owns a 32x32 pixel tile.
//   setup : the centre-update kernel has appended to the tile's list every cluster whose window
//           [int(k)-S, int(k)+S) meets the tile; the CTA reads that list (one coalesced 32-byte record per
//           cluster), sorts it by cluster index (upstream visits clusters in ascending n with a
//           strict '<', so ascending order reproduces its tie-breaking), stages their centres in shared
//           memory and tabulates the separable spatial terms ex2[c][x], ey2[c][y] (+inf outside the
//           window, so that an out-of-window candidate can never win).
//   main  : a warp handles a 16x8 patch, 2x2 pixels per thread, and visits only the candidates whose
//           window meets the patch (6.7 per pixel on average at S=20; 4.0 is the minimum).  The distance arithmetic is upstream's float32 sequence, evaluated with the sm_100
//           packed f32x2 instructions (FADD2/FMUL2/FFMA2: same rounding per element, half the issue
//           slots), never contracted:   d = ((t0*t0 + t1*t1) + t2*t2) + (ex2+ey2)/xywt.
//           The IEEE division by the constant xywt is a multiply when xywt is a power of two, otherwise
//           Markstein's correctly rounded q = a*r; q += fma(-q, w, a)*r with r = RN(1/w).
//   accum : per-warp REDUX sums of (L,a,b,x,y,1) per label, shared-memory accumulators per candidate,
//           one 64-bit global atomic per field per candidate per CTA.
// Algorithmic traffic: 3 B/px read + 4 B/px written per iteration.
#include "slic.cuh"
#include "slic_generic.cuh"
#include <cfloat>
#include <cmath>

namespace spx {

constexpr int TW = TILE_W;   // 32: two warps side by side, 16 px each
constexpr int TH = TILE_H;   // 32: two warp rows x 8 rows x two steps
constexpr int CAP = TILE_CAP;
constexpr int NT = 128;      // threads per CTA
constexpr int MIN_S = 14;    // below this the candidate lists outgrow CAP too often: generic kernel instead

typedef unsigned long long u64;

__device__ __forceinline__ u64 pack2(float lo, float hi)
{
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 add2(u64 a, u64 b)
{
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// add of two PRODUCTS.  ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with -fmad=false (scalar
// mul.rn/add.rn are left alone, as the PTX ISA promises); a flush-to-zero mismatch between the two stops it.  The
// operands here are squares of differences of an 8-bit value and a float32 mean (zero or >= 2^-48, never subnormal),
// so .ftz does not change any result.  tests/test_abi.py checks the SASS for stray FFMA2.
__device__ __forceinline__ u64 add2_products(u64 a, u64 b)
{
    u64 r;
    asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b)
{
    u64 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c)
{
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// ---- shared memory layout -----------------------------------------------------------------------
// One 416-byte record per candidate, so that a single record offset addresses everything:
//   +0   (-kc0,-kc0) (-kc1,-kc1)      16 B      +16  (-kc2,-kc2)  8 B   (+24 pad)
//   +32  ex2[32]  float               128 B     (x-kx)^2 [pre-divided by xywt when it is a power of two], +inf outside
//   +160 ey2[32]  float2 (v,v)        256 B     (y-ky)^2 likewise, duplicated so a row's term is a ready f32x2 operand
constexpr int REC = 416;
constexpr int REC_EX = 32;
constexpr int REC_EY = 160;
struct TileSmem {
    __align__(16) unsigned char rec[CAP * REC];
    int4 win[CAP];             // window: x_lo, x_hi, y_lo, y_hi (half open)
    unsigned acc[CAP][8];      // 0 c0, 1 c1, 2 c2, 3 count, 4 x (tile relative), 5 y (tile relative)
    float kxy[CAP][2];         // centre position per slot
    int n[CAP];                // cluster index per slot, ascending
    int list[CAP];
};

struct TileArgs {
    const int* tl_count; const TileRec* tl_rec;
    unsigned long long* acc;
    int w, h, S, Kcap, tiles_x, lp;
    size_t ipitch;
    float xywt, rcp_xywt;
};

// per-label sums of one warp step.  Every thread contributes up to four pixels (slot s[j], words wA/wB/wC[j]:
// wA = c0 | c1<<16, wB = c2 | count<<16, wC = xrel | yrel<<16).  Labels are visited in ascending slot order:
// REDUX.MIN picks the next one (result lands in a uniform register), REDUX.SUM adds its three words over the
// warp, and lanes 0..5 add one field each to the CTA's shared accumulators.
__device__ __forceinline__ void warp_accumulate4(unsigned (*acc)[8], const int (&s)[4], const unsigned (&wA)[4],
                                                 const unsigned (&wB)[4], const unsigned (&wC)[4], int lane)
{
    // thread-local merge: everything equal to the first pixel's slot is summed up front (the common case)
    unsigned pA = wA[0], pB = wB[0], pC = wC[0];
    unsigned key0 = (unsigned)s[0];               // -1 -> 0xffffffff = nothing
    unsigned kj[3];
    bool left = false;
#pragma unroll
    for (int j = 1; j < 4; j++) {
        const bool e = s[j] == s[0];
        pA += e ? wA[j] : 0u; pB += e ? wB[j] : 0u; pC += e ? wC[j] : 0u;
        kj[j - 1] = e ? 0xffffffffu : (unsigned)s[j];
        left |= !e && s[j] >= 0;
    }
    // primary pass
    for (;;) {
        const unsigned cur = __reduce_min_sync(0xffffffffu, key0);
        if (cur == 0xffffffffu) break;
        const bool e = key0 == cur;
        const unsigned A = __reduce_add_sync(0xffffffffu, e ? pA : 0u);
        const unsigned B = __reduce_add_sync(0xffffffffu, e ? pB : 0u);
        const unsigned Cc = __reduce_add_sync(0xffffffffu, e ? pC : 0u);
        key0 = e ? 0xffffffffu : key0;
        if (lane < 6) {
            const unsigned wsel = lane < 2 ? A : lane < 4 ? B : Cc;
            atomicAdd(&acc[cur][lane], (lane & 1) ? (wsel >> 16) : (wsel & 0xffffu));
        }
    }
    // leftovers: pixels whose slot differs from the thread's first pixel (threads on a label boundary): few lanes,
    // so plain shared-memory atomics are cheaper than another round of warp reductions
    if (__any_sync(0xffffffffu, left)) {
#pragma unroll
        for (int j = 0; j < 3; j++) {
            if (kj[j] != 0xffffffffu) {
                unsigned* dst = acc[kj[j]];
                atomicAdd(dst + 0, wA[j + 1] & 0xffffu); atomicAdd(dst + 1, wA[j + 1] >> 16);
                atomicAdd(dst + 2, wB[j + 1] & 0xffffu); atomicAdd(dst + 3, 1u);
                atomicAdd(dst + 4, wC[j + 1] & 0xffffu); atomicAdd(dst + 5, wC[j + 1] >> 16);
            }
        }
    }
}

__device__ __forceinline__ void acc_global_pixel(const TileArgs& t, int n, int c0, int c1, int c2, int x, int y)
{
    atomicAdd(&t.acc[n], (u64)c0); atomicAdd(&t.acc[(size_t)t.Kcap + n], (u64)c1);
    atomicAdd(&t.acc[2 * (size_t)t.Kcap + n], (u64)c2); atomicAdd(&t.acc[3 * (size_t)t.Kcap + n], (u64)x);
    atomicAdd(&t.acc[4 * (size_t)t.Kcap + n], (u64)y); atomicAdd(&t.acc[5 * (size_t)t.Kcap + n], 1ull);
}

// exact per-pixel path for a tile whose candidate list overflowed (kept out of line: it is cold)
__device__ __noinline__ void tile_fallback(const SlicGeom& g, const SlicArrays& a, const unsigned char* img, int* labels,
                                           int parity, int tx0, int ty0, int warp, int lane)
{
    for (int r = warp; r < TH; r += NT / 32) {
        const int x = tx0 + lane, y = ty0 + r;
        slic_pixel_generic<100>(g, a, img, labels, nullptr, parity, x, y, x < g.w && y < g.h);
    }
}

template <bool POW2>
__global__ void __launch_bounds__(NT, 5) slic_tile_kernel(const __grid_constant__ TileArgs t, const __grid_constant__ SlicGeom g, const __grid_constant__ SlicArrays a,
                                                           const unsigned char* __restrict__ img, int* __restrict__ labels,
                                                           int parity, int* __restrict__ d_overflow)
{
    __shared__ TileSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = t.S;
    const int tile = blockIdx.y * t.tiles_x + blockIdx.x;
    const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH;

    // ---- first pixels in flight before anything else (lanes 8 x 4, 2x2 pixels per thread, warps 2 x 2)
    const int lx = lane & 7, ly = lane >> 3;
    const int wx = warp & 1, wy = warp >> 1;
    const int xo = 16 * wx + 2 * lx;           // tile-relative x of the thread's left pixels
    const int x = tx0 + xo;
    unsigned raw[2][3];
    auto load_rows = [&](int yy, unsigned (&dst)[2][3]) {
#pragma unroll
        for (int r = 0; r < 2; r++) {
            dst[r][0] = dst[r][1] = dst[r][2] = 0;
            if (yy + r < t.h) {
                const unsigned short* p = reinterpret_cast<const unsigned short*>(img + (size_t)(yy + r) * t.ipitch + (size_t)x * 3);
                dst[r][0] = __ldg(p); dst[r][1] = __ldg(p + 1); dst[r][2] = __ldg(p + 2);
            }
        }
    };
    load_rows(ty0 + 8 * wy + 2 * ly, raw);

    // ---- candidates: the centre-update kernel appended every cluster whose window touches this tile
    const int cnt = t.tl_count[tile];
    if (cnt > CAP) {
        if (tid == 0) atomicAdd(d_overflow, 1);
        tile_fallback(g, a, img, labels, parity, tx0, ty0, warp, lane);
        return;
    }
    float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0;
    if (tid < cnt) {
        const float4* src = reinterpret_cast<const float4*>(&t.tl_rec[(size_t)tile * CAP + tid]);
        r0 = src[0]; r1 = src[1];
        sm.list[tid] = __float_as_int(r1.y);
    }
    __syncthreads();
    // ---- sort by cluster index (upstream's visiting order), stage centres
    if (tid < cnt) {
        const int n = __float_as_int(r1.y);
        int rank = 0;
        for (int j = 0; j < cnt; j++) rank += sm.list[j] < n ? 1 : 0;
        sm.n[rank] = n;
        u64* r64 = reinterpret_cast<u64*>(sm.rec + rank * REC);
        r64[0] = pack2(-r0.z, -r0.z); r64[1] = pack2(-r0.w, -r0.w); r64[2] = pack2(-r1.x, -r1.x);
        const int ix = __float_as_int(r1.z), iy = __float_as_int(r1.w);
        sm.win[rank] = make_int4(ix - S, ix + S, iy - S, iy + S);
        sm.kxy[rank][0] = r0.x; sm.kxy[rank][1] = r0.y;
    }
    for (int i = tid; i < cnt * 8; i += NT) (&sm.acc[0][0])[i] = 0u;
    __syncthreads();
    // ---- separable spatial tables: thread (c, j) fills ex2[c][j] and ey2[c][j]
    {
        const float inf = __int_as_float(0x7f800000);
        for (int i = tid; i < cnt * 32; i += NT) {
            const int c = i >> 5, j = i & 31;
            const int4 wn = sm.win[c];
            const int px = tx0 + j, py = ty0 + j;
            const float ex = __fsub_rn((float)px, sm.kxy[c][0]), ey = __fsub_rn((float)py, sm.kxy[c][1]);
            float vx = __fmul_rn(ex, ex), vy = __fmul_rn(ey, ey);
            if (POW2) { vx = __fmul_rn(vx, t.rcp_xywt); vy = __fmul_rn(vy, t.rcp_xywt); }
            vx = (px >= wn.x && px < wn.y) ? vx : inf;
            vy = (py >= wn.z && py < wn.w) ? vy : inf;
            reinterpret_cast<float*>(sm.rec + c * REC + REC_EX)[j] = vx;
            reinterpret_cast<u64*>(sm.rec + c * REC + REC_EY)[j] = pack2(vy, vy);
        }
    }
    __syncthreads();

    // ---- main: a warp owns a 16x8 patch per step; the CTA's 4 warps cover 32x16, two steps cover the tile
    const u64 R2 = pack2(t.rcp_xywt, t.rcp_xywt);
    const float nw = -t.xywt;
    const u64 NW2 = pack2(nw, nw);
    const int X0 = tx0 + 16 * wx, X1 = X0 + 16;
    const unsigned char* thr_ex = sm.rec + REC_EX + xo * 4;
#pragma unroll 1
    for (int step = 0; step < TH / 16; step++) {
        const int yo = 16 * step + 8 * wy + 2 * ly;
        const int y = ty0 + yo;
        unsigned cur[2][3];
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int k = 0; k < 3; k++) cur[r][k] = raw[r][k];
        if (step + 1 < TH / 16) load_rows(y + 16, raw);      // prefetch the next step's pixels
        // unpack without the (quarter-rate) I2F unit: byte -> float as (0x4B0000bb as float) - 2^23, built with PRMT.
        // w01 = [c00 c01 c02 c10], w2 = [c11 c12 0 0]   (pixel (r,0) = bytes 0..2, pixel (r,1) = bytes 3..5)
        u64 P[2][3];
        unsigned w01[2], w2[2];
        {
            const unsigned MAGIC = 0x4B000100u;          // bytes: 00 01 00 4B
            const u64 M2 = pack2(-8388608.0f, -8388608.0f);
#pragma unroll
            for (int r = 0; r < 2; r++) {
                w01[r] = cur[r][0] | (cur[r][1] << 16);
                w2[r] = cur[r][2];
                const float a0 = __uint_as_float(__byte_perm(w01[r], MAGIC, 0x7640)), b0 = __uint_as_float(__byte_perm(w01[r], MAGIC, 0x7643));
                const float a1 = __uint_as_float(__byte_perm(w01[r], MAGIC, 0x7641)), b1 = __uint_as_float(__byte_perm(w2[r], MAGIC, 0x7640));
                const float a2 = __uint_as_float(__byte_perm(w01[r], MAGIC, 0x7642)), b2 = __uint_as_float(__byte_perm(w2[r], MAGIC, 0x7641));
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
        int slot[2][2] = {{-1, -1}, {-1, -1}};
        const unsigned char* thr_ey = sm.rec + REC_EY + yo * 8;
#pragma unroll 1
        for (int half = 0; half < 2; half++) {
            unsigned m = half ? m1 : m0;
            while (m) {
                const int c = (__ffs(m) - 1) + 32 * half;
                m &= m - 1;
                const int ro = c * REC;
                const ulonglong2 ca = *reinterpret_cast<const ulonglong2*>(sm.rec + ro);
                const u64 cb = *reinterpret_cast<const u64*>(sm.rec + ro + 16);
                const u64 exv = *reinterpret_cast<const u64*>(thr_ex + ro);
                const ulonglong2 eyv = *reinterpret_cast<const ulonglong2*>(thr_ey + ro);
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    const u64 t0 = add2(P[r][0], ca.x), t1 = add2(P[r][1], ca.y), t2 = add2(P[r][2], cb);
                    u64 d = add2_products(mul2(t0, t0), mul2(t1, t1));
                    d = add2_products(d, mul2(t2, t2));
                    u64 sxy = add2(exv, r ? eyv.y : eyv.x);
                    if (!POW2) {
                        const u64 q0 = mul2(sxy, R2);
                        const u64 e = fma2(q0, NW2, sxy);
                        sxy = fma2(e, R2, q0);
                    }
                    d = add2(d, sxy);
                    float d0, d1;
                    unpack2(d, d0, d1);
                    if (d0 < best[r][0]) { best[r][0] = d0; slot[r][0] = c; }
                    if (d1 < best[r][1]) { best[r][1] = d1; slot[r][1] = c; }
                }
            }
        }
        // ---- labels (a pixel no window covers keeps its label) and accumulation words
        int sj[4];
        unsigned wA[4], wB[4], wC[4];
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const int yy = y + r;
            int* lrow = labels + (size_t)yy * t.lp + x;
            const bool in0 = x < t.w && yy < t.h, in1 = x + 1 < t.w && yy < t.h;
            int n0 = -1, n1 = -1;
            if (in0) n0 = slot[r][0] >= 0 ? sm.n[slot[r][0]] : lrow[0];
            if (in1) n1 = slot[r][1] >= 0 ? sm.n[slot[r][1]] : lrow[1];
            if (in1) *reinterpret_cast<int2*>(lrow) = make_int2(n0, n1);
            else if (in0) lrow[0] = n0;
            const unsigned MAGIC = 0x4B000100u;
            wA[2 * r] = __byte_perm(w01[r], MAGIC, 0x4140);          // c00 | c01 << 16
            wB[2 * r] = __byte_perm(w01[r], MAGIC, 0x4542);          // c02 | 1 << 16
            wA[2 * r + 1] = __byte_perm(w01[r], w2[r], 0x6463);      // c10 | c11 << 16
            wB[2 * r + 1] = __byte_perm(w2[r], MAGIC, 0x4541);       // c12 | 1 << 16
            wC[2 * r] = (unsigned)xo | ((unsigned)(yo + r) << 16);
            wC[2 * r + 1] = wC[2 * r] + 1u;
            sj[2 * r] = in0 ? slot[r][0] : -1;
            sj[2 * r + 1] = in1 ? slot[r][1] : -1;
            if (in0 && slot[r][0] < 0) acc_global_pixel(t, n0, wA[2 * r] & 0xffff, wA[2 * r] >> 16, wB[2 * r] & 0xffff, x, yy);
            if (in1 && slot[r][1] < 0) acc_global_pixel(t, n1, wA[2 * r + 1] & 0xffff, wA[2 * r + 1] >> 16, wB[2 * r + 1] & 0xffff, x + 1, yy);
        }
        warp_accumulate4(sm.acc, sj, wA, wB, wC, lane);
    }
    __syncthreads();
    // ---- flush: one global atomic per field per candidate that received pixels
    for (int i = tid; i < cnt * 6; i += NT) {
        const int c = i / 6, f = i - 6 * c;
        const unsigned count = sm.acc[c][3];
        if (count == 0) continue;
        const int n = sm.n[c];
        u64 v = sm.acc[c][f];
        int field = f;                               // global order: c0, c1, c2, x, y, count
        if (f == 3) field = 5;
        else if (f == 4) { field = 3; v += (u64)count * (u64)tx0; }
        else if (f == 5) { field = 4; v += (u64)count * (u64)ty0; }
        atomicAdd(&t.acc[(size_t)field * t.Kcap + n], v);
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
    return g.alg == SPX_SLIC && g.C == 3 && g.S >= MIN_S && (is_pow2(g.xywt) || markstein_ok(g.xywt));
}

int slic_tile_assign(const SlicGeom& g, const SlicArrays& a, const unsigned char* d_img, int* d_labels, const int* d_K,
                     int parity, int* d_overflow, cudaStream_t st)
{
    (void)d_K;
    dim3 grid(g.tiles_x, g.tiles_y);
    TileArgs t;
    t.tl_count = a.tl_count[parity]; t.tl_rec = a.tl_rec[parity]; t.acc = a.acc;
    t.w = g.w; t.h = g.h; t.S = g.S; t.Kcap = g.Kcap; t.tiles_x = g.tiles_x; t.lp = g.lp;
    t.ipitch = g.ipitch; t.xywt = g.xywt; t.rcp_xywt = 1.0f / g.xywt;
    if (is_pow2(g.xywt)) slic_tile_kernel<true><<<grid, NT, 0, st>>>(t, g, a, d_img, d_labels, parity, d_overflow);
    else slic_tile_kernel<false><<<grid, NT, 0, st>>>(t, g, a, d_img, d_labels, parity, d_overflow);
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

extern "C" int spx_selftest_division(float w, int n, int device, int* mismatches)
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
}
