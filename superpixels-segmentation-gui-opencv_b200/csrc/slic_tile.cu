// slic_tile.cu -- tiled fast path of the SLIC assignment + centre accumulation (3-channel, SLIC=100).
//
// One CTA owns a 64x32 pixel tile.
//   setup : the CTA collects the clusters whose window [int(k)-S, int(k)+S) meets the tile from the
//           centre bins, sorts them by cluster index (upstream visits clusters in ascending n with a
//           strict '<', so ascending order reproduces its tie-breaking), stages their centres in shared
//           memory and tabulates the separable spatial terms ex2[c][x], ey2[c][y] (+inf outside the
//           window, so that an out-of-window candidate can never win).
//   main  : a warp handles an 8x8 patch, 2 horizontally adjacent pixels per thread, and visits only the
//           candidates whose window meets the patch (5.8 per pixel on average at S=20; 4.0 is the
//           minimum).  The distance arithmetic is upstream's float32 sequence, evaluated with the sm_100
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

constexpr int TW = 64;      // tile width  (8 warps x 8 px)
constexpr int TH = 32;      // tile height (4 steps x 8 rows)
constexpr int CAP = 64;     // candidate capacity per tile
constexpr int MIN_S = 14;   // below this the candidate lists outgrow CAP: generic kernel instead

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

struct TileSmem {
    u64 cenA[CAP][2];          // (-kc0,-kc0), (-kc1,-kc1)
    u64 cenB[CAP];             // (-kc2,-kc2)
    float ex[CAP][TW];         // (x-kx)^2 [scaled by 1/xywt when it is a power of two], +inf outside the window
    u64 ey[CAP][TH];           // ((y-ky)^2, same) duplicated into both halves
    int4 win[CAP];             // window: x_lo, x_hi, y_lo, y_hi (half open)
    unsigned acc[CAP][8];      // L, a, count... see FIELD order below
    int n[CAP];                // cluster index per slot, ascending
    int list[CAP];
    int cnt;
};
// shared accumulator fields: 0 c0, 1 c1, 2 c2, 3 count, 4 x (tile relative), 5 y (tile relative)

// all 32 lanes call; slot < 0 = nothing to add.  wA = c0 | c1<<16, wB = c2 | count<<16, wC = xrel | yrel<<16
__device__ __forceinline__ void warp_accumulate(unsigned (*acc)[8], int slot, unsigned wA, unsigned wB, unsigned wC, int lane)
{
    unsigned rem = __ballot_sync(0xffffffffu, slot >= 0);
    while (rem) {
        const int leader = __ffs(rem) - 1;
        const int cur = __shfl_sync(0xffffffffu, slot, leader);
        const bool eq = slot == cur;
        const unsigned m = __ballot_sync(0xffffffffu, eq);
        const unsigned A = __reduce_add_sync(0xffffffffu, eq ? wA : 0u);
        const unsigned B = __reduce_add_sync(0xffffffffu, eq ? wB : 0u);
        const unsigned Cc = __reduce_add_sync(0xffffffffu, eq ? wC : 0u);
        if (lane < 6) {
            const unsigned wsel = lane < 2 ? A : lane < 4 ? B : Cc;
            const unsigned v = (lane & 1) ? (wsel >> 16) : (wsel & 0xffffu);
            atomicAdd(&acc[cur][lane], v);
        }
        rem &= ~m;
    }
}

template <bool POW2>
__global__ void __launch_bounds__(256, 3) slic_tile_kernel(SlicGeom g, SlicArrays a, const unsigned char* __restrict__ img,
                                                            int* __restrict__ labels, int parity, float rcp_xywt,
                                                            int* __restrict__ d_overflow)
{
    __shared__ TileSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = g.S;
    const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH;
    const int tx1 = min(tx0 + TW, g.w), ty1 = min(ty0 + TH, g.h);
    if (tid == 0) sm.cnt = 0;
    __syncthreads();

    // ---- candidates: clusters with int(kx) in (tx0-S, tx1+S) and int(ky) in (ty0-S, ty1+S)
    {
        const int bx0 = max(tx0 - S + 1, 0) / S, bx1 = min((tx1 + S - 1) / S, g.nbx - 1);
        const int by0 = max(ty0 - S + 1, 0) / S, by1 = min((ty1 + S - 1) / S, g.nby - 1);
        const int nbw = bx1 - bx0 + 1, nbins = nbw * (by1 - by0 + 1);
        const int* head = a.head[parity];
        for (int b = tid; b < nbins; b += 256) {
            const int by = by0 + b / nbw, bx = bx0 + b % nbw;
            for (int n = head[by * g.nbx + bx]; n >= 0; n = a.next[n]) {
                const int ix = a.ikx[n], iy = a.iky[n];
                if (ix - S < tx1 && ix + S > tx0 && iy - S < ty1 && iy + S > ty0) {
                    const int i = atomicAdd(&sm.cnt, 1);
                    if (i < CAP) sm.list[i] = n;
                }
            }
        }
    }
    __syncthreads();
    const int cnt = sm.cnt;
    if (cnt > CAP) {
        // too many candidates for the shared-memory tables: exact per-pixel path for this tile
        if (tid == 0) atomicAdd(d_overflow, 1);
        for (int r = warp; r < TH; r += 8)
            for (int c0 = 0; c0 < TW; c0 += 32) {
                const int x = tx0 + c0 + lane, y = ty0 + r;
                slic_pixel_generic<100>(g, a, img, labels, nullptr, parity, x, y, x < g.w && y < g.h);
            }
        return;
    }
    // ---- sort by cluster index, stage centres
    if (tid < cnt) {
        const int n = sm.list[tid];
        int rank = 0;
        for (int j = 0; j < cnt; j++) rank += sm.list[j] < n ? 1 : 0;
        sm.n[rank] = n;
        const float k0 = -a.kc[n], k1 = -a.kc[(size_t)g.Kcap + n], k2 = -a.kc[2 * (size_t)g.Kcap + n];
        sm.cenA[rank][0] = pack2(k0, k0);
        sm.cenA[rank][1] = pack2(k1, k1);
        sm.cenB[rank] = pack2(k2, k2);
        const int ix = a.ikx[n], iy = a.iky[n];
        sm.win[rank] = make_int4(ix - S, ix + S, iy - S, iy + S);
    }
    for (int i = tid; i < cnt * 8; i += 256) (&sm.acc[0][0])[i] = 0u;
    __syncthreads();
    // ---- separable spatial tables
    const float inf = __int_as_float(0x7f800000);
    for (int i = tid; i < cnt * TW; i += 256) {
        const int c = i >> 6, xx = i & (TW - 1);
        const int x = tx0 + xx;
        const int4 wn = sm.win[c];
        const float e = __fsub_rn((float)x, a.kx[sm.n[c]]);
        float v = __fmul_rn(e, e);
        if (POW2) v = __fmul_rn(v, rcp_xywt);
        sm.ex[c][xx] = (x >= wn.x && x < wn.y) ? v : inf;
    }
    for (int i = tid; i < cnt * TH; i += 256) {
        const int c = i >> 5, yy = i & (TH - 1);
        const int y = ty0 + yy;
        const int4 wn = sm.win[c];
        const float e = __fsub_rn((float)y, a.ky[sm.n[c]]);
        float v = __fmul_rn(e, e);
        if (POW2) v = __fmul_rn(v, rcp_xywt);
        v = (y >= wn.z && y < wn.w) ? v : inf;
        sm.ey[c][yy] = pack2(v, v);
    }
    __syncthreads();

    // ---- main: warp = 8 px x 8 rows per step, 2 px per thread
    const int lx = lane & 3, ly = lane >> 2;
    const int xo = 8 * warp + 2 * lx;          // tile-relative x of the thread's first pixel
    const int x = tx0 + xo;
    const u64 R2 = pack2(rcp_xywt, rcp_xywt);
    const float nw = -g.xywt;
    const u64 NW2 = pack2(nw, nw);
    const int X0 = tx0 + 8 * warp, X1 = X0 + 8;
#pragma unroll 1
    for (int step = 0; step < TH / 8; step++) {
        const int yo = 8 * step + ly;
        const int y = ty0 + yo;
        const bool in0 = x < g.w && y < g.h, in1 = x + 1 < g.w && y < g.h;
        unsigned b01 = 0, b23 = 0, b45 = 0;
        if (y < g.h) {
            const unsigned short* p = reinterpret_cast<const unsigned short*>(img + (size_t)y * g.ipitch + (size_t)x * 3);
            b01 = p[0]; b23 = p[1]; b45 = p[2];
        }
        const int c00 = b01 & 255, c01 = b01 >> 8, c02 = b23 & 255, c10 = b23 >> 8, c11 = b45 & 255, c12 = b45 >> 8;
        const u64 P0 = pack2((float)c00, (float)c10), P1 = pack2((float)c01, (float)c11), P2 = pack2((float)c02, (float)c12);
        // candidates whose window meets this warp's patch
        const int Y0 = ty0 + 8 * step, Y1 = Y0 + 8;
        unsigned m0, m1 = 0;
        {
            int4 wn = sm.win[lane];
            m0 = __ballot_sync(0xffffffffu, lane < cnt && wn.x < X1 && wn.y > X0 && wn.z < Y1 && wn.w > Y0);
            if (cnt > 32) {
                wn = sm.win[lane + 32];
                m1 = __ballot_sync(0xffffffffu, lane + 32 < cnt && wn.x < X1 && wn.y > X0 && wn.z < Y1 && wn.w > Y0);
            }
        }
        float best0 = FLT_MAX, best1 = FLT_MAX;
        int slot0 = -1, slot1 = -1;
#pragma unroll 1
        for (int half = 0; half < 2; half++) {
            unsigned m = half ? m1 : m0;
            while (m) {
                const int c = (__ffs(m) - 1) + 32 * half;
                m &= m - 1;
                const ulonglong2 ca = *reinterpret_cast<const ulonglong2*>(&sm.cenA[c][0]);
                const u64 cb = sm.cenB[c];
                const u64 exv = *reinterpret_cast<const u64*>(&sm.ex[c][xo]);
                const u64 eyv = sm.ey[c][yo];
                const u64 t0 = add2(P0, ca.x), t1 = add2(P1, ca.y), t2 = add2(P2, cb);
                u64 d = add2_products(mul2(t0, t0), mul2(t1, t1));
                d = add2_products(d, mul2(t2, t2));
                u64 s = add2(exv, eyv);
                if (!POW2) {
                    const u64 q0 = mul2(s, R2);
                    const u64 e = fma2(q0, NW2, s);
                    s = fma2(e, R2, q0);
                }
                d = add2(d, s);
                float d0, d1;
                unpack2(d, d0, d1);
                if (d0 < best0) { best0 = d0; slot0 = c; }
                if (d1 < best1) { best1 = d1; slot1 = c; }
            }
        }
        // ---- labels (a pixel no window covers keeps its label)
        int n0 = -1, n1 = -1;
        int* lrow = labels + (size_t)y * g.lp + x;
        if (in0) n0 = slot0 >= 0 ? sm.n[slot0] : lrow[0];
        if (in1) n1 = slot1 >= 0 ? sm.n[slot1] : lrow[1];
        if (in1) *reinterpret_cast<int2*>(lrow) = make_int2(n0, n1);
        else if (in0) lrow[0] = n0;
        // ---- accumulate (channels, count, x, y) per label
        {
            const int s0 = in0 ? slot0 : -1, s1 = in1 ? slot1 : -1;
            const bool same = s0 == s1;
            const unsigned a0 = (unsigned)c00 | ((unsigned)c01 << 16), b0 = (unsigned)c02 | (1u << 16), w0 = (unsigned)xo | ((unsigned)yo << 16);
            const unsigned a1 = (unsigned)c10 | ((unsigned)c11 << 16), b1 = (unsigned)c12 | (1u << 16), w1 = (unsigned)(xo + 1) | ((unsigned)yo << 16);
            warp_accumulate(sm.acc, s0, same ? a0 + a1 : a0, same ? b0 + b1 : b0, same ? w0 + w1 : w0, lane);
            const int sec = same ? -1 : s1;
            if (__any_sync(0xffffffffu, sec >= 0)) warp_accumulate(sm.acc, sec, a1, b1, w1, lane);
            // uncovered pixels (rare): straight to the global accumulators of the label they keep
            if (in0 && slot0 < 0) {
                atomicAdd(&a.acc[n0], (u64)c00); atomicAdd(&a.acc[(size_t)g.Kcap + n0], (u64)c01);
                atomicAdd(&a.acc[2 * (size_t)g.Kcap + n0], (u64)c02); atomicAdd(&a.acc[3 * (size_t)g.Kcap + n0], (u64)x);
                atomicAdd(&a.acc[4 * (size_t)g.Kcap + n0], (u64)y); atomicAdd(&a.acc[5 * (size_t)g.Kcap + n0], 1ull);
            }
            if (in1 && slot1 < 0) {
                atomicAdd(&a.acc[n1], (u64)c10); atomicAdd(&a.acc[(size_t)g.Kcap + n1], (u64)c11);
                atomicAdd(&a.acc[2 * (size_t)g.Kcap + n1], (u64)c12); atomicAdd(&a.acc[3 * (size_t)g.Kcap + n1], (u64)(x + 1));
                atomicAdd(&a.acc[4 * (size_t)g.Kcap + n1], (u64)y); atomicAdd(&a.acc[5 * (size_t)g.Kcap + n1], 1ull);
            }
        }
    }
    __syncthreads();
    // ---- flush: one global atomic per field per candidate that received pixels
    for (int i = tid; i < cnt * 6; i += 256) {
        const int c = i / 6, f = i - 6 * c;
        const unsigned count = sm.acc[c][3];
        if (count == 0) continue;
        const int n = sm.n[c];
        u64 v = sm.acc[c][f];
        int field = f;                               // global order: c0, c1, c2, x, y, count
        if (f == 3) field = 5;
        else if (f == 4) { field = 3; v += (u64)count * (u64)tx0; }
        else if (f == 5) { field = 4; v += (u64)count * (u64)ty0; }
        atomicAdd(&a.acc[(size_t)field * g.Kcap + n], v);
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
    dim3 grid(cdiv(g.w, TW), cdiv(g.h, TH));
    const float r = 1.0f / g.xywt;
    if (is_pow2(g.xywt)) slic_tile_kernel<true><<<grid, 256, 0, st>>>(g, a, d_img, d_labels, parity, r, d_overflow);
    else slic_tile_kernel<false><<<grid, 256, 0, st>>>(g, a, d_img, d_labels, parity, r, d_overflow);
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
