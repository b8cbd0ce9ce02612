// slic_generic.cuh -- exact per-pixel assignment + accumulation for every SLIC-family configuration.
// One thread per pixel; walks the bin lists in global memory.  Used by the generic kernel (slic.cu) and
// as the in-kernel fallback of the tiled kernel (slic_tile.cu) when a tile has too many candidates.
#pragma once
#include "slic.cuh"
#include <cfloat>

namespace spx {

// must be called by all 32 lanes of a converged warp (warp-aggregated accumulation)
template <int ALG>   // 100 SLIC, 101 SLICO, 102 MSLIC
__device__ __forceinline__ void slic_pixel_generic(const SlicGeom& g, const SlicArrays& a, const unsigned char* __restrict__ img,
                                                   int* __restrict__ labels, float* __restrict__ dcplane, int parity,
                                                   int x, int y, bool inside)
{
    int label = -1;
    int pix[SPX_MAXC] = {0, 0, 0, 0};
    float lastdc = 0.f;
    if (inside) {
        const unsigned char* p = img + (size_t)y * g.ipitch + (size_t)x * g.C;
        float fpix[SPX_MAXC];
        for (int c = 0; c < g.C; c++) { pix[c] = p[c]; fpix[c] = (float)pix[c]; }
        const float fx = (float)x, fy = (float)y;
        const int* head = parity ? a.head[1] : a.head[0];
        int bx = x / g.S, by = y / g.S;
        float best = FLT_MAX;
        int bestn = -1, lastn = -1;
        for (int yy = max(by - g.radius, 0); yy <= min(by + g.radius, g.nby - 1); yy++)
            for (int xx = max(bx - g.radius, 0); xx <= min(bx + g.radius, g.nbx - 1); xx++)
                for (int n = head[yy * g.nbx + xx]; n >= 0; n = a.next[n]) {
                    int off = a.ioff[n];
                    int ix = a.ikx[n], iy = a.iky[n];
                    if (x < ix - off || x >= ix + off || y < iy - off || y >= iy + off) continue;
                    float dc = 0.f;
                    for (int c = 0; c < g.C; c++) {
                        float t = __fsub_rn(fpix[c], a.kc[(size_t)c * g.Kcap + n]);
                        dc = __fadd_rn(dc, __fmul_rn(t, t));
                    }
                    float ex = __fsub_rn(fx, a.kx[n]), ey = __fsub_rn(fy, a.ky[n]);
                    float dxy = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
                    float d;
                    if (ALG == 101) {
                        d = __fadd_rn(__fdiv_rn(dc, a.maxc[n]), __fdiv_rn(dxy, g.xywt));
                        if (n > lastn) { lastn = n; lastdc = dc; }
                    } else {
                        d = __fadd_rn(dc, __fdiv_rn(dxy, g.xywt));
                        if (ALG == 102) d = __fdiv_rn(d, a.adaptk[n]);
                    }
                    if (d < best || (d == best && n < bestn)) { best = d; bestn = n; }
                }
        size_t li = (size_t)y * g.lp + x;
        if (bestn >= 0) { labels[li] = bestn; label = bestn; }
        else label = labels[li];                       // covered by no window: keeps its label
        if (ALG == 101) {
            size_t pi = (size_t)y * g.lp + x;                      // the plane has the label pitch
            if (lastn >= 0) dcplane[pi] = lastdc; else lastdc = dcplane[pi];
        }
    }
    // warp-aggregated accumulation of (channels, x, y, count) and the SLICO maximum
    int key = inside ? label : -1 - (int)(threadIdx.x & 31);
    unsigned m = __match_any_sync(0xffffffffu, key);
    bool leader = inside && (int)(threadIdx.x & 31) == __ffs(m) - 1;
    for (int c = 0; c < g.C; c++) {
        int s = __reduce_add_sync(m, pix[c]);
        if (leader) atomicAdd(&a.acc[(size_t)c * g.Kcap + label], (unsigned long long)s);
    }
    int sx = __reduce_add_sync(m, inside ? x : 0), sy = __reduce_add_sync(m, inside ? y : 0);
    if (leader) {
        atomicAdd(&a.acc[(size_t)g.C * g.Kcap + label], (unsigned long long)sx);
        atomicAdd(&a.acc[(size_t)(g.C + 1) * g.Kcap + label], (unsigned long long)sy);
        atomicAdd(&a.acc[(size_t)(g.C + 2) * g.Kcap + label], (unsigned long long)__popc(m));
    }
    if (ALG == 101) {
        unsigned mx = __reduce_max_sync(m, __float_as_uint(lastdc));   // dc >= 0: float order == uint order
        if (leader) atomicMax(&a.maxc_acc[label], mx);
    }
}

}  // namespace spx
