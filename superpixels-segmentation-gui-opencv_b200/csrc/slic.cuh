// slic.cuh -- device-side state of one SuperpixelSLIC object (SLIC / SLICO / MSLIC).
#pragma once
#include "common.cuh"

namespace spx {

#define SPX_MAXC 4

// centre / accumulator arrays, all of capacity Kcap
struct SlicArrays {
    float* kx; float* ky;            // centre position (float32 means, as upstream)
    float* kc;                       // centre colour, channel-major: kc[c*Kcap + k]
    int* ikx; int* iky; int* ioff;   // (int)kx, (int)ky, window half-size (S, or int(S*adaptk) for MSLIC)
    float* adaptk;                   // MSLIC scale
    float* maxc; float* maxxy;       // SLICO running maxima (read by the assignment of this iteration)
    unsigned* maxc_acc; unsigned* maxxy_acc;   // SLICO maxima being accumulated (float bits, >= 0)
    unsigned long long* acc;         // accumulators, field-major: acc[f*Kcap + k], f = c0..cC-1, x, y, count
    int* next;                       // per-bin linked lists of centres
    int* head[2];                    // bin heads, double buffered by iteration parity
    long long* area;                 // MSLIC fixed-point manifold area per cluster
    int* ms_scratch;                 // MSLIC adapt step: [0..1] total area (64 bit), [2] splits, [3] new K, [4..] splits per block
    int* tl_count[2];                // tiled path: candidates per 32x32 tile, double buffered like head[]
    struct TileRec* tl_rec[2];       // tiled path: candidate records, TILE_CAP per tile
};

// one candidate of a tile: everything the tiled assignment kernel needs about a cluster (32 bytes)
struct TileRec { float kx, ky, k0, k1, k2; int n, ix, iy; };
constexpr int TILE_W = 32;           // tile of the tiled assignment kernel
constexpr int TILE_H = 32;
constexpr int TILE_CAP = 48;         // candidate capacity per tile (more -> exact per-pixel fallback for that tile)

struct SlicGeom {
    int w, h, C, S, K, Kcap;
    int nbx, nby;                    // bins of S x S pixels
    int lp;                          // label pitch (elements), a multiple of TILE_W; rows are padded to the tile grid
    int p4;                          // pitch (pixels) of the padded Lab1 word image of the tiled path
    size_t ipitch;                   // image pitch (bytes)
    int alg;                         // 100/101/102
    int radius;                      // bin search radius (1, or more for MSLIC)
    float xywt;                      // SLIC/MSLIC: (S/ruler)^2 ; SLICO: S*S
    int tiles_x, tiles_y;            // grid of TILE_W x TILE_H tiles
    int tile_on;                     // the tiled path is in use: centres are also scattered into the tile lists
    int ty_lo, ty_hi;                // tile rows this object assigns (strip mode; [0, tiles_y) otherwise)
    int own_y0, own_y1;              // pixel rows whose seeds this object owns (strip mode; [0, h) otherwise)
};

}  // namespace spx
