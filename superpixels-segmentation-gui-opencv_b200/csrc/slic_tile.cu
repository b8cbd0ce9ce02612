// slic_tile.cu -- tiled fast path of the SLIC-family assignment (3-channel images).  (stub: filled in next)
#include "slic.cuh"
namespace spx {
bool slic_tile_supported(const SlicGeom&) { return false; }
int slic_tile_assign(const SlicGeom&, const SlicArrays&, const unsigned char*, int*, const int*, int, int*, cudaStream_t)
{
    set_error("tiled SLIC path not available");
    return SPX_ERR_INTERNAL;
}
}  // namespace spx
