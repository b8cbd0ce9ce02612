// stubs.cu -- entry points not built yet (SEEDS): fail loudly, never fall back to a CPU path.
#include "common.cuh"
#define NOTYET(name) { spx::set_error(name ": not implemented in this build"); return SPX_ERR_INTERNAL; }
extern "C" {
int spx_seeds_create(int, int, int, int, int, int, int, int, int, spx_seeds_t**) NOTYET("createSuperpixelSEEDS")
int spx_seeds_iterate(spx_seeds_t*, const uint8_t*, size_t, int) NOTYET("SuperpixelSEEDS::iterate")
int spx_seeds_get_number_of_superpixels(spx_seeds_t*, int*) NOTYET("SuperpixelSEEDS::getNumberOfSuperpixels")
int spx_seeds_get_labels(spx_seeds_t*, int32_t*, size_t) NOTYET("SuperpixelSEEDS::getLabels")
int spx_seeds_get_label_contour_mask(spx_seeds_t*, uint8_t*, size_t, int) NOTYET("SuperpixelSEEDS::getLabelContourMask")
int spx_seeds_destroy(spx_seeds_t*) { return SPX_OK; }
}
