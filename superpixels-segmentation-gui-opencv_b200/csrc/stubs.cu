// stubs.cu -- entry points not built yet (LSC, SEEDS): fail loudly, never fall back to a CPU path.
#include "common.cuh"
#define NOTYET(name) { spx::set_error(name ": not implemented in this build"); return SPX_ERR_INTERNAL; }
extern "C" {
int spx_lsc_create(const uint8_t*, size_t, int, int, int, int, float, int, spx_lsc_t**) NOTYET("createSuperpixelLSC")
int spx_lsc_iterate(spx_lsc_t*, int) NOTYET("SuperpixelLSC::iterate")
int spx_lsc_enforce_label_connectivity(spx_lsc_t*, int) NOTYET("SuperpixelLSC::enforceLabelConnectivity")
int spx_lsc_get_number_of_superpixels(spx_lsc_t*, int*) NOTYET("SuperpixelLSC::getNumberOfSuperpixels")
int spx_lsc_get_labels(spx_lsc_t*, int32_t*, size_t) NOTYET("SuperpixelLSC::getLabels")
int spx_lsc_get_label_contour_mask(spx_lsc_t*, uint8_t*, size_t, int) NOTYET("SuperpixelLSC::getLabelContourMask")
int spx_lsc_set_labels(spx_lsc_t*, const int32_t*, size_t, int) NOTYET("SuperpixelLSC::setLabels")
int spx_lsc_destroy(spx_lsc_t*) { return SPX_OK; }
int spx_seeds_create(int, int, int, int, int, int, int, int, int, spx_seeds_t**) NOTYET("createSuperpixelSEEDS")
int spx_seeds_iterate(spx_seeds_t*, const uint8_t*, size_t, int) NOTYET("SuperpixelSEEDS::iterate")
int spx_seeds_get_number_of_superpixels(spx_seeds_t*, int*) NOTYET("SuperpixelSEEDS::getNumberOfSuperpixels")
int spx_seeds_get_labels(spx_seeds_t*, int32_t*, size_t) NOTYET("SuperpixelSEEDS::getLabels")
int spx_seeds_get_label_contour_mask(spx_seeds_t*, uint8_t*, size_t, int) NOTYET("SuperpixelSEEDS::getLabelContourMask")
int spx_seeds_destroy(spx_seeds_t*) { return SPX_OK; }
}
