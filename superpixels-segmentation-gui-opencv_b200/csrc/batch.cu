// batch.cu -- batch of independent frames on one device (BASELINE.json configs[4]).
// One engine + one stream per slot; frames are dealt round-robin to the slots, so the small
// per-iteration kernels of one frame overlap the large kernels of the others.
#include "common.cuh"
#include <vector>

extern "C" int spx_slic_batch_dev(const void* const* d_frames, int n_frames, int w, int h, int C, int alg, int S, float ruler,
                                  int iters, int min_element_size, void* const* d_labels, void* const* d_masks,
                                  int thick_line, int n_streams, long long* launches_out)
{
    using namespace spx;
    SPX_REQUIRE(d_frames && d_labels && n_frames >= 0, SPX_ERR_BADARG, "slic_batch: bad argument");
    if (n_streams < 1) n_streams = 1;
    if (n_streams > n_frames) n_streams = n_frames > 0 ? n_frames : 1;
    std::vector<spx_slic_t*> eng(n_streams, nullptr);
    int rc = SPX_OK;
    long long launches = 0;
    for (int i = 0; i < n_frames && rc == SPX_OK; i++) {
        int s = i % n_streams;
        if (!eng[s]) rc = spx_slic_create_dev(d_frames[i], (size_t)w * C, w, h, C, alg, S, ruler, nullptr, &eng[s]);
        else rc = spx_slic_reset_dev(eng[s], d_frames[i], (size_t)w * C);
        if (rc == SPX_OK) rc = spx_slic_iterate(eng[s], iters);
        if (rc == SPX_OK && min_element_size > 0) rc = spx_slic_enforce_label_connectivity(eng[s], min_element_size);
        if (rc == SPX_OK) rc = spx_slic_get_labels_dev(eng[s], d_labels[i], (size_t)w * 4);
        if (rc == SPX_OK && d_masks) rc = spx_slic_get_label_contour_mask_dev(eng[s], d_masks[i], (size_t)w, thick_line);
    }
    for (int s = 0; s < n_streams; s++)
        if (eng[s]) {
            if (rc == SPX_OK) rc = spx_slic_synchronize(eng[s]);
            long long l = 0;
            spx_slic_launch_count(eng[s], &l);
            launches += l;
            spx_slic_destroy(eng[s]);
        }
    if (launches_out) *launches_out = launches;
    return rc;
}
