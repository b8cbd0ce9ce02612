// batch.cu -- batch of independent frames on one device (BASELINE.json configs[4]).
// One engine + one stream per slot; frames are dealt round-robin to the slots, so the small per-iteration
// kernels of one frame overlap the large kernels of the others.  Engines (device buffers, captured graphs)
// are kept in a per-process pool keyed by device and configuration, so a steady stream of batches pays
// no allocation or graph-capture cost after the first one.
#include "common.cuh"
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

namespace {
typedef std::tuple<int, int, int, int, int, int, float> PoolKey;   // device, w, h, C, alg, S, ruler
std::mutex g_mu;
std::map<PoolKey, std::vector<spx_slic_t*>> g_pool;
}  // namespace

extern "C" int spx_slic_batch_release(void) try
{
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto& kv : g_pool)
        for (spx_slic_t* e : kv.second) spx_slic_destroy(e);
    g_pool.clear();
    return SPX_OK;
} SPX_API_CATCH

extern "C" int spx_slic_batch_dev(const void* const* d_frames, int n_frames, int w, int h, int C, int alg, int S, float ruler,
                                  int iters, int min_element_size, void* const* d_labels, void* const* d_masks,
                                  int thick_line, int n_streams, long long* launches_out) try
{
    using namespace spx;
    SPX_REQUIRE(d_frames && d_labels && n_frames >= 0, SPX_ERR_BADARG, "slic_batch: bad argument");
    if (n_frames == 0) { if (launches_out) *launches_out = 0; return SPX_OK; }
    if (n_streams < 1) n_streams = 1;
    if (n_streams > n_frames) n_streams = n_frames;
    int dev = 0;
    SPX_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_mu);
    std::vector<spx_slic_t*>& eng = g_pool[PoolKey(dev, w, h, C, alg, S, ruler)];
    std::vector<long long> l0(n_streams, 0);
    std::vector<bool> fresh(n_streams, false);
    int rc = SPX_OK;
    for (int s = 0; s < n_streams && s < (int)eng.size(); s++) spx_slic_launch_count(eng[s], &l0[s]);
    for (int i = 0; i < n_frames && rc == SPX_OK; i++) {
        const int s = i % n_streams;
        if (s >= (int)eng.size()) {
            spx_slic_t* e = nullptr;
            rc = spx_slic_create_dev(d_frames[i], (size_t)w * C, w, h, C, alg, S, ruler, nullptr, &e);
            if (rc != SPX_OK) break;
            spx_slic_set_throughput_mode(e, 1);                    // several frames in flight: the instruction-lean centre update
            eng.push_back(e);
        } else {
            rc = spx_slic_reset_dev(eng[s], d_frames[i], (size_t)w * C);
        }
        if (rc == SPX_OK) rc = spx_slic_iterate(eng[s], iters);
        if (rc == SPX_OK && min_element_size > 0) rc = spx_slic_enforce_label_connectivity(eng[s], min_element_size);
        if (rc == SPX_OK) rc = spx_slic_get_labels_dev(eng[s], d_labels[i], (size_t)w * 4);
        if (rc == SPX_OK && d_masks) rc = spx_slic_get_label_contour_mask_dev(eng[s], d_masks[i], (size_t)w, thick_line);
    }
    long long launches = 0;
    for (int s = 0; s < n_streams && s < (int)eng.size(); s++) {
        int rs = spx_slic_synchronize(eng[s]);
        if (rc == SPX_OK) rc = rs;
        long long l = 0;
        spx_slic_launch_count(eng[s], &l);
        launches += l - l0[s];
    }
    if (launches_out) *launches_out = launches;
    return rc;
} SPX_API_CATCH
