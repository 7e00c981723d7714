// placeholder until the IVF kernels land
#include "ivf.h"
namespace b2vs {
int ivf_search(b2vs_index*, const void*, int, int, int, const b2vs_search_params&, float*, int64_t*, cudaStream_t) { set_error("ivf not built yet"); return B2VS_EUNSUP; }
void ivf_fill_info(const b2vs_index*, b2vs_index_info*) {}
void ivf_last_stats(const b2vs_index*, b2vs_search_stats* s) { *s = b2vs_search_stats{}; }
void ivf_destroy(b2vs_index*) {}
}
using namespace b2vs;
extern "C" int b2vs_ivfflat_build(int, int, int, int, const void*, int64_t, int64_t, const b2vs_ivf_params*, void*, b2vs_index**) { set_error("ivf not built yet"); return B2VS_EUNSUP; }
extern "C" int b2vs_ivfpq_build(int, int, int, int, const void*, int64_t, int64_t, const b2vs_ivf_params*, void*, b2vs_index**) { set_error("ivf not built yet"); return B2VS_EUNSUP; }
extern "C" int b2vs_kmeans_fit(int, int, int, const void*, int64_t, int, int, uint64_t, float*, int32_t*, void*) { set_error("ivf not built yet"); return B2VS_EUNSUP; }
extern "C" int b2vs_ivf_list_sizes_host(const b2vs_index*, int32_t*) { return B2VS_EUNSUP; }
extern "C" int b2vs_ivf_centroids_host(const b2vs_index*, float*) { return B2VS_EUNSUP; }
