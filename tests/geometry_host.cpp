// Host build of csrc/geometry_core.h for the CPU test suite (tests/test_geometry_core.py compiles it with g++):
// checks the per-grain arithmetic of gg_region_center against the reference's golden vectors without a GPU.
// TEST INFRASTRUCTURE ONLY — the product runs the same header inside geometry.cu on the device.
#include <cmath>
#include <cstddef>
#include <limits>
#include "../graingraphnn_b200/csrc/geometry_core.h"

struct FetchHost {
    const float* xj; int ld; const float* off; float factor; const int32_t* col;
    void operator()(int32_t k, float* x, float* y) const {
        int32_t j = col[k];
        *x = gg_global_coord(xj[(size_t)j * ld], off ? off[2 * (size_t)j] : 0.f, factor);
        *y = gg_global_coord(xj[(size_t)j * ld + 1], off ? off[2 * (size_t)j + 1] : 0.f, factor);
    }
};

extern "C" void region_center_host(const float* xj, int ld, const float* off, float factor, const int32_t* rowptr,
                                   const int32_t* col, const int32_t* key, int n_grain, double* centers, float* xg, int ld_g) {
    for (int g = 0; g < n_grain; ++g) {
        int beg = rowptr[g], end = rowptr[g + 1];
        double cx, cy;
        bool ok;
        if (key) {
            GGRegionWalk<FetchHost> w{key + beg, end - beg, FetchHost{xj, ld, factor > 1.f ? off : nullptr, factor, col + beg}, -1};
            ok = gg_region_center_one(w, &cx, &cy);
        } else {                                  // col already in dict order
            GGOrderedWalk<FetchHost> w{end - beg, FetchHost{xj, ld, factor > 1.f ? off : nullptr, factor, col + beg}, 0};
            ok = gg_region_center_one(w, &cx, &cy);
        }
        if (ok) {
            centers[2 * g] = cx; centers[2 * g + 1] = cy;
            xg[(size_t)g * ld_g] = gg_patch_coord((float)cx, factor);
            xg[(size_t)g * ld_g + 1] = gg_patch_coord((float)cy, factor);
        } else {
            centers[2 * g] = centers[2 * g + 1] = std::numeric_limits<double>::quiet_NaN();
        }
    }
}
