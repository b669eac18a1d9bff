// geometry.cu — geometry feedback of a rollout step: grain centres from the joint positions.
//
// Replaces, for the periodic boundary, the per-grain Python loop of graph.update (graph_datastruct.py:672-708) that
// graph_trajectory.GNN_update (graph_trajectory.py:1010-1098) runs on the host after every NN step, and the per-grain
// tensor assignment of test.py:556-559 that writes the centres back into the grain features.  The reference keeps this
// on the CPU (X.detach().numpy()), which forces a device->host->device round trip per step; here the joint rows never
// leave HBM.  HBM-bound integer / fp64 work: ~6 joints per grain, one thread per grain, rows of 8 B gathered through the
// grain->joint CSR (L2-resident: 16 B per joint at 2 joints per grain).
#include "common.cuh"
#include "geometry_core.h"

namespace {

__global__ void fill_i32_kernel(int32_t* p, int32_t v, int32_t n) {
    int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// rank[j] = position of joint j's first appearance as a target of the grain->joint edges (its place in the reference's
// joint2vertex dict, graph_trajectory.py:1062-1080).
__global__ void joint_rank_kernel(const int64_t* __restrict__ gj_dst, int64_t E, int32_t n_joint, int32_t* rank) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    int64_t j = gj_dst[e];
    if (j >= 0 && j < n_joint) atomicMin(&rank[j], (int32_t)e);
}

// key[k] = rank[col[k]] in CSR order, so the per-grain walk scans contiguous ints.
__global__ void edge_key_kernel(const int32_t* __restrict__ col, const int32_t* __restrict__ rank, int64_t E, int32_t* key) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < E) key[e] = rank[col[e]];
}

struct FetchJoint {
    const float* xj; int32_t ld; const float* off; float factor; const int32_t* col;
    __device__ __forceinline__ void operator()(int32_t k, float* x, float* y) const {
        int32_t j = col[k];
        float2 p = *reinterpret_cast<const float2*>(xj + (size_t)j * ld);
        float ox = 0.f, oy = 0.f;
        if (off) { float2 o = *reinterpret_cast<const float2*>(off + 2 * (size_t)j); ox = o.x; oy = o.y; }
        *x = gg_global_coord(p.x, ox, factor);
        *y = gg_global_coord(p.y, oy, factor);
    }
};

// Once per topology: the joints of every grain in dict order (selection by repeated minimum over the keys, ~6 per grain).
__global__ void region_sort_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                   const int32_t* __restrict__ key, int32_t n_grain, int32_t* __restrict__ col_sorted) {
    int32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_grain) return;
    int32_t beg = rowptr[g], end = rowptr[g + 1], last = -1;
    for (int32_t i = beg; i < end; ++i) {
        int32_t best = 0x7fffffff, at = beg;
        for (int32_t k = beg; k < end; ++k) {
            int32_t v = key[k];
            if (v > last && v < best) { best = v; at = k; }
        }
        last = best;
        col_sorted[i] = col[at];
    }
}

template <bool ORDERED>
__global__ void __launch_bounds__(128)
region_center_kernel(const float* __restrict__ xj, int32_t ld_j, const float* __restrict__ off, float factor,
                     const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const int32_t* __restrict__ key,
                     int32_t n_grain, double* __restrict__ centers, float* __restrict__ x_grain, int32_t ld_g) {
    int32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_grain) return;
    int32_t beg = rowptr[g], end = rowptr[g + 1];
    double cx, cy;
    bool ok;
    if (ORDERED) {
        GGOrderedWalk<FetchJoint> w{end - beg, FetchJoint{xj, ld_j, off, factor, col + beg}, 0};
        ok = gg_region_center_one(w, &cx, &cy);
    } else {
        GGRegionWalk<FetchJoint> w{key + beg, end - beg, FetchJoint{xj, ld_j, off, factor, col + beg}, -1};
        ok = gg_region_center_one(w, &cx, &cy);
    }
    if (ok) {
        if (centers) { centers[2 * (size_t)g] = cx; centers[2 * (size_t)g + 1] = cy; }
        if (x_grain) {
            x_grain[(size_t)g * ld_g]     = gg_patch_coord((float)cx, factor);
            x_grain[(size_t)g * ld_g + 1] = gg_patch_coord((float)cy, factor);
        }
    } else if (centers) {
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        centers[2 * (size_t)g] = nan; centers[2 * (size_t)g + 1] = nan;
    }
}

}  // namespace

extern "C" int gg_joint_rank(const int64_t* gj_dst, int64_t n_edges, int32_t n_joint, int32_t* rank, void* stream) {
    if (n_edges < 0 || n_joint < 0 || (n_edges > 0 && !gj_dst) || (n_joint > 0 && !rank)) return GG_EINVAL;
    if (n_edges > 0x7fffffffLL) return GG_EINVAL;
    if (n_edges == 0 && n_joint == 0) return 0;
    cudaStream_t s = GG_STREAM(stream);
    if (n_joint > 0) fill_i32_kernel<<<(n_joint + 255) / 256, 256, 0, s>>>(rank, 0x7fffffff, n_joint);
    if (n_edges > 0) joint_rank_kernel<<<(unsigned)((n_edges + 255) / 256), 256, 0, s>>>(gj_dst, n_edges, n_joint, rank);
    GG_LAUNCH_OK();
    return 0;
}

extern "C" int gg_region_key(const int32_t* col, const int32_t* rank, int64_t n_edges, int32_t* key, void* stream) {
    if (n_edges < 0 || (n_edges > 0 && (!col || !rank || !key))) return GG_EINVAL;
    if (n_edges == 0) return 0;
    if (n_edges > 0) edge_key_kernel<<<(unsigned)((n_edges + 255) / 256), 256, 0, GG_STREAM(stream)>>>(col, rank, n_edges, key);
    GG_LAUNCH_OK();
    return 0;
}

extern "C" int gg_region_center(const float* x_joint, int32_t ld_j, const float* joint_offset, float domain_factor,
                                const int32_t* rowptr, const int32_t* col, const int32_t* key, int32_t n_grain,
                                double* centers, float* x_grain, int32_t ld_g, void* stream) {
    if (n_grain < 0 || ld_j < 2 || (ld_j & 1) || !x_joint || !rowptr || (x_grain && ld_g < 2)) return GG_EINVAL;
    if ((reinterpret_cast<uintptr_t>(x_joint) & 7u) || (joint_offset && (reinterpret_cast<uintptr_t>(joint_offset) & 7u)))
        return GG_EALIGN;
    if (domain_factor > 1.0f && !joint_offset) return GG_EINVAL;
    if (n_grain == 0) return 0;
    const float* off = domain_factor > 1.0f ? joint_offset : nullptr;
    if (key)
        region_center_kernel<false><<<(n_grain + 127) / 128, 128, 0, GG_STREAM(stream)>>>(
            x_joint, ld_j, off, domain_factor, rowptr, col, key, n_grain, centers, x_grain, ld_g);
    else
        region_center_kernel<true><<<(n_grain + 127) / 128, 128, 0, GG_STREAM(stream)>>>(
            x_joint, ld_j, off, domain_factor, rowptr, col, nullptr, n_grain, centers, x_grain, ld_g);
    GG_LAUNCH_OK();
    return 0;
}

extern "C" int gg_region_sort(const int32_t* rowptr, const int32_t* col, const int32_t* key, int32_t n_grain,
                              int32_t* col_sorted, void* stream) {
    if (n_grain < 0 || (n_grain > 0 && (!rowptr || !col_sorted))) return GG_EINVAL;
    if (n_grain == 0) return 0;
    region_sort_kernel<<<(n_grain + 127) / 128, 128, 0, GG_STREAM(stream)>>>(rowptr, col, key, n_grain, col_sorted);
    GG_LAUNCH_OK();
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// QoI bookkeeping of GNN_update (graph_trajectory.py:1041-1051, :1100-1103), on the device next to the features it reads.
//   area_counts[g] = area_g * s^2 / area_sum,  area_sum = sum(area * mask) / (lxd / patch)^2        (live grains, else NaN)
//   extraV[g]      = mask_g * extraV_g / 20 * s^3
//   vertex_area[j] = mesh^2 * sum over the grains g of joint j of area_counts[g] / (number of joints of g)
// area_sum comes from the caller (a float64 reduction over the grain rows); both kernels are one thread per row.
namespace {
__global__ void area_counts_kernel(const float* __restrict__ xg, int ld_g, const float* __restrict__ mask, int ld_m, int n_grain,
                                   double s, double area_sum, double v_scale, double* __restrict__ area_counts, double* __restrict__ extra_v) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_grain) return;
    const double m = mask ? (double)mask[(size_t)g * ld_m] : 1.0;
    const double area = (double)xg[(size_t)g * ld_g + 3], ev = (double)xg[(size_t)g * ld_g + 4];
    area_counts[g] = m > 0.0 ? area * (s * s) / area_sum : nan("");
    extra_v[g] = m * ev / v_scale * (s * s * s);
}
__global__ void vertex_area_kernel(const int32_t* __restrict__ rowptr_j, const int32_t* __restrict__ col_g,
                                   const int32_t* __restrict__ rowptr_g, const double* __restrict__ area_counts, double mesh2,
                                   int n_joint, double* __restrict__ vertex_area) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_joint) return;
    double acc = 0.0;
    for (int e = rowptr_j[j]; e < rowptr_j[j + 1]; ++e) {
        const int g = col_g[e];
        const int deg = rowptr_g[g + 1] - rowptr_g[g];
        const double a = area_counts[g];
        if (deg > 0 && a == a) acc += a / (double)deg * mesh2;
    }
    vertex_area[j] = acc;
}
}  // namespace

extern "C" int gg_area_bookkeeping(const float* x_grain, int32_t ld_g, const float* mask_grain, int32_t ld_m, int32_t n_grain,
                                   double s, double area_sum, double v_scale, double* area_counts, double* extra_v,
                                   const int32_t* rowptr_j, const int32_t* col_g, const int32_t* rowptr_g, int32_t n_joint,
                                   double mesh2, double* vertex_area, void* stream) {
    if (n_grain < 0 || n_joint < 0 || ld_g < 5 || !x_grain || !area_counts || !extra_v || !(area_sum == area_sum)) return GG_EINVAL;
    if (n_grain > 0) {
        area_counts_kernel<<<(n_grain + 255) / 256, 256, 0, GG_STREAM(stream)>>>(x_grain, ld_g, mask_grain, ld_m, n_grain, s, area_sum, v_scale,
                                                                                area_counts, extra_v);
        GG_LAUNCH_OK();
    }
    if (vertex_area && n_joint > 0) {
        if (!rowptr_j || !col_g || !rowptr_g) return GG_EINVAL;
        vertex_area_kernel<<<(n_joint + 255) / 256, 256, 0, GG_STREAM(stream)>>>(rowptr_j, col_g, rowptr_g, area_counts, mesh2, n_joint, vertex_area);
        GG_LAUNCH_OK();
    }
    return 0;
}
