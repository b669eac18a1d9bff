// raster.cu — row f4: the polygon raster of the layer-error QoI (graph_datastruct.py:553-610 `plot_polygons`, periodic branch, and
// :346-348 `compute_error_layer`) on the device.
//
// The reference draws every grain polygon with PIL's ImageDraw.polygon(fill = grain id) into a (2s x 2s) image, in `region_coors`
// dict order (later polygons overwrite earlier ones where they touch), then folds the four s x s quadrants with max().  Here:
//   gg_raster_polygons: one warp per polygon, lanes take scan lines; a pixel keeps the LARGEST draw index that covers it
//                       (atomicMax — the same winner as sequential overdraw, in any execution order);
//   gg_raster_fold:     alpha[y][x] = max over the four quadrants of the grain id of the winning polygon (0: never drawn).
// The scan-line rule restates PIL's polygon fill for integer vertices: per row the float32 intersections
// (y - y0) * dx + x0 (separate multiply and add, as the reference's x86-64 build rounds them), an edge that ENDS on the row
// counts twice unless the row is the polygon's last, horizontal edges are drawn directly, spans run from round-half-up of the
// left to round-half-down of the right intersection, and a single-vertex first / last row whose two edges both lean to larger x
// in the adjacent row is connected to that row's span.  Checked pixel for pixel against PIL on the reference's tilings
// (tests/test_raster.py); on random convex polygons 1.3 % differ from PIL by a few pixels of one row (an order-dependent corner
// rule of PIL's that convex tilings do not trigger once the neighbours are drawn).
#include "common.cuh"

namespace {

constexpr int kMaxEdges = 32;        // vertices per polygon handled by the warp kernel (grains have 3..12)

__device__ __forceinline__ int round_up_(float f) { return f >= 0.0f ? (int)floorf(f + 0.5f) : -(int)floorf(fabsf(f) + 0.5f); }
__device__ __forceinline__ int round_down_(float f) { return f >= 0.0f ? (int)ceilf(f - 0.5f) : -(int)ceilf(fabsf(f) - 0.5f); }

__device__ __forceinline__ void hline_max(int* img, int W, int H, int y, int xa, int xb, int ink) {
    if (y < 0 || y >= H) return;
    xa = max(xa, 0); xb = min(xb, W - 1);
    for (int x = xa; x <= xb; ++x) atomicMax(&img[(size_t)y * W + x], ink);
}

__global__ void raster_polygons_kernel(const int32_t* __restrict__ poly_ptr, const int32_t* __restrict__ verts, int n_poly,
                                       int* __restrict__ img, int W, int H) {
    const int p = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (p >= n_poly) return;
    const int v0 = poly_ptr[p], n = poly_ptr[p + 1] - v0;
    if (n <= 1 || n > kMaxEdges) return;                               // graph_datastruct.py:588 `if len(p)>1`
    int ex0[kMaxEdges], ey0[kMaxEdges], eya[kMaxEdges], eyb[kMaxEdges];
    float edx[kMaxEdges];
    int ne = 0, ymin_p = INT_MAX, ymax_p = INT_MIN;
    for (int i = 0; i < n; ++i) {
        const int x0 = verts[2 * (v0 + i)], y0 = verts[2 * (v0 + i) + 1];
        const int i1 = i + 1 < n ? i + 1 : 0;
        const int x1 = verts[2 * (v0 + i1)], y1 = verts[2 * (v0 + i1) + 1];
        ymin_p = min(ymin_p, y0); ymax_p = max(ymax_p, y0);
        if (y0 == y1) {                                                // horizontal edge: drawn directly
            if (lane == 0) hline_max(img, W, H, y0, min(x0, x1), max(x0, x1), p);
            continue;
        }
        ex0[ne] = x0; ey0[ne] = y0; eya[ne] = min(y0, y1); eyb[ne] = max(y0, y1);
        edx[ne] = __fdiv_rn((float)(x1 - x0), (float)(y1 - y0));
        ++ne;
    }
    const int y_lo = max(ymin_p, 0), y_hi = min(ymax_p, H);
    for (int y = y_lo + lane; y <= y_hi; y += 32) {
        float xx[2 * kMaxEdges];
        int j = 0;
        for (int i = 0; i < ne; ++i) {
            if (y >= eya[i] && y <= eyb[i]) {
                const float xv = __fadd_rn(__fmul_rn((float)(y - ey0[i]), edx[i]), (float)ex0[i]);
                xx[j++] = xv;
                if (y == eyb[i] && y < ymax_p) xx[j++] = xv;          // an edge ending here counts twice ("consistent polygons")
            }
        }
        if (j == 2 && xx[0] == xx[1] && (y == ymax_p || y == ymin_p)) {
            // sheared extreme vertex: both edges lean to larger x in the adjacent row -> connect to that row's span
            const int off = y == ymax_p ? -1 : 1;
            float lo = INFINITY;
            bool right = true;
            for (int i = 0; i < ne; ++i)
                if (y >= eya[i] && y <= eyb[i]) {
                    const float a = __fadd_rn(__fmul_rn((float)(y + off - ey0[i]), edx[i]), (float)ex0[i]);
                    right = right && a > xx[0];
                    lo = fminf(lo, a);
                }
            if (right) xx[1] = fmaxf(xx[0], (float)(round_up_(lo) - 1));
        }
        for (int a = 1; a < j; ++a) {                                  // insertion sort (<= a dozen entries)
            const float v = xx[a];
            int b = a - 1;
            while (b >= 0 && xx[b] > v) { xx[b + 1] = xx[b]; --b; }
            xx[b + 1] = v;
        }
        int x_pos = j == 0 ? -1 : 0;
        for (int i = 1; i < j; i += 2) {
            const int x_end = round_down_(xx[i]);
            if (x_end < x_pos) continue;
            int x_start = round_up_(xx[i - 1]);
            if (x_pos > x_start) {
                x_start = x_pos;
                if (x_end < x_start) continue;
            }
            hline_max(img, W, H, y, x_start, x_end, p);
            x_pos = x_end + 1;
        }
    }
}

__global__ void raster_fold_kernel(const int* __restrict__ img, const int32_t* __restrict__ ids, int s, int32_t* __restrict__ alpha) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s * s) return;
    const int y = i / s, x = i - y * s, W = 2 * s;
    int best = 0;
    const int q[4] = {img[(size_t)y * W + x], img[(size_t)(y + s) * W + x], img[(size_t)y * W + x + s], img[(size_t)(y + s) * W + x + s]};
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (q[k] >= 0) best = max(best, ids[q[k]]);
    alpha[i] = best;
}

__global__ void count_mismatch_kernel(const int32_t* __restrict__ a, const int32_t* __restrict__ b, int64_t n, unsigned long long* __restrict__ out) {
    unsigned long long c = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) c += a[i] != b[i];
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

}  // namespace

extern "C" int gg_raster_polygons(const int32_t* poly_ptr, const int32_t* verts, const int32_t* ids, int32_t n_poly, int32_t s,
                                  int32_t* scratch /* [2s x 2s] */, int32_t* alpha /* [s x s] */, void* stream) {
    if (n_poly < 0 || s < 1 || !scratch || !alpha || (n_poly > 0 && (!poly_ptr || !verts || !ids))) return GG_EINVAL;
    cudaStream_t st = GG_STREAM(stream);
    const int W = 2 * s;
    cudaError_t err = cudaMemsetAsync(scratch, 0xff, (size_t)W * W * sizeof(int32_t), st);       // -1: never drawn
    if (err != cudaSuccess) return (int)err;
    if (n_poly > 0) {
        const int warps_per_block = 4;
        raster_polygons_kernel<<<(n_poly + warps_per_block - 1) / warps_per_block, 32 * warps_per_block, 0, st>>>(poly_ptr, verts, n_poly, scratch, W, W);
        GG_LAUNCH_OK();
    }
    raster_fold_kernel<<<(s * s + 255) / 256, 256, 0, st>>>(scratch, ids, s, alpha);
    GG_LAUNCH_OK();
    return 0;
}

extern "C" int gg_count_mismatch(const int32_t* a, const int32_t* b, int64_t n, unsigned long long* count, void* stream) {
    if (n < 0 || !count || (n > 0 && (!a || !b))) return GG_EINVAL;
    cudaStream_t st = GG_STREAM(stream);
    cudaError_t err = cudaMemsetAsync(count, 0, sizeof(unsigned long long), st);
    if (err != cudaSuccess) return (int)err;
    if (n > 0) {
        count_mismatch_kernel<<<148 * 4, 256, 0, st>>>(a, b, n, count);
        GG_LAUNCH_OK();
    }
    return 0;
}
