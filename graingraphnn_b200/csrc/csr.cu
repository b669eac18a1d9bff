// csr.cu — kernel family (a): stable dst-sorted CSR of one edge type.
//
// Replaces PyG's per-call COO gather/scatter bookkeeping (MessagePassing.propagate, called at
// periodGATconv.py:174) with an index structure built once per topology change.  Integer-exact:
// row i lists the in-edges of node i in original edge order (== numpy.argsort(dst, kind="stable")).
//
// Four passes, all HBM-bound integer work (no shared-memory staging needed: every access is a
// coalesced stream except the per-row counters, which live in L2):
//   1. histogram of targets (int atomics; counts are order-independent, hence deterministic)
//   2. exclusive scan -> rowptr   (two-level: 2048-element tiles, tile sums scanned by one block)
//   3. slot claim with an atomic cursor per row (order inside a row is arbitrary here ...)
//   4. ... and is made canonical by sorting each row's edge ids ascending (rows are 3-8 long).
#include "common.cuh"

namespace {

constexpr int kScanTile = 2048;  // elements per scan block (256 threads x 8)

__global__ void csr_histogram(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t E,
                              int n_src, int n_dst, int* __restrict__ count, int* __restrict__ status) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= E) return;
    int64_t d = dst[e], s = src[e];
    if (d < 0 || d >= n_dst || s < 0 || s >= n_src) { atomicExch(status, GG_ERANGE); return; }
    atomicAdd(&count[d], 1);
}

// tile-local exclusive scan; tile_sum[b] = total of tile b
__global__ void scan_tiles(const int* __restrict__ in, int* __restrict__ out, int n, int* __restrict__ tile_sum) {
    __shared__ int warp_tot[8];
    const int base = blockIdx.x * kScanTile + threadIdx.x * 8;
    int v[8], run = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] = (base + i < n) ? in[base + i] : 0; }
#pragma unroll
    for (int i = 0; i < 8; ++i) { int t = v[i]; v[i] = run; run += t; }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < warp; ++w) woff += warp_tot[w];
    const int excl = woff + incl - run;
#pragma unroll
    for (int i = 0; i < 8; ++i) if (base + i < n) out[base + i] = v[i] + excl;
    if (threadIdx.x == 255) tile_sum[blockIdx.x] = woff + incl;
}

// single block: exclusive scan of the tile sums in place (n_tiles is small: N / 2048)
__global__ void scan_tile_sums(int* __restrict__ tile_sum, int n_tiles) {
    __shared__ int carry;
    __shared__ int warp_tot[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n_tiles; base += blockDim.x) {
        int i = base + threadIdx.x;
        int x = (i < n_tiles) ? tile_sum[i] : 0;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        int incl = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        int woff = 0;
        for (int w = 0; w < warp; ++w) woff += warp_tot[w];
        if (i < n_tiles) tile_sum[i] = carry + woff + incl - x;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry += woff + incl;
        __syncthreads();
    }
}

__global__ void scan_add_offsets(int* __restrict__ out, int n, const int* __restrict__ tile_sum) {
    int i = blockIdx.x * kScanTile + threadIdx.x;
    const int off = tile_sum[blockIdx.x];
#pragma unroll
    for (int k = 0; k < 8; ++k, i += 256) if (i < n) out[i] += off;
}

__global__ void csr_claim(const int64_t* __restrict__ dst, int64_t E, int n_dst, const int* __restrict__ rowptr,
                          int* __restrict__ cursor, int* __restrict__ perm) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= E) return;
    int64_t d = dst[e];
    if (d < 0 || d >= n_dst) return;  // already reported by the histogram pass
    int slot = rowptr[d] + atomicAdd(&cursor[d], 1);
    perm[slot] = (int)e;
}

// canonical order inside each row: ascending original edge id; then col = src[perm]
__global__ void csr_sort_rows(const int64_t* __restrict__ src, int n_dst, const int* __restrict__ rowptr,
                              int* __restrict__ perm, int* __restrict__ col) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_dst) return;
    const int b = rowptr[r], e = rowptr[r + 1];
    for (int i = b + 1; i < e; ++i) {   // insertion sort: rows have 3-8 entries on grain graphs
        int key = perm[i], j = i - 1;
        while (j >= b && perm[j] > key) { perm[j + 1] = perm[j]; --j; }
        perm[j + 1] = key;
    }
    for (int i = b; i < e; ++i) col[i] = (int)src[perm[i]];
}

// work items of the gather kernel: node i contributes max(1, ceil(deg_i / dcap)) chunks of <= dcap in-edges
__global__ void items_count(const int* __restrict__ rowptr, int n_dst, int dcap, int* __restrict__ cnt) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n_dst) return;
    if (i == n_dst) { cnt[i] = 0; return; }
    const int deg = rowptr[i + 1] - rowptr[i];
    cnt[i] = deg == 0 ? 1 : (deg + dcap - 1) / dcap;
}
__global__ void items_fill(const int* __restrict__ rowptr, const int* __restrict__ item_ptr, int n_dst, int dcap, int4* __restrict__ items) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_dst) return;
    const int b = rowptr[i], e = rowptr[i + 1];
    int o = item_ptr[i];
    if (b == e) { items[o] = make_int4(i, b, 0x300, 0); return; }          // no in-edges: one empty item, first and last
    for (int pos = b; pos < e; pos += dcap, ++o) {
        const int c = min(dcap, e - pos);
        items[o] = make_int4(i, pos, c | (pos == b ? 0x100 : 0) | (pos + c >= e ? 0x200 : 0), 0);
    }
}

__global__ void permute_f32(const float* __restrict__ src, const int* __restrict__ perm, float* __restrict__ out, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = __ldg(&src[perm[i]]);
}

}  // namespace

extern "C" size_t gg_csr_workspace_bytes(int64_t n_edges, int32_t n_dst) {
    (void)n_edges;
    size_t n = (size_t)n_dst + 1;
    size_t tiles = (n + kScanTile - 1) / kScanTile;
    return (n + tiles + 64) * sizeof(int);   // counts/cursor [n_dst+1] + tile sums
}

extern "C" int gg_csr_build(const int64_t* edge_index, int64_t E, int32_t n_src, int32_t n_dst,
                            int32_t* rowptr, int32_t* col, int32_t* perm, int32_t* status,
                            void* workspace, size_t workspace_bytes, void* stream) {
    if (E < 0 || n_dst < 0 || n_src < 0 || !rowptr || !status || !workspace) return GG_EINVAL;
    if (E > 0 && (!edge_index || !col || !perm)) return GG_EINVAL;
    if (E > 0x7fffffffLL) return GG_EINVAL;
    if (workspace_bytes < gg_csr_workspace_bytes(E, n_dst)) return GG_ENOSPC;
    cudaStream_t st = GG_STREAM(stream);
    const int n = n_dst + 1;
    const int tiles = (n + kScanTile - 1) / kScanTile;
    int* count = static_cast<int*>(workspace);
    int* tile_sum = count + n;
    const int64_t* src = edge_index;
    const int64_t* dst = edge_index + E;
    cudaError_t err;
    if ((err = cudaMemsetAsync(count, 0, sizeof(int) * n, st)) != cudaSuccess) return (int)err;
    if ((err = cudaMemsetAsync(status, 0, sizeof(int), st)) != cudaSuccess) return (int)err;
    const int eb = (int)((E + 255) / 256);
    if (E > 0) { csr_histogram<<<eb, 256, 0, st>>>(src, dst, E, n_src, n_dst, count, status); GG_LAUNCH_OK(); }
    scan_tiles<<<tiles, 256, 0, st>>>(count, rowptr, n, tile_sum); GG_LAUNCH_OK();
    if (tiles > 1) {
        scan_tile_sums<<<1, 1024, 0, st>>>(tile_sum, tiles); GG_LAUNCH_OK();
        scan_add_offsets<<<tiles, 256, 0, st>>>(rowptr, n, tile_sum); GG_LAUNCH_OK();
    }
    if (E > 0) {
        if ((err = cudaMemsetAsync(count, 0, sizeof(int) * n, st)) != cudaSuccess) return (int)err;
        csr_claim<<<eb, 256, 0, st>>>(dst, E, n_dst, rowptr, count, perm); GG_LAUNCH_OK();
        if (n_dst > 0) { csr_sort_rows<<<(n_dst + 127) / 128, 128, 0, st>>>(src, n_dst, rowptr, perm, col); GG_LAUNCH_OK(); }
    }
    return 0;
}

extern "C" int gg_permute_f32(const float* src, const int32_t* perm, float* out, int64_t n, void* stream) {
    if (n < 0 || (n > 0 && (!src || !perm || !out))) return GG_EINVAL;
    if (n == 0) return 0;
    permute_f32<<<(unsigned)((n + 255) / 256), 256, 0, GG_STREAM(stream)>>>(src, perm, out, n);
    GG_LAUNCH_OK();
    return 0;
}

extern "C" int gg_csr_items(const int32_t* rowptr, int32_t n_dst, int32_t dcap, int32_t* item_ptr, int32_t* items,
                            void* workspace, size_t workspace_bytes, void* stream) {
    if (n_dst < 0 || dcap < 1 || dcap > 255 || !rowptr || !item_ptr || !workspace) return GG_EINVAL;
    if (n_dst > 0 && !items) return GG_EINVAL;
    if (workspace_bytes < gg_csr_workspace_bytes(0, n_dst)) return GG_ENOSPC;
    if (items && !gg_aligned16(items)) return GG_EALIGN;
    cudaStream_t st = GG_STREAM(stream);
    const int n = n_dst + 1;
    const int tiles = (n + kScanTile - 1) / kScanTile;
    int* count = static_cast<int*>(workspace);
    int* tile_sum = count + n;
    items_count<<<(n + 255) / 256, 256, 0, st>>>(rowptr, n_dst, dcap, count); GG_LAUNCH_OK();
    scan_tiles<<<tiles, 256, 0, st>>>(count, item_ptr, n, tile_sum); GG_LAUNCH_OK();
    if (tiles > 1) {
        scan_tile_sums<<<1, 1024, 0, st>>>(tile_sum, tiles); GG_LAUNCH_OK();
        scan_add_offsets<<<tiles, 256, 0, st>>>(item_ptr, n, tile_sum); GG_LAUNCH_OK();
    }
    if (n_dst > 0) { items_fill<<<(n_dst + 255) / 256, 256, 0, st>>>(rowptr, item_ptr, n_dst, dcap, reinterpret_cast<int4*>(items)); GG_LAUNCH_OK(); }
    return 0;
}
