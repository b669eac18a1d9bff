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
#ifndef GG_UNIT_TILES
#define GG_UNIT_TILES 4
#endif
constexpr int kUnitTiles = GG_UNIT_TILES;   // tiles per unit of the tiled gather (units go round-robin over the CTAs)

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

// ---- tile index of the warp-specialised gather (gather_tiled.cu) ------------------------------------------------
// nz_flag / scan / nz_fill: the targets WITH in-edges, ascending (nz), and where their rows start (nzptr, nzptr[NZ] = E)
__global__ void nz_flag(const int* __restrict__ rowptr, int n_dst, int* __restrict__ flag) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n_dst) return;
    flag[i] = (i < n_dst && rowptr[i + 1] > rowptr[i]) ? 1 : 0;
}
__global__ void nz_fill(const int* __restrict__ rowptr, const int* __restrict__ pos, int n_dst, int* __restrict__ nz,
                        int* __restrict__ nzptr, int* __restrict__ nz_count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n_dst) return;
    const int o = pos[i];
    if (i == n_dst) { nzptr[o] = rowptr[n_dst]; nz_count[0] = o; return; }
    const int b = rowptr[i];
    if (rowptr[i + 1] > b) { nz[o] = i; nzptr[o] = b; }
}
// Units and tiles (see gather_tiled.cu).  Unit k = targets (indices into nz) whose first in-edge lies in [k B, (k+1) B); its
// edges [nzptr[i_lo], nzptr[i_hi]) are cut into tiles of <= ecap edges.  Units are laid out CTA-major: slot b * Uc + k' holds
// unit k = b + k' * n_ctas, so an exclusive scan over the slots orders the tiles by CTA, then by unit, then by position.
__device__ __forceinline__ int nz_lower_bound(const int* __restrict__ nzptr, int nzc, int x) {   // first i in [0, nzc] with nzptr[i] >= x
    int lo = 0, hi = nzc;                            // nzptr[nzc] = E >= every x asked for
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(&nzptr[mid]) >= x) hi = mid; else lo = mid + 1;
    }
    return lo;
}
__device__ __forceinline__ int nz_owner(const int* __restrict__ nzptr, int nzc, int e) {         // largest i in [0, nzc) with nzptr[i] <= e
    int lo = 0, hi = nzc - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(&nzptr[mid]) <= e) lo = mid; else hi = mid - 1;
    }
    return lo;
}
__global__ void unit_tiles_count(const int* __restrict__ nzptr, const int* __restrict__ nz_count, int E, int B, int ecap,
                                 int n_units, int n_ctas, int Uc, int* __restrict__ cnt) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot > n_ctas * Uc) return;
    int nt = 0;
    const int k = Uc > 0 ? slot / Uc + (slot % Uc) * n_ctas : n_units;
    if (slot < n_ctas * Uc && k < n_units) {
        const int nzc = nz_count[0];
        const int i_lo = nz_lower_bound(nzptr, nzc, k * B), i_hi = nz_lower_bound(nzptr, nzc, min((k + 1) * B, E));
        if (i_lo < i_hi) nt = (__ldg(&nzptr[i_hi]) - __ldg(&nzptr[i_lo]) + ecap - 1) / ecap;
    }
    cnt[slot] = nt;
}
__global__ void unit_tiles_fill(const int* __restrict__ nzptr, const int* __restrict__ nz_count, int E, int B, int ecap,
                                int n_units, int n_ctas, int Uc, const int* __restrict__ pos, int4* __restrict__ tiles,
                                int* __restrict__ cta_ptr) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot > n_ctas * Uc) return;
    if (Uc == 0) { for (int b = 0; b <= n_ctas; ++b) cta_ptr[b] = 0; return; }     // no edges: slot 0 is the only thread
    if (slot == n_ctas * Uc) { cta_ptr[n_ctas] = pos[slot]; return; }
    if (slot % Uc == 0) cta_ptr[slot / Uc] = pos[slot];
    const int k = slot / Uc + (slot % Uc) * n_ctas;
    if (k >= n_units) return;
    const int nzc = nz_count[0];
    const int i_lo = nz_lower_bound(nzptr, nzc, k * B), i_hi = nz_lower_bound(nzptr, nzc, min((k + 1) * B, E));
    if (i_lo >= i_hi) return;
    const int lo = __ldg(&nzptr[i_lo]), hi = __ldg(&nzptr[i_hi]);
    int o = pos[slot];
    for (int e0 = lo; e0 < hi; e0 += ecap, ++o) {
        const int ne = min(ecap, hi - e0);
        tiles[o] = make_int4(e0, ne, nz_owner(nzptr, nzc, e0), nz_owner(nzptr, nzc, e0 + ne - 1));
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

extern "C" int gg_csr_compact(const int32_t* rowptr, int32_t n_dst, int32_t* nz, int32_t* nzptr, int32_t* nz_count,
                              int32_t* scratch, void* workspace, size_t workspace_bytes, void* stream) {
    if (n_dst < 0 || !rowptr || !nzptr || !nz_count || !scratch || !workspace) return GG_EINVAL;
    if (n_dst > 0 && !nz) return GG_EINVAL;
    if (workspace_bytes < gg_csr_workspace_bytes(0, n_dst)) return GG_ENOSPC;
    cudaStream_t st = GG_STREAM(stream);
    const int n = n_dst + 1;
    const int tiles = (n + kScanTile - 1) / kScanTile;
    int* flag = static_cast<int*>(workspace);
    int* tile_sum = flag + n;
    nz_flag<<<(n + 255) / 256, 256, 0, st>>>(rowptr, n_dst, flag); GG_LAUNCH_OK();
    scan_tiles<<<tiles, 256, 0, st>>>(flag, scratch, n, tile_sum); GG_LAUNCH_OK();
    if (tiles > 1) {
        scan_tile_sums<<<1, 1024, 0, st>>>(tile_sum, tiles); GG_LAUNCH_OK();
        scan_add_offsets<<<tiles, 256, 0, st>>>(scratch, n, tile_sum); GG_LAUNCH_OK();
    }
    nz_fill<<<(n + 255) / 256, 256, 0, st>>>(rowptr, scratch, n_dst, nz, nzptr, nz_count); GG_LAUNCH_OK();
    return 0;
}

extern "C" int64_t gg_csr_tiles_capacity(int64_t n_edges, int32_t ecap, int32_t n_ctas) {
    if (n_edges < 0 || ecap < 1 || n_ctas < 1) return 0;
    const int64_t B = (int64_t)kUnitTiles * ecap, n_units = (n_edges + B - 1) / B;
    return n_units + n_edges / ecap + 1;              // every unit ends in at most one partial tile
}

extern "C" size_t gg_csr_tiles_scratch_ints(int64_t n_edges, int32_t ecap, int32_t n_ctas) {
    if (n_edges < 0 || ecap < 1 || n_ctas < 1) return 0;
    const int64_t B = (int64_t)kUnitTiles * ecap, n_units = (n_edges + B - 1) / B, Uc = (n_units + n_ctas - 1) / n_ctas;
    const int64_t n = (int64_t)n_ctas * Uc + 1;
    return (size_t)(2 * n + (n + kScanTile - 1) / kScanTile + 64);
}

extern "C" int gg_csr_tiles(const int32_t* nzptr, const int32_t* nz_count, int64_t n_edges, int32_t ecap, int32_t n_ctas,
                            int32_t* tiles, int32_t* cta_ptr, int32_t* scratch, void* stream) {
    if (n_edges < 0 || n_edges > 0x7fffffffLL || ecap < 1 || n_ctas < 1 || !nzptr || !nz_count || !cta_ptr || !scratch) return GG_EINVAL;
    if (n_edges > 0 && (!tiles || !gg_aligned16(tiles))) return GG_EINVAL;
    cudaStream_t st = GG_STREAM(stream);
    const int B = kUnitTiles * ecap;
    const int n_units = (int)((n_edges + B - 1) / B);
    const int Uc = (n_units + n_ctas - 1) / n_ctas;
    const int n = n_ctas * Uc + 1;
    const int scan_blocks = (n + kScanTile - 1) / kScanTile;
    int* cnt = scratch;
    int* pos = cnt + n;
    int* tile_sum = pos + n;
    unit_tiles_count<<<(n + 127) / 128, 128, 0, st>>>(nzptr, nz_count, (int)n_edges, B, ecap, n_units, n_ctas, Uc, cnt); GG_LAUNCH_OK();
    scan_tiles<<<scan_blocks, 256, 0, st>>>(cnt, pos, n, tile_sum); GG_LAUNCH_OK();
    if (scan_blocks > 1) {
        scan_tile_sums<<<1, 1024, 0, st>>>(tile_sum, scan_blocks); GG_LAUNCH_OK();
        scan_add_offsets<<<scan_blocks, 256, 0, st>>>(pos, n, tile_sum); GG_LAUNCH_OK();
    }
    unit_tiles_fill<<<(n + 127) / 128, 128, 0, st>>>(nzptr, nz_count, (int)n_edges, B, ecap, n_units, n_ctas, Uc, pos,
                                                     reinterpret_cast<int4*>(tiles), cta_ptr); GG_LAUNCH_OK();
    return 0;
}
