// edge.cu — per-edge and per-node streaming kernels around the cells:
//   (a11) wrapped edge length, (d) node heads and the fused joint-pair edge-event head, (a12) feature update,
//   (e) halo row pack/unpack.  All are HBM-bound streams; none benefits from tensor cores.
#include "common.cuh"

namespace {

// test.py:562-575 — one thread per edge; x rows are only 8/11 floats so the two gathers hit L2-resident lines.
__global__ void edge_length_kernel(const float* __restrict__ xs, int lds, const float* __restrict__ xd, int ldd,
                                   const int64_t* __restrict__ ei, int64_t E, float* __restrict__ out) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int64_t s = ei[e], d = ei[E + e];
    float dx = __ldg(&xs[s * lds]) - __ldg(&xd[d * ldd]);
    float dy = __ldg(&xs[s * lds + 1]) - __ldg(&xd[d * ldd + 1]);
    dx = (float)((dx < -0.5f) - (dx > 0.5f)) + dx;
    dy = (float)((dy < -0.5f) - (dy > 0.5f)) + dy;
    out[e] = sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));   // no FMA contraction: matches torch
}

__global__ void permute_kernel(const float* __restrict__ src, const int* __restrict__ perm, float* __restrict__ out, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = src[perm[i]];
}

struct HeadW { float w[4][128]; float b[4]; int act[4]; };   // passed by value (kernel parameter space)

// models.py:433-452 — 8 lanes per node, each lane C/8 channels; y = act(W h + b)
template <int NV>
__global__ void node_head_kernel(const float* __restrict__ h, int ldh, const float* __restrict__ W,
                                 const float* __restrict__ b, int n_out, int act0, int act1, int act2, int act3,
                                 float* __restrict__ y, int ldy, const float* __restrict__ area_in, int ld_area,
                                 float area_scale, float* __restrict__ area_out, int M) {
    constexpr int C = 32 * NV;
    const int sub = threadIdx.x & 7;
    const int64_t m = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 3;
    const bool ok = m < M;
    float4 hv[NV];
#pragma unroll
    for (int r = 0; r < NV; ++r) hv[r] = ok ? ldg4(h + m * ldh + 4 * (sub + 8 * r)) : make_float4(0, 0, 0, 0);
    const int acts[4] = {act0, act1, act2, act3};
    for (int j = 0; j < n_out; ++j) {
        float acc = 0.f;
#pragma unroll
        for (int r = 0; r < NV; ++r) {
            float4 w = ldg4(W + j * C + 4 * (sub + 8 * r));
            acc = fmaf(hv[r].x, w.x, acc); acc = fmaf(hv[r].y, w.y, acc);
            acc = fmaf(hv[r].z, w.z, acc); acc = fmaf(hv[r].w, w.w, acc);
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        if (ok && sub == 0) {
            const float raw = acc + __ldg(&b[j]);
            float v = raw;
            if (acts[j] == 1) v = tanhf(raw); else if (acts[j] == 2) v = fmaxf(raw, 0.f);
            y[m * ldy + j] = v;
            if (j == 0 && area_out) area_out[m] = tanhf(raw) / area_scale + __ldg(&area_in[m * ld_area]);
        }
    }
}

// models.py:595-609 — 8 lanes per edge: gather h[src], h[dst] (C/8 channels per lane each), three dot products.
template <int NV>
__global__ void edge_head_kernel(const float* __restrict__ h, int ldh, const int64_t* __restrict__ ei, int64_t E,
                                 const float* __restrict__ ea, const float* __restrict__ W1, const float* __restrict__ b1,
                                 const float* __restrict__ W2, const float* __restrict__ b2,
                                 float* __restrict__ edge_event, float* __restrict__ edge) {
    constexpr int C = 32 * NV;
    constexpr int LDW = 2 * C + 1;
    const int sub = threadIdx.x & 7;
    const int64_t e = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 3;
    const bool ok = e < E;
    int64_t s = 0, d = 0;
    if (ok) { s = ei[e]; d = ei[E + e]; }
    float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int r = 0; r < NV; ++r) {
        const int c0 = 4 * (sub + 8 * r);
        float4 hs = ok ? ldg4(h + s * ldh + c0) : make_float4(0, 0, 0, 0);
        float4 hd = ok ? ldg4(h + d * ldh + c0) : make_float4(0, 0, 0, 0);
        const float hsv[4] = {hs.x, hs.y, hs.z, hs.w}, hdv[4] = {hd.x, hd.y, hd.z, hd.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {   // weight rows are 2C+1 long -> scalar (unaligned) loads, L1-resident
            acc[0] = fmaf(hsv[k], __ldg(&W2[c0 + k]), acc[0]);
            acc[1] = fmaf(hsv[k], __ldg(&W1[c0 + k]), acc[1]);
            acc[2] = fmaf(hsv[k], __ldg(&W1[LDW + c0 + k]), acc[2]);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            acc[0] = fmaf(hdv[k], __ldg(&W2[C + c0 + k]), acc[0]);
            acc[1] = fmaf(hdv[k], __ldg(&W1[C + c0 + k]), acc[1]);
            acc[2] = fmaf(hdv[k], __ldg(&W1[LDW + C + c0 + k]), acc[2]);
        }
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 1);
        acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 2);
        acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 4);
    }
    if (ok && sub == 0) {
        const float a = __ldg(&ea[e]);
        edge_event[e] = fmaf(a, __ldg(&W2[2 * C]), acc[0]) + __ldg(&b2[0]);
        if (edge) {
            edge[2 * e]     = tanhf(fmaf(a, __ldg(&W1[2 * C]), acc[1]) + __ldg(&b1[0]));
            edge[2 * e + 1] = tanhf(fmaf(a, __ldg(&W1[LDW + 2 * C]), acc[2]) + __ldg(&b1[1]));
        }
    }
}

// models.py:510-516 and test.py:401-402 (z += dz); the clamp of test.py:405-407 is the second kernel.
__global__ void feature_update_kernel(float* __restrict__ xj, int ldj, int nj, const float* __restrict__ yj,
                                      float* __restrict__ xg, int ldg, int ng, const float* __restrict__ yg, float dz, int last_g) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nj) {
        float* r = xj + (int64_t)i * ldj;
        const float dx = yj[2 * i], dy = yj[2 * i + 1];
        r[0] += dx / 5.0f; r[1] += dy / 5.0f;
        r[2] += dz;
        r[6] = dx; r[7] = dy;
    } else if (i < nj + ng) {
        i -= nj;
        float* r = xg + (int64_t)i * ldg;
        const float ds = yg[2 * i], dv = yg[2 * i + 1];
        r[3] += ds / 20.0f;
        r[4] = dv;
        r[last_g] = ds;
        r[2] += dz;
    }
}

// Ensemble form (block-diagonal batch of independent rollouts, each with its own span): per-node z increments; the clamp of
// test.py:405-407 is applied per node, which is the same thing as per graph because all nodes of a graph carry one z.
__global__ void feature_update_batched_kernel(float* __restrict__ xj, int ldj, int nj, const float* __restrict__ yj,
                                              float* __restrict__ xg, int ldg, int ng, const float* __restrict__ yg,
                                              const float* __restrict__ dzj, const float* __restrict__ dzg, float z_max, int last_g) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nj) {
        float* r = xj + (int64_t)i * ldj;
        const float dx = yj[2 * i], dy = yj[2 * i + 1];
        r[0] += dx / 5.0f; r[1] += dy / 5.0f;
        const float z = r[2] + dzj[i];
        r[2] = z > z_max ? z_max : z;
        r[6] = dx; r[7] = dy;
    } else if (i < nj + ng) {
        i -= nj;
        float* r = xg + (int64_t)i * ldg;
        const float ds = yg[2 * i], dv = yg[2 * i + 1];
        r[3] += ds / 20.0f;
        r[4] = dv;
        r[last_g] = ds;
        const float z = r[2] + dzg[i];
        r[2] = z > z_max ? z_max : z;
    }
}

__global__ void z_probe_kernel(const float* __restrict__ xg, float z_max, int* __restrict__ flag) {
    *flag = (xg[2] > z_max) ? 1 : 0;
}

__global__ void z_clamp_kernel(float* __restrict__ xj, int ldj, int nj, float* __restrict__ xg, int ldg, int ng,
                               float z_max, const int* __restrict__ flag) {
    if (*flag == 0) return;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nj) xj[(int64_t)i * ldj + 2] = z_max;
    else if (i < nj + ng) xg[(int64_t)(i - nj) * ldg + 2] = z_max;
}

// (e) halo pack: one float4 per thread, rows gathered by index
__global__ void gather_rows_kernel(const float* __restrict__ src, int lds, const int* __restrict__ idx, int n, int w4,
                                   float* __restrict__ out, int ldo) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= (int64_t)n * w4) return;
    const int r = (int)(t / w4), c = (int)(t % w4) * 4;
    const float4 v = ldg4(src + (int64_t)__ldg(&idx[r]) * lds + c);
    *reinterpret_cast<float4*>(out + (int64_t)r * ldo + c) = v;
}

__global__ void scatter_rows_kernel(const float* __restrict__ src, int lds, const int* __restrict__ idx, int n, int w4,
                                    float* __restrict__ out, int ldo) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= (int64_t)n * w4) return;
    const int r = (int)(t / w4), c = (int)(t % w4) * 4;
    const float4 v = ldg4(src + (int64_t)r * lds + c);
    *reinterpret_cast<float4*>(out + (int64_t)__ldg(&idx[r]) * ldo + c) = v;
}

}  // namespace

extern "C" int gg_edge_length(const float* x_src, int32_t ld_src, const float* x_dst, int32_t ld_dst,
                              const int64_t* edge_index, int64_t E, const int32_t* perm,
                              float* out, float* out_csr, void* stream) {
    if (E < 0 || ld_src < 2 || ld_dst < 2) return GG_EINVAL;
    if (E == 0) return 0;
    if (!x_src || !x_dst || !edge_index || !out || (out_csr && !perm)) return GG_EINVAL;
    const unsigned nb = (unsigned)((E + 255) / 256);
    edge_length_kernel<<<nb, 256, 0, GG_STREAM(stream)>>>(x_src, ld_src, x_dst, ld_dst, edge_index, E, out);
    GG_LAUNCH_OK();
    if (out_csr) { permute_kernel<<<nb, 256, 0, GG_STREAM(stream)>>>(out, perm, out_csr, E); GG_LAUNCH_OK(); }
    return 0;
}

extern "C" int gg_node_head(const float* h, int32_t ldh, int32_t C, const float* W, const float* b, int32_t n_out,
                            const int32_t* act_host, float* y, int32_t ldy,
                            const float* area_in, int32_t ld_area, float area_scale, float* area_out,
                            int32_t M, void* stream) {
    if (M < 0 || n_out < 1 || n_out > 4 || C % 32 || C < 32 || C > 128 || !act_host) return GG_EINVAL;
    if (M == 0) return 0;
    if (!h || !W || !b || !y || (area_out && !area_in)) return GG_EINVAL;
    if (!gg_aligned16(h) || (ldh & 3) || !gg_aligned16(W)) return GG_EALIGN;
    int a[4] = {0, 0, 0, 0};
    for (int j = 0; j < n_out; ++j) a[j] = act_host[j];
    const unsigned nb = (unsigned)(((int64_t)M * 8 + 255) / 256);
    cudaStream_t st = GG_STREAM(stream);
#define GG_NH(NV) node_head_kernel<NV><<<nb, 256, 0, st>>>(h, ldh, W, b, n_out, a[0], a[1], a[2], a[3], y, ldy, area_in, ld_area, area_scale, area_out, M)
    switch (C / 32) { case 1: GG_NH(1); break; case 2: GG_NH(2); break; case 3: GG_NH(3); break; default: GG_NH(4); }
#undef GG_NH
    GG_LAUNCH_OK();
    return 0;
}

extern "C" int gg_edge_head(const float* h, int32_t ldh, int32_t C, const int64_t* edge_index, int64_t E,
                            const float* eattr, const float* W1, const float* b1, const float* W2, const float* b2,
                            float* edge_event, float* edge, void* stream) {
    if (E < 0 || C % 32 || C < 32 || C > 128) return GG_EINVAL;
    if (E == 0) return 0;
    if (!h || !edge_index || !eattr || !W1 || !b1 || !W2 || !b2 || !edge_event) return GG_EINVAL;
    if (!gg_aligned16(h) || (ldh & 3)) return GG_EALIGN;
    const unsigned nb = (unsigned)((E * 8 + 255) / 256);
    cudaStream_t st = GG_STREAM(stream);
#define GG_EH(NV) edge_head_kernel<NV><<<nb, 256, 0, st>>>(h, ldh, edge_index, E, eattr, W1, b1, W2, b2, edge_event, edge)
    switch (C / 32) { case 1: GG_EH(1); break; case 2: GG_EH(2); break; case 3: GG_EH(3); break; default: GG_EH(4); }
#undef GG_EH
    GG_LAUNCH_OK();
    return 0;
}

extern "C" int gg_feature_update(float* x_joint, int32_t ld_j, int32_t n_joint, const float* y_joint,
                                 float* x_grain, int32_t ld_g, int32_t n_grain, int32_t n_grain_feat, const float* y_grain,
                                 float dz, float z_max, int32_t* scratch, void* stream) {
    if (n_joint < 0 || n_grain < 1 || ld_j < 8 || n_grain_feat < 6 || ld_g < n_grain_feat || !scratch) return GG_EINVAL;
    if (!x_joint || !x_grain || !y_joint || !y_grain) return GG_EINVAL;
    cudaStream_t st = GG_STREAM(stream);
    const int n = n_joint + n_grain;
    feature_update_kernel<<<(n + 255) / 256, 256, 0, st>>>(x_joint, ld_j, n_joint, y_joint, x_grain, ld_g, n_grain, y_grain, dz, n_grain_feat - 1);
    GG_LAUNCH_OK();
    z_probe_kernel<<<1, 1, 0, st>>>(x_grain, z_max, scratch);
    GG_LAUNCH_OK();
    z_clamp_kernel<<<(n + 255) / 256, 256, 0, st>>>(x_joint, ld_j, n_joint, x_grain, ld_g, n_grain, z_max, scratch);
    GG_LAUNCH_OK();
    return 0;
}

extern "C" int gg_feature_update_batched(float* x_joint, int32_t ld_j, int32_t n_joint, const float* y_joint,
                                         float* x_grain, int32_t ld_g, int32_t n_grain, int32_t n_grain_feat, const float* y_grain,
                                         const float* dz_joint, const float* dz_grain, float z_max, void* stream) {
    if (n_joint < 0 || n_grain < 1 || ld_j < 8 || n_grain_feat < 6 || ld_g < n_grain_feat) return GG_EINVAL;
    if (!x_joint || !x_grain || !y_joint || !y_grain || !dz_joint || !dz_grain) return GG_EINVAL;
    const int n = n_joint + n_grain;
    feature_update_batched_kernel<<<(n + 255) / 256, 256, 0, GG_STREAM(stream)>>>(x_joint, ld_j, n_joint, y_joint, x_grain, ld_g, n_grain,
                                                                                  y_grain, dz_joint, dz_grain, z_max, n_grain_feat - 1);
    GG_LAUNCH_OK();
    return 0;
}

static int rows_args_ok(const float* src, int32_t ld_src, const int32_t* idx, int32_t n, int32_t width,
                        float* out, int32_t ld_out) {
    if (n < 0 || width < 0 || (width & 3)) return GG_EINVAL;
    if (n == 0 || width == 0) return 1;
    if (!src || !idx || !out) return GG_EINVAL;
    if (!gg_aligned16(src) || !gg_aligned16(out) || (ld_src & 3) || (ld_out & 3)) return GG_EALIGN;
    return 0;
}

extern "C" int gg_gather_rows(const float* src, int32_t ld_src, const int32_t* idx, int32_t n, int32_t width,
                              float* out, int32_t ld_out, void* stream) {
    int rc = rows_args_ok(src, ld_src, idx, n, width, out, ld_out);
    if (rc) return rc < 0 ? rc : 0;
    const int64_t t = (int64_t)n * (width / 4);
    gather_rows_kernel<<<(unsigned)((t + 255) / 256), 256, 0, GG_STREAM(stream)>>>(src, ld_src, idx, n, width / 4, out, ld_out);
    GG_LAUNCH_OK();
    return 0;
}

extern "C" int gg_scatter_rows(const float* src, int32_t ld_src, const int32_t* idx, int32_t n, int32_t width,
                               float* out, int32_t ld_out, void* stream) {
    int rc = rows_args_ok(src, ld_src, idx, n, width, out, ld_out);
    if (rc) return rc < 0 ? rc : 0;
    const int64_t t = (int64_t)n * (width / 4);
    scatter_rows_kernel<<<(unsigned)((t + 255) / 256), 256, 0, GG_STREAM(stream)>>>(src, ld_src, idx, n, width / 4, out, ld_out);
    GG_LAUNCH_OK();
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// (a7) SAGEConv mean aggregation (heterogclstm.py:52-54 -> PyG SAGEConv(aggr='mean')): out[i, :] = mean over the
// in-edges of i of src[col[e], :], 0 for rows without in-edges.  One warp per target row, float4 columns, the row's
// edges summed in CSR (= original) order and divided once, like scatter(..., reduce='mean') = sum / max(count, 1).
// ---------------------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256)
segment_mean_kernel(const float* __restrict__ src, int ld_src, int width, const int* __restrict__ rowptr,
                    const int* __restrict__ col, int n_dst, float* __restrict__ out, int ld_out) {
    const int row = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (row >= n_dst) return;
    const int beg = __ldg(&rowptr[row]), end = __ldg(&rowptr[row + 1]);
    const float inv = end > beg ? 1.0f / (float)(end - beg) : 0.f;
    for (int c = 4 * lane; c < width; c += 128) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int e = beg; e < end; ++e) {
            const float4 v = ldg4(src + (size_t)__ldg(&col[e]) * ld_src + c);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        *reinterpret_cast<float4*>(out + (size_t)row * ld_out + c) = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
    }
}
}  // namespace

extern "C" int gg_segment_mean(const float* src, int32_t ld_src, int32_t width, const int32_t* rowptr, const int32_t* col,
                               int32_t n_dst, float* out, int32_t ld_out, void* stream) {
    if (n_dst < 0 || width < 0 || (width & 3)) return GG_EINVAL;
    if (n_dst == 0 || width == 0) return 0;
    if (!src || !rowptr || !col || !out) return GG_EINVAL;
    if (!gg_aligned16(src) || !gg_aligned16(out) || (ld_src & 3) || (ld_out & 3)) return GG_EALIGN;
    const unsigned nb = (unsigned)(((int64_t)n_dst * 32 + 255) / 256);
    segment_mean_kernel<<<nb, 256, 0, GG_STREAM(stream)>>>(src, ld_src, width, rowptr, col, n_dst, out, ld_out);
    GG_LAUNCH_OK();
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// per-edge periodic wrap of the source displacement (periodGATconv.py:209-210): r = p_j - p_i; +1 where r < -0.5,
// -1 where r > 0.5, per coordinate.  2 bits per coordinate (0: none, 1: +1, 2: -1), x | y << 2 | z << 4, CSR order.
// Computed once per step / forward and shared by every gate, cell and model that reads the same positions.
// ---------------------------------------------------------------------------------------------------------
namespace {
__device__ __forceinline__ int wrap_code2(float r) { return (r < -0.5f) ? 1 : ((r > 0.5f) ? 2 : 0); }
__global__ void edge_wrap_kernel(const float* __restrict__ ps, int ld_ps, const float* __restrict__ pd, int ld_pd,
                                 const int* __restrict__ rowptr, const int* __restrict__ col, int n_dst, int* __restrict__ wrap) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_dst) return;
    const float x = __ldg(pd + (size_t)i * ld_pd), y = __ldg(pd + (size_t)i * ld_pd + 1), z = __ldg(pd + (size_t)i * ld_pd + 2);
    const int b = __ldg(&rowptr[i]), e = __ldg(&rowptr[i + 1]);
    for (int k = b; k < e; ++k) {
        const float* pj = ps + (size_t)__ldg(&col[k]) * ld_ps;
        wrap[k] = wrap_code2(__ldg(pj) - x) | (wrap_code2(__ldg(pj + 1) - y) << 2) | (wrap_code2(__ldg(pj + 2) - z) << 4);
    }
}
// One pass per edge type and step over the CSR rows: wrap code (periodGATconv.py:209-210) and wrapped 2-D length
// (test.py:562-575, same arithmetic as edge_length_kernel: bit-identical) of every in-edge, written in CSR order and - through
// perm - in the reference's original edge order.  Replaces edge_wrap + edge_length + permute (3 launches) in the rollout step.
__global__ void edge_refresh_kernel(const float* __restrict__ ps, int ld_ps, const float* __restrict__ pd, int ld_pd,
                                    const int* __restrict__ rowptr, const int* __restrict__ col, const int* __restrict__ perm, int n_dst,
                                    int* __restrict__ wrap, float* __restrict__ ea_csr, float* __restrict__ ea_orig) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_dst) return;
    const float x = __ldg(pd + (size_t)i * ld_pd), y = __ldg(pd + (size_t)i * ld_pd + 1), z = __ldg(pd + (size_t)i * ld_pd + 2);
    const int b = __ldg(&rowptr[i]), e = __ldg(&rowptr[i + 1]);
    for (int k = b; k < e; ++k) {
        const float* pj = ps + (size_t)__ldg(&col[k]) * ld_ps;
        float dx = __ldg(pj) - x, dy = __ldg(pj + 1) - y;
        wrap[k] = wrap_code2(dx) | (wrap_code2(dy) << 2) | (wrap_code2(__ldg(pj + 2) - z) << 4);
        dx = (float)((dx < -0.5f) - (dx > 0.5f)) + dx;
        dy = (float)((dy < -0.5f) - (dy > 0.5f)) + dy;
        const float len = sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));   // no FMA contraction: matches torch
        ea_csr[k] = len;
        ea_orig[__ldg(&perm[k])] = len;
    }
}
}  // namespace

extern "C" int gg_edge_refresh(const float* pos_src, int32_t ld_pos_src, const float* pos_dst, int32_t ld_pos_dst,
                               const int32_t* rowptr, const int32_t* col, const int32_t* perm, int32_t n_dst,
                               int32_t* wrap_csr, float* eattr_csr, float* eattr, void* stream) {
    if (n_dst < 0 || ld_pos_src < 3 || ld_pos_dst < 3) return GG_EINVAL;
    if (n_dst == 0) return 0;
    if (!pos_src || !pos_dst || !rowptr || !col || !perm || !wrap_csr || !eattr_csr || !eattr) return GG_EINVAL;
    edge_refresh_kernel<<<(n_dst + 255) / 256, 256, 0, GG_STREAM(stream)>>>(pos_src, ld_pos_src, pos_dst, ld_pos_dst, rowptr, col, perm,
                                                                          n_dst, wrap_csr, eattr_csr, eattr);
    GG_LAUNCH_OK();
    return 0;
}

extern "C" int gg_edge_wrap(const float* pos_src, int32_t ld_pos_src, const float* pos_dst, int32_t ld_pos_dst,
                            const int32_t* rowptr, const int32_t* col, int32_t n_dst, int32_t* wrap_csr, void* stream) {
    if (n_dst < 0 || ld_pos_src < 3 || ld_pos_dst < 3) return GG_EINVAL;
    if (n_dst == 0) return 0;
    if (!pos_src || !pos_dst || !rowptr || !col || !wrap_csr) return GG_EINVAL;
    edge_wrap_kernel<<<(n_dst + 255) / 256, 256, 0, GG_STREAM(stream)>>>(pos_src, ld_pos_src, pos_dst, ld_pos_dst, rowptr, col, n_dst, wrap_csr);
    GG_LAUNCH_OK();
    return 0;
}
