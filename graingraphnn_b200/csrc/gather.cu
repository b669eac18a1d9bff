// gather.cu — kernel family (b): fused periodic-attention gather / segment-softmax / aggregate.
//
// Takes over PeriodConv.message + PyG softmax + scatter-add (periodGATconv.py:204-236, :174) for ONE edge type
// and ALL gates of a cell, operating on per-node projections (see gg_node_proj) instead of per-edge GEMMs.
//
// Mapping: one warp per (target node, gate quad).  Inside the warp, 8 lanes own one gate: lane `sub` holds the
// float4 chunks sub, sub+8, ... of that gate's C channels, so each 8-lane group reads 128 contiguous bytes per
// load instruction (LDG.E.128) and a 4-gate row (G*C*4 B = 1.5 KB at C=96) is fetched by 3 fully-coalesced
// requests.  Scores need a reduction over 8 lanes only (3 shuffles).  The edges of a row are visited in CSR order
// (= original edge order, stable sort), so the per-row sum has the same association as the CPU index_add_.
// No atomics, no scratch in global memory; per-warp scores live in shared memory (up to 32 in-edges; longer rows
// take a recompute path).
//
// HBM-bound: algorithmic bytes per launch = 4*G*C*(2*N_src + 2*N_dst) + 12*(N_src+N_dst) + 4*(N_dst+1) + 8*E
// (read K,V once per source, Q once and write agg once per target; positions; rowptr; col + edge attr).
#include "common.cuh"
#include <math_constants.h>

namespace {

constexpr int kWarpsPerBlock = 8;
constexpr int kMaxDegSmem = 32;

struct GatherParams {
    const float* P_src; int ld_src, k_off, v_off;
    const float* P_dst; int ld_dst, q_off, qx_off;
    const float* pos_src; int ld_ps;
    const float* pos_dst; int ld_pd;
    const int* rowptr; const int* col; const float* ea;
    const float* Wv3;
    int n_dst, G, quads, weighted;
    float* agg; float* agg_lo; int ld_agg; float* ea_out;
    float sqrt_c;
};

template <int NV>
__device__ __forceinline__ float dot_group(const float4 (&a)[NV], const float4 (&b)[NV]) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < NV; ++r) {
        s = fmaf(a[r].x, b[r].x, s); s = fmaf(a[r].y, b[r].y, s);
        s = fmaf(a[r].z, b[r].z, s); s = fmaf(a[r].w, b[r].w, s);
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    return s;
}

__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

__device__ __forceinline__ float wrapf(float r) { return (float)((r < -0.5f) - (r > 0.5f)); }

template <int NV>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
pgat_gather_kernel(const GatherParams p) {
    constexpr int C = 32 * NV;
    __shared__ float s_score[kWarpsPerBlock][4][kMaxDegSmem];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane >> 3, sub = lane & 7;
    const int64_t unit = (int64_t)blockIdx.x * kWarpsPerBlock + warp;     // (node, quad)
    const int node = (int)(unit / p.quads), quad = (int)(unit % p.quads);
    if (node >= p.n_dst) return;                                          // warp-uniform
    const int gate = quad * 4 + grp;
    const bool active = gate < p.G;                                       // idle 8-lane groups when G % 4 != 0
    const int gcol = (active ? gate : 0) * C;                             // idle groups shadow gate 0 (no stores)

    const int beg = __ldg(&p.rowptr[node]), end = __ldg(&p.rowptr[node + 1]);
    const int deg = end - beg;

    // per-lane slices of the value-displacement weights (x,y,z columns of lin_value), fixed for the kernel
    float4 wv[NV][4];
#pragma unroll
    for (int r = 0; r < NV; ++r)
#pragma unroll
        for (int k = 0; k < 4; ++k) wv[r][k] = ldg4(p.Wv3 + (size_t)(gcol + 4 * (sub + 8 * r) + k) * 4);

    const float pix = __ldg(&p.pos_dst[(size_t)node * p.ld_pd]);
    const float piy = __ldg(&p.pos_dst[(size_t)node * p.ld_pd + 1]);
    const float piz = __ldg(&p.pos_dst[(size_t)node * p.ld_pd + 2]);

    float smax = -CUDART_INF_F;
    float4 qx = make_float4(0, 0, 0, 0);
    float4 q[NV];
    if (p.weighted) {
        const float* qrow = p.P_dst + (size_t)node * p.ld_dst;
#pragma unroll
        for (int r = 0; r < NV; ++r) q[r] = ldg4(qrow + p.q_off + gcol + 4 * (sub + 8 * r));
        qx = ldg4(qrow + p.qx_off + 4 * (active ? gate : 0));
        // ---- pass 1: scores --------------------------------------------------------------------------
        for (int e = beg; e < end; ++e) {
            const int j = __ldg(&p.col[e]);
            const float* krow = p.P_src + (size_t)j * p.ld_src + p.k_off + gcol;
            float4 k[NV];
#pragma unroll
            for (int r = 0; r < NV; ++r) k[r] = ldg4(krow + 4 * (sub + 8 * r));
            const float* pj = p.pos_src + (size_t)j * p.ld_ps;
            const float wx = wrapf(__ldg(pj) - pix), wy = wrapf(__ldg(pj + 1) - piy), wz = wrapf(__ldg(pj + 2) - piz);
            float s = dot_group<NV>(q, k);
            s = fmaf(qx.x, wx, s); s = fmaf(qx.y, wy, s); s = fmaf(qx.z, wz, s);
            s = fmaf(qx.w, __ldg(&p.ea[e]), s);
            s /= p.sqrt_c;
            smax = fmaxf(smax, s);
            if (sub == 0 && e - beg < kMaxDegSmem) s_score[warp][grp][e - beg] = s;
        }
        __syncwarp();
    }

    // ---- pass 1.5: softmax denominator (PyG: exp(s - max) / (sum + 1e-16)) ----------------------------------
    float den_eps = 1.0f;
    if (p.weighted) {
        float den = 0.f;
        if (deg <= kMaxDegSmem) {
            for (int e = 0; e < deg; ++e) den += expf(s_score[warp][grp][e] - smax);
        } else {
            for (int e = beg; e < end; ++e) {   // long-row path: recompute the scores
                const int j = __ldg(&p.col[e]);
                const float* krow = p.P_src + (size_t)j * p.ld_src + p.k_off + gcol;
                float4 k[NV];
#pragma unroll
                for (int r = 0; r < NV; ++r) k[r] = ldg4(krow + 4 * (sub + 8 * r));
                const float* pj = p.pos_src + (size_t)j * p.ld_ps;
                const float wx = wrapf(__ldg(pj) - pix), wy = wrapf(__ldg(pj + 1) - piy), wz = wrapf(__ldg(pj + 2) - piz);
                float s = dot_group<NV>(q, k);
                s = fmaf(qx.x, wx, s); s = fmaf(qx.y, wy, s); s = fmaf(qx.z, wz, s);
                s = fmaf(qx.w, __ldg(&p.ea[e]), s);
                den += expf(s / p.sqrt_c - smax);
            }
        }
        den_eps = den + 1e-16f;
    }

    // ---- pass 2: weighted aggregation of relu(V_j + Wv3 (w - p_i)) ------------------------------------------
    float4 acc[NV];
#pragma unroll
    for (int r = 0; r < NV; ++r) acc[r] = make_float4(0, 0, 0, 0);
    float ea_acc = 0.f;
    for (int e = beg; e < end; ++e) {
        const int j = __ldg(&p.col[e]);
        const float* vrow = p.P_src + (size_t)j * p.ld_src + p.v_off + gcol;
        float4 v[NV];
#pragma unroll
        for (int r = 0; r < NV; ++r) v[r] = ldg4(vrow + 4 * (sub + 8 * r));
        const float* pj = p.pos_src + (size_t)j * p.ld_ps;
        const float pjx = __ldg(pj), pjy = __ldg(pj + 1), pjz = __ldg(pj + 2);
        const float tx = wrapf(pjx - pix) - pix, ty = wrapf(pjy - piy) - piy, tz = wrapf(pjz - piz) - piz;
        const float a = __ldg(&p.ea[e]);
        float alpha = 1.0f;
        if (p.weighted) {
            float s;
            if (e - beg < kMaxDegSmem) s = s_score[warp][grp][e - beg];
            else {
                const float* krow = p.P_src + (size_t)j * p.ld_src + p.k_off + gcol;
                float4 k[NV];
#pragma unroll
                for (int r = 0; r < NV; ++r) k[r] = ldg4(krow + 4 * (sub + 8 * r));
                s = dot_group<NV>(q, k);
                s = fmaf(qx.x, wrapf(pjx - pix), s); s = fmaf(qx.y, wrapf(pjy - piy), s); s = fmaf(qx.z, wrapf(pjz - piz), s);
                s = fmaf(qx.w, a, s);
                s /= p.sqrt_c;
            }
            alpha = expf(s - smax) / den_eps;
        }
        ea_acc = fmaf(alpha, a, ea_acc);
#pragma unroll
        for (int r = 0; r < NV; ++r) {
            float x0 = fmaf(wv[r][0].z, tz, fmaf(wv[r][0].y, ty, fmaf(wv[r][0].x, tx, v[r].x)));
            float x1 = fmaf(wv[r][1].z, tz, fmaf(wv[r][1].y, ty, fmaf(wv[r][1].x, tx, v[r].y)));
            float x2 = fmaf(wv[r][2].z, tz, fmaf(wv[r][2].y, ty, fmaf(wv[r][2].x, tx, v[r].z)));
            float x3 = fmaf(wv[r][3].z, tz, fmaf(wv[r][3].y, ty, fmaf(wv[r][3].x, tx, v[r].w)));
            acc[r].x = fmaf(alpha, fmaxf(x0, 0.f), acc[r].x);
            acc[r].y = fmaf(alpha, fmaxf(x1, 0.f), acc[r].y);
            acc[r].z = fmaf(alpha, fmaxf(x2, 0.f), acc[r].z);
            acc[r].w = fmaf(alpha, fmaxf(x3, 0.f), acc[r].w);
        }
    }
    if (active) {
        float* orow = p.agg + (size_t)node * p.ld_agg + gcol;
        if (p.agg_lo) {   // TF32 split for the tensor-core gate GEMM: agg = hi + lo, both TF32-representable
            float* lrow = p.agg_lo + (size_t)node * p.ld_agg + gcol;
#pragma unroll
            for (int r = 0; r < NV; ++r) {
                const float4 hi = make_float4(tf32_rna(acc[r].x), tf32_rna(acc[r].y), tf32_rna(acc[r].z), tf32_rna(acc[r].w));
                const float4 lo = make_float4(tf32_rna(acc[r].x - hi.x), tf32_rna(acc[r].y - hi.y), tf32_rna(acc[r].z - hi.z), tf32_rna(acc[r].w - hi.w));
                *reinterpret_cast<float4*>(orow + 4 * (sub + 8 * r)) = hi;
                *reinterpret_cast<float4*>(lrow + 4 * (sub + 8 * r)) = lo;
            }
        } else
#pragma unroll
        for (int r = 0; r < NV; ++r) *reinterpret_cast<float4*>(orow + 4 * (sub + 8 * r)) = acc[r];
        if (sub == 0) p.ea_out[(size_t)node * p.G + gate] = ea_acc;
    }
}

}  // namespace

extern "C" int gg_pgat_gather(const float* P_src, int32_t ld_src, int32_t k_off, int32_t v_off,
                              const float* P_dst, int32_t ld_dst, int32_t q_off, int32_t qx_off,
                              const float* pos_src, int32_t ld_pos_src, const float* pos_dst, int32_t ld_pos_dst,
                              const int32_t* rowptr, const int32_t* col, const float* eattr_csr,
                              const float* Wv3, int32_t n_dst, int32_t G, int32_t C, int32_t weighted,
                              float* agg, float* agg_lo, int32_t ld_agg, float* ea, void* stream) {
    if (n_dst < 0 || G < 1 || G > 64 || C % 32 || C < 32 || C > 128) return GG_EINVAL;
    if (n_dst == 0) return 0;
    if (!P_src || !pos_src || !pos_dst || !rowptr || !Wv3 || !agg || !ea) return GG_EINVAL;
    if (weighted && !P_dst) return GG_EINVAL;
    if ((ld_src | k_off | v_off | ld_agg) & 3) return GG_EALIGN;
    if (weighted && ((ld_dst | q_off | qx_off) & 3)) return GG_EALIGN;
    if (!gg_aligned16(P_src) || !gg_aligned16(agg) || !gg_aligned16(Wv3) || (weighted && !gg_aligned16(P_dst))) return GG_EALIGN;
    GatherParams p;
    p.P_src = P_src; p.ld_src = ld_src; p.k_off = k_off; p.v_off = v_off;
    p.P_dst = P_dst; p.ld_dst = ld_dst; p.q_off = q_off; p.qx_off = qx_off;
    p.pos_src = pos_src; p.ld_ps = ld_pos_src; p.pos_dst = pos_dst; p.ld_pd = ld_pos_dst;
    p.rowptr = rowptr; p.col = col; p.ea = eattr_csr; p.Wv3 = Wv3;
    p.n_dst = n_dst; p.G = G; p.quads = (G + 3) / 4; p.weighted = weighted ? 1 : 0;
    p.agg = agg; p.agg_lo = agg_lo; p.ld_agg = ld_agg; p.ea_out = ea;
    if (agg_lo && !gg_aligned16(agg_lo)) return GG_EALIGN;
    p.sqrt_c = sqrtf((float)C);
    const int64_t units = (int64_t)n_dst * p.quads;
    const unsigned nb = (unsigned)((units + kWarpsPerBlock - 1) / kWarpsPerBlock);
    cudaStream_t st = GG_STREAM(stream);
    switch (C / 32) {
        case 1: pgat_gather_kernel<1><<<nb, kWarpsPerBlock * 32, 0, st>>>(p); break;
        case 2: pgat_gather_kernel<2><<<nb, kWarpsPerBlock * 32, 0, st>>>(p); break;
        case 3: pgat_gather_kernel<3><<<nb, kWarpsPerBlock * 32, 0, st>>>(p); break;
        default: pgat_gather_kernel<4><<<nb, kWarpsPerBlock * 32, 0, st>>>(p); break;
    }
    GG_LAUNCH_OK();
    return 0;
}
