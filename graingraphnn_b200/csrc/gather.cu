// gather.cu — kernel family (b): fused periodic-attention gather / segment-softmax / aggregate.
//
// Takes over PeriodConv.message + PyG softmax + scatter-add (periodGATconv.py:204-236, :174) for ONE edge type
// and ALL gates of a cell, operating on per-node projections (see gg_node_proj) instead of per-edge GEMMs.
//
// Mapping: one warp per (target node, gate quad).  Inside the warp, 8 lanes own one gate: lane `sub` holds the
// float4 chunks sub, sub+8, ... of that gate's C channels, so each 8-lane group reads 128 contiguous bytes per
// load instruction (LDG.E.128) and a 4-gate row (G*C*4 B = 1.5 KB at C=96) is fetched by 3 fully-coalesced
// requests.  Scores need a reduction over 8 lanes only (3 shuffles).
//
// The kernel is HBM/L2-latency bound (every row is a dependent gather: rowptr -> col -> K/V rows), so it is built
// around memory-level parallelism rather than occupancy:
//   * per-edge metadata (source id, edge length, wrap flags of the periodic displacement) is fetched by ONE lane per
//     edge — lanes 0..deg-1 work in parallel — and broadcast by shuffle, instead of every lane chasing col[e];
//   * edges are processed in chunks of CH: all K rows of a chunk are requested before the first score is reduced,
//     and the V row of an edge is requested as soon as its K registers are dead, so a warp keeps CH x NV 128-bit
//     loads in flight per lane (CH=4, C=96: 6 KB per warp);
//   * rows longer than CH use an online softmax (running max / denominator, rescaled accumulators), so nothing is
//     recomputed or re-read and no scratch memory is needed for any in-degree;
//   * the per-target part of the value displacement, -Wv3 p_i, is hoisted out of the edge loop; the per-edge part
//     Wv3 w_e only exists for edges that cross a periodic/patch boundary (w_e != 0, warp-uniform branch);
//   * streaming operands (Q rows, aggregated outputs) bypass L1 (ld.global.nc.L1::no_allocate / st.global.cs) so the
//     re-used K/V rows of neighbouring targets stay cached.
// No atomics; the edges of a row are accumulated in CSR (= original, stable) order.
//
// Algorithmic bytes per launch = 4*G*C*(2*N_src + 2*N_dst) + 12*(N_src+N_dst) + 4*(N_dst+1) + 8*E
// (read K,V once per source, Q once and write agg once per target; positions; rowptr; col + edge attr).
#include "common.cuh"
#include <math_constants.h>
#include <stdlib.h>

namespace {

constexpr int kWarpsPerBlock = 8;

struct GatherParams {
    const float* P_src; int ld_src, k_off, v_off;
    const float* P_dst; int ld_dst, q_off, qx_off;
    const float* pos_src; int ld_ps;
    const float* pos_dst; int ld_pd;
    const int* rowptr; const int* col; const float* ea;
    const float* Wv3;
    int n_dst, G, quads, weighted;
    float* agg; float* agg_lo; int ld_agg; float* ea_out;
    float inv_sqrt_c;
};

__device__ __forceinline__ float4 ldg4_stream(const float* p) {      // read-once data: do not allocate in L1
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg4_stream(float* p, const float4& v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

__device__ __forceinline__ float group_sum8(float s) {
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    return s;
}

// wrap flags of one coordinate difference packed in 2 bits: 0 -> 0, 1 -> +1, 2 -> -1   (periodGATconv.py:210)
__device__ __forceinline__ int wrap_code(float r) { return (r < -0.5f) ? 1 : ((r > 0.5f) ? 2 : 0); }
__device__ __forceinline__ float wrap_val(int code) { return code == 0 ? 0.f : (code == 1 ? 1.f : -1.f); }

template <int NV>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 2)
pgat_gather_kernel(const GatherParams p) {
    constexpr int C = 32 * NV;
    constexpr int CH = NV >= 4 ? 3 : 4;                  // edges per load batch (register budget: CH*NV float4 buffers)
    extern __shared__ __align__(16) float s_wv[];        // [3][GCpad]: x, y, z columns of lin_value for every gate channel

    const int GCp = p.quads * 4 * C;
    for (int i = threadIdx.x; i < GCp; i += blockDim.x) {
        const bool in = i < p.G * C;
        s_wv[i] = in ? __ldg(&p.Wv3[(size_t)i * 4 + 0]) : 0.f;
        s_wv[GCp + i] = in ? __ldg(&p.Wv3[(size_t)i * 4 + 1]) : 0.f;
        s_wv[2 * GCp + i] = in ? __ldg(&p.Wv3[(size_t)i * 4 + 2]) : 0.f;
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane >> 3, sub = lane & 7;
    const int64_t unit = (int64_t)blockIdx.x * kWarpsPerBlock + warp;     // (node, quad)
    const int node = (int)(unit / p.quads), quad = (int)(unit % p.quads);
    if (node >= p.n_dst) return;                                          // warp-uniform
    const int gate = quad * 4 + grp;
    const bool active = gate < p.G;                                       // idle 8-lane groups when G % 4 != 0
    const int gcol = (active ? gate : 0) * C;                             // idle groups shadow gate 0 (no stores)
    const int wcol = (quad * 4 + grp) * C;                                // column in the (zero-padded) smem copy of Wv3

    const int beg = __ldg(&p.rowptr[node]), end = __ldg(&p.rowptr[node + 1]);

    const float* pd = p.pos_dst + (size_t)node * p.ld_pd;
    const float pix = __ldg(pd), piy = __ldg(pd + 1), piz = __ldg(pd + 2);

    float4 q[NV];
    float4 qx = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.weighted) {
        const float* qrow = p.P_dst + (size_t)node * p.ld_dst;
#pragma unroll
        for (int r = 0; r < NV; ++r) q[r] = ldg4_stream(qrow + p.q_off + gcol + 4 * (sub + 8 * r));
        qx = ldg4(qrow + p.qx_off + 4 * (active ? gate : 0));
    } else {
#pragma unroll
        for (int r = 0; r < NV; ++r) q[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    }

    // per-target part of the displaced value input: V_j + Wv3 (w_e - p_i) = (V_j - Wv3 p_i) + Wv3 w_e
    float4 vp[NV];
#pragma unroll
    for (int r = 0; r < NV; ++r) {
        const int c = wcol + 4 * (sub + 8 * r);
        const float4 wx = *reinterpret_cast<const float4*>(s_wv + c);
        const float4 wy = *reinterpret_cast<const float4*>(s_wv + GCp + c);
        const float4 wz = *reinterpret_cast<const float4*>(s_wv + 2 * GCp + c);
        vp[r].x = -fmaf(wz.x, piz, fmaf(wy.x, piy, wx.x * pix));
        vp[r].y = -fmaf(wz.y, piz, fmaf(wy.y, piy, wx.y * pix));
        vp[r].z = -fmaf(wz.z, piz, fmaf(wy.z, piy, wx.z * pix));
        vp[r].w = -fmaf(wz.w, piz, fmaf(wy.w, piy, wx.w * pix));
    }

    float m_run = -CUDART_INF_F, l_run = 0.f, ea_acc = 0.f;
    float4 acc[NV];
#pragma unroll
    for (int r = 0; r < NV; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int base = beg; base < end; base += 32) {
        // ---- metadata of up to 32 edges, one edge per lane ----------------------------------------------------
        const int my = base + lane;
        int mj = 0, mw = 0;
        float ma = 0.f;
        if (my < end) {
            mj = __ldg(&p.col[my]);
            ma = __ldg(&p.ea[my]);
            const float* pj = p.pos_src + (size_t)mj * p.ld_ps;
            mw = wrap_code(__ldg(pj) - pix) | (wrap_code(__ldg(pj + 1) - piy) << 2) | (wrap_code(__ldg(pj + 2) - piz) << 4);
        }
        const int cnt = min(32, end - base);
        for (int c0 = 0; c0 < cnt; c0 += CH) {
            int j[CH];
            bool on[CH];
#pragma unroll
            for (int e = 0; e < CH; ++e) {
                j[e] = __shfl_sync(0xffffffffu, mj, min(c0 + e, 31));
                on[e] = c0 + e < cnt;                                  // warp-uniform
            }
            // ---- phase 1: all K rows of the chunk in flight, then the scores ------------------------------------
            float s[CH];
            float4 buf[CH][NV];
            if (p.weighted) {
#pragma unroll
                for (int e = 0; e < CH; ++e)
                    if (on[e]) {
                        const float* krow = p.P_src + (size_t)j[e] * p.ld_src + p.k_off + gcol;
#pragma unroll
                        for (int r = 0; r < NV; ++r) buf[e][r] = ldg4(krow + 4 * (sub + 8 * r));
                    }
            }
            float m_new = m_run;
#pragma unroll
            for (int e = 0; e < CH; ++e) {
                s[e] = 0.f;
                if (on[e]) {
                    if (p.weighted) {
                        float d = 0.f;
#pragma unroll
                        for (int r = 0; r < NV; ++r) {
                            d = fmaf(q[r].x, buf[e][r].x, d); d = fmaf(q[r].y, buf[e][r].y, d);
                            d = fmaf(q[r].z, buf[e][r].z, d); d = fmaf(q[r].w, buf[e][r].w, d);
                        }
                        d = group_sum8(d);
                        const int wc = __shfl_sync(0xffffffffu, mw, min(c0 + e, 31));
                        d = fmaf(qx.x, wrap_val(wc & 3), d);
                        d = fmaf(qx.y, wrap_val((wc >> 2) & 3), d);
                        d = fmaf(qx.z, wrap_val((wc >> 4) & 3), d);
                        d = fmaf(qx.w, __shfl_sync(0xffffffffu, ma, min(c0 + e, 31)), d);
                        s[e] = d * p.inv_sqrt_c;
                        m_new = fmaxf(m_new, s[e]);
                    }
                    // the K registers of this edge are dead: request its V row right away
                    const float* vrow = p.P_src + (size_t)j[e] * p.ld_src + p.v_off + gcol;
#pragma unroll
                    for (int r = 0; r < NV; ++r) buf[e][r] = ldg4(vrow + 4 * (sub + 8 * r));
                }
            }
            // ---- online softmax bookkeeping (PyG: exp(s - max) / (sum + 1e-16), periodGATconv.py:227) ------------
            if (p.weighted && m_new > m_run) {
                if (m_run != -CUDART_INF_F) {                           // rescale what earlier chunks accumulated
                    const float scale = expf(m_run - m_new);
                    l_run *= scale; ea_acc *= scale;
#pragma unroll
                    for (int r = 0; r < NV; ++r) { acc[r].x *= scale; acc[r].y *= scale; acc[r].z *= scale; acc[r].w *= scale; }
                }
                m_run = m_new;
            }
            // ---- phase 2: weighted accumulation of relu(V_j + Wv3 (w_e - p_i)) -----------------------------------
#pragma unroll
            for (int e = 0; e < CH; ++e) {
                if (!on[e]) continue;
                const float pe = p.weighted ? expf(s[e] - m_run) : 1.f;
                const int wc = __shfl_sync(0xffffffffu, mw, min(c0 + e, 31));
                l_run += pe;
                ea_acc = fmaf(pe, __shfl_sync(0xffffffffu, ma, min(c0 + e, 31)), ea_acc);
                if (wc != 0) {                                          // edge crosses a periodic / patch boundary
                    const float tx = wrap_val(wc & 3), ty = wrap_val((wc >> 2) & 3), tz = wrap_val((wc >> 4) & 3);
#pragma unroll
                    for (int r = 0; r < NV; ++r) {
                        const int c = wcol + 4 * (sub + 8 * r);
                        const float4 wx = *reinterpret_cast<const float4*>(s_wv + c);
                        const float4 wy = *reinterpret_cast<const float4*>(s_wv + GCp + c);
                        const float4 wz = *reinterpret_cast<const float4*>(s_wv + 2 * GCp + c);
                        buf[e][r].x += fmaf(wz.x, tz, fmaf(wy.x, ty, wx.x * tx));
                        buf[e][r].y += fmaf(wz.y, tz, fmaf(wy.y, ty, wx.y * tx));
                        buf[e][r].z += fmaf(wz.z, tz, fmaf(wy.z, ty, wx.z * tx));
                        buf[e][r].w += fmaf(wz.w, tz, fmaf(wy.w, ty, wx.w * tx));
                    }
                }
#pragma unroll
                for (int r = 0; r < NV; ++r) {
                    acc[r].x = fmaf(pe, fmaxf(buf[e][r].x + vp[r].x, 0.f), acc[r].x);
                    acc[r].y = fmaf(pe, fmaxf(buf[e][r].y + vp[r].y, 0.f), acc[r].y);
                    acc[r].z = fmaf(pe, fmaxf(buf[e][r].z + vp[r].z, 0.f), acc[r].z);
                    acc[r].w = fmaf(pe, fmaxf(buf[e][r].w + vp[r].w, 0.f), acc[r].w);
                }
            }
        }
    }
    if (p.weighted) {
        const float inv = 1.0f / (l_run + 1e-16f);
        ea_acc *= inv;
#pragma unroll
        for (int r = 0; r < NV; ++r) { acc[r].x *= inv; acc[r].y *= inv; acc[r].z *= inv; acc[r].w *= inv; }
    }
    if (active) {
        float* orow = p.agg + (size_t)node * p.ld_agg + gcol;
        if (p.agg_lo) {   // TF32 split for the tensor-core gate GEMM: agg = hi + lo, both TF32-representable
            float* lrow = p.agg_lo + (size_t)node * p.ld_agg + gcol;
#pragma unroll
            for (int r = 0; r < NV; ++r) {
                const float4 hi = make_float4(tf32_rna(acc[r].x), tf32_rna(acc[r].y), tf32_rna(acc[r].z), tf32_rna(acc[r].w));
                const float4 lo = make_float4(tf32_rna(acc[r].x - hi.x), tf32_rna(acc[r].y - hi.y), tf32_rna(acc[r].z - hi.z), tf32_rna(acc[r].w - hi.w));
                stg4_stream(orow + 4 * (sub + 8 * r), hi);
                stg4_stream(lrow + 4 * (sub + 8 * r), lo);
            }
        } else {
#pragma unroll
            for (int r = 0; r < NV; ++r) stg4_stream(orow + 4 * (sub + 8 * r), acc[r]);
        }
        if (sub == 0) p.ea_out[(size_t)node * p.G + gate] = ea_acc;
    }
}


// =====================================================================================================================
// TMA-staged variant (the default on sm_100): every warp runs its own double-buffered bulk-copy pipeline.
//
// Work item = (target node, gate quad, chunk of <= DCAP in-edges).  For item k+1 twelve lanes each issue ONE
// cp.async.bulk (UBLKCP) — the Q row, the QX block and position of the target, and per edge the position, K row and
// V row of the source — into the warp's shared-memory slot (k+1)&1, completing on that slot's mbarrier, while the warp
// computes item k out of slot k&1.  The source ids of item k+2 are fetched (plain loads) in the same iteration, so
// the dependent chain rowptr -> col -> row address never stalls the warp: 3-deep software pipeline, no cross-warp
// synchronisation.  Registers hold only the per-target state (Q, -Wv3 p_i, accumulators), so 10 warps x 2 slots x 10.6 KB
// fill the 227 KB of shared memory of one persistent CTA per SM and keep ~100 KB of loads in flight per SM.
// =====================================================================================================================
constexpr int DCAP = 3;               // in-edges per item (joints of a grain network have exactly 3)

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_init_(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx_(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_(uint32_t bar, uint32_t parity) {
    uint32_t done, spins = 0;
    long long t0 = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (!done && (++spins & 0xfffu) == 0) {              // a lost copy must surface as an error, not as a hung GPU
            if (t0 == 0) t0 = clock64();
            else if (clock64() - t0 > 8000000000LL) __trap();
        }
    } while (!done);
}

struct Item { int node, ebeg, cnt, first, last, valid; };

// warp-uniform walk over this warp's target nodes and their edge chunks; the rowptr pair of the NEXT node is always
// already requested when the current one starts.
struct Cursor {
    int node, stride, n_nodes;
    int end, pos, started;
    int nbeg, nend;
    const int* rowptr;
    __device__ __forceinline__ void prefetch(int n) {
        if (n < n_nodes) { nbeg = __ldg(&rowptr[n]); nend = __ldg(&rowptr[n + 1]); }
    }
    __device__ __forceinline__ void init(int first, int stride_, int n_nodes_, const int* rp) {
        node = first; stride = stride_; n_nodes = n_nodes_; rowptr = rp; started = 0;
        nbeg = nend = 0;
        prefetch(node);
        pos = nbeg; end = nend;
        prefetch(node + stride);
    }
    __device__ __forceinline__ Item next() {
        Item it;
        it.valid = 0; it.node = 0; it.ebeg = 0; it.cnt = 0; it.first = 0; it.last = 0;
        if (node >= n_nodes) return it;
        if (started && pos >= end) {                       // move to the next node
            node += stride;
            if (node >= n_nodes) return it;
            pos = nbeg; end = nend; started = 0;
            prefetch(node + stride);
        }
        it.valid = 1;
        it.node = node;
        it.ebeg = pos; it.cnt = min(DCAP, end - pos);
        it.first = !started; it.last = pos + it.cnt >= end;
        pos += it.cnt; started = 1;
        return it;
    }
};

__device__ __forceinline__ float4 lds4(uint32_t addr) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr));
    return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// periodGATconv.py:210: +1 where r < -0.5, -1 where r > 0.5
__device__ __forceinline__ float wrapf(float r) { return (r < -0.5f ? 1.f : 0.f) - (r > 0.5f ? 1.f : 0.f); }

// One edge of an item out of the staged slot.  Phase 1 (score) and phase 2 (accumulate) are separate calls.
template <int NV>
struct TmaGatherState {
    float4 q[NV], vp[NV], acc[NV];
    float4 qx;
    float pix, piy, piz, m_run, l_run, ea_acc;
};

template <int NV>
__global__ void __launch_bounds__(384, 1)
pgat_gather_tma_kernel(const GatherParams p) {
    constexpr int C = 32 * NV;
    constexpr int ROWB = 4 * C * 4;                         // one row of 4 gates x C floats
    // slot layout: [Q row | QX (64 B)] [P_i 16 B | P_j 3 x 16 B] [edge 0: K row | V row] [edge 1 ...] [edge 2 ...]
    constexpr int OFF_PI = ROWB + 64, OFF_PJ = ROWB + 80, OFF_KV = ROWB + 128;
    constexpr int SLOT = ROWB * (1 + 2 * DCAP) + 128;
    constexpr float LOG2E = 1.4426950408889634f;
    extern __shared__ __align__(128) uint8_t smem[];
    const int n_warps = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = p.G, GC = G * C;                          // G <= 4 here (host dispatch)
    float* s_wv = reinterpret_cast<float*>(smem);           // [3][4*C], zero beyond G*C
    constexpr uint32_t WVB = 3 * 4 * C * 4;
    uint8_t* my = smem + WVB + (size_t)warp * (2 * SLOT);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WVB + (size_t)n_warps * (2 * SLOT)) + 2 * warp;

    for (int i = threadIdx.x; i < 4 * C; i += blockDim.x) {
        const bool in = i < GC;
        s_wv[i] = in ? __ldg(&p.Wv3[(size_t)i * 4 + 0]) : 0.f;
        s_wv[4 * C + i] = in ? __ldg(&p.Wv3[(size_t)i * 4 + 1]) : 0.f;
        s_wv[8 * C + i] = in ? __ldg(&p.Wv3[(size_t)i * 4 + 2]) : 0.f;
    }
    if (lane == 0) {
        mbar_init_(smem_addr(&bars[0]), 1);
        mbar_init_(smem_addr(&bars[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int grp = lane >> 3, sub = lane & 7;
    const bool active = grp < G;
    const int gsel = active ? grp : 0;                      // idle 8-lane groups (G < 4) shadow gate 0 and store nothing
    const uint32_t slot_addr0 = smem_addr(my);
    const uint32_t bar_addr0 = smem_addr(&bars[0]);
    const uint32_t wv_addr = smem_addr(s_wv) + 4u * (grp * C + 4 * sub);
    const uint32_t lane_off = 4u * (gsel * C + 4 * sub);    // byte offset of this lane's first float4 inside a staged row
    const uint32_t rowb = (uint32_t)GC * 4u;                // bytes of one K / V / Q row
    const int w = p.weighted;
    const bool kv_adjacent = w && p.v_off == p.k_off + GC;  // one copy brings K|V
    const bool qx_adjacent = w && p.qx_off == p.q_off + GC; // one copy brings Q|QX
    const float sc2 = p.inv_sqrt_c * LOG2E;                 // scores are kept in log2 units: exp(x) = ex2(x log2 e)

    Cursor cur;
    cur.init(blockIdx.x * n_warps + warp, gridDim.x * n_warps, p.n_dst, p.rowptr);

    auto load_meta = [&](const Item& it, int& mj, float& ma) {
        mj = 0; ma = 0.f;
        if (it.valid && lane < it.cnt) { mj = __ldg(&p.col[it.ebeg + lane]); ma = __ldg(&p.ea[it.ebeg + lane]); }
    };
    auto issue = [&](const Item& it, int mj, int slot) {
        const uint32_t base = slot_addr0 + slot * SLOT, bar = bar_addr0 + 8u * slot;
        const uint32_t tx = (uint32_t)it.cnt * (16u + rowb + (w ? rowb : 0u)) + (it.first ? 16u + (w ? rowb + 16u * G : 0u) : 0u);
        if (lane == 0) mbar_expect_tx_(bar, tx);
        __syncwarp();
        if (lane < it.cnt) {                                             // per edge: position and K|V rows of the source
            const float* prow = p.P_src + (size_t)mj * p.ld_src;
            const uint32_t kv = base + OFF_KV + lane * (2 * ROWB);
            bulk_g2s(base + OFF_PJ + 16 * lane, p.pos_src + (size_t)mj * p.ld_ps, 16, bar);
            if (kv_adjacent) {
                bulk_g2s(kv, prow + p.k_off, 2 * rowb, bar);
            } else {
                if (w) bulk_g2s(kv, prow + p.k_off, rowb, bar);
                bulk_g2s(kv + rowb, prow + p.v_off, rowb, bar);
            }
        } else if (lane == DCAP && it.first) {
            bulk_g2s(base + OFF_PI, p.pos_dst + (size_t)it.node * p.ld_pd, 16, bar);
        } else if (lane == DCAP + 1 && it.first && w) {
            const float* qrow = p.P_dst + (size_t)it.node * p.ld_dst;
            if (qx_adjacent) {
                bulk_g2s(base, qrow + p.q_off, rowb + 16u * G, bar);
            } else {
                bulk_g2s(base, qrow + p.q_off, rowb, bar);
                bulk_g2s(base + rowb, qrow + p.qx_off, 16u * G, bar);
            }
        }
    };

    TmaGatherState<NV> st;
#pragma unroll
    for (int r = 0; r < NV; ++r) { st.q[r] = st.vp[r] = st.acc[r] = make_float4(0.f, 0.f, 0.f, 0.f); }
    st.qx = make_float4(0.f, 0.f, 0.f, 0.f);
    st.pix = st.piy = st.piz = 0.f; st.m_run = -CUDART_INF_F; st.l_run = 0.f; st.ea_acc = 0.f;

    // phase 1 of edge e: score in log2 units (wrap vector w_e = (wx, wy, wz) already known)
    auto score = [&](uint32_t sl, int e, float a, float wx, float wy, float wz) -> float {
        const uint32_t krow = sl + OFF_KV + e * (2 * ROWB) + lane_off;
        float d = 0.f;
#pragma unroll
        for (int r = 0; r < NV; ++r) {
            const float4 k = lds4(krow + 128 * r);
            d = fmaf(st.q[r].x, k.x, d); d = fmaf(st.q[r].y, k.y, d); d = fmaf(st.q[r].z, k.z, d); d = fmaf(st.q[r].w, k.w, d);
        }
        d = group_sum8(d);
        d = fmaf(st.qx.x, wx, d); d = fmaf(st.qx.y, wy, d); d = fmaf(st.qx.z, wz, d); d = fmaf(st.qx.w, a, d);
        return d * sc2;
    };
    // phase 2 of edge e.  relu(v + vp) = max(v, -vp) + vp, so the loop accumulates pe * max(V_j + Wv3 w_e, nvp) with
    // nvp = Wv3 p_i and the target adds vp * sum(pe) once at the end (2 instead of 3 instructions per channel and edge).
    auto accumulate = [&](uint32_t sl, int e, float pe, float a, float wx, float wy, float wz, bool wrapped) {
        st.l_run += pe;
        st.ea_acc = fmaf(pe, a, st.ea_acc);
        const uint32_t vrow = sl + OFF_KV + e * (2 * ROWB) + rowb + lane_off;
        float4 v[NV];
#pragma unroll
        for (int r = 0; r < NV; ++r) v[r] = lds4(vrow + 128 * r);
        if (wrapped) {                                       // edge crosses a periodic / patch boundary (warp-uniform)
#pragma unroll
            for (int r = 0; r < NV; ++r) {
                const float4 ax = lds4(wv_addr + 128 * r), ay = lds4(wv_addr + 16 * C + 128 * r), az = lds4(wv_addr + 32 * C + 128 * r);
                v[r].x += fmaf(az.x, wz, fmaf(ay.x, wy, ax.x * wx));
                v[r].y += fmaf(az.y, wz, fmaf(ay.y, wy, ax.y * wx));
                v[r].z += fmaf(az.z, wz, fmaf(ay.z, wy, ax.z * wx));
                v[r].w += fmaf(az.w, wz, fmaf(ay.w, wy, ax.w * wx));
            }
        }
#pragma unroll
        for (int r = 0; r < NV; ++r) {
            st.acc[r].x = fmaf(pe, fmaxf(v[r].x, st.vp[r].x), st.acc[r].x);
            st.acc[r].y = fmaf(pe, fmaxf(v[r].y, st.vp[r].y), st.acc[r].y);
            st.acc[r].z = fmaf(pe, fmaxf(v[r].z, st.vp[r].z), st.acc[r].z);
            st.acc[r].w = fmaf(pe, fmaxf(v[r].w, st.vp[r].w), st.acc[r].w);
        }
    };

    // one item out of slot SLOT_ID (compile-time, so every shared-memory address is base + immediate)
    auto compute = [&](const Item& itC, float maC, const uint32_t sl, const uint32_t bar, const uint32_t parity) {
        mbar_wait_(bar, parity);
        if (itC.first) {
            const float4 pi4 = lds4(sl + OFF_PI);
            st.pix = pi4.x; st.piy = pi4.y; st.piz = pi4.z;
            if (w) {
#pragma unroll
                for (int r = 0; r < NV; ++r) st.q[r] = lds4(sl + lane_off + 128 * r);
                st.qx = lds4(sl + rowb + 16 * gsel);
            }
#pragma unroll
            for (int r = 0; r < NV; ++r) {   // vp holds nvp = Wv3 p_i  (V_j + Wv3 (w_e - p_i) > 0  <=>  V_j + Wv3 w_e > nvp)
                const float4 ax = lds4(wv_addr + 128 * r), ay = lds4(wv_addr + 16 * C + 128 * r), az = lds4(wv_addr + 32 * C + 128 * r);
                st.vp[r].x = fmaf(az.x, st.piz, fmaf(ay.x, st.piy, ax.x * st.pix));
                st.vp[r].y = fmaf(az.y, st.piz, fmaf(ay.y, st.piy, ax.y * st.pix));
                st.vp[r].z = fmaf(az.z, st.piz, fmaf(ay.z, st.piy, ax.z * st.pix));
                st.vp[r].w = fmaf(az.w, st.piz, fmaf(ay.w, st.piy, ax.w * st.pix));
                st.acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            st.m_run = -CUDART_INF_F; st.l_run = 0.f; st.ea_acc = 0.f;
        }
        // wrap vector of edge `lane` (lanes < cnt), computed once and broadcast: periodGATconv.py:209-210
        float lwx = 0.f, lwy = 0.f, lwz = 0.f;
        if (lane < itC.cnt) {
            const float4 pj = lds4(sl + OFF_PJ + 16 * lane);
            lwx = wrapf(pj.x - st.pix); lwy = wrapf(pj.y - st.piy); lwz = wrapf(pj.z - st.piz);
        }
        const unsigned wrapped_mask = __ballot_sync(0xffffffffu, lwx != 0.f || lwy != 0.f || lwz != 0.f);
        float s[DCAP], a[DCAP], wx[DCAP], wy[DCAP], wz[DCAP];
        float m_new = st.m_run;
#pragma unroll
        for (int e = 0; e < DCAP; ++e) {
            a[e] = __shfl_sync(0xffffffffu, maC, e);
            wx[e] = wy[e] = wz[e] = 0.f; s[e] = 0.f;
        }
        if (wrapped_mask) {                                  // rare: some edge of the item crosses a boundary
#pragma unroll
            for (int e = 0; e < DCAP; ++e) {
                wx[e] = __shfl_sync(0xffffffffu, lwx, e); wy[e] = __shfl_sync(0xffffffffu, lwy, e); wz[e] = __shfl_sync(0xffffffffu, lwz, e);
            }
        }
        if (w) {
            if (itC.cnt == DCAP) {                           // the common case, branch-free over the edges
#pragma unroll
                for (int e = 0; e < DCAP; ++e) { s[e] = score(sl, e, a[e], wx[e], wy[e], wz[e]); m_new = fmaxf(m_new, s[e]); }
            } else {
#pragma unroll
                for (int e = 0; e < DCAP; ++e)
                    if (e < itC.cnt) { s[e] = score(sl, e, a[e], wx[e], wy[e], wz[e]); m_new = fmaxf(m_new, s[e]); }
            }
            if (m_new > st.m_run) {                          // online softmax: rescale what earlier chunks accumulated
                if (st.m_run != -CUDART_INF_F) {
                    const float scale = ex2_approx(st.m_run - m_new);
                    st.l_run *= scale; st.ea_acc *= scale;
#pragma unroll
                    for (int r = 0; r < NV; ++r) { st.acc[r].x *= scale; st.acc[r].y *= scale; st.acc[r].z *= scale; st.acc[r].w *= scale; }
                }
                st.m_run = m_new;
            }
        }
        if (itC.cnt == DCAP) {
#pragma unroll
            for (int e = 0; e < DCAP; ++e)
                accumulate(sl, e, w ? ex2_approx(s[e] - st.m_run) : 1.f, a[e], wx[e], wy[e], wz[e], (wrapped_mask >> e) & 1u);
        } else {
#pragma unroll
            for (int e = 0; e < DCAP; ++e)
                if (e < itC.cnt) accumulate(sl, e, w ? ex2_approx(s[e] - st.m_run) : 1.f, a[e], wx[e], wy[e], wz[e], (wrapped_mask >> e) & 1u);
        }
        if (itC.last) {
            const float inv = w ? 1.0f / (st.l_run + 1e-16f) : 1.0f;      // PyG softmax: exp(s - max) / (sum + 1e-16)
            const float lsum = st.l_run;
            if (active) {
                float* orow = p.agg + (size_t)itC.node * p.ld_agg + grp * C + 4 * sub;
                float* lrow = p.agg_lo ? p.agg_lo + (size_t)itC.node * p.ld_agg + grp * C + 4 * sub : nullptr;
#pragma unroll
                for (int r = 0; r < NV; ++r) {
                    // sum pe * relu(.) = sum pe * max(., nvp) - nvp * sum pe
                    const float4 o = make_float4(fmaf(-st.vp[r].x, lsum, st.acc[r].x) * inv, fmaf(-st.vp[r].y, lsum, st.acc[r].y) * inv,
                                                 fmaf(-st.vp[r].z, lsum, st.acc[r].z) * inv, fmaf(-st.vp[r].w, lsum, st.acc[r].w) * inv);
                    if (lrow) {   // TF32 split for the tensor-core gate GEMM: agg = hi + lo, both TF32-representable
                        const float4 hi = make_float4(tf32_rna(o.x), tf32_rna(o.y), tf32_rna(o.z), tf32_rna(o.w));
                        const float4 lo = make_float4(tf32_rna(o.x - hi.x), tf32_rna(o.y - hi.y), tf32_rna(o.z - hi.z), tf32_rna(o.w - hi.w));
                        stg4_stream(orow + 32 * r, hi);
                        stg4_stream(lrow + 32 * r, lo);
                    } else {
                        stg4_stream(orow + 32 * r, o);
                    }
                }
                if (sub == 0) p.ea_out[(size_t)itC.node * G + grp] = st.ea_acc * inv;
            }
        }
        __syncwarp();                                        // every lane is done with this slot before it is refilled
    };

    // 3-deep software pipeline, unrolled by two so that slot ids are compile-time constants:
    //   stage M: source ids of item k+2 | stage I: bulk copies of item k+1 | stage C: compute item k
    Item itA = cur.next(), itB, itN;
    int mjA, mjB, mjN; float maA, maB, maN;
    load_meta(itA, mjA, maA);
    itB = cur.next();
    load_meta(itB, mjB, maB);
    if (itA.valid) issue(itA, mjA, 0);
    uint32_t par = 0u;
    while (itA.valid) {
        itN = cur.next(); load_meta(itN, mjN, maN);
        if (itB.valid) issue(itB, mjB, 1);
        compute(itA, maA, slot_addr0, bar_addr0, par);
        itA = itN; mjA = mjN; maA = maN;                     // A now holds item k+2 (to be issued into slot 0)
        if (!itB.valid) break;
        itN = cur.next(); load_meta(itN, mjN, maN);
        if (itA.valid) issue(itA, mjA, 0);
        compute(itB, maB, slot_addr0 + SLOT, bar_addr0 + 8u, par);
        itB = itN; mjB = mjN; maB = maN;
        par ^= 1u;
    }
}

static int tma_gather_warps(int C, int quads, size_t* smem_out) {
    const size_t rowb = (size_t)4 * C * 4;
    const size_t slot = rowb * (1 + 2 * DCAP) + 128;
    (void)quads;
    const size_t wv = (size_t)3 * 4 * C * 4;
    const size_t budget = 227 * 1024;
    int warps = 12;                                          // __launch_bounds__(384)
    while (warps > 0 && wv + (size_t)warps * (2 * slot + 16) > budget) --warps;
    *smem_out = wv + (size_t)warps * (2 * slot + 16);
    return warps;
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
    p.inv_sqrt_c = 1.0f / sqrtf((float)C);
    const int64_t units = (int64_t)n_dst * p.quads;
    cudaStream_t st = GG_STREAM(stream);
    cudaError_t err = cudaSuccess;
    // TMA-staged pipeline (default): needs 16-byte aligned position rows for the bulk copies and room for >= 4 warps
    static const int force_ldg = []() { const char* e = getenv("GG_GATHER"); return e && e[0] == 'l' ? 1 : 0; }();
    size_t tma_smem = 0;
    const int tma_warps = tma_gather_warps(C, p.quads, &tma_smem);
    const bool tma_ok = !force_ldg && tma_warps >= 4 && p.quads == 1 && n_dst < (1 << 30) && !((ld_pos_src | ld_pos_dst) & 3) && ld_pos_src >= 4 && ld_pos_dst >= 4 && gg_aligned16(pos_src) && gg_aligned16(pos_dst)
                        && (!weighted || !(qx_off & 3)) && gg_device_is_sm100();
    if (tma_ok) {
        static int n_sms = 0;
        if (n_sms == 0) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev); }
        const int64_t want = ((int64_t)n_dst + tma_warps - 1) / tma_warps;
        const unsigned grid = (unsigned)(want < n_sms ? want : n_sms);
#define GG_GATHER_TMA(NV)                                                                                          \
    do {                                                                                                           \
        err = cudaFuncSetAttribute(pgat_gather_tma_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tma_smem); \
        if (err != cudaSuccess) return (int)err;                                                                   \
        pgat_gather_tma_kernel<NV><<<grid, tma_warps * 32, tma_smem, st>>>(p);                                     \
    } while (0)
        switch (C / 32) {
            case 1: GG_GATHER_TMA(1); break;
            case 2: GG_GATHER_TMA(2); break;
            case 3: GG_GATHER_TMA(3); break;
            default: GG_GATHER_TMA(4); break;
        }
#undef GG_GATHER_TMA
        GG_LAUNCH_OK();
        return 0;
    }
    const unsigned nb = (unsigned)((units + kWarpsPerBlock - 1) / kWarpsPerBlock);
    const size_t smem = (size_t)3 * p.quads * 4 * C * sizeof(float);
    if (smem > 200 * 1024) return GG_EINVAL;
#define GG_GATHER(NV)                                                                                              \
    do {                                                                                                           \
        if (smem > 48 * 1024) err = cudaFuncSetAttribute(pgat_gather_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (err != cudaSuccess) return (int)err;                                                                   \
        pgat_gather_kernel<NV><<<nb, kWarpsPerBlock * 32, smem, st>>>(p);                                          \
    } while (0)
    switch (C / 32) {
        case 1: GG_GATHER(1); break;
        case 2: GG_GATHER(2); break;
        case 3: GG_GATHER(3); break;
        default: GG_GATHER(4); break;
    }
#undef GG_GATHER
    GG_LAUNCH_OK();
    return 0;
}
