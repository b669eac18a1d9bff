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
    int nb, raw_k; const int* items; const int* item_ptr; const int* wrap;   // flat work list (+ per-node offsets) and per-edge wrap codes (TMA path)
    const float* Wv3;
    int n_dst, G, quads, weighted;
    float* agg; int ld_agg; float* ea_out;
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
#pragma unroll
        for (int r = 0; r < NV; ++r) stg4_stream(orow + 4 * (sub + 8 * r), acc[r]);
        if (sub == 0) p.ea_out[(size_t)node * p.G + gate] = ea_acc;
    }
}


// =====================================================================================================================
// TMA-staged variant (the default on sm_100): every warp runs its own double-buffered bulk-copy pipeline over a FLAT list
// of work items.
//
// Work item = (target node, chunk of <= DCAP in-edges), precomputed once per topology by gg_csr_items as int4
// {node, first edge, count | first << 8 | last << 9, -}.  The periodic wrap of every edge is precomputed once per step by
// gg_edge_wrap (2 bits per coordinate), so the kernel touches no source positions.  Per item the warp issues ONE
// cp.async.bulk per edge (the K|V row pair of the source, 3 KB) into its shared-memory slot (k+1)&1, completing on that
// slot's mbarrier, while it computes item k out of slot k&1; Q | QX and the target position of item k+1 travel by plain
// 128-bit loads into a second register set, and the descriptors / source ids of items k+3 / k+2 are fetched in the same
// iteration, so the dependent chain item -> col -> row address never stalls the warp (4-deep software pipeline, no
// cross-warp synchronisation).  Wv3 lives in registers.  12 warps x 2 slots x 9 KB per persistent CTA.
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

struct Meta { int col, wrap; float ea; };            // per lane: edge `lane` of an item (lanes >= cnt: unused)

template <int NV>
__global__ void __launch_bounds__(384, 1)
pgat_gather_tma_kernel(const GatherParams p) {
    constexpr bool RAW = false;                             // raw-score mode has its own kernel (pgat_gather_raw_kernel)
    constexpr int C = 32 * NV;
    constexpr int ROWB = 4 * C * 4;                         // one row of 4 gates x C floats (bytes)
    constexpr int SLOT = DCAP * 2 * ROWB;                   // per edge: [K row | V row]   (RAW: [16 raw floats | V row])
    constexpr int NQ = RAW ? 4 : NV;                        // float4 registers of the target's query
    constexpr float LOG2E = 1.4426950408889634f;
    extern __shared__ __align__(128) uint8_t smem[];
    const int n_warps = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = p.G, GC = G * C;                          // G <= 4 here (host dispatch)
    uint8_t* my = smem + (size_t)warp * (2 * SLOT);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)n_warps * (2 * SLOT)) + 2 * warp;
    if (lane == 0) {
        mbar_init_(smem_addr(&bars[0]), 1);
        mbar_init_(smem_addr(&bars[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    const int grp = lane >> 3, sub = lane & 7;
    const bool active = grp < G;
    const int gsel = active ? grp : 0;                      // idle 8-lane groups (G < 4) shadow gate 0 and store nothing
    const uint32_t slot_addr0 = smem_addr(my);
    const uint32_t bar_addr0 = smem_addr(&bars[0]);
    const uint32_t lane_off = 4u * (gsel * C + 4 * sub);    // byte offset of this lane's first float4 inside a staged row
    const uint32_t rowb = (uint32_t)GC * 4u;                // bytes of one K / V / Q row
    const int w = p.weighted;
    const bool kv_adjacent = w && p.v_off == p.k_off + (RAW ? 16 : GC);  // one copy brings K|V (RAW: raw features | V)
    const uint32_t kbytes = RAW ? 64u : rowb;               // bytes staged in front of the V row
    const float sc2 = p.inv_sqrt_c * LOG2E;                 // scores are kept in log2 units: exp(x) = ex2(x log2 e)

    // x, y, z columns of lin_value for this lane's channels (zero for idle groups): Wv3 is [G*C][4]
    float4 wvx[NV], wvy[NV], wvz[NV];
#pragma unroll
    for (int r = 0; r < NV; ++r) {
        float4 t[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) t[i] = active ? ldg4(p.Wv3 + (size_t)(grp * C + 4 * (sub + 8 * r) + i) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        wvx[r] = make_float4(t[0].x, t[1].x, t[2].x, t[3].x);
        wvy[r] = make_float4(t[0].y, t[1].y, t[2].y, t[3].y);
        wvz[r] = make_float4(t[0].z, t[1].z, t[2].z, t[3].z);
    }

    // Block-cyclic ownership: warp gw owns the node blocks gw, gw + W, ... of NB consecutive targets, i.e. contiguous ranges
    // [item_ptr[NB b], item_ptr[NB (b+1)]) of the item list (all chunks of a node stay in one warp: the softmax state lives
    // in registers), while at any moment the W warps of the grid sweep ~NB W neighbouring targets, whose K|V rows are shared
    // through L2 (a contiguous block per warp was measured to read 2.4x more DRAM bytes).
    const int NB = p.nb;
    const int4* __restrict__ items = reinterpret_cast<const int4*>(p.items);
    const int64_t W = (int64_t)gridDim.x * n_warps, gw = (int64_t)blockIdx.x * n_warps + warp;
    const int64_t n_blocks = ((int64_t)p.n_dst + NB - 1) / NB;
    int64_t blk = gw;                                        // block whose range [cur, cur_end) is being handed out
    int cur = 0, cur_end = 0, nxt = 0, nxt_end = 0;
    auto block_range = [&](int64_t b, int& lo, int& hi) {
        lo = hi = 0;
        if (b < n_blocks) {
            const int64_t n0 = b * NB, n1 = n0 + NB < p.n_dst ? n0 + NB : p.n_dst;
            lo = __ldg(&p.item_ptr[n0]); hi = __ldg(&p.item_ptr[n1]);
        }
    };
    block_range(blk, cur, cur_end);
    block_range(blk + W, nxt, nxt_end);
    // next item descriptor of this warp ({-1,..} when the warp has run out of work)
    auto load_desc = [&]() -> int4 {
        while (cur >= cur_end) {
            if (blk >= n_blocks) return make_int4(-1, 0, 0, 0);
            blk += W; cur = nxt; cur_end = nxt_end;
            block_range(blk + W, nxt, nxt_end);
        }
        return __ldg(items + cur++);
    };
    auto load_meta = [&](const int4& d) -> Meta {
        Meta m; m.col = 0; m.wrap = 0; m.ea = 0.f;
        if (d.x >= 0 && lane < (d.z & 0xff)) {
            m.col = __ldg(&p.col[d.y + lane]); m.ea = __ldg(&p.ea[d.y + lane]); m.wrap = __ldg(&p.wrap[d.y + lane]);
        }
        return m;
    };
    // register set of the NEXT item's target: Q row chunks, QX, position
    float4 qn[NQ], qxn = make_float4(0.f, 0.f, 0.f, 0.f);
    float pnx = 0.f, pny = 0.f, pnz = 0.f;
#pragma unroll
    for (int r = 0; r < NQ; ++r) qn[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    auto issue_copies = [&](const int4& d, const Meta& m, int slot) {
        const int cnt = d.z & 0xff;
        if (cnt > 0) {
            const uint32_t base = slot_addr0 + slot * SLOT, bar = bar_addr0 + 8u * slot;
            if (lane == 0) mbar_expect_tx_(bar, (uint32_t)cnt * (w ? kbytes + rowb : rowb));
            __syncwarp();
            if (lane < cnt) {                                            // per edge: K|V rows of the source
                const float* prow = p.P_src + (size_t)m.col * p.ld_src;
                const uint32_t kv = base + lane * (2 * ROWB);
                if (kv_adjacent) {
                    bulk_g2s(kv, prow + p.k_off, kbytes + rowb, bar);
                } else {
                    if (w) bulk_g2s(kv, prow + p.k_off, kbytes, bar);
                    bulk_g2s(kv + kbytes, prow + p.v_off, rowb, bar);
                }
            }
        }
    };
    auto prefetch_target = [&](const int4& d) {
        if (d.z & 0x100) {                                               // first chunk of a target: its Q | QX and position
            const float* pd = p.pos_dst + (size_t)d.x * p.ld_pd;
            pnx = __ldg(pd); pny = __ldg(pd + 1); pnz = __ldg(pd + 2);
            if (w) {
                const float* qrow = p.P_dst + (size_t)d.x * p.ld_dst;
                if (RAW) {
#pragma unroll
                    for (int r = 0; r < NQ; ++r) qn[r] = ldg4(qrow + p.q_off + gsel * 16 + 4 * r);
                } else {
#pragma unroll
                    for (int r = 0; r < NQ; ++r) qn[r] = ldg4_stream(qrow + p.q_off + gsel * C + 4 * (sub + 8 * r));
                    qxn = ldg4(qrow + p.qx_off + 4 * gsel);
                }
            }
        }
    };

    // per-target state
    float4 q[NQ], vp[NV], acc[NV], qx = make_float4(0.f, 0.f, 0.f, 0.f);
    float m_run = -CUDART_INF_F, l_run = 0.f, ea_acc = 0.f;
#pragma unroll
    for (int r = 0; r < NV; ++r) { vp[r] = acc[r] = make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
    for (int r = 0; r < NQ; ++r) q[r] = make_float4(0.f, 0.f, 0.f, 0.f);

    auto wrapv = [](int code) -> float { return code == 0 ? 0.f : (code == 1 ? 1.f : -1.f); };   // gg_edge_wrap: 1 -> +1, 2 -> -1

    int4 d0 = load_desc(), d1 = load_desc(), d2 = load_desc();
    Meta m0 = load_meta(d0), m1 = load_meta(d1);
    if (d0.x >= 0) { issue_copies(d0, m0, 0); prefetch_target(d0); }
    uint32_t par = 0u;                                       // bit s = phase of slot s
    int slot = 0;
    while (d0.x >= 0) {
        const int4 d3 = load_desc();
        const Meta m2 = load_meta(d2);
        const int cnt = d0.z & 0xff;
        if (d1.x >= 0) issue_copies(d1, m1, slot ^ 1);       // the next item's rows are in flight before anything can stall
        if (d0.z & 0x100) {                                  // adopt the prefetched target registers (before they are reused)
#pragma unroll
            for (int r = 0; r < NQ; ++r) q[r] = qn[r];
#pragma unroll
            for (int r = 0; r < NV; ++r) {                   // vp = Wv3 p_i:  V_j + Wv3 (w_e - p_i) > 0  <=>  V_j + Wv3 w_e > vp
                vp[r].x = fmaf(wvz[r].x, pnz, fmaf(wvy[r].x, pny, wvx[r].x * pnx));
                vp[r].y = fmaf(wvz[r].y, pnz, fmaf(wvy[r].y, pny, wvx[r].y * pnx));
                vp[r].z = fmaf(wvz[r].z, pnz, fmaf(wvy[r].z, pny, wvx[r].z * pnx));
                vp[r].w = fmaf(wvz[r].w, pnz, fmaf(wvy[r].w, pny, wvx[r].w * pnx));
                acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            qx = RAW ? q[0] : qxn;                           // RAW: Q'[0:3] = Wk3^T q
            m_run = -CUDART_INF_F; l_run = 0.f; ea_acc = 0.f;
        }
        if (d1.x >= 0) prefetch_target(d1);

        // ---------------------------------------------------------------- compute item d0 out of `slot`
        const uint32_t sl = slot_addr0 + slot * SLOT;
        if (cnt > 0) {
            mbar_wait_(bar_addr0 + 8u * slot, (par >> slot) & 1u);
            par ^= 1u << slot;
            float a[DCAP], s[DCAP];
            int wc[DCAP];
#pragma unroll
            for (int e = 0; e < DCAP; ++e) { a[e] = __shfl_sync(0xffffffffu, m0.ea, e); wc[e] = __shfl_sync(0xffffffffu, m0.wrap, e); s[e] = 0.f; }
            float m_new = m_run;
            if (w) {
#pragma unroll
                for (int e = 0; e < DCAP; ++e) {
                    if (e < cnt) {                           // warp-uniform
                        float dd = 0.f, d2_ = 0.f;           // two partial sums: shorter dependency chains
                        if (RAW) {
                            const uint32_t xrow = sl + e * (2 * ROWB);          // same 64 bytes for every lane (broadcast)
                            float4 x0 = lds4(xrow), x1 = lds4(xrow + 16), x2 = lds4(xrow + 32), x3 = lds4(xrow + 48);
                            x3.w = a[e];                                         // Q'[15] = We . q multiplies the edge length
                            dd = fmaf(q[0].x, x0.x, dd); d2_ = fmaf(q[0].y, x0.y, d2_); dd = fmaf(q[0].z, x0.z, dd); d2_ = fmaf(q[0].w, x0.w, d2_);
                            dd = fmaf(q[1].x, x1.x, dd); d2_ = fmaf(q[1].y, x1.y, d2_); dd = fmaf(q[1].z, x1.z, dd); d2_ = fmaf(q[1].w, x1.w, d2_);
                            dd = fmaf(q[2].x, x2.x, dd); d2_ = fmaf(q[2].y, x2.y, d2_); dd = fmaf(q[2].z, x2.z, dd); d2_ = fmaf(q[2].w, x2.w, d2_);
                            dd = fmaf(q[3].x, x3.x, dd); d2_ = fmaf(q[3].y, x3.y, d2_); dd = fmaf(q[3].z, x3.z, dd); d2_ = fmaf(q[3].w, x3.w, d2_);
                            dd += d2_;
                        } else {
                            const uint32_t krow = sl + e * (2 * ROWB) + lane_off;
#pragma unroll
                            for (int r = 0; r < NV; ++r) {
                                const float4 k = lds4(krow + 128 * r);
                                dd = fmaf(q[r].x, k.x, dd); d2_ = fmaf(q[r].y, k.y, d2_); dd = fmaf(q[r].z, k.z, dd); d2_ = fmaf(q[r].w, k.w, d2_);
                            }
                            dd = group_sum8(dd + d2_);
                        }
                        if (wc[e]) {                         // periodGATconv.py:209-211: the wrapped displacement enters the key
                            dd = fmaf(qx.x, wrapv(wc[e] & 3), dd); dd = fmaf(qx.y, wrapv((wc[e] >> 2) & 3), dd); dd = fmaf(qx.z, wrapv((wc[e] >> 4) & 3), dd);
                        }
                        if (!RAW) dd = fmaf(qx.w, a[e], dd);
                        s[e] = dd * sc2;
                        m_new = fmaxf(m_new, s[e]);
                    }
                }
                if (m_new > m_run) {                         // online softmax: rescale what earlier chunks accumulated
                    if (m_run != -CUDART_INF_F) {
                        const float scale = ex2_approx(m_run - m_new);
                        l_run *= scale; ea_acc *= scale;
#pragma unroll
                        for (int r = 0; r < NV; ++r) { acc[r].x *= scale; acc[r].y *= scale; acc[r].z *= scale; acc[r].w *= scale; }
                    }
                    m_run = m_new;
                }
            }
            // relu(v + Wv3 (w_e - p_i)) = max(V_j + Wv3 w_e, vp) - vp: accumulate pe * max(., vp); the target subtracts vp * sum(pe) once
#pragma unroll
            for (int e = 0; e < DCAP; ++e) {
                if (e < cnt) {
                    const float pe = w ? ex2_approx(s[e] - m_run) : 1.f;
                    l_run += pe;
                    ea_acc = fmaf(pe, a[e], ea_acc);
                    const uint32_t vrow = sl + e * (2 * ROWB) + kbytes + lane_off;
                    float4 v[NV];
#pragma unroll
                    for (int r = 0; r < NV; ++r) v[r] = lds4(vrow + 128 * r);
                    if (wc[e]) {                             // edge crosses a periodic / patch boundary (warp-uniform, rare)
                        const float tx = wrapv(wc[e] & 3), ty = wrapv((wc[e] >> 2) & 3), tz = wrapv((wc[e] >> 4) & 3);
#pragma unroll
                        for (int r = 0; r < NV; ++r) {
                            v[r].x += fmaf(wvz[r].x, tz, fmaf(wvy[r].x, ty, wvx[r].x * tx));
                            v[r].y += fmaf(wvz[r].y, tz, fmaf(wvy[r].y, ty, wvx[r].y * tx));
                            v[r].z += fmaf(wvz[r].z, tz, fmaf(wvy[r].z, ty, wvx[r].z * tx));
                            v[r].w += fmaf(wvz[r].w, tz, fmaf(wvy[r].w, ty, wvx[r].w * tx));
                        }
                    }
#pragma unroll
                    for (int r = 0; r < NV; ++r) {
                        acc[r].x = fmaf(pe, fmaxf(v[r].x, vp[r].x), acc[r].x);
                        acc[r].y = fmaf(pe, fmaxf(v[r].y, vp[r].y), acc[r].y);
                        acc[r].z = fmaf(pe, fmaxf(v[r].z, vp[r].z), acc[r].z);
                        acc[r].w = fmaf(pe, fmaxf(v[r].w, vp[r].w), acc[r].w);
                    }
                }
            }
        }
        if (d0.z & 0x200) {                                  // last chunk of the target: normalise and store
            const float inv = w ? 1.0f / (l_run + 1e-16f) : 1.0f;      // PyG softmax: exp(s - max) / (sum + 1e-16)
            if (active) {
                float* orow = p.agg + (size_t)d0.x * p.ld_agg + grp * C + 4 * sub;
#pragma unroll
                for (int r = 0; r < NV; ++r)                 // sum pe * relu(.) = sum pe * max(., vp) - vp * sum pe
                    stg4_stream(orow + 32 * r, make_float4(fmaf(-vp[r].x, l_run, acc[r].x) * inv, fmaf(-vp[r].y, l_run, acc[r].y) * inv,
                                                           fmaf(-vp[r].z, l_run, acc[r].z) * inv, fmaf(-vp[r].w, l_run, acc[r].w) * inv));
                if (sub == 0) p.ea_out[(size_t)d0.x * G + grp] = ea_acc * inv;
            }
        }
        __syncwarp();                                        // every lane is done with this slot before it is refilled
        d0 = d1; m0 = m1; d1 = d2; m1 = m2; d2 = d3;
        slot ^= 1;
    }
}

// =====================================================================================================================
// Raw-score variant (cells without hidden state, i.e. the encoder): q_i . (Wk x_j) = x_j . (Wk^T q_i), so no key rows exist.
// The staged source row is [raw features (16 floats) | V row]; the target carries Q' = [Wk[:, :F]^T q | We . q] (16 floats per
// gate) and the score is the 16-term dot product x_j . Q' (the last term multiplies the edge length), computed redundantly by
// the 8 lanes of a gate group: no key rows, no shuffles.  An item needs ~5 KB of shared memory instead of 9 KB, which buys a
// THIRD slot per warp: two items are in flight while one is computed (the classic kernel is bound by DRAM latency with one
// item of lookahead).  Q' and the target position travel in the slot too (bulk copies), so no register set is tied up.
// =====================================================================================================================
constexpr int RAW_NSLOT = 3;

template <int NV>
__global__ void __launch_bounds__(384, 1)
pgat_gather_raw_kernel(const GatherParams p) {
    constexpr int C = 32 * NV;
    constexpr int ROWB = 4 * C * 4;                         // V row of 4 gates (bytes)
    constexpr int ES = 64 + ROWB;                           // per edge: [16 raw floats | V row]
    constexpr int HDR = 320;                                // [Q' of 4 gates (256 B) | target position (16 B) | pad]
    constexpr int SLOT = HDR + DCAP * ES;
    constexpr int D = RAW_NSLOT - 1;                        // items in flight ahead of the one being computed
    constexpr float LOG2E = 1.4426950408889634f;
    extern __shared__ __align__(128) uint8_t smem[];
    const int n_warps = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = p.G, GC = G * C;
    uint8_t* my = smem + (size_t)warp * (RAW_NSLOT * SLOT);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)n_warps * (RAW_NSLOT * SLOT)) + RAW_NSLOT * warp;
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < RAW_NSLOT; ++i) mbar_init_(smem_addr(&bars[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    const int grp = lane >> 3, sub = lane & 7;
    const bool active = grp < G;
    const int gsel = active ? grp : 0;
    const uint32_t slot_addr0 = smem_addr(my);
    const uint32_t bar_addr0 = smem_addr(&bars[0]);
    const uint32_t lane_off = 4u * (gsel * C + 4 * sub);
    const uint32_t rowb = (uint32_t)GC * 4u;
    const float sc2 = p.inv_sqrt_c * LOG2E;

    float4 wvx[NV], wvy[NV], wvz[NV];
#pragma unroll
    for (int r = 0; r < NV; ++r) {
        float4 t[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) t[i] = active ? ldg4(p.Wv3 + (size_t)(grp * C + 4 * (sub + 8 * r) + i) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        wvx[r] = make_float4(t[0].x, t[1].x, t[2].x, t[3].x);
        wvy[r] = make_float4(t[0].y, t[1].y, t[2].y, t[3].y);
        wvz[r] = make_float4(t[0].z, t[1].z, t[2].z, t[3].z);
    }

    // block-cyclic ownership of the item list (see pgat_gather_tma_kernel)
    const int NB = p.nb;
    const int4* __restrict__ items = reinterpret_cast<const int4*>(p.items);
    const int64_t W = (int64_t)gridDim.x * n_warps, gw = (int64_t)blockIdx.x * n_warps + warp;
    const int64_t n_blocks = ((int64_t)p.n_dst + NB - 1) / NB;
    int64_t blk = gw;
    int cur = 0, cur_end = 0, nxt = 0, nxt_end = 0;
    auto block_range = [&](int64_t b, int& lo, int& hi) {
        lo = hi = 0;
        if (b < n_blocks) {
            const int64_t n0 = b * NB, n1 = n0 + NB < p.n_dst ? n0 + NB : p.n_dst;
            lo = __ldg(&p.item_ptr[n0]); hi = __ldg(&p.item_ptr[n1]);
        }
    };
    block_range(blk, cur, cur_end);
    block_range(blk + W, nxt, nxt_end);
    auto load_desc = [&]() -> int4 {
        while (cur >= cur_end) {
            if (blk >= n_blocks) return make_int4(-1, 0, 0, 0);
            blk += W; cur = nxt; cur_end = nxt_end;
            block_range(blk + W, nxt, nxt_end);
        }
        return __ldg(items + cur++);
    };
    auto load_meta = [&](const int4& d) -> Meta {
        Meta m; m.col = 0; m.wrap = 0; m.ea = 0.f;
        if (d.x >= 0 && lane < (d.z & 0xff)) {
            m.col = __ldg(&p.col[d.y + lane]); m.ea = __ldg(&p.ea[d.y + lane]); m.wrap = __ldg(&p.wrap[d.y + lane]);
        }
        return m;
    };
    auto issue = [&](const int4& d, const Meta& m, int slot) {
        if (d.x < 0) return;
        const int cnt = d.z & 0xff;
        const bool first = d.z & 0x100;
        const uint32_t tx = (uint32_t)cnt * (64u + rowb) + (first ? 64u * G + 16u : 0u);
        if (tx == 0) return;
        const uint32_t base = slot_addr0 + slot * SLOT, bar = bar_addr0 + 8u * slot;
        if (lane == 0) mbar_expect_tx_(bar, tx);
        __syncwarp();
        if (lane < cnt) {                                                // per edge: [raw features | V] of the source
            bulk_g2s(base + HDR + lane * ES, p.P_src + (size_t)m.col * p.ld_src + p.k_off, 64u + rowb, bar);
        } else if (lane == DCAP && first) {
            bulk_g2s(base, p.P_dst + (size_t)d.x * p.ld_dst + p.q_off, 64u * G, bar);
        } else if (lane == DCAP + 1 && first) {
            bulk_g2s(base + 256, p.pos_dst + (size_t)d.x * p.ld_pd, 16, bar);
        }
    };

    float4 q[4], vp[NV], acc[NV];
    float m_run = -CUDART_INF_F, l_run = 0.f, ea_acc = 0.f;
#pragma unroll
    for (int r = 0; r < NV; ++r) { vp[r] = acc[r] = make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
    for (int r = 0; r < 4; ++r) q[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    auto wrapv = [](int code) -> float { return code == 0 ? 0.f : (code == 1 ? 1.f : -1.f); };

    // descriptor / metadata queues: d[i], m[i] belong to the item i positions ahead of the one being computed
    int4 d[D + 2];
    Meta m[D + 1];
#pragma unroll
    for (int i = 0; i < D + 2; ++i) d[i] = load_desc();
#pragma unroll
    for (int i = 0; i < D + 1; ++i) m[i] = load_meta(d[i]);
#pragma unroll
    for (int i = 0; i < D; ++i) issue(d[i], m[i], i);
    uint32_t par = 0u;
    int slot = 0;
    while (d[0].x >= 0) {
        const int4 d_new = load_desc();
        const Meta m_new = load_meta(d[D + 1]);
        { int s2 = slot + D; if (s2 >= RAW_NSLOT) s2 -= RAW_NSLOT; issue(d[D], m[D], s2); }   // slot of the item computed last

        const int4 d0 = d[0];
        const Meta m0 = m[0];
        const int cnt = d0.z & 0xff;
        const bool first = d0.z & 0x100;
        const uint32_t sl = slot_addr0 + slot * SLOT;
        if (cnt > 0 || first) {
            mbar_wait_(bar_addr0 + 8u * slot, (par >> slot) & 1u);
            par ^= 1u << slot;
        }
        if (first) {
            const float4 pi = lds4(sl + 256);
#pragma unroll
            for (int r = 0; r < 4; ++r) q[r] = lds4(sl + gsel * 64 + 16 * r);
#pragma unroll
            for (int r = 0; r < NV; ++r) {                   // vp = Wv3 p_i
                vp[r].x = fmaf(wvz[r].x, pi.z, fmaf(wvy[r].x, pi.y, wvx[r].x * pi.x));
                vp[r].y = fmaf(wvz[r].y, pi.z, fmaf(wvy[r].y, pi.y, wvx[r].y * pi.x));
                vp[r].z = fmaf(wvz[r].z, pi.z, fmaf(wvy[r].z, pi.y, wvx[r].z * pi.x));
                vp[r].w = fmaf(wvz[r].w, pi.z, fmaf(wvy[r].w, pi.y, wvx[r].w * pi.x));
                acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            m_run = -CUDART_INF_F; l_run = 0.f; ea_acc = 0.f;
        }
        if (cnt > 0) {
            float a[DCAP], s[DCAP];
            int wc[DCAP];
#pragma unroll
            for (int e = 0; e < DCAP; ++e) { a[e] = __shfl_sync(0xffffffffu, m0.ea, e); wc[e] = __shfl_sync(0xffffffffu, m0.wrap, e); s[e] = 0.f; }
            float m_new_ = m_run;
#pragma unroll
            for (int e = 0; e < DCAP; ++e) {
                if (e < cnt) {                               // warp-uniform
                    const uint32_t xrow = sl + HDR + e * ES;                 // same 64 bytes for every lane (broadcast)
                    float4 x0 = lds4(xrow), x1 = lds4(xrow + 16), x2 = lds4(xrow + 32), x3 = lds4(xrow + 48);
                    x3.w = a[e];                                             // Q'[15] = We . q multiplies the edge length
                    float dd = 0.f, d2_ = 0.f;
                    dd = fmaf(q[0].x, x0.x, dd); d2_ = fmaf(q[0].y, x0.y, d2_); dd = fmaf(q[0].z, x0.z, dd); d2_ = fmaf(q[0].w, x0.w, d2_);
                    dd = fmaf(q[1].x, x1.x, dd); d2_ = fmaf(q[1].y, x1.y, d2_); dd = fmaf(q[1].z, x1.z, dd); d2_ = fmaf(q[1].w, x1.w, d2_);
                    dd = fmaf(q[2].x, x2.x, dd); d2_ = fmaf(q[2].y, x2.y, d2_); dd = fmaf(q[2].z, x2.z, dd); d2_ = fmaf(q[2].w, x2.w, d2_);
                    dd = fmaf(q[3].x, x3.x, dd); d2_ = fmaf(q[3].y, x3.y, d2_); dd = fmaf(q[3].z, x3.z, dd); d2_ = fmaf(q[3].w, x3.w, d2_);
                    dd += d2_;
                    if (wc[e]) {                             // Q'[0:3] = Wk3^T q meets the wrap vector (periodGATconv.py:209-211)
                        dd = fmaf(q[0].x, wrapv(wc[e] & 3), dd); dd = fmaf(q[0].y, wrapv((wc[e] >> 2) & 3), dd); dd = fmaf(q[0].z, wrapv((wc[e] >> 4) & 3), dd);
                    }
                    s[e] = dd * sc2;
                    m_new_ = fmaxf(m_new_, s[e]);
                }
            }
            if (m_new_ > m_run) {                            // online softmax: rescale what earlier chunks accumulated
                if (m_run != -CUDART_INF_F) {
                    const float scale = ex2_approx(m_run - m_new_);
                    l_run *= scale; ea_acc *= scale;
#pragma unroll
                    for (int r = 0; r < NV; ++r) { acc[r].x *= scale; acc[r].y *= scale; acc[r].z *= scale; acc[r].w *= scale; }
                }
                m_run = m_new_;
            }
#pragma unroll
            for (int e = 0; e < DCAP; ++e) {
                if (e < cnt) {
                    const float pe = ex2_approx(s[e] - m_run);
                    l_run += pe;
                    ea_acc = fmaf(pe, a[e], ea_acc);
                    const uint32_t vrow = sl + HDR + e * ES + 64 + lane_off;
                    float4 v[NV];
#pragma unroll
                    for (int r = 0; r < NV; ++r) v[r] = lds4(vrow + 128 * r);
                    if (wc[e]) {
                        const float tx = wrapv(wc[e] & 3), ty = wrapv((wc[e] >> 2) & 3), tz = wrapv((wc[e] >> 4) & 3);
#pragma unroll
                        for (int r = 0; r < NV; ++r) {
                            v[r].x += fmaf(wvz[r].x, tz, fmaf(wvy[r].x, ty, wvx[r].x * tx));
                            v[r].y += fmaf(wvz[r].y, tz, fmaf(wvy[r].y, ty, wvx[r].y * tx));
                            v[r].z += fmaf(wvz[r].z, tz, fmaf(wvy[r].z, ty, wvx[r].z * tx));
                            v[r].w += fmaf(wvz[r].w, tz, fmaf(wvy[r].w, ty, wvx[r].w * tx));
                        }
                    }
#pragma unroll
                    for (int r = 0; r < NV; ++r) {
                        acc[r].x = fmaf(pe, fmaxf(v[r].x, vp[r].x), acc[r].x);
                        acc[r].y = fmaf(pe, fmaxf(v[r].y, vp[r].y), acc[r].y);
                        acc[r].z = fmaf(pe, fmaxf(v[r].z, vp[r].z), acc[r].z);
                        acc[r].w = fmaf(pe, fmaxf(v[r].w, vp[r].w), acc[r].w);
                    }
                }
            }
        }
        if (d0.z & 0x200) {                                  // last chunk of the target: normalise and store
            const float inv = 1.0f / (l_run + 1e-16f);       // PyG softmax: exp(s - max) / (sum + 1e-16)
            if (active) {
                float* orow = p.agg + (size_t)d0.x * p.ld_agg + grp * C + 4 * sub;
#pragma unroll
                for (int r = 0; r < NV; ++r)
                    stg4_stream(orow + 32 * r, make_float4(fmaf(-vp[r].x, l_run, acc[r].x) * inv, fmaf(-vp[r].y, l_run, acc[r].y) * inv,
                                                           fmaf(-vp[r].z, l_run, acc[r].z) * inv, fmaf(-vp[r].w, l_run, acc[r].w) * inv));
                if (sub == 0) p.ea_out[(size_t)d0.x * G + grp] = ea_acc * inv;
            }
        }
        __syncwarp();                                        // every lane is done with this slot before it is refilled
#pragma unroll
        for (int i = 0; i < D + 1; ++i) d[i] = d[i + 1];
        d[D + 1] = d_new;
#pragma unroll
        for (int i = 0; i < D; ++i) m[i] = m[i + 1];
        m[D] = m_new;
        if (++slot == RAW_NSLOT) slot = 0;
    }
}

static int raw_gather_warps(int C, size_t* smem_out) {
    const size_t slot = 320 + (size_t)DCAP * (64 + 4 * C * 4);
    const size_t per_warp = RAW_NSLOT * slot + 8 * RAW_NSLOT;
    int warps = 12;                                          // __launch_bounds__(384)
    while (warps > 0 && (size_t)warps * per_warp > 227 * 1024) --warps;
    *smem_out = (size_t)warps * per_warp;
    return warps;
}

static int tma_gather_warps(int C, size_t* smem_out) {
    const size_t slot = (size_t)DCAP * 2 * 4 * C * 4;
    const size_t budget = 227 * 1024;
    int warps = 12;                                          // __launch_bounds__(384)
    while (warps > 0 && (size_t)warps * (2 * slot + 16) > budget) --warps;
    *smem_out = (size_t)warps * (2 * slot + 16);
    return warps;
}

}  // namespace

extern "C" int gg_gather_dcap(void) { return DCAP; }

extern "C" int gg_pgat_gather(const float* P_src, int32_t ld_src, int32_t k_off, int32_t v_off,
                              const float* P_dst, int32_t ld_dst, int32_t q_off, int32_t qx_off,
                              const float* pos_src, int32_t ld_pos_src, const float* pos_dst, int32_t ld_pos_dst,
                              const int32_t* rowptr, const int32_t* col, const float* eattr_csr,
                              const int32_t* items, const int32_t* item_ptr, const int32_t* wrap_csr, int32_t raw_k,
                              const float* Wv3, int32_t n_dst, int32_t G, int32_t C, int32_t weighted,
                              float* agg, int32_t ld_agg, float* ea, void* stream) {
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
    p.items = items; p.item_ptr = item_ptr; p.wrap = wrap_csr; p.raw_k = raw_k;
    if (raw_k != 0 && (raw_k != 16 || !weighted || v_off != k_off + 16)) return GG_EINVAL;
    static const int nb_env = []() { const char* e = getenv("GG_GATHER_NB"); return e ? atoi(e) : 0; }();
    p.nb = nb_env > 0 ? nb_env : 8;                          // measured on B200, 250k targets: 1: 4.25 | 4: 3.87 | 8: 3.70 | 16: 3.96 | 64: 5.50 ms per step
    p.n_dst = n_dst; p.G = G; p.quads = (G + 3) / 4; p.weighted = weighted ? 1 : 0;
    p.agg = agg; p.ld_agg = ld_agg; p.ea_out = ea;
    p.inv_sqrt_c = 1.0f / sqrtf((float)C);
    const int64_t units = (int64_t)n_dst * p.quads;
    cudaStream_t st = GG_STREAM(stream);
    cudaError_t err = cudaSuccess;
    // TMA-staged pipeline (default): needs the flat item list (gg_csr_items), the per-edge wrap codes (gg_edge_wrap) and G <= 4
    static const int force_ldg = []() { const char* e = getenv("GG_GATHER"); return e && e[0] == 'l' ? 1 : 0; }();
    size_t tma_smem = 0;
    const int tma_warps = tma_gather_warps(C, &tma_smem);
    const bool tma_ok = !force_ldg && items && item_ptr && wrap_csr && tma_warps >= 4 && p.quads == 1 && gg_device_is_sm100();
    if (tma_ok) {
        static int n_sms = 0;
        if (n_sms == 0) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev); }
        static const int sm_cap = []() { const char* e = getenv("GG_GATHER_SMS"); return e ? atoi(e) : 0; }();   // experiments
        const int use_sms = sm_cap > 0 && sm_cap < n_sms ? sm_cap : n_sms;
        const int64_t want = ((int64_t)n_dst + tma_warps - 1) / tma_warps;
        const unsigned grid = (unsigned)(want < use_sms ? want : use_sms);
#define GG_GATHER_TMA(NV)                                                                                          \
    do {                                                                                                           \
        err = cudaFuncSetAttribute(pgat_gather_tma_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tma_smem); \
        if (err != cudaSuccess) return (int)err;                                                                   \
        pgat_gather_tma_kernel<NV><<<grid, tma_warps * 32, tma_smem, st>>>(p);                                     \
    } while (0)
        if (raw_k) {
            size_t raw_smem = 0;
            const int raw_warps = raw_gather_warps(C, &raw_smem);
            const int64_t want_r = ((int64_t)n_dst + raw_warps - 1) / raw_warps;
            const unsigned grid_r = (unsigned)(want_r < use_sms ? want_r : use_sms);
#define GG_GATHER_RAW(NV)                                                                                          \
    do {                                                                                                           \
        err = cudaFuncSetAttribute(pgat_gather_raw_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)raw_smem); \
        if (err != cudaSuccess) return (int)err;                                                                   \
        pgat_gather_raw_kernel<NV><<<grid_r, raw_warps * 32, raw_smem, st>>>(p);                                   \
    } while (0)
            switch (C / 32) {
                case 1: GG_GATHER_RAW(1); break;
                case 2: GG_GATHER_RAW(2); break;
                case 3: GG_GATHER_RAW(3); break;
                default: GG_GATHER_RAW(4); break;
            }
#undef GG_GATHER_RAW
        } else {
            switch (C / 32) {
                case 1: GG_GATHER_TMA(1); break;
                case 2: GG_GATHER_TMA(2); break;
                case 3: GG_GATHER_TMA(3); break;
                default: GG_GATHER_TMA(4); break;
            }
        }
#undef GG_GATHER_TMA
        GG_LAUNCH_OK();
        return 0;
    }
    if (raw_k) return GG_EINVAL;                 // raw-score mode exists in the bulk-copy kernel only
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
