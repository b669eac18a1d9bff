// gather_tiled.cu — kernel family (b), warp-specialised form: fused periodic-attention gather / segment softmax / aggregate
// (PeriodConv.message + PyG softmax + scatter-add, periodGATconv.py:204-236, :174) for ONE edge type and ALL gates of a cell.
//
// Why this form: the per-warp pipelines of gather.cu spend ~350 of their ~570 instructions per target on bookkeeping (work-list
// queues, address arithmetic, copy issue loops) and run 3 warps per scheduler, so they are bound by instruction issue at ~30 %
// of the HBM roofline although every copy has landed before it is waited for (ncu: profiles/r1_gather_v4_ncu_summary.csv).
// Here the irregular part lives in dedicated producer warps and the consumers execute straight-line arithmetic out of shared
// memory:
//
//   * Targets are grouped into UNITS by the CSR position of their first in-edge (unit k: first edge in [k B, (k+1) B), B = 4 ECAP; measured 2 / 4 / 8 tiles per unit: 4 is best)
//     and the edges of a unit are cut into TILES of ECAP consecutive in-edges (the last one shorter) — independent of where
//     target rows begin and end, so any in-degree works with a fixed shared-memory slot count.  gg_csr_compact lists the targets
//     with in-edges (nz, nzptr); gg_csr_tiles emits the tile list {first edge, edges, first target, last target} ordered by
//     CTA: unit k belongs to CTA k mod n_ctas (neighbouring units, which share source rows, are in flight together on
//     different SMs and meet in L2), and cta_ptr[b] .. cta_ptr[b + 1] are the tiles of CTA b in processing order.
//   * Persistent CTA per SM.  NP producer warps: per tile, one cp.async.bulk per DISTINCT source row brings it into the stage
//     (neighbouring targets share sources; __match_any_sync finds the repeats inside a tile and every edge records the slot its
//     row is staged in) and one per starting target brings its block (Q|QX|position, or Q' with the position in three spare
//     slots); edge lengths, wrap codes, row slots and the target descriptors are written to the stage by the producer lanes.
//     Everything of a tile completes on ONE mbarrier (expect_tx); NS stages, full/empty barrier pair per stage.
//   * Three row forms (template MODE): 0 = [K | V] with Q | QX targets; 1 = raw scores without hidden state (encoder):
//     [16 raw features | V], Q' of 16 floats per gate; 2 = raw scores on the whole cell input (decoder): [X padded to 32 | h | V],
//     Q' of 32 + C floats per gate.  q.(Wk x) = x.(Wk^T q): the raw forms need no key rows, 2 KB instead of 3 KB per decoder edge.
//   * NC consumer warps: target i (index in the compacted list) belongs to warp i mod NC in EVERY tile, so a row that straddles a
//     tile boundary (always inside one unit, hence inside one CTA) stays with its warp and the online-softmax state (running
//     max, denominator, accumulators) simply stays in registers across tiles.  8 lanes per gate, 128-bit shared-memory loads, no atomics, edges of a row accumulate in CSR order.
//   * Targets without in-edges get their zero rows from a short pass at the end.
//
// Algorithmic bytes per launch: see gather.cu (identical data, identical arithmetic).
#include "common.cuh"
#include <math_constants.h>
#include <stdlib.h>

#ifdef GG_TILED_PROFILE
__device__ unsigned long long g_tiled_prof[8];   // [0] consumer wait, [1] consumer busy, [2] producer wait, [3] producer issue, [4] consumer warps, [5] producer warps, [6] clocks and [7] ns of consumer warp 0 of CTA 0 (SM clock of the launch)
__device__ unsigned long long g_clk_ring[256];     // (clocks, ns) of the last 128 launches, in launch order
__device__ unsigned int g_clk_n;
__device__ __forceinline__ unsigned long long gg_globaltimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define GG_PROF_T0() long long prof_t = clock64()
#define GG_PROF_ADD(var) do { const long long now_ = clock64(); var += now_ - prof_t; prof_t = now_; } while (0)
#else
#define GG_PROF_T0() do {} while (0)
#define GG_PROF_ADD(var) do {} while (0)
#endif

#ifndef GG_PRODUCER_SLEEP_NS
#define GG_PRODUCER_SLEEP_NS 250u
#endif
#ifndef GG_ENC_ROW_PAD
#define GG_ENC_ROW_PAD 16u
#endif

namespace {

constexpr int NP = 4;                 // producer warps (exactly one warpgroup: setmaxnreg is per warpgroup)
constexpr int NC = 12;                // consumer warps
constexpr int kThreads = 32 * (NP + NC);
constexpr int CH = 3;                 // edges per softmax chunk (joints have exactly 3 in-edges)
struct FastTag { static constexpr bool value = true; };
struct SlowTag { static constexpr bool value = false; };

struct TiledSeg {                                    // one edge type (all gates of the cell)
    const float* P_src; int ld_src, k_off;
    const float* P_dst; int ld_dst, q_off;          // target block at q_off: Q | QX (or Q') | position
    const int* rowptr; const int* col; const float* ea; const int* wrap;
    const int* nz; const int* nzptr; const int4* tiles; const int* cta_ptr;
    const float* Wv3;
    int n_dst, n_edges;
    float* agg; int ld_agg; float* ea_out;
};
constexpr int kMaxSeg = 3;
// One launch serves up to kMaxSeg edge types of a cell (same width, gates and row form, hence the same stage layout): every
// persistent CTA walks its tiles of segment 0, then of segment 1, ... without draining the stage ring in between, so a cell
// pays one launch ramp and one tail instead of three.
struct TiledParams {
    TiledSeg seg[kMaxSeg];
    int n_seg;
    float inv_sqrt_c;
};

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
__device__ __forceinline__ void mbar_arrive_(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
// Producer-side wait: a waiting producer warp must not poll — its polls take issue slots from the consumer warps of the same
// scheduler (ncu round 2: the suspend-time hint of try_wait returns after ~50 ns, 32 polls per tile and warp, 12 % of all
// instructions issued).  The producers run NS - 1 stages ahead, so they can afford to sleep between polls.
__device__ __forceinline__ void mbar_wait_sleepy_(uint32_t bar, uint32_t parity) {
    uint32_t done, spins = 0;
    long long t0 = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity), "r"(100000u) : "memory");
        if (!done) {
            asm volatile("nanosleep.u32 %0;" ::"r"(GG_PRODUCER_SLEEP_NS));
            if ((++spins & 0xffu) == 0) {                    // a lost arrival must surface as an error, not as a hung GPU
                if (t0 == 0) t0 = clock64();
                else if (clock64() - t0 > 8000000000LL) __trap();
            }
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
__device__ __forceinline__ float rcp_nr(float x) {            // 1 / x for x in [1e-16, 1e3]: approximate reciprocal + one Newton step (<= 1 ulp)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return fmaf(r, fmaf(-x, r, 1.0f), r);
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
// packed fp32 pairs (sm_100 FFMA2 / FMUL2): a float4 travels as two 64-bit registers
typedef unsigned long long u64;
struct P4 { u64 lo, hi; };
__device__ __forceinline__ u64 pack2(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fmax2(u64 a, u64 b) {           // no packed max on sm_100: two FMNMX on the halves
    float ax, ay, bx, by;
    unpack2(a, ax, ay); unpack2(b, bx, by);
    return pack2(fmaxf(ax, bx), fmaxf(ay, by));
}
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float hsum2(u64 a, u64 b) { float x, y; unpack2(fadd2(a, b), x, y); return x + y; }
__device__ __forceinline__ P4 lds4p(uint32_t addr) {
    P4 r;
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(r.lo), "=l"(r.hi) : "r"(addr));
    return r;
}
__device__ __forceinline__ P4 ldg4p(const float* p) { const float4 v = ldg4(p); P4 r; r.lo = pack2(v.x, v.y); r.hi = pack2(v.z, v.w); return r; }
__device__ __forceinline__ void lds2(uint32_t addr, int& a, int& b) { asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(addr)); }
__device__ __forceinline__ void sts2(uint32_t addr, int a, int b) { asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory"); }
__device__ __forceinline__ void stg4p_stream(float* p, u64 a, u64 b) {
    asm volatile("st.global.cs.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ float wrapv(int code) { return code == 0 ? 0.f : (code == 1 ? 1.f : -1.f); }   // gg_edge_wrap: 1 -> +1, 2 -> -1

// Tile geometry and shared-memory layout of one stage (byte offsets; every block 16-byte aligned), fixed at compile time per
// (width, gates, mode) so that every address in the consumer loop is a constant offset.
//   rows [ECAP x ES] | headers [HCAP x HB] | target descriptors [ECAP x 8] | edge metadata [ECAP x 8] | info [16]
// rawk: floats of the raw block in front of V (0: K | V form; 16: encoder; 32 + C: cells with hidden state)
constexpr uint32_t cfg_es(int G, int C, int rawk) { return rawk ? 4u * rawk + (uint32_t)G * C * 4u : 2u * (uint32_t)G * C * 4u; }   // [raw | V] or [K | V]
constexpr uint32_t cfg_qb(int G, int C, int rawk) { return rawk ? 4u * rawk * G : (uint32_t)G * C * 4u + 16u * G; }            // Q' (rawk per gate) or Q | QX
// target block: Q | QX | position (x, y, z, -), or Q' alone — there the position rides in three spare slots of the first gate's Q'
constexpr uint32_t cfg_hb(int G, int C, int rawk) { return cfg_qb(G, C, rawk) + (rawk ? 0u : 16u); }
// Stride of a staged row.  Encoder form (16 raw floats in front): every lane of a gate group scores ANOTHER edge of the chunk, so the
// lanes of one 128-bit shared load read the same 16 bytes of three different rows; with the natural stride (304 words = 16 mod 32)
// rows an even number of slots apart sit in the same banks (ncu round 1: 1.3 conflicts per shared load).  16 bytes of padding make
// the stride 308 words = 20 mod 32: rows up to 7 slots apart land in 8 different bank quads.
constexpr uint32_t cfg_ess(int G, int C, int rawk) { return cfg_es(G, C, rawk) + (rawk == 16 ? GG_ENC_ROW_PAD : 0u); }
constexpr uint32_t cfg_stage_bytes(int G, int C, int raw, int ecap, int hcap) {
    const uint32_t e8 = (((uint32_t)ecap * 8u) + 15u) & ~15u;
    return ((uint32_t)ecap * cfg_ess(G, C, raw) + (uint32_t)hcap * cfg_hb(G, C, raw) + 2u * e8 + 16u + 127u) & ~127u;
}
constexpr int cfg_hcap(int ecap) { return ecap / 3 + 2; }
// The largest tile (multiple of 6 edges: joints have 3 in-edges, grains ~6) that fits 227 KB with three stages, else with two
// (measured on the bench graph, encoder form: 54 edges x 3 stages 179 / 152 / 179 us per launch, 36 x 4: 192 / 164 / 192 us).
// what = 0: ECAP, 1: stages
#ifndef GG_DEC_ECAP
#define GG_DEC_ECAP 0      // (measurement override of the decoder form's tile: edges per tile, stages)
#define GG_DEC_NS 0
#endif
constexpr int cfg_pick(int G, int C, int raw, int what) {
    if (GG_DEC_ECAP && raw > 16) return what == 0 ? GG_DEC_ECAP : GG_DEC_NS;
    for (int ns = 3; ns >= 2; --ns)
        for (int ecap = 60; ecap >= 12; ecap -= 6)
            if ((long long)ns * cfg_stage_bytes(G, C, raw, ecap, cfg_hcap(ecap)) <= 227 * 1024 - 256) return what == 0 ? ecap : ns;
    return 0;
}
// MODE: 0 = K | V rows, 1 = raw scores without hidden state (16-float raw block), 2 = raw scores on [X padded to 32 | h]
__host__ __device__ constexpr int mode_rawk(int mode, int C) { return mode == 0 ? 0 : (mode == 1 ? 16 : 32 + C); }
template <int NV, int G_, int MODE>
struct TCfg {
    static constexpr int C = 32 * NV, G = G_, RAW = mode_rawk(MODE, 32 * NV);
    static constexpr int ECAP = cfg_pick(G, C, RAW, 0), NS = cfg_pick(G, C, RAW, 1), HCAP = cfg_hcap(ECAP);
    static constexpr uint32_t ES = cfg_es(G, C, RAW), ESS = cfg_ess(G, C, RAW), QB = cfg_qb(G, C, RAW), HB = cfg_hb(G, C, RAW);
    static constexpr uint32_t POS = RAW == 0 ? QB : (RAW == 16 ? 48u : 112u);   // byte offset of x, y, z inside the target block
    static constexpr uint32_t HDR = (uint32_t)ECAP * ESS;
    static constexpr uint32_t TD = HDR + (uint32_t)HCAP * HB;            // {node, lo | hi << 8 | starts << 16 | ends << 17 | header slot << 24}
    static constexpr uint32_t E8 = (((uint32_t)ECAP * 8u) + 15u) & ~15u;
    static constexpr uint32_t EM = TD + E8;                              // {edge length, wrap code}
    static constexpr uint32_t INFO = EM + E8;                            // {targets, index of the first one in nz, -, -}
    static constexpr uint32_t BYTES = cfg_stage_bytes(G, C, RAW, ECAP, HCAP);
    static_assert(ECAP >= 12 && ECAP <= 60 && NS >= 2, "tile does not fit");
    static_assert(BYTES == ((INFO + 16u + 127u) & ~127u), "layout");
};
// shapes with a compiled kernel: the encoder (3 live gates, raw scores), the decoder (4 gates), single convolutions (1 gate)
constexpr bool cfg_supported(int G, int mode) { return mode == 1 ? G == 3 : (mode == 2 ? G == 4 : (G == 4 || G == 1)); }

template <int NV, int G, int MODE>
__global__ void __launch_bounds__(kThreads, 1)
pgat_gather_tiled_kernel(const __grid_constant__ TiledParams P) {
    using K = TCfg<NV, G, MODE>;
    constexpr bool RAW = MODE != 0, RAWH = MODE == 2;        // raw scores; ... on the full cell input (with hidden state)
    constexpr int RAWK = K::RAW;
    constexpr int C = 32 * NV, GC = G * C;
    constexpr int NQ = MODE == 1 ? 4 : (RAWH ? NV + 1 : NV); // float4 registers of the target's query (raw: this lane's share of Q')
    constexpr float LOG2E = 1.4426950408889634f;
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int ECAP = K::ECAP, HCAP = K::HCAP, NS = K::NS;
    (void)ECAP;
    uint32_t smem0 = smem_addr(smem);
    asm volatile("mov.u32 %0, %0;" : "+r"(smem0));            // opaque: ptxas otherwise re-derives the shared window base (S2UR SR_CgaCtaId,
                                                              // ~30 clk on the critical path) in front of every group of shared loads
    const uint32_t bar0 = smem0 + (uint32_t)NS * K::BYTES;    // full[NS] | empty[NS]
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (NS + s); };
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init_(full_bar(s), NP); mbar_init_(empty_bar(s), NC); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int grid = gridDim.x;

    if (warp < NP) {
        // =========================================================================================== producers
        // All NP producer warps work on EVERY tile, in order (so none can run more than one stage ring ahead: a parity wait
        // cannot tell "two phases behind" from "done"): slot s / target j of a tile belongs to warp s mod NP, lane s / NP, which
        // spreads the warp-serial issue of the bulk copies (~50 clk each) over NP warps.  Tile descriptors are read two tiles
        // ahead and the per-edge / per-target metadata one tile ahead, so the chain tiles -> col -> row address is never waited for.
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");          // registers go to the consumer warpgroups
        struct Meta { int col, wrap, tn, ta, tb, t0, colA, colB; float ea; };
        const TiledSeg* p = &P.seg[0];
        int f0 = 0, n_tiles = 0;
        const int mine = warp + NP * lane;                           // the slot and the target this lane serves in every tile
        auto load_desc = [&](int t) -> int4 {                        // {first edge, edges, first target, last target}
            return t < n_tiles ? __ldg(&p->tiles[f0 + t]) : make_int4(0, 0, 0, -1);
        };
        auto load_meta = [&](const int4& d) -> Meta {
            Meta m;
            const int e0 = d.x, ne = d.y, cnt = d.w - d.z + 1;
            m.col = 0; m.wrap = 0; m.ea = 0.f; m.tn = 0; m.ta = 0; m.tb = 0; m.t0 = 0;
            m.colA = lane < ne ? __ldg(&p->col[e0 + lane]) : -1 - lane;               // every slot's source, for the duplicate search
            m.colB = (ECAP > 32 && 32 + lane < ne) ? __ldg(&p->col[e0 + 32 + lane]) : -33 - lane;
            if (mine < ne) { m.col = __ldg(&p->col[e0 + mine]); m.ea = __ldg(&p->ea[e0 + mine]); m.wrap = __ldg(&p->wrap[e0 + mine]); }
            if (mine < cnt) { m.tn = __ldg(&p->nz[d.z + mine]); m.ta = __ldg(&p->nzptr[d.z + mine]); m.tb = __ldg(&p->nzptr[d.z + mine + 1]); }
            if (cnt > 0) m.t0 = __ldg(&p->nzptr[d.z]);
            return m;
        };
        int stage = 0; uint32_t phase = 0;
        long long prof_wait = 0, prof_busy = 0;
        (void)prof_wait; (void)prof_busy;
        GG_PROF_T0();
        for (int sg = 0; sg < P.n_seg; ++sg) {
        p = &P.seg[sg];
        f0 = __ldg(&p->cta_ptr[blockIdx.x]); n_tiles = __ldg(&p->cta_ptr[blockIdx.x + 1]) - f0;   // this CTA's tiles of the segment
        int4 d_cur = load_desc(0), d_nxt = load_desc(1);
        Meta m_cur = load_meta(d_cur);
        for (int t = 0; t < n_tiles; ++t) {
            const int4 d_n2 = load_desc(t + 2);
            const Meta m_nxt = load_meta(d_nxt);
            const int e0 = d_cur.x, ne = d_cur.y, e1 = e0 + ne;
            const int cnt = d_cur.w - d_cur.z + 1;
            const int fs = m_cur.t0 >= e0 ? 1 : 0;                           // does the first target START in this tile?
            const uint32_t base = smem0 + (uint32_t)stage * K::BYTES, bar = full_bar(stage);
            // what the consumer needs to know about target `mine`
            const int starts = m_cur.ta >= e0 ? 1 : 0, ends = m_cur.tb <= e1 ? 1 : 0;
            const int h = mine - (1 - fs);
            const bool has_hdr = mine < cnt && starts && h < HCAP;           // else: continuing row, or more starts than header slots
            const int n_hdr = __popc(__ballot_sync(0xffffffffu, has_hdr));
            // Neighbouring targets share sources (grain -> joint: ~48 % of a tile's edges repeat a source of the same tile, the other
            // types 20-25 %): within each group of 32 slots a source row is staged ONCE, at the slot of its first edge, and every
            // edge carries the slot its row lives in (bits 8.. of the wrap word).  This form is bound by L2 -> shared-memory bytes.
            const int leadA = __ffs(__match_any_sync(0xffffffffu, m_cur.colA)) - 1;
            const int leadB = ECAP > 32 ? 32 + __ffs(__match_any_sync(0xffffffffu, m_cur.colB)) - 1 : 0;
            const int la = __shfl_sync(0xffffffffu, leadA, mine & 31), lb = __shfl_sync(0xffffffffu, leadB, mine & 31);
            const int row_slot = (ECAP > 32 && mine >= 32) ? lb : la;         // where the row of edge `mine` is staged
            const bool copies_row = mine < ne && row_slot == mine;
            const int n_rows = __popc(__ballot_sync(0xffffffffu, copies_row));
            GG_PROF_ADD(prof_busy);
            mbar_wait_sleepy_(empty_bar(stage), phase ^ 1u);
            GG_PROF_ADD(prof_wait);
            if (mine < ne) sts2(base + K::EM + 8u * mine, __float_as_int(m_cur.ea), m_cur.wrap | ((row_slot * (int)K::ESS) << 8));   // wrap code | byte offset of the staged row
            if (mine < cnt)
                sts2(base + K::TD + 8u * mine, m_cur.tn,
                     (max(m_cur.ta, e0) - e0) | ((min(m_cur.tb, e1) - e0) << 8) | (starts << 16) | (ends << 17) | ((has_hdr ? h : 255) << 24));
            if (threadIdx.x == 0) sts2(base + K::INFO, cnt, d_cur.z);
            __syncwarp();
            if (lane == 0) mbar_expect_tx_(bar, (uint32_t)n_rows * K::ES + (uint32_t)n_hdr * K::HB);
            __syncwarp();
            if (copies_row) bulk_g2s(base + (uint32_t)mine * K::ESS, p->P_src + (size_t)m_cur.col * p->ld_src + p->k_off, K::ES, bar);
            if (has_hdr) {
                const uint32_t hd = base + K::HDR + (uint32_t)h * K::HB;
                bulk_g2s(hd, p->P_dst + (size_t)m_cur.tn * p->ld_dst + p->q_off, K::HB, bar);     // Q | QX (Q') and the position behind it
            }
            d_cur = d_nxt; d_nxt = d_n2; m_cur = m_nxt;
            if (++stage == NS) { stage = 0; phase ^= 1u; }
        }
        }
#ifdef GG_TILED_PROFILE
        GG_PROF_ADD(prof_busy);
        if (lane == 0) { atomicAdd(&g_tiled_prof[2], (unsigned long long)prof_wait); atomicAdd(&g_tiled_prof[3], (unsigned long long)prof_busy); atomicAdd(&g_tiled_prof[5], 1ull); }
#endif
        return;
    }

    // =============================================================================================== consumers
    asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");             // 4 x 32 x 56 + 12 x 32 x 152 = 64 K registers
    const int cw = warp - NP;
    const int grp = lane >> 3, sub = lane & 7;
    const bool active = grp < G;
    const int gsel = active ? grp : 0;                       // idle 8-lane groups (G < 4) shadow gate 0 and store nothing
    const uint32_t lane_off = 4u * (gsel * C + 4 * sub);     // byte offset of this lane's first float4 inside a staged row
    const uint32_t v_off = RAW ? 4u * RAWK : (uint32_t)GC * 4u;   // V row behind the raw block / the K row
    const float sc2 = P.inv_sqrt_c * LOG2E;                  // scores are kept in log2 units: exp(x) = ex2(x log2 e)
    const int me = sub % 3;                                  // raw-score mode: the edge of a chunk this lane scores (lanes 0..2 of a group publish)
    const int src0 = lane & 24;                              // lane `sub == 0` of this gate group

    P4 wvx[NV], wvy[NV], wvz[NV];                            // x, y, z columns of lin_value for this lane's channels (per segment)

    // per-target state (survives tile boundaries)
    P4 q[NQ], vp[NV], acc[NV];
    float4 qx = make_float4(0.f, 0.f, 0.f, 0.f);
    float m_run = -CUDART_INF_F, l_run = 0.f, ea_acc = 0.f;
#pragma unroll
    for (int r = 0; r < NV; ++r) { vp[r].lo = vp[r].hi = 0ull; acc[r].lo = acc[r].hi = 0ull; }
#pragma unroll
    for (int r = 0; r < NQ; ++r) q[r].lo = q[r].hi = 0ull;

    int stage = 0; uint32_t phase = 0;
    long long prof_wait = 0, prof_busy = 0;
    (void)prof_wait; (void)prof_busy;
    GG_PROF_T0();
#ifdef GG_TILED_PROFILE
    const long long prof_c0 = clock64();
    const unsigned long long prof_n0 = gg_globaltimer();
#endif
    for (int sg = 0; sg < P.n_seg; ++sg) {
    const TiledSeg& ps = P.seg[sg];
    const int n_tiles = __ldg(&ps.cta_ptr[blockIdx.x + 1]) - __ldg(&ps.cta_ptr[blockIdx.x]);   // this CTA's tiles of the segment
    // What the per-target code reads of the segment lives in registers: indexed constant-bank loads (c[0x0][R + off], ~30 clk)
    // would otherwise sit on the critical path of every target (measured: +8 % on the launch).
    struct { const float* P_dst; float* agg; float* ea_out; const int* rowptr; const float* Wv3; int n_dst, ld_dst, q_off, ld_agg; } p;
    p.P_dst = ps.P_dst; p.agg = ps.agg; p.ea_out = ps.ea_out; p.rowptr = ps.rowptr; p.Wv3 = ps.Wv3;
    p.n_dst = ps.n_dst; p.ld_dst = ps.ld_dst; p.q_off = ps.q_off; p.ld_agg = ps.ld_agg;
    asm volatile("" : "+l"(p.P_dst), "+l"(p.agg), "+l"(p.ea_out), "+r"(p.n_dst), "+r"(p.ld_dst), "+r"(p.q_off), "+r"(p.ld_agg));
    // x, y, z columns of lin_value for this lane's channels (zero for idle groups): Wv3 is [G*C][4]; packed pairs
#pragma unroll
    for (int r = 0; r < NV; ++r) {
        float4 t4[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) t4[i] = active ? ldg4(p.Wv3 + (size_t)(grp * C + 4 * (sub + 8 * r) + i) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        wvx[r].lo = pack2(t4[0].x, t4[1].x); wvx[r].hi = pack2(t4[2].x, t4[3].x);
        wvy[r].lo = pack2(t4[0].y, t4[1].y); wvy[r].hi = pack2(t4[2].y, t4[3].y);
        wvz[r].lo = pack2(t4[0].z, t4[1].z); wvz[r].hi = pack2(t4[2].z, t4[3].z);
    }
    for (int t = 0; t < n_tiles; ++t) {
        const uint32_t base = smem0 + (uint32_t)stage * K::BYTES;
        GG_PROF_ADD(prof_busy);
        mbar_wait_(full_bar(stage), phase);
        GG_PROF_ADD(prof_wait);
        int cnt, i_first;
        lds2(base + K::INFO, cnt, i_first);
        int j = cw - i_first % NC;
        if (j < 0) j += NC;
        for (; j < cnt; j += NC) {
            int node, td;
            lds2(base + K::TD + 8u * j, node, td);
            if (node >= p.n_dst) continue;                   // rows behind the owned ones (slab partition) are not computed
            const int lo = td & 0xff, hi = (td >> 8) & 0xff;
            if (td & 0x10000) {
                // ---- the target starts here: its query, position, fresh softmax state
                const int hs = (td >> 24) & 0xff;
                float4 pi;
                if (hs != 255) {
                    const uint32_t hd = base + K::HDR + (uint32_t)hs * K::HB;
                    if (RAWH) {                               // this lane's share of Q'_gate: float4 slots sub, sub + 8, ...
#pragma unroll
                        for (int r = 0; r < NQ; ++r) q[r] = lds4p(hd + gsel * (4 * RAWK) + 16 * (sub + 8 * r));
                        qx = lds4(hd + gsel * (4 * RAWK));    // Q'[0:3] = Wk3^T q
                    } else if (RAW) {
#pragma unroll
                        for (int r = 0; r < NQ; ++r) q[r] = lds4p(hd + gsel * 64 + 16 * r);
                    } else {
#pragma unroll
                        for (int r = 0; r < NQ; ++r) q[r] = lds4p(hd + lane_off + 128 * r);
                        qx = lds4(hd + (uint32_t)GC * 4u + 16u * gsel);
                    }
                    pi = lds4(hd + K::POS);
                } else {                                      // more starting targets than header slots (runs of in-degree < 3): plain loads
                    const float* qrow = p.P_dst + (size_t)node * p.ld_dst + p.q_off;
                    if (RAWH) {
#pragma unroll
                        for (int r = 0; r < NQ; ++r) q[r] = ldg4p(qrow + gsel * RAWK + 4 * (sub + 8 * r));
                        qx = ldg4(qrow + gsel * RAWK);
                    } else if (RAW) {
#pragma unroll
                        for (int r = 0; r < NQ; ++r) q[r] = ldg4p(qrow + gsel * 16 + 4 * r);
                    } else {
#pragma unroll
                        for (int r = 0; r < NQ; ++r) q[r] = ldg4p(qrow + gsel * C + 4 * (sub + 8 * r));
                        qx = ldg4(qrow + GC + 4 * gsel);
                    }
                    pi = ldg4(qrow + K::POS / 4);
                }
                if (MODE == 1) { unpack2(q[0].lo, qx.x, qx.y); float dummy; unpack2(q[0].hi, qx.z, dummy); }   // Q'[0:3] = Wk3^T q
                const u64 px = pack2(pi.x, pi.x), py = pack2(pi.y, pi.y), pz = pack2(pi.z, pi.z);
#pragma unroll
                for (int r = 0; r < NV; ++r) {                // vp = Wv3 p_i:  V_j + Wv3 (w_e - p_i) > 0  <=>  V_j + Wv3 w_e > vp
                    vp[r].lo = ffma2(wvz[r].lo, pz, ffma2(wvy[r].lo, py, fmul2(wvx[r].lo, px)));
                    vp[r].hi = ffma2(wvz[r].hi, pz, ffma2(wvy[r].hi, py, fmul2(wvx[r].hi, px)));
                }
            }
            // One chunk of <= CH consecutive in-edges of the target.  FAST = the chunk is full and none of its edges crosses a patch
            // boundary (decided warp-uniformly from the staged edge words): straight-line code, no clamped slots, no per-edge
            // predicates, no wrap corrections — the compiler schedules the loads of all three edges ahead of the arithmetic.
            // WHOLE = the chunk is the target's whole row (it starts and ends in this tile with exactly CH edges — every joint): the
            // running state is written, never read (no reset, no rescale, the first edge multiplies instead of accumulating).
            auto chunk = [&](auto fast_tag, auto whole_tag, const int (&ai)[CH], const int (&ww)[CH], const int n_e) {
                constexpr bool FAST = decltype(fast_tag)::value, WHOLE = decltype(whole_tag)::value;
                float ae[CH], sc[CH];
                int wc[CH];
                uint32_t ro[CH];                              // shared-memory address of the edge's source row (shared by duplicates)
#pragma unroll
                for (int e = 0; e < CH; ++e) { ae[e] = __int_as_float(ai[e]); wc[e] = FAST ? 0 : (ww[e] & 0xff); ro[e] = base + ((uint32_t)ww[e] >> 8); }
                if (RAWH) {
                    // x_j . Q'_i over the 32 + C input slots: lane `sub` of a gate group owns the float4 slots sub, sub + 8, ...;
                    // the four gate groups read the same addresses (broadcast).  Slot 31 (We . q) meets the edge length.
#pragma unroll
                    for (int e = 0; e < CH; ++e) {
                        const uint32_t xrow = ro[e] + 16u * sub;
                        u64 da = 0ull, db = 0ull;
#pragma unroll
                        for (int r = 0; r < NQ; ++r) {
                            P4 xx = lds4p(xrow + 128 * r);
                            if (r == 0 && sub == 7) { float x30, x31; unpack2(xx.hi, x30, x31); xx.hi = pack2(x30, ae[e]); }
                            if (r == 0) { da = fmul2(q[r].lo, xx.lo); db = fmul2(q[r].hi, xx.hi); }
                            else { da = ffma2(q[r].lo, xx.lo, da); db = ffma2(q[r].hi, xx.hi, db); }
                        }
                        float dd = group_sum8(hsum2(da, db));
                        if (!FAST && wc[e]) {
                            dd = fmaf(qx.x, wrapv(wc[e] & 3), dd); dd = fmaf(qx.y, wrapv((wc[e] >> 2) & 3), dd); dd = fmaf(qx.z, wrapv((wc[e] >> 4) & 3), dd);
                        }
                        sc[e] = dd * sc2;
                    }
                } else if (RAW) {
                    // each lane scores ONE edge of the chunk (edge `me`) for its gate; lanes 0..2 of the group publish
                    const float my_ae = me == 0 ? ae[0] : (me == 1 ? ae[1] : ae[2]);
                    const uint32_t row = me == 0 ? ro[0] : (me == 1 ? ro[1] : ro[2]);
                    const P4 x0 = lds4p(row), x1 = lds4p(row + 16), x2 = lds4p(row + 32);
                    P4 x3 = lds4p(row + 48);
                    { float x14, x15; unpack2(x3.hi, x14, x15); x3.hi = pack2(x14, my_ae); }    // Q'[15] = We . q multiplies the edge length
                    u64 da = fmul2(q[0].lo, x0.lo), db = fmul2(q[0].hi, x0.hi);
                    da = ffma2(q[1].lo, x1.lo, da); db = ffma2(q[1].hi, x1.hi, db);
                    da = ffma2(q[2].lo, x2.lo, da); db = ffma2(q[2].hi, x2.hi, db);
                    da = ffma2(q[3].lo, x3.lo, da); db = ffma2(q[3].hi, x3.hi, db);
                    float dd = hsum2(da, db);
                    if (!FAST && (wc[0] | wc[1] | wc[2])) {   // periodGATconv.py:209-211: the wrapped displacement enters the key
                        const int my_wc = me == 0 ? wc[0] : (me == 1 ? wc[1] : wc[2]);
                        dd = fmaf(qx.x, wrapv(my_wc & 3), dd); dd = fmaf(qx.y, wrapv((my_wc >> 2) & 3), dd); dd = fmaf(qx.z, wrapv((my_wc >> 4) & 3), dd);
                    }
                    dd *= sc2;
#pragma unroll
                    for (int e = 0; e < CH; ++e) sc[e] = __shfl_sync(0xffffffffu, dd, src0 + e);
                } else {
#pragma unroll
                    for (int e = 0; e < CH; ++e) {
                        const uint32_t krow = ro[e] + lane_off;
                        u64 da = 0ull, db = 0ull;
#pragma unroll
                        for (int r = 0; r < NV; ++r) {
                            const P4 kk = lds4p(krow + 128 * r);
                            da = ffma2(q[r].lo, kk.lo, da); db = ffma2(q[r].hi, kk.hi, db);
                        }
                        float dd = group_sum8(hsum2(da, db));
                        dd = fmaf(qx.w, ae[e], dd);
                        if (!FAST && wc[e]) {
                            dd = fmaf(qx.x, wrapv(wc[e] & 3), dd); dd = fmaf(qx.y, wrapv((wc[e] >> 2) & 3), dd); dd = fmaf(qx.z, wrapv((wc[e] >> 4) & 3), dd);
                        }
                        sc[e] = dd * sc2;
                    }
                }
                if (!FAST) {
                    if (n_e < 2) sc[1] = -CUDART_INF_F;
                    if (n_e < 3) sc[2] = -CUDART_INF_F;
                }
                if (WHOLE) { m_run = fmaxf(sc[0], fmaxf(sc[1], sc[2])); l_run = 0.f; ea_acc = 0.f; }
                const float m_new = WHOLE ? m_run : fmaxf(fmaxf(m_run, sc[0]), fmaxf(sc[1], sc[2]));
                if (!WHOLE && m_new > m_run) {                // online softmax: rescale what earlier chunks accumulated
                    if (m_run != -CUDART_INF_F) {
                        const float scale = ex2_approx(m_run - m_new);
                        const u64 s2 = pack2(scale, scale);
                        l_run *= scale; ea_acc *= scale;
#pragma unroll
                        for (int r = 0; r < NV; ++r) { acc[r].lo = fmul2(acc[r].lo, s2); acc[r].hi = fmul2(acc[r].hi, s2); }
                    }
                    m_run = m_new;
                }
                // relu(v + Wv3 (w_e - p_i)) = max(V_j + Wv3 w_e, vp) - vp: accumulate pe * max(., vp); vp * sum(pe) is subtracted once
#pragma unroll
                for (int e = 0; e < CH; ++e) {
                    if (FAST || e < n_e) {
                        const float pe = ex2_approx(sc[e] - m_run);
                        const u64 pe2 = pack2(pe, pe);
                        l_run += pe;
                        ea_acc = fmaf(pe, ae[e], ea_acc);
                        const uint32_t vrow = ro[e] + v_off + lane_off;
                        P4 v[NV];
#pragma unroll
                        for (int r = 0; r < NV; ++r) v[r] = lds4p(vrow + 128 * r);
                        if (!FAST && wc[e]) {                 // edge crosses a periodic / patch boundary (warp-uniform)
                            const float fx = wrapv(wc[e] & 3), fy = wrapv((wc[e] >> 2) & 3), fz = wrapv((wc[e] >> 4) & 3);
                            const u64 tx = pack2(fx, fx), ty = pack2(fy, fy), tz = pack2(fz, fz);
#pragma unroll
                            for (int r = 0; r < NV; ++r) {
                                v[r].lo = ffma2(wvz[r].lo, tz, ffma2(wvy[r].lo, ty, ffma2(wvx[r].lo, tx, v[r].lo)));
                                v[r].hi = ffma2(wvz[r].hi, tz, ffma2(wvy[r].hi, ty, ffma2(wvx[r].hi, tx, v[r].hi)));
                            }
                        }
#pragma unroll
                        for (int r = 0; r < NV; ++r) {
                            if (WHOLE && e == 0) {
                                acc[r].lo = fmul2(pe2, fmax2(v[r].lo, vp[r].lo));
                                acc[r].hi = fmul2(pe2, fmax2(v[r].hi, vp[r].hi));
                            } else {
                                acc[r].lo = ffma2(pe2, fmax2(v[r].lo, vp[r].lo), acc[r].lo);
                                acc[r].hi = ffma2(pe2, fmax2(v[r].hi, vp[r].hi), acc[r].hi);
                            }
                        }
                    }
                }
            };
            bool whole = false;
            if ((td & 0x30000) == 0x30000 && hi - lo == CH) { // the whole row in one full chunk
                int ai[CH], ww[CH];
                const uint32_t em = base + K::EM + 8u * lo;
#pragma unroll
                for (int e = 0; e < CH; ++e) lds2(em + 8u * e, ai[e], ww[e]);
                if (((ww[0] | ww[1] | ww[2]) & 0xff) == 0) { chunk(FastTag{}, FastTag{}, ai, ww, CH); whole = true; }
            }
            if (!whole && (td & 0x10000)) {                   // fresh softmax state of a row that is walked chunk by chunk
#pragma unroll
                for (int r = 0; r < NV; ++r) acc[r].lo = acc[r].hi = 0ull;
                m_run = -CUDART_INF_F; l_run = 0.f; ea_acc = 0.f;
            }
            for (int s0 = whole ? hi : lo; s0 < hi; s0 += CH) {
                const int n_e = hi - s0;                      // >= 1, warp-uniform
                int ai[CH], ww[CH];
                const uint32_t em = base + K::EM + 8u * s0;
                if (n_e >= CH) {
#pragma unroll
                    for (int e = 0; e < CH; ++e) lds2(em + 8u * e, ai[e], ww[e]);
                    if (((ww[0] | ww[1] | ww[2]) & 0xff) == 0) { chunk(FastTag{}, SlowTag{}, ai, ww, CH); continue; }
                } else {                                      // slots beyond the row are clamped to its last edge
                    lds2(em, ai[0], ww[0]);
                    lds2(em + (n_e > 1 ? 8u : 0u), ai[1], ww[1]);
                    ai[2] = ai[1]; ww[2] = ww[1];
                }
                chunk(SlowTag{}, SlowTag{}, ai, ww, n_e);
            }
            if (td & 0x20000) {
                // ---- the target ends here: normalise and store (PyG softmax: exp(s - max) / (sum + 1e-16))
                const float inv = rcp_nr(l_run + 1e-16f);
                if (active) {
                    float* orow = p.agg + (size_t)node * p.ld_agg + grp * C + 4 * sub;
                    const u64 nl2 = pack2(-l_run, -l_run), inv2 = pack2(inv, inv);
#pragma unroll
                    for (int r = 0; r < NV; ++r)              // sum pe * relu(.) = sum pe * max(., vp) - vp * sum pe
                        stg4p_stream(orow + 32 * r, fmul2(ffma2(vp[r].lo, nl2, acc[r].lo), inv2), fmul2(ffma2(vp[r].hi, nl2, acc[r].hi), inv2));
                    if (sub == 0) p.ea_out[(size_t)node * G + grp] = ea_acc * inv;
                }
            }
        }
        __syncwarp();                                         // every lane is done with the stage before it is handed back
        if (lane == 0) mbar_arrive_(empty_bar(stage));
        if (++stage == NS) { stage = 0; phase ^= 1u; }
    }

    // ---- targets without in-edges: zero rows (PyG scatter-add leaves them 0)
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t b0 = ((int64_t)blockIdx.x * NC + cw) * 32; b0 < p.n_dst; b0 += (int64_t)grid * NC * 32) {
        const int n = (int)b0 + lane;
        const bool empty = n < p.n_dst && __ldg(&p.rowptr[n + 1]) == __ldg(&p.rowptr[n]);
        unsigned mask = __ballot_sync(0xffffffffu, empty);
        while (mask) {
            const int bit = __ffs(mask) - 1;
            mask &= mask - 1;
            const int node = (int)b0 + bit;
            if (active) {
                float* orow = p.agg + (size_t)node * p.ld_agg + grp * C + 4 * sub;
#pragma unroll
                for (int r = 0; r < NV; ++r) stg4_stream(orow + 32 * r, z4);
                if (sub == 0) p.ea_out[(size_t)node * G + grp] = 0.f;
            }
        }
    }
    }   // segments
#ifdef GG_TILED_PROFILE
    GG_PROF_ADD(prof_busy);
    if (lane == 0) { atomicAdd(&g_tiled_prof[0], (unsigned long long)prof_wait); atomicAdd(&g_tiled_prof[1], (unsigned long long)prof_busy); atomicAdd(&g_tiled_prof[4], 1ull); }
    if (blockIdx.x == 0 && cw == 0 && lane == 0) {
        const unsigned long long dc = (unsigned long long)(clock64() - prof_c0), dn = gg_globaltimer() - prof_n0;
        g_tiled_prof[6] = dc; g_tiled_prof[7] = dn;
        const unsigned int slot = atomicAdd(&g_clk_n, 1u) & 127u;
        g_clk_ring[2 * slot] = dc; g_clk_ring[2 * slot + 1] = dn;
    }
#endif
}

int host_mode(int C, int raw_k) { return raw_k == 0 ? 0 : (raw_k == 16 ? 1 : (raw_k == 32 + C ? 2 : -1)); }
int host_ecap(int G, int C, int raw_k) {
    const int mode = host_mode(C, raw_k);
    if (mode < 0 || G < 1 || G > 4 || C % 32 || C < 32 || C > 128 || !cfg_supported(G, mode)) return 0;
    return cfg_pick(G, C, raw_k, 0);
}

}  // namespace

extern "C" int gg_gather_tile_ecap(int32_t G, int32_t C, int32_t raw_k) {
    return host_ecap(G, C, raw_k);
}

// number of persistent CTAs the tile lists are laid out for: one per SM (GG_GATHER_SMS caps it for experiments)
extern "C" int gg_gather_ctas(void) {
    static int n = 0;
    if (n == 0) {
        int dev = 0, sms = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) return 0;
        const char* e = getenv("GG_GATHER_SMS");
        const int cap = e ? atoi(e) : 0;
        n = cap > 0 && cap < sms ? cap : sms;
    }
    return n;
}

extern "C" int gg_pgat_gather_tiled_multi(const gg_gather_segment* segs, int32_t n_seg, int32_t n_ctas, int32_t ecap,
                                          int32_t raw_k, int32_t G, int32_t C, void* stream) {
    if (!segs || n_seg < 1 || n_seg > kMaxSeg) return GG_EINVAL;
    const int my_ecap = host_ecap(G, C, raw_k);
    if (my_ecap == 0 || ecap != my_ecap || n_ctas < 1) return GG_EINVAL;
    TiledParams p;
    p.n_seg = 0;
    p.inv_sqrt_c = 1.0f / sqrtf((float)C);
    for (int i = 0; i < n_seg; ++i) {
        const gg_gather_segment& g = segs[i];
        if (g.n_dst < 0 || g.n_edges < 0 || g.n_edges > 0x7fffffffLL) return GG_EINVAL;
        if (g.n_dst == 0) continue;
        if (!g.P_src || !g.P_dst || !g.rowptr || !g.Wv3 || !g.agg || !g.ea || !g.cta_ptr) return GG_EINVAL;
        if (g.n_edges > 0 && (!g.col || !g.eattr_csr || !g.wrap_csr || !g.nz || !g.nzptr || !g.tiles || !gg_aligned16(g.tiles))) return GG_EINVAL;
        if ((g.ld_src | g.k_off | g.ld_dst | g.q_off | g.ld_agg) & 3) return GG_EALIGN;
        if (!gg_aligned16(g.P_src) || !gg_aligned16(g.P_dst) || !gg_aligned16(g.agg) || !gg_aligned16(g.Wv3)) return GG_EALIGN;
        TiledSeg& t = p.seg[p.n_seg++];
        t.P_src = g.P_src; t.ld_src = g.ld_src; t.k_off = g.k_off;
        t.P_dst = g.P_dst; t.ld_dst = g.ld_dst; t.q_off = g.q_off;
        t.rowptr = g.rowptr; t.col = g.col; t.ea = g.eattr_csr; t.wrap = g.wrap_csr;
        t.nz = g.nz; t.nzptr = g.nzptr; t.tiles = reinterpret_cast<const int4*>(g.tiles); t.cta_ptr = g.cta_ptr;
        t.Wv3 = g.Wv3;
        t.n_dst = g.n_dst; t.n_edges = (int)g.n_edges;
        t.agg = g.agg; t.ld_agg = g.ld_agg; t.ea_out = g.ea;
    }
    if (p.n_seg == 0) return 0;
    if (!gg_device_is_sm100()) return GG_EARCH;
    const unsigned grid = (unsigned)n_ctas;                  // the tile lists are ordered for exactly this many CTAs
    cudaStream_t st = GG_STREAM(stream);
    cudaError_t err = cudaSuccess;
#define GG_TILED(NV, GV, RAWV)                                                                                                     \
    do {                                                                                                                            \
        using KC = TCfg<NV, GV, RAWV>;                                                                                              \
        const size_t smem = (size_t)KC::NS * KC::BYTES + 16 * KC::NS;                                                               \
        err = cudaFuncSetAttribute(pgat_gather_tiled_kernel<NV, GV, RAWV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (err != cudaSuccess) return (int)err;                                                                                    \
        pgat_gather_tiled_kernel<NV, GV, RAWV><<<grid, kThreads, smem, st>>>(p);                                                    \
    } while (0)
#define GG_TILED_NV(GV, RAWV)                        \
    switch (C / 32) {                                \
        case 1: GG_TILED(1, GV, RAWV); break;        \
        case 2: GG_TILED(2, GV, RAWV); break;        \
        case 3: GG_TILED(3, GV, RAWV); break;        \
        default: GG_TILED(4, GV, RAWV); break;       \
    }
    if (raw_k == 16) { GG_TILED_NV(3, 1) }
    else if (raw_k) { GG_TILED_NV(4, 2) }
    else if (G == 4) { GG_TILED_NV(4, 0) }
    else { GG_TILED_NV(1, 0) }
#undef GG_TILED_NV
#undef GG_TILED
    GG_LAUNCH_OK();
    return 0;
}

extern "C" int gg_pgat_gather_tiled(const float* P_src, int32_t ld_src, int32_t k_off,
                                    const float* P_dst, int32_t ld_dst, int32_t q_off,
                                    const int32_t* rowptr, const int32_t* col, const float* eattr_csr, const int32_t* wrap_csr,
                                    const int32_t* nz, const int32_t* nzptr, const int32_t* tiles, const int32_t* cta_ptr,
                                    int32_t n_ctas, int32_t ecap, int64_t n_edges,
                                    int32_t raw_k, const float* Wv3, int32_t n_dst, int32_t G, int32_t C,
                                    float* agg, int32_t ld_agg, float* ea, void* stream) {
    gg_gather_segment g;
    g.P_src = P_src; g.ld_src = ld_src; g.k_off = k_off;
    g.P_dst = P_dst; g.ld_dst = ld_dst; g.q_off = q_off;
    g.rowptr = rowptr; g.col = col; g.eattr_csr = eattr_csr; g.wrap_csr = wrap_csr;
    g.nz = nz; g.nzptr = nzptr; g.tiles = tiles; g.cta_ptr = cta_ptr;
    g.n_edges = n_edges; g.Wv3 = Wv3; g.n_dst = n_dst;
    g.agg = agg; g.ld_agg = ld_agg; g.ea = ea;
    return gg_pgat_gather_tiled_multi(&g, 1, n_ctas, ecap, raw_k, G, C, stream);
}

// Debug builds (-DGG_TILED_PROFILE): (SM clocks, ns) that consumer warp 0 of CTA 0 spent in each of the last <= 128 launches, in launch
// order, WITHOUT disturbing the launch sequence (one synchronisation when it is read); returns the number of launches since the
// last call and resets the counter.
extern "C" int gg_gather_tiled_clocks(unsigned long long* out) {
#ifdef GG_TILED_PROFILE
    unsigned int n = 0, zero = 0;
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out, g_clk_ring, sizeof(unsigned long long) * 256) != cudaSuccess) return -1;
    if (cudaMemcpyFromSymbol(&n, g_clk_n, sizeof(n)) != cudaSuccess) return -1;
    cudaMemcpyToSymbol(g_clk_n, &zero, sizeof(zero));
    return (int)n;
#else
    (void)out;
    return -1;
#endif
}

// Debug builds (-DGG_TILED_PROFILE): cycles the consumer / producer warps spent waiting and working since the last call, summed
// over warps: out[0..5] = consumer wait, consumer busy, producer wait, producer busy, consumer warps, producer warps.
extern "C" int gg_gather_tiled_profile(unsigned long long* out) {
#ifdef GG_TILED_PROFILE
    unsigned long long z[8] = {0};
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out, g_tiled_prof, sizeof(z)) != cudaSuccess) return GG_EINVAL;
    cudaMemcpyToSymbol(g_tiled_prof, z, sizeof(z));
    return 0;
#else
    (void)out;
    return GG_EINVAL;
#endif
}
