// topology.cu — row f1: the topology update of a rollout step on the device (GrainNN_classifier.update, models.py:614-845 without
// nucleation; switching_edge_index :899-1053; delete_grain_index :864-896; the two-sided sweep :716-727 / :745-755).
//
//   gg_topology_lists   data-parallel: the ascending position lists per joint / grain of the two edge arrays (what every
//                       `(E == p).nonzero()` scan of the reference returns), fixed capacity per node
//   gg_topology_update  one CTA: a walking thread takes the eliminations in the reference's order (topology_core.h — the same routine the
//                       CPU suite checks against the reference's own outputs: in-place edits, appended edges, -1 for deleted columns),
//                       the plain switches of the step run concurrently in conflict-free rounds on fourteen worker warps, a helper
//                       warp sorts and loops over whole lists, a look-ahead warp pulls the walker's table entries into L1;
//                       candidates come straight from gg_select_events' device buffers
//   (cleanup, models.py:846-862, is a stable compaction of the columns that are not -1: the caller's stream compaction)
// What it removes is the host round trip of the full prediction arrays and edge lists (20 + 36 MB at 1.2 10^5 grains) and the host's
// O(E) indexing and compaction per step; an elimination costs the walker ~85 us, a switch a fraction of a microsecond amortised.
#include "common.cuh"
#ifdef GG_TOPO_PROFILE
#include <stdio.h>
#endif

// ---- look-ahead of the walking thread (hooks of topology_core.h).  The walk is a chain of dependent look-ups, each an L2 round
// trip (~0.3 us): ~100 per event.  A second warp of the same CTA reads the table entries of the next few events ahead of the
// walker, so that the walker finds them in the SM's L1.  It only loads (every index bounds-checked, since it may read a list while
// the walker edits it) and never stores to global memory; the walker's results do not depend on it.
#define GG_HINT_SMALL 32
struct TopoHint {
    int32_t small_list[GG_HINT_SMALL];     // copy of a short event list (the caller's array may live in the walker's local memory)
    const int32_t* edges;                  // a long event list, in global memory
    int n, k, epoch, done;
    int grain, gepoch;                     // the grain of the NEXT elimination (its joints' entries are read ahead too)
};
extern __shared__ __align__(16) unsigned char gg_topo_smem[];
static __host__ __device__ __forceinline__ void topo_hint_begin(const int32_t* edges, int n) {
#ifdef __CUDA_ARCH__
    volatile TopoHint* h = reinterpret_cast<volatile TopoHint*>(gg_topo_smem);
    if (n <= GG_HINT_SMALL) { for (int i = 0; i < n; ++i) h->small_list[i] = edges[i]; h->edges = nullptr; }
    else h->edges = edges;
    h->n = n; h->k = 0;
    __threadfence();                       // the list (written by this thread) is visible before the epoch moves
    h->epoch = h->epoch + 1;
#endif
}
static __host__ __device__ __forceinline__ void topo_hint_at(int k) {
#ifdef __CUDA_ARCH__
    reinterpret_cast<volatile TopoHint*>(gg_topo_smem)->k = k;
#endif
}
#ifdef GG_TOPO_PROFILE
__device__ long long g_topo_clk[8];
__device__ long long g_topo_last;
static __host__ __device__ __forceinline__ void topo_mark(int b) {
#ifdef __CUDA_ARCH__
    const long long now = clock64();
    g_topo_clk[b] += now - g_topo_last;
    g_topo_last = now;
#endif
}
#define GG_TOPO_MARK(b) topo_mark(b)
#endif
static __host__ __device__ __forceinline__ void topo_hint_grain(int g) {
#ifdef __CUDA_ARCH__
    volatile TopoHint* h = reinterpret_cast<volatile TopoHint*>(gg_topo_smem);
    h->grain = g;
    h->gepoch = h->gepoch + 1;
#endif
}
#define GG_TOPO_HINT_BEGIN(edges, n) topo_hint_begin(edges, n)
#define GG_TOPO_HINT_AT(k) topo_hint_at(k)
#define GG_TOPO_HINT_GRAIN(g) topo_hint_grain((int)(g))

// ---- helper warp.  The sorts of the walk and its loops over whole lists (the touched joints before and after the switches, the
// candidate list, the sweep's change list) are data-parallel; the walking thread hands the long ones to a third warp of the CTA
// through a request slot in shared memory and waits.  Same integer results by construction, the same single-rounding float
// operations per element; short lists (and lists in the walker's local memory) stay with the walker.
struct GGTopo;
enum { TOPO_OP_PRE = 1, TOPO_OP_POST, TOPO_OP_SORT_PAIRS, TOPO_OP_L1_MARK, TOPO_OP_L1_COMPACT, TOPO_OP_SWEEP, TOPO_OP_SWITCH_PAR };
struct TopoSvc { int req, ack, op, n, flag, ret, enabled, par_enabled; const void* a; void* b; void* c;
                 int par_seq, n_dirty, par_pending, par_err;
                 int gcap; unsigned long long* gkey; };                 // sort keys of lists too long for shared memory: the work area of the concurrent switches      // the concurrent switches: job counter of the worker warps, shared change-list counter
constexpr int kTopoSortCap = 2048;         // elements the helper warp sorts in shared memory; longer lists fall back to the walker
constexpr int kTopoSvcMin = 48;            // shorter lists are not worth the hand-over
constexpr int kTopoParMin = 64;            // ... nor worth the rounds of the concurrent switches
constexpr int kTopoThreads = 512;          // warp 0 walker | 1 look-ahead | 2 helper | 2..15 the worker group of the concurrent switches
constexpr int kTopoWorkers = kTopoThreads - 64;
__device__ __forceinline__ volatile TopoSvc* topo_svc() { return reinterpret_cast<volatile TopoSvc*>(gg_topo_smem + sizeof(TopoHint)); }
__device__ __forceinline__ bool topo_sort_fits(int n) {        // keys of the helper warp's sort: shared memory, else the global area
    int n2 = 1;
    while (n2 < n) n2 <<= 1;
    return n2 <= kTopoSortCap || n2 <= topo_svc()->gcap;
}
static __device__ int topo_service_call(int op, const void* a, void* b, void* c, int n, int flag) {
    volatile TopoSvc* s = topo_svc();
    s->op = op; s->a = a; s->b = b; s->c = c; s->n = n; s->flag = flag;
    __threadfence_block();
    const int seq = s->req + 1;
    s->req = seq;
    while (s->ack != seq) { }
    __threadfence_block();
    return s->ret;
}
static __host__ __device__ __forceinline__ int topo_switch_pre_dev(GGTopo& t, const int32_t* edges, int n, int32_t* touched);
static __host__ __device__ __forceinline__ void topo_switch_post_dev(GGTopo& t, const int32_t* touched, int nt);
static __host__ __device__ __forceinline__ void topo_sort_pairs_dev(int32_t* id, float* val, int n, bool by_value);
static __host__ __device__ __forceinline__ void topo_l1_mark_dev(GGTopo& t, const int32_t* l1, int n);
static __host__ __device__ __forceinline__ int topo_l1_compact_dev(GGTopo& t, int32_t* l1, float* logit, int n);
static __host__ __device__ __forceinline__ int topo_sweep_collect_dev(GGTopo& t, int32_t* cand);
static __host__ __device__ __forceinline__ bool topo_switch_parallel_dev(GGTopo& t, const int32_t* edges, int n);
#define GG_TOPO_SWITCH_PRE(t, edges, n, touched) topo_switch_pre_dev(t, edges, n, touched)
#define GG_TOPO_SWITCH_POST(t, touched, nt) topo_switch_post_dev(t, touched, nt)
#define GG_TOPO_SORT_PAIRS(id, val, n, by_value) topo_sort_pairs_dev(id, val, n, by_value)
#define GG_TOPO_L1_MARK(t, l1, n) topo_l1_mark_dev(t, l1, n)
#define GG_TOPO_L1_COMPACT(t, l1, logit, n) topo_l1_compact_dev(t, l1, logit, n)
#define GG_TOPO_SWEEP_COLLECT(t, cand) topo_sweep_collect_dev(t, cand)
#define GG_TOPO_SWITCH_PARALLEL(t, edges, n) topo_switch_parallel_dev(t, edges, n)
#include "topology_core.h"

// the walker's side of the hooks
static __host__ __device__ __forceinline__ int topo_switch_pre_dev(GGTopo& t, const int32_t* edges, int n, int32_t* touched) {
#ifdef __CUDA_ARCH__
    if (topo_svc()->enabled && n >= kTopoSvcMin && topo_sort_fits(2 * n) && __isGlobal(edges)) return topo_service_call(TOPO_OP_PRE, edges, touched, nullptr, n, 0);
#endif
    return gg_topo_switch_pre_seq(t, edges, n, touched);
}
static __host__ __device__ __forceinline__ void topo_switch_post_dev(GGTopo& t, const int32_t* touched, int nt) {
#ifdef __CUDA_ARCH__
    if (topo_svc()->enabled && nt >= kTopoSvcMin) { topo_service_call(TOPO_OP_POST, touched, nullptr, nullptr, nt, 0); return; }
#endif
    gg_topo_switch_post_seq(t, touched, nt);
}
static __host__ __device__ __forceinline__ void topo_sort_pairs_dev(int32_t* id, float* val, int n, bool by_value) {
#ifdef __CUDA_ARCH__
    if (topo_svc()->enabled && n >= kTopoSvcMin && topo_sort_fits(n)) { topo_service_call(TOPO_OP_SORT_PAIRS, nullptr, id, val, n, by_value ? 1 : 0); return; }
#endif
    gg_topo_sort_pairs(id, val, n, by_value);
}
static __host__ __device__ __forceinline__ void topo_l1_mark_dev(GGTopo& t, const int32_t* l1, int n) {
#ifdef __CUDA_ARCH__
    if (topo_svc()->enabled && n >= kTopoSvcMin) { topo_service_call(TOPO_OP_L1_MARK, l1, nullptr, nullptr, n, 0); return; }
#endif
    gg_topo_l1_mark_seq(t, l1, n);
}
static __host__ __device__ __forceinline__ int topo_l1_compact_dev(GGTopo& t, int32_t* l1, float* logit, int n) {
#ifdef __CUDA_ARCH__
    if (topo_svc()->enabled && n >= kTopoSvcMin) return topo_service_call(TOPO_OP_L1_COMPACT, nullptr, l1, logit, n, 0);
#endif
    return gg_topo_l1_compact_seq(t, l1, logit, n);
}
static __host__ __device__ __forceinline__ bool topo_switch_parallel_dev(GGTopo& t, const int32_t* edges, int n) {
#ifdef __CUDA_ARCH__
    volatile TopoSvc* s = topo_svc();
    if (!s->par_enabled || n < kTopoParMin || n > GG_TOPO_PAR_MAX || !t.par_work || !__isGlobal(edges)) return false;
    s->n_dirty = t.n_dirty;
    const int err = topo_service_call(TOPO_OP_SWITCH_PAR, edges, t.par_work, nullptr, n, 0);
    t.n_dirty = s->n_dirty;
    if (err) t.err = err;
    return true;
#else
    return false;
#endif
}
static __host__ __device__ __forceinline__ int topo_sweep_collect_dev(GGTopo& t, int32_t* cand) {
#ifdef __CUDA_ARCH__
    if (topo_svc()->enabled && !t.dirty_all && t.n_dirty >= kTopoSvcMin && topo_sort_fits(t.n_dirty)) return topo_service_call(TOPO_OP_SWEEP, t.dirty_list, cand, nullptr, t.n_dirty, 0);
#endif
    return gg_topo_sweep_collect_seq(t, cand);
}

namespace {

__global__ void topo_fill_lists(const int64_t* __restrict__ a, int64_t cap, int64_t n, int32_t* __restrict__ l0, int32_t* __restrict__ c0, int cap0,
                                int64_t n0, int32_t* __restrict__ l1, int32_t* __restrict__ c1, int cap1, int64_t n1, int* __restrict__ status) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int64_t u = a[e], v = a[cap + e];
    if (u >= 0) {
        if (u >= n0) { atomicExch(status, GG_TOPO_CAPACITY); return; }
        const int k = atomicAdd(&c0[u], 1);
        if (k < cap0) l0[u * cap0 + k] = (int32_t)e; else atomicExch(status, GG_TOPO_LIST_OVERFLOW);
    }
    if (v >= 0) {
        if (v >= n1) { atomicExch(status, GG_TOPO_CAPACITY); return; }
        const int k = atomicAdd(&c1[v], 1);
        if (k < cap1) l1[v * cap1 + k] = (int32_t)e; else atomicExch(status, GG_TOPO_LIST_OVERFLOW);
    }
}
__global__ void topo_sort_lists(int32_t* __restrict__ l, int32_t* __restrict__ c, int cap, int64_t n_nodes) {
    const int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (v >= n_nodes) return;
    const int n = min(c[v], cap);
    c[v] = n;
    int32_t* p = l + v * cap;
    for (int a = 1; a < n; ++a) {
        const int32_t x = p[a];
        int b = a - 1;
        while (b >= 0 && p[b] > x) { p[b + 1] = p[b]; --b; }
        p[b + 1] = x;
    }
}
__global__ void topo_active(const float* __restrict__ y, int ld, int n, uint8_t* __restrict__ act) {    // models.py:502-503: y[:, 0] > -10
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) act[i] = y[(size_t)i * ld] > -10.0f ? 1 : 0;
}

__global__ void topo_seed_two_sided(const int32_t* __restrict__ cnt, int n_grain, uint8_t* __restrict__ flag, int32_t* __restrict__ list, int32_t* __restrict__ n) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_grain) return;
    const int c = cnt[g];
    if (c > 0 && c <= 2) { flag[g] = 1; list[atomicAdd(n, 1)] = g; }
}

struct TopoArgs {
    GGTopo t;
    const int32_t* ge_count; const int32_t* ge_ids; const float* ge_vals; int ge_cap;       // gg_select_events buffers: grains (id, area)
    const int32_t* l1_count; const int32_t* l1_ids; const float* l1_vals; int l1_cap;       // edges (column, logit)
    int32_t* ge_sorted; int32_t* l1_work; float* l1_logit_work;
    int64_t* switching_list; int32_t* grain_event_out; int32_t* work;
    int64_t* result;                                                                         // {n_pp, n_pq, n_switch, n_grain_event, err, n_ge_in, n_l1_in}
    const int32_t* n_seed;                                                                   // grains with one or two joints on entry (topo_seed_two_sided)
    int prefetch, service, parallel, gcap;
};

// What the switch of edge column e will look at, read by 8 lanes: lanes 0-3 its first end point, 4-7 the second; of each four, lane 0
// walks the joint's own entries and its three grains, lanes 1-3 one joint neighbour each with that neighbour's lists.
// What an event will look at around joint p: j = 0 the joint's own entries and its three grains, j = 1..3 one joint neighbour each
// with that neighbour's lists.
__device__ void topo_prefetch_joint(const GGTopo& t, int64_t p, int j, unsigned& sink) {
    const GGRows& pp = t.pp;
    const GGRows& pq = t.pq;
    if (p < 0 || p >= t.n_joint) return;
    auto joint_row = [&](int64_t v) {
        const int64_t jr = t.jrow ? t.jrow[v] : v;
        if (jr >= 0 && jr < t.n_joint) sink += __float_as_uint(t.xj[jr * t.ld_xj]) + __float_as_uint(t.xj[jr * t.ld_xj + t.col_dxy]);
    };
    if (j == 0) {
        sink += t.act_j[p] + pp.ahead_cnt[p] + __float_as_uint(t.yj[2 * p]) + pp.cnt[1][p] + pp.list[1][p * GG_TOPO_CAP_J];
        joint_row(p);
        const int c = min(pq.cnt[0][p], GG_TOPO_CAP_J);
        for (int i = 0; i < c; ++i) {
            const int32_t pos = pq.list[0][p * GG_TOPO_CAP_J + i];
            if (pos < 0 || pos >= pq.cap) continue;
            sink += (unsigned)pq.a[pos];
            const int64_t q = pq.a[pq.cap + pos];
            if (q >= 0 && q < t.n_grain) sink += pq.cnt[1][q] + pq.list[1][q * GG_TOPO_CAP_G] + t.dirty_flag[q] + __float_as_uint(t.yg[q * t.ld_yg]);
        }
    } else {
        const int c = min(pp.cnt[0][p], GG_TOPO_CAP_J);
        if (j - 1 >= c) return;
        const int32_t pos = pp.list[0][p * GG_TOPO_CAP_J + j - 1];
        if (pos < 0 || pos >= pp.cap) return;
        sink += (unsigned)pp.a[pos] + pp.ahead_flag[pos];
        const int64_t n = pp.a[pp.cap + pos];
        if (n < 0 || n >= t.n_joint) return;
        sink += pp.ahead_cnt[n] + pp.cnt[1][n] + pp.list[1][n * GG_TOPO_CAP_J];
        joint_row(n);
        const int cq = min(pq.cnt[0][n], GG_TOPO_CAP_J);
        for (int i = 0; i < cq; ++i) {
            const int32_t p2 = pq.list[0][n * GG_TOPO_CAP_J + i];
            if (p2 >= 0 && p2 < pq.cap) sink += (unsigned)pq.a[pq.cap + p2];
        }
        const int cp = min(pp.cnt[0][n], GG_TOPO_CAP_J);
        for (int i = 0; i < cp; ++i) {
            const int32_t p3 = pp.list[0][n * GG_TOPO_CAP_J + i];
            if (p3 >= 0 && p3 < pp.cap) sink += (unsigned)pp.a[pp.cap + p3] + (unsigned)pp.a[p3] + pp.ahead_flag[p3];
        }
    }
}
// The switch of edge column e, read by 8 lanes: lanes 0-3 its first end point, 4-7 the second.
__device__ void topo_prefetch_event(const GGTopo& t, int32_t e, int role, unsigned& sink) {
    if (e < 0 || e >= t.pp.cap) return;
    if (role == 0) sink += t.pp.ahead_flag[e];
    topo_prefetch_joint(t, t.pp.a[(role >> 2) * t.pp.cap + e], role & 3, sink);
}
// The elimination of grain g: lane i takes the i-th joint around it, four lanes per joint as above.
__device__ void topo_prefetch_grain(const GGTopo& t, int g, int lane, unsigned& sink) {
    if (g < 0 || g >= t.n_grain) return;
    sink += t.act_g[g] + t.dirty_flag[g];
    const int c = min(t.pq.cnt[1][g], 8);
    const int i = lane >> 2;
    if (i >= c) return;
    const int32_t pos = t.pq.list[1][(int64_t)g * GG_TOPO_CAP_G + i];
    if (pos < 0 || pos >= t.pq.cap) return;
    topo_prefetch_joint(t, t.pq.a[pos], lane & 3, sink);
}

// ---- helper warp: bitonic sort of 64-bit keys in shared memory + the list loops
__device__ __forceinline__ uint32_t topo_ord_i32(int32_t v) { return (uint32_t)v ^ 0x80000000u; }              // signed order -> unsigned order
__device__ __forceinline__ uint32_t topo_ord_f32(float v) {                                                     // float order -> unsigned order (-0 = +0)
    uint32_t u = __float_as_uint(v);
    if (u == 0x80000000u) u = 0u;
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float topo_unord_f32(uint32_t o) { return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o); }
__device__ void topo_warp_sort(unsigned long long* key, int n, int lane) {          // ascending; key[n .. n2) are padded with ~0
    int n2 = 1;
    while (n2 < n) n2 <<= 1;
    for (int i = n + lane; i < n2; i += 32) key[i] = ~0ull;
    __syncwarp();
    for (int k = 2; k <= n2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = lane; i < n2; i += 32) {
                const int x = i ^ j;
                if (x > i) {
                    const unsigned long long a = key[i], b = key[x];
                    if ((a > b) == ((i & k) == 0)) { key[i] = b; key[x] = a; }
                }
            }
            __syncwarp();
        }
}
// order-preserving compaction of the elements with keep(i): every lane calls with its own i per chunk of 32; returns the new count
template <class Keep, class Move>
__device__ int topo_warp_compact(int n, int lane, Keep keep, Move move) {
    int w = 0;
    for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        const bool k = i < n && keep(i);
        const unsigned m = __ballot_sync(0xffffffffu, k);
        move(i, k ? w + __popc(m & ((1u << lane) - 1u)) : -1);       // reads of the chunk happen before its writes (two phases inside)
        w += __popc(m);
    }
    return w;
}

// ---- the plain switches of a list in conflict-free rounds (worker group: warps 2..15, named barrier 1).
// A switch writes the lists, counters and coordinates of its two end points and the joint lists of the (<= 4) grains around them; it
// reads, besides, the entries of the end points' joint neighbours.  Every round each pending event records that footprint on the
// CURRENT tables and marks it with its rank (atomicMin into hashed mark tables: a collision only adds a false conflict); an event
// runs when no event of lower rank that is still pending writes what it reads, reads what it writes, or shares a grain with it.
// The lowest pending rank always runs, so the rounds end; what an event sees is what it would see in the sequential order, because
// a pending lower-rank event can only grow into joints that lower-rank footprints already hold (a switch permutes adjacency
// among its own six joints).  The per-event code is the sequential one (gg_topo_switch_one).
__device__ __forceinline__ void topo_par_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kTopoWorkers) : "memory"); }
__device__ void topo_parallel_switch(const GGTopo& t0, const int32_t* edges, int n, int32_t* work, int tid) {
    volatile TopoSvc* sv = topo_svc();
    GGTopo t = t0;                                               // private copy: error fields, the shared change-list counter
    t.n_dirty_shared = const_cast<int32_t*>(&sv->n_dirty);
    t.dirty_all = false;
    const int64_t M = gg_topo_par_marks(n);
    const uint32_t mask = (uint32_t)(M - 1);
    int32_t* done = work;
    int32_t* fp = work + n;
    int32_t* mark_r = fp + (int64_t)n * GG_TOPO_PAR_FP;
    int32_t* mark_w = mark_r + M;
    int32_t* mark_g = mark_w + M;
    for (int64_t i = tid; i < 3 * M; i += kTopoWorkers) mark_r[i] = 0x7FFFFFFF;
    for (int i = tid; i < n; i += kTopoWorkers) done[i] = 0;
    if (tid == 0) { sv->par_err = 0; sv->par_pending = 0; }
    topo_par_bar();
    for (;;) {
        // (a) footprints of the pending events on the current tables, marked with their ranks
        for (int i = tid; i < n; i += kTopoWorkers) {
            if (done[i]) continue;
            int32_t* f = fp + (int64_t)i * GG_TOPO_PAR_FP;
            const int32_t e = edges[i];
            int nw = 0, nj = 0, ng = 0;
            int32_t* J = f + 3;
            int32_t* G = f + 3 + 18;
            const int64_t pe[2] = {t.pp.get(0, e), t.pp.get(1, e)};
            for (int k = 0; k < 2; ++k) if (pe[k] >= 0 && (k == 0 || pe[1] != pe[0])) J[nj++] = (int32_t)pe[k];
            nw = nj;
            for (int k = 0; k < nw; ++k) {
                const int64_t p = J[k];
                int c; const int32_t* l = t.pp.at(0, p, &c);
                for (int q = 0; q < c && q < GG_TOPO_CAP_J; ++q) {
                    const int32_t v = (int32_t)t.pp.get(1, l[q]);
                    bool seen = v < 0;
                    for (int z = 0; z < nj; ++z) seen = seen || J[z] == v;
                    if (!seen && nj < 18) J[nj++] = v;
                }
                l = t.pq.at(0, p, &c);
                for (int q = 0; q < c && q < GG_TOPO_CAP_J; ++q) {
                    const int32_t g = (int32_t)t.pq.get(1, l[q]);
                    bool seen = g < 0;
                    for (int z = 0; z < ng; ++z) seen = seen || G[z] == g;
                    if (!seen && ng < 6) G[ng++] = g;
                }
            }
            f[0] = nw; f[1] = nj; f[2] = ng;
            for (int z = 0; z < nj; ++z) atomicMin(&mark_r[(uint32_t)J[z] & mask], i);
            for (int z = 0; z < nw; ++z) atomicMin(&mark_w[(uint32_t)J[z] & mask], i);
            for (int z = 0; z < ng; ++z) atomicMin(&mark_g[(uint32_t)G[z] & mask], i);
        }
        topo_par_bar();
        // (b) the events no pending lower rank interferes with run now
        for (int i = tid; i < n; i += kTopoWorkers) {
            if (done[i]) continue;
            const int32_t* f = fp + (int64_t)i * GG_TOPO_PAR_FP;
            const int nw = f[0], nj = f[1], ng = f[2];
            const int32_t* J = f + 3;
            const int32_t* G = f + 3 + 18;
            bool ok = true;
            for (int z = 0; z < nj; ++z) ok = ok && mark_w[(uint32_t)J[z] & mask] >= i;
            for (int z = 0; z < nw; ++z) ok = ok && mark_r[(uint32_t)J[z] & mask] >= i;
            for (int z = 0; z < ng; ++z) ok = ok && mark_g[(uint32_t)G[z] & mask] >= i;
            if (!ok) continue;
            gg_topo_switch_one(t, edges[i], -1, nullptr, 0);
            const int err = t.err ? t.err : (t.pp.err ? t.pp.err : t.pq.err);
            if (err) atomicCAS(const_cast<int*>(&sv->par_err), 0, err);
            done[i] = 2;
        }
        topo_par_bar();
        // (c) clear the marks of this round, count what is left
        for (int i = tid; i < n; i += kTopoWorkers) {
            if (done[i] == 1) continue;
            const int32_t* f = fp + (int64_t)i * GG_TOPO_PAR_FP;
            for (int z = 0; z < f[1]; ++z) { mark_r[(uint32_t)f[3 + z] & mask] = 0x7FFFFFFF; mark_w[(uint32_t)f[3 + z] & mask] = 0x7FFFFFFF; }
            for (int z = 0; z < f[2]; ++z) mark_g[(uint32_t)f[3 + 18 + z] & mask] = 0x7FFFFFFF;
            if (done[i] == 2) done[i] = 1;
            else atomicAdd(const_cast<int*>(&sv->par_pending), 1);
        }
        topo_par_bar();
        const int left = sv->par_pending, err = sv->par_err;
        topo_par_bar();
        if (tid == 0) sv->par_pending = 0;
        if (left == 0 || err) break;
    }
    topo_par_bar();
}
// warps 3..15: wait for a job of the concurrent switches, take part, wait again
__device__ void topo_worker_warp(const GGTopo& t) {
    volatile TopoSvc* s = topo_svc();
    volatile TopoHint* h = reinterpret_cast<volatile TopoHint*>(gg_topo_smem);
    const int lane = threadIdx.x & 31;
    int seen = 0;
    for (;;) {
        int j = 0;
        if (lane == 0) { while ((j = s->par_seq) == seen && !h->done) __nanosleep(500); if (j == seen) j = -1; }
        j = __shfl_sync(0xffffffffu, j, 0);
        if (j < 0) return;
        seen = j;
        __threadfence_block();
        topo_parallel_switch(t, static_cast<const int32_t*>(const_cast<const void*>(s->a)), s->n, static_cast<int32_t*>(const_cast<void*>(s->b)), (int)threadIdx.x - 64);
    }
}

__device__ void topo_service_warp(const GGTopo& t) {
    __shared__ unsigned long long key_smem[kTopoSortCap];
    volatile TopoSvc* s = topo_svc();
    volatile TopoHint* h = reinterpret_cast<volatile TopoHint*>(gg_topo_smem);
    const int lane = threadIdx.x & 31;
    const GGRows& pp = t.pp;
    int done = 0;
    for (;;) {
        int r = 0;
        if (lane == 0) { while ((r = s->req) == done && !h->done) __nanosleep(40); if (r == done) r = -1; }
        r = __shfl_sync(0xffffffffu, r, 0);
        if (r < 0) return;
        __threadfence_block();
        const int op = s->op, n = s->n, flag = s->flag;
        const void* a = const_cast<const void*>(s->a);
        void* b = const_cast<void*>(s->b);
        void* c = const_cast<void*>(s->c);
        int ret = 0;
        const int need = op == TOPO_OP_PRE ? 2 * n : n;          // keys this request sorts: shared memory when they fit, else the global area
        int need2 = 1;
        while (need2 < need) need2 <<= 1;
        unsigned long long* key = need2 <= kTopoSortCap ? key_smem : const_cast<unsigned long long*>(s->gkey);
        if (op == TOPO_OP_PRE) {                                 // gg_topo_switch_pre_seq
            const int32_t* edges = static_cast<const int32_t*>(a);
            int32_t* touched = static_cast<int32_t*>(b);
            for (int k = lane; k < n; k += 32) {
                const int32_t e = edges[k];
                const int64_t u = pp.get(0, e), v = pp.get(1, e);
                key[2 * k] = topo_ord_i32((int32_t)u); key[2 * k + 1] = topo_ord_i32((int32_t)v);
                pp.ahead_flag[e] |= 1;
                atomicAdd(&pp.ahead_cnt[u], 1);
                atomicAdd(&pp.ahead_cnt[v], 1);
            }
            __syncwarp();
            topo_warp_sort(key, 2 * n, lane);
            const int nt = topo_warp_compact(2 * n, lane, [&](int i) { return i == 0 || key[i] != key[i - 1]; },
                                             [&](int i, int w) { if (w >= 0) touched[w] = (int32_t)((uint32_t)key[i] ^ 0x80000000u); });
            __syncwarp();
            for (int i = lane; i < nt; i += 32) {
                const int64_t p = touched[i];
                float* x = gg_topo_xrow(t, p);
                x[0] = gg_tsub(x[0], gg_tdiv(t.yj[2 * p], 5.0f));
                x[1] = gg_tsub(x[1], gg_tdiv(t.yj[2 * p + 1], 5.0f));
            }
            ret = nt;
        } else if (op == TOPO_OP_POST) {                         // gg_topo_switch_post_seq
            const int32_t* touched = static_cast<const int32_t*>(a);
            for (int i = lane; i < n; i += 32) {
                const int64_t p = touched[i];
                float* x = gg_topo_xrow(t, p);
                const float y0 = gg_tmul(5.0f, gg_tsub(x[0], x[0])), y1 = gg_tmul(5.0f, gg_tsub(x[1], x[1]));
                t.yj[2 * p] = y0; t.yj[2 * p + 1] = y1;
                x[t.col_dxy] = y0; x[t.col_dxy + 1] = y1;
            }
        } else if (op == TOPO_OP_SORT_PAIRS) {                   // gg_topo_sort_pairs (a total order: any correct sort gives the same array)
            int32_t* id = static_cast<int32_t*>(b);
            float* val = static_cast<float*>(c);
            for (int i = lane; i < n; i += 32)
                key[i] = flag ? ((unsigned long long)(~topo_ord_f32(val[i])) << 32) | topo_ord_i32(id[i])
                              : ((unsigned long long)topo_ord_i32(id[i]) << 32) | __float_as_uint(val[i]);
            __syncwarp();
            topo_warp_sort(key, n, lane);
            for (int i = lane; i < n; i += 32) {
                const unsigned long long kk = key[i];
                if (flag) { id[i] = (int32_t)((uint32_t)kk ^ 0x80000000u); val[i] = topo_unord_f32(~(uint32_t)(kk >> 32)); }
                else { id[i] = (int32_t)((uint32_t)(kk >> 32) ^ 0x80000000u); val[i] = __uint_as_float((uint32_t)kk); }
            }
        } else if (op == TOPO_OP_L1_MARK) {
            const int32_t* l1 = static_cast<const int32_t*>(a);
            for (int i = lane; i < n; i += 32) pp.ahead_flag[l1[i]] |= 4;
        } else if (op == TOPO_OP_L1_COMPACT) {                   // gg_topo_l1_compact_seq
            int32_t* l1 = static_cast<int32_t*>(b);
            float* logit = static_cast<float*>(c);
            int32_t e = 0; float lg = 0.f;
            ret = topo_warp_compact(n, lane,
                                    [&](int i) { e = l1[i]; lg = logit[i]; const uint8_t f = pp.ahead_flag[e]; pp.ahead_flag[e] = f & 1; return !(f & 2); },
                                    [&](int i, int w) { __syncwarp(); if (w >= 0) { l1[w] = e; logit[w] = lg; } __syncwarp(); });
        } else if (op == TOPO_OP_SWITCH_PAR) {                   // release the worker warps and take part (threads 0..31 of the group)
            if (lane == 0) { __threadfence_block(); s->par_seq = s->par_seq + 1; }
            __syncwarp();
            topo_parallel_switch(t, static_cast<const int32_t*>(a), n, static_cast<int32_t*>(b), lane);
            ret = s->par_err;
        } else if (op == TOPO_OP_SWEEP) {                        // gg_topo_sweep_collect_seq, the change-list branch
            const int32_t* dirty = static_cast<const int32_t*>(a);
            int32_t* cand = static_cast<int32_t*>(b);
            const int nc = topo_warp_compact(n, lane, [&](int i) { const int cg = t.pq.cnt[1][dirty[i]]; return cg > 0 && cg <= 2; },
                                             [&](int i, int w) { if (w >= 0) key[w] = topo_ord_i32(dirty[i]); });
            __syncwarp();
            topo_warp_sort(key, nc, lane);
            for (int i = lane; i < nc; i += 32) cand[i] = (int32_t)((uint32_t)key[i] ^ 0x80000000u);
            for (int i = lane; i < n; i += 32) t.dirty_flag[dirty[i]] = 0;
            ret = nc;
        }
        __threadfence_block();
        __syncwarp();
        if (lane == 0) { s->ret = ret; __threadfence_block(); s->ack = r; }
        done = r;
    }
}

#ifndef GG_TOPO_LOOKAHEAD
#define GG_TOPO_LOOKAHEAD 1      // events between the walker and the next wave of four (measured: 1: 5.0 ms, 3: 5.1, 6: 5.35, 12: 6.05 for 320 events)
#endif
constexpr int kTopoLookAhead = GG_TOPO_LOOKAHEAD;

__device__ void topo_prefetch_warp(const GGTopo& t) {
    volatile TopoHint* h = reinterpret_cast<volatile TopoHint*>(gg_topo_smem);
    const int lane = threadIdx.x & 31;
    int seen = 0, gseen = 0;
    unsigned sink = 0;
    for (;;) {
        if (h->done) break;
        const int gep = h->gepoch;
        if (gep != gseen) { gseen = gep; topo_prefetch_grain(t, h->grain, lane, sink); __syncwarp(); }
        const int ep = h->epoch;
        if (ep == seen) { __nanosleep(200); continue; }
        seen = ep;
        const int n = h->n;
        const int32_t* edges = const_cast<const int32_t*>(h->edges);
        for (int base = 0; base < n; base += 4) {
            while (h->k + kTopoLookAhead < base && h->epoch == ep && !h->done) __nanosleep(100);
            if (h->epoch != ep || h->done) break;
            const int idx = base + (lane >> 3);
            if (idx < n) topo_prefetch_event(t, edges ? __ldcg(edges + idx) : h->small_list[idx], lane & 7, sink);
            __syncwarp();
        }
    }
    if (sink == 0x9E3779B9u) h->small_list[0] = (int32_t)sink;      // keeps the loads
}

__global__ void __launch_bounds__(kTopoThreads, 1) topo_update_kernel(TopoArgs A) {
    {
        volatile TopoHint* h = reinterpret_cast<volatile TopoHint*>(gg_topo_smem);
        if (threadIdx.x == 0) {
            h->n = 0; h->k = 0; h->epoch = 0; h->done = 0; h->edges = nullptr; h->grain = -1; h->gepoch = 0;
            volatile TopoSvc* sv = topo_svc();
            sv->req = 0; sv->ack = 0; sv->enabled = A.service; sv->par_enabled = A.service && A.parallel; sv->par_seq = 0;
            sv->gkey = reinterpret_cast<unsigned long long*>(A.t.par_work); sv->gcap = A.gcap;
        }
        __syncthreads();
        if (threadIdx.x >= 96) { if (A.service && A.parallel) topo_worker_warp(A.t); return; }
        if (threadIdx.x >= 64) { if (A.service) topo_service_warp(A.t); return; }
        if (threadIdx.x >= 32) { if (A.prefetch) topo_prefetch_warp(A.t); return; }
    }
    if (threadIdx.x != 0) return;
#ifdef GG_TOPO_PROFILE
    for (int i = 0; i < 8; ++i) g_topo_clk[i] = 0;
    g_topo_last = clock64();
#endif
    GGTopo& t = A.t;
    t.preseeded = true; t.n_dirty = *A.n_seed;
    const int n_ge = min(*A.ge_count, A.ge_cap), n_l1 = min(*A.l1_count, A.l1_cap);
    A.result[5] = *A.ge_count; A.result[6] = *A.l1_count;
    // test.py:414-416: candidates sorted by predicted area ascending (equal areas: by grain id)
    for (int i = 0; i < n_ge; ++i) A.ge_sorted[i] = i;
    for (int a = 1; a < n_ge; ++a) {
        const int32_t o = A.ge_sorted[a];
        const float v = A.ge_vals[o]; const int32_t id = A.ge_ids[o];
        int b = a - 1;
        while (b >= 0 && (A.ge_vals[A.ge_sorted[b]] > v || (A.ge_vals[A.ge_sorted[b]] == v && A.ge_ids[A.ge_sorted[b]] > id))) { A.ge_sorted[b + 1] = A.ge_sorted[b]; --b; }
        A.ge_sorted[b + 1] = o;
    }
    for (int i = 0; i < n_ge; ++i) A.ge_sorted[i] = A.ge_ids[A.ge_sorted[i]];
    // NOTE: ge_sorted now holds grain ids; the in-place rewrite above is safe because slot i is read (as an index) before it is written
    for (int i = 0; i < n_l1; ++i) { A.l1_work[i] = A.l1_ids[i]; A.l1_logit_work[i] = A.l1_vals[i]; }
    GGTopoResult r = gg_topo_update(t, A.ge_sorted, n_ge, A.l1_work, A.l1_logit_work, n_l1, A.switching_list, A.grain_event_out, A.work);
    A.result[0] = t.pp.n; A.result[1] = t.pq.n; A.result[2] = r.n_switch; A.result[3] = r.n_grain_event; A.result[4] = r.err;
    reinterpret_cast<volatile TopoHint*>(gg_topo_smem)->done = 1;
#ifdef GG_TOPO_PROFILE
    printf("topo clk: sort %lld | eliminations %lld | sweeps %lld | L1 sort %lld | switch pre %lld main %lld post %lld | last sweep %lld\n",
           g_topo_clk[0], g_topo_clk[1], g_topo_clk[2], g_topo_clk[3], g_topo_clk[4], g_topo_clk[5], g_topo_clk[6], g_topo_clk[7]);
#endif
}

}  // namespace

extern "C" int gg_topology_lists(const int64_t* edges, int64_t cap, int64_t n, int32_t* list0, int32_t* cnt0, int32_t cap0, int64_t n0,
                                 int32_t* list1, int32_t* cnt1, int32_t cap1, int64_t n1, int32_t* status, void* stream) {
    if (cap < n || n < 0 || n0 < 0 || n1 < 0 || cap0 < 1 || cap1 < 1 || !cnt0 || !cnt1 || !list0 || !list1 || !status || (n > 0 && !edges)) return GG_EINVAL;
    cudaStream_t st = GG_STREAM(stream);
    cudaError_t err;
    if ((err = cudaMemsetAsync(cnt0, 0, sizeof(int32_t) * (size_t)n0, st)) != cudaSuccess) return (int)err;
    if ((err = cudaMemsetAsync(cnt1, 0, sizeof(int32_t) * (size_t)n1, st)) != cudaSuccess) return (int)err;
    if ((err = cudaMemsetAsync(status, 0, sizeof(int32_t), st)) != cudaSuccess) return (int)err;
    if (n > 0) { topo_fill_lists<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(edges, cap, n, list0, cnt0, cap0, n0, list1, cnt1, cap1, n1, status); GG_LAUNCH_OK(); }
    if (n0 > 0) { topo_sort_lists<<<(unsigned)((n0 + 255) / 256), 256, 0, st>>>(list0, cnt0, cap0, n0); GG_LAUNCH_OK(); }
    if (n1 > 0) { topo_sort_lists<<<(unsigned)((n1 + 255) / 256), 256, 0, st>>>(list1, cnt1, cap1, n1); GG_LAUNCH_OK(); }
    return 0;
}

extern "C" int gg_topology_caps(int32_t* cap_joint, int32_t* cap_grain) {
    if (cap_joint) *cap_joint = GG_TOPO_CAP_J;
    if (cap_grain) *cap_grain = GG_TOPO_CAP_G;
    return 0;
}

extern "C" int64_t gg_topology_work_ints(int64_t n_l1, int64_t n_ge, int64_t n_grain) { return gg_topo_work_ints(n_l1, n_ge, n_grain); }

extern "C" int gg_topology_update(int64_t* pp, int64_t cap_pp, int64_t n_pp, int64_t* pq, int64_t cap_pq, int64_t n_pq,
                                  int32_t* pp_list0, int32_t* pp_cnt0, int32_t* pp_list1, int32_t* pp_cnt1,
                                  int32_t* pq_list0, int32_t* pq_cnt0, int32_t* pq_list1, int32_t* pq_cnt1,
                                  int32_t* ahead_cnt, uint8_t* ahead_flag,
                                  float* x_joint, int32_t ld_xj, const int32_t* joint_row, int32_t col_dxy,
                                  float* y_joint, const float* y_grain, int32_t ld_yg,
                                  float* mask_grain, float* mask_joint, uint8_t* act_grain, uint8_t* act_joint,
                                  int32_t n_joint, int32_t n_grain,
                                  const int32_t* ge_count, const int32_t* ge_ids, const float* ge_vals, int32_t ge_cap,
                                  const int32_t* l1_count, const int32_t* l1_ids, const float* l1_vals, int32_t l1_cap,
                                  uint8_t* dirty_flag, int32_t* dirty_list, int32_t* scratch,
                                  int32_t* ge_sorted, int32_t* l1_work, float* l1_logit_work,
                                  int64_t* switching_list, int32_t* grain_event_out, int32_t* work, int64_t* result, void* stream) {
    if (!pp || !pq || !x_joint || !y_joint || !y_grain || !mask_grain || !mask_joint || !act_grain || !act_joint || !result) return GG_EINVAL;
    if (!ge_count || !l1_count || !switching_list || !grain_event_out || !work || !scratch || !dirty_flag || !dirty_list) return GG_EINVAL;
    if (n_pp > cap_pp || n_pq > cap_pq || n_joint < 0 || n_grain < 0 || ld_xj < col_dxy + 2) return GG_EINVAL;
    cudaStream_t st = GG_STREAM(stream);
    topo_active<<<(n_grain + 255) / 256, 256, 0, st>>>(y_grain, ld_yg, n_grain, act_grain); GG_LAUNCH_OK();
    topo_active<<<(n_joint + 255) / 256, 256, 0, st>>>(y_joint, 2, n_joint, act_joint); GG_LAUNCH_OK();
    TopoArgs A;
    memset(&A, 0, sizeof(A));
    A.t.pp.a = pp; A.t.pp.cap = cap_pp; A.t.pp.n = n_pp;
    A.t.pp.list[0] = pp_list0; A.t.pp.cnt[0] = pp_cnt0; A.t.pp.lcap[0] = GG_TOPO_CAP_J;
    A.t.pp.list[1] = pp_list1; A.t.pp.cnt[1] = pp_cnt1; A.t.pp.lcap[1] = GG_TOPO_CAP_J;
    A.t.pp.ahead_cnt = ahead_cnt; A.t.pp.ahead_flag = ahead_flag;
    A.t.pq.a = pq; A.t.pq.cap = cap_pq; A.t.pq.n = n_pq;
    A.t.pq.list[0] = pq_list0; A.t.pq.cnt[0] = pq_cnt0; A.t.pq.lcap[0] = GG_TOPO_CAP_J;
    A.t.pq.list[1] = pq_list1; A.t.pq.cnt[1] = pq_cnt1; A.t.pq.lcap[1] = GG_TOPO_CAP_G;
    A.t.xj = x_joint; A.t.ld_xj = ld_xj; A.t.jrow = joint_row; A.t.col_dxy = col_dxy;
    A.t.yj = y_joint; A.t.yg = y_grain; A.t.ld_yg = ld_yg;
    A.t.mask_g = mask_grain; A.t.ld_mg = 1; A.t.mask_j = mask_joint; A.t.ld_mj = 1;
    A.t.act_g = act_grain; A.t.act_j = act_joint; A.t.n_joint = n_joint; A.t.n_grain = n_grain;
    A.t.dirty_flag = dirty_flag; A.t.dirty_list = dirty_list; A.t.scratch = scratch;
    A.ge_count = ge_count; A.ge_ids = ge_ids; A.ge_vals = ge_vals; A.ge_cap = ge_cap;
    A.l1_count = l1_count; A.l1_ids = l1_ids; A.l1_vals = l1_vals; A.l1_cap = l1_cap;
    A.ge_sorted = ge_sorted; A.l1_work = l1_work; A.l1_logit_work = l1_logit_work;
    A.switching_list = switching_list; A.grain_event_out = grain_event_out; A.work = work; A.result = result;
    // the first two-sided sweep looks at the grains that have one or two joints NOW plus those the events touch, not at every grain
    int32_t* n_seed = scratch + n_grain + 2 * (l1_cap + ge_cap) + 120;          // (the tail of scratch: 128 spare ints)
    cudaError_t err = cudaMemsetAsync(n_seed, 0, sizeof(int32_t), st);
    if (err != cudaSuccess) return (int)err;
    topo_seed_two_sided<<<(n_grain + 255) / 256, 256, 0, st>>>(pq_cnt1, n_grain, dirty_flag, dirty_list, n_seed); GG_LAUNCH_OK();
    A.n_seed = n_seed;
    { const char* e = getenv("GG_TOPO_PREFETCH"); A.prefetch = !(e && e[0] == '0'); }        // (measurement switches)
    { const char* e = getenv("GG_TOPO_SERVICE"); A.service = !(e && e[0] == '0'); }
    { const char* e = getenv("GG_TOPO_PARALLEL"); A.parallel = !(e && e[0] == '0'); }
    A.t.par_work = work + gg_topo_seq_ints(l1_cap, ge_cap, n_grain);
    { int64_t g = 1; while (2 * g <= gg_topo_par_ints(l1_cap) / 2) g <<= 1; A.gcap = (int)(g > (1 << 20) ? (1 << 20) : g); }
    static bool carve_set = false;
    if (!carve_set) {                     // the walk lives on L1 hits (look-ahead warp): keep shared memory at what the helper warp needs
        cudaFuncSetAttribute(topo_update_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 10);
        carve_set = true;
    }
    topo_update_kernel<<<1, kTopoThreads, sizeof(TopoHint) + sizeof(TopoSvc), st>>>(A);
    GG_LAUNCH_OK();
    return 0;
}
