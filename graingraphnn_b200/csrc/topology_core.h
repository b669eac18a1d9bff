// topology_core.h — the topology update of a rollout step as ONE sequential routine over device-resident
// tables, written once for the device kernel (topology.cu: a single thread walks the events, everything else — the position
// lists before, the stable compaction after — is data-parallel) and for the host harness of the CPU test suite
// (tests/topology_host.cpp compiles this header with g++ and checks it against the reference's own outputs,
// tests/golden/topology_golden.npz).
//
// Follows GrainNN_classifier.update (models.py:614-845 without the optional nucleation branch :771-835), switching_edge_index
// (:899-1053), delete_grain_index (:864-896) and the two-sided sweep (:716-727, :745-755) decision for decision, with the
// reference's in-place discipline on the edge arrays: edits overwrite positions, new edges are appended, -1 marks deleted columns
// until the final stable compaction — which is what makes the result comparable position for position.  Every `(E == p).nonzero()`
// scan of the reference is answered from an ascending position list kept per joint / grain (fixed capacity: a joint has three
// joint and three grain neighbours, a grain at most GG_TOPO_CAP_G joints).  Float arithmetic is IEEE single, one rounding per
// operation like torch's (no contraction: -ffp-contract=off on the host, __f*_rn on the device).
// Orders: eliminations by predicted area ascending (test.py:416), switches by probability descending (models.py:730-731) with
// equal probabilities — logits that saturate or collide in fp32 included — in ascending edge position, as the reference's sort
// leaves them: the caller hands over sigmoid(logit) as torch computes it, not the logit.
#pragma once
#include <stdint.h>
#ifndef __CUDACC__
#define GG_TD inline
#include <cmath>
#else
#define GG_TD __host__ __device__ __forceinline__
#endif

// Hooks of the device build (topology.cu defines them before including this header): the walking thread announces the list of
// switching events it is about to process and its position in it, so that a second warp can pull the table entries of the events
// to come into the SM's L1 (results are unaffected; the host build compiles them away).
#ifndef GG_TOPO_HINT_BEGIN
#define GG_TOPO_HINT_BEGIN(edges, n) ((void)0)
#define GG_TOPO_HINT_AT(k) ((void)0)
#endif
#ifndef GG_TOPO_HINT_GRAIN
#define GG_TOPO_HINT_GRAIN(g) ((void)0)     // the grain whose elimination comes after the current one
#endif
// The data-parallel parts of the walk (the two sorts, the loops over the touched joints before and after the switches of a list,
// the bookkeeping of the candidate list) go through these; the device build hands the long ones to a helper warp.
#ifndef GG_TOPO_SWITCH_PRE
#define GG_TOPO_SWITCH_PRE(t, edges, n, touched) gg_topo_switch_pre_seq(t, edges, n, touched)
#define GG_TOPO_SWITCH_POST(t, touched, nt) gg_topo_switch_post_seq(t, touched, nt)
#define GG_TOPO_SORT_PAIRS(id, val, n, by_value) gg_topo_sort_pairs(id, val, n, by_value)
#define GG_TOPO_L1_MARK(t, l1, n) gg_topo_l1_mark_seq(t, l1, n)
#define GG_TOPO_L1_COMPACT(t, l1, logit, n) gg_topo_l1_compact_seq(t, l1, logit, n)
#define GG_TOPO_SWEEP_COLLECT(t, cand) gg_topo_sweep_collect_seq(t, cand)
#endif
#ifndef GG_TOPO_SWITCH_PARALLEL
#define GG_TOPO_SWITCH_PARALLEL(t, edges, n) false      // device build: the plain switches of a list run in conflict-free rounds
#endif
#ifndef GG_TOPO_MARK
#define GG_TOPO_MARK(bucket) ((void)0)     // profiling build of the device kernel: time since the last mark goes to `bucket`
#endif

#define GG_TOPO_CAP_J 8        // positions per joint and row (3 in a valid tiling, 4 transiently inside a switch)
#define GG_TOPO_CAP_G 32       // joints per grain

enum GGTopoError {
    GG_TOPO_OK = 0, GG_TOPO_LIST_OVERFLOW = 1, GG_TOPO_NOT_TWO_SIDED = 2, GG_TOPO_NO_COMMON_GRAIN = 3, GG_TOPO_BAD_VALENCE = 4,
    GG_TOPO_ACROSS_MISMATCH = 5, GG_TOPO_GROW = 6, GG_TOPO_CAPACITY = 7
};

GG_TD float gg_tsub(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fsub_rn(a, b);
#else
    volatile float r = a - b; return r;
#endif
}
GG_TD float gg_tadd(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fadd_rn(a, b);
#else
    volatile float r = a + b; return r;
#endif
}
GG_TD float gg_tmul(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fmul_rn(a, b);
#else
    volatile float r = a * b; return r;
#endif
}
GG_TD float gg_tdiv(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fdiv_rn(a, b);
#else
    volatile float r = a / b; return r;
#endif
}
// periodic_move (models.py:1097-1100) in float32: (p - [rel > .5]) + [rel < -.5], rel = p - pc
GG_TD float gg_wrap_to(float p, float pc) {
    const float rel = gg_tsub(p, pc);
    return gg_tadd(gg_tsub(p, rel > 0.5f ? 1.0f : 0.0f), rel < -0.5f ? 1.0f : 0.0f);
}

// One [2, cap] int64 edge array + ascending position lists per (row, value).
struct GGRows {
    int64_t* a;            // row r at a + r * cap
    int64_t cap, n;        // capacity / used columns (appended edges grow n)
    int32_t* list[2];      // list[r][v * lcap[r] + k]
    int32_t* cnt[2];
    int32_t lcap[2];
    int32_t* ahead_cnt;    // (pp only, inside switch) occurrences of every joint among the end points of the events still to come
    uint8_t* ahead_flag;   // (pp only) bit 0: column is one of the events still to come; bit 2: column is a switching candidate of this
                           // step (L1), bit 1: ... that was a switched side of an eliminated grain and leaves L1 (models.py:713)
    int err;

    GG_TD int64_t get(int r, int64_t pos) const { return a[r * cap + pos]; }
    GG_TD const int32_t* at(int r, int64_t v, int* n_out) const { *n_out = cnt[r][v]; return list[r] + v * lcap[r]; }
    GG_TD void list_remove(int r, int64_t v, int32_t pos) {
        int32_t* l = list[r] + v * lcap[r];
        int c = cnt[r][v], k = 0;
        while (k < c && l[k] != pos) ++k;
        if (k == c) return;
        for (; k + 1 < c; ++k) l[k] = l[k + 1];
        cnt[r][v] = c - 1;
    }
    GG_TD void list_insert(int r, int64_t v, int32_t pos) {       // bisect.insort: keeps the list ascending
        int32_t* l = list[r] + v * lcap[r];
        int c = cnt[r][v];
        if (c >= lcap[r]) { err = GG_TOPO_LIST_OVERFLOW; return; }
        int k = c;
        while (k > 0 && l[k - 1] > pos) { l[k] = l[k - 1]; --k; }
        l[k] = pos;
        cnt[r][v] = c + 1;
    }
    GG_TD void set(int r, int64_t pos, int64_t v) {
        const int64_t old = a[r * cap + pos];
        if (old == v) return;
        if (old >= 0) list_remove(r, old, (int32_t)pos);
        if (v >= 0) list_insert(r, v, (int32_t)pos);
        if (ahead_flag && (ahead_flag[pos] & 1)) {
            if (old >= 0) --ahead_cnt[old];
            if (v >= 0) ++ahead_cnt[v];
        }
        a[r * cap + pos] = v;
    }
    GG_TD void kill(int64_t pos) { set(0, pos, -1); set(1, pos, -1); }
    GG_TD void append(int64_t v0, int64_t v1) {
        if (n >= cap) { err = GG_TOPO_CAPACITY; return; }
        const int64_t pos = n++;
        a[pos] = -1; a[cap + pos] = -1;
        set(0, pos, v0); set(1, pos, v1);
    }
};

struct GGTopo {
    GGRows pp, pq;                 // joint->joint, joint->grain
    float* xj; int ld_xj;          // joint features (x, y in columns 0..1; dx, dy in columns col_dxy, col_dxy + 1)
    const int32_t* jrow;           // joint id -> row of xj (NULL: identity; an engine that keeps its rows in another order passes its map)
    int col_dxy;                   // first of the two prediction columns of the joint features (6)
    float* yj;                     // [Nj, 2] predicted joint displacements (y_dict['joint'])
    const float* yg; int ld_yg;    // y_dict['grain'] (column 0 orders the sides of a vanishing grain)
    float* mask_g; int ld_mg;      // [Ng] / [Nj] live flags (set to 0 here)
    float* mask_j; int ld_mj;
    const uint8_t* act_g; const uint8_t* act_j;
    int32_t n_joint, n_grain;
    uint8_t* dirty_flag; int32_t* dirty_list; int32_t n_dirty; bool dirty_all;   // grains whose joint count changed since the last two-side check
    int32_t* n_dirty_shared;       // device, events running concurrently: the list is appended through this counter (else NULL)
    int32_t* par_work;             // device: work area of the concurrent switches (gg_topo_par_ints ints behind the sequential work area)
    bool preseeded;                // dirty_list already holds the grains that have one or two joints on entry (a data-parallel pre-pass):
                                   // the first sweep then looks at them and at the grains the events touched instead of at every grain
    int32_t* scratch;              // >= 2 * max events + n_grain ints
    int err;
};

GG_TD float* gg_topo_xrow(const GGTopo& t, int64_t p) { return t.xj + (int64_t)(t.jrow ? t.jrow[p] : p) * t.ld_xj; }

GG_TD void gg_topo_dirty_add(GGTopo& t, int64_t g) {
    if (t.dirty_all || g < 0) return;
#ifdef __CUDA_ARCH__
    if (t.n_dirty_shared) {                                        // test-and-set of the grain's flag byte, then an atomic append
        unsigned* w = reinterpret_cast<unsigned*>(t.dirty_flag + (g & ~(int64_t)3));
        const unsigned bit = 1u << (8 * (unsigned)(g & 3));
        if (!(atomicOr(w, bit) & (0xFFu << (8 * (unsigned)(g & 3))))) t.dirty_list[atomicAdd(t.n_dirty_shared, 1)] = (int32_t)g;
        return;
    }
#endif
    if (!t.dirty_flag[g]) { t.dirty_flag[g] = 1; t.dirty_list[t.n_dirty++] = (int32_t)g; }
}
GG_TD void gg_topo_pq_set_grain(GGTopo& t, int64_t pos, int64_t g) {
    gg_topo_dirty_add(t, t.pq.get(1, pos));
    gg_topo_dirty_add(t, g);
    t.pq.set(1, pos, g);
}
// positions e with pp[0, e] == p1 and (pp[1, e] == p2) == equal, ascending; returns how many (out holds <= GG_TOPO_CAP_J)
GG_TD int gg_topo_pp_between(const GGTopo& t, int64_t p1, int64_t p2, bool equal, int32_t* out) {
    int c, k = 0;
    const int32_t* l = t.pp.at(0, p1, &c);
    for (int i = 0; i < c; ++i)
        if ((t.pp.get(1, l[i]) == p2) == equal) out[k++] = l[i];
    return k;
}

// delete_grain_index (models.py:864-896): a two-sided grain and its two joints disappear, their outer neighbours are linked
GG_TD void gg_topo_delete_grain(GGTopo& t, int64_t grain) {
    int c;
    const int32_t* lg = t.pq.at(1, grain, &c);
    if (c != 2) { t.err = GG_TOPO_NOT_TWO_SIDED; return; }
    const int64_t p1 = t.pq.get(0, lg[0]), p2 = t.pq.get(0, lg[1]);
    int32_t tmp[GG_TOPO_CAP_J];
    if (gg_topo_pp_between(t, p1, p2, false, tmp) < 1) { t.err = GG_TOPO_BAD_VALENCE; return; }
    const int64_t n1 = t.pp.get(1, tmp[0]);
    if (gg_topo_pp_between(t, p2, p1, false, tmp) < 1) { t.err = GG_TOPO_BAD_VALENCE; return; }
    const int64_t n2 = t.pp.get(1, tmp[0]);
    t.pp.append(n1, n2);
    t.pp.append(n2, n1);
    t.mask_g[(int64_t)grain * t.ld_mg] = 0.f;
    t.mask_j[p1 * t.ld_mj] = 0.f;
    t.mask_j[p2 * t.ld_mj] = 0.f;
    int32_t copy[GG_TOPO_CAP_G];
    {
        int n; const int32_t* l = t.pq.at(1, grain, &n);
        for (int i = 0; i < n; ++i) copy[i] = l[i];
        for (int i = 0; i < n; ++i) t.pq.kill(copy[i]);
    }
    const int64_t pj[2] = {p1, p2};
    for (int s = 0; s < 2; ++s) {
        const int64_t j = pj[s];
        int n; const int32_t* l = t.pq.at(0, j, &n);
        for (int i = 0; i < n; ++i) copy[i] = l[i];
        for (int i = 0; i < n; ++i) { gg_topo_dirty_add(t, t.pq.get(1, copy[i])); t.pq.kill(copy[i]); }
        int n0; const int32_t* l0 = t.pp.at(0, j, &n0);
        for (int i = 0; i < n0; ++i) copy[i] = l0[i];
        int n1_; const int32_t* l1 = t.pp.at(1, j, &n1_);
        for (int i = 0; i < n1_; ++i) copy[n0 + i] = l1[i];
        for (int i = 0; i < n0 + n1_; ++i) t.pp.kill(copy[i]);
    }
    if (t.pp.err) t.err = t.pp.err;
    if (t.pq.err) t.err = t.pq.err;
}

GG_TD void gg_topo_sort_i32(int32_t* v, int n) {                 // ascending; heapsort: the lists live in global memory and one thread sorts them
    for (int start = n / 2 - 1; start >= 0; --start) {
        int root = start; const int32_t x = v[root];
        for (;;) { int c = 2 * root + 1; if (c >= n) break; if (c + 1 < n && v[c + 1] > v[c]) ++c; if (v[c] <= x) break; v[root] = v[c]; root = c; }
        v[root] = x;
    }
    for (int end = n - 1; end > 0; --end) {
        const int32_t x = v[end]; v[end] = v[0];
        int root = 0;
        for (;;) { int c = 2 * root + 1; if (c >= end) break; if (c + 1 < end && v[c + 1] > v[c]) ++c; if (v[c] <= x) break; v[root] = v[c]; root = c; }
        v[root] = x;
    }
}
// (id, value) pairs: by_value = false: id ascending; true: value descending, equal values by id ascending (a total order: heapsort)
GG_TD bool gg_topo_pair_after(int32_t ia, float va, int32_t ib, float vb, bool by_value) {      // does a come after b?
    if (!by_value) return ia > ib;
    return va < vb || (va == vb && ia > ib);
}
GG_TD void gg_topo_sort_pairs(int32_t* id, float* val, int n, bool by_value) {
    for (int start = n / 2 - 1; start >= 0; --start) {
        int root = start; const int32_t xi = id[root]; const float xv = val[root];
        for (;;) { int c = 2 * root + 1; if (c >= n) break; if (c + 1 < n && gg_topo_pair_after(id[c + 1], val[c + 1], id[c], val[c], by_value)) ++c;
                   if (!gg_topo_pair_after(id[c], val[c], xi, xv, by_value)) break; id[root] = id[c]; val[root] = val[c]; root = c; }
        id[root] = xi; val[root] = xv;
    }
    for (int end = n - 1; end > 0; --end) {
        const int32_t xi = id[end]; const float xv = val[end]; id[end] = id[0]; val[end] = val[0];
        int root = 0;
        for (;;) { int c = 2 * root + 1; if (c >= end) break; if (c + 1 < end && gg_topo_pair_after(id[c + 1], val[c + 1], id[c], val[c], by_value)) ++c;
                   if (!gg_topo_pair_after(id[c], val[c], xi, xv, by_value)) break; id[root] = id[c]; val[root] = val[c]; root = c; }
        id[root] = xi; val[root] = xv;
    }
}

// The two-sided sweep (models.py:716-727 / :745-755): every grain left with one or two joints is deleted; returns how many were
// appended to `out` (ascending grain ids).
// candidates of a sweep: the grains with one or two joints among those whose joint count changed (all grains on the first sweep of
// a step that was not pre-seeded), ascending; the change list is emptied
GG_TD int gg_topo_sweep_collect_seq(GGTopo& t, int32_t* cand) {
    int nc = 0;
    if (t.dirty_all) {                                            // first check of the step: every grain (torch.unique over E_pq[1])
        for (int32_t g = 0; g < t.n_grain; ++g) { const int c = t.pq.cnt[1][g]; if (c > 0 && c <= 2) cand[nc++] = g; }
    } else {
        for (int i = 0; i < t.n_dirty; ++i) { const int32_t g = t.dirty_list[i]; const int c = t.pq.cnt[1][g]; if (c > 0 && c <= 2) cand[nc++] = g; }
        gg_topo_sort_i32(cand, nc);
    }
    for (int i = 0; i < t.n_dirty; ++i) t.dirty_flag[t.dirty_list[i]] = 0;
    return nc;
}
GG_TD int gg_topo_delete_two_sided(GGTopo& t, int32_t* out) {
    int32_t* cand = t.scratch;
    const int nc = GG_TOPO_SWEEP_COLLECT(t, cand);
    t.n_dirty = 0;
    t.dirty_all = false;
    // the candidate list lives in scratch, which delete_grain does not touch
    for (int i = 0; i < nc && !t.err; ++i) { gg_topo_delete_grain(t, cand[i]); out[i] = cand[i]; }
    return nc;
}

GG_TD bool gg_topo_inside(const float* t_, const float* v1, const float* v2, const float* v3) {   // point_in_triangle, models.py:1055-1072
    const float a[2] = {gg_wrap_to(v1[0], t_[0]), gg_wrap_to(v1[1], t_[1])};
    const float b[2] = {gg_wrap_to(v2[0], t_[0]), gg_wrap_to(v2[1], t_[1])};
    const float c[2] = {gg_wrap_to(v3[0], t_[0]), gg_wrap_to(v3[1], t_[1])};
#define GG_SIGN(A, B, C) gg_tsub(gg_tmul(gg_tsub((A)[0], (C)[0]), gg_tsub((B)[1], (C)[1])), gg_tmul(gg_tsub((B)[0], (C)[0]), gg_tsub((A)[1], (C)[1])))
    const float d0 = GG_SIGN(t_, a, b), d1 = GG_SIGN(t_, b, c), d2 = GG_SIGN(t_, c, a);
#undef GG_SIGN
    const bool neg = d0 < 0 || d1 < 0 || d2 < 0, pos = d0 > 0 || d1 > 0 || d2 > 0;
    return !(neg && pos);
}

// Before the switches of a list (models.py:905-907 and the bookkeeping of "the events still to come"); returns |touched|.
GG_TD int gg_topo_switch_pre_seq(GGTopo& t, const int32_t* edges, int n_edges, int32_t* touched) {
    GGRows& pp = t.pp;
    int nt = 0;
    for (int k = 0; k < n_edges; ++k) { touched[nt++] = (int32_t)pp.get(0, edges[k]); touched[nt++] = (int32_t)pp.get(1, edges[k]); }
    gg_topo_sort_i32(touched, nt);
    { int w = 0; for (int i = 0; i < nt; ++i) if (i == 0 || touched[i] != touched[w - 1]) touched[w++] = touched[i]; nt = w; }
    for (int i = 0; i < nt; ++i) {
        const int64_t p = touched[i];
        float* x = gg_topo_xrow(t, p);
        x[0] = gg_tsub(x[0], gg_tdiv(t.yj[2 * p], 5.0f));
        x[1] = gg_tsub(x[1], gg_tdiv(t.yj[2 * p + 1], 5.0f));
    }
    for (int k = 0; k < n_edges; ++k) {
        pp.ahead_flag[edges[k]] |= 1;
        ++pp.ahead_cnt[pp.get(0, edges[k])];
        ++pp.ahead_cnt[pp.get(1, edges[k])];
    }
    return nt;
}
// After them: y <- 5 (x - before) with `before` a VIEW of x in the reference (it has followed every move): exactly zero (:1046-1050)
GG_TD void gg_topo_switch_post_seq(GGTopo& t, const int32_t* touched, int nt) {
    for (int i = 0; i < nt; ++i) {
        const int64_t p = touched[i];
        float* x = gg_topo_xrow(t, p);
        const float y0 = gg_tmul(5.0f, gg_tsub(x[0], x[0])), y1 = gg_tmul(5.0f, gg_tsub(x[1], x[1]));
        t.yj[2 * p] = y0; t.yj[2 * p + 1] = y1;
        x[t.col_dxy] = y0; x[t.col_dxy + 1] = y1;
    }
}
// `L1 = [e for e in L1 if e not in sides]` after every elimination (models.py:713) is applied once, before the switches: the
// candidates are flagged, an elimination flags those of its switched sides that are candidates, the compaction drops them.
GG_TD void gg_topo_l1_mark_seq(GGTopo& t, const int32_t* L1, int n) { for (int i = 0; i < n; ++i) t.pp.ahead_flag[L1[i]] |= 4; }
GG_TD int gg_topo_l1_compact_seq(GGTopo& t, int32_t* L1, float* logit, int n) {      // keeps the order; clears the flags
    int w = 0;
    for (int i = 0; i < n; ++i) {
        const int32_t e = L1[i];
        const uint8_t f = t.pp.ahead_flag[e];
        t.pp.ahead_flag[e] = f & 1;
        if (!(f & 2)) { L1[w] = e; logit[w] = logit[i]; ++w; }
    }
    return w;
}

// One switching event: edge column e of pp (the loop body of switching_edge_index, models.py:909-1040).
GG_TD int gg_topo_switch_one(GGTopo& t, int32_t e, int64_t elim_grain, int32_t* forced, int n_forced) {
    GGRows& pp = t.pp;
    GGRows& pq = t.pq;
    const int64_t p1 = pp.get(0, e), p2 = pp.get(1, e);
    if (p1 >= 0 && p2 >= 0 && t.act_j[p1] && t.act_j[p2]) {
        int c1, c2;
        int32_t at_q1[GG_TOPO_CAP_J], at_q2[GG_TOPO_CAP_J], at_n1[GG_TOPO_CAP_J], at_n2[GG_TOPO_CAP_J];
        int64_t q1[GG_TOPO_CAP_J], q2[GG_TOPO_CAP_J];
        { const int32_t* l = pq.at(0, p1, &c1); for (int i = 0; i < c1; ++i) { at_q1[i] = l[i]; q1[i] = pq.get(1, l[i]); } }
        { const int32_t* l = pq.at(0, p2, &c2); for (int i = 0; i < c2; ++i) { at_q2[i] = l[i]; q2[i] = pq.get(1, l[i]); } }
        const int nn1 = gg_topo_pp_between(t, p1, p2, false, at_n1), nn2 = gg_topo_pp_between(t, p2, p1, false, at_n2);
        if (c1 != 3 || c2 != 3 || nn1 < 2 || nn2 < 2) { t.err = GG_TOPO_BAD_VALENCE; return n_forced; }
        int64_t n1[2] = {pp.get(1, at_n1[0]), pp.get(1, at_n1[1])}, n2[2] = {pp.get(1, at_n2[0]), pp.get(1, at_n2[1])};
        // grains: the two shared ones shrink, the unshared one of each joint grows across (:925-945)
        int64_t grow1 = -1, grow2 = -1, shrink[2];
        int ng1 = 0, ng2 = 0, ns = 0;
        for (int i = 0; i < 3; ++i) {
            int m = 0; for (int j = 0; j < 3; ++j) m += q2[j] == q1[i];
            if (m != 1) { grow1 = q1[i]; ++ng1; }
            if (m != 0) { if (ns < 2) shrink[ns] = q1[i]; ++ns; }
        }
        for (int i = 0; i < 3; ++i) {
            int m = 0; for (int j = 0; j < 3; ++j) m += q1[j] == q2[i];
            if (m != 1) { grow2 = q2[i]; ++ng2; }
        }
        if (ns != 2) { t.err = GG_TOPO_NO_COMMON_GRAIN; return n_forced; }
        const int64_t shrink_a = shrink[0], shrink_b = shrink[1];
        int32_t slots1[2 * GG_TOPO_CAP_J], slots2[2 * GG_TOPO_CAP_J];
        int ns1 = 0, ns2 = 0;
        for (int i = 0; i < 3; ++i) if (q1[i] == shrink_a) slots1[ns1++] = at_q1[i];
        for (int i = 0; i < 3; ++i) if (q1[i] == shrink_b) slots1[ns1++] = at_q1[i];
        for (int i = 0; i < 3; ++i) if (q2[i] == shrink_a) slots2[ns2++] = at_q2[i];
        for (int i = 0; i < 3; ++i) if (q2[i] == shrink_b) slots2[ns2++] = at_q2[i];
        // order the outer neighbours: the one that touches shrink_a first (:947-975)
        for (int side = 0; side < 2; ++side) {
            int64_t* nn = side == 0 ? n1 : n2;
            int32_t* at = side == 0 ? at_n1 : at_n2;
            int c; const int32_t* l = pq.at(0, nn[0], &c);
            bool touches = false;
            for (int i = 0; i < c; ++i) touches = touches || pq.get(1, l[i]) == shrink_a;
            if (!touches) { const int64_t tn = nn[0]; nn[0] = nn[1]; nn[1] = tn; const int32_t ta = at[0]; at[0] = at[1]; at[1] = ta; }
        }
        int64_t a1 = n1[0], b1 = n1[1], a2 = n2[0], b2 = n2[1];
        if (!(elim_grain < 0 && (a1 == a2 || b1 == b2))) {
            if (a1 == a2 && shrink_a != elim_grain) forced[n_forced++] = (int32_t)shrink_a;
            if (b1 == b2 && shrink_b != elim_grain) forced[n_forced++] = (int32_t)shrink_b;
            // both ends collapse onto the midpoint (:989-996)
            float* x1 = gg_topo_xrow(t, p1);
            float* x2 = gg_topo_xrow(t, p2);
            const float m0 = gg_tmul(0.5f, gg_tadd(x1[0], gg_wrap_to(x2[0], x1[0])));
            const float m1 = gg_tmul(0.5f, gg_tadd(x1[1], gg_wrap_to(x2[1], x1[1])));
            const float w0 = gg_wrap_to(m0, x2[0]), w1 = gg_wrap_to(m1, x2[1]);
            x1[0] = m0; x1[1] = m1; x2[0] = w0; x2[1] = w1;
            bool swap = gg_topo_inside(x2, x1, gg_topo_xrow(t, a1), gg_topo_xrow(t, a2));
            const int32_t* ah = pp.ahead_cnt;
            if (ah[a2] > 0 && !(ah[b2] > 0)) swap = false;
            if (ah[b2] > 0 && !(ah[a2] > 0)) swap = true;
            if (ah[a1] > 0 && !(ah[b1] > 0)) swap = true;
            if (ah[b1] > 0 && !(ah[a1] > 0)) swap = false;
            if (swap) {
                for (int i = 0; i < ns1 / 2; ++i) { const int32_t s = slots1[i]; slots1[i] = slots1[ns1 - 1 - i]; slots1[ns1 - 1 - i] = s; }
                for (int i = 0; i < ns2 / 2; ++i) { const int32_t s = slots2[i]; slots2[i] = slots2[ns2 - 1 - i]; slots2[ns2 - 1 - i] = s; }
                { const int32_t s = at_n1[0]; at_n1[0] = at_n1[1]; at_n1[1] = s; }
                { const int32_t s = at_n2[0]; at_n2[0] = at_n2[1]; at_n2[1] = s; }
                { const int64_t s = a1; a1 = b1; b1 = s; }
                { const int64_t s = a2; a2 = b2; b2 = s; }
            }
            if (ng1 != 1 || ng2 != 1 || ns1 < 2 || ns2 < 1) { t.err = GG_TOPO_GROW; return n_forced; }
            GG_TOPO_MARK(elim_grain < 0 ? 5 : 1);
            gg_topo_pq_set_grain(t, slots1[1], grow2);
            gg_topo_pq_set_grain(t, slots2[0], grow1);
            pp.set(0, at_n1[1], p2);
            pp.set(0, at_n2[0], p1);
            int32_t tmp[GG_TOPO_CAP_J];
            int c = gg_topo_pp_between(t, a2, p2, true, tmp);
            for (int i = 0; i < c; ++i) pp.set(1, tmp[i], p1);
            c = gg_topo_pp_between(t, b1, p1, true, tmp);
            for (int i = 0; i < c; ++i) pp.set(1, tmp[i], p2);
            GG_TOPO_MARK(elim_grain < 0 ? 6 : 1);
        }
    }
    // this event is no longer "to come"
    pp.ahead_flag[e] &= (uint8_t)~1u;
    { const int64_t u = pp.get(0, e), v = pp.get(1, e); if (u >= 0) --pp.ahead_cnt[u]; if (v >= 0) --pp.ahead_cnt[v]; }

    return n_forced;
}

// switching_edge_index (models.py:899-1053) over the edge columns `edges[0..n_edges)`; elim_grain < 0: plain neighbour switching.
// Forced eliminations are appended to forced[] (returns the new count).
GG_TD int gg_topo_switch(GGTopo& t, const int32_t* edges, int n_edges, int64_t elim_grain, int32_t* forced, int n_forced) {
    GGRows& pp = t.pp;
    GGRows& pq = t.pq;
    // touched = sorted set of the end points; every touched joint first steps back by its predicted displacement (:905-907);
    // the end points of the events still to come are counted per joint and kept current by GGRows::set
    int32_t* touched = t.scratch + t.n_grain;                     // (scratch[0 .. n_grain) is the two-sided sweep's)
    const int nt = GG_TOPO_SWITCH_PRE(t, edges, n_edges, touched);
    GG_TOPO_MARK(elim_grain < 0 ? 4 : 1);
    if (!(elim_grain < 0 && GG_TOPO_SWITCH_PARALLEL(t, edges, n_edges))) {
        GG_TOPO_HINT_BEGIN(edges, n_edges);
        for (int k = 0; k < n_edges && !t.err; ++k) {
            GG_TOPO_HINT_AT(k);
            n_forced = gg_topo_switch_one(t, edges[k], elim_grain, forced, n_forced);
        }
    }
    GG_TOPO_MARK(elim_grain < 0 ? 5 : 1);
    if (t.err) {                                                  // leave the ahead tables clean
        for (int k = 0; k < n_edges; ++k) if (pp.ahead_flag[edges[k]] & 1) {
            pp.ahead_flag[edges[k]] &= (uint8_t)~1u;
            const int64_t u = pp.get(0, edges[k]), v = pp.get(1, edges[k]);
            if (u >= 0) --pp.ahead_cnt[u];
            if (v >= 0) --pp.ahead_cnt[v];
        }
    }
    GG_TOPO_SWITCH_POST(t, touched, nt);
    if (pp.err) t.err = pp.err;
    if (pq.err) t.err = pq.err;
    GG_TOPO_MARK(elim_grain < 0 ? 6 : 1);
    return n_forced;
}

// GrainNN_classifier.update (models.py:614-768).  grain_event[0..n_ge): candidate grains sorted by predicted area ascending;
// L1[0..n_l1): candidate edge columns with their probabilities, ANY order (sorted here: probability descending, then column ascending).
// Outputs: switching_list [<= n_l1][2], grain_event_out (the input followed by the forced / two-sided eliminations).
// work: int32 [3 * (n_l1 + n_ge) + 64 + ...] see gg_topo_work_ints.
struct GGTopoResult { int32_t n_switch, n_grain_event, err; };

#define GG_TOPO_PAR_MAX 16384      // the concurrent-switch path of the device build takes lists up to this long
#define GG_TOPO_PAR_FP 28          // ints of an event's footprint record: {writes, joints, grains, joints[18], grains[6], pad}
GG_TD int64_t gg_topo_par_marks(int64_t n) { int64_t m = 4096; while (m < 64 * n) m <<= 1; return m; }      // slots of one mark table
GG_TD int64_t gg_topo_seq_ints(int64_t n_l1, int64_t n_ge, int64_t n_grain) { return 4 * n_l1 + 2 * n_grain + 2 * n_ge + 4 * GG_TOPO_CAP_G + 64; }
GG_TD int64_t gg_topo_par_ints(int64_t n_l1) {
    const int64_t n = n_l1 < GG_TOPO_PAR_MAX ? n_l1 : GG_TOPO_PAR_MAX;
    return n * (1 + GG_TOPO_PAR_FP) + 3 * gg_topo_par_marks(n);
}
GG_TD int64_t gg_topo_work_ints(int64_t n_l1, int64_t n_ge, int64_t n_grain) { return gg_topo_seq_ints(n_l1, n_ge, n_grain) + gg_topo_par_ints(n_l1); }

GG_TD GGTopoResult gg_topo_update(GGTopo& t, const int32_t* grain_event, int n_ge, int32_t* L1, float* L1_logit, int n_l1,
                                  int64_t* switching_list, int32_t* grain_event_out, int32_t* work) {
    GGTopoResult res = {0, 0, 0};
    for (int i = 0; i < n_ge; ++i) grain_event_out[i] = grain_event[i];
    int n_out = n_ge;
    int32_t* forced = work;                                       // <= 2 per event
    int32_t* sides = work + 2 * (n_l1 + n_ge) + 16;               // <= CAP_G
    int32_t* ord = sides + GG_TOPO_CAP_G;
    int32_t* removed = ord + GG_TOPO_CAP_G;                       // two-sided sweep output, <= n_grain
    int64_t around[GG_TOPO_CAP_G], across[GG_TOPO_CAP_G];
    if (t.preseeded) t.dirty_all = false;
    else { t.dirty_all = true; t.n_dirty = 0; }
    // L1 ascending by column first (the reference's nonzero order), so that ties of the later sort are by column
    GG_TOPO_SORT_PAIRS(L1, L1_logit, n_l1, false);
    GG_TOPO_L1_MARK(t, L1, n_l1);
    GG_TOPO_MARK(0);
    int n_unexpected = 0;
    int32_t* unexpected = work + n_l1 + n_ge + 8;                 // forced + swept grains, in the reference's order
    for (int gi = 0; gi < n_ge && !t.err; ++gi) {                  // models.py:638-727
        const int64_t grain = grain_event[gi];
        GG_TOPO_HINT_GRAIN(gi + 1 < n_ge ? grain_event[gi + 1] : -1);
        if (!t.act_g[grain]) continue;
        int na; const int32_t* lg = t.pq.at(1, grain, &na);
        if (na == 0 || na > GG_TOPO_CAP_G) continue;
        bool all_active = true;
        for (int i = 0; i < na; ++i) { around[i] = t.pq.get(0, lg[i]); all_active = all_active && t.act_j[around[i]]; }
        if (!all_active) continue;
        int n_sides = 0, n_across = 0;
        for (int i = 0; i < na && !t.err; ++i)
            for (int j = i + 1; j < na; ++j) {                    // itertools.combinations(around, 2)
                int64_t p1 = around[i], p2 = around[j];
                if (p1 > p2) { const int64_t s = p1; p1 = p2; p2 = s; }
                int32_t tmp[GG_TOPO_CAP_J];
                const int c = gg_topo_pp_between(t, p1, p2, true, tmp);
                if (!c) continue;
                for (int k = 0; k < c && n_sides < GG_TOPO_CAP_G; ++k) sides[n_sides++] = tmp[k];
                int64_t g1[GG_TOPO_CAP_J], g2[GG_TOPO_CAP_J];
                int n1 = 0, n2 = 0, cc;
                { const int32_t* l = t.pq.at(0, p1, &cc); for (int k = 0; k < cc; ++k) if (t.pq.get(1, l[k]) != grain) g1[n1++] = t.pq.get(1, l[k]); }
                { const int32_t* l = t.pq.at(0, p2, &cc); for (int k = 0; k < cc; ++k) if (t.pq.get(1, l[k]) != grain) g2[n2++] = t.pq.get(1, l[k]); }
                if (n1 < 2) { t.err = GG_TOPO_BAD_VALENCE; break; }
                bool in0 = false, in1 = false;
                for (int k = 0; k < n2; ++k) { in0 = in0 || g2[k] == g1[0]; in1 = in1 || g2[k] == g1[1]; }
                if (in0) across[n_across++] = g1[0];
                else if (in1) across[n_across++] = g1[1];
                else { t.err = GG_TOPO_NO_COMMON_GRAIN; break; }
            }
        if (t.err) break;
        if (n_across != na) { t.err = GG_TOPO_ACROSS_MISMATCH; break; }
        bool distinct = true;
        for (int i = 0; i < n_across; ++i) for (int j = i + 1; j < n_across; ++j) distinct = distinct && across[i] != across[j];
        if (!distinct) continue;
        // torch.sort(y['grain'][across, 0]) ascending (stable on ties), all but the last two sides switch (:686-690)
        for (int i = 0; i < n_across; ++i) ord[i] = i;
        for (int a = 1; a < n_across; ++a) {
            const int32_t o = ord[a];
            const float v = t.yg[across[o] * t.ld_yg];
            int b = a - 1;
            while (b >= 0 && t.yg[across[ord[b]] * t.ld_yg] > v) { ord[b + 1] = ord[b]; --b; }
            ord[b + 1] = o;
        }
        int32_t sw[GG_TOPO_CAP_G];
        const int n_sw = n_across >= 2 ? n_across - 2 : 0;
        for (int i = 0; i < n_sw; ++i) sw[i] = sides[ord[i]];
        const int nf = gg_topo_switch(t, sw, n_sw, grain, forced, 0);
        if (t.err) break;
        for (int i = 0; i < nf; ++i) unexpected[n_unexpected++] = forced[i];
        gg_topo_delete_grain(t, grain);
        for (int i = 0; i < nf && !t.err; ++i) gg_topo_delete_grain(t, forced[i]);
        if (t.err) break;
        for (int k = 0; k < n_sw; ++k) if (t.pp.ahead_flag[sw[k]] & 4) t.pp.ahead_flag[sw[k]] |= 2;     // L1 = [e for e in L1 if e not in sides]
        GG_TOPO_MARK(1);
        gg_topo_delete_two_sided(t, removed);                     // (its victims are not reported, models.py:716-727)
        GG_TOPO_MARK(2);
    }
    n_l1 = GG_TOPO_L1_COMPACT(t, L1, L1_logit, n_l1);             // (also on an error: the flags are left clean)
    if (!t.err && n_l1 > 0) {                                      // models.py:730-740
        // probability descending = logit descending; equal logits keep ascending column order
        GG_TOPO_SORT_PAIRS(L1, L1_logit, n_l1, true);
        int w = 0;
        for (int i = 0; i < n_l1; ++i) if (t.pp.get(0, L1[i]) != -1) { L1[w] = L1[i]; L1_logit[w] = L1_logit[i]; ++w; }
        n_l1 = w;
        GG_TOPO_MARK(3);
        gg_topo_switch(t, L1, n_l1, -1, forced, 0);
        for (int i = 0; i < n_l1; ++i) { switching_list[2 * i] = t.pp.get(0, L1[i]); switching_list[2 * i + 1] = t.pp.get(1, L1[i]); }
        res.n_switch = n_l1;
    }
    if (!t.err) {
        const int n = gg_topo_delete_two_sided(t, removed);
        for (int i = 0; i < n; ++i) unexpected[n_unexpected++] = removed[i];
        GG_TOPO_MARK(7);
    }
    for (int i = 0; i < n_unexpected; ++i) grain_event_out[n_out++] = unexpected[i];
    res.n_grain_event = n_out;
    res.err = t.err;
    return res;
}
