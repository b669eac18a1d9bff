// Host build of csrc/topology_core.h for the CPU test suite (tests/test_topology_core.py compiles it with g++ -ffp-contract=off):
// the sequential surgery that topology.cu runs in one device thread, checked here against the reference's own outputs
// (tests/golden/topology_golden.npz) without a GPU.  TEST INFRASTRUCTURE ONLY — the product runs the same header on the device.
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../graingraphnn_b200/csrc/topology_core.h"

static void build_lists(GGRows& r, int64_t n0, int64_t n1, std::vector<int32_t>& l0, std::vector<int32_t>& c0,
                        std::vector<int32_t>& l1, std::vector<int32_t>& c1, int cap0, int cap1) {
    l0.assign((size_t)n0 * cap0, 0); c0.assign((size_t)n0, 0); l1.assign((size_t)n1 * cap1, 0); c1.assign((size_t)n1, 0);
    r.list[0] = l0.data(); r.cnt[0] = c0.data(); r.lcap[0] = cap0;
    r.list[1] = l1.data(); r.cnt[1] = c1.data(); r.lcap[1] = cap1;
    for (int64_t e = 0; e < r.n; ++e) {                      // ascending positions: appended in order
        const int64_t u = r.a[e], v = r.a[r.cap + e];
        if (u >= 0 && c0[u] < cap0) l0[u * cap0 + c0[u]++] = (int32_t)e;
        if (v >= 0 && c1[v] < cap1) l1[v * cap1 + c1[v]++] = (int32_t)e;
    }
}

// pp / pq: int64 [2, cap] with n used columns.  Returns the error code; n_out = {n_pp, n_pq, n_switch, n_grain_event}.
extern "C" int topology_update_host(int64_t* pp, int64_t cap_pp, int64_t n_pp, int64_t* pq, int64_t cap_pq, int64_t n_pq,
                                    float* xj, int ld_xj, int col_dxy, float* yj, const float* yg, int ld_yg,
                                    float* mask_g, float* mask_j, const uint8_t* act_g, const uint8_t* act_j,
                                    int n_joint, int n_grain, const int32_t* grain_event, int n_ge,
                                    int32_t* L1, float* L1_logit, int n_l1, int64_t* switching_list, int32_t* grain_event_out,
                                    int64_t* n_out, int preseed) {
    GGTopo t;
    memset(&t, 0, sizeof(t));
    t.pp.a = pp; t.pp.cap = cap_pp; t.pp.n = n_pp;
    t.pq.a = pq; t.pq.cap = cap_pq; t.pq.n = n_pq;
    std::vector<int32_t> a0, a1, a2, a3, b0, b1, b2, b3;
    build_lists(t.pp, n_joint, n_joint, a0, b0, a1, b1, GG_TOPO_CAP_J, GG_TOPO_CAP_J);
    build_lists(t.pq, n_joint, n_grain, a2, b2, a3, b3, GG_TOPO_CAP_J, GG_TOPO_CAP_G);
    std::vector<int32_t> ahead_cnt((size_t)n_joint, 0);
    std::vector<uint8_t> ahead_flag((size_t)cap_pp, 0), dirty_flag((size_t)n_grain, 0);
    std::vector<int32_t> dirty_list((size_t)n_grain, 0), scratch((size_t)n_grain + 2 * (size_t)(n_l1 + n_ge + GG_TOPO_CAP_G) + 64, 0);
    std::vector<int32_t> work((size_t)gg_topo_work_ints(n_l1, n_ge, n_grain), 0);
    t.pp.ahead_cnt = ahead_cnt.data(); t.pp.ahead_flag = ahead_flag.data();
    t.xj = xj; t.ld_xj = ld_xj; t.col_dxy = col_dxy; t.yj = yj; t.yg = yg; t.ld_yg = ld_yg;
    t.mask_g = mask_g; t.ld_mg = 1; t.mask_j = mask_j; t.ld_mj = 1; t.act_g = act_g; t.act_j = act_j;
    t.n_joint = n_joint; t.n_grain = n_grain;
    t.dirty_flag = dirty_flag.data(); t.dirty_list = dirty_list.data(); t.scratch = scratch.data();
    if (preseed) {                                           // what topo_seed_two_sided does on the device
        t.preseeded = true;
        for (int g = 0; g < n_grain; ++g) { const int c = t.pq.cnt[1][g]; if (c > 0 && c <= 2) { dirty_flag[g] = 1; dirty_list[t.n_dirty++] = g; } }
    }
    GGTopoResult r = gg_topo_update(t, grain_event, n_ge, L1, L1_logit, n_l1, switching_list, grain_event_out, work.data());
    n_out[0] = t.pp.n; n_out[1] = t.pq.n; n_out[2] = r.n_switch; n_out[3] = r.n_grain_event;
    return r.err;
}
