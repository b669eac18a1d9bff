"""CPU oracle of the rollout's topology update.  TEST INFRASTRUCTURE ONLY.

A restatement, in the reference's op order, of GrainNN_classifier.update (models.py:614-845, without the optional nucleation
branch :771-835, which needs nucleation_prob > 1e-6), switching_edge_index (:899-1053), delete_grain_index (:864-896),
cleanup (:846-862) and the geometry helpers point_in_triangle / periodic_move (:1055-1108).  Edge arrays keep the
reference's in-place discipline — edits overwrite positions, new edges are appended, `-1` marks deleted rows until the final
stable compaction — so the outputs are comparable array for array.  Pinned by tests/golden/topology_golden.npz, the outputs of
the reference's own update (oracle/make_golden_topology.py); tests/test_topology_golden.py.
Lookups are O(E) scans like the reference's; a device version replaces them by adjacency tables.
"""
import itertools

import torch

JJ, JG, GJ = ('joint', 'connect', 'joint'), ('joint', 'pull', 'grain'), ('grain', 'push', 'joint')
JOINT_SCALE = 5          # models.py:546


def _where(cond):
    return cond.nonzero().view(-1)


def _wrap_to(p, pc):
    """periodic_move, models.py:1097-1100."""
    rel = p - pc
    return p - 1 * (rel > 0.5) + 1 * (rel < -0.5)


def _inside(t, v1, v2, v3):
    """point_in_triangle, models.py:1055-1072: t inside (or on) the triangle of the three points moved next to t."""
    def sign(a, b, c):
        return (a[0] - c[0]) * (b[1] - c[1]) - (b[0] - c[0]) * (a[1] - c[1])
    a, b, c = _wrap_to(v1, t), _wrap_to(v2, t), _wrap_to(v3, t)
    d = (sign(t, a, b), sign(t, b, c), sign(t, c, a))
    neg = bool(d[0] < 0) or bool(d[1] < 0) or bool(d[2] < 0)
    pos = bool(d[0] > 0) or bool(d[1] > 0) or bool(d[2] > 0)
    return not (neg and pos)


class _Topology:
    def __init__(self, x_dict, edge_index_dict, y_dict, mask, active_grains, active_joints):
        self.x, self.y, self.mask = x_dict, y_dict, mask
        self.pp = edge_index_dict[JJ]            # E_pp, edited in place / re-bound on append
        self.pq = edge_index_dict[JG]            # E_pq
        self.active_grains, self.active_joints = active_grains, active_joints

    # ---- lookups (ascending positions, like `.nonzero()` on the reference's masks)
    def joints_of(self, grain):
        return self.pq[0][_where(self.pq[1] == grain)]

    def has_pq(self, joint, grain):
        return len(_where((self.pq[0] == joint) & (self.pq[1] == grain))) > 0

    # ---- models.py:864-896
    def delete_grain(self, grain):
        around = self.joints_of(grain)
        assert len(around) == 2, around
        p1, p2 = around
        n1 = self.pp[1][_where((self.pp[0] == p1) & (self.pp[1] != p2))][0]
        n2 = self.pp[1][_where((self.pp[0] == p2) & (self.pp[1] != p1))][0]
        self.pp = torch.cat([self.pp, torch.tensor([[n1, n2], [n2, n1]])], dim=-1)
        self.mask['grain'][grain] = 0
        self.mask['joint'][p1] = 0
        self.mask['joint'][p2] = 0
        self.pq[:, _where(self.pq[1] == grain)] = -1
        for j in (p1, p2):
            self.pq[:, _where(self.pq[0] == j)] = -1
            self.pp[:, _where(self.pp[0] == j)] = -1
            self.pp[:, _where(self.pp[1] == j)] = -1

    def delete_two_sided(self):
        """models.py:716-727 / :745-755."""
        grains, counts = torch.unique(self.pq[1, :], return_counts=True)
        left = grains[counts <= 2]
        for g in left:
            self.delete_grain(g)
        return left

    # ---- models.py:899-1053
    def switch(self, edges, elim_grain):
        pp, pq, x, y = self.pp, self.pq, self.x['joint'], self.y['joint']
        forced = []
        touched = torch.unique(pp.T[edges].view(-1))
        before = {}
        for p in touched:
            x[p, :2] -= y[p] / JOINT_SCALE
            before[int(p)] = x[p, :2]                                  # a view: it follows later moves (as in the reference)
        for k in range(len(edges)):
            p1, p2 = pp.T[edges][k]
            if p1 not in self.active_joints or p2 not in self.active_joints:
                continue
            at_q1, at_q2 = _where(pq[0] == p1), _where(pq[0] == p2)
            q1, q2 = pq[1][at_q1], pq[1][at_q2]
            at_n1, at_n2 = _where((pp[0] == p1) & (pp[1] != p2)), _where((pp[0] == p2) & (pp[1] != p1))
            n1, n2 = pp[1][at_n1], pp[1][at_n2]
            shared1 = sum(q1 == g for g in q2)
            shared2 = sum(q2 == g for g in q1)
            grow1 = q1[(1 - shared1).nonzero(as_tuple=True)]           # grain of p1 only: p2's new neighbour
            grow2 = q2[(1 - shared2).nonzero(as_tuple=True)]
            shrink_a, shrink_b = q1[shared1.nonzero(as_tuple=True)]
            slots1 = [at_q1[i] for i in range(3) if q1[i] == shrink_a] + [at_q1[i] for i in range(3) if q1[i] == shrink_b]
            slots2 = [at_q2[i] for i in range(3) if q2[i] == shrink_a] + [at_q2[i] for i in range(3) if q2[i] == shrink_b]
            # order the two other neighbours of each end: first the one on the side of shrink_a
            if self.has_pq(n1[0], shrink_a):
                n1, at_n1 = [n1[0], n1[1]], [at_n1[0], at_n1[1]]
            else:
                n1, at_n1 = [n1[1], n1[0]], [at_n1[1], at_n1[0]]
            if self.has_pq(n2[0], shrink_a):
                n2, at_n2 = [n2[0], n2[1]], [at_n2[0], at_n2[1]]
            else:
                n2, at_n2 = [n2[1], n2[0]], [at_n2[1], at_n2[0]]
            a1, b1 = n1
            a2, b2 = n2
            if elim_grain is None and (a1 == a2 or b1 == b2):
                continue
            if a1 == a2 and shrink_a != elim_grain:
                forced.append(shrink_a)
            if b1 == b2 and shrink_b != elim_grain:
                forced.append(shrink_b)
            # both ends collapse onto the midpoint (models.py:989-996)
            x1, x2 = x[p1, :2], x[p2, :2]
            mid = 0.5 * (x1 + _wrap_to(x2, x1))
            x[p1, :2], x[p2, :2] = mid, _wrap_to(mid, x2)
            swap = _inside(x[p2, :2], x[p1, :2], x[a1, :2], x[a2, :2])
            ahead = torch.unique(pp.T[edges][k:].view(-1))
            if a2 in ahead and b2 not in ahead:
                swap = False
            if b2 in ahead and a2 not in ahead:
                swap = True
            if a1 in ahead and b1 not in ahead:
                swap = True
            if b1 in ahead and a1 not in ahead:
                swap = False
            if swap:
                slots1.reverse(); slots2.reverse(); at_n1.reverse(); at_n2.reverse()
                a1, b1 = b1, a1
                a2, b2 = b2, a2
            pq[1, slots1[1]] = grow2
            pq[1, slots2[0]] = grow1
            pp[0, at_n1[1]] = p2
            pp[0, at_n2[0]] = p1
            pp[1][_where((pp[0] == a2) & (pp[1] == p2))] = p1
            pp[1][_where((pp[0] == b1) & (pp[1] == p1))] = p2
        for p in touched:
            y[p] = JOINT_SCALE * (x[p, :2] - before[int(p)])
            x[p, 6:8] = y[p]
        return forced


def topology_update(x_dict, edge_index_dict, y_dict, mask, active_grains, active_joints, threshold=0.6):
    """models.py:614-845 with nucleation_prob = 0.  Mutates x_dict / y_dict / mask like the reference; returns
    (x_dict, new edge_index_dict, switching_list)."""
    t = _Topology(x_dict, {k: v.clone() for k, v in edge_index_dict.items()}, y_dict, mask, active_grains, active_joints)
    prob = torch.sigmoid(y_dict['edge_event'])
    L1 = _where((prob > threshold) & (t.pp[0] < t.pp[1]))                                   # :627-629
    unexpected = []
    for grain in y_dict['grain_event']:                                                     # :638-727
        if grain not in active_grains:
            continue
        around = t.joints_of(grain)
        if any(p not in active_joints for p in around) or len(around) == 0:
            continue
        sides, across = [], []
        for p1, p2 in itertools.combinations(around, 2):                                    # torch.combinations order
            if p1 > p2:
                p1, p2 = p2, p1
            at = _where((t.pp[0] == p1) & (t.pp[1] == p2))
            if len(at) > 0:
                sides.append(at)
                g1 = t.pq[1][_where((t.pq[0] == p1) & (t.pq[1] != grain))]
                g2 = t.pq[1][_where((t.pq[0] == p2) & (t.pq[1] != grain))]
                if g1[0] in g2:
                    across.append(g1[0])
                elif g1[1] in g2:
                    across.append(g1[1])
                else:
                    raise KeyError
        sides, across = torch.cat(sides), torch.tensor(across)
        assert len(across) == len(around)
        if len(torch.unique(across)) != len(across):
            continue
        _, order = torch.sort(y_dict['grain'][across, 0])
        sides = sides[order[:-2]]                                                           # all but the two sides that survive
        forced = t.switch(sides, elim_grain=grain)
        unexpected.extend(forced)
        for g in [grain] + forced:
            t.delete_grain(g)
        for e in sides:
            if e in L1:
                L1 = L1[L1 != e]
        t.delete_two_sided()
    _, order = torch.sort(prob[L1], dim=0, descending=True, stable=True)                     # :730-731 (ties: unspecified in the reference; edge order here)
    L1 = L1[order]
    for e in L1:
        if t.pp[0, e] == -1:
            L1 = L1[L1 != e]
    t.switch(L1, elim_grain=None)
    switching_list = t.pp.T[L1]
    unexpected.extend(t.delete_two_sided())
    if len(unexpected) > 0:
        y_dict['grain_event'] = torch.cat([y_dict['grain_event'], torch.tensor(unexpected)])
    keep_pq, keep_pp = _where(t.pq[0] != -1), _where(t.pp[0] != -1)                          # cleanup :846-862
    out = {JJ: t.pp[:, keep_pp], JG: t.pq[:, keep_pq]}
    out[GJ] = torch.flip(out[JG], dims=[0])                                                  # :841
    return x_dict, out, switching_list
