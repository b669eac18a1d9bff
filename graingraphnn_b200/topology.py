"""Host-side topology update of a rollout step with O(1) lookups.

Same decisions, same edge arrays (position for position), same moved joints as the reference's
`GrainNN_classifier.update` (models.py:614-845, nucleation included), `switching_edge_index` (:899-1053),
`delete_grain_index` (:864-896) and `cleanup` (:846-862) — but every `(E == p).nonzero()` scan of the reference
(O(E) each, O(events x E) per step: the step-time floor at >= 10^5 grains) is answered from position lists kept per joint
and per grain, so a step costs O(E) once for the index plus O(1) per lookup.  The host stays Python, as in the reference;
the edge arrays keep the reference's in-place discipline (edits overwrite positions, new edges are appended, -1 marks
deleted rows until the final stable compaction), which is what makes the result comparable array for array
(tests/test_topology_golden.py: the reference's own outputs, and the O(E)-scan oracle on denser event sets).
Input candidates may come from the device (`EventSelector.fetch()`), or are selected here like models.py:627-629.
"""
import bisect
import itertools

import numpy as np
import torch

JJ, JG, GJ = ('joint', 'connect', 'joint'), ('joint', 'pull', 'grain'), ('grain', 'push', 'joint')
JOINT_SCALE = 5          # models.py:546


_F0, _F1, _HALF = np.float32(0.0), np.float32(1.0), np.float32(0.5)


def _wrap_to(p, pc):
    """periodic_move (models.py:1097-1100) on float32 numpy rows: p - 1*(rel > .5) + 1*(rel < -.5), every step in float32
    (torch promotes float32 - int64 to float32; numpy would go to float64, hence the explicit casts)."""
    rel = p - pc
    return (p - (rel > _HALF).astype(np.float32)) + (rel < -_HALF).astype(np.float32)


def _inside(t, v1, v2, v3):
    def sign(a, b, c):                                     # point_in_triangle, models.py:1055-1072 (float32 scalars)
        return (a[0] - c[0]) * (b[1] - c[1]) - (b[0] - c[0]) * (a[1] - c[1])
    a, b, c = _wrap_to(v1, t), _wrap_to(v2, t), _wrap_to(v3, t)
    d = (sign(t, a, b), sign(t, b, c), sign(t, c, a))
    neg = bool(d[0] < 0) or bool(d[1] < 0) or bool(d[2] < 0)
    pos = bool(d[0] > 0) or bool(d[1] > 0) or bool(d[2] > 0)
    return not (neg and pos)


class _Rows:
    """One [2, E] edge array with, per endpoint value, the ascending list of positions holding it (what `.nonzero()` on
    the reference's masks returns).  Row 0 / row 1 are indexed separately."""

    def __init__(self, edges, n0, n1):
        self.a = edges.numpy().astype(np.int64).copy()
        self.n = (n0, n1)
        # positions grouped by value: one stable argsort per row, made when the row is first queried (from the array as it is
        # THEN — edits before that need no bookkeeping); a value's Python list is made on first use and kept current from
        # then on, so a step touches O(events) lists, not O(N)
        self.order, self.starts, self.lists = [None, None], [None, None], ({}, {})

    def _index(self, r):
        order = np.argsort(self.a[r], kind='stable')
        self.order[r] = order
        # sized from the values actually present: ids appended after construction (nucleation) are indexed like any other
        top = max(self.n[r], int(self.a[r].max()) + 1 if self.a.shape[1] else 0)
        self.starts[r] = np.searchsorted(self.a[r][order], np.arange(top + 1))

    def at(self, r, v):
        v = int(v)
        if self.order[r] is None:
            self._index(r)
        hit = self.lists[r].get(v)
        if hit is None:
            hit = self.lists[r][v] = self.order[r][self.starts[r][v]:self.starts[r][v + 1]].tolist() if v + 1 < len(self.starts[r]) else []
        return hit

    def counts(self, r):
        """Occurrences of every value in row r of the CURRENT array (vectorised)."""
        vals = self.a[r]
        return np.bincount(vals[vals >= 0], minlength=self.n[r])

    def get(self, r, pos):
        return int(self.a[r, pos])

    def set(self, r, pos, v):
        old, v = int(self.a[r, pos]), int(v)
        if old == v:
            return
        if self.order[r] is not None:                       # row not indexed yet: the index will read the edited array
            if old >= 0:
                self.at(r, old).remove(pos)
            if v >= 0:
                bisect.insort(self.at(r, v), pos)
        self.a[r, pos] = v

    def kill(self, pos):
        self.set(0, pos, -1)
        self.set(1, pos, -1)

    def append(self, v0, v1):
        pos = self.a.shape[1]
        self.a = np.concatenate([self.a, np.array([[-1], [-1]], dtype=np.int64)], axis=1)
        self.set(0, pos, v0)
        self.set(1, pos, v1)

    def compact(self):
        return torch.from_numpy(self.a[:, self.a[0] != -1].copy())


class _Surgery:
    def __init__(self, x_dict, edge_index_dict, y_dict, mask, active_grains, active_joints):
        nj, ng = x_dict['joint'].shape[0], x_dict['grain'].shape[0]
        self.x, self.y, self.mask = x_dict, y_dict, mask
        self.pp = _Rows(edge_index_dict[JJ], nj, nj)
        self.pq = _Rows(edge_index_dict[JG], nj, ng)
        self.act_g = np.zeros(ng, dtype=bool)
        self.act_g[np.asarray(active_grains)] = True
        self.act_j = np.zeros(nj, dtype=bool)
        self.act_j[np.asarray(active_joints)] = True
        self.dirty = None                                   # grains whose joint count changed since the last two-side check

    def pp_between(self, p1, p2, equal=True):
        """positions e with pp[0,e] == p1 and pp[1,e] == p2 (or != p2), ascending."""
        return [e for e in self.pp.at(0, p1) if (self.pp.get(1, e) == p2) == equal]

    def pq_set_grain(self, pos, g):
        for v in (self.pq.get(1, pos), int(g)):
            if v >= 0 and self.dirty is not None:
                self.dirty.add(v)
        self.pq.set(1, pos, g)

    def delete_grain(self, grain):                          # models.py:864-896
        grain = int(grain)
        around = [self.pq.get(0, e) for e in self.pq.at(1, grain)]
        assert len(around) == 2, around
        p1, p2 = around
        n1 = self.pp.get(1, self.pp_between(p1, p2, equal=False)[0])
        n2 = self.pp.get(1, self.pp_between(p2, p1, equal=False)[0])
        self.pp.append(n1, n2)
        self.pp.append(n2, n1)
        self.mask['grain'][grain] = 0
        self.mask['joint'][p1] = 0
        self.mask['joint'][p2] = 0
        for e in list(self.pq.at(1, grain)):
            self.pq.kill(e)
        for j in (p1, p2):
            for e in list(self.pq.at(0, j)):
                if self.dirty is not None:
                    self.dirty.add(self.pq.get(1, e))
                self.pq.kill(e)
            for e in list(self.pp.at(0, j)) + list(self.pp.at(1, j)):
                self.pp.kill(e)

    def delete_two_sided(self):                             # models.py:716-727 / :745-755
        if self.dirty is None:                              # first check of the step: every grain (torch.unique over E_pq[1])
            cnt = self.pq.counts(1)
            cand = np.nonzero((cnt > 0) & (cnt <= 2))[0].tolist()
        else:
            cand = sorted(g for g in self.dirty if 0 < len(self.pq.at(1, g)) <= 2)
        self.dirty = set()
        for g in cand:
            self.delete_grain(g)
        return cand

    def switch(self, edges, elim_grain):                    # models.py:899-1053
        # joint rows as float32 numpy views of the torch tensors (same memory): IEEE single arithmetic like torch's, without
        # the per-op overhead of 0-d tensors
        pp, pq, x, y = self.pp, self.pq, self.x['joint'].numpy(), self.y['joint'].numpy()
        assert x.dtype == np.float32 and y.dtype == np.float32
        forced = []
        edges = [int(e) for e in edges]
        edges_arr = np.asarray(edges, dtype=np.int64)
        touched = sorted({pp.get(r, e) for e in edges for r in (0, 1)})
        before = {}
        for p in touched:
            x[p, :2] -= y[p] / np.float32(JOINT_SCALE)
            before[p] = x[p, :2]                            # a view: it follows later moves (as in the reference)
        for k, e in enumerate(edges):
            p1, p2 = pp.get(0, e), pp.get(1, e)
            if not (self.act_j[p1] and self.act_j[p2]):
                continue
            at_q1, at_q2 = list(pq.at(0, p1)), list(pq.at(0, p2))
            q1, q2 = [pq.get(1, i) for i in at_q1], [pq.get(1, i) for i in at_q2]
            at_n1, at_n2 = self.pp_between(p1, p2, equal=False), self.pp_between(p2, p1, equal=False)
            n1, n2 = [pp.get(1, i) for i in at_n1], [pp.get(1, i) for i in at_n2]
            grow1 = [g for g in q1 if q2.count(g) != 1]
            grow2 = [g for g in q2 if q1.count(g) != 1]
            shrink_a, shrink_b = [g for g in q1 if q2.count(g) != 0]
            slots1 = [at_q1[i] for i in range(3) if q1[i] == shrink_a] + [at_q1[i] for i in range(3) if q1[i] == shrink_b]
            slots2 = [at_q2[i] for i in range(3) if q2[i] == shrink_a] + [at_q2[i] for i in range(3) if q2[i] == shrink_b]
            if not any(pq.get(1, i) == shrink_a for i in pq.at(0, n1[0])):
                n1, at_n1 = [n1[1], n1[0]], [at_n1[1], at_n1[0]]
            else:
                n1, at_n1 = n1[:2], at_n1[:2]
            if not any(pq.get(1, i) == shrink_a for i in pq.at(0, n2[0])):
                n2, at_n2 = [n2[1], n2[0]], [at_n2[1], at_n2[0]]
            else:
                n2, at_n2 = n2[:2], at_n2[:2]
            (a1, b1), (a2, b2) = n1, n2
            if elim_grain is None and (a1 == a2 or b1 == b2):
                continue
            if a1 == a2 and shrink_a != elim_grain:
                forced.append(shrink_a)
            if b1 == b2 and shrink_b != elim_grain:
                forced.append(shrink_b)
            x1, x2 = x[p1, :2], x[p2, :2]                   # both ends collapse onto the midpoint (models.py:989-996)
            mid = _HALF * (x1 + _wrap_to(x2, x1))
            x[p1, :2], x[p2, :2] = mid, _wrap_to(mid, x2)
            swap = _inside(x[p2, :2], x[p1, :2], x[a1, :2], x[a2, :2])
            ahead = set(pp.a[:, edges_arr[k:]].ravel().tolist())      # endpoints of the events still to come, as they are NOW
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
            assert len(grow1) == 1 and len(grow2) == 1      # the reference assigns 1-element tensors here
            self.pq_set_grain(slots1[1], grow2[0])
            self.pq_set_grain(slots2[0], grow1[0])
            pp.set(0, at_n1[1], p2)
            pp.set(0, at_n2[0], p1)
            for i in self.pp_between(a2, p2):
                pp.set(1, i, p1)
            for i in self.pp_between(b1, p1):
                pp.set(1, i, p2)
        for p in touched:
            y[p] = np.float32(JOINT_SCALE) * (x[p, :2] - before[p])
            x[p, 6:8] = y[p]
        return forced


def _unit_towards(p, pc):
    """periodic_norm, models.py:1102-1108."""
    rel = p - pc
    rel = rel - 1 * (rel > 0.5) + 1 * (rel < -0.5)
    return torch.nn.functional.normalize(rel, p=2.0, dim=0, eps=1e-6)


def _nucleate(s, nucleation_prob):
    """models.py:771-835: a new grain opens at every live junction drawn by `torch.rand` (same draws, in the same order, as the
    reference: one vector over the junctions, then two angles per site).  The junction keeps its neighbour 0 and becomes one
    corner of the new triangular grain; two new junctions take over neighbours 1 and 2.  x_dict / mask entries are re-bound to
    grown tensors exactly as the reference re-binds them."""
    x, mask, pp, pq = s.x, s.mask, s.pp, s.pq
    draws = torch.rand(x['joint'].size(dim=0))
    sites = ((draws < nucleation_prob) & (mask['joint'][:, 0] > 0)).nonzero().view(-1)
    n_grain, n_joint = mask['grain'].size(dim=0), mask['joint'].size(dim=0)
    for junction in sites:
        j = int(junction)
        mask['joint'] = torch.cat((mask['joint'], torch.tensor([1, 1]).view(-1, 1)))
        mask['grain'] = torch.cat((mask['grain'], torch.tensor([1]).view(-1, 1)))
        (sx, sy, sz), dz = x['joint'][j, :3], x['joint'][j, -1]
        theta_x, theta_z = torch.rand(2) * torch.pi / 2
        area = 0.004
        reach = torch.sqrt(area * 4 / 3 / torch.sqrt(torch.tensor(3)))
        grain_row = torch.tensor([sx, sy, sz, area, 0, torch.cos(theta_x), torch.sin(theta_x),
                                  torch.cos(theta_z), torch.sin(theta_z), area, dz])
        x['grain'] = torch.cat((x['grain'], grain_row.view(1, -1)), dim=0)
        j1, j2 = n_joint, n_joint + 1
        nb = [pp.get(1, e) for e in pp.at(0, j)]
        nb0, nb1, nb2 = nb
        opposite = [0, 0, 0]                                 # the grain of the junction that neighbour k does NOT touch
        for g in [pq.get(1, e) for e in pq.at(0, j)]:
            for k in range(3):
                if not any(pq.get(1, e) == g for e in pq.at(0, nb[k])):
                    opposite[k] = g
        g0, g1, g2 = opposite
        assert g0 != g1 and g1 != g2 and g0 != g2
        centre = x['joint'][j, :2].clone()
        row1, row2 = x['joint'][j].clone(), x['joint'][j].clone()
        x['joint'][j, :2] = centre + _unit_towards(x['joint'][nb0, :2], centre) * reach
        row1[:2] = centre + _unit_towards(x['joint'][nb1, :2], centre) * reach
        row2[:2] = centre + _unit_towards(x['joint'][nb2, :2], centre) * reach
        x['joint'][j, -2:] = 0
        row1[-2:] = 0
        row2[-2:] = 0
        x['joint'] = torch.cat((x['joint'], row1.view(1, -1), row2.view(1, -1)), dim=0)
        for e in list(pq.at(0, j)):
            pq.kill(e)
        for e in s.pp_between(nb1, j):
            pp.set(1, e, j1)
        for e in s.pp_between(nb2, j):
            pp.set(1, e, j2)
        for e in s.pp_between(j, nb1):
            pp.set(0, e, j1)
        for e in s.pp_between(j, nb2):
            pp.set(0, e, j2)
        for a, b in ((j, j1), (j, j2), (j1, j), (j1, j2), (j2, j), (j2, j1)):
            pp.append(a, b)
        for a, b in ((j, n_grain), (j1, n_grain), (j2, n_grain), (j1, g0), (j2, g0), (j, g1), (j2, g1), (j, g2), (j1, g2)):
            pq.append(a, b)
        n_grain += 1
        n_joint += 2


def topology_update(x_dict, edge_index_dict, y_dict, mask, active_grains, active_joints, threshold=0.6, L1=None,
                    nucleation_prob=0.0):
    """The reference's GrainNN_classifier.update on CPU tensors (nucleation_prob > 1e-6: new grains open at random junctions,
    drawn from torch's global generator like the reference's; x_dict / mask entries are then re-bound to grown tensors).  Mutates x_dict / y_dict / mask like
    the reference and returns (x_dict, new edge_index_dict, switching_list).  y_dict['grain_event']: grain ids sorted by area
    (test.py:414-416).  L1 (optional): the candidate edges, ascending ids (EventSelector.fetch()['L1']); selected here from
    y_dict['edge_event'] when absent (models.py:627-629)."""
    s = _Surgery(x_dict, edge_index_dict, y_dict, mask, active_grains, active_joints)
    prob = torch.sigmoid(y_dict['edge_event'])
    if L1 is None:
        pp0 = edge_index_dict[JJ]
        L1 = ((prob > threshold) & (pp0[0] < pp0[1])).nonzero().view(-1)
    L1 = [int(e) for e in L1]
    unexpected = []
    for grain in [int(g) for g in y_dict['grain_event']]:                                    # models.py:638-727
        if not s.act_g[grain]:
            continue
        around = [s.pq.get(0, e) for e in s.pq.at(1, grain)]
        if len(around) == 0 or not all(s.act_j[p] for p in around):
            continue
        sides, across = [], []
        for p1, p2 in itertools.combinations(around, 2):
            if p1 > p2:
                p1, p2 = p2, p1
            at = s.pp_between(p1, p2)
            if at:
                sides.extend(at)
                g1 = [s.pq.get(1, e) for e in s.pq.at(0, p1) if s.pq.get(1, e) != grain]
                g2 = [s.pq.get(1, e) for e in s.pq.at(0, p2) if s.pq.get(1, e) != grain]
                if g1[0] in g2:
                    across.append(g1[0])
                elif g1[1] in g2:
                    across.append(g1[1])
                else:
                    raise KeyError
        assert len(across) == len(around)
        if len(set(across)) != len(across):
            continue
        _, order = torch.sort(y_dict['grain'][torch.tensor(across), 0])
        sides = [sides[int(i)] for i in order[:-2]]
        forced = s.switch(sides, elim_grain=grain)
        unexpected.extend(forced)
        for g in [grain] + forced:
            s.delete_grain(g)
        L1 = [e for e in L1 if e not in sides]
        s.delete_two_sided()
    if L1:
        # models.py:730-731.  Equal probabilities (logits that saturate or collide in fp32): the reference's torch.sort is not a stable
        # one beyond 16 elements and leaves their order to the implementation; here, and in the device update, they keep edge order
        _, order = torch.sort(prob[torch.tensor(L1)], dim=0, descending=True, stable=True)
        L1 = [L1[int(i)] for i in order]
    L1 = [e for e in L1 if s.pp.get(0, e) != -1]
    s.switch(L1, elim_grain=None)
    switching_list = torch.from_numpy(s.pp.a[:, L1].T.copy()) if L1 else torch.zeros(0, 2, dtype=torch.int64)
    unexpected.extend(s.delete_two_sided())
    if unexpected:
        y_dict['grain_event'] = torch.cat([y_dict['grain_event'], torch.tensor(unexpected)])
    if nucleation_prob > 1e-6:                                                               # models.py:771-835
        _nucleate(s, nucleation_prob)
    out = {JJ: s.pp.compact(), JG: s.pq.compact()}
    out[GJ] = torch.flip(out[JG], dims=[0])                                                  # models.py:841
    return x_dict, out, switching_list
