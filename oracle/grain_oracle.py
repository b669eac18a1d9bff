"""CPU oracle for the GrainGNN rollout message-passing path.  TEST INFRASTRUCTURE ONLY.

A pure-torch restatement, in the reference's own op order, of the functions in SURVEY.md §8(a).
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this module; the product (`graingraphnn_b200/`) never does and has no CPU fallback.

Parity status: the arithmetic of this path lives in un-vendored third-party wheels
(torch-geometric 2.1.0, torch-scatter 2.1.0, torch-sparse 0.6.15; pinned in /root/reference/README.md:24-25)
that are absent from this image, and the reference ships no tests or golden NN outputs.  The oracle is
therefore pinned two ways (tests/test_oracle_golden.py): (1) against golden vectors produced by the
reference's OWN periodGATconv.py / heteropgclstm.py / heterogclstm.py / models.py imported unmodified from
/root/reference and executed on `oracle/pyg_stub` (our restatement of the PyG 2.1.0 surface they call) by
`oracle/make_golden.py`; (2) against the KATs the reference does hold: parameter counts 1,204,612 /
1,204,806 (model/regressor0_logfile:40, model/classifier1_logfile:40) and the pickled edge lengths
(graphs/40_40/*.pkl `edge_weight_dicts`).  Against real PyG wheels it is "parity unpinned".

Every function works in the dtype of its inputs (fp32 for timing / parity, fp64 for error analysis) and
takes weights as a flat `state_dict` with the reference's key names.
"""
import math
from typing import Dict, Tuple

import numpy as np
import torch
import torch.nn.functional as F

EdgeType = Tuple[str, str, str]
GATES = ('i', 'f', 'c', 'o')


def _key(edge_type: EdgeType) -> str:
    return '__'.join(edge_type)  # PyG HeteroConv ModuleDict key


def _lin(sd, prefix, x):
    """PyG Linear: weight is [out, in]  (periodGATconv.py:119-131)."""
    return F.linear(x, sd[prefix + '.weight'], sd.get(prefix + '.bias'))


def segment_softmax(s, index, n):
    """PyG 2.1.0 utils.softmax as called at periodGATconv.py:227: exp(s-max)/(sum+1e-16)."""
    idx = index.view(-1, 1).expand_as(s)
    m = torch.full((n, s.shape[1]), float('-inf'), dtype=s.dtype).scatter_reduce(0, idx, s, 'amax', include_self=True)
    m = torch.where(torch.isinf(m), torch.zeros_like(m), m)
    p = (s - m.index_select(0, index)).exp()
    den = torch.zeros((n, s.shape[1]), dtype=s.dtype).index_add_(0, index, p)
    return p / (den.index_select(0, index) + 1e-16)


def period_conv(sd, prefix, x_src, x_dst, edge_index, edge_attr, weighted=True):
    """One PeriodConv call (heads=1, concat, root_weight, no beta).

    periodGATconv.py:157-201 (forward) and :204-236 (message); `weighted=False` is the
    periodconv.py:235 variant (softmax computed but not applied)."""
    src, dst = edge_index[0], edge_index[1]
    n_dst, C = x_dst.shape[0], sd[prefix + '.lin_l2.weight'].shape[0]
    x_j, x_i = x_src.index_select(0, src), x_dst.index_select(0, dst)       # propagate: j=source, i=target
    rel = x_j[:, :3] - x_i[:, :3]                                           # :209
    reloc = -1 * (rel > 0.5) + 1 * (rel < -0.5) + rel                       # :210
    x_j = torch.cat([reloc, x_j[:, 3:]], dim=1)                             # :211
    query = _lin(sd, prefix + '.lin_query', x_i)                            # :216
    key = _lin(sd, prefix + '.lin_key', x_j)                                # :217
    value = _lin(sd, prefix + '.lin_l2', F.relu(_lin(sd, prefix + '.lin_value', x_j)))  # :218
    e = F.linear(edge_attr, sd[prefix + '.lin_edge.weight'])                # :222 (no bias)
    key = key + e                                                           # :224
    alpha = (query * key).sum(dim=-1, keepdim=True) / math.sqrt(C)          # :226
    alpha = segment_softmax(alpha, dst, n_dst)                              # :227
    out = value + e                                                         # :231-233
    if weighted:
        out = out * alpha                                                   # :235
    out = torch.zeros((n_dst, C), dtype=out.dtype).index_add_(0, dst, out)  # aggr='add'
    return out + _lin(sd, prefix + '.lin_skip', x_dst)                      # :186,:192


def hetero_conv(sd, prefix, xin, edge_index_dict, edge_attr_dict, weighted=True):
    """PyG HeteroConv(aggr='sum'): dict order, (x_src, x_dst), stack(...).sum(0) per dst type."""
    outs: Dict[str, list] = {}
    for et, ei in edge_index_dict.items():
        s, _, d = et
        o = period_conv(sd, f'{prefix}.convs.{_key(et)}', xin[s], xin[d], ei, edge_attr_dict[et], weighted)
        outs.setdefault(d, []).append(o)
    return {d: (v[0] if len(v) == 1 else torch.stack(v, 0).sum(0)) for d, v in outs.items()}


def pgclstm_cell(sd, prefix, x_dict, edge_index_dict, edge_attr_dict, h=None, c=None, weighted=True):
    """HeteroPGCLSTM.forward, heteropgclstm.py:148-183 (gates :111-146).  No peephole; o uses OLD h."""
    C = sd[f'{prefix}.b_i.{next(iter(x_dict))}'].shape[1]
    if h is None:
        h = {t: torch.zeros(x.shape[0], C, dtype=x.dtype) for t, x in x_dict.items()}   # :101-104
    if c is None:
        c = {t: torch.zeros(x.shape[0], C, dtype=x.dtype) for t, x in x_dict.items()}   # :106-109
    xin = {t: torch.cat([x, h[t]], dim=1) for t, x in x_dict.items()}                   # :112
    pre = {}
    for g in GATES:
        conv = hetero_conv(sd, f'{prefix}.conv_{g}', xin, edge_index_dict, edge_attr_dict, weighted)
        pre[g] = {t: conv[t] + sd[f'{prefix}.b_{g}.{t}'] for t in x_dict}
    i = {t: torch.sigmoid(pre['i'][t]) for t in x_dict}
    f = {t: torch.sigmoid(pre['f'][t]) for t in x_dict}
    tt = {t: torch.tanh(pre['c'][t]) for t in x_dict}
    c2 = {t: f[t] * c[t] + i[t] * tt[t] for t in x_dict}                                # :133
    o = {t: torch.sigmoid(pre['o'][t]) for t in x_dict}
    h2 = {t: o[t] * torch.tanh(c2[t]) for t in x_dict}                                  # :145
    return h2, c2


def pgc_cell(sd, prefix, x_dict, edge_index_dict, edge_attr_dict, h=None, c=None, weighted=True):
    """HeteroPGC.forward, heteropgclstm.py:243-284: relu(conv_i([X,h]) + b_i); c passes through."""
    C = sd[f'{prefix}.b_i.{next(iter(x_dict))}'].shape[1]
    if h is None:
        h = {t: torch.zeros(x.shape[0], C, dtype=x.dtype) for t, x in x_dict.items()}
    if c is None:
        c = {t: torch.zeros(x.shape[0], C, dtype=x.dtype) for t, x in x_dict.items()}
    xin = {t: torch.cat([x, h[t]], dim=1) for t, x in x_dict.items()}
    conv = hetero_conv(sd, f'{prefix}.conv_i', xin, edge_index_dict, edge_attr_dict, weighted)
    return {t: torch.relu(conv[t] + sd[f'{prefix}.b_i.{t}']) for t in x_dict}, c


def sage_conv(sd, prefix, x_src, x_dst, edge_index):
    """PyG 2.1.0 SAGEConv defaults as built at heterogclstm.py:52-54: lin_l(mean_j x_j) + lin_r(x_i)."""
    src, dst = edge_index[0], edge_index[1]
    n = x_dst.shape[0]
    s = torch.zeros((n, x_src.shape[1]), dtype=x_src.dtype).index_add_(0, dst, x_src.index_select(0, src))
    cnt = torch.zeros(n, dtype=x_src.dtype).index_add_(0, dst, torch.ones(dst.shape[0], dtype=x_src.dtype))
    mean = s / cnt.clamp(min=1).view(-1, 1)
    return _lin(sd, prefix + '.lin_l', mean) + F.linear(x_dst, sd[prefix + '.lin_r.weight'])


def gclstm_cell(sd, prefix, x_dict, edge_index_dict, h=None, c=None):
    """HeteroGCLSTM.forward, heterogclstm.py:162-196 (gates :125-160); W_* params are never read."""
    C = sd[f'{prefix}.b_i.{next(iter(x_dict))}'].shape[1]
    if h is None:
        h = {t: torch.zeros(x.shape[0], C, dtype=x.dtype) for t, x in x_dict.items()}
    if c is None:
        c = {t: torch.zeros(x.shape[0], C, dtype=x.dtype) for t, x in x_dict.items()}
    xin = {t: torch.cat([x, h[t]], dim=1) for t, x in x_dict.items()}
    pre = {}
    for g in GATES:
        outs: Dict[str, list] = {}
        for et, ei in edge_index_dict.items():
            s, _, d = et
            outs.setdefault(d, []).append(sage_conv(sd, f'{prefix}.conv_{g}.convs.{_key(et)}', xin[s], xin[d], ei))
        pre[g] = {d: (v[0] if len(v) == 1 else torch.stack(v, 0).sum(0)) + sd[f'{prefix}.b_{g}.{d}']
                  for d, v in outs.items()}
    i = {t: torch.sigmoid(pre['i'][t]) for t in x_dict}
    f = {t: torch.sigmoid(pre['f'][t]) for t in x_dict}
    c2 = {t: f[t] * c[t] + i[t] * torch.tanh(pre['c'][t]) for t in x_dict}
    o = {t: torch.sigmoid(pre['o'][t]) for t in x_dict}
    return {t: o[t] * torch.tanh(c2[t]) for t in x_dict}, c2


def encode_decode(sd, x_dict, edge_index_dict, edge_attr_dict, weighted=True):
    """SeqGCLSTM encoder(None) -> decoder(enc state), layers=1, seq_len=1  (models.py:219-289, :422-424)."""
    h, c = pgclstm_cell(sd, 'gclstm_encoder.cell_list.0', x_dict, edge_index_dict, edge_attr_dict, None, None, weighted)
    return pgclstm_cell(sd, 'gclstm_decoder.cell_list.0', x_dict, edge_index_dict, edge_attr_dict, h, c, weighted)


def regressor_forward(sd, x_dict, edge_index_dict, edge_attr_dict, return_state=False):
    """GrainNN_regressor.forward, models.py:401-467 (history=False, edge_len=False)."""
    h, c = encode_decode(sd, x_dict, edge_index_dict, edge_attr_dict)
    y = {t: F.linear(h[t], sd[f'linear.{t}.weight'], sd[f'linear.{t}.bias']) for t in h}   # :433
    y['joint'] = torch.tanh(y['joint'])                                                    # :443
    y['grain_area'] = torch.tanh(y['grain'][:, 0]) / 20 + x_dict['grain'][:, 3]            # :445
    y['grain'][:, 0] = torch.tanh(y['grain'][:, 0])                                        # :450
    y['grain'][:, 1] = F.relu(y['grain'][:, 1])                                            # :452
    return (y, h, c) if return_state else y


def classifier_forward(sd, x_dict, edge_index_dict, edge_attr_dict, return_state=False):
    """GrainNN_classifier.forward, models.py:572-611 (history=False). Outputs in ORIGINAL jj edge order."""
    h, c = encode_decode(sd, x_dict, edge_index_dict, edge_attr_dict)
    jj = ('joint', 'connect', 'joint')
    src, dst = edge_index_dict[jj][0], edge_index_dict[jj][1]
    pair = torch.cat([h['joint'][src], h['joint'][dst], edge_attr_dict[jj]], dim=-1)       # :602
    y = {'edge_event': F.linear(pair, sd['lin2.weight'], sd['lin2.bias']).view(-1)}        # :607
    y['edge'] = torch.tanh(F.linear(pair, sd['lin1.weight'], sd['lin1.bias']))             # :609
    return (y, h, c) if return_state else y


def regressor_update(x_dict, y_dict, span=6, train_frames=120):
    """Feature update, in place: models.py:503-516 (periodic branch) + test.py:401-407."""
    x_dict['joint'][:, :2] += y_dict['joint'] / 5
    x_dict['grain'][:, 3] += y_dict['grain'][:, 0] / 20
    x_dict['grain'][:, 4] = y_dict['grain'][:, 1]
    x_dict['joint'][:, 6:8] = y_dict['joint']
    x_dict['grain'][:, -1] = y_dict['grain'][:, 0]
    x_dict['grain'][:, 2] += span / (train_frames + 1)
    x_dict['joint'][:, 2] += span / (train_frames + 1)
    if x_dict['grain'][0, 2] > train_frames / (train_frames + 1):
        x_dict['grain'][:, 2] = train_frames / (train_frames + 1)
        x_dict['joint'][:, 2] = train_frames / (train_frames + 1)


def edge_attr_rebuild(x_dict, edge_index_dict):
    """Wrapped 2-D edge length per edge type, test.py:562-575."""
    out = {}
    for et, index in edge_index_dict.items():
        rel = x_dict[et[0]][index[0], :2] - x_dict[et[-1]][index[-1], :2]
        rel = -1 * (rel > 0.5) + 1 * (rel < -0.5) + rel
        out[et] = torch.sqrt(rel[:, 0] ** 2 + rel[:, 1] ** 2).view(-1, 1)
    return out


def nn_step(sd_r, sd_c, x_dict, edge_index_dict, edge_attr_dict, span=6):
    """One fixed-topology rollout step ("nn-step"): test.py:382-383, :400-407, :562-575.
    Mutates x_dict in place, returns (pred, new edge_attr_dict)."""
    pred = regressor_forward(sd_r, x_dict, edge_index_dict, edge_attr_dict)
    pred.update(classifier_forward(sd_c, x_dict, edge_index_dict, edge_attr_dict))
    regressor_update(x_dict, pred, span)
    return pred, edge_attr_rebuild(x_dict, edge_index_dict)


def region_center(x_joint, gj_edge_index, n_grain, joint_offset=None, domain_factor=1):
    """Grain centres from the joint positions — SURVEY.md §8 row f2, the geometry feedback of a rollout step.

    Restates, for the periodic boundary, test.py:471-476 (patch -> global joint coordinates in fp32),
    graph_trajectory.py:1036-1039 (vertices = fp32 joint rows), :1062-1075 (vertex2joint / joint2vertex: joints in the
    order of their first appearance as a target of the grain->joint edges; grains are 1-based regions; a later joint with
    the same grain triple takes the place of the earlier one, as the reference's dict does) and
    graph_datastruct.py:672-708 (per region: chain of periodic_move :55-72 against the PREVIOUS moved vertex, shift by
    +1 along an axis where any vertex is <= -eps, np.mean).  The numpy scalar types of the reference are kept: the first
    vertex stays float32 (and its +1 shift rounds in float32), later vertices become float64 through `x += int64`.
    Returns centers float64 [n_grain, 2] with NaN rows for grains the reference skips (<= 1 vertex)."""
    eps = 1e-12                                                            # graph_datastruct.py:36
    xj = x_joint.detach().cpu() if isinstance(x_joint, torch.Tensor) else torch.as_tensor(np.asarray(x_joint))
    X = xj[:, :2].clone().float()
    if domain_factor > 1:                                                  # test.py:472-474
        X = (X + torch.as_tensor(np.asarray(joint_offset), dtype=torch.float32)) / domain_factor
    X_j = X.numpy()
    gj = gj_edge_index.cpu().numpy() if isinstance(gj_edge_index, torch.Tensor) else np.asarray(gj_edge_index)
    vertex2joint = {}
    for grain, joint in gj.T:                                              # graph_trajectory.py:1062-1064
        vertex2joint.setdefault(int(joint), set()).add(int(grain) + 1)
    joint2vertex = dict((tuple(sorted(v)), k) for k, v in vertex2joint.items())    # :1080
    region_coors = {}
    for k, v in joint2vertex.items():                                      # graph_datastruct.py:672-678
        for region in set(k):
            region_coors.setdefault(region, []).append(X_j[v])
    centers = np.full((n_grain, 2), np.nan)
    for region, verts in region_coors.items():                             # :682-708
        if len(verts) <= 1:
            continue
        for i in range(1, len(verts)):
            x, y = verts[i]
            xc, yc = verts[i - 1]
            rel_x, rel_y = x - xc, y - yc
            x += -1 * (rel_x > 0.5) + 1 * (rel_x < -0.5)                  # np.float32 += np.int64 -> np.float64
            y += -1 * (rel_y > 0.5) + 1 * (rel_y < -0.5)
            verts[i] = [x, y]
        inbound = [True, True]
        for vert in verts:
            inbound = [i and (j > -eps) for i, j in zip(inbound, vert)]
        moved = [[i + 1 * (not j) for i, j in zip(vert, inbound)] for vert in verts]
        x, y = zip(*moved)
        centers[region - 1] = [np.mean(x), np.mean(y)]
    return centers


def area_bookkeeping(x_grain, mask_grain, gj_edge_index, lxd, patch_size=40, mesh_size=0.08, v_scale=20):
    """graph_trajectory.py:1041-1051 (`area_counts`, `extraV_traj`) and :1100-1103 (`vertex_area`), in the reference's numpy
    arithmetic.  -> (area_counts {grain id (1-based): value}, extraV [Ng], vertex_area {joint: value})."""
    import numpy as np
    from collections import defaultdict
    X_g = x_grain[:, 3:5].detach().numpy()
    mask_g = mask_grain.detach().numpy().reshape(mask_grain.shape[0], -1)[:, 0]
    s = (patch_size / mesh_size) + 1
    area_counts = {}
    area_sum = np.sum(X_g[:, 0] * mask_g) / (lxd / patch_size) ** 2
    for idx, area in enumerate(X_g[:, 0]):
        if mask_g[idx] > 0:
            area_counts[idx + 1] = area * s ** 2 / area_sum
    extra = mask_g * X_g[:, 1] / v_scale * s ** 3
    gj = gj_edge_index.numpy()
    regions = defaultdict(list)
    seen = set()
    for g, j in gj.T:
        if (int(g), int(j)) not in seen:
            seen.add((int(g), int(j)))
            regions[int(g) + 1].append(int(j))
    vertex_area = defaultdict(float)
    for region, verts in regions.items():
        for v in verts:
            vertex_area[v] += area_counts[region] / len(verts) * mesh_size ** 2
    return area_counts, extra, vertex_area


def grain_xy_writeback(x_grain, centers, domain_factor=1):
    """test.py:556-559: grain (x, y) <- fp32(region centre), `(.. * domain_factor) % 1` on scaled patches.  In place."""
    for g in range(centers.shape[0]):
        if not np.isnan(centers[g, 0]):
            x_grain[g, :2] = torch.FloatTensor(centers[g])
            if domain_factor > 1:
                x_grain[g, :2] = (x_grain[g, :2] * domain_factor) % 1
    return x_grain


def event_candidates(y_dict, jj_edge_index, mask_grain=None, edge_threshold=0.6, area_threshold=1e-4):
    """Event candidates exactly as the reference's host code selects them: models.py:627-629 (edges that may switch) and
    test.py:414-416 (grains that vanish, sorted by predicted area)."""
    src, dst = jj_edge_index[0], jj_edge_index[1]
    prob = torch.sigmoid(y_dict['edge_event'])
    L1 = ((prob > edge_threshold) & (src < dst)).nonzero().view(-1)
    alive = torch.ones_like(y_dict['grain_area'], dtype=torch.bool) if mask_grain is None else (mask_grain.reshape(len(mask_grain), -1)[:, 0] > 0)
    grain_event = (alive & (y_dict['grain_area'] < area_threshold)).nonzero().view(-1)
    grain_event = grain_event[torch.argsort(y_dict['grain_area'][grain_event])]
    return L1, grain_event


def csr_by_dst(edge_index, n_dst):
    """Stable dst-sorted CSR (the index structure kernel (a) must reproduce bit-exactly).
    rowptr[n_dst+1], col[E] = src in (dst, original-edge-id) order, perm[E] = original edge id."""
    ei = edge_index.numpy() if isinstance(edge_index, torch.Tensor) else np.asarray(edge_index)
    perm = np.argsort(ei[1], kind='stable')
    rowptr = np.zeros(n_dst + 1, dtype=np.int64)
    np.cumsum(np.bincount(ei[1], minlength=n_dst), out=rowptr[1:])
    return rowptr.astype(np.int32), ei[0][perm].astype(np.int32), perm.astype(np.int32)


# ----------------------------------------------------------------------------- weights
def param_shapes(kind='regressor', C=96, f_grain=11, f_joint=8,
                 edge_types=(('grain', 'push', 'joint'), ('joint', 'pull', 'grain'), ('joint', 'connect', 'joint'))):
    """Reference state_dict layout in registration order -> {key: shape}."""
    D = {'grain': f_grain + C, 'joint': f_joint + C}
    shapes = {}
    for part in ('gclstm_encoder', 'gclstm_decoder'):
        p = f'{part}.cell_list.0'
        for g in GATES:
            for et in edge_types:
                q = f'{p}.conv_{g}.convs.{_key(et)}'
                s, d = D[et[0]], D[et[2]]
                shapes[q + '.lin_key.weight'] = (C, s); shapes[q + '.lin_key.bias'] = (C,)
                shapes[q + '.lin_query.weight'] = (C, d); shapes[q + '.lin_query.bias'] = (C,)
                shapes[q + '.lin_value.weight'] = (C, s); shapes[q + '.lin_value.bias'] = (C,)
                shapes[q + '.lin_l2.weight'] = (C, C); shapes[q + '.lin_l2.bias'] = (C,)
                shapes[q + '.lin_edge.weight'] = (C, 1)
                shapes[q + '.lin_skip.weight'] = (C, d); shapes[q + '.lin_skip.bias'] = (C,)
            for t in ('grain', 'joint'):
                shapes[f'{p}.b_{g}.{t}'] = (1, C)
    if kind == 'regressor':
        for t in ('grain', 'joint'):
            shapes[f'linear.{t}.weight'] = (2, C); shapes[f'linear.{t}.bias'] = (2,)
    else:
        shapes['lin1.weight'] = (2, 2 * C + 1); shapes['lin1.bias'] = (2,)
        shapes['lin2.weight'] = (1, 2 * C + 1); shapes['lin2.bias'] = (1,)
    return shapes


def synth_state_dict(kind='regressor', seed=0, dtype=torch.float32, gain=1.0, **kw):
    """Deterministic stand-in weights with the reference's exact keys/shapes (the shipped .pt files are
    absent, .MISSING_LARGE_BLOBS:2-3).  U(-g/sqrt(fan_in), g/sqrt(fan_in)) like PyG/torch Linear defaults."""
    gen = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in param_shapes(kind, **kw).items():
        fan_in = shp[-1] if len(shp) == 2 and not k.split('.')[-2].startswith('b_') else shp[-1]
        bound = gain / math.sqrt(max(fan_in, 1))
        if k.endswith('lin_edge.weight'):
            bound = gain
        sd[k] = ((torch.rand(shp, generator=gen, dtype=torch.float64) * 2 - 1) * bound).to(dtype)
    return sd
