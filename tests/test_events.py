"""Event candidates (first stage): host logic on the CPU, the selection kernel on the GPU against the
reference's host scans (models.py:627-629, test.py:414-416, restated in oracle/grain_oracle.event_candidates)."""
import numpy as np
import pytest
import torch

import grain_oracle as orc
from util import ET, load_graph

from graingraphnn_b200.events import sigmoid_threshold_band


@pytest.mark.parametrize('thr', [0.6, 0.5, 0.3, 0.9])
def test_threshold_band_brackets_every_decision_of_torch_sigmoid(thr):
    lo, hi = sigmoid_threshold_band(thr)
    ulps = int(np.float32(hi).view(np.int32)) - int(np.float32(lo).view(np.int32))
    assert -1 <= ulps <= 16                                             # a crossing (hi just below lo) or a band of a few floats
    g = torch.Generator().manual_seed(0)
    t = torch.randn(2_000_000, generator=g) * 4
    below, above = float(np.nextafter(np.float32(lo), np.float32(-np.inf))), float(np.nextafter(np.float32(max(lo, hi)), np.float32(np.inf)))
    t[:64], t[64:128], t[128:192] = lo, below, above
    t[-3:] = torch.tensor([lo, below, above])                            # vector body and scalar tail positions
    dec = torch.sigmoid(t) > thr
    assert bool(((t >= lo) | ~dec).all())                                # device test keeps a superset ...
    out = (t < lo) | (t > hi)
    assert torch.equal(dec[out], (t >= lo)[out])                         # ... and is exact outside the band
    assert not dec[64:128].any() and dec[128:192].all() and not dec[-2] and dec[-1]


def test_oracle_event_candidates_follow_the_reference_lines():
    _, ei, _ = load_graph('c1')
    jj = ei[ET[2]]
    g = torch.Generator().manual_seed(1)
    y = {'edge_event': torch.randn(jj.shape[1], generator=g), 'grain_area': torch.rand(118, generator=g) * 3e-4}
    mask = torch.ones(118, 1); mask[::7] = 0
    L1, ge = orc.event_candidates(y, jj, mask)
    assert all(jj[0, e] < jj[1, e] and torch.sigmoid(y['edge_event'])[e] > 0.6 for e in L1) and torch.all(L1[1:] > L1[:-1])
    assert all(mask[gidx, 0] > 0 and y['grain_area'][gidx] < 1e-4 for gidx in ge)
    assert torch.all(y['grain_area'][ge][1:] >= y['grain_area'][ge][:-1])
    n_ref = sum(1 for gidx in range(118) if mask[gidx, 0] > 0 and y['grain_area'][gidx] < 1e-4)
    assert len(ge) == n_ref and n_ref > 5


def test_select_events_validates_arguments_without_gpu():
    from graingraphnn_b200 import _lib
    L = _lib.lib()
    cnt = (np.zeros(1, np.int32)).ctypes.data
    assert L.gg_select_events(None, -1, 1, 0.0, 0, None, None, None, 0, 0, cnt, None, None, None) == -1
    assert L.gg_select_events(None, 0, 1, 0.0, 2, None, None, None, 0, 0, cnt, None, None, None) == -1     # unknown mode
    assert L.gg_select_events(None, 0, 1, 0.0, 0, cnt, None, None, 0, 0, cnt, None, None, None) == -1      # src without dst
    assert L.gg_select_events(None, 0, 1, 0.0, 0, None, None, None, 0, 8, cnt, None, None, None) == -1     # cap without buffers
    assert L.gg_select_events(None, 0, 1, 0.0, 0, None, None, None, 0, 0, None, None, None, None) == -1    # no counter


# ------------------------------------------------------------------------------------------------------------- GPU
def _dev():
    return torch.device('cuda:0')


@pytest.mark.gpu
@pytest.mark.parametrize('cap', [4096, 4])
def test_selected_events_equal_the_reference_scans(cap):
    from graingraphnn_b200.events import EventSelector
    _, ei, _ = load_graph('c2')
    jj = ei[ET[2]]
    g = torch.Generator().manual_seed(2)
    y = {'edge_event': torch.randn(jj.shape[1], generator=g) * 2, 'grain_area': torch.rand(1043, generator=g) * 2e-3}
    mask = torch.ones(1043, 1); mask[::5] = 0
    sel = EventSelector(_dev(), cap_edges=cap, cap_grains=cap)
    lo, hi = sel.logit_min, sel.logit_band_hi
    y['edge_event'][:4] = torch.tensor([lo, float(np.nextafter(np.float32(lo), np.float32(-np.inf))),
                                        float(np.nextafter(np.float32(max(lo, hi)), np.float32(np.inf))), 50.0])
    L1_ref, ge_ref = orc.event_candidates(y, jj, mask)
    sel.select_edge_events(y['edge_event'].to(_dev()), jj.to(_dev()))
    sel.select_grain_events(y['grain_area'].to(_dev()), mask.to(_dev()))
    out = sel.fetch()
    band = (y['edge_event'] >= lo) & (y['edge_event'] <= hi)
    outside = lambda ids: ids[~band[ids]]                                   # noqa: E731
    assert torch.equal(outside(out['L1']), outside(L1_ref)) and len(L1_ref) > 300
    assert torch.equal(out['L1_logit'], y['edge_event'][out['L1']])
    assert torch.equal(out['grain_event'], ge_ref) and len(ge_ref) > 20
    assert torch.equal(out['grain_event_area'], y['grain_area'][out['grain_event_ids']])
    assert sel.d2h_bytes <= 16 + 8 * (int((y['edge_event'] >= lo).sum()) + len(ge_ref))      # (id, value) pairs only


@pytest.mark.gpu
def test_engine_steps_leave_the_event_candidates_on_the_device():
    """RolloutEngine.enable_event_selection: after each step fetch_events() equals the reference's scans of the full
    predictions, eager and replayed from a CUDA graph; strided views (area of the owned rows) included."""
    from graingraphnn_b200.engine import RolloutEngine
    x, ei, ea = load_graph('c2')
    x['grain'][:, 3] *= 0.004                        # small areas, so that some grains fall below 1e-4
    dev = _dev()
    eng = RolloutEngine.from_state_dicts(orc.synth_state_dict('regressor', 1), orc.synth_state_dict('classifier', 2), dev)
    eng.set_graph({k: v.to(dev) for k, v in x.items()}, {k: v.to(dev) for k, v in ei.items()}, {k: v.to(dev) for k, v in ea.items()})
    mask = torch.ones(x['grain'].shape[0], 1); mask[::9] = 0
    eng.enable_event_selection(mask)
    n_events = 0
    for step in range(3):
        if step == 1:
            eng.capture(span=6, warmup=1)
        pred = eng.step(6)
        y = {k: pred[k].cpu() for k in ('edge_event', 'grain_area')}
        L1_ref, ge_ref = orc.event_candidates(y, ei[ET[2]], mask)
        out = eng.fetch_events()
        band = (y['edge_event'] >= eng._events.logit_min) & (y['edge_event'] <= eng._events.logit_band_hi)
        assert torch.equal(out['L1'][~band[out['L1']]], L1_ref[~band[L1_ref]])
        assert torch.equal(out['grain_event'], ge_ref)
        n_events += len(L1_ref) + len(ge_ref)
    assert n_events > 0


@pytest.mark.gpu
def test_select_events_full_size_counts():
    """6 M logits (the 10^6-grain configuration's jj edges): candidate set equals a torch formulation on the device."""
    from graingraphnn_b200.events import EventSelector
    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(3)
    n = 5_992_920
    v = torch.randn(n, generator=g, device=dev) * 1.5 - 3.0
    src = torch.randint(0, 2_000_000, (n,), generator=g, device=dev)
    dst = torch.randint(0, 2_000_000, (n,), generator=g, device=dev)
    sel = EventSelector(dev, cap_edges=1 << 16)
    sel.select_edge_events(v, torch.stack([src, dst]))
    out = sel.fetch()
    want = ((v >= sel.logit_min) & (src < dst)).nonzero().view(-1).cpu()
    keep = torch.sigmoid(v[want].cpu()) > 0.6
    assert torch.equal(out['L1'], want[keep]) and len(want) > 10_000
