"""CPU: host-side logic of the product — module API, state_dict layout, lazy materialisation, deepcopy, weight packing
(checked by running a torch emulation of the kernels on the packed buffers against the oracle), and the no-CPU-fallback rule."""
import copy
import os

import pytest
import torch

import grain_oracle as orc
from emulate import emu_cell
from util import ET, GOLDEN, load_graph, rel_err

from graingraphnn_b200 import _lib
from graingraphnn_b200.cell import pad_features
from graingraphnn_b200.engine import _Hyper
from graingraphnn_b200.heteropgclstm import HeteroPGC, HeteroPGCLSTM
from graingraphnn_b200.models import GrainNN_classifier, GrainNN_regressor
from graingraphnn_b200.periodconv import PeriodConv as SumConv
from graingraphnn_b200.periodGATconv import PeriodConv


def hyper():
    return _Hyper({'grain': list(range(11)), 'joint': list(range(8))}, {'grain': [0, 1], 'joint': [0, 1]}, 96,
                  (['grain', 'joint', 'mask'], list(ET)), 'cpu')


def test_state_dict_layout_equals_reference():
    """Keys, order and shapes against the state_dict of the reference's OWN models (tests/golden/reference_state_dict_layout.json,
    dumped by oracle/make_golden_layout.py from the unmodified models.py), the oracle's restatement and the bench helper."""
    import json
    from graingraphnn_b200.weights import state_dict_shapes
    with open(os.path.join(GOLDEN, 'reference_state_dict_layout.json')) as f:
        ref = json.load(f)
    R = GrainNN_regressor(hyper())
    C = GrainNN_classifier(hyper(), R)
    for m, kind, n in ((R, 'regressor', 1204612), (C, 'classifier', 1204806)):
        shapes = orc.param_shapes(kind)
        assert [(k, tuple(s)) for k, s in ref[kind]['keys']] == list(shapes.items()) == list(state_dict_shapes(kind).items())
        assert ref[kind]['n_params'] == n
        sd = m.state_dict()
        assert list(sd.keys()) == list(shapes.keys())
        assert all(tuple(sd[k].shape) == shapes[k] for k in shapes)
        assert sum(p.numel() for p in m.parameters()) == n
        res = m.load_state_dict(orc.synth_state_dict(kind, 7))
        assert not res.missing_keys and not res.unexpected_keys


def test_classifier_deepcopies_regressor_cells():
    R = GrainNN_regressor(hyper())
    R.load_state_dict(orc.synth_state_dict('regressor', 1))
    C = GrainNN_classifier(hyper(), R)
    a = R.gclstm_encoder.cell_list[0].conv_i.conv(ET[0]).lin_key.weight
    b = C.gclstm_encoder.cell_list[0].conv_i.conv(ET[0]).lin_key.weight
    assert torch.equal(a, b) and a.data_ptr() != b.data_ptr()


def test_lazy_periodconv_materialises_on_load_and_survives_deepcopy():
    conv = PeriodConv(in_channels=(-1, -1), out_channels=96)
    assert not conv.lin_key.materialized
    conv2 = copy.deepcopy(conv)
    sd = {k[len('gclstm_decoder.cell_list.0.conv_i.convs.grain__push__joint.'):]: v
          for k, v in orc.synth_state_dict('regressor', 3).items()
          if k.startswith('gclstm_decoder.cell_list.0.conv_i.convs.grain__push__joint.')}
    conv2.load_state_dict(sd)
    assert conv2.lin_key.weight.shape == (96, 107) and conv2.lin_query.weight.shape == (96, 104)
    assert conv2.lin_edge.weight.shape == (96, 1) and conv2.lin_beta is None


def test_constructor_rejects_unsupported_variants():
    with pytest.raises(NotImplementedError):
        PeriodConv(8, 96, heads=2)
    with pytest.raises(NotImplementedError):
        PeriodConv(8, 96, beta=True)
    assert SumConv(8, 96).weighted is False and PeriodConv(8, 96).weighted is True


def test_no_cpu_fallback():
    x, ei, ea = load_graph('c1')
    R = GrainNN_regressor(hyper())
    with pytest.raises(RuntimeError, match='no CPU'):
        R(x, ei, ea)
    conv = PeriodConv(8, 96)
    with pytest.raises(RuntimeError, match='no CPU'):
        conv(x['joint'], ei[ET[2]], ea[ET[2]])


def test_no_cpu_fallback_in_the_widened_rows():
    """Geometry feedback and event selection refuse CPU tensors; the host topology update is host code by design and never
    imports the oracle."""
    import inspect
    from graingraphnn_b200 import events, geometry, topology
    x, ei, _ = load_graph('c1')
    with pytest.raises(RuntimeError, match='no CPU'):
        geometry.RegionIndex(ei[ET[0]], 118, 236)

    class _Idx:                      # region_center checks its tensors before it touches the index
        n_grain, n_joint = 118, 236
    with pytest.raises(RuntimeError, match='no CPU'):
        geometry.region_center(x['joint'], _Idx())
    sel = object.__new__(events.EventSelector)
    sel._buf = {'edge': (None, None, None, 0, None)}
    sel.device = torch.device('cpu')
    with pytest.raises(RuntimeError, match='no CPU'):
        sel._select('edge', torch.zeros(4), 0.0, 0)
    for mod in (events, geometry, topology):
        src = inspect.getsource(mod)
        assert 'grain_oracle' not in src and 'topology_oracle' not in src and 'import oracle' not in src


def _csr(ei, x):
    csr = {}
    for e in ET:
        rp, col, perm = orc.csr_by_dst(ei[e], x[e[2]].shape[0])
        csr[e] = (torch.from_numpy(rp), torch.from_numpy(col), torch.from_numpy(perm).long())
    return csr


@pytest.mark.parametrize('cls,gates,mode,raw', [(HeteroPGCLSTM, ('i', 'f', 'c', 'o'), _lib.GG_GATE_LSTM, False),
                                                (HeteroPGCLSTM, ('i', 'f', 'c', 'o'), _lib.GG_GATE_LSTM, True),
                                                (HeteroPGC, ('i',), _lib.GG_GATE_RELU, False)])
def test_packed_cell_algebra_matches_oracle(cls, gates, mode, raw):
    """fp64 emulation of the three kernels on the packed (fp32-rounded) weights == oracle cell to ~1e-7; raw: the raw-score
    layout with hidden state (source row [X padded to 32 | h | V], Q' of 32 + C floats per gate, We . q in slot 31)."""
    x, ei, ea = load_graph('c1', torch.float64)
    sd = orc.synth_state_dict('regressor', 1)
    cell = cls({'grain': 11, 'joint': 8}, 96, (['grain', 'joint'], list(ET)))
    pre = 'gclstm_decoder.cell_list.0.'
    cell.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)
                          and any(f'.conv_{g}.' in '.' + k[len(pre):] or k[len(pre):].startswith(f'b_{g}.') or
                                  k[len(pre):].startswith(f'conv_{g}.') for g in gates)})
    sd64 = {k: v.double() for k, v in sd.items()}
    h0, c0 = orc.pgclstm_cell(sd64, 'gclstm_encoder.cell_list.0', x, ei, ea)
    csr = _csr(ei, x)
    eac = {e: ea[e].reshape(-1)[csr[e][2]] for e in ET}
    pk = cell.packed(gates, True, 'cpu', raw=raw)
    assert pk.raw_k == (128 if raw else 0) and (not raw or (pk.we_slot == 31 and pk.ncols['joint'] == 2 * (128 + 384) + 2 * 4 * 128 and pk.ncols['grain'] == 1024))
    xpad = {t: pad_features(x[t].float(), pk.k1p[t]).double() for t in x}
    hh, cc = emu_cell(pk, xpad, h0, c0, {e: csr[e][:2] for e in ET}, eac, mode)
    if cls is HeteroPGCLSTM:
        href, cref = orc.pgclstm_cell(sd64, 'gclstm_decoder.cell_list.0', x, ei, ea, h0, c0)
        assert all(rel_err(cc[t], cref[t]) < 1e-6 for t in x)
    else:
        href, _ = orc.pgc_cell(sd64, 'gclstm_decoder.cell_list.0', x, ei, ea, h0, c0)
    assert all(rel_err(hh[t], href[t]) < 1e-6 for t in x)


@pytest.mark.parametrize('raw', [False, True])
def test_packed_encoder_drops_forget_gate_and_hidden_columns(raw):
    """Encoder packing (h = c = 0): classic layout and the raw-score layout (no key projection, Q' of 16 floats per gate)."""
    x, ei, ea = load_graph('c1', torch.float64)
    sd = orc.synth_state_dict('classifier', 2)
    C = GrainNN_classifier(hyper())
    C.load_state_dict(sd)
    cell = C.gclstm_encoder.cell_list[0]
    pk = cell.packed(('i', 'c', 'o'), False, 'cpu', raw=raw)
    assert pk.G == 3 and pk.Wcat['grain'].shape[1] == 12 and pk.Wcat['joint'].shape[1] == 8
    if raw:
        assert pk.raw_k == 16 and pk.ncols['joint'] == 2 * (16 + 3 * 96) + 2 * 3 * 16 and pk.ncols['grain'] == 16 + 3 * 96 + 3 * 16
        assert all(pk.voff[e] == pk.koff[e] + 16 for e in ET)
    else:
        assert pk.raw_k == 0 and pk.ncols['joint'] == 6 * 3 * 96 + 2 * (3 * 4 + 4) and pk.ncols['grain'] == 3 * 3 * 96 + 3 * 4 + 4
    assert all(pk.posoff[e] == (pk.qoff[e] + 12 if raw else pk.qxoff[e] + 3 * 4) for e in ET)     # position: slots 12..14 of the first Q', or behind Q | QX
    csr = _csr(ei, x)
    eac = {e: ea[e].reshape(-1)[csr[e][2]] for e in ET}
    xpad = {t: pad_features(x[t].float(), pk.k1p[t]).double() for t in x}
    hh, cc = emu_cell(pk, xpad, None, None, {e: csr[e][:2] for e in ET}, eac, _lib.GG_GATE_LSTM0)
    href, cref = orc.pgclstm_cell({k: v.double() for k, v in sd.items()}, 'gclstm_encoder.cell_list.0', x, ei, ea)
    assert all(rel_err(hh[t], href[t]) < 1e-6 and rel_err(cc[t], cref[t]) < 1e-6 for t in x)


@pytest.mark.parametrize('weighted', [True, False])
def test_packed_single_conv_matches_oracle(weighted):
    x, ei, ea = load_graph('c1', torch.float64)
    sd = orc.synth_state_dict('regressor', 3)
    pre = 'gclstm_decoder.cell_list.0.conv_i.convs.grain__push__joint.'
    conv = (PeriodConv if weighted else SumConv)(in_channels=(-1, -1), out_channels=96)
    conv.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)})
    xs = torch.cat([x['grain'], torch.randn(118, 96, dtype=torch.float64)], 1)
    xd = torch.cat([x['joint'], torch.randn(236, 96, dtype=torch.float64)], 1)
    ref = orc.period_conv({k: v.double() for k, v in sd.items()}, pre[:-1], xs, xd, ei[ET[0]], ea[ET[0]], weighted)
    pk = conv._pack(107, 104, False, 'cpu')
    rp, col, perm = orc.csr_by_dst(ei[ET[0]], 236)
    et = pk.edge_types[0]
    xpad = {'s': pad_features(xs.float(), pk.k1p['s']).double(), 'd': pad_features(xd.float(), pk.k1p['d']).double()}
    out, _ = emu_cell(pk, xpad, None, None, {et: (torch.from_numpy(rp), torch.from_numpy(col))},
                      {et: ea[ET[0]].reshape(-1)[torch.from_numpy(perm).long()]}, _lib.GG_GATE_RAW)
    assert rel_err(out['d'], ref) < 1e-6


def test_sage_cells_have_the_reference_state_dict_layout():
    """Keys / shapes / order of HeteroGCLSTM equal what the reference module produced (tests/golden/gclstm_state_dict.pt,
    written by oracle/make_golden.py from /root/reference/heterogclstm.py); HeteroGC holds only conv_i."""
    import os
    from util import GOLDEN
    from graingraphnn_b200.heterogclstm import HeteroGC, HeteroGCLSTM
    gsd = torch.load(os.path.join(GOLDEN, 'gclstm_state_dict.pt'))
    cell = HeteroGCLSTM({'grain': 11, 'joint': 8}, 96, (['grain', 'joint'], list(ET)))
    sd = cell.state_dict()
    assert list(sd.keys()) == list(gsd.keys())
    assert all(sd[k].shape == gsd[k].shape for k in gsd)
    assert not cell.load_state_dict(gsd).missing_keys
    cell2 = copy.deepcopy(cell)
    assert torch.equal(cell2.conv_o.conv(ET[1]).lin_l.weight, gsd['conv_o.convs.joint__pull__grain.lin_l.weight'])
    gc = HeteroGC({'grain': 11, 'joint': 8}, 96, (['grain', 'joint'], list(ET)))
    assert sorted(gc.state_dict()) == sorted(k for k in gsd if k.startswith('conv_i.')) or \
        all(k.startswith('conv_i.') for k in gc.state_dict())
    assert gc.conv_i.conv(ET[0]).lin_l.weight.shape == (96, 11) and gc.conv_i.conv(ET[0]).lin_r.weight.shape == (96, 8)
    with pytest.raises(RuntimeError):
        cell({'grain': torch.zeros(3, 11), 'joint': torch.zeros(4, 8)}, {})      # CPU tensors: no fallback


def test_collate_offsets_edges_like_the_reference_loader():
    """Block-diagonal batch (BASELINE config 5): node features stacked, edge indices offset per graph (data_loader.py:113-162)."""
    from graingraphnn_b200.ensemble import collate
    x1, ei1, ea1 = load_graph('c1')
    x2, ei2, ea2 = load_graph('c2')
    x, ei, ea, ptr = collate([(x1, ei1, ea1), (x2, ei2, ea2), (x1, ei1, ea1)])
    ng1, nj1, ng2, nj2 = x1['grain'].shape[0], x1['joint'].shape[0], x2['grain'].shape[0], x2['joint'].shape[0]
    assert ptr['grain'] == [0, ng1, ng1 + ng2, 2 * ng1 + ng2] and ptr['joint'] == [0, nj1, nj1 + nj2, 2 * nj1 + nj2]
    for e in ET:
        E1, E2 = ei1[e].shape[1], ei2[e].shape[1]
        assert ptr[e] == [0, E1, E1 + E2, 2 * E1 + E2]
        off = torch.tensor([[ptr[e[0]][1]], [ptr[e[2]][1]]])
        assert torch.equal(ei[e][:, E1:E1 + E2], ei2[e] + off)
        off = torch.tensor([[ptr[e[0]][2]], [ptr[e[2]][2]]])
        assert torch.equal(ei[e][:, E1 + E2:], ei1[e] + off)
        assert torch.equal(ea[e][E1:E1 + E2].reshape(-1), ea2[e].reshape(-1))
    assert torch.equal(x['grain'][ng1:ng1 + ng2], x2['grain'])


@pytest.mark.skipif(not os.path.exists('/root/reference/models.py'), reason='/root/reference is not mounted')
def test_reference_models_construct_over_these_cells():
    """the reference's unmodified models.py / parameters.py with this package's modules swapped in under the
    reference's module names (test.py:16, :162-184) — the models construct, every state_dict key of the seeded stand-ins loads
    strictly, the parameter counts are the log files' (regressor0_logfile:40, classifier1_logfile:40).  A subprocess, so that the
    reference's top-level `models` does not shadow anything in this session.  (The forward needs a GPU, where the reference is
    not mounted; tests/test_gpu_parity.py covers it through this package's own models.py, which is the same few lines.)"""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import sys, types, torch
sys.path.insert(0, %r); sys.path.insert(0, %r)
import ref_shims
ref_shims.install_plot_stubs()
import graingraphnn_b200
from graingraphnn_b200 import heteropgclstm, heterogclstm, periodGATconv, periodconv
from graingraphnn_b200.weights import synth_state_dict
for m in (heteropgclstm, heterogclstm, periodGATconv, periodconv):
    sys.modules[m.__name__.rsplit('.', 1)[1]] = m
sys.path.insert(0, '/root/reference')
import models, parameters
assert models.__file__.startswith('/root/reference')
ET = [('grain', 'push', 'joint'), ('joint', 'pull', 'grain'), ('joint', 'connect', 'joint')]
hp, hpc = parameters.regressor(0), parameters.classifier_transfered(1)
hp.metadata = (['grain', 'joint', 'mask'], ET); hp.device = 'cpu'
hp.features = {'grain': list(range(11)), 'joint': list(range(8))}          # the widths of the reference's graphs (graphs/40_40)
hp.targets = {'grain': ['darea', 'extraV'], 'joint': ['dx', 'dy']}
hpc.metadata, hpc.features, hpc.device = hp.metadata, hp.features, 'cpu'
R = models.GrainNN_regressor(hp)
r = R.load_state_dict(synth_state_dict('regressor', 1)); assert not r.missing_keys and not r.unexpected_keys
C = models.GrainNN_classifier(hpc, R)
r = C.load_state_dict(synth_state_dict('classifier', 2)); assert not r.missing_keys and not r.unexpected_keys
assert type(R.gclstm_encoder.cell_list[0]) is heteropgclstm.HeteroPGCLSTM
assert sum(p.numel() for p in R.parameters()) == 1204612 and sum(p.numel() for p in C.parameters()) == 1204806
print('OK')
''' % (root, os.path.join(root, 'oracle'))
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith('OK'), out.stderr[-2000:]
