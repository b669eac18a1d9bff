"""GPU (-m gpu): the CUDA path, called through the boundary modules / C ABI, against the CPU oracle and the committed
golden vectors.  Bars: indices bit-exact; floating point within 1e-4 relative (max |a-b| / max |ref|) per step."""
import numpy as np
import pytest
import torch

import grain_oracle as orc
from util import ET, SHORT, load_golden, load_graph, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


def dev():
    assert torch.cuda.is_available()
    return torch.device('cuda:0')


def to_dev(d):
    return {k: v.to(dev()) for k, v in d.items()}


def hyper():
    from graingraphnn_b200.engine import _Hyper
    return _Hyper({'grain': list(range(11)), 'joint': list(range(8))}, {'grain': [0, 1], 'joint': [0, 1]}, 96,
                  (['grain', 'joint', 'mask'], list(ET)), 'cuda')


def models(seed_r=1, seed_c=2):
    from graingraphnn_b200.models import GrainNN_classifier, GrainNN_regressor
    R = GrainNN_regressor(hyper())
    R.load_state_dict(orc.synth_state_dict('regressor', seed_r))
    C = GrainNN_classifier(hyper(), R)
    C.load_state_dict(orc.synth_state_dict('classifier', seed_c))
    return R.to(dev()).eval(), C.to(dev()).eval()


def random_graph(n_src, n_dst, E, seed, hub=None):
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(0, n_src, (E,), generator=g)
    dst = torch.randint(0, n_dst, (E,), generator=g)
    if hub is not None:
        dst[: E // 3] = hub          # one long row (exercises the > 32 in-edge path)
    return torch.stack([src, dst])


# ------------------------------------------------------------------------------------------------ (a) CSR, bit-exact
@pytest.mark.parametrize('case', ['c1', 'c2', 'random', 'hub', 'empty', 'big'])
def test_csr_build_bit_exact(case):
    from graingraphnn_b200.graph import build_csr
    if case in ('c1', 'c2'):
        x, ei, _ = load_graph(case)
        todo = [(ei[e], x[e[0]].shape[0], x[e[2]].shape[0]) for e in ET]
    elif case == 'random':
        todo = [(random_graph(50, 70, 400, 0), 50, 70), (random_graph(7, 5000, 3000, 1), 7, 5000)]
    elif case == 'hub':
        todo = [(random_graph(100, 40, 900, 2, hub=11), 100, 40)]
    elif case == 'empty':
        todo = [(torch.zeros(2, 0, dtype=torch.int64), 10, 10)]
    else:
        todo = [(random_graph(200000, 300000, 1500000, 3), 200000, 300000)]
    for ei, ns, nd in todo:
        g = build_csr(ei.to(dev()), ns, nd)
        rp, col, perm = orc.csr_by_dst(ei, nd)
        assert np.array_equal(g.rowptr.cpu().numpy(), rp)
        assert np.array_equal(g.perm.cpu().numpy(), perm)
        assert np.array_equal(g.col.cpu().numpy(), col)


def test_csr_rejects_out_of_range_endpoints():
    from graingraphnn_b200.graph import build_csr
    ei = torch.tensor([[0, 1, 2], [0, 9, 1]])
    with pytest.raises(IndexError):
        build_csr(ei.to(dev()), 3, 3)


# ------------------------------------------------------------------------------------------------ (a11) edge length
@pytest.mark.parametrize('name', ['c1', 'c2'])
def test_edge_length_bit_exact(name):
    from graingraphnn_b200.graph import build_csr, edge_length
    x, ei, ea = load_graph(name)
    ref = orc.edge_attr_rebuild(x, ei)
    xd = to_dev(x)
    for e in ET:
        g = build_csr(ei[e].to(dev()), x[e[0]].shape[0], x[e[2]].shape[0])
        out, out_csr = edge_length(xd[e[0]], xd[e[2]], ei[e].to(dev()), g)
        # IEEE sub / mul / add / sqrt on the GPU; torch's CPU sqrt is off by one ulp on ~0.5 % of the edges
        # (checked against numpy's correctly-rounded sqrt), so the bar here is <= 1 ulp, and exact vs numpy.
        d = x[e[0]][ei[e][0], :2].numpy() - x[e[2]][ei[e][1], :2].numpy()
        d = d + ((d < -0.5).astype(np.float32) - (d > 0.5).astype(np.float32))
        exact = np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(np.float32))
        assert np.array_equal(out.cpu().numpy().reshape(-1), exact)
        assert float((out.cpu() - ref[e]).abs().max()) <= float(np.spacing(np.float32(ref[e].max())))
        assert torch.equal(out_csr.cpu(), out.cpu().reshape(-1)[g.perm.cpu().long()])


def test_edge_refresh_equals_edge_length_and_edge_wrap_bit_for_bit():
    """gg_edge_refresh (one pass over the CSR rows) == gg_edge_length + gg_permute + gg_edge_wrap, bit for bit."""
    from graingraphnn_b200 import _lib
    from graingraphnn_b200._lib import check, ptr
    from graingraphnn_b200.cell import pad_features
    from graingraphnn_b200.graph import build_csr, edge_length, edge_wrap
    x, ei, _ = load_graph('c2')
    L = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    xp = {t: pad_features(v.to(dev()), (v.shape[1] + 3) // 4 * 4) for t, v in x.items()}
    for e in ET:
        eid = ei[e].to(dev())
        g = build_csr(eid, x[e[0]].shape[0], x[e[2]].shape[0])
        ea, ea_csr = edge_length(xp[e[0]], xp[e[2]], eid, g)
        wr = edge_wrap(g, xp[e[0]], xp[e[2]])
        E = eid.shape[1]
        ea2, ea_csr2 = torch.full((E, 1), -1.0, device=dev()), torch.full((E,), -1.0, device=dev())
        wr2 = torch.full((E,), -1, dtype=torch.int32, device=dev())
        check(L.gg_edge_refresh(ptr(xp[e[0]]), xp[e[0]].stride(0), ptr(xp[e[2]]), xp[e[2]].stride(0), ptr(g.rowptr), ptr(g.col),
                                ptr(g.perm), g.n_dst, ptr(wr2), ptr(ea_csr2), ptr(ea2), st), 'gg_edge_refresh')
        assert torch.equal(ea2, ea) and torch.equal(ea_csr2, ea_csr) and torch.equal(wr2, wr[:E])


# ------------------------------------------------------------------------------------------------ (b)+(c) single conv
@pytest.mark.parametrize('tag', ['gat', 'sum'])
@pytest.mark.parametrize('et', [ET[0], ET[2]])
def test_period_conv_matches_golden_and_oracle(tag, et):
    from graingraphnn_b200.periodconv import PeriodConv as SumConv
    from graingraphnn_b200.periodGATconv import PeriodConv
    x, ei, ea = load_graph('c1')
    g = load_golden('c1')
    h, _ = orc.encode_decode(orc.synth_state_dict('regressor', 1), x, ei, ea)
    xin = {t: torch.cat([x[t], h[t]], 1) for t in x}
    sd = orc.synth_state_dict('regressor', 3)
    pre = f'gclstm_decoder.cell_list.0.conv_i.convs.{"__".join(et)}.'
    conv = (PeriodConv if tag == 'gat' else SumConv)(in_channels=(-1, -1), out_channels=96)
    conv.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)})
    conv = conv.to(dev())
    xx = xin[et[0]].to(dev()) if et[0] == et[2] else (xin[et[0]].to(dev()), xin[et[2]].to(dev()))
    out = conv(xx, ei[et].to(dev()), ea[et].to(dev()))
    assert rel_err(out, g[f'conv_{tag}_{SHORT[et]}']) < TOL


@pytest.mark.parametrize('C', [32, 64, 128])
def test_period_conv_other_widths_and_ragged_graph(C):
    """Widths of the reference's hyper-parameter grid (parameters.py:18-21), a ragged random graph with empty rows and one
    row of > 32 in-edges (recompute path)."""
    from graingraphnn_b200.periodGATconv import PeriodConv
    torch.manual_seed(C)
    ns, nd, E = 300, 200, 1500
    ei = random_graph(ns, nd, E, C, hub=5)
    ei = ei[:, ei[1] != 17]                                # node 17 has no in-edges
    xs, xd = torch.rand(ns, 10), torch.rand(nd, 7)
    ea = torch.rand(ei.shape[1], 1) * 0.1
    conv = PeriodConv(in_channels=(10, 7), out_channels=C)
    sd = {k: v.clone() for k, v in conv.state_dict().items()}
    ref = orc.period_conv({'c.' + k: v for k, v in sd.items()}, 'c', xs, xd, ei, ea)
    out = conv.to(dev())((xs.to(dev()), xd.to(dev())), ei.to(dev()), ea.to(dev()))
    assert rel_err(out, ref) < TOL
    skip = xd @ sd['lin_skip.weight'].t() + sd['lin_skip.bias']
    assert rel_err(out[17].cpu(), skip[17]) < 1e-5 and torch.isfinite(out).all()


def test_period_conv_without_edges_returns_the_skip_term():
    """An edge type that lost all its edges (E = 0): PyG's scatter-add leaves zeros, the output is lin_skip alone."""
    from graingraphnn_b200.periodGATconv import PeriodConv
    torch.manual_seed(3)
    xs, xd = torch.rand(40, 10), torch.rand(25, 7)
    conv = PeriodConv(in_channels=(10, 7), out_channels=96)
    sd = {k: v.clone() for k, v in conv.state_dict().items()}
    ei = torch.zeros(2, 0, dtype=torch.int64)
    out = conv.to(dev())((xs.to(dev()), xd.to(dev())), ei.to(dev()), torch.zeros(0, 1, device=dev()))
    skip = xd @ sd['lin_skip.weight'].t() + sd['lin_skip.bias']
    assert out.shape == (25, 96) and rel_err(out, skip) < 1e-5


# ------------------------------------------------------------------------------------------------ (c) tensor-core GEMM
@pytest.mark.parametrize('M,N,K2', [(300, 2336, 96), (1000, 1168, 96), (129, 1752, 0), (5000, 256, 64), (128, 32, 32)])
def test_tcgen05_node_proj_matches_fp32(M, N, K2):
    """3xTF32 tcgen05 projection == fp32 SIMT projection == float64 reference to ~1e-6 (plain TF32 would be ~5e-4)."""
    from graingraphnn_b200 import _lib
    from graingraphnn_b200._lib import check, ptr
    from graingraphnn_b200.packing import split_tf32
    L = _lib.lib()
    assert L.gg_tc_supported() == 1
    g = torch.Generator().manual_seed(M + N)
    K1 = 12
    x = torch.rand(M, K1, generator=g) * 2 - 1
    x[:, 11] = 0
    h = torch.rand(M, K2, generator=g) - 0.5 if K2 else None
    W = (torch.rand(N, 32 + K2, generator=g) - 0.5) / 5
    W[:, K1:32] = 0
    b = torch.rand(N, generator=g)
    a = torch.cat([x, torch.zeros(M, 32 - K1)] + ([h] if K2 else []), 1)
    ref = a.double() @ W.double().t() + b.double()
    d = dev()
    xd, hd, bd = x.to(d), (h.to(d) if K2 else None), b.to(d)
    whi, wlo = split_tf32(W.to(d))
    kp = 32 + K2
    ahi, alo = torch.empty(M, kp, device=d), torch.empty(M, kp, device=d)
    out = torch.full((M, N), float('nan'), device=d)
    st = torch.cuda.current_stream().cuda_stream
    check(L.gg_split_tf32(ptr(xd), K1, K1, ptr(hd), K2, K2, M, ptr(ahi), ptr(alo), kp, 32, st), 'gg_split_tf32')
    assert rel_err(ahi + alo, a) < 1e-6
    for k_first in (K1, 32):             # with and without skipping the MMAs of the zero padding behind the features
        out.fill_(float('nan'))
        check(L.gg_node_proj_tc(ptr(ahi), ptr(alo), kp, k_first, ptr(whi), ptr(wlo), N, ptr(bd), ptr(out), N, M, 0, st), 'gg_node_proj_tc')
        torch.cuda.synchronize()
        assert torch.isfinite(out).all()
        assert rel_err(out, ref) < 2e-6, rel_err(out, ref)
    # the projection with the split fused in (no A_hi / A_lo arrays): same operands, same bar
    out.fill_(float('nan'))
    check(L.gg_node_proj_fused(ptr(xd), K1, K1, ptr(hd), K2, K2, ptr(whi), ptr(wlo), N, ptr(bd), ptr(out), N, M, 0, st), 'gg_node_proj_fused')
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    assert rel_err(out, ref) < 2e-6, rel_err(out, ref)
    # the fp32 CUDA-core kernel on the same (unsplit) operands
    Wd = torch.cat([W[:, :K1], W[:, 32:]], 1).contiguous().to(d)
    out2 = torch.empty(M, N, device=d)
    check(L.gg_node_proj(ptr(xd), K1, K1, ptr(hd), K2, K2, ptr(Wd), K1 + K2, ptr(bd), ptr(out2), N, M, N, st), 'gg_node_proj')
    assert rel_err(out2, ref) < 2e-6
    assert rel_err(out, out2) < 2e-6


# ------------------------------------------------------------------------------------------------ cells
def test_pgclstm_cell_encoder_and_decoder_states():
    x, ei, ea = load_graph('c1')
    g = load_golden('c1')
    R, _ = models()
    xd, eid, ead = to_dev(x), to_dev(ei), to_dev(ea)
    enc = R.gclstm_encoder(xd, eid, ead, None)
    dec = R.gclstm_decoder(xd, eid, ead, enc)
    assert rel_err(enc[0][0]['joint'], g['r_enc_h_joint']) < TOL and rel_err(enc[0][1]['grain'], g['r_enc_c_grain']) < TOL
    for k, v in (('r_dec_h_joint', dec[0][0]['joint']), ('r_dec_h_grain', dec[0][0]['grain']),
                 ('r_dec_c_joint', dec[0][1]['joint']), ('r_dec_c_grain', dec[0][1]['grain'])):
        assert rel_err(v, g[k]) < TOL, k


def ragged_hetero_graph(seed, ng=220, nj=400):
    """Edge lists that exercise the tile logic of the warp-specialised gather: a row of 300 in-edges (spans several tiles),
    rows without in-edges, a long run of rows with a single in-edge (more starting targets than header slots)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for et in ET:
        ns, nd = (ng if et[0] == 'grain' else nj), (ng if et[2] == 'grain' else nj)
        src = torch.randint(0, ns, (3 * nd,), generator=g)
        dst = torch.randint(0, nd, (3 * nd,), generator=g)
        hub = torch.full((300,), 7, dtype=torch.int64)
        ones = torch.arange(40, 140)                                   # a run of in-degree ~1 rows
        keep = (dst < 40) | (dst >= 140)
        src, dst = src[keep], dst[keep]
        keep = (dst != 150) & (dst != 151) & (dst != nd - 1)           # rows without in-edges (the last row among them)
        src, dst = src[keep], dst[keep]
        src = torch.cat([src, torch.randint(0, ns, (300,), generator=g), torch.randint(0, ns, (100,), generator=g)])
        dst = torch.cat([dst, hub, ones])
        perm = torch.randperm(src.numel(), generator=g)
        out[et] = torch.stack([src[perm], dst[perm]])
    return out


@pytest.mark.parametrize('gather', ['tiled', 'items'])
def test_pgclstm_cell_on_ragged_graph(gather, monkeypatch):
    """Encoder form (no state: raw-score gather, 3 live gates) and decoder form (4 gates) of the cell on a ragged graph,
    through the warp-specialised gather and through the per-warp item-list gather."""
    from graingraphnn_b200.heteropgclstm import HeteroPGCLSTM
    monkeypatch.setenv('GG_GATHER', gather)
    torch.manual_seed(5)
    ng, nj = 220, 400
    x = {'grain': torch.rand(ng, 11), 'joint': torch.rand(nj, 8)}
    ei = ragged_hetero_graph(11, ng, nj)
    ea = {e: torch.rand(ei[e].shape[1], 1) * 0.2 for e in ET}
    sd = orc.synth_state_dict('regressor', 4)
    pre = 'gclstm_decoder.cell_list.0.'
    cell = HeteroPGCLSTM({'grain': 11, 'joint': 8}, 96, (['grain', 'joint'], list(ET)))
    cell.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)})
    cell = cell.to(dev())
    h0 = {t: torch.rand(v.shape[0], 96) - 0.5 for t, v in x.items()}
    c0 = {t: torch.rand(v.shape[0], 96) - 0.5 for t, v in x.items()}
    for hh0, cc0 in ((None, None), (h0, c0)):
        href, cref = orc.pgclstm_cell(sd, pre[:-1], x, ei, ea, hh0, cc0)
        hh, cc = cell(to_dev(x), to_dev(ei), to_dev(ea), None if hh0 is None else to_dev(hh0), None if cc0 is None else to_dev(cc0))
        for t in x:
            assert torch.isfinite(hh[t]).all()
            assert rel_err(hh[t], href[t]) < TOL and rel_err(cc[t], cref[t]) < TOL, (t, hh0 is None)


def test_pgclstm_cell_with_h_but_without_c():
    from graingraphnn_b200.heteropgclstm import HeteroPGCLSTM
    x, ei, ea = load_graph('c1')
    sd = orc.synth_state_dict('regressor', 1)
    pre = 'gclstm_decoder.cell_list.0.'
    cell = HeteroPGCLSTM({'grain': 11, 'joint': 8}, 96, (['grain', 'joint'], list(ET)))
    cell.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)})
    h0 = {t: torch.rand(v.shape[0], 96) - 0.5 for t, v in x.items()}
    href, cref = orc.pgclstm_cell(sd, pre[:-1], x, ei, ea, h0, None)
    hh, cc = cell.to(dev())(to_dev(x), to_dev(ei), to_dev(ea), to_dev(h0), None)
    assert all(rel_err(hh[t], href[t]) < TOL and rel_err(cc[t], cref[t]) < TOL for t in x)


def test_pgc_cell_matches_golden():
    from graingraphnn_b200.heteropgclstm import HeteroPGC
    x, ei, ea = load_graph('c1')
    g = load_golden('c1')
    h, c = orc.encode_decode(orc.synth_state_dict('regressor', 1), x, ei, ea)
    sd = orc.synth_state_dict('regressor', 4)
    pre = 'gclstm_decoder.cell_list.0.'
    cell = HeteroPGC({'grain': 11, 'joint': 8}, 96, (['grain', 'joint'], list(ET)))
    cell.load_state_dict({k[len(pre):]: v for k, v in sd.items()
                          if k.startswith(pre + 'conv_i.') or k.startswith(pre + 'b_i.')})
    hh, cc = cell.to(dev())(to_dev(x), to_dev(ei), to_dev(ea), to_dev(h), to_dev(c))
    assert rel_err(hh['joint'], g['pgc_h_joint']) < TOL and rel_err(hh['grain'], g['pgc_h_grain']) < TOL
    assert torch.equal(cc['grain'].cpu(), c['grain'])


def test_gclstm_sage_cell_matches_reference_golden():
    """HeteroGCLSTM (SAGEConv gates, heterogclstm.py:125-196) with the state_dict the reference module produced."""
    import os
    from util import GOLDEN
    from graingraphnn_b200.heterogclstm import HeteroGCLSTM
    x, ei, ea = load_graph('c1')
    g = load_golden('c1')
    h, c = orc.encode_decode(orc.synth_state_dict('regressor', 1), x, ei, ea)
    gsd = torch.load(os.path.join(GOLDEN, 'gclstm_state_dict.pt'))
    cell = HeteroGCLSTM({'grain': 11, 'joint': 8}, 96, (['grain', 'joint'], list(ET)))
    res = cell.load_state_dict(gsd)
    assert not res.missing_keys and not res.unexpected_keys
    hh, cc = cell.to(dev())(to_dev(x), to_dev(ei), to_dev(h), to_dev(c))
    assert rel_err(hh['joint'], g['gclstm_h_joint']) < TOL and rel_err(cc['grain'], g['gclstm_c_grain']) < TOL
    href, cref = orc.gclstm_cell({'cell.' + k: v for k, v in gsd.items()}, 'cell', x, ei, h, c)
    assert all(rel_err(hh[t], href[t]) < TOL and rel_err(cc[t], cref[t]) < TOL for t in x)
    # zero-initialised states (h_dict = c_dict = None, heterogclstm.py:101-109)
    hh0, cc0 = cell(to_dev(x), to_dev(ei))
    href0, cref0 = orc.gclstm_cell({'cell.' + k: v for k, v in gsd.items()}, 'cell', x, ei)
    assert all(rel_err(hh0[t], href0[t]) < TOL and rel_err(cc0[t], cref0[t]) < TOL for t in x)


def test_gc_sage_layer_matches_oracle_on_ragged_graph():
    """HeteroGC = relu(HeteroConv{SAGEConv}(x)) (heterogclstm.py:236-275) incl. rows without in-edges (mean = 0)."""
    from graingraphnn_b200.heterogclstm import HeteroGC
    torch.manual_seed(3)
    n = {'grain': 57, 'joint': 130}
    x = {'grain': torch.rand(57, 11), 'joint': torch.rand(130, 8)}
    ei = {ET[0]: random_graph(57, 130, 300, 1), ET[1]: random_graph(130, 57, 200, 2, hub=5), ET[2]: random_graph(130, 130, 390, 3)}
    ei[ET[2]] = ei[ET[2]][:, ei[ET[2]][1] % 7 != 0]          # some joints receive nothing over this edge type
    cell = HeteroGC(n_in := {'grain': 11, 'joint': 8}, 64, (['grain', 'joint'], list(ET)))
    sd = cell.state_dict()
    ref = {}
    for et in ET:
        o = orc.sage_conv(sd, 'conv_i.convs.' + '__'.join(et), x[et[0]], x[et[2]], ei[et])
        ref[et[2]] = ref.get(et[2], 0) + o
    out = cell.to(dev())(to_dev(x), to_dev(ei))
    assert set(out) == {'grain', 'joint'} and out['joint'].shape == (130, 64)
    assert all(rel_err(out[t], torch.relu(ref[t])) < TOL for t in n)


# ------------------------------------------------------------------------------------------------ models
@pytest.mark.parametrize('gemm', ['tc', 'simt'])
@pytest.mark.parametrize('name', ['c1', 'c2'])
def test_regressor_and_classifier_forward_match_reference_golden(name, gemm, monkeypatch):
    monkeypatch.setenv('GG_GEMM', gemm)      # tcgen05 3xTF32 GEMMs vs fp32 CUDA-core GEMMs: both must meet the bar
    x, ei, ea = load_graph(name)
    g = load_golden(name)
    R, C = models()
    xd, eid, ead = to_dev(x), to_dev(ei), to_dev(ea)
    y = R(xd, eid, ead)
    yc = C(xd, eid, ead)
    for k, v in (('r_joint', y['joint']), ('r_grain', y['grain']), ('r_grain_area', y['grain_area']),
                 ('c_edge_event', yc['edge_event']), ('c_edge', yc['edge'])):
        assert rel_err(v, g[k]) < TOL, (k, rel_err(v, g[k]))
    # event decisions (models.py:626-628: sigmoid(edge_event) > 0.6; test.py:418: area < 1e-4) are identical
    assert torch.equal(torch.sigmoid(yc['edge_event']).cpu() > 0.6, torch.sigmoid(g['c_edge_event']) > 0.6)
    assert torch.equal(y['grain_area'].cpu() < 1e-4, g['r_grain_area'] < 1e-4)


@pytest.mark.parametrize('rawh', ['1', '0'])
def test_engine_three_rollout_steps_match_reference_golden(rawh, monkeypatch):
    """RolloutEngine (resident state + fused step) over 3 steps vs reference Rmodel.update + edge-attr rebuild; decoder gather
    in the raw-score form ([input | V] rows, default) and in the K | V form (GG_RAWH=0)."""
    from graingraphnn_b200.engine import RolloutEngine
    monkeypatch.setenv('GG_RAWH', rawh)
    x, ei, ea = load_graph('c1')
    g = load_golden('c1')
    eng = RolloutEngine.from_state_dicts(orc.synth_state_dict('regressor', 1), orc.synth_state_dict('classifier', 2), dev())
    eng.set_graph(to_dev(x), to_dev(ei), to_dev(ea))
    for step in range(3):
        pred = eng.step(span=6)
        assert rel_err(pred['edge_event'], g[f'step{step}_edge_event']) < TOL
        assert rel_err(pred['grain_area'], g[f'step{step}_grain_area']) < TOL
        assert rel_err(eng.x['joint'], g[f'step{step}_x_joint']) < TOL
        assert rel_err(eng.x['grain'], g[f'step{step}_x_grain']) < TOL
        assert rel_err(eng.edge_attr[ET[2]], g[f'step{step}_ea_jj']) < TOL
        assert torch.equal(torch.sigmoid(pred['edge_event']).cpu() > 0.6, torch.sigmoid(g[f'step{step}_edge_event']) > 0.6)


def test_engine_cuda_graph_replay_equals_eager():
    from graingraphnn_b200.engine import RolloutEngine
    x, ei, ea = load_graph('c1')
    sd_r, sd_c = orc.synth_state_dict('regressor', 1), orc.synth_state_dict('classifier', 2)
    outs = []
    for use_graph in (False, True):
        eng = RolloutEngine.from_state_dicts(sd_r, sd_c, dev())
        eng.set_graph(to_dev(x), to_dev(ei), to_dev(ea))
        if use_graph:
            eng.capture(span=6, warmup=1)      # the warm-up step runs; the capture itself executes nothing
            pred = eng.step(6)
        else:
            for _ in range(2):
                pred = eng.step(6)
        torch.cuda.synchronize()
        outs.append((pred['edge_event'].clone(), eng.x['joint'].clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


def test_feature_update_exact_including_z_clamp():
    from graingraphnn_b200.heads import feature_update
    x, _, _ = load_graph('c1')
    for z0 in (0.0, 118.0 / 121.0):
        xo = {k: v.clone() for k, v in x.items()}
        xo['grain'][:, 2] = z0; xo['joint'][:, 2] = z0
        y = {'joint': torch.rand(236, 2) - 0.5, 'grain': torch.rand(118, 2) - 0.5}
        xd = to_dev(xo)
        feature_update(xd['joint'], xd['grain'], y['joint'].to(dev()), y['grain'].to(dev()), 6 / 121, 120 / 121)
        orc.regressor_update(xo, y, span=6)
        assert torch.equal(xd['joint'].cpu(), xo['joint']) and torch.equal(xd['grain'].cpu(), xo['grain'])


def test_full_size_properties_100k_grains():
    """At a BASELINE-size domain (C3, ~10^5 grains) the oracle is too slow; check size-independent properties instead:
    permutation invariance of the edge order, zero attention leakage across rows, finite outputs."""
    from graingraphnn_b200.synth import honeycomb_graph
    from graingraphnn_b200.engine import RolloutEngine
    x, ei = honeycomb_graph(320, 320, seed=0)            # 102,400 grains
    sd_r, sd_c = orc.synth_state_dict('regressor', 1), orc.synth_state_dict('classifier', 2)
    eng = RolloutEngine.from_state_dicts(sd_r, sd_c, dev())
    eng.set_graph(to_dev(x), to_dev(ei))
    p1 = {k: v.clone() for k, v in eng.step(6).items()}
    gen = torch.Generator().manual_seed(1)
    ei2, perms = {}, {}
    for e in ET:
        perms[e] = torch.randperm(ei[e].shape[1], generator=gen)
        ei2[e] = ei[e][:, perms[e]]
    eng2 = RolloutEngine.from_state_dicts(sd_r, sd_c, dev())
    eng2.set_graph(to_dev(x), to_dev(ei2))
    p2 = eng2.step(6)
    assert all(torch.isfinite(v).all() for v in p1.values())
    # node outputs: same up to summation order inside a row; edge outputs: permuted the same way
    assert rel_err(p2['joint'], p1['joint']) < 1e-5 and rel_err(p2['grain_area'], p1['grain_area']) < 1e-5
    assert rel_err(p2['edge_event'], p1['edge_event'][perms[ET[2]].to(dev())]) < 1e-5


def test_ensemble_of_rollouts_with_per_graph_span_matches_oracle_per_member():
    """BASELINE config 5: a block-diagonal batch of independent rollouts (different sizes, different spans) steps every
    member exactly as the oracle steps it alone; two steps, the second one through the clamp of z (test.py:405-407)."""
    from graingraphnn_b200.ensemble import EnsembleEngine
    sd_r, sd_c = orc.synth_state_dict('regressor', 1), orc.synth_state_dict('classifier', 2)
    members, spans = [], (6, 12, 120)
    for i, name in enumerate(('c1', 'c2', 'c1')):
        x, ei, ea = load_graph(name)
        x = {t: v.clone() for t, v in x.items()}
        g = torch.Generator().manual_seed(100 + i)
        x['joint'][:, :2] = (x['joint'][:, :2] + 0.01 * torch.rand(x['joint'].shape[0], 2, generator=g)) % 1.0
        members.append((x, ei, None))
    eng = EnsembleEngine.from_state_dicts(sd_r, sd_c, device=dev())
    eng.set_graphs([(to_dev(x), to_dev(ei), None) for x, ei, _ in members])
    ref_x = [{t: v.clone() for t, v in m[0].items()} for m in members]
    ref_ea = [orc.edge_attr_rebuild(m[0], m[1]) for m in members]
    for step in range(2):
        pred = eng.split(eng.step(list(spans)))
        for i, (m, sp) in enumerate(zip(members, spans)):
            ref, ref_ea[i] = orc.nn_step(sd_r, sd_c, ref_x[i], m[1], ref_ea[i], sp)
            for k in ('joint', 'grain', 'grain_area', 'edge_event'):
                assert rel_err(pred[i][k], ref[k]) < TOL, (step, i, k)
            got = eng.member_features(i)
            for t in ('joint', 'grain'):
                assert rel_err(got[t], ref_x[i][t]) < TOL, (step, i, t)
    assert abs(float(eng.member_features(2)['grain'][0, 2]) - 120 / 121) < 1e-6      # span 120 twice: clamped


def test_engine_topology_change_between_steps_matches_oracle():
    """Dynamic topology (models.py:840-841 rebinds new edge_index tensors after an elimination): step, drop every edge that
    touches two grains (their rows stay, with in-degree 0: dead nodes keep flowing through the NN), set_topology, step again."""
    from graingraphnn_b200.engine import RolloutEngine
    x, ei, ea = load_graph('c1')
    sd_r, sd_c = orc.synth_state_dict('regressor', 1), orc.synth_state_dict('classifier', 2)
    eng = RolloutEngine.from_state_dicts(sd_r, sd_c, dev())
    eng.set_graph(to_dev(x), to_dev(ei), to_dev(ea))
    xo = {k: v.clone() for k, v in x.items()}
    pred = eng.step(6)
    ref, ea1 = orc.nn_step(sd_r, sd_c, xo, ei, ea, 6)
    assert rel_err(pred['edge_event'], ref['edge_event']) < TOL
    dead = torch.tensor([5, 77])
    ei2 = {ET[0]: ei[ET[0]][:, ~torch.isin(ei[ET[0]][0], dead)], ET[1]: ei[ET[1]][:, ~torch.isin(ei[ET[1]][1], dead)],
           ET[2]: ei[ET[2]][:, 12:]}                      # also drop a few joint-joint edges: in-degree 2 rows
    eng.set_topology(to_dev(ei2))
    pred = eng.step(6)
    ref, _ = orc.nn_step(sd_r, sd_c, xo, ei2, orc.edge_attr_rebuild(xo, ei2), 6)
    for k in ('joint', 'grain', 'grain_area', 'edge_event'):
        assert torch.isfinite(pred[k]).all() and rel_err(pred[k], ref[k]) < TOL, k
    assert rel_err(eng.x['grain'], xo['grain']) < TOL and rel_err(eng.x['joint'], xo['joint']) < TOL
    assert pred['edge_event'].shape[0] == ei2[ET[2]].shape[1]


# ------------------------------------------------------------------------------------------ geometry feedback (row f2)
@pytest.mark.parametrize('name,factor', [('c1', 1), ('c2', 3), ('syn', 1)])
def test_region_center_bit_exact_vs_reference_gnn_update(name, factor):
    """gg_joint_rank + gg_region_key + gg_region_center against the output of the reference's own GNN_update / graph.update
    (tests/golden/geometry_golden.npz): index arrays, float64 centres and the fp32 write-back, all bit-exact."""
    from graingraphnn_b200.geometry import RegionIndex, region_center
    from test_geometry_core import load_case, region_index_numpy
    c = load_case(name)
    ng, nj = c['xg0'].shape[0], c['xj'].shape[0]
    idx = RegionIndex(torch.from_numpy(c['gj']).to(dev()), ng, nj)
    rowptr, col, key, rank = region_index_numpy(c['gj'], ng, nj)
    assert np.array_equal(idx.rowptr.cpu().numpy(), rowptr) and np.array_equal(idx.col.cpu().numpy(), col)
    assert np.array_equal(idx.rank.cpu().numpy()[:nj], rank) and np.array_equal(idx.key.cpu().numpy()[:len(key)], key)
    xj = torch.from_numpy(c['xj']).to(dev())
    xg = torch.from_numpy(c['xg0'].copy()).to(dev())
    off = None if c['off'] is None else torch.from_numpy(c['off']).to(dev())
    centers = region_center(xj, idx, xg, off, factor)
    assert np.array_equal(centers.cpu().numpy(), c['center'], equal_nan=True)
    assert np.array_equal(xg.cpu().numpy(), c['xg_out'])
    assert torch.equal(xj.cpu(), torch.from_numpy(c['xj']))                  # the joint rows are read only
    grain_of = np.repeat(np.arange(ng), np.diff(rowptr))
    assert np.array_equal(idx.col_sorted.cpu().numpy()[:len(col)], col[np.lexsort((key, grain_of))])      # gg_region_sort
    c_key = region_center(xj, idx, None, off, factor, presorted=False)                                    # walk by key
    assert torch.equal(torch.nan_to_num(c_key, nan=-7.0), torch.nan_to_num(centers, nan=-7.0))
    # centres only (no write-back), and write-back only (no centres)
    c2 = region_center(xj, idx, None, off, factor)
    assert torch.equal(torch.nan_to_num(c2, nan=-7.0), torch.nan_to_num(centers, nan=-7.0))
    xg2 = torch.from_numpy(c['xg0'].copy()).to(dev())
    assert region_center(xj, idx, xg2, off, factor, want_centers=False) is None and torch.equal(xg2, xg)


def test_engine_steps_with_geometry_feedback_match_oracle():
    """Rollout steps in the reference's order (test.py:382-407, :471-476, :556-575): NN step, feature update, grain centres
    from the moved joints, write-back, edge lengths from the new coordinates — eager and replayed from a CUDA graph."""
    from graingraphnn_b200.engine import RolloutEngine
    x, ei, ea = load_graph('c1')
    sd_r, sd_c = orc.synth_state_dict('regressor', 1), orc.synth_state_dict('classifier', 2)
    xo = {k: v.clone() for k, v in x.items()}
    eao = {k: v.clone() for k, v in ea.items()}
    ref = []
    for step in range(3):
        pred = orc.regressor_forward(sd_r, xo, ei, eao)
        pred.update(orc.classifier_forward(sd_c, xo, ei, eao))
        orc.regressor_update(xo, pred, 6)
        cen = orc.region_center(xo['joint'], ei[ET[0]], xo['grain'].shape[0])
        orc.grain_xy_writeback(xo['grain'], cen)
        eao = orc.edge_attr_rebuild(xo, ei)
        ref.append((pred['edge_event'].clone(), xo['grain'].clone(), xo['joint'].clone(), cen.copy(), eao[ET[1]].clone()))
    for use_graph in (False, True):
        eng = RolloutEngine.from_state_dicts(sd_r, sd_c, dev())
        eng.set_graph(to_dev(x), to_dev(ei), to_dev(ea))
        eng.enable_geometry_feedback()
        first = 0
        if use_graph:
            eng.capture(span=6, warmup=1)
            first = 1                                   # the warm-up step was step 0
        for step in range(first, 3):
            pred = eng.step(6)
            ev, xg, xj, cen, ea_jg = ref[step]
            assert rel_err(pred['edge_event'], ev) < TOL
            assert rel_err(eng.x['grain'], xg) < TOL and rel_err(eng.x['joint'], xj) < TOL
            assert np.nanmax(np.abs(eng.centers.cpu().numpy() - cen)) < 1e-5
            assert rel_err(eng.edge_attr[ET[1]], ea_jg) < TOL
        # the centres written are exactly the centres of the joints the engine holds
        cen_exact = orc.region_center(eng.x['joint'].cpu(), ei[ET[0]], 118)
        assert np.array_equal(eng.centers.cpu().numpy(), cen_exact, equal_nan=True)
        assert torch.equal(eng.x['grain'][:, :2].cpu(), torch.from_numpy(cen_exact).float())


def test_region_center_full_size_properties_100k_grains():
    """~10^5 grains: against an independent vectorised fp64 formulation on the device (every joint unwrapped against the
    grain's first joint; equal to the chain unwrap for grains narrower than half the domain), compared on the circle."""
    from graingraphnn_b200.synth import honeycomb_graph
    from graingraphnn_b200.geometry import RegionIndex, region_center
    x, ei = honeycomb_graph(320, 320, seed=0)
    xj, xg = x['joint'].to(dev()), x['grain'].to(dev())
    gj = ei[ET[0]].to(dev())
    ng, nj = xg.shape[0], xj.shape[0]
    idx = RegionIndex(gj, ng, nj)
    xg_out = xg.clone()
    centers = region_center(xj, idx, xg_out)
    deg = (idx.rowptr[1:] - idx.rowptr[:-1]).long()
    assert int(deg.min()) >= 2 and not torch.isnan(centers).any()
    grain_of = torch.repeat_interleave(torch.arange(ng, device=dev()), deg)
    key = idx.key[:gj.shape[1]].long()
    first_key = torch.full((ng,), 2 ** 31 - 1, device=dev(), dtype=torch.long).scatter_reduce(0, grain_of, key, 'amin')
    first_joint = gj[1][first_key]                                           # key = edge position of the first appearance
    p = xj[idx.col.long(), :2].double()
    p0 = xj[first_joint, :2].double()[grain_of]
    d = p - p0
    d = d - torch.round(d)
    mean = torch.zeros(ng, 2, dtype=torch.float64, device=dev()).index_add_(0, grain_of, d) / deg[:, None] + xj[first_joint, :2].double()
    diff = centers - mean
    diff = diff - torch.round(diff)
    # not 1e-12: the reference keeps the FIRST vertex of a region in float32, so its +1 seam shift rounds at fp32 precision
    # (graph_datastruct.py:701-703 on an np.float32 scalar) and moves the mean by <= 2^-24 — reproduced, not corrected
    assert float(diff.abs().max()) < 6e-8
    assert torch.equal(xg_out[:, :2], centers.float()) and torch.equal(xg_out[:, 2:], xg[:, 2:])
    assert float(centers.min()) > -1e-12 and float(centers.max()) < 2.0
