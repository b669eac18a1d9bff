"""CPU: the oracle against the golden vectors produced by the reference's own modules (oracle/make_golden.py) and
against the KATs the reference holds for this path."""
import math
import os

import numpy as np
import pytest
import torch

import grain_oracle as orc
from util import ET, GOLDEN, SHORT, load_golden, load_graph, rel_err


@pytest.mark.parametrize('name', ['c1', 'c2'])
def test_forward_matches_reference_modules(name):
    x, ei, ea = load_graph(name)
    g = load_golden(name)
    sd_r, sd_c = orc.synth_state_dict('regressor', 1), orc.synth_state_dict('classifier', 2)
    y, h, c = orc.regressor_forward(sd_r, x, ei, ea, return_state=True)
    yc, hc, _ = orc.classifier_forward(sd_c, x, ei, ea, return_state=True)
    pairs = [('r_joint', y['joint']), ('r_grain', y['grain']), ('r_grain_area', y['grain_area']),
             ('c_edge_event', yc['edge_event']), ('c_edge', yc['edge']), ('r_dec_h_joint', h['joint']),
             ('r_dec_h_grain', h['grain']), ('r_dec_c_joint', c['joint']), ('r_dec_c_grain', c['grain']),
             ('c_dec_h_joint', hc['joint'])]
    for k, v in pairs:
        assert torch.equal(v, g[k]), (k, rel_err(v, g[k]))   # same op order on the same torch build: bit-identical


def test_single_conv_pgc_and_sage_cells():
    x, ei, ea = load_graph('c1')
    g = load_golden('c1')
    sd_r = orc.synth_state_dict('regressor', 1)
    h, c = orc.encode_decode(sd_r, x, ei, ea)
    xin = {t: torch.cat([x[t], h[t]], 1) for t in x}
    sd3 = orc.synth_state_dict('regressor', 3)
    for tag, weighted in (('gat', True), ('sum', False)):
        for et in (ET[0], ET[2]):
            out = orc.period_conv(sd3, f'gclstm_decoder.cell_list.0.conv_i.convs.{"__".join(et)}',
                                  xin[et[0]], xin[et[2]], ei[et], ea[et], weighted)
            assert torch.equal(out, g[f'conv_{tag}_{SHORT[et]}'])
    sd4 = orc.synth_state_dict('regressor', 4)
    hh, _ = orc.pgc_cell(sd4, 'gclstm_decoder.cell_list.0', x, ei, ea, h, c)
    assert torch.equal(hh['joint'], g['pgc_h_joint']) and torch.equal(hh['grain'], g['pgc_h_grain'])
    gsd = torch.load(os.path.join(GOLDEN, 'gclstm_state_dict.pt'))
    hh, cc = orc.gclstm_cell({'cell.' + k: v for k, v in gsd.items()}, 'cell', x, ei, h, c)
    assert rel_err(hh['joint'], g['gclstm_h_joint']) < 1e-6 and rel_err(cc['grain'], g['gclstm_c_grain']) < 1e-6


def test_three_rollout_steps_match_reference_update():
    x, ei, ea = load_graph('c1')
    g = load_golden('c1')
    sd_r, sd_c = orc.synth_state_dict('regressor', 1), orc.synth_state_dict('classifier', 2)
    for step in range(3):
        pred, ea = orc.nn_step(sd_r, sd_c, x, ei, ea, span=6)
        assert torch.equal(pred['edge_event'], g[f'step{step}_edge_event'])
        assert torch.equal(pred['grain_area'], g[f'step{step}_grain_area'])
        assert torch.equal(x['joint'], g[f'step{step}_x_joint']) and torch.equal(x['grain'], g[f'step{step}_x_grain'])
        assert torch.equal(ea[ET[2]], g[f'step{step}_ea_jj'])


def test_parameter_counts_match_reference_logfiles():
    # model/regressor0_logfile:40 and model/classifier1_logfile:40
    assert sum(math.prod(s) for s in orc.param_shapes('regressor').values()) == 1204612
    assert sum(math.prod(s) for s in orc.param_shapes('classifier').values()) == 1204806


def test_edge_length_formula_reproduces_pickled_edge_weights():
    # KAT (2) of SURVEY.md §8c: test.py:562-575 applied to the pickled coordinates == pickled edge_weight_dicts
    x, ei, ea = load_graph('c1')
    new = orc.edge_attr_rebuild(x, ei)
    for et in ET:
        assert float((new[et] - ea[et]).abs().max()) < 1e-7


def test_fixture_graph_invariants():
    for name, ng in (('c1', 118), ('c2', 1043)):
        x, ei, _ = load_graph(name)
        assert x['grain'].shape == (ng, 11) and x['joint'].shape == (2 * ng, 8)
        for et in ET:
            assert ei[et].shape == (2, 6 * ng)
        for et in (ET[0], ET[2]):   # every joint has exactly 3 grain and 3 joint in-edges
            assert (np.bincount(ei[et][1].numpy(), minlength=2 * ng) == 3).all()


def test_csr_by_dst_is_stable():
    ei = torch.tensor([[5, 4, 3, 2, 1, 0], [1, 0, 1, 0, 1, 3]])
    rp, col, perm = orc.csr_by_dst(ei, 5)
    assert rp.tolist() == [0, 2, 5, 5, 6, 6] and col.tolist() == [4, 2, 5, 3, 1, 0] and perm.tolist() == [1, 3, 0, 2, 4, 5]


def test_zero_in_degree_rows_are_skip_plus_bias():
    x, ei, ea = load_graph('c1')
    sd = orc.synth_state_dict('regressor', 3)
    p = 'gclstm_encoder.cell_list.0.conv_i.convs.joint__connect__joint'
    xin = torch.cat([x['joint'], torch.zeros(x['joint'].shape[0], 96)], 1)
    keep = ei[ET[2]][1] != 7
    out = orc.period_conv(sd, p, xin, xin, ei[ET[2]][:, keep], ea[ET[2]][keep])
    skip = xin @ sd[p + '.lin_skip.weight'].t() + sd[p + '.lin_skip.bias']
    assert torch.equal(out[7], skip[7]) and torch.isfinite(out).all()
