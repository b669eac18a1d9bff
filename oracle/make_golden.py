"""Generate tests/golden/* by running the REFERENCE's own modules (imported unmodified from /root/reference)
on oracle/pyg_stub.  Run once in the build container:  python oracle/make_golden.py

TEST INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box, so everything the -m gpu tests need
(the two fixture graphs, converted to npz, and the reference's outputs on them) is committed under tests/golden/.
Weights: the shipped regressor0.pt / classifier1.pt are absent (.MISSING_LARGE_BLOBS:2-3), so seeded stand-in
state_dicts with the reference's exact keys (grain_oracle.synth_state_dict) are loaded with load_state_dict;
tests regenerate the same tensors from the seed instead of storing 10 MB of weights.
"""
import os
import sys

import dill
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, 'pyg_stub'))
import ref_shims  # noqa: E402

ref_shims.install_plot_stubs()
ref_shims.add_reference_to_path()
import grain_oracle as orc  # noqa: E402
from models import GrainNN_classifier, GrainNN_regressor  # noqa: E402  (reference models.py)
from parameters import classifier_transfered, regressor  # noqa: E402
from heteropgclstm import HeteroPGC  # noqa: E402
from heterogclstm import HeteroGCLSTM  # noqa: E402
import periodconv  # noqa: E402
import periodGATconv  # noqa: E402

OUT = os.path.join(HERE, '..', 'tests', 'golden')
ET = [('grain', 'push', 'joint'), ('joint', 'pull', 'grain'), ('joint', 'connect', 'joint')]
SHORT = {ET[0]: 'gj', ET[1]: 'jg', ET[2]: 'jj'}


def load_graph(path, domain_factor):
    with open(path, 'rb') as f:
        g = dill.load(f)[0]
    x = {k: torch.FloatTensor(v) for k, v in g.feature_dicts.items()}            # data_loader.py:102
    ei = {k: torch.LongTensor(v) for k, v in g.edge_index_dicts.items()}
    ea = {k: torch.FloatTensor(v) for k, v in g.edge_weight_dicts.items()}
    assert list(ei.keys()) == ET and list(x.keys()) == ['grain', 'joint']
    if domain_factor > 1:   # test.py:29-55 scale_feature_patchs, periodic branch (called at test.py:310-312)
        for k in ea:
            ea[k] *= domain_factor
        x['grain'][:, :2] *= domain_factor
        x['joint'][:, :2] *= domain_factor
        x['joint'][:, :2] = x['joint'][:, :2] - torch.floor(x['joint'][:, :2])
        x['grain'][:, :2] = x['grain'][:, :2] - (x['grain'][:, :2] - x['grain'][:, :2] % 1)
    return g, x, ei, ea


def build_models(g):
    hp, hpc = regressor(0), classifier_transfered(1)                                 # test.py:162-163
    hp.metadata = (['grain', 'joint', 'mask'], ET); hp.features = g.features; hp.targets = g.targets; hp.device = 'cpu'
    hpc.metadata = hp.metadata; hpc.features = hp.features; hpc.device = 'cpu'
    R = GrainNN_regressor(hp)
    R.load_state_dict(orc.synth_state_dict('regressor', seed=1)); R.eval()           # test.py:177-179
    C = GrainNN_classifier(hpc, R)
    C.load_state_dict(orc.synth_state_dict('classifier', seed=2)); C.eval()          # test.py:182-184
    assert sum(p.numel() for p in R.parameters()) == 1204612                          # regressor0_logfile:40
    assert sum(p.numel() for p in C.parameters()) == 1204806                          # classifier1_logfile:40
    return R, C


def states(model, x, ei, ea):
    enc = model.gclstm_encoder(x, ei, ea, None)
    dec = model.gclstm_decoder(x, ei, ea, enc)
    return enc[-1], dec[-1]


@torch.no_grad()
def main():
    os.makedirs(OUT, exist_ok=True)
    cases = {'c1': ('/root/reference/graphs/40_40/seed10020_G1.904_R0.558_span6.pkl', 1),
             'c2': ('/root/reference/graphs/120_120/seed0_G10.0_R2.0_span6.pkl', 3)}
    for name, (path, factor) in cases.items():
        g, x, ei, ea = load_graph(path, factor)
        graph = {'x_grain': x['grain'].numpy(), 'x_joint': x['joint'].numpy()}
        for et in ET:
            graph['ei_' + SHORT[et]] = ei[et].numpy().astype(np.int32)
            graph['ea_' + SHORT[et]] = ea[et].numpy()
        # pickled edge lengths before patch scaling (KAT for the a11 formula)
        np.savez_compressed(os.path.join(OUT, f'{name}_graph.npz'), **graph)

        R, C = build_models(g)
        gold = {}
        (he, ce), (hd, cd) = states(R, x, ei, ea)
        y = R({k: v.clone() for k, v in x.items()}, ei, ea)
        yc = C(x, ei, ea)
        (hce, cce), (hcd, ccd) = states(C, x, ei, ea)
        gold.update({'r_joint': y['joint'], 'r_grain': y['grain'], 'r_grain_area': y['grain_area'],
                     'c_edge_event': yc['edge_event'], 'c_edge': yc['edge'],
                     'r_enc_h_joint': he['joint'], 'r_enc_c_grain': ce['grain'],
                     'r_dec_h_joint': hd['joint'], 'r_dec_h_grain': hd['grain'],
                     'r_dec_c_joint': cd['joint'], 'r_dec_c_grain': cd['grain'],
                     'c_dec_h_joint': hcd['joint']})
        if name == 'c1':
            # single PeriodConv calls (attention + un-weighted variants) on the g->j and j->j types
            xin = {t: torch.cat([x[t], hd[t]], 1) for t in x}
            for mod, tag in ((periodGATconv, 'gat'), (periodconv, 'sum')):
                for et in (ET[0], ET[2]):
                    conv = mod.PeriodConv(in_channels=(-1, -1), out_channels=96)
                    sd = {k.split('lin_', 1)[0] + 'lin_' + k.split('lin_', 1)[1]: v for k, v in
                          orc.synth_state_dict('regressor', seed=3).items()
                          if k.startswith(f'gclstm_decoder.cell_list.0.conv_i.convs.{"__".join(et)}.')}
                    sd = {k.rsplit('.', 2)[-2] + '.' + k.rsplit('.', 2)[-1]: v for k, v in sd.items()}
                    conv.load_state_dict(sd)
                    xx = xin[et[0]] if et[0] == et[2] else (xin[et[0]], xin[et[2]])
                    gold[f'conv_{tag}_{SHORT[et]}'] = conv(xx, ei[et], ea[et])
            # HeteroPGC (single relu conv layer) and HeteroGCLSTM (SAGEConv gates)
            pgc = HeteroPGC({'grain': 11, 'joint': 8}, 96, (['grain', 'joint'], ET))
            sd = {k[len('gclstm_decoder.cell_list.0.'):]: v for k, v in orc.synth_state_dict('regressor', seed=4).items()
                  if k.startswith('gclstm_decoder.cell_list.0.') and ('.conv_i.' in k or '.b_i.' in k)}
            pgc.load_state_dict(sd)
            hh, _ = pgc(x, ei, ea, hd, cd)
            gold['pgc_h_joint'], gold['pgc_h_grain'] = hh['joint'], hh['grain']
            torch.manual_seed(5)
            gcl = HeteroGCLSTM({'grain': 11, 'joint': 8}, 96, (['grain', 'joint'], ET))
            hh, cc = gcl(x, ei, hd, cd)            # materialises the lazy SAGEConv weights with seed 5
            gsd = {k: v.clone() for k, v in gcl.state_dict().items()}
            torch.save(gsd, os.path.join(OUT, 'gclstm_state_dict.pt'))
            gold['gclstm_h_joint'], gold['gclstm_c_grain'] = hh['joint'], cc['grain']
            # three fixed-topology rollout steps: reference Rmodel.update (models.py:473-516) + the
            # test.py:401-407 z update + test.py:562-575 edge-attr rebuild (restated: test.py is a script)
            xs = {k: v.clone() for k, v in x.items()}
            eas = {k: v.clone() for k, v in ea.items()}
            for step in range(3):
                pred = R(xs, ei, eas)
                pred.update(C(xs, ei, eas))
                R.update(xs, pred, {'domain_offset': 0, 'domain_factor': 1})
                xs['grain'][:, 2] += 6 / 121
                xs['joint'][:, 2] += 6 / 121
                eas = {}
                for et, index in ei.items():
                    rel = xs[et[0]][index[0], :2] - xs[et[-1]][index[-1], :2]
                    rel = -1 * (rel > 0.5) + 1 * (rel < -0.5) + rel
                    eas[et] = torch.sqrt(rel[:, 0] ** 2 + rel[:, 1] ** 2).view(-1, 1)
                gold[f'step{step}_edge_event'] = pred['edge_event']
                gold[f'step{step}_grain_area'] = pred['grain_area']
                gold[f'step{step}_x_joint'] = xs['joint'].clone()
                gold[f'step{step}_x_grain'] = xs['grain'].clone()
                gold[f'step{step}_ea_jj'] = eas[ET[2]].clone()
        np.savez_compressed(os.path.join(OUT, f'{name}_golden.npz'), **{k: v.numpy() for k, v in gold.items()})
        print(name, {k: tuple(v.shape) for k, v in gold.items()})


if __name__ == '__main__':
    main()
