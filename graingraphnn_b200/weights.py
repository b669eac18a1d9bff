"""Seeded stand-in weights with the reference's `state_dict` layout, for benchmarks and demos when `model/regressor0.pt` /
`model/classifier1.pt` are not at hand (they are absent from the reference mount, .MISSING_LARGE_BLOBS:2-3).

The key layout follows the module registration order of models.py:351-399 / :529-570, heteropgclstm.py:48-82 and
periodGATconv.py:119-143; `load_weights` loads the real files unchanged when a path is given.
Values: U(-g / sqrt(fan_in), g / sqrt(fan_in)) like torch / PyG `Linear` defaults, from one seeded generator in key order
(the test oracle draws its stand-ins the same way, so equal seeds give equal weights)."""
import math

import torch

GATES = ('i', 'f', 'c', 'o')
EDGE_TYPES = (('grain', 'push', 'joint'), ('joint', 'pull', 'grain'), ('joint', 'connect', 'joint'))


def state_dict_shapes(kind='regressor', C=96, f_grain=11, f_joint=8, edge_types=EDGE_TYPES):
    D = {'grain': f_grain + C, 'joint': f_joint + C}
    shapes = {}
    for part in ('gclstm_encoder', 'gclstm_decoder'):
        p = f'{part}.cell_list.0'
        for g in GATES:
            for et in edge_types:
                q = f'{p}.conv_{g}.convs.{"__".join(et)}'
                s, d = D[et[0]], D[et[2]]
                for lin, shp in (('lin_key', (C, s)), ('lin_query', (C, d)), ('lin_value', (C, s)), ('lin_l2', (C, C))):
                    shapes[f'{q}.{lin}.weight'], shapes[f'{q}.{lin}.bias'] = shp, (C,)
                shapes[f'{q}.lin_edge.weight'] = (C, 1)
                shapes[f'{q}.lin_skip.weight'], shapes[f'{q}.lin_skip.bias'] = (C, d), (C,)
            for t in ('grain', 'joint'):
                shapes[f'{p}.b_{g}.{t}'] = (1, C)
    if kind == 'regressor':
        for t in ('grain', 'joint'):
            shapes[f'linear.{t}.weight'], shapes[f'linear.{t}.bias'] = (2, C), (2,)
    else:
        shapes['lin1.weight'], shapes['lin1.bias'] = (2, 2 * C + 1), (2,)
        shapes['lin2.weight'], shapes['lin2.bias'] = (1, 2 * C + 1), (1,)
    return shapes


def synth_state_dict(kind='regressor', seed=0, gain=1.0, dtype=torch.float32, head_gain=1.0, **kw):
    """head_gain scales the regressor's output heads (`linear.{grain,joint}`): untrained heads move every joint by ~0.1 patch per
    step in one direction (tanh(.) / 5, models.py:503-516), which tears the tiling apart within five steps (every grain-joint edge
    then crosses a patch boundary); a trained model moves joints by O(1e-3).  Benchmarks use head_gain = 0.02 so that the
    geometry stays a valid tiling over the timed steps; parity tests keep 1.0 (the oracle's stand-ins)."""
    gen = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in state_dict_shapes(kind, **kw).items():
        bound = gain if k.endswith('lin_edge.weight') else gain / math.sqrt(max(shp[-1], 1))
        sd[k] = ((torch.rand(shp, generator=gen, dtype=torch.float64) * 2 - 1) * bound).to(dtype)
        if k.startswith('linear.'):
            sd[k] = sd[k] * head_gain
    return sd


def load_weights(regressor_pt=None, classifier_pt=None, seeds=(1, 2), head_gain=1.0):
    """(sd_regressor, sd_classifier, description): the reference's files when given (test.py:178, :183), else seeded stand-ins."""
    if regressor_pt and classifier_pt:
        return (torch.load(regressor_pt, map_location='cpu'), torch.load(classifier_pt, map_location='cpu'),
                f'{regressor_pt} / {classifier_pt}')
    return (synth_state_dict('regressor', seeds[0], head_gain=head_gain), synth_state_dict('classifier', seeds[1]),
            'seeded stand-ins with the reference state_dict layout (regressor0.pt/classifier1.pt absent)'
            + (f'; regressor heads scaled by {head_gain} (per-step joint displacement of the order a trained model produces)' if head_gain != 1.0 else ''))
