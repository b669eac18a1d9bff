"""Shared helpers for the test-suite (graph fixtures, tolerances)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
ET = [('grain', 'push', 'joint'), ('joint', 'pull', 'grain'), ('joint', 'connect', 'joint')]
SHORT = {ET[0]: 'gj', ET[1]: 'jg', ET[2]: 'jj'}


def load_graph(name, dtype=torch.float32, device='cpu'):
    z = np.load(os.path.join(GOLDEN, f'{name}_graph.npz'))
    x = {'grain': torch.from_numpy(z['x_grain']).to(dtype).to(device),
         'joint': torch.from_numpy(z['x_joint']).to(dtype).to(device)}
    ei = {et: torch.from_numpy(z['ei_' + SHORT[et]].astype(np.int64)).to(device) for et in ET}
    ea = {et: torch.from_numpy(z['ea_' + SHORT[et]]).to(dtype).to(device) for et in ET}
    return x, ei, ea


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, f'{name}_golden.npz'))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def rel_err(a, b):
    """max |a-b| / max(|b|) — the 'relative' of the 1e-4 bar: scaled by the tensor's own magnitude."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))
