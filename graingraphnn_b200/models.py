"""Task models of the rollout path: `SeqGCLSTM`, `GrainNN_regressor`, `GrainNN_classifier`.

Host-side mirror of the reference's models.py for the hot path only: constructors, `forward(x_dict, edge_index_dict,
edge_attr)` and the `state_dict` layout follow models.py:151-301, :351-467, :529-611, so `regressor0.pt` /
`classifier1.pt` load unchanged.  `GrainNN_classifier.update` (the topology update, models.py:614-1053)
keeps the reference's signature and results; its decisions are made by `topology.topology_update` (host, position lists).
"""
import copy

import torch
from torch import nn

from .cell import _as_f32c, require_cuda
from .heads import edge_head, feature_update, node_head
from .heteropgclstm import HeteroPGCLSTM


class SeqGCLSTM(nn.Module):
    """Stack of graph-LSTM cells run for seq_len = 1 (models.py:151-301). Layer 0 is a HeteroPGCLSTM."""

    def __init__(self, in_channels_dict, out_channels, num_layers, metadata, device, bias=True, return_all_layers=True):
        super().__init__()
        out_channels = self._extend_for_multilayer(out_channels, num_layers)
        if not len(out_channels) == num_layers:
            raise ValueError('Inconsistent list length.')
        self.in_channels_dict, self.out_channels, self.num_layers = in_channels_dict, out_channels, num_layers
        self.metadata, self.device, self.bias, self.return_all_layers = metadata, device, bias, return_all_layers
        cells = []
        for i in range(num_layers):
            if i == 0:
                cells.append(HeteroPGCLSTM(in_channels_dict=in_channels_dict, out_channels=out_channels[i],
                                           metadata=metadata, bias=bias, device=device))
            else:
                from .heterogclstm import HeteroGCLSTM
                cur = {t: out_channels[i - 1] for t in in_channels_dict}
                cells.append(HeteroGCLSTM(in_channels_dict=cur, out_channels=out_channels[i],
                                          metadata=metadata, bias=bias, device=device))
        self.cell_list = nn.ModuleList(cells)

    def forward(self, x_dict, edge_index_dict, edge_attr, hidden_state):
        if self.num_layers > 1:
            # models.py:254-258 passes edge_attr= to HeteroGCLSTM.forward, which has no such parameter
            # (heterogclstm.py:162-168): the reference raises TypeError here too.
            raise TypeError("HeteroGCLSTM.forward() got an unexpected keyword argument 'edge_attr' "
                            '(layers > 1 is unreachable in the reference, SURVEY.md finding 1)')
        h = c = None
        if hidden_state is not None:
            h, c = hidden_state[0]
        h, c = self.cell_list[0](x_dict=x_dict, edge_index_dict=edge_index_dict, edge_attr=edge_attr, h_dict=h, c_dict=c)
        return [[h, c]]

    def _init_hidden(self, x_dict):
        return [[self.cell_list[i]._set_hidden_state(x_dict, None), self.cell_list[i]._set_hidden_state(x_dict, None)]
                for i in range(self.num_layers)]

    @staticmethod
    def _extend_for_multilayer(param, num_layers):
        return param if isinstance(param, list) else [param] * num_layers


def _in_channels(hyper):
    return {t: len(f) for t, f in hyper.features.items()}


class GrainNN_regressor(nn.Module):
    """Encoder/decoder regressor (models.py:351-467); `history` / `edge_len` variants are not on the rollout path."""

    def __init__(self, hyper, history=False, edge_len=False):
        super().__init__()
        if history or edge_len:
            raise NotImplementedError('history / edge_len variants are not used by test.py rollouts')
        self.in_channels_dict = _in_channels(hyper)
        self.out_channels, self.num_layer = hyper.layer_size, hyper.layers
        self.metadata, self.out_win = hyper.metadata, getattr(hyper, 'out_win', 1)
        self.device, self.seq_len = getattr(hyper, 'device', 'cuda'), getattr(hyper, 'window', 1)
        self.history, self.edge_len = history, edge_len
        self.gclstm_encoder = SeqGCLSTM(self.in_channels_dict, self.out_channels, self.num_layer, self.metadata, self.device)
        self.gclstm_decoder = SeqGCLSTM(self.in_channels_dict, self.out_channels, self.num_layer, self.metadata, self.device)
        self.dim = {'joint': 2, 'grain': 1}
        self.linear = nn.ModuleDict({t: nn.Linear(self.out_channels, len(targets)) for t, targets in hyper.targets.items()})
        self.scaling = {'grain': 20, 'joint': 5}

    @torch.no_grad()
    def forward(self, x_dict, edge_index_dict, edge_attr):
        hidden = self.gclstm_encoder(x_dict, edge_index_dict, edge_attr, None)          # models.py:422
        hidden = self.gclstm_decoder(x_dict, edge_index_dict, edge_attr, hidden)        # :424
        h_dict, _ = hidden[-1]
        xg = x_dict['grain']
        yj, _ = node_head(h_dict['joint'], self.linear['joint'].weight, self.linear['joint'].bias, [1, 1])       # :433,:443
        yg, area = node_head(h_dict['grain'], self.linear['grain'].weight, self.linear['grain'].bias, [1, 2],
                             area_in=_as_f32c(xg)[:, 3], area_scale=self.scaling['grain'])                        # :445-452
        return {'grain': yg, 'joint': yj, 'grain_area': area}

    @torch.no_grad()
    def update(self, x_dict, y_dict, geometry_scaling, span=None, train_frames=120):
        """models.py:473-516, periodic-domain branch.  With `span` given the z advance of test.py:401-407 is fused in."""
        if 'melt_left' in geometry_scaling:
            raise NotImplementedError('moving melt-pool window (models.py:480-498) is outside the rollout hot path')
        geometry_scaling['active_grains'] = (y_dict['grain'][:, 0] > -10).nonzero().view(-1)       # :502
        geometry_scaling['active_joints'] = (y_dict['joint'][:, 0] > -10).nonzero().view(-1)       # :503
        xj, xg = x_dict['joint'], x_dict['grain']
        require_cuda(xj, "x_dict['joint']")
        dz = 0.0 if span is None else span / (train_frames + 1)
        z_max = train_frames / (train_frames + 1) if span is not None else float('inf')
        feature_update(xj, xg, y_dict['joint'].contiguous(), y_dict['grain'].contiguous(), dz, z_max)


class GrainNN_classifier(nn.Module):
    """Edge-event classifier sharing the regressor's encoder/decoder architecture (models.py:529-611)."""

    threshold = 0.6          # edge-event probability threshold of `update`; the reference's driver sets it (test.py:187-191)

    def __init__(self, hyper, regressor=None, history=False):
        super().__init__()
        if history:
            raise NotImplementedError('history variant is not used by test.py rollouts')
        self.in_channels_dict = _in_channels(hyper)
        self.out_channels, self.num_layer = hyper.layer_size, hyper.layers
        self.metadata, self.out_win = hyper.metadata, getattr(hyper, 'out_win', 1)
        self.seq_len, self.device, self.history = getattr(hyper, 'window', 1), getattr(hyper, 'device', 'cuda'), history
        self.dim = {'joint': 2, 'grain': 1}
        self.scaling = {'grain': 20, 'joint': 5}
        if regressor:
            self.gclstm_encoder = copy.deepcopy(regressor.gclstm_encoder)               # models.py:550-552
            self.gclstm_decoder = copy.deepcopy(regressor.gclstm_decoder)
        else:
            self.gclstm_encoder = SeqGCLSTM(self.in_channels_dict, self.out_channels, self.num_layer, self.metadata, self.device)
            self.gclstm_decoder = SeqGCLSTM(self.in_channels_dict, self.out_channels, self.num_layer, self.metadata, self.device)
        self.lin1 = nn.Linear(2 * self.out_channels + 1, 2)
        self.lin2 = nn.Linear(2 * self.out_channels + 1, 1)

    @torch.no_grad()
    def forward(self, x_dict, edge_index_dict, edge_attr):
        hidden = self.gclstm_encoder(x_dict, edge_index_dict, edge_attr, None)
        hidden = self.gclstm_decoder(x_dict, edge_index_dict, edge_attr, hidden)
        h_dict, _ = hidden[-1]
        jj = ('joint', 'connect', 'joint')
        ev, ed = edge_head(h_dict['joint'], edge_index_dict[jj], edge_attr[jj],
                           self.lin1.weight, self.lin1.bias, self.lin2.weight, self.lin2.bias)           # :595-609
        return {'edge_event': ev, 'edge': ed}

    @torch.no_grad()
    def update(self, x_dict, edge_index_dict, edge_attr, y_dict, mask, geometry_scaling, nucleation_prob=0.0):
        """Topology update of a rollout step with the reference's signature and results (models.py:614-845; called at
        test.py:426): grain elimination, neighbour switching, cleanup.  Host code, like the reference's — tensors on a CUDA
        device are brought to the host and the results put back — but every lookup is answered from position lists instead
        of an O(E) scan (topology.py).  `self.threshold` is set by the caller (test.py:187; class default 0.6).  Returns
        (x_dict, edge_index_dict, switching_list); `x_dict['joint']`, `y_dict` and `mask` are updated in place as the
        reference does and the entries of `edge_index_dict` are re-bound to the cleaned-up arrays (models.py:838-841); the
        caller's edge TENSORS are left as they were (the reference also writes -1 into them on the way, :877-880 — nothing
        reads them afterwards); with nucleation (:771-835) `x_dict` / `mask` entries are re-bound to grown tensors."""
        from .topology import topology_update
        dev = x_dict['joint'].device
        host = lambda d: {k: (v.cpu() if isinstance(v, torch.Tensor) else v) for k, v in d.items()}   # noqa: E731
        xh, yh, mh, eh = host(x_dict), host(y_dict), host(mask), host(edge_index_dict)
        _, new_ei, pairs = topology_update(xh, eh, yh, mh, geometry_scaling['active_grains'].cpu(),
                                           geometry_scaling['active_joints'].cpu(), threshold=self.threshold,
                                           nucleation_prob=float(nucleation_prob))
        for t in ('joint', 'grain'):
            if xh[t].shape != x_dict[t].shape:                 # nucleation grew the node set: re-bind, as the reference does
                x_dict[t] = xh[t].to(x_dict[t].device)
                mask[t] = mh[t].to(mask[t].device)
            elif x_dict[t].device.type != 'cpu':               # on the host the tensors were edited in place already
                x_dict[t].copy_(xh[t])
                mask[t].copy_(mh[t])
            if y_dict[t].device.type != 'cpu':
                y_dict[t].copy_(yh[t])
        y_dict['grain_event'] = yh['grain_event'].to(y_dict['grain_event'].device)
        for e, v in new_ei.items():
            edge_index_dict[e] = v.to(dev)
        return x_dict, edge_index_dict, pairs.to(dev)

