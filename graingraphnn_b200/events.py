"""Event candidates of a rollout step on the device (first stage).

The reference decides on the host, from the full prediction arrays, which joint-joint edges switch
(`L1 = ((sigmoid(edge_event) > threshold) & (src < dst)).nonzero()`, models.py:627-629, later sorted by probability
:730-731) and which grains vanish (`((mask > 0) & (grain_area < threshold)).nonzero()` sorted by area, test.py:414-416).
`EventSelector` leaves only those candidates on the device (gg_select_events: one streaming pass per array) so that a step
ships a few (id, value) pairs to the host topology update instead of 7 floats per grain.  The probability test runs on the
logit: the device keeps every logit >= the smallest float32 that `torch.sigmoid` (CPU — the reference's arithmetic) can put
above the threshold, and the host applies the reference's own `sigmoid(v) > threshold` to those few survivors.  Identical to
the reference outside a band of a few ulps around the crossing, where torch's vectorised and scalar sigmoid paths disagree
with each other (the reference's decision there depends on the element's position in the tensor; `sigmoid_threshold_band`).
No CPU fallback for the selection itself.
"""
import numpy as np
import torch

from . import _lib
from ._lib import check, ptr


def sigmoid_threshold_band(threshold, window=2048):
    """(x_lo, x_hi): every float32 x < x_lo has torch.sigmoid(x) <= threshold and every x > x_hi has sigmoid(x) > threshold,
    on the CPU (the reference's arithmetic), whichever of torch's code paths evaluates it.  torch's CPU sigmoid is monotone
    only up to an ulp AND its vectorised body and scalar tail round differently, so for the few floats inside [x_lo, x_hi]
    the reference's own decision depends on where in the tensor the element sits; outside the band it does not.
    A bisection locates the crossing, then every float32 within `window` ulps of it is evaluated on both paths."""
    f = lambda v: bool(torch.sigmoid(torch.tensor([v], dtype=torch.float32))[0] > threshold)   # noqa: E731
    lo, hi = np.float32(-90.0), np.float32(90.0)
    if f(float(lo)) or not f(float(hi)):
        raise ValueError(f'probability threshold {threshold} is outside (0, 1)')
    while True:
        mid = np.float32((np.float64(lo) + np.float64(hi)) / 2)
        if mid == lo or mid == hi:
            break
        if f(float(mid)):
            hi = mid
        else:
            lo = mid
    b = int(np.float32(hi).view(np.int32))
    key = b if b >= 0 else -(b & 0x7fffffff)                                  # monotone integer key of a float32
    keys = np.arange(key - window, key + window + 1, dtype=np.int64)
    bits = np.where(keys >= 0, keys, (-keys) | 0x80000000).astype(np.uint32)
    xs = torch.from_numpy(bits.view(np.float32).copy())
    vec = torch.sigmoid(torch.cat([xs, xs.new_zeros(64)]))[:len(xs)] > threshold            # vectorised body
    sca = torch.stack([torch.sigmoid(xs[i:i + 1])[0] for i in range(len(xs))]) > threshold   # scalar tail
    any_pass, all_pass = (vec | sca).numpy(), (vec & sca).numpy()
    first = int(np.argmax(any_pass))
    last_fail = len(xs) - 1 - int(np.argmax(~all_pass[::-1]))
    assert any_pass.any() and first > 8 and last_fail < len(xs) - 8 and not all_pass.all()
    return float(xs[first]), float(xs[last_fail])


def sigmoid_threshold_logit(threshold):
    """Smallest float32 logit that ANY evaluation path of torch.sigmoid puts above `threshold`: the device keeps v >= this."""
    return sigmoid_threshold_band(threshold)[0]


class EventSelector:
    """Preallocated candidate buffers for one graph; `select_*` enqueue on the current stream (capturable), `fetch()`
    synchronises and returns the reference's host-side lists."""

    def __init__(self, device, edge_threshold=0.6, area_threshold=1e-4, cap_edges=4096, cap_grains=4096, on_grow=None):
        self.device = torch.device(device)
        self.on_grow = on_grow        # called when a candidate buffer is replaced: a captured step still writes the old one
        self._retired = []            # replaced buffers stay allocated: a CUDA graph captured earlier may still write into them
        self.edge_threshold, self.area_threshold = edge_threshold, area_threshold
        self.logit_min, self.logit_band_hi = sigmoid_threshold_band(edge_threshold)
        self._buf = {}
        for name, cap in (('edge', cap_edges), ('grain', cap_grains)):
            self._alloc(name, cap)
        self.d2h_bytes = 0

    def _alloc(self, name, cap):
        self._buf[name] = (torch.zeros(1, dtype=torch.int32, device=self.device),
                           torch.empty(cap, dtype=torch.int32, device=self.device),
                           torch.empty(cap, dtype=torch.float32, device=self.device), cap, None)

    def _select(self, name, values, thr, mode, src=None, dst=None, mask=None):
        if not values.is_cuda:
            raise RuntimeError('graingraphnn_b200 runs on CUDA tensors only (no CPU fallback)')
        count, ids, vals, cap, _ = self._buf[name]
        n, ld = values.shape[0], (values.stride(0) if values.dim() else 1)
        args = (values, thr, mode, src, dst, mask)
        with torch.cuda.device(self.device):
            check(_lib.lib().gg_select_events(ptr(values), n, max(ld, 1), float(thr), mode, ptr(src), ptr(dst), ptr(mask),
                                              mask.stride(0) if mask is not None else 0, cap, ptr(count), ptr(ids), ptr(vals),
                                              torch.cuda.current_stream().cuda_stream), 'gg_select_events')
        self._buf[name] = (count, ids, vals, cap, args)

    def select_edge_events(self, edge_event, jj_edge_index):
        """edge_event [E] fp32 logits (original jj edge order), jj_edge_index [2,E] int64."""
        ei = jj_edge_index
        self._select('edge', edge_event.reshape(-1), self.logit_min, 0, ei[0].contiguous(), ei[1].contiguous())

    def select_grain_events(self, grain_area, mask_grain=None):
        """grain_area [Ng] fp32, mask_grain (optional) [Ng] or [Ng,1] fp32: live grains (> 0)."""
        m = None if mask_grain is None else (mask_grain if mask_grain.dim() == 1 else mask_grain[:, 0])
        self._select('grain', grain_area.reshape(-1), self.area_threshold, 1, mask=m)

    def _fetch(self, name):
        count, ids, vals, cap, args = self._buf[name]
        n = int(count.item())
        self.d2h_bytes += 4
        if n > cap:                                   # rare: grow and repeat the pass on the same inputs
            self._retired.append(self._buf[name][:3])
            self._alloc(name, 2 * n)
            if self.on_grow is not None:              # the owner drops its captured graph (it holds the old pointers)
                self.on_grow()
            values, thr, mode, src, dst, mask = args
            self._select(name, values, thr, mode, src, dst, mask)
            return self._fetch(name)
        i, v = ids[:n].cpu(), vals[:n].cpu()
        self.d2h_bytes += 8 * n
        if name == 'edge':                            # the reference's own test on the survivors (decides the few-ulp band)
            keep = torch.sigmoid(v) > self.edge_threshold
            i, v = i[keep], v[keep]
        order = torch.argsort(i)                      # `nonzero` order (ids are unique)
        return i[order].long(), v[order]

    def fetch(self):
        """-> {'L1': int64 ids ascending (models.py:629), 'L1_logit': their logits, 'grain_event': int64 ids sorted by
        area as test.py:416 does, 'grain_event_area': areas in `nonzero` order}."""
        L1, logit = self._fetch('edge')
        g, area = self._fetch('grain')
        return {'L1': L1, 'L1_logit': logit, 'grain_event': g[torch.argsort(area)], 'grain_event_ids': g, 'grain_event_area': area}
