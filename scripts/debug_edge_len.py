import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')]
import torch, numpy as np
import grain_oracle as orc
from util import ET, load_graph
from graingraphnn_b200.graph import build_csr, edge_length
x, ei, ea = load_graph('c1')
ref = orc.edge_attr_rebuild(x, ei)
d = torch.device('cuda:0')
for e in ET:
    out, _ = edge_length(x[e[0]].to(d), x[e[2]].to(d), ei[e].to(d), None)
    o = out.cpu()
    diff = (o - ref[e]).abs()
    bad = (diff > 0).nonzero()[:, 0]
    print(e, 'n_bad', bad.numel(), 'max diff', diff.max().item())
    for b in bad[:5].tolist():
        s, t = ei[e][0, b].item(), ei[e][1, b].item()
        print('  edge', b, o[b].item(), ref[e][b].item(), np.float32(o[b].item()).view(np.uint32) - np.float32(ref[e][b].item()).view(np.uint32),
              x[e[0]][s, :2].tolist(), x[e[2]][t, :2].tolist())
