"""state_dict keys and shapes of the REFERENCE's own GrainNN_regressor / GrainNN_classifier (models.py imported unmodified
from /root/reference on oracle/pyg_stub, hyper-parameters from parameters.py as test.py:162-173 sets them up) ->
tests/golden/reference_state_dict_layout.json.  Run once in the build container.  TEST INFRASTRUCTURE ONLY."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402


def main():
    g, _, _, _ = mg.load_graph('/root/reference/graphs/40_40/seed10020_G1.904_R0.558_span6.pkl', 1)
    R, C = mg.build_models(g)
    out = {}
    for name, m in (('regressor', R), ('classifier', C)):
        out[name] = {'keys': [[k, list(v.shape)] for k, v in m.state_dict().items()],
                     'n_params': sum(p.numel() for p in m.parameters())}
    path = os.path.join(HERE, '..', 'tests', 'golden', 'reference_state_dict_layout.json')
    with open(path, 'w') as f:
        json.dump(out, f)
    print(path, {k: (len(v['keys']), v['n_params']) for k, v in out.items()})


if __name__ == '__main__':
    main()
