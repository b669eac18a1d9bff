"""Import shims that let the read-only reference tree (/root/reference) be imported in the
build container, where matplotlib / h5py / termcolor / tvtk / torch_geometric are absent.

TEST INFRASTRUCTURE ONLY.  Used by oracle/make_golden.py (fixture generation, run once in the
container that has /root/reference) and by tests that are skipped when /root/reference is absent.
Nothing in the product path (graingraphnn_b200/) imports this file.
"""
import sys
import types

import numpy as np


def _mod(name):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        sys.modules[name] = m
    return m


def install_plot_stubs():
    """matplotlib/h5py/termcolor stand-ins: enough surface for graph_datastruct.py:13-21 and
    graph_trajectory.py:10-15 to import; none of the plotting entry points are ever called."""
    if 'matplotlib' in sys.modules and getattr(sys.modules['matplotlib'], '__gg_stub__', False):
        return
    try:
        import matplotlib  # noqa: F401  (a real install wins)
        return
    except ImportError:
        pass
    mpl = _mod('matplotlib')
    mpl.__gg_stub__ = True
    plt = _mod('matplotlib.pyplot')
    plt.rcParams = {}
    cm = _mod('matplotlib.cm')
    colors = _mod('matplotlib.colors')

    def get_cmap(name, n=256):
        def cmap(x):
            x = np.asarray(x, dtype=float)
            return np.stack([x, x, x, np.ones_like(x)], axis=-1)
        return cmap

    class ListedColormap:  # noqa: D401
        def __init__(self, colors_, *a, **k):
            self.colors = colors_

    cm.get_cmap = get_cmap
    colors.ListedColormap = ListedColormap
    mpl.pyplot, mpl.cm, mpl.colors = plt, cm, colors
    for name in ('matplotlib.patches', 'matplotlib.collections', 'mpl_toolkits',
                 'mpl_toolkits.axes_grid1', 'h5py'):
        _mod(name)
    tc = _mod('termcolor')
    tc.colored = lambda s, *a, **k: s


def add_reference_to_path(ref='/root/reference'):
    if ref not in sys.path:
        sys.path.insert(0, ref)
