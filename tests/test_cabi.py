"""CPU: the C-ABI shared library builds, loads and exports every symbol include/graingnn_b200.h declares."""
import ctypes
import os
import re

from graingraphnn_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'graingnn_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(gg_[a-z0-9_]+)\s*\(', src)))


def test_library_builds_and_exports_declared_symbols():
    path = build.build()
    assert os.path.exists(path)
    h = ctypes.CDLL(path)
    names = declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(h, n), f'{n} declared in include/graingnn_b200.h but not exported'


def test_python_binding_covers_the_header():
    bound = set(_lib._PROTOS) | set(_lib._OPTIONAL)
    assert set(declared_symbols()) <= bound


def test_error_strings_and_argument_checks_without_gpu():
    L = _lib.lib()
    assert L.gg_version() >= 100
    assert b'invalid argument' in L.gg_error_string(-1)
    assert b'aligned' in L.gg_error_string(-3)
    # argument validation happens before any CUDA call, so it is testable on a CPU-only box
    assert L.gg_csr_build(None, -1, 0, 0, None, None, None, None, None, 0, None) == -1
    assert L.gg_pgat_gather(None, 0, 0, 0, None, 0, 0, 0, None, 0, None, 0, None, None, None, None, None, None, 0, None, 5, 4, 100, 1, None, 0, None, None) == -1
    assert L.gg_gather_dcap() == 3 and L.gg_csr_items(None, -1, 3, None, None, None, 0, None) == -1
    assert L.gg_edge_wrap(None, 0, None, 0, None, None, 5, None, None) == -1
    assert L.gg_node_proj(None, 0, 0, None, 0, 0, None, 0, None, None, 0, 0, 0, None) == 0   # empty problem is a no-op


def test_tiled_gather_geometry_and_argument_checks_without_gpu():
    """Tile sizes of the warp-specialised gather per (gates, width, row form) and host-side validation of its entry points."""
    L = _lib.lib()
    assert L.gg_gather_tile_ecap(3, 96, 16) == 54          # encoder: [16 raw | V] rows, three stages of 54 edges
    assert L.gg_gather_tile_ecap(4, 96, 128) == 24         # decoder: [input (32 + C) | V] rows
    assert L.gg_gather_tile_ecap(4, 96, 0) == 18           # decoder, K | V rows
    assert L.gg_gather_tile_ecap(1, 96, 0) > 0 and L.gg_gather_tile_ecap(2, 96, 0) == 0 and L.gg_gather_tile_ecap(4, 96, 64) == 0
    assert L.gg_csr_tiles_capacity(1000, 24, 148) >= (1000 + 23) // 24
    assert L.gg_csr_tiles_scratch_ints(1000, 24, 148) > 0
    assert L.gg_csr_tiles(None, None, -1, 24, 148, None, None, None, None) == -1
    assert L.gg_csr_compact(None, -1, None, None, None, None, None, 0, None) == -1
    # wrong tile size for the shape -> invalid argument, before any CUDA call
    assert L.gg_pgat_gather_tiled(None, 0, 0, None, 0, 0, None, None, None, None, None, None, None, None, 148, 99, 10, 16, None,
                                  5, 3, 96, None, 0, None, None) == -1
    assert L.gg_node_proj_fused(None, 8, 3, None, 0, 0, None, None, 32, None, None, 32, 10, 0, None) == -1   # K1 % 4 != 0
    assert L.gg_edge_refresh(None, 0, None, 0, None, None, None, 5, None, None, None, None) == -1
