"""ctypes binding of libgraingnn_b200.so (the C ABI declared in include/graingnn_b200.h).

There is no CPU fallback: if the library is missing, `lib()` raises; if a kernel rejects its arguments or the
launch fails, `check()` raises RuntimeError with the library's message.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int32, c_int64, c_size_t, c_void_p

from . import build as _build

_LIB = None

GG_GATE_RAW, GG_GATE_RELU, GG_GATE_LSTM, GG_GATE_LSTM0 = 0, 1, 2, 3


class AggInput(Structure):
    _fields_ = [('agg', c_void_p), ('ld_agg', c_int32), ('ea', c_void_p), ('rowptr', c_void_p),
                ('W2', c_void_p), ('We', c_void_p), ('b2', c_void_p), ('weighted', c_int32)]


class GatherSegment(Structure):
    """gg_gather_segment (include/graingnn_b200.h): one edge type of a multi-segment tiled gather launch."""
    _fields_ = [('P_src', c_void_p), ('ld_src', c_int32), ('k_off', c_int32),
                ('P_dst', c_void_p), ('ld_dst', c_int32), ('q_off', c_int32),
                ('rowptr', c_void_p), ('col', c_void_p), ('eattr_csr', c_void_p), ('wrap_csr', c_void_p),
                ('nz', c_void_p), ('nzptr', c_void_p), ('tiles', c_void_p), ('cta_ptr', c_void_p),
                ('n_edges', c_int64), ('Wv3', c_void_p), ('n_dst', c_int32),
                ('agg', c_void_p), ('ld_agg', c_int32), ('ea', c_void_p)]


_P, _I, _L, _F, _S = c_void_p, c_int32, c_int64, c_float, c_size_t
_PROTOS = {
    'gg_error_string': (c_char_p, [_I]),
    'gg_version': (_I, []),
    'gg_device_is_sm100': (_I, []),
    'gg_csr_workspace_bytes': (_S, [_L, _I]),
    'gg_csr_build': (_I, [_P, _L, _I, _I, _P, _P, _P, _P, _P, _S, _P]),
    'gg_permute_f32': (_I, [_P, _P, _P, _L, _P]),
    'gg_edge_length': (_I, [_P, _I, _P, _I, _P, _L, _P, _P, _P, _P]),
    'gg_node_proj': (_I, [_P, _I, _I, _P, _I, _I, _P, _I, _P, _P, _I, _I, _I, _P]),
    'gg_pgat_gather': (_I, [_P, _I, _I, _I, _P, _I, _I, _I, _P, _I, _P, _I, _P, _P, _P, _P, _P, _P, _I, _P, _I, _I, _I, _I, _P, _I, _P, _P]),
    'gg_gather_dcap': (_I, []),
    'gg_csr_items': (_I, [_P, _I, _I, _P, _P, _P, _S, _P]),
    'gg_gather_tile_ecap': (_I, [_I, _I, _I]),
    'gg_csr_compact': (_I, [_P, _I, _P, _P, _P, _P, _P, _S, _P]),
    'gg_gather_ctas': (_I, []),
    'gg_csr_tiles_capacity': (_L, [_L, _I, _I]),
    'gg_csr_tiles_scratch_ints': (_S, [_L, _I, _I]),
    'gg_csr_tiles': (_I, [_P, _P, _L, _I, _I, _P, _P, _P, _P]),
    'gg_pgat_gather_tiled': (_I, [_P, _I, _I, _P, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _L, _I, _P, _I, _I, _I, _P, _I, _P, _P]),
    'gg_pgat_gather_tiled_multi': (_I, [POINTER(GatherSegment), _I, _I, _I, _I, _I, _I, _P]),
    'gg_edge_wrap': (_I, [_P, _I, _P, _I, _P, _P, _I, _P, _P]),
    'gg_edge_refresh': (_I, [_P, _I, _P, _I, _P, _P, _P, _I, _P, _P, _P, _P]),
    'gg_gate_update': (_I, [POINTER(AggInput), _I, _P, _I, _I, _P, _I, _P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    'gg_node_head': (_I, [_P, _I, _I, _P, _P, _I, POINTER(c_int32), _P, _I, _P, _I, _F, _P, _I, _P]),
    'gg_edge_head': (_I, [_P, _I, _I, _P, _L, _P, _P, _P, _P, _P, _P, _P, _P]),
    'gg_feature_update': (_I, [_P, _I, _I, _P, _P, _I, _I, _I, _P, _F, _F, _P, _P]),
    'gg_feature_update_batched': (_I, [_P, _I, _I, _P, _P, _I, _I, _I, _P, _P, _P, _F, _P]),
    'gg_segment_mean': (_I, [_P, _I, _I, _P, _P, _I, _P, _I, _P]),
    'gg_select_events': (_I, [_P, _L, _I, _F, _I, _P, _P, _P, _I, _I, _P, _P, _P, _P]),
    'gg_joint_rank': (_I, [_P, _L, _I, _P, _P]),
    'gg_region_key': (_I, [_P, _P, _L, _P, _P]),
    'gg_region_sort': (_I, [_P, _P, _P, _I, _P, _P]),
    'gg_region_center': (_I, [_P, _I, _P, _F, _P, _P, _P, _I, _P, _P, _I, _P]),
    'gg_area_bookkeeping': (_I, [_P, _I, _P, _I, _I, c_double, c_double, c_double, _P, _P, _P, _P, _P, _I, c_double, _P, _P]),
    'gg_topology_caps': (_I, [_P, _P]),
    'gg_topology_work_ints': (_L, [_L, _L, _L]),
    'gg_topology_lists': (_I, [_P, _L, _L, _P, _P, _I, _L, _P, _P, _I, _L, _P, _P]),
    'gg_topology_update': (_I, [_P, _L, _L, _P, _L, _L, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _I, _P, _P, _I, _P, _P, _P, _P, _I, _I,
                                _P, _P, _P, _I, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    'gg_raster_polygons': (_I, [_P, _P, _P, _I, _I, _P, _P, _P]),
    'gg_count_mismatch': (_I, [_P, _P, _L, _P, _P]),
    'gg_gather_rows': (_I, [_P, _I, _P, _I, _I, _P, _I, _P]),
    'gg_scatter_rows': (_I, [_P, _I, _P, _I, _I, _P, _I, _P]),
}
# entry points of the tcgen05 path (gemm_tc.cu); bound when present
_OPTIONAL = {
    'gg_tc_supported': (_I, []),
    'gg_node_proj_tc': (_I, [_P, _P, _I, _I, _P, _P, _I, _P, _P, _I, _I, _I, _P]),
    'gg_split_tf32': (_I, [_P, _I, _I, _P, _I, _I, _I, _P, _P, _I, _I, _P]),
    'gg_node_proj_fused': (_I, [_P, _I, _I, _P, _I, _I, _P, _P, _I, _P, _P, _I, _I, _I, _P]),
    'gg_gate_update_tc': (_I, [POINTER(AggInput), _I, _P, _I, _I, _P, _I, _P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
}


def lib_path():
    return _build.lib_path()


def lib():
    """Load (once) and return the ctypes handle. Raises if the shared library has not been built."""
    global _LIB
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            raise RuntimeError(
                f'{path} not found: build it with `python -m graingraphnn_b200.build` '
                '(or __graft_entry__.build()); graingraphnn_b200 has no CPU fallback')
        h = ctypes.CDLL(path)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(h, name)
            fn.restype, fn.argtypes = res, args
        for name, (res, args) in _OPTIONAL.items():
            if hasattr(h, name):
                fn = getattr(h, name)
                fn.restype, fn.argtypes = res, args
        _LIB = h
    return _LIB


def exported_symbols():
    return list(_PROTOS)


# kernels launched by one successful call of each entry point (for bench.py's gpu_launches accounting)
KERNELS_PER_CALL = {'gg_csr_build': 6, 'gg_csr_items': 5, 'gg_csr_compact': 5, 'gg_csr_tiles': 5, 'gg_pgat_gather_tiled': 1, 'gg_pgat_gather_tiled_multi': 1, 'gg_edge_wrap': 1, 'gg_edge_refresh': 1, 'gg_permute_f32': 1, 'gg_edge_length': 2, 'gg_node_proj': 1, 'gg_pgat_gather': 1,
                    'gg_gate_update': 1, 'gg_node_head': 1, 'gg_edge_head': 1, 'gg_feature_update': 3, 'gg_feature_update_batched': 1,
                    'gg_gather_rows': 1, 'gg_scatter_rows': 1, 'gg_select_events': 1, 'gg_joint_rank': 2, 'gg_region_key': 1, 'gg_region_sort': 1, 'gg_region_center': 1, 'gg_area_bookkeeping': 2, 'gg_topology_lists': 6, 'gg_topology_update': 5, 'gg_raster_polygons': 2, 'gg_count_mismatch': 1, 'gg_segment_mean': 1, 'gg_node_proj_tc': 1, 'gg_node_proj_fused': 1, 'gg_split_tf32': 1, 'gg_gate_update_tc': 1}
LAUNCHES = [0]


def check(rc, what=''):
    LAUNCHES[0] += KERNELS_PER_CALL.get(what, 0)
    if rc != 0:
        msg = lib().gg_error_string(int(rc)).decode()
        raise RuntimeError(f'{what or "graingnn_b200"} failed ({rc}): {msg}')


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else c_void_p(t.data_ptr())
