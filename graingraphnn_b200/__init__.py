"""graingraphnn_b200 — B200-native (sm_100a) implementation of the GrainGNN rollout message-passing path.

Boundary modules (same names / constructors / forward signatures / state_dict layout as YigongQin/GrainGraphNN):
    periodGATconv.PeriodConv, periodconv.PeriodConv, heteropgclstm.{HeteroPGCLSTM, HeteroPGC},
    heterogclstm.{HeteroGCLSTM, HeteroGC}, models.{SeqGCLSTM, GrainNN_regressor, GrainNN_classifier}
Everything numeric runs in libgraingnn_b200.so (include/graingnn_b200.h); there is no CPU fallback.
"""
from . import _lib  # noqa: F401
from ._lib import lib, lib_path  # noqa: F401

__version__ = '0.1.0'
