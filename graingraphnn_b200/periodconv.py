"""`periodconv.PeriodConv` — the un-weighted variant of the periodic convolution (reference periodconv.py, identical to
periodGATconv.py except that the attention weights are computed but not applied, :235).  Same kernels, `weighted=0`."""
from .periodGATconv import PeriodConv as _AttentionPeriodConv


class PeriodConv(_AttentionPeriodConv):
    weighted = False
