from .conv import HeteroConv, MessagePassing, SAGEConv  # noqa: F401
from .dense.linear import Linear  # noqa: F401
