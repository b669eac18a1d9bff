from .message_passing import MessagePassing  # noqa: F401
from .sage_conv import SAGEConv  # noqa: F401
from .hetero_conv import HeteroConv  # noqa: F401
