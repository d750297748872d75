from .message_passing import MessagePassing  # noqa: F401
from .gcn_conv import GCNConv  # noqa: F401
