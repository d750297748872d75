from .conv import GCNConv, MessagePassing  # noqa: F401
