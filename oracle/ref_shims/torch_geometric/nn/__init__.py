from .conv import MessagePassing, GraphConv  # noqa: F401
