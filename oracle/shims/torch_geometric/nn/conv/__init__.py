from .message_passing import MessagePassing  # noqa: F401
from .transformer_conv import TransformerConv  # noqa: F401
