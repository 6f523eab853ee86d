"""dd_b200: host-side binding of libdynamo_b200.so (hand-written sm_100a kernels for Dynamo-Depth's hot path)."""
from . import _lib  # noqa: F401
from ._lib import DynamoB200Error, load  # noqa: F401
