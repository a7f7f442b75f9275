"""gymnasium subset: the two space types the hot path inspects."""
from . import spaces  # noqa: F401
from .spaces import Space  # noqa: F401
