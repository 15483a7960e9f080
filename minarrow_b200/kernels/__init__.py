"""Mirror of the reference's `kernels` module tree for the hot path:
arithmetic (leaf kernels), bitmask, routing, broadcast, plus the null-aware reductions."""
from . import arithmetic, bitmask, broadcast, reduce, routing  # noqa: F401
