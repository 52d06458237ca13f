from ._core import AdamOptimizer, exponential_decay  # noqa: F401
