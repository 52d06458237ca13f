from tensorflow._core import HParams  # noqa: F401
