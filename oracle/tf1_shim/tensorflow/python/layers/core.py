from tensorflow._core import Dense, dense, dropout  # noqa: F401
