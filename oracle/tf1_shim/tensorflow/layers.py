from ._core import dense, conv1d, batch_normalization, max_pooling1d, dropout, Dense, Conv1D, BatchNormalization, Layer  # noqa: F401
