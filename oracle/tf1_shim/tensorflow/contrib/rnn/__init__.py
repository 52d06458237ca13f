from tensorflow._core import RNNCell, GRUCell, MultiRNNCell, OutputProjectionWrapper, ResidualWrapper  # noqa: F401
