from tensorflow._core import _zero_state_tensors, _linear, RNNCell, GRUCell, MultiRNNCell, ResidualWrapper  # noqa: F401
