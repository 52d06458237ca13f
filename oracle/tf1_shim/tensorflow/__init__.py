"""Eager stand-in for `import tensorflow as tf` (TF r1.4 surface used by /root/reference/models) — see _core.py."""
from ._core import (Tensor, Variable, TensorShape, Dimension, TensorArray, GraphKeys, DType,  # noqa: F401
                    float32, float64, int32, int64, bool_ as bool,
                    variable_scope, name_scope, control_dependencies, get_variable, get_variable_scope, get_collection,
                    add_to_collection, trainable_variables, global_variables, placeholder, reset_default_graph,
                    truncated_normal_initializer, constant_initializer, zeros_initializer, ones_initializer, glorot_uniform_initializer,
                    shape, tile, concat, expand_dims, squeeze, reshape, transpose, split, identity, zeros, ones, zeros_like, fill,
                    cast, one_hot, reduce_mean, reduce_sum, reduce_all, equal, logical_or, logical_and, logical_not, abs, minimum,
                    maximum, matmul, tanh, sigmoid, sqrt, rsqrt, square, exp, log, cumsum, clip_by_value, where, reverse_sequence,
                    reverse, cond, assert_equal, clip_by_global_norm, set_float_precision, shim_state)
from . import nn, layers, train, contrib  # noqa: F401,E402

__version__ = "1.4.0-shim"
