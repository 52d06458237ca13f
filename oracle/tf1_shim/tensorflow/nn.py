from ._core import relu, softsign, sigmoid, tanh, softmax, embedding_lookup, bidirectional_dynamic_rnn, dynamic_rnn  # noqa: F401
