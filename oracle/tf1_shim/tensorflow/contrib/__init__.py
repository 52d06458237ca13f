from . import rnn, seq2seq, training  # noqa: F401
