from tensorflow._core import (_bahdanau_score, _BaseAttentionMechanism, BahdanauAttention, BahdanauMonotonicAttention,  # noqa: F401
                        AttentionWrapperState, AttentionMechanism, monotonic_attention, safe_cumprod)


class AttentionWrapper:  # models/modules.py:6 imports the name and never uses it
    def __init__(self, *a, **k):
        raise NotImplementedError("the reference ships its own AttentionWrapper (models/rnn_wrappers.py)")
