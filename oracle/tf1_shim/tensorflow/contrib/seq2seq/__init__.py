from tensorflow._core import (Helper, BasicDecoder, BasicDecoderOutput, dynamic_decode, BahdanauAttention, BahdanauMonotonicAttention,  # noqa: F401
                      AttentionWrapperState, AttentionMechanism, tile_batch, monotonic_attention, safe_cumprod)
