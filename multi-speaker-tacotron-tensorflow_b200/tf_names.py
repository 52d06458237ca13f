"""TensorFlow variable names of the reference graph <-> this repo's parameter names.

The reference builds its graph under ``tf.variable_scope('model')`` (train.py:145, synthesizer.py:49) and
``'inference'`` (models/tacotron.py:28); every variable below that is named by TF's layer / RNN-cell scoping rules
(r1.4): functional ``tf.layers.*`` without a name take ``dense``, ``dense_1`` ... per enclosing scope, RNN cells take the
snake-cased class name, ``MultiRNNCell`` adds ``cell_<i>``, ``dynamic_decode`` adds ``decoder``.  The table produced
here is what ``tools/make_reference_golden.py`` uses to load OUR parameters into the reference's model code, and what a
``model.ckpt`` importer needs.  The names were obtained by executing the reference's model-building code
(see oracle/tf1_shim); a real TF checkpoint could not be inspected in this environment.

Shape differences: ``attention_score_bias`` / ``attention_g`` are TF scalars and 1-element vectors here;
``moving_variance`` is ``moving_var`` here.
"""
from __future__ import annotations

from typing import Dict

_DEC = "decoder/output_projection_wrapper"
_CELL0 = _DEC + "/multi_rnn_cell/cell_0/output_projection_wrapper"
_ATT = _CELL0 + "/concat_output_and_attention_wrapper/attention_wrapper"
_PRE = _ATT + "/decoder_prenet_wrapper"

_GRU = {"gates/kernel": "gates_kernel", "gates/bias": "gates_bias", "candidate/kernel": "cand_kernel", "candidate/bias": "cand_bias"}
_CONV = {"conv1d/kernel": "kernel", "conv1d/bias": "bias", "batch_normalization/gamma": "gamma", "batch_normalization/beta": "beta",
         "batch_normalization/moving_mean": "moving_mean", "batch_normalization/moving_variance": "moving_var"}


def _cbhg(tf_scope: str, ours: str, bank_size: int, depth: int, has_dense: bool) -> Dict[str, str]:
    m: Dict[str, str] = {}
    for k in range(1, bank_size + 1):
        for a, b in _CONV.items():
            m["%s/conv_bank/conv1d_%d/%s" % (tf_scope, k, a)] = "%s/bank_%d/%s" % (ours, k, b)
    for i in (1, 2):
        for a, b in _CONV.items():
            m["%s/proj_%d/%s" % (tf_scope, i, a)] = "%s/proj_%d/%s" % (ours, i, b)
    if has_dense:                                                       # modules.py:72-73
        m[tf_scope + "/dense/kernel"] = ours + "/highway_in/kernel"
        m[tf_scope + "/dense/bias"] = ours + "/highway_in/bias"
    for i in range(1, depth + 1):
        for g in "HT":
            for p in ("kernel", "bias"):
                m["%s/highway_%d/%s/%s" % (tf_scope, i, g, p)] = "%s/highway_%d/%s_%s" % (ours, i, g, p)
    for d in ("fw", "bw"):
        for a, b in _GRU.items():
            m["%s/bidirectional_rnn/%s/gru_cell/%s" % (tf_scope, d, a)] = "%s/gru_%s/%s" % (ours, d, b)
    return m


def tf_to_ours(hp, num_speakers: int = 1, prefix: str = "model/inference/") -> Dict[str, str]:
    """{TF variable name: name in params.param_specs(hp, num_speakers)}."""
    from .params import speaker_mode
    mode = speaker_mode(hp, num_speakers)
    m: Dict[str, str] = {"embedding": "embedding"}
    n_dense = 0                                                          # unnamed tf.layers.dense calls directly under 'inference'
    sites = ["before_highway", "encoder_rnn_init_state", "attention_rnn_init_state"] + \
            ["decoder_rnn_init_states%d" % (i + 1) for i in range(hp.dec_layer_num)]
    if mode in ("simple", "deepvoice"):
        m["speaker_embedding"] = "speaker_embedding"
    if mode == "deepvoice":                                              # tacotron.py:68-79, in call order
        for s in sites:
            tf_dense = "dense" if n_dense == 0 else "dense_%d" % n_dense
            m[tf_dense + "/kernel"] = "speaker/%s/kernel" % s
            m[tf_dense + "/bias"] = "speaker/%s/bias" % s
            n_dense += 1
    elif mode == "deepvoice_table":                                      # modules.py:11-15
        for s in sites:
            m[s] = "speaker/%s/table" % s
    for i in range(1, len(hp.enc_prenet_sizes) + 1):
        for p in ("kernel", "bias"):
            m["prenet/dense_%d/%s" % (i, p)] = "enc_prenet/dense_%d/%s" % (i, p)
    m.update(_cbhg("encoder_cbhg", "enc_cbhg", hp.enc_bank_size, hp.enc_highway_depth, hp.enc_proj_sizes[-1] != hp.enc_rnn_size))
    m["memory_layer/kernel"] = "attention/memory_kernel"
    for i in range(1, len(hp.dec_prenet_sizes) + 1):
        for p in ("kernel", "bias"):
            m["%s/decoder_prenet/dense_%d/%s" % (_PRE, i, p)] = "dec_prenet/dense_%d/%s" % (i, p)
    for a, b in _GRU.items():
        m["%s/gru_cell/%s" % (_PRE, a)] = "attention_gru/" + b
    att = {"bah_mon": "bahdanau_monotonic_attention", "bah_norm": "bahdanau_attention", "bah": "bahdanau_attention"}[hp.attention_type]
    m["%s/%s/query_layer/kernel" % (_ATT, att)] = "attention/query_kernel"
    m["%s/%s/attention_v" % (_ATT, att)] = "attention/v"
    if hp.attention_type == "bah_mon":
        m["%s/%s/attention_score_bias" % (_ATT, att)] = "attention/score_bias"
    elif hp.attention_type == "bah_norm":
        m["%s/%s/attention_g" % (_ATT, att)] = "attention/g"
        m["%s/%s/attention_b" % (_ATT, att)] = "attention/b"
    m[_CELL0 + "/kernel"] = "concat_proj/kernel"
    m[_CELL0 + "/bias"] = "concat_proj/bias"
    for i in range(1, hp.dec_layer_num + 1):
        for a, b in _GRU.items():
            m["%s/multi_rnn_cell/cell_%d/gru_cell/%s" % (_DEC, i, a)] = "dec_gru_%d/%s" % (i, b)
    m[_DEC + "/kernel"] = "mel_proj/kernel"
    m[_DEC + "/bias"] = "mel_proj/bias"
    m.update(_cbhg("post_cbhg", "post_cbhg", hp.post_bank_size, hp.post_highway_depth, hp.post_proj_sizes[-1] != hp.post_rnn_size))
    tf_dense = "dense" if n_dense == 0 else "dense_%d" % n_dense       # tacotron.py:235
    m[tf_dense + "/kernel"] = "linear/kernel"
    m[tf_dense + "/bias"] = "linear/bias"
    return {prefix + k: v for k, v in m.items()}
