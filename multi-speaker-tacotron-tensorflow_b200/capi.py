"""ctypes binding of ``libtaco_b200.so`` (the C ABI declared in ``include/taco_capi.h``).

The library is the product: there is no CPU or PyTorch fallback.  If the shared object is missing it is built in-tree
with nvcc (``build.py``); if that is impossible, or a GPU call is made without a CUDA device, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional, Sequence, Tuple

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libtaco_b200.so")

TACO_ABI_VERSION = 4
TACO_SCALARS_RAW_BYTES = 128
ATT_TYPES = {"bah_mon": 0, "bah": 1, "bah_norm": 2}
SPK_MODES = {"none": 0, "simple": 1, "deepvoice": 2, "deepvoice_table": 3}
PREC = {"fp32": 0, "tf32": 1, "bf16": 2}


class TacoConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("num_symbols", C.c_int32), ("embedding_size", C.c_int32),
        ("num_speakers", C.c_int32), ("speaker_mode", C.c_int32), ("speaker_embedding_size", C.c_int32),
        ("enc_prenet_sizes", C.c_int32 * 2), ("enc_bank_size", C.c_int32), ("enc_bank_channels", C.c_int32),
        ("enc_proj_sizes", C.c_int32 * 2), ("enc_proj_width", C.c_int32), ("enc_highway_depth", C.c_int32),
        ("enc_rnn_size", C.c_int32), ("attention_type", C.c_int32), ("attention_size", C.c_int32),
        ("attention_state_size", C.c_int32), ("dec_prenet_sizes", C.c_int32 * 2), ("dec_layer_num", C.c_int32),
        ("dec_rnn_size", C.c_int32), ("post_bank_size", C.c_int32), ("post_bank_channels", C.c_int32),
        ("post_proj_sizes", C.c_int32 * 2), ("post_proj_width", C.c_int32), ("post_highway_depth", C.c_int32),
        ("post_rnn_size", C.c_int32), ("num_mels", C.c_int32), ("num_freq", C.c_int32),
        ("reduction_factor", C.c_int32), ("precision", C.c_int32), ("device", C.c_int32),
        ("prioritize_loss", C.c_int32), ("priority_lo", C.c_int32), ("priority_hi", C.c_int32),
    ]


class TacoParamEntry(C.Structure):
    _fields_ = [("name", C.c_char_p), ("offset", C.c_int64), ("numel", C.c_int64), ("trainable", C.c_int32)]


class TacoBatch(C.Structure):
    _fields_ = [
        ("N", C.c_int32), ("T_in", C.c_int32), ("T_out", C.c_int32),
        ("inputs", C.c_void_p), ("input_lengths", C.c_void_p), ("speaker_id", C.c_void_p),
        ("mel_targets", C.c_void_p), ("linear_targets", C.c_void_p), ("loss_coeff", C.c_void_p),
        ("manual_alignments", C.c_void_p), ("decoder_steps", C.c_int32), ("rnn_decoder_test_mode", C.c_int32),
        ("linear_targets_bf16", C.c_int32),
    ]


class TacoStepScalars(C.Structure):
    _fields_ = [("loss", C.c_float), ("mel_loss", C.c_float), ("linear_loss", C.c_float),
                ("loss_without_coeff", C.c_float), ("grad_norm", C.c_float), ("learning_rate", C.c_float)]


class TacoGemmDesc(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("B", C.c_void_p), ("C", C.c_void_p),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("lda", C.c_int32), ("ldb", C.c_int32), ("ldc", C.c_int32),
        ("transA", C.c_int32), ("ctap", C.c_int32), ("transB", C.c_int32),
        ("alpha", C.c_float), ("accumulate", C.c_int32),
        ("bias", C.c_void_p), ("act", C.c_int32),
        ("mask_period", C.c_int32), ("mask_lo", C.c_int32), ("mask_hi", C.c_int32),
        ("remap_period", C.c_int32), ("remap_outer", C.c_int64), ("remap_inner", C.c_int64),
        ("colsum", C.c_void_p), ("colsumsq", C.c_void_p), ("split_k", C.c_int32),
        ("tap_table", C.c_void_p), ("tap_rows", C.c_int32),
        ("A16", C.c_void_p), ("B16", C.c_void_p), ("C16", C.c_void_p), ("ldc16", C.c_int32),
    ]


# every symbol include/taco_capi.h declares, with its signature
_SIGNATURES = {
    "taco_last_error": (C.c_char_p, []),
    "taco_abi_version": (C.c_int, []),
    "taco_launch_count": (C.c_int64, []),
    "taco_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(TacoConfig)]),
    "taco_destroy": (C.c_int, [C.c_void_p]),
    "taco_bind_params": (C.c_int, [C.c_void_p, C.POINTER(TacoParamEntry), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]),
    "taco_workspace_bytes": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_size_t)]),
    "taco_bind_workspace": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "taco_ws_region": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_size_t), C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                 C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "taco_forward": (C.c_int, [C.c_void_p, C.POINTER(TacoBatch), C.c_void_p]),
    "taco_backward": (C.c_int, [C.c_void_p, C.POINTER(TacoBatch), C.c_void_p]),
    "taco_optimizer_step": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_float, C.c_int32, C.c_float, C.c_float, C.c_float,
                                      C.c_void_p]),
    "taco_cbhg_forward": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "taco_cbhg_backward": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "taco_decoder_forward": (C.c_int, [C.c_void_p, C.POINTER(TacoBatch), C.c_void_p, C.c_void_p]),
    "taco_decoder_backward": (C.c_int, [C.c_void_p, C.POINTER(TacoBatch), C.c_void_p, C.c_void_p, C.c_void_p]),
    "taco_highway_combine": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "taco_batch_norm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "taco_dp_bucket": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "taco_dp_wait_bucket": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "taco_read_scalars": (C.c_int, [C.c_void_p, C.POINTER(TacoStepScalars), C.c_void_p]),
    "taco_copy_scalars_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "taco_finish_scalars": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(TacoStepScalars)]),
    "taco_gl_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "taco_gl_destroy": (C.c_int, [C.c_void_p]),
    "taco_gl_workspace_bytes": (C.c_size_t, [C.c_void_p]),
    "taco_gl_inv_spectrogram": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_float,
                                          C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "taco_audio_spectrogram": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_int32,
                                         C.c_void_p, C.c_void_p, C.c_void_p]),
    "taco_crc32c": (C.c_uint32, [C.c_void_p, C.c_size_t, C.c_uint32]),
    "taco_profile": (C.c_int, [C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "taco_gemm": (C.c_int, [C.POINTER(TacoGemmDesc), C.c_int32, C.c_int32, C.c_void_p]),
}
DECLARED_SYMBOLS = tuple(_SIGNATURES)

_lib: Optional[C.CDLL] = None


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load (building first if needed) the shared library and attach the signatures."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise RuntimeError("libtaco_b200.so is missing and building was disabled")
        from . import build as _build
        _build.build(verbose=False)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError here = the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.taco_abi_version() != TACO_ABI_VERSION:
        raise RuntimeError("libtaco_b200.so ABI %d != binding ABI %d" % (lib.taco_abi_version(), TACO_ABI_VERSION))
    _lib = lib
    return lib


class TacoError(RuntimeError):
    pass


def check(rc: int) -> None:
    if rc != 0:
        msg = load().taco_last_error()
        raise TacoError("libtaco_b200 error %d: %s" % (rc, msg.decode(errors="replace") if msg else "?"))


def make_config(hp, num_speakers: int, speaker_mode: str, precision: str, device: int, num_symbols: int) -> TacoConfig:
    cfg = TacoConfig()
    cfg.abi_version = TACO_ABI_VERSION
    cfg.num_symbols = num_symbols
    cfg.embedding_size = hp.embedding_size
    cfg.num_speakers = num_speakers
    cfg.speaker_mode = SPK_MODES[speaker_mode]
    cfg.speaker_embedding_size = hp.speaker_embedding_size
    if len(hp.enc_prenet_sizes) != 2 or len(hp.dec_prenet_sizes) != 2 or len(hp.enc_proj_sizes) != 2 or len(hp.post_proj_sizes) != 2:
        raise TacoError("the kernels are built for two prenet layers and two CBHG projections")
    if hp.dec_layer_num != 2:
        raise TacoError("the decoder sequencing is built for dec_layer_num == 2")
    cfg.enc_prenet_sizes[:] = hp.enc_prenet_sizes
    cfg.enc_bank_size, cfg.enc_bank_channels = hp.enc_bank_size, hp.enc_bank_channel_size
    cfg.enc_proj_sizes[:] = hp.enc_proj_sizes
    cfg.enc_proj_width, cfg.enc_highway_depth, cfg.enc_rnn_size = hp.enc_proj_width, hp.enc_highway_depth, hp.enc_rnn_size
    if hp.attention_type not in ATT_TYPES:
        raise TacoError(" [!] Unkown attention type: {}".format(hp.attention_type))
    cfg.attention_type = ATT_TYPES[hp.attention_type]
    cfg.attention_size, cfg.attention_state_size = hp.attention_size, hp.attention_state_size
    cfg.dec_prenet_sizes[:] = hp.dec_prenet_sizes
    cfg.dec_layer_num, cfg.dec_rnn_size = hp.dec_layer_num, hp.dec_rnn_size
    cfg.post_bank_size, cfg.post_bank_channels = hp.post_bank_size, hp.post_bank_channel_size
    cfg.post_proj_sizes[:] = hp.post_proj_sizes
    cfg.post_proj_width, cfg.post_highway_depth, cfg.post_rnn_size = hp.post_proj_width, hp.post_highway_depth, hp.post_rnn_size
    if hp.enc_maxpool_width != 2 or hp.post_maxpool_width != 2:
        raise TacoError("max-pool width 2 is the only width built")
    cfg.num_mels, cfg.num_freq, cfg.reduction_factor = hp.num_mels, hp.num_freq, hp.reduction_factor
    cfg.precision = PREC[precision]
    cfg.device = device
    cfg.prioritize_loss = 1 if hp.prioritize_loss else 0
    cfg.priority_hi = int(5000 / (hp.sample_rate * 0.5) * hp.num_freq)     # reference: tacotron.py:284-285
    cfg.priority_lo = int(165 / (hp.sample_rate * 0.5) * hp.num_freq)
    return cfg
