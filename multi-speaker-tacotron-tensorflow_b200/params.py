"""Parameter inventory of the Tacotron hot path, in TensorFlow layouts.

One ordered table describes every tensor the reference's graph creates
(reference: models/tacotron.py:34-94,101-112,127-181,219-235 and
models/modules.py:11-131).  Layout conventions are TF's:
dense ``kernel[in,out]``, conv ``kernel[k,in,out]``, GRU ``gates_kernel[(x;h),
(r|u)]`` / ``cand_kernel[(x;h), h]``, batch-norm ``gamma,beta[C]`` trainable and
``moving_mean,moving_var[C]`` as non-trainable state.

All trainable tensors live in ONE flat fp32 buffer (so gradient all-reduce,
global-norm clip and Adam are single passes); batch-norm moving statistics live
in a second flat buffer.  Offsets are multiples of 4 elements (16 bytes) so the
kernels may use 128-bit accesses on any tensor start.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

import torch

NUM_SYMBOLS = 80  # reference: text/symbols.py:13 (PAD, EOS + jamo + punctuation + space)
PAD_ID, EOS_ID = 0, 1

ALIGN = 8  # elements (16 bytes in the bf16 mirror of the flat buffer)


@dataclass(frozen=True)
class ParamSpec:
    name: str
    shape: Tuple[int, ...]
    init: str              # glorot | zeros | ones | const:<v> | tnormal:<std> | glorot1d
    trainable: bool = True

    @property
    def numel(self) -> int:
        n = 1
        for s in self.shape:
            n *= s
        return n


def _dense(prefix, cin, cout, bias=True, bias_init="zeros"):
    out = [ParamSpec(prefix + "/kernel", (cin, cout), "glorot")]
    if bias:
        out.append(ParamSpec(prefix + "/bias", (cout,), bias_init))
    return out


def _conv_bn(prefix, k, cin, cout):
    return [
        ParamSpec(prefix + "/kernel", (k, cin, cout), "glorot"),
        ParamSpec(prefix + "/bias", (cout,), "zeros"),
        ParamSpec(prefix + "/gamma", (cout,), "ones"),
        ParamSpec(prefix + "/beta", (cout,), "zeros"),
        ParamSpec(prefix + "/moving_mean", (cout,), "zeros", trainable=False),
        ParamSpec(prefix + "/moving_var", (cout,), "ones", trainable=False),
    ]


def _gru(prefix, cin, h):
    # reference: TF r1.4 GRUCell — gate bias initialised to 1, candidate bias to 0
    return [
        ParamSpec(prefix + "/gates_kernel", (cin + h, 2 * h), "glorot"),
        ParamSpec(prefix + "/gates_bias", (2 * h,), "ones"),
        ParamSpec(prefix + "/cand_kernel", (cin + h, h), "glorot"),
        ParamSpec(prefix + "/cand_bias", (h,), "zeros"),
    ]


def _cbhg(prefix, cin, bank_size, bank_ch, proj_sizes, proj_width, depth, rnn):
    """Storage order note: the per-k bank tensors of one kind are emitted back to back so that
    e.g. all bank biases form one contiguous [bank_size*bank_ch] vector in the flat buffer (the
    kernels address the concatenated conv bank, modules.py:42-44, as one [.., K*C] matrix).
    The named TF-layout tensors stay individually addressable."""
    specs: List[ParamSpec] = []
    bank = [_conv_bn("%s/bank_%d" % (prefix, k), k, cin, bank_ch) for k in range(1, bank_size + 1)]
    for field in range(6):          # kernel, bias, gamma, beta, moving_mean, moving_var
        specs += [b[field] for b in bank]
    c = bank_size * bank_ch
    for i, p in enumerate(proj_sizes):
        specs += _conv_bn("%s/proj_%d" % (prefix, i + 1), proj_width, c, p)
        c = p
    if c != rnn:  # reference: modules.py:72-73
        specs += _dense(prefix + "/highway_in", c, rnn)
    for i in range(depth):
        specs += [
            ParamSpec("%s/highway_%d/H_kernel" % (prefix, i + 1), (rnn, rnn), "glorot"),
            ParamSpec("%s/highway_%d/H_bias" % (prefix, i + 1), (rnn,), "zeros"),
            ParamSpec("%s/highway_%d/T_kernel" % (prefix, i + 1), (rnn, rnn), "glorot"),
            ParamSpec("%s/highway_%d/T_bias" % (prefix, i + 1), (rnn,), "const:-1.0"),
        ]
    specs += _gru(prefix + "/gru_fw", rnn, rnn)
    specs += _gru(prefix + "/gru_bw", rnn, rnn)
    return specs


def speaker_mode(hp, num_speakers: int) -> str:
    """none | simple | deepvoice | deepvoice_table  (reference: tacotron.py:41-94)."""
    if num_speakers <= 1:
        return "none"
    if hp.model_type == "simple":
        return "simple"
    if hp.model_type == "deepvoice":
        return "deepvoice_table" if hp.speaker_embedding_size == 1 else "deepvoice"
    raise ValueError(" [!] Unkown multi-speaker model type: {}".format(hp.model_type))


def param_specs(hp, num_speakers: int = 1) -> List[ParamSpec]:
    mode = speaker_mode(hp, num_speakers)
    E = hp.embedding_size
    S = hp.speaker_embedding_size
    spk_cat = S if mode == "simple" else 0
    enc_out = 2 * hp.enc_rnn_size
    A = hp.attention_size
    ha = hp.attention_state_size
    hd = hp.dec_rnn_size
    r = hp.reduction_factor
    specs: List[ParamSpec] = [ParamSpec("embedding", (NUM_SYMBOLS, E), "tnormal:0.5")]

    if mode in ("simple", "deepvoice"):
        specs.append(ParamSpec("speaker_embedding", (num_speakers, S), "tnormal:0.5"))
    sites = [("before_highway", hp.enc_prenet_sizes[-1]),
             ("encoder_rnn_init_state", 2 * hp.enc_rnn_size),
             ("attention_rnn_init_state", ha)]
    sites += [("decoder_rnn_init_states%d" % (i + 1), hd) for i in range(hp.dec_layer_num)]
    if mode == "deepvoice":
        for nm, d in sites:  # tf.layers.dense(speaker_embed, d, softsign)  tacotron.py:68-79
            specs += _dense("speaker/" + nm, S, d)
    elif mode == "deepvoice_table":
        for nm, d in sites:  # get_embed tables  modules.py:11-15
            specs.append(ParamSpec("speaker/" + nm + "/table", (num_speakers, d), "tnormal:0.1"))

    c = E
    for i, sz in enumerate(hp.enc_prenet_sizes):
        specs += _dense("enc_prenet/dense_%d" % (i + 1), c, sz)
        c = sz
    specs += _cbhg("enc_cbhg", c, hp.enc_bank_size, hp.enc_bank_channel_size,
                   hp.enc_proj_sizes, hp.enc_proj_width, hp.enc_highway_depth, hp.enc_rnn_size)

    specs.append(ParamSpec("attention/memory_kernel", (enc_out, A), "glorot"))
    specs.append(ParamSpec("attention/query_kernel", (ha, A), "glorot"))
    specs.append(ParamSpec("attention/v", (A,), "glorot1d"))
    if hp.attention_type == "bah_mon":
        specs.append(ParamSpec("attention/score_bias", (1,), "zeros"))
    elif hp.attention_type == "bah_norm":
        specs.append(ParamSpec("attention/g", (1,), "const:%r" % math.sqrt(1.0 / A)))
        specs.append(ParamSpec("attention/b", (A,), "zeros"))
    elif hp.attention_type != "bah":
        raise ValueError(" [!] Unkown attention type: {}".format(hp.attention_type))

    c = hp.num_mels + enc_out  # decoder prenet sees [x_t ; context]  rnn_wrappers.py:249
    for i, sz in enumerate(hp.dec_prenet_sizes):
        specs += _dense("dec_prenet/dense_%d" % (i + 1), c, sz)
        c = sz
    specs += _gru("attention_gru", c + spk_cat, ha)
    specs += _dense("concat_proj", ha + enc_out + spk_cat, hd)
    for i in range(hp.dec_layer_num):
        specs += _gru("dec_gru_%d" % (i + 1), hd, hd)
    specs += _dense("mel_proj", hd, hp.num_mels * r)

    specs += _cbhg("post_cbhg", hp.num_mels, hp.post_bank_size, hp.post_bank_channel_size,
                   hp.post_proj_sizes, hp.post_proj_width, hp.post_highway_depth, hp.post_rnn_size)
    specs += _dense("linear", 2 * hp.post_rnn_size + spk_cat, hp.num_freq)
    return specs


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


@dataclass
class Layout:
    """Offsets of every tensor inside the two flat buffers."""
    specs: List[ParamSpec]
    offsets: Dict[str, int]
    n_trainable: int     # padded length of the trainable buffer
    n_state: int         # padded length of the BN-state buffer

    def spec(self, name: str) -> ParamSpec:
        for s in self.specs:
            if s.name == name:
                return s
        raise KeyError(name)


def make_layout(specs: Sequence[ParamSpec]) -> Layout:
    offs: Dict[str, int] = {}
    nt = ns = 0
    for s in specs:
        if s.trainable:
            offs[s.name] = nt
            nt += _round_up(s.numel, ALIGN)
        else:
            offs[s.name] = ns
            ns += _round_up(s.numel, ALIGN)
    return Layout(list(specs), offs, nt, max(ns, ALIGN))


def _fans(shape: Tuple[int, ...]) -> Tuple[int, int]:
    if len(shape) == 1:
        return shape[0], shape[0]
    if len(shape) == 2:
        return shape[0], shape[1]
    rf = 1
    for s in shape[:-2]:
        rf *= s
    return shape[-2] * rf, shape[-1] * rf


def init_tensor(spec: ParamSpec, gen: torch.Generator, dtype=torch.float32) -> torch.Tensor:
    """TF-1.x default initialisers (SURVEY.md §8c items 1,6,9,13)."""
    kind = spec.init
    if kind == "zeros":
        return torch.zeros(spec.shape, dtype=dtype)
    if kind == "ones":
        return torch.ones(spec.shape, dtype=dtype)
    if kind.startswith("const:"):
        return torch.full(spec.shape, float(kind[6:]), dtype=dtype)
    if kind in ("glorot", "glorot1d"):
        fi, fo = _fans(spec.shape)
        lim = math.sqrt(6.0 / (fi + fo))
        return ((torch.rand(spec.shape, generator=gen, dtype=torch.float64) * 2 - 1) * lim).to(dtype)
    if kind.startswith("tnormal:"):
        std = float(kind[8:])
        t = torch.randn(spec.shape, generator=gen, dtype=torch.float64)
        for _ in range(16):  # resample beyond 2 sigma, as tf.truncated_normal does
            bad = t.abs() > 2
            if not bad.any():
                break
            t = torch.where(bad, torch.randn(spec.shape, generator=gen, dtype=torch.float64), t)
        return (t.clamp(-2, 2) * std).to(dtype)
    raise ValueError(kind)


def init_params(hp, num_speakers: int = 1, seed: int = 4321, dtype=torch.float32,
                randomize_bn_state: bool = False) -> Dict[str, torch.Tensor]:
    """Name -> CPU tensor, deterministic in (hp, num_speakers, seed)."""
    gen = torch.Generator().manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    for s in param_specs(hp, num_speakers):
        t = init_tensor(s, gen, dtype)
        if randomize_bn_state and not s.trainable:
            if s.name.endswith("moving_mean"):
                t = (torch.randn(s.shape, generator=gen, dtype=torch.float64) * 0.1).to(dtype)
            else:
                t = (torch.rand(s.shape, generator=gen, dtype=torch.float64) + 0.5).to(dtype)
        out[s.name] = t
    return out


def flatten(named: Dict[str, torch.Tensor], layout: Layout, device="cpu"):
    """Pack a name->tensor dict into (trainable_flat, state_flat) fp32 buffers."""
    flat = torch.zeros(layout.n_trainable, dtype=torch.float32, device=device)
    state = torch.zeros(layout.n_state, dtype=torch.float32, device=device)
    for s in layout.specs:
        dst = flat if s.trainable else state
        o = layout.offsets[s.name]
        dst[o:o + s.numel] = named[s.name].reshape(-1).to(device=device, dtype=torch.float32)
    return flat, state


def views(flat: torch.Tensor, state: torch.Tensor, layout: Layout) -> Dict[str, torch.Tensor]:
    """Zero-copy named views into the flat buffers."""
    out: Dict[str, torch.Tensor] = {}
    for s in layout.specs:
        src = flat if s.trainable else state
        o = layout.offsets[s.name]
        out[s.name] = src[o:o + s.numel].view(s.shape)
    return out
