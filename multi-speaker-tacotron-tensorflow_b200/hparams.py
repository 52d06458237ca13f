"""TensorFlow-free hyper-parameter container for the B200 Tacotron hot path.

Mirrors the key set and active defaults of the reference's ``hparams.py``
(reference: hparams.py:7-150 — the ``basic_params`` dict, the sample-rate
override at :26-29 and the *active* "Deep Voice 2" override block at :83-94)
and the ``params.json`` round trip of ``utils/__init__.py:100-126``.

The reference wraps the dict in ``tf.contrib.training.HParams``; callers only
use attribute access, ``values()`` and ``to_json()``, which is what this class
offers.  Nothing here touches TensorFlow or CUDA.
"""
from __future__ import annotations

import copy
import json
import os
from typing import Any, Dict, Iterable

PARAMS_NAME = "params.json"  # reference: utils/__init__.py:13

# Keys grouped the way the reference groups them; values are the *effective*
# defaults after the reference's own override blocks have run.
_AUDIO = dict(
    num_mels=80,
    num_freq=1025,
    sample_rate=24000,        # 20000 overridden to 24000 (hparams.py:26-29)
    frame_length_ms=50,
    frame_shift_ms=12.5,
    preemphasis=0.97,
    min_level_db=-100,
    ref_level_db=20,
)

_MODEL = dict(
    model_type="single",      # single | simple | deepvoice (hparams.py:33)
    speaker_embedding_size=16,
    embedding_size=256,
    dropout_prob=0.8,         # 0.5 overridden by the active DV2 block (hparams.py:85)
    # encoder
    enc_prenet_sizes=[256, 128],
    enc_bank_size=16,
    enc_bank_channel_size=128,
    enc_maxpool_width=2,
    enc_highway_depth=4,
    enc_rnn_size=128,
    enc_proj_sizes=[128, 128],
    enc_proj_width=3,
    # attention
    attention_type="bah_mon",
    attention_size=256,
    attention_state_size=256,
    # decoder
    dec_layer_num=2,
    dec_rnn_size=256,
    dec_prenet_sizes=[256, 128],
    post_bank_size=8,
    post_bank_channel_size=256,
    post_maxpool_width=2,
    post_highway_depth=4,
    post_rnn_size=256,        # 128 overridden by the active DV2 block (hparams.py:91)
    post_proj_sizes=[256, 80],
    post_proj_width=3,
    reduction_factor=4,       # reference default; BASELINE configs override to 5
)

_TRAIN = dict(
    batch_size=16,
    adam_beta1=0.9,
    adam_beta2=0.999,
    use_fixed_test_inputs=False,
    initial_learning_rate=0.002,
    decay_learning_rate_mode=0,
    initial_data_greedy=True,
    initial_phase_step=8000,
    main_data_greedy_factor=0,
    main_data=[""],
    prioritize_loss=False,
    recognition_loss_coeff=0.2,
    ignore_recognition_level=1,
)

_EVAL = dict(
    min_tokens=50,
    min_iters=30,
    max_iters=200,
    skip_inadequate=False,
    griffin_lim_iters=60,
    power=1.5,
)

_TEXT = dict(cleaners="korean_cleaners")


def default_values() -> Dict[str, Any]:
    out: Dict[str, Any] = {}
    for block in (_TEXT, _AUDIO, _MODEL, _TRAIN, _EVAL):
        out.update(copy.deepcopy(block))
    return out


class HParams:
    """Attribute bag with the small API surface the reference relies on."""

    def __init__(self, **kwargs: Any) -> None:
        self.__dict__["_v"] = {}
        for k, v in kwargs.items():
            self._v[k] = v

    # attribute protocol -------------------------------------------------
    def __getattr__(self, name: str) -> Any:
        try:
            return self.__dict__["_v"][name]
        except KeyError as e:
            raise AttributeError(name) from e

    def __setattr__(self, name: str, value: Any) -> None:
        self._v[name] = value

    def __contains__(self, name: str) -> bool:
        return name in self._v

    # HParams-like helpers -----------------------------------------------
    def values(self) -> Dict[str, Any]:
        return dict(self._v)

    def keys(self) -> Iterable[str]:
        return self._v.keys()

    def add_hparam(self, name: str, value: Any) -> None:
        if name in self._v:
            raise ValueError("hyperparameter %r already exists" % name)
        self._v[name] = value

    def set_hparam(self, name: str, value: Any) -> None:
        if name not in self._v:
            raise KeyError(name)
        self._v[name] = value

    def override(self, **kwargs: Any) -> "HParams":
        """Return a copy with some values replaced (unknown keys are errors)."""
        new = HParams(**copy.deepcopy(self._v))
        for k, v in kwargs.items():
            if k not in new._v and k != "num_speakers":
                raise KeyError("unknown hyperparameter %r" % k)
            new._v[k] = v
        return new

    def parse(self, spec: str) -> "HParams":
        """``name=value,name=value`` overrides, values parsed as JSON (lists such as ``enc_proj_sizes=[128,128]`` included:
        only commas outside brackets separate items)."""
        if not spec:
            return self
        items, depth, cur = [], 0, ""
        for ch in spec:
            if ch in "[{(":
                depth += 1
            elif ch in "]})":
                depth -= 1
            if ch == "," and depth == 0:
                items.append(cur); cur = ""
            else:
                cur += ch
        if cur.strip():
            items.append(cur)
        for item in items:
            name, _, raw = item.partition("=")
            name = name.strip()
            if name not in self._v:
                raise KeyError("unknown hyperparameter %r" % name)
            try:
                val = json.loads(raw)
            except ValueError:
                val = raw
            self._v[name] = val
        return self

    def to_json(self) -> str:
        return json.dumps(self._v, sort_keys=True)

    def __repr__(self) -> str:
        return "HParams(%s)" % ", ".join("%s=%r" % kv for kv in sorted(self._v.items()))


hparams = HParams(**default_values())


def hparams_debug_string(hp: HParams | None = None) -> str:
    vals = (hp or hparams).values()
    return "Hyperparameters:\n" + "\n".join("    %s: %s" % (k, vals[k]) for k in sorted(vals))


# params.json round trip (reference: utils/__init__.py:100-126) --------------
_NEVER_RELOADED = ("job_name", "num_workers", "display", "is_train", "load_path")


def save_hparams(model_dir: str, hp: HParams) -> str:
    path = os.path.join(model_dir, PARAMS_NAME)
    os.makedirs(model_dir, exist_ok=True)
    with open(path, "w", encoding="utf-8") as f:
        json.dump(hp.values(), f, indent=4, sort_keys=True, ensure_ascii=False)
    return path


def load_hparams(hp: HParams, load_path: str, skip_list: Iterable[str] = ()) -> HParams:
    path = os.path.join(load_path, PARAMS_NAME)
    with open(path, encoding="utf-8") as f:
        stored = json.load(f)
    skip = set(skip_list)
    for key, value in stored.items():
        if key in skip or key not in hp:
            continue
        if key in _NEVER_RELOADED:
            continue
        if getattr(hp, key) != value:
            setattr(hp, key, value)
    return hp


def stft_parameters(hp: HParams):
    """(n_fft, hop_length, win_length) — reference: audio/__init__.py:118-122."""
    n_fft = (hp.num_freq - 1) * 2
    hop_length = int(hp.frame_shift_ms / 1000 * hp.sample_rate)
    win_length = int(hp.frame_length_ms / 1000 * hp.sample_rate)
    return n_fft, hop_length, win_length
