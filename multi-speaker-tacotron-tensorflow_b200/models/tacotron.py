"""Reference-facing mirror of ``models/tacotron.py``'s ``Tacotron`` (reference lines :16-343) on top of the CUDA engine.

TF graph mode is gone, so the protocol becomes eager but keeps names, argument meaning and error behaviour:
    model = create_model(hparams)
    model.initialize(inputs, input_lengths, num_speakers, speaker_id, mel_targets, linear_targets, loss_coeff, ...)
        -> runs the forward; sets .mel_outputs .linear_outputs .alignments (+ the attributes of tacotron.py:242-251)
    model.add_loss()              -> .loss .mel_loss .linear_loss .loss_without_coeff   (tacotron.py:274-302)
    model.add_optimizer(step)     -> applies clip + Adam (+ BN update); sets .learning_rate (tacotron.py:305-336)
``train.py``'s ``sess.run([global_step, loss, optimize])`` is ``model.train_step(batch)``.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .. import params as P
from ..engine import Engine


class Tacotron:
    def __init__(self, hparams, precision: str = "tf32", device: int = 0, seed: int = 4321):
        self._hparams = hparams
        self._precision = precision
        self._device = device
        self._seed = seed
        self.engine: Optional[Engine] = None
        self.num_speakers = None
        self.is_manual_attention = False          # tacotron.py:120-125 (placeholders become plain attributes)
        self.manual_alignments = None
        self._loss_ready = False

    # ---- tacotron.py:21-271 -----------------------------------------------------------------------------
    def initialize(self, inputs, input_lengths, num_speakers, speaker_id, mel_targets=None, linear_targets=None,
                   loss_coeff=None, rnn_decoder_test_mode=False, is_randomly_initialized=False):
        hp = self._hparams
        is_training = linear_targets is not None                      # tacotron.py:26
        self.is_randomly_initialized = is_randomly_initialized
        if num_speakers > 1 and hp.model_type not in ("deepvoice", "simple"):
            raise Exception(" [!] Unkown multi-speaker model type: {}".format(hp.model_type))     # tacotron.py:87-88
        if self.engine is None or self.num_speakers != num_speakers:
            self.engine = Engine(hp, num_speakers, precision=self._precision, device=self._device, seed=self._seed)
        self.num_speakers = num_speakers
        manual = self.manual_alignments if self.is_manual_attention else None
        steps = 0 if is_training else hp.max_iters                     # tacotron.py:207-210
        out = self.engine.forward(inputs, input_lengths, speaker_id, mel_targets, linear_targets, loss_coeff,
                                  decoder_steps=steps, rnn_decoder_test_mode=rnn_decoder_test_mode, manual_alignments=manual)
        self.inputs, self.speaker_id, self.input_lengths, self.loss_coeff = inputs, speaker_id, input_lengths, loss_coeff
        self.mel_outputs, self.linear_outputs, self.alignments = out["mel_outputs"], out["linear_outputs"], out["alignments"]
        self.mel_targets, self.linear_targets = mel_targets, linear_targets
        self.final_decoder_state = None
        self._loss_ready = False
        return self

    # ---- tacotron.py:274-302 ----------------------------------------------------------------------------
    def add_loss(self):
        """Loss + gradients (the engine fuses the L1 losses with their gradient and runs the backward pass)."""
        if self.linear_targets is None:
            raise RuntimeError("add_loss: initialize must have been called with targets")
        self.engine.backward()
        sc = self.engine.scalars()
        self.loss, self.mel_loss, self.linear_loss = sc["loss"], sc["mel_loss"], sc["linear_loss"]
        self.loss_without_coeff = sc["loss_without_coeff"]
        self._loss_ready = True

    # ---- tacotron.py:305-336 ----------------------------------------------------------------------------
    def add_optimizer(self, global_step: Optional[int] = None, allreduce=None):
        if not self._loss_ready:
            raise RuntimeError("add_optimizer: add_loss must have been called")
        if global_step is not None:
            self.engine.global_step = int(global_step)
        scale = allreduce(self.engine.grads) if allreduce is not None else 1.0
        self.gradients = self.engine.named_gradients()
        self.engine.optimizer_step(self.is_randomly_initialized, scale)
        self.learning_rate = self.engine.scalars()["learning_rate"]
        self._loss_ready = False

    def train_step(self, batch, num_speakers: int = 1, is_randomly_initialized: bool = True, allreduce=None):
        self.initialize(batch["inputs"], batch["input_lengths"], num_speakers, batch.get("speaker_id"),
                        batch["mel_targets"], batch["linear_targets"], batch.get("loss_coeff"),
                        is_randomly_initialized=is_randomly_initialized)
        self.add_loss()
        self.add_optimizer(allreduce=allreduce)
        return self.loss

    def get_dummy_feed_dict(self):                                     # tacotron.py:338-343
        return {"is_manual_attention": False, "manual_alignments": np.zeros([1, 1, 1])}

    # ---- checkpoints (train.py:175,242-244: tf.train.Saver -> a torch state file in the same directory layout) ----
    def state_dict(self):
        e = self.engine
        return dict(params=e.params.cpu(), bn_state=e.bn_state.cpu(), adam_m=e.adam_m.cpu(), adam_v=e.adam_v.cpu(),
                    global_step=e.global_step, adam_step=e.adam_step, names=[s.name for s in e.specs], hparams=self._hparams.values(),
                    num_speakers=self.num_speakers)

    def load_state_dict(self, sd, reset_step: bool = False):
        e = self.engine
        if [s.name for s in e.specs] != list(sd["names"]):
            raise RuntimeError("checkpoint tensor inventory does not match this model's hyper-parameters")
        e.params.copy_(sd["params"]); e.bn_state.copy_(sd["bn_state"])
        e.adam_m.copy_(sd["adam_m"]); e.adam_v.copy_(sd["adam_v"])
        e.adam_step = int(sd.get("adam_step", sd["global_step"]))      # the optimizer state is restored either way
        e.global_step = 0 if reset_step else int(sd["global_step"])    # train.py:194-205 (--initialize_path resets the step)
