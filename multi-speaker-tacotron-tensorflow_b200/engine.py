"""Device-side driver of the hot path: owns the flat parameter / optimizer buffers and the workspace (torch tensors as
the device container) and calls the C ABI.  This is the layer the reference-facing ``models.Tacotron`` mirror sits on.

reference call sites it serves: train.py:145-155,217-219 (train step), synthesizer.py:47-54,166-167 (inference).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import capi
from . import params as P


def _require_cuda(device: int) -> torch.device:
    if not torch.cuda.is_available():
        raise capi.TacoError("libtaco_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", device)


class PendingScalars:
    """Handle of an in-flight scalar read-back (Engine.scalars_async)."""

    def __init__(self, engine, raw, event):
        self._engine, self._raw, self._event = engine, raw, event

    def get(self) -> Dict[str, float]:
        self._event.synchronize()
        out = capi.TacoStepScalars()
        capi.check(self._engine.lib.taco_finish_scalars(self._engine._h, self._raw.data_ptr(), C.byref(out)))
        return {k: float(getattr(out, k)) for k, _ in capi.TacoStepScalars._fields_}


class Engine:
    def __init__(self, hp, num_speakers: int = 1, precision: str = "fp32", device: int = 0, seed: int = 4321,
                 named_params: Optional[Dict[str, torch.Tensor]] = None, randomize_bn_state: bool = False):
        self.lib = capi.load()
        self.dev = _require_cuda(device)
        self.hp = hp
        self.num_speakers = num_speakers
        self.speaker_mode = P.speaker_mode(hp, num_speakers)
        self.precision = precision
        self.specs = P.param_specs(hp, num_speakers)
        self.layout = P.make_layout(self.specs)
        if named_params is None:
            named_params = P.init_params(hp, num_speakers, seed, randomize_bn_state=randomize_bn_state)
        with torch.cuda.device(self.dev):
            self.params, self.bn_state = P.flatten(named_params, self.layout, device=self.dev)
            self.grads = torch.zeros_like(self.params)
            self.adam_m = torch.zeros_like(self.params)
            self.adam_v = torch.zeros_like(self.params)
        self._cfg = capi.make_config(hp, num_speakers, self.speaker_mode, precision, device, P.NUM_SYMBOLS)
        self._h = C.c_void_p()
        capi.check(self.lib.taco_create(C.byref(self._h), C.byref(self._cfg)))
        self._names = [s.name.encode() for s in self.specs]          # keep the C strings alive
        table = (capi.TacoParamEntry * len(self.specs))()
        for i, s in enumerate(self.specs):
            table[i].name = self._names[i]
            table[i].offset = self.layout.offsets[s.name]
            table[i].numel = s.numel
            table[i].trainable = 1 if s.trainable else 0
        self._table = table
        capi.check(self.lib.taco_bind_params(self._h, table, len(self.specs), self.params.data_ptr(), self.grads.data_ptr(),
                                             self.adam_m.data_ptr(), self.adam_v.data_ptr(), self.bn_state.data_ptr(),
                                             self.layout.n_trainable, self.layout.n_state))
        self._ws: Optional[torch.Tensor] = None
        self._ws_key = None
        self._batch = None
        self._keep = []
        self.global_step = 0
        self.adam_step = 0          # Adam updates applied to adam_m / adam_v (TF: beta1_power, beta2_power)

    # ---- lifetime ---------------------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.taco_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- parameters -------------------------------------------------------------------------------------
    def named_parameters(self) -> Dict[str, torch.Tensor]:
        return P.views(self.params, self.bn_state, self.layout)

    def named_gradients(self) -> Dict[str, torch.Tensor]:
        return {k: v for k, v in P.views(self.grads, self.bn_state, self.layout).items() if self.layout.spec(k).trainable}

    def load_named(self, named: Dict[str, torch.Tensor]) -> None:
        flat, state = P.flatten(named, self.layout, device=self.dev)
        self.params.copy_(flat)
        self.bn_state.copy_(state)

    # ---- workspace --------------------------------------------------------------------------------------
    def _ensure_workspace(self, N: int, T_in: int, T_out_or_steps: int, training: bool) -> None:
        key = (N, T_in, T_out_or_steps, training)
        if key == self._ws_key:
            return
        nbytes = C.c_size_t()
        capi.check(self.lib.taco_workspace_bytes(self._h, N, T_in, T_out_or_steps, 1 if training else 0, C.byref(nbytes)))
        need = int(nbytes.value)
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            with torch.cuda.device(self.dev):
                self._ws = torch.zeros(need, dtype=torch.uint8, device=self.dev)
        else:
            self._ws.zero_()
        capi.check(self.lib.taco_bind_workspace(self._h, self._ws.data_ptr(), self._ws.numel()))
        self._ws_key = key

    def region(self, name: str) -> torch.Tensor:
        """Zero-copy view of a named workspace region (valid until the next shape change)."""
        off, numel, ndim = C.c_size_t(), C.c_int64(), C.c_int32()
        dims, strides = (C.c_int64 * 4)(), (C.c_int64 * 4)()
        capi.check(self.lib.taco_ws_region(self._h, name.encode(), C.byref(off), C.byref(numel), dims, strides, C.byref(ndim)))
        nv = ndim.value                       # sign / offset encode the element type: -n double, 100+n bf16 mirror, n fp32
        if nv >= 100:
            nd, dt, esz = nv - 100, torch.bfloat16, 2
        else:
            nd = abs(nv)
            dt, esz = (torch.float64, 8) if nv < 0 else (torch.float32, 4)
        base = self._ws.view(dt)
        return torch.as_strided(base, [dims[i] for i in range(nd)], [strides[i] for i in range(nd)], off.value // esz)

    # ---- block level (taco_cbhg_* / taco_decoder_*: SURVEY.md 8b) -------------------------------------------
    def plan(self, N: int, T_in: int, T_out: int, training: bool = True) -> None:
        """Size and bind the workspace for a batch shape without running anything (block-level calls need a planned shape)."""
        self._ensure_workspace(N, T_in, T_out if training else T_out // self.hp.reduction_factor, training)

    def cbhg_forward(self, which: int, inputs, input_lengths=None, before_highway=None, rnn_init_state=None, is_training: bool = True):
        x = self._f32(inputs)
        N, T, _ = x.shape
        H2 = 2 * (self.hp.post_rnn_size if which else self.hp.enc_rnn_size)
        out = torch.empty(N, T, H2, device=self.dev, dtype=torch.float32)
        L, bh, h0 = self._i32(input_lengths), self._f32(before_highway), self._f32(rnn_init_state)
        ptr = lambda t: None if t is None else t.data_ptr()
        self._blk_keep = [x, L, bh, h0]
        capi.check(self.lib.taco_cbhg_forward(self._h, which, x.data_ptr(), ptr(L), ptr(bh), ptr(h0), N, T, 1 if is_training else 0,
                                              out.data_ptr(), self._stream()))
        return out

    def cbhg_backward(self, which: int, d_outputs, input_lengths=None, want_before: bool = False, want_init_state: bool = False):
        """Gradients accumulate into ``self.grads`` (zero it first); returns (d_inputs, d_before_highway, d_rnn_init_state)."""
        dy = self._f32(d_outputs)
        N, T, _ = dy.shape
        cin = self.hp.num_mels if which else self.hp.enc_prenet_sizes[-1]
        p2 = self.hp.post_proj_sizes[-1] if which else self.hp.enc_proj_sizes[-1]
        H2 = 2 * (self.hp.post_rnn_size if which else self.hp.enc_rnn_size)
        dx = torch.empty(N, T, cin, device=self.dev, dtype=torch.float32)
        db = torch.empty(N, p2, device=self.dev) if want_before else None
        dh = torch.empty(N, H2, device=self.dev) if want_init_state else None
        L = self._i32(input_lengths)
        ptr = lambda t: None if t is None else t.data_ptr()
        capi.check(self.lib.taco_cbhg_backward(self._h, which, dy.data_ptr(), ptr(L), N, T, dx.data_ptr(), ptr(db), ptr(dh), self._stream()))
        return dx, db, dh

    def _batch_struct(self, inputs, input_lengths, speaker_id, mel_targets, linear_targets, decoder_steps, manual_alignments, test_mode):
        b = capi.TacoBatch()
        N, T_in = inputs.shape
        b.N, b.T_in, b.T_out = N, T_in, (mel_targets.shape[1] if mel_targets is not None else 0)
        ptr = lambda t: None if t is None else t.data_ptr()
        b.inputs, b.input_lengths, b.speaker_id = ptr(inputs), ptr(input_lengths), ptr(speaker_id)
        b.mel_targets, b.linear_targets, b.loss_coeff = ptr(mel_targets), ptr(linear_targets), None
        b.manual_alignments, b.decoder_steps, b.rnn_decoder_test_mode = ptr(manual_alignments), decoder_steps, 1 if test_mode else 0
        return b

    def decoder_forward(self, memory, inputs, input_lengths, speaker_id=None, mel_targets=None, linear_targets=None, decoder_steps: int = 0,
                        manual_alignments=None, rnn_decoder_test_mode: bool = False):
        """The attention decoder on a given encoder memory [N,T_in,2*enc_rnn] (``taco_decoder_forward``): teacher-forced when targets
        are given (pass ``linear_targets`` - any tensor of the right shape - to select the training plan), free-running otherwise."""
        mem = self._f32(memory)
        inputs, input_lengths, speaker_id = self._i32(inputs), self._i32(input_lengths), self._i32(speaker_id)
        mel_targets, linear_targets, manual_alignments = self._f32(mel_targets), self._f32(linear_targets), self._f32(manual_alignments)
        training = linear_targets is not None
        T_out = mel_targets.shape[1] if mel_targets is not None else 0
        self._ensure_workspace(inputs.shape[0], inputs.shape[1], T_out if training else (decoder_steps or T_out // self.hp.reduction_factor), training)
        b = self._batch_struct(inputs, input_lengths, speaker_id, mel_targets, linear_targets, decoder_steps, manual_alignments, rnn_decoder_test_mode)
        self._batch = b
        self._keep = [mem, inputs, input_lengths, speaker_id, mel_targets, linear_targets, manual_alignments]
        capi.check(self.lib.taco_decoder_forward(self._h, C.byref(b), mem.data_ptr(), self._stream()))
        return dict(mel_outputs=self.region("mel_outputs"), alignments=self.region("alignments"))

    def decoder_backward(self, d_mel_outputs):
        dm = self._f32(d_mel_outputs)
        N, T_in = self._batch.N, self._batch.T_in
        d_mem = torch.empty(N, T_in, 2 * self.hp.enc_rnn_size, device=self.dev, dtype=torch.float32)
        capi.check(self.lib.taco_decoder_backward(self._h, C.byref(self._batch), dm.data_ptr(), d_mem.data_ptr(), self._stream()))
        return d_mem

    # ---- hot path ---------------------------------------------------------------------------------------
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.dev).cuda_stream

    def _i32(self, t):
        return None if t is None else t.to(device=self.dev, dtype=torch.int32).contiguous()

    def _f32(self, t):
        return None if t is None else t.to(device=self.dev, dtype=torch.float32).contiguous()

    def forward(self, inputs, input_lengths, speaker_id=None, mel_targets=None, linear_targets=None, loss_coeff=None,
                decoder_steps: int = 0, rnn_decoder_test_mode: bool = False, manual_alignments=None):
        inputs, input_lengths, speaker_id = self._i32(inputs), self._i32(input_lengths), self._i32(speaker_id)
        # linear targets may arrive as bfloat16 (half the host->device bytes of a step); everything else is fp32
        lin16 = linear_targets is not None and linear_targets.dtype == torch.bfloat16
        mel_targets = self._f32(mel_targets)
        linear_targets = linear_targets.to(device=self.dev).contiguous() if lin16 else self._f32(linear_targets)
        loss_coeff, manual_alignments = self._f32(loss_coeff), self._f32(manual_alignments)
        N, T_in = inputs.shape
        training = linear_targets is not None
        T_out = mel_targets.shape[1] if mel_targets is not None else 0
        self._ensure_workspace(N, T_in, T_out if training else (decoder_steps or (T_out // self.hp.reduction_factor)), training)
        b = capi.TacoBatch()
        b.N, b.T_in, b.T_out = N, T_in, T_out
        ptr = lambda t: None if t is None else t.data_ptr()
        b.inputs, b.input_lengths, b.speaker_id = ptr(inputs), ptr(input_lengths), ptr(speaker_id)
        b.mel_targets, b.linear_targets, b.loss_coeff = ptr(mel_targets), ptr(linear_targets), ptr(loss_coeff)
        b.manual_alignments = ptr(manual_alignments)
        b.decoder_steps = decoder_steps
        b.rnn_decoder_test_mode = 1 if rnn_decoder_test_mode else 0
        b.linear_targets_bf16 = 1 if lin16 else 0
        self._batch = b
        self._keep = [inputs, input_lengths, speaker_id, mel_targets, linear_targets, loss_coeff, manual_alignments]
        capi.check(self.lib.taco_forward(self._h, C.byref(b), self._stream()))
        return dict(mel_outputs=self.region("mel_outputs"), linear_outputs=self.region("linear_outputs"),
                    alignments=self.region("alignments"))

    def backward(self) -> None:
        if self._batch is None:
            raise capi.TacoError("backward() before forward()")
        capi.check(self.lib.taco_backward(self._h, C.byref(self._batch), self._stream()))

    def optimizer_step(self, is_randomly_initialized: bool = True, grad_scale: float = 1.0) -> None:
        hp = self.hp
        capi.check(self.lib.taco_optimizer_step(self._h, self.global_step, self.adam_step, 1 if is_randomly_initialized else 0,
                                                float(hp.initial_learning_rate), int(hp.decay_learning_rate_mode),
                                                float(hp.adam_beta1), float(hp.adam_beta2), float(grad_scale), self._stream()))
        self.global_step += 1
        self.adam_step += 1

    def scalars(self) -> Dict[str, float]:
        out = capi.TacoStepScalars()
        capi.check(self.lib.taco_read_scalars(self._h, C.byref(out), self._stream()))
        return {k: float(getattr(out, k)) for k, _ in capi.TacoStepScalars._fields_}

    def scalars_async(self) -> "PendingScalars":
        """Enqueue the read-back of this step's scalars without synchronising; ``.get()`` on the returned handle waits for
        it.  Lets a training loop fetch the loss of every step (train.py:217-226) while the next step is already enqueued."""
        if not hasattr(self, "_raw_ring"):
            self._raw_ring = [torch.empty(capi.TACO_SCALARS_RAW_BYTES, dtype=torch.uint8).pin_memory() for _ in range(8)]
            self._raw_next = 0
        raw = self._raw_ring[self._raw_next % len(self._raw_ring)]
        self._raw_next += 1
        capi.check(self.lib.taco_copy_scalars_async(self._h, raw.data_ptr(), self._stream()))
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.dev))
        return PendingScalars(self, raw, ev)

    def train_step(self, batch: Dict[str, torch.Tensor], is_randomly_initialized: bool = True, allreduce=None,
                   after_forward=None) -> None:
        """forward + loss/backward + (gradient all-reduce) + clip/Adam/BN update — the body of train.py:217-219.
        ``after_forward()`` is called once the forward pass is enqueued: the place to issue the host->device copy of the NEXT
        batch (gated on an event recorded there).  Measured at C2 (tools/e2e_diag.py): 113 MB of H2D traffic beside the
        forward pass costs 0.68 ms per step, beside the backward pass 0.08 ms."""
        self.forward(batch["inputs"], batch["input_lengths"], batch.get("speaker_id"), batch["mel_targets"],
                     batch["linear_targets"], batch.get("loss_coeff"))
        if after_forward is not None:
            after_forward()
        self.backward()
        scale = 1.0
        if allreduce is not None:
            scale = allreduce(self.grads)
        self.optimizer_step(is_randomly_initialized, scale)

    def launch_count(self) -> int:
        return int(self.lib.taco_launch_count())
