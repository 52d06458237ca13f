"""Block-level mirror of the reference's ``models/modules.py`` (prenet :18-25, cbhg :27-96, highwaynet :105-120, conv1d :123-131)
over the C ABI's operator / block entry points.

The reference builds these blocks as TensorFlow graph fragments that create their variables under a scope; here a block is a
call on device tensors with the variables passed in (``params``: name -> tensor, the names of ``params.py``), executed by the
same CUDA kernels the whole model runs (``taco_gemm`` with its fused epilogues, ``taco_batch_norm``, ``taco_highway_combine``,
``taco_cbhg_forward``).  Argument names and meaning follow the reference; sizes it passes positionally (bank size, widths ...)
are read from the shapes of the variables.  There is no PyTorch arithmetic in these functions: torch is the tensor container.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from .. import capi

_ACT = {None: 0, "relu": 1, "sigmoid": 2, "tanh": 3, "softsign": 4}


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _dense(x2d: torch.Tensor, kernel: torch.Tensor, bias: Optional[torch.Tensor], act, precision: str) -> torch.Tensor:
    """tf.layers.dense on a [rows, in] matrix: one taco_gemm call with the bias / activation epilogue."""
    rows, cin = x2d.shape
    cout = kernel.shape[1]
    out = torch.empty(rows, cout, device=x2d.device, dtype=torch.float32)
    d = capi.TacoGemmDesc()
    d.A, d.B, d.C = x2d.data_ptr(), kernel.data_ptr(), out.data_ptr()
    d.M, d.N, d.K, d.lda, d.ldb, d.ldc = rows, cout, cin, cin, cout, cout
    d.alpha, d.split_k = 1.0, 1
    d.bias = bias.data_ptr() if bias is not None else None
    d.act = _ACT[act]
    capi.check(capi.load().taco_gemm(C.byref(d), 1, capi.PREC[precision], _stream(x2d)))
    return out


def prenet(inputs: torch.Tensor, is_training: bool, params: Dict[str, torch.Tensor], scope: str = "prenet", precision: str = "fp32") -> torch.Tensor:
    """modules.py:18-25: relu(dense) per layer ``<scope>/dense_<i>``; ``tf.layers.dropout`` is called without ``training=True`` there,
    so it is the identity in both modes (``is_training`` is accepted and ignored, like the reference's ``drop_rate``)."""
    x = inputs.reshape(-1, inputs.shape[-1]).contiguous()
    i = 1
    while "%s/dense_%d/kernel" % (scope, i) in params:
        x = _dense(x, params["%s/dense_%d/kernel" % (scope, i)], params["%s/dense_%d/bias" % (scope, i)], "relu", precision)
        i += 1
    return x.reshape(*inputs.shape[:-1], x.shape[-1])


def highwaynet(inputs: torch.Tensor, params: Dict[str, torch.Tensor], scope: str, precision: str = "fp32") -> torch.Tensor:
    """modules.py:105-120: H = relu(dense), T = sigmoid(dense, bias -1), y = H*T + x*(1-T)."""
    x = inputs.reshape(-1, inputs.shape[-1]).contiguous()
    H = _dense(x, params[scope + "/H_kernel"], params[scope + "/H_bias"], "relu", precision)
    T = _dense(x, params[scope + "/T_kernel"], params[scope + "/T_bias"], "sigmoid", precision)
    y = torch.empty_like(x)
    capi.check(capi.load().taco_highway_combine(H.data_ptr(), T.data_ptr(), x.data_ptr(), y.data_ptr(), x.shape[0], x.shape[1], _stream(x)))
    return y.reshape(inputs.shape)


def conv1d(inputs: torch.Tensor, activation, is_training: bool, params: Dict[str, torch.Tensor], scope: str, precision: str = "fp32"):
    """modules.py:123-131: tf.layers.conv1d(padding='same') -> activation -> tf.layers.batch_normalization(training=is_training).
    The convolution is ONE taco_gemm over the zero-padded time layout (tap addressing, pad rows masked in the epilogue); returns
    (outputs [N,T,C], batch_mean, batch_var) - the batch moments feed the moving-statistics update in training mode."""
    kernel, bias = params[scope + "/kernel"], params[scope + "/bias"]
    k, cin, cout = kernel.shape
    N, T, _ = inputs.shape
    left = (k - 1) // 2
    Tp = T + k - 1
    rows = N * Tp
    xp = torch.zeros(rows + 2 * k, cin, device=inputs.device, dtype=torch.float32)          # padded layout + slack rows for the taps
    xp[k:k + rows].view(N, Tp, cin)[:, left:left + T].copy_(inputs)
    raw = torch.empty(rows, cout, device=inputs.device, dtype=torch.float32)
    d = capi.TacoGemmDesc()
    d.A, d.B, d.C = xp[k - left:].data_ptr(), kernel.data_ptr(), raw.data_ptr()
    d.M, d.N, d.K, d.lda, d.ldb, d.ldc = rows, cout, k * cin, cin, cout, cout
    d.ctap, d.alpha, d.split_k = cin, 1.0, 1
    d.bias, d.act = bias.data_ptr(), _ACT[activation]
    d.mask_period, d.mask_lo, d.mask_hi = Tp, left, left + T
    capi.check(capi.load().taco_gemm(C.byref(d), 1, capi.PREC[precision], _stream(inputs)))
    y = raw.view(N, Tp, cout)[:, left:left + T].contiguous()
    out = torch.empty_like(y)
    mean = torch.empty(cout, device=inputs.device); var = torch.empty(cout, device=inputs.device)
    scratch = torch.empty(4 * cout, device=inputs.device, dtype=torch.float64)
    capi.check(capi.load().taco_batch_norm(y.data_ptr(), params[scope + "/gamma"].data_ptr(), params[scope + "/beta"].data_ptr(),
                                          params[scope + "/moving_mean"].data_ptr(), params[scope + "/moving_var"].data_ptr(), N, T, cout,
                                          1 if is_training else 0, out.data_ptr(), mean.data_ptr(), var.data_ptr(), scratch.data_ptr(), _stream(inputs)))
    return out, mean, var


def cbhg(engine, inputs: torch.Tensor, input_lengths: Optional[torch.Tensor], is_training: bool, scope: str,
         before_highway: Optional[torch.Tensor] = None, encoder_rnn_init_state: Optional[torch.Tensor] = None) -> torch.Tensor:
    """modules.py:27-96 as one call (``taco_cbhg_forward``): ``scope`` is ``'encoder_cbhg'`` / ``'enc_cbhg'`` or ``'post_cbhg'``; bank
    size, channel counts, projection sizes, highway depth and RNN size are those of the engine's hyper-parameters (the reference
    passes the same hparams fields positionally, tacotron.py:105-112,219-224).  The engine's workspace must be planned for a batch
    whose text length (encoder) / frame count (post-net) equals ``inputs.shape[1]``: see ``Engine.plan``."""
    which = 1 if scope.startswith("post") else 0
    return engine.cbhg_forward(which, inputs, input_lengths, before_highway, encoder_rnn_init_state, is_training)
