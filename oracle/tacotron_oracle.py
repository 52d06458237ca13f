"""CPU ORACLE — test infrastructure only, never the product path.

PARITY PINNING: the reference's arithmetic lives in TensorFlow 1.x (pinned ``tensorflow==1.3.0`` in
requirements.txt:101; the code needs r1.4 because of ``BahdanauMonotonicAttention``, models/tacotron.py:5), which cannot
be installed here, and the reference ships no tests, golden vectors or fixtures.  This file restates the reference graph
op for op from its source, using the TF r1.4 semantics listed in SURVEY.md §8(c) items 1-13.  It is pinned two ways:

  1. against the reference's OWN model code: ``tools/make_reference_golden.py`` imports the unmodified
     ``/root/reference/{hparams.py,models/*.py,text/symbols.py}`` over ``oracle/tf1_shim`` (an eager stand-in for the TF
     r1.4 Python API) and writes ``tests/golden/ref_*.npz``; ``tests/test_reference_golden.py`` requires this oracle to
     reproduce every output, loss, gradient, Adam update and batch-norm statistic of those runs (fp32: <=5e-6; fp64:
     <=1e-12) for all speaker modes, attention types, manual attention, rnn_decoder_test_mode, prioritize_loss, both LR
     schedules and ragged lengths.  That pins the wiring, sizes, order of operations, feeding rules, loss and optimizer
     recipe to the reference's code.  What it cannot pin is TensorFlow's library arithmetic itself (GRUCell, monotonic
     attention, batch norm, SAME padding ... are restated in the shim from the TF sources, independently of this file):
     status "pinned to the reference's Python, TF kernels restated twice and cross-checked" — not "ran TensorFlow";
  2. by the self-tests in ``tests/test_oracle.py`` (recursive vs closed-form monotonic attention, independent scalar GRU,
     torch.nn.functional equivalents, fp64 finite differences, hand-computed Adam step).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import this
module.

Every function cites the reference lines it follows.  Tensors are torch CPU tensors (fp32 or fp64); the three recurrences
are Python loops, mirroring the reference's ``tf.while_loop`` dispatch structure.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor

BN_MOMENTUM = 0.99   # tf.layers.batch_normalization default
BN_EPS = 1e-3        # tf.layers.batch_normalization default
FLT_TINY = float(np.finfo(np.float32).tiny)


# --------------------------------------------------------------------------
# building blocks (reference: models/modules.py)
# --------------------------------------------------------------------------
def dense(x: Tensor, w: Tensor, b: Optional[Tensor] = None) -> Tensor:
    """tf.layers.dense: y = x . W[in,out] + b on the last axis."""
    y = x @ w
    return y if b is None else y + b


def prenet(x: Tensor, P: Dict[str, Tensor], scope: str, n_layers: int) -> Tensor:
    """modules.py:18-25.  tf.layers.dropout is called without training=True, so it is the
    identity in both modes (SURVEY.md §2 'latent bugs'); the oracle mirrors that."""
    for i in range(n_layers):
        x = torch.relu(dense(x, P["%s/dense_%d/kernel" % (scope, i + 1)],
                             P["%s/dense_%d/bias" % (scope, i + 1)]))
    return x


def conv1d_same(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    """tf.layers.conv1d(padding='same'): cross-correlation, W[k,in,out], zero pad
    left=(k-1)//2, right=k-1-left.  x is [N,T,Cin] -> [N,T,Cout]."""
    k = w.shape[0]
    left = (k - 1) // 2
    xp = F.pad(x.transpose(1, 2), (left, k - 1 - left))           # [N,Cin,T+k-1]
    y = F.conv1d(xp, w.permute(2, 1, 0).contiguous(), b)          # weight [Cout,Cin,k]
    return y.transpose(1, 2)


def batch_norm(x: Tensor, P: Dict[str, Tensor], scope: str, is_training: bool,
               new_state: Optional[Dict[str, Tensor]]) -> Tensor:
    """tf.layers.batch_normalization(axis=-1, momentum=.99, eps=1e-3), non-fused 3-D path:
    training -> biased batch moments over (N,T) (pad frames included, modules.py:131);
    moving <- moving*.99 + batch*.01 (biased variance)."""
    g, bt = P[scope + "/gamma"], P[scope + "/beta"]
    if is_training:
        mean = x.mean(dim=(0, 1))
        var = ((x - mean) ** 2).mean(dim=(0, 1))
        if new_state is not None:
            with torch.no_grad():
                new_state[scope + "/moving_mean"] = P[scope + "/moving_mean"] * BN_MOMENTUM + mean * (1 - BN_MOMENTUM)
                new_state[scope + "/moving_var"] = P[scope + "/moving_var"] * BN_MOMENTUM + var * (1 - BN_MOMENTUM)
    else:
        mean, var = P[scope + "/moving_mean"], P[scope + "/moving_var"]
    return (x - mean) * torch.rsqrt(var + BN_EPS) * g + bt


def conv1d_bn(x, P, scope, activation, is_training, new_state):
    """modules.py:123-131 — conv -> activation -> batch-norm (activation BEFORE BN)."""
    y = conv1d_same(x, P[scope + "/kernel"], P[scope + "/bias"])
    if activation is not None:
        y = activation(y)
    return batch_norm(y, P, scope, is_training, new_state)


def maxpool_same_2(x: Tensor) -> Tensor:
    """tf.layers.max_pooling1d(pool=2, stride=1, 'same'): pad right with -inf."""
    nxt = torch.cat([x[:, 1:], torch.full_like(x[:, :1], -float("inf"))], dim=1)
    # torch.where (not torch.maximum): on ties the gradient goes to the FIRST element of the window, as TF's CPU
    # MaxPoolGrad does; torch.maximum would split it 50/50.  Ties are common here (BN of ReLU zeros).
    return torch.where(x >= nxt, x, nxt)


def highwaynet(x: Tensor, P: Dict[str, Tensor], scope: str) -> Tensor:
    """modules.py:105-120."""
    H = torch.relu(dense(x, P[scope + "/H_kernel"], P[scope + "/H_bias"]))
    T = torch.sigmoid(dense(x, P[scope + "/T_kernel"], P[scope + "/T_bias"]))
    return H * T + x * (1.0 - T)


def gru_cell(x: Tensor, h: Tensor, P: Dict[str, Tensor], scope: str) -> Tensor:
    """TF r1.3/1.4 GRUCell: [r,u]=sigmoid([x,h].Wg+bg); c=tanh([x,r*h].Wc+bc);
    h'=u*h+(1-u)*c.  Reset is applied BEFORE the candidate matmul."""
    H = h.shape[-1]
    gates = torch.sigmoid(torch.cat([x, h], -1) @ P[scope + "/gates_kernel"] + P[scope + "/gates_bias"])
    r, u = gates[..., :H], gates[..., H:]
    c = torch.tanh(torch.cat([x, r * h], -1) @ P[scope + "/cand_kernel"] + P[scope + "/cand_bias"])
    return u * h + (1.0 - u) * c


def reverse_sequence(x: Tensor, lengths: Tensor) -> Tensor:
    """tf.reverse_sequence(seq_axis=1, batch_axis=0)."""
    N, T = x.shape[:2]
    t = torch.arange(T).unsqueeze(0).expand(N, T)
    L = lengths.view(N, 1).to(torch.long)
    idx = torch.where(t < L, L - 1 - t, t)
    return torch.gather(x, 1, idx.unsqueeze(-1).expand_as(x))


def dynamic_rnn(x: Tensor, lengths: Optional[Tensor], h0: Tensor, P, scope) -> Tensor:
    """tf.nn.dynamic_rnn with sequence_length: steps >= length emit zeros and copy state."""
    N, T, _ = x.shape
    h = h0
    outs = []
    for t in range(T):
        hn = gru_cell(x[:, t], h, P, scope)
        if lengths is not None:
            m = (t < lengths).view(N, 1)
            outs.append(torch.where(m, hn, torch.zeros_like(hn)))
            h = torch.where(m, hn, h)
        else:
            outs.append(hn)
            h = hn
    return torch.stack(outs, 1)


def bidirectional_gru(x, lengths, P, scope, init_state=None):
    """modules.py:82-96 (tf.nn.bidirectional_dynamic_rnn)."""
    N = x.shape[0]
    H = P[scope + "/gru_fw/cand_bias"].shape[0]
    if init_state is not None:
        h_fw, h_bw = init_state[:, :H], init_state[:, H:]           # tf.split(state, 2, 1)
    else:
        h_fw = h_bw = x.new_zeros(N, H)
    out_fw = dynamic_rnn(x, lengths, h_fw, P, scope + "/gru_fw")
    if lengths is not None:
        xr = reverse_sequence(x, lengths)
        out_bw = reverse_sequence(dynamic_rnn(xr, lengths, h_bw, P, scope + "/gru_bw"), lengths)
    else:
        out_bw = dynamic_rnn(x.flip(1), None, h_bw, P, scope + "/gru_bw").flip(1)
    return torch.cat([out_fw, out_bw], -1)


def cbhg(x, lengths, is_training, P, scope, bank_size, n_proj, depth,
         before_highway=None, rnn_init_state=None, new_state=None, taps=None):
    """modules.py:27-96."""
    bank = torch.cat([conv1d_bn(x, P, "%s/bank_%d" % (scope, k), torch.relu, is_training, new_state)
                      for k in range(1, bank_size + 1)], -1)
    pooled = maxpool_same_2(bank)
    y = pooled
    for i in range(n_proj):
        act = None if i == n_proj - 1 else torch.relu
        y = conv1d_bn(y, P, "%s/proj_%d" % (scope, i + 1), act, is_training, new_state)
    hw = y + x
    if before_highway is not None:
        hw = hw + before_highway.unsqueeze(1)
    if (scope + "/highway_in/kernel") in P:
        hw = dense(hw, P[scope + "/highway_in/kernel"], P[scope + "/highway_in/bias"])
    if taps is not None:
        taps[scope + "/bank"] = bank
        taps[scope + "/highway_input"] = hw
    for i in range(depth):
        hw = highwaynet(hw, P, "%s/highway_%d" % (scope, i + 1))
    if taps is not None:
        taps[scope + "/rnn_input"] = hw
    return bidirectional_gru(hw, lengths, P, scope, rnn_init_state)


# --------------------------------------------------------------------------
# attention (reference: TF r1.4 attention_wrapper.py, called from rnn_wrappers.py:304-317)
# --------------------------------------------------------------------------
def safe_cumprod_exclusive(x: Tensor) -> Tensor:
    """exp(cumsum(log(clip(x, tiny, 1)), exclusive=True)) along axis 1."""
    lg = torch.log(torch.clamp(x, FLT_TINY, 1.0))
    cs = torch.cumsum(lg, 1) - lg
    return torch.exp(cs)


def monotonic_attention_parallel(p: Tensor, prev: Tensor) -> Tensor:
    cp = safe_cumprod_exclusive(1.0 - p)
    return p * cp * torch.cumsum(prev / torch.clamp(cp, 1e-10, 1.0), 1)


def monotonic_attention_recursive(p: Tensor, prev: Tensor) -> Tensor:
    """Raffel et al. 2017 definition: q_j=(1-p_{j-1})q_{j-1}+prev_j; a_j=p_j q_j (self-test)."""
    N, T = p.shape
    out = []
    q = torch.zeros_like(p[:, 0])
    for j in range(T):
        q = (q * (1.0 - p[:, j - 1]) if j > 0 else q) + prev[:, j]
        out.append(p[:, j] * q)
    return torch.stack(out, 1)


def attention_scores(q: Tensor, keys: Tensor, P, attention_type: str) -> Tensor:
    """_bahdanau_score (+ score bias for the monotonic variant)."""
    v = P["attention/v"]
    if attention_type == "bah_norm":
        nv = P["attention/g"] * v * torch.rsqrt((v * v).sum())
        return (nv * torch.tanh(keys + q.unsqueeze(1) + P["attention/b"])).sum(-1)
    s = (v * torch.tanh(keys + q.unsqueeze(1))).sum(-1)
    if attention_type == "bah_mon":
        s = s + P["attention/score_bias"]
    return s


def attention_probabilities(score: Tensor, prev: Tensor, attention_type: str) -> Tensor:
    if attention_type == "bah_mon":
        return monotonic_attention_parallel(torch.sigmoid(score), prev)
    return torch.softmax(score, -1)


# --------------------------------------------------------------------------
# whole forward (reference: models/tacotron.py:21-251)
# --------------------------------------------------------------------------
def softsign(x):
    return x / (1.0 + x.abs())


def speaker_vectors(P, hp, mode: str, speaker_id: Optional[Tensor]):
    """tacotron.py:41-94 -> dict(speaker_embed, before_highway, enc_init, att_init, dec_init[list])."""
    out = dict(speaker_embed=None, before_highway=None, enc_init=None, att_init=None, dec_init=None)
    if mode == "none":
        return out
    names = ["before_highway", "encoder_rnn_init_state", "attention_rnn_init_state"] + \
            ["decoder_rnn_init_states%d" % (i + 1) for i in range(hp.dec_layer_num)]
    if mode == "simple":
        out["speaker_embed"] = P["speaker_embedding"][speaker_id.long()]
        return out
    if mode == "deepvoice":
        e = P["speaker_embedding"][speaker_id.long()]
        vecs = [softsign(dense(e, P["speaker/%s/kernel" % n], P["speaker/%s/bias" % n])) for n in names]
    else:
        vecs = [P["speaker/%s/table" % n][speaker_id.long()] for n in names]
    out.update(before_highway=vecs[0], enc_init=vecs[1], att_init=vecs[2], dec_init=vecs[3:])
    return out


def forward(P: Dict[str, Tensor], hp, inputs: Tensor, input_lengths: Tensor,
            num_speakers: int = 1, speaker_id: Optional[Tensor] = None,
            mel_targets: Optional[Tensor] = None, linear_targets: Optional[Tensor] = None,
            rnn_decoder_test_mode: bool = False, manual_alignments: Optional[Tensor] = None,
            max_iters: Optional[int] = None, update_bn: bool = True,
            speaker_mode: Optional[str] = None, want_taps: bool = False):
    """Returns dict(mel_outputs[N,To,M], linear_outputs[N,To,F], alignments[N,Ti,Td],
    new_bn_state{...}, taps{...}).  is_training := linear_targets is not None (tacotron.py:26)."""
    is_training = linear_targets is not None
    if speaker_mode is None:
        from importlib import import_module
        speaker_mode = import_module("multi-speaker-tacotron-tensorflow_b200.params").speaker_mode(hp, num_speakers)
    new_state: Optional[Dict[str, Tensor]] = {} if (is_training and update_bn) else None
    taps: Optional[Dict[str, Tensor]] = {} if want_taps else None
    N, Ti = inputs.shape
    r, M = hp.reduction_factor, hp.num_mels
    dt = P["embedding"].dtype

    spk = speaker_vectors(P, hp, speaker_mode, speaker_id)

    # encoder -------------------------------------------------------------
    x = P["embedding"][inputs.long()]                                              # tacotron.py:34-39
    x = prenet(x, P, "enc_prenet", len(hp.enc_prenet_sizes))                       # :101-103
    memory = cbhg(x, input_lengths, is_training, P, "enc_cbhg", hp.enc_bank_size,
                  len(hp.enc_proj_sizes), hp.enc_highway_depth,
                  before_highway=spk["before_highway"], rnn_init_state=spk["enc_init"],
                  new_state=new_state, taps=taps)                                  # :105-112
    keys = memory @ P["attention/memory_kernel"]                                   # memory_layer, no bias, no mask

    # decoder -------------------------------------------------------------
    ha = spk["att_init"] if spk["att_init"] is not None else memory.new_zeros(N, hp.attention_state_size)
    hd = [spk["dec_init"][i] if spk["dec_init"] is not None else memory.new_zeros(N, hp.dec_rnn_size)
          for i in range(hp.dec_layer_num)]
    ctx = memory.new_zeros(N, memory.shape[-1])                                    # rnn_wrappers.py:209
    if hp.attention_type == "bah_mon":
        align = F.one_hot(torch.zeros(N, dtype=torch.long), Ti).to(dt)             # initial_alignments
    else:
        align = memory.new_zeros(N, Ti)
    xt = memory.new_zeros(N, M)                                                    # helpers.py:70-72
    if is_training:
        Td = mel_targets.shape[1] // r
        if max_iters is not None:
            Td = min(Td, max_iters)
        feed = mel_targets[:, r - 1::r, :]                                         # helpers.py:44
    else:
        Td = max_iters if max_iters is not None else hp.max_iters                  # helpers.py:29 never fires
    outs, hist = [], []
    for t in range(Td):
        u = torch.cat([xt, ctx], -1)                                               # rnn_wrappers.py:249
        z = prenet(u, P, "dec_prenet", len(hp.dec_prenet_sizes))                   # :367-369
        if spk["speaker_embed"] is not None:
            z = torch.cat([z, spk["speaker_embed"]], -1)                           # :371-376
        ha = gru_cell(z, ha, P, "attention_gru")                                   # :251
        q = ha @ P["attention/query_kernel"]
        score = attention_scores(q, keys, P, hp.attention_type)
        computed = attention_probabilities(score, align, hp.attention_type)
        align = manual_alignments[:, t, :] if manual_alignments is not None else computed  # :313-317
        ctx = torch.bmm(align.unsqueeze(1), memory).squeeze(1)                     # :333-334
        hist.append(align)
        cat = [ha, ctx] + ([spk["speaker_embed"]] if spk["speaker_embed"] is not None else [])
        y = dense(torch.cat(cat, -1), P["concat_proj/kernel"], P["concat_proj/bias"])   # :405-415, tacotron.py:170
        for i in range(hp.dec_layer_num):                                          # ResidualWrapper(GRUCell)
            hd[i] = gru_cell(y, hd[i], P, "dec_gru_%d" % (i + 1))
            y = y + hd[i]
        o = dense(y, P["mel_proj/kernel"], P["mel_proj/bias"])                     # tacotron.py:178-179
        outs.append(o)
        if is_training and not rnn_decoder_test_mode:
            xt = feed[:, t, :]                                                     # helpers.py:66
        else:
            xt = o[:, -M:]                                                         # helpers.py:31,64
    dec = torch.stack(outs, 1)                                                     # [N,Td,M*r]
    mel = dec.reshape(N, Td * r, M)                                                # tacotron.py:213-214
    alignments = torch.stack(hist, 2)                                              # [N,Ti,Td]  :238-239

    # post-net ------------------------------------------------------------
    post = cbhg(mel, None, is_training, P, "post_cbhg", hp.post_bank_size,
                len(hp.post_proj_sizes), hp.post_highway_depth, new_state=new_state, taps=taps)
    if spk["speaker_embed"] is not None:                                           # :226-233 (simple only)
        post = torch.cat([spk["speaker_embed"].unsqueeze(1).expand(-1, post.shape[1], -1), post], -1)
    linear = dense(post, P["linear/kernel"], P["linear/bias"])                     # :235
    if taps is not None:
        taps.update(memory=memory, keys=keys, decoder_outputs=dec, post_outputs=post)
    return dict(mel_outputs=mel, linear_outputs=linear, alignments=alignments,
                new_bn_state=new_state or {}, taps=taps or {},
                final_state=dict(ha=ha, hd=hd, ctx=ctx, align=align))


# --------------------------------------------------------------------------
# loss / optimiser (reference: models/tacotron.py:274-336)
# --------------------------------------------------------------------------
def losses(out, mel_targets, linear_targets, loss_coeff, hp):
    mel_l = (mel_targets - out["mel_outputs"]).abs()
    l1 = (linear_targets - out["linear_outputs"]).abs()
    c = loss_coeff.view(-1, 1, 1)
    if hp.prioritize_loss:
        hi = int(5000 / (hp.sample_rate * 0.5) * hp.num_freq)
        lo = int(165 / (hp.sample_rate * 0.5) * hp.num_freq)
        lp = l1[:, :, lo:hi]
        loss = (mel_l * c).mean() + 0.5 * (l1 * c).mean() + 0.5 * (lp * c).mean()
        linear_loss = 0.5 * (l1.mean() + lp.mean())
    else:
        loss = (mel_l * c).mean() + (l1 * c).mean()
        linear_loss = l1.mean()
    mel_loss = mel_l.mean()
    return dict(loss=loss, mel_loss=mel_loss, linear_loss=linear_loss,
                loss_without_coeff=mel_loss + linear_loss)


def learning_rate(hp, global_step: int, is_randomly_initialized: bool) -> float:
    """tacotron.py:314-326 with step = global_step + 1."""
    step = float(global_step + 1)
    if hp.decay_learning_rate_mode == 0:
        w = 4000.0 if is_randomly_initialized else 40000.0
        return hp.initial_learning_rate * w ** 0.5 * min(step * w ** -1.5, step ** -0.5)
    return hp.initial_learning_rate * 0.95 ** (step / 3000.0)


def clip_by_global_norm(grads: Dict[str, Tensor], clip_norm: float = 1.0):
    gn = math.sqrt(sum(float((g.double() ** 2).sum()) for g in grads.values()))
    scale = clip_norm * min(1.0 / gn, 1.0 / clip_norm) if gn > 0 else 1.0
    return {k: g * scale for k, g in grads.items()}, gn


def adam_step(P, grads, m, v, t: int, lr: float, b1=0.9, b2=0.999, eps=1e-8):
    """tf.train.AdamOptimizer: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); p -= lr_t*m/(sqrt(v)+eps)."""
    lr_t = lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    for k, g in grads.items():
        m[k] = b1 * m[k] + (1 - b1) * g
        v[k] = b2 * v[k] + (1 - b2) * g * g
        P[k] = P[k] - lr_t * m[k] / (v[k].sqrt() + eps)
    return P, m, v


def train_step(P, m, v, hp, batch, global_step: int, is_randomly_initialized=True,
               num_speakers=1, speaker_mode=None):
    """One reference training step (train.py:217-219): forward, loss, grads, clip, Adam, BN update.
    P/m/v are name->tensor dicts (trainable + BN state in P); returns scalars + updated dicts."""
    names = [k for k in P if not (k.endswith("moving_mean") or k.endswith("moving_var"))]
    leaf = {k: (P[k].detach().clone().requires_grad_(True) if k in names else P[k]) for k in P}
    out = forward(leaf, hp, batch["inputs"], batch["input_lengths"], num_speakers,
                  batch.get("speaker_id"), batch["mel_targets"], batch["linear_targets"],
                  speaker_mode=speaker_mode)
    ls = losses(out, batch["mel_targets"], batch["linear_targets"], batch["loss_coeff"], hp)
    gl = torch.autograd.grad(ls["loss"], [leaf[k] for k in names], allow_unused=True)
    grads = {k: (g if g is not None else torch.zeros_like(P[k])) for k, g in zip(names, gl)}
    clipped, gn = clip_by_global_norm(grads, 1.0)
    lr = learning_rate(hp, global_step, is_randomly_initialized)
    newP = {k: P[k].detach() for k in P}
    sub = {k: newP[k] for k in names}
    sub, m, v = adam_step(sub, clipped, m, v, global_step + 1, lr, hp.adam_beta1, hp.adam_beta2)
    newP.update(sub)
    newP.update({k: t.detach() for k, t in out["new_bn_state"].items()})
    return dict(loss=float(ls["loss"]), mel_loss=float(ls["mel_loss"]), linear_loss=float(ls["linear_loss"]),
                loss_without_coeff=float(ls["loss_without_coeff"]), grad_norm=gn, lr=lr,
                grads=grads, params=newP, m=m, v=v, outputs=out)
