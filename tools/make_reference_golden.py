"""Generates tests/golden/ref_*.npz by executing the reference's OWN model code.

    python tools/make_reference_golden.py            # needs /root/reference (this container only)

The unmodified modules ``/root/reference/{hparams,models/*,text/symbols}.py`` are imported with ``oracle/tf1_shim`` on
``sys.path`` as ``tensorflow`` (TensorFlow 1.x itself is not installable here — see oracle/tf1_shim/tensorflow/_core.py for
what that does and does not pin).  For every case below: ``create_model(hparams).initialize(...)`` → ``add_loss()`` →
``add_optimizer(global_step)`` is driven exactly as ``train.py:143-155`` / ``synthesizer.py:48-55`` do, on OUR parameter
values (loaded through the TF variable names of ``tf_names.tf_to_ours``) and seeded inputs; outputs, losses, gradients,
the post-step parameters and batch-norm statistics are written as fixtures.  The fixtures are what
``tests/test_reference_golden.py`` (CPU oracle) and ``tests/test_gpu_parity.py`` (CUDA path) are checked against.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("TACO_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: (hparam overrides, num_speakers, batch kwargs, mode)
    "ref_train_single": (dict(), 1, dict(), "train"),
    "ref_infer_single": (dict(max_iters=6), 1, dict(), "infer"),
    "ref_train_deepvoice": (dict(model_type="deepvoice"), 3, dict(speakers=[2, 0]), "train"),
    "ref_infer_deepvoice": (dict(model_type="deepvoice", max_iters=4), 3, dict(speakers=[1, 2]), "infer"),
    "ref_train_simple": (dict(model_type="simple"), 4, dict(speakers=[3, 1]), "train"),
    "ref_train_deepvoice_table": (dict(model_type="deepvoice", speaker_embedding_size=1), 3, dict(speakers=[0, 2]), "train"),
    "ref_train_bah_norm": (dict(attention_type="bah_norm"), 1, dict(), "train"),
    "ref_train_bah": (dict(attention_type="bah"), 1, dict(), "train"),
    "ref_train_test_mode": (dict(), 1, dict(), "train_test_mode"),
    "ref_train_prioritize_lr1": (dict(prioritize_loss=True, decay_learning_rate_mode=1), 1, dict(), "train"),
    "ref_infer_manual_attention": (dict(max_iters=5), 1, dict(), "infer_manual"),
    "ref_train_ragged": (dict(), 1, dict(N=3, Ti=13, To=20, lengths=[13, 9, 5], seed=77), "train"),
    # two consecutive training steps (the train op run twice): Adam slots / beta powers, BN statistics and the global step
    # carry over in the variable store; the fixture holds the SECOND step's outputs, loss, gradients and the state after it
    "ref_train_two_steps": (dict(model_type="deepvoice"), 3, dict(speakers=[1, 2]), "train_x2"),
}


# BASELINE.json configs[0] ("single-speaker forward batch=2, 50-char text -> 200 mel frames, r=5 - the reference's own CPU
# case") at its exact size, forward (free-running, 40 decoder steps) and one training step.  Checked against the CPU oracle
# only (tests/test_reference_golden.py); the CUDA path meets the oracle at this size in test_forward_backward_vs_oracle.
# The 1025-bin linear outputs are stored every 8th bin to keep the fixtures small.
C1_CASES = {
    "ref_c1_infer": (dict(max_iters=40), 1, dict(N=2, Ti=50, To=200, lengths=[50, 37], seed=501), "infer"),
    "ref_c1_train": (dict(), 1, dict(N=2, Ti=50, To=200, lengths=[50, 37], seed=502), "train"),
}
C1_LINEAR_STRIDE = 8


def golden_batch(N=2, Ti=11, To=15, lengths=None, seed=2024, speakers=None):
    g = torch.Generator().manual_seed(seed)
    inp = torch.randint(2, 80, (N, Ti), generator=g, dtype=torch.int32)
    L = torch.tensor(lengths if lengths is not None else [Ti, Ti - 4][:N], dtype=torch.int32)
    for n in range(N):
        inp[n, L[n] - 1] = 1
        inp[n, L[n]:] = 0
    b = dict(inputs=inp, input_lengths=L, mel_targets=torch.rand(N, To, 80, generator=g),
             linear_targets=torch.rand(N, To, 1025, generator=g), loss_coeff=torch.rand(N, generator=g) + 0.5)
    if speakers is not None:
        b["speaker_id"] = torch.tensor(speakers, dtype=torch.int32)
    return b


def golden_params(tb, hp, num_speakers=1, seed=99):
    named = tb.params.init_params(hp, num_speakers, seed=seed, randomize_bn_state=True)
    g = torch.Generator().manual_seed(seed + 1)
    for k, v in named.items():
        if k.endswith("/gamma"):
            v.add_(torch.randn(v.shape, generator=g) * 0.2)
        elif k.endswith(("/beta", "/bias", "_bias", "score_bias")):
            v.add_(torch.randn(v.shape, generator=g) * 0.1)
    return named


def our_hparams(tb, overrides):
    return tb.hparams.override(reduction_factor=5, **overrides)


def manual_alignments(N, Td, Ti):
    """A diagonal-ish hand-made alignment, [N, Td, Ti] as AttentionWrapper indexes it (rnn_wrappers.py:315)."""
    a = torch.zeros(N, Td, Ti)
    for t in range(Td):
        a[:, t, min(Ti - 1, 2 * t)] = 0.75
        a[:, t, min(Ti - 1, 2 * t + 1)] += 0.25
    return a


def run_reference(tb, hp_over, num_speakers, batch, named, mode, double=False):
    """Drives the reference's model code (imported from REF) on `named` parameters; returns numpy results."""
    shim = os.path.join(ROOT, "oracle", "tf1_shim")
    for p in (REF, shim):
        if p not in sys.path:
            sys.path.insert(0, p)
    import warnings
    warnings.filterwarnings("ignore", category=SyntaxWarning)
    import tensorflow as tf                                   # the shim
    from hparams import hparams as ref_hp                     # reference/hparams.py (defaults are part of what is pinned)
    from models import create_model                           # reference/models/__init__.py:6
    import models.tacotron as ref_tacotron
    ref_tacotron.log = lambda *a, **k: None                   # the model logs its dimensions; keep the generator quiet

    tf.set_float_precision(double)
    ours = our_hparams(tb, hp_over)
    saved = dict(ref_hp.values())
    for k, v in dict(reduction_factor=5, **hp_over).items():
        ref_hp.set_hparam(k, v)
    # every model hyper-parameter must agree between reference/hparams.py (+ overrides) and our hparams mirror
    for k in ("embedding_size", "enc_prenet_sizes", "enc_bank_size", "enc_bank_channel_size", "enc_maxpool_width", "enc_highway_depth",
              "enc_rnn_size", "enc_proj_sizes", "enc_proj_width", "attention_type", "attention_size", "attention_state_size",
              "dec_layer_num", "dec_rnn_size", "dec_prenet_sizes", "post_bank_size", "post_bank_channel_size", "post_maxpool_width",
              "post_highway_depth", "post_rnn_size", "post_proj_sizes", "post_proj_width", "num_mels", "num_freq", "speaker_embedding_size",
              "initial_learning_rate", "adam_beta1", "adam_beta2", "sample_rate", "max_iters"):
        assert getattr(ref_hp, k) == getattr(ours, k), (k, getattr(ref_hp, k), getattr(ours, k))

    name_map = tb.tf_names.tf_to_ours(ours, num_speakers, prefix="model/inference/")
    used = set()
    ft = torch.float64 if double else torch.float32

    def provider(full, shape, initializer):
        if full not in name_map:
            if full.startswith("model/inference/"):
                raise KeyError("the reference created a variable we have no parameter for: " + full)
            return None
        used.add(full)
        return named[name_map[full]].to(ft).reshape(shape)

    feeds = {}
    is_train = mode.startswith("train")
    Td = batch["mel_targets"].shape[1] // 5 if is_train else ref_hp.max_iters
    if mode == "infer_manual":
        feeds = {"is_manual_attention": True,
                 "manual_alignments": manual_alignments(batch["inputs"].shape[0], Td, batch["inputs"].shape[1]).to(ft)}
    tf.reset_default_graph(provider=provider, feeds=feeds)
    T = lambda x: tf.Tensor(x.to(ft) if x.dtype.is_floating_point else x)  # noqa: E731
    res = {}
    try:
        global_step = tf.Variable(0, name="global_step", trainable=False)          # train.py:143

        def build():
            with tf.variable_scope("model"):                                        # train.py:145 / synthesizer.py:49
                model = create_model(ref_hp)
                spk = T(batch["speaker_id"]) if num_speakers > 1 else None
                if is_train:
                    model.initialize(T(batch["inputs"]), T(batch["input_lengths"]), num_speakers, spk,
                                     T(batch["mel_targets"]), T(batch["linear_targets"]), T(batch["loss_coeff"]),
                                     rnn_decoder_test_mode=(mode == "train_test_mode"), is_randomly_initialized=True)
                    model.add_loss()
                    model.add_optimizer(global_step)
                else:
                    model.initialize(T(batch["inputs"]), T(batch["input_lengths"]), num_speakers, spk)
            return model
        model = build()
        if mode == "train_x2":
            # "sess.run(train_op)" a second time: the eager stand-in evaluates while it builds, so the graph is built again over
            # the SAME variable store (parameters, Adam slots, beta powers, BN statistics, global step all persist); only the
            # naming counters and the already-executed update ops of the first trace are dropped
            st0 = tf.shim_state()
            st0.opened.clear(); st0.collections.clear()
            model = build()
        missing = set(name_map) - used
        assert not missing, "parameters the reference never created: %s" % sorted(missing)
        res.update(mel_outputs=model.mel_outputs.numpy(), linear_outputs=model.linear_outputs.numpy(),
                   alignments=model.alignments.numpy())
        if is_train:
            tvars = tf.trainable_variables()
            grads = {name_map[v.name]: (np.zeros(tuple(v.t.shape)) if g is None else g.numpy()) for g, v in zip(model.gradients, tvars)}
            res["scalars"] = np.array([float(model.loss), float(model.mel_loss), float(model.linear_loss), float(model.loss_without_coeff),
                                       float(np.sqrt(sum((g.astype(np.float64) ** 2).sum() for g in grads.values()))),
                                       float(model.learning_rate)])
            res["grads"] = grads
            # `optimize` already ran (eager): parameters and BN statistics below are the post-step values
            st = tf.shim_state().vars
            res["params_after"] = {name_map[k]: v.numpy() for k, v in st.items() if k in name_map}
            res["global_step_after"] = int(global_step)
    finally:
        for k, v in saved.items():
            ref_hp.set_hparam(k, v)
        tf.set_float_precision(False)
    return res


FULL_SMALL = ("attention/v", "mel_proj/bias", "embedding")
FULL_MORE = FULL_SMALL + ("attention/query_kernel", "enc_cbhg/gru_bw/cand_kernel", "post_cbhg/bank_3/kernel", "dec_gru_2/gates_kernel")


def pack(res, full_grads=FULL_SMALL):
    out = dict(mel_outputs=res["mel_outputs"], linear_outputs=res["linear_outputs"], alignments=res["alignments"])
    if "scalars" in res:
        names = sorted(res["grads"])
        out.update(scalars=res["scalars"], grad_names=np.array(names),
                   grad_norms=np.array([float(np.linalg.norm(res["grads"][k].astype(np.float64))) for k in names]),
                   global_step_after=res["global_step_after"])
        pa = res["params_after"]
        for k in full_grads:
            if k in res["grads"]:
                out["grad:" + k] = res["grads"][k]
                out["after:" + k] = pa[k]
        for k in ("enc_cbhg/proj_1/moving_mean", "enc_cbhg/proj_1/moving_var", "post_cbhg/bank_8/moving_mean", "post_cbhg/proj_2/moving_var"):
            out["after:" + k] = pa[k]
        for k in pa:
            if k.startswith("speaker"):
                out["grad:" + k] = res["grads"][k] if k in res["grads"] else np.zeros_like(pa[k])
    return out


def audio_signal(n=14400, seed=11):
    """A speech-like test signal: a few drifting harmonics under a slow envelope plus a little noise."""
    rng = np.random.RandomState(seed)
    t = np.arange(n) / 24000.0
    f0 = 140.0 + 30.0 * np.sin(2 * np.pi * 1.5 * t)
    ph = 2 * np.pi * np.cumsum(f0) / 24000.0
    y = sum(a * np.sin(h * ph + rng.rand() * 6.28) for h, a in ((1, 0.5), (2, 0.3), (3, 0.2), (5, 0.1), (9, 0.05), (17, 0.02)))
    y = y * (0.3 + 0.7 * np.sin(2 * np.pi * 2.0 * t) ** 2) + 0.01 * rng.randn(n)
    return (0.5 * y / np.abs(y).max()).astype(np.float32)


def run_reference_audio(gl_iters=4, seed=7):
    """The reference's audio/__init__.py, unmodified, over the librosa 0.5.1 stand-in (oracle/tf1_shim/librosa) and the real
    scipy.  ``np.complex`` (removed from numpy >= 1.24, used at audio/__init__.py:78) is re-attached for the import."""
    shim = os.path.join(ROOT, "oracle", "tf1_shim")
    for p in (REF, shim):
        if p not in sys.path:
            sys.path.insert(0, p)
    import warnings
    warnings.filterwarnings("ignore", category=SyntaxWarning)
    if not hasattr(np, "complex"):
        np.complex = complex
    import audio as ref_audio                                   # reference/audio/__init__.py
    y = audio_signal()
    spec = ref_audio.spectrogram(y)                             # [num_freq, T]
    mel = ref_audio.melspectrogram(y)                           # [num_mels, T]
    saved = ref_audio.hparams.griffin_lim_iters
    ref_audio.hparams.set_hparam("griffin_lim_iters", gl_iters)
    try:
        np.random.seed(seed)
        wav = ref_audio.inv_spectrogram(spec)                   # draws np.random.rand(*S.shape) once (audio/__init__.py:77)
    finally:
        ref_audio.hparams.set_hparam("griffin_lim_iters", saved)
    np.random.seed(seed)
    phase = np.random.rand(*spec.shape)
    return dict(wav_in=y, spectrogram=spec.astype(np.float32), melspectrogram=mel.astype(np.float32), phase=phase.astype(np.float32),
                wav_out=wav.astype(np.float32), n_iters=gl_iters, mel_basis=ref_audio._build_mel_basis().astype(np.float32))


TEXT_SAMPLES = ["안녕하세요, 반갑습니다.", "오늘 날씨가 참 좋네요!", "값이 얼마예요? 읽다, 앉아; 닭: 삶 (괜찮아) - 끝.", "  앞뒤 공백  ",
                "그리고 그는 천천히 걸어갔다.", "뭐라고요? 아니, 괜찮아요... 정말로!"]


def run_reference_text():
    """The reference's text front end (text/__init__.py text_to_sequence / sequence_to_text), unmodified, over the `jamo`
    stand-in, on sentences that need no normalisation (plain Hangul + the punctuation of the symbol table)."""
    shim = os.path.join(ROOT, "oracle", "tf1_shim")
    for p in (REF, shim):
        if p not in sys.path:
            sys.path.insert(0, p)
    import warnings
    warnings.filterwarnings("ignore", category=SyntaxWarning)
    import tensorflow  # noqa: F401  (hparams.py, imported by text/__init__.py, needs the name)
    from text import text_to_sequence, sequence_to_text
    from text.symbols import symbols
    out = {"symbols": symbols, "cases": []}
    for s in TEXT_SAMPLES:
        seq = [int(v) for v in text_to_sequence(s)]
        out["cases"].append({"text": s, "sequence": seq, "round_trip": sequence_to_text(seq, skip_eos_and_pad=True, combine_jamo=True)})
    return out


def batch_examples(seed=5, n=5, n_mel=4, n_lin=6):
    """Examples as DataFeeder hands them to _prepare_batch: (tokens, loss_coeff, mel[T, n_mel], linear[T, n_lin], speaker, T)."""
    rng = np.random.RandomState(seed)
    out = []
    for i in range(n):
        T, L = int(rng.randint(7, 23)), int(rng.randint(3, 12))
        tok = rng.randint(2, 80, L).astype(np.int32); tok[-1] = 1
        out.append((tok, float(rng.choice([1.0, 0.2])), rng.rand(T, n_mel).astype(np.float32), rng.rand(T, n_lin).astype(np.float32),
                    int(rng.randint(0, 3)), T))
    return out


def run_reference_batch(reduction_factor=5):
    """The reference's datasets/datafeeder.py `_prepare_batch` (and its padding helpers, :289-328), unmodified, imported over
    the stand-ins (tensorflow / librosa / jamo / tinytag are only touched at import time on this path)."""
    shim = os.path.join(ROOT, "oracle", "tf1_shim")
    for p in (REF, shim):
        if p not in sys.path:
            sys.path.insert(0, p)
    import warnings
    warnings.filterwarnings("ignore", category=SyntaxWarning)
    if not hasattr(np, "complex"):
        np.complex = complex
    import datasets.datafeeder as ref_df
    out = {}
    for tag, data_type in (("plain", None), ("train", "train")):
        res = ref_df._prepare_batch(batch_examples(), reduction_factor, np.random.RandomState(11), data_type)
        for name, arr in zip(("inputs", "input_lengths", "loss_coeff", "mel_targets", "linear_targets", "speaker_id"), res):
            out["%s:%s" % (tag, name)] = np.asarray(arr)
    out["round_up"] = np.array([ref_df._round_up(x, 5) for x in range(0, 13)])
    return out


def trim_alignments(Ti=12, Td=14):
    """Attention matrices [T_in, T_dec] covering the branches of the end-trimming rule (synthesizer.py:242-262)."""
    rng = np.random.RandomState(4)
    def diag(stop, hold):                      # attention walks forward, then sits on position `stop` for `hold` steps
        a = np.full((Ti, Td), 0.01)
        pos = [min(stop, t) for t in range(Td)]
        for t in range(Td):
            a[pos[t] if t < stop + hold else min(Ti - 1, stop + 1 + (t - stop - hold) // 2), t] = 1.0
        return a
    return [diag(8, 2), diag(8, 7), diag(Ti - 1, 3), diag(5, 1), rng.rand(Ti, Td), np.eye(Ti, Td) + 0.001]


def run_reference_trim(r=5):
    """The reference's synthesizer.plot_graph_and_save_audio (attention_trim=True, end_of_sentence=True), unmodified, over
    the stand-ins; its `save_audio` is replaced by a recorder, so the number of spectrogram frames that survive the trim can
    be read off the waveform length (hop * (frames - 1))."""
    shim = os.path.join(ROOT, "oracle", "tf1_shim")
    for p in (REF, shim):
        if p not in sys.path:
            sys.path.insert(0, p)
    import warnings
    warnings.filterwarnings("ignore", category=SyntaxWarning)
    if not hasattr(np, "complex"):
        np.complex = complex
    import synthesizer as ref_syn
    from hparams import hparams as ref_hp
    saved = (ref_hp.reduction_factor, ref_hp.griffin_lim_iters)
    ref_hp.set_hparam("reduction_factor", r); ref_hp.set_hparam("griffin_lim_iters", 0)
    got = []
    ref_syn.save_audio = lambda audio, path, sample_rate=None: got.append(len(audio))
    out = []
    try:
        rng = np.random.RandomState(9)
        for k, al in enumerate(trim_alignments()):
            Td = al.shape[1]
            for seq_len in (al.shape[0], al.shape[0] - 3):
                spec = rng.rand(Td * r, ref_hp.num_freq).astype(np.float32)
                ref_syn.plot_graph_and_save_audio((0, (spec, al, None, "text", list(range(seq_len)))), attention_trim=True,
                                                  end_of_sentence=True)
                out.append({"alignment": k, "sequence_len": seq_len, "frames": got[-1] // 300 + 1})
    finally:
        ref_hp.set_hparam("reduction_factor", saved[0]); ref_hp.set_hparam("griffin_lim_iters", saved[1])
    return {"r": r, "cases": out}


def main():
    sys.path.insert(0, ROOT)
    import tacotron_b200 as tb
    os.makedirs(OUT, exist_ok=True)
    for name, (over, S, bk, mode) in CASES.items():
        hp = our_hparams(tb, over)
        named = golden_params(tb, hp, S)
        b = golden_batch(**bk)
        res = run_reference(tb, over, S, b, named, mode)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **pack(res, FULL_MORE if name == "ref_train_single" else FULL_SMALL))
        print(name, "mel", res["mel_outputs"].shape, "loss" if "scalars" in res else "", res.get("scalars", [""])[0])
    for name, (over, S, bk, mode) in C1_CASES.items():
        hp = our_hparams(tb, over)
        res = run_reference(tb, over, S, golden_batch(**bk), golden_params(tb, hp, S), mode)
        packed = pack(res, ("attention/v", "mel_proj/bias"))
        packed["linear_outputs"] = np.ascontiguousarray(packed["linear_outputs"][:, :, ::C1_LINEAR_STRIDE])
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **packed)
        print(name, "mel", res["mel_outputs"].shape, res.get("scalars", [""])[0])
    import json
    with open(os.path.join(OUT, "ref_text_small.json"), "w", encoding="utf-8") as f:
        json.dump(run_reference_text(), f, ensure_ascii=False, indent=1)
    print("ref_text_small", len(TEXT_SAMPLES), "sentences")
    with open(os.path.join(OUT, "ref_trim_small.json"), "w") as f:
        json.dump(run_reference_trim(), f, indent=1)
    print("ref_trim_small")
    np.savez_compressed(os.path.join(OUT, "ref_batch_small.npz"), **run_reference_batch())
    print("ref_batch_small")
    a = run_reference_audio()
    np.savez_compressed(os.path.join(OUT, "ref_audio_small.npz"), **a)
    print("ref_audio_small", a["spectrogram"].shape, a["melspectrogram"].shape, a["wav_out"].shape)


if __name__ == "__main__":
    main()
