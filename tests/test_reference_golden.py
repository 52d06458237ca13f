"""The CPU oracle against fixtures produced by the reference's OWN model code (tests/golden/ref_*.npz).

The fixtures come from ``tools/make_reference_golden.py``: /root/reference's unmodified ``models/*.py`` + ``hparams.py``
executed over ``oracle/tf1_shim`` (an eager stand-in for the TF r1.4 API — TensorFlow itself is not installable here).
This is what pins ``oracle/tacotron_oracle.py``: every output, loss, gradient, Adam update and batch-norm statistic the
reference's code produces on seeded inputs must be reproduced by the oracle restatement.

Tolerances (fp32 on both sides, different summation orders): outputs 5e-6 max-abs, scalars 1e-6 (global gradient norm 2e-5), per-tensor gradient
norms 1e-4 relative (with a floor of 1e-6 of the global norm: gradients that are mathematically zero or sums of nearly
cancelling terms, such as a conv bias in front of batch norm, are rounding noise), full
gradient tensors 2e-5 relative L2, post-step parameters / BN statistics 2e-6 max-abs.
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import make_reference_golden as mr  # noqa: E402
from oracle import tacotron_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
HAVE_REF = os.path.isdir(os.path.join(mr.REF, "models"))


def _oracle_case(tb, name):
    over, S, bk, mode = (mr.CASES.get(name) or mr.C1_CASES[name])
    hp = mr.our_hparams(tb, over)
    named = mr.golden_params(tb, hp, S)
    b = mr.golden_batch(**bk)
    if not mode.startswith("train"):
        ma = mr.manual_alignments(b["inputs"].shape[0], hp.max_iters, b["inputs"].shape[1]) if mode == "infer_manual" else None
        with torch.no_grad():
            out = O.forward(named, hp, b["inputs"], b["input_lengths"], S, b.get("speaker_id"), manual_alignments=ma, max_iters=hp.max_iters)
        return dict(outputs=out)
    names = [k for k in named if not k.endswith(("moving_mean", "moving_var"))]
    m = {k: torch.zeros_like(named[k]) for k in names}
    v = {k: torch.zeros_like(named[k]) for k in names}
    step = 0
    if mode == "train_x2":                       # the fixture describes the SECOND of two consecutive steps
        first = O.train_step({k: t.clone() for k, t in named.items()}, m, v, hp, b, 0, True, S)
        named, m, v, step = first["params"], first["m"], first["v"], 1
    leaf = {k: (named[k].clone().requires_grad_(True) if k in names else named[k]) for k in named}
    out = O.forward(leaf, hp, b["inputs"], b["input_lengths"], S, b.get("speaker_id"), b["mel_targets"], b["linear_targets"],
                    rnn_decoder_test_mode=(mode == "train_test_mode"))
    ls = O.losses(out, b["mel_targets"], b["linear_targets"], b["loss_coeff"], hp)
    gl = torch.autograd.grad(ls["loss"], [leaf[k] for k in names], allow_unused=True)
    grads = {k: (g if g is not None else torch.zeros_like(named[k])) for k, g in zip(names, gl)}
    clipped, gn = O.clip_by_global_norm(grads, 1.0)
    lr = O.learning_rate(hp, step, True)
    P = {k: named[k].detach().clone() for k in names}
    P, m, v = O.adam_step(P, clipped, m, v, step + 1, lr, hp.adam_beta1, hp.adam_beta2)
    after = dict(P)
    after.update({k: t.detach() for k, t in out["new_bn_state"].items()})
    return dict(outputs=out, losses={k: float(x) for k, x in ls.items()}, grads=grads, grad_norm=gn, lr=lr, after=after, steps=step + 1)


@pytest.mark.parametrize("name", sorted(mr.CASES) + sorted(mr.C1_CASES))
def test_oracle_reproduces_reference_run(tb, name):
    """(ref_c1_*: BASELINE.json configs[0] at its exact size — batch 2, 50 tokens, 200 mel frames, r=5; linear bins strided.)"""
    g = np.load(os.path.join(GOLD, name + ".npz"))
    res = _oracle_case(tb, name)
    for k in ("mel_outputs", "linear_outputs", "alignments"):
        got = res["outputs"][k].detach().numpy()
        if k == "linear_outputs" and name in mr.C1_CASES:
            got = got[:, :, ::mr.C1_LINEAR_STRIDE]
        assert got.shape == g[k].shape, (k, got.shape, g[k].shape)
        assert np.abs(got - g[k]).max() <= 5e-6, (k, np.abs(got - g[k]).max())
    if "scalars" not in g.files:
        return
    ls = res["losses"]
    want = g["scalars"]
    got = [ls["loss"], ls["mel_loss"], ls["linear_loss"], ls["loss_without_coeff"], res["grad_norm"], res["lr"]]
    for a, b_, what in zip(got, want, ("loss", "mel_loss", "linear_loss", "loss_without_coeff", "grad_norm", "lr")):
        rel = 2e-5 if what == "grad_norm" else 1e-6          # the norm sums 9 M fp32 gradient entries computed in a different order
        assert abs(a - b_) <= rel * max(1.0, abs(b_)), (what, a, b_)
    assert sorted(res["grads"]) == list(g["grad_names"])                  # the reference trains exactly our parameter set
    for k, n in zip(g["grad_names"], g["grad_norms"]):
        mine = float(res["grads"][k].double().norm())
        rel_n = 5e-4 if name in mr.C1_CASES else 1e-4        # 200-frame reductions in front of batch norm cancel more
        assert abs(mine - n) <= max(rel_n * n, 1e-6 * float(want[4])), (k, mine, n)     # floor: 1e-6 of the global gradient norm
    for key in g.files:
        if key.startswith("grad:"):
            ref = g[key]
            d = np.linalg.norm(res["grads"][key[5:]].numpy() - ref)
            assert d <= (5e-4 if name in mr.C1_CASES else 2e-5) * max(np.linalg.norm(ref), 1e-3), (key, d)
        elif key.startswith("after:"):
            assert np.abs(res["after"][key[6:]].numpy() - g[key]).max() <= 2e-6, key
    assert int(g["global_step_after"]) == res["steps"]


def test_existing_oracle_fixture_equals_reference_run(tb):
    """tests/golden/tacotron_train_small.npz (written from the oracle in an earlier round) and ref_train_single.npz (written
    by the reference's code) describe the same parameters and batch: they must agree."""
    a = np.load(os.path.join(GOLD, "tacotron_train_small.npz"))
    b = np.load(os.path.join(GOLD, "ref_train_single.npz"))
    for k in ("mel_outputs", "linear_outputs", "alignments"):
        assert np.abs(a[k] - b[k]).max() <= 5e-6, k
    assert np.abs(a["scalars"] - b["scalars"]).max() <= 1e-6
    assert np.abs(a["grad_attention_v"] - b["grad:attention/v"]).max() <= 1e-7
    assert np.abs(a["param_after_attention_v"] - b["after:attention/v"]).max() <= 1e-6
    assert np.abs(a["bn_after_enc_p1_mean"] - b["after:enc_cbhg/proj_1/moving_mean"]).max() <= 1e-6
    c = np.load(os.path.join(GOLD, "tacotron_infer_small.npz"))
    d = np.load(os.path.join(GOLD, "ref_infer_single.npz"))
    for k in ("mel_outputs", "linear_outputs", "alignments"):
        assert np.abs(c[k] - d[k]).max() <= 5e-6, k


def test_tf_variable_name_table_covers_every_parameter(tb):
    """tf_names.tf_to_ours is a bijection onto params.param_specs for every speaker mode / attention type."""
    for over, S in ((dict(), 1), (dict(model_type="deepvoice"), 3), (dict(model_type="simple"), 2),
                    (dict(model_type="deepvoice", speaker_embedding_size=1), 3), (dict(attention_type="bah_norm"), 1),
                    (dict(attention_type="bah"), 1)):
        hp = mr.our_hparams(tb, over)
        table = tb.tf_names.tf_to_ours(hp, S)
        ours = [s.name for s in tb.params.param_specs(hp, S)]
        assert sorted(table.values()) == sorted(ours), over
        assert len(set(table.values())) == len(table)
        assert all(k.startswith("model/inference/") for k in table)


@pytest.mark.skipif(not HAVE_REF, reason="the reference tree is only mounted in the build container")
@pytest.mark.parametrize("name", ["ref_train_deepvoice", "ref_infer_manual_attention"])
def test_fixtures_regenerate_from_the_reference(tb, name):
    """The committed fixtures are reproducible: re-running the reference's code here gives the stored arrays."""
    over, S, bk, mode = mr.CASES[name]
    hp = mr.our_hparams(tb, over)
    res = mr.run_reference(tb, over, S, mr.golden_batch(**bk), mr.golden_params(tb, hp, S), mode)
    g = np.load(os.path.join(GOLD, name + ".npz"))
    for k in ("mel_outputs", "linear_outputs", "alignments"):
        assert np.array_equal(res[k], g[k]), k
    if "scalars" in g.files:
        assert np.array_equal(res["scalars"], g["scalars"])


@pytest.mark.skipif(not HAVE_REF, reason="the reference tree is only mounted in the build container")
@pytest.mark.parametrize("case", ["ref_train_deepvoice", "ref_c1_train"])
def test_reference_run_in_float64_matches_oracle_in_float64(tb, case):
    """Rounding-free structural check: the reference's code over the shim in float64 against the oracle in float64 (also at
    the BASELINE configs[0] size, where the fp32 fixtures differ from the fp32 oracle by up to 1e-4 in some gradients)."""
    over, S, bk, mode = (mr.CASES.get(case) or mr.C1_CASES[case])
    hp = mr.our_hparams(tb, over)
    named = mr.golden_params(tb, hp, S)
    b = mr.golden_batch(**bk)
    res = mr.run_reference(tb, over, S, b, named, mode, double=True)
    P = {k: v.double() for k, v in named.items()}
    names = [k for k in P if not k.endswith(("moving_mean", "moving_var"))]
    leaf = {k: (P[k].clone().requires_grad_(True) if k in names else P[k]) for k in P}
    out = O.forward(leaf, hp, b["inputs"], b["input_lengths"], S, b.get("speaker_id"), b["mel_targets"].double(), b["linear_targets"].double())
    ls = O.losses(out, b["mel_targets"].double(), b["linear_targets"].double(), b["loss_coeff"].double(), hp)
    gl = torch.autograd.grad(ls["loss"], [leaf[k] for k in names], allow_unused=True)
    for k in ("mel_outputs", "linear_outputs", "alignments"):
        assert np.abs(out[k].detach().numpy() - res[k]).max() <= 1e-12, k
    assert abs(float(ls["loss"]) - res["scalars"][0]) <= 1e-13
    for k, gg in zip(names, gl):
        ref = res["grads"][k]
        mine = np.zeros_like(ref) if gg is None else gg.numpy()
        assert np.abs(mine - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max()), k


# ---- audio: the reference's audio/__init__.py over the librosa stand-in + real scipy (ref_audio_small.npz) ----------
def test_audio_oracle_reproduces_reference_audio_code(tb):
    from oracle import griffin_lim_oracle as G
    from importlib import import_module
    a = np.load(os.path.join(GOLD, "ref_audio_small.npz"))
    assert a["spectrogram"].shape == (1025, 1 + 14400 // 300) and a["melspectrogram"].shape == (80, 49)
    assert np.abs(G.spectrogram(a["wav_in"]) - a["spectrogram"]).max() <= 2e-5
    assert np.abs(G.melspectrogram(a["wav_in"]) - a["melspectrogram"]).max() <= 2e-6
    w = G.inv_spectrogram(a["spectrogram"].T, a["phase"].T, n_iters=int(a["n_iters"]))
    assert w.shape == a["wav_out"].shape and np.abs(w - a["wav_out"]).max() <= 5e-6 * np.abs(a["wav_out"]).max()
    # the product's host-side mel table (a constant; no oracle import there) against the reference's _build_mel_basis()
    audio = import_module("multi-speaker-tacotron-tensorflow_b200.audio")
    B = audio.build_mel_basis(tb.hparams)
    assert B.shape == a["mel_basis"].shape and np.abs(B - a["mel_basis"]).max() <= 1e-8
    assert np.abs(G.slaney_mel_basis() - a["mel_basis"]).max() <= 1e-8


@pytest.mark.skipif(not HAVE_REF, reason="the reference tree is only mounted in the build container")
def test_audio_fixture_regenerates_from_the_reference():
    a = np.load(os.path.join(GOLD, "ref_audio_small.npz"))
    b = mr.run_reference_audio()
    for k in ("spectrogram", "melspectrogram", "wav_out", "phase"):
        assert np.array_equal(a[k], b[k]), k


# ---- text: the reference's text front end over the `jamo` stand-in (ref_text_small.json) -----------------------------
def test_tokenizer_matches_reference_text_front_end(tb):
    import json
    with open(os.path.join(GOLD, "ref_text_small.json"), encoding="utf-8") as f:
        ref = json.load(f)
    T = tb.text
    assert T.symbols == ref["symbols"] and len(T.symbols) == tb.params.NUM_SYMBOLS == 80
    assert T.symbols[tb.params.PAD_ID] == "_" and T.symbols[tb.params.EOS_ID] == "~"
    for case in ref["cases"]:
        seq = T.text_to_sequence(case["text"])
        assert seq.dtype == np.int32 and list(seq) == case["sequence"], case["text"]
        assert seq[-1] == tb.params.EOS_ID and (seq[:-1] > 1).all()
        assert T.sequence_to_text(seq, skip_eos_and_pad=True, combine_jamo=True) == case["round_trip"]
    batch, L = T.texts_to_batch([c["text"] for c in ref["cases"][:3]])
    assert batch.shape == (3, max(len(c["sequence"]) for c in ref["cases"][:3])) and list(L) == [len(c["sequence"]) for c in ref["cases"][:3]]
    assert (batch[0, L[0]:] == tb.params.PAD_ID).all()
    # symbols outside the table (digits, Latin letters: the reference would read them out, see text.py) are dropped
    assert list(T.text_to_sequence("A1 가")) == [T._symbol_to_id[" "], T._symbol_to_id["ᄀ"], T._symbol_to_id["ᅡ"], 1]


@pytest.mark.skipif(not HAVE_REF, reason="the reference tree is only mounted in the build container")
def test_text_fixture_regenerates_from_the_reference():
    import json
    with open(os.path.join(GOLD, "ref_text_small.json"), encoding="utf-8") as f:
        assert json.load(f) == mr.run_reference_text()


# ---- input pipeline: the reference's _prepare_batch (datasets/datafeeder.py:289-328) ---------------------------------
def test_prepare_batch_matches_reference_datafeeder(tb):
    from importlib import import_module
    df = import_module("multi-speaker-tacotron-tensorflow_b200.datasets.datafeeder")
    g = np.load(os.path.join(GOLD, "ref_batch_small.npz"))
    for tag, data_type in (("plain", None), ("train", "train")):
        feed = df.prepare_batch(mr.batch_examples(), 5, np.random.RandomState(11), data_type)
        for name in ("inputs", "input_lengths", "loss_coeff", "mel_targets", "linear_targets", "speaker_id"):
            ref = g["%s:%s" % (tag, name)]
            assert feed[name].shape == ref.shape and feed[name].dtype == ref.dtype, (tag, name, feed[name].shape, ref.shape)
            assert np.array_equal(feed[name], ref), (tag, name)
        assert feed["mel_targets"].shape[1] % 5 == 0 and feed["mel_targets"].shape[1] > max(x[5] for x in mr.batch_examples())
    assert [df._round_up(x, 5) for x in range(0, 13)] == list(g["round_up"])
    # written in place into staging buffers (the pinned ring of the DataFeeder): same values, views into the buffers
    big = {k: np.full(v.size + 7, -1, v.dtype) for k, v in feed.items()}
    feed2 = df.prepare_batch(mr.batch_examples(), 5, np.random.RandomState(11), "train", out=big)
    for k in feed:
        assert np.array_equal(feed2[k], feed[k]) and np.shares_memory(feed2[k], big[k]), k


@pytest.mark.skipif(not HAVE_REF, reason="the reference tree is only mounted in the build container")
def test_batch_fixture_regenerates_from_the_reference():
    g = np.load(os.path.join(GOLD, "ref_batch_small.npz"))
    again = mr.run_reference_batch()
    assert sorted(g.files) == sorted(again) and all(np.array_equal(g[k], again[k]) for k in g.files)


# ---- hyper-parameters: defaults and the params.json contract (hparams.py, utils/__init__.py:100-126) -------------------
@pytest.mark.skipif(not HAVE_REF, reason="the reference tree is only mounted in the build container")
def test_hparams_defaults_and_params_json_interoperate_with_the_reference(tb, tmp_path):
    shim = os.path.join(ROOT, "oracle", "tf1_shim")
    for p in (mr.REF, shim):
        if p not in sys.path:
            sys.path.insert(0, p)
    import warnings
    warnings.filterwarnings("ignore", category=SyntaxWarning)
    import tensorflow  # noqa: F401  (the stand-in; reference/hparams.py builds a tf.contrib.training.HParams)
    from hparams import hparams as ref_hp
    import utils as ref_utils
    ref_values = ref_hp.values()
    mine = tb.hparams.values()
    assert sorted(ref_values) == sorted(mine), set(ref_values) ^ set(mine)           # same keys ...
    for k, v in ref_values.items():
        assert mine[k] == v, (k, mine[k], v)                                         # ... same defaults (active override block included)
    # the reference writes params.json, we read it
    d1 = str(tmp_path / "ref_run")
    os.makedirs(d1)
    ref_hp.set_hparam("batch_size", 24); ref_hp.set_hparam("model_type", "deepvoice")
    try:
        ref_utils.save_hparams(d1, ref_hp)
    finally:
        ref_hp.set_hparam("batch_size", ref_values["batch_size"]); ref_hp.set_hparam("model_type", ref_values["model_type"])
    hp = tb.hparams.override()
    tb.load_hparams(hp, d1)
    assert hp.batch_size == 24 and hp.model_type == "deepvoice" and hp.enc_prenet_sizes == ref_values["enc_prenet_sizes"]
    # we write params.json, the reference reads it
    d2 = str(tmp_path / "our_run")
    tb.save_hparams(d2, tb.hparams.override(reduction_factor=5, attention_type="bah_norm"))
    try:
        ref_utils.load_hparams(ref_hp, d2)
        assert ref_hp.reduction_factor == 5 and ref_hp.attention_type == "bah_norm"
    finally:
        for k, v in ref_values.items():
            ref_hp.set_hparam(k, v)


# ---- synthesis post-processing: the end-trimming rule of synthesizer.py:242-262 ---------------------------------------
def test_attention_trim_matches_reference_synthesizer(tb):
    import json
    from importlib import import_module
    syn = import_module("multi-speaker-tacotron-tensorflow_b200.synthesizer")
    with open(os.path.join(GOLD, "ref_trim_small.json")) as f:
        ref = json.load(f)
    als = mr.trim_alignments()
    assert len(ref["cases"]) == 2 * len(als)
    for c in ref["cases"]:
        al = als[c["alignment"]]
        keep = syn.attention_trim_frames(al, c["sequence_len"], ref["r"])
        assert min(keep, al.shape[1] * ref["r"]) == c["frames"], c        # spec[:keep] as the reference slices wav[:spec_end_idx]


@pytest.mark.skipif(not HAVE_REF, reason="the reference tree is only mounted in the build container")
def test_trim_fixture_regenerates_from_the_reference():
    import json
    with open(os.path.join(GOLD, "ref_trim_small.json")) as f:
        assert json.load(f) == mr.run_reference_trim()
