"""CPU-side checks: hparams mirror, parameter layout, the C-ABI library loads and exports every declared symbol (no
compute calls without a GPU), golden fixtures are reproducible from the oracle, and the data-parallel plumbing (gloo)."""
import ctypes
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_hparams_defaults_and_roundtrip(tb, tmp_path):
    hp = tb.hparams
    # effective defaults of the reference after its own override blocks (hparams.py:26-29,83-94)
    assert hp.sample_rate == 24000 and hp.post_rnn_size == 256 and hp.dropout_prob == 0.8 and hp.reduction_factor == 4
    assert hp.attention_type == "bah_mon" and hp.model_type == "single" and hp.max_iters == 200
    h2 = hp.override(reduction_factor=5)
    assert h2.reduction_factor == 5 and hp.reduction_factor == 4
    with pytest.raises(KeyError):
        hp.override(not_a_key=1)
    p = tb.save_hparams(str(tmp_path), h2)
    assert json.load(open(p))["reduction_factor"] == 5
    h3 = tb.load_hparams(tb.hparams.override(), str(tmp_path))
    assert h3.reduction_factor == 5
    assert "reduction_factor: 5" in tb.hparams_debug_string(h3)
    from importlib import import_module
    hpm = import_module("multi-speaker-tacotron-tensorflow_b200.hparams")
    assert hpm.stft_parameters(hp) == (2048, 300, 1200)          # audio/__init__.py:118-122 at 24 kHz


def test_layout_alignment_and_bank_contiguity(tb, hp5):
    specs = tb.params.param_specs(hp5, 1)
    lay = tb.params.make_layout(specs)
    assert all(o % 8 == 0 for o in lay.offsets.values())      # 16 bytes in the bf16 mirror of the flat buffer
    for pf, Kb, Cb in (("enc_cbhg", 16, 128), ("post_cbhg", 8, 256)):
        for f in ("bias", "gamma", "beta", "moving_mean", "moving_var"):
            offs = [lay.offsets["%s/bank_%d/%s" % (pf, k, f)] for k in range(1, Kb + 1)]
            assert all(b - a == Cb for a, b in zip(offs, offs[1:]))
    named = tb.params.init_params(hp5, 1, seed=3)
    flat, state = tb.params.flatten(named, lay)
    views = tb.params.views(flat, state, lay)
    assert all(torch.equal(views[k], named[k]) for k in named)
    assert float(named["enc_cbhg/highway_1/T_bias"][0]) == -1.0 and float(named["dec_gru_1/gates_bias"][0]) == 1.0
    assert named["embedding"].abs().max() <= 1.0 + 1e-6          # truncated normal, sigma 0.5, cut at 2 sigma


def test_speaker_modes(tb, hp5):
    P = tb.params
    assert P.speaker_mode(hp5, 1) == "none"
    assert P.speaker_mode(hp5.override(model_type="simple"), 3) == "simple"
    assert P.speaker_mode(hp5.override(model_type="deepvoice"), 3) == "deepvoice"
    assert P.speaker_mode(hp5.override(model_type="deepvoice", speaker_embedding_size=1), 3) == "deepvoice_table"
    with pytest.raises(ValueError, match="Unkown multi-speaker model type"):
        P.speaker_mode(hp5, 2)                                    # model_type='single' with >1 speakers (tacotron.py:87-88)
    n_simple = sum(s.numel for s in P.param_specs(hp5.override(model_type="simple"), 3))
    assert n_simple == 9336610 + 3 * 16 + 16 * 768 + 16 * 256 + 16 * 1025


def test_library_exports_every_declared_symbol(tb):
    capi = tb.capi
    lib = capi.load()
    assert lib.taco_abi_version() == capi.TACO_ABI_VERSION
    header = open(os.path.join(ROOT, "include", "taco_capi.h")).read()
    declared = set(re.findall(r"\b(taco_[a-z_0-9]+)\s*\(", header))
    declared -= {"taco_model_s", "taco_gl_s"}
    assert declared == set(capi.DECLARED_SYMBOLS), declared ^ set(capi.DECLARED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name)
    # error behaviour without touching the GPU
    h = ctypes.c_void_p()
    cfg = capi.make_config(tb.hparams, 1, "none", "fp32", 0, 80)
    cfg.abi_version = 999
    assert lib.taco_create(ctypes.byref(h), ctypes.byref(cfg)) == -1
    assert b"ABI version" in lib.taco_last_error()
    cfg.abi_version = capi.TACO_ABI_VERSION
    assert lib.taco_create(ctypes.byref(h), ctypes.byref(cfg)) == 0
    nbytes = ctypes.c_size_t()
    assert lib.taco_workspace_bytes(h, 32, 128, 800, 1, ctypes.byref(nbytes)) == 0 and nbytes.value > 1 << 30
    assert lib.taco_workspace_bytes(h, 0, 128, 800, 1, ctypes.byref(nbytes)) == -2
    assert lib.taco_forward(h, None, None) == -1
    assert lib.taco_destroy(h) == 0


def test_engine_fails_loudly_without_cuda(tb, hp5):
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(tb.capi.TacoError, match="no CPU fallback"):
        tb.Engine(hp5, 1)


def test_config_rejects_unbuilt_variants(tb, hp5):
    with pytest.raises(tb.capi.TacoError, match="Unkown attention type"):
        tb.capi.make_config(hp5.override(attention_type="luong"), 1, "none", "fp32", 0, 80)
    with pytest.raises(tb.capi.TacoError):
        tb.capi.make_config(hp5.override(dec_prenet_sizes=[256, 128, 64]), 1, "none", "fp32", 0, 80)


def test_golden_fixtures_reproduce_from_oracle(tb, hp5):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_golden as mg
    from oracle import tacotron_oracle as O
    gold = np.load(os.path.join(ROOT, "tests", "golden", "tacotron_train_small.npz"))
    named = mg.golden_params(hp5)
    b = mg.golden_batch()
    out = O.forward(named, hp5, b["inputs"], b["input_lengths"], 1, None, b["mel_targets"], b["linear_targets"], speaker_mode="none")
    assert np.abs(out["mel_outputs"].numpy() - gold["mel_outputs"]).max() < 1e-5
    assert np.abs(out["linear_outputs"].numpy() - gold["linear_outputs"]).max() < 1e-5
    ls = O.losses(out, b["mel_targets"], b["linear_targets"], b["loss_coeff"], hp5)
    assert abs(float(ls["loss"]) - gold["scalars"][0]) < 1e-5


def test_bench_reference_arm_line_shape():
    # the reference arm runs the CPU oracle; use a tiny override so the check stays fast
    code = ("import bench, json; bench.CFG.update(N=2, T_in=8, T_out=10); "
            "import argparse; a=argparse.Namespace(gpus=1, steps=1, warmup=0); bench.run_reference(a)")
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "mel-frames/s" and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["value"] > 0


_DP_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["TACO_ROOT"])
import tacotron_b200
from importlib import import_module
D = import_module("multi-speaker-tacotron-tensorflow_b200.dist")
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
r = dist.get_rank()
flat = torch.arange(8, dtype=torch.float32) * (r + 1)
scale = D.allreduce_sum_(flat)
assert abs(scale - 0.5) < 1e-9
assert torch.allclose(flat * scale, torch.arange(8, dtype=torch.float32) * 1.5)
st = torch.full((4,), float(r)); D.average_bn_state_(st); assert torch.allclose(st, torch.full((4,), 0.5))
assert D.shard_rows(64, r, 2) == ((0, 32) if r == 0 else (32, 64)) and D.shard_rows(5, 1, 2) == (3, 5)
dist.destroy_process_group()
print("rank", r, "ok")
'''


def test_data_parallel_plumbing_gloo_world2(tmp_path):
    script = tmp_path / "dp_worker.py"
    script.write_text(_DP_WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29541", TACO_ROOT=ROOT)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=300)
        assert p.returncode == 0, out[-2000:]


# ---- input pipeline (datasets/datafeeder.py mirror) --------------------------------------------------------------------
def _write_examples(tmp_path, name, n, rng, r=5):
    from importlib import import_module
    df = import_module("multi-speaker-tacotron-tensorflow_b200.datasets.datafeeder")
    d = tmp_path / name
    d.mkdir()
    for i in range(n):
        frames = int(rng.randint(150, 400))
        ntok = int(rng.randint(50, 90))
        df.write_example(str(d / ("ex%03d.npz" % i)), rng.randint(2, 80, ntok), rng.rand(frames, 80), rng.rand(frames, 1025),
                         loss_coeff=1.0 if i % 3 else 0.5)
    return str(d)


def test_prepare_batch_follows_reference_padding_rules():
    """datafeeder.py:289-328: tokens padded with 0 to the longest row; targets padded with 0 to round_up(longest + 1, r)."""
    import numpy as np
    from importlib import import_module
    df = import_module("multi-speaker-tacotron-tensorflow_b200.datasets.datafeeder")
    rng = np.random.RandomState(0)
    batch = []
    for frames, ntok, spk in ((7, 3, 0), (10, 5, 1), (4, 2, 0)):
        batch.append((rng.randint(2, 80, ntok).astype(np.int32), 0.5 + spk, rng.rand(frames, 80).astype(np.float32),
                      rng.rand(frames, 1025).astype(np.float32), spk, frames))
    ref = [tuple(x) for x in batch]
    feed = df.prepare_batch(list(batch), 5, rng, data_type="test")
    assert feed["inputs"].shape == (3, 5) and feed["inputs"].dtype == np.int32
    assert feed["mel_targets"].shape == (3, 15, 80) and feed["linear_targets"].shape == (3, 15, 1025)      # round_up(10 + 1, 5)
    for i, x in enumerate(ref):
        assert (feed["inputs"][i, :len(x[0])] == x[0]).all() and (feed["inputs"][i, len(x[0]):] == 0).all()
        assert feed["input_lengths"][i] == len(x[0]) and feed["loss_coeff"][i] == np.float32(x[1]) and feed["speaker_id"][i] == x[4]
        assert (feed["mel_targets"][i, :x[5]] == x[2]).all() and (feed["mel_targets"][i, x[5]:] == 0).all()
        assert (feed["linear_targets"][i, :x[5]] == x[3]).all() and (feed["linear_targets"][i, x[5]:] == 0).all()
    # exact multiple: the reference still adds one frame before rounding (max_len + 1)
    b2 = [(np.arange(2, 6, dtype=np.int32), 1, np.ones((10, 80), np.float32), np.ones((10, 1025), np.float32), 0, 10)]
    assert df.prepare_batch(b2, 5, rng)["mel_targets"].shape[1] == 15
    b3 = [(np.arange(2, 6, dtype=np.int32), 1, np.ones((9, 80), np.float32), np.ones((9, 1025), np.float32), 0, 9)]
    assert df.prepare_batch(b3, 5, rng)["mel_targets"].shape[1] == 10
    # in-place assembly into (oversized) staging buffers returns views of them
    out = dict(inputs=np.full((4, 9), 7, np.int32), input_lengths=np.zeros(4, np.int32), loss_coeff=np.zeros(4, np.float32),
               mel_targets=np.full((4, 20, 80), 3, np.float32), linear_targets=np.full((4, 20, 1025), 3, np.float32),
               speaker_id=np.zeros(4, np.int32))
    f2 = df.prepare_batch([tuple(x) for x in ref], 5, rng, data_type="test", out=out)
    for k in feed:
        assert np.array_equal(f2[k], feed[k]), k
    assert np.shares_memory(f2["mel_targets"], out["mel_targets"])


def test_datafeeder_groups_sorts_and_feeds(tmp_path):
    """One group = batch_size * batches_per_group examples, bucketed by frame count (datafeeder.py:211-242); every batch obeys
    the length filter (:44-45); ranks draw different streams; the test split is the same on every rank."""
    import types
    import numpy as np
    from importlib import import_module
    tb = import_module("multi-speaker-tacotron-tensorflow_b200")
    df = import_module("multi-speaker-tacotron-tensorflow_b200.datasets.datafeeder")
    rng = np.random.RandomState(1)
    dirs = [_write_examples(tmp_path, "spk_a", 24, rng), _write_examples(tmp_path, "spk_b", 24, rng)]
    hp = tb.hparams.override(reduction_factor=5, initial_phase_step=0)
    cfg = types.SimpleNamespace(random_seed=123, skip_path_filter=False)
    feeders = [df.DataFeeder(dirs, hp, cfg, batches_per_group=4, data_type="train", batch_size=4, rank=r, world=2, log=lambda *_: None)
               for r in range(2)]
    assert feeders[0].path_dict == feeders[1].path_dict                       # identical train/test split on both ranks
    assert all(len(v) == 24 - 4 for v in feeders[0].path_dict.values())      # the last batch_size paths are held out (:64-67)
    test_feeder = df.DataFeeder(dirs, hp, cfg, batches_per_group=2, data_type="test", batch_size=4, log=lambda *_: None)
    # (reference quirk kept: 'test' takes the last batch_size paths of the UNshuffled listing while 'train' drops the last
    # batch_size of the shuffled one, datafeeder.py:36-37,64-67 - the two splits are not disjoint)
    assert all(len(v) == 4 for v in test_feeder.path_dict.values())
    groups = [f._next_group() for f in feeders]
    for grp in groups:
        assert len(grp) == 4 and all(len(b) == 4 for b in grp)
        spans = sorted((min(x[-1] for x in b), max(x[-1] for x in b)) for b in grp)
        assert all(spans[i][1] <= spans[i + 1][0] for i in range(3))         # bucketed: batches cover disjoint frame ranges
        assert {x[4] for b in grp for x in b} == {0, 1}                      # both speakers present, ids by directory order
    assert [x[-1] for b in groups[0] for x in b] != [x[-1] for b in groups[1] for x in b]
    f = feeders[0]
    f.start_in_session(None, 0)
    try:
        seen = 0
        for _ in range(6):
            b = f.next_batch()
            assert b["inputs"].shape[0] == 4 and b["mel_targets"].shape[1] % 5 == 0 and b["mel_targets"].shape[2] == 80
            assert b["linear_targets"].shape[:2] == b["mel_targets"].shape[:2] and b["speaker_id"].shape == (4,)
            L = b["input_lengths"].numpy()
            assert (b["inputs"].numpy()[np.arange(4), L - 1] != 0).all()
            nz = (np.abs(b["mel_targets"].numpy()).sum(-1) > 0).sum(1)
            assert nz.max() + 1 <= b["mel_targets"].shape[1] < nz.max() + 1 + 5
            seen += 1
        assert seen == 6
    finally:
        f.stop()
    assert test_feeder.static_batches is not None and len(test_feeder.static_batches) == 2


def test_librosa_trim_end_matches_framewise_definition(tb):
    """synthesizer.librosa_trim_end against librosa 0.5.1's effects.trim written out frame by frame."""
    from importlib import import_module
    syn = import_module("multi-speaker-tacotron-tensorflow_b200.synthesizer")
    rng = np.random.RandomState(0)
    y = np.concatenate([rng.randn(30000) * 0.3, rng.randn(20000) * 1e-4]).astype(np.float32)      # speech, then near-silence
    fl, hop, top_db = 5120, 256, 50.0
    frames = [y[i:i + fl].astype(np.float64) for i in range(0, len(y) - fl + 1, hop)]
    mse = np.array([np.mean(f * f) for f in frames])
    db = 10 * np.log10(np.maximum(1e-10, mse)) - 10 * np.log10(max(1e-10, mse.max()))
    last = np.flatnonzero(db > -top_db)[-1]
    want = min(len(y), (last + 1) * hop)
    got = syn.librosa_trim_end(y)
    assert got == want and 24000 < got < 31000
    assert syn.librosa_trim_end(np.zeros(100, np.float32)) == 100                 # shorter than a frame: untouched
    assert syn.librosa_trim_end(rng.randn(20000).astype(np.float32)) >= 20000 - fl   # no silence: (almost) nothing cut


def test_generate_endpoint_md5_cache_and_errors(tb, tmp_path):
    """app.py mirror: /generate?text=&speaker_id= -> wav, cached under <root>/<model>/<md5>.<speaker>.0.wav (app.py:55-99)."""
    import hashlib
    import threading
    import urllib.error
    import urllib.parse
    import urllib.request
    from importlib import import_module
    app = import_module("multi-speaker-tacotron-tensorflow_b200.app")

    class FakeSynthesizer:
        calls = []

        def synthesize(self, texts=None, tokens=None, paths=None, speaker_ids=None, attention_trim=False, **_kw):
            self.calls.append((texts, tokens, speaker_ids, attention_trim))
            if texts and "fail" in texts[0]:
                raise RuntimeError("synthesis failed")
            os.makedirs(os.path.dirname(paths[0]), exist_ok=True)
            with open(paths[0], "wb") as f:
                f.write(b"RIFF" + (texts[0] if texts else " ".join(map(str, tokens[0]))).encode("utf-8"))
            return [True]

    syn = FakeSynthesizer()
    server = app.make_server(syn, str(tmp_path / "logs" / "son_2017"), port=0, audio_root=str(tmp_path / "audio"), host="127.0.0.1")
    port = server.server_address[1]
    th = threading.Thread(target=server.serve_forever, daemon=True)
    th.start()
    try:
        text = "안녕하세요"
        url = "http://127.0.0.1:%d/generate?text=%s&speaker_id=1" % (port, urllib.parse.quote(text))
        for _ in range(2):
            with urllib.request.urlopen(url) as r:
                assert r.status == 200 and r.headers["Content-Type"] == "audio/wav"
                md5 = hashlib.md5(text.encode("utf-8")).hexdigest()
                assert md5 in r.headers["Content-Disposition"] and r.read() == b"RIFF" + text.encode("utf-8")
        assert len(syn.calls) == 1 and syn.calls[0] == ([text], None, [1], True)             # second request came from the cache
        assert os.path.exists(str(tmp_path / "audio" / "son_2017" / ("%s.1.0.wav" % md5)))
        with urllib.request.urlopen("http://127.0.0.1:%d/generate?tokens=5+9+1" % port) as r:   # speaker_id defaults to 0
            assert r.read() == b"RIFF5 9 1"
        with urllib.request.urlopen("http://127.0.0.1:%d/generate?speaker_id=0" % port) as r:
            assert json.loads(r.read()) == {}
        for bad in ("/generate?text=fail&speaker_id=0", "/generate?text=x&speaker_id=abc"):
            with pytest.raises(urllib.error.HTTPError) as e:
                urllib.request.urlopen("http://127.0.0.1:%d%s" % (port, bad))
            assert e.value.code == 400 and json.loads(e.value.read()) == {"success": False}
        with pytest.raises(urllib.error.HTTPError) as e:
            urllib.request.urlopen("http://127.0.0.1:%d/" % port)
        assert e.value.code == 404
    finally:
        server.shutdown()
        server.server_close()


def test_product_path_never_touches_the_oracle():
    """The oracle (and the TF / librosa / jamo stand-ins under it) is test infrastructure: no module of the shipped package,
    the command-line wrappers or the C sources may import, load or link anything under oracle/ (bench.py's cpu_baseline /
    --impl reference legs and __graft_entry__.smoke() are the two allowed callers outside tests/)."""
    import ast
    pkg = os.path.join(ROOT, "multi-speaker-tacotron-tensorflow_b200")
    files = [os.path.join(d, f) for d, _, fs in os.walk(pkg) for f in fs if f.endswith(".py")]
    files += [os.path.join(ROOT, f) for f in ("train.py", "synthesizer.py", "app.py", "tacotron_b200.py")]
    for path in files:
        tree = ast.parse(open(path, encoding="utf-8").read())
        for node in ast.walk(tree):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                names = [node.module or ""]
            elif isinstance(node, ast.Call) and getattr(node.func, "id", getattr(node.func, "attr", "")) == "import_module" and node.args \
                    and isinstance(node.args[0], ast.Constant):
                names = [str(node.args[0].value)]
            for n in names:
                assert not (n == "oracle" or n.startswith("oracle.") or n.split(".")[0] in ("tensorflow", "librosa", "jamo")), (path, n)
    for d, _, fs in os.walk(os.path.join(pkg, "csrc")):
        for f in fs:
            assert "oracle" not in open(os.path.join(d, f), encoding="utf-8").read(), f
    # bench.py: the oracle only inside the CPU legs
    src = open(os.path.join(ROOT, "bench.py"), encoding="utf-8").read()
    tree = ast.parse(src)
    for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)]:
        uses = any(isinstance(n, ast.ImportFrom) and (n.module or "").startswith("oracle") for n in ast.walk(fn))
        assert uses == (fn.name in ("cpu_baseline", "synth_rtf_cpu")), fn.name


def test_decoder_wavefront_chunk_boundaries(tb):
    """The time chunks of the decoder wavefront (model_decoder.cu: wave_chunks) partition [0, Td) in order, keep every chunk
    >= 8 steps, never exceed the requested count, and end with the short chunk that trims the pipeline's fill / drain."""
    lib = tb.capi.load()
    lib.taco_debug_wave_chunks.restype = ctypes.c_int
    lib.taco_debug_wave_chunks.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(ctypes.c_int32), ctypes.c_int32]
    buf = (ctypes.c_int32 * 64)()
    for Td in list(range(1, 70)) + [160, 161, 200, 999, 1000]:
        for want in (1, 2, 3, 4, 5, 8, 16):
            n = lib.taco_debug_wave_chunks(Td, want, buf, 64)
            ch = [(buf[2 * i], buf[2 * i + 1]) for i in range(n)]
            assert 1 <= n <= max(1, want) and ch[0][0] == 0 and ch[-1][1] == Td, (Td, want, ch)
            assert all(a[1] == b[0] for a, b in zip(ch, ch[1:])) and all(t1 > t0 for t0, t1 in ch), (Td, want, ch)
            if n > 1:
                assert min(t1 - t0 for t0, t1 in ch) >= 8 or Td < 16, (Td, want, ch)
                assert ch[-1][1] - ch[-1][0] <= max(t1 - t0 for t0, t1 in ch[:-1]), (Td, want, ch)      # the last chunk is the short one
    assert lib.taco_debug_wave_chunks(160, 4, buf, 64) == 4 and [buf[i] for i in range(8)] == [0, 56, 56, 112, 112, 144, 144, 160]


def test_manual_attention_second_pass_matches_the_reference_indexing(tb):
    """synthesizer.py:171-196 of the reference, restated literally (its own variable names), against manual_alignments_from:
    mode 1 = zeros + a one at (argmax decoder step of every INPUT position, that position); mode 3 = the same ones written into
    the transposed soft alignments; mode 2 cannot run in the reference (np.pow)."""
    from importlib import import_module
    syn = import_module("multi-speaker-tacotron-tensorflow_b200.synthesizer")
    rng = np.random.RandomState(5)
    alignments = rng.rand(3, 7, 11).astype(np.float32)             # [N, T_in (E), T_dec (D)]
    alignments /= alignments.sum(1, keepdims=True)
    # --- the reference's lines for mode 1
    alignments_T = np.transpose(alignments, [0, 2, 1])
    new_alignments = np.zeros_like(alignments_T)
    for idx in range(len(alignments)):
        argmax = alignments[idx].argmax(1)
        new_alignments[idx][(argmax, range(len(argmax)))] = 1
    got = syn.manual_alignments_from(alignments, 1)
    assert got.shape == (3, 11, 7) and np.array_equal(got, new_alignments)
    assert got.sum() <= 3 * 7 and (got.sum(1) <= 1).all()            # at most one decoder step per input position
    # --- mode 3
    new_alignments = np.transpose(alignments, [0, 2, 1]).copy()
    for idx in range(len(alignments)):
        argmax = alignments[idx].argmax(1)
        new_alignments[idx][(argmax, range(len(argmax)))] = 1
    got3 = syn.manual_alignments_from(alignments, 3)
    assert np.array_equal(got3, new_alignments) and not np.array_equal(got3, got)
    assert np.array_equal(alignments_T, np.transpose(alignments, [0, 2, 1]))   # the input is not modified
    with pytest.raises(NotImplementedError, match="np.pow"):
        syn.manual_alignments_from(alignments, 2)


def test_hparams_parse_keeps_list_values_and_feeder_thread_joins(tb, tmp_path):
    hp = tb.hparams.override()
    hp.parse("enc_proj_sizes=[128,128],reduction_factor=5,post_proj_sizes=[256, 80]")
    assert hp.enc_proj_sizes == [128, 128] and hp.post_proj_sizes == [256, 80] and hp.reduction_factor == 5
    with pytest.raises(KeyError):
        hp.parse("no_such_hparam=1")
    # threading.Thread has a private _stop() on Python 3.12: the feeder's stop flag must not shadow it (join() would raise)
    from importlib import import_module
    df = import_module("multi-speaker-tacotron-tensorflow_b200.datasets.datafeeder")
    import types
    d = _write_examples(tmp_path, "spk", 24, np.random.RandomState(0))
    feeder = df.DataFeeder([d], tb.hparams.override(reduction_factor=5, initial_phase_step=0), types.SimpleNamespace(random_seed=1, skip_path_filter=False),
                           batches_per_group=2, data_type="train", batch_size=4, log=lambda *_: None)
    assert not callable(getattr(feeder, "_stop_event", None)) and callable(feeder._stop)
    feeder.start_in_session(None, 0)
    feeder.stop()
    feeder.join(timeout=30)
    assert not feeder.is_alive()


def test_bench_legs_job_reports_a_failed_child_instead_of_raising():
    """bench.py runs the C3 / C5 legs of a multi-GPU run as a child job; whatever happens in it must come back as an error
    record in the JSON line, never as an exception that loses the headline measurement (here: no GPU, so the child fails)."""
    import argparse
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)
    out = bench.run_legs_job(argparse.Namespace(precision="bf16", targets="bf16"), 2)
    assert set(out) == {"c3", "c5"}
    if not __import__("torch").cuda.is_available():
        assert "error" in out["c3"] and "error" in out["c5"] and "legs job" in out["c3"]["error"]
