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
    assert all(o % 4 == 0 for o in lay.offsets.values())
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
