"""One C4 synthesis (free-running decoder + post-net + linear + 60-iteration Griffin-Lim) and one analysis front-end call between
cudaProfilerStart/Stop (run under `ncu --profile-from-start off`): the launch list behind profiles/r2_synth_launches.*"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tacotron_b200 as tb
from importlib import import_module
import bench

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
Engine = import_module("multi-speaker-tacotron-tensorflow_b200.engine").Engine
audio = import_module("multi-speaker-tacotron-tensorflow_b200.audio")
hp = tb.hparams.override(reduction_factor=5)
eng = Engine(hp, 1, precision=prec, randomize_bn_state=True)
gl = audio.GriffinLim(hp, max_frames=1000)
tok, L, phase = bench.synth_inputs()
phase = phase.to(eng.dev)
wav_in = torch.from_numpy((np.random.RandomState(0).randn(299700) * 0.05).astype(np.float32)).to(eng.dev)
for _ in range(2):
    out = eng.forward(tok, L, decoder_steps=200)
    wav = gl.inv_spectrogram(out["linear_outputs"][0], phase, n_iters=60)
    lin, mel = gl.spectrograms(wav_in)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
out = eng.forward(tok, L, decoder_steps=200)
wav = gl.inv_spectrogram(out["linear_outputs"][0], phase, n_iters=60)
lin, mel = gl.spectrograms(wav_in)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("frames", out["linear_outputs"].shape, "samples", wav.shape, "analysis", lin.shape, mel.shape)
