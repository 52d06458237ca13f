"""Data-parallel numerical equivalence on real GPUs (needs >= 2 CUDA devices: run under `gpurun --gpus 2`).

SURVEY.md 4 / 8e: the only collective of the path is ONE all-reduce of the flat gradient per step; the replicas must stay
bit-identical and follow the reference's single-device update applied to the mean gradient."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _run(world, precision):
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT="29577",
                   DP_PRECISION=precision)
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "dp_worker.py")], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        out, _ = p.communicate(timeout=900)
        outs.append(out)
        assert p.returncode == 0, out[-3000:]
    assert any("DP_OK" in o for o in outs), outs[0][-2000:]
    return [line for o in outs for line in o.splitlines() if "DP_OK" in line][0]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_two_gpu_data_parallel_steps_match_the_oracle_mean_gradient_step(precision):
    print(_run(2, precision))
