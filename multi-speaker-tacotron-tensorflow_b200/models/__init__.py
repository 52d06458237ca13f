"""Mirror of the reference's ``models`` package surface (models/__init__.py:6-17)."""
import os
from glob import glob

from .tacotron import Tacotron


def create_model(hparams):
    return Tacotron(hparams)


def get_most_recent_checkpoint(checkpoint_dir):
    """Latest ``model.ckpt-<step>.pt`` in a run directory (the reference globs ``*.ckpt-*.data-*``, models/__init__.py:10-17)."""
    paths = [p for p in glob(os.path.join(checkpoint_dir, "*.ckpt-*.pt"))]
    if not paths:
        raise FileNotFoundError(" [!] No checkpoint found in {}".format(checkpoint_dir))
    idxes = [int(os.path.basename(p).split("-")[1].split(".")[0]) for p in paths]
    max_idx = max(idxes)
    lastest_checkpoint = os.path.join(checkpoint_dir, "model.ckpt-{}.pt".format(max_idx))
    print(" [*] Found lastest checkpoint: {}".format(lastest_checkpoint))
    return lastest_checkpoint
