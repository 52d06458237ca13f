"""Mirror of the reference's ``models`` package surface (models/__init__.py:6-17)."""
import os
from glob import glob

from .tacotron import Tacotron


def create_model(hparams):
    return Tacotron(hparams)


def get_most_recent_checkpoint(checkpoint_dir):
    """Latest checkpoint of a run directory: ``model.ckpt-<step>.pt`` (this repo) or a TensorFlow ``model.ckpt-<step>``
    prefix written by the reference (it globs ``*.ckpt-*.data-*``, models/__init__.py:10-17); ``tf_checkpoint.load_any``
    reads either.  On equal steps the native file wins."""
    found = {}
    for p in glob(os.path.join(checkpoint_dir, "*.ckpt-*.data-*")):
        step = int(os.path.basename(p).split("-")[1].split(".")[0])
        found[step] = os.path.join(checkpoint_dir, "model.ckpt-{}".format(step))
    for p in glob(os.path.join(checkpoint_dir, "*.ckpt-*.pt")):
        step = int(os.path.basename(p).split("-")[1].split(".")[0])
        found[step] = os.path.join(checkpoint_dir, "model.ckpt-{}.pt".format(step))
    if not found:
        raise FileNotFoundError(" [!] No checkpoint found in {}".format(checkpoint_dir))
    lastest_checkpoint = found[max(found)]
    print(" [*] Found lastest checkpoint: {}".format(lastest_checkpoint))
    return lastest_checkpoint
