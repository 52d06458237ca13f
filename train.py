#!/usr/bin/env python
"""``python train.py --data_paths=datasets/a,datasets/b`` — the reference's training command line (train.py:281-297) on
the B200 engine.  Multi-GPU: ``torchrun --nproc-per-node N train.py ...``."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from importlib import import_module  # noqa: E402

if __name__ == "__main__":
    import_module("multi-speaker-tacotron-tensorflow_b200.train").main()
