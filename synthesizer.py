#!/usr/bin/env python
"""``python synthesizer.py --load_path logs/<run> --tokens "5 9 23 1"`` — the reference's synthesis command line
(synthesizer.py:372-389) on the B200 engine."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from importlib import import_module  # noqa: E402

if __name__ == "__main__":
    import_module("multi-speaker-tacotron-tensorflow_b200.synthesizer").main()
