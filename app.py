#!/usr/bin/env python
"""``python app.py --load_path logs/<run> --num_speakers N --port 5000`` — the reference's serving command line
(app.py:121-134): ``/generate?text=&speaker_id=`` answers with a wav, cached by the md5 of the text."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from importlib import import_module  # noqa: E402

if __name__ == "__main__":
    raise SystemExit(import_module("multi-speaker-tacotron-tensorflow_b200.app").main())
