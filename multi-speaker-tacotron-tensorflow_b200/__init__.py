"""B200-native Tacotron mel-synthesis hot path (drop-in for the reference's models/tacotron.py forward + train step).

The directory name carries a hyphen (it is fixed by the project layout), so import it with
``importlib.import_module("multi-speaker-tacotron-tensorflow_b200")`` or through the root-level ``tacotron_b200`` shim.
"""
from .hparams import HParams, hparams, hparams_debug_string, load_hparams, save_hparams  # noqa: F401
from . import params  # noqa: F401
from . import tf_names  # noqa: F401
from . import tf_checkpoint  # noqa: F401
from . import text  # noqa: F401
from . import capi  # noqa: F401   (ctypes binding; the shared library is only loaded on first use)
from .engine import Engine  # noqa: F401
from .models import Tacotron, create_model, get_most_recent_checkpoint  # noqa: F401

__all__ = ["HParams", "hparams", "hparams_debug_string", "load_hparams", "save_hparams", "params", "capi", "Engine",
           "Tacotron", "create_model", "get_most_recent_checkpoint"]
