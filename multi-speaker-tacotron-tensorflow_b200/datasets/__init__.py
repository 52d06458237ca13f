"""Input pipeline of the training path: the reference's ``datasets/datafeeder.py`` re-built around
pinned staging buffers and an asynchronous host->device copy instead of a TF FIFOQueue."""
from .datafeeder import DataFeeder, get_path_dict, prepare_batch, write_example  # noqa: F401
