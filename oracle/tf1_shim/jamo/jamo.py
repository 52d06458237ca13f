from . import _jamo_char_to_hcj  # noqa: F401
