from . import _absent as _jamo_char_to_hcj  # noqa: F401
