"""Import stub for the `jamo` package (absent here).  text/korean.py only needs the names at import time; the symbol
table the model uses (text/symbols.py -> ALL_SYMBOLS, 80 entries) is built from literal code-point ranges."""


def _absent(*_a, **_k):
    raise NotImplementedError("jamo is not installed; only the symbol table of the reference's text front end is used")


h2j = j2h = hangul_to_jamo = j2hcj = _absent
