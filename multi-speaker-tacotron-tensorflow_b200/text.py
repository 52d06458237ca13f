"""Token ids of the model's input: the reference's symbol table and the jamo decomposition of Hangul text.

The model's vocabulary is fixed by ``text/korean.py:11-21`` / ``text/symbols.py:13``: PAD ``_`` (0), EOS ``~`` (1), the 19 lead
/ 21 vowel / 27 tail conjoining jamo (U+1100.., U+1161.., U+11A8..), the punctuation ``!'(),-.:;?`` and the space — 80
symbols.  ``text_to_sequence`` reproduces the path ``text/__init__.py:23-58`` → ``cleaners.korean_cleaners`` →
``korean.tokenize`` for text that needs no normalisation: every Hangul syllable is decomposed by Unicode arithmetic (what the
``jamo`` package's ``hangul_to_jamo`` does), symbols outside the table are dropped, EOS is appended.

Not reproduced (out of scope): ``korean.normalize`` — the dictionary substitutions, English letter /
word read-outs, quote splitting (nltk) and number-to-Korean conversion (``text/korean.py:151-304``, ``ko_dictionary.py``).
Digits and Latin letters are therefore dropped like any other symbol outside the table; normalise such text before calling,
or pass ``Synthesizer(text_to_sequence=...)`` a full front end.
"""
from __future__ import annotations

from typing import Iterable, List

import numpy as np

PAD, EOS = "_", "~"
PUNC, SPACE = "!'(),-.:;?", " "
JAMO_LEADS = "".join(chr(c) for c in range(0x1100, 0x1113))
JAMO_VOWELS = "".join(chr(c) for c in range(0x1161, 0x1176))
JAMO_TAILS = "".join(chr(c) for c in range(0x11A8, 0x11C3))
ALL_SYMBOLS = PAD + EOS + JAMO_LEADS + JAMO_VOWELS + JAMO_TAILS + PUNC + SPACE
symbols = ALL_SYMBOLS
_symbol_to_id = {s: i for i, s in enumerate(ALL_SYMBOLS)}
_id_to_symbol = {i: s for i, s in enumerate(ALL_SYMBOLS)}

_SBASE, _NV, _NT = 0xAC00, 21, 28           # Hangul syllable block: 19 leads x 21 vowels x 28 tails (tail 0 = none)


def hangul_to_jamo(text: str) -> str:
    """Precomposed syllables (U+AC00..U+D7A3) -> lead + vowel (+ tail) conjoining jamo; everything else unchanged."""
    out = []
    for ch in text:
        c = ord(ch) - _SBASE
        if 0 <= c < 19 * _NV * _NT:
            lead, vowel, tail = c // (_NV * _NT), (c // _NT) % _NV, c % _NT
            out.append(chr(0x1100 + lead) + chr(0x1161 + vowel) + (chr(0x11A7 + tail) if tail else ""))
        else:
            out.append(ch)
    return "".join(out)


def jamo_to_korean(text: str) -> str:
    """Inverse of hangul_to_jamo on well-formed input (text/korean.py:54-81): lead + vowel (+ tail) -> one syllable."""
    out, i, n = [], 0, len(text)
    while i < n:
        ch = text[i]
        if ch in JAMO_LEADS and i + 1 < n and text[i + 1] in JAMO_VOWELS:
            lead, vowel, tail, used = JAMO_LEADS.index(ch), JAMO_VOWELS.index(text[i + 1]), 0, 2
            if i + 2 < n and text[i + 2] in JAMO_TAILS:
                tail, used = JAMO_TAILS.index(text[i + 2]) + 1, 3
            out.append(chr(_SBASE + (lead * _NV + vowel) * _NT + tail))
            i += used
        else:
            out.append(ch)
            i += 1
    return "".join(out)


def text_to_sequence(text: str, as_token: bool = False):
    """text/__init__.py:23-58 for text that needs no normalisation: int32 ids ending in EOS (or the token string)."""
    ids = [_symbol_to_id[s] for s in hangul_to_jamo(text.strip()) if s in _symbol_to_id and s not in (PAD, EOS)]
    ids.append(_symbol_to_id[EOS])
    if as_token:
        return sequence_to_text(ids, combine_jamo=True)
    return np.array(ids, dtype=np.int32)


def sequence_to_text(sequence: Iterable[int], skip_eos_and_pad: bool = False, combine_jamo: bool = False) -> str:
    """text/__init__.py:61-79."""
    result = "".join(_id_to_symbol[int(i)] for i in sequence
                     if int(i) in _id_to_symbol and not (skip_eos_and_pad and _id_to_symbol[int(i)] in (EOS, PAD)))
    return jamo_to_korean(result) if combine_jamo else result


def texts_to_batch(texts: List[str]):
    """Padded int32 batch + lengths as train.py's create_batch_inputs_from_texts builds them for evaluation."""
    seqs = [text_to_sequence(t) for t in texts]
    L = np.array([len(s) for s in seqs], dtype=np.int32)
    out = np.zeros((len(seqs), int(L.max())), dtype=np.int32)
    for i, s in enumerate(seqs):
        out[i, :len(s)] = s
    return out, L
