"""TEST INFRASTRUCTURE — stand-in for the `jamo` package (absent here) so that the reference's text front end
(/root/reference/text/*.py) can be imported and run unmodified by tools/make_reference_golden.py.  Restates what
jamo 0.4's functions do for modern Hangul: syllable <-> conjoining-jamo conversion by Unicode arithmetic."""
_SBASE, _NV, _NT = 0xAC00, 21, 28


def hangul_to_jamo(hangul_string):
    return (_ for _ in _decompose(hangul_string))


def _decompose(s):
    for ch in s:
        c = ord(ch) - _SBASE
        if 0 <= c < 19 * _NV * _NT:
            yield chr(0x1100 + c // (_NV * _NT))
            yield chr(0x1161 + (c // _NT) % _NV)
            if c % _NT:
                yield chr(0x11A7 + c % _NT)
        else:
            yield ch


def h2j(hangul_string):
    return "".join(_decompose(hangul_string))


def j2h(lead, vowel, tail=None):
    t = (ord(tail) - 0x11A7) if tail else 0
    return chr(_SBASE + ((ord(lead) - 0x1100) * _NV + (ord(vowel) - 0x1161)) * _NT + t)


_LEAD_HCJ = "ㄱㄲㄴㄷㄸㄹㅁㅂㅃㅅㅆㅇㅈㅉㅊㅋㅌㅍㅎ"
_TAIL_HCJ = "ㄱㄲㄳㄴㄵㄶㄷㄹㄺㄻㄼㄽㄾㄿㅀㅁㅂㅄㅅㅆㅇㅈㅊㅋㅌㅍㅎ"


def _jamo_char_to_hcj(char):
    o = ord(char)
    if 0x1100 <= o <= 0x1112:
        return _LEAD_HCJ[o - 0x1100]
    if 0x1161 <= o <= 0x1175:
        return chr(0x314F + o - 0x1161)
    if 0x11A8 <= o <= 0x11C2:
        return _TAIL_HCJ[o - 0x11A8]
    return char


def j2hcj(jamo):
    return "".join(_jamo_char_to_hcj(c) for c in jamo)
