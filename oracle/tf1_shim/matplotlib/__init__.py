"""Import stub for matplotlib (absent here): the reference's utils/plot.py configures it at import time; the fixture paths
never draw."""


def use(*_a, **_k):
    pass


def rc(*_a, **_k):
    pass
