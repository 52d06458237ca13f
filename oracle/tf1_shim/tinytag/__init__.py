"""Import stub for `tinytag` (absent here): the reference's audio/get_duration.py imports it at module level; nothing on the
paths the fixtures exercise calls it."""


class TinyTag:
    @staticmethod
    def get(*_a, **_k):
        raise NotImplementedError("tinytag is not installed; only importing the reference's datafeeder module needs the name")
