"""Minimal stand-in for `colorama` (absent from this image): the reference's scripts only concatenate these strings."""


class _Codes(object):
    def __getattr__(self, name):
        return ""


Style = _Codes()
Fore = _Codes()
Back = _Codes()


def init(*a, **k):
    pass
