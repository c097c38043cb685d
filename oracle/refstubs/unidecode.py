"""Stub for the `unidecode` package (absent in this image); identity is enough for the
sBayes hot path, which only uses it to sanitise names (sbayes/util.py)."""


def unidecode(s):
    return s
