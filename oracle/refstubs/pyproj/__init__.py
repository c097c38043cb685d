"""Stub: CRS(...) -> None makes sbayes.preprocessing.ComputeNetwork use Euclidean distances."""


def CRS(*args, **kwargs):
    return None
