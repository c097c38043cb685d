"""Shim mapping the small part of ruamel.yaml that sBayes uses onto PyYAML."""
import yaml as _yaml


class YAML:
    def __init__(self, typ=None, **kwargs):
        pass

    def load(self, f):
        if hasattr(f, "read"):
            return _yaml.safe_load(f)
        with open(f, "r") as fh:
            return _yaml.safe_load(fh)

    def dump(self, data, stream=None):
        return _yaml.safe_dump(data, stream)
