"""Plugin loader used when the reference tree is not importable (GPU box, unit tests).

Same contract as the reference's ``utils/utils.py:41-51``: a config is ``{"target": "a.b.C",
"params": {...}}`` and is instantiated by importing the dotted path.
"""
import importlib


def get_obj_from_str(string):
    module, cls = string.rsplit(".", 1)
    return getattr(importlib.import_module(module), cls)


def instantiate_from_config(config):
    if "target" not in config:
        raise KeyError("Expected key `target` to instantiate.")
    return get_obj_from_str(config["target"])(**config.get("params", dict()))
