"""Module-level ``__getattr__`` helper for the overlay files.

An overlay file shadows the reference file of the same dotted path (both trees are namespace
packages).  Names the overlay does not define are looked up in the shadowed reference file, so
e.g. ``from modules.diffusionmodules.model import Encoder`` (used by the reference's VQGAN
baselines, out of scope here) keeps working when the reference tree is on ``sys.path``.
"""
import importlib.util
import os
import sys


def make_getattr(mod_name, mod_file):
    rel = mod_name.replace(".", os.sep) + ".py"
    cache = {}

    def __getattr__(name):
        if "mod" not in cache:
            cache["mod"] = None
            here = os.path.abspath(mod_file)
            for root in list(sys.path) + [os.getcwd()]:
                cand = os.path.abspath(os.path.join(root or ".", rel))
                if cand != here and os.path.exists(cand):
                    spec = importlib.util.spec_from_file_location(mod_name + "__reference", cand)
                    m = importlib.util.module_from_spec(spec)
                    spec.loader.exec_module(m)
                    cache["mod"] = m
                    break
        m = cache["mod"]
        if m is not None and hasattr(m, name):
            return getattr(m, name)
        raise AttributeError(f"module {mod_name!r} has no attribute {name!r}")

    return __getattr__
