"""Run a script of the reference tree (train.py, scripts/tools/*.py) with the B200 overlay active.

    cd /path/to/DynamicVectorQuantization
    python -m dynamicvectorquantization_b200.launch train.py --gpus -1 --base configs/stage1/dqvae-dual-r-05_imagenet.yml

Why a launcher: ``python train.py`` puts the script's directory (the reference root) at ``sys.path[0]``, AHEAD of
anything on ``PYTHONPATH``.  ``modules`` / ``models`` are namespace packages in both trees, and a sub-module is
taken from the first portion that has it - so with a plain PYTHONPATH overlay the reference's own files would win
and the job would silently train the stock PyTorch implementation.  Here the overlay is inserted at index 0, the
script's directory right behind it (what ``python script.py`` would have put first), and the script then runs as
``__main__`` under ``runpy``.  ``assert_overlay_active()`` is the check the model classes run on construction.
"""
import os
import runpy
import sys

from . import configs

# dotted paths that must resolve to this package once the overlay is active (Python face)
OVERLAID = ("modules.dynamic_modules.EncoderDual", "modules.dynamic_modules.EncoderTriple",
            "modules.dynamic_modules.DecoderPositional", "modules.vector_quantization.quantize2_mask",
            "modules.diffusionmodules.model", "models.stage1_dynamic.dqvae_dual_feat",
            "models.stage1_dynamic.dqvae_dual_entropy", "models.stage1_dynamic.dqvae_triple_feat")


def assert_overlay_active(paths=OVERLAID):
    """Import every overlaid dotted path and check that it came from the overlay tree, not from the reference."""
    import importlib
    for name in paths:
        mod = importlib.import_module(name)
        f = os.path.abspath(getattr(mod, "__file__", "") or "")
        if not f.startswith(os.path.abspath(configs.OVERLAY) + os.sep):
            raise RuntimeError(f"{name} resolved to {f}, not to the B200 overlay ({configs.OVERLAY}): the reference "
                               f"tree is ahead of the overlay on sys.path - start the job with "
                               f"`python -m dynamicvectorquantization_b200.launch <script> ...`")


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] in ("-h", "--help"):
        print(__doc__)
        return 0
    script = os.path.abspath(argv[0])
    configs.activate_overlay()                       # overlay at sys.path[0], repository root behind it
    script_dir = os.path.dirname(script)
    if script_dir in sys.path:
        sys.path.remove(script_dir)
    sys.path.insert(1, script_dir)                   # where `python script.py` looks first - now second
    assert_overlay_active()
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main())
