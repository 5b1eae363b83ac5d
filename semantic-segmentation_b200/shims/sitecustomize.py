"""Launcher glue for the unmodified reference train.py on torch >= 2.0 (SURVEY.md §7 hard part 8): argv spelling of the
local rank, the NumPy 2 `np.int` alias, torch.load's torch-1.x default, and binding the overlay `network` package.

torch.distributed.launch / torchrun export LOCAL_RANK (and pass the dashed ``--local-rank``), while train.py:112 only
declares ``--local_rank`` and train.py:298 reads it from argv. When this directory is on PYTHONPATH, Python imports
this module at start-up and the two spellings are reconciled without touching the reference file."""
import os
import sys

if sys.argv and os.path.basename(sys.argv[0]) == "train.py":
    argv = []
    have = False
    for a in sys.argv[1:]:
        if a.startswith("--local-rank"):
            a = a.replace("--local-rank", "--local_rank", 1)
        if a.startswith("--local_rank"):
            have = True
        argv.append(a)
    if not have and "LOCAL_RANK" in os.environ:
        argv += ["--local_rank", os.environ["LOCAL_RANK"]]
    sys.argv[1:] = argv
    # `python train.py` puts the script's own directory (the reference checkout, with its network/ package) in FRONT of
    # PYTHONPATH, after this module has run: bind the overlay's `network` package now, while the overlay directory
    # (first on PYTHONPATH) still wins, so that train.py's `import network` finds it in sys.modules.
    if os.environ.get("B200SEG_OVERLAY", "1") != "0":
        try:
            import network as _network  # noqa: F401
            if "semantic-segmentation_b200" not in os.path.abspath(_network.__file__):
                del sys.modules["network"]
        except Exception as _e:  # noqa
            print("b200seg sitecustomize: could not pre-import the overlay network package: %r" % (_e,), file=sys.stderr)
    # torch >= 2.6 defaults torch.load(weights_only=True); the reference's checkpoints (train.py:371, loss/optimizer.py)
    # carry numpy scalars (mean_iu ...) written by its own save code: restore the torch 1.x default it was written for
    try:
        import functools as _ft
        import torch as _torch
        _orig_load = _torch.load

        @_ft.wraps(_orig_load)
        def _load(*a, **k):
            k.setdefault("weights_only", False)
            return _orig_load(*a, **k)
        _torch.load = _load
    except Exception:
        pass
    try:
        import numpy as _np
        if not hasattr(_np, "int"):
            _np.int = int          # network/hrnetv2.py:315 uses the alias NumPy 2 removed
    except Exception:
        pass
