"""Run one of the reference's scripts UNCHANGED on the libst3d path:

    python -m st3d.run /path/to/second_approach.py --n_views 8 --size 512 ...

The script's `from pytorch3d... import ...`, `from utils import *`, `from losses import *` and
`from style_transfer import *` then resolve to the modules under `compat/` (same names, same call
surface) instead of the third-party library and the reference's torch-only helpers.
Set ST3D_KEEP_REFERENCE_HELPERS=1 to keep the script's own utils/losses/style_transfer modules and swap
only the `pytorch3d` package.  ST3D_SEED=<int> seeds torch / random / numpy before the script starts (the scripts
draw their cameras from the global generators and never seed them); ST3D_VGG_RANDOM_INIT=1 makes `get_vgg()` build
the VGG-19 with seeded random weights instead of downloading the ImageNet checkpoint (offline boxes).
"""
import os
import runpy
import sys

PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMPAT = os.path.join(PKG, "compat")


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit(__doc__)
    script = os.path.abspath(argv[0])
    sys.argv = [script] + argv[1:]
    if os.environ.get("ST3D_SEED") is not None:
        import random

        import numpy as np
        import torch
        seed = int(os.environ["ST3D_SEED"])
        torch.manual_seed(seed)
        random.seed(seed)
        np.random.seed(seed)
    for p in (PKG,):
        if p not in sys.path:
            sys.path.insert(0, p)
    if os.environ.get("ST3D_KEEP_REFERENCE_HELPERS") == "1":
        # only the pytorch3d package comes from compat: expose it through a directory holding nothing else
        import importlib.util
        spec = importlib.util.spec_from_file_location("pytorch3d", os.path.join(COMPAT, "pytorch3d", "__init__.py"),
                                                      submodule_search_locations=[os.path.join(COMPAT, "pytorch3d")])
        mod = importlib.util.module_from_spec(spec)
        sys.modules["pytorch3d"] = mod
        spec.loader.exec_module(mod)
        sys.path.insert(0, os.path.dirname(script))
    else:
        sys.path.insert(0, os.path.dirname(script))
        sys.path.insert(0, COMPAT)          # ahead of the script's own directory
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
