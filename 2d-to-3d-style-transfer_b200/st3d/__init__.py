"""st3d -- host-side binding of libst3d.so (hand-written sm_100a kernels behind include/st3d.h).

There is no CPU fallback: every op raises if the CUDA library is missing or a tensor is not on a
CUDA device.  The CPU restatement used by the tests lives in /oracle and is never imported here.
"""
from ._lib import lib, check, library_path, St3dError  # noqa: F401
from . import ops  # noqa: F401
