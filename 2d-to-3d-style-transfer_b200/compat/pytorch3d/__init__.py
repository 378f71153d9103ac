"""`pytorch3d`-named compatibility package for the call surface EmaMule/2D-to-3D-Style-Transfer uses
(first_approach.py:15-16, second_approach.py:15-16, utils.py:6-9, losses.py:3).

The real PyTorch3D is a third-party dependency of the reference; this package is NOT it.  It exposes
the same names with the same argument meaning, and routes the per-iteration hot path (rasterize,
texture sample, shade, blend, and their backward) to the hand-written sm_100a kernels of libst3d
through `st3d.functional`.  Only the configuration the reference exercises is accelerated; anything
else raises NotImplementedError instead of silently falling back.
"""
__version__ = "0.0+st3d"
