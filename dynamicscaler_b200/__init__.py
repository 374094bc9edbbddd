"""dynamicscaler_b200 -- B200 (sm_100a) implementation of DynamicScaler's per-step
panorama tiling path, behind the reference's own Python signatures.

Module map (reference module -> drop-in here):
    utils.shift_window_utils          -> dynamicscaler_b200.shift_window_utils
    utils.panorama_tensor_utils       -> dynamicscaler_b200.panorama_tensor_utils
    utils.ring_panorama_tensor_utils  -> dynamicscaler_b200.ring_panorama_tensor_utils
    utils.tensor_utils                -> dynamicscaler_b200.tensor_utils
    utils.diffusion_utils             -> dynamicscaler_b200.diffusion_utils
``install()`` registers them under the reference's names so the unmodified
t2v/i2v_sphere_panorama_pipeline call sites pick them up.  ``step`` holds the additive
batched step driver (gather-all / merge-all, sharded over GPUs).

Everything computes on the GPU through libdynscaler_sm100.so (C ABI in
include/dynscaler.h); there is no CPU fallback -- importing works without a GPU, calling
does not.
"""
from ._lib import lib, LibraryMissing  # noqa: F401

__all__ = ["install", "lib", "LibraryMissing"]


def install():
    """Register the drop-in modules under the reference's module names
    (``utils.shift_window_utils`` ...) so ``pipeline/*_pipeline.py`` import them."""
    import sys
    import types
    from . import (diffusion_utils, panorama_tensor_utils, ring_panorama_tensor_utils, shift_window_utils,
                   tensor_utils)

    pkg = sys.modules.get("utils")
    if pkg is None:
        try:                                   # the reference's own ``utils`` package, when its checkout is on
            import utils as pkg                # sys.path: its other modules (multi_prompt_utils, ...) stay reachable
        except ImportError:
            pkg = types.ModuleType("utils")
            pkg.__path__ = []
            sys.modules["utils"] = pkg
    for name, mod in (("shift_window_utils", shift_window_utils),
                      ("panorama_tensor_utils", panorama_tensor_utils),
                      ("ring_panorama_tensor_utils", ring_panorama_tensor_utils),
                      ("tensor_utils", tensor_utils),
                      ("diffusion_utils", diffusion_utils)):
        sys.modules["utils." + name] = mod
        setattr(pkg, name, mod)
