"""Drop-in boundary (level 1): the replacement modules expose the reference's classes
and functions with the same parameter names, order and defaults.  Compared against the reference
sources when the checkout is present (build container); the registration under the reference's
module names is checked everywhere.  CPU only, nothing is launched."""
import importlib
import inspect
import os
import sys
import types

import pytest

REF = "/root/reference"

PAIRS = {
    "utils.shift_window_utils": ("dynamicscaler_b200.shift_window_utils",
                                 ["get_dimension_slices_and_sizes", "RingLatent", "RingImageTensor"]),
    "utils.panorama_tensor_utils": ("dynamicscaler_b200.panorama_tensor_utils", ["PanoramaTensor", "PanoramaLatentProxy"]),
    "utils.ring_panorama_tensor_utils": ("dynamicscaler_b200.ring_panorama_tensor_utils",
                                         ["RingPanoramaTensor", "RingPanoramaLatentProxy", "RingLatentProxy"]),
    "utils.tensor_utils": ("dynamicscaler_b200.tensor_utils", ["mix_latents_with_mask", "load_image_tensor_from_path"]),
    "utils.diffusion_utils": ("dynamicscaler_b200.diffusion_utils",
                              ["resize_video_latent", "padding_latents_at_front", "padding_latents_at_end"]),
}


def _params(fn):
    """(name, default) of the public parameters; private ``_x`` keyword extensions are ours."""
    return [(p.name, p.default) for p in inspect.signature(fn).parameters.values() if not p.name.startswith("_")]


def _public_methods(cls):
    return {n: f for n, f in inspect.getmembers(cls, inspect.isfunction) if not n.startswith("_") or n == "__init__"}


def test_install_registers_the_reference_module_names():
    saved = {k: v for k, v in sys.modules.items() if k == "utils" or k.startswith("utils.")}
    try:
        for k in saved:
            del sys.modules[k]
        import dynamicscaler_b200
        dynamicscaler_b200.install()
        for ref_name, (ours, names) in PAIRS.items():
            mod = importlib.import_module(ref_name)
            assert mod is importlib.import_module(ours)
            for n in names:
                assert hasattr(mod, n), (ref_name, n)
    finally:
        for k in [k for k in sys.modules if k == "utils" or k.startswith("utils.")]:
            del sys.modules[k]
        sys.modules.update(saved)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
def test_signatures_match_the_reference():
    saved_path = list(sys.path)
    saved = {k: v for k, v in sys.modules.items() if k == "utils" or k.startswith("utils.") or k.startswith("pytorch_lightning")}
    try:
        for k in [k for k in sys.modules if k == "utils" or k.startswith("utils.")]:
            del sys.modules[k]
        pl = types.ModuleType("pytorch_lightning")                      # only a type hint needs it
        import torch
        pl.LightningModule = torch.nn.Module
        pl.seed_everything = lambda *a, **k: None
        plu = types.ModuleType("pytorch_lightning.utilities")
        plu.rank_zero_only = lambda f: f
        sys.modules.setdefault("pytorch_lightning", pl)
        sys.modules.setdefault("pytorch_lightning.utilities", plu)
        sys.path.insert(0, REF)
        sys.dont_write_bytecode = True
        for ref_name, (ours_name, names) in PAIRS.items():
            ref_mod = importlib.import_module(ref_name)
            assert ref_mod.__file__.startswith(REF), ref_mod.__file__
            ours = importlib.import_module(ours_name)
            for n in names:
                r, o = getattr(ref_mod, n), getattr(ours, n)
                if inspect.isclass(r):
                    rm, om = _public_methods(r), _public_methods(o)
                    missing = set(rm) - set(om)
                    assert not missing, (n, missing)
                    for m in rm:
                        assert _params(rm[m]) == _params(om[m]), (n, m, _params(rm[m]), _params(om[m]))
                else:
                    assert _params(r) == _params(o), (n, _params(r), _params(o))
    finally:
        sys.path[:] = saved_path
        for k in [k for k in sys.modules if k == "utils" or k.startswith("utils.")]:
            del sys.modules[k]
        sys.modules.update(saved)
