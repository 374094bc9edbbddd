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


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
def test_unmodified_pipelines_import_the_dropins():
    """``install()`` + the reference's UNMODIFIED pipeline modules: their call sites
    (t2v/i2v_sphere_panorama_pipeline.py) resolve to the replacement classes.  diffusers /
    pytorch_lightning are absent here, so they are stubbed like in SURVEY §8c; nothing is computed."""
    import torch
    saved_path = list(sys.path)
    prefixes = ("utils", "pipeline", "lvdm", "diffusers", "pytorch_lightning")
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in prefixes}
    try:
        for k in list(saved):
            del sys.modules[k]

        def mod(name, **attrs):
            m = types.ModuleType(name)
            m.__dict__.update(attrs)
            sys.modules[name] = m
            return m

        mod("pytorch_lightning", LightningModule=torch.nn.Module, seed_everything=lambda *a, **k: None)
        mod("pytorch_lightning.utilities", rank_zero_only=lambda f: f)

        class DiffusionPipeline:
            def register_modules(self, **kw):
                self.__dict__.update(kw)

        class _Mixin:
            pass

        logging = mod("diffusers.logging", get_logger=lambda name: __import__("logging").getLogger(name))
        d = mod("diffusers", DiffusionPipeline=DiffusionPipeline, logging=logging,
                ConfigMixin=type("ConfigMixin", (), {}), SchedulerMixin=type("SchedulerMixin", (), {}))
        d.__path__ = []
        mod("diffusers.configuration_utils", ConfigMixin=d.ConfigMixin, register_to_config=lambda f: f)
        mod("diffusers.schedulers", __path__=[])
        mod("diffusers.schedulers.scheduling_utils", SchedulerMixin=d.SchedulerMixin)
        mod("diffusers.utils", BaseOutput=type("BaseOutput", (dict,), {}), logging=logging, __path__=[])
        mod("diffusers.utils.torch_utils", randn_tensor=lambda *a, **k: None)
        sys.path.insert(0, REF)
        sys.dont_write_bytecode = True
        import dynamicscaler_b200
        dynamicscaler_b200.install()
        t2v = importlib.import_module("pipeline.t2v_sphere_panorama_pipeline")
        i2v = importlib.import_module("pipeline.i2v_sphere_panorama_pipeline")
        from dynamicscaler_b200 import (diffusion_utils, panorama_tensor_utils, ring_panorama_tensor_utils,
                                        shift_window_utils, tensor_utils)
        assert t2v.PanoramaLatentProxy is panorama_tensor_utils.PanoramaLatentProxy
        assert t2v.PanoramaTensor is panorama_tensor_utils.PanoramaTensor
        assert t2v.RingLatent is shift_window_utils.RingLatent and t2v.RingImageTensor is shift_window_utils.RingImageTensor
        assert t2v.mix_latents_with_mask is tensor_utils.mix_latents_with_mask
        assert t2v.resize_video_latent is diffusion_utils.resize_video_latent
        assert i2v.RingPanoramaLatentProxy is ring_panorama_tensor_utils.RingPanoramaLatentProxy
        assert i2v.PanoramaTensor is panorama_tensor_utils.PanoramaTensor
        # the reference's other utils modules are still the reference's own
        assert importlib.import_module("utils.multi_prompt_utils").__file__.startswith(REF)
        assert hasattr(t2v.VC2_Pipeline_T2V_SpherePano, "basic_sample_shift_shpere_panorama")
    finally:
        sys.path[:] = saved_path
        for k in [k for k in sys.modules if k.split(".")[0] in prefixes]:
            del sys.modules[k]
        sys.modules.update(saved)
