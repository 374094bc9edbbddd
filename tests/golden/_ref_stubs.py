"""Stubs that let the reference's pipeline modules import in the build container:
``pytorch_lightning`` and ``diffusers`` are not installed; only the names the reference touches at
import time are provided.  Used by the golden generators only."""
import os
import sys
import types

import torch

REF = os.environ.get("DS_REFERENCE", "/root/reference")
if REF not in sys.path:
    sys.path.insert(0, REF)

# ---- stubs for packages that are not installed ---------------------------------
_pl = types.ModuleType("pytorch_lightning")
_pl.LightningModule = torch.nn.Module
_pl.seed_everything = lambda *a, **k: None
_plu = types.ModuleType("pytorch_lightning.utilities")
_plu.rank_zero_only = lambda f: f
sys.modules.setdefault("pytorch_lightning", _pl)
sys.modules.setdefault("pytorch_lightning.utilities", _plu)


class _Bar:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def update(self, *a):
        pass


class _DiffusionPipeline:
    def __init__(self):
        pass

    def register_modules(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    @property
    def _execution_device(self):
        return torch.device("cpu")

    def to(self, *a, **k):
        return self

    def progress_bar(self, *a, **k):
        return _Bar()


_df = types.ModuleType("diffusers")
_df.DiffusionPipeline = _DiffusionPipeline
_df.ConfigMixin = type("ConfigMixin", (), {})
_df.SchedulerMixin = type("SchedulerMixin", (), {})
_df.logging = types.SimpleNamespace(get_logger=lambda name: types.SimpleNamespace(
    info=lambda *a, **k: None, warning=lambda *a, **k: None))
_cu = types.ModuleType("diffusers.configuration_utils")
_cu.ConfigMixin = _df.ConfigMixin
_cu.register_to_config = lambda f: f
_su = types.ModuleType("diffusers.schedulers.scheduling_utils")
_su.SchedulerMixin = _df.SchedulerMixin
_ut = types.ModuleType("diffusers.utils")
_ut.BaseOutput = type("BaseOutput", (), {})
_ut.logging = _df.logging
_sc = types.ModuleType("diffusers.schedulers")
for name, mod in (("diffusers", _df), ("diffusers.configuration_utils", _cu), ("diffusers.schedulers", _sc),
                  ("diffusers.schedulers.scheduling_utils", _su), ("diffusers.utils", _ut)):
    sys.modules.setdefault(name, mod)
for missing in ("kornia", "open_clip", "imageio", "decord", "omegaconf"):
    if missing not in sys.modules:
        try:
            __import__(missing)
        except Exception:
            sys.modules[missing] = types.ModuleType(missing)

