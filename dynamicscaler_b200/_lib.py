"""ctypes binding of libdynscaler_sm100.so (C ABI: include/dynscaler.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``csrc/Makefile``.  If it
is missing every call raises :class:`LibraryMissing` -- there is no other code path.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdynscaler_sm100.so")

DS_F32, DS_BF16, DS_F16 = 0, 1, 2
DS_MASK_NONE, DS_MASK_PLANE, DS_MASK_FULL = 0, 1, 2
DS_MAX_PAYLOADS = 4
ABI_VERSION = 7

_vp, _i, _ll, _f = C.c_void_p, C.c_int, C.c_longlong, C.c_float

# name -> argtypes; mirrors include/dynscaler.h one to one (tests/test_abi.py checks it
# against the header text).
PROTOTYPES = {
    "ds_abi_version": [],
    "ds_error_string": [_i],
    "ds_peer_table_size": [],
    "ds_pm_peer_sync": [_vp, _i, _i, _vp],
    "ds_tune": [_i, _ll],
    "ds_ring_gather": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "ds_ring_scatter": [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "ds_ring_merge_plan": [_vp, _vp, _i, _i, _i, _i, _i, _vp],
    "ds_gather_nearest": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "ds_gather_bilinear": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "ds_merge_nearest": [_vp, _vp, _i, _vp, _i, _i, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "ds_merge_nearest_peer": [_vp, _vp, _i, _vp, _i, _i, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _vp],
    "ds_merge_nearest_ddim": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _i,
                              _i, _i, _vp, _i, _vp],
    "ds_renoise_mix": [_vp, _vp, _vp, _vp, _f, _f, _f, _f, _ll, _ll, _i, _i, _vp],
    "ds_merge_bilinear": [_vp, _vp, _i, _vp, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "ds_pm_gather": [_vp, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _vp],
    "ds_pm_merge_nearest": [_vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _vp],
    "ds_pm_merge_bilinear": [_vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _vp],
    "ds_pm_merge_bilinear_tiled": [_vp, _vp, _i, _vp, _i, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _vp],
    "ds_mix_mask": [_vp, _vp, _vp, _vp, _f, _f, _ll, _i, _i, _vp],
    "ds_pm_merge_nearest_ddim": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _vp, _i, _i,
                                 _i, _i, _i, _i, _i, _i, _i, _vp, _i, _vp],
    "ds_pm_merge_bilinear_tiled_ddim": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _i,
                                        _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _vp],
    "ds_pm_gather_renoise": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _vp],
    "ds_resize_planes": [_vp, _vp, _ll, _i, _i, _i, _i, _i, _i, _i, _vp],
    "ds_select_merged": [_vp, _vp, _vp, _ll, _ll, _ll, _i, _vp],
    "ds_tile_accumulate": [_vp, _vp, _vp, _ll, _ll, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "ds_tile_normalise": [_vp, _vp, _vp, _ll, _ll, _ll, _vp],
    "ds_repeat_renoise": [_vp, _vp, _vp, _f, _f, _ll, _i, _ll, _i, _vp],
}


PEER_DEFERRED_PUBLISH = [False]   # mirror of the library's DS_TUNE_PEER_MERGE_PUBLISH switch (set through PanoramaStepper)
DS_MAX_PEERS = 8
DS_PEER_AFTER_GATHER, DS_PEER_AFTER_MERGE, DS_PEER_WAIT_MERGES = 0, 1, 2


class PeerTable(C.Structure):
    """Mirror of ``ds_peer_table_t`` (include/dynscaler.h); tests/test_abi.py checks the layout
    against the header."""
    _fields_ = [("merges_done", C.c_int32 * DS_MAX_PEERS), ("gathers_done", C.c_int32 * DS_MAX_PEERS),
                ("my_merges", C.c_int32), ("my_gathers", C.c_int32), ("issued_gathers", C.c_int32),
                ("issued_merges", C.c_int32), ("error", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32),
                ("reserved", C.c_int32),
                ("pm", (C.c_uint64 * DS_MAX_PAYLOADS) * DS_MAX_PEERS), ("mask", C.c_uint64 * DS_MAX_PEERS),
                ("ctl", C.c_uint64 * DS_MAX_PEERS), ("mc_pm", C.c_uint64 * DS_MAX_PAYLOADS), ("mc_mask", C.c_uint64)]


class LibraryMissing(RuntimeError):
    pass


class _Lib:
    """Lazy handle: loads on first attribute access so that importing the package on a
    machine without the built library (or without a GPU) still works."""

    def __init__(self):
        self._dll = None
        self.launches = 0          # kernels launched through this binding (bench.py reports it)

    def load(self):
        if self._dll is None:
            if not os.path.exists(LIB_PATH):
                raise LibraryMissing(
                    f"{LIB_PATH} not built -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(nvcc, sm_100a).  dynamicscaler_b200 has no CPU or PyTorch fallback.")
            dll = C.CDLL(LIB_PATH)
            for name, argtypes in PROTOTYPES.items():
                fn = getattr(dll, name)
                fn.argtypes = argtypes
                fn.restype = C.c_char_p if name == "ds_error_string" else C.c_int
            if dll.ds_abi_version() != ABI_VERSION:
                raise LibraryMissing(f"{LIB_PATH}: ABI {dll.ds_abi_version()} != expected {ABI_VERSION}; rebuild")
            self._dll = dll
        return self._dll

    def __getattr__(self, name):
        if name.startswith("ds_"):
            fn = getattr(self.load(), name)
            self.__dict__[name] = fn           # next lookup is a plain instance attribute
            return fn
        raise AttributeError(name)

    def check(self, rc: int, what: str = ""):
        self.launches += 1
        if rc != 0:
            msg = self.load().ds_error_string(rc).decode()
            raise RuntimeError(f"{what or 'dynscaler'} failed ({rc}): {msg}")


lib = _Lib()


def dtype_code(dtype: torch.dtype) -> int:
    if dtype == torch.float32:
        return DS_F32
    if dtype == torch.bfloat16:
        return DS_BF16
    if dtype == torch.float16:                 # the reference's use_fp16 mode (gen_pano_360.py:191-213)
        return DS_F16
    raise TypeError(f"dynamicscaler_b200 supports float32, bfloat16 and float16 latents, got {dtype}")


def require_cuda(t: torch.Tensor, what: str = "tensor"):
    if not t.is_cuda:
        raise RuntimeError(f"dynamicscaler_b200: {what} must live on a CUDA device (got {t.device}); "
                           "this package has no CPU path")


class _NoGuard:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def on(device):
    """``with on(dev):`` -- make ``dev`` current for a launch; free when it already is (the usual
    case: switching devices costs several microseconds per call)."""
    idx = device.index
    if idx is None or idx == torch.cuda.current_device():
        return _NO_GUARD
    return torch.cuda.device(device)


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_ptr(device) -> int:
    """cudaStream_t of torch's current stream on ``device`` (the raw C accessor when torch has it: the
    unmodified call sites make ~300 launches per step, so every microsecond of host work counts)."""
    if _raw_stream is not None:
        idx = device.index
        return _raw_stream(torch.cuda.current_device() if idx is None else idx)
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def ptr_array(tensors):
    """Host array of device pointers (the ``h_panos`` / ``h_payloads`` arguments)."""
    arr = (C.c_void_p * max(len(tensors), 1))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr
