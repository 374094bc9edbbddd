"""The C-ABI library loads on a GPU-less box and exports exactly what include/dynscaler.h
declares, with the argument lists the ctypes binding uses.  No compute calls."""
import ctypes
import os
import re

import pytest

from dynamicscaler_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "dynscaler.h")).read()

CTYPE = {"int": ctypes.c_int, "long long": ctypes.c_longlong, "float": ctypes.c_float}


def declared():
    text = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(int|const char\*)\s+(ds_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        args = [a.strip() for a in m.group(3).split(",")]
        args = [] if args == ["void"] else args
        types = []
        for a in args:
            a = re.sub(r"\s+", " ", a)
            if "*" in a:
                types.append(ctypes.c_void_p)
            else:
                types.append(CTYPE[" ".join(a.split(" ")[:-1])])
        out[m.group(2)] = types
    return out


def test_header_matches_binding():
    decl = declared()
    assert set(decl) == set(_lib.PROTOTYPES), set(decl) ^ set(_lib.PROTOTYPES)
    for name, types in decl.items():
        assert types == _lib.PROTOTYPES[name], name


def test_library_exports_every_symbol():
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    dll = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared():
        assert hasattr(dll, name), name
    assert _lib.lib.ds_abi_version() == _lib.ABI_VERSION
    assert _lib.lib.ds_error_string(-4) == b"dynscaler: illegal ring window"


def test_argument_errors_need_no_gpu():
    """Argument validation happens before any launch, so it is observable here."""
    L = _lib.lib
    assert L.ds_ring_gather(None, None, None, 1, 0, 0, 0, 1, 1, 8, 16, 1, 4, 4, 0, None) == -1
    one = ctypes.c_void_p(16)
    assert L.ds_ring_gather(one, one, None, 2, 0, 0, 0, 1, 1, 8, 16, 1, 4, 4, 0, None) == -2      # n>1 needs d_wins
    assert L.ds_ring_gather(one, one, None, 1, 40, 0, 0, 1, 1, 8, 16, 1, 4, 4, 0, None) == -4     # left >= 2W
    assert L.ds_ring_scatter(one, one, 0, 0, 0, 1, 1, 8, 16, 1, 9, 4, 0, None) == -4              # h > H on a write
    assert L.ds_ring_gather(one, one, None, 1, 0, 0, 0, 1, 1, 8, 16, 1, 4, 4, 7, None) == -3      # dtype
    assert L.ds_gather_nearest(one, one, None, None, 0, 1, 1, 1, 8, 1, 4, 0, None) == -1
    assert L.ds_mix_mask(one, one, one, one, 0.5, 0.5, 10, 3, 0, None) == -2                      # n % mask_hw


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    fresh = _lib._Lib()
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.LibraryMissing):
        fresh.ds_abi_version


def test_cpu_tensors_are_rejected():
    import torch
    from dynamicscaler_b200.shift_window_utils import RingLatent
    from dynamicscaler_b200.panorama_tensor_utils import PanoramaLatentProxy
    from dynamicscaler_b200.tensor_utils import mix_latents_with_mask
    with pytest.raises(RuntimeError, match="no CPU path"):
        RingLatent(torch.zeros(1, 1, 2, 8, 16)).get_window_latent(0, 4, 0, 4)
    with pytest.raises(RuntimeError, match="no CPU path"):
        PanoramaLatentProxy(torch.zeros(1, 1, 2, 8, 16)).get_view_tensor_no_interpolate(90, 0, 0, 4, 4)
    with pytest.raises(RuntimeError, match="no CPU path"):
        mix_latents_with_mask(torch.zeros(1, 1, 2, 4, 4), torch.zeros(1, 1, 2, 4, 4), torch.zeros(1, 4, 4), 0.5)


def test_reference_assertions_fire_before_any_launch():
    import torch
    from dynamicscaler_b200.shift_window_utils import RingLatent, RingImageTensor
    from dynamicscaler_b200.panorama_tensor_utils import PanoramaTensor, PanoramaLatentProxy
    from dynamicscaler_b200.tensor_utils import mix_latents_with_mask
    r = RingLatent(torch.zeros(1, 1, 2, 8, 16))
    with pytest.raises(AssertionError, match="Invalid pos_left"):
        r.get_window_latent(5, 5)
    with pytest.raises(AssertionError, match="Invalid pos_top"):
        r.get_window_latent(0, 4, 3, 17)
    with pytest.raises(AssertionError, match="Invalid frame_begin"):
        r.get_window_latent(0, 4, 0, 4, 2, 5)
    with pytest.raises(AssertionError, match="not legal"):
        RingLatent(torch.zeros(1, 2, 8, 16))
    with pytest.raises(AssertionError):
        PanoramaTensor(torch.zeros(3, 8, 15))                       # W != 2H
    with pytest.raises(AssertionError):
        PanoramaLatentProxy(torch.zeros(8, 16))
    with pytest.raises(AssertionError, match="RingImageTensor"):
        RingImageTensor(None, torch.zeros(3, 8, 16), height=8, width=12)
    with pytest.raises(NotImplementedError):
        mix_latents_with_mask(torch.zeros(1, 1, 2, 4, 4), torch.zeros(1, 1, 2, 4, 4), torch.zeros(4, 4), 0.5)


def test_peer_table_layout_matches_header(tmp_path):
    """ctypes mirror of ds_peer_table_t vs the C header (compiled with gcc) and the library."""
    import subprocess
    src = tmp_path / "off.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "dynscaler.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(ds_peer_table_t),'
                   'offsetof(ds_peer_table_t, gathers_done), offsetof(ds_peer_table_t, my_merges),'
                   'offsetof(ds_peer_table_t, error), offsetof(ds_peer_table_t, pm),'
                   'offsetof(ds_peer_table_t, mask), offsetof(ds_peer_table_t, mc_mask));return 0;}\n')
    exe = tmp_path / "off"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    T = _lib.PeerTable
    want = [ctypes.sizeof(T), T.gathers_done.offset, T.my_merges.offset, T.error.offset, T.pm.offset,
            T.mask.offset, T.mc_mask.offset]
    assert got == want
    assert _lib.lib.ds_peer_table_size() == ctypes.sizeof(T)
