"""CPU: the C-ABI library loads and exports every symbol include/ttb.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "ttb.h")).read()
    return sorted(set(re.findall(r"\b(ttb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    import __graft_entry__ as ge
    ge.build()
    libpath = os.path.join(ROOT, "tortoise_tts_b200", "libttb.so")
    assert os.path.exists(libpath)
    lib = ctypes.CDLL(libpath)
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), n
    lib.ttb_version.restype = ctypes.c_int
    assert lib.ttb_version() >= 100


def test_binding_lists_every_symbol():
    from tortoise_tts_b200 import lib as L
    assert sorted(L.SYMBOLS) == _declared()


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tortoise_tts_b200.api import TextToSpeech
    with pytest.raises(RuntimeError):
        TextToSpeech(state_dicts={})
