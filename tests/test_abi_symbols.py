"""The C-ABI library loads on a CPU-only box and exports every symbol include/kpopcount.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_functions():
    with open(os.path.join(ROOT, "include", "kpopcount.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(kpc_[a-z0-9_]+)\s*\(", text)
    return sorted(set(n for n in names if n != "kpc_sink_fn"))


def test_library_exports_every_declared_symbol():
    from kpop_b200 import _native
    assert os.path.exists(_native.LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(_native.LIB_PATH)
    funcs = declared_functions()
    assert len(funcs) >= 25
    for name in funcs:
        assert hasattr(lib, name), f"{name} is declared in include/kpopcount.h but not exported"
    assert set(funcs) == set(_native.PROTOTYPES), "the ctypes binding and the header disagree"


def test_backend_is_cuda_and_there_is_no_fallback():
    """Without a GPU the product must fail loudly (KPC_E_CUDA), never count on the CPU."""
    import torch
    from kpop_b200 import KMerCounter, KPopCountError, _native
    assert _native.load().kpc_backend() == b"cuda"
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(KPopCountError) as e:
        KMerCounter(k=12, label="x")
    assert e.value.code == _native.KPC_E_CUDA


def test_k_range_is_checked_at_creation():
    from kpop_b200 import Content, KMerCounter, KPopCountError, _native
    for content, k in ((Content.DNA_ds, 31), (Content.DNA_ss, 31), (Content.Protein, 13)):
        with pytest.raises(KPopCountError) as e:
            KMerCounter(k=k, content=content, label="x")
        assert e.value.code == _native.KPC_E_K_RANGE


def test_spectra_filename():
    from kpop_b200 import spectra_filename
    assert spectra_filename("pre") == "pre.KPopSpectra.txt"
    assert spectra_filename("/dev/stdout") == "/dev/stdout"
    assert spectra_filename("/devx") == "/devx.KPopSpectra.txt"
