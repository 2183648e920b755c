"""The C-ABI shared library loads and exports every symbol include/seqrec_b200.h declares (no compute calls: CPU box)."""
import ctypes
import os
import re

import pytest

from seq_recommendations_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "seqrec_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\bint\s+(seqrec_\w+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return ctypes.CDLL(_lib.LIB_PATH)


def test_header_declares_the_expected_surface():
    syms = header_symbols()
    assert len(syms) >= 20
    for must in ("seqrec_gather_rows", "seqrec_scatter_add_rows", "seqrec_rnn_forward", "seqrec_rnn_backward",
                 "seqrec_ce_forward", "seqrec_ce_backward", "seqrec_adagrad", "seqrec_topk"):
        assert must in syms


def test_library_exports_every_declared_symbol(lib):
    for s in header_symbols():
        assert hasattr(lib, s), "libseqrec_b200.so does not export %s" % s


def test_python_binding_covers_the_header():
    assert sorted(_lib.SIGNATURES) == header_symbols()


def test_abi_version_and_no_torch_types(lib):
    lib.seqrec_abi_version.restype = ctypes.c_int
    assert lib.seqrec_abi_version() >= 1
    text = open(os.path.join(ROOT, "include", "seqrec_b200.h")).read()
    assert "torch" not in text.lower().replace("pytorch allocates", "") and "at::" not in text


def test_missing_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from seq_recommendations_b200 import engine
    with pytest.raises(_lib.SeqrecError):
        engine.HotPath("GRU", "tanh", 10, 8, 10)
