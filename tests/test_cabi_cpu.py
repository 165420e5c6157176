"""CPU-only checks of the boundary: the C-ABI library loads, exports every
symbol include/lite_b200.h declares, and fails loudly without a GPU."""
import os
import re

import pytest

import lite_b200
from lite_b200.capi import SYMBOLS
from lite_b200.host import HOST_SYMBOLS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _built():
    return os.path.exists(lite_b200.lib_path())


def test_header_symbols_are_bound():
    hdr = open(os.path.join(ROOT, "include", "lite_b200.h")).read()
    declared = set(re.findall(r"\b(lite_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(SYMBOLS)


def test_host_header_symbols_are_bound():
    hdr = open(os.path.join(ROOT, "include", "lite_b200_host.h")).read()
    declared = set(re.findall(r"\b(lite_host_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(HOST_SYMBOLS)


@pytest.mark.skipif(not _built(), reason="libliteb200.so not built (run __graft_entry__.build())")
def test_library_exports_every_symbol():
    L = lite_b200.load_library()
    for s in SYMBOLS + HOST_SYMBOLS:
        assert hasattr(L, s), s


@pytest.mark.skipif(not _built(), reason="libliteb200.so not built")
def test_ensemble_calls_fail_loudly_without_a_context_or_gpu():
    """No CPU path behind lite_smf_*: without a context every call is an error."""
    L = lite_b200.load_library()
    assert L.lite_smf_run(None, 1, 1, None) != 0
    assert L.lite_smf_create(None, 8, 16, 0.0, 0.0, 1.0, 1.0, 5) != 0
    assert b"context" in L.lite_last_error() or b"lite_smf_create" in L.lite_last_error()


@pytest.mark.skipif(not _built(), reason="libliteb200.so not built")
def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(lite_b200.LiteError, match="no CPU fallback"):
        lite_b200.Context(0)


def test_shard_ranges_partition():
    for n in (0, 1, 7, 10**8 + 3):
        for g in (1, 2, 3, 8):
            tot = 0
            prev_end = 0
            for r in range(g):
                j0, c = lite_b200.shard_range(n, r, g)
                assert j0 == prev_end
                prev_end = j0 + c
                tot += c
            assert tot == n
