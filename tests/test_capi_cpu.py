"""CPU-side checks of the C-ABI boundary: the library builds, loads, exports every symbol that
include/openvis_b200.h declares, and refuses to run without an sm_100 device (no fallback)."""
import os
import re

import pytest
import torch

from openvis_b200 import _lib as L
from openvis_b200 import build as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    B.build()
    return L.load()


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "openvis_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ovis_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    # and the python binding knows every one of them
    assert sorted(L.SIGNATURES) == names


def test_version_and_plan(lib):
    assert lib.ovis_version() == 100
    splits, q_pad, o_n, ml_n = L.xattn_plan(1, 100, 529920)
    assert splits >= 1 and q_pad == 128
    # (max, sum) partials + the tile-skip bitmap (512 words per group and query tile)
    assert o_n == splits * 8 * 128 * 32 and ml_n == splits * 8 * 128 * 2 + 512
    s2, qp2, _, _ = L.xattn_plan(36, 200, 920)
    assert s2 >= 1 and qp2 == 256


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_device_no_fallback(lib):
    assert lib.ovis_device_check() == 2          # OVIS_ERR_ARCH
    with pytest.raises(L.OvisError):
        L.device_check()
    with pytest.raises(L.OvisError):
        L.nchw_to_tokens_f16(torch.zeros(1, 256, 2, 2))     # CPU tensor: refused, never computed on the host
