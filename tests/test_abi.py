"""CPU tests of the drop-in boundary: the shared library loads, exports every symbol the header
declares, and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from nquant_android_b200 import _build, _lib
    _build.build()
    return _lib.load()


def header_functions():
    text = open(os.path.join(ROOT, "include", "nquant_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nq_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    names = header_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/nquant_b200.h but not exported"


def test_binding_lists_every_symbol():
    from nquant_android_b200 import _lib
    assert sorted(_lib.SYMBOLS) == header_functions()


def test_image_info_mirror_matches(lib):
    from nquant_android_b200 import _lib
    assert lib.nq_sizeof_image_info() == ctypes.sizeof(_lib.ImageInfo)


def test_header_cites_reference_lines():
    text = open(os.path.join(ROOT, "include", "nquant_b200.h")).read()
    for cite in ["PnnQuantizer.java:409", "PnnQuantizer.java:35", "PnnLABQuantizer.java:24", "PnnQuantizer.java:458",
                 "GilbertCurve.java:282"]:
        assert cite in text


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert lib.nq_device_count() == 0
    assert not lib.nq_create(0)
    assert b"no CPU fallback" in lib.nq_last_error()
    from nquant_android_b200.quantizer import PnnQuantizer, NQuantError
    q = PnnQuantizer(np.zeros(16, dtype=np.uint32), 4, 4)
    with pytest.raises(NQuantError):
        q.convert(16, True)


def test_gilbert_order_argument_errors(lib):
    out = (ctypes.c_uint32 * 4)()
    assert lib.nq_gilbert_order(0, 4, out) == -2
    assert lib.nq_gilbert_order(2, 2, None) == -2
    assert lib.nq_gilbert_order(2, 2, out) == 0
    assert list(out) == [0, 2, 3, 1]


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "nquant_android_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in text and "nq_oracle" not in text and "oracle/" not in text, f
