"""The C-ABI library loads and exports every symbol include/cilqr_b200.h declares; without a GPU
it refuses to create a handle instead of falling back to a CPU path."""
import ctypes as C
import os
import re

import pytest

import cilqr_b200 as cb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "cilqr_b200.h")).read()
    return sorted(set(re.findall(r"\b(cilqr_b200_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported():
    lib = cb.load_library()
    names = _declared()
    assert len(names) >= 20
    assert sorted(cb.EXPORTS) == names
    for n in names:
        assert hasattr(lib, n), n


def test_params_struct_layout():
    assert C.sizeof(cb.CilqrParams) == 30 * 8 + 4 * 4
    p = cb.CilqrParams.from_dict(cb.get_scenario("two_straight").params)
    assert p.dt == 0.1 and p.reference_point == 0 and p.max_iter == 100


def test_no_gpu_means_no_handle():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    scn = cb.get_scenario("two_straight")
    with pytest.raises(cb.CilqrError) as e:
        cb.BatchSolver([cb.scenario.template_data(scn)], 4, 30, 3)
    assert e.value.code == -4  # CILQR_ERR_NO_DEVICE: the product has no CPU fallback
    assert "no CPU fallback" in str(e.value) or "sm_" in str(e.value)


def test_argument_validation_without_gpu():
    lib = cb.load_library()
    h = C.c_void_p()
    p = cb.CilqrParams.from_dict(cb.get_scenario("two_straight").params)
    assert lib.cilqr_b200_create(None, 0, 4, 30, 3, 0, C.byref(h)) == -1
    assert lib.cilqr_b200_create(C.byref(p), 0, 0, 30, 3, 0, C.byref(h)) == -1
    assert lib.cilqr_b200_create(C.byref(p), 0, 4, 0, 3, 0, C.byref(h)) == -1
    assert lib.cilqr_b200_create(C.byref(p), 0, 4, 30, 3, 7, C.byref(h)) == -1
    assert b"dtype" in lib.cilqr_b200_last_error()
    assert lib.cilqr_b200_destroy(None) == 0


def test_product_does_not_touch_the_oracle():
    """The shipped path must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "toy-example-of-ilqr_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.lower() or f in ("__init__.py",) and "oracle" not in txt, (dirpath, f)
    assert "oracle" not in open(os.path.join(ROOT, "include", "cilqr_b200.h")).read().lower()
