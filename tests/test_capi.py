"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol
include/psi_b200.h declares, and refuses to compute without a device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import psi_b200 as P
from psi_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "psi_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(psi_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    names = declared_symbols()
    assert len(names) >= 30
    L = ctypes.CDLL(P.lib_path())
    for n in names:
        assert hasattr(L, n), "libpsi_b200.so does not export " + n
    assert sorted(capi.SYMBOLS) == names, "ctypes binding and header disagree"
    assert b"sm_100a" in P.lib().psi_version()


def test_params_struct_layout_matches_oracle_mirror():
    from oracle.params_ref import PsiParams as OraclePsiParams
    assert ctypes.sizeof(P.PsiParams) == ctypes.sizeof(OraclePsiParams)
    for (n1, t1), (n2, t2) in zip(P.PsiParams._fields_, OraclePsiParams._fields_):
        assert n1 == n2 and ctypes.sizeof(t1) == ctypes.sizeof(t2)


def test_params_generate_validation():
    with pytest.raises(P.PsiError):
        P.params_generate(1000, 4296540161, 3)         # not a power of two
    with pytest.raises(P.PsiError):
        P.params_generate(16384, 65537 * 3, 3)         # not prime / not 1 mod 2N
    with pytest.raises(P.PsiError):
        P.params_generate(16384, 4296540161, 3, 9)     # sizeQ above PSI_MAX_LIMBS
    p = P.params_generate(16384, 4296540161, 3)
    assert (p.N, p.L, p.Lp) == (16384, 4, 4)


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="GPU present")
def test_no_cpu_fallback():
    p = P.params_generate(1024, 4296540161, 2)
    with pytest.raises(P.PsiError) as ei:
        P.CryptoContext(p)
    assert ei.value.status == capi.PSI_ERR_NO_DEVICE
    assert "no CPU path" in str(ei.value)
    out = ctypes.c_double()
    assert P.lib().psi_bench_imad_peak(0, ctypes.byref(out)) != 0


def test_product_does_not_import_oracle():
    """The product path must never route through the oracle."""
    pkg = os.path.join(ROOT, "nested-hashing-psi_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|#include\s+\".*oracle|libpsi_oracle|orc_[a-z_]+\(", text,
                                     flags=re.M), f
