"""CPU: the C-ABI library loads and exports every symbol include/unomol_b200.h declares; without a GPU the
compute entry points fail loudly instead of falling back."""
import ctypes
import os
import re
import numpy as np
import pytest
from conftest import ROOT, golden_input


def test_header_symbols_are_exported():
    from unomol_b200 import capi
    hdr = open(os.path.join(ROOT, "include", "unomol_b200.h")).read()
    declared = set(re.findall(r"\b(unomol_b200_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(capi.lib, name), "missing export %s" % name
    assert declared == set(capi.EXPORTS)


def test_struct_layouts():
    from unomol_b200 import capi
    assert ctypes.sizeof(capi.TwoInt) == 24          # reference TwoInts, TwoElectronInts.hpp:20-23
    assert capi.lib.unomol_b200_strerror(-2).decode().startswith("CUDA")


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from unomol_b200 import capi
    from unomol_b200.basis import Basis
    b = Basis.from_patin(golden_input("3g.h2o"))
    with pytest.raises(capi.UnomolError):
        capi.Handle(b)


def test_product_does_not_import_oracle():
    """the product path must never route through oracle/ (tier rule 3)"""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "unomol_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b|#include\s+\"[^\"]*oracle", txt, flags=re.M):
                    bad.append(f)
    assert not bad, bad


def test_basis_reader_matches_oracle(oracle):
    from unomol_b200.basis import Basis
    for name in ["3g.h2o", "631.nh3", "dh95.co2", "tz2p.sf6"]:
        b = Basis.from_patin(golden_input(name))
        o = oracle.basis(golden_input(name))
        assert (b.nshell, b.nbf, b.ncen) == (o.nshell, o.nbf, o.ncen)
        assert np.array_equal(b.off, o.off) and np.array_equal(b.lv, o.lv)
        np.testing.assert_allclose(b.coef, o.coef, rtol=1e-14)
        np.testing.assert_allclose(b.xyz, o.xyz, rtol=0, atol=0)


def test_water_cluster_shape():
    from unomol_b200.basis import water_cluster
    b = water_cluster(154)
    assert (b.nbf, b.nshell, b.nelec) == (2002, 1386, 1540)
    b2 = water_cluster(154)
    assert np.array_equal(b.xyz, b2.xyz)                        # seeded
