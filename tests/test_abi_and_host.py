"""CPU tests: the C-ABI library loads and exports every symbol the header declares; host-side logic."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def header_symbols():
    text = open(os.path.join(ROOT, "include", "naiveb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nb200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_header_symbol(pkg):
    L = pkg._lib.load()
    names = header_symbols()
    assert len(names) >= 30
    for name in names:
        assert hasattr(L, name), f"{name} declared in include/naiveb200.h but not exported"
    # and the ctypes table binds exactly the header's functions
    assert sorted(pkg._lib.SIGNATURES) == names


def test_library_is_sm100a_only():
    import subprocess
    lib = os.path.join(ROOT, "naivedynamics.jl_b200", "libnaiveb200.so")
    out = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_device(pkg):
    L = pkg._lib.load()
    if L.nb200_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.NB200Error, match="no CPU fallback"):
        pkg.Handle(1000)


def test_product_never_touches_the_oracle():
    pkgdir = os.path.join(ROOT, "naivedynamics.jl_b200")
    for dirpath, _, files in os.walk(pkgdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl", "Makefile")):
                text = open(os.path.join(dirpath, f)).read()
                assert "nd_oracle" not in text and "oracle/" not in text, f"{f} references the oracle"


def test_spheres_bvh_specs_validation(pkg):
    # BVHTraverse.jl:74-80, same messages
    with pytest.raises(ValueError, match="evenly divides"):
        pkg.SpheresBVHSpecs(neighbor_distance=0.1, atom_count=10, floattype=np.float32, atomsperleaf=3)
    with pytest.raises(ValueError, match="more than one leaf"):
        pkg.SpheresBVHSpecs(neighbor_distance=0.1, atom_count=8, floattype=np.float32, atomsperleaf=8)
    with pytest.raises(ValueError, match="not implemented"):  # Float64 -> mortoncodes! error (:285-287)
        pkg.SpheresBVHSpecs(neighbor_distance=0.1, atom_count=8, floattype=np.float64, atomsperleaf=1)
    s = pkg.SpheresBVHSpecs(neighbor_distance=0.1, atom_count=5000, floattype=np.float32, atomsperleaf=4)
    assert (s.leaves_count, s.branches_count) == (1250, 1249)
    assert s.neighbor_distance.dtype == np.float32


def test_collect_objects_shapes_and_rules(pkg):
    c = pkg.GenericRandomCollector(objectnumber=100, minDim=(0.0, 0.0, 0.0), maxDim=(1.0, 2.0, 3.0), temperature=0.01,
                                   randomvelocity=False, minmass=1.0, maxmass=5.0, minimumdistance=0.001,
                                   mincharge=-1e-9, maxcharge=1e-9, seed=3)
    s = pkg.collect_objects(c)
    assert s.position.shape == (100, 3) and s.position.dtype == np.float32
    assert np.all(s.position >= 0) and np.all(s.position[:, 1] < 2) and s.position[:, 2].max() > 1
    # !randomvelocity: v = T/N * 3N * kb / m  (MDInput.jl:329-333)
    assert np.allclose(s.velocity[:, 0], 0.01 * 3 / s.mass, rtol=1e-6)
    assert s.index.tolist() == list(range(1, 101)) and np.all(s.force == 0)
    p = pkg.generate_positions(c)
    assert p.shape == (100, 3)


def test_pairlist_sorted_canonicalises(pkg):
    pl = pkg.PairList(np.array([3, 1, 2], np.int32), np.array([1, 2, 1], np.int32), np.array([.3, .1, .2], np.float32))
    s = pl.sorted()
    assert s.a.tolist() == [1, 1, 1] and s.b.tolist() == [2, 2, 3]
    assert len(pl) == 3 and pl[0] == (3, 1, np.float32(.3))
