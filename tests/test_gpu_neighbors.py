"""GPU parity tests of the neighbour search against the oracle (bit-exact pair set and distances)."""
import numpy as np
import pytest

from conftest import uniform_positions

pytestmark = pytest.mark.gpu


def assert_same_pairs(oracle, got, ref):
    """got, ref: (a, b, d).  Canonical (min,max) pair set and Float32 d bit-identical."""
    cg, cr = oracle.canonical(*got), oracle.canonical(*ref)
    assert len(cg[0]) == len(cr[0]), (len(cg[0]), len(cr[0]))
    assert np.array_equal(cg[0], cr[0]) and np.array_equal(cg[1], cr[1])
    assert np.array_equal(cg[2].view(np.uint32), cr[2].view(np.uint32))


def assert_same_oriented(got, ref):
    """The (a, b) ORIENTATION must match the reference traversal too (a = earlier in its sort)."""
    g = np.stack([got[0], got[1]], 1)
    r = np.stack([ref[0], ref[1]], 1)
    g = g[np.lexsort((g[:, 1], g[:, 0]))]
    r = r[np.lexsort((r[:, 1], r[:, 0]))]
    assert np.array_equal(g, r)


def search(pkg, x, r, apl=4):
    spec = pkg.SpheresBVHSpecs(neighbor_distance=r, atom_count=len(x), floattype=np.float32, atomsperleaf=apl)
    pl = pkg.leafbuild_traverse_bvh(x, spec)
    return pl.a, pl.b, pl.d


def test_position8_golden(pkg, oracle, golden8):
    # test/BVHTraverse.jl:153-209
    p8, r = golden8["position8"], golden8["neighbor_distance"]
    got = search(pkg, p8, r, apl=1)
    assert len(got[0]) == golden8["expected_pairs"]
    ref = oracle.build_traverse_bvh(p8, r, 1)
    assert_same_pairs(oracle, got, ref)
    assert_same_oriented(got, ref)
    naive = oracle.brute_force(p8, r, "sqrt")  # threshold_pairs(unique_pairs(position8), r)
    assert_same_pairs(oracle, got, naive)
    is_paired = np.zeros(9, bool)
    is_paired[got[0]] = True
    is_paired[got[1]] = True
    assert is_paired[1:].sum() == 8


@pytest.mark.parametrize("d", [1, 2, 3, 4, 5])
def test_pos5000_over_distances(pkg, oracle, d):
    # test/BVHTraverse.jl:222-243 — r up to 1.0: rows of thousands of neighbours, many segments per leaf
    x = uniform_positions(5000, 5000)
    r = np.float32(0.00001 * 10 ** d)
    got = search(pkg, x, r)
    assert_same_pairs(oracle, got, oracle.brute_force(x, r, "d2"))
    if d <= 4:
        ref = oracle.leafbuild_traverse_bvh(x, r, 4, nthreads=4)
        assert_same_pairs(oracle, got, ref)
        assert_same_oriented(got, ref)


@pytest.mark.parametrize("apl", [1, 2, 4, 5, 10, 20, 1000])
def test_pos5000_over_atomsperleaf(pkg, oracle, apl):
    # test/BVHTraverse.jl:246-265: result independent of atomsperleaf
    x = uniform_positions(5000, 5001)
    got = search(pkg, x, 0.1, apl)
    ref = oracle.build_traverse_bvh(x, 0.1, 1, nthreads=4) if apl == 1 else oracle.leafbuild_traverse_bvh(x, 0.1, apl, nthreads=4)
    assert_same_pairs(oracle, got, ref)
    assert_same_oriented(got, ref)


@pytest.mark.parametrize("n", [10, 100, 1000, 2000, 3000, 20000])
def test_over_sizes(pkg, oracle, n):
    # test/BVHTraverse.jl:268-287
    x = uniform_positions(n, n)
    got = search(pkg, x, 0.1, 5)
    ref = oracle.leafbuild_traverse_bvh(x, 0.1, 5, nthreads=4)
    assert_same_pairs(oracle, got, ref)
    assert_same_oriented(got, ref)
    assert_same_pairs(oracle, got, oracle.brute_force(x, 0.1, "d2"))


@pytest.mark.parametrize("n", [2, 3, 31, 32, 33, 63, 64, 65, 97])
def test_ragged_sizes(big_handle, oracle, n):
    x = uniform_positions(n, 1000 + n)
    cnt = big_handle.neighbors(x, 0.4)
    got = big_handle.get_pairs()
    assert cnt == len(got[0])
    assert_same_pairs(oracle, got, oracle.brute_force(x, 0.4, "d2"))


def test_edge_cases(big_handle, oracle):
    rng = np.random.default_rng(1)
    # no pairs at all
    x = uniform_positions(2000, 1)
    assert big_handle.neighbors(x, 1e-7) == 0
    a, b, d = big_handle.get_pairs()
    assert len(a) == 0
    # cutoff 0: d2 < 0 is never true, even for coincident atoms
    xc = np.repeat(uniform_positions(50, 2), 3, axis=0)
    assert big_handle.neighbors(xc, 0.0) == 0
    # coincident atoms, tiny cutoff: each triple gives 3 pairs with d == 0
    assert big_handle.neighbors(xc, 1e-6) == 150
    got = big_handle.get_pairs()
    assert np.all(got[2] == 0)
    assert_same_pairs(oracle, got, oracle.brute_force(xc, 1e-6, "d2"))
    # everything within reach of everything (rows of n-1 entries)
    xa = (0.5 + 0.01 * rng.random((700, 3))).astype(np.float32)
    assert big_handle.neighbors(xa, 0.5) == 700 * 699 // 2
    assert_same_pairs(oracle, big_handle.get_pairs(), oracle.brute_force(xa, 0.5, "d2"))
    # clustered + background (imbalanced leaves), atoms outside the [0,1] box
    xk = np.concatenate([uniform_positions(3000, 3) * 1.5 - 0.25,
                         (0.3 + 0.004 * rng.standard_normal((3000, 3))).astype(np.float32)]).astype(np.float32)
    big_handle.neighbors(xk, 0.01)
    assert_same_pairs(oracle, big_handle.get_pairs(), oracle.brute_force(xk, 0.01, "d2"))
    # atoms on a lattice with the cutoff EXACTLY at a lattice distance (strict <, no FMA)
    g = np.stack(np.meshgrid(*[np.arange(16)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.float32) * np.float32(0.0625)
    for r in (0.0625, 0.125, np.float32(0.0625) * np.float32(np.sqrt(2))):
        big_handle.neighbors(g, r)
        assert_same_pairs(oracle, big_handle.get_pairs(), oracle.brute_force(g, r, "d2"))


def test_pair_set_is_independent_of_the_curve(big_handle, oracle):
    x = uniform_positions(20_000, 21)
    ref = oracle.brute_force(x, 0.05, "d2")
    for curve in (0, 1):
        big_handle.set_curve(curve)
        big_handle.neighbors(x, 0.05)
        assert_same_pairs(oracle, big_handle.get_pairs(), ref)
    big_handle.set_curve(1)


def test_half_and_directed_lists_agree(pkg, big_handle, oracle):
    # NB200_LIST_HALF (default: each pair once, row of the Morton-earlier atom, like the reference's traversal)
    # and NB200_LIST_DIRECTED (each pair in both rows) must export the same pairs, orientation and counts.
    rng = np.random.default_rng(5)
    cases = [(uniform_positions(20_000, 31), 0.05), (uniform_positions(33, 32), 0.4), (uniform_positions(2, 33), 2.0),
             (np.concatenate([uniform_positions(3000, 34), (0.6 + 0.003 * rng.standard_normal((3000, 3)))]).astype(np.float32), 0.012),
             ((0.5 + 0.01 * rng.random((500, 3))).astype(np.float32), 0.5)]
    try:
        for x, r in cases:
            ref = oracle.brute_force(x, r, "d2")
            out = {}
            for mode in (pkg._lib.NB200_LIST_HALF, pkg._lib.NB200_LIST_DIRECTED):
                big_handle.set_list_mode(mode)
                cnt = big_handle.neighbors(x, r)
                got = big_handle.get_pairs()
                st = big_handle.get_stats()
                assert st["list_half"] == mode and st["n_pairs"] == cnt == len(got[0])
                assert st["n_entries"] == (cnt if mode == pkg._lib.NB200_LIST_HALF else 2 * cnt)
                assert_same_pairs(oracle, got, ref)
                out[mode] = (got, big_handle.get_neighbor_counts())
            assert_same_oriented(out[0][0], out[1][0])
            assert np.array_equal(out[0][1], out[1][1])
            full = np.bincount(np.concatenate([ref[0], ref[1]]) - 1, minlength=len(x))
            assert np.array_equal(out[1][1], full)
    finally:
        big_handle.set_list_mode(pkg._lib.NB200_LIST_HALF)


def test_stride4_and_index_base(big_handle, oracle):
    x = uniform_positions(3000, 9)
    x4 = np.concatenate([x, np.full((3000, 1), 7.0, np.float32)], 1)
    n3 = big_handle.neighbors(x, 0.05)
    p3 = big_handle.get_pairs(index_base=1)
    n4 = big_handle.neighbors(x4, 0.05)
    p4 = big_handle.get_pairs(index_base=0)
    assert n3 == n4
    assert_same_pairs(oracle, p3, (p4[0] + 1, p4[1] + 1, p4[2]))


def test_errors(pkg, big_handle):
    x = uniform_positions(100, 1)
    with pytest.raises(pkg.NB200Error, match="n_max"):
        big_handle.neighbors(uniform_positions(1_200_000, 1), 0.001)
    with pytest.raises(pkg.NB200Error, match="cutoff"):
        big_handle.neighbors(x, -1.0)
    h = pkg.Handle(128)
    with pytest.raises(pkg.NB200Error, match="no neighbour list"):
        h.get_pairs()
    h.neighbors(x, 0.3)
    npairs = h.pair_count()
    small = (np.empty(3, np.int32), np.empty(3, np.int32), np.empty(3, np.float32))
    with pytest.raises(pkg.NB200Error) as e:  # caller's buffer too small: two-call protocol
        h.get_pairs(out=small)
    assert e.value.code == pkg._lib.NB200_ERR_CAPACITY and npairs > 3
    h.close()


def test_regrow_protocol(pkg, oracle):
    # a handle created with a tiny pair hint must regrow transparently
    x = uniform_positions(4000, 4)
    h = pkg.Handle(4096, pair_capacity_hint=16)
    cnt = h.neighbors(x, 0.1)
    assert h.get_stats()["regrows"] >= 1
    got = h.get_pairs()
    assert cnt == len(got[0])
    assert_same_pairs(oracle, got, oracle.brute_force(x, 0.1, "d2"))
    h.close()


def test_c1_config_10k(pkg, oracle, golden_digests):
    # BASELINE config 1: 10k uniform points, r = 0.1, atomsperleaf = 4 (BVHBenchSuite.jl:117-120)
    x = uniform_positions(10_000, 20250313)
    got = search(pkg, x, 0.1, 4)
    ref = oracle.leafbuild_traverse_bvh(x, 0.1, 4, nthreads=8)
    assert_same_pairs(oracle, got, ref)
    assert_same_oriented(got, ref)
    g, d = golden_digests["c1_10k_uniform_r0.1"], oracle.digest_pairs(*got)   # and the committed golden vector
    assert (d["count"], d["xor"], d["sum"]) == (g["count"], g["xor"], g["sum"])


def test_full_size_1m_digest(big_handle, oracle, golden_digests):
    # BASELINE config 3 size: 1M atoms, rho* = 0.8, rc = 2.5 sigma -> r = 0.02321.  The O(N^2) oracle cannot
    # run here; the independent O(N) cell-grid search gives count / xor / sum digests of the exact pair set.
    n = 1_000_000
    x = uniform_positions(n, 3)
    r = np.float32(2.5 * (0.8 / n) ** (1 / 3))
    cnt = big_handle.neighbors(x, r)
    a, b, d = big_handle.get_pairs()
    ref = oracle.cellgrid_digest(x, r, per_atom=True)
    got = oracle.digest_pairs(a, b, d)
    assert cnt == ref["count"] == got["count"]
    assert got["xor"] == ref["xor"] and got["sum"] == ref["sum"]
    g = golden_digests["c3_size_1m_uniform_seed3"]   # the committed golden vector of the same input
    assert (got["count"], got["xor"], got["sum"]) == (g["count"], g["xor"], g["sum"])
    assert np.array_equal(big_handle.get_neighbor_counts(), ref["per_atom"])
    # size-independent properties: no self pairs, no duplicates, every d below the cutoff
    assert np.all(a != b) and np.all(d < r)
    key = np.minimum(a, b).astype(np.int64) * (n + 1) + np.maximum(a, b)
    assert len(np.unique(key)) == len(key)
