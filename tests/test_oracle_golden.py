"""CPU tests: the oracle against the reference's own golden vector and test properties."""
import numpy as np
import pytest

from conftest import uniform_positions


def test_position8_codes_and_order(oracle, golden8):
    p8 = golden8["position8"]
    q = np.floor(p8 / np.float32(1 / 1023)).astype(np.int32)  # round(Int32, p / binwidth, RoundDown)
    assert q.tolist() == golden8["quantized"]
    codes = oracle.mortoncodes(p8)
    assert codes.tolist() == golden8["morton_codes"]
    assert oracle.sortperm(codes).tolist() == golden8["sorted_original_ids"]


def test_position8_tree_structure(oracle, golden8):
    # test/BVHTraverse.jl:180-187: last leaf skip == 0, leaves 1..7 skip != 0
    T = oracle.tree(golden8["position8"], golden8["neighbor_distance"], 1, leaf_variant=False)
    assert T["skip"][7] == 0
    assert all(T["skip"][i] != 0 for i in range(7))
    assert T["index"].tolist() == golden8["sorted_original_ids"]
    # root forced to [0,1]^3 (BVHTraverse.jl:856-858)
    assert T["min"][8].tolist() == [0, 0, 0] and T["max"][8].tolist() == [1, 1, 1]


def traversable(T, n_leaves):
    """is_traversable (test/BVHTraverse.jl:87-116): the rope walk from the root visits every leaf once."""
    left, skip = T["left"], T["skip"]
    seen = []
    k = n_leaves + 1
    guard = 0
    while k != 0:
        guard += 1
        assert guard < 10 * len(left) + 10, "rope walk does not terminate"
        if left[k - 1] == 0:
            seen.append(k)
            k = skip[k - 1]
        else:
            k = left[k - 1]
    return seen == list(range(1, n_leaves + 1))


def test_position8_all_28_pairs(oracle, golden8):
    p8, r = golden8["position8"], golden8["neighbor_distance"]
    for nt in (1, 3, 8):
        a, b, d = oracle.build_traverse_bvh(p8, r, 1, nthreads=nt)
        assert len(a) == golden8["expected_pairs"]
        ca = oracle.canonical(a, b, d)
        cb = oracle.canonical(*oracle.brute_force(p8, r, "sqrt"))  # threshold_pairs(unique_pairs(p8), r)
        assert all(np.array_equal(u, v) for u, v in zip(ca, cb))
    a, b, d = oracle.leafbuild_traverse_bvh(p8, r, 2)
    assert len(a) == 28


def test_repeated_threaded_builds_are_traversable(oracle, golden8):
    # the reference's race detector (test/BVHTraverse.jl:189-191): 100 rebuilds must all be traversable
    for _ in range(100):
        T = oracle.tree(golden8["position8"], 10.0, 1, leaf_variant=False, nthreads=8)
        assert traversable(T, 8)
    x = uniform_positions(4000, 3)
    for nt in (1, 8):
        T = oracle.tree(x, 0.05, 4, leaf_variant=True, nthreads=nt)
        assert traversable(T, 1000)


@pytest.mark.parametrize("d", [1, 2, 3, 4, 5])
def test_bvh_equals_alltoall_over_distances(oracle, d):
    # test/BVHTraverse.jl:222-243: pos5000, r = 1e-5 * 10^d
    x = uniform_positions(5000, 5000)
    r = np.float32(0.00001 * 10 ** d)
    A = oracle.canonical(*oracle.leafbuild_traverse_bvh(x, r, 4, nthreads=4))
    B = oracle.canonical(*oracle.brute_force(x, r, "d2"))
    assert all(np.array_equal(u, v) for u, v in zip(A, B))
    Q = oracle.canonical(*oracle.build_traverse_bvh(x, r, 1, nthreads=4))
    assert all(np.array_equal(u, v) for u, v in zip(Q, B))


@pytest.mark.parametrize("apl", [2, 4, 5, 10, 20, 1000])
def test_bvh_equals_alltoall_over_atomsperleaf(oracle, apl):
    # test/BVHTraverse.jl:246-265 (apl = 1 is the atom-query variant, covered above)
    x = uniform_positions(5000, 5001)
    A = oracle.canonical(*oracle.leafbuild_traverse_bvh(x, 0.1, apl, nthreads=3))
    B = oracle.canonical(*oracle.brute_force(x, 0.1, "d2"))
    assert all(np.array_equal(u, v) for u, v in zip(A, B))


@pytest.mark.parametrize("n", [10, 100, 1000, 2000, 3000, 20000])
def test_bvh_equals_alltoall_over_sizes(oracle, n):
    # test/BVHTraverse.jl:268-287, apl = 5, r = 0.1
    x = uniform_positions(n, n)
    A = oracle.canonical(*oracle.leafbuild_traverse_bvh(x, 0.1, 5, nthreads=4))
    B = oracle.canonical(*oracle.brute_force(x, 0.1, "d2"))
    assert all(np.array_equal(u, v) for u, v in zip(A, B))
    # the independent O(N) cell-grid search agrees too (it is what checks the 1M-atom GPU runs)
    dg = oracle.cellgrid_digest(x, 0.1)
    assert dg["count"] == len(B[0]) and dg["xor"] == oracle.digest_pairs(*B)["xor"] and dg["sum"] == oracle.digest_pairs(*B)["sum"]


def test_diary_pair_counts_are_in_range(oracle, golden8):
    # devdiary.md:1321,1302 — unseeded data, so only the magnitude can be checked
    ka = golden8["diary_known_answers"]
    n1 = len(oracle.brute_force(uniform_positions(5000, 1), 0.1)[0])
    n2 = len(oracle.brute_force(uniform_positions(1000, 2), 0.03)[0])
    assert abs(n1 - ka["pos5000_r0.1_pairs"]) < 0.03 * ka["pos5000_r0.1_pairs"]
    assert abs(n2 - ka["pos1000_r0.03_pairs"]) < 25


def test_predicates_differ_only_at_the_cutoff(oracle):
    # BVH predicate d2 < fl(r*r) vs all-to-all predicate sqrt(d2) < r
    x = uniform_positions(5000, 20250313)
    a, b, d = oracle.brute_force(x, 0.1, "d2")
    a2, b2, d2 = oracle.brute_force(x, 0.1, "sqrt")
    s1 = set(zip(a.tolist(), b.tolist()))
    s2 = set(zip(a2.tolist(), b2.tolist()))
    diff = s1 ^ s2
    assert len(diff) <= 3
    dd = dict(zip(zip(a.tolist(), b.tolist()), d.tolist()))
    for p in diff:
        if p in dd:
            assert abs(dd[p] - 0.1) < 2e-8


def test_spec_errors(oracle):
    with pytest.raises(oracle.SpecError, match="evenly divides"):
        oracle.spec(0.1, 10, 3)
    with pytest.raises(oracle.SpecError, match="more than one leaf"):
        oracle.spec(0.1, 10, 10)
    with pytest.raises(oracle.SpecError):  # leafTreeData needs apl >= 2 (BVHTraverse.jl:378)
        oracle.leafbuild_traverse_bvh(uniform_positions(16, 1), 0.1, 1)
    assert oracle.spec(0.1, 5000, 4) == (1250, 1249)


def test_literal_forces_and_verlet(oracle):
    # Forces.jl / Simulator.jl have no reference test: PARITY UNPINNED.  Hand-computed values.
    a = np.array([1, 1, 2], np.int32); b = np.array([2, 3, 3], np.int32)
    d = np.array([0.01, 0.02, 0.005], np.float32)
    f = oracle.force_lennardjones(3, a, b, d)
    def lj(dd):
        dd = float(np.float32(dd))
        pre = float(np.float32(np.float32(24) * np.float32(-1e10)) / np.float32(dd))
        return pre * ((2 * 0.0001 / dd) ** 12.0 - (0.0001 / dd) ** 6.0)
    e1 = np.float32(np.float32(0) + lj(0.01)); e1 = np.float32(float(e1) + lj(0.02))
    assert np.allclose(f[0], e1, rtol=1e-6) and f[0, 0] == f[0, 1] == f[0, 2]
    assert np.allclose(f[1], np.float32(lj(0.005)), rtol=1e-6) and np.all(f[2] == 0)
    q = np.array([1.0, -2.0, 0.5], np.float32)
    fc = oracle.force_coulomb(3, a, b, d, q)
    # sequential semantics: f1 += q1q2/d^2; f2 -= f1; f1 += q1q3/d^2; f3 -= f1; f2 += q2q3/d^2; f3 -= f2
    f1 = np.float32(-2.0) / np.float32(np.float32(0.01) * np.float32(0.01)); f2 = -f1
    f1 = f1 + np.float32(0.5) / np.float32(np.float32(0.02) * np.float32(0.02)); f3 = -f1
    f2 = f2 + np.float32(-1.0) / np.float32(np.float32(0.005) * np.float32(0.005)); f3 = f3 - f2
    assert np.allclose(fc[:, 0], [f1, f2, f3], rtol=1e-6)
    s = oracle.sum_forces(f, fc)
    assert np.array_equal(s, f + fc)
    # Verlet body: x += v dt + a dt^2/2 ; v += (a + a')dt/2, then reflect
    pos = np.array([[0.5, 0.5, 0.99]], np.float32); vel = np.array([[0.1, -0.2, 0.3]], np.float32)
    F = np.array([[1.0, 0.0, 0.0]], np.float32); Fn = np.array([[3.0, 0.0, 0.0]], np.float32)
    m = np.array([2.0], np.float32)
    p2, v2 = oracle.verlet(pos, vel, F, Fn, m, 0.1)
    assert np.allclose(p2, [[0.5 + 0.01 + 0.5 * 0.01 / 2, 0.48, 1.02]], rtol=1e-6)
    assert np.allclose(v2, [[0.1 + (0.5 + 1.5) * 0.1 / 2, -0.2, 0.3]], rtol=1e-6)
    p3, v3 = oracle.boundary_reflect(p2, v2, (0, 0, 0), (1, 1, 1))
    assert p3[0, 2] == 1.0 and v3[0, 2] == -v2[0, 2] and p3[0, 0] == p2[0, 0]


def test_rescale_velocity_restatement_by_hand(oracle):
    # rescale_velocity! (Simulator.jl:119-144) on two atoms, worked by hand in Float32:
    #   v1 = (3,4,0) -> |v| = 5, m = 2 ; v2 = (0,0,2) -> |v| = 2, m = 1 ; objects = 2
    #   Ti = (2/(3*2*1)) * (5*2/2 + 2*1/2) = (1/3) * 6 = 2 ;  beta = sqrt(1 + gamma*(Tf/Ti - 1))
    v = np.array([[3, 4, 0], [0, 0, 2]], np.float32)
    m = np.array([2, 1], np.float32)
    out, ti, beta = oracle.rescale_velocity(v, 8.0, 1.0, m, 2)
    assert abs(ti - 2.0) < 1e-6 and abs(beta - 2.0) < 1e-6          # sqrt(8/2)
    assert np.allclose(out, 2.0 * v, rtol=1e-6)
    out, ti, beta = oracle.rescale_velocity(v, 8.0, 0.0, m, 2)      # gamma = 0: no rescaling (the docstring, :113-116)
    assert beta == 1.0 and np.array_equal(out, v)
    out, ti, beta = oracle.rescale_velocity(v, 8.0, 0.5, m, 2)
    assert abs(beta - np.sqrt(1 + 0.5 * 3.0)) < 1e-6


def test_philox4x32_10_known_answers(oracle):
    # the three known-answer vectors the Random123 library publishes for philox4x32-10 (kat_vectors): zeros, all ones,
    # and the digits of pi; they pin the counter-based generator behind nb200_collect_objects / oracle.collect_objects
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
        ((0xFFFFFFFF,) * 4, (0xFFFFFFFF, 0xFFFFFFFF), (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
        ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
         (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
    ]
    for ctr, key, want in kat:
        got = oracle.philox4x32_10(np.array([ctr], np.uint32), key)[0]
        assert tuple(int(x) for x in got) == want


def test_collect_objects_restatement_properties(oracle):
    # collect_objects (MDInput.jl:305-369): ranges of the draws, the velocity rule worked by hand, and the re-draw loop
    # of generate_pruned_positions! (:260-283) leaving no two atoms closer than minimumdistance
    n, lo, hi = 3000, (-1.0, 0.0, 2.0), (3.0, 1.0, 2.5)
    r = oracle.collect_objects(n, 42, lo, hi, 1.0, 2.0, -0.5, 0.5, 0.72, True, 0.05)
    p = r["position"]
    assert p.dtype == np.float32 and (p >= np.float32(lo)).all() and (p <= np.float32(hi)).all()
    assert r["rounds"] >= 1 and r["redrawn"] >= r["rounds"]
    assert len(oracle.brute_force(p, 0.05, "d2")[0]) == 0
    assert r["mass"].min() >= 1.0 and r["mass"].max() <= 2.0 and abs(r["mass"].mean() - 1.5) < 0.05
    assert r["charge"].min() >= -0.5 and r["charge"].max() <= 0.5 and abs(r["charge"].mean()) < 0.05
    for d in range(3):   # veldist sums to one per axis: sum_i v_i m_i = 3 n T
        assert np.isclose((r["velocity"][:, d].astype(np.float64) * r["mass"]).sum(), 3 * n * 0.72, rtol=1e-4)
    # the same seed gives the same system; atoms that were never re-drawn keep their first position
    again = oracle.collect_objects(n, 42, lo, hi, 1.0, 2.0, -0.5, 0.5, 0.72, True, 0.05)
    assert np.array_equal(again["position"], p)
    first = oracle.collect_objects(n, 42, lo, hi, 1.0, 2.0, -0.5, 0.5, 0.72, True, 0.0)
    assert first["rounds"] == 0 and (first["position"] == p).all(axis=1).sum() >= n - r["redrawn"]
    # randomvelocity = false (:329-334): v = T / n * 3 * n / mass on every axis, Float32 until the Float64 division
    fixed = oracle.collect_objects(8, 1, (0, 0, 0), (1, 1, 1), 2.0, 2.0, 0.0, 0.0, 0.5, False, 0.0)
    f = np.float32
    want = f(np.float64((f(0.5) / f(8)) * f(3) * f(8)) / 2.0)
    assert np.all(fixed["velocity"] == want) and np.all(fixed["mass"] == 2.0) and np.all(fixed["charge"] == 0.0)
    with pytest.raises(RuntimeError, match="Objects could not be placed"):
        oracle.collect_objects(500, 3, (0, 0, 0), (1, 1, 1), 1.0, 2.0, -1.0, 1.0, 1.0, True, 0.5, max_rounds=3)


def test_oracle_reproduces_the_committed_pair_digests(oracle, golden_digests):
    # tests/golden/pair_digests.json pins the exact pair sets the GPU tests compare against; every search the oracle
    # has (all-pairs loop, the restated reference BVH, the independent cell grid) must reproduce them
    for name, g in golden_digests.items():
        x = uniform_positions(g["n"], g["seed"])
        r = np.float32(g["cutoff"])
        want = (g["count"], g["xor"], g["sum"])
        cg = oracle.cellgrid_digest(x, r)
        assert (cg["count"], cg["xor"], cg["sum"]) == want, name
        if g["n"] <= 20_000:
            for pairs in (oracle.brute_force(x, r, "d2"), oracle.leafbuild_traverse_bvh(x, r, 4, nthreads=4)):
                d = oracle.digest_pairs(*pairs)
                assert (d["count"], d["xor"], d["sum"]) == want, name
