"""GPU parity tests of the device-side system setup (nb200_collect_objects; MDInput.jl:175-190, 228-283, 305-369).

The draws are a pure function of (seed, atom, round), restated in oracle/nd_oracle.py, so the generated system is
compared bit for bit; at sizes beyond the oracle's O(N^2) pair loop the result is checked through its properties."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


CASES = [
    # n, seed, box_min, box_max, mass range, charge range, temperature, randomvelocity, minimumdistance
    (5000, 12345, (0, 0, 0), (1, 1, 1), (1.0, 2.0), (-1.0, 1.0), 0.72, True, 0.02),
    (5000, 2**40 + 17, (-1, 0, 2), (3, 1, 2.5), (0.5, 0.5), (0.0, 1e-3), 300.0, False, 0.03),
    (4097, 7, (0, 0, 0), (1, 1, 1), (1.0, 4.0), (-2.0, 2.0), 1.0, True, 0.0),
    (2, 99, (0, 0, 0), (1, 1, 1), (1.0, 2.0), (-1.0, 1.0), 1.0, True, 0.5),
    (33, 5, (0, 0, 0), (1, 1, 1), (1.0, 2.0), (-1.0, 1.0), 1.0, True, 0.3),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"n{c[0]}-r{c[8]}")
def test_collect_objects_matches_the_oracle_bit_for_bit(big_handle, oracle, case):
    n, seed, lo, hi, (m0, m1), (q0, q1), temp, rnd, mind = case
    h = big_handle
    h.set_box(lo, hi)
    h.set_forcefield(eps=0.0, sigma=1.0, kcoul=0.0, cutoff=0.02, shift=True)
    mass, charge, rounds, redrawn = h.collect_objects(n, seed, m0, m1, q0, q1, temp, rnd, mind)
    ref = oracle.collect_objects(n, seed, lo, hi, m0, m1, q0, q1, temp, rnd, mind)
    assert (rounds, redrawn) == (ref["rounds"], ref["redrawn"])
    assert np.array_equal(bits(mass), bits(ref["mass"]))
    assert np.array_equal(bits(charge), bits(ref["charge"]))
    assert np.array_equal(bits(h.get_positions()), bits(ref["position"]))
    assert np.array_equal(bits(h.get_velocities()), bits(ref["velocity"]))
    if mind > 0 and n >= 1000:
        assert rounds >= 1, "the case is meant to exercise the re-draw loop"
    # the system is resident exactly as after set_system: the step loop runs on it
    h.step(2, 1e-4)
    assert np.isfinite(h.get_positions()).all()
    h.set_box((0, 0, 0), (1, 1, 1))


def test_collect_objects_same_seed_same_system_other_seed_other_system(big_handle):
    h = big_handle
    h.set_box((0, 0, 0), (1, 1, 1))
    h.set_forcefield(eps=0.0, sigma=1.0, kcoul=0.0, cutoff=0.01, shift=True)
    args = (1.0, 2.0, -1.0, 1.0, 0.5, True, 0.01)
    h.collect_objects(20000, 1, *args)
    a = h.get_positions()
    h.collect_objects(20000, 1, *args)
    b = h.get_positions()
    h.collect_objects(20000, 2, *args)
    c = h.get_positions()
    assert np.array_equal(bits(a), bits(b))
    assert not np.array_equal(bits(a), bits(c))


def test_collect_objects_200k_properties(big_handle):
    """Beyond the O(N^2) oracle: every atom inside the box, no two atoms closer than minimumdistance (checked with an
    independent k-d tree in Float64 on the Float32 coordinates), draws spread like Uniform(a, b)."""
    from scipy.spatial import cKDTree
    n, mind = 200_000, 0.012          # ~0.7 too-close partners per atom at the first draw: many re-draw rounds
    h = big_handle
    h.set_box((0, 0, 0), (1, 1, 1))
    h.set_forcefield(eps=0.0, sigma=1.0, kcoul=0.0, cutoff=mind, shift=True)
    mass, charge, rounds, redrawn = h.collect_objects(n, 20250313, 1.0, 3.0, -1.0, 1.0, 0.72, True, mind)
    pos = h.get_positions()
    vel = h.get_velocities()
    assert rounds >= 3 and redrawn > n // 10
    assert (pos >= 0).all() and (pos <= 1).all()
    x = pos.astype(np.float64)
    close = cKDTree(x).query_pairs(mind * (1 - 1e-6), output_type="ndarray")
    d2 = ((x[close[:, 0]] - x[close[:, 1]]) ** 2).sum(1) if len(close) else np.zeros(0)
    assert not (d2 < np.float64(np.float32(mind) * np.float32(mind)) * (1 - 1e-6)).any()
    assert h.pair_count() == 0      # the list left by collect_objects' forces is built at cutoff = minimumdistance
    for v, a, b in ((mass, 1.0, 3.0), (charge, -1.0, 1.0)):
        assert v.min() >= a and v.max() <= b
        assert abs(v.mean() - (a + b) / 2) < 6 * (b - a) / np.sqrt(12 * n)
    # velocity rule (MDInput.jl:319-336): sum over atoms of v*m per axis = 3 n T (veldist sums to one)
    for d in range(3):
        assert np.isclose((vel[:, d].astype(np.float64) * mass).sum(), 3 * n * 0.72, rtol=1e-4)


def test_collect_objects_reports_an_impossible_packing(pkg, big_handle):
    h = big_handle
    h.set_box((0, 0, 0), (1, 1, 1))
    h.set_forcefield(eps=0.0, sigma=1.0, kcoul=0.0, cutoff=0.02, shift=True)
    with pytest.raises(pkg.NB200Error) as e:
        h.collect_objects(4000, 3, 1.0, 2.0, -1.0, 1.0, 1.0, True, 0.2, max_rounds=4)
    assert e.value.code == pkg._lib.NB200_ERR_STATE
    assert "Objects could not be placed" in str(e.value)
    with pytest.raises(pkg.NB200Error):
        h.collect_objects(100, 3, 2.0, 1.0, -1.0, 1.0, 1.0, True, 0.0)      # Uniform(2, 1)
    # the handle stays usable
    _, _, rounds, _ = h.collect_objects(4000, 3, 1.0, 2.0, -1.0, 1.0, 1.0, True, 0.01)
    assert rounds >= 0 and len(h.get_positions()) == 4000


def test_collect_objects_through_the_reference_interface(pkg):
    """collect_objects(Collector) -> simulate_bvh!(sys, spec, bvhspec, clct) as a user of the reference writes it
    (docs/src/tutorial.md:13-30), with the setup on the device."""
    clct = pkg.GenericRandomCollector(objectnumber=8192, minDim=(0.0, 0.0, 0.0), maxDim=(1.0, 1.0, 1.0), temperature=0.01,
                                      randomvelocity=True, minmass=1.0, maxmass=2.0, minimumdistance=0.02, mincharge=-1.0,
                                      maxcharge=1.0, seed=11)
    sys = pkg.collect_objects(clct, backend=pkg.B200Backend())
    assert sys.position.shape == (8192, 3) and sys.velocity.shape == (8192, 3)
    spec = pkg.SimSpec(duration=5, stepwidth=1e-5)
    bvhspec = pkg.SpheresBVHSpecs(neighbor_distance=0.03, atom_count=8192, floattype=np.float32, atomsperleaf=4)
    first = sys.position.copy()
    poslog = pkg.simulate_bvh_(sys, spec, bvhspec, clct)
    assert len(poslog) == 6 and np.array_equal(poslog[0], first)
    assert not np.array_equal(poslog[-1], first)
