"""GPU tests of the force kernels and the device step loop against the oracle."""
import numpy as np
import pytest

from conftest import uniform_positions

pytestmark = pytest.mark.gpu

FORCE_RTOL = 1e-5  # north star: forces and energies within 1e-5 relative (Float32)


# Net forces that are not the result of a cancellation (|F| > 5 % of sum_j |F_ij|) must hold the relative bound on |F|
# itself, atom by atom; the factor 4 is the rsqrt.approx pair evaluation (2 ulp in 1/r, squared and cubed: pair_force.cuh).
STRONG_FORCE_RTOL = 4e-5


def strong_force_rel_error(f, f64, scale):
    mag = np.linalg.norm(f64, axis=1)
    strong = mag > 0.05 * scale
    assert strong.sum() > len(f) // 20, "the test system has too few atoms with a strong net force"
    worst = float((np.linalg.norm(f - f64, axis=1)[strong] / mag[strong]).max())
    print(f"max relative net-force error over {int(strong.sum())} strong atoms: {worst:.3e}")
    return worst


def lattice(m, jitter, seed, margin=0.1):
    rng = np.random.default_rng(seed)
    g = (np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij"), -1).reshape(-1, 3) + 0.5) / m
    a = (1 - 2 * margin) / m
    x = margin + (1 - 2 * margin) * g + jitter * a * (rng.random(g.shape) - 0.5)
    return x.astype(np.float32), a


@pytest.mark.parametrize("list_mode", [1, 0])  # NB200_LIST_HALF (reaction scattered), NB200_LIST_DIRECTED (owner computes)
@pytest.mark.parametrize("with_charge", [False, True])
def test_physical_forces_match_fp64_oracle(pkg, oracle, with_charge, list_mode):
    x, a = lattice(20, 0.3, 1)
    n = len(x)
    sigma = a / 1.1
    rc = 2.5 * sigma
    q = ((np.random.default_rng(2).random(n) - 0.5) * 0.2).astype(np.float32) if with_charge else None
    kc = 0.05 * sigma if with_charge else 0.0
    h = pkg.Handle(n)
    h.set_list_mode(list_mode)
    h.set_forcefield(eps=1.0, sigma=sigma, kcoul=kc, cutoff=rc, shift=True)
    h.set_system(x, None, None, q)
    pa, pb, pd = h.get_pairs()
    ref = oracle.brute_force(x, rc, "d2")
    assert len(pa) == len(ref[0])
    f64, pe64, scale = oracle.forces_physical_f64(x, q, pa, pb, 1.0, sigma, kc, rc, True)
    f = h.get_forces()
    # tolerance is relative to sum_j |F_ij| — the conditioning scale of a Float32 sum (DESIGN.md "Forces")
    err = np.abs(f - f64).max(axis=1) / scale
    assert err.max() < FORCE_RTOL, err.max()
    # and relative to the net force itself for the bulk of atoms
    big = np.linalg.norm(f64, axis=1) > 1e-3 * scale
    rel = np.linalg.norm(f - f64, axis=1)[big] / np.linalg.norm(f64, axis=1)[big]
    assert np.median(rel) < FORCE_RTOL
    # atoms whose net force is not a cancellation (|F| > 5 % of sum_j |F_ij|): the MAXIMUM relative error of the net force
    assert strong_force_rel_error(f, f64, scale) < STRONG_FORCE_RTOL
    ke, pe = h.get_energies()
    assert ke == 0.0
    assert abs(pe - pe64.sum()) <= FORCE_RTOL * np.abs(pe64).sum()
    # Newton's third law: the net force on the whole system must be ~0 in either list form
    assert np.abs(f.sum(0)).max() < 1e-4 * np.abs(f).sum(0).max()
    h.close()


@pytest.mark.parametrize("list_mode", [1, 0])
@pytest.mark.parametrize("with_charge", [False, True])
def test_fused_traversal_forces_match_the_tile_kernel_and_the_oracle(pkg, oracle, with_charge, list_mode):
    # The step loop evaluates the pair forces inside the traversal (nb200_set_fused_force, default on); the separate
    # kernel reads the same tiles back from the list.  After the same steps both must agree with the fp64 oracle on
    # the exact pair set of the CURRENT positions, and the list the fused traversal wrote must be that pair set.
    x, a = lattice(24, 0.1, 11)
    n = len(x)
    sigma = a / 1.1
    rc = 2.5 * sigma
    rng = np.random.default_rng(12)
    q = ((rng.random(n) - 0.5) * 0.2).astype(np.float32) if with_charge else None
    kc = 0.05 * sigma if with_charge else 0.0
    v = (rng.standard_normal((n, 3)) * 0.8 * sigma).astype(np.float32)  # T* ~ 0.64 in reduced units
    mass = np.full(n, 1.0 / sigma ** 2, np.float32)                      # m* = 1 with lengths in box units
    out = {}
    for fused in (True, False):
        h = pkg.Handle(n)
        h.set_list_mode(list_mode)
        h.set_fused_force(fused)
        h.set_forcefield(eps=1.0, sigma=sigma, kcoul=kc, cutoff=rc, shift=True)
        h.set_system(x, v, mass, q)
        h.step(3, 0.002)
        xs = h.get_positions()
        f = h.get_forces()
        pa, pb, pd = h.get_pairs()
        ref = oracle.canonical(*oracle.brute_force(xs, rc, "d2"))
        got = oracle.canonical(pa, pb, pd)
        assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])
        assert np.array_equal(got[2].view(np.uint32), ref[2].view(np.uint32))
        f64, pe64, scale = oracle.forces_physical_f64(xs, q, pa, pb, 1.0, sigma, kc, rc, True)
        err = np.abs(f - f64).max(axis=1) / scale
        assert err.max() < FORCE_RTOL, (fused, err.max())
        assert strong_force_rel_error(f, f64, scale) < STRONG_FORCE_RTOL, fused
        ke, pe = h.get_energies()  # re-runs the tile kernel with energies on the list the step wrote
        assert abs(pe - pe64.sum()) <= FORCE_RTOL * np.abs(pe64).sum()
        f2 = h.get_forces()
        assert (np.abs(f2 - f64).max(axis=1) / scale).max() < FORCE_RTOL
        out[fused] = (xs, f)
        h.close()
    # same trajectory either way (Float32 sums in a different order: agreement, not identity)
    assert np.abs(out[True][0] - out[False][0]).max() < 1e-6
    assert np.abs(out[True][1] - out[False][1]).max() <= 2e-5 * np.abs(out[False][1]).max()


def test_literal_reference_forces(pkg, oracle):
    # Forces.jl:6-66 behind the reference's own signatures, on a list from the search
    x = uniform_positions(3000, 12)
    spec = pkg.SpheresBVHSpecs(neighbor_distance=0.05, atom_count=3000, floattype=np.float32, atomsperleaf=4)
    pl = pkg.leafbuild_traverse_bvh(x, spec)
    n = len(x)
    f = np.zeros((n, 3), np.float32)
    pkg.force_lennardjones_(f, pl, x)
    ref = oracle.force_lennardjones(n, pl.a, pl.b, pl.d)
    denom = np.maximum(np.abs(ref), 1e-30)
    assert (np.abs(f - ref) / denom)[ref != 0].max() < 1e-5
    assert np.array_equal(f == 0, ref == 0)
    # empty list: force zeroed, early return (Forces.jl:34-36)
    f[:] = 5
    pkg.force_lennardjones_(f, (np.empty(0, np.int32), np.empty(0, np.int32), np.empty(0, np.float32)), x)
    assert np.all(f == 0)
    # literal Coulomb is order dependent: same list order -> bit-identical result
    q = ((np.random.default_rng(3).random(n) - 0.5)).astype(np.float32)
    sub = slice(0, 20000)
    fc = np.zeros((n, 3), np.float32)
    pkg.force_coulomb_(fc, (pl.a[sub], pl.b[sub], pl.d[sub]), q)
    rc = oracle.force_coulomb(n, pl.a[sub], pl.b[sub], pl.d[sub], q)
    assert np.array_equal(fc.view(np.uint32), rc.view(np.uint32))
    s = np.zeros_like(f)
    pkg.sum_forces_(s, ref, rc)
    assert np.array_equal(s, oracle.sum_forces(ref, rc))


def test_literal_verlet_update_is_bit_exact(big_handle, oracle):
    rng = np.random.default_rng(8)
    n = 10_000
    pos = rng.random((n, 3)).astype(np.float32)
    vel = (rng.standard_normal((n, 3)) * 0.3).astype(np.float32)
    f = rng.standard_normal((n, 3)).astype(np.float32)
    fn = rng.standard_normal((n, 3)).astype(np.float32)
    m = rng.uniform(1, 5, n).astype(np.float32)
    for dt in (1.0, 0.37, 0.005):
        p_ref, v_ref = oracle.verlet(pos, vel, f, fn, m, dt)
        p_ref, v_ref = oracle.boundary_reflect(p_ref, v_ref, (0, 0, 0), (1, 1, 1))
        p, v = big_handle.verlet_update(pos, vel, f, fn, m, dt, (0, 0, 0), (1, 1, 1))
        assert np.array_equal(p.view(np.uint32), p_ref.view(np.uint32))
        assert np.array_equal(v.view(np.uint32), v_ref.view(np.uint32))


def test_force_free_simulate_bvh_is_bit_exact(pkg, oracle):
    # simulate_bvh! (Simulator.jl:327-379) never computes forces: with F == 0 the device loop must give
    # the reference trajectory bit for bit (Verlet body + boundary_reflect!), rebuilding the list each step.
    n = 1024
    c = pkg.GenericRandomCollector(objectnumber=n, minDim=(0.0, 0.0, 0.0), maxDim=(1.0, 1.0, 1.0), temperature=0.01,
                                   randomvelocity=False, minmass=1.0, maxmass=5.0, minimumdistance=0.0001,
                                   mincharge=-1e-9, maxcharge=1e-9, seed=4)
    sysm = pkg.collect_objects(c)
    sysm.velocity[:] = (np.random.default_rng(5).standard_normal((n, 3)) * 0.05).astype(np.float32)
    pos0, vel0 = sysm.position.copy(), sysm.velocity.copy()
    bvhspec = pkg.SpheresBVHSpecs(neighbor_distance=0.03, atom_count=n, floattype=np.float32, atomsperleaf=4)
    simspec = pkg.SimSpec(duration=40, stepwidth=1, currentstep=1, logLength=10, vDamp=1, threshold=0.03)
    poslog = pkg.simulate_bvh_(sysm, simspec, bvhspec, c)
    assert len(poslog) == 41
    p, v = pos0.copy(), vel0.copy()
    zero = np.zeros_like(p)
    for step in range(40):
        p, v = oracle.verlet(p, v, zero, zero, sysm.mass, 1.0)
        p, v = oracle.boundary_reflect(p, v, c.minDim, c.maxDim)
        assert np.array_equal(poslog[step + 1].view(np.uint32), p.view(np.uint32)), step
    assert np.array_equal(sysm.velocity.view(np.uint32), v.view(np.uint32))
    # and the list of the last step is the exact pair set of the final positions
    h = pkg.get_handle(n)
    got = h.get_pairs()
    ref = oracle.brute_force(p, 0.03, "d2")
    assert len(got[0]) == len(ref[0])


def test_rescale_velocity_matches_the_reference_formula(pkg, oracle):
    # rescale_velocity! (Simulator.jl:119-144): Ti from the SPEED (as written), beta = (1 + gamma (Tf/Ti - 1))^0.5.
    # PARITY UNPINNED in the reference (no test there); checked against the oracle's Float32 restatement.
    n = 5000
    rng = np.random.default_rng(8)
    v = (rng.standard_normal((n, 3)) * 0.2).astype(np.float32)
    mass = rng.uniform(1.0, 5.0, n).astype(np.float32)
    for tf, gamma in ((0.05, 1.0), (0.3, 0.25)):
        want, ti, beta = oracle.rescale_velocity(v, tf, gamma, mass, n)
        got = v.copy()
        pkg.rescale_velocity_(got, tf, gamma, mass, n)
        assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max(), (tf, gamma)
    # physical variant on a resident system: gamma = 1 lands exactly on the target kinetic temperature
    h = pkg.Handle(n)
    h.set_forcefield(0.0, 1.0, 0.0, 1e-6, True)
    h.set_system(rng.random((n, 3)).astype(np.float32), v, mass, None)
    h.rescale_velocity(0.7, 1.0, physical=True)
    vs = h.get_velocities().astype(np.float64)
    t_kin = (mass[:, None] * vs * vs).sum() / (3 * n)
    assert abs(t_kin - 0.7) < 1e-5 * 0.7
    h.close()


def test_simulate_call_logs_frames_and_rescales(pkg):
    # nb200_simulate = the simulate!/simulate_bvh! loop in one call: frames every log_every steps (copied out on a
    # second stream), rescale_velocity! every rescale_every steps (Simulator.jl:241-245).  Must equal the manual loop.
    x, a = lattice(12, 0.3, 31)
    n = len(x)
    sigma = a / 1.1
    rng = np.random.default_rng(32)
    v = (rng.standard_normal((n, 3)) * 0.4 * sigma).astype(np.float32)
    mass = np.full(n, 1.0 / sigma ** 2, np.float32)
    q = (0.1 * (rng.random(n) - 0.5)).astype(np.float32)
    dt, nsteps, log_every, rescale_every, tf, gamma = 0.003, 23, 3, 5, 0.02, 0.5

    def fresh():
        h = pkg.Handle(n)
        h.set_forcefield(1.0, sigma, 0.3 * sigma, 2.5 * sigma, True)
        h.set_system(x, v, mass, q)
        return h
    h1 = fresh()
    frames = h1.simulate(nsteps, dt, log_every=log_every, rescale_every=rescale_every, target_temperature=tf, gamma=gamma)
    assert frames.shape == (nsteps // log_every, n, 3)
    h2 = fresh()
    k = 0
    for s in range(1, nsteps + 1):
        h2.step(1, dt)
        if s % rescale_every == 0:
            h2.rescale_velocity(tf, gamma)
        if s % log_every == 0:
            assert np.abs(frames[k] - h2.get_positions()).max() < 2e-6, (s, k)
            k += 1
    assert k == len(frames)
    assert np.abs(h1.get_velocities() - h2.get_velocities()).max() < 1e-4 * np.abs(h2.get_velocities()).max()
    assert len(h1.simulate(4, dt, log_every=0)) == 0  # no log requested
    h1.close(); h2.close()


def test_md_steps_follow_the_fp64_integrator(pkg, oracle):
    x, a = lattice(12, 0.2, 3)
    n = len(x)
    sigma = a / 1.12
    rc = 2.5 * sigma
    rng = np.random.default_rng(4)
    v = (rng.standard_normal((n, 3)) * 0.5 * sigma).astype(np.float32)
    # lengths are in box units and time in LJ tau, so the LJ mass m* enters as m*/sigma^2
    m = (rng.uniform(1, 2, n) / sigma ** 2).astype(np.float32)
    q = ((rng.random(n) - 0.5) * 0.2).astype(np.float32)
    kc = 0.05 * sigma
    dt = 0.005
    h = pkg.Handle(n)
    h.set_forcefield(1.0, sigma, kc, rc, True)
    h.set_system(x, v, m, q)
    ke0, pe0 = h.get_energies()
    h.step(20, dt)
    p_gpu, v_gpu = h.get_positions(), h.get_velocities()
    ke1, pe1 = h.get_energies()
    p64, v64, f64, en = oracle.md_steps_f64(x, v, m, q, 20, dt, rc, 1.0, sigma, kc, True, (0, 0, 0), (1, 1, 1))
    assert np.abs(p_gpu - p64).max() < 2e-5 * sigma
    assert np.abs(v_gpu - v64).max() < 1e-4 * np.abs(v64).max()
    assert abs(ke1 - en["ke"]) < 1e-4 * abs(en["ke"]) and abs(pe1 - en["pe"]) < 1e-4 * abs(en["pe"])
    assert abs((ke1 + pe1) - (ke0 + pe0)) < 1e-3 * abs(ke0 + pe0)  # hot start: KE doubles in 20 steps
    h.close()


def test_nve_energy_drift_1000_steps(pkg):
    # BASELINE config 2 shape at reduced size for test time: LJ fluid, rho* = 0.8442, T* = 0.72, rc = 2.5 sigma,
    # dt = 0.005 tau, 1000 steps, neighbour rebuild every step.  Bound on |dE/E0| stated here: 2e-3.
    m = 24
    n = m ** 3
    rho = 0.8442
    L = (n / rho) ** (1 / 3)              # box edge in sigma
    sigma = 0.8 / L                        # lattice occupies [0.1, 0.9]^3 of the unit box
    x, _ = lattice(m, 0.02, 11)
    rng = np.random.default_rng(1)
    v = rng.standard_normal((n, 3)) * np.sqrt(0.72)
    v -= v.mean(0)
    v = (v * sigma).astype(np.float32)    # velocities in box units per tau
    mass = np.full(n, 1.0 / sigma ** 2, np.float32)  # m* = 1 with lengths in box units, time in tau
    h = pkg.Handle(n)
    h.set_forcefield(1.0, sigma, 0.0, 2.5 * sigma, True)
    h.set_system(x, v, mass, None)
    dt = 0.005
    ke0, pe0 = h.get_energies()
    e0 = ke0 + pe0
    worst = 0.0
    for _ in range(10):
        h.step(100, dt)
        ke, pe = h.get_energies()
        e = ke + pe
        worst = max(worst, abs((e - e0) / e0))
    assert np.isfinite(e) and worst < 2e-3, worst
    assert h.get_stats()["steps_done"] == 1000
    h.close()


def test_step_host_roundtrip(pkg):
    x, a = lattice(10, 0.1, 6)
    n = len(x)
    sigma = a / 1.12
    v = (np.random.default_rng(2).standard_normal((n, 3)) * 0.3 * sigma).astype(np.float32)
    mass = np.full(n, 1.0 / sigma ** 2, np.float32)
    h1 = pkg.Handle(n)
    h1.set_forcefield(1.0, sigma, 0.0, 2.5 * sigma, True)
    h1.set_system(x, v, mass)
    h1.step(5, 0.002)
    p_ref, v_ref = h1.get_positions(), h1.get_velocities()
    h2 = pkg.Handle(n)
    h2.set_forcefield(1.0, sigma, 0.0, 2.5 * sigma, True)
    h2.set_system(x, v, mass)
    xb, vb = x.copy(), v.copy()
    for _ in range(5):  # one host round trip per step, like the reference's poslog push
        h2.step_host(xb, vb, 1, 0.002)
    assert np.abs(xb - p_ref).max() < 1e-6 and np.abs(vb - v_ref).max() < 1e-4 * np.abs(v_ref).max()
    h1.close(); h2.close()


@pytest.mark.parametrize("every", [2, 5, 16])
def test_resort_interval_keeps_the_pair_set_exact(pkg, oracle, every):
    # nb200_set_resort_interval: between full re-sorts the atoms keep their order and only the leaf boxes are refreshed
    # (the reference's TreeData! update path); the list must still be the exact pair set of the current positions and
    # the trajectory must match the sort-every-step loop.
    x, a = lattice(14, 0.4, 21)
    n = len(x)
    sigma = a / 1.1
    rng = np.random.default_rng(22)
    v = (rng.standard_normal((n, 3)) * 1.5 * sigma).astype(np.float32)  # hot: atoms move ~1 % of a leaf per step
    mass = np.full(n, 1.0 / sigma ** 2, np.float32)
    ref = pkg.Handle(n)
    ref.set_forcefield(1.0, sigma, 0.0, 2.5 * sigma, True)
    ref.set_system(x, v, mass)
    h = pkg.Handle(n)
    h.set_forcefield(1.0, sigma, 0.0, 2.5 * sigma, True)
    h.set_resort_interval(every)
    h.set_system(x, v, mass)
    for nsteps in (1, 3, 7, 9):
        h.step(nsteps, 0.004)
        ref.step(nsteps, 0.004)
        p = h.get_positions()
        got = h.get_pairs()
        want = oracle.brute_force(p, np.float32(2.5 * sigma), "d2")
        cg, cw = oracle.canonical(*got), oracle.canonical(*want)
        assert len(cg[0]) == len(cw[0]) and np.array_equal(cg[0], cw[0]) and np.array_equal(cg[1], cw[1])
        assert np.array_equal(cg[2].view(np.uint32), cw[2].view(np.uint32))
        assert np.abs(p - ref.get_positions()).max() < 2e-6
    h.close(); ref.close()


def test_leapfrog_host_async_matches_step_loop(pkg):
    # nb200_leapfrog_host_async iterated (one search per call, host round trip every step) must reproduce the
    # device-resident loop; three interleaved replicas exercise the overlap of independent handles.
    x, a = lattice(10, 0.1, 8)
    n = len(x)
    sigma = a / 1.12
    rng = np.random.default_rng(4)
    mass = np.full(n, 1.0 / sigma ** 2, np.float32)
    q = (0.1 * (rng.random(n) - 0.5)).astype(np.float32)
    dt, nsteps, R = 0.002, 6, 3
    vs = [(rng.standard_normal((n, 3)) * 0.3 * sigma).astype(np.float32) for _ in range(R)]
    refs = []
    for r in range(R):
        h = pkg.Handle(n)
        h.set_forcefield(1.0, sigma, float(sigma) * 0.5, 2.5 * sigma, True)
        h.set_system(x, vs[r], mass, q)
        h.step(nsteps, dt)
        refs.append((h.get_positions(), h.get_velocities()))
        h.close()
    hs, xb, vb = [], [], []
    for r in range(R):
        h = pkg.Handle(n)
        h.set_forcefield(1.0, sigma, float(sigma) * 0.5, 2.5 * sigma, True)
        h.set_system(x, vs[r], mass, q)
        hs.append(h); xb.append(x.copy()); vb.append(vs[r].copy())
    for s in range(nsteps):
        for r in range(R):
            hs[r].sync()
            hs[r].leapfrog_host_async(xb[r].ctypes.data, vb[r].ctypes.data, 3, n, dt, s > 0)
    for r in range(R):
        hs[r].sync()
        assert np.abs(xb[r] - refs[r][0]).max() < 1e-6, r
        v_sync = hs[r].get_velocities()  # closes the pending half kick with F(x(t)) recomputed on demand
        assert np.abs(v_sync - refs[r][1]).max() < 1e-4 * np.abs(refs[r][1]).max(), r
        assert hs[r].get_stats()["steps_done"] == nsteps
        hs[r].close()


def test_c3_config_1m_list_stays_exact_through_the_step_loop(pkg, oracle):
    # BASELINE config 3 at full size (1M LJ + Coulomb atoms, rc = 2.5 sigma): after MD steps — full re-sort every step,
    # then with the re-sort interval — the list on the device must be the exact pair set of the positions on the device.
    # The O(N^2) oracle cannot run at this size; the independent O(N) cell-grid search gives order-independent digests.
    import sys
    from conftest import ROOT
    sys.path.insert(0, ROOT)
    from bench import make_workload
    w = make_workload("c3")
    n = w["n"]
    h = pkg.Handle(n)
    h.set_forcefield(w["eps"], w["sigma"], w["kcoul"], w["cutoff"], True)
    h.set_system(w["pos"], w["vel"], w["mass"], w["charge"])
    ke0, pe0 = h.get_energies()
    for every, nsteps in ((1, 6), (8, 11)):
        h.set_resort_interval(every)
        h.step(nsteps, w["dt"])
        p = h.get_positions()
        a, b, d = h.get_pairs()
        ref = oracle.cellgrid_digest(p, np.float32(w["cutoff"]), per_atom=True)
        got = oracle.digest_pairs(a, b, d)
        assert got["count"] == ref["count"] == h.pair_count()
        assert got["xor"] == ref["xor"] and got["sum"] == ref["sum"]
        assert np.array_equal(h.get_neighbor_counts(), ref["per_atom"])
    ke, pe = h.get_energies()
    assert abs((ke + pe) - (ke0 + pe0)) < 1e-3 * abs(ke0 + pe0)  # NVE over 17 steps from the un-equilibrated lattice
    h.close()


def test_c2_config_100k_nve_1000_steps(pkg):
    # BASELINE config 2: 100k-particle LJ fluid (46^3 = 97 336 atoms), rho* = 0.8442, T* = 0.72, rc = 2.5 sigma,
    # dt = 0.005 tau, velocity-Verlet NVE, 1000 steps, neighbour rebuild every step.  Drift bound: 2e-3 relative.
    import sys
    from conftest import ROOT
    sys.path.insert(0, ROOT)
    from bench import make_workload
    w = make_workload("c2")
    h = pkg.Handle(w["n"])
    h.set_forcefield(w["eps"], w["sigma"], 0.0, w["cutoff"], True)
    h.set_system(w["pos"], w["vel"], w["mass"], None)
    ke0, pe0 = h.get_energies()
    e0 = ke0 + pe0
    worst = 0.0
    for _ in range(10):
        h.step(100, w["dt"])
        ke, pe = h.get_energies()
        worst = max(worst, abs((ke + pe - e0) / e0))
    assert np.isfinite(ke + pe) and worst < 2e-3, worst
    # temperature stays physical: T* = 2 KE / (3 N) in eps units
    assert 0.3 < 2 * ke / (3 * w["n"]) < 1.5
    h.close()


def test_c5_style_clustered_gas_digest(pkg, oracle):
    # BASELINE config 5 shape at 1M atoms: half uniform background, half in Gaussian clusters (sigma 0.01),
    # clamped to the box; r chosen for ~20 neighbours on average.  Stress test of traversal imbalance:
    # parity through the independent cell-grid digest.
    rng = np.random.default_rng(5)
    n = 1_000_000
    bg = rng.random((n // 2, 3))
    centres = rng.random((512, 3))
    cl = centres[rng.integers(0, 512, n - n // 2)] + 0.01 * rng.standard_normal((n - n // 2, 3))
    x = np.clip(np.concatenate([bg, cl]), 0.0, np.nextafter(1.0, 0.0)).astype(np.float32)
    r = np.float32(0.006)
    h = pkg.Handle(n)
    cnt = h.neighbors(x, r)
    a, b, d = h.get_pairs()
    ref = oracle.cellgrid_digest(x, r)
    got = oracle.digest_pairs(a, b, d)
    assert cnt == ref["count"] == got["count"] and got["xor"] == ref["xor"] and got["sum"] == ref["sum"]
    assert 5 < 2 * cnt / n < 200
    h.close()


def test_async_step_loop_reports_a_neighbour_buffer_overflow(pkg, oracle):
    # The asynchronous step loop cannot regrow the neighbour buffer: a step that needs more slots sets a sticky device
    # flag, nb200_sync returns NB200_ERR_PAIR_OVERFLOW and regrows, and the caller reloads the system and retries.
    n = 6000
    x = uniform_positions(n, 77)
    v = ((0.5 - x) * 0.2).astype(np.float32)  # force-free collapse towards the centre: the pair count explodes
    h = pkg.Handle(n, pair_capacity_hint=64)
    h.set_forcefield(0.0, 1.0, 0.0, 0.03, True)
    h.set_system(x, v, None, None)
    p0 = h.pair_count()
    with pytest.raises(pkg.NB200Error) as e:
        h.step(4, 1.0)  # x -> 0.5 + 0.2 (x - 0.5): density x125
    assert e.value.code == pkg._lib.NB200_ERR_PAIR_OVERFLOW
    # retry after the regrow: same system, same steps
    h.set_system(x, v, None, None)
    h.step(4, 1.0)
    p = h.get_positions()
    assert np.abs(p - (x + 4 * v)).max() < 1e-5
    got = h.get_pairs()
    ref = oracle.brute_force(p, np.float32(0.03), "d2")
    assert len(got[0]) == len(ref[0]) > 20 * p0
    h.close()


def test_a_handled_overflow_of_a_search_is_not_reported_by_a_later_step_loop(pkg):
    # nb200_neighbors / nb200_set_system regrow the buffer and retry on their own; the overflow they handled must not
    # surface as NB200_ERR_PAIR_OVERFLOW of the next step loop on the same handle.
    n = 6000
    x = (0.1 + 0.8 * uniform_positions(n, 78)).astype(np.float32)   # clear of the walls: no reflection below
    h = pkg.Handle(n, pair_capacity_hint=64)
    before = h.get_stats()["regrows"]
    assert h.neighbors(x, 0.1) > 50_000                # far beyond the 4k slots the hint asked for
    assert h.get_stats()["regrows"] > before          # the search did overflow and regrow
    h.set_forcefield(0.0, 1.0, 0.0, 0.03, True)
    h.set_system(x, np.full((n, 3), 1e-4, np.float32), None, None)
    h.step(3, 1.0)                                     # used to fail with the stale overflow flag of the search
    assert np.abs(h.get_positions() - (x + np.float32(3e-4))).max() < 1e-6
    h.close()


def test_a_replaced_system_does_not_inherit_the_reports_of_unsynced_steps(pkg):
    # asynchronous steps flag problems on the device for the next nb200_sync; loading a new system before that sync
    # replaces the state they ran on, so their report must not be pinned on the new system
    n = 4096
    x = (0.1 + 0.8 * uniform_positions(n, 79)).astype(np.float32)
    v = np.full((n, 3), 1e-3, np.float32)
    h = pkg.Handle(n)
    h.set_forcefield(0.0, 1.0, 0.0, 0.03, True)
    h.set_list_reuse(1e-4, 10)
    h.set_system(x, v, None, None)
    h.step_async(10, 1.0)                           # every atom moves 1e-3 per step, skin/2 = 5e-5: flagged on the device
    h.set_system(x, np.zeros_like(v), None, None)   # never synced; a new system takes its place
    h.step(3, 1.0)                                  # must not raise
    assert np.array_equal(h.get_positions(), x)
    h.set_system(x, v, None, None)                  # the same violation IS reported to a caller who syncs
    with pytest.raises(pkg.NB200Error, match="list reuse"):
        h.step(10, 1.0)
    h.close()


def test_list_reuse_with_a_skin_matches_the_rebuild_every_step_loop(pkg, oracle):
    # nb200_set_list_reuse (list reuse across steps): the list is built with cutoff + skin every k-th step and the
    # force kernel re-applies the exact predicate at the cutoff, so every step evaluates exactly the pairs of a fresh search.
    x, a = lattice(14, 0.3, 41)
    n = len(x)
    sigma = a / 1.1
    rc = 2.5 * sigma
    rng = np.random.default_rng(42)
    v = (rng.standard_normal((n, 3)) * 0.8 * sigma).astype(np.float32)
    mass = np.full(n, 1.0 / sigma ** 2, np.float32)
    q = (0.1 * (rng.random(n) - 0.5)).astype(np.float32)
    dt = 0.004

    def fresh(reuse):
        h = pkg.Handle(n)
        h.set_forcefield(1.0, sigma, 0.2 * sigma, rc, True)
        if reuse:
            h.set_list_reuse(0.4 * sigma, 6)
        h.set_system(x, v, mass, q)
        return h
    h0, h1 = fresh(False), fresh(True)
    assert h1.pair_count() > h0.pair_count()  # the skin list is longer
    for nsteps in (1, 4, 9, 3):  # ends on steps with a fresh list and with a reused one
        h0.step(nsteps, dt)
        h1.step(nsteps, dt)
        p0, p1 = h0.get_positions(), h1.get_positions()
        assert np.abs(p1 - p0).max() < 2e-6
        # forces of the reused list against the fp64 oracle over the exact pair set of the current positions
        ra, rb, rd = oracle.brute_force(p1, np.float32(rc), "d2")
        f64, pe64, scale = oracle.forces_physical_f64(p1, q, ra, rb, 1.0, sigma, 0.2 * sigma, rc, True)
        assert (np.abs(h1.get_forces() - f64).max(axis=1) / scale).max() < FORCE_RTOL
        e0, e1 = h0.get_energies(), h1.get_energies()
        assert abs(e1[1] - pe64.sum()) <= FORCE_RTOL * np.abs(pe64).sum()
        assert abs(e1[0] - e0[0]) <= 1e-5 * abs(e0[0])
    h0.close()
    # a skin that is too small for the interval is detected on the device and reported at the sync
    h2 = pkg.Handle(n)
    h2.set_forcefield(1.0, sigma, 0.0, rc, True)
    h2.set_list_reuse(0.002 * sigma, 20)
    h2.set_system(x, v, mass, None)
    with pytest.raises(pkg.NB200Error, match="list reuse"):
        h2.step(20, dt)
    with pytest.raises(pkg.NB200Error):
        h2.set_list_reuse(0.0, 4)
    h1.close(); h2.close()


def test_a_handled_synchronous_regrow_is_not_reported_as_a_step_overflow(pkg):
    # Asynchronous steps, then — without nb200_sync — a call that searches synchronously and has to regrow the neighbour
    # buffer (a larger cutoff), then more asynchronous steps: the regrow was handled inside that call, so the next nb200_sync
    # must not report a list overflow (the sticky flag is only news for step loops that could not regrow).
    x, a = lattice(24, 0.1, 5)
    n = len(x)
    sigma = a / 1.1
    rng = np.random.default_rng(6)
    v = (rng.standard_normal((n, 3)) * 0.3 * sigma).astype(np.float32)
    mass = np.full(n, 1.0 / sigma ** 2, np.float32)
    h = pkg.Handle(n, pair_capacity_hint=8 * n)  # a small neighbour buffer: the first search regrows it to what 1.6 sigma needs
    h.set_forcefield(eps=1.0, sigma=sigma, kcoul=0.0, cutoff=1.6 * sigma, shift=True)
    h.set_system(x, v, mass, None)
    h.step_async(4, 0.002)
    cap0 = h.get_stats()["entry_capacity"]
    h.set_forcefield(eps=1.0, sigma=sigma, kcoul=0.0, cutoff=7.0 * sigma, shift=True)  # ~80x the pairs: the list must regrow
    h.step_async(4, 0.0005)
    h.sync()  # raises NB200Error on a (spurious) overflow
    assert h.get_stats()["entry_capacity"] > cap0, "the test did not force a regrow"
    assert np.isfinite(h.get_forces()).all()
    h.close()
