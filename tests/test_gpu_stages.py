"""GPU parity tests of the individual stages, through the C ABI."""
import numpy as np
import pytest

from conftest import uniform_positions

pytestmark = pytest.mark.gpu


def spread10(v):
    v = v.astype(np.uint32) & 0x3FF
    v = (v | (v << 16)) & 0x030000FF
    v = (v | (v << 8)) & 0x0300F00F
    v = (v | (v << 4)) & 0x030C30C3
    v = (v | (v << 2)) & 0x09249249
    return v


def morton30_numpy(x, lo=0.0, hi=1.0):
    scale = np.float32(1024.0) / (np.float32(hi) - np.float32(lo))
    q = np.floor((x - np.float32(lo)) * scale).astype(np.int64)
    q = np.clip(q, 0, 1023).astype(np.uint32)
    return spread10(q[:, 0]) | (spread10(q[:, 1]) << 1) | (spread10(q[:, 2]) << 2)


def hilbert30_numpy(x, lo=0.0, hi=1.0, bits=10):
    """Skilling's axes-to-transpose + interleave: numpy twin of hilbert_interleave (csrc/atoms.cu)."""
    scale = np.float32(1024.0) / (np.float32(hi) - np.float32(lo))
    q = np.clip(np.floor((x - np.float32(lo)) * scale).astype(np.int64), 0, 1023).astype(np.uint32)
    X = [q[:, 0].copy(), q[:, 1].copy(), q[:, 2].copy()]
    Q = np.uint32(1 << (bits - 1))
    while Q > 1:
        P = np.uint32(Q - 1)
        for i in range(3):
            inv = (X[i] & Q) != 0
            t = np.where(inv, np.uint32(0), (X[0] ^ X[i]) & P)
            X[0] = np.where(inv, X[0] ^ P, X[0]) ^ t
            if i:
                X[i] = X[i] ^ t
        Q = np.uint32(Q >> 1)
    X[1] ^= X[0]
    X[2] ^= X[1]
    t = np.zeros_like(X[0])
    Q = np.uint32(1 << (bits - 1))
    while Q > 1:
        t = np.where((X[2] & Q) != 0, t ^ np.uint32(Q - 1), t)
        Q = np.uint32(Q >> 1)
    X = [v ^ t for v in X]
    return (spread10(X[0]) << 2) | (spread10(X[1]) << 1) | spread10(X[2])


def hilbert_is_a_curve(bits=4):
    """Property of the twin itself: consecutive Hilbert indices are face-neighbouring lattice points."""
    m = 1 << bits
    g = np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij"), -1).reshape(-1, 3)
    x = ((g + 0.5) / 1024.0).astype(np.float32)          # the first m lattice cells per axis of the 1024 grid
    key = hilbert30_numpy(x)
    order = np.argsort(key)
    step = np.abs(np.diff(g[order], axis=0)).sum(1)
    return len(np.unique(key)) == m ** 3 and np.all(step == 1)


def test_hilbert_twin_is_a_hilbert_curve():
    assert hilbert_is_a_curve(3) and hilbert_is_a_curve(4)


@pytest.mark.parametrize("n", [1, 33, 1000, 100_003])
def test_hilbert_keys_match_numpy(big_handle, n):
    x = uniform_positions(n, 41 + n)
    if n > 2:
        x[1] = [1.0, 1.0, 1.0]
        x[2] = [-0.5, 2.0, 0.5]
    big_handle.set_curve(1)
    assert np.array_equal(big_handle.sort_keys(x), hilbert30_numpy(x))
    big_handle.set_curve(0)
    assert np.array_equal(big_handle.sort_keys(x), morton30_numpy(x))
    big_handle.set_curve(1)


@pytest.mark.parametrize("n", [1, 31, 32, 33, 1000, 100_003])
def test_morton30_matches_numpy(big_handle, n):
    x = uniform_positions(n, 11 + n)
    x[0] = [0.0, 0.0, 0.0]
    if n > 2:
        x[1] = [1.0, 1.0, 1.0]      # clamps to 1023
        x[2] = [-0.5, 2.0, 0.5]     # out of box: clamped
    assert np.array_equal(big_handle.morton30(x), morton30_numpy(x))
    x4 = np.concatenate([x, np.ones((n, 1), np.float32)], 1)
    assert np.array_equal(big_handle.morton30(x4), morton30_numpy(x))


@pytest.mark.parametrize("n", [1, 2, 255, 4096, 4097, 70_001, 1_000_000])
@pytest.mark.parametrize("kind", ["random32", "morton30", "fewkeys"])
def test_radix_sort_is_a_stable_sort(big_handle, n, kind):
    rng = np.random.default_rng(n * 7 + len(kind))
    if kind == "random32":
        keys = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    elif kind == "morton30":
        keys = morton30_numpy(rng.random((n, 3)).astype(np.float32))
    else:
        keys = rng.integers(0, 5, n).astype(np.uint32) * np.uint32(0x01010101)
    vals = np.arange(n, dtype=np.uint32)
    k2, v2 = big_handle.sort_pairs(keys, vals)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k2, keys[order])
    assert np.array_equal(v2, order.astype(np.uint32))


def check_tree(tree, n_atoms, positions_sorted=None):
    nL = tree["n_leaves"]
    child, nbox, lbox = tree["node_child"], tree["node_box"], tree["leaf_box"]
    assert nL == (n_atoms + 31) // 32
    if nL == 1:
        return
    assert tree["root"] == 0
    seen_leaf = np.zeros(nL, bool)
    seen_node = np.zeros(nL - 1, bool)
    stack = [0]
    while stack:
        k = stack.pop()
        assert not seen_node[k]
        seen_node[k] = True
        left, right, first, last = child[k]
        lo, hi = nbox[k, :3], nbox[k, 3:]
        covered = []
        for c in (left, right):
            if c < 0:
                leaf = ~c
                assert not seen_leaf[leaf]
                seen_leaf[leaf] = True
                cb = lbox[leaf]
                covered.append((leaf, leaf))
            else:
                cb = nbox[c]
                covered.append((child[c][2], child[c][3]))
                stack.append(c)
            assert np.all(cb[:3] >= lo) and np.all(cb[3:] <= hi)
        # children tile the parent's leaf range, in order
        assert covered[0][0] == first and covered[1][1] == last and covered[0][1] + 1 == covered[1][0]
        # the parent box is exactly the union
        boxes = [lbox[~c] if c < 0 else nbox[c] for c in (left, right)]
        assert np.array_equal(np.minimum(boxes[0][:3], boxes[1][:3]), lo)
        assert np.array_equal(np.maximum(boxes[0][3:], boxes[1][3:]), hi)
    assert seen_leaf.all() and seen_node.all()
    assert child[0][2] == 0 and child[0][3] == nL - 1


@pytest.mark.parametrize("curve", [1, 0])
@pytest.mark.parametrize("n", [33, 64, 1000, 4097, 100_000])
def test_tree_is_a_valid_hierarchy(big_handle, n, curve):
    x = uniform_positions(n, 90 + n)
    big_handle.set_curve(curve)
    big_handle.neighbors(x, 0.01)
    tree = big_handle.get_tree()
    check_tree(tree, n)
    ids = big_handle.get_sorted_ids()
    assert sorted(ids.tolist()) == list(range(n))
    keys = (hilbert30_numpy if curve else morton30_numpy)(x[ids])
    big_handle.set_curve(1)
    # the pipeline sorts by the key bits the atom count needs: log2(n)+4 bits, whole 8-bit passes from the top
    bits = 4
    while (1 << (bits - 4)) < n and bits < 30:
        bits += 1
    passes = min(4, max(2, (bits + 7) // 8))
    keys = keys.astype(np.int64) >> (0 if passes == 4 else 30 - 8 * passes)
    assert np.all(np.diff(keys) >= 0)                            # sorted by the curve's key
    same = np.diff(keys) == 0
    assert np.all(np.diff(ids)[same] > 0)                        # stable: ties keep original order
    # leaf boxes are the tight bounds of their 32 atoms
    xs = x[ids]
    for leaf in (0, tree["n_leaves"] // 2, tree["n_leaves"] - 1):
        seg = xs[leaf * 32:(leaf + 1) * 32]
        assert np.array_equal(tree["leaf_box"][leaf, :3], seg.min(0)) and np.array_equal(tree["leaf_box"][leaf, 3:], seg.max(0))


def test_tree_with_many_duplicate_keys(big_handle):
    # all atoms in one Morton cell: the index tie-break (BVHTraverse.jl:675) must still give a valid tree
    rng = np.random.default_rng(5)
    x = (0.5 + 1e-5 * rng.random((5000, 3))).astype(np.float32)
    big_handle.neighbors(x, 1e-6)
    check_tree(big_handle.get_tree(), 5000)
    for _ in range(20):  # race detector, like the reference's 100 rebuilds (test/BVHTraverse.jl:189-191)
        big_handle.neighbors(uniform_positions(3000, 77), 0.02)
        check_tree(big_handle.get_tree(), 3000)


def test_argument_and_state_errors_of_the_tuning_and_loop_calls(pkg):
    """Every entry point reports bad arguments / call order through its status code and nb200_last_error (the
    reference's error("...") style, BVHTraverse.jl:75,79,286) instead of crashing."""
    E = pkg._lib
    h = pkg.Handle(4096)
    for call, code in ((lambda: h.set_list_mode(7), E.NB200_ERR_BAD_ARG), (lambda: h.set_resort_interval(0), E.NB200_ERR_BAD_ARG),
                       (lambda: h.set_curve(3), E.NB200_ERR_BAD_ARG),
                       (lambda: h.rescale_velocity(1.0, 1.0), E.NB200_ERR_STATE),       # no system loaded
                       (lambda: h.simulate(3, 0.1), E.NB200_ERR_STATE),
                       (lambda: h.mg_search_force(), E.NB200_ERR_STATE),                # nb200_mg_set_owned not called
                       (lambda: h.mg_search_force_async(), E.NB200_ERR_STATE),
                       (lambda: h.mg_sync(), E.NB200_ERR_STATE),
                       (lambda: h.mg_publication(), E.NB200_ERR_STATE)):
        with pytest.raises(pkg.NB200Error) as e:
            call()
        assert e.value.code == code, e.value
    x = uniform_positions(1000, 3)
    h.set_forcefield(0.0, 1.0, 0.0, 0.05, True)
    h.set_system(x, np.zeros_like(x), None, None)
    with pytest.raises(pkg.NB200Error) as e:  # poslog too small for the frames the run logs
        h.simulate(10, 0.1, log_every=2, out=np.empty((2, 1000, 3), np.float32))
    assert e.value.code == E.NB200_ERR_CAPACITY
    with pytest.raises(pkg.NB200Error) as e:
        h.leapfrog_host_async(0, 0, 3, 1000, 0.1, False)  # xyz is NULL (vel may be: positions-only exchange)
    assert e.value.code == E.NB200_ERR_BAD_ARG
    # multi-GPU slab tables must describe this handle
    h.mg_set_owned(x)
    with pytest.raises(pkg.NB200Error) as e:
        h.mg_connect(2, 0, [0, 1000], [999, 1000], direct_base=[h.mg_publication()[0], h.mg_publication()[0]])
    assert e.value.code == E.NB200_ERR_BAD_ARG
    with pytest.raises(pkg.NB200Error) as e:
        h.mg_connect(2, 0, [0, 1000], [1000, 1000])  # neither a pointer nor an IPC handle for peer 1
    assert e.value.code == E.NB200_ERR_BAD_ARG
    # world 1 needs no connection: search, then an asynchronous step
    ng, ne = h.mg_search_force()
    assert ng == 0 and ne >= 0
    h.mg_integrate(0.01)
    h.mg_search_force_async()
    assert h.mg_sync() == (0, ne)
    h.close()
