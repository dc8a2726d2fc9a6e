"""Host-side layout generators vs post-reset states recorded from the reference
(tests/golden/*.npz hold init_grid/init_agents produced with known generator seeds)."""
import numpy as np
import pytest

from multigrid_b200 import layouts as L
from multigrid_b200.core.constants import Direction
from oracle import mg_oracle as O
from tests.golden_util import load_case


def rngs(seed, b):
    """The deterministic-oracle recipe of tests/golden/make_golden.py::make_env."""
    return (np.random.default_rng(seed * 1000 + b),
            np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed * 7919 + b))))


CASES = [
    ("empty8_n4", 2, lambda n: L.EmptyLayout(n, size=8)),
    ("empty16_n8_v9", 4, lambda n: L.EmptyLayout(n, size=16)),
    ("empty6r_n3_nooverlap_all", 5,
     lambda n: L.EmptyLayout(n, size=6, agent_start_pos=None, agent_start_dir=None)),
    ("empty5_n1", 6, lambda n: L.EmptyLayout(n, size=5)),
    ("bup_n2", 3, lambda n: L.BlockedUnlockPickupLayout(n)),
    ("playground_n3", 8, lambda n: L.PlaygroundLayout(n)),
    ("rbd_n2_autoreset", 42, lambda n: L.RedBlueDoorsLayout(n, size=6, max_steps=40)),
    ("lh6_n4", 52, lambda n: L.LockedHallwayLayout(n, num_rooms=6)),
]


@pytest.mark.parametrize("name,seed,make", CASES)
def test_layout_matches_reference_reset(name, seed, make):
    d, meta = load_case(name)
    layout = make(meta["n"])
    assert (layout.width, layout.height, layout.max_steps) == (meta["W"], meta["H"], meta["max_steps"])
    assert layout.hook == meta["hook"]
    M = (1 << 64) - 1
    for b in range(meta["B"]):
        layout_rng, order_rng = rngs(seed, b)
        grid, agents, _ = layout.generate(layout_rng, order_rng)
        np.testing.assert_array_equal(grid, d["init_grid"][b], err_msg=f"{name} env {b}")
        np.testing.assert_array_equal(O.unpack_agents(agents), d["init_agents"][b])
        # the order stream must be left exactly where the reference's reset leaves it
        st = order_rng.bit_generator.state["state"]["state"]
        assert [st & M, st >> 64] == [int(v) for v in d["pcg_state"][b]]


def test_generate_pool_shapes_and_determinism():
    layout = L.BlockedUnlockPickupLayout(2)
    g1, a1, info = L.generate_pool(layout, 5, seed=7)
    g2, a2, _ = L.generate_pool(layout, 5, seed=7)
    assert g1.shape == (5, 11, 6, 3) and a1.shape == (5, 2, 8)
    np.testing.assert_array_equal(g1, g2)
    np.testing.assert_array_equal(a1, a2)
    assert info[0]["mission"].startswith("pick up the ")
    assert len({g.tobytes() for g in g1}) > 1
    g, a, _ = L.generate_pool(L.EmptyLayout(4), 100, seed=0)
    assert g.shape[0] == 1 and (a[0, :, 1:3] == 1).all() and (a[0, :, 0] == int(Direction.right)).all()


def _numpy_rngs(K, seed, odd_buffer):
    gens = [np.random.default_rng([seed, k]) for k in range(K)]
    if odd_buffer:  # leave a buffered upper half in every second generator (numpy's has_uint32 state)
        for g in gens[::2]:
            g.integers(0, 10)
    return gens


@pytest.mark.parametrize("size,n,seed", [(5, 1, 0), (5, 4, 1), (6, 3, 2), (8, 8, 3), (16, 12, 4), (7, 20, 5), (3, 1, 6)])
@pytest.mark.parametrize("odd_buffer", [False, True])
def test_device_layout_function_matches_host_generator(size, n, seed, odd_buffer):
    """gen_layout_empty_random (the function the CUDA layout kernel runs; here compiled for the CPU by
    tests/hostsim) against EmptyLayout.generate driven by real numpy generators: same grids, same agent
    records, same generator state afterwards (PCG64 state AND the buffered 32-bit half) -- i.e. the
    in-kernel Generator.integers (Lemire + pcg64_next32 buffering) is numpy's."""
    from multigrid_b200.env import layout_generator_words
    from tests.hostsim.sim import gen_layouts_empty_random
    if n > (size - 2) ** 2 - 1:
        pytest.skip("more agents than free cells")
    K = 300
    layout = L.EmptyLayout(n, size=size, agent_start_pos=None, agent_start_dir=None)
    st, inc, buf = layout_generator_words(_numpy_rngs(K, seed, odd_buffer))
    grid, agents, st2, buf2 = gen_layouts_empty_random(size, size, n, st, inc, buf)
    host = _numpy_rngs(K, seed, odd_buffer)
    for k in range(K):
        g, a, _ = layout.generate(host[k], None)
        np.testing.assert_array_equal(grid[k], g, err_msg=f"layout {k}")
        np.testing.assert_array_equal(agents[k], a, err_msg=f"layout {k}")
    st_h, _, buf_h = layout_generator_words(host)
    np.testing.assert_array_equal(st2, st_h)
    np.testing.assert_array_equal(buf2, buf_h)
    assert len({a.tobytes() for a in agents}) > 1


def test_in_kernel_integers_matches_numpy_for_awkward_ranges():
    """Ranges whose Lemire threshold is non-zero (3, 5, 6, 7, 14 cells ...) exercise the rejection loop."""
    from multigrid_b200.env import layout_generator_words
    from tests.hostsim.sim import gen_layouts_empty_random
    for size in (5, 7, 9, 16, 100):  # integers(0, size): thresholds (2^32 - size) % size
        K = 2000
        gens = [np.random.default_rng([99, size, k]) for k in range(K)]
        st, inc, buf = layout_generator_words(gens)
        grid, agents, st2, buf2 = gen_layouts_empty_random(size, size, 1, st, inc, buf)
        layout = L.EmptyLayout(1, size=size, agent_start_pos=None, agent_start_dir=None)
        for k in range(0, K, 7):
            g, a, _ = layout.generate(gens[k], None)
            np.testing.assert_array_equal(agents[k], a)


@pytest.mark.parametrize("S,n,seed", [(6, 2, 0), (6, 5, 1), (4, 1, 2), (5, 3, 3), (9, 8, 4)])
def test_device_bup_layout_function_matches_host_generator(S, n, seed):
    """gen_layout_bup (the function of the BUP layout kernel, CPU build) against
    BlockedUnlockPickupLayout.generate (pinned to the reference's reset by the fixtures above) with real
    numpy generators: grid, agents, box colour, both generators' states afterwards."""
    from multigrid_b200.env import layout_generator_words
    from tests.hostsim.sim import gen_layouts_bup
    K = 400
    mk_l = lambda: [np.random.default_rng([seed, k]) for k in range(K)]                      # noqa: E731
    mk_o = lambda: [np.random.Generator(np.random.PCG64(np.random.SeedSequence(1000 * seed + k))) for k in range(K)]  # noqa: E731
    lg, og = mk_l(), mk_o()
    for g in lg[::2]:
        g.integers(0, 10)  # a buffered 32-bit half in every second layout generator
    st, inc, buf = layout_generator_words(lg)
    ost, oinc, _ = layout_generator_words(og)
    grid, agents, st2, buf2, ost2, info = gen_layouts_bup(S, n, st, inc, buf, ost, oinc)
    layout = L.BlockedUnlockPickupLayout(n, room_size=S)
    colors = [c.value for c in L._COLORS]
    for k in range(K):
        g, a, inf = layout.generate(lg[k], og[k])
        np.testing.assert_array_equal(grid[k], g, err_msg=f"layout {k}")
        np.testing.assert_array_equal(agents[k], a, err_msg=f"layout {k}")
        assert inf["mission"] == f"pick up the {colors[info[k]]} box"
    st_h, _, buf_h = layout_generator_words(lg)
    ost_h, _, _ = layout_generator_words(og)
    np.testing.assert_array_equal(st2, st_h)
    np.testing.assert_array_equal(buf2, buf_h)
    np.testing.assert_array_equal(ost2, ost_h)


@pytest.mark.parametrize("size,n,seed", [(6, 2, 0), (8, 3, 1), (4, 4, 2), (11, 7, 3)])
def test_device_rbd_layout_function_matches_host_generator(size, n, seed):
    """gen_layout_red_blue_doors (CPU build of the kernel's function) vs RedBlueDoorsLayout.generate."""
    from multigrid_b200.env import layout_generator_words
    from tests.hostsim.sim import gen_layouts_red_blue_doors
    K = 400
    lg = [np.random.default_rng([seed, k, 5]) for k in range(K)]
    for g in lg[1::2]:
        g.integers(0, 3)
    st, inc, buf = layout_generator_words(lg)
    grid, agents, st2, buf2 = gen_layouts_red_blue_doors(size, n, st, inc, buf)
    layout = L.RedBlueDoorsLayout(n, size=size)
    for k in range(K):
        g, a, _ = layout.generate(lg[k], None)
        np.testing.assert_array_equal(grid[k], g, err_msg=f"layout {k}")
        np.testing.assert_array_equal(agents[k], a, err_msg=f"layout {k}")
    st_h, _, buf_h = layout_generator_words(lg)
    np.testing.assert_array_equal(st2, st_h)
    np.testing.assert_array_equal(buf2, buf_h)


@pytest.mark.parametrize("rooms,S,mhk,mkpr,n,seed", [(6, 5, 1, 2, 4, 0), (2, 5, 1, 2, 2, 1), (4, 4, 2, 1, 3, 2),
                                                      (6, 6, 3, 3, 8, 3), (4, 7, 1, 4, 1, 4)])
def test_device_locked_hallway_layout_function_matches_host_generator(rooms, S, mhk, mkpr, n, seed):
    """gen_layout_locked_hallway (CPU build of the kernel's function; in-kernel Generator.shuffle) vs
    LockedHallwayLayout.generate with real numpy generators."""
    from multigrid_b200.env import layout_generator_words
    from tests.hostsim.sim import gen_layouts_locked_hallway
    K = 400
    lg = [np.random.default_rng([seed, k, 9]) for k in range(K)]
    for g in lg[1::2]:
        g.integers(0, 3)
    st, inc, buf = layout_generator_words(lg)
    grid, agents, st2, buf2 = gen_layouts_locked_hallway(rooms, S, mhk, mkpr, n, st, inc, buf)
    layout = L.LockedHallwayLayout(n, num_rooms=rooms, room_size=S, max_hallway_keys=mhk, max_keys_per_room=mkpr)
    for k in range(K):
        g, a, _ = layout.generate(lg[k], None)
        np.testing.assert_array_equal(grid[k], g, err_msg=f"layout {k}")
        np.testing.assert_array_equal(agents[k], a, err_msg=f"layout {k}")
    st_h, _, buf_h = layout_generator_words(lg)
    np.testing.assert_array_equal(st2, st_h)
    np.testing.assert_array_equal(buf2, buf_h)


@pytest.mark.parametrize("S,rows,cols,n,seed", [(7, 3, 3, 3, 0), (7, 3, 3, 8, 1), (7, 2, 4, 2, 2), (6, 4, 4, 5, 3), (8, 1, 2, 1, 4)])
def test_device_playground_layout_function_matches_host_generator(S, rows, cols, n, seed):
    """gen_layout_playground (CPU build of the kernel's function: connect_all, add_object, place_in_room with
    both generators) vs PlaygroundLayout.generate with real numpy generators."""
    from multigrid_b200.env import layout_generator_words
    from tests.hostsim.sim import gen_layouts_playground
    K = 250
    lg = [np.random.default_rng([seed, k, 3]) for k in range(K)]
    og = [np.random.Generator(np.random.PCG64(np.random.SeedSequence(500 * seed + k))) for k in range(K)]
    for g in lg[::2]:
        g.integers(0, 10)
    st, inc, buf = layout_generator_words(lg)
    ost, oinc, _ = layout_generator_words(og)
    grid, agents, st2, buf2, ost2 = gen_layouts_playground(S, rows, cols, n, st, inc, buf, ost, oinc)
    layout = L.PlaygroundLayout(n, room_size=S, num_rows=rows, num_cols=cols)
    for k in range(K):
        g, a, _ = layout.generate(lg[k], og[k])
        np.testing.assert_array_equal(grid[k], g, err_msg=f"layout {k}")
        np.testing.assert_array_equal(agents[k], a, err_msg=f"layout {k}")
    st_h, _, buf_h = layout_generator_words(lg)
    ost_h, _, _ = layout_generator_words(og)
    np.testing.assert_array_equal(st2, st_h)
    np.testing.assert_array_equal(buf2, buf_h)
    np.testing.assert_array_equal(ost2, ost_h)
