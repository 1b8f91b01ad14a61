"""Poisson inputs (InputModel, src/models.cpp:863-903; generator seeding src/models.hpp:347,366), CPU side:
the product draws the random spikes on the host with libstdc++'s std::mt19937 (csrc/host/poisson.cpp) and
hands them to the device as an overlay. Here the overlay drives the CPU restatement instead of the device and
the result is compared with the reference's own run (tests/golden/poisson.*)."""
import ctypes as C

import numpy as np

import sanafe_b200 as sfe
from helpers import Oracle, check_against_golden, golden, load_chip


def product_overlay(chip, steps):
    src = sfe.lib().sfe_poisson_create(C.byref(chip.tables))
    assert src, sfe.lib().sfe_last_error()
    cols = sfe.lib().sfe_poisson_cols(src)
    bits = np.zeros((steps, cols), dtype=np.uint8)
    assert sfe.lib().sfe_poisson_fill(src, bits.ctypes.data, steps) == 0
    return src, bits


def test_lowering_assigns_unit_ordinals_and_columns():
    chip = load_chip("poisson", device=-1)
    t = chip.tables
    assert t.n_inputs == 3 and t.n_poisson_cols == 3 and t.input_seed_base == 0
    descs = [t.inputs[k] for k in range(3)]
    # in.0 and in.1 share core 0.0's demo_input (1st InputModel of the chip), in.2 sits on core 1.2 (7th)
    assert [(d.unit, d.share_count, d.share_rank) for d in descs] == [(0, 2, 0), (0, 2, 1), (6, 1, 0)]
    assert sorted(d.poisson_col for d in descs) == [0, 1, 2]
    assert [d.poisson for d in descs] == [0.3, 0.3, 0.65]


def test_product_draws_reproduce_the_reference_run():
    chip = load_chip("poisson", device=-1)
    g = golden("poisson")
    src, bits = product_overlay(chip, g["steps"])
    oracle = Oracle(chip)
    oracle.set_input_overlay(bits)
    rd, out = oracle.run(g["steps"])
    check_against_golden("poisson", chip, rd, out, potential_rtol=0.0, energy_rtol=1e-12)
    sfe.lib().sfe_poisson_destroy(src)


def test_streams_continue_across_fills_and_agree_with_the_restatement():
    """Two fills == one fill (the generators are not rewound), and the product's libstdc++ draws equal the
    restatement's own plain-C MT19937 spike for spike."""
    chip = load_chip("poisson", device=-1)
    src, whole = product_overlay(chip, 120)
    sfe.lib().sfe_poisson_destroy(src)
    src = sfe.lib().sfe_poisson_create(C.byref(chip.tables))
    a = np.zeros((50, 3), dtype=np.uint8)
    b = np.zeros((70, 3), dtype=np.uint8)
    assert sfe.lib().sfe_poisson_fill(src, a.ctypes.data, 50) == 0
    assert sfe.lib().sfe_poisson_fill(src, b.ctypes.data, 70) == 0
    sfe.lib().sfe_poisson_destroy(src)
    assert np.array_equal(whole, np.concatenate([a, b]))
    with_overlay, own = Oracle(chip), Oracle(chip)
    with_overlay.set_input_overlay(whole)
    _, o1 = with_overlay.run(120)
    _, o2 = own.run(120)
    assert np.array_equal(o1["fired_bits"], o2["fired_bits"])
    assert 0.2 < whole[:, :2].mean() < 0.4 and 0.55 < whole[:, 2].mean() < 0.75


def test_seed_base_shifts_the_streams():
    """The reference seeds by a process-wide construction counter: a chip created after another one that
    instantiated 8 input units draws different spikes (set_input_seed_base reproduces either situation)."""
    import os
    from helpers import ROOT, golden_flat
    cwd = os.getcwd()
    os.chdir(ROOT)
    try:
        arch, net = sfe.load_flat(golden_flat("poisson"))
        chip = sfe.SpikingChip(arch, device=-1)
        chip.set_input_seed_base(8)
        chip.load(net)
    finally:
        os.chdir(cwd)
    assert chip.tables.input_seed_base == 8
    src, shifted = product_overlay(chip, 64)
    sfe.lib().sfe_poisson_destroy(src)
    base = load_chip("poisson", device=-1)
    src, first = product_overlay(base, 64)
    sfe.lib().sfe_poisson_destroy(src)
    assert not np.array_equal(shifted, first)
    # unit 0 of a chip with base 8 has the seed (9) that ordinal 8 would have with base 0: the restatement agrees
    _, o1 = Oracle(chip).run(64)
    with_overlay = Oracle(chip)
    with_overlay.set_input_overlay(shifted)
    _, o2 = with_overlay.run(64)
    assert np.array_equal(o1["fired_bits"], o2["fired_bits"])


def test_engine_generator_equals_libstdcxx():
    """csrc/mt19937.cuh (one source for host and device; the experimental device-side draws use it) against
    std::mt19937 + std::uniform_real_distribution<double>: same doubles, with the state interleaved or not."""
    L = sfe.lib()
    for seed in (1, 2, 7, 1000, 4294967295):
        for stride in (1, 3, 128):
            want, got = np.zeros(5000), np.zeros(5000)
            L.sfe_poisson_reference_draws(seed, want.ctypes.data, want.size)
            L.sfe_mt19937_draws(seed, got.ctypes.data, got.size, stride)
            assert np.array_equal(want, got), (seed, stride)
            assert 0.0 <= got.min() and got.max() < 1.0


def test_glibc_rand_restatement_equals_libc():
    """The TrueNorth threshold jitter draws `std::rand() & random_mask` (src/models.cpp:757): the product restates
    glibc's generator (csrc/host/poisson.cpp) instead of touching the process-global one. Checked against libc itself."""
    import ctypes
    libc = ctypes.CDLL("libc.so.6")
    libc.rand.restype = ctypes.c_int
    for seed in (1, 12345):
        libc.srand(seed)
        want = np.array([libc.rand() for _ in range(2000)], dtype=np.uint32)
        got = np.zeros(2000, dtype=np.uint32)
        sfe.lib().sfe_glibc_rand_draws(seed, 0, got.ctypes.data, got.size)
        assert np.array_equal(got, want), seed
    tail = np.zeros(100, dtype=np.uint32)
    sfe.lib().sfe_glibc_rand_draws(12345, 1900, tail.ctypes.data, tail.size)
    assert np.array_equal(tail, want[1900:])
