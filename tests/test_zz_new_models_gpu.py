"""Device tests of what was added late in round 1: Poisson inputs, LIF file noise, model-defined neuron traces
(sim(neuron_trace=...), `sim -x`) and the batched design-space sweep. Their CPU side — lowering, host draws,
restatement vs the reference's goldens — is pinned in test_oracle_vs_reference.py, test_poisson_inputs.py and
test_dse_batch.py. Collected last (see helpers.NEW_GOLDEN_CASES).

The first nine tests have run green on a B200 (profiles/r1_pytest_new_models_gpu.log, r1_pytest_new_models_gpu_2.log).
The ones after them (the reference's unit-test vectors on the device, Ctrl-C, the two experimental paths) were written
once the round's GPU minutes were spent: their CPU halves are pinned (tests/test_reference_unit_vectors.py,
test_python_builders.py, test_poisson_inputs.py), the device halves have not executed yet; the two marked xfail drive
code paths that are off by default."""
import numpy as np
import pytest

import sanafe_b200 as sfe
from helpers import NEW_GOLDEN_CASES, Oracle, check_against_golden, golden, load_chip

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", NEW_GOLDEN_CASES)
def test_engine_matches_reference_new_cases(name):
    chip = load_chip(name, device=0)
    g = golden(name)
    rd, out = chip.sim_raw(g["steps"], "simple", steps=True, fired=True, potentials=True, neuron_traces=True)
    assert ("neuron_traces" in out) == ("neuron_traces_shape" in g)
    check_against_golden(name, chip, rd, out, potential_rtol=0.0, energy_rtol=1e-9)


def test_poisson_sim_calls_continue_the_streams():
    """sim(100) + reset() + sim(200) draws the same random spikes as one sim(300): neither the generators nor
    the spike-train cursor are rewound by reset (InputModel::reset, src/models.hpp:358)."""
    a = load_chip("poisson", device=0)
    b = load_chip("poisson", device=0)
    _, out_a = a.sim_raw(300, steps=True, fired=True)
    _, out_b1 = b.sim_raw(100, "detailed", steps=True, fired=True)
    _, out_b2 = b.sim_raw(200, steps=True, fired=True)
    n_in = 3  # in.0..2 are device neurons of cores 0.0 / 1.2; compare the input neurons' raster columns only
    t = a.tables
    in_idx = [a.neuron_index("in", k) for k in range(n_in)]
    whole = out_a["fired_bits"]
    parts = np.concatenate([out_b1["fired_bits"], out_b2["fired_bits"]])
    for i in in_idx:
        assert np.array_equal((whole[:, i >> 5] >> (i & 31)) & 1, (parts[:, i >> 5] >> (i & 31)) & 1)
    assert np.array_equal(whole, parts)
    assert t.n_poisson_cols == 3


def test_poisson_engine_refuses_to_step_without_an_overlay():
    chip = load_chip("poisson", device=0)
    assert sfe.lib().sfe_engine_enqueue(chip.engine, 1) != 0
    assert b"Poisson" in sfe.lib().sfe_last_error()


@pytest.mark.parametrize("mode", ["grid", "streams"])
def test_dse_batch_equals_lone_runs_and_the_restatement(tmp_path, monkeypatch, mode):
    """BASELINE config 5: chips stepped side by side by sfe_batch_sim give, bit for bit, what each gives when
    run alone, and what the CPU restatement gives for the same design point. grid: the batch is an outer grid
    dimension of the step kernels (one launch per phase and step for all chips, sfe_engine_batch_enqueue);
    streams: one stream per chip (SFE_BATCH_GRID=0)."""
    from sanafe_b200 import dse
    monkeypatch.setenv("SFE_BATCH_GRID", "1" if mode == "grid" else "0")
    points = [(64, 1.0), (64, 2.0), (200, 1.0), (200, 0.5), (32, 1.5), (32, 1.0)]
    steps = 40
    batch = dse.Sweep(points, str(tmp_path / "batch"), device=0, host_threads=4)
    rds = batch.sim(steps)
    rds2 = batch.sim(steps)  # state persists across batched calls like across sim() calls
    lone = dse.Sweep(points, str(tmp_path / "lone"), device=0, host_threads=4, share=batch)
    host = dse.Sweep(points, str(tmp_path / "host"), device=-1, host_threads=4, share=batch)
    for k, point in enumerate(points):
        rd_l, _ = lone.chips[k].sim_raw(steps)
        rd_l2, _ = lone.chips[k].sim_raw(steps)
        oracle = Oracle(host.chips[k])
        rd_o, _ = oracle.run(steps)
        rd_o2, _ = oracle.run(steps)
        for got, alone, want in ((rds[k], rd_l, rd_o), (rds2[k], rd_l2, rd_o2)):
            for key in ("timestep_start", "spikes", "packets_sent", "neurons_updated", "neurons_fired"):
                assert getattr(got, key) == getattr(alone, key) == getattr(want, key), (point, key)
            for key in ("total_energy", "synapse_energy", "soma_energy", "network_energy", "sim_time"):
                assert getattr(got, key) == getattr(alone, key), (point, key)
                assert abs(getattr(got, key) - getattr(want, key)) <= 1e-9 * abs(getattr(want, key)), (point, key)
        assert rds[k].spikes > 0


def test_neuron_trace_surfaces(tmp_path):
    """Model-defined traces (LIF `u` of log_u neurons) through the three user surfaces: the ctypes binding's and the
    pybind11 module's sim(neuron_trace=...) and the command line's -x (neurons.csv, src/chip.cpp:1478-1517, 1664-1702)."""
    import os
    import subprocess
    from helpers import GOLDEN, ROOT, golden_flat
    ref = np.load(os.path.join(GOLDEN, "noise.neuron_traces.npy"))
    chip = load_chip("noise", device=0)
    assert chip.trace_names() == ["a.0/u", "a.1/u", "a.2/u", "a.3/u", "a.4/u", "b.0/u", "q.0/u"]
    res = chip.sim(150, timing_model="detailed", neuron_trace=True, spike_trace=True)
    assert np.array_equal(np.asarray(res["neuron_trace"]["u"]), ref)
    # pybind11 module, file sink
    from sanafe_b200 import sanafecpp_b200 as m
    cwd = os.getcwd()
    os.chdir(ROOT)
    try:
        arch, net = m.load_flat(golden_flat("noise"))
        pchip = m.SpikingChip(arch)
        pchip.load(net)
        path = str(tmp_path / "neurons.csv")
        out = pchip.sim(150, timing_model="simple", neuron_trace=path)
        mem = pchip.sim(5, timing_model="simple", neuron_trace=True)
    finally:
        os.chdir(cwd)
    lines = open(path).read().splitlines()
    assert lines[0] == "timestep," + "".join(f"neuron {n}," for n in chip.trace_names())
    got = np.asarray([[float(x) for x in ln.split(",")[1:-1]] for ln in lines[1:]])
    assert got.shape == ref.shape and np.allclose(got, ref, rtol=1e-5, atol=1e-12)  # CSV keeps 6 digits
    assert out["timesteps_executed"] == 150 and len(mem["neuron_trace"]["u"]) == 5
    # command line: -x next to the YAML descriptions
    src = os.path.join(GOLDEN, "src")
    sim = os.path.join(ROOT, "sana-fe_b200", "sanafe_b200", "sim")
    cli = subprocess.run([sim, "-x", "-o", str(tmp_path / "cli"), os.path.join(src, "noise_arch.yaml"),
                          os.path.join(src, "noise_snn.yaml"), "150"], capture_output=True, text=True, timeout=120, cwd=ROOT)
    assert cli.returncode == 0, cli.stderr
    cli_lines = (tmp_path / "cli" / "neurons.csv").read_text().splitlines()
    assert cli_lines == lines


def test_taps_on_a_partitioned_chip_are_refused_loudly():
    """`taps` lines are replayed from the whole chip's raster by one kernel: a partitioned chip must refuse them with
    a message, never run something else."""
    with pytest.raises(sfe.SanafeError, match="taps"):
        load_chip("taps", device=0, partition=(0, 2))


def test_per_neuron_bias_patches_between_sim_calls():
    """The DVS-gesture loop of the reference (scripts/tcad2025/dvs_gesture.py): MappedNeuron.set_attributes(bias)
    on many neurons, then sim(). The patches are collected on the host and uploaded as one vector before the next
    step; results must equal the CPU restatement given the same biases."""
    dev = load_chip("synth_soma", device=0)
    host = load_chip("synth_soma", device=-1)
    oracle = Oracle(host)
    t = host.tables
    n = t.n_neurons
    bias = np.array([t.neuron_bias[i] for i in range(n)])
    rng = np.random.default_rng(3)
    for round_ in range(3):
        rd_d, out_d = dev.sim_raw(8, steps=True, fired=True, potentials=True)
        rd_h, out_h = oracle.run(8)
        assert np.array_equal(out_d["fired_bits"], out_h["fired_bits"]), round_
        assert np.array_equal(out_d["potentials"], out_h["potentials"]), round_
        assert rd_d.spikes == rd_h.spikes
        for i in rng.choice(n, size=200, replace=False):
            group, offset = "pop", int(i)
            if dev.neuron_index(group, offset) != i:
                continue  # (device order = group order for this single-group network; be safe)
            bias[i] = float(rng.integers(0, 3)) * 64.0
            dev.set_neuron_attribute(group, offset, "bias", bias[i])
        oracle.set_bias(bias)


def test_pybind_in_memory_trace_formats():
    """In-memory traces of the pybind11 module have the reference's shapes (src/pytrace.hpp): spike_trace = per step a
    list of NeuronAddress, potential_trace = per step a list of floats, neuron_trace = {name: per step values},
    perf_trace = {column: per step value}, message_trace = per step a list of message dicts; entries not asked for
    are None (src/pymodule.cpp:691-702)."""
    import os
    from helpers import ROOT, golden_flat
    from sanafe_b200 import sanafecpp_b200 as m
    cwd = os.getcwd()
    os.chdir(ROOT)
    try:
        arch, net = m.load_flat(golden_flat("noise"))
        chip = m.SpikingChip(arch)
        chip.load(net)
    finally:
        os.chdir(cwd)
    res = chip.sim(20, timing_model="simple", spike_trace=True, potential_trace=True, perf_trace=True, message_trace=True)
    assert res["neuron_trace"] is None
    spikes = res["spike_trace"]
    assert len(spikes) == 20 and all(isinstance(step, list) for step in spikes)
    first = next(a for step in spikes for a in step)
    assert isinstance(first, m.NeuronAddress) and isinstance(first.group_name, str) and first.neuron_offset >= 0
    assert len(res["potential_trace"]) == 20 and len(res["potential_trace"][0]) == 8
    assert set(res["perf_trace"]) >= {"timestep", "fired", "updated", "hops", "spikes", "sim_time", "total_energy"}
    msgs = res["message_trace"]
    assert len(msgs) == 20 and sum(len(step) for step in msgs) > 0
    one = next(msg for step in msgs for msg in step)
    assert {"timestep", "mid", "spikes", "hops", "generation_delay", "placeholder", "src_neuron_group_id"} <= set(one)
    none = chip.sim(1, timing_model="simple")
    assert all(none[k] is None for k in ("spike_trace", "potential_trace", "neuron_trace", "perf_trace", "message_trace"))


def test_reference_unit_vectors_on_the_device(tmp_path):
    """The reference's per-model unit-test expectations (tests/unit/test_*.cpp), restated on a one-neuron rig, on
    the device (the CPU restatement passes the same list in test_reference_unit_vectors.py)."""
    import reference_unit_vectors as vectors
    import unit_rig as rig
    vectors.check_all(tmp_path, 0, rig.device_runner)


@pytest.mark.timeout(180)
def test_ctrl_c_interrupts_a_long_sim():
    """The reference's Python loop polls for signals every 100 ms (src/pymodule.cpp:629-652): a KeyboardInterrupt
    must surface while the device runs, and the chip must stay usable afterwards."""
    import os
    import signal
    import threading
    import time
    from helpers import ROOT, golden_flat
    from sanafe_b200 import sanafecpp_b200 as m
    cwd = os.getcwd()
    os.chdir(ROOT)
    try:
        arch, net = m.load_flat(golden_flat("example"))
        chip = m.SpikingChip(arch)
        chip.load(net)
    finally:
        os.chdir(cwd)
    chip.sim(100, timing_model="simple")  # warm-up
    threading.Timer(0.5, lambda: os.kill(os.getpid(), signal.SIGINT)).start()
    t0 = time.time()
    with pytest.raises(KeyboardInterrupt):
        chip.sim(3_000_000, timing_model="simple")
    assert time.time() - t0 < 60.0
    res = chip.sim(10, timing_model="simple")
    assert res["timesteps_executed"] == 10 and res["timestep_start"] > 101


@pytest.mark.parametrize("device_draws", ["1", "0"])
def test_poisson_draws_device_and_host(monkeypatch, device_draws):
    """Poisson inputs against the reference's golden with the draws made on the device (poisson_kernel, the default)
    and on the host with libstdc++'s generator (SFE_DEVICE_POISSON=0, the cross-check): the same spikes either way."""
    monkeypatch.setenv("SFE_DEVICE_POISSON", device_draws)
    chip = load_chip("poisson", device=0)
    g = golden("poisson")
    rd, out = chip.sim_raw(g["steps"], "simple", steps=True, fired=True, potentials=True)
    check_against_golden("poisson", chip, rd, out, potential_rtol=0.0, energy_rtol=1e-9)


def test_oversized_core_accumulates_in_hbm(monkeypatch):
    """A core whose dendrite cells do not fit the shared-memory budget accumulates in HBM (acc_global). Forced here
    with a zero budget on goldens of every accumulation mode: ordered fp64 (frac), exact with a delay ring
    (synth_delay), the charge-loss quirk; cores with 4-byte records keep their shared-memory path (synth_small)."""
    monkeypatch.setenv("SFE_ACC_SMEM_LIMIT", "0")
    for name in ("frac", "synth_delay", "synth_quirk", "synth_small", "dvs"):
        chip = load_chip(name, device=0)
        g = golden(name)
        rd, out = chip.sim_raw(g["steps"], "simple", steps=True, fired=True, potentials=True)
        check_against_golden(name, chip, rd, out, potential_rtol=0.0, energy_rtol=1e-9)
