"""BASELINE config 5 (design-space sweep), CPU side: the generated TrueNorth-shaped chip equals the
reference's arch/truenorth.yaml when its costs are zeroed, the conv SNN lowers to the expected shape, the batch
loader (worker threads) produces the tables of a sequential load, and cost multipliers scale energies while leaving
the spikes alone (checked on the CPU restatement)."""
import ctypes as C
import os

import numpy as np
import pytest

import sanafe_b200 as sfe
from helpers import REFERENCE_ROOT, Oracle
from sanafe_b200 import dse


def _records(ptr, count):
    """Field values of a ctypes struct array (padding bytes are not part of the comparison)."""
    def value(x):
        return tuple(value(getattr(x, name)) for name, _ in x._fields_) if hasattr(x, "_fields_") else x
    return tuple(value(ptr[k]) for k in range(count))


def _bytes(ptr, n_bytes):
    return C.string_at(ptr, n_bytes) if n_bytes else b""  # an empty table may be a NULL pointer


def table_bytes(t):
    """The lowered tables that decide the simulation, in a comparable form."""
    return (_records(t.cores, t.n_cores), _records(t.soma_classes, t.n_soma_classes),
            _records(t.cost_classes, t.n_cost_classes), _records(t.axons_in, t.n_axons_in),
            _records(t.inputs, t.n_inputs), _records(t.noise, t.n_noise),
            _bytes(t.neuron_class, 4 * t.n_neurons), _bytes(t.neuron_aux, 4 * t.n_neurons),
            _bytes(t.neuron_bias, 8 * t.n_neurons), _bytes(t.neuron_potential0, 8 * t.n_neurons),
            _bytes(t.axon_out_begin, 4 * (t.n_neurons + 1)), _bytes(t.axon_out_target, 4 * t.n_axons_out),
            _bytes(t.input_spikes, t.n_input_spikes), _bytes(t.probes, 4 * t.n_probes),
            _bytes(t.u_probes, 4 * t.n_u_probes), _bytes(t.noise_values, 8 * t.n_noise_values),
            _bytes(t.syn_weight, 8 * t.n_synapses), _bytes(t.syn_meta, 4 * t.n_synapses))


def test_sweep_points():
    pts = dse.sweep_points()
    assert len(pts) == 1024 and len(set(pts)) == 1024
    assert {p[0] for p in pts} == {8 * k for k in range(1, 33)}
    assert max(dse.cores_needed(p[0]) for p in pts) <= 4096


@pytest.mark.skipif(not os.path.exists(os.path.join(REFERENCE_ROOT, "arch", "truenorth.yaml")),
                    reason="reference tree not present")
def test_generated_chip_has_the_shape_of_truenorth_yaml(tmp_path):
    """Same network on arch/truenorth.yaml and on the generated chip with every cost multiplied by 0:
    identical lowered tables (the shipped file has all costs 0)."""
    text, cores = dse.snn_yaml(192)
    (tmp_path / "snn.yaml").write_text(text)
    (tmp_path / "arch.yaml").write_text(dse.arch_yaml(0.0, tiles=4096))
    tabs = []
    for arch_path in (os.path.join(REFERENCE_ROOT, "arch", "truenorth.yaml"), str(tmp_path / "arch.yaml")):
        arch = sfe.load_arch(arch_path)
        chip = sfe.SpikingChip(arch, device=-1)
        chip.load(sfe.load_net(str(tmp_path / "snn.yaml"), arch))
        tabs.append((chip, table_bytes(chip.tables)))
    assert tabs[0][1] == tabs[1][1]
    t = tabs[0][0].tables
    assert t.n_neurons == sum(dse.LAYERS.values()) == 10043
    assert t.n_synapses == 3600 * 9 + 5408 * 144 + 5408 * 11
    assert t.mapped_cores == cores == dse.cores_needed(192)


def test_batch_load_equals_sequential_load_and_costs_only_scale_energy(tmp_path):
    points = [(64, 1.0), (64, 2.0), (200, 1.0), (256, 0.5)]
    sweep = dse.Sweep(points, str(tmp_path), device=-1, host_threads=3)
    # sequential load of the same design points
    for (npc, mult), chip in zip(points, sweep.chips):
        (tmp_path / "a.yaml").write_text(dse.arch_yaml(mult, tiles=dse.cores_needed(npc)))
        (tmp_path / "n.yaml").write_text(dse.snn_yaml(npc)[0])
        arch = sfe.load_arch(str(tmp_path / "a.yaml"))
        lone = sfe.SpikingChip(arch, device=-1)
        lone.load(sfe.load_net(str(tmp_path / "n.yaml"), arch))
        assert table_bytes(lone.tables) == table_bytes(chip.tables), (npc, mult)
    steps = 24
    runs = [Oracle(chip).run(steps) for chip in sweep.chips]
    (rd1, o1), (rd2, o2), (rd3, o3), _ = runs
    # a cost multiplier changes no spike; energy and simulated time scale with it (powers of two: exactly)
    assert np.array_equal(o1["fired_bits"], o2["fired_bits"]) and rd1.spikes == rd2.spikes > 0
    assert rd2.total_energy == 2.0 * rd1.total_energy and rd2.sim_time == 2.0 * rd1.sim_time
    # a different mapping changes messages and time, not the network's behaviour
    assert rd3.neurons_fired == rd1.neurons_fired and rd3.spikes == rd1.spikes
    assert rd3.packets_sent != rd1.packets_sent


def test_batch_sim_needs_a_device(tmp_path):
    sweep = dse.Sweep([(128, 1.0), (128, 2.0)], str(tmp_path), device=-1)
    with pytest.raises(sfe.SanafeError, match="sfe_batch_sim: chip 0: .*no CUDA device"):
        sweep.sim(3)


def test_sweep_is_dealt_round_robin_to_ranks():
    """More GPUs = replicas only: every design point lands on exactly one rank, 128 per rank at 8 GPUs."""
    import sys
    from helpers import ROOT
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import dse_sweep
    pts = dse.sweep_points()
    parts = [dse_sweep.points_of_rank(pts, r, 8) for r in range(8)]
    assert sorted(sum(parts, [])) == sorted(pts) and all(len(p) == 128 for p in parts)


def test_bench_deals_128_design_points_to_every_gpu():
    """bench.py at N GPUs: 128 design points per rank, disjoint, all 1024 at N = 8, every mapping on every rank."""
    import sys
    from helpers import ROOT
    sys.path.insert(0, ROOT)
    import bench
    every = set(dse.sweep_points())
    for world in (1, 2, 4, 8):
        parts = [bench.dse_points_of_rank(r, world) for r in range(world)]
        assert all(len(p) == 128 for p in parts)
        flat = [pt for p in parts for pt in p]
        assert len(set(flat)) == 128 * world and set(flat) <= every
        for p in parts:
            assert {pt[0] for pt in p} == {8 * k for k in range(1, 33)}
    assert set(pt for r in range(8) for pt in bench.dse_points_of_rank(r, 8)) == every
