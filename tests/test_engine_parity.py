"""Parity tests proper: the CUDA engine, called through the C ABI, against
(a) golden outputs of the reference itself and (b) the CPU restatement on larger
seeded synthetic networks. Bit-exact rasters / counters / potentials; energy and
latency to 1e-9 relative (north_star allows 1e-6); Hodgkin-Huxley potentials to
1e-9 relative (device exp/pow vs glibc)."""
import ctypes as C

import numpy as np
import pytest

import sanafe_b200 as sfe
from helpers import GOLDEN_CASES, Oracle, check_against_golden, golden, load_chip, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_engine_matches_reference(name):
    chip = load_chip(name, device=0)
    g = golden(name)
    rd, out = chip.sim_raw(g["steps"], "simple", steps=True, fired=True, potentials=True)
    check_against_golden(name, chip, rd, out, potential_rtol=1e-9 if name == "hh" else 0.0, energy_rtol=1e-9)
    assert abs(chip.get_power() - g["summary"]["power"]) <= 1e-9 * abs(g["summary"]["power"])


def test_forced_ordered_mode_matches_reference(monkeypatch):
    """SFE_FORCE_ORDERED=1 (the measurement knob of bench.py): every core takes the ordered fp64 path of the message
    phase, whatever its certificate allows; results are those of the reference all the same."""
    monkeypatch.setenv("SFE_FORCE_ORDERED", "1")
    for name in ("synth_small", "synth_delay"):
        chip = load_chip(name, device=0)
        g = golden(name)
        rd, out = chip.sim_raw(g["steps"], "simple", steps=True, fired=True, potentials=True)
        check_against_golden(name, chip, rd, out, potential_rtol=0.0, energy_rtol=1e-9)


def test_benchmark_shape_matches_the_reference(tmp_path):
    """The benchmark's own shape (32 of the 1024 cores x 1024 LIF neurons, fan-out 8 x 125, the
    generator and parameters of bench.py) against the reference's engine: hash over the rasters of 40 timesteps and the
    counters bit-exact, energy and simulated time to 1e-9. The golden was made by the reference itself
    (tests/golden/make_goldens.py --bench-sample); bench.py repeats the comparison live on the GPU box."""
    import json
    import os
    import sys
    from helpers import GOLDEN, ROOT
    sys.path.insert(0, ROOT)
    import bench
    from sanafe_b200 import archgen
    with open(os.path.join(GOLDEN, "bench_sample32.hash.json")) as f:
        g = json.load(f)
    h, cores = g["hash"], g["sample_cores"]
    assert g["spec"] == bench.sample_spec(cores), "bench.py's workload changed: regenerate the golden"
    flat = str(tmp_path / "arch.jsonl")
    archgen.write_flat(archgen.loihi_large(tiles=(cores + 3) // 4), flat)
    arch, _ = sfe.load_flat(flat)
    chip = sfe.SpikingChip(arch, device=0)
    chip.load_synthetic(sfe.SynthSpec(**bench.sample_spec(cores)), generate_on_device=True)
    rd, out = chip.sim_raw(int(h["hash_steps"]), "simple", steps=True, fired=True)
    assert f"{bench.raster_hash(out['fired_bits']):016x}" == h["raster_hash"]
    assert (rd.spikes, rd.packets_sent, rd.neurons_updated, rd.neurons_fired) == (
        h["hash_spikes"], h["hash_packets"], h["hash_updated"], h["hash_fired"])
    assert rel_err(rd.total_energy, h["hash_energy"]) <= 1e-9 and rel_err(rd.sim_time, h["hash_sim_time"]) <= 1e-9


def test_sim_calls_continue_state():
    """sim() twice == sim() once (state and the timestep counter persist, src/chip.cpp:481,553)."""
    a = load_chip("synth_delay", device=0)
    b = load_chip("synth_delay", device=0)
    rd_a, out_a = a.sim_raw(40, steps=True, fired=True, potentials=True)
    rd_b1, out_b1 = b.sim_raw(15, steps=True, fired=True, potentials=True)
    rd_b2, out_b2 = b.sim_raw(25, steps=True, fired=True, potentials=True)
    assert rd_b2.timestep_start == 16
    assert np.array_equal(out_a["fired_bits"], np.concatenate([out_b1["fired_bits"], out_b2["fired_bits"]]))
    assert np.array_equal(out_a["potentials"], np.concatenate([out_b1["potentials"], out_b2["potentials"]]))
    assert rd_a.spikes == rd_b1.spikes + rd_b2.spikes


def synth_spec(**kw):
    base = dict(cores=32, neurons_per_core=256, dest_cores=6, syn_per_axon=48, seed=11, bias_permille=100,
                bias=128.0, threshold=64.0, reset=0.0, leak_decay=0.9, w_min=-8, w_max=8, max_delay=0,
                log_spikes=1, log_potential_n=512)
    base.update(kw)
    return sfe.SynthSpec(**base)


def loihi_large_flat(tmp_path, tiles):
    """loihi_large-shaped architecture from the generated-architecture helper."""
    from sanafe_b200 import archgen
    path = str(tmp_path / "arch.jsonl")
    archgen.write_flat(archgen.loihi_large(tiles=tiles), path)
    return path


@pytest.mark.parametrize("kw", [dict(), dict(max_delay=3, seed=5), dict(neurons_per_core=1024, syn_per_axon=125, dest_cores=8, cores=16)])
def test_device_generated_synthetic_matches_restatement(tmp_path, kw):
    """Bulk path: synapses generated on the device from the spec == the same network
    materialised on the host and run through the CPU restatement."""
    spec = synth_spec(**kw)
    arch, _ = sfe.load_flat(loihi_large_flat(tmp_path, tiles=(spec.cores + 3) // 4))
    dev = sfe.SpikingChip(arch, device=0)
    dev.load_synthetic(spec, generate_on_device=True)
    host = sfe.SpikingChip(arch, device=-1)
    host.load_synthetic(spec, generate_on_device=False)
    steps = 25
    rd_d, out_d = dev.sim_raw(steps, steps=True, fired=True, potentials=True)
    rd_h, out_h = Oracle(host).run(steps)
    assert np.array_equal(out_d["fired_bits"], out_h["fired_bits"])
    assert np.array_equal(out_d["potentials"], out_h["potentials"])
    for key in ("neurons_fired", "neurons_updated", "packets_sent", "total_hops", "spike_count"):
        assert np.array_equal(out_d["steps"][key], out_h["steps"][key]), key
    for key in ("sim_time", "total_energy", "synapse_energy", "soma_energy", "network_energy"):
        assert rel_err(out_d["steps"][key], out_h["steps"][key]) <= 1e-9, key
    assert rd_d.spikes == rd_h.spikes and rd_d.spikes > 0


def test_bias_patch_and_reset():
    """MappedNeuron.set_attributes(bias) between sim() calls + reset() (scripts/tcad2025/dvs_gesture.py loop)."""
    chip = load_chip("synth_soma", device=0)
    ref = load_chip("synth_soma", device=-1)
    oracle = Oracle(ref)
    n = chip.tables.n_neurons
    bias = np.array([chip.tables.neuron_bias[i] for i in range(n)])
    for round_ in range(3):
        rd_d, out_d = chip.sim_raw(10, steps=True, fired=True, potentials=True)
        rd_h, out_h = oracle.run(10)
        assert np.array_equal(out_d["fired_bits"], out_h["fired_bits"]), round_
        assert np.array_equal(out_d["potentials"], out_h["potentials"]), round_
        bias = np.roll(bias, 17)
        assert sfe.lib().sfe_engine_set_bias(chip.engine, bias.ctypes.data, n) == 0
        oracle.set_bias(bias)
        if round_ == 1:
            chip.reset()
            oracle.reset()


@pytest.mark.parametrize("name,world", [("synth_delay", 2), ("synth_soma", 3), ("dvs", 4), ("frac", 2)])
def test_partitioned_engines_match_reference(name, world):
    """Multi-GPU path emulated on one device: `world` engines each simulate a contiguous core
    range; fired-raster slices are exchanged between the neuron and message phases. The merged
    result must equal the reference's golden (rasters/counters bit-exact)."""
    import hashlib
    from helpers import golden_spikes

    def make(device, rank, w):
        import os
        from helpers import ROOT, golden_description
        cwd = os.getcwd()
        os.chdir(ROOT)
        try:
            arch, net = golden_description(name)
            chip = sfe.SpikingChip(arch, device=device)
            chip.set_partition(rank, w)
            chip.load(net)
        finally:
            os.chdir(cwd)
        return chip

    g = golden(name)
    steps = min(g["steps"], 150)
    part = sfe.PartitionedChip(make, world)
    rasters = []
    for _ in range(steps):
        part.step()
        rasters.append(part.raster())
    rec = part.collect()
    ps = g["per_step"]
    for key, col in (("fired", "neurons_fired"), ("updated", "neurons_updated"), ("packets", "packets_sent"),
                     ("spikes", "spike_count")):
        assert np.array_equal(rec[col], np.asarray(ps[key][:steps], dtype=np.int64)), (name, key)
    for key in ("sim_time", "synapse_energy", "dendrite_energy", "soma_energy", "network_energy", "total_energy"):
        assert rel_err(rec[key], ps[key][:steps]) <= 1e-9, (name, key)
    # raster rows in reference trace order
    n = part.chips[0].tables.n_neurons
    words = (n + 31) // 32
    bits = np.zeros((steps, words * 4), dtype=np.uint8)
    for s, r in enumerate(rasters):
        bits[s, :len(r)] = r
    text = part.chips[0].format_spikes(bits.view(np.uint32), 1)
    want = "".join(line + "\n" for line in golden_spikes(name).split("\n") if line and int(line.rsplit(",", 1)[1]) <= steps)
    assert text == want


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["synth_small", "synth_delay", "synth_soma", "synth_quirk"])
def test_fused_step_kernel_equals_two_kernel_step(name, monkeypatch):
    """With SFE_FUSED_STEP=1 sfe_engine_enqueue(K >= 2) runs the fused step kernel (one launch per step: message phase of step t, fold of
    step t, neuron phase of step t + 1 core by core, ready flag) where sfe_chip_sim runs the two-kernel step. Same
    per-step records, final potentials and last raster, also across two consecutive batches and a reset."""
    import ctypes as C
    import sanafe_b200 as sfe
    L = sfe.lib()
    g = golden(name)
    steps = g["steps"]
    ref = load_chip(name, device=0)
    rd_ref, out_ref = ref.sim_raw(steps, "simple", steps=True, fired=True)
    monkeypatch.setenv("SFE_FUSED_STEP", "1")
    chip = load_chip(name, device=0)
    eng = chip.engine
    n = chip.tables.n_neurons
    first = steps // 3
    launches0 = L.sfe_engine_launch_count(eng)
    recs = []
    for count in (first, steps - first):
        assert L.sfe_engine_enqueue(eng, count) == 0, L.sfe_last_error()
        buf = np.zeros(count, dtype=sfe.STEP_DTYPE)
        got = L.sfe_engine_collect_records(eng, buf.ctypes.data, count)
        assert got == count, L.sfe_last_error()
        recs.append(buf)
    assert L.sfe_engine_launch_count(eng) - launches0 == steps + 4, "the fused step kernel did not run (steps + 2 launches per batch)"
    assert L.sfe_engine_exchange_error(eng) == 0
    got = np.concatenate(recs)
    for key in ("neurons_fired", "neurons_updated", "packets_sent", "total_hops", "spike_count"):
        assert np.array_equal(got[key], out_ref["steps"][key]), (name, key)
    for key in ("sim_time", "synapse_energy", "dendrite_energy", "soma_energy", "network_energy", "total_energy"):
        assert rel_err(got[key], out_ref["steps"][key]) <= 1e-12, (name, key)
    pa, pb = np.zeros(n), np.zeros(n)
    assert L.sfe_engine_read_potentials(eng, pa.ctypes.data, n) == 0
    assert L.sfe_engine_read_potentials(ref.engine, pb.ctypes.data, n) == 0
    assert np.array_equal(pa, pb), name
    words = (n + 31) // 32
    fa = np.zeros(words, dtype=np.uint32)
    assert L.sfe_engine_read_fired(eng, fa.ctypes.data, words) == 0
    assert np.array_equal(fa, out_ref["fired_bits"][-1]), name
    # and the golden itself through the fused path after a reset of both engines: same evolution again
    chip.reset()
    ref.reset()
    assert L.sfe_engine_enqueue(eng, 10) == 0, L.sfe_last_error()
    buf = np.zeros(10, dtype=sfe.STEP_DTYPE)
    assert L.sfe_engine_collect_records(eng, buf.ctypes.data, 10) == 10
    _, out2 = ref.sim_raw(10, "simple", steps=True)
    assert np.array_equal(buf["spike_count"], out2["steps"]["spike_count"]) and np.array_equal(buf["neurons_fired"], out2["steps"]["neurons_fired"])
