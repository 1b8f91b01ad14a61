"""The pybind11 module (sanafecpp_b200) mirrors the reference's `sanafecpp` surface:
same function / method / keyword names and result-dict keys (src/pymodule.cpp:850-1213)."""
import hashlib
import inspect
import os

import pytest

from helpers import REFERENCE_ROOT, ROOT, golden, golden_flat, golden_spikes, rel_err


def module():
    from sanafe_b200 import sanafecpp_b200
    return sanafecpp_b200


def test_surface_matches_reference_names():
    m = module()
    for name in ("load_arch", "load_net", "SpikingChip", "Architecture", "Network"):
        assert hasattr(m, name), name
    doc = m.SpikingChip.sim.__doc__
    for kw in ("timesteps", "timing_model", "processing_threads", "scheduler_threads", "spike_trace",
               "potential_trace", "neuron_trace", "perf_trace", "message_trace", "write_trace_headers"):
        assert kw in doc, kw
    assert "overwrite" in m.SpikingChip.load.__doc__


def test_no_gpu_no_fallback():
    import sanafe_b200 as sfe
    if sfe.lib().sfe_device_count() > 0:
        pytest.skip("a GPU is present")
    m = module()
    arch, net = m.load_flat(golden_flat("example"))
    with pytest.raises(RuntimeError, match="no CUDA device"):
        m.SpikingChip(arch)


@pytest.mark.gpu
def test_sim_result_dict_and_traces(tmp_path):
    m = module()
    os.chdir(ROOT)
    arch, net = m.load_flat(golden_flat("dvs"))
    chip = m.SpikingChip(arch)
    chip.load(net)
    spikes_path = str(tmp_path / "spikes.csv")
    res = chip.sim(1000, timing_model="detailed", spike_trace=spikes_path, perf_trace=True)
    g = golden("dvs")
    s = g["summary"]
    assert set(res) >= {"timestep_start", "timesteps_executed", "energy", "sim_time", "spikes", "packets_sent",
                        "neurons_updated", "neurons_fired", "perf_trace"}
    assert set(res["energy"]) == {"total", "synapse", "dendrite", "soma", "network"}
    assert res["spikes"] == s["spikes"] and res["neurons_fired"] == s["neurons_fired"]
    assert res["packets_sent"] == s["packets_sent"] and res["neurons_updated"] == s["neurons_updated"]
    assert rel_err(res["energy"]["total"], s["total_energy"]) <= 1e-9
    assert rel_err(res["sim_time"], g["detailed"]["sim_time"]) <= 1e-9
    text = open(spikes_path).read()
    assert text.startswith("neuron,timestep\n")
    # the reference's own spikes.csv of this run has this md5
    assert hashlib.md5(text.encode()).hexdigest() == "52d77e728bc457796781163eb622f88f"
    assert abs(chip.get_power() - res["energy"]["total"] / res["sim_time"]) <= 1e-12 * chip.get_power()
