"""The object-by-object builder surface of the pybind11 module (reference: src/pymodule.cpp:850-1212): Network /
NeuronGroup / Neuron / Connection / Architecture / Tile / Core, MappedNeuron.set_attributes, Network.save. Runs
without a GPU: networks are built in Python, lowered on a host-only chip and checked (a) against the tables the
YAML front-end produces for the same network and (b), through the CPU restatement, against the reference's own
outputs (tests/golden/example.*)."""
import ctypes as C
import os

import numpy as np
import pytest

import sanafe_b200 as sfe
from helpers import GOLDEN, REFERENCE_ROOT, ROOT, Oracle, check_against_golden, golden
from test_dse_batch import table_bytes


def module():
    from sanafe_b200 import sanafecpp_b200
    return sanafecpp_b200


class HostChip:
    """A pybind11 SpikingChip (host-only) seen through the ctypes view of the C ABI (tables, format_spikes)."""

    def __init__(self, arch, net):
        self.chip = module().SpikingChip(arch, device=-1)
        sfe.lib().sfe_chip_set_input_seed_base(self.chip._handle, 0)
        self.chip.load(net)

    @property
    def tables(self):
        return sfe.lib().sfe_chip_tables(self.chip._handle).contents

    def format_spikes(self, fired_bits, timestep_start):
        view = sfe.SpikingChip.__new__(sfe.SpikingChip)
        view._h = self.chip._handle
        try:
            return sfe.SpikingChip.format_spikes(view, fired_bits, timestep_start)
        finally:
            view._h = None  # the pybind object owns the chip


def example_arch_path():
    """arch/example_chip.yaml restated for the box without the reference tree."""
    return os.path.join(GOLDEN, "src", "example_arch.yaml")


def build_example(m, arch):
    """snn/example_snn.yaml, object by object (what tutorial-style scripts do)."""
    net = m.Network()
    inp = net.create_neuron_group("in", 2, log_spikes=True)
    out = net.create_neuron_group("out", 2)
    inp[0].set_attributes(log_spikes=False)
    inp[1].set_attributes(model_attributes={"spikes": [1, 0, 1]})
    for n in out:
        # (the reference binds soma_attributes= to the dendrite unit and vice versa: Appendix B-11)
        n.set_attributes(dendrite_attributes={"threshold": 2}, model_attributes={"log_u": True}, log_potential=True)
    out[1].connect_to_neuron(out[1], {"weight": -4})
    inp.connect_neurons_dense(out, {"weight": [-1, 2, 1, 3]})
    core00, core01 = arch.tiles[0].cores[0], arch.tiles[0].cores[1]
    inp[0].set_attributes(soma_hw_name="demo_input")
    inp[0].map_to_core(core00)
    inp[1].set_attributes(soma_hw_name="demo_input")
    inp[1].map_to_core(core01)
    for n in out.neurons:
        n.map_to_core(core00)
    return net


def test_builder_surface_names():
    m = module()
    for name in ("Network", "NeuronGroup", "Neuron", "Connection", "NeuronAddress", "Architecture", "Tile", "Core",
                 "MappedNeuron", "SpikingChip", "BufferPosition", "HardwareMappingError", "framework_attributes",
                 "model_attributes", "load_arch", "load_net"):
        assert hasattr(m, name), name
    for kw in ("group_name", "neuron_count", "model_attributes", "default_synapse_hw_name", "default_dendrite_hw_name",
               "log_potential", "log_spikes", "soma_hw_name"):
        assert kw in m.Network.create_neuron_group.__doc__, kw
    for kw in ("dest_group", "attributes", "input_width", "input_height", "input_channels", "kernel_width",
               "kernel_height", "kernel_count", "stride_width", "stride_height"):
        assert kw in m.NeuronGroup.connect_neurons_conv2d.__doc__, kw
    for kw in ("model_attributes", "soma_attributes", "dendrite_attributes", "log_spikes"):
        assert kw in m.MappedNeuron.set_attributes.__doc__, kw
    assert "leaky_integrate_and_fire" in m.model_attributes and "energy_spike_out" in m.framework_attributes


def test_example_network_built_in_python_reproduces_the_reference(tmp_path):
    m = module()
    arch = m.load_arch(example_arch_path())
    net = build_example(m, arch)
    assert repr(net["out"]).startswith("sanafe::NeuronGroup(name=out") and len(net.groups) == 2
    assert [c.post_neuron.neuron_offset for c in net["in"][1].edges_out] == [0, 1]
    assert net["out"][1].edges_out[0].synapse_attributes == {"weight": -4}
    built = HostChip(arch, net)
    # the same network through the YAML front-end
    loaded = HostChip(arch, m.load_net(os.path.join(GOLDEN, "src", "example_snn.yaml"), arch))
    assert table_bytes(built.tables) == table_bytes(loaded.tables)
    # ... and against the reference's own run of arch/example_chip.yaml + snn/example_snn.yaml
    g = golden("example")
    rd, out = Oracle(built).run(g["steps"])
    check_against_golden("example", built, rd, out, potential_rtol=0.0, energy_rtol=1e-12)


@pytest.mark.skipif(not os.path.exists(os.path.join(REFERENCE_ROOT, "snn", "example_snn.yaml")), reason="reference tree not present")
def test_restated_example_files_equal_the_reference_files():
    m = module()
    ref_arch = m.load_arch(os.path.join(REFERENCE_ROOT, "arch", "example_chip.yaml"))
    ref = HostChip(ref_arch, m.load_net(os.path.join(REFERENCE_ROOT, "snn", "example_snn.yaml"), ref_arch))
    arch = m.load_arch(example_arch_path())
    mine = HostChip(arch, m.load_net(os.path.join(GOLDEN, "src", "example_snn.yaml"), arch))
    assert table_bytes(ref.tables) == table_bytes(mine.tables)


def test_save_round_trip(tmp_path):
    """Network.save writes a file that loads back to the same lowered tables (built, DVS-like conv and noise nets)."""
    m = module()
    arch = m.load_arch(example_arch_path())
    net = build_example(m, arch)
    path = str(tmp_path / "saved.yaml")
    net.save(path)
    again = m.load_net(path, arch)
    first, second = HostChip(arch, net), HostChip(arch, again)  # (the tables live as long as their chip)
    assert table_bytes(first.tables) == table_bytes(second.tables)
    text = open(path).read()
    assert "network:" in text and "mappings:" in text and '"out.1 -> out.1"' in text
    # a loaded network with hyper-edges, unit-specific attributes, non-default units and float weights
    os.chdir(ROOT)
    src = os.path.join(GOLDEN, "src")
    for arch_file, net_file in (("noise_arch.yaml", "noise_snn.yaml"), ("example_arch.yaml", "frac_snn.yaml"),
                                ("example_arch.yaml", "poisson_snn.yaml")):
        a = m.load_arch(os.path.join(src, arch_file))
        n1 = m.load_net(os.path.join(src, net_file), a)
        n1.save(path)
        n2 = m.load_net(path, a)
        c1, c2 = HostChip(a, n1), HostChip(a, n2)
        t1, t2 = c1.tables, c2.tables
        assert table_bytes(t1) == table_bytes(t2), net_file
        assert t1.n_poisson_cols == t2.n_poisson_cols and t1.n_noise == t2.n_noise and t1.n_u_probes == t2.n_u_probes
    # the saved file read by an independent YAML parser (PyYAML, oracle/yaml_to_flat.py) and run by the REFERENCE
    # engine gives the reference's result for the original description
    from helpers import have_reference_binary, run_reference
    if have_reference_binary():
        import sys
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import yaml_to_flat
        a = m.load_arch(os.path.join(src, "noise_arch.yaml"))
        m.load_net(os.path.join(src, "noise_snn.yaml"), a).save(path)
        flat = str(tmp_path / "saved.jsonl")
        yaml_to_flat.convert(os.path.join(src, "noise_arch.yaml"), path, flat)
        summary = run_reference(flat, str(tmp_path), 150, "simple")
        want = golden("noise")["summary"]
        for key in ("spikes", "packets_sent", "neurons_fired", "neurons_updated", "total_energy", "sim_time"):
            assert summary[key] == want[key], key
    with pytest.raises(RuntimeError, match="not mapped"):
        unmapped = m.Network()
        unmapped.create_neuron_group("g", 1)
        unmapped.save(path)


def test_conv_sparse_and_errors(tmp_path):
    m = module()
    arch = m.load_arch(example_arch_path())
    net = m.Network()
    a = net.create_neuron_group("a", 16, model_attributes={"threshold": 1.5, "bias": 1})
    b = net.create_neuron_group("b", 8, soma_hw_name="demo_soma_alt")
    c = net.create_neuron_group("c", 3)
    a.connect_neurons_conv2d(b, {"weight": np.arange(1, 9, dtype=np.float32) / 4}, input_width=4, input_height=4,
                             input_channels=1, kernel_width=2, kernel_height=2, kernel_count=2, stride_width=2, stride_height=2)
    b.connect_neurons_sparse(c, {"weight": [1, -2, 3]}, [(0, 0), (7, 2), (3, 1)])
    for n in list(a) + list(b) + list(c):
        n.map_to_core(arch.tiles[1].cores[2])
    built = HostChip(arch, net)
    # the same through YAML
    (tmp_path / "n.yaml").write_text("""network:
  name: x
  groups:
  - name: a
    attributes: {threshold: 1.5, bias: 1}
    neurons:
    - {0..15: {}}
  - name: b
    attributes: {}
    neurons:
    - {0..7: {}}
  - name: c
    attributes: {}
    neurons:
    - {0..2: {}}
  edges:
  - a -> b: {type: conv2d, input_width: 4, input_height: 4, input_channels: 1, kernel_width: 2, kernel_height: 2, kernel_count: 2, stride_width: 2, stride_height: 2, weight: [0.25, 0.5, 0.75, 1.0, 1.25, 1.5, 1.75, 2.0]}
  - b.0 -> c.0: {weight: 1}
  - b.7 -> c.2: {weight: -2}
  - b.3 -> c.1: {weight: 3}
mappings:
- {a: {core: '1.2'}}
- {b: {core: '1.2', soma: demo_soma_alt}}
- {c: {core: '1.2'}}
""")
    loaded = HostChip(arch, m.load_net(str(tmp_path / "n.yaml"), arch))
    assert built.tables.n_synapses == loaded.tables.n_synapses == 4 * 4 * 2 + 3
    assert table_bytes(built.tables) == table_bytes(loaded.tables)
    # error behaviour of the reference's builders
    with pytest.raises(ValueError, match="Reserved neuron attribute"):
        a[0].set_attributes(model_attributes={"soma_hw_name": "x"})
    with pytest.raises(IndexError):
        net["nope"]
    with pytest.raises(IndexError):
        a[16]
    with pytest.raises(ValueError):
        a.connect_neurons_dense(b, {"weight": [1, 2, 3]})  # needs 16 x 8 values
    with pytest.raises(ValueError, match="1D list"):
        a.connect_neurons_dense(b, {"weight": 3})
    # float32 narrowing of Python floats
    a[0].set_attributes(model_attributes={"threshold": 0.1})
    a[0].map_to_core(arch.tiles[0].cores[0])
    chip = HostChip(arch, net)
    t = chip.tables
    thresholds = {t.soma_classes[k].threshold for k in range(t.n_soma_classes)}
    assert float(np.float32(0.1)) in thresholds and 0.1 not in thresholds


def test_mapped_neuron_set_attributes():
    m = module()
    arch = m.load_arch(example_arch_path())
    chip = m.SpikingChip(arch, device=-1)
    chip.load(build_example(m, arch))
    groups = chip.mapped_neuron_groups
    assert sorted(groups) == ["in", "out"] and [len(groups[g]) for g in ("in", "out")] == [2, 2]
    t = sfe.lib().sfe_chip_tables(chip._handle).contents
    idx = sfe.lib().sfe_chip_neuron_index(chip._handle, b"out", 1)
    groups["out"][1].set_attributes(model_attributes={"bias": 0.5})
    assert t.neuron_bias[idx] == 0.5
    groups["out"][1].set_attributes(dendrite_attributes={"threshold": 7})
    t = sfe.lib().sfe_chip_tables(chip._handle).contents
    assert t.soma_classes[t.neuron_class[idx]].threshold == 7.0
    groups["in"][0].set_attributes(log_spikes=True)
    with pytest.raises(ValueError, match="Reserved neuron attribute"):
        groups["out"][0].set_attributes(model_attributes={"log_spikes": True})
    with pytest.raises(RuntimeError, match="before SpikingChip.load"):
        groups["in"][1].set_attributes(model_attributes={"spikes": [1, 1]})


def test_sim_errors_cross_the_worker_thread():
    """sim() runs the device loop on a worker thread (signal polling on the caller's): errors raised there must
    arrive as the exception the caller sees — here the no-fallback rule on a host-only chip."""
    m = module()
    arch = m.load_arch(example_arch_path())
    chip = m.SpikingChip(arch, device=-1)
    chip.load(build_example(m, arch))
    with pytest.raises(RuntimeError, match="no CUDA device"):
        chip.sim(10, timing_model="simple")


def test_second_load_needs_overwrite():
    """SpikingChip.load(net, overwrite=False) on a loaded chip maps a second network next to the first in the
    reference (src/chip.cpp:129-138); this engine refuses instead of silently replacing."""
    m = module()
    arch = m.load_arch(example_arch_path())
    chip = m.SpikingChip(arch, device=-1)
    chip.load(build_example(m, arch))
    with pytest.raises(RuntimeError, match="overwrite=True"):
        chip.load(build_example(m, arch))
    chip.load(build_example(m, arch), overwrite=True)
    assert sorted(chip.mapped_neuron_groups) == ["in", "out"]


def test_import_sanafe_shim():
    """`import sanafe` (PYTHONPATH=sana-fe_b200) gives the reference's top-level names."""
    import sanafe
    for name in ("load_arch", "load_net", "Network", "SpikingChip", "Architecture", "HardwareMappingError"):
        assert hasattr(sanafe, name), name
    arch = sanafe.load_arch(example_arch_path())
    net = sanafe.Network()
    group = net.create_neuron_group("g", 2, model_attributes={"threshold": 1.0})
    group[0].connect_to_neuron(group[1], {"weight": 1})
    for n in group:
        n.map_to_core(arch.tiles[0].cores[0])
    chip = sanafe.SpikingChip(arch, device=-1)
    chip.load(net)
    assert list(chip.mapped_neuron_groups) == ["g"]


def test_neuron_address_objects():
    """NeuronAddress (what in-memory spike traces hold, src/pytrace.hpp:121-141) is picklable as in the reference."""
    import pickle
    m = module()
    net = m.Network()
    g = net.create_neuron_group("grp", 2)
    g[0].connect_to_neuron(g[1])
    address = g[0].edges_out[0].post_neuron
    assert (address.group_name, address.neuron_offset) == ("grp", 1)
    again = pickle.loads(pickle.dumps(address))
    assert (again.group_name, again.neuron_offset) == ("grp", 1)
