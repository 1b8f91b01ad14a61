"""The in-tree C++ YAML reader (csrc/host/yaml*.cpp) against the independent PyYAML path
the oracle uses (oracle/yaml_to_flat.py): both must lower to identical tables / results.
Needs the reference's arch/ and snn/ files, so it runs in this container only."""
import os

import numpy as np
import pytest

import sanafe_b200 as sfe
from helpers import GOLDEN, REFERENCE_ROOT, ROOT, Oracle, golden, golden_flat

pytestmark = pytest.mark.skipif(not os.path.exists(REFERENCE_ROOT), reason="needs the reference tree")

SRC = os.path.join(GOLDEN, "src")
CASES = {
    "example": (f"{REFERENCE_ROOT}/arch/example_chip.yaml", f"{REFERENCE_ROOT}/snn/example_snn.yaml"),
    "dvs": (f"{REFERENCE_ROOT}/arch/loihi.yaml", f"{REFERENCE_ROOT}/snn/dvs.yaml"),
    "hh": (f"{SRC}/hh_arch.yaml", f"{SRC}/hh_snn.yaml"),
    "truenorth": (f"{REFERENCE_ROOT}/arch/truenorth.yaml", f"{SRC}/tn_snn.yaml"),
    "frac": (f"{REFERENCE_ROOT}/arch/example_chip.yaml", f"{SRC}/frac_snn.yaml"),
}


def arrays(t):
    def arr(ptr, n):
        return np.ctypeslib.as_array(ptr, shape=(n,)).copy() if n else np.zeros(0)
    return {
        "neuron_class": arr(t.neuron_class, t.n_neurons), "neuron_bias": arr(t.neuron_bias, t.n_neurons),
        "axon_out_begin": arr(t.axon_out_begin, t.n_neurons + 1),
        "syn_weight": arr(t.syn_weight, t.n_synapses), "syn_meta": arr(t.syn_meta, t.n_synapses),
        "probes": arr(t.probes, t.n_probes),
    }


@pytest.mark.parametrize("name", list(CASES))
def test_yaml_reader_equals_pyyaml_path(name):
    arch_path, net_path = CASES[name]
    cwd = os.getcwd()
    os.chdir(ROOT)
    try:
        arch = sfe.load_arch(arch_path)
        net = sfe.load_net(net_path, arch)
        chip_y = sfe.SpikingChip(arch, device=-1)
        chip_y.load(net)
        arch_f, net_f = sfe.load_flat(golden_flat(name))
        chip_f = sfe.SpikingChip(arch_f, device=-1)
        chip_f.load(net_f)
    finally:
        os.chdir(cwd)
    ty, tf = chip_y.tables, chip_f.tables
    for field in ("n_neurons", "n_synapses", "n_axons_in", "n_soma_classes", "mapped_cores", "mapped_tiles", "sync_delay"):
        assert getattr(ty, field) == getattr(tf, field), field
    ay, af = arrays(ty), arrays(tf)
    for key in ay:
        assert np.array_equal(ay[key], af[key]), key
    steps = min(golden(name)["steps"], 100)
    rd_y, out_y = Oracle(chip_y).run(steps)
    rd_f, out_f = Oracle(chip_f).run(steps)
    assert np.array_equal(out_y["fired_bits"], out_f["fired_bits"])
    for key in out_y["steps"].dtype.names:
        assert np.array_equal(out_y["steps"][key], out_f["steps"][key]), key


def test_yaml_reader_loads_every_shipped_architecture():
    for fname in sorted(os.listdir(f"{REFERENCE_ROOT}/arch")):
        if fname.endswith(".yaml"):
            assert sfe.load_arch(f"{REFERENCE_ROOT}/arch/{fname}") is not None, fname


def test_yaml_errors():
    with pytest.raises(sfe.SanafeError, match="Failed to open architecture file"):
        sfe.load_arch("/nonexistent/arch.yaml")
    arch = sfe.load_arch(f"{REFERENCE_ROOT}/arch/example_chip.yaml")
    with pytest.raises(sfe.SanafeError, match="failed to open"):
        sfe.load_net("/nonexistent/net.yaml", arch)
