"""Out-of-tree device models (include/sfe_device_model.h): the replacement of the reference's hardware-unit plugin
ABI (dlopen + create_<model>(), src/plugins.cpp:45-98; PipelineUnit, src/pipeline.hpp:69-301). The worked example is
the reference's Hodgkin-Huxley plugin built out of tree (sana-fe_b200/plugins/hodgkin_huxley_device.cu): registered
under the model's name, or named by a unit's `plugin:` path, it takes the place of the functor compiled into the
engine and must reproduce the reference's golden."""
import ctypes as C
import os

import numpy as np
import pytest

import sanafe_b200 as sfe
from helpers import ROOT, check_against_golden, golden, golden_flat

PLUGIN = os.path.join(ROOT, "sana-fe_b200", "sanafe_b200", "plugins", "libhodgkin_huxley_b200.so")


@pytest.fixture
def registered_hh():
    L = sfe.lib()
    assert L.sfe_load_device_model(b"hodgkin_huxley", PLUGIN.encode()) == 0, L.sfe_last_error()
    yield
    L.sfe_unregister_device_model(b"hodgkin_huxley")


def load_hh(device, flat=None):
    cwd = os.getcwd()
    os.chdir(ROOT)
    try:
        arch, net = sfe.load_flat(flat or golden_flat("hh"))
        chip = sfe.SpikingChip(arch, device=device)
        chip.set_input_seed_base(0)
        chip.load(net)
    finally:
        os.chdir(cwd)
    return chip


def test_registry_entry_points():
    L = sfe.lib()
    assert L.sfe_device_model_registered(b"hodgkin_huxley") == 0
    assert L.sfe_register_device_model(b"x", None) == -1 and b"null descriptor" in L.sfe_last_error()
    assert L.sfe_load_device_model(b"hodgkin_huxley", b"/nonexistent/lib.so") == -1
    assert b"could not load library" in L.sfe_last_error()
    # a library without the factory symbol (the engine's own shared library)
    own = os.path.join(ROOT, "sana-fe_b200", "sanafe_b200", "libsanafe_b200.so").encode()
    assert L.sfe_load_device_model(b"hodgkin_huxley", own) == -1
    assert b"does not export sfe_device_model_hodgkin_huxley" in L.sfe_last_error()
    assert L.sfe_load_device_model(b"hodgkin_huxley", PLUGIN.encode()) == 0, L.sfe_last_error()
    assert L.sfe_device_model_registered(b"hodgkin_huxley") == 1
    # registering the descriptor by hand, as an embedding application would
    plug = C.CDLL(PLUGIN)
    plug.sfe_device_model_hodgkin_huxley.restype = C.c_void_p
    assert L.sfe_register_device_model(b"hh_again", plug.sfe_device_model_hodgkin_huxley()) == 0
    assert L.sfe_unregister_device_model(b"hh_again") == 0
    assert L.sfe_unregister_device_model(b"hh_again") == -1
    assert L.sfe_unregister_device_model(b"hodgkin_huxley") == 0
    assert L.sfe_device_model_registered(b"hodgkin_huxley") == 0


def test_lowering_uses_the_registered_model(registered_hh):
    """With the model registered, the Hodgkin-Huxley units of the golden description lower to device-model
    instances (one per neuron) whose state words / parameters carry the neurons' attributes."""
    chip = load_hh(device=-1)
    tb = chip.tables
    n = tb.n_neurons
    classes = [tb.soma_classes[tb.neuron_class[i]].model for i in range(n)]
    hh_neurons = [i for i in range(n) if classes[i] == 5]  # SFE_SOMA_DEVICE_MODEL
    assert hh_neurons and tb.n_hh == 0
    assert tb.n_device_models == 1 and tb.n_device_instances == len(hh_neurons)
    assert sorted(tb.neuron_aux[i] for i in hh_neurons) == list(range(len(hh_neurons)))
    # without the registration the same description uses the functor compiled into the engine
    sfe.lib().sfe_unregister_device_model(b"hodgkin_huxley")
    chip2 = load_hh(device=-1)
    assert chip2.tables.n_device_models == 0 and chip2.tables.n_hh == len(hh_neurons)
    sfe.lib().sfe_load_device_model(b"hodgkin_huxley", PLUGIN.encode())


def flat_with_plugin_path(tmp_path):
    text = open(golden_flat("hh")).read().replace("oracle/_ref/libhodgkin_huxley.so", PLUGIN)
    assert PLUGIN in text
    path = tmp_path / "hh_out_of_tree.jsonl"
    path.write_text(text)
    return str(path)


def test_plugin_path_of_the_description_is_loaded(tmp_path):
    """`plugin: <library>` on a unit: the library is dlopen'ed at load() and asked for sfe_device_model_<model>."""
    L = sfe.lib()
    assert L.sfe_device_model_registered(b"hodgkin_huxley") == 0
    chip = load_hh(device=-1, flat=flat_with_plugin_path(tmp_path))
    assert chip.tables.n_device_models == 1
    assert L.sfe_device_model_registered(b"hodgkin_huxley") == 1
    L.sfe_unregister_device_model(b"hodgkin_huxley")


def test_host_only_plugin_is_refused(tmp_path):
    text = open(golden_flat("hh")).read().replace('"model":"hodgkin_huxley"', '"model":"some_host_model"')
    path = tmp_path / "host_plugin.jsonl"
    path.write_text(text)
    if '"some_host_model"' not in text:
        pytest.skip("flat format changed")
    with pytest.raises(sfe.SanafeError, match="has no device model"):
        load_hh(device=-1, flat=str(path))


@pytest.mark.gpu
@pytest.mark.parametrize("route", ["registered", "plugin_path"])
def test_out_of_tree_hodgkin_huxley_matches_reference(route, tmp_path):
    """The reference's golden of snn/hh_example with the soma model running as an out-of-tree device kernel."""
    L = sfe.lib()
    if route == "registered":
        assert L.sfe_load_device_model(b"hodgkin_huxley", PLUGIN.encode()) == 0, L.sfe_last_error()
        chip = load_hh(device=0)
    else:
        chip = load_hh(device=0, flat=flat_with_plugin_path(tmp_path))
    try:
        assert chip.tables.n_device_models == 1
        g = golden("hh")
        rd, out = chip.sim_raw(g["steps"], "simple", steps=True, fired=True, potentials=True)
        check_against_golden("hh", chip, rd, out, potential_rtol=1e-9, energy_rtol=1e-9)
        # reset(): the model's reset mask zeroes V, m, n, h (plugins/hodgkin_huxley.cpp:70-86)
        chip.reset()
        rd2, out2 = chip.sim_raw(5, "simple", steps=True, fired=True, potentials=True)
        assert np.all(np.isfinite(out2["potentials"]))
    finally:
        L.sfe_unregister_device_model(b"hodgkin_huxley")


def test_pybind_module_registers_device_models():
    """The drop-in Python module exposes the programmatic side of the plugin loader."""
    from sanafe_b200 import sanafecpp_b200 as m
    assert not m.device_model_registered("hodgkin_huxley")
    m.load_device_model("hodgkin_huxley", PLUGIN)
    assert m.device_model_registered("hodgkin_huxley")
    m.unregister_device_model("hodgkin_huxley")
    assert not m.device_model_registered("hodgkin_huxley")
    with pytest.raises(RuntimeError, match="could not load library"):
        m.load_device_model("hodgkin_huxley", "/nonexistent/lib.so")
