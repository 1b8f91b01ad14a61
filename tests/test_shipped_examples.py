"""The example descriptions the reference ships besides BASELINE's configs (arch/demo_with_dendrites.yaml +
snn/dendrite.yaml: the `taps` dendrite demo; snn/input_net.yaml: two input neurons sharing one `input` unit;
snn/conv.yaml: a conv2d hyper-edge), run by the reference itself (oracle/_ref/sanafe_ref, fed through PyYAML) and
by this repo's YAML front-end + lowering + CPU restatement: same totals, same spike rows. Needs the reference tree
and its compiled engine, so it runs in the build container only."""
import os
import sys

import pytest

import sanafe_b200 as sfe
import numpy as np
from helpers import (REFERENCE_ROOT, ROOT, Oracle, have_reference_binary, load_ref_potentials, load_ref_spikes, rel_err,
                     run_reference)

pytestmark = pytest.mark.skipif(not (os.path.exists(REFERENCE_ROOT) and have_reference_binary()),
                                reason="needs the reference tree and oracle/_ref/sanafe_ref")

CASES = {
    "dendrite": ("arch/demo_with_dendrites.yaml", "snn/dendrite.yaml", 50),
    "input_net": ("arch/example_chip.yaml", "snn/input_net.yaml", 40),
    "conv": ("arch/example_chip.yaml", "snn/conv.yaml", 60),
}


@pytest.mark.parametrize("name", list(CASES))
def test_shipped_example_matches_reference(name, tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import yaml_to_flat
    arch_rel, net_rel, steps = CASES[name]
    arch_path, net_path = os.path.join(REFERENCE_ROOT, arch_rel), os.path.join(REFERENCE_ROOT, net_rel)
    flat = str(tmp_path / "case.jsonl")
    yaml_to_flat.convert(arch_path, net_path, flat)
    want = run_reference(flat, str(tmp_path), steps, "simple", per_step=True)
    want_spikes = load_ref_spikes(str(tmp_path))
    arch = sfe.load_arch(arch_path)
    chip = sfe.SpikingChip(arch, device=-1)
    chip.set_input_seed_base(0)
    chip.load(sfe.load_net(net_path, arch))
    rd, out = Oracle(chip).run(steps)
    for key in ("spikes", "packets_sent", "neurons_updated", "neurons_fired"):
        assert getattr(rd, key) == want[key], (name, key, getattr(rd, key), want[key])
    for key in ("total_energy", "synapse_energy", "dendrite_energy", "soma_energy", "network_energy", "sim_time"):
        assert rel_err(getattr(rd, key), want[key]) <= 1e-12, (name, key)
    assert chip.format_spikes(out["fired_bits"], 1) == want_spikes, name
    want_potentials = load_ref_potentials(str(tmp_path))
    if want_potentials is not None:
        assert np.array_equal(out["potentials"], want_potentials), (name, "potentials")
