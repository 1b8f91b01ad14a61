"""A one-neuron test rig: restates the reference's per-model unit tests (tests/unit/test_*.cpp call
`model.update(address, current, time)` directly) at the only level this engine has — a chip running timesteps.

`currents[k]` is the synaptic current the reference test passes to its (k+1)-th update() call (None = no input).
The rig delivers it with one driver neuron per call that fires exactly once, through a synapse of that weight. A
spike sent in step t reaches the soma in step t+1 (buffer before the soma), so update call k corresponds to
timestep k+2; the target is also updated in step 1 without input (v = 0, idle unless forced/biased)."""
import os

import numpy as np

import sanafe_b200 as sfe
from helpers import GOLDEN, ROOT
from sanafe_b200 import dse

STATUS = {0: "idle", 1: "idle", 2: "updated", 3: "fired"}


def _attrs(d):
    return "{" + ", ".join(f"{k}: {v}" for k, v in d.items()) + "}"


def lif_network(attrs, currents, soma="demo_soma_default"):
    """Target on core 0.0 of the demo chip, drivers on the `demo_input` units of cores 0.1 .. 1.3 (an input unit
    holds ONE spike train)."""
    assert len(currents) <= 6
    lines = ["network:", "  name: unit_rig", "  groups:",
             f"  - {{name: target, attributes: {_attrs(dict(attrs, log_spikes='true', log_potential='true'))}, neurons: [{{0: {{}}}}]}}"]
    # (a far-from-threshold sink behind the target keeps at least one synapse on the chip in every case)
    lines.append("  - {name: sink, attributes: {threshold: 1000000.0}, neurons: [{0: {}}]}")
    edges = ["  - {target.0 -> sink.0: {weight: 0.125}}"]
    maps = [f"- {{target.0: {{core: '0.0', soma: {soma}}}}}", "- {sink.0: {core: '1.3', soma: demo_soma_default}}"]  # its own core: soma / dendrite units hold per-unit state
    for k, c in enumerate(currents):
        if c is None:
            continue
        train = [0] * k + [1]
        lines.append(f"  - {{name: drv{k}, attributes: {{}}, neurons: [{{0: {{spikes: {train}}}}}]}}")
        edges.append(f"  - {{drv{k}.0 -> target.0: {{weight: {c!r}}}}}")
        core = k + 1
        maps.append(f"- {{drv{k}.0: {{core: '{core // 4}.{core % 4}', soma: demo_input}}}}")
    return "\n".join(lines + (["  edges:"] + edges if edges else ["  edges: []"]) + ["mappings:"] + maps) + "\n"


def truenorth_network(attrs, currents):
    """Drivers are TrueNorth neurons that count up to their firing step and then drop out (bias 1, threshold k+1,
    hard reset far below zero)."""
    lines = ["network:", "  name: unit_rig", "  groups:",
             f"  - {{name: target, attributes: {_attrs(dict(attrs, log_spikes='true', log_potential='true'))}, neurons: [{{0: {{}}}}]}}"]
    lines.append("  - {name: sink, attributes: {threshold: 1000000.0}, neurons: [{0: {}}]}")
    edges = ["  - {target.0 -> sink.0: {weight: 0.125}}"]
    assert len(currents) <= 6
    maps = ["- {target.0: {core: '0.0'}}", "- {sink.0: {core: '7.0'}}"]
    for k, c in enumerate(currents):
        if c is None:
            continue
        lines.append(f"  - {{name: drv{k}, attributes: {{bias: 1, threshold: {k + 1}, reset: -1000000, reset_mode: hard}}, neurons: [{{0: {{}}}}]}}")
        edges.append(f"  - {{drv{k}.0 -> target.0: {{weight: {c!r}}}}}")
        maps.append(f"- {{drv{k}.0: {{core: '{k + 1}.0'}}}}")
    return "\n".join(lines + (["  edges:"] + edges if edges else ["  edges: []"]) + ["mappings:"] + maps) + "\n"


def load(tmp_path, arch_text_or_path, net_text, device):
    tmp_path = str(tmp_path)
    if "\n" in arch_text_or_path:
        arch_path = os.path.join(tmp_path, "arch.yaml")
        with open(arch_path, "w") as f:
            f.write(arch_text_or_path)
    else:
        arch_path = arch_text_or_path
    net_path = os.path.join(tmp_path, "net.yaml")
    with open(net_path, "w") as f:
        f.write(net_text)
    cwd = os.getcwd()
    os.chdir(ROOT)
    try:
        arch = sfe.load_arch(arch_path)
        chip = sfe.SpikingChip(arch, device=device)
        chip.set_input_seed_base(0)
        chip.load(sfe.load_net(net_path, arch))
    finally:
        os.chdir(cwd)
    return chip


def run(chip, steps, runner):
    """(status names, potentials) of target.0 per timestep; runner(chip, steps) -> (RunData, traces)."""
    _, out = runner(chip, steps)
    i = chip.neuron_index("target", 0)
    names = chip.probe_names()
    col = names.index("target.0")
    return [STATUS[int(s)] for s in out["status"][:, i]], [float(v) for v in out["potentials"][:, col]]


def oracle_runner(chip, steps):
    from helpers import Oracle
    return Oracle(chip).run(steps, status=True)


def device_runner(chip, steps):
    return chip.sim_raw(steps, "simple", steps=True, fired=True, potentials=True, status=True)


def demo_arch():
    return os.path.join(GOLDEN, "src", "example_arch.yaml")


def truenorth_arch():
    return dse.arch_yaml(1.0, tiles=8)
