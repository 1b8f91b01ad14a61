"""Design-space sweep of BASELINE config 5: many independent simulations of one synthetic conv SNN
on TrueNorth-shaped chips, batched on a GPU (sfe_batch_load / sfe_batch_sim, include/sanafe_b200.h).

The design points are (neurons mapped per core) x (multiplier on every energy / latency figure of the
chip). The architecture has the shape of the reference's arch/truenorth.yaml (64 x 64 mesh of tiles with
one 256-neuron core: `truenorth` soma, `current_based` synapses, `accumulator` dendrite, buffer before the
soma); that file ships with all costs 0 "for the sweep to override", so base costs are given here. The
network has the layer shapes of the reference's DVS gesture net (snn/dvs.yaml: 32x32x1 -> conv 3x3x16
stride 2 -> conv 3x3x32 -> dense 11) with integer weights and TrueNorth somas (no `random_mask`, hence no
std::rand(): SURVEY Appendix B-7). Both are written as the reference's YAML formats and go through the
ordinary front-end. The GPU box has no reference tree: nothing here reads it.
"""
import ctypes as C
import os

import numpy as np

from . import RunData, SanafeError, SpikingChip, TIMING, lib, load_arch, load_net

# One published figure anchors the base costs: 26 pJ per synaptic event (Merolla et al., "A million
# spiking-neuron integrated circuit with a scalable communication network and interface", Science 2014);
# the others are same-order placeholders that the sweep scales.
BASE_COSTS = {
    "energy_hop": 2.0e-12, "latency_hop": 1.0e-9,
    "energy_message_in": 1.0e-12, "latency_message_in": 2.0e-9,
    "energy_access_neuron": 4.0e-12, "latency_access_neuron": 1.0e-9,
    "energy_update_neuron": 6.0e-12, "latency_update_neuron": 2.0e-9,
    "energy_spike_out": 10.0e-12, "latency_spike_out": 4.0e-9,
    "energy_process_spike": 26.0e-12, "latency_process_spike": 1.5e-9,
    "energy_update": 0.5e-12, "latency_update": 0.5e-9,
    "energy_message_out": 8.0e-12, "latency_message_out": 3.0e-9,
}

LAYERS = {"in": 32 * 32, "c1": 15 * 15 * 16, "c2": 13 * 13 * 32, "out": 11}


def arch_yaml(multiplier=1.0, tiles=4096):
    """Architecture YAML of a TrueNorth-shaped chip with every cost = base x multiplier. `tiles` may be cut
    below 64 x 64 (mesh dimensions, hence coordinates and hop counts, stay those of the full chip)."""
    c = {k: repr(v * multiplier) for k, v in BASE_COSTS.items()}
    return f"""architecture:
  name: truenorth_chip
  attributes:
    topology: mesh
    width: 64
    height: 64
    link_buffer_size: 1
    sync_model: fixed
    latency_sync: 0.0
  tile:
    - name: truenorth_tile[0..{tiles - 1}]
      attributes:
        energy_north_hop: {c['energy_hop']}
        latency_north_hop: {c['latency_hop']}
        energy_east_hop: {c['energy_hop']}
        latency_east_hop: {c['latency_hop']}
        energy_south_hop: {c['energy_hop']}
        latency_south_hop: {c['latency_hop']}
        energy_west_hop: {c['energy_hop']}
        latency_west_hop: {c['latency_hop']}
      core:
        - name: truenorth_core
          attributes:
            buffer_position: soma
            buffer_inside_unit: false
            max_neurons_supported: 256
          axon_in:
            - name: core_in
              attributes:
                energy_message_in: {c['energy_message_in']}
                latency_message_in: {c['latency_message_in']}
          soma:
            - name: core_soma
              attributes:
                model: truenorth
                energy_access_neuron: {c['energy_access_neuron']}
                latency_access_neuron: {c['latency_access_neuron']}
                energy_update_neuron: {c['energy_update_neuron']}
                latency_update_neuron: {c['latency_update_neuron']}
                energy_spike_out: {c['energy_spike_out']}
                latency_spike_out: {c['latency_spike_out']}
          synapse:
            - name: core_synapses
              attributes:
                model: current_based
                energy_process_spike: {c['energy_process_spike']}
                latency_process_spike: {c['latency_process_spike']}
          dendrite:
            - name: core_dendrites
              attributes:
                model: accumulator
                energy_update: {c['energy_update']}
                latency_update: {c['latency_update']}
          axon_out:
            - name: core_out
              attributes:
                energy_message_out: {c['energy_message_out']}
                latency_message_out: {c['latency_message_out']}
"""


def _flow(values):
    return "[" + ", ".join(str(int(v)) for v in values) + "]"


def snn_yaml(neurons_per_core, seed=1, active_fraction=0.15):
    """SNN YAML of the conv net, mapped `neurons_per_core` neurons to a core, layer after layer, cores in
    id order (a layer never shares a core with the next one). The input layer is bias-driven (a seeded
    `active_fraction` of the pixels fire every other step); weights are small seeded integers."""
    if not 1 <= neurons_per_core <= 256:
        raise ValueError("neurons_per_core must be in 1..256")
    rng = np.random.default_rng(seed)
    bias = (rng.random(LAYERS["in"]) < active_fraction).astype(int)
    w1 = rng.integers(-2, 4, size=3 * 3 * 1 * 16)
    w2 = rng.integers(-2, 3, size=3 * 3 * 16 * 32)
    w3 = rng.integers(-3, 4, size=LAYERS["c2"] * LAYERS["out"])
    lines = ["network:", "  name: dse_conv", "  groups:"]
    lines += ["  - name: in",
              "    attributes: {threshold: 2, reset: 0, log_spikes: true}",
              "    neurons:"]
    # runs of equal bias keep the file short
    start = 0
    for i in range(1, LAYERS["in"] + 1):
        if i == LAYERS["in"] or bias[i] != bias[start]:
            key = f"{start}..{i - 1}" if i - 1 > start else f"{start}"
            lines.append(f"    - {{{key}: {{bias: {bias[start]}}}}}")
            start = i
    for name, thr in (("c1", 3), ("c2", 4), ("out", 6)):
        lines += [f"  - name: {name}",
                  f"    attributes: {{threshold: {thr}, reset: 0, leak: 1, reverse_threshold: -8, "
                  f"reverse_reset_mode: saturate, log_spikes: true{', log_potential: true' if name == 'out' else ''}}}",
                  "    neurons:",
                  f"    - {{0..{LAYERS[name] - 1}: {{}}}}"]
    lines += ["  edges:",
              "  - in -> c1:", "      type: conv2d", "      input_width: 32", "      input_height: 32",
              "      input_channels: 1", "      kernel_width: 3", "      kernel_height: 3", "      kernel_count: 16",
              "      stride_width: 2", "      stride_height: 2", f"      weight: {_flow(w1)}",
              "  - c1 -> c2:", "      type: conv2d", "      input_width: 15", "      input_height: 15",
              "      input_channels: 16", "      kernel_width: 3", "      kernel_height: 3", "      kernel_count: 32",
              "      stride_width: 1", "      stride_height: 1", f"      weight: {_flow(w2)}",
              "  - c2 -> out:", "      type: dense", f"      weight: {_flow(w3)}",
              "mappings:"]
    core = 0
    for name in ("in", "c1", "c2", "out"):
        n = LAYERS[name]
        for a in range(0, n, neurons_per_core):
            b = min(n, a + neurons_per_core) - 1
            key = f"{name}.{a}..{b}" if b > a else f"{name}.{a}"
            lines.append(f"- {{'{key}': {{core: '{core}.0'}}}}")
            core += 1
    if core > 4096:
        raise ValueError(f"{core} cores needed, the chip has 4096")
    return "\n".join(lines) + "\n", core


def cores_needed(neurons_per_core):
    return sum((n + neurons_per_core - 1) // neurons_per_core for n in LAYERS.values())


def sweep_points(n_mappings=32, n_multipliers=32):
    """The design points of config 5: neurons per core 8, 16, ... 256 x cost multipliers 2^(k/8 - 2)."""
    npcs = [256 * (i + 1) // n_mappings for i in range(n_mappings)]
    mults = [2.0 ** (k / 8.0 - 2.0) for k in range(n_multipliers)]
    return [(npc, m) for npc in npcs for m in mults]


class Sweep:
    """Chips of a list of design points on one device (device < 0: host-only chips, lowering only)."""

    def __init__(self, points, workdir, device=0, seed=1, host_threads=0, share=None):
        """share: another Sweep whose parsed architectures / networks are reused (descriptions are only
        read by load)."""
        self.points = list(points)
        self.host_threads = host_threads
        os.makedirs(workdir, exist_ok=True)
        self._arch, self._net = (share._arch, share._net) if share is not None else ({}, {})
        self.chips = []
        # Descriptions first, in parallel: the C library drops the GIL while it parses, and parsing the conv net
        # (about a second) is what a sweep spends most of its host time on.
        arch_jobs, net_jobs = {}, {}
        for npc, mult in self.points:
            tiles = cores_needed(npc)
            akey = (mult, tiles)
            if akey not in self._arch and akey not in arch_jobs:
                path = os.path.join(workdir, f"arch_m{len(self._arch) + len(arch_jobs)}.yaml")
                with open(path, "w") as f:
                    f.write(arch_yaml(mult, tiles=tiles))
                arch_jobs[akey] = path
            # a parsed network only keeps core ADDRESSES (Neuron::map_to_core, src/network.cpp:141-149), so one
            # parse serves every cost variant of the chip
            if npc not in self._net and npc not in net_jobs:
                path = os.path.join(workdir, f"snn_n{npc}.yaml")
                if not os.path.exists(path):
                    with open(path, "w") as f:
                        f.write(snn_yaml(npc, seed)[0])
                net_jobs[npc] = (path, akey)
        for akey, path in arch_jobs.items():
            self._arch[akey] = load_arch(path)
        if net_jobs:
            from concurrent.futures import ThreadPoolExecutor
            workers = max(1, min(len(net_jobs), host_threads or (os.cpu_count() or 8)))
            with ThreadPoolExecutor(max_workers=workers) as pool:
                parsed = pool.map(lambda job: load_net(job[0], self._arch[job[1]]), net_jobs.values())
                for npc, net in zip(net_jobs, parsed):
                    self._net[npc] = net
        nets = []
        for npc, mult in self.points:
            chip = SpikingChip(self._arch[(mult, cores_needed(npc))], device=device)
            self.chips.append(chip)
            nets.append(self._net[npc])
        n = len(self.chips)
        chip_arr = (C.c_void_p * n)(*[c._h for c in self.chips])
        net_arr = (C.c_void_p * n)(*[x._h for x in nets])
        if lib().sfe_batch_load(chip_arr, net_arr, n, host_threads) != 0:
            raise SanafeError(lib().sfe_last_error().decode())
        self._chip_arr = chip_arr

    def sim(self, timesteps, timing_model="simple"):
        """One batched run of all chips; returns the list of RunData (one per design point)."""
        n = len(self.chips)
        out = (RunData * n)()
        if lib().sfe_batch_sim(self._chip_arr, n, timesteps, TIMING[timing_model], None, out, self.host_threads) != 0:
            raise SanafeError(lib().sfe_last_error().decode())
        return list(out)
