"""Generated architectures, written as flat JSON-lines descriptions.

The GPU box has no copy of the reference tree, so the Loihi-like architecture of
BASELINE config 4 is described here programmatically instead of being read from
arch/loihi_large.yaml: a 2-D mesh of tiles with four cores each, the buffer inside
the dendrite unit, LIF + input somas, three current-based synapse units, an
accumulator and a delay-line dendrite. Energy / latency figures are the published
Loihi estimates (Davies et al., "Loihi: A Neuromorphic Manycore Processor with
On-Chip Learning", IEEE Micro 2018) that the reference's Loihi descriptions use.
"""
import json


def _unit(model, **costs):
    attrs = {"model": model}
    attrs.update(costs)
    return {"model": model, "log_energy": False, "log_latency": False, "update_every_timestep": False,
            "attrs": attrs}


def loihi_large(tiles=1024, width=256, height=128, cores_per_tile=4, buffer_position="dendrite",
                buffer_inside_unit=True):
    """Records of a Loihi-like mesh; `tiles` may be cut below width*height (the mesh
    dimensions, hence tile coordinates and hop costs, stay those of the full chip)."""
    recs = [["noc", "loihi_chip", {"width": width, "height": height, "link_buffer_size": 16,
                                   "sync": [[1, 0.6e-6], [2, 1.0e-6], [4, 1.4e-6], [29, 1.8e-6]]}]]
    soma_lif = _unit("leaky_integrate_fire", energy_access_neuron=51.2e-12, latency_access_neuron=6.0e-9,
                     energy_update_neuron=21.6e-12, latency_update_neuron=3.7e-9,
                     energy_spike_out=69.3e-12, latency_spike_out=30.0e-9)
    soma_in = _unit("input", energy_access_neuron=0.0, latency_access_neuron=0.0, energy_update_neuron=0.0,
                    latency_update_neuron=0.0, energy_spike_out=0.0, latency_spike_out=0.0)
    for t in range(tiles):
        recs.append(["tile", f"loihi_tile[{t}]", {
            "energy_north_hop": 4.2e-12, "latency_north_hop": 6.5e-9, "energy_east_hop": 3.0e-12,
            "latency_east_hop": 4.1e-9, "energy_south_hop": 4.2e-12, "latency_south_hop": 6.5e-9,
            "energy_west_hop": 3.0e-12, "latency_west_hop": 4.1e-9, "log_energy": False}])
        for c in range(cores_per_tile):
            recs.append(["core", f"loihi_core[{c}]", {"buffer_position": buffer_position,
                                                     "buffer_inside_unit": buffer_inside_unit,
                                                     "max_neurons_supported": 1024, "log_energy": False}])
            recs.append(["axon_in", "loihi_in", 0.0, 16.0e-9])
            recs.append(["unit", "synapse", "loihi_dense_synapse",
                         _unit("current_based", energy_process_spike=35.5e-12, latency_process_spike=3.8e-9)])
            recs.append(["unit", "synapse", "loihi_sparse_synapse",
                         _unit("current_based", energy_process_spike=33.6e-12, latency_process_spike=4.7e-9)])
            recs.append(["unit", "synapse", "loihi_conv_synapse",
                         _unit("current_based", latency_process_spike=3.1e-9, energy_process_spike=24.0e-12)])
            recs.append(["unit", "dendrite", "loihi_dendrites",
                         _unit("accumulator", energy_update=0.0, latency_update=0.0)])
            recs.append(["unit", "dendrite", "loihi_dendrites_delay",
                         _unit("accumulator_with_delay", energy_update=0.0, latency_update=0.0)])
            recs.append(["unit", "soma", "loihi_lif", soma_lif])
            recs.append(["unit_range", "soma", "loihi_inputs", 0, 1023, soma_in])
            recs.append(["axon_out", "loihi_out", 111.0e-12, 5.1e-9])
    recs.append(["end_arch"])
    return recs


def write_flat(records, path, synth=None):
    with open(path, "w") as f:
        for r in records:
            f.write(json.dumps(r, separators=(",", ":")))
            f.write("\n")
        if synth is not None:
            f.write(json.dumps(["synth", synth], separators=(",", ":")) + "\n")
        f.write('["end_net"]\n')
