// lower.hpp — load-time lowering: Architecture + SpikingNetwork -> sfe_tables.
//
// This is the new engine's restatement of SpikingChip::load (reference
// src/chip.cpp:129-408, 1263-1391; src/core.cpp:61-184; src/mapped.cpp:27-188;
// src/pipeline.cpp:59-85): it reproduces every index the reference assigns
// (in-core neuron order, per-unit addresses, connection order, one axon per
// (pre-neuron, destination core) pair) and emits them as SoA / CSR arrays.
#ifndef SFE_LOWER_HPP_
#define SFE_LOWER_HPP_

#include <map>
#include <optional>
#include <string>
#include <vector>

#include "desc.hpp"
#include "sanafe_b200.h"

namespace sfe
{

struct HostTables
{
    std::vector<sfe_tile_desc> tiles;
    std::vector<sfe_core_desc> cores;
    std::vector<sfe_soma_class> soma_classes;
    std::vector<sfe_cost_class> cost_classes;
    std::vector<uint32_t> neuron_class, neuron_aux;
    std::vector<double> neuron_bias, neuron_potential0;
    std::vector<uint32_t> axon_out_begin, axon_out_target;
    std::vector<sfe_input_desc> inputs;
    std::vector<uint8_t> input_spikes;
    std::vector<sfe_hh_init> hh;
    std::vector<uint32_t> probes;
    std::vector<sfe_axon_in> axons_in;
    std::vector<uint32_t> axon_src;
    std::vector<double> syn_weight;
    std::vector<uint32_t> syn_meta;
    std::optional<sfe_synth_spec> synth;
    std::vector<sfe_noise_desc> noise;  // LIF file noise: one record per neuron of a unit with a stream
    std::vector<double> noise_values;
    std::vector<uint32_t> u_probes;     // LIF neurons with log_u, in trace order
    std::vector<uint32_t> neuron_taps;  // "taps" dendrites (empty when the network has none)
    std::vector<sfe_taps_desc> taps;
    std::vector<double> taps_values;
    // out-of-tree soma device models (include/sfe_device_model.h): one block per model, instances in device order
    struct DeviceModelBlock
    {
        const sfe_device_model_desc *desc{nullptr};
        std::vector<uint32_t> neurons;
        std::vector<std::vector<double>> state_rows, param_rows; // per instance, as the lowering fills them
        std::vector<double> state_init, params;                  // SoA, built by finalize_view
        bool numbered{false}; // neuron_aux already rewritten to chip-wide instance indices
    };
    std::vector<DeviceModelBlock> device_models;
    std::vector<sfe_device_model_block> device_model_view;
    uint32_t input_seed_base{0}; // "input" units created in this process before this chip (set by the chip)
    uint32_t n_poisson_cols{0};
    uint32_t n_rand_cols{0};     // TrueNorth neurons with random_mask != 0
    sfe_tables view{};

    // host-only naming: device index <-> (group, offset); groups in lexicographic order
    struct NeuronName
    {
        uint32_t group;
        uint32_t offset;
        bool log_spikes;
        bool log_potential;
    };
    std::vector<std::string> group_names;
    std::vector<NeuronName> names;                      // by device index
    std::vector<std::vector<uint32_t>> group_to_device; // [group][offset] -> device index
    std::vector<std::string> core_names;                // "tile.core" per core id
    std::vector<std::string> soma_model_of_neuron_unit; // reserved

    void finalize_view(const Architecture &arch);
    int64_t find_neuron(const std::string &group, uint64_t offset) const;
};

// Network described object by object.
void lower_network(const Architecture &arch, const SpikingNetwork &net, HostTables &out);
// Bulk path: the synthetic network of sfe_synth.h. materialize = also fill the
// synapse arrays on the host (tests, CPU oracle); otherwise they are generated
// on the device from the spec.
void lower_synthetic(const Architecture &arch, const SynthRequest &req, bool materialize, HostTables &out);

// Number of built-in "input" soma units an architecture instantiates (every InputModel the
// reference's SpikingChip constructor creates advances the process-wide seed counter,
// src/chip.cpp:61-92, src/models.hpp:347,366).
uint32_t count_input_units(const Architecture &arch);

// MappedNeuron.set_attributes for one numeric soma attribute: patches the host
// tables (splitting the neuron's parameter class when needed). Returns true if
// the class table changed (device must re-upload classes + neuron_class).
enum class PatchKind
{
    ignored,  // the neuron's model does not know the key (the reference ignores it too)
    bias,     // host bias table changed: upload the vector before the next step
    classes,  // soma class table / per-neuron class ids changed
    potential // the membrane potential itself was set (LIF "potential")
};
PatchKind patch_neuron_attribute(HostTables &t, uint32_t neuron, const std::string &name, double value);

} // namespace sfe
#endif
