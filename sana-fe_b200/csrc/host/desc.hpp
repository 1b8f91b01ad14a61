// desc.hpp — host-side description objects: Architecture and SpikingNetwork.
//
// Mirrors the reference's builder surface (names, argument meaning, error
// behaviour) so that front-ends and bindings written against the reference map
// one to one:
//   Architecture / TileConfiguration / CoreConfiguration   src/arch.hpp:70-203, src/arch.cpp:24-180
//   SpikingNetwork / NeuronGroup / Neuron / Connection     src/network.hpp:30-195, src/network.cpp:55-605
//   ModelAttribute                                         src/attribute.hpp:41-178
// These objects are only read during SpikingChip construction and load(); the
// per-timestep path never touches them.
#ifndef SFE_DESC_HPP_
#define SFE_DESC_HPP_

#include <cstddef>
#include <cstdint>
#include <map>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <variant>
#include <vector>

#include "sfe_synth.h"

namespace sfe
{

class HardwareMappingError : public std::runtime_error // src/mapped.hpp:30-38
{
public:
    explicit HardwareMappingError(const std::string &m) : std::runtime_error(m) {}
};

// ---- attributes -----------------------------------------------------------
struct Attr
{
    std::variant<bool, int, double, std::string, std::vector<Attr>> value;
    std::optional<std::string> name;
    bool forward_to_synapse{true};
    bool forward_to_dendrite{true};
    bool forward_to_soma{true};

    Attr() : value(false) {}
    static Attr of(bool v) { Attr a; a.value = v; return a; }
    static Attr of(int v) { Attr a; a.value = v; return a; }
    static Attr of(double v) { Attr a; a.value = v; return a; }
    static Attr of(std::string v) { Attr a; a.value = std::move(v); return a; }

    bool is_list() const { return std::holds_alternative<std::vector<Attr>>(value); }
    bool is_string() const { return std::holds_alternative<std::string>(value); }
    // Cast rules of src/attribute.hpp:43-94
    bool as_bool() const
    {
        if (auto *b = std::get_if<bool>(&value)) return *b;
        if (auto *i = std::get_if<int>(&value)) return *i != 0;
        throw std::runtime_error("Error: Attribute " + name.value_or("") + " cannot be cast to a bool ()");
    }
    int as_int() const
    {
        if (auto *i = std::get_if<int>(&value)) return *i;
        throw std::bad_variant_access();
    }
    double as_double() const
    {
        if (auto *d = std::get_if<double>(&value)) return *d;
        if (auto *i = std::get_if<int>(&value)) return static_cast<double>(*i);
        throw std::runtime_error("Error: Attribute" + (name ? " " + *name : std::string()) +
                " cannot be cast to a double");
    }
    const std::string &as_string() const
    {
        if (auto *s = std::get_if<std::string>(&value)) return *s;
        throw std::bad_variant_access();
    }
    const std::vector<Attr> &as_list() const
    {
        if (auto *l = std::get_if<std::vector<Attr>>(&value)) return *l;
        throw std::bad_variant_access();
    }
};
using AttrMap = std::map<std::string, Attr>;

template <typename T> struct LookupTable // src/utils.hpp:19-45
{
    std::map<size_t, T> values;
    T get(size_t x) const
    {
        if (values.empty()) throw std::runtime_error("Table is empty");
        auto it = values.upper_bound(x);
        if (it == values.begin()) return values.begin()->second;
        --it;
        return it->second;
    }
};

// ---- architecture ---------------------------------------------------------
enum BufferPosition : uint8_t // src/arch.hpp:41-49
{
    buffer_before_dendrite_unit = 0U,
    buffer_inside_dendrite_unit = 1U,
    buffer_before_soma_unit = 2U,
    buffer_inside_soma_unit = 3U,
    buffer_before_axon_out_unit = 4U,
};
BufferPosition parse_buffer_position(const std::string &pos, bool inside); // src/pipeline.cpp:268-310

struct ModelInfo // src/arch.hpp:51-59
{
    AttrMap model_attributes;
    std::optional<std::string> plugin_library_path;
    std::string name;
    bool log_energy{false};
    bool log_latency{false};
    bool update_every_timestep{false};
};

// One entry of a core's synapse/dendrite/soma section. A `name[a..b]` entry
// (src/yaml_arch.cpp:188-218) is kept as ONE family of b-a+1 identical units
// instead of being expanded: arch/loihi_large.yaml declares 4096 cores x 1030
// units, of which a mapped network touches a handful per core.
struct PipelineUnitConfiguration
{
    ModelInfo model_info;
    std::string name; // full name, or the part before '[' for a family
    bool implements_synapse{false};
    bool implements_dendrite{false};
    bool implements_soma{false};
    bool is_range{false};
    int range_first{0};
    int range_last{0};

    // index of `unit_name` inside this entry, or -1
    int match(const std::string &unit_name) const;
    std::string instance_name(int i) const
    {
        return is_range ? name + "[" + std::to_string(i) + "]" : name;
    }
    int first_instance() const { return is_range ? range_first : 0; }
};

struct AxonInConfiguration { std::string name; double energy_message_in{0.0}; double latency_message_in{0.0}; };
struct AxonOutConfiguration { std::string name; double energy_message_out{0.0}; double latency_message_out{0.0}; };

struct CoreAddress { size_t parent_tile_id{}; size_t offset_within_tile{}; size_t id{}; };

struct CorePipelineConfiguration
{
    BufferPosition buffer_position{buffer_before_soma_unit};
    size_t max_neurons_supported{1024};
    bool log_energy{false};
};

struct CoreConfiguration
{
    CorePipelineConfiguration pipeline;
    std::string name;
    CoreAddress address;
    std::vector<AxonInConfiguration> axon_in;
    std::vector<PipelineUnitConfiguration> pipeline_hw;
    std::vector<AxonOutConfiguration> axon_out;

    AxonInConfiguration &create_axon_in(std::string n, double energy, double latency)
    {
        axon_in.push_back({std::move(n), energy, latency});
        return axon_in.back();
    }
    PipelineUnitConfiguration &create_hardware_unit(std::string n, const ModelInfo &m)
    {
        PipelineUnitConfiguration u;
        u.model_info = m;
        u.name = std::move(n);
        pipeline_hw.push_back(std::move(u));
        return pipeline_hw.back();
    }
    PipelineUnitConfiguration &create_hardware_unit_range(
            std::string base, int first, int last, const ModelInfo &m)
    {
        PipelineUnitConfiguration &u = create_hardware_unit(std::move(base), m);
        u.is_range = true;
        u.range_first = first;
        u.range_last = last;
        return u;
    }
    AxonOutConfiguration &create_axon_out(std::string n, double energy, double latency)
    {
        axon_out.push_back({std::move(n), energy, latency});
        return axon_out.back();
    }
    // yaml_merge_or_create_hardware_unit (src/yaml_arch.cpp:149-186): a name that
    // appears in several sections is ONE unit implementing several functions
    PipelineUnitConfiguration &merge_or_create_hardware_unit(const std::string &n, ModelInfo m,
            const std::string &section, bool is_range = false, int first = 0, int last = 0);
};

struct TilePowerMetrics
{
    double energy_north_hop{0.0}, latency_north_hop{0.0};
    double energy_east_hop{0.0}, latency_east_hop{0.0};
    double energy_south_hop{0.0}, latency_south_hop{0.0};
    double energy_west_hop{0.0}, latency_west_hop{0.0};
    bool log_energy{false};
};

struct TileConfiguration
{
    std::vector<CoreConfiguration> cores;
    TilePowerMetrics power_metrics;
    std::string name;
    size_t id{};
    size_t x{};
    size_t y{};
};

struct NetworkOnChipConfiguration
{
    LookupTable<double> ts_sync_delay_table;
    size_t width_in_tiles{1};
    size_t height_in_tiles{1};
    size_t link_buffer_size{0};
};

class Architecture
{
public:
    std::vector<TileConfiguration> tiles;
    LookupTable<double> ts_sync_delay_table;
    std::string name;
    size_t core_count{0};
    size_t max_cores_per_tile{0};
    size_t noc_width_in_tiles{1};
    size_t noc_height_in_tiles{1};
    size_t noc_buffer_size{0};

    Architecture(std::string n, const NetworkOnChipConfiguration &noc)
            : ts_sync_delay_table(noc.ts_sync_delay_table), name(std::move(n)),
              noc_width_in_tiles(noc.width_in_tiles), noc_height_in_tiles(noc.height_in_tiles),
              noc_buffer_size(noc.link_buffer_size)
    {
    }
    TileConfiguration &create_tile(std::string n, const TilePowerMetrics &m); // src/arch.cpp:89-104
    CoreConfiguration &create_core(std::string n, size_t parent_tile_id,
            const CorePipelineConfiguration &p); // src/arch.cpp:119-150
    std::vector<CoreConfiguration *> cores();
    std::vector<const CoreConfiguration *> cores() const;
};

// ---- network ----------------------------------------------------------------
struct NeuronConfiguration // src/network.hpp:30-38
{
    AttrMap model_attributes;
    std::optional<std::string> soma_hw_name;
    std::optional<std::string> default_synapse_hw_name;
    std::optional<std::string> dendrite_hw_name;
    std::optional<bool> log_spikes;
    std::optional<bool> log_potential;
};

struct NeuronAddress
{
    std::string group_name;
    std::optional<size_t> neuron_offset;
};

struct Conv2DParameters // src/network.hpp:47-57
{
    int input_width{}, input_height{}, input_channels{};
    int kernel_width{}, kernel_height{}, kernel_count{1};
    int stride_width{1}, stride_height{1};
};

class SpikingNetwork;
class NeuronGroup;

struct Connection // src/network.hpp:182-193 (only synapse_attributes reach the hardware: src/core.cpp:170-184)
{
    AttrMap synapse_attributes;
    AttrMap dendrite_attributes;
    std::string synapse_hw_name;
    NeuronAddress pre_neuron;
    NeuronAddress post_neuron;
    size_t id{};
};

bool is_reserved_neuron_attribute(const std::string &name); // src/attribute.hpp:24-36

class Neuron
{
public:
    std::vector<Connection> edges_out;
    AttrMap model_attributes;
    std::string soma_hw_name;
    std::string default_synapse_hw_name;
    std::string dendrite_hw_name;
    std::string parent_group_name;
    SpikingNetwork *parent_net{nullptr};
    size_t offset{};
    std::optional<CoreAddress> core_address;
    size_t mapping_order{};
    bool log_spikes{false};
    bool log_potential{false};

    Neuron(size_t off, SpikingNetwork &net, std::string group, const NeuronConfiguration &config);
    size_t connect_to_neuron(Neuron &dest);              // src/network.cpp:173-192
    void map_to_core(const CoreConfiguration &core);     // src/network.cpp:85-92
    void set_attributes(const NeuronConfiguration &c);   // src/network.cpp:94-131
};

class NeuronGroup
{
public:
    std::vector<Neuron> neurons;
    NeuronConfiguration default_neuron_config;
    std::string name;

    NeuronGroup(std::string n, SpikingNetwork &net, size_t count, const NeuronConfiguration &c);
    using AttrLists = std::map<std::string, std::vector<Attr>>;
    void connect_neurons_dense(NeuronGroup &dest, const AttrLists &lists);                 // src/network.cpp:567-605
    void connect_neurons_sparse(NeuronGroup &dest, const AttrLists &lists,
            const std::vector<std::pair<size_t, size_t>> &pairs);                          // src/network.cpp:229-276
    void connect_neurons_conv2d(NeuronGroup &dest, const AttrLists &lists,
            const Conv2DParameters &conv);                                                  // src/network.cpp:279-565
};

class SpikingNetwork
{
public:
    // std::map: iteration is lexicographic by group name, which fixes the
    // connection mapping order and the trace order
    std::map<std::string, std::unique_ptr<NeuronGroup>> groups;
    std::string name;
    size_t mapping_count{0};

    explicit SpikingNetwork(std::string n = "") : name(std::move(n)) {}
    SpikingNetwork(const SpikingNetwork &) = delete;
    SpikingNetwork &operator=(const SpikingNetwork &) = delete;
    NeuronGroup &create_neuron_group(const std::string &n, size_t count,
            const NeuronConfiguration &c); // src/network.cpp:143-160
    NeuronGroup &group(const std::string &n)
    {
        auto it = groups.find(n);
        if (it == groups.end()) throw std::out_of_range("map::at");
        return *it->second;
    }
    size_t update_mapping_count() { return ++mapping_count; }
};

// front-ends
struct SynthRequest // a "synth" record of a flat file: generator spec + unit names
{
    sfe_synth_spec spec{};
    std::string soma_hw_name, synapse_hw_name, dendrite_hw_name;
};
void load_flat_file(const std::string &path, std::unique_ptr<Architecture> &arch,
        std::unique_ptr<SpikingNetwork> &net, std::optional<SynthRequest> &synth);
std::unique_ptr<Architecture> load_arch_yaml(const std::string &path);
std::unique_ptr<SpikingNetwork> load_net_yaml(const std::string &path, Architecture &arch);
std::unique_ptr<SpikingNetwork> load_net_netlist(const std::string &path, Architecture &arch); // legacy format

} // namespace sfe
#endif
