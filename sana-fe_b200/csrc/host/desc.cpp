// desc.cpp — builder methods of the description objects (see desc.hpp).
#include "desc.hpp"

#include <algorithm>
#include <unordered_set>

namespace sfe
{

BufferPosition parse_buffer_position(const std::string &pos, const bool inside)
{
    // src/pipeline.cpp:268-310
    if (pos == "dendrite") return inside ? buffer_inside_dendrite_unit : buffer_before_dendrite_unit;
    if (pos == "soma") return inside ? buffer_inside_soma_unit : buffer_before_soma_unit;
    if (pos == "axon_out") return buffer_before_axon_out_unit;
    throw std::invalid_argument("Error: Buffer position not supported");
}

int PipelineUnitConfiguration::match(const std::string &unit_name) const
{
    if (!is_range) return (unit_name == name) ? 0 : -1;
    // "base[i]" with range_first <= i <= range_last
    if (unit_name.size() < name.size() + 3 || unit_name.compare(0, name.size(), name) != 0 ||
            unit_name[name.size()] != '[' || unit_name.back() != ']')
        return -1;
    const std::string idx = unit_name.substr(name.size() + 1, unit_name.size() - name.size() - 2);
    if (idx.empty() || idx.find_first_not_of("0123456789") != std::string::npos) return -1;
    const long i = std::stol(idx);
    return (i >= range_first && i <= range_last) ? static_cast<int>(i) : -1;
}

PipelineUnitConfiguration &CoreConfiguration::merge_or_create_hardware_unit(const std::string &n, ModelInfo m,
        const std::string &section, const bool is_range, const int first, const int last)
{
    auto set_flag = [&](PipelineUnitConfiguration &hw) {
        if (section == "synapse") hw.implements_synapse = true;
        else if (section == "dendrite") hw.implements_dendrite = true;
        else if (section == "soma") hw.implements_soma = true;
        else throw std::runtime_error("Section not recognized");
    };
    for (PipelineUnitConfiguration &hw : pipeline_hw)
    {
        if (hw.name == n && hw.is_range == is_range && hw.range_first == first && hw.range_last == last)
        {
            set_flag(hw);
            // std::map::merge keeps existing keys
            for (auto &kv : m.model_attributes) hw.model_info.model_attributes.insert(kv);
            if (m.plugin_library_path.has_value()) hw.model_info.plugin_library_path = m.plugin_library_path;
            return hw;
        }
    }
    PipelineUnitConfiguration &hw =
            is_range ? create_hardware_unit_range(n, first, last, m) : create_hardware_unit(n, m);
    set_flag(hw);
    return hw;
}

TileConfiguration &Architecture::create_tile(std::string n, const TilePowerMetrics &m)
{
    TileConfiguration t;
    t.power_metrics = m;
    t.name = std::move(n);
    t.id = tiles.size();
    t.x = t.id / noc_height_in_tiles;
    t.y = t.id % noc_height_in_tiles;
    tiles.push_back(std::move(t));
    return tiles.back();
}

CoreConfiguration &Architecture::create_core(
        std::string n, const size_t parent_tile_id, const CorePipelineConfiguration &p)
{
    if (parent_tile_id >= tiles.size()) throw std::invalid_argument("Tile ID out of range");
    TileConfiguration &tile = tiles[parent_tile_id];
    CoreConfiguration c;
    c.pipeline = p;
    c.name = std::move(n);
    c.address = {parent_tile_id, tile.cores.size(), core_count++};
    max_cores_per_tile = std::max(max_cores_per_tile, c.address.offset_within_tile + 1);
    tile.cores.push_back(std::move(c));
    return tile.cores.back();
}

std::vector<CoreConfiguration *> Architecture::cores()
{
    std::vector<CoreConfiguration *> out;
    for (auto &t : tiles)
        for (auto &c : t.cores) out.push_back(&c);
    return out;
}

std::vector<const CoreConfiguration *> Architecture::cores() const
{
    std::vector<const CoreConfiguration *> out;
    for (const auto &t : tiles)
        for (const auto &c : t.cores) out.push_back(&c);
    return out;
}

bool is_reserved_neuron_attribute(const std::string &name)
{
    static const std::unordered_set<std::string> reserved = {"soma_hw_name", "default_synapse_hw_name",
            "dendrite_hw_name", "log_spikes", "log_potential", "log_v"};
    return reserved.count(name) > 0;
}

Neuron::Neuron(const size_t off, SpikingNetwork &net, std::string group, const NeuronConfiguration &config)
        : parent_group_name(std::move(group)), parent_net(&net), offset(off)
{
    set_attributes(config);
}

void Neuron::set_attributes(const NeuronConfiguration &c)
{
    if (c.default_synapse_hw_name) default_synapse_hw_name = *c.default_synapse_hw_name;
    if (c.dendrite_hw_name) dendrite_hw_name = *c.dendrite_hw_name;
    if (c.soma_hw_name) soma_hw_name = *c.soma_hw_name;
    if (c.log_spikes) log_spikes = *c.log_spikes;
    if (c.log_potential) log_potential = *c.log_potential;
    for (const auto &[key, attribute] : c.model_attributes)
    {
        if (is_reserved_neuron_attribute(key))
        {
            throw std::invalid_argument("Reserved neuron attribute '" + key +
                    "' cannot be used as a model attribute. Pass it as a direct argument instead.");
        }
        model_attributes.insert_or_assign(key, attribute);
    }
}

size_t Neuron::connect_to_neuron(Neuron &dest)
{
    edges_out.emplace_back();
    Connection &edge = edges_out.back();
    edge.id = edges_out.size() - 1;
    edge.pre_neuron = {parent_group_name, offset};
    edge.post_neuron = {dest.parent_group_name, dest.offset};
    edge.synapse_hw_name = dest.default_synapse_hw_name;
    return edge.id;
}

void Neuron::map_to_core(const CoreConfiguration &core)
{
    core_address = core.address;
    mapping_order = parent_net->update_mapping_count();
}

NeuronGroup::NeuronGroup(std::string n, SpikingNetwork &net, const size_t count, const NeuronConfiguration &c)
        : default_neuron_config(c), name(std::move(n))
{
    neurons.reserve(count);
    for (size_t i = 0; i < count; ++i) neurons.emplace_back(i, net, name, c);
}

NeuronGroup &SpikingNetwork::create_neuron_group(
        const std::string &n, const size_t count, const NeuronConfiguration &c)
{
    if (groups.find(n) != groups.end())
        throw std::invalid_argument("Group: " + n + " already exists in SNN.");
    auto g = std::make_unique<NeuronGroup>(n, *this, count, c);
    NeuronGroup &ref = *g;
    groups.emplace(n, std::move(g));
    return ref;
}

void NeuronGroup::connect_neurons_sparse(NeuronGroup &dest, const AttrLists &lists,
        const std::vector<std::pair<size_t, size_t>> &pairs)
{
    size_t edge_idx = 0;
    for (const auto &[source_id, dest_id] : pairs)
    {
        if (source_id >= neurons.size()) throw std::invalid_argument("Error: src id is out of range.");
        if (dest_id >= dest.neurons.size()) throw std::invalid_argument("Error: dest nid is out of range.");
        Neuron &source = neurons[source_id];
        const size_t idx = source.connect_to_neuron(dest.neurons[dest_id]);
        Connection &con = source.edges_out[idx];
        AttrMap attributes;
        for (const auto &[key, values] : lists)
        {
            if (values.size() != pairs.size())
                throw std::invalid_argument("Error: Length of attribute list != number of defined edges.");
            attributes[key] = values.at(edge_idx);
        }
        con.synapse_attributes = attributes;
        con.dendrite_attributes = attributes;
        ++edge_idx;
    }
}

void NeuronGroup::connect_neurons_dense(NeuronGroup &dest, const AttrLists &lists)
{
    for (size_t s = 0; s < neurons.size(); ++s)
    {
        Neuron &source = neurons[s];
        source.edges_out.reserve(source.edges_out.size() + dest.neurons.size());
        for (size_t d = 0; d < dest.neurons.size(); ++d)
        {
            const size_t list_index = (s * dest.neurons.size()) + d;
            const size_t idx = source.connect_to_neuron(dest.neurons[d]);
            Connection &con = source.edges_out[idx];
            for (const auto &[key, values] : lists)
            {
                if (values.size() <= list_index)
                    throw std::invalid_argument("Not enough entries defined for attribute");
                const Attr &a = values[list_index];
                if (a.forward_to_synapse) con.synapse_attributes[key] = a;
                if (a.forward_to_dendrite) con.dendrite_attributes[key] = a;
            }
        }
    }
}

void NeuronGroup::connect_neurons_conv2d(NeuronGroup &dest, const AttrLists &lists, const Conv2DParameters &cv)
{
    auto require_positive = [](int v, const char *n) {
        if (v <= 0)
            throw std::invalid_argument("Error: Conv2D parameter '" + std::string(n) + "' must be > 0 (got " +
                    std::to_string(v) + ").");
    };
    require_positive(cv.input_width, "input_width");
    require_positive(cv.input_height, "input_height");
    require_positive(cv.input_channels, "input_channels");
    require_positive(cv.kernel_width, "kernel_width");
    require_positive(cv.kernel_height, "kernel_height");
    require_positive(cv.kernel_count, "kernel_count");
    require_positive(cv.stride_width, "stride_width");
    require_positive(cv.stride_height, "stride_height");
    if (cv.kernel_width > cv.input_width || cv.kernel_height > cv.input_height)
        throw std::invalid_argument("Error: Conv2D kernel larger than input with zero padding.");
    // No padding; output size (W - K)/S + 1   src/network.cpp:404-420
    const int out_w = (cv.input_width - cv.kernel_width) / cv.stride_width + 1;
    const int out_h = (cv.input_height - cv.kernel_height) / cv.stride_height + 1;
    const int out_c = cv.kernel_count;
    const size_t expect_in = static_cast<size_t>(cv.input_channels) * cv.input_width * cv.input_height;
    const size_t expect_out = static_cast<size_t>(out_c) * out_w * out_h;
    if (expect_in != neurons.size())
        throw std::invalid_argument("Expected " + std::to_string(expect_in) +
                " neurons in source group for convolution but there are " + std::to_string(neurons.size()) +
                " neurons.\n");
    if (expect_out != dest.neurons.size())
        throw std::invalid_argument("Expected " + std::to_string(expect_out) +
                " neurons in dest group for convolution but there are " + std::to_string(dest.neurons.size()) +
                " neurons.\n");
    // Size every source neuron's edge list up front (the same loop nest, counting): 3.5 M Connection objects of
    // the DVS net otherwise move several times while their vectors grow
    {
        std::vector<uint32_t> fan_out(neurons.size(), 0);
        for (int y_out = 0; y_out < out_h; ++y_out)
            for (int x_out = 0; x_out < out_w; ++x_out)
                for (int c_in = 0; c_in < cv.input_channels; ++c_in)
                    for (int ky = 0; ky < cv.kernel_height; ++ky)
                    {
                        const int y = y_out * cv.stride_height + ky;
                        if (y < 0 || y >= cv.input_height) continue;
                        for (int kx = 0; kx < cv.kernel_width; ++kx)
                        {
                            const int x = x_out * cv.stride_width + kx;
                            if (x < 0 || x >= cv.input_width) continue;
                            ++fan_out[(static_cast<size_t>(c_in) * cv.input_width * cv.input_height) +
                                    (static_cast<size_t>(y) * cv.input_width) + x];
                        }
                    }
        for (size_t i = 0; i < neurons.size(); ++i)
            neurons[i].edges_out.reserve(neurons[i].edges_out.size() + static_cast<size_t>(fan_out[i]) * out_c);
    }
    // Loop nest c_out -> y_out -> x_out -> c_in -> ky -> kx fixes the edge
    // creation order (src/network.cpp:300-370); indices are channel-major for
    // neurons and [y][x][c_in][c_out] for the filter (src/network.cpp:490-530)
    for (int c_out = 0; c_out < out_c; ++c_out)
        for (int y_out = 0; y_out < out_h; ++y_out)
            for (int x_out = 0; x_out < out_w; ++x_out)
            {
                const size_t dest_idx = (static_cast<size_t>(c_out) * out_w * out_h) +
                        (static_cast<size_t>(y_out) * out_w) + x_out;
                Neuron &d = dest.neurons.at(dest_idx);
                for (int c_in = 0; c_in < cv.input_channels; ++c_in)
                    for (int ky = 0; ky < cv.kernel_height; ++ky)
                    {
                        const int y = y_out * cv.stride_height + ky;
                        if (y < 0 || y >= cv.input_height) continue;
                        for (int kx = 0; kx < cv.kernel_width; ++kx)
                        {
                            const int x = x_out * cv.stride_width + kx;
                            if (x < 0 || x >= cv.input_width) continue;
                            const size_t src_idx = (static_cast<size_t>(c_in) * cv.input_width * cv.input_height) +
                                    (static_cast<size_t>(y) * cv.input_width) + x;
                            const size_t filter_idx =
                                    ((static_cast<size_t>(ky) * cv.kernel_width + kx) * cv.input_channels + c_in) *
                                            cv.kernel_count +
                                    c_out;
                            Neuron &s = neurons.at(src_idx);
                            const size_t idx = s.connect_to_neuron(d);
                            Connection &con = s.edges_out[idx];
                            for (const auto &[key, values] : lists)
                            {
                                if (values.size() <= filter_idx)
                                    throw std::invalid_argument("Not enough entries defined for attribute");
                                const Attr &a = values.at(filter_idx);
                                if (a.forward_to_dendrite) con.dendrite_attributes[key] = a;
                                if (a.forward_to_synapse) con.synapse_attributes[key] = a;
                            }
                        }
                    }
            }
}

} // namespace sfe
