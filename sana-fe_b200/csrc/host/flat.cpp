// flat.cpp — reader for the flat JSON-lines description (one record per line).
//
// The flat form is a pre-digested Architecture + SNN description: typing of
// scalars, attribute forwarding flags and range expansion have already been
// decided by whoever wrote it (oracle/yaml_to_flat.py for the oracle fixtures,
// sanafe_b200.archgen for generated architectures). Records are replayed through
// the same builder calls the YAML front-end uses.
#include <fstream>

#include "desc.hpp"
#include "json.hpp"

namespace sfe
{
namespace
{
Attr to_attr(const Json &j)
{
    Attr a;
    if (j.is_bool()) a.value = j.boolean();
    else if (j.is_int()) a.value = static_cast<int>(j.integer());
    else if (j.is_double()) a.value = j.number();
    else if (j.is_string()) a.value = j.str();
    else if (j.is_list())
    {
        std::vector<Attr> l;
        l.reserve(j.list().size());
        for (const Json &e : j.list()) l.push_back(to_attr(e));
        a.value = std::move(l);
    }
    else if (j.is_map())
    {
        std::vector<Attr> l;
        for (const auto &kv : j.map())
        {
            Attr e = to_attr(kv.second);
            e.name = kv.first;
            l.push_back(std::move(e));
        }
        a.value = std::move(l);
    }
    else throw std::runtime_error("flat description: null attribute value");
    return a;
}

AttrMap to_attr_map(const Json &j) // [[key, value, fwd_syn, fwd_den, fwd_soma], ...]
{
    AttrMap m;
    for (const Json &e : j.list())
    {
        const JsonList &t = e.list();
        Attr a = to_attr(t.at(1));
        a.forward_to_synapse = t.at(2).boolean();
        a.forward_to_dendrite = t.at(3).boolean();
        a.forward_to_soma = t.at(4).boolean();
        m[t.at(0).str()] = std::move(a);
    }
    return m;
}

NeuronConfiguration to_neuron_config(const Json &j)
{
    NeuronConfiguration c;
    if (const Json *p = j.find("soma_hw_name")) c.soma_hw_name = p->str();
    if (const Json *p = j.find("default_synapse_hw_name")) c.default_synapse_hw_name = p->str();
    if (const Json *p = j.find("dendrite_hw_name")) c.dendrite_hw_name = p->str();
    if (const Json *p = j.find("log_spikes")) c.log_spikes = p->boolean();
    if (const Json *p = j.find("log_potential")) c.log_potential = p->boolean();
    if (const Json *p = j.find("attrs")) c.model_attributes = to_attr_map(*p);
    return c;
}

NeuronGroup::AttrLists to_attr_lists(const Json &j) // [[name, [values], fs, fd, fso], ...]
{
    NeuronGroup::AttrLists out;
    for (const Json &e : j.list())
    {
        const JsonList &t = e.list();
        std::vector<Attr> vals;
        vals.reserve(t.at(1).list().size());
        for (const Json &v : t.at(1).list())
        {
            Attr a = to_attr(v);
            a.forward_to_synapse = t.at(2).boolean();
            a.forward_to_dendrite = t.at(3).boolean();
            a.forward_to_soma = t.at(4).boolean();
            vals.push_back(std::move(a));
        }
        out[t.at(0).str()] = std::move(vals);
    }
    return out;
}

ModelInfo to_model_info(const Json &a)
{
    ModelInfo mi;
    mi.name = a.at("model").str();
    if (const Json *p = a.find("plugin")) mi.plugin_library_path = p->str();
    mi.log_energy = a.at("log_energy").boolean();
    mi.log_latency = a.at("log_latency").boolean();
    mi.update_every_timestep = a.at("update_every_timestep").boolean();
    for (const auto &kv : a.at("attrs").map()) mi.model_attributes[kv.first] = to_attr(kv.second);
    return mi;
}
} // namespace

void load_flat_file(const std::string &path, std::unique_ptr<Architecture> &arch,
        std::unique_ptr<SpikingNetwork> &net, std::optional<SynthRequest> &synth)
{
    std::ifstream in(path);
    if (!in) throw std::invalid_argument("Error: description file: failed to open (" + path + ").");
    std::string line;
    CoreConfiguration *core = nullptr;
    size_t tile_id = 0;
    size_t line_no = 0;
    while (std::getline(in, line))
    {
        ++line_no;
        if (line.empty()) continue;
        try
        {
            JsonReader reader(line.data(), line.data() + line.size());
            const Json rec = reader.parse();
            const JsonList &r = rec.list();
            const std::string &kind = r.at(0).str();
            if (kind == "noc")
            {
                const Json &a = r.at(2);
                NetworkOnChipConfiguration noc;
                noc.width_in_tiles = a.at("width").integer();
                noc.height_in_tiles = a.at("height").integer();
                noc.link_buffer_size = a.at("link_buffer_size").integer();
                for (const Json &kv : a.at("sync").list())
                    noc.ts_sync_delay_table.values[kv.list().at(0).integer()] = kv.list().at(1).number();
                arch = std::make_unique<Architecture>(r.at(1).str(), noc);
            }
            else if (kind == "tile")
            {
                const Json &a = r.at(2);
                TilePowerMetrics m;
                m.energy_north_hop = a.at("energy_north_hop").number();
                m.latency_north_hop = a.at("latency_north_hop").number();
                m.energy_east_hop = a.at("energy_east_hop").number();
                m.latency_east_hop = a.at("latency_east_hop").number();
                m.energy_south_hop = a.at("energy_south_hop").number();
                m.latency_south_hop = a.at("latency_south_hop").number();
                m.energy_west_hop = a.at("energy_west_hop").number();
                m.latency_west_hop = a.at("latency_west_hop").number();
                m.log_energy = a.at("log_energy").boolean();
                tile_id = arch->create_tile(r.at(1).str(), m).id;
            }
            else if (kind == "core")
            {
                const Json &a = r.at(2);
                CorePipelineConfiguration pc;
                pc.buffer_position =
                        parse_buffer_position(a.at("buffer_position").str(), a.at("buffer_inside_unit").boolean());
                pc.max_neurons_supported = a.at("max_neurons_supported").integer();
                pc.log_energy = a.at("log_energy").boolean();
                core = &arch->create_core(r.at(1).str(), tile_id, pc);
            }
            else if (kind == "axon_in") core->create_axon_in(r.at(1).str(), r.at(2).number(), r.at(3).number());
            else if (kind == "axon_out") core->create_axon_out(r.at(1).str(), r.at(2).number(), r.at(3).number());
            else if (kind == "unit")
                core->merge_or_create_hardware_unit(r.at(2).str(), to_model_info(r.at(3)), r.at(1).str());
            else if (kind == "unit_range")
                core->merge_or_create_hardware_unit(r.at(2).str(), to_model_info(r.at(5)), r.at(1).str(), true,
                        static_cast<int>(r.at(3).integer()), static_cast<int>(r.at(4).integer()));
            else if (kind == "end_arch") net = std::make_unique<SpikingNetwork>("net");
            else if (kind == "group")
                net->create_neuron_group(r.at(1).str(), r.at(2).integer(), to_neuron_config(r.at(3)));
            else if (kind == "neurons")
            {
                NeuronGroup &g = net->group(r.at(1).str());
                const NeuronConfiguration c = to_neuron_config(r.at(4));
                for (long long i = r.at(2).integer(); i <= r.at(3).integer(); ++i) g.neurons.at(i).set_attributes(c);
            }
            else if (kind == "edge")
            {
                Neuron &src = net->group(r.at(1).str()).neurons.at(r.at(2).integer());
                Neuron &dst = net->group(r.at(3).str()).neurons.at(r.at(4).integer());
                Connection &con = src.edges_out[src.connect_to_neuron(dst)];
                con.synapse_attributes = to_attr_map(r.at(5).at("synapse_attrs"));
                con.dendrite_attributes = to_attr_map(r.at(5).at("dendrite_attrs"));
            }
            else if (kind == "conv2d")
            {
                const Json &p = r.at(3);
                Conv2DParameters c;
                c.input_width = p.at("input_width").integer();
                c.input_height = p.at("input_height").integer();
                c.input_channels = p.at("input_channels").integer();
                c.kernel_width = p.at("kernel_width").integer();
                c.kernel_height = p.at("kernel_height").integer();
                c.kernel_count = p.at("kernel_count").integer();
                c.stride_width = p.at("stride_width").integer();
                c.stride_height = p.at("stride_height").integer();
                net->group(r.at(1).str()).connect_neurons_conv2d(net->group(r.at(2).str()), to_attr_lists(r.at(4)), c);
            }
            else if (kind == "dense")
                net->group(r.at(1).str()).connect_neurons_dense(net->group(r.at(2).str()), to_attr_lists(r.at(3)));
            else if (kind == "sparse")
            {
                std::vector<std::pair<size_t, size_t>> pairs;
                for (const Json &pr : r.at(3).list())
                    pairs.emplace_back(pr.list().at(0).integer(), pr.list().at(1).integer());
                net->group(r.at(1).str())
                        .connect_neurons_sparse(net->group(r.at(2).str()), to_attr_lists(r.at(4)), pairs);
            }
            else if (kind == "map")
            {
                NeuronGroup &g = net->group(r.at(1).str());
                const Json &hw = r.at(6);
                for (long long i = r.at(2).integer(); i <= r.at(3).integer(); ++i)
                {
                    Neuron &n = g.neurons.at(i);
                    if (const Json *p = hw.find("synapse")) n.default_synapse_hw_name = p->str();
                    if (const Json *p = hw.find("dendrite")) n.dendrite_hw_name = p->str();
                    if (const Json *p = hw.find("soma")) n.soma_hw_name = p->str();
                    const size_t t = r.at(4).integer();
                    const size_t c = r.at(5).integer();
                    if (t >= arch->tiles.size()) throw std::invalid_argument("Tile ID >= tile count");
                    if (c >= arch->tiles[t].cores.size()) throw std::invalid_argument("Core ID >= core count");
                    n.map_to_core(arch->tiles[t].cores[c]);
                }
            }
            else if (kind == "synth")
            {
                const Json &j = r.at(1);
                SynthRequest q;
                q.spec.cores = j.at("cores").integer();
                q.spec.neurons_per_core = j.at("neurons_per_core").integer();
                q.spec.dest_cores = j.at("dest_cores").integer();
                q.spec.syn_per_axon = j.at("syn_per_axon").integer();
                q.spec.seed = j.at("seed").integer();
                q.spec.bias_permille = j.at("bias_permille").integer();
                q.spec.bias = j.at("bias").number();
                q.spec.threshold = j.at("threshold").number();
                q.spec.reset = j.at("reset").number();
                q.spec.leak_decay = j.at("leak_decay").number();
                q.spec.w_min = j.at("w_min").integer();
                q.spec.w_max = j.at("w_max").integer();
                q.spec.max_delay = j.at("max_delay").integer();
                q.spec.log_spikes = j.at("log_spikes").integer();
                q.spec.log_potential_n = j.at("log_potential_n").integer();
                q.soma_hw_name = j.at("soma_hw_name").str();
                q.synapse_hw_name = j.at("synapse_hw_name").str();
                q.dendrite_hw_name = j.at("dendrite_hw_name").str();
                synth = q;
            }
            else if (kind == "end_net") break;
            else throw std::invalid_argument("unknown record '" + kind + "'");
        }
        catch (const std::exception &e)
        {
            throw std::invalid_argument(path + ":" + std::to_string(line_no) + ": " + e.what());
        }
    }
    if (!arch) throw std::invalid_argument(path + ": no architecture defined");
    if (!net) net = std::make_unique<SpikingNetwork>("net");
}
} // namespace sfe
