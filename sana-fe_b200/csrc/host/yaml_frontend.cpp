// yaml_frontend.cpp — Architecture and SNN description readers over yaml.hpp.
//
// Same accepted formats, defaults and dispatch as the reference's RapidYAML
// front-end (arch: src/yaml_arch.cpp:28-594; SNN: src/yaml_snn.cpp:61-1056;
// attributes and scalar typing: src/yaml_common.cpp:103-320), expressed against
// the builder API of desc.hpp.
#include <algorithm>
#include <cerrno>
#include <cstdlib>
#include <fstream>
#include <sstream>

#include "desc.hpp"
#include "yaml.hpp"

namespace sfe
{
namespace
{
using yaml::Node;

[[noreturn]] void parse_error(const std::string &what, const Node &n)
{
    throw std::invalid_argument("Error: " + what + " (Line " + std::to_string(n.line) + ").\n");
}

// Scalar typing order int -> double -> bool -> string  (src/yaml_common.cpp:205-263)
bool parse_int(const std::string &s, int &out)
{
    if (s.empty()) return false;
    size_t i = (s[0] == '-' || s[0] == '+') ? 1 : 0;
    if (i >= s.size()) return false;
    int base = 10;
    if (s.size() > i + 2 && s[i] == '0' && (s[i + 1] == 'x' || s[i + 1] == 'X')) base = 16;
    else if (s.size() > i + 2 && s[i] == '0' && (s[i + 1] == 'b' || s[i + 1] == 'B')) base = 2;
    else if (s.size() > i + 2 && s[i] == '0' && (s[i + 1] == 'o' || s[i + 1] == 'O')) base = 8;
    const size_t digits = base == 10 ? i : i + 2;
    for (size_t k = digits; k < s.size(); ++k)
    {
        const char c = s[k];
        const bool ok = base == 16 ? std::isxdigit(static_cast<unsigned char>(c)) != 0
                : base == 2        ? (c == '0' || c == '1')
                : base == 8        ? (c >= '0' && c <= '7')
                                   : (c >= '0' && c <= '9');
        if (!ok) return false;
    }
    errno = 0;
    const long long v = std::strtoll((std::string(s[0] == '-' ? "-" : "") + s.substr(digits)).c_str(), nullptr, base);
    if (errno != 0 || v > 2147483647LL || v < -2147483648LL) return false;
    out = static_cast<int>(v);
    return true;
}

bool parse_double(const std::string &s, double &out)
{
    if (s.empty()) return false;
    const char c = s[0];
    if (!(std::isdigit(static_cast<unsigned char>(c)) || c == '-' || c == '+' || c == '.')) return false;
    char *end = nullptr;
    errno = 0;
    out = std::strtod(s.c_str(), &end);
    return end != nullptr && *end == '\0' && end != s.c_str();
}

bool parse_bool(const std::string &s, bool &out)
{
    if (s == "true" || s == "True" || s == "TRUE") { out = true; return true; }
    if (s == "false" || s == "False" || s == "FALSE") { out = false; return true; }
    return false;
}

Attr typed(const std::string &s)
{
    int i = 0;
    double d = 0.0;
    bool b = false;
    if (parse_int(s, i)) return Attr::of(i);
    if (parse_double(s, d)) return Attr::of(d);
    if (parse_bool(s, b)) return Attr::of(b);
    return Attr::of(s);
}

Attr to_attr(const Node &n) // yaml_parse_attribute  src/yaml_common.cpp:143-203
{
    Attr a;
    if (n.is_seq())
    {
        std::vector<Attr> l;
        l.reserve(n.seq.size());
        for (const Node &e : n.seq) l.push_back(to_attr(e));
        a.value = std::move(l);
    }
    else if (n.is_map())
    {
        std::vector<Attr> l;
        for (const auto &kv : n.map)
        {
            Attr e = to_attr(kv.second);
            e.name = kv.first;
            l.push_back(std::move(e));
        }
        a.value = std::move(l);
    }
    else if (n.is_scalar()) a = typed(n.scalar);
    else throw std::invalid_argument("Invalid attribute YAML node.\n");
    return a;
}

bool skip_key(const std::string &k) // src/yaml_common.cpp:30-36
{
    return k == "soma_hw_name" || k == "default_synapse_hw_name" || k == "dendrite_hw_name" || k == "log_spikes" ||
            k == "log_potential" || k == "synapse" || k == "dendrite" || k == "soma";
}

// description_parse_model_attributes_yaml  src/yaml_common.cpp:103-141
AttrMap model_attributes(const Node &n)
{
    AttrMap out;
    if (n.is_seq())
    {
        for (const Node &e : n.seq)
        {
            AttrMap sub = model_attributes(e);
            out.insert(sub.begin(), sub.end()); // existing keys win
        }
    }
    else if (n.is_map())
    {
        for (const auto &kv : n.map)
            if (!skip_key(kv.first)) out[kv.first] = to_attr(kv.second);
    }
    else
        throw std::invalid_argument("Error: Model attributes must be an ordered map or mapping of named attributes.\n");
    return out;
}

template <typename T> T required(const Node &n, const char *key);
const Node &required_node(const Node &n, const char *key)
{
    if (!n.is_map()) parse_error("Node should be a mapping", n);
    const Node *c = n.find(key);
    if (c == nullptr) parse_error(std::string("Value for key '") + key + "' not defined", n);
    return *c;
}
template <> std::string required<std::string>(const Node &n, const char *key)
{
    const Node &c = required_node(n, key);
    if (!c.is_scalar()) parse_error(std::string("Expected scalar for '") + key + "'", c);
    return c.scalar;
}
template <> double required<double>(const Node &n, const char *key)
{
    double d = 0.0;
    const Node &c = required_node(n, key);
    if (!c.is_scalar() || !parse_double(c.scalar, d)) parse_error(std::string("Expected a number for '") + key + "'", c);
    return d;
}
template <> int required<int>(const Node &n, const char *key)
{
    int i = 0;
    const Node &c = required_node(n, key);
    if (!c.is_scalar() || !parse_int(c.scalar, i)) parse_error(std::string("Expected an integer for '") + key + "'", c);
    return i;
}

bool optional_bool(const Node &n, const char *key, bool fallback)
{
    const Node *c = n.find(key);
    if (c == nullptr) return fallback;
    bool b = false;
    int i = 0;
    if (c->is_scalar() && parse_bool(c->scalar, b)) return b;
    if (c->is_scalar() && parse_int(c->scalar, i)) return i != 0;
    parse_error(std::string("Expected a bool for '") + key + "'", *c);
}

std::pair<size_t, size_t> parse_range(const std::string &s) // yaml_parse_range  src/yaml_common.cpp:266-319
{
    const size_t delim = s.find("..");
    if (delim == std::string::npos) throw std::runtime_error("Range delimiter '..' not found");
    const size_t open = s.find('['), close = s.find(']');
    const size_t start = open != std::string::npos ? open + 1 : 0;
    const size_t end = close != std::string::npos ? close : s.size();
    if (end <= start || delim <= start || delim >= end) throw std::runtime_error("Invalid range format");
    size_t first = 0, last = 0;
    try
    {
        first = std::stoull(s.substr(start, delim - start));
        last = std::stoull(s.substr(delim + 2, end - delim - 2));
    }
    catch (const std::exception &)
    {
        throw std::runtime_error("Invalid range string, failed to convert");
    }
    if (first > last) throw std::runtime_error("Invalid range; first > last");
    return {first, last};
}

bool has_range(const std::string &s) { return s.find("..") != std::string::npos; }

template <typename F> void for_each_entry(const Node &section, F &&f)
{
    if (section.is_seq())
        for (const Node &e : section.seq) f(e);
    else f(section);
}

// ------------------------------------------------------------------ architecture
void parse_core(const Node &core_node, size_t tile_id, Architecture &arch, const std::string &name)
{
    const Node &attrs = required_node(core_node, "attributes");
    CorePipelineConfiguration pc;
    pc.buffer_position = parse_buffer_position(required<std::string>(attrs, "buffer_position"),
            optional_bool(attrs, "buffer_inside_unit", false));
    pc.max_neurons_supported = static_cast<size_t>(required<int>(attrs, "max_neurons_supported"));
    pc.log_energy = optional_bool(attrs, "log_energy", false);
    CoreConfiguration &core = arch.create_core(name, tile_id, pc);
    // fixed section order (src/yaml_arch.cpp:246-293)
    for (const char *section : {"axon_in", "synapse", "dendrite", "soma", "axon_out"})
    {
        const Node *sec = core_node.find(section);
        if (sec == nullptr) parse_error(std::string("No ") + section + " section defined", core_node);
        for_each_entry(*sec, [&](const Node &unit) {
            const std::string uname = required<std::string>(unit, "name");
            const bool ranged = has_range(uname);
            std::pair<size_t, size_t> range{0, 0};
            if (ranged) range = parse_range(uname);
            const std::string base = uname.substr(0, uname.find('['));
            const std::string sec_name(section);
            if (sec_name == "axon_in" || sec_name == "axon_out")
            {
                const Node *ua = unit.find("attributes");
                if (ua == nullptr) parse_error("No attributes section defined", unit);
                for (size_t i = range.first; i <= range.second; ++i)
                {
                    const std::string full = ranged ? base + "[" + std::to_string(i) + "]" : uname;
                    if (sec_name == "axon_in")
                        core.create_axon_in(full, required<double>(*ua, "energy_message_in"), required<double>(*ua, "latency_message_in"));
                    else
                        core.create_axon_out(full, required<double>(*ua, "energy_message_out"), required<double>(*ua, "latency_message_out"));
                }
                return;
            }
            const Node &ua = required_node(unit, "attributes");
            ModelInfo mi;
            mi.name = required<std::string>(ua, "model");
            mi.log_energy = optional_bool(ua, "log_energy", false);
            mi.log_latency = optional_bool(ua, "log_latency", false);
            mi.update_every_timestep = optional_bool(ua, "update_every_timestep", false);
            if (const Node *p = ua.find("plugin"))
            {
                if (!p->is_scalar()) parse_error("Expected plugin path to be string", *p);
                mi.plugin_library_path = p->scalar;
            }
            mi.model_attributes = model_attributes(ua);
            if (ranged)
                core.merge_or_create_hardware_unit(base, mi, sec_name, true, static_cast<int>(range.first), static_cast<int>(range.second));
            else core.merge_or_create_hardware_unit(uname, mi, sec_name);
        });
    }
}

void parse_tile(const Node &tile_node, Architecture &arch)
{
    const std::string tile_name = required<std::string>(tile_node, "name");
    std::pair<size_t, size_t> range{0, 0};
    if (has_range(tile_name)) range = parse_range(tile_name);
    const Node &ta = required_node(tile_node, "attributes");
    for (size_t t = range.first; t <= range.second; ++t)
    {
        TilePowerMetrics m;
        m.energy_north_hop = required<double>(ta, "energy_north_hop");
        m.latency_north_hop = required<double>(ta, "latency_north_hop");
        m.energy_east_hop = required<double>(ta, "energy_east_hop");
        m.latency_east_hop = required<double>(ta, "latency_east_hop");
        m.energy_south_hop = required<double>(ta, "energy_south_hop");
        m.latency_south_hop = required<double>(ta, "latency_south_hop");
        m.energy_west_hop = required<double>(ta, "energy_west_hop");
        m.latency_west_hop = required<double>(ta, "latency_west_hop");
        m.log_energy = optional_bool(ta, "log_energy", false);
        const size_t tile_id = arch.create_tile(tile_name.substr(0, tile_name.find('[')) + "[" + std::to_string(t) + "]", m).id;
        const Node *cores = tile_node.find("core");
        if (cores == nullptr) parse_error("No core section defined", tile_node);
        for_each_entry(*cores, [&](const Node &core_node) {
            const std::string core_name = required<std::string>(core_node, "name");
            std::pair<size_t, size_t> cr{0, 0};
            if (has_range(core_name)) cr = parse_range(core_name);
            for (size_t c = cr.first; c <= cr.second; ++c)
                parse_core(core_node, tile_id, arch, core_name.substr(0, core_name.find('[')) + "[" + std::to_string(c) + "]");
        });
    }
}

// ------------------------------------------------------------------------ network
NeuronConfiguration neuron_attributes(const Node &n, const NeuronConfiguration &base)
{
    // yaml_parse_neuron_attributes  src/yaml_snn.cpp:331-394
    NeuronConfiguration cfg = base;
    if (n.is_seq())
    {
        for (const Node &e : n.seq) cfg = neuron_attributes(e, cfg);
        return cfg;
    }
    if (!n.is_map())
        throw std::invalid_argument("Error: Model attributes must be an ordered map or mapping of named attributes.\n");
    if (n.find("log_potential") != nullptr) cfg.log_potential = optional_bool(n, "log_potential", false);
    if (n.find("log_spikes") != nullptr) cfg.log_spikes = optional_bool(n, "log_spikes", false);
    if (const Node *p = n.find("synapse_hw_name")) cfg.default_synapse_hw_name = p->scalar;
    if (const Node *p = n.find("dendrite_hw_name")) cfg.dendrite_hw_name = p->scalar;
    if (const Node *p = n.find("soma_hw_name")) cfg.soma_hw_name = p->scalar;
    for (auto &[key, a] : model_attributes(n))
    {
        Attr v = a;
        v.forward_to_dendrite = true;
        v.forward_to_soma = true;
        cfg.model_attributes[key] = v;
    }
    if (const Node *d = n.find("dendrite"))
        for (auto &[key, a] : model_attributes(*d))
        {
            Attr v = a;
            v.forward_to_synapse = false;
            v.forward_to_soma = false;
            cfg.model_attributes[key] = v;
        }
    if (const Node *so = n.find("soma"))
        for (auto &[key, a] : model_attributes(*so))
        {
            Attr v = a;
            v.forward_to_synapse = false;
            v.forward_to_dendrite = false;
            cfg.model_attributes[key] = v;
        }
    return cfg;
}

size_t count_neurons(const Node &neurons) // description_count_neurons  src/yaml_snn.cpp:226-278
{
    if (!neurons.is_seq()) parse_error("Invalid neuron format, should be list", neurons);
    size_t count = 0;
    auto add = [&](const std::string &id) {
        if (has_range(id))
        {
            const auto r = parse_range(id);
            count += r.second - r.first + 1;
        }
        else ++count;
    };
    for (const Node &entry : neurons.seq)
    {
        if (entry.is_map())
            for (const auto &kv : entry.map) add(kv.first);
        else if (entry.is_seq())
            for (const Node &e : entry.seq)
                for (const auto &kv : e.map) add(kv.first);
        else add(entry.scalar);
    }
    return count;
}

void edge_attributes(Connection &edge, const Node &n) // description_parse_edge_attributes  src/yaml_snn.cpp:831-878
{
    if (n.is_seq())
    {
        for (const Node &e : n.seq) edge_attributes(edge, e);
        return;
    }
    if (const Node *s = n.find("synapse"))
        for (auto &[key, a] : model_attributes(*s))
        {
            Attr v = a;
            v.forward_to_dendrite = false;
            v.forward_to_soma = false;
            edge.synapse_attributes[key] = v;
        }
    if (const Node *d = n.find("dendrite"))
        for (auto &[key, a] : model_attributes(*d))
        {
            Attr v = a;
            v.forward_to_synapse = false;
            v.forward_to_soma = false;
            edge.dendrite_attributes[key] = v;
        }
    for (auto &[key, a] : model_attributes(n))
    {
        edge.synapse_attributes[key] = a;
        edge.dendrite_attributes[key] = a;
    }
}

std::string trim_ws(const std::string &s)
{
    size_t a = 0, b = s.size();
    while (a < b && std::isspace(static_cast<unsigned char>(s[a]))) ++a;
    while (b > a && std::isspace(static_cast<unsigned char>(s[b - 1]))) --b;
    return s.substr(a, b - a);
}

void parse_edge(const std::string &description, const Node &attrs, SpikingNetwork &net)
{
    // description_parse_edge_description  src/yaml_snn.cpp:396-448
    const size_t arrow = description.find("->");
    if (arrow == std::string::npos) parse_error("Edge is not formatted correctly: " + description, attrs);
    const std::string src = trim_ws(description.substr(0, arrow));
    const std::string dst = trim_ws(description.substr(arrow + 2));
    const size_t sdot = src.find('.'), ddot = dst.find('.');
    const bool s_neuron = sdot != std::string::npos, d_neuron = ddot != std::string::npos;
    if (s_neuron != d_neuron) parse_error("No target neuron defined in edge:" + description, attrs);
    const std::string sg = src.substr(0, sdot), dg = dst.substr(0, ddot);
    if (net.groups.find(sg) == net.groups.end()) parse_error("Invalid source neuron group:" + sg, attrs);
    if (net.groups.find(dg) == net.groups.end()) parse_error("Invalid target neuron group:" + dg, attrs);
    NeuronGroup &source_group = net.group(sg);
    NeuronGroup &target_group = net.group(dg);
    if (s_neuron)
    {
        const size_t so = std::stoull(src.substr(sdot + 1)), to = std::stoull(dst.substr(ddot + 1));
        if (so >= source_group.neurons.size()) parse_error("Invalid source neuron id: " + sg + "." + std::to_string(so), attrs);
        if (to >= target_group.neurons.size()) parse_error("Invalid target neuron id: " + dg + "." + std::to_string(to), attrs);
        Neuron &s = source_group.neurons[so];
        Connection &edge = s.edges_out[s.connect_to_neuron(target_group.neurons[to])];
        edge_attributes(edge, attrs);
        return;
    }
    // hyper-edge  src/yaml_snn.cpp:529-829
    std::string type;
    if (attrs.is_seq())
    {
        for (const Node &a : attrs.seq)
            if (const Node *t = a.find("type")) type = t->scalar;
    }
    else type = required<std::string>(attrs, "type");
    if (type.empty()) parse_error("No hyperedge type specified.", attrs);
    const AttrMap all = model_attributes(attrs);
    NeuronGroup::AttrLists lists;
    auto as_list = [&](const std::string &name, const Attr &a, const char *what) {
        if (!a.is_list()) parse_error(std::string("Attribute must be a list with an entry for each ") + what + " (name: " + name + ")", attrs);
        lists[name] = a.as_list();
    };
    if (type == "conv2d")
    {
        Conv2DParameters cv;
        for (const auto &[name, a] : all)
        {
            if (name == "input_height") cv.input_height = a.as_int();
            else if (name == "input_width") cv.input_width = a.as_int();
            else if (name == "input_channels") cv.input_channels = a.as_int();
            else if (name == "kernel_width") cv.kernel_width = a.as_int();
            else if (name == "kernel_height") cv.kernel_height = a.as_int();
            else if (name == "kernel_count") cv.kernel_count = a.as_int();
            else if (name == "stride_width") cv.stride_width = a.as_int();
            else if (name == "stride_height") cv.stride_height = a.as_int();
            else if (name != "type") as_list(name, a, "kernel connection");
        }
        source_group.connect_neurons_conv2d(target_group, lists, cv);
    }
    else if (type == "dense")
    {
        for (const auto &[name, a] : all)
            if (name != "type") as_list(name, a, "connection");
        source_group.connect_neurons_dense(target_group, lists);
    }
    else if (type == "sparse")
    {
        std::vector<std::pair<size_t, size_t>> pairs;
        for (const auto &[name, a] : all)
        {
            if (name == "source_target_pairs")
            {
                if (!a.is_list()) parse_error("Source/target pair must be a list of pairs", attrs);
                for (const Attr &p : a.as_list())
                {
                    if (!p.is_list() || p.as_list().size() != 2) parse_error("Invalid source/target format: expected [source, target]", attrs);
                    pairs.emplace_back(static_cast<size_t>(p.as_list()[0].as_int()), static_cast<size_t>(p.as_list()[1].as_int()));
                }
            }
            else if (name != "type") as_list(name, a, "connection pair");
        }
        source_group.connect_neurons_sparse(target_group, lists, pairs);
    }
    else parse_error("Invalid hyperedge type: " + type, attrs);
}

void mapping_info(const Node &info, Neuron &n, std::string &core_name) // src/yaml_snn.cpp:988-1028
{
    if (info.is_seq())
    {
        for (const Node &f : info.seq) mapping_info(f, n, core_name);
        return;
    }
    if (!info.is_map()) parse_error("Expected attributes to be map", info);
    if (const Node *p = info.find("synapse")) n.default_synapse_hw_name = p->scalar;
    if (const Node *p = info.find("dendrite")) n.dendrite_hw_name = p->scalar;
    if (const Node *p = info.find("soma")) n.soma_hw_name = p->scalar;
    if (const Node *p = info.find("core")) core_name = p->scalar;
}
} // namespace

std::unique_ptr<Architecture> load_arch_yaml(const std::string &path)
{
    Node doc;
    try
    {
        doc = yaml::parse_file(path);
    }
    catch (const std::invalid_argument &e)
    {
        if (std::string(e.what()).rfind("failed to open", 0) == 0)
            throw std::invalid_argument("Failed to open architecture file: " + path); // src/arch.cpp:106-117
        throw;
    }
    const Node *arch_node = doc.find("architecture");
    if (arch_node == nullptr) throw std::invalid_argument("Error: No architecture section defined (Line 1).\n");
    const std::string name = required<std::string>(*arch_node, "name");
    if (name.find('[') != std::string::npos) parse_error("Multiple architectures not supported", *arch_node);
    const Node &na = required_node(*arch_node, "attributes");
    // description_parse_noc_configuration_yaml  src/yaml_arch.cpp:425-510
    NetworkOnChipConfiguration noc;
    noc.width_in_tiles = static_cast<size_t>(required<int>(na, "width"));
    noc.height_in_tiles = static_cast<size_t>(required<int>(na, "height"));
    noc.link_buffer_size = static_cast<size_t>(required<int>(na, "link_buffer_size"));
    std::string model = "fixed";
    if (const Node *m = na.find("sync_model")) model = m->scalar;
    const Node *delay = na.find("latency_sync");
    auto as_double = [&](const Node &n) {
        double d = 0.0;
        if (!n.is_scalar() || !parse_double(n.scalar, d)) parse_error("Expected a number", n);
        return d;
    };
    if (model == "fixed") noc.ts_sync_delay_table.values[0] = delay != nullptr ? as_double(*delay) : 0.0;
    else if (model == "table")
    {
        if (delay == nullptr) parse_error("Attribute 'latency_sync' required when 'table' synchronization model is chosen.", na);
        if (delay->is_seq())
        {
            size_t i = 0;
            for (const Node &v : delay->seq) noc.ts_sync_delay_table.values[i++] = as_double(v);
        }
        else if (delay->is_map())
            for (const auto &kv : delay->map) noc.ts_sync_delay_table.values[std::stoull(kv.first)] = as_double(kv.second);
        else noc.ts_sync_delay_table.values[0] = as_double(*delay);
    }
    else parse_error("Unknown sync_model: " + model, na);
    auto arch = std::make_unique<Architecture>(name, noc);
    const Node *tiles = arch_node->find("tile");
    if (tiles == nullptr) parse_error("No tile section defined", *arch_node);
    for_each_entry(*tiles, [&](const Node &tile) { parse_tile(tile, *arch); });
    return arch;
}

std::unique_ptr<SpikingNetwork> load_net_yaml(const std::string &path, Architecture &arch)
{
    Node doc;
    try
    {
        doc = yaml::parse_file(path);
    }
    catch (const std::invalid_argument &e)
    {
        if (std::string(e.what()).rfind("failed to open", 0) == 0)
            throw std::invalid_argument("Error: Network file: failed to open (" + path + ")."); // src/network.cpp:200-205
        throw;
    }
    if (!doc.is_map()) throw std::invalid_argument("Error: Mapped network file has invalid format (Line 1).\n");
    const Node *net_node = doc.find("network");
    if (net_node == nullptr) parse_error("No top-level 'network' section defined", doc);
    std::string net_name;
    if (const Node *nm = net_node->find("name"))
    {
        net_name = nm->scalar;
        if (net_name.find('[') != std::string::npos) parse_error("Multiple networks not supported", *net_node);
    }
    auto net = std::make_unique<SpikingNetwork>(net_name);
    const Node *groups = net_node->find("groups");
    if (groups == nullptr) parse_error("No neuron groups specified", *net_node);
    const Node *edges = net_node->find("edges");
    if (edges == nullptr) parse_error("No edges section specified", *net_node);
    if (!groups->is_seq()) parse_error("Neuron group section does not define a list of groups", *groups);
    for (const Node &g : groups->seq)
    {
        const std::string gname = required<std::string>(g, "name");
        const Node *neurons = g.find("neurons");
        if (neurons == nullptr) parse_error("No neurons section defined.", g);
        const size_t count = count_neurons(*neurons);
        NeuronConfiguration def;
        if (const Node *a = g.find("attributes")) def = neuron_attributes(*a, NeuronConfiguration{});
        NeuronGroup &group = net->create_neuron_group(gname, count, def);
        for (const Node &entry : neurons->seq)
        {
            auto apply = [&](const std::string &id, const Node &attrs) {
                const NeuronConfiguration cfg = neuron_attributes(attrs, group.default_neuron_config);
                if (has_range(id))
                {
                    const auto r = parse_range(id);
                    for (size_t i = r.first; i <= r.second; ++i) group.neurons.at(i).set_attributes(cfg);
                }
                else group.neurons.at(std::stoull(id)).set_attributes(cfg);
            };
            if (entry.is_map())
                for (const auto &kv : entry.map) apply(kv.first, kv.second);
            else if (entry.is_seq())
                for (const Node &e : entry.seq)
                    for (const auto &kv : e.map) apply(kv.first, kv.second);
            // a plain "a..b" entry only counts (src/yaml_snn.cpp:226-302)
        }
    }
    if (!edges->is_seq()) parse_error("Edges section does not define a list of edges", *edges);
    for (const Node &entry : edges->seq)
        for (const auto &kv : entry.map) parse_edge(kv.first, kv.second, *net);
    // mappings  src/yaml_snn.cpp:880-1056
    const Node *mappings = doc.find("mappings");
    if (mappings == nullptr) parse_error("No 'mappings' section defined", doc);
    if (!mappings->is_seq()) parse_error("Mappings must be given as a sequence / list.", *mappings);
    for (const Node &m : mappings->seq)
    {
        if (!m.is_map()) parse_error("Expected mapping to be defined in the format: <group>.<neuron>: [<attributes>]", m);
        if (m.map.size() > 1) parse_error("Should be one entry per mapping", m);
        for (const auto &kv : m.map)
        {
            const std::string &addr = kv.first;
            const size_t dot = addr.find('.');
            const std::string gname = addr.substr(0, dot);
            if (net->groups.find(gname) == net->groups.end()) parse_error("While mapping, group not found (" + gname + ")", kv.second);
            NeuronGroup &group = net->group(gname);
            size_t first = 0, last = group.neurons.empty() ? 0 : group.neurons.size() - 1;
            if (dot != std::string::npos)
            {
                const std::string ns = addr.substr(dot + 1);
                if (has_range(ns)) std::tie(first, last) = parse_range(ns);
                else first = last = std::stoull(ns);
            }
            for (size_t o = first; o <= last; ++o)
            {
                if (o >= group.neurons.size()) parse_error("Invalid neuron id: " + gname + "." + std::to_string(o), kv.second);
                Neuron &n = group.neurons[o];
                std::string core_address;
                mapping_info(kv.second, n, core_address);
                const size_t cdot = core_address.find('.');
                const size_t tile_id = std::stoull(core_address.substr(0, cdot));
                const size_t core_off = std::stoull(core_address.substr(cdot + 1));
                if (tile_id >= arch.tiles.size()) parse_error("Tile ID >= tile count", kv.second);
                if (core_off >= arch.tiles[tile_id].cores.size()) parse_error("Core ID >= core count", kv.second);
                n.map_to_core(arch.tiles[tile_id].cores[core_off]);
            }
        }
    }
    return net;
}


// ---------------------------------------------------------------------------------------------
// Legacy netlist front-end (`sim -n`, sanafe.load_net(path, arch, use_netlist_format=True)):
// one entry per line, `g <count> <attrs>` / `n <group>.<neuron> <attrs>` / `e <g>.<n>-><g>.<n> <attrs>`
// / `& <g>.<n>@<tile>.<core>`; attributes are whitespace-separated key=value fields or one embedded
// YAML flow collection. Restates src/netlist.cpp:38-617 (reader only).
// ---------------------------------------------------------------------------------------------
namespace
{
std::vector<std::string> netlist_fields(const std::string &line) // netlist_get_fields  src/netlist.cpp:71-94
{
    std::vector<std::string> fields;
    const char *delim = " \t\r\n";
    size_t start = line.find_first_not_of(delim);
    while (start != std::string::npos)
    {
        const size_t end = line.find_first_of(delim, start);
        fields.push_back(line.substr(start, end == std::string::npos ? std::string::npos : end - start));
        start = end == std::string::npos ? std::string::npos : line.find_first_not_of(delim, end);
    }
    return fields;
}

size_t netlist_to_index(const std::string &s) // field_to_int  src/netlist.cpp (std::from_chars on the whole field)
{
    size_t pos = 0;
    unsigned long long v = 0;
    try
    {
        v = std::stoull(s, &pos);
    }
    catch (const std::exception &)
    {
        pos = 0;
    }
    if (pos != s.size() || s.empty()) throw std::runtime_error("Error: Invalid integer field in netlist (" + s + ")");
    return static_cast<size_t>(v);
}

// netlist_parse_attribute_value  src/netlist.cpp:286-320: int, then double, then bool (0/1 only,
// i.e. never reached), then string - each must consume the whole value
Attr netlist_value(const std::string &text)
{
    {
        std::stringstream ss(text);
        int v = 0;
        if ((ss >> v) && ss.eof()) return Attr::of(v);
    }
    {
        std::stringstream ss(text);
        double v = 0.0;
        if ((ss >> v) && ss.eof()) return Attr::of(v);
    }
    return Attr::of(text);
}

AttrMap netlist_attributes(const std::vector<std::string> &fields, const size_t first, const int line_number)
{
    AttrMap out;
    if (first >= fields.size() || fields[first].empty()) return out;
    const char open = fields[first][0];
    if (open == '[' || open == '{')
    {
        // netlist_parse_embedded_json  src/netlist.cpp:378-414: the fields are glued back together and
        // parsed as one YAML flow collection up to its closing bracket
        std::string all;
        for (size_t k = first; k < fields.size(); ++k) all += fields[k] + ' ';
        const char close = open == '[' ? ']' : '}';
        int depth = 0;
        size_t end = std::string::npos;
        for (size_t k = 0; k < all.size(); ++k)
        {
            if (all[k] == open) ++depth;
            else if (all[k] == close && --depth == 0)
            {
                end = k;
                break;
            }
        }
        if (end == std::string::npos)
            throw std::runtime_error("Error: Line " + std::to_string(line_number) + ": embedded attributes are not terminated");
        return model_attributes(yaml::parse(all.substr(0, end + 1)));
    }
    for (size_t k = first; k < fields.size(); ++k)
    {
        const std::string &field = fields[k];
        const size_t eq = field.find('=');
        if (field.size() < 3 || eq == std::string::npos || eq == 0 || eq + 1 >= field.size()) continue; // logged and skipped
        Attr a = netlist_value(field.substr(eq + 1));
        a.name = field.substr(0, eq);
        out.insert({field.substr(0, eq), a}); // the first occurrence of a key wins (std::map::insert)
    }
    return out;
}

std::pair<std::string, size_t> netlist_neuron(const std::string &field) // netlist_parse_neuron_field  :96-110
{
    const size_t dot = field.find('.');
    if (dot == std::string::npos) throw std::runtime_error("Error: Invalid neuron format");
    return {field.substr(0, dot), netlist_to_index(field.substr(dot + 1))};
}

NeuronConfiguration netlist_neuron_config(const AttrMap &attrs) // the reserved names  :421-446, 491-515
{
    NeuronConfiguration cfg;
    auto get = [&](const char *key) -> const Attr * {
        auto it = attrs.find(key);
        return it == attrs.end() ? nullptr : &it->second;
    };
    if (const Attr *a = get("synapse_hw_name")) cfg.default_synapse_hw_name = a->as_string();
    if (const Attr *a = get("dendrite_hw_name")) cfg.dendrite_hw_name = a->as_string();
    if (const Attr *a = get("soma_hw_name")) cfg.soma_hw_name = a->as_string();
    if (const Attr *a = get("log_spikes")) cfg.log_spikes = a->as_bool();
    if (const Attr *a = get("log_v")) cfg.log_potential = a->as_bool();
    return cfg;
}

Neuron &netlist_lookup(SpikingNetwork &net, const std::string &group, const size_t id, const int line_number)
{
    if (net.groups.find(group) == net.groups.end())
        throw std::invalid_argument("Error: Line " + std::to_string(line_number) + ": Group (" + group + ") not in groups.");
    NeuronGroup &g = net.group(group);
    if (id >= g.neurons.size())
        throw std::invalid_argument("Error: Line " + std::to_string(line_number) + ": Trying to access neuron (" + group + "." +
                std::to_string(id) + ") but group " + group + " only allocates " + std::to_string(g.neurons.size()) +
                " neuron(s).");
    return g.neurons[id];
}
} // namespace

std::unique_ptr<SpikingNetwork> load_net_netlist(const std::string &path, Architecture &arch)
{
    std::ifstream fp(path);
    if (!fp.is_open()) throw std::invalid_argument("Error: Network file: failed to open (" + path + ")."); // src/network.cpp:200-205
    auto net = std::make_unique<SpikingNetwork>("");
    std::string line;
    int line_number = 1;
    for (; std::getline(fp, line); ++line_number)
    {
        const std::vector<std::string> fields = netlist_fields(line);
        if (fields.empty()) continue;
        const char type = fields[0][0];
        if (type == '#') continue; // comment (netlist_read_network_entry  src/netlist.cpp:186-194)
        if (fields.size() < 2) throw std::invalid_argument("Error: Line " + std::to_string(line_number) + ": fields < 2");
        if (type == 'g')
        {
            // groups are named by their position  :416-465
            AttrMap attrs = netlist_attributes(fields, 2, line_number);
            NeuronConfiguration cfg = netlist_neuron_config(attrs);
            for (auto it = attrs.begin(); it != attrs.end();)
                it = is_reserved_neuron_attribute(it->first) ? attrs.erase(it) : std::next(it);
            cfg.model_attributes = attrs;
            net->create_neuron_group(std::to_string(net->groups.size()), netlist_to_index(fields[1]), cfg);
        }
        else if (type == 'n')
        {
            // per-neuron attributes; the reserved names stay in model_attributes here (:467-517), so a
            // netlist that names e.g. soma_hw_name per neuron fails at load like it does in the reference
            const auto [group, id] = netlist_neuron(fields[1]);
            Neuron &n = netlist_lookup(*net, group, id, line_number);
            const AttrMap attrs = netlist_attributes(fields, 2, line_number);
            NeuronConfiguration cfg = netlist_neuron_config(attrs);
            cfg.model_attributes = attrs;
            n.set_attributes(cfg);
        }
        else if (type == 'e')
        {
            const size_t arrow = fields[1].find("->");
            if (arrow == std::string::npos) throw std::runtime_error("Invalid edge format");
            const auto [sg, sn] = netlist_neuron(fields[1].substr(0, arrow));
            const auto [dg, dn] = netlist_neuron(fields[1].substr(arrow + 2));
            Neuron &src = netlist_lookup(*net, sg, sn, line_number);
            Neuron &dst = netlist_lookup(*net, dg, dn, line_number);
            const AttrMap attrs = netlist_attributes(fields, 2, line_number);
            const size_t idx = src.connect_to_neuron(dst);
            src.edges_out[idx].synapse_attributes = attrs;
            src.edges_out[idx].dendrite_attributes = attrs;
        }
        else if (type == '&')
        {
            const size_t at = fields[1].find('@');
            if (at == std::string::npos) throw std::runtime_error("Invalid mapping format");
            const auto [group, id] = netlist_neuron(fields[1].substr(0, at));
            const std::string core = fields[1].substr(at + 1);
            const size_t dot = core.find('.');
            if (dot == std::string::npos) throw std::runtime_error("Error: Invalid neuron format");
            const size_t tile_id = netlist_to_index(core.substr(0, dot)), core_off = netlist_to_index(core.substr(dot + 1));
            if (tile_id >= arch.tiles.size() || core_off >= arch.tiles[tile_id].cores.size())
                throw std::runtime_error("Error: Couldn't parse mapping.");
            netlist_lookup(*net, group, id, line_number).map_to_core(arch.tiles[tile_id].cores[core_off]);
        }
        else
            throw std::invalid_argument("Invalid description entry type");
    }
    return net;
}

} // namespace sfe
