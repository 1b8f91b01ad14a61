// yaml_writer.cpp — Network.save(path): writes a SpikingNetwork as the SNN YAML format.
//
// Same document layout as the reference's writer (src/yaml_snn.cpp:1058-1416 network section,
// :1418-1555 mappings section): `network: {name, groups: [{name, attributes, neurons: [run: {...}]}],
// edges: [{"g.i -> h.j": {attrs}}]}` followed by `mappings: [{"g.i": {core: "t.c", synapse, dendrite,
// soma}}]` in mapping order; neurons with identical settings are written as one `a..b` run; attributes
// that are forwarded to some hardware units only go under `synapse:` / `dendrite:` / `soma:`.
// Differences, on purpose: a neuron attribute is omitted when it EQUALS the group default (the
// reference's test is inverted, src/yaml_snn.cpp:1320-1326, and drops exactly the overrides), doubles
// are written with 17 significant digits, and the file is rewritten rather than merged with previous
// content. What is written loads back (here and in the reference) to a network that lowers to the same
// tables: tests/test_python_builders.py.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <functional>
#include <sstream>

#include "handles.hpp"
#include "yaml.hpp"

namespace sfe
{
namespace
{
bool same_attr(const Attr &a, const Attr &b)
{
    if (a.forward_to_synapse != b.forward_to_synapse || a.forward_to_dendrite != b.forward_to_dendrite ||
            a.forward_to_soma != b.forward_to_soma || a.value.index() != b.value.index())
        return false;
    if (a.is_list())
    {
        const std::vector<Attr> &x = a.as_list(), &y = b.as_list();
        if (x.size() != y.size()) return false;
        for (size_t i = 0; i < x.size(); ++i)
            if (x[i].name != y[i].name || !same_attr(x[i], y[i])) return false;
        return true;
    }
    if (const bool *v = std::get_if<bool>(&a.value)) return *v == std::get<bool>(b.value);
    if (const int *v = std::get_if<int>(&a.value)) return *v == std::get<int>(b.value);
    if (const double *v = std::get_if<double>(&a.value)) return *v == std::get<double>(b.value);
    return a.as_string() == b.as_string();
}

bool same_attrs(const AttrMap &a, const AttrMap &b)
{
    if (a.size() != b.size()) return false;
    auto ia = a.begin();
    auto ib = b.begin();
    for (; ia != a.end(); ++ia, ++ib)
        if (ia->first != ib->first || !same_attr(ia->second, ib->second)) return false;
    return true;
}

std::string scalar(const std::string &s) // a string that must read back as this string
{
    bool plain = !s.empty() && s.front() != ' ' && s.back() != ' ';
    for (const char ch : s)
        if (std::string("[]{},:#&*!|>'\"%@`\n\t").find(ch) != std::string::npos) plain = false;
    if (plain)
    {
        // would it be typed as a number or a bool by the reader (src/yaml_common.cpp:205-263)?
        char *end = nullptr;
        std::strtod(s.c_str(), &end);
        if (end != nullptr && *end == '\0') plain = false;
        if (s == "true" || s == "false" || s == "True" || s == "False" || s == "null" || s == "~" || s == "-") plain = false;
    }
    if (plain) return s;
    std::string out = "\"";
    for (const char ch : s)
    {
        if (ch == '"' || ch == '\\') out += '\\';
        if (ch == '\n') out += "\\n";
        else if (ch == '\t') out += "\\t";
        else out += ch;
    }
    return out + "\"";
}

std::string number(const double v)
{
    char buf[64];
    std::snprintf(buf, sizeof(buf), "%.17g", v);
    std::string s = buf;
    // keep the type: a double must not read back as an int
    if (s.find_first_of(".eEn") == std::string::npos) s += ".0";
    return s;
}

void write_value(std::ostream &out, const Attr &a) // flow style
{
    if (const bool *b = std::get_if<bool>(&a.value)) out << (*b ? "true" : "false");
    else if (const int *i = std::get_if<int>(&a.value)) out << *i;
    else if (const double *d = std::get_if<double>(&a.value)) out << number(*d);
    else if (const std::string *s = std::get_if<std::string>(&a.value)) out << scalar(*s);
    else
    {
        const std::vector<Attr> &list = a.as_list();
        out << "[";
        for (size_t k = 0; k < list.size(); ++k)
        {
            if (k != 0) out << ", ";
            if (list[k].name.has_value() && !list[k].name->empty())
            {
                out << "{" << scalar(*list[k].name) << ": ";
                write_value(out, list[k]);
                out << "}";
            }
            else write_value(out, list[k]);
        }
        out << "]";
    }
}

// "key: value" items of a flow mapping; unit-specific attributes nest under synapse/dendrite/soma
std::vector<std::string> flow_items(const AttrMap &attrs, const AttrMap *defaults)
{
    std::vector<std::string> items;
    std::map<std::string, std::vector<std::string>> sections;
    for (const auto &[key, a] : attrs)
    {
        if (defaults != nullptr)
        {
            const auto d = defaults->find(key);
            if (d != defaults->end() && same_attr(d->second, a)) continue;
        }
        std::ostringstream v;
        v << scalar(key) << ": ";
        write_value(v, a);
        if (a.forward_to_synapse && a.forward_to_dendrite && a.forward_to_soma) items.push_back(v.str());
        else
        {
            if (a.forward_to_synapse) sections["synapse"].push_back(v.str());
            if (a.forward_to_dendrite) sections["dendrite"].push_back(v.str());
            if (a.forward_to_soma) sections["soma"].push_back(v.str());
        }
    }
    for (const auto &[name, sub] : sections)
    {
        std::string s = name + ": {";
        for (size_t k = 0; k < sub.size(); ++k) s += (k != 0 ? ", " : "") + sub[k];
        items.push_back(s + "}");
    }
    return items;
}

std::string flow_map(const std::vector<std::string> &items)
{
    std::string s = "{";
    for (size_t k = 0; k < items.size(); ++k) s += (k != 0 ? ", " : "") + items[k];
    return s + "}";
}

std::string address(const NeuronAddress &a)
{
    return a.neuron_offset.has_value() ? a.group_name + "." + std::to_string(*a.neuron_offset) : a.group_name;
}
} // namespace

void save_net_yaml(const SpikingNetwork &net, const std::string &path)
{
    std::ostringstream out;
    out << "network:\n  name: " << (net.name.empty() ? std::string("\" \"") : scalar(net.name)) << "\n  groups:\n";
    for (const auto &[gname, gptr] : net.groups)
    {
        const NeuronGroup &g = *gptr;
        const NeuronConfiguration &def = g.default_neuron_config;
        std::vector<std::string> items = flow_items(def.model_attributes, nullptr);
        if (def.log_spikes.has_value()) items.push_back(std::string("log_spikes: ") + (*def.log_spikes ? "true" : "false"));
        if (def.log_potential.has_value()) items.push_back(std::string("log_potential: ") + (*def.log_potential ? "true" : "false"));
        out << "  - name: " << scalar(g.name) << "\n    attributes: " << flow_map(items) << "\n    neurons:\n";
        const bool def_spikes = def.log_spikes.value_or(false), def_pot = def.log_potential.value_or(false);
        size_t run_start = 0;
        for (size_t i = 1; i <= g.neurons.size(); ++i)
        {
            if (i < g.neurons.size())
            {
                const Neuron &a = g.neurons[run_start], &b = g.neurons[i];
                if (same_attrs(a.model_attributes, b.model_attributes) && a.log_spikes == b.log_spikes &&
                        a.log_potential == b.log_potential)
                    continue;
            }
            const Neuron &n = g.neurons[run_start];
            std::vector<std::string> nitems = flow_items(n.model_attributes, &def.model_attributes);
            if (n.log_spikes != def_spikes) nitems.push_back(std::string("log_spikes: ") + (n.log_spikes ? "true" : "false"));
            if (n.log_potential != def_pot) nitems.push_back(std::string("log_potential: ") + (n.log_potential ? "true" : "false"));
            out << "    - {" << run_start;
            if (i - 1 > run_start) out << ".." << (i - 1);
            out << ": " << flow_map(nitems) << "}\n";
            run_start = i;
        }
    }
    size_t n_edges = 0;
    for (const auto &[gname, gptr] : net.groups)
        for (const Neuron &n : gptr->neurons) n_edges += n.edges_out.size();
    out << (n_edges == 0 ? "  edges: []\n" : "  edges:\n"); // (an empty block would read back as null, not as a list)
    for (const auto &[gname, gptr] : net.groups)
        for (const Neuron &n : gptr->neurons)
            for (const Connection &con : n.edges_out)
                out << "  - {" << scalar(address(con.pre_neuron) + " -> " + address(con.post_neuron)) << ": "
                    << flow_map(flow_items(con.synapse_attributes, nullptr)) << "}\n";
    // mappings, in mapping order (src/yaml_snn.cpp:1477-1490)
    std::vector<const Neuron *> all;
    for (const auto &[gname, gptr] : net.groups)
        for (const Neuron &n : gptr->neurons) all.push_back(&n);
    std::stable_sort(all.begin(), all.end(), [](const Neuron *a, const Neuron *b) { return a->mapping_order < b->mapping_order; });
    out << "mappings:\n";
    for (const Neuron *n : all)
    {
        if (!n->core_address.has_value()) throw std::runtime_error("Error: Neuron not mapped, can't save."); // :1509-1514
        out << "- {" << scalar(n->parent_group_name + "." + std::to_string(n->offset)) << ": {core: \""
            << n->core_address->parent_tile_id << "." << n->core_address->offset_within_tile << "\"";
        if (!n->default_synapse_hw_name.empty()) out << ", synapse: " << scalar(n->default_synapse_hw_name);
        if (!n->dendrite_hw_name.empty()) out << ", dendrite: " << scalar(n->dendrite_hw_name);
        if (!n->soma_hw_name.empty()) out << ", soma: " << scalar(n->soma_hw_name);
        out << "}}\n";
    }
    // An existing file keeps its other top-level sections; `network` and `mappings` are replaced
    // (yaml_write_network / yaml_write_mappings_file re-open the file as a tree, src/yaml_snn.cpp:1058-1110,1477-1555;
    // a file that does not parse is an error). The kept sections are copied as text, ahead of the two written here.
    std::string kept;
    {
        std::ifstream old(path);
        if (old.is_open())
        {
            std::ostringstream buf;
            buf << old.rdbuf();
            const std::string text = buf.str();
            bool blank = true;
            for (const char ch : text) blank = blank && (ch == ' ' || ch == '\n' || ch == '\r' || ch == '\t');
            if (!blank)
            {
                yaml::Node doc;
                try
                {
                    doc = yaml::parse(text);
                }
                catch (const std::exception &e)
                {
                    throw std::runtime_error("Error: existing file is not valid YAML, cannot save network into it (" + path + "): " + e.what());
                }
                if (!doc.is_map()) throw std::runtime_error("Error: existing file is not a YAML mapping, cannot save network into it (" + path + ")");
                // top-level blocks = runs of lines from one column-0 key to the next
                std::istringstream lines(text);
                bool keep = false;
                for (std::string line; std::getline(lines, line);)
                {
                    const bool key_line = !line.empty() && line[0] != ' ' && line[0] != '\t' && line[0] != '#' && line[0] != '-' &&
                            line.find(':') != std::string::npos;
                    if (key_line)
                    {
                        const std::string key = line.substr(0, line.find(':'));
                        keep = key != "network" && key != "mappings";
                    }
                    if (keep) kept += line + "\n";
                }
            }
        }
    }
    std::ofstream fp(path);
    if (!fp.is_open()) throw std::runtime_error("Failed to open YAML file for writing: " + path);
    fp << kept << out.str();
}

// ---------------------------------------------------------------------------
// Legacy netlist format: SpikingNetwork::save(path, use_netlist_format = true)
// (src/network.cpp:606-703; line builders src/netlist.cpp:619-851). Same lines as the reference writes, with the
// reference's losses: groups are written by position (the loader names them "0", "1", ...), a group's log_potential is
// written under a key the loader does not read (`log_v` is what it reads, src/netlist.cpp:437-441), numbers print
// through ModelAttribute::print (src/attribute.hpp:129-155: doubles as `std::scientific`, 6 digits).
// ---------------------------------------------------------------------------
namespace
{
std::string netlist_print(const Attr &a) // ModelAttribute::print
{
    if (const bool *b = std::get_if<bool>(&a.value)) return *b ? "true" : "false";
    if (const int *i = std::get_if<int>(&a.value)) return std::to_string(*i);
    if (const double *d = std::get_if<double>(&a.value))
    {
        std::ostringstream ss;
        ss << std::scientific << *d;
        return ss.str();
    }
    if (const std::string *str = std::get_if<std::string>(&a.value)) return *str;
    throw std::runtime_error("Printing vectors not yet supported");
}

// netlist_attributes_to_netlist  src/netlist.cpp:774-851: ` key=value` per attribute that differs from the group's
// default; as soon as one attribute is a list, all of them go into one YAML flow mapping instead
std::string netlist_attributes(const AttrMap &attrs, const AttrMap &defaults)
{
    bool nested = false;
    for (const auto &[key, a] : attrs) nested = nested || a.is_list();
    if (nested)
    {
        const std::vector<std::string> items = flow_items(attrs, &defaults);
        return " " + flow_map(items);
    }
    std::string out;
    for (const auto &[key, a] : attrs)
    {
        const auto d = defaults.find(key);
        if (d != defaults.end() && same_attr(d->second, a)) continue;
        out += " " + key + "=" + netlist_print(a);
    }
    return out;
}
} // namespace

void save_net_netlist(const SpikingNetwork &net, const std::string &path)
{
    std::ofstream out(path);
    if (!out.is_open()) throw std::invalid_argument("Error: Couldn't open net file to save to.");
    // create_group_name_to_id_mapping: position in the (lexicographic) group map
    std::map<std::string, size_t> group_id;
    for (const auto &[name, g] : net.groups) group_id.emplace(name, group_id.size());
    const AttrMap no_defaults;
    // save_groups_to_netlist / netlist_group_to_netlist
    for (const auto &[name, g] : net.groups)
    {
        const NeuronConfiguration &c = g->default_neuron_config;
        std::string line = "g " + std::to_string(g->neurons.size());
        if (c.default_synapse_hw_name.has_value() && !c.default_synapse_hw_name->empty()) line += " synapse_hw_name=" + *c.default_synapse_hw_name;
        if (c.dendrite_hw_name.has_value() && !c.dendrite_hw_name->empty()) line += " dendrite_hw_name=" + *c.dendrite_hw_name;
        if (c.log_potential.value_or(false)) line += " log_potential=1";
        if (c.log_spikes.value_or(false)) line += " log_spikes=1";
        if (c.soma_hw_name.has_value() && !c.soma_hw_name->empty()) line += " soma_hw_name=" + *c.soma_hw_name;
        out << line << netlist_attributes(c.model_attributes, no_defaults) << "\n";
    }
    // save_neurons_to_netlist: every neuron followed by its outgoing connections
    for (const auto &[name, g] : net.groups)
    {
        const NeuronConfiguration &c = g->default_neuron_config;
        for (const Neuron &n : g->neurons)
        {
            std::string line = "n " + std::to_string(group_id.at(name)) + "." + std::to_string(n.offset);
            // (is_unique_attribute: written when the group has no default or a different one)
            auto unique_name = [&](const char *key, const std::string &value, const std::optional<std::string> &dflt) {
                if (!value.empty() && (!dflt.has_value() || *dflt != value)) line += std::string(" ") + key + "=" + value;
            };
            unique_name("soma_hw_name", n.soma_hw_name, c.soma_hw_name);
            unique_name("synapse_hw_name", n.default_synapse_hw_name, c.default_synapse_hw_name);
            unique_name("dendrite_hw_name", n.dendrite_hw_name, c.dendrite_hw_name);
            if (n.log_spikes && (!c.log_spikes.has_value() || !*c.log_spikes)) line += " log_spikes=1";
            if (n.log_potential && (!c.log_potential.has_value() || !*c.log_potential)) line += " log_potential=1";
            out << line << netlist_attributes(n.model_attributes, c.model_attributes) << "\n";
            for (const Connection &con : n.edges_out)
            {
                auto end = [&](const NeuronAddress &a) {
                    return std::to_string(group_id.at(a.group_name)) + (a.neuron_offset.has_value() ? "." + std::to_string(*a.neuron_offset) : "");
                };
                out << "e " << end(con.pre_neuron) << "->" << end(con.post_neuron)
                    << netlist_attributes(con.synapse_attributes, no_defaults) << "\n";
            }
        }
    }
    // save_mappings_to_netlist: in mapping order
    std::vector<const Neuron *> all;
    for (const auto &[name, g] : net.groups)
        for (const Neuron &n : g->neurons) all.push_back(&n);
    std::stable_sort(all.begin(), all.end(), [](const Neuron *a, const Neuron *b) { return a->mapping_order < b->mapping_order; });
    for (const Neuron *n : all)
    {
        if (n->core_address.has_value())
            out << "& " << group_id.at(n->parent_group_name) << "." << n->offset << "@" << n->core_address->parent_tile_id << "."
                << n->core_address->offset_within_tile;
        out << "\n";
    }
}

} // namespace sfe
