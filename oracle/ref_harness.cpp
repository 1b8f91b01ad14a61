// ref_harness.cpp — drives the UNMODIFIED reference engine (compiled from
// /root/reference/src, see oracle/Makefile) as the parity oracle and as the CPU
// baseline. TEST INFRASTRUCTURE ONLY: nothing in the product links or calls this.
//
// The reference's YAML front-end needs RapidYAML (network-fetched; absent
// offline), so descriptions arrive in a "flat" JSON-lines form produced by
// oracle/yaml_to_flat.py (PyYAML -> flat). This file replays those records
// through the reference's public builder API (Architecture::create_tile /
// create_core, CoreConfiguration::create_*, SpikingNetwork::create_neuron_group,
// Neuron::set_attributes / connect_to_neuron / map_to_core,
// NeuronGroup::connect_neurons_*). Every data structure and all arithmetic from
// load() onwards are the reference's own.
//
// usage: sanafe_ref <flat.jsonl> --steps N [--timing simple|detailed]
//                   [--threads N] [--out DIR] [--traces] [--per-step]
//                   [--dump-map] [--reps R] [--hash-steps H]
// --hash-steps H: before the timed calls, H timesteps are run one by one and the spikes of each (get_spikes, i.e.
// neurons with log_spikes) are folded into an order-independent 64-bit raster hash, sum of
// sfe_mix64(timestep << 32 | neuron offset) mod 2^64, written to summary.json with the counters of those H steps:
// bench.py runs the GPU engine on the same network and steps and compares (parity at the benchmark's own shape).
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <variant>
#include <vector>

#ifdef HAVE_OPENMP
#include <omp.h>
#endif

#include "arch.hpp"
#include "attribute.hpp"
#include "chip.hpp"
#include "core.hpp"
#include "mapped.hpp"
#include "models.hpp"
#include "netlist.hpp"
#include "network.hpp"
#include "pipeline.hpp"
#include "tile.hpp"

#include "sfe_synth.h"

// ---------------------------------------------------------------------------
// Entry points the engine TUs reference but whose definitions live in the
// RapidYAML-dependent files we cannot build (see stub/shim.hpp).
// ---------------------------------------------------------------------------
namespace sanafe
{
Architecture description_parse_arch_file_yaml(std::ifstream &)
{
    throw std::runtime_error("YAML front-end not built into the oracle");
}
SpikingNetwork yaml_parse_network_file(std::ifstream &, Architecture &)
{
    throw std::runtime_error("YAML front-end not built into the oracle");
}
void yaml_write_network(std::filesystem::path, const SpikingNetwork &)
{
    throw std::runtime_error("YAML writer not built into the oracle");
}
void yaml_write_mappings_file(std::filesystem::path, const SpikingNetwork &)
{
    throw std::runtime_error("YAML writer not built into the oracle");
}
SpikingNetwork netlist_parse_file(std::ifstream &, Architecture &)
{
    throw std::runtime_error("netlist front-end not built into the oracle");
}
std::string netlist_group_to_netlist(const NeuronGroup &)
{
    return {};
}
std::string netlist_neuron_to_netlist(const Neuron &, const SpikingNetwork &,
        const std::map<std::string, size_t> &)
{
    return {};
}
std::string netlist_mapping_to_netlist(
        const Neuron &, const std::map<std::string, size_t> &)
{
    return {};
}
std::string netlist_connection_to_netlist(
        const Connection &, const std::map<std::string, size_t> &)
{
    return {};
}
}

// ---------------------------------------------------------------------------
// Minimal JSON reader (objects keep insertion order)
// ---------------------------------------------------------------------------
struct JVal;
using JList = std::vector<JVal>;
using JMap = std::vector<std::pair<std::string, JVal>>;
struct JVal
{
    std::variant<std::monostate, bool, long long, double, std::string,
            std::shared_ptr<JList>, std::shared_ptr<JMap>>
            v;
    bool is_null() const { return std::holds_alternative<std::monostate>(v); }
    bool is_bool() const { return std::holds_alternative<bool>(v); }
    bool is_int() const { return std::holds_alternative<long long>(v); }
    bool is_double() const { return std::holds_alternative<double>(v); }
    bool is_str() const { return std::holds_alternative<std::string>(v); }
    bool is_list() const { return std::holds_alternative<std::shared_ptr<JList>>(v); }
    bool is_map() const { return std::holds_alternative<std::shared_ptr<JMap>>(v); }
    bool b() const { return std::get<bool>(v); }
    long long i() const { return std::get<long long>(v); }
    double d() const { return is_int() ? static_cast<double>(i()) : std::get<double>(v); }
    const std::string &s() const { return std::get<std::string>(v); }
    const JList &list() const { return *std::get<std::shared_ptr<JList>>(v); }
    const JMap &map() const { return *std::get<std::shared_ptr<JMap>>(v); }
    const JVal *find(const std::string &k) const
    {
        if (!is_map()) return nullptr;
        for (const auto &kv : map())
            if (kv.first == k) return &kv.second;
        return nullptr;
    }
    const JVal &at(const std::string &k) const
    {
        const JVal *p = find(k);
        if (p == nullptr) throw std::runtime_error("flat: missing key " + k);
        return *p;
    }
};

struct JParser
{
    const char *p;
    const char *end;
    void ws()
    {
        while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p;
    }
    JVal parse()
    {
        ws();
        if (p >= end) throw std::runtime_error("json: unexpected end");
        JVal out;
        if (*p == '{')
        {
            ++p;
            auto m = std::make_shared<JMap>();
            ws();
            if (*p == '}') { ++p; out.v = m; return out; }
            for (;;)
            {
                ws();
                JVal k = parse();
                ws();
                if (*p != ':') throw std::runtime_error("json: expected ':'");
                ++p;
                JVal val = parse();
                m->emplace_back(k.s(), std::move(val));
                ws();
                if (*p == ',') { ++p; continue; }
                if (*p == '}') { ++p; break; }
                throw std::runtime_error("json: expected ',' or '}'");
            }
            out.v = m;
            return out;
        }
        if (*p == '[')
        {
            ++p;
            auto l = std::make_shared<JList>();
            ws();
            if (*p == ']') { ++p; out.v = l; return out; }
            for (;;)
            {
                l->push_back(parse());
                ws();
                if (*p == ',') { ++p; continue; }
                if (*p == ']') { ++p; break; }
                throw std::runtime_error("json: expected ',' or ']'");
            }
            out.v = l;
            return out;
        }
        if (*p == '"')
        {
            ++p;
            std::string s;
            while (p < end && *p != '"')
            {
                if (*p == '\\')
                {
                    ++p;
                    switch (*p)
                    {
                    case 'n': s += '\n'; break;
                    case 't': s += '\t'; break;
                    case 'u':
                    {
                        unsigned code = std::stoul(std::string(p + 1, p + 5), nullptr, 16);
                        s += static_cast<char>(code & 0x7f);
                        p += 4;
                        break;
                    }
                    default: s += *p; break;
                    }
                    ++p;
                }
                else
                {
                    s += *p++;
                }
            }
            ++p;
            out.v = std::move(s);
            return out;
        }
        if (!std::strncmp(p, "true", 4)) { p += 4; out.v = true; return out; }
        if (!std::strncmp(p, "false", 5)) { p += 5; out.v = false; return out; }
        if (!std::strncmp(p, "null", 4)) { p += 4; return out; }
        // Number: an int unless it has '.', 'e', 'E', "inf" or "nan"
        const char *q = p;
        bool is_float = false;
        while (q < end && (std::isdigit(static_cast<unsigned char>(*q)) || *q == '-' ||
                                  *q == '+' || *q == '.' || *q == 'e' || *q == 'E' ||
                                  std::isalpha(static_cast<unsigned char>(*q))))
        {
            if (*q == '.' || *q == 'e' || *q == 'E' || std::isalpha(static_cast<unsigned char>(*q)))
                is_float = true;
            ++q;
        }
        const std::string tok(p, q);
        p = q;
        if (is_float)
            out.v = std::strtod(tok.c_str(), nullptr);
        else
            out.v = std::strtoll(tok.c_str(), nullptr, 10);
        return out;
    }
};

static JVal parse_json_line(const std::string &line)
{
    JParser jp{line.data(), line.data() + line.size()};
    return jp.parse();
}

// ---------------------------------------------------------------------------
// JSON -> reference attribute types
// ---------------------------------------------------------------------------
static sanafe::ModelAttribute to_attr(const JVal &j)
{
    sanafe::ModelAttribute a;
    if (j.is_bool())
        a.value = j.b();
    else if (j.is_int())
        a.value = static_cast<int>(j.i());
    else if (j.is_double())
        a.value = j.d();
    else if (j.is_str())
        a.value = j.s();
    else if (j.is_list())
    {
        std::vector<sanafe::ModelAttribute> l;
        for (const JVal &e : j.list()) l.push_back(to_attr(e));
        a.value = std::move(l);
    }
    else if (j.is_map())
    {
        // Same representation as yaml_parse_attribute_map: list of named attrs
        std::vector<sanafe::ModelAttribute> l;
        for (const auto &kv : j.map())
        {
            sanafe::ModelAttribute e = to_attr(kv.second);
            e.name = kv.first;
            l.push_back(std::move(e));
        }
        a.value = std::move(l);
    }
    else
        throw std::runtime_error("flat: null attribute");
    return a;
}

// attrs: [[key, value, fwd_synapse, fwd_dendrite, fwd_soma], ...]
static std::map<std::string, sanafe::ModelAttribute> to_attr_map(const JVal &j)
{
    std::map<std::string, sanafe::ModelAttribute> m;
    for (const JVal &e : j.list())
    {
        const JList &t = e.list();
        sanafe::ModelAttribute a = to_attr(t.at(1));
        a.forward_to_synapse = t.at(2).b();
        a.forward_to_dendrite = t.at(3).b();
        a.forward_to_soma = t.at(4).b();
        m[t.at(0).s()] = std::move(a);
    }
    return m;
}

static sanafe::NeuronConfiguration to_neuron_config(const JVal &j)
{
    sanafe::NeuronConfiguration c;
    if (const JVal *p = j.find("soma_hw_name")) c.soma_hw_name = p->s();
    if (const JVal *p = j.find("default_synapse_hw_name")) c.default_synapse_hw_name = p->s();
    if (const JVal *p = j.find("dendrite_hw_name")) c.dendrite_hw_name = p->s();
    if (const JVal *p = j.find("log_spikes")) c.log_spikes = p->b();
    if (const JVal *p = j.find("log_potential")) c.log_potential = p->b();
    if (const JVal *p = j.find("attrs")) c.model_attributes = to_attr_map(*p);
    return c;
}

static std::map<std::string, std::vector<sanafe::ModelAttribute>> to_attr_lists(const JVal &j)
{
    // lists: [[name, [values...], fwd_syn, fwd_den, fwd_soma], ...]
    std::map<std::string, std::vector<sanafe::ModelAttribute>> out;
    for (const JVal &e : j.list())
    {
        const JList &t = e.list();
        std::vector<sanafe::ModelAttribute> vals;
        vals.reserve(t.at(1).list().size());
        for (const JVal &v : t.at(1).list())
        {
            sanafe::ModelAttribute a = to_attr(v);
            a.forward_to_synapse = t.at(2).b();
            a.forward_to_dendrite = t.at(3).b();
            a.forward_to_soma = t.at(4).b();
            vals.push_back(std::move(a));
        }
        out[t.at(0).s()] = std::move(vals);
    }
    return out;
}

// ---------------------------------------------------------------------------
// Architecture records (dispatch of src/yaml_arch.cpp, data from the flat file)
// ---------------------------------------------------------------------------
static void set_flag(sanafe::PipelineUnitConfiguration &hw, const std::string &section)
{
    if (section == "synapse") hw.implements_synapse = true;
    else if (section == "dendrite") hw.implements_dendrite = true;
    else if (section == "soma") hw.implements_soma = true;
    else throw std::runtime_error("flat: bad unit section " + section);
}

struct Loaded
{
    std::unique_ptr<sanafe::Architecture> arch;
    std::unique_ptr<sanafe::SpikingNetwork> net;
    size_t synapses{0};
};

static void build_synth(const JVal &spec_j, sanafe::Architecture &arch,
        sanafe::SpikingNetwork &net, size_t &synapses)
{
    sfe_synth_spec s{};
    s.cores = spec_j.at("cores").i();
    s.neurons_per_core = spec_j.at("neurons_per_core").i();
    s.dest_cores = spec_j.at("dest_cores").i();
    s.syn_per_axon = spec_j.at("syn_per_axon").i();
    s.seed = spec_j.at("seed").i();
    s.bias_permille = spec_j.at("bias_permille").i();
    s.bias = spec_j.at("bias").d();
    s.threshold = spec_j.at("threshold").d();
    s.reset = spec_j.at("reset").d();
    s.leak_decay = spec_j.at("leak_decay").d();
    s.w_min = spec_j.at("w_min").i();
    s.w_max = spec_j.at("w_max").i();
    s.max_delay = spec_j.at("max_delay").i();
    s.log_spikes = spec_j.at("log_spikes").i();
    s.log_potential_n = spec_j.at("log_potential_n").i();

    sanafe::NeuronConfiguration cfg;
    cfg.soma_hw_name = spec_j.at("soma_hw_name").s();
    cfg.default_synapse_hw_name = spec_j.at("synapse_hw_name").s();
    cfg.dendrite_hw_name = spec_j.at("dendrite_hw_name").s();
    cfg.log_spikes = (s.log_spikes != 0);
    cfg.log_potential = false;
    auto mk = [](double v) {
        sanafe::ModelAttribute a;
        a.value = v;
        return a;
    };
    cfg.model_attributes["threshold"] = mk(s.threshold);
    cfg.model_attributes["reset"] = mk(s.reset);
    cfg.model_attributes["leak_decay"] = mk(s.leak_decay);

    const size_t total = static_cast<size_t>(s.cores) * s.neurons_per_core;
    sanafe::NeuronGroup &pop = net.create_neuron_group("pop", total, cfg);
    for (size_t n = 0; n < total; ++n)
    {
        sanafe::NeuronConfiguration nc;
        if (sfe_synth_has_bias(&s, n)) nc.model_attributes["bias"] = mk(s.bias);
        if (n < s.log_potential_n) nc.log_potential = true;
        if (!nc.model_attributes.empty() || nc.log_potential.has_value())
            pop.neurons[n].set_attributes(nc);
    }
    for (size_t n = 0; n < total; ++n)
    {
        const uint32_t home = static_cast<uint32_t>(n / s.neurons_per_core);
        sanafe::Neuron &src = pop.neurons[n];
        src.edges_out.reserve(static_cast<size_t>(s.dest_cores) * s.syn_per_axon);
        for (uint32_t k = 0; k < s.dest_cores; ++k)
        {
            const uint32_t dc = sfe_synth_dest_core(&s, home, k);
            const uint64_t axon = n * s.dest_cores + k;
            const sfe_synth_axon ap = sfe_synth_axon_params(&s, axon);
            for (uint32_t j = 0; j < s.syn_per_axon; ++j)
            {
                const uint64_t syn = axon * s.syn_per_axon + j;
                const size_t post = static_cast<size_t>(dc) * s.neurons_per_core +
                        sfe_synth_post(&s, &ap, j);
                const size_t idx = src.connect_to_neuron(pop.neurons[post]);
                sanafe::Connection &con = src.edges_out[idx];
                sanafe::ModelAttribute w;
                w.value = static_cast<double>(sfe_synth_weight(&s, syn));
                con.synapse_attributes["weight"] = w;
                if (s.max_delay > 0)
                {
                    sanafe::ModelAttribute d;
                    d.value = static_cast<int>(sfe_synth_delay(&s, syn));
                    con.synapse_attributes["delay"] = d;
                }
                ++synapses;
            }
        }
    }
    auto cores = arch.cores();
    for (size_t n = 0; n < total; ++n)
        pop.neurons[n].map_to_core(cores.at(n / s.neurons_per_core));
}

static Loaded load_flat(const std::string &path)
{
    std::ifstream in(path);
    if (!in) throw std::runtime_error("cannot open " + path);
    Loaded L;
    std::string line;
    sanafe::CoreConfiguration *core = nullptr;
    size_t tile_id = 0;
    while (std::getline(in, line))
    {
        if (line.empty()) continue;
        const JVal rec = parse_json_line(line);
        const JList &r = rec.list();
        const std::string &kind = r.at(0).s();
        if (kind == "noc")
        {
            const JVal &a = r.at(2);
            sanafe::NetworkOnChipConfiguration noc;
            noc.width_in_tiles = a.at("width").i();
            noc.height_in_tiles = a.at("height").i();
            noc.link_buffer_size = a.at("link_buffer_size").i();
            for (const JVal &kv : a.at("sync").list())
                noc.ts_sync_delay_table.values[kv.list().at(0).i()] = kv.list().at(1).d();
            L.arch = std::make_unique<sanafe::Architecture>(r.at(1).s(), noc);
        }
        else if (kind == "tile")
        {
            const JVal &a = r.at(2);
            sanafe::TilePowerMetrics m;
            m.energy_north_hop = a.at("energy_north_hop").d();
            m.latency_north_hop = a.at("latency_north_hop").d();
            m.energy_east_hop = a.at("energy_east_hop").d();
            m.latency_east_hop = a.at("latency_east_hop").d();
            m.energy_south_hop = a.at("energy_south_hop").d();
            m.latency_south_hop = a.at("latency_south_hop").d();
            m.energy_west_hop = a.at("energy_west_hop").d();
            m.latency_west_hop = a.at("latency_west_hop").d();
            m.log_energy = a.at("log_energy").b();
            tile_id = L.arch->create_tile(r.at(1).s(), m).id;
        }
        else if (kind == "core")
        {
            const JVal &a = r.at(2);
            sanafe::CorePipelineConfiguration pc;
            pc.buffer_position = sanafe::pipeline_parse_buffer_pos_str(
                    a.at("buffer_position").s(), a.at("buffer_inside_unit").b());
            pc.max_neurons_supported = a.at("max_neurons_supported").i();
            pc.log_energy = a.at("log_energy").b();
            core = &L.arch->create_core(r.at(1).s(), tile_id, pc);
        }
        else if (kind == "axon_in")
        {
            sanafe::AxonInPowerMetrics m;
            m.energy_message_in = r.at(2).d();
            m.latency_message_in = r.at(3).d();
            core->create_axon_in(r.at(1).s(), m);
        }
        else if (kind == "axon_out")
        {
            sanafe::AxonOutPowerMetrics m;
            m.energy_message_out = r.at(2).d();
            m.latency_message_out = r.at(3).d();
            core->create_axon_out(r.at(1).s(), m);
        }
        else if (kind == "unit" || kind == "unit_range")
        {
            // ["unit", section, name, {model, plugin?, log_energy, log_latency,
            //   update_every_timestep, attrs:{...}}]
            // ["unit_range", section, base, first, last, {...}] = base[first..last]
            const bool is_range = (kind == "unit_range");
            const std::string &section = r.at(1).s();
            const JVal &a = is_range ? r.at(5) : r.at(3);
            const long long first = is_range ? r.at(3).i() : 0;
            const long long last = is_range ? r.at(4).i() : 0;
            for (long long ui = first; ui <= last; ++ui)
            {
            const std::string name = is_range ? r.at(2).s() + "[" + std::to_string(ui) + "]" : r.at(2).s();
            sanafe::ModelInfo mi;
            mi.name = a.at("model").s();
            if (const JVal *p = a.find("plugin")) mi.plugin_library_path = p->s();
            mi.log_energy = a.at("log_energy").b();
            mi.log_latency = a.at("log_latency").b();
            mi.update_every_timestep = a.at("update_every_timestep").b();
            for (const auto &kv : a.at("attrs").map())
                mi.model_attributes[kv.first] = to_attr(kv.second);
            // src/yaml_arch.cpp:149-186 (merge same-named unit across sections)
            bool exists = false;
            for (auto &hw : core->pipeline_hw)
            {
                if (hw.name == name)
                {
                    exists = true;
                    set_flag(hw, section);
                    hw.model_info.model_attributes.merge(mi.model_attributes);
                    if (mi.plugin_library_path.has_value())
                        hw.model_info.plugin_library_path = mi.plugin_library_path;
                    break;
                }
            }
            if (!exists) set_flag(core->create_hardware_unit(name, mi), section);
            }
        }
        else if (kind == "end_arch")
        {
            L.net = std::make_unique<sanafe::SpikingNetwork>("net");
        }
        else if (kind == "group")
        {
            L.net->create_neuron_group(r.at(1).s(), r.at(2).i(), to_neuron_config(r.at(3)));
        }
        else if (kind == "neurons")
        {
            // ["neurons", group, first, last, config] — config already merged
            // with the group default (src/yaml_snn.cpp:304-329)
            sanafe::NeuronGroup &g = L.net->groups.at(r.at(1).s());
            const sanafe::NeuronConfiguration c = to_neuron_config(r.at(4));
            for (long long i = r.at(2).i(); i <= r.at(3).i(); ++i) g.neurons.at(i).set_attributes(c);
        }
        else if (kind == "edge")
        {
            sanafe::Neuron &src = L.net->groups.at(r.at(1).s()).neurons.at(r.at(2).i());
            sanafe::Neuron &dst = L.net->groups.at(r.at(3).s()).neurons.at(r.at(4).i());
            const size_t idx = src.connect_to_neuron(dst);
            sanafe::Connection &con = src.edges_out[idx];
            con.synapse_attributes = to_attr_map(r.at(5).at("synapse_attrs"));
            con.dendrite_attributes = to_attr_map(r.at(5).at("dendrite_attrs"));
        }
        else if (kind == "conv2d")
        {
            sanafe::NeuronGroup &sg = L.net->groups.at(r.at(1).s());
            sanafe::NeuronGroup &dg = L.net->groups.at(r.at(2).s());
            const JVal &p = r.at(3);
            sanafe::Conv2DParameters c;
            c.input_width = p.at("input_width").i();
            c.input_height = p.at("input_height").i();
            c.input_channels = p.at("input_channels").i();
            c.kernel_width = p.at("kernel_width").i();
            c.kernel_height = p.at("kernel_height").i();
            c.kernel_count = p.at("kernel_count").i();
            c.stride_width = p.at("stride_width").i();
            c.stride_height = p.at("stride_height").i();
            sg.connect_neurons_conv2d(dg, to_attr_lists(r.at(4)), c);
        }
        else if (kind == "dense")
        {
            sanafe::NeuronGroup &sg = L.net->groups.at(r.at(1).s());
            sanafe::NeuronGroup &dg = L.net->groups.at(r.at(2).s());
            sg.connect_neurons_dense(dg, to_attr_lists(r.at(3)));
        }
        else if (kind == "sparse")
        {
            sanafe::NeuronGroup &sg = L.net->groups.at(r.at(1).s());
            sanafe::NeuronGroup &dg = L.net->groups.at(r.at(2).s());
            std::vector<std::pair<size_t, size_t>> pairs;
            for (const JVal &pr : r.at(3).list())
                pairs.emplace_back(pr.list().at(0).i(), pr.list().at(1).i());
            sg.connect_neurons_sparse(dg, to_attr_lists(r.at(4)), pairs);
        }
        else if (kind == "map")
        {
            // ["map", group, first, last, tile, core_offset, {synapse?,dendrite?,soma?}]
            sanafe::NeuronGroup &g = L.net->groups.at(r.at(1).s());
            const JVal &hw = r.at(6);
            for (long long i = r.at(2).i(); i <= r.at(3).i(); ++i)
            {
                sanafe::Neuron &n = g.neurons.at(i);
                if (const JVal *p = hw.find("synapse")) n.default_synapse_hw_name = p->s();
                if (const JVal *p = hw.find("dendrite")) n.dendrite_hw_name = p->s();
                if (const JVal *p = hw.find("soma")) n.soma_hw_name = p->s();
                n.map_to_core(L.arch->tiles.at(r.at(4).i()).cores.at(r.at(5).i()));
            }
        }
        else if (kind == "synth")
        {
            build_synth(r.at(1), *L.arch, *L.net, L.synapses);
        }
        else if (kind == "end_net")
        {
            break;
        }
        else
        {
            throw std::runtime_error("flat: unknown record " + kind);
        }
    }
    L.synapses = 0;
    for (const auto &[name, g] : L.net->groups)
        for (const auto &n : g.neurons) L.synapses += n.edges_out.size();
    return L;
}

// ---------------------------------------------------------------------------
static void dump_map(sanafe::SpikingChip &chip, const std::string &path)
{
    // Structure of the mapped chip, for cross-checking the new engine's lowering
    std::ofstream out(path);
    for (const sanafe::Tile &t : chip.tiles)
    {
        for (const sanafe::Core &c : t.cores)
        {
            if (c.neurons.empty() && c.axons_in.empty()) continue;
            out << "core " << c.id << " tile " << t.id << " x " << t.x << " y " << t.y
                << " neurons " << c.neurons.size() << " axons_in " << c.axons_in.size()
                << " axons_out " << c.axons_out.size() << "\n";
            for (const sanafe::MappedNeuron &n : c.neurons)
            {
                out << " n " << n.parent_group_name << "." << n.offset << " soma "
                    << n.soma_hw->name << "@" << n.mapped_soma_hw_address << " dend "
                    << n.dendrite_hw->name << "@" << n.mapped_dendrite_hw_address << " out";
                for (size_t a : n.axon_out_addresses)
                {
                    const sanafe::AxonOutModel &ao = c.axons_out[a];
                    out << " " << ao.dest_tile_id << "." << ao.dest_core_offset << ":"
                        << ao.dest_axon_id;
                }
                out << "\n";
            }
            for (size_t a = 0; a < c.axons_in.size(); ++a)
            {
                out << " a " << a;
                for (size_t s : c.axons_in[a].synapse_addresses)
                {
                    const sanafe::MappedConnection &con = *c.connections_in[s];
                    const sanafe::MappedNeuron &pre = con.pre_neuron_ref;
                    const sanafe::MappedNeuron &post = con.post_neuron_ref;
                    out << " " << pre.parent_group_name << "." << pre.offset << ">"
                        << post.mapped_offset_within_core << "/" << con.synapse_hw->name << "@"
                        << con.mapped_synapse_hw_address;
                }
                out << "\n";
            }
        }
    }
}

static void write_rundata(std::ostream &o, const sanafe::RunData &rd)
{
    char buf[1024];
    std::snprintf(buf, sizeof(buf),
            "\"timestep_start\": %ld, \"timesteps_executed\": %ld, \"spikes\": %ld, "
            "\"packets_sent\": %ld, \"neurons_updated\": %ld, \"neurons_fired\": %ld, "
            "\"total_energy\": %.17g, \"synapse_energy\": %.17g, \"dendrite_energy\": %.17g, "
            "\"soma_energy\": %.17g, \"network_energy\": %.17g, \"sim_time\": %.17g",
            rd.timestep_start, rd.timesteps_executed, rd.spikes, rd.packets_sent,
            rd.neurons_updated, rd.neurons_fired, rd.total_energy, rd.synapse_energy,
            rd.dendrite_energy, rd.soma_energy, rd.network_energy, rd.sim_time);
    o << buf;
}

int main(int argc, char **argv)
{
    if (argc < 2)
    {
        std::fprintf(stderr, "usage: %s flat.jsonl --steps N [--timing simple|detailed] "
                             "[--threads N] [--out DIR] [--traces] [--per-step] [--dump-map] [--reps R]\n",
                argv[0]);
        return 2;
    }
    std::string flat = argv[1];
    long steps = 1;
    std::string timing = "simple";
    int threads = 1;
    std::string out_dir = ".";
    bool traces = false;
    bool per_step = false;
    bool want_map = false;
    int reps = 1;
    long hash_steps = 0;
    for (int i = 2; i < argc; ++i)
    {
        const std::string a = argv[i];
        if (a == "--steps") steps = std::atol(argv[++i]);
        else if (a == "--timing") timing = argv[++i];
        else if (a == "--threads") threads = std::atoi(argv[++i]);
        else if (a == "--out") out_dir = argv[++i];
        else if (a == "--traces") traces = true;
        else if (a == "--per-step") per_step = true;
        else if (a == "--dump-map") want_map = true;
        else if (a == "--reps") reps = std::atoi(argv[++i]);
        else if (a == "--hash-steps") hash_steps = std::atol(argv[++i]);
        else { std::fprintf(stderr, "unknown flag %s\n", a.c_str()); return 2; }
    }
#ifdef HAVE_OPENMP
    omp_set_num_threads(threads); // == CLI -N (src/arg_parsing.cpp:82-95)
#endif
    std::filesystem::create_directories(out_dir);
    try
    {
        using clk = std::chrono::steady_clock;
        const auto t0 = clk::now();
        Loaded L = load_flat(flat);
        const auto t1 = clk::now();
        sanafe::SpikingChip chip(*L.arch);
        chip.load(*L.net);
        const auto t2 = clk::now();
        if (want_map) dump_map(chip, out_dir + "/map.txt");

        const sanafe::TimingModel tm = sanafe::parse_timing_model(timing);
        sanafe::TraceFlags tf;
        if (traces)
        {
            tf.record_spikes = true;
            tf.record_potentials = true;
            tf.record_neuron_state = true;
            tf.record_perf = true;
            tf.record_messages = true;
        }
        size_t neurons = 0;
        for (const auto &[name, g] : chip.mapped_neuron_groups) neurons += g.size();

        sanafe::RunData total(1);
        double best_wall = 1e300;
        double wall_sum = 0.0;
        std::string calls_json;
        unsigned long long raster_hash = 0ull;
        sanafe::RunData hashed(1);
        for (long t = 1; t <= hash_steps; ++t)
        {
            const sanafe::RunData rd = chip.sim(1, tm, 0, tf, out_dir);
            for (const auto &addr : chip.get_spikes())
                raster_hash += sfe_mix64((static_cast<unsigned long long>(t) << 32) |
                        static_cast<unsigned long long>(addr.neuron_offset.value()));
            hashed.total_energy += rd.total_energy;
            hashed.sim_time += rd.sim_time;
            hashed.spikes += rd.spikes;
            hashed.packets_sent += rd.packets_sent;
            hashed.neurons_updated += rd.neurons_updated;
            hashed.neurons_fired += rd.neurons_fired;
        }
        if (per_step)
        {
            // Full-precision per-step records through the public getters
            // (CSV traces only keep 6 digits, src/chip.cpp:1648-1650)
            std::ofstream steps_f(out_dir + "/steps.csv");
            std::ofstream pot_f(out_dir + "/potentials_full.csv");
            std::ofstream spk_f(out_dir + "/spikes_full.csv");
            std::ofstream trc_f(out_dir + "/traces_full.csv");
            steps_f << "timestep,fired,updated,packets,spikes,sim_time,synapse_energy,"
                       "dendrite_energy,soma_energy,network_energy,total_energy\n";
            const auto w0 = clk::now();
            for (long t = 1; t <= steps; ++t)
            {
                const sanafe::RunData rd = chip.sim(1, tm, 0, tf, out_dir);
                char buf[512];
                std::snprintf(buf, sizeof(buf), "%ld,%ld,%ld,%ld,%ld,%.17g,%.17g,%.17g,%.17g,%.17g,%.17g\n",
                        t, rd.neurons_fired, rd.neurons_updated, rd.packets_sent, rd.spikes,
                        rd.sim_time, rd.synapse_energy, rd.dendrite_energy, rd.soma_energy,
                        rd.network_energy, rd.total_energy);
                steps_f << buf;
                const auto pots = chip.get_potentials();
                if (!pots.empty())
                {
                    pot_f << t;
                    for (double v : pots)
                    {
                        std::snprintf(buf, sizeof(buf), ",%.17g", v);
                        pot_f << buf;
                    }
                    pot_f << "\n";
                }
                for (const auto &addr : chip.get_spikes())
                    spk_f << addr.group_name << "." << addr.neuron_offset.value() << "," << t << "\n";
                const auto trs = chip.get_traces();
                if (!trs.empty())
                {
                    trc_f << t;
                    for (const auto &[name, vals] : trs)
                        for (double v : vals)
                        {
                            std::snprintf(buf, sizeof(buf), ",%s=%.17g", name.c_str(), v);
                            trc_f << buf;
                        }
                    trc_f << "\n";
                }
                total.total_energy += rd.total_energy;
                total.synapse_energy += rd.synapse_energy;
                total.dendrite_energy += rd.dendrite_energy;
                total.soma_energy += rd.soma_energy;
                total.network_energy += rd.network_energy;
                total.sim_time += rd.sim_time;
                total.spikes += rd.spikes;
                total.packets_sent += rd.packets_sent;
                total.neurons_updated += rd.neurons_updated;
                total.neurons_fired += rd.neurons_fired;
                total.timesteps_executed += 1;
            }
            best_wall = std::chrono::duration<double>(clk::now() - w0).count();
            wall_sum = best_wall;
        }
        else
        {
            // reps > 1: consecutive sim() calls on the same chip (state carries
            // on); wall of each call recorded, totals are of the LAST call
            std::vector<double> walls;
            for (int r = 0; r < reps; ++r)
            {
                const auto w0 = clk::now();
                total = chip.sim(steps, tm, 0, tf, out_dir);
                const double w = std::chrono::duration<double>(clk::now() - w0).count();
                walls.push_back(w);
                wall_sum += w;
                best_wall = w; // earlier calls are warm-up: report the LAST call (its RunData is what we print)
                char cb[96];
                std::snprintf(cb, sizeof(cb), "%s[%ld, %.6f]", r == 0 ? "" : ", ", total.spikes, w);
                calls_json += cb; // every call's synaptic events and wall: the caller takes the best rate
            }
        }
        std::ofstream sf(out_dir + "/summary.json");
        sf << "{";
        write_rundata(sf, total);
        char buf[512];
        std::snprintf(buf, sizeof(buf),
                ", \"power\": %.17g, \"timing\": \"%s\", \"threads\": %d, \"neurons\": %zu, "
                "\"synapses\": %zu, \"mapped_cores\": %zu, \"mapped_tiles\": %zu, "
                "\"build_s\": %.6f, \"load_s\": %.6f, \"sim_wall_s\": %.6f, \"sim_wall_total_s\": %.6f, "
                "\"reps\": %d",
                chip.get_power(), timing.c_str(), threads, neurons, L.synapses, chip.mapped_cores,
                chip.mapped_tiles, std::chrono::duration<double>(t1 - t0).count(),
                std::chrono::duration<double>(t2 - t1).count(), best_wall, wall_sum, reps);
        sf << buf;
        if (hash_steps > 0)
        {
            std::snprintf(buf, sizeof(buf),
                    ", \"hash_steps\": %ld, \"raster_hash\": \"%016llx\", \"hash_spikes\": %ld, \"hash_packets\": %ld, "
                    "\"hash_updated\": %ld, \"hash_fired\": %ld, \"hash_energy\": %.17g, \"hash_sim_time\": %.17g",
                    hash_steps, raster_hash, hashed.spikes, hashed.packets_sent, hashed.neurons_updated,
                    hashed.neurons_fired, hashed.total_energy, hashed.sim_time);
            sf << buf;
        }
        sf << ", \"calls\": [" << calls_json << "]}\n";
        sf.close();
        std::ifstream back(out_dir + "/summary.json");
        std::cout << back.rdbuf();
    }
    catch (const std::exception &e)
    {
        std::fprintf(stderr, "sanafe_ref: exception: %s\n", e.what());
        return 1;
    }
    return 0;
}
