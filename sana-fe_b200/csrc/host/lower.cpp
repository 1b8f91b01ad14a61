// lower.cpp — see lower.hpp.
#include "lower.hpp"

#include "device_model.hpp"
#include "sfe_device_model.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <set>
#include <unordered_map>

namespace sfe
{
namespace
{
constexpr size_t kLifMaxCompartments = 1024;   // src/models.hpp:29
constexpr size_t kTrueNorthMaxNeurons = 4096;  // src/models.hpp:283
constexpr size_t kMaxDelay = 5;                // src/models.hpp:158

enum class UnitModel
{
    current_based,
    accumulator,
    accumulator_with_delay,
    lif,
    truenorth,
    input,
    hodgkin_huxley,
    taps,
    neurofem,
    device_model // a soma model registered out of tree (include/sfe_device_model.h)
};

struct UnitKey
{
    int family;
    int instance;
    bool operator<(const UnitKey &o) const
    {
        return family != o.family ? family < o.family : instance < o.instance;
    }
    bool operator==(const UnitKey &o) const { return family == o.family && instance == o.instance; }
};

struct UnitState // one used hardware unit instance of one core
{
    UnitModel model{};
    size_t neuron_count{0};     // PipelineUnit::add_neuron     src/pipeline.cpp:79-85
    size_t connection_count{0}; // PipelineUnit::add_connection src/pipeline.cpp:59-77
    std::vector<uint8_t> delays; // accumulator_with_delay: by synapse address (src/models.cpp:133-152)
    // "input" / HH keep ONE state per unit
    std::vector<uint8_t> spikes;
    double rate{0.0};
    double poisson{0.0};
    double hh_m{0.0}, hh_n{0.0}, hh_h{0.0}, hh_i{0.0};
    // "taps" (MultiTapModel1D): vector sizes evolve exactly as set_attribute_neuron leaves them (src/models.cpp:259-318)
    size_t n_taps{1};
    std::vector<double> time_constants{0.0}, space_constants{};
    std::vector<int> synapse_to_tap; // by synapse address (set_attribute_edge "tap", src/models.cpp:320-332)
    std::vector<uint8_t> synapse_to_compartment; // "neurofem": by synapse address (plugins/neurofem.cpp:120-136), default 0
    uint32_t noise_off{0}, noise_len{0}; // LIF unit with a noise file: its entries in HostTables::noise_values
    bool noise_loaded{false};
    const sfe_device_model_desc *device_model{nullptr}; // UnitModel::device_model: the registered descriptor
    std::vector<uint32_t> sharing; // device-order list of neurons mapped to this unit (filled later)
};

struct LConn
{
    uint32_t post_core;
    uint32_t post_in_core;
    UnitKey syn;
    uint32_t syn_addr;
    double weight;
};

struct LNeuron
{
    const Neuron *n;
    uint32_t group;
    UnitKey dend, soma;
    UnitState *dend_unit{nullptr}, *soma_unit{nullptr}; // into LCore::units (stable addresses)
    size_t dend_addr, soma_addr;
    sfe_soma_class cls;
    double bias{0.0};
    double potential0{0.0};
    std::vector<double> dm_state, dm_params; // out-of-tree device model: this instance's initial state words / parameters
    std::vector<LConn> out;
    std::vector<std::pair<uint32_t, uint32_t>> axons_out; // (dest core, axon index in dest core)
};

struct LAxonIn
{
    uint32_t src_core;
    uint32_t src_in_core;
    std::vector<std::pair<uint32_t, uint32_t>> syn; // (src neuron's connection index) -> resolved later
};

struct LCore
{
    const CoreConfiguration *cfg;
    std::vector<LNeuron> neurons;
    std::map<UnitKey, UnitState> units; // (node-based: element addresses are stable)
    std::vector<LAxonIn> axons_in;
    std::unordered_map<std::string, std::pair<UnitKey, UnitState *>> synapse_hw; // get_hw results by unit name
};

// The out-of-tree device model a unit resolves to, or null: plugin_get_hw (src/plugins.cpp:85-98) with device code in
// place of host virtuals. A model registered under the unit's model name wins (sfe_register_device_model /
// sfe_load_device_model); otherwise the unit's `plugin:` library is loaded and asked for sfe_device_model_<model>().
// *why receives the reason a named library did not yield a model.
const sfe_device_model_desc *resolve_device_model(const PipelineUnitConfiguration &u, std::string *why)
{
    const std::string &m = u.model_info.name;
    if (const sfe_device_model_desc *d = find_device_model(m)) return d;
    if (!u.model_info.plugin_library_path.has_value()) return nullptr;
    return load_device_model(m, *u.model_info.plugin_library_path, why);
}

UnitModel parse_model(const PipelineUnitConfiguration &u)
{
    const std::string &m = u.model_info.name;
    if (u.model_info.plugin_library_path.has_value())
    {
        // Plugins are host C++ virtuals in the reference (src/plugins.cpp:45-98). The device path needs device code:
        // a model registered through include/sfe_device_model.h (also what a `plugin:` library built with that header
        // provides), else one of the reference's own plugins, whose functors are compiled into the engine.
        // Anything else cannot run "with no CPU fallback".
        std::string why;
        if (resolve_device_model(u, &why) != nullptr) return UnitModel::device_model;
        if (m == "hodgkin_huxley") return UnitModel::hodgkin_huxley;
        if (m == "neurofem") return UnitModel::neurofem;
        throw std::runtime_error("Plugin model '" + m + "' (" + *u.model_info.plugin_library_path +
                ") has no device model: " + why + ". Host-only plugins are not supported by the B200 engine; see "
                "include/sfe_device_model.h");
    }
    if (find_device_model(m) != nullptr && m != "current_based" && m != "accumulator" && m != "accumulator_with_delay" &&
            m != "leaky_integrate_fire" && m != "truenorth" && m != "input" && m != "taps")
        return UnitModel::device_model;
    if (m == "current_based") return UnitModel::current_based;
    if (m == "accumulator") return UnitModel::accumulator;
    if (m == "accumulator_with_delay") return UnitModel::accumulator_with_delay;
    if (m == "leaky_integrate_fire") return UnitModel::lif;
    if (m == "truenorth") return UnitModel::truenorth;
    if (m == "input") return UnitModel::input;
    if (m == "taps") return UnitModel::taps;
    throw std::invalid_argument("Pipeline model not supported (" + m + ")\n"); // src/models.cpp:964-966
}

// Number of "input"-model soma units created before unit `key` of core `core_id` when the chip is
// built (SpikingChip::SpikingChip, src/chip.cpp:61-92: tiles, cores, pipeline units in order;
// name[a..b] families expand ascending). The reference seeds every InputModel's std::mt19937 with
// a running instance counter (src/models.hpp:347,366), so this ordinal fixes the Poisson stream.
bool is_builtin_input(const PipelineUnitConfiguration &u)
{
    return !u.model_info.plugin_library_path.has_value() && u.model_info.name == "input";
}

uint32_t input_unit_ordinal(const Architecture &arch, const size_t core_id, const UnitKey &key)
{
    uint32_t ordinal = 0;
    for (const TileConfiguration &tile : arch.tiles)
        for (const CoreConfiguration &core : tile.cores)
        {
            const bool here = core.address.id == core_id;
            for (size_t f = 0; f < core.pipeline_hw.size(); ++f)
            {
                const PipelineUnitConfiguration &u = core.pipeline_hw[f];
                if (here && static_cast<int>(f) == key.family)
                    return ordinal + static_cast<uint32_t>(key.instance - u.first_instance());
                if (is_builtin_input(u)) ordinal += u.is_range ? static_cast<uint32_t>(u.range_last - u.range_first + 1) : 1u;
            }
            if (here) throw std::logic_error("input_unit_ordinal: unit not found in its core");
        }
    throw std::logic_error("input_unit_ordinal: core not found");
}

// Core::get_hw  src/core.cpp:61-97
UnitKey get_hw(const CoreConfiguration &core, const std::string &hw_name, const bool is_synapse,
        const bool is_dendrite, const bool is_soma)
{
    const bool first_available = hw_name.empty();
    for (size_t f = 0; f < core.pipeline_hw.size(); ++f)
    {
        const PipelineUnitConfiguration &hw = core.pipeline_hw[f];
        if ((is_synapse && !hw.implements_synapse) || (is_dendrite && !hw.implements_dendrite) ||
                (is_soma && !hw.implements_soma))
            continue;
        if (first_available) return {static_cast<int>(f), hw.first_instance()};
        const int inst = hw.match(hw_name);
        if (inst >= 0) return {static_cast<int>(f), inst};
    }
    throw HardwareMappingError("Could not find h/w (with name:" + hw_name + ") that implements synapse:" +
            std::to_string(static_cast<int>(is_synapse)) + ", dendrite:" +
            std::to_string(static_cast<int>(is_dendrite)) + ", soma:" + std::to_string(static_cast<int>(is_soma)));
}

double unit_double(const PipelineUnitConfiguration &u, const char *key, const char *what)
{
    auto it = u.model_info.model_attributes.find(key);
    if (it == u.model_info.model_attributes.end())
        throw std::runtime_error(std::string(what) + " unit '" + u.name + "' does not simulate energy/latency or " +
                "provide a default cost (" + key + ") in the architecture description.");
    return it->second.as_double();
}

uint32_t parse_reset_mode(const std::string &s) // src/models.cpp:905-931
{
    if (s == "none") return SFE_RESET_NONE;
    if (s == "soft") return SFE_RESET_SOFT;
    if (s == "hard") return SFE_RESET_HARD;
    if (s == "saturate") return SFE_RESET_SATURATE;
    throw std::invalid_argument("Reset mode not recognized");
}

// A unit that appears in several sections of a core implements several pipeline functions (src/pipeline.hpp:307-420:
// its input interface is that of the first function, its output interface that of the last). The built-in models each
// implement one function; the combined unit among the shipped models is the "neurofem" plugin (dendrite + soma,
// plugins/neurofem.cpp:33-37, arch/neurofem.yaml).
void check_unit_shape(const PipelineUnitConfiguration &u)
{
    const int functions = int(u.implements_synapse) + int(u.implements_dendrite) + int(u.implements_soma);
    if (functions == 1) return;
    if (u.implements_synapse && u.implements_soma && !u.implements_dendrite) // src/pipeline.hpp:343-350
        throw std::logic_error("Invalid pipeline configuration: h/w supports synapse and soma but not dendrite functionality. "
                               "To fix this, either add this to the core's dendrite section, or remove from either the "
                               "synapse or soma sections.");
    const bool neurofem = u.model_info.plugin_library_path.has_value() && u.model_info.name == "neurofem";
    if (neurofem && u.implements_dendrite && u.implements_soma && !u.implements_synapse) return;
    throw std::runtime_error("Hardware unit '" + u.name + "' (model " + u.model_info.name +
            ") is listed in several sections of its core: of the models this engine carries only 'neurofem' is a combined "
            "(dendrite + soma) unit");
}

void soma_defaults(const PipelineUnitConfiguration &u, const UnitModel model, sfe_soma_class &c)
{
    // defaults of LoihiCompartment / TrueNorthNeuron  src/models.hpp:222-241, 284-299
    std::memset(&c, 0, sizeof(c));
    c.reset_mode = SFE_RESET_HARD;
    c.reverse_reset_mode = SFE_RESET_NONE;
    switch (model)
    {
    case UnitModel::lif:
        c.model = SFE_SOMA_LIF;
        c.leak = 1.0;
        break;
    case UnitModel::truenorth:
        c.model = SFE_SOMA_TRUENORTH;
        c.flags = SFE_SOMA_LEAK_TOWARDS_ZERO;
        break;
    case UnitModel::input:
        c.model = SFE_SOMA_INPUT;
        break;
    case UnitModel::hodgkin_huxley:
        c.model = SFE_SOMA_HH;
        break;
    case UnitModel::device_model:
        c.model = SFE_SOMA_DEVICE_MODEL;
        break;
    case UnitModel::neurofem:
        c.model = SFE_SOMA_NEUROFEM; // NeuroFEMNeuron defaults  plugins/neurofem.cpp:64-84: everything 0 but dt
        c.nf_dt = 1.0e-3;
        c.flags = SFE_SOMA_IS_DENDRITE;
        break;
    default:
        throw std::runtime_error("Unit '" + u.name + "' is not a soma model");
    }
    // src/pipeline.cpp:206-266: all three metrics or none
    c.energy_access = unit_double(u, "energy_access_neuron", "Soma");
    c.energy_update = unit_double(u, "energy_update_neuron", "Soma");
    c.energy_spike_out = unit_double(u, "energy_spike_out", "Soma");
    c.latency_access = unit_double(u, "latency_access_neuron", "Soma");
    c.latency_update = unit_double(u, "latency_update_neuron", "Soma");
    c.latency_spike_out = unit_double(u, "latency_spike_out", "Soma");
}

// One pass over a LIF unit's noise file, entry by entry as LoihiLifModel::loihi_read_noise_stream /
// loihi_generate_noise (src/models.cpp:589-650) would deliver them before rewinding: a line is consumed per
// update (an unparsable line yields 0), the stream rewinds when the next read would hit the end of the file,
// and the value is cut to `noise_bits` bits with bit 8 as the sign (src/models.hpp:273-275).
std::vector<double> read_noise_file(const std::string &path, const int noise_bits)
{
    std::ifstream f(path);
    if (!f.is_open()) throw std::runtime_error("Failed to open noise stream"); // src/models.cpp:360-365
    const long sign_mask = 0x100;
    const long random_mask = (1L << noise_bits) - 1L;
    std::vector<double> values;
    while (!(f.eof() || f.peek() == std::ifstream::traits_type::eof()))
    {
        std::string line;
        if (!std::getline(f, line)) break;
        int parsed = 0;
        std::istringstream iss(line);
        if (!(iss >> parsed)) parsed = 0;
        long v = parsed;
        const long sign_bit = v & sign_mask;
        v &= random_mask;
        if (sign_bit != 0) v |= ~random_mask;
        values.push_back(static_cast<double>(v));
    }
    if (values.empty()) throw std::runtime_error("Couldn't read noise entry from file"); // src/models.cpp:620-624
    return values;
}

// MultiTapModel1D::set_attribute_neuron  src/models.cpp:259-318, branch for branch (the vectors' sizes depend on
// the order in which the attributes arrive, which is the std::map order of the attribute names)
void set_taps_attribute(UnitState &u, const std::string &key, const Attr &a)
{
    auto doubles = [](const Attr &list) {
        std::vector<double> v;
        for (const Attr &e : list.as_list()) v.push_back(e.as_double());
        return v;
    };
    if (key == "taps")
    {
        const size_t n = static_cast<size_t>(a.as_int());
        if (n == 0) throw std::invalid_argument("Number of taps must be > 0\n");
        u.n_taps = n;
        u.time_constants.resize(n);
        u.space_constants.resize(n - 1);
    }
    else if (key == "time_constants")
    {
        u.time_constants = doubles(a);
        if (u.time_constants.size() < u.n_taps)
            throw std::invalid_argument("Expected " + std::to_string(u.n_taps) + " but received " +
                    std::to_string(u.time_constants.size()) + "time constants.");
        if (u.time_constants.size() > u.n_taps) u.space_constants.resize(u.n_taps - 1);
    }
    else if (key == "space_constants")
    {
        const size_t previous = u.time_constants.size();
        u.space_constants = doubles(a);
        if (u.space_constants.size() < u.n_taps - 1)
            throw std::invalid_argument("Expected " + std::to_string(u.n_taps - 1) + " but received " +
                    std::to_string(previous) + "time constants.");
        if (u.space_constants.size() > u.n_taps - 1) u.time_constants.resize(u.n_taps);
    }
}

// LoihiLifModel / TrueNorthModel / InputModel / HodgkinHuxley ::set_attribute_neuron
// (src/models.cpp:375-439, 664-722, 832-853; plugins/hodgkin_huxley.cpp:94-114)
void set_soma_attribute(LNeuron &ln, UnitState &unit, const UnitModel model, const std::string &key, const Attr &a)
{
    sfe_soma_class &c = ln.cls;
    auto set_flag = [&](uint32_t flag, bool on) { c.flags = on ? (c.flags | flag) : (c.flags & ~flag); };
    if (model == UnitModel::lif)
    {
        if (key == "threshold") c.threshold = a.as_double();
        else if (key == "reverse_threshold") c.reverse_threshold = a.as_double();
        else if (key == "reset") c.reset = a.as_double();
        else if (key == "reverse_reset") c.reverse_reset = a.as_double();
        else if (key == "reset_mode") c.reset_mode = parse_reset_mode(a.as_string());
        else if (key == "reverse_reset_mode") c.reverse_reset_mode = parse_reset_mode(a.as_string());
        else if (key == "leak_decay") c.leak = a.as_double();
        else if (key == "log_u") set_flag(SFE_SOMA_LOG_U, a.as_bool());
        else if (key == "input_decay") c.input_decay = a.as_double();
        else if (key == "bias") ln.bias = a.as_double();
        else if (key == "force_update" || key == "force_update_every_timestep")
            set_flag(SFE_SOMA_FORCE_UPDATE, a.as_bool());
        else if (key == "refractory_delay") c.refractory_delay = a.as_int();
        else if (key == "potential") ln.potential0 = a.as_double();
    }
    else if (model == UnitModel::truenorth)
    {
        if (key == "threshold") c.threshold = a.as_double();
        else if (key == "reverse_threshold") c.reverse_threshold = a.as_double();
        else if (key == "reset") c.reset = a.as_double();
        else if (key == "reverse_reset") c.reverse_reset = a.as_double();
        else if (key == "reset_mode") c.reset_mode = parse_reset_mode(a.as_string());
        else if (key == "reverse_reset_mode") c.reverse_reset_mode = parse_reset_mode(a.as_string());
        else if (key == "leak") c.leak = a.as_double();
        else if (key == "bias") ln.bias = a.as_double();
        else if (key == "force_update_every_timestep" || key == "force_update")
            set_flag(SFE_SOMA_FORCE_UPDATE, a.as_bool());
        else if (key == "leak_towards_zero") set_flag(SFE_SOMA_LEAK_TOWARDS_ZERO, a.as_bool());
        else if (key == "random_mask")
        {
            const int mask = a.as_int();
            if (mask < 0) throw std::invalid_argument("random_mask < 0; must be unsigned.");
            // the jitter draws from the process-global std::rand() (src/models.cpp:757): reproduced for the reference's
            // single-threaded order (one draw per such neuron and timestep, cores and neurons in order; csrc/host/poisson.cpp)
            c.random_mask = static_cast<uint32_t>(mask);
        }
    }
    else if (model == UnitModel::input)
    {
        if (key == "spikes")
        {
            unit.spikes.clear();
            for (const Attr &e : a.as_list()) unit.spikes.push_back(e.as_bool() ? 1 : 0);
        }
        else if (key == "poisson")
        {
            unit.poisson = a.as_double();
        }
        else if (key == "rate")
        {
            unit.rate = a.as_double();
            // InputModel::update takes timestep % (long)(1 / rate) (src/models.cpp:887-890): a rate above 1 divides by zero
            // (the reference's host build dies with SIGFPE); refuse it here instead of handing the device a zero divisor
            if (unit.rate > 1.0)
                throw std::invalid_argument("input rate " + std::to_string(unit.rate) + " > 1: the period (long)(1 / rate) would be 0");
        }
    }
    else if (model == UnitModel::neurofem)
    {
        // NeuroFEMModel::set_attribute_neuron  plugins/neurofem.cpp:138-190
        if (key == "threshold") c.threshold = a.as_double();
        else if (key == "reset") c.reset = a.as_double();
        else if (key == "lambda_d") c.input_decay = a.as_double();
        else if (key == "lambda_v") c.leak = a.as_double();
        else if (key == "bias") ln.bias = a.as_double();
        else if (key == "dt") c.nf_dt = a.as_double();
        else if (key == "kp") c.nf_kp = a.as_double();
        else if (key == "ki") c.nf_ki = a.as_double();
        else if (key == "sigma_v")
        {
            // the noise term is sigma_v * N(0,1) drawn from a generator seeded by std::random_device
            // (plugins/neurofem.cpp:26-29, 303): only sigma_v = 0 has a reference result to match
            if (a.as_double() != 0.0)
                throw std::runtime_error("neurofem sigma_v != 0 adds noise from a std::random_device-seeded generator "
                                         "(plugins/neurofem.cpp:26-29), which no run of the reference reproduces; use sigma_v: 0");
        }
        // force_update / force_soma_update are stored but never read by the model's update()
    }
    else if (model == UnitModel::device_model)
    {
        // set_attribute_neuron of an out-of-tree model: the attribute initialises the state word / parameter of that
        // name; names the model does not declare are ignored (as PipelineUnit subclasses ignore unknown keys)
        const sfe_device_model_desc &d = *unit.device_model;
        if (!std::holds_alternative<double>(a.value) && !std::holds_alternative<int>(a.value)) return;
        for (uint32_t w = 0; w < d.n_state; ++w)
            if (d.state_names[w] != nullptr && key == d.state_names[w]) ln.dm_state[w] = a.as_double();
        for (uint32_t q = 0; q < d.n_params; ++q)
            if (d.param_names[q] != nullptr && key == d.param_names[q]) ln.dm_params[q] = a.as_double();
    }
    else if (model == UnitModel::hodgkin_huxley)
    {
        if (key == "m") unit.hh_m = a.as_double();
        else if (key == "n") unit.hh_n = a.as_double();
        else if (key == "h") unit.hh_h = a.as_double();
        else if (key == "current") unit.hh_i = a.as_double();
    }
}

// Smallest s in [0, 40] with w * 2^s integral for every w, or -1
int exact_shift(const std::vector<double> &weights, size_t begin, size_t end)
{
    int shift = 0;
    for (size_t i = begin; i < end; ++i)
    {
        const double w = weights[i];
        if (!std::isfinite(w)) return -1;
        while (shift <= 40)
        {
            const double scaled = std::ldexp(w, shift);
            if (scaled == std::nearbyint(scaled)) break;
            ++shift;
        }
        if (shift > 40) return -1;
    }
    return shift;
}

uint32_t pack_hop(const TileConfiguration &src, const TileConfiguration &dst)
{
    // sim_estimate_network_costs  src/chip.cpp:1127-1169
    const uint32_t dx = static_cast<uint32_t>(src.x > dst.x ? src.x - dst.x : dst.x - src.x);
    const uint32_t dy = static_cast<uint32_t>(src.y > dst.y ? src.y - dst.y : dst.y - src.y);
    if (dx > 0xfff || dy > 0xfff) throw std::runtime_error("NoC larger than 4096 tiles per dimension");
    const uint32_t east = src.x < dst.x ? 1u : 0u;
    const uint32_t north = src.y < dst.y ? 1u : 0u;
    return dx | (dy << 12) | (east << 24) | (north << 25);
}

struct CostKey
{
    double v[4];
    uint32_t per_message;
    uint32_t den_is_soma;
    bool operator<(const CostKey &o) const
    {
        const int c = std::memcmp(v, o.v, sizeof(v));
        if (c != 0) return c < 0;
        return per_message != o.per_message ? per_message < o.per_message : den_is_soma < o.den_is_soma;
    }
};

uint32_t intern_cost(std::map<CostKey, uint32_t> &index, std::vector<sfe_cost_class> &classes, const CostKey &k,
        const bool share)
{
    if (share)
    {
        auto it = index.find(k);
        if (it != index.end()) return it->second;
    }
    sfe_cost_class c{};
    c.syn_energy = k.v[0];
    c.syn_latency = k.v[1];
    c.den_energy = k.v[2];
    c.den_latency = k.v[3];
    c.per_message = k.per_message;
    c.den_is_soma = k.den_is_soma;
    classes.push_back(c);
    const uint32_t id = static_cast<uint32_t>(classes.size() - 1);
    if (share) index[k] = id;
    return id;
}

uint32_t intern_soma_class(std::map<std::string, uint32_t> &index, std::vector<sfe_soma_class> &classes,
        const sfe_soma_class &c)
{
    const std::string key(reinterpret_cast<const char *>(&c), sizeof(c));
    auto it = index.find(key);
    if (it != index.end()) return it->second;
    classes.push_back(c);
    return index[key] = static_cast<uint32_t>(classes.size() - 1);
}

void fill_arch_tables(const Architecture &arch, HostTables &out)
{
    out.tiles.clear();
    out.cores.clear();
    out.core_names.clear();
    for (const TileConfiguration &t : arch.tiles)
    {
        sfe_tile_desc d{};
        d.x = static_cast<uint32_t>(t.x);
        d.y = static_cast<uint32_t>(t.y);
        d.energy_east = t.power_metrics.energy_east_hop;
        d.energy_west = t.power_metrics.energy_west_hop;
        d.energy_south = t.power_metrics.energy_south_hop;
        d.energy_north = t.power_metrics.energy_north_hop;
        d.latency_east = t.power_metrics.latency_east_hop;
        d.latency_west = t.power_metrics.latency_west_hop;
        d.latency_south = t.power_metrics.latency_south_hop;
        d.latency_north = t.power_metrics.latency_north_hop;
        out.tiles.push_back(d);
        for (const CoreConfiguration &c : t.cores)
        {
            sfe_core_desc cd{};
            cd.id = static_cast<uint32_t>(c.address.id);
            cd.tile = static_cast<uint32_t>(t.id);
            cd.offset = static_cast<uint32_t>(c.address.offset_within_tile);
            cd.buffer_pos = c.pipeline.buffer_position;
            cd.dend_in_msg = c.pipeline.buffer_position > buffer_before_dendrite_unit ? 1 : 0;
            cd.ring = 1;
            if (!c.axon_in.empty())
            {
                // latency from unit 0 (Message.dest_axon_hw = 0, src/message.cpp:58);
                // energy from the LAST unit, which only ever counts messages when it is
                // also unit 0 (src/chip.cpp:1214-1221)
                cd.latency_axon_in = c.axon_in.front().latency_message_in;
                cd.energy_axon_in = c.axon_in.size() == 1 ? c.axon_in.front().energy_message_in : 0.0;
            }
            if (!c.axon_out.empty())
            {
                cd.latency_axon_out = c.axon_out.front().latency_message_out; // src/core.cpp:147
                cd.energy_axon_out = c.axon_out.size() == 1 ? c.axon_out.front().energy_message_out : 0.0;
            }
            out.cores.push_back(cd);
            out.core_names.push_back(std::to_string(t.id) + "." + std::to_string(cd.offset));
        }
    }
}

// Where the time-step buffer sits decides which units run in the message pipeline and which in the neuron pipeline
// (src/mapped.cpp:40-58, 171-187). Supported: inside the dendrite unit, before the soma unit, and inside the soma unit
// for neurons whose soma keeps its own double-buffered inputs (the combined "neurofem" unit; "input" somas, which
// take no input). The other combinations have no meaningful result in the reference either:
//  * inside the soma unit / before axon_out with a separate soma unit put the soma in the MESSAGE pipeline: it is
//    updated once per synaptic event; the LIF model throws at the second update of a step or a skipped step
//    (src/models.cpp:502-511), so such a chip stops at the first neuron that receives two spikes;
//  * before the dendrite unit keeps only the LAST synaptic current of a step in the buffer
//    (core.timestep_buffer[post] = pipeline_output, src/chip.cpp:759) - an order-dependent loss, not a model.
void check_buffer_position(const CoreConfiguration &c)
{
    if (c.pipeline.buffer_position == buffer_before_axon_out_unit || c.pipeline.buffer_position == buffer_before_dendrite_unit)
    {
        throw std::runtime_error("Core '" + c.name +
                "': the buffer positions before the dendrite unit (keeps only the last synaptic current of a step) and "
                "before axon_out (updates the soma once per synaptic event; the built-in LIF model throws on it) are not "
                "supported by the B200 engine (supported: dendrite+buffer_inside_unit, soma, soma+buffer_inside_unit)");
    }
}
} // namespace

int64_t HostTables::find_neuron(const std::string &group, const uint64_t offset) const
{
    const auto it = std::lower_bound(group_names.begin(), group_names.end(), group);
    if (it == group_names.end() || *it != group) return -1;
    const auto &v = group_to_device[static_cast<size_t>(it - group_names.begin())];
    return offset < v.size() ? static_cast<int64_t>(v[offset]) : -1;
}

void HostTables::finalize_view(const Architecture &arch)
{
    sfe_tables &v = view;
    std::memset(&v, 0, sizeof(v));
    v.abi_version = SFE_ABI_VERSION;
    v.noc_width = static_cast<uint32_t>(arch.noc_width_in_tiles);
    v.noc_height = static_cast<uint32_t>(arch.noc_height_in_tiles);
    v.noc_buffer_size = static_cast<uint32_t>(arch.noc_buffer_size);
    v.max_cores_per_tile = static_cast<uint32_t>(arch.max_cores_per_tile);
    v.n_tiles = static_cast<uint32_t>(tiles.size());
    v.n_cores = static_cast<uint32_t>(cores.size());
    v.n_neurons = static_cast<uint32_t>(neuron_class.size());
    v.n_axons_in = static_cast<uint32_t>(axons_in.size());
    v.n_soma_classes = static_cast<uint32_t>(soma_classes.size());
    v.n_cost_classes = static_cast<uint32_t>(cost_classes.size());
    v.n_inputs = static_cast<uint32_t>(inputs.size());
    v.input_seed_base = input_seed_base;
    v.n_poisson_cols = n_poisson_cols;
    v.noise = noise.data();
    v.noise_values = noise_values.data();
    v.n_noise = static_cast<uint32_t>(noise.size());
    v.n_noise_values = noise_values.size();
    v.u_probes = u_probes.data();
    v.n_u_probes = static_cast<uint32_t>(u_probes.size());
    v.neuron_taps = taps.empty() ? nullptr : neuron_taps.data();
    v.taps = taps.data();
    v.taps_values = taps_values.data();
    v.n_taps_units = static_cast<uint32_t>(taps.size());
    v.n_taps_values = static_cast<uint32_t>(taps_values.size());
    v.n_hh = static_cast<uint32_t>(hh.size());
    v.n_probes = static_cast<uint32_t>(probes.size());
    v.n_axons_out = axon_out_target.size();
    uint64_t nsyn = 0;
    uint32_t mapped_cores = 0;
    std::set<uint32_t> tiles_used;
    for (const sfe_core_desc &c : cores)
    {
        nsyn += c.syn_count;
        if (c.neuron_count > 0)
        {
            ++mapped_cores;
            tiles_used.insert(c.tile);
        }
    }
    v.n_synapses = nsyn;
    v.mapped_cores = mapped_cores;
    v.mapped_tiles = static_cast<uint32_t>(tiles_used.size());
    v.sync_delay = arch.ts_sync_delay_table.get(v.mapped_tiles); // src/chip.cpp:562-574
    v.tiles = tiles.data();
    v.cores = cores.data();
    v.soma_classes = soma_classes.data();
    v.cost_classes = cost_classes.data();
    v.neuron_class = neuron_class.data();
    v.neuron_aux = neuron_aux.data();
    v.neuron_bias = neuron_bias.data();
    v.neuron_potential0 = neuron_potential0.data();
    v.axon_out_begin = axon_out_begin.data();
    v.axon_out_target = axon_out_target.data();
    v.inputs = inputs.data();
    v.input_spikes = input_spikes.data();
    v.n_input_spikes = input_spikes.size();
    v.hh = hh.data();
    v.probes = probes.data();
    v.axons_in = axons_in.data();
    v.axon_src = axon_src.data();
    v.syn_weight = syn_weight.empty() ? nullptr : syn_weight.data();
    v.syn_meta = syn_meta.empty() ? nullptr : syn_meta.data();
    v.synth = synth.has_value() ? &*synth : nullptr;
    // out-of-tree device models: instance rows -> SoA blocks; neuron_aux becomes the chip-wide instance index
    device_model_view.clear();
    uint32_t first = 0;
    for (DeviceModelBlock &b : device_models)
    {
        const uint32_t n = static_cast<uint32_t>(b.neurons.size());
        b.state_init.assign(static_cast<size_t>(b.desc->n_state) * n, 0.0);
        b.params.assign(static_cast<size_t>(b.desc->n_params) * n, 0.0);
        for (uint32_t k = 0; k < n; ++k)
        {
            for (uint32_t w = 0; w < b.desc->n_state; ++w) b.state_init[static_cast<size_t>(w) * n + k] = b.state_rows[k][w];
            for (uint32_t q = 0; q < b.desc->n_params; ++q) b.params[static_cast<size_t>(q) * n + k] = b.param_rows[k][q];
            if (!b.numbered) neuron_aux[b.neurons[k]] = first + k;
        }
        b.numbered = true;
        device_model_view.push_back({b.desc, n, 0u, b.state_init.data(), b.params.data(), b.neurons.data()});
        first += n;
    }
    v.device_models = device_model_view.data();
    v.n_device_models = static_cast<uint32_t>(device_model_view.size());
    v.n_device_instances = first;
    v.n_rand_cols = n_rand_cols;
}

uint32_t count_input_units(const Architecture &arch)
{
    uint32_t n = 0;
    for (const TileConfiguration &tile : arch.tiles)
        for (const CoreConfiguration &core : tile.cores)
            for (const PipelineUnitConfiguration &u : core.pipeline_hw)
                if (is_builtin_input(u)) n += u.is_range ? static_cast<uint32_t>(u.range_last - u.range_first + 1) : 1u;
    return n;
}

#define SFE_PHASE_MARK(x) do { if (std::getenv("SFE_LOWER_PROFILE")) { auto now_ = std::chrono::steady_clock::now(); std::fprintf(stderr, "phase %s: %.3f\n", x, std::chrono::duration<double>(now_ - t_phase_).count()); t_phase_ = now_; } } while (0)
void lower_network(const Architecture &arch, const SpikingNetwork &net, HostTables &out)
{
    auto t_phase_ = std::chrono::steady_clock::now();
    out = HostTables{};
    fill_arch_tables(arch, out);
    const std::vector<const CoreConfiguration *> core_cfgs = arch.cores();
    std::vector<LCore> lcores(core_cfgs.size());
    for (size_t i = 0; i < core_cfgs.size(); ++i) lcores[i].cfg = core_cfgs[i];

    // ---- groups, in std::map (lexicographic) order -------------------------
    std::vector<const NeuronGroup *> groups;
    for (const auto &[name, g] : net.groups)
    {
        out.group_names.push_back(name);
        groups.push_back(g.get());
    }
    SFE_PHASE_MARK("0");
    // ---- map_neurons: mapping_order fixes the in-core order (src/chip.cpp:186-234)
    struct Pending
    {
        const Neuron *n;
        uint32_t group;
    };
    std::map<std::string, std::pair<uint32_t, uint32_t>> noise_files; // (path, bits) -> (offset, length)
    std::vector<Pending> order;
    for (size_t gi = 0; gi < groups.size(); ++gi)
        for (const Neuron &n : groups[gi]->neurons) order.push_back({&n, static_cast<uint32_t>(gi)});
    std::sort(order.begin(), order.end(),
            [](const Pending &a, const Pending &b) { return a.n->mapping_order < b.n->mapping_order; });
    // (group, offset) -> (core, index in core)
    std::vector<std::vector<std::pair<uint32_t, uint32_t>>> where(groups.size());
    for (size_t gi = 0; gi < groups.size(); ++gi)
        where[gi].assign(groups[gi]->neurons.size(), {UINT32_MAX, UINT32_MAX});

    for (const Pending &p : order)
    {
        const Neuron &n = *p.n;
        if (!n.core_address.has_value())
            throw HardwareMappingError("Neuron: " + n.parent_group_name + "." + std::to_string(n.offset) + " not mapped.");
        LCore &lc = lcores.at(n.core_address->id);
        const CoreConfiguration &cfg = *lc.cfg;
        // Core::map_neuron  src/core.cpp:116-168
        if (lc.neurons.size() >= cfg.pipeline.max_neurons_supported)
            throw HardwareMappingError("Error: Exceeded maximum neurons per core.");
        if (cfg.pipeline_hw.empty()) throw std::runtime_error("Error: No units defined");
        check_buffer_position(cfg);
        LNeuron ln{};
        ln.n = &n;
        ln.group = p.group;
        ln.dend = get_hw(cfg, n.dendrite_hw_name, false, true, false);
        ln.soma = get_hw(cfg, n.soma_hw_name, false, false, true);
        if (cfg.axon_out.empty()) throw std::runtime_error("Error: No axon out units defined");
        const PipelineUnitConfiguration &dend_cfg = cfg.pipeline_hw[ln.dend.family];
        const PipelineUnitConfiguration &soma_cfg = cfg.pipeline_hw[ln.soma.family];
        check_unit_shape(dend_cfg);
        check_unit_shape(soma_cfg);
        UnitState &dend = lc.units[ln.dend];
        dend.model = parse_model(dend_cfg);
        ln.dend_addr = dend.neuron_count++;
        UnitState &soma = lc.units[ln.soma];
        soma.model = parse_model(soma_cfg);
        ln.soma_addr = soma.neuron_count++;
        ln.dend_unit = &lc.units[ln.dend]; // (looked up again: inserting the soma unit cannot move it, but `dend` may alias)
        ln.soma_unit = &soma;
        if (dend.model != UnitModel::accumulator && dend.model != UnitModel::accumulator_with_delay &&
                dend.model != UnitModel::taps && dend.model != UnitModel::neurofem)
            throw std::runtime_error("Unit '" + dend_cfg.name + "' is not a dendrite model");
        const bool combined = soma.model == UnitModel::neurofem;
        if (combined)
        {
            // the combined unit double-buffers its two accumulators itself (plugins/neurofem.cpp:204-215): with the
            // buffer inside the dendrite unit, before or inside the soma unit it runs once per synaptic event in the
            // message pipeline and once per timestep in the neuron pipeline, and the chip-managed buffer carries nothing
            if (!(ln.dend == ln.soma))
                throw std::runtime_error("a 'neurofem' unit is a combined dendrite + soma unit: map the neuron's dendrite "
                                         "and soma to the same unit (dendrite_hw_name = soma_hw_name)");
            // Core::map_neuron maps a neuron to a combined unit once (src/core.cpp:155-165): one address
            --soma.neuron_count;
            ln.soma_addr = ln.dend_addr;
        }
        else if (dend.model == UnitModel::neurofem)
        {
            // e.g. the "input" neurons of arch/neurofem.yaml, whose default dendrite unit is the first one of the core:
            // they take a slot of the combined unit, which never runs for them as long as no edge reaches them
            if (soma.model != UnitModel::input || cfg.pipeline.buffer_position < buffer_before_soma_unit)
                throw std::runtime_error("a 'neurofem' unit as the dendrite of another soma model is only supported for "
                                         "'input' neurons with the buffer before or inside the soma unit");
        }
        else if (cfg.pipeline.buffer_position == buffer_inside_soma_unit && soma.model != UnitModel::input)
            throw std::runtime_error("Core '" + cfg.name + "': with the buffer inside the soma unit a separate soma unit is "
                    "updated once per synaptic event (src/mapped.cpp:52-57); the built-in LIF model throws on that "
                    "(src/models.cpp:502-511). Supported somas there: 'neurofem' (combined unit) and 'input'");
        if (dend.model == UnitModel::neurofem && ln.dend_addr >= 1024)
            throw std::runtime_error("Error: Mapped too many neurons for NeuroFEM (" + std::to_string(ln.dend_addr + 1) +
                    "> 1024)"); // plugins/neurofem.cpp:99-112
        // capacity of the built-in models' state tables (src/models.hpp:29,283)
        if (ln.dend_addr >= kLifMaxCompartments)
            throw std::out_of_range("dendrite unit '" + dend_cfg.name + "' holds at most 1024 neurons");
        if ((soma.model == UnitModel::lif && ln.soma_addr >= kLifMaxCompartments) ||
                (soma.model == UnitModel::truenorth && ln.soma_addr >= kTrueNorthMaxNeurons))
            throw std::out_of_range("soma unit '" + soma_cfg.name + "' is full");
        soma_defaults(soma_cfg, soma.model, ln.cls);
        if (soma.model == UnitModel::device_model)
        {
            std::string why;
            soma.device_model = resolve_device_model(soma_cfg, &why);
            if (soma.device_model == nullptr) throw std::runtime_error("device model '" + soma_cfg.model_info.name + "' vanished: " + why);
            const sfe_device_model_desc &d = *soma.device_model;
            ln.dm_state.assign(d.n_state, 0.0);
            ln.dm_params.assign(d.n_params, 0.0);
            for (uint32_t w = 0; w < d.n_state; ++w) ln.dm_state[w] = d.state_init != nullptr ? d.state_init[w] : 0.0;
            for (uint32_t q = 0; q < d.n_params; ++q) ln.dm_params[q] = d.param_init != nullptr ? d.param_init[q] : 0.0;
            // set_attribute_hw: the unit's own attributes in the architecture description come first (src/pipeline.cpp:87-130)
            for (const auto &[key, a] : soma_cfg.model_info.model_attributes) set_soma_attribute(ln, soma, soma.model, key, a);
            if (dend.model == UnitModel::taps)
                throw std::runtime_error("an out-of-tree soma model behind a 'taps' dendrite is not supported");
        }
        if (soma.model == UnitModel::lif && soma_cfg.model_info.model_attributes.count("noise") != 0)
        {
            // LoihiLifModel::set_attribute_hw  src/models.cpp:351-373
            if (!soma.noise_loaded)
            {
                const auto bits = soma_cfg.model_info.model_attributes.find("noise_bits");
                const int noise_bits = bits != soma_cfg.model_info.model_attributes.end() ? bits->second.as_int() : 7;
                const std::string path = soma_cfg.model_info.model_attributes.at("noise").as_string();
                auto [it, fresh] = noise_files.try_emplace(path + "\n" + std::to_string(noise_bits), 0u, 0u);
                if (fresh)
                {
                    const std::vector<double> values = read_noise_file(path, noise_bits);
                    it->second = {static_cast<uint32_t>(out.noise_values.size()), static_cast<uint32_t>(values.size())};
                    out.noise_values.insert(out.noise_values.end(), values.begin(), values.end());
                }
                soma.noise_off = it->second.first;
                soma.noise_len = it->second.second;
                soma.noise_loaded = true;
            }
            ln.cls.flags |= SFE_SOMA_NOISE;
        }
        ln.cls.dend_model = dend.model == UnitModel::accumulator ? SFE_DEND_ACCUMULATOR
                : dend.model == UnitModel::taps                  ? SFE_DEND_TAPS
                : dend.model == UnitModel::neurofem              ? SFE_DEND_NEUROFEM
                                                                 : SFE_DEND_ACCUMULATOR_DELAY;
        // a combined unit is pushed once into the neuron pipeline (src/mapped.cpp:171-187) and its output interface is
        // the soma's: no separate dendrite cost
        ln.cls.dend_in_neuron = (cfg.pipeline.buffer_position <= buffer_inside_dendrite_unit && !combined) ? 1 : 0;
        if (ln.cls.dend_in_neuron != 0)
        {
            ln.cls.dend_energy_update = unit_double(dend_cfg, "energy_update", "Dendrite");
            ln.cls.dend_latency_update = unit_double(dend_cfg, "latency_update", "Dendrite");
        }
        // MappedNeuron::set_attributes  src/mapped.cpp:113-166
        for (const auto &[key, a] : n.model_attributes)
        {
            if (is_reserved_neuron_attribute(key))
                throw std::invalid_argument("Reserved neuron attribute '" + key +
                        "' cannot be used as a model attribute. Pass it as a direct argument instead (if supported).");
            // the accumulators take no per-neuron attributes (src/models.hpp:80,133); "taps" does
            if (a.forward_to_dendrite && dend.model == UnitModel::taps) set_taps_attribute(dend, key, a);
            if (a.forward_to_soma || (combined && a.forward_to_dendrite)) set_soma_attribute(ln, soma, soma.model, key, a);
        }
        where[p.group][n.offset] = {static_cast<uint32_t>(n.core_address->id), static_cast<uint32_t>(lc.neurons.size())};
        lc.neurons.push_back(std::move(ln));
    }

    SFE_PHASE_MARK("1");
    // ---- device indices: cores in id order, neurons in in-core order --------
    uint32_t next = 0;
    for (size_t c = 0; c < lcores.size(); ++c)
    {
        out.cores[c].neuron_begin = next;
        out.cores[c].neuron_count = static_cast<uint32_t>(lcores[c].neurons.size());
        next += out.cores[c].neuron_count;
    }
    const uint32_t n_neurons = next;

    SFE_PHASE_MARK("2");
    // ---- map_connections  src/chip.cpp:334-380 --------------------------------
    size_t last_post_group = SIZE_MAX;
    for (size_t gi = 0; gi < groups.size(); ++gi)
    {
        for (const Neuron &src : groups[gi]->neurons)
        {
            const auto [src_core, src_idx] = where[gi][src.offset];
            LNeuron &pre = lcores[src_core].neurons[src_idx];
            pre.out.reserve(src.edges_out.size());
            for (const Connection &con : src.edges_out)
            {
                if (!con.post_neuron.neuron_offset.has_value())
                    throw std::invalid_argument("Post neuron doesn't specify group offset");
                // consecutive edges mostly stay in one post group: remember the last lookup
                if (last_post_group == SIZE_MAX || out.group_names[last_post_group] != con.post_neuron.group_name)
                {
                    const auto git = std::lower_bound(out.group_names.begin(), out.group_names.end(), con.post_neuron.group_name);
                    if (git == out.group_names.end() || *git != con.post_neuron.group_name) throw std::out_of_range("map::at");
                    last_post_group = static_cast<size_t>(git - out.group_names.begin());
                }
                const size_t pg = last_post_group;
                const auto [post_core, post_idx] = where[pg].at(*con.post_neuron.neuron_offset);
                LCore &pc = lcores[post_core];
                LNeuron &post = pc.neurons[post_idx];
                // get_synapse_hw_name  src/chip.cpp:308-332
                const std::string &hw_name =
                        !con.synapse_hw_name.empty() ? con.synapse_hw_name : post.n->default_synapse_hw_name;
                LConn lcn{};
                lcn.post_core = post_core;
                lcn.post_in_core = post_idx;
                // Core::get_hw + the unit checks once per (core, synapse unit name), not per connection
                auto cached = pc.synapse_hw.find(hw_name);
                if (cached == pc.synapse_hw.end())
                {
                    const UnitKey key = get_hw(*pc.cfg, hw_name, true, false, false);
                    const PipelineUnitConfiguration &syn_cfg = pc.cfg->pipeline_hw[key.family];
                    check_unit_shape(syn_cfg);
                    UnitState &fresh = pc.units[key];
                    fresh.model = parse_model(syn_cfg);
                    if (fresh.model != UnitModel::current_based)
                        throw std::runtime_error("Unit '" + syn_cfg.name + "' is not a synapse model");
                    cached = pc.synapse_hw.emplace(hw_name, std::make_pair(key, &fresh)).first;
                }
                lcn.syn = cached->second.first;
                UnitState &syn = *cached->second.second;
                lcn.syn_addr = static_cast<uint32_t>(syn.connection_count++);
                lcn.weight = 0.0;
                UnitState &dend = *post.dend_unit;
                // MappedConnection::set_attributes  src/mapped.cpp:60-89
                for (const auto &[key, a] : con.synapse_attributes)
                {
                    if (a.forward_to_synapse && (key == "w" || key == "weight")) lcn.weight = a.as_double();
                    if (a.forward_to_dendrite && dend.model == UnitModel::taps && key == "tap")
                    {
                        if (dend.synapse_to_tap.size() <= lcn.syn_addr) dend.synapse_to_tap.resize(lcn.syn_addr + 1, 0);
                        dend.synapse_to_tap[lcn.syn_addr] = a.as_int();
                    }
                    if (a.forward_to_dendrite && dend.model == UnitModel::neurofem && key == "compartment")
                    {
                        // NeuroFEMModel::set_attribute_edge  plugins/neurofem.cpp:114-136 (by SYNAPSE address)
                        const int compartment = a.as_int();
                        if (compartment < 0 || compartment > 1) throw std::runtime_error("Error: compartment must be 0 or 1");
                        if (dend.synapse_to_compartment.size() <= lcn.syn_addr) dend.synapse_to_compartment.resize(lcn.syn_addr + 1, 0);
                        dend.synapse_to_compartment[lcn.syn_addr] = static_cast<uint8_t>(compartment);
                    }
                    if (a.forward_to_dendrite && dend.model == UnitModel::accumulator_with_delay)
                    {
                        // the dendrite's delay table is addressed by the SYNAPSE address
                        if (dend.delays.size() <= lcn.syn_addr) dend.delays.resize(lcn.syn_addr + 1, 0);
                        if (key == "delay" || key == "d")
                        {
                            const int delay = a.as_int();
                            if (static_cast<size_t>(delay) > kMaxDelay) throw std::runtime_error("Error: delay > max delay\n");
                            dend.delays[lcn.syn_addr] = static_cast<uint8_t>(delay);
                        }
                    }
                }
                // InputModel::update throws the moment a non-zero current reaches an input neuron
                // (src/models.cpp:866-874); a device engine cannot raise mid-step, so the edge is refused here
                if (post.soma_unit->model == UnitModel::input &&
                        (dend.model == UnitModel::neurofem || pc.cfg->pipeline.buffer_position == buffer_inside_soma_unit))
                    throw std::runtime_error("edge into an 'input' neuron of a core whose buffer sits inside the soma unit (or "
                            "whose dendrite is a combined unit): the reference would update the input unit once per synaptic "
                            "event, advancing its spike train; not supported");
                if (lcn.weight != 0.0 && post.soma_unit->model == UnitModel::input)
                    throw std::runtime_error("Current sent to input neuron which cannot be processed (" +
                            std::to_string(lcn.weight) + "): edge " + con.pre_neuron.group_name + "." +
                            std::to_string(con.pre_neuron.neuron_offset.value_or(0)) + " -> " + con.post_neuron.group_name + "." +
                            std::to_string(*con.post_neuron.neuron_offset));
                pre.out.push_back(lcn);
            }
        }
    }

    SFE_PHASE_MARK("3");
    // ---- map_axons  src/chip.cpp:382-408, 1263-1391 -----------------------------
    // One axon per (pre-neuron, destination core). The reference visits a
    // neuron's destination cores in std::set<Core*> (heap address) order; that
    // order only shows in message ids / detailed timing. We use core id order.
    for (size_t c = 0; c < lcores.size(); ++c)
    {
        for (size_t i = 0; i < lcores[c].neurons.size(); ++i)
        {
            LNeuron &pre = lcores[c].neurons[i];
            std::set<uint32_t> dests;
            for (const LConn &k : pre.out) dests.insert(k.post_core);
            std::unordered_map<uint32_t, uint32_t> axon_of;
            for (const uint32_t d : dests)
            {
                LAxonIn a;
                a.src_core = static_cast<uint32_t>(c);
                a.src_in_core = static_cast<uint32_t>(i);
                lcores[d].axons_in.push_back(std::move(a));
                const uint32_t idx = static_cast<uint32_t>(lcores[d].axons_in.size() - 1);
                axon_of[d] = idx;
                pre.axons_out.emplace_back(d, idx);
            }
            for (size_t k = 0; k < pre.out.size(); ++k)
                lcores[pre.out[k].post_core].axons_in[axon_of[pre.out[k].post_core]].syn.emplace_back(
                        static_cast<uint32_t>(k), 0u);
        }
    }

    SFE_PHASE_MARK("4");
    // ---- emit neuron arrays ---------------------------------------------------
    out.neuron_class.resize(n_neurons);
    out.neuron_aux.assign(n_neurons, 0);
    out.neuron_bias.resize(n_neurons);
    out.neuron_potential0.resize(n_neurons);
    out.names.resize(n_neurons);
    out.group_to_device.resize(groups.size());
    for (size_t gi = 0; gi < groups.size(); ++gi) out.group_to_device[gi].resize(groups[gi]->neurons.size());
    std::map<std::string, uint32_t> class_index;
    // neurons sharing an input / HH unit, in update (= in-core) order
    for (size_t c = 0; c < lcores.size(); ++c)
        for (size_t i = 0; i < lcores[c].neurons.size(); ++i)
            lcores[c].units[lcores[c].neurons[i].soma].sharing.push_back(out.cores[c].neuron_begin + static_cast<uint32_t>(i));
    for (size_t c = 0; c < lcores.size(); ++c)
    {
        for (size_t i = 0; i < lcores[c].neurons.size(); ++i)
        {
            LNeuron &ln = lcores[c].neurons[i];
            const uint32_t dev = out.cores[c].neuron_begin + static_cast<uint32_t>(i);
            out.neuron_class[dev] = intern_soma_class(class_index, out.soma_classes, ln.cls);
            out.neuron_bias[dev] = ln.bias;
            out.neuron_potential0[dev] = ln.potential0;
            out.names[dev] = {ln.group, static_cast<uint32_t>(ln.n->offset), ln.n->log_spikes, ln.n->log_potential};
            out.group_to_device[ln.group][ln.n->offset] = dev;
            const UnitState &soma = lcores[c].units[ln.soma];
            if (soma.model == UnitModel::input)
            {
                sfe_input_desc d{};
                d.spikes_off = static_cast<uint32_t>(out.input_spikes.size());
                d.spikes_len = static_cast<uint32_t>(soma.spikes.size());
                out.input_spikes.insert(out.input_spikes.end(), soma.spikes.begin(), soma.spikes.end());
                d.share_count = static_cast<uint32_t>(soma.sharing.size());
                d.share_rank = static_cast<uint32_t>(
                        std::find(soma.sharing.begin(), soma.sharing.end(), dev) - soma.sharing.begin());
                d.rate = soma.rate;
                d.poisson = soma.poisson;
                d.unit = input_unit_ordinal(arch, c, ln.soma);
                d.poisson_col = soma.poisson > 0.0 ? out.n_poisson_cols++ : 0xFFFFFFFFu;
                out.neuron_aux[dev] = static_cast<uint32_t>(out.inputs.size());
                out.inputs.push_back(d);
            }
            else if (soma.model == UnitModel::lif && soma.noise_len > 0)
            {
                sfe_noise_desc d{};
                d.off = soma.noise_off;
                d.len = soma.noise_len;
                d.share_count = static_cast<uint32_t>(soma.sharing.size());
                d.share_rank = static_cast<uint32_t>(
                        std::find(soma.sharing.begin(), soma.sharing.end(), dev) - soma.sharing.begin());
                out.neuron_aux[dev] = static_cast<uint32_t>(out.noise.size());
                out.noise.push_back(d);
            }
            else if (soma.model == UnitModel::truenorth && ln.cls.random_mask != 0u)
                out.neuron_aux[dev] = out.n_rand_cols++; // its column in the rand() overlay: device order = the reference's update order
            else if (soma.model == UnitModel::device_model)
            {
                // one instance per neuron; neuron_aux is made chip-wide once every block is complete (below)
                size_t b = 0;
                while (b < out.device_models.size() && out.device_models[b].desc != soma.device_model) ++b;
                if (b == out.device_models.size())
                {
                    out.device_models.emplace_back();
                    out.device_models.back().desc = soma.device_model;
                }
                HostTables::DeviceModelBlock &blk = out.device_models[b];
                out.neuron_aux[dev] = static_cast<uint32_t>(blk.neurons.size());
                blk.neurons.push_back(dev);
                blk.state_rows.push_back(ln.dm_state);
                blk.param_rows.push_back(ln.dm_params);
            }
            else if (soma.model == UnitModel::hodgkin_huxley)
            {
                if (soma.sharing.size() != 1)
                    throw std::runtime_error("hodgkin_huxley keeps one neuron of state per hardware unit "
                                             "(plugins/hodgkin_huxley.cpp:94-118); map one neuron per unit");
                out.neuron_aux[dev] = static_cast<uint32_t>(out.hh.size());
                out.hh.push_back({soma.hh_m, soma.hh_n, soma.hh_h, soma.hh_i});
            }
        }
    }
    // "taps" dendrites keep ONE line of state per hardware unit, whatever neuron is being updated (SURVEY Appendix
    // B-6; the shipped snn/dendrite.yaml maps its input neurons to the same unit as the neuron with the line): one
    // descriptor per unit, shared by the neurons mapped to it
    for (size_t c = 0; c < lcores.size(); ++c)
    {
        std::map<UnitKey, uint32_t> unit_index;
        for (size_t i = 0; i < lcores[c].neurons.size(); ++i)
        {
            const LNeuron &ln = lcores[c].neurons[i];
            const UnitState &dend = *ln.dend_unit;
            if (dend.model != UnitModel::taps) continue;
            if (out.neuron_taps.empty()) out.neuron_taps.assign(n_neurons, 0xFFFFFFFFu);
            auto [it, fresh] = unit_index.try_emplace(ln.dend, static_cast<uint32_t>(out.taps.size()));
            if (fresh)
            {
                sfe_taps_desc d{};
                d.n_taps = static_cast<uint32_t>(dend.n_taps);
                d.const_off = static_cast<uint32_t>(out.taps_values.size());
                for (size_t k = 0; k < dend.n_taps; ++k) out.taps_values.push_back(dend.time_constants.at(k));
                for (size_t k = 0; k + 1 < dend.n_taps; ++k) out.taps_values.push_back(dend.space_constants.at(k));
                out.taps.push_back(d);
            }
            out.neuron_taps[out.cores[c].neuron_begin + i] = it->second;
        }
    }
    // potential probes in trace order: lexicographic group, offset (src/chip.cpp:1632-1662)
    for (size_t gi = 0; gi < groups.size(); ++gi)
        for (const Neuron &n : groups[gi]->neurons)
            if (n.log_potential) out.probes.push_back(out.group_to_device[gi][n.offset]);
    // model-defined traces, same order: the only built-in producer is LIF's `u` for log_u neurons
    // (LoihiLifModel::get_neuron_traces, src/models.cpp:653-662)
    for (size_t gi = 0; gi < groups.size(); ++gi)
        for (const Neuron &n : groups[gi]->neurons)
        {
            const uint32_t dev = out.group_to_device[gi][n.offset];
            const sfe_soma_class &cls = out.soma_classes[out.neuron_class[dev]];
            if (cls.model == SFE_SOMA_LIF && (cls.flags & SFE_SOMA_LOG_U) != 0) out.u_probes.push_back(dev);
        }

    SFE_PHASE_MARK("5");
    // ---- emit axons-out (per neuron, in sending order) -------------------------
    std::vector<uint32_t> core_axon_base(lcores.size());
    uint32_t axon_total = 0;
    for (size_t c = 0; c < lcores.size(); ++c)
    {
        core_axon_base[c] = axon_total;
        out.cores[c].axon_in_begin = axon_total;
        out.cores[c].axon_in_count = static_cast<uint32_t>(lcores[c].axons_in.size());
        axon_total += out.cores[c].axon_in_count;
    }
    out.axon_out_begin.assign(n_neurons + 1, 0);
    for (size_t c = 0; c < lcores.size(); ++c)
        for (size_t i = 0; i < lcores[c].neurons.size(); ++i)
        {
            const uint32_t dev = out.cores[c].neuron_begin + static_cast<uint32_t>(i);
            out.axon_out_begin[dev] = static_cast<uint32_t>(out.axon_out_target.size());
            for (const auto &[d, idx] : lcores[c].neurons[i].axons_out)
                out.axon_out_target.push_back(core_axon_base[d] + idx);
        }
    out.axon_out_begin[n_neurons] = static_cast<uint32_t>(out.axon_out_target.size());

    SFE_PHASE_MARK("6");
    // ---- emit axons-in + synapses, per destination core -------------------------
    out.axons_in.resize(axon_total);
    out.axon_src.resize(axon_total);
    std::map<CostKey, uint32_t> cost_index;
    uint64_t syn_total = 0;
    for (size_t c = 0; c < lcores.size(); ++c)
        for (const LAxonIn &a : lcores[c].axons_in) syn_total += a.syn.size();
    out.syn_weight.reserve(syn_total);
    out.syn_meta.reserve(syn_total);
    for (size_t c = 0; c < lcores.size(); ++c)
    {
        LCore &lc = lcores[c];
        sfe_core_desc &cd = out.cores[c];
        cd.syn_begin = out.syn_weight.size();
        uint32_t max_delay = 0;
        bool core_has_taps = false, core_fixed_slots = false, core_has_delay = false;
        const size_t n_families = lc.cfg->pipeline_hw.size();
        std::vector<CostKey> pair_key(n_families * n_families);
        std::vector<uint8_t> pair_known(n_families * n_families, 0);
        // dendrite delay tables of the core's units, by family/instance (looked up once per unit, not per synapse)
        const UnitState *last_dend = nullptr;
        UnitKey last_dend_key{-1, -1};
        for (size_t ai = 0; ai < lc.axons_in.size(); ++ai)
        {
            const LAxonIn &a = lc.axons_in[ai];
            const LNeuron &pre = lcores[a.src_core].neurons[a.src_in_core];
            sfe_axon_in &rec = out.axons_in[core_axon_base[c] + ai];
            rec.syn_off = static_cast<uint32_t>(out.syn_weight.size() - cd.syn_begin);
            rec.syn_count = static_cast<uint32_t>(a.syn.size());
            rec.hop = pack_hop(arch.tiles[out.cores[a.src_core].tile], arch.tiles[cd.tile]);
            out.axon_src[core_axon_base[c] + ai] = out.cores[a.src_core].neuron_begin + a.src_in_core;
            bool uniform = true;
            CostKey first{};
            CostKey total{};
            total.per_message = 1;
            for (size_t s = 0; s < a.syn.size(); ++s)
            {
                const LConn &k = pre.out[a.syn[s].first];
                const LNeuron &post = lc.neurons[k.post_in_core];
                // default costs depend on the (synapse unit, dendrite unit) families only: looked up once per pair
                const size_t pair = static_cast<size_t>(k.syn.family) * n_families + static_cast<size_t>(post.dend.family);
                if (!pair_known[pair])
                {
                    const PipelineUnitConfiguration &syn_cfg = lc.cfg->pipeline_hw[k.syn.family];
                    const PipelineUnitConfiguration &den_cfg = lc.cfg->pipeline_hw[post.dend.family];
                    CostKey fresh{};
                    fresh.v[0] = unit_double(syn_cfg, "energy_process_spike", "Synapse");
                    fresh.v[1] = unit_double(syn_cfg, "latency_process_spike", "Synapse");
                    const bool den_combined = den_cfg.model_info.plugin_library_path.has_value() && den_cfg.model_info.name == "neurofem";
                    if (cd.dend_in_msg != 0 && den_combined)
                    {
                        // a combined dendrite + soma unit in the message pipeline: its output interface is the soma's, and
                        // with the status unset an event costs one neuron access (src/pipeline.hpp:631-667)
                        fresh.v[2] = unit_double(den_cfg, "energy_access_neuron", "Soma");
                        fresh.v[3] = unit_double(den_cfg, "latency_access_neuron", "Soma");
                        fresh.den_is_soma = 1;
                    }
                    else if (cd.dend_in_msg != 0)
                    {
                        fresh.v[2] = unit_double(den_cfg, "energy_update", "Dendrite");
                        fresh.v[3] = unit_double(den_cfg, "latency_update", "Dendrite");
                    }
                    pair_key[pair] = fresh;
                    pair_known[pair] = 1;
                }
                const CostKey key = pair_key[pair];
                if (s == 0) first = key;
                else if (std::memcmp(first.v, key.v, sizeof(key.v)) != 0 || first.den_is_soma != key.den_is_soma) uniform = false;
                if (s == 0) total.den_is_soma = key.den_is_soma;
                else if (total.den_is_soma != key.den_is_soma)
                    throw std::runtime_error("an axon whose synapses reach both a combined (dendrite + soma) unit and a plain "
                                             "dendrite unit is not supported");
                // per-message totals in the reference's summation order
                // (src/chip.cpp:766-789: per synapse (0.0 + syn) + den, then message += that)
                total.v[0] += key.v[0];
                total.v[2] += key.v[2];
                total.v[1] += (0.0 + key.v[1]) + key.v[3];
                if (last_dend == nullptr || !(last_dend_key == post.dend))
                {
                    last_dend = &lc.units[post.dend];
                    last_dend_key = post.dend;
                }
                const UnitState &dend = *last_dend;
                uint32_t delay = 0;
                if (dend.model == UnitModel::accumulator_with_delay && k.syn_addr < dend.delays.size())
                    delay = dend.delays[k.syn_addr];
                if (dend.model == UnitModel::accumulator_with_delay && delay > 0) core_has_delay = true;
                if (dend.model == UnitModel::neurofem)
                {
                    // the compartment (0: u1, 1: u2) rides in the delay field: the dendrite slot of the synapse
                    delay = k.syn_addr < dend.synapse_to_compartment.size() ? dend.synapse_to_compartment[k.syn_addr] : 0u;
                    core_fixed_slots = true;
                }
                max_delay = std::max(max_delay, delay);
                if (dend.model == UnitModel::taps)
                {
                    // MultiTapModel1D::input_current  src/models.cpp:215-235: tap of the synapse address, default 0;
                    // a tap outside the line throws when the first current arrives — refused here
                    const int tap = k.syn_addr < dend.synapse_to_tap.size() ? dend.synapse_to_tap[k.syn_addr] : 0;
                    if (tap < 0 || static_cast<size_t>(tap) >= dend.n_taps)
                        throw std::logic_error("Tap should be >= 0 and less than taps.\n");
                    delay = static_cast<uint32_t>(tap); // rides in the delay field of syn_meta
                    core_has_taps = true;
                }
                out.syn_weight.push_back(k.weight);
                out.syn_meta.push_back(k.post_in_core | (delay << 16));
            }
            rec.cost_class = uniform ? intern_cost(cost_index, out.cost_classes, first, true)
                                     : intern_cost(cost_index, out.cost_classes, total, false);
        }
        cd.syn_count = out.syn_weight.size() - cd.syn_begin;
        cd.ring = max_delay + 1;
        for (const LNeuron &ln : lc.neurons)
            if (ln.soma_unit->model == UnitModel::neurofem) core_fixed_slots = true;
        if (core_fixed_slots)
        {
            if (core_has_delay || core_has_taps)
                throw std::runtime_error("Core '" + lc.cfg->name + "': 'neurofem' units cannot share a core with synapses that "
                                         "carry accumulator delays or dendritic taps");
            cd.fixed_slots = 1;
            cd.ring = 2; // u1 and u2
        }
        if (cd.syn_count > 0xffffffffull) throw std::runtime_error("more than 2^32 synapses on one core");
        // exactness certificate: every weight is k * 2^-s and the
        // per-post-neuron sums stay small enough for integer smem accumulators
        const int shift = exact_shift(out.syn_weight, cd.syn_begin, cd.syn_begin + cd.syn_count);
        cd.acc_mode = SFE_ACC_ORDERED;
        cd.weight_shift = 0;
        if (shift >= 0 && !core_has_taps) // a tap line adds onto evolving fp64 state: in-order adds only
        {
            std::vector<double> sum_abs(lc.neurons.size(), 0.0);
            std::vector<uint64_t> fan_in(lc.neurons.size(), 0);
            for (uint64_t s = cd.syn_begin; s < cd.syn_begin + cd.syn_count; ++s)
            {
                const uint32_t post = SFE_SYN_POST(out.syn_meta[s]);
                sum_abs[post] += std::fabs(std::ldexp(out.syn_weight[s], shift));
                fan_in[post] += 1;
            }
            double worst_sum = 0.0;
            uint64_t worst_fan = 0;
            for (size_t i = 0; i < sum_abs.size(); ++i)
            {
                worst_sum = std::max(worst_sum, sum_abs[i]);
                worst_fan = std::max(worst_fan, fan_in[i]);
            }
            cd.weight_shift = shift;
            if (worst_sum < 65536.0 && worst_fan < 32768) cd.acc_mode = SFE_ACC_PACKED17;
            else if (worst_sum < 524288.0 && worst_fan < 4096) cd.acc_mode = SFE_ACC_PACKED32;
            else if (worst_sum < 2147483648.0) cd.acc_mode = SFE_ACC_DUAL32;
        }
    }
    if (out.cost_classes.empty()) out.cost_classes.push_back(sfe_cost_class{});
    if (out.soma_classes.empty()) out.soma_classes.push_back(sfe_soma_class{});
    SFE_PHASE_MARK("end");
    out.finalize_view(arch);
}

void lower_synthetic(const Architecture &arch, const SynthRequest &req, const bool materialize, HostTables &out)
{
    out = HostTables{};
    fill_arch_tables(arch, out);
    const sfe_synth_spec &s = req.spec;
    const std::vector<const CoreConfiguration *> core_cfgs = arch.cores();
    const uint32_t C = s.cores, P = s.neurons_per_core, D = s.dest_cores, S = s.syn_per_axon;
    if (C == 0 || C > core_cfgs.size()) throw HardwareMappingError("synthetic network needs more cores than the architecture has");
    if (P == 0 || (P & (P - 1)) != 0 || P > 4096) throw std::invalid_argument("synthetic: neurons_per_core must be a power of two <= 4096");
    if (D > C || S > P) throw std::invalid_argument("synthetic: dest_cores <= cores and syn_per_axon <= neurons_per_core");
    if (s.max_delay > kMaxDelay) throw std::runtime_error("Error: delay > max delay\n");
    if (static_cast<uint64_t>(P) * D * S > 0xffffffffull) throw std::runtime_error("more than 2^32 synapses on one core");

    // All cores share one hardware description in practice; resolve units per core
    // anyway so that errors surface exactly as for a described network.
    sfe_soma_class cls{};
    sfe_cost_class cost{};
    for (uint32_t c = 0; c < C; ++c)
    {
        const CoreConfiguration &cfg = *core_cfgs[c];
        if (P > cfg.pipeline.max_neurons_supported) throw HardwareMappingError("Error: Exceeded maximum neurons per core.");
        check_buffer_position(cfg);
        const UnitKey dk = get_hw(cfg, req.dendrite_hw_name, false, true, false);
        const UnitKey sk = get_hw(cfg, req.soma_hw_name, false, false, true);
        const UnitKey yk = get_hw(cfg, req.synapse_hw_name, true, false, false);
        const PipelineUnitConfiguration &den_cfg = cfg.pipeline_hw[dk.family];
        const PipelineUnitConfiguration &soma_cfg = cfg.pipeline_hw[sk.family];
        const PipelineUnitConfiguration &syn_cfg = cfg.pipeline_hw[yk.family];
        const UnitModel dm = parse_model(den_cfg);
        if (parse_model(soma_cfg) != UnitModel::lif) throw std::runtime_error("synthetic: soma unit must be leaky_integrate_fire");
        if (parse_model(syn_cfg) != UnitModel::current_based) throw std::runtime_error("synthetic: synapse unit must be current_based");
        if (dm != UnitModel::accumulator && dm != UnitModel::accumulator_with_delay)
            throw std::runtime_error("synthetic: dendrite unit must be an accumulator");
        if (P > kLifMaxCompartments) throw std::out_of_range("soma unit '" + soma_cfg.name + "' is full");
        sfe_soma_class k{};
        soma_defaults(soma_cfg, UnitModel::lif, k);
        k.threshold = s.threshold;
        k.reset = s.reset;
        k.leak = s.leak_decay;
        k.dend_model = dm == UnitModel::accumulator ? SFE_DEND_ACCUMULATOR : SFE_DEND_ACCUMULATOR_DELAY;
        k.dend_in_neuron = cfg.pipeline.buffer_position <= buffer_inside_dendrite_unit ? 1 : 0;
        if (k.dend_in_neuron != 0)
        {
            k.dend_energy_update = unit_double(den_cfg, "energy_update", "Dendrite");
            k.dend_latency_update = unit_double(den_cfg, "latency_update", "Dendrite");
        }
        sfe_cost_class q{};
        q.syn_energy = unit_double(syn_cfg, "energy_process_spike", "Synapse");
        q.syn_latency = unit_double(syn_cfg, "latency_process_spike", "Synapse");
        if (out.cores[c].dend_in_msg != 0)
        {
            q.den_energy = unit_double(den_cfg, "energy_update", "Dendrite");
            q.den_latency = unit_double(den_cfg, "latency_update", "Dendrite");
        }
        if (c == 0)
        {
            cls = k;
            cost = q;
        }
        else if (std::memcmp(&cls, &k, sizeof(k)) != 0 || std::memcmp(&cost, &q, sizeof(q)) != 0)
            throw std::runtime_error("synthetic: the mapped cores must share one hardware description");
        // without a delay line the dendrite cannot hold delays
        if (s.max_delay > 0 && dm != UnitModel::accumulator_with_delay)
            throw std::runtime_error("synthetic: max_delay > 0 needs an accumulator_with_delay dendrite");
    }
    out.soma_classes.push_back(cls);
    out.cost_classes.push_back(cost);

    const uint32_t N = C * P;
    out.group_names = {"pop"};
    out.group_to_device.resize(1);
    out.group_to_device[0].resize(N);
    out.names.resize(N);
    out.neuron_class.assign(N, 0);
    out.neuron_aux.assign(N, 0);
    out.neuron_bias.resize(N);
    out.neuron_potential0.assign(N, 0.0);
    for (uint32_t n = 0; n < N; ++n)
    {
        out.neuron_bias[n] = sfe_synth_has_bias(&s, n) ? s.bias : 0.0;
        out.names[n] = {0, n, s.log_spikes != 0, n < s.log_potential_n};
        out.group_to_device[0][n] = n;
        if (n < s.log_potential_n) out.probes.push_back(n);
    }
    // Sources of destination core d, ascending core id: {(d - k) mod C, k < D}.
    // rank_of(d, src) = position of src in that sorted list.
    auto rank_of = [&](uint32_t d, uint32_t src) -> uint32_t {
        if (D == C) return src;
        if (d + 1 >= D) return src - (d + 1 - D);          // no wrap: sources d-D+1 .. d
        const uint32_t wrapped = D - 1 - d;                // sources 0..d and C-wrapped..C-1
        return src <= d ? src : (d + 1) + (src - (C - wrapped));
    };
    const uint32_t axons_per_core = P * D;
    out.axons_in.resize(static_cast<size_t>(C) * axons_per_core);
    out.axon_src.resize(out.axons_in.size());
    out.axon_out_begin.resize(N + 1);
    out.axon_out_target.resize(static_cast<size_t>(N) * D);
    for (uint32_t c = 0; c < C; ++c)
    {
        sfe_core_desc &cd = out.cores[c];
        cd.neuron_begin = c * P;
        cd.neuron_count = P;
        cd.axon_in_begin = c * axons_per_core;
        cd.axon_in_count = axons_per_core;
        cd.syn_begin = static_cast<uint64_t>(c) * axons_per_core * S;
        cd.syn_count = static_cast<uint64_t>(axons_per_core) * S;
        cd.ring = s.max_delay + 1;
        cd.weight_shift = 0;
        cd.acc_mode = SFE_ACC_DUAL32; // refined below / re-certified on the device
    }
    std::vector<std::pair<uint32_t, uint32_t>> order(D); // (destination core, k)
    for (uint32_t n = 0; n < N; ++n)
    {
        const uint32_t home = n / P;
        out.axon_out_begin[n] = n * D;
        // The reference visits a neuron's destination cores in std::set<Core*> order,
        // which in practice is ascending (tile, core) order (checked against the
        // reference's own mapping dump); messages are sent in that order.
        for (uint32_t k = 0; k < D; ++k) order[k] = {sfe_synth_dest_core(&s, home, k), k};
        std::sort(order.begin(), order.end());
        for (uint32_t q = 0; q < D; ++q)
        {
            const uint32_t d = order[q].first;
            const uint32_t slot = rank_of(d, home) * P + (n % P);
            const uint32_t axon_id = d * axons_per_core + slot;
            out.axon_out_target[static_cast<size_t>(n) * D + q] = axon_id;
            sfe_axon_in &rec = out.axons_in[axon_id];
            rec.syn_off = slot * S;
            rec.syn_count = S;
            rec.hop = pack_hop(arch.tiles[out.cores[home].tile], arch.tiles[out.cores[d].tile]);
            rec.cost_class = 0;
            out.axon_src[axon_id] = n;
        }
    }
    out.axon_out_begin[N] = N * D;

    out.synth = s;
    if (materialize)
    {
        const uint64_t total = static_cast<uint64_t>(C) * axons_per_core * S;
        out.syn_weight.resize(total);
        out.syn_meta.resize(total);
        for (uint32_t n = 0; n < N; ++n)
        {
            const uint32_t home = n / P;
            for (uint32_t k = 0; k < D; ++k)
            {
                const uint64_t axon = static_cast<uint64_t>(n) * D + k;
                const uint32_t d = sfe_synth_dest_core(&s, home, k);
                const uint32_t axon_id = d * axons_per_core + rank_of(d, home) * P + (n % P);
                const uint64_t base = out.cores[d].syn_begin + out.axons_in[axon_id].syn_off;
                const sfe_synth_axon ap = sfe_synth_axon_params(&s, axon);
                for (uint32_t j = 0; j < S; ++j)
                {
                    const uint64_t syn = axon * S + j;
                    out.syn_weight[base + j] = static_cast<double>(sfe_synth_weight(&s, syn));
                    out.syn_meta[base + j] = sfe_synth_post(&s, &ap, j) | (sfe_synth_delay(&s, syn) << 16);
                }
            }
        }
    }
    // exactness certificate from the generator's bounds: integer weights, fan-in
    // per (post, core) is at most P*D (every source neuron of every source core
    // hits a post at most once per axon)
    {
        const double wmax = static_cast<double>(std::max(std::abs(s.w_min), std::abs(s.w_max)));
        const double fan_bound = static_cast<double>(P) * D; // axons into the core; each hits a post <= once
        const uint32_t mode = (fan_bound < 32768.0 && fan_bound * wmax < 65536.0)  ? SFE_ACC_PACKED17
                : (fan_bound < 4096.0 && fan_bound * wmax < 524288.0)              ? SFE_ACC_PACKED32
                : (fan_bound * wmax < 2147483648.0)                                ? SFE_ACC_DUAL32
                                                                                   : SFE_ACC_ORDERED;
        for (uint32_t c = 0; c < C; ++c) out.cores[c].acc_mode = mode;
    }
    out.finalize_view(arch);
}

// Post-load MappedNeuron::set_attributes (src/mapped.cpp:113-166) for one numeric attribute: the reference forwards
// it to the neuron's soma (and dendrite) unit, whose set_attribute_neuron knows a model-specific key set
// (src/models.cpp:375-439 LIF, 664-722 TrueNorth, 832-853 input; plugins/hodgkin_huxley.cpp:94-114;
// plugins/neurofem.cpp:138-190). Keys a model does not know are ignored there and here; keys it honours but this
// engine cannot change once the tables are on the device raise instead of being dropped silently.
PatchKind patch_neuron_attribute(HostTables &t, const uint32_t neuron, const std::string &name, const double value)
{
    if (neuron >= t.neuron_class.size()) throw std::out_of_range("neuron index out of range");
    sfe_soma_class c = t.soma_classes[t.neuron_class[neuron]];
    auto set_flag = [&](const uint32_t flag) { c.flags = value != 0.0 ? (c.flags | flag) : (c.flags & ~flag); };
    auto frozen = [&](const char *why) -> PatchKind {
        throw std::runtime_error("attribute '" + name + "' cannot be changed after load(): " + why);
    };
    switch (c.model)
    {
    case SFE_SOMA_LIF:
        if (name == "bias") { t.neuron_bias[neuron] = value; return PatchKind::bias; }
        if (name == "potential") { t.neuron_potential0[neuron] = value; return PatchKind::potential; }
        if (name == "threshold") c.threshold = value;
        else if (name == "reverse_threshold") c.reverse_threshold = value;
        else if (name == "reset") c.reset = value;
        else if (name == "reverse_reset") c.reverse_reset = value;
        else if (name == "leak_decay") c.leak = value;
        else if (name == "input_decay") c.input_decay = value;
        else if (name == "refractory_delay") c.refractory_delay = static_cast<int32_t>(value);
        else if (name == "force_update" || name == "force_update_every_timestep") set_flag(SFE_SOMA_FORCE_UPDATE);
        else if (name == "log_u") return frozen("the model-defined trace columns are fixed when the network is loaded");
        else return PatchKind::ignored; // e.g. "leak" (a TrueNorth key)
        break;
    case SFE_SOMA_TRUENORTH:
        if (name == "bias") { t.neuron_bias[neuron] = value; return PatchKind::bias; }
        if (name == "threshold") c.threshold = value;
        else if (name == "reverse_threshold") c.reverse_threshold = value;
        else if (name == "reset") c.reset = value;
        else if (name == "reverse_reset") c.reverse_reset = value;
        else if (name == "leak") c.leak = value;
        else if (name == "force_update" || name == "force_update_every_timestep") set_flag(SFE_SOMA_FORCE_UPDATE);
        else if (name == "leak_towards_zero") set_flag(SFE_SOMA_LEAK_TOWARDS_ZERO);
        else if (name == "random_mask")
        {
            if (value < 0.0) throw std::invalid_argument("random_mask < 0; must be unsigned.");
            if ((value != 0.0) != (c.random_mask != 0u))
                return frozen("switching the threshold jitter on or off changes the chip's rand() stream (one draw per such neuron and step)");
            c.random_mask = static_cast<uint32_t>(value);
        }
        else return PatchKind::ignored; // e.g. "leak_decay" (a LIF key)
        break;
    case SFE_SOMA_NEUROFEM:
        if (name == "bias") { t.neuron_bias[neuron] = value; return PatchKind::bias; }
        if (name == "threshold") c.threshold = value;
        else if (name == "reset") c.reset = value;
        else if (name == "lambda_d") c.input_decay = value;
        else if (name == "lambda_v") c.leak = value;
        else if (name == "dt") c.nf_dt = value;
        else if (name == "kp") c.nf_kp = value;
        else if (name == "ki") c.nf_ki = value;
        else if (name == "sigma_v")
        {
            if (value != 0.0) return frozen("sigma_v != 0 adds std::random_device noise that no reference run reproduces");
            return PatchKind::ignored;
        }
        else return PatchKind::ignored;
        break;
    case SFE_SOMA_INPUT:
        if (name == "rate" || name == "poisson" || name == "spikes")
            return frozen("an 'input' unit keeps one spike source per hardware unit, lowered at load()");
        return PatchKind::ignored;
    case SFE_SOMA_HH:
        if (name == "m" || name == "n" || name == "h" || name == "current")
            return frozen("the Hodgkin-Huxley state of a unit is initialised at load()");
        return PatchKind::ignored;
    default:
        return frozen("the neuron's soma is an out-of-tree device model");
    }
    // split the class: find an identical one or append
    for (size_t i = 0; i < t.soma_classes.size(); ++i)
    {
        if (std::memcmp(&t.soma_classes[i], &c, sizeof(c)) == 0)
        {
            t.neuron_class[neuron] = static_cast<uint32_t>(i);
            return PatchKind::classes;
        }
    }
    t.soma_classes.push_back(c);
    t.neuron_class[neuron] = static_cast<uint32_t>(t.soma_classes.size() - 1);
    t.view.soma_classes = t.soma_classes.data();
    t.view.n_soma_classes = static_cast<uint32_t>(t.soma_classes.size());
    return PatchKind::classes;
}

} // namespace sfe
