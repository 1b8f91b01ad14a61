// pymodule.cpp — pybind11 module `sanafecpp_b200`: the reference's Python surface
// (src/pymodule.cpp:850-1213) over the C ABI of this engine. Same names, keyword
// arguments and result dictionary as `sanafecpp`, so `sanafe/__init__.py`'s
// `from sanafecpp import *` can be pointed at it:
//   load_arch(path), load_net(path, arch, use_netlist_format=False)
//   SpikingChip(arch).load(net, overwrite=False)
//   SpikingChip.sim(timesteps=1, timing_model="detailed", processing_threads=0,
//                   scheduler_threads=0, spike_trace=None, potential_trace=None,
//                   neuron_trace=None, perf_trace=None, message_trace=None,
//                   write_trace_headers=True) -> dict
//   SpikingChip.reset(), SpikingChip.get_power()
// The GIL is released while the device runs (src/pymodule.cpp:629-652).
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <chrono>
#include <condition_variable>
#include <fstream>
#include <mutex>
#include <thread>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "handles.hpp"
#include "sanafe_b200.h"
#include "sfe_device_model.h"

namespace py = pybind11;

namespace
{
struct ArchHandle
{
    sfe_arch *h{nullptr};
    ~ArchHandle() { sfe_arch_free(h); }
};
struct NetHandle
{
    sfe_net *h{nullptr};
    ~NetHandle() { sfe_net_free(h); }
};

[[noreturn]] void raise_last()
{
    // the exception class the reference would have thrown (pybind11's default translation then matches:
    // invalid_argument -> ValueError, out_of_range -> IndexError, HardwareMappingError registered below)
    const std::string msg = sfe_last_error();
    switch (sfe_last_error_kind())
    {
    case SFE_ERROR_INVALID_ARGUMENT: throw std::invalid_argument(msg);
    case SFE_ERROR_OUT_OF_RANGE: throw std::out_of_range(msg);
    case SFE_ERROR_HARDWARE_MAPPING: throw sfe::HardwareMappingError(msg);
    default: throw std::runtime_error(msg);
    }
}

// ---- Python values <-> attributes (src/pymodule.cpp:119-248) -------------------------------------
// Note the reference's narrowing: a Python float becomes a C float first, and a
// Python bool is an int.
sfe::Attr to_attr(const py::object &value)
{
    sfe::Attr a;
    bool is_int = false, is_float = false;
    if (py::hasattr(value, "dtype"))
    {
        const std::string kind = value.attr("dtype").attr("kind").cast<std::string>();
        is_int = kind == "i" || kind == "u";
        is_float = kind == "f";
    }
    else
    {
        is_int = py::isinstance<py::int_>(value);
        is_float = py::isinstance<py::float_>(value);
    }
    if (py::isinstance<py::str>(value)) a.value = value.cast<std::string>();
    else if (is_int) a.value = value.cast<int>();
    else if (is_float) a.value = static_cast<double>(value.cast<float>());
    else if (py::isinstance<py::dict>(value))
    {
        std::vector<sfe::Attr> list;
        for (auto item : value.cast<py::dict>())
        {
            sfe::Attr e = to_attr(py::reinterpret_borrow<py::object>(item.second));
            e.name = item.first.cast<std::string>();
            list.push_back(std::move(e));
        }
        a.value = std::move(list);
    }
    else if (py::isinstance<py::iterable>(value))
    {
        std::vector<sfe::Attr> list;
        for (auto it = py::iter(value); it != py::iterator::sentinel(); ++it)
            list.push_back(to_attr(py::reinterpret_borrow<py::object>(*it)));
        a.value = std::move(list);
    }
    else throw std::invalid_argument("Error: dict has unsupported type");
    return a;
}

sfe::AttrMap to_attr_map(const py::dict &d, bool to_synapse = true, bool to_dendrite = true, bool to_soma = true)
{
    sfe::AttrMap m;
    for (auto kv : d)
    {
        const std::string key = kv.first.cast<std::string>();
        sfe::Attr a = to_attr(py::reinterpret_borrow<py::object>(kv.second));
        a.forward_to_synapse = to_synapse;
        a.forward_to_dendrite = to_dendrite;
        a.forward_to_soma = to_soma;
        a.name = key;
        m[key] = std::move(a);
    }
    return m;
}

sfe::NeuronGroup::AttrLists to_attr_lists(const py::dict &d) // src/pymodule.cpp:90-117
{
    sfe::NeuronGroup::AttrLists lists;
    for (auto kv : d)
    {
        const std::string key = kv.first.cast<std::string>();
        if (!py::isinstance<py::iterable>(kv.second))
            throw std::invalid_argument("Error: Each attribute must be provided as a 1D list/array of values. "
                                        "Multi-dimensional arrays must be flattened in C-order and storing channels as "
                                        "the last dim");
        std::vector<sfe::Attr> &list = lists[key];
        for (auto it = py::iter(kv.second); it != py::iterator::sentinel(); ++it)
            list.push_back(to_attr(py::reinterpret_borrow<py::object>(*it)));
    }
    return lists;
}

py::object from_attr(const sfe::Attr &a)
{
    if (const bool *b = std::get_if<bool>(&a.value)) return py::cast(*b);
    if (const int *i = std::get_if<int>(&a.value)) return py::cast(*i);
    if (const double *d = std::get_if<double>(&a.value)) return py::cast(*d);
    if (const std::string *s = std::get_if<std::string>(&a.value)) return py::cast(*s);
    const std::vector<sfe::Attr> &list = a.as_list();
    if (!list.empty() && list[0].name.has_value())
    {
        py::dict d;
        for (const sfe::Attr &e : list)
        {
            if (!e.name.has_value()) throw std::runtime_error("Error: Sub-attribute is unnamed.\n");
            d[py::str(*e.name)] = from_attr(e);
        }
        return d;
    }
    py::list l;
    for (const sfe::Attr &e : list) l.append(from_attr(e));
    return l;
}

py::dict from_attr_map(const sfe::AttrMap &m)
{
    py::dict d;
    for (const auto &[key, a] : m) d[py::str(key)] = from_attr(a);
    return d;
}

// ---- builder objects (src/pymodule.cpp:873-1173). References into a description keep its handle alive. ----
struct GroupRef
{
    std::shared_ptr<NetHandle> owner;
    sfe::NeuronGroup *g;
};
struct NeuronRef
{
    std::shared_ptr<NetHandle> owner;
    sfe::Neuron *n;
};

sfe::SpikingNetwork &described(const std::shared_ptr<NetHandle> &net)
{
    if (net->h == nullptr || !net->h->net) throw std::runtime_error("this network has no object description (generated networks are built on the device)");
    return *net->h->net;
}

std::string group_info(const sfe::NeuronGroup &g)
{
    return "sanafe::NeuronGroup(name=" + g.name + " neurons=" + std::to_string(g.neurons.size()) + ")";
}
std::string neuron_info(const sfe::Neuron &n)
{
    std::string s = "sanafe::Neuron(nid=" + n.parent_group_name + "." + std::to_string(n.offset) + " connections_out=" +
            std::to_string(n.edges_out.size());
    if (n.core_address.has_value())
        s += " core=" + std::to_string(n.core_address->parent_tile_id) + "." + std::to_string(n.core_address->offset_within_tile);
    return s + ")";
}

class Chip;
// A mapped neuron of a chip (src/mapped.hpp): only what Python can do with it — set_attributes
struct MappedRef
{
    std::shared_ptr<Chip> chip;
    std::string group;
    size_t offset;
};

int parse_timing(const std::string &s) // parse_timing_model  src/chip.cpp:1833-1858
{
    if (s == "simple") return SFE_TIMING_SIMPLE;
    if (s == "detailed") return SFE_TIMING_DETAILED;
    if (s == "cycle") return SFE_TIMING_CYCLE;
    throw std::invalid_argument("Error: unsupported timing model: " + s);
}

void write_sink(const py::object &sink, const std::string &text)
{
    if (py::isinstance<py::str>(sink))
    {
        std::ofstream f(sink.cast<std::string>(), std::ios::app);
        if (!f) throw std::runtime_error("Error: Couldn't open trace file for writing.");
        f << text;
    }
    else sink.attr("write")(text);
}

class Chip : public std::enable_shared_from_this<Chip>
{
public:
    sfe_chip *handle() { return h_; }
    // SpikingChip.mapped_neuron_groups  src/pymodule.cpp:1183-1191: {group name: [MappedNeuron, ...]}
    py::dict mapped_neuron_groups()
    {
        py::dict groups;
        std::string text(sfe_chip_group_names(h_, nullptr, 0) + 1, '\0');
        sfe_chip_group_names(h_, text.data(), text.size());
        std::istringstream in(text.c_str());
        for (std::string line; std::getline(in, line);)
        {
            const size_t tab = line.rfind('\t');
            const std::string name = line.substr(0, tab);
            const size_t count = std::stoul(line.substr(tab + 1));
            py::list neurons;
            for (size_t k = 0; k < count; ++k) neurons.append(MappedRef{shared_from_this(), name, k});
            groups[py::str(name)] = neurons;
        }
        return groups;
    }

    explicit Chip(const std::shared_ptr<ArchHandle> &arch, int device) : arch_(arch)
    {
        h_ = sfe_chip_create(arch->h, device);
        if (h_ == nullptr) raise_last();
    }
    ~Chip() { sfe_chip_destroy(h_); }
    Chip(const Chip &) = delete;
    Chip &operator=(const Chip &) = delete;

    void load(const std::shared_ptr<NetHandle> &net, bool overwrite)
    {
        // SpikingChip::load  src/chip.cpp:129-138: overwrite=false maps the network NEXT TO what is already on the
        // chip. This engine lowers one description per load: a second load has to replace the first.
        if (loaded_ && !overwrite)
            throw std::runtime_error("SpikingChip.load: mapping a second network next to the loaded one is not supported by "
                                     "the B200 engine; pass overwrite=True to replace it");
        if (sfe_chip_load(h_, net->h) != 0) raise_last();
        loaded_ = true;
    }
    void reset()
    {
        if (sfe_chip_reset(h_) != 0) raise_last();
    }
    double get_power() { return sfe_chip_get_power(h_); }

    py::dict sim(long timesteps, const std::string &timing_model, int /*processing_threads: the device has its own*/,
            int scheduler_threads, const py::object &spike_trace,
            const py::object &potential_trace, const py::object &neuron_trace, const py::object &perf_trace,
            const py::object &message_trace, bool write_trace_headers)
    {
        const sfe_tables *t = sfe_chip_tables(h_);
        if (t == nullptr) throw std::runtime_error("no network loaded");
        // host threads of the detailed timing model (src/pymodule.cpp:549-560); 0 = one per host core
        sfe_chip_set_scheduler_threads(h_, scheduler_threads > 0 ? static_cast<uint32_t>(scheduler_threads) : 0u);
        const size_t words = (static_cast<size_t>(t->n_neurons) + 31) / 32;
        const bool want_spikes = !spike_trace.is_none(), want_pot = !potential_trace.is_none(), want_perf = !perf_trace.is_none();
        const bool want_msgs = !message_trace.is_none();
        const bool want_traces = !neuron_trace.is_none();
        std::vector<double> utraces(want_traces ? static_cast<size_t>(t->n_u_probes) * timesteps : 0);
        std::vector<uint8_t> status(want_msgs ? static_cast<size_t>(t->n_neurons) * timesteps : 0);
        std::vector<uint32_t> fired(want_spikes ? words * timesteps : 0);
        std::vector<double> pots(want_pot ? static_cast<size_t>(t->n_probes) * timesteps : 0);
        std::vector<sfe_step_record> steps(want_perf ? timesteps : 0);
        sfe_trace_request req{};
        req.fired_bits = fired.empty() ? nullptr : fired.data();
        req.potentials = pots.empty() ? nullptr : pots.data();
        req.steps = steps.empty() ? nullptr : steps.data();
        req.status = status.empty() ? nullptr : status.data();
        req.neuron_traces = utraces.empty() ? nullptr : utraces.data();
        sfe_run_data rd{};
        const int timing = parse_timing(timing_model);
        // The device loop runs on a worker thread with the GIL released; this thread wakes every 100 ms to let
        // Python deliver signals (Ctrl-C), as the reference's loop does (src/pymodule.cpp:629-652), and asks the
        // engine to stop at its next batch boundary when one arrives.
        int rc = 0, kind = SFE_ERROR_RUNTIME;
        std::string error;
        bool interrupted = false;
        {
            std::mutex mu;
            std::condition_variable cv;
            bool finished = false;
            std::thread worker([&]() {
                const int r = sfe_chip_sim(h_, timesteps, timing, &req, &rd);
                const std::lock_guard<std::mutex> lock(mu);
                rc = r;
                if (r != 0)
                {
                    error = sfe_last_error(); // thread-local: carry it over
                    kind = sfe_last_error_kind();
                }
                finished = true;
                cv.notify_all();
            });
            for (;;)
            {
                {
                    py::gil_scoped_release release;
                    std::unique_lock<std::mutex> lock(mu);
                    if (cv.wait_for(lock, std::chrono::milliseconds(100), [&] { return finished; })) break;
                }
                if (!interrupted && PyErr_CheckSignals() != 0)
                {
                    interrupted = true;
                    sfe_chip_request_stop(h_);
                }
            }
            {
                py::gil_scoped_release release;
                worker.join();
            }
        }
        if (interrupted) throw py::error_already_set();
        if (rc != 0)
        {
            switch (kind)
            {
            case SFE_ERROR_INVALID_ARGUMENT: throw std::invalid_argument(error);
            case SFE_ERROR_OUT_OF_RANGE: throw std::out_of_range(error);
            case SFE_ERROR_HARDWARE_MAPPING: throw sfe::HardwareMappingError(error);
            default: throw std::runtime_error(error);
            }
        }
        // result dictionary  src/pymodule.cpp:268-288, 698-705
        py::dict energy;
        energy["total"] = rd.total_energy;
        energy["synapse"] = rd.synapse_energy;
        energy["dendrite"] = rd.dendrite_energy;
        energy["soma"] = rd.soma_energy;
        energy["network"] = rd.network_energy;
        py::dict out;
        out["timestep_start"] = rd.timestep_start;
        out["timesteps_executed"] = rd.timesteps_executed;
        out["energy"] = energy;
        out["sim_time"] = rd.sim_time;
        out["spikes"] = rd.spikes;
        out["packets_sent"] = rd.packets_sent;
        out["neurons_updated"] = rd.neurons_updated;
        out["neurons_fired"] = rd.neurons_fired;
        // the five trace entries always exist; they stay None unless the trace was kept in memory (src/pymodule.cpp:691-702)
        for (const char *key : {"spike_trace", "potential_trace", "neuron_trace", "perf_trace", "message_trace"}) out[key] = py::none();
        if (want_spikes)
        {
            const size_t need = sfe_chip_format_spikes(h_, fired.data(), timesteps, rd.timestep_start, nullptr, 0);
            std::string text(need + 1, '\0');
            sfe_chip_format_spikes(h_, fired.data(), timesteps, rd.timestep_start, text.data(), need + 1);
            text.resize(need);
            if (py::isinstance<py::bool_>(spike_trace))
            {
                // in-memory: one list of NeuronAddress per timestep (src/pytrace.hpp:98-141)
                std::vector<py::list> per_step(timesteps);
                std::istringstream in(text);
                std::string row;
                while (std::getline(in, row))
                {
                    const size_t comma = row.rfind(','), dot = row.rfind('.', comma);
                    const long ts = std::stol(row.substr(comma + 1));
                    sfe::NeuronAddress address; // PySpikeTrace hands out NeuronAddress objects (src/pytrace.hpp:121-141)
                    address.group_name = row.substr(0, dot);
                    address.neuron_offset = std::stoul(row.substr(dot + 1, comma - dot - 1));
                    per_step[ts - rd.timestep_start].append(py::cast(address));
                }
                py::list all;
                for (auto &l : per_step) all.append(l);
                out["spike_trace"] = all;
            }
            else write_sink(spike_trace, (write_trace_headers ? std::string("neuron,timestep\n") : std::string()) + text);
        }
        if (want_pot)
        {
            if (py::isinstance<py::bool_>(potential_trace))
            {
                py::list all;
                for (long s = 0; s < timesteps; ++s)
                {
                    py::list row;
                    for (uint32_t p = 0; p < t->n_probes; ++p) row.append(pots[static_cast<size_t>(s) * t->n_probes + p]);
                    all.append(row);
                }
                out["potential_trace"] = all;
            }
            else
            {
                std::ostringstream text; // src/chip.cpp:1454-1476, 1632-1662 (default ostream precision)
                if (write_trace_headers)
                {
                    const size_t need = sfe_chip_probe_names(h_, nullptr, 0);
                    std::string names(need + 1, '\0');
                    sfe_chip_probe_names(h_, names.data(), need + 1);
                    names.resize(need);
                    text << "timestep,";
                    std::istringstream in(names);
                    std::string nm;
                    while (std::getline(in, nm)) text << "neuron " << nm << ",";
                    text << "\n";
                }
                if (t->n_probes > 0)
                    for (long s = 0; s < timesteps; ++s)
                    {
                        text << (rd.timestep_start + s) << ",";
                        for (uint32_t p = 0; p < t->n_probes; ++p) text << pots[static_cast<size_t>(s) * t->n_probes + p] << ",";
                        text << "\n";
                    }
                write_sink(potential_trace, text.str());
            }
        }
        if (want_traces)
        {
            const uint32_t n_tr = t->n_u_probes;
            if (py::isinstance<py::bool_>(neuron_trace))
            {
                // in memory: {trace name: [values of the traced neurons, per timestep]}  src/pytrace.hpp:205-223
                py::dict data;
                if (n_tr > 0)
                {
                    py::list all;
                    for (long s = 0; s < timesteps; ++s)
                    {
                        py::list row;
                        for (uint32_t p = 0; p < n_tr; ++p) row.append(utraces[static_cast<size_t>(s) * n_tr + p]);
                        all.append(row);
                    }
                    data["u"] = all;
                }
                out["neuron_trace"] = data;
            }
            else
            {
                std::ostringstream text; // src/chip.cpp:1478-1517, 1664-1702
                if (write_trace_headers)
                {
                    const size_t need = sfe_chip_trace_names(h_, nullptr, 0);
                    std::string names(need + 1, '\0');
                    sfe_chip_trace_names(h_, names.data(), need + 1);
                    names.resize(need);
                    text << "timestep,";
                    std::istringstream in(names);
                    std::string nm;
                    while (std::getline(in, nm)) text << "neuron " << nm << ",";
                    text << "\n";
                }
                for (long s = 0; s < timesteps; ++s)
                {
                    text << (rd.timestep_start + s) << ",";
                    for (uint32_t p = 0; p < n_tr; ++p) text << utraces[static_cast<size_t>(s) * n_tr + p] << ",";
                    if (n_tr > 0) text << "\n";
                }
                write_sink(neuron_trace, text.str());
            }
        }
        if (want_msgs)
        {
            // messages.csv rows  src/chip.cpp:1586-1608, 1731-1764
            const size_t need = sfe_chip_format_messages(h_, status.data(), timesteps, rd.timestep_start, timing, nullptr, 0);
            std::string text(need + 1, '\0');
            sfe_chip_format_messages(h_, status.data(), timesteps, rd.timestep_start, timing, text.data(), text.size());
            text.resize(need);
            if (py::isinstance<py::bool_>(message_trace))
            {
                // in memory: per timestep a list of message dictionaries (message_to_dict, src/pytrace.cpp:17-53) with
                // the fields the device trace carries (those of messages.csv, at its 6-digit precision); tile
                // coordinates, global core ids and axon ids are not reconstructed
                std::vector<py::list> per_step(timesteps);
                std::istringstream in(text);
                for (std::string line; std::getline(in, line);)
                {
                    std::vector<std::string> f;
                    std::istringstream cells(line);
                    for (std::string cell; std::getline(cells, cell, ',');) f.push_back(cell);
                    if (f.size() < 16) continue;
                    auto split = [](const std::string &x) {
                        const size_t dot = x.rfind('.');
                        return std::make_pair(x.substr(0, dot), x.substr(dot + 1));
                    };
                    py::dict msg;
                    const long ts = std::stol(f[0]);
                    msg["timestep"] = ts;
                    msg["mid"] = std::stol(f[1]);
                    const auto src_neuron = split(f[2]);
                    msg["src_neuron_group_id"] = src_neuron.first;
                    msg["src_neuron_offset"] = std::stoul(src_neuron.second);
                    const auto src_hw = split(f[3]);
                    msg["src_tile_id"] = std::stoul(src_hw.first);
                    msg["src_core_offset"] = std::stoul(src_hw.second);
                    const bool placeholder = f[4] == "x.x";
                    msg["placeholder"] = placeholder;
                    if (!placeholder)
                    {
                        const auto dest_hw = split(f[4]);
                        msg["dest_tile_id"] = std::stoul(dest_hw.first);
                        msg["dest_core_offset"] = std::stoul(dest_hw.second);
                    }
                    msg["hops"] = std::stol(f[5]);
                    msg["spikes"] = std::stol(f[6]);
                    msg["send_timestamp"] = std::stod(f[7]);
                    msg["received_timestamp"] = std::stod(f[8]);
                    msg["processed_timestamp"] = std::stod(f[9]);
                    msg["generation_delay"] = std::stod(f[10]);
                    msg["processing_delay"] = std::stod(f[11]);
                    msg["network_delay"] = std::stod(f[12]);
                    msg["blocking_delay"] = std::stod(f[13]);
                    if (ts >= rd.timestep_start && ts < rd.timestep_start + timesteps) per_step[ts - rd.timestep_start].append(msg);
                }
                py::list all;
                for (auto &l : per_step) all.append(l);
                out["message_trace"] = all;
            }
            else
            {
                const std::string header = "timestep,mid,src_neuron,src_hw,dest_hw,hops,spikes,send_timestamp,received_timestamp,"
                                           "processed_timestamp,generation_delay,processing_delay,network_delay,blocking_delay,"
                                           "min_hop_delay,messages_along_route\n";
                write_sink(message_trace, (write_trace_headers ? header : std::string()) + text);
            }
        }
        if (want_perf)
        {
            if (py::isinstance<py::bool_>(perf_trace))
            {
                py::dict cols;
                py::list ts, fired_l, updated, packets, hops, spikes, sim_time, syn, den, soma, net, total;
                for (long s = 0; s < timesteps; ++s)
                {
                    const sfe_step_record &r = steps[s];
                    ts.append(rd.timestep_start + s);
                    fired_l.append(r.neurons_fired);
                    updated.append(r.neurons_updated);
                    packets.append(r.packets_sent);
                    hops.append(r.total_hops);
                    spikes.append(r.spike_count);
                    sim_time.append(r.sim_time);
                    syn.append(r.synapse_energy);
                    den.append(r.dendrite_energy);
                    soma.append(r.soma_energy);
                    net.append(r.network_energy);
                    total.append(r.total_energy);
                }
                cols["timestep"] = ts;
                cols["fired"] = fired_l;
                cols["updated"] = updated;
                cols["packets"] = packets;
                cols["hops"] = hops;
                cols["spikes"] = spikes;
                cols["sim_time"] = sim_time;
                cols["synapse_energy"] = syn;
                cols["dendrite_energy"] = den;
                cols["soma_energy"] = soma;
                cols["network_energy"] = net;
                cols["total_energy"] = total;
                out["perf_trace"] = cols;
            }
            else
            {
                std::ostringstream text; // src/chip.cpp:1557-1585, 1704-1729
                if (write_trace_headers)
                    text << "timestep,fired,updated,packets,hops,spikes,sim_time,synapse_energy,dendrite_energy,"
                            "soma_energy,network_energy,total_energy\n";
                for (long s = 0; s < timesteps; ++s)
                {
                    const sfe_step_record &r = steps[s];
                    text << (rd.timestep_start + s) << "," << r.neurons_fired << "," << r.neurons_updated << "," << r.packets_sent
                         << "," << r.total_hops << "," << r.spike_count << "," << std::scientific << r.sim_time << ","
                         << r.synapse_energy << "," << r.dendrite_energy << "," << r.soma_energy << "," << r.network_energy
                         << "," << r.total_energy << std::defaultfloat << "\n";
                }
                write_sink(perf_trace, text.str());
            }
        }
        return out;
    }

private:
    std::shared_ptr<ArchHandle> arch_;
    sfe_chip *h_{nullptr};
    bool loaded_{false};
};
} // namespace

PYBIND11_MODULE(sanafecpp_b200, m)
{
    m.doc() = "SANA-FE time-step engine on B200 (drop-in for the sanafecpp surface it covers)";
    py::register_exception<sfe::HardwareMappingError>(m, "HardwareMappingError");
    py::enum_<sfe::BufferPosition>(m, "BufferPosition") // src/pymodule.cpp:862-868
            .value("buffer_before_dendrite_unit", sfe::buffer_before_dendrite_unit)
            .value("buffer_before_soma_unit", sfe::buffer_before_soma_unit)
            .value("buffer_before_axon_out_unit", sfe::buffer_before_axon_out_unit);
    {
        // attribute documentation tables (src/pipeline.hpp:182-205, src/models.cpp:969-987): names as in the
        // reference, descriptions in this module's own words
        py::dict fw;
        for (const char *k : {"force_update", "synapse_hw_name", "dendrite_hw_name", "soma_hw_name", "model", "plugin",
                     "energy_message_in", "latency_message_in", "energy_access_neuron", "latency_access_neuron",
                     "energy_update_neuron", "latency_update_neuron", "energy_spike_out", "latency_spike_out",
                     "energy_process_spike", "latency_process_spike", "energy_update", "latency_update",
                     "energy_message_out", "latency_message_out"})
            fw[k] = "framework attribute (see the architecture description format)";
        m.attr("framework_attributes") = fw;
        py::dict models;
        auto doc = [](std::initializer_list<const char *> names) {
            py::dict d;
            for (const char *n : names) d[n] = "model attribute";
            return d;
        };
        models["current_based"] = doc({"weight", "w"});
        models["accumulator"] = py::none();
        models["accumulator_with_delay"] = py::none();
        models["input"] = doc({"rate", "poisson", "spikes"});
        models["leaky_integrate_and_fire"] = doc({"bias", "force_update", "force_update_every_timestep", "leak_decay", "log_u",
                "noise", "noise_bits", "refractory_delay", "reset_mode", "reverse_reset_mode", "reset", "reverse_reset",
                "reverse_threshold", "threshold"});
        models["truenorth"] = doc({"bias", "force_update", "leak", "leak_towards_zero", "reset_mode", "reverse_reset_mode",
                "reset", "reverse_reset", "reverse_threshold", "threshold"});
        m.attr("model_attributes") = models;
    }

    // ---- architecture (src/pymodule.cpp:1114-1172) ----------------------------------------------------
    py::class_<sfe::CoreConfiguration>(m, "Core")
            .def_readonly("name", &sfe::CoreConfiguration::name)
            .def_property_readonly("id", [](const sfe::CoreConfiguration &c) { return c.address.id; })
            .def_property_readonly("parent_tile_id", [](const sfe::CoreConfiguration &c) { return c.address.parent_tile_id; })
            .def_property_readonly("offset_within_tile", [](const sfe::CoreConfiguration &c) { return c.address.offset_within_tile; })
            .def("__repr__", [](const sfe::CoreConfiguration &c) {
                return "sanafe::Core(name=" + c.name + " address=" + std::to_string(c.address.parent_tile_id) + "." +
                        std::to_string(c.address.offset_within_tile) + ")";
            });
    py::class_<sfe::TileConfiguration>(m, "Tile")
            .def_readonly("name", &sfe::TileConfiguration::name)
            .def_readonly("id", &sfe::TileConfiguration::id)
            .def_readwrite("cores", &sfe::TileConfiguration::cores)
            .def("__repr__", [](const sfe::TileConfiguration &t) {
                return "sanafe::Tile(name=" + t.name + " cores=" + std::to_string(t.cores.size()) + ")";
            });
    py::class_<ArchHandle, std::shared_ptr<ArchHandle>>(m, "Architecture")
            .def("__repr__", [](const ArchHandle &a) {
                return "sanafe::Architecture(name=" + a.h->arch->name + ", tiles=" + std::to_string(a.h->arch->tiles.size()) + ")";
            })
            // as in the reference these hand out COPIES (std::vector members bound by value): all a script does
            // with a Core is Neuron.map_to_core(core), which reads its address
            .def_property_readonly("tiles", [](const ArchHandle &a) { return a.h->arch->tiles; })
            .def("cores", [](const ArchHandle &a) {
                std::vector<sfe::CoreConfiguration> out;
                for (const sfe::CoreConfiguration *c : static_cast<const sfe::Architecture &>(*a.h->arch).cores()) out.push_back(*c);
                return out;
            })
            .def(
                    "create_tile",
                    [](ArchHandle &a, std::string name, double en, double ln, double ee, double le, double es, double ls,
                            double ew, double lw, bool log_energy) {
                        sfe::TilePowerMetrics pm;
                        pm.energy_north_hop = en; pm.latency_north_hop = ln;
                        pm.energy_east_hop = ee; pm.latency_east_hop = le;
                        pm.energy_south_hop = es; pm.latency_south_hop = ls;
                        pm.energy_west_hop = ew; pm.latency_west_hop = lw;
                        pm.log_energy = log_energy;
                        return a.h->arch->create_tile(std::move(name), pm);
                    },
                    py::arg("name"), py::arg("energy_north_hop") = 0.0, py::arg("latency_north_hop") = 0.0,
                    py::arg("energy_east_hop") = 0.0, py::arg("latency_east_hop") = 0.0, py::arg("energy_south_hop") = 0.0,
                    py::arg("latency_south_hop") = 0.0, py::arg("energy_west_hop") = 0.0, py::arg("latency_west_hop") = 0.0,
                    py::arg("log_energy") = false)
            .def(
                    "create_core",
                    [](ArchHandle &a, std::string name, size_t parent_tile_id, const std::string &buffer_position,
                            bool buffer_inside_unit, size_t max_neurons_supported, bool log_energy) {
                        sfe::CorePipelineConfiguration pc{};
                        pc.buffer_position = sfe::parse_buffer_position(buffer_position, buffer_inside_unit);
                        pc.max_neurons_supported = max_neurons_supported;
                        pc.log_energy = log_energy;
                        return a.h->arch->create_core(std::move(name), parent_tile_id, pc);
                    },
                    py::arg("name"), py::arg("parent_tile_id"), py::arg("buffer_position") = "soma",
                    py::arg("buffer_inside_unit") = false, py::arg("max_neurons_supported") = 1024, py::arg("log_energy") = false);

    // ---- network (src/pymodule.cpp:873-1112) --------------------------------------------------------------
    py::class_<sfe::NeuronAddress>(m, "NeuronAddress")
            .def_readonly("group_name", &sfe::NeuronAddress::group_name)
            .def_readonly("neuron_offset", &sfe::NeuronAddress::neuron_offset)
            .def("__repr__", [](const sfe::NeuronAddress &a) {
                return a.group_name + "." + (a.neuron_offset ? std::to_string(*a.neuron_offset) : std::string("*"));
            })
            .def(py::pickle([](const sfe::NeuronAddress &a) { return py::make_tuple(a.group_name, a.neuron_offset); },
                    [](const py::tuple &t) {
                        sfe::NeuronAddress a;
                        a.group_name = t[0].cast<std::string>();
                        a.neuron_offset = t[1].cast<std::optional<size_t>>();
                        return a;
                    }));
    py::class_<sfe::Connection>(m, "Connection")
            .def_readonly("pre_neuron", &sfe::Connection::pre_neuron)
            .def_readonly("post_neuron", &sfe::Connection::post_neuron)
            .def_readonly("synapse_hw_name", &sfe::Connection::synapse_hw_name)
            .def_property_readonly("synapse_attributes", [](const sfe::Connection &c) { return from_attr_map(c.synapse_attributes); })
            .def("__repr__", [](const sfe::Connection &c) {
                return "sanafe::Connection(pre_neuron=" + c.pre_neuron.group_name + "." + std::to_string(c.pre_neuron.neuron_offset.value_or(0)) +
                        " post_neuron=" + c.post_neuron.group_name + "." + std::to_string(c.post_neuron.neuron_offset.value_or(0)) + ")";
            });
    py::class_<NeuronRef>(m, "Neuron")
            .def("__repr__", [](const NeuronRef &r) { return neuron_info(*r.n); })
            .def("get_id", [](const NeuronRef &r) { return r.n->offset; })
            .def("map_to_core", [](const NeuronRef &r, const sfe::CoreConfiguration &core) { r.n->map_to_core(core); })
            .def(
                    "set_attributes",
                    // The reference binds the keyword names `soma_attributes` / `dendrite_attributes` to the C++
                    // parameters in swapped order (src/pymodule.cpp:1016-1042 vs :466-490):
                    // what a script passes as soma_attributes= is forwarded to the DENDRITE unit and vice versa.
                    // A drop-in keeps that.
                    [](const NeuronRef &r, std::optional<std::string> soma_hw_name, std::optional<std::string> default_synapse_hw_name,
                            std::optional<std::string> dendrite_hw_name, std::optional<bool> log_spikes, std::optional<bool> log_potential,
                            const py::dict &model_attributes, const py::dict &bound_as_soma_attributes,
                            const py::dict &bound_as_dendrite_attributes) {
                        sfe::NeuronConfiguration c;
                        c.soma_hw_name = std::move(soma_hw_name);
                        c.default_synapse_hw_name = std::move(default_synapse_hw_name);
                        c.dendrite_hw_name = std::move(dendrite_hw_name);
                        c.log_spikes = log_spikes;
                        c.log_potential = log_potential;
                        c.model_attributes = to_attr_map(model_attributes);
                        for (auto &kv : to_attr_map(bound_as_soma_attributes, false, true, false)) c.model_attributes.insert(kv);
                        for (auto &kv : to_attr_map(bound_as_dendrite_attributes, false, false, true)) c.model_attributes.insert(kv);
                        r.n->set_attributes(c);
                    },
                    py::arg("soma_hw_name") = py::none(), py::arg("default_synapse_hw_name") = py::none(),
                    py::arg("dendrite_hw_name") = py::none(), py::arg("log_spikes") = py::none(), py::arg("log_potential") = py::none(),
                    py::arg("model_attributes") = py::dict(), py::arg("soma_attributes") = py::dict(),
                    py::arg("dendrite_attributes") = py::dict())
            .def(
                    "connect_to_neuron",
                    [](const NeuronRef &r, const NeuronRef &dest, std::optional<py::dict> attr) -> size_t {
                        const size_t idx = r.n->connect_to_neuron(*dest.n);
                        sfe::Connection &con = r.n->edges_out[idx];
                        for (const auto &[key, a] : to_attr_map(attr.value_or(py::dict())))
                        {
                            if (a.forward_to_synapse) con.synapse_attributes[key] = a;
                            if (a.forward_to_dendrite) con.dendrite_attributes[key] = a;
                        }
                        return idx;
                    },
                    py::arg("dest"), py::arg("attributes") = py::none())
            .def_property_readonly("edges_out", [](const NeuronRef &r) { return r.n->edges_out; });
    py::class_<GroupRef>(m, "NeuronGroup")
            .def("__repr__", [](const GroupRef &g) { return group_info(*g.g); })
            .def("get_name", [](const GroupRef &g) { return g.g->name; })
            .def("connect_neurons_dense",
                    [](const GroupRef &g, const GroupRef &dest, const py::dict &attributes) {
                        g.g->connect_neurons_dense(*dest.g, to_attr_lists(attributes));
                    },
                    py::arg("dest_group"), py::arg("attributes"))
            .def("connect_neurons_sparse",
                    [](const GroupRef &g, const GroupRef &dest, const py::dict &attributes, const py::object &pairs) {
                        if (!py::isinstance<py::iterable>(pairs))
                            throw std::invalid_argument("Error: must provide connectivity as a list of source/destination pairs, "
                                                        "providing the offsets within the groups.");
                        std::vector<std::pair<size_t, size_t>> vec;
                        for (auto it = py::iter(pairs); it != py::iterator::sentinel(); ++it)
                        {
                            if (!py::isinstance<py::iterable>(*it)) throw py::value_error("Error: each entry in the src/dest id list must be a 2-tuple");
                            const auto seq = py::reinterpret_borrow<py::object>(*it).cast<py::sequence>();
                            if (seq.size() != 2) throw py::value_error("Expected a 2-tuple");
                            vec.emplace_back(seq[0].cast<size_t>(), seq[1].cast<size_t>());
                        }
                        g.g->connect_neurons_sparse(*dest.g, to_attr_lists(attributes), vec);
                    },
                    py::arg("dest_group"), py::arg("attributes"), py::arg("src_dest_id_pairs"))
            .def("connect_neurons_conv2d",
                    [](const GroupRef &g, const GroupRef &dest, const py::dict &attributes, int input_width, int input_height,
                            int input_channels, int kernel_width, int kernel_height, int kernel_count, int stride_width, int stride_height) {
                        sfe::Conv2DParameters cv;
                        cv.input_width = input_width; cv.input_height = input_height; cv.input_channels = input_channels;
                        cv.kernel_width = kernel_width; cv.kernel_height = kernel_height; cv.kernel_count = kernel_count;
                        cv.stride_width = stride_width; cv.stride_height = stride_height;
                        g.g->connect_neurons_conv2d(*dest.g, to_attr_lists(attributes), cv);
                    },
                    py::arg("dest_group"), py::arg("attributes"), py::arg("input_width"), py::arg("input_height"),
                    py::arg("input_channels"), py::arg("kernel_width"), py::arg("kernel_height"), py::arg("kernel_count") = 1,
                    py::arg("stride_width") = 1, py::arg("stride_height") = 1)
            .def_property_readonly("neurons",
                    [](const GroupRef &g) {
                        py::list l;
                        for (sfe::Neuron &n : g.g->neurons) l.append(NeuronRef{g.owner, &n});
                        return l;
                    })
            .def("__getitem__",
                    [](const GroupRef &g, const py::object &index) -> py::object {
                        if (py::isinstance<py::int_>(index))
                        {
                            const size_t i = index.cast<size_t>();
                            if (i >= g.g->neurons.size()) throw py::index_error();
                            return py::cast(NeuronRef{g.owner, &g.g->neurons[i]});
                        }
                        if (py::isinstance<py::slice>(index))
                        {
                            size_t start = 0, stop = 0, step = 0, len = 0;
                            if (!index.cast<py::slice>().compute(g.g->neurons.size(), &start, &stop, &step, &len)) throw py::error_already_set();
                            py::list l;
                            for (size_t k = 0; k < len; ++k) l.append(NeuronRef{g.owner, &g.g->neurons[start + k * step]});
                            return l;
                        }
                        throw py::type_error("Index must be int or slice");
                    })
            .def("__len__", [](const GroupRef &g) { return g.g->neurons.size(); })
            .def("__iter__", [](const GroupRef &g) {
                py::list l;
                for (sfe::Neuron &n : g.g->neurons) l.append(NeuronRef{g.owner, &n});
                return py::iter(l);
            });
    py::class_<NetHandle, std::shared_ptr<NetHandle>>(m, "Network")
            .def(py::init([]() {
                auto n = std::make_shared<NetHandle>();
                n->h = sfe_net_create("");
                if (n->h == nullptr) raise_last();
                return n;
            }))
            .def("__repr__", [](const std::shared_ptr<NetHandle> &n) {
                return "sanafe::SpikingNetwork(groups=" + std::to_string(described(n).groups.size()) + ")";
            })
            .def(
                    "create_neuron_group",
                    [](const std::shared_ptr<NetHandle> &n, const std::string &group_name, int neuron_count, const py::dict &model_attributes,
                            const std::string &default_synapse_hw_name, const std::string &default_dendrite_hw_name, bool log_potential,
                            bool log_spikes, const std::string &soma_hw_name) {
                        sfe::NeuronConfiguration c;
                        c.default_synapse_hw_name = default_synapse_hw_name;
                        c.dendrite_hw_name = default_dendrite_hw_name;
                        c.log_potential = log_potential;
                        c.log_spikes = log_spikes;
                        c.soma_hw_name = soma_hw_name;
                        c.model_attributes = to_attr_map(model_attributes);
                        return GroupRef{n, &described(n).create_neuron_group(group_name, static_cast<size_t>(neuron_count), c)};
                    },
                    py::arg("group_name"), py::arg("neuron_count"), py::arg("model_attributes") = py::dict(),
                    py::arg("default_synapse_hw_name") = "", py::arg("default_dendrite_hw_name") = "", py::arg("log_potential") = false,
                    py::arg("log_spikes") = false, py::arg("soma_hw_name") = "")
            .def(
                    "save",
                    [](const std::shared_ptr<NetHandle> &n, const std::string &path, bool use_netlist_format) {
                        if (use_netlist_format)
                        {
                            if (sfe_net_save_netlist(n->h, path.c_str()) != 0) raise_last();
                            return;
                        }
                        if (sfe_net_save_yaml(n->h, path.c_str()) != 0) raise_last();
                    },
                    py::arg("path"), py::arg("use_netlist_format") = false)
            .def_property_readonly("groups",
                    [](const std::shared_ptr<NetHandle> &n) {
                        py::dict d;
                        for (auto &[name, g] : described(n).groups) d[py::str(name)] = GroupRef{n, g.get()};
                        return d;
                    })
            .def("__getitem__", [](const std::shared_ptr<NetHandle> &n, const std::string &name) {
                auto it = described(n).groups.find(name);
                if (it == described(n).groups.end()) throw py::index_error();
                return GroupRef{n, it->second.get()};
            });
    py::class_<MappedRef>(m, "MappedNeuron")
            .def("__repr__", [](const MappedRef &r) { return "sanafe::MappedNeuron(" + r.group + "." + std::to_string(r.offset) + ")"; })
            .def(
                    "set_attributes",
                    // MappedNeuron::set_attributes  src/mapped.cpp:113-166 via pyset_attributes_mapped (src/pymodule.cpp:502-527).
                    // The running engine takes numeric soma parameters (per-neuron bias on a fast path, the others by
                    // re-classing the neuron); anything structural has to be set before load().
                    [](const MappedRef &r, const py::dict &model_attributes, const py::dict &bound_as_soma_attributes,
                            const py::dict &bound_as_dendrite_attributes, std::optional<bool> log_spikes) {
                        sfe::AttrMap all = to_attr_map(model_attributes);
                        for (auto &kv : to_attr_map(bound_as_soma_attributes, false, true, false)) all.insert(kv);
                        for (auto &kv : to_attr_map(bound_as_dendrite_attributes, false, false, true)) all.insert(kv);
                        sfe_chip *chip = r.chip->handle();
                        if (log_spikes.has_value() && sfe_chip_set_neuron_log_spikes(chip, r.group.c_str(), r.offset, *log_spikes ? 1 : 0) != 0) raise_last();
                        for (const auto &[key, a] : all)
                        {
                            if (sfe::is_reserved_neuron_attribute(key))
                                throw std::invalid_argument("Reserved neuron attribute '" + key + "' cannot be used as a model attribute. "
                                        "Pass it as a direct argument instead (if supported).");
                            if (!a.forward_to_soma) continue; // the built-in dendrite models take no per-neuron attributes
                            if (a.is_list() || a.is_string())
                                throw std::runtime_error("MappedNeuron.set_attributes: '" + key + "' is not a numeric soma parameter; set it on the "
                                        "Neuron before SpikingChip.load()");
                            if (sfe_chip_set_neuron_attribute(chip, r.group.c_str(), r.offset, key.c_str(), a.as_double()) != 0) raise_last();
                        }
                    },
                    py::arg("model_attributes") = py::dict(), py::arg("soma_attributes") = py::dict(),
                    py::arg("dendrite_attributes") = py::dict(), py::arg("log_spikes") = py::none());
    m.def(
            "load_arch",
            [](const std::string &path) {
                auto a = std::make_shared<ArchHandle>();
                a->h = sfe_arch_load_yaml(path.c_str());
                if (a->h == nullptr) raise_last();
                return a;
            },
            py::arg("path"));
    m.def(
            "load_net",
            [](const std::string &path, const std::shared_ptr<ArchHandle> &arch, bool use_netlist_format) {
                auto n = std::make_shared<NetHandle>();
                n->h = use_netlist_format ? sfe_net_load_netlist(path.c_str(), arch->h) : sfe_net_load_yaml(path.c_str(), arch->h);
                if (n->h == nullptr) raise_last();
                return n;
            },
            py::arg("path"), py::arg("arch"), py::arg("use_netlist_format") = false);
    m.def(
            "load_flat",
            [](const std::string &path) {
                auto a = std::make_shared<ArchHandle>();
                auto n = std::make_shared<NetHandle>();
                if (sfe_load_flat(path.c_str(), &a->h, &n->h) != 0) raise_last();
                return py::make_tuple(a, n);
            },
            py::arg("path"));
    // Out-of-tree hardware-unit models (include/sfe_device_model.h): the programmatic side of the reference's plugin
    // loader (plugin_get_hw, src/plugins.cpp:85-98). A unit's `plugin:` path is loaded by SpikingChip.load() itself.
    m.def(
            "load_device_model",
            [](const std::string &model, const std::string &library_path) {
                if (sfe_load_device_model(model.c_str(), library_path.c_str()) != 0) raise_last();
            },
            py::arg("model"), py::arg("library_path"));
    m.def(
            "unregister_device_model",
            [](const std::string &model) {
                if (sfe_unregister_device_model(model.c_str()) != 0) raise_last();
            },
            py::arg("model"));
    m.def(
            "device_model_registered", [](const std::string &model) { return sfe_device_model_registered(model.c_str()) != 0; },
            py::arg("model"));
    py::class_<Chip, std::shared_ptr<Chip>>(m, "SpikingChip")
            .def(py::init<const std::shared_ptr<ArchHandle> &, int>(), py::arg("arch"), py::arg("device") = 0)
            .def_property_readonly("mapped_neuron_groups", &Chip::mapped_neuron_groups)
            // the C-ABI handle (an address), for callers that mix this module with the C ABI
            .def_property_readonly("_handle", [](Chip &c) { return reinterpret_cast<uintptr_t>(c.handle()); })
            .def("load", &Chip::load, py::arg("net"), py::arg("overwrite") = false)
            .def("sim", &Chip::sim, py::arg("timesteps") = 1, py::arg("timing_model") = "detailed",
                    py::arg("processing_threads") = 0, py::arg("scheduler_threads") = 0, py::arg("spike_trace") = py::none(),
                    py::arg("potential_trace") = py::none(), py::arg("neuron_trace") = py::none(),
                    py::arg("perf_trace") = py::none(), py::arg("message_trace") = py::none(),
                    py::arg("write_trace_headers") = true)
            .def("reset", &Chip::reset)
            .def("get_power", &Chip::get_power);
}
