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

#include <fstream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "sanafe_b200.h"

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
    throw std::runtime_error(sfe_last_error());
}

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

class Chip
{
public:
    explicit Chip(const std::shared_ptr<ArchHandle> &arch, int device) : arch_(arch)
    {
        h_ = sfe_chip_create(arch->h, device);
        if (h_ == nullptr) raise_last();
    }
    ~Chip() { sfe_chip_destroy(h_); }
    Chip(const Chip &) = delete;
    Chip &operator=(const Chip &) = delete;

    void load(const std::shared_ptr<NetHandle> &net, bool /*overwrite*/)
    {
        if (sfe_chip_load(h_, net->h) != 0) raise_last();
    }
    void reset()
    {
        if (sfe_chip_reset(h_) != 0) raise_last();
    }
    double get_power() { return sfe_chip_get_power(h_); }

    py::dict sim(long timesteps, const std::string &timing_model, int, int, const py::object &spike_trace,
            const py::object &potential_trace, const py::object &neuron_trace, const py::object &perf_trace,
            const py::object &message_trace, bool write_trace_headers)
    {
        const sfe_tables *t = sfe_chip_tables(h_);
        if (t == nullptr) throw std::runtime_error("no network loaded");
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
        int rc = 0;
        {
            py::gil_scoped_release release;
            rc = sfe_chip_sim(h_, timesteps, timing, &req, &rd);
        }
        if (rc != 0) raise_last();
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
        if (want_spikes)
        {
            const size_t need = sfe_chip_format_spikes(h_, fired.data(), timesteps, rd.timestep_start, nullptr, 0);
            std::string text(need + 1, '\0');
            sfe_chip_format_spikes(h_, fired.data(), timesteps, rd.timestep_start, text.data(), need + 1);
            text.resize(need);
            if (py::isinstance<py::bool_>(spike_trace))
            {
                // in-memory: one list of (group, offset) per timestep (src/pytrace.hpp)
                std::vector<py::list> per_step(timesteps);
                std::istringstream in(text);
                std::string row;
                while (std::getline(in, row))
                {
                    const size_t comma = row.rfind(','), dot = row.rfind('.', comma);
                    const long ts = std::stol(row.substr(comma + 1));
                    per_step[ts - rd.timestep_start].append(py::make_tuple(row.substr(0, dot), std::stoul(row.substr(dot + 1, comma - dot - 1))));
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
                py::list rows;
                std::istringstream in(text);
                for (std::string line; std::getline(in, line);) rows.append(py::str(line).attr("split")(","));
                out["message_trace"] = rows;
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
};
} // namespace

PYBIND11_MODULE(sanafecpp_b200, m)
{
    m.doc() = "SANA-FE time-step engine on B200 (drop-in for the sanafecpp surface it covers)";
    py::class_<ArchHandle, std::shared_ptr<ArchHandle>>(m, "Architecture");
    py::class_<NetHandle, std::shared_ptr<NetHandle>>(m, "Network");
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
    py::class_<Chip>(m, "SpikingChip")
            .def(py::init<const std::shared_ptr<ArchHandle> &, int>(), py::arg("arch"), py::arg("device") = 0)
            .def("load", &Chip::load, py::arg("net"), py::arg("overwrite") = false)
            .def("sim", &Chip::sim, py::arg("timesteps") = 1, py::arg("timing_model") = "detailed",
                    py::arg("processing_threads") = 0, py::arg("scheduler_threads") = 0, py::arg("spike_trace") = py::none(),
                    py::arg("potential_trace") = py::none(), py::arg("neuron_trace") = py::none(),
                    py::arg("perf_trace") = py::none(), py::arg("message_trace") = py::none(),
                    py::arg("write_trace_headers") = true)
            .def("reset", &Chip::reset)
            .def("get_power", &Chip::get_power);
}
