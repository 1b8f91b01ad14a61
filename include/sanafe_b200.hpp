// sanafe_b200.hpp — header-only C++17 mirror of the reference's C++ user surface over the C ABI (sanafe_b200.h).
//
// What it replaces (same names, argument meaning and error behaviour, so code written against the reference's
// headers ports by changing the namespace):
//   sanafe::load_arch / sanafe::load_net            src/arch.cpp:106-117, src/network.cpp:194-222
//   sanafe::SpikingChip                              src/chip.hpp:56-107
//       SpikingChip(const Architecture &)            src/chip.cpp:61-92
//       load(const SpikingNetwork &)                 src/chip.cpp:129-138
//       sim(timesteps, timing_model, scheduler_thread_count, trace_flags, output_dir) -> RunData
//                                                    src/chip.cpp:477-537 (trace files opened by the first sim() of a
//                                                    chip and appended by later ones; writers :849-1764)
//       reset(), get_power(), get_total_timesteps()  src/chip.cpp:576-621
//       sim_output_run_summary(output_dir, RunData)  src/chip.cpp:849-899
//   sanafe::TraceFlags, sanafe::RunData, TimingModel src/chip.hpp:46-54,215-233, src/arch.hpp
// Errors come back from the C ABI as status codes; they are rethrown here as the exception class the reference throws
// (sfe_last_error_kind). The `sim` command line (csrc/host/sim_main.cpp) is written on top of this header.
#ifndef SANAFE_B200_HPP_
#define SANAFE_B200_HPP_

#include <algorithm>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "sanafe_b200.h"

namespace sanafe_b200
{
enum TimingModel : int // src/arch.hpp (timing_model_simple / detailed / cycle_accurate)
{
    timing_model_simple = SFE_TIMING_SIMPLE,
    timing_model_detailed = SFE_TIMING_DETAILED,
    timing_model_cycle_accurate = SFE_TIMING_CYCLE
};

struct TraceFlags // src/chip.hpp:46-54
{
    bool record_spikes{false};
    bool record_potentials{false};
    bool record_neuron_state{false};
    bool record_perf{false};
    bool record_messages{false};
};

using RunData = sfe_run_data; // src/chip.hpp:215-233

class HardwareMappingError : public std::runtime_error // src/mapped.hpp:30-38
{
public:
    explicit HardwareMappingError(const std::string &m) : std::runtime_error(m) {}
};

[[noreturn]] inline void throw_last_error()
{
    const std::string what = sfe_last_error();
    switch (sfe_last_error_kind())
    {
    case SFE_ERROR_INVALID_ARGUMENT: throw std::invalid_argument(what);
    case SFE_ERROR_OUT_OF_RANGE: throw std::out_of_range(what);
    case SFE_ERROR_HARDWARE_MAPPING: throw HardwareMappingError(what);
    default: throw std::runtime_error(what);
    }
}

class Architecture
{
public:
    explicit Architecture(sfe_arch *h) : h_(h) {}
    Architecture(Architecture &&o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    Architecture(const Architecture &) = delete;
    Architecture &operator=(const Architecture &) = delete;
    ~Architecture() { if (h_ != nullptr) sfe_arch_free(h_); }
    sfe_arch *handle() const { return h_; }
private:
    sfe_arch *h_;
};

class SpikingNetwork
{
public:
    explicit SpikingNetwork(sfe_net *h) : h_(h) {}
    SpikingNetwork(SpikingNetwork &&o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    SpikingNetwork(const SpikingNetwork &) = delete;
    SpikingNetwork &operator=(const SpikingNetwork &) = delete;
    ~SpikingNetwork() { if (h_ != nullptr) sfe_net_free(h_); }
    sfe_net *handle() const { return h_; }
    void save(const std::filesystem::path &path, const bool use_netlist_format = false) const // src/network.cpp:705-718
    {
        if ((use_netlist_format ? sfe_net_save_netlist(h_, path.c_str()) : sfe_net_save_yaml(h_, path.c_str())) != 0) throw_last_error();
    }
private:
    sfe_net *h_;
};

inline Architecture load_arch(const std::filesystem::path &path)
{
    sfe_arch *a = sfe_arch_load_yaml(path.c_str());
    if (a == nullptr) throw_last_error();
    return Architecture(a);
}

inline SpikingNetwork load_net(const std::filesystem::path &path, Architecture &arch, const bool use_netlist_format = false)
{
    sfe_net *n = use_netlist_format ? sfe_net_load_netlist(path.c_str(), arch.handle()) : sfe_net_load_yaml(path.c_str(), arch.handle());
    if (n == nullptr) throw_last_error();
    return SpikingNetwork(n);
}

class SpikingChip
{
public:
    // device: CUDA device ordinal (the reference has no such argument: one process, host threads)
    explicit SpikingChip(const Architecture &arch, const int device = 0) : h_(sfe_chip_create(arch.handle(), device))
    {
        if (h_ == nullptr) throw_last_error();
    }
    SpikingChip(const SpikingChip &) = delete;
    SpikingChip(SpikingChip &&) = delete;
    SpikingChip &operator=(const SpikingChip &) = delete;
    SpikingChip &operator=(SpikingChip &&) = delete;
    ~SpikingChip() { if (h_ != nullptr) sfe_chip_destroy(h_); }

    void load(const SpikingNetwork &net)
    {
        if (sfe_chip_load(h_, net.handle()) != 0) throw_last_error();
    }

    // scheduler_thread_count: host threads of the detailed timing model (0 = one per host core)
    RunData sim(const long timesteps = 1, const TimingModel timing_model = timing_model_detailed, const int scheduler_thread_count = 1,
            const TraceFlags trace_flags = TraceFlags(), const std::string &output_dir = "")
    {
        if (timesteps <= 0) throw std::invalid_argument("Time-steps must be > 0");
        sfe_chip_set_scheduler_threads(h_, scheduler_thread_count > 0 ? static_cast<uint32_t>(scheduler_thread_count) : 0u);
        if (total_timesteps_ <= 0) open_traces(trace_flags, output_dir); // src/chip.cpp:493-516
        const sfe_tables *t = sfe_chip_tables(h_);
        const size_t n = t->n_neurons, words = (n + 31) / 32, probes = t->n_probes;
        const bool any_step_trace = spike_trace_.is_open() || potential_trace_.is_open() || perf_trace_.is_open() || message_trace_.is_open();
        // a chunk of timesteps per call keeps the host buffers bounded (the status bytes are one per neuron and step)
        const long chunk = std::max<long>(1, std::min<long>(1024, static_cast<long>((64u << 20) / std::max<size_t>(n, 1))));
        RunData total{};
        bool first = true;
        for (long done = 0; done < timesteps;)
        {
            const long batch = std::min(chunk, timesteps - done);
            sfe_trace_request req{};
            if (spike_trace_.is_open())
            {
                fired_.assign(static_cast<size_t>(batch) * words, 0u);
                req.fired_bits = fired_.data();
            }
            if (potential_trace_.is_open() && probes > 0)
            {
                pots_.assign(static_cast<size_t>(batch) * probes, 0.0);
                req.potentials = pots_.data();
            }
            if (message_trace_.is_open())
            {
                status_.assign(static_cast<size_t>(batch) * n, 0);
                req.status = status_.data();
            }
            if (perf_trace_.is_open())
            {
                steps_.assign(static_cast<size_t>(batch), sfe_step_record{});
                req.steps = steps_.data();
            }
            if (neuron_trace_.is_open() && n_neuron_traces_ > 0)
            {
                utraces_.assign(static_cast<size_t>(batch) * n_neuron_traces_, 0.0);
                req.neuron_traces = utraces_.data();
            }
            RunData rd{};
            const bool want = any_step_trace || (neuron_trace_.is_open() && n_neuron_traces_ > 0);
            if (sfe_chip_sim(h_, batch, timing_model, want ? &req : nullptr, &rd) != 0) throw_last_error();
            write_traces(rd, batch, timing_model, probes);
            if (first) total = rd;
            else accumulate(total, rd);
            first = false;
            done += batch;
        }
        total_timesteps_ += timesteps;
        return total;
    }

    void reset() // src/chip.cpp:576-600
    {
        if (sfe_chip_reset(h_) != 0) throw_last_error();
    }
    double get_power() const { return sfe_chip_get_power(h_); } // src/chip.cpp:607-621
    long get_total_timesteps() const noexcept { return total_timesteps_; }
    const sfe_tables *tables() const { return sfe_chip_tables(h_); }
    sfe_chip *handle() const { return h_; }

    // MappedNeuron::set_attributes for one numeric attribute after load (src/mapped.cpp:113-166)
    void set_neuron_attribute(const std::string &group, const size_t offset, const std::string &name, const double value)
    {
        if (sfe_chip_set_neuron_attribute(h_, group.c_str(), offset, name.c_str(), value) != 0) throw_last_error();
    }

    // sim_output_run_summary / sim_format_run_summary  src/chip.cpp:849-899 (the wall-time split is this engine's:
    // device + host scheduler)
    static void format_run_summary(std::ostream &out, const RunData &rd)
    {
        out << "build_git_version: 'sanafe-b200'\n";
        out << "timesteps_executed: " << rd.timesteps_executed << "\n";
        out << "total_spikes: " << rd.spikes << "\n";
        out << "total_messages_sent: " << rd.packets_sent << "\n";
        out << "total_neurons_updated: " << rd.neurons_updated << "\n";
        out << "total_neurons_fired: " << rd.neurons_fired << "\n";
        out << "sim_time: " << std::scientific << rd.sim_time << "\n";
        out << "energy:\n";
        out << "  synapse:" << std::scientific << rd.synapse_energy << "\n";
        out << "  dendrite:" << std::scientific << rd.dendrite_energy << "\n";
        out << "  soma:" << std::scientific << rd.soma_energy << "\n";
        out << "  network: " << std::scientific << rd.network_energy << "\n";
        out << "  total: " << std::scientific << rd.total_energy << "\n";
        out << "wall_time:\n";
        out << "  device: " << std::fixed << (rd.wall_time - rd.scheduler_wall_time) << "\n";
        out << "  scheduler: " << std::fixed << rd.scheduler_wall_time << "\n";
    }
    void sim_output_run_summary(const std::filesystem::path &output_dir, const RunData &rd) const
    {
        std::ofstream f(output_dir / "run_summary.yaml");
        if (f.is_open()) format_run_summary(f, rd);
    }

    static void accumulate(RunData &total, const RunData &rd)
    {
        total.timesteps_executed += rd.timesteps_executed;
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
        total.wall_time += rd.wall_time;
        total.scheduler_wall_time += rd.scheduler_wall_time;
    }

private:
    static std::ofstream open_trace(const std::filesystem::path &path)
    {
        std::ofstream f(path);
        if (!f.is_open()) throw std::runtime_error("Error: Couldn't open trace file for writing."); // src/chip.cpp:901-975
        return f;
    }
    std::string names(size_t (*fn)(const sfe_chip *, char *, size_t)) const
    {
        std::string text(fn(h_, nullptr, 0) + 1, '\0');
        fn(h_, text.data(), text.size());
        text.resize(std::strlen(text.c_str()));
        return text;
    }
    void open_traces(const TraceFlags &flags, const std::filesystem::path &dir)
    {
        if (flags.record_spikes)
        {
            spike_trace_ = open_trace(dir / "spikes.csv");
            spike_trace_ << "neuron,timestep\n"; // src/chip.cpp:1447-1452
        }
        if (flags.record_potentials)
        {
            potential_trace_ = open_trace(dir / "potentials.csv");
            potential_trace_ << "timestep,"; // src/chip.cpp:1454-1476
            std::istringstream lines(names(sfe_chip_probe_names));
            for (std::string line; std::getline(lines, line);) potential_trace_ << "neuron " << line << ",";
            potential_trace_ << "\n";
        }
        if (flags.record_neuron_state)
        {
            neuron_trace_ = open_trace(dir / "neurons.csv");
            neuron_trace_ << "timestep,"; // src/chip.cpp:1478-1517
            std::istringstream lines(names(sfe_chip_trace_names));
            for (std::string line; std::getline(lines, line); ++n_neuron_traces_) neuron_trace_ << "neuron " << line << ",";
            neuron_trace_ << "\n";
        }
        if (flags.record_perf)
        {
            perf_trace_ = open_trace(dir / "perf.csv");
            perf_trace_ << "timestep,fired,updated,packets,hops,spikes,sim_time,synapse_energy,dendrite_energy,soma_energy,"
                           "network_energy,total_energy\n"; // src/chip.cpp:1519-1538
        }
        if (flags.record_messages)
        {
            message_trace_ = open_trace(dir / "messages.csv");
            message_trace_ << "timestep,mid,src_neuron,src_hw,dest_hw,hops,spikes,send_timestamp,received_timestamp,"
                              "processed_timestamp,generation_delay,processing_delay,network_delay,blocking_delay,min_hop_delay,"
                              "messages_along_route\n"; // src/chip.cpp:1540-1560
        }
    }
    void write_traces(const RunData &rd, const long batch, const int timing_model, const size_t probes)
    {
        if (neuron_trace_.is_open())
            for (long s = 0; s < batch; ++s)
            {
                // sim_trace_record_neuron_traces  src/chip.cpp:1664-1702 (default ostream precision)
                neuron_trace_ << (rd.timestep_start + s) << ",";
                for (size_t p = 0; p < n_neuron_traces_; ++p) neuron_trace_ << utraces_[static_cast<size_t>(s) * n_neuron_traces_ + p] << ',';
                if (n_neuron_traces_ > 0) neuron_trace_ << "\n";
            }
        if (spike_trace_.is_open())
        {
            text_.resize(sfe_chip_format_spikes(h_, fired_.data(), batch, rd.timestep_start, nullptr, 0) + 1);
            sfe_chip_format_spikes(h_, fired_.data(), batch, rd.timestep_start, text_.data(), text_.size());
            spike_trace_ << text_.c_str();
        }
        for (long s = 0; s < batch; ++s)
        {
            if (potential_trace_.is_open())
            {
                potential_trace_ << (rd.timestep_start + s) << ",";
                for (size_t p = 0; p < probes; ++p) potential_trace_ << pots_[static_cast<size_t>(s) * probes + p] << ',';
                if (probes > 0) potential_trace_ << "\n";
            }
            if (perf_trace_.is_open())
            {
                const sfe_step_record &r = steps_[static_cast<size_t>(s)];
                perf_trace_ << (rd.timestep_start + s) << "," << r.neurons_fired << "," << r.neurons_updated << "," << r.packets_sent << ","
                            << r.total_hops << "," << r.spike_count << "," << std::scientific << r.sim_time << "," << r.synapse_energy
                            << "," << r.dendrite_energy << "," << r.soma_energy << "," << r.network_energy << "," << r.total_energy << "\n";
            }
        }
        if (message_trace_.is_open())
        {
            text_.resize(sfe_chip_format_messages(h_, status_.data(), batch, rd.timestep_start, timing_model, nullptr, 0) + 1);
            sfe_chip_format_messages(h_, status_.data(), batch, rd.timestep_start, timing_model, text_.data(), text_.size());
            message_trace_ << text_.c_str();
        }
    }

    sfe_chip *h_{nullptr};
    long total_timesteps_{0};
    std::ofstream spike_trace_, potential_trace_, neuron_trace_, perf_trace_, message_trace_;
    size_t n_neuron_traces_{0};
    std::vector<uint32_t> fired_;
    std::vector<double> pots_, utraces_;
    std::vector<uint8_t> status_;
    std::vector<sfe_step_record> steps_;
    std::string text_;
};

} // namespace sanafe_b200
#endif
