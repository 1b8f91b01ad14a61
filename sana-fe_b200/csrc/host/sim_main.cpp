// sim_main.cpp — the `sim` command line of SANA-FE on the B200 engine.
//
//   sim [-mnopstvxNS] <arch description> <network description> <timesteps>
//
// Mirrors the reference's front-end (src/main.cpp:27-98, src/arg_parsing.cpp:31-147) over the C
// ABI: same flags, same five outputs in the output directory (spikes.csv, potentials.csv,
// perf.csv, messages.csv, run_summary.yaml; writers src/chip.cpp:849-899,1447-1764), including
// the reference's quirk that `-s` alone switches on the spike, potential, perf and message traces
// (src/main.cpp:63-67) while -p / -v / -m are parsed but not consulted; -x writes neurons.csv (model-defined
// neuron traces: LIF `u` of the log_u neurons). Not available: the `cycle` timing model (Booksim2).
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "sanafe_b200.h"

namespace
{
struct Flags
{
    std::filesystem::path output_dir{"."};
    bool record_spikes{false}, record_potentials{false}, record_perf{false}, record_messages{false};
    bool record_neuron_state{false}, use_netlist{false};
    int timing_model{SFE_TIMING_DETAILED}; // the reference's default (src/arg_parsing.hpp)
    int consumed{0};
    int scheduler_threads{0}; // -S: host threads of the detailed timing model (0 = one per host core)
};

int parse_timing(const std::string &name)
{
    if (name == "simple") return SFE_TIMING_SIMPLE;
    if (name == "detailed") return SFE_TIMING_DETAILED;
    if (name == "cycle") return SFE_TIMING_CYCLE;
    throw std::invalid_argument("Error: Timing model not recognized (" + name + ")");
}

[[noreturn]] void fail(const std::string &what)
{
    std::fprintf(stderr, "Error: runtime exception thrown: %s\n", what.c_str());
    std::exit(1);
}
} // namespace

int main(int argc, char *argv[])
{
    std::vector<std::string> args(argv + 1, argv + argc);
    Flags flags;
    size_t idx = 0;
    while (idx < args.size() && (args.size() - idx) > 3 && !args[idx].empty() && args[idx][0] == '-' && args[idx].size() >= 2)
    {
        int used = 1;
        switch (args[idx][1])
        {
        case 'o': flags.output_dir = args.at(idx + 1); used = 2; std::printf("Writing output to %s\n", flags.output_dir.c_str()); break;
        case 'm': flags.record_messages = true; break;
        case 'n': flags.use_netlist = true; break;
        case 'p': flags.record_perf = true; break;
        case 's': flags.record_spikes = true; break;
        case 't':
            try { flags.timing_model = parse_timing(args.at(idx + 1)); }
            catch (const std::exception &e) { std::fprintf(stderr, "%s\n", e.what()); return 1; }
            used = 2;
            break;
        case 'v': flags.record_potentials = true; break;
        case 'x': flags.record_neuron_state = true; break;
        case 'N': used = 2; std::printf("Processing threads: the device engine ignores -N\n"); break;
        case 'S':
            try { flags.scheduler_threads = std::stoi(args.at(idx + 1)); }
            catch (const std::exception &) { std::fprintf(stderr, "Error: invalid scheduler thread count: %s\n", args.at(idx + 1).c_str()); return 1; }
            used = 2;
            std::printf("Scheduler threads: %d\n", flags.scheduler_threads);
            break;
        default: std::printf("Error: Flag %c not recognized.\n", args[idx][1]); break;
        }
        idx += static_cast<size_t>(used);
    }
    if (args.size() - idx < 3)
    {
        std::printf("Usage: ./sim [-mnopstvNS] <arch description> <network description> <timesteps>\n");
        return 0;
    }
    const std::string arch_file = args[idx], net_file = args[idx + 1];
    long timesteps = 0;
    try { timesteps = std::stol(args[idx + 2]); }
    catch (const std::exception &) { std::fprintf(stderr, "Error: invalid argument thrown: Error: Invalid time-step format: %s\n", args[idx + 2].c_str()); return 1; }
    if (timesteps <= 0) { std::fprintf(stderr, "Error: invalid argument thrown: Time-steps must be > 0\n"); return 1; }

    std::printf("Running SANA-FE simulation (B200 engine, ABI %d)\n", sfe_abi_version());
    sfe_arch *arch = sfe_arch_load_yaml(arch_file.c_str());
    if (arch == nullptr) { std::fprintf(stderr, "%s\n", sfe_last_error()); return 1; }
    std::printf("Architecture initialized.\n");
    sfe_net *net = flags.use_netlist ? sfe_net_load_netlist(net_file.c_str(), arch) : sfe_net_load_yaml(net_file.c_str(), arch);
    if (net == nullptr) { std::fprintf(stderr, "%s\n", sfe_last_error()); return 1; }
    std::printf("Network initialized.\n");
    sfe_chip *chip = sfe_chip_create(arch, 0);
    if (chip == nullptr || sfe_chip_load(chip, net) != 0) fail(sfe_last_error());
    const sfe_tables *t = sfe_chip_tables(chip);
    sfe_chip_set_scheduler_threads(chip, flags.scheduler_threads > 0 ? static_cast<uint32_t>(flags.scheduler_threads) : 0u);

    // src/main.cpp:63-67: every trace hangs off -s
    const bool traces = flags.record_spikes;
    std::filesystem::create_directories(flags.output_dir);
    std::ofstream spikes, potentials, perf, messages, neuron_trace;
    size_t n_traces = 0;
    if (flags.record_neuron_state)
    {
        // sim_trace_open_neuron_trace / sim_trace_write_neuron_trace_header  src/chip.cpp:916-929, 1478-1517
        neuron_trace.open(flags.output_dir / "neurons.csv");
        std::string names(sfe_chip_trace_names(chip, nullptr, 0) + 1, '\0');
        sfe_chip_trace_names(chip, names.data(), names.size());
        names.resize(std::strlen(names.c_str()));
        neuron_trace << "timestep,";
        std::istringstream lines(names);
        for (std::string line; std::getline(lines, line); ++n_traces) neuron_trace << "neuron " << line << ",";
        neuron_trace << "\n";
    }
    std::string probe_names;
    if (traces)
    {
        spikes.open(flags.output_dir / "spikes.csv");
        potentials.open(flags.output_dir / "potentials.csv");
        perf.open(flags.output_dir / "perf.csv");
        messages.open(flags.output_dir / "messages.csv");
        spikes << "neuron,timestep\n";
        probe_names.resize(sfe_chip_probe_names(chip, nullptr, 0) + 1);
        sfe_chip_probe_names(chip, probe_names.data(), probe_names.size());
        probe_names.resize(std::strlen(probe_names.c_str()));
        potentials << "timestep,";
        {
            std::istringstream names(probe_names);
            for (std::string line; std::getline(names, line);) potentials << "neuron " << line << ",";
        }
        potentials << "\n";
        perf << "timestep,fired,updated,packets,hops,spikes,sim_time,synapse_energy,dendrite_energy,soma_energy,"
                "network_energy,total_energy\n";
        messages << "timestep,mid,src_neuron,src_hw,dest_hw,hops,spikes,send_timestamp,received_timestamp,"
                    "processed_timestamp,generation_delay,processing_delay,network_delay,blocking_delay,min_hop_delay,"
                    "messages_along_route\n";
    }

    std::printf("Running simulation.\n");
    const size_t n = t->n_neurons, words = (n + 31) / 32, probes = t->n_probes;
    const long chunk = std::max<long>(1, std::min<long>(1024, static_cast<long>((64u << 20) / std::max<size_t>(n, 1))));
    sfe_run_data total{};
    bool first = true;
    std::vector<uint32_t> fired;
    std::vector<double> pots, utraces;
    std::vector<uint8_t> status;
    std::vector<sfe_step_record> steps;
    std::string text;
    for (long done = 0; done < timesteps;)
    {
        const long batch = std::min(chunk, timesteps - done);
        sfe_trace_request req{};
        if (traces)
        {
            fired.assign(static_cast<size_t>(batch) * words, 0u);
            pots.assign(static_cast<size_t>(batch) * probes, 0.0);
            status.assign(static_cast<size_t>(batch) * n, 0);
            steps.assign(static_cast<size_t>(batch), sfe_step_record{});
            req.fired_bits = fired.data();
            req.potentials = probes > 0 ? pots.data() : nullptr;
            req.status = status.data();
            req.steps = steps.data();
        }
        if (n_traces > 0)
        {
            utraces.assign(static_cast<size_t>(batch) * n_traces, 0.0);
            req.neuron_traces = utraces.data();
        }
        sfe_run_data rd{};
        if (sfe_chip_sim(chip, batch, flags.timing_model, (traces || n_traces > 0) ? &req : nullptr, &rd) != 0) fail(sfe_last_error());
        if (flags.record_neuron_state)
            for (long s = 0; s < batch; ++s)
            {
                // sim_trace_record_neuron_traces  src/chip.cpp:1664-1702 (default ostream precision)
                neuron_trace << (rd.timestep_start + s) << ",";
                for (size_t p = 0; p < n_traces; ++p) neuron_trace << utraces[static_cast<size_t>(s) * n_traces + p] << ',';
                if (n_traces > 0) neuron_trace << "\n";
            }
        if (traces)
        {
            text.resize(sfe_chip_format_spikes(chip, fired.data(), batch, rd.timestep_start, nullptr, 0) + 1);
            sfe_chip_format_spikes(chip, fired.data(), batch, rd.timestep_start, text.data(), text.size());
            spikes << text.c_str();
            for (long s = 0; s < batch; ++s)
            {
                potentials << (rd.timestep_start + s) << ",";
                for (size_t p = 0; p < probes; ++p) potentials << pots[static_cast<size_t>(s) * probes + p] << ',';
                if (probes > 0) potentials << "\n";
                const sfe_step_record &r = steps[static_cast<size_t>(s)];
                perf << (rd.timestep_start + s) << "," << r.neurons_fired << "," << r.neurons_updated << "," << r.packets_sent << ","
                     << r.total_hops << "," << r.spike_count << "," << std::scientific << r.sim_time << "," << r.synapse_energy
                     << "," << r.dendrite_energy << "," << r.soma_energy << "," << r.network_energy << "," << r.total_energy << "\n";
            }
            text.resize(sfe_chip_format_messages(chip, status.data(), batch, rd.timestep_start, flags.timing_model, nullptr, 0) + 1);
            sfe_chip_format_messages(chip, status.data(), batch, rd.timestep_start, flags.timing_model, text.data(), text.size());
            messages << text.c_str();
        }
        if (first) total = rd;
        else
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
        first = false;
        done += batch;
    }

    // sim_format_run_summary  src/chip.cpp:872-899 (the wall-time split is this engine's: device + host scheduler)
    auto summary = [&](std::ostream &out) {
        out << "build_git_version: 'sanafe-b200'\n";
        out << "timesteps_executed: " << total.timesteps_executed << "\n";
        out << "total_spikes: " << total.spikes << "\n";
        out << "total_messages_sent: " << total.packets_sent << "\n";
        out << "total_neurons_updated: " << total.neurons_updated << "\n";
        out << "total_neurons_fired: " << total.neurons_fired << "\n";
        out << "sim_time: " << std::scientific << total.sim_time << "\n";
        out << "energy:\n";
        out << "  synapse:" << std::scientific << total.synapse_energy << "\n";
        out << "  dendrite:" << std::scientific << total.dendrite_energy << "\n";
        out << "  soma:" << std::scientific << total.soma_energy << "\n";
        out << "  network: " << std::scientific << total.network_energy << "\n";
        out << "  total: " << std::scientific << total.total_energy << "\n";
        out << "wall_time:\n";
        out << "  device: " << std::fixed << (total.wall_time - total.scheduler_wall_time) << "\n";
        out << "  scheduler: " << std::fixed << total.scheduler_wall_time << "\n";
    };
    std::printf("***** Run Summary *****\n");
    summary(std::cout);
    std::ofstream summary_file(flags.output_dir / "run_summary.yaml");
    if (summary_file.is_open()) summary(summary_file);
    std::printf("Average power consumption: %f W.\n", sfe_chip_get_power(chip));
    std::printf("Run finished.\n");
    sfe_chip_destroy(chip);
    sfe_net_free(net);
    sfe_arch_free(arch);
    return 0;
}
