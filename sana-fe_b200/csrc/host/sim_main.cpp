// sim_main.cpp — the `sim` command line of SANA-FE on the B200 engine.
//
//   sim [-mnopstvxNS] <arch description> <network description> <timesteps>
//
// Mirrors the reference's front-end (src/main.cpp:27-98, src/arg_parsing.cpp:31-147): same flags, same outputs in the
// output directory (spikes.csv, potentials.csv, perf.csv, messages.csv, run_summary.yaml; writers
// src/chip.cpp:849-899,1447-1764), including the reference's quirk that `-s` alone switches on the spike, potential,
// perf and message traces (src/main.cpp:63-67) while -p / -v / -m are parsed but not consulted; -x writes neurons.csv
// (model-defined neuron traces: LIF `u` of the log_u neurons). Not available: the `cycle` timing model (Booksim2).
// Written against the C++ mirror of the reference's user surface (include/sanafe_b200.hpp), as src/main.cpp is
// written against sanafe::SpikingChip.
#include <cstdio>
#include <cstdlib>
#include <filesystem>
#include <iostream>
#include <string>
#include <vector>

#include "sanafe_b200.hpp"

namespace
{
struct Flags
{
    std::filesystem::path output_dir{"."};
    bool record_spikes{false}, record_potentials{false}, record_perf{false}, record_messages{false};
    bool record_neuron_state{false}, use_netlist{false};
    sanafe_b200::TimingModel timing_model{sanafe_b200::timing_model_detailed}; // the reference's default (src/arg_parsing.hpp)
    int scheduler_threads{0}; // -S: host threads of the detailed timing model (0 = one per host core)
};

sanafe_b200::TimingModel parse_timing(const std::string &name)
{
    if (name == "simple") return sanafe_b200::timing_model_simple;
    if (name == "detailed") return sanafe_b200::timing_model_detailed;
    if (name == "cycle") return sanafe_b200::timing_model_cycle_accurate;
    throw std::invalid_argument("Error: Timing model not recognized (" + name + ")");
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
    try
    {
        sanafe_b200::Architecture arch = sanafe_b200::load_arch(arch_file);
        std::printf("Architecture initialized.\n");
        sanafe_b200::SpikingNetwork net = sanafe_b200::load_net(net_file, arch, flags.use_netlist);
        std::printf("Network initialized.\n");
        sanafe_b200::SpikingChip hw(arch);
        hw.load(net);

        // src/main.cpp:63-67: every trace hangs off -s
        sanafe_b200::TraceFlags trace_flags;
        trace_flags.record_spikes = trace_flags.record_potentials = trace_flags.record_perf = trace_flags.record_messages =
                flags.record_spikes;
        trace_flags.record_neuron_state = flags.record_neuron_state;
        std::filesystem::create_directories(flags.output_dir);
        std::printf("Running simulation.\n");
        const sanafe_b200::RunData run_summary = hw.sim(timesteps, flags.timing_model, flags.scheduler_threads, trace_flags, flags.output_dir);
        std::printf("***** Run Summary *****\n");
        sanafe_b200::SpikingChip::format_run_summary(std::cout, run_summary);
        hw.sim_output_run_summary(flags.output_dir, run_summary);
        std::printf("Average power consumption: %f W.\n", hw.get_power());
        std::printf("Run finished.\n");
    }
    catch (const std::invalid_argument &e)
    {
        std::fprintf(stderr, "Error: invalid argument thrown: %s\n", e.what());
        return 1;
    }
    catch (const std::exception &e)
    {
        std::fprintf(stderr, "Error: runtime exception thrown: %s\n", e.what());
        return 1;
    }
    return 0;
}
