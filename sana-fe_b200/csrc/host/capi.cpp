// capi.cpp — the C ABI (include/sanafe_b200.h), description level.
//
// sfe_chip mirrors the reference's SpikingChip (src/chip.hpp:56-107): it is built
// from an Architecture, load() lowers a SpikingNetwork to device tables, sim()
// runs timesteps and returns RunData; state persists across sim() calls and
// reset() zeroes model state without rewinding the timestep counter.
#include <atomic>
#include <charconv>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <mutex>
#include <sstream>
#include <thread>
#include <vector>

#include "../engine.hpp"
#include "desc.hpp"
#include "glibc_rand.hpp"
#include "handles.hpp"
#include "lower.hpp"
#include "schedule.hpp"
#include "sanafe_b200.h"

namespace sfe
{
namespace
{
thread_local std::string g_last_error;
thread_local int g_last_error_kind = SFE_ERROR_RUNTIME;
}
void set_last_error(const std::string &msg)
{
    g_last_error = msg;
    g_last_error_kind = SFE_ERROR_RUNTIME;
}
} // namespace sfe

struct sfe_chip
{
    sfe::Architecture arch;
    int device;
    sfe::HostTables tables;
    bool loaded{false};
    sfe_engine *engine{nullptr};
    // SpikingChip running totals (src/chip.cpp:535-547, 602-621)
    double total_energy{0.0};
    double total_sim_time{0.0};
    uint32_t rank{0}, world{1};
    std::unique_ptr<sfe::DetailedScheduler> scheduler; // built on first use
    // Timesteps are scheduled independently of each other (the NoC state starts empty every step,
    // src/schedule.cpp:208-232), so a batch of steps is spread over host threads: the reference's `-S` /
    // scheduler_threads. 0 = one per host core.
    std::vector<std::unique_ptr<sfe::DetailedScheduler>> scheduler_pool;
    uint32_t scheduler_threads{0};
    long next_mid{0};                                  // total_messages_sent (src/chip.hpp:126): ids of traced messages
    // Poisson inputs: "input" units this process had created before this chip (InputModel::instance_counter is a
    // process-wide static, src/models.hpp:366) and the host generators of this chip's Poisson units
    uint32_t input_seed_base{0};
    sfe_poisson *poisson{nullptr};
    // TrueNorth threshold jitter: the chip's stream of rand() values (the reference draws from the process-global
    // generator, seed 1 unless somebody called srand(): one value per jittered neuron and timestep, in chip order)
    sfe::GlibcRand jitter{1u};
    // per-neuron bias patches (MappedNeuron.set_attributes in a per-frame loop, scripts/tcad2025/dvs_gesture.py) are
    // collected in the host table and uploaded as ONE vector before the next step instead of one copy per neuron
    bool bias_dirty{false};
    // sfe_chip_format_messages: text of the last sizing call (see there)
    std::string sized_messages;
    bool sized_messages_valid{false};
    const uint8_t *sized_status{nullptr};
    int64_t sized_steps{0}, sized_start{0};
    int sized_timing{0};
    long sized_first_mid{0}, sized_next_mid{0};
    explicit sfe_chip(const sfe::Architecture &a, int dev) : arch(a), device(dev) {}
};

namespace
{
template <typename F> auto guarded(F &&f, decltype(f()) on_error) -> decltype(f())
{
    try
    {
        return f();
    }
    // the reference reports errors as C++ exceptions; the class survives the ABI as a "kind" so that bindings can
    // raise what the reference's would (pybind11: invalid_argument -> ValueError, out_of_range -> IndexError,
    // HardwareMappingError registered by name, src/pymodule.cpp:870-871)
    catch (const sfe::HardwareMappingError &e)
    {
        sfe::set_last_error(e.what());
        sfe::g_last_error_kind = SFE_ERROR_HARDWARE_MAPPING;
    }
    catch (const std::invalid_argument &e)
    {
        sfe::set_last_error(e.what());
        sfe::g_last_error_kind = SFE_ERROR_INVALID_ARGUMENT;
    }
    catch (const std::out_of_range &e)
    {
        sfe::set_last_error(e.what());
        sfe::g_last_error_kind = SFE_ERROR_OUT_OF_RANGE;
    }
    catch (const std::exception &e)
    {
        sfe::set_last_error(e.what());
    }
    catch (...)
    {
        sfe::set_last_error("unknown C++ exception");
    }
    return on_error;
}

std::atomic<uint32_t> g_input_units_created{0};

// Runs job(scheduler, s) for s in [0, steps) on the chip's scheduler threads (each thread owns a scheduler);
// timesteps are independent for the timing model (the NoC state starts empty every step).
template <typename Job> void parallel_steps(sfe_chip *c, const int64_t steps, Job &&job)
{
    uint32_t threads = c->scheduler_threads != 0 ? c->scheduler_threads : std::max(1u, std::thread::hardware_concurrency());
    threads = static_cast<uint32_t>(std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(threads, 64), steps / 4)));
    while (c->scheduler_pool.size() < threads)
        c->scheduler_pool.push_back(std::make_unique<sfe::DetailedScheduler>(c->tables.view));
    if (threads == 1)
    {
        for (int64_t s = 0; s < steps; ++s) job(*c->scheduler_pool[0], s);
        return;
    }
    std::atomic<int64_t> next{0};
    std::exception_ptr failure;
    std::mutex mu;
    auto worker = [&](sfe::DetailedScheduler *sched) {
        try
        {
            // small blocks of consecutive steps, handed out dynamically (steps differ in their message counts)
            for (int64_t s0 = next.fetch_add(8); s0 < steps; s0 = next.fetch_add(8))
                for (int64_t s = s0; s < std::min<int64_t>(steps, s0 + 8); ++s) job(*sched, s);
        }
        catch (...)
        {
            const std::lock_guard<std::mutex> lock(mu);
            if (!failure) failure = std::current_exception();
        }
    };
    std::vector<std::thread> pool;
    for (uint32_t w = 1; w < threads; ++w) pool.emplace_back(worker, c->scheduler_pool[w].get());
    worker(c->scheduler_pool[0].get());
    for (std::thread &th : pool) th.join();
    if (failure) std::rethrow_exception(failure);
}

// sim_time of `steps` consecutive timesteps from their status bytes
void schedule_steps(sfe_chip *c, const uint8_t *status, const int64_t steps, double *sim_time)
{
    const size_t n = c->tables.view.n_neurons;
    parallel_steps(c, steps, [&](sfe::DetailedScheduler &sched, const int64_t s) {
        sim_time[s] = sched.schedule_step(status + static_cast<size_t>(s) * n);
    });
}

int flush_bias(sfe_chip *c)
{
    if (!c->bias_dirty || c->engine == nullptr) return 0;
    c->bias_dirty = false;
    return sfe_engine_set_bias_staged(c->engine, c->tables.neuron_bias.data(), c->tables.neuron_bias.size());
}

int attach_engine(sfe_chip *c)
{
    if (c->engine != nullptr)
    {
        sfe_engine_destroy(c->engine);
        c->engine = nullptr;
    }
    // the host schedulers cache per-axon tables of the previous network
    c->scheduler.reset();
    c->scheduler_pool.clear();
    // (a second load() restarts the Poisson streams; the reference's generators would run on)
    sfe_poisson_destroy(c->poisson);
    c->poisson = nullptr;
    c->jitter = sfe::GlibcRand(1u); // (a second load() restarts the jitter stream, like the Poisson streams)
    c->tables.input_seed_base = c->input_seed_base;
    c->tables.view.input_seed_base = c->input_seed_base;
    if (c->tables.view.n_poisson_cols > 0)
    {
        c->poisson = sfe_poisson_create(&c->tables.view);
        if (c->poisson == nullptr) return -1;
    }
    c->total_energy = 0.0;
    c->total_sim_time = 0.0;
    if (c->device < 0) return 0; // host-only chip: lowering + table export
    c->engine = sfe_engine_create_partitioned(&c->tables.view, c->device, c->rank, c->world);
    return c->engine != nullptr ? 0 : -1;
}
} // namespace

extern "C" const char *sfe_last_error(void)
{
    return sfe::g_last_error.c_str();
}

extern "C" int sfe_last_error_kind(void)
{
    return sfe::g_last_error_kind;
}

extern "C" int sfe_abi_version(void)
{
    return SFE_ABI_VERSION;
}

extern "C" sfe_arch *sfe_arch_load_yaml(const char *path)
{
    return guarded(
            [&]() -> sfe_arch * {
                auto a = std::make_unique<sfe_arch>();
                a->arch = sfe::load_arch_yaml(path);
                return a.release();
            },
            nullptr);
}

extern "C" sfe_net *sfe_net_load_yaml(const char *path, sfe_arch *arch)
{
    return guarded(
            [&]() -> sfe_net * {
                if (arch == nullptr || !arch->arch) throw std::invalid_argument("sfe_net_load_yaml: null architecture");
                auto n = std::make_unique<sfe_net>();
                n->net = sfe::load_net_yaml(path, *arch->arch);
                return n.release();
            },
            nullptr);
}

// load_net(path, arch, use_netlist_format = true)  src/network.cpp:194-222, src/netlist.cpp:38-69
extern "C" sfe_net *sfe_net_load_netlist(const char *path, sfe_arch *arch)
{
    return guarded(
            [&]() -> sfe_net * {
                if (arch == nullptr || !arch->arch) throw std::invalid_argument("sfe_net_load_netlist: null architecture");
                auto n = std::make_unique<sfe_net>();
                n->net = sfe::load_net_netlist(path, *arch->arch);
                return n.release();
            },
            nullptr);
}

extern "C" int sfe_load_flat(const char *path, sfe_arch **arch, sfe_net **net)
{
    return guarded(
            [&]() -> int {
                auto a = std::make_unique<sfe_arch>();
                auto n = std::make_unique<sfe_net>();
                sfe::load_flat_file(path, a->arch, n->net, n->synth);
                *arch = a.release();
                *net = n.release();
                return 0;
            },
            -1);
}

extern "C" void sfe_arch_free(sfe_arch *a)
{
    delete a;
}
extern "C" void sfe_net_free(sfe_net *n)
{
    delete n;
}

extern "C" sfe_chip *sfe_chip_create(const sfe_arch *arch, int device)
{
    return guarded(
            [&]() -> sfe_chip * {
                if (arch == nullptr || !arch->arch) throw std::invalid_argument("sfe_chip_create: null architecture");
                if (device >= 0 && sfe_device_count() <= 0)
                    throw std::runtime_error("no CUDA device: the B200 engine has no CPU fallback "
                                             "(pass device < 0 for a host-only chip that can only lower/export tables)");
                auto *chip = new sfe_chip(*arch->arch, device);
                chip->input_seed_base = g_input_units_created.fetch_add(sfe::count_input_units(chip->arch));
                return chip;
            },
            nullptr);
}

extern "C" int sfe_chip_set_partition(sfe_chip *c, uint32_t rank, uint32_t world)
{
    if (world == 0 || rank >= world)
    {
        sfe::set_last_error("sfe_chip_set_partition: need rank < world");
        return -1;
    }
    c->rank = rank;
    c->world = world;
    return 0;
}

extern "C" int sfe_chip_set_input_seed_base(sfe_chip *c, uint32_t base)
{
    if (c->loaded)
    {
        sfe::set_last_error("sfe_chip_set_input_seed_base: call before load");
        return -1;
    }
    c->input_seed_base = base;
    return 0;
}

extern "C" void sfe_chip_destroy(sfe_chip *c)
{
    if (c == nullptr) return;
    if (c->engine != nullptr) sfe_engine_destroy(c->engine);
    sfe_poisson_destroy(c->poisson);
    delete c;
}

extern "C" int sfe_chip_load(sfe_chip *c, const sfe_net *net)
{
    return guarded(
            [&]() -> int {
                if (net == nullptr) throw std::invalid_argument("sfe_chip_load: null network");
                if (net->synth.has_value()) sfe::lower_synthetic(c->arch, *net->synth, true, c->tables);
                else sfe::lower_network(c->arch, *net->net, c->tables);
                c->loaded = true;
                return attach_engine(c);
            },
            -1);
}

extern "C" int sfe_chip_load_synthetic(sfe_chip *c, const sfe_synth_spec *spec, int generate_on_device)
{
    return guarded(
            [&]() -> int {
                sfe::SynthRequest req;
                req.spec = *spec;
                // unit names follow BASELINE config 4
                req.soma_hw_name = "loihi_lif";
                req.synapse_hw_name = "loihi_dense_synapse";
                req.dendrite_hw_name = "loihi_dendrites_delay";
                sfe::lower_synthetic(c->arch, req, generate_on_device == 0, c->tables);
                c->loaded = true;
                return attach_engine(c);
            },
            -1);
}

extern "C" int sfe_chip_sim(sfe_chip *c, int64_t timesteps, int timing_model, const sfe_trace_request *req,
        sfe_run_data *out)
{
    return guarded(
            [&]() -> int {
                if (!c->loaded) throw std::runtime_error("sfe_chip_sim: no network loaded");
                if (c->engine == nullptr)
                    throw std::runtime_error("no CUDA device: the B200 engine has no CPU fallback");
                const auto wall0 = std::chrono::steady_clock::now();
                sfe_engine_request_stop(c->engine, 0);
                if (flush_bias(c) != 0) return -1;
                if (timing_model == SFE_TIMING_CYCLE)
                    throw std::runtime_error("the cycle-accurate timing model needs Booksim2 (third-party, not "
                                             "available): out of scope");
                sfe_run_data rd;
                const bool detailed = timing_model != SFE_TIMING_SIMPLE;
                const uint32_t rand_cols = c->tables.view.n_rand_cols;
                if (!detailed && c->poisson == nullptr && rand_cols == 0)
                {
                    if (sfe_engine_run(c->engine, timesteps, req, &rd) != 0) return -1;
                }
                else
                {
                    // Chunked run with per-step records.
                    // Detailed model: the device runs the timesteps and hands back one status byte per neuron and
                    // step; the host scheduler (src/schedule.cpp:208-620 restated in schedule.cpp) turns them into
                    // per-step sim_time.
                    // Poisson inputs: the host draws the chunk's random spikes with the reference's generator
                    // (poisson.cpp) and uploads them as an overlay before the chunk is enqueued.
                    if (detailed && c->world > 1) throw std::runtime_error("detailed timing is not available on a partitioned chip");
                    const size_t n = c->tables.view.n_neurons;
                    const size_t words = (n + 31) / 32;
                    const uint32_t cols = c->poisson != nullptr ? sfe_poisson_cols(c->poisson) : 0u;
                    int64_t batch_cap = 4096;
                    if (detailed) batch_cap = std::min<int64_t>(batch_cap, (64ll << 20) / static_cast<int64_t>(std::max<size_t>(n, 1)));
                    if (cols > 0) batch_cap = std::min<int64_t>(batch_cap, (32ll << 20) / static_cast<int64_t>(cols));
                    if (rand_cols > 0) batch_cap = std::min<int64_t>(batch_cap, (8ll << 20) / static_cast<int64_t>(rand_cols));
                    batch_cap = std::max<int64_t>(1, batch_cap);
                    std::vector<uint8_t> status, overlay;
                    std::vector<uint32_t> jitter;
                    std::vector<double> sched_time;
                    std::vector<sfe_step_record> recs;
                    std::memset(&rd, 0, sizeof(rd));
                    double sched_s = 0.0;
                    for (int64_t done = 0; done < timesteps;)
                    {
                        const int64_t batch = std::min<int64_t>(batch_cap, timesteps - done);
                        recs.resize(static_cast<size_t>(batch));
                        sfe_trace_request sub{};
                        sub.steps = recs.data();
                        if (detailed)
                        {
                            status.resize(static_cast<size_t>(batch) * n);
                            sub.status = status.data();
                        }
                        else if (req != nullptr && req->status != nullptr) sub.status = req->status + static_cast<size_t>(done) * n;
                        if (req != nullptr)
                        {
                            if (req->fired_bits != nullptr) sub.fired_bits = req->fired_bits + static_cast<size_t>(done) * words;
                            if (req->potentials != nullptr) sub.potentials = req->potentials + static_cast<size_t>(done) * c->tables.view.n_probes;
                            if (req->neuron_traces != nullptr)
                                sub.neuron_traces = req->neuron_traces + static_cast<size_t>(done) * c->tables.view.n_u_probes;
                        }
                        // draws on the device by default (sfe_engine_fill_input_overlay); SFE_DEVICE_POISSON=0: on the
                        // host with libstdc++'s generator, the cross-check
                        const char *opt = std::getenv("SFE_DEVICE_POISSON");
                        const bool device_draws = opt == nullptr || std::atoi(opt) != 0;
                        if (c->poisson != nullptr && device_draws)
                        {
                            if (sfe_engine_fill_input_overlay(c->engine, batch) != 0) return -1;
                        }
                        else if (c->poisson != nullptr)
                        {
                            overlay.resize(static_cast<size_t>(batch) * cols);
                            if (sfe_poisson_fill(c->poisson, overlay.data(), batch) != 0) return -1;
                            if (sfe_engine_set_input_overlay(c->engine, overlay.data(), batch, cols) != 0) return -1;
                        }
                        if (rand_cols > 0)
                        {
                            // TrueNorthModel::truenorth_threshold_and_reset  src/models.cpp:749-759: one rand() per
                            // jittered neuron and timestep, in the order a single processing thread updates them
                            jitter.resize(static_cast<size_t>(batch) * rand_cols);
                            for (uint32_t &x : jitter) x = c->jitter.next();
                            if (sfe_engine_set_rand_overlay(c->engine, jitter.data(), batch, rand_cols) != 0) return -1;
                        }
                        sfe_run_data part;
                        if (sfe_engine_run(c->engine, batch, &sub, &part) != 0) return -1;
                        if (done == 0) rd.timestep_start = part.timestep_start;
                        const auto s0 = std::chrono::steady_clock::now();
                        if (detailed)
                        {
                            sched_time.resize(static_cast<size_t>(batch));
                            schedule_steps(c, status.data(), batch, sched_time.data());
                        }
                        for (int64_t b = 0; b < batch; ++b)
                        {
                            sfe_step_record &r = recs[static_cast<size_t>(b)];
                            if (detailed) r.sim_time = sched_time[static_cast<size_t>(b)];
                            // update_run_data  src/chip.cpp:462-475
                            rd.total_energy += r.total_energy;
                            rd.synapse_energy += r.synapse_energy;
                            rd.dendrite_energy += r.dendrite_energy;
                            rd.soma_energy += r.soma_energy;
                            rd.network_energy += r.network_energy;
                            rd.sim_time += r.sim_time;
                            rd.spikes += r.spike_count;
                            rd.packets_sent += r.packets_sent;
                            rd.neurons_updated += r.neurons_updated;
                            rd.neurons_fired += r.neurons_fired;
                            if (req != nullptr && req->steps != nullptr) req->steps[done + b] = r;
                            if (detailed && req != nullptr && req->status != nullptr)
                                std::memcpy(req->status + static_cast<size_t>(done + b) * n, status.data() + static_cast<size_t>(b) * n, n);
                        }
                        if (detailed) sched_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - s0).count();
                        done += batch;
                    }
                    rd.timesteps_executed = timesteps;
                    rd.scheduler_wall_time = sched_s;
                }
                rd.wall_time = std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count();
                c->total_energy += rd.total_energy;   // sim_update_total_energy_and_counts
                c->total_sim_time += rd.sim_time;     // retire_timestep
                if (out != nullptr) *out = rd;
                return 0;
            },
            -1);
}

namespace
{
// n independent jobs over a pool of host threads; job(k) returns 0 or -1 (+ the thread's last error)
template <typename Job> int run_batch(const uint32_t n, uint32_t host_threads, const char *what, Job &&job)
{
    // default: one worker per host core (a worker spins in cudaStreamSynchronize while its chip's batch runs)
    if (host_threads == 0) host_threads = std::min<uint32_t>(n, std::max(1u, std::thread::hardware_concurrency()));
    host_threads = std::max<uint32_t>(1u, std::min<uint32_t>(host_threads, n));
    std::atomic<uint32_t> next{0};
    std::mutex mu;
    uint32_t first_bad = UINT32_MAX;
    std::string first_msg;
    auto worker = [&]() {
        for (uint32_t k = next.fetch_add(1); k < n; k = next.fetch_add(1))
        {
            int rc = -1;
            try
            {
                rc = job(k);
            }
            catch (const std::exception &e)
            {
                sfe::set_last_error(e.what());
            }
            if (rc != 0)
            {
                const std::lock_guard<std::mutex> lock(mu);
                if (k < first_bad)
                {
                    first_bad = k;
                    first_msg = sfe_last_error();
                }
            }
        }
    };
    std::vector<std::thread> pool;
    for (uint32_t w = 1; w < host_threads; ++w) pool.emplace_back(worker);
    worker();
    for (std::thread &t : pool) t.join();
    if (first_bad == UINT32_MAX) return 0;
    sfe::set_last_error(std::string(what) + ": chip " + std::to_string(first_bad) + ": " + first_msg);
    return -1;
}
} // namespace

extern "C" int sfe_batch_load(sfe_chip *const *chips, const sfe_net *const *nets, uint32_t n, uint32_t host_threads)
{
    if (n == 0) return 0;
    if (chips == nullptr || nets == nullptr)
    {
        sfe::set_last_error("sfe_batch_load: null argument");
        return -1;
    }
    return run_batch(n, host_threads, "sfe_batch_load", [&](uint32_t k) { return sfe_chip_load(chips[k], nets[k]); });
}

extern "C" int sfe_batch_sim(sfe_chip *const *chips, uint32_t n, int64_t timesteps, int timing_model,
        const sfe_trace_request *reqs, sfe_run_data *out, uint32_t host_threads)
{
    if (n == 0) return 0;
    if (chips == nullptr)
    {
        sfe::set_last_error("sfe_batch_sim: null argument");
        return -1;
    }
    for (uint32_t a = 0; a < n; ++a)
        for (uint32_t b = a + 1; b < n; ++b)
            if (chips[a] == chips[b])
            {
                sfe::set_last_error("sfe_batch_sim: the same chip appears twice");
                return -1;
            }
    // SFE_BATCH_GRID=1: the whole batch as one launch per phase and step (sfe_engine_batch_enqueue: the concatenation
    // of the chips' grids) when only totals are asked for and the chips are in lock step. Measured on a B200 with 128
    // design points of the config-5 sweep it is the slower of the two (609 against 809 simulations/s: every phase
    // becomes a barrier across all chips, and the batched kernels read their tables through a pointer), so one stream
    // per chip with a pool of host threads stays the default.
    {
        const char *opt = std::getenv("SFE_BATCH_GRID");
        bool grid = (opt != nullptr && std::atoi(opt) != 0) && n > 1 && timing_model == SFE_TIMING_SIMPLE && timesteps > 0;
        for (uint32_t k = 0; k < n && grid; ++k)
        {
            const sfe_chip *c = chips[k];
            grid = c != nullptr && c->loaded && c->engine != nullptr && c->poisson == nullptr && c->world == 1 &&
                    c->tables.view.n_rand_cols == 0;
            if (grid && reqs != nullptr)
                grid = reqs[k].steps == nullptr && reqs[k].fired_bits == nullptr && reqs[k].potentials == nullptr &&
                        reqs[k].status == nullptr && reqs[k].neuron_traces == nullptr;
        }
        if (grid)
        {
            const int rc = guarded(
                    [&]() -> int {
                        const auto wall0 = std::chrono::steady_clock::now();
                        std::vector<sfe_engine *> engines(n);
                        std::vector<sfe_run_data> total(n);
                        for (uint32_t k = 0; k < n; ++k)
                        {
                            sfe_engine_request_stop(chips[k]->engine, 0);
                            if (flush_bias(chips[k]) != 0) return -1;
                            engines[k] = chips[k]->engine;
                            std::memset(&total[k], 0, sizeof(sfe_run_data));
                        }
                        for (int64_t done = 0; done < timesteps;)
                        {
                            const int64_t chunk = std::min<int64_t>(2048, timesteps - done); // (the device log holds 4096 steps)
                            const int enq = sfe_engine_batch_enqueue(engines.data(), n, chunk);
                            if (enq < 0) return -1;
                            if (enq > 0)
                            {
                                if (done == 0) return 1; // not a batch the grid form can run: nothing has been stepped
                                throw std::runtime_error("sfe_batch_sim: the chips of the batch fell out of lock step");
                            }
                            for (uint32_t k = 0; k < n; ++k)
                            {
                                sfe_run_data part;
                                if (sfe_engine_collect(engines[k], &part) != 0) return -1;
                                if (done == 0) total[k].timestep_start = part.timestep_start;
                                total[k].timesteps_executed += part.timesteps_executed;
                                total[k].total_energy += part.total_energy;
                                total[k].synapse_energy += part.synapse_energy;
                                total[k].dendrite_energy += part.dendrite_energy;
                                total[k].soma_energy += part.soma_energy;
                                total[k].network_energy += part.network_energy;
                                total[k].sim_time += part.sim_time;
                                total[k].spikes += part.spikes;
                                total[k].packets_sent += part.packets_sent;
                                total[k].neurons_updated += part.neurons_updated;
                                total[k].neurons_fired += part.neurons_fired;
                            }
                            done += chunk;
                        }
                        const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count();
                        for (uint32_t k = 0; k < n; ++k)
                        {
                            total[k].wall_time = wall;
                            chips[k]->total_energy += total[k].total_energy; // sim_update_total_energy_and_counts
                            chips[k]->total_sim_time += total[k].sim_time;   // retire_timestep
                            if (out != nullptr) out[k] = total[k];
                        }
                        return 0;
                    },
                    -1);
            if (rc <= 0) return rc;
        }
    }
    return run_batch(n, host_threads, "sfe_batch_sim", [&](uint32_t k) {
        return sfe_chip_sim(chips[k], timesteps, timing_model, reqs != nullptr ? &reqs[k] : nullptr,
                out != nullptr ? &out[k] : nullptr);
    });
}

extern "C" int sfe_chip_set_scheduler_threads(sfe_chip *c, uint32_t threads)
{
    c->scheduler_threads = threads;
    return 0;
}

extern "C" void sfe_chip_request_stop(sfe_chip *c)
{
    if (c != nullptr && c->engine != nullptr) sfe_engine_request_stop(c->engine, 1);
}

extern "C" int sfe_chip_reset(sfe_chip *c)
{
    if (c->engine == nullptr)
    {
        sfe::set_last_error("sfe_chip_reset: chip has no device engine");
        return -1;
    }
    return sfe_engine_reset(c->engine);
}

extern "C" double sfe_chip_get_power(sfe_chip *c)
{
    return c->total_sim_time > 0.0 ? c->total_energy / c->total_sim_time : 0.0; // src/chip.cpp:607-621
}

extern "C" const sfe_tables *sfe_chip_tables(const sfe_chip *c)
{
    return c->loaded ? &c->tables.view : nullptr;
}

extern "C" sfe_engine *sfe_chip_engine(sfe_chip *c)
{
    if (c != nullptr) flush_bias(c); // callers that step the engine themselves see the patched biases
    return c->engine;
}

extern "C" int64_t sfe_chip_neuron_index(const sfe_chip *c, const char *group, uint64_t offset)
{
    return c->tables.find_neuron(group, offset);
}

// MappedNeuron::set_attributes(..., log_spikes)  src/mapped.cpp:113-124
extern "C" int sfe_chip_set_neuron_log_spikes(sfe_chip *c, const char *group, uint64_t offset, int on)
{
    return guarded(
            [&]() -> int {
                const int64_t idx = c->tables.find_neuron(group, offset);
                if (idx < 0) throw std::out_of_range(std::string("no mapped neuron ") + group + "." + std::to_string(offset));
                c->tables.names[static_cast<size_t>(idx)].log_spikes = on != 0;
                return 0;
            },
            -1);
}

// SpikingChip::mapped_neuron_groups  src/chip.hpp:104: "name<TAB>neuron count" per line, lexicographic order
extern "C" size_t sfe_chip_group_names(const sfe_chip *c, char *buf, size_t cap)
{
    const sfe::HostTables &t = c->tables;
    std::string out;
    for (size_t g = 0; g < t.group_names.size(); ++g)
        out += t.group_names[g] + "\t" + std::to_string(t.group_to_device[g].size()) + "\n";
    if (buf != nullptr && cap > 0)
    {
        const size_t n = std::min(cap - 1, out.size());
        std::memcpy(buf, out.data(), n);
        buf[n] = '\0';
    }
    return out.size();
}

extern "C" sfe_net *sfe_net_create(const char *name)
{
    return guarded(
            [&]() -> sfe_net * {
                auto n = std::make_unique<sfe_net>();
                n->net = std::make_unique<sfe::SpikingNetwork>(name != nullptr ? name : "");
                return n.release();
            },
            nullptr);
}

extern "C" int sfe_net_save_yaml(const sfe_net *net, const char *path)
{
    return guarded(
            [&]() -> int {
                if (net == nullptr || !net->net) throw std::invalid_argument("sfe_net_save_yaml: no described network");
                sfe::save_net_yaml(*net->net, path);
                return 0;
            },
            -1);
}

// SpikingNetwork::save(path, use_netlist_format = true)  src/network.cpp:606-703
extern "C" int sfe_net_save_netlist(const sfe_net *net, const char *path)
{
    return guarded(
            [&]() -> int {
                if (net == nullptr || !net->net) throw std::invalid_argument("sfe_net_save_netlist: no described network");
                sfe::save_net_netlist(*net->net, path);
                return 0;
            },
            -1);
}

extern "C" int sfe_chip_set_neuron_attribute(sfe_chip *c, const char *group, uint64_t offset, const char *name,
        double value)
{
    return guarded(
            [&]() -> int {
                const int64_t idx = c->tables.find_neuron(group, offset);
                if (idx < 0) throw std::out_of_range(std::string("no mapped neuron ") + group + "." + std::to_string(offset));
                const sfe::PatchKind kind = sfe::patch_neuron_attribute(c->tables, static_cast<uint32_t>(idx), name, value);
                if (c->engine == nullptr) return 0;
                if (kind == sfe::PatchKind::bias) c->bias_dirty = true; // uploaded with the next sfe_chip_sim / sfe_chip_engine call
                else if (kind == sfe::PatchKind::classes)
                    return sfe_engine_update_classes(c->engine, c->tables.soma_classes.data(),
                            static_cast<uint32_t>(c->tables.soma_classes.size()), c->tables.neuron_class.data());
                else if (kind == sfe::PatchKind::potential)
                    return sfe_engine_set_neuron_potential(c->engine, static_cast<uint32_t>(idx), value);
                return 0;
            },
            -1);
}

// Spike rows in the reference's trace order and format (src/chip.cpp:1610-1630):
// for every step, groups in lexicographic order, offsets ascending, only neurons
// with log_spikes. Returns the number of bytes needed (excluding the NUL).
extern "C" size_t sfe_chip_format_spikes(const sfe_chip *c, const uint32_t *fired_bits, int64_t timesteps,
        int64_t timestep_start, char *buf, size_t cap)
{
    const sfe::HostTables &t = c->tables;
    const size_t words = (t.names.size() + 31) / 32;
    std::string out;
    for (int64_t s = 0; s < timesteps; ++s)
    {
        const uint32_t *bits = fired_bits + static_cast<size_t>(s) * words;
        for (size_t g = 0; g < t.group_names.size(); ++g)
        {
            for (size_t o = 0; o < t.group_to_device[g].size(); ++o)
            {
                const uint32_t i = t.group_to_device[g][o];
                if (t.names[i].log_spikes && ((bits[i >> 5] >> (i & 31)) & 1u))
                {
                    out += t.group_names[g];
                    out += '.';
                    out += std::to_string(o);
                    out += ',';
                    out += std::to_string(timestep_start + s);
                    out += '\n';
                }
            }
        }
    }
    if (buf != nullptr && cap > 0)
    {
        const size_t n = std::min(cap - 1, out.size());
        std::memcpy(buf, out.data(), n);
        buf[n] = '\0';
    }
    return out.size();
}

// Rows of messages.csv (sim_trace_record_message, src/chip.cpp:1731-1764) for `timesteps` steps,
// rebuilt from the per-neuron status bytes of those steps; every other message field is a
// load-time constant of the lowered tables. SFE_TIMING_DETAILED runs the NoC scheduler so that
// the timestamps / network / blocking delays are filled; SFE_TIMING_SIMPLE leaves them at their
// defaults (-inf / 0), like the reference. Message ids continue the chip's lifetime counter
// (total_messages_sent, src/chip.cpp:815) in the single-thread creation order.
extern "C" size_t sfe_chip_format_messages(sfe_chip *c, const uint8_t *status, int64_t timesteps, int64_t timestep_start,
        int timing_model, char *buf, size_t cap)
{
    return guarded(
            [&]() -> size_t {
                if (!c->loaded) throw std::runtime_error("sfe_chip_format_messages: no network loaded");
                const sfe::HostTables &t = c->tables;
                const size_t n = t.view.n_neurons;
                // The sizing call (buf == NULL) and the call that fetches the text do the same work: keep the text of
                // the sizing call and hand it out if the very next call asks for the same rows.
                const bool same_as_sized = c->sized_messages_valid && c->sized_status == status && c->sized_steps == timesteps &&
                        c->sized_start == timestep_start && c->sized_timing == timing_model && c->sized_first_mid == c->next_mid;
                if (buf != nullptr && cap > 0 && same_as_sized)
                {
                    const std::string &text = c->sized_messages;
                    const size_t k = std::min(cap - 1, text.size());
                    std::memcpy(buf, text.data(), k);
                    buf[k] = '\0';
                    c->next_mid = c->sized_next_mid;
                    c->sized_messages_valid = false;
                    const size_t size = text.size();
                    c->sized_messages.clear();
                    c->sized_messages.shrink_to_fit();
                    return size;
                }
                c->sized_messages_valid = false;
                // Message ids run on from step to step (total_messages_sent, src/chip.cpp:815): the id a step starts
                // with is known from the spikes alone (one message per axon-out of every fired neuron), so the steps
                // can be traced and formatted independently, on the scheduler threads, and concatenated.
                const long first_mid = c->next_mid;
                std::vector<long> start_mid(static_cast<size_t>(timesteps) + 1, first_mid);
                for (int64_t s = 0; s < timesteps; ++s)
                {
                    const uint8_t *st = status + static_cast<size_t>(s) * n;
                    long messages = 0;
                    for (size_t i = 0; i < n; ++i)
                        if (st[i] == SFE_STATUS_FIRED) messages += t.view.axon_out_begin[i + 1] - t.view.axon_out_begin[i];
                    start_mid[static_cast<size_t>(s) + 1] = start_mid[static_cast<size_t>(s)] + messages;
                }
                std::vector<std::string> texts(static_cast<size_t>(timesteps));
                parallel_steps(c, timesteps, [&](sfe::DetailedScheduler &sched, const int64_t s) {
                    std::string &out = texts[static_cast<size_t>(s)];
                    // numbers exactly as an ostream with default settings prints them (src/chip.cpp:1731-1764 streams
                    // the fields with operator<<): integers in decimal, doubles as %g with 6 significant digits
                    auto put_int = [&out](const long long v) {
                        char tmp[24];
                        const auto r = std::to_chars(tmp, tmp + sizeof(tmp), v);
                        out.append(tmp, r.ptr);
                    };
                    auto put_double = [&out](const double v) {
                        char tmp[40];
                        const auto r = std::to_chars(tmp, tmp + sizeof(tmp), v, std::chars_format::general, 6);
                        out.append(tmp, r.ptr);
                    };
                    thread_local std::vector<sfe::MessageRecord> recs; // capacity kept from step to step
                    recs.clear();
                    long mid = start_mid[static_cast<size_t>(s)];
                    sched.trace_step(status + static_cast<size_t>(s) * n, timing_model == SFE_TIMING_DETAILED, recs, mid);
                    if (mid != start_mid[static_cast<size_t>(s) + 1]) throw std::logic_error("message count of a step disagrees with its spikes");
                    out.reserve(recs.size() * 128);
                    for (const sfe::MessageRecord &m : recs)
                    {
                        const sfe::HostTables::NeuronName &nm = t.names[m.src_neuron];
                        put_int(timestep_start + s);
                        out += ',';
                        put_int(m.mid);
                        out += ',';
                        out += t.group_names[nm.group];
                        out += '.';
                        put_int(nm.offset);
                        out += ',';
                        out += t.core_names[m.src_core];
                        out += ',';
                        if (m.placeholder) out += "x.x,";
                        else
                        {
                            out += t.core_names[m.dest_core];
                            out += ',';
                        }
                        put_int(m.hops);
                        out += ',';
                        put_int(m.spikes);
                        for (const double v : {m.sent, m.received, m.processed, m.generation_delay, m.processing_delay,
                                     m.network_delay, m.blocking_delay, m.min_hop_delay, m.messages_along_route})
                        {
                            out += ',';
                            put_double(v);
                        }
                        out += '\n';
                    }
                });
                c->next_mid = start_mid[static_cast<size_t>(timesteps)];
                std::string out;
                {
                    size_t total = 0;
                    for (const std::string &x : texts) total += x.size();
                    out.reserve(total);
                    for (const std::string &x : texts) out += x;
                }
                const std::string &text = out;
                if (buf != nullptr && cap > 0)
                {
                    const size_t k = std::min(cap - 1, text.size());
                    std::memcpy(buf, text.data(), k);
                    buf[k] = '\0';
                }
                else
                {
                    // sizing call: the ids are not consumed; remember the text for the call that fetches it
                    c->sized_next_mid = c->next_mid;
                    c->next_mid = first_mid;
                    c->sized_status = status;
                    c->sized_steps = timesteps;
                    c->sized_start = timestep_start;
                    c->sized_timing = timing_model;
                    c->sized_first_mid = first_mid;
                    const size_t size = out.size();
                    c->sized_messages = std::move(out);
                    c->sized_messages_valid = true;
                    return size;
                }
                return text.size();
            },
            static_cast<size_t>(0));
}

// header of potentials.csv (src/chip.cpp:1454-1476): "timestep,neuron g.o,..."
extern "C" size_t sfe_chip_probe_names(const sfe_chip *c, char *buf, size_t cap)
{
    const sfe::HostTables &t = c->tables;
    std::string out;
    for (uint32_t i : t.probes)
    {
        out += t.group_names[t.names[i].group];
        out += '.';
        out += std::to_string(t.names[i].offset);
        out += '\n';
    }
    if (buf != nullptr && cap > 0)
    {
        const size_t n = std::min(cap - 1, out.size());
        std::memcpy(buf, out.data(), n);
        buf[n] = '\0';
    }
    return out.size();
}

extern "C" size_t sfe_chip_trace_names(const sfe_chip *c, char *buf, size_t cap)
{
    const sfe::HostTables &t = c->tables;
    std::string out;
    for (uint32_t i : t.u_probes)
    {
        out += t.group_names[t.names[i].group];
        out += '.';
        out += std::to_string(t.names[i].offset);
        out += "/u\n";
    }
    if (buf != nullptr && cap > 0)
    {
        const size_t n = std::min(cap - 1, out.size());
        std::memcpy(buf, out.data(), n);
        buf[n] = '\0';
    }
    return out.size();
}

// Detailed timing model over a caller-supplied status trace ([timesteps][n_neurons]
// SFE_STATUS_* bytes, e.g. sfe_trace_request.status of an earlier run): per-step
// sim_time of schedule_messages_timestep_detailed (src/schedule.cpp:208-292). Pure
// host code (the scheduler is host-side in the reference too).
extern "C" int sfe_chip_schedule_detailed(sfe_chip *c, const uint8_t *status, int64_t timesteps, double *sim_time)
{
    return guarded(
            [&]() -> int {
                if (!c->loaded) throw std::runtime_error("sfe_chip_schedule_detailed: no network loaded");
                schedule_steps(c, status, timesteps, sim_time);
                return 0;
            },
            -1);
}
