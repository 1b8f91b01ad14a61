// engine.cu — the per-timestep hot path of SANA-FE as hand-written sm_100a kernels.
//
// Replaces SpikingChip::sim_hw_timestep and everything under it (reference
// src/chip.cpp:1053-1108): one timestep = three kernels on one stream, launched with
// programmatic dependent launch, no host synchronisation inside a run.
//
//   soma_kernel      neuron phase           (src/chip.cpp:624-654,710-736,802-834; models.cpp)
//                    CTA per segment of 512 neurons of a core, two neurons per thread, every
//                    load issued before the first use; SoA state; warp-ballot spike raster;
//                    fired neurons raise the inbox bits of their axons; warp/block reductions
//                    give the per-segment counters / energy / generation-delay sum
//   fanout_kernel    message phase          (src/chip.cpp:656-764,1127-1169; models.cpp:29-131)
//                    persistent CTAs draw work items (inbox slices of destination cores) from
//                    a ticket; inbox bitmask (= message arrival order) -> CTA-wide list of
//                    active axons -> their CSR segments streamed from HBM (4-byte lossless
//                    records through per-warp cp.async rings, or fp64 weight + u32 meta through
//                    TMA bulk copies) -> shared-memory dendrite accumulators (exact fixed point,
//                    or ordered fp64 when the weights are not provably order-independent).
//                    On a partitioned chip the kernel starts with the raster exchange over
//                    peer memory and gathers its inbox from the exchanged raster.
//   finalize_kernel  energy, counters, simple timing (src/chip.cpp:1028-1051,1171-1261;
//                    src/schedule.cpp:61-102); appends one sfe_step_record to the device log
//
// Load-time kernels: synth_generate_kernel, certify_kernel (exactness certificate),
// pack_q4_kernel (4-byte records). The host half of the engine (tables, work items, streams,
// the C ABI of include/sanafe_b200.h, NCCL / CUDA-IPC plumbing) follows the kernels.
//
// All arithmetic that decides a spike is IEEE fp64 without FMA contraction
// (compiled with -fmad=false), so rasters are bit-identical to the reference.
// There is no CPU fallback: every entry point fails when no CUDA device exists.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "engine.hpp"
#include "mt19937.cuh"
#include "sanafe_b200.h"
#include "sfe_device_model.h"

namespace sfe
{
void set_last_error(const std::string &msg);
}

namespace
{
// ---------------------------------------------------------------------------
// Device-side tables
// ---------------------------------------------------------------------------
struct CoreDev
{
    uint32_t neuron_begin, neuron_count;
    uint32_t axon_begin, axon_count;  // into axons_in[] (unpadded global ids)
    uint32_t inbox_word_begin;        // padded: each core's inbox starts on a word
    uint32_t fired_word_begin;        // padded: each core's raster starts on a word
    uint32_t dend_base;               // into din*: layout [slot][k] per core
    uint32_t ring;                    // delay-ring slots
    uint32_t acc_mode;                // SFE_ACC_*
    uint32_t dend_in_msg;
    uint32_t fixed_slots;             // the delay field of a synapse is its dendrite slot ("neurofem" compartments)
    uint32_t fast_soma, pad_fast;     // fused step kernel: the core's neuron phase runs through soma_core
    uint32_t tile;
    uint32_t seg_begin, seg_count;    // neuron-phase segments of this core
    uint32_t item_begin, item_count;  // message-phase work items (inbox slices) of this core
    uint32_t q4;                      // synapses of this core also exist as 4-byte records (syn_q4)
    uint32_t active_idx;              // position in the active-core list (per-core step partials)
    uint32_t acc_global;              // 1: the core's dendrite cells do not fit in shared memory - the message phase
                                      // accumulates straight into the HBM arrays (din32 / dcnt32 / din64)
    unsigned long long syn_begin;
    double scale, inv_scale;          // 2^shift, 2^-shift
    double lat_axon_in, e_axon_in, lat_axon_out, e_axon_out;
    double e_east, e_west, e_south, e_north; // hop energies of the core's (destination) tile
};

struct StatsN // written by soma_kernel, one per core
{
    uint32_t updated, fired, packets, pad;
    double soma_e, dend_e, gen_sum;
    double dup_e; // energy of combined dendrite + soma units: part of soma_e AND dend_e, counted once in the total
};

struct FanItem // one unit of message-phase work: a slice of one core's inbox
{
    uint32_t core, word_lo, word_hi, pad;
};

struct StatsM // written by fanout_kernel, one per work item
{
    uint32_t msgs, pad;
    unsigned long long events, hop_e, hop_w, hop_n, hop_s;
    double syn_e, den_e, proc;
    double dup_e; // part of den_e spent in combined dendrite + soma units (also soma energy)
};

struct SomaSegment;
// "taps" dendrites (MultiTapModel1D, src/models.cpp:167-259): the
// currents of a tap line are added onto evolving fp64 state in arrival order, so each line is replayed by one
// thread of a small kernel of its own (taps_kernel) from the step's fired raster; the message phase still does
// the accounting of those events and the neuron phase reads the line's output instead of its accumulator cell.
struct TapsUnit
{
    uint32_t n_taps, const_off; // time constants [n_taps] then space constants [n_taps-1] in taps_values
    uint32_t state_off;         // the line's voltages in DevState::tap_v / tap_next
    uint32_t syn_begin, syn_count; // its incoming synapses in arrival order (taps_syn)
    uint32_t pad;
};
struct TapsSyn
{
    double weight;
    uint32_t src_bit; // raster bit of the source neuron
    uint32_t tap;
    uint32_t post;    // device index of the neuron the synapse belongs to (neurons of a unit share its line)
    uint32_t pad;
};

// Poisson inputs drawn on the device (the default; SFE_DEVICE_POISSON=0 draws on the host with libstdc++,
// csrc/host/poisson.cpp, the cross-check): one MT19937 per Poisson input unit (csrc/mt19937.cuh, states
// interleaved), one thread per unit filling the unit's columns of the overlay for a chunk of timesteps.
struct PoissonUnit
{
    double probability;
    uint32_t share_count;
    uint32_t col_begin; // the unit's col_of_rank[share_count] in poisson_cols
};

__global__ void poisson_kernel(const PoissonUnit *units, const uint32_t *cols, uint32_t n_units, uint32_t *mt, uint32_t *idx,
        uint8_t *overlay, const long long n_steps, const uint32_t n_cols)
{
    const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_units) return;
    const PoissonUnit d = units[u];
    uint32_t cursor = idx[u];
    for (long long step = 0; step < n_steps; ++step)
        for (uint32_t r = 0; r < d.share_count; ++r)
        {
            // InputModel::update draws on every update of the unit (src/models.cpp:880-885)
            const double x = sfe::mt_canonical(mt + u, n_units, &cursor);
            const uint32_t col = cols[d.col_begin + r];
            if (col != 0xFFFFFFFFu) overlay[static_cast<size_t>(step) * n_cols + col] = d.probability > x ? 1 : 0;
        }
    idx[u] = cursor;
}

struct DevTables
{
    const CoreDev *cores;
    const struct SomaSegment *soma_segments; // one record per neuron-phase segment (local cores)
    uint32_t partitioned;             // 1: this engine simulates a core range; spikes arrive via the exchanged raster
    uint32_t inbox_lo, inbox_hi;      // local inbox word range [lo, hi)
    uint32_t inbox_words, pad_inbox;  // words of one inbox (the inbox is double-buffered by step parity)
    uint32_t n_soma_classes;
    const uint32_t *fanout_core_list; // cores that have axons-in
    const FanItem *fan_items;         // message-phase work items, grouped by core
    const uint32_t *fan_order;        // ticket -> item id, heaviest first
    uint32_t n_fan_items;
    const uint32_t *active_core_list; // cores with neurons or axons-in (ascending id)
    uint32_t n_active_cores;
    const sfe_soma_class *classes;
    const sfe_cost_class *costs;
    const uint32_t *neuron_class;
    const uint32_t *neuron_aux;
    const uint32_t *axon_out_begin;
    const uint32_t *axon_out_bit;     // padded inbox bit index of each axon-out
    // partitioned chip: raster bit of the source neuron of every local axon-in, indexed by
    // (inbox bit position - 32 * inbox_lo); 0xFFFFFFFF for padding bits
    const uint32_t *axon_src;
    const uint32_t *word_src;         // per local inbox word: how it derives from the raster (kWordGather / kWordEmpty / a raster word)
    const sfe_input_desc *inputs;
    const uint8_t *input_spikes;
    // Poisson inputs: host-drawn overlay (sfe_engine_set_input_overlay), row = step index - overlay_step0
    const uint8_t *overlay;
    long long overlay_step0, overlay_steps;
    uint32_t overlay_cols, pad_overlay;
    // TrueNorth threshold jitter: rand() values [step - rand_step0][jittered neuron] (sfe_engine_set_rand_overlay)
    const uint32_t *rand_overlay;
    long long rand_step0, rand_steps;
    uint32_t rand_cols, pad_rand;
    const sfe_axon_in *axons_in;
    const double *syn_w;
    const uint32_t *syn_meta;
    // Lossless 4-byte synapse records of cores whose certificate allows it (PACKED32, no delay
    // ring, <= 4096 neurons): bits 12..31 = weight * 2^shift as a 20-bit two's-complement integer
    // (exact by the certificate), bits 0..11 = post-synaptic neuron. Same padded indexing as
    // syn_w / syn_meta; a third of their bytes per synaptic event.
    const uint32_t *syn_q4;
    const uint32_t *probes;
    const TapsUnit *taps_units;       // "taps" dendrites
    const TapsSyn *taps_syn;
    const double *taps_values;
    const uint32_t *neuron_taps;      // per neuron: its taps unit or 0xFFFFFFFF
    uint32_t n_taps_units, pad_taps;
    const uint32_t *u_probes;         // LIF neurons whose input current u is traced (log_u)
    const sfe_noise_desc *noise;      // LIF file noise: per neuron of a unit with a stream
    const double *noise_values;
    uint32_t n_u_probes, pad_probes;
    uint32_t n_cores, n_probes, n_neurons, n_cost_classes, n_fanout_cores;
    uint32_t n_soma_segments, pad_segments;
    double sync_delay;
};

// Multi-GPU exchange over peer memory (NVLink), flag-in-data: a raster word travels as an 8-byte pair (word, epoch + 1)
// that the neuron-phase warp which produced it stores straight into every rank's double-buffered raster. There is
// no fence, no separate arrival flag and no publishing step: a reader polls the pair it needs until its epoch matches
// (an aligned 8-byte store arrives whole), so the exchange costs one NVLink crossing after the segment that fired.
constexpr int kMaxPeers = 8;
struct Exchange
{
    uint2 *ll[kMaxPeers];         // [2][fired_words] (word, epoch + 1) pairs of every rank (own entry: local memory)
    // Flow control of the two-deep pair buffers: done[q][r] (in rank q's memory) = timesteps rank r has completed.
    // Words of step t go into the buffer step t - 2 used, so a rank stores them only once every rank has completed
    // step t - 2 (ranks are tied together only along their data dependencies: without this a rank that needs
    // nothing from a slow peer would run ahead and overwrite words the peer has not read yet).
    uint32_t *done[kMaxPeers];
    uint32_t n_peers;             // 0: exchange done by the host side (NCCL / device copies) into a plain raster
    uint32_t rank;
    uint32_t fired_words;
    uint32_t slice_words;         // words per rank slice (multiple of 4)
    uint32_t *error;              // sticky: a word did not arrive in time
};

// What changes from launch to launch (host-side counters, passed by value next to the tables and the state; a batched
// launch derives a chip's copy from the one its chips - in lock step - share).
struct StepVars
{
    unsigned long long step_seq; // steps enqueued on the engine so far: parity of the double-buffered per-step arrays,
                                 // tag of the raster exchange
    long long steps_done;        // timesteps simulated before this step (T = steps_done + 1)
    long long fold_step;         // 0-based index of the step whose record this launch appends (slot of the device log)
    uint32_t fold_parity;        // parity of the step the stand-alone / piggy-backed fold works on
    uint32_t fuse_next;          // fused step kernel: this launch also runs the neuron phase of the next step
    uint32_t *work;              // ticket counter of this step's message phase (a slot of the engine's pool, see prepare_tickets)
    // CTAs of THIS chip in the running launch: neuron-phase kernel = segments + fold CTAs, message-phase kernel =
    // persistent CTAs drawing work items
    uint32_t soma_grid, fanout_grid;
};

struct StepPartial;
struct DevState
{
    Exchange x;
    double *v, *u, *bias;
    int32_t *refractory;
    uint8_t *status;
    uint32_t *fired_bits;  // padded per core (write base of the neuron phase)
    const uint32_t *fired_global; // raster of the whole chip (after the exchange when partitioned)
    uint32_t *inbox;       // padded per core
    uint32_t *din32;       // PACKED32: packed | DUAL32: sum
    uint32_t *dcnt32;      // DUAL32: count | ORDERED: has flag
    double *din64;         // ORDERED: value
    double *tap_v, *tap_next; // "taps": voltages of every line, scratch of the same shape
    long long *tap_steps;     // [n_taps_units] timesteps the line has been advanced to
    double *tap_buf;          // [n_neurons] tap 0 of the neuron's line after the last event that targeted the neuron
    uint32_t *tap_has;        // [n_neurons] the soma has an input waiting
    double *hh;            // [5][n_hh]: V, m, n, h, I
    uint32_t n_hh;
    double *nf_u2, *nf_uint; // "neurofem" neurons: u2 and the integrated error (potential = v, u1 = u)
    // out-of-tree soma device models (sfe_device_model.h), indexed by chip-wide instance (= neuron_aux of the neuron):
    // what the model's own kernel left for this step, and where get_potential() of the instance lives (or null)
    const uint8_t *plug_status;
    const double *const *plug_potential;
    StatsN *stats_n;
    StatsM *stats_m;
    sfe_step_record *log;  // ring of step records
    uint32_t log_cap;
    double *probe_out;     // [n_probes] potentials of the current step
    long long *step;       // [0] = timesteps simulated so far (T-1 during step T), [1] = log cursor
    uint32_t *final_ticket;
    // fused step kernel: a core is complete when all its work items of the step are done; the CTA that completes
    // it folds its statistics and runs its neuron phase of the NEXT step; the CTA that completes the last core
    // publishes the next step's raster (ready flag / peer exchange) and folds the chip
    uint32_t *core_done;      // [n_cores] work items of the core finished in this step
    uint32_t *cores_finished; // cores completed in this step
    uint32_t *ready;          // [0] = e + 1 once the neuron phase of step epoch e is complete on this GPU (unpartitioned chip)
    unsigned long long *timeline; // diagnostic (SFE_TIMELINE=1): [64 steps][grid][16] %globaltimer stamps of every CTA
    struct StepPartial *core_partials; // [2][n_active_cores], double-buffered by step parity
    struct StepPartial *partials;
};

// ---------------------------------------------------------------------------
// Small device helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Fire-and-forget OR into global memory. Written in PTX: when a kernel also contains atomics whose
// result is used, the compiler may emit the returning form (ATOMG) for plain atomicOr calls too,
// and a neuron-phase thread would then wait for every inbox bit it raises.
__device__ __forceinline__ void red_or(uint32_t *p, const uint32_t v)
{
    asm volatile("red.global.or.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_add(uint32_t *p, const uint32_t v)
{
    asm volatile("red.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed_sys(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Programmatic dependent launch: every kernel of the step lets its successor be scheduled right
// away (its CTAs fill SM slots as they free up and run their table-only prologue) and waits for
// its predecessor's completion + memory flush before touching anything a kernel writes.
__device__ __forceinline__ void griddep_launch_dependents()
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void griddep_wait()
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long v;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
    return v;
}

// One raster word of the current step from this rank's pair buffer: polls until the producer's store has landed
// (bounded: a missing peer sets the sticky error flag instead of hanging the GPU).
__device__ __forceinline__ uint32_t ll_word(const uint2 *pairs, const uint32_t index, const uint32_t want, uint32_t *error)
{
    const uint2 *p = pairs + index;
    uint32_t word, flag;
    asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(word), "=r"(flag) : "l"(p) : "memory");
    if (flag == want) return word;
    // time limit on the SM's own cycle counter (%globaltimer is a chip-wide register: polling it is slow); ~2 s
    const long long t0 = clock64();
    for (;;)
    {
        __nanosleep(20);
        asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(word), "=r"(flag) : "l"(p) : "memory");
        if (flag == want) return word;
        if (ld_relaxed_sys(error) != 0u) return 0u; // somebody already gave up: the run is invalid, finish it quickly
        if (clock64() - t0 > 4000000000ll)
        {
            // error[0] = 1, error[1..3] = which word, the epoch it was expected to carry, what it carried (diagnostics)
            if (atomicCAS(error, 0u, 1u) == 0u)
            {
                error[1] = index;
                error[2] = want;
                error[3] = flag;
            }
            return 0u;
        }
    }
}
__device__ __forceinline__ void ll_store(uint2 *p, const uint32_t word, const uint32_t flag)
{
    asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(word), "r"(flag) : "memory");
}

// Unpartitioned chip, fused step kernel: the launch that ran the neuron phase of step epoch e raises ready = e + 1
// (release); the launch that processes step e's messages waits for it (acquire). Bounded like ll_word.
__device__ __forceinline__ void ready_publish(uint32_t *ready, const unsigned long long epoch)
{
    __threadfence();
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(ready), "r"(static_cast<uint32_t>(epoch) + 1u) : "memory");
}
__device__ __forceinline__ void ready_wait(const DevState &s, const unsigned long long epoch);

__device__ __forceinline__ double warp_sum(double x)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
__device__ __forceinline__ unsigned long long warp_sum(unsigned long long x)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
__device__ __forceinline__ uint32_t warp_sum(uint32_t x)
{
    return __reduce_add_sync(0xffffffffu, x);
}
__device__ __forceinline__ double warp_max(double x)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, o));
    return x;
}

// Block-wide sum, result valid in thread 0. `scratch` holds >= 32 elements.
template <typename T> __device__ __forceinline__ T block_sum(T x, T *scratch)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
    x = warp_sum(x);
    __syncthreads();
    if (lane == 0) scratch[warp] = x;
    __syncthreads();
    if (warp == 0)
    {
        x = lane < nwarps ? scratch[lane] : T(0);
        x = warp_sum(x);
    }
    return x;
}

// ---------------------------------------------------------------------------
// Soma models (device functors). Same operation order as the reference; no FMA.
// ---------------------------------------------------------------------------
// LoihiLifModel::update  src/models.cpp:497-567
template <bool kNoise = false>
__device__ __forceinline__ int lif_update(const sfe_soma_class &c, double &v, double &u, int &refractory,
        const double bias, const bool has_in, const double in, const long long steps_done, const double noise = 0.0)
{
    int state = SFE_STATUS_IDLE;
    if ((fabs(v) > 0.0) || has_in || (fabs(bias) > 0.0) || (c.flags & SFE_SOMA_FORCE_UPDATE)) state = SFE_STATUS_UPDATED;
    if (steps_done > 0)
    {
        u = u * c.input_decay;
        v = v * c.leak;
    }
    v = static_cast<double>(__double2int_rz(v * 64.0)) / 64.0; // loihi_quantize (C int cast)
    if constexpr (kNoise) v = v + noise;                        // src/models.cpp:535-539
    if (refractory <= 0)
    {
        v = v + bias;
        u = u + (has_in ? in : 0.0);
        v = v + u;
        if (v > c.threshold)
        {
            if (c.reset_mode == SFE_RESET_HARD) v = c.reset;
            else if (c.reset_mode == SFE_RESET_SOFT) v = v - c.threshold;
            refractory = c.refractory_delay;
            state = SFE_STATUS_FIRED;
        }
        if (v < c.reverse_threshold)
        {
            if (c.reverse_reset_mode == SFE_RESET_SOFT) v = v - c.reverse_threshold;
            else if (c.reverse_reset_mode == SFE_RESET_HARD) v = c.reverse_reset;
            else if (c.reverse_reset_mode == SFE_RESET_SATURATE) v = c.reverse_threshold;
        }
    }
    refractory = max(0, refractory - 1);
    return state;
}

// TrueNorthModel::update  src/models.cpp:724-830. rnd = this update's `rand() & random_mask` (0 without jitter): it
// shifts the value the thresholds are compared with, not the potential (truenorth_threshold_and_reset, :745-759)
__device__ __forceinline__ int truenorth_update(
        const sfe_soma_class &c, double &v, const double bias, const bool has_in, const double in, const uint32_t rnd = 0u)
{
    int state = SFE_STATUS_IDLE;
    if ((fabs(v) > 0.0) || has_in || (fabs(bias) > 0.0) || (c.flags & SFE_SOMA_FORCE_UPDATE)) state = SFE_STATUS_UPDATED;
    if (c.flags & SFE_SOMA_LEAK_TOWARDS_ZERO)
    {
        if (v > 0.0) v = v - c.leak;
        else if (v < 0.0) v = v + c.leak;
    }
    else v = v + c.leak;
    v = v + bias;
    if (has_in) v = v + in;
    const double compared = rnd != 0u ? v + static_cast<double>(rnd) : v;
    if (compared >= c.threshold)
    {
        if (c.reset_mode == SFE_RESET_HARD) v = c.reset;
        else if (c.reset_mode == SFE_RESET_SOFT) v = v - c.threshold;
        else if (c.reset_mode == SFE_RESET_SATURATE) v = c.threshold;
        state = SFE_STATUS_FIRED;
    }
    else if (compared <= c.reverse_threshold)
    {
        if (c.reverse_reset_mode == SFE_RESET_HARD) v = c.reverse_reset;
        else if (c.reverse_reset_mode == SFE_RESET_SOFT) v = v + c.reverse_threshold;
        else if (c.reverse_reset_mode == SFE_RESET_SATURATE) v = c.reverse_threshold;
    }
    return state;
}

// InputModel::update  src/models.cpp:863-903 (per-unit cursor shared by share_count neurons)
__device__ __forceinline__ int input_update(const DevTables &t, const sfe_input_desc &d, const long long step_idx,
        const long long timestep)
{
    bool send = false;
    const unsigned long long cursor = static_cast<unsigned long long>(step_idx) * d.share_count + d.share_rank;
    if (cursor < d.spikes_len) send = t.input_spikes[d.spikes_off + cursor] != 0;
    // poisson > U(0,1): the draw was made on the host with the reference's generator (poisson.cpp)
    if (d.poisson > 0.0) // (engine creation checked poisson_col; check_overlay() that the overlay covers this step)
    {
        const long long row = step_idx - t.overlay_step0;
        if (row >= 0 && row < t.overlay_steps && t.overlay[static_cast<size_t>(row) * t.overlay_cols + d.poisson_col] != 0)
            send = true;
    }
    if ((d.rate > 0.0) && ((timestep % static_cast<long long>(1.0 / d.rate)) == 0)) send = true;
    return send ? SFE_STATUS_FIRED : SFE_STATUS_IDLE;
}

// HodgkinHuxley::update  plugins/hodgkin_huxley.cpp:116-170 — the reference plugin
// compiled as a device functor. State layout hh[5][n_hh] = V, m, n, h, I.
__device__ __forceinline__ int hh_update(double *hh, const uint32_t n_hh, const uint32_t k)
{
    const double C_m = 10.0, g_Na = 1200.0, g_K = 360.0, g_L = 3.0, V_Na = 50.0, V_K = -77.0, V_L = 54.387, dt = 0.1;
    double V = hh[k], m = hh[n_hh + k], n = hh[2 * n_hh + k], h = hh[3 * n_hh + k];
    const double I = hh[4 * n_hh + k];
    const double alpha_n = (0.01 * (V + 55)) / (1 - exp(-0.1 * (V + 55)));
    const double alpha_m = (0.1 * (V + 40)) / (1 - exp(-0.1 * (V + 40)));
    const double alpha_h = 0.07 * exp(-0.05 * (V + 65));
    const double beta_n = 0.125 * exp(-0.01125 * (V + 55));
    const double beta_m = 4 * exp(-0.05556 * (V + 65));
    const double beta_h = 1 / (1 + exp(-0.1 * (V + 35)));
    const double tau_n = 1 / (alpha_n + beta_n);
    const double tau_m = 1 / (alpha_m + beta_m);
    const double tau_h = 1 / (alpha_h + beta_h);
    const double pm = alpha_m / (alpha_m + beta_m);
    const double pn = alpha_n / (alpha_n + beta_n);
    const double ph = alpha_h / (alpha_h + beta_h);
    const double n4 = pow(n, 4.0);
    const double m3 = pow(m, 3.0);
    const double denominator = g_L + g_K * n4 + g_Na * (m3 * h);
    const double tau_V = C_m / denominator;
    const double Vinf = ((g_L) * V_L + g_K * n4 * V_K + g_Na * m3 * h * V_Na + I) / denominator;
    const double prev_V = V;
    V = Vinf + (V - Vinf) * exp(-1 * dt / tau_V);
    m = pm + (m - pm) * exp(-1 * dt / tau_m);
    n = pn + (n - pn) * exp(-1 * dt / tau_n);
    h = ph + (h - ph) * exp(-1 * dt / tau_h);
    hh[k] = V;
    hh[n_hh + k] = m;
    hh[2 * n_hh + k] = n;
    hh[3 * n_hh + k] = h;
    return ((prev_V < 25) && (V > 25)) ? SFE_STATUS_FIRED : SFE_STATUS_UPDATED;
}

// NeuroFEMModel::update + process_fem  plugins/neurofem.cpp:192-317 — the reference plugin compiled as a device functor,
// for the call of the neuron phase (the first update of a timestep): the two accumulators filled by the previous
// message phase are consumed, then the compartment is advanced. sigma_v = 0 is enforced at load: the noise term
// sigma_v * N(0,1) is +-0 and leaves the potential unchanged. lambda_v rides in c.leak, lambda_d in c.input_decay.
__device__ __forceinline__ int neurofem_update(const sfe_soma_class &c, double &v, double &u1, double &u2, double &u_int,
        const double bias, const double acc1, const double acc2)
{
    const double lambda_d = c.input_decay, lambda_v = c.leak, dt = c.nf_dt;
    u1 = u1 - lambda_d * dt * u1;
    u2 = u2 - lambda_d * dt * u2;
    u1 = u1 + acc1;
    u2 = u2 + lambda_d * acc2;
    const double u_err = u1 + bias;
    u_int = u_int + dt * u_err;
    v = v - (lambda_v * dt * v);
    v = v + (dt * c.nf_kp * u_err) + (dt * c.nf_ki * u_int) + (dt * u2) + 0.0 - acc2;
    if (v > c.threshold)
    {
        v = c.reset;
        return SFE_STATUS_FIRED;
    }
    return SFE_STATUS_UPDATED;
}

// ---------------------------------------------------------------------------
// K5: energy, counters, simple timing model. One thread per active core; the last
// CTA to finish (ticket) folds the per-CTA partials in a fixed order and appends the
// step record, so every floating-point sum is formed in the same order on every run.
// ---------------------------------------------------------------------------
constexpr int kFinalThreads = 256;

struct StepPartial
{
    unsigned long long fired, updated, packets, hops, events;
    double syn_e, den_e, soma_e, net_e, max_gen, max_proc;
    double dup_e;
};

__device__ __forceinline__ StepPartial load_partial(const StepPartial *p) // L2 (coherent) loads
{
    StepPartial r;
    r.fired = __ldcg(&p->fired);
    r.updated = __ldcg(&p->updated);
    r.packets = __ldcg(&p->packets);
    r.hops = __ldcg(&p->hops);
    r.events = __ldcg(&p->events);
    r.syn_e = __ldcg(&p->syn_e);
    r.den_e = __ldcg(&p->den_e);
    r.soma_e = __ldcg(&p->soma_e);
    r.net_e = __ldcg(&p->net_e);
    r.max_gen = __ldcg(&p->max_gen);
    r.max_proc = __ldcg(&p->max_proc);
    r.dup_e = __ldcg(&p->dup_e);
    return r;
}

__device__ __forceinline__ void fold_partial(StepPartial &b, const StepPartial &x)
{
    b.fired += x.fired;
    b.updated += x.updated;
    b.packets += x.packets;
    b.hops += x.hops;
    b.events += x.events;
    b.syn_e += x.syn_e;
    b.den_e += x.den_e;
    b.soma_e += x.soma_e;
    b.net_e += x.net_e;
    b.dup_e += x.dup_e;
    b.max_gen = fmax(b.max_gen, x.max_gen);
    b.max_proc = fmax(b.max_proc, x.max_proc);
}

__device__ __forceinline__ StepPartial warp_fold(StepPartial p) // xor butterfly: a fixed order
{
    p.fired = warp_sum(p.fired);
    p.updated = warp_sum(p.updated);
    p.packets = warp_sum(p.packets);
    p.hops = warp_sum(p.hops);
    p.events = warp_sum(p.events);
    p.syn_e = warp_sum(p.syn_e);
    p.den_e = warp_sum(p.den_e);
    p.soma_e = warp_sum(p.soma_e);
    p.net_e = warp_sum(p.net_e);
    p.dup_e = warp_sum(p.dup_e);
    p.max_gen = warp_max(p.max_gen);
    p.max_proc = warp_max(p.max_proc);
    return p;
}

// Step statistics of one core, computed by a full warp: the lanes fetch the core's segment /
// work-item statistics in parallel (one round trip instead of seg_count + item_count dependent
// ones) and fold them with a butterfly; every sum is formed in the same order on every run.
// The result is valid in every lane. L2 loads: the statistics may come from other CTAs of the
// running kernel (fused finalize).
__device__ __forceinline__ StepPartial fold_core(
        const DevTables &t, const DevState &s, const uint32_t ci, const int lane, const uint32_t parity)
{
    StepPartial p = {0ull, 0ull, 0ull, 0ull, 0ull, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    const CoreDev &core = t.cores[ci];
    StatsN n = {0u, 0u, 0u, 0u, 0.0, 0.0, 0.0, 0.0};
    for (uint32_t g = lane; g < core.seg_count; g += 32)
    {
        const StatsN *x = s.stats_n + static_cast<size_t>(parity) * t.n_soma_segments + core.seg_begin + g;
        n.updated += __ldcg(&x->updated);
        n.fired += __ldcg(&x->fired);
        n.packets += __ldcg(&x->packets);
        n.soma_e += __ldcg(&x->soma_e);
        n.dend_e += __ldcg(&x->dend_e);
        n.gen_sum += __ldcg(&x->gen_sum);
        n.dup_e += __ldcg(&x->dup_e);
    }
    StatsM m = {0u, 0u, 0ull, 0ull, 0ull, 0ull, 0ull, 0.0, 0.0, 0.0, 0.0};
    for (uint32_t g = lane; g < core.item_count; g += 32)
    {
        const StatsM *x = s.stats_m + core.item_begin + g;
        m.msgs += __ldcg(&x->msgs);
        m.events += __ldcg(&x->events);
        m.hop_e += __ldcg(&x->hop_e);
        m.hop_w += __ldcg(&x->hop_w);
        m.hop_n += __ldcg(&x->hop_n);
        m.hop_s += __ldcg(&x->hop_s);
        m.syn_e += __ldcg(&x->syn_e);
        m.den_e += __ldcg(&x->den_e);
        m.proc += __ldcg(&x->proc);
        m.dup_e += __ldcg(&x->dup_e);
    }
    n.updated = warp_sum(n.updated);
    n.fired = warp_sum(n.fired);
    n.packets = warp_sum(n.packets);
    n.soma_e = warp_sum(n.soma_e);
    n.dend_e = warp_sum(n.dend_e);
    n.gen_sum = warp_sum(n.gen_sum);
    n.dup_e = warp_sum(n.dup_e);
    m.msgs = warp_sum(m.msgs);
    m.events = warp_sum(m.events);
    m.hop_e = warp_sum(m.hop_e);
    m.hop_w = warp_sum(m.hop_w);
    m.hop_n = warp_sum(m.hop_n);
    m.hop_s = warp_sum(m.hop_s);
    m.syn_e = warp_sum(m.syn_e);
    m.den_e = warp_sum(m.den_e);
    m.proc = warp_sum(m.proc);
    m.dup_e = warp_sum(m.dup_e);
    n.gen_sum += static_cast<double>(n.packets) * core.lat_axon_out;
    p.fired = n.fired;
    p.updated = n.updated;
    p.packets = n.packets;
    p.hops = m.hop_e + m.hop_w + m.hop_n + m.hop_s;
    p.events = m.events;
    p.syn_e = m.syn_e;
    p.den_e = n.dend_e + m.den_e;
    p.soma_e = n.soma_e + m.dup_e; // events through a combined dendrite + soma unit cost it a neuron access each
    p.dup_e = n.dup_e + m.dup_e;
    // sim_calculate_tile_energy / sim_calculate_core_energy  src/chip.cpp:1189-1261
    double hop = static_cast<double>(m.hop_e) * core.e_east;
    hop += static_cast<double>(m.hop_w) * core.e_west;
    hop += static_cast<double>(m.hop_s) * core.e_south;
    hop += static_cast<double>(m.hop_n) * core.e_north;
    p.net_e = hop + static_cast<double>(m.msgs) * core.e_axon_in + static_cast<double>(n.packets) * core.e_axon_out;
    p.max_gen = n.gen_sum;
    p.max_proc = m.proc;
    return p;
}

// Appends the step record (one thread). schedule_messages_timestep_simple  src/schedule.cpp:61-102
__device__ __forceinline__ void append_step_record(const DevTables &t, const DevState &s, const StepVars &sv, const StepPartial &b)
{
    sfe_step_record r;
    r.neurons_fired = static_cast<long long>(b.fired);
    r.neurons_updated = static_cast<long long>(b.updated);
    r.packets_sent = static_cast<long long>(b.packets);
    r.total_hops = static_cast<long long>(b.hops);
    r.spike_count = static_cast<long long>(b.events);
    r.synapse_energy = b.syn_e;
    r.dendrite_energy = b.den_e;
    r.soma_energy = b.soma_e;
    r.network_energy = b.net_e;
    r.total_energy = b.net_e + b.syn_e + b.den_e + b.soma_e - b.dup_e; // a unit's energy once (src/chip.cpp:1224-1261)
    r.sim_time = fmax(b.max_proc, b.max_gen) + t.sync_delay;
    // the record's slot follows from the step number the host passed with the launch (the folds of consecutive
    // steps may overlap in the fused step kernel: nothing here reads what another fold writes)
    s.log[sv.fold_step % s.log_cap] = r;
    s.step[1] = sv.fold_step + 1;
    s.step[0] = sv.fold_step + 1;
}

// Chip-wide fold by one warp over the per-core partials of the step (fused step kernel).
__device__ __forceinline__ void fold_chip(const DevTables &t, const DevState &s, const StepVars &sv, const int lane)
{
    const StepPartial *partials = s.core_partials + static_cast<size_t>(sv.step_seq & 1ull) * t.n_active_cores;
    StepPartial b = {0ull, 0ull, 0ull, 0ull, 0ull, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (uint32_t k = lane; k < t.n_active_cores; k += 32) fold_partial(b, load_partial(&partials[k]));
    b = warp_fold(b);
    if (lane == 0) append_step_record(t, s, sv, b);
}

// The stand-alone form of the fold: one WARP per active core, `nblocks` CTAs numbered `block`.
// Runs as its own kernel (finalize_kernel) or as extra CTAs of the NEXT step's neuron-phase kernel
// (the step's statistics are complete by then and nothing on the critical path waits for the fold).
// `scratch`: (kFinalThreads / 32) StepPartial + one uint32 of shared memory, supplied by the caller
// (the neuron-phase kernel lends its class cache: an extra static array would push three of its
// CTAs past the 16 KB shared-memory carve-out and cost it a third of its occupancy).
constexpr size_t kFinalScratchBytes = (kFinalThreads / 32) * sizeof(StepPartial) + 16;
__device__ __forceinline__ void finalize_body(const DevTables &t, const DevState &s, const StepVars &sv, const uint32_t block,
        const uint32_t nblocks, unsigned char *scratch)
{
    StepPartial *warp_part = reinterpret_cast<StepPartial *>(scratch);
    uint32_t &ticket_s = *reinterpret_cast<uint32_t *>(scratch + (kFinalThreads / 32) * sizeof(StepPartial));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    StepPartial p = {0ull, 0ull, 0ull, 0ull, 0ull, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    const uint32_t a = block * (kFinalThreads / 32) + warp;
    if (a < t.n_active_cores) p = fold_core(t, s, t.active_core_list[a], lane, sv.fold_parity);
    if (lane == 0) warp_part[warp] = p;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        StepPartial b = warp_part[0];
        for (int w = 1; w < kFinalThreads / 32; ++w) fold_partial(b, warp_part[w]);
        s.partials[block] = b;
        __threadfence();
        ticket_s = atomicAdd(s.final_ticket, 1u);
    }
    __syncthreads();
    if (ticket_s != nblocks - 1) return;
    // ---- last CTA: fold the per-CTA partials and append the step record ------------------
    __threadfence();
    StepPartial b = {0ull, 0ull, 0ull, 0ull, 0ull, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (uint32_t k = threadIdx.x; k < nblocks; k += kFinalThreads) fold_partial(b, load_partial(&s.partials[k]));
    b = warp_fold(b);
    __syncthreads(); // warp_part is reused
    if (lane == 0) warp_part[warp] = b;
    __syncthreads();
    if (threadIdx.x != 0) return;
    b = warp_part[0];
    for (int w = 1; w < kFinalThreads / 32; ++w) fold_partial(b, warp_part[w]);
    append_step_record(t, s, sv, b);
    *s.final_ticket = 0u;
}

// ---------------------------------------------------------------------------
// K1: neuron phase. One CTA per segment of kSomaThreads consecutive neurons of a core
// (one neuron per thread: every load of the step is issued at once).
// kExotic = the chip maps input / Hodgkin-Huxley somas (kept out of the LIF /
// TrueNorth instantiation so the common case stays at low register pressure).
// ---------------------------------------------------------------------------
// Diagnostic build (make variant NAME=tl EXTRA_NVFLAGS=-DSFE_TIMELINE_ALL=1, run with SFE_TIMELINE=1): the two-kernel
// step leaves %globaltimer stamps too (tools/timeline_partitioned.py). Costs registers: never on in the shipped build.
#ifndef SFE_TIMELINE_ALL
#define SFE_TIMELINE_ALL 0
#endif
constexpr bool kTimelineAll = SFE_TIMELINE_ALL != 0;
constexpr size_t kTimelineWords = 64ull * 1024ull * 16ull; // [64 steps][<= 1024 CTAs][16 stamps]; the neuron-phase kernel uses a second block
constexpr int kSomaThreads = 256;
#ifndef SFE_SOMA_PER_THREAD
#define SFE_SOMA_PER_THREAD 2
#endif
constexpr int kSomaPerThread = SFE_SOMA_PER_THREAD; // neurons per thread: a segment is kSomaThreads * kSomaPerThread neurons

constexpr int kClassCache = 24; // soma parameter classes cached in shared memory (160 B each)

// everything the neuron phase needs about its core, in one record (one dependent load
// less than going through the core table)
struct SomaSegment
{
    uint32_t k0;               // first neuron of the segment within the core
    uint32_t neuron_begin, neuron_count;
    uint32_t fired_word_begin;
    uint32_t dend_base, ring, acc_mode, fixed_slots;
    double inv_scale;
};

struct SomaScratch // per-CTA reduction scratch of the neuron phase
{
    double part_d[kSomaThreads / 32][4];
    uint32_t part_u[kSomaThreads / 32][3];
};

// The neuron phase of one segment (kSomaThreads * kSomaPerThread consecutive neurons of a core) by a whole CTA of
// kSomaThreads threads: the body of soma_kernel, also run by the fused step kernel for the cores whose message
// phase has just finished. `steps_done` = timesteps simulated before this step, `parity` = parity of the step (the
// per-segment statistics and the inbox are double-buffered by it). class_cache: shared copy of the class table or null.
template <bool kExotic>
__device__ __forceinline__ void soma_segment(const DevTables &t, const DevState &s, const StepVars &sv, const uint32_t seg_idx,
        const sfe_soma_class *class_cache, SomaScratch &scr, const long long steps_done, const uint32_t parity)
{
    uint32_t *const inbox = s.inbox + static_cast<size_t>(parity) * t.inbox_words;
    const SomaSegment core = t.soma_segments[seg_idx];
    const int lane = threadIdx.x & 31;
    const bool classes_cached = class_cache != nullptr;
    const long long T = steps_done + 1;

    uint32_t n_updated = 0, n_fired = 0, n_packets = 0;
    double soma_e = 0.0, dend_e = 0.0, lat_sum = 0.0, dup_e = 0.0;
    // the dendrite slot read this step: the rotating slot of a delay line, or slot 0 of fixed slots ("neurofem" cores)
    const uint32_t slot = (core.ring > 1 && core.fixed_slots == 0u) ? static_cast<uint32_t>(T % core.ring) : 0u;

    // Every global load of the step is issued before the first use, speculatively: which of them
    // a neuron needs depends on its class, but waiting for the class first would put three
    // dependent memory round trips where one suffices (the unused ones cost little: the neuron
    // phase is latency-bound, not bandwidth-bound). A thread owns kSomaPerThread neurons
    // (kSomaThreads apart) and has the loads of all of them in flight at once.
    uint32_t cid_r[kSomaPerThread], a0_r[kSomaPerThread], a1_r[kSomaPerThread], sum_r[kSomaPerThread], cnt_r[kSomaPerThread];
    double v_r[kSomaPerThread], u_r[kSomaPerThread], bias_r[kSomaPerThread];
    int refr_r[kSomaPerThread];
#pragma unroll
    for (int r = 0; r < kSomaPerThread; ++r)
    {
        const uint32_t k = core.k0 + r * kSomaThreads + threadIdx.x;
        cid_r[r] = a0_r[r] = a1_r[r] = sum_r[r] = cnt_r[r] = 0u;
        v_r[r] = u_r[r] = bias_r[r] = 0.0;
        refr_r[r] = 0;
        if (k < core.neuron_count)
        {
            const uint32_t i = core.neuron_begin + k;
            const uint32_t d = core.dend_base + slot * core.neuron_count + k;
            cid_r[r] = __ldg(t.neuron_class + i);
            a0_r[r] = __ldg(t.axon_out_begin + i);
            a1_r[r] = __ldg(t.axon_out_begin + i + 1);
            v_r[r] = s.v[i];
            u_r[r] = s.u[i];
            bias_r[r] = s.bias[i];
            refr_r[r] = s.refractory[i];
            sum_r[r] = s.din32[d];
            if (core.acc_mode != SFE_ACC_PACKED32 && core.acc_mode != SFE_ACC_PACKED17) cnt_r[r] = s.dcnt32[d];
        }
    }
    __syncthreads(); // class cache filled
#pragma unroll
    for (int r = 0; r < kSomaPerThread; ++r)
    {
        const uint32_t k = core.k0 + r * kSomaThreads + threadIdx.x;
        const bool valid = k < core.neuron_count;
        int st = SFE_STATUS_IDLE;
        const uint32_t cid = cid_r[r], a0 = a0_r[r], a1 = a1_r[r], raw_sum = sum_r[r], raw_cnt = cnt_r[r];
        const double v0 = v_r[r], u0 = u_r[r], bias0 = bias_r[r];
        const int refr0 = refr_r[r];
        const uint32_t d = core.dend_base + slot * core.neuron_count + k;
        if (valid)
        {
            const uint32_t i = core.neuron_begin + k;
            const sfe_soma_class c = classes_cached ? class_cache[cid] : t.classes[cid];
            // ---- dendrite output for this step ------------------------------
            bool has_in = false;
            double in = 0.0;
            bool tap_line = false;
            if constexpr (kExotic) tap_line = c.dend_model == SFE_DEND_TAPS;
            bool neurofem = false;
            if constexpr (kExotic) neurofem = c.model == SFE_SOMA_NEUROFEM;
            double nf_acc1 = 0.0, nf_acc2 = 0.0;
            if (neurofem)
            {
                // combined dendrite + soma unit: two accumulators (slot 0: u1, slot 1: u2) filled by the previous
                // message phase, consumed here (next_*_dendritic_accumulator.value_or(0.0), plugins/neurofem.cpp:204-215)
                const uint32_t d1 = d + core.neuron_count;
                if (core.acc_mode == SFE_ACC_ORDERED)
                {
                    if (raw_cnt != 0u)
                    {
                        nf_acc1 = s.din64[d];
                        s.din64[d] = 0.0;
                        s.dcnt32[d] = 0u;
                    }
                    if (s.dcnt32[d1] != 0u)
                    {
                        nf_acc2 = s.din64[d1];
                        s.din64[d1] = 0.0;
                        s.dcnt32[d1] = 0u;
                    }
                }
                else
                {
                    const uint32_t raw1 = s.din32[d1];
                    int sum0, sum1;
                    if (core.acc_mode == SFE_ACC_PACKED17)
                    {
                        sum0 = static_cast<int>(raw_sum - (((raw_sum + 0x10000u) >> 17) << 17));
                        sum1 = static_cast<int>(raw1 - (((raw1 + 0x10000u) >> 17) << 17));
                    }
                    else if (core.acc_mode == SFE_ACC_PACKED32)
                    {
                        sum0 = (static_cast<int>(raw_sum << 12)) >> 12;
                        sum1 = (static_cast<int>(raw1 << 12)) >> 12;
                    }
                    else
                    {
                        sum0 = static_cast<int>(raw_sum);
                        sum1 = static_cast<int>(raw1);
                        if (raw_cnt != 0u) s.dcnt32[d] = 0u;
                        if (s.dcnt32[d1] != 0u) s.dcnt32[d1] = 0u;
                    }
                    nf_acc1 = static_cast<double>(sum0) * core.inv_scale;
                    nf_acc2 = static_cast<double>(sum1) * core.inv_scale;
                    if (raw_sum != 0u) s.din32[d] = 0u;
                    if (raw1 != 0u) s.din32[d1] = 0u;
                }
            }
            else if (tap_line)
            {
                // the line's tap 0 after the last event of the previous step (taps_kernel); the accumulator cell
                // the message phase filled for this neuron is ignored
                if (s.tap_has[i] != 0u)
                {
                    has_in = true;
                    in = s.tap_buf[i];
                    s.tap_has[i] = 0u;
                }
            }
            else if (c.dend_in_neuron && c.dend_model == SFE_DEND_ACCUMULATOR)
            {
                // buffer inside a plain accumulator: the charge is zeroed before it is
                // read (src/models.cpp:78-82)
                has_in = true;
            }
            else if (core.acc_mode == SFE_ACC_PACKED17)
            {
                if (raw_sum != 0u)
                {
                    // cell = sum + count * 2^17 with |sum| < 2^16: the count is the cell rounded to 2^17
                    const uint32_t count = (raw_sum + 0x10000u) >> 17;
                    has_in = true; // a non-zero cell has count >= 1 (count * 2^17 - |sum| > 0)
                    in = static_cast<double>(static_cast<int>(raw_sum - (count << 17))) * core.inv_scale;
                    s.din32[d] = 0u; // consumed (the message phase accumulates into zeroed slots)
                }
            }
            else if (core.acc_mode == SFE_ACC_PACKED32)
            {
                if (raw_sum != 0u)
                {
                    const int sum = (static_cast<int>(raw_sum << 12)) >> 12; // sign-extend 20 bits
                    has_in = true;                                           // count = (raw - sum) >> 20 > 0
                    in = static_cast<double>(sum) * core.inv_scale;
                    s.din32[d] = 0u; // consumed (the message phase accumulates into zeroed slots)
                }
            }
            else if (core.acc_mode == SFE_ACC_DUAL32)
            {
                if (raw_cnt != 0u)
                {
                    has_in = true;
                    in = static_cast<double>(static_cast<int>(raw_sum)) * core.inv_scale;
                    s.din32[d] = 0u;
                    s.dcnt32[d] = 0u;
                }
            }
            else
            {
                if (raw_cnt != 0u)
                {
                    has_in = true;
                    in = s.din64[d];
                    // consumed: a delay-ring slot is reused ring steps later, and a core that accumulates in HBM
                    // (acc_global) adds onto these cells directly
                    s.din64[d] = 0.0;
                    s.dcnt32[d] = 0u;
                }
            }
            double lat = 0.0;
            if (c.dend_in_neuron)
            {
                dend_e += c.dend_energy_update;
                lat += c.dend_latency_update;
            }
            // ---- soma ---------------------------------------------------------
            const double bias = bias0;
            if (c.model == SFE_SOMA_LIF)
            {
                double v = v0, u = u0;
                int refr = refr0;
                bool noisy = false;
                if constexpr (kExotic) noisy = (c.flags & SFE_SOMA_NOISE) != 0u;
                if (noisy)
                {
                    // the unit's stream: one entry per update of any of its neurons, rewinding at the end
                    const sfe_noise_desc nd = t.noise[t.neuron_aux[i]];
                    const unsigned long long cursor =
                            static_cast<unsigned long long>(steps_done) * nd.share_count + nd.share_rank;
                    st = lif_update<true>(c, v, u, refr, bias, has_in, in, steps_done, t.noise_values[nd.off + cursor % nd.len]);
                }
                else st = lif_update(c, v, u, refr, bias, has_in, in, steps_done);
                s.v[i] = v;
                s.u[i] = u;
                s.refractory[i] = refr;
            }
            else if (c.model == SFE_SOMA_TRUENORTH)
            {
                double v = v0;
                uint32_t rnd = 0u;
                if constexpr (kExotic)
                    if (c.random_mask != 0u)
                    {
                        // (check_overlay made sure the overlay covers this step)
                        const long long row = steps_done - t.rand_step0;
                        if (row >= 0 && row < t.rand_steps)
                            rnd = __ldg(t.rand_overlay + static_cast<size_t>(row) * t.rand_cols + t.neuron_aux[i]) & c.random_mask;
                    }
                st = truenorth_update(c, v, bias, has_in, in, rnd);
                s.v[i] = v;
            }
            else if constexpr (kExotic)
            {
                if (c.model == SFE_SOMA_INPUT) st = input_update(t, t.inputs[t.neuron_aux[i]], steps_done, T);
                else if (c.model == SFE_SOMA_NEUROFEM)
                {
                    double v = v0, u1 = u0, u2 = s.nf_u2[i], u_int = s.nf_uint[i];
                    st = neurofem_update(c, v, u1, u2, u_int, bias, nf_acc1, nf_acc2);
                    s.v[i] = v;
                    s.u[i] = u1;
                    s.nf_u2[i] = u2;
                    s.nf_uint[i] = u_int;
                }
                else if (c.model == SFE_SOMA_DEVICE_MODEL) st = s.plug_status[t.neuron_aux[i]]; // written by the model's own kernel
                else st = hh_update(s.hh, s.n_hh, t.neuron_aux[i]);
            }
            s.status[i] = static_cast<uint8_t>(st);
            // ---- default soma costs  src/pipeline.hpp:631-714 -----------------
            double e = c.energy_access, l = c.latency_access;
            if (st >= SFE_STATUS_UPDATED)
            {
                e += c.energy_update;
                l += c.latency_update;
                ++n_updated;
            }
            if (st == SFE_STATUS_FIRED)
            {
                e += c.energy_spike_out;
                l += c.latency_spike_out;
                ++n_fired;
            }
            soma_e += e;
            if constexpr (kExotic)
                if ((c.flags & SFE_SOMA_IS_DENDRITE) != 0u)
                {
                    dend_e += e; // a combined unit's energy counts in the dendrite and the soma bucket, once in the total
                    dup_e += e;
                }
            lat_sum += lat + l;
        }
        // spike raster: one ballot per 32 neurons
        const uint32_t ballot = __ballot_sync(0xffffffffu, st == SFE_STATUS_FIRED);
        if (lane == 0 && k < ((core.neuron_count + 31u) & ~31u)) s.fired_bits[core.fired_word_begin + (k >> 5)] = ballot;
        // Partitioned chip, peer-memory exchange: lane q hands the word to rank q as a (word, epoch + 1) pair - the
        // whole exchange (ll_word on the reading side)
        if (static_cast<uint32_t>(lane) < s.x.n_peers && k - lane < ((core.neuron_count + 31u) & ~31u))
            ll_store(s.x.ll[lane] + (sv.step_seq & 1ull) * s.x.fired_words + core.fired_word_begin + ((k - lane) >> 5), ballot,
                    static_cast<uint32_t>(sv.step_seq) + 1u);
        // pipeline_process_axon_out  src/chip.cpp:802-834: one message per axon of a fired
        // neuron: raise the inbox bit of every target (loads batched eight at a time).
        if (st == SFE_STATUS_FIRED) n_packets += a1 - a0;
        if (st == SFE_STATUS_FIRED && t.partitioned == 0u)
        {
            for (uint32_t base = a0; base < a1; base += 8u)
            {
                uint32_t bit[8];
#pragma unroll
                for (int x = 0; x < 8; ++x)
                    if (base + x < a1) bit[x] = __ldg(t.axon_out_bit + base + x);
#pragma unroll
                for (int x = 0; x < 8; ++x)
                    if (base + x < a1) red_or(&inbox[bit[x] >> 5], 1u << (bit[x] & 31));
            }
        }
    }
    // ---- per-segment reductions: warp shuffles, one barrier -------------------------
    double(*part_d)[4] = scr.part_d;
    uint32_t(*part_u)[3] = scr.part_u;
    n_updated = warp_sum(n_updated);
    n_fired = warp_sum(n_fired);
    n_packets = warp_sum(n_packets);
    soma_e = warp_sum(soma_e);
    dend_e = warp_sum(dend_e);
    lat_sum = warp_sum(lat_sum);
    if constexpr (kExotic) dup_e = warp_sum(dup_e);
    const int warp = threadIdx.x >> 5;
    if (lane == 0)
    {
        part_u[warp][0] = n_updated;
        part_u[warp][1] = n_fired;
        part_u[warp][2] = n_packets;
        part_d[warp][0] = soma_e;
        part_d[warp][1] = dend_e;
        part_d[warp][2] = lat_sum;
        part_d[warp][3] = dup_e;
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        StatsN out = {0u, 0u, 0u, 0u, 0.0, 0.0, 0.0, 0.0};
        for (int w = 0; w < kSomaThreads / 32; ++w) // fixed order: deterministic sums
        {
            out.updated += part_u[w][0];
            out.fired += part_u[w][1];
            out.packets += part_u[w][2];
            out.soma_e += part_d[w][0];
            out.dend_e += part_d[w][1];
            out.gen_sum += part_d[w][2]; // finalize adds packets * latency_axon_out per core
            out.dup_e += part_d[w][3];
        }
        // gen_sum: sum of generation delays incl. the trailing placeholder message
        // (src/chip.cpp:640-652, 821-823; src/schedule.cpp:81)
        s.stats_n[static_cast<size_t>(parity) * t.n_soma_segments + seg_idx] = out; // double-buffered by step parity
    }
    __syncthreads(); // the scratch is free again (a CTA may run several segments in a row)
}

template <bool kExotic>
__device__ __forceinline__ void soma_body(const DevTables &t, const DevState &s, const StepVars &sv, const uint32_t block)
{
    __shared__ sfe_soma_class class_cache[kClassCache];
    __shared__ SomaScratch scr;
    unsigned long long *tl = nullptr;
    if constexpr (kTimelineAll)
        if (s.timeline != nullptr && threadIdx.x == 0 && block < 1024u)
            tl = s.timeline + kTimelineWords + ((sv.step_seq & 63ull) * 1024ull + block) * 16ull;
    if (tl != nullptr) tl[0] = global_timer_ns();
    griddep_launch_dependents();
    if (block >= sv.soma_grid) return;
    if (block >= t.n_soma_segments)
    {
        // extra CTAs: the fold of the PREVIOUS step (its statistics are complete once the previous
        // message phase has finished), off the critical path of this step
        griddep_wait();
        static_assert(sizeof(class_cache) >= kFinalScratchBytes, "the class cache doubles as the fold's scratch");
        finalize_body(t, s, sv, block - t.n_soma_segments, sv.soma_grid - t.n_soma_segments,
                reinterpret_cast<unsigned char *>(class_cache));
        return;
    }
    const bool classes_cached = t.n_soma_classes <= kClassCache;
    if (classes_cached)
    {
        // 8-byte words of the class table, cooperatively
        const unsigned long long *src = reinterpret_cast<const unsigned long long *>(t.classes);
        unsigned long long *dst = reinterpret_cast<unsigned long long *>(class_cache);
        for (uint32_t x = threadIdx.x; x < t.n_soma_classes * (sizeof(sfe_soma_class) / 8); x += kSomaThreads) dst[x] = __ldg(src + x);
    }
    griddep_wait(); // everything above reads load-time tables only
    if (tl != nullptr) tl[1] = global_timer_ns();
    if (s.x.n_peers > 0u && threadIdx.x < s.x.n_peers)
    {
        // the message phase of every earlier step has completed on this rank (stream order): tell the peers ...
        const uint32_t completed = static_cast<uint32_t>(sv.step_seq);
        if (block == 0) asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(s.x.done[threadIdx.x] + s.x.rank), "r"(completed) : "memory");
        // ... and hold this step's words back until every peer has completed the step whose buffer they reuse
        const uint32_t *peer_done = s.x.done[s.x.rank] + threadIdx.x;
        const long long t0 = clock64();
        while (static_cast<int32_t>(ld_relaxed_sys(peer_done) + 1u - completed) < 0)
        {
            __nanosleep(100);
            if (ld_relaxed_sys(s.x.error) != 0u) break;
            if (clock64() - t0 > 4000000000ll)
            {
                if (atomicCAS(s.x.error, 0u, 2u) == 0u)
                {
                    s.x.error[1] = threadIdx.x;
                    s.x.error[2] = completed;
                    s.x.error[3] = ld_relaxed_sys(peer_done);
                }
                break;
            }
        }
    }
    soma_segment<kExotic>(t, s, sv, block, classes_cached ? class_cache : nullptr, scr, sv.steps_done,
            static_cast<uint32_t>(sv.step_seq & 1ull));
    if (tl != nullptr) tl[2] = global_timer_ns();
}

#ifndef SFE_SOMA_CTAS
#define SFE_SOMA_CTAS 3
#endif
template <bool kExotic>
__global__ void __launch_bounds__(kSomaThreads, SFE_SOMA_CTAS) soma_kernel(const DevTables t, const DevState s, const StepVars sv)
{
    soma_body<kExotic>(t, s, sv, blockIdx.x);
}

// Design-space batches: many independent chips stepped by ONE launch per phase (the reference runs such sweeps as
// separate processes, scripts/tcad2025/compare_nemo_perf.py:52-101). The grid is the concatenation of the chips' own
// grids: a per-CTA map says which chip a CTA works for and which of that chip's CTAs it is. A slot carries the chip's
// tables and state as the single-chip kernels take them by value; the chips are in lock step, so what changes from step
// to step comes once with the launch.
struct BatchSlot
{
    DevTables t;
    DevState s;
    uint32_t tma_off, n_segments, final_grid, fanout_grid;
    uint32_t *work_pool; // the chip's ticket counters (kWorkPool of them)
    unsigned long long pad_to_16;
};
static_assert(sizeof(BatchSlot) % 16 == 0 || (sizeof(BatchSlot) + 8) % 16 == 0, "slots are copied in 8-byte words");

// A CTA of a batched launch first brings its chip's slot into shared memory (one coalesced read): the tables and the
// state are then a shared-memory window instead of a chain of dependent global loads behind every pointer.
__device__ __forceinline__ const BatchSlot &load_slot(const BatchSlot *slots, const uint32_t chip, unsigned long long *window)
{
    const unsigned long long *src = reinterpret_cast<const unsigned long long *>(slots + chip);
    for (uint32_t x = threadIdx.x; x < sizeof(BatchSlot) / 8; x += blockDim.x) window[x] = __ldg(src + x);
    __syncthreads();
    return *reinterpret_cast<const BatchSlot *>(window);
}
struct BatchStep
{
    long long steps_done;
    unsigned long long step_seq;
    long long fold_step;
    uint32_t fold, fold_parity; // the neuron-phase launch also folds the previous step
};
constexpr uint32_t kWorkPool = 8192; // >= 2 x the device log capacity (steps that can be enqueued before a collect)

__device__ __forceinline__ StepVars batch_vars(const BatchSlot &b, const BatchStep &a)
{
    StepVars sv;
    sv.steps_done = a.steps_done;
    sv.step_seq = a.step_seq;
    sv.fold_step = a.fold_step;
    sv.fold_parity = a.fold_parity;
    sv.fuse_next = 0u;
    sv.work = b.work_pool + (a.step_seq % kWorkPool);
    sv.soma_grid = b.n_segments + (a.fold != 0u ? b.final_grid : 0u);
    sv.fanout_grid = b.fanout_grid;
    return sv;
}

template <bool kExotic>
__global__ void __launch_bounds__(kSomaThreads, 3) soma_kernel_batch(const BatchSlot *__restrict__ slots, const uint2 *__restrict__ cta_map, const BatchStep a)
{
    __shared__ __align__(16) unsigned long long window[sizeof(BatchSlot) / 8];
    const uint2 where = __ldg(cta_map + blockIdx.x); // (chip, CTA of the chip)
    const BatchSlot &b = load_slot(slots, where.x, window);
    const StepVars sv = batch_vars(b, a);
    soma_body<kExotic>(b.t, b.s, sv, where.y);
}

// potentials of the probed neurons, after the neuron phase (src/chip.cpp:1071-1082)
__global__ void probe_kernel(const DevTables t, const DevState s)
{
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= t.n_probes + t.n_u_probes) return;
    if (p >= t.n_probes)
    {
        // model-defined trace: LIF input current (LoihiLifModel::get_neuron_traces, src/models.cpp:653-662)
        s.probe_out[p] = s.u[t.u_probes[p - t.n_probes]];
        return;
    }
    const uint32_t i = t.probes[p];
    const uint32_t model = t.classes[t.neuron_class[i]].model;
    double v = s.v[i];
    if (model == SFE_SOMA_HH) v = s.hh[t.neuron_aux[i]];
    else if (model == SFE_SOMA_INPUT) v = 0.0; // PipelineUnit::get_potential default
    else if (model == SFE_SOMA_DEVICE_MODEL)
    {
        const double *where = s.plug_potential[t.neuron_aux[i]];
        v = where != nullptr ? *where : 0.0;
    }
    s.probe_out[p] = v;
}

// Out-of-tree soma device models, first half of their step: the dendrite output of every instance for this timestep
// (what the neuron phase computes inline for the built-in somas), handed to the model's own kernel as plain arrays.
// One thread per instance; the accumulator cell is consumed here.
struct PlugGather
{
    uint32_t cell0;       // the neuron's accumulator cell in slot 0
    uint32_t stride;      // cells per slot (neurons of the core)
    uint32_t ring;
    uint32_t acc_mode;
    uint32_t fixed_slots;
    uint32_t charge_lost; // plain accumulator with the buffer inside the dendrite unit: a value, but always 0.0
    double inv_scale;
};
__global__ void plug_gather_kernel(const PlugGather *g, const uint32_t n, const DevState s, const StepVars sv, double *current_in, uint8_t *has_in)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const PlugGather r = g[k];
    const long long T = sv.steps_done + 1;
    const uint32_t slot = (r.ring > 1 && r.fixed_slots == 0u) ? static_cast<uint32_t>(T % r.ring) : 0u;
    const uint32_t d = r.cell0 + slot * r.stride;
    bool has = false;
    double in = 0.0;
    if (r.charge_lost != 0u) has = true; // src/models.cpp:78-82
    else if (r.acc_mode == SFE_ACC_PACKED17)
    {
        const uint32_t raw = s.din32[d];
        if (raw != 0u)
        {
            const uint32_t count = (raw + 0x10000u) >> 17;
            has = true;
            in = static_cast<double>(static_cast<int>(raw - (count << 17))) * r.inv_scale;
            s.din32[d] = 0u;
        }
    }
    else if (r.acc_mode == SFE_ACC_PACKED32)
    {
        const uint32_t raw = s.din32[d];
        if (raw != 0u)
        {
            has = true;
            in = static_cast<double>((static_cast<int>(raw << 12)) >> 12) * r.inv_scale;
            s.din32[d] = 0u;
        }
    }
    else if (r.acc_mode == SFE_ACC_DUAL32)
    {
        if (s.dcnt32[d] != 0u)
        {
            has = true;
            in = static_cast<double>(static_cast<int>(s.din32[d])) * r.inv_scale;
            s.din32[d] = 0u;
            s.dcnt32[d] = 0u;
        }
    }
    else if (s.dcnt32[d] != 0u)
    {
        has = true;
        in = s.din64[d];
        s.din64[d] = 0.0;
        s.dcnt32[d] = 0u;
    }
    current_in[k] = in;
    has_in[k] = has ? 1 : 0;
}

// One thread per tap line: MultiTapModel1D::update (src/models.cpp:237-257) for every event of the step that
// targets the line, in arrival order (the list is sorted that way at load): catch the line up to this timestep
// (calculate_next_state, :167-202, one RC step per elapsed timestep), add the current to the synapse's tap
// (input_current, :215-235); the value of tap 0 after the last event is what the soma reads next step.
__global__ void taps_kernel(const DevTables t, const DevState s, const StepVars sv)
{
    const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= t.n_taps_units) return;
    const TapsUnit d = t.taps_units[u];
    const long long T = sv.steps_done + 1;
    const uint32_t n = d.n_taps;
    double *v = s.tap_v + d.state_off, *next = s.tap_next + d.state_off;
    const double *tc = t.taps_values + d.const_off, *sc = tc + n;
    long long steps = s.tap_steps[u];
    for (uint32_t e = d.syn_begin; e < d.syn_begin + d.syn_count; ++e)
    {
        const TapsSyn y = t.taps_syn[e];
        if (((s.fired_global[y.src_bit >> 5] >> (y.src_bit & 31u)) & 1u) == 0u) continue;
        while (steps < T)
        {
            ++steps;
            for (uint32_t k = 0; k < n; ++k) next[k] = v[k] * tc[k];
            for (uint32_t src = 0; src < n; ++src)
            {
                if (src > 0)
                {
                    const double proximal = v[src] * sc[src - 1];
                    next[src - 1] = next[src - 1] + proximal;
                    next[src] = next[src] - proximal;
                }
                if (src + 1 < n)
                {
                    const double distal = v[src] * sc[src];
                    next[src + 1] = next[src + 1] + distal;
                    next[src] = next[src] - distal;
                }
            }
            for (uint32_t k = 0; k < n; ++k) v[k] = next[k];
        }
        v[y.tap] = v[y.tap] + y.weight;
        // what the event returns is buffered for ITS neuron (src/chip.cpp:738-764: timestep_buffer[post] = result)
        s.tap_buf[y.post] = v[0];
        s.tap_has[y.post] = 1u;
    }
    s.tap_steps[u] = steps;
}

__device__ __forceinline__ void ready_wait(const DevState &s, const unsigned long long epoch)
{
    if (threadIdx.x == 0 && ld_relaxed_sys(s.x.error) == 0u)
    {
        const uint32_t want = static_cast<uint32_t>(epoch) + 1u;
        const long long t0 = clock64();
        while (static_cast<int32_t>(ld_acquire_gpu(s.ready) - want) < 0)
        {
            __nanosleep(64);
            if (clock64() - t0 > 2000000000ll) // ~1 s: the producing launch never came (a host-side sequencing bug)
            {
                atomicExch(s.x.error, 2u);
                break;
            }
        }
    }
    __syncthreads();
}

// Multi-GPU: a rank sees the spikes of the whole chip as a fired-bit raster (SURVEY 8e:
// partition by destination core, exchange the fired-source set). The message phase derives the
// inbox word of 32 local axons by looking up the raster bit of each axon's source neuron.
// Raster lookups are plain (L1-cached, coherent-path) loads. The raster changes only between
// kernels (NCCL / copies) or, with the peer-memory exchange, before this CTA's acquire of the
// arrival flags; every lookup comes after that point, so cached lines are never stale. Not
// __ldg: the non-coherent path is outside the memory model's guarantees for data written
// while the kernel runs.
// How a message-phase thread reads raster word `index` of the current step: from the plain raster (NCCL / host-side
// exchange: complete before the kernel starts; plain L1-cached loads, never __ldg - the raster is rewritten between
// kernels), or from this rank's pair buffer of the peer-memory exchange (ll_word: polls until the word has landed).
struct RasterView
{
    const uint32_t *plain;
    const uint2 *pairs; // non-null: peer-memory exchange
    uint32_t want;      // epoch + 1
    uint32_t *error;
    __device__ __forceinline__ uint32_t word(const uint32_t index) const
    {
        return pairs != nullptr ? ll_word(pairs, index, want, error) : plain[index];
    }
};

// One inbox word = 32 local axons-in. Load-time table word_src tells how to derive it from the raster: an index
// below kWordGather = the 32 axons are fed by 32 consecutive neurons that fill one raster word (every dense
// projection between word-aligned populations: the word is copied), kWordEmpty = padding only, kWordGather = look
// up the raster bit of each axon's source neuron (axon_src).
constexpr uint32_t kWordGather = 0xFFFFFFFFu, kWordEmpty = 0xFFFFFFFEu;
__device__ __forceinline__ uint32_t gather_inbox_word(const uint32_t *__restrict__ src32, const RasterView &raster)
{
    uint32_t word = 0u;
#pragma unroll 1
    for (int h = 0; h < 2; ++h) // two halves: 16 source ids + 16 raster words in flight per thread
    {
        uint4 v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = __ldg(reinterpret_cast<const uint4 *>(src32) + 4 * h + q);
        uint32_t half = 0u;
#pragma unroll
        for (int q = 0; q < 4; ++q)
        {
            const uint32_t src[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
#pragma unroll
            for (int k = 0; k < 4; ++k)
            {
                // padding axons point at 0xFFFFFFFF
                const uint32_t r = src[k] != 0xFFFFFFFFu ? raster.word(src[k] >> 5) : 0u;
                half |= ((r >> (src[k] & 31u)) & 1u) << (4 * q + k);
            }
        }
        word |= half << (16 * h);
    }
    return word;
}
__device__ __forceinline__ uint32_t inbox_word_from_raster(const DevTables &t, const size_t local_word, const RasterView &raster)
{
    const uint32_t ws = __ldg(t.word_src + local_word);
    if (ws < kWordEmpty) return raster.word(ws);
    if (ws == kWordEmpty) return 0u;
    return gather_inbox_word(t.axon_src + (local_word << 5), raster);
}

// ---------------------------------------------------------------------------
// K3: message phase. One CTA per destination core.
// ---------------------------------------------------------------------------
constexpr int kFanoutThreads = 256;
constexpr int kFanoutWarps = kFanoutThreads / 32;
constexpr int kListCap = 2048; // active axons listed per round (16 KB of shared memory)
constexpr int kListExtra = 512; // of which, for cores with 4-byte records: extra chunks of axons with > 128 synapses
constexpr int kCostCache = 32;  // cost classes cached in shared memory

struct FanoutCounters
{
    uint32_t msgs;
    unsigned long long events, hop_e, hop_w, hop_n, hop_s;
    double syn_e, den_e, proc;
    double dup_e; // part of den_e spent in combined dendrite + soma units (only tracked outside the q4-only kernel)
};

template <bool kDup>
__device__ __forceinline__ void account_axon(
        FanoutCounters &c, const sfe_axon_in &ax, const sfe_cost_class &cc, const double lat_axon_in)
{
    // receive_message / sim_estimate_network_costs  src/chip.cpp:694-708,1127-1169
    const unsigned long long dx = SFE_HOP_DX(ax.hop), dy = SFE_HOP_DY(ax.hop);
    if (SFE_HOP_EAST(ax.hop)) c.hop_e += dx; else c.hop_w += dx;
    if (SFE_HOP_NORTH(ax.hop)) c.hop_n += dy; else c.hop_s += dy;
    c.msgs += 1;
    c.events += ax.syn_count;
    if (cc.per_message)
    {
        c.syn_e += cc.syn_energy;
        c.den_e += cc.den_energy;
        if (kDup && cc.den_is_soma != 0u) c.dup_e += cc.den_energy;
        c.proc += lat_axon_in + cc.syn_latency;
    }
    else
    {
        const double n = static_cast<double>(ax.syn_count);
        c.syn_e += n * cc.syn_energy;
        c.den_e += n * cc.den_energy;
        if (kDup && cc.den_is_soma != 0u) c.dup_e += n * cc.den_energy;
        c.proc += lat_axon_in + n * (cc.syn_latency + cc.den_latency);
    }
}

struct ChunkCursor
{
    uint32_t e, j0, count, stride;
    uint2 ent; // (padded segment offset, synapse count) of list entry e
};

// dendrite slot of a synaptic event: the slot of a delay line that is read at timestep T + 1 + delay, or - fixed slots -
// the compartment the synapse names in its delay field
__device__ __forceinline__ uint32_t dendrite_slot(const uint32_t ring, const bool fixed, const long long T, const uint32_t m)
{
    if (ring <= 1u) return 0u;
    return fixed ? SFE_SYN_DELAY(m) : static_cast<uint32_t>((T + 1 + SFE_SYN_DELAY(m)) % ring);
}

__device__ __forceinline__ void accumulate_one(uint32_t *acc32, uint32_t *cnt32, const uint32_t P, const uint32_t ring,
        const bool fixed_slots, const long long T, const double scale, const uint32_t packed_one, const double w, const uint32_t m)
{
    const uint32_t post = SFE_SYN_POST(m);
    const uint32_t sl = dendrite_slot(ring, fixed_slots, T, m);
    const int fixed = __double2int_rn(w * scale);
    // packed cells: sum + count * 2^20 (PACKED32) or sum + count * 2^17 (PACKED17)
    if (packed_one != 0u) atomicAdd(&acc32[sl * P + post], packed_one + static_cast<uint32_t>(fixed));
    else
    {
        atomicAdd(&acc32[sl * P + post], static_cast<uint32_t>(fixed));
        atomicAdd(&cnt32[sl * P + post], 1u);
    }
}

// ---- TMA (cp.async.bulk) + mbarrier primitives for the staged variant ------------
__device__ __forceinline__ uint32_t smem_addr(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(const uint32_t bar, const uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(const uint32_t bar, const uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(const uint32_t dst, const void *src, const uint32_t bytes, const uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(const uint32_t bar, const uint32_t parity)
{
    asm volatile("{\n"
                 ".reg .pred P1;\n"
                 "LAB_WAIT:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
                 "@P1 bra DONE;\n"
                 "bra LAB_WAIT;\n"
                 "DONE:\n"
                 "}" ::"r"(bar),
                 "r"(parity)
                 : "memory");
}

// The neuron phase of a WHOLE core by one CTA of kFanoutThreads threads, for the fused step kernel, where it sits on
// the step's critical path (it runs when the core's last work item retires, and the next step's raster is published
// after the last core): ONE memory round trip for the state of all the core's neurons - every array is brought into
// shared memory (idle at that point: accumulators, rings and list of the message phase) with 16-byte cp.async.cg -
// then the updates from shared memory, the coalesced write-back, and a second round trip only for the inbox bits of
// the neurons that fired (compacted, one thread per fired neuron, its axons' loads batched). soma_segment does the same
// work 512 neurons at a time with two dependent round trips each and relies on thousands of CTAs in flight to hide
// them. Cores it serves: LIF / TrueNorth somas (no input / plugin / noise / taps), exact accumulation modes, up to
// kCoreSomaMax neurons, 4-neuron aligned (CoreDev::fast_soma, decided at load).
constexpr uint32_t kCoreSomaMax = 1024;
constexpr uint32_t kCoreSomaPerThread = kCoreSomaMax / 256;
__host__ __device__ constexpr size_t core_soma_smem(const uint32_t P) // v, u, bias (f64); refractory, class, sum, count, axon_out_begin (+1)
{
    return static_cast<size_t>(P) * (3 * 8 + 5 * 4) + 16 + kClassCache * sizeof(sfe_soma_class) + 512;
}

__device__ __forceinline__ void cp_async16(const uint32_t dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

__device__ __forceinline__ void soma_core(const DevTables &t, const DevState &s, const CoreDev &core, unsigned char *smem,
        const long long steps_done, const uint32_t parity)
{
    const uint32_t P = core.neuron_count, nb = core.neuron_begin;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long T = steps_done + 1;
    const uint32_t slot = (core.ring > 1 && core.fixed_slots == 0u) ? static_cast<uint32_t>(T % core.ring) : 0u;
    const uint32_t d0 = core.dend_base + slot * P;
    const bool counted = core.acc_mode == SFE_ACC_DUAL32;
    // shared-memory layout
    double *v_s = reinterpret_cast<double *>(smem);
    double *u_s = v_s + P;
    double *b_s = u_s + P;
    int32_t *refr_s = reinterpret_cast<int32_t *>(b_s + P);
    uint32_t *cid_s = reinterpret_cast<uint32_t *>(refr_s + P);
    uint32_t *sum_s = cid_s + P;
    uint32_t *cnt_s = sum_s + P;
    uint32_t *aob_s = cnt_s + P; // P + 1 entries (+3 pad)
    sfe_soma_class *class_s = reinterpret_cast<sfe_soma_class *>(aob_s + P + 4);
    uint32_t *fired_list = reinterpret_cast<uint32_t *>(class_s + kClassCache); // [<= 96] overflow -> direct path
    __shared__ uint32_t n_fired_s;
    __shared__ double red_d[kFanoutWarps][3];
    __shared__ uint32_t red_u[kFanoutWarps][3];
    if (threadIdx.x == 0) n_fired_s = 0u;
    // ---- one round trip: everything the core's neurons need ------------------------------------
    for (uint32_t x = threadIdx.x; x < P / 2; x += kFanoutThreads) // 16 bytes = 2 doubles
    {
        cp_async16(smem_addr(v_s + 2 * x), s.v + nb + 2 * x);
        cp_async16(smem_addr(u_s + 2 * x), s.u + nb + 2 * x);
        cp_async16(smem_addr(b_s + 2 * x), s.bias + nb + 2 * x);
    }
    for (uint32_t x = threadIdx.x; x < P / 4; x += kFanoutThreads) // 16 bytes = 4 words
    {
        cp_async16(smem_addr(refr_s + 4 * x), s.refractory + nb + 4 * x);
        cp_async16(smem_addr(cid_s + 4 * x), t.neuron_class + nb + 4 * x);
        cp_async16(smem_addr(sum_s + 4 * x), s.din32 + d0 + 4 * x);
        if (counted) cp_async16(smem_addr(cnt_s + 4 * x), s.dcnt32 + d0 + 4 * x);
        cp_async16(smem_addr(aob_s + 4 * x), t.axon_out_begin + nb + 4 * x);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (threadIdx.x == 0) aob_s[P] = __ldg(t.axon_out_begin + nb + P);
    const bool classes_cached = t.n_soma_classes <= kClassCache;
    if (classes_cached)
    {
        const unsigned long long *src = reinterpret_cast<const unsigned long long *>(t.classes);
        unsigned long long *dst = reinterpret_cast<unsigned long long *>(class_s);
        for (uint32_t x = threadIdx.x; x < t.n_soma_classes * (sizeof(sfe_soma_class) / 8); x += kFanoutThreads) dst[x] = __ldg(src + x);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    uint32_t *const inbox = s.inbox + static_cast<size_t>(parity) * t.inbox_words;
    uint32_t n_updated = 0, n_fired = 0, n_packets = 0;
    double soma_e = 0.0, dend_e = 0.0, lat_sum = 0.0;
#pragma unroll
    for (uint32_t r = 0; r < kCoreSomaPerThread; ++r)
    {
        const uint32_t k = r * kFanoutThreads + threadIdx.x;
        const bool valid = k < P;
        int st = SFE_STATUS_IDLE;
        if (valid)
        {
            const uint32_t i = nb + k;
            const sfe_soma_class &c = classes_cached ? class_s[cid_s[k]] : t.classes[cid_s[k]];
            const uint32_t raw_sum = sum_s[k];
            bool has_in = false;
            double in = 0.0;
            if (c.dend_in_neuron && c.dend_model == SFE_DEND_ACCUMULATOR) has_in = true; // charge lost (src/models.cpp:78-82)
            else if (core.acc_mode == SFE_ACC_PACKED17)
            {
                if (raw_sum != 0u)
                {
                    const uint32_t count = (raw_sum + 0x10000u) >> 17;
                    has_in = true;
                    in = static_cast<double>(static_cast<int>(raw_sum - (count << 17))) * core.inv_scale;
                    s.din32[d0 + k] = 0u;
                }
            }
            else if (core.acc_mode == SFE_ACC_PACKED32)
            {
                if (raw_sum != 0u)
                {
                    has_in = true;
                    in = static_cast<double>((static_cast<int>(raw_sum << 12)) >> 12) * core.inv_scale;
                    s.din32[d0 + k] = 0u;
                }
            }
            else if (cnt_s[k] != 0u) // DUAL32
            {
                has_in = true;
                in = static_cast<double>(static_cast<int>(raw_sum)) * core.inv_scale;
                s.din32[d0 + k] = 0u;
                s.dcnt32[d0 + k] = 0u;
            }
            double lat = 0.0;
            if (c.dend_in_neuron)
            {
                dend_e += c.dend_energy_update;
                lat += c.dend_latency_update;
            }
            double v = v_s[k];
            if (c.model == SFE_SOMA_LIF)
            {
                double u = u_s[k];
                int refr = refr_s[k];
                st = lif_update(c, v, u, refr, b_s[k], has_in, in, steps_done);
                s.u[i] = u;
                s.refractory[i] = refr;
            }
            else st = truenorth_update(c, v, b_s[k], has_in, in);
            s.v[i] = v;
            s.status[i] = static_cast<uint8_t>(st);
            double e = c.energy_access, l = c.latency_access;
            if (st >= SFE_STATUS_UPDATED)
            {
                e += c.energy_update;
                l += c.latency_update;
                ++n_updated;
            }
            if (st == SFE_STATUS_FIRED)
            {
                e += c.energy_spike_out;
                l += c.latency_spike_out;
                ++n_fired;
                n_packets += aob_s[k + 1] - aob_s[k];
            }
            soma_e += e;
            lat_sum += lat + l;
        }
        const uint32_t ballot = __ballot_sync(0xffffffffu, st == SFE_STATUS_FIRED);
        if (lane == 0 && k < ((P + 31u) & ~31u)) s.fired_bits[core.fired_word_begin + (k >> 5)] = ballot;
        // unpartitioned chip: the fired neurons are listed, their axons' inbox bits raised below
        if (t.partitioned == 0u && ballot != 0u)
        {
            uint32_t base = 0u;
            if (lane == 0) base = atomicAdd(&n_fired_s, __popc(ballot));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (st == SFE_STATUS_FIRED)
            {
                const uint32_t at = base + __popc(ballot & ((1u << lane) - 1u));
                if (at < 96u) fired_list[at] = k;
                else
                {
                    // (more spikes than the list holds: raise this neuron's bits right here)
                    for (uint32_t a = aob_s[k]; a < aob_s[k + 1]; ++a)
                    {
                        const uint32_t bit = __ldg(t.axon_out_bit + a);
                        red_or(&inbox[bit >> 5], 1u << (bit & 31));
                    }
                }
            }
        }
    }
    __syncthreads();
    if (t.partitioned == 0u)
    {
        // pipeline_process_axon_out  src/chip.cpp:802-834: one message per axon of a fired neuron
        const uint32_t n_list = min(n_fired_s, 96u);
        // 8 lanes per fired neuron: lane j of the group raises the bits of axons j, j + 8, ... (one batch of loads in
        // flight for the whole core instead of a dependent chain per neuron)
        for (uint32_t f = threadIdx.x >> 3; f < n_list; f += kFanoutThreads >> 3)
        {
            const uint32_t k = fired_list[f];
            for (uint32_t a = aob_s[k] + (threadIdx.x & 7u); a < aob_s[k + 1]; a += 8u)
            {
                const uint32_t bit = __ldg(t.axon_out_bit + a);
                red_or(&inbox[bit >> 5], 1u << (bit & 31));
            }
        }
    }
    // ---- per-core statistics: warp shuffles, one barrier, fixed order ------------------------
    n_updated = warp_sum(n_updated);
    n_fired = warp_sum(n_fired);
    n_packets = warp_sum(n_packets);
    soma_e = warp_sum(soma_e);
    dend_e = warp_sum(dend_e);
    lat_sum = warp_sum(lat_sum);
    if (lane == 0)
    {
        red_u[warp][0] = n_updated;
        red_u[warp][1] = n_fired;
        red_u[warp][2] = n_packets;
        red_d[warp][0] = soma_e;
        red_d[warp][1] = dend_e;
        red_d[warp][2] = lat_sum;
    }
    __syncthreads();
    if (threadIdx.x < core.seg_count)
    {
        // the core's total goes into its first segment's record, the other segments' records are empty
        StatsN out = {0u, 0u, 0u, 0u, 0.0, 0.0, 0.0, 0.0};
        if (threadIdx.x == 0)
            for (int w = 0; w < kFanoutWarps; ++w)
            {
                out.updated += red_u[w][0];
                out.fired += red_u[w][1];
                out.packets += red_u[w][2];
                out.soma_e += red_d[w][0];
                out.dend_e += red_d[w][1];
                out.gen_sum += red_d[w][2];
            }
        s.stats_n[static_cast<size_t>(parity) * t.n_soma_segments + core.seg_begin + threadIdx.x] = out;
    }
    __syncthreads();
}

// Streaming variants of the exact-mode message phase (selected at engine creation,
// SFE_FANOUT=scalar|vector|tma):
//   kStreamScalar  8 scalar loads per 128-synapse chunk, next chunk's loads issued
//                  before the current chunk's atomics
//   kStreamTma     each warp owns a ring of kTmaStages shared-memory stages filled by
//                  cp.async.bulk (TMA) with mbarrier completion: lane 0 issues two bulk
//                  copies per chunk (weights, meta), the warp consumes from shared memory
//   kStreamQ4      engines whose cores ALL carry 4-byte records: the register-pipelined q4
//                  stream only (it is also a branch of the other two for mixed engines)
constexpr int kStreamScalar = 0, kStreamTma = 2, kStreamQ4 = 3;
constexpr int kQ4CtasPerSm = 4;
#ifndef SFE_Q4_STAGES
#define SFE_Q4_STAGES 6 // measured on C4, one B200: 4 stages 126.3, 6 stages 125.6, 8 stages 132.7 us per step
#endif
#ifndef SFE_Q4_BULK
#define SFE_Q4_BULK 0
#endif
constexpr int kQ4Stages = SFE_Q4_STAGES; // per-warp ring of 512-byte chunks (4-byte records)
constexpr int kQ4StageBytes = 512;       // 128 records
// SFE_Q4_BULK: the q4-only instantiation fills its rings with cp.async.bulk (one TMA copy per chunk, issued by lane 0,
// mbarrier completion) instead of one 16-byte cp.async per lane
constexpr bool kQ4Bulk = SFE_Q4_BULK != 0;
constexpr int kQ4BarBytes = kQ4Bulk ? kFanoutWarps * kQ4Stages * 8 : 0;
static_assert(kQ4Stages * kQ4StageBytes <= 6144, "a warp's 4-byte-record ring reuses its TMA stages in mixed engines");
constexpr int kTmaStages = 4;
constexpr int kTmaStageBytes = 128 * 12; // 128 fp64 weights + 128 u32 meta words

// kFused: the message phase also folds the step (fused_finalize, SFE_FUSED_FINALIZE=1)
template <int V, bool kFused>
__device__ __forceinline__ void fanout_body(const DevTables &t, const DevState &s, const StepVars &sv, const uint32_t tma_off, const uint32_t block)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double part_d[kFanoutWarps][4];
    __shared__ unsigned long long part_l[kFanoutWarps][6];
    __shared__ sfe_cost_class cost_cache[kCostCache];
    __shared__ uint32_t next_item, list_n, list_total, core_event;
    __shared__ uint32_t scan_w[kFanoutWarps];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if constexpr (kTimelineAll && !kFused)
        if (s.timeline != nullptr && threadIdx.x == 0) s.timeline[((sv.step_seq & 63ull) * 1024ull + block) * 16ull + 12] = global_timer_ns();
    griddep_launch_dependents();
    if (block >= sv.fanout_grid) return;
    const bool costs_cached = t.n_cost_classes <= kCostCache;
    if (costs_cached)
        for (uint32_t x = threadIdx.x; x < t.n_cost_classes; x += kFanoutThreads) cost_cache[x] = t.costs[x];
    const sfe_cost_class *cost_table = costs_cached ? cost_cache : t.costs;
    unsigned char *tma_base = smem_raw + tma_off;
    unsigned long long *tma_bars = reinterpret_cast<unsigned long long *>(tma_base + kFanoutWarps * kTmaStages * kTmaStageBytes);
    // staging area at tma_off: the TMA stages + mbarriers (kStreamTma; a warp's 4-byte-record ring
    // reuses its stages), or the cp.async rings alone when the engine has 4-byte records
    const uint32_t list_off = tma_off + (V == kStreamTma ? kFanoutWarps * kTmaStages * (kTmaStageBytes + 8)
                                                          : (t.syn_q4 != nullptr ? kFanoutWarps * kQ4Stages * kQ4StageBytes + kQ4BarBytes : 0));
    if constexpr (V == kStreamTma)
    {
        if (threadIdx.x < kFanoutWarps * kTmaStages) mbar_init(smem_addr(tma_bars + threadIdx.x), 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    unsigned long long *q4_bars = reinterpret_cast<unsigned long long *>(tma_base + kFanoutWarps * kQ4Stages * kQ4StageBytes);
    uint32_t q4_phase = 0u;
    (void) q4_phase;
    (void) q4_bars;
    if constexpr (V == kStreamQ4 && kQ4Bulk)
    {
        if (threadIdx.x < kFanoutWarps * kQ4Stages) mbar_init(smem_addr(q4_bars + threadIdx.x), 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint32_t tma_phase = 0u; // bit st = parity the next wait on stage st of this warp expects
    (void) tma_phase;
    // Two-kernel step: everything above reads load-time tables / initialises shared memory only; the data of the
    // neuron-phase kernel is awaited here. Fused step kernel (kFused): no grid dependency at all - the kernel waits,
    // below, for the READY FLAG of its step (raised by the launch that ran the step's neuron phase: the previous
    // fused launch, or the publish kernel after a stand-alone neuron phase), so its CTAs start on the SMs the
    // previous launch frees while that launch still folds its step.
    if constexpr (!kFused)
    {
        // Still table-only: walk the descriptor chain of this CTA's first work item (its block index, see below) so
        // that the lines sit in L1 when the item is opened after the wait - three dependent loads off the critical path.
        if (block < t.n_fan_items)
        {
            const FanItem *it0 = t.fan_items + __ldg(t.fan_order + block);
            const uint32_t c0 = __ldg(&it0->core), w0 = __ldg(&it0->word_lo);
            const CoreDev *cd0 = t.cores + c0;
            if (threadIdx.x < 2u) asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const char *>(cd0) + 128u * threadIdx.x));
            if (t.partitioned != 0u)
            {
                const size_t local = static_cast<size_t>(__ldg(&cd0->inbox_word_begin)) + w0 - t.inbox_lo + threadIdx.x;
                asm volatile("prefetch.global.L1 [%0];" ::"l"(t.word_src + local));
            }
        }
        griddep_wait();
    }
    if constexpr (kTimelineAll && !kFused)
        if (s.timeline != nullptr && threadIdx.x == 0) s.timeline[((sv.step_seq & 63ull) * 1024ull + block) * 16ull + 13] = global_timer_ns();
    const long long T = sv.steps_done + 1;
    uint32_t *const inbox = s.inbox + static_cast<size_t>(sv.step_seq & 1ull) * t.inbox_words;

    // Partitioned chip: the inbox is derived from the exchanged raster. With the peer-memory exchange there is nothing
    // to wait for here: every raster word is awaited where it is read (RasterView / ll_word).
    RasterView raster;
    raster.plain = s.fired_global;
    raster.pairs = s.x.n_peers > 0u ? s.x.ll[s.x.rank] + (sv.step_seq & 1ull) * s.x.fired_words : nullptr;
    raster.want = static_cast<uint32_t>(sv.step_seq) + 1u;
    raster.error = s.x.error;
    const bool gather = t.partitioned != 0u;
    // Work items are handed out through an atomic ticket counter of the step's own (sv.work points at the step's slot
    // of a pool the host clears per batch: launches of consecutive steps overlap in the fused step kernel, and a
    // straggler's last, failed draw must not eat a ticket of a later step). Fused step kernel: the first item of a CTA
    // is its block index - no atomic, and its descriptors are fetched before the ready flag is awaited.
    bool first_item = true;
    (void) first_item;
    unsigned long long *const tl = ((kFused || kTimelineAll) && s.timeline != nullptr)
            ? s.timeline + ((sv.step_seq & 63ull) * (kTimelineAll ? 1024ull : sv.fanout_grid) + block) * 16ull
            : nullptr;
    uint32_t tl_item = 0u;
    auto stamp = [&](const uint32_t k) {
        if ((kFused || kTimelineAll) && tl != nullptr && threadIdx.x == 0 && k < 16u) tl[k] = global_timer_ns();
    };
    stamp(0);

    // Persistent CTAs: cores (heaviest first) are handed out through an atomic ticket,
    // so the grid is one resident wave and no SM idles behind a wave boundary.
    for (;;)
    {
    __syncthreads(); // previous core fully retired (smem accumulators, next_item)
    if (first_item)
    {
        if (threadIdx.x == 0) next_item = block; // no atomic for the first item; its descriptors are on their way already
        if constexpr (!kFused) first_item = false;
    }
    else if (threadIdx.x == 0) next_item = atomicAdd(sv.work, 1u) + sv.fanout_grid;
    __syncthreads();
    const uint32_t ticket = next_item;
    if (ticket >= t.n_fan_items) break;
    const uint32_t item_id = t.fan_order[ticket];
    const FanItem item = t.fan_items[item_id];
    const uint32_t ci = item.core;
    const CoreDev core = t.cores[ci];
    if (kFused && first_item)
    {
        // the descriptors of the first item are in registers: now wait for the step's raster / inbox to be complete
        ready_wait(s, sv.step_seq);
        first_item = false;
        stamp(1);
    }
    stamp(2u + 4u * tl_item);
    // kStreamQ4 instantiation: every core of the engine is certified for 4-byte records, the
    // other accumulation modes and streaming variants are compiled out (fewer registers)
    const uint32_t acc_mode = V == kStreamQ4 ? static_cast<uint32_t>(SFE_ACC_PACKED17) : core.acc_mode;
    const bool is_q4 = V == kStreamQ4 || core.q4 != 0u;
    const uint32_t P = core.neuron_count;
    const uint32_t cells = P * core.ring;
    // does accumulated charge survive? (not for a plain accumulator with the buffer
    // inside the dendrite unit — the neuron phase zeroes it; only counters matter)
    const bool accumulate = core.dend_in_msg != 0 && cells > 0;

    uint32_t *acc32 = reinterpret_cast<uint32_t *>(smem_raw);          // PACKED32 / DUAL32 sum
    uint32_t *cnt32 = acc32 + cells;                                   // DUAL32 count / ORDERED has
    double *acc64 = reinterpret_cast<double *>(smem_raw);              // ORDERED (cnt32 placed after)
    if (acc_mode == SFE_ACC_ORDERED) cnt32 = reinterpret_cast<uint32_t *>(acc64 + cells);
    // a core too large for shared-memory accumulators works on the HBM arrays directly: the exact modes add with
    // global atomics (sums commute), the ordered mode is one warp that owns the core's cells for the whole phase
    const bool in_hbm = V != kStreamQ4 && core.acc_global != 0u;
    if (in_hbm)
    {
        acc32 = s.din32 + core.dend_base;
        cnt32 = s.dcnt32 + core.dend_base;
        acc64 = s.din64 + core.dend_base;
    }

    if (accumulate && !in_hbm)
    {
        if (acc_mode == SFE_ACC_PACKED32 || acc_mode == SFE_ACC_PACKED17)
            for (uint32_t x = threadIdx.x; x < cells; x += kFanoutThreads) acc32[x] = 0u;
        else if (acc_mode == SFE_ACC_DUAL32)
            for (uint32_t x = threadIdx.x; x < 2 * cells; x += kFanoutThreads) acc32[x] = 0u;
        else
            for (uint32_t x = threadIdx.x; x < cells; x += kFanoutThreads)
            {
                // a delay-line slot that already holds charge keeps adding to it in
                // event order (value_or(0.0) + w, src/models.cpp:123-125): seed from HBM
                acc64[x] = core.ring > 1 ? s.din64[core.dend_base + x] : 0.0;
                cnt32[x] = core.ring > 1 ? s.dcnt32[core.dend_base + x] : 0u;
            }
    }
    __syncthreads();

    FanoutCounters cnt = {0u, 0ull, 0ull, 0ull, 0ull, 0ull, 0.0, 0.0, 0.0, 0.0};
    const uint32_t n_words = item.word_hi; // this item's slice of the core's inbox: [word_lo, word_hi)
    const double *__restrict__ w_base = t.syn_w + core.syn_begin;
    const uint32_t *__restrict__ m_base = t.syn_meta + core.syn_begin;
    const uint32_t ring = V == kStreamQ4 ? 1u : core.ring;
    const bool fixed_slots = V != kStreamQ4 && core.fixed_slots != 0u;

    // Ordered mode, shared-memory accumulators, TMA instantiation: the CTA builds the list of active axons like the exact
    // modes do (it comes out in arrival order: ascending inbox bit), then ONE warp walks the list in order, its segments
    // prefetched through the warp's TMA ring, and adds in synapse order - the reference's sequential loop without a
    // dependent global load on its path (the stand-alone loop further down pays two round trips per message).
    const bool ordered_stream = V == kStreamTma && acc_mode == SFE_ACC_ORDERED && !in_hbm;
    if (acc_mode != SFE_ACC_ORDERED || ordered_stream)
    {
        // Exact fixed-point accumulation: any order gives the reference's sums bit for
        // bit (load-time certificate). Per round (normally one per core):
        //   pass 1  one inbox word per thread, block scan of the popcounts -> every
        //           active axon gets a slot in a CTA-wide shared-memory list
        //   pass 2  the axon records of the listed axons are fetched with independent
        //           loads (one global latency for the whole core), accounted, and the
        //           list is rewritten as (segment offset, synapse count)
        //   stream  warp w takes entries w, w+8, ... and streams their CSR segments
        const uint32_t packed = acc_mode == SFE_ACC_PACKED32 ? (1u << 20) : acc_mode == SFE_ACC_PACKED17 ? (1u << 17) : 0u;
        uint2 *list = reinterpret_cast<uint2 *>(smem_raw + list_off);
        for (uint32_t wb = item.word_lo; wb < n_words;)
        {
            const uint32_t wi = wb + threadIdx.x;
            uint32_t word = 0u;
            if (wi < n_words)
                word = gather ? inbox_word_from_raster(t, core.inbox_word_begin + wi - t.inbox_lo, raster)
                              : inbox[core.inbox_word_begin + wi];
            const uint32_t pc = __popc(word);
            uint32_t incl = pc;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += up;
            }
            if (lane == 31) scan_w[warp] = incl;
            __syncthreads();
            uint32_t base = incl - pc, total = 0u;
#pragma unroll
            for (int w = 0; w < kFanoutWarps; ++w)
            {
                const uint32_t x = scan_w[w];
                if (w < warp) base += x;
                total += x;
            }
            // 4-byte records: the tail of the list is kept for the extra chunks of axons with more than 128 synapses
            const bool ok = base + pc <= (is_q4 ? kListCap - kListExtra : kListCap);
            const uint32_t accepted = __syncthreads_count(ok); // ok is monotone in the word index
            if (threadIdx.x == accepted || (accepted == kFanoutThreads && threadIdx.x == 0))
                list_n = list_total = accepted == kFanoutThreads ? total : base;
            if (ok && word != 0u)
            {
                if (!gather) inbox[core.inbox_word_begin + wi] = 0u; // consume
                uint32_t bits = word, slot = base;
                while (bits != 0u)
                {
                    const uint32_t b = __ffs(bits) - 1;
                    bits &= bits - 1u;
                    list[slot++] = make_uint2((wi << 5) + b, 0u);
                }
            }
            __syncthreads();
            const uint32_t n_list = list_n;
            for (uint32_t e = threadIdx.x; e < n_list; e += kFanoutThreads)
            {
                const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(t.axons_in + core.axon_begin + list[e].x));
                sfe_axon_in ax;
                ax.syn_off = raw.x;
                ax.syn_count = raw.y;
                ax.hop = raw.z;
                ax.cost_class = raw.w;
                account_axon<V != kStreamQ4>(cnt, ax, cost_table[ax.cost_class], core.lat_axon_in);
                if (!is_q4) list[e] = make_uint2(ax.syn_off, ax.syn_count);
                else
                {
                    // 4-byte records: the stream works on chunks of <= 128 records, entry = (first record, number of
                    // 16-byte pieces). The further chunks of a long axon go to the tail of the list (the order of
                    // exact accumulation is free); chunks that do not fit there are added right here, record by record.
                    list[e] = make_uint2(ax.syn_off, (min(ax.syn_count, 128u) + 3u) >> 2);
                    if (ax.syn_count > 128u && accumulate)
                    {
                        const uint32_t extra = (ax.syn_count - 1u) >> 7;
                        const uint32_t at = atomicAdd(&list_total, extra);
                        for (uint32_t k = 0; k < extra; ++k)
                        {
                            const uint32_t off = ax.syn_off + 128u * (k + 1u);
                            const uint32_t w4 = (min(ax.syn_count - 128u * (k + 1u), 128u) + 3u) >> 2;
                            if (at + k < kListCap) list[at + k] = make_uint2(off, w4);
                            else
                                for (uint32_t j = 0; j < 4u * w4; ++j)
                                {
                                    const uint32_t qv = __ldg(t.syn_q4 + core.syn_begin + off + j);
                                    atomicAdd(&acc32[(qv & 0x3FFCu) >> 2], qv >> 14);
                                }
                        }
                    }
                }
            }
            __syncthreads();
            wb += accepted;
            stamp(3u + 4u * tl_item);
            if (!accumulate || static_cast<uint32_t>(warp) >= n_list || (ordered_stream && warp != 0)) continue;

            // ---- stream the segments: chunk = 128 consecutive synapses of one axon ------
            // Segments are padded to 4 synapses and 16-byte aligned in HBM (engine-side layout).
            ChunkCursor cur;
            cur.e = ordered_stream ? 0u : warp;
            cur.j0 = 0u;
            cur.count = n_list;
            cur.stride = ordered_stream ? 1u : kFanoutWarps;
            cur.ent = list[cur.e];
            if constexpr (V == kStreamQ4 && kQ4Bulk)
            {
                // 4-byte records through the TMA engine: lane 0 issues ONE bulk copy per chunk into the warp's ring
                // (mbarrier completion), every lane reads back its 16 bytes (lane-major layout, q4_position) and adds
                // its four records; a warp barrier precedes the refill of a stage.
                const uint32_t acc_s = smem_addr(acc32);
                const uint32_t ring_w = smem_addr(tma_base + warp * kQ4Stages * kQ4StageBytes);
                const uint32_t bar0 = smem_addr(q4_bars + warp * kQ4Stages);
                const uint32_t *q_core = t.syn_q4 + core.syn_begin;
                const uint32_t n_total = min(list_total, static_cast<uint32_t>(kListCap));
                const uint32_t my_chunks = (n_total - warp + kFanoutWarps - 1u) / kFanoutWarps;
                const uint2 *my_list = list + warp;
                auto add_q4 = [&](const uint32_t qv) {
                    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(acc_s + (qv & 0x3FFCu)), "r"(qv >> 14) : "memory");
                };
                auto issue = [&](const uint32_t c, const int st) {
                    if (c < my_chunks && lane == 0)
                    {
                        const uint2 ent = my_list[c * kFanoutWarps];
                        mbar_expect_tx(bar0 + 8u * st, ent.y * 16u);
                        bulk_g2s(ring_w + st * kQ4StageBytes, q_core + ent.x, ent.y * 16u, bar0 + 8u * st);
                    }
                };
#pragma unroll
                for (int st = 0; st < kQ4Stages; ++st) issue(st, st);
                for (uint32_t c0 = 0; c0 < my_chunks; c0 += kQ4Stages)
                {
#pragma unroll
                    for (int st = 0; st < kQ4Stages; ++st)
                    {
                        const uint32_t c = c0 + st;
                        if (c >= my_chunks) break;
                        mbar_wait(bar0 + 8u * st, (q4_phase >> st) & 1u);
                        q4_phase ^= 1u << st;
                        if (lane < my_list[c * kFanoutWarps].y)
                        {
                            uint4 rec;
                            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                                         : "=r"(rec.x), "=r"(rec.y), "=r"(rec.z), "=r"(rec.w)
                                         : "r"(ring_w + st * kQ4StageBytes + 16 * lane)
                                         : "memory");
                            add_q4(rec.x);
                            add_q4(rec.y);
                            add_q4(rec.z);
                            add_q4(rec.w);
                        }
                        __syncwarp(); // every lane has read the stage before it is refilled
                        issue(c + kQ4Stages, st);
                    }
                }
            }
            else if (is_q4)
            {
                // 4-byte records (q4_record). Each warp owns a ring of kQ4Stages 512-byte chunks in shared
                // memory filled by cp.async: lane l moves the l-th 16 bytes of a chunk and, thanks to the
                // lane-major layout of the table (q4_position), those are the four records it accumulates
                // - it waits for its own copies only (no warp barrier), reads them back with one 16-byte
                // load and adds each with one red.shared: address = record & 0x3FFC (+ the accumulators'
                // base as an immediate), addend = record >> 14. Pad records are zero words (add 0 to cell
                // 0), so a lane never tests its records one by one. Nothing is held in registers while in
                // flight: seven chunks per warp are always on their way.
                uint32_t acc_s = smem_addr(acc32);
                uint32_t ring_s = smem_addr(
                        tma_base + warp * (V == kStreamTma ? kTmaStages * kTmaStageBytes : kQ4Stages * kQ4StageBytes) + 16 * lane);
                const uint32_t *q_lane = t.syn_q4 + core.syn_begin + 4u * lane;
                const uint32_t n_total = min(list_total, static_cast<uint32_t>(kListCap));
                // this warp's chunks: list entries warp, warp + 8, ...; chunk c lives in stage c % kQ4Stages
                const uint32_t my_chunks = (n_total - warp + kFanoutWarps - 1u) / kFanoutWarps;
                uint32_t list_s = smem_addr(list + warp); // entry of chunk c: list_s + 64 * c
                uint32_t lane_r = lane;
                // opaque copies: the five values the loop lives on stay in registers (the compiler otherwise re-derives
                // the shared-window addresses and the lane's base pointer from the kernel parameters for every chunk)
                asm volatile("mov.u64 %0, %0;" : "+l"(q_lane));
                asm volatile("mov.u32 %0, %0;" : "+r"(acc_s));
                asm volatile("mov.u32 %0, %0;" : "+r"(ring_s));
                asm volatile("mov.u32 %0, %0;" : "+r"(list_s));
                asm volatile("mov.u32 %0, %0;" : "+r"(lane_r));
                auto add_q4 = [&](const uint32_t qv) {
                    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(acc_s + (qv & 0x3FFCu)), "r"(qv >> 14) : "memory");
                };
                // copy of chunk c (entry at list_s + 64 * (c - c_base) + off) into stage st
                auto issue = [&](const bool live, const uint32_t entry_s, const int st) {
                    if (live)
                    {
                        uint32_t ex, ey;
                        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(ex), "=r"(ey) : "r"(entry_s) : "memory");
                        if (lane_r < ey)
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring_s + st * kQ4StageBytes),
                                         "l"(q_lane + ex)
                                         : "memory");
                    }
                    asm volatile("cp.async.commit_group;" ::: "memory"); // empty groups keep the count uniform
                };
#pragma unroll
                for (int st = 0; st < kQ4Stages; ++st) issue(static_cast<uint32_t>(st) < my_chunks, list_s + 64u * st, st);
                for (uint32_t left = my_chunks; left != 0u;)
                {
#pragma unroll
                    for (int st = 0; st < kQ4Stages; ++st)
                    {
                        if (left == 0u) break;
                        asm volatile("cp.async.wait_group %0;" ::"n"(kQ4Stages - 1) : "memory");
                        uint32_t w4;
                        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w4) : "r"(list_s + 64u * st + 4u) : "memory");
                        if (lane_r < w4)
                        {
                            uint4 rec;
                            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                                         : "=r"(rec.x), "=r"(rec.y), "=r"(rec.z), "=r"(rec.w)
                                         : "r"(ring_s + st * kQ4StageBytes)
                                         : "memory");
                            add_q4(rec.x);
                            add_q4(rec.y);
                            add_q4(rec.z);
                            add_q4(rec.w);
                        }
                        // the lane refills its own 16 bytes: no barrier needed
                        issue(left > static_cast<uint32_t>(kQ4Stages), list_s + 64u * (st + kQ4Stages), st);
                        --left;
                    }
                    list_s += 64u * kQ4Stages;
                }
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            else if constexpr (V == kStreamQ4)
            {
            }
            else if constexpr (V == kStreamTma)
            {
                unsigned char *stage0 = tma_base + warp * kTmaStages * kTmaStageBytes;
                const uint32_t bar0 = smem_addr(tma_bars + warp * kTmaStages);
                auto issue = [&](const int st) -> uint32_t {
                    if (cur.e >= cur.count) return 0u;
                    const uint32_t rem = cur.ent.y - cur.j0;
                    const uint32_t base_syn = cur.ent.x + cur.j0;
                    const uint32_t n4 = (min(rem, 128u) + 3u) & ~3u;
                    if (lane == 0)
                    {
                        const uint32_t bar = bar0 + 8u * st;
                        const uint32_t dst = smem_addr(stage0 + st * kTmaStageBytes);
                        mbar_expect_tx(bar, n4 * 12u);
                        bulk_g2s(dst, w_base + base_syn, n4 * 8u, bar);
                        bulk_g2s(dst + 1024u, m_base + base_syn, n4 * 4u, bar);
                    }
                    cur.j0 += 128u;
                    if (cur.j0 >= cur.ent.y)
                    {
                        cur.j0 = 0u;
                        cur.e += cur.stride;
                        if (cur.e < cur.count) cur.ent = list[cur.e];
                    }
                    return rem;
                };
                uint32_t rem_s[kTmaStages];
#pragma unroll
                for (int st = 0; st < kTmaStages; ++st) rem_s[st] = issue(st);
                bool done = false;
                while (!done)
                {
#pragma unroll
                    for (int st = 0; st < kTmaStages; ++st)
                    {
                        if (rem_s[st] == 0u)
                        {
                            done = true;
                            break;
                        }
                        mbar_wait(bar0 + 8u * st, (tma_phase >> st) & 1u);
                        tma_phase ^= 1u << st;
                        const double *ws = reinterpret_cast<const double *>(stage0 + st * kTmaStageBytes);
                        const uint32_t *ms = reinterpret_cast<const uint32_t *>(stage0 + st * kTmaStageBytes + 1024);
                        if (ordered_stream)
                        {
                            // synapse order: 32 at a time, lanes that hit the same cell one after the other, lowest first
#pragma unroll 1
                            for (int u = 0; u < 4; ++u)
                            {
                                const uint32_t j = lane + 32u * u;
                                const bool valid = j < rem_s[st];
                                double w = 0.0;
                                uint32_t cell = 0xffffffffu - lane; // distinct dummies for idle lanes
                                if (valid)
                                {
                                    w = ws[j];
                                    const uint32_t m = ms[j];
                                    cell = dendrite_slot(ring, fixed_slots, T, m) * P + SFE_SYN_POST(m);
                                }
                                // Do two lanes hit the same cell? Every lane tags its cell (the "has a value" word, set to 1
                                // below either way) and reads the tag back: all lanes see their own tag <=> no duplicate.
                                // (match.any resolves one distinct value per pass: ~10x the cost of this test, and the synapses
                                // of one axon almost never share a post-synaptic neuron.)
                                if (valid) cnt32[cell] = static_cast<uint32_t>(lane) + 2u;
                                __syncwarp();
                                const bool lost = valid && cnt32[cell] != static_cast<uint32_t>(lane) + 2u;
                                if (__ballot_sync(0xffffffffu, lost) == 0u)
                                {
                                    if (valid)
                                    {
                                        acc64[cell] = acc64[cell] + w;
                                        cnt32[cell] = 1u;
                                    }
                                    __syncwarp();
                                    continue;
                                }
                                const uint32_t peers = __match_any_sync(0xffffffffu, cell);
                                const int rank = __popc(peers & ((1u << lane) - 1u));
                                const int rounds = __reduce_max_sync(0xffffffffu, __popc(peers));
                                for (int r = 0; r < rounds; ++r)
                                {
                                    if (valid && rank == r)
                                    {
                                        acc64[cell] = acc64[cell] + w;
                                        cnt32[cell] = 1u;
                                    }
                                    __syncwarp();
                                }
                            }
                        }
                        else
                        {
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                            {
                                const uint32_t j = lane + 32u * u;
                                if (j < rem_s[st]) accumulate_one(acc32, cnt32, P, ring, fixed_slots, T, core.scale, packed, ws[j], ms[j]);
                            }
                        }
                        __syncwarp(); // every lane is done with the stage before it is refilled
                        rem_s[st] = issue(st);
                    }
                }
            }
            else
            {
                double cw[4], nw[4];
                uint32_t cm[4], nm[4];
                uint32_t cvalid = 0u, nvalid = 0u;
                bool more = true;
                while (more || cvalid != 0u)
                {
                    nvalid = 0u;
                    if (more)
                    {
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                        {
                            const uint32_t j = cur.j0 + lane + 32u * u;
                            if (j < cur.ent.y)
                            {
                                nw[u] = __ldg(w_base + cur.ent.x + j);
                                nm[u] = __ldg(m_base + cur.ent.x + j);
                                nvalid |= 1u << u;
                            }
                        }
                        cur.j0 += 128u;
                        if (cur.j0 >= cur.ent.y)
                        {
                            cur.j0 = 0u;
                            cur.e += cur.stride;
                            more = cur.e < cur.count;
                            if (more) cur.ent = list[cur.e];
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (cvalid & (1u << u)) accumulate_one(acc32, cnt32, P, ring, fixed_slots, T, core.scale, packed, cw[u], cm[u]);
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                    {
                        cw[u] = nw[u];
                        cm[u] = nm[u];
                    }
                    cvalid = nvalid;
                }
            }
        }
    }
    else if (V != kStreamQ4 && warp == 0)
    {
        // Ordered mode: one warp replays the messages in arrival order and adds
        // in synapse order, exactly like the reference's sequential loop.
        for (uint32_t wi = 0; wi < n_words; ++wi)
        {
            uint32_t word = 0u;
            if (gather)
            {
                const uint32_t src = __ldg(t.axon_src + (static_cast<size_t>(core.inbox_word_begin + wi - t.inbox_lo) << 5) + lane);
                const uint32_t r = src != 0xFFFFFFFFu ? raster.word(src >> 5) : 0u;
                word = __ballot_sync(0xffffffffu, ((r >> (src & 31u)) & 1u) != 0u);
            }
            else
            {
                if (lane == 0)
                {
                    word = inbox[core.inbox_word_begin + wi];
                    if (word != 0u) inbox[core.inbox_word_begin + wi] = 0u;
                }
                word = __shfl_sync(0xffffffffu, word, 0);
            }
            while (word != 0u)
            {
                const uint32_t b = __ffs(word) - 1;
                word &= word - 1u;
                const sfe_axon_in ax = t.axons_in[core.axon_begin + (wi << 5) + b];
                if (lane == 0) account_axon<true>(cnt, ax, cost_table[ax.cost_class], core.lat_axon_in);
                if (!accumulate) continue;
                for (uint32_t j0 = 0; j0 < ax.syn_count; j0 += 32)
                {
                    const uint32_t j = j0 + lane;
                    const bool valid = j < ax.syn_count;
                    double w = 0.0;
                    uint32_t cell = 0xffffffffu - lane; // distinct dummies for idle lanes
                    if (valid)
                    {
                        w = __ldg(w_base + ax.syn_off + j);
                        const uint32_t m = __ldg(m_base + ax.syn_off + j);
                        cell = dendrite_slot(ring, fixed_slots, T, m) * P + SFE_SYN_POST(m);
                    }
                    // lanes that hit the same cell add one after the other, lowest lane first
                    const uint32_t peers = __match_any_sync(0xffffffffu, cell);
                    const int rank = __popc(peers & ((1u << lane) - 1u));
                    const int rounds = __reduce_max_sync(0xffffffffu, __popc(peers));
                    for (int r = 0; r < rounds; ++r)
                    {
                        if (valid && rank == r)
                        {
                            if (in_hbm)
                            {
                                // L2 accesses: a cell may have been written by another lane a moment ago
                                __stcg(&acc64[cell], __ldcg(&acc64[cell]) + w);
                                __stcg(&cnt32[cell], 1u);
                            }
                            else
                            {
                                acc64[cell] = acc64[cell] + w;
                                cnt32[cell] = 1u;
                            }
                        }
                        __syncwarp();
                    }
                }
            }
        }
    }
    __syncthreads();

    // ---- write the dendrite state back (coalesced) ------------------------------
    if (accumulate && !in_hbm)
    {
        const uint32_t base = core.dend_base;
        // Exact modes: several items (inbox slices) of one core may run on different CTAs,
        // and integer sums commute, so partial sums are merged with fire-and-forget global
        // atomics; the neuron phase zeroes a slot when it consumes it.
        if (acc_mode == SFE_ACC_PACKED32 || acc_mode == SFE_ACC_PACKED17)
        {
            for (uint32_t x = threadIdx.x; x < cells; x += kFanoutThreads)
                if (acc32[x] != 0u) red_add(&s.din32[base + x], acc32[x]);
        }
        else if (acc_mode == SFE_ACC_DUAL32)
        {
            for (uint32_t x = threadIdx.x; x < cells; x += kFanoutThreads)
                if (cnt32[x] != 0u)
                {
                    red_add(&s.din32[base + x], acc32[x]);
                    red_add(&s.dcnt32[base + x], cnt32[x]);
                }
        }
        else
        {
            for (uint32_t x = threadIdx.x; x < cells; x += kFanoutThreads)
            {
                s.din64[base + x] = acc64[x];
                s.dcnt32[base + x] = cnt32[x];
            }
        }
    }

    // ---- per-item counters: warp shuffles, one barrier, fixed order ---------------------
    {
        const unsigned long long v_msgs = warp_sum(static_cast<unsigned long long>(cnt.msgs));
        const unsigned long long v_events = warp_sum(cnt.events);
        const unsigned long long v_he = warp_sum(cnt.hop_e);
        const unsigned long long v_hw = warp_sum(cnt.hop_w);
        const unsigned long long v_hn = warp_sum(cnt.hop_n);
        const unsigned long long v_hs = warp_sum(cnt.hop_s);
        const double v_se = warp_sum(cnt.syn_e);
        const double v_de = warp_sum(cnt.den_e);
        const double v_pr = warp_sum(cnt.proc);
        double v_dup = 0.0;
        if constexpr (V != kStreamQ4) v_dup = warp_sum(cnt.dup_e);
        if (lane == 0)
        {
            part_d[warp][3] = v_dup;
            part_l[warp][0] = v_msgs;
            part_l[warp][1] = v_events;
            part_l[warp][2] = v_he;
            part_l[warp][3] = v_hw;
            part_l[warp][4] = v_hn;
            part_l[warp][5] = v_hs;
            part_d[warp][0] = v_se;
            part_d[warp][1] = v_de;
            part_d[warp][2] = v_pr;
        }
        __syncthreads();
        if (threadIdx.x == 0)
        {
            StatsM out = {0u, 0u, 0ull, 0ull, 0ull, 0ull, 0ull, 0.0, 0.0, 0.0, 0.0};
            for (int w = 0; w < kFanoutWarps; ++w)
            {
                out.msgs += static_cast<uint32_t>(part_l[w][0]);
                out.events += part_l[w][1];
                out.hop_e += part_l[w][2];
                out.hop_w += part_l[w][3];
                out.hop_n += part_l[w][4];
                out.hop_s += part_l[w][5];
                out.syn_e += part_d[w][0];
                out.den_e += part_d[w][1];
                out.proc += part_d[w][2];
                out.dup_e += part_d[w][3];
            }
            s.stats_m[item_id] = out;
        }
    }
    stamp(4u + 4u * tl_item);
    if constexpr (kFused)
    {
        // ---- core completion: the CTA that retires the last work item of a core folds the core's step statistics
        // and runs the core's neuron phase of the NEXT step; the CTA that completes the last core publishes that
        // step's raster (ready flag, or the peer-memory exchange) and folds the chip
        if (threadIdx.x == 0)
        {
            __threadfence(); // this item's statistics and merged charge before the count
            const uint32_t done = atomicAdd(&s.core_done[ci], 1u) + 1u;
            core_event = done == core.item_count ? 1u : 0u;
            if (core_event != 0u) s.core_done[ci] = 0u; // ready for the next step
        }
        __syncthreads();
        if (core_event != 0u)
        {
            __threadfence();
            const uint32_t parity = static_cast<uint32_t>(sv.step_seq & 1ull);
            if (warp == 0)
            {
                const StepPartial p = fold_core(t, s, ci, lane, parity);
                if (lane == 0) s.core_partials[static_cast<size_t>(parity) * t.n_active_cores + core.active_idx] = p;
            }
            if (sv.fuse_next != 0u && core.seg_count > 0u && core.fast_soma != 0u)
            {
                __syncthreads(); // (warp 0 is done with the fold; the whole dynamic shared memory is idle)
                soma_core(t, s, core, smem_raw, sv.steps_done + 1, parity ^ 1u);
            }
            else if (sv.fuse_next != 0u && core.seg_count > 0u)
            {
                // the class table goes into the (now idle) list region, the reduction scratch behind it
                sfe_soma_class *class_cache = reinterpret_cast<sfe_soma_class *>(smem_raw + list_off);
                SomaScratch *scr = reinterpret_cast<SomaScratch *>(smem_raw + list_off + kClassCache * sizeof(sfe_soma_class));
                const bool classes_cached = t.n_soma_classes <= kClassCache;
                if (classes_cached)
                {
                    const unsigned long long *src = reinterpret_cast<const unsigned long long *>(t.classes);
                    unsigned long long *dst = reinterpret_cast<unsigned long long *>(class_cache);
                    for (uint32_t x = threadIdx.x; x < t.n_soma_classes * (sizeof(sfe_soma_class) / 8); x += kFanoutThreads) dst[x] = __ldg(src + x);
                }
                __syncthreads();
                for (uint32_t g = 0; g < core.seg_count; ++g)
                    soma_segment<false>(t, s, sv, core.seg_begin + g, classes_cached ? class_cache : nullptr, *scr, sv.steps_done + 1,
                            parity ^ 1u);
            }
            __syncthreads();
            if (threadIdx.x == 0)
            {
                __threadfence(); // the core's partial, state, raster words and inbox bits before the count
                const uint32_t finished = atomicAdd(s.cores_finished, 1u) + 1u;
                core_event = finished == t.n_active_cores ? 2u : 0u;
                if (core_event != 0u) *s.cores_finished = 0u;
            }
            __syncthreads();
            if (core_event == 2u)
            {
                __threadfence();
                if (sv.fuse_next != 0u)
                {
                    if (threadIdx.x == 0) ready_publish(s.ready, sv.step_seq + 1ull);
                }
                stamp(14);
                if (warp == 0) fold_chip(t, s, sv, lane);
            }
        }
    }
    stamp(5u + 4u * tl_item);
    if (tl_item < 2u) ++tl_item;
    } // persistent loop
    stamp(15);
    // completion of this launch implies completion of the one before it (whose last CTA may still be folding when
    // this one's work is done): stream order as seen by the host and by the next non-fused launch stays transitive
    if constexpr (kFused) griddep_wait();
}

template <int V, bool kFused>
__global__ void __launch_bounds__(kFanoutThreads, (V == kStreamQ4 && !kFused) ? kQ4CtasPerSm : 3) fanout_kernel(const DevTables t, const DevState s, const StepVars sv, const uint32_t tma_off)
{
    fanout_body<V, kFused>(t, s, sv, tma_off, blockIdx.x);
}

// batched launch (see soma_kernel_batch): 3 CTAs per SM - the tables are read through a pointer instead of the
// constant bank, and the chips of a batch are small and latency-bound, so registers matter more than occupancy
template <int V>
__global__ void __launch_bounds__(kFanoutThreads, 3) fanout_kernel_batch(const BatchSlot *__restrict__ slots, const uint2 *__restrict__ cta_map, const BatchStep a)
{
    __shared__ __align__(16) unsigned long long window[sizeof(BatchSlot) / 8];
    const uint2 where = __ldg(cta_map + blockIdx.x); // (chip, CTA of the chip)
    const BatchSlot &b = load_slot(slots, where.x, window);
    const StepVars sv = batch_vars(b, a);
    fanout_body<V, false>(b.t, b.s, sv, b.tma_off, where.y);
}

// Stand-alone finalize (engines without message-phase work items): one WARP per active core.
__global__ void __launch_bounds__(kFinalThreads) finalize_kernel(const DevTables t, const DevState s, const StepVars sv)
{
    __shared__ __align__(16) unsigned char scratch[kFinalScratchBytes];
    griddep_launch_dependents();
    griddep_wait();
    finalize_body(t, s, sv, blockIdx.x, gridDim.x, scratch);
}

// ---------------------------------------------------------------------------
// Bulk synthetic network: synapse generation + exactness certificate on device
// ---------------------------------------------------------------------------
// Device synapse layout: a segment (the synapses of one axon-in) is padded to a multiple of 4 records, and a segment
// of 64 or more records starts on a multiple of 16 records: its 4-byte records then begin on a 64-byte boundary,
// the granularity at which HBM is read (a 500-byte segment at an arbitrary 16-byte offset touches 8.6 such blocks on
// average, an aligned one 8).
__host__ __device__ __forceinline__ unsigned long long segment_start(const unsigned long long off, const uint32_t count)
{
    return count >= 64u ? (off + 15ull) & ~15ull : off;
}
__host__ __device__ __forceinline__ uint32_t segment_stride(const uint32_t count) // of equal segments laid end to end
{
    const uint32_t padded = (count + 3u) & ~3u;
    return count >= 64u ? (padded + 15u) & ~15u : padded;
}

__global__ void synth_generate_kernel(const sfe_synth_spec sp, double *__restrict__ syn_w,
        uint32_t *__restrict__ syn_meta, const unsigned long long total, const uint32_t first_core)
{
    const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
    const uint32_t C = sp.cores, P = sp.neurons_per_core, D = sp.dest_cores, S = sp.syn_per_axon;
    const uint32_t Sp = segment_stride(S); // device layout of equal segments
    const unsigned long long per_core = (static_cast<unsigned long long>(P) * D * Sp + 15ull) & ~15ull; // cores start 64-byte aligned
    for (unsigned long long g = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; g < total;
            g += stride)
    {
        const uint32_t d = first_core + static_cast<uint32_t>(g / per_core);
        const unsigned long long rem = g % per_core;
        const uint32_t slot = static_cast<uint32_t>(rem / Sp);
        const uint32_t j = static_cast<uint32_t>(rem % Sp);
        if (j >= S || slot >= P * D)
        {
            syn_w[g] = 0.0;
            syn_meta[g] = 0u;
            continue;
        }
        const uint32_t r = slot / P, p = slot % P;
        // r-th source core of destination d, ascending core id (see lower_synthetic)
        uint32_t src;
        if (D == C) src = r;
        else if (d + 1 >= D) src = d + 1 - D + r;
        else
        {
            const uint32_t wrapped = D - 1 - d;
            src = r <= d ? r : C - wrapped + (r - (d + 1));
        }
        const uint32_t k = (d + C - src) % C;
        const unsigned long long n = static_cast<unsigned long long>(src) * P + p;
        const unsigned long long axon = n * D + k;
        const unsigned long long syn = axon * S + j;
        const sfe_synth_axon ap = sfe_synth_axon_params(&sp, axon);
        syn_w[g] = static_cast<double>(sfe_synth_weight(&sp, syn));
        syn_meta[g] = sfe_synth_post(&sp, &ap, j) | (sfe_synth_delay(&sp, syn) << 16);
    }
}

// per core: worst per-post fan-in and sum |w * 2^shift| -> accumulation mode
__global__ void __launch_bounds__(256) certify_kernel(
        CoreDev *cores, const uint32_t *core_list, const sfe_axon_in *axons, const double *syn_w, const uint32_t *syn_meta)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint32_t bad;
    const uint32_t ci = core_list[blockIdx.x];
    CoreDev &core = cores[ci];
    const uint32_t P = core.neuron_count;
    unsigned long long *sum_abs = reinterpret_cast<unsigned long long *>(smem_raw);
    uint32_t *fan = reinterpret_cast<uint32_t *>(sum_abs + P);
    for (uint32_t x = threadIdx.x; x < P; x += blockDim.x)
    {
        sum_abs[x] = 0ull;
        fan[x] = 0u;
    }
    if (threadIdx.x == 0) bad = 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (uint32_t a = warp; a < core.axon_count; a += nwarps)
    {
        const sfe_axon_in ax = axons[core.axon_begin + a];
        for (uint32_t j = lane; j < ax.syn_count; j += 32)
        {
            const double scaled = fabs(syn_w[core.syn_begin + ax.syn_off + j] * core.scale);
            if (!(scaled < 2147483648.0) || scaled != rint(scaled)) bad = 1u;
            const uint32_t post = SFE_SYN_POST(syn_meta[core.syn_begin + ax.syn_off + j]);
            atomicAdd(&sum_abs[post], static_cast<unsigned long long>(scaled));
            atomicAdd(&fan[post], 1u);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        unsigned long long worst_sum = 0ull;
        uint32_t worst_fan = 0u;
        for (uint32_t x = 0; x < P; ++x)
        {
            worst_sum = sum_abs[x] > worst_sum ? sum_abs[x] : worst_sum;
            worst_fan = fan[x] > worst_fan ? fan[x] : worst_fan;
        }
        uint32_t mode = SFE_ACC_ORDERED;
        if (bad == 0u)
        {
            if (worst_sum < 65536ull && worst_fan < 32768u) mode = SFE_ACC_PACKED17;
            else if (worst_sum < 524288ull && worst_fan < 4096u) mode = SFE_ACC_PACKED32;
            else if (worst_sum < 2147483648ull) mode = SFE_ACC_DUAL32;
        }
        core.acc_mode = mode;
    }
}

// Where record j of a segment of `count` records sits in the 4-byte table. A segment is cut into
// chunks of 128 records; a chunk of n records (n4 = n rounded up to 4) is stored lane-major:
// with W = n4 / 4, record u*W + l of the chunk is at position 4*l + u. The lane that moves the
// l-th 16-byte piece of the chunk into shared memory then owns four records whose indices are W
// apart: it reads them back with one 16-byte load, needs no warp barrier, and for a fixed u the
// lanes touch consecutive records (bank-friendly for strided post-synaptic indices).
__host__ __device__ __forceinline__ uint32_t q4_position(const uint32_t j, const uint32_t count)
{
    const uint32_t chunk = j & ~127u, n = min(count - chunk, 128u), w = ((n + 3u) & ~3u) >> 2, r = j - chunk;
    return chunk + 4u * (r % w) + r / w;
}

// The 4-byte record of a synapse of a PACKED17 core: bits 14..31 = weight * 2^shift + 2^17 (the certificate bounds
// |weight * 2^shift| below 2^16, so the field is never 0), bits 2..13 = post-synaptic neuron * 4 (its byte offset in
// the core's shared-memory accumulators), bits 0..1 = 0. The message phase needs one AND for the address and one
// shift for the addend (sum + count * 2^17 in one atomic). An all-zero word is a pad record: it adds 0 to cell 0.
__host__ __device__ __forceinline__ uint32_t q4_record(const int fixed, const uint32_t post)
{
    return (static_cast<uint32_t>(fixed + (1 << 17)) << 14) | ((post & 0xFFFu) << 2);
}

// 4-byte records of the certified cores (one CTA column per core, warps stride over its axons)
__global__ void __launch_bounds__(256) pack_q4_kernel(const CoreDev *cores, const uint32_t *core_list, const sfe_axon_in *axons,
        const double *syn_w, const uint32_t *syn_meta, uint32_t *syn_q4)
{
    const CoreDev &core = cores[core_list[blockIdx.y]];
    if (core.q4 == 0u) return;
    const int lane = threadIdx.x & 31;
    const uint32_t warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nwarps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t a = warp; a < core.axon_count; a += nwarps)
    {
        const sfe_axon_in ax = axons[core.axon_begin + a];
        for (uint32_t j = lane; j < ax.syn_count; j += 32)
        {
            const size_t at = core.syn_begin + ax.syn_off + j;
            const int fixed = __double2int_rn(syn_w[at] * core.scale);
            // lane-major inside every 128-record chunk: the lane that copies 16 bytes of a chunk
            // owns records l, l+W, l+2W, l+3W of it (W = quarter width), see q4_position
            syn_q4[core.syn_begin + ax.syn_off + q4_position(j, ax.syn_count)] = q4_record(fixed, SFE_SYN_POST(syn_meta[at]));
        }
    }
}

} // namespace

// ===========================================================================
// Host side of the engine
// ===========================================================================
#define SFE_CUDA(call)                                                                              \
    do                                                                                              \
    {                                                                                               \
        const cudaError_t err__ = (call);                                                           \
        if (err__ != cudaSuccess)                                                                   \
        {                                                                                           \
            sfe::set_last_error(std::string("CUDA error: ") + cudaGetErrorString(err__) + " at " +  \
                    __FILE__ + ":" + std::to_string(__LINE__) + " (" #call ")");                    \
            return -1;                                                                              \
        }                                                                                           \
    } while (0)

struct sfe_engine
{
    uint2 *p2p_block{nullptr};           // [2][fired_words] (word, epoch + 1) pairs, exported over CUDA IPC
    void *p2p_peer_base[kMaxPeers] = {}; // opened peer blocks (own entry stays null)
    bool p2p_on{false};
    bool q4_any{false};
    int device{0};
    cudaStream_t stream{nullptr};
    bool own_stream{false};
    std::vector<void *> allocs;
    size_t device_bytes{0};
    DevTables t{};
    DevState s{};
    StepVars sv{};
    CoreDev *d_cores{nullptr};
    uint32_t *d_neuron_class{nullptr};
    sfe_soma_class *d_classes{nullptr};
    uint32_t n_classes_cap{0};
    std::vector<CoreDev> h_cores;
    std::vector<uint32_t> soma_list, fanout_list, all_soma_list, active_list;
    bool fused_ok{false}; // the fused step kernel can run this engine's enqueue-only batches
    bool tickets_prepared{false}; // a batch entry point has cleared the ticket counters of the steps it enqueues
    uint32_t *d_work_pool{nullptr}; // kWorkPool ticket counters, one per step in flight (slot = step_seq % kWorkPool)
    std::vector<uint32_t> fired_word_begin; // per core
    uint32_t fired_words{0}, inbox_words{0};
    uint32_t dend_cells{0};
    size_t fanout_smem{0};
    uint32_t tma_off{0};
    int fanout_variant{kStreamTma};
    unsigned fanout_grid{1};
    unsigned final_grid{1};
    uint32_t n_segments{0};
    bool exotic{false};
    bool neurofem{false}; // the chip maps "neurofem" neurons (extra state arrays)
    // multi-GPU partition (contiguous core ranges)
    uint32_t rank{0}, world{1};
    std::vector<uint32_t> owner;      // rank that simulates each core
    uint32_t slice_words{0};          // raster words per rank slice (equal for all ranks)
    // traced runs (sfe_engine_run with fired_bits): every step of a batch writes its raster into its own slot of a
    // device ring, copied to the host once per batch - no copy between the two kernels of a step
    uint32_t *d_fired_ring{nullptr};
    size_t fired_ring_steps{0};
    void *pinned_fired{nullptr};
    size_t pinned_fired_bytes{0};
    bool raster_identity{false}; // the padded raster layout equals device-index bit order (every core starts on a word)
    uint32_t *d_fired_local{nullptr}; // this rank's raster slice (slice_words)
    uint32_t *d_fired_global{nullptr};// world * slice_words
    bool external_exchange{false};
    uint32_t n_neurons{0}, n_probes{0}, n_cores{0}, n_hh{0}, n_u_probes{0};
    int64_t total_timesteps{0};
    int64_t launches{0};
    int64_t log_read{0}; // steps already collected from the device log
    uint32_t log_cap{0};
    std::vector<sfe_core_desc> core_desc; // host copy (neuron ranges)
    std::vector<double> potential0;
    std::vector<sfe_hh_init> hh_init;
    // pinned staging
    void *pinned{nullptr};
    size_t pinned_bytes{0};
    cudaEvent_t ev_begin{nullptr}, ev_end{nullptr};
    // per-launch timing of the message-phase kernel between time_begin/time_end
    bool timing{false};
    bool time_launches{true};
    // The fold of a step (finalize) rides on the next step's neuron-phase kernel; the last step of
    // a batch is folded by the stand-alone kernel when somebody needs the records (flush_fold).
    bool piggyback{true}, pending_fold{false};
    uint32_t pending_parity{0};
    int64_t pending_step{0};
    // whole-vector bias uploads are double-buffered and travel on their own stream, so that the
    // upload for the next step overlaps the kernels of the current one
    double *bias_buf[2] = {nullptr, nullptr};
    double *bias_stage{nullptr}; // pinned staging of sfe_engine_set_bias_staged
    int bias_cur{0}, bias_pending{-1};
    cudaStream_t copy_stream{nullptr};
    cudaEvent_t bias_ready[2] = {nullptr, nullptr}, bias_free[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used{0};
    bool ordered_any{false}, dual_any{false};
    // cooperative cancellation of a long sfe_engine_run (Ctrl-C in the Python binding): checked between batches
    std::atomic<bool> stop_requested{false};
    // device-side Poisson draws
    PoissonUnit *d_poisson_units{nullptr};
    uint32_t *d_poisson_cols{nullptr}, *d_mt{nullptr}, *d_mt_idx{nullptr};
    uint32_t n_poisson_units{0};
    // out-of-tree soma device models
    struct PlugModel
    {
        const sfe_device_model_desc *desc{nullptr};
        uint32_t n{0}, first{0}; // instances, chip-wide index of the first
        double *d_state{nullptr}, *d_params{nullptr};
        std::vector<double> state_init;
    };
    std::vector<PlugModel> plug;
    BatchSlot *d_batch_slots{nullptr}; // device copy of the slots of the last batch this engine led
    uint32_t batch_slots_cap{0};
    uint2 *d_batch_map{nullptr};       // (chip, CTA of the chip) of every CTA of the batched neuron- and message-phase launch
    size_t batch_map_cap{0};
    uint32_t n_plug{0};
    bool plug_failed{false}; // a model's launch function reported an error
    PlugGather *d_plug_gather{nullptr};
    double *d_plug_in{nullptr};
    uint8_t *d_plug_has{nullptr}, *d_plug_status{nullptr};
    uint32_t n_taps_units{0}; // "taps" dendrites
    size_t tap_cells{0};
    // Poisson overlay (host-drawn, see poisson.cpp)
    uint32_t n_poisson_cols{0};
    uint8_t *d_overlay{nullptr};
    size_t overlay_cap{0};
    // TrueNorth threshold jitter overlay
    uint32_t n_rand_cols{0};
    uint32_t *d_rand_overlay{nullptr};
    size_t rand_overlay_cap{0};

    template <typename T> int alloc(T **p, size_t count)
    {
        const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
        void *q = nullptr;
        SFE_CUDA(cudaMalloc(&q, bytes));
        SFE_CUDA(cudaMemsetAsync(q, 0, bytes, stream));
        allocs.push_back(q);
        device_bytes += bytes;
        *p = static_cast<T *>(q);
        return 0;
    }
    template <typename T> int upload(const T **dst, const T *src, size_t count)
    {
        T *p = nullptr;
        if (alloc(&p, count) != 0) return -1;
        if (count > 0 && src != nullptr) SFE_CUDA(cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, stream));
        *dst = p;
        return 0;
    }
    int ensure_pinned(size_t bytes)
    {
        if (bytes <= pinned_bytes) return 0;
        if (pinned != nullptr) cudaFreeHost(pinned);
        pinned = nullptr;
        pinned_bytes = 0;
        SFE_CUDA(cudaMallocHost(&pinned, bytes));
        pinned_bytes = bytes;
        return 0;
    }
};

extern "C" int sfe_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess)
    {
        cudaGetLastError();
        return 0;
    }
    return n;
}

static int engine_init_state(sfe_engine *e)
{
    // initial potentials and HH state (also what reset() does not restore: the
    // reference's reset() zeroes, it does not re-apply initial attributes)
    SFE_CUDA(cudaMemcpyAsync(e->s.v, e->potential0.data(), e->n_neurons * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    if (e->n_hh > 0)
    {
        std::vector<double> hh(5 * static_cast<size_t>(e->n_hh), 0.0);
        for (uint32_t k = 0; k < e->n_hh; ++k)
        {
            hh[e->n_hh + k] = e->hh_init[k].m;
            hh[2 * e->n_hh + k] = e->hh_init[k].n;
            hh[3 * e->n_hh + k] = e->hh_init[k].h;
            hh[4 * e->n_hh + k] = e->hh_init[k].current;
        }
        SFE_CUDA(cudaMemcpyAsync(e->s.hh, hh.data(), hh.size() * sizeof(double), cudaMemcpyHostToDevice, e->stream));
        SFE_CUDA(cudaStreamSynchronize(e->stream));
    }
    return 0;
}

static int engine_build(sfe_engine *e, const sfe_tables *tb)
{
    SFE_CUDA(cudaSetDevice(e->device));
    SFE_CUDA(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    e->own_stream = true;
    SFE_CUDA(cudaEventCreate(&e->ev_begin));
    SFE_CUDA(cudaEventCreate(&e->ev_end));
    e->n_neurons = tb->n_neurons;
    e->n_probes = tb->n_probes;
    e->n_u_probes = tb->n_u_probes;
    e->n_cores = tb->n_cores;
    e->n_hh = tb->n_hh;
    e->core_desc.assign(tb->cores, tb->cores + tb->n_cores);
    for (uint32_t k = 0; k < tb->n_soma_classes; ++k)
        if (tb->soma_classes[k].model == SFE_SOMA_INPUT || tb->soma_classes[k].model == SFE_SOMA_HH ||
                tb->soma_classes[k].model == SFE_SOMA_NEUROFEM || tb->soma_classes[k].model == SFE_SOMA_DEVICE_MODEL ||
                (tb->soma_classes[k].model == SFE_SOMA_TRUENORTH && tb->soma_classes[k].random_mask != 0u) || (tb->soma_classes[k].flags & SFE_SOMA_NOISE) != 0u)
            e->exotic = true;
    for (uint32_t k = 0; k < tb->n_soma_classes; ++k)
        if (tb->soma_classes[k].model == SFE_SOMA_NEUROFEM) e->neurofem = true;
    e->potential0.assign(tb->neuron_potential0, tb->neuron_potential0 + tb->n_neurons);
    if (tb->n_hh > 0) e->hh_init.assign(tb->hh, tb->hh + tb->n_hh);

    // ---- partition + raster layout (host/partition.cpp) --------------------------
    e->owner.assign(tb->n_cores, 0);
    e->fired_word_begin.assign(tb->n_cores, 0);
    if (sfe_plan_partition(tb, e->world, e->owner.data(), e->fired_word_begin.data(), &e->slice_words) != 0) return -1;
    auto is_local = [&](uint32_t c) { return e->owner[c] == e->rank; };

    // ---- per-core device descriptors, padded bit layouts ---------------------
    e->h_cores.resize(tb->n_cores);
    std::vector<uint32_t> inbox_word_begin(tb->n_cores);
    uint32_t inbox_words = 0, dend_cells = 0;
    uint32_t inbox_lo = UINT32_MAX, inbox_hi = 0;
    size_t smem_max = 0;
    for (uint32_t c = 0; c < tb->n_cores; ++c)
    {
        const sfe_core_desc &cd = tb->cores[c];
        CoreDev &d = e->h_cores[c];
        std::memset(&d, 0, sizeof(d));
        d.neuron_begin = cd.neuron_begin;
        d.neuron_count = cd.neuron_count;
        d.axon_begin = cd.axon_in_begin;
        d.axon_count = cd.axon_in_count;
        d.inbox_word_begin = inbox_words;
        d.fired_word_begin = e->fired_word_begin[c];
        d.dend_base = dend_cells;
        d.ring = cd.ring == 0 ? 1 : cd.ring;
        d.acc_mode = cd.acc_mode;
        d.dend_in_msg = cd.dend_in_msg;
        d.fixed_slots = cd.fixed_slots;
        d.tile = cd.tile;
        d.syn_begin = cd.syn_begin;
        d.scale = std::ldexp(1.0, cd.weight_shift);
        d.inv_scale = std::ldexp(1.0, -cd.weight_shift);
        d.lat_axon_in = cd.latency_axon_in;
        d.e_axon_in = cd.energy_axon_in;
        d.lat_axon_out = cd.latency_axon_out;
        d.e_axon_out = cd.energy_axon_out;
        const sfe_tile_desc &tile = tb->tiles[cd.tile];
        d.e_east = tile.energy_east;
        d.e_west = tile.energy_west;
        d.e_south = tile.energy_south;
        d.e_north = tile.energy_north;
        if (cd.neuron_count > 0x10000u)
        {
            sfe::set_last_error("core " + std::to_string(c) + " maps " + std::to_string(cd.neuron_count) +
                    " neurons: the synapse records address at most 65536 neurons per core");
            return -1;
        }
        inbox_word_begin[c] = inbox_words;
        if (is_local(c) && cd.axon_in_count > 0)
        {
            inbox_lo = std::min(inbox_lo, inbox_words);
            inbox_hi = std::max(inbox_hi, inbox_words + (cd.axon_in_count + 31) / 32);
        }
        inbox_words += (cd.axon_in_count + 31) / 32;
        dend_cells += cd.neuron_count * d.ring;
        if (cd.neuron_count > 0) e->all_soma_list.push_back(c);
        if (!is_local(c)) continue;
        if (cd.neuron_count > 0) e->soma_list.push_back(c);
        if (cd.axon_in_count > 0) e->fanout_list.push_back(c);
        if (cd.acc_mode == SFE_ACC_ORDERED) e->ordered_any = true;
        if (cd.acc_mode == SFE_ACC_DUAL32) e->dual_any = true;
    }
    e->inbox_words = inbox_words;
    e->fired_words = e->world * e->slice_words;
    e->dend_cells = dend_cells;
    e->t.partitioned = e->world > 1 ? 1u : 0u;
    e->t.inbox_lo = inbox_lo == UINT32_MAX ? 0u : inbox_lo;
    e->t.inbox_hi = inbox_hi;

    // ---- static tables -----------------------------------------------------------
    if (e->upload(&e->t.classes, tb->soma_classes, tb->n_soma_classes) != 0) return -1;
    e->d_classes = const_cast<sfe_soma_class *>(e->t.classes);
    e->n_classes_cap = tb->n_soma_classes;
    if (e->upload(&e->t.costs, tb->cost_classes, tb->n_cost_classes) != 0) return -1;
    if (e->upload(&e->t.neuron_class, tb->neuron_class, tb->n_neurons) != 0) return -1;
    e->d_neuron_class = const_cast<uint32_t *>(e->t.neuron_class);
    if (e->upload(&e->t.neuron_aux, tb->neuron_aux, tb->n_neurons) != 0) return -1;
    if (e->upload(&e->t.axon_out_begin, tb->axon_out_begin, static_cast<size_t>(tb->n_neurons) + 1) != 0) return -1;
    {
        // axon-out targets as padded inbox bit positions
        std::vector<uint32_t> axon_core(tb->n_axons_in);
        for (uint32_t c = 0; c < tb->n_cores; ++c)
            for (uint32_t a = 0; a < tb->cores[c].axon_in_count; ++a) axon_core[tb->cores[c].axon_in_begin + a] = c;
        std::vector<uint32_t> bits(tb->n_axons_out);
        for (uint64_t a = 0; a < tb->n_axons_out; ++a)
        {
            const uint32_t id = tb->axon_out_target[a];
            const uint32_t c = axon_core[id];
            bits[a] = inbox_word_begin[c] * 32u + (id - tb->cores[c].axon_in_begin);
        }
        if (e->upload(&e->t.axon_out_bit, bits.data(), bits.size()) != 0) return -1;
        if (e->world > 1)
        {
            // source neuron (as a raster bit) of every axon-in of this rank's inbox range
            const uint64_t lo = static_cast<uint64_t>(e->t.inbox_lo) * 32u, hi = static_cast<uint64_t>(e->t.inbox_hi) * 32u;
            std::vector<uint32_t> src(std::max<uint64_t>(hi > lo ? hi - lo : 0, 32), 0xFFFFFFFFu);
            for (uint32_t c = 0; c < tb->n_cores; ++c)
                for (uint32_t k = 0; k < tb->cores[c].neuron_count; ++k)
                {
                    const uint32_t i = tb->cores[c].neuron_begin + k;
                    for (uint32_t a = tb->axon_out_begin[i]; a < tb->axon_out_begin[i + 1]; ++a)
                        if (bits[a] >= lo && bits[a] < hi) src[bits[a] - lo] = e->fired_word_begin[c] * 32u + k;
                }
            if (e->upload(&e->t.axon_src, src.data(), src.size()) != 0) return -1;
            // per inbox word: copied raster word, padding only, or gathered bit by bit (see inbox_word_from_raster)
            std::vector<uint32_t> word_src(src.size() / 32u, kWordGather);
            for (size_t w = 0; w < word_src.size(); ++w)
            {
                const uint32_t *a = src.data() + 32u * w;
                bool empty = true, copied = a[0] != 0xFFFFFFFFu && (a[0] & 31u) == 0u;
                for (uint32_t j = 0; j < 32u; ++j)
                {
                    empty = empty && a[j] == 0xFFFFFFFFu;
                    copied = copied && a[j] == a[0] + j;
                }
                if (empty) word_src[w] = kWordEmpty;
                else if (copied) word_src[w] = a[0] >> 5;
            }
            if (e->upload(&e->t.word_src, word_src.data(), word_src.size()) != 0) return -1;
        }
        SFE_CUDA(cudaStreamSynchronize(e->stream));
    }
    if (e->upload(&e->t.inputs, tb->inputs, tb->n_inputs) != 0) return -1;
    if (e->upload(&e->t.input_spikes, tb->input_spikes, tb->n_input_spikes) != 0) return -1;
    if (tb->n_taps_units != 0)
    {
        if (e->world > 1 || tb->syn_weight == nullptr || tb->syn_meta == nullptr)
        {
            sfe::set_last_error("'taps' dendrites need an unpartitioned chip with host synapse tables");
            return -1;
        }
        // raster bit of every neuron, and every line's incoming synapses in arrival order (axons-in of its core
        // in order, synapses of an axon in order: exactly the order the message phase of the reference visits)
        std::vector<uint32_t> raster_bit(tb->n_neurons);
        for (uint32_t c = 0; c < tb->n_cores; ++c)
            for (uint32_t k = 0; k < tb->cores[c].neuron_count; ++k)
                raster_bit[tb->cores[c].neuron_begin + k] = e->fired_word_begin[c] * 32u + k;
        std::vector<std::vector<TapsSyn>> incoming(tb->n_taps_units);
        for (uint32_t c = 0; c < tb->n_cores; ++c)
        {
            const sfe_core_desc &cd = tb->cores[c];
            for (uint32_t a = 0; a < cd.axon_in_count; ++a)
            {
                const sfe_axon_in &ax = tb->axons_in[cd.axon_in_begin + a];
                const uint32_t src = tb->axon_src[cd.axon_in_begin + a];
                for (uint32_t j = 0; j < ax.syn_count; ++j)
                {
                    const uint64_t at = cd.syn_begin + ax.syn_off + j;
                    const uint32_t post = cd.neuron_begin + SFE_SYN_POST(tb->syn_meta[at]);
                    const uint32_t unit = tb->neuron_taps[post];
                    if (unit == 0xFFFFFFFFu) continue;
                    incoming[unit].push_back({tb->syn_weight[at], raster_bit[src], SFE_SYN_TAP(tb->syn_meta[at]), post, 0u});
                }
            }
        }
        std::vector<TapsUnit> units(tb->n_taps_units);
        std::vector<TapsSyn> flat;
        uint32_t cells = 0;
        for (uint32_t u = 0; u < tb->n_taps_units; ++u)
        {
            units[u] = {tb->taps[u].n_taps, tb->taps[u].const_off, cells, static_cast<uint32_t>(flat.size()),
                    static_cast<uint32_t>(incoming[u].size()), 0u};
            cells += tb->taps[u].n_taps;
            for (const TapsSyn &y : incoming[u])
            {
                if (y.tap >= tb->taps[u].n_taps)
                {
                    sfe::set_last_error("sfe_engine_create: synapse addressed to a tap outside its line");
                    return -1;
                }
                flat.push_back(y);
            }
        }
        if (e->upload(&e->t.taps_units, units.data(), units.size()) != 0) return -1;
        if (e->upload(&e->t.taps_syn, flat.data(), flat.size()) != 0) return -1;
        if (e->upload(&e->t.taps_values, tb->taps_values, tb->n_taps_values) != 0) return -1;
        if (e->upload(&e->t.neuron_taps, tb->neuron_taps, tb->n_neurons) != 0) return -1;
        e->t.n_taps_units = tb->n_taps_units;
        e->n_taps_units = tb->n_taps_units;
        e->tap_cells = cells;
        if (e->alloc(&e->s.tap_v, cells) != 0) return -1;
        if (e->alloc(&e->s.tap_next, cells) != 0) return -1;
        if (e->alloc(&e->s.tap_steps, tb->n_taps_units) != 0) return -1;
        if (e->alloc(&e->s.tap_buf, tb->n_neurons) != 0) return -1;
        if (e->alloc(&e->s.tap_has, tb->n_neurons) != 0) return -1;
        e->exotic = true; // the neuron phase reads the lines' outputs in its `exotic` instantiation
    }
    e->n_poisson_cols = tb->n_poisson_cols;
    e->n_rand_cols = tb->n_rand_cols;
    for (uint32_t k = 0; k < tb->n_inputs; ++k)
        if (tb->inputs[k].poisson > 0.0 && tb->inputs[k].poisson_col >= tb->n_poisson_cols)
        {
            sfe::set_last_error("sfe_engine_create: Poisson input neuron with poisson_col >= n_poisson_cols");
            return -1;
        }
    if (tb->n_poisson_cols > 0)
    {
        // generators for sfe_engine_fill_input_overlay: one per input unit with poisson > 0, seeded like the
        // reference's (input_seed_base + unit + 1), states interleaved word-major
        std::vector<PoissonUnit> units;
        std::vector<uint32_t> cols, seeds;
        std::vector<uint32_t> unit_of; // unit ordinal of units[k]
        for (uint32_t k = 0; k < tb->n_inputs; ++k)
        {
            const sfe_input_desc &d = tb->inputs[k];
            if (!(d.poisson > 0.0)) continue;
            size_t slot = 0;
            while (slot < unit_of.size() && unit_of[slot] != d.unit) ++slot;
            if (slot == unit_of.size())
            {
                unit_of.push_back(d.unit);
                seeds.push_back(tb->input_seed_base + d.unit + 1u);
                units.push_back({d.poisson, d.share_count, static_cast<uint32_t>(cols.size())});
                cols.insert(cols.end(), d.share_count, 0xFFFFFFFFu);
            }
            if (d.share_rank < units[slot].share_count) cols[units[slot].col_begin + d.share_rank] = d.poisson_col;
        }
        const uint32_t n_units = static_cast<uint32_t>(units.size());
        std::vector<uint32_t> mt(static_cast<size_t>(sfe::kMtWords) * n_units), idx(n_units);
        for (uint32_t u = 0; u < n_units; ++u) sfe::mt_seed(mt.data() + u, n_units, &idx[u], seeds[u]);
        e->n_poisson_units = n_units;
        if (e->alloc(&e->d_poisson_units, n_units) != 0 || e->alloc(&e->d_poisson_cols, cols.size()) != 0 ||
                e->alloc(&e->d_mt, mt.size()) != 0 || e->alloc(&e->d_mt_idx, n_units) != 0)
            return -1;
        SFE_CUDA(cudaMemcpyAsync(e->d_poisson_units, units.data(), n_units * sizeof(PoissonUnit), cudaMemcpyHostToDevice, e->stream));
        SFE_CUDA(cudaMemcpyAsync(e->d_poisson_cols, cols.data(), cols.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, e->stream));
        SFE_CUDA(cudaMemcpyAsync(e->d_mt, mt.data(), mt.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, e->stream));
        SFE_CUDA(cudaMemcpyAsync(e->d_mt_idx, idx.data(), n_units * sizeof(uint32_t), cudaMemcpyHostToDevice, e->stream));
        SFE_CUDA(cudaStreamSynchronize(e->stream)); // the host vectors go out of scope
    }
    // Device synapse layout: every axon segment starts on a multiple of 4 synapses
    // (32-byte aligned weights, 16-byte aligned meta words) so the message phase can
    // use 128-bit loads; the axon records carry the padded offsets.
    std::vector<sfe_axon_in> dev_axons(tb->axons_in, tb->axons_in + tb->n_axons_in);
    uint64_t padded_total = 0;
    for (uint32_t c = 0; c < tb->n_cores; ++c)
    {
        const sfe_core_desc &cd = tb->cores[c];
        e->h_cores[c].syn_begin = padded_total;
        if (!is_local(c)) continue; // another rank holds this core's synapses
        uint64_t off = 0;
        for (uint32_t a = 0; a < cd.axon_in_count; ++a)
        {
            sfe_axon_in &ax = dev_axons[cd.axon_in_begin + a];
            if (off > 0xffffffffull)
            {
                sfe::set_last_error("more than 2^32 padded synapses on one core");
                return -1;
            }
            off = segment_start(off, ax.syn_count);
            ax.syn_off = static_cast<uint32_t>(off);
            off += (static_cast<uint64_t>(ax.syn_count) + 3u) & ~3ull;
        }
        padded_total += (off + 15ull) & ~15ull; // every core starts on a 64-byte boundary of the 4-byte table
    }
    if (e->upload(&e->t.axons_in, dev_axons.data(), dev_axons.size()) != 0) return -1;
    SFE_CUDA(cudaStreamSynchronize(e->stream));
    if (e->upload(&e->t.probes, tb->probes, tb->n_probes) != 0) return -1;
    if (e->upload(&e->t.u_probes, tb->u_probes, tb->n_u_probes) != 0) return -1;
    if (e->upload(&e->t.noise, tb->noise, tb->n_noise) != 0) return -1;
    if (e->upload(&e->t.noise_values, tb->noise_values, tb->n_noise_values) != 0) return -1;
    for (uint32_t k = 0; k < tb->n_noise; ++k)
        if (tb->noise[k].len == 0 || tb->noise[k].share_count == 0 ||
                static_cast<uint64_t>(tb->noise[k].off) + tb->noise[k].len > tb->n_noise_values)
        {
            sfe::set_last_error("sfe_engine_create: malformed sfe_noise_desc");
            return -1;
        }
    // heaviest destination cores first (longest-processing-time order for the ticket queue)
    std::stable_sort(e->fanout_list.begin(), e->fanout_list.end(),
            [&](uint32_t a, uint32_t b) { return tb->cores[a].syn_count > tb->cores[b].syn_count; });

    if (e->upload(&e->t.fanout_core_list, e->fanout_list.data(), e->fanout_list.size()) != 0) return -1;
    {
        // this rank's cores that do anything in a step (ascending id)
        std::vector<uint32_t> active;
        for (uint32_t c = 0; c < tb->n_cores; ++c)
        {
            if (!is_local(c) || (tb->cores[c].neuron_count == 0 && tb->cores[c].axon_in_count == 0)) continue;
            e->h_cores[c].active_idx = static_cast<uint32_t>(active.size());
            active.push_back(c);
        }
        e->t.n_active_cores = static_cast<uint32_t>(active.size());
        e->active_list = active;
        if (e->upload(&e->t.active_core_list, active.data(), active.size()) != 0) return -1;
        SFE_CUDA(cudaStreamSynchronize(e->stream));
    }
    {
        const CoreDev *p = nullptr;
        if (e->upload(&p, e->h_cores.data(), e->h_cores.size()) != 0) return -1;
        e->t.cores = p;
        e->d_cores = const_cast<CoreDev *>(p);
    }
    e->t.n_cores = tb->n_cores;
    e->t.n_probes = tb->n_probes;
    e->t.n_u_probes = tb->n_u_probes;
    e->t.n_neurons = tb->n_neurons;
    e->t.n_cost_classes = tb->n_cost_classes;
    e->t.n_soma_classes = tb->n_soma_classes;
    e->t.n_fanout_cores = static_cast<uint32_t>(e->fanout_list.size());
    e->t.sync_delay = tb->sync_delay;

    // ---- synapses: copied, or generated on the device ---------------------------
    double *d_w = nullptr;
    uint32_t *d_m = nullptr;
    if (e->alloc(&d_w, padded_total) != 0) return -1;
    if (e->alloc(&d_m, padded_total) != 0) return -1;
    e->t.syn_w = d_w;
    e->t.syn_meta = d_m;
    if (tb->syn_weight != nullptr && tb->n_synapses > 0)
    {
        // repack into the padded layout, one core at a time (bounded staging)
        std::vector<double> pw;
        std::vector<uint32_t> pm;
        for (uint32_t c = 0; c < tb->n_cores; ++c)
        {
            const sfe_core_desc &cd = tb->cores[c];
            if (cd.axon_in_count == 0 || !is_local(c)) continue;
            const sfe_axon_in &last = dev_axons[cd.axon_in_begin + cd.axon_in_count - 1];
            const uint64_t core_padded = last.syn_off + ((static_cast<uint64_t>(last.syn_count) + 3u) & ~3ull);
            pw.assign(core_padded, 0.0);
            pm.assign(core_padded, 0u);
            for (uint32_t a = 0; a < cd.axon_in_count; ++a)
            {
                const sfe_axon_in &src = tb->axons_in[cd.axon_in_begin + a];
                const sfe_axon_in &dst = dev_axons[cd.axon_in_begin + a];
                std::memcpy(pw.data() + dst.syn_off, tb->syn_weight + cd.syn_begin + src.syn_off, src.syn_count * sizeof(double));
                std::memcpy(pm.data() + dst.syn_off, tb->syn_meta + cd.syn_begin + src.syn_off, src.syn_count * sizeof(uint32_t));
            }
            SFE_CUDA(cudaMemcpyAsync(d_w + e->h_cores[c].syn_begin, pw.data(), core_padded * sizeof(double), cudaMemcpyHostToDevice, e->stream));
            SFE_CUDA(cudaMemcpyAsync(d_m + e->h_cores[c].syn_begin, pm.data(), core_padded * sizeof(uint32_t), cudaMemcpyHostToDevice, e->stream));
            SFE_CUDA(cudaStreamSynchronize(e->stream));
        }
    }
    else if (tb->n_synapses > 0 && padded_total > 0)
    {
        if (tb->synth == nullptr)
        {
            sfe::set_last_error("tables carry neither synapse arrays nor a synthetic spec");
            return -1;
        }
        const uint32_t first_local = e->fanout_list.empty() ? 0u : *std::min_element(e->fanout_list.begin(), e->fanout_list.end());
        synth_generate_kernel<<<148 * 16, 256, 0, e->stream>>>(*tb->synth, d_w, d_m, padded_total, first_local);
        SFE_CUDA(cudaGetLastError());
        // certificate computed where the synapses are
        uint32_t max_p = 0;
        for (uint32_t c : e->fanout_list) max_p = std::max(max_p, tb->cores[c].neuron_count);
        if (!e->fanout_list.empty())
        {
            certify_kernel<<<static_cast<unsigned>(e->fanout_list.size()), 256, max_p * 12, e->stream>>>(
                    e->d_cores, e->t.fanout_core_list, e->t.axons_in, d_w, d_m);
            SFE_CUDA(cudaGetLastError());
        }
        SFE_CUDA(cudaMemcpyAsync(e->h_cores.data(), e->d_cores, e->h_cores.size() * sizeof(CoreDev), cudaMemcpyDeviceToHost, e->stream));
        SFE_CUDA(cudaStreamSynchronize(e->stream));
        e->ordered_any = e->dual_any = false;
        for (uint32_t c : e->fanout_list)
        {
            if (e->h_cores[c].acc_mode == SFE_ACC_ORDERED) e->ordered_any = true;
            if (e->h_cores[c].acc_mode == SFE_ACC_DUAL32) e->dual_any = true;
        }
    }

    // SFE_FORCE_ORDERED=1 (measurement knob): every core accumulates in the ordered fp64 mode - the path a network with
    // non-dyadic weights takes - whatever its certificate allows. Results do not change (ordered adds are always right).
    if (const char *v = std::getenv("SFE_FORCE_ORDERED"))
        if (std::atoi(v) != 0)
        {
            for (uint32_t c : e->fanout_list) e->h_cores[c].acc_mode = SFE_ACC_ORDERED;
            e->ordered_any = !e->fanout_list.empty();
            SFE_CUDA(cudaMemcpyAsync(e->d_cores, e->h_cores.data(), e->h_cores.size() * sizeof(CoreDev), cudaMemcpyHostToDevice, e->stream));
            SFE_CUDA(cudaStreamSynchronize(e->stream));
        }

    // ---- compact 4-byte synapse records for the cores whose certificate allows them -------
    {
        const char *env = std::getenv("SFE_SYN_Q4");
        const bool want = env == nullptr || std::atoi(env) != 0;
        bool any = false;
        for (uint32_t c : e->fanout_list)
        {
            CoreDev &d = e->h_cores[c];
            d.q4 = (want && d.acc_mode == SFE_ACC_PACKED17 && d.ring == 1 && d.neuron_count <= 4096 && d.dend_in_msg != 0) ? 1u : 0u;
            any = any || d.q4 != 0u;
        }
        e->q4_any = any;
        if (any)
        {
            uint32_t *d_q = nullptr;
            if (e->alloc(&d_q, padded_total) != 0) return -1;
            e->t.syn_q4 = d_q;
            SFE_CUDA(cudaMemcpyAsync(e->d_cores, e->h_cores.data(), e->h_cores.size() * sizeof(CoreDev), cudaMemcpyHostToDevice, e->stream));
            pack_q4_kernel<<<dim3(16, static_cast<unsigned>(e->fanout_list.size())), 256, 0, e->stream>>>(
                    e->d_cores, e->t.fanout_core_list, e->t.axons_in, d_w, d_m, d_q);
            SFE_CUDA(cudaGetLastError());
            SFE_CUDA(cudaStreamSynchronize(e->stream));
        }
    }

    // ---- neuron-phase segments (after certification: they carry the accumulation mode)
    {
        std::vector<SomaSegment> segs;
        for (uint32_t c : e->soma_list)
        {
            e->h_cores[c].seg_begin = static_cast<uint32_t>(segs.size());
            const CoreDev &d = e->h_cores[c];
            for (uint32_t k0 = 0; k0 < tb->cores[c].neuron_count; k0 += kSomaThreads * kSomaPerThread)
            {
                SomaSegment g;
                g.k0 = k0;
                g.neuron_begin = d.neuron_begin;
                g.neuron_count = d.neuron_count;
                g.fired_word_begin = d.fired_word_begin;
                g.dend_base = d.dend_base;
                g.ring = d.ring;
                g.acc_mode = d.acc_mode;
                g.fixed_slots = d.fixed_slots;
                g.inv_scale = d.inv_scale;
                segs.push_back(g);
            }
            e->h_cores[c].seg_count = static_cast<uint32_t>(segs.size()) - e->h_cores[c].seg_begin;
        }
        e->n_segments = static_cast<uint32_t>(segs.size());
        e->t.n_soma_segments = e->n_segments;
        if (e->upload(&e->t.soma_segments, segs.data(), segs.size()) != 0) return -1;
        SFE_CUDA(cudaMemcpyAsync(e->d_cores, e->h_cores.data(), e->h_cores.size() * sizeof(CoreDev), cudaMemcpyHostToDevice, e->stream));
        SFE_CUDA(cudaStreamSynchronize(e->stream));
    }
    // ---- state ---------------------------------------------------------------------
    if (e->alloc(&e->s.v, tb->n_neurons) != 0) return -1;
    if (e->alloc(&e->s.u, tb->n_neurons) != 0) return -1;
    if (e->alloc(&e->s.bias, tb->n_neurons) != 0) return -1;
    if (e->alloc(&e->s.refractory, tb->n_neurons) != 0) return -1;
    if (e->alloc(&e->s.status, tb->n_neurons) != 0) return -1;
    if (e->alloc(&e->d_fired_global, e->fired_words) != 0) return -1;
    if (e->world > 1)
    {
        if (e->alloc(&e->d_fired_local, e->slice_words) != 0) return -1;
    }
    else e->d_fired_local = e->d_fired_global;
    // the neuron phase indexes the raster with chip-wide word offsets: bias the write
    // base so that this rank's slice lands at the start of its local buffer
    e->s.fired_bits = e->d_fired_local - static_cast<size_t>(e->rank) * e->slice_words * (e->world > 1 ? 1 : 0);
    e->s.fired_global = e->d_fired_global;
    e->raster_identity = e->world == 1;
    for (uint32_t c : e->soma_list)
        if (e->fired_word_begin[c] * 32ull != tb->cores[c].neuron_begin) e->raster_identity = false;
    if (e->alloc(&e->s.inbox, 2 * static_cast<size_t>(inbox_words)) != 0) return -1; // double-buffered by step parity
    e->t.inbox_words = inbox_words;
    if (e->alloc(&e->s.din32, dend_cells) != 0) return -1;
    if (e->alloc(&e->s.dcnt32, (e->ordered_any || e->dual_any) ? dend_cells : 1) != 0) return -1;
    if (e->alloc(&e->s.din64, e->ordered_any ? dend_cells : 1) != 0) return -1;
    if (e->alloc(&e->s.hh, 5 * static_cast<size_t>(tb->n_hh)) != 0) return -1;
    e->s.n_hh = tb->n_hh;
    if (e->alloc(&e->s.nf_u2, e->neurofem ? tb->n_neurons : 1) != 0) return -1;
    if (e->alloc(&e->s.nf_uint, e->neurofem ? tb->n_neurons : 1) != 0) return -1;
    if (e->alloc(&e->s.stats_n, 2 * static_cast<size_t>(e->n_segments)) != 0) return -1; // double-buffered by step parity
    e->log_cap = 4096;
    e->s.log_cap = e->log_cap;
    if (e->alloc(&e->s.log, e->log_cap) != 0) return -1;
    if (e->alloc(&e->s.probe_out, static_cast<size_t>(tb->n_probes) + tb->n_u_probes) != 0) return -1;
    if (e->alloc(&e->s.step, 2) != 0) return -1;
    if (e->alloc(&e->d_work_pool, kWorkPool) != 0) return -1;
    e->sv.work = e->d_work_pool;
    if (e->alloc(&e->s.final_ticket, 1) != 0) return -1;
    if (e->alloc(&e->s.core_done, tb->n_cores) != 0) return -1;
    if (e->alloc(&e->s.cores_finished, 1) != 0) return -1;
    if (e->alloc(&e->s.ready, 1) != 0) return -1;
    if (std::getenv("SFE_TIMELINE") != nullptr)
        if (e->alloc(&e->s.timeline, 2 * kTimelineWords) != 0) return -1; // 64 steps x up to 1024 CTAs x 16 stamps, message + neuron phase
    if (e->alloc(&e->s.core_partials, 2 * static_cast<size_t>(e->t.n_active_cores)) != 0) return -1;
    if (e->alloc(&e->s.x.error, 4) != 0) return -1;
    e->final_grid = std::max<unsigned>(1u, (e->t.n_active_cores + kFinalThreads / 32 - 1) / (kFinalThreads / 32));
    if (e->alloc(&e->s.partials, e->final_grid) != 0) return -1;
    SFE_CUDA(cudaMemcpyAsync(e->s.bias, tb->neuron_bias, tb->n_neurons * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    if (engine_init_state(e) != 0) return -1;

    // ---- out-of-tree soma device models (sfe_device_model.h) -----------------------------------------
    if (tb->n_device_models > 0)
    {
        std::vector<PlugGather> gather(tb->n_device_instances);
        std::vector<const double *> potential(tb->n_device_instances, nullptr);
        uint32_t first = 0;
        for (uint32_t b = 0; b < tb->n_device_models; ++b)
        {
            const sfe_device_model_block &blk = tb->device_models[b];
            if (blk.desc == nullptr || blk.desc->abi_version != SFE_DEVICE_MODEL_ABI || blk.desc->launch == nullptr)
            {
                sfe::set_last_error("sfe_engine_create: malformed device-model block");
                return -1;
            }
            sfe_engine::PlugModel m;
            m.desc = blk.desc;
            m.n = blk.n_instances;
            m.first = first;
            if (e->alloc(&m.d_state, static_cast<size_t>(m.desc->n_state) * m.n) != 0) return -1;
            if (e->alloc(&m.d_params, static_cast<size_t>(m.desc->n_params) * m.n) != 0) return -1;
            m.state_init.assign(blk.state_init, blk.state_init + static_cast<size_t>(m.desc->n_state) * m.n);
            if (!m.state_init.empty())
                SFE_CUDA(cudaMemcpyAsync(m.d_state, m.state_init.data(), m.state_init.size() * sizeof(double), cudaMemcpyHostToDevice, e->stream));
            if (m.desc->n_params > 0 && m.n > 0)
                SFE_CUDA(cudaMemcpyAsync(m.d_params, blk.params, static_cast<size_t>(m.desc->n_params) * m.n * sizeof(double),
                        cudaMemcpyHostToDevice, e->stream));
            for (uint32_t k = 0; k < m.n; ++k)
            {
                const uint32_t i = blk.neurons[k];
                uint32_t c = 0;
                while (c + 1 < tb->n_cores && !(i >= tb->cores[c].neuron_begin && i < tb->cores[c].neuron_begin + tb->cores[c].neuron_count)) ++c;
                const CoreDev &d = e->h_cores[c];
                const sfe_soma_class &cls = tb->soma_classes[tb->neuron_class[i]];
                if (cls.model != SFE_SOMA_DEVICE_MODEL || tb->neuron_aux[i] != first + k)
                {
                    sfe::set_last_error("sfe_engine_create: device-model block does not match the neuron tables");
                    return -1;
                }
                gather[first + k] = {d.dend_base + (i - d.neuron_begin), d.neuron_count, d.ring, d.acc_mode, d.fixed_slots,
                        (cls.dend_in_neuron != 0u && cls.dend_model == SFE_DEND_ACCUMULATOR) ? 1u : 0u, d.inv_scale};
                if (m.desc->potential_state >= 0) potential[first + k] = m.d_state + static_cast<size_t>(m.desc->potential_state) * m.n + k;
            }
            first += m.n;
            e->plug.push_back(std::move(m));
        }
        e->n_plug = first;
        if (e->alloc(&e->d_plug_gather, first) != 0 || e->alloc(&e->d_plug_in, first) != 0 || e->alloc(&e->d_plug_has, first) != 0 ||
                e->alloc(&e->d_plug_status, first) != 0)
            return -1;
        const double **d_pot = nullptr;
        if (e->alloc(&d_pot, first) != 0) return -1;
        SFE_CUDA(cudaMemcpyAsync(e->d_plug_gather, gather.data(), first * sizeof(PlugGather), cudaMemcpyHostToDevice, e->stream));
        SFE_CUDA(cudaMemcpyAsync(d_pot, potential.data(), first * sizeof(double *), cudaMemcpyHostToDevice, e->stream));
        SFE_CUDA(cudaStreamSynchronize(e->stream));
        e->s.plug_status = e->d_plug_status;
        e->s.plug_potential = d_pot;
    }

    // ---- shared memory of the message phase ------------------------------------------
    // A core whose dendrite cells exceed the shared-memory budget accumulates in HBM instead (acc_global): slower,
    // but such a core is rare (an ordered core with delays beyond ~2.8k neurons, an exact one beyond ~24k) and the
    // reference has no such limit. SFE_ACC_SMEM_LIMIT (bytes) lowers the budget (tests).
    size_t acc_budget = 160 * 1024;
    if (const char *v = std::getenv("SFE_ACC_SMEM_LIMIT")) acc_budget = static_cast<size_t>(std::max(0, std::atoi(v)));
    bool any_global = false;
    for (uint32_t c : e->fanout_list)
    {
        CoreDev &d = e->h_cores[c];
        const size_t cells = static_cast<size_t>(d.neuron_count) * d.ring;
        const size_t need = (d.acc_mode == SFE_ACC_PACKED32 || d.acc_mode == SFE_ACC_PACKED17) ? cells * 4 : d.acc_mode == SFE_ACC_DUAL32 ? cells * 8 : cells * 12;
        if (need > acc_budget && d.q4 == 0u)
        {
            d.acc_global = 1u;
            any_global = true;
            continue;
        }
        smem_max = std::max(smem_max, need);
    }
    if (any_global)
    {
        SFE_CUDA(cudaMemcpyAsync(e->d_cores, e->h_cores.data(), e->h_cores.size() * sizeof(CoreDev), cudaMemcpyHostToDevice, e->stream));
        SFE_CUDA(cudaStreamSynchronize(e->stream));
    }
    if (const char *v = std::getenv("SFE_FANOUT"))
    {
        const std::string name(v);
        if (name == "scalar") e->fanout_variant = kStreamScalar;
        else if (name == "tma") e->fanout_variant = kStreamTma;
        else
        {
            sfe::set_last_error("SFE_FANOUT must be scalar or tma");
            return -1;
        }
    }
    {
        // every core certified for 4-byte records: the q4-only instantiation (unless a variant was forced)
        bool all_q4 = !e->fanout_list.empty();
        for (uint32_t c : e->fanout_list) all_q4 = all_q4 && e->h_cores[c].q4 != 0u;
        if (all_q4 && std::getenv("SFE_FANOUT") == nullptr) e->fanout_variant = kStreamQ4;
    }
    e->tma_off = static_cast<uint32_t>((smem_max + 127) & ~static_cast<size_t>(127));
    size_t soma_smem = 0;
    {
        // fused step kernel: which cores' neuron phase runs through soma_core (see there), and the shared memory it needs
        const char *fused = std::getenv("SFE_FUSED_STEP");
        const char *fast = std::getenv("SFE_FAST_SOMA");
        const bool want = (fused != nullptr && std::atoi(fused) != 0) && (fast == nullptr || std::atoi(fast) != 0) && !e->exotic &&
                e->n_taps_units == 0;
        for (uint32_t c : e->soma_list)
        {
            CoreDev &d = e->h_cores[c];
            d.fast_soma = (want && d.neuron_count <= kCoreSomaMax && (d.neuron_count & 3u) == 0u && (d.neuron_begin & 3u) == 0u &&
                                  (d.dend_base & 3u) == 0u && d.acc_mode != SFE_ACC_ORDERED)
                    ? 1u
                    : 0u;
            if (d.fast_soma != 0u) soma_smem = std::max(soma_smem, core_soma_smem(d.neuron_count));
        }
    }
    smem_max = e->tma_off +
            (e->fanout_variant == kStreamTma ? kFanoutWarps * kTmaStages * (kTmaStageBytes + 8)
                                             : (e->q4_any ? kFanoutWarps * kQ4Stages * kQ4StageBytes + kQ4BarBytes : 0)) +
            kListCap * sizeof(uint2);
    smem_max = std::max(smem_max, soma_smem);
    e->fanout_smem = smem_max;
    if (smem_max > 200 * 1024)
    {
        sfe::set_last_error("the message phase needs more than 200 KB of shared memory (SFE_ACC_SMEM_LIMIT too high?)");
        return -1;
    }
    // The opt-in limit is a property of the FUNCTION (per device), not of this engine: several engines
    // live in one process (design-space batches, emulated partitions), so it only ever grows — an engine
    // with a small need must not lower it under one that needs more.
    {
        static std::mutex mu;
        static int raised[64] = {};
        const std::lock_guard<std::mutex> lock(mu);
        const int dev = std::max(0, std::min(e->device, 63));
        const int need = static_cast<int>(std::max<size_t>(smem_max, 1024));
        if (need > raised[dev])
        {
            SFE_CUDA(cudaFuncSetAttribute(fanout_kernel<kStreamScalar, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, need));
            SFE_CUDA(cudaFuncSetAttribute(fanout_kernel<kStreamTma, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, need));
            SFE_CUDA(cudaFuncSetAttribute(fanout_kernel<kStreamQ4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, need));
            SFE_CUDA(cudaFuncSetAttribute(fanout_kernel<kStreamScalar, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, need));
            SFE_CUDA(cudaFuncSetAttribute(fanout_kernel<kStreamTma, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, need));
            SFE_CUDA(cudaFuncSetAttribute(fanout_kernel<kStreamQ4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, need));
            SFE_CUDA(cudaFuncSetAttribute(fanout_kernel_batch<kStreamScalar>, cudaFuncAttributeMaxDynamicSharedMemorySize, need));
            SFE_CUDA(cudaFuncSetAttribute(fanout_kernel_batch<kStreamTma>, cudaFuncAttributeMaxDynamicSharedMemorySize, need));
            SFE_CUDA(cudaFuncSetAttribute(fanout_kernel_batch<kStreamQ4>, cudaFuncAttributeMaxDynamicSharedMemorySize, need));
            raised[dev] = need;
        }
    }
    {
        int per_sm = 1, sms = 1;
        if (e->fanout_variant == kStreamQ4)
            SFE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fanout_kernel<kStreamQ4, false>, kFanoutThreads, smem_max));
        else if (e->fanout_variant == kStreamTma)
            SFE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fanout_kernel<kStreamTma, false>, kFanoutThreads, smem_max));
        else
            SFE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fanout_kernel<kStreamScalar, false>, kFanoutThreads, smem_max));
        {
            // the fused step kernel carries the neuron phase too (80 registers, 3 CTAs per SM): one grid size serves
            // both kernels of an engine, so the smaller occupancy decides
            const char *fused = std::getenv("SFE_FUSED_STEP");
            const bool want_fused = (fused != nullptr && std::atoi(fused) != 0) && !e->exotic && e->n_taps_units == 0;
            int per_sm_fused = per_sm;
            if (want_fused && e->fanout_variant == kStreamQ4)
                SFE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_fused, fanout_kernel<kStreamQ4, true>, kFanoutThreads, smem_max));
            else if (want_fused && e->fanout_variant == kStreamTma)
                SFE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_fused, fanout_kernel<kStreamTma, true>, kFanoutThreads, smem_max));
            else if (want_fused)
                SFE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_fused, fanout_kernel<kStreamScalar, true>, kFanoutThreads, smem_max));
            per_sm = std::max(1, std::min(per_sm, per_sm_fused));
        }
        SFE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, e->device));
        if (const char *v = std::getenv("SFE_FANOUT_CTAS_PER_SM")) per_sm = std::max(1, std::min(per_sm, std::atoi(v)));
        // ---- message-phase work items: split every core's inbox into slices so that there
        // are several items per resident CTA (no tail behind the slowest core, and a rank
        // that owns few cores still fills the GPU)
        const size_t slots = static_cast<size_t>(per_sm) * sms;
        // The split is chosen with a small cost model: a CTA that processes k items pays k times a
        // fixed latency (ticket, descriptors, inbox scan, axon records, epilogue: ~5 us of dependent
        // round trips) plus its share of the streaming time, and the step ends with the CTA that
        // drew the most items. Few cores -> exactly one item per CTA; many cores -> enough items
        // per CTA that the last wave is short.
        size_t split = 1;
        if (!e->fanout_list.empty())
        {
            double syn = 0.0;
            for (uint32_t c : e->fanout_list)
                syn += static_cast<double>(tb->cores[c].syn_count) * (e->h_cores[c].q4 != 0u ? 4.0 : 12.0); // bytes per record
            const double n_cores_f = static_cast<double>(e->fanout_list.size());
            const double core_bytes = 0.1 * syn / n_cores_f; // nominal 10 % activity
            const double fixed_us = 5.0, hbm_bytes_per_us = 6.5e6, cta_bytes_per_us = 3.0e4;
            double best = 0.0;
            for (size_t cand = 1; cand <= 16; ++cand)
            {
                const double items = n_cores_f * static_cast<double>(cand);
                const double per_cta = std::ceil(items / static_cast<double>(slots));
                const double rate = std::min(cta_bytes_per_us, hbm_bytes_per_us / std::min<double>(items, static_cast<double>(slots)));
                const double cost = per_cta * (fixed_us + core_bytes / static_cast<double>(cand) / rate);
                // fewer items unless clearly better; a single wave should fill the resident CTAs
                if (cand == 1 || cost < best * 0.98 || (per_cta == 1.0 && cost <= best * 1.001))
                {
                    best = cost;
                    split = cand;
                }
            }
        }
        if (const char *v = std::getenv("SFE_FANOUT_SPLIT")) split = std::max(1, std::atoi(v));
        // parts per core: `split`, capped by the core's inbox size (>= 8 words per slice; an ordered core is one item)
        std::vector<uint32_t> parts_of(tb->n_cores, 1u);
        size_t n_items = 0;
        for (uint32_t c : e->fanout_list)
        {
            const CoreDev &d = e->h_cores[c];
            const uint32_t words = (d.axon_count + 31u) / 32u;
            parts_of[c] = d.acc_mode == SFE_ACC_ORDERED ? 1u : static_cast<uint32_t>(std::max<size_t>(1, std::min<size_t>(split, words / 8)));
            n_items += parts_of[c];
        }
        std::vector<FanItem> items;
        std::vector<double> weight;
        for (uint32_t c : e->fanout_list)
        {
            const CoreDev &d = e->h_cores[c];
            const uint32_t words = (d.axon_count + 31u) / 32u;
            const uint32_t parts = parts_of[c];
            e->h_cores[c].item_begin = static_cast<uint32_t>(items.size());
            e->h_cores[c].item_count = parts;
            for (uint32_t k = 0; k < parts; ++k)
            {
                FanItem it;
                it.core = c;
                it.word_lo = static_cast<uint32_t>(static_cast<uint64_t>(words) * k / parts);
                it.word_hi = static_cast<uint32_t>(static_cast<uint64_t>(words) * (k + 1) / parts);
                it.pad = 0;
                items.push_back(it);
                weight.push_back(static_cast<double>(tb->cores[c].syn_count) / parts);
            }
        }
        // active cores without axons-in get one empty item each: in the fused step kernel every core passes through
        // the same completion path (fold, neuron phase of the next step), and an empty item costs next to nothing
        for (uint32_t c : e->active_list)
        {
            if (e->h_cores[c].axon_count != 0) continue;
            e->h_cores[c].item_begin = static_cast<uint32_t>(items.size());
            e->h_cores[c].item_count = 1;
            items.push_back(FanItem{c, 0u, 0u, 0u});
            weight.push_back(0.0);
        }
        std::vector<uint32_t> order(items.size());
        for (uint32_t i = 0; i < order.size(); ++i) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return weight[a] > weight[b]; });
        e->t.n_fan_items = static_cast<uint32_t>(items.size());
        if (e->upload(&e->t.fan_items, items.data(), items.size()) != 0) return -1;
        if (e->upload(&e->t.fan_order, order.data(), order.size()) != 0) return -1;
        if (e->alloc(&e->s.stats_m, items.size()) != 0) return -1;
        SFE_CUDA(cudaMemcpyAsync(e->d_cores, e->h_cores.data(), e->h_cores.size() * sizeof(CoreDev), cudaMemcpyHostToDevice, e->stream));
        SFE_CUDA(cudaStreamSynchronize(e->stream));
        e->fanout_grid = static_cast<unsigned>(std::max<size_t>(1, std::min<size_t>(items.size(), slots)));
        // The fused step kernel (one launch per step, see enqueue_fused) serves enqueue-only batches of engines whose
        // somas are all LIF / TrueNorth (the kernel carries the plain neuron-phase instantiation); SFE_FUSED_STEP=0
        // keeps the two-kernel step everywhere.
        {
            // Measured (profiles/r2_fused_step.md): the neuron phase of a core run by ONE CTA is a chain of memory round
            // trips (~14 us from a core's last item to its raster words, against ~6 us for the stand-alone kernel that
            // spreads the same work over hundreds of CTAs), so the two-kernel step wins at every scale tried (C4 on
            // 1/2/4/8 GPUs, a 128-core chip). The fused kernel stays as an option: SFE_FUSED_STEP=1.
            const char *fused = std::getenv("SFE_FUSED_STEP");
            e->fused_ok = (fused != nullptr && std::atoi(fused) != 0) && !e->exotic && e->n_taps_units == 0 && e->world == 1 &&
                    !e->soma_list.empty() && !e->fanout_list.empty();
            const char *piggy = std::getenv("SFE_PIGGYBACK_FINALIZE"); // 0: a finalize kernel per step
            e->piggyback = piggy == nullptr || std::atoi(piggy) != 0;
        }
    }
    SFE_CUDA(cudaStreamSynchronize(e->stream));
    return 0;
}

extern "C" sfe_engine *sfe_engine_create_partitioned(const sfe_tables *tables, int device, uint32_t rank, uint32_t world)
{
    if (tables == nullptr || tables->abi_version != SFE_ABI_VERSION)
    {
        sfe::set_last_error("sfe_engine_create: bad tables / ABI version");
        return nullptr;
    }
    if (world == 0 || rank >= world)
    {
        sfe::set_last_error("sfe_engine_create_partitioned: need rank < world");
        return nullptr;
    }
    if (sfe_device_count() <= 0)
    {
        sfe::set_last_error("no CUDA device: the B200 engine has no CPU fallback");
        return nullptr;
    }
    sfe_engine *e = new sfe_engine();
    e->device = device;
    e->rank = rank;
    e->world = world;
    if (engine_build(e, tables) != 0)
    {
        sfe_engine_destroy(e);
        return nullptr;
    }
    return e;
}

extern "C" sfe_engine *sfe_engine_create(const sfe_tables *tables, int device)
{
    return sfe_engine_create_partitioned(tables, device, 0, 1);
}

extern "C" void sfe_engine_destroy(sfe_engine *e)
{
    if (e == nullptr) return;
    cudaSetDevice(e->device);
    if (e->stream != nullptr) cudaStreamSynchronize(e->stream);
    for (void *&p : e->p2p_peer_base)
        if (p != nullptr) cudaIpcCloseMemHandle(p);
    if (e->copy_stream != nullptr)
    {
        cudaStreamSynchronize(e->copy_stream);
        cudaStreamDestroy(e->copy_stream);
        for (int k = 0; k < 2; ++k)
        {
            cudaEventDestroy(e->bias_ready[k]);
            cudaEventDestroy(e->bias_free[k]);
        }
    }
    for (void *p : e->allocs) cudaFree(p);
    if (e->d_overlay != nullptr) cudaFree(e->d_overlay);
    if (e->d_rand_overlay != nullptr) cudaFree(e->d_rand_overlay);
    if (e->pinned != nullptr) cudaFreeHost(e->pinned);
    if (e->pinned_fired != nullptr) cudaFreeHost(e->pinned_fired);
    if (e->d_fired_ring != nullptr) cudaFree(e->d_fired_ring);
    if (e->bias_stage != nullptr) cudaFreeHost(e->bias_stage);
    if (e->ev_begin != nullptr) cudaEventDestroy(e->ev_begin);
    if (e->ev_end != nullptr) cudaEventDestroy(e->ev_end);
    for (cudaEvent_t ev : e->ev_pool) cudaEventDestroy(ev);
    if (e->own_stream && e->stream != nullptr) cudaStreamDestroy(e->stream);
    delete e;
}

extern "C" int sfe_engine_set_stream(sfe_engine *e, void *stream)
{
    SFE_CUDA(cudaStreamSynchronize(e->stream));
    if (e->own_stream) cudaStreamDestroy(e->stream);
    e->own_stream = false;
    e->stream = static_cast<cudaStream_t>(stream);
    return 0;
}

// Kernels of the step are launched with programmatic stream serialisation (SFE_PDL=0 turns it
// off): the launch latency and the table-only prologue of a kernel overlap the tail of the
// kernel before it; griddepcontrol.wait inside the kernel keeps the data dependencies.
template <typename... KArgs, typename... Args>
static void launch_step_kernel(sfe_engine *e, void (*kernel)(KArgs...), dim3 grid, unsigned block, size_t smem, Args... args)
{
    static const bool pdl = [] {
        const char *v = std::getenv("SFE_PDL");
        return v == nullptr || std::atoi(v) != 0;
    }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = e->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Out-of-tree soma device models: dendrite outputs of their instances, then every model's own update kernel, all
// ahead of the neuron-phase kernel on the engine's stream (plain launches: fully ordered).
static void launch_device_models(sfe_engine *e)
{
    if (e->plug.empty() || e->n_plug == 0) return;
    plug_gather_kernel<<<(e->n_plug + 127) / 128, 128, 0, e->stream>>>(e->d_plug_gather, e->n_plug, e->s, e->sv, e->d_plug_in, e->d_plug_has);
    ++e->launches;
    for (const sfe_engine::PlugModel &m : e->plug)
    {
        sfe_device_model_launch a{};
        a.stream = e->stream;
        a.n_instances = m.n;
        a.n_state = m.desc->n_state;
        a.n_params = m.desc->n_params;
        a.state = m.d_state;
        a.params = m.d_params;
        a.current_in = e->d_plug_in + m.first;
        a.has_in = e->d_plug_has + m.first;
        a.status = e->d_plug_status + m.first;
        a.timestep = e->total_timesteps + 1;
        if (m.desc->launch(&a) != 0) e->plug_failed = true;
        ++e->launches;
    }
}

template <typename... KArgs, typename... Args>
static void launch_step_kernel(sfe_engine *e, void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, Args... args)
{
    launch_step_kernel(e, kernel, dim3(grid), block, smem, args...);
}

static void launch_soma(sfe_engine *e)
{
    e->sv.steps_done = e->total_timesteps;
    launch_device_models(e);
    unsigned grid = e->n_segments;
    if (e->pending_fold)
    {
        grid += e->final_grid; // extra CTAs fold the previous step
        e->sv.fold_parity = e->pending_parity;
        e->sv.fold_step = e->pending_step;
        e->pending_fold = false;
    }
    e->sv.soma_grid = grid;
    if (e->exotic) launch_step_kernel(e, soma_kernel<true>, grid, kSomaThreads, 0, e->t, e->s, e->sv);
    else launch_step_kernel(e, soma_kernel<false>, grid, kSomaThreads, 0, e->t, e->s, e->sv);
}

// "taps" lines of the step: after the neuron phase has written the raster, before the message
// phase; a plain launch, i.e. fully ordered after the neuron phase like the probe kernel
static void launch_taps(sfe_engine *e)
{
    if (e->n_taps_units == 0) return;
    e->sv.steps_done = e->total_timesteps;
    taps_kernel<<<(e->n_taps_units + 127) / 128, 128, 0, e->stream>>>(e->t, e->s, e->sv);
    ++e->launches;
}

// The stand-alone fold of a step whose records are needed now (end of a batch, collect, reset...).
static void flush_fold(sfe_engine *e)
{
    if (!e->pending_fold) return;
    e->sv.fold_parity = e->pending_parity;
    e->sv.fold_step = e->pending_step;
    launch_step_kernel(e, finalize_kernel, e->final_grid, kFinalThreads, 0, e->t, e->s, e->sv);
    ++e->launches;
    e->pending_fold = false;
}

// Clears the ticket counters of the next `count` steps (one memset per batch, ahead of its first kernel: it waits for
// every earlier launch, stragglers included, and nothing of the batch has started).
static int prepare_tickets(sfe_engine *e, const int64_t count)
{
    if (count <= 0 || e->d_work_pool == nullptr) return 0;
    const uint32_t first = static_cast<uint32_t>(e->sv.step_seq % kWorkPool);
    const uint32_t n = static_cast<uint32_t>(std::min<int64_t>(count, kWorkPool));
    const uint32_t head = std::min(n, kWorkPool - first);
    SFE_CUDA(cudaMemsetAsync(e->d_work_pool + first, 0, head * sizeof(uint32_t), e->stream));
    if (n > head) SFE_CUDA(cudaMemsetAsync(e->d_work_pool, 0, (n - head) * sizeof(uint32_t), e->stream));
    return 0;
}

static void launch_fanout(sfe_engine *e, const bool fused = false)
{
    e->sv.steps_done = e->total_timesteps;
    e->sv.work = e->d_work_pool + (e->sv.step_seq % kWorkPool);
    const unsigned grid = e->fanout_grid;
    e->sv.fanout_grid = grid;
    if (e->fanout_variant == kStreamQ4)
    {
        if (fused) launch_step_kernel(e, fanout_kernel<kStreamQ4, true>, grid, kFanoutThreads, e->fanout_smem, e->t, e->s, e->sv, e->tma_off);
        else launch_step_kernel(e, fanout_kernel<kStreamQ4, false>, grid, kFanoutThreads, e->fanout_smem, e->t, e->s, e->sv, e->tma_off);
    }
    else if (e->fanout_variant == kStreamTma)
    {
        if (fused) launch_step_kernel(e, fanout_kernel<kStreamTma, true>, grid, kFanoutThreads, e->fanout_smem, e->t, e->s, e->sv, e->tma_off);
        else launch_step_kernel(e, fanout_kernel<kStreamTma, false>, grid, kFanoutThreads, e->fanout_smem, e->t, e->s, e->sv, e->tma_off);
    }
    else
    {
        if (fused) launch_step_kernel(e, fanout_kernel<kStreamScalar, true>, grid, kFanoutThreads, e->fanout_smem, e->t, e->s, e->sv, e->tma_off);
        else launch_step_kernel(e, fanout_kernel<kStreamScalar, false>, grid, kFanoutThreads, e->fanout_smem, e->t, e->s, e->sv, e->tma_off);
    }
}

// Raises the ready flag of the step whose neuron phase a stand-alone soma_kernel has just run (unpartitioned chip), or
// pushes the rank's raster slice to every peer and raises the arrival flags there: what the fused step kernel does
// itself for every later step of a batch.
__global__ void __launch_bounds__(256) publish_kernel(const DevTables t, const DevState s, const StepVars sv)
{
    (void) t;
    griddep_launch_dependents();
    griddep_wait();
    if (threadIdx.x == 0) ready_publish(s.ready, sv.step_seq);
}

static int apply_pending_bias(sfe_engine *e);
static int check_overlay(const sfe_engine *e);

// `timesteps` steps of the fused step kernel (enqueue-only paths: no per-step traces): one stand-alone neuron phase
// for the first step, then ONE launch per step that processes the step's messages, folds it and - all but the last -
// runs the neuron phase of the next step core by core as the cores' message phases finish. timesteps + 2 launches.
static int enqueue_fused(sfe_engine *e, const int64_t timesteps)
{
    if (prepare_tickets(e, timesteps) != 0) return -1;
    if (check_overlay(e) != 0) return -1;
    if (apply_pending_bias(e) != 0) return -1;
    launch_soma(e); // (the fold of an earlier two-kernel step rides along)
    ++e->launches;
    launch_step_kernel(e, publish_kernel, 1u, 256u, 0, e->t, e->s, e->sv);
    ++e->launches;
    for (int64_t i = 0; i < timesteps; ++i)
    {
        const bool timed = e->timing && e->time_launches && e->ev_used + 2 <= 8192;
        if (timed)
        {
            while (e->ev_pool.size() < e->ev_used + 2)
            {
                cudaEvent_t ev;
                cudaEventCreate(&ev);
                e->ev_pool.push_back(ev);
            }
            cudaEventRecord(e->ev_pool[e->ev_used], e->stream);
        }
        e->sv.fuse_next = i + 1 < timesteps ? 1u : 0u;
        e->sv.fold_step = e->total_timesteps;
        launch_fanout(e, true);
        ++e->launches;
        if (timed)
        {
            cudaEventRecord(e->ev_pool[e->ev_used + 1], e->stream);
            e->ev_used += 2;
        }
        ++e->sv.step_seq;
        ++e->total_timesteps;
    }
    e->sv.fuse_next = 0u;
    return 0;
}

static void launch_finalize(sfe_engine *e)
{
    if (e->piggyback && !e->soma_list.empty())
    {
        // folded by extra CTAs of the next neuron-phase kernel (or by flush_fold)
        e->pending_fold = true;
        e->pending_parity = static_cast<uint32_t>(e->sv.step_seq & 1ull);
        e->pending_step = e->total_timesteps;
    }
    else
    {
        e->sv.fold_parity = static_cast<uint32_t>(e->sv.step_seq & 1ull);
        e->sv.fold_step = e->total_timesteps;
        launch_step_kernel(e, finalize_kernel, e->final_grid, kFinalThreads, 0, e->t, e->s, e->sv);
        ++e->launches;
    }
    ++e->sv.step_seq;
}

static int apply_pending_bias(sfe_engine *e);

// A chip with Poisson inputs can only step through timesteps its overlay covers: stepping past it
// would silently drop the random spikes.
static int check_overlay(const sfe_engine *e)
{
    if (e->n_rand_cols != 0 && !(e->total_timesteps >= e->t.rand_step0 && e->total_timesteps < e->t.rand_step0 + e->t.rand_steps))
    {
        sfe::set_last_error("this chip has TrueNorth neurons with a random threshold mask: call sfe_engine_set_rand_overlay for the "
                            "steps to be simulated first; sfe_chip_sim does it itself");
        return -1;
    }
    if (e->n_poisson_cols == 0) return 0;
    if (e->total_timesteps >= e->t.overlay_step0 && e->total_timesteps < e->t.overlay_step0 + e->t.overlay_steps) return 0;
    sfe::set_last_error("this chip has Poisson inputs: call sfe_engine_set_input_overlay (see sfe_poisson_fill) for the "
                        "steps to be simulated first; sfe_chip_sim does it itself");
    return -1;
}

static int enqueue_step(sfe_engine *e, bool probes)
{
    if (check_overlay(e) != 0) return -1;
    if (apply_pending_bias(e) != 0) return -1;
    if (!e->soma_list.empty())
    {
        launch_soma(e);
        ++e->launches;
    }
    launch_taps(e);
    if (probes && e->n_probes + e->n_u_probes > 0)
    {
        probe_kernel<<<(e->n_probes + e->n_u_probes + 255) / 256, 256, 0, e->stream>>>(e->t, e->s);
        ++e->launches;
    }
    if (!e->fanout_list.empty())
    {
        const bool timed = e->timing && e->time_launches && e->ev_used + 2 <= 8192;
        if (timed)
        {
            while (e->ev_pool.size() < e->ev_used + 2)
            {
                cudaEvent_t ev;
                cudaEventCreate(&ev);
                e->ev_pool.push_back(ev);
            }
            cudaEventRecord(e->ev_pool[e->ev_used], e->stream);
        }
        launch_fanout(e);
        ++e->launches;
        if (timed)
        {
            cudaEventRecord(e->ev_pool[e->ev_used + 1], e->stream);
            e->ev_used += 2;
        }
    }
    launch_finalize(e);
    ++e->total_timesteps;
    return 0;
}

extern "C" int sfe_engine_enqueue(sfe_engine *e, int64_t timesteps)
{
    SFE_CUDA(cudaSetDevice(e->device));
    if (e->world > 1)
    {
        sfe::set_last_error("sfe_engine_enqueue: a partitioned engine needs the raster exchange between the phases");
        return -1;
    }
    if (e->total_timesteps - e->log_read + timesteps > e->log_cap)
    {
        sfe::set_last_error("sfe_engine_enqueue: more than " + std::to_string(e->log_cap) +
                " uncollected steps; call sfe_engine_collect");
        return -1;
    }
    if (e->fused_ok && timesteps >= 2)
    {
        if (enqueue_fused(e, timesteps) != 0) return -1;
    }
    else
    {
        if (prepare_tickets(e, timesteps) != 0) return -1;
        for (int64_t i = 0; i < timesteps; ++i)
            if (enqueue_step(e, false) != 0) return -1;
    }
    SFE_CUDA(cudaGetLastError());
    return 0;
}

// `timesteps` steps of n independent chips with ONE launch per phase and step (blockIdx.y = chip) instead of one per
// chip: design-space sweeps of small chips are bound by the number of launches, not by the work in them. The chips
// must be in lock step (same timestep counters) on one device, unpartitioned, with plain LIF / TrueNorth / ... somas of
// one kernel instantiation, no 'taps', no out-of-tree models, no Poisson inputs, no pending bias upload.
// Returns 0 = enqueued (collect every engine as usual), 1 = this set of chips cannot run as one batch (nothing was
// enqueued: step them one by one), -1 = error.
extern "C" int sfe_engine_batch_enqueue(sfe_engine *const *engines, uint32_t n, int64_t timesteps)
{
    if (n == 0 || timesteps <= 0) return 0;
    if (engines == nullptr || engines[0] == nullptr) return 1;
    sfe_engine *lead = engines[0];
    for (uint32_t k = 0; k < n; ++k)
    {
        const sfe_engine *e = engines[k];
        if (e == nullptr || e->device != lead->device || e->world != 1 || e->n_taps_units != 0 || !e->plug.empty() ||
                e->exotic != lead->exotic || e->fanout_variant != lead->fanout_variant || e->fused_ok || !e->piggyback ||
                e->total_timesteps != lead->total_timesteps || e->sv.step_seq != lead->sv.step_seq ||
                e->pending_fold != lead->pending_fold || (e->pending_fold && (e->pending_parity != lead->pending_parity || e->pending_step != lead->pending_step)) ||
                e->bias_pending >= 0 || e->n_poisson_cols != 0 || e->n_rand_cols != 0 || e->soma_list.empty() || e->fanout_list.empty() || e->timing ||
                e->total_timesteps - e->log_read + timesteps > e->log_cap)
            return 1;
        for (uint32_t j = 0; j < k; ++j)
            if (engines[j] == e) return 1;
    }
    SFE_CUDA(cudaSetDevice(lead->device));
    if (lead->batch_slots_cap < n)
    {
        if (lead->alloc(&lead->d_batch_slots, n) != 0) return -1;
        lead->batch_slots_cap = n;
    }
    std::vector<BatchSlot> slots(n);
    std::vector<uint2> map; // neuron-phase CTAs of all chips (segments, then fold CTAs), then the message-phase CTAs
    size_t max_smem = 0;
    for (uint32_t k = 0; k < n; ++k)
    {
        sfe_engine *e = engines[k];
        if (prepare_tickets(e, timesteps) != 0) return -1;
        BatchSlot &b = slots[k];
        std::memset(&b, 0, sizeof(b));
        b.t = e->t;
        b.s = e->s;
        b.tma_off = e->tma_off;
        b.n_segments = e->n_segments;
        b.final_grid = e->final_grid;
        b.fanout_grid = e->fanout_grid;
        b.work_pool = e->d_work_pool;
        max_smem = std::max(max_smem, e->fanout_smem);
        for (uint32_t x = 0; x < e->n_segments + e->final_grid; ++x) map.push_back(make_uint2(k, x));
    }
    const unsigned soma_ctas = static_cast<unsigned>(map.size());
    for (uint32_t k = 0; k < n; ++k)
        for (uint32_t x = 0; x < engines[k]->fanout_grid; ++x) map.push_back(make_uint2(k, x));
    const unsigned fan_ctas = static_cast<unsigned>(map.size()) - soma_ctas;
    if (lead->batch_map_cap < map.size())
    {
        if (lead->alloc(&lead->d_batch_map, map.size()) != 0) return -1;
        lead->batch_map_cap = map.size();
    }
    // everything the chips' own streams hold (loads, ticket memsets, earlier steps) before the batch touches their state
    for (uint32_t k = 0; k < n; ++k) SFE_CUDA(cudaStreamSynchronize(engines[k]->stream));
    SFE_CUDA(cudaMemcpy(lead->d_batch_slots, slots.data(), n * sizeof(BatchSlot), cudaMemcpyHostToDevice));
    SFE_CUDA(cudaMemcpy(lead->d_batch_map, map.data(), map.size() * sizeof(uint2), cudaMemcpyHostToDevice));
    const uint2 *soma_map = lead->d_batch_map, *fan_map = lead->d_batch_map + soma_ctas;
    for (int64_t i = 0; i < timesteps; ++i)
    {
        BatchStep a{};
        a.steps_done = lead->total_timesteps;
        a.step_seq = lead->sv.step_seq;
        a.fold = lead->pending_fold ? 1u : 0u;
        a.fold_parity = lead->pending_parity;
        a.fold_step = lead->pending_step;
        // (without a step to fold - the first launch after a collect - the chips' fold CTAs leave at once)
        if (lead->exotic) launch_step_kernel(lead, soma_kernel_batch<true>, soma_ctas, kSomaThreads, 0, lead->d_batch_slots, soma_map, a);
        else launch_step_kernel(lead, soma_kernel_batch<false>, soma_ctas, kSomaThreads, 0, lead->d_batch_slots, soma_map, a);
        if (lead->fanout_variant == kStreamQ4) launch_step_kernel(lead, fanout_kernel_batch<kStreamQ4>, fan_ctas, kFanoutThreads, max_smem, lead->d_batch_slots, fan_map, a);
        else if (lead->fanout_variant == kStreamTma) launch_step_kernel(lead, fanout_kernel_batch<kStreamTma>, fan_ctas, kFanoutThreads, max_smem, lead->d_batch_slots, fan_map, a);
        else launch_step_kernel(lead, fanout_kernel_batch<kStreamScalar>, fan_ctas, kFanoutThreads, max_smem, lead->d_batch_slots, fan_map, a);
        lead->launches += 2;
        for (uint32_t k = 0; k < n; ++k)
        {
            sfe_engine *e = engines[k];
            // (launch_finalize: the fold of this step rides on the next neuron-phase launch, or on flush_fold)
            e->pending_fold = true;
            e->pending_parity = static_cast<uint32_t>(e->sv.step_seq & 1ull);
            e->pending_step = e->total_timesteps;
            ++e->sv.step_seq;
            ++e->total_timesteps;
        }
    }
    SFE_CUDA(cudaGetLastError());
    // whatever the chips' own streams do next (the fold of the last step, collects) comes after the batch
    cudaEvent_t done;
    SFE_CUDA(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
    SFE_CUDA(cudaEventRecord(done, lead->stream));
    for (uint32_t k = 1; k < n; ++k) SFE_CUDA(cudaStreamWaitEvent(engines[k]->stream, done, 0));
    SFE_CUDA(cudaEventDestroy(done));
    return 0;
}

// The overlay of the next n_steps steps drawn on the device (poisson_kernel) instead of uploaded.
// Advances the device generators; do not mix with host-drawn overlays on one engine (two streams of draws).
extern "C" int sfe_engine_fill_input_overlay(sfe_engine *e, int64_t n_steps)
{
    SFE_CUDA(cudaSetDevice(e->device));
    if (e->n_poisson_cols == 0 || n_steps < 0)
    {
        sfe::set_last_error("sfe_engine_fill_input_overlay: this chip has no Poisson inputs");
        return -1;
    }
    SFE_CUDA(cudaStreamSynchronize(e->stream)); // steps already enqueued may still be reading the previous overlay
    const size_t bytes = static_cast<size_t>(n_steps) * e->n_poisson_cols;
    if (bytes > e->overlay_cap)
    {
        if (e->d_overlay != nullptr) cudaFree(e->d_overlay);
        e->d_overlay = nullptr;
        e->overlay_cap = 0;
        void *q = nullptr;
        SFE_CUDA(cudaMalloc(&q, bytes));
        e->d_overlay = static_cast<uint8_t *>(q);
        e->overlay_cap = bytes;
    }
    if (bytes > 0)
    {
        SFE_CUDA(cudaMemsetAsync(e->d_overlay, 0, bytes, e->stream));
        poisson_kernel<<<(e->n_poisson_units + 63) / 64, 64, 0, e->stream>>>(e->d_poisson_units, e->d_poisson_cols,
                e->n_poisson_units, e->d_mt, e->d_mt_idx, e->d_overlay, n_steps, e->n_poisson_cols);
        ++e->launches;
        SFE_CUDA(cudaGetLastError());
    }
    e->t.overlay = e->d_overlay;
    e->t.overlay_cols = e->n_poisson_cols;
    e->t.overlay_step0 = e->total_timesteps;
    e->t.overlay_steps = n_steps;
    return 0;
}

extern "C" int sfe_engine_set_rand_overlay(sfe_engine *e, const uint32_t *values, int64_t n_steps, uint32_t n_cols)
{
    SFE_CUDA(cudaSetDevice(e->device));
    if (n_cols != e->n_rand_cols || n_steps < 0 || (values == nullptr && n_steps > 0 && n_cols > 0))
    {
        sfe::set_last_error("sfe_engine_set_rand_overlay: need values[n_steps][" + std::to_string(e->n_rand_cols) + "]");
        return -1;
    }
    SFE_CUDA(cudaStreamSynchronize(e->stream)); // steps already enqueued may still be reading the previous overlay
    const size_t count = static_cast<size_t>(n_steps) * n_cols;
    if (count > e->rand_overlay_cap)
    {
        if (e->d_rand_overlay != nullptr) cudaFree(e->d_rand_overlay);
        e->d_rand_overlay = nullptr;
        e->rand_overlay_cap = 0;
        void *q = nullptr;
        SFE_CUDA(cudaMalloc(&q, count * sizeof(uint32_t)));
        e->d_rand_overlay = static_cast<uint32_t *>(q);
        e->rand_overlay_cap = count;
    }
    if (count > 0) SFE_CUDA(cudaMemcpy(e->d_rand_overlay, values, count * sizeof(uint32_t), cudaMemcpyHostToDevice));
    e->t.rand_overlay = e->d_rand_overlay;
    e->t.rand_cols = n_cols;
    e->t.rand_step0 = e->total_timesteps;
    e->t.rand_steps = n_steps;
    return 0;
}

extern "C" void sfe_engine_request_stop(sfe_engine *e, int on)
{
    e->stop_requested.store(on != 0, std::memory_order_relaxed);
}

extern "C" int sfe_engine_set_input_overlay(sfe_engine *e, const uint8_t *bits, int64_t n_steps, uint32_t n_cols)
{
    SFE_CUDA(cudaSetDevice(e->device));
    if (n_cols != e->n_poisson_cols || n_steps < 0 || (bits == nullptr && n_steps > 0 && n_cols > 0))
    {
        sfe::set_last_error("sfe_engine_set_input_overlay: need bits[n_steps][" + std::to_string(e->n_poisson_cols) + "]");
        return -1;
    }
    // steps already enqueued may still be reading the previous overlay
    SFE_CUDA(cudaStreamSynchronize(e->stream));
    const size_t bytes = static_cast<size_t>(n_steps) * n_cols;
    if (bytes > e->overlay_cap)
    {
        if (e->d_overlay != nullptr) cudaFree(e->d_overlay);
        e->d_overlay = nullptr;
        e->overlay_cap = 0;
        void *q = nullptr;
        SFE_CUDA(cudaMalloc(&q, bytes));
        e->d_overlay = static_cast<uint8_t *>(q);
        e->overlay_cap = bytes;
    }
    if (bytes > 0) SFE_CUDA(cudaMemcpy(e->d_overlay, bits, bytes, cudaMemcpyHostToDevice));
    e->t.overlay = e->d_overlay;
    e->t.overlay_cols = n_cols;
    e->t.overlay_step0 = e->total_timesteps;
    e->t.overlay_steps = n_steps;
    return 0;
}

static void add_record(sfe_run_data &rd, const sfe_step_record &r)
{
    // update_run_data  src/chip.cpp:462-475 (same accumulation order: step by step)
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
}

extern "C" int sfe_engine_exchange_error(sfe_engine *e);
static int collect_records(sfe_engine *e, std::vector<sfe_step_record> &out)
{
    const int64_t pending = e->total_timesteps - e->log_read;
    out.resize(static_cast<size_t>(pending));
    flush_fold(e);
    SFE_CUDA(cudaStreamSynchronize(e->stream));
    if (e->plug_failed)
    {
        sfe::set_last_error("an out-of-tree device model failed to launch its update kernel; results are invalid");
        return -1;
    }
    if (e->p2p_on && sfe_engine_exchange_error(e) != 0) return -1; // (the message names the word / rank that did not arrive)
    int64_t done = 0;
    while (done < pending)
    {
        const int64_t pos = (e->log_read + done) % e->log_cap;
        const int64_t chunk = std::min<int64_t>(pending - done, e->log_cap - pos);
        SFE_CUDA(cudaMemcpy(out.data() + done, e->s.log + pos, chunk * sizeof(sfe_step_record), cudaMemcpyDeviceToHost));
        done += chunk;
    }
    e->log_read += pending;
    return 0;
}

extern "C" int sfe_engine_collect(sfe_engine *e, sfe_run_data *out)
{
    SFE_CUDA(cudaSetDevice(e->device));
    sfe_run_data rd;
    std::memset(&rd, 0, sizeof(rd));
    rd.timestep_start = e->log_read + 1;
    std::vector<sfe_step_record> recs;
    if (collect_records(e, recs) != 0) return -1;
    rd.timesteps_executed = static_cast<int64_t>(recs.size());
    for (const sfe_step_record &r : recs) add_record(rd, r);
    if (out != nullptr) *out = rd;
    return 0;
}

extern "C" int sfe_engine_run(sfe_engine *e, int64_t timesteps, const sfe_trace_request *req, sfe_run_data *out)
{
    SFE_CUDA(cudaSetDevice(e->device));
    if (e->total_timesteps != e->log_read)
    {
        sfe::set_last_error("sfe_engine_run: uncollected enqueued steps; call sfe_engine_collect first");
        return -1;
    }
    if (e->world > 1)
    {
        sfe::set_last_error("sfe_engine_run: a partitioned engine is stepped with sfe_engine_enqueue_neuron_phase / "
                            "exchange / sfe_engine_enqueue_message_phase");
        return -1;
    }
    sfe_run_data rd;
    std::memset(&rd, 0, sizeof(rd));
    rd.timestep_start = e->total_timesteps + 1; // src/chip.cpp:481
    rd.timesteps_executed = timesteps;
    const bool want_fired = req != nullptr && req->fired_bits != nullptr;
    const bool want_pot = req != nullptr && req->potentials != nullptr && e->n_probes > 0;
    const bool want_status = req != nullptr && req->status != nullptr;
    const bool want_u = req != nullptr && req->neuron_traces != nullptr && e->n_u_probes > 0;
    const size_t words = (static_cast<size_t>(e->n_neurons) + 31) / 32;
    // per-step staging in pinned memory, drained once per batch (no per-step host sync)
    const size_t fired_bytes = e->fired_words * sizeof(uint32_t);
    const size_t per_step = (want_pot ? e->n_probes * sizeof(double) : 0) + (want_u ? e->n_u_probes * sizeof(double) : 0) +
            (want_status ? e->n_neurons : 0);
    const size_t per_step_all = per_step + (want_fired ? fired_bytes : 0);
    const int64_t batch_cap = per_step_all == 0 ? e->log_cap
                                                : std::max<int64_t>(1, std::min<int64_t>(e->log_cap, (256ll << 20) / static_cast<int64_t>(per_step_all)));
    std::vector<sfe_step_record> recs;
    int64_t done = 0;
    while (done < timesteps)
    {
        // (a chip with Poisson inputs is only interrupted before its first batch: the random spikes of the whole call
        // have been drawn already, and stopping in the middle would leave the generators ahead of the timestep counter)
        if (e->stop_requested.load(std::memory_order_relaxed) && (done == 0 || (e->n_poisson_cols == 0 && e->n_rand_cols == 0)))
        {
            // the steps simulated so far stay simulated (state and timestep counter), like an interrupted reference run
            sfe::set_last_error("simulation interrupted after " + std::to_string(done) + " of " + std::to_string(timesteps) + " timesteps");
            return -1;
        }
        const int64_t batch = std::min<int64_t>(batch_cap, timesteps - done);
        if (per_step > 0 && e->ensure_pinned(per_step * static_cast<size_t>(batch)) != 0) return -1;
        if (want_fired)
        {
            if (static_cast<size_t>(batch) > e->fired_ring_steps)
            {
                if (e->d_fired_ring != nullptr) cudaFree(e->d_fired_ring);
                e->d_fired_ring = nullptr;
                e->fired_ring_steps = 0;
                void *q = nullptr;
                SFE_CUDA(cudaMalloc(&q, fired_bytes * static_cast<size_t>(batch)));
                e->d_fired_ring = static_cast<uint32_t *>(q);
                e->fired_ring_steps = static_cast<size_t>(batch);
            }
            if (fired_bytes * static_cast<size_t>(batch) > e->pinned_fired_bytes)
            {
                if (e->pinned_fired != nullptr) cudaFreeHost(e->pinned_fired);
                e->pinned_fired = nullptr;
                e->pinned_fired_bytes = 0;
                SFE_CUDA(cudaMallocHost(&e->pinned_fired, fired_bytes * static_cast<size_t>(batch)));
                e->pinned_fired_bytes = fired_bytes * static_cast<size_t>(batch);
            }
        }
        if (prepare_tickets(e, batch) != 0) return -1;
        for (int64_t b = 0; b < batch; ++b)
        {
            if (want_fired)
            {
                // this step's raster slot (an unpartitioned engine: the neuron phase writes it, taps lines read it)
                e->s.fired_bits = e->d_fired_ring + static_cast<size_t>(b) * e->fired_words;
                e->s.fired_global = e->s.fired_bits;
            }
            if (check_overlay(e) != 0) return -1;
            if (apply_pending_bias(e) != 0) return -1;
            if (!e->soma_list.empty())
            {
                launch_soma(e);
                ++e->launches;
            }
            launch_taps(e);
            unsigned char *stage = static_cast<unsigned char *>(e->pinned) + per_step * static_cast<size_t>(b);
            if (want_pot || want_u)
            {
                probe_kernel<<<(e->n_probes + e->n_u_probes + 255) / 256, 256, 0, e->stream>>>(e->t, e->s);
                ++e->launches;
            }
            if (want_pot)
            {
                SFE_CUDA(cudaMemcpyAsync(stage, e->s.probe_out, e->n_probes * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
                stage += e->n_probes * sizeof(double);
            }
            if (want_u)
            {
                SFE_CUDA(cudaMemcpyAsync(stage, e->s.probe_out + e->n_probes, e->n_u_probes * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
                stage += e->n_u_probes * sizeof(double);
            }
            if (want_status)
                SFE_CUDA(cudaMemcpyAsync(stage, e->s.status, e->n_neurons, cudaMemcpyDeviceToHost, e->stream));
            if (!e->fanout_list.empty())
            {
                launch_fanout(e);
                ++e->launches;
            }
            launch_finalize(e);
            ++e->total_timesteps;
        }
        if (want_fired)
        {
            // the batch's rasters in one copy; the last one stays readable through sfe_engine_read_raster
            SFE_CUDA(cudaMemcpyAsync(e->pinned_fired, e->d_fired_ring, fired_bytes * static_cast<size_t>(batch), cudaMemcpyDeviceToHost, e->stream));
            SFE_CUDA(cudaMemcpyAsync(e->d_fired_global, e->d_fired_ring + static_cast<size_t>(batch - 1) * e->fired_words, fired_bytes,
                    cudaMemcpyDeviceToDevice, e->stream));
            e->s.fired_bits = e->d_fired_local;
            e->s.fired_global = e->d_fired_global;
        }
        SFE_CUDA(cudaGetLastError());
        if (collect_records(e, recs) != 0) return -1;
        for (int64_t b = 0; b < batch; ++b)
        {
            const sfe_step_record &r = recs[static_cast<size_t>(b)];
            add_record(rd, r);
            if (req != nullptr && req->steps != nullptr) req->steps[done + b] = r;
            const unsigned char *stage = static_cast<const unsigned char *>(e->pinned) + per_step * static_cast<size_t>(b);
            if (want_fired && e->raster_identity)
            {
                std::memcpy(req->fired_bits + static_cast<size_t>(done + b) * words,
                        static_cast<const uint32_t *>(e->pinned_fired) + static_cast<size_t>(b) * e->fired_words, words * sizeof(uint32_t));
            }
            else if (want_fired)
            {
                // un-pad: per-core word-aligned raster -> device-index bit order
                const uint32_t *padded = static_cast<const uint32_t *>(e->pinned_fired) + static_cast<size_t>(b) * e->fired_words;
                uint32_t *dst = req->fired_bits + static_cast<size_t>(done + b) * words;
                std::memset(dst, 0, words * sizeof(uint32_t));
                for (uint32_t c : e->soma_list)
                {
                    const sfe_core_desc &cd = e->core_desc[c];
                    const uint32_t *src = padded + e->fired_word_begin[c];
                    if ((cd.neuron_begin & 31u) == 0)
                    {
                        const uint32_t nw = (cd.neuron_count + 31) / 32;
                        for (uint32_t x = 0; x < nw; ++x) dst[(cd.neuron_begin >> 5) + x] |= src[x];
                    }
                    else
                    {
                        for (uint32_t k = 0; k < cd.neuron_count; ++k)
                            if ((src[k >> 5] >> (k & 31)) & 1u)
                            {
                                const uint32_t i = cd.neuron_begin + k;
                                dst[i >> 5] |= 1u << (i & 31);
                            }
                    }
                }
            }
            if (want_pot)
            {
                std::memcpy(req->potentials + static_cast<size_t>(done + b) * e->n_probes, stage, e->n_probes * sizeof(double));
                stage += e->n_probes * sizeof(double);
            }
            if (want_u)
            {
                std::memcpy(req->neuron_traces + static_cast<size_t>(done + b) * e->n_u_probes, stage, e->n_u_probes * sizeof(double));
                stage += e->n_u_probes * sizeof(double);
            }
            if (want_status) std::memcpy(req->status + static_cast<size_t>(done + b) * e->n_neurons, stage, e->n_neurons);
        }
        done += batch;
    }
    if (out != nullptr) *out = rd;
    return 0;
}

extern "C" int sfe_engine_reset(sfe_engine *e)
{
    // SpikingChip::reset  src/chip.cpp:576-600: pipeline buffers and model state are
    // zeroed (LIF: u and v only; HH: V, m, n, h); timestep counters keep running
    SFE_CUDA(cudaSetDevice(e->device));
    flush_fold(e);
    SFE_CUDA(cudaMemsetAsync(e->s.v, 0, e->n_neurons * sizeof(double), e->stream));
    SFE_CUDA(cudaMemsetAsync(e->s.u, 0, e->n_neurons * sizeof(double), e->stream));
    SFE_CUDA(cudaMemsetAsync(e->s.status, 0, e->n_neurons, e->stream));
    SFE_CUDA(cudaMemsetAsync(e->s.din32, 0, std::max<size_t>(e->dend_cells, 1) * sizeof(uint32_t), e->stream));
    if (e->ordered_any || e->dual_any) SFE_CUDA(cudaMemsetAsync(e->s.dcnt32, 0, std::max<size_t>(e->dend_cells, 1) * sizeof(uint32_t), e->stream));
    if (e->ordered_any) SFE_CUDA(cudaMemsetAsync(e->s.din64, 0, std::max<size_t>(e->dend_cells, 1) * sizeof(double), e->stream));
    if (e->n_hh > 0) SFE_CUDA(cudaMemsetAsync(e->s.hh, 0, 4 * static_cast<size_t>(e->n_hh) * sizeof(double), e->stream));
    if (e->neurofem)
    {
        // NeuroFEMModel::reset  plugins/neurofem.cpp:319-335 (the accumulators are the dendrite cells cleared above)
        SFE_CUDA(cudaMemsetAsync(e->s.nf_u2, 0, e->n_neurons * sizeof(double), e->stream));
        SFE_CUDA(cudaMemsetAsync(e->s.nf_uint, 0, e->n_neurons * sizeof(double), e->stream));
    }
    for (const sfe_engine::PlugModel &m : e->plug)
        for (uint32_t w = 0; w < m.desc->n_state; ++w) // the model's reset(): the state words it declared
            if ((m.desc->reset_mask >> w) & 1u)
                SFE_CUDA(cudaMemsetAsync(m.d_state + static_cast<size_t>(w) * m.n, 0, m.n * sizeof(double), e->stream));
    if (e->n_taps_units > 0)
    {
        // MultiTapModel1D::reset  src/models.cpp:340-348: voltages only (the lines' step counters run on)
        SFE_CUDA(cudaMemsetAsync(e->s.tap_v, 0, std::max<size_t>(e->tap_cells, 1) * sizeof(double), e->stream));
        SFE_CUDA(cudaMemsetAsync(e->s.tap_has, 0, e->n_neurons * sizeof(uint32_t), e->stream));
    }
    SFE_CUDA(cudaStreamSynchronize(e->stream));
    return 0;
}

// The neuron phase of the next enqueued step switches to the most recently uploaded bias vector.
static int apply_pending_bias(sfe_engine *e)
{
    if (e->bias_pending < 0) return 0;
    const int nxt = e->bias_pending;
    SFE_CUDA(cudaEventRecord(e->bias_free[e->bias_cur], e->stream)); // every step enqueued so far read the old buffer
    SFE_CUDA(cudaStreamWaitEvent(e->stream, e->bias_ready[nxt], 0));
    e->s.bias = e->bias_buf[nxt];
    e->bias_cur = nxt;
    e->bias_pending = -1;
    return 0;
}

// Asynchronous: the vector is copied (from pinned memory: by DMA, overlapping the kernels of the
// steps already enqueued) into the inactive one of two device buffers; it takes effect with the
// next step that is enqueued. The host buffer must stay untouched until that step has run.
extern "C" int sfe_engine_set_bias(sfe_engine *e, const double *bias, size_t n)
{
    SFE_CUDA(cudaSetDevice(e->device));
    if (n != e->n_neurons)
    {
        sfe::set_last_error("sfe_engine_set_bias: expected one bias per neuron");
        return -1;
    }
    if (e->copy_stream == nullptr)
    {
        SFE_CUDA(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
        e->bias_buf[0] = e->s.bias;
        if (e->alloc(&e->bias_buf[1], e->n_neurons) != 0) return -1;
        SFE_CUDA(cudaStreamSynchronize(e->stream)); // alloc() clears on the engine's stream
        for (int k = 0; k < 2; ++k)
        {
            SFE_CUDA(cudaEventCreateWithFlags(&e->bias_ready[k], cudaEventDisableTiming));
            SFE_CUDA(cudaEventCreateWithFlags(&e->bias_free[k], cudaEventDisableTiming));
        }
        e->bias_cur = 0;
    }
    const int nxt = e->bias_pending >= 0 ? e->bias_pending : 1 - e->bias_cur; // a not yet adopted upload is overwritten
    SFE_CUDA(cudaStreamWaitEvent(e->copy_stream, e->bias_free[nxt], 0)); // no-op until the event has been recorded
    SFE_CUDA(cudaMemcpyAsync(e->bias_buf[nxt], bias, n * sizeof(double), cudaMemcpyHostToDevice, e->copy_stream));
    SFE_CUDA(cudaEventRecord(e->bias_ready[nxt], e->copy_stream));
    e->bias_pending = nxt;
    return 0;
}

// The same from pageable host memory (the chip object's bias table after MappedNeuron.set_attributes patches): the
// vector is first copied into an engine-owned pinned buffer, so the host->device transfer is a DMA like the one above
// and the caller's memory is free again on return.
extern "C" int sfe_engine_set_bias_staged(sfe_engine *e, const double *bias, size_t n)
{
    SFE_CUDA(cudaSetDevice(e->device));
    if (n != e->n_neurons)
    {
        sfe::set_last_error("sfe_engine_set_bias_staged: expected one bias per neuron");
        return -1;
    }
    if (e->bias_stage == nullptr) SFE_CUDA(cudaMallocHost(reinterpret_cast<void **>(&e->bias_stage), std::max<size_t>(n, 1) * sizeof(double)));
    if (e->copy_stream != nullptr) SFE_CUDA(cudaStreamSynchronize(e->copy_stream)); // the previous upload has left the buffer
    std::memcpy(e->bias_stage, bias, n * sizeof(double));
    return sfe_engine_set_bias(e, e->bias_stage, n);
}

extern "C" int sfe_engine_set_neuron_bias(sfe_engine *e, uint32_t neuron, double bias)
{
    SFE_CUDA(cudaSetDevice(e->device));
    if (neuron >= e->n_neurons)
    {
        sfe::set_last_error("sfe_engine_set_neuron_bias: neuron out of range");
        return -1;
    }
    if (apply_pending_bias(e) != 0) return -1;
    SFE_CUDA(cudaMemcpyAsync(e->s.bias + neuron, &bias, sizeof(double), cudaMemcpyHostToDevice, e->stream));
    SFE_CUDA(cudaStreamSynchronize(e->stream));
    return 0;
}

extern "C" int sfe_engine_set_neuron_potential(sfe_engine *e, uint32_t neuron, double potential)
{
    SFE_CUDA(cudaSetDevice(e->device));
    if (neuron >= e->n_neurons)
    {
        sfe::set_last_error("sfe_engine_set_neuron_potential: neuron out of range");
        return -1;
    }
    SFE_CUDA(cudaMemcpyAsync(e->s.v + neuron, &potential, sizeof(double), cudaMemcpyHostToDevice, e->stream));
    SFE_CUDA(cudaStreamSynchronize(e->stream));
    return 0;
}

int sfe_engine_update_classes(sfe_engine *e, const sfe_soma_class *classes, uint32_t n_classes,
        const uint32_t *neuron_class)
{
    SFE_CUDA(cudaSetDevice(e->device));
    if (n_classes > e->n_classes_cap)
    {
        sfe_soma_class *p = nullptr;
        const uint32_t cap = std::max(n_classes, e->n_classes_cap * 2);
        if (e->alloc(&p, cap) != 0) return -1;
        e->d_classes = p;
        e->t.classes = p;
        e->n_classes_cap = cap;
    }
    e->t.n_soma_classes = n_classes;
    SFE_CUDA(cudaMemcpyAsync(e->d_classes, classes, n_classes * sizeof(sfe_soma_class), cudaMemcpyHostToDevice, e->stream));
    SFE_CUDA(cudaMemcpyAsync(e->d_neuron_class, neuron_class, e->n_neurons * sizeof(uint32_t), cudaMemcpyHostToDevice, e->stream));
    SFE_CUDA(cudaStreamSynchronize(e->stream));
    return 0;
}

extern "C" int sfe_engine_read_potentials(sfe_engine *e, double *out, size_t n)
{
    SFE_CUDA(cudaSetDevice(e->device));
    if (n != e->n_neurons)
    {
        sfe::set_last_error("sfe_engine_read_potentials: expected one slot per neuron");
        return -1;
    }
    SFE_CUDA(cudaStreamSynchronize(e->stream));
    SFE_CUDA(cudaMemcpy(out, e->s.v, n * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

// The last step's fired-bit raster as the device keeps it (every core starts on a word boundary,
// sfe_engine_raster_layout gives each core's first word): one D2H copy, no host-side repacking.
extern "C" int sfe_engine_read_raster(sfe_engine *e, uint32_t *words, size_t n_words)
{
    SFE_CUDA(cudaSetDevice(e->device));
    if (n_words != e->fired_words)
    {
        sfe::set_last_error("sfe_engine_read_raster: expected " + std::to_string(e->fired_words) + " words");
        return -1;
    }
    const uint32_t *src = e->world > 1 ? e->d_fired_global : e->d_fired_local;
    if (e->p2p_on && e->sv.step_seq > 0)
    {
        // peer-memory exchange: the ranks store (word, epoch + 1) pairs straight into this rank's double-buffered raster.
        // Ranks are tied together only along their data dependencies, so a peer this rank needs nothing from may still
        // be working on the step: wait (bounded) until the words of every core carry the step's tag.
        std::vector<uint2> pairs(n_words);
        SFE_CUDA(cudaStreamSynchronize(e->stream));
        const uint32_t want = static_cast<uint32_t>(e->sv.step_seq - 1ull) + 1u;
        const uint2 *from = e->p2p_block + ((e->sv.step_seq - 1ull) & 1ull) * e->fired_words;
        for (int attempt = 0;; ++attempt)
        {
            SFE_CUDA(cudaMemcpy(pairs.data(), from, n_words * sizeof(uint2), cudaMemcpyDeviceToHost));
            bool complete = true;
            for (size_t c = 0; c < e->core_desc.size() && complete; ++c)
                for (uint32_t w = 0; w < (e->core_desc[c].neuron_count + 31u) / 32u; ++w)
                    if (pairs[e->fired_word_begin[c] + w].y != want)
                    {
                        complete = false;
                        break;
                    }
            if (complete) break;
            if (attempt > 20000)
            {
                sfe::set_last_error("sfe_engine_read_raster: a peer has not delivered its raster words of the last step");
                return -1;
            }
            std::this_thread::sleep_for(std::chrono::microseconds(100));
        }
        for (size_t k = 0; k < n_words; ++k) words[k] = pairs[k].y == want ? pairs[k].x : 0u; // (pad words are never written)
        return 0;
    }
    SFE_CUDA(cudaMemcpyAsync(words, src, n_words * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
    SFE_CUDA(cudaStreamSynchronize(e->stream));
    return 0;
}

extern "C" int sfe_engine_read_fired(sfe_engine *e, uint32_t *bits, size_t n_words)
{
    SFE_CUDA(cudaSetDevice(e->device));
    const size_t words = (static_cast<size_t>(e->n_neurons) + 31) / 32;
    if (n_words != words)
    {
        sfe::set_last_error("sfe_engine_read_fired: wrong word count");
        return -1;
    }
    std::vector<uint8_t> status(e->n_neurons);
    SFE_CUDA(cudaStreamSynchronize(e->stream));
    SFE_CUDA(cudaMemcpy(status.data(), e->s.status, e->n_neurons, cudaMemcpyDeviceToHost));
    std::memset(bits, 0, words * sizeof(uint32_t));
    for (uint32_t i = 0; i < e->n_neurons; ++i)
        if (status[i] == SFE_STATUS_FIRED) bits[i >> 5] |= 1u << (i & 31);
    return 0;
}

extern "C" int64_t sfe_engine_total_timesteps(const sfe_engine *e)
{
    return e->total_timesteps;
}
extern "C" int64_t sfe_engine_launch_count(const sfe_engine *e)
{
    return e->launches;
}
extern "C" size_t sfe_engine_device_bytes(const sfe_engine *e)
{
    return e->device_bytes;
}

// per-launch events of the message-phase kernel inside a timed region: on by default; off
// for a throughput measurement (an event between two kernels keeps the second from being
// launched early)
extern "C" int sfe_engine_time_launches(sfe_engine *e, int on)
{
    e->time_launches = on != 0;
    return 0;
}

extern "C" int sfe_engine_time_begin(sfe_engine *e)
{
    SFE_CUDA(cudaSetDevice(e->device));
    e->timing = true;
    e->ev_used = 0;
    SFE_CUDA(cudaEventRecord(e->ev_begin, e->stream));
    return 0;
}

extern "C" int sfe_engine_time_end(sfe_engine *e, float *ms_total, float *ms_fanout)
{
    SFE_CUDA(cudaSetDevice(e->device));
    flush_fold(e); // the timed region ends with the last step's record written
    SFE_CUDA(cudaEventRecord(e->ev_end, e->stream));
    SFE_CUDA(cudaEventSynchronize(e->ev_end));
    float ms = 0.f;
    SFE_CUDA(cudaEventElapsedTime(&ms, e->ev_begin, e->ev_end));
    if (ms_total != nullptr) *ms_total = ms;
    float fan = 0.f;
    for (size_t i = 0; i + 1 < e->ev_used; i += 2)
    {
        float x = 0.f;
        SFE_CUDA(cudaEventElapsedTime(&x, e->ev_pool[i], e->ev_pool[i + 1]));
        fan += x;
    }
    if (ms_fanout != nullptr) *ms_fanout = fan;
    e->timing = false;
    e->ev_used = 0;
    return 0;
}

// ---- peer-memory exchange (CUDA IPC over NVLink) ------------------------------------------
// One process per GPU: every rank exports its exchange block, the 64-byte handles travel
// through the caller's control plane (torch.distributed / MPI), every rank opens the others.
extern "C" int sfe_engine_p2p_export(sfe_engine *e, void *handle64)
{
    SFE_CUDA(cudaSetDevice(e->device));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size is part of the ABI");
    if (e->world > kMaxPeers)
    {
        sfe::set_last_error("the peer-memory exchange supports up to 8 ranks (one NVSwitch box)");
        return -1;
    }
    if (e->p2p_block == nullptr)
        if (e->alloc(&e->p2p_block, 2 * static_cast<size_t>(e->fired_words) + kMaxPeers / 2) != 0) return -1; // pairs + done[8]
    SFE_CUDA(cudaStreamSynchronize(e->stream));
    cudaIpcMemHandle_t h;
    SFE_CUDA(cudaIpcGetMemHandle(&h, e->p2p_block));
    std::memcpy(handle64, &h, sizeof(h));
    return 0;
}

extern "C" int sfe_engine_p2p_attach(sfe_engine *e, const void *handles)
{
    SFE_CUDA(cudaSetDevice(e->device));
    if (e->p2p_block == nullptr)
    {
        sfe::set_last_error("sfe_engine_p2p_attach: call sfe_engine_p2p_export first");
        return -1;
    }
    if (e->soma_list.empty())
    {
        sfe::set_last_error("sfe_engine_p2p_attach: every rank of the peer-memory exchange must own neurons (its neuron phase reports "
                            "the rank's progress to the peers)");
        return -1;
    }
    SFE_CUDA(cudaStreamSynchronize(e->stream));
    for (uint32_t q = 0; q < e->world; ++q)
    {
        uint2 *base = e->p2p_block;
        if (q != e->rank)
        {
            cudaIpcMemHandle_t h;
            std::memcpy(&h, static_cast<const unsigned char *>(handles) + 64 * static_cast<size_t>(q), sizeof(h));
            void *opened = nullptr;
            SFE_CUDA(cudaIpcOpenMemHandle(&opened, h, cudaIpcMemLazyEnablePeerAccess));
            e->p2p_peer_base[q] = opened;
            base = static_cast<uint2 *>(opened);
        }
        e->s.x.ll[q] = base;
        e->s.x.done[q] = reinterpret_cast<uint32_t *>(base + 2 * static_cast<size_t>(e->fired_words));
    }
    e->s.x.n_peers = e->world;
    e->s.x.rank = e->rank;
    e->s.x.fired_words = e->fired_words;
    e->s.x.slice_words = e->slice_words;
    e->p2p_on = true;
    return 0;
}

// Callers synchronise all ranks (host barrier) before detaching: a peer may still be storing.
extern "C" int sfe_engine_p2p_detach(sfe_engine *e)
{
    SFE_CUDA(cudaSetDevice(e->device));
    SFE_CUDA(cudaStreamSynchronize(e->stream));
    for (uint32_t q = 0; q < kMaxPeers; ++q)
        if (e->p2p_peer_base[q] != nullptr)
        {
            cudaIpcCloseMemHandle(e->p2p_peer_base[q]);
            e->p2p_peer_base[q] = nullptr;
        }
    e->s.x.n_peers = 0;
    e->p2p_on = false;
    return 0;
}

// diagnostic: the %globaltimer stamps the CTAs of the fused step kernel left for the last 64 steps (SFE_TIMELINE=1):
// out[64][grid][16], returns the grid size (0: not recorded)
extern "C" int sfe_engine_read_timeline(sfe_engine *e, unsigned long long *out, size_t cap_words)
{
    SFE_CUDA(cudaSetDevice(e->device));
    if (e->s.timeline == nullptr) return 0;
    const size_t words = 64ull * e->fanout_grid * 16ull;
    if (cap_words < words) return -1;
    SFE_CUDA(cudaStreamSynchronize(e->stream));
    SFE_CUDA(cudaMemcpy(out, e->s.timeline, words * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return static_cast<int>(e->fanout_grid);
}

// diagnostic build (SFE_TIMELINE_ALL): both stamp blocks as they are, [2][64][1024][16]; returns the words copied
extern "C" int64_t sfe_engine_read_timeline_raw(sfe_engine *e, unsigned long long *out, size_t cap_words)
{
    if (cudaSetDevice(e->device) != cudaSuccess) return -1;
    if (e->s.timeline == nullptr || cap_words < 2 * kTimelineWords) return 0;
    if (cudaStreamSynchronize(e->stream) != cudaSuccess) return -1;
    if (cudaMemcpy(out, e->s.timeline, 2 * kTimelineWords * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return static_cast<int64_t>(2 * kTimelineWords);
}

// 0 = fine, 1 = a peer failed to arrive within the in-kernel time limit (results invalid)
extern "C" int sfe_engine_exchange_error(sfe_engine *e)
{
    SFE_CUDA(cudaSetDevice(e->device));
    uint32_t v[4] = {0, 0, 0, 0};
    SFE_CUDA(cudaStreamSynchronize(e->stream));
    SFE_CUDA(cudaMemcpy(v, e->s.x.error, sizeof(v), cudaMemcpyDeviceToHost));
    if (v[0] == 2u)
        sfe::set_last_error("raster exchange: rank " + std::to_string(v[1]) + " did not complete step " + std::to_string(v[2] >= 2 ? v[2] - 2 : 0) +
                " in time (rank " + std::to_string(e->rank) + " holds its words back; the peer reported " + std::to_string(v[3]) + " completed steps)");
    else if (v[0] != 0u)
        sfe::set_last_error("raster exchange: word " + std::to_string(v[1]) + " (slice of rank " +
                std::to_string(e->slice_words != 0 ? v[1] / e->slice_words : 0) + ") did not arrive at rank " + std::to_string(e->rank) +
                " in time: expected epoch tag " + std::to_string(v[2]) + ", found " + std::to_string(v[3]));
    return static_cast<int>(v[0]);
}

// ---- multi-GPU: one engine per rank, spikes exchanged as a fired-bit raster ---------------
// step = enqueue_neuron_phase -> all-gather(local slice -> global raster) -> enqueue_message_phase
extern "C" int sfe_engine_enqueue_neuron_phase(sfe_engine *e)
{
    SFE_CUDA(cudaSetDevice(e->device));
    if (e->total_timesteps - e->log_read + 1 > e->log_cap)
    {
        sfe::set_last_error("more than " + std::to_string(e->log_cap) + " uncollected steps; call sfe_engine_collect_records");
        return -1;
    }
    if (check_overlay(e) != 0) return -1;
    if (apply_pending_bias(e) != 0) return -1;
    if (!e->tickets_prepared && prepare_tickets(e, 1) != 0) return -1;
    if (!e->soma_list.empty())
    {
        launch_soma(e);
        ++e->launches;
    }
    SFE_CUDA(cudaGetLastError());
    return 0;
}

// SFE_PHASE_PROFILE diagnostic: events dropped between the kernels of a partitioned step
static std::vector<cudaEvent_t> g_prof_events;
static bool g_prof_on = false;
static void prof_mark(sfe_engine *e)
{
    if (!g_prof_on) return;
    cudaEvent_t ev;
    cudaEventCreate(&ev);
    cudaEventRecord(ev, e->stream);
    g_prof_events.push_back(ev);
}

extern "C" int sfe_engine_enqueue_message_phase(sfe_engine *e)
{
    SFE_CUDA(cudaSetDevice(e->device));
    // with the peer-memory exchange every rank launches the message phase, work or not: its
    // first CTA publishes this rank's raster slice and all ranks stay within one step of each
    // other (which is what the double-buffered raster assumes)
    if (!e->fanout_list.empty() || e->p2p_on)
    {
        launch_fanout(e);
        ++e->launches;
    }
    prof_mark(e);
    launch_finalize(e);
    ++e->total_timesteps;
    SFE_CUDA(cudaGetLastError());
    return 0;
}

extern "C" void *sfe_engine_fired_local_ptr(sfe_engine *e, size_t *n_bytes)
{
    if (n_bytes != nullptr) *n_bytes = static_cast<size_t>(e->slice_words) * sizeof(uint32_t);
    return e->d_fired_local;
}

extern "C" void *sfe_engine_fired_global_ptr(sfe_engine *e, size_t *n_bytes)
{
    if (n_bytes != nullptr) *n_bytes = static_cast<size_t>(e->fired_words) * sizeof(uint32_t);
    return e->d_fired_global;
}

// Let the caller own the exchange buffers (e.g. torch tensors handed to NCCL):
// local = slice_words u32, global = world * slice_words u32, both device memory.
extern "C" int sfe_engine_set_exchange_buffers(sfe_engine *e, void *local, void *global)
{
    SFE_CUDA(cudaSetDevice(e->device));
    SFE_CUDA(cudaStreamSynchronize(e->stream));
    e->d_fired_local = static_cast<uint32_t *>(local);
    e->d_fired_global = static_cast<uint32_t *>(global);
    e->s.fired_bits = e->d_fired_local - static_cast<size_t>(e->rank) * e->slice_words * (e->world > 1 ? 1 : 0);
    e->s.fired_global = e->d_fired_global;
    e->external_exchange = true;
    return 0;
}

// per-step records of this rank's partition since the last collect (counts / energies are
// partial sums over the local cores, sim_time the local maximum): the caller sums the former
// over ranks and takes the maximum of the latter, step by step
extern "C" int64_t sfe_engine_collect_records(sfe_engine *e, sfe_step_record *out, int64_t cap)
{
    if (cudaSetDevice(e->device) != cudaSuccess) return -1;
    const int64_t pending = e->total_timesteps - e->log_read;
    if (pending > cap)
    {
        sfe::set_last_error("sfe_engine_collect_records: buffer too small");
        return -1;
    }
    std::vector<sfe_step_record> recs;
    if (collect_records(e, recs) != 0) return -1;
    std::memcpy(out, recs.data(), recs.size() * sizeof(sfe_step_record));
    return pending;
}

extern "C" int sfe_engine_partition_info(const sfe_engine *e, uint32_t *rank, uint32_t *world, uint32_t *slice_words,
        uint32_t *local_cores, uint64_t *local_neurons)
{
    if (rank != nullptr) *rank = e->rank;
    if (world != nullptr) *world = e->world;
    if (slice_words != nullptr) *slice_words = e->slice_words;
    if (local_cores != nullptr) *local_cores = static_cast<uint32_t>(e->soma_list.size());
    if (local_neurons != nullptr)
    {
        uint64_t n = 0;
        for (uint32_t c : e->soma_list) n += e->core_desc[c].neuron_count;
        *local_neurons = n;
    }
    return 0;
}

extern "C" int sfe_engine_synchronize(sfe_engine *e)
{
    SFE_CUDA(cudaSetDevice(e->device));
    flush_fold(e);
    SFE_CUDA(cudaStreamSynchronize(e->stream));
    return 0;
}

// plain cudaMemcpy (any direction, unified addressing); for tests and small exchanges
// Copy between any two of host / device memory, complete on return. (cudaMemcpy alone does not
// wait for a device-to-device copy, and the engines' streams are non-blocking: without the
// synchronisation a kernel enqueued right after could read the destination too early.)
extern "C" int sfe_device_memcpy(void *dst, const void *src, size_t bytes)
{
    SFE_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDefault));
    SFE_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
    return 0;
}

// first raster word of every core in the (global) fired raster
extern "C" int sfe_engine_raster_layout(const sfe_engine *e, uint32_t *word_begin, size_t n_cores)
{
    if (n_cores != e->fired_word_begin.size())
    {
        sfe::set_last_error("sfe_engine_raster_layout: expected one slot per core");
        return -1;
    }
    std::memcpy(word_begin, e->fired_word_begin.data(), n_cores * sizeof(uint32_t));
    return 0;
}

// pinned host memory for buffers that cross the ABI every step (bias vectors, rasters)
extern "C" void *sfe_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess)
    {
        cudaGetLastError();
        sfe::set_last_error("sfe_host_alloc: cudaMallocHost failed");
        return nullptr;
    }
    return p;
}

extern "C" void sfe_host_free(void *p)
{
    if (p != nullptr) cudaFreeHost(p);
}

// The last `n` step records of the device log, by the DEVICE cursor. For steps that
// were replayed from a CUDA graph (they bypass the host-side counters); re-synchronises
// the host counters with the device.
extern "C" int64_t sfe_engine_read_log_tail(sfe_engine *e, sfe_step_record *out, int64_t n)
{
    if (cudaSetDevice(e->device) != cudaSuccess) return -1;
    flush_fold(e);
    if (cudaStreamSynchronize(e->stream) != cudaSuccess) return -1;
    long long counters[2] = {0, 0};
    if (cudaMemcpy(counters, e->s.step, sizeof(counters), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    const int64_t cursor = counters[1];
    if (n > cursor || n > e->log_cap)
    {
        sfe::set_last_error("sfe_engine_read_log_tail: not that many records in the device log");
        return -1;
    }
    for (int64_t i = 0; i < n; ++i)
    {
        const int64_t pos = (cursor - n + i) % e->log_cap;
        if (cudaMemcpy(out + i, e->s.log + pos, sizeof(sfe_step_record), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    }
    e->total_timesteps = counters[0];
    e->log_read = cursor;
    return n;
}

// ===========================================================================
// NCCL exchange driven from C++ (no Python in the per-step loop)
// ===========================================================================
// libnccl is loaded at run time (dlopen) so the single-GPU library has no NCCL
// dependency. Only four entry points are used; their signatures and the enum
// values below are those of NCCL 2.x's public nccl.h.
#include <dlfcn.h>

namespace
{
constexpr int kNcclUniqueIdBytes = 128;
struct NcclUniqueId
{
    char internal[kNcclUniqueIdBytes];
};
constexpr int kNcclUint32 = 3; // ncclDataType_t::ncclUint32
using NcclComm = void *;
struct NcclApi
{
    void *lib{nullptr};
    int (*GetUniqueId)(NcclUniqueId *){nullptr};
    int (*CommInitRank)(NcclComm *, int, NcclUniqueId, int){nullptr};
    int (*AllGather)(const void *, void *, size_t, int, NcclComm, cudaStream_t){nullptr};
    int (*CommDestroy)(NcclComm){nullptr};
    const char *(*GetErrorString)(int){nullptr};
};
NcclApi g_nccl;

int nccl_load(const char *path)
{
    if (g_nccl.lib != nullptr) return 0;
    const char *candidates[] = {path, std::getenv("SFE_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *c : candidates)
    {
        if (c == nullptr || *c == '\0') continue;
        g_nccl.lib = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib != nullptr) break;
    }
    if (g_nccl.lib == nullptr)
    {
        sfe::set_last_error("could not load libnccl.so.2 (pass its path or set SFE_NCCL_LIB)");
        return -1;
    }
    g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(dlsym(g_nccl.lib, "ncclGetUniqueId"));
    g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(dlsym(g_nccl.lib, "ncclCommInitRank"));
    g_nccl.AllGather = reinterpret_cast<decltype(g_nccl.AllGather)>(dlsym(g_nccl.lib, "ncclAllGather"));
    g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(dlsym(g_nccl.lib, "ncclCommDestroy"));
    g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(dlsym(g_nccl.lib, "ncclGetErrorString"));
    if (g_nccl.GetUniqueId == nullptr || g_nccl.CommInitRank == nullptr || g_nccl.AllGather == nullptr ||
            g_nccl.CommDestroy == nullptr || g_nccl.GetErrorString == nullptr)
    {
        sfe::set_last_error("libnccl lacks an expected symbol");
        return -1;
    }
    return 0;
}
std::vector<std::pair<sfe_engine *, NcclComm>> g_comms; // engine -> communicator

NcclComm comm_of(sfe_engine *e)
{
    for (auto &kv : g_comms)
        if (kv.first == e) return kv.second;
    return nullptr;
}
} // namespace

// rank 0 creates the id (128 bytes) and ships it to the other ranks (torch.distributed / any channel)
extern "C" int sfe_nccl_get_unique_id(void *out128, const char *libnccl_path)
{
    if (nccl_load(libnccl_path) != 0) return -1;
    NcclUniqueId id;
    const int rc = g_nccl.GetUniqueId(&id);
    if (rc != 0)
    {
        sfe::set_last_error(std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(rc));
        return -1;
    }
    std::memcpy(out128, id.internal, kNcclUniqueIdBytes);
    return 0;
}

extern "C" int sfe_engine_comm_init(sfe_engine *e, const void *unique_id128, const char *libnccl_path)
{
    SFE_CUDA(cudaSetDevice(e->device));
    if (nccl_load(libnccl_path) != 0) return -1;
    NcclUniqueId id;
    std::memcpy(id.internal, unique_id128, kNcclUniqueIdBytes);
    NcclComm comm = nullptr;
    const int rc = g_nccl.CommInitRank(&comm, static_cast<int>(e->world), id, static_cast<int>(e->rank));
    if (rc != 0)
    {
        sfe::set_last_error(std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(rc));
        return -1;
    }
    g_comms.emplace_back(e, comm);
    return 0;
}

extern "C" int sfe_engine_comm_destroy(sfe_engine *e)
{
    for (size_t i = 0; i < g_comms.size(); ++i)
        if (g_comms[i].first == e)
        {
            g_nccl.CommDestroy(g_comms[i].second);
            g_comms.erase(g_comms.begin() + static_cast<long>(i));
            break;
        }
    return 0;
}

// `timesteps` full steps of a partitioned chip, enqueued back to back on the engine's
// stream: neuron phase -> ncclAllGather(raster slices) over NVLink -> expand + message phase
extern "C" int sfe_engine_enqueue_partitioned(sfe_engine *e, int64_t timesteps)
{
    SFE_CUDA(cudaSetDevice(e->device));
    NcclComm comm = comm_of(e);
    if (e->world > 1 && comm == nullptr && !e->p2p_on)
    {
        sfe::set_last_error("sfe_engine_enqueue_partitioned: call sfe_engine_p2p_attach or sfe_engine_comm_init first");
        return -1;
    }
    if (e->total_timesteps - e->log_read + timesteps > e->log_cap)
    {
        sfe::set_last_error("more than " + std::to_string(e->log_cap) + " uncollected steps; collect first");
        return -1;
    }
    // SFE_PHASE_PROFILE=1: bracket every kernel with events and print the mean device time of
    // each (diagnostic only)
    static const bool profile = std::getenv("SFE_PHASE_PROFILE") != nullptr;
    const int64_t prof_steps = profile ? std::min<int64_t>(timesteps, 128) : 0;
    if (prepare_tickets(e, timesteps) != 0) return -1;
    e->tickets_prepared = true;
    for (int64_t s = 0; s < timesteps; ++s)
    {
        g_prof_on = s < prof_steps;
        prof_mark(e);
        if (sfe_engine_enqueue_neuron_phase(e) != 0) return -1;
        prof_mark(e);
        if (e->world > 1 && !e->p2p_on)
        {
            const int rc = g_nccl.AllGather(e->d_fired_local, e->d_fired_global, e->slice_words, kNcclUint32, comm, e->stream);
            if (rc != 0)
            {
                sfe::set_last_error(std::string("ncclAllGather: ") + g_nccl.GetErrorString(rc));
                return -1;
            }
            ++e->launches;
        }
        prof_mark(e);
        if (sfe_engine_enqueue_message_phase(e) != 0) return -1; // marks after expand and after fanout
        prof_mark(e);
    }
    g_prof_on = false;
    e->tickets_prepared = false;
    if (prof_steps > 0)
    {
        SFE_CUDA(cudaStreamSynchronize(e->stream));
        constexpr int kMarks = 5;
        double acc[kMarks] = {0, 0, 0, 0, 0};
        for (int64_t s = 0; s < prof_steps; ++s)
            for (int k = 0; k < kMarks; ++k)
            {
                const size_t a = static_cast<size_t>(s) * kMarks + k;
                if (a + 1 >= g_prof_events.size()) break;
                float ms = 0.f;
                cudaEventElapsedTime(&ms, g_prof_events[a], g_prof_events[a + 1]);
                acc[k] += ms;
            }
        const double d = 1e3 / static_cast<double>(prof_steps);
        std::fprintf(stderr, "[sfe phase profile] rank %d: soma %.1f us, exchange %.1f us, fanout %.1f us, "
                "finalize %.1f us, gap %.1f us (mean of %lld steps)\n",
                e->rank, acc[0] * d, acc[1] * d, acc[2] * d, acc[3] * d, acc[4] * d, static_cast<long long>(prof_steps));
        for (cudaEvent_t ev : g_prof_events) cudaEventDestroy(ev);
        g_prof_events.clear();
    }
    return 0;
}
