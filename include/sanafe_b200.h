/* sanafe_b200.h — C ABI of the B200-native SANA-FE time-step engine.
 *
 * Drop-in boundary for ONE path of the reference: SpikingChip::load / sim /
 * step / sim_hw_timestep and everything they call (reference src/chip.cpp,
 * pipeline.*, models.cpp, schedule.cpp "simple"). Plain pointers and sizes
 * only; no C++ or torch types cross this boundary; every function returns an
 * int status (0 = ok) or a handle/NULL and leaves a message for sfe_last_error().
 * No exception crosses the ABI.
 *
 * Two levels:
 *   description level  sfe_arch_* / sfe_net_* / sfe_chip_*   mirrors the reference's
 *                      Architecture / SpikingNetwork / SpikingChip surface
 *   table level        sfe_tables + sfe_engine_*             the lowered SoA/CSR form the
 *                      device kernels consume (also what oracle/sfe_oracle.c consumes)
 */
#ifndef SANAFE_B200_H_
#define SANAFE_B200_H_

#include <stddef.h>
#include <stdint.h>

#include "sfe_synth.h"

#ifdef __cplusplus
extern "C" {
#endif

#define SFE_ABI_VERSION 7

/* ---- enums (values follow the reference where it has them) -------------- */
/* src/arch.hpp:41-49 BufferPosition */
enum
{
    SFE_BUF_BEFORE_DENDRITE = 0,
    SFE_BUF_INSIDE_DENDRITE = 1,
    SFE_BUF_BEFORE_SOMA = 2,
    SFE_BUF_INSIDE_SOMA = 3,
    SFE_BUF_BEFORE_AXON_OUT = 4
};
/* src/arch.hpp:61-68 NeuronResetModes */
enum
{
    SFE_RESET_NONE = 0,
    SFE_RESET_SOFT = 1,
    SFE_RESET_HARD = 2,
    SFE_RESET_SATURATE = 3
};
/* src/mapped.hpp:22-28 NeuronStatus */
enum
{
    SFE_STATUS_UNSET = 0,
    SFE_STATUS_IDLE = 1,
    SFE_STATUS_UPDATED = 2,
    SFE_STATUS_FIRED = 3
};
/* src/chip.hpp:40-45 TimingModel */
enum
{
    SFE_TIMING_SIMPLE = 0,
    SFE_TIMING_DETAILED = 1,
    SFE_TIMING_CYCLE = 2
};
/* soma models: src/models.cpp:933-967 + plugins/hodgkin_huxley.cpp */
enum
{
    SFE_SOMA_LIF = 0,       /* "leaky_integrate_fire"   src/models.cpp:497-567 */
    SFE_SOMA_TRUENORTH = 1, /* "truenorth"              src/models.cpp:799-830 */
    SFE_SOMA_INPUT = 2,     /* "input"                  src/models.cpp:863-903 */
    SFE_SOMA_HH = 3,        /* "hodgkin_huxley" plugin  plugins/hodgkin_huxley.cpp:116-170 */
    SFE_SOMA_NEUROFEM = 4,  /* "neurofem" plugin (combined dendrite + soma unit)  plugins/neurofem.cpp:192-317, sigma_v = 0 */
    SFE_SOMA_DEVICE_MODEL = 5 /* a soma model registered with sfe_register_device_model: its update runs as a kernel of
                               * its own device code image before the neuron phase, see sfe_device_model.h */
};
/* dendrite models */
enum
{
    SFE_DEND_ACCUMULATOR = 0,      /* src/models.cpp:71-94  */
    SFE_DEND_ACCUMULATOR_DELAY = 1,/* src/models.cpp:96-131 */
    SFE_DEND_TAPS = 2,             /* "taps" MultiTapModel1D  src/models.cpp:167-259 */
    SFE_DEND_NEUROFEM = 3          /* dendrite half of the combined "neurofem" unit: two accumulators per neuron (u1, u2),
                                    * double-buffered over the timestep; a synapse names its compartment (0/1) in the
                                    * delay field of syn_meta  plugins/neurofem.cpp:120-136, 227-248 */
};
/* how one core's message phase accumulates synaptic charge */
enum
{
    SFE_ACC_PACKED32 = 0, /* exact fixed point: 20-bit sum + 12-bit count in one u32 smem atomic */
    SFE_ACC_DUAL32 = 1,   /* exact fixed point: int32 sum + u32 count, two smem atomics */
    SFE_ACC_ORDERED = 2,  /* sequential fp64 adds in message arrival order (any weights) */
    SFE_ACC_PACKED17 = 3  /* exact fixed point: cell = sum + count * 2^17 (|sum| < 2^16, count < 2^15): one u32 smem atomic
                           * whose addend is a plain shift of the 4-byte synapse record (see syn_q4 in engine.cu) */
};
/* sfe_soma_class.flags */
#define SFE_SOMA_FORCE_UPDATE 1u
#define SFE_SOMA_LEAK_TOWARDS_ZERO 2u
#define SFE_SOMA_LOG_U 4u
#define SFE_SOMA_NOISE 8u /* LIF unit with a file noise stream: neuron_aux indexes sfe_tables.noise */
#define SFE_SOMA_IS_DENDRITE 16u /* the soma unit is also the neuron's dendrite unit (a combined unit): its energy counts in the
                                  * dendrite AND the soma bucket of a step, once in the total (src/chip.cpp:1224-1245) */

/* ---- lowered tables ------------------------------------------------------ */
typedef struct sfe_tile_desc
{
    uint32_t x, y; /* x = id / noc_height, y = id % noc_height  src/arch.cpp:78-104 */
    double energy_east, energy_west, energy_south, energy_north;
    double latency_east, latency_west, latency_south, latency_north;
} sfe_tile_desc;

typedef struct sfe_core_desc
{
    uint32_t id;          /* global core id (creation order) */
    uint32_t tile;
    uint32_t offset;      /* offset within tile */
    uint32_t buffer_pos;  /* SFE_BUF_* */
    uint32_t neuron_begin, neuron_count;   /* neurons in in-core mapped order */
    uint32_t axon_in_begin, axon_in_count; /* axons-in in creation (= arrival) order */
    uint64_t syn_begin;   /* first synapse of this core in the synapse arrays */
    uint64_t syn_count;
    uint32_t acc_mode;    /* SFE_ACC_* */
    int32_t weight_shift; /* exact modes: every weight is k * 2^-weight_shift */
    uint32_t ring;        /* dendrite ring slots = max synaptic delay + 1 */
    uint32_t dend_in_msg; /* 1: dendrite unit is in the message pipeline (buffer_pos >= 1) */
    uint32_t fixed_slots; /* 1: the delay field of a synapse IS its dendrite slot (compartment of a "neurofem" unit); the
                           * neuron phase reads slots 0..ring-1 of the step instead of the rotating slot of a delay line */
    uint32_t pad;
    /* axon units with the reference's quirks folded in:
     * latency: axon_in_hw[0] / axon_out_hw[0]; energy: only when the core has
     * exactly one such unit (the reference counts the LAST unit's energy while
     * crediting activity to unit 0). */
    double energy_axon_in, latency_axon_in;
    double energy_axon_out, latency_axon_out;
} sfe_core_desc;

typedef struct sfe_soma_class
{
    uint32_t model;         /* SFE_SOMA_* */
    uint32_t reset_mode;    /* SFE_RESET_* */
    uint32_t reverse_reset_mode;
    int32_t refractory_delay;
    uint32_t flags;         /* SFE_SOMA_* flags */
    uint32_t random_mask;   /* truenorth only: threshold jitter `rand() & random_mask` (src/models.cpp:749-759). Non-zero: the
                             * neuron's neuron_aux is its column in the per-step overlay of glibc rand() values
                             * (sfe_engine_set_rand_overlay; sfe_chip_sim draws them, one per such neuron and timestep in chip
                             * order - what the reference's process-global std::rand() yields with one processing thread) */
    uint32_t dend_model;    /* SFE_DEND_* of the neuron's dendrite unit */
    uint32_t dend_in_neuron;/* 1: dendrite unit is in the neuron pipeline (buffer_pos <= 1) */
    double threshold, reverse_threshold, reset, reverse_reset;
    double leak;            /* LIF leak_decay | truenorth leak */
    double input_decay;
    /* default costs of the soma unit (src/pipeline.hpp:631-714) */
    double energy_access, energy_update, energy_spike_out;
    double latency_access, latency_update, latency_spike_out;
    /* default costs of the dendrite unit when it runs in the neuron pipeline */
    double dend_energy_update, dend_latency_update;
    /* "neurofem" (plugins/neurofem.cpp:64-84): lambda_v rides in `leak`, lambda_d in `input_decay`; kp, ki, dt here */
    double nf_kp, nf_ki, nf_dt;
} sfe_soma_class;

/* per-event default costs of one (synapse unit, dendrite unit) pairing; an axon
 * whose synapses mix units gets a class with per_message = 1 holding totals */
typedef struct sfe_cost_class
{
    double syn_energy, syn_latency; /* src/pipeline.hpp:511-572 */
    double den_energy, den_latency; /* src/pipeline.hpp:574-629 (0 if dendrite not in message pipeline) */
    uint32_t per_message;
    uint32_t den_is_soma; /* 1: the "dendrite" of the pairing is a combined dendrite + soma unit: every event costs it one soma
                           * access (src/pipeline.hpp:631-667 with status unset), counted in the dendrite and soma buckets */
} sfe_cost_class;

/* one axon-in = one (pre-neuron, destination core) pair = one message per spike */
typedef struct sfe_axon_in
{
    uint32_t syn_off;    /* relative to core.syn_begin */
    uint32_t syn_count;  /* = Message.spikes  src/message.cpp:52 */
    uint32_t hop;        /* dx | dy<<12 | east<<24 | north<<25   src/chip.cpp:1127-1169 */
    uint32_t cost_class;
} sfe_axon_in;
#define SFE_HOP_DX(h) ((h) & 0xfffu)
#define SFE_HOP_DY(h) (((h) >> 12) & 0xfffu)
#define SFE_HOP_EAST(h) (((h) >> 24) & 1u)
#define SFE_HOP_NORTH(h) (((h) >> 25) & 1u)

/* synapse meta word: post-neuron offset within core | delay << 16 */
#define SFE_SYN_POST(m) ((m) & 0xffffu)
#define SFE_SYN_DELAY(m) (((m) >> 16) & 0x7u)
#define SFE_SYN_TAP(m) ((m) >> 16) /* "taps" dendrites: the field holds the tap index instead of a delay */

typedef struct sfe_input_desc /* "input" soma: state is per hardware UNIT (src/models.cpp:832-903) */
{
    uint32_t spikes_off, spikes_len; /* into sfe_tables.input_spikes (1 byte per step) */
    uint32_t share_count, share_rank;/* neurons sharing the unit, this neuron's rank among them */
    double rate;
    double poisson;                  /* spike when poisson > U(0,1) drawn from the UNIT's std::mt19937 (src/models.cpp:880-885) */
    uint32_t unit;                   /* ordinal of the unit among the chip's "input" units in hardware order: its generator is
                                      * seeded with input_seed_base + unit + 1 (InputModel::instance_counter, src/models.hpp:347,366) */
    uint32_t poisson_col;            /* column of this neuron in the per-step Poisson overlay, or 0xFFFFFFFF (poisson == 0) */
} sfe_input_desc;

/* LIF file noise (LoihiLifModel::loihi_generate_noise, src/models.cpp:589-650): the unit reads one entry of its
 * file per update of any of its compartments and rewinds at the end of the file. The entries, already masked and
 * sign-extended, sit in sfe_tables.noise_values[off .. off+len); at step index s the neuron of rank share_rank among
 * the share_count neurons of the unit consumes entry (s * share_count + share_rank) mod len. */
typedef struct sfe_noise_desc
{
    uint32_t off, len;
    uint32_t share_count, share_rank;
} sfe_noise_desc;

/* "taps" dendrite (MultiTapModel1D, src/models.cpp:167-259): a 1-D RC line of n_taps compartments, ONE line per
 * hardware unit whatever neuron is updated (neurons mapped to the same unit share their descriptor). taps_values[const_off .. +n_taps) are the time constants, the next n_taps-1 values the space constants.
 * A synapse names its tap in the delay field of syn_meta. */
typedef struct sfe_taps_desc
{
    uint32_t n_taps;
    uint32_t const_off;
} sfe_taps_desc;

typedef struct sfe_hh_init /* Hodgkin-Huxley plugin initial state, one neuron per unit */
{
    double m, n, h, current;
} sfe_hh_init;

/* Out-of-tree soma device models (sfe_device_model.h): one block per model the chip uses. Every neuron whose soma
 * class has model == SFE_SOMA_DEVICE_MODEL is one instance; its neuron_aux is its index in the concatenation of all
 * blocks (block b covers [sum of n_instances before b, +n_instances)). */
struct sfe_device_model_desc;
typedef struct sfe_device_model_block
{
    const struct sfe_device_model_desc *desc;
    uint32_t n_instances;
    uint32_t pad;
    const double *state_init; /* [desc->n_state][n_instances] */
    const double *params;     /* [desc->n_params][n_instances] */
    const uint32_t *neurons;  /* [n_instances] device index of every instance */
} sfe_device_model_block;

typedef struct sfe_tables
{
    uint32_t abi_version;
    uint32_t noc_width, noc_height, noc_buffer_size, max_cores_per_tile;
    uint32_t n_tiles, n_cores, n_neurons, n_axons_in, n_soma_classes, n_cost_classes;
    uint32_t n_inputs, n_hh, n_probes;
    uint32_t mapped_tiles, mapped_cores;
    uint64_t n_synapses, n_axons_out;
    double sync_delay; /* LookupTable::get(mapped_tiles)  src/chip.cpp:562-574, src/utils.hpp:19-45 */

    const sfe_tile_desc *tiles;
    const sfe_core_desc *cores;
    const sfe_soma_class *soma_classes;
    const sfe_cost_class *cost_classes;

    /* neurons, indexed by device index (cores in id order, in-core mapped order) */
    const uint32_t *neuron_class;
    const uint32_t *neuron_aux;      /* index into inputs / hh (by model), else 0 */
    const double *neuron_bias;
    const double *neuron_potential0; /* initial potential (LIF attribute "potential") */
    const uint32_t *axon_out_begin;  /* n_neurons + 1 */
    const uint32_t *axon_out_target; /* n_axons_out: global axon-in ids, in the order messages are sent */
    const sfe_input_desc *inputs;
    const uint8_t *input_spikes;
    uint64_t n_input_spikes;
    const sfe_hh_init *hh;
    const uint32_t *probes;          /* device indices whose potential is traced each step */

    /* axons-in, grouped by destination core in arrival order */
    const sfe_axon_in *axons_in;
    const uint32_t *axon_src;        /* source neuron (device index) of each axon-in */

    /* synapses, grouped by destination core, axon, creation order. NULL when the
     * tables describe a synthetic network that is generated on the device. */
    const double *syn_weight;
    const uint32_t *syn_meta;
    const sfe_synth_spec *synth;     /* non-NULL: generate synapses from this spec */

    /* Poisson inputs: "input" units created in this process before this chip (the reference's counter is a
     * process-wide static) and the number of neurons with poisson > 0 (= columns of the overlay) */
    uint32_t input_seed_base, n_poisson_cols;

    /* LIF file noise streams */
    const sfe_noise_desc *noise;
    const double *noise_values;
    uint64_t n_noise_values;
    uint32_t n_noise;
    /* model-defined neuron traces (src/chip.cpp:1478-1517, 1664-1702): device indices of the LIF neurons with
     * log_u, whose input current `u` is traced each step, in trace order */
    uint32_t n_u_probes;
    const uint32_t *u_probes;

    /* "taps" dendrites: per neuron the index of its sfe_taps_desc (0xFFFFFFFF: none); NULL when n_taps_units == 0 */
    const uint32_t *neuron_taps;
    const sfe_taps_desc *taps;
    const double *taps_values;
    uint32_t n_taps_units, n_taps_values;

    /* out-of-tree soma device models used by this chip */
    const sfe_device_model_block *device_models;
    uint32_t n_device_models, n_device_instances;

    /* TrueNorth neurons with random_mask != 0 (= columns of the rand() overlay) */
    uint32_t n_rand_cols, pad_rand;
} sfe_tables;

/* one record per simulated timestep (src/timestep.hpp:21-42) */
typedef struct sfe_step_record
{
    int64_t neurons_fired, neurons_updated, packets_sent, total_hops, spike_count;
    double sim_time, synapse_energy, dendrite_energy, soma_energy, network_energy, total_energy;
} sfe_step_record;

/* src/chip.hpp:215-233 RunData */
typedef struct sfe_run_data
{
    int64_t timestep_start, timesteps_executed;
    int64_t spikes, packets_sent, neurons_updated, neurons_fired;
    double total_energy, synapse_energy, dendrite_energy, soma_energy, network_energy;
    double sim_time;
    double wall_time;          /* host wall-clock seconds of the sim() call */
    double scheduler_wall_time;/* of which: host-side detailed scheduler (0 for the simple model) */
} sfe_run_data;

/* What sim() should hand back besides the totals. Any pointer may be NULL.
 * Buffers are caller-owned HOST memory. */
typedef struct sfe_trace_request
{
    sfe_step_record *steps;   /* [timesteps] per-step records */
    uint32_t *fired_bits;     /* [timesteps][ceil(n_neurons/32)] fired bitmask per step, device index order */
    double *potentials;       /* [timesteps][n_probes] */
    uint8_t *status;          /* [timesteps][n_neurons] SFE_STATUS_* (detailed-timing feed / debugging) */
    double *neuron_traces;    /* [timesteps][n_u_probes] model-defined traces (LIF `u` of the log_u neurons) */
} sfe_trace_request;

const char *sfe_last_error(void);
/* class of the C++ exception behind sfe_last_error() (the reference reports errors as exceptions) */
enum { SFE_ERROR_RUNTIME = 0, SFE_ERROR_INVALID_ARGUMENT = 1, SFE_ERROR_OUT_OF_RANGE = 2, SFE_ERROR_HARDWARE_MAPPING = 3 };
int sfe_last_error_kind(void);
int sfe_abi_version(void);
/* number of CUDA devices visible; 0 when there is no GPU (never a CPU fallback) */
int sfe_device_count(void);

/* ---- table level --------------------------------------------------------- */
typedef struct sfe_engine sfe_engine;

/* Copies the tables to device `device` (synapses generated there when
 * tables->synth is set) and zero-initialises the state. Fails when no CUDA
 * device is present. part_rank/part_count: this engine simulates the cores
 * whose index satisfies the contiguous core partition of rank part_rank. */
sfe_engine *sfe_engine_create(const sfe_tables *tables, int device);
void sfe_engine_destroy(sfe_engine *e);
/* `stream` is a cudaStream_t (0 = the engine's own stream) */
int sfe_engine_set_stream(sfe_engine *e, void *stream);
/* Runs `timesteps` steps. Traces as requested. Blocks until results are on the host. */
int sfe_engine_run(sfe_engine *e, int64_t timesteps, const sfe_trace_request *req, sfe_run_data *out);
/* Enqueue steps without any host synchronisation or read-back (benchmark / graph use). */
int sfe_engine_enqueue(sfe_engine *e, int64_t timesteps);
/* `timesteps` steps of n independent chips in lock step with one launch per phase and step for the whole batch
 * (design-space sweeps; sfe_batch_sim uses it). 0 = enqueued, 1 = these chips cannot run as one batch (nothing enqueued),
 * -1 = error. Collect every engine afterwards as after sfe_engine_enqueue. */
int sfe_engine_batch_enqueue(sfe_engine *const *engines, uint32_t n, int64_t timesteps);
/* Collect per-step records of the steps enqueued since the last collect. */
int sfe_engine_collect(sfe_engine *e, sfe_run_data *out);
/* reset(): zero model state, keep the timestep counter (src/chip.cpp:576-600) */
int sfe_engine_reset(sfe_engine *e);
/* Per-neuron bias patch from HOST memory (MappedNeuron.set_attributes fast path,
 * src/pymodule.cpp:1176-1181); bias has n_neurons entries in device-index order.
 * sfe_engine_set_bias is asynchronous and double-buffered: the vector is copied on a separate
 * stream into the inactive device buffer (overlapping the steps already enqueued) and takes
 * effect with the next step that is enqueued; keep the host buffer unchanged until then. */
int sfe_engine_set_bias(sfe_engine *e, const double *bias, size_t n);
/* the same from pageable memory: staged through an engine-owned pinned buffer; the caller's vector is free on return */
int sfe_engine_set_bias_staged(sfe_engine *e, const double *bias, size_t n);
/* Poisson inputs are drawn on the host (the reference's generator is libstdc++'s std::mt19937 +
 * uniform_real_distribution, one draw per update of a neuron of the unit, src/models.cpp:863-903) and reach
 * the device as an overlay: bits[step][col] != 0 makes the input neuron with that poisson_col spike in that
 * step. Covers the next `n_steps` steps to be enqueued; sfe_chip_sim does this itself. */
int sfe_engine_set_input_overlay(sfe_engine *e, const uint8_t *bits, int64_t n_steps, uint32_t n_cols);
/* rand() values of the TrueNorth neurons with random_mask != 0 for the next n_steps timesteps: values[n_steps][n_cols],
 * n_cols = sfe_tables.n_rand_cols. An engine with such neurons refuses to step past its overlay. */
int sfe_engine_set_rand_overlay(sfe_engine *e, const uint32_t *values, int64_t n_steps, uint32_t n_cols);
/* Cooperative cancellation, callable from another thread: a running sfe_engine_run / sfe_chip_sim returns -1
 * ("simulation interrupted") at its next batch boundary (at most 4096 steps); the steps done so far stay done.
 * The reference polls PyErr_CheckSignals every 100 ms inside its loop (src/pymodule.cpp:629-652). on = 0 re-arms. */
void sfe_engine_request_stop(sfe_engine *e, int on);
/* The same overlay drawn on the device by one MT19937 per input unit (csrc/mt19937.cuh, checked on the host against
 * libstdc++ and on a B200 against the reference's golden); sfe_chip_sim uses it unless SFE_DEVICE_POISSON=0. Do not
 * mix host-drawn and device-drawn overlays on one engine: they are two separate streams of draws. */
int sfe_engine_fill_input_overlay(sfe_engine *e, int64_t n_steps);
/* Host-only source of that overlay (no device needed): one std::mt19937 per Poisson unit, seeded
 * tables->input_seed_base + unit + 1. sfe_poisson_fill writes bits[n_steps][sfe_poisson_cols] for the next
 * n_steps steps and advances the streams (they are not rewound by a reset: InputModel::reset, src/models.hpp:358). */
typedef struct sfe_poisson sfe_poisson;
sfe_poisson *sfe_poisson_create(const sfe_tables *tables);
void sfe_poisson_destroy(sfe_poisson *p);
uint32_t sfe_poisson_cols(const sfe_poisson *p);
int sfe_poisson_fill(sfe_poisson *p, uint8_t *bits, int64_t n_steps);
/* glibc's rand() (TYPE_3 additive feedback generator) restated: out[k] = the (skip + k)-th value after srand(seed).
 * The reference's TrueNorth threshold jitter draws from the process-global std::rand() (src/models.cpp:757). */
void sfe_glibc_rand_draws(uint32_t seed, uint64_t skip, uint32_t *out, size_t n);
/* cross-check hooks: n draws of U(0,1) from seed through libstdc++'s std::mt19937 + uniform_real_distribution, and
 * through the engine's own host/device MT19937 (csrc/mt19937.cuh, state interleaved with `stride`) */
void sfe_poisson_reference_draws(uint32_t seed, double *out, size_t n);
void sfe_mt19937_draws(uint32_t seed, double *out, size_t n, size_t stride);
int sfe_engine_set_neuron_bias(sfe_engine *e, uint32_t neuron, double bias);
/* LIF "potential" attribute after load (src/models.cpp:433-436): sets the membrane potential of one neuron */
int sfe_engine_set_neuron_potential(sfe_engine *e, uint32_t neuron, double potential);
int sfe_engine_read_potentials(sfe_engine *e, double *out, size_t n);
int sfe_engine_read_fired(sfe_engine *e, uint32_t *bits, size_t n_words);
/* the last step's raster in the device's padded layout (world * slice words; every core starts on
 * a word boundary, see sfe_engine_raster_layout): one device->host copy, nothing repacked */
int sfe_engine_read_raster(sfe_engine *e, uint32_t *words, size_t n_words);
int64_t sfe_engine_total_timesteps(const sfe_engine *e);
/* kernel launches issued so far (for bench.py's gpu_launches) */
int64_t sfe_engine_launch_count(const sfe_engine *e);
/* device time (ms) of the steps enqueued between the last two timing marks */
/* whether time_begin/time_end also bracket every message-phase launch with events (default 1);
 * 0 for throughput runs: an event between two kernels defeats programmatic dependent launch */
int sfe_engine_time_launches(sfe_engine *e, int on);
int sfe_engine_time_begin(sfe_engine *e);
int sfe_engine_time_end(sfe_engine *e, float *ms_total, float *ms_fanout);
/* ---- multi-GPU: the chip's cores are partitioned into `world` contiguous ranges (balanced by
 * synapse count); rank r simulates range r. One timestep is
 *     sfe_engine_enqueue_neuron_phase   local neuron phase, writes this rank's raster slice
 *     all-gather of the slices          (NCCL by the caller: local ptr -> global ptr, equal sizes)
 *     sfe_engine_enqueue_message_phase  raises the local inbox bits of every fired neuron of the
 *                                       chip, then message phase + energy/timing of the local cores
 * Per-step records are partial (local cores): sum counts/energies over ranks, max of sim_time. */
sfe_engine *sfe_engine_create_partitioned(const sfe_tables *tables, int device, uint32_t rank, uint32_t world);
int sfe_engine_enqueue_neuron_phase(sfe_engine *e);
int sfe_engine_enqueue_message_phase(sfe_engine *e);
void *sfe_engine_fired_local_ptr(sfe_engine *e, size_t *n_bytes);
void *sfe_engine_fired_global_ptr(sfe_engine *e, size_t *n_bytes);
/* caller-owned exchange buffers (device memory): local = slice, global = world * slice */
int sfe_engine_set_exchange_buffers(sfe_engine *e, void *local, void *global);
int64_t sfe_engine_collect_records(sfe_engine *e, sfe_step_record *out, int64_t cap);
/* NCCL exchange driven from C++ (libnccl.so.2 is dlopen'ed; pass its path or NULL): rank 0
 * obtains a 128-byte unique id and ships it to the other ranks, every rank then calls
 * sfe_engine_comm_init; sfe_engine_enqueue_partitioned enqueues whole steps (neuron phase,
 * ncclAllGather of the raster slices over NVLink, message phase) without host round trips. */
int sfe_nccl_get_unique_id(void *out128, const char *libnccl_path);
int sfe_engine_comm_init(sfe_engine *e, const void *unique_id128, const char *libnccl_path);
int sfe_engine_comm_destroy(sfe_engine *e);
int sfe_engine_enqueue_partitioned(sfe_engine *e, int64_t timesteps);
/* Peer-memory exchange (CUDA IPC over NVLink / NVSwitch), one process per GPU, up to 8 ranks:
 * the neuron-phase kernel of every rank stores its raster words directly into every rank's
 * double-buffered raster and raises an arrival flag there; the message phase begins by waiting
 * (bounded) for all flags — no collective launch between the phases. Every rank exports a
 * 64-byte handle, the caller gathers the handles of all ranks in rank order (its own control
 * plane) and hands them to attach. After attach, sfe_engine_enqueue_partitioned needs no NCCL
 * communicator. All ranks must enqueue the same number of steps and synchronise on the host
 * before detach/destroy. sfe_engine_exchange_error: 1 if a peer failed to arrive in time. */
int sfe_engine_p2p_export(sfe_engine *e, void *handle64);
int sfe_engine_p2p_attach(sfe_engine *e, const void *handles_world_x_64);
int sfe_engine_p2p_detach(sfe_engine *e);
int sfe_engine_exchange_error(sfe_engine *e);
/* diagnostic (SFE_TIMELINE=1 at engine creation): %globaltimer stamps of every CTA of the fused step kernel for the
 * last 64 steps, out[64][grid][16]; returns the grid size, 0 when nothing was recorded */
int sfe_engine_read_timeline(sfe_engine *e, unsigned long long *out, size_t cap_words);
int64_t sfe_engine_read_timeline_raw(sfe_engine *e, unsigned long long *out, size_t cap_words); /* diagnostic builds */
/* last n records of the device log (for steps replayed from a captured CUDA graph) */
int64_t sfe_engine_read_log_tail(sfe_engine *e, sfe_step_record *out, int64_t n);
/* Host-only plan of a chip split over `world` GPUs (no CUDA call): owner[c] = rank of core c
 * (contiguous ranges balanced by synapse + neuron work; the reference keeps all cores in one
 * process, src/chip.cpp:586-618), fired_word_begin[c] = first word of core c in the exchanged
 * fired-bit raster (world equal slices of *slice_words words). Outputs may be NULL. */
int sfe_plan_partition(const sfe_tables *tables, uint32_t world, uint32_t *owner, uint32_t *fired_word_begin,
        uint32_t *slice_words);
int sfe_engine_partition_info(const sfe_engine *e, uint32_t *rank, uint32_t *world, uint32_t *slice_words,
        uint32_t *local_cores, uint64_t *local_neurons);
int sfe_engine_raster_layout(const sfe_engine *e, uint32_t *word_begin, size_t n_cores);
int sfe_engine_synchronize(sfe_engine *e);
int sfe_device_memcpy(void *dst, const void *src, size_t bytes);
/* pinned host memory for buffers that cross the ABI every step */
void *sfe_host_alloc(size_t bytes);
void sfe_host_free(void *p);
size_t sfe_engine_device_bytes(const sfe_engine *e);

/* ---- description level --------------------------------------------------- */
typedef struct sfe_arch sfe_arch;
typedef struct sfe_net sfe_net;
typedef struct sfe_chip sfe_chip;

/* load_arch / load_net  (src/arch.cpp:106-117, src/network.cpp:194-222) */
sfe_arch *sfe_arch_load_yaml(const char *path);
sfe_net *sfe_net_load_yaml(const char *path, sfe_arch *arch);
/* the legacy netlist format (`sim -n`; load_net(..., use_netlist_format=true), src/netlist.cpp:38-617) */
sfe_net *sfe_net_load_netlist(const char *path, sfe_arch *arch);
/* flat JSON-lines description (oracle/yaml_to_flat.py); returns both objects */
int sfe_load_flat(const char *path, sfe_arch **arch, sfe_net **net);
void sfe_arch_free(sfe_arch *a);
void sfe_net_free(sfe_net *n);

/* SpikingChip(arch)  src/chip.cpp:61-104. device < 0: host-only chip (lowering
 * and table export work; sim() fails with "no CUDA device"). */
sfe_chip *sfe_chip_create(const sfe_arch *arch, int device);
/* before load(): this chip object simulates partition `rank` of `world` (multi-GPU) */
int sfe_chip_set_partition(sfe_chip *c, uint32_t rank, uint32_t world);
/* Poisson inputs are seeded by how many "input" units the PROCESS created before this chip's (the reference's
 * InputModel::instance_counter is a process-wide static, src/models.hpp:366); sfe_chip_create keeps that count.
 * This overrides it for the chip (0 = as if the chip were the first one of a fresh process). Call before load. */
int sfe_chip_set_input_seed_base(sfe_chip *c, uint32_t base);

/* ---- design-space exploration: independent chips side by side on one GPU (BASELINE config 5) ----------
 * The reference runs a sweep as separate processes (scripts/, one `sim` per design point). Here the n chips of a
 * sweep live in one process; every chip owns a stream, and `host_threads` worker threads (0 = one per host core, at most n) drive
 * sfe_chip_load / sfe_chip_sim of different chips concurrently, so the kernels of that many chips overlap on the
 * device. Each chip goes through exactly the code path of a lone chip: its results are those of running it alone.
 * reqs: NULL or n trace requests; out: NULL or n records. Returns 0, or -1 with sfe_last_error() naming the first
 * chip that failed (the others still ran). */
int sfe_batch_load(sfe_chip *const *chips, const sfe_net *const *nets, uint32_t n, uint32_t host_threads);
int sfe_batch_sim(sfe_chip *const *chips, uint32_t n, int64_t timesteps, int timing_model,
        const sfe_trace_request *reqs, sfe_run_data *out, uint32_t host_threads);
void sfe_chip_destroy(sfe_chip *c);
/* SpikingChip::load  src/chip.cpp:129-138 */
int sfe_chip_load(sfe_chip *c, const sfe_net *net);
/* bulk extension: map the synthetic network of `spec` (sfe_synth.h) without
 * materialising per-edge objects */
int sfe_chip_load_synthetic(sfe_chip *c, const sfe_synth_spec *spec, int generate_on_device);
/* SpikingChip::sim  src/chip.cpp:477-533 */
int sfe_chip_sim(sfe_chip *c, int64_t timesteps, int timing_model, const sfe_trace_request *req,
        sfe_run_data *out);
/* host-side detailed timing model (src/schedule.cpp:208-620) over a status trace
 * ([timesteps][n_neurons] SFE_STATUS_* bytes): per-step sim_time */
int sfe_chip_schedule_detailed(sfe_chip *c, const uint8_t *status, int64_t timesteps, double *sim_time);
/* Host threads of the detailed timing model (the reference's -S / scheduler_threads): timesteps are scheduled
 * independently, a batch of them is spread over this many threads. 0 (default) = one per host core. */
int sfe_chip_set_scheduler_threads(sfe_chip *c, uint32_t threads);
int sfe_chip_reset(sfe_chip *c);       /* src/chip.cpp:576-600 */
double sfe_chip_get_power(sfe_chip *c);/* src/chip.cpp:607-621 */
const sfe_tables *sfe_chip_tables(const sfe_chip *c);
sfe_engine *sfe_chip_engine(sfe_chip *c);
/* group/offset <-> device index, trace order (lexicographic group, offset; src/chip.cpp:1616-1629) */
int64_t sfe_chip_neuron_index(const sfe_chip *c, const char *group, uint64_t offset);
/* MappedNeuron.set_attributes for a numeric soma attribute (bias, threshold, ...). Bias patches are collected in the
 * host table and reach the device as one vector before the next sfe_chip_sim (or sfe_chip_engine) call, so a
 * per-frame loop over a thousand input neurons costs one upload, not a thousand. */
int sfe_chip_set_neuron_attribute(sfe_chip *c, const char *group, uint64_t offset, const char *name,
        double value);
/* Spike rows of a fired-bit raster in the reference's trace order and format,
 * "group.offset,timestep\n" (src/chip.cpp:1610-1630). Returns the bytes needed. */
size_t sfe_chip_format_spikes(const sfe_chip *c, const uint32_t *fired_bits, int64_t timesteps,
        int64_t timestep_start, char *buf, size_t cap);
/* Rows of messages.csv (sim_trace_record_message, src/chip.cpp:1731-1764; order of
 * sim_sort_and_record_messages, :439-456) for `timesteps` steps, rebuilt from the per-neuron status
 * bytes of those steps (sfe_trace_request.status). SFE_TIMING_DETAILED runs the NoC scheduler so
 * that the timestamps and network / blocking delays are filled, SFE_TIMING_SIMPLE leaves them at
 * -inf / 0 like the reference. Message ids continue over the chip's lifetime in single-thread
 * creation order. Returns the text length; a call with buf = NULL only sizes the buffer (it does
 * not consume message ids). */
size_t sfe_chip_format_messages(sfe_chip *chip, const uint8_t *status, int64_t timesteps, int64_t timestep_start,
        int timing_model, char *buf, size_t cap);
/* "group.offset\n" of every potential probe, in potentials.csv column order (src/chip.cpp:1454-1476) */
size_t sfe_chip_probe_names(const sfe_chip *c, char *buf, size_t cap);
/* header cells of neurons.csv (src/chip.cpp:1478-1517): "group.offset/trace" of every model-defined trace column
 * (sfe_trace_request.neuron_traces), one per line */
size_t sfe_chip_trace_names(const sfe_chip *c, char *buf, size_t cap);
/* SpikingChip::mapped_neuron_groups (src/chip.hpp:104): "group name<TAB>neuron count" per line, lexicographic */
size_t sfe_chip_group_names(const sfe_chip *c, char *buf, size_t cap);
/* MappedNeuron::set_attributes(..., log_spikes) (src/mapped.cpp:113-124) */
int sfe_chip_set_neuron_log_spikes(sfe_chip *c, const char *group, uint64_t offset, int on);
/* sfe_engine_request_stop for a chip; sfe_chip_sim re-arms it when it starts */
void sfe_chip_request_stop(sfe_chip *c);
/* an empty network for the object-by-object builders, and Network.save (src/network.cpp:693-712; YAML) */
sfe_net *sfe_net_create(const char *name);
int sfe_net_save_yaml(const sfe_net *net, const char *path);
/* SpikingNetwork::save(path, use_netlist_format = true): the legacy netlist format, with the reference's losses (groups
 * become "0", "1", ...; doubles print with 6 digits)  src/network.cpp:606-703, src/netlist.cpp:619-851 */
int sfe_net_save_netlist(const sfe_net *net, const char *path);

#ifdef __cplusplus
}
#endif
#endif /* SANAFE_B200_H_ */
