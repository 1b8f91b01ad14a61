/* sfe_synth.h — counter-based generator for the synthetic random LIF network
 * (BASELINE.json configs[3]: "synthetic random LIF net on arch/loihi_large.yaml:
 * 1M neurons x 1000 fan-out").
 *
 * Shape follows the reference's own random-network experiment
 * (scripts/tcad2025/random_network.py:47-124: one population "pop", neurons
 * mapped core by core, every sending neuron targets `dest_cores` cores with
 * `syn_per_axon` distinct post-neurons per core, a fixed fraction of neurons is
 * bias-driven). Python's Mersenne-Twister `random.sample` is replaced by a
 * stateless 64-bit mixer so that every element (bias flag, post-neuron, weight,
 * delay) is a pure function of (seed, index): the reference harness (C++), the
 * CPU restatement (C), numpy tests and the CUDA bulk loader all evaluate the very
 * same function and never have to exchange a 12 GB array.
 *
 * Plain C99, usable from C, C++ and CUDA (__host__ __device__).
 */
#ifndef SFE_SYNTH_H_
#define SFE_SYNTH_H_

#include <stdint.h>

#ifdef __CUDACC__
#define SFE_HD __host__ __device__ __forceinline__
#else
#define SFE_HD static inline
#endif

typedef struct sfe_synth_spec
{
    uint32_t cores;            /* C: mapped cores 0..C-1 (core id order = tile-major) */
    uint32_t neurons_per_core; /* P: power of two, <= 4096 */
    uint32_t dest_cores;       /* D: each neuron sends to cores (home+k) mod C, k<D; D<=C */
    uint32_t syn_per_axon;     /* S: distinct post-neurons per destination core; S<=P */
    uint64_t seed;
    uint32_t bias_permille;    /* fraction of bias-driven neurons, in 1/1000 */
    double bias;               /* bias of driven neurons (others 0) */
    double threshold;
    double reset;
    double leak_decay;
    int32_t w_min;             /* integer weights uniform in [w_min, w_max] */
    int32_t w_max;
    uint32_t max_delay;        /* synaptic delays uniform in [0, max_delay] (<=5) */
    uint32_t log_spikes;       /* 1: every neuron logs spikes */
    uint32_t log_potential_n;  /* first n neurons of the population log potentials */
} sfe_synth_spec;

SFE_HD uint64_t sfe_mix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

SFE_HD uint64_t sfe_hash(uint64_t seed, uint64_t stream, uint64_t idx)
{
    return sfe_mix64(sfe_mix64(seed + stream * 0xD6E8FEB86659FD93ull) ^ idx);
}

/* 1 if neuron n (global index) is bias-driven. */
SFE_HD int sfe_synth_has_bias(const sfe_synth_spec *s, uint64_t n)
{
    return (sfe_hash(s->seed, 1, n) % 1000ull) < s->bias_permille;
}

/* Destination core of the k-th axon of a neuron whose home core is c. */
SFE_HD uint32_t sfe_synth_dest_core(const sfe_synth_spec *s, uint32_t c, uint32_t k)
{
    return (c + k) % s->cores;
}

/* Per-axon constants for the post-neuron bijection (axon = n*D + k). */
typedef struct sfe_synth_axon
{
    uint32_t start, stride, mask;
} sfe_synth_axon;

SFE_HD sfe_synth_axon sfe_synth_axon_params(const sfe_synth_spec *s, uint64_t axon)
{
    const uint64_t h = sfe_hash(s->seed, 2, axon);
    const uint32_t P = s->neurons_per_core;
    sfe_synth_axon a;
    a.start = (uint32_t) (h % P);
    a.stride = ((uint32_t) ((h >> 20) % P)) | 1u;
    a.mask = (uint32_t) ((h >> 40) % P);
    return a;
}

/* j-th post-neuron (offset within the destination core) of an axon. P is a power
 * of two, stride is odd => j -> (start + j*stride) mod P is a bijection, and so is
 * the xor with mask < P: the S post-neurons of an axon are distinct. */
SFE_HD uint32_t sfe_synth_post(const sfe_synth_spec *s, const sfe_synth_axon *a, uint32_t j)
{
    const uint32_t P = s->neurons_per_core;
    return ((a->start + j * a->stride) & (P - 1u)) ^ a->mask;
}

/* Weight and delay of synapse `syn` = (n*D + k)*S + j. */
SFE_HD int32_t sfe_synth_weight(const sfe_synth_spec *s, uint64_t syn)
{
    const uint64_t span = (uint64_t) ((int64_t) s->w_max - (int64_t) s->w_min + 1);
    return s->w_min + (int32_t) (sfe_hash(s->seed, 3, syn) % span);
}

SFE_HD uint32_t sfe_synth_delay(const sfe_synth_spec *s, uint64_t syn)
{
    if (s->max_delay == 0)
    {
        return 0;
    }
    return (uint32_t) (sfe_hash(s->seed, 4, syn) % (uint64_t) (s->max_delay + 1u));
}

#endif /* SFE_SYNTH_H_ */
