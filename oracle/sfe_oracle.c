/* sfe_oracle.c — CPU restatement of the reference's per-timestep algorithm.
 *
 * TEST INFRASTRUCTURE ONLY. Imported/linked only by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg, and there only as the checker. The product
 * (libsanafe_b200.so) never links or calls it.
 *
 * Parity status: PINNED. This restatement is checked against outputs of the
 * reference itself (oracle/_ref/sanafe_ref, the unmodified reference engine
 * compiled from /root/reference/src) on the fixtures under tests/golden/ —
 * spike rasters, per-step counters, potentials and model-defined traces bit-exact,
 * energies/latencies to 1e-12 relative (tests/test_oracle_vs_reference.py, 12 cases) —
 * and against the expectations of the reference's own per-model unit tests
 * (tests/test_reference_unit_vectors.py).
 *
 * It walks the same lowered tables (include/sanafe_b200.h) as the CUDA engine,
 * strictly sequentially and in the reference's own order, so every floating-point
 * sum that decides a spike is formed exactly as the reference forms it:
 *   neuron phase    src/chip.cpp:624-654, 710-736, 802-834
 *   message phase   src/chip.cpp:656-764, 1127-1169
 *   soma models     src/models.cpp:441-567 (LIF, incl. the file noise stream :589-650), 724-830 (TrueNorth),
 *                   863-903 (input: spike trains, rate, Poisson with an own MT19937), plugins/hodgkin_huxley.cpp:116-170,
 *                   plugins/neurofem.cpp:192-317 (combined dendrite + soma unit, sigma_v = 0)
 *   dendrites       src/models.cpp:71-94 (accumulator), 96-131 (accumulator_with_delay), 167-259 (taps)
 *   default costs   src/pipeline.hpp:511-731
 *   energy/counters src/chip.cpp:1028-1051, 1171-1261
 *   simple timing   src/schedule.cpp:61-102
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "sanafe_b200.h"

typedef struct sfe_oracle
{
    const sfe_tables *t;
    int64_t total_timesteps;
    /* soma state */
    double *v, *u, *bias;
    int32_t *refractory;
    uint8_t *status;
    /* dendrite state: buffer (buffer before soma) or ring (buffer inside dendrite) */
    double *buf;      /* [n_neurons]                       timestep_buffer current */
    uint8_t *buf_has; /* [n_neurons]                       timestep_buffer has_value */
    double *acc;      /* accumulator charge (per neuron) */
    int64_t *acc_step;
    double *ring;     /* [n_neurons * 6] absolute-step ring of the delay line */
    uint8_t *ring_has;
    /* HH state */
    double *hh_v, *hh_m, *hh_n, *hh_h, *hh_i;
    /* NeuroFEM (plugins/neurofem.cpp:62-90): potential = v, u1 = u; u2, u_integrated and the two accumulators that the
     * message phase of a step fills for the neuron phase of the next (next_u1/u2_dendritic_accumulator.value_or(0)) */
    double *nf_u2, *nf_uint, *nf_acc;
    /* inbox */
    uint8_t *axon_active;
    /* per-step scratch */
    double *core_gen, *core_proc;
    int64_t *tile_e, *tile_w, *tile_n, *tile_s;
    int64_t *core_msgs;
    double *core_syn_e, *core_den_e, *core_soma_e, *core_axout_e;
    double *core_dup_e; /* energy of combined dendrite + soma units: in both buckets, once in the total */
    double total_sim_time, total_energy;
    /* Poisson inputs: one MT19937 per input unit (own restatement, below), or an external overlay */
    struct mt19937 *poisson_gen; /* [n_poisson_units], indexed by sfe_input_desc.unit */
    uint32_t n_poisson_units;
    /* "taps" dendrites (MultiTapModel1D): voltages of every line, a scratch line, steps simulated per line */
    double *tap_v, *tap_next;
    uint32_t *tap_off;           /* [n_taps_units] offset of the unit's line in tap_v */
    int64_t *tap_steps;          /* [n_taps_units] timesteps_simulated */
    /* TrueNorth threshold jitter: the process-global rand() of the reference, restated (below) and drawn as the
     * neurons are updated - one processing thread, cores and neurons in order */
    uint32_t rand_state[31];
    int rand_front, rand_rear;
    const uint8_t *overlay;      /* caller-owned [overlay_steps][overlay_cols], NULL: draw here */
    int64_t overlay_step0, overlay_steps;
    uint32_t overlay_cols;
} sfe_oracle;

static void oracle_srand(sfe_oracle *o, uint32_t seed);

/* MT19937 (Matsumoto & Nishimura 1998) as std::mt19937 specifies it, and the 53-bit canonical
 * double libstdc++'s std::uniform_real_distribution<double>{0,1} builds from two 32-bit outputs
 * (lo first): U = (lo + hi * 2^32) / 2^64, nudged below 1.0. InputModel: src/models.hpp:347-366. */
struct mt19937
{
    uint32_t mt[624];
    int idx;
    int seeded;
};
static void mt_seed(struct mt19937 *g, uint32_t seed)
{
    g->mt[0] = seed;
    for (int i = 1; i < 624; ++i) g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t) i;
    g->idx = 624;
    g->seeded = 1;
}
static uint32_t mt_next(struct mt19937 *g)
{
    if (g->idx >= 624)
    {
        for (int i = 0; i < 624; ++i)
        {
            const uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
            g->mt[i] = g->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        g->idx = 0;
    }
    uint32_t y = g->mt[g->idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}
static double mt_canonical(struct mt19937 *g)
{
    const double lo = (double) mt_next(g);
    const double hi = (double) mt_next(g);
    double u = (lo + hi * 4294967296.0) / 18446744073709551616.0;
    if (u >= 1.0) u = nextafter(1.0, 0.0);
    return u;
}

#define RING 6 /* max_delay 5 + 1  (src/models.hpp:158) */

static void *zalloc(size_t n, size_t sz)
{
    return calloc(n ? n : 1, sz);
}

sfe_oracle *sfe_oracle_create(const sfe_tables *t)
{
    if (t == NULL || t->abi_version != SFE_ABI_VERSION) return NULL;
    if (t->n_synapses != 0 && (t->syn_weight == NULL || t->syn_meta == NULL)) return NULL; /* needs host synapse arrays */
    sfe_oracle *o = (sfe_oracle *) zalloc(1, sizeof(*o));
    const size_t n = t->n_neurons;
    o->t = t;
    o->v = (double *) zalloc(n, sizeof(double));
    o->u = (double *) zalloc(n, sizeof(double));
    o->bias = (double *) zalloc(n, sizeof(double));
    o->refractory = (int32_t *) zalloc(n, sizeof(int32_t));
    o->status = (uint8_t *) zalloc(n, 1);
    o->buf = (double *) zalloc(n, sizeof(double));
    o->buf_has = (uint8_t *) zalloc(n, 1);
    o->acc = (double *) zalloc(n, sizeof(double));
    o->acc_step = (int64_t *) zalloc(n, sizeof(int64_t));
    o->ring = (double *) zalloc(n * RING, sizeof(double));
    o->ring_has = (uint8_t *) zalloc(n * RING, 1);
    o->hh_v = (double *) zalloc(t->n_hh, sizeof(double));
    o->hh_m = (double *) zalloc(t->n_hh, sizeof(double));
    o->hh_n = (double *) zalloc(t->n_hh, sizeof(double));
    o->hh_h = (double *) zalloc(t->n_hh, sizeof(double));
    o->hh_i = (double *) zalloc(t->n_hh, sizeof(double));
    o->nf_u2 = (double *) zalloc(n, sizeof(double));
    o->nf_uint = (double *) zalloc(n, sizeof(double));
    o->nf_acc = (double *) zalloc(2 * n, sizeof(double));
    o->core_dup_e = (double *) zalloc(t->n_cores, sizeof(double));
    o->axon_active = (uint8_t *) zalloc(t->n_axons_in, 1);
    o->core_gen = (double *) zalloc(t->n_cores, sizeof(double));
    o->core_proc = (double *) zalloc(t->n_cores, sizeof(double));
    o->tile_e = (int64_t *) zalloc(t->n_tiles, sizeof(int64_t));
    o->tile_w = (int64_t *) zalloc(t->n_tiles, sizeof(int64_t));
    o->tile_n = (int64_t *) zalloc(t->n_tiles, sizeof(int64_t));
    o->tile_s = (int64_t *) zalloc(t->n_tiles, sizeof(int64_t));
    o->core_msgs = (int64_t *) zalloc(t->n_cores, sizeof(int64_t));
    o->core_syn_e = (double *) zalloc(t->n_cores, sizeof(double));
    o->core_den_e = (double *) zalloc(t->n_cores, sizeof(double));
    o->core_soma_e = (double *) zalloc(t->n_cores, sizeof(double));
    o->core_axout_e = (double *) zalloc(t->n_cores, sizeof(double));
    for (size_t i = 0; i < n; ++i)
    {
        o->bias[i] = t->neuron_bias[i];
        o->v[i] = t->neuron_potential0[i];
    }
    {
        size_t total = 0, widest = 1;
        o->tap_off = (uint32_t *) zalloc(t->n_taps_units, sizeof(uint32_t));
        o->tap_steps = (int64_t *) zalloc(t->n_taps_units, sizeof(int64_t));
        for (uint32_t k = 0; k < t->n_taps_units; ++k)
        {
            o->tap_off[k] = (uint32_t) total;
            total += t->taps[k].n_taps;
            if (t->taps[k].n_taps > widest) widest = t->taps[k].n_taps;
        }
        o->tap_v = (double *) zalloc(total, sizeof(double));
        o->tap_next = (double *) zalloc(widest, sizeof(double));
    }
    for (uint32_t k = 0; k < t->n_inputs; ++k)
        if (t->inputs[k].poisson > 0.0 && t->inputs[k].unit + 1 > o->n_poisson_units) o->n_poisson_units = t->inputs[k].unit + 1;
    o->poisson_gen = (struct mt19937 *) zalloc(o->n_poisson_units, sizeof(struct mt19937));
    oracle_srand(o, 1u); /* a fresh reference process: srand() is never called */
    for (size_t i = 0; i < t->n_hh; ++i)
    {
        o->hh_m[i] = t->hh[i].m;
        o->hh_n[i] = t->hh[i].n;
        o->hh_h[i] = t->hh[i].h;
        o->hh_i[i] = t->hh[i].current;
    }
    return o;
}

void sfe_oracle_destroy(sfe_oracle *o)
{
    if (o == NULL) return;
    free(o->v); free(o->u); free(o->bias); free(o->refractory); free(o->status);
    free(o->buf); free(o->buf_has); free(o->acc); free(o->acc_step); free(o->ring); free(o->ring_has);
    free(o->hh_v); free(o->hh_m); free(o->hh_n); free(o->hh_h); free(o->hh_i);
    free(o->nf_u2); free(o->nf_uint); free(o->nf_acc); free(o->core_dup_e);
    free(o->axon_active); free(o->core_gen); free(o->core_proc);
    free(o->tile_e); free(o->tile_w); free(o->tile_n); free(o->tile_s); free(o->core_msgs);
    free(o->core_syn_e); free(o->core_den_e); free(o->core_soma_e); free(o->core_axout_e);
    free(o->poisson_gen);
    free(o->tap_v); free(o->tap_next); free(o->tap_off); free(o->tap_steps);
    free(o);
}

/* Use an externally drawn Poisson overlay (the product's sfe_poisson_fill) for the next n_steps steps
 * instead of the oracle's own generator: lets the tests check the product's draws on the CPU. */
void sfe_oracle_set_input_overlay(sfe_oracle *o, const uint8_t *bits, int64_t n_steps, uint32_t n_cols)
{
    o->overlay = bits;
    o->overlay_step0 = o->total_timesteps;
    o->overlay_steps = n_steps;
    o->overlay_cols = n_cols;
}

void sfe_oracle_set_bias(sfe_oracle *o, const double *bias, size_t n)
{
    memcpy(o->bias, bias, n * sizeof(double));
}

/* SpikingChip::reset  src/chip.cpp:576-600: buffers, model state and statuses are
 * cleared; timestep counters are not. LIF::reset zeroes u and v only. */
void sfe_oracle_reset(sfe_oracle *o)
{
    const size_t n = o->t->n_neurons;
    memset(o->v, 0, n * sizeof(double));
    memset(o->u, 0, n * sizeof(double));
    memset(o->status, 0, n);
    memset(o->buf_has, 0, n);
    memset(o->acc, 0, n * sizeof(double));
    memset(o->ring_has, 0, n * RING);
    memset(o->ring, 0, n * RING * sizeof(double));
    memset(o->nf_u2, 0, n * sizeof(double)); /* NeuroFEMModel::reset  plugins/neurofem.cpp:319-335 */
    memset(o->nf_uint, 0, n * sizeof(double));
    memset(o->nf_acc, 0, 2 * n * sizeof(double));
    for (size_t i = 0; i < o->t->n_hh; ++i)
    {
        o->hh_v[i] = o->hh_m[i] = o->hh_n[i] = o->hh_h[i] = 0.0; /* plugins/hodgkin_huxley.cpp:71-87 */
    }
    if (o->t->n_taps_units != 0) /* MultiTapModel1D::reset  src/models.cpp:340-348: voltages only */
    {
        const sfe_taps_desc *last = &o->t->taps[o->t->n_taps_units - 1];
        memset(o->tap_v, 0, ((size_t) o->tap_off[o->t->n_taps_units - 1] + last->n_taps) * sizeof(double));
    }
}

/* LoihiLifModel::update  src/models.cpp:497-567 */
static int lif_update(sfe_oracle *o, const sfe_soma_class *c, size_t i, int has_in, double in, int64_t steps_done)
{
    double v = o->v[i], u = o->u[i];
    const double bias = o->bias[i];
    int state = SFE_STATUS_IDLE;
    if ((fabs(v) > 0.0) || has_in || (fabs(bias) > 0.0) || (c->flags & SFE_SOMA_FORCE_UPDATE)) state = SFE_STATUS_UPDATED;
    if (steps_done > 0)
    {
        u *= c->input_decay;
        v *= c->leak;
    }
    v = (double) ((int) (v * 64.0)) / 64.0; /* loihi_quantize: C cast, truncation toward zero */
    if (c->flags & SFE_SOMA_NOISE)
    {
        /* loihi_generate_noise  src/models.cpp:535-539, 589-650: the unit consumes one file entry per
         * update of any of its compartments (in-core order) and rewinds at the end of the file */
        const sfe_noise_desc *nd = &o->t->noise[o->t->neuron_aux[i]];
        const uint64_t cursor = (uint64_t) steps_done * nd->share_count + nd->share_rank;
        v += o->t->noise_values[nd->off + cursor % nd->len];
    }
    if (o->refractory[i] <= 0)
    {
        v += bias;
        u += has_in ? in : 0.0;
        v += u;
        if (v > c->threshold)
        {
            if (c->reset_mode == SFE_RESET_HARD) v = c->reset;
            else if (c->reset_mode == SFE_RESET_SOFT) v -= c->threshold;
            o->refractory[i] = c->refractory_delay;
            state = SFE_STATUS_FIRED;
        }
        if (v < c->reverse_threshold)
        {
            if (c->reverse_reset_mode == SFE_RESET_SOFT) v -= c->reverse_threshold;
            else if (c->reverse_reset_mode == SFE_RESET_HARD) v = c->reverse_reset;
            else if (c->reverse_reset_mode == SFE_RESET_SATURATE) v = c->reverse_threshold;
        }
    }
    o->refractory[i] = o->refractory[i] - 1 > 0 ? o->refractory[i] - 1 : 0;
    o->v[i] = v;
    o->u[i] = u;
    return state;
}

/* rand() of glibc (stdlib/random_r.c, TYPE_3): 31 words seeded by a Park-Miller LCG from srand's argument (1 when the
 * program never calls srand), word[f] += word[r] with f three ahead of r, result = word >> 1, 310 values discarded */
static uint32_t oracle_rand(sfe_oracle *o)
{
    uint32_t *s = o->rand_state;
    s[o->rand_front] += s[o->rand_rear];
    const uint32_t result = s[o->rand_front] >> 1;
    if (++o->rand_front == 31) o->rand_front = 0;
    if (++o->rand_rear == 31) o->rand_rear = 0;
    return result;
}
static void oracle_srand(sfe_oracle *o, uint32_t seed)
{
    int64_t word = seed == 0 ? 1 : (int32_t) seed;
    o->rand_state[0] = (uint32_t) word;
    for (int i = 1; i < 31; ++i)
    {
        word = (16807 * word) % 2147483647;
        if (word < 0) word += 2147483647;
        o->rand_state[i] = (uint32_t) word;
    }
    o->rand_front = 3;
    o->rand_rear = 0;
    for (int i = 0; i < 310; ++i) (void) oracle_rand(o);
}

/* TrueNorthModel::update  src/models.cpp:724-830 */
static int truenorth_update(sfe_oracle *o, const sfe_soma_class *c, size_t i, int has_in, double in)
{
    double v = o->v[i];
    const double bias = o->bias[i];
    int state = SFE_STATUS_IDLE;
    if ((fabs(v) > 0.0) || has_in || (fabs(bias) > 0.0) || (c->flags & SFE_SOMA_FORCE_UPDATE)) state = SFE_STATUS_UPDATED;
    if (c->flags & SFE_SOMA_LEAK_TOWARDS_ZERO)
    {
        if (v > 0.0) v -= c->leak;
        else if (v < 0.0) v += c->leak;
    }
    else v += c->leak;
    v += bias;
    if (has_in) v += in;
    /* truenorth_threshold_and_reset  src/models.cpp:745-759: the jitter shifts the compared value only */
    double compared = v;
    if (c->random_mask != 0) compared += (double) (oracle_rand(o) & c->random_mask);
    if (compared >= c->threshold)
    {
        if (c->reset_mode == SFE_RESET_HARD) v = c->reset;
        else if (c->reset_mode == SFE_RESET_SOFT) v -= c->threshold;
        else if (c->reset_mode == SFE_RESET_SATURATE) v = c->threshold;
        state = SFE_STATUS_FIRED;
    }
    else if (compared <= c->reverse_threshold)
    {
        if (c->reverse_reset_mode == SFE_RESET_HARD) v = c->reverse_reset;
        else if (c->reverse_reset_mode == SFE_RESET_SOFT) v += c->reverse_threshold;
        else if (c->reverse_reset_mode == SFE_RESET_SATURATE) v = c->reverse_threshold;
    }
    o->v[i] = v;
    return state;
}

/* InputModel::update  src/models.cpp:863-903. The unit's spike-train cursor is
 * shared by the share_count neurons mapped to it: at step index s (0-based) the
 * neuron of rank r consumes element s*share_count + r. */
static int input_update(sfe_oracle *o, const sfe_input_desc *d, int64_t step_idx, int64_t timestep)
{
    const sfe_tables *t = o->t;
    int send = 0;
    const uint64_t cursor = (uint64_t) step_idx * d->share_count + d->share_rank;
    if (cursor < d->spikes_len) send = t->input_spikes[d->spikes_off + cursor] != 0;
    /* `poisson_probability > uniform_distribution(gen)`: one draw per update of the unit; the neurons
     * of a unit are updated in rank order, so drawing as we go reproduces the unit's stream. Draws of
     * units with probability 0 can never be observed and are skipped. */
    if (d->poisson > 0.0)
    {
        if (o->overlay != NULL)
        {
            const int64_t row = step_idx - o->overlay_step0;
            if (row >= 0 && row < o->overlay_steps && o->overlay[(size_t) row * o->overlay_cols + d->poisson_col]) send = 1;
        }
        else
        {
            struct mt19937 *g = &o->poisson_gen[d->unit];
            if (!g->seeded) mt_seed(g, t->input_seed_base + d->unit + 1u);
            if (d->poisson > mt_canonical(g)) send = 1;
        }
    }
    if ((d->rate > 0.0) && ((timestep % (long int) (1.0 / d->rate)) == 0)) send = 1;
    return send ? SFE_STATUS_FIRED : SFE_STATUS_IDLE;
}

/* HodgkinHuxley::update  plugins/hodgkin_huxley.cpp:116-170 */
static int hh_update(sfe_oracle *o, size_t k)
{
    const double C_m = 10.0, g_Na = 1200.0, g_K = 360.0, g_L = 3.0, V_Na = 50.0, V_K = -77.0, V_L = 54.387, dt = 0.1;
    double V = o->hh_v[k], m = o->hh_m[k], n = o->hh_n[k], h = o->hh_h[k];
    const double I = o->hh_i[k];
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
    const double denominator = g_L + g_K * (pow(n, 4)) + g_Na * (pow(m, 3) * h);
    const double tau_V = C_m / denominator;
    const double Vinf = ((g_L) * V_L + g_K * (pow(n, 4)) * V_K + g_Na * (pow(m, 3)) * h * V_Na + I) / denominator;
    const double prev_V = V;
    V = Vinf + (V - Vinf) * exp(-1 * dt / tau_V);
    m = pm + (m - pm) * exp(-1 * dt / tau_m);
    n = pn + (n - pn) * exp(-1 * dt / tau_n);
    h = ph + (h - ph) * exp(-1 * dt / tau_h);
    o->hh_v[k] = V; o->hh_m[k] = m; o->hh_n[k] = n; o->hh_h[k] = h;
    return ((prev_V < 25) && (V > 25)) ? SFE_STATUS_FIRED : SFE_STATUS_UPDATED;
}

/* NeuroFEMModel::update + process_fem  plugins/neurofem.cpp:192-317 for the call of the neuron phase (the first call of
 * a timestep): the accumulators filled by the previous message phase are consumed, then the compartment is updated.
 * sigma_v = 0 (enforced at load): the noise term is +-0 and leaves the potential unchanged. */
static int neurofem_update(sfe_oracle *o, const sfe_soma_class *c, size_t i)
{
    const double lambda_d = c->input_decay, lambda_v = c->leak, dt = c->nf_dt;
    const double acc1 = o->nf_acc[2 * i], acc2 = o->nf_acc[2 * i + 1];
    o->nf_acc[2 * i] = 0.0;
    o->nf_acc[2 * i + 1] = 0.0;
    double u1 = o->u[i], u2 = o->nf_u2[i], u_int = o->nf_uint[i], v = o->v[i];
    u1 -= lambda_d * dt * u1;
    u2 -= lambda_d * dt * u2;
    u1 += acc1;
    u2 += lambda_d * acc2;
    const double u_err = u1 + o->bias[i];
    u_int += dt * u_err;
    v = v - (lambda_v * dt * v);
    v = v + (dt * c->nf_kp * u_err) + (dt * c->nf_ki * u_int) + (dt * u2) + (0.0 * 0.0) - acc2;
    int state = SFE_STATUS_UPDATED;
    if (v > c->threshold)
    {
        v = c->reset;
        state = SFE_STATUS_FIRED;
    }
    o->u[i] = u1; o->nf_u2[i] = u2; o->nf_uint[i] = u_int; o->v[i] = v;
    return state;
}

/* MultiTapModel1D::update  src/models.cpp:237-257 with calculate_next_state (:167-202) and input_current
 * (:215-235): catch the line up to timestep T (one RC step per elapsed timestep), add the current to the
 * synapse's tap, return the voltage of tap 0. */
static double taps_update(sfe_oracle *o, uint32_t unit, int64_t T, double current, uint32_t tap)
{
    const sfe_taps_desc *d = &o->t->taps[unit];
    const size_t n = d->n_taps;
    double *v = o->tap_v + o->tap_off[unit], *next = o->tap_next;
    const double *tc = o->t->taps_values + d->const_off, *sc = tc + n;
    while (o->tap_steps[unit] < T)
    {
        ++o->tap_steps[unit];
        for (size_t k = 0; k < n; ++k) next[k] = v[k] * tc[k];
        for (size_t src = 0; src < n; ++src)
        {
            if (src > 0)
            {
                const double proximal = v[src] * sc[src - 1];
                next[src - 1] += proximal;
                next[src] -= proximal;
            }
            if (src < n - 1)
            {
                const double distal = v[src] * sc[src];
                next[src + 1] += distal;
                next[src] -= distal;
            }
        }
        for (size_t k = 0; k < n; ++k) v[k] = next[k];
    }
    v[tap] += current;
    return v[0];
}

static double potential_of(const sfe_oracle *o, size_t i)
{
    const sfe_soma_class *c = &o->t->soma_classes[o->t->neuron_class[i]];
    if (c->model == SFE_SOMA_HH) return o->hh_v[o->t->neuron_aux[i]];
    if (c->model == SFE_SOMA_INPUT) return 0.0; /* PipelineUnit::get_potential default  src/pipeline.hpp:109-112 */
    return o->v[i];
}

static void one_step(sfe_oracle *o, sfe_step_record *rec, uint32_t *fired_bits, double *potentials, uint8_t *status_out,
        double *neuron_traces)
{
    const sfe_tables *t = o->t;
    const int64_t T = ++o->total_timesteps;   /* SpikingChip::step  src/chip.cpp:549-560 */
    const int64_t steps_done = T - 1;         /* == cx.timesteps_simulated for every compartment */
    memset(rec, 0, sizeof(*rec));
    /* sim_reset_measurements  src/chip.cpp:1393-1445 */
    memset(o->tile_e, 0, t->n_tiles * sizeof(int64_t));
    memset(o->tile_w, 0, t->n_tiles * sizeof(int64_t));
    memset(o->tile_n, 0, t->n_tiles * sizeof(int64_t));
    memset(o->tile_s, 0, t->n_tiles * sizeof(int64_t));
    memset(o->core_msgs, 0, t->n_cores * sizeof(int64_t));
    memset(o->core_gen, 0, t->n_cores * sizeof(double));
    memset(o->core_proc, 0, t->n_cores * sizeof(double));
    memset(o->core_syn_e, 0, t->n_cores * sizeof(double));
    memset(o->core_den_e, 0, t->n_cores * sizeof(double));
    memset(o->core_soma_e, 0, t->n_cores * sizeof(double));
    memset(o->core_axout_e, 0, t->n_cores * sizeof(double));
    memset(o->core_dup_e, 0, t->n_cores * sizeof(double));
    if (fired_bits) memset(fired_bits, 0, ((t->n_neurons + 31) / 32) * sizeof(uint32_t));

    /* ---- process_neurons ------------------------------------------------ */
    for (uint32_t ci = 0; ci < t->n_cores; ++ci)
    {
        const sfe_core_desc *core = &t->cores[ci];
        double next_delay = 0.0; /* core.next_message_generation_delay */
        double gen_sum = 0.0;
        for (uint32_t k = 0; k < core->neuron_count; ++k)
        {
            const size_t i = core->neuron_begin + k;
            const sfe_soma_class *c = &t->soma_classes[t->neuron_class[i]];
            int has_in = 0;
            double in = 0.0;
            double lat = 0.0; /* execute_pipeline: total_latency{0.0}, += per unit */
            if (c->model == SFE_SOMA_NEUROFEM)
            {
                /* combined unit: its own double-buffered accumulators, read inside neurofem_update */
            }
            else if (c->dend_in_neuron)
            {
                /* buffer inside the dendrite unit: dendrite.update(addr, nullopt, nullopt, T) */
                if (c->dend_model == SFE_DEND_ACCUMULATOR)
                {
                    /* src/models.cpp:78-91: first touch at T zeroes the charge that the
                     * message phase of T-1 accumulated */
                    if (o->acc_step[i] < T) { o->acc[i] = 0.0; o->acc_step[i] = T; }
                    has_in = 1;
                    in = o->acc[i];
                }
                else
                {
                    /* src/models.cpp:102-117: shift the delay line up to T */
                    const size_t slot = i * RING + (size_t) (T % RING);
                    has_in = o->ring_has[slot];
                    in = o->ring[slot];
                    o->ring_has[slot] = 0;
                    o->ring[slot] = 0.0;
                }
                o->core_den_e[ci] += c->dend_energy_update;
                lat += c->dend_latency_update;
            }
            else
            {
                /* read then clear the chip-managed buffer  src/chip.cpp:713-723 */
                has_in = o->buf_has[i];
                in = o->buf[i];
                o->buf_has[i] = 0;
                o->buf[i] = 0.0;
            }
            int st;
            switch (c->model)
            {
            case SFE_SOMA_LIF: st = lif_update(o, c, i, has_in, in, steps_done); break;
            case SFE_SOMA_TRUENORTH: st = truenorth_update(o, c, i, has_in, in); break;
            case SFE_SOMA_INPUT: st = input_update(o, &t->inputs[t->neuron_aux[i]], steps_done, T); break;
            case SFE_SOMA_NEUROFEM: st = neurofem_update(o, c, i); break;
            default: st = hh_update(o, t->neuron_aux[i]); break;
            }
            o->status[i] = (uint8_t) st;
            /* calculate_soma_default_energy_latency  src/pipeline.hpp:631-714 */
            double e = c->energy_access, l = c->latency_access;
            if (st == SFE_STATUS_UPDATED || st == SFE_STATUS_FIRED)
            {
                e += c->energy_update;
                l += c->latency_update;
                rec->neurons_updated++;
            }
            if (st == SFE_STATUS_FIRED)
            {
                e += c->energy_spike_out;
                l += c->latency_spike_out;
                rec->neurons_fired++;
            }
            o->core_soma_e[ci] += e;
            if (c->flags & SFE_SOMA_IS_DENDRITE) /* src/chip.cpp:1224-1245: the unit's energy in both buckets */
            {
                o->core_den_e[ci] += e;
                o->core_dup_e[ci] += e;
            }
            lat += l;
            next_delay += lat;
            if (st == SFE_STATUS_FIRED)
            {
                if (fired_bits) fired_bits[i >> 5] |= 1u << (i & 31);
                /* pipeline_process_axon_out  src/chip.cpp:802-834 */
                for (uint32_t a = t->axon_out_begin[i]; a < t->axon_out_begin[i + 1]; ++a)
                {
                    o->axon_active[t->axon_out_target[a]] = 1;
                    o->core_axout_e[ci] += core->energy_axon_out;
                    gen_sum += next_delay + core->latency_axon_out; /* m.generation_delay */
                    next_delay = 0.0;
                    rec->packets_sent++;
                }
            }
        }
        if (next_delay != 0.0) gen_sum += next_delay; /* placeholder message  src/chip.cpp:640-652 */
        o->core_gen[ci] = gen_sum;
    }
    if (status_out) memcpy(status_out, o->status, t->n_neurons);
    if (potentials)
        for (uint32_t p = 0; p < t->n_probes; ++p) potentials[p] = potential_of(o, t->probes[p]);
    if (neuron_traces) /* sim_trace_record_neuron_traces  src/chip.cpp:1664-1702: LIF `u` of the log_u neurons */
        for (uint32_t p = 0; p < t->n_u_probes; ++p) neuron_traces[p] = o->u[t->u_probes[p]];

    /* ---- process_messages ------------------------------------------------ */
    for (uint32_t ci = 0; ci < t->n_cores; ++ci)
    {
        const sfe_core_desc *core = &t->cores[ci];
        const double *w = t->syn_weight + core->syn_begin;
        const uint32_t *meta = t->syn_meta + core->syn_begin;
        for (uint32_t k = 0; k < core->axon_in_count; ++k)
        {
            const uint32_t aid = core->axon_in_begin + k;
            if (!o->axon_active[aid]) continue;
            o->axon_active[aid] = 0;
            const sfe_axon_in *ax = &t->axons_in[aid];
            const sfe_cost_class *cc = &t->cost_classes[ax->cost_class];
            /* receive_message + sim_estimate_network_costs  src/chip.cpp:694-708, 1127-1169 */
            const int64_t dx = SFE_HOP_DX(ax->hop), dy = SFE_HOP_DY(ax->hop);
            if (SFE_HOP_EAST(ax->hop)) o->tile_e[core->tile] += dx; else o->tile_w[core->tile] += dx;
            if (SFE_HOP_NORTH(ax->hop)) o->tile_n[core->tile] += dy; else o->tile_s[core->tile] += dy;
            rec->total_hops += dx + dy;
            /* process_message  src/chip.cpp:738-764 */
            o->core_msgs[ci]++;
            double lat = core->latency_axon_in;
            for (uint32_t s = 0; s < ax->syn_count; ++s)
            {
                const double weight = w[ax->syn_off + s];
                const uint32_t m = meta[ax->syn_off + s];
                const size_t post = core->neuron_begin + SFE_SYN_POST(m);
                rec->spike_count++; /* ++spikes_processed  src/pipeline.hpp:479 */
                if (core->dend_in_msg)
                {
                    const sfe_soma_class *pc = &t->soma_classes[t->neuron_class[post]];
                    if (pc->dend_model == SFE_DEND_NEUROFEM)
                    {
                        /* next_u{1,2}_dendritic_accumulator.value_or(0.0) + current  plugins/neurofem.cpp:227-248 */
                        double *cell = &o->nf_acc[2 * post + (SFE_SYN_DELAY(m) & 1u)];
                        *cell = *cell + weight;
                    }
                    else if (pc->dend_model == SFE_DEND_TAPS)
                    {
                        /* buffer before the soma: the value the last event of the step returns is what the soma reads */
                        o->buf[post] = taps_update(o, t->neuron_taps[post], T, weight, SFE_SYN_TAP(m));
                        o->buf_has[post] = 1;
                    }
                    else if (pc->dend_model == SFE_DEND_ACCUMULATOR)
                    {
                        if (o->acc_step[post] < T) { o->acc[post] = 0.0; o->acc_step[post] = T; }
                        o->acc[post] = o->acc[post] + weight;
                        if (!pc->dend_in_neuron) { o->buf[post] = o->acc[post]; o->buf_has[post] = 1; }
                    }
                    else
                    {
                        /* next_accumulated_charges[delay] += w: reaches the soma at T+1+delay */
                        const size_t slot = post * RING + (size_t) ((T + 1 + SFE_SYN_DELAY(m)) % RING);
                        o->ring[slot] = (o->ring_has[slot] ? o->ring[slot] : 0.0) + weight;
                        o->ring_has[slot] = 1;
                    }
                }
                if (!cc->per_message)
                {
                    o->core_syn_e[ci] += cc->syn_energy;
                    o->core_den_e[ci] += cc->den_energy;
                    if (cc->den_is_soma)
                    {
                        o->core_soma_e[ci] += cc->den_energy;
                        o->core_dup_e[ci] += cc->den_energy;
                    }
                    lat += (0.0 + cc->syn_latency) + cc->den_latency;
                }
            }
            if (cc->per_message)
            {
                o->core_syn_e[ci] += cc->syn_energy;
                o->core_den_e[ci] += cc->den_energy;
                if (cc->den_is_soma)
                {
                    o->core_soma_e[ci] += cc->den_energy;
                    o->core_dup_e[ci] += cc->den_energy;
                }
                lat += cc->syn_latency;
            }
            o->core_proc[ci] += lat; /* message_processing_latencies[dest_core_id]  src/schedule.cpp:82 */
        }
    }

    /* ---- sim_calculate_ts_energy  src/chip.cpp:1171-1261 -------------------- */
    {
        uint32_t ci = 0;
        for (uint32_t ti = 0; ti < t->n_tiles; ++ti)
        {
            const sfe_tile_desc *tile = &t->tiles[ti];
            double tile_energy = (double) o->tile_e[ti] * tile->energy_east;
            tile_energy += (double) o->tile_w[ti] * tile->energy_west;
            tile_energy += (double) o->tile_s[ti] * tile->energy_south;
            tile_energy += (double) o->tile_n[ti] * tile->energy_north;
            rec->network_energy += tile_energy;
            while (ci < t->n_cores && t->cores[ci].tile == ti)
            {
                const sfe_core_desc *core = &t->cores[ci];
                const double axon_in_energy = (double) o->core_msgs[ci] * core->energy_axon_in;
                rec->network_energy += axon_in_energy;
                const double pipeline = ((o->core_syn_e[ci] + o->core_den_e[ci]) + o->core_soma_e[ci]) - o->core_dup_e[ci];
                rec->synapse_energy += o->core_syn_e[ci];
                rec->dendrite_energy += o->core_den_e[ci];
                rec->soma_energy += o->core_soma_e[ci];
                rec->network_energy += o->core_axout_e[ci];
                double core_energy = axon_in_energy;
                core_energy += pipeline;
                core_energy += o->core_axout_e[ci];
                tile_energy += core_energy;
                ++ci;
            }
            rec->total_energy += tile_energy;
        }
    }
    /* ---- schedule_messages_timestep_simple  src/schedule.cpp:61-102 ---------- */
    {
        double max_gen = 0.0, max_proc = 0.0;
        for (uint32_t ci = 0; ci < t->n_cores; ++ci)
        {
            if (o->core_gen[ci] > max_gen) max_gen = o->core_gen[ci];
            if (o->core_proc[ci] > max_proc) max_proc = o->core_proc[ci];
        }
        rec->sim_time = (max_proc > max_gen ? max_proc : max_gen) + t->sync_delay;
    }
    o->total_sim_time += rec->sim_time;
    o->total_energy += rec->total_energy;
}

int sfe_oracle_run(sfe_oracle *o, int64_t timesteps, const sfe_trace_request *req, sfe_run_data *out)
{
    const sfe_tables *t = o->t;
    const size_t words = (t->n_neurons + 31) / 32;
    sfe_run_data rd;
    memset(&rd, 0, sizeof(rd));
    rd.timestep_start = o->total_timesteps + 1; /* RunData rd(total_timesteps + 1)  src/chip.cpp:481 */
    rd.timesteps_executed = timesteps;
    for (int64_t s = 0; s < timesteps; ++s)
    {
        sfe_step_record rec;
        one_step(o, &rec, (req && req->fired_bits) ? req->fired_bits + (size_t) s * words : NULL,
                (req && req->potentials) ? req->potentials + (size_t) s * t->n_probes : NULL,
                (req && req->status) ? req->status + (size_t) s * t->n_neurons : NULL,
                (req && req->neuron_traces) ? req->neuron_traces + (size_t) s * t->n_u_probes : NULL);
        if (req && req->steps) req->steps[s] = rec;
        /* update_run_data  src/chip.cpp:462-475 */
        rd.total_energy += rec.total_energy;
        rd.synapse_energy += rec.synapse_energy;
        rd.dendrite_energy += rec.dendrite_energy;
        rd.soma_energy += rec.soma_energy;
        rd.network_energy += rec.network_energy;
        rd.sim_time += rec.sim_time;
        rd.spikes += rec.spike_count;
        rd.packets_sent += rec.packets_sent;
        rd.neurons_updated += rec.neurons_updated;
        rd.neurons_fired += rec.neurons_fired;
    }
    if (out) *out = rd;
    return 0;
}

double sfe_oracle_get_power(const sfe_oracle *o) /* src/chip.cpp:607-621 */
{
    return o->total_sim_time > 0.0 ? o->total_energy / o->total_sim_time : 0.0;
}

void sfe_oracle_read_potentials(const sfe_oracle *o, double *out)
{
    for (size_t i = 0; i < o->t->n_neurons; ++i) out[i] = potential_of(o, i);
}
