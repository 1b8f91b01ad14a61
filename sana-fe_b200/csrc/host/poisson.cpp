// poisson.cpp — host-side source of the Poisson input spikes.
//
// InputModel::update (src/models.cpp:863-903) draws `uniform_distribution(gen)` on EVERY update of
// the unit and spikes when `poisson > U`. `gen` is a std::mt19937 seeded with the 1-based,
// process-wide construction index of the unit (src/models.hpp:347,366) and the distribution is
// libstdc++'s std::uniform_real_distribution<double>{0,1} (two 32-bit draws per value). This file
// is compiled with the same libstdc++, so using the very same standard classes reproduces the
// reference's stream bit for bit. The draws reach the device as a byte overlay
// (sfe_engine_set_input_overlay): one row per timestep, one column per Poisson neuron.
//
// A unit shared by `share_count` neurons is updated share_count times per step, in in-core order:
// the neuron of rank r sees draw number step*share_count + r of its unit.
#include <algorithm>
#include <map>
#include <memory>
#include <random>
#include <stdexcept>
#include <vector>

#include "engine.hpp"
#include "glibc_rand.hpp"
#include "mt19937.cuh"
#include "sanafe_b200.h"

struct sfe_poisson
{
    struct Unit
    {
        std::mt19937 gen;
        std::uniform_real_distribution<double> uniform{0.0, 1.0};
        double probability{0.0};
        uint32_t share_count{0};
        std::vector<uint32_t> col_of_rank; // overlay column of the neuron of rank r, or 0xFFFFFFFF
    };
    std::vector<Unit> units; // only units with poisson > 0 (the draws of the others are never observed)
    uint32_t n_cols{0};
};

extern "C" sfe_poisson *sfe_poisson_create(const sfe_tables *t)
{
    try
    {
        if (t == nullptr) throw std::invalid_argument("sfe_poisson_create: null tables");
        auto p = std::make_unique<sfe_poisson>();
        p->n_cols = t->n_poisson_cols;
        std::map<uint32_t, size_t> index; // unit ordinal -> slot
        for (uint32_t k = 0; k < t->n_inputs; ++k)
        {
            const sfe_input_desc &d = t->inputs[k];
            if (!(d.poisson > 0.0)) continue;
            if (d.poisson_col >= t->n_poisson_cols)
                throw std::invalid_argument("sfe_poisson_create: poisson_col out of range");
            auto [it, fresh] = index.try_emplace(d.unit, p->units.size());
            if (fresh)
            {
                sfe_poisson::Unit u;
                u.gen.seed(t->input_seed_base + d.unit + 1u); // InputModel() : gen(++instance_counter)
                u.probability = d.poisson;
                u.share_count = d.share_count;
                u.col_of_rank.assign(d.share_count, 0xFFFFFFFFu);
                p->units.push_back(std::move(u));
            }
            sfe_poisson::Unit &u = p->units[it->second];
            if (d.share_count != u.share_count || d.share_rank >= u.share_count || d.poisson != u.probability)
                throw std::invalid_argument("sfe_poisson_create: neurons of one input unit disagree about the unit");
            u.col_of_rank[d.share_rank] = d.poisson_col;
        }
        return p.release();
    }
    catch (const std::exception &e)
    {
        sfe::set_last_error(e.what());
        return nullptr;
    }
}

extern "C" void sfe_poisson_destroy(sfe_poisson *p)
{
    delete p;
}

extern "C" uint32_t sfe_poisson_cols(const sfe_poisson *p)
{
    return p != nullptr ? p->n_cols : 0u;
}

extern "C" int sfe_poisson_fill(sfe_poisson *p, uint8_t *bits, int64_t n_steps)
{
    if (p == nullptr || (bits == nullptr && p->n_cols != 0 && n_steps > 0) || n_steps < 0)
    {
        sfe::set_last_error("sfe_poisson_fill: bad arguments");
        return -1;
    }
    const size_t cols = p->n_cols;
    std::fill(bits, bits + static_cast<size_t>(n_steps) * cols, uint8_t{0});
    for (sfe_poisson::Unit &u : p->units)
        for (int64_t s = 0; s < n_steps; ++s)
            for (uint32_t r = 0; r < u.share_count; ++r)
            {
                const bool spike = u.probability > u.uniform(u.gen);
                const uint32_t col = u.col_of_rank[r];
                if (col != 0xFFFFFFFFu) bits[static_cast<size_t>(s) * cols + col] = spike ? 1 : 0;
            }
    return 0;
}

// ---- cross-check hooks (tests): the same seed through libstdc++ and through mt19937.cuh's host compilation ----
extern "C" void sfe_poisson_reference_draws(uint32_t seed, double *out, size_t n)
{
    std::mt19937 gen(seed);
    std::uniform_real_distribution<double> uniform{0.0, 1.0};
    for (size_t k = 0; k < n; ++k) out[k] = uniform(gen);
}

extern "C" void sfe_mt19937_draws(uint32_t seed, double *out, size_t n, size_t stride)
{
    if (stride == 0) stride = 1;
    std::vector<uint32_t> mt(static_cast<size_t>(sfe::kMtWords) * stride, 0u);
    uint32_t idx = 0;
    sfe::mt_seed(mt.data(), stride, &idx, seed);
    for (size_t k = 0; k < n; ++k) out[k] = sfe::mt_canonical(mt.data(), stride, &idx);
}

// glibc's rand() / srand() (stdlib/random_r.c, TYPE_3: x^31 + x^3 + 1 additive feedback over 32-bit words, seeded by the
// Park-Miller LCG, first 310 outputs discarded, result = word >> 1) restated, so that the TrueNorth threshold jitter
// (`std::rand() & random_mask`, src/models.cpp:757) does not depend on - or disturb - the host process's own generator.
namespace sfe
{
GlibcRand::GlibcRand(const uint32_t seed)
{
    int32_t r[34];
    r[0] = static_cast<int32_t>(seed == 0u ? 1u : seed);
    for (int i = 1; i < 31; ++i)
    {
        // r[i] = (16807 * r[i-1]) % 2147483647 without overflow (Schrage), negative results wrapped
        const long hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
        long word = 16807 * lo - 2836 * hi;
        if (word < 0) word += 2147483647;
        r[i] = static_cast<int32_t>(word);
    }
    for (int i = 0; i < 31; ++i) state_[i] = static_cast<uint32_t>(r[i]);
    front_ = 3;
    rear_ = 0;
    for (int i = 0; i < 310; ++i) (void) next();
}

uint32_t GlibcRand::next()
{
    state_[front_] += state_[rear_];
    const uint32_t result = state_[front_] >> 1;
    front_ = (front_ + 1) % 31;
    rear_ = (rear_ + 1) % 31;
    return result;
}
} // namespace sfe

extern "C" void sfe_glibc_rand_draws(uint32_t seed, uint64_t skip, uint32_t *out, size_t n)
{
    sfe::GlibcRand gen(seed);
    for (uint64_t k = 0; k < skip; ++k) (void) gen.next();
    for (size_t k = 0; k < n; ++k) out[k] = gen.next();
}
