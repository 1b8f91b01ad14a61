// partition.cpp — how a chip is split over the GPUs of one box.
//
// The reference simulates all cores in one process (src/chip.cpp:586-618 loops over
// every core under OpenMP); here each rank owns a contiguous range of cores (core-major
// order = the order of the lowered tables), balanced by the synapse + neuron work of
// the range. The spikes of a step travel between ranks as a fired-bit raster with one
// equal-sized slice per rank, so that the exchange is a plain all-gather.
#include <algorithm>
#include <cmath>
#include <vector>

#include "../engine.hpp"
#include "sanafe_b200.h"

extern "C" int sfe_plan_partition(const sfe_tables *tb, uint32_t world, uint32_t *owner, uint32_t *fired_word_begin,
        uint32_t *slice_words)
{
    if (tb == nullptr || world == 0)
    {
        sfe::set_last_error("sfe_plan_partition: need tables and world >= 1");
        return -1;
    }
    std::vector<uint32_t> own(tb->n_cores, 0);
    if (world > 1)
    {
        auto weight = [&](uint32_t c) {
            return static_cast<double>(tb->cores[c].syn_count) + 64.0 * tb->cores[c].neuron_count;
        };
        double total = 0.0;
        for (uint32_t c = 0; c < tb->n_cores; ++c) total += weight(c);
        double before = 0.0;
        uint32_t last = 0;
        for (uint32_t c = 0; c < tb->n_cores; ++c)
        {
            // a core goes to the rank its midpoint of cumulative work falls into; never backwards
            const double w = weight(c);
            if (w > 0.0 && total > 0.0)
                last = std::max(last, std::min(world - 1, static_cast<uint32_t>((before + 0.5 * w) * world / total)));
            own[c] = last;
            before += w;
        }
    }
    // raster: one equal-sized slice per rank, every core starts on a word boundary
    std::vector<uint32_t> words(world, 0);
    for (uint32_t c = 0; c < tb->n_cores; ++c) words[own[c]] += (tb->cores[c].neuron_count + 31) / 32;
    // slices are multiples of 4 words: ranks push them to their peers with 16-byte stores
    const uint32_t slice = (std::max<uint32_t>(1, *std::max_element(words.begin(), words.end())) + 3u) & ~3u;
    std::vector<uint32_t> fill(world, 0);
    for (uint32_t c = 0; c < tb->n_cores; ++c)
    {
        if (owner != nullptr) owner[c] = own[c];
        if (fired_word_begin != nullptr) fired_word_begin[c] = own[c] * slice + fill[own[c]];
        fill[own[c]] += (tb->cores[c].neuron_count + 31) / 32;
    }
    if (slice_words != nullptr) *slice_words = slice;
    return 0;
}
