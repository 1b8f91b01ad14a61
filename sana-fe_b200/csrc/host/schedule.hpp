// schedule.hpp — host-side "detailed" timing model (see schedule.cpp).
#ifndef SFE_SCHEDULE_HPP_
#define SFE_SCHEDULE_HPP_

#include <cstdint>
#include <memory>
#include <vector>

#include "sanafe_b200.h"

namespace sfe
{
// One row of the reference's message trace (src/chip.cpp:1731-1764).
struct MessageRecord
{
    long mid{-1};           // placeholder_mid for the trailing "processing only" entry of a core
    bool placeholder{true};
    uint32_t src_neuron{0}; // device index (a placeholder names the last neuron of its core)
    uint32_t src_core{0}, dest_core{0};
    uint32_t hops{0}, spikes{0};
    double sent, received, processed; // -inf unless the detailed model filled them
    double generation_delay{0.0}, processing_delay{0.0}, network_delay{0.0}, blocking_delay{0.0};
    double min_hop_delay{0.0}, messages_along_route{0.0};
};

class DetailedScheduler
{
public:
    explicit DetailedScheduler(const sfe_tables &t);
    ~DetailedScheduler();
    // sim_time of one timestep from the per-neuron status bytes of that step
    double schedule_step(const uint8_t *status) { return run_step(status, true, nullptr, nullptr); }
    // The step's messages in the order the reference traces them (sim_sort_and_record_messages,
    // src/chip.cpp:439-456: by message id, placeholders last). detailed = run the NoC scheduler so
    // that the timestamps are filled; *next_mid is the chip-lifetime message counter
    // (total_messages_sent, src/chip.cpp:815; ids follow the single-thread creation order).
    double trace_step(const uint8_t *status, bool detailed, std::vector<MessageRecord> &out, long &next_mid)
    {
        return run_step(status, detailed, &out, &next_mid);
    }

private:
    double run_step(const uint8_t *status, bool schedule, std::vector<MessageRecord> *trace, long *next_mid);
    const sfe_tables &t_;
    std::vector<uint32_t> axon_core_; // destination core of every axon-in
    std::vector<double> axon_proc_;   // processing delay of the message that targets it
    struct Scratch;                   // per-step work arrays, kept between steps (schedule.cpp)
    std::unique_ptr<Scratch> scratch_;
};
} // namespace sfe
#endif
