// schedule.hpp — host-side "detailed" timing model (see schedule.cpp).
#ifndef SFE_SCHEDULE_HPP_
#define SFE_SCHEDULE_HPP_

#include <cstdint>
#include <vector>

#include "sanafe_b200.h"

namespace sfe
{
class DetailedScheduler
{
public:
    explicit DetailedScheduler(const sfe_tables &t);
    // sim_time of one timestep from the per-neuron status bytes of that step
    double schedule_step(const uint8_t *status);

private:
    const sfe_tables &t_;
    std::vector<uint32_t> axon_core_; // destination core of every axon-in
    std::vector<double> axon_proc_;   // processing delay of the message that targets it
};
} // namespace sfe
#endif
