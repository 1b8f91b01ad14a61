// engine.hpp — internal (non-ABI) hooks between the host layer and engine.cu.
#ifndef SFE_ENGINE_HPP_
#define SFE_ENGINE_HPP_

#include <string>

#include "sanafe_b200.h"

namespace sfe
{
void set_last_error(const std::string &msg);
}
// re-upload the soma parameter classes + per-neuron class ids (MappedNeuron.set_attributes)
int sfe_engine_update_classes(sfe_engine *e, const sfe_soma_class *classes, uint32_t n_classes,
        const uint32_t *neuron_class);
#endif
