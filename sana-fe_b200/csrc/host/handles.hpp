// handles.hpp — what the C ABI's opaque description handles hold. Internal to the package: the C++
// front-ends that build descriptions object by object (the pybind11 module mirrors the reference's
// Python builder classes, src/pymodule.cpp:850-1173) reach the description through these, everything
// else goes through include/sanafe_b200.h.
#ifndef SFE_HANDLES_HPP_
#define SFE_HANDLES_HPP_

#include <memory>
#include <optional>

#include "desc.hpp"

struct sfe_arch
{
    std::unique_ptr<sfe::Architecture> arch;
};
struct sfe_net
{
    std::unique_ptr<sfe::SpikingNetwork> net;
    std::optional<sfe::SynthRequest> synth;
};

namespace sfe
{
// Network.save(path)  src/network.cpp:693-712 (YAML format only)
void save_net_yaml(const SpikingNetwork &net, const std::string &path);
void save_net_netlist(const SpikingNetwork &net, const std::string &path); // legacy format (lossy, as the reference's)
} // namespace sfe
#endif
