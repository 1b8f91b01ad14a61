// yaml_frontend.cpp — Architecture / SNN YAML readers (placeholder until the
// in-tree YAML subset reader lands; see DESIGN.md "next" rows).
#include "desc.hpp"

namespace sfe
{
std::unique_ptr<Architecture> load_arch_yaml(const std::string &path)
{
    throw std::runtime_error("YAML architecture reader not built yet (" + path + "); use the flat description");
}
std::unique_ptr<SpikingNetwork> load_net_yaml(const std::string &path, Architecture &)
{
    throw std::runtime_error("YAML network reader not built yet (" + path + "); use the flat description");
}
} // namespace sfe
