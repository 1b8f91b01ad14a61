// Force-included (-include) ahead of every reference translation unit. The
// reference's yaml_arch.hpp / yaml_snn.hpp need RapidYAML (network-fetched,
// absent offline); their include guards are pre-defined on the command line and
// the four entry points the engine TUs reference are declared here instead
// (src/arch.cpp:116, src/network.cpp:218,703-704). Definitions: ref_harness.cpp.
#pragma once
#include <filesystem>
#include <fstream>
namespace sanafe
{
class Architecture;
class SpikingNetwork;
Architecture description_parse_arch_file_yaml(std::ifstream &fp);
SpikingNetwork yaml_parse_network_file(std::ifstream &fp, Architecture &arch);
void yaml_write_network(std::filesystem::path path, const SpikingNetwork &network);
void yaml_write_mappings_file(std::filesystem::path path, const SpikingNetwork &network);
}
