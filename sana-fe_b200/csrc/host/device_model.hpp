// device_model.hpp — host registry of out-of-tree device models (include/sfe_device_model.h, src/plugins.cpp:22-98)
#ifndef SFE_DEVICE_MODEL_HPP_
#define SFE_DEVICE_MODEL_HPP_

#include <string>

struct sfe_device_model_desc;

namespace sfe
{
// the descriptor registered under `name`, or null
const sfe_device_model_desc *find_device_model(const std::string &name);
// dlopen(path) + sfe_device_model_<name>() + registration; null with the reason in *why
const sfe_device_model_desc *load_device_model(const std::string &name, const std::string &path, std::string *why);
} // namespace sfe
#endif
