// device_model.cpp — registry of out-of-tree device models (include/sfe_device_model.h).
//
// The counterpart of the reference's plugin loader (src/plugins.cpp:22-98): there, `plugin_get_hw(model, path)`
// dlopen's the library once per model name, looks up `create_<model>` and calls it for every unit instance. Here the
// library exports `sfe_device_model_<model>()`, which returns a descriptor of DEVICE code (state layout, attribute
// names, a launch function); the descriptor is registered under the model name and the lowering resolves units
// against the registry.
#include <dlfcn.h>

#include <map>
#include <mutex>
#include <string>

#include "device_model.hpp"
#include "sanafe_b200.h"
#include "sfe_device_model.h"

namespace sfe
{
void set_last_error(const std::string &msg);

namespace
{
std::mutex g_mu;
std::map<std::string, const sfe_device_model_desc *> &registry()
{
    static std::map<std::string, const sfe_device_model_desc *> r;
    return r;
}
std::map<std::string, void *> &libraries() // path -> dlopen handle (kept open for the life of the process)
{
    static std::map<std::string, void *> l;
    return l;
}

std::string check_desc(const sfe_device_model_desc *d)
{
    if (d == nullptr) return "null descriptor";
    if (d->abi_version != SFE_DEVICE_MODEL_ABI)
        return "descriptor ABI version " + std::to_string(d->abi_version) + ", this engine speaks " + std::to_string(SFE_DEVICE_MODEL_ABI);
    if (d->n_state > 32) return "more than 32 state words per instance";
    if (d->launch == nullptr) return "no launch function";
    if (d->potential_state >= static_cast<int32_t>(d->n_state)) return "potential_state outside the state words";
    if ((d->n_state > 0 && d->state_names == nullptr) || (d->n_params > 0 && d->param_names == nullptr)) return "missing attribute names";
    return "";
}
} // namespace

const sfe_device_model_desc *find_device_model(const std::string &name)
{
    const std::lock_guard<std::mutex> lock(g_mu);
    const auto it = registry().find(name);
    return it == registry().end() ? nullptr : it->second;
}

// plugin_init_hw  src/plugins.cpp:45-83: load the library, look the factory symbol up, remember it under the model name
const sfe_device_model_desc *load_device_model(const std::string &name, const std::string &path, std::string *why)
{
    void *lib = nullptr;
    {
        const std::lock_guard<std::mutex> lock(g_mu);
        const auto it = libraries().find(path);
        if (it != libraries().end()) lib = it->second;
    }
    if (lib == nullptr)
    {
        lib = dlopen(path.c_str(), RTLD_LAZY | RTLD_LOCAL);
        if (lib == nullptr)
        {
            if (why != nullptr) *why = std::string("could not load library: ") + dlerror();
            return nullptr;
        }
        const std::lock_guard<std::mutex> lock(g_mu);
        libraries()[path] = lib;
    }
    dlerror();
    using Factory = const sfe_device_model_desc *(*) (void);
    const std::string symbol = "sfe_device_model_" + name;
    const Factory factory = reinterpret_cast<Factory>(dlsym(lib, symbol.c_str()));
    if (factory == nullptr)
    {
        const bool host_plugin = dlsym(lib, ("create_" + name).c_str()) != nullptr;
        if (why != nullptr)
            *why = "the library does not export " + symbol +
                    (host_plugin ? " (it is a host plugin of the reference, create_" + name + "(): host virtuals cannot run on the device; "
                                   "build the model with include/sfe_device_model.h)"
                                 : "");
        return nullptr;
    }
    const sfe_device_model_desc *d = factory();
    const std::string bad = check_desc(d);
    if (!bad.empty())
    {
        if (why != nullptr) *why = symbol + "(): " + bad;
        return nullptr;
    }
    const std::lock_guard<std::mutex> lock(g_mu);
    registry()[name] = d;
    return d;
}

} // namespace sfe

extern "C" int sfe_register_device_model(const char *model_name, const sfe_device_model_desc *desc)
{
    if (model_name == nullptr || *model_name == '\0')
    {
        sfe::set_last_error("sfe_register_device_model: empty model name");
        return -1;
    }
    const std::string bad = sfe::check_desc(desc);
    if (!bad.empty())
    {
        sfe::set_last_error("sfe_register_device_model(" + std::string(model_name) + "): " + bad);
        return -1;
    }
    const std::lock_guard<std::mutex> lock(sfe::g_mu);
    sfe::registry()[model_name] = desc;
    return 0;
}

extern "C" int sfe_unregister_device_model(const char *model_name)
{
    const std::lock_guard<std::mutex> lock(sfe::g_mu);
    if (model_name == nullptr || sfe::registry().erase(model_name) == 0)
    {
        sfe::set_last_error("sfe_unregister_device_model: no such model");
        return -1;
    }
    return 0;
}

extern "C" int sfe_load_device_model(const char *model_name, const char *library_path)
{
    if (model_name == nullptr || library_path == nullptr)
    {
        sfe::set_last_error("sfe_load_device_model: null argument");
        return -1;
    }
    std::string why;
    if (sfe::load_device_model(model_name, library_path, &why) == nullptr)
    {
        sfe::set_last_error("sfe_load_device_model(" + std::string(model_name) + ", " + library_path + "): " + why);
        return -1;
    }
    return 0;
}

extern "C" int sfe_device_model_registered(const char *model_name)
{
    return model_name != nullptr && sfe::find_device_model(model_name) != nullptr ? 1 : 0;
}
