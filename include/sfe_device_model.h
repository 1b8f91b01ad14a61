/* sfe_device_model.h — out-of-tree hardware-unit models for the B200 SANA-FE engine.
 *
 * What it replaces: the reference loads a hardware-unit plugin as host C++ — `dlopen(plugin)` +
 * `extern "C" PipelineUnit *create_<model>()` (src/plugins.cpp:45-98) — and calls its virtuals
 * (src/pipeline.hpp:69-301: set_attribute_hw / set_attribute_neuron / update / reset / get_potential) once per
 * neuron per timestep. A device engine cannot call host virtuals per neuron, so a plugin model here is DEVICE code:
 * the plugin library is built with nvcc from a functor (the `update` arithmetic) plus the kernel template and the
 * SFE_DEVICE_SOMA_MODEL macro of this header, and exports
 *
 *     extern "C" const sfe_device_model_desc *sfe_device_model_<model>(void);     // <-> create_<model>()
 *
 * The engine finds it exactly where the reference finds `create_<model>`: a hardware unit of the architecture
 * description names `model: <model>` and `plugin: <path to the library>`; at load() the library is dlopen'ed and the
 * symbol looked up. A model can also be registered programmatically (sfe_register_device_model), which is what a
 * host application embedding the engine would do. A registered model takes precedence over the models compiled into
 * the engine under the same name (the shipped example registers the reference's Hodgkin-Huxley plugin out of tree:
 * sana-fe_b200/plugins/hodgkin_huxley_device.cu).
 *
 * Scope of this version: SOMA units (SomaUnit, src/pipeline.hpp:469-489) — update(neuron_address, current_in,
 * timestep) -> status. Synapse and dendrite units keep the built-in models (their arithmetic is fused into the
 * message-phase kernel).
 *
 * Data model. Every neuron mapped to a unit of the model is one INSTANCE with `n_state` fp64 state words and
 * `n_params` fp64 parameters, stored SoA on the device: word w of instance k is at `base[w * n_instances + k]`.
 *   - set_attribute_neuron(name, value) at load time fills the state word / parameter whose name matches
 *     (state_names / param_names); other names are ignored, as PipelineUnit subclasses ignore unknown keys.
 *     set_attribute_hw (attributes of the unit in the architecture file) is applied the same way first.
 *   - reset() zeroes the state words selected by reset_mask (src/chip.cpp:576-600 calls every unit's reset()).
 *   - get_potential() returns state word `potential_state` (potential traces), or 0.0 when it is < 0.
 *   A reference plugin that keeps ONE state per hardware unit whatever neuron is updated (the shipped
 *   hodgkin_huxley.cpp does) is reproduced with one neuron per unit.
 *
 * Step protocol (engine side, csrc/engine.cu: launch_device_models): before the neuron-phase kernel of a timestep
 * the engine gathers, for every instance, the dendrite output of the step (current_in / has_in — the
 * std::optional<double> of update()), then calls `launch` once per model with plain device pointers and the engine's
 * stream; the model's kernel writes one status byte per instance (SFE_STATUS_IDLE / UPDATED / FIRED = NeuronStatus
 * idle / updated / fired). The neuron-phase kernel then treats the instance like any other soma: default
 * energy/latency costs by status (src/pipeline.hpp:631-714), spike raster, axon-out messages.
 */
#ifndef SFE_DEVICE_MODEL_H_
#define SFE_DEVICE_MODEL_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFE_DEVICE_MODEL_ABI 1

/* values of the status bytes a model writes (same numbering as SFE_STATUS_* of sanafe_b200.h) */
#define SFE_MODEL_IDLE 1
#define SFE_MODEL_UPDATED 2
#define SFE_MODEL_FIRED 3

typedef struct sfe_device_model_launch
{
    void *stream;             /* cudaStream_t the kernel must be enqueued on */
    uint32_t n_instances;     /* neurons of this model on this engine */
    uint32_t n_state, n_params;
    uint32_t pad;
    double *state;            /* [n_state][n_instances], device memory */
    const double *params;     /* [n_params][n_instances], device memory */
    const double *current_in; /* [n_instances] dendrite output of this timestep (meaningful where has_in != 0) */
    const uint8_t *has_in;    /* [n_instances] 1: current_in holds a value (std::optional has_value) */
    uint8_t *status;          /* [n_instances] out: SFE_MODEL_IDLE / UPDATED / FIRED */
    int64_t timestep;         /* 1-based, the `timestep` argument of PipelineUnit::update */
} sfe_device_model_launch;

typedef struct sfe_device_model_desc
{
    uint32_t abi_version;            /* SFE_DEVICE_MODEL_ABI */
    uint32_t n_state, n_params;      /* fp64 words per instance; n_state <= 32 */
    uint32_t reset_mask;             /* bit w: state word w is zeroed by reset() */
    int32_t potential_state;         /* state word returned by get_potential(), -1: 0.0 */
    uint32_t pad;
    const char *const *state_names;  /* [n_state] attribute name that initialises the word (NULL / "" = none) */
    const char *const *param_names;  /* [n_params] */
    const double *state_init;        /* [n_state] defaults, NULL = zeros */
    const double *param_init;        /* [n_params] defaults, NULL = zeros */
    /* enqueue the model's update kernel for one timestep; 0 on success */
    int (*launch)(const sfe_device_model_launch *args);
} sfe_device_model_desc;

/* Registers `desc` (which must outlive its use) under `model_name`; replaces an earlier registration of the name.
 * Returns 0, or -1 with sfe_last_error() set (bad ABI version, too many state words, no launch function). */
int sfe_register_device_model(const char *model_name, const sfe_device_model_desc *desc);
/* Removes a registration (chips already loaded keep using the descriptor). 0, or -1 if the name is unknown. */
int sfe_unregister_device_model(const char *model_name);
/* dlopen(`library_path`) and register what `sfe_device_model_<model_name>()` returns: the counterpart of the
 * reference's plugin_init_hw (src/plugins.cpp:45-83). load() does this itself for units with a `plugin:` path. */
int sfe_load_device_model(const char *model_name, const char *library_path);
/* 1 if a model of that name is registered */
int sfe_device_model_registered(const char *model_name);

#ifdef __cplusplus
}
#endif

/* ------------------------------------------------------------------------------------------------------------
 * Plugin-author side (CUDA): write a functor, instantiate the model with SFE_DEVICE_SOMA_MODEL.
 *
 *   struct MyModel
 *   {
 *       static constexpr int kState = 2, kParams = 1;
 *       static constexpr unsigned kResetMask = 0x3;          // reset() zeroes both state words
 *       static constexpr int kPotentialState = 0;
 *       static const char *state_name(int w) { static const char *n[] = {"v", "w"}; return n[w]; }
 *       static const char *param_name(int p) { static const char *n[] = {"gain"}; return n[p]; }
 *       static double state_init(int) { return 0.0; }
 *       static double param_init(int) { return 1.0; }
 *       // one neuron, one timestep: `s` / `p` index the instance's state words / parameters
 *       __device__ static int update(sfe_model_words<double> s, sfe_model_words<const double> p, bool has_in, double in,
 *               long long timestep);                          // returns SFE_MODEL_IDLE / UPDATED / FIRED
 *   };
 *   SFE_DEVICE_SOMA_MODEL(my_model, MyModel)
 *
 * Build: nvcc -shared -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -fmad=false -I<repo>/include -o libmy_model.so my_model.cu
 * (-fmad=false: the reference's x86-64 build has no FMA contraction; keep it for bit-equal potentials).
 * ------------------------------------------------------------------------------------------------------------ */
#ifdef __CUDACC__
#include <cuda_runtime.h>

template <typename T> struct sfe_model_words /* word w of this instance: base[w * stride] */
{
    T *base;
    uint32_t stride;
    __device__ __forceinline__ T &operator[](const int w) const { return base[static_cast<size_t>(w) * stride]; }
};

template <class Model> __global__ void __launch_bounds__(128) sfe_device_soma_kernel(const sfe_device_model_launch a)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.n_instances) return;
    const sfe_model_words<double> s = {a.state + k, a.n_instances};
    const sfe_model_words<const double> p = {a.params + k, a.n_instances};
    a.status[k] = static_cast<uint8_t>(Model::update(s, p, a.has_in[k] != 0, a.current_in[k], static_cast<long long>(a.timestep)));
}

template <class Model> int sfe_device_soma_launch(const sfe_device_model_launch *a)
{
    if (a->n_instances == 0) return 0;
    sfe_device_soma_kernel<Model><<<(a->n_instances + 127) / 128, 128, 0, static_cast<cudaStream_t>(a->stream)>>>(*a);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

#define SFE_DEVICE_SOMA_MODEL(name, Model)                                                                        \
    extern "C" const sfe_device_model_desc *sfe_device_model_##name(void)                                         \
    {                                                                                                             \
        static const char *state_names[Model::kState > 0 ? Model::kState : 1];                                    \
        static const char *param_names[Model::kParams > 0 ? Model::kParams : 1];                                  \
        static double state_init[Model::kState > 0 ? Model::kState : 1];                                          \
        static double param_init[Model::kParams > 0 ? Model::kParams : 1];                                        \
        static sfe_device_model_desc desc;                                                                        \
        for (int w = 0; w < Model::kState; ++w)                                                                   \
        {                                                                                                         \
            state_names[w] = Model::state_name(w);                                                                \
            state_init[w] = Model::state_init(w);                                                                 \
        }                                                                                                         \
        for (int p = 0; p < Model::kParams; ++p)                                                                  \
        {                                                                                                         \
            param_names[p] = Model::param_name(p);                                                                \
            param_init[p] = Model::param_init(p);                                                                 \
        }                                                                                                         \
        desc.abi_version = SFE_DEVICE_MODEL_ABI;                                                                  \
        desc.n_state = Model::kState;                                                                             \
        desc.n_params = Model::kParams;                                                                           \
        desc.reset_mask = Model::kResetMask;                                                                      \
        desc.potential_state = Model::kPotentialState;                                                            \
        desc.state_names = state_names;                                                                           \
        desc.param_names = param_names;                                                                           \
        desc.state_init = state_init;                                                                             \
        desc.param_init = param_init;                                                                             \
        desc.launch = &sfe_device_soma_launch<Model>;                                                             \
        return &desc;                                                                                             \
    }
#endif /* __CUDACC__ */

#endif /* SFE_DEVICE_MODEL_H_ */
