"""sanafe_b200 — Python face of the B200-native SANA-FE time-step engine.

A thin ctypes binding over the C ABI in include/sanafe_b200.h (libsanafe_b200.so,
built in-tree by sana-fe_b200/Makefile). The class and method names follow the
reference's Python module (src/pymodule.cpp:850-1213): `load_arch`, `load_net`,
`SpikingChip(arch).load(net)`, `.sim(timesteps, timing_model=..., spike_trace=...,
potential_trace=..., perf_trace=...) -> dict`, `.reset()`, `.get_power()`.

There is no CPU fallback: constructing a SpikingChip on a machine without a CUDA
device raises. (`device=-1` builds a host-only chip that can lower a network and
export its tables — used by the CPU-side tests — but cannot simulate.)
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SFE_LIB_PATH: an experiment build of the same library (sana-fe_b200/Makefile, target `variant`)
_LIB_PATH = os.environ.get("SFE_LIB_PATH") or os.path.join(_HERE, "libsanafe_b200.so")


class SynthSpec(C.Structure):
    """sfe_synth_spec (include/sfe_synth.h)."""
    _fields_ = [("cores", C.c_uint32), ("neurons_per_core", C.c_uint32), ("dest_cores", C.c_uint32),
                ("syn_per_axon", C.c_uint32), ("seed", C.c_uint64), ("bias_permille", C.c_uint32),
                ("bias", C.c_double), ("threshold", C.c_double), ("reset", C.c_double),
                ("leak_decay", C.c_double), ("w_min", C.c_int32), ("w_max", C.c_int32),
                ("max_delay", C.c_uint32), ("log_spikes", C.c_uint32), ("log_potential_n", C.c_uint32)]


class TileDesc(C.Structure):
    _fields_ = [("x", C.c_uint32), ("y", C.c_uint32),
                ("energy_east", C.c_double), ("energy_west", C.c_double), ("energy_south", C.c_double),
                ("energy_north", C.c_double), ("latency_east", C.c_double), ("latency_west", C.c_double),
                ("latency_south", C.c_double), ("latency_north", C.c_double)]


class CoreDesc(C.Structure):
    _fields_ = [("id", C.c_uint32), ("tile", C.c_uint32), ("offset", C.c_uint32), ("buffer_pos", C.c_uint32),
                ("neuron_begin", C.c_uint32), ("neuron_count", C.c_uint32),
                ("axon_in_begin", C.c_uint32), ("axon_in_count", C.c_uint32),
                ("syn_begin", C.c_uint64), ("syn_count", C.c_uint64),
                ("acc_mode", C.c_uint32), ("weight_shift", C.c_int32), ("ring", C.c_uint32),
                ("dend_in_msg", C.c_uint32), ("fixed_slots", C.c_uint32), ("pad", C.c_uint32),
                ("energy_axon_in", C.c_double), ("latency_axon_in", C.c_double),
                ("energy_axon_out", C.c_double), ("latency_axon_out", C.c_double)]


class SomaClass(C.Structure):
    _fields_ = [("model", C.c_uint32), ("reset_mode", C.c_uint32), ("reverse_reset_mode", C.c_uint32),
                ("refractory_delay", C.c_int32), ("flags", C.c_uint32), ("random_mask", C.c_uint32),
                ("dend_model", C.c_uint32), ("dend_in_neuron", C.c_uint32),
                ("threshold", C.c_double), ("reverse_threshold", C.c_double), ("reset", C.c_double),
                ("reverse_reset", C.c_double), ("leak", C.c_double), ("input_decay", C.c_double),
                ("energy_access", C.c_double), ("energy_update", C.c_double), ("energy_spike_out", C.c_double),
                ("latency_access", C.c_double), ("latency_update", C.c_double), ("latency_spike_out", C.c_double),
                ("dend_energy_update", C.c_double), ("dend_latency_update", C.c_double),
                ("nf_kp", C.c_double), ("nf_ki", C.c_double), ("nf_dt", C.c_double)]


class CostClass(C.Structure):
    _fields_ = [("syn_energy", C.c_double), ("syn_latency", C.c_double), ("den_energy", C.c_double),
                ("den_latency", C.c_double), ("per_message", C.c_uint32), ("den_is_soma", C.c_uint32)]


class AxonIn(C.Structure):
    _fields_ = [("syn_off", C.c_uint32), ("syn_count", C.c_uint32), ("hop", C.c_uint32), ("cost_class", C.c_uint32)]


class InputDesc(C.Structure):
    _fields_ = [("spikes_off", C.c_uint32), ("spikes_len", C.c_uint32), ("share_count", C.c_uint32),
                ("share_rank", C.c_uint32), ("rate", C.c_double), ("poisson", C.c_double),
                ("unit", C.c_uint32), ("poisson_col", C.c_uint32)]


class NoiseDesc(C.Structure):
    _fields_ = [("off", C.c_uint32), ("len", C.c_uint32), ("share_count", C.c_uint32), ("share_rank", C.c_uint32)]


class TapsDesc(C.Structure):
    _fields_ = [("n_taps", C.c_uint32), ("const_off", C.c_uint32)]


class HHInit(C.Structure):
    _fields_ = [("m", C.c_double), ("n", C.c_double), ("h", C.c_double), ("current", C.c_double)]


class Tables(C.Structure):
    """sfe_tables (include/sanafe_b200.h)."""
    _fields_ = [("abi_version", C.c_uint32), ("noc_width", C.c_uint32), ("noc_height", C.c_uint32),
                ("noc_buffer_size", C.c_uint32), ("max_cores_per_tile", C.c_uint32),
                ("n_tiles", C.c_uint32), ("n_cores", C.c_uint32), ("n_neurons", C.c_uint32),
                ("n_axons_in", C.c_uint32), ("n_soma_classes", C.c_uint32), ("n_cost_classes", C.c_uint32),
                ("n_inputs", C.c_uint32), ("n_hh", C.c_uint32), ("n_probes", C.c_uint32),
                ("mapped_tiles", C.c_uint32), ("mapped_cores", C.c_uint32),
                ("n_synapses", C.c_uint64), ("n_axons_out", C.c_uint64), ("sync_delay", C.c_double),
                ("tiles", C.POINTER(TileDesc)), ("cores", C.POINTER(CoreDesc)),
                ("soma_classes", C.POINTER(SomaClass)), ("cost_classes", C.POINTER(CostClass)),
                ("neuron_class", C.POINTER(C.c_uint32)), ("neuron_aux", C.POINTER(C.c_uint32)),
                ("neuron_bias", C.POINTER(C.c_double)), ("neuron_potential0", C.POINTER(C.c_double)),
                ("axon_out_begin", C.POINTER(C.c_uint32)), ("axon_out_target", C.POINTER(C.c_uint32)),
                ("inputs", C.POINTER(InputDesc)), ("input_spikes", C.POINTER(C.c_uint8)),
                ("n_input_spikes", C.c_uint64), ("hh", C.POINTER(HHInit)), ("probes", C.POINTER(C.c_uint32)),
                ("axons_in", C.POINTER(AxonIn)), ("axon_src", C.POINTER(C.c_uint32)),
                ("syn_weight", C.POINTER(C.c_double)), ("syn_meta", C.POINTER(C.c_uint32)),
                ("synth", C.POINTER(SynthSpec)), ("input_seed_base", C.c_uint32), ("n_poisson_cols", C.c_uint32),
                ("noise", C.POINTER(NoiseDesc)), ("noise_values", C.POINTER(C.c_double)),
                ("n_noise_values", C.c_uint64), ("n_noise", C.c_uint32), ("n_u_probes", C.c_uint32),
                ("u_probes", C.POINTER(C.c_uint32)), ("neuron_taps", C.POINTER(C.c_uint32)),
                ("taps", C.POINTER(TapsDesc)), ("taps_values", C.POINTER(C.c_double)),
                ("n_taps_units", C.c_uint32), ("n_taps_values", C.c_uint32),
                ("device_models", C.c_void_p), ("n_device_models", C.c_uint32), ("n_device_instances", C.c_uint32),
                ("n_rand_cols", C.c_uint32), ("pad_rand", C.c_uint32)]


class StepRecord(C.Structure):
    _fields_ = [("neurons_fired", C.c_int64), ("neurons_updated", C.c_int64), ("packets_sent", C.c_int64),
                ("total_hops", C.c_int64), ("spike_count", C.c_int64), ("sim_time", C.c_double),
                ("synapse_energy", C.c_double), ("dendrite_energy", C.c_double), ("soma_energy", C.c_double),
                ("network_energy", C.c_double), ("total_energy", C.c_double)]


STEP_DTYPE = np.dtype([("neurons_fired", "<i8"), ("neurons_updated", "<i8"), ("packets_sent", "<i8"),
                       ("total_hops", "<i8"), ("spike_count", "<i8"), ("sim_time", "<f8"),
                       ("synapse_energy", "<f8"), ("dendrite_energy", "<f8"), ("soma_energy", "<f8"),
                       ("network_energy", "<f8"), ("total_energy", "<f8")])


class RunData(C.Structure):
    _fields_ = [("timestep_start", C.c_int64), ("timesteps_executed", C.c_int64), ("spikes", C.c_int64),
                ("packets_sent", C.c_int64), ("neurons_updated", C.c_int64), ("neurons_fired", C.c_int64),
                ("total_energy", C.c_double), ("synapse_energy", C.c_double), ("dendrite_energy", C.c_double),
                ("soma_energy", C.c_double), ("network_energy", C.c_double), ("sim_time", C.c_double),
                ("wall_time", C.c_double), ("scheduler_wall_time", C.c_double)]


class TraceRequest(C.Structure):
    _fields_ = [("steps", C.c_void_p), ("fired_bits", C.c_void_p), ("potentials", C.c_void_p),
                ("status", C.c_void_p), ("neuron_traces", C.c_void_p)]


_lib = None


def lib():
    """The loaded C ABI. Raises if the in-tree library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise ImportError(f"{_LIB_PATH} missing: build it with `make -C sana-fe_b200` "
                          "(or __graft_entry__.build()); there is no Python/CPU fallback")
    L = C.CDLL(_LIB_PATH)
    vp, i64, u32, u64, dbl, sz, cstr = C.c_void_p, C.c_int64, C.c_uint32, C.c_uint64, C.c_double, C.c_size_t, C.c_char_p
    sigs = {
        "sfe_last_error": (cstr, []), "sfe_abi_version": (C.c_int, []), "sfe_device_count": (C.c_int, []),
        "sfe_engine_create": (vp, [C.POINTER(Tables), C.c_int]), "sfe_engine_destroy": (None, [vp]),
        "sfe_engine_set_stream": (C.c_int, [vp, vp]),
        "sfe_engine_run": (C.c_int, [vp, i64, C.POINTER(TraceRequest), C.POINTER(RunData)]),
        "sfe_engine_enqueue": (C.c_int, [vp, i64]), "sfe_engine_collect": (C.c_int, [vp, C.POINTER(RunData)]),
        "sfe_engine_reset": (C.c_int, [vp]), "sfe_engine_set_bias": (C.c_int, [vp, vp, sz]),
        "sfe_engine_set_bias_staged": (C.c_int, [vp, vp, sz]),
        "sfe_engine_set_neuron_bias": (C.c_int, [vp, u32, dbl]),
        "sfe_engine_read_potentials": (C.c_int, [vp, vp, sz]), "sfe_engine_read_fired": (C.c_int, [vp, vp, sz]),
        "sfe_engine_read_raster": (C.c_int, [vp, vp, sz]),
        "sfe_engine_total_timesteps": (i64, [vp]), "sfe_engine_launch_count": (i64, [vp]),
        "sfe_engine_time_launches": (C.c_int, [vp, C.c_int]),
        "sfe_engine_time_begin": (C.c_int, [vp]),
        "sfe_engine_time_end": (C.c_int, [vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
        "sfe_engine_create_partitioned": (vp, [C.POINTER(Tables), C.c_int, u32, u32]),
        "sfe_engine_set_exchange_buffers": (C.c_int, [vp, vp, vp]),
        "sfe_engine_collect_records": (i64, [vp, vp, i64]),
        "sfe_engine_partition_info": (C.c_int, [vp, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32), C.POINTER(u32), C.POINTER(u64)]),
        "sfe_chip_set_partition": (C.c_int, [vp, u32, u32]),
        "sfe_chip_set_input_seed_base": (C.c_int, [vp, u32]),
        "sfe_engine_set_input_overlay": (C.c_int, [vp, vp, i64, u32]),
        "sfe_engine_fill_input_overlay": (C.c_int, [vp, i64]),
        "sfe_poisson_create": (vp, [C.POINTER(Tables)]), "sfe_poisson_destroy": (None, [vp]),
        "sfe_poisson_cols": (u32, [vp]), "sfe_poisson_fill": (C.c_int, [vp, vp, i64]),
        "sfe_poisson_reference_draws": (None, [u32, vp, sz]), "sfe_mt19937_draws": (None, [u32, vp, sz, sz]),
        "sfe_glibc_rand_draws": (None, [u32, u64, vp, sz]), "sfe_engine_set_rand_overlay": (C.c_int, [vp, vp, i64, u32]),
        "sfe_engine_read_log_tail": (i64, [vp, vp, i64]),
        "sfe_nccl_get_unique_id": (C.c_int, [vp, cstr]),
        "sfe_engine_comm_init": (C.c_int, [vp, vp, cstr]),
        "sfe_engine_comm_destroy": (C.c_int, [vp]),
        "sfe_engine_enqueue_partitioned": (C.c_int, [vp, i64]),
        "sfe_engine_raster_layout": (C.c_int, [vp, vp, sz]),
        "sfe_engine_p2p_export": (C.c_int, [vp, vp]),
        "sfe_engine_p2p_attach": (C.c_int, [vp, vp]),
        "sfe_engine_p2p_detach": (C.c_int, [vp]),
        "sfe_engine_exchange_error": (C.c_int, [vp]),
        "sfe_engine_read_timeline": (C.c_int, [vp, vp, sz]),
        "sfe_engine_read_timeline_raw": (i64, [vp, vp, sz]),
        "sfe_plan_partition": (C.c_int, [vp, C.c_uint32, vp, vp, vp]),
        "sfe_engine_synchronize": (C.c_int, [vp]),
        "sfe_device_memcpy": (C.c_int, [vp, vp, sz]),
        "sfe_host_alloc": (vp, [sz]), "sfe_host_free": (None, [vp]),
        "sfe_engine_enqueue_neuron_phase": (C.c_int, [vp]), "sfe_engine_enqueue_message_phase": (C.c_int, [vp]),
        "sfe_engine_fired_local_ptr": (vp, [vp, C.POINTER(sz)]), "sfe_engine_fired_global_ptr": (vp, [vp, C.POINTER(sz)]),
        "sfe_engine_device_bytes": (sz, [vp]),
        "sfe_arch_load_yaml": (vp, [cstr]), "sfe_net_load_yaml": (vp, [cstr, vp]),
        "sfe_net_load_netlist": (vp, [cstr, vp]),
        "sfe_load_flat": (C.c_int, [cstr, C.POINTER(vp), C.POINTER(vp)]),
        "sfe_arch_free": (None, [vp]), "sfe_net_free": (None, [vp]),
        "sfe_chip_create": (vp, [vp, C.c_int]), "sfe_chip_destroy": (None, [vp]),
        "sfe_chip_load": (C.c_int, [vp, vp]),
        "sfe_chip_load_synthetic": (C.c_int, [vp, C.POINTER(SynthSpec), C.c_int]),
        "sfe_chip_sim": (C.c_int, [vp, i64, C.c_int, C.POINTER(TraceRequest), C.POINTER(RunData)]),
        "sfe_chip_schedule_detailed": (C.c_int, [vp, vp, i64, vp]),
        "sfe_chip_set_scheduler_threads": (C.c_int, [vp, u32]),
        "sfe_chip_reset": (C.c_int, [vp]), "sfe_chip_get_power": (dbl, [vp]),
        "sfe_chip_tables": (C.POINTER(Tables), [vp]), "sfe_chip_engine": (vp, [vp]),
        "sfe_chip_neuron_index": (i64, [vp, cstr, u64]),
        "sfe_chip_set_neuron_attribute": (C.c_int, [vp, cstr, u64, cstr, dbl]),
        "sfe_chip_format_messages": (sz, [vp, vp, i64, i64, C.c_int, C.c_char_p, sz]),
        "sfe_chip_format_spikes": (sz, [vp, vp, i64, i64, C.c_char_p, sz]),
        "sfe_chip_probe_names": (sz, [vp, C.c_char_p, sz]),
        "sfe_chip_trace_names": (sz, [vp, C.c_char_p, sz]),
        "sfe_chip_group_names": (sz, [vp, C.c_char_p, sz]),
        "sfe_chip_set_neuron_log_spikes": (C.c_int, [vp, cstr, u64, C.c_int]),
        "sfe_chip_request_stop": (None, [vp]), "sfe_engine_request_stop": (None, [vp, C.c_int]),
        "sfe_last_error_kind": (C.c_int, []),
        "sfe_net_create": (vp, [cstr]), "sfe_net_save_yaml": (C.c_int, [vp, cstr]), "sfe_net_save_netlist": (C.c_int, [vp, cstr]),
        "sfe_batch_load": (C.c_int, [vp, vp, u32, u32]),
        "sfe_batch_sim": (C.c_int, [vp, u32, i64, C.c_int, vp, vp, u32]),
        # out-of-tree device models (include/sfe_device_model.h)
        "sfe_register_device_model": (C.c_int, [cstr, vp]), "sfe_unregister_device_model": (C.c_int, [cstr]),
        "sfe_load_device_model": (C.c_int, [cstr, cstr]), "sfe_device_model_registered": (C.c_int, [cstr]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(L, name)  # AttributeError if the ABI header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


ABI_SYMBOLS = None  # filled lazily by tests from include/sanafe_b200.h


class SanafeError(RuntimeError):
    pass


def _check(rc):
    if rc != 0:
        raise SanafeError(lib().sfe_last_error().decode())


def _ptr(h):
    if not h:
        raise SanafeError(lib().sfe_last_error().decode())
    return h


TIMING = {"simple": 0, "detailed": 1, "cycle": 2}


class Architecture:
    def __init__(self, handle):
        self._h = handle

    def __del__(self):
        if getattr(self, "_h", None):
            try:
                lib().sfe_arch_free(self._h)
            except Exception:  # noqa: BLE001 - interpreter shutdown: the module may already be torn down
                pass
            self._h = None


class Network:
    def __init__(self, handle):
        self._h = handle

    def save(self, path, use_netlist_format=False):
        """sanafe.Network.save (src/network.cpp:705-718): YAML, or the legacy netlist format (lossy, as the reference's)."""
        fn = lib().sfe_net_save_netlist if use_netlist_format else lib().sfe_net_save_yaml
        _check(fn(self._h, os.fsencode(path)))

    def __del__(self):
        if getattr(self, "_h", None):
            try:
                lib().sfe_net_free(self._h)
            except Exception:  # noqa: BLE001
                pass
            self._h = None


def load_arch(path):
    """sanafe.load_arch (src/arch.cpp:106-117)."""
    return Architecture(_ptr(lib().sfe_arch_load_yaml(os.fsencode(path))))


def load_net(path, arch, use_netlist_format=False):
    """sanafe.load_net (src/network.cpp:194-222)."""
    if use_netlist_format:
        return Network(_ptr(lib().sfe_net_load_netlist(os.fsencode(path), arch._h)))
    return Network(_ptr(lib().sfe_net_load_yaml(os.fsencode(path), arch._h)))


def load_flat(path):
    """(Architecture, Network) from a flat JSON-lines description."""
    a, n = C.c_void_p(), C.c_void_p()
    _check(lib().sfe_load_flat(os.fsencode(path), C.byref(a), C.byref(n)))
    return Architecture(a.value), Network(n.value)


class SpikingChip:
    """sanafe.SpikingChip (src/pymodule.cpp:1174-1212; src/chip.hpp:56-107)."""

    def __init__(self, arch, device=0):
        self._arch = arch
        self._h = _ptr(lib().sfe_chip_create(arch._h, device))
        self._timestep = 0

    def __del__(self):
        if getattr(self, "_h", None):
            try:
                lib().sfe_chip_destroy(self._h)
            except Exception:  # noqa: BLE001
                pass
            self._h = None

    def set_partition(self, rank, world):
        """Multi-GPU: this chip simulates core range `rank` of `world` (call before load)."""
        _check(lib().sfe_chip_set_partition(self._h, rank, world))

    def set_input_seed_base(self, base=0):
        """Poisson inputs: seed the generators as if `base` "input" units had been created in this
        process before this chip's (the reference counts them process-wide). Call before load."""
        _check(lib().sfe_chip_set_input_seed_base(self._h, base))

    def load(self, net, overwrite=False):
        """SpikingChip.load (src/chip.cpp:129-138). overwrite=False on a loaded chip would map a second network
        next to the first in the reference; this engine lowers one description per load and refuses that."""
        if getattr(self, "_loaded", False) and not overwrite:
            raise SanafeError("SpikingChip.load: mapping a second network next to the loaded one is not supported by "
                              "the B200 engine; pass overwrite=True to replace it")
        _check(lib().sfe_chip_load(self._h, net._h))
        self._loaded = True

    def load_synthetic(self, spec, generate_on_device=True):
        _check(lib().sfe_chip_load_synthetic(self._h, C.byref(spec), 1 if generate_on_device else 0))

    @property
    def tables(self):
        t = lib().sfe_chip_tables(self._h)
        if not t:
            raise SanafeError("no network loaded")
        return t.contents

    @property
    def engine(self):
        return lib().sfe_chip_engine(self._h)

    def neuron_index(self, group, offset):
        return lib().sfe_chip_neuron_index(self._h, str(group).encode(), offset)

    def set_neuron_attribute(self, group, offset, name, value):
        _check(lib().sfe_chip_set_neuron_attribute(self._h, str(group).encode(), offset, name.encode(), float(value)))

    def reset(self):
        _check(lib().sfe_chip_reset(self._h))

    def get_power(self):
        return lib().sfe_chip_get_power(self._h)

    def probe_names(self):
        n = lib().sfe_chip_probe_names(self._h, None, 0)
        buf = C.create_string_buffer(n + 1)
        lib().sfe_chip_probe_names(self._h, buf, n + 1)
        return buf.value.decode().split("\n")[:-1] if n else []

    def trace_names(self):
        """Columns of the model-defined neuron traces ("group.offset/trace"), in trace order."""
        n = lib().sfe_chip_trace_names(self._h, None, 0)
        buf = C.create_string_buffer(n + 1)
        lib().sfe_chip_trace_names(self._h, buf, n + 1)
        return buf.value.decode().split("\n")[:-1] if n else []

    def sim_raw(self, timesteps, timing_model="simple", steps=False, fired=False, potentials=False, status=False,
                neuron_traces=False):
        """One sfe_chip_sim call; returns (RunData, dict of numpy traces)."""
        t = self.tables
        req = TraceRequest()
        out = {}
        if steps:
            out["steps"] = np.zeros(timesteps, dtype=STEP_DTYPE)
            req.steps = out["steps"].ctypes.data
        if fired:
            out["fired_bits"] = np.zeros((timesteps, (t.n_neurons + 31) // 32), dtype=np.uint32)
            req.fired_bits = out["fired_bits"].ctypes.data
        if potentials and t.n_probes:
            out["potentials"] = np.zeros((timesteps, t.n_probes), dtype=np.float64)
            req.potentials = out["potentials"].ctypes.data
        if status:
            out["status"] = np.zeros((timesteps, t.n_neurons), dtype=np.uint8)
            req.status = out["status"].ctypes.data
        if neuron_traces and t.n_u_probes:
            out["neuron_traces"] = np.zeros((timesteps, t.n_u_probes), dtype=np.float64)
            req.neuron_traces = out["neuron_traces"].ctypes.data
        rd = RunData()
        _check(lib().sfe_chip_sim(self._h, timesteps, TIMING[timing_model], C.byref(req), C.byref(rd)))
        return rd, out

    def format_spikes(self, fired_bits, timestep_start):
        fired_bits = np.ascontiguousarray(fired_bits, dtype=np.uint32)
        steps = fired_bits.shape[0]
        n = lib().sfe_chip_format_spikes(self._h, fired_bits.ctypes.data, steps, timestep_start, None, 0)
        buf = C.create_string_buffer(n + 1)
        lib().sfe_chip_format_spikes(self._h, fired_bits.ctypes.data, steps, timestep_start, buf, n + 1)
        return buf.value.decode()

    MESSAGE_HEADER = ("timestep,mid,src_neuron,src_hw,dest_hw,hops,spikes,send_timestamp,received_timestamp,"
                      "processed_timestamp,generation_delay,processing_delay,network_delay,blocking_delay,"
                      "min_hop_delay,messages_along_route\n")

    def format_messages(self, status, timestep_start, timing_model="simple"):
        """Rows of messages.csv for the steps whose status bytes are given (src/chip.cpp:1731-1764)."""
        status = np.ascontiguousarray(status, dtype=np.uint8)
        steps = status.shape[0]
        tm = TIMING[timing_model] if isinstance(timing_model, str) else int(timing_model)
        n = lib().sfe_chip_format_messages(self._h, status.ctypes.data, steps, timestep_start, tm, None, 0)
        buf = C.create_string_buffer(n + 1)
        lib().sfe_chip_format_messages(self._h, status.ctypes.data, steps, timestep_start, tm, buf, n + 1)
        return buf.value.decode()

    def sim(self, timesteps=1, timing_model="detailed", processing_threads=0, scheduler_threads=0,
            spike_trace=None, potential_trace=None, neuron_trace=None, perf_trace=None, message_trace=None,
            write_trace_headers=True):
        """sanafe.SpikingChip.sim (src/pymodule.cpp:549-706): returns the same result dict."""
        rd, tr = self.sim_raw(timesteps, timing_model, steps=bool(perf_trace), fired=bool(spike_trace),
                              potentials=bool(potential_trace), status=bool(message_trace),
                              neuron_traces=bool(neuron_trace))
        result = {
            "timestep_start": rd.timestep_start, "timesteps_executed": rd.timesteps_executed,
            "energy": {"total": rd.total_energy, "synapse": rd.synapse_energy, "dendrite": rd.dendrite_energy,
                       "soma": rd.soma_energy, "network": rd.network_energy},
            "sim_time": rd.sim_time, "spikes": rd.spikes, "packets_sent": rd.packets_sent,
            "neurons_updated": rd.neurons_updated, "neurons_fired": rd.neurons_fired,
        }
        if spike_trace:
            text = self.format_spikes(tr["fired_bits"], rd.timestep_start)
            if spike_trace is True:
                rows = [r.split(",") for r in text.split("\n") if r]
                per_step = [[] for _ in range(timesteps)]
                for addr, ts in rows:
                    g, o = addr.rsplit(".", 1)
                    per_step[int(ts) - rd.timestep_start].append((g, int(o)))
                result["spike_trace"] = per_step
            else:
                self._write_text(spike_trace, ("neuron,timestep\n" if write_trace_headers else "") + text)
        if potential_trace:
            pots = tr.get("potentials", np.zeros((timesteps, 0)))
            if potential_trace is True:
                result["potential_trace"] = pots.tolist()
            else:
                names = self.probe_names()
                hdr = "timestep," + "".join(f"neuron {n}," for n in names) + "\n"
                body = "".join(f"{rd.timestep_start + s}," + "".join(f"{v:g}," for v in pots[s]) + "\n"
                               for s in range(timesteps)) if names else ""
                self._write_text(potential_trace, (hdr if write_trace_headers else "") + body)
        if neuron_trace:
            # src/pymodule.cpp:664-668 / src/chip.cpp:1478-1517, 1664-1702
            vals = tr.get("neuron_traces", np.zeros((timesteps, 0)))
            if neuron_trace is True:
                result["neuron_trace"] = {"u": vals.tolist()} if vals.shape[1] else {}  # src/pytrace.hpp:205-212
            else:
                names = self.trace_names()
                hdr = "timestep," + "".join(f"neuron {n}," for n in names) + "\n"
                body = "".join(f"{rd.timestep_start + s}," + "".join(f"{v:g}," for v in vals[s]) + ("\n" if names else "")
                               for s in range(timesteps))
                self._write_text(neuron_trace, (hdr if write_trace_headers else "") + body)
        if message_trace:
            text = self.format_messages(tr["status"], rd.timestep_start, timing_model)
            if message_trace is True:
                result["message_trace"] = [r.split(",") for r in text.split("\n") if r]
            else:
                self._write_text(message_trace, (self.MESSAGE_HEADER if write_trace_headers else "") + text)
        if perf_trace:
            st = tr["steps"]
            if perf_trace is True:
                result["perf_trace"] = {k: st[k].tolist() for k in st.dtype.names}
            else:
                hdr = ("timestep,fired,updated,packets,hops,spikes,sim_time,synapse_energy,dendrite_energy,"
                       "soma_energy,network_energy,total_energy\n")
                body = "".join(
                    f"{rd.timestep_start + s},{r['neurons_fired']},{r['neurons_updated']},{r['packets_sent']},"
                    f"{r['total_hops']},{r['spike_count']},{r['sim_time']:e},{r['synapse_energy']:e},"
                    f"{r['dendrite_energy']:e},{r['soma_energy']:e},{r['network_energy']:e},{r['total_energy']:e}\n"
                    for s, r in enumerate(st))
                self._write_text(perf_trace, (hdr if write_trace_headers else "") + body)
        return result

    @staticmethod
    def _write_text(sink, text):
        if isinstance(sink, (str, os.PathLike)):
            with open(sink, "a") as f:
                f.write(text)
        else:
            sink.write(text)


def merge_partition_records(per_rank):
    """Combine the per-step records of all partitions of one chip: counts and energies
    are partial sums over each rank's cores (summed in rank order), sim_time is the
    local maximum (simple timing model: max over cores, src/schedule.cpp:89-98)."""
    out = per_rank[0].copy()
    for rec in per_rank[1:]:
        for name in out.dtype.names:
            if name == "sim_time":
                out[name] = np.maximum(out[name], rec[name])
            else:
                out[name] = out[name] + rec[name]
    return out


def run_data_from_records(records, timestep_start=1):
    """RunData totals from per-step records, accumulated step by step (src/chip.cpp:462-475)."""
    rd = RunData()
    rd.timestep_start = timestep_start
    rd.timesteps_executed = len(records)
    for r in records:
        rd.total_energy += float(r["total_energy"])
        rd.synapse_energy += float(r["synapse_energy"])
        rd.dendrite_energy += float(r["dendrite_energy"])
        rd.soma_energy += float(r["soma_energy"])
        rd.network_energy += float(r["network_energy"])
        rd.sim_time += float(r["sim_time"])
        rd.spikes += int(r["spike_count"])
        rd.packets_sent += int(r["packets_sent"])
        rd.neurons_updated += int(r["neurons_updated"])
        rd.neurons_fired += int(r["neurons_fired"])
    return rd


class PartitionedChip:
    """One simulated chip split over `world` engines (one per GPU; all on one device
    when `devices` repeats an index, which is how the single-GPU tests emulate the
    multi-GPU path). `exchange` moves every rank's raster slice into every rank's
    global raster between the two phases of a step: device copies here, an NCCL
    all-gather in bench.py."""

    def __init__(self, make_chip, world, devices=None):
        self.world = world
        self.chips = []
        for r in range(world):
            chip = make_chip(devices[r] if devices else 0, r, world)
            self.chips.append(chip)
        L = lib()
        self.local, self.glob = [], []
        for c in self.chips:
            nb = C.c_size_t()
            self.local.append(L.sfe_engine_fired_local_ptr(c.engine, C.byref(nb)))
            self.slice_bytes = nb.value
            self.glob.append(L.sfe_engine_fired_global_ptr(c.engine, C.byref(nb)))
            self.global_bytes = nb.value

    def exchange(self):
        L = lib()
        for c in self.chips:
            _check(L.sfe_engine_synchronize(c.engine))
        for dst in range(self.world):
            for src in range(self.world):
                _check(L.sfe_device_memcpy(self.glob[dst] + src * self.slice_bytes, self.local[src], self.slice_bytes))

    def step(self):
        L = lib()
        for c in self.chips:
            _check(L.sfe_engine_enqueue_neuron_phase(c.engine))
        self.exchange()
        for c in self.chips:
            _check(L.sfe_engine_enqueue_message_phase(c.engine))

    def raster(self):
        """Global fired raster of the last step in device-index bit order."""
        L = lib()
        c0 = self.chips[0]
        t = c0.tables
        _check(L.sfe_engine_synchronize(c0.engine))
        padded = np.zeros(self.global_bytes // 4, dtype=np.uint32)
        _check(L.sfe_device_memcpy(padded.ctypes.data, self.glob[0], self.global_bytes))
        begin = np.zeros(t.n_cores, dtype=np.uint32)
        _check(L.sfe_engine_raster_layout(c0.engine, begin.ctypes.data, t.n_cores))
        bits = np.zeros(t.n_neurons, dtype=bool)
        for c in range(t.n_cores):
            cd = t.cores[c]
            if cd.neuron_count == 0:
                continue
            words = padded[begin[c]:begin[c] + (cd.neuron_count + 31) // 32]
            unpacked = np.unpackbits(words.view(np.uint8), bitorder="little")[:cd.neuron_count]
            bits[cd.neuron_begin:cd.neuron_begin + cd.neuron_count] = unpacked.astype(bool)
        return np.packbits(bits, bitorder="little").view(np.uint8)

    def collect(self):
        """Merged per-step records of the steps run since the last collect."""
        L = lib()
        per_rank = []
        for c in self.chips:
            n = L.sfe_engine_total_timesteps(c.engine)
            buf = np.zeros(max(n, 1), dtype=STEP_DTYPE)
            got = L.sfe_engine_collect_records(c.engine, buf.ctypes.data, len(buf))
            if got < 0:
                raise SanafeError(L.sfe_last_error().decode())
            per_rank.append(buf[:got])
        return merge_partition_records(per_rank)
