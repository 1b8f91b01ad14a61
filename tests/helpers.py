"""Shared test helpers: the oracle (CPU restatement + reference binary) and fixtures.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may touch oracle/.
"""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sana-fe_b200"))
import sanafe_b200 as sfe  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "sanafe_ref")
REF_HH_PLUGIN = os.path.join(ORACLE_DIR, "_ref", "libhodgkin_huxley.so")
REFERENCE_ROOT = "/root/reference"

_oracle = None


def oracle_lib():
    """ctypes handle of oracle/libsfe_oracle.so (the CPU restatement)."""
    global _oracle
    if _oracle is None:
        path = os.path.join(ORACLE_DIR, "libsfe_oracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "port"])
        L = C.CDLL(path)
        L.sfe_oracle_create.restype = C.c_void_p
        L.sfe_oracle_create.argtypes = [C.POINTER(sfe.Tables)]
        L.sfe_oracle_destroy.argtypes = [C.c_void_p]
        L.sfe_oracle_run.restype = C.c_int
        L.sfe_oracle_run.argtypes = [C.c_void_p, C.c_int64, C.POINTER(sfe.TraceRequest), C.POINTER(sfe.RunData)]
        L.sfe_oracle_reset.argtypes = [C.c_void_p]
        L.sfe_oracle_set_bias.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.sfe_oracle_get_power.restype = C.c_double
        L.sfe_oracle_get_power.argtypes = [C.c_void_p]
        L.sfe_oracle_read_potentials.argtypes = [C.c_void_p, C.c_void_p]
        L.sfe_oracle_set_input_overlay.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_uint32]
        _oracle = L
    return _oracle


class Oracle:
    """CPU restatement driven over a chip's lowered tables."""

    def __init__(self, chip):
        self.chip = chip  # keeps the tables alive
        self.t = chip.tables
        self.h = oracle_lib().sfe_oracle_create(C.byref(self.t))
        if not self.h:
            raise RuntimeError("sfe_oracle_create failed (tables without host synapse arrays?)")

    def __del__(self):
        if getattr(self, "h", None):
            oracle_lib().sfe_oracle_destroy(self.h)
            self.h = None

    def run(self, timesteps, steps=True, fired=True, potentials=True, status=False):
        t = self.t
        req = sfe.TraceRequest()
        out = {}
        if steps:
            out["steps"] = np.zeros(timesteps, dtype=sfe.STEP_DTYPE)
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
        if t.n_u_probes:
            out["neuron_traces"] = np.zeros((timesteps, t.n_u_probes), dtype=np.float64)
            req.neuron_traces = out["neuron_traces"].ctypes.data
        rd = sfe.RunData()
        rc = oracle_lib().sfe_oracle_run(self.h, timesteps, C.byref(req), C.byref(rd))
        assert rc == 0
        return rd, out

    def reset(self):
        oracle_lib().sfe_oracle_reset(self.h)

    def set_input_overlay(self, bits):
        """Poisson spikes drawn elsewhere (the product's sfe_poisson_fill) for the next len(bits) steps,
        instead of the restatement's own MT19937."""
        self._overlay = np.ascontiguousarray(bits, dtype=np.uint8)  # the oracle keeps the pointer
        oracle_lib().sfe_oracle_set_input_overlay(self.h, self._overlay.ctypes.data, self._overlay.shape[0],
                                                  self._overlay.shape[1] if self._overlay.ndim == 2 else 0)

    def set_bias(self, bias):
        bias = np.ascontiguousarray(bias, dtype=np.float64)
        oracle_lib().sfe_oracle_set_bias(self.h, bias.ctypes.data, bias.size)


def have_reference_binary():
    return os.path.exists(REF_BIN)


def run_reference(flat_path, out_dir, steps, timing="simple", threads=1, per_step=True, traces=False, dump_map=False):
    """Run the reference engine itself (oracle/_ref/sanafe_ref) on a flat description."""
    cmd = [REF_BIN, flat_path, "--steps", str(steps), "--timing", timing, "--threads", str(threads), "--out", out_dir]
    if per_step:
        cmd.append("--per-step")
    if traces:
        cmd.append("--traces")
    if dump_map:
        cmd.append("--dump-map")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"sanafe_ref failed: {res.stderr[-2000:]}")
    with open(os.path.join(out_dir, "summary.json")) as f:
        return json.load(f)


def load_ref_steps(out_dir):
    """steps.csv of a --per-step reference run as a structured array (full precision)."""
    return np.genfromtxt(os.path.join(out_dir, "steps.csv"), delimiter=",", names=True, dtype=None, encoding=None)


def load_ref_potentials(out_dir):
    path = os.path.join(out_dir, "potentials_full.csv")
    if not os.path.exists(path) or os.path.getsize(path) == 0:
        return None
    arr = np.loadtxt(path, delimiter=",", ndmin=2)
    return arr[:, 1:]


def load_ref_traces(out_dir):
    """traces_full.csv of a --per-step reference run (rows "t,name=value,..." from SpikingChip::get_traces):
    the values per step, in file order, or None when the run has no model-defined traces."""
    path = os.path.join(out_dir, "traces_full.csv")
    if not os.path.exists(path) or os.path.getsize(path) == 0:
        return None
    rows = []
    with open(path) as f:
        for line in f:
            rows.append([float(x.split("=", 1)[1]) for x in line.strip().split(",")[1:]])
    return np.asarray(rows, dtype=np.float64)


def load_ref_spikes(out_dir):
    with open(os.path.join(out_dir, "spikes_full.csv")) as f:
        return f.read()


def rundata_dict(rd):
    return {name: getattr(rd, name) for name, _ in rd._fields_}


# ---------------------------------------------------------------------------
# golden fixtures (tests/golden/, generated by make_goldens.py from the reference)
# ---------------------------------------------------------------------------
import gzip
import hashlib
import tempfile

GOLDEN_CASES = ["example", "dvs", "hh", "synth_small", "synth_delay", "synth_quirk", "synth_soma", "truenorth", "frac"]
# Cases added late in round 1 (Poisson inputs, LIF file noise + model-defined traces): pinned on the CPU against
# the reference here; their device tests live in tests/test_zz_new_models_gpu.py (collected last; green on a
# B200, profiles/r1_pytest_new_models_gpu.log).
NEW_GOLDEN_CASES = ["poisson", "noise", "taps", "neurofem", "truenorth_rand"]
ORACLE_ONLY_CASES = []
_flat_cache = {}


def golden_flat(name):
    """Path of the flat description of a golden case (gunzipped to a temp file if needed)."""
    plain = os.path.join(GOLDEN, name + ".jsonl")
    if os.path.exists(plain):
        return plain
    if name not in _flat_cache:
        tmp = tempfile.NamedTemporaryFile(prefix=name + "_", suffix=".jsonl", delete=False)
        with gzip.open(plain + ".gz", "rb") as f:
            tmp.write(f.read())
        tmp.close()
        _flat_cache[name] = tmp.name
    return _flat_cache[name]


def golden(name):
    with open(os.path.join(GOLDEN, name + ".golden.json")) as f:
        return json.load(f)


def golden_spikes(name):
    plain = os.path.join(GOLDEN, name + ".spikes.txt")
    if os.path.exists(plain):
        with open(plain) as f:
            return f.read()
    with gzip.open(plain + ".gz", "rt") as f:
        return f.read()


_description_cache = {}


def golden_description(name):
    """(Architecture, Network) of a golden case, parsed once per process: descriptions are only read by load()
    (the DVS net alone takes seconds to parse and several tests put it on more than one chip)."""
    if name not in _description_cache:
        cwd = os.getcwd()
        os.chdir(ROOT)  # plugin / noise-file paths inside flat files are relative to the repo root
        try:
            _description_cache[name] = sfe.load_flat(golden_flat(name))
        finally:
            os.chdir(cwd)
    return _description_cache[name]


def load_chip(name, device, partition=None):
    cwd = os.getcwd()
    os.chdir(ROOT)  # plugin paths inside flat files are relative to the repo root
    try:
        arch, net = golden_description(name)
        chip = sfe.SpikingChip(arch, device=device)
        chip.set_input_seed_base(0)  # the goldens come from a fresh reference process each
        if partition is not None:
            chip.set_partition(*partition)
        chip.load(net)
    finally:
        os.chdir(cwd)
    return chip


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = np.maximum(np.abs(b), 1e-300)
    return float(np.max(np.where(a == b, 0.0, np.abs(a - b) / scale))) if a.size else 0.0


def check_against_golden(name, chip, rd, out, potential_rtol=0.0, energy_rtol=1e-9):
    """Spike raster + counters bit-exact, potentials to `potential_rtol` (0 = bit-exact),
    energy / latency to `energy_rtol` (north_star: 1e-6), against the reference's outputs."""
    g = golden(name)
    s = g["summary"]
    for key in ("spikes", "packets_sent", "neurons_updated", "neurons_fired"):
        assert getattr(rd, key) == s[key], (name, key, getattr(rd, key), s[key])
    for key in ("total_energy", "synapse_energy", "dendrite_energy", "soma_energy", "network_energy", "sim_time"):
        assert rel_err(getattr(rd, key), s[key]) <= energy_rtol, (name, key, getattr(rd, key), s[key])
    ps = g["per_step"]
    st = out["steps"]
    for key, col in (("fired", "neurons_fired"), ("updated", "neurons_updated"), ("packets", "packets_sent"),
                     ("spikes", "spike_count")):
        assert np.array_equal(st[col], np.asarray(ps[key], dtype=np.int64)), (name, key)
    for key in ("sim_time", "synapse_energy", "dendrite_energy", "soma_energy", "network_energy", "total_energy"):
        assert rel_err(st[key], ps[key]) <= energy_rtol, (name, key, rel_err(st[key], ps[key]))
    text = chip.format_spikes(out["fired_bits"], 1)
    assert text.count("\n") == g["spike_rows"], (name, text.count("\n"), g["spike_rows"])
    assert hashlib.md5(text.encode()).hexdigest() == g["spikes_md5"], name
    assert text == golden_spikes(name), name
    if "neuron_traces_shape" in g and "neuron_traces" in out:
        ref = np.load(os.path.join(GOLDEN, name + ".neuron_traces.npy"))
        assert list(out["neuron_traces"].shape) == g["neuron_traces_shape"], (name, out["neuron_traces"].shape)
        assert np.array_equal(out["neuron_traces"], ref), (name, "model-defined traces (u) differ")
    if "potentials_shape" in g:
        pots = out["potentials"]
        assert list(pots.shape) == g["potentials_shape"], (name, pots.shape)
        full = os.path.join(GOLDEN, name + ".potentials.npy")
        if potential_rtol == 0.0:
            sha = hashlib.sha256(np.ascontiguousarray(pots, dtype="<f8").tobytes()).hexdigest()
            if sha != g["potentials_sha256"] and os.path.exists(full):
                ref = np.load(full)
                bad = np.argwhere(ref != pots)
                raise AssertionError((name, "potentials differ", bad[:5].tolist(), rel_err(pots, ref)))
            assert sha == g["potentials_sha256"], name
        else:
            ref = np.load(full)
            assert rel_err(pots, ref) <= potential_rtol, (name, rel_err(pots, ref))
