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
        rd = sfe.RunData()
        rc = oracle_lib().sfe_oracle_run(self.h, timesteps, C.byref(req), C.byref(rd))
        assert rc == 0
        return rd, out

    def reset(self):
        oracle_lib().sfe_oracle_reset(self.h)

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


def load_ref_spikes(out_dir):
    with open(os.path.join(out_dir, "spikes_full.csv")) as f:
        return f.read()


def rundata_dict(rd):
    return {name: getattr(rd, name) for name, _ in rd._fields_}
