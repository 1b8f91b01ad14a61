#!/usr/bin/env python3
"""bench.py — headline benchmark of the SANA-FE time-step hot path on B200.

Workload (BASELINE.json configs[3], the configuration the metric is quoted on):
synthetic random LIF network on a loihi_large-shaped chip — 1,048,576 neurons on
1024 cores x 1024, fan-out 1000 (8 destination cores x 125 synapses) = 1.05e9
synapses, ~10 % of neurons firing per step, simple timing model. A "step" is one
simulated timestep (neuron phase -> message phase -> energy/timing reduction).

  value   synaptic events/s over the timed steps, network resident in HBM
  e2e     same metric through the public API with HOST buffers every step: the
          per-neuron bias vector goes host->device (the reference's per-frame
          MappedNeuron.set_attributes pattern, scripts/tcad2025/dvs_gesture.py) and
          the spike raster + step record come back device->host
  roofline  message-phase kernel (fanout_kernel): `frac` = bytes the kernel really
          moves (4 B lossless record/synaptic event + 16 B/message) over its
          CUDA-event duration and the measured HBM peak; `canonical` = the same with
          SURVEY.md 8d's 12 B/event (exceeds the peak: the records are narrowed)
  cpu_baseline  the reference's own C++ simulator (oracle/_ref/sanafe_ref, built
          unmodified from the reference sources) on a bounded sample of the same
          generator (32 of the 1024 cores), all host cores, best of 3 calls
  parity_vs_reference  the GPU engine on that very sample and the same timesteps:
          raster hash + counters equal the reference's, energies to 1e-9
  raster_sha  sha256 of the rasters of 16 timesteps after a reset (same value at
          every GPU count: the partitioned engine gives the unpartitioned rasters)

`--impl reference` times only that CPU reference arm.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "sana-fe_b200"))

METRIC = "synaptic_events_per_s"
UNIT = "synaptic events/s"
WORKLOAD = "synthetic LIF, loihi_large: 1024 cores x 1024 neurons, fan-out 8x125, ~10% activity, simple timing"

FULL = dict(cores=1024, neurons_per_core=1024, dest_cores=8, syn_per_axon=125, seed=1, bias_permille=100,
            bias=128.0, threshold=64.0, reset=0.0, leak_decay=0.9, w_min=-9, w_max=8, max_delay=0,
            log_spikes=0, log_potential_n=0)
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "sanafe_ref")
SAMPLE_CORES = 32   # 32 cores x 1024 neurons x fan-out 1000 (32.8 M synapses, ~22 GB in the reference)
HASH_STEPS = 40     # timesteps compared raster by raster between the reference and the GPU engine on the sample
RASTER_SHA_STEPS = 16


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per fanout_kernel launch from the newest committed
    `ncu --set full` capture of this same workload (profiles/r*_ncu_full_summary.json), or (None, None)."""
    import glob
    paths = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_full_summary.json")))
    if not paths:
        return None, None
    path = paths[-1]
    def gb(text):
        val, unit = text.split()[:2]
        return float(val) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[unit]
    vals = [gb(r["dram__bytes_read.sum"]) + gb(r["dram__bytes_write.sum"])
            for r in json.load(open(path)) if "fanout_kernel" in r["Kernel Name"]]
    if not vals:
        return None, None
    return sum(vals) / len(vals), os.path.relpath(path, ROOT) + " (ncu --set full, DRAM read+write bytes per launch)"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.samples = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            parts = [p.strip() for p in s.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def sample_spec(cores):
    """Bounded sample of the same generator for the CPU reference: `cores` of the
    1024 cores, everything else (neurons/core, fan-in 1000 per neuron, weights) unchanged."""
    s = dict(FULL)
    s["cores"] = cores
    s["dest_cores"] = min(FULL["dest_cores"], cores)
    s["log_spikes"] = 1  # get_spikes() only reports neurons with log_spikes (no trace is written: no cost in the timed calls)
    return s


def sample_cores():
    """32 cores need ~22 GB in the reference (686 B per synapse); fall back to 8 on a small host."""
    try:
        with open("/proc/meminfo") as f:
            avail_kb = next(int(line.split()[1]) for line in f if line.startswith("MemAvailable"))
        return SAMPLE_CORES if avail_kb > 40 * 1024 * 1024 else 8
    except Exception:  # noqa: BLE001
        return 8


def mix64(x):
    """sfe_mix64 (include/sfe_synth.h) on a numpy uint64 array."""
    import numpy as np
    x = x.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def raster_hash(fired_bits):
    """Order-independent hash of a raster [steps][words] (u32, device index order): sum of
    mix64(timestep << 32 | neuron) mod 2^64, timestep 1-based. oracle/ref_harness.cpp --hash-steps computes the same."""
    import numpy as np
    total = np.uint64(0)
    with np.errstate(over="ignore"):
        for t in range(fired_bits.shape[0]):
            bits = np.unpackbits(np.ascontiguousarray(fired_bits[t]).view(np.uint8), bitorder="little")
            idx = np.flatnonzero(bits).astype(np.uint64)
            total = total + mix64((np.uint64(t + 1) << np.uint64(32)) | idx).sum(dtype=np.uint64)
    return int(total)


def run_reference_sample(steps, warmup, cores=None, hash_steps=HASH_STEPS, calls=3):
    """The reference's own simulator (CPU, OpenMP over cores, -N = all host cores): `hash_steps` timesteps hashed
    raster by raster, then 1 warm-up + `calls` timed sim() calls of `steps` timesteps; best rate of the timed calls."""
    from sanafe_b200 import archgen
    threads = os.cpu_count() or 1
    if not os.path.exists(REF_BIN):
        return None
    cores = cores or sample_cores()
    tmp = tempfile.mkdtemp(prefix="sfe_ref_")
    flat = os.path.join(tmp, "sample.jsonl")
    spec = sample_spec(cores)
    spec.update(soma_hw_name="loihi_lif", synapse_hw_name="loihi_dense_synapse", dendrite_hw_name="loihi_dendrites_delay")
    archgen.write_flat(archgen.loihi_large(tiles=(cores + 3) // 4), flat, synth=spec)
    out = os.path.join(tmp, "out")
    reps = (1 if warmup > 0 else 0) + calls
    res = subprocess.run([REF_BIN, flat, "--steps", str(steps), "--timing", "simple", "--threads", str(threads),
                          "--out", out, "--reps", str(reps), "--hash-steps", str(hash_steps)], capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stderr[-2000:])
        return None
    with open(os.path.join(out, "summary.json")) as f:
        summ = json.load(f)
    timed = summ["calls"][1:] if warmup > 0 and len(summ["calls"]) > 1 else summ["calls"]
    events, wall = max(timed, key=lambda c: c[0] / c[1])  # load/mapping excluded
    return {"value": events / wall, "unit": UNIT, "cores": threads, "kind": "reference",
            "sample": f"{cores} of 1024 cores ({cores * FULL['neurons_per_core']} neurons x fan-out "
                      f"{min(FULL['dest_cores'], cores) * FULL['syn_per_axon']}, {summ['synapses']} synapses), best of {len(timed)} "
                      f"sim() calls of {steps} steps (after {len(summ['calls']) - len(timed)} warm-up call and {hash_steps} hashed steps): "
                      f"{events} synaptic events in {wall:.3f} s on {threads} threads",
            "ms_per_step": 1e3 * wall / steps, "steps_per_s": steps / wall, "sample_cores": cores,
            "hash": {k: summ.get(k) for k in ("hash_steps", "raster_hash", "hash_spikes", "hash_packets", "hash_updated",
                                              "hash_fired", "hash_energy", "hash_sim_time")}}


def gpu_sample_parity(ref, device):
    """The GPU engine on the reference arm's sample, same timesteps: raster hash and counters must be equal,
    energy / simulated time to 1e-9 (north_star: rasters and counts bit-exact, totals within 1e-6)."""
    import sanafe_b200 as sfe
    from sanafe_b200 import archgen
    h = ref["hash"]
    if not h or not h.get("hash_steps"):
        return None
    cores = ref["sample_cores"]
    tmp = tempfile.mkdtemp(prefix="sfe_parity_")
    flat = os.path.join(tmp, "arch.jsonl")
    archgen.write_flat(archgen.loihi_large(tiles=(cores + 3) // 4), flat)
    arch, _ = sfe.load_flat(flat)
    chip = sfe.SpikingChip(arch, device=device)
    chip.load_synthetic(sfe.SynthSpec(**sample_spec(cores)), generate_on_device=True)
    rd, out = chip.sim_raw(int(h["hash_steps"]), "simple", steps=True, fired=True)
    got = {"raster_hash": f"{raster_hash(out['fired_bits']):016x}", "hash_spikes": rd.spikes, "hash_packets": rd.packets_sent,
           "hash_updated": rd.neurons_updated, "hash_fired": rd.neurons_fired, "hash_energy": rd.total_energy,
           "hash_sim_time": rd.sim_time}
    exact = all(got[k] == h[k] for k in ("raster_hash", "hash_spikes", "hash_packets", "hash_updated", "hash_fired"))
    rel = max(abs(got[k] - h[k]) / max(abs(h[k]), 1e-300) for k in ("hash_energy", "hash_sim_time"))
    return {"ok": bool(exact and rel <= 1e-9), "steps": int(h["hash_steps"]), "sample_cores": cores,
            "raster_hash_reference": h["raster_hash"], "raster_hash_gpu": got["raster_hash"],
            "counters_equal": bool(exact), "energy_time_max_rel_err": rel,
            "synaptic_events": got["hash_spikes"], "neurons_fired": got["hash_fired"]}


def canonical_raster(words, layout, neuron_counts):
    """The device raster (every core on a word boundary, per-rank slices when partitioned) as core-by-core words."""
    import numpy as np
    parts = [words[b:b + (n + 31) // 32] for b, n in zip(layout, neuron_counts) if n > 0]
    return np.concatenate(parts) if parts else np.zeros(0, dtype=np.uint32)


def raster_sha(L, eng, tb, step_once, drain, steps=RASTER_SHA_STEPS):
    """sha256 over the rasters of `steps` timesteps after a reset: with bias-driven LIF neurons the evolution from the
    zero state does not depend on how many steps ran before, so every GPU count must print the same value."""
    import hashlib
    import numpy as np
    assert L.sfe_engine_reset(eng) == 0, L.sfe_last_error()
    nb = C.c_size_t()
    L.sfe_engine_fired_global_ptr(eng, C.byref(nb))
    words = np.zeros(nb.value // 4, dtype=np.uint32)
    layout = np.zeros(tb.n_cores, dtype=np.uint32)
    assert L.sfe_engine_raster_layout(eng, layout.ctypes.data, tb.n_cores) == 0, L.sfe_last_error()
    counts = [tb.cores[c].neuron_count for c in range(tb.n_cores)]
    sha = hashlib.sha256()
    for _ in range(steps):
        step_once()
        assert L.sfe_engine_read_raster(eng, words.ctypes.data, len(words)) == 0, L.sfe_last_error()
        sha.update(canonical_raster(words, layout, counts).tobytes())
    drain()
    return sha.hexdigest()


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    base = run_reference_sample(max(args.steps, 1), args.warmup)
    if base is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/sanafe_ref not built (needs /root/reference)"}))
        return 0
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": base["sample"]},
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--cores", type=int, default=FULL["cores"], help="scale the workload down (debug only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dse", action="store_true", help="skip the design-space-sweep side measurement")
    ap.add_argument("--no-extras", action="store_true", help="skip the record-format variants and the small configurations (N=1 only)")
    ap.add_argument("--quick", action="store_true",
                    help="device-resident throughput and roofline only (what the record-format variants run in a child process)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import numpy as np
    import sanafe_b200 as sfe
    from sanafe_b200 import archgen

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    L = sfe.lib()
    if L.sfe_device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")

    spec_d = dict(FULL)
    spec_d["cores"] = args.cores
    spec_d["dest_cores"] = min(spec_d["dest_cores"], args.cores)
    spec = sfe.SynthSpec(**spec_d)
    tmp = tempfile.mkdtemp(prefix=f"sfe_bench_{rank}_")
    flat = os.path.join(tmp, "arch.jsonl")
    archgen.write_flat(archgen.loihi_large(tiles=max(1, (args.cores + 3) // 4) if args.cores < 4096 else 1024), flat)
    arch, _ = sfe.load_flat(flat)
    chip = sfe.SpikingChip(arch, device=local_rank)
    peak, peak_src = peaks()

    if world > 1:
        return partitioned_arm(args, sfe, chip, spec, rank, local_rank, world, peak, peak_src)

    t0 = time.time()
    chip.load_synthetic(spec, generate_on_device=True)
    load_s = time.time() - t0
    eng = chip.engine
    tb = chip.tables
    n = tb.n_neurons

    # ---- device-resident throughput ------------------------------------------------
    rd = sfe.RunData()
    assert L.sfe_engine_enqueue(eng, args.warmup) == 0, L.sfe_last_error()
    assert L.sfe_engine_collect(eng, C.byref(rd)) == 0, L.sfe_last_error()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = L.sfe_engine_launch_count(eng)
    ms_total, ms_fan = C.c_float(), C.c_float()
    # timed region: K steps back to back, CUDA events on the engine's stream at both ends only
    assert L.sfe_engine_time_launches(eng, 0) == 0
    assert L.sfe_engine_time_begin(eng) == 0
    assert L.sfe_engine_enqueue(eng, args.steps) == 0, L.sfe_last_error()
    assert L.sfe_engine_time_end(eng, C.byref(ms_total), None) == 0, L.sfe_last_error()
    launches = L.sfe_engine_launch_count(eng) - launches0
    assert L.sfe_engine_collect(eng, C.byref(rd)) == 0, L.sfe_last_error()
    events, messages = rd.spikes, rd.packets_sent
    seconds = ms_total.value / 1e3
    value = events / seconds

    # ---- roofline of the dominant kernel (message phase): a second pass of K steps with CUDA
    # events around every fanout_kernel launch (live, same process, same stream)
    rd2 = sfe.RunData()
    ms_total2 = C.c_float()
    assert L.sfe_engine_time_launches(eng, 1) == 0
    assert L.sfe_engine_time_begin(eng) == 0
    assert L.sfe_engine_enqueue(eng, args.steps) == 0, L.sfe_last_error()
    assert L.sfe_engine_time_end(eng, C.byref(ms_total2), C.byref(ms_fan)) == 0, L.sfe_last_error()
    assert L.sfe_engine_collect(eng, C.byref(rd2)) == 0, L.sfe_last_error()
    clocks = sampler.stop()  # sampled from before the timed region to the end of the roofline pass
    # The kernel's bytes: the engine stores certified cores' synapses as lossless 4-byte records, so one launch moves
    # 4 B per synaptic event + 16 B per message (the layout bytes); `frac` is that over the kernel's CUDA-event time
    # and the measured peak. SURVEY 8(d)'s canonical figure (12 B per event: fp64 weight + post index) is kept as
    # `canonical`; with narrowed records it exceeds the peak and is not a fraction of anything real.
    record_bytes = 4.0 if os.environ.get("SFE_SYN_Q4", "1") != "0" and os.environ.get("SFE_FORCE_ORDERED", "0") == "0" else 12.0
    layout_bytes = record_bytes * rd2.spikes + 16.0 * rd2.packets_sent  # over args.steps launches
    fan_bytes = 12.0 * rd2.spikes + 16.0 * rd2.packets_sent
    fan_s = ms_fan.value / 1e3
    achieved = layout_bytes / fan_s / 1e9 if fan_s > 0 else 0.0
    step_bytes = record_bytes * events + 16.0 * messages + 48.0 * n * args.steps
    traffic, traffic_src = ncu_traffic() if args.cores == FULL["cores"] else (None, None)
    roofline = {"bound": "hbm", "kernel": "fanout_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peak_src,
                "kernel_ms_per_launch": ms_fan.value / max(args.steps, 1),
                "kernel_share_of_step": ms_fan.value / ms_total2.value if ms_total2.value > 0 else None,
                "algorithmic_bytes_per_launch": layout_bytes / max(args.steps, 1),
                "bytes_per_unit": f"{int(record_bytes)} B per synaptic event (lossless record) + 16 B per message",
                "ncu_dram_frac": (traffic / (fan_s / max(args.steps, 1)) / 1e9 / peak) if (traffic and fan_s > 0) else None,
                "canonical": {"bytes_per_launch": fan_bytes / max(args.steps, 1),
                              "achieved": fan_bytes / fan_s / 1e9 if fan_s > 0 else None,
                              "frac": fan_bytes / fan_s / 1e9 / peak if fan_s > 0 else None,
                              "note": "SURVEY 8d: 12 B per event + 16 B per message; above 1.0 because the records are narrowed"},
                "whole_step": {"achieved": step_bytes / (ms_total.value / 1e3) / 1e9,
                               "frac": step_bytes / (ms_total.value / 1e3) / 1e9 / peak,
                               "bytes_per_step": step_bytes / max(args.steps, 1)}}

    if args.quick:
        print(json.dumps({"ms_per_step": 1e3 * seconds / args.steps, "value": value, "events_per_step": events / args.steps,
                          "kernel_ms_per_launch": roofline["kernel_ms_per_launch"], "bytes_per_launch": roofline["algorithmic_bytes_per_launch"],
                          "achieved": roofline["achieved"], "frac": roofline["frac"], "bytes_per_unit": roofline["bytes_per_unit"]}))
        return 0

    # ---- raster checksum (before the end-to-end legs patch the biases) ---------------------------------
    def step_once():
        assert L.sfe_engine_enqueue(eng, 1) == 0, L.sfe_last_error()

    rds = sfe.RunData()
    sha = raster_sha(L, eng, tb, step_once, lambda: L.sfe_engine_collect(eng, C.byref(rds)))

    # ---- end to end through the reference's own call, with HOST buffers ---------------------------
    # The loop of scripts/tcad2025/dvs_gesture.py:140-151: per frame, MappedNeuron.set_attributes(bias) on 1024 input
    # neurons (sfe_chip_set_neuron_attribute; the patches reach the device as one bias vector when sim() starts),
    # then SpikingChip.sim(frame of timesteps) = sfe_chip_sim with a perf trace (per-step records) and a spike trace
    # (fired-bit raster of every step, device -> host) into host buffers.
    frame_steps = max(10, min(args.steps, 100))
    n_frames = 3
    words_per_step = (n + 31) // 32
    e2e_fired = np.zeros((frame_steps, words_per_step), dtype=np.uint32)
    e2e_steps = np.zeros(frame_steps, dtype=sfe.STEP_DTYPE)
    req = sfe.TraceRequest()
    req.steps = e2e_steps.ctypes.data
    req.fired_bits = e2e_fired.ctypes.data
    base_bias = np.ctypeslib.as_array(tb.neuron_bias, shape=(n,))[:1024].copy()
    rde = sfe.RunData()

    def e2e_frame(f):
        for i in range(1024):  # a different bias pattern every frame (rotation keeps the driven fraction)
            assert L.sfe_chip_set_neuron_attribute(chip._h, b"pop", i, b"bias", float(base_bias[(i + f) % 1024])) == 0
        assert L.sfe_chip_sim(chip._h, frame_steps, 0, C.byref(req), C.byref(rde)) == 0, L.sfe_last_error()
        return rde.spikes

    e2e_frame(0)
    t_e2e = time.perf_counter()
    e2e_events = sum(e2e_frame(f + 1) for f in range(n_frames))
    e2e_s = time.perf_counter() - t_e2e
    e2e = {"value": e2e_events / e2e_s, "unit": UNIT,
           "h2d_bytes_per_step": 8 * n // frame_steps, "d2h_bytes_per_step": 4 * words_per_step + 88,
           "steps": frame_steps * n_frames, "ms_per_step": 1e3 * e2e_s / (frame_steps * n_frames),
           "call": "sfe_chip_set_neuron_attribute x1024 + sfe_chip_sim(frame) per frame (the reference's dvs_gesture.py loop)",
           "frame_steps": frame_steps, "frames": n_frames,
           "note": "per frame: 1024 bias patches -> one 8 MB bias vector host->device; per step: raster + step record device->host"}

    # ---- the same metric with per-STEP host traffic through the table-level C ABI (round-1 e2e leg, kept) ----
    bias_ptr = [L.sfe_host_alloc(8 * n) for _ in range(2)]
    for ptr in bias_ptr:
        np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(n,))[:] = np.ctypeslib.as_array(tb.neuron_bias, shape=(n,))
    per_step_n = max(10, min(args.steps, 100))
    nb = C.c_size_t()
    L.sfe_engine_fired_global_ptr(eng, C.byref(nb))
    words = nb.value // 4
    raster_ptr = L.sfe_host_alloc(4 * words)

    def per_step_loop(count):
        total = 0
        assert L.sfe_engine_set_bias(eng, bias_ptr[0], n) == 0
        for s_ in range(count):
            assert L.sfe_engine_enqueue(eng, 1) == 0, L.sfe_last_error()
            assert L.sfe_engine_set_bias(eng, bias_ptr[(s_ + 1) & 1], n) == 0      # inputs of the next step
            assert L.sfe_engine_read_raster(eng, raster_ptr, words) == 0, L.sfe_last_error()  # results of this step
            assert L.sfe_engine_collect(eng, C.byref(rde)) == 0
            total += rde.spikes
        return total

    per_step_loop(3)
    t_ps = time.perf_counter()
    ps_events = per_step_loop(per_step_n)
    L.sfe_engine_synchronize(eng)
    ps_s = time.perf_counter() - t_ps
    for ptr in bias_ptr:
        L.sfe_host_free(ptr)
    L.sfe_host_free(raster_ptr)
    e2e["per_step_host_io"] = {"value": ps_events / ps_s, "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 4 * words + 88,
                               "ms_per_step": 1e3 * ps_s / per_step_n,
                               "call": "sfe_engine_set_bias + sfe_engine_enqueue(1) + sfe_engine_read_raster + sfe_engine_collect per step"}

    cpu = None if args.no_cpu_baseline else run_reference_sample(100, 1)
    parity = gpu_sample_parity(cpu, local_rank) if cpu else None
    dse = None if args.no_dse else dse_side_measurement()
    if not args.no_extras and args.cores == FULL["cores"]:
        # the same workload through the other two message-phase paths (child processes: the record format is chosen at load)
        roofline["variants"] = {
            "records_12B_tma": variant_measurement({"SFE_SYN_Q4": "0"}, "12-byte records (fp64 weight + meta word) through cp.async.bulk + mbarrier: "
                                                                         "the path of cores with delays or without an exactness certificate"),
            "ordered_fp64": variant_measurement({"SFE_FORCE_ORDERED": "1"}, "every core forced to the ordered mode (what non-dyadic weights select): one warp "
                                                                             "per core replays the messages in arrival order with fp64 adds", steps=5),
        }
    small = None if args.no_extras else small_configs(local_rank)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * seconds / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD if args.cores == FULL["cores"] else f"DEBUG scale: {args.cores} cores",
                   "neurons": n, "synapses": int(tb.n_synapses), "timing_model": "simple",
                   "l2_policy": "inputs larger than L2 (12.6 GB of synapse tables; ~1.2 GB streamed per step)",
                   "parallelism": "1 GPU", "activity": rd.neurons_fired / float(n * args.steps), "load_s": load_s},
        "timesteps_per_s": args.steps / seconds, "events_per_step": events / args.steps,
        "roofline": roofline,
        "cpu_baseline": ({k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")} if cpu else None),
        "parity_vs_reference": parity, "raster_sha": sha,
        "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        # BASELINE configs[4] (design-space sweep batched on one GPU): a side measurement, not the headline metric
        "dse": dse,
        # BASELINE configs[0..2] through the drop-in Python module, next to the reference's simulator on the same files
        "small_configs": small,
    }
    print(json.dumps(line))
    return 0


def variant_measurement(env, what, steps=20):
    """bench.py --quick in a child process with `env` (the synapse record format is chosen when the chip is loaded)."""
    cmd = [sys.executable, os.path.abspath(__file__), "--quick", "--steps", str(steps), "--warmup", "3"]
    child_env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    child_env.update(env)
    try:
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=child_env)
        if res.returncode != 0:
            return {"error": (res.stderr or res.stdout)[-300:]}
        out = json.loads(res.stdout.strip().splitlines()[-1])
        out.update({"what": what, "env": env, "steps": steps})
        return out
    except Exception as e:  # noqa: BLE001 - a side measurement must never fail the bench
        return {"error": repr(e)[:300]}


def small_configs(device):
    """BASELINE configs[0..2] (example chip, DVS gesture with detailed timing, Hodgkin-Huxley plugin) as timesteps/s
    through SpikingChip.sim() of the drop-in pybind11 module, next to the reference's own simulator (oracle/_ref) on the
    same description files with all host threads. Descriptions: tests/golden/*.jsonl (made from the reference's files)."""
    import gzip
    try:
        from sanafe_b200 import sanafecpp_b200 as m
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)[:200]}
    threads = os.cpu_count() or 1
    cases = [("config1_example", "example", 100, "simple"), ("config2_dvs", "dvs", 1000, "detailed"), ("config3_hh", "hh", 1000, "simple")]
    out = {}
    cwd = os.getcwd()
    os.chdir(ROOT)  # plugin paths inside the flat files are relative to the repo root
    try:
        for key, name, steps, timing in cases:
            try:
                flat = os.path.join(ROOT, "tests", "golden", name + ".jsonl")
                if not os.path.exists(flat):
                    tmp = tempfile.NamedTemporaryFile(prefix=name + "_", suffix=".jsonl", delete=False)
                    with gzip.open(flat + ".gz", "rb") as f:
                        tmp.write(f.read())
                    tmp.close()
                    flat = tmp.name
                arch, net = m.load_flat(flat)
                chip = m.SpikingChip(arch)
                chip.load(net)
                chip.sim(steps, timing_model=timing)  # warm-up call (first launches, pinned buffers, scheduler pool)
                best, sched = None, 0.0
                for _ in range(3):
                    t0 = time.perf_counter()
                    res = chip.sim(steps, timing_model=timing)
                    wall = time.perf_counter() - t0
                    if best is None or wall < best:
                        best = wall
                entry = {"timesteps": steps, "timing_model": timing, "timesteps_per_s": steps / best, "us_per_timestep": 1e6 * best / steps,
                         "call": "sanafecpp_b200.SpikingChip.sim(timesteps, timing_model) - best of 3 calls after a warm-up call",
                         "synaptic_events": int(res["spikes"])}
                if os.path.exists(REF_BIN):
                    tmpd = tempfile.mkdtemp(prefix="sfe_small_")
                    ref = subprocess.run([REF_BIN, flat, "--steps", str(steps), "--timing", timing, "--threads", str(threads),
                                          "--out", tmpd, "--reps", "4"], capture_output=True, text=True, timeout=300)
                    if ref.returncode == 0:
                        with open(os.path.join(tmpd, "summary.json")) as f:
                            summ = json.load(f)
                        ref_wall = min(c[1] for c in summ["calls"][1:])
                        entry["reference"] = {"timesteps_per_s": steps / ref_wall, "us_per_timestep": 1e6 * ref_wall / steps, "threads": threads,
                                              "synaptic_events": int(summ["calls"][-1][0]),
                                              "call": "oracle/_ref/sanafe_ref (the reference's SpikingChip.sim), best of 3 calls after a warm-up call"}
                        entry["speedup_vs_reference"] = ref_wall / best
                        entry["events_equal"] = int(summ["calls"][-1][0]) == int(res["spikes"])
                out[key] = entry
            except Exception as e:  # noqa: BLE001 - a side measurement must never fail the bench
                out[key] = {"error": repr(e)[:300]}
    finally:
        os.chdir(cwd)
    return out


def dse_side_measurement():
    """128 design points of the config-5 sweep (tools/dse_sweep.py: one after another, batched with one stream per chip,
    and as one launch per phase for all chips) in a child process, so that nothing it does can cost the headline line; returns its JSON
    or the reason it is missing."""
    import subprocess
    cmd = [sys.executable, os.path.join(ROOT, "tools", "dse_sweep.py"), "--mappings", "16", "--multipliers", "8",
           "--steps", "200", "--threads", "16"]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    try:
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
        if res.returncode != 0:
            return {"error": (res.stderr or res.stdout)[-300:]}
        return json.loads(res.stdout.strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001 - a side measurement must never fail the bench
        return {"error": repr(e)[:300]}


def dse_points_of_rank(rank, world):
    """128 design points of the 32 x 32 sweep for every rank: an evenly strided subset of 128 x world points (all 1024 at
    world = 8), dealt round-robin, so that every rank holds every mapping (neurons per core) at several cost multipliers."""
    from sanafe_b200 import dse
    every = dse.sweep_points()
    sel = every[::max(1, 8 // world)][:128 * world]
    return sel[rank::world]


def dse_partitioned(rank, world, local_rank, dist, steps=200):
    """BASELINE configs[4] at N GPUs: 128 design points of the 32 x 32 sweep per GPU (all 1024 at N = 8), replicas only -
    independent simulations, no exchange. Every rank times its own batch; rank 0 reports the slowest rank's wall time.
    No collective inside: a rank that fails reports its error instead of leaving the others at a barrier."""
    mine_out = None
    try:
        from sanafe_b200 import dse
        mine = dse_points_of_rank(rank, world)
        t0 = time.time()
        sweep = dse.Sweep(mine, tempfile.mkdtemp(prefix=f"dse_{rank}_"), device=local_rank, host_threads=16)
        load_s = time.time() - t0
        sweep.sim(10)  # warm-up
        t0 = time.time()
        rds = sweep.sim(steps)
        wall = time.time() - t0
        mine_out = {"points": len(mine), "wall_s": wall, "load_s": load_s, "events": sum(r.spikes for r in rds)}
    except Exception as e:  # noqa: BLE001
        mine_out = {"error": repr(e)[:300]}
    gathered = [None] * world
    dist.all_gather_object(gathered, mine_out)
    if rank != 0:
        return None
    bad = [g for g in gathered if g is None or "error" in g]
    if bad:
        return {"error": str(bad[0])}
    points = sum(g["points"] for g in gathered)
    wall = max(g["wall_s"] for g in gathered)
    return {"workload": "config 5: design-space sweep of the conv SNN on TrueNorth-shaped chips"
                        + (" (all 32 x 32 design points)" if points == 1024 else ""),
            "design_points": points, "points_per_gpu": points // world, "n_gpus": world, "steps": steps,
            "wall_s": round(wall, 4), "sims_per_s": round(points / wall, 1), "timesteps_per_s": round(points * steps / wall, 1),
            "synaptic_events_per_s": round(sum(g["events"] for g in gathered) / wall, 1),
            "load_s_max": round(max(g["load_s"] for g in gathered), 2),
            "mode": "sfe_batch_sim: one stream per chip, 16 host threads per GPU; replicas only across GPUs"}


def nccl_library_path():
    """The NCCL the image ships with torch (dlopen'ed by the engine)."""
    try:
        import nvidia.nccl
        cand = os.path.join(os.path.dirname(nvidia.nccl.__file__), "lib", "libnccl.so.2")
        if os.path.exists(cand):
            return cand
    except Exception:
        pass
    return "libnccl.so.2"


def partitioned_arm(args, sfe, chip, spec, rank, local_rank, world, peak, peak_src):
    """N > 1: ONE chip, its cores partitioned over the ranks (strong scaling). Every step:
    local neuron phase -> ncclAllGather of the fired raster over NVLink -> local message
    phase; the whole run is enqueued from C++ (sfe_engine_enqueue_partitioned), torch.distributed
    only carries the NCCL id and the final reductions."""
    import numpy as np
    import torch
    import torch.distributed as dist

    L = sfe.lib()
    torch.cuda.set_device(local_rank)
    dist.init_process_group("gloo")
    chip.set_partition(rank, world)
    t0 = time.time()
    chip.load_synthetic(spec, generate_on_device=True)
    load_s = time.time() - t0
    eng = chip.engine
    tb = chip.tables
    n = tb.n_neurons
    slice_words = C.c_uint32()
    L.sfe_engine_partition_info(eng, None, None, C.byref(slice_words), None, None)
    exchange = os.environ.get("SFE_EXCHANGE", "p2p")
    if exchange == "p2p":
        # fused exchange: the neuron-phase kernel stores the raster into every peer over NVLink
        mine = np.zeros(64, dtype=np.uint8)
        assert L.sfe_engine_p2p_export(eng, mine.ctypes.data) == 0, L.sfe_last_error()
        everyone = [torch.zeros(64, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(everyone, torch.from_numpy(mine))
        handles = np.concatenate([h.numpy() for h in everyone])
        assert L.sfe_engine_p2p_attach(eng, handles.ctypes.data) == 0, L.sfe_last_error()
        dist.barrier()
    else:
        lib_path = nccl_library_path().encode()
        uid = np.zeros(128, dtype=np.uint8)
        if rank == 0:
            assert L.sfe_nccl_get_unique_id(uid.ctypes.data, lib_path) == 0, L.sfe_last_error()
        uid_t = torch.from_numpy(uid)
        dist.broadcast(uid_t, src=0)
        assert L.sfe_engine_comm_init(eng, uid_t.numpy().ctypes.data, lib_path) == 0, L.sfe_last_error()

    def collect():
        buf = np.zeros(4096, dtype=sfe.STEP_DTYPE)
        got = L.sfe_engine_collect_records(eng, buf.ctypes.data, len(buf))
        assert got >= 0, L.sfe_last_error()
        return buf[:got]

    assert L.sfe_engine_enqueue_partitioned(eng, args.warmup) == 0, L.sfe_last_error()
    assert L.sfe_engine_synchronize(eng) == 0
    collect()
    dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = L.sfe_engine_launch_count(eng)
    ms_total, ms_fan = C.c_float(), C.c_float()
    assert L.sfe_engine_time_begin(eng) == 0
    assert L.sfe_engine_enqueue_partitioned(eng, args.steps) == 0, L.sfe_last_error()
    assert L.sfe_engine_time_end(eng, C.byref(ms_total), C.byref(ms_fan)) == 0, L.sfe_last_error()
    dist.barrier()
    seconds = ms_total.value / 1e3
    clocks = sampler.stop()
    launches = L.sfe_engine_launch_count(eng) - launches0
    recs = collect()
    events = int(recs["spike_count"].sum())
    messages = int(recs["packets_sent"].sum())
    fired = int(recs["neurons_fired"].sum())
    tmax = torch.tensor([seconds], dtype=torch.float64)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tot = torch.tensor([float(events), float(messages), float(fired), float(launches), ms_fan.value / 1e3], dtype=torch.float64)
    dist.all_reduce(tot)
    seconds = float(tmax.item())
    events_all, messages_all, fired_all, launches_all = (int(x) for x in tot.tolist()[:4])
    fan_s_mean = tot.tolist()[4] / world
    def step_once():
        assert L.sfe_engine_enqueue_partitioned(eng, 1) == 0, L.sfe_last_error()

    if os.environ.get("SFE_TIMELINE"):
        # diagnostic build (SFE_TIMELINE_ALL): every rank keeps the stamps of its last 64 steps
        raw = np.zeros(2 * 64 * 1024 * 16, dtype=np.uint64)
        if L.sfe_engine_read_timeline_raw(eng, raw.ctypes.data, raw.size) > 0:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            np.save(os.path.join(ROOT, "gpurun_out", f"timeline_n{world}_r{rank}.npy"), raw.reshape(2, 64, 1024, 16)[:, :, :640, :])
    dist.barrier()
    sha = raster_sha(L, eng, tb, step_once, collect)
    shas = [None] * world
    dist.all_gather_object(shas, sha)
    xerr = L.sfe_engine_exchange_error(eng)
    xmsg = L.sfe_last_error().decode() if xerr != 0 else ""
    dist.barrier()
    dse = None if args.no_dse else dse_partitioned(rank, world, local_rank, dist)
    if exchange == "p2p":
        L.sfe_engine_p2p_detach(eng)
    else:
        L.sfe_engine_comm_destroy(eng)
    if xerr != 0:
        raise SystemExit(f"bench.py: rank {rank}: {xmsg}; results invalid")
    if rank != 0:
        dist.destroy_process_group()
        return 0
    record_bytes = 4.0 if os.environ.get("SFE_SYN_Q4", "1") != "0" else 12.0  # lossless 4-byte records, as at N = 1
    step_bytes = record_bytes * events_all + 16.0 * messages_all + 48.0 * n * args.steps
    fan_bytes = record_bytes * events_all + 16.0 * messages_all
    line = {
        "metric": METRIC, "value": events_all / seconds, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * seconds / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD if args.cores == FULL["cores"] else f"DEBUG scale: {args.cores} cores",
                   "neurons": n, "synapses": int(tb.n_synapses), "timing_model": "simple",
                   "l2_policy": "inputs larger than L2 (synapse tables of each partition >> 126 MB)",
                   "parallelism": f"cores partitioned over {world} GPUs; fired raster ({4 * slice_words.value * world} B/step) "
                                  + ("stored into every peer by the neuron-phase kernel over NVLink (CUDA IPC), flag barrier"
                                     if exchange == "p2p" else "exchanged with ncclAllGather enqueued from C++"),
                   "activity": fired_all / float(n * args.steps), "load_s": load_s},
        "timesteps_per_s": args.steps / seconds, "events_per_step": events_all / args.steps,
        "roofline": {"bound": "hbm", "kernel": "fanout_kernel (all ranks)",
                     "achieved": fan_bytes / fan_s_mean / 1e9 if fan_s_mean > 0 else None,
                     "peak": peak * world, "unit": "GB/s",
                     "frac": (fan_bytes / fan_s_mean / 1e9 / (peak * world)) if fan_s_mean > 0 else None,
                     "traffic": None, "peak_source": peak_src,
                     "kernel_ms_per_launch": 1e3 * fan_s_mean / args.steps,
                     "whole_step": {"achieved": step_bytes / seconds / 1e9, "frac": step_bytes / seconds / 1e9 / (peak * world)}},
        "cpu_baseline": None,
        "raster_sha": sha, "raster_sha_equal_on_all_ranks": all(x == sha for x in shas),
        "e2e": {"value": events_all / seconds, "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0, "note": "N>1 reports the device-timed value; the host-buffer e2e leg is the N=1 run"},
        "gpu_launches": launches_all, "clocks": clocks,
        # BASELINE configs[4]: 128 independent design points per GPU (1024 at 8 GPUs), a side measurement
        "dse": dse,
    }
    print(json.dumps(line))
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
