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
  roofline  message-phase kernel (fanout_kernel): algorithmic bytes 12 B/synaptic
          event + 16 B/message over its CUDA-event duration
  cpu_baseline  the reference's own C++ simulator (oracle/_ref/sanafe_ref, built
          unmodified from the reference sources) on a bounded sample of the same
          generator, all host cores

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


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per fanout_kernel launch from the committed
    `ncu --set full` capture of this same workload (profiles/), or None."""
    path = os.path.join(ROOT, "profiles", "r1_ncu_full_summary.json")
    if not os.path.exists(path):
        return None
    def gb(text):
        val, unit = text.split()[:2]
        return float(val) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[unit]
    vals = [gb(r["dram__bytes_read.sum"]) + gb(r["dram__bytes_write.sum"])
            for r in json.load(open(path)) if "fanout_kernel" in r["Kernel Name"]]
    return sum(vals) / len(vals) if vals else None


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
    s.update(soma_hw_name="loihi_lif", synapse_hw_name="loihi_dense_synapse", dendrite_hw_name="loihi_dendrites_delay")
    return s


def run_reference_sample(steps, warmup, cores=8):
    """The reference's own simulator (CPU, OpenMP over cores, -N = all host cores)."""
    from sanafe_b200 import archgen
    threads = os.cpu_count() or 1
    if not os.path.exists(REF_BIN):
        return None
    tmp = tempfile.mkdtemp(prefix="sfe_ref_")
    flat = os.path.join(tmp, "sample.jsonl")
    archgen.write_flat(archgen.loihi_large(tiles=(cores + 3) // 4), flat, synth=sample_spec(cores))
    out = os.path.join(tmp, "out")
    reps = 2 if warmup > 0 else 1
    res = subprocess.run([REF_BIN, flat, "--steps", str(steps), "--timing", "simple", "--threads", str(threads),
                          "--out", out, "--reps", str(reps)], capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stderr[-2000:])
        return None
    with open(os.path.join(out, "summary.json")) as f:
        summ = json.load(f)
    wall = summ["sim_wall_s"]  # last sim() call (earlier ones warm up); load/mapping excluded
    return {"value": summ["spikes"] / wall, "unit": UNIT, "cores": threads, "kind": "reference",
            "sample": f"{cores} of 1024 cores (8192 neurons x fan-out {FULL['dest_cores'] * FULL['syn_per_axon']}, "
                      f"{summ['synapses']} synapses), {steps} steps (after {reps - 1} warm-up sim() call), "
                      f"{summ['spikes']} synaptic events, {wall:.3f} s",
            "ms_per_step": 1e3 * wall / steps, "steps_per_s": steps / wall}


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
    ap.add_argument("--no-dse", action="store_true", help="skip the design-space-sweep side measurement (N=1 only)")
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
    # canonical algorithmic bytes = 12 B per synaptic event (fp64 weight + post index)
    # + 16 B per message. The engine stores certified cores' synapses as lossless 4-byte records,
    # so the canonical figure can exceed the peak; `moved` is what the kernel really pulls from HBM.
    fan_bytes = 12.0 * rd2.spikes + 16.0 * rd2.packets_sent          # over args.steps launches
    record_bytes = 4.0 if os.environ.get("SFE_SYN_Q4", "1") != "0" else 12.0
    layout_bytes = record_bytes * rd2.spikes + 16.0 * rd2.packets_sent
    fan_s = ms_fan.value / 1e3
    achieved = fan_bytes / fan_s / 1e9 if fan_s > 0 else 0.0
    step_bytes = 12.0 * events + 16.0 * messages + 48.0 * n * args.steps
    traffic = ncu_traffic() if args.cores == FULL["cores"] else None
    roofline = {"bound": "hbm", "kernel": "fanout_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "traffic_source": "profiles/r1_ncu_full_summary.json (ncu --set full, DRAM read+write bytes per launch)",
                "peak_source": peak_src,
                "kernel_ms_per_launch": ms_fan.value / max(args.steps, 1),
                "kernel_share_of_step": ms_fan.value / ms_total2.value if ms_total2.value > 0 else None,
                "algorithmic_bytes_per_launch": fan_bytes / max(args.steps, 1),
                "note": "canonical bytes (12 B/event, SURVEY 8d) over the kernel time; the engine reads lossless "
                        f"{int(record_bytes)}-byte synapse records, see `moved`",
                "moved": {"bytes_per_launch_layout": layout_bytes / max(args.steps, 1),
                          "achieved": layout_bytes / fan_s / 1e9 if fan_s > 0 else None,
                          "frac": layout_bytes / fan_s / 1e9 / peak if fan_s > 0 else None,
                          "ncu_dram_frac": (traffic / (fan_s / max(args.steps, 1)) / 1e9 / peak) if (traffic and fan_s > 0) else None},
                "whole_step": {"achieved": step_bytes / (ms_total.value / 1e3) / 1e9,
                               "frac": step_bytes / (ms_total.value / 1e3) / 1e9 / peak,
                               "bytes_per_step": step_bytes / max(args.steps, 1)}}

    # ---- end to end through the public C-ABI with HOST buffers ---------------------------
    # per step: the step's bias vector host->device from pinned memory (8 B per neuron), one
    # timestep, the spike raster + the step record device->host. The upload of step t+1's inputs
    # is issued before step t's results are awaited (set_bias is double-buffered), the way a
    # streaming caller would use the API.
    bias_ptr = [L.sfe_host_alloc(8 * n) for _ in range(2)]
    for ptr in bias_ptr:
        np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(n,))[:] = np.ctypeslib.as_array(tb.neuron_bias, shape=(n,))
    e2e_steps = max(10, min(args.steps, 100))
    nb = C.c_size_t()
    L.sfe_engine_fired_global_ptr(eng, C.byref(nb))
    words = nb.value // 4
    raster_ptr = L.sfe_host_alloc(4 * words)
    rde = sfe.RunData()

    def e2e_loop(count):
        total = 0
        assert L.sfe_engine_set_bias(eng, bias_ptr[0], n) == 0
        for s in range(count):
            assert L.sfe_engine_enqueue(eng, 1) == 0, L.sfe_last_error()
            assert L.sfe_engine_set_bias(eng, bias_ptr[(s + 1) & 1], n) == 0      # inputs of the next step
            assert L.sfe_engine_read_raster(eng, raster_ptr, words) == 0, L.sfe_last_error()  # results of this step
            assert L.sfe_engine_collect(eng, C.byref(rde)) == 0
            total += rde.spikes
        return total

    e2e_loop(3)
    t_e2e = time.perf_counter()
    e2e_events = e2e_loop(e2e_steps)
    L.sfe_engine_synchronize(eng)
    e2e_s = time.perf_counter() - t_e2e
    for ptr in bias_ptr:
        L.sfe_host_free(ptr)
    L.sfe_host_free(raster_ptr)
    e2e = {"value": e2e_events / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 4 * words + 88,
           "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
           "note": "C-ABI calls with pinned host buffers; step t+1's bias upload overlaps step t (double-buffered set_bias)"}

    cpu = None if args.no_cpu_baseline else run_reference_sample(100, 1)  # ~1 s timed; loading its 8 M synapses dominates
    dse = None if args.no_dse else dse_side_measurement()
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
        "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        # BASELINE configs[4] (design-space sweep batched on one GPU): a side measurement, not the headline metric
        "dse": dse,
    }
    print(json.dumps(line))
    return 0


def dse_side_measurement():
    """A 16-point slice of the config-5 sweep (tools/dse_sweep.py) in a child process, so that nothing it does can
    cost the headline line; returns its JSON or the reason it is missing."""
    import subprocess
    cmd = [sys.executable, os.path.join(ROOT, "tools", "dse_sweep.py"), "--mappings", "2", "--multipliers", "8",
           "--steps", "200", "--threads", "16"]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    try:
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=180, env=env)
        if res.returncode != 0:
            return {"error": (res.stderr or res.stdout)[-300:]}
        return json.loads(res.stdout.strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001 - a side measurement must never fail the bench
        return {"error": repr(e)[:300]}


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
    xerr = L.sfe_engine_exchange_error(eng)
    dist.barrier()
    if exchange == "p2p":
        L.sfe_engine_p2p_detach(eng)
    else:
        L.sfe_engine_comm_destroy(eng)
    if xerr != 0:
        raise SystemExit("bench.py: a peer did not arrive at the raster exchange in time; results invalid")
    if rank != 0:
        dist.destroy_process_group()
        return 0
    step_bytes = 12.0 * events_all + 16.0 * messages_all + 48.0 * n * args.steps
    fan_bytes = 12.0 * events_all + 16.0 * messages_all
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
        "e2e": {"value": events_all / seconds, "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0, "note": "N>1 reports the device-timed value; the host-buffer e2e leg is the N=1 run"},
        "gpu_launches": launches_all, "clocks": clocks,
    }
    print(json.dumps(line))
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
