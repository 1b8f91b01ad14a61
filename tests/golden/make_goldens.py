#!/usr/bin/env python3
"""Regenerates tests/golden/ by running the REFERENCE ITSELF (oracle/_ref/sanafe_ref,
the unmodified reference engine built from /root/reference/src).

Run here (container with /root/reference); the outputs are committed so that the
GPU box, which has no /root/reference, can check the CUDA engine and the CPU
restatement against them. Every case writes
    <case>.jsonl[.gz]     flat description fed to the reference AND to the new engine
    <case>.golden.json    RunData totals + per-step records (full precision)
    <case>.spikes.txt[.gz] spike rows "group.offset,timestep" in reference trace order
    <case>.potentials.npy  probed potentials per step (float64), or a sha256 + head
"""
import gzip
import hashlib
import json
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import yaml_to_flat  # noqa: E402
from helpers import load_ref_potentials, load_ref_spikes, load_ref_steps, load_ref_traces, run_reference  # noqa: E402

REF = "/root/reference"
SRC = os.path.join(HERE, "src")

SYNTH_DEFAULT = {
    "cores": 8, "neurons_per_core": 64, "dest_cores": 4, "syn_per_axon": 16, "seed": 1,
    "bias_permille": 100, "bias": 128.0, "threshold": 64.0, "reset": 0.0, "leak_decay": 0.9,
    "w_min": -8, "w_max": 8, "max_delay": 0, "log_spikes": 1, "log_potential_n": 96,
    "soma_hw_name": "loihi_lif", "synapse_hw_name": "loihi_dense_synapse",
    "dendrite_hw_name": "loihi_dendrites_delay",
}

CASES = {
    # BASELINE configs[0]
    "example": dict(arch=f"{REF}/arch/example_chip.yaml", net=f"{REF}/snn/example_snn.yaml", steps=100),
    # BASELINE configs[1] (simple timing here; detailed sim_time is a separate golden)
    "dvs": dict(arch=f"{REF}/arch/loihi.yaml", net=f"{REF}/snn/dvs.yaml", steps=1000, gz=True, threads=8,
                pot_hash=True),
    # BASELINE configs[2]
    "hh": dict(arch=f"{SRC}/hh_arch.yaml", net=f"{SRC}/hh_snn.yaml", steps=1000),
    # BASELINE configs[3], scaled down: loihi_large (buffer inside dendrite) + delay dendrite
    "synth_small": dict(arch=f"{REF}/arch/loihi_large.yaml", max_tiles=2, synth=dict(SYNTH_DEFAULT), steps=60),
    "synth_delay": dict(arch=f"{REF}/arch/loihi_large.yaml", max_tiles=4,
                        synth=dict(SYNTH_DEFAULT, cores=12, dest_cores=5, max_delay=5, seed=7, w_min=-6, w_max=9),
                        steps=80),
    # the charge-loss quirk: plain accumulator with the buffer inside the dendrite unit
    "synth_quirk": dict(arch=f"{REF}/arch/loihi_large.yaml", max_tiles=2,
                        synth=dict(SYNTH_DEFAULT, dendrite_hw_name="loihi_dendrites"), steps=30),
    # buffer before soma + accumulator on loihi.yaml, wider cores
    "synth_soma": dict(arch=f"{REF}/arch/loihi.yaml", max_tiles=4,
                       synth=dict(SYNTH_DEFAULT, cores=16, neurons_per_core=128, dest_cores=6, syn_per_axon=40,
                                  dendrite_hw_name="loihi_dendrites", seed=3, w_min=-12, w_max=10), steps=60),
    # BASELINE configs[4] soma model
    "truenorth": dict(arch=f"{REF}/arch/truenorth.yaml", net=f"{SRC}/tn_snn.yaml", max_tiles=8, steps=120),
    # TrueNorth threshold jitter: std::rand() & random_mask (one processing thread: the draws follow the update order)
    "truenorth_rand": dict(arch=f"{REF}/arch/truenorth.yaml", net=f"{SRC}/tn_rand_snn.yaml", max_tiles=8, steps=200),
    # ordered fp64 accumulation
    "frac": dict(arch=f"{REF}/arch/example_chip.yaml", net=f"{SRC}/frac_snn.yaml", steps=200),
    # Poisson inputs: libstdc++ mt19937 streams seeded by the InputModel construction order
    # LIF file noise stream + model-defined traces (log_u)
    "noise": dict(arch=f"{SRC}/noise_arch.yaml", net=f"{SRC}/noise_snn.yaml", steps=150),
    # "taps" dendrites (MultiTapModel1D): 1-D RC lines, synapses addressed to taps
    "taps": dict(arch=f"{SRC}/taps_arch.yaml", net=f"{SRC}/taps_snn.yaml", steps=200),
    "poisson": dict(arch=f"{REF}/arch/example_chip.yaml", net=f"{SRC}/poisson_snn.yaml", steps=300),
    # the NeuroFEM plugin: a combined dendrite + soma unit with two compartments, buffer inside the soma unit, sigma_v = 0
    "neurofem": dict(arch=f"{SRC}/neurofem_arch.yaml", net=f"{SRC}/neurofem_snn.yaml", steps=400),
}


def make_case(name, spec):
    flat = os.path.join(HERE, name + ".jsonl")
    yaml_to_flat.convert(spec["arch"], spec.get("net"), flat, max_tiles=spec.get("max_tiles"),
                         synth=spec.get("synth"))
    out_dir = tempfile.mkdtemp(prefix="golden_" + name)
    summary = run_reference(flat, out_dir, spec["steps"], "simple", threads=spec.get("threads", 1), per_step=True)
    steps = load_ref_steps(out_dir)
    golden = {
        "case": name, "steps": spec["steps"], "timing": "simple",
        "source": {k: (v.replace(REF, "<reference>").replace(ROOT, "<repo>") if isinstance(v, str) else v)
                   for k, v in spec.items() if k in ("arch", "net", "max_tiles", "synth")},
        "summary": {k: summary[k] for k in (
            "timesteps_executed", "spikes", "packets_sent", "neurons_updated", "neurons_fired", "total_energy",
            "synapse_energy", "dendrite_energy", "soma_energy", "network_energy", "sim_time", "power", "neurons",
            "synapses", "mapped_cores", "mapped_tiles")},
        "per_step": {k: [float(x) if steps[k].dtype.kind == "f" else int(x) for x in np.atleast_1d(steps[k])]
                     for k in steps.dtype.names},
    }
    spikes = load_ref_spikes(out_dir)
    golden["spikes_md5"] = hashlib.md5(spikes.encode()).hexdigest()
    golden["spike_rows"] = spikes.count("\n")
    pots = load_ref_potentials(out_dir)
    if pots is not None:
        golden["potentials_shape"] = list(pots.shape)
        golden["potentials_sha256"] = hashlib.sha256(np.ascontiguousarray(pots, dtype="<f8").tobytes()).hexdigest()
        if spec.get("pot_hash"):
            np.save(os.path.join(HERE, name + ".potentials_head.npy"), pots[:25])
        else:
            np.save(os.path.join(HERE, name + ".potentials.npy"), pots)
    trs = load_ref_traces(out_dir)
    if trs is not None:
        golden["neuron_traces_shape"] = list(trs.shape)
        np.save(os.path.join(HERE, name + ".neuron_traces.npy"), trs)
    # detailed timing totals for the same run (host scheduler parity, NEXT row f-1)
    if name in ("example", "dvs", "frac", "truenorth", "synth_delay"):
        out2 = tempfile.mkdtemp(prefix="golden_det_" + name)
        det = run_reference(flat, out2, spec["steps"], "detailed", threads=spec.get("threads", 1), per_step=True)
        dsteps = load_ref_steps(out2)
        golden["detailed"] = {"sim_time": det["sim_time"],
                              "per_step_sim_time": [float(x) for x in np.atleast_1d(dsteps["sim_time"])]}
        shutil.rmtree(out2)
    with open(os.path.join(HERE, name + ".golden.json"), "w") as f:
        json.dump(golden, f)
    if spec.get("gz"):
        with open(flat, "rb") as fi, gzip.open(flat + ".gz", "wb", compresslevel=9) as fo:
            shutil.copyfileobj(fi, fo)
        os.remove(flat)
        with gzip.open(os.path.join(HERE, name + ".spikes.txt.gz"), "wt", compresslevel=9) as f:
            f.write(spikes)
    else:
        with open(os.path.join(HERE, name + ".spikes.txt"), "w") as f:
            f.write(spikes)
    shutil.rmtree(out_dir)
    s = golden["summary"]
    print(f"{name}: neurons={s['neurons']} synapses={s['synapses']} events={s['spikes']} fired={s['neurons_fired']} "
          f"energy={s['total_energy']!r} sim_time={s['sim_time']!r}")


# Trace files written by the reference's OWN writers (sim_trace_* in src/chip.cpp) with one
# processing thread (message ids follow the creation order): tests/golden/traces/<case>.<timing>.*
TRACE_CASES = [("example", "simple", 100), ("example", "detailed", 100), ("synth_small", "detailed", 12),
               ("synth_delay", "simple", 12), ("truenorth", "detailed", 20), ("hh", "simple", 200)]


def make_traces():
    import subprocess
    out_root = os.path.join(HERE, "traces")
    os.makedirs(out_root, exist_ok=True)
    for name, timing, steps in TRACE_CASES:
        flat = os.path.join(HERE, name + ".jsonl")
        tmp = tempfile.mkdtemp(prefix=f"traces_{name}_{timing}_")
        subprocess.run([os.path.join(ROOT, "oracle", "_ref", "sanafe_ref"), flat, "--steps", str(steps), "--timing", timing,
                        "--threads", "1", "--out", tmp, "--traces"], check=True, stdout=subprocess.DEVNULL)
        kinds = ("messages", "perf", "potentials", "spikes") if name == "hh" else ("messages", "perf")
        for kind in kinds:
            src = os.path.join(tmp, kind + ".csv")
            if kind == "perf" and name != "hh":
                shutil.copy(src, os.path.join(out_root, f"{name}.{timing}.perf.csv"))
            else:
                with open(src, "rb") as fi, gzip.open(os.path.join(out_root, f"{name}.{timing}.{kind}.csv.gz"), "wb", compresslevel=9) as fo:
                    shutil.copyfileobj(fi, fo)
        shutil.rmtree(tmp)
        print(f"traces: {name} {timing} {steps} steps")


if __name__ == "__main__":
    os.chdir(ROOT)  # plugin paths in the flat files are relative to the repo root
    if "--traces" in sys.argv[1:]:
        make_traces()
        sys.exit(0)
    if "--bench-sample" in sys.argv[1:]:
        # 32 cores x 1024 neurons x fan-out 1000 of the benchmark workload through the reference's
        # own engine (needs ~22 GB of host memory), 40 timesteps, every raster hashed -> bench_sample32.hash.json
        sys.path.insert(0, ROOT)
        import bench
        ref = bench.run_reference_sample(1, 0, cores=32, hash_steps=40, calls=1)
        with open(os.path.join(HERE, "bench_sample32.hash.json"), "w") as f:
            json.dump({"what": "BASELINE.md section 3 sample of the benchmark workload: 32 of the 1024 cores x 1024 LIF neurons, "
                               "fan-out 8 x 125 (32.8 M synapses), generator include/sfe_synth.h with bench.FULL's parameters; the "
                               "reference's own engine (oracle/_ref/sanafe_ref) run for 40 timesteps, every raster hashed "
                               "(bench.raster_hash)",
                       "made_by": "tests/golden/make_goldens.py --bench-sample", "sample_cores": 32,
                       "spec": bench.sample_spec(32), "hash": ref["hash"]}, f, indent=1)
        sys.exit(0)
    only = sys.argv[1:]
    for case_name, case_spec in CASES.items():
        if only and case_name not in only:
            continue
        make_case(case_name, case_spec)
