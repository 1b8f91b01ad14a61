#!/usr/bin/env python3
"""BASELINE config 5 on one GPU: a slice of the 32 x 32 design-space sweep (neurons per core x cost
multiplier) of the synthetic conv SNN on TrueNorth-shaped chips, run as a batch (sfe_batch_sim: one stream per
chip, worker threads keep that many chips in flight) and, for comparison, one chip after another.

    python tools/dse_sweep.py --mappings 4 --multipliers 8 --steps 200 --threads 16

Prints one JSON line: simulations/s, timesteps/s and synaptic events/s over all design points, batched and
sequential (wall clock around the sim calls; loading is reported separately)."""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sana-fe_b200"))
import sanafe_b200 as sfe  # noqa: E402
from sanafe_b200 import dse  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mappings", type=int, default=4)
    ap.add_argument("--multipliers", type=int, default=8)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--threads", type=int, default=16)
    ap.add_argument("--device", type=int, default=0)
    args = ap.parse_args()
    if sfe.lib().sfe_device_count() <= 0:
        raise SystemExit("no CUDA device: the engine has no CPU fallback")
    every = dse.sweep_points()
    npcs = sorted({p[0] for p in every})
    mults = sorted({p[1] for p in every})
    pick_n = [npcs[(i * len(npcs)) // args.mappings] for i in range(args.mappings)]
    pick_m = [mults[(i * len(mults)) // args.multipliers] for i in range(args.multipliers)]
    points = [(n, m) for n in pick_n for m in pick_m]
    work = tempfile.mkdtemp(prefix="dse_")
    t0 = time.time()
    sweep = dse.Sweep(points, work, device=args.device, host_threads=args.threads)
    load_s = time.time() - t0
    sweep.sim(10)  # warm-up (first launches, pinned buffers)
    out = {"workload": "config 5 slice", "design_points": len(points), "steps": args.steps,
           "neurons_per_core": pick_n, "host_threads": args.threads, "load_s": round(load_s, 2)}
    for label, threads in (("sequential", 1), ("batched", args.threads)):
        sweep.host_threads = threads
        t0 = time.time()
        rds = sweep.sim(args.steps)
        wall = time.time() - t0
        events = sum(r.spikes for r in rds)
        out[label] = {"wall_s": round(wall, 4), "sims_per_s": round(len(points) / wall, 2),
                      "timesteps_per_s": round(len(points) * args.steps / wall, 1),
                      "synaptic_events_per_s": round(events / wall, 1)}
    out["batched_over_sequential"] = round(out["sequential"]["wall_s"] / out["batched"]["wall_s"], 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
