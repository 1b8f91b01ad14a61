#!/usr/bin/env python3
"""BASELINE config 5 on one GPU: a slice of the 32 x 32 design-space sweep (neurons per core x cost
multiplier) of the synthetic conv SNN on TrueNorth-shaped chips, run as a batch (sfe_batch_sim: one stream per
chip, worker threads keep that many chips in flight) and, for comparison, one chip after another.

    python tools/dse_sweep.py --mappings 4 --multipliers 8 --steps 200 --threads 16

Prints one JSON line: simulations/s, timesteps/s and synaptic events/s over all design points, batched and
sequential (wall clock around the sim calls; loading is reported separately).

More GPUs: replicas only (independent simulations, no exchange). Under
`python -m torch.distributed.run --nproc-per-node N tools/dse_sweep.py ...` rank r takes design points r, r+N, ...
on device LOCAL_RANK; the ranks meet at a gloo barrier on both sides of the timed region and rank 0 prints the
line with the maximum wall time and the summed counts."""
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


def points_of_rank(points, rank, world):
    """Design points dealt round-robin (mappings and multipliers both vary fastest across ranks)."""
    return points[rank::world]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mappings", type=int, default=4)
    ap.add_argument("--multipliers", type=int, default=8)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--threads", type=int, default=16)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--batched-only", action="store_true", help="skip the sequential and one-stream-per-chip runs")
    ap.add_argument("--csv", default=None, help="write one row per design point (the sweep's actual result) to this file")
    args = ap.parse_args()
    if sfe.lib().sfe_device_count() <= 0:
        raise SystemExit("no CUDA device: the engine has no CPU fallback")
    every = dse.sweep_points()
    npcs = sorted({p[0] for p in every})
    mults = sorted({p[1] for p in every})
    pick_n = [npcs[(i * len(npcs)) // args.mappings] for i in range(args.mappings)]
    pick_m = [mults[(i * len(mults)) // args.multipliers] for i in range(args.multipliers)]
    points = [(n, m) for n in pick_n for m in pick_m]
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch.distributed as dist  # control plane only (barriers, gathering the ranks' numbers)
        dist.init_process_group("gloo")
        args.device = int(os.environ.get("LOCAL_RANK", "0"))
    all_points = len(points)
    points = points_of_rank(points, rank, world)
    work = tempfile.mkdtemp(prefix="dse_")
    t0 = time.time()
    sweep = dse.Sweep(points, work, device=args.device, host_threads=args.threads)
    load_s = time.time() - t0
    sweep.sim(10)  # warm-up (first launches, pinned buffers)
    out = {"workload": "config 5" + (" slice" if all_points < 1024 else " (all 32 x 32 design points)"), "design_points": all_points, "n_gpus": world, "steps": args.steps,
           "neurons_per_core": pick_n, "host_threads": args.threads, "load_s": round(load_s, 2)}
    # sequential: one chip after another; batched: one stream per chip, host threads keep that many in flight (the
    # default of sfe_batch_sim); grid: the batch as one launch per phase and step for all chips (SFE_BATCH_GRID=1)
    modes = (("sequential", 1, "0"), ("batched", args.threads, "0"), ("grid", args.threads, "1"))
    if args.batched_only:
        modes = modes[1:2]
    for label, threads, grid in modes:
        sweep.host_threads = threads
        os.environ["SFE_BATCH_GRID"] = grid
        if dist is not None:
            dist.barrier()
        t0 = time.time()
        rds = sweep.sim(args.steps)
        wall = time.time() - t0
        events = sum(r.spikes for r in rds)
        if dist is not None:
            dist.barrier()
            gathered = [None] * world
            dist.all_gather_object(gathered, (wall, events))
            wall = max(g[0] for g in gathered)
            events = sum(g[1] for g in gathered)
        out[label] = {"wall_s": round(wall, 4), "sims_per_s": round(all_points / wall, 2),
                      "timesteps_per_s": round(all_points * args.steps / wall, 1),
                      "synaptic_events_per_s": round(events / wall, 1)}
    if "sequential" in out:
        out["batched_over_sequential"] = round(out["sequential"]["wall_s"] / out["batched"]["wall_s"], 2)
        out["batched_over_grid"] = round(out["grid"]["wall_s"] / out["batched"]["wall_s"], 2)
    if args.csv:
        # what a design-space exploration is after: energy, simulated time and activity of every design point
        # (of the last, batched, run of --steps timesteps)
        rows = [(npc, mult, r.total_energy, r.sim_time, r.spikes, r.packets_sent, r.neurons_fired)
                for (npc, mult), r in zip(points, rds)]
        if dist is not None:
            gathered = [None] * world
            dist.all_gather_object(gathered, rows)
            rows = [row for part in gathered for row in part]
        if rank == 0:
            with open(args.csv, "w") as f:
                f.write("neurons_per_core,cost_multiplier,energy_j,sim_time_s,synaptic_events,messages,neurons_fired\n")
                for row in sorted(rows):
                    f.write(",".join(repr(x) for x in row) + "\n")
    if rank == 0:
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
