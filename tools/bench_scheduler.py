#!/usr/bin/env python3
"""Host-side cost of the `detailed` timing model (no GPU needed): schedules the DVS golden's timesteps from the
CPU restatement's status bytes with 1, 2, 4, 8 and all host threads and checks that every thread count gives the
same per-step sim_time. These are the scheduler numbers quoted in DESIGN.md §8 (f-1).

    python tools/bench_scheduler.py [--case dvs] [--steps 400]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sana-fe_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import sanafe_b200 as sfe  # noqa: E402
from helpers import Oracle, load_chip  # noqa: E402  (the restatement only supplies the status bytes)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="dvs")
    ap.add_argument("--steps", type=int, default=400)
    args = ap.parse_args()
    chip = load_chip(args.case, device=-1)
    rd, out = Oracle(chip).run(args.steps, status=True, potentials=False)
    status = np.ascontiguousarray(out["status"])
    print(f"{args.case}: {args.steps} timesteps, {rd.packets_sent / args.steps:.0f} messages per timestep")
    first = None
    for threads in (1, 2, 4, 8, 0):
        sfe.lib().sfe_chip_set_scheduler_threads(chip._h, threads)
        sim_time = np.zeros(args.steps)
        t0 = time.perf_counter()
        rc = sfe.lib().sfe_chip_schedule_detailed(chip._h, status.ctypes.data, args.steps, sim_time.ctypes.data)
        dt = time.perf_counter() - t0
        assert rc == 0, sfe.lib().sfe_last_error()
        first = sim_time if first is None else first
        label = threads if threads else f"all ({os.cpu_count()})"
        print(f"  threads {label}: {1e3 * dt / args.steps:.3f} ms per timestep, identical to 1 thread: {np.array_equal(first, sim_time)}")


if __name__ == "__main__":
    main()
