"""Per-rank kernel costs of the partitioned step, measured on ONE GPU: the full bench
workload split over `world` engines that all live on device 0 (the exchange is a device
copy). Run it under `ncu --metrics gpu__time_duration.sum` to get per-kernel durations of
every rank without paying for `world` GPUs; plain, it checks the merged records against
an unpartitioned engine.

    python tools/emulate_partition.py --world 8 --steps 6 [--cores 1024] [--check]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sana-fe_b200"))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import sanafe_b200 as sfe  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--cores", type=int, default=bench.FULL["cores"])
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    import tempfile
    from sanafe_b200 import archgen
    spec_d = dict(bench.FULL)
    spec_d["cores"] = args.cores
    spec_d["dest_cores"] = min(spec_d["dest_cores"], args.cores)
    spec = sfe.SynthSpec(**spec_d)
    flat = os.path.join(tempfile.mkdtemp(prefix="sfe_emul_"), "arch.jsonl")
    archgen.write_flat(archgen.loihi_large(tiles=max(1, (args.cores + 3) // 4) if args.cores < 4096 else 1024), flat)
    arch, _ = sfe.load_flat(flat)

    def make(device, rank, world):
        chip = sfe.SpikingChip(arch, device=device)
        chip.set_partition(rank, world)
        chip.load_synthetic(spec, generate_on_device=True)
        return chip

    part = sfe.PartitionedChip(make, args.world)
    for _ in range(args.steps):
        part.step()
    recs = part.collect()
    print("partitioned: events/step", recs["spike_count"].tolist(), "fired", recs["neurons_fired"].tolist())
    if args.check:
        whole = sfe.SpikingChip(arch, device=0)
        whole.load_synthetic(spec, generate_on_device=True)
        rd = whole.sim_raw(args.steps)[0]
        ok = int(recs["spike_count"].sum()) == rd.spikes and int(recs["neurons_fired"].sum()) == rd.neurons_fired
        rel = abs(float(recs["total_energy"].sum()) - rd.total_energy) / rd.total_energy
        print("whole chip:  events", rd.spikes, "fired", rd.neurons_fired, "match", ok, "energy rel", rel)
        assert ok and rel < 1e-9


if __name__ == "__main__":
    main()
