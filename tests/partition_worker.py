"""Worker of tests/test_partition_gloo.py: one process per rank of a world_size-2 `gloo` group.

Exercises the host side of the N > 1 path without a GPU: every rank plans the same partition
(sfe_plan_partition), packs the spikes of ITS cores into its raster slice, the slices are
all-gathered (the product does this with ncclAllGather on device buffers), every rank decodes
the global raster back to neuron ids, and the per-rank partial step records are merged.
The spikes come from the oracle (tests may use it as the checker / stimulus)."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import helpers  # noqa: E402
import sanafe_b200 as sfe  # noqa: E402


def main():
    case, steps, out_path = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    chip = helpers.load_chip(case, device=-1)  # host-only: lowering + tables, no engine
    t = chip.tables
    L = sfe.lib()
    owner = np.zeros(t.n_cores, dtype=np.uint32)
    word_begin = np.zeros(t.n_cores, dtype=np.uint32)
    slice_words = C.c_uint32()
    assert L.sfe_plan_partition(C.byref(t), world, owner.ctypes.data, word_begin.ctypes.data, C.byref(slice_words)) == 0
    sw = slice_words.value
    # every rank must plan the same split
    plans = [None] * world
    dist.all_gather_object(plans, (owner.tolist(), word_begin.tolist(), sw))
    assert all(p == plans[0] for p in plans)

    oracle = helpers.Oracle(chip)
    rd, out = oracle.run(steps, potentials=False)
    fired_bits = out["fired_bits"]  # (steps, words) in device (core-major) neuron order
    recs = out["steps"]
    n = t.n_neurons
    neuron_fired = np.unpackbits(fired_bits.view(np.uint8), axis=1, bitorder="little")[:, :n].astype(bool)

    mismatches = 0
    local_cores = [c for c in range(t.n_cores) if owner[c] == rank]
    for s in range(steps):
        mine = np.zeros(sw, dtype=np.uint32)
        for c in local_cores:
            cd = t.cores[c]
            if cd.neuron_count == 0:
                continue
            bits = neuron_fired[s, cd.neuron_begin:cd.neuron_begin + cd.neuron_count]
            padded = np.zeros(((cd.neuron_count + 31) // 32) * 32, dtype=np.uint8)
            padded[:cd.neuron_count] = bits
            words = np.packbits(padded, bitorder="little").view(np.uint32)
            off = int(word_begin[c]) - rank * sw
            assert 0 <= off and off + len(words) <= sw
            mine[off:off + len(words)] = words
        gathered = [torch.zeros(sw, dtype=torch.int32) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(mine.view(np.int32)))
        raster = np.concatenate([g.numpy().view(np.uint32) for g in gathered])
        # decode: every rank sees every neuron of the chip that fired
        decoded = np.zeros(n, dtype=bool)
        for c in range(t.n_cores):
            cd = t.cores[c]
            if cd.neuron_count == 0:
                continue
            nw = (cd.neuron_count + 31) // 32
            w = raster[word_begin[c]:word_begin[c] + nw]
            decoded[cd.neuron_begin:cd.neuron_begin + cd.neuron_count] = \
                np.unpackbits(w.view(np.uint8), bitorder="little")[:cd.neuron_count].astype(bool)
        mismatches += int((decoded != neuron_fired[s]).sum())

    # partial records: this rank's share of the counts (cores it owns), sim_time = local max
    share = np.zeros(steps, dtype=sfe.STEP_DTYPE)
    frac = len(local_cores) / float(t.n_cores)
    for name in share.dtype.names:
        if name == "sim_time":
            share[name] = recs[name] * (1.0 if rank == world - 1 else 0.5)
        elif share.dtype[name].kind == "i":
            share[name] = recs[name] // world + (recs[name] % world if rank == 0 else 0)
        else:
            share[name] = recs[name] * frac
    parts = [None] * world
    dist.all_gather_object(parts, share)
    merged = sfe.merge_partition_records(parts)
    total = sfe.run_data_from_records(merged)
    fracs = [None] * world
    dist.all_gather_object(fracs, frac)
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump({"mismatches": mismatches, "slice_words": sw, "owner": owner.tolist(),
                       "spikes": total.spikes, "ref_spikes": int(rd.spikes),
                       "fired": total.neurons_fired, "ref_fired": int(rd.neurons_fired),
                       "sim_time": total.sim_time, "ref_sim_time": float(rd.sim_time),
                       "energy": total.total_energy, "ref_energy": float(rd.total_energy),
                       "frac_sum": sum(fracs)}, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
