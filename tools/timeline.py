#!/usr/bin/env python3
"""Per-CTA timeline of the fused step kernel on the benchmark workload (SFE_TIMELINE=1): where a step's time goes.
Writes gpurun_out/timeline.npy [64 steps][grid][16 stamps] (ns) and prints per-phase means."""
import ctypes as C
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sana-fe_b200"))
sys.path.insert(0, ROOT)
os.environ["SFE_TIMELINE"] = "1"
import sanafe_b200 as sfe  # noqa: E402
from sanafe_b200 import archgen  # noqa: E402
import bench  # noqa: E402

cores = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
L = sfe.lib()
spec_d = dict(bench.FULL)
spec_d["cores"] = cores
tmp = tempfile.mkdtemp()
flat = os.path.join(tmp, "arch.jsonl")
archgen.write_flat(archgen.loihi_large(tiles=max(1, (cores + 3) // 4)), flat)
arch, _ = sfe.load_flat(flat)
chip = sfe.SpikingChip(arch, device=0)
chip.load_synthetic(sfe.SynthSpec(**spec_d), generate_on_device=True)
eng = chip.engine
rd = sfe.RunData()
assert L.sfe_engine_enqueue(eng, 100) == 0
assert L.sfe_engine_collect(eng, C.byref(rd)) == 0
buf = np.zeros(64 * 1024 * 16, dtype=np.uint64)
grid = L.sfe_engine_read_timeline(eng, buf.ctypes.data, buf.size)
assert grid > 0, "no timeline recorded"
tl = buf[: 64 * grid * 16].reshape(64, grid, 16).astype(np.int64)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.save(os.path.join(ROOT, "gpurun_out", "timeline.npy"), tl)
steps = [k for k in range(64) if k not in (99 % 64,)][8:40]  # steady-state steps (not the last of the batch)
t0 = tl[:, :, 0].min(axis=1)  # first CTA entry of each step's launch
def us(x):
    return x / 1e3
rows = []
for k in steps:
    a = tl[k]
    rows.append([us(a[:, 1].max() - t0[k]), us(np.median(a[:, 1]) - t0[k]),           # ready observed (last / median CTA)
                 us(np.median(a[:, 3] - a[:, 2])), us(np.median(a[:, 4] - a[:, 3])), us(np.median(a[:, 5] - a[:, 4])),
                 us(np.median(np.where(a[:, 6] > 0, a[:, 9] - a[:, 6], 0))),
                 us(a[:, 15].max() - t0[k]), us(np.median(a[:, 15]) - t0[k]),
                 us(t0[(k + 1) % 64] - t0[k]) if (k + 1) % 64 in steps or True else 0])
r = np.array(rows)
names = ["ready seen (last CTA)", "ready seen (median)", "item1 prologue", "item1 stream", "item1 completion+soma",
         "item2 total", "kernel end (last CTA)", "kernel end (median)", "launch-to-launch"]
for n, c in zip(names, r.mean(axis=0)):
    print(f"{n:>28}: {c:8.1f} us")
