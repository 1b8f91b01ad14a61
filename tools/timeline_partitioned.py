#!/usr/bin/env python3
"""Critical path of a two-kernel step from the %globaltimer stamps of a diagnostic build (SFE_TIMELINE_ALL=1, SFE_TIMELINE=1):
reads gpurun_out/timeline_n<N>_r<rank>.npy ([2 = message | neuron phase][64 steps][CTA][16 stamps], ns) written by
bench.py and prints, per rank, the mean over steady-state steps of every phase boundary relative to the start of the
step's neuron-phase kernel.

message-phase stamps: 12 entry, 13 after griddepcontrol.wait, 0 raster visible (after the exchange wait), 2 first item
begins, 3 its list is ready (inbox scan + axon records), 4 its stream + write-back done, 5 item end, 15 CTA exit.
neuron-phase stamps: 0 entry, 1 after griddepcontrol.wait, 2 segment done, 3 raster slice published (last CTA)."""
import glob
import sys

import numpy as np

pattern = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/timeline_n*_r*.npy"
for path in sorted(glob.glob(pattern)):
    tl = np.load(path).astype(np.int64)
    fan, soma = tl[0], tl[1]
    rows = []
    for k in range(8, 56):
        f, so = fan[k], soma[k]
        live_f = f[:, 12] > 0
        live_s = so[:, 0] > 0
        if not live_f.any() or not live_s.any():
            continue
        t0 = so[live_s, 0].min()
        nxt = soma[(k + 1) % 64]
        nxt0 = nxt[nxt[:, 0] > 0, 0].min() if (nxt[:, 0] > 0).any() else t0
        if nxt0 <= t0 or nxt0 - t0 > 1_000_000:
            continue
        def rel(a):
            return (a - t0) / 1e3
        pub = so[:, 3].max()
        rows.append([rel(so[live_s, 1].max()), rel(so[live_s, 2].max()), rel(pub) if pub > 0 else np.nan,
                     rel(f[live_f, 12].min()), rel(np.median(f[live_f, 13])), rel(f[live_f, 13].max()),
                     rel(np.median(f[live_f, 0])), rel(f[live_f, 0].max()),
                     rel(np.median(f[live_f, 3])), rel(np.median(f[live_f, 4])), rel(f[live_f, 4].max()),
                     rel(np.median(f[live_f, 15])), rel(f[live_f, 15].max()), rel(nxt0)])
    if not rows:
        print(path, "no usable steps")
        continue
    r = np.nanmean(np.array(rows), axis=0)
    names = ["soma: last CTA past griddep wait", "soma: last segment done", "soma: slice published",
             "fanout: first CTA entry", "fanout: past griddep wait (median)", "fanout: past griddep wait (last)",
             "fanout: raster visible (median)", "fanout: raster visible (last)",
             "fanout: item 1 list ready (median)", "fanout: item 1 streamed (median)", "fanout: item 1 streamed (last)",
             "fanout: CTA exit (median)", "fanout: CTA exit (last)", "next step's soma kernel starts"]
    print(path, f"({len(rows)} steps; us after the first neuron-phase CTA started)")
    for n, c in zip(names, r):
        print(f"  {n:>38}: {c:7.1f}")
