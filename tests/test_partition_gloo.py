"""N > 1 host logic on CPU: world_size-2 `gloo` group (tests/partition_worker.py)."""
import ctypes as C
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import helpers
import sanafe_b200 as sfe

HERE = os.path.dirname(os.path.abspath(__file__))


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("case", ["synth_delay", "example"])
def test_world2_raster_exchange_and_record_merge(case, tmp_path):
    out = tmp_path / "out.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(HERE, "partition_worker.py"), case, "12", str(out)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert res.returncode == 0, res.stderr[-2000:]
    r = json.loads(out.read_text())
    assert r["mismatches"] == 0
    assert r["spikes"] == r["ref_spikes"] and r["fired"] == r["ref_fired"]
    assert r["sim_time"] == pytest.approx(r["ref_sim_time"], rel=1e-12)
    assert r["energy"] == pytest.approx(r["ref_energy"], rel=1e-9)
    assert sorted(set(r["owner"])) in ([0, 1], [0])  # both ranks own cores (or a one-core chip)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_plan_partition_is_contiguous_balanced_and_disjoint(world):
    chip = helpers.load_chip("synth_small", device=-1)
    t = chip.tables
    owner = np.zeros(t.n_cores, dtype=np.uint32)
    begin = np.zeros(t.n_cores, dtype=np.uint32)
    sw = C.c_uint32()
    assert sfe.lib().sfe_plan_partition(C.byref(t), world, owner.ctypes.data, begin.ctypes.data, C.byref(sw)) == 0
    assert (np.diff(owner.astype(np.int64)) >= 0).all() and owner.max() < world
    # raster words of different cores never overlap and stay inside their owner's slice
    spans = []
    for c in range(t.n_cores):
        nw = (t.cores[c].neuron_count + 31) // 32
        if nw:
            assert owner[c] * sw.value <= begin[c] and begin[c] + nw <= (owner[c] + 1) * sw.value
            spans.append((int(begin[c]), int(begin[c]) + nw))
    spans.sort()
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))
    # balance: no rank holds more than twice the mean synapse work (cores are indivisible)
    work = np.zeros(world)
    for c in range(t.n_cores):
        work[owner[c]] += t.cores[c].syn_count + 64.0 * t.cores[c].neuron_count
    if world <= t.n_cores // 2:
        assert work.max() <= 2.0 * work.sum() / world
