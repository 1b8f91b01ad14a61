"""Detailed timing model (reference src/schedule.cpp:208-620): the host scheduler is
fed per-step status bytes and must reproduce the reference's per-step sim_time.
CPU test: status bytes from the CPU restatement; GPU test: through sfe_chip_sim."""
import numpy as np
import pytest

import sanafe_b200 as sfe
from helpers import Oracle, golden, load_chip, rel_err

CASES = ["example", "frac", "truenorth", "synth_delay", "dvs"]


@pytest.mark.parametrize("name", CASES)
def test_host_scheduler_matches_reference(name):
    g = golden(name)
    steps = g["steps"] if name != "dvs" else 200
    chip = load_chip(name, device=-1)
    rd, out = Oracle(chip).run(steps, status=True)
    sim_time = np.zeros(steps)
    status = np.ascontiguousarray(out["status"])
    rc = sfe.lib().sfe_chip_schedule_detailed(chip._h, status.ctypes.data, steps, sim_time.ctypes.data)
    assert rc == 0, sfe.lib().sfe_last_error()
    want = np.asarray(g["detailed"]["per_step_sim_time"][:steps])
    assert rel_err(sim_time, want) <= 1e-9, (name, rel_err(sim_time, want))


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_engine_detailed_timing(name):
    g = golden(name)
    chip = load_chip(name, device=0)
    rd, out = chip.sim_raw(g["steps"], "detailed", steps=True, fired=True)
    want = np.asarray(g["detailed"]["per_step_sim_time"])
    assert rel_err(out["steps"]["sim_time"], want) <= 1e-9
    assert rel_err(rd.sim_time, g["detailed"]["sim_time"]) <= 1e-9
    assert rd.neurons_fired == g["summary"]["neurons_fired"] and rd.spikes == g["summary"]["spikes"]
    assert rd.scheduler_wall_time > 0.0
