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


def test_scheduler_threads_do_not_change_results():
    """Timesteps are scheduled independently (the reference's scheduler threads, `-S`): any number of host threads
    gives bit-identical per-step sim_time."""
    chip = load_chip("synth_delay", device=-1)
    steps = 80
    rd, out = Oracle(chip).run(steps, status=True)
    status = np.ascontiguousarray(out["status"])
    results = []
    for threads in (1, 3, 8, 0):
        assert sfe.lib().sfe_chip_set_scheduler_threads(chip._h, threads) == 0
        sim_time = np.zeros(steps)
        assert sfe.lib().sfe_chip_schedule_detailed(chip._h, status.ctypes.data, steps, sim_time.ctypes.data) == 0
        results.append(sim_time)
    assert all(np.array_equal(results[0], r) for r in results[1:])
    assert rel_err(results[0], golden("synth_delay")["detailed"]["per_step_sim_time"][:steps]) <= 1e-9
