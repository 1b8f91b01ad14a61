"""Pins the CPU restatement (oracle/sfe_oracle.c) — and with it the new engine's
load-time lowering, which it shares — against outputs of the reference itself
(tests/golden/, produced by oracle/_ref/sanafe_ref = the unmodified reference
engine). Runs without a GPU."""
import pytest

from helpers import GOLDEN_CASES, NEW_GOLDEN_CASES, ORACLE_ONLY_CASES, Oracle, check_against_golden, golden, load_chip


@pytest.mark.parametrize("name", GOLDEN_CASES + NEW_GOLDEN_CASES + ORACLE_ONLY_CASES)
def test_restatement_matches_reference(name):
    chip = load_chip(name, device=-1)
    g = golden(name)
    t = chip.tables
    assert t.n_neurons == g["summary"]["neurons"]
    assert t.n_synapses == g["summary"]["synapses"]
    assert t.mapped_cores == g["summary"]["mapped_cores"]
    assert t.mapped_tiles == g["summary"]["mapped_tiles"]
    rd, out = Oracle(chip).run(g["steps"])
    # same glibc exp/pow as the reference plugin: HH potentials are bit-exact too
    check_against_golden(name, chip, rd, out, potential_rtol=0.0, energy_rtol=1e-12)
