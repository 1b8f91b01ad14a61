"""The reference's per-model unit-test expectations (tests/unit/test_*.cpp) against the CPU restatement and the
lowering it shares with the engine. The same vectors run on the device in test_zz_new_models_gpu.py."""
import reference_unit_vectors as vectors
import unit_rig as rig


def poisson_aware_oracle(chip, steps):
    return rig.oracle_runner(chip, steps)


def test_lif_vectors(tmp_path):
    vectors.check_lif(tmp_path, -1, poisson_aware_oracle)


def test_truenorth_vectors(tmp_path):
    vectors.check_truenorth(tmp_path, -1, poisson_aware_oracle)


def test_synapse_and_dendrite_vectors(tmp_path):
    vectors.check_synapse_and_dendrite(tmp_path, -1, poisson_aware_oracle)


def test_input_vectors(tmp_path):
    vectors.check_input(tmp_path, -1, poisson_aware_oracle)


def test_multitap_vectors(tmp_path):
    vectors.check_taps(tmp_path, -1, poisson_aware_oracle)
