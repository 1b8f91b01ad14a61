"""The expectations of the reference's own per-model unit tests (tests/unit/test_loihi_lif.cpp,
test_truenorth.cpp, test_accumulator.cpp, test_current_based_synapse.cpp, test_inputmodel.cpp), restated on the
one-neuron rig (unit_rig.py). `check_all(tmp_path, device, runner)` runs them on the CPU restatement
(test_reference_unit_vectors.py) or on the device (test_zz_new_models_gpu.py). Update call k of a reference test is
timestep k+2 here (see unit_rig); step index in the lists below = timestep - 1."""
import os

import pytest

import sanafe_b200 as sfe
import unit_rig as rig

LIF_BASE = dict(threshold=64.0, reset=0.0, reset_mode="hard", leak_decay=1.0, input_decay=0.0, bias=0.0)


def lif(tmp_path, device, runner, attrs, currents, steps=None, arch=None):
    chip = rig.load(tmp_path, arch or rig.demo_arch(), rig.lif_network(attrs, currents), device)
    return rig.run(chip, steps or len(currents) + 1, runner)


def tn(tmp_path, device, runner, attrs, currents, steps=None):
    chip = rig.load(tmp_path, rig.truenorth_arch(), rig.truenorth_network(attrs, currents), device)
    return rig.run(chip, steps or len(currents) + 1, runner)


def noisy_arch(tmp_path, entries, name):
    """The demo chip with a noise file on its default LIF soma (set_attribute_hw "noise", src/models.cpp:351-366)."""
    path = os.path.join(str(tmp_path), name)
    with open(path, "w") as f:
        f.write(entries)
    text = open(rig.demo_arch()).read()
    return text.replace("{name: demo_soma_default, attributes: {model: leaky_integrate_fire,",
                        "{name: demo_soma_default, attributes: {model: leaky_integrate_fire, noise: " + path + ",")


def check_lif(tmp_path, device, runner):
    # FiresWhenAboveThreshold (test_loihi_lif.cpp:28-45): 80 > 64 -> fired, hard reset to 0
    st, v = lif(tmp_path, device, runner, LIF_BASE, [80.0])
    assert st == ["idle", "fired"] and v[1] == 0.0
    # DoesNotFireBelowThreshold (:47-64) and StableWithoutInput (:66-86): updated, potential stays 50
    st, v = lif(tmp_path, device, runner, LIF_BASE, [50.0, None])
    assert st == ["idle", "updated", "updated"] and v[1] == 50.0 and v[2] == 50.0
    # LeakAndQuantizeReducesPotential (:113-126)
    st, v = lif(tmp_path, device, runner, dict(leak_decay=0.5, threshold=100.0), [80.0, None])
    assert v[1] == 80.0 and v[2] == 40.0 and v[2] < v[1]
    # FiresWithSoftReset (:128-139): 25 - 20 = 5 > 0
    st, v = lif(tmp_path, device, runner, dict(threshold=20.0, reset_mode="soft", reset=5.0), [25.0])
    assert st[1] == "fired" and v[1] == 5.0
    # ReverseThresholdBranches (:141-161), one neuron per mode: soft subtracts the (zero) reverse threshold,
    # hard goes to reverse_reset (default 0), saturate clamps at the reverse threshold
    for mode, want in (("soft", -10.0), ("hard", 0.0), ("saturate", 0.0)):
        st, v = lif(tmp_path, device, runner,
                    dict(threshold=100.0, reverse_threshold=0.0, reset=0.0, reverse_reset_mode=mode), [-10.0])
        assert st[1] == "updated" and v[1] == want, mode
    # AddsInputCurrentWhenProvided (:220-227), SetForceSomaUpdate (:318-327)
    st, v = lif(tmp_path, device, runner, dict(threshold=100.0), [2.0])
    assert v[1] == 2.0
    st, v = lif(tmp_path, device, runner, dict(force_update="true"), [], steps=2)
    assert st == ["updated", "updated"]
    # GenerateNoiseFromFile (:163-180): "10", an unparsable line (0), "20"; then the stream rewinds
    st, v = lif(tmp_path, device, runner, dict(threshold=100.0), [10.0], steps=5,
                arch=noisy_arch(tmp_path, "10\ninvalid\n20\n", "noise_test.txt"))
    assert v == [10.0, 20.0, 40.0, 50.0, 50.0]
    # NoiseGeneratesSignBit (:270-283): 256 has bit 8 set -> (256 & 0x7f) | ~0x7f = -128
    st, v = lif(tmp_path, device, runner, dict(threshold=10.0), [1.0],
                arch=noisy_arch(tmp_path, "256\n", "noise_signbit.txt"))
    assert v == [-128.0, -255.0]
    # NoiseEOFTriggersReset (:285-299): a one-line file is read again and again
    st, v = lif(tmp_path, device, runner, dict(threshold=100.0), [None, None],
                arch=noisy_arch(tmp_path, "5\n", "noise_eof.txt"))
    assert v == [5.0, 10.0, 15.0]
    # NoiseFileFailsToOpen (:88-94) / NoiseFileEmptyThrows (:254-268): raised when the chip is loaded
    with pytest.raises(sfe.SanafeError, match="Failed to open noise stream"):
        lif(tmp_path, device, runner, dict(threshold=10.0), [1.0],
            arch=open(rig.demo_arch()).read().replace("{model: leaky_integrate_fire, energy_access_neuron: 20.0e-12",
                                                      "{model: leaky_integrate_fire, noise: nonexistent.txt, energy_access_neuron: 20.0e-12"))
    with pytest.raises(sfe.SanafeError, match="Couldn't read noise entry"):
        lif(tmp_path, device, runner, dict(threshold=10.0), [5.0], arch=noisy_arch(tmp_path, "", "noise_empty.txt"))


def check_truenorth(tmp_path, device, runner):
    # SetThresholdAndUpdateFires (test_truenorth.cpp:38-45)
    st, v = tn(tmp_path, device, runner, dict(threshold=0.5, reset_mode="hard", reset=0.0), [1.0])
    assert st[1] == "fired" and v[1] == 0.0
    # LeakReducesPotential (:46-56): leak 0.5 towards zero
    st, v = tn(tmp_path, device, runner, dict(threshold=10.0, leak=0.5, leak_towards_zero="true"), [2.0, None])
    assert v[1] == 2.0 and v[2] == 1.5
    # LeakTowardsZeroBothDirections (:76-93)
    st, v = tn(tmp_path, device, runner, dict(threshold=10.0, leak=1.0, leak_towards_zero="true"), [3.0, None])
    assert v[1] == 3.0 and v[2] == 2.0
    st, v = tn(tmp_path, device, runner, dict(threshold=10.0, leak=1.0, leak_towards_zero="true"), [-3.0, None])
    assert v[1] == -3.0 and v[2] == -2.0
    # LeakWithoutTowardsZeroIncreasesPotential (:94-103)
    st, v = tn(tmp_path, device, runner, dict(threshold=10.0, leak=1.0, leak_towards_zero="false"), [], steps=3)
    assert v == [1.0, 2.0, 3.0]
    # ThresholdAndResetModes (:104-114): soft subtracts the threshold, saturate clamps at it
    st, v = tn(tmp_path, device, runner, dict(threshold=1.0, reset=0.0, reset_mode="soft"), [2.0])
    assert st[1] == "fired" and v[1] == 1.0
    st, v = tn(tmp_path, device, runner, dict(threshold=1.0, reset=0.0, reset_mode="saturate"), [2.0])
    assert st[1] == "fired" and v[1] == 1.0
    # ReverseResetModes (:115-129)
    for mode, want in (("hard", -2.0), ("soft", -5.0), ("saturate", 0.0)):
        st, v = tn(tmp_path, device, runner,
                   dict(threshold=10.0, reverse_threshold=0.0, reverse_reset=-2.0, reverse_reset_mode=mode), [-5.0])
        assert v[1] == want, mode
    # RandomizedThresholdAffectsPotential (:130-137): no mask -> plain hard reset
    st, v = tn(tmp_path, device, runner, dict(threshold=5.0, reset_mode="hard", reset=0.0), [10.0])
    assert st[1] == "fired" and v[1] == 0.0
    # RandomMaskNegativeThrows (:168-173)
    with pytest.raises(sfe.SanafeError, match="random_mask < 0"):
        tn(tmp_path, device, runner, dict(threshold=1.0, random_mask=-1), [])
    # RandomMaskEnablesRandomizedThreshold (:175-186): srand(1), mask 0xFF, no input -> the first rand() value
    # (1804289383 & 0xFF = 103) lifts the compared potential over the threshold: fired, hard reset to 0
    st, v = tn(tmp_path, device, runner, dict(threshold=1.0, reset_mode="hard", reset=0.0, random_mask=255), [], steps=1)
    assert st[0] == "fired" and v[0] == 0.0
    # (the jitter shifts what is compared, not the potential: with a threshold out of reach nothing accumulates)
    st, v = tn(tmp_path, device, runner, dict(threshold=1000.0, random_mask=255), [], steps=3)
    assert st == ["idle", "idle", "idle"] and v == [0.0, 0.0, 0.0]


def check_synapse_and_dendrite(tmp_path, device, runner):
    # CurrentBasedSynapseModel ReadReturnsWeight (test_current_based_synapse.cpp): the weight is the current
    st, v = lif(tmp_path, device, runner, dict(threshold=100.0), [1.23])
    assert v[1] == 1.23
    # AccumulatorModel AccumulatesChargeOverTime (test_accumulator.cpp:22-39): 2 + 3 arriving in one step -> 5
    net = rig.lif_network(dict(threshold=100.0), [2.0, 3.0]).replace("spikes: [0, 1]", "spikes: [1]")
    chip = rig.load(tmp_path, rig.demo_arch(), net, device)
    st, v = rig.run(chip, 2, runner)
    assert v[1] == 5.0


def check_input(tmp_path, device, runner):
    def input_net(attr):
        return ("network:\n  name: in\n  groups:\n"
                f"  - {{name: target, attributes: {{log_spikes: true, log_potential: true}}, neurons: [{{0: {attr}}}]}}\n"
                "  - {name: sink, attributes: {threshold: 1000000.0}, neurons: [{0: {}}]}\n"
                "  edges:\n  - {target.0 -> sink.0: {weight: 0.125}}\n"
                "mappings:\n- {target.0: {core: '0.0', soma: demo_input}}\n- {sink.0: {core: '0.1', soma: demo_soma_default}}\n")
    # GeneratesSpikeWhenSpikeValueSet / NoSpikeWhenSpikeValueZero (test_inputmodel.cpp:30-42); the train ends -> idle
    st, _ = rig.run(rig.load(tmp_path, rig.demo_arch(), input_net("{spikes: [1]}"), device), 3, runner)
    assert st == ["fired", "idle", "idle"]
    st, _ = rig.run(rig.load(tmp_path, rig.demo_arch(), input_net("{spikes: [0]}"), device), 2, runner)
    assert st == ["idle", "idle"]
    # GeneratesSpikeWithPoisson (:70-78): probability 1.0 > U in [0, 1) always; GeneratesSpikeWithRate (:79-87)
    st, _ = rig.run(rig.load(tmp_path, rig.demo_arch(), input_net("{poisson: 1.0}"), device), 4, runner)
    assert st == ["fired"] * 4
    st, _ = rig.run(rig.load(tmp_path, rig.demo_arch(), input_net("{rate: 1.0}"), device), 4, runner)
    assert st == ["fired"] * 4
    # ExternalCurrentThrows (:51-54): refused when the chip is loaded
    bad = rig.lif_network({}, [3.5]).replace("soma: demo_soma_default", "soma: demo_input")
    with pytest.raises(sfe.SanafeError, match="Current sent to input neuron"):
        rig.load(tmp_path, rig.demo_arch(), bad, device)
    # ModelParseResetMode / ModelGetPipelineUnit (:88-128): unknown names are invalid arguments
    with pytest.raises(sfe.SanafeError, match="Reset mode not recognized"):
        lif(tmp_path, device, runner, dict(reset_mode="invalid"), [])
    with pytest.raises(sfe.SanafeError, match="Pipeline model not supported"):
        lif(tmp_path, device, runner, {}, [],
            arch=open(rig.demo_arch()).read().replace("model: accumulator", "model: invalid_model"))


def taps_arch():
    return open(rig.demo_arch()).read().replace("model: accumulator", "model: taps")


def check_taps(tmp_path, device, runner):
    """test_multitap.cpp — CPU restatement only: the device engine refuses `taps` dendrites for now."""
    line = "dendrite: {taps: 2, time_constants: [1.0, 1.0], space_constants: [0.0]}"
    # InputCurrentAdds (:86-96): the current lands on tap 0 and is what the soma receives
    st, v = lif(tmp_path, device, runner, {"threshold": 100.0, line.split(":")[0]: line.split(": ", 1)[1]}, [1.5], arch=taps_arch())
    assert v[1] == 1.5
    # CalculateNextStateChangesVoltages (:128-139): time constant 0.5 halves the line per step (a zero-weight event
    # at the next step makes the lazily updated line catch up and report tap 0)
    half = {"threshold": 100.0, "leak_decay": 0.0, "dendrite": "{taps: 2, time_constants: [0.5, 0.5], space_constants: [0.0]}"}
    st, v = lif(tmp_path, device, runner, half, [2.0, 0.0], arch=taps_arch())
    assert v[1] == 2.0 and v[2] == 1.0 and v[2] < v[1]
    # InputCurrentToMappedTap (:97-107): a current addressed to tap 1 reaches tap 0 only through the space constant
    net = rig.lif_network({"threshold": 100.0, "leak_decay": 0.0,
                           "dendrite": "{taps: 2, time_constants: [1.0, 1.0], space_constants: [0.25]}"}, [2.0, 0.0])
    net = net.replace("{weight: 2.0}", "{weight: 2.0, tap: 1}")
    chip = rig.load(tmp_path, taps_arch(), net, device)
    st, v = rig.run(chip, 3, runner)
    assert v[1] == 0.0 and v[2] == 0.5  # step 2: tap 1 holds 2.0; step 3: a quarter of it has moved to tap 0
    # TapsZeroThrows (:37-43), TimeConstantsTooFewThrows (:152-158), InvalidTapThrows (:108-118)
    with pytest.raises(sfe.SanafeError, match="Number of taps must be > 0"):
        lif(tmp_path, device, runner, {"dendrite": "{taps: 0}"}, [1.0], arch=taps_arch())
    with pytest.raises(sfe.SanafeError, match="time constants"):
        lif(tmp_path, device, runner, {"dendrite": "{taps: 3, time_constants: [0.9, 0.8]}"}, [1.0], arch=taps_arch())
    bad = rig.lif_network({"dendrite": "{taps: 1}"}, [1.0]).replace("{weight: 1.0}", "{weight: 1.0, tap: 5}")
    with pytest.raises(sfe.SanafeError, match="Tap should be >= 0 and less than taps"):
        rig.load(tmp_path, taps_arch(), bad, device)
    # TimeConstantsResizeLargerVector / SpaceConstantsResizeLargerVector (:62-85): longer lists are accepted
    st, v = lif(tmp_path, device, runner, {"threshold": 100.0, "dendrite": "{taps: 2, time_constants: [0.5, 0.5, 0.5], space_constants: [0.4, 0.4, 0.4]}"},
                [1.0], arch=taps_arch())
    assert v[1] == 1.0


def check_all(tmp_path, device, runner):
    check_lif(tmp_path, device, runner)
    check_truenorth(tmp_path, device, runner)
    check_synapse_and_dendrite(tmp_path, device, runner)
    check_input(tmp_path, device, runner)
