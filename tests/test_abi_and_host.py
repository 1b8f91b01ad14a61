"""CPU-side checks: the C-ABI library loads and exports every symbol the header
declares; host-only chips lower networks but refuse to simulate (no CPU fallback);
the generated loihi_large architecture equals the reference's arch/loihi_large.yaml."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import sanafe_b200 as sfe
from helpers import GOLDEN, REFERENCE_ROOT, ROOT, Oracle, golden_flat, load_chip


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "sanafe_b200.h")).read()
    # the host half of the device-model header (its CUDA half holds templates for plugin authors, not exports)
    text += open(os.path.join(ROOT, "include", "sfe_device_model.h")).read().split("#ifdef __CUDACC__")[0]
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(sfe_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(os.path.join(ROOT, "sana-fe_b200", "sanafe_b200", "libsanafe_b200.so"))
    names = declared_symbols()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sfe.lib().sfe_abi_version() == 7


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to create a device chip; a host-only
    chip (device=-1) lowers and exports tables but sim() fails loudly."""
    if sfe.lib().sfe_device_count() > 0:
        pytest.skip("a GPU is present")
    arch, net = sfe.load_flat(golden_flat("example"))
    with pytest.raises(sfe.SanafeError, match="no CUDA device"):
        sfe.SpikingChip(arch, device=0)
    chip = sfe.SpikingChip(arch, device=-1)
    chip.load(net)
    with pytest.raises(sfe.SanafeError, match="no CUDA device"):
        chip.sim_raw(1)


def test_product_does_not_link_the_oracle():
    lib = os.path.join(ROOT, "sana-fe_b200", "sanafe_b200", "libsanafe_b200.so")
    out = subprocess.run(["nm", "-D", lib], capture_output=True, text=True).stdout
    assert "sfe_oracle" not in out
    for root, _, files in os.walk(os.path.join(ROOT, "sana-fe_b200")):
        for f in files:
            if f.endswith((".cpp", ".hpp", ".cu", ".py", "Makefile")):
                src = open(os.path.join(root, f)).read()
                assert "oracle/" not in src.replace("oracle/_ref/libhodgkin_huxley.so", "").replace(
                    "oracle/yaml_to_flat.py", ""), f


def test_mapping_errors_mirror_the_reference():
    arch, net = sfe.load_flat(golden_flat("example"))
    chip = sfe.SpikingChip(arch, device=-1)
    chip.load(net)
    assert chip.neuron_index("in", 1) >= 0
    assert chip.neuron_index("nope", 0) == -1
    # synthetic network larger than the architecture
    spec = sfe.SynthSpec(cores=64, neurons_per_core=64, dest_cores=2, syn_per_axon=4, seed=1, bias_permille=100,
                         bias=1.0, threshold=1.0, reset=0.0, leak_decay=1.0, w_min=-1, w_max=1, max_delay=0,
                         log_spikes=0, log_potential_n=0)
    with pytest.raises(sfe.SanafeError):
        chip.load_synthetic(spec, generate_on_device=False)


def test_lowered_tables_invariants():
    chip = load_chip("dvs", device=-1)
    t = chip.tables
    begins = np.ctypeslib.as_array(t.axon_out_begin, shape=(t.n_neurons + 1,))
    assert np.all(np.diff(begins.astype(np.int64)) >= 0) and begins[-1] == t.n_axons_out
    targets = np.ctypeslib.as_array(t.axon_out_target, shape=(t.n_axons_out,))
    # one axon-in per (pre-neuron, destination core): every axon-in has exactly one sender
    assert t.n_axons_out == t.n_axons_in and len(np.unique(targets)) == t.n_axons_in
    total = 0
    for c in range(t.n_cores):
        cd = t.cores[c]
        total += cd.syn_count
        for a in range(cd.axon_in_begin, cd.axon_in_begin + min(cd.axon_in_count, 50)):
            ax = t.axons_in[a]
            assert ax.syn_off + ax.syn_count <= cd.syn_count
    assert total == t.n_synapses


@pytest.mark.skipif(not os.path.exists(REFERENCE_ROOT), reason="needs the reference tree (this container only)")
def test_generated_loihi_large_equals_reference_yaml(tmp_path):
    """sanafe_b200.archgen.loihi_large() is what the GPU box uses instead of
    arch/loihi_large.yaml; both must lower the synthetic network to identical tables."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import yaml_to_flat
    from sanafe_b200 import archgen

    synth = dict(cores=8, neurons_per_core=64, dest_cores=4, syn_per_axon=16, seed=1, bias_permille=100, bias=128.0,
                 threshold=64.0, reset=0.0, leak_decay=0.9, w_min=-8, w_max=8, max_delay=2, log_spikes=1,
                 log_potential_n=8, soma_hw_name="loihi_lif", synapse_hw_name="loihi_dense_synapse",
                 dendrite_hw_name="loihi_dendrites_delay")
    ref_flat = str(tmp_path / "ref.jsonl")
    gen_flat = str(tmp_path / "gen.jsonl")
    yaml_to_flat.convert(os.path.join(REFERENCE_ROOT, "arch", "loihi_large.yaml"), None, ref_flat, max_tiles=3, synth=synth)
    archgen.write_flat(archgen.loihi_large(tiles=3), gen_flat, synth=synth)
    results = []
    for path in (ref_flat, gen_flat):
        arch, net = sfe.load_flat(path)
        chip = sfe.SpikingChip(arch, device=-1)
        chip.load(net)
        rd, out = Oracle(chip).run(30)
        results.append((rd, out, chip))
    (rd_a, out_a, _), (rd_b, out_b, _) = results
    assert np.array_equal(out_a["fired_bits"], out_b["fired_bits"])
    assert np.array_equal(out_a["potentials"], out_b["potentials"])
    for key in out_a["steps"].dtype.names:
        assert np.array_equal(out_a["steps"][key], out_b["steps"][key]), key


def test_benchmark_sample_golden_matches_bench_py():
    """tests/golden/bench_sample32.hash.json (made by the reference's engine on BASELINE.md section 3's sample) describes
    the workload bench.py runs today: same generator parameters, and bench.raster_hash is the hash the harness computes."""
    import json
    import sys
    sys.path.insert(0, ROOT)
    import bench
    with open(os.path.join(GOLDEN, "bench_sample32.hash.json")) as f:
        g = json.load(f)
    assert g["spec"] == bench.sample_spec(g["sample_cores"])
    assert g["hash"]["hash_steps"] == bench.HASH_STEPS and len(g["hash"]["raster_hash"]) == 16
    # raster_hash on a tiny raster, by hand: timestep 1, neuron 3 and timestep 2, neuron 0
    bits = np.zeros((2, 1), dtype=np.uint32)
    bits[0, 0] = 1 << 3
    bits[1, 0] = 1
    want = (int(bench.mix64(np.array([(1 << 32) | 3], dtype=np.uint64))[0]) + int(bench.mix64(np.array([2 << 32], dtype=np.uint64))[0])) % (1 << 64)
    assert bench.raster_hash(bits) == want
