"""Accept / reject expectations of the reference's parser unit tests (tests/unit/test_yaml_arch.cpp,
tests/unit/test_yaml_snn.cpp; GoogleTest + RapidYAML are fetched dependencies, so they cannot run here), restated
against this repo's in-tree YAML front-end through the Python surface. Each case names the reference test it
restates. The reference throws YamlDescriptionParsingError / std::invalid_argument etc.; here every rejected
description must raise, with a message that says what is wrong."""
import os

import pytest

from helpers import GOLDEN


def m():
    from sanafe_b200 import sanafecpp_b200
    return sanafecpp_b200


TILE_ATTRS = ("{energy_north_hop: 1.0, latency_north_hop: 1.0, energy_east_hop: 1.0, latency_east_hop: 1.0, "
              "energy_south_hop: 1.0, latency_south_hop: 1.0, energy_west_hop: 1.0, latency_west_hop: 1.0}")
CORE_BODY = """
          attributes: {buffer_position: soma, max_neurons_supported: 10}
          axon_in:
            - name: axin
              attributes: {energy_message_in: 0.0, latency_message_in: 0.0}
          synapse:
            - name: syn
              attributes: {model: current_based, energy_process_spike: 1.0, latency_process_spike: 1.0}
          dendrite:
            - name: dend
              attributes: {model: accumulator, energy_update: 0.0, latency_update: 0.0, update_every_timestep: true}
          soma:
            - name: soma
              attributes: {model: leaky_integrate_fire, energy_access_neuron: 1.0, latency_access_neuron: 1.0,
                           energy_update_neuron: 1.0, latency_update_neuron: 1.0, energy_spike_out: 1.0, latency_spike_out: 1.0}
          axon_out:
            - name: axout
              attributes: {energy_message_out: 1.0, latency_message_out: 1.0}
"""


def arch_text(tile="tile0", core="core0", width=1, body=CORE_BODY, with_core=True, with_tile=True):
    text = f"architecture:\n  name: a\n  attributes: {{link_buffer_size: 1, width: {width}, height: 1}}\n"
    if with_tile:
        text += f"  tile:\n    - name: {tile}\n      attributes: {TILE_ATTRS}\n"
        if with_core:
            text += f"      core:\n        - name: {core}{body}"
    return text


def load_arch(tmp_path, text):
    p = tmp_path / "arch.yaml"
    p.write_text(text)
    return m().load_arch(str(p))


def load_net(tmp_path, text, arch=None):
    arch = arch or load_arch(tmp_path, arch_text(tile="tile[0..1]", width=2))
    p = tmp_path / "net.yaml"
    p.write_text(text)
    return m().load_net(str(p), arch), arch


# ---- test_yaml_arch.cpp ---------------------------------------------------------------------------------------
def test_parses_basic_architecture(tmp_path):
    """ParsesBasicArchitecture (:149-288), LoadArchFromFile_VerifiesNestedStructure (:580-595)"""
    arch = load_arch(tmp_path, arch_text())
    # unit names always carry an index, range notation or not (test_yaml_arch.cpp:226-230)
    assert len(arch.tiles) == 1 and arch.tiles[0].name == "tile0[0]"
    cores = arch.tiles[0].cores
    assert len(cores) == 1 and cores[0].name == "core0[0]" and cores[0].id == 0
    assert len(arch.cores()) == 1


def test_tile_and_core_range_notation(tmp_path):
    """ParsesTileRangeNotation (:290-368), ParsesCoreRangeNotation (:370-447)"""
    arch = load_arch(tmp_path, arch_text(tile="tile[0..2]", width=3))
    assert [t.name for t in arch.tiles] == ["tile[0]", "tile[1]", "tile[2]"]
    assert [t.id for t in arch.tiles] == [0, 1, 2]
    arch = load_arch(tmp_path, arch_text(core="core[0..3]"))
    assert [c.name for c in arch.cores()] == ["core[0]", "core[1]", "core[2]", "core[3]"]
    assert [c.offset_within_tile for c in arch.cores()] == [0, 1, 2, 3]


@pytest.mark.parametrize("case,text", [
    ("MissingTileSectionThrows (:449-468)", arch_text(with_tile=False)),
    ("MissingCoreSectionThrows (:470-500)", arch_text(with_core=False)),
    ("MissingSomaSectionThrows (:502-559)", arch_text(body=CORE_BODY[:CORE_BODY.index("          soma:")] + CORE_BODY[CORE_BODY.index("          axon_out:"):])),
    ("ParseAxonInAttributes_Invalid (:42-56)", arch_text(body=CORE_BODY.replace(", latency_message_in: 0.0", ""))),
    ("ParseAxonOutAttributes_Invalid (:77-91)", arch_text(body=CORE_BODY.replace(", latency_message_out: 1.0", ""))),
])
def test_rejected_architectures(tmp_path, case, text):
    with pytest.raises(Exception) as e:
        load_arch(tmp_path, text)
    assert str(e.value).strip(), case


def test_architecture_metrics_reach_the_lowered_tables(tmp_path):
    """ParseAxonInAttributes_Valid / ParseAxonOutAttributes_Valid (:23-40, 58-75), DescriptionParseTileMetricsYaml_Valid
    (:117-147), ParseProcessingUnitAttributesWithPlugin (:93-115): the parsed figures, read back from the tables a chip
    lowers them to (the reference tests read them from the parser's structs)."""
    import sanafe_b200 as sfe
    tile_attrs = ("{energy_north_hop: 1.0, latency_north_hop: 2.0, energy_east_hop: 3.0, latency_east_hop: 4.0, "
                  "energy_south_hop: 5.0, latency_south_hop: 6.0, energy_west_hop: 7.0, latency_west_hop: 8.0, log_energy: true}")
    body = CORE_BODY.replace("{energy_message_in: 0.0, latency_message_in: 0.0}", "{energy_message_in: 7.89, latency_message_in: 0.12}")
    body = body.replace("{energy_message_out: 1.0, latency_message_out: 1.0}", "{energy_message_out: 7.89, latency_message_out: 0.12}")
    arch = load_arch(tmp_path, arch_text(body=body).replace(TILE_ATTRS, tile_attrs))
    net, _ = load_net(tmp_path, "network:\n  name: n\n  groups:\n    - name: g\n      neurons: [0]\n  edges: []\nmappings:\n  - g: {core: 0.0}\n", arch)
    chip = m().SpikingChip(arch, device=-1)
    chip.load(net)
    t = sfe.lib().sfe_chip_tables(chip._handle).contents
    tile, core = t.tiles[0], t.cores[0]
    assert (tile.energy_north, tile.latency_north, tile.energy_east, tile.latency_east) == (1.0, 2.0, 3.0, 4.0)
    assert (tile.energy_south, tile.latency_south, tile.energy_west, tile.latency_west) == (5.0, 6.0, 7.0, 8.0)
    assert (core.energy_axon_in, core.latency_axon_in, core.energy_axon_out, core.latency_axon_out) == (7.89, 0.12, 7.89, 0.12)
    # a unit that names a plugin library: the path is kept and asked for a device model when a neuron is mapped to the unit
    plug = arch_text(body=CORE_BODY.replace("{model: leaky_integrate_fire,", '{model: "testmodel", log_energy: true, log_latency: false, plugin: "plugin.so",'))
    arch2 = load_arch(tmp_path, plug)
    net2, _ = load_net(tmp_path, "network:\n  name: n\n  groups:\n    - name: g\n      neurons: [0]\n  edges: []\nmappings:\n  - g: {core: 0.0}\n", arch2)
    with pytest.raises(Exception, match=r"testmodel.*plugin\.so"):
        m().SpikingChip(arch2, device=-1).load(net2)


def test_arch_file_not_open():
    """LoadArchFromFile_FileNotOpen (:561-566)"""
    with pytest.raises(Exception):
        m().load_arch("/nonexistent/arch.yaml")


# ---- test_yaml_snn.cpp ----------------------------------------------------------------------------------------
GROUPS = """network:
  name: example
  groups:
    - name: Input
      neurons:
        - 0..1
    - name: Output
      neurons:
        - 0..1
"""
MAPPINGS = """mappings:
  - Input: {core: 0.0}
  - Output: {core: 1.0}
"""


def test_parse_full_network_section(tmp_path):
    """ParseFullNetworkSection (:187-229), ParseEdgeDescription_Valid / _WithWhitespace / _ExtremeWhitespace
    (:23-40, 376-393, 61-78), CountNeurons_WithRangesAndSingles (:80-96)"""
    net, _ = load_net(tmp_path, GROUPS + "  edges:\n    - Input.0 -> Output.0: [weight: -1.0]\n"
                                         "    -    Input.1   ->   Output.1  : [weight: -2.0]\n" + MAPPINGS)
    assert sorted(net.groups) == ["Input", "Output"] and len(net["Input"]) == 2 and len(net["Output"]) == 2
    e0, e1 = net["Input"][0].edges_out, net["Input"][1].edges_out
    assert len(e0) == 1 and len(e1) == 1
    assert (e0[0].post_neuron.group_name, e0[0].post_neuron.neuron_offset) == ("Output", 0)
    assert (e1[0].post_neuron.group_name, e1[0].post_neuron.neuron_offset) == ("Output", 1)
    assert e0[0].synapse_attributes == {"weight": -1.0} and e1[0].synapse_attributes == {"weight": -2.0}
    net, _ = load_net(tmp_path, "network:\n  name: n\n  groups:\n    - name: g\n      neurons: [0..2, 3, 4..9]\n"
                                "  edges: []\nmappings:\n  - g: {core: 0.0}\n")
    assert len(net["g"]) == 10


@pytest.mark.parametrize("style", ["- {0: [log_spikes: true, threshold: 2.5]}", "- {0: {log_spikes: true, threshold: 2.5}}",
                                   "- 0: [log_spikes: true, threshold: 2.5]", "- 0: {log_spikes: true, threshold: 2.5}"])
def test_neuron_attribute_styles(tmp_path, style):
    """ParseNeuronSimAttributesListOfMapsFlow / MapFlow / ListOfMapsInline / MapInline (:113-185)"""
    import sanafe_b200 as sfe
    net, arch = load_net(tmp_path, "network:\n  name: n\n  groups:\n    - name: g\n      neurons:\n        " + style +
                         "\n  edges: []\nmappings:\n  - g: {core: 0.0}\n")
    chip = m().SpikingChip(arch, device=-1)
    chip.load(net)
    t = sfe.lib().sfe_chip_tables(chip._handle).contents
    assert t.soma_classes[t.neuron_class[0]].threshold == 2.5
    assert sfe.lib().sfe_chip_format_spikes(chip._handle, (sfe.C.c_uint32 * 1)(1), 1, 1, None, 0) == len("g.0,1\n")


def test_unit_specific_attributes(tmp_path):
    """ParseNeuronAttributes_UnitSpecificModelAttributes (:443-476), ParseEdgeAttributes_UnitSpecific (:640-669),
    ParseNeuronAttributes_HardwareUnits (:425-441), ParseMappingInfo_AllHardwareUnits (:1044-1079)"""
    import sanafe_b200 as sfe
    text = ("network:\n  name: n\n  groups:\n    - name: Input\n      neurons:\n        - 0: {soma: {threshold: 7.0}, dendrite: {threshold: 9.0}}\n"
            "    - name: Output\n      neurons:\n        - 0\n  edges:\n    - Input.0 -> Output.0:\n        synapse:\n          weight: 1.5\n"
            "        dendrite:\n          weight: 99\n" + "mappings:\n  - Input.0: {core: 0.0, soma: soma, dendrite: dend, synapse: syn}\n  - Output.0: {core: 0.0}\n")
    net, arch = load_net(tmp_path, text)
    assert net["Input"][0].edges_out[0].synapse_attributes == {"weight": 1.5}
    chip = m().SpikingChip(arch, device=-1)
    chip.load(net)
    t = sfe.lib().sfe_chip_tables(chip._handle).contents
    i = sfe.lib().sfe_chip_neuron_index(chip._handle, b"Input", 0)
    assert t.soma_classes[t.neuron_class[i]].threshold == 7.0  # the soma-specific value, not the dendrite's
    assert t.syn_weight[0] == 1.5


REJECTED_NETS = [
    ("ParseEdgeDescription_MissingDotThrows (:42-59)", GROUPS + "  edges:\n    - Input0 -> Output.0: [weight: 1]\n" + MAPPINGS),
    ("ParseEdgeDescription_NoArrowThrows (:341-355)", GROUPS + "  edges:\n    - Input.0 Output.0: [weight: 1]\n" + MAPPINGS),
    ("CountNeurons_InvalidFormatThrows (:98-111)", GROUPS.replace("- 0..1\n    - name: Output", "- 0..x\n    - name: Output") + "  edges: []\n" + MAPPINGS),
    ("ParseNetworkSection_MissingGroupsThrows (:498-511)", "network:\n  name: n\n  edges: []\n" + MAPPINGS),
    ("ParseNetworkSection_MissingEdgesThrows (:513-529)", GROUPS + MAPPINGS),
    ("ParseNeuronConnection_InvalidSourceGroup (:531-549)", GROUPS + "  edges:\n    - Nope.0 -> Output.0: [weight: 1]\n" + MAPPINGS),
    ("ParseNeuronConnection_InvalidTargetGroup (:551-569)", GROUPS + "  edges:\n    - Input.0 -> Nope.0: [weight: 1]\n" + MAPPINGS),
    ("ParseNeuronConnection_OutOfBoundsNeuronId (:571-592)", GROUPS + "  edges:\n    - Input.7 -> Output.0: [weight: 1]\n" + MAPPINGS),
    ("ParseHyperedge_NoTypeThrows (:594-615)", GROUPS + "  edges:\n    - Input -> Output: [weight: [1, 2, 3, 4]]\n" + MAPPINGS),
    ("ParseHyperedge_InvalidTypeThrows (:617-638)", GROUPS + "  edges:\n    - Input -> Output: [type: bogus, weight: [1, 2, 3, 4]]\n" + MAPPINGS),
    ("ParseDenseHyperedge_NonListAttributeThrows (:805-828)", GROUPS + "  edges:\n    - Input -> Output: [type: dense, weight: 1.0]\n" + MAPPINGS),
    ("ParseSparseHyperedge_NonListPairsThrows (:855-878)", GROUPS + "  edges:\n    - Input -> Output: [type: sparse, source_target_pairs: 3, weight: [1]]\n" + MAPPINGS),
    ("ParseSparseHyperedge_InvalidPairFormat (:830-853)", GROUPS + "  edges:\n    - Input -> Output: [type: sparse, source_target_pairs: [[0, 1, 1]], weight: [1]]\n" + MAPPINGS),
    ("ParseMappingSection_InvalidNeuronGroup (:671-696)", GROUPS + "  edges: []\nmappings:\n  - Nope: {core: 0.0}\n"),
    ("ParseMappingSection_OutOfBoundsTile (:698-723)", GROUPS + "  edges: []\nmappings:\n  - Input: {core: 9.0}\n  - Output: {core: 0.0}\n"),
    ("ParseMappingSection_NotSequenceThrows (:1081-1108)", GROUPS + "  edges: []\nmappings:\n  Input: {core: 0.0}\n"),
    ("ParseMapping_MultipleEntriesThrows (:1110-1138)", GROUPS + "  edges: []\nmappings:\n  - {Input: {core: 0.0}, Output: {core: 0.0}}\n"),
    ("ParseEdgesSection_NotSequenceThrows (:1140-1158)", GROUPS + "  edges: {Input.0 -> Output.0: [weight: 1]}\n" + MAPPINGS),
    ("ParseNeuronSection_NotSequenceThrows (:1160-1177)", GROUPS.replace("      neurons:\n        - 0..1\n    - name: Output", "      neurons: {0..1: {}}\n    - name: Output") + "  edges: []\n" + MAPPINGS),
    ("ParseNeuronGroupSection_NotSequenceThrows (:1179-1195)", "network:\n  name: n\n  groups: {name: g}\n  edges: []\nmappings: []\n"),
    ("ParseNetworkFile_MissingNetworkSection (:915-939)", MAPPINGS),
    ("ParseNetworkFile_MissingMappingsSection (:941-969)", GROUPS + "  edges: []\n"),
    ("ParseNetworkFile_InvalidTopLevelFormat (:971-990)", "- just\n- a\n- list\n"),
    ("ParseNeuronGroup_NoNeuronsSection (:1008-1024)", "network:\n  name: n\n  groups:\n    - name: g\n  edges: []\nmappings: []\n"),
    ("Conv2D_WrongOutputNeuronCount (:1407-1438)", GROUPS + "  edges:\n    - Input -> Output: [type: conv2d, input_width: 2, input_height: 1, input_channels: 1, "
                                                              "kernel_width: 1, kernel_height: 1, kernel_count: 3, weight: [1, 2, 3]]\n" + MAPPINGS),
    ("Conv2D_WrongInputNeuronCount (:1440-1471)", GROUPS + "  edges:\n    - Input -> Output: [type: conv2d, input_width: 3, input_height: 1, input_channels: 1, "
                                                             "kernel_width: 2, kernel_height: 1, kernel_count: 1, weight: [1, 2]]\n" + MAPPINGS),
]


@pytest.mark.parametrize("case,text", REJECTED_NETS, ids=[c.split(" ")[0] for c, _ in REJECTED_NETS])
def test_rejected_networks(tmp_path, case, text):
    with pytest.raises(Exception) as e:
        load_net(tmp_path, text)
    assert str(e.value).strip(), case


def test_net_file_not_open(tmp_path):
    """ParseNetworkFile_FileNotOpen (:905-913)"""
    arch = load_arch(tmp_path, arch_text())
    with pytest.raises(Exception):
        m().load_net("/nonexistent/net.yaml", arch)


def test_mapping_ranges_and_whole_groups(tmp_path):
    """ParseMappingSection_NeuronRange (:725-755), ParseMapping_AllNeuronsInGroup (:1374-1405),
    ParseEdgeDescription_HyperedgeNoNeuronOffset (:357-374), ParseConv2dHyperedge_AllParameters (:773-803)"""
    import sanafe_b200 as sfe
    text = ("network:\n  name: n\n  groups:\n    - name: a\n      neurons: [0..8]\n    - name: b\n      neurons: [0..3]\n"
            "  edges:\n    - a -> b: [type: conv2d, input_width: 3, input_height: 3, input_channels: 1, kernel_width: 2, kernel_height: 2, "
            "kernel_count: 1, stride_width: 1, stride_height: 1, weight: [1, 2, 3, 4]]\n"
            "mappings:\n  - a.0..3: {core: 0.0}\n  - a.4..8: {core: 1.0}\n  - b: {core: 1.0}\n")
    net, arch = load_net(tmp_path, text)
    chip = m().SpikingChip(arch, device=-1)
    chip.load(net)
    t = sfe.lib().sfe_chip_tables(chip._handle).contents
    assert [t.cores[c].neuron_count for c in range(2)] == [4, 9] and t.n_synapses == 16


def test_serialize_network(tmp_path):
    """SerializeNetworkToYaml (:295-339), WriteEdgeFormat (:284-293), SerializeNeuronRuns_MultipleRuns (:1244-1289),
    WriteNetwork_EmptyNetworkName (:1197-1223), WriteMappings_NeuronNotMapped (:992-1006)"""
    mod = m()
    arch = load_arch(tmp_path, arch_text(tile="tile[0..1]", width=2))
    net = mod.Network()
    g = net.create_neuron_group("g", 6, model_attributes={"threshold": 2.0})
    for i in (2, 3):
        g[i].set_attributes(model_attributes={"bias": 1.5})
    g[0].connect_to_neuron(g[5], {"weight": 3})
    path = str(tmp_path / "out.yaml")
    with pytest.raises(RuntimeError, match="not mapped"):
        net.save(path)
    for n in g:
        n.map_to_core(arch.tiles[1].cores[0])
    net.save(path)
    text = open(path).read()
    assert 'name: " "' in text                                   # empty network name
    assert "{0..1: {}}" in text and "{2..3: {bias: 1.5}}" in text and "{4..5: {}}" in text  # three runs
    assert '"g.0 -> g.5": {weight: 3}' in text
    assert text.count('core: "1.0"') == 6
    again = mod.load_net(path, arch)
    assert len(again["g"]) == 6 and again["g"][0].edges_out[0].post_neuron.neuron_offset == 5


# ---- test_connect_neurons_sparse.cpp ----------------------------------------------------------------------------
def sparse_groups(n_src, n_dst):
    net = m().Network()
    return net, net.create_neuron_group("src", n_src), net.create_neuron_group("dst", n_dst)


def weight_of(connection):
    return connection.synapse_attributes["weight"]


def test_sparse_attributes_are_indexed_by_edge_position():
    """AttributesIndexedByEdgePositionNotSourceId (:40-70)"""
    net, src, dst = sparse_groups(3, 3)
    src.connect_neurons_sparse(dst, {"weight": [10.0, 20.0, 30.0]}, [(2, 0), (0, 1), (1, 2)])
    assert [weight_of(src[i].edges_out[0]) for i in (2, 0, 1)] == [10.0, 20.0, 30.0]


def test_sparse_edges_from_one_source_get_distinct_attributes():
    """MultipleEdgesFromSameSourceGetDistinctAttributes (:75-100)"""
    net, src, dst = sparse_groups(2, 3)
    src.connect_neurons_sparse(dst, {"weight": [1.0, 2.0, 3.0]}, [(0, 0), (0, 1), (1, 2)])
    assert [weight_of(e) for e in src[0].edges_out] == [1.0, 2.0] and weight_of(src[1].edges_out[0]) == 3.0


def test_sparse_large_source_id_small_edge_count():
    """LargeSourceIdSmallEdgeCountDoesNotOverrun (:103-120)"""
    net, src, dst = sparse_groups(8, 2)
    src.connect_neurons_sparse(dst, {"weight": [100.0, 200.0]}, [(5, 0), (7, 1)])
    assert weight_of(src[5].edges_out[0]) == 100.0 and weight_of(src[7].edges_out[0]) == 200.0
    with pytest.raises(ValueError):
        src.connect_neurons_sparse(dst, {"weight": [1.0]}, [(8, 0)])  # source id out of range


# ---- test_basic_input.cpp (the command line's required arguments) -------------------------------------------------
def run_sim(*args):
    import subprocess
    from helpers import ROOT
    sim = os.path.join(ROOT, "sana-fe_b200", "sanafe_b200", "sim")
    return subprocess.run([sim, *args], capture_output=True, text=True, timeout=60)


@pytest.mark.parametrize("case,args,needle", [
    ("MissingArguments (:23-27)", ["arch.yaml"], "Usage: ./sim"),
    ("InvalidTimestepNonNumeric (:29-33)", ["arch.yaml", "net.yaml", "abc"], "Invalid time-step format"),
    ("InvalidTimestepNegative (:35-39)", ["arch.yaml", "net.yaml", "-10"], "Time-steps must be > 0"),
    ("InvalidTimestepZero (:41-45)", ["arch.yaml", "net.yaml", "0"], "Time-steps must be > 0"),
    ("FileDoesNotExist (:47-55)", ["nonexistent_arch.yaml", "net.yaml", "100"], "nonexistent_arch.yaml"),
])
def test_command_line_required_arguments(case, args, needle):
    res = run_sim(*args)
    assert needle in (res.stdout + res.stderr), (case, res.stdout, res.stderr)
    if "Usage" not in needle:
        assert res.returncode != 0, case


# ---- more of test_yaml_snn.cpp (restated in round 2) -------------------------------------------------------------
MORE_REJECTED_NETS = [
    ("ParseMultipleNetworks (:258-282)", GROUPS.replace("name: example", "name: example[0..2]") + "  edges: []\n" + MAPPINGS),
    ("ParseNeuronSection_InvalidNeuronId (:478-496)", "network:\n  name: test\n  groups:\n    - name: Input\n      neurons:\n        - 0..1\n"
                                                      "        - 5: {weight: 1.0}\n  edges: []\nmappings:\n  - Input: {core: 0.0}\n"),
    ("ParseSparseHyperedge_InvalidPairTypeThrows (:880-903)", GROUPS + "  edges:\n    - Input -> Output:\n        type: sparse\n"
                                                              "        source_target_pairs: [0]\n" + MAPPINGS),
]


@pytest.mark.parametrize("case,text", MORE_REJECTED_NETS, ids=[c.split(" ")[0] for c, _ in MORE_REJECTED_NETS])
def test_more_rejected_networks(tmp_path, case, text):
    with pytest.raises(Exception) as e:
        load_net(tmp_path, text)
    assert str(e.value).strip(), case


def test_more_accepted_networks(tmp_path):
    """ParseNeuronGroup_EmptyName (:1026-1042), ParseHyperedgeType_FromSequence (:757-771: the attributes of a hyper-edge as a
    sequence of one-key maps), ParseNeuronAttributes_HardwareUnits (:425-441) and ParseMappingInfo_AllHardwareUnits
    (:1044-1079: unit names given with the mapping)."""
    net, _ = load_net(tmp_path, 'network:\n  name: test\n  groups:\n    - name: ""\n      neurons:\n        - 0\n  edges: []\nmappings: []\n')
    assert list(net.groups) == [""] and len(net[""]) == 1
    net, _ = load_net(tmp_path, GROUPS + "  edges:\n    - Input -> Output:\n        - type: dense\n        - weight: [1.0, 2.0, 3.0, 4.0]\n" + MAPPINGS)
    assert [c.synapse_attributes["weight"] for n in net["Input"] for c in n.edges_out] == [1.0, 2.0, 3.0, 4.0]
    text = ("network:\n  name: test\n  groups:\n    - name: Input\n      attributes: {synapse_hw_name: syn, dendrite_hw_name: dend, soma_hw_name: soma}\n"
            "      neurons:\n        - 0\n    - name: Other\n      neurons:\n        - 0\n  edges: []\n"
            "mappings:\n  - Input.0: {core: 0.0}\n  - Other.0:\n      core: 0.0\n      synapse: syn\n      dendrite: dend\n      soma: soma\n")
    net, arch = load_net(tmp_path, text)
    chip = m().SpikingChip(arch, device=-1)
    chip.load(net)  # the named units exist on the core: both ways of naming them resolve
    # ... and the names are really used: a unit the core does not have is refused when the neuron is mapped
    for bad in (text.replace("soma_hw_name: soma}", "soma_hw_name: nope}"), text.replace("      soma: soma\n", "      soma: nope\n")):
        assert bad != text
        net, arch = load_net(tmp_path, bad)
        with pytest.raises(Exception, match="nope"):
            m().SpikingChip(arch, device=-1).load(net)


def test_save_into_an_existing_file(tmp_path):
    """WriteNetwork_PreservesOtherSections (:1291-1332), WriteMappings_PreservesNetworkSection (:1334-1372),
    WriteNetwork_ExistingFileWithInvalidYAML (:1225-1242): saving into an existing file replaces its `network` and `mappings`
    sections, keeps the others, and refuses a file that is not YAML."""
    mod = m()
    arch = load_arch(tmp_path, arch_text())
    net = mod.Network()  # (the Python class has no name argument, src/pymodule.cpp:1120-1122: an unnamed network is written as " ")
    g = net.create_neuron_group("TestGroup", 1)
    g[0].map_to_core(arch.tiles[0].cores[0])
    path = tmp_path / "preserve.yaml"
    path.write_text("\ncustom_section:\n  data: should_be_preserved\nnetwork:\n  name: old\n  groups: []\n  edges: []\n")
    net.save(str(path))
    text = path.read_text()
    assert "custom_section" in text and "should_be_preserved" in text and "TestGroup" in text and "name: old" not in text
    again = mod.load_net(str(path), arch)
    assert list(again.groups) == ["TestGroup"]
    bad = tmp_path / "invalid.yaml"
    bad.write_text("this is not valid: yaml: content\n[[[")
    with pytest.raises(RuntimeError):
        net.save(str(bad))
