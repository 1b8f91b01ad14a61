#!/usr/bin/env python3
"""yaml_to_flat.py — PyYAML -> "flat" JSON-lines description for the oracle.

TEST INFRASTRUCTURE ONLY (see oracle/README.md). The reference parses its
Architecture / SNN YAML files with RapidYAML, which is fetched from the network
by its build and is absent here. This script is the independent front-end for
the ORACLE: it parses the same files with PyYAML and writes one record per line;
oracle/ref_harness.cpp replays the records through the reference's public
builder API. The new engine does NOT use this path (it has its own C++ YAML
reader), so a reader bug cannot cancel out between the two sides.

Dispatch rules restated from the reference (file:line):
  * scalar typing int -> double -> bool -> string   src/yaml_common.cpp:205-263
  * skip keys for model attributes                  src/yaml_common.cpp:30-36
  * arch sections, name[a..b] ranges, unit merging  src/yaml_arch.cpp:149-293,346-423
  * NoC / sync table                                src/yaml_arch.cpp:425-510
  * neuron attribute forwarding                     src/yaml_snn.cpp:331-394
  * edges / hyper-edges                             src/yaml_snn.cpp:396-878
  * mappings                                        src/yaml_snn.cpp:880-1056
"""
import argparse
import json
import re
import sys

import yaml

SKIP_KEYS = {"soma_hw_name", "default_synapse_hw_name", "dendrite_hw_name",
             "log_spikes", "log_potential", "synapse", "dendrite", "soma"}

_INT_RE = re.compile(r"^[+-]?(0x[0-9a-fA-F]+|0b[01]+|0o[0-7]+|[0-9]+)$")


def typed(text):
    """Type a YAML scalar the way src/yaml_common.cpp:205-263 does."""
    if not isinstance(text, str):
        raise ValueError(f"expected scalar, got {type(text)}")
    t = text.strip()
    if _INT_RE.match(t):
        return int(t, 0) if not t.lstrip("+-").isdigit() else int(t, 10)
    try:
        if t and t.lower() not in ("nan", "inf", "-inf", "+inf", "infinity"):
            return float(t)
    except ValueError:
        pass
    if t in ("true", "True", "TRUE"):
        return True
    if t in ("false", "False", "FALSE"):
        return False
    return text


def to_bool(text):
    v = typed(text)
    if isinstance(v, bool):
        return v
    if isinstance(v, int):
        return v != 0
    raise ValueError(f"not a bool: {text!r}")


def attr_value(node):
    """yaml_parse_attribute: scalar / list / map (ordered)."""
    if isinstance(node, list):
        return [attr_value(e) for e in node]
    if isinstance(node, dict):
        return {str(k): attr_value(v) for k, v in node.items()}
    return typed(node)


def model_attributes(node):
    """description_parse_model_attributes_yaml (src/yaml_common.cpp:103-141):
    flatten a map, or a list of maps; std::map::insert keeps the FIRST value of
    a duplicated key across list entries."""
    out = {}
    if isinstance(node, list):
        for e in node:
            for k, v in model_attributes(e).items():
                out.setdefault(k, v)
    elif isinstance(node, dict):
        for k, v in node.items():
            if str(k) not in SKIP_KEYS:
                out[str(k)] = attr_value(v)
    else:
        raise ValueError("Model attributes must be a map or list of maps")
    return out


def parse_range(s):
    m = re.search(r"(\d+)\.\.(\d+)", s)
    if not m:
        raise ValueError(f"bad range {s}")
    a, b = int(m.group(1)), int(m.group(2))
    if a > b:
        raise ValueError("Invalid range; first > last")
    return a, b


def as_list(node):
    return node if isinstance(node, list) else [node]


class Flat:
    def __init__(self, out):
        self.out = out

    def emit(self, *rec):
        self.out.write(json.dumps(list(rec), separators=(",", ":")))
        self.out.write("\n")


# ---------------------------------------------------------------- architecture
def convert_arch(arch_doc, flat, max_tiles=None, plugin_map=None):
    arch = arch_doc["architecture"]
    a = arch["attributes"]
    sync_model = a.get("sync_model", "fixed")
    sync = []
    if sync_model == "fixed":
        sync.append([0, float(typed(a["latency_sync"])) if "latency_sync" in a else 0.0])
    elif sync_model == "table":
        node = a["latency_sync"]
        if isinstance(node, list):
            sync = [[i, float(typed(v))] for i, v in enumerate(node)]
        elif isinstance(node, dict):
            sync = [[int(k), float(typed(v))] for k, v in node.items()]
        else:
            sync = [[0, float(typed(node))]]
    else:
        raise ValueError("Unknown sync_model: " + sync_model)
    flat.emit("noc", arch["name"], {
        "width": int(typed(a["width"])), "height": int(typed(a["height"])),
        "link_buffer_size": int(typed(a["link_buffer_size"])), "sync": sync})

    tiles_made = 0
    for tile in as_list(arch["tile"]):
        tname = tile["name"]
        t0, t1 = parse_range(tname) if ".." in tname else (0, 0)
        for t in range(t0, t1 + 1):
            if max_tiles is not None and tiles_made >= max_tiles:
                break
            tiles_made += 1
            ta = tile["attributes"]
            metrics = {k: float(typed(ta[k])) for k in (
                "energy_north_hop", "latency_north_hop", "energy_east_hop", "latency_east_hop",
                "energy_south_hop", "latency_south_hop", "energy_west_hop", "latency_west_hop")}
            metrics["log_energy"] = to_bool(ta["log_energy"]) if "log_energy" in ta else False
            flat.emit("tile", tname.split("[")[0] + f"[{t}]", metrics)
            for core in as_list(tile["core"]):
                cname = core["name"]
                c0, c1 = parse_range(cname) if ".." in cname else (0, 0)
                for c in range(c0, c1 + 1):
                    convert_core(core, cname.split("[")[0] + f"[{c}]", flat, plugin_map)
    flat.emit("end_arch")


def convert_core(core, name, flat, plugin_map):
    ca = core["attributes"]
    flat.emit("core", name, {
        "buffer_position": ca["buffer_position"],
        "buffer_inside_unit": to_bool(ca["buffer_inside_unit"]) if "buffer_inside_unit" in ca else False,
        "max_neurons_supported": int(typed(ca["max_neurons_supported"])),
        "log_energy": to_bool(ca["log_energy"]) if "log_energy" in ca else False})
    for section in ("axon_in", "synapse", "dendrite", "soma", "axon_out"):
        if section not in core:
            raise ValueError(f"No {section} section defined")
        for unit in as_list(core[section]):
            uname = unit["name"]
            u0, u1 = parse_range(uname) if ".." in uname else (0, 0)
            if ".." in uname and section not in ("axon_in", "axon_out"):
                # keep name[a..b] as ONE record (loihi.yaml has 1024 input units per core)
                ua = unit["attributes"]
                info = {"model": str(ua["model"]),
                        "log_energy": to_bool(ua["log_energy"]) if "log_energy" in ua else False,
                        "log_latency": to_bool(ua["log_latency"]) if "log_latency" in ua else False,
                        "update_every_timestep": to_bool(ua["update_every_timestep"])
                        if "update_every_timestep" in ua else False,
                        "attrs": model_attributes(ua)}
                if "plugin" in ua:
                    p = str(ua["plugin"])
                    info["plugin"] = (plugin_map or {}).get(p, p)
                flat.emit("unit_range", section, uname.split("[")[0], u0, u1, info)
                continue
            for u in range(u0, u1 + 1):
                full = uname.split("[")[0] + f"[{u}]" if ".." in uname else uname
                ua = unit["attributes"]
                if section == "axon_in":
                    flat.emit("axon_in", full, float(typed(ua["energy_message_in"])),
                              float(typed(ua["latency_message_in"])))
                elif section == "axon_out":
                    flat.emit("axon_out", full, float(typed(ua["energy_message_out"])),
                              float(typed(ua["latency_message_out"])))
                else:
                    info = {"model": str(ua["model"]),
                            "log_energy": to_bool(ua["log_energy"]) if "log_energy" in ua else False,
                            "log_latency": to_bool(ua["log_latency"]) if "log_latency" in ua else False,
                            "update_every_timestep": to_bool(ua["update_every_timestep"])
                            if "update_every_timestep" in ua else False,
                            "attrs": model_attributes(ua)}
                    if "plugin" in ua:
                        p = str(ua["plugin"])
                        info["plugin"] = (plugin_map or {}).get(p, p)
                    flat.emit("unit", section, full, info)


# --------------------------------------------------------------------- network
def attrs_with_flags(attrs, syn, den, soma):
    return {k: [v, syn, den, soma] for k, v in attrs.items()}


def neuron_config(node, template):
    """yaml_parse_neuron_attributes (src/yaml_snn.cpp:331-394). `template` and
    the result are dicts: optional hw names / log flags + 'attrs' {k: [v,fs,fd,fso]}."""
    cfg = {k: (dict(v) if k == "attrs" else v) for k, v in template.items()}
    cfg.setdefault("attrs", {})
    if isinstance(node, list):
        for e in node:
            cfg = neuron_config(e, cfg)
        return cfg
    if node is None or (isinstance(node, str) and node.strip() == ""):
        raise ValueError("Model attributes must be a map or list of maps")
    if "log_potential" in node:
        cfg["log_potential"] = to_bool(node["log_potential"])
    if "log_spikes" in node:
        cfg["log_spikes"] = to_bool(node["log_spikes"])
    if "synapse_hw_name" in node:
        cfg["default_synapse_hw_name"] = str(node["synapse_hw_name"])
    if "dendrite_hw_name" in node:
        cfg["dendrite_hw_name"] = str(node["dendrite_hw_name"])
    if "soma_hw_name" in node:
        cfg["soma_hw_name"] = str(node["soma_hw_name"])
    cfg["attrs"].update(attrs_with_flags(model_attributes(node), True, True, True))
    if "dendrite" in node:
        cfg["attrs"].update(attrs_with_flags(model_attributes(node["dendrite"]), False, True, False))
    if "soma" in node:
        cfg["attrs"].update(attrs_with_flags(model_attributes(node["soma"]), False, False, True))
    return cfg


def config_record(cfg):
    rec = {k: v for k, v in cfg.items() if k != "attrs"}
    rec["attrs"] = [[k, v[0], v[1], v[2], v[3]] for k, v in sorted(cfg["attrs"].items())]
    return rec


def count_neurons(neurons_node):
    n = 0
    for entry in neurons_node:
        ids = list(entry.keys()) if isinstance(entry, dict) else (
            [k for e in entry for k in e.keys()] if isinstance(entry, list) else [entry])
        for i in ids:
            i = str(i)
            if ".." in i:
                a, b = parse_range(i)
                n += b - a + 1
            else:
                n += 1
    return n


def edge_attrs(node, syn, den):
    """description_parse_edge_attributes (src/yaml_snn.cpp:831-878)."""
    if isinstance(node, list):
        for e in node:
            edge_attrs(e, syn, den)
        return
    if "synapse" in node:
        syn.update(attrs_with_flags(model_attributes(node["synapse"]), True, False, False))
    if "dendrite" in node:
        den.update(attrs_with_flags(model_attributes(node["dendrite"]), False, True, False))
    for k, v in model_attributes(node).items():
        syn[k] = [v, True, True, True]
        den[k] = [v, True, True, True]


def flag_list(d):
    return [[k, v[0], v[1], v[2], v[3]] for k, v in sorted(d.items())]


CONV_KEYS = ("input_height", "input_width", "input_channels", "kernel_width",
             "kernel_height", "kernel_count", "stride_width", "stride_height")


def convert_net(net_doc, flat):
    net = net_doc["network"]
    group_sizes = {}
    for g in net["groups"]:
        gname = str(g["name"])
        count = count_neurons(g["neurons"])
        default = neuron_config(g["attributes"], {}) if "attributes" in g else {"attrs": {}}
        flat.emit("group", gname, count, config_record(default))
        group_sizes[gname] = count
        for entry in g["neurons"]:
            if not isinstance(entry, (dict, list)):
                continue  # plain "a..b" only counts (src/yaml_snn.cpp:226-278)
            items = entry.items() if isinstance(entry, dict) else [kv for e in entry for kv in e.items()]
            for nid, attrs in items:
                cfg = neuron_config(attrs, default)
                nid = str(nid)
                first, last = parse_range(nid) if ".." in nid else (int(nid), int(nid))
                flat.emit("neurons", gname, first, last, config_record(cfg))

    for entry in net["edges"]:
        for desc, attrs in entry.items():
            src, dst = [s.strip() for s in str(desc).split("->")]
            if "." in src:
                sg, sn = src.split(".", 1)
                dg, dn = dst.split(".", 1)
                syn, den = {}, {}
                edge_attrs(attrs, syn, den)
                flat.emit("edge", sg, int(sn), dg, int(dn),
                          {"synapse_attrs": flag_list(syn), "dendrite_attrs": flag_list(den)})
                continue
            ma = model_attributes(attrs)
            etype = ma.get("type")
            lists = []
            for k, v in sorted(ma.items()):
                if k == "type" or k in CONV_KEYS or k == "source_target_pairs":
                    continue
                if not isinstance(v, list):
                    raise ValueError(f"Attribute must be a list (name: {k})")
                lists.append([k, v, True, True, True])
            if etype == "conv2d":
                params = {"kernel_count": 1, "stride_width": 1, "stride_height": 1}
                params.update({k: ma[k] for k in CONV_KEYS if k in ma})
                flat.emit("conv2d", src, dst, params, lists)
            elif etype == "dense":
                flat.emit("dense", src, dst, lists)
            elif etype == "sparse":
                flat.emit("sparse", src, dst, ma["source_target_pairs"], lists)
            else:
                raise ValueError(f"Invalid hyperedge type: {etype}")

    for entry in net_doc["mappings"]:
        if len(entry) != 1:
            raise ValueError("Should be one entry per mapping")
        for addr, info in entry.items():
            addr = str(addr)
            gname, _, nstr = addr.partition(".")
            if nstr:
                first, last = parse_range(nstr) if ".." in nstr else (int(nstr), int(nstr))
            else:
                first, last = 0, group_sizes[gname] - 1
            hw = {}
            core = None
            for field in as_list(info):
                for k in ("synapse", "dendrite", "soma"):
                    if k in field:
                        hw[k] = str(field[k])
                if "core" in field:
                    core = str(field["core"])
            tile_s, _, core_s = core.partition(".")
            flat.emit("map", gname, first, last, int(tile_s), int(core_s), hw)
    flat.emit("end_net")


def load_yaml(path):
    with open(path) as f:
        return yaml.load(f, Loader=getattr(yaml, "CBaseLoader", yaml.BaseLoader))


def convert(arch_path, net_path, out_path, max_tiles=None, plugin_map=None, synth=None):
    with open(out_path, "w") as out:
        flat = Flat(out)
        convert_arch(load_yaml(arch_path), flat, max_tiles, plugin_map)
        if synth is not None:
            flat.emit("synth", synth)
            flat.emit("end_net")
        else:
            convert_net(load_yaml(net_path), flat)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("arch")
    ap.add_argument("net", nargs="?")
    ap.add_argument("-o", "--out", required=True)
    ap.add_argument("--max-tiles", type=int, default=None)
    ap.add_argument("--synth", default=None, help="JSON synthetic-network spec instead of a net file")
    args = ap.parse_args()
    convert(args.arch, args.net, args.out, args.max_tiles,
            synth=json.loads(args.synth) if args.synth else None)


if __name__ == "__main__":
    sys.exit(main())
