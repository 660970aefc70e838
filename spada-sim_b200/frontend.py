"""CLI grammar and config -- mirror of the reference's src/frontend.rs.

``Cli`` (frontend.rs:52-75): positional ``<simulator> <accelerator> <category> <workload>
<configuration>`` plus ``-p/--preprocess``; the three enums are case-insensitive.
``parse_config`` (frontend.rs:77-85) prints the path and deserialises ``OmegaConfig``
(frontend.rs:8-23); every key is required, as with serde.
"""
from __future__ import annotations

import argparse
import json
from dataclasses import dataclass, fields
from typing import List

MODES = ["AccurateSimu", "TrafficModel", "BReuseCounter"]
ACCELERATORS = ["Ip", "Op", "MultiRow", "Spada"]
CATEGORIES = ["SS", "NN"]


@dataclass
class OmegaConfig:
    ss_filepath: str
    nn_filepath: str
    pe_num: int
    at_num: int
    lane_num: int
    cache_size: int
    word_byte: int
    block_shape: List[int]
    mem_latency: int
    cache_latency: int
    freq: float
    channel: int
    bandwidth_per_channel: float


@dataclass
class Cli:
    simulator: str
    accelerator: str
    category: str
    workload: str
    configuration: str
    preprocess: bool = False


def _enum(variants):
    table = {v.lower(): v for v in variants}

    def conv(s: str) -> str:
        try:
            return table[s.lower()]
        except KeyError:
            raise argparse.ArgumentTypeError(f"'{s}' isn't a valid value [possible values: {', '.join(variants)}]")
    return conv


def parse_args(argv=None) -> Cli:
    p = argparse.ArgumentParser(prog="spada-sim", description="B200-native drop-in for the spada-sim SpGEMM path")
    p.add_argument("simulator", type=_enum(MODES), help="The simulator to use.")
    p.add_argument("accelerator", type=_enum(ACCELERATORS), help="The accelerator to simulate.")
    p.add_argument("category", type=_enum(CATEGORIES), help="The workload category to search for the workload.")
    p.add_argument("workload", help="The workload name.")
    p.add_argument("configuration", help="Configuration file path.")
    p.add_argument("-p", "--preprocess", action="store_true", help="Preprocessing.")
    ns = p.parse_args(argv)
    return Cli(ns.simulator, ns.accelerator, ns.category, ns.workload, ns.configuration, ns.preprocess)


def parse_config(config_fp: str) -> OmegaConfig:
    print(config_fp)
    with open(config_fp, "r") as f:
        raw = json.load(f)
    missing = [f.name for f in fields(OmegaConfig) if f.name not in raw]
    if missing:
        raise KeyError(f"missing field `{missing[0]}`")
    cfg = OmegaConfig(**{f.name: raw[f.name] for f in fields(OmegaConfig)})
    if len(cfg.block_shape) != 2:
        raise ValueError("block_shape must have 2 entries")
    return cfg
