"""Load the committed golden fixtures (tests/golden/, made by oracle/make_golden.py)."""
import gzip
import json
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def dec(o):
    if isinstance(o, dict):
        if "d" in o and len(o) == 1:
            return {dec(k): dec(v) for k, v in o["d"]}
        if "t" in o and len(o) == 1:
            return tuple(dec(v) for v in o["t"])
        if "a" in o and len(o) == 1:
            return np.array(o["a"], dtype=np.float64)
        return {k: dec(v) for k, v in o.items()}
    if isinstance(o, list):
        return [dec(v) for v in o]
    return o


def load_reference(tag):
    with gzip.open(os.path.join(GOLD, f"{tag}_reference.json.gz"), "rt") as fh:
        return dec(json.load(fh))


def load_inputs(tag):
    z = np.load(os.path.join(GOLD, f"{tag}_inputs.npz"))
    return {k: z[k] for k in z.files}


def oracle_topology(inp):
    import we_oracle as wo

    return wo.OracleTopology(
        inp["masses"], inp["charges"], inp["resids"], list(inp["resnames"]),
        list(inp["types"]), inp["bonds"], list(inp["names"]),
    )


def plain(d):
    """nested defaultdict -> plain dict"""
    if isinstance(d, dict):
        return {k: plain(v) for k, v in d.items()}
    return d


TAGS = ["arginine", "aspirin", "synth_ortho", "synth_tric", "synth_sparse"]
