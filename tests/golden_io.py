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


def write_gro(path, names, resnames, resids, positions, dimensions, title="written by tests/golden_io.py"):
    """GRO text (nm, %8.3f) of one or more frames: positions [F, N, 3] in Angstrom, dimensions [F, 6]."""
    import numpy as np

    positions = np.asarray(positions)
    if positions.ndim == 2:
        positions, dimensions = positions[None], np.asarray(dimensions)[None]
    with open(path, "w") as fh:
        for P, D in zip(positions, dimensions):
            fh.write(f"{title}\n{len(names):5d}\n")
            nm = np.asarray(P, dtype=np.float64) / 10.0
            for k in range(len(names)):
                fh.write(f"{int(resids[k]) % 100000:5d}{str(resnames[k]):<5s}{str(names[k]):>5s}{(k + 1) % 100000:5d}"
                         f"{nm[k, 0]:8.3f}{nm[k, 1]:8.3f}{nm[k, 2]:8.3f}\n")
            fh.write(f"{D[0] / 10:10.5f}{D[1] / 10:10.5f}{D[2] / 10:10.5f}\n")
