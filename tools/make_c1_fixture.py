"""TEST INFRASTRUCTURE -- BASELINE.json config C1 as a committed fixture.

    python tools/make_c1_fixture.py          # build container only (/root/reference is mounted here)

Reads the reference's example input `example_inputs/lysozyme_solution/system.gro` with the product's GRO
reader and stores what that file holds - atom names, residue names / numbers, the single frame of
positions (float32, Angstrom) and the box - as `tests/golden/c1_lysozyme_inputs.npz`, because
/root/reference does not exist on the GPU box.  The tests write a GRO file back from these arrays
(`%8.3f` nm reproduces the original numbers) and run the `waterShells` path on it.
"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from waterentropy_b200.io.gro import GroFile  # noqa: E402

SRC = "/root/reference/example_inputs/lysozyme_solution/system.gro"
g = GroFile(SRC)
out = os.path.join(REPO, "tests", "golden", "c1_lysozyme_inputs.npz")
np.savez_compressed(out, names=np.array(g.names), resnames=np.array(g.resnames), resids=g.resids.astype(np.int32),
                    positions=g.positions, dimensions=g.dimensions)
print(out, os.path.getsize(out), "bytes;", g.n_atoms, "atoms,", g.n_frames, "frame(s)")
