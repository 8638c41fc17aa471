"""Stage trajectory frames into (pinned) host batches for `WaterEntropyEngine.submit`.

Replaces the reference's per-task `system.trajectory[index]` seek
(`recipes/interfacial_solvent.py:285`): frames are gathered into float32
`[F, N, 3]` position / force blocks and `[F, 6]` cell blocks.
"""
from __future__ import annotations

import numpy as np


def _pinned(shape, dtype):
    try:
        import torch

        if torch.cuda.is_available():
            t = torch.empty(shape, dtype=getattr(torch, np.dtype(dtype).name), pin_memory=True)
            return t.numpy(), t
    except Exception:  # pragma: no cover
        pass
    a = np.empty(shape, dtype=dtype)
    return a, a


def iter_batches(system, indices, batch):
    """Yield (positions, forces, dimensions, frame_ids) for consecutive slices of `indices`."""
    indices = list(indices)
    n_atoms = system.n_atoms if hasattr(system, "n_atoms") else len(system.atoms)
    for b0 in range(0, len(indices), batch):
        sel = indices[b0 : b0 + batch]
        P, keepP = _pinned((len(sel), n_atoms, 3), np.float32)
        F, keepF = _pinned((len(sel), n_atoms, 3), np.float32)
        if hasattr(system, "frame_arrays"):  # waterentropy_b200.io.Universe: read into pinned memory
            _, _, D = system.frame_arrays(sel, P, F)
            yield keepP, keepF, np.ascontiguousarray(D), sel
            continue
        D = np.empty((len(sel), 6), dtype=np.float32)
        ids = []
        for k, i in enumerate(sel):
            if hasattr(system, "frame"):  # SyntheticSystem
                P[k], F[k], D[k] = system.frame(i)
                ids.append(i)
            else:  # MDAnalysis-like Universe
                ts = system.trajectory[i]
                P[k] = system.atoms.positions
                F[k] = system.atoms.forces
                D[k] = system.dimensions
                ids.append(int(ts.frame))
        yield keepP, keepF, D, ids
