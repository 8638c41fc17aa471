"""Stage trajectory frames into pinned host batches for `WaterEntropyEngine.submit`.

Replaces the reference's per-task `system.trajectory[index]` seek
(`recipes/interfacial_solvent.py:285`): frames are gathered into float32
`[F, N, 3]` position / force blocks and `[F, 6]` cell blocks.  The blocks come
from a small ring of page-locked buffers that is recycled as the engine retires
batches (`we_frames_retired`), so host memory stays O(ring) however long the
trajectory is - the reference holds one frame at a time.
"""
from __future__ import annotations

import numpy as np

RING_SLOTS = 6  # the engine keeps four batches in flight; two more are being filled / just submitted


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


def _read_into(system, sel, P, F):
    """Fill P[:len(sel)] (and F unless None) with the frames `sel`; returns (dimensions, frame ids)."""
    n = len(sel)
    if hasattr(system, "frame_arrays"):  # waterentropy_b200.io.Universe: read straight into the block
        _, _, D = system.frame_arrays(sel, P, F)
        return np.ascontiguousarray(D, dtype=np.float32), list(sel)
    D = np.empty((n, 6), dtype=np.float32)
    ids = []
    for k, i in enumerate(sel):
        if hasattr(system, "frame"):  # SyntheticSystem
            p, f, D[k] = system.frame(i)
            P[k] = p
            if F is not None:
                F[k] = f
            ids.append(i)
        else:  # MDAnalysis-like Universe
            ts = system.trajectory[i]
            P[k] = system.atoms.positions
            if F is not None:
                F[k] = system.atoms.forces
            D[k] = system.dimensions
            ids.append(int(ts.frame))
    return D, ids


def iter_batches(system, indices, batch, engine=None, forces=True, slots=RING_SLOTS):
    """Yield (positions, forces, dimensions, frame_ids) for consecutive slices of `indices`.

    The caller submits every batch to `engine` before asking for the next one; a ring slot is
    refilled only after the engine has retired the batch that used it.  Without an engine every
    batch gets fresh buffers (the caller keeps them alive)."""
    indices = list(indices)
    if not indices:
        return
    n_atoms = system.n_atoms if hasattr(system, "n_atoms") else len(system.atoms)
    # frames already resident in page-locked memory (io.Universe.pin()): hand out views, no staging copy
    views = getattr(system, "frame_views", None)
    n_batches = (len(indices) + batch - 1) // batch
    ring = []
    ends = []
    base = len(engine.frame_ids) if engine is not None else 0
    done = 0
    for b, b0 in enumerate(range(0, len(indices), batch)):
        sel = indices[b0:b0 + batch]
        if views is not None:
            got = views(sel, forces)
            if got is not None:
                done += len(sel)
                yield got[0], got[1], got[2], list(sel)
                continue
        if engine is None:
            slot = (_pinned((len(sel), n_atoms, 3), np.float32), _pinned((len(sel), n_atoms, 3), np.float32) if forces else None)
        else:
            s = b % slots
            if s >= len(ring):
                rows = min(batch, len(indices))
                ring.append((_pinned((rows, n_atoms, 3), np.float32), _pinned((rows, n_atoms, 3), np.float32) if forces else None))
                ends.append(0)
            elif ends[s]:
                engine.wait_retired(ends[s])  # the batch that last used this slot is off the host buffers
            slot = ring[s]
        (P, keepP), fslot = slot
        F, keepF = fslot if fslot is not None else (None, None)
        D, ids = _read_into(system, sel, P, F)
        n = len(sel)
        done += n
        if engine is not None:
            ends[b % slots] = base + done
        yield (keepP[:n], None if keepF is None else keepF[:n], D, ids)
