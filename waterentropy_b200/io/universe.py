"""A minimal in-memory `Universe` for the file formats of the reference's fixtures.

MDAnalysis is the reference's reader (`cli/runWaterEntropy.py:37`); it is not
installable here, so the CLI and the tests load trajectories with this class.
It exposes only what the hot path reads -- topology arrays, float32 frames --
with MDAnalysis attribute names, so a real `MDAnalysis.Universe` can be passed
to the recipes instead without any change.
"""
from __future__ import annotations

import os

import numpy as np


class _Timestep:
    def __init__(self, frame):
        self.frame = frame


class _Bonds:
    def __init__(self, pairs):
        self.indices = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)


class _Atoms:
    def __init__(self, u):
        self._u = u

    def __len__(self):
        return self._u.n_atoms

    masses = property(lambda s: s._u._masses)
    charges = property(lambda s: s._u._charges)
    resids = property(lambda s: s._u._resids)
    resnames = property(lambda s: s._u._resnames)
    types = property(lambda s: s._u._types)
    names = property(lambda s: s._u._names)
    indices = property(lambda s: np.arange(s._u.n_atoms))
    bonds = property(lambda s: s._u.bonds)
    positions = property(lambda s: s._u._current()[0])
    forces = property(lambda s: s._u._current()[1])


class _Trajectory:
    def __init__(self, u):
        self._u = u

    @property
    def n_frames(self):
        return self._u._source.n_frames

    def __len__(self):
        return self.n_frames

    def __repr__(self):  # what `print(u.trajectory)` shows (cli/run_find_Nc.py:36)
        return f"<{type(self._u._source).__name__} with {self.n_frames} frames of {self._u.n_atoms} atoms>"

    @property
    def ts(self):
        return _Timestep(self._u._frame)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return (self[k] for k in range(*i.indices(self.n_frames)))
        i = int(i)
        if i < 0:
            i += self.n_frames
        if not 0 <= i < self.n_frames:
            raise IndexError(f"frame {i} out of range [0, {self.n_frames})")
        self._u._frame = i
        return _Timestep(i)

    def __iter__(self):
        return self[0 : self.n_frames]


class _ArraySource:
    """Frames held in memory (forces optional: the shells path does not read them)."""

    def __init__(self, positions, forces, dimensions):
        self.P = np.ascontiguousarray(positions, dtype=np.float32)
        self.F = None if forces is None else np.ascontiguousarray(forces, dtype=np.float32)
        self.dimensions = np.ascontiguousarray(dimensions, dtype=np.float32)
        self.n_frames = self.P.shape[0]
        self.n_atoms = self.P.shape[1]
        self.has_forces = self.F is not None
        self._keep = None  # pinned torch tensors backing P / F after pin()

    def read(self, indices, out_pos=None, out_frc=None):
        idx = np.asarray(list(indices), dtype=np.int64)
        if out_pos is not None:
            out_pos[: len(idx)] = self.P[idx]
            if out_frc is not None:
                if self.F is None:
                    raise ValueError("the trajectory holds no forces; the covariance path needs them")
                out_frc[: len(idx)] = self.F[idx]
            return out_pos[: len(idx)], None if out_frc is None else out_frc[: len(idx)], self.dimensions[idx]
        return self.P[idx], None if self.F is None else self.F[idx], self.dimensions[idx]

    def pin(self):
        """Move the frames into page-locked memory so that batches can be handed to the engine in place."""
        import torch

        if self._keep is None and torch.cuda.is_available():
            tp = torch.from_numpy(self.P).pin_memory()
            tf = None if self.F is None else torch.from_numpy(self.F).pin_memory()
            self._keep = (tp, tf)
            self.P = tp.numpy()
            self.F = None if tf is None else tf.numpy()

    def views(self, indices, forces=True):
        """(positions, forces, dimensions) views of a contiguous run of frames in pinned memory, or None."""
        if self._keep is None or not len(indices):
            return None
        i0, n = int(indices[0]), len(indices)
        if list(indices) != list(range(i0, i0 + n)):
            return None
        tp, tf = self._keep
        if forces and tf is None:
            raise ValueError("the trajectory holds no forces; the covariance path needs them")
        return tp[i0:i0 + n], (tf[i0:i0 + n] if forces else None), self.dimensions[i0:i0 + n]


class Universe:
    """`Universe(topology, coordinates)` for AMBER prmtop + NetCDF and prmtop + GROMACS TRR
    (frames streamed from memory-mapped files), or `Universe.from_arrays(...)`."""

    def __init__(self, topology=None, coordinates=None):
        self._frame = 0
        self._cache = (-1, None, None)
        self.has_charges = True
        if topology is None:
            return
        ext_t = os.path.splitext(str(topology))[1].lower()
        ext_c = os.path.splitext(str(coordinates))[1].lower()
        if ext_t in (".prmtop", ".parm7", ".top") and ext_c in (".nc", ".ncdf"):
            from .amber import AmberTopology, NetCDFTrajectory

            t = AmberTopology(topology)
            x = NetCDFTrajectory(coordinates)
        elif ext_t == ".gro":
            # positions + names only: a template topology (no charges), good for the shells path
            from .gro import GroFile, GroTopology

            g = GroFile(topology)
            t = GroTopology(g.names, g.resnames, g.resids, g.positions[0], g.dimensions[0])
            if ext_c == ".gro":
                x = g if str(coordinates) == str(topology) else GroFile(coordinates)
            elif ext_c == ".trr":
                from .gromacs import TRRTrajectory

                x = TRRTrajectory(coordinates)
            else:
                raise NotImplementedError(f"no reader for coordinates {ext_c!r} with a GRO topology")
        elif ext_c == ".trr":
            from .gromacs import open_gromacs

            t, x = open_gromacs(topology, coordinates)
        else:
            raise NotImplementedError(
                f"no reader for topology {ext_t!r} + coordinates {ext_c!r} "
                "(supported: AMBER prmtop + NetCDF, prmtop + GROMACS TRR, GRO + GRO / TRR for the shells path)")
        if x.n_atoms != t.n_atoms:
            raise ValueError(f"trajectory has {x.n_atoms} atoms, topology {t.n_atoms}")
        self._set_topology(t.names, t.types, t.resnames, t.resids, t.masses, t.charges, t.bonds)
        self.has_charges = bool(getattr(t, "has_charges", True))
        self._source = x

    @classmethod
    def from_arrays(cls, names, types, resnames, resids, masses, charges, bonds, positions, forces, dimensions):
        u = cls()
        u._set_topology(names, types, resnames, resids, masses, charges, bonds)
        u._source = _ArraySource(positions, forces, dimensions)
        return u

    def _set_topology(self, names, types, resnames, resids, masses, charges, bonds):
        self._names = np.asarray(names, dtype=object)
        self._types = np.asarray(types, dtype=object)
        self._resnames = np.asarray(resnames, dtype=object)
        self._resids = np.asarray(resids, dtype=np.int64)
        self._masses = np.asarray(masses, dtype=np.float64)
        self._charges = np.asarray(charges, dtype=np.float64)
        self.bonds = _Bonds(bonds)
        self.n_atoms = len(self._masses)
        self.atoms = _Atoms(self)
        self.trajectory = _Trajectory(self)

    def _current(self):
        if self._cache[0] != self._frame:
            P, F, _ = self._source.read([self._frame])
            self._cache = (self._frame, P[0], None if F is None else F[0])
        return self._cache[1], self._cache[2]

    @property
    def dimensions(self):
        return self._source.dimensions[self._frame]

    @property
    def has_forces(self):
        return bool(getattr(self._source, "has_forces", True))

    # fast path used by waterentropy_b200.frames: a batch of frames straight into (pinned) buffers
    def frame_arrays(self, indices, out_pos=None, out_frc=None):
        return self._source.read(indices, out_pos, out_frc)

    def pin(self):
        """In-memory trajectories: page-lock the frames once; the recipes then submit them in place
        (no staging copy, zero-copy force reads)."""
        if hasattr(self._source, "pin"):
            self._source.pin()
        return self

    def frame_views(self, indices, forces=True):
        fn = getattr(self._source, "views", None)
        return fn(indices, forces) if fn else None
