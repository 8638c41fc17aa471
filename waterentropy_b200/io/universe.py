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
    """Frames held in memory."""

    def __init__(self, positions, forces, dimensions):
        self.P = np.ascontiguousarray(positions, dtype=np.float32)
        self.F = np.ascontiguousarray(forces, dtype=np.float32)
        self.dimensions = np.ascontiguousarray(dimensions, dtype=np.float32)
        self.n_frames = self.P.shape[0]

    def read(self, indices, out_pos=None, out_frc=None):
        idx = np.asarray(list(indices), dtype=np.int64)
        if out_pos is not None:
            out_pos[: len(idx)] = self.P[idx]
            out_frc[: len(idx)] = self.F[idx]
            return out_pos[: len(idx)], out_frc[: len(idx)], self.dimensions[idx]
        return self.P[idx], self.F[idx], self.dimensions[idx]


class Universe:
    """`Universe(topology, coordinates)` for AMBER prmtop + NetCDF and prmtop + GROMACS TRR
    (frames streamed from memory-mapped files), or `Universe.from_arrays(...)`."""

    def __init__(self, topology=None, coordinates=None):
        self._frame = 0
        self._cache = (-1, None, None)
        if topology is None:
            return
        ext_t = os.path.splitext(str(topology))[1].lower()
        ext_c = os.path.splitext(str(coordinates))[1].lower()
        if ext_t in (".prmtop", ".parm7", ".top") and ext_c in (".nc", ".ncdf"):
            from .amber import AmberTopology, NetCDFTrajectory

            t = AmberTopology(topology)
            x = NetCDFTrajectory(coordinates)
        elif ext_c == ".trr":
            from .gromacs import open_gromacs

            t, x = open_gromacs(topology, coordinates)
        else:
            raise NotImplementedError(
                f"no reader for topology {ext_t!r} + coordinates {ext_c!r} "
                "(supported: AMBER prmtop + NetCDF, prmtop + GROMACS TRR)")
        if not x.has_forces:
            raise ValueError("the trajectory holds no forces; the covariance path needs them")
        if x.n_atoms != t.n_atoms:
            raise ValueError(f"trajectory has {x.n_atoms} atoms, topology {t.n_atoms}")
        self._set_topology(t.names, t.types, t.resnames, t.resids, t.masses, t.charges, t.bonds)
        self._source = x

    @classmethod
    def from_arrays(cls, names, types, resnames, resids, masses, charges, bonds, positions, forces, dimensions):
        u = cls()
        if forces is None:
            raise ValueError("the trajectory holds no forces; the covariance path needs them")
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
            self._cache = (self._frame, P[0], F[0])
        return self._cache[1], self._cache[2]

    @property
    def dimensions(self):
        return self._source.dimensions[self._frame]

    # fast path used by waterentropy_b200.frames: a batch of frames straight into (pinned) buffers
    def frame_arrays(self, indices, out_pos=None, out_frc=None):
        return self._source.read(indices, out_pos, out_frc)
