"""Minimal GROMACS trajectory reader: TRR (XDR, positions + forces + box).

The reference's second fixture (`tests/input_files/gromacs/aspirin_solution`)
pairs a TRR trajectory with an AMBER PRMTOP topology file; MDAnalysis' TRR
reader (libxdrfile) is replaced by this pure-NumPy parser.  Units follow
MDAnalysis: nm -> Angstrom (x10), kJ mol^-1 nm^-1 -> kJ mol^-1 A^-1 (/10), box
vectors -> [lx, ly, lz, alpha, beta, gamma] in float32.
"""
from __future__ import annotations

import os
import struct

import numpy as np


class TRRTrajectory:
    """GROMACS TRR: one pass over the frame headers builds an index, frames are then read on
    demand from a memory map (`read(indices)`)."""

    def __init__(self, path: str):
        self._data = np.memmap(path, dtype=np.uint8, mode="r")
        data = self._data
        off = 0
        self._frames = []  # (x offset, f offset or -1, dtype, box)
        self.n_atoms = 0
        while off < len(data):
            magic, _slen = struct.unpack_from(">ii", data, off)
            if magic != 1993:
                raise ValueError(f"{path}: bad TRR magic {magic} at byte {off}")
            off += 8
            (n,) = struct.unpack_from(">i", data, off)
            off += 4 + ((n + 3) // 4) * 4
            (ir, e, box, vir, pres, top, sym, xs, vs, fs, natoms, _step, _nre) = struct.unpack_from(">13i", data, off)
            off += 52
            real = 8 if (box == 72 or (box == 0 and natoms and max(xs, vs, fs) == natoms * 24)) else 4
            dt = ">f8" if real == 8 else ">f4"
            off += 2 * real  # time, lambda
            cell = np.zeros((3, 3))
            if box:
                cell = np.frombuffer(data, dtype=dt, count=9, offset=off).reshape(3, 3).astype(np.float64)
                off += box
            off += vir + pres
            xo = off if xs else -1
            off += xs + vs
            fo = off if fs else -1
            off += fs + ir + e + top + sym
            if xo >= 0:
                self._frames.append((xo, fo, dt, _dimensions(cell * 10.0)))
                self.n_atoms = natoms
        self.n_frames = len(self._frames)
        self.has_forces = self.n_frames > 0 and all(f[1] >= 0 for f in self._frames)
        self.dimensions = np.stack([f[3] for f in self._frames]).astype(np.float32) if self._frames else np.zeros((0, 6), np.float32)

    def read(self, indices, out_pos=None, out_frc=None):
        idx = np.asarray(list(indices), dtype=np.int64)
        n, na = len(idx), self.n_atoms
        P = out_pos if out_pos is not None else np.empty((n, na, 3), dtype=np.float32)
        F = out_frc if out_frc is not None else np.empty((n, na, 3), dtype=np.float32)
        if n and not self._convert_native(idx, P, F):
            for k, i in enumerate(idx):
                xo, fo, dt, _ = self._frames[int(i)]
                P[k] = np.frombuffer(self._data, dtype=dt, count=na * 3, offset=xo).reshape(na, 3)
                if self.has_forces:
                    F[k] = np.frombuffer(self._data, dtype=dt, count=na * 3, offset=fo).reshape(na, 3)
            P[:n] *= np.float32(10.0)  # nm -> A
            if self.has_forces:
                F[:n] /= np.float32(10.0)  # kJ/mol/nm -> kJ/mol/A
        return P[:n], (F[:n] if self.has_forces else None), self.dimensions[idx]

    def _convert_native(self, idx, P, F):
        """Single-precision frames through `we_convert_be_f32` (host threads: XDR byte swap + the float32 unit
        conversion in one pass, straight into the caller's blocks); False for double-precision files."""
        from .. import _lib

        if os.environ.get("WE_NATIVE_READER", "1") == "0":
            return False
        frames = [self._frames[int(i)] for i in idx]
        if any(fr[2] != ">f4" for fr in frames) or P.dtype != np.float32 or not P.flags.c_contiguous:
            return False
        if self.has_forces and (F.dtype != np.float32 or not F.flags.c_contiguous):
            return False
        base = self._data.ctypes.data
        n_each = self.n_atoms * 3
        _lib.convert_be_f32([base + fr[0] for fr in frames], [P[k].ctypes.data for k in range(len(frames))], n_each, 1, 10.0)
        if self.has_forces:
            _lib.convert_be_f32([base + fr[1] for fr in frames], [F[k].ctypes.data for k in range(len(frames))], n_each, 2, 10.0)
        return True


def read_trr(path: str):
    """Whole file: (positions [F,N,3] f32, forces [F,N,3] f32 or None, dimensions [F,6] f32)."""
    t = TRRTrajectory(path)
    return t.read(range(t.n_frames))


def _dimensions(cell):
    a, b, c = cell
    la, lb, lc = np.linalg.norm(a), np.linalg.norm(b), np.linalg.norm(c)
    if la == 0 or lb == 0 or lc == 0:
        return np.zeros(6)

    def ang(u, v, lu, lv):
        return np.degrees(np.arccos(np.clip(np.dot(u, v) / (lu * lv), -1.0, 1.0)))

    alpha, beta, gamma = ang(b, c, lb, lc), ang(a, c, la, lc), ang(a, b, la, lb)
    out = np.array([la, lb, lc, alpha, beta, gamma])
    out[3:] = np.where(np.abs(out[3:] - 90.0) < 1e-5, 90.0, out[3:])
    return out


def open_gromacs(topology: str, trajectory: str):
    """(AmberTopology, TRRTrajectory) with the compatibility checks of `load_gromacs`."""
    ext = os.path.splitext(str(topology))[1].lower()
    if ext not in (".prmtop", ".parm7", ".top"):
        raise NotImplementedError(
            f"topology {topology!r}: a GRO/TPR file carries no masses, charges or bonds that this reader "
            "can use; pass the PRMTOP topology of the system (TPR parsing is not implemented)")
    from .amber import AmberTopology

    t = AmberTopology(topology)
    x = TRRTrajectory(trajectory)
    if x.n_frames == 0:
        raise ValueError(f"{trajectory}: no coordinate frames")
    if x.n_atoms != t.n_atoms:
        raise ValueError(f"{trajectory}: {x.n_atoms} atoms, topology has {t.n_atoms}")
    return t, x


def load_gromacs(topology: str, trajectory: str):
    """(names, types, resnames, resids, masses, charges, bonds, positions, forces, dimensions)."""
    ext = os.path.splitext(str(topology))[1].lower()
    if ext not in (".prmtop", ".parm7", ".top"):
        raise NotImplementedError(
            f"topology {topology!r}: a GRO/TPR file carries no masses, charges or bonds that this reader "
            "can use; pass the PRMTOP topology of the system (TPR parsing is not implemented)")
    from .amber import AmberTopology

    t = AmberTopology(topology)
    pos, frc, dims = read_trr(trajectory)
    if pos is None:
        raise ValueError(f"{trajectory}: no coordinate frames")
    if pos.shape[1] != t.n_atoms:
        raise ValueError(f"{trajectory}: {pos.shape[1]} atoms, topology has {t.n_atoms}")
    return t.names, t.types, t.resnames, t.resids, t.masses, t.charges, t.bonds, pos, frc, dims
