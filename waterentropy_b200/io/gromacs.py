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


def read_trr(path: str):
    """Return (positions [F,N,3] f32 or None, forces [F,N,3] f32 or None, dimensions [F,6] f32)."""
    data = open(path, "rb").read()
    off = 0
    P, F, D = [], [], []
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
        x = f = None
        if xs:
            x = np.frombuffer(data, dtype=dt, count=natoms * 3, offset=off).reshape(natoms, 3)
            off += xs
        off += vs
        if fs:
            f = np.frombuffer(data, dtype=dt, count=natoms * 3, offset=off).reshape(natoms, 3)
            off += fs
        off += ir + e + top + sym
        if x is not None:
            P.append((x.astype(np.float32) * np.float32(10.0)))
            F.append(None if f is None else (f.astype(np.float32) / np.float32(10.0)))
            D.append(_dimensions(cell * 10.0))
    forces = None if (not F or any(v is None for v in F)) else np.stack(F)
    return (np.stack(P) if P else None), forces, np.stack(D).astype(np.float32)


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
