"""Minimal AMBER readers (PRMTOP topology + NetCDF3 trajectory).

MDAnalysis is what the reference uses to parse these files
(`/root/reference/tests/input_files/load_inputs.py:10-21`); it is not a
dependency here.  Only the fields the interfacial-solvent hot path consumes are
read: names, charges, masses, types, residues and bonds from the PRMTOP, and
float32 coordinates / forces / cell from the NetCDF trajectory.

Unit conventions follow MDAnalysis so that numbers match the reference:
  * charges  = CHARGE / 18.2223 (float64)
  * forces   = kcal mol^-1 A^-1 * float32(4.184), multiplied in float32
  * cell     = float32 [lx, ly, lz, alpha, beta, gamma]
"""
from __future__ import annotations

import re

import os

import numpy as np


def _parse_format(fmt: str):
    m = re.match(r"\(?(\d+)([aAiIeEfF])(\d+)", fmt.strip().lstrip("("))
    if not m:
        raise ValueError(f"unsupported PRMTOP format {fmt!r}")
    return int(m.group(1)), m.group(2).lower(), int(m.group(3))


def read_prmtop_sections(path: str) -> dict:
    """Return {FLAG: list of raw fixed-width string fields}."""
    sections: dict[str, list[str]] = {}
    kinds: dict[str, str] = {}
    flag = None
    width = 0
    with open(path, "r", encoding="ascii", errors="replace") as fh:
        for line in fh:
            if line.startswith("%VERSION") or line.startswith("%COMMENT"):
                continue
            if line.startswith("%FLAG"):
                flag = line[5:].strip()
                sections[flag] = []
                continue
            if line.startswith("%FORMAT"):
                _, kind, width = _parse_format(line[7:].strip())
                kinds[flag] = kind
                continue
            if flag is None:
                continue
            body = line.rstrip("\n")
            for i in range(0, len(body), width):
                field = body[i : i + width]
                if kinds[flag] == "a":
                    if len(field.strip()) or len(field) == width:
                        sections[flag].append(field)
                elif field.strip():
                    sections[flag].append(field)
    sections["__kinds__"] = kinds  # type: ignore[assignment]
    return sections


class AmberTopology:
    """Frame-invariant arrays parsed from a PRMTOP file."""

    def __init__(self, path: str):
        sec = read_prmtop_sections(path)
        pointers = [int(x) for x in sec["POINTERS"]]
        n = pointers[0]
        self.n_atoms = n
        self.names = np.array([s.strip() for s in sec["ATOM_NAME"][:n]], dtype=object)
        self.types = np.array(
            [s.strip() for s in sec["AMBER_ATOM_TYPE"][:n]], dtype=object
        )
        self.charges = np.array([float(x) for x in sec["CHARGE"][:n]]) / 18.2223
        self.masses = np.array([float(x) for x in sec["MASS"][:n]], dtype=np.float64)
        res_labels = [s.strip() for s in sec["RESIDUE_LABEL"]]
        res_ptr = [int(x) - 1 for x in sec["RESIDUE_POINTER"]] + [n]
        resids = np.empty(n, dtype=np.int64)
        resnames = np.empty(n, dtype=object)
        for r, name in enumerate(res_labels):
            resids[res_ptr[r] : res_ptr[r + 1]] = r + 1
            resnames[res_ptr[r] : res_ptr[r + 1]] = name
        self.resids = resids
        self.resnames = resnames
        bonds = []
        for key in ("BONDS_INC_HYDROGEN", "BONDS_WITHOUT_HYDROGEN"):
            vals = [int(x) for x in sec.get(key, [])]
            for i in range(0, len(vals) - 2, 3):
                a, b = vals[i] // 3, vals[i + 1] // 3
                bonds.append((a, b) if a < b else (b, a))
        # MDAnalysis canonicalises each bond to (low, high) and de-duplicates
        self.bonds = np.array(sorted(set(bonds)), dtype=np.int64).reshape(-1, 2)


class NetCDFTrajectory:
    """AMBER NetCDF3 trajectory: coordinates, forces, cell.  Frames are read on demand from a
    memory-mapped file (`read(indices)`), so trajectories larger than RAM can be streamed."""

    def __init__(self, path: str):
        from scipy.io import netcdf_file

        self._nc = netcdf_file(path, "r", mmap=True)
        var = self._nc.variables
        self._xyz = var["coordinates"]
        self.n_frames, self.n_atoms = int(self._xyz.shape[0]), int(self._xyz.shape[1])
        self._frc = var["forces"] if "forces" in var else None
        self.has_forces = self._frc is not None
        if "cell_lengths" in var:
            lengths = np.array(var["cell_lengths"][:], dtype=np.float64)
            angles = np.array(var["cell_angles"][:], dtype=np.float64)
            self.dimensions = np.concatenate([lengths, angles], axis=1).astype(np.float32)
        else:
            self.dimensions = np.zeros((self.n_frames, 6), dtype=np.float32)

    def read(self, indices, out_pos=None, out_frc=None):
        """float32 native-endian (positions, forces, dimensions) of the given frames."""
        idx = np.asarray(list(indices), dtype=np.int64)
        n = len(idx)
        P = out_pos if out_pos is not None else np.empty((n, self.n_atoms, 3), dtype=np.float32)
        F = out_frc if out_frc is not None else np.empty((n, self.n_atoms, 3), dtype=np.float32)
        if n and not self._convert_native(idx, P, F):
            for k, i in enumerate(idx):
                P[k] = self._xyz[int(i)]  # big-endian file -> native float32
                if self._frc is not None:
                    F[k] = self._frc[int(i)]
            if self._frc is not None:
                F[:n] *= np.float32(4.184)  # kcal -> kJ, float32 like MDAnalysis
        return P[:n], (F[:n] if self._frc is not None else None), self.dimensions[idx]

    def _convert_native(self, idx, P, F):
        """The frames' records through `we_convert_be_f32` (host threads: byte swap + the float32 kcal -> kJ
        factor in one pass, straight into the caller's blocks).  False when the layout is not the plain one
        (then the NumPy path above does the same arithmetic)."""
        from .. import _lib

        if os.environ.get("WE_NATIVE_READER", "1") == "0":
            return False
        n_each = self.n_atoms * 3

        def addrs(var, out):
            data = var.data
            if data.dtype != np.dtype(">f4") or out.dtype != np.float32 or not out.flags.c_contiguous:
                return None
            src, dst = [], []
            for k, i in enumerate(idx):
                blk = data[int(i)]
                if not blk.flags.c_contiguous or blk.size != n_each:
                    return None
                src.append(blk.ctypes.data)
                dst.append(out[k].ctypes.data)
            return src, dst

        jobs = [(addrs(self._xyz, P), 0, 1.0)]
        if self._frc is not None:
            jobs.append((addrs(self._frc, F), 1, 4.184))
        if any(j[0] is None for j in jobs):
            return False
        for (src, dst), op, factor in jobs:
            _lib.convert_be_f32(src, dst, n_each, op, factor)
        return True

    def close(self):
        self._xyz = self._frc = None  # drop the memory-mapped views before closing the file
        nc, self._nc = getattr(self, "_nc", None), None
        if nc is not None:
            import warnings

            with warnings.catch_warnings():
                warnings.simplefilter("ignore", RuntimeWarning)
                nc.close()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    @property
    def positions(self):
        return self.read(range(self.n_frames))[0]

    @property
    def forces(self):
        return self.read(range(self.n_frames))[1]
