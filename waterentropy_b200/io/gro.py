"""GROMACS GRO coordinates + a template topology.

BASELINE.json's first configuration is the reference's own example,
`example_inputs/lysozyme_solution/system.gro` (33,945 atoms: lysozyme, 10,659 SOL waters, 8 CL), which
upstream loads with `MDAnalysis.Universe(system.tpr, system.trr)` (`cli/runWaterEntropy.py:37`).  Its
force trajectory is not shipped (`.MISSING_LARGE_BLOBS`) and a TPR file needs MDAnalysis' parser, so the
one frame that exists - the GRO file - is read here.  A GRO file has names, residues, positions and the
box, but no masses, charges, atom types or bonds; the template below supplies what the SHELLS path needs
(`get_interfacial_shells`, `cli/run_find_Nc.py`: heavy-atom selection, bonded exclusions, fragments,
like / non-like):

* masses from the element guessed from the atom name (what MDAnalysis' guessers do for formats
  without masses),
* atom types = atom names (a TPR would carry force-field types; what matters to the path is only that
  water oxygens and solute atoms differ, `analysis/shell_labels.py:102`),
* bonds: O-H1, O-H2 for every water residue; distance-based for everything else, the MDAnalysis
  `guess_bonds` criterion d < 0.55 (r_i + r_j) on the first frame with minimum-image distances.

Charges are NOT guessed (they decide hydrogen bonds, `analysis/HB.py:85-137`): a Universe built from a
GRO topology reports `has_charges = False` and the entropy recipes refuse it.  Units as MDAnalysis:
nm -> Angstrom.
"""
from __future__ import annotations

import numpy as np

from ..system import WATER_RESNAMES

# element -> (mass, van der Waals radius used by the bond guess); MDAnalysis tables
_ELEMENTS = {
    "H": (1.008, 1.10), "C": (12.011, 1.70), "N": (14.007, 1.55), "O": (15.999, 1.52), "S": (32.06, 1.80),
    "P": (30.974, 1.80), "F": (18.998, 1.47), "CL": (35.45, 1.75), "BR": (79.904, 1.85), "I": (126.904, 1.98),
    "NA": (22.99, 2.27), "K": (39.098, 2.75), "MG": (24.305, 1.73), "CA": (40.078, 2.31), "ZN": (65.38, 1.39),
    "FE": (55.845, 1.63), "CU": (63.546, 1.40), "MN": (54.938, 1.61), "LI": (6.94, 1.82), "SE": (78.971, 1.90),
}
_ION_RESNAMES = {"CL", "CL-", "NA", "NA+", "K", "K+", "MG", "CA", "ZN", "LI", "BR", "I", "F", "SOD", "CLA", "POT", "CAL"}


def guess_element(name: str, resname: str = "") -> str:
    """Element from a PDB/GRO atom name: leading digits dropped; two-letter symbols only for ions (their
    residue is the atom) and a few unambiguous names - 'CA' in a protein is an alpha carbon."""
    n = "".join(ch for ch in name.upper() if ch.isalpha())
    r = resname.upper().strip("+-0123456789")
    if not n:
        return "X"
    if (resname.upper() in _ION_RESNAMES or r == n) and n[:2] in _ELEMENTS:
        return n[:2]
    if n[:2] in ("BR", "SE", "ZN", "FE", "MG", "MN", "CU", "LI") and n[:2] in _ELEMENTS and len(n) == 2 and resname.upper() in _ION_RESNAMES:
        return n[:2]
    return n[0] if n[0] in _ELEMENTS else "X"


class GroFile:
    """Frames of a GRO file (one or more, concatenated): positions float32 [F, N, 3] in Angstrom,
    dimensions float32 [F, 6]; names / resnames / resids from the first frame."""

    def __init__(self, path: str):
        with open(path, "r") as fh:
            lines = fh.read().splitlines()
        frames, dims = [], []
        i = 0
        self.names = self.resnames = self.resids = None
        while i + 2 < len(lines) + 0 and i < len(lines):
            if not lines[i].strip() and i + 1 >= len(lines):
                break
            n = int(lines[i + 1].split()[0])
            body = lines[i + 2 : i + 2 + n]
            if len(body) != n or i + 2 + n >= len(lines):
                raise ValueError(f"{path}: truncated frame at line {i + 1}")
            # fixed columns: resid(5) resname(5) name(5) number(5) then x y z with a common field width
            first = body[0]
            dec = [k for k, ch in enumerate(first[20:]) if ch == "."]
            width = dec[1] - dec[0] if len(dec) > 1 else 8
            xyz = np.empty((n, 3), dtype=np.float64)
            for k, ln in enumerate(body):
                xyz[k, 0] = float(ln[20 : 20 + width])
                xyz[k, 1] = float(ln[20 + width : 20 + 2 * width])
                xyz[k, 2] = float(ln[20 + 2 * width : 20 + 3 * width])
            if self.names is None:
                self.names = [ln[10:15].strip() for ln in body]
                self.resnames = [ln[5:10].strip() for ln in body]
                raw = np.array([int(ln[0:5]) for ln in body], dtype=np.int64)
                # the 5-digit residue number wraps at 100000: unwrap like MDAnalysis' GRO parser
                wraps = np.concatenate([[0], np.cumsum(np.diff(raw) < -50000)])
                self.resids = raw + 100000 * wraps
            frames.append((xyz * 10.0).astype(np.float32))
            box = [float(v) for v in lines[i + 2 + n].split()]
            dims.append(_gro_box(box))
            i += n + 3
        if not frames:
            raise ValueError(f"{path}: no frames")
        self.positions = np.stack(frames)
        self.dimensions = np.stack(dims).astype(np.float32)
        self.n_frames, self.n_atoms = self.positions.shape[:2]
        self.has_forces = False

    def read(self, indices, out_pos=None, out_frc=None):
        idx = np.asarray(list(indices), dtype=np.int64)
        if out_frc is not None:
            raise ValueError("a GRO file holds no forces; the covariance path needs them")
        if out_pos is not None:
            out_pos[: len(idx)] = self.positions[idx]
            return out_pos[: len(idx)], None, self.dimensions[idx]
        return self.positions[idx], None, self.dimensions[idx]


def _gro_box(box):
    """GRO box line (v1x v2y v3z [v1y v1z v2x v2z v3x v3y], nm) -> [lx, ly, lz, alpha, beta, gamma] (A, deg)."""
    if len(box) == 3:
        return np.array([box[0] * 10, box[1] * 10, box[2] * 10, 90.0, 90.0, 90.0])
    from .gromacs import _dimensions

    v1 = np.array([box[0], box[3], box[4]]) * 10
    v2 = np.array([box[5], box[1], box[6]]) * 10
    v3 = np.array([box[7], box[8], box[2]]) * 10
    return _dimensions(np.stack([v1, v2, v3]))


def guess_bonds(positions, dimensions, elements, resnames, resids):
    """[n_bonds, 2] int64.  Water residues: O bonded to its hydrogens.  Everything else: pairs closer than
    0.55 (r_i + r_j) (minimum image), hydrogens keeping only their nearest partner."""
    from scipy.spatial import cKDTree

    n = len(elements)
    is_water = np.array([r in WATER_RESNAMES for r in resnames])
    bonds = []
    # waters: heavy atom of the residue - its hydrogens
    w = np.nonzero(is_water)[0]
    if len(w):
        rid = np.asarray(resids)[w]
        start = np.concatenate([[0], np.nonzero(np.diff(rid) != 0)[0] + 1, [len(w)]])
        el = np.array(elements, dtype=object)
        for a, b in zip(start[:-1], start[1:]):
            atoms = w[a:b]
            heavy = [i for i in atoms if el[i] != "H"]
            if len(heavy) == 1:
                bonds.extend((heavy[0], i) for i in atoms if el[i] == "H")
    rest = np.nonzero(~is_water)[0]
    if len(rest) > 1:
        radii = np.array([_ELEMENTS.get(e, (0.0, 1.5))[1] for e in elements])
        box = np.asarray(dimensions[:3], dtype=np.float64)
        ortho = bool(np.all(np.asarray(dimensions[3:]) == 90.0))
        p = np.asarray(positions, dtype=np.float64)[rest]
        if ortho:
            tree = cKDTree(np.mod(p, box), boxsize=box)
        else:  # solutes are whole molecules in practice; without an orthorhombic box no image search is done
            tree = cKDTree(p)
        pairs = tree.query_pairs(0.55 * 2 * radii[rest].max(), output_type="ndarray")
        if len(pairs):
            i, j = rest[pairs[:, 0]], rest[pairs[:, 1]]
            d = p[pairs[:, 0]] - p[pairs[:, 1]]
            if ortho:
                d -= box * np.round(d / box)
            dist = np.sqrt((d * d).sum(axis=1))
            ok = dist < 0.55 * (radii[i] + radii[j])
            el = np.array(elements, dtype=object)
            ok &= ~((el[i] == "H") & (el[j] == "H"))
            i, j, dist = i[ok], j[ok], dist[ok]
            # a hydrogen has one bond: keep its shortest
            keep = np.ones(len(i), dtype=bool)
            for side in (i, j):
                hmask = el[side] == "H"
                order = np.lexsort((dist, side))
                s_sorted = side[order]
                firsts = np.ones(len(order), dtype=bool)
                firsts[1:] = s_sorted[1:] != s_sorted[:-1]
                drop = order[~firsts & hmask[order]]
                keep[drop] = False
            bonds.extend(zip(i[keep].tolist(), j[keep].tolist()))
    if not bonds:
        return np.zeros((0, 2), dtype=np.int64)
    b = np.array(bonds, dtype=np.int64)
    return np.unique(np.sort(b, axis=1), axis=0)


class GroTopology:
    """Template topology for the atoms of a GRO file (see the module docstring)."""

    def __init__(self, names, resnames, resids, positions0, dimensions0):
        self.names = [str(x) for x in names]
        self.resnames = [str(x) for x in resnames]
        self.resids = np.asarray(resids, dtype=np.int64)
        self.n_atoms = len(self.names)
        self.elements = [guess_element(a, r) for a, r in zip(self.names, self.resnames)]
        self.masses = np.array([_ELEMENTS.get(e, (0.0, 0.0))[0] for e in self.elements], dtype=np.float64)
        self.types = list(self.names)
        self.charges = np.zeros(self.n_atoms, dtype=np.float64)
        self.has_charges = False
        self.bonds = guess_bonds(positions0, dimensions0, self.elements, self.resnames, self.resids)
