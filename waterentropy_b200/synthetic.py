"""Deterministic synthetic solvated systems for the BASELINE.json configurations.

The reference ships no trajectory large enough for the headline metric (its
only usable fixture has 900 waters, `tests/input_files/amber/arginine_solution`),
so the bench and the large parity tests run on synthetic TIP3P boxes with an
optional protein-like solute:

* waters: rigid TIP3P (r_OH 0.9572 A, HOH 104.52 deg), atom order O,H1,H2,
  bonds O-H1, O-H2, H1-H2 (as in the reference's AMBER fixture), masses
  16.0/1.008/1.008, charges -0.834/+0.417, resname WAT, types OW/HW;
  O sites on a lattice (fractional grid of the cell) with a per-frame rigid
  displacement N(0, 0.35 A)^3 and a fresh uniform random rotation; molecules
  are written whole (the reference never unwraps, recipes/forces_torques.py:40);
* solute: chains of residues on a 5 A self-avoiding lattice snake; every residue
  has 5-7 heavy atoms with their hydrogens *after* them in index order (donor
  rule `analysis/HB.py:111-115`), globally unique resids;
* forces: float32 N(0, 60 kJ mol^-1 A^-1)^3 per atom;
* box: orthorhombic or triclinic `[L, L, L, 60, 60, 90]`.

Every frame is a pure function of (seed, frame index), so ranks can generate
their own shard and the CPU baseline can regenerate any sub-sample.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

WATER_RESNAMES = frozenset(
    "H2O HOH OH2 HHO OHH TIP T3P T4P T5P SOL WAT TIP2 TIP3 TIP4".split()
)

_AMINO = "ALA ARG ASN ASP CYS GLN GLU GLY HIS ILE LEU LYS MET PHE PRO SER THR TRP TYR VAL".split()
_NUCLEO = "DA DC DG DT".split()

_R_OH = 0.9572
_HOH = np.deg2rad(104.52)
# TIP3P template, O at origin, molecule in the xy plane
_WATER_TEMPLATE = np.array(
    [
        [0.0, 0.0, 0.0],
        [_R_OH * np.sin(_HOH / 2), _R_OH * np.cos(_HOH / 2), 0.0],
        [-_R_OH * np.sin(_HOH / 2), _R_OH * np.cos(_HOH / 2), 0.0],
    ]
)

# residue template: (name, type, mass, charge, parent index in template or -1, offset)
# hydrogens directly follow the heavy atom they are bonded to.
_RES_TEMPLATE = [
    ("N", "N", 14.01, -0.4157, -1, (-1.20, 0.30, 0.00)),
    ("H", "H", 1.008, 0.2719, 0, (-1.60, 1.22, 0.10)),
    ("CA", "CX", 12.01, 0.0337, 0, (0.00, -0.45, 0.00)),
    ("HA", "H1", 1.008, 0.0823, 2, (0.05, -1.10, 0.88)),
    ("CB", "CT", 12.01, -0.1825, 2, (0.10, -1.35, -1.22)),
    ("HB2", "HC", 1.008, 0.0603, 4, (1.05, -1.90, -1.25)),
    ("HB3", "HC", 1.008, 0.0603, 4, (-0.75, -2.05, -1.30)),
    ("CG", "CT", 12.01, -0.0100, 4, (0.05, -0.55, -2.52)),
    ("HG", "HC", 1.008, 0.0400, 7, (0.90, 0.15, -2.60)),
    ("OG", "OH", 16.00, -0.6546, 7, (-1.15, 0.20, -2.70)),
    ("HO", "HO", 1.008, 0.4275, 9, (-1.20, 0.75, -3.50)),
    ("C", "C", 12.01, 0.5973, 2, (1.25, 0.40, 0.05)),
    ("O", "O", 16.00, -0.5679, 11, (1.35, 1.62, 0.10)),
]
# short residue = first 7 entries + C, O  (5 heavy atoms)
_SHORT = [0, 1, 2, 3, 4, 5, 6, 11, 12]


@dataclass
class SyntheticSystem:
    """Topology arrays + frame generator.  Duck-types the parts of an
    MDAnalysis Universe the hot path reads (see `waterentropy_b200.system`)."""

    names: list
    types: list
    resnames: list
    resids: np.ndarray
    masses: np.ndarray
    charges: np.ndarray
    bonds: np.ndarray
    dimensions: np.ndarray  # float32 [6]
    base: np.ndarray  # float64 [N,3] reference positions (solute) / O sites
    water_O: np.ndarray  # atom indices of water oxygens
    solute_atoms: np.ndarray
    seed: int
    sigma_water: float = 0.35
    sigma_solute: float = 0.15
    force_sigma: float = 60.0
    label: str = ""
    _matrix: np.ndarray = field(default=None, repr=False)

    @property
    def n_atoms(self):
        return len(self.masses)

    @property
    def n_waters(self):
        return len(self.water_O)

    def frame(self, f: int):
        """(positions float32 [N,3], forces float32 [N,3], dimensions float32 [6])"""
        rng = np.random.default_rng([self.seed, 7919, int(f)])
        n = self.n_atoms
        pos = np.empty((n, 3), dtype=np.float64)
        if len(self.solute_atoms):
            pos[self.solute_atoms] = self.base[self.solute_atoms] + rng.normal(
                0.0, self.sigma_solute, (len(self.solute_atoms), 3)
            )
        W = len(self.water_O)
        if W:
            q = rng.normal(size=(W, 4))
            q /= np.linalg.norm(q, axis=1)[:, None]
            a, b, c, d = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
            R = np.empty((W, 3, 3))
            R[:, 0, 0] = a * a + b * b - c * c - d * d
            R[:, 0, 1] = 2 * (b * c - a * d)
            R[:, 0, 2] = 2 * (b * d + a * c)
            R[:, 1, 0] = 2 * (b * c + a * d)
            R[:, 1, 1] = a * a - b * b + c * c - d * d
            R[:, 1, 2] = 2 * (c * d - a * b)
            R[:, 2, 0] = 2 * (b * d - a * c)
            R[:, 2, 1] = 2 * (c * d + a * b)
            R[:, 2, 2] = a * a - b * b - c * c + d * d
            mol = np.einsum("wij,aj->wai", R, _WATER_TEMPLATE)
            centre = self.base[self.water_O] + rng.normal(0.0, self.sigma_water, (W, 3))
            mol += centre[:, None, :]
            idx = self.water_O[:, None] + np.arange(3)[None, :]
            pos[idx.reshape(-1)] = mol.reshape(-1, 3)
        frc = rng.normal(0.0, self.force_sigma, (n, 3)).astype(np.float32)
        return pos.astype(np.float32), frc, self.dimensions.copy()

    def frames(self, start: int, count: int):
        P = np.empty((count, self.n_atoms, 3), dtype=np.float32)
        F = np.empty((count, self.n_atoms, 3), dtype=np.float32)
        D = np.empty((count, 6), dtype=np.float32)
        for k in range(count):
            P[k], F[k], D[k] = self.frame(start + k)
        return P, F, D


def box_matrix(dimensions):
    """float64 lower-triangular cell matrix (rows = cell vectors)."""
    lx, ly, lz, al, be, ga = [float(v) for v in dimensions]
    ca, cb, cg = np.cos(np.deg2rad([al, be, ga]))
    sg = np.sin(np.deg2rad(ga))
    if al == 90.0:
        ca = 0.0
    if be == 90.0:
        cb = 0.0
    if ga == 90.0:
        cg, sg = 0.0, 1.0
    m = np.zeros((3, 3))
    m[0, 0] = lx
    m[1, 0] = ly * cg
    m[1, 1] = ly * sg
    m[2, 0] = lz * cb
    m[2, 1] = lz * (ca - cb * cg) / sg
    m[2, 2] = np.sqrt(lz * lz - m[2, 0] ** 2 - m[2, 1] ** 2)
    return m


def _snake(n_nodes: int, spacing: float, centre, rng):
    """Self-avoiding boustrophedon path through a cubic lattice ball."""
    side = int(np.ceil((n_nodes * 1.0) ** (1 / 3))) + 1
    pts = []
    for ix in range(side):
        ys = range(side) if ix % 2 == 0 else range(side - 1, -1, -1)
        for ny, iy in enumerate(ys):
            flip = (ix * side + ny) % 2 == 1
            zs = range(side) if not flip else range(side - 1, -1, -1)
            for iz in zs:
                pts.append((ix, iy, iz))
    pts = np.array(pts[:n_nodes], dtype=np.float64)
    pts -= pts.mean(axis=0)
    return pts * spacing + np.asarray(centre)[None, :]


def make_system(
    n_waters: int,
    n_residues: int = 0,
    n_chains: int = 1,
    triclinic: bool = False,
    seed: int = 20261019,
    nucleic_fraction: float = 0.0,
    density: float = 0.03344,
    label: str = "",
    n_ions: int = 0,
) -> SyntheticSystem:
    """Build a solvated system with exactly `n_waters` waters."""
    rng = np.random.default_rng([seed, 104729])
    # ---- solute -----------------------------------------------------------
    names, types, resnames, resids, masses, charges, bonds, base = [], [], [], [], [], [], [], []
    res_per_chain = [n_residues // max(n_chains, 1)] * max(n_chains, 1) if n_residues else []
    for k in range(n_residues - sum(res_per_chain)):
        res_per_chain[k] += 1
    solute_volume = n_residues * 125.0 * 1.3 + n_ions * 30.0
    vol = n_waters / density + solute_volume
    shape = np.sqrt(0.5) if triclinic else 1.0
    L = (vol / shape) ** (1 / 3)
    dims = np.array(
        [L, L, L, 60.0, 60.0, 90.0] if triclinic else [L, L, L, 90.0, 90.0, 90.0],
        dtype=np.float32,
    )
    M = box_matrix(dims)
    cell_centre = 0.5 * (M[0] + M[1] + M[2])
    resid = 0
    if n_residues:
        nodes = _snake(n_residues, 5.0, cell_centre, rng)
        node = 0
        for ch, nres in enumerate(res_per_chain):
            prev_c = None
            for r in range(nres):
                resid += 1
                nuc = rng.random() < nucleic_fraction
                rname = _NUCLEO[rng.integers(4)] if nuc else _AMINO[rng.integers(20)]
                tmpl = list(range(len(_RES_TEMPLATE))) if rng.random() < 0.6 else _SHORT
                q = rng.normal(size=4)
                q /= np.linalg.norm(q)
                a, b, c, d = q
                R = np.array(
                    [
                        [a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                        [2 * (b * c + a * d), a * a - b * b + c * c - d * d, 2 * (c * d - a * b)],
                        [2 * (b * d - a * c), 2 * (c * d + a * b), a * a - b * b - c * c + d * d],
                    ]
                )
                first = len(names)
                local = {}
                for t in tmpl:
                    nm, ty, ms, qq, parent, off = _RES_TEMPLATE[t]
                    local[t] = len(names)
                    names.append(nm)
                    types.append(ty)
                    resnames.append(rname)
                    resids.append(resid)
                    masses.append(ms)
                    charges.append(qq)
                    base.append(nodes[node] + R @ np.array(off))
                    if parent >= 0:
                        bonds.append((local[parent], local[t]))
                if prev_c is not None:
                    bonds.append((prev_c, first))  # peptide-like link C(i-1) - N(i)
                prev_c = local[11]
                node += 1
    n_solute_atoms = len(names)
    for k in range(n_ions):
        resid += 1
        names.append("Cl-")
        types.append("Cl-")
        resnames.append("Cl-")
        resids.append(resid)
        masses.append(35.45)
        charges.append(-1.0)
        # ions on a sparse shell around the solute / cell centre
        u = rng.normal(size=3)
        u /= np.linalg.norm(u)
        radius = (n_residues * 125.0 * 3 / (4 * np.pi)) ** (1 / 3) + 6.0 if n_residues else 0.25 * L
        base.append(cell_centre + u * radius * (1.0 + 0.3 * rng.random()))
    solute_atoms = np.arange(len(names), dtype=np.int64)
    solute_xyz = np.array(base, dtype=np.float64).reshape(-1, 3)
    heavy_xyz = (
        solute_xyz[np.array(masses) > 2.0] if len(names) else np.zeros((0, 3))
    )
    # ---- water lattice ----------------------------------------------------
    n_side = int(np.ceil((n_waters * (1.0 + solute_volume / vol) * 1.02 + 8) ** (1 / 3)))
    while True:
        g = (np.arange(n_side) + 0.5) / n_side
        frac = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3)
        sites = frac @ M
        if len(heavy_xyz):
            from scipy.spatial import cKDTree

            tree = cKDTree(heavy_xyz)
            dmin, _ = tree.query(sites, k=1)
            sites = sites[dmin > 2.9]
        if len(sites) >= n_waters:
            break
        n_side += 1
    keep = np.sort(rng.choice(len(sites), size=n_waters, replace=False))
    sites = sites[keep]
    first_water = len(names)
    resid0 = resid
    W = n_waters
    names += ["O", "H1", "H2"] * W
    types += ["OW", "HW", "HW"] * W
    resnames += ["WAT"] * (3 * W)
    resids = np.concatenate(
        [np.array(resids, dtype=np.int64), resid0 + 1 + np.repeat(np.arange(W), 3)]
    )
    masses = np.concatenate([np.array(masses, dtype=np.float64), np.tile([16.0, 1.008, 1.008], W)])
    charges = np.concatenate(
        [np.array(charges, dtype=np.float64), np.tile([-0.834, 0.417, 0.417], W)]
    )
    water_O = first_water + 3 * np.arange(W, dtype=np.int64)
    wb = np.stack(
        [
            np.stack([water_O, water_O + 1], axis=1),
            np.stack([water_O, water_O + 2], axis=1),
            np.stack([water_O + 1, water_O + 2], axis=1),
        ],
        axis=1,
    ).reshape(-1, 2)
    bonds = np.concatenate([np.array(bonds, dtype=np.int64).reshape(-1, 2), wb])
    base_all = np.zeros((len(names), 3))
    if len(solute_atoms):
        base_all[solute_atoms] = solute_xyz
    base_all[water_O] = sites
    return SyntheticSystem(
        names=names,
        types=types,
        resnames=resnames,
        resids=resids,
        masses=masses,
        charges=charges,
        bonds=bonds,
        dimensions=dims,
        base=base_all,
        water_O=water_O,
        solute_atoms=solute_atoms,
        seed=seed,
        label=label,
        _matrix=M,
    )


# BASELINE.json configs 2-5 (config 1 is the reference's own lysozyme example)
CONFIGS = {
    "C2": dict(n_waters=10_000, n_residues=0, triclinic=False, frames=1_000,
               label="C2 pure TIP3P, 10k waters x 1k frames"),
    "C3": dict(n_waters=30_000, n_residues=129, n_chains=1, n_ions=8, triclinic=False, frames=5_000,
               label="C3 solvated lysozyme-size protein, 30k waters x 5k frames"),
    "C4": dict(n_waters=100_000, n_residues=1_250, n_chains=4, n_ions=40, triclinic=True,
               nucleic_fraction=0.35, frames=10_000,
               label="C4 solvated protein-DNA complex, 100k waters x 10k frames, triclinic"),
    "C5": dict(n_waters=300_000, n_residues=6_000, n_chains=12, n_ions=200, triclinic=False, frames=2_000,
               label="C5 1M-atom membrane/protein, 300k waters x 2k frames"),
}


def make_config(name: str, seed_offset: int = 0, scale: float = 1.0) -> SyntheticSystem:
    """Build one of the named BASELINE.json configurations (optionally scaled
    down by `scale` in waters and residues for quick tests)."""
    cfg = dict(CONFIGS[name])
    cfg.pop("frames")
    cfg["n_waters"] = max(8, int(round(cfg["n_waters"] * scale)))
    cfg["n_residues"] = int(round(cfg.get("n_residues", 0) * scale))
    cfg["n_ions"] = int(round(cfg.get("n_ions", 0) * scale))
    seed = 20261019 + int(name[1:]) + seed_offset
    return make_system(seed=seed, **cfg)
