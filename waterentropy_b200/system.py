"""Frame-invariant topology tables for the device path.

The reference re-derives these every frame with MDAnalysis selection strings
(`utils/selections.py:8-105`, `recipes/interfacial_solvent.py:298-299`,
`maths/trig.py:59-62`, `analysis/HB.py:111-115`).  They depend on the topology
only, so they are computed once on the host and uploaded with `we_create`.

`SystemTopology.from_universe(u)` accepts anything that duck-types the parts
of an MDAnalysis Universe the hot path reads (`u.atoms.masses/.charges/.resids/
.resnames/.types`, `u.bonds` or `u.atoms.bonds` with `.indices`), including
`waterentropy_b200.io.Universe` and `waterentropy_b200.synthetic.SyntheticSystem`.
"""
from __future__ import annotations

import numpy as np

from . import _lib

WATER_RESNAMES = frozenset(
    "H2O HOH OH2 HHO OHH TIP T3P T4P T5P SOL WAT TIP2 TIP3 TIP4".split()
)  # MDAnalysis "water" selection keyword

RECIPE_INTERFACIAL = 0
RECIPE_BULK = 1
PREFIX = ["0_", "", "1_", "2_", "X_"]  # device label prefix codes


def _components(n, bonds):
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components

    if len(bonds) == 0:
        return np.arange(n, dtype=np.int64)
    a = coo_matrix((np.ones(len(bonds), dtype=np.int8), (bonds[:, 0], bonds[:, 1])), shape=(n, n))
    _, lab = connected_components(a, directed=False)
    # relabel so that fragments are numbered by their first atom
    first = np.full(lab.max() + 1, n, dtype=np.int64)
    np.minimum.at(first, lab, np.arange(n))
    order = np.argsort(first)
    remap = np.empty_like(order)
    remap[order] = np.arange(len(order))
    return remap[lab].astype(np.int64)


class SystemTopology:
    def __init__(self, masses, charges, resids, resnames, types, bonds, names=None):
        n = len(masses)
        self.n_atoms = n
        self.masses = np.ascontiguousarray(masses, dtype=np.float64)
        self.charges = np.ascontiguousarray(charges, dtype=np.float64)
        self.resids = np.ascontiguousarray(resids, dtype=np.int64)
        if np.abs(self.resids).max(initial=0) >= 2**31:
            raise ValueError("resids must fit in int32")
        resnames = np.asarray([str(r) for r in resnames], dtype=object)
        types = np.asarray([str(t) for t in types], dtype=object)
        self.names = None if names is None else [str(x) for x in names]
        self.resname_table, rid = np.unique(resnames.astype(str), return_inverse=True)
        self.type_table, tid = np.unique(types.astype(str), return_inverse=True)
        if len(self.resname_table) >= 8192:
            raise ValueError("more than 8191 distinct residue names")
        self.resname_table = [str(s) for s in self.resname_table]
        self.resname_id = np.ascontiguousarray(rid, dtype=np.int32)
        self.type_id = np.ascontiguousarray(tid, dtype=np.int32)
        water_ids = [i for i, r in enumerate(self.resname_table) if r in WATER_RESNAMES]
        self.is_water = np.ascontiguousarray(np.isin(self.resname_id, water_ids), dtype=np.uint8)
        b = np.asarray(bonds, dtype=np.int64).reshape(-1, 2)
        if len(b):
            b = np.unique(np.sort(b, axis=1), axis=0)  # MDAnalysis: (low, high), no duplicates
        self.bonds = np.ascontiguousarray(b, dtype=np.int32)
        self.fragment_id = _components(n, b)
        self.is_heavy = (self.masses >= 2.0) & (self.masses <= 999.0)  # "mass 2 to 999"
        self._derive()

    # ------------------------------------------------------------------
    @classmethod
    def from_universe(cls, u):
        if isinstance(u, SystemTopology):
            return u
        cached = getattr(u, "_we_topology", None)
        if cached is not None:
            return cached
        atoms = getattr(u, "atoms", u)
        if hasattr(u, "masses") and not hasattr(u, "atoms"):  # SyntheticSystem / array bundle
            top = cls(u.masses, u.charges, u.resids, u.resnames, u.types, u.bonds, getattr(u, "names", None))
        else:
            bonds_obj = getattr(u, "bonds", None)
            if bonds_obj is None:
                bonds_obj = atoms.bonds
            bonds = bonds_obj.indices if hasattr(bonds_obj, "indices") else bonds_obj
            names = getattr(atoms, "names", None)
            top = cls(atoms.masses, atoms.charges, atoms.resids, atoms.resnames, atoms.types, bonds, names)
        try:
            u._we_topology = top
        except Exception:  # pragma: no cover - objects that refuse attributes
            pass
        return top

    # ------------------------------------------------------------------
    def _derive(self):
        n = self.n_atoms
        frag = self.fragment_id
        heavy = self.is_heavy
        nonwater = self.is_water == 0
        # --- utils/selections.py:34-56  find_solute_molecules ---------------------
        n_frag = int(frag.max()) + 1 if n else 0
        frag_has_nonwater = np.zeros(n_frag, dtype=bool)
        frag_has_nonwater[frag[nonwater]] = True
        cand_resids = np.unique(self.resids[frag_has_nonwater[frag]])
        sel = nonwater & heavy
        rid, cnt = np.unique(self.resids[sel], return_counts=True)
        count_of = dict(zip(rid.tolist(), cnt.tolist()))
        solute = []
        for r in cand_resids.tolist():
            k = count_of.get(r, 0)
            if k > 1:
                solute.append(r)
            elif k == 1:
                a = np.nonzero(sel & (self.resids == r))[0][0]
                if np.floor(self.masses[a]) != 16:  # a lone non-oxygen heavy atom is a solute
                    solute.append(r)
        self.solute_resids = solute
        # --- recipes/interfacial_solvent.py:39-43: UAs of the whole fragments -------
        if solute:
            in_res = np.isin(self.resids, np.array(solute))
            frag_sel = np.zeros(n_frag, dtype=bool)
            frag_sel[frag[in_res]] = True
            ua = np.nonzero(frag_sel[frag] & heavy)[0]
            order = np.lexsort((ua, frag[ua]))
            self.solute_ua = np.ascontiguousarray(ua[order], dtype=np.int32)
        else:
            self.solute_ua = np.zeros(0, dtype=np.int32)
        # --- water centres and their molecules ------------------------------------
        centres = np.nonzero(heavy & (self.is_water == 1))[0]
        self.centres = np.ascontiguousarray(centres, dtype=np.int32)
        order = np.argsort(frag, kind="stable")
        frag_sorted = frag[order]
        starts = np.searchsorted(frag_sorted, np.arange(n_frag))
        ends = np.searchsorted(frag_sorted, np.arange(n_frag), side="right")
        heavy_per_frag = np.bincount(frag[heavy], minlength=n_frag)
        cf = frag[centres]
        if len(centres) and (heavy_per_frag[cf] != 1).any():
            bad = centres[np.nonzero(heavy_per_frag[cf] != 1)[0][0]]
            raise NotImplementedError(
                f"water atom {bad} belongs to a molecule with several united atoms; the per-frame recipes record "
                "single-UA solvent only (molecules with several united atoms go through "
                "recipes.forces_torques.get_forces_torques, reference recipes/forces_torques.py:58-93)"
            )
        sizes = (ends - starts)[cf]
        self.frag_ptr = np.ascontiguousarray(np.concatenate([[0], np.cumsum(sizes)]), dtype=np.int32)
        if len(centres):
            flat = np.repeat(starts[cf], sizes) + (np.arange(int(sizes.sum())) - np.repeat(self.frag_ptr[:-1], sizes))
            self.frag_atoms = np.ascontiguousarray(order[flat], dtype=np.int32)  # ascending atom index
        else:
            self.frag_atoms = np.zeros(0, dtype=np.int32)
        # --- covariance keys ------------------------------------------------------
        cls_ids = np.unique(self.resname_id[centres]) if len(centres) else np.zeros(0, dtype=np.int32)
        self.class_resname_ids = [int(c) for c in cls_ids]
        self.n_classes = max(1, len(self.class_resname_ids))
        self.centre_class = np.full(n, -1, dtype=np.int32)
        for k, c in enumerate(self.class_resname_ids):
            self.centre_class[(self.resname_id == c) & heavy & (self.is_water == 1)] = k
        if self.n_classes > 1:
            eligible = heavy.copy()
        else:
            eligible = heavy & ~np.isin(self.resname_id, self.class_resname_ids)
        keys = np.stack([self.resname_id[eligible].astype(np.int64), self.resids[eligible]], axis=1)
        self.bin_of_atom = np.full(n, -1, dtype=np.int32)
        if len(keys):
            uniq, inv = np.unique(keys, axis=0, return_inverse=True)
            self.bin_of_atom[np.nonzero(eligible)[0]] = inv.reshape(-1).astype(np.int32)
            self.pair_bins = [(int(a), int(b)) for a, b in uniq]  # (resname_id, resid)
        else:
            self.pair_bins = []
        self.n_pair_bins = max(1, len(self.pair_bins))
        self.heavy_idx = np.nonzero(heavy)[0].astype(np.int32)

    # ------------------------------------------------------------------
    def c_struct(self, recipe: int):
        """(we_topology, keepalive) for `we_create`."""
        t = _lib.we_topology()
        keep = []

        def ptr(a):
            keep.append(a)
            return a.ctypes.data

        t.n_atoms = self.n_atoms
        t.masses = ptr(self.masses)
        t.charges = ptr(self.charges)
        t.resid = ptr(np.ascontiguousarray(self.resids, dtype=np.int32))
        t.resname_id = ptr(self.resname_id)
        t.type_id = ptr(self.type_id)
        t.is_water = ptr(self.is_water)
        t.n_bonds = len(self.bonds)
        t.bonds = ptr(self.bonds if len(self.bonds) else np.zeros((1, 2), dtype=np.int32))
        t.n_solute_ua = len(self.solute_ua)
        t.solute_ua = ptr(self.solute_ua if len(self.solute_ua) else np.zeros(1, dtype=np.int32))
        t.n_centres = len(self.centres)
        t.centres = ptr(self.centres if len(self.centres) else np.zeros(1, dtype=np.int32))
        t.frag_ptr = ptr(self.frag_ptr)
        t.frag_atoms = ptr(self.frag_atoms if len(self.frag_atoms) else np.zeros(1, dtype=np.int32))
        t.bin_of_atom = ptr(self.bin_of_atom)
        t.centre_class = ptr(self.centre_class)
        t.n_pair_bins = self.n_pair_bins
        t.n_classes = self.n_classes
        t.recipe = recipe
        return t, keep

    # ---- decoding of device tables ------------------------------------------
    def label_string(self, code: int) -> str:
        return PREFIX[code >> 13] + self.resname_table[code & 0x1FFF]
