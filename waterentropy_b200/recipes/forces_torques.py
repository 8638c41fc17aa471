"""Force / torque covariance of one molecule on the B200 path.

Drop-in for `waterEntropy/recipes/forces_torques.py:14-93` (`get_forces_torques`): same positional
arguments, mutates `covariances`.  The arithmetic - centre of mass, inertia tensor, LAPACK-faithful
principal axes, rotated mass-weighted force, inertia-weighted torque, the outer products, and for
molecules with several united atoms the bonded-axes frames of `maths/transformations.py:166-353` -
runs in the CUDA kernels `k_force_torque` (the device function `k_cov` uses) and
`k_force_torque_multi`; the host only gathers the atoms and the bond lists.
"""
from __future__ import annotations

import numpy as np

from waterentropy_b200 import _lib


def _device():
    try:
        import torch

        if torch.cuda.is_available():
            return torch.cuda.current_device()
    except Exception:  # pragma: no cover
        pass
    return 0


def molecule_force_torque(positions, forces, masses, mol_ptr=None, eig_flavour=None, device=None):
    """Rotated mass-weighted forces, inertia-weighted torques and their covariance matrices for molecules
    given as compact arrays (`positions` / `forces` float32 [n, 3], `masses` [n], `mol_ptr` CSR, default
    one molecule).  Returns dict(F [m,3], T [m,3], Fcov [m,3,3], Tcov [m,3,3])."""
    L = _lib.load()
    pos = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
    frc = np.ascontiguousarray(forces, dtype=np.float32).reshape(-1, 3)
    m = np.ascontiguousarray(masses, dtype=np.float64)
    if mol_ptr is None:
        mol_ptr = [0, len(m)]
    ptr = np.ascontiguousarray(mol_ptr, dtype=np.int32)
    if len(pos) != len(m) or len(frc) != len(m) or ptr[0] != 0 or ptr[-1] != len(m):
        raise ValueError("positions, forces, masses and mol_ptr disagree")
    flavour = _lib.host_eig_flavour()[0] if eig_flavour is None else _lib._EIG_NAMES[str(eig_flavour).lower()]
    n = len(ptr) - 1
    out = np.empty((n, 24), dtype=np.float64)
    _lib.check(L.we_molecule_force_torque(_device() if device is None else device, pos.ctypes.data, frc.ctypes.data,
                                          m.ctypes.data, ptr.ctypes.data, n, flavour, out.ctypes.data))
    return dict(F=out[:, 0:3].copy(), T=out[:, 3:6].copy(), Fcov=out[:, 6:15].reshape(n, 3, 3).copy(),
                Tcov=out[:, 15:24].reshape(n, 3, 3).copy())


def _adjacency(top):
    """sorted neighbour lists of every atom (CSR), cached on the topology"""
    adj = getattr(top, "_we_adjacency", None)
    if adj is None:
        b = np.asarray(top.bonds, dtype=np.int64).reshape(-1, 2)
        both = np.concatenate([b, b[:, ::-1]]) if len(b) else np.zeros((0, 2), dtype=np.int64)
        order = np.lexsort((both[:, 1], both[:, 0]))
        both = both[order]
        ptr = np.zeros(top.n_atoms + 1, dtype=np.int64)
        np.add.at(ptr, both[:, 0] + 1, 1)
        adj = (np.cumsum(ptr), both[:, 1].copy())
        top._we_adjacency = adj
    return adj


def multi_ua_force_torque(system, molecules, eig_flavour=None, device=None):
    """`recipes/forces_torques.py:14-93` for molecules with more than one united atom (`multiple_UAs`), in the
    current frame of `system`: `molecules` = list of atom-index arrays.  Runs in the CUDA kernel
    `k_force_torque_multi`.  Returns a list of dict(F, T, Fcov, Tcov (whole molecule); uas (atom indices of the
    united atoms that have a bonded-axes frame), ua_F, ua_T, ua_Fcov [k,3,3], ua_Tcov [k,3,3])."""
    from waterentropy_b200.system import SystemTopology

    L = _lib.load()
    top = SystemTopology.from_universe(system)
    ptr, nbr = _adjacency(top)
    masses = top.masses
    heavy = (masses >= 2.0) & (masses <= 999.0)          # "mass 2 to 999"
    light = (masses >= 1.0) & (masses <= 1.1)            # "mass 1 to 1.1"
    mols = [np.unique(np.asarray(m, dtype=np.int64)) for m in molecules]
    pool = set()
    for m in mols:
        pool.update(m.tolist())
        for a in m[heavy[m]].tolist():
            pool.update(nbr[ptr[a]:ptr[a + 1]].tolist())
    pool = np.array(sorted(pool), dtype=np.int64)
    where = {int(a): k for k, a in enumerate(pool)}
    mol_ptr, mol_atoms, ua_ptr, ua_atom = [0], [], [0], []
    hv_ptr, hv, lt_ptr, lt, hin_ptr, hin = [0], [], [0], [], [0], []
    for m in mols:
        inside = set(m.tolist())
        mol_atoms.extend(where[a] for a in m.tolist())
        mol_ptr.append(len(mol_atoms))
        for a in m[heavy[m]].tolist():
            ua_atom.append(where[a])
            nb = nbr[ptr[a]:ptr[a + 1]]
            hv.extend(where[b] for b in nb[heavy[nb]].tolist())
            lt.extend(where[b] for b in nb[light[nb]].tolist())
            hin.extend(where[b] for b in nb[light[nb]].tolist() if b in inside)
            hv_ptr.append(len(hv)); lt_ptr.append(len(lt)); hin_ptr.append(len(hin))
        ua_ptr.append(len(ua_atom))
    i32 = lambda x: np.ascontiguousarray(x if len(x) else [0], dtype=np.int32)  # noqa: E731
    pos = np.ascontiguousarray(np.asarray(system.atoms.positions, dtype=np.float32)[pool])
    frc = np.ascontiguousarray(np.asarray(system.atoms.forces, dtype=np.float32)[pool])
    mass = np.ascontiguousarray(masses[pool], dtype=np.float64)
    box3 = np.ascontiguousarray(np.asarray(system.dimensions, dtype=np.float32)[:3])
    arrs = [i32(x) for x in (mol_ptr, mol_atoms, ua_ptr, ua_atom, hv_ptr, hv, lt_ptr, lt, hin_ptr, hin)]
    n_mol, n_ua = len(mols), len(ua_atom)
    out_mol = np.empty((n_mol, 24), dtype=np.float64)
    out_ua = np.empty((max(n_ua, 1), 25), dtype=np.float64)
    flavour = _lib.host_eig_flavour()[0] if eig_flavour is None else _lib._EIG_NAMES[str(eig_flavour).lower()]
    _lib.check(L.we_molecule_force_torque_multi(
        _device() if device is None else device, pos.ctypes.data, frc.ctypes.data, mass.ctypes.data, len(pool),
        arrs[0].ctypes.data, arrs[1].ctypes.data, n_mol, arrs[2].ctypes.data, arrs[3].ctypes.data,
        arrs[4].ctypes.data, arrs[5].ctypes.data, arrs[6].ctypes.data, arrs[7].ctypes.data, arrs[8].ctypes.data,
        arrs[9].ctypes.data, box3.ctypes.data, flavour, out_mol.ctypes.data, out_ua.ctypes.data))
    res = []
    for k in range(n_mol):
        u0, u1 = ua_ptr[k], ua_ptr[k + 1]
        rows = out_ua[u0:u1]
        keep = rows[:, 24] != 0.0 if u1 > u0 else np.zeros(0, dtype=bool)
        rows = rows[keep]
        res.append(dict(
            F=out_mol[k, 0:3].copy(), T=out_mol[k, 3:6].copy(), Fcov=out_mol[k, 6:15].reshape(3, 3).copy(),
            Tcov=out_mol[k, 15:24].reshape(3, 3).copy(),
            uas=pool[np.asarray(ua_atom[u0:u1], dtype=np.int64)][keep] if u1 > u0 else np.zeros(0, dtype=np.int64),
            ua_F=rows[:, 0:3].copy(), ua_T=rows[:, 3:6].copy(),
            ua_Fcov=rows[:, 6:15].reshape(-1, 3, 3).copy(), ua_Tcov=rows[:, 15:24].reshape(-1, 3, 3).copy()))
    return res


def get_forces_torques(covariances, molecule, nearest: str, system):
    """Reference: `recipes/forces_torques.py:14-93`.  `molecule`: an atom group (`.indices`, `.resnames`;
    a bare index list is accepted too) of ONE molecule in the current frame of `system`.
    * one united atom (water): force / torque covariance matrices, halved, under `(nearest, resname)`;
    * `multiple_UAs` (several heavy atoms, one residue): whole-molecule matrices (not halved) under
      `(nearest, resname)`, then the stacked bonded-axes matrices of its united atoms under
      `(molecule, molecule)` exactly as upstream (`:87-93`);
    * `polymer` (several residues): whole-molecule matrices only, like upstream."""
    idx = np.asarray(getattr(molecule, "indices", molecule), dtype=np.int64)
    masses = np.asarray(system.atoms.masses, dtype=np.float64)[idx]
    heavy = int(((masses >= 2.0) & (masses <= 999.0)).sum())
    resnames = getattr(molecule, "resnames", None)
    name = resnames[0] if resnames is not None else system.atoms.resnames[idx[0]]
    if heavy == 0:
        # utils/selections.py:89-105 -> "no_UA": upstream then fails inside get_axes with an unbound scale
        raise ValueError("molecule has no united atom (no atom with 2 <= mass <= 999)")
    if heavy == 1:
        pos = np.asarray(system.atoms.positions, dtype=np.float32)[idx]
        frc = np.asarray(system.atoms.forces, dtype=np.float32)[idx]
        r = molecule_force_torque(pos, frc, masses)
        covariances.add_data(nearest, name, r["Fcov"][0], r["Tcov"][0])
        return
    n_res = len(np.unique(np.asarray(system.atoms.resids)[idx]))
    if n_res > 1:
        # "polymer": get_axes keeps the all-atom inertia tensor for the moments (transformations.py:113-119)
        # and no united-atom frames are built; the whole-molecule kernel path with all-atom moments
        pos = np.asarray(system.atoms.positions, dtype=np.float32)[np.sort(idx)]
        frc = np.asarray(system.atoms.forces, dtype=np.float32)[np.sort(idx)]
        m = np.asarray(system.atoms.masses, dtype=np.float64)[np.sort(idx)]
        w = molecule_force_torque(pos, frc, m)
        covariances.add_data(nearest, name, w["Fcov"][0] * 4.0, w["Tcov"][0] * 4.0)  # scale_covariance = 1
        return
    r = multi_ua_force_torque(system, [idx])[0]
    covariances.add_data(nearest, name, r["Fcov"], r["Tcov"])
    if len(r["uas"]) == 0:
        raise ValueError("need at least one array to concatenate")  # np.concatenate([]) upstream (:88-92)
    try:
        hash(molecule)
        key = molecule            # upstream keys this entry by the AtomGroup itself (:87-93)
    except TypeError:
        key = tuple(int(i) for i in idx)   # a bare index list / array was handed in
    covariances.add_data(key, key, r["ua_Fcov"], r["ua_Tcov"])
