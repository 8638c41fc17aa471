"""Force / torque covariance of one molecule on the B200 path.

Drop-in for `waterEntropy/recipes/forces_torques.py:14-55` (`get_forces_torques`, the
single-united-atom branch that water takes): same positional arguments, mutates `covariances`.
The arithmetic - centre of mass, inertia tensor, LAPACK-faithful principal axes, rotated
mass-weighted force, inertia-weighted torque, the two outer products - runs in the CUDA kernel
`k_force_torque` (the device function `k_cov` uses); nothing is computed on the host.
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


def get_forces_torques(covariances, molecule, nearest: str, system):
    """Reference: `recipes/forces_torques.py:14-55`.  `molecule`: an atom group (`.indices`, `.resnames`;
    a bare index list is accepted too) of ONE molecule in the current frame of `system`; its force and
    torque covariance matrices are folded into `covariances` under `(nearest, molecule resname)`."""
    idx = np.asarray(getattr(molecule, "indices", molecule), dtype=np.int64)
    masses = np.asarray(system.atoms.masses, dtype=np.float64)[idx]
    heavy = int(((masses >= 2.0) & (masses <= 999.0)).sum())
    if heavy != 1:
        # utils/selections.py:89-105 -> "multiple_UAs": the bonded-axes branch of forces_torques.py:58-93,
        # "not applicable to water" upstream and not part of this path
        raise NotImplementedError("only single-united-atom molecules (water) are supported on the B200 path")
    pos = np.asarray(system.atoms.positions, dtype=np.float32)[idx]
    frc = np.asarray(system.atoms.forces, dtype=np.float32)[idx]
    r = molecule_force_torque(pos, frc, masses)
    resnames = getattr(molecule, "resnames", None)
    name = resnames[0] if resnames is not None else system.atoms.resnames[idx[0]]
    covariances.add_data(nearest, name, r["Fcov"][0], r["Tcov"][0])
