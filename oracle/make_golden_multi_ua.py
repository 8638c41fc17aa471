"""TEST INFRASTRUCTURE ONLY -- goldens of the `multiple_UAs` branch, from the UNMODIFIED reference.

    python oracle/make_golden_multi_ua.py        (build container: /root/reference is mounted)

`recipes/forces_torques.py:58-93` (bonded-axes frames of every united atom of a molecule with more than
one heavy atom) is "not applicable to water" upstream and no reference test reaches it.  This script
imports `/root/reference/waterEntropy` on top of `oracle/mda_shim`, rebuilds the Universes of the committed
input fixtures (`tests/golden/*_inputs.npz`) and calls the reference's `get_forces_torques` on single
residues / single-residue molecules of them, recording everything it puts into the CovarianceCollection:
the whole-molecule matrices (key `(nearest, resname)`) and the stacked united-atom matrices (key
`(molecule, molecule)`).  Output: `tests/golden/multi_ua_reference.json.gz`.
Only outputs of the reference are stored, none of its source.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path[:0] = [os.path.join(HERE, "mda_shim"), REPO, os.path.join(REPO, "tests"), "/root/reference"]
os.environ.setdefault("WE_SHIM_INLINE", "1")

import MDAnalysis  # noqa: E402  (the shim)
from waterEntropy.entropy.convariances import CovarianceCollection  # noqa: E402
from waterEntropy.recipes.forces_torques import get_forces_torques  # noqa: E402
import waterEntropy.utils.selections as Select  # noqa: E402

import golden_io  # noqa: E402
from make_golden import write_json  # noqa: E402

# (fixture, frames, how many molecules at most)
CASES = [("arginine", [0, 3, 9], 8), ("aspirin", [0, 4, 10], 8), ("synth_ortho", [0, 2], 12),
         ("synth_tric", [1], 12)]


def universe(inp):
    return MDAnalysis.Universe.from_arrays(
        list(inp["names"]), list(inp["types"]), inp["charges"], inp["masses"], inp["resids"],
        list(inp["resnames"]), inp["bonds"], inp["positions"], inp["forces"], inp["dimensions"])


def molecules(u):
    """single residues of the non-water part that the reference classifies `multiple_UAs`"""
    out = []
    sol = u.select_atoms("not water")
    for rid in sorted(set(int(r) for r in sol.resids)):
        mol = u.select_atoms(f"resid {rid}")
        if Select.guess_length_scale(mol) == "multiple_UAs":
            out.append(mol)
    return out


def main():
    cases = []
    for tag, frames, limit in CASES:
        inp = golden_io.load_inputs(tag)
        u = universe(inp)
        for fr in frames:
            u.trajectory[fr]
            for mol in molecules(u)[:limit]:
                cov = CovarianceCollection()
                with np.errstate(all="ignore"):
                    get_forces_torques(cov, mol, "NEAR", u)
                whole = ("NEAR", mol.resnames[0])
                rec = {
                    "fixture": tag, "frame": int(fr), "indices": [int(i) for i in mol.indices],
                    "whole_F": np.array(cov.forces[whole]), "whole_T": np.array(cov.torques[whole]),
                    "ua_F": np.array(cov.forces[(mol, mol)]), "ua_T": np.array(cov.torques[(mol, mol)]),
                    "counts": [int(cov.counts[whole]), int(cov.counts[(mol, mol)])],
                }
                cases.append(rec)
    # "polymer" molecules (several residues): whole-molecule matrices only, not halved
    polymers = []
    for tag, frames in (("arginine", [0, 5]), ("synth_ortho", [1])):
        inp = golden_io.load_inputs(tag)
        u = universe(inp)
        for fr in frames:
            u.trajectory[fr]
            for mol in u.select_atoms("not water").fragments[:3]:
                if Select.guess_length_scale(mol) != "polymer":
                    continue
                cov = CovarianceCollection()
                get_forces_torques(cov, mol, "NEAR", u)
                key = ("NEAR", mol.resnames[0])
                polymers.append({"fixture": tag, "frame": int(fr), "indices": [int(i) for i in mol.indices],
                                 "whole_F": np.array(cov.forces[key]), "whole_T": np.array(cov.torques[key]),
                                 "keys": len(cov.forces)})
    # ---- bonding cases no fixture molecule has: united atoms without a heavy neighbour (get_bonded_axes falls
    # back to the principal axes of the united atom, transformations.py:341-343) and a molecule none of whose
    # united atoms gets a frame (np.concatenate([]) raises, forces_torques.py:88-92) ----
    rng = np.random.default_rng(20261020)
    names = ["N", "H1", "H2", "H3", "O", "H", "C", "O", "P", "O1", "O2", "O3", "O4", "H4"]
    masses = [14.007, 1.008, 1.008, 1.008, 15.999, 1.008, 12.011, 15.999, 30.974, 15.999, 15.999, 15.999, 15.999, 1.008]
    resids = [1, 1, 1, 1, 1, 1, 2, 2, 3, 3, 3, 3, 3, 3]
    resnames = ["NHO"] * 6 + ["CMO"] * 2 + ["PHO"] * 6
    bonds = [(0, 1), (0, 2), (0, 3), (4, 5), (6, 7), (8, 9), (8, 10), (8, 11), (8, 12), (12, 13)]
    base = np.array([[5.0, 5.0, 5.0], [5.6, 5.7, 5.1], [4.3, 5.6, 5.2], [5.1, 4.2, 5.5], [8.0, 6.0, 5.5], [8.7, 6.5, 5.9],
                     [12.0, 12.0, 12.0], [13.1, 12.2, 11.9], [17.0, 3.0, 18.5], [18.2, 3.6, 19.1], [16.1, 4.1, 17.9],
                     [16.3, 2.1, 19.6], [17.6, 2.0, 17.4], [18.3, 2.4, 16.9]])
    P = np.stack([base + rng.normal(0, 0.15, base.shape) for _ in range(3)]).astype(np.float32)
    P[2] += np.float32(11.0)  # near / across the box faces: get_vector's wrapping matters
    Fr = rng.normal(0, 40.0, P.shape).astype(np.float32)
    D = np.tile(np.array([20.0, 19.0, 21.0, 90, 90, 90], dtype=np.float32), (3, 1))
    charges = np.zeros(len(names))
    np.savez_compressed(os.path.join(golden_io.GOLD, "odd_ua_inputs.npz"), names=np.array(names, dtype="U8"),
                        types=np.array(names, dtype="U8"), resnames=np.array(resnames, dtype="U8"),
                        resids=np.asarray(resids, dtype=np.int64), masses=np.asarray(masses), charges=charges,
                        bonds=np.asarray(bonds, dtype=np.int64), positions=P, forces=Fr, dimensions=D)
    u = universe(golden_io.load_inputs("odd_ua"))
    odd = []
    for fr in range(3):
        u.trajectory[fr]
        for rid in (1, 2, 3):
            mol = u.select_atoms(f"resid {rid}")
            assert Select.guess_length_scale(mol) == "multiple_UAs"
            cov = CovarianceCollection()
            rec = {"fixture": "odd_ua", "frame": fr, "indices": [int(i) for i in mol.indices]}
            try:
                with np.errstate(all="ignore"):
                    get_forces_torques(cov, mol, "NEAR", u)
                rec.update(ua_F=np.array(cov.forces[(mol, mol)]), ua_T=np.array(cov.torques[(mol, mol)]))
            except ValueError as exc:
                rec["raises"] = str(exc)
            whole = ("NEAR", mol.resnames[0])
            rec.update(whole_F=np.array(cov.forces[whole]), whole_T=np.array(cov.torques[whole]))
            odd.append(rec)
    write_json("multi_ua", {"cases": cases, "polymers": polymers, "odd": odd})
    print(len(odd), "odd united-atom molecules,", sum("raises" in r for r in odd), "raise")
    print(len(polymers), "polymer molecules")
    n_ua = sum(len(c["ua_F"]) for c in cases)
    print(f"{len(cases)} molecules, {n_ua} united-atom frames -> tests/golden/multi_ua_reference.json.gz")


if __name__ == "__main__":
    main()
