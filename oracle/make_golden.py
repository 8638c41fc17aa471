"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/* from the UNMODIFIED reference.

Run in the build container (where /root/reference is mounted):

    python oracle/make_golden.py

It imports `/root/reference/waterEntropy` on top of `oracle/mda_shim` (the
stand-in for the uninstallable MDAnalysis/dask) and records, for the
reference's own fixture and for three small synthetic systems,

  * the inputs (topology + float32 frames) as .npz, because /root/reference does
    not exist on the GPU box,
  * what the reference computes from them (.json): per-frame
    `_entropy_per_step` results, all-centre neighbour lists / RAD shells / HB
    acceptors, bulk-recipe results and the public-API return values.

The oracle (`oracle/we_oracle.py`) and the CUDA path are both tested against
these files.  None of the reference's source is copied; only its outputs.
"""
from __future__ import annotations

import gzip
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path[:0] = [os.path.join(HERE, "mda_shim"), REPO, "/root/reference"]
os.environ.setdefault("WE_SHIM_INLINE", "1")

import MDAnalysis  # noqa: E402  (the shim)
import waterEntropy.analysis.HB as HBond  # noqa: E402
import waterEntropy.analysis.RAD as RADShell  # noqa: E402
from waterEntropy.analysis.shells import ShellCollection  # noqa: E402
import waterEntropy.maths.trig as Trig  # noqa: E402
import waterEntropy.recipes.bulk_water as Bulk  # noqa: E402
import waterEntropy.recipes.interfacial_solvent as Inter  # noqa: E402

from waterentropy_b200 import synthetic  # noqa: E402

GOLD = os.path.join(REPO, "tests", "golden")


def enc(o):
    """JSON-encode nested dicts with tuple / int keys, tuples and arrays."""
    if isinstance(o, dict):
        return {"d": [[enc(k), enc(v)] for k, v in o.items()]}
    if isinstance(o, tuple):
        return {"t": [enc(v) for v in o]}
    if isinstance(o, np.ndarray):
        return {"a": o.tolist()}
    if isinstance(o, (list,)):
        return [enc(v) for v in o]
    if isinstance(o, (np.integer,)):
        return int(o)
    if isinstance(o, (np.floating,)):
        return float(o)
    if isinstance(o, (np.str_,)):
        return str(o)
    return o


def write_json(tag, obj):
    """gzip with mtime=0 so regenerating identical content gives identical bytes"""
    raw = json.dumps(enc(obj)).encode()
    with open(os.path.join(GOLD, f"{tag}_reference.json.gz"), "wb") as fh:
        with gzip.GzipFile(fileobj=fh, mode="wb", mtime=0, compresslevel=9) as gz:
            gz.write(raw)


def cov_dict(cov):
    return {
        "forces": {k: np.array(v) for k, v in cov.forces.items()},
        "torques": {k: np.array(v) for k, v in cov.torques.items()},
        "counts": dict(cov.counts),
    }


def lab_dict(lab):
    return {
        "labelled_shell_counts": lab.labelled_shell_counts,
        "resid_labelled_shell_counts": lab.resid_labelled_shell_counts,
    }


def vib_dict(vib):
    return {
        "translational_S": dict(vib.translational_S),
        "rotational_S": dict(vib.rotational_S),
        "translational_freq": dict(vib.translational_freq),
        "rotational_freq": dict(vib.rotational_freq),
    }


def all_centres(system, frame):
    """Neighbour lists, RAD shells and HB acceptors of every heavy atom."""
    system.trajectory[frame]
    heavy = system.select_atoms("mass 2 to 999")
    shells = ShellCollection()
    HBs = HBond.HBCollection()
    out = {"nbr_idx": {}, "nbr_dist": {}, "nbr_total": {}, "shell": {}, "acceptor": {}}
    for atom in heavy:
        idx, dist = Trig.get_sorted_neighbours(atom.index, system)
        out["nbr_idx"][atom.index] = [int(v) for v in idx[:25]]
        out["nbr_dist"][atom.index] = [float(v) for v in dist[:25]]
        out["nbr_total"][atom.index] = len(idx)
        shell = RADShell.get_RAD_shell(atom, system, shells, idx, dist)
        out["shell"][atom.index] = [int(v) for v in shell.UA_shell]
    for atom in heavy:
        shell = shells.find_shell(atom.index)
        HBond.get_shell_HB_acceptors(shell, system, HBs)
        acc = HBs.find_acceptor(atom.index)
        if acc:
            out["acceptor"][atom.index] = {int(d): int(a) for d, a in acc.items()}
    return out


def record(system, tag, step_frames, centre_frames, bulk_frames, api_runs):
    ref = {"per_step": {}, "centres": {}, "bulk": {}, "api": {}, "shells_api": {}}
    for f in step_frames:
        fsi, cov, lab, n = Inter._entropy_per_step((f, system))
        ref["per_step"][f] = {
            "frame_solvent_indices": fsi, "cov": cov_dict(cov), "labels": lab_dict(lab), "n": n,
        }
    for f in centre_frames:
        ref["centres"][f] = all_centres(system, f)
    for f in bulk_frames:
        so, cov, vib = Bulk.get_bulk_water_orient_entropy(system, f, f + 1, 1)
        ref["bulk"][f] = {"Sorient": so, "cov_scaled": cov_dict(cov), "vib": vib_dict(vib)}
    for (s, e, st) in api_runs:
        so, cov, vib, fsi, n = Inter.get_interfacial_water_orient_entropy(system, s, e, st)
        key = f"{s}:{e}:{st}"
        ref["api"][key] = {
            "Sorient": so, "cov_scaled": cov_dict(cov), "vib": vib_dict(vib),
            "frame_solvent_indices": fsi, "n_frames": n,
        }
        ref["shells_api"][key] = Inter.get_interfacial_shells(system, s, e, st)
    write_json(tag, ref)
    print(tag, "reference outputs written")


def save_inputs(tag, names, types, resnames, resids, masses, charges, bonds, P, F, D):
    np.savez_compressed(
        os.path.join(GOLD, f"{tag}_inputs.npz"),
        names=np.array(names, dtype="U8"), types=np.array(types, dtype="U8"),
        resnames=np.array(resnames, dtype="U8"), resids=np.asarray(resids, dtype=np.int64),
        masses=np.asarray(masses, dtype=np.float64), charges=np.asarray(charges, dtype=np.float64),
        bonds=np.asarray(bonds, dtype=np.int64), positions=P, forces=F, dimensions=D,
    )


def main():
    os.makedirs(GOLD, exist_ok=True)
    # ---- the reference's own fixture -------------------------------------
    d = "/root/reference/tests/input_files/amber/arginine_solution/"
    u = MDAnalysis.Universe(d + "system.prmtop", d + "system.nc")
    top = u._top
    tr = u.trajectory
    save_inputs("arginine", top.names, top.types, top.resnames, top.resids, top.masses,
                top.charges, top.bonds, tr._pos, tr._frc, tr._dim)
    record(u, "arginine", step_frames=range(10), centre_frames=[0, 2, 7], bulk_frames=[0, 3],
           api_runs=[(0, 4, 2), (0, 10, 1)])
    # ---- the reference's second fixture (GROMACS TRR + PRMTOP; no reference test uses it) ----
    from waterentropy_b200.io.gromacs import load_gromacs

    d = "/root/reference/tests/input_files/gromacs/aspirin_solution/"
    names, types, resnames, resids, masses, charges, bonds, P, F, D = load_gromacs(d + "system.prmtop", d + "system.trr")
    save_inputs("aspirin", names, types, resnames, resids, masses, charges, bonds, P, F, D)
    u = MDAnalysis.Universe.from_arrays(names, types, charges, masses, resids, resnames, bonds, P, F, D)
    record(u, "aspirin", step_frames=range(0, 11, 2), centre_frames=[0, 10], bulk_frames=[5], api_runs=[(0, 11, 3)])
    # ---- small synthetic systems -----------------------------------------
    cases = {
        "synth_ortho": dict(n_waters=420, n_residues=8, n_chains=2, n_ions=2, triclinic=False, seed=11),
        "synth_tric": dict(n_waters=520, n_residues=10, n_chains=1, n_ions=1, triclinic=True,
                           nucleic_fraction=0.4, seed=12),
        "synth_sparse": dict(n_waters=90, n_residues=3, n_chains=1, n_ions=1, triclinic=False,
                             density=0.0045, seed=13),
    }
    for tag, kw in cases.items():
        s = synthetic.make_system(**kw)
        P, F, D = s.frames(0, 3)
        save_inputs(tag, s.names, s.types, s.resnames, s.resids, s.masses, s.charges, s.bonds, P, F, D)
        u = MDAnalysis.Universe.from_arrays(s.names, s.types, s.charges, s.masses, s.resids,
                                            s.resnames, s.bonds, P, F, D)
        try:
            record(u, tag, step_frames=range(3), centre_frames=[0, 1], bulk_frames=[0],
                   api_runs=[(0, 3, 1)])
        except Exception as exc:  # sparse systems may legitimately raise
            print(tag, "reference raised", type(exc).__name__, exc)
            write_json(tag, {"raises": type(exc).__name__})


if __name__ == "__main__":
    main()
