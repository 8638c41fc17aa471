"""Interfacial-water recipes on the B200 path.

Drop-in for `waterEntropy/recipes/interfacial_solvent.py` (same function names,
positional order and return objects; reference lines cited per function).  The
per-frame work -- neighbour search, RAD shells, HB labelling, shell labels,
force/torque covariances, count tables -- runs in the CUDA kernels behind
`WaterEntropyEngine`; this module only stages frames and rebuilds the
reference's Python objects from the device tables.
"""
from __future__ import annotations

from waterentropy_b200 import distributed as wdist
from waterentropy_b200.engine import WaterEntropyEngine
from waterentropy_b200.entropy.orientations import Orientations
from waterentropy_b200.entropy.vibrations import Vibrations
from waterentropy_b200.frames import iter_batches
from waterentropy_b200.system import SystemTopology
from waterentropy_b200.utils.helpers import nested_dict


def _device():
    try:
        import torch

        if torch.cuda.is_available():
            return torch.cuda.current_device()
    except Exception:  # pragma: no cover
        pass
    return 0


# ---- context cache -----------------------------------------------------------------------------
# Creating a context costs ~0.2 s on a C4-sized system (device workspace, pinned record buffers, the first
# short batch that sizes the label table) - more than the kernels of a hundred frames.  The reference's unit
# of work is `_entropy_per_step`, ONE frame per call (dask maps it over the frames), and users call the recipes
# block after block; so the last context is kept and reused (after `we_reset`) by the next call with the same
# topology object, recipe, device and options.  One context at most is kept; WE_ENGINE_CACHE=0 disables it.
_CACHE = {"key": None, "eng": None, "created": 0, "reused": 0}
_ENV_KEYS = ("WE_EIG_FLAVOUR", "WE_COPY_GATE", "WE_COV_STREAM", "WE_GRAPHS", "WE_GATE", "WE_LIB_PATH")


def _cache_key(top, recipe, kw):
    import os

    return (id(top), recipe, _device(), tuple(sorted(kw.items())), tuple(os.environ.get(k) for k in _ENV_KEYS))


def _acquire(system, recipe, **kw):
    import os

    top = SystemTopology.from_universe(system)
    key = _cache_key(top, recipe, kw)
    eng = None
    if _CACHE["eng"] is not None:
        eng, _CACHE["eng"] = _CACHE["eng"], None
        if _CACHE["key"] == key and eng.top is top and os.environ.get("WE_ENGINE_CACHE", "1") != "0":
            try:
                eng.reset()
                _CACHE["reused"] += 1
                eng._cache_key = key
                return eng
            except Exception:  # a context left in an error state is not reused
                pass
        eng.close()
    eng = WaterEntropyEngine(top, recipe, device=_device(), **kw)
    _CACHE["created"] += 1
    eng._cache_key = key
    return eng


def _release(eng):
    """Hand a context back after a successful run (instead of `close`)."""
    import os

    if os.environ.get("WE_ENGINE_CACHE", "1") == "0" or getattr(eng, "_cache_key", None) is None:
        eng.close()
        return
    if _CACHE["eng"] is not None:
        _CACHE["eng"].close()
    _CACHE["key"], _CACHE["eng"] = eng._cache_key, eng


def drop_cached_context():
    """Free the cached context (device workspace, pinned buffers) now."""
    if _CACHE["eng"] is not None:
        _CACHE["eng"].close()
    _CACHE["key"], _CACHE["eng"] = None, None


import atexit  # noqa: E402

atexit.register(drop_cached_context)


def _run(system, indices, recipe, keep_debug=False, keep_records=True, keep_shells=False, shells_only=False):
    if recipe == "interfacial" and len(SystemTopology.from_universe(system).solute_ua) == 0:
        # upstream ends in MDAnalysis' SelectionError here: `select_atoms("resid ")` with an empty
        # residue list (utils/selections.py:19-20); use the bulk recipe for solute-free systems
        raise ValueError("no solute molecules in the system: the interfacial recipe needs a solute "
                         "(use get_bulk_water_orient_entropy for pure water)")
    if not shells_only and getattr(system, "has_forces", True) is False:
        raise ValueError("the trajectory holds no forces; the covariance path needs them")
    if not shells_only and getattr(system, "has_charges", True) is False:
        raise ValueError("the topology holds no charges (GRO template): hydrogen bonds cannot be assigned; "
                         "only the shells path (get_interfacial_shells / waterShells) runs on it")
    eng = _acquire(system, recipe, keep_records=keep_records, keep_debug=keep_debug, keep_shells=keep_shells,
                   shells_only=shells_only)
    # whole kernel batches from a recycled ring of pinned buffers: reading batch k + 1 overlaps the
    # kernels of batch k, host memory stays O(ring)
    try:
        for P, F, D, ids in iter_batches(system, indices, eng.batch_frames, engine=eng, forces=not shells_only):
            eng.submit(P, F, D, ids)
        eng.sync()
    except Exception:
        eng.close()
        raise
    return eng


def get_interfacial_water_orient_entropy(system, start: int, end: int, step: int, temperature=298,
                                         parallel=False, client=None):
    """Reference: `recipes/interfacial_solvent.py:181-270`.

    Returns `(Sorient_dict, covariances, vibrations, frame_solvent_indices, n_frames)`.
    `parallel=True` shards the frames over the ranks of an initialised
    `torch.distributed` job (one process per GPU) and combines the per-GPU
    tables with NCCL; the dask `client` argument of the reference is accepted
    and ignored."""
    indices = list(range(start, end, step))
    world = wdist.world() if parallel else (0, 1)
    mine = wdist.shard_frames(indices, *world)
    eng = _run(system, mine, "interfacial")
    Sorients = Orientations()
    if world[1] > 1:
        covariances, entries, frame_solvent_indices = wdist.combine_raw(eng, download_entries=False)
        if entries is None:
            Sorients.add_device(eng)  # NCCL: the tables were merged on the device and are reduced there
        else:
            # host combine (gloo): from the merged label entries with array operations, the same numbers,
            # bit for bit, as Orientations().add_data(HBLabelCollection)
            Sorients.add_entries(eng.top, False, entries)
    else:
        # one GPU: the table is reduced on the device (we_orient_entropy), the label entries stay there
        covariances, frame_solvent_indices = eng.covariances(), eng.frame_solvent_indices()
        Sorients.add_device(eng)
    _release(eng)
    n_frames = len(indices)
    vibrations = Vibrations(temperature)
    vibrations.add_data(covariances, diagonalise=True)
    return Sorients.resid_labelled_Sorient, covariances, vibrations, frame_solvent_indices, n_frames


def _entropy_per_step(args):
    """Reference: `recipes/interfacial_solvent.py:273-345` -- one frame, fresh accumulators.
    Returns `(frame_solvent_indices, CovarianceCollection, HBLabelCollection, 1)`."""
    index, system = args
    eng = _run(system, [index], "interfacial")
    out = (eng.frame_solvent_indices(), eng.covariances(), eng.hb_labels(), 1)
    _release(eng)
    return out


def get_interfacial_shells(system, start: int, end: int, step: int):
    """Reference: `recipes/interfacial_solvent.py:104-146` -- RAD shells of the interfacial
    waters that have a non-like neighbour: `{frame: {water_idx: [shell idx ...]}}`."""
    indices = list(range(start, end, step))
    # upstream stops at the interfacial waters' own shells here (no HBs, labels or neighbours' shells,
    # no forces), so failures that only those later steps can raise must not surface: shells_only
    eng = _run(system, indices, "interfacial", keep_shells=True, shells_only=True)
    out = nested_dict()
    for k, frame in enumerate(eng.frame_ids):
        for atom, shell in eng.frame_shells(k).items():
            out[frame][atom] = shell
    _release(eng)
    return out


def find_interfacial_solvent(solutes, system, shells=None):
    """Reference: `recipes/interfacial_solvent.py:25-66` -- waters in the RAD shell of any
    solute united atom that has a water among its 20 nearest neighbours, for the
    CURRENT frame of `system`.  `solutes`: atom group (`.indices`) or index list
    restricting the solute atoms; whole molecules are used, as upstream."""
    top = SystemTopology.from_universe(system)
    frame = int(system.trajectory.ts.frame) if hasattr(system, "trajectory") else 0
    eng = _run(system, [frame], "interfacial", keep_debug=True, keep_records=False)
    sel = getattr(solutes, "indices", solutes)
    frag = set(int(top.fragment_id[a]) for a in sel)
    flags = eng.debug(0, "flags")
    all_shells = eng.shells(0)
    slot = {int(a): s for s, a in enumerate(top.heavy_idx)}
    found = set()
    for a in top.heavy_idx:
        a = int(a)
        if int(top.fragment_id[a]) in frag and (flags[slot[a]] & 1):
            found.update(n for n in all_shells[a] if top.is_water[n])
    _release(eng)
    return list(found)


def save_solvent_indices(frame, atom_idx, nearest_resid, nearest_resname, frame_solvent_indices):
    """Reference: `recipes/interfacial_solvent.py:69-88`."""
    if nearest_resid not in frame_solvent_indices[frame][nearest_resname]:
        frame_solvent_indices[frame][nearest_resname][nearest_resid] = []
    frame_solvent_indices[frame][nearest_resname][nearest_resid].append(atom_idx)
    return frame_solvent_indices


def print_frame_solvent_dicts(frame_solvent_indices: dict):
    """Reference: `recipes/interfacial_solvent.py:91-101`."""
    for frame, by_name in sorted(list(frame_solvent_indices.items())):
        for resname, by_resid in sorted(list(by_name.items())):
            for resid, solvents in sorted(list(by_resid.items())):
                print(frame, resname, resid, len(solvents), solvents)


def print_frame_solvent_shells(frame_solvent_shells: dict):
    """Reference: `recipes/interfacial_solvent.py:169-178`."""
    for frame, by_atom in sorted(list(frame_solvent_shells.items())):
        for atom_idx, shell in sorted(list(by_atom.items())):
            print(frame, atom_idx, shell, len(shell))
