"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the waterEntropy per-frame hot path.

NumPy/Python layer over `oracle/we_oracle.c`.  Together they restate the
reference's algorithm for the path `BASELINE.json:north_star` names; they are
the checker for the CUDA implementation and the CPU baseline of `bench.py`.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU arms may import
this module; the product package never does.

Parity status: PINNED.  `tests/test_oracle_golden.py` checks every function
here against (a) the literal golden values in the reference's own tests and
(b) vectors produced by running the unmodified reference through
`oracle/mda_shim` (`oracle/make_golden.py`, committed with its outputs in
`tests/golden/`).  Triclinic boxes are the exception: no reference test covers
them and MDAnalysis itself is absent, so the triclinic minimum image is
"parity unpinned" (restated from MDAnalysis' documented algorithm).

Reference lines restated (under /root/reference/waterEntropy/):
  utils/selections.py:34-56,89-105        -> OracleTopology (solutes, scales)
  recipes/interfacial_solvent.py:25-66    -> interfacial_frame (gate + collect)
  recipes/interfacial_solvent.py:273-345  -> interfacial_frame
  recipes/bulk_water.py:37-72             -> bulk_frame
  analysis/shell_labels.py:10-105         -> _shell_labels / _nearest_nonlike
  analysis/HB.py:142-169                  -> _hb_for (which centres get HBs)
  analysis/HB_labels.py:37-156,299-344    -> LabelCounts / _hb_labels
  recipes/forces_torques.py:28-55         -> forces_torques
  maths/transformations.py:12-79,104-135  -> forces_torques
  entropy/convariances.py:21-108          -> CovMeans
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from collections import defaultdict

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_SO = os.path.join(_BUILD, "libwe_oracle.so")
_SRC = os.path.join(_HERE, "we_oracle.c")

MAX_NBR = 25
WATER_RESNAMES = frozenset(
    "H2O HOH OH2 HHO OHH TIP T3P T4P T5P SOL WAT TIP2 TIP3 TIP4".split()
)


def build(force: bool = False) -> str:
    """Compile the C restatement with gcc (libm powf/pow, no FMA contraction)."""
    if (
        force
        or not os.path.exists(_SO)
        or os.path.getmtime(_SO) < os.path.getmtime(_SRC)
    ):
        os.makedirs(_BUILD, exist_ok=True)
        subprocess.check_call(
            [
                "gcc", "-O2", "-fno-builtin", "-ffp-contract=off", "-fopenmp", "-fPIC",
                "-shared", "-o", _SO, _SRC, "-lm",
            ]
        )
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        P = ctypes.c_void_p
        I = ctypes.c_int
        L.we_oracle_rad.restype = I
        L.we_oracle_rad.argtypes = [I, P, P, I, P, P, P, P, I, P, P, P, P, P, P, P, P]
        L.we_oracle_hb.restype = I
        L.we_oracle_hb.argtypes = [I, P, P, P, I, P, P, P, P, P, P]
        L.we_oracle_angle_cos.restype = ctypes.c_float
        L.we_oracle_angle_cos.argtypes = [P, P, P, P]
        L.we_oracle_distance.restype = ctypes.c_double
        L.we_oracle_distance.argtypes = [P, P, P]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def nested_dict():
    return defaultdict(nested_dict)


def to_plain(d):
    """Recursively convert nested defaultdicts to plain dicts (for comparisons)."""
    if isinstance(d, dict):
        return {k: to_plain(v) for k, v in d.items()}
    return d


# --------------------------------------------------------------------------
# topology (frame-invariant)
# --------------------------------------------------------------------------
class OracleTopology:
    """Frame-invariant tables.  Mirrors what the reference recomputes every
    frame through MDAnalysis selection strings."""

    def __init__(self, masses, charges, resids, resnames, types, bonds, names=None):
        self.n_atoms = n = len(masses)
        self.masses = np.asarray(masses, dtype=np.float64)
        self.charges = np.asarray(charges, dtype=np.float64)
        self.resids = np.asarray(resids, dtype=np.int64)
        self.resnames = [str(r) for r in resnames]
        self.types = [str(t) for t in types]
        self.names = None if names is None else [str(x) for x in names]
        bonds = np.asarray(bonds, dtype=np.int64).reshape(-1, 2)
        lo = np.minimum(bonds[:, 0], bonds[:, 1])
        hi = np.maximum(bonds[:, 0], bonds[:, 1])
        self.bonds = np.unique(np.stack([lo, hi], axis=1), axis=0)
        self.is_water = np.array(
            [r in WATER_RESNAMES for r in self.resnames], dtype=np.uint8
        )
        adj = [[] for _ in range(n)]
        for a, b in self.bonds:
            adj[a].append(int(b))
            adj[b].append(int(a))
        self.adj = adj
        # fragments = connected components
        frag = -np.ones(n, dtype=np.int64)
        members = []
        for s in range(n):
            if frag[s] >= 0:
                continue
            fid = len(members)
            stack, comp = [s], []
            frag[s] = fid
            while stack:
                v = stack.pop()
                comp.append(v)
                for w in adj[v]:
                    if frag[w] < 0:
                        frag[w] = fid
                        stack.append(w)
            members.append(sorted(comp))
        self.fragment_id = frag
        self.fragment_members = members
        # heavy atoms: "mass 2 to 999"
        heavy = (self.masses >= 2.0) & (self.masses <= 999.0)
        self.is_heavy = heavy
        self.heavy_idx = np.nonzero(heavy)[0].astype(np.int32)
        self.heavy_slot = -np.ones(n, dtype=np.int64)
        self.heavy_slot[self.heavy_idx] = np.arange(len(self.heavy_idx))
        ptr = [0]
        ex = []
        for i in self.heavy_idx:
            ex.extend(adj[i])
            ptr.append(len(ex))
        self.excl_ptr = np.array(ptr, dtype=np.int32)
        self.excl = np.array(ex if ex else [0], dtype=np.int32)
        # HB donors of heavy atom X: bonds (X, D) with D the second (higher)
        # index, 1.0 < mass_D < 1.1 and q_D > 0   (analysis/HB.py:111-115)
        self.donors = {}
        for i in self.heavy_idx:
            ds = [
                j
                for j in sorted(adj[i])
                if j > i and 1.0 < self.masses[j] < 1.1 and self.charges[j] > 0
            ]
            self.donors[int(i)] = ds
        self.solute_resids = self._find_solute_molecules()
        self.solute_UAs = self._solute_UAs()
        self.water_centres = np.array(
            [i for i in self.heavy_idx if self.is_water[i]], dtype=np.int64
        )

    # utils/selections.py:34-56
    def _find_solute_molecules(self):
        nonwater = np.nonzero(self.is_water == 0)[0]
        out = []
        seen_frag = []
        for i in nonwater:
            f = self.fragment_id[i]
            if f not in seen_frag:
                seen_frag.append(f)
        seen_frag.sort(key=lambda f: self.fragment_members[f][0])
        nonwater_mask = self.is_water == 0
        for f in seen_frag:
            mem = self.fragment_members[f]
            for rid in sorted(set(int(self.resids[a]) for a in mem)):
                sel = np.nonzero(nonwater_mask & (self.resids == rid) & self.is_heavy)[0]
                if len(sel) > 1:
                    out.append(rid)
                if len(sel) == 1 and np.floor(self.masses[sel[0]]) != 16:
                    out.append(rid)
        return out

    # recipes/interfacial_solvent.py:39-43, :298-299
    def _solute_UAs(self):
        if not self.solute_resids:
            return np.zeros(0, dtype=np.int64)
        sel = np.nonzero(np.isin(self.resids, np.array(self.solute_resids)))[0]
        frags = []
        for a in sel:
            f = self.fragment_id[a]
            if f not in frags:
                frags.append(f)
        frags.sort(key=lambda f: self.fragment_members[f][0])
        uas = []
        for f in frags:
            uas.extend(a for a in self.fragment_members[f] if self.is_heavy[a])
        return np.array(uas, dtype=np.int64)

    def bonded_heavy(self, atom):
        """utils/selections.py:59-73"""
        if self.masses[atom] < 1.1:
            hv = [j for j in sorted(self.adj[atom]) if self.is_heavy[j]]
            return hv[0]
        return atom


# --------------------------------------------------------------------------
# accumulators (plain restatements of the reference's containers)
# --------------------------------------------------------------------------
class CovMeans:
    """entropy/convariances.py:10-108 -- running means keyed (nearest, resname)."""

    def __init__(self):
        self.forces = {}
        self.torques = {}
        self.counts = {}

    def add(self, key, F, T):
        if key not in self.forces:
            self.forces[key] = F
            self.torques[key] = T
            self.counts[key] = 1
        else:
            n = self.counts[key]
            self.forces[key] = (self.forces[key] * n + F) / (n + 1)
            self.torques[key] = (self.torques[key] * n + T) / (n + 1)
            self.counts[key] = n + 1

    def merge(self, other):
        for key in other.forces:
            if key not in self.forces:
                self.forces[key] = other.forces[key]
                self.torques[key] = other.torques[key]
                self.counts[key] = other.counts[key]
            else:
                a, b = self.counts[key], other.counts[key]
                self.forces[key] = (self.forces[key] * a + other.forces[key] * b) / (a + b)
                self.torques[key] = (self.torques[key] * a + other.torques[key] * b) / (a + b)
                self.counts[key] = a + b


class LabelCounts:
    """analysis/HB_labels.py:15-296 -- the two nested count tables."""

    def __init__(self):
        self.labelled_shell_counts = nested_dict()
        self.resid_labelled_shell_counts = nested_dict()

    def add(self, resid, resname, labels, donates_to, accepts_from, N_w):
        key = tuple(sorted(labels))
        for entry in (
            self.labelled_shell_counts[resname][key],
            self.resid_labelled_shell_counts[resid][resname][key],
        ):
            if "shell_count" not in entry:
                entry["shell_count"] = 1
                entry["N_w"] = N_w
            else:
                entry["shell_count"] += 1
            for a in donates_to:
                entry["donates_to"][a] = entry["donates_to"].get(a, 0) + 1
            for d in accepts_from:
                entry["accepts_from"][d] = entry["accepts_from"].get(d, 0) + 1

    @staticmethod
    def _merge_entry(dst, src):
        for k, v in src.items():
            if k in ("donates_to", "accepts_from"):
                for lab, c in v.items():
                    dst[k][lab] = dst[k].get(lab, 0) + c
            elif k in dst and k != "N_w":
                dst[k] += v
            else:
                dst[k] = v

    def merge(self, other):
        for resid, a in other.resid_labelled_shell_counts.items():
            for resname, b in a.items():
                for key, src in b.items():
                    self._merge_entry(
                        self.resid_labelled_shell_counts[resid][resname][key], src
                    )
        for resname, b in other.labelled_shell_counts.items():
            for key, src in b.items():
                self._merge_entry(self.labelled_shell_counts[resname][key], src)


# --------------------------------------------------------------------------
# per-frame kernels
# --------------------------------------------------------------------------
def rad_frame(top: OracleTopology, pos, dims, centres=None):
    """Neighbour lists + RAD shells.  `centres`: heavy-atom *atom indices*
    (None = every heavy atom).  Returns dict of arrays indexed by ordinal."""
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    dims = np.ascontiguousarray(dims, dtype=np.float32)
    if centres is None:
        slots = None
        nc = len(top.heavy_idx)
        atoms = top.heavy_idx.astype(np.int64)
    else:
        atoms = np.asarray(centres, dtype=np.int64)
        slots = np.ascontiguousarray(top.heavy_slot[atoms], dtype=np.int32)
        if (slots < 0).any():
            raise ValueError("centre is not a heavy atom")
        nc = len(atoms)
    out = dict(
        atoms=atoms,
        nbr_total=np.zeros(nc, dtype=np.int32),
        nbr_idx=-np.ones((nc, MAX_NBR), dtype=np.int32),
        nbr_dist=np.zeros((nc, MAX_NBR), dtype=np.float64),
        shell_n=np.zeros(nc, dtype=np.int32),
        shell_idx=-np.ones((nc, MAX_NBR), dtype=np.int32),
        gate=np.zeros(nc, dtype=np.uint8),
        status=np.zeros(nc, dtype=np.int32),
    )
    rc = lib().we_oracle_rad(
        top.n_atoms, _p(pos), _p(dims), len(top.heavy_idx), _p(top.heavy_idx),
        _p(top.excl_ptr), _p(top.excl), _p(top.is_water), nc,
        None if slots is None else _p(slots),
        _p(out["nbr_total"]), _p(out["nbr_idx"]), _p(out["nbr_dist"]),
        _p(out["shell_n"]), _p(out["shell_idx"]), _p(out["gate"]), _p(out["status"]),
    )
    if rc != 0:
        raise MemoryError("we_oracle_rad failed")
    return out


def hb_frame(top: OracleTopology, pos, dims, don_x, don_h, shl_n, shl_idx):
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    dims = np.ascontiguousarray(dims, dtype=np.float32)
    don_x = np.ascontiguousarray(don_x, dtype=np.int32)
    don_h = np.ascontiguousarray(don_h, dtype=np.int32)
    shl_n = np.ascontiguousarray(shl_n, dtype=np.int32)
    shl_idx = np.ascontiguousarray(shl_idx, dtype=np.int32)
    nd = len(don_x)
    acc = -np.ones(nd, dtype=np.int32)
    st = np.zeros(nd, dtype=np.int32)
    if nd:
        lib().we_oracle_hb(
            top.n_atoms, _p(pos), _p(dims), _p(top.charges), nd, _p(don_x), _p(don_h),
            _p(shl_n), _p(shl_idx), _p(acc), _p(st),
        )
    return acc, st


class FrameState:
    """All-centre RAD + HB state of one frame (device analogue: the per-frame
    shell / acceptor arrays).  Computing every heavy centre up front is
    equivalent to the reference's lazy memoisation because each shell / HB is a
    pure function of the frame."""

    def __init__(self, top: OracleTopology, pos, dims):
        self.top = top
        self.pos = np.ascontiguousarray(pos, dtype=np.float32)
        self.dims = np.ascontiguousarray(dims, dtype=np.float32)
        self.rad = rad_frame(top, self.pos, self.dims)
        don_x, don_h = [], []
        for i in top.heavy_idx:
            for h in top.donors[int(i)]:
                don_x.append(int(i))
                don_h.append(h)
        self.don_x = np.array(don_x, dtype=np.int32)
        self.don_h = np.array(don_h, dtype=np.int32)
        slots = top.heavy_slot[self.don_x] if len(don_x) else np.zeros(0, dtype=np.int64)
        self.acc, self.hb_status = hb_frame(
            top, self.pos, self.dims, self.don_x, self.don_h,
            self.rad["shell_n"][slots], self.rad["shell_idx"][slots],
        )
        self.don_of = defaultdict(list)  # X -> [(D, A, status)]
        for x, h, a, s in zip(self.don_x, self.don_h, self.acc, self.hb_status):
            self.don_of[int(x)].append((int(h), int(a), int(s)))

    def shell(self, atom):
        s = self.top.heavy_slot[atom]
        st = int(self.rad["status"][s])
        if st == 1:
            raise ValueError(f"Atom coordinates of {atom} in neighbour list")
        if st == 2:
            raise ValueError("not enough values to unpack (no neighbours in cut-off)")
        n = int(self.rad["shell_n"][s])
        return [int(v) for v in self.rad["shell_idx"][s, :n]]

    def neighbours(self, atom):
        s = self.top.heavy_slot[atom]
        st = int(self.rad["status"][s])
        if st == 1:
            raise ValueError(f"Atom coordinates of {atom} in neighbour list")
        if st == 2:
            raise ValueError("not enough values to unpack (no neighbours in cut-off)")
        return s

    def donations(self, atom):
        """[(D, A)] for heavy atom `atom` (analysis/HB.py:96-139)."""
        out = []
        for d, a, st in self.don_of.get(int(atom), []):
            if st == 1:
                raise ValueError("Atom coordinates in neighbour list (HB donor)")
            if st == 2:
                raise ValueError("not enough values to unpack (HB donor)")
            out.append((d, a))
        return out


class LazyFrameState:
    """Memoised per-centre RAD + HB state: shells and acceptors are computed on demand and cached, which is
    exactly what the reference does per frame (`ShellCollection.find_shell`, analysis/RAD.py:86;
    `HBCollection`, analysis/HB.py:142-169).  Same interface as `FrameState`; `n_evaluated` counts the
    distinct centres whose neighbour search + RAD test ran - the reference's real per-frame work."""

    def __init__(self, top: OracleTopology, pos, dims):
        self.top = top
        self.pos = np.ascontiguousarray(pos, dtype=np.float32)
        self.dims = np.ascontiguousarray(dims, dtype=np.float32)
        H = len(top.heavy_idx)
        self.known = np.zeros(H, dtype=bool)
        self.rad = dict(
            nbr_total=np.zeros(H, dtype=np.int32), nbr_idx=-np.ones((H, MAX_NBR), dtype=np.int32),
            nbr_dist=np.zeros((H, MAX_NBR), dtype=np.float64), shell_n=np.zeros(H, dtype=np.int32),
            shell_idx=-np.ones((H, MAX_NBR), dtype=np.int32), gate=np.zeros(H, dtype=np.uint8),
            status=np.zeros(H, dtype=np.int32))
        self.don_of = {}
        self.n_evaluated = 0
        self.n_hb_centres = 0

    def ensure(self, atoms):
        """RAD state of the heavy atoms `atoms` (one batched call for those not yet known)."""
        slots = np.unique(self.top.heavy_slot[np.asarray(list(atoms), dtype=np.int64)])
        slots = slots[~self.known[slots]]
        if len(slots) == 0:
            return
        r = rad_frame(self.top, self.pos, self.dims, centres=self.top.heavy_idx[slots])
        for k in ("nbr_total", "nbr_idx", "nbr_dist", "shell_n", "shell_idx", "gate", "status"):
            self.rad[k][slots] = r[k]
        self.known[slots] = True
        self.n_evaluated += len(slots)

    def ensure_hb(self, atoms):
        """Acceptors of the donor hydrogens of `atoms` (needs their shells)."""
        todo = [int(a) for a in atoms if int(a) not in self.don_of]
        if not todo:
            return
        self.ensure(todo)
        don_x, don_h = [], []
        for x in todo:
            self.don_of[x] = []
            for h in self.top.donors[x]:
                don_x.append(x)
                don_h.append(h)
        self.n_hb_centres += len(todo)
        if not don_x:
            return
        don_x = np.array(don_x, dtype=np.int32)
        don_h = np.array(don_h, dtype=np.int32)
        slots = self.top.heavy_slot[don_x]
        acc, st = hb_frame(self.top, self.pos, self.dims, don_x, don_h, self.rad["shell_n"][slots],
                           self.rad["shell_idx"][slots])
        # a donor whose centre has no valid shell inherits the centre's failure
        for x, h, a, s, rs in zip(don_x, don_h, acc, st, self.rad["status"][slots]):
            self.don_of[int(x)].append((int(h), int(a), int(rs) if rs else int(s)))

    def _slot(self, atom):
        self.ensure([atom])
        s = self.top.heavy_slot[atom]
        st = int(self.rad["status"][s])
        if st == 1:
            raise ValueError(f"Atom coordinates of {atom} in neighbour list")
        if st == 2:
            raise ValueError("not enough values to unpack (no neighbours in cut-off)")
        return s

    def shell(self, atom):
        s = self._slot(atom)
        return [int(v) for v in self.rad["shell_idx"][s, : int(self.rad["shell_n"][s])]]

    def neighbours(self, atom):
        return self._slot(atom)

    def donations(self, atom):
        self.ensure_hb([atom])
        out = []
        for d, a, st in self.don_of.get(int(atom), []):
            if st == 1:
                raise ValueError("Atom coordinates in neighbour list (HB donor)")
            if st == 2:
                raise ValueError("not enough values to unpack (HB donor)")
            out.append((d, a))
        return out

    def prefetch_interfacial(self):
        """Batch the closure of the interfacial recipe (solute united atoms -> interfacial waters -> their
        shell members -> donors of both) so the C routine sees large centre lists; the result is what the
        one-at-a-time memoisation would hold."""
        top = self.top
        self.ensure(top.solute_UAs)
        slots = top.heavy_slot[np.asarray(top.solute_UAs, dtype=np.int64)]
        ok = (self.rad["status"][slots] == 0) & (self.rad["gate"][slots] == 1)
        members = self.rad["shell_idx"][slots[ok]]
        valid = np.arange(MAX_NBR)[None, :] < self.rad["shell_n"][slots[ok]][:, None]
        cand = np.unique(members[valid])
        inter = cand[top.is_water[cand] == 1]
        self.ensure(inter)
        s2 = top.heavy_slot[inter]
        good = self.rad["status"][s2] == 0
        m2 = self.rad["shell_idx"][s2[good]]
        v2 = np.arange(MAX_NBR)[None, :] < self.rad["shell_n"][s2[good]][:, None]
        mem = np.unique(m2[v2]) if len(m2) else np.zeros(0, dtype=np.int64)
        self.ensure(mem)
        self.ensure_hb(np.unique(np.concatenate([inter, mem])) if len(mem) else inter)
        return inter


def _nearest_nonlike(top, centre, shell):
    """analysis/shell_labels.py:90-105"""
    for n in shell:
        if top.resnames[n] != top.resnames[centre] and top.types[n] != top.types[centre]:
            return n
    return None


def _shell_labels(top, fs: FrameState, centre, shell):
    """analysis/shell_labels.py:10-87 -> (nearest_nonlike, labels, N_w) or None"""
    nn = _nearest_nonlike(top, centre, shell)
    if nn is None:
        return None
    labels = []
    N_w = 0
    for n in shell:
        if n == nn:
            labels.append(f"0_{top.resnames[n]}")
        if n != nn and top.resnames[n] != top.resnames[centre]:
            labels.append(top.resnames[n])
        if n != nn and top.resnames[n] == top.resnames[centre]:
            N_w += 1
            nshell = fs.shell(n)
            nnn = _nearest_nonlike(top, n, nshell)
            if nnn is None:
                labels.append(f"2_{top.resnames[n]}")
            elif top.resids[nnn] == top.resids[nn]:
                labels.append(f"1_{top.resnames[n]}")
            else:
                labels.append(f"X_{top.resnames[n]}")
    return nn, labels, N_w


def _hb_labels(top, fs: FrameState, centre, shell, labels):
    """analysis/HB.py:142-169 + analysis/HB_labels.py:299-344.
    Returns (donates_to_labels, accepts_from_labels)."""
    donates, accepts = [], []
    if not labels:
        return donates, accepts
    # accepting_from[centre] restricted to donors whose heavy atom is in the
    # centre's shell: exactly the donors of shell members that point at centre
    for n in shell:
        for d, a in fs.donations(n):
            if a == centre:
                accepts.append(labels[shell.index(n)])
    for d, a in fs.donations(centre):
        donates.append(labels[shell.index(a)])
    return donates, accepts


def forces_torques(top: OracleTopology, pos, frc, waters):
    """recipes/forces_torques.py:28-55 for a list of single-UA molecules
    (given by their heavy atom).  Returns dict with F, T (vectors), Fcov, Tcov,
    axes, moi -- batched NumPy restatement (np.linalg.eig == LAPACK dgeev)."""
    waters = list(waters)
    M = len(waters)
    if M == 0:
        z = np.zeros((0, 3, 3))
        return dict(F=np.zeros((0, 3)), T=np.zeros((0, 3)), Fcov=z, Tcov=z, axes=z, moi=np.zeros((0, 3)))
    Fv = np.empty((M, 3))
    Tv = np.empty((M, 3))
    AX = np.empty((M, 3, 3))
    MOI = np.empty((M, 3))
    # group by fragment size for batching
    by_size = defaultdict(list)
    for k, w in enumerate(waters):
        mem = top.fragment_members[top.fragment_id[w]]
        by_size[len(mem)].append((k, mem))
    for size, items in by_size.items():
        ks = np.array([k for k, _ in items])
        idx = np.array([mem for _, mem in items], dtype=np.int64)  # [m, size]
        p32 = pos[idx]  # [m,size,3] float32
        f32 = frc[idx]
        mass = top.masses[idx]  # [m,size]
        coords = p32.astype(np.float64)
        com = (coords * mass[:, :, None]).sum(axis=1) / mass.sum(axis=1)[:, None]
        p = p32 - com[:, None, :]  # float32 - float64 -> float64
        tens = np.zeros((len(ks), 3, 3))
        x, y, z = p[..., 0], p[..., 1], p[..., 2]
        tens[:, 0, 0] = (mass * (y**2 + z**2)).sum(axis=1)
        tens[:, 0, 1] = tens[:, 1, 0] = -(mass * x * y).sum(axis=1)
        tens[:, 0, 2] = tens[:, 2, 0] = -(mass * x * z).sum(axis=1)
        tens[:, 1, 1] = (mass * (x**2 + z**2)).sum(axis=1)
        tens[:, 1, 2] = tens[:, 2, 1] = -(mass * y * z).sum(axis=1)
        tens[:, 2, 2] = (mass * (x**2 + y**2)).sum(axis=1)
        e_val, e_vec = np.linalg.eig(tens)
        e_val = e_val.real
        e_vec = e_vec.real
        order = np.argsort(e_val, axis=1)[:, ::-1]
        axes = np.take_along_axis(e_vec, order[:, None, :], axis=2).transpose(0, 2, 1)
        hand = np.einsum("mi,mi->m", np.cross(axes[:, 0], axes[:, 1]), axes[:, 2])
        axes = np.where((hand < 0)[:, None, None], -axes, axes)
        order2 = np.argsort(np.abs(e_val), axis=1)[:, ::-1]
        moi_axis = np.take_along_axis(e_val, order2, axis=1)
        # torque (maths/transformations.py:12-37)
        rc = np.einsum("mak,mik->mai", p, axes)
        rf = np.einsum("mak,mik->mai", f32.astype(np.float64), axes)
        cross = np.cross(rc, rf)
        Tq = (cross / np.sqrt(moi_axis)[:, None, :]).sum(axis=1)
        # mass weighted force (maths/transformations.py:40-64)
        fs32 = f32[:, 0, :]
        for a in range(1, size):
            fs32 = fs32 + f32[:, a, :]  # float32 sequential sum
        rot = np.einsum("mk,mik->mi", fs32.astype(np.float64), axes)
        Fm = rot / (mass.sum(axis=1) ** 0.5)[:, None]
        Fv[ks] = Fm
        Tv[ks] = Tq
        AX[ks] = axes
        MOI[ks] = moi_axis
    Fcov = Fv[:, :, None] * Fv[:, None, :] * 0.25
    Tcov = Tv[:, :, None] * Tv[:, None, :] * 0.25
    return dict(F=Fv, T=Tv, Fcov=Fcov, Tcov=Tcov, axes=AX, moi=MOI)


# --------------------------------------------------------------------------
# `multiple_UAs` molecules (recipes/forces_torques.py:14-93 with maths/transformations.py:82-353):
# whole-molecule matrices in the principal-axis frame + one pair of matrices per united atom in its
# bonded-axes frame.  Plain loops; float32 / float64 placement follows NumPy's promotion of the
# reference's expressions (positions, forces, box are float32; masses, centres of mass float64).
# --------------------------------------------------------------------------
def _get_vector(a, b, dims3):
    """maths/trig.py:139-157.  Note the first test: EVERY negative component gets +box."""
    delta = b - a
    out = []
    for delt, box in zip(delta, dims3):
        if delt < 0 and delt < 0.5 * box:
            delt = delt + box
        if delt > 0 and delt > 0.5 * box:
            delt = delt - box
        out.append(delt)
    return np.array(out)


def _principal_axes_moi(pos32, masses):
    """shim AtomGroup.principal_axes / moment_of_inertia + maths/transformations.py:113-133"""
    com = (pos32.astype(np.float64) * masses[:, None]).sum(axis=0) / masses.sum()
    p = pos32 - com
    t = np.zeros((3, 3))
    t[0][0] = (masses * (p[:, 1] ** 2 + p[:, 2] ** 2)).sum()
    t[0][1] = t[1][0] = -(masses * p[:, 0] * p[:, 1]).sum()
    t[0][2] = t[2][0] = -(masses * p[:, 0] * p[:, 2]).sum()
    t[1][1] = (masses * (p[:, 0] ** 2 + p[:, 2] ** 2)).sum()
    t[1][2] = t[2][1] = -(masses * p[:, 1] * p[:, 2]).sum()
    t[2][2] = (masses * (p[:, 0] ** 2 + p[:, 1] ** 2)).sum()
    e_val, e_vec = np.linalg.eig(t)
    order = np.argsort(e_val)[::-1]
    axes = e_vec[:, order].T
    if np.dot(np.cross(axes[0], axes[1]), axes[2]) < 0:
        axes = axes * -1
    return com, t, axes


def _moi_axis(tensor):
    ev = np.linalg.eig(tensor)[0]
    return ev[np.abs(ev).argsort()[::-1]]


def _custom_axes(a, b_list, c, dims3):
    """maths/transformations.py:166-228"""
    axis1 = np.zeros(3)
    for b in b_list:
        ab = _get_vector(a, b, dims3)
        axis1 += np.divide(ab, np.sqrt((ab**2).sum(axis=-1)))
    ac = _get_vector(b_list[0], c, dims3) if len(b_list) > 2 else _get_vector(a, c, dims3)
    ac_norm = np.divide(ac, np.sqrt((ac**2).sum(axis=-1)))
    axis2 = np.cross(ac_norm, axis1) if len(b_list) > 2 else np.cross(axis1, ac_norm)
    axis3 = np.cross(axis1, axis2)
    return np.array((axis1, axis2, axis3))


def _custom_PI_MOI(pos32, masses, custom_axes, centre32, dims3):
    """maths/transformations.py:231-283 (incl. the column-wise division by the ROW norms of :239-240)"""
    norm2 = np.sum(custom_axes**2, axis=1)
    PI = custom_axes / norm2**0.5
    RR = _get_vector(pos32[0], centre32, dims3)
    for i in range(3):
        if np.dot(PI[i], RR) < 0:
            PI[i] = -PI[i]
    translated = pos32 - centre32
    mi = np.zeros(3)
    for coord, mass in zip(translated, masses):
        mi += np.sum(np.cross(PI, coord) ** 2 * mass, axis=1)
    return PI, mi


def _torque_force(pos32, frc32, masses, centre, axes, mi_axis):
    """maths/transformations.py:12-64"""
    rc = np.tensordot(pos32 - centre, axes.T, axes=1)
    rf = np.tensordot(frc32, axes.T, axes=1)
    torque = np.sum(np.divide(np.cross(rc, rf), np.sqrt(mi_axis)), axis=0)
    fs = np.sum(frc32, axis=0)
    force = np.tensordot(fs, axes.T, axes=1) / np.sum(masses) ** 0.5
    return force, torque


def bonded_axes(top: OracleTopology, pos, dims3, atom):
    """maths/transformations.py:286-353 -> (axes, MI axis) or (None, position) as upstream."""
    heavy = [j for j in sorted(top.adj[atom]) if 2.0 <= top.masses[j] <= 999.0]
    light = [j for j in sorted(top.adj[atom]) if 1.0 <= top.masses[j] <= 1.1]
    ua_all = [atom] + heavy + light
    a = pos[atom]
    custom = None
    vec = a
    if len(heavy) == 2:
        custom = _custom_axes(a, [pos[heavy[0]]], pos[heavy[1]], dims3)
    if len(heavy) == 1 and len(light) >= 1:
        custom = _custom_axes(a, [pos[heavy[0]]], pos[light[0]], dims3)
    if len(heavy) > 2:
        custom = _custom_axes(a, pos[heavy], pos[heavy[1]], dims3)
    if len(heavy) == 1 and len(light) == 1:
        custom = _custom_axes(a, [pos[heavy[0]]], np.zeros(3), dims3)
    if len(heavy) == 0:
        _, tens, custom = _principal_axes_moi(pos[ua_all], top.masses[ua_all])
        vec = _moi_axis(tens)
    if custom is not None:
        custom, vec = _custom_PI_MOI(pos[ua_all], top.masses[ua_all], custom, a, dims3)
    return custom, vec


def forces_torques_multi(top: OracleTopology, pos, frc, dims, indices):
    """recipes/forces_torques.py:14-93 for ONE `multiple_UAs` molecule (atom indices ascending).
    Returns dict(whole_F, whole_T [3,3]; ua_F, ua_T [k,3,3]; uas = the united atoms that got a frame)."""
    idx = np.asarray(sorted(int(i) for i in indices), dtype=np.int64)
    dims3 = np.asarray(dims, dtype=np.float32)[:3]
    p32, f32, m = pos[idx], frc[idx], top.masses[idx]
    uas = [int(i) for i in idx if 2.0 <= top.masses[i] <= 999.0]
    inside = set(int(i) for i in idx)
    # maths/transformations.py:82-101: united-atom masses (hydrogens bonded INSIDE the molecule)
    ua_masses = []
    for i in idx:
        if top.masses[i] > 1.1:
            mm = top.masses[i]
            for h in sorted(top.adj[int(i)]):
                if h in inside and 1.0 <= top.masses[h] <= 1.1:
                    mm += top.masses[h]
            ua_masses.append(mm)
    com, _, axes = _principal_axes_moi(p32, m)
    tens = np.zeros((3, 3))
    for coord, mass in zip(pos[uas], ua_masses):  # maths/transformations.py:138-163
        dx, dy, dz = coord[0] - com[0], coord[1] - com[1], coord[2] - com[2]
        tens[0][0] += (abs(dy) ** 2 + abs(dz) ** 2) * mass
        tens[0][1] -= dx * dy * mass
        tens[1][0] -= dx * dy * mass
        tens[1][1] += (abs(dx) ** 2 + abs(dz) ** 2) * mass
        tens[0][2] -= dx * dz * mass
        tens[2][0] -= dx * dz * mass
        tens[2][2] += (abs(dx) ** 2 + abs(dy) ** 2) * mass
        tens[1][2] -= dy * dz * mass
        tens[2][1] -= dy * dz * mass
    F, T = _torque_force(p32, f32, m, com, axes, _moi_axis(tens))
    out = dict(whole_F=np.outer(F, F), whole_T=np.outer(T, T), uas=[], ua_F=[], ua_T=[])
    for ua in uas:
        custom, mi = bonded_axes(top, pos, dims3, ua)
        if custom is None:
            continue
        F, T = _torque_force(p32, f32, m, pos[ua], custom, mi)
        out["uas"].append(ua)
        out["ua_F"].append(np.outer(F, F))
        out["ua_T"].append(np.outer(T, T))
    if not out["uas"]:
        raise ValueError("need at least one array to concatenate")  # np.concatenate([]) upstream (:88-92)
    out["ua_F"] = np.array(out["ua_F"])
    out["ua_T"] = np.array(out["ua_T"])
    return out


def find_interfacial_solvent(top: OracleTopology, fs: FrameState, solute_UAs=None):
    """recipes/interfacial_solvent.py:25-66 -> sorted unique water atom indices"""
    uas = top.solute_UAs if solute_UAs is None else solute_UAs
    found = set()
    for a in uas:
        s = fs.neighbours(a)  # raises like get_sorted_neighbours would
        if fs.rad["gate"][s]:
            for n in fs.shell(a):
                if top.is_water[n]:
                    found.add(n)
    return sorted(found)


def interfacial_frame(top: OracleTopology, pos, frc, dims, frame_id, lazy=False):
    """recipes/interfacial_solvent.py:273-345 (`_entropy_per_step`).
    Returns (frame_solvent_indices, CovMeans, LabelCounts, 1, detail).  `lazy`: evaluate only the
    centres the reference's memoisation touches (needed at full size; identical results)."""
    if lazy:
        fs = LazyFrameState(top, pos, dims)
        fs.prefetch_interfacial()
    else:
        fs = FrameState(top, pos, dims)
    fsi = nested_dict()
    cov = CovMeans()
    lab = LabelCounts()
    detail = dict(recorded=[], nearest=[], labels=[], donates=[], accepts=[], N_w=[], shells={})
    for X in find_interfacial_solvent(top, fs):
        shell = fs.shell(X)
        # HB step touches X and every member (errors propagate)
        fs.donations(X)
        for n in shell:
            fs.shell(n)
            fs.donations(n)
        res = _shell_labels(top, fs, X, shell)
        if res is None:
            continue
        nn, labels, N_w = res
        don, acc = _hb_labels(top, fs, X, shell, labels)
        resid, resname = int(top.resids[nn]), top.resnames[nn]
        lab.add(resid, resname, labels, don, acc, N_w)
        if resid not in fsi[frame_id][resname]:
            fsi[frame_id][resname][resid] = []
        fsi[frame_id][resname][resid].append(int(X))
        detail["recorded"].append(int(X))
        detail["nearest"].append(int(nn))
        detail["labels"].append(labels)
        detail["donates"].append(don)
        detail["accepts"].append(acc)
        detail["N_w"].append(N_w)
        detail["shells"][int(X)] = shell
    ft = forces_torques(top, np.asarray(pos, dtype=np.float32), np.asarray(frc, dtype=np.float32), detail["recorded"])
    for k, X in enumerate(detail["recorded"]):
        nn = detail["nearest"][k]
        key = (f"{top.resnames[nn]}_{int(top.resids[nn])}", top.resnames[X])
        cov.add(key, ft["Fcov"][k], ft["Tcov"][k])
    detail["ft"] = ft
    detail["n_rad_centres"] = int(getattr(fs, "n_evaluated", len(top.heavy_idx)))
    return fsi, cov, lab, 1, detail


def bulk_frame(top: OracleTopology, pos, frc, dims, cov: CovMeans, lab: LabelCounts):
    """recipes/bulk_water.py:37-72 for one frame; accumulates into cov / lab."""
    fs = FrameState(top, pos, dims)
    recorded = []
    detail = dict(recorded=recorded, labels=[], donates=[], accepts=[])
    for X in top.water_centres:
        X = int(X)
        shell = fs.shell(X)
        labels = [top.resnames[n] for n in shell]
        if set(labels) == {top.resnames[X]}:
            fs.donations(X)
            for n in shell:
                fs.shell(n)
                fs.donations(n)
            don, acc = _hb_labels(top, fs, X, shell, labels)
            lab.add(top.resnames[X], top.resnames[X], labels, don, acc, len(labels))
            recorded.append(X)
            detail["labels"].append(labels)
            detail["donates"].append(don)
            detail["accepts"].append(acc)
    ft = forces_torques(top, np.asarray(pos, dtype=np.float32), np.asarray(frc, dtype=np.float32), recorded)
    for k, X in enumerate(recorded):
        cov.add((top.resnames[X], top.resnames[X]), ft["Fcov"][k], ft["Tcov"][k])
    detail["ft"] = ft
    return detail


MDA_CONVERSION = (1000**2 * 6.02214086e26) / ((6.022e23**2) * (1e-10**2))


def interfacial_run(top, positions, forces, dimensions, frames):
    """recipes/interfacial_solvent.py:181-270 up to (not including) the entropy
    post-processing: merged (frame_solvent_indices, CovMeans, LabelCounts, n)."""
    fsi = nested_dict()
    cov = CovMeans()
    lab = LabelCounts()
    n = 0
    for f in frames:
        r = interfacial_frame(top, positions[f], forces[f], dimensions[f], f)
        fsi.update(r[0])
        cov.merge(r[1])
        lab.merge(r[2])
        n += r[3]
    return fsi, cov, lab, n


def bulk_run(top, positions, forces, dimensions, frames):
    cov = CovMeans()
    lab = LabelCounts()
    for f in frames:
        bulk_frame(top, positions[f], forces[f], dimensions[f], cov, lab)
    return cov, lab
