"""Per-GPU engine: owns one C-ABI context and rebuilds the reference's result objects.

`WaterEntropyEngine` is the batched replacement of the reference's per-frame
worker `_entropy_per_step` (`recipes/interfacial_solvent.py:273-345`) and of the
frame loop of `recipes/bulk_water.py:37-72`: frames go in as float32 arrays,
the accumulated `CovarianceCollection`, `HBLabelCollection` and
`frame_solvent_indices` come out.  All arithmetic happens in the CUDA kernels
behind `include/waterentropy_b200.h`; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

from collections import defaultdict

import numpy as np

from . import _lib
from .analysis.HB_labels import HBLabelCollection
from .entropy.convariances import CovarianceCollection
from .system import PREFIX, RECIPE_BULK, RECIPE_INTERFACIAL, SystemTopology
from .utils.helpers import nested_dict

ENTRY_DTYPE = np.dtype(
    [
        ("hash", "<u8"), ("hash2", "<u8"), ("first_seen_inv", "<u8"), ("shell_count", "<u8"),
        ("nearest_resid", "<i4"), ("nearest_resname_id", "<i4"), ("n_labels", "<i4"), ("n_w", "<i4"),
        ("labels", "<u2", (32,)), ("donates", "<u4", (25,)), ("accepts", "<u4", (25,)), ("pad", "<u4", (50,)),
    ]
)
assert ENTRY_DTYPE.itemsize == 512

DBG = dict(shell_n=(0, np.int32), shell_idx=(1, np.int32), nearest=(2, np.int32), flags=(3, np.int32),
           acceptor=(4, np.int32), nbr_idx=(5, np.int32), nbr_dist=(6, np.float64), ft=(7, np.float64),
           rec=(8, np.int32))


def _host_ptr(a):
    """Raw host pointer of a contiguous numpy array or CPU torch tensor."""
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    return a.ctypes.data


def decode_cov_tables(top, recipe, s, c) -> CovarianceCollection:
    """Dense per-bin sums [B,2,3,3] + counts [B] -> the reference's CovarianceCollection
    (`entropy/convariances.py:10-19`): key (f"{resname}_{resid}", water resname), mean = sum / count."""
    out = CovarianceCollection()
    idx = np.nonzero(c)[0]
    # every bin's means in one array (the same float64 division per element as `set_sums`); the dictionaries
    # hold views of it, so `Vibrations.add_data` can scale and evaluate all bins with array operations
    means = np.asarray(s, dtype=np.float64)[idx] / np.asarray(c)[idx].astype(np.float64)[:, None, None, None]
    keys = []
    for k, b in enumerate(idx.tolist()):
        cls = b % top.n_classes
        centre_name = top.resname_table[top.class_resname_ids[cls]]
        if recipe == RECIPE_BULK:
            key = (centre_name, centre_name)
        else:
            rn, rid = top.pair_bins[b // top.n_classes]
            key = (f"{top.resname_table[rn]}_{rid}", centre_name)
        out.forces[key] = means[k, 0]
        out.torques[key] = means[k, 1]
        out.counts[key] = int(c[b])
        keys.append(key)
    out._stacked = (keys, means)
    return out


def decode_label_entries(top, recipe, entries) -> HBLabelCollection:
    """Device label-table entries -> the reference's HBLabelCollection
    (`analysis/HB_labels.py:15-156`).  Label codes become strings, the tuple is sorted the way
    Python sorts strings, and entries are fed in order of first appearance because N_w is frozen
    by the first insert (`analysis/HB_labels.py:58,71`).

    What `HBLabelCollection.add_counts` would do entry by entry, on plain Python lists (the columns of the
    structured array are converted once): this is most of the reference's unit of work on this path,
    `_entropy_per_step` - one frame per call, ~1,400 entries on C4 (34 -> 9 ms per call)."""
    out = HBLabelCollection()
    if len(entries) == 0:
        return out
    e = entries[np.argsort(~entries["first_seen_inv"], kind="stable")]
    cache = top.__dict__.setdefault("_we_label_strings", {})
    names = top.resname_table
    bulk = recipe == RECIPE_BULK
    by_name, by_resid = out.labelled_shell_counts, out.resid_labelled_shell_counts
    for n, codes, don, acc, rn, rid, n_w, count in zip(
            e["n_labels"].tolist(), e["labels"].tolist(), e["donates"].tolist(), e["accepts"].tolist(),
            e["nearest_resname_id"].tolist(), e["nearest_resid"].tolist(), e["n_w"].tolist(), e["shell_count"].tolist()):
        labels = []
        for c in codes[:n]:
            lab = cache.get(c)
            if lab is None:
                lab = cache[c] = top.label_string(c)
            labels.append(lab)
        key = tuple(sorted(labels))
        resname = names[rn]
        resid = resname if bulk else rid
        for entry in (by_name[resname][key], by_resid[resid][resname][key]):
            if "shell_count" in entry:
                entry["shell_count"] += count
            else:
                entry["shell_count"] = count
                entry["N_w"] = n_w
            for field, src in (("donates_to", don), ("accepts_from", acc)):
                dst = None
                for p in range(n):
                    v = src[p]
                    if v:
                        if dst is None:
                            dst = entry[field]
                        lab = labels[p]
                        dst[lab] = dst.get(lab, 0) + v
    return out


class WaterEntropyEngine:
    def __init__(self, system, recipe="interfacial", device=0, chunk_frames=0, table_log2=0,
                 keep_records=True, keep_debug=False, search_radius=0.0, max_box=0.0, force_upload=0,
                 keep_shells=False, eig_flavour=None, shells_only=False, pos_upload=0):
        """`eig_flavour`: "avx512" / "avx2" / None.  The reference's principal axes carry the eigenvector
        signs `numpy.linalg.eig` returns on the host it runs on; None replays the OpenBLAS kernel family
        of THIS host's NumPy (`_lib.host_eig_flavour`), so the off-diagonal covariance elements are the
        ones the reference would produce here."""
        self.lib = _lib.load()
        self.top = SystemTopology.from_universe(system)
        self.recipe = RECIPE_BULK if recipe in ("bulk", RECIPE_BULK) else RECIPE_INTERFACIAL
        t, self._keep = self.top.c_struct(self.recipe)
        if eig_flavour is None:
            self.eig_flavour, self.eig_source, self.eig_pinned = _lib.host_eig_flavour()
        else:
            self.eig_flavour = _lib._EIG_NAMES[str(eig_flavour).lower()]
            self.eig_source, self.eig_pinned = f"eig_flavour={eig_flavour}", True
        o = _lib.we_options(device=device, chunk_frames=chunk_frames, table_log2=table_log2,
                            keep_records=int(keep_records), keep_debug=int(keep_debug),
                            search_radius=search_radius, max_box=max_box, force_upload=force_upload,
                            keep_shells=int(keep_shells), eig_flavour=self.eig_flavour,
                            shells_only=int(shells_only), pos_upload=int(pos_upload))
        self._ctx = C.c_void_p()
        _lib.check(self.lib.we_create(C.byref(t), C.byref(o), C.byref(self._ctx)))
        self.device = device
        self.keep_records = keep_records
        self.keep_debug = keep_debug
        self.shells_only = bool(shells_only)
        self.frame_ids = []
        self._inflight = []  # (frames submitted up to and including this call, buffers): kept alive until retired

    @property
    def batch_frames(self):
        """Frames per kernel batch; a reader that submits whole batches keeps the device busiest."""
        return _lib.check(self.lib.we_batch_frames(self._ctx))

    # ------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self.lib.we_destroy(self._ctx)
            self._ctx = C.c_void_p()
        self._entry_buf = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        _lib.check(self.lib.we_reset(self._ctx))
        self.frame_ids = []
        self._inflight = []

    # ------------------------------------------------------------------
    @staticmethod
    def _check_frames(a, name, shape, cuda_device=None):
        """Layout the kernels assume: float32, C-contiguous, the right shape (and device)."""
        if hasattr(a, "data_ptr"):  # torch tensor
            import torch

            if a.dtype != torch.float32 or not a.is_contiguous():
                raise ValueError(f"{name} must be a contiguous float32 tensor, got {a.dtype}, contiguous={a.is_contiguous()}")
            if cuda_device is None and a.device.type != "cpu":
                raise ValueError(f"{name} is on {a.device}; submit() takes host memory (use submit_device)")
            if cuda_device is not None and (a.device.type != "cuda" or a.device.index != cuda_device):
                raise ValueError(f"{name} is on {a.device}; this engine runs on cuda:{cuda_device}")
        elif cuda_device is not None:
            raise ValueError(f"{name} must be a CUDA tensor")
        if tuple(a.shape) != tuple(shape):
            raise ValueError(f"{name} must have shape {tuple(shape)}, got {tuple(a.shape)}")

    def _retire(self):
        """Drop references to host buffers the device no longer reads (we_frames_retired)."""
        if self._inflight:
            done = int(self.lib.we_frames_retired(self._ctx))
            while self._inflight and self._inflight[0][0] <= done:
                self._inflight.pop(0)

    def wait_retired(self, n_frames):
        """Block until the first `n_frames` submitted frames no longer need their input buffers."""
        _lib.check(self.lib.we_wait_frames_retired(self._ctx, int(n_frames)))
        self._retire()

    def submit(self, positions, forces, dimensions, frame_ids):
        """Queue frames held in host memory (numpy arrays or CPU/pinned torch tensors).
        positions/forces float32 [F,N,3], dimensions float32 [F,6], frame_ids int [F].
        `forces` may be None for a shells_only engine."""
        n = len(frame_ids)
        if n == 0:
            return
        fid = np.ascontiguousarray(frame_ids, dtype=np.int64)
        if not hasattr(dimensions, "data_ptr"):
            dimensions = np.ascontiguousarray(dimensions, dtype=np.float32)
        if not hasattr(positions, "data_ptr"):
            positions = np.ascontiguousarray(positions, dtype=np.float32)
        if forces is not None and not hasattr(forces, "data_ptr"):
            forces = np.ascontiguousarray(forces, dtype=np.float32)
        shape = (n, self.top.n_atoms, 3)
        self._check_frames(positions, "positions", shape)
        self._check_frames(dimensions, "dimensions", (n, 6))
        if forces is None:
            if not self.shells_only:
                raise ValueError("forces are required (only a shells_only engine runs without them)")
        else:
            self._check_frames(forces, "forces", shape)
        self._retire()
        _lib.check(self.lib.we_submit(self._ctx, _host_ptr(positions), None if forces is None else _host_ptr(forces),
                                      _host_ptr(dimensions), fid.ctypes.data, n))
        self.frame_ids.extend(int(x) for x in fid)
        self._inflight.append((len(self.frame_ids), (positions, forces, dimensions, fid)))

    def submit_device(self, d_positions, d_forces, dimensions, frame_ids):
        """Queue frames whose positions / forces are CUDA tensors on this engine's device."""
        n = len(frame_ids)
        if n == 0:
            return
        fid = np.ascontiguousarray(frame_ids, dtype=np.int64)
        dims = np.ascontiguousarray(dimensions, dtype=np.float32)
        shape = (n, self.top.n_atoms, 3)
        self._check_frames(d_positions, "positions", shape, self.device)
        self._check_frames(d_forces, "forces", shape, self.device)
        self._check_frames(dims, "dimensions", (n, 6))
        self._retire()
        _lib.check(self.lib.we_submit_device(self._ctx, d_positions.data_ptr(), d_forces.data_ptr(),
                                             dims.ctypes.data, fid.ctypes.data, n))
        self.frame_ids.extend(int(x) for x in fid)
        self._inflight.append((len(self.frame_ids), (d_positions, d_forces, dims, fid)))

    def sync(self):
        try:
            _lib.check(self.lib.we_sync(self._ctx))
        finally:
            self._inflight = []

    def diag(self):
        d = _lib.we_diag()
        _lib.check(self.lib.we_get_diag(self._ctx, C.byref(d)))
        return {n: int(getattr(d, n)) for n, _ in d._fields_}

    # ---- measurement ---------------------------------------------------------
    KERNELS = ["k_bin", "k_scan", "k_scatter", "k_rad", "k_hb", "k_list1", "k_label", "k_cov",
               "k_rad_pass1", "k_rad_pass2", "k_seed_list2", "k_gate", "k_fetch_light"]

    def profile(self, enable=True):
        _lib.check(self.lib.we_profile(self._ctx, int(enable)))

    def kernel_times(self):
        """{kernel: (total ms, launches)} measured with CUDA events on the compute stream."""
        ms = np.zeros(len(self.KERNELS), dtype=np.float64)
        n = np.zeros(len(self.KERNELS), dtype=np.uint64)
        _lib.check(self.lib.we_kernel_times(self._ctx, ms.ctypes.data, n.ctypes.data))
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(self.KERNELS)}

    def span_start(self):
        _lib.check(self.lib.we_span_start(self._ctx))

    def span_stop(self):
        ms = C.c_double()
        _lib.check(self.lib.we_span_stop(self._ctx, C.byref(ms)))
        return ms.value

    # ---- raw tables --------------------------------------------------------
    def cov_tables(self):
        nb = self.lib.we_n_bins(self._ctx)
        s = np.zeros((nb, 2, 3, 3), dtype=np.float64)
        c = np.zeros(nb, dtype=np.uint64)
        _lib.check(self.lib.we_copy_cov(self._ctx, s.ctypes.data, c.ctypes.data))
        return s, c

    def snapshot_cov(self):
        """Enqueue a stream-ordered copy of the covariance tables (state after everything submitted
        so far); returns a ticket for `wait_cov`.  Does not wait."""
        return _lib.check(self.lib.we_snapshot_cov(self._ctx))

    def wait_cov(self, ticket):
        nb = self.lib.we_n_bins(self._ctx)
        s = np.zeros((nb, 2, 3, 3), dtype=np.float64)
        c = np.zeros(nb, dtype=np.uint64)
        _lib.check(self.lib.we_wait_cov(self._ctx, ticket, s.ctypes.data, c.ctypes.data))
        return s, c

    def vibrational_entropies(self, temperature=298):
        """{(nearest, water resname): (translational S [3], rotational S [3])} of every occupied bin, computed
        on the device from the covariance tables (entropy/vibrations.py:142-192, diagonalise=True)."""
        from .entropy.vibrations import Vibrations

        nb = self.lib.we_n_bins(self._ctx)
        out = np.zeros((nb, 2, 3), dtype=np.float64)
        _lib.check(self.lib.we_vib_entropy(self._ctx, float(temperature), Vibrations(temperature).mda_conversion(),
                                           out.ctypes.data))
        _, cnt = self.cov_tables()
        cov = decode_cov_tables(self.top, self.recipe, np.zeros((nb, 2, 3, 3)), cnt)
        keys = list(cov.counts.keys())
        return {k: (out[b, 0].copy(), out[b, 1].copy()) for k, b in zip(keys, np.nonzero(cnt)[0])}

    def orient_entropy(self, bulk=None):
        """`(resid_labelled_Sorient, resname_labelled_Sorient)` computed on the device from this context's
        label table (`we_orient_entropy`; reference `entropy/orientations.py:298-370`): the label entries
        never come to the host.  Rows are `[Sor, count, N_c, N_w, Nc_eff, pbias, Sor_Nc, Sor_Nw]`."""
        bulk = (self.recipe == RECIPE_BULK) if bulk is None else bool(bulk)
        top = self.top
        rank = getattr(top, "_we_label_rank", None)
        if rank is None:  # Python's string order over the whole label alphabet (topology only)
            nres = len(top.resname_table)
            codes = np.array([(p << 13) | r for p in range(len(PREFIX)) for r in range(nres)], dtype=np.int64)
            strings = [top.label_string(int(c)) for c in codes]
            order = sorted(range(len(codes)), key=strings.__getitem__)
            rank = np.zeros(65536, dtype=np.uint16)
            rank[codes[order]] = np.arange(len(order), dtype=np.uint16)
            top._we_label_rank = rank
        n_entries = _lib.check(self.lib.we_n_label_entries(self._ctx))
        tables = []
        for by_resname in ((1, 1) if bulk else (0, 1)):
            cap = max(1, n_entries)
            keys = np.zeros((cap, 2), dtype=np.int32)
            rows = np.zeros((cap, 8), dtype=np.float64)
            ng = C.c_int32(0)
            _lib.check(self.lib.we_orient_entropy(self._ctx, rank.ctypes.data, by_resname, cap, keys.ctypes.data,
                                                  rows.ctypes.data, C.byref(ng)))
            tables.append((keys[:ng.value], rows[:ng.value]))
        by_resid, by_resname_t = nested_dict(), nested_dict()
        for (keys, rows), table in zip(tables, (by_resid, by_resname_t)):
            for (resid, rn), row in zip(keys.tolist(), rows.tolist()):
                name = top.resname_table[rn]
                row[1] = int(row[1])
                if table is by_resname_t:
                    table[name] = row
                else:
                    table[name if bulk else int(resid)][name] = row
        return by_resid, by_resname_t

    def device_table_ptrs(self):
        a, b = C.c_void_p(), C.c_void_p()
        _lib.check(self.lib.we_device_tables(self._ctx, C.byref(a), C.byref(b)))
        return a.value, b.value, self.lib.we_n_bins(self._ctx)

    def label_entries(self, copy=True):
        """The label table's occupied entries (structured array, ENTRY_DTYPE).  Read into page-locked memory
        (tens of MB once several ranks' tables are merged).  `copy=False` returns a view of that staging
        buffer, valid until the next call on this engine (what the recipes use: they consume it at once)."""
        n = _lib.check(self.lib.we_n_label_entries(self._ctx))
        if n == 0:
            return np.zeros(0, dtype=ENTRY_DTYPE)
        buf = getattr(self, "_entry_buf", None)
        if buf is None or buf.numel() < n * ENTRY_DTYPE.itemsize:
            try:
                import torch

                buf = torch.empty(int(n * 1.25 + 1024) * ENTRY_DTYPE.itemsize, dtype=torch.uint8, pin_memory=True)
            except Exception:  # pragma: no cover
                buf = None
            self._entry_buf = buf
        if buf is None:  # pragma: no cover
            out = np.zeros(n, dtype=ENTRY_DTYPE)
            got = _lib.check(self.lib.we_copy_label_entries(self._ctx, out.ctypes.data, len(out)))
            return out[:got]
        view = buf.numpy().view(ENTRY_DTYPE)
        got = _lib.check(self.lib.we_copy_label_entries(self._ctx, view.ctypes.data, len(view)))
        return view[:got].copy() if copy else view[:got]

    def label_entries_device(self):
        """(device pointer, count) of this rank's compacted label entries (512 B each)."""
        p = C.c_void_p()
        n = _lib.check(self.lib.we_label_entries_device(self._ctx, C.byref(p)))
        return p.value, n

    def merge_label_entries_device(self, ptr, n):
        _lib.check(self.lib.we_merge_label_entries(self._ctx, ptr, n))

    def frame_records(self, k):
        n = _lib.check(self.lib.we_frame_record_count(self._ctx, k))
        out = np.zeros((max(n, 1), 2), dtype=np.int32)
        _lib.check(self.lib.we_copy_frame_records(self._ctx, k, out.ctypes.data, len(out)))
        return out[:n]

    def frame_records_range(self, k0, n):
        """(counts [n], pairs [total, 2]) of the submitted frames k0 .. k0 + n - 1 in one call."""
        counts = np.zeros(max(n, 1), dtype=np.int32)
        cap = max(1, n * max(1, len(self.top.centres)))
        if not hasattr(self, "_rec_buf") or len(self._rec_buf) < cap:
            self._rec_buf = np.empty((cap, 2), dtype=np.int32)
        total = self.lib.we_copy_frame_records_range(self._ctx, k0, n, counts.ctypes.data, self._rec_buf.ctypes.data, cap)
        if total < 0:
            _lib.check(int(total))
        return counts[:n], self._rec_buf[:total]

    def frame_shells(self, k):
        """{water atom: [shell atoms, distance order]} of the recorded waters of submitted frame k."""
        n = _lib.check(self.lib.we_frame_record_count(self._ctx, k))
        out = np.zeros((max(n, 1), 26), dtype=np.int32)
        _lib.check(self.lib.we_copy_frame_shells(self._ctx, k, out.ctypes.data, len(out)))
        rec = self.frame_records(k)
        return {int(rec[i, 0]): [int(v) for v in out[i, 1:1 + out[i, 0]]] for i in range(n)}

    def debug(self, k, field):
        fid, dt = DBG[field]
        nbytes = self.lib.we_debug_size(self._ctx, k, fid)
        if nbytes < 0:
            _lib.check(-10)
        out = np.zeros(nbytes // np.dtype(dt).itemsize, dtype=dt)
        _lib.check(self.lib.we_debug_copy(self._ctx, k, fid, out.ctypes.data, nbytes))
        if field in ("shell_idx", "nbr_idx", "nbr_dist"):
            out = out.reshape(-1, 25)
        elif field == "ft":
            out = out.reshape(-1, 6)
        elif field == "rec":
            out = out.reshape(-1, 4)
        return out

    def shells(self, k):
        """{atom index: [shell atom indices]} of every heavy atom in submitted frame k."""
        n = self.debug(k, "shell_n")
        idx = self.debug(k, "shell_idx")
        return {int(a): [int(v) for v in idx[s, : n[s]]] for s, a in enumerate(self.top.heavy_idx)}

    # ---- reference-shaped results ----------------------------------------------
    def covariances(self, tables=None) -> CovarianceCollection:
        """CovarianceCollection with mean = device sum / count (unscaled units)."""
        s, c = self.cov_tables() if tables is None else tables
        return decode_cov_tables(self.top, self.recipe, s, c)

    def hb_labels(self, entries=None) -> HBLabelCollection:
        return decode_label_entries(self.top, self.recipe, self.label_entries(copy=False) if entries is None else entries)

    def frame_solvent_indices(self):
        """frame_solvent_indices[frame][resname][resid] = [water atom idx ...] (ascending)."""
        top = self.top
        out = nested_dict()
        n = len(self.frame_ids)
        if n == 0:
            return out
        counts, pairs = self.frame_records_range(0, n)
        frame_of = np.repeat(np.arange(n), counts)
        keep = pairs[:, 1] >= 0
        frame_of, atoms, near = frame_of[keep], pairs[keep, 0], pairs[keep, 1]
        if len(atoms) == 0:
            return out
        # records are ascending in the water index inside a frame; a stable sort by (frame, resname, resid) of
        # the nearest solute atom keeps that order inside every residue's list, as upstream's appends do
        rn, rid = top.resname_id[near].astype(np.int64), top.resids[near].astype(np.int64)
        # one stable sort on a combined (frame, resname, resid) key instead of a three-key lexsort
        rid0 = int(rid.min())
        key = (frame_of.astype(np.int64) * (int(rn.max()) + 1) + rn) * (int(rid.max()) - rid0 + 1) + (rid - rid0)
        order = np.argsort(key, kind="stable")
        key, atoms = key[order], atoms[order]
        cut = np.nonzero(key[1:] != key[:-1])[0] + 1
        starts = np.concatenate([[0], cut])
        ends = np.concatenate([cut, [len(atoms)]])
        atoms_l = atoms.tolist()
        # the residue lists in one C-level pass, then one dict per (frame, resname)
        lists = list(map(atoms_l.__getitem__, map(slice, starts.tolist(), ends.tolist())))
        g_frame, g_rn, g_rid = frame_of[order][starts], rn[order][starts], rid[order][starts]
        cut2 = np.nonzero((g_frame[1:] != g_frame[:-1]) | (g_rn[1:] != g_rn[:-1]))[0] + 1
        s2 = np.concatenate([[0], cut2]).tolist()
        e2 = np.concatenate([cut2, [len(starts)]]).tolist()
        rid_l = g_rid.tolist()
        names = top.resname_table
        fids = self.frame_ids
        for s0, e0, fk, a in zip(s2, e2, g_frame[s2].tolist(), g_rn[s2].tolist()):
            out[fids[fk]][names[a]] = defaultdict(nested_dict, zip(rid_l[s0:e0], lists[s0:e0]))
        return out
