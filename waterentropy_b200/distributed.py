"""Frame sharding and table combination across the GPUs of one box.

The reference parallelises over frames with dask (`client.map` of one frame per
worker, `recipes/interfacial_solvent.py:219-249`) and merges the per-frame
Python objects on the client.  Here every rank (one process per GPU, launched
with torchrun) analyses a contiguous block of frames and the per-GPU tables are
combined once at the end:

* dense covariance sums + counts  -> one all-reduce (SUM) on the device tables
  (replaces `CovarianceCollection.merge`, `entropy/convariances.py:86-108`);
* sparse label-count entries      -> all-gather of the compact entry arrays,
  merged by their 128-bit key hash (replaces `HBLabelCollection.merge`,
  `analysis/HB_labels.py:158-296`);
* per-frame recorded-water lists  -> gathered as plain objects (tiny).
"""
from __future__ import annotations

import numpy as np


def world():
    """(rank, world_size) of the initialised torch.distributed job, else (0, 1)."""
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:  # pragma: no cover
        pass
    return 0, 1


def shard_frames(indices, rank, world_size):
    """Contiguous block of `indices` for `rank` (keeps frame order per rank)."""
    indices = list(indices)
    n = len(indices)
    base, extra = divmod(n, world_size)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return indices[lo:hi]


def merge_label_entries(entry_arrays):
    """Merge structured label-entry arrays (one per rank) by their 128-bit key (hash, hash2):
    counts add, `first_seen_inv` takes the maximum (= earliest first appearance), the key
    fields come from any member of the group.  Vectorised (sort + reduceat)."""
    parts = [e for e in entry_arrays if len(e)]
    if not parts:
        return entry_arrays[0][:0]
    allv = np.concatenate(parts)
    order = np.lexsort((allv["hash2"], allv["hash"]))
    allv = allv[order]
    new = np.ones(len(allv), dtype=bool)
    new[1:] = (allv["hash"][1:] != allv["hash"][:-1]) | (allv["hash2"][1:] != allv["hash2"][:-1])
    starts = np.nonzero(new)[0]
    out = allv[starts].copy()
    out["shell_count"] = np.add.reduceat(allv["shell_count"], starts)
    out["donates"] = np.add.reduceat(allv["donates"], starts, axis=0)
    out["accepts"] = np.add.reduceat(allv["accepts"], starts, axis=0)
    out["first_seen_inv"] = np.maximum.reduceat(allv["first_seen_inv"], starts)
    return out


def _device_tensor(ptr, n, dtype, device):
    import torch

    class _Wrap:
        pass

    w = _Wrap()
    w.__cuda_array_interface__ = {
        "shape": (n,), "typestr": np.dtype(dtype).str, "data": (ptr, False), "version": 2,
    }
    return torch.as_tensor(w, device=f"cuda:{device}")


def allreduce_host(s, c):
    """SUM all-reduce of host copies of the covariance tables (gloo / CPU path)."""
    import torch
    import torch.distributed as dist

    ts = torch.from_numpy(s)
    tc = torch.from_numpy(c.view(np.int64))
    dist.all_reduce(ts, op=dist.ReduceOp.SUM)
    dist.all_reduce(tc, op=dist.ReduceOp.SUM)
    return s, c


def allreduce_cov(eng):
    """SUM all-reduce of the dense covariance tables over the job: in place on the device
    tables through NCCL (NVLink / NVSwitch), or on host copies for non-NCCL backends."""
    import torch
    import torch.distributed as dist

    if dist.get_backend() == "nccl":
        p_sum, p_cnt, nb = eng.device_table_ptrs()
        t_sum = _device_tensor(p_sum, nb * 18, np.float64, eng.device)
        t_cnt = _device_tensor(p_cnt, nb, np.int64, eng.device)
        torch.cuda.synchronize(eng.device)
        dist.all_reduce(t_sum, op=dist.ReduceOp.SUM)
        dist.all_reduce(t_cnt, op=dist.ReduceOp.SUM)
        torch.cuda.synchronize(eng.device)
        return eng.cov_tables()
    return allreduce_host(*eng.cov_tables())


def allgather_entries(entries):
    """All-gather variable-length structured arrays (padded to the largest)."""
    import torch
    import torch.distributed as dist

    ws = dist.get_world_size()
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    n = torch.tensor([len(entries)], dtype=torch.int64, device=dev)
    sizes = [torch.zeros_like(n) for _ in range(ws)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    buf = np.zeros(cap, dtype=entries.dtype)
    buf[: len(entries)] = entries
    t = torch.from_numpy(buf.view(np.uint8)).to(dev)
    outs = [torch.zeros_like(t) for _ in range(ws)]
    dist.all_gather(outs, t)
    return [o.cpu().numpy().view(entries.dtype)[:k] for o, k in zip(outs, sizes)]


def merge_labels_nccl(eng, root_only=False, download=True):
    """Job-wide label table, merged on the device: all-gather of the compacted entries over NCCL,
    then `we_merge_label_entries` folds the other ranks' entries into this rank's table.
    Returns the merged entries (host structured array); with `root_only` only rank 0 merges and
    downloads (the other ranks return their own entries)."""
    import torch
    import torch.distributed as dist

    ws, rank = dist.get_world_size(), dist.get_rank()
    dev = torch.device("cuda", eng.device)
    ptr, n = eng.label_entries_device()
    sizes_t = torch.zeros(ws, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(sizes_t, torch.tensor([n], dtype=torch.int64, device=dev))
    sizes = [int(x) for x in sizes_t.tolist()]
    cap = max(max(sizes), 1)
    mine = torch.empty(cap * 512, dtype=torch.uint8, device=dev)
    if n:
        mine[: n * 512].copy_(_device_tensor(ptr, n * 512, np.uint8, eng.device))
    if root_only:
        # only rank 0 needs the other ranks' entries: one gather over NVLink instead of an all-gather
        outs = [torch.empty_like(mine) for _ in range(ws)] if rank == 0 else None
        dist.gather(mine, outs, dst=0)
        torch.cuda.synchronize(dev)
        if rank != 0:
            return eng.label_entries(copy=False)[:0]
    else:
        flat = torch.empty(ws * cap * 512, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(flat, mine)
        torch.cuda.synchronize(dev)
        outs = [flat[r * cap * 512:(r + 1) * cap * 512] for r in range(ws)]
    for r in range(ws):
        if r != rank and sizes[r]:
            eng.merge_label_entries_device(outs[r].data_ptr(), sizes[r])
    if not download:  # the merged table stays on the device (Orientations.add_device reduces it there)
        return None
    return eng.label_entries(copy=False)


def combine(eng):
    """(CovarianceCollection, HBLabelCollection, frame_solvent_indices) of the whole job."""
    cov, entries, fsi = combine_raw(eng)
    return cov, eng.hb_labels(entries), fsi


def combine_raw(eng, download_entries=True):
    """(CovarianceCollection, merged label entries, frame_solvent_indices) of the whole job.  With NCCL the
    label tables are merged on the device; `download_entries=False` then leaves the merged table there and
    returns None in its place."""
    import torch.distributed as dist

    tables = allreduce_cov(eng)
    if dist.get_backend() == "nccl":
        merged = merge_labels_nccl(eng, download=download_entries)
    else:
        merged = merge_label_entries(allgather_entries(eng.label_entries()))
    fsi_local = {k: {a: dict(b) for a, b in v.items()} for k, v in eng.frame_solvent_indices().items()}
    gathered = [None] * dist.get_world_size()
    dist.all_gather_object(gathered, fsi_local)
    from .utils.helpers import nested_dict

    fsi = nested_dict()
    for part in gathered:
        for frame, by_name in part.items():
            for resname, by_resid in by_name.items():
                for resid, atoms in by_resid.items():
                    fsi[frame][resname][resid] = atoms
    return eng.covariances(tables), merged, fsi
