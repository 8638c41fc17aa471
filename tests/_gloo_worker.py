"""Worker for tests/test_distributed_gloo.py (world_size 2, gloo, CPU): every rank analyses its
frame shard with the CPU oracle, packs the results into the device table formats, and the product's
combine path (all-reduce of the dense covariance table, all-gather + merge of the label entries)
must reproduce the single-process result."""
import hashlib
import os
import sys

import numpy as np
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path[:0] = [REPO, os.path.join(REPO, "oracle"), HERE]

import golden_io as gio  # noqa: E402
import we_oracle as wo  # noqa: E402
from waterentropy_b200 import distributed as wd  # noqa: E402
from waterentropy_b200.engine import ENTRY_DTYPE, decode_cov_tables, decode_label_entries  # noqa: E402
from waterentropy_b200.system import PREFIX, RECIPE_INTERFACIAL, SystemTopology  # noqa: E402


def encode(st, cov, lab, first_stamp):
    """oracle results -> (cov_sum, cov_cnt, entries) in the device layouts"""
    nb = st.n_pair_bins * st.n_classes
    s = np.zeros((nb, 2, 3, 3))
    c = np.zeros(nb, dtype=np.uint64)
    name_id = {n: i for i, n in enumerate(st.resname_table)}
    pair = {p: i for i, p in enumerate(st.pair_bins)}
    for (nearest, wat), n in cov.counts.items():
        rn, rid = nearest.rsplit("_", 1)
        b = pair[(name_id[rn], int(rid))] * st.n_classes + st.class_resname_ids.index(name_id[wat])
        s[b, 0], s[b, 1], c[b] = cov.forces[(nearest, wat)] * n, cov.torques[(nearest, wat)] * n, n
    rows = []
    for resid, a in lab.resid_labelled_shell_counts.items():
        for resname, bb in a.items():
            for key, e in bb.items():
                codes = []
                for label in key:
                    pfx = max((p for p in PREFIX if p and label.startswith(p)), key=len, default="")
                    codes.append((PREFIX.index(pfx) << 13) | name_id[label[len(pfx):]])
                codes.sort()
                strs = [st.label_string(cd) for cd in codes]
                ent = np.zeros((), dtype=ENTRY_DTYPE)
                h = hashlib.blake2b(repr((resid, resname, tuple(codes), e["N_w"])).encode(), digest_size=16).digest()
                ent["hash"], ent["hash2"] = int.from_bytes(h[:8], "little") | 1, int.from_bytes(h[8:], "little") | 1
                ent["nearest_resid"], ent["nearest_resname_id"] = resid, name_id[resname]
                ent["n_labels"], ent["n_w"], ent["shell_count"] = len(codes), e["N_w"], e["shell_count"]
                ent["labels"][: len(codes)] = codes
                ent["first_seen_inv"] = np.uint64(~np.uint64(first_stamp))
                for field, dst in (("donates_to", "donates"), ("accepts_from", "accepts")):
                    for label, cnt in e.get(field, {}).items():
                        ent[dst][strs.index(label)] += cnt
                rows.append(ent)
    return s, c, np.array(rows, dtype=ENTRY_DTYPE) if rows else np.zeros(0, dtype=ENTRY_DTYPE)


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    inp = gio.load_inputs("arginine")
    top = gio.oracle_topology(inp)
    st = SystemTopology(inp["masses"], inp["charges"], inp["resids"], list(inp["resnames"]),
                        list(inp["types"]), inp["bonds"])
    frames = list(range(0, 5))
    mine = wd.shard_frames(frames, rank, world)
    fsi, cov, lab, n = wo.interfacial_run(top, inp["positions"], inp["forces"], inp["dimensions"], mine)
    s, c, ent = encode(st, cov, lab, first_stamp=(mine[0] if mine else 0) * st.n_atoms)
    s, c = wd.allreduce_host(s, c)
    merged = wd.merge_label_entries(wd.allgather_entries(ent))
    got_cov = decode_cov_tables(st, RECIPE_INTERFACIAL, s, c)
    got_lab = decode_label_entries(st, RECIPE_INTERFACIAL, merged)
    # single-process truth
    fsi_all, cov_all, lab_all, _ = wo.interfacial_run(top, inp["positions"], inp["forces"], inp["dimensions"], frames)
    assert gio.plain(got_lab.resid_labelled_shell_counts) == gio.plain(lab_all.resid_labelled_shell_counts)
    assert gio.plain(got_lab.labelled_shell_counts) == gio.plain(lab_all.labelled_shell_counts)
    assert set(got_cov.counts) == set(cov_all.counts)
    for k in cov_all.counts:
        assert got_cov.counts[k] == cov_all.counts[k]
        assert np.allclose(got_cov.forces[k], cov_all.forces[k], rtol=1e-12)
        assert np.allclose(got_cov.torques[k], cov_all.torques[k], rtol=1e-12)
    gathered = [None] * world
    dist.all_gather_object(gathered, sorted(mine))
    assert sorted(sum(gathered, [])) == frames
    dist.barrier()
    if rank == 0:
        print("GLOO_OK", len(merged), int(c.sum()))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
