"""The recipe call with `parallel=True` on N GPUs (NCCL) against the unmodified reference's return values.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/recipe_parallel_check.py

Every rank analyses its frame shard, the tables are all-reduced / merged on the device, `S_orient` is reduced on
the device from the merged label table; every rank must return what `tests/golden/*_reference.json.gz` holds for
the whole job (reference `recipes/interfacial_solvent.py:181-270`).  Needs >= 2 GPUs, so it is a tool run under
gpurun --gpus 2 (log in profiles/), not part of `pytest -m gpu` (one GPU).
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "tests")]
os.environ.setdefault("WE_EIG_FLAVOUR", "avx512")  # the goldens' BLAS family

import golden_io as gio  # noqa: E402
from waterentropy_b200.io import Universe  # noqa: E402
from waterentropy_b200.recipes.interfacial_solvent import get_interfacial_water_orient_entropy  # noqa: E402

RTOL = 1e-9


def close(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.allclose(a, b, rtol=RTOL, atol=RTOL * max(np.abs(b).max(), 1e-300))


def main():
    lr = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    rank, world = dist.get_rank(), dist.get_world_size()
    checked = 0
    for tag in ("arginine", "aspirin", "synth_tric"):
        inp, ref = gio.load_inputs(tag), gio.load_reference(tag)
        u = Universe.from_arrays(inp["names"], inp["types"], inp["resnames"], inp["resids"], inp["masses"],
                                 inp["charges"], inp["bonds"], inp["positions"], inp["forces"], inp["dimensions"])
        for key, api in ref["api"].items():
            s, e, st = [int(v) for v in key.split(":")]
            so, cov, vib, fsi, n = get_interfacial_water_orient_entropy(u, s, e, st, parallel=True)
            assert n == api["n_frames"]
            assert gio.plain(fsi) == api["frame_solvent_indices"], (tag, key, "frame_solvent_indices")
            got = gio.plain(so)
            assert got.keys() == api["Sorient"].keys()
            for resid in got:
                for resname in got[resid]:
                    a, b = got[resid][resname], api["Sorient"][resid][resname]
                    assert a[1] == b[1] and np.allclose(a, b, rtol=RTOL, atol=1e-12), (tag, key, resid, a, b)
            assert set(cov.counts) == set(api["cov_scaled"]["counts"])
            for k in cov.counts:
                assert cov.counts[k] == api["cov_scaled"]["counts"][k]
                assert close(cov.forces[k], api["cov_scaled"]["forces"][k]), (tag, key, k)
                assert close(cov.torques[k], api["cov_scaled"]["torques"][k]), (tag, key, k)
                assert np.allclose(vib.translational_S[k], api["vib"]["translational_S"][k], rtol=RTOL)
                assert np.allclose(vib.rotational_S[k], api["vib"]["rotational_S"][k], rtol=RTOL)
            checked += 1
    dist.barrier()
    print(f"rank {rank} of {world}: {checked} parallel recipe runs equal to the reference's return values", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
