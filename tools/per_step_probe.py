"""Wall clock of the reference's unit of work, `_entropy_per_step((index, system))` - one frame per call - on C4,
with and without the recipes' context cache:   python tools/per_step_probe.py [config] [calls]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from waterentropy_b200 import synthetic
from waterentropy_b200.io import Universe
from waterentropy_b200.recipes import interfacial_solvent as rec

cfg = sys.argv[1] if len(sys.argv) > 1 else "C4"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
s = synthetic.make_config(cfg)
P, F, D = s.frames(0, n)
hP, hF = torch.from_numpy(P).pin_memory(), torch.from_numpy(F).pin_memory()
u = Universe.from_arrays(s.names, s.types, s.resnames, s.resids, s.masses, s.charges, s.bonds, hP.numpy(), hF.numpy(), D)
u._source._keep = (hP, hF)
rec._entropy_per_step((0, u))  # CUDA module, topology tables
for cache in ("1", "0"):
    os.environ["WE_ENGINE_CACHE"] = cache
    rec.drop_cached_context()
    ts = []
    for i in range(n):
        t0 = time.perf_counter()
        rec._entropy_per_step((i, u))
        ts.append(time.perf_counter() - t0)
    print(f"{cfg} WE_ENGINE_CACHE={cache}: first call {ts[0] * 1e3:.1f} ms, later calls median {sorted(ts[1:])[len(ts[1:]) // 2] * 1e3:.1f} ms", flush=True)
rec.drop_cached_context()
