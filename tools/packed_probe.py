"""Packed vs whole-frame uploads, submit-only loop (diagnostic): wall per 96-frame step, per-kernel device
times (profile mode), host time per submit call."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from waterentropy_b200 import synthetic
from waterentropy_b200.engine import WaterEntropyEngine

s = synthetic.make_config("C4"); F = 96
P, Fr, D = s.frames(0, F); ids = list(range(F))
hP, hF = torch.from_numpy(P).pin_memory(), torch.from_numpy(Fr).pin_memory()
for name, pu, fu, fm in (("whole", 1, 0, 0), ("packed fm0", 2, 0, 0), ("packed fm1", 2, 0, 1), ("packed fm2", 2, 0, 2)):
    os.environ["WE_FETCH_MODE"] = str(fm)
    for prof in (False, True):
        eng = WaterEntropyEngine(s, "interfacial", device=0, pos_upload=pu, force_upload=fu)
        for _ in range(3): eng.submit(hP, hF, D, ids)
        eng.sync(); eng.profile(prof)
        d0 = eng.diag()
        t0 = time.perf_counter(); hs = []
        for _ in range(10):
            t1 = time.perf_counter(); eng.submit(hP, hF, D, ids); hs.append(time.perf_counter() - t1)
        t2 = time.perf_counter()
        eng.sync(); dt = (time.perf_counter() - t0) / 10 * 1e3
        d1 = eng.diag()
        print("%-14s prof=%d wall %.2f ms/step; host in submit %.2f ms/step (last sync %.2f ms); pack %.2f ms/step" % (
            name, prof, dt, sum(hs) / 10 * 1e3, (time.perf_counter() - t2) * 1e3, (d1["host_pack_us"] - d0["host_pack_us"]) / 10 / 1e3), flush=True)
        if prof:
            kt = eng.kernel_times(); nb = max(1, kt["k_bin"][1])
            print("   per batch:", {k: round(v[0] / nb, 4) for k, v in kt.items()}, "sum %.3f" % (sum(v[0] for v in kt.values()) / nb), flush=True)
        eng.close()
