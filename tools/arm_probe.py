"""Per-kernel times of one step in the device-resident arm vs the host-buffer (e2e) arms (diagnostic)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from waterentropy_b200 import synthetic
from waterentropy_b200.engine import WaterEntropyEngine

s = synthetic.make_config("C4"); F = 96
P, Fr, D = s.frames(0, F); ids = list(range(F))
hP, hF = torch.from_numpy(P).pin_memory(), torch.from_numpy(Fr).pin_memory()
dP, dF = hP.cuda(), hF.cuda()
for name, kw, dev, gate, cov in (("device, no records", dict(keep_records=False), True, 1, 1),
                                 ("device + records", dict(keep_records=True), True, 1, 1),
                                 ("host zero-copy + records", dict(keep_records=True), False, 1, 1),
                                 ("host zero-copy, no records", dict(keep_records=False), False, 1, 1),
                                 ("host zero-copy + records, ungated", dict(keep_records=True), False, 0, 1)):
    os.environ["WE_COPY_GATE"] = str(gate); os.environ["WE_COV_STREAM"] = str(cov)
    for prof in (False, True):
        eng = WaterEntropyEngine(s, "interfacial", device=0, **kw)
        sub = (lambda: eng.submit_device(dP, dF, D, ids)) if dev else (lambda: eng.submit(hP, hF, D, ids))
        for _ in range(3): sub()
        eng.sync(); eng.profile(prof)
        t0 = time.perf_counter()
        for _ in range(10): sub()
        eng.sync(); dt = (time.perf_counter() - t0) / 10 * 1e3
        if prof:
            kt = eng.kernel_times(); nb = max(1, kt["k_bin"][1])
            print("   per batch:", {k: round(v[0] / nb, 4) for k, v in kt.items()}, "sum %.3f" % (sum(v[0] for v in kt.values()) / nb), flush=True)
        else:
            print("%-28s wall %.2f ms/step" % (name, dt), flush=True)
        eng.close()
