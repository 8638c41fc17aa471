"""Where does the host-buffer arm lose time against the device-resident arm? (diagnostic)"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from waterentropy_b200 import synthetic
from waterentropy_b200.engine import WaterEntropyEngine

s = synthetic.make_config("C4"); F = 96
P, Fr, D = s.frames(0, F); ids = list(range(F))
hP, hF = torch.from_numpy(P).pin_memory(), torch.from_numpy(Fr).pin_memory()
dP, dF = hP.cuda(), hF.cuda()

def run(label, dev=False, env=None, **kw):
    for k, v in (env or {}).items(): os.environ[k] = str(v)
    eng = WaterEntropyEngine(s, "interfacial", device=0, **kw)
    sub = (lambda: eng.submit_device(dP, dF, D, ids)) if dev else (lambda: eng.submit(hP, hF, D, ids))
    for _ in range(3): sub()
    eng.sync()
    t0 = time.perf_counter()
    for _ in range(10): sub()
    eng.sync()
    dt = (time.perf_counter() - t0) / 10 * 1e3
    print("%-52s %.2f ms/step" % (label, dt), flush=True)
    eng.close()
    for k in (env or {}): os.environ.pop(k, None)

os.environ["WE_DIAG_TIMING"] = "1"
run("device-resident + records", dev=True, keep_records=True)
run("device-resident", dev=True, keep_records=False)
