"""Host time of every call of a short streaming run (C4, batches of `batch_frames`): where the fixed cost
of a fresh context goes.   python tools/submit_probe.py [config] [n_batches]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from waterentropy_b200 import synthetic
from waterentropy_b200.engine import WaterEntropyEngine

cfg = sys.argv[1] if len(sys.argv) > 1 else "C4"
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 6
s = synthetic.make_config(cfg)
t0 = time.perf_counter()
eng = WaterEntropyEngine(s, "interfacial", device=0)
B = eng.batch_frames
print(f"create {time.perf_counter() - t0:.3f} s, batch_frames {B}")
P, F, D = s.frames(0, B)
hP, hF = torch.from_numpy(P).pin_memory(), torch.from_numpy(F).pin_memory()
for rep in range(2):
    for k in range(nb):
        t0 = time.perf_counter()
        eng.submit(hP, hF, D, list(range(k * B, k * B + B)))
        print(f"rep {rep} submit {k}: {1e3 * (time.perf_counter() - t0):.2f} ms")
    t0 = time.perf_counter()
    eng.sync()
    print(f"rep {rep} sync: {1e3 * (time.perf_counter() - t0):.2f} ms")
    t0 = time.perf_counter()
    e = eng.label_entries()
    print(f"rep {rep} label_entries ({len(e)}): {1e3 * (time.perf_counter() - t0):.2f} ms")
    t0 = time.perf_counter()
    fsi = eng.frame_solvent_indices()
    print(f"rep {rep} frame_solvent_indices: {1e3 * (time.perf_counter() - t0):.2f} ms")
    eng.reset()
