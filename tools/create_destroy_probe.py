"""Host time of context creation, first submission, read-out and destruction on C4 (diagnostic; typical 35 / 53 / 12 / 21 ms,
with sporadic 0.2-0.7 s spikes in any of them on the shared GPU boxes: pinned-memory operations of the host)."""
import sys, time
sys.path.insert(0, "/root/repo")
import torch
from waterentropy_b200 import synthetic
from waterentropy_b200.engine import WaterEntropyEngine
s = synthetic.make_config("C4")
P, F, D = s.frames(0, 1)
hP, hF = torch.from_numpy(P).pin_memory(), torch.from_numpy(F).pin_memory()
for i in range(5):
    t0 = time.perf_counter(); eng = WaterEntropyEngine(s, "interfacial", device=0); t1 = time.perf_counter()
    eng.submit(hP, hF, D, [0]); t2 = time.perf_counter(); eng.sync(); t3 = time.perf_counter()
    e = eng.hb_labels(); c = eng.covariances(); t4 = time.perf_counter()
    eng.close(); t5 = time.perf_counter()
    print(f"create {1e3*(t1-t0):.1f}  submit {1e3*(t2-t1):.1f}  sync {1e3*(t3-t2):.1f}  read {1e3*(t4-t3):.1f}  close {1e3*(t5-t4):.1f} ms", flush=True)
