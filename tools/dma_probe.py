"""Does an unrelated copy in flight slow the kernels of the device-resident arm? (diagnostic)
Runs the device arm alone, then with a background stream doing (a) host->device copies from pinned memory,
(b) device->device copies of the same size, (c) device->host copies."""
import os, sys, time, threading
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from waterentropy_b200 import synthetic
from waterentropy_b200.engine import WaterEntropyEngine

s = synthetic.make_config("C4"); F = 96
P, Fr, D = s.frames(0, F); ids = list(range(F))
dP, dF = torch.from_numpy(P).cuda(), torch.from_numpy(Fr).cuda()
eng = WaterEntropyEngine(s, "interfacial", device=0, keep_records=False)
nbytes = 90 * 1024 * 1024
hsrc = torch.empty(nbytes, dtype=torch.uint8).pin_memory(); hsrc.fill_(1)
ddst = torch.empty(nbytes, dtype=torch.uint8, device="cuda"); dsrc = torch.empty_like(ddst)
hdst = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
side = torch.cuda.Stream()
stop = False

def bg(kind):
    with torch.cuda.stream(side):
        while not stop:
            if kind == "h2d": ddst.copy_(hsrc, non_blocking=True)
            elif kind == "d2d": ddst.copy_(dsrc, non_blocking=True)
            elif kind == "d2h": hdst.copy_(dsrc, non_blocking=True)
            side.synchronize()
            time.sleep(0.0007)      # ~70 % duty cycle like the e2e arm

def run(label, kind=None):
    global stop
    stop = False
    th = None
    if kind:
        th = threading.Thread(target=bg, args=(kind,)); th.start()
    for _ in range(3): eng.submit_device(dP, dF, D, ids)
    eng.sync()
    eng.span_start()
    for _ in range(10): eng.submit_device(dP, dF, D, ids)
    ms = eng.span_stop() / 10
    eng.sync()
    stop = True
    if th: th.join()
    print("%-34s %.2f ms per 96-frame step" % (label, ms), flush=True)

run("device arm alone")
run("with background H2D (pinned)", "h2d")
run("with background D2D", "d2d")
run("with background D2H (pinned)", "d2h")
run("device arm alone (again)")
eng.close()
