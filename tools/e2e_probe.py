"""Break down one e2e step per phase (diagnostic; run under torchrun for N > 1)."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from waterentropy_b200 import synthetic, distributed as wdist
from waterentropy_b200.engine import WaterEntropyEngine
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1: dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
s = synthetic.make_config("C4"); F = 96
P, Fr, D = s.frames(rank * F, F); ids = list(range(rank * F, rank * F + F))
eng = WaterEntropyEngine(s, "interfacial", device=lr, keep_records=True)
hP, hF = torch.from_numpy(P).pin_memory(), torch.from_numpy(Fr).pin_memory()
for it in range(6):
    t = [time.perf_counter()]
    eng.submit(hP, hF, D, ids); t.append(time.perf_counter())
    eng.sync(); t.append(time.perf_counter())
    tabs = wdist.allreduce_cov(eng) if world > 1 else eng.cov_tables(); t.append(time.perf_counter())
    k0 = len(eng.frame_ids) - F
    n = sum(len(eng.frame_records(k0 + k)) for k in range(F)); t.append(time.perf_counter())
    if rank == 0: print(it, "submit %.1f sync %.1f tables %.1f records %.1f ms" % tuple(1e3 * (b - a) for a, b in zip(t, t[1:])), flush=True)
t0 = time.perf_counter(); e = eng.label_entries(); t1 = time.perf_counter()
if world > 1:
    g = wdist.allgather_entries(e); t2 = time.perf_counter(); m = wdist.merge_label_entries(g); t3 = time.perf_counter()
    if rank == 0: print("entries %.1f allgather %.1f merge %.1f ms" % (1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2)))
elif rank == 0: print("entries %.1f ms" % (1e3*(t1-t0)))
if world > 1: dist.destroy_process_group()
