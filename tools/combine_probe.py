"""Time the parts of the end-of-job combine (torchrun, NCCL): all-reduce of the covariance tables, all-gather of
the label entries, device merge, label read-out.   torchrun --nproc-per-node N tools/combine_probe.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from waterentropy_b200 import distributed as wdist
from waterentropy_b200 import synthetic
from waterentropy_b200.engine import WaterEntropyEngine

rank, lr = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
s = synthetic.make_config("C4")
F = 96
P, Fr, D = s.frames(rank * F, F)
eng = WaterEntropyEngine(s, "interfacial", device=lr)
hP, hF = torch.from_numpy(P).pin_memory(), torch.from_numpy(Fr).pin_memory()
for rep in range(3):
    eng.submit(hP, hF, D, list(range(rank * F, rank * F + F)))
    eng.sync()
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter(); wdist.allreduce_cov(eng); torch.cuda.synchronize(); t1 = time.perf_counter()
    merged = wdist.merge_labels_nccl(eng, root_only=True); torch.cuda.synchronize(); t2 = time.perf_counter()
    if rank == 0:
        print(f"rep {rep}: allreduce_cov {1e3 * (t1 - t0):.2f} ms, merge_labels_nccl {1e3 * (t2 - t1):.2f} ms, {len(merged)} entries", flush=True)
    eng.reset()
dist.destroy_process_group()
