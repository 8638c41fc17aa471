"""cProfile of the reference's unit of work, `_entropy_per_step` (one frame per call), on C4 (diagnostic)."""
import os, sys, time, cProfile, pstats
sys.path.insert(0, "/root/repo")
import torch
from waterentropy_b200 import synthetic
from waterentropy_b200.io import Universe
from waterentropy_b200.recipes import interfacial_solvent as rec
s = synthetic.make_config("C4"); n = 12
P, F, D = s.frames(0, n)
hP, hF = torch.from_numpy(P).pin_memory(), torch.from_numpy(F).pin_memory()
u = Universe.from_arrays(s.names, s.types, s.resnames, s.resids, s.masses, s.charges, s.bonds, hP.numpy(), hF.numpy(), D)
u._source._keep = (hP, hF)
for i in range(3): rec._entropy_per_step((i, u))
pr = cProfile.Profile(); pr.enable()
t0 = time.perf_counter()
for i in range(3, 11): rec._entropy_per_step((i, u))
dt = (time.perf_counter() - t0) / 8
pr.disable()
print("per call %.1f ms" % (dt * 1e3))
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
