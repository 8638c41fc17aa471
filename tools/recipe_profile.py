"""Where the wall clock of the public recipe call goes (cProfile), C4 by default:
    python tools/recipe_profile.py [config] [frames]"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from waterentropy_b200 import synthetic
from waterentropy_b200.io import Universe
from waterentropy_b200.recipes.interfacial_solvent import get_interfacial_water_orient_entropy

cfg = sys.argv[1] if len(sys.argv) > 1 else "C4"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 96
s = synthetic.make_config(cfg)
P, F, D = s.frames(0, n)
hP, hF = torch.from_numpy(P).pin_memory(), torch.from_numpy(F).pin_memory()
u = Universe.from_arrays(s.names, s.types, s.resnames, s.resids, s.masses, s.charges, s.bonds, hP.numpy(), hF.numpy(), D)
u._source._keep = (hP, hF)
get_interfacial_water_orient_entropy(u, 0, n, 1)  # warm-up
t0 = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
get_interfacial_water_orient_entropy(u, 0, n, 1)
pr.disable()
print("wall", round(time.perf_counter() - t0, 3), "s for", n, "frames of", cfg)
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
