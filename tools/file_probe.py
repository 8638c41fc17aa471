"""Trajectory FILE -> result: the recipe call on an io.Universe backed by an AMBER NetCDF file of a synthetic
configuration (written here, read back from the page cache), with the NumPy reader and with the native threaded
reader (we_convert_be_f32).     python tools/file_probe.py [config] [frames] [--no-gpu]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.io import netcdf_file

from waterentropy_b200 import synthetic
from waterentropy_b200.io import Universe
from waterentropy_b200.io.amber import NetCDFTrajectory

args = [a for a in sys.argv[1:] if not a.startswith("--")]
cfg = args[0] if args else "C4"
n = int(args[1]) if len(args) > 1 else 96
gpu = "--no-gpu" not in sys.argv
s = synthetic.make_config(cfg)
path = f"/tmp/we_probe_{cfg}_{n}.nc"
t0 = time.perf_counter()
nc = netcdf_file(path, "w", version=2)
nc.createDimension("frame", None)
nc.createDimension("atom", s.n_atoms)
nc.createDimension("spatial", 3)
nc.createDimension("cell_spatial", 3)
nc.createDimension("cell_angular", 3)
xyz = nc.createVariable("coordinates", "f", ("frame", "atom", "spatial"))
frc = nc.createVariable("forces", "f", ("frame", "atom", "spatial"))
cl = nc.createVariable("cell_lengths", "d", ("frame", "cell_spatial"))
ca = nc.createVariable("cell_angles", "d", ("frame", "cell_angular"))
for k in range(n):
    p, f, d = s.frame(k)
    xyz[k] = p
    frc[k] = f
    cl[k] = d[:3]
    ca[k] = d[3:]
nc.close()
size = os.path.getsize(path)
print(f"{cfg}: {s.n_atoms} atoms, {n} frames, {size / 1e6:.0f} MB written in {time.perf_counter() - t0:.1f} s", flush=True)

for native in ("0", "1"):
    os.environ["WE_NATIVE_READER"] = native
    tr = NetCDFTrajectory(path)
    P = np.empty((32, s.n_atoms, 3), np.float32)
    F = np.empty((32, s.n_atoms, 3), np.float32)
    tr.read(range(min(32, n)), P, F)  # page cache warm
    t0 = time.perf_counter()
    for b0 in range(0, n, 32):
        tr.read(range(b0, min(n, b0 + 32)), P, F)
    dt = time.perf_counter() - t0
    print(f"reader alone, native={native}: {n / dt:8.0f} frames/s  {2 * 12 * s.n_atoms * n / dt / 1e9:6.2f} GB/s  "
          f"({s.n_waters * n / dt:.3g} water.frames/s)", flush=True)
    tr.close()

if gpu:
    from waterentropy_b200.recipes.interfacial_solvent import get_interfacial_water_orient_entropy

    for native in ("0", "1"):
        os.environ["WE_NATIVE_READER"] = native
        u = Universe()
        u._set_topology(s.names, s.types, s.resnames, s.resids, s.masses, s.charges, s.bonds)
        u._source = NetCDFTrajectory(path)
        get_interfacial_water_orient_entropy(u, 0, min(n, 32), 1)  # warm-up: context, page cache
        runs = []
        for _ in range(3):
            t0 = time.perf_counter()
            out = get_interfacial_water_orient_entropy(u, 0, n, 1)
            runs.append(time.perf_counter() - t0)
        dt = sorted(runs)[1]
        print(f"recipe on the file, native={native}: {dt:.3f} s for {n} frames = {s.n_waters * n / dt:.3g} water.frames/s "
              f"(runs {[round(r, 3) for r in runs]})", flush=True)
os.remove(path)
