#!/usr/bin/env python
"""bench.py -- water-molecule.frames/sec of the per-frame hot path (RAD + HB + covariance).

Contract (see the task statement):  python bench.py --gpus N --steps K --warmup W
prints ONE JSON line on rank 0.  A "step" is one pass of the hot path over one batch
of synthetic frames of the workload BASELINE.json quotes the metric on (C4: 100k-water
solvated protein-DNA complex in a triclinic box); per-GPU work is fixed (weak scaling,
every rank analyses its own frames, tables combined with one NCCL exchange).

  value : whole-job water.frames/s with the frames already resident in HBM, timed with
          CUDA events on the library's compute stream (max over ranks).
  e2e   : the same metric through the public host-buffer API (`WaterEntropyEngine.submit`
          from pinned memory -> sync -> covariance table + label table + recorded-water
          lists back on the host, plus the cross-rank combine when N > 1).
  roofline / cpu_baseline : see DESIGN.md "Measurement".

`--impl reference` times the CPU restatement of the reference (oracle/, the reference is
pure Python on top of an uninstallable MDAnalysis) on all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

from waterentropy_b200 import synthetic  # noqa: E402

METRIC = "water-molecule.frames/sec (RAD+HB+covariance)"
UNIT = "water.frames/s"


def load_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU every 20 ms on a thread (NVML) while the
    timed region runs; falls back to one `nvidia-smi` query if NVML is unavailable."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, device):
        self.device = device
        self.samples = []
        self.mask = 0
        self.max_mhz = None
        self._stop = None
        self._thread = None

    def start(self):
        import threading

        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.device)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception:
            return
        self._stop = threading.Event()

        def loop():
            while not self._stop.is_set():
                try:
                    self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                    self.mask |= int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    try:
                        self.mask |= int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                    except Exception:
                        pass
                self._stop.wait(0.02)

        self._thread = threading.Thread(target=loop, daemon=True)
        self._thread.start()

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": []}
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=2)
        if not self.samples:
            try:
                q = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits",
                                    "-i", str(self.device)], capture_output=True, text=True, timeout=10).stdout
                a, b = [float(x) for x in q.strip().split(",")[:2]]
                out.update(sm_mhz=a, sm_max_mhz=b, samples=1, note="single nvidia-smi sample after the timed region")
            except Exception:
                pass
            return out
        out["sm_mhz"] = statistics.median(self.samples)
        out["samples"] = len(self.samples)
        out["reasons"] = sorted(v for k, v in self.REASONS.items() if self.mask & k)
        return out


# ----------------------------------------------------------------------------- CPU arm
def _oracle(threads):
    """The CPU restatement (oracle/), OpenMP over the centres of a frame with `threads` threads."""
    os.environ["OMP_NUM_THREADS"] = str(max(1, int(threads)))  # read when libgomp starts: before the first call
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import we_oracle as wo

    return wo


class CpuArm:
    """The reference's per-frame work on the host cores, with the C + NumPy port of its algorithm.

    What is timed is what `_entropy_per_step` (recipes/interfacial_solvent.py:273-345) evaluates for a frame -
    not every heavy atom: the reference memoises shells lazily, so only the solute united atoms, the
    interfacial waters they gate in, those waters' shell members and the donors of both are ever searched
    (`LazyFrameState`); for the bulk recipe every water is a centre (recipes/bulk_water.py:37-72).  Synthetic
    frame generation stays outside the clock.  A step is ONE frame, complete, on all cores (the per-centre
    work is independent, like the reference's frames); if complete frames would not fit the time budget the
    timed steps evaluate a random sample of the centres of the measured closure and scale by its size."""

    def __init__(self, cfg_name, scale, threads):
        self.wo = _oracle(threads)
        self.threads = threads
        self.sys = synthetic.make_config(cfg_name, scale=scale)
        s = self.sys
        self.top = self.wo.OracleTopology(s.masses, s.charges, s.resids, s.resnames, s.types, s.bonds)
        self.bulk = len(s.solute_atoms) == 0
        self.H = len(self.top.heavy_idx)
        self.closure = None  # heavy-atom centres the reference evaluates per frame (measured)

    def full_frame(self, k):
        """(seconds, centres evaluated, recorded waters) of one complete frame."""
        wo, top = self.wo, self.top
        P, F, D = self.sys.frame(k)                      # untimed
        t0 = time.perf_counter()
        if self.bulk:
            det = wo.bulk_frame(top, P, F, D, wo.CovMeans(), wo.LabelCounts())
            n_c = self.H
        else:
            det = wo.interfacial_frame(top, P, F, D, k, lazy=True)[4]
            n_c = det["n_rad_centres"]
        dt = time.perf_counter() - t0
        self.closure = n_c
        self._closure_frame = k
        return dt, n_c, len(det["recorded"])

    def sampled_frame(self, k, m, seed=7):
        """Seconds one frame would take, from `m` random centres of the closure (full per-centre cost:
        brute-force neighbour list + sort + RAD, the HBs of their donors, force/torque of the waters among
        them at the recorded fraction), scaled by closure / m."""
        wo, top = self.wo, self.top
        P, F, D = self.sys.frame(k)                      # untimed
        rng = np.random.default_rng([seed, k])
        m = min(m, self.H)
        centres = np.sort(rng.choice(top.heavy_idx, size=m, replace=False))
        t0 = time.perf_counter()
        r = wo.rad_frame(top, P, D, centres=centres)
        dx, dh, rows = [], [], []
        for i, a in enumerate(centres):
            for h in top.donors[int(a)]:
                dx.append(int(a)); dh.append(h); rows.append(i)
        if dx:
            wo.hb_frame(top, P, D, dx, dh, r["shell_n"][rows], r["shell_idx"][rows])
        waters = [int(a) for a in centres if top.is_water[a]][: max(1, m // 8)]
        wo.forces_torques(top, P, F, waters)
        dt = time.perf_counter() - t0
        return dt * self.closure / m


def cpu_rate(cfg_name, scale, threads, steps, warmup, budget_s, sample_centres=2048):
    """water.frames/s of the oracle port: `steps` timed frames after `warmup` (>= 1: the first frame is always
    complete and measures the closure).  Returns (rate, seconds per frame, description dict)."""
    arm = CpuArm(cfg_name, scale, threads)
    W = arm.sys.n_waters
    t_first, closure, recorded = arm.full_frame(0)
    total = max(1, warmup) + steps
    complete = t_first * (total - 1) <= budget_s
    times = []
    for k in range(1, total):
        if complete:
            dt = arm.full_frame(k)[0]
        else:
            dt = arm.sampled_frame(k, sample_centres)
        if k >= max(1, warmup):
            times.append(dt)
    if not times:
        times = [t_first]
    per_frame = statistics.median(times)
    rate = W / per_frame
    what = ("every timed step is one complete frame" if complete else
            f"timed steps: {sample_centres} random centres per frame at full per-centre cost, scaled to the closure "
            f"measured on a complete frame ({t_first:.1f} s, {W / t_first:.0f} water.frames/s unscaled)")
    info = {
        "sample": f"{'bulk' if arm.bulk else 'interfacial'} recipe, 1 frame per step on {threads} threads (OpenMP over centres); "
                  f"the reference's own per-frame work: {closure} of {arm.H} heavy-atom centres are searched "
                  f"({'every water' if arm.bulk else 'solute united atoms + interfacial waters + their shell members'}), "
                  f"{recorded} waters recorded; {what}; synthetic frame generation untimed",
        "centres_evaluated_per_frame": int(closure), "heavy_atom_centres": int(arm.H),
        "closure_fraction": closure / arm.H, "complete_frames": bool(complete),
        "first_complete_frame_s": t_first, "frame_s": [round(t, 3) for t in times],
        "all_centres_equivalent": rate * closure / arm.H,
    }
    return rate, per_frame, info


def reference_python_fixture(max_frames=4):
    """The UNMODIFIED reference (staged by oracle/stage_ref.py into oracle/_ref, running on oracle/mda_shim)
    through its own public entry point on its own arginine fixture (2,737 atoms, 900 waters): the only
    workload it finishes in seconds - its cost per centre grows with the atom count."""
    ref_dir = os.path.join(REPO, "oracle", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "waterEntropy")):
        return {"unavailable": "oracle/_ref is not staged (python oracle/stage_ref.py in the build container)"}
    code = (
        "import sys, os, time, json\n"
        f"REPO={REPO!r}\n"
        "sys.path[:0]=[os.path.join(REPO,'oracle','mda_shim'), os.path.join(REPO,'oracle','_ref'), REPO, os.path.join(REPO,'tests')]\n"
        "os.environ.setdefault('WE_SHIM_INLINE','1')\n"
        "import MDAnalysis, golden_io as gio\n"
        "import waterEntropy.recipes.interfacial_solvent as Inter\n"
        "inp=gio.load_inputs('arginine')\n"
        "u=MDAnalysis.Universe.from_arrays(inp['names'],inp['types'],inp['charges'],inp['masses'],inp['resids'],inp['resnames'],inp['bonds'],inp['positions'],inp['forces'],inp['dimensions'])\n"
        f"n={int(max_frames)}\n"
        "t0=time.perf_counter(); r=Inter.get_interfacial_water_orient_entropy(u,0,n,1); dt=time.perf_counter()-t0\n"
        "print(json.dumps({'seconds':dt,'frames':r[4],'recorded':int(sum(r[1].counts.values())),'file':Inter.__file__}))\n")
    try:
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
        d = json.loads(out.stdout.strip().splitlines()[-1])
    except Exception as exc:  # pragma: no cover
        return {"unavailable": f"reference run failed: {exc}"}
    return {"value": 900 * d["frames"] / d["seconds"], "unit": UNIT, "cores": 1, "kind": "reference",
            "sample": f"unmodified reference ({os.path.relpath(d['file'], REPO)}) on oracle/mda_shim, serial "
                      f"get_interfacial_water_orient_entropy over {d['frames']} frames of its arginine fixture "
                      f"(2,737 atoms, 900 waters, {d['recorded']} recorded); not comparable across workloads: "
                      "its per-centre cost grows with the atom count",
            "seconds": d["seconds"]}


def workload_config(args, s=None, recipe=None, input_bytes=None):
    """`config` of the JSON line: the same keys in both arms so the driver's same-config check holds."""
    s = s or synthetic.make_config(args.config, scale=args.scale)
    recipe = recipe or ("interfacial" if len(s.solute_atoms) else "bulk")
    input_bytes = input_bytes if input_bytes is not None else 24 * s.n_atoms * args.frames_per_step
    return {"workload": synthetic.CONFIGS[args.config]["label"], "scale": args.scale,
            "n_atoms": s.n_atoms, "n_waters": s.n_waters, "frames_per_step_per_gpu": args.frames_per_step,
            "recipe": recipe,
            "l2": f"inputs {input_bytes / 1e6:.0f} MB per step per GPU "
                  + ("> 126 MB L2 (no flush needed)" if input_bytes > 126e6 else "<= L2: reduce reliance, see DESIGN.md")}


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = args.cpu_cores or (os.cpu_count() or 1)
    rate, frame_s, info = cpu_rate(args.config, args.scale, cores, args.steps, args.warmup, args.cpu_budget)
    sample = info.pop("sample")
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": frame_s * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, **info,
                         "reference_python": reference_python_fixture()},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference = pure Python on MDAnalysis (not installable offline); the timed arm is oracle/ (C + NumPy "
                "restatement of the same algorithm, pinned against the reference's goldens) doing the reference's "
                "per-frame work on all host cores; cpu_baseline.reference_python is the unmodified reference itself "
                "on its own fixture",
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm
def time_recipe(s, recipe, hP, hF, D, Fps, W, repeats=3):
    """Wall clock of the public recipe entry point - `get_interfacial_water_orient_entropy(u, 0, F, 1)` (or the
    bulk recipe) - on an in-memory Universe holding the step's frames in page-locked memory: engine creation,
    topology upload, batch staging, kernels, table read-out, rebuilding the reference's Python objects and the
    entropy post-processing all inside the clock.  One warm-up call, one timed first call that creates its
    context, then the median of `repeats` calls that reuse it (the recipes keep the last context)."""
    from waterentropy_b200.io import Universe
    from waterentropy_b200.recipes.bulk_water import get_bulk_water_orient_entropy
    from waterentropy_b200.recipes.interfacial_solvent import get_interfacial_water_orient_entropy

    u = Universe.from_arrays(s.names, s.types, s.resnames, s.resids, s.masses, s.charges, s.bonds,
                             hP.numpy(), hF.numpy(), D)
    u._source._keep = (hP, hF)   # the frames are already page-locked: hand them over in place
    call = (lambda: get_interfacial_water_orient_entropy(u, 0, Fps, 1)) if recipe == "interfacial" else \
           (lambda: get_bulk_water_orient_entropy(u, 0, Fps, 1))
    from waterentropy_b200.recipes.interfacial_solvent import drop_cached_context

    call()                      # loads the CUDA module, builds the topology tables
    drop_cached_context()
    t0 = time.perf_counter()
    call()                      # a first call: creates its context (device workspace, pinned buffers)
    first = time.perf_counter() - t0
    runs = []
    for _ in range(repeats):    # further calls on the same system reuse that context (recipes' context cache)
        t0 = time.perf_counter()
        out = call()
        runs.append(time.perf_counter() - t0)
    sec = statistics.median(runs)
    return {"value": W * Fps / sec, "unit": UNIT, "seconds": sec, "runs_s": [round(t, 4) for t in runs], "frames": Fps,
            "first_call_seconds": round(first, 4),
            "call": ("get_interfacial_water_orient_entropy(u, 0, %d, 1)" % Fps) if recipe == "interfacial" else
                    ("get_bulk_water_orient_entropy(u, 0, %d, 1)" % Fps),
            "counts_bins": len(out[1].counts),
            "timed": "whole call, wall clock: staging in whole kernel batches, kernels, read-out, CovarianceCollection "
                     "rebuild, Orientations (device) + Vibrations; `seconds` = median of repeated calls, which reuse the "
                     "context of the call before (what per-frame `_entropy_per_step` calls and block-wise analyses "
                     "see); `first_call_seconds` = a call that creates its context"}


def time_file_recipe(s, recipe, P, Fr, D, Fps, W, repeats=3):
    """The same recipe call on a trajectory FILE: the step's frames written as an AMBER NetCDF-3 file (big-endian
    float32 records, forces in kcal like AMBER writes them), read back through `io.Universe`'s streaming reader
    into the pinned staging ring.  File in the page cache (just written); `numpy_reader_seconds` is the same call
    with the pure-NumPy conversion instead of the library's threaded one (`we_convert_be_f32`)."""
    import tempfile

    from scipy.io import netcdf_file

    from waterentropy_b200.io import Universe
    from waterentropy_b200.io.amber import NetCDFTrajectory
    from waterentropy_b200.recipes.bulk_water import get_bulk_water_orient_entropy
    from waterentropy_b200.recipes.interfacial_solvent import get_interfacial_water_orient_entropy

    path = os.path.join(tempfile.gettempdir(), "we_bench_%d.nc" % os.getpid())
    try:
        nc = netcdf_file(path, "w", version=2)
        for name, size in (("frame", None), ("atom", s.n_atoms), ("spatial", 3), ("cell_spatial", 3), ("cell_angular", 3)):
            nc.createDimension(name, size)
        xyz = nc.createVariable("coordinates", "f", ("frame", "atom", "spatial"))
        frc = nc.createVariable("forces", "f", ("frame", "atom", "spatial"))
        cl = nc.createVariable("cell_lengths", "d", ("frame", "cell_spatial"))
        ca = nc.createVariable("cell_angles", "d", ("frame", "cell_angular"))
        for k in range(Fps):
            xyz[k], frc[k], cl[k], ca[k] = P[k], Fr[k], D[k][:3], D[k][3:]
        nc.close()
        size = os.path.getsize(path)
        out = {}
        for native in ("1", "0"):
            os.environ["WE_NATIVE_READER"] = native
            u = Universe()
            u._set_topology(s.names, s.types, s.resnames, s.resids, s.masses, s.charges, s.bonds)
            u._source = NetCDFTrajectory(path)
            call = (lambda: get_interfacial_water_orient_entropy(u, 0, Fps, 1)) if recipe == "interfacial" else \
                   (lambda: get_bulk_water_orient_entropy(u, 0, Fps, 1))
            call()
            runs = []
            for _ in range(repeats):
                t0 = time.perf_counter()
                call()
                runs.append(time.perf_counter() - t0)
            out[native] = (statistics.median(runs), runs)
            u._source.close()
        os.environ.pop("WE_NATIVE_READER", None)
        sec, runs = out["1"]
        return {"value": W * Fps / sec, "unit": UNIT, "seconds": sec, "runs_s": [round(t, 4) for t in runs], "frames": Fps,
                "file_bytes": size, "numpy_reader_seconds": round(out["0"][0], 4),
                "timed": "whole recipe call on an io.Universe streaming an AMBER NetCDF file (page cache): threaded byte "
                         "swap + unit conversion into the pinned ring, upload, kernels, read-out, result objects"}
    finally:
        if os.path.exists(path):
            os.remove(path)


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from waterentropy_b200 import distributed as wdist
    from waterentropy_b200.engine import WaterEntropyEngine

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    # one slice of the host cores per rank: the submitting thread, the pinned-memory reader and the driver's
    # helper threads of a rank stay off the other ranks' cores (8 ranks share the box's 32 vCPUs)
    affinity = None
    if world > 1 and hasattr(os, "sched_setaffinity"):
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = max(1, len(cores) // world)
            slot = local_rank
            if os.environ.get("WE_BENCH_AFFINITY") == "reverse":   # probe: does the feed follow the cores?
                slot = world - 1 - local_rank
            elif os.environ.get("WE_BENCH_AFFINITY") == "swap":    # probe: halves exchanged
                slot = (local_rank + world // 2) % world
            mine = cores[slot * per:(slot + 1) * per] or cores
            os.sched_setaffinity(0, mine)
            affinity = [mine[0], mine[-1]]
        except OSError:
            affinity = None
    s = synthetic.make_config(args.config, scale=args.scale)
    N, W = s.n_atoms, s.n_waters
    recipe = "interfacial" if len(s.solute_atoms) else "bulk"  # solute-free configs (C2): bulk recipe
    Fps = args.frames_per_step
    # every rank analyses its own frames
    P, Fr, D = s.frames(rank * Fps, Fps)
    ids = list(range(rank * Fps, rank * Fps + Fps))
    input_bytes = P.nbytes + Fr.nbytes

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(x, op=None):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op or dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- device-resident arm (value) ----------------
    # every output of the path is produced in the timed region: covariance tables, label table and (interfacial
    # recipe) the per-frame recorded-water lists, which k_label writes to pinned host memory
    eng = WaterEntropyEngine(s, recipe, device=local_rank, keep_records=(recipe == "interfacial"), chunk_frames=args.chunk,
                             search_radius=args.search_radius)
    dP, dF = torch.from_numpy(P).cuda(), torch.from_numpy(Fr).cuda()
    for _ in range(args.warmup):
        eng.submit_device(dP, dF, D, ids)
        eng.sync()
    eng.profile(True)
    launches0 = eng.diag()["kernel_launches"]
    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    eng.span_start()
    for _ in range(args.steps):
        eng.submit_device(dP, dF, D, ids)
    dev_ms = eng.span_stop()
    eng.sync()
    barrier()
    dev_ms = maxr(dev_ms)
    diag = eng.diag()
    launches = diag["kernel_launches"] - launches0
    ktimes = eng.kernel_times()
    eng.profile(False)
    value = world * W * Fps * args.steps / (dev_ms * 1e-3)
    # the same K steps once more without the per-kernel event timers (the library then replays each batch as two
    # CUDA graphs): what a caller of we_submit_device sees; reported beside `value`, which stays the profiled region
    for _ in range(2):
        eng.submit_device(dP, dF, D, ids)
    eng.sync()
    barrier()
    eng.span_start()
    for _ in range(args.steps):
        eng.submit_device(dP, dF, D, ids)
    dev_ms_plain = maxr(eng.span_stop())
    eng.sync()
    # roofline of the dominant kernel.  k_rad runs as up to three dependent passes per batch
    # (solute atoms -> interfacial waters -> their shell members); they are one kernel and are
    # accounted together: time per batch = sum of the passes, launches = number of batches.
    rad_ms = sum(ktimes[k][0] for k in ("k_rad", "k_rad_pass1", "k_rad_pass2"))
    merged = {k: v for k, v in ktimes.items() if not k.startswith("k_rad")}
    merged["k_rad"] = (rad_ms, ktimes["k_rad"][1])
    top_k = max(merged, key=lambda k: merged[k][0])
    k_ms, k_n = merged[top_k]
    n_batches = max(1, ktimes["k_bin"][1])
    chunk = max(1, round(Fps * args.steps / n_batches))
    k_n = n_batches
    alg_bytes = (24 * N + 24) * chunk  # 24 B per atom + 24 B of box per frame
    peak, peak_src = load_peaks()
    achieved = alg_bytes / (k_ms / k_n * 1e-3) / 1e9 if k_n else 0.0
    traffic = None
    warp_inst = None
    tp = os.path.join(REPO, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            if tj.get("kernel") == top_k and tj.get("config") == args.config and tj.get("frames_per_launch") == chunk:
                traffic = tj.get("dram_bytes_per_launch")
                warp_inst = tj.get("warp_instructions_per_launch")
        except Exception:
            pass
    total_k = sum(v[0] for v in merged.values())
    roofline = {"bound": "hbm", "kernel": top_k, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "frames_per_launch": chunk,
                "kernel_ms_per_launch": k_ms / k_n if k_n else None,
                "kernel_share_of_step": k_ms / total_k if total_k else None,
                "kernel_ms_per_batch": {k: round(v[0] / n_batches, 4) for k, v in ktimes.items()}}
    if warp_inst and k_n and k_ms > 0:
        # the resource that actually binds the kernel: warp-instruction issue slots (4 schedulers per SM at the SM
        # clock); instructions from the committed ncu capture of the same command, time measured live above
        props = torch.cuda.get_device_properties(local_rank)
        sm_hz = 1e3 * float(getattr(props, "clock_rate", 0) or 1965000)  # kHz -> Hz; B200 boost clock as fallback
        slots = props.multi_processor_count * 4 * sm_hz
        roofline["issue"] = {"warp_instructions_per_launch": warp_inst, "issue_slots_per_s": slots,
                             "frac": warp_inst / (slots * (k_ms / k_n) * 1e-3),
                             "note": "the kernel is bound by instruction issue, not by HBM: share of the issue slots "
                                     "of the time its launches took"}
    eng.close()
    del dP, dF

    # ---------------- end-to-end arm (host buffers through the public API) ----------------
    eng = WaterEntropyEngine(s, recipe, device=local_rank, keep_records=(recipe == "interfacial"), chunk_frames=args.chunk,
                             force_upload=args.force_upload, search_radius=args.search_radius, pos_upload=args.pos_upload)
    hP, hF = torch.from_numpy(P).pin_memory(), torch.from_numpy(Fr).pin_memory()
    d2h = 0

    def read_step(ticket, k0):
        """host read of one finished step: its covariance snapshot + its recorded-water lists"""
        nonlocal d2h
        tables = eng.wait_cov(ticket)
        # the bulk recipe returns no per-frame lists (recipes/bulk_water.py:86-90)
        nrec = len(eng.frame_records_range(k0, Fps)[1]) if recipe == "interfacial" else 0
        d2h = tables[0].nbytes + tables[1].nbytes + Fps * 4 + nrec * 8
        return nrec

    def e2e_steps(k_steps):
        """K steps through the host API.  Every step copies its frames host -> device from pinned
        memory and its result (stream-ordered covariance snapshot + recorded-water lists) is read
        back on the host; the read of step k overlaps the kernels of step k+1, as in a real run
        that streams a long trajectory."""
        prev, nrec = None, 0
        for _ in range(k_steps):
            k0 = len(eng.frame_ids)
            eng.submit(hP, hF, D, ids)
            ticket = eng.snapshot_cov()
            if prev is not None:
                nrec = read_step(*prev)
            prev = (ticket, k0)
        eng.sync()  # drains the last batches' recorded-water lists
        if prev is not None:
            nrec = read_step(*prev)
        return nrec

    def e2e_finalize():
        """end of job, inside the timed region: the cross-rank combine (one NCCL all-reduce of the
        covariance tables, all-gather + device merge of the label entries) and the label table read"""
        if world > 1:
            wdist.allreduce_cov(eng)
            return wdist.merge_labels_nccl(eng, root_only=True)
        return eng.label_entries(copy=False)

    e2e_steps(args.warmup)
    e2e_finalize()
    eng.reset()  # the warm-up merge folded the other ranks' tables into this one
    # The host side of this pool is shared (other tenants' copies and page-cache traffic): single e2e
    # measurements of the same build differ by up to 2x while the device-timed value does not.  The K-step
    # region is therefore timed `--e2e-repeats` times (default 3), each a complete job (K steps + combine +
    # label read, tables reset in between); the fastest is reported and all are listed in e2e["runs_ms"].
    e2e_runs, combine_runs, combine_wait_runs, local_runs = [], [], [], []
    dma_bytes, light_atoms, pack_ms = 0, 0, 0.0
    for _ in range(max(1, args.e2e_repeats)):
        barrier()
        dma0, pack0 = eng.diag()["h2d_dma_bytes"], eng.diag()["host_pack_us"]
        t0 = time.perf_counter()
        nrec = e2e_steps(args.steps)
        t1 = time.perf_counter()
        entries = e2e_finalize()
        t2 = time.perf_counter()
        dg = eng.diag()  # counted by the library: bytes the copy engine moved, light atoms read in place
        dma_bytes, light_atoms = (dg["h2d_dma_bytes"] - dma0) / args.steps, dg["light_atoms_fetched"] / args.steps
        pack_ms = (dg["host_pack_us"] - pack0) / args.steps / 1e3
        local_runs.append(t1 - t0)  # this rank's own K steps, before the combine and the closing barrier
        barrier()
        e2e_runs.append(maxr(time.perf_counter() - t0))
        # the exchange itself = what the LAST rank to arrive spends in it (MIN over ranks); ranks that finish
        # their steps earlier (faster host feed) also wait in the collective for the others (MAX over ranks)
        combine_runs.append(maxr(t2 - t1, dist.ReduceOp.MIN if world > 1 else None))
        combine_wait_runs.append(maxr(t2 - t1))
        eng.reset()
    e2e_s = statistics.median(e2e_runs)
    # per rank: what the host fed this GPU and how busy its kernels were (the e2e limiter at N > 1 is the box's
    # aggregate host->device feed, not a collective)
    my_e2e_ms = statistics.median(local_runs) / args.steps * 1e3
    my_h2d = dma_bytes / (my_e2e_ms * 1e-3) / 1e9
    my_busy = (dev_ms / args.steps) / my_e2e_ms
    if world > 1:
        t = torch.tensor([my_h2d, my_busy], dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allr, t)
        per_rank = {"h2d_gbs": [round(float(x[0]), 2) for x in allr], "kernel_busy_frac": [round(float(x[1]), 3) for x in allr]}
    else:
        per_rank = {"h2d_gbs": [round(my_h2d, 2)], "kernel_busy_frac": [round(my_busy, 3)]}
    per_rank["cpu_affinity_of_rank0"] = affinity
    clk = clocks.stop()
    e2e_val = world * W * Fps * args.steps / e2e_s
    # bytes that cross PCIe per step: positions, boxes and frame ids are copied; the forces either are
    # copied whole (force_upload 1 / bulk recipe) or stay in pinned host memory and k_cov reads the atoms
    # of the recorded waters in place (zero-copy, the default for the interfacial recipe)
    zero_copy = recipe == "interfacial" and args.force_upload in (0, 3)
    packed = light_atoms > 0 or dma_bytes < 0.9 * P.nbytes
    position_path = ("packed: host threads gather the heavy atoms into pinned staging, the copy engine moves those; "
                     "the light atoms the path reads are read in place from the caller's pinned buffer"
                     if packed else "whole position frames copied")
    if zero_copy:
        h2d = dma_bytes + int(nrec) * 3 * 12 + light_atoms * 12
        force_path = "zero-copy: recorded waters' atoms read in place from pinned host memory"
    elif args.force_upload == 2:
        h2d = dma_bytes + int(nrec) * 3 * 12 + light_atoms * 12
        force_path = "recorded waters' atoms gathered by the host, uploaded compactly"
    else:
        h2d = dma_bytes + light_atoms * 12
        force_path = "whole force frames copied"
    final_bytes = int(entries.nbytes)
    eng.close()

    # ---------------- the call a user makes: the recipe itself (N = 1) ----------------
    e2e_recipe = None
    if world == 1 and not args.no_recipe:
        e2e_recipe = time_recipe(s, recipe, hP, hF, D, Fps, W)
        if not args.no_file:
            try:
                e2e_recipe["from_file"] = time_file_recipe(s, recipe, P, Fr, D, Fps, W)
            except Exception as exc:  # pragma: no cover - a full /tmp must not cost the bench line
                e2e_recipe["from_file"] = {"value": None, "unavailable": str(exc)}

    # ---------------- CPU baseline (rank 0, N = 1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = args.cpu_cores or (os.cpu_count() or 1)
        try:
            rate, _, info = cpu_rate(args.config, args.scale, cores, 1, 1, 25.0)
            sample = info.pop("sample")
            cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, **info,
                   "reference_python": reference_python_fixture()}
        except Exception as exc:  # pragma: no cover
            cpu = {"value": None, "unit": UNIT, "cores": cores, "kind": "port", "sample": f"failed: {exc}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, s, recipe, input_bytes),
            "value_arm": "device-resident inputs; all outputs produced in the timed region (covariance and label "
                         "tables on the device, per-frame recorded-water lists written to pinned host memory)",
            "ms_per_step_untimed_kernels": dev_ms_plain / args.steps,
            "clocks": clk,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "h2d_dma_bytes_per_step": int(dma_bytes), "h2d_zero_copy_atoms_per_step": int(light_atoms + (int(nrec) * 3 if zero_copy else 0)),
                    "host_input_bytes_per_step": int(input_bytes + Fps * 32), "position_path": position_path, "host_pack_ms_per_step": round(pack_ms, 3),
                    "force_path": force_path,
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s / args.steps * 1e3,
                    "runs_ms_per_step": [round(t / args.steps * 1e3, 3) for t in e2e_runs],
                    "combine_ms": round(statistics.median(combine_runs) * 1e3, 3),
                    "combine_runs_ms": [round(t * 1e3, 3) for t in combine_runs],
                    "combine_incl_wait_for_slowest_rank_ms": round(statistics.median(combine_wait_runs) * 1e3, 3),
                    "d2h_bytes_at_end_of_job": final_bytes,
                    "timed": "K x (submit from pinned host memory; stream-ordered covariance snapshot + recorded-water "
                             "lists read on the host, overlapping the next step) + end-of-job combine and label-table read "
                             "(combine_ms: all-reduce + gather + device merge + label read as the last rank to arrive sees it, inside the timed region); "
                             "MEDIAN of runs_ms_per_step (complete jobs)",
                    "recipe": e2e_recipe, "per_rank": per_rank},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "diag": {**{k: diag[k] for k in ("wide_searches", "shrink_retries", "table_entries", "eig_fallbacks")},
                     "recorded_waters_per_step": int(nrec)},
        }
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C4", choices=sorted(synthetic.CONFIGS))
    ap.add_argument("--scale", type=float, default=1.0, help="scale waters/residues of the config (tests only)")
    ap.add_argument("--frames-per-step", type=int, default=96)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--cpu-sample", type=int, default=2048, help="CPU arm: centres per sampled step (only when complete frames exceed --cpu-budget)")
    ap.add_argument("--cpu-budget", type=float, default=240.0, help="--impl reference: seconds the timed steps may take")
    ap.add_argument("--search-radius", type=float, default=0.0, help="first-pass candidate radius in A (0 = library default); tuning only, results do not depend on it")
    ap.add_argument("--no-file", action="store_true", help="skip e2e.recipe.from_file (the recipe on a NetCDF file written to the temp dir)")
    ap.add_argument("--no-recipe", action="store_true", help="skip the e2e.recipe measurement")
    ap.add_argument("--cpu-cores", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-repeats", type=int, default=3, help="timed repetitions of the K-step e2e job; the fastest is reported")
    ap.add_argument("--pos-upload", type=int, default=0, help="e2e arm: 0 auto, 1 whole frames, 2 packed (heavy atoms by DMA, light atoms read in place)")
    ap.add_argument("--force-upload", type=int, default=0, help="e2e arm: 0 auto, 1 whole frames, 2 host gather, 3 zero-copy")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist

            dist.destroy_process_group()


if __name__ == "__main__":
    main()
