"""CPU tests (no GPU): host-side logic, numerics header, ABI surface, consumers."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import golden_io as gio
import we_oracle as wo
from waterentropy_b200 import _build, _lib
from waterentropy_b200.analysis.HB_labels import HBLabelCollection
from waterentropy_b200.entropy.convariances import CovarianceCollection
from waterentropy_b200.entropy.orientations import Orientations
from waterentropy_b200.entropy.vibrations import Vibrations
from waterentropy_b200.system import SystemTopology

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hostlib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("hn") / "libhost_numerics.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fno-builtin", "-ffp-contract=off", "-shared",
                           "-fPIC", "-o", so, os.path.join(REPO, "tests", "host_numerics.cpp")])
    return ctypes.CDLL(so)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_powf2_matches_libm(hostlib):
    libm = ctypes.CDLL("libm.so.6")
    libm.powf.restype = ctypes.c_float
    libm.powf.argtypes = [ctypes.c_float, ctypes.c_float]
    rng = np.random.default_rng(1)
    x = np.concatenate([
        rng.uniform(0.0, 30.0, 200_000), rng.uniform(0.0, 1e-3, 1000), 10.0 ** rng.uniform(-30, 18, 5000),
        [0.0, 1.0, 2.0, 1e-40, 3e-39, 1.5e19, 3e19, np.inf],
    ]).astype(np.float32)
    out = np.empty_like(x)
    hostlib.h_powf2(_p(x), len(x), _p(out))
    ref = np.array([libm.powf(float(v), 2.0) for v in x], dtype=np.float32)
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))
    fast = np.empty_like(x)
    hostlib.h_powf2_fast(_p(x), len(x), _p(fast))
    assert np.array_equal(fast.view(np.uint32), ref.view(np.uint32))
    # and NumPy's float32 scalar power is that same libm call
    assert all(np.float32(v) ** 2 == r for v, r in zip(x[:20000], ref[:20000]))


def test_pow2_f64_matches_libm_and_numpy(hostlib):
    """we_pow2 (we_numerics.cuh) == glibc pow(x, 2.0) == NumPy's float64 scalar `x ** 2` (RAD.py:53-57, HB.py:96)."""
    rng = np.random.default_rng(11)
    x = np.concatenate([
        np.exp(rng.uniform(-7, 7, 150_000)),
        rng.uniform(0.05, 12.0, 150_000),                 # distances / inverse distances in Angstrom
        1.0 + rng.uniform(-1e-12, 1e-12, 1000), [1.0, 0.5, 2.0, float.fromhex("0x1.b0860c015e44p-9")],
    ]).astype(np.float64)
    out = np.empty_like(x)
    hostlib.h_pow2(_p(x), len(x), _p(out))
    libm = ctypes.CDLL("libm.so.6")
    libm.pow.restype = ctypes.c_double
    libm.pow.argtypes = [ctypes.c_double, ctypes.c_double]
    ref = np.array([libm.pow(float(v), 2.0) for v in x])
    assert np.array_equal(out.view(np.uint64), ref.view(np.uint64))
    assert (ref != x * x).sum() > 50                      # the emulation is not just x*x
    assert all(np.float64(v) ** 2 == r for v, r in zip(x[:20000], ref[:20000]))


def test_pow_tables_are_this_libms():
    """csrc/we_pow64_tables.h holds exactly the constants of the libm the reference would run on here."""
    import importlib.util

    libm_path = "/lib/x86_64-linux-gnu/libm.so.6"
    if not os.path.exists(libm_path):
        pytest.skip("no x86-64 glibc libm at the usual path")
    spec = importlib.util.spec_from_file_location("make_pow_tables", os.path.join(REPO, "tools", "make_pow_tables.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.render(libm_path) == open(mod.DST).read()


def test_angle_matches_oracle_and_numpy(hostlib):
    rng = np.random.default_rng(2)
    n = 50_000
    dim = np.array([31.2, 30.5, 29.9], dtype=np.float32)
    abc = rng.uniform(-5, 36, (n, 9)).astype(np.float32)
    abc[: n // 2, 3:6] = abc[: n // 2, 0:3] + rng.normal(0, 3, (n // 2, 3)).astype(np.float32)
    abc[: n // 2, 6:9] = abc[: n // 2, 0:3] + rng.normal(0, 3, (n // 2, 3)).astype(np.float32)
    out = np.empty(n, dtype=np.float32)
    hostlib.h_angle_cos(_p(abc), _p(dim), n, _p(out))
    L = wo.lib()
    ref = np.array([L.we_oracle_angle_cos(_p(abc[i, 0:3].copy()), _p(abc[i, 3:6].copy()), _p(abc[i, 6:9].copy()), _p(dim))
                    for i in range(n)], dtype=np.float32)
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))

    def np_angle(a, b, c, d):  # the reference's formula, maths/trig.py:112-121, on float32 arrays
        ba, bc, ac = np.abs(a - b), np.abs(c - b), np.abs(c - a)
        ba = np.where(ba > 0.5 * d, ba - d, ba)
        bc = np.where(bc > 0.5 * d, bc - d, bc)
        ac = np.where(ac > 0.5 * d, ac - d, ac)
        dba, dbc, dac = np.sqrt((ba**2).sum(axis=-1)), np.sqrt((bc**2).sum(axis=-1)), np.sqrt((ac**2).sum(axis=-1))
        return (dac**2 - dbc**2 - dba**2) / (-2 * dbc * dba)

    for i in range(3000):
        assert np.float32(np_angle(abc[i, 0:3], abc[i, 3:6], abc[i, 6:9], dim)).view(np.uint32) == out[i].view(np.uint32)


def test_distance_matches_oracle(hostlib):
    rng = np.random.default_rng(3)
    n = 20_000
    dims = np.array([31.169188, 31.017628, 29.992, 90, 90, 90], dtype=np.float32)
    pairs = rng.uniform(-10, 45, (n, 6)).astype(np.float32)
    out = np.empty(n)
    hostlib.h_dist_ortho(_p(pairs), _p(dims), n, _p(out))
    L = wo.lib()
    ref = np.array([L.we_oracle_distance(_p(pairs[i, 0:3].copy()), _p(pairs[i, 3:6].copy()), _p(dims)) for i in range(n)])
    assert np.array_equal(out, ref)


def test_arccos_rule_matches_numpy(hostlib):
    # float(degrees(arccos(c))) > 90 (analysis/HB.py:91,135) as a float32 threshold
    edge = np.array([0xB3800000, 0xB37FFFFF, 0xB3800001, 0xB3000000, 0xBF800000, 0xBF800001, 0x80000000, 0,
                     0x3F800000, 0xBF7FFFFF], dtype=np.uint32).view(np.float32)
    rng = np.random.default_rng(4)
    vals = np.concatenate([edge, rng.uniform(-1.2, 1.2, 100_000).astype(np.float32),
                           (rng.uniform(-4e-7, 4e-7, 20_000)).astype(np.float32)])
    with np.errstate(invalid="ignore"):
        want = np.array([float(np.degrees(np.arccos(c))) > 90 for c in vals])
    got = np.array([bool(hostlib.h_angle_gt_90(ctypes.c_float(float(c)))) for c in vals])
    assert np.array_equal(want, got)


@pytest.mark.parametrize("tag", gio.TAGS)
def test_topology_matches_oracle(tag):
    inp = gio.load_inputs(tag)
    ot = gio.oracle_topology(inp)
    st = SystemTopology(inp["masses"], inp["charges"], inp["resids"], list(inp["resnames"]),
                        list(inp["types"]), inp["bonds"], list(inp["names"]))
    assert sorted(set(st.solute_resids)) == sorted(set(ot.solute_resids))
    assert sorted(st.solute_ua.tolist()) == sorted(int(a) for a in ot.solute_UAs)
    assert st.centres.tolist() == [int(a) for a in ot.water_centres]
    assert st.heavy_idx.tolist() == ot.heavy_idx.tolist()
    for k, c in enumerate(st.centres):
        assert st.frag_atoms[st.frag_ptr[k]:st.frag_ptr[k + 1]].tolist() == ot.fragment_members[ot.fragment_id[c]]
    assert np.array_equal(st.is_water, ot.is_water)


def test_abi_exports_every_declared_symbol():
    so = _build.build()
    lib = ctypes.CDLL(so)
    header = open(os.path.join(REPO, "include", "waterentropy_b200.h")).read()
    declared = set(re.findall(r"\b(we_[a-z0-9_]+)\s*\(", header))
    declared -= {"we_ctx"}
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert set(_lib.EXPORTS) <= declared
    assert _lib.load().we_version() == 100


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from waterentropy_b200.engine import WaterEntropyEngine

    inp = gio.load_inputs("synth_sparse")
    st = SystemTopology(inp["masses"], inp["charges"], inp["resids"], list(inp["resnames"]),
                        list(inp["types"]), inp["bonds"])
    with pytest.raises(_lib.WaterEntropyError, match="no CUDA device"):
        WaterEntropyEngine(st)


def _labels_from_golden(d):
    lab = HBLabelCollection()
    for resid, a in d["resid_labelled_shell_counts"].items():
        for resname, b in a.items():
            for key, e in b.items():
                lab.add_counts(resid, resname, list(key), e["N_w"], e["shell_count"],
                               e.get("donates_to", {}), e.get("accepts_from", {}))
    return lab


@pytest.mark.parametrize("tag", gio.TAGS)
def test_consumers_match_reference(tag):
    """Orientations / Vibrations (consumers of the path's output) vs the reference's numbers."""
    ref = gio.load_reference(tag)
    inp = gio.load_inputs(tag)
    top = gio.oracle_topology(inp)
    for key, api in ref["api"].items():
        s, e, st = [int(v) for v in key.split(":")]
        fsi, cov, lab, n = wo.interfacial_run(top, inp["positions"], inp["forces"], inp["dimensions"], range(s, e, st))
        mine = HBLabelCollection()
        for resid, a in lab.resid_labelled_shell_counts.items():
            for resname, b in a.items():
                for k, en in b.items():
                    mine.add_counts(resid, resname, list(k), en["N_w"], en["shell_count"],
                                    dict(en.get("donates_to", {})), dict(en.get("accepts_from", {})))
        assert gio.plain(mine.resid_labelled_shell_counts) == gio.plain(lab.resid_labelled_shell_counts)
        assert gio.plain(mine.labelled_shell_counts) == gio.plain(lab.labelled_shell_counts)
        so = Orientations()
        so.add_data(mine)
        got = gio.plain(so.resid_labelled_Sorient)
        assert got.keys() == api["Sorient"].keys()
        for resid in got:
            for resname in got[resid]:
                assert got[resid][resname] == pytest.approx(api["Sorient"][resid][resname], rel=1e-12, abs=1e-12)
        cc = CovarianceCollection()
        for k in cov.counts:
            cc.forces[k], cc.torques[k], cc.counts[k] = cov.forces[k].copy(), cov.torques[k].copy(), cov.counts[k]
        vib = Vibrations(298)
        vib.add_data(cc, diagonalise=True)
        for k in cc.counts:
            assert np.allclose(cc.forces[k], api["cov_scaled"]["forces"][k], rtol=1e-11)
            assert np.allclose(vib.translational_S[k], api["vib"]["translational_S"][k], rtol=1e-10)
            assert np.allclose(vib.rotational_S[k], api["vib"]["rotational_S"][k], rtol=1e-10)
            assert np.allclose(vib.translational_freq[k], api["vib"]["translational_freq"][k], rtol=1e-10)


def test_reference_literal_sorient():
    # tests/recipes/test_interfacial_solvent.py:62-97 and test_bulk_water.py:34-45
    inp = gio.load_inputs("arginine")
    top = gio.oracle_topology(inp)
    fsi, cov, lab, n = wo.interfacial_run(top, inp["positions"], inp["forces"], inp["dimensions"], [0, 2])
    mine = HBLabelCollection()
    mine.merge(lab)
    so = Orientations()
    so.add_data(mine)
    S = so.resid_labelled_Sorient
    assert S[1]["ACE"] == pytest.approx([2.2473807716251804, 13, 6.467857142857143, 5.257142857142858,
                                         2.9594964080590276, 0.10765314493056646, 10.670272583734757, 8.183120170171872])
    assert S[3]["NME"] == pytest.approx([0.9503950365967891, 12, 6.538461538461538, 5.1923076923076925,
                                         2.050259026687599, 0.07010328362114078, 10.77132783200613, 8.187096360116465])
    cov2, lab2 = wo.bulk_run(top, inp["positions"], inp["forces"], inp["dimensions"], [0])
    mine2 = HBLabelCollection()
    mine2.merge(lab2)
    so2 = Orientations()
    so2.add_data(mine2)
    assert so2.resid_labelled_Sorient["WAT"]["WAT"] == pytest.approx(
        [11.295694179188464, 867, 6.053509907760105, 6.053509907760105, 6.053509907760105,
         0.2291068586055006, 9.775154915649155, 9.775154915649155])
    cc = CovarianceCollection()
    k = ("ACE_1", "WAT")
    cc.forces[k], cc.torques[k], cc.counts[k] = cov.forces[k].copy(), cov.torques[k].copy(), cov.counts[k]
    vib = Vibrations(298)
    vib.add_data(cc)
    assert np.allclose(vib.translational_S[k], [16.787066, 15.49228514, 18.34936798])
    assert np.allclose(vib.rotational_S[k], [5.13190029, 11.11787766, 7.50395398])


def test_shard_frames_and_merge():
    from waterentropy_b200 import distributed as wd
    from waterentropy_b200.engine import ENTRY_DTYPE

    idx = list(range(3, 50, 2))
    parts = [wd.shard_frames(idx, r, 4) for r in range(4)]
    assert sum(parts, []) == idx
    assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    a = np.zeros(3, dtype=ENTRY_DTYPE)
    b = np.zeros(2, dtype=ENTRY_DTYPE)
    a["hash"], a["hash2"], a["shell_count"] = [1, 2, 3], [9, 9, 9], [5, 6, 7]
    b["hash"], b["hash2"], b["shell_count"] = [2, 4], [9, 9], [10, 1]
    a["donates"][1, 0] = 3
    b["donates"][0, 0] = 4
    a["first_seen_inv"][1], b["first_seen_inv"][0] = 100, 200
    m = wd.merge_label_entries([a, b])
    got = {int(e["hash"]): e for e in m}
    assert sorted(got) == [1, 2, 3, 4]
    assert got[2]["shell_count"] == 16 and got[2]["donates"][0] == 7 and got[2]["first_seen_inv"] == 200


REF_FILES = "/root/reference/tests/input_files"


@pytest.mark.skipif(not os.path.isdir(REF_FILES), reason="reference fixtures only exist in the build container")
def test_streaming_readers_reproduce_golden_inputs():
    """io.Universe (prmtop + NetCDF, prmtop + TRR; memory-mapped, frames on demand) vs the committed inputs."""
    from waterentropy_b200.frames import iter_batches
    from waterentropy_b200.io import Universe

    for tag, top, crd in (("arginine", "amber/arginine_solution/system.prmtop", "amber/arginine_solution/system.nc"),
                          ("aspirin", "gromacs/aspirin_solution/system.prmtop", "gromacs/aspirin_solution/system.trr")):
        u = Universe(os.path.join(REF_FILES, top), os.path.join(REF_FILES, crd))
        g = gio.load_inputs(tag)
        n = g["positions"].shape[0]
        assert u.trajectory.n_frames == n
        P, F, D = u.frame_arrays(range(n))
        assert np.array_equal(P, g["positions"]) and np.array_equal(F, g["forces"]) and np.array_equal(D, g["dimensions"])
        ts = u.trajectory[n - 1]
        assert ts.frame == n - 1 and np.array_equal(u.atoms.positions, g["positions"][n - 1])
        got = [ids for _, _, _, ids in iter_batches(u, range(0, n, 2), 3)]
        assert sum(got, []) == list(range(0, n, 2))
        st = SystemTopology.from_universe(u)
        assert st.n_atoms == len(g["masses"]) and np.allclose(st.charges, g["charges"])
    with pytest.raises(NotImplementedError):
        Universe(os.path.join(REF_FILES, "gromacs/aspirin_solution/system.gro"),
                 os.path.join(REF_FILES, "gromacs/aspirin_solution/system.trr"))


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` runs without a GPU (it times the oracle port on the host cores) and
    prints one JSON line with the keys the driver reads."""
    import json
    import sys

    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--config", "C2", "--cpu-sample", "8", "--cpu-cores", "2"],
                         capture_output=True, text=True, timeout=300, cwd=REPO)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "water.frames/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == 2
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
