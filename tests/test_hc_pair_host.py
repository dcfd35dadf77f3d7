"""CPU: the pair-lane HalfCheetah algorithm (envpool_b200/csrc/mujoco_pair.cuh -- the source the
CUDA kernel compiles) built as plain C++ with two host threads playing the two lanes of an env,
checked against the oracle's restatement of the same pipeline (oracle/mjc_oracle.c, PARITY
UNPINNED against MuJoCo itself).  Covers what the lane split could get wrong --
which lane owns which body / capsule / limit, the duplicated root block staying bit-identical in
both lanes (the control flow of the solver relies on it), the pair sums, the shared-memory /
overflow row storage -- without a GPU.  Tolerance: summation order and libm only, 1e-9 per env
step of 5 mj_steps (measured 5e-13)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hc_pair_host", "pair_host.cc")
SO = os.path.join(HERE, "hc_pair_host", "libhc_pair_host.so")
HDRS = [os.path.join(HERE, "..", "envpool_b200", "csrc", f)
        for f in ("mujoco_pair.cuh", "mujoco_model.h")]


@pytest.fixture(scope="module")
def pair_host():
    deps = [SRC] + HDRS
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.run(["g++", "-std=c++20", "-O2", "-fPIC", "-shared", "-pthread", "-w", SRC,
                        "-o", SO], check=True)
    H = ctypes.CDLL(SO)
    vp = ctypes.c_void_p
    H.hc_pair_host_step.argtypes = [ctypes.c_char_p, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int]
    return H


def test_pair_lanes_match_oracle(pair_host):
    from envpool_b200 import _capi
    from oracle.oracle_lib import MjcSim, lib

    blob = _capi.hc_model_blob()
    L = lib()
    L.mjc_warm_mut.restype = ctypes.POINTER(ctypes.c_double)
    L.mjc_warm_mut.argtypes = [ctypes.c_void_p]
    sim = MjcSim()
    warm = np.ctypeslib.as_array(L.mjc_warm_mut(sim.d), shape=(9,))
    rng = np.random.default_rng(0)
    worst, rows_seen = 0.0, set()
    for trial in range(24):
        # near the floor, folded joints, fast: contacts on both legs, the torso and the head,
        # joint limits, warm starts from the previous step
        sim.qpos[:] = rng.uniform(-0.1, 0.1, 9)
        sim.qpos[1] = rng.uniform(-0.2, 0.3)
        sim.qpos[2] = rng.uniform(-1.5, 1.5)
        sim.qpos[3:] = rng.uniform(-1.2, 1.2, 6)
        sim.qvel[:] = rng.normal(0, 2.0, 9)
        warm[:] = 0
        for t in range(25):
            a = rng.uniform(-1.2, 1.2, 6)   # beyond ctrlrange: the clamp is on the path
            q, v, w = sim.qpos.copy(), sim.qvel.copy(), warm.copy()
            sim.step(a, 5)
            ks = (27, 9, 2)[t % 3]          # 2: most rows take the overflow path
            bad = pair_host.hc_pair_host_step(blob, q.ctypes.data, v.ctypes.data, w.ctypes.data,
                                              a.ctypes.data, 5, ks)
            assert bad == 0, "the duplicated root state of the two lanes is not bit-identical"
            rows_seen.add(sim.nefc)
            for g, r in ((q, sim.qpos), (v, sim.qvel), (w, warm)):
                err = float(np.max(np.abs(g - r) / (1 + np.abs(r))))
                worst = max(worst, err)
                assert err <= 1e-9, (trial, t, ks, sim.nefc, err)
    assert 0 in rows_seen and max(rows_seen) >= 8, rows_seen   # free flight and multi-contact
    print("pair-lane host build vs oracle: worst rel err", worst, "nefc seen", sorted(rows_seen))


def test_model_blob_matches_oracle_constants():
    """The engine's compiled model (epb_hc_model) against the oracle's independent compile."""
    from envpool_b200 import _capi
    from oracle.oracle_lib import MjcSim

    blob = np.frombuffer(_capi.hc_model_blob(), dtype=np.uint8)
    d = blob[: (blob.size // 8) * 8].view(np.float64)
    mass = d[0:7]
    np.testing.assert_allclose(mass.sum(), 14.0, rtol=1e-12)       # settotalmass, xml:52
    want = [6.2502, 1.5435, 1.5874, 1.0954, 1.4381, 1.2008, 0.8845]  # MuJoCo's body masses
    np.testing.assert_allclose(mass, want, atol=5e-5)
    c = MjcSim().constants()
    # HcModel starts with double mass[7] comx[7] comz[7] iyy[7] bposx[7] bposz[7] | armature,
    # damping, stiffness, rlo, rhi [9] each | gear[6] | 5 geom arrays [8] | dof_invweight0[9]
    # body_invw_tran[7] | grad timestep gravity mu meaninertia ...  (csrc/mujoco_model.h)
    np.testing.assert_allclose(mass, c["mass"], rtol=1e-13)
    np.testing.assert_allclose(d[7:14], c["com"][:, 0], rtol=1e-13, atol=1e-16)
    np.testing.assert_allclose(d[14:21], c["com"][:, 1], rtol=1e-13, atol=1e-16)
    np.testing.assert_allclose(d[21:28], c["iyy"], rtol=1e-13)
    np.testing.assert_allclose(d[133:142], c["dof_invweight0"], rtol=1e-11)
    np.testing.assert_allclose(d[142:149], c["body_invweight0"][:, 0], rtol=1e-11)
    np.testing.assert_allclose(d[153], c["meaninertia"], rtol=1e-13)
    assert d[149] == 0.046 and d[150] == 0.01 and d[151] == -9.81 and d[152] == 0.4


def test_row_storage_does_not_change_the_arithmetic(pair_host):
    """Rows in the shared-memory slab or in the thread-local overflow: same values, bit for bit
    (ks = 27: all rows in the slab; ks = 0 / 3: all / most rows in the overflow)."""
    from envpool_b200 import _capi

    blob = _capi.hc_model_blob()
    rng = np.random.default_rng(7)
    for _ in range(40):
        q0 = rng.uniform(-0.1, 0.1, 9)
        q0[1] = rng.uniform(-0.25, 0.2)
        q0[2] = rng.uniform(-1.5, 1.5)
        q0[3:] = rng.uniform(-1.3, 1.3, 6)
        v0, w0, a = rng.normal(0, 2.0, 9), rng.normal(0, 3.0, 9), rng.uniform(-1, 1, 6)
        outs = []
        for ks in (27, 3, 0):
            q, v, w = q0.copy(), v0.copy(), w0.copy()
            assert pair_host.hc_pair_host_step(blob, q.ctypes.data, v.ctypes.data, w.ctypes.data,
                                               a.ctypes.data, 5, ks) == 0
            outs.append(np.concatenate([q, v, w]))
        assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
