"""Pin the HalfCheetah restatement (oracle/mjc_oracle.c) against MuJoCo itself.

TEST INFRASTRUCTURE, and the one piece of it that could not be run in the build container or on
the GPU box of rounds 1-2: neither has `mujoco` (profiles/r2_probe_gpu_box_packages.txt).  It is
committed so that the pin is one command away wherever MuJoCo 3.6.0 (the reference's pinned
version, envpool/workspace0.bzl:561-572) exists:

    python oracle/pin_halfcheetah_with_mujoco.py record  <half_cheetah_envpool.xml>
    python oracle/pin_halfcheetah_with_mujoco.py compare

`record` (needs `import mujoco`) writes tests/golden/halfcheetah_mujoco.npz:
  * the compiled model constants the restatement derives by hand (body masses / CoMs / inertias,
    dof_invweight0, body_invweight0, armature / damping / stiffness / ranges, geom frames,
    meaninertia, the option block) -- what mj_loadXML makes of the XML the reference loads
    (mujoco/gym/mujoco_env.h:50-58,87);
  * teacher-forced single mj_steps: (qpos, qvel, qacc_warmstart, ctrl) -> (qpos, qvel, qacc,
    qacc_warmstart, nefc) from contact-rich random states, the protocol of the reference's own
    alignment test (mujoco/gym/mujoco_gym_align_test.py:120-171);
  * free runs: 64 env steps of 5 mj_steps from reset-like states with random actions
    (mujoco_gym_align_test.py:193-206).
`compare` (needs only numpy + the oracle) replays both on the restatement and reports the worst
errors against the bars VERDICT r1 set: 1e-9 teacher-forced, 5e-3 over 64 free steps; exit code
0 only if both hold and every model constant matches to 1e-9.  When that passes, rows a16-a18 of
SURVEY section 8 are pinned, the four open points of DESIGN.md section 3 are settled, and the
same file pins the CUDA kernels through tests/test_gpu_halfcheetah.py (they are held to the
restatement at 1e-9 per env step)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "halfcheetah_mujoco.npz")
N_TF, N_FREE, T_FREE, FRAME_SKIP = 512, 32, 64, 5


def contact_rich_states(rng, n):
    q = rng.uniform(-0.1, 0.1, size=(n, 9))
    q[:, 1] = rng.uniform(-0.25, 0.3, n)      # torso height offset: legs on / in the floor
    q[:, 2] = rng.uniform(-1.5, 1.5, n)       # pitch
    q[:, 3:] = rng.uniform(-1.3, 1.3, size=(n, 6))   # beyond the joint ranges: limits active
    v = rng.normal(0, 2.0, size=(n, 9))
    w = rng.normal(0, 5.0, size=(n, 9)) * (rng.random((n, 1)) < 0.5)   # half cold, half warm
    a = rng.uniform(-1.2, 1.2, size=(n, 6))   # beyond ctrlrange: the clamp is on the path
    return q, v, w, a


def record(xml_path):
    import mujoco

    m = mujoco.MjModel.from_xml_path(xml_path)
    d = mujoco.MjData(m)
    assert (m.nq, m.nv, m.nu) == (9, 9, 6), (m.nq, m.nv, m.nu)
    out = {"mujoco_version": np.array(mujoco.__version__), "xml": np.array(open(xml_path).read())}
    for name in ("body_mass", "body_ipos", "body_iquat", "body_inertia", "body_pos",
                 "body_invweight0",
                 "dof_invweight0", "dof_armature", "dof_damping", "jnt_stiffness", "jnt_range",
                 "geom_pos", "geom_quat", "geom_size", "geom_friction", "geom_solref",
                 "geom_solimp", "jnt_solref", "jnt_solimp", "actuator_gear", "actuator_ctrlrange"):
        out["model_" + name] = np.array(getattr(m, name))
    out["model_meaninertia"] = np.array(m.stat.meaninertia)
    out["model_opt"] = np.array([m.opt.timestep, m.opt.tolerance, m.opt.ls_tolerance,
                                 m.opt.iterations, m.opt.ls_iterations, m.opt.impratio,
                                 m.opt.solver, m.opt.cone, m.opt.integrator, m.opt.gravity[2]])
    rng = np.random.default_rng(2024)
    q, v, w, a = contact_rich_states(rng, N_TF)
    tf = {k: np.zeros((N_TF, 9)) for k in ("q1", "v1", "qacc", "w1")}
    nefc = np.zeros(N_TF, dtype=np.int64)
    for i in range(N_TF):
        mujoco.mj_resetData(m, d)
        d.qpos[:], d.qvel[:], d.qacc_warmstart[:], d.ctrl[:] = q[i], v[i], w[i], a[i]
        mujoco.mj_step(m, d)
        tf["q1"][i], tf["v1"][i], tf["qacc"][i], tf["w1"][i] = (d.qpos, d.qvel, d.qacc,
                                                                  d.qacc_warmstart)
        nefc[i] = d.nefc
    out.update(tf_q0=q, tf_v0=v, tf_w0=w, tf_ctrl=a, tf_nefc=nefc,
               **{"tf_" + k: x for k, x in tf.items()})
    q0 = rng.uniform(-0.1, 0.1, size=(N_FREE, 9))
    v0 = rng.normal(0, 0.1, size=(N_FREE, 9))
    acts = rng.uniform(-1, 1, size=(N_FREE, T_FREE, 6))
    traj = np.zeros((N_FREE, T_FREE, 18))
    for i in range(N_FREE):
        mujoco.mj_resetData(m, d)
        d.qpos[:], d.qvel[:] = q0[i], v0[i]
        mujoco.mj_forward(m, d)                       # MujocoReset: mujoco_env.h:128-130
        for t in range(T_FREE):
            d.ctrl[:] = acts[i, t]
            for _ in range(FRAME_SKIP):
                mujoco.mj_step(m, d)                  # MujocoStep: mujoco_env.h:137-148
            traj[i, t, :9], traj[i, t, 9:] = d.qpos, d.qvel
    out.update(free_q0=q0, free_v0=v0, free_actions=acts, free_traj=traj)
    np.savez_compressed(GOLDEN, **out)
    print("wrote", GOLDEN, "from MuJoCo", mujoco.__version__)


def compare():
    sys.path.insert(0, ROOT)
    import ctypes

    from oracle.oracle_lib import MjcSim, lib

    g = np.load(GOLDEN)
    sim = MjcSim()
    L = lib()
    L.mjc_warm_mut.restype = ctypes.POINTER(ctypes.c_double)
    L.mjc_warm_mut.argtypes = [ctypes.c_void_p]
    warm = np.ctypeslib.as_array(L.mjc_warm_mut(sim.d), shape=(9,))
    ok = True
    # ---- model constants (bodies 1..7 of MuJoCo = the restatement's 0..6; body 0 is the world)
    c = sim.constants()
    checks = {
        "body_mass": (c["mass"], g["model_body_mass"][1:]),
        "body CoM x": (c["com"][:, 0], g["model_body_ipos"][1:, 0]),
        "body CoM z": (c["com"][:, 1], g["model_body_ipos"][1:, 2]),
        "dof_invweight0": (c["dof_invweight0"], g["model_dof_invweight0"]),
        "body_invweight0": (c["body_invweight0"], g["model_body_invweight0"][1:]),
        "meaninertia": (np.array(c["meaninertia"]), g["model_meaninertia"]),
    }
    # planar inertia about y: MuJoCo stores PRINCIPAL inertias in the inertial frame body_iquat;
    # rotate back into the body frame (whose y axis is the world's for this planar model)
    iyy = []
    for b in range(1, 8):
        w_, x_, y_, z_ = g["model_body_iquat"][b]
        R = np.array([[1 - 2 * (y_ * y_ + z_ * z_), 2 * (x_ * y_ - z_ * w_), 2 * (x_ * z_ + y_ * w_)],
                      [2 * (x_ * y_ + z_ * w_), 1 - 2 * (x_ * x_ + z_ * z_), 2 * (y_ * z_ - x_ * w_)],
                      [2 * (x_ * z_ - y_ * w_), 2 * (y_ * z_ + x_ * w_), 1 - 2 * (x_ * x_ + y_ * y_)]])
        iyy.append((R @ np.diag(g["model_body_inertia"][b]) @ R.T)[1, 1])
    checks["body inertia (y)"] = (c["iyy"], np.array(iyy))
    for name, (have, want) in checks.items():
        err = float(np.max(np.abs(have - want) / (1e-300 + np.abs(want) + 1e-12)))
        print(f"  model {name:18s} worst rel err {err:.3e}")
        ok &= err <= 1e-9
    # ---- teacher-forced single mj_steps
    worst = 0.0
    for i in range(len(g["tf_q0"])):
        sim.qpos[:], sim.qvel[:], warm[:] = g["tf_q0"][i], g["tf_v0"][i], g["tf_w0"][i]
        sim.step(g["tf_ctrl"][i], 1)
        for have, want in ((sim.qpos, g["tf_q1"][i]), (sim.qvel, g["tf_v1"][i]),
                           (warm, g["tf_w1"][i])):
            worst = max(worst, float(np.max(np.abs(have - want) / (1 + np.abs(want)))))
    print(f"  teacher-forced mj_step: worst rel err {worst:.3e} over {len(g['tf_q0'])} states "
          f"(nefc up to {int(g['tf_nefc'].max())}); bar 1e-9")
    ok &= worst <= 1e-9
    # ---- free runs
    worst = 0.0
    for i in range(len(g["free_q0"])):
        sim.qpos[:], sim.qvel[:], warm[:] = g["free_q0"][i], g["free_v0"][i], 0.0
        for t in range(g["free_traj"].shape[1]):
            sim.step(g["free_actions"][i, t], FRAME_SKIP)
            have = np.concatenate([sim.qpos, sim.qvel])
            want = g["free_traj"][i, t]
            worst = max(worst, float(np.max(np.abs(have - want) / (1 + np.abs(want)))))
    print(f"  free runs, {g['free_traj'].shape[1]} env steps: worst rel err {worst:.3e}; "
          f"bar 5e-3 (mujoco_gym_align_test.py:42-43,93-94)")
    ok &= worst <= 5e-3
    print("PINNED" if ok else "NOT PINNED", "against MuJoCo", str(g["mujoco_version"]))
    return 0 if ok else 1


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "record":
        record(sys.argv[2])
    elif len(sys.argv) >= 2 and sys.argv[1] == "compare":
        sys.exit(compare())
    else:
        print(__doc__)
        sys.exit(2)
