"""CPU, only where /root/reference is mounted (this container; skipped on the GPU box): the
REFERENCE's own Python layer -- envpool/python/{api,env_spec,dm_envpool,gymnasium_envpool,
envpool,data}.py, envpool/registration.py and the family packages' __init__/registration,
imported unmodified -- runs on top of THIS repo's pybind11 modules.  That is the drop-in
boundary of SURVEY.md 8(b) seen from the reference's side.

optree / dm_env / gymnasium are not installed here; tests/refstubs holds stand-ins for the
handful of names that layer uses, and the check runs in a subprocess (tests/ref_layer_check.py)
so they never shadow anything in this process."""
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("ENVPOOL_REFERENCE_ROOT", "/root/reference")

pytestmark = pytest.mark.skipif(
    not os.path.isdir(os.path.join(REF, "envpool", "python")),
    reason="the reference checkout is not mounted here")


@pytest.fixture(scope="module")
def report(engine_built):
    out = subprocess.run([sys.executable, os.path.join(HERE, "ref_layer_check.py")],
                         capture_output=True, text=True, timeout=300)
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("REPORT ")]
    assert lines, out.stdout[-2000:] + out.stderr[-4000:]
    return json.loads(lines[-1][7:])


def test_reference_layer_was_the_one_imported(report):
    assert report["reference_file"].startswith(REF)
    mro = report["adapter_mro"]
    # the reference's adapter class sits directly on OUR pybind pool class
    assert mro[1] == "envpool_b200.classic_control.classic_control_envpool._CartPoleEnvPool"
    assert "envpool.python.dm_envpool.DMEnvPoolMixin" in mro
    assert "envpool.python.envpool.EnvPoolMixin" in mro


def test_reference_make_spec_over_our_modules_equals_ours(report):
    assert not report["errors"], {t: report["tasks"][t].get("error") for t in report["errors"]}
    assert len(report["tasks"]) >= 24
    for v in ("v3", "v4", "v5"):   # via the reference's mujoco/gym registration.py
        hc = report["tasks"][f"HalfCheetah-{v}"]
        assert hc["obs_space"]["shape"] == [17] and hc["dm_action"]["shape"] == [6]
        assert hc["reward_threshold"] == 4800.0
    for task, e in report["tasks"].items():
        assert e["in_reference_registry"], task
        for k in ("config_equal", "state_keys_equal", "action_keys_equal", "obs_space_equal",
                  "act_space_equal"):
            assert e[k], (task, k, e.get("config_diff"))
    cp = report["tasks"]["CartPole-v1"]
    assert cp["obs_space"]["shape"] == [4] and cp["act_space"]["n"] == 2
    assert cp["reward_threshold"] == 475.0
    assert cp["dm_obs_fields"] == ["env_id", "players", "obs"]
    assert report["tasks"]["Acrobot-v1"]["dm_obs_fields"] == ["env_id", "players", "obs", "state"]
    assert report["tasks"]["Pendulum-v1"]["act_space"]["type"] == "Box"


def test_reference_dm_fold_over_our_key_order(report):
    f = report["dm_fold"]
    assert f["obs_is_same_object"] and f["players_env_id"] == [0, 1, 2]
    assert f["last"] == [False, True, True]


def test_reference_pool_construction_reaches_our_engine(report):
    """envpool.make_gymnasium() of the REFERENCE constructs OUR pool: with a GPU it steps,
    without one the engine refuses loudly (there is no CPU fallback to fall into)."""
    p = report["pool"]
    if p["ok"]:
        assert p["obs_shape"] == [4, 4] and p["env_id"] == [0, 1, 2, 3]
        assert p["reward"] == [1.0, 1.0, 1.0, 1.0]
    else:
        assert "cuda" in p["error"].lower(), p["error"]
