#!/bin/bash
# Round 2, first GPU call: (1) probe the GPU box for MuJoCo / gymnasium / envpool / dm_control
# (VERDICT r1 task 1), (2) first run of the staged device RNG corner-case tests.
mkdir -p gpurun_out/r2_probe
O=gpurun_out/r2_probe
{
  echo "== python packages =="
  for m in mujoco gymnasium gym envpool dm_control dm_env optree mujoco_py jax brax; do
    python - <<PY 2>&1 | tail -1
try:
    import $m
    print("$m", "PRESENT", getattr($m, "__version__", "?"), getattr($m, "__file__", "?"))
except Exception as e:
    print("$m", "ABSENT", type(e).__name__, str(e)[:80])
PY
  done
  echo "== pip download mujoco==3.6.0 (no network expected) =="
  timeout 60 python -m pip download --no-deps -d /tmp/mjdl mujoco==3.6.0 2>&1 | tail -3
  echo "== wheelhouse / filesystem search =="
  ls /opt/wheelhouse 2>/dev/null | grep -i -E "mujoco|gymnasium|dm_control|envpool|optree|dm_env" || echo "no matching wheel in /opt/wheelhouse"
  find / -xdev \( -iname "libmujoco*" -o -iname "mujoco*.whl" -o -iname "mujoco.h" \) 2>/dev/null | head -20
  echo "(end of find)"
  echo "== host =="
  nproc; lscpu | grep -E "Model name|Socket|Thread|Core" ; nvidia-smi -L
} > $O/probe.txt 2>&1
ENVPOOL_B200_RUN_STAGED=1 timeout 600 python -m pytest tests/test_gpu_rng_corner_cases.py -x -q > $O/pytest_rng_corner.txt 2>&1
echo "rng corner exit $?" >> $O/probe.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1
echo "pytest gpu exit $?" >> $O/probe.txt
timeout 300 python bench.py --steps 20 --warmup 3 > $O/bench20.txt 2>&1
tail -5 $O/probe.txt
