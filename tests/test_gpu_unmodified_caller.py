"""SURVEY 8f-4 / INTEGRATION.md section 1: the reference's OWN caller — src/motion_planning.cpp, unmodified, built in
this repo's container by tests/unmodified_caller/Makefile against the drop-in CILQRSolver (Eigen-typed branch of
host/cilqr_solver_compat.hpp) — runs its receding-horizon loop on the GPU through libcilqr_b200.so.

The binary draws the ego state and the applied control of every tick on its (stubbed) figure; the test reads them back
and compares them with the oracle driven through the same loop.  The reference adds N(0, 0.02 m) noise to half of the
obstacle samples from std::random_device (src/motion_planning.cpp:163-171), so the comparison is to a few
centimetres, not to rounding; bit-level parity is the subject of the other tests."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import cilqr_b200 as cb
from oracle import oracle_py as op

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
BIN = os.path.join(ROOT, "tests", "_build", "motion_planning_b200")


def _write_config(path, name, sim_time):
    """The scenario as a YAML file in the reference's layout, with a shorter max_simulation_time (input data)."""
    from test_host_scenario_cpp import write_yaml
    cfg = dict(cb.templates.TEMPLATES[name])
    cfg["max_simulation_time"] = sim_time
    write_yaml(path, cfg)


@pytest.mark.parametrize("name", ["two_straight", "three_bend"])
def test_reference_main_runs_on_the_gpu_through_the_drop_in(tmp_path, name):
    if not os.path.exists(BIN):
        pytest.skip("tests/_build/motion_planning_b200 not built (needs the reference tree at build time)")
    ticks = 25
    cfg = tmp_path / "scenario.yaml"
    _write_config(str(cfg), name, 0.1 * ticks - 0.05)
    r = subprocess.run([BIN, "-c", str(cfg)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    # one block of TEXT lines per tick: x, y, v, yaw, acc, steer
    blocks = [b for b in r.stdout.split("TICK\n") if "TEXT" in b]
    assert len(blocks) == ticks, (len(blocks), r.stdout[-500:])
    got = np.array([[float(v) for v in re.findall(r"= (-?[0-9.]+)", b)] for b in blocks])
    assert got.shape == (ticks, 6)
    # the oracle through the same loop, without the noise
    scn = cb.get_scenario(name)
    N = scn.cfg["lqr/N"]
    o = op.Solver(scn.params, N)
    x0, t, exp = scn.x0.copy(), 0.0, []
    for _ in range(ticks):
        pb = cb.single_problem(scn, N, tick=int(t / scn.dt), x0=x0)
        res = o.solve(pb.templates[0], pb.ref_velo[0], pb.n_obs[0], pb.obs[0], pb.borders[0], pb.x0[0])
        x0 = res.x[1].copy()
        exp.append(list(x0) + list(res.u[0]))
        t += scn.dt
    exp = np.array(exp)
    err = np.abs(got - exp)
    print("%s: max |ego - oracle| over %d ticks: x %.3f y %.3f v %.3f yaw %.3f acc %.3f steer %.3f"
          % ((name, ticks) + tuple(err.max(axis=0))))
    assert err[:, 0].max() < 0.1 and err[:, 1].max() < 0.1 and err[:, 2].max() < 0.1 and err[:, 3].max() < 0.03
    # the first tick is noise-free for the ego and nearly so for the obstacles: two printed decimals
    assert np.all(err[0, :4] <= 0.011)
