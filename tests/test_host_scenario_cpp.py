"""C++ host layer (host/scenario_host.hpp + headless_planner): the YAML-subset GlobalConfig, spline,
reference line and obstacle-track builder against the numpy scenario prep (which is itself pinned to the
reference's spline code by tests/golden/reference_lines.npz); and, on a GPU, the headless replay of the
reference's simulator loop against the oracle loop."""
import os
import subprocess

import numpy as np
import pytest

import cilqr_b200 as cb
from oracle import oracle_py as op

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "toy-example-of-ilqr_b200", "host", "headless_planner")


def write_yaml(path, cfg):
    """A scenario file in the reference's YAML layout (config/scenario_*.yaml) from a flat template."""
    def sec(prefix):
        return {k[len(prefix) + 1:]: v for k, v in cfg.items() if k.startswith(prefix + "/")}
    def fmt(v):
        if isinstance(v, bool):
            return "true" if v else "false"
        if isinstance(v, str):
            return '"%s"' % v
        if isinstance(v, list):
            return "[" + ", ".join(repr(float(e)) for e in v) + "]"
        return repr(v)
    with open(path, "w") as f:
        f.write("max_simulation_time: %r\ndelta_t: %r\n\n" % (cfg["max_simulation_time"], cfg["delta_t"]))
        for s in ("lqr", "iteration", "vehicle"):
            f.write("%s:\n" % s)
            for k, v in sec(s).items():
                f.write("  %s: %s   # %s\n" % (k, fmt(v), k))
            f.write("\n")
        f.write("laneline:\n  reference:\n    x: %s\n    y: %s\n" % (fmt(cfg["laneline/reference/x"]), fmt(cfg["laneline/reference/y"])))
        f.write("  border: %s\n  center_line: %s\n\n" % (fmt(cfg["laneline/border"]), fmt(cfg["laneline/center_line"])))
        f.write("initial_condition:\n  # [x, y, v, yaw]\n")
        for row in cfg["initial_condition"]:
            f.write("  - %s  # vehicle\n" % fmt(row))
        f.write("\nvisualization:\n  y_lim: [-5, 13]\n  show_obstacle_boundary: true\n")


def _need_bin():
    if not os.path.exists(BIN):
        pytest.skip("headless_planner not built")


@pytest.mark.parametrize("name", cb.templates.TEMPLATE_ORDER)
def test_cpp_scenario_prep_matches_numpy(tmp_path, name):
    _need_bin()
    cfg = cb.templates.TEMPLATES[name]
    path = str(tmp_path / "scenario.yaml")
    write_yaml(path, cfg)
    out = subprocess.run([BIN, "-c", path, "-d"], capture_output=True, text=True, check=True).stdout.split("\n")
    scn = cb.get_scenario(name)
    M = int(out[0].split()[1])
    assert M == scn.ref.size()
    ref = np.array([[float(v) for v in l.split()] for l in out[1:1 + M]])
    assert np.abs(ref[:, 0] - scn.ref.x).max() < 1e-9 and np.abs(ref[:, 1] - scn.ref.y).max() < 1e-9
    assert np.abs(ref[:, 2] - scn.ref.yaw).max() < 1e-9
    b = [float(v) for v in out[1 + M].split()[1:]]
    assert b == list(scn.borders)
    nveh, T = [int(v) for v in out[2 + M].split()[1:]]
    assert nveh == len(scn.ic) and T == scn.tracks.shape[1]
    tr = np.array([[float(v) for v in l.split()] for l in out[3 + M:3 + M + nveh * T]]).reshape(nveh, T, 3)
    assert np.abs(tr[1:] - scn.tracks).max() < 1e-9
    last = out[3 + M + nveh * T].split()
    assert int(last[1]) == cfg["lqr/N"] and last[3] == cfg["vehicle/reference_point"] and last[5] == cfg["lqr/slove_type"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["two_borrow", "three_straight"])
def test_headless_replay_matches_oracle_loop(tmp_path, name):
    _need_bin()
    cfg = cb.templates.TEMPLATES[name]
    path = str(tmp_path / "scenario.yaml")
    write_yaml(path, cfg)
    ticks = 6
    rows = subprocess.run([BIN, "-c", path, "-t", str(ticks)], capture_output=True, text=True, check=True).stdout.strip().split("\n")
    assert rows[0].startswith("t,x,y,v,yaw")
    got = np.array([[float(v) for v in r.split(",")] for r in rows[1:]])
    assert got.shape[0] == ticks
    scn = cb.get_scenario(name)
    N = cfg["lqr/N"]
    o = op.Solver(scn.params, N)
    x0, t = scn.x0.copy(), 0.0
    for i in range(ticks):
        index = int(t / scn.dt)  # motion_planning.cpp:181
        pb = cb.single_problem(scn, N, tick=index, x0=x0)
        r = o.solve(pb.templates[0], pb.ref_velo[0], pb.n_obs[0], pb.obs[0], pb.borders[0], pb.x0[0])
        assert np.abs(got[i, 1:5] - x0).max() < 1e-5, i
        if int(got[i, 7]) != r.iters:
            break  # a decision flipped: later ticks start from different states
        assert np.abs(got[i, 5:7] - r.u[0]).max() < 1e-6
        x0 = r.x[1].copy()
        t += scn.dt
    assert i >= 2
