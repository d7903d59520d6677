"""The C++ drop-in class (host/cilqr_solver_compat.hpp) driven like motion_planning.cpp drives the
reference: construct from a config, solve per tick, apply x.row(1) — against the oracle."""
import os
import subprocess

import numpy as np
import pytest

import cilqr_b200 as cb
from oracle import oracle_py as op

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "toy-example-of-ilqr_b200", "host", "compat_demo")


def _dump(path, scn, N):
    cfg = dict(scn.cfg)
    cfg["lqr/N"] = N
    keys = [k for k, v in cfg.items() if not isinstance(v, list)]
    with open(path, "w") as f:
        f.write("%d\n" % len(keys))
        for k in keys:
            v = cfg[k]
            t = "b" if isinstance(v, bool) else "i" if isinstance(v, int) else "d" if isinstance(v, float) else "s"
            f.write("%s %s %s\n" % (k, t, int(v) if t == "b" else repr(v) if t == "d" else v))
        f.write("%d\n" % scn.ref.size())
        for a, b, c in zip(scn.ref.x, scn.ref.y, scn.ref.yaw):
            f.write("%r %r %r\n" % (float(a), float(b), float(c)))
        f.write("%d %d\n" % (scn.tracks.shape[0], scn.tracks.shape[1]))
        for tr in scn.tracks:
            for a, b, c in tr:
                f.write("%r %r %r\n" % (float(a), float(b), float(c)))
        f.write(" ".join(repr(float(v)) for v in list(scn.x0) + [scn.target_velocity] + list(scn.borders)) + "\n")


@pytest.mark.parametrize("name,ticks", [("two_borrow", 1), ("three_straight", 4)])
def test_cpp_dropin_matches_oracle(tmp_path, name, ticks):
    if not os.path.exists(DEMO):
        pytest.skip("compat_demo not built")
    scn = cb.get_scenario(name)
    N = 30
    path = str(tmp_path / "scenario.txt")
    _dump(path, scn, N)
    out = subprocess.run([DEMO, path, str(ticks)], capture_output=True, text=True, check=True).stdout.splitlines()
    o = op.Solver(scn.params, N)
    x0 = scn.x0.copy()
    pos = 0
    for tick in range(ticks):
        hdr = out[pos].split()
        u = np.array([[float(v) for v in l.split()[1:]] for l in out[pos + 1: pos + 1 + N]])
        x = np.array([[float(v) for v in l.split()[1:]] for l in out[pos + 1 + N: pos + 2 + 2 * N]])
        pos += 2 + 2 * N
        pb = cb.single_problem(scn, N, tick=tick, x0=x0)
        r = o.solve(pb.templates[0], pb.ref_velo[0], pb.n_obs[0], pb.obs[0], pb.borders[0], pb.x0[0])
        assert int(hdr[5]) == r.iters and int(hdr[3]) == r.status
        assert np.abs(u - r.u).max() < 1e-6 and np.abs(x - r.x).max() < 1e-6
        x0 = r.x[1].copy()
