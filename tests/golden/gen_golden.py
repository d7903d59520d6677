"""Generates tests/golden/*.npz from oracle/_ref — the reference's own sources compiled in place
(oracle/Makefile target `ref`).  Run here, in the build container, where /root/reference exists:

    python tests/golden/gen_golden.py

The fixtures are small (first solves of the four YAML templates at the shipped N = 30 and the
benchmark N = 50, a warm-started receding-horizon sequence, an ALM solve, and per-stage outputs
on a perturbed trajectory) and travel with the repo to the GPU box, where the reference tree
does not exist.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cilqr_b200 as cb  # noqa: E402
from oracle import ref_py as rp  # noqa: E402
from helpers import perturbed_trajectories  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def problem_arrays(pb, b=0):
    td = pb.templates[pb.tmpl[b]]
    return td, (td, pb.ref_velo[b], pb.n_obs[b], pb.obs[b], pb.borders[b])


def main():
    assert rp.available(), "build oracle/_ref first (make -C oracle ref)"
    # 1. first solves of the four templates
    for name in cb.templates.TEMPLATE_ORDER:
        scn = cb.get_scenario(name)
        for N in (30, 50):
            pb = cb.single_problem(scn, N)
            td, args = problem_arrays(pb)
            s = rp.RefSolver(dict(td.params, use_last_solution=0), N)
            u, x, st = s.solve(*args, pb.x0[0])
            np.savez_compressed(os.path.join(OUT, "solve_%s_N%d.npz" % (name, N)), u=u, x=x, status=st,
                                x0=pb.x0[0], ref_velo=pb.ref_velo[0], borders=pb.borders[0], obs=pb.obs[0],
                                wx=td.wx, wy=td.wy, wyaw=td.wyaw)
    # 2. warm-started sequence (three_straight ships use_last_solution: true)
    scn = cb.get_scenario("three_straight")
    N = 30
    s = rp.RefSolver(scn.params, N)
    x0 = scn.x0.copy()
    us, xs, sts, x0s = [], [], [], []
    for tick in range(6):
        pb = cb.single_problem(scn, N, tick=tick, x0=x0)
        td, args = problem_arrays(pb)
        u, x, st = s.solve(*args, pb.x0[0])
        us.append(u); xs.append(x); sts.append(st); x0s.append(x0.copy())
        x0 = x[1].copy()
    np.savez_compressed(os.path.join(OUT, "warm_three_straight_N30.npz"), u=np.array(us), x=np.array(xs),
                        status=np.array(sts), x0=np.array(x0s))
    # 3. ALM solve (slove_type: "alm", same scalars as two_straight / two_borrow)
    for name in ("two_straight", "two_borrow"):
        scn = cb.get_scenario(name)
        pb = cb.single_problem(scn, 30)
        td, args = problem_arrays(pb)
        s = rp.RefSolver(dict(td.params, solve_type=1, use_last_solution=0), 30)
        u, x, st = s.solve(*args, pb.x0[0])
        np.savez_compressed(os.path.join(OUT, "alm_%s_N30.npz" % name), u=u, x=x, status=st)
    # 4. per-stage outputs on perturbed trajectories of synthetic instances (both vehicle models)
    pb = cb.synthetic_batch("C3", 8, N=50)
    u, x = perturbed_trajectories(pb, seed=5)
    rec = dict(u=u, x=x)
    keys = ("J", "lx", "lu", "lxx", "luu", "A", "B", "d", "K", "dV", "status", "ref_pts", "fwd_u", "fwd_x")
    acc = {k: [] for k in keys}
    for b in range(pb.B):
        td, args = problem_arrays(pb, b)
        s = rp.RefSolver(td.params, pb.N)
        acc["J"].append(s.total_cost(*args, u[b], x[b]))
        r = s.backward_pass(*args, u[b], x[b], 0.5)
        for k in ("lx", "lu", "lxx", "luu", "d", "K", "dV", "status"):
            acc[k].append(r[k])
        acc["A"].append(r["A"].reshape(pb.N, 4, 4))
        acc["B"].append(r["B"].reshape(pb.N, 4, 2))
        acc["ref_pts"].append(s.ref_points(td, x[b]))
        nu, nx = s.forward_pass(u[b], x[b], r["d"], r["K"], 0.25)
        acc["fwd_u"].append(nu)
        acc["fwd_x"].append(nx)
    rec.update({k: np.array(v) for k, v in acc.items()})
    np.savez_compressed(os.path.join(OUT, "stages_C3_B8_N50.npz"), **rec)
    # 5. reference-line sampling of the reference's own spline code
    lines = {}
    for name in cb.templates.TEMPLATE_ORDER:
        cfg = cb.templates.TEMPLATES[name]
        wx, wy, wyaw, lon = rp.reference_line(cfg["laneline/reference/x"], cfg["laneline/reference/y"],
                                              cfg["laneline/center_line"][0])
        lines[name + "_x"], lines[name + "_y"], lines[name + "_yaw"] = wx, wy, wyaw
    np.savez_compressed(os.path.join(OUT, "reference_lines.npz"), **lines)
    print("wrote", sorted(f for f in os.listdir(OUT) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
