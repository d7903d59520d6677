"""Host-side scenario construction: the step *before* the hot path.

Restates, Eigen-free and in numpy, what the reference does between reading a YAML and
calling CILQRSolver::solve:
  * natural cubic spline + ReferenceLine sampling every 0.1 m with a lateral offset
    (src/cubic_spline.cpp:17-169, src/utils.cpp:21-35, :60-67),
  * road borders [max(border), min(border)] (src/motion_planning.cpp:101-103),
  * constant-speed obstacle tracks along the nearest centre line, oncoming if yaw0 > pi/2
    (src/motion_planning.cpp:121-160; the random noise of :163-171 is off),
and generates the synthetic batches C1..C4 of SURVEY §8d from a counter-based RNG keyed
(seed, instance, draw), so any slice of a batch can be regenerated independently on any rank.
"""
import math
from dataclasses import dataclass, field

import numpy as np

from . import templates as T

# ---------------------------------------------------------------------------
# solver scalars
# ---------------------------------------------------------------------------
PARAM_FIELDS = [
    ("dt", "d"),
    ("w_pos", "d"), ("w_vel", "d"), ("w_yaw", "d"), ("w_acc", "d"), ("w_stl", "d"),
    ("obstacle_exp_q1", "d"), ("obstacle_exp_q2", "d"), ("state_exp_q1", "d"), ("state_exp_q2", "d"),
    ("alm_rho_init", "d"), ("alm_gamma", "d"), ("max_rho", "d"), ("max_mu", "d"),
    ("init_lamb", "d"), ("lamb_decay", "d"), ("lamb_amplify", "d"), ("max_lamb", "d"),
    ("convergence_threshold", "d"), ("accept_step_threshold", "d"),
    ("wheelbase", "d"), ("width", "d"), ("length", "d"),
    ("velo_max", "d"), ("velo_min", "d"), ("yaw_lim", "d"), ("acc_max", "d"), ("acc_min", "d"),
    ("stl_lim", "d"), ("d_safe", "d"),
    ("max_iter", "i"), ("solve_type", "i"), ("reference_point", "i"), ("use_last_solution", "i"),
]


def params_from_config(cfg):
    """Flat GlobalConfig map -> the scalars the reference constructor reads (cpp:17-83)."""
    st = cfg["lqr/slove_type"]
    solve_type = 1 if st == "alm" else 0  # anything else falls back to barrier (cpp:38-41)
    return {
        "dt": cfg["delta_t"],
        "w_pos": cfg["lqr/w_pos"], "w_vel": cfg["lqr/w_vel"], "w_yaw": cfg["lqr/w_yaw"],
        "w_acc": cfg["lqr/w_acc"], "w_stl": cfg["lqr/w_stl"],
        "obstacle_exp_q1": cfg["lqr/obstacle_exp_q1"], "obstacle_exp_q2": cfg["lqr/obstacle_exp_q2"],
        "state_exp_q1": cfg["lqr/state_exp_q1"], "state_exp_q2": cfg["lqr/state_exp_q2"],
        "alm_rho_init": cfg["lqr/alm_rho_init"], "alm_gamma": cfg["lqr/alm_gamma"],
        "max_rho": cfg["lqr/max_rho"], "max_mu": cfg["lqr/max_mu"],
        "init_lamb": cfg["iteration/init_lamb"], "lamb_decay": cfg["iteration/lamb_decay"],
        "lamb_amplify": cfg["iteration/lamb_amplify"], "max_lamb": cfg["iteration/max_lamb"],
        "convergence_threshold": cfg["iteration/convergence_threshold"],
        "accept_step_threshold": cfg["iteration/accept_step_threshold"],
        "wheelbase": cfg["vehicle/wheelbase"], "width": cfg["vehicle/width"], "length": cfg["vehicle/length"],
        "velo_max": cfg["vehicle/velo_max"], "velo_min": cfg["vehicle/velo_min"], "yaw_lim": cfg["vehicle/yaw_lim"],
        "acc_max": cfg["vehicle/acc_max"], "acc_min": cfg["vehicle/acc_min"], "stl_lim": cfg["vehicle/stl_lim"],
        "d_safe": cfg["vehicle/d_safe"],
        "max_iter": int(cfg["iteration/max_iter"]),
        "solve_type": solve_type,
        "reference_point": 0 if cfg["vehicle/reference_point"] == "rear_center" else 1,
        "use_last_solution": 1 if cfg["lqr/use_last_solution"] else 0,
    }


# ---------------------------------------------------------------------------
# spline / reference line
# ---------------------------------------------------------------------------
class CubicSpline1D:
    """Natural cubic spline (src/cubic_spline.cpp:17-124)."""

    def __init__(self, x, y):
        self.x = np.asarray(x, dtype=np.float64)
        self.a = np.asarray(y, dtype=np.float64)
        n = len(self.x)
        h = np.diff(self.x)
        if np.any(h < 0):
            raise ValueError("x coordinates must be sorted in ascending order")
        A = np.zeros((n, n))
        A[0, 0] = 1.0
        for i in range(n - 1):
            if i != n - 2:
                A[i + 1, i + 1] = 2.0 * (h[i] + h[i + 1])
            A[i + 1, i] = h[i]
            A[i, i + 1] = h[i]
        A[0, 1] = 0.0
        A[n - 1, n - 2] = 0.0
        A[n - 1, n - 1] = 1.0
        Bv = np.zeros(n)
        for i in range(n - 2):
            Bv[i + 1] = 3.0 * (self.a[i + 2] - self.a[i + 1]) / h[i + 1] - 3.0 * (self.a[i + 1] - self.a[i]) / h[i]
        self.c = np.linalg.solve(A, Bv)
        self.d = (self.c[1:] - self.c[:-1]) / (3.0 * h)
        self.b = (self.a[1:] - self.a[:-1]) / h - h * (self.c[1:] + 2 * self.c[:-1]) / 3.0
        self.h = h

    def _seg(self, s):
        s = np.asarray(s, dtype=np.float64)
        if np.any(s < self.x[0]) or np.any(s > self.x[-1]):
            raise ValueError("received value out of the pre-defined range")
        # upper_bound - 1; the reference indexes one past the coefficient arrays when s equals
        # the last knot exactly (undefined behaviour) — clamp to the last segment instead.
        idx = np.searchsorted(self.x, s, side="right") - 1
        idx = np.minimum(idx, len(self.x) - 2)
        return idx, s - self.x[idx]

    def position(self, s):
        i, dx = self._seg(s)
        return self.a[i] + self.b[i] * dx + self.c[i] * dx ** 2 + self.d[i] * dx ** 3

    def d1(self, s):
        i, dx = self._seg(s)
        return self.b[i] + 2.0 * self.c[i] * dx + 3.0 * self.d[i] * dx ** 2

    def d2(self, s):
        i, dx = self._seg(s)
        return 2.0 * self.c[i] + 6.0 * self.d[i] * dx


class CubicSpline2D:
    """src/cubic_spline.cpp:126-169."""

    def __init__(self, x, y):
        x = np.asarray(x, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        ds = np.hypot(np.diff(x), np.diff(y))
        self.s = np.concatenate([[0.0], np.cumsum(ds)])
        self.sx = CubicSpline1D(self.s, x)
        self.sy = CubicSpline1D(self.s, y)

    def position(self, s):
        return self.sx.position(s), self.sy.position(s)

    def yaw(self, s):
        return np.arctan2(self.sy.d1(s), self.sx.d1(s))


class ReferenceLine:
    """src/utils.cpp:21-35: the spline sampled every `accuracy` metres, shifted laterally by `width`."""

    def __init__(self, x, y, width=0.0, accuracy=0.1):
        self.spline = CubicSpline2D(x, y)
        self.delta_d = width
        ss = []
        s = 0.0
        end = self.spline.s[-1]
        while s <= end:  # s accumulates by repeated addition, as in the reference loop
            ss.append(s)
            s += accuracy
        self.longitude = np.array(ss)
        px, py = self.spline.position(self.longitude)
        yaw = self.spline.yaw(self.longitude)
        self.x = px - width * np.sin(yaw)
        self.y = py + width * np.cos(yaw)
        self.yaw = yaw

    def size(self):
        return len(self.x)

    def length(self):
        return float(self.spline.s[-1])

    def calc_position(self, s):
        """src/utils.cpp:60-67, vectorised: (x, y, yaw) at arc length s."""
        px, py = self.spline.position(s)
        yaw = self.spline.yaw(s)
        return px - self.delta_d * np.sin(yaw), py + self.delta_d * np.cos(yaw), yaw


# ---------------------------------------------------------------------------
# one scenario (a YAML template), as motion_planning.cpp prepares it
# ---------------------------------------------------------------------------
class Scenario:
    def __init__(self, cfg, extra_ticks=None):
        self.cfg = cfg
        self.params = params_from_config(cfg)
        rx, ry = cfg["laneline/reference/x"], cfg["laneline/reference/y"]
        self.center_lines = [ReferenceLine(rx, ry, w) for w in cfg["laneline/center_line"]]
        bw = sorted(cfg["laneline/border"], reverse=True)
        self.borders = np.array([bw[0], bw[-1]])  # motion_planning.cpp:101-103
        self.ref = self.center_lines[0]           # solve() is given center_lines[0] (:195)
        self.target_velocity = cfg["vehicle/target_velocity"]
        self.dt = cfg["delta_t"]
        self.ic = np.array(cfg["initial_condition"], dtype=np.float64)
        self.x0 = self.ic[0].copy()
        self.lane_of, self.start_s = self._assign_lanes()
        # tick times accumulate like `for (t = 0; t < T + 10; t += dt)` (:143)
        ts = []
        t = 0.0
        tmax = cfg["max_simulation_time"] + 10 if extra_ticks is None else extra_ticks * self.dt
        while t < tmax:
            ts.append(t)
            t += self.dt
        self.ticks = np.array(ts)
        self.tracks = np.stack([self.track(i, self.ticks) for i in range(1, len(self.ic))]) if len(self.ic) > 1 \
            else np.zeros((0, len(ts), 3))

    def _assign_lanes(self):
        """Nearest centre line and arc length per vehicle (motion_planning.cpp:121-141)."""
        lane_of, start_s = [], []
        for ic in self.ic:
            line_num, s0, min_diff = 0, self.center_lines[0].length(), -1.0
            for l, cl in enumerate(self.center_lines):
                d = np.hypot(cl.x - ic[0], cl.y - ic[1])
                inc = np.nonzero(d[1:] > d[:-1])[0]
                if len(inc):
                    i = inc[0] + 1
                    last = d[i - 1]
                    if min_diff < 0 or last < min_diff:
                        min_diff, line_num, s0 = last, l, cl.longitude[i - 1]
            lane_of.append(line_num)
            start_s.append(s0)
        return lane_of, start_s

    def track(self, veh, t, ds0=0.0, speed=None):
        """(x, y, yaw)[len(t)] of vehicle `veh` (motion_planning.cpp:143-160), optional jitter."""
        cl = self.center_lines[self.lane_of[veh]]
        v = self.ic[veh][2] if speed is None else speed
        s_end = cl.longitude[-1]
        if self.ic[veh][3] <= math.pi / 2:
            s = np.minimum(self.start_s[veh] + ds0 + t * v, s_end)
            s = np.maximum(s, cl.longitude[0])
            x, y, yaw = cl.calc_position(s)
        else:
            s = np.maximum(self.start_s[veh] + ds0 - t * v, cl.longitude[0])
            s = np.minimum(s, s_end)
            x, y, yaw = cl.calc_position(s)
            yaw = np.fmod(yaw + math.pi, 2 * math.pi)
        return np.stack([x, y, yaw], axis=-1)

    def obstacles_at(self, tick, N):
        """get_sub_routing_lines (src/utils.cpp:88-103): tracks from `tick` on, first N+1 samples."""
        sub = self.tracks[:, tick:, :]
        if sub.shape[1] < N + 1:
            raise IndexError("Index out of range")  # RoutingLine::operator[] (src/utils.cpp:53-55)
        return sub[:, : N + 1, :]


# ---------------------------------------------------------------------------
# batches
# ---------------------------------------------------------------------------
@dataclass
class TemplateData:
    params: dict
    wx: np.ndarray
    wy: np.ndarray
    wyaw: np.ndarray


@dataclass
class BatchProblem:
    """B independent solve() calls in the host layout of include/cilqr_b200.h."""
    templates: list
    N: int
    x0: np.ndarray        # [B][4]
    ref_velo: np.ndarray  # [B]
    borders: np.ndarray   # [B][2]
    tmpl: np.ndarray      # [B] int32
    n_obs: np.ndarray     # [B] int32
    obs: np.ndarray       # [B][max_obs][obs_len][3]
    name: str = ""
    meta: dict = field(default_factory=dict)

    @property
    def B(self):
        return self.x0.shape[0]

    @property
    def max_obs(self):
        return self.obs.shape[1]

    @property
    def obs_len(self):
        return self.obs.shape[2]

    def slice(self, lo, hi):
        return BatchProblem(self.templates, self.N, self.x0[lo:hi], self.ref_velo[lo:hi], self.borders[lo:hi],
                            self.tmpl[lo:hi], self.n_obs[lo:hi], self.obs[lo:hi], self.name, self.meta)


def template_data(scn):
    return TemplateData(scn.params, scn.ref.x.copy(), scn.ref.y.copy(), scn.ref.yaw.copy())


def single_problem(scn, N, tick=0, x0=None):
    """The solve() call motion_planning.cpp makes at simulation tick `tick`, as a batch of one."""
    obs = scn.obstacles_at(tick, N)
    return BatchProblem([template_data(scn)], N,
                        np.asarray(scn.x0 if x0 is None else x0, dtype=np.float64)[None, :].copy(),
                        np.array([scn.target_velocity], dtype=np.float64), scn.borders[None, :].copy(),
                        np.zeros(1, np.int32), np.array([obs.shape[0]], np.int32),
                        np.ascontiguousarray(obs[None]), name="single")


_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix64(z):
    z = z.astype(np.uint64)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def u01(seed, inst, draw):
    """Counter-based uniform [0,1): splitmix64 finaliser of (seed, instance id, draw index)."""
    inst = np.asarray(inst, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + np.uint64(0x9E3779B97F4A7C15) * (inst * np.uint64(64) + np.uint64(draw) + np.uint64(1))
        z = _mix64(_mix64(z))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def uniform(seed, inst, draw, lo, hi):
    return lo + (hi - lo) * u01(seed, inst, draw)


DEFAULT_SEED = 20261017
_SCN_CACHE = {}


def get_scenario(name):
    if name not in _SCN_CACHE:
        _SCN_CACHE[name] = Scenario(T.TEMPLATES[name])
    return _SCN_CACHE[name]


def _jittered_tracks(scn, ids, N, seed, draw0, extra=()):
    """Obstacle tracks of template `scn` for instances `ids`: every template obstacle shifted by
    ds ~ U(-8, 8) m along its lane with speed v + U(-1, 1) clamped >= 0.5."""
    t = np.arange(N + 1, dtype=np.float64) * scn.dt
    n = len(scn.ic) - 1
    obs = np.zeros((len(ids), n + len(extra), N + 1, 3))
    for j in range(n):
        ds = uniform(seed, ids, draw0 + 2 * j, -8.0, 8.0)
        v = np.maximum(scn.ic[j + 1][2] + uniform(seed, ids, draw0 + 2 * j + 1, -1.0, 1.0), 0.5)
        obs[:, j] = scn.track(j + 1, t[None, :], ds0=ds[:, None], speed=v[:, None])
    return obs


def synthetic_batch(config, B, N=None, seed=DEFAULT_SEED, first_id=0):
    """Synthetic batches of SURVEY §8d.  `first_id` lets a rank generate its own slice
    [first_id, first_id + B) of a larger batch."""
    ids = np.arange(first_id, first_id + B, dtype=np.uint64)
    if config == "C1":
        N = 50 if N is None else N
        scn = get_scenario("two_straight")
        tds = [template_data(scn)]
        x0 = np.stack([uniform(seed, ids, 0, -5.0, 5.0), uniform(seed, ids, 1, -0.6, 0.6),
                       uniform(seed, ids, 2, 5.0, 10.0), uniform(seed, ids, 3, -0.05, 0.05)], axis=1)
        ref_velo = uniform(seed, ids, 4, 6.0, 10.0)
        obs = _jittered_tracks(scn, ids, N, seed, 8)
        tmpl = np.zeros(B, np.int32)
        n_obs = np.full(B, obs.shape[1], np.int32)
        borders = np.tile(scn.borders, (B, 1))
    elif config == "C2":
        N = 100 if N is None else N
        scn = get_scenario("two_straight")
        tds = [template_data(scn)]
        x0 = np.stack([uniform(seed, ids, 0, -5.0, 5.0), uniform(seed, ids, 1, -0.6, 0.6),
                       uniform(seed, ids, 2, 5.0, 10.0), uniform(seed, ids, 3, -0.05, 0.05)], axis=1)
        ref_velo = uniform(seed, ids, 4, 6.0, 10.0)
        t = np.arange(N + 1, dtype=np.float64) * scn.dt
        obs = np.zeros((B, 3, N + 1, 3))
        for j in range(3):
            ox = x0[:, 0] + uniform(seed, ids, 8 + 3 * j, 10.0, 90.0)
            lane = np.where(u01(seed, ids, 9 + 3 * j) < 0.5, 0.0, 3.6)
            v = uniform(seed, ids, 10 + 3 * j, 0.0, 7.0)
            obs[:, j, :, 0] = ox[:, None] + v[:, None] * t[None, :]
            obs[:, j, :, 1] = lane[:, None]
        tmpl = np.zeros(B, np.int32)
        n_obs = np.full(B, 3, np.int32)
        borders = np.tile(scn.borders, (B, 1))
    elif config == "C3":
        N = 50 if N is None else N
        scns = [get_scenario(n) for n in T.TEMPLATE_ORDER]
        tds = [template_data(s) for s in scns]
        for td in tds:  # first solve only: warm start off for all
            td.params = dict(td.params, use_last_solution=0)
        max_obs = max(len(s.ic) - 1 for s in scns)
        tmpl = (ids % np.uint64(4)).astype(np.int32)
        x0 = np.zeros((B, 4))
        ref_velo = np.zeros(B)
        borders = np.zeros((B, 2))
        n_obs = np.zeros(B, np.int32)
        obs = np.zeros((B, max_obs, N + 1, 3))
        for ti, s in enumerate(scns):
            sel = np.nonzero(tmpl == ti)[0]
            if len(sel) == 0:
                continue
            sid = ids[sel]
            # ego jitter as C1, applied in the lane frame of the template's initial condition
            ds = uniform(seed, sid, 0, -5.0, 5.0)
            dl = uniform(seed, sid, 1, -0.6, 0.6)
            cl = s.center_lines[s.lane_of[0]]
            s_ego = np.clip(s.start_s[0] + ds, cl.longitude[0], cl.longitude[-1])
            ex, ey, eyaw = cl.calc_position(s_ego)
            x0[sel, 0] = ex - dl * np.sin(eyaw)
            x0[sel, 1] = ey + dl * np.cos(eyaw)
            x0[sel, 2] = np.maximum(s.ic[0][2] + uniform(seed, sid, 2, -2.0, 2.0), 0.5)
            x0[sel, 3] = eyaw + uniform(seed, sid, 3, -0.05, 0.05)
            ref_velo[sel] = s.target_velocity + uniform(seed, sid, 4, -2.0, 2.0)
            borders[sel] = s.borders
            o = _jittered_tracks(s, sid, N, seed, 8)
            obs[sel, : o.shape[1]] = o
            n_obs[sel] = o.shape[1]
    elif config == "C4":
        N = 200 if N is None else N
        scn = get_scenario("two_borrow")
        tds = [template_data(scn)]
        x0 = np.stack([uniform(seed, ids, 0, -5.0, 5.0), uniform(seed, ids, 1, -0.6, 0.6),
                       uniform(seed, ids, 2, 5.0, 10.0), uniform(seed, ids, 3, -0.05, 0.05)], axis=1)
        ref_velo = uniform(seed, ids, 4, 6.0, 10.0)
        base = _jittered_tracks(scn, ids, N, seed, 8)
        t = np.arange(N + 1, dtype=np.float64) * scn.dt
        extra = np.zeros((B, 1, N + 1, 3))
        ox = uniform(seed, ids, 40, 60.0, 160.0)
        v = uniform(seed, ids, 41, 2.0, 8.0)
        extra[:, 0, :, 0] = ox[:, None] - v[:, None] * t[None, :]
        extra[:, 0, :, 1] = 3.6
        extra[:, 0, :, 2] = math.fmod(0.0 + math.pi, 2 * math.pi)
        obs = np.concatenate([base, extra], axis=1)
        tmpl = np.zeros(B, np.int32)
        n_obs = np.full(B, obs.shape[1], np.int32)
        borders = np.tile(scn.borders, (B, 1))
    else:
        raise ValueError("unknown synthetic config %r" % (config,))
    return BatchProblem(tds, N, np.ascontiguousarray(x0), np.ascontiguousarray(ref_velo),
                        np.ascontiguousarray(borders), tmpl, n_obs, np.ascontiguousarray(obs), name=config,
                        meta={"seed": seed, "first_id": first_id})
