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


# ---------------------------------------------------------------------------
# synthetic batches: a small descriptor per scenario template, expanded into problem arrays
# either here (numpy) or in place on the device (cilqr_b200_synth_generate) — the same numbers bit for bit:
# every floating-point step below is a single IEEE add / multiply / min / max / floor / fmod, and the device
# kernel performs the same steps in the same order with contraction off.
# ---------------------------------------------------------------------------
@dataclass
class LaneTable:
    """A centre line as sampled by ReferenceLine (every 0.1 m): positions, yaw, arc length of each sample, and the
    left normal (-sin yaw, cos yaw) of each sample.  Poses between samples are linear interpolations."""
    x: np.ndarray
    y: np.ndarray
    yaw: np.ndarray
    lon: np.ndarray
    nx: np.ndarray
    ny: np.ndarray


def lane_table(cl):
    return LaneTable(cl.x.copy(), cl.y.copy(), cl.yaw.copy(), cl.longitude.copy(), -np.sin(cl.yaw), np.cos(cl.yaw))


def lane_pose(tb, s):
    """(x, y, yaw, sample index) at arc length s: sample i = floor((s - lon[0]) * 10) clamped to [0, M-2],
    f = (s - lon[i]) * 10, value = v[i] + f * (v[i+1] - v[i])."""
    s = np.asarray(s, dtype=np.float64)
    i = np.clip(np.floor((s - tb.lon[0]) * 10.0).astype(np.int64), 0, len(tb.lon) - 2)
    f = (s - tb.lon[i]) * 10.0
    return (tb.x[i] + f * (tb.x[i + 1] - tb.x[i]), tb.y[i] + f * (tb.y[i + 1] - tb.y[i]),
            tb.yaw[i] + f * (tb.yaw[i + 1] - tb.yaw[i]), i)


@dataclass
class SynthObstacle:
    kind: int = 0          # 0: follows a lane table; 1: straight line at constant y
    # kind 0: starts at start_s + U(draw; -8, 8) on `lane`, speed max(speed + U(draw + 1; -1, 1), 0.5), drives
    # towards decreasing s with yaw + pi when `oncoming` (src/motion_planning.cpp:149-158)
    lane: int = 0
    oncoming: int = 0
    draw: int = 0
    start_s: float = 0.0
    speed: float = 0.0
    # kind 1: x = [ego x0 +] U(draw; x_lo, x_hi) + direction * v t, v = U(draw_v; v_lo, v_hi),
    # y = y0, or y1 when two_lanes and u01(draw_lane) >= 0.5; constant yaw
    x_lo: float = 0.0
    x_hi: float = 0.0
    v_lo: float = 0.0
    v_hi: float = 0.0
    y0: float = 0.0
    y1: float = 0.0
    yaw: float = 0.0
    direction: float = 1.0
    rel_to_ego: int = 0
    two_lanes: int = 0
    draw_lane: int = 0
    draw_v: int = 0


@dataclass
class SynthTemplate:
    ego_kind: int = 0      # 0: x0 = (U0(-5,5), U1(-.6,.6), U2(5,10), U3(-.05,.05)), ref_velo = U4(6,10)
    #                        1: lane frame: s = ego_s + U0(-5,5) on ego_lane, lateral U1(-.6,.6) along the sample's
    #                           normal, v = max(ego_v + U2(-2,2), 0.5), yaw = lane yaw + U3(-.05,.05),
    #                           ref_velo = target_velocity + U4(-2,2)
    ego_lane: int = 0
    ego_s: float = 0.0
    ego_v: float = 0.0
    target_velocity: float = 0.0
    borders: tuple = (0.0, 0.0)
    obstacles: list = field(default_factory=list)


@dataclass
class SynthSpec:
    name: str
    N: int
    templates: list        # TemplateData per scenario template (instance i uses template i % len(templates))
    synth: list            # SynthTemplate per scenario template
    lanes: list            # LaneTable list the descriptors index into

    @property
    def max_obs(self):
        return max(len(t.obstacles) for t in self.synth)


def _lane_obstacles(scn, lane0, draw0):
    out = []
    for j in range(len(scn.ic) - 1):
        veh = j + 1
        out.append(SynthObstacle(kind=0, lane=lane0 + scn.lane_of[veh], oncoming=int(scn.ic[veh][3] > math.pi / 2),
                                 draw=draw0 + 2 * j, start_s=float(scn.start_s[veh]), speed=float(scn.ic[veh][2])))
    return out


def synth_spec(config, N=None):
    """Descriptor of BASELINE config C1..C4."""
    if config in ("C1", "C2", "C4"):
        scn = get_scenario("two_borrow" if config == "C4" else "two_straight")
        lanes = [lane_table(cl) for cl in scn.center_lines]
        st = SynthTemplate(ego_kind=0, target_velocity=float(scn.target_velocity), borders=tuple(scn.borders))
        if config == "C1":
            N = 50 if N is None else N
            st.obstacles = _lane_obstacles(scn, 0, 8)
        elif config == "C2":
            N = 100 if N is None else N
            st.obstacles = [SynthObstacle(kind=1, draw=8 + 3 * j, draw_lane=9 + 3 * j, draw_v=10 + 3 * j, x_lo=10.0,
                                          x_hi=90.0, v_lo=0.0, v_hi=7.0, y0=0.0, y1=3.6, two_lanes=1, rel_to_ego=1)
                            for j in range(3)]
        else:
            N = 200 if N is None else N
            st.obstacles = _lane_obstacles(scn, 0, 8) + [
                SynthObstacle(kind=1, draw=40, draw_v=41, x_lo=60.0, x_hi=160.0, v_lo=2.0, v_hi=8.0, y0=3.6,
                              yaw=math.fmod(0.0 + math.pi, 2 * math.pi), direction=-1.0)]
        return SynthSpec(config, N, [template_data(scn)], [st], lanes)
    if config == "C3":
        N = 50 if N is None else N
        scns = [get_scenario(n) for n in T.TEMPLATE_ORDER]
        tds = [template_data(s) for s in scns]
        for td in tds:  # first solve only: warm start off for all
            td.params = dict(td.params, use_last_solution=0)
        lanes, synth = [], []
        for s in scns:
            lane0 = len(lanes)
            lanes += [lane_table(cl) for cl in s.center_lines]
            synth.append(SynthTemplate(ego_kind=1, ego_lane=lane0 + s.lane_of[0], ego_s=float(s.start_s[0]),
                                       ego_v=float(s.ic[0][2]), target_velocity=float(s.target_velocity),
                                       borders=tuple(s.borders), obstacles=_lane_obstacles(s, lane0, 8)))
        return SynthSpec(config, N, tds, synth, lanes)
    raise ValueError("unknown synthetic config %r" % (config,))


def generate_host(spec, B, seed=DEFAULT_SEED, first_id=0):
    """Expands a SynthSpec into the problem arrays of instances [first_id, first_id + B) with numpy."""
    ids = np.arange(first_id, first_id + B, dtype=np.uint64)
    N, nt, max_obs = spec.N, len(spec.synth), spec.max_obs
    tmpl = (ids % np.uint64(nt)).astype(np.int32)
    x0, ref_velo, borders = np.zeros((B, 4)), np.zeros(B), np.zeros((B, 2))
    n_obs = np.zeros(B, np.int32)
    obs = np.zeros((B, max_obs, N + 1, 3))
    for ti, st in enumerate(spec.synth):
        sel = np.nonzero(tmpl == ti)[0]
        if len(sel) == 0:
            continue
        sid = ids[sel]
        dt = spec.templates[ti].params["dt"]
        t = np.arange(N + 1, dtype=np.float64) * dt
        if st.ego_kind == 0:
            ex0 = np.stack([uniform(seed, sid, 0, -5.0, 5.0), uniform(seed, sid, 1, -0.6, 0.6),
                            uniform(seed, sid, 2, 5.0, 10.0), uniform(seed, sid, 3, -0.05, 0.05)], axis=1)
            rv = uniform(seed, sid, 4, 6.0, 10.0)
        else:
            tb = spec.lanes[st.ego_lane]
            s_ego = np.minimum(np.maximum(st.ego_s + uniform(seed, sid, 0, -5.0, 5.0), tb.lon[0]), tb.lon[-1])
            dl = uniform(seed, sid, 1, -0.6, 0.6)
            px, py, pyaw, i = lane_pose(tb, s_ego)
            ex0 = np.stack([px + dl * tb.nx[i], py + dl * tb.ny[i],
                            np.maximum(st.ego_v + uniform(seed, sid, 2, -2.0, 2.0), 0.5),
                            pyaw + uniform(seed, sid, 3, -0.05, 0.05)], axis=1)
            rv = st.target_velocity + uniform(seed, sid, 4, -2.0, 2.0)
        x0[sel], ref_velo[sel], borders[sel] = ex0, rv, np.asarray(st.borders)
        n_obs[sel] = len(st.obstacles)
        for j, ob in enumerate(st.obstacles):
            if ob.kind == 0:
                tb = spec.lanes[ob.lane]
                ds = uniform(seed, sid, ob.draw, -8.0, 8.0)
                v = np.maximum(ob.speed + uniform(seed, sid, ob.draw + 1, -1.0, 1.0), 0.5)
                s0 = (ob.start_s + ds)[:, None]
                tv = t[None, :] * v[:, None]
                if ob.oncoming:
                    sj = np.minimum(np.maximum(s0 - tv, tb.lon[0]), tb.lon[-1])
                else:
                    sj = np.maximum(np.minimum(s0 + tv, tb.lon[-1]), tb.lon[0])
                px, py, pyaw, _ = lane_pose(tb, sj)
                if ob.oncoming:
                    pyaw = np.fmod(pyaw + math.pi, 2 * math.pi)
                obs[sel, j, :, 0], obs[sel, j, :, 1], obs[sel, j, :, 2] = px, py, pyaw
            else:
                ox = uniform(seed, sid, ob.draw, ob.x_lo, ob.x_hi)
                if ob.rel_to_ego:
                    ox = ex0[:, 0] + ox
                v = uniform(seed, sid, ob.draw_v, ob.v_lo, ob.v_hi)
                y = np.full(len(sel), ob.y0)
                if ob.two_lanes:
                    y = np.where(u01(seed, sid, ob.draw_lane) < 0.5, ob.y0, ob.y1)
                obs[sel, j, :, 0] = ox[:, None] + ob.direction * (t[None, :] * v[:, None])
                obs[sel, j, :, 1] = y[:, None]
                obs[sel, j, :, 2] = ob.yaw
    return BatchProblem(spec.templates, N, x0, ref_velo, borders, tmpl, n_obs, obs, name=spec.name,
                        meta={"seed": seed, "first_id": first_id})


def synthetic_batch(config, B, N=None, seed=DEFAULT_SEED, first_id=0):
    """Synthetic batches of SURVEY 8d, generated on the host.  `first_id` lets a rank generate its own slice
    [first_id, first_id + B) of a larger batch.  (BatchSolver.generate does the same on the device.)"""
    return generate_host(synth_spec(config, N), B, seed, first_id)
