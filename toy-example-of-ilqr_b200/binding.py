"""ctypes binding of libcilqr_b200.so (include/cilqr_b200.h) and the host-side mirror of the
reference's solver interface.

  BatchSolver   — one handle on one GPU: upload / solve_resident / download, the fused
                  host-buffer solve(), and the per-stage operators the parity tests use.
  CILQRSolver   — same constructor/solve() surface as the reference class
                  (include/cilqr_solver.hpp:31-42): CILQRSolver(config) and
                  solve(x0, ref_waypoints, ref_velo, obs_preds, road_borders) -> (u [N][2], x [N+1][4]),
                  a batch of one on top of BatchSolver.

There is no CPU fallback: if the shared library is missing or no sm_100 device is present the
constructors raise.
"""
import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from .scenario import PARAM_FIELDS, BatchProblem, params_from_config

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcilqr_b200.so")

MAX_TEMPLATES = 8
STATUS_NAMES = ["RUNNING", "CONVERGED", "BACKWARD_PASS_FAIL", "FORWARD_PASS_FAIL", "FORWARD_PASS_SMALL_STEP"]
EXIT_NAMES = ["MAX_ITER", "CONVERGED", "MAX_LAMB"]

# every symbol include/cilqr_b200.h declares (tests/test_abi.py checks the header against this)
EXPORTS = [
    "cilqr_b200_last_error", "cilqr_b200_version", "cilqr_b200_create", "cilqr_b200_destroy",
    "cilqr_b200_set_stream", "cilqr_b200_set_template", "cilqr_b200_reset", "cilqr_b200_solve_batch",
    "cilqr_b200_upload", "cilqr_b200_solve_resident", "cilqr_b200_download", "cilqr_b200_counters",
    "cilqr_b200_set_option", "cilqr_b200_enable_trace", "cilqr_b200_get_trace", "cilqr_b200_simulate",
    "cilqr_b200_stage_times",
    "cilqr_b200_stage_init", "cilqr_b200_stage_ref_match", "cilqr_b200_stage_cost", "cilqr_b200_stage_derivs",
    "cilqr_b200_stage_backward", "cilqr_b200_stage_forward", "cilqr_b200_bench_backward",
    "cilqr_b200_bench_tile_records",
    "cilqr_b200_synth_set_lanes", "cilqr_b200_synth_generate", "cilqr_b200_synth_download",
]


class CilqrParams(C.Structure):
    _fields_ = [(n, C.c_double if t == "d" else C.c_int32) for n, t in PARAM_FIELDS]

    @classmethod
    def from_dict(cls, d):
        p = cls()
        for n, t in PARAM_FIELDS:
            setattr(p, n, float(d[n]) if t == "d" else int(d[n]))
        return p


SYNTH_MAX_OBS = 16


class SynthObstacleC(C.Structure):
    _fields_ = [("kind", C.c_int32), ("lane", C.c_int32), ("oncoming", C.c_int32), ("draw", C.c_int32),
                ("start_s", C.c_double), ("speed", C.c_double),
                ("x_lo", C.c_double), ("x_hi", C.c_double), ("v_lo", C.c_double), ("v_hi", C.c_double),
                ("y0", C.c_double), ("y1", C.c_double), ("yaw", C.c_double), ("direction", C.c_double),
                ("rel_to_ego", C.c_int32), ("two_lanes", C.c_int32), ("draw_lane", C.c_int32), ("draw_v", C.c_int32)]


class SynthTemplateC(C.Structure):
    _fields_ = [("ego_kind", C.c_int32), ("ego_lane", C.c_int32), ("n_obs", C.c_int32), ("reserved", C.c_int32),
                ("ego_s", C.c_double), ("ego_v", C.c_double), ("target_velocity", C.c_double),
                ("borders", C.c_double * 2), ("dt", C.c_double), ("obs", SynthObstacleC * SYNTH_MAX_OBS)]


class Counters(C.Structure):
    _fields_ = [("total_iters", C.c_int64), ("total_trials", C.c_int64), ("rounds", C.c_int32),
                ("launches", C.c_int32), ("exits", C.c_int32 * 3), ("reserved", C.c_int32)]


class CilqrError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("cilqr_b200 error %d: %s" % (code, msg))
        self.code = code


LIB_PARITY_PATH = os.path.join(_HERE, "libcilqr_b200_parity.so")
_lib = None
_lib_parity = None


def load_library(flavour="fast"):
    """Loads libcilqr_b200.so; raises (never falls back) when it has not been built.
    flavour "parity": libcilqr_b200_parity.so, the same kernels compiled with -DCILQR_PARITY -fmad=false
    (reference operation order, portable transcendentals) for bit-for-bit comparisons with the CPU."""
    global _lib, _lib_parity
    if flavour == "parity":
        if _lib_parity is None:
            if not os.path.exists(LIB_PARITY_PATH):
                raise ImportError("%s is missing: run `python __graft_entry__.py` first" % LIB_PARITY_PATH)
            _lib_parity = C.CDLL(LIB_PARITY_PATH)
            _lib_parity.cilqr_b200_last_error.restype = C.c_char_p
            _lib_parity.cilqr_b200_version.restype = C.c_char_p
        return _lib_parity
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: run `python __graft_entry__.py` (nvcc, sm_100a) first; "
                              "cilqr_b200 has no CPU fallback" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.cilqr_b200_last_error.restype = C.c_char_p
        _lib.cilqr_b200_version.restype = C.c_char_p
    return _lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int32))


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


@dataclass
class SolveResult:
    u: np.ndarray
    x: np.ndarray
    J: np.ndarray          # [B][2] initial / final cost
    K: np.ndarray
    d: np.ndarray
    step_cost: np.ndarray
    status: np.ndarray
    iters: np.ndarray
    exit_reason: np.ndarray


class BatchSolver:
    def __init__(self, templates, max_batch, N, max_obs, dtype="f64", device=0, flavour="fast"):
        self.lib = load_library(flavour)
        self.N, self.max_batch, self.max_obs = int(N), int(max_batch), int(max_obs)
        self.dtype = {"f64": 0, "f32": 1}[dtype]
        self.h = C.c_void_p()
        p0 = CilqrParams.from_dict(templates[0].params)
        self._ck(self.lib.cilqr_b200_create(C.byref(p0), int(device), self.max_batch, self.N, self.max_obs,
                                            self.dtype, C.byref(self.h)))
        for t, td in enumerate(templates):
            self.set_template(t, td.params, td.wx, td.wy, td.wyaw)

    # -- plumbing ----------------------------------------------------------
    def _ck(self, rc):
        if rc != 0:
            raise CilqrError(rc, self.lib.cilqr_b200_last_error().decode())

    def close(self):
        if self.h:
            self.lib.cilqr_b200_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream):
        self._ck(self.lib.cilqr_b200_set_stream(self.h, C.c_void_p(cuda_stream)))

    def set_template(self, t, params=None, wx=None, wy=None, wyaw=None):
        p = CilqrParams.from_dict(params) if params is not None else None
        wx, wy, wyaw = _f64(wx), _f64(wy), _f64(wyaw)
        self._ck(self.lib.cilqr_b200_set_template(self.h, int(t), C.byref(p) if p is not None else None,
                                                  _dp(wx), _dp(wy), _dp(wyaw), 0 if wx is None else len(wx)))

    def reset(self):
        self._ck(self.lib.cilqr_b200_reset(self.h))

    def counters(self):
        c = Counters()
        self._ck(self.lib.cilqr_b200_counters(self.h, C.byref(c)))
        return {"total_iters": int(c.total_iters), "total_trials": int(c.total_trials), "rounds": int(c.rounds),
                "launches": int(c.launches), "exits": dict(zip(EXIT_NAMES, list(c.exits)))}

    OPT_WIDE_SEARCH, OPT_RUN_AHEAD, OPT_PREFETCH_BELOW, OPT_BENCH_PREFETCH, OPT_PROFILE_STAGES, OPT_PIPELINE = 0, 1, 2, 3, 4, 5
    OPT_STAGED_BACKWARD = 6
    OPT_REPACK = 7
    OPT_WIDE_STEP = 8
    OPT_LOOKAHEAD = 9
    OPT_FUSED_BACKWARD = 10
    STAGES = ["derivs", "backward", "forward", "ref_match", "cost", "decide"]

    def stage_times(self):
        ms = np.zeros(6)
        n = np.zeros(6, np.int32)
        self._ck(self.lib.cilqr_b200_stage_times(self.h, _dp(ms), _ip(n)))
        return {k: (float(a), int(b)) for k, a, b in zip(self.STAGES, ms, n)}

    def set_option(self, option, value):
        self._ck(self.lib.cilqr_b200_set_option(self.h, int(option), int(value)))

    def enable_trace(self, cap):
        self._trace_cap = int(cap)
        self._ck(self.lib.cilqr_b200_enable_trace(self.h, int(cap)))

    def get_trace(self, B):
        """(status, alpha, cost), each [B][cap]: outcome of every iter_step of the last solve."""
        cap = self._trace_cap
        st, al, co = np.empty((B, cap), np.int32), np.empty((B, cap), np.int32), np.empty((B, cap))
        self._ck(self.lib.cilqr_b200_get_trace(self.h, int(B), _ip(st), _ip(al), _dp(co)))
        return st, al, co

    # -- solve ---------------------------------------------------------------
    def _alloc_out(self, B, want_gains=True, pinned=None):
        N = self.N
        out = SolveResult(
            u=np.empty((B, N, 2)), x=np.empty((B, N + 1, 4)), J=np.empty((B, 2)),
            K=np.empty((B, N, 2, 4)) if want_gains else None, d=np.empty((B, N, 2)) if want_gains else None,
            step_cost=np.empty((B, N + 1)), status=np.empty(B, np.int32), iters=np.empty(B, np.int32),
            exit_reason=np.empty(B, np.int32))
        return out

    def upload(self, pb: BatchProblem):
        self._pb_keep = (_f64(pb.x0), _f64(pb.ref_velo), _f64(pb.borders), _i32(pb.tmpl), _i32(pb.n_obs), _f64(pb.obs))
        x0, rv, bd, tm, no, ob = self._pb_keep
        self._ck(self.lib.cilqr_b200_upload(self.h, pb.B, _dp(x0), _dp(rv), _dp(bd), _ip(tm), _ip(no), _dp(ob),
                                            int(pb.obs_len)))

    # -- synthetic workloads generated on the device ------------------
    def generate(self, spec, B, seed=None, first_id=0, keep_yaw=False):
        """cilqr_b200_synth_generate: instances [first_id, first_id + B) of the SynthSpec (scenario.synth_spec),
        the device twin of scenario.generate_host."""
        from .scenario import DEFAULT_SEED
        seed = DEFAULT_SEED if seed is None else seed
        if getattr(self, "_synth_spec", None) is not spec:
            off = np.zeros(len(spec.lanes) + 1, np.int32)
            for i, tb in enumerate(spec.lanes):
                off[i + 1] = off[i] + len(tb.x)
            cat = [_f64(np.concatenate([getattr(tb, f) for tb in spec.lanes])) for f in ("x", "y", "yaw", "lon", "nx", "ny")]
            self._ck(self.lib.cilqr_b200_synth_set_lanes(self.h, len(spec.lanes), _ip(off), *[_dp(a) for a in cat]))
            arr = (SynthTemplateC * len(spec.synth))()
            for t, st in enumerate(spec.synth):
                c = arr[t]
                c.ego_kind, c.ego_lane, c.n_obs = int(st.ego_kind), int(st.ego_lane), len(st.obstacles)
                c.ego_s, c.ego_v, c.target_velocity = float(st.ego_s), float(st.ego_v), float(st.target_velocity)
                c.borders[0], c.borders[1] = float(st.borders[0]), float(st.borders[1])
                c.dt = float(spec.templates[t].params["dt"])
                for j, ob in enumerate(st.obstacles):
                    for name, ctype in SynthObstacleC._fields_:
                        v = getattr(ob, name)
                        setattr(c.obs[j], name, int(v) if ctype is C.c_int32 else float(v))
            self._synth_spec, self._synth_arr = spec, arr
        self._ck(self.lib.cilqr_b200_synth_generate(self.h, int(B), C.c_uint64(int(first_id)), C.c_uint64(int(seed)),
                                                    len(spec.synth), self._synth_arr, int(bool(keep_yaw))))

    def synth_download(self, B):
        """The resident problem data: x0 [B][4], ref_velo [B], borders [B][2], tmpl [B], n_obs [B] and the obstacle
        samples as stored on the device, [B][max_obs][N+1][4]."""
        x0, rv, bd = np.empty((B, 4)), np.empty(B), np.empty((B, 2))
        tm, no = np.empty(B, np.int32), np.empty(B, np.int32)
        ob = np.empty((B, self.max_obs, self.N + 1, 4))
        self._ck(self.lib.cilqr_b200_synth_download(self.h, int(B), _dp(x0), _dp(rv), _dp(bd), _ip(tm), _ip(no), _dp(ob)))
        return x0, rv, bd, tm, no, ob

    def solve_resident(self, B):
        self._ck(self.lib.cilqr_b200_solve_resident(self.h, int(B)))

    def download(self, B, out=None, want_gains=True):
        out = out or self._alloc_out(B, want_gains)
        self._ck(self.lib.cilqr_b200_download(self.h, int(B), _dp(out.u), _dp(out.x), _dp(out.J), _dp(out.K),
                                              _dp(out.d), _dp(out.step_cost), _ip(out.status), _ip(out.iters),
                                              _ip(out.exit_reason)))
        return out

    def download_counts(self, B):
        """iter_step total and exit histogram of the first B resident instances (only the two int arrays cross PCIe)."""
        it, ex = np.empty(B, np.int32), np.empty(B, np.int32)
        self._ck(self.lib.cilqr_b200_download(self.h, int(B), None, None, None, None, None, None, None, _ip(it), _ip(ex)))
        return {"iters": int(it.sum(dtype=np.int64)), "exits": dict(zip(EXIT_NAMES, np.bincount(ex, minlength=3)[:3].tolist()))}

    def solve(self, pb: BatchProblem, out=None, want_gains=True):
        """cilqr_b200_solve_batch: host buffers in, host buffers out, copies included."""
        B = pb.B
        out = out or self._alloc_out(B, want_gains)
        x0, rv, bd, tm, no, ob = (_f64(pb.x0), _f64(pb.ref_velo), _f64(pb.borders), _i32(pb.tmpl), _i32(pb.n_obs),
                                  _f64(pb.obs))
        self._ck(self.lib.cilqr_b200_solve_batch(
            self.h, B, _dp(x0), _dp(rv), _dp(bd), _ip(tm), _ip(no), _dp(ob), int(pb.obs_len), _dp(out.u), _dp(out.x),
            _dp(out.J), _dp(out.K), _dp(out.d), _dp(out.step_cost), _ip(out.status), _ip(out.iters),
            _ip(out.exit_reason)))
        return out

    def simulate(self, x0, ref_velo, borders, tmpl, n_obs, tracks, ticks):
        """Closed receding-horizon loop on the device (motion_planning.cpp:180-197): returns the ego state
        at every tick [B][ticks+1][4], iter_step counts and final status per tick [B][ticks]."""
        x0, ref_velo, borders, tracks = _f64(x0), _f64(ref_velo), _f64(borders), _f64(tracks)
        tmpl, n_obs = _i32(tmpl), _i32(n_obs)
        B = x0.shape[0]
        ego = np.empty((B, ticks + 1, 4))
        iters, status = np.empty((B, ticks), np.int32), np.empty((B, ticks), np.int32)
        self._ck(self.lib.cilqr_b200_simulate(self.h, B, _dp(x0), _dp(ref_velo), _dp(borders), _ip(tmpl), _ip(n_obs),
                                              _dp(tracks), int(tracks.shape[2]), int(ticks), _dp(ego), _ip(iters),
                                              _ip(status)))
        return ego, iters, status

    # -- stage operators -----------------------------------------------------
    def stage_init(self, x0, tmpl=None, warm=False, last_u=None):
        x0 = _f64(x0)
        B = x0.shape[0]
        u, x = np.empty((B, self.N, 2)), np.empty((B, self.N + 1, 4))
        lu = _f64(last_u)
        self._ck(self.lib.cilqr_b200_stage_init(self.h, B, _dp(x0), _ip(_i32(tmpl)), int(bool(warm)), _dp(lu), _dp(u), _dp(x)))
        return u, x

    def stage_ref_match(self, x, tmpl=None):
        x = _f64(x)
        B = x.shape[0]
        idx = np.empty((B, self.N + 1), np.int32)
        self._ck(self.lib.cilqr_b200_stage_ref_match(self.h, B, _dp(x), _ip(_i32(tmpl)), _ip(idx)))
        return idx

    def stage_cost(self, pb: BatchProblem, u, x, alm_mu=None, alm_rho=None):
        B = pb.B
        J, sc = np.empty(B), np.empty((B, self.N + 1))
        a = (_f64(u), _f64(x), _f64(pb.ref_velo), _f64(pb.borders), _i32(pb.tmpl), _i32(pb.n_obs), _f64(pb.obs),
             _f64(alm_mu), _f64(alm_rho))
        self._ck(self.lib.cilqr_b200_stage_cost(self.h, B, _dp(a[0]), _dp(a[1]), _dp(a[2]), _dp(a[3]), _ip(a[4]),
                                                _ip(a[5]), _dp(a[6]), int(pb.obs_len), _dp(a[7]), _dp(a[8]), _dp(J),
                                                _dp(sc)))
        return J, sc

    def stage_derivs(self, pb: BatchProblem, u, x, alm_mu=None, alm_rho=None):
        B, N = pb.B, self.N
        lx, lu = np.empty((B, N + 1, 4)), np.empty((B, N, 2))
        lxx, luu = np.empty((B, N + 1, 4, 4)), np.empty((B, N, 2, 2))
        A, Bm = np.empty((B, N, 4, 4)), np.empty((B, N, 4, 2))
        mun = np.empty((B, N, 8 + 2 * self.max_obs)) if alm_mu is not None else None
        a = (_f64(u), _f64(x), _f64(pb.ref_velo), _f64(pb.borders), _i32(pb.tmpl), _i32(pb.n_obs), _f64(pb.obs),
             _f64(alm_mu), _f64(alm_rho))
        self._ck(self.lib.cilqr_b200_stage_derivs(self.h, B, _dp(a[0]), _dp(a[1]), _dp(a[2]), _dp(a[3]), _ip(a[4]),
                                                  _ip(a[5]), _dp(a[6]), int(pb.obs_len), _dp(a[7]), _dp(a[8]),
                                                  _dp(lx), _dp(lu), _dp(lxx), _dp(luu), _dp(A), _dp(Bm), _dp(mun)))
        return dict(lx=lx, lu=lu, lxx=lxx, luu=luu, A=A, B=Bm, mu_next=mun)

    def stage_backward(self, lx, lu, lxx, luu, A, Bm, lamb):
        a = [_f64(v) for v in (lx, lu, lxx, luu, A, Bm)]
        B, N = a[0].shape[0], self.N
        lamb = _f64(np.broadcast_to(np.asarray(lamb, dtype=np.float64), (B,)))
        d, K, dV, st = np.empty((B, N, 2)), np.empty((B, N, 2, 4)), np.empty((B, 2)), np.empty(B, np.int32)
        self._ck(self.lib.cilqr_b200_stage_backward(self.h, B, *[_dp(v) for v in a], _dp(lamb), _dp(d), _dp(K),
                                                    _dp(dV), _ip(st)))
        return d, K, dV, st

    def stage_forward(self, u, x, d, K, alpha, tmpl=None):
        a = [_f64(v) for v in (u, x, d, K)]
        B, N = a[0].shape[0], self.N
        alpha = _f64(np.broadcast_to(np.asarray(alpha, dtype=np.float64), (B,)))
        nu, nx = np.empty((B, N, 2)), np.empty((B, N + 1, 4))
        self._ck(self.lib.cilqr_b200_stage_forward(self.h, B, *[_dp(v) for v in a], _dp(alpha), _ip(_i32(tmpl)),
                                                   _dp(nu), _dp(nx)))
        return nu, nx

    # -- roofline leg ----------------------------------------------------------
    def bench_backward(self, B, lamb=0.0, reps=20, flush_l2=True):
        ms = np.empty(reps, np.float32)
        nbytes = C.c_double()
        self._ck(self.lib.cilqr_b200_bench_backward(self.h, int(B), C.c_double(lamb), int(reps), int(bool(flush_l2)),
                                                    ms.ctypes.data_as(C.POINTER(C.c_float)), C.byref(nbytes)))
        return ms, nbytes.value

    def bench_tile_records(self, B0, B):
        self._ck(self.lib.cilqr_b200_bench_tile_records(self.h, int(B0), int(B)))


class CILQRSolver:
    """Mirror of the reference class (include/cilqr_solver.hpp:31-42) over the C ABI.

    `config` is the flat "section/key" map GlobalConfig serves (src/global_config.cpp:22-92);
    `ref_waypoints` needs .x/.y/.yaw (ReferenceLine, include/utils.hpp:44-46); `obs_preds` is a
    list of tracks, each an array [T][3] of (x, y, yaw) per tick (RoutingLine).
    A track shorter than N+1 raises IndexError like RoutingLine::operator[] throws
    std::out_of_range (src/utils.cpp:53-55).
    """

    def __init__(self, config, dtype="f64", device=0, max_obs=16):
        self.params = params_from_config(config)
        self.N = int(config["lqr/N"])
        if int(config.get("lqr/nx", 4)) != 4 or int(config.get("lqr/nu", 2)) != 2:
            raise ValueError("the bicycle model has nx = 4, nu = 2")
        from .scenario import TemplateData
        self._td = TemplateData(self.params, np.zeros(1), np.zeros(1), np.zeros(1))
        self._solver = BatchSolver([self._td], 1, self.N, max_obs, dtype, device)
        self._max_obs = max_obs
        self._wp_id = None
        self.last = None

    def solve(self, x0, ref_waypoints, ref_velo, obs_preds, road_borders):
        # the reference reads the line afresh on every solve(): re-upload whenever its content changed
        line = tuple(np.array(v, dtype=np.float64) for v in (ref_waypoints.x, ref_waypoints.y, ref_waypoints.yaw))
        if self._wp_id is None or not all(np.array_equal(a, b) for a, b in zip(line, self._wp_id)):
            self._solver.set_template(0, None, *line)
            self._wp_id = line
        n = len(obs_preds)
        if n > self._max_obs:
            raise ValueError("more obstacles (%d) than max_obs=%d" % (n, self._max_obs))
        obs = np.zeros((1, max(self._max_obs, 1), self.N + 1, 3))
        for j, tr in enumerate(obs_preds):
            tr = np.asarray(tr, dtype=np.float64)
            if tr.shape[0] < self.N + 1:
                raise IndexError("Index out of range")
            obs[0, j] = tr[: self.N + 1]
        pb = BatchProblem([self._td], self.N, np.asarray(x0, dtype=np.float64).reshape(1, 4),
                          np.array([float(ref_velo)]), np.asarray(road_borders, dtype=np.float64).reshape(1, 2),
                          np.zeros(1, np.int32), np.array([n], np.int32), obs)
        self.last = self._solver.solve(pb)
        return self.last.u[0], self.last.x[0]

    def close(self):
        self._solver.close()
