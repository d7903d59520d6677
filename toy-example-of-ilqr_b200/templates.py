"""The four scenario templates of the reference, as flat "section/key" maps.

These restate the *values* of /root/reference/config/scenario_{two_straight,two_borrow,
three_straight,three_bend}.yaml in the key space GlobalConfig builds from them
(src/global_config.cpp:22-92), including its defaults (alm_* :34-37, reference_point :54-55).
tests/test_templates.py checks them against the YAML files whenever the reference tree is
mounted; on the GPU box (no reference tree) these tables are the source of truth.
Every shipped YAML has lqr/N = 30 (config/*.yaml:5); the benchmark configs override N.
"""
import copy

_COMMON = {
    "delta_t": 0.1,
    "lqr/N": 30, "lqr/nx": 4, "lqr/nu": 2,
    "lqr/w_pos": 1.0, "lqr/w_vel": 1.0, "lqr/w_yaw": 20.0, "lqr/w_acc": 0.5,
    "lqr/slove_type": "barrier",
    "lqr/alm_gamma": 0.0, "lqr/max_rho": 20.0,
    "lqr/obstacle_exp_q1": 5.5, "lqr/obstacle_exp_q2": 5.75, "lqr/state_exp_q1": 3.0,
    "iteration/max_iter": 100, "iteration/init_lamb": 0.0, "iteration/lamb_decay": 0.5,
    "iteration/lamb_amplify": 2.0, "iteration/max_lamb": 1000.0,
    "iteration/convergence_threshold": 0.01, "iteration/accept_step_threshold": 0.5,
    "vehicle/wheelbase": 2.8, "vehicle/width": 2.0, "vehicle/length": 4.5,
    "vehicle/velo_min": 0.0, "vehicle/yaw_lim": 1.57, "vehicle/acc_max": 3.0, "vehicle/acc_min": -3.0,
}


def _mk(**kw):
    d = copy.deepcopy(_COMMON)
    d.update(kw)
    return d


TEMPLATES = {
    # config/scenario_two_straight.yaml
    "two_straight": _mk(**{
        "max_simulation_time": 12.0,
        "lqr/w_stl": 20.0, "lqr/alm_rho_init": 20.0, "lqr/max_mu": 120.0,
        "lqr/state_exp_q2": 3.5, "lqr/use_last_solution": False,
        "vehicle/reference_point": "rear_center", "vehicle/target_velocity": 8.0,
        "vehicle/velo_max": 15.0, "vehicle/stl_lim": 0.12, "vehicle/d_safe": 1.0,
        "laneline/reference/x": [-10.0, 0.0, 50.0, 100.0, 150.0, 200.0],
        "laneline/reference/y": [0.0, 0.0, 0.0, 0.0, 0.0, 0.0],
        "laneline/border": [-1.8, 1.8, 5.4],
        "laneline/center_line": [0.0, 3.6],
        "initial_condition": [[0, 0, 8.0, 0], [30, 0, 3.0, 0], [35, 3.6, 5, 0], [15, 3.6, 2.5, 0]],
    }),
    # config/scenario_two_borrow.yaml
    "two_borrow": _mk(**{
        "max_simulation_time": 15.0,
        "lqr/w_stl": 50.0, "lqr/alm_rho_init": 20.0, "lqr/max_mu": 120.0,
        "lqr/state_exp_q2": 3.5, "lqr/use_last_solution": False,
        "vehicle/reference_point": "gravity_center", "vehicle/target_velocity": 8.0,
        "vehicle/velo_max": 15.0, "vehicle/stl_lim": 0.12, "vehicle/d_safe": 0.9,
        "laneline/reference/x": [-10.0, 0.0, 50.0, 100.0, 150.0, 200.0],
        "laneline/reference/y": [0.0, 0.0, 0.0, 0.0, 0.0, 0.0],
        "laneline/border": [-1.8, 1.8, 5.4],
        "laneline/center_line": [0.0, 3.6],
        "initial_condition": [[0, 0, 8.0, 0], [20, 0, 3.0, 0], [30, 0, 6.5, 0], [35, 3.6, 8.0, 3.1416],
                              [110, 3.6, 3.0, 3.1416]],
    }),
    # config/scenario_three_straight.yaml
    "three_straight": _mk(**{
        "max_simulation_time": 10.0,
        "lqr/w_stl": 30.0, "lqr/alm_rho_init": 0.0, "lqr/max_mu": 100.0,
        "lqr/state_exp_q2": 3.6, "lqr/use_last_solution": True,
        "vehicle/reference_point": "gravity_center", "vehicle/target_velocity": 9.0,
        "vehicle/velo_max": 10.0, "vehicle/stl_lim": 0.12, "vehicle/d_safe": 0.9,
        "laneline/reference/x": [-10.0, 0.0, 50.0, 100.0, 150.0, 200.0],
        "laneline/reference/y": [7.2, 7.2, 7.2, 7.2, 7.2, 7.2],
        "laneline/border": [-9.0, -5.4, -1.8, 1.8],
        "laneline/center_line": [0.0, -3.6, -7.2],
        "initial_condition": [[0, 0, 6.0, 0], [20, 0, 3.0, 0], [40, 0, 3.0, 0], [50, 0, 3.0, 0],
                              [0, 3.6, 5.0, 0], [35, 3.6, 5.0, 0], [50, 3.6, 5.0, 0], [5, 7.2, 7, 0],
                              [50, 7.2, 6.0, 0]],
    }),
    # config/scenario_three_bend.yaml
    "three_bend": _mk(**{
        "max_simulation_time": 15.0,
        "lqr/w_stl": 25.0, "lqr/alm_rho_init": 20.0, "lqr/max_mu": 100.0,
        "lqr/state_exp_q2": 3.5, "lqr/use_last_solution": False,
        "vehicle/reference_point": "gravity_center", "vehicle/target_velocity": 8.0,
        "vehicle/velo_max": 10.0, "vehicle/stl_lim": 0.2, "vehicle/d_safe": 0.8,
        "laneline/reference/x": [-20.0, -5.0, 10.0, 20.0, 35.0, 70.0, 100.0, 150.0],
        "laneline/reference/y": [1.0, 1.0, 1.0, 5.0, 6.5, 0.0, 0.0, 0.0],
        "laneline/border": [-1.8, 1.8, 5.4, 9.0],
        "laneline/center_line": [0.0, 3.6, 7.2],
        "initial_condition": [[-10, 1, 4, 0], [10, 0, 4, 0.5236], [25, 10, 4, 0.5236], [-15, 9, 8, 0]],
    }),
}

TEMPLATE_ORDER = ["two_straight", "two_borrow", "three_straight", "three_bend"]


def flatten_yaml(doc):
    """A parsed reference YAML document -> the flat map GlobalConfig::load_file builds
    (src/global_config.cpp:22-92), defaults included."""
    lqr, it, veh, lane = doc["lqr"], doc["iteration"], doc["vehicle"], doc["laneline"]
    m = {
        "max_simulation_time": float(doc["max_simulation_time"]),
        "delta_t": float(doc["delta_t"]),
        "lqr/N": int(lqr["N"]), "lqr/nx": int(lqr["nx"]), "lqr/nu": int(lqr["nu"]),
        "lqr/w_pos": float(lqr["w_pos"]), "lqr/w_vel": float(lqr["w_vel"]), "lqr/w_yaw": float(lqr["w_yaw"]),
        "lqr/w_acc": float(lqr["w_acc"]), "lqr/w_stl": float(lqr["w_stl"]),
        "lqr/slove_type": str(lqr["slove_type"]),
        "lqr/alm_rho_init": float(lqr.get("alm_rho_init", 1.0)),
        "lqr/alm_gamma": float(lqr.get("alm_gamma", 0.0)),
        "lqr/max_rho": float(lqr.get("max_rho", 100.0)),
        "lqr/max_mu": float(lqr.get("max_mu", 1000.0)),
        "lqr/obstacle_exp_q1": float(lqr["obstacle_exp_q1"]), "lqr/obstacle_exp_q2": float(lqr["obstacle_exp_q2"]),
        "lqr/state_exp_q1": float(lqr["state_exp_q1"]), "lqr/state_exp_q2": float(lqr["state_exp_q2"]),
        "lqr/use_last_solution": bool(lqr["use_last_solution"]),
        "iteration/max_iter": int(it["max_iter"]), "iteration/init_lamb": float(it["init_lamb"]),
        "iteration/lamb_decay": float(it["lamb_decay"]), "iteration/lamb_amplify": float(it["lamb_amplify"]),
        "iteration/max_lamb": float(it["max_lamb"]),
        "iteration/convergence_threshold": float(it["convergence_threshold"]),
        "iteration/accept_step_threshold": float(it["accept_step_threshold"]),
        "vehicle/reference_point": str(veh.get("reference_point", "gravity_center")),
        "vehicle/target_velocity": float(veh["target_velocity"]),
        "vehicle/wheelbase": float(veh["wheelbase"]), "vehicle/width": float(veh["width"]),
        "vehicle/length": float(veh["length"]), "vehicle/velo_max": float(veh["velo_max"]),
        "vehicle/velo_min": float(veh["velo_min"]), "vehicle/yaw_lim": float(veh["yaw_lim"]),
        "vehicle/acc_max": float(veh["acc_max"]), "vehicle/acc_min": float(veh["acc_min"]),
        "vehicle/stl_lim": float(veh["stl_lim"]), "vehicle/d_safe": float(veh["d_safe"]),
        "laneline/reference/x": [float(v) for v in lane["reference"]["x"]],
        "laneline/reference/y": [float(v) for v in lane["reference"]["y"]],
        "laneline/border": [float(v) for v in lane["border"]],
        "laneline/center_line": [float(v) for v in lane["center_line"]],
        "initial_condition": [[float(v) for v in row] for row in doc["initial_condition"]],
    }
    return m


def load_yaml(path):
    import yaml
    with open(path) as f:
        return flatten_yaml(yaml.safe_load(f))
