// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// C driver around the reference's OWN CILQRSolver, compiled from the unmodified
// sources under /root/reference (see Makefile target `ref`; -fno-access-control
// lets this file reach the private stages).  What is substituted and why:
//   * Eigen            -> shim/Eigen/{Core,Dense}  (absent from the image)
//   * fmt, spdlog      -> shim no-ops              (log text only)
//   * matplotlibcpp.h  -> shim stub                (plot helpers in utils.cpp, never called)
//   * GlobalConfig     -> the definitions below: same class (include/global_config.hpp),
//                         but the map is filled from a Params struct instead of yaml-cpp
// The solver arithmetic itself — every line of src/cilqr_solver.cpp and
// src/utils.cpp:262-439 — is the reference's.
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>

#include "cilqr_solver.hpp"
#include "global_config.hpp"

// ---- GlobalConfig without yaml-cpp (declarations: include/global_config.hpp) ----
GlobalConfig* GlobalConfig::instance = nullptr;
void GlobalConfig::load_file(const std::string&) {}
bool GlobalConfig::has_key(std::string key_str) { return config_map.find(key_str) != config_map.end(); }
GlobalConfig* GlobalConfig::get_instance(const std::string&) {
    if (instance == nullptr) instance = new GlobalConfig();
    return instance;
}
template <typename T>
T GlobalConfig::get_config(const std::string& key) const {
    auto it = config_map.find(key);
    if (it != config_map.end()) {
        try {
            return std::any_cast<T>(it->second);
        } catch (const std::bad_any_cast&) {
        }
    }
    return T();
}
void GlobalConfig::destroy_instance() {
    delete instance;
    instance = nullptr;
}
template std::string GlobalConfig::get_config<std::string>(const std::string&) const;
template int GlobalConfig::get_config<int>(const std::string&) const;
template double GlobalConfig::get_config<double>(const std::string&) const;
template bool GlobalConfig::get_config<bool>(const std::string&) const;

namespace {

// same field order as cilqr_oracle::Params / cilqr_params_t
struct Params {
    double dt;
    double w_pos, w_vel, w_yaw, w_acc, w_stl;
    double obstacle_exp_q1, obstacle_exp_q2, state_exp_q1, state_exp_q2;
    double alm_rho_init, alm_gamma, max_rho, max_mu;
    double init_lamb, lamb_decay, lamb_amplify, max_lamb;
    double convergence_threshold, accept_step_threshold;
    double wheelbase, width, length;
    double velo_max, velo_min, yaw_lim, acc_max, acc_min, stl_lim, d_safe;
    int32_t max_iter, solve_type, reference_point, use_last_solution;
};

struct RefSolver {
    GlobalConfig* cfg;
    CILQRSolver* solver;
    int N;
};

ReferenceLine make_line(int M, const double* wx, const double* wy, const double* wyaw) {
    ReferenceLine rl(std::vector<double>{0.0, 1.0}, std::vector<double>{0.0, 0.0});
    rl.x.assign(wx, wx + M);
    rl.y.assign(wy, wy + M);
    rl.yaw.assign(wyaw, wyaw + M);
    rl.longitude.assign(size_t(M), 0.0);
    return rl;
}

std::vector<RoutingLine> make_obs(int n_obs, int obs_len, const double* obs) {
    std::vector<RoutingLine> out(static_cast<size_t>(n_obs));
    for (int j = 0; j < n_obs; ++j)
        for (int k = 0; k < obs_len; ++k) {
            const double* p = obs + (size_t(j) * obs_len + k) * 3;
            out[j].x.push_back(p[0]);
            out[j].y.push_back(p[1]);
            out[j].yaw.push_back(p[2]);
        }
    return out;
}

Eigen::MatrixX2d load_u(int N, const double* u) {
    Eigen::MatrixX2d m = Eigen::MatrixX2d::Zero(N, 2);
    for (int i = 0; i < N; ++i)
        for (int c = 0; c < 2; ++c) m(i, c) = u[i * 2 + c];
    return m;
}
Eigen::MatrixX4d load_x(int rows, const double* x) {
    Eigen::MatrixX4d m = Eigen::MatrixX4d::Zero(rows, 4);
    for (int i = 0; i < rows; ++i)
        for (int c = 0; c < 4; ++c) m(i, c) = x[i * 4 + c];
    return m;
}
void store(const Eigen::Dense& m, double* dst) {
    if (!dst) return;
    for (Eigen::Index i = 0; i < m.rows(); ++i)
        for (Eigen::Index j = 0; j < m.cols(); ++j) dst[i * m.cols() + j] = m(i, j);
}

}  // namespace

extern "C" {

int ref_sizeof_params() { return int(sizeof(Params)); }

void* ref_solver_create(const Params* p, int N) {
    // one private GlobalConfig per solver (the singleton accessor is bypassed)
    auto* cfg = new GlobalConfig();
    auto& m = cfg->config_map;
    m["delta_t"] = p->dt;
    m["lqr/N"] = N;
    m["lqr/nx"] = 4;
    m["lqr/nu"] = 2;
    m["lqr/w_pos"] = p->w_pos;
    m["lqr/w_vel"] = p->w_vel;
    m["lqr/w_yaw"] = p->w_yaw;
    m["lqr/w_acc"] = p->w_acc;
    m["lqr/w_stl"] = p->w_stl;
    m["lqr/slove_type"] = std::string(p->solve_type == 1 ? "alm" : "barrier");
    m["lqr/alm_rho_init"] = p->alm_rho_init;
    m["lqr/alm_gamma"] = p->alm_gamma;
    m["lqr/max_rho"] = p->max_rho;
    m["lqr/max_mu"] = p->max_mu;
    m["lqr/obstacle_exp_q1"] = p->obstacle_exp_q1;
    m["lqr/obstacle_exp_q2"] = p->obstacle_exp_q2;
    m["lqr/state_exp_q1"] = p->state_exp_q1;
    m["lqr/state_exp_q2"] = p->state_exp_q2;
    m["lqr/use_last_solution"] = bool(p->use_last_solution != 0);
    m["iteration/max_iter"] = int(p->max_iter);
    m["iteration/init_lamb"] = p->init_lamb;
    m["iteration/lamb_decay"] = p->lamb_decay;
    m["iteration/lamb_amplify"] = p->lamb_amplify;
    m["iteration/max_lamb"] = p->max_lamb;
    m["iteration/convergence_threshold"] = p->convergence_threshold;
    m["iteration/accept_step_threshold"] = p->accept_step_threshold;
    m["vehicle/reference_point"] = std::string(p->reference_point == 0 ? "rear_center" : "gravity_center");
    m["vehicle/wheelbase"] = p->wheelbase;
    m["vehicle/width"] = p->width;
    m["vehicle/length"] = p->length;
    m["vehicle/velo_max"] = p->velo_max;
    m["vehicle/velo_min"] = p->velo_min;
    m["vehicle/yaw_lim"] = p->yaw_lim;
    m["vehicle/acc_max"] = p->acc_max;
    m["vehicle/acc_min"] = p->acc_min;
    m["vehicle/stl_lim"] = p->stl_lim;
    m["vehicle/d_safe"] = p->d_safe;
    auto* s = new RefSolver{cfg, new CILQRSolver(cfg), N};
    return s;
}

void ref_solver_destroy(void* h) {
    auto* s = static_cast<RefSolver*>(h);
    delete s->solver;
    delete s->cfg;
    delete s;
}

// The public call: CILQRSolver::solve.  info = {current_solve_status}.
int ref_solver_solve(void* h, int M, const double* wx, const double* wy, const double* wyaw, double ref_velo,
                     int n_obs, int obs_len, const double* obs, const double* borders, const double* x0,
                     double* u_out, double* x_out, int32_t* info) {
    auto* s = static_cast<RefSolver*>(h);
    ReferenceLine rl = make_line(M, wx, wy, wyaw);
    std::vector<RoutingLine> ob = make_obs(n_obs, obs_len, obs);
    Eigen::Vector4d x0v{x0[0], x0[1], x0[2], x0[3]};
    Eigen::Vector2d bd{borders[0], borders[1]};
    try {
        auto [u, x] = s->solver->solve(x0v, rl, ref_velo, ob, bd);
        store(u, u_out);
        store(x, x_out);
    } catch (const std::out_of_range&) {
        return -2;
    }
    if (info) info[0] = int32_t(s->solver->current_solve_status);
    return 0;
}

// ---- private stages (reached through -fno-access-control) --------------------
int ref_total_cost(void* h, int M, const double* wx, const double* wy, const double* wyaw, double ref_velo,
                   int n_obs, int obs_len, const double* obs, const double* borders, const double* u,
                   const double* x, double* J) {
    auto* s = static_cast<RefSolver*>(h);
    ReferenceLine rl = make_line(M, wx, wy, wyaw);
    std::vector<RoutingLine> ob = make_obs(n_obs, obs_len, obs);
    Eigen::Vector2d bd{borders[0], borders[1]};
    *J = s->solver->get_total_cost(load_u(s->N, u), load_x(s->N + 1, x), rl, ref_velo, ob, bd);
    return 0;
}

int ref_ref_points(void* h, int M, const double* wx, const double* wy, const double* wyaw, int rows,
                   const double* x, double* pts /*[rows][3]*/) {
    auto* s = static_cast<RefSolver*>(h);
    ReferenceLine rl = make_line(M, wx, wy, wyaw);
    store(s->solver->get_ref_exact_points(load_x(rows, x), rl), pts);
    return 0;
}

// backward_pass: fills the cached l_* (fresh solver status RUNNING => recomputed), A/B, d, K, delta_V.
int ref_backward_pass(void* h, int M, const double* wx, const double* wy, const double* wyaw, double ref_velo,
                      int n_obs, int obs_len, const double* obs, const double* borders, const double* u,
                      const double* x, double lamb, double* lx, double* lu, double* lxx, double* luu, double* A,
                      double* Bm, double* d, double* K, double* dV, int32_t* status) {
    auto* s = static_cast<RefSolver*>(h);
    ReferenceLine rl = make_line(M, wx, wy, wyaw);
    std::vector<RoutingLine> ob = make_obs(n_obs, obs_len, obs);
    Eigen::Vector2d bd{borders[0], borders[1]};
    Eigen::MatrixX2d um = load_u(s->N, u);
    Eigen::MatrixX4d xm = load_x(s->N + 1, x);
    s->solver->current_solve_status = LQRSolveStatus::RUNNING;
    auto [dd, KK, dv] = s->solver->backward_pass(um, xm, lamb, rl, ref_velo, ob, bd);
    store(s->solver->l_x, lx);
    store(s->solver->l_u, lu);
    store(s->solver->l_xx, lxx);
    store(s->solver->l_uu, luu);
    auto [dfdx, dfdu] = utils::get_kinematic_model_derivatives(xm, um, s->solver->dt, s->solver->wheelbase,
                                                               uint32_t(s->N), s->solver->reference_point);
    store(dfdx, A);
    store(dfdu, Bm);
    store(dd, d);
    store(KK, K);
    if (dV) {
        dV[0] = dv[0];
        dV[1] = dv[1];
    }
    if (status) *status = int32_t(s->solver->current_solve_status);
    return 0;
}

int ref_forward_pass(void* h, const double* u, const double* x, const double* d, const double* K, double alpha,
                     double* new_u, double* new_x) {
    auto* s = static_cast<RefSolver*>(h);
    Eigen::MatrixX2d dm = load_u(s->N, d);
    Eigen::MatrixX4d Km = load_x(s->N * 2, K);
    auto [nu, nx] = s->solver->forward_pass(load_u(s->N, u), load_x(s->N + 1, x), dm, Km, alpha);
    store(nu, new_u);
    store(nx, new_x);
    return 0;
}

// ---- scenario preprocessing of the reference (src/utils.cpp:21-35, src/cubic_spline.cpp) ----
// Samples ReferenceLine(x, y, width) with the reference's own code; returns the number of
// waypoints written (<= cap).
int ref_reference_line(int n, const double* x, const double* y, double width, int cap, double* wx, double* wy,
                       double* wyaw, double* longitude) {
    ReferenceLine rl(std::vector<double>(x, x + n), std::vector<double>(y, y + n), width);
    int M = int(rl.size());
    for (int i = 0; i < M && i < cap; ++i) {
        wx[i] = rl.x[size_t(i)];
        wy[i] = rl.y[size_t(i)];
        wyaw[i] = rl.yaw[size_t(i)];
        if (longitude) longitude[i] = rl.longitude[size_t(i)];
    }
    return M;
}

}  // extern "C"
