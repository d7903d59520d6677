// C++ drop-in for the reference's solver class, on top of the C ABI (include/cilqr_b200.h).
//
// Mirrors include/cilqr_solver.hpp:31-42 of the reference:
//     explicit CILQRSolver(const GlobalConfig* const config);
//     std::tuple<MatrixX2d, MatrixX4d> solve(const Vector4d& x0, const ReferenceLine& ref_waypoints,
//                                            double ref_velo, const std::vector<RoutingLine>& obs_preds,
//                                            const Vector2d& road_boaders);
// so that src/motion_planning.cpp:178 and :194-197 compile unchanged against it:
//   * where <Eigen/Core> exists the Eigen types are used as they are (column-major matrices
//     filled from the library's row-major output);
//   * elsewhere (this image has no Eigen) the minimal fixed-purpose types below stand in, with the
//     members motion_planning.cpp touches (row(i), operator()(i,j), rows()).
// ReferenceLine / RoutingLine need only their public x, y, yaw vectors (include/utils.hpp:44-46,
// :65-67); GlobalConfig only get_config<T>(key) (include/global_config.hpp:33-34).
// A batch of one: the batched entry points are in the C ABI.
#pragma once

#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <tuple>
#include <cstring>
#include <vector>

#include "../../include/cilqr_b200.h"

#if defined(CILQR_COMPAT_USE_EIGEN) && __has_include(<Eigen/Core>)
#include <Eigen/Core>
namespace cilqr_compat {
using Vector2d = Eigen::Vector2d;
using Vector4d = Eigen::Vector4d;
using MatrixX2d = Eigen::MatrixX2d;
using MatrixX4d = Eigen::MatrixX4d;
}  // namespace cilqr_compat
#else
namespace cilqr_compat {
using Vector2d = std::array<double, 2>;
using Vector4d = std::array<double, 4>;
// rows x C matrix, row-major, with the accessors the reference's caller uses
template <int C>
struct MatrixXC {
    int n_rows = 0;
    std::vector<double> v;
    MatrixXC() = default;
    explicit MatrixXC(int rows, int /*cols*/ = C) : n_rows(rows), v(size_t(rows) * C, 0.0) {}
    int rows() const { return n_rows; }
    int cols() const { return C; }
    double& operator()(int i, int j) { return v[size_t(i) * C + j]; }
    double operator()(int i, int j) const { return v[size_t(i) * C + j]; }
    std::array<double, C> row(int i) const {
        std::array<double, C> r;
        for (int j = 0; j < C; ++j) r[j] = v[size_t(i) * C + j];
        return r;
    }
};
using MatrixX2d = MatrixXC<2>;
using MatrixX4d = MatrixXC<4>;
}  // namespace cilqr_compat
#endif

namespace cilqr_compat {

// include/cilqr_solver.hpp:23-29
enum class LQRSolveStatus { RUNNING, CONVERGED, BACKWARD_PASS_FAIL, FORWARD_PASS_FAIL, FORWARD_PASS_SMALL_STEP };

// Fills cilqr_params_t from anything with the reference's get_config<T>(key) (src/cilqr_solver.cpp:18-72).
template <typename Config>
cilqr_params_t params_from_config(const Config* c, int* N_out) {
    cilqr_params_t p{};
    p.dt = c->template get_config<double>("delta_t");
    *N_out = c->template get_config<int>("lqr/N");
    if (c->template get_config<int>("lqr/nx") != 4 || c->template get_config<int>("lqr/nu") != 2)
        throw std::invalid_argument("cilqr_b200: the bicycle model has nx = 4, nu = 2");
    p.w_pos = c->template get_config<double>("lqr/w_pos");
    p.w_vel = c->template get_config<double>("lqr/w_vel");
    p.w_yaw = c->template get_config<double>("lqr/w_yaw");
    p.w_acc = c->template get_config<double>("lqr/w_acc");
    p.w_stl = c->template get_config<double>("lqr/w_stl");
    p.use_last_solution = c->template get_config<bool>("lqr/use_last_solution") ? 1 : 0;
    const std::string st = c->template get_config<std::string>("lqr/slove_type");
    p.solve_type = st == "alm" ? 1 : 0;  // anything else: barrier, as the reference falls back (cpp:38-41)
    // the reference reads only the set that matches the mode (cpp:42-52); reading both is harmless
    p.obstacle_exp_q1 = c->template get_config<double>("lqr/obstacle_exp_q1");
    p.obstacle_exp_q2 = c->template get_config<double>("lqr/obstacle_exp_q2");
    p.state_exp_q1 = c->template get_config<double>("lqr/state_exp_q1");
    p.state_exp_q2 = c->template get_config<double>("lqr/state_exp_q2");
    p.alm_rho_init = c->template get_config<double>("lqr/alm_rho_init");
    p.alm_gamma = c->template get_config<double>("lqr/alm_gamma");
    p.max_rho = c->template get_config<double>("lqr/max_rho");
    p.max_mu = c->template get_config<double>("lqr/max_mu");
    p.max_iter = c->template get_config<int>("iteration/max_iter");
    p.init_lamb = c->template get_config<double>("iteration/init_lamb");
    p.lamb_decay = c->template get_config<double>("iteration/lamb_decay");
    p.lamb_amplify = c->template get_config<double>("iteration/lamb_amplify");
    p.max_lamb = c->template get_config<double>("iteration/max_lamb");
    p.convergence_threshold = c->template get_config<double>("iteration/convergence_threshold");
    p.accept_step_threshold = c->template get_config<double>("iteration/accept_step_threshold");
    p.wheelbase = c->template get_config<double>("vehicle/wheelbase");
    p.width = c->template get_config<double>("vehicle/width");
    p.length = c->template get_config<double>("vehicle/length");
    p.velo_max = c->template get_config<double>("vehicle/velo_max");
    p.velo_min = c->template get_config<double>("vehicle/velo_min");
    p.yaw_lim = c->template get_config<double>("vehicle/yaw_lim");
    p.acc_max = c->template get_config<double>("vehicle/acc_max");
    p.acc_min = c->template get_config<double>("vehicle/acc_min");
    p.stl_lim = c->template get_config<double>("vehicle/stl_lim");
    p.d_safe = c->template get_config<double>("vehicle/d_safe");
    p.reference_point = c->template get_config<std::string>("vehicle/reference_point") == "rear_center" ? 0 : 1;
    return p;
}

class CILQRSolver {
  public:
    CILQRSolver() = delete;
    template <typename Config>
    explicit CILQRSolver(const Config* const config, int device = 0, int max_obs = 16, int dtype = CILQR_F64)
        : max_obs_(max_obs) {
        cilqr_params_t p = params_from_config(config, &N_);
        check(cilqr_b200_create(&p, device, 1, N_, max_obs, dtype, &h_));
    }
    CILQRSolver(const CILQRSolver&) = delete;
    CILQRSolver& operator=(const CILQRSolver&) = delete;
    ~CILQRSolver() { cilqr_b200_destroy(h_); }

    // RefLine: .x, .y, .yaw vectors and size(); Routing: .x, .y, .yaw vectors (throws
    // std::out_of_range for a track shorter than N+1, like RoutingLine::operator[], src/utils.cpp:53-55)
    template <typename RefLine, typename Routing>
    std::tuple<MatrixX2d, MatrixX4d> solve(const Vector4d& x0, const RefLine& ref_waypoints, double ref_velo,
                                           const std::vector<Routing>& obs_preds, const Vector2d& road_boaders) {
        // the reference reads the line afresh on every solve(); the device copy is refreshed whenever the
        // CONTENT differs from what was uploaded last (a line edited in place, or another line at the same
        // address, must not be served from a stale copy)
        if (!same_line(ref_waypoints.x, line_x_) || !same_line(ref_waypoints.y, line_y_) ||
            !same_line(ref_waypoints.yaw, line_yaw_)) {
            check(cilqr_b200_set_template(h_, 0, nullptr, ref_waypoints.x.data(), ref_waypoints.y.data(),
                                          ref_waypoints.yaw.data(), int(ref_waypoints.x.size())));
            line_x_.assign(ref_waypoints.x.begin(), ref_waypoints.x.end());
            line_y_.assign(ref_waypoints.y.begin(), ref_waypoints.y.end());
            line_yaw_.assign(ref_waypoints.yaw.begin(), ref_waypoints.yaw.end());
        }
        const int n = int(obs_preds.size());
        if (n > max_obs_) throw std::invalid_argument("cilqr_b200: more obstacles than max_obs");
        const int L = N_ + 1;
        obs_.assign(size_t(max_obs_ > 0 ? max_obs_ : 1) * L * 3, 0.0);
        for (int j = 0; j < n; ++j) {
            const Routing& r = obs_preds[size_t(j)];
            if (r.x.size() < size_t(L) || r.y.size() < size_t(L) || r.yaw.size() < size_t(L))
                throw std::out_of_range("Index out of range");
            for (int k = 0; k < L; ++k) {
                double* o = &obs_[(size_t(j) * L + k) * 3];
                o[0] = r.x[size_t(k)];
                o[1] = r.y[size_t(k)];
                o[2] = r.yaw[size_t(k)];
            }
        }
        const double x0v[4] = {x0[0], x0[1], x0[2], x0[3]};
        const double bd[2] = {road_boaders[0], road_boaders[1]};
        const int32_t tmpl = 0, nobs = n;
        u_.assign(size_t(N_) * 2, 0.0);
        x_.assign(size_t(N_ + 1) * 4, 0.0);
        int32_t status = 0;
        check(cilqr_b200_solve_batch(h_, 1, x0v, &ref_velo, bd, &tmpl, &nobs, obs_.data(), L, u_.data(), x_.data(),
                                     J_, nullptr, nullptr, nullptr, &status, &iters_, &exit_));
        status_ = static_cast<LQRSolveStatus>(status);
        MatrixX2d u(N_, 2);
        MatrixX4d x(N_ + 1, 4);
        for (int i = 0; i < N_; ++i)
            for (int c = 0; c < 2; ++c) u(i, c) = u_[size_t(i) * 2 + c];
        for (int i = 0; i <= N_; ++i)
            for (int c = 0; c < 4; ++c) x(i, c) = x_[size_t(i) * 4 + c];
        return std::make_tuple(u, x);
    }

    // extras the reference keeps private
    LQRSolveStatus status() const { return status_; }
    int iterations() const { return iters_; }
    double initial_cost() const { return J_[0]; }
    double final_cost() const { return J_[1]; }
    int horizon() const { return N_; }

  private:
    template <typename V>
    static bool same_line(const V& a, const std::vector<double>& b) {
        return a.size() == b.size() && (b.empty() || std::memcmp(a.data(), b.data(), b.size() * sizeof(double)) == 0);
    }
    static void check(int rc) {
        if (rc == CILQR_ERR_RANGE) throw std::out_of_range(cilqr_b200_last_error());
        if (rc != 0) throw std::runtime_error(std::string("cilqr_b200: ") + cilqr_b200_last_error());
    }
    cilqr_handle_t* h_ = nullptr;
    int N_ = 0, max_obs_ = 0;
    std::vector<double> line_x_, line_y_, line_yaw_;  // the reference line as last uploaded
    std::vector<double> obs_, u_, x_;
    double J_[2] = {0, 0};
    int32_t iters_ = 0, exit_ = 0;
    LQRSolveStatus status_ = LQRSolveStatus::RUNNING;
};

}  // namespace cilqr_compat
