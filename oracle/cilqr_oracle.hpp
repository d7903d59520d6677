// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement (Eigen-free C++17, templated on the scalar) of the CILQR hot
// path of PuYuuu/toy-example-of-iLQR.  Only tests/, __graft_entry__.smoke() and
// the cpu_baseline / --impl reference legs of bench.py may build, link or call
// anything in oracle/.  The CUDA library (libcilqr_b200.so) never does.
//
// Parity pin: the reference ships no tests or golden vectors, and its real
// dependency (Eigen3) is absent from this image.  The pin used instead is
// oracle/_ref: the reference's own, unmodified src/cilqr_solver.cpp +
// src/utils.cpp + src/cubic_spline.cpp compiled in place against the small
// Eigen-API stand-in under oracle/shim/ (see oracle/Makefile).
// tests/test_oracle_vs_ref.py holds this restatement to that build; the
// committed fixtures under tests/golden/ were generated from it.
//
// Every function cites the reference lines it follows (paths relative to
// /root/reference).  All sums are written in source order; the library is
// built with -ffp-contract=off to mirror the reference's plain x86-64 -O3.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace cilqr_oracle {

// Transcendentals of the restatement.  Default flavour: glibc's (std::), as the reference calls them; this is
// the flavour pinned bit-for-bit to the reference sources (oracle/_ref).  -DCILQR_ORACLE_PMATH ("pm" flavour,
// liboracle_pm.so): the portable implementations the PARITY build of the CUDA library uses too
// (toy-example-of-ilqr_b200/csrc/cilqr_pmath.h — pure IEEE add/mul/div/sqrt/fma, identical bits under gcc and nvcc),
// so that the GPU parity build can be compared with this restatement bit for bit.  sq(x) = std::pow(x, 2).
#ifdef CILQR_ORACLE_PMATH
}  // namespace cilqr_oracle
#include "../toy-example-of-ilqr_b200/csrc/cilqr_pmath.h"
namespace cilqr_oracle {
namespace om {
template <typename T> inline T sin(T v) { return cilqr_pm::pm_sin(v); }
template <typename T> inline T cos(T v) { return cilqr_pm::pm_cos(v); }
template <typename T> inline T tan(T v) { return cilqr_pm::pm_tan(v); }
template <typename T> inline T atan(T v) { return cilqr_pm::pm_atan(v); }
template <typename T> inline T exp(T v) { return cilqr_pm::pm_exp(v); }
template <typename T> inline T hypot(T a, T b) { return cilqr_pm::pm_hypot(a, b); }
template <typename T> inline T sq(T v) { return v * v; }
}  // namespace om
#else
namespace om {
using std::sin;
using std::cos;
using std::tan;
using std::atan;
using std::exp;
using std::hypot;
template <typename T> inline T sq(T v) { return std::pow(v, T(2)); }
}  // namespace om
#endif


// Plain-data mirror of the scalars the reference constructor reads
// (src/cilqr_solver.cpp:17-83).  Field order is shared with the ctypes
// Structure in tests/ (python side: cilqr_params fields).
struct Params {
    double dt;
    double w_pos, w_vel, w_yaw, w_acc, w_stl;
    double obstacle_exp_q1, obstacle_exp_q2, state_exp_q1, state_exp_q2;
    double alm_rho_init, alm_gamma, max_rho, max_mu;
    double init_lamb, lamb_decay, lamb_amplify, max_lamb;
    double convergence_threshold, accept_step_threshold;
    double wheelbase, width, length;
    double velo_max, velo_min, yaw_lim, acc_max, acc_min, stl_lim, d_safe;
    int32_t max_iter;
    int32_t solve_type;        // 0 = barrier, 1 = alm   (cpp:33-41)
    int32_t reference_point;   // 0 = rear_center, 1 = gravity_center (cpp:72-76)
    int32_t use_last_solution; // cpp:32
};

// include/cilqr_solver.hpp:23-29
enum Status : int32_t {
    RUNNING = 0,
    CONVERGED = 1,
    BACKWARD_PASS_FAIL = 2,
    FORWARD_PASS_FAIL = 3,
    FORWARD_PASS_SMALL_STEP = 4,
};

// How solve() left its loop (cpp:127-148).
enum ExitReason : int32_t { EXIT_MAX_ITER = 0, EXIT_CONVERGED = 1, EXIT_MAX_LAMB = 2 };

constexpr double kEps = 1e-5;  // include/utils.hpp:28

// include/utils.hpp:110-117 — note sign(0) = sign(-0) = +1.
template <typename T>
inline int sign_of(T v) {
    return v < 0 ? -1 : 1;
}

// One planning problem, reference wire format flattened:
// waypoints = ReferenceLine::{x,y,yaw} (include/utils.hpp:44-46),
// obs[j][k] = RoutingLine::operator[] of obstacle j at tick k (src/utils.cpp:52-58).
template <typename T>
struct Problem {
    const T* wx = nullptr;
    const T* wy = nullptr;
    const T* wyaw = nullptr;
    int M = 0;
    T ref_velo = 0;
    int n_obs = 0;
    int obs_len = 0;       // samples per obstacle track (must be >= N+1)
    const T* obs = nullptr;  // [n_obs][obs_len][3]
    T border_up = 0;       // road_boaders[0]
    T border_lo = 0;       // road_boaders[1]
};

// ---------------------------------------------------------------------------
// L1: model and geometry (src/utils.cpp:262-439)
// ---------------------------------------------------------------------------

// src/utils.cpp:262-283
template <typename T>
inline void kinematic_propagate(const T x[4], const T u[2], T dt, T wheelbase, int ref_point,
                                T out[4]) {
    T beta = om::atan(om::tan(u[1]) / 2);
    if (ref_point == 0) {
        out[0] = x[0] + x[2] * om::cos(x[3]) * dt;
        out[1] = x[1] + x[2] * om::sin(x[3]) * dt;
        out[2] = x[2] + u[0] * dt;
        out[3] = x[3] + x[2] * om::tan(u[1]) * dt / wheelbase;
    } else {
        out[0] = x[0] + x[2] * om::cos(beta + x[3]) * dt;
        out[1] = x[1] + x[2] * om::sin(beta + x[3]) * dt;
        out[2] = x[2] + u[0] * dt;
        out[3] = x[3] + 2 * x[2] * om::sin(beta) * dt / wheelbase;
    }
}

// src/utils.cpp:285-342 — one step's A (4x4 row-major) and B (4x2 row-major).
// Reproduces the reference's beta mismatch: the Jacobian uses
// atan(tan(delta/2)) while the step uses atan(tan(delta)/2).
template <typename T>
inline void model_derivatives(const T x[4], const T u[2], T dt, T wheelbase, int ref_point,
                              T A[16], T B[8]) {
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) A[r * 4 + c] = (r == c) ? T(1) : T(0);
    for (int i = 0; i < 8; ++i) B[i] = 0;
    T velo = x[2], yaw = x[3], delta = u[1];
    T beta = om::atan(om::tan(delta / 2));
    T tan_d = om::tan(delta);
    T tan_sq = tan_d * tan_d;
    T beta_over_stl = T(0.5) * (1 + tan_sq) / (1 + T(0.25) * tan_sq);
    if (ref_point == 0) {
        A[0 * 4 + 2] = om::cos(yaw) * dt;
        A[0 * 4 + 3] = velo * (-om::sin(yaw)) * dt;
        A[1 * 4 + 2] = om::sin(yaw) * dt;
        A[1 * 4 + 3] = velo * om::cos(yaw) * dt;
        A[3 * 4 + 2] = om::tan(delta) * dt / wheelbase;
        B[2 * 2 + 0] = dt;
        B[3 * 2 + 1] = (velo * dt / wheelbase) / (om::cos(delta) * om::cos(delta));
    } else {
        A[0 * 4 + 2] = om::cos(beta + yaw) * dt;
        A[0 * 4 + 3] = velo * (-om::sin(beta + yaw)) * dt;
        A[1 * 4 + 2] = om::sin(beta + yaw) * dt;
        A[1 * 4 + 3] = velo * om::cos(beta + yaw) * dt;
        A[3 * 4 + 2] = 2 * om::sin(beta) * dt / wheelbase;
        B[0 * 2 + 1] = velo * (-om::sin(beta + yaw)) * dt * beta_over_stl;
        B[1 * 2 + 1] = velo * om::cos(beta + yaw) * dt * beta_over_stl;
        B[2 * 2 + 0] = dt;
        B[3 * 2 + 1] = (2 * velo * dt / wheelbase) * om::cos(beta) * beta_over_stl;
    }
}

// src/utils.cpp:344-361
template <typename T>
inline void front_rear_centers(const T x[4], T wheelbase, int ref_point, T front[2], T rear[2]) {
    T yaw = x[3];
    T wv[2] = {wheelbase * om::cos(yaw), wheelbase * om::sin(yaw)};
    if (ref_point == 0) {
        front[0] = x[0] + wv[0];
        front[1] = x[1] + wv[1];
        rear[0] = x[0];
        rear[1] = x[1];
    } else {
        front[0] = x[0] + T(0.5) * wv[0];
        front[1] = x[1] + T(0.5) * wv[1];
        rear[0] = x[0] - T(0.5) * wv[0];
        rear[1] = x[1] - T(0.5) * wv[1];
    }
}

// src/utils.cpp:387-393 with ego_pnt_radius = width/2 (cpp:330); obs_attr =
// {ego width, ego length, d_safe} (cpp:78).
template <typename T>
inline void ellipse_scales(const Params& p, T ab[2]) {
    T radius = T(0.5) * T(p.width);
    ab[0] = T(0.5) * T(p.length) + T(p.d_safe) * 6 + radius;
    ab[1] = T(0.5) * T(p.width) + T(p.d_safe) + radius;
}

// src/utils.cpp:395-407
template <typename T>
inline T ellipse_margin(const T pnt[2], const T obs[3], const T ab[2]) {
    T theta = obs[2];
    T dx = pnt[0] - obs[0], dy = pnt[1] - obs[1];
    T c = om::cos(theta), s = om::sin(theta);
    T xs = c * dx + s * dy;
    T ys = -s * dx + c * dy;
    return 1 - (om::sq(xs) / om::sq(ab[0]) + om::sq(ys) / om::sq(ab[1]));
}

// src/utils.cpp:409-439 — gradient of the margin w.r.t. the point.
template <typename T>
inline void ellipse_margin_grad(const T pnt[2], const T obs[3], const T ab[2], T g[2]) {
    T theta = obs[2];
    T dx = pnt[0] - obs[0], dy = pnt[1] - obs[1];
    T c = om::cos(theta), s = om::sin(theta);
    T xs = c * dx + s * dy;
    T ys = -s * dx + c * dy;
    T gs[2] = {-2 * xs / om::sq(ab[0]), -2 * ys / om::sq(ab[1])};
    // rotation^T * gs, then identity * that (:427-436)
    g[0] = c * gs[0] + (-s) * gs[1];
    g[1] = s * gs[0] + c * gs[1];
}

// cpp:326-335 — (front, rear) margins.
template <typename T>
inline void obstacle_constr(const Params& p, const T x[4], const T obs[3], T c[2]) {
    T fr[2], rr[2], ab[2];
    front_rear_centers(x, T(p.wheelbase), p.reference_point, fr, rr);
    ellipse_scales(p, ab);
    c[0] = ellipse_margin(fr, obs, ab);
    c[1] = ellipse_margin(rr, obs, ab);
}

// cpp:715-739 with src/utils.cpp:363-385 — d(margin)/d(state) for front and rear.
template <typename T>
inline void obstacle_constr_grad(const Params& p, const T x[4], const T obs[3], T gf[4], T gr[4]) {
    T fr[2], rr[2], ab[2], gpf[2], gpr[2];
    front_rear_centers(x, T(p.wheelbase), p.reference_point, fr, rr);
    ellipse_scales(p, ab);
    ellipse_margin_grad(fr, obs, ab, gpf);
    ellipse_margin_grad(rr, obs, ab, gpr);
    T yaw = x[3];
    T half = T(0.5) * T(p.wheelbase);
    // 4x2 Jacobians (rows = state component, cols = point component)
    T Jf[4][2] = {{1, 0}, {0, 1}, {0, 0}, {half * (-om::sin(yaw)), half * om::cos(yaw)}};
    T Jr[4][2] = {{1, 0}, {0, 1}, {0, 0}, {-half * (-om::sin(yaw)), -half * om::cos(yaw)}};
    if (p.reference_point == 0) {
        Jf[3][0] = T(p.wheelbase) * (-om::sin(yaw));
        Jf[3][1] = T(p.wheelbase) * om::cos(yaw);
        Jr[3][0] = 0;
        Jr[3][1] = 0;
    }
    for (int r = 0; r < 4; ++r) {
        gf[r] = Jf[r][0] * gpf[0] + Jf[r][1] * gpf[1];
        gr[r] = Jr[r][0] * gpr[0] + Jr[r][1] * gpr[1];
    }
}

// ---------------------------------------------------------------------------
// L2: solver (src/cilqr_solver.cpp)
// ---------------------------------------------------------------------------

template <typename T>
struct BackwardResult {
    std::vector<T> d;  // [N][2]
    std::vector<T> K;  // [N][2][4]
    T dV[2];
};

// Per-iteration record for lockstep comparisons.
struct IterTrace {
    int32_t status;      // status after iter_step
    int32_t alpha_index; // accepted / converged trial index, -1 if none
    int32_t effective;   // effective_flag after iter_step
    double ori_cost;
    double new_cost;     // cost returned by iter_step
    double lamb_after;   // lambda after the schedule update
};

template <typename T>
class Solver {
  public:
    Solver(const Params& params, int horizon) : p(params), N(horizon) {
        status = RUNNING;
        first_solve = true;
        alm_rho = 0;
    }

    Params p;
    int N;
    int32_t status;
    bool first_solve;
    std::vector<T> last_u;             // [N][2]
    std::vector<T> l_x, l_u, l_xx, l_uu;  // [(N+1)][4], [N][2], [(N+1)][4][4], [N][2][2]
    std::vector<T> alm_mu, alm_mu_next;   // [N][8+2*n_obs]
    T alm_rho;
    // results of the most recent solve
    BackwardResult<T> last_bw;
    std::vector<IterTrace> trace;
    int32_t iters = 0;
    int32_t exit_reason = 0;
    T final_lamb = 0;
    T J_init = 0;
    T J_final = 0;

    // cpp:289-314 — nearest waypoint by monotone first-local-minimum scan.
    // Returns indices; ref[i] = (wx, wy, wyaw)[idx[i]].
    void ref_match(const Problem<T>& pb, const T* x, int rows, std::vector<int>& idx) const {
        idx.assign(rows, 0);
        uint16_t start = 0;
        for (int i = 0; i < rows; ++i) {
            int32_t min_idx = -1;
            T min_d = std::numeric_limits<T>::max();
            for (size_t j = start; j < size_t(pb.M); ++j) {
                T cur = om::hypot(x[i * 4 + 0] - pb.wx[j], x[i * 4 + 1] - pb.wy[j]);
                if (min_idx < 0 || cur < min_d) {
                    min_idx = int32_t(j);
                    min_d = cur;
                } else {
                    break;
                }
            }
            idx[i] = min_idx;
            start = uint16_t(min_idx);
        }
    }

    T alm_item(T c, T rho, T mu) const {  // hpp:81-83
        return rho * om::sq(std::max(c + mu / rho, T(0))) / 2;
    }
    T exp_barrier(T c, T q1, T q2) const { return q1 * om::exp(q2 * c); }  // hpp:80

    // The eight box constraints of step k in the reference's order
    // (cpp:222-241 / :507-519): acc up/lo, steer up/lo, velocity up/lo, lateral up/lo.
    void box_constraints(const Problem<T>& pb, const T* uk, const T* xk, const T* ref, T c[8],
                         T* d_sign_out, T* hyp_out) const {
        c[0] = uk[0] - T(p.acc_max);
        c[1] = T(p.acc_min) - uk[0];
        c[2] = uk[1] - T(p.stl_lim);
        c[3] = -T(p.stl_lim) - uk[1];
        c[4] = xk[2] - T(p.velo_max);
        c[5] = T(p.velo_min) - xk[2];
        T d_sign = (xk[1] - ref[1]) * om::cos(ref[2]) - (xk[0] - ref[0]) * om::sin(ref[2]);
        T hyp = om::hypot(xk[0] - ref[0], xk[1] - ref[1]);
        T cur_d = sign_of(d_sign) * hyp;
        c[6] = cur_d - (pb.border_up - T(p.width) / 2);
        c[7] = (pb.border_lo + T(p.width) / 2) - cur_d;
        if (d_sign_out) *d_sign_out = d_sign;
        if (hyp_out) *hyp_out = hyp;
    }

    // cpp:199-287.  step_cost (optional, [N+1]) receives the per-step split:
    // step k holds the state term of x_k, the control term of u_k (k < N) and
    // the constraint terms of step k (k >= 1).
    T total_cost(const Problem<T>& pb, const T* u, const T* x, T* step_cost = nullptr) const {
        std::vector<int> idx;
        ref_match(pb, x, N + 1, idx);
        const T Q[4] = {T(p.w_pos), T(p.w_pos), T(p.w_vel), T(p.w_yaw)};
        const T R[2] = {T(p.w_acc), T(p.w_stl)};
        if (step_cost)
            for (int k = 0; k <= N; ++k) step_cost[k] = 0;
        T states_devt = 0;
        for (int k = 0; k <= N; ++k) {
            T ref[4] = {pb.wx[idx[k]], pb.wy[idx[k]], pb.ref_velo, pb.wyaw[idx[k]]};
            T s = 0;
            for (int c = 0; c < 4; ++c) {
                T e = x[k * 4 + c] - ref[c];
                s += e * Q[c] * e;
            }
            states_devt += s;
            if (step_cost) step_cost[k] += s;
        }
        T ctrl_energy = 0;
        for (int k = 0; k < N; ++k) {
            T s = 0;
            for (int c = 0; c < 2; ++c) s += u[k * 2 + c] * R[c] * u[k * 2 + c];
            ctrl_energy += s;
            if (step_cost) step_cost[k] += s;
        }
        T J_prime = states_devt + ctrl_energy;

        T J_barrier = 0;
        int ncol = 8 + 2 * pb.n_obs;
        for (int k = 1; k <= N; ++k) {
            const T* uk = u + (k - 1) * 2;
            const T* xk = x + k * 4;
            T ref[3] = {pb.wx[idx[k]], pb.wy[idx[k]], pb.wyaw[idx[k]]};
            T c[8];
            box_constraints(pb, uk, xk, ref, c, nullptr, nullptr);
            T Jk = 0;
            if (p.solve_type == 0) {
                Jk = exp_barrier(c[0], T(p.state_exp_q1), T(p.state_exp_q2)) +
                     exp_barrier(c[1], T(p.state_exp_q1), T(p.state_exp_q2)) +
                     exp_barrier(c[2], T(p.state_exp_q1), T(p.state_exp_q2)) +
                     exp_barrier(c[3], T(p.state_exp_q1), T(p.state_exp_q2)) +
                     exp_barrier(c[4], T(p.state_exp_q1), T(p.state_exp_q2)) +
                     exp_barrier(c[5], T(p.state_exp_q1), T(p.state_exp_q2)) +
                     exp_barrier(c[6], T(p.state_exp_q1), T(p.state_exp_q2)) +
                     exp_barrier(c[7], T(p.state_exp_q1), T(p.state_exp_q2));
            } else {
                const T* mu = alm_mu.data() + size_t(k - 1) * ncol;
                Jk = alm_item(c[0], alm_rho, mu[0]) + alm_item(c[1], alm_rho, mu[1]) +
                     alm_item(c[2], alm_rho, mu[2]) + alm_item(c[3], alm_rho, mu[3]) +
                     alm_item(c[4], alm_rho, mu[4]) + alm_item(c[5], alm_rho, mu[5]) +
                     alm_item(c[6], alm_rho, mu[6]) + alm_item(c[7], alm_rho, mu[7]);
            }
            for (int j = 0; j < pb.n_obs; ++j) {
                const T* ob = pb.obs + (size_t(j) * pb.obs_len + k) * 3;
                T oc[2];
                obstacle_constr(p, xk, ob, oc);
                if (p.solve_type == 0) {
                    Jk += exp_barrier(oc[0], T(p.obstacle_exp_q1), T(p.obstacle_exp_q2));
                    Jk += exp_barrier(oc[1], T(p.obstacle_exp_q1), T(p.obstacle_exp_q2));
                } else {
                    const T* mu = alm_mu.data() + size_t(k - 1) * ncol;
                    Jk += alm_item(oc[0], alm_rho, mu[8 + 2 * j]);
                    Jk += alm_item(oc[1], alm_rho, mu[9 + 2 * j]);
                }
            }
            J_barrier += Jk;
            if (step_cost) step_cost[k] += Jk;
        }
        return J_prime + J_barrier;
    }

    // One constraint's contribution (cpp:692-699 barrier, :701-713 ALM):
    // grad += g, hess += H, with c_dot of length n (2 or 4).
    void add_constraint(T c, const T* c_dot, int n, T q1, T q2, T mu, T* grad, T* hess) const {
        if (p.solve_type == 0) {
            T b = exp_barrier(c, q1, q2);
            T q2sq = om::sq(q2);
            for (int r = 0; r < n; ++r) grad[r] += q2 * b * c_dot[r];
            for (int r = 0; r < n; ++r)
                for (int cc = 0; cc < n; ++cc) hess[r * n + cc] += q2sq * b * (c_dot[r] * c_dot[cc]);
        } else {
            if ((c + mu / alm_rho) > 0) {
                T bd[4];
                for (int r = 0; r < n; ++r) bd[r] = alm_rho * (c + mu / alm_rho) * c_dot[r];
                for (int r = 0; r < n; ++r) grad[r] += bd[r];
                for (int r = 0; r < n; ++r)
                    for (int cc = 0; cc < n; ++cc) hess[r * n + cc] += bd[r] * c_dot[cc];
            }
        }
    }

    // cpp:463-690.  Returns false when the barrier-mode cache was reused (:469-474).
    bool cost_derivatives(const Problem<T>& pb, const T* u, const T* x) {
        if (p.solve_type == 0 && status != RUNNING && status != FORWARD_PASS_SMALL_STEP) {
            status = RUNNING;
            return false;
        }
        status = RUNNING;
        l_x.assign(size_t(N + 1) * 4, 0);
        l_u.assign(size_t(N) * 2, 0);
        l_xx.assign(size_t(N + 1) * 16, 0);
        l_uu.assign(size_t(N) * 4, 0);
        std::vector<int> idx;
        ref_match(pb, x, N + 1, idx);
        const T Q[4] = {T(p.w_pos), T(p.w_pos), T(p.w_vel), T(p.w_yaw)};
        const T R[2] = {T(p.w_acc), T(p.w_stl)};
        int ncol = 8 + 2 * pb.n_obs;

        for (int k = 1; k <= N; ++k) {
            const T* uk = u + (k - 1) * 2;
            const T* xk = x + k * 4;
            T ref[3] = {pb.wx[idx[k]], pb.wy[idx[k]], pb.wyaw[idx[k]]};
            T c[8], d_sign, hyp;
            box_constraints(pb, uk, xk, ref, c, &d_sign, &hyp);
            const T cu[4][2] = {{1, 0}, {-1, 0}, {0, 1}, {0, -1}};
            T cx[4][4] = {{0, 0, 1, 0}, {0, 0, -1, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};
            // lateral gradient (:527-533); hypot is re-evaluated per component in the reference
            cx[2][0] = (xk[0] - ref[0]) / om::hypot(xk[0] - ref[0], xk[1] - ref[1]);
            cx[2][1] = (xk[1] - ref[1]) / om::hypot(xk[0] - ref[0], xk[1] - ref[1]);
            if (d_sign < 0)
                for (int r = 0; r < 4; ++r) cx[2][r] = -1 * cx[2][r];
            for (int r = 0; r < 4; ++r) cx[3][r] = -1 * cx[2][r];

            const T* mu = (p.solve_type == 1) ? alm_mu.data() + size_t(k - 1) * ncol : nullptr;
            T* gu = l_u.data() + size_t(k - 1) * 2;
            T* Hu = l_uu.data() + size_t(k - 1) * 4;
            T* gx = l_x.data() + size_t(k) * 4;
            T* Hx = l_xx.data() + size_t(k) * 16;
            T q1 = T(p.state_exp_q1), q2 = T(p.state_exp_q2);
            for (int m = 0; m < 4; ++m) add_constraint(c[m], cu[m], 2, q1, q2, mu ? mu[m] : T(0), gu, Hu);
            for (int m = 0; m < 4; ++m)
                add_constraint(c[4 + m], cx[m], 4, q1, q2, mu ? mu[4 + m] : T(0), gx, Hx);
            if (p.solve_type == 1) {
                T* mn = alm_mu_next.data() + size_t(k - 1) * ncol;
                for (int m = 0; m < 8; ++m)
                    mn[m] = std::min(std::max(mu[m] + alm_rho * c[m], T(0)), T(p.max_mu));
            }
            for (int j = 0; j < pb.n_obs; ++j) {
                const T* ob = pb.obs + (size_t(j) * pb.obs_len + k) * 3;
                T oc[2], gf[4], gr[4];
                obstacle_constr(p, xk, ob, oc);
                obstacle_constr_grad(p, xk, ob, gf, gr);
                // the reference sums front+rear first, then adds to the row (:662-664)
                T g2[4] = {0, 0, 0, 0}, H2[16] = {0};
                T oq1 = T(p.obstacle_exp_q1), oq2 = T(p.obstacle_exp_q2);
                add_constraint(oc[0], gf, 4, oq1, oq2, mu ? mu[8 + 2 * j] : T(0), g2, H2);
                add_constraint(oc[1], gr, 4, oq1, oq2, mu ? mu[9 + 2 * j] : T(0), g2, H2);
                for (int r = 0; r < 4; ++r) gx[r] += g2[r];
                for (int r = 0; r < 16; ++r) Hx[r] += H2[r];
                if (p.solve_type == 1) {
                    T* mn = alm_mu_next.data() + size_t(k - 1) * ncol;
                    mn[8 + 2 * j] = std::min(std::max(mu[8 + 2 * j] + alm_rho * oc[0], T(0)), T(p.max_mu));
                    mn[9 + 2 * j] = std::min(std::max(mu[9 + 2 * j] + alm_rho * oc[1], T(0)), T(p.max_mu));
                }
            }
        }
        // prime part added last: l = l_prime + l_barrier (:686-689)
        for (int k = 0; k < N; ++k) {
            for (int c = 0; c < 2; ++c) {
                l_u[k * 2 + c] = 2 * (u[k * 2 + c] * R[c]) + l_u[k * 2 + c];
                l_uu[k * 4 + c * 2 + c] = 2 * R[c] + l_uu[k * 4 + c * 2 + c];
            }
        }
        for (int k = 0; k <= N; ++k) {
            T ref[4] = {pb.wx[idx[k]], pb.wy[idx[k]], pb.ref_velo, pb.wyaw[idx[k]]};
            for (int c = 0; c < 4; ++c) {
                l_x[k * 4 + c] = 2 * (x[k * 4 + c] - ref[c]) * Q[c] + l_x[k * 4 + c];
                l_xx[k * 16 + c * 4 + c] = 2 * Q[c] + l_xx[k * 16 + c * 4 + c];
            }
        }
        return true;
    }

    // The Riccati recursion of cpp:391-439 on explicit inputs (dense reference
    // layout).  Sets status to BACKWARD_PASS_FAIL on a non-PD Quu (:415-420).
    // lx [(N+1)][4], lu [N][2], lxx [(N+1)][16], luu [N][4], A [N][16], B [N][8].
    static int32_t riccati(int N, const T* lx, const T* lu, const T* lxx, const T* luu, const T* A,
                           const T* B, T lamb, BackwardResult<T>& out) {
        out.d.assign(size_t(N) * 2, 0);
        out.K.assign(size_t(N) * 8, 0);
        out.dV[0] = out.dV[1] = 0;
        T Vx[4], Vxx[16];
        for (int c = 0; c < 4; ++c) Vx[c] = lx[size_t(N) * 4 + c];
        for (int c = 0; c < 16; ++c) Vxx[c] = lxx[size_t(N) * 16 + c];
        for (int i = N - 1; i >= 0; --i) {
            const T* Ai = A + size_t(i) * 16;
            const T* Bi = B + size_t(i) * 8;
            T Qx[4], Qu[2], Qxx[16], Quu[4], Qux[8];
            T AtV[16], BtV[8];
            for (int r = 0; r < 4; ++r) {
                T s = 0;
                for (int k = 0; k < 4; ++k) s += Ai[k * 4 + r] * Vx[k];
                Qx[r] = lx[size_t(i) * 4 + r] + s;
            }
            for (int r = 0; r < 2; ++r) {
                T s = 0;
                for (int k = 0; k < 4; ++k) s += Bi[k * 2 + r] * Vx[k];
                Qu[r] = lu[size_t(i) * 2 + r] + s;
            }
            for (int r = 0; r < 4; ++r)
                for (int c = 0; c < 4; ++c) {
                    T s = 0;
                    for (int k = 0; k < 4; ++k) s += Ai[k * 4 + r] * Vxx[k * 4 + c];
                    AtV[r * 4 + c] = s;
                }
            for (int r = 0; r < 2; ++r)
                for (int c = 0; c < 4; ++c) {
                    T s = 0;
                    for (int k = 0; k < 4; ++k) s += Bi[k * 2 + r] * Vxx[k * 4 + c];
                    BtV[r * 4 + c] = s;
                }
            for (int r = 0; r < 4; ++r)
                for (int c = 0; c < 4; ++c) {
                    T s = 0;
                    for (int k = 0; k < 4; ++k) s += AtV[r * 4 + k] * Ai[k * 4 + c];
                    Qxx[r * 4 + c] = lxx[size_t(i) * 16 + r * 4 + c] + s;
                }
            for (int r = 0; r < 2; ++r)
                for (int c = 0; c < 2; ++c) {
                    T s = 0;
                    for (int k = 0; k < 4; ++k) s += BtV[r * 4 + k] * Bi[k * 2 + c];
                    Quu[r * 2 + c] = (luu[size_t(i) * 4 + r * 2 + c] + s) + lamb * (r == c ? T(1) : T(0));
                }
            for (int r = 0; r < 2; ++r)
                for (int c = 0; c < 4; ++c) {
                    T s = 0;
                    for (int k = 0; k < 4; ++k) s += BtV[r * 4 + k] * Ai[k * 4 + c];
                    Qux[r * 4 + c] = T(0) + s;  // l_ux is identically zero (:79-80)
                }
            // Eigen::LLT (lower, unblocked) on the 2x2: NumericalIssue iff a pivot <= 0; NaN passes.
            {
                T a00 = Quu[0];
                if (a00 <= T(0)) return BACKWARD_PASS_FAIL;
                T l00 = std::sqrt(a00);
                T l10 = Quu[2] / l00;
                T a11 = Quu[3] - l10 * l10;
                if (a11 <= T(0)) return BACKWARD_PASS_FAIL;
            }
            // Matrix2d::inverse(): adjugate times 1/det (:421)
            T det = Quu[0] * Quu[3] - Quu[2] * Quu[1];
            T invdet = T(1) / det;
            T inv[4] = {Quu[3] * invdet, -Quu[1] * invdet, -Quu[2] * invdet, Quu[0] * invdet};
            T* di = out.d.data() + size_t(i) * 2;
            T* Ki = out.K.data() + size_t(i) * 8;
            for (int r = 0; r < 2; ++r) di[r] = (-inv[r * 2 + 0]) * Qu[0] + (-inv[r * 2 + 1]) * Qu[1];
            for (int r = 0; r < 2; ++r)
                for (int c = 0; c < 4; ++c)
                    Ki[r * 4 + c] = (-inv[r * 2 + 0]) * Qux[0 * 4 + c] + (-inv[r * 2 + 1]) * Qux[1 * 4 + c];
            // value update (:427-432), left-to-right products, regularised Quu
            T KtQuu[8];  // 4x2
            for (int r = 0; r < 4; ++r)
                for (int c = 0; c < 2; ++c)
                    KtQuu[r * 2 + c] = Ki[0 * 4 + r] * Quu[0 * 2 + c] + Ki[1 * 4 + r] * Quu[1 * 2 + c];
            T nVx[4], nVxx[16];
            for (int r = 0; r < 4; ++r) {
                T t1 = KtQuu[r * 2 + 0] * di[0] + KtQuu[r * 2 + 1] * di[1];
                T t2 = Ki[0 * 4 + r] * Qu[0] + Ki[1 * 4 + r] * Qu[1];
                T t3 = Qux[0 * 4 + r] * di[0] + Qux[1 * 4 + r] * di[1];
                nVx[r] = ((Qx[r] + t1) + t2) + t3;
            }
            for (int r = 0; r < 4; ++r)
                for (int c = 0; c < 4; ++c) {
                    T t1 = KtQuu[r * 2 + 0] * Ki[0 * 4 + c] + KtQuu[r * 2 + 1] * Ki[1 * 4 + c];
                    T t2 = Ki[0 * 4 + r] * Qux[0 * 4 + c] + Ki[1 * 4 + r] * Qux[1 * 4 + c];
                    T t3 = Qux[0 * 4 + r] * Ki[0 * 4 + c] + Qux[1 * 4 + r] * Ki[1 * 4 + c];
                    nVxx[r * 4 + c] = ((Qxx[r * 4 + c] + t1) + t2) + t3;
                }
            for (int c = 0; c < 4; ++c) Vx[c] = nVx[c];
            for (int c = 0; c < 16; ++c) Vxx[c] = nVxx[c];
            // expected reduction (:435-436)
            T hd0 = T(0.5) * di[0], hd1 = T(0.5) * di[1];
            T r0 = hd0 * Quu[0] + hd1 * Quu[2];
            T r1 = hd0 * Quu[1] + hd1 * Quu[3];
            out.dV[0] += r0 * di[0] + r1 * di[1];
            out.dV[1] += di[0] * Qu[0] + di[1] * Qu[1];
        }
        return RUNNING;
    }

    // src/utils.cpp:285-342 over the horizon: A [N][16], B [N][8].
    void dyn_derivatives(const T* u, const T* x, std::vector<T>& A, std::vector<T>& B) const {
        A.assign(size_t(N) * 16, 0);
        B.assign(size_t(N) * 8, 0);
        for (int i = 0; i < N; ++i)
            model_derivatives(x + i * 4, u + i * 2, T(p.dt), T(p.wheelbase), p.reference_point,
                              A.data() + size_t(i) * 16, B.data() + size_t(i) * 8);
    }

    // cpp:383-440
    void backward_pass(const Problem<T>& pb, const T* u, const T* x, T lamb, BackwardResult<T>& out) {
        cost_derivatives(pb, u, x);
        std::vector<T> A, B;
        dyn_derivatives(u, x, A, B);
        int32_t st = riccati(N, l_x.data(), l_u.data(), l_xx.data(), l_uu.data(), A.data(), B.data(),
                             lamb, out);
        if (st == BACKWARD_PASS_FAIL) status = BACKWARD_PASS_FAIL;
    }

    // cpp:442-461
    void forward_pass(const T* u, const T* x, const T* d, const T* K, T alpha, T* new_u,
                      T* new_x) const {
        for (int c = 0; c < 4; ++c) new_x[c] = x[c];
        for (int i = 0; i < N; ++i) {
            T dx[4];
            for (int c = 0; c < 4; ++c) dx[c] = new_x[i * 4 + c] - x[i * 4 + c];
            T nu_i[2];
            for (int r = 0; r < 2; ++r) {
                T s = 0;
                for (int c = 0; c < 4; ++c) s += K[size_t(i) * 8 + r * 4 + c] * dx[c];
                nu_i[r] = (u[i * 2 + r] + s) + alpha * d[i * 2 + r];
            }
            new_u[i * 2 + 0] = nu_i[0];
            new_u[i * 2 + 1] = nu_i[1];
            kinematic_propagate(new_x + i * 4, nu_i, T(p.dt), T(p.wheelbase), p.reference_point,
                                new_x + (i + 1) * 4);
        }
    }

    // cpp:155-197 (cold) and :163-180 (warm)
    void init_trajectory(const T* x0, bool warm, std::vector<T>& u, std::vector<T>& x) const {
        u.assign(size_t(N) * 2, 0);
        x.assign(size_t(N + 1) * 4, 0);
        if (warm) {
            for (int i = 0; i < N - 1; ++i) {
                u[i * 2 + 0] = last_u[(i + 1) * 2 + 0];
                u[i * 2 + 1] = last_u[(i + 1) * 2 + 1];
            }
            u[(N - 1) * 2 + 0] = last_u[(N - 1) * 2 + 0];
            u[(N - 1) * 2 + 1] = last_u[(N - 1) * 2 + 1];
        }
        for (int c = 0; c < 4; ++c) x[c] = x0[c];
        for (int i = 0; i < N; ++i)
            kinematic_propagate(x.data() + i * 4, u.data() + i * 2, T(p.dt), T(p.wheelbase),
                                p.reference_point, x.data() + (i + 1) * 4);
    }

    // cpp:337-381.  Returns the cost iter_step returns; new_u/new_x as returned.
    T iter_step(const Problem<T>& pb, const std::vector<T>& u, const std::vector<T>& x, T lamb,
                bool& effective, std::vector<T>& new_u, std::vector<T>& new_x, int& alpha_index,
                T& ori_cost_out) {
        T ori_cost = total_cost(pb, u.data(), x.data());
        ori_cost_out = ori_cost;
        alpha_index = -1;
        backward_pass(pb, u.data(), x.data(), lamb, last_bw);
        if (status == BACKWARD_PASS_FAIL) {
            new_u = u;
            new_x = x;
            return ori_cost;  // effective keeps its previous value (:345-347)
        }
        T new_J = std::numeric_limits<T>::max();
        new_u.assign(size_t(N) * 2, 0);
        new_x.assign(size_t(N + 1) * 4, 0);
        effective = false;
        int ai = 0;
        for (T alpha = 1; alpha > T(1e-6); alpha *= T(0.5), ++ai) {
            forward_pass(u.data(), x.data(), last_bw.d.data(), last_bw.K.data(), alpha, new_u.data(),
                         new_x.data());
            new_J = total_cost(pb, new_u.data(), new_x.data());
            const T actual = ori_cost - new_J;
            if (std::fabs(alpha - T(1)) < T(kEps) && std::fabs(actual) < T(p.convergence_threshold)) {
                status = CONVERGED;
                alpha_index = ai;
                return new_J;
            }
            T approx = -(alpha * alpha * last_bw.dV[0] + alpha * last_bw.dV[1]);
            if (actual > T(0) && (approx < T(0) || actual / approx > T(p.accept_step_threshold))) {
                if (std::fabs(alpha - T(1)) > T(kEps)) status = FORWARD_PASS_SMALL_STEP;
                effective = true;
                alpha_index = ai;
                return new_J;
            }
        }
        // all trials rejected (:377-380); the ALM update runs in both modes
        if (p.solve_type == 1) {
            alm_mu = alm_mu_next;
            alm_rho = std::min((1 + T(p.alm_gamma)) * alm_rho, T(p.max_rho));
        }
        status = FORWARD_PASS_FAIL;
        return new_J;
    }

    // cpp:85-153
    void solve(const Problem<T>& pb, const T* x0, std::vector<T>& u, std::vector<T>& x,
               bool keep_trace = false) {
        if (p.solve_type == 1 && (!p.use_last_solution || (p.use_last_solution && first_solve))) {
            alm_rho = T(p.alm_rho_init);
            alm_mu.assign(size_t(N) * (8 + 2 * pb.n_obs), 0);
            alm_mu_next.assign(size_t(N) * (8 + 2 * pb.n_obs), 0);
        }
        status = RUNNING;
        if (!first_solve && p.use_last_solution) {
            init_trajectory(x0, true, u, x);
        } else {
            init_trajectory(x0, false, u, x);
            first_solve = false;
        }
        J_init = total_cost(pb, u.data(), x.data());
        T lamb = T(p.init_lamb);
        bool effective = false;
        trace.clear();
        iters = 0;
        exit_reason = EXIT_MAX_ITER;
        std::vector<T> nu, nx;
        for (int itr = 0; itr < p.max_iter; ++itr) {
            int ai;
            T ori;
            T newJ = iter_step(pb, u, x, lamb, effective, nu, nx, ai, ori);
            if (effective) {
                x = nx;
                u = nu;
            }
            if (status == BACKWARD_PASS_FAIL || status == FORWARD_PASS_FAIL) {
                lamb = std::max(T(p.lamb_amplify), lamb * T(p.lamb_amplify));
            } else if (status == RUNNING) {
                lamb *= T(p.lamb_decay);
            }
            iters = itr + 1;
            if (keep_trace)
                trace.push_back(IterTrace{status, ai, effective ? 1 : 0, double(ori), double(newJ),
                                          double(lamb)});
            if (lamb > T(p.max_lamb)) {
                exit_reason = EXIT_MAX_LAMB;
                break;
            } else if (status == CONVERGED) {
                exit_reason = EXIT_CONVERGED;
                break;
            }
        }
        last_u = u;
        final_lamb = lamb;
        J_final = total_cost(pb, u.data(), x.data());
    }
};

}  // namespace cilqr_oracle
