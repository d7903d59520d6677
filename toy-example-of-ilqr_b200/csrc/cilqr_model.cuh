// Device functions of the CILQR hot path: bicycle model, model Jacobians, ego
// circle centres, obstacle ellipse margin + gradient, barrier / augmented-
// Lagrangian terms.  Written for sm_100a; everything stays in registers, the
// callers own the memory layout.  Reference behaviour cited per function
// (paths relative to the reference tree).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#ifdef CILQR_PARITY
#include "cilqr_pmath.h"
#endif

namespace cilqr {

// ---- build flavours -----------------------------------------------------------
// default        the fast build: CUDA libdevice transcendentals, FMA contraction, the algebraic shortcuts
//                listed in DESIGN.md section 3 (each a few ulp from the reference's operation sequence).
// CILQR_PARITY   libcilqr_b200_parity.so (also needs nvcc -fmad=false): every stage evaluates the reference's
//                own operation sequence (the order the tests' CPU restatement is written in, which is
//                bit-identical to the reference sources) with the portable transcendentals of
//                cilqr_pmath.h, so that a whole free-running solve can be compared with the CPU bit for
//                bit.  Same kernels, same work lists / trial pool / verdict logic; only the arithmetic
//                inside the device functions differs.  Not a performance build.
#if defined(CILQR_PARITY) || defined(CILQR_EXPERIMENT_DIRECT)
constexpr bool kParity = true;  // (host side: one-thread rollout kernels only)
#else
constexpr bool kParity = false;
#endif

// ---- scalar math, overloaded on the compute type --------------------------
#ifdef CILQR_PARITY
template <typename T> __device__ __forceinline__ T m_sin(T v) { return cilqr_pm::pm_sin(v); }
template <typename T> __device__ __forceinline__ T m_cos(T v) { return cilqr_pm::pm_cos(v); }
template <typename T> __device__ __forceinline__ void m_sincos(T v, T* s, T* c) { *s = cilqr_pm::pm_sin(v); *c = cilqr_pm::pm_cos(v); }
template <typename T> __device__ __forceinline__ T m_tan(T v) { return cilqr_pm::pm_tan(v); }
template <typename T> __device__ __forceinline__ T m_atan(T v) { return cilqr_pm::pm_atan(v); }
template <typename T> __device__ __forceinline__ T m_exp(T v) { return cilqr_pm::pm_exp(v); }
template <typename T> __device__ __forceinline__ T m_hypot(T a, T b) { return cilqr_pm::pm_hypot(a, b); }
#else
__device__ __forceinline__ double m_sin(double v) { return sin(v); }
__device__ __forceinline__ float m_sin(float v) { return sinf(v); }
__device__ __forceinline__ double m_cos(double v) { return cos(v); }
__device__ __forceinline__ float m_cos(float v) { return cosf(v); }
__device__ __forceinline__ void m_sincos(double v, double* s, double* c) { sincos(v, s, c); }
__device__ __forceinline__ void m_sincos(float v, float* s, float* c) { sincosf(v, s, c); }
__device__ __forceinline__ double m_tan(double v) { return tan(v); }
__device__ __forceinline__ float m_tan(float v) { return tanf(v); }
__device__ __forceinline__ double m_atan(double v) { return atan(v); }
__device__ __forceinline__ float m_atan(float v) { return atanf(v); }
// exp() for the barrier terms, without a slow path.  A step's cost holds 8 + 2 n_obs independent
// exponentials; CUDA's exp() ends in a data-dependent branch (range check), which keeps the compiler
// from interleaving them, and one evaluation is a ~170-cycle dependent chain on an otherwise idle
// scheduler (ncu: 17-22 warp-cycles per issued instruction in the cost / derivative kernels).  Same
// scheme as any libm exp — k = rint(x log2 e), r = x - k ln2 (two-part ln2), exp(r) by a degree-13
// polynomial (|r| <= ln2/2: truncation 6e-18 relative), scaling by 2^k in two halves so that
// overflow gives inf and underflow 0 — within 1 ulp, the accuracy class of CUDA's exp(); NaN is kept.
__device__ __forceinline__ double m_exp(double x) {
    const double xc = fmin(fmax(x, -1100.0), 1100.0);
    const double shifter = 6755399441055744.0;  // 1.5 * 2^52: the sum's low mantissa bits hold rint()
    const double t = __fma_rn(xc, 1.4426950408889634, shifter);
    const double kf = t - shifter;
    const int k = __double2loint(t);
    double r = __fma_rn(kf, -6.93147180369123816490e-01, xc);
    r = __fma_rn(kf, -1.90821492927058770002e-10, r);
    double p = 1.6059043836821613e-10;
    p = __fma_rn(p, r, 2.0876756987868100e-09);
    p = __fma_rn(p, r, 2.5052108385441720e-08);
    p = __fma_rn(p, r, 2.7557319223985888e-07);
    p = __fma_rn(p, r, 2.7557319223985893e-06);
    p = __fma_rn(p, r, 2.4801587301587302e-05);
    p = __fma_rn(p, r, 1.9841269841269841e-04);
    p = __fma_rn(p, r, 1.3888888888888889e-03);
    p = __fma_rn(p, r, 8.3333333333333332e-03);
    p = __fma_rn(p, r, 4.1666666666666664e-02);
    p = __fma_rn(p, r, 1.6666666666666666e-01);
    p = __fma_rn(p, r, 0.5);
    p = __fma_rn(p, r, 1.0);
    p = __fma_rn(p, r, 1.0);
    const int k1 = k >> 1, k2 = k - k1;
    const double s1 = __hiloint2double((k1 + 1023) << 20, 0), s2 = __hiloint2double((k2 + 1023) << 20, 0);
    const double v = (p * s1) * s2;
    return x == x ? v : x;
}
__device__ __forceinline__ float m_exp(float v) { return expf(v); }
__device__ __forceinline__ double m_hypot(double a, double b) { return hypot(a, b); }
__device__ __forceinline__ float m_hypot(float a, float b) { return hypotf(a, b); }
#endif
__device__ __forceinline__ double m_sqrt(double v) { return sqrt(v); }
__device__ __forceinline__ float m_sqrt(float v) { return sqrtf(v); }
__device__ __forceinline__ double m_fma(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ float m_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double m_fabs(double v) { return fabs(v); }
__device__ __forceinline__ float m_fabs(float v) { return fabsf(v); }
// Loads that must be ISSUED where they are written: the compiler otherwise sinks a software
// prefetch down to its first use (seen in the ncu source view: the serial Riccati / rollout
// chains then stall a full memory latency per step).  volatile asm keeps program order.
__device__ __forceinline__ double ld_early(const double* p) {
    double v;
    asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float ld_early(const float* p) {
    float v;
    asm volatile("ld.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ int ld_early(const int* p) {
    int v;
    asm volatile("ld.global.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// std::max / std::min semantics (NaN in the first argument is kept), as the
// reference uses them (cpp:82, :120, :378, :623).
template <typename T>
__device__ __forceinline__ T std_max(T a, T b) { return (a < b) ? b : a; }
template <typename T>
__device__ __forceinline__ T std_min(T a, T b) { return (b < a) ? b : a; }

// Per-template solver scalars in the compute type (converted once on the host
// from cilqr_params_t; derived ellipse semi-axes follow src/utils.cpp:387-393
// with ego_pnt_radius = width/2, src/cilqr_solver.cpp:330, :78).
template <typename T>
struct DevParams {
    T dt, wheelbase, width;
    T Q[4];  // diag(w_pos, w_pos, w_vel, w_yaw)  (cpp:23-27)
    T R[2];  // diag(w_acc, w_stl)                (cpp:28-30)
    T obs_q1, obs_q2, st_q1, st_q2;
    T acc_max, acc_min, stl_lim, velo_max, velo_min;
    T ell_a2, ell_b2;  // reciprocals of the squared semi-axes, 1/a^2 and 1/b^2
    T alm_rho_init, alm_gamma, max_rho, max_mu;
    T init_lamb, lamb_decay, lamb_amplify, max_lamb, conv_thr, accept_thr;
    int max_iter, solve_type, ref_point, use_last;
    int wp_off, wp_len;  // slice of the handle's waypoint table
};

// src/utils.cpp:262-283 — one Euler step of the rear-axle (ref_point 0) or centre-of-gravity (1)
// bicycle model; all right-hand sides use the old state.  The step is split into its
// transcendental part and its algebra so that the rollout kernels can evaluate the two sin/cos
// pairs of a step — of the yaw and of the steering angle — on two lanes at once, and so that BOTH
// vehicle models run the same instruction stream (a batch may mix them lane by lane; as two
// branches the warp would execute both chains one after the other on every step):
//   rear:    heading = (sin, cos)(yaw),        turn = tan(steer) = sin(steer) / cos(steer)
//   gravity: heading = (sin, cos)(beta + yaw), turn = sin(beta),  beta = atan(tan(steer) / 2)
// with beta's sine and cosine taken straight from tan(beta) = tan(steer) / 2 (beta lies in
// (-pi/2, pi/2), so cos(beta) = 1 / sqrt(1 + tan^2(beta)) > 0) and the heading by the angle-addition
// formulas — no atan and no third sin/cos on the recurrence.  Each of these is within 2-3 ulp of the
// reference's operation sequence (absolute error <= 4e-16 on the heading), the accuracy class of
// CUDA's own tan() / atan().
// `mixed`: warp-uniform, some lane of the warp runs the gravity model (a pure rear-axle warp skips
// the beta algebra altogether).
template <typename T>
__device__ __forceinline__ void step_trig(int ref_point, bool mixed, T s_yaw, T c_yaw, T s_st, T c_st, T* s_head,
                                          T* c_head, T* turn) {
    const T td = s_st / c_st;
    *s_head = s_yaw;
    *c_head = c_yaw;
    *turn = td;
    if (mixed) {
        const T tb = T(0.5) * td;
        const T cb = T(1) / m_sqrt(T(1) + tb * tb);
        const T sb = tb * cb;
        if (ref_point != 0) {
            *s_head = sb * c_yaw + cb * s_yaw;
            *c_head = cb * c_yaw - sb * s_yaw;
            *turn = sb;
        }
    }
}
// warp vote over the lanes that are currently active
__device__ __forceinline__ bool any_gravity_lane(int ref_point) { return __any_sync(__activemask(), ref_point != 0); }
template <typename T>
__device__ __forceinline__ void step_from_trig(const T x[4], T acc, T dt, T dt_over_wb, int ref_point, T s_head,
                                               T c_head, T turn, T out[4]) {
    // dt_over_wb = dt / wheelbase, formed once per trajectory: the reference's (v tan(steer) dt) / L
    // becomes (v tan(steer)) (dt / L) — one rounding apart, and no division left on the yaw recurrence,
    // which is the critical path of every rollout (yaw -> feedback -> steer -> tan -> yaw)
    out[0] = x[0] + x[2] * c_head * dt;
    out[1] = x[1] + x[2] * s_head * dt;
    out[2] = x[2] + acc * dt;
    out[3] = ref_point == 0 ? x[3] + x[2] * turn * dt_over_wb : x[3] + 2 * x[2] * turn * dt_over_wb;
}
template <typename T>
__device__ __forceinline__ void propagate(const T x[4], T acc, T steer, T dt, T wheelbase,
                                          int ref_point, T out[4]) {
#ifdef CILQR_PARITY
    // the reference's own sequence (src/utils.cpp:262-283)
    const T beta = m_atan(m_tan(steer) / 2);
    if (ref_point == 0) {
        out[0] = x[0] + x[2] * m_cos(x[3]) * dt;
        out[1] = x[1] + x[2] * m_sin(x[3]) * dt;
        out[2] = x[2] + acc * dt;
        out[3] = x[3] + x[2] * m_tan(steer) * dt / wheelbase;
    } else {
        out[0] = x[0] + x[2] * m_cos(beta + x[3]) * dt;
        out[1] = x[1] + x[2] * m_sin(beta + x[3]) * dt;
        out[2] = x[2] + acc * dt;
        out[3] = x[3] + 2 * x[2] * m_sin(beta) * dt / wheelbase;
    }
#elif defined(CILQR_EXPERIMENT_DIRECT)
    // development switch: the reference's sequence with libdevice transcendentals (measures what the shared-trig
    // formulation below costs in accuracy)
    const T beta = m_atan(m_tan(steer) / 2);
    if (ref_point == 0) {
        out[0] = x[0] + x[2] * m_cos(x[3]) * dt;
        out[1] = x[1] + x[2] * m_sin(x[3]) * dt;
        out[2] = x[2] + acc * dt;
        out[3] = x[3] + x[2] * m_tan(steer) * dt / wheelbase;
    } else {
        out[0] = x[0] + x[2] * m_cos(beta + x[3]) * dt;
        out[1] = x[1] + x[2] * m_sin(beta + x[3]) * dt;
        out[2] = x[2] + acc * dt;
        out[3] = x[3] + 2 * x[2] * m_sin(beta) * dt / wheelbase;
    }
#else
    T sy, cy, ss, cs, s, c, turn;
    m_sincos(x[3], &sy, &cy);
    m_sincos(steer, &ss, &cs);
    step_trig(ref_point, any_gravity_lane(ref_point), sy, cy, ss, cs, &s, &c, &turn);
    step_from_trig(x, acc, dt, dt / wheelbase, ref_point, s, c, turn, out);
#endif
}

// src/utils.cpp:285-342 — the non-trivial entries of A = df/dx (identity plus
// a02 a03 a12 a13 a32) and B = df/du (b01 b11 b20 b31).  Keeps the reference's
// quirk: the Jacobian's beta is atan(tan(steer/2)), not the step's
// atan(tan(steer)/2), while d(beta)/d(steer) is that of the step's beta.
template <typename T>
__device__ __forceinline__ void model_jacobians(T velo, T yaw, T steer, T dt, T wheelbase,
                                                int ref_point, T a[5], T b[4]) {
    if (ref_point == 0) {
        T s, c;
        m_sincos(yaw, &s, &c);
        a[0] = c * dt;
        a[1] = velo * (-s) * dt;
        a[2] = s * dt;
        a[3] = velo * c * dt;
        a[4] = m_tan(steer) * dt / wheelbase;
        T cd = m_cos(steer);
        b[0] = 0;
        b[1] = 0;
        b[2] = dt;
        b[3] = (velo * dt / wheelbase) / (cd * cd);
    } else {
        T beta = m_atan(m_tan(steer / 2));
        T td = m_tan(steer);
        T t2 = td * td;
        T dbeta = T(0.5) * (1 + t2) / (1 + T(0.25) * t2);
        T s, c;
        m_sincos(beta + yaw, &s, &c);
        a[0] = c * dt;
        a[1] = velo * (-s) * dt;
        a[2] = s * dt;
        a[3] = velo * c * dt;
        a[4] = 2 * m_sin(beta) * dt / wheelbase;
        b[0] = velo * (-s) * dt * dbeta;
        b[1] = velo * c * dt * dbeta;
        b[2] = dt;
        b[3] = (2 * velo * dt / wheelbase) * m_cos(beta) * dbeta;
    }
}

// Ego geometry shared by all obstacles of one step: circle centres
// (src/utils.cpp:344-361) and the yaw rows of their 4x2 Jacobians (:363-385).
template <typename T>
struct EgoCircles {
    T fx, fy, rx, ry;      // front / rear circle centres
    T jf0, jf1, jr0, jr1;  // d(front)/d(yaw), d(rear)/d(yaw)
};

template <typename T>
__device__ __forceinline__ EgoCircles<T> ego_circles(const T x[4], T wheelbase, int ref_point) {
    EgoCircles<T> e;
    T s, c;
    m_sincos(x[3], &s, &c);
    T wx = wheelbase * c, wy = wheelbase * s;
    T half = T(0.5) * wheelbase;
    if (ref_point == 0) {
        e.fx = x[0] + wx;
        e.fy = x[1] + wy;
        e.rx = x[0];
        e.ry = x[1];
        e.jf0 = wheelbase * (-s);
        e.jf1 = wheelbase * c;
        e.jr0 = 0;
        e.jr1 = 0;
    } else {
        e.fx = x[0] + T(0.5) * wx;
        e.fy = x[1] + T(0.5) * wy;
        e.rx = x[0] - T(0.5) * wx;
        e.ry = x[1] - T(0.5) * wy;
        e.jf0 = half * (-s);
        e.jf1 = half * c;
        e.jr0 = -half * (-s);
        e.jr1 = -half * c;
    }
    return e;
}

// src/utils.cpp:395-407 and :409-439 for one circle centre against one
// obstacle sample (ox, oy, sin/cos of its yaw): margin c = 1 - (xs^2/a^2 + ys^2/b^2)
// and, when wanted, its gradient w.r.t. the point.
template <typename T, bool kGrad>
__device__ __forceinline__ T ellipse_margin(T px, T py, T ox, T oy, T so, T co, T inv_a2, T inv_b2, T* gx, T* gy) {
    // inv_a2 = 1/a^2, inv_b2 = 1/b^2 are per-template constants: the reference's divisions by a^2, b^2
    // become multiplications (<= 1 ulp apart; fp64 division costs ~30 issue slots on the SM)
    T dx = px - ox, dy = py - oy;
    T xs = co * dx + so * dy;
    T ys = -so * dx + co * dy;
#ifdef CILQR_PARITY
    // parity build: inv_a2 / inv_b2 hold a^2 and b^2 themselves and the reference's divisions are kept
    T margin = 1 - ((xs * xs) / inv_a2 + (ys * ys) / inv_b2);
    if (kGrad) {
        T g0 = -2 * xs / inv_a2, g1 = -2 * ys / inv_b2;
        *gx = co * g0 + (-so) * g1;
        *gy = so * g0 + co * g1;
    }
#else
    T margin = 1 - ((xs * xs) * inv_a2 + (ys * ys) * inv_b2);
    if (kGrad) {
        T g0 = -2 * xs * inv_a2, g1 = -2 * ys * inv_b2;
        *gx = co * g0 + (-so) * g1;
        *gy = so * g0 + co * g1;
    }
#endif
    return margin;
}

// Lateral offset to the matched waypoint (cpp:235-241): signed *distance to the
// waypoint*, sign from the cross product; sign(0) = +1 (include/utils.hpp:110-117).
template <typename T>
__device__ __forceinline__ T lateral_offset(T px, T py, T rx, T ry, T s, T c, T* d_sign, T* hyp) {
    // s, c = sin / cos of the matched waypoint's yaw (tabulated once per template)
    T ds = (py - ry) * c - (px - rx) * s;
    T h = m_hypot(px - rx, py - ry);
    *d_sign = ds;
    *hyp = h;
    return (ds < 0 ? T(-1) : T(1)) * h;
}

// include/cilqr_solver.hpp:80 and :81-83
template <typename T>
__device__ __forceinline__ T exp_barrier(T c, T q1, T q2) {
    return q1 * m_exp(q2 * c);
}
template <typename T>
__device__ __forceinline__ T alm_item(T c, T rho, T mu) {
    T v = std_max(c + mu / rho, T(0));
    return rho * (v * v) / 2;
}

}  // namespace cilqr
