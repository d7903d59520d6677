// ORACLE — TEST INFRASTRUCTURE ONLY (see cilqr_oracle.hpp header).
// extern "C" surface over the restatement so tests/ and bench.py's cpu_baseline
// leg can drive it through ctypes.  All arrays cross the boundary as double in
// the reference's per-solve layout (row-major): u [N][2], x [N+1][4],
// K [N][2][4], d [N][2], obs [n_obs][obs_len][3].  dtype 0 computes in fp64,
// dtype 1 converts to float, computes in fp32 and widens the results.
#include "cilqr_oracle.hpp"

#include <atomic>
#include <thread>

using namespace cilqr_oracle;

namespace {

template <typename T>
std::vector<T> cast_in(const double* src, size_t n) {
    std::vector<T> v(n);
    for (size_t i = 0; i < n; ++i) v[i] = T(src[i]);
    return v;
}
template <typename T>
void cast_out(const std::vector<T>& v, double* dst) {
    if (!dst) return;
    for (size_t i = 0; i < v.size(); ++i) dst[i] = double(v[i]);
}
template <typename T>
void cast_out(const T* v, size_t n, double* dst) {
    if (!dst) return;
    for (size_t i = 0; i < n; ++i) dst[i] = double(v[i]);
}

// Owns the converted copies a Problem<T> points into.
template <typename T>
struct ProblemStore {
    std::vector<T> wx, wy, wyaw, obs;
    Problem<T> pb;
    ProblemStore(int M, const double* wx_, const double* wy_, const double* wyaw_, double ref_velo,
                 int n_obs, int obs_len, const double* obs_, const double* borders)
        : wx(cast_in<T>(wx_, M)),
          wy(cast_in<T>(wy_, M)),
          wyaw(cast_in<T>(wyaw_, M)),
          obs(cast_in<T>(obs_, size_t(n_obs) * obs_len * 3)) {
        pb.wx = wx.data();
        pb.wy = wy.data();
        pb.wyaw = wyaw.data();
        pb.M = M;
        pb.ref_velo = T(ref_velo);
        pb.n_obs = n_obs;
        pb.obs_len = obs_len;
        pb.obs = obs.data();
        pb.border_up = T(borders[0]);
        pb.border_lo = T(borders[1]);
    }
};

// dtype codes of the C surface: 0 = fp64, 1 = fp32, 2 = x87 long double (64-bit mantissa; the "truth" the
// parity tests measure fp64 implementations against; glibc flavour only).
#ifdef CILQR_ORACLE_PMATH
#define ORACLE_DISPATCH(dtype, FN, ...) ((dtype) == 0 ? FN<double>(__VA_ARGS__) : FN<float>(__VA_ARGS__))
#else
#define ORACLE_DISPATCH(dtype, FN, ...) \
    ((dtype) == 0 ? FN<double>(__VA_ARGS__) : (dtype) == 1 ? FN<float>(__VA_ARGS__) : FN<long double>(__VA_ARGS__))
#endif

struct AnySolver {
    int dtype;
    int N;
    virtual ~AnySolver() {}
};
template <typename T>
struct TypedSolver : AnySolver {
    Solver<T> s;
    TypedSolver(const Params& p, int N_) : s(p, N_) {}
};
template <typename T>
AnySolver* make_solver(const Params* p, int N) {
    return new TypedSolver<T>(*p, N);
}

template <typename T>
int do_total_cost(const Params* p, int N, int M, const double* wx, const double* wy,
                  const double* wyaw, double ref_velo, int n_obs, int obs_len, const double* obs,
                  const double* borders, const double* u, const double* x, const double* alm_mu,
                  double alm_rho, double* J, double* step_cost) {
    ProblemStore<T> st(M, wx, wy, wyaw, ref_velo, n_obs, obs_len, obs, borders);
    Solver<T> s(*p, N);
    if (p->solve_type == 1) {
        s.alm_rho = T(alm_rho);
        s.alm_mu = cast_in<T>(alm_mu, size_t(N) * (8 + 2 * n_obs));
    }
    auto uu = cast_in<T>(u, size_t(N) * 2);
    auto xx = cast_in<T>(x, size_t(N + 1) * 4);
    std::vector<T> sc(N + 1);
    *J = double(s.total_cost(st.pb, uu.data(), xx.data(), sc.data()));
    cast_out(sc, step_cost);
    return 0;
}

template <typename T>
int do_cost_derivs(const Params* p, int N, int M, const double* wx, const double* wy,
                   const double* wyaw, double ref_velo, int n_obs, int obs_len, const double* obs,
                   const double* borders, const double* u, const double* x, const double* alm_mu,
                   double alm_rho, double* lx, double* lu, double* lxx, double* luu,
                   double* alm_mu_next) {
    ProblemStore<T> st(M, wx, wy, wyaw, ref_velo, n_obs, obs_len, obs, borders);
    Solver<T> s(*p, N);
    if (p->solve_type == 1) {
        s.alm_rho = T(alm_rho);
        s.alm_mu = cast_in<T>(alm_mu, size_t(N) * (8 + 2 * n_obs));
        s.alm_mu_next.assign(size_t(N) * (8 + 2 * n_obs), 0);
    }
    auto uu = cast_in<T>(u, size_t(N) * 2);
    auto xx = cast_in<T>(x, size_t(N + 1) * 4);
    s.cost_derivatives(st.pb, uu.data(), xx.data());
    cast_out(s.l_x, lx);
    cast_out(s.l_u, lu);
    cast_out(s.l_xx, lxx);
    cast_out(s.l_uu, luu);
    if (p->solve_type == 1) cast_out(s.alm_mu_next, alm_mu_next);
    return 0;
}

template <typename T>
int do_riccati(int N, const double* lx, const double* lu, const double* lxx, const double* luu,
               const double* A, const double* B, double lamb, double* d, double* K, double* dV,
               int32_t* status) {
    auto a = cast_in<T>(lx, size_t(N + 1) * 4);
    auto b = cast_in<T>(lu, size_t(N) * 2);
    auto c = cast_in<T>(lxx, size_t(N + 1) * 16);
    auto e = cast_in<T>(luu, size_t(N) * 4);
    auto f = cast_in<T>(A, size_t(N) * 16);
    auto g = cast_in<T>(B, size_t(N) * 8);
    BackwardResult<T> out;
    *status = Solver<T>::riccati(N, a.data(), b.data(), c.data(), e.data(), f.data(), g.data(),
                                 T(lamb), out);
    cast_out(out.d, d);
    cast_out(out.K, K);
    dV[0] = double(out.dV[0]);
    dV[1] = double(out.dV[1]);
    return 0;
}

template <typename T>
void solve_one(Solver<T>& s, int M, const double* wx, const double* wy, const double* wyaw,
               double ref_velo, int n_obs, int obs_len, const double* obs, const double* borders,
               const double* x0, double* u_out, double* x_out, double* K_out, double* d_out,
               double* J_out /*[2] init, final*/, double* step_cost_out, int32_t* info /*[4]*/,
               double* lamb_out, bool keep_trace) {
    ProblemStore<T> st(M, wx, wy, wyaw, ref_velo, n_obs, obs_len, obs, borders);
    T x0t[4] = {T(x0[0]), T(x0[1]), T(x0[2]), T(x0[3])};
    std::vector<T> u, x;
    s.solve(st.pb, x0t, u, x, keep_trace);
    cast_out(u, u_out);
    cast_out(x, x_out);
    cast_out(s.last_bw.K, K_out);
    cast_out(s.last_bw.d, d_out);
    if (J_out) {
        J_out[0] = double(s.J_init);
        J_out[1] = double(s.J_final);
    }
    if (step_cost_out) {
        std::vector<T> sc(s.N + 1);
        s.total_cost(st.pb, u.data(), x.data(), sc.data());
        cast_out(sc, step_cost_out);
    }
    if (info) {
        info[0] = s.status;
        info[1] = s.iters;
        info[2] = s.exit_reason;
        info[3] = 0;
    }
    if (lamb_out) *lamb_out = double(s.final_lamb);
}


template <typename T>
int do_propagate(const Params* p, const double* x, const double* u, double* out) {
    T xf[4] = {T(x[0]), T(x[1]), T(x[2]), T(x[3])};
    T uf[2] = {T(u[0]), T(u[1])};
    T of[4];
    kinematic_propagate<T>(xf, uf, T(p->dt), T(p->wheelbase), p->reference_point, of);
    for (int c = 0; c < 4; ++c) out[c] = double(of[c]);
    return 0;
}

template <typename T>
int do_dyn_derivs(const Params* p, int N, const double* u, const double* x, double* A, double* B) {
    Solver<T> s(*p, N);
    auto uu = cast_in<T>(u, size_t(N) * 2);
    auto xx = cast_in<T>(x, size_t(N + 1) * 4);
    std::vector<T> a, b;
    s.dyn_derivatives(uu.data(), xx.data(), a, b);
    cast_out(a, A);
    cast_out(b, B);
    return 0;
}

template <typename T>
int do_ref_match(int M, const double* wx, const double* wy, int rows, const double* x, int32_t* idx_out) {
    Params p{};
    std::vector<int> idx;
    Solver<T> s(p, rows - 1);
    auto a = cast_in<T>(wx, M), b = cast_in<T>(wy, M);
    auto xx = cast_in<T>(x, size_t(rows) * 4);
    Problem<T> pb;
    pb.wx = a.data();
    pb.wy = b.data();
    pb.M = M;
    s.ref_match(pb, xx.data(), rows, idx);
    for (int i = 0; i < rows; ++i) idx_out[i] = idx[i];
    return 0;
}

template <typename T>
int do_forward(const Params* p, int N, const double* u, const double* x, const double* d, const double* K,
               double alpha, double* new_u, double* new_x) {
    Solver<T> s(*p, N);
    auto a = cast_in<T>(u, size_t(N) * 2), b = cast_in<T>(x, size_t(N + 1) * 4);
    auto c = cast_in<T>(d, size_t(N) * 2), e = cast_in<T>(K, size_t(N) * 8);
    std::vector<T> nu(size_t(N) * 2), nx(size_t(N + 1) * 4);
    s.forward_pass(a.data(), b.data(), c.data(), e.data(), T(alpha), nu.data(), nx.data());
    cast_out(nu, new_u);
    cast_out(nx, new_x);
    return 0;
}

template <typename T>
int do_solver_solve(AnySolver* a, int M, const double* wx, const double* wy, const double* wyaw, double ref_velo,
                    int n_obs, int obs_len, const double* obs, const double* borders, const double* x0,
                    double* u_out, double* x_out, double* K_out, double* d_out, double* J_out,
                    double* step_cost_out, int32_t* info, double* lamb_out, double* trace, int trace_cap) {
    Solver<T>& s = static_cast<TypedSolver<T>*>(a)->s;
    solve_one(s, M, wx, wy, wyaw, ref_velo, n_obs, obs_len, obs, borders, x0, u_out, x_out, K_out, d_out, J_out,
              step_cost_out, info, lamb_out, trace != nullptr);
    if (trace) {
        int n = std::min<int>(int(s.trace.size()), trace_cap);
        for (int i = 0; i < n; ++i) {
            const IterTrace& t = s.trace[i];
            double* row = trace + size_t(i) * 6;
            row[0] = t.status;
            row[1] = t.alpha_index;
            row[2] = t.effective;
            row[3] = t.ori_cost;
            row[4] = t.new_cost;
            row[5] = t.lamb_after;
        }
        if (info) info[3] = n;
    }
    return 0;
}

template <typename T>
int do_batch_one(const Params& prm, int N, int M, const double* wx, const double* wy, const double* wyaw,
                 double ref_velo, int n_obs, int obs_len, const double* ob, const double* borders, const double* x0,
                 double* uo, double* xo, double* Ko, double* dout, double* J2, int32_t* info, int trace_cap,
                 std::vector<IterTrace>* tr_out) {
    Solver<T> s(prm, N);
    solve_one(s, M, wx, wy, wyaw, ref_velo, n_obs, obs_len, ob, borders, x0, uo, xo, Ko, dout, J2, nullptr, info,
              nullptr, trace_cap > 0);
    *tr_out = s.trace;
    return 0;
}

}  // namespace

extern "C" {

int oracle_sizeof_params() { return int(sizeof(Params)); }

int oracle_propagate(const Params* p, int dtype, const double* x, const double* u, double* out) {
    return ORACLE_DISPATCH(dtype, do_propagate, p, x, u, out);
}

int oracle_dyn_derivs(const Params* p, int dtype, int N, const double* u, const double* x, double* A,
                      double* B) {
    return ORACLE_DISPATCH(dtype, do_dyn_derivs, p, N, u, x, A, B);
}

int oracle_ref_match(int dtype, int M, const double* wx, const double* wy, int rows, const double* x,
                     int32_t* idx_out) {
    return ORACLE_DISPATCH(dtype, do_ref_match, M, wx, wy, rows, x, idx_out);
}

int oracle_total_cost(const Params* p, int dtype, int N, int M, const double* wx, const double* wy,
                      const double* wyaw, double ref_velo, int n_obs, int obs_len, const double* obs,
                      const double* borders, const double* u, const double* x, const double* alm_mu,
                      double alm_rho, double* J, double* step_cost) {
    return ORACLE_DISPATCH(dtype, do_total_cost, p, N, M, wx, wy, wyaw, ref_velo, n_obs, obs_len, obs, borders, u, x,
                           alm_mu, alm_rho, J, step_cost);
}

int oracle_cost_derivs(const Params* p, int dtype, int N, int M, const double* wx, const double* wy,
                       const double* wyaw, double ref_velo, int n_obs, int obs_len, const double* obs,
                       const double* borders, const double* u, const double* x, const double* alm_mu,
                       double alm_rho, double* lx, double* lu, double* lxx, double* luu,
                       double* alm_mu_next) {
    return ORACLE_DISPATCH(dtype, do_cost_derivs, p, N, M, wx, wy, wyaw, ref_velo, n_obs, obs_len, obs, borders, u, x,
                           alm_mu, alm_rho, lx, lu, lxx, luu, alm_mu_next);
}

int oracle_riccati(int dtype, int N, const double* lx, const double* lu, const double* lxx,
                   const double* luu, const double* A, const double* B, double lamb, double* d,
                   double* K, double* dV, int32_t* status) {
    return ORACLE_DISPATCH(dtype, do_riccati, N, lx, lu, lxx, luu, A, B, lamb, d, K, dV, status);
}

int oracle_forward(const Params* p, int dtype, int N, const double* u, const double* x,
                   const double* d, const double* K, double alpha, double* new_u, double* new_x) {
    return ORACLE_DISPATCH(dtype, do_forward, p, N, u, x, d, K, alpha, new_u, new_x);
}

// Stateful solver (warm start, cached derivatives, ALM multipliers persist
// between solve() calls exactly as in the reference object).
void* oracle_solver_create(const Params* p, int dtype, int N) {
    AnySolver* a = ORACLE_DISPATCH(dtype, make_solver, p, N);
    a->dtype = dtype;
    a->N = N;
    return a;
}

void oracle_solver_destroy(void* h) { delete static_cast<AnySolver*>(h); }

// info = {status, iters, exit_reason, n_trace}.  trace (optional) receives up to
// trace_cap rows of {status, alpha_index, effective, ori_cost, new_cost, lamb_after}.
int oracle_solver_solve(void* h, int M, const double* wx, const double* wy, const double* wyaw,
                        double ref_velo, int n_obs, int obs_len, const double* obs,
                        const double* borders, const double* x0, double* u_out, double* x_out,
                        double* K_out, double* d_out, double* J_out, double* step_cost_out,
                        int32_t* info, double* lamb_out, double* trace, int trace_cap) {
    auto* a = static_cast<AnySolver*>(h);
    if (obs_len < a->N + 1 && n_obs > 0) return -2;  // RoutingLine::operator[] would throw (utils.cpp:52-58)
    return ORACLE_DISPATCH(a->dtype, do_solver_solve, a, M, wx, wy, wyaw, ref_velo, n_obs, obs_len, obs, borders, x0,
                           u_out, x_out, K_out, d_out, J_out, step_cost_out, info, lamb_out, trace, trace_cap);
}

// Batch of independent first solves, one fresh solver per instance, instances
// fanned out over nthreads host threads (the reference is single-threaded per
// solve).  Template tables: params[T], waypoints concatenated with offsets
// wp_off[T+1].  Per-instance arrays in the reference layout.  Outputs may be NULL.
int oracle_solve_batch(int dtype, int N, int n_tmpl, const Params* params, const int32_t* wp_off,
                       const double* wx, const double* wy, const double* wyaw, int B,
                       const int32_t* tmpl, const double* x0, const double* ref_velo,
                       const double* borders, const int32_t* n_obs, int max_obs, int obs_len,
                       const double* obs, double* u_out, double* x_out, double* K_out, double* d_out,
                       double* J_out, int32_t* status_out, int32_t* iters_out, int32_t* exit_out,
                       int trace_cap, int32_t* tr_status, int32_t* tr_alpha, double* tr_cost, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    std::atomic<int> next(0);
    std::atomic<int> bad(0);
    auto copy_trace = [&](const std::vector<IterTrace>& tr, int b) {
        for (int i = 0; i < trace_cap && i < int(tr.size()); ++i) {
            if (tr_status) tr_status[size_t(b) * trace_cap + i] = tr[i].status;
            if (tr_alpha) tr_alpha[size_t(b) * trace_cap + i] = tr[i].alpha_index;
            if (tr_cost) tr_cost[size_t(b) * trace_cap + i] = tr[i].new_cost;
        }
    };
    auto worker = [&]() {
        for (;;) {
            int b = next.fetch_add(1);
            if (b >= B) break;
            int t = tmpl ? tmpl[b] : 0;
            if (t < 0 || t >= n_tmpl) {
                bad = 1;
                continue;
            }
            int off = wp_off[t], M = wp_off[t + 1] - wp_off[t];
            int32_t info[4];
            double J2[2];
            const double* ob = obs + size_t(b) * max_obs * obs_len * 3;
            double* uo = u_out ? u_out + size_t(b) * N * 2 : nullptr;
            double* xo = x_out ? x_out + size_t(b) * (N + 1) * 4 : nullptr;
            double* Ko = K_out ? K_out + size_t(b) * N * 8 : nullptr;
            double* dout = d_out ? d_out + size_t(b) * N * 2 : nullptr;
            std::vector<IterTrace> tr;
            ORACLE_DISPATCH(dtype, do_batch_one, params[t], N, M, wx + off, wy + off, wyaw + off, ref_velo[b], n_obs[b],
                            obs_len, ob, borders + size_t(b) * 2, x0 + size_t(b) * 4, uo, xo, Ko, dout, J2, info,
                            trace_cap, &tr);
            copy_trace(tr, b);
            if (J_out) {
                J_out[size_t(b) * 2 + 0] = J2[0];
                J_out[size_t(b) * 2 + 1] = J2[1];
            }
            if (status_out) status_out[b] = info[0];
            if (iters_out) iters_out[b] = info[1];
            if (exit_out) exit_out[b] = info[2];
        }
    };
    std::vector<std::thread> pool;
    for (int i = 1; i < nthreads; ++i) pool.emplace_back(worker);
    worker();
    for (auto& th : pool) th.join();
    return bad ? -1 : 0;
}

#ifdef CILQR_ORACLE_PMATH
// Test hook (tests/test_pmath_cpu.py): the portable transcendentals evaluated over an array.
// fn: 0 sin, 1 cos, 2 tan, 3 atan, 4 exp, 5 hypot(a, b).
int oracle_pmath_eval(int fn, int n, const double* a, const double* b, double* out) {
    for (int i = 0; i < n; ++i) {
        switch (fn) {
            case 0: out[i] = cilqr_pm::pm_sin(a[i]); break;
            case 1: out[i] = cilqr_pm::pm_cos(a[i]); break;
            case 2: out[i] = cilqr_pm::pm_tan(a[i]); break;
            case 3: out[i] = cilqr_pm::pm_atan(a[i]); break;
            case 4: out[i] = cilqr_pm::pm_exp(a[i]); break;
            case 5: out[i] = cilqr_pm::pm_hypot(a[i], b[i]); break;
            default: return -1;
        }
    }
    return 0;
}
#endif

}  // extern "C"
