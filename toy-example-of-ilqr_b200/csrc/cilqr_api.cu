// libcilqr_b200.so — C ABI (include/cilqr_b200.h) over the sm_100a kernels.
// One handle = one device, one stream, all buffers allocated at create time.
// There is deliberately no CPU path: without a Blackwell device every entry
// point fails with CILQR_ERR_NO_DEVICE.
#include "../../include/cilqr_b200.h"
#include "cilqr_kernels.cuh"

#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace cilqr;

namespace {

thread_local std::string g_err;

// NVTX range for the lifetime of the object (header-only NVTX3: a no-op unless a profiler is attached).  The
// ranges mark where the host issues each stage K0..K7 of a round, the whole solve, and the copies around it.
struct Nvtx {
    explicit Nvtx(const char* name) { nvtxRangePushA(name); }
    ~Nvtx() { nvtxRangePop(); }
    Nvtx(const Nvtx&) = delete;
    Nvtx& operator=(const Nvtx&) = delete;
};

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CK(expr)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(CILQR_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

#ifndef CILQR_COST_MINB_T
#define CILQR_COST_MINB_T 8
#endif
constexpr int kGridCap = 148 * 32;  // grid-stride kernels: at most 32 CTAs of 128 threads per SM (queued beyond residency)
// smallest batch that is still repacked: below it a round is a latency chain whatever the slots'
// order (measured: 8192 instances 20.7 -> 19.5 ms with repacks down to 4096, 4096 unchanged)
constexpr int kRepackMinBatch = 4096;
// largest handle that gets the spare gains / the doubled trial pool of the look-ahead rounds (the latency regime
// ends at prefetch_below = 16384 instances)
constexpr int kLookaheadMaxBatch = 16384;
// batches up to this size use them by default (measured: profiles/r02_lookahead.txt)
constexpr int kLookaheadDefault = 512;

// Type-erased part of a handle; the typed buffers live in Impl<T>.
struct Base {
    int dtype = 0, device = 0, max_batch = 0, N = 0, max_obs = 0, Bs = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cilqr_params_t params[CILQR_B200_MAX_TEMPLATES];
    bool tmpl_set[CILQR_B200_MAX_TEMPLATES] = {false};
    int wp_off[CILQR_B200_MAX_TEMPLATES] = {0}, wp_len[CILQR_B200_MAX_TEMPLATES] = {0};
    std::vector<double> h_wx, h_wy, h_wyaw;  // concatenated tables
    bool any_alm = false;
    int max_rounds = 0;
    int run_ahead = 3;
    int prefetch_below = 16384;  // batches up to this size use the latency-regime kernel variants (measured crossover ~16-18k)
    int bench_prefetch = -1;  // stage operator / roofline leg: -1 = the variant the solver uses at that batch
    unsigned scan_epoch = 0;  // tags the look-back words of one verdict launch (never 0, 30 bits)
    int pipeline = 1;  // latency regime: rollout and waypoint match as one two-stage kernel
    int staged = 1;    // latency-bound batches: backward pass fed by bulk async copies into shared memory
    int wide_step = 1; // bandwidth-bound rounds widen a line search step by step (2, 4, 8, 6 alphas) instead of all at once
    int repack = 1;    // survivors moved into a dense prefix whenever they are down to half of the slots in use
    int fused_backward = 1;  // bandwidth-bound rounds: the backward pass computes the control half of the records itself
    // look-ahead rounds (k_adopt): batches up to lookahead_below run next iteration's backward pass alongside the
    // line search's cost / verdict kernels, on a second stream
    int lookahead = 1;
    int lookahead_below = 0;  // set at create: min(kLookaheadMaxBatch, max_batch) when the spare buffers exist
    cudaStream_t stream_b = nullptr;
    cudaEvent_t ev_fork[4] = {nullptr, nullptr, nullptr, nullptr}, ev_join[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_mid[4] = {nullptr, nullptr, nullptr, nullptr};
    std::vector<cudaEvent_t> prof_ev_b;  // stage profile of the second stream
    std::vector<int> prof_stage_b;
    // optional in-step stage profile: CUDA events around every stage launch of one solve
    int profile = 0;
    std::vector<cudaEvent_t> prof_ev;
    std::vector<int> prof_stage;  // stage id that starts at event i (-1 = end marker)
    double stage_ms[6] = {0, 0, 0, 0, 0, 0};
    int stage_launches[6] = {0, 0, 0, 0, 0, 0};
    volatile int* h_ctl = nullptr;  // mapped pinned, read as one 64-bit word: rounds completed << 32 | instances active after it
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    double* stage = nullptr;  // device staging in host layout
    size_t stage_bytes = 0;
    int* istage = nullptr;    // device staging for int arrays [max_batch]
    float* flush = nullptr;
    size_t flush_n = 0;
    cilqr_counters_t counters{};
    int launches = 0;
    virtual ~Base() {}
};

template <typename T>
struct Impl : Base {
    Dev<T> D{};
    T* obs_plain = nullptr;   // [max_obs][N+1][4][Bs], the obstacle window of a plain solve
    T* obs_tracks = nullptr;  // [max_obs][track_cap][4][Bs], full tracks of the simulation
    int track_cap = 0;
    T* hist_x = nullptr;
    int* hist_iters = nullptr;
    int* hist_status = nullptr;
    int hist_cap = 0;
    // synthetic workloads: centre-line tables and template descriptors on the device
    double* syn_tab = nullptr;  // x | y | yaw | lon | nx | ny, each syn_cap samples
    size_t syn_cap = 0;
    int* syn_off = nullptr;
    int syn_lanes = 0;
    cilqr_synth_template_t* syn_tm = nullptr;  // [CILQR_B200_MAX_TEMPLATES]
    DevParams<T>* dP = nullptr;
    T* d_wp = nullptr;  // wx | wy | wyaw, each kMaxWaypoints? (sized on demand)
    size_t wp_cap = 0;
    std::vector<void*> allocs;
};

template <typename T>
DevParams<T> convert_params(const cilqr_params_t& p) {
    DevParams<T> d{};
    d.dt = T(p.dt);
    d.wheelbase = T(p.wheelbase);
    d.width = T(p.width);
    d.Q[0] = T(p.w_pos);
    d.Q[1] = T(p.w_pos);
    d.Q[2] = T(p.w_vel);
    d.Q[3] = T(p.w_yaw);
    d.R[0] = T(p.w_acc);
    d.R[1] = T(p.w_stl);
    d.obs_q1 = T(p.obstacle_exp_q1);
    d.obs_q2 = T(p.obstacle_exp_q2);
    d.st_q1 = T(p.state_exp_q1);
    d.st_q2 = T(p.state_exp_q2);
    d.acc_max = T(p.acc_max);
    d.acc_min = T(p.acc_min);
    d.stl_lim = T(p.stl_lim);
    d.velo_max = T(p.velo_max);
    d.velo_min = T(p.velo_min);
    // src/utils.cpp:387-393 with obs_attr = {width, length, d_safe} and radius = width/2
    T radius = T(0.5) * T(p.width);
    T a = T(0.5) * T(p.length) + T(p.d_safe) * 6 + radius;
    T b = T(0.5) * T(p.width) + T(p.d_safe) + radius;
#ifdef CILQR_PARITY
    d.ell_a2 = a * a;  // parity build: the squares themselves, divided by as in the reference (ellipse_margin)
    d.ell_b2 = b * b;
#else
    d.ell_a2 = T(1) / (a * a);
    d.ell_b2 = T(1) / (b * b);
#endif
    d.alm_rho_init = T(p.alm_rho_init);
    d.alm_gamma = T(p.alm_gamma);
    d.max_rho = T(p.max_rho);
    d.max_mu = T(p.max_mu);
    d.init_lamb = T(p.init_lamb);
    d.lamb_decay = T(p.lamb_decay);
    d.lamb_amplify = T(p.lamb_amplify);
    d.max_lamb = T(p.max_lamb);
    d.conv_thr = T(p.convergence_threshold);
    d.accept_thr = T(p.accept_step_threshold);
    d.max_iter = p.max_iter;
    d.solve_type = p.solve_type;
    d.ref_point = p.reference_point;
    d.use_last = p.use_last_solution;
    return d;
}

int check_params(const cilqr_params_t* p) {
    if (!p) return fail(CILQR_ERR_INVALID, "params is NULL");
    if (!(p->dt > 0)) return fail(CILQR_ERR_INVALID, "delta_t must be positive");
    if (!(p->wheelbase > 0)) return fail(CILQR_ERR_INVALID, "vehicle/wheelbase must be positive");
    if (p->max_iter < 0 || p->max_iter > 100000) return fail(CILQR_ERR_INVALID, "iteration/max_iter out of range");
    if (p->solve_type != 0 && p->solve_type != 1) return fail(CILQR_ERR_INVALID, "solve_type must be 0 (barrier) or 1 (alm)");
    if (p->reference_point != 0 && p->reference_point != 1)
        return fail(CILQR_ERR_INVALID, "reference_point must be 0 (rear_center) or 1 (gravity_center)");
    return 0;
}

template <typename T, typename P>
int dalloc(Impl<T>* h, P** p, size_t count) {
    void* q = nullptr;
    size_t bytes = std::max<size_t>(count, 1) * sizeof(P);
    CK(cudaMalloc(&q, bytes));
    CK(cudaMemsetAsync(q, 0, bytes, h->stream));
    h->allocs.push_back(q);
    *p = static_cast<P*>(q);
    return 0;
}

template <typename T>
int alloc_alm(Impl<T>* h) {
    if (h->D.mu) return 0;
    size_t n = size_t(h->N) * h->D.alm_cols * h->Bs;
    int rc;
    if ((rc = dalloc(h, &h->D.mu, n))) return rc;
    if ((rc = dalloc(h, &h->D.mu_next, n))) return rc;
    if ((rc = dalloc(h, &h->D.rho, h->Bs))) return rc;
    return 0;
}

template <typename T>
int upload_templates(Impl<T>* h) {
    // waypoints
    size_t total = h->h_wx.size();
    if (total > h->wp_cap) {
        size_t cap = std::max<size_t>(total, 4096);
        T* q = nullptr;
        CK(cudaMalloc(&q, cap * 5 * sizeof(T)));
        h->allocs.push_back(q);
        h->d_wp = q;
        h->wp_cap = cap;
    }
    if (total) {
        std::vector<T> tmp(h->wp_cap * 3, T(0));
        for (size_t i = 0; i < total; ++i) {
            tmp[i] = T(h->h_wx[i]);
            tmp[h->wp_cap + i] = T(h->h_wy[i]);
            tmp[2 * h->wp_cap + i] = T(h->h_wyaw[i]);
        }
        CK(cudaMemcpyAsync(h->d_wp, tmp.data(), tmp.size() * sizeof(T), cudaMemcpyHostToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    h->D.wx = h->d_wp;
    h->D.wy = h->d_wp + h->wp_cap;
    h->D.wyaw = h->d_wp + 2 * h->wp_cap;
    h->D.wsin = h->d_wp + 3 * h->wp_cap;
    h->D.wcos = h->d_wp + 4 * h->wp_cap;
    if (total) {
        k_wp_sincos<T><<<(unsigned(total) + 127) / 128, 128, 0, h->stream>>>(
            h->d_wp + 2 * h->wp_cap, h->d_wp + 3 * h->wp_cap, h->d_wp + 4 * h->wp_cap, int(total));
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(h->stream));
    }
    std::vector<DevParams<T>> hp(CILQR_B200_MAX_TEMPLATES);
    h->any_alm = false;
    int max_iter = 0;
    for (int t = 0; t < CILQR_B200_MAX_TEMPLATES; ++t) {
        const cilqr_params_t& src = h->tmpl_set[t] ? h->params[t] : h->params[0];
        hp[t] = convert_params<T>(src);
        hp[t].wp_off = h->wp_off[t];
        hp[t].wp_len = h->wp_len[t];
        if (h->tmpl_set[t]) {
            h->any_alm |= src.solve_type == 1;
            max_iter = std::max(max_iter, src.max_iter);
        }
    }
    CK(cudaMemcpyAsync(h->dP, hp.data(), hp.size() * sizeof(DevParams<T>), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (h->any_alm) {
        int rc = alloc_alm(h);
        if (rc) return rc;
    }
    h->max_rounds = max_iter * kNumAlphas + 8;
    return 0;
}

template <typename T>
int create_impl(const cilqr_params_t* params, int device, int max_batch, int N, int max_obs, cilqr_handle_t** out) {
    auto* h = new Impl<T>();
    h->dtype = std::is_same<T, float>::value ? CILQR_F32 : CILQR_F64;
    h->device = device;
    h->max_batch = max_batch;
    h->N = N;
    h->max_obs = max_obs;
    h->Bs = (max_batch + 127) / 128 * 128;
    int rc = 0;
    auto body = [&]() -> int {
        CK(cudaSetDevice(device));
        CK(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
        h->stream = h->own_stream;
        CK(cudaEventCreate(&h->t0));
        CK(cudaEventCreate(&h->t1));
        Dev<T>& D = h->D;
        D.N = N;
        D.Bs = h->Bs;
        // trial pool: room for every instance's single trial plus wide searches of the stragglers
        D.Vs = int(std::min<size_t>(size_t(h->Bs) * 4, size_t(h->Bs) + (size_t(1) << 21)));
        const bool la = max_batch <= kLookaheadMaxBatch;
        if (la) D.Vs *= 2;  // a look-ahead solve alternates between two halves of the pool
        D.spec = 0;
        D.pool_base = 0;
        D.pool_cap = D.Vs;
        D.round_id = 0;
        D.max_obs = max_obs;
        D.alm_cols = 8 + 2 * max_obs;
        D.wide_mode = 1;
        D.trace_cap = 0;
        const size_t Bs = h->Bs, Vs = D.Vs;
        int r;
        {
            int* hc = nullptr;
            CK(cudaHostAlloc(&hc, 4 * sizeof(int), cudaHostAllocMapped));
            hc[0] = hc[1] = hc[2] = hc[3] = 0;
            h->h_ctl = hc;
            int* dc = nullptr;
            CK(cudaHostGetDevicePointer(&dc, hc, 0));
            D.h_ctl = dc;
        }
        if ((r = dalloc(h, &D.ctl, CTL_WORDS))) return r;
        if ((r = dalloc(h, &h->dP, CILQR_B200_MAX_TEMPLATES))) return r;
        D.P = h->dP;
        if ((r = dalloc(h, &D.ref_velo, Bs))) return r;
        if ((r = dalloc(h, &D.borders, 2 * Bs))) return r;
        if ((r = dalloc(h, &D.tmpl, Bs))) return r;
        if ((r = dalloc(h, &D.n_obs, Bs))) return r;
        if ((r = dalloc(h, &h->obs_plain, size_t(max_obs) * (N + 1) * 4 * Bs))) return r;
        D.obs = h->obs_plain;
        D.obs_len = N + 1;
        D.obs_off = 0;
        if ((r = dalloc(h, &D.x0, 4 * Bs))) return r;
        if ((r = dalloc(h, &D.X, size_t(N + 1) * 4 * Bs))) return r;
        if ((r = dalloc(h, &D.U, size_t(N) * 2 * Bs))) return r;
        if ((r = dalloc(h, &D.ridx, size_t(N + 1) * Bs))) return r;
        if ((r = dalloc(h, &D.sc, size_t(N + 1) * kScPlanes * Bs))) return r;
        if ((r = dalloc(h, &D.Xt, size_t(N + 1) * 4 * Vs))) return r;
        if ((r = dalloc(h, &D.Ut, size_t(N) * 2 * Vs))) return r;
        if ((r = dalloc(h, &D.ridx_t, size_t(N + 1) * Vs))) return r;
        if ((r = dalloc(h, &D.sc_t, size_t(N + 1) * kScPlanes * Vs))) return r;
        if ((r = dalloc(h, &D.t_inst, Vs))) return r;
        if ((r = dalloc(h, &D.t_aidx, Vs))) return r;
        if ((r = dalloc(h, &D.t_done, Vs))) return r;
        if ((r = dalloc(h, &D.J_t, Vs))) return r;
        if ((r = dalloc(h, &D.t_first, Bs))) return r;
        if ((r = dalloc(h, &D.t_count, Bs))) return r;
        if ((r = dalloc(h, &D.commit_src, Bs))) return r;
        if ((r = dalloc(h, &D.wide, Bs))) return r;
        if ((r = dalloc(h, &D.act, 2 * Bs))) return r;
        if ((r = dalloc(h, &D.scan_state, Bs / 128 + 1))) return r;
        if ((r = dalloc(h, &D.swap_src, Bs + 8))) return r;
        if ((r = dalloc(h, &D.swap_dst, Bs + 8))) return r;
        if ((r = dalloc(h, &D.rec, size_t(N + 1) * kRecFields * Bs))) return r;
        if ((r = dalloc(h, &D.Kg, size_t(N) * 8 * Bs * (la ? 2 : 1)))) return r;
        if ((r = dalloc(h, &D.dg, size_t(N) * 2 * Bs * (la ? 2 : 1)))) return r;
        if ((r = dalloc(h, &D.dV, 2 * Bs * (la ? 2 : 1)))) return r;
        if (la) {
            if ((r = dalloc(h, &D.gsel, Bs))) return r;
            if ((r = dalloc(h, &D.job_round, Bs))) return r;
            if ((r = dalloc(h, &D.job_src, Bs))) return r;
            if ((r = dalloc(h, &D.job_lamb, Bs))) return r;
            if ((r = dalloc(h, &D.job_ok, Bs))) return r;
            if ((r = dalloc(h, &D.cur_src, Bs))) return r;
            if ((r = dalloc(h, &D.t_job, Vs))) return r;
            if ((r = dalloc(h, &D.jlamb_t, Vs))) return r;
            if ((r = dalloc(h, &D.jok_t, Vs))) return r;
            if ((r = dalloc(h, &D.rec_t, size_t(N + 1) * kRecFields * Vs))) return r;
            if ((r = dalloc(h, &D.Kg_t, size_t(N) * 8 * Vs))) return r;
            if ((r = dalloc(h, &D.dg_t, size_t(N) * 2 * Vs))) return r;
            if ((r = dalloc(h, &D.dV_t, 2 * Vs))) return r;
            CK(cudaStreamCreateWithFlags(&h->stream_b, cudaStreamNonBlocking));
            for (int i = 0; i < 4; ++i) {
                CK(cudaEventCreateWithFlags(&h->ev_fork[i], cudaEventDisableTiming));
                CK(cudaEventCreateWithFlags(&h->ev_join[i], cudaEventDisableTiming));
                CK(cudaEventCreateWithFlags(&h->ev_mid[i], cudaEventDisableTiming));
            }
            h->lookahead_below = kLookaheadMaxBatch;
        }
        if ((r = dalloc(h, &D.lamb, Bs))) return r;
        if ((r = dalloc(h, &D.J_cur, Bs))) return r;
        if ((r = dalloc(h, &D.J_init, Bs))) return r;
        if ((r = dalloc(h, &D.alpha, Bs))) return r;
        if ((r = dalloc(h, &D.status, Bs))) return r;
        if ((r = dalloc(h, &D.phase, Bs))) return r;
        if ((r = dalloc(h, &D.aidx, Bs))) return r;
        if ((r = dalloc(h, &D.iters, Bs))) return r;
        if ((r = dalloc(h, &D.exit_reason, Bs))) return r;
        if ((r = dalloc(h, &D.rec_valid, Bs))) return r;
        if ((r = dalloc(h, &D.last_u, size_t(N) * 2 * Bs))) return r;
        if ((r = dalloc(h, &D.first, Bs))) return r;
        if ((r = dalloc(h, &h->istage, Bs))) return r;
        // staging: one chunk of trajectories in host layout; 64 MiB or one trajectory's largest array
        size_t per_traj = std::max<size_t>(size_t(N + 1) * 16, size_t(max_obs) * (N + 1) * 3) * sizeof(double);
        h->stage_bytes = std::max<size_t>(size_t(64) << 20, per_traj * 4);
        if ((r = dalloc(h, reinterpret_cast<char**>(&h->stage), h->stage_bytes))) return r;
        h->params[0] = *params;
        h->tmpl_set[0] = true;
        if ((r = upload_templates(h))) return r;
        CK(cudaMemsetAsync(D.first, 0, Bs * sizeof(int), h->stream));
        return 0;
    };
    rc = body();
    if (rc == 0) {
        // first = 1 everywhere
        std::vector<int> ones(h->Bs, 1);
        cudaError_t e = cudaMemcpyAsync(h->D.first, ones.data(), ones.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) rc = fail(CILQR_ERR_CUDA, "initialising warm-start flags: %s", cudaGetErrorString(e));
    }
    if (rc) {
        for (void* p : h->allocs) cudaFree(p);
        delete h;
        return rc;
    }
    *out = reinterpret_cast<cilqr_handle_t*>(static_cast<Base*>(h));
    return 0;
}

inline Base* base(cilqr_handle_t* h) { return reinterpret_cast<Base*>(h); }

#define DISPATCH(h, fn, ...)                                                        \
    (base(h)->dtype == CILQR_F64 ? fn(static_cast<Impl<double>*>(base(h)), ##__VA_ARGS__) \
                                 : fn(static_cast<Impl<float>*>(base(h)), ##__VA_ARGS__))

inline dim3 grid1(int B) { return dim3((B + 127) / 128); }
inline dim3 grid2(int B, int rows) { return dim3((B + 127) / 128, rows); }
// grid-stride launches: enough CTAs for `count` items, capped at a multiple of the 148 SMs
inline dim3 gs1(int count) { return dim3(std::max(1, std::min((count + 127) / 128, kGridCap))); }
inline dim3 gs2(int count, int rows) { return dim3(std::max(1, std::min((count + 127) / 128, kGridCap)), rows); }
// Which feed of the backward recursion a launch over n trajectories uses (all return the same bits):
//   2  staged through shared memory by bulk async copies — small batches (whole-batch tiles)
//   1  next step's record prefetched into registers — latency-bound lists, and bandwidth-bound launches
//      large enough to fill the GPU at 3 CTAs / SM (fp64 measured on the tiled layout: 6815 against
//      6390 GB/s at 262144, 5349 against 5976 at 65536; fp32: ahead at every size, 5466 against 4233
//      at 65536, 6365 against 6112 at 262144)
//   0  plain streaming — fp64 in between
template <typename T>
int backward_variant(const Base* h, int n, int B, bool lat) {
    if (lat) return (h->staged && B <= h->prefetch_below) ? 2 : 1;
    if (sizeof(T) == 4) return 1;
    return n >= 196608 ? 1 : 0;
}
// k_backward in the bandwidth regime: 64-thread CTAs.  The kernel is one resident wave (<= 4 CTAs of 128 threads per
// SM at its register count), so its time is that of the fullest SM: 65536 trajectories in 128-thread CTAs are 512
// CTAs = 3.46 per SM, i.e. some SMs stream 4 CTAs while the others idle after 3 (measured 0.84-0.91 of the copy
// bandwidth at that batch against 0.97-1.04 at 262144); at 64 threads the imbalance is one CTA in seven.
static const int kBwThreads = [] { const char* e = getenv("CILQR_BW_THREADS"); const int v = e ? atoi(e) : 0; return (v == 32 || v == 64 || v == 128) ? v : 64; }();
inline dim3 bw_grid(int n) { return dim3(std::max(1, std::min((n + kBwThreads - 1) / kBwThreads, 2 * kGridCap))); }
// k_backward_staged: one warp per tile of 32 instances, at most 16 warps per SM
inline dim3 staged_grid(int B) { return dim3(std::max(1, std::min((B + 31) / 32, 148 * 16))); }
// The staged backward pass over `count` trajectories (tiles of 32): one warp per tile.
template <typename T>
inline void launch_staged_backward(Base* h, cudaStream_t st, const Dev<T>& D, int count, int B, int solver) {
    k_backward_staged<T><<<staged_grid(count), 32, 0, st>>>(D, B, solver);
    h->launches++;
}
// step-parallel stages (k_cost, k_derivs): x = step, y = blocks of trajectories
inline dim3 gk(int count, int rows) { return dim3(rows, std::max(1, std::min((count + 127) / 128, kGridCap))); }

template <typename... KArgs, typename... Args>
inline void launch_kernel(Base* h, void (*kernel)(KArgs...), dim3 grid, dim3 block, Args&&... args) {
    kernel<<<grid, block, 0, h->stream>>>(std::forward<Args>(args)...);
    h->launches++;
}
#define LAUNCH(h, kernel, grid, block, ...) launch_kernel(h, kernel, grid, block, __VA_ARGS__)
template <typename... KArgs, typename... Args>
inline void launch_kernel_on(Base* h, cudaStream_t st, void (*kernel)(KArgs...), dim3 grid, dim3 block, Args&&... args) {
    kernel<<<grid, block, 0, st>>>(std::forward<Args>(args)...);
    h->launches++;
}
#define LAUNCH_ON(h, st, kernel, grid, block, ...) launch_kernel_on(h, st, kernel, grid, block, __VA_ARGS__)
// cost / derivative kernels: the ALM paths are compiled in only when a template asks for them
#define LAUNCH_COST(h, minb, grid, ...)                                     \
    do {                                                                    \
        if ((h)->any_alm) LAUNCH(h, (k_cost<T, minb, true>), grid, 128, __VA_ARGS__);  \
        else LAUNCH(h, (k_cost<T, minb, false>), grid, 128, __VA_ARGS__);   \
    } while (0)
#define LAUNCH_DERIVS(h, part, grid, ...)                                   \
    do {                                                                    \
        if ((h)->any_alm) LAUNCH(h, (k_derivs<T, part, true>), grid, 128, __VA_ARGS__, 0);  \
        else LAUNCH(h, (k_derivs<T, part, false>), grid, 128, __VA_ARGS__, 0); \
    } while (0)

// host layout (double [B][E_src rows]) -> device SoA, chunked through the staging buffer.
// rows_src/rows_dst/inner: when the source has more rows per block than the destination keeps
// (obstacle tracks longer than N+1), only the first rows_dst rows of each block are kept.
template <typename T>
__global__ void k_pack_rows(const double* __restrict__ src, T* __restrict__ dst, int B, int blocks, int rows_src,
                            int rows_dst, int inner, int inner_dst, size_t Bs, int b_off) {
    __shared__ double tile[32][33];
    const int E = blocks * rows_src * inner;
    int e0 = blockIdx.x * 32, b0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {
        int b = b0 + r, e = e0 + threadIdx.x;
        if (b < B && e < E) tile[r][threadIdx.x] = src[size_t(b) * E + e];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
        int e = e0 + r, b = b0 + threadIdx.x;
        if (b < B && e < E) {
            int c = e % inner;
            int row = (e / inner) % rows_src;
            int blk = e / (inner * rows_src);
            if (row < rows_dst)
                dst[(size_t(blk) * rows_dst + row) * inner_dst * Bs + size_t(c) * Bs + b_off + b] = T(tile[threadIdx.x][r]);
        }
    }
}

template <typename T>
int pack_to_device(Impl<T>* h, const double* src, T* dst, int B, int blocks, int rows_src, int rows_dst, int inner,
                   int inner_dst = 0) {
    if (inner_dst == 0) inner_dst = inner;
    if (!src) return fail(CILQR_ERR_INVALID, "NULL input array");
    const size_t E = size_t(blocks) * rows_src * inner;
    if (E == 0 || B == 0) return 0;
    size_t per = E * sizeof(double);
    int chunk = int(std::min<size_t>(size_t(B), std::max<size_t>(h->stage_bytes / per, 1)));
    chunk = std::min(chunk, 65535 * 32);  // gridDim.y of the transpose kernel
    if (per > h->stage_bytes) return fail(CILQR_ERR_INVALID, "one trajectory's array (%zu bytes) exceeds the staging buffer", per);
    for (int b0 = 0; b0 < B; b0 += chunk) {
        int nb = std::min(chunk, B - b0);
        CK(cudaMemcpyAsync(h->stage, src + size_t(b0) * E, size_t(nb) * per, cudaMemcpyHostToDevice, h->stream));
        dim3 g((unsigned(E) + 31) / 32, (nb + 31) / 32), blk(32, 8);
        k_pack_rows<T><<<g, blk, 0, h->stream>>>(h->stage, dst, nb, blocks, rows_src, rows_dst, inner, inner_dst, h->Bs, b0);
        h->launches++;
    }
    CK(cudaGetLastError());
    return 0;
}

template <typename T>
__global__ void k_unpack_rows(const T* __restrict__ src, double* __restrict__ dst, int B, int E, size_t Bs,
                              int b_off) {
    __shared__ double tile[32][33];
    int e0 = blockIdx.x * 32, b0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {
        int e = e0 + r, b = b0 + threadIdx.x;
        if (b < B && e < E) tile[r][threadIdx.x] = double(src[size_t(e) * Bs + b_off + b]);
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
        int b = b0 + r, e = e0 + threadIdx.x;
        if (b < B && e < E) dst[size_t(b) * E + e] = tile[threadIdx.x][r];
    }
}

template <typename T>
int unpack_to_host(Impl<T>* h, const T* src, double* dst, int B, int E, size_t stride = 0) {
    if (!dst || E == 0 || B == 0) return 0;
    if (stride == 0) stride = h->Bs;
    size_t per = size_t(E) * sizeof(double);
    int chunk = int(std::min<size_t>(size_t(B), std::max<size_t>(h->stage_bytes / per, 1)));
    chunk = std::min(chunk, 65535 * 32);  // gridDim.y of the transpose kernel
    for (int b0 = 0; b0 < B; b0 += chunk) {
        int nb = std::min(chunk, B - b0);
        dim3 g((E + 31) / 32, (nb + 31) / 32), blk(32, 8);
        k_unpack_rows<T><<<g, blk, 0, h->stream>>>(src, h->stage, nb, E, stride, b0);
        h->launches++;
        CK(cudaMemcpyAsync(dst + size_t(b0) * E, h->stage, size_t(nb) * per, cudaMemcpyDeviceToHost, h->stream));
    }
    CK(cudaGetLastError());
    return 0;
}

template <typename T>
int upload_ints(Impl<T>* h, const int32_t* src, int* dst, int B, int fill) {
    if (src) {
        CK(cudaMemcpyAsync(dst, src, size_t(B) * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    } else {
        k_pack_int<<<grid1(B), 128, 0, h->stream>>>(nullptr, dst, B, fill);
        h->launches++;
    }
    return 0;
}

template <typename T>
int download_ints(Impl<T>* h, const int* src, int32_t* dst, int B) {
    if (!dst) return 0;
    CK(cudaMemcpyAsync(dst, src, size_t(B) * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    return 0;
}

int check_batch(Base* h, int B) {
    if (!h) return fail(CILQR_ERR_INVALID, "handle is NULL");
    if (B < 0 || B > h->max_batch) return fail(CILQR_ERR_INVALID, "batch %d outside [0, max_batch=%d]", B, h->max_batch);
    return 0;
}

int check_tmpl_nobs(Base* h, int B, const int32_t* tmpl, const int32_t* n_obs, int obs_len) {
    if (tmpl)
        for (int b = 0; b < B; ++b)
            if (tmpl[b] < 0 || tmpl[b] >= CILQR_B200_MAX_TEMPLATES || !h->tmpl_set[tmpl[b]])
                return fail(CILQR_ERR_INVALID, "instance %d uses template %d, which was never set", b, tmpl[b]);
    for (int t = 0; t < CILQR_B200_MAX_TEMPLATES; ++t)
        if (h->tmpl_set[t] && h->wp_len[t] <= 0 && (!tmpl || std::find(tmpl, tmpl + B, t) != tmpl + B))
            return fail(CILQR_ERR_INVALID, "template %d has no reference line (cilqr_b200_set_template)", t);
    bool any = false;
    if (n_obs)
        for (int b = 0; b < B; ++b) {
            if (n_obs[b] < 0 || n_obs[b] > h->max_obs)
                return fail(CILQR_ERR_INVALID, "instance %d has n_obs=%d outside [0, max_obs=%d]", b, n_obs[b], h->max_obs);
            any |= n_obs[b] > 0;
        }
    if (any && obs_len < h->N + 1)
        return fail(CILQR_ERR_RANGE, "obstacle tracks hold %d samples, need N+1=%d (RoutingLine index out of range)", obs_len, h->N + 1);
    return 0;
}

template <typename T>
int upload_problem_data(Impl<T>* h, int B, const double* ref_velo, const double* borders, const int32_t* tmpl,
                        const int32_t* n_obs, const double* obs, int obs_len) {
    int rc;
    h->D.obs = h->obs_plain;
    h->D.obs_len = h->N + 1;
    h->D.obs_off = 0;
    if ((rc = check_tmpl_nobs(h, B, tmpl, n_obs, obs_len))) return rc;
    if ((rc = upload_ints(h, tmpl, h->D.tmpl, B, 0))) return rc;
    if ((rc = upload_ints(h, n_obs, h->D.n_obs, B, 0))) return rc;
    if ((rc = pack_to_device(h, ref_velo, h->D.ref_velo, B, 1, 1, 1, 1))) return rc;
    if ((rc = pack_to_device(h, borders, h->D.borders, B, 1, 2, 2, 1))) return rc;
    bool any = false;
    if (n_obs)
        for (int b = 0; b < B && !any; ++b) any = n_obs[b] > 0;
    if (any && h->max_obs > 0) {
        if ((rc = pack_to_device(h, obs, h->obs_plain, B, h->max_obs, obs_len, h->N + 1, 3, 4))) return rc;
        LAUNCH(h, k_obs_sincos<T>, gs2(B, h->max_obs * (h->N + 1)), 128, h->obs_plain, B, size_t(h->Bs));
    }
    return 0;
}

template <typename T>
int do_upload(Impl<T>* h, int B, const double* x0, const double* ref_velo, const double* borders,
              const int32_t* tmpl, const int32_t* n_obs, const double* obs, int obs_len) {
    int rc;
    CK(cudaSetDevice(h->device));
    if ((rc = upload_problem_data(h, B, ref_velo, borders, tmpl, n_obs, obs, obs_len))) return rc;
    if ((rc = pack_to_device(h, x0, h->D.x0, B, 1, 4, 4, 1))) return rc;
    return 0;
}

// waypoint match + per-step costs of a set of trajectories (0: current, 1: the trial pool)
// stage profile: an event in front of stage `id` (0 derivs, 1 backward, 2 forward, 3 ref_match, 4 cost,
// 5 decide; -1 closes the round)
inline void mark_stage(Base* h, int id) {
    if (!h->profile) return;
    size_t i = h->prof_stage.size();
    if (i >= h->prof_ev.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        h->prof_ev.push_back(e);
    }
    cudaEventRecord(h->prof_ev[i], h->stream);
    h->prof_stage.push_back(id);
}

// the same on the second stream of a look-ahead solve (its own event list)
inline void mark_stage_b(Base* h, int id) {
    if (!h->profile) return;
    size_t i = h->prof_stage_b.size();
    if (i >= h->prof_ev_b.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        h->prof_ev_b.push_back(e);
    }
    cudaEventRecord(h->prof_ev_b[i], h->stream_b);
    h->prof_stage_b.push_back(id);
}

// `count`: an upper bound of the trajectories the launch has to cover (all B instances, or the slots of
// the trial pool that can be in use this round); `lat`: latency-regime kernel variants.
template <typename T>
void launch_cost(Impl<T>* h, int B, int trial, int count, bool lat, bool matched = false) {
    if (trial) mark_stage(h, 3);
    // waypoint scan window: 16 lanes per trajectory while the batch is latency-bound (one probe
    // usually covers a step's advance), 8 lanes in the throughput regime
    if (matched) {
        // k_rollout_match already wrote the matches of the trial pool
    } else if (lat) {
        LAUNCH(h, (k_ref_match<T, 16>), gs1(count * 16), 128, h->D, B, trial);
    } else {
        LAUNCH(h, (k_ref_match<T, 8>), gs1(count * 8), 128, h->D, B, trial);
    }
    if (trial) mark_stage(h, 4);
    if (lat) {
        LAUNCH_COST(h, 4, gk(count, h->N + 1), h->D, B, trial);
    } else {
        LAUNCH_COST(h, CILQR_COST_MINB_T, gk(count, h->N + 1), h->D, B, trial ? 2 : 0);
        if (trial) LAUNCH(h, k_sum_trials<T>, gs1(count), 128, h->D, B);
    }
}

// Exchange the slot pairs of repack `level` in every per-instance array (see k_plan_repack): moves the
// survivors into the prefix, and — applied a second time — back.  m_bound >= the number of pairs.
template <typename T>
void swap_instances(Impl<T>* h, int level, int offset, int m_bound) {
    const Dev<T>& D = h->D;
    const size_t Bs = D.Bs;
    const int N = D.N;
    const int* src = D.swap_src + offset;
    const int* dst = D.swap_dst + offset;
    const int* m = D.ctl + CTL_NSWAP + level;
    const int gx = std::max(1, std::min((m_bound + 127) / 128, kGridCap));
    auto rows_t = [&](T* p, int rows) {
        if (p) LAUNCH(h, k_swap_rows<T>, dim3(gx, std::min(rows, 1024)), 128, p, Bs, rows, src, dst, m);
    };
    auto rows_i = [&](int* p, int rows) {
        if (p) LAUNCH(h, k_swap_rows<int>, dim3(gx, std::min(rows, 1024)), 128, p, Bs, rows, src, dst, m);
    };
    // problem data
    rows_t(D.ref_velo, 1);
    rows_t(D.borders, 2);
    rows_i(D.tmpl, 1);
    rows_i(D.n_obs, 1);
    if (D.max_obs > 0) rows_t(D.obs, D.max_obs * D.obs_len * 4);
    rows_t(D.x0, 4);
    // trajectory, records, gains
    rows_t(D.X, (N + 1) * 4);
    rows_t(D.U, N * 2);
    rows_i(D.ridx, N + 1);
    rows_t(D.sc, (N + 1) * kScPlanes);
    LAUNCH(h, k_swap_records<T>, dim3(gx, std::min((N + 1) * kRecFields, 1024)), 128, D, src, dst, m);
    rows_t(D.Kg, N * 8);
    rows_t(D.dg, N * 2);
    rows_t(D.dV, 2);
    // solver state
    rows_t(D.lamb, 1);
    rows_t(D.J_cur, 1);
    rows_t(D.J_init, 1);
    rows_t(D.alpha, 1);
    rows_i(D.status, 1);
    rows_i(D.phase, 1);
    rows_i(D.aidx, 1);
    rows_i(D.iters, 1);
    rows_i(D.exit_reason, 1);
    rows_i(D.rec_valid, 1);
    rows_i(D.wide, 1);
    rows_i(D.commit_src, 1);
    rows_i(D.t_first, 1);
    rows_i(D.t_count, 1);
    rows_t(D.last_u, N * 2);
    rows_i(D.first, 1);
    if (D.mu) {
        rows_t(D.mu, N * D.alm_cols);
        rows_t(D.mu_next, N * D.alm_cols);
        rows_t(D.rho, 1);
    }
    if (D.trace_cap > 0) {
        rows_i(D.tr_status, D.trace_cap);
        rows_i(D.tr_alpha, D.trace_cap);
        rows_t(D.tr_cost, D.trace_cap);
    }
}

template <typename T>
int do_solve_resident(Impl<T>* h, int Bfull) {
    CK(cudaSetDevice(h->device));
    if (Bfull == 0) return 0;
    Nvtx solve_range("cilqr solve_resident");
    int B = Bfull;  // slots in use: shrinks when the survivors are repacked into a prefix
    const int N = h->N;
    h->launches = 0;
    CK(cudaStreamSynchronize(h->stream));
    if (h->stream_b) CK(cudaStreamSynchronize(h->stream_b));  // (nothing of a solve that ended in an error may still be running)
    // progress words written by the verdict kernel: [0] rounds completed << 32 | instances running,
    // [1] rounds completed << 32 | entries on the next round's work list
    volatile unsigned long long* progress = reinterpret_cast<volatile unsigned long long*>(h->h_ctl);
    progress[0] = static_cast<unsigned long long>(unsigned(B));
    progress[1] = static_cast<unsigned long long>(unsigned(B));
    CK(cudaMemsetAsync(h->D.ctl, 0, CTL_WORDS * sizeof(int), h->stream));
    {
        Nvtx r("K0 init trajectory + K1/K2 initial cost");
        LAUNCH(h, k_init<T>, gs1(B), 128, h->D, B, -1, 1);
        launch_cost(h, B, 0, B, B <= h->prefetch_below);
        LAUNCH(h, k_sum_cost<T>, gs1(B), 128, h->D, B, 0);
    }
    // One round = one line-search step for every running instance.  The verdict kernel publishes
    // (rounds completed, instances still running, length of the next work list) into mapped host
    // memory; the host keeps at most run_ahead rounds queued beyond the last count it has seen and
    // stops at zero.  Work lists only ever shrink, so the last length the host has seen bounds every
    // later round: it sizes the grids (no tail of empty CTAs once most instances are done) and picks
    // the kernel variants (all variants of a stage return the same bits, so a big batch switches to
    // the latency-regime kernels for its stragglers).
    int launched = 0;
    bool fused_rounds = false;  // rounds have run whose records hold no control half (see `fused` below)
    int level = 0, repack_bound[kRepackLevels], repack_off[kRepackLevels + 1] = {0};
    // Look-ahead rounds (see k_adopt): the whole solve of a latency-bound batch.  Needs the piped rollout and the
    // staged backward pass (their look-ahead forms are the ones written), no augmented-Lagrangian template (its
    // multiplier updates re-cost the current trajectory between iterations) and the spare buffers of create.
    // A larger batch (of a handle that has the buffers) starts with sequential rounds — with thousands of instances the
    // stages are throughput-bound and the two streams only get in each other's way — and moves to look-ahead rounds
    // once its work list is down to that size: la_ok = the solve may, la = it is doing so.
    const int la_switch = std::min(h->lookahead_below, h->lookahead > 1 ? h->lookahead : kLookaheadDefault);
    // (the switch in the middle of a solve only on request — an explicit bound: it is worth 5 % on C1 x 4096 in fp32 and
    // nothing in fp64, profiles/r02_lookahead.txt — so by default a larger batch runs exactly the sequential rounds)
    const bool la_ok = !kParity && h->lookahead && Bfull <= h->lookahead_below && Bfull <= h->prefetch_below && h->staged &&
                       h->pipeline && N + 1 <= kPipeMaxSteps && !h->any_alm && (Bfull <= la_switch || h->lookahead > 1);
    bool la = la_ok && Bfull <= la_switch;
    // sequential rounds of such a solve alternate between the halves of the trial pool too, so that the trial accepted
    // in the last one is still in place when the first look-ahead round reads it
    struct PoolGuard {
        Dev<T>& D;
        ~PoolGuard() { D.pool_base = 0, D.pool_cap = D.Vs; }
    } pool_guard{h->D};
    if (la_ok) h->D.pool_cap = h->D.Vs / 2;
    Dev<T> Dl = h->D;  // the launch arguments of a look-ahead round
    static const int la_serial = getenv("CILQR_LA_SERIAL") ? atoi(getenv("CILQR_LA_SERIAL")) : 0;  // debugging: 1 = everything on one stream
    const cudaStream_t sb = la_serial ? h->stream : h->stream_b;
    static const int la_cost_late = getenv("CILQR_LA_COST_LATE") ? atoi(getenv("CILQR_LA_COST_LATE")) : 1;
    if (la_ok) {
        Dl.spec = 1;
        Dl.wide_step = 0;
        static const int spec_all_env = getenv("CILQR_LA_SPEC_ALL") ? atoi(getenv("CILQR_LA_SPEC_ALL")) : 1024;
        Dl.spec_all_below = spec_all_env;
    }
    if (la) {
        // jobs of round 0: every instance needs the backward pass of its initial trajectory
        Dl.round_id = 0;
        Dl.pool_base = 0;
        LAUNCH(h, k_adopt<T>, gs1(B), 128, Dl, 0, 1);
    }
    // the spin below must not outlive a device fault or a stalled kernel: every kSpinCheck polls the stream is
    // queried (a sticky error, or an idle stream whose progress words still say "rounds outstanding", ends the
    // solve with CILQR_ERR_CUDA), and a round that makes no progress for kStallSeconds is reported as a stall
    constexpr unsigned kSpinCheck = 1u << 16;
    constexpr double kStallSeconds = 60.0;
    unsigned spins = 0;
    int last_done = -1;
    auto last_progress = std::chrono::steady_clock::now();
    while (launched < h->max_rounds) {
        const unsigned long long w = progress[0];
        const int done = int(w >> 32);
        if (done > 0 && unsigned(w) == 0u) break;
        if (launched - done > h->run_ahead) {  // spin on the mapped words
            if (done != last_done) {
                last_done = done;
                spins = 0;
                last_progress = std::chrono::steady_clock::now();
            } else if (++spins % kSpinCheck == 0) {
                const cudaError_t q = cudaStreamQuery(h->stream);
                if (q != cudaSuccess && q != cudaErrorNotReady)
                    return fail(CILQR_ERR_CUDA, "solve aborted after %d of %d queued rounds: %s", done, launched, cudaGetErrorString(q));
                if (q == cudaSuccess && int(progress[0] >> 32) == done)
                    return fail(CILQR_ERR_CUDA, "solve stalled: the stream is idle but only %d of %d queued rounds reported", done, launched);
                const double idle = std::chrono::duration<double>(std::chrono::steady_clock::now() - last_progress).count();
                if (idle > kStallSeconds)
                    return fail(CILQR_ERR_CUDA, "solve stalled: no round completed for %.0f s (%d of %d queued rounds done)", idle, done, launched);
            }
            continue;
        }
        const int n_bound = std::max(1, std::min(B, int(unsigned(progress[1]))));
        // Repack: the survivors of a large batch have thinned out to half of the slots in use -> move
        // them into a dense prefix (swap_instances) and carry on as a batch of that size.
        if (!la && h->repack && level < kRepackLevels && B > (h->repack > 1 ? h->repack : kRepackMinBatch) && size_t(n_bound) * 2 <= size_t(B)) {
            LAUNCH(h, k_plan_repack<T>, dim3(1), kPlanThreads, h->D, launched & 1, level, h->D.swap_src + repack_off[level],
                   h->D.swap_dst + repack_off[level]);
            swap_instances(h, level, repack_off[level], n_bound);
            repack_bound[level] = n_bound;
            repack_off[level + 1] = repack_off[level] + n_bound;  // <= B / 2 pairs per level: < Bs in total
            ++level;
            B = n_bound;
        }
        const int par = launched & 1;  // which of the two work lists this round reads
        if (la_ok && !la && launched > 0 && n_bound <= la_switch) {
            // from here on look-ahead rounds: what the verdict kernel of the last sequential round left behind is taken
            // over as if no job had run (every instance that needs a backward pass gets one for this round)
            la = true;
            Dl.round_id = launched;
            Dl.pool_base = par * Dl.pool_cap;
            LAUNCH(h, k_adopt<T>, gs1(n_bound), 128, Dl, par, 1);
        }
        h->D.pool_base = la_ok ? par * h->D.pool_cap : 0;
        if (la) {
            // stream S: rollouts + match | costs, verdict |        adopt
            // stream B:                  | derivatives, recursion /
            const int trial_bound = int(std::min<long long>(Dl.pool_cap, (long long)n_bound * kNumAlphas));
            const int e = launched & 3;
            Dl.round_id = launched;
            Dl.pool_base = par * Dl.pool_cap;
            mark_stage(h, 2);
            nvtxRangePushA("K6+K1 rollouts, waypoint match");
            if (launched > 0) {  // (round 0 has no trials: every instance starts with a backward job)
                const int blocks = std::max(1, std::min((trial_bound + kPipeTrials - 1) / kPipeTrials, kGridCap));
                // (16 scan lanes per trial while the trials are likely to fit one wave of one block per SM: a round uses
                // about one slot per instance, not the 20 the bound allows for; more than that only costs a second pass)
                const bool narrow = h->pipeline == 8 || (h->pipeline == 1 && std::min(trial_bound, 3 * n_bound + 256) > 148 * kPipeTrials);
                if (narrow) LAUNCH(h, (k_rollout_match<T, 8, true>), dim3(blocks), pipe_threads(8), Dl, B);
                else LAUNCH(h, (k_rollout_match<T, 16, true>), dim3(blocks), pipe_threads(16), Dl, B);
            }
            nvtxRangePop();
            CK(cudaEventRecord(h->ev_fork[e], h->stream));
            CK(cudaStreamWaitEvent(sb, h->ev_fork[e], 0));
            nvtxRangePushA("K3+K4+K5 backward jobs (second stream)");
            mark_stage_b(h, 0);
            {
                // blocks [0, y_list): the work list (copies of accepted trials, the instances' own jobs); from y_list
                // on: the speculative jobs of this round's trial slots
                const int y_list = std::max(1, std::min((n_bound + 127) / 128, kGridCap / 2));
                // (grid-stride over the slots in use, which the host does not know: usually one per instance)
                const int y_slots = launched > 0 ? std::max(1, std::min((trial_bound + 127) / 128, 2 * y_list + 8)) : 0;
                LAUNCH_ON(h, sb, (k_derivs<T, -1, false>), dim3(2 * (N + 1), y_list + y_slots), 128, Dl, B, 2, par, y_list);
            }
            if (la_cost_late) CK(cudaEventRecord(h->ev_mid[e], sb));
            mark_stage_b(h, 1);
            launch_staged_backward<T>(h, sb, Dl, B + (launched > 0 ? std::min(trial_bound, 2 * n_bound + 1024) : 0), B, 2);
            if (la_serial == 2) CK(cudaDeviceSynchronize());
            mark_stage_b(h, -1);
            CK(cudaEventRecord(h->ev_join[e], sb));
            nvtxRangePop();
            nvtxRangePushA("K2 trial costs, K7 verdict, adopt");
            // the step costs wait for the derivative kernel: both are step-parallel fp64 work and slow each other down
            // (measured), whereas the recursion that follows on the other stream occupies a handful of warps
            if (la_cost_late) CK(cudaStreamWaitEvent(h->stream, h->ev_mid[e], 0));
            mark_stage(h, 4);
            if (launched > 0) LAUNCH(h, (k_cost<T, 4, false>), gk(trial_bound, N + 1), 128, Dl, B, 1);
            mark_stage(h, 5);
            h->scan_epoch = (h->scan_epoch % 0x3fffffffu) + 1u;
            LAUNCH(h, k_decide<T>, dim3((n_bound + 127) / 128), 128, Dl, B, par, h->scan_epoch);
            CK(cudaStreamWaitEvent(h->stream, h->ev_join[e], 0));
            Dl.round_id = launched + 1;
            Dl.pool_base = (par ^ 1) * Dl.pool_cap;
            LAUNCH(h, k_adopt<T>, gs1(n_bound), 128, Dl, par ^ 1, 0);
            nvtxRangePop();
            mark_stage(h, -1);
            ++launched;
            continue;
        }
        const int trial_bound = int(std::min<long long>(h->D.pool_cap, (long long)n_bound * kNumAlphas));
        const bool lat = n_bound <= h->prefetch_below;
        mark_stage(h, 0);
        nvtxRangePushA("K3+K4 derivatives");
        // Bandwidth-bound rounds of a barrier-type solve: the backward pass computes the control half of the records
        // (l_u, l_uu, A, B) from the trajectory itself (k_backward<T, true, true>), so only the state half of the
        // derivative stage is launched and 14 of the 28 record fields are neither written nor read back.
        const bool fused = !kParity && !lat && h->fused_backward && !h->any_alm;
        if (lat) {
            if (fused_rounds) {
                // the cached records of this solve have no control half yet: the latency-regime kernels read all 28 fields
                LAUNCH_DERIVS(h, 1, gk(n_bound, N + 1), h->D, B, 5, par);
                fused_rounds = false;
            }
            LAUNCH_DERIVS(h, -1, gk(n_bound, 2 * (N + 1)), h->D, B, 1, par);
        } else if (fused) {
            LAUNCH_DERIVS(h, 0, gk(n_bound, N + 1), h->D, B, 4, par);
            fused_rounds = true;
        } else {
            LAUNCH_DERIVS(h, 0, gk(n_bound, N + 1), h->D, B, 1, par);
            LAUNCH_DERIVS(h, 1, gk(n_bound, N + 1), h->D, B, 1, par);
        }
        if (h->any_alm) {
            LAUNCH_COST(h, 4, gk(B, N + 1), h->D, B, 0);
            LAUNCH(h, k_sum_cost<T>, gs1(B), 128, h->D, B, 1);
        }
        nvtxRangePop();
        mark_stage(h, 1);
        nvtxRangePushA("K5 backward pass");
        h->D.wide_step = (!lat && h->wide_step) ? 1 : 0;
        switch (fused ? 3 : backward_variant<T>(h, n_bound, B, lat)) {
#ifndef CILQR_PARITY
            case 3:
                LAUNCH(h, (k_backward<T, true, true>), bw_grid(n_bound), kBwThreads, h->D, B, 1, par);
                break;
#endif
            case 2:  // small batch: one warp per tile of 32 instances, records staged through shared memory
                launch_staged_backward<T>(h, h->stream, h->D, B, B, 1);
                break;
            case 1:  // a short work list is spread over one warp per scheduler (see k_backward)
                if (lat) LAUNCH(h, (k_backward<T, true>), gs1(std::max(n_bound, 148 * 128)), 128, h->D, B, 1, par);
                else LAUNCH(h, (k_backward<T, true>), bw_grid(n_bound), kBwThreads, h->D, B, 1, par);
                break;
            default:
                LAUNCH(h, (k_backward<T, false>), bw_grid(n_bound), kBwThreads, h->D, B, 1, par);
        }
        nvtxRangePop();
        mark_stage(h, 2);
        nvtxRangePushA("K6+K1+K2 rollouts, waypoint match, trial costs");
        // (the parity build keeps to the one-thread rollout: the two-lane kernels split the step's trigonometry
        // in a way that is a few ulp from the reference's sequence)
        const bool piped = !kParity && lat && h->pipeline && N + 1 <= kPipeMaxSteps;
        if (piped) {
            const int blocks = std::max(1, std::min((trial_bound + kPipeTrials - 1) / kPipeTrials, kGridCap));
            // 16 scan lanes per trial when the whole trial pool fits one wave at two blocks per SM
            const bool narrow = h->pipeline == 8 || (h->pipeline == 1 && trial_bound > 2 * 148 * kPipeTrials);
            if (narrow) {
                LAUNCH(h, (k_rollout_match<T, 8>), dim3(blocks), pipe_threads(8), h->D, B);
            } else {
                LAUNCH(h, (k_rollout_match<T, 16>), dim3(blocks), pipe_threads(16), h->D, B);
            }
        } else if (lat && !kParity) {
            LAUNCH(h, k_forward2<T>, gs1(2 * trial_bound), 128, h->D, B);  // two lanes per trial slot
        } else {
            LAUNCH(h, (k_forward<T, true>), gs1(trial_bound), 128, h->D, B, 1);  // rollout + waypoint match
        }
        launch_cost(h, B, 1, trial_bound, lat, piped || !lat || kParity);
        nvtxRangePop();
        mark_stage(h, 5);
        Nvtx r7("K7 verdict");
        h->scan_epoch = (h->scan_epoch % 0x3fffffffu) + 1u;
        // one CTA per chunk of the work list, never fewer (no striding): a CTA that went on to a second
        // chunk would wait, in the look-back, on chunks whose CTAs cannot start before it exits
        LAUNCH(h, k_decide<T>, dim3((n_bound + 127) / 128), 128, h->D, B, par, h->scan_epoch);
        mark_stage(h, -1);
        ++launched;
    }
    // commit a step accepted in the last round
    if (la) {
        // (every instance is done: no jobs left, only trial slots waiting to be copied; then the current copy of
        // the gains goes back into the first one)
        Dl.round_id = launched;
        const dim3 g_fin = gk(B, 2 * (N + 1));
        LAUNCH(h, (k_derivs<T, -1, false>), g_fin, 128, Dl, B, 2, launched & 1, int(g_fin.y));
        LAUNCH(h, k_gains_home<T>, gs2(B, N * 10 + 2), 128, Dl, B);
        LAUNCH(h, k_pack_int, grid1(B), 128, static_cast<const int*>(nullptr), h->D.gsel, B, 0);
    } else {
        LAUNCH_DERIVS(h, -1, gk(B, 2 * (N + 1)), h->D, B, 1, launched & 1);
    }
    // every instance back into its own slot
    while (level > 0) {
        --level;
        swap_instances(h, level, repack_off[level], repack_bound[level]);
    }
    B = Bfull;
    LAUNCH(h, k_store_last_u<T>, gs2(B, N), 128, h->D, B);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    if (h->profile) {
        for (int i = 0; i < 6; ++i) {
            h->stage_ms[i] = 0;
            h->stage_launches[i] = 0;
        }
        if (const char* dump = getenv("CILQR_PROFILE_DUMP")) {  // development: every stage interval of the solve, in order
            if (FILE* f = fopen(dump, "w")) {
                for (int st = 0; st < 2; ++st) {
                    const auto& ev = st ? h->prof_ev_b : h->prof_ev;
                    const auto& id = st ? h->prof_stage_b : h->prof_stage;
                    for (size_t i = 0; i + 1 < id.size(); ++i) {
                        float ms = 0, t0 = 0;
                        cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
                        cudaEventElapsedTime(&t0, h->prof_ev[0], ev[i]);
                        fprintf(f, "%d %d %.2f %.2f\n", st, id[i], t0 * 1e3f, ms * 1e3f);
                    }
                }
                fclose(f);
            }
        }
        for (size_t i = 0; i + 1 < h->prof_stage.size(); ++i) {
            int id = h->prof_stage[i];
            if (id < 0) continue;
            float ms = 0;
            if (cudaEventElapsedTime(&ms, h->prof_ev[i], h->prof_ev[i + 1]) == cudaSuccess) {
                h->stage_ms[id] += ms;
                h->stage_launches[id] += 1;
            }
        }
        h->prof_stage.clear();
        for (size_t i = 0; i + 1 < h->prof_stage_b.size(); ++i) {
            int id = h->prof_stage_b[i];
            if (id < 0) continue;
            float ms = 0;
            if (cudaEventElapsedTime(&ms, h->prof_ev_b[i], h->prof_ev_b[i + 1]) == cudaSuccess) {
                h->stage_ms[id] += ms;
                h->stage_launches[id] += 1;
            }
        }
        h->prof_stage_b.clear();
    }
    int ctl[CTL_WORDS];
    CK(cudaMemcpy(ctl, h->D.ctl, sizeof ctl, cudaMemcpyDeviceToHost));
    h->counters.rounds = ctl[CTL_ROUND];
    {
        unsigned long long t;
        memcpy(&t, &ctl[CTL_TRIALS], sizeof t);
        h->counters.total_trials = int64_t(t);
    }
    h->counters.launches = h->launches;
    return 0;
}

template <typename T>
int do_download(Impl<T>* h, int B, double* u_out, double* x_out, double* J_out, double* K_out, double* d_out,
                double* step_cost_out, int32_t* status_out, int32_t* iters_out, int32_t* exit_out) {
    CK(cudaSetDevice(h->device));
    const int N = h->N;
    int rc;
    if ((rc = unpack_to_host(h, h->D.U, u_out, B, N * 2))) return rc;
    if ((rc = unpack_to_host(h, h->D.X, x_out, B, (N + 1) * 4))) return rc;
    if ((rc = unpack_to_host(h, h->D.Kg, K_out, B, N * 8))) return rc;
    if ((rc = unpack_to_host(h, h->D.dg, d_out, B, N * 2))) return rc;
    if ((rc = unpack_to_host(h, h->D.sc, step_cost_out, B, N + 1))) return rc;
    if (J_out) {
        // J_init and J_cur are two [Bs] arrays; emit [B][2]
        std::vector<double> tmp(size_t(B) * 2);
        std::vector<T> a(B), c(B);
        CK(cudaMemcpyAsync(a.data(), h->D.J_init, size_t(B) * sizeof(T), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(c.data(), h->D.J_cur, size_t(B) * sizeof(T), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        for (int b = 0; b < B; ++b) {
            J_out[size_t(b) * 2] = double(a[b]);
            J_out[size_t(b) * 2 + 1] = double(c[b]);
        }
    }
    if ((rc = download_ints(h, h->D.status, status_out, B))) return rc;
    if ((rc = download_ints(h, h->D.iters, iters_out, B))) return rc;
    if ((rc = download_ints(h, h->D.exit_reason, exit_out, B))) return rc;
    // counters
    std::vector<int> it(B), ex(B);
    CK(cudaMemcpyAsync(it.data(), h->D.iters, size_t(B) * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(ex.data(), h->D.exit_reason, size_t(B) * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->counters.total_iters = 0;
    h->counters.exits[0] = h->counters.exits[1] = h->counters.exits[2] = 0;
    int it_max = 0;
    for (int b = 0; b < B; ++b) {
        h->counters.total_iters += it[b];
        it_max = std::max(it_max, it[b]);
        if (ex[b] >= 0 && ex[b] < 3) h->counters.exits[ex[b]]++;
    }
    // batch summary (the reference logs one line per solve, cpp:128-148; a batch gets one line in all), on request:
    // CILQR_B200_LOG=1 in the environment
    static const bool log_summary = [] { const char* e = getenv("CILQR_B200_LOG"); return e && *e && *e != '0'; }();
    if (log_summary)
        fprintf(stderr, "[cilqr_b200] %d solves (%s, N=%d): %lld iter_steps (mean %.1f, max %d), %lld line-search trials, "
                "%d device rounds, %d launches; exits: %d converged, %d max_iter, %d max_lamb\n",
                B, h->dtype == CILQR_F64 ? "fp64" : "fp32", h->N, (long long)h->counters.total_iters,
                double(h->counters.total_iters) / std::max(B, 1), it_max, (long long)h->counters.total_trials, h->counters.rounds,
                h->counters.launches, h->counters.exits[EX_CONVERGED], h->counters.exits[EX_MAX_ITER], h->counters.exits[EX_MAX_LAMB]);
    return 0;
}

// ---- stage operators -------------------------------------------------------

template <typename T>
int stage_init(Impl<T>* h, int B, const double* x0, const int32_t* tmpl, int warm, const double* last_u,
               double* u_out, double* x_out) {
    CK(cudaSetDevice(h->device));
    int rc;
    const int N = h->N;
    if ((rc = upload_ints(h, tmpl, h->D.tmpl, B, 0))) return rc;
    if ((rc = pack_to_device(h, x0, h->D.x0, B, 1, 4, 4, 1))) return rc;
    if (warm) {
        if ((rc = pack_to_device(h, last_u, h->D.last_u, B, 1, N * 2, N * 2, 1))) return rc;
    }
    LAUNCH(h, k_init<T>, gs1(B), 128, h->D, B, warm ? 1 : 0, 0);
    if ((rc = unpack_to_host(h, h->D.U, u_out, B, N * 2))) return rc;
    if ((rc = unpack_to_host(h, h->D.X, x_out, B, (N + 1) * 4))) return rc;
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

template <typename T>
int load_traj(Impl<T>* h, int B, const double* u, const double* x) {
    int rc;
    const int N = h->N;
    if (u && (rc = pack_to_device(h, u, h->D.U, B, 1, N * 2, N * 2, 1))) return rc;
    if (x && (rc = pack_to_device(h, x, h->D.X, B, 1, (N + 1) * 4, (N + 1) * 4, 1))) return rc;
    return 0;
}

template <typename T>
int stage_ref_match(Impl<T>* h, int B, const double* x, const int32_t* tmpl, int32_t* idx_out) {
    CK(cudaSetDevice(h->device));
    int rc;
    if ((rc = check_tmpl_nobs(h, B, tmpl, nullptr, 0))) return rc;
    if ((rc = upload_ints(h, tmpl, h->D.tmpl, B, 0))) return rc;
    if ((rc = load_traj(h, B, nullptr, x))) return rc;
    constexpr int G = 8;
    LAUNCH(h, k_ref_match<T, G>, gs1(B * G), 128, h->D, B, 0);
    // ridx is [N+1][Bs] ints: transpose on the host (test path only)
    const int N = h->N;
    std::vector<int> tmp(size_t(N + 1) * h->Bs);
    CK(cudaMemcpyAsync(tmp.data(), h->D.ridx, tmp.size() * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (int b = 0; b < B; ++b)
        for (int k = 0; k <= N; ++k) idx_out[size_t(b) * (N + 1) + k] = tmp[size_t(k) * h->Bs + b];
    return 0;
}

template <typename T>
int load_alm(Impl<T>* h, int B, const double* alm_mu, const double* alm_rho) {
    if (!h->any_alm) return 0;
    if (!alm_mu || !alm_rho) return fail(CILQR_ERR_INVALID, "ALM template needs alm_mu and alm_rho");
    int rc;
    if ((rc = pack_to_device(h, alm_mu, h->D.mu, B, 1, h->N * h->D.alm_cols, h->N * h->D.alm_cols, 1))) return rc;
    if ((rc = pack_to_device(h, alm_rho, h->D.rho, B, 1, 1, 1, 1))) return rc;
    return 0;
}

template <typename T>
int stage_cost(Impl<T>* h, int B, const double* u, const double* x, const double* ref_velo, const double* borders,
               const int32_t* tmpl, const int32_t* n_obs, const double* obs, int obs_len, const double* alm_mu,
               const double* alm_rho, double* J_out, double* step_cost_out) {
    CK(cudaSetDevice(h->device));
    int rc;
    if ((rc = upload_problem_data(h, B, ref_velo, borders, tmpl, n_obs, obs, obs_len))) return rc;
    if ((rc = load_traj(h, B, u, x))) return rc;
    if ((rc = load_alm(h, B, alm_mu, alm_rho))) return rc;
    launch_cost(h, B, 0, B, B <= h->prefetch_below);
    LAUNCH(h, k_sum_cost<T>, gs1(B), 128, h->D, B, 0);
    if ((rc = unpack_to_host(h, h->D.J_cur, J_out, B, 1))) return rc;
    if ((rc = unpack_to_host(h, h->D.sc, step_cost_out, B, h->N + 1))) return rc;
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

// dense per-stage outputs are written by a kernel straight into the staging buffer, chunk by chunk
template <typename T>
int stage_derivs(Impl<T>* h, int B, const double* u, const double* x, const double* ref_velo, const double* borders,
                 const int32_t* tmpl, const int32_t* n_obs, const double* obs, int obs_len, const double* alm_mu,
                 const double* alm_rho, double* lx, double* lu, double* lxx, double* luu, double* A, double* Bm,
                 double* alm_mu_next) {
    CK(cudaSetDevice(h->device));
    int rc;
    const int N = h->N;
    if ((rc = upload_problem_data(h, B, ref_velo, borders, tmpl, n_obs, obs, obs_len))) return rc;
    if ((rc = load_traj(h, B, u, x))) return rc;
    if ((rc = load_alm(h, B, alm_mu, alm_rho))) return rc;
    constexpr int G = 8;
    LAUNCH(h, k_ref_match<T, G>, gs1(B * G), 128, h->D, B, 0);
    LAUNCH_DERIVS(h, -1, gk(B, 2 * (N + 1)), h->D, B, 0, 0);
    // dense conversion on device into a temporary allocation (test path; not part of create-time budget)
    size_t n_lx = size_t(B) * (N + 1) * 4, n_lu = size_t(B) * N * 2, n_lxx = size_t(B) * (N + 1) * 16,
           n_luu = size_t(B) * N * 4, n_A = size_t(B) * N * 16, n_B = size_t(B) * N * 8;
    double* tmp = nullptr;
    size_t total = n_lx + n_lu + n_lxx + n_luu + n_A + n_B;
    CK(cudaMalloc(&tmp, total * sizeof(double)));
    double *p_lx = tmp, *p_lu = p_lx + n_lx, *p_lxx = p_lu + n_lu, *p_luu = p_lxx + n_lxx, *p_A = p_luu + n_luu,
           *p_B = p_A + n_A;
    LAUNCH(h, k_records_to_dense<T>, grid2(B, N + 1), 128, h->D, B, p_lx, p_lu, p_lxx, p_luu, p_A, p_B);
    cudaError_t e = cudaSuccess;
    auto get = [&](double* dst, const double* src, size_t n) {
        if (dst && e == cudaSuccess) e = cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    };
    get(lx, p_lx, n_lx);
    get(lu, p_lu, n_lu);
    get(lxx, p_lxx, n_lxx);
    get(luu, p_luu, n_luu);
    get(A, p_A, n_A);
    get(Bm, p_B, n_B);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(tmp);
    if (e != cudaSuccess) return fail(CILQR_ERR_CUDA, "stage_derivs download: %s", cudaGetErrorString(e));
    if (h->any_alm && alm_mu_next) {
        if ((rc = unpack_to_host(h, h->D.mu_next, alm_mu_next, B, N * h->D.alm_cols))) return rc;
        CK(cudaStreamSynchronize(h->stream));
    }
    return 0;
}

template <typename T>
int stage_backward(Impl<T>* h, int B, const double* lx, const double* lu, const double* lxx, const double* luu,
                   const double* A, const double* Bm, const double* lamb, double* d_out, double* K_out,
                   double* dV_out, int32_t* status_out) {
    CK(cudaSetDevice(h->device));
    int rc;
    const int N = h->N;
    size_t n_lx = size_t(B) * (N + 1) * 4, n_lu = size_t(B) * N * 2, n_lxx = size_t(B) * (N + 1) * 16,
           n_luu = size_t(B) * N * 4, n_A = size_t(B) * N * 16, n_B = size_t(B) * N * 8;
    size_t total = n_lx + n_lu + n_lxx + n_luu + n_A + n_B;
    double* tmp = nullptr;
    CK(cudaMalloc(&tmp, total * sizeof(double)));
    double *p_lx = tmp, *p_lu = p_lx + n_lx, *p_lxx = p_lu + n_lu, *p_luu = p_lxx + n_lxx, *p_A = p_luu + n_luu,
           *p_B = p_A + n_A;
    cudaError_t e = cudaSuccess;
    auto put = [&](double* dst, const double* src, size_t n) {
        if (e == cudaSuccess) e = cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    };
    put(p_lx, lx, n_lx);
    put(p_lu, lu, n_lu);
    put(p_lxx, lxx, n_lxx);
    put(p_luu, luu, n_luu);
    put(p_A, A, n_A);
    put(p_B, Bm, n_B);
    if (e != cudaSuccess) {
        cudaFree(tmp);
        return fail(CILQR_ERR_CUDA, "stage_backward upload: %s", cudaGetErrorString(e));
    }
    LAUNCH(h, k_records_from_dense<T>, grid2(B, N + 1), 128, h->D, B, p_lx, p_lu, p_lxx, p_luu, p_A, p_B);
    if ((rc = pack_to_device(h, lamb, h->D.lamb, B, 1, 1, 1, 1))) {
        cudaFree(tmp);
        return rc;
    }
    {
        const int variant = h->bench_prefetch >= 0 ? h->bench_prefetch : backward_variant<T>(h, B, B, B <= h->prefetch_below);
        if (variant == 2) {
            launch_staged_backward<T>(h, h->stream, h->D, B, B, 0);
#ifndef CILQR_PARITY
        } else if (variant == 3) {
            LAUNCH(h, (k_backward<T, true, true>), bw_grid(B), kBwThreads, h->D, B, 0, 0);
#endif
        } else if (variant == 1) {
            LAUNCH(h, (k_backward<T, true>), bw_grid(B), kBwThreads, h->D, B, 0, 0);
        } else {
            LAUNCH(h, (k_backward<T, false>), bw_grid(B), kBwThreads, h->D, B, 0, 0);
        }
    }
    e = cudaStreamSynchronize(h->stream);
    cudaFree(tmp);
    if (e != cudaSuccess) return fail(CILQR_ERR_CUDA, "stage_backward: %s", cudaGetErrorString(e));
    if ((rc = unpack_to_host(h, h->D.dg, d_out, B, N * 2))) return rc;
    if ((rc = unpack_to_host(h, h->D.Kg, K_out, B, N * 8))) return rc;
    if ((rc = unpack_to_host(h, h->D.dV, dV_out, B, 2))) return rc;
    if ((rc = download_ints(h, h->D.status, status_out, B))) return rc;
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

template <typename T>
int stage_forward(Impl<T>* h, int B, const double* u, const double* x, const double* d, const double* K,
                  const double* alpha, const int32_t* tmpl, double* new_u, double* new_x) {
    CK(cudaSetDevice(h->device));
    int rc;
    const int N = h->N;
    if ((rc = upload_ints(h, tmpl, h->D.tmpl, B, 0))) return rc;
    if ((rc = load_traj(h, B, u, x))) return rc;
    if ((rc = pack_to_device(h, d, h->D.dg, B, 1, N * 2, N * 2, 1))) return rc;
    if ((rc = pack_to_device(h, K, h->D.Kg, B, 1, N * 8, N * 8, 1))) return rc;
    if ((rc = pack_to_device(h, alpha, h->D.alpha, B, 1, 1, 1, 1))) return rc;
    LAUNCH(h, (k_forward<T, false>), gs1(B), 128, h->D, B, 0);
    if ((rc = unpack_to_host(h, h->D.Ut, new_u, B, N * 2, size_t(h->D.Vs)))) return rc;
    if ((rc = unpack_to_host(h, h->D.Xt, new_x, B, (N + 1) * 4, size_t(h->D.Vs)))) return rc;
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

template <typename T>
int bench_backward(Impl<T>* h, int B, double lamb, int reps, int flush_l2, float* ms_out, double* bytes_per_launch) {
    CK(cudaSetDevice(h->device));
    if (flush_l2 && !h->flush) {
        h->flush_n = (size_t(256) << 20) / sizeof(float);
        CK(cudaMalloc(&h->flush, h->flush_n * sizeof(float)));
        h->allocs.push_back(h->flush);
        CK(cudaMemsetAsync(h->flush, 0, h->flush_n * sizeof(float), h->stream));
    }
    std::vector<T> l(B, T(lamb));
    CK(cudaMemcpyAsync(h->D.lamb, l.data(), size_t(B) * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (int r = 0; r < reps; ++r) {
        if (flush_l2) {
            k_flush_l2<<<148 * 8, 256, 0, h->stream>>>(h->flush, h->flush_n);
        }
        CK(cudaEventRecord(h->t0, h->stream));
        const int variant = h->bench_prefetch >= 0 ? h->bench_prefetch : backward_variant<T>(h, B, B, B <= h->prefetch_below);
        if (variant == 2) {
            launch_staged_backward<T>(h, h->stream, h->D, B, B, 0);
#ifndef CILQR_PARITY
        } else if (variant == 3) {
            LAUNCH(h, (k_backward<T, true, true>), bw_grid(B), kBwThreads, h->D, B, 0, 0);
#endif
        } else if (variant == 1) {
            LAUNCH(h, (k_backward<T, true>), bw_grid(B), kBwThreads, h->D, B, 0, 0);
        } else {
            LAUNCH(h, (k_backward<T, false>), bw_grid(B), kBwThreads, h->D, B, 0, 0);
        }
        CK(cudaEventRecord(h->t1, h->stream));
        CK(cudaEventSynchronize(h->t1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, h->t0, h->t1));
        if (ms_out) ms_out[r] = ms;
    }
    CK(cudaGetLastError());
    if (bytes_per_launch) *bytes_per_launch = double(38 * h->N + 18) * sizeof(T) * double(B);
    return 0;
}

template <typename T>
int bench_tile(Impl<T>* h, int B0, int B) {
    CK(cudaSetDevice(h->device));
    if (B0 <= 0 || B0 > B) return fail(CILQR_ERR_INVALID, "need 0 < B0 <= B");
    LAUNCH(h, k_tile_records<T>, grid2(B, (h->N + 1) * kRecFields), 128, h->D, B0, B);
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

template <typename T>
int do_reset(Impl<T>* h) {
    CK(cudaSetDevice(h->device));
    std::vector<int> ones(h->Bs, 1);
    CK(cudaMemcpyAsync(h->D.first, ones.data(), ones.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

// Receding-horizon closed loop of src/motion_planning.cpp:180-197 for B scenarios, on the device.
template <typename T>
int do_simulate(Impl<T>* h, int B, const double* x0, const double* ref_velo, const double* borders,
                const int32_t* tmpl, const int32_t* n_obs, const double* tracks, int track_len, int ticks,
                double* ego_out, int32_t* iters_out, int32_t* status_out) {
    CK(cudaSetDevice(h->device));
    const int N = h->N;
    if (ticks < 1) return fail(CILQR_ERR_INVALID, "ticks must be >= 1");
    bool any = false;
    if (n_obs)
        for (int b = 0; b < B && !any; ++b) any = n_obs[b] > 0;
    std::vector<int> offs(ticks);
    {
        const double dt = h->params[0].dt;
        double t = 0.;
        for (int i = 0; i < ticks; ++i, t += dt) offs[i] = int(size_t(t / dt));
    }
    if (any && track_len < offs[ticks - 1] + N + 1)
        return fail(CILQR_ERR_RANGE, "obstacle tracks hold %d samples, the last tick needs %d (RoutingLine index out of range)",
                    track_len, offs[ticks - 1] + N + 1);
    int rc;
    if ((rc = check_tmpl_nobs(h, B, tmpl, n_obs, track_len))) return rc;
    // set-up allocations (grow-only): full tracks and the per-tick history
    if (track_len > h->track_cap) {
        if ((rc = dalloc(h, &h->obs_tracks, size_t(h->max_obs) * track_len * 4 * h->Bs))) return rc;
        h->track_cap = track_len;
    }
    if (ticks > h->hist_cap) {
        if ((rc = dalloc(h, &h->hist_x, size_t(ticks + 1) * 4 * h->Bs))) return rc;
        if ((rc = dalloc(h, &h->hist_iters, size_t(ticks) * h->Bs))) return rc;
        if ((rc = dalloc(h, &h->hist_status, size_t(ticks) * h->Bs))) return rc;
        h->hist_cap = ticks;
    }
    if ((rc = upload_ints(h, tmpl, h->D.tmpl, B, 0))) return rc;
    if ((rc = upload_ints(h, n_obs, h->D.n_obs, B, 0))) return rc;
    if ((rc = pack_to_device(h, ref_velo, h->D.ref_velo, B, 1, 1, 1, 1))) return rc;
    if ((rc = pack_to_device(h, borders, h->D.borders, B, 1, 2, 2, 1))) return rc;
    if ((rc = pack_to_device(h, x0, h->D.x0, B, 1, 4, 4, 1))) return rc;
    if (any && h->max_obs > 0) {
        if ((rc = pack_to_device(h, tracks, h->obs_tracks, B, h->max_obs, track_len, track_len, 3, 4))) return rc;
        LAUNCH(h, k_obs_sincos<T>, gs2(B, h->max_obs * track_len), 128, h->obs_tracks, B, size_t(h->Bs));
    }
    if ((rc = do_reset(h))) return rc;  // a new simulation starts with is_first_solve == true
    h->D.obs = h->obs_tracks;
    h->D.obs_len = track_len;
    int64_t total_iters = 0, total_trials = 0;
    int rounds = 0, launches = 0;
    // the obstacle window of tick i starts at sample size_t(t / delta_t) with t accumulated in floating point,
    // exactly as the reference's loop does (motion_planning.cpp:180-181: for delta_t = 0.1 the sequence is
    // 0,1,2,3,4,5,5,6,... — the truncation of 0.6 / 0.1 = 5.999...)
    for (int t = 0; t < ticks; ++t) {
        h->D.obs_off = offs[t];
        if ((rc = do_solve_resident(h, B))) break;
        LAUNCH(h, k_advance<T>, gs1(B), 128, h->D, B, t, h->hist_x, h->hist_iters, h->hist_status);
        total_trials += h->counters.total_trials;
        rounds += h->counters.rounds;
        launches += h->launches;
    }
    h->D.obs = h->obs_plain;
    h->D.obs_len = N + 1;
    h->D.obs_off = 0;
    if (rc) return rc;
    CK(cudaGetLastError());
    // history back in host layout: ego [B][ticks+1][4], iters / status [B][ticks]
    if ((rc = unpack_to_host(h, h->hist_x, ego_out, B, (ticks + 1) * 4))) return rc;
    std::vector<int> hi(size_t(ticks) * h->Bs), hs(size_t(ticks) * h->Bs);
    CK(cudaMemcpyAsync(hi.data(), h->hist_iters, hi.size() * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(hs.data(), h->hist_status, hs.size() * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (int b = 0; b < B; ++b)
        for (int t = 0; t < ticks; ++t) {
            total_iters += hi[size_t(t) * h->Bs + b];
            if (iters_out) iters_out[size_t(b) * ticks + t] = hi[size_t(t) * h->Bs + b];
            if (status_out) status_out[size_t(b) * ticks + t] = hs[size_t(t) * h->Bs + b];
        }
    h->counters.total_iters = total_iters;
    h->counters.total_trials = total_trials;
    h->counters.rounds = rounds;
    h->counters.launches = launches;
    return 0;
}

template <typename T>
int do_synth_set_lanes(Impl<T>* h, int n_lanes, const int32_t* off, const double* x, const double* y, const double* yaw,
                       const double* lon, const double* nx, const double* ny) {
    CK(cudaSetDevice(h->device));
    const size_t total = size_t(off[n_lanes]);
    if (total > h->syn_cap || !h->syn_off) {
        double* q = nullptr;
        CK(cudaMalloc(&q, std::max<size_t>(total, 1) * 6 * sizeof(double)));
        h->allocs.push_back(q);
        h->syn_tab = q;
        h->syn_cap = std::max<size_t>(total, 1);
        int* o = nullptr;
        CK(cudaMalloc(&o, 1024 * sizeof(int)));
        h->allocs.push_back(o);
        h->syn_off = o;
    }
    const double* src[6] = {x, y, yaw, lon, nx, ny};
    for (int i = 0; i < 6; ++i)
        CK(cudaMemcpyAsync(h->syn_tab + size_t(i) * h->syn_cap, src[i], total * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->syn_off, off, size_t(n_lanes + 1) * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->syn_lanes = n_lanes;
    return 0;
}

template <typename T>
int do_synth_generate(Impl<T>* h, int B, uint64_t first_id, uint64_t seed, int n_tmpl, const cilqr_synth_template_t* tmpls,
                      int keep_yaw) {
    CK(cudaSetDevice(h->device));
    if (!h->syn_tm) {
        int rc = dalloc(h, &h->syn_tm, CILQR_B200_MAX_TEMPLATES);
        if (rc) return rc;
    }
    for (int t = 0; t < n_tmpl; ++t) {
        if (!h->tmpl_set[t] || h->wp_len[t] <= 0)
            return fail(CILQR_ERR_INVALID, "synthetic template %d has no solver template / reference line (cilqr_b200_set_template)", t);
        const cilqr_synth_template_t& st = tmpls[t];
        if (st.n_obs < 0 || st.n_obs > h->max_obs || st.n_obs > CILQR_B200_SYNTH_MAX_OBS)
            return fail(CILQR_ERR_INVALID, "synthetic template %d has %d obstacles (max_obs %d)", t, st.n_obs, h->max_obs);
        bool lanes_ok = st.ego_kind == 0 || (st.ego_lane >= 0 && st.ego_lane < h->syn_lanes);
        for (int j = 0; j < st.n_obs; ++j)
            lanes_ok = lanes_ok && (st.obs[j].kind != 0 || (st.obs[j].lane >= 0 && st.obs[j].lane < h->syn_lanes));
        if (!lanes_ok) return fail(CILQR_ERR_INVALID, "synthetic template %d refers to a lane table that was not set", t);
    }
    CK(cudaMemcpyAsync(h->syn_tm, tmpls, size_t(n_tmpl) * sizeof(cilqr_synth_template_t), cudaMemcpyHostToDevice, h->stream));
    h->D.obs = h->obs_plain;
    h->D.obs_len = h->N + 1;
    h->D.obs_off = 0;
    SynthLanes L{h->syn_tab, h->syn_tab + h->syn_cap, h->syn_tab + 2 * h->syn_cap, h->syn_tab + 3 * h->syn_cap,
                 h->syn_tab + 4 * h->syn_cap, h->syn_tab + 5 * h->syn_cap, h->syn_off};
    LAUNCH(h, k_synth_instances<T>, gs1(B), 128, h->D, B, (unsigned long long)first_id, (unsigned long long)seed, n_tmpl, h->syn_tm, L);
    if (h->max_obs > 0) {
        LAUNCH(h, k_synth_obstacles<T>, gs2(B, h->max_obs * (h->N + 1)), 128, h->D, h->obs_plain, B, (unsigned long long)first_id,
               (unsigned long long)seed, n_tmpl, h->syn_tm, L);
        if (!keep_yaw) LAUNCH(h, k_obs_sincos<T>, gs2(B, h->max_obs * (h->N + 1)), 128, h->obs_plain, B, size_t(h->Bs));
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

template <typename T>
int do_synth_download(Impl<T>* h, int B, double* x0, double* ref_velo, double* borders, int32_t* tmpl, int32_t* n_obs, double* obs) {
    CK(cudaSetDevice(h->device));
    int rc;
    if ((rc = unpack_to_host(h, h->D.x0, x0, B, 4))) return rc;
    if ((rc = unpack_to_host(h, h->D.ref_velo, ref_velo, B, 1))) return rc;
    if ((rc = unpack_to_host(h, h->D.borders, borders, B, 2))) return rc;
    if ((rc = download_ints(h, h->D.tmpl, tmpl, B))) return rc;
    if ((rc = download_ints(h, h->D.n_obs, n_obs, B))) return rc;
    if (h->max_obs > 0 && (rc = unpack_to_host(h, h->obs_plain, obs, B, h->max_obs * (h->N + 1) * 4))) return rc;
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

template <typename T>
int do_set_template(Impl<T>* h, int t, const cilqr_params_t* params, const double* wx, const double* wy,
                    const double* wyaw, int M) {
    CK(cudaSetDevice(h->device));
    if (params) {
        int rc = check_params(params);
        if (rc) return rc;
        h->params[t] = *params;
        h->tmpl_set[t] = true;
    } else if (!h->tmpl_set[t]) {
        h->params[t] = h->params[0];
        h->tmpl_set[t] = true;
    }
    if (wx || wy || wyaw) {
        if (!(wx && wy && wyaw)) return fail(CILQR_ERR_INVALID, "reference line needs x, y and yaw");
        if (M <= 0 || M > 65535)
            return fail(CILQR_ERR_INVALID, "reference line must hold 1..65535 waypoints (uint16_t index in the reference), got %d", M);
        // rebuild the concatenated table with template t replaced
        std::vector<double> nx, ny, nyaw;
        int off[CILQR_B200_MAX_TEMPLATES], len[CILQR_B200_MAX_TEMPLATES];
        for (int i = 0; i < CILQR_B200_MAX_TEMPLATES; ++i) {
            off[i] = int(nx.size());
            if (i == t) {
                nx.insert(nx.end(), wx, wx + M);
                ny.insert(ny.end(), wy, wy + M);
                nyaw.insert(nyaw.end(), wyaw, wyaw + M);
                len[i] = M;
            } else {
                int o = h->wp_off[i], l = h->wp_len[i];
                nx.insert(nx.end(), h->h_wx.begin() + o, h->h_wx.begin() + o + l);
                ny.insert(ny.end(), h->h_wy.begin() + o, h->h_wy.begin() + o + l);
                nyaw.insert(nyaw.end(), h->h_wyaw.begin() + o, h->h_wyaw.begin() + o + l);
                len[i] = l;
            }
        }
        h->h_wx.swap(nx);
        h->h_wy.swap(ny);
        h->h_wyaw.swap(nyaw);
        for (int i = 0; i < CILQR_B200_MAX_TEMPLATES; ++i) {
            h->wp_off[i] = off[i];
            h->wp_len[i] = len[i];
        }
    }
    return upload_templates(h);
}

template <typename T>
int do_set_option(Impl<T>* h, int option, int value) {
    switch (option) {
        case CILQR_OPT_WIDE_SEARCH:
            h->D.wide_mode = value ? 1 : 0;
            return 0;
        case CILQR_OPT_RUN_AHEAD:
            if (value < 0 || value > 64) return fail(CILQR_ERR_INVALID, "run-ahead must be in [0, 64]");
            h->run_ahead = value;
            return 0;
        case CILQR_OPT_PREFETCH_BELOW:
            h->prefetch_below = value;
            return 0;
        case CILQR_OPT_PROFILE_STAGES:
            h->profile = value ? 1 : 0;
            h->prof_stage.clear();
            return 0;
        case CILQR_OPT_PIPELINE:
            h->pipeline = value;
            return 0;
        case CILQR_OPT_STAGED_BACKWARD:
            h->staged = value ? 1 : 0;
            return 0;
        case CILQR_OPT_WIDE_STEP:
            h->wide_step = value ? 1 : 0;
            return 0;
        case CILQR_OPT_REPACK:
            h->repack = value < 0 ? 0 : value;  // > 1: smallest batch that is still repacked (development)
            return 0;
        case CILQR_OPT_FUSED_BACKWARD:
            h->fused_backward = value ? 1 : 0;
            return 0;
        case CILQR_OPT_LOOKAHEAD:  // 0 off, 1 the default batch bound, > 1 an explicit one
            h->lookahead = value < 0 ? 0 : value;
            return 0;
        case CILQR_OPT_BENCH_PREFETCH:
            h->bench_prefetch = value < 0 ? -1 : (value > 3 ? 3 : value);
            return 0;
        default:
            return fail(CILQR_ERR_INVALID, "unknown option %d", option);
    }
}

template <typename T>
int do_enable_trace(Impl<T>* h, int cap) {
    CK(cudaSetDevice(h->device));
    if (cap < 0 || cap > 100000) return fail(CILQR_ERR_INVALID, "trace capacity out of range");
    if (cap > 0 && cap != h->D.trace_cap) {
        int rc;
        if ((rc = dalloc(h, &h->D.tr_status, size_t(cap) * h->Bs))) return rc;
        if ((rc = dalloc(h, &h->D.tr_alpha, size_t(cap) * h->Bs))) return rc;
        if ((rc = dalloc(h, &h->D.tr_cost, size_t(cap) * h->Bs))) return rc;
        CK(cudaStreamSynchronize(h->stream));
    }
    h->D.trace_cap = cap;
    return 0;
}

template <typename T>
int do_get_trace(Impl<T>* h, int B, int32_t* status, int32_t* alpha, double* cost) {
    CK(cudaSetDevice(h->device));
    const int cap = h->D.trace_cap;
    if (cap <= 0) return fail(CILQR_ERR_INVALID, "trace is not enabled");
    std::vector<int> a(size_t(cap) * h->Bs), c(size_t(cap) * h->Bs);
    CK(cudaMemcpyAsync(a.data(), h->D.tr_status, a.size() * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(c.data(), h->D.tr_alpha, c.size() * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    int rc;
    if ((rc = unpack_to_host(h, h->D.tr_cost, cost, B, cap))) return rc;
    CK(cudaStreamSynchronize(h->stream));
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < cap; ++i) {
            if (status) status[size_t(b) * cap + i] = a[size_t(i) * h->Bs + b];
            if (alpha) alpha[size_t(b) * cap + i] = c[size_t(i) * h->Bs + b];
        }
    return 0;
}

template <typename T>
int do_destroy(Impl<T>* h) {
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    for (void* p : h->allocs) cudaFree(p);
    if (h->h_ctl) cudaFreeHost(const_cast<int*>(h->h_ctl));
    for (auto& e : h->prof_ev) cudaEventDestroy(e);
    for (auto& e : h->prof_ev_b) cudaEventDestroy(e);
    for (int i = 0; i < 4; ++i) {
        if (h->ev_fork[i]) cudaEventDestroy(h->ev_fork[i]);
        if (h->ev_join[i]) cudaEventDestroy(h->ev_join[i]);
        if (h->ev_mid[i]) cudaEventDestroy(h->ev_mid[i]);
    }
    if (h->stream_b) cudaStreamDestroy(h->stream_b);
    if (h->t0) cudaEventDestroy(h->t0);
    if (h->t1) cudaEventDestroy(h->t1);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
    return 0;
}

}  // namespace

extern "C" {

const char* cilqr_b200_last_error(void) { return g_err.c_str(); }
#ifdef CILQR_PARITY
const char* cilqr_b200_version(void) { return "cilqr_b200 0.2 (sm_100a, parity build: reference operation order, portable transcendentals, no FMA contraction)"; }
#else
const char* cilqr_b200_version(void) { return "cilqr_b200 0.2 (sm_100a)"; }
#endif

int cilqr_b200_create(const cilqr_params_t* params, int device, int max_batch, int N, int max_obs, int dtype,
                      cilqr_handle_t** out) {
    if (!out) return fail(CILQR_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int rc = check_params(params);
    if (rc) return rc;
    if (max_batch <= 0) return fail(CILQR_ERR_INVALID, "max_batch must be positive");
    if (max_batch > kMaxBatch)
        return fail(CILQR_ERR_INVALID, "max_batch %d exceeds %d (the verdict kernel's per-round tally holds 24-bit counts); "
                    "shard larger batches over several handles", max_batch, kMaxBatch);
    if (N < 1 || N > 4096) return fail(CILQR_ERR_INVALID, "horizon N must be in [1, 4096]");
    if (max_obs < 0 || max_obs > 64) return fail(CILQR_ERR_INVALID, "max_obs must be in [0, 64]");
    if (dtype != CILQR_F64 && dtype != CILQR_F32) return fail(CILQR_ERR_INVALID, "dtype must be CILQR_F64 or CILQR_F32");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(CILQR_ERR_NO_DEVICE, "no CUDA device (%s); cilqr_b200 has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= count) return fail(CILQR_ERR_INVALID, "device %d outside [0, %d)", device, count);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(CILQR_ERR_NO_DEVICE, "device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major, prop.minor);
    return dtype == CILQR_F64 ? create_impl<double>(params, device, max_batch, N, max_obs, out)
                              : create_impl<float>(params, device, max_batch, N, max_obs, out);
}

int cilqr_b200_destroy(cilqr_handle_t* h) {
    if (!h) return 0;
    return DISPATCH(h, do_destroy);
}

int cilqr_b200_set_stream(cilqr_handle_t* h, void* cuda_stream) {
    if (!h) return fail(CILQR_ERR_INVALID, "handle is NULL");
    Base* b = base(h);
    b->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : b->own_stream;
    return 0;
}

int cilqr_b200_set_template(cilqr_handle_t* h, int tmpl, const cilqr_params_t* params, const double* wx,
                            const double* wy, const double* wyaw, int M) {
    if (!h) return fail(CILQR_ERR_INVALID, "handle is NULL");
    if (tmpl < 0 || tmpl >= CILQR_B200_MAX_TEMPLATES) return fail(CILQR_ERR_INVALID, "template id %d outside [0, %d)", tmpl, CILQR_B200_MAX_TEMPLATES);
    return DISPATCH(h, do_set_template, tmpl, params, wx, wy, wyaw, M);
}

int cilqr_b200_reset(cilqr_handle_t* h) {
    if (!h) return fail(CILQR_ERR_INVALID, "handle is NULL");
    return DISPATCH(h, do_reset);
}

int cilqr_b200_upload(cilqr_handle_t* h, int B, const double* x0, const double* ref_velo, const double* borders,
                      const int32_t* tmpl, const int32_t* n_obs, const double* obs, int obs_len) {
    int rc = check_batch(base(h), B);
    if (rc) return rc;
    if (B == 0) return 0;
    rc = DISPATCH(h, do_upload, B, x0, ref_velo, borders, tmpl, n_obs, obs, obs_len);
    if (rc) return rc;
    CK(cudaStreamSynchronize(base(h)->stream));
    return 0;
}

int cilqr_b200_solve_resident(cilqr_handle_t* h, int B) {
    int rc = check_batch(base(h), B);
    if (rc) return rc;
    return DISPATCH(h, do_solve_resident, B);
}

int cilqr_b200_download(cilqr_handle_t* h, int B, double* u_out, double* x_out, double* J_out, double* K_out,
                        double* d_out, double* step_cost_out, int32_t* status_out, int32_t* iters_out,
                        int32_t* exit_out) {
    int rc = check_batch(base(h), B);
    if (rc) return rc;
    if (B == 0) return 0;
    rc = DISPATCH(h, do_download, B, u_out, x_out, J_out, K_out, d_out, step_cost_out, status_out, iters_out, exit_out);
    if (rc) return rc;
    CK(cudaStreamSynchronize(base(h)->stream));
    return 0;
}

int cilqr_b200_solve_batch(cilqr_handle_t* h, int B, const double* x0, const double* ref_velo, const double* borders,
                           const int32_t* tmpl, const int32_t* n_obs, const double* obs, int obs_len, double* u_out,
                           double* x_out, double* J_out, double* K_out, double* d_out, double* step_cost_out,
                           int32_t* status_out, int32_t* iters_out, int32_t* exit_out) {
    int rc = check_batch(base(h), B);
    if (rc) return rc;
    if (B == 0) return 0;
    rc = DISPATCH(h, do_upload, B, x0, ref_velo, borders, tmpl, n_obs, obs, obs_len);
    if (rc) return rc;
    rc = DISPATCH(h, do_solve_resident, B);
    if (rc) return rc;
    rc = DISPATCH(h, do_download, B, u_out, x_out, J_out, K_out, d_out, step_cost_out, status_out, iters_out, exit_out);
    if (rc) return rc;
    CK(cudaStreamSynchronize(base(h)->stream));
    return 0;
}

int cilqr_b200_simulate(cilqr_handle_t* h, int B, const double* x0, const double* ref_velo, const double* borders,
                        const int32_t* tmpl, const int32_t* n_obs, const double* tracks, int track_len, int ticks,
                        double* ego_out, int32_t* iters_out, int32_t* status_out) {
    int rc = check_batch(base(h), B);
    if (rc) return rc;
    if (B == 0) return 0;
    return DISPATCH(h, do_simulate, B, x0, ref_velo, borders, tmpl, n_obs, tracks, track_len, ticks, ego_out,
                    iters_out, status_out);
}

int cilqr_b200_synth_set_lanes(cilqr_handle_t* h, int n_lanes, const int32_t* lane_off, const double* x, const double* y,
                               const double* yaw, const double* lon, const double* nx, const double* ny) {
    if (!h) return fail(CILQR_ERR_INVALID, "handle is NULL");
    if (n_lanes < 1 || n_lanes > 1023 || !lane_off || !x || !y || !yaw || !lon || !nx || !ny)
        return fail(CILQR_ERR_INVALID, "need 1..1023 lane tables and all six arrays");
    for (int l = 0; l < n_lanes; ++l)
        if (lane_off[l + 1] - lane_off[l] < 2) return fail(CILQR_ERR_INVALID, "lane table %d holds fewer than 2 samples", l);
    return DISPATCH(h, do_synth_set_lanes, n_lanes, lane_off, x, y, yaw, lon, nx, ny);
}

int cilqr_b200_synth_generate(cilqr_handle_t* h, int B, uint64_t first_id, uint64_t seed, int n_tmpl,
                              const cilqr_synth_template_t* tmpls, int keep_yaw) {
    int rc = check_batch(base(h), B);
    if (rc) return rc;
    if (B == 0) return 0;
    if (!tmpls || n_tmpl < 1 || n_tmpl > CILQR_B200_MAX_TEMPLATES) return fail(CILQR_ERR_INVALID, "need 1..%d template descriptors", CILQR_B200_MAX_TEMPLATES);
    return DISPATCH(h, do_synth_generate, B, first_id, seed, n_tmpl, tmpls, keep_yaw);
}

int cilqr_b200_synth_download(cilqr_handle_t* h, int B, double* x0, double* ref_velo, double* borders, int32_t* tmpl,
                              int32_t* n_obs, double* obs) {
    int rc = check_batch(base(h), B);
    if (rc) return rc;
    if (B == 0) return 0;
    return DISPATCH(h, do_synth_download, B, x0, ref_velo, borders, tmpl, n_obs, obs);
}

int cilqr_b200_stage_times(cilqr_handle_t* h, double* ms_out, int32_t* launches_out) {
    if (!h || !ms_out) return fail(CILQR_ERR_INVALID, "NULL argument");
    for (int i = 0; i < 6; ++i) {
        ms_out[i] = base(h)->stage_ms[i];
        if (launches_out) launches_out[i] = base(h)->stage_launches[i];
    }
    return 0;
}

int cilqr_b200_counters(cilqr_handle_t* h, cilqr_counters_t* out) {
    if (!h || !out) return fail(CILQR_ERR_INVALID, "NULL argument");
    *out = base(h)->counters;
    return 0;
}

int cilqr_b200_set_option(cilqr_handle_t* h, int option, int value) {
    if (!h) return fail(CILQR_ERR_INVALID, "handle is NULL");
    return DISPATCH(h, do_set_option, option, value);
}

int cilqr_b200_enable_trace(cilqr_handle_t* h, int cap) {
    if (!h) return fail(CILQR_ERR_INVALID, "handle is NULL");
    return DISPATCH(h, do_enable_trace, cap);
}

int cilqr_b200_get_trace(cilqr_handle_t* h, int B, int32_t* status, int32_t* alpha, double* cost) {
    int rc = check_batch(base(h), B);
    if (rc) return rc;
    if (B == 0) return 0;
    return DISPATCH(h, do_get_trace, B, status, alpha, cost);
}

int cilqr_b200_stage_init(cilqr_handle_t* h, int B, const double* x0, const int32_t* tmpl, int warm,
                          const double* last_u, double* u_out, double* x_out) {
    int rc = check_batch(base(h), B);
    if (rc) return rc;
    if (B == 0) return 0;
    return DISPATCH(h, stage_init, B, x0, tmpl, warm, last_u, u_out, x_out);
}

int cilqr_b200_stage_ref_match(cilqr_handle_t* h, int B, const double* x, const int32_t* tmpl, int32_t* idx_out) {
    int rc = check_batch(base(h), B);
    if (rc) return rc;
    if (B == 0) return 0;
    return DISPATCH(h, stage_ref_match, B, x, tmpl, idx_out);
}

int cilqr_b200_stage_cost(cilqr_handle_t* h, int B, const double* u, const double* x, const double* ref_velo,
                          const double* borders, const int32_t* tmpl, const int32_t* n_obs, const double* obs,
                          int obs_len, const double* alm_mu, const double* alm_rho, double* J_out,
                          double* step_cost_out) {
    int rc = check_batch(base(h), B);
    if (rc) return rc;
    if (B == 0) return 0;
    return DISPATCH(h, stage_cost, B, u, x, ref_velo, borders, tmpl, n_obs, obs, obs_len, alm_mu, alm_rho, J_out, step_cost_out);
}

int cilqr_b200_stage_derivs(cilqr_handle_t* h, int B, const double* u, const double* x, const double* ref_velo,
                            const double* borders, const int32_t* tmpl, const int32_t* n_obs, const double* obs,
                            int obs_len, const double* alm_mu, const double* alm_rho, double* lx, double* lu,
                            double* lxx, double* luu, double* A, double* Bm, double* alm_mu_next) {
    int rc = check_batch(base(h), B);
    if (rc) return rc;
    if (B == 0) return 0;
    return DISPATCH(h, stage_derivs, B, u, x, ref_velo, borders, tmpl, n_obs, obs, obs_len, alm_mu, alm_rho, lx, lu, lxx, luu, A, Bm, alm_mu_next);
}

int cilqr_b200_stage_backward(cilqr_handle_t* h, int B, const double* lx, const double* lu, const double* lxx,
                              const double* luu, const double* A, const double* Bm, const double* lamb,
                              double* d_out, double* K_out, double* dV_out, int32_t* status_out) {
    int rc = check_batch(base(h), B);
    if (rc) return rc;
    if (B == 0) return 0;
    if (!lx || !lu || !lxx || !luu || !A || !Bm || !lamb) return fail(CILQR_ERR_INVALID, "NULL input array");
    return DISPATCH(h, stage_backward, B, lx, lu, lxx, luu, A, Bm, lamb, d_out, K_out, dV_out, status_out);
}

int cilqr_b200_stage_forward(cilqr_handle_t* h, int B, const double* u, const double* x, const double* d,
                             const double* K, const double* alpha, const int32_t* tmpl, double* new_u,
                             double* new_x) {
    int rc = check_batch(base(h), B);
    if (rc) return rc;
    if (B == 0) return 0;
    return DISPATCH(h, stage_forward, B, u, x, d, K, alpha, tmpl, new_u, new_x);
}

int cilqr_b200_bench_backward(cilqr_handle_t* h, int B, double lamb, int reps, int flush_l2, float* ms_out,
                              double* bytes_per_launch) {
    int rc = check_batch(base(h), B);
    if (rc) return rc;
    if (B == 0 || reps <= 0) return fail(CILQR_ERR_INVALID, "need B > 0 and reps > 0");
    return DISPATCH(h, bench_backward, B, lamb, reps, flush_l2, ms_out, bytes_per_launch);
}

int cilqr_b200_bench_tile_records(cilqr_handle_t* h, int B0, int B) {
    int rc = check_batch(base(h), B);
    if (rc) return rc;
    return DISPATCH(h, bench_tile, B0, B);
}

}  // extern "C"
