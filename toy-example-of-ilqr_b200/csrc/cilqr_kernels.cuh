// Batched CILQR kernels for sm_100a.
//
// Memory layout ("step-major SoA"): every per-trajectory array is stored as
// [step][field][batch] with the batch index innermost and a batch stride that is
// a multiple of 128 (the derivative records: [step][tile of 32][field][32], see
// rec_at).  Consecutive lanes of a warp own consecutive trajectories, so every
// global access below is a fully coalesced 128-byte (fp32) or 256-byte (fp64)
// row, whichever stage is running:
//   * step-parallel stages (cost, derivatives): one thread per (trajectory, step);
//   * serial chains (rollouts, Riccati recursion): one thread per trajectory, the
//     4x4 / 4x2 / 2x2 blocks held in registers, A and B in their sparse form
//     (5 + 4 non-trivial entries), V_xx / l_xx symmetric (10 entries); small
//     batches feed the recursion from a shared-memory ring filled by bulk
//     asynchronous copies (k_backward_staged);
//   * waypoint matching: G lanes per trajectory scan a G-wide window of the
//     reference line per probe and pick the first local minimum by ballot
//     (k_ref_match, k_rollout_match), or the rollout thread scans on one by one
//     (k_forward in bandwidth-bound rounds).
// No tensor cores: the largest contraction is 4x4x4.
//
// Line search as a trial pool.  The reference tries alpha = 1, 1/2, ... one after
// the other and keeps the first that passes (cpp:354-372).  Here every searching
// instance claims `count` consecutive slots of a trial pool per round — one slot
// normally, more while it is in a streak of rejected steps — the pool is rolled
// out / matched / costed in bulk, and the verdict kernel walks the instance's
// slots in alpha order and takes the first that passes: same decision, up to
// 20x fewer dependent rounds for the stragglers.
//
// Work lists and repack.  A round touches only the instances still running (a
// list kept in instance order by the verdict kernel), and whenever they are down
// to half of the slots in use they are swapped into a dense prefix of every
// array (k_plan_repack, k_swap_rows) — and back at the end of the solve.
//
// Bandwidth-bound rounds fuse the control half of the derivative records (l_u, l_uu, A, B: functions
// of v, yaw, u alone) into the backward pass (riccati_fused): it is neither stored nor read back.
//
// Look-ahead rounds (small batches, k_adopt).  The next iteration's derivatives and backward
// pass are run one round early on a second stream, as "jobs" next to the cost and verdict
// kernels of the line search they depend on — one per possible outcome of the verdict — and
// the job the verdict asks for is adopted: same operations on the same values, one link of
// the iteration's latency chain off the critical path.
//
// All kernels are grid-stride over device-side counts; the host sizes the grids
// from the last list length it has seen (an upper bound: lists only shrink).
#pragma once

#include "../../include/cilqr_b200.h"
#include "cilqr_model.cuh"

#include <type_traits>

namespace cilqr {

#ifndef CILQR_PARITY
constexpr int kRecFields = 28;  // compact derivative record per step
constexpr int kRecLx = 0;       // l_x   [4]
constexpr int kRecLxx = 4;      // l_xx  sym: 00 01 02 03 11 12 13 22 23 33
constexpr int kRecLu = 14;      // l_u   [2]
constexpr int kRecLuu = 16;     // l_uu  sym: 00 01 11
constexpr int kRecA = 19;       // A     a02 a03 a12 a13 a32
constexpr int kRecB = 24;       // B     b01 b11 b20 b31
constexpr int kVN = 16;         // value-function Hessian carried by the recursion: the full 4x4, as in the reference (see riccati_step)
constexpr int kScPlanes = 1;    // per-step cost rows
#else
// Parity build: l_xx and V_xx are kept as the full 4x4 the reference carries (its products do not
// return bitwise-symmetric matrices, and in ALM mode l_xx itself is not bitwise symmetric), and the
// step costs are kept split into the three sums the reference forms (get_total_cost, cpp:211-213, :286).
constexpr int kRecFields = 34;
constexpr int kRecLx = 0;    // l_x   [4]
constexpr int kRecLxx = 4;   // l_xx  [4][4] row-major
constexpr int kRecLu = 20;   // l_u   [2]
constexpr int kRecLuu = 22;  // l_uu  00 01 11 (both off-diagonals are exactly zero in the reference)
constexpr int kRecA = 25;    // A     a02 a03 a12 a13 a32 (the other entries are exactly 0 / 1)
constexpr int kRecB = 30;    // B     b01 b11 b20 b31
constexpr int kVN = 16;
constexpr int kScPlanes = 4;  // plane 0: step total; 1: state term; 2: control term; 3: constraint terms
#endif

// The records are stored tiled, [step][tile of 32 instances][field][32]: the 28 fields of an
// instance's step sit a constant 32 scalars apart (immediate offsets, no per-field address
// arithmetic), a warp of consecutive instances still reads a fully coalesced row per field, and one
// step's record of a whole tile is one contiguous 28 * 32 * sizeof(T) block — one bulk copy.
constexpr int kRecTile = 32;
// fp64: field c of an instance sits rf(c) = 32 c scalars after its field 0 (a warp reads one 256-byte row per field).
// fp32 (fast build): fields are interleaved in groups of four — [field quad][instance][4] — so that a lane fetches four
// fields of its instance with ONE 128-bit load and a warp still reads contiguous 512-byte rows: the fp32 recursion is
// bound by its instruction stream (ncu: issue slots 55 % active at 76 % of DRAM throughput with 4-byte loads), and this
// quarters its load instructions.  28 fields = 7 quads.  rl() = scalars between consecutive instances of a tile.
template <typename T>
__host__ __device__ constexpr bool rec_quads() {
#ifdef CILQR_PARITY
    return false;
#else
    return sizeof(T) == 4;
#endif
}
template <typename T>
__host__ __device__ constexpr int rl() { return rec_quads<T>() ? 4 : 1; }
template <typename T>
__host__ __device__ constexpr int rf(int c) { return rec_quads<T>() ? (c >> 2) * (4 * kRecTile) + (c & 3) : c * kRecTile; }

constexpr int kNumAlphas = 20;  // alpha = 2^0 .. 2^-19  (cpp:354)
constexpr int kRepackLevels = 10;

enum Phase : int { PH_BACKWARD = 0, PH_SEARCH = 1, PH_DONE = 2 };
enum Status : int { ST_RUNNING = 0, ST_CONVERGED = 1, ST_BWD_FAIL = 2, ST_FWD_FAIL = 3, ST_SMALL_STEP = 4 };
enum ExitReason : int { EX_MAX_ITER = 0, EX_CONVERGED = 1, EX_MAX_LAMB = 2 };
// device-side loop control words
enum Ctl : int {
    CTL_NV = 0, CTL_ACTIVE = 1, CTL_TICKET = 2, CTL_ROUND = 3,
    CTL_NACT = 5,   // [2] entries in the two work lists (round parity)
    CTL_CHUNK = 7,  // next chunk of the work list to hand out in the verdict kernel
    CTL_TALLY = 8,  // [2 words, one 64-bit counter] verdict kernel: blocks done << 48 | trials << 24 | running
    CTL_TRIALS = 10,  // [2 words, one 64-bit counter] line-search trials evaluated over the whole solve
    CTL_NV2 = 12,     // [2] look-ahead solves: trial slots in use, by round parity (the kernels of round r still read
                      // theirs while the verdict kernel of round r re-arms the other one for round r + 1's claims)
    CTL_NSWAP = 16,  // [kRepackLevels] slot pairs exchanged by each repack of the solve
    CTL_WORDS = 32
};
// The tally packs the round's running instances and trials into 24 bits each and the finished blocks into 16
// (one block per 128 work-list entries): cilqr_b200_create refuses batches beyond this.
constexpr int kMaxBatch = (1 << 23) - 128;

// Everything a kernel needs, passed by value.
template <typename T>
struct Dev {
    int N, Bs, Vs, max_obs, alm_cols;
    int wide_mode;     // 0: one alpha per round; 1: adaptive (all remaining alphas while in a rejection streak)
    int wide_step;     // 1: in a streak, a round evaluates a0 + 2 more alphas instead of all that remain (2, 4, 8, 6:
                       // bandwidth-bound rounds, where the 29 % of trials that full widening wastes cost real time)
    int trace_cap;     // iterations recorded per instance (0 = off)
    const DevParams<T>* P;  // [CILQR_B200_MAX_TEMPLATES]
    const T* wx;
    const T* wy;
    const T* wyaw;
    const T* wsin;  // sin / cos of the waypoint yaws, evaluated once per template
    const T* wcos;
    // problem data, stride Bs
    T* ref_velo;  // [Bs]
    T* borders;   // [2][Bs]
    int* tmpl;    // [Bs]
    int* n_obs;   // [Bs]
    T* obs;       // [max_obs][obs_len][4][Bs]  (x, y, sin yaw, cos yaw); step k reads sample obs_off + k
    int obs_len;  // N+1 for a plain solve, the track length in the receding-horizon simulation
    int obs_off;  // simulation tick (utils::get_sub_routing_lines, src/utils.cpp:88-103)
    T* x0;        // [4][Bs]
    // current trajectory of every instance, stride Bs
    T* X;       // [N+1][4][Bs]
    T* U;       // [N][2][Bs]
    int* ridx;  // [N+1][Bs]   matched waypoint per step
    T* sc;      // [N+1][Bs]   per-step cost
    // trial pool, stride Vs
    T* Xt;
    T* Ut;
    int* ridx_t;
    T* sc_t;
    int* t_inst;  // [Vs] owning instance
    int* t_aidx;  // [Vs] alpha index
    int* t_done;  // [Vs] step costs stored so far (self-resetting)
    T* J_t;       // [Vs] total cost of the trial
    int* t_first;     // [Bs] first slot claimed this round
    int* t_count;     // [Bs] slots claimed this round
    int* commit_src;  // [Bs] trial slot to copy into the current trajectory, -1 = none
    // derivative records and gains, stride Bs
    T* rec;  // [N+1][Bs / 32][28][32]  (rec_at)
    T* Kg;   // [N][8][Bs]
    T* dg;   // [N][2][Bs]
    T* dV;   // [2][Bs]
    // solver state per instance
    T* lamb;
    T* J_cur;
    T* J_init;
    T* alpha;  // [Bs] explicit step length for the stage operator
    int* status;
    int* phase;
    int* aidx;
    int* iters;
    int* exit_reason;
    int* rec_valid;
    int* wide;
    // work lists, [2][Bs]: the instances round r has to touch are act[(r & 1) * Bs + 0 .. ctl[CTL_NACT + (r & 1)]),
    // in increasing instance order (so a warp's accesses stay as coalesced as the survivors allow);
    // the verdict kernel of round r filters its list into the one of round r + 1 (stable, single
    // pass: per-chunk counts chained through scan_state with decoupled look-back)
    int* act;
    unsigned long long* scan_state;  // [Bs / 128 + 1]  epoch << 34 | flag << 32 | count
    // repack (large batches): slot pairs (src beyond the prefix, dst a hole inside it), the levels of
    // one solve stacked one after the other (each level has at most half the pairs of the one before)
    int* swap_src;  // [Bs + 8]
    int* swap_dst;
    T* last_u;   // [N][2][Bs]
    int* first;  // [Bs]
    // augmented-Lagrangian state (allocated only when a template asks for it)
    T* mu;       // [N][alm_cols][Bs]
    T* mu_next;  // [N][alm_cols][Bs]
    T* rho;      // [Bs]
    // loop control
    int* ctl;             // [CTL_WORDS]
    volatile int* h_ctl;  // mapped pinned host memory, two 64-bit words: rounds completed << 32 | {instances running, next list length}
    // Look-ahead rounds (latency-bound batches, see k_adopt): the backward pass an instance will need next is run
    // one round early, next to the cost / verdict kernels of the line search it depends on, and adopted when the
    // verdict turns out as expected.
    int spec;        // 1 inside a look-ahead solve
    int pool_base;   // first slot of the half of the trial pool this round uses (0 outside look-ahead solves)
    int pool_cap;    // slots in it (Vs outside look-ahead solves)
    int* gsel;       // [Bs] the instance's current gains: copy 0 / 1 of (Kg, dg, dV: allocated twice), or 2 = still in the
                     // arrays of the trial slot cur_src (adopted last round, copied into copy 0 during this one)
    int round_id;    // the round a launch belongs to (k_adopt: the round its jobs are for)
    int* job_round;  // [Bs] the round the instance's job is for (an instance dropped from the work list keeps a stale one)
    int* job_src;    // [Bs] the trajectory this round's backward job differentiates: -2 no job, -1 the current one, >= 0 a trial slot
    T* job_lamb;     // [Bs] the regularisation the job runs with
    int* job_ok;     // [Bs] the job's recursion met no non-PD Q_uu
    int spec_all_below;  // work lists up to this length give every trial of a line search a job, longer ones only alpha = 1
    int* cur_src;    // [Bs] trial slot (other half of the pool) still holding the current trajectory, -1 once it is copied
    // speculative jobs, one per trial slot: the backward pass that follows if that trial is the accepted one
    int* t_job;      // [Vs] 1: this round differentiates the slot's trajectory and runs the recursion on it
    T* jlamb_t;      // [Vs] with this regularisation
    int* jok_t;      // [Vs] the recursion met no non-PD Q_uu
    T* rec_t;        // [N+1][Vs / 32][28][32]  derivative records of the slots' trajectories (layout of rec)
    T* Kg_t;         // [N][8][Vs]   gains of the slots' jobs
    T* dg_t;         // [N][2][Vs]
    T* dV_t;         // [2][Vs]
    // optional per-iteration trace, [trace_cap][Bs]
    int* tr_status;
    int* tr_alpha;
    T* tr_cost;
};

// A set of trajectories: either the instances' current ones or the trial pool.
template <typename T>
struct View {
    T* X;
    T* U;
    int* ridx;
    T* sc;
    size_t stride;
    const int* inst;  // trial slot -> instance, nullptr for identity
};
template <typename T>
__device__ __forceinline__ View<T> view_of(const Dev<T>& D, int trial) {
    View<T> v;
    if (trial) {
        v.X = D.Xt; v.U = D.Ut; v.ridx = D.ridx_t; v.sc = D.sc_t; v.stride = size_t(D.Vs); v.inst = D.t_inst;
    } else {
        v.X = D.X; v.U = D.U; v.ridx = D.ridx; v.sc = D.sc; v.stride = size_t(D.Bs); v.inst = nullptr;
    }
    return v;
}
// the control word counting the trial slots of the round D.round_id
template <typename T>
__device__ __forceinline__ int nv_index(const Dev<T>& D) { return D.spec ? CTL_NV2 + (D.round_id & 1) : int(CTL_NV); }
template <typename T>
__device__ __forceinline__ int view_count(const Dev<T>& D, int trial, int B) {
    if (!trial) return B;
    int nv = D.ctl[nv_index(D)];
    return nv < D.pool_cap ? nv : D.pool_cap;
}

// field 0 of the record of step k of instance b; field c is rf<T>(c) scalars further on, the same
// instance's step k - 1 is kRecFields * Bs scalars back
template <typename T>
__device__ __forceinline__ T* rec_at(const Dev<T>& D, int k, int b) {
    return D.rec + (size_t(k) * (D.Bs / kRecTile) + size_t(b / kRecTile)) * (kRecFields * kRecTile) + (b % kRecTile) * rl<T>();
}

__device__ __forceinline__ size_t at(size_t stride, int step, int field, int nfields, int i) {
    return (size_t(step) * nfields + field) * stride + i;
}

// The instance's current copy of the gains: 0 outside look-ahead solves (k_adopt flips it when a job is adopted).
template <typename T>
__device__ __forceinline__ int gains_sel(const Dev<T>& D, int b) { return D.spec ? D.gsel[b] : 0; }
// this round's backward job of instance b (look-ahead solves), -2 = none
template <typename T>
__device__ __forceinline__ int job_of(const Dev<T>& D, int b) { return D.job_round[b] == D.round_id ? D.job_src[b] : -2; }
template <typename T>
__device__ __forceinline__ T* Kg_of(const Dev<T>& D, int sel) { return D.Kg + size_t(sel) * D.N * 8 * D.Bs; }
template <typename T>
__device__ __forceinline__ T* dg_of(const Dev<T>& D, int sel) { return D.dg + size_t(sel) * D.N * 2 * D.Bs; }
template <typename T>
__device__ __forceinline__ T* dV_of(const Dev<T>& D, int sel) { return D.dV + size_t(sel) * 2 * D.Bs; }

// Where instance b's current gains are: `*stride` scalars between consecutive rows; K, d, dV point at the instance's
// (or slot's) column.
template <typename T>
struct GainsAt {
    const T* K;
    const T* d;
    const T* dV;
    size_t stride;
};
template <typename T>
__device__ __forceinline__ GainsAt<T> gains_at(const Dev<T>& D, int b) {
    const int g = gains_sel(D, b);
    GainsAt<T> G;
    if (g == 2) {
        const int cs = D.cur_src[b];
        G.K = D.Kg_t + cs;
        G.d = D.dg_t + cs;
        G.dV = D.dV_t + cs;
        G.stride = size_t(D.Vs);
    } else {
        G.K = Kg_of(D, g) + b;
        G.d = dg_of(D, g) + b;
        G.dV = dV_of(D, g) + b;
        G.stride = size_t(D.Bs);
    }
    return G;
}
// field 0 of the record of step k of trial slot v (look-ahead solves; same tiling as rec_at, batch stride Vs)
template <typename T>
__device__ __forceinline__ T* rec_t_at(const Dev<T>& D, int k, int v) {
    return D.rec_t + (size_t(k) * (D.Vs / kRecTile) + size_t(v / kRecTile)) * (kRecFields * kRecTile) + (v % kRecTile) * rl<T>();
}

// The per-template solver scalars (2 KB for all templates) copied into shared memory at the start of a kernel:
// every thread of the step-parallel stages reads a dozen of them behind the load of its instance's template id, and
// from global memory that is one more level of dependent loads on kernels that are bound by exactly that
// (ncu, k_derivs / k_cost at 262144 instances: long-scoreboard stalls, 35-45 % of the warps active).
// Usage: `__shared__ DevParams<T> sP[CILQR_B200_MAX_TEMPLATES]; D.P = stage_params(D, sP);` (D is the kernel's own copy).
template <typename T>
__device__ __forceinline__ const DevParams<T>* stage_params(const Dev<T>& D, DevParams<T>* smem) {
    constexpr int kWords = int(sizeof(DevParams<T>) * CILQR_B200_MAX_TEMPLATES / sizeof(unsigned));
    static_assert(sizeof(DevParams<T>) % sizeof(unsigned) == 0, "word copy");
    const unsigned* src = reinterpret_cast<const unsigned*>(D.P);
    unsigned* dst = reinterpret_cast<unsigned*>(smem);
    const int tid = threadIdx.x + threadIdx.y * blockDim.x, nt = blockDim.x * blockDim.y;
    for (int i = tid; i < kWords; i += nt) dst[i] = src[i];
    __syncthreads();
    return smem;
}

// Total cost of trajectory i from its per-step costs sc [kScPlanes][N+1][stride].  Default build: the step
// totals in step order.  Parity build: the reference's three sums (states, controls, constraints: cpp:211-213,
// :217-286) each in step order, then (states + controls) + constraints.  kCg: read through L2 (the values
// were just written by other threads of the same kernel).
template <typename T, bool kCg>
__device__ __forceinline__ T sum_step_costs(const T* sc, size_t stride, int N, int i) {
    auto ld = [&](size_t idx) { return kCg ? __ldcg(sc + idx) : sc[idx]; };
#ifndef CILQR_PARITY
    T J = 0;
    for (int k = 0; k <= N; ++k) J += ld(size_t(k) * stride + i);
    return J;
#else
    T part[3];
    for (int p = 0; p < 3; ++p) {
        T a = 0;
        const int k0 = p == 2 ? 1 : 0, k1 = p == 1 ? N - 1 : N;
        for (int k = k0; k <= k1; ++k) a += ld((size_t(p + 1) * (N + 1) + k) * stride + i);
        part[p] = a;
    }
    return (part[0] + part[1]) + part[2];
#endif
}

// ---------------------------------------------------------------------------
// K0  initial trajectory: get_init_traj (cpp:155-161, :182-197) or the shifted
//     warm start get_init_traj_increment (cpp:163-180); also resets the
//     per-instance solver state the way solve() does (cpp:88-109).
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) k_init(Dev<T> D, int B, int force_warm, int reset_state) {
    const size_t Bs = D.Bs;
    const int N = D.N;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
        const DevParams<T>& P = D.P[D.tmpl[b]];
        const T p_dt = P.dt, p_wb = P.wheelbase;
        const int p_ref = P.ref_point;
        bool warm = force_warm > 0 || (force_warm < 0 && P.use_last && !D.first[b]);
        T x[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            x[c] = D.x0[size_t(c) * Bs + b];
            D.X[at(Bs, 0, c, 4, b)] = x[c];
        }
        for (int i = 0; i < N; ++i) {
            T a = 0, s = 0;
            if (warm) {
                int src = (i + 1 < N) ? i + 1 : N - 1;
                a = D.last_u[at(Bs, src, 0, 2, b)];
                s = D.last_u[at(Bs, src, 1, 2, b)];
            }
            D.U[at(Bs, i, 0, 2, b)] = a;
            D.U[at(Bs, i, 1, 2, b)] = s;
            T nx[4];
            propagate(x, a, s, p_dt, p_wb, p_ref, nx);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                x[c] = nx[c];
                D.X[at(Bs, i + 1, c, 4, b)] = nx[c];
            }
        }
        if (reset_state) {
            if (P.solve_type == 1 && D.mu && (!P.use_last || D.first[b])) {  // cpp:88-93
                D.rho[b] = P.alm_rho_init;
                for (int i = 0; i < N * D.alm_cols; ++i) {
                    D.mu[size_t(i) * Bs + b] = 0;
                    D.mu_next[size_t(i) * Bs + b] = 0;
                }
            }
            D.first[b] = 0;
            D.status[b] = ST_RUNNING;
            D.lamb[b] = P.init_lamb;
            D.iters[b] = 0;
            D.phase[b] = P.max_iter > 0 ? PH_BACKWARD : PH_DONE;
            D.aidx[b] = 0;
            D.rec_valid[b] = 0;
            D.wide[b] = 0;
            D.exit_reason[b] = EX_MAX_ITER;
            D.commit_src[b] = -1;
            D.t_count[b] = 0;
            D.dV[b] = 0;
            D.dV[Bs + b] = 0;
            if (D.gsel) {
                D.gsel[b] = 0;
                D.job_src[b] = -2;
                D.job_round[b] = -1;
                D.job_ok[b] = 0;
                D.cur_src[b] = -1;
            }
            // round 0 works on every instance
            D.act[b] = b;
            if (b == 0) {
                D.ctl[CTL_NACT + 0] = B;
                D.ctl[CTL_NACT + 1] = 0;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// K1  get_ref_exact_points (cpp:289-314).  For each step the reference scans
//     j = start, start+1, ... and stops at the first j whose successor is not
//     strictly closer (or at the last waypoint); the next step starts there.
//     Equivalent formulation used here: first j >= start with
//     !(dist[j+1] < dist[j]) (NaN stops the scan, as in the reference); distances
//     are compared squared (no hypot on the serial chain; the two orderings can
//     differ only for waypoints equidistant to within an ulp).
//     G lanes per trajectory evaluate a window of G waypoints per probe.
// ---------------------------------------------------------------------------
// squared distance to a waypoint, with the operation order spelled out: every kernel that scans
// (windowed or one by one) must compare the same bits
template <typename T>
__device__ __forceinline__ T wp_dist2(T px, T py, T wx, T wy) {
#if defined(CILQR_PARITY) || defined(CILQR_EXPERIMENT_HYPOT)
    return m_hypot(px - wx, py - wy);  // the reference compares hypot() values (cpp:300-309)
#else
    const T ex = px - wx, ey = py - wy;
    return m_fma(ex, ex, ey * ey);
#endif
}
// the same scan, one waypoint at a time (inside the throughput-regime rollout, one thread per trial)
template <typename T>
__device__ __forceinline__ int match_from(const T* __restrict__ wx, const T* __restrict__ wy, int M, int start, T px, T py) {
    int j = start;
    T dj = wp_dist2(px, py, __ldg(wx + j), __ldg(wy + j));
    while (j + 1 < M) {
        const T dn = wp_dist2(px, py, __ldg(wx + j + 1), __ldg(wy + j + 1));
        if (!(dn < dj)) break;
        ++j;
        dj = dn;
    }
    return j;
}

template <typename T, int G>
__global__ void __launch_bounds__(128) k_ref_match(Dev<T> D, int B, int trial) {
    const View<T> V = view_of(D, trial);
    const int count = view_count(D, trial, B);
    const int lane = threadIdx.x & 31;
    const int sub = lane % G;
    const int grp_shift = lane - sub;  // first lane of my group inside the warp
    const unsigned grp_mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << grp_shift);
    const int groups_per_warp = 32 / G;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    for (int base = warp * groups_per_warp; base < count; base += n_warps * groups_per_warp) {
        const int v = base + lane / G;
        const bool live = v < count;
        const int vv = live ? v : count - 1;  // keep addresses valid for idle groups
        const int b = V.inst ? V.inst[vv] : vv;
        const DevParams<T>& P = D.P[D.tmpl[b]];
        const int M = P.wp_len;
        const T* wx = D.wx + P.wp_off;
        const T* wy = D.wy + P.wp_off;
        int start = 0;
        // the position of step k+1 is fetched while step k scans (the scan is a serial chain)
        T npx = V.X[at(V.stride, 0, 0, 4, vv)];
        T npy = V.X[at(V.stride, 0, 1, 4, vv)];
        for (int k = 0; k <= D.N; ++k) {
            const T px = npx, py = npy;
            {
                const int kn = k < D.N ? k + 1 : k;
                npx = ld_early(V.X + at(V.stride, kn, 0, 4, vv));
                npy = ld_early(V.X + at(V.stride, kn, 1, 4, vv));
            }
            int found = -1;
            bool done = !live;
            // all groups of the warp iterate together; finished groups idle
            while (!__all_sync(0xffffffffu, done)) {
                int j = start + sub;
                int jc = j < M ? j : M - 1;
                // squared distance: the scan only compares distances, and x -> sqrt(x) is monotone
                T dj = wp_dist2(px, py, __ldg(wx + jc), __ldg(wy + jc));
                T dn = __shfl_down_sync(0xffffffffu, dj, 1, G);
                bool stop = !done && (sub < G - 1) && (j + 1 >= M || !(dn < dj));
                unsigned m = __ballot_sync(0xffffffffu, stop) & grp_mask;
                if (!done) {
                    if (m) {
                        found = start + (__ffs(m) - 1 - grp_shift);
                        done = true;
                    } else {
                        start += G - 1;
                    }
                }
            }
            if (live) {
                if (sub == 0) V.ridx[size_t(k) * V.stride + v] = found;
                start = found;
            }
        }
    }
}

// do the active lanes of the warp hold the same value?
__device__ __forceinline__ bool same_in_warp(int v) {
    int pred;
    __match_all_sync(__activemask(), v, &pred);
    return pred != 0;
}

// acc up/lo, steer up/lo in the reference's order (cpp:222-228)
template <typename T>
__device__ __forceinline__ void ctrl_constraints(const DevParams<T>& P, T acc, T steer, T c[4]) {
    c[0] = acc - P.acc_max;
    c[1] = P.acc_min - acc;
    c[2] = steer - P.stl_lim;
    c[3] = -P.stl_lim - steer;
}

// ---------------------------------------------------------------------------
// K2  get_total_cost (cpp:199-287), one thread per (trajectory, step k):
//     state term of x_k, control term of u_k (k < N), constraint terms of
//     step k (k >= 1: u_{k-1}, x_k, ref_k, obstacles at tick k).
// ---------------------------------------------------------------------------
// kMinBlocks: 8 CTAs/SM (64 registers, small spills) in the throughput regime, where the kernel is
// fp64-latency bound and more resident warps pay (+11 % whole-solve at B = 262 144); 4 (128 registers, four
// obstacles in flight per trip) for latency-bound batches, where occupancy is irrelevant.
// cost of step k of trajectory v of a view (instance b), with waypoint match ri
// kAlm: the batch may hold augmented-Lagrangian instances; false compiles the ALM paths out, which
// leaves the barrier terms free of branches (independent exponentials interleave).
template <typename T, bool kAlm, int kOb>
__device__ __forceinline__ T step_cost_of(const Dev<T>& D, const View<T>& V, int b, int v, int k, int ri, T* parts = nullptr) {
    const int N = D.N;
    const size_t Bs = D.Bs;
    const DevParams<T>& P = D.P[D.tmpl[b]];
    T x[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) x[c] = V.X[at(V.stride, k, c, 4, v)];
    const T rx = D.wx[P.wp_off + ri], ry = D.wy[P.wp_off + ri], ryaw = D.wyaw[P.wp_off + ri];
    const T ref[4] = {rx, ry, D.ref_velo[b], ryaw};
    T cost = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        T e = x[c] - ref[c];
        cost += e * P.Q[c] * e;
    }
    if (parts) parts[0] = cost, parts[1] = 0, parts[2] = 0;
    if (k < N) {
        T a = V.U[at(V.stride, k, 0, 2, v)], s = V.U[at(V.stride, k, 1, 2, v)];
        T ce = a * P.R[0] * a;
        ce += s * P.R[1] * s;
        cost += ce;
        if (parts) parts[1] = ce;
    }
    if (k >= 1) {
        T a = V.U[at(V.stride, k - 1, 0, 2, v)], s = V.U[at(V.stride, k - 1, 1, 2, v)];
        T c[8];
        ctrl_constraints(P, a, s, c);
        c[4] = x[2] - P.velo_max;
        c[5] = P.velo_min - x[2];
        T d_sign, hyp;
        T cur_d = lateral_offset(x[0], x[1], rx, ry, D.wsin[P.wp_off + ri], D.wcos[P.wp_off + ri], &d_sign, &hyp);
        c[6] = cur_d - (D.borders[b] - P.width / 2);
        c[7] = (D.borders[Bs + b] + P.width / 2) - cur_d;
        T Jk = 0;
        const bool alm = kAlm && P.solve_type == 1;
        const T rho = alm ? D.rho[b] : T(0);
        const T* mu = alm ? D.mu + size_t(k - 1) * D.alm_cols * Bs + b : nullptr;
        if (!alm) {
#pragma unroll
            for (int m = 0; m < 8; ++m) Jk += exp_barrier(c[m], P.st_q1, P.st_q2);
        } else {
#pragma unroll
            for (int m = 0; m < 8; ++m) Jk += alm_item(c[m], rho, mu[size_t(m) * Bs]);
        }
        const int no = D.n_obs[b];
        if (no > 0) {
            EgoCircles<T> e = ego_circles(x, P.wheelbase, P.ref_point);
            // Q obstacles per trip: their loads are issued together and their barrier terms are
            // independent, so the exponentials interleave; the sums keep the reference's order
            auto trip = [&](int j, auto qc) {
                constexpr int Q = decltype(qc)::value;
                T item[2 * Q];
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    const int jj = j + q < no ? j + q : j;
                    // obstacle sample (x, y, sin yaw, cos yaw): the sin/cos of the input yaw is
                    // evaluated once per upload (k_obs_sincos), not once per cost evaluation
                    const T* ob = D.obs + (size_t(jj) * D.obs_len + D.obs_off + k) * 4 * Bs + b;
                    const T ox = ob[0], oy = ob[Bs], so = ob[2 * Bs], co = ob[3 * Bs];
                    T cf = ellipse_margin<T, false>(e.fx, e.fy, ox, oy, so, co, P.ell_a2, P.ell_b2, nullptr, nullptr);
                    T cr = ellipse_margin<T, false>(e.rx, e.ry, ox, oy, so, co, P.ell_a2, P.ell_b2, nullptr, nullptr);
                    if (!alm) {
                        item[2 * q] = exp_barrier(cf, P.obs_q1, P.obs_q2);
                        item[2 * q + 1] = exp_barrier(cr, P.obs_q1, P.obs_q2);
                    } else {
                        item[2 * q] = alm_item(cf, rho, mu[size_t(8 + 2 * jj) * Bs]);
                        item[2 * q + 1] = alm_item(cr, rho, mu[size_t(9 + 2 * jj) * Bs]);
                    }
                }
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    if (j + q < no) {
                        Jk += item[2 * q];
                        Jk += item[2 * q + 1];
                    }
                }
            };
            if (kOb == 2 && same_in_warp(no)) {
                // throughput regime (the kernel is bound by its fp64 instructions): whole pairs, then a last single
                // obstacle on its own instead of a pair with a discarded duplicate (3 obstacles: 6 instead of 8 barrier
                // terms) — where the warp agrees on the count; a mixed warp would serialise the extra trip
                int j = 0;
                for (; j + 2 <= no; j += 2) trip(j, std::integral_constant<int, 2>());
                if (j < no) trip(j, std::integral_constant<int, 1>());
            } else {
                // latency regime: one trip for up to kOb obstacles (the padding of the last trip is idle issue slots there)
                for (int j = 0; j < no; j += kOb) trip(j, std::integral_constant<int, kOb>());
            }
        }
        cost += Jk;
        if (parts) parts[2] = Jk;
    }
    return cost;
}

// trial > 0: the trial pool.  trial == 1 also totals each trial's step costs in this kernel (the thread
// that stores a trial's last step cost sums them in step order: one launch less per round, for
// latency-bound batches); trial == 2 leaves the totals to k_sum_trials (bandwidth-bound batches: no
// fence and no atomic per thread).
template <typename T, int kMinBlocks, bool kAlm>
__global__ void __launch_bounds__(128, kMinBlocks) k_cost(Dev<T> D, int B, int trial) {
    __shared__ DevParams<T> sP[CILQR_B200_MAX_TEMPLATES];
    D.P = stage_params(D, sP);
    const View<T> V = view_of(D, trial);
    const int count = view_count(D, trial, B);
    const int N = D.N;
    // grid: x = step, y = blocks of trajectories (so the blocks that have work are dispatched first
    // when only the head of the trial pool is in use)
    const int k = blockIdx.x;
    const int v_first = trial ? D.pool_base : 0, v_end = v_first + count;
    for (int v = v_first + blockIdx.y * blockDim.x + threadIdx.x; v < v_end; v += gridDim.y * blockDim.x) {
        const int b = V.inst ? V.inst[v] : v;
#ifdef CILQR_PARITY
        T parts[3];
        const T cost = step_cost_of<T, kAlm, (kMinBlocks <= 4 ? 4 : 2)>(D, V, b, v, k, V.ridx[size_t(k) * V.stride + v], parts);
        for (int p = 0; p < 3; ++p) V.sc[(size_t(p + 1) * (N + 1) + k) * V.stride + v] = parts[p];
#else
        const T cost = step_cost_of<T, kAlm, (kMinBlocks <= 4 ? 4 : 2)>(D, V, b, v, k, V.ridx[size_t(k) * V.stride + v]);
#endif
        V.sc[size_t(k) * V.stride + v] = cost;
        if (trial == 1) {
            // the thread that stores the last step cost of a trial sums them in step order (fixed
            // order: the total is deterministic whichever thread ends up doing it)
            __threadfence();
            if (atomicAdd(&D.t_done[v], 1) == N) {
                __threadfence();
                D.J_t[v] = sum_step_costs<T, true>(V.sc, V.stride, N, v);
                D.t_done[v] = 0;
            }
        }
    }
}

// Total cost of every trial of the pool, step costs summed in step order (same bits as k_cost's own sum).
template <typename T>
__global__ void __launch_bounds__(128) k_sum_trials(Dev<T> D, int B) {
    const int count = view_count(D, 1, B);
    const size_t Vs = D.Vs;
    for (int v = D.pool_base + blockIdx.x * blockDim.x + threadIdx.x; v < D.pool_base + count; v += gridDim.x * blockDim.x) {
        D.J_t[v] = sum_step_costs<T, false>(D.sc_t, Vs, D.N, v);
    }
}

// J = sum of the step costs of the current trajectory.  mode 0: every instance,
// sets J_cur and J_init (after the initial rollout); mode 1: instances whose
// multipliers just changed (ALM, cpp:342 recomputes ori_cost every iter_step).
template <typename T>
__global__ void __launch_bounds__(128) k_sum_cost(Dev<T> D, int B, int mode) {
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
        if (mode == 1 && (D.phase[b] != PH_BACKWARD || D.rec_valid[b])) continue;
        const T J = sum_step_costs<T, false>(D.sc, D.Bs, D.N, b);
        D.J_cur[b] = J;
        if (mode == 0) D.J_init[b] = J;
    }
}

// One constraint's gradient / Gauss-Newton Hessian weight: returns (g, h) with
// grad += g * c_dot and hess += h * c_dot c_dot^T.  Barrier: g = q2*b, h = q2^2*b
// (cpp:692-699); ALM: g = rho*(c + mu/rho) if that is > 0 else 0, h = g (cpp:701-713).
template <typename T>
__device__ __forceinline__ void constraint_weights(bool alm, T c, T q1, T q2, T rho, T mu, T* g, T* h) {
    if (!alm) {
        T bv = exp_barrier(c, q1, q2);
        *g = q2 * bv;
        *h = (q2 * q2) * bv;
    } else {
        T t = c + mu / rho;
        T w = (t > 0) ? rho * t : T(0);
        *g = w;
        *h = w;
    }
}

#ifdef CILQR_PARITY
// Parity build: one constraint's contribution formed exactly as the reference forms it — barrier
// cpp:692-699 (grad += q2 b c_dot, hess += q2^2 b c_dot c_dot^T), ALM cpp:701-713 (b_dot = rho (c + mu/rho) c_dot,
// hess += b_dot c_dot^T) — over all n entries of c_dot, zeros included.
template <typename T, int n>
__device__ __forceinline__ void add_constraint_ref(bool alm, T c, const T* c_dot, T q1, T q2, T rho, T mu, T* grad, T* hess) {
    if (!alm) {
        const T b = exp_barrier(c, q1, q2);
        const T q2sq = q2 * q2;
#pragma unroll
        for (int r = 0; r < n; ++r) grad[r] += q2 * b * c_dot[r];
#pragma unroll
        for (int r = 0; r < n; ++r)
#pragma unroll
            for (int cc = 0; cc < n; ++cc) hess[r * n + cc] += q2sq * b * (c_dot[r] * c_dot[cc]);
    } else if ((c + mu / rho) > 0) {
        T bd[n];
#pragma unroll
        for (int r = 0; r < n; ++r) bd[r] = rho * (c + mu / rho) * c_dot[r];
#pragma unroll
        for (int r = 0; r < n; ++r) grad[r] += bd[r];
#pragma unroll
        for (int r = 0; r < n; ++r)
#pragma unroll
            for (int cc = 0; cc < n; ++cc) hess[r * n + cc] += bd[r] * c_dot[cc];
    }
}
#endif

// ---------------------------------------------------------------------------
// K3 + K4  get_total_cost_derivatives_and_Hessians (cpp:463-690) and
//     get_kinematic_model_derivatives (src/utils.cpp:285-342), one thread per
//     (instance, step k), writing the compact record of step k:
//     l_x[k], l_xx[k] (constraints of x_k if k >= 1), and for k < N
//     l_u[k], l_uu[k] (constraints of u_k, i.e. the reference's step k+1), A_k, B_k.
//     In the solver (masked != 0) the same thread first commits an accepted
//     trial into the current trajectory ("x = new_x; u = new_u", cpp:113-116),
//     then differentiates only where the record is stale — the reference's
//     cache rule for rejected steps (cpp:469-474).
// ---------------------------------------------------------------------------
#ifndef CILQR_PARITY
// The control half of a step's record in the barrier solve type — l_u, l_uu (constraints of u_k: the reference's step
// k + 1) and the model Jacobians A_k, B_k — from (v_k, yaw_k, u_k): out[0..13] in record order (l_u 2, l_uu 3, A 5, B 4).
// Shared by the derivative kernel (which stores it) and the fused backward pass (which does not).
template <typename T>
__device__ __forceinline__ void control_fields(const DevParams<T>& P, T velo, T yaw, T ua, T us, T* out) {
    T c[4];
    ctrl_constraints(P, ua, us, c);
    T g[4], h[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) constraint_weights(false, c[m], P.st_q1, P.st_q2, T(0), T(0), &g[m], &h[m]);
    // (zA / zS: the zero entries of c_dot, as in the state half: an overflowed acceleration barrier
    // poisons the steering entries and vice versa, both poison the off-diagonal)
    const T zA = (g[0] + g[1]) * T(0), zS = (g[2] + g[3]) * T(0);
    const T zAh = (h[0] + h[1]) * T(0), zSh = (h[2] + h[3]) * T(0);
    const T gu0 = (g[0] + (-g[1])) + zS;
    const T gu1 = (g[2] + (-g[3])) + zA;
    const T hu0 = (h[0] + h[1]) + zSh;
    const T hu1 = (h[2] + h[3]) + zAh;
    out[0] = 2 * (ua * P.R[0]) + gu0;
    out[1] = 2 * (us * P.R[1]) + gu1;
    out[2] = 2 * P.R[0] + hu0;
    out[3] = zAh + zSh;
    out[4] = 2 * P.R[1] + hu1;
    model_jacobians(velo, yaw, us, P.dt, P.wheelbase, P.ref_point, out + 5, out + 10);
}
#endif

// The work of one thread of the derivative stage: half `part` (0 = state terms l_x, l_xx; 1 = control terms and
// model Jacobians l_u, l_uu, A, B) of step k of instance b.  masked != 0 (inside the solver): first commit an
// accepted trial, then differentiate only where the record is stale.
template <typename T, bool kAlm, bool kPairs = true>
__device__ __forceinline__ void derivs_item(const Dev<T>& D, int b, int k, int part, int masked, int slot = -1) {
    const int N = D.N;
    const size_t Bs = D.Bs, Vs = D.Vs;

    T x[4], ua = 0, us = 0;
    int ri = 0;
    // src: trial slot to commit; tsrc: trial slot to differentiate (-1 = the current trajectory's own arrays).
    // Look-ahead rounds.  masked == 2: the slot to commit is the one k_adopt took over from the verdict kernel, and what
    // is differentiated (into the instance's record) is the instance's own job: a trial of its running line search, or
    // the current trajectory where its record is stale.  masked == 3: the speculative job of trial slot `slot` (owned
    // by instance b): its trajectory, into the slot's own record.
    int src = -1, tsrc = -1;
    bool diff = true;
    if (masked == 1 || masked == 4) {
        // (4: a round whose backward pass computes the control half of the records itself — only the state half of
        // this stage is launched, and it copies the accepted controls as well)
        src = tsrc = D.commit_src[b];
    } else if (masked == 5) {
        // the control half of the records that are still valid, after rounds that did not store it (first
        // latency-regime round of a solve that began in the bandwidth regime)
        if (!D.rec_valid[b] || D.commit_src[b] >= 0) return;
    } else if (masked == 2) {
        src = tsrc = D.cur_src[b];
        const int job = job_of(D, b);
        diff = job >= 0 || (job == -1 && !D.rec_valid[b]);
        if (job >= 0) tsrc = job;  // the instance's own job speculates on a trial of its running line search
        if (src >= 0 && D.gsel[b] == 2) {
            // the slot was adopted together with its job: its record and gains become the instance's (copy 0)
            const T* rs = rec_t_at(D, k, src);
            T* rd = rec_at(D, k, b);
            if (part == 0) {
#pragma unroll
                for (int c = kRecLx; c < kRecLu; ++c) rd[rf<T>(c)] = rs[rf<T>(c)];
                if (k < N) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) D.Kg[at(Bs, k, c, 8, b)] = D.Kg_t[at(Vs, k, c, 8, src)];
                }
            } else if (k < N) {
#pragma unroll
                for (int c = kRecLu; c < kRecFields; ++c) rd[rf<T>(c)] = rs[rf<T>(c)];
                D.dg[at(Bs, k, 0, 2, b)] = D.dg_t[at(Vs, k, 0, 2, src)];
                D.dg[at(Bs, k, 1, 2, b)] = D.dg_t[at(Vs, k, 1, 2, src)];
                if (k == 0) {
                    D.dV[b] = D.dV_t[src];
                    D.dV[Bs + b] = D.dV_t[Vs + src];
                }
            }
        }
    } else if (masked == 3) {
        tsrc = slot;
    }
    if (src >= 0) {
        // the accepted trial becomes the current trajectory; part 1 reads the state from the
        // trial slot too, so it never races with part 0's copy
#pragma unroll
        for (int c = 0; c < 4; ++c) x[c] = D.Xt[at(Vs, k, c, 4, src)];
        if (part == 0) {
#pragma unroll
            for (int c = 0; c < 4; ++c) D.X[at(Bs, k, c, 4, b)] = x[c];
            ri = D.ridx_t[size_t(k) * Vs + src];
            D.ridx[size_t(k) * Bs + b] = ri;
            for (int p = 0; p < kScPlanes; ++p)
                D.sc[(size_t(p) * (N + 1) + k) * Bs + b] = D.sc_t[(size_t(p) * (N + 1) + k) * Vs + src];
            if (masked == 4 && k < N) {
                D.U[at(Bs, k, 0, 2, b)] = D.Ut[at(Vs, k, 0, 2, src)];
                D.U[at(Bs, k, 1, 2, b)] = D.Ut[at(Vs, k, 1, 2, src)];
            }
        } else if (k < N) {
            ua = D.Ut[at(Vs, k, 0, 2, src)];
            us = D.Ut[at(Vs, k, 1, 2, src)];
            D.U[at(Bs, k, 0, 2, b)] = ua;
            D.U[at(Bs, k, 1, 2, b)] = us;
        }
    }
    if ((masked == 1 || masked == 4) && (D.phase[b] != PH_BACKWARD || D.rec_valid[b])) return;
    if (masked == 2 && !diff) return;
    if (tsrc < 0) {
#pragma unroll
        for (int c = 0; c < 4; ++c) x[c] = D.X[at(Bs, k, c, 4, b)];
        if (part == 0) {
            ri = D.ridx[size_t(k) * Bs + b];
        } else if (k < N) {
            ua = D.U[at(Bs, k, 0, 2, b)];
            us = D.U[at(Bs, k, 1, 2, b)];
        }
    } else if (tsrc != src) {
#pragma unroll
        for (int c = 0; c < 4; ++c) x[c] = D.Xt[at(Vs, k, c, 4, tsrc)];
        if (part == 0) {
            ri = D.ridx_t[size_t(k) * Vs + tsrc];
        } else if (k < N) {
            ua = D.Ut[at(Vs, k, 0, 2, tsrc)];
            us = D.Ut[at(Vs, k, 1, 2, tsrc)];
        }
    }
    const DevParams<T>& P = D.P[D.tmpl[b]];
    const bool alm = kAlm && P.solve_type == 1;
    const T rho = alm ? D.rho[b] : T(0);
    T* rec = masked == 3 ? rec_t_at(D, k, slot) : rec_at(D, k, b);
    if (part == 0) {
    const T rx = D.wx[P.wp_off + ri], ry = D.wy[P.wp_off + ri], ryaw = D.wyaw[P.wp_off + ri];
    const T ref[4] = {rx, ry, D.ref_velo[b], ryaw};
#ifdef CILQR_PARITY
    // the reference's accumulation, literally (cpp:497-689): velocity up / lo, border up / lo into the
    // row, then per obstacle (front + rear) summed first and added to the row, the prime part last
    T gx[4] = {0, 0, 0, 0};
    T Hx[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (k >= 1) {
        const T* mu = alm ? D.mu + size_t(k - 1) * D.alm_cols * Bs + b : nullptr;
        T* mun = alm ? D.mu_next + size_t(k - 1) * D.alm_cols * Bs + b : nullptr;
        T d_sign, hyp;
        const T cur_d = lateral_offset(x[0], x[1], rx, ry, D.wsin[P.wp_off + ri], D.wcos[P.wp_off + ri], &d_sign, &hyp);
        const T cc4[4] = {x[2] - P.velo_max, P.velo_min - x[2], cur_d - (D.borders[b] - P.width / 2),
                          (D.borders[Bs + b] + P.width / 2) - cur_d};
        T cx[4][4] = {{0, 0, 1, 0}, {0, 0, -1, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};
        cx[2][0] = (x[0] - rx) / hyp;
        cx[2][1] = (x[1] - ry) / hyp;
        if (d_sign < 0)
            for (int r = 0; r < 4; ++r) cx[2][r] = -1 * cx[2][r];
        for (int r = 0; r < 4; ++r) cx[3][r] = -1 * cx[2][r];
#pragma unroll
        for (int m = 0; m < 4; ++m)
            add_constraint_ref<T, 4>(alm, cc4[m], cx[m], P.st_q1, P.st_q2, rho, alm ? mu[size_t(4 + m) * Bs] : T(0), gx, Hx);
        if (alm)
            for (int m = 0; m < 4; ++m)
                mun[size_t(4 + m) * Bs] = std_min(std_max(mu[size_t(4 + m) * Bs] + rho * cc4[m], T(0)), P.max_mu);
        const int no = D.n_obs[b];
        if (no > 0) {
            const EgoCircles<T> e = ego_circles(x, P.wheelbase, P.ref_point);
            for (int j = 0; j < no; ++j) {
                const T* ob = D.obs + (size_t(j) * D.obs_len + D.obs_off + k) * 4 * Bs + b;
                const T ox = ob[0], oy = ob[Bs], so = ob[2 * Bs], co = ob[3 * Bs];
                T gfx, gfy, grx, gry;
                const T cf = ellipse_margin<T, true>(e.fx, e.fy, ox, oy, so, co, P.ell_a2, P.ell_b2, &gfx, &gfy);
                const T cr = ellipse_margin<T, true>(e.rx, e.ry, ox, oy, so, co, P.ell_a2, P.ell_b2, &grx, &gry);
                // 4x2 centre Jacobians times the point gradients (utils.cpp:363-385, cpp:728-736)
                const T gf[4] = {T(1) * gfx + T(0) * gfy, T(0) * gfx + T(1) * gfy, T(0) * gfx + T(0) * gfy,
                                 e.jf0 * gfx + e.jf1 * gfy};
                const T gr[4] = {T(1) * grx + T(0) * gry, T(0) * grx + T(1) * gry, T(0) * grx + T(0) * gry,
                                 e.jr0 * grx + e.jr1 * gry};
                T g2[4] = {0, 0, 0, 0};
                T H2[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
                add_constraint_ref<T, 4>(alm, cf, gf, P.obs_q1, P.obs_q2, rho, alm ? mu[size_t(8 + 2 * j) * Bs] : T(0), g2, H2);
                add_constraint_ref<T, 4>(alm, cr, gr, P.obs_q1, P.obs_q2, rho, alm ? mu[size_t(9 + 2 * j) * Bs] : T(0), g2, H2);
                for (int r = 0; r < 4; ++r) gx[r] += g2[r];
                for (int r = 0; r < 16; ++r) Hx[r] += H2[r];
                if (alm) {
                    mun[size_t(8 + 2 * j) * Bs] = std_min(std_max(mu[size_t(8 + 2 * j) * Bs] + rho * cf, T(0)), P.max_mu);
                    mun[size_t(9 + 2 * j) * Bs] = std_min(std_max(mu[size_t(9 + 2 * j) * Bs] + rho * cr, T(0)), P.max_mu);
                }
            }
        }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) rec[rf<T>(kRecLx + c)] = 2 * (x[c] - ref[c]) * P.Q[c] + gx[c];
#pragma unroll
    for (int c = 0; c < 4; ++c) Hx[c * 4 + c] = 2 * P.Q[c] + Hx[c * 4 + c];
#pragma unroll
    for (int c = 0; c < 16; ++c) rec[rf<T>(kRecLxx + c)] = Hx[c];
#else
    T gx[4] = {0, 0, 0, 0};
    T H[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // 00 01 02 03 11 12 13 22 23 33
    if (k >= 1) {
        const T* mu = alm ? D.mu + size_t(k - 1) * D.alm_cols * Bs + b : nullptr;
        T* mun = alm ? D.mu_next + size_t(k - 1) * D.alm_cols * Bs + b : nullptr;
        // velocity bounds: c_dot = (0,0,+-1,0)
        T cv[2] = {x[2] - P.velo_max, P.velo_min - x[2]};
        // zV / zB / zO: what the zero entries of a constraint's c_dot contribute in the reference, which forms
        // every product (cpp:696-698): 0 for a finite barrier weight, NaN for one that overflowed (0 * inf).
        // Summed per constraint family and added below to the entries that family does not otherwise touch,
        // so that an overflowing barrier poisons the same entries as in the reference.
        T zV = 0, zB = 0, zO = 0;     // gradient side (weights g)
        T zVh = 0, zBh = 0, zOh = 0;  // Hessian side (weights h = q2 g can overflow where g has not)
        T g, h;
        constraint_weights(alm, cv[0], P.st_q1, P.st_q2, rho, alm ? mu[size_t(4) * Bs] : T(0), &g, &h);
        gx[2] += g;
        H[7] += h;
        zV += g;
        zVh += h;
        constraint_weights(alm, cv[1], P.st_q1, P.st_q2, rho, alm ? mu[size_t(5) * Bs] : T(0), &g, &h);
        gx[2] += -g;
        H[7] += h;
        zV += g;
        zVh += h;
        // road borders: c_dot = +-(px-rx, py-ry)/hypot, flipped when d_sign < 0 (cpp:527-533)
        T d_sign, hyp;
        T cur_d = lateral_offset(x[0], x[1], rx, ry, D.wsin[P.wp_off + ri], D.wcos[P.wp_off + ri], &d_sign, &hyp);
        T cp[2] = {cur_d - (D.borders[b] - P.width / 2), (D.borders[Bs + b] + P.width / 2) - cur_d};
        T n0 = (x[0] - rx) / hyp, n1 = (x[1] - ry) / hyp;
        if (d_sign < 0) {
            n0 = -n0;
            n1 = -n1;
        }
        constraint_weights(alm, cp[0], P.st_q1, P.st_q2, rho, alm ? mu[size_t(6) * Bs] : T(0), &g, &h);
        gx[0] += g * n0;
        gx[1] += g * n1;
        H[0] += h * (n0 * n0);
        H[1] += h * (n0 * n1);
        H[4] += h * (n1 * n1);
        zB += g;
        zBh += h;
        constraint_weights(alm, cp[1], P.st_q1, P.st_q2, rho, alm ? mu[size_t(7) * Bs] : T(0), &g, &h);
        gx[0] += g * (-n0);
        gx[1] += g * (-n1);
        H[0] += h * (n0 * n0);
        H[1] += h * (n0 * n1);
        H[4] += h * (n1 * n1);
        zB += g;
        zBh += h;
        if (alm) {
            mun[size_t(4) * Bs] = std_min(std_max(mu[size_t(4) * Bs] + rho * cv[0], T(0)), P.max_mu);
            mun[size_t(5) * Bs] = std_min(std_max(mu[size_t(5) * Bs] + rho * cv[1], T(0)), P.max_mu);
            mun[size_t(6) * Bs] = std_min(std_max(mu[size_t(6) * Bs] + rho * cp[0], T(0)), P.max_mu);
            mun[size_t(7) * Bs] = std_min(std_max(mu[size_t(7) * Bs] + rho * cp[1], T(0)), P.max_mu);
        }
        // obstacles: front and rear circle against each ellipse (cpp:648-664)
        const int no = D.n_obs[b];
        if (no > 0) {
            EgoCircles<T> e = ego_circles(x, P.wheelbase, P.ref_point);
            // two obstacles per trip (independent exponentials interleave); accumulation in the
            // reference's order
            // (whole pairs, then a last single obstacle on its own: no discarded duplicate)
            auto trip = [&](int j, auto qc) {
                constexpr int Q = decltype(qc)::value;
                T tg[Q][3], tH[Q][6];
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    const int jj = j + q;
                    // obstacle sample (x, y, sin yaw, cos yaw): the sin/cos of the input yaw is
                    // evaluated once per upload (k_obs_sincos), not once per cost evaluation
                    const T* ob = D.obs + (size_t(jj) * D.obs_len + D.obs_off + k) * 4 * Bs + b;
                    const T ox = ob[0], oy = ob[Bs], so = ob[2 * Bs], co = ob[3 * Bs];
                    T gfx, gfy, grx, gry;
                    T cf = ellipse_margin<T, true>(e.fx, e.fy, ox, oy, so, co, P.ell_a2, P.ell_b2, &gfx, &gfy);
                    T cr = ellipse_margin<T, true>(e.rx, e.ry, ox, oy, so, co, P.ell_a2, P.ell_b2, &grx, &gry);
                    // chain through the 4x2 centre Jacobians: (gx, gy, 0, yaw row)
                    T f3 = e.jf0 * gfx + e.jf1 * gfy;
                    T r3 = e.jr0 * grx + e.jr1 * gry;
                    T gf, hf, gr, hr;
                    constraint_weights(alm, cf, P.obs_q1, P.obs_q2, rho, alm ? mu[size_t(8 + 2 * jj) * Bs] : T(0), &gf, &hf);
                    constraint_weights(alm, cr, P.obs_q1, P.obs_q2, rho, alm ? mu[size_t(9 + 2 * jj) * Bs] : T(0), &gr, &hr);
                    // front + rear first, then into the row (cpp:662-664)
                    tg[q][0] = gf * gfx + gr * grx;
                    tg[q][1] = gf * gfy + gr * gry;
                    tg[q][2] = gf * f3 + gr * r3;
                    tH[q][0] = hf * (gfx * gfx) + hr * (grx * grx);
                    tH[q][1] = hf * (gfx * gfy) + hr * (grx * gry);
                    tH[q][2] = hf * (gfx * f3) + hr * (grx * r3);
                    tH[q][3] = hf * (gfy * gfy) + hr * (gry * gry);
                    tH[q][4] = hf * (gfy * f3) + hr * (gry * r3);
                    tH[q][5] = hf * (f3 * f3) + hr * (r3 * r3);
                    zO += gf + gr;
                    zOh += hf + hr;
                    if (alm) {
                        mun[size_t(8 + 2 * jj) * Bs] =
                            std_min(std_max(mu[size_t(8 + 2 * jj) * Bs] + rho * cf, T(0)), P.max_mu);
                        mun[size_t(9 + 2 * jj) * Bs] =
                            std_min(std_max(mu[size_t(9 + 2 * jj) * Bs] + rho * cr, T(0)), P.max_mu);
                    }
                }
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    gx[0] += tg[q][0];
                    gx[1] += tg[q][1];
                    gx[3] += tg[q][2];
                    H[0] += tH[q][0];
                    H[1] += tH[q][1];
                    H[3] += tH[q][2];
                    H[4] += tH[q][3];
                    H[6] += tH[q][4];
                    H[9] += tH[q][5];
                }
            };
            // (no padded pair for a mixed warp here, unlike step_cost_of: a third copy of the body costs this kernel more
            // in spills than the extra trip costs the mixed warps of C3 — measured)
            if (kPairs) {
                int j = 0;
                for (; j + 2 <= no; j += 2) trip(j, std::integral_constant<int, 2>());
                if (j < no) trip(j, std::integral_constant<int, 1>());
            } else {
                // bandwidth-bound rounds in fp64: one obstacle per trip — the kernel is short of registers, not of
                // independent work (408 -> 356 bytes of spills), and a warp of mixed obstacle counts (C3) wastes no half
                // trips: ~1 % of a C1 / C2 solve; fp32 keeps the pairs (C4, 5 obstacles: 454 against 466 ms)
                for (int j = 0; j < no; ++j) trip(j, std::integral_constant<int, 1>());
            }
        }
        // velocity terms have c_dot = (0, 0, +-1, 0), border terms (n0, n1, 0, 0), obstacle terms (gx, gy, 0, g3)
        zV *= T(0);
        zB *= T(0);
        zO *= T(0);
        zVh *= T(0);
        zBh *= T(0);
        zOh *= T(0);
        gx[0] += zV;
        gx[1] += zV;
        gx[2] += zB + zO;
        gx[3] += zV + zB;
        const T zVB = zVh + zBh, zAll = zVB + zOh;
        H[0] += zVh;        // 00
        H[1] += zVh;        // 01
        H[2] += zAll;       // 02
        H[3] += zVB;        // 03
        H[4] += zVh;        // 11
        H[5] += zAll;       // 12
        H[6] += zVB;        // 13
        H[7] += zBh + zOh;  // 22
        H[8] += zAll;       // 23
        H[9] += zVB;        // 33
    }
    // prime objective: l_x = 2 (x - ref) Q, l_xx = 2 Q (cpp:493-494), summed with the constraint part
#pragma unroll
    for (int c = 0; c < 4; ++c) rec[rf<T>(kRecLx + c)] = 2 * (x[c] - ref[c]) * P.Q[c] + gx[c];
    H[0] += 2 * P.Q[0];
    H[4] += 2 * P.Q[1];
    H[7] += 2 * P.Q[2];
    H[9] += 2 * P.Q[3];
#pragma unroll
    for (int c = 0; c < 10; ++c) rec[rf<T>(kRecLxx + c)] = H[c];
#endif
    }  // part 0

    if (part == 1 && k < N) {
        T c[4];
        ctrl_constraints(P, ua, us, c);
        const T* mu = alm ? D.mu + size_t(k) * D.alm_cols * Bs + b : nullptr;
#ifdef CILQR_PARITY
        const T cu[4][2] = {{1, 0}, {-1, 0}, {0, 1}, {0, -1}};
        T gu[2] = {0, 0}, Hu[4] = {0, 0, 0, 0};
#pragma unroll
        for (int m = 0; m < 4; ++m)
            add_constraint_ref<T, 2>(alm, c[m], cu[m], P.st_q1, P.st_q2, rho, alm ? mu[size_t(m) * Bs] : T(0), gu, Hu);
        if (alm) {
            T* mun = D.mu_next + size_t(k) * D.alm_cols * Bs + b;
            for (int m = 0; m < 4; ++m)
                mun[size_t(m) * Bs] = std_min(std_max(mu[size_t(m) * Bs] + rho * c[m], T(0)), P.max_mu);
        }
        rec[rf<T>(kRecLu + 0)] = 2 * (ua * P.R[0]) + gu[0];
        rec[rf<T>(kRecLu + 1)] = 2 * (us * P.R[1]) + gu[1];
        rec[rf<T>(kRecLuu + 0)] = 2 * P.R[0] + Hu[0];
        rec[rf<T>(kRecLuu + 1)] = Hu[1];
        rec[rf<T>(kRecLuu + 2)] = 2 * P.R[1] + Hu[3];
#else
        if (!alm) {
            T f[14];
            control_fields(P, x[2], x[3], ua, us, f);
#pragma unroll
            for (int c2 = 0; c2 < 14; ++c2) rec[rf<T>(kRecLu + c2)] = f[c2];
        } else {
        T g[4], h[4];
#pragma unroll
        for (int m = 0; m < 4; ++m)
            constraint_weights(alm, c[m], P.st_q1, P.st_q2, rho, alm ? mu[size_t(m) * Bs] : T(0), &g[m], &h[m]);
        // (zA / zS: the zero entries of c_dot, as in the state half: an overflowed acceleration barrier
        // poisons the steering entries and vice versa, both poison the off-diagonal)
        const T zA = (g[0] + g[1]) * T(0), zS = (g[2] + g[3]) * T(0);
        const T zAh = (h[0] + h[1]) * T(0), zSh = (h[2] + h[3]) * T(0);
        T gu0 = (g[0] + (-g[1])) + zS;
        T gu1 = (g[2] + (-g[3])) + zA;
        T hu0 = (h[0] + h[1]) + zSh;
        T hu1 = (h[2] + h[3]) + zAh;
        if (alm) {
            T* mun = D.mu_next + size_t(k) * D.alm_cols * Bs + b;
#pragma unroll
            for (int m = 0; m < 4; ++m)
                mun[size_t(m) * Bs] = std_min(std_max(mu[size_t(m) * Bs] + rho * c[m], T(0)), P.max_mu);
        }
        rec[rf<T>(kRecLu + 0)] = 2 * (ua * P.R[0]) + gu0;
        rec[rf<T>(kRecLu + 1)] = 2 * (us * P.R[1]) + gu1;
        rec[rf<T>(kRecLuu + 0)] = 2 * P.R[0] + hu0;
        rec[rf<T>(kRecLuu + 1)] = zAh + zSh;
        rec[rf<T>(kRecLuu + 2)] = 2 * P.R[1] + hu1;
#endif
        T ja[5], jb[4];
        model_jacobians(x[2], x[3], us, P.dt, P.wheelbase, P.ref_point, ja, jb);
#pragma unroll
        for (int c2 = 0; c2 < 5; ++c2) rec[rf<T>(kRecA + c2)] = ja[c2];
#pragma unroll
        for (int c2 = 0; c2 < 4; ++c2) rec[rf<T>(kRecB + c2)] = jb[c2];
#ifndef CILQR_PARITY
        }  // alm
#endif
    }
}

template <typename T, int kPart, bool kAlm>
// CTAs per SM of the two halves in the throughput regime: the state half at 6 (80 registers, 156 B of spills)
// beats 8 (64 registers, 276 B) and 4 (128, none): 262144 instances 243.1 -> 234.7 ms; the control half stays at 8.
#ifndef CILQR_DERIVS0_MINB
#define CILQR_DERIVS0_MINB 6
#endif
#ifndef CILQR_DERIVS1_MINB
#define CILQR_DERIVS1_MINB 8
#endif
__global__ void __launch_bounds__(128, kPart < 0 ? 4 : (kPart == 0 ? CILQR_DERIVS0_MINB : CILQR_DERIVS1_MINB)) k_derivs(Dev<T> D, int B, int masked, int par, int slot_y0 = 0) {
    __shared__ DevParams<T> sP[CILQR_B200_MAX_TEMPLATES];
    // (blocks with nothing to do leave before staging the parameters)
    if (masked == 2 && int(blockIdx.y) >= slot_y0 && (int(blockIdx.y) - slot_y0) * int(blockDim.x) >= view_count(D, 1, B)) return;
    if (masked == 2 && int(blockIdx.y) < slot_y0 && int(blockIdx.y) * int(blockDim.x) >= D.ctl[CTL_NACT + par]) return;
    D.P = stage_params(D, sP);
    const size_t Bs = D.Bs;
    // two independent halves per (instance, step): part 0 = state terms (l_x, l_xx), part 1 = control
    // terms and model Jacobians (l_u, l_uu, A, B).  kPart < 0: one launch, the half taken from
    // blockIdx.y (latency-bound batches: one launch less per round).  kPart = 0 / 1: one launch per
    // half, so that each half gets its own register budget and occupancy (throughput regime).
    // grid: x = step (and half), y = blocks of work-list entries
    const int k = kPart < 0 ? int(blockIdx.x >> 1) : int(blockIdx.x);
    const int part = kPart < 0 ? int(blockIdx.x & 1) : kPart;
    // solver: only the instances on this round's work list (running, or with a step to commit)
    const int* list = masked ? D.act + size_t(par) * Bs : nullptr;
    const int n = masked ? D.ctl[CTL_NACT + par] : B;
    if (masked == 2 && int(blockIdx.y) >= slot_y0) {
        // look-ahead rounds: the blocks from slot_y0 on take the speculative jobs of this round's trial slots
        const int nv = view_count(D, 1, B);
        for (int i = (blockIdx.y - slot_y0) * blockDim.x + threadIdx.x; i < nv; i += (gridDim.y - slot_y0) * blockDim.x) {
            const int v = D.pool_base + i;
            if (D.t_job[v]) derivs_item<T, kAlm>(D, D.t_inst[v], k, part, 3, v);
        }
        return;
    }
    const int ny = masked == 2 ? slot_y0 : int(gridDim.y);
    for (int idx = blockIdx.y * blockDim.x + threadIdx.x; idx < n; idx += ny * blockDim.x)
        derivs_item<T, kAlm, (kPart < 0 || sizeof(T) == 4)>(D, list ? list[idx] : idx, k, part, masked);
}

// solve()'s bookkeeping after an iter_step (cpp:113-141): lambda schedule,
// iteration count, the three exits.  alpha_idx / cost feed the optional trace.
template <typename T>
__device__ __forceinline__ void end_iteration(const Dev<T>& D, const DevParams<T>& P, int b, int status,
                                              int alpha_idx, T cost) {
    T lamb = D.lamb[b];
    if (status == ST_BWD_FAIL || status == ST_FWD_FAIL) {
        lamb = std_max(P.lamb_amplify, lamb * P.lamb_amplify);
    } else if (status == ST_RUNNING) {
        lamb *= P.lamb_decay;
    }
    D.lamb[b] = lamb;
    D.status[b] = status;
    const int it = D.iters[b];
    if (it < D.trace_cap) {
        D.tr_status[size_t(it) * D.Bs + b] = status;
        D.tr_alpha[size_t(it) * D.Bs + b] = alpha_idx;
        D.tr_cost[size_t(it) * D.Bs + b] = cost;
    }
    D.iters[b] = it + 1;
    D.aidx[b] = 0;
    if (lamb > P.max_lamb) {
        D.phase[b] = PH_DONE;
        D.exit_reason[b] = EX_MAX_LAMB;
    } else if (status == ST_CONVERGED) {
        D.phase[b] = PH_DONE;
        D.exit_reason[b] = EX_CONVERGED;
    } else if (it + 1 >= P.max_iter) {
        D.phase[b] = PH_DONE;
        D.exit_reason[b] = EX_MAX_ITER;
    } else {
        D.phase[b] = PH_BACKWARD;
    }
}

// V_xx at the horizon = l_xx[N] (cpp:395-396), from the record's l_xx fields (offsets rf<T>)
template <typename T>
__device__ __forceinline__ void load_terminal_V(const T* rec, T* V) {
#ifdef CILQR_PARITY
#pragma unroll
    for (int c = 0; c < 16; ++c) V[c] = rec[rf<T>(kRecLxx + c)];
#else
    int e = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = r; c < 4; ++c, ++e) {
            const T v = rec[rf<T>(kRecLxx + e)];
            V[r * 4 + c] = v;
            V[c * 4 + r] = v;
        }
#endif
}

template <typename T>
__device__ __forceinline__ constexpr T kEps();
template <>
__device__ __forceinline__ constexpr double kEps<double>() { return 2.220446049250313e-16; }
template <>
__device__ __forceinline__ constexpr float kEps<float>() { return 1.1920929e-07f; }

// One step of the recursion: consumes the 28-scalar record r of step i and the value function (Vx, V) of
// step i+1, produces the gains K (2x4), d (2), and overwrites (Vx, V) with step i's.  Returns false —
// leaving Vx, V, dV untouched — when Q_uu + lambda*I fails the LLT test.  Shared by every variant of
// the backward kernel, so they all return the same bits.
#ifdef CILQR_PARITY
// Parity build: the same step as the reference writes it (cpp:398-437) — dense 4x4 / 4x2 products, inner sums
// over k = 0..3 from zero, (A^T V) A and (B^T V) A / B associated from the left, the full (not symmetrised)
// V_xx, Eigen::LLT's test sequence, the adjugate inverse, the three-term value update.
template <typename T>
__device__ __forceinline__ bool riccati_step(const T* r, T lamb, T* Vx, T* V, T& dV0, T& dV1, T* K, T& d0, T& d1) {
    T A[16], B[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) A[i] = (i % 5 == 0) ? T(1) : T(0);
#pragma unroll
    for (int i = 0; i < 8; ++i) B[i] = 0;
    A[0 * 4 + 2] = r[kRecA + 0];
    A[0 * 4 + 3] = r[kRecA + 1];
    A[1 * 4 + 2] = r[kRecA + 2];
    A[1 * 4 + 3] = r[kRecA + 3];
    A[3 * 4 + 2] = r[kRecA + 4];
    B[0 * 2 + 1] = r[kRecB + 0];
    B[1 * 2 + 1] = r[kRecB + 1];
    B[2 * 2 + 0] = r[kRecB + 2];
    B[3 * 2 + 1] = r[kRecB + 3];
    const T luu[4] = {r[kRecLuu + 0], r[kRecLuu + 1], r[kRecLuu + 1], r[kRecLuu + 2]};
    T Qx[4], Qu[2], Qxx[16], Quu[4], Qux[8], AtV[16], BtV[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        T s = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) s += A[k * 4 + i] * Vx[k];
        Qx[i] = r[kRecLx + i] + s;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        T s = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) s += B[k * 2 + i] * Vx[k];
        Qu[i] = r[kRecLu + i] + s;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            T s = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) s += A[k * 4 + i] * V[k * 4 + c];
            AtV[i * 4 + c] = s;
        }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            T s = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) s += B[k * 2 + i] * V[k * 4 + c];
            BtV[i * 4 + c] = s;
        }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            T s = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) s += AtV[i * 4 + k] * A[k * 4 + c];
            Qxx[i * 4 + c] = r[kRecLxx + i * 4 + c] + s;
        }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            T s = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) s += BtV[i * 4 + k] * B[k * 2 + c];
            Quu[i * 2 + c] = (luu[i * 2 + c] + s) + lamb * (i == c ? T(1) : T(0));
        }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            T s = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) s += BtV[i * 4 + k] * A[k * 4 + c];
            Qux[i * 4 + c] = T(0) + s;
        }
    {
        const T a00 = Quu[0];
        if (a00 <= T(0)) return false;
        const T l00 = m_sqrt(a00);
        const T l10 = Quu[2] / l00;
        const T a11 = Quu[3] - l10 * l10;
        if (a11 <= T(0)) return false;
    }
    const T det = Quu[0] * Quu[3] - Quu[2] * Quu[1];
    const T invdet = T(1) / det;
    const T inv[4] = {Quu[3] * invdet, -Quu[1] * invdet, -Quu[2] * invdet, Quu[0] * invdet};
    T d[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) d[i] = (-inv[i * 2 + 0]) * Qu[0] + (-inv[i * 2 + 1]) * Qu[1];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) K[i * 4 + c] = (-inv[i * 2 + 0]) * Qux[0 * 4 + c] + (-inv[i * 2 + 1]) * Qux[1 * 4 + c];
    d0 = d[0];
    d1 = d[1];
    T KtQuu[8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 2; ++c) KtQuu[i * 2 + c] = K[0 * 4 + i] * Quu[0 * 2 + c] + K[1 * 4 + i] * Quu[1 * 2 + c];
    T nVx[4], nV[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const T t1 = KtQuu[i * 2 + 0] * d[0] + KtQuu[i * 2 + 1] * d[1];
        const T t2 = K[0 * 4 + i] * Qu[0] + K[1 * 4 + i] * Qu[1];
        const T t3 = Qux[0 * 4 + i] * d[0] + Qux[1 * 4 + i] * d[1];
        nVx[i] = ((Qx[i] + t1) + t2) + t3;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const T t1 = KtQuu[i * 2 + 0] * K[0 * 4 + c] + KtQuu[i * 2 + 1] * K[1 * 4 + c];
            const T t2 = K[0 * 4 + i] * Qux[0 * 4 + c] + K[1 * 4 + i] * Qux[1 * 4 + c];
            const T t3 = Qux[0 * 4 + i] * K[0 * 4 + c] + Qux[1 * 4 + i] * K[1 * 4 + c];
            nV[i * 4 + c] = ((Qxx[i * 4 + c] + t1) + t2) + t3;
        }
#pragma unroll
    for (int i = 0; i < 4; ++i) Vx[i] = nVx[i];
#pragma unroll
    for (int i = 0; i < 16; ++i) V[i] = nV[i];
    const T hd0 = T(0.5) * d[0], hd1 = T(0.5) * d[1];
    const T r0 = hd0 * Quu[0] + hd1 * Quu[2];
    const T r1 = hd0 * Quu[1] + hd1 * Quu[3];
    dV0 += r0 * d[0] + r1 * d[1];
    dV1 += d[0] * Qu[0] + d[1] * Qu[1];
    return true;
}
#else
// The value-function Hessian is carried as the full 4x4 and updated with the reference's three-term form WITHOUT
// symmetrising it, although in exact arithmetic it is symmetric: the reference's recursion amplifies the
// antisymmetric part of V_xx (rounding noise to begin with) by roughly rho(A - BK) rho(A + BK) per step, and on long
// horizons with stiff barrier terms (N >= 100: BASELINE configs C2, C4) that mode grows until Q_uu fails the
// positive-definiteness test — backward_pass returns BACKWARD_PASS_FAIL and solve() raises lambda (cpp:415-420,
// :118-120).  A symmetric recursion (10 entries) is cheaper and numerically better behaved, but it sails through
// where the reference fails (measured: tests/test_gpu_truth_bound.py), i.e. it is a different algorithm on exactly
// the configs the benchmark names.  Same products as the reference, A and B in their sparse form, sums in the
// reference's order; FMA contraction is the only liberty taken.
template <typename T>
__device__ __forceinline__ bool riccati_step(const T* r, T lamb, T* Vx, T* V, T& dV0, T& dV1, T* K, T& d0, T& d1) {
    const T a02 = r[kRecA + 0], a03 = r[kRecA + 1], a12 = r[kRecA + 2], a13 = r[kRecA + 3], a32 = r[kRecA + 4];
    const T b01 = r[kRecB + 0], b11 = r[kRecB + 1], b20 = r[kRecB + 2], b31 = r[kRecB + 3];
    // P = A^T V: rows 0 and 1 are V's, rows 2 and 3 mix in columns 2 and 3 of A (A = I + {02, 03, 12, 13, 32})
    T P2[4], P3[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        P2[c] = a02 * V[0 * 4 + c] + a12 * V[1 * 4 + c] + V[2 * 4 + c] + a32 * V[3 * 4 + c];
        P3[c] = a03 * V[0 * 4 + c] + a13 * V[1 * 4 + c] + V[3 * 4 + c];
    }
    // Q_xx = l_xx + P A (l_xx symmetric, upper triangle in the record)
    T Qxx[16];
    {
        const T* P[4] = {V, V + 4, P2, P3};
        const int sym[16] = {0, 1, 2, 3, 1, 4, 5, 6, 2, 5, 7, 8, 3, 6, 8, 9};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const T* p = P[i];
            Qxx[i * 4 + 0] = r[kRecLxx + sym[i * 4 + 0]] + p[0];
            Qxx[i * 4 + 1] = r[kRecLxx + sym[i * 4 + 1]] + p[1];
            Qxx[i * 4 + 2] = r[kRecLxx + sym[i * 4 + 2]] + (a02 * p[0] + a12 * p[1] + p[2] + a32 * p[3]);
            Qxx[i * 4 + 3] = r[kRecLxx + sym[i * 4 + 3]] + (a03 * p[0] + a13 * p[1] + p[3]);
        }
    }
    // Q_x = l_x + A^T V_x ; Q_u = l_u + B^T V_x
    T Qx[4];
    Qx[0] = r[kRecLx + 0] + Vx[0];
    Qx[1] = r[kRecLx + 1] + Vx[1];
    Qx[2] = r[kRecLx + 2] + (a02 * Vx[0] + a12 * Vx[1] + Vx[2] + a32 * Vx[3]);
    Qx[3] = r[kRecLx + 3] + (a03 * Vx[0] + a13 * Vx[1] + Vx[3]);
    const T Qu0 = r[kRecLu + 0] + b20 * Vx[2];
    const T Qu1 = r[kRecLu + 1] + (b01 * Vx[0] + b11 * Vx[1] + b31 * Vx[3]);
    // G = B^T V (2x4): row 0 = b20 V[2][.], row 1 = b01 V[0][.] + b11 V[1][.] + b31 V[3][.]
    const T G00 = b20 * V[8], G01 = b20 * V[9], G02 = b20 * V[10], G03 = b20 * V[11];
    const T G10 = b01 * V[0] + b11 * V[4] + b31 * V[12];
    const T G11 = b01 * V[1] + b11 * V[5] + b31 * V[13];
    const T G12 = b01 * V[2] + b11 * V[6] + b31 * V[14];
    const T G13 = b01 * V[3] + b11 * V[7] + b31 * V[15];
    // Q_ux = G A (2x4)
    T Qux[8];
    Qux[0] = G00;
    Qux[1] = G01;
    Qux[2] = a02 * G00 + a12 * G01 + G02 + a32 * G03;
    Qux[3] = a03 * G00 + a13 * G01 + G03;
    Qux[4] = G10;
    Qux[5] = G11;
    Qux[6] = a02 * G10 + a12 * G11 + G12 + a32 * G13;
    Qux[7] = a03 * G10 + a13 * G11 + G13;
    // Q_uu = l_uu + G B + lambda I  (both off-diagonals kept, as the reference computes them)
    const T Quu00 = (r[kRecLuu + 0] + G02 * b20) + lamb;
    const T Quu01 = r[kRecLuu + 1] + (G00 * b01 + G01 * b11 + G03 * b31);
    const T Quu10 = r[kRecLuu + 1] + G12 * b20;
    const T Quu11 = (r[kRecLuu + 2] + (G10 * b01 + G11 * b11 + G13 * b31)) + lamb;
    // LLT positive-definiteness test (Eigen::LLT, lower, unblocked: fail iff a00 <= 0 or
    // a11 - (a10 / sqrt(a00))^2 <= 0; NaN passes), as a predicate: nothing is stored when it fails.
    // The sqrt / divide sequence is ~40 instructions on a serial chain that is bound by its
    // instruction stream, and the test passes by a wide margin on almost every step, so a
    // sufficient condition is tried first: in floating point (a10 / sqrt(a00))^2 is
    // a10^2 / a00 within 5 roundings, so a00 > 0 and a00 a11 - a10^2 > 64 eps a10^2 guarantees that
    // the exact sequence passes too.  Anything else (near-singular, non-positive, NaN, inf) takes the
    // exact sequence, so the verdict is always the reference's.
    bool not_pd = false;
    {
        const T a10sq = Quu10 * Quu10;
        const bool surely_pd = Quu00 > T(0) && (Quu00 * Quu11 - a10sq) > T(64) * kEps<T>() * a10sq;
        if (!surely_pd) {
            const T l10 = Quu10 / m_sqrt(Quu00);
            not_pd = (Quu00 <= T(0)) || (Quu11 - l10 * l10 <= T(0));
        }
    }
    const T invdet = T(1) / (Quu00 * Quu11 - Quu10 * Quu01);
    const T i00 = Quu11 * invdet, i01 = -Quu01 * invdet, i10 = -Quu10 * invdet, i11 = Quu00 * invdet;
    d0 = (-i00) * Qu0 + (-i01) * Qu1;
    d1 = (-i10) * Qu0 + (-i11) * Qu1;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        K[c] = (-i00) * Qux[c] + (-i01) * Qux[4 + c];
        K[4 + c] = (-i10) * Qux[c] + (-i11) * Qux[4 + c];
    }
    if (not_pd) return false;
    // value function update (cpp:427-432), regularised Q_uu
    T M0[4], M1[4];  // K^T Q_uu, columns 0 and 1
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        M0[c] = K[c] * Quu00 + K[4 + c] * Quu10;
        M1[c] = K[c] * Quu01 + K[4 + c] * Quu11;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        T t1 = M0[c] * d0 + M1[c] * d1;
        T t2 = K[c] * Qu0 + K[4 + c] * Qu1;
        T t3 = Qux[c] * d0 + Qux[4 + c] * d1;
        Vx[c] = ((Qx[c] + t1) + t2) + t3;
    }
#pragma unroll
    for (int rr = 0; rr < 4; ++rr)
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            T t1 = M0[rr] * K[cc] + M1[rr] * K[4 + cc];
            T t2 = K[rr] * Qux[cc] + K[4 + rr] * Qux[4 + cc];
            T t3 = Qux[rr] * K[cc] + Qux[4 + rr] * K[4 + cc];
            V[rr * 4 + cc] = ((Qxx[rr * 4 + cc] + t1) + t2) + t3;
        }
    // expected cost reduction (cpp:435-436)
    const T h0 = T(0.5) * d0, h1 = T(0.5) * d1;
    dV0 += (h0 * Quu00 + h1 * Quu10) * d0 + (h0 * Quu01 + h1 * Quu11) * d1;
    dV1 += d0 * Qu0 + d1 * Qu1;
    return true;
}

#endif

// The 28 (34) fields of one record into registers, from global or shared memory (p = the instance's field 0).
// fp32 quad layout: seven 128-bit loads.  kEarly: issued where written (software prefetch, see ld_early).
__device__ __forceinline__ float4 ld_early4(const float* p) {
    float4 v;
    asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
template <typename T, bool kEarly>
__device__ __forceinline__ void load_record(const T* p, T* r) {
    if constexpr (rec_quads<T>()) {
        static_assert(!rec_quads<T>() || kRecFields % 4 == 0, "whole quads");
#pragma unroll
        for (int q = 0; q < kRecFields / 4; ++q) {
            const float4 v = kEarly ? ld_early4(p + q * (4 * kRecTile)) : *reinterpret_cast<const float4*>(p + q * (4 * kRecTile));
            r[4 * q + 0] = v.x;
            r[4 * q + 1] = v.y;
            r[4 * q + 2] = v.z;
            r[4 * q + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int c = 0; c < kRecFields; ++c) r[c] = kEarly ? ld_early(p + rf<T>(c)) : p[rf<T>(c)];
    }
}

// The Riccati recursion of one trajectory.  Streams the 28-scalar record of each
// step (l_x 4, l_xx 10, l_u 2, l_uu 3, A 5, B 4) and writes K (8) and d (2):
// (38 N + 18) * sizeof(T) algorithmic bytes per trajectory.  Q_uu + lambda*I is
// tested exactly like Eigen::LLT (lower, unblocked): fail iff a pivot <= 0, NaN
// passes (cpp:415-420); the inverse is the adjugate times 1/det (cpp:421).
// Returns false on a non-PD Q_uu (d, K rows not reached are zeroed, as in the reference).
#ifndef CILQR_PARITY
// Fused flavour (bandwidth-bound rounds, barrier solve type): the control half of every record (l_u, l_uu, A, B: 14 of
// the 28 fields) is computed here from (v, yaw, u) of the current trajectory instead of being written by the
// derivative kernel and read back — 32 instead of 112 bytes read per step in fp64, no write at all — with the next
// step's operands prefetched one step ahead like the plain register-prefetch flavour.  Same entry formulas
// (control_fields), same recursion (riccati_step).
template <typename T>
__device__ __forceinline__ bool riccati_fused(const Dev<T>& D, const DevParams<T>& P, int b, T lamb) {
    const int N = D.N;
    const size_t Bs = D.Bs;
    static_assert(kRecLu == 14, "state half first");
    const T* rec = rec_at(D, N, b);
    T Vx[4], V[kVN];
#pragma unroll
    for (int c = 0; c < 4; ++c) Vx[c] = rec[rf<T>(kRecLx + c)];
    load_terminal_V(rec, V);
    T dV0 = 0, dV1 = 0;
    bool failed = false;
    int i = N - 1;
    // operands of step i: the state half of its record, v_i, yaw_i, u_i
    T nxt[18];
    auto fetch = [&](int step, T* o) {
        const T* rp = rec_at(D, step, b);
        if constexpr (rec_quads<T>()) {
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const float4 v = ld_early4(rp + q * (4 * kRecTile));
                o[4 * q + 0] = v.x, o[4 * q + 1] = v.y, o[4 * q + 2] = v.z, o[4 * q + 3] = v.w;
            }
            o[12] = ld_early(rp + rf<T>(12));
            o[13] = ld_early(rp + rf<T>(13));
        } else {
#pragma unroll
            for (int c = 0; c < 14; ++c) o[c] = ld_early(rp + rf<T>(c));
        }
        o[14] = ld_early(D.X + at(Bs, step, 2, 4, b));
        o[15] = ld_early(D.X + at(Bs, step, 3, 4, b));
        o[16] = ld_early(D.U + at(Bs, step, 0, 2, b));
        o[17] = ld_early(D.U + at(Bs, step, 1, 2, b));
    };
    fetch(i, nxt);
    for (; i >= 0; --i) {
        T r[kRecFields];
#pragma unroll
        for (int c = 0; c < 14; ++c) r[c] = nxt[c];
        const T velo = nxt[14], yaw = nxt[15], ua = nxt[16], us = nxt[17];
        fetch(i > 0 ? i - 1 : 0, nxt);
        control_fields(P, velo, yaw, ua, us, r + 14);
        T K[8], d0, d1;
        if (!riccati_step(r, lamb, Vx, V, dV0, dV1, K, d0, d1)) {
            failed = true;
            break;
        }
        D.dg[at(Bs, i, 0, 2, b)] = d0;
        D.dg[at(Bs, i, 1, 2, b)] = d1;
#pragma unroll
        for (int c = 0; c < 8; ++c) D.Kg[at(Bs, i, c, 8, b)] = K[c];
    }
    if (failed) {
        for (; i >= 0; --i) {
            D.dg[at(Bs, i, 0, 2, b)] = 0;
            D.dg[at(Bs, i, 1, 2, b)] = 0;
#pragma unroll
            for (int c = 0; c < 8; ++c) D.Kg[at(Bs, i, c, 8, b)] = 0;
        }
    }
    D.dV[b] = dV0;
    D.dV[Bs + b] = dV1;
    return !failed;
}
#endif

template <typename T, bool kPrefetch>
__device__ __forceinline__ bool riccati(const Dev<T>& D, int b, T lamb) {
    const int N = D.N;
    const size_t Bs = D.Bs;
    const T* rec = rec_at(D, N, b);
    T Vx[4], V[kVN];
#pragma unroll
    for (int c = 0; c < 4; ++c) Vx[c] = rec[rf<T>(kRecLx + c)];
    load_terminal_V(rec, V);
    T dV0 = 0, dV1 = 0;
    bool failed = false;
    int i = N - 1;
    T nxt[kRecFields];
    if (kPrefetch) load_record<T, true>(rec - size_t(kRecFields) * Bs, nxt);
    for (; i >= 0; --i) {
        rec -= size_t(kRecFields) * Bs;
        T r[kRecFields];
        if (kPrefetch) {
            // record of step i was fetched during step i+1; fetch step i-1 now (latency-bound batches)
#pragma unroll
            for (int c = 0; c < kRecFields; ++c) r[c] = nxt[c];
            load_record<T, true>(rec - size_t(i > 0 ? kRecFields : 0) * Bs, nxt);
        } else {
            load_record<T, false>(rec, r);
        }
        T K[8], d0, d1;
        if (!riccati_step(r, lamb, Vx, V, dV0, dV1, K, d0, d1)) {
            failed = true;
            break;
        }
        D.dg[at(Bs, i, 0, 2, b)] = d0;
        D.dg[at(Bs, i, 1, 2, b)] = d1;
#pragma unroll
        for (int c = 0; c < 8; ++c) D.Kg[at(Bs, i, c, 8, b)] = K[c];
    }
    if (failed) {
        for (; i >= 0; --i) {
            D.dg[at(Bs, i, 0, 2, b)] = 0;
            D.dg[at(Bs, i, 1, 2, b)] = 0;
#pragma unroll
            for (int c = 0; c < 8; ++c) D.Kg[at(Bs, i, c, 8, b)] = 0;
        }
    }
    D.dV[b] = dV0;
    D.dV[Bs + b] = dV1;
    return !failed;
}

// Warp-collective: every lane asks for `want` consecutive trial-pool slots for its instance b
// (alphas a0, a0+1, ...); exclusive prefix sum over the warp, one atomic for the warp's total.
template <typename T>
__device__ __forceinline__ void claim_slots(const Dev<T>& D, int b, int want, int a0, int lane) {
    int incl = want;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    int total = __shfl_sync(0xffffffffu, incl, 31);
    int base = 0;
    if (lane == 31 && total > 0) base = atomicAdd(&D.ctl[nv_index(D)], total);
    base = __shfl_sync(0xffffffffu, base, 31);
    if (want > 0) {
        const int r0 = base + incl - want;  // position inside this round's (half of the) pool
        const int room = D.pool_cap - r0;
        const int cnt = room <= 0 ? 0 : (want < room ? want : room);
        const int v0 = D.pool_base + r0;
        D.t_first[b] = v0;
        D.t_count[b] = cnt;
        for (int i = 0; i < cnt; ++i) {
            D.t_inst[v0 + i] = b;
            D.t_aidx[v0 + i] = a0 + i;
        }
    }
}

// What backward_pass leaves behind in the solver for an instance that was in PH_BACKWARD (ok = the
// recursion met no non-PD Q_uu), and how many alphas the instance evaluates this round.
template <typename T>
__device__ __forceinline__ void after_backward(const Dev<T>& D, int b, bool ran, bool ok, int ph, int* want, int* a0) {
    if (ran) {
        D.rec_valid[b] = 1;
        if (!ok) {
            end_iteration(D, D.P[D.tmpl[b]], b, ST_BWD_FAIL, -1, D.J_cur[b]);  // cpp:345-347, :118-120
            ph = D.phase[b];
        } else {
            D.status[b] = ST_RUNNING;  // set by the derivative stage (cpp:472/475)
            D.phase[b] = ph = PH_SEARCH;
            D.aidx[b] = 0;
        }
    }
    if (ph == PH_SEARCH) {
        *a0 = D.aidx[b];
        const int rem = kNumAlphas - *a0;
        int w = 1;
        if (D.wide_mode && D.wide[b]) w = D.wide_step ? (rem < *a0 + 2 ? rem : *a0 + 2) : rem;
        *want = w;
    }
}

// ---------------------------------------------------------------------------
// K5  backward_pass (cpp:383-440), one thread per trajectory.
//     solver == 0: the stand-alone operator (roofline leg, stage tests).
//     solver != 0: instances in PH_BACKWARD run the recursion and start their
//     line search; then every searching instance claims its trial-pool slots
//     for this round (warp-aggregated, one atomic per warp).
// ---------------------------------------------------------------------------
template <typename T, bool kPrefetch, bool kFused = false>
__global__ void __launch_bounds__(128) k_backward(Dev<T> D, int B, int solver, int par) {
#ifndef CILQR_PARITY
    __shared__ DevParams<T> sP[kFused ? CILQR_B200_MAX_TEMPLATES : 1];
    if (kFused) D.P = stage_params(D, sP);
#endif
    const int lane = threadIdx.x & 31;
    const int n_threads = gridDim.x * blockDim.x;
    // solver: only the instances on this round's work list
    const int* list = solver ? D.act + size_t(par) * D.Bs : nullptr;
    const int n = solver ? D.ctl[CTL_NACT + par] : B;
    // a short list is spread over all warps of the grid (lpw consecutive entries per warp) instead of
    // filling a few warps: the chain latency is the same and the scattered loads of a sparse list
    // do not queue up in a handful of SMs
    const int n_warps = n_threads >> 5;
    int lpw = (n + n_warps - 1) / n_warps;
    lpw = (lpw < 1 ? 1 : (lpw > 32 ? 32 : lpw));
    if (n > 16384) lpw = 32;  // bandwidth-bound: full warps, whole rows
    const int per_pass = n_warps * lpw;
    const int rounds = (n + per_pass - 1) / per_pass;
    const int first_idx = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * lpw + lane;
    for (int it = 0, idx = first_idx; it < rounds; ++it, idx += per_pass) {
        const bool in = lane < lpw && idx < n;
        const int b = in ? (list ? list[idx] : idx) : 0;
        int want = 0, a0 = 0;
        if (in) {
            auto recursion = [&]() {
#ifndef CILQR_PARITY
                if constexpr (kFused) return riccati_fused<T>(D, D.P[D.tmpl[b]], b, D.lamb[b]);
#endif
                return riccati<T, kPrefetch>(D, b, D.lamb[b]);
            };
            if (!solver) {
                bool ok = recursion();
                D.status[b] = ok ? ST_RUNNING : ST_BWD_FAIL;
            } else {
                D.commit_src[b] = -1;  // consumed by the derivative stage just before
                D.t_count[b] = 0;
                const int ph = D.phase[b];
                bool ok = true;
                if (ph == PH_BACKWARD) ok = recursion();
                after_backward(D, b, ph == PH_BACKWARD, ok, ph, &want, &a0);
            }
        }
        if (solver) claim_slots(D, b, want, a0, lane);
    }
}

// ---------------------------------------------------------------------------
// K5, staged variant for latency-bound batches: one warp per tile of 32 consecutive instances.
//     The 28 rows of a step's record (32 scalars each, contiguous in the step-major layout) are
//     brought into a shared-memory ring by bulk asynchronous copies (cp.async.bulk, completion on
//     an mbarrier), kStages steps ahead of the recursion, so the serial chain reads its operands
//     with immediate-offset shared-memory loads: no per-load address arithmetic and no registers
//     tied up in software prefetch.  Same arithmetic (riccati_step) and the same bits as the
//     other variants.  Walks every tile of the batch (a work list would break the tiles up);
//     tiles without an instance in PH_BACKWARD skip the recursion.
// ---------------------------------------------------------------------------
constexpr int kStagedStages = 3;

__device__ __forceinline__ unsigned smem_addr(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_load(unsigned dst_smem, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// Slot claiming for a tile that owns a private region [tile_base, tile_base + tile_slots) of the trial pool
// (k_solve_tiles): same layout of the claims as claim_slots, no global counter.  Returns the slots in use.
template <typename T>
__device__ __forceinline__ int claim_slots_tile(const Dev<T>& D, int b, int want, int a0, int lane, int tile_base, int tile_slots) {
    int incl = want;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (want > 0) {
        const int off = incl - want;
        const int room = tile_slots - off;
        const int cnt = room <= 0 ? 0 : (want < room ? want : room);
        D.t_first[b] = tile_base + off;
        D.t_count[b] = cnt;
        for (int i = 0; i < cnt; ++i) {
            D.t_inst[tile_base + off + i] = b;
            D.t_aidx[tile_base + off + i] = a0 + i;
        }
    }
    return total < tile_slots ? total : tile_slots;
}

// Shared-memory state of the staged backward pass of one warp.
template <typename T>
struct StagedRing {
    T (*stage)[kRecFields][32];     // [kStagedStages]
    unsigned long long* full;       // [kStagedStages] mbarriers
    unsigned issued, consumed;      // ring positions (steps), running over all tiles this warp handles
};

// The backward pass of tile `tile` (32 consecutive instances, lane = instance) by the calling warp; see
// k_backward_staged.  solver: 0 the stand-alone operator; 1 inside the sequential rounds (state transitions and slot
// claims follow: tile_slots < 0: slots claimed from the global trial pool (claim_slots); otherwise from the tile's
// private region and the number of slots in use is returned); 2 look-ahead rounds, the instances' own jobs (into the
// spare copy of their gains); 3 look-ahead rounds, the speculative jobs of a tile of 32 trial slots (tile counts
// slots of this round's half of the pool; records and gains are the slots' own arrays).
template <typename T>
__device__ __forceinline__ int backward_tile(const Dev<T>& D, StagedRing<T>& R, int tile, int B, int solver, int lane,
                                             int tile_base, int tile_slots) {
    const int N = D.N;
    constexpr unsigned kStepBytes = kRecFields * 32 * sizeof(T);
    // index of my trajectory in the arrays the pass works on (instances: stride Bs; trial slots: stride Vs)
    const bool slots = solver == 3;
    const size_t S = slots ? size_t(D.Vs) : size_t(D.Bs);
    const int first = slots ? D.pool_base + tile * 32 : tile * 32;
    const int end = slots ? D.pool_base + view_count(D, 1, B) : B;
    const int b = first + lane;
    const bool in = b < end;
    int ph = PH_DONE;
    bool run = false;
    if (in) {
        if (!solver) {
            run = true;
        } else if (solver == 3) {
            run = D.t_job[b] != 0;
        } else if (solver == 2) {
            run = job_of(D, b) != -2;
        } else {
            D.commit_src[b] = -1;  // consumed by the derivative stage just before
            D.t_count[b] = 0;
            ph = D.phase[b];
            run = ph == PH_BACKWARD;
        }
    }
    // where the gains go: the instance's arrays (sequential rounds, stand-alone operator), the spare copy of the
    // instance's gains (copy 1 while the current ones are copy 0 or still in a slot, else copy 0), the slot's arrays
    int gs = 0;
    if (solver == 2 && in) gs = D.gsel[b] == 1 ? 0 : 1;
    T* const Kbase = slots ? D.Kg_t : Kg_of(D, gs);
    T* const dbase = slots ? D.dg_t : dg_of(D, gs);
    T* const dVbase = slots ? D.dV_t : dV_of(D, gs);
    T* const recbase = slots ? D.rec_t : D.rec;
    bool ok = true;
    if (__any_sync(0xffffffffu, run)) {
        // field 0 of the tile's record of step 0; the record of step k is k * kRecFields * S scalars further on
        const T* tile_rec = recbase + size_t(first / kRecTile) * (kRecFields * kRecTile);
        // the tile's record of step `step` (one contiguous block) -> ring slot issued % kStagedStages
        auto issue = [&](int step) {
            const unsigned s = R.issued % kStagedStages;
            if (lane == 0) {
                const unsigned bar = smem_addr(&R.full[s]);
                mbar_arrive_expect_tx(bar, kStepBytes);
                bulk_load(smem_addr(&R.stage[s][0][0]), tile_rec + size_t(step) * kRecFields * S, kStepBytes, bar);
            }
            ++R.issued;
        };
        for (int j = 0; j < kStagedStages && j < N; ++j) issue(N - 1 - j);
        T lamb = T(0);
        if (run) lamb = solver == 3 ? D.jlamb_t[b] : (solver == 2 ? D.job_lamb[b] : D.lamb[b]);
        T Vx[4], V[kVN];
        {
            const T* rec = tile_rec + size_t(N) * kRecFields * S + (in ? lane : 0) * rl<T>();
#pragma unroll
            for (int c = 0; c < 4; ++c) Vx[c] = rec[rf<T>(kRecLx + c)];
            load_terminal_V(rec, V);
        }
        T dV0 = 0, dV1 = 0;
        T* Kp = Kbase + size_t(N) * 8 * S + (in ? b : first);
        T* dp = dbase + size_t(N) * 2 * S + (in ? b : first);
        for (int i = N - 1; i >= 0; --i) {
            const unsigned s = R.consumed % kStagedStages, parity = (R.consumed / kStagedStages) & 1u;
            ++R.consumed;
            Kp -= 8 * S;
            dp -= 2 * S;
            mbar_wait(smem_addr(&R.full[s]), parity);
            if (run) {
                T K[8] = {0, 0, 0, 0, 0, 0, 0, 0}, d0 = 0, d1 = 0;
                if (ok) {
                    T r[kRecFields];
                    load_record<T, false>(&R.stage[s][0][0] + lane * rl<T>(), r);
                    ok = riccati_step(r, lamb, Vx, V, dV0, dV1, K, d0, d1);
                    if (!ok) {
                        // rows not reached stay zero, as in the reference (cpp:392-393, :418)
                        d0 = d1 = 0;
#pragma unroll
                        for (int c = 0; c < 8; ++c) K[c] = 0;
                    }
                }
                dp[0] = d0;
                dp[S] = d1;
#pragma unroll
                for (int c = 0; c < 8; ++c) Kp[size_t(c) * S] = K[c];
            }
            __syncwarp();  // every lane is done with slot s
            if (i - kStagedStages >= 0) issue(i - kStagedStages);
        }
        if (run) {
            dVbase[b] = dV0;
            dVbase[S + b] = dV1;
        }
    }
    if (in && !solver) D.status[b] = ok ? ST_RUNNING : ST_BWD_FAIL;
    if (solver >= 2) {
        if (run) (solver == 3 ? D.jok_t : D.job_ok)[b] = ok;
        return 0;
    }
    int want = 0, a0 = 0;
    if (in && solver) after_backward(D, b, run, ok, ph, &want, &a0);
    if (!solver) return 0;
    if (tile_slots < 0) {
        claim_slots(D, in ? b : 0, want, a0, lane);
        return 0;
    }
    return claim_slots_tile(D, in ? b : 0, want, a0, lane, tile_base, tile_slots);
}

template <typename T>
__global__ void __launch_bounds__(32) k_backward_staged(Dev<T> D, int B, int solver) {
    __shared__ __align__(128) T stage[kStagedStages][kRecFields][32];
    __shared__ __align__(8) unsigned long long full[kStagedStages];
    const int lane = threadIdx.x;
    static_assert(kRecTile == 32, "one warp per record tile");
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kStagedStages; ++s) mbar_init(smem_addr(&full[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    StagedRing<T> R{stage, full, 0u, 0u};
    const int n_tiles = (B + 31) / 32;
    // look-ahead rounds: the tiles of this round's trial slots (speculative jobs) first, then the instances' own jobs
    const int n_slot_tiles = solver == 2 ? (view_count(D, 1, B) + 31) / 32 : 0;
    for (int t = blockIdx.x; t < n_slot_tiles + n_tiles; t += gridDim.x) {
        const bool st = t < n_slot_tiles;
        backward_tile(D, R, st ? t : t - n_slot_tiles, B, st ? 3 : solver, lane, 0, -1);
    }
}

// ---------------------------------------------------------------------------
// K6  forward_pass (cpp:442-461): u' = u + K (x' - x) + alpha d, x' = f(x', u'),
//     one thread per trial slot.  In the solver alpha = 2^-aidx of the slot;
//     the stage operator passes explicit alphas (slot v = instance v).
// ---------------------------------------------------------------------------
// kMatch: the waypoint match of each new position follows in the same thread (bandwidth-bound rounds:
// no second pass over the trial pool; the scan of a step is ~9 waypoints at the usual speeds).
template <typename T, bool kMatch>
__global__ void __launch_bounds__(128) k_forward(Dev<T> D, int B, int solver) {
    __shared__ DevParams<T> sP[CILQR_B200_MAX_TEMPLATES];
    D.P = stage_params(D, sP);
    const int N = D.N;
    const size_t Bs = D.Bs, Vs = D.Vs;
    const int count = solver ? view_count(D, 1, B) : B;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < count; v += gridDim.x * blockDim.x) {
        const int b = solver ? D.t_inst[v] : v;
        // solver scalars into registers: the stores below could alias *P as far as the compiler knows,
        // and a reload per step puts a memory latency on the serial chain
        const DevParams<T>* Pp = D.P + D.tmpl[b];
        const T p_dt = Pp->dt, p_wb = Pp->wheelbase;
        const int p_ref = Pp->ref_point;
        const T alpha = solver ? T(1) / T(1 << D.t_aidx[v]) : D.alpha[b];
        const T* wx = D.wx + Pp->wp_off;
        const T* wy = D.wy + Pp->wp_off;
        const int wp_len = Pp->wp_len;
        int match = 0;
        if (kMatch) {
            // x'_0 = x_0: the match of step 0 is the current trajectory's
            match = D.ridx[b];
            D.ridx_t[v] = match;
        }
        T xn[4];
        // operands of step i+1 are fetched while step i computes: the rollout is one serial
        // dependency chain, so load latency must stay off it
        T cx[4], cu[2], cd[2], cK[8];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            cx[c] = D.X[at(Bs, 0, c, 4, b)];
            xn[c] = cx[c];
            D.Xt[at(Vs, 0, c, 4, v)] = xn[c];
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            cu[r] = D.U[at(Bs, 0, r, 2, b)];
            cd[r] = D.dg[at(Bs, 0, r, 2, b)];
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) cK[c] = D.Kg[at(Bs, 0, c, 8, b)];
        for (int i = 0; i < N; ++i) {
            T nxx[4], nxu[2], nxd[2], nxK[8];
            const int ip = i + 1 < N ? i + 1 : i;
#pragma unroll
            for (int c = 0; c < 4; ++c) nxx[c] = ld_early(D.X + at(Bs, ip, c, 4, b));
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                nxu[r] = ld_early(D.U + at(Bs, ip, r, 2, b));
                nxd[r] = ld_early(D.dg + at(Bs, ip, r, 2, b));
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) nxK[c] = ld_early(D.Kg + at(Bs, ip, c, 8, b));
            T dx[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) dx[c] = xn[c] - cx[c];
            T un[2];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                T s = 0;
#pragma unroll
                for (int c = 0; c < 4; ++c) s += cK[r * 4 + c] * dx[c];
                un[r] = (cu[r] + s) + alpha * cd[r];
                D.Ut[at(Vs, i, r, 2, v)] = un[r];
            }
            T nx[4];
            propagate(xn, un[0], un[1], p_dt, p_wb, p_ref, nx);
            if (kMatch) {
                match = match_from(wx, wy, wp_len, match, nx[0], nx[1]);
                D.ridx_t[size_t(i + 1) * Vs + v] = match;
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                xn[c] = nx[c];
                D.Xt[at(Vs, i + 1, c, 4, v)] = nx[c];
                cx[c] = nxx[c];
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                cu[r] = nxu[r];
                cd[r] = nxd[r];
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) cK[c] = nxK[c];
        }
    }
}

// K6 on two lanes per trial slot (latency-bound batches).  A rollout step holds two independent
// sin/cos evaluations (~220 clk each on B200) on an otherwise short dependency chain; lane 0 of a
// pair takes the heading, lane 1 the steering side, they swap results by shuffle and both apply the
// same algebra (step_from_trig), so the state stays replicated and the bits equal k_forward's.
template <typename T>
__global__ void __launch_bounds__(128) k_forward2(Dev<T> D, int B) {
    const int N = D.N;
    const size_t Bs = D.Bs, Vs = D.Vs;
    const int count = view_count(D, 1, B);
    const int role = threadIdx.x & 1;
    const int pair = (blockIdx.x * blockDim.x + threadIdx.x) >> 1;
    const int n_pairs = (gridDim.x * blockDim.x) >> 1;
    const int warp_first = pair - ((threadIdx.x & 31) >> 1);
    for (int base = warp_first; base < count; base += n_pairs) {
        const int v = base + ((threadIdx.x & 31) >> 1);
        const bool live = v < count;
        const int vv = live ? v : count - 1;
        const int b = D.t_inst[vv];
        const DevParams<T>* Pp = D.P + D.tmpl[b];
        const T p_dt = Pp->dt, p_dtw = Pp->dt / Pp->wheelbase;
        const int p_ref = Pp->ref_point;
        const bool mixed = __any_sync(0xffffffffu, p_ref != 0);
        const T alpha = T(1) / T(1 << D.t_aidx[vv]);
        T xn[4], cx[4], cu[2], cd[2], cK[8];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            cx[c] = D.X[at(Bs, 0, c, 4, b)];
            xn[c] = cx[c];
            if (live && role == 0) D.Xt[at(Vs, 0, c, 4, v)] = xn[c];
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            cu[r] = D.U[at(Bs, 0, r, 2, b)];
            cd[r] = D.dg[at(Bs, 0, r, 2, b)];
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) cK[c] = D.Kg[at(Bs, 0, c, 8, b)];
        for (int i = 0; i < N; ++i) {
            T nxx[4], nxu[2], nxd[2], nxK[8];
            const int ip = i + 1 < N ? i + 1 : i;
#pragma unroll
            for (int c = 0; c < 4; ++c) nxx[c] = ld_early(D.X + at(Bs, ip, c, 4, b));
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                nxu[r] = ld_early(D.U + at(Bs, ip, r, 2, b));
                nxd[r] = ld_early(D.dg + at(Bs, ip, r, 2, b));
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) nxK[c] = ld_early(D.Kg + at(Bs, ip, c, 8, b));
            T dx[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) dx[c] = xn[c] - cx[c];
            T un[2];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                T s = 0;
#pragma unroll
                for (int c = 0; c < 4; ++c) s += cK[r * 4 + c] * dx[c];
                un[r] = (cu[r] + s) + alpha * cd[r];
            }
            if (live && role == 0) {
                D.Ut[at(Vs, i, 0, 2, v)] = un[0];
                D.Ut[at(Vs, i, 1, 2, v)] = un[1];
            }
            // the two sin/cos pairs of the step (yaw on lane 0, steering angle on lane 1), then swapped
            T s_head, c_head, turn;
            {
                T s, c;
                m_sincos(role ? un[1] : xn[3], &s, &c);
                const T os = __shfl_xor_sync(0xffffffffu, s, 1), oc = __shfl_xor_sync(0xffffffffu, c, 1);
                step_trig(p_ref, mixed, role ? os : s, role ? oc : c, role ? s : os, role ? c : oc, &s_head, &c_head, &turn);
            }
            T nx[4];
            step_from_trig(xn, un[0], p_dt, p_dtw, p_ref, s_head, c_head, turn, nx);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                xn[c] = nx[c];
                if (live && role == 0) D.Xt[at(Vs, i + 1, c, 4, v)] = nx[c];
                cx[c] = nxx[c];
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                cu[r] = nxu[r];
                cd[r] = nxd[r];
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) cK[c] = nxK[c];
        }
    }
}

// K6 + K1 as a two-stage pipeline inside one block (latency-bound batches).  The waypoint match of
// step k needs only the position x'_k, so it can trail the rollout instead of following it as a
// second kernel:
//   warp 0          rolls 16 trials out, two lanes each as in k_forward2 (lane r of a pair also owns
//                   control row r: its feedback row, u and d), writes every new position into a
//                   shared-memory ring covering the whole horizon and publishes its progress
//                   (fence, then a counter);
//   G/2 scan warps  match positions as they appear (G lanes per trial, k_ref_match<T,G>'s scan),
//                   each at its own pace.
// A round then pays about max(rollout, scan) instead of their sum.  Needs N + 1 <= kPipeMaxSteps
// (the host falls back to the two kernels otherwise).  A third stage for the step costs was measured
// and dropped: one cost item is ~3-6 us of serial fp64 work, so hiding it takes >= 8 lanes per trial
// at the rollout's register count, which no longer fits a benchmark round into one wave
// (profiles/r01_pipeline_findings.txt).
constexpr int kPipeTrials = 16;
constexpr int kPipeMaxSteps = 128;
__host__ __device__ constexpr int pipe_threads(int G) { return 32 + kPipeTrials * G; }

__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// One group of up to kPipeTrials trials [base, base + kPipeTrials) of the pool (count = slots in use overall): the
// calling warp is either the group's roller (two lanes per trial) or scan warp number `scan_warp` (G lanes per
// trial).  Positions travel through pos[step][trial][2]; step k is handed over by an arrival on the mbarrier
// bars[k] (release) that the scan warps wait for (acquire) in phase `parity` — every barrier completes exactly
// one phase per call, so the caller alternates parity and separates calls by a block-wide barrier.
template <typename T, int G, bool kLa>
__device__ __forceinline__ void rollout_match_group(const Dev<T>& D, T (*pos)[kPipeTrials][2], unsigned long long* bars,
                                                    unsigned parity, int base, int count, bool roller, int scan_warp, int lane) {
    const int N = D.N;
    const size_t Bs = D.Bs, Vs = D.Vs;
    // rollout lanes: slot = lane / 2, role = lane & 1; scan lanes: slot = (32 / G) scan_warp + lane / G
    const int role = lane & 1;
    const int sub = lane % G;
    const int grp_shift = lane - sub;
    const unsigned grp_mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << grp_shift);
    const int slot = roller ? (lane >> 1) : (scan_warp * (32 / G) + lane / G);
    const int v = base + slot;
    const bool live = v < count;
    const int vv = live ? v : count - 1;
    const int b = D.t_inst[vv];
    const DevParams<T>* Pp = D.P + D.tmpl[b];
    if (roller) {
        const T p_dt = Pp->dt, p_dtw = Pp->dt / Pp->wheelbase;
        const int p_ref = Pp->ref_point;
        const bool mixed = __any_sync(0xffffffffu, p_ref != 0);
        const T alpha = T(1) / T(1 << D.t_aidx[vv]);
        // the trajectory the search starts from: the instance's own arrays or, in a look-ahead solve, the trial
        // slot of the previous round that was accepted and is being copied there meanwhile; its gains: the
        // instance's current copy
        // (kLa = false, the sequential rounds: always the instance's own arrays — base pointers from the parameter
        // bank, no registers on a chain that has none to spare)
        const int cs = kLa ? D.cur_src[b] : -1;
        const T* Xc = (kLa && cs >= 0) ? D.Xt + cs : D.X + b;
        const T* Uc = (kLa && cs >= 0) ? D.Ut + cs : D.U + b;
        const size_t xs = (kLa && cs >= 0) ? Vs : Bs;
        GainsAt<T> Gc;
        if (kLa) {
            Gc = gains_at(D, b);
        } else {
            Gc.K = D.Kg + b;
            Gc.d = D.dg + b;
            Gc.stride = Bs;
        }
        const T* Kc = Gc.K;
        const T* dc = Gc.d;
        const size_t gst = Gc.stride;
        // lane `role` of a pair owns control row `role`: its feedback row, u and d
        T xn[4], cx[4], cK[4], cu, cd;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            cx[c] = Xc[size_t(c) * xs];
            xn[c] = cx[c];
        }
        if (role == 0) {
            pos[0][slot][0] = xn[0];
            pos[0][slot][1] = xn[1];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_addr(&bars[0]));
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (live && role == 0) D.Xt[at(Vs, 0, c, 4, v)] = xn[c];
        cu = Uc[size_t(role) * xs];
        cd = dc[size_t(role) * gst];
#pragma unroll
        for (int c = 0; c < 4; ++c) cK[c] = Kc[size_t(role * 4 + c) * gst];
        for (int i = 0; i < N; ++i) {
            T nxx[4], nxK[4], nxu, nxd;
            const int ip = i + 1 < N ? i + 1 : i;
#pragma unroll
            for (int c = 0; c < 4; ++c) nxx[c] = ld_early(Xc + size_t(ip * 4 + c) * xs);
            nxu = ld_early(Uc + size_t(ip * 2 + role) * xs);
            nxd = ld_early(dc + size_t(ip * 2 + role) * gst);
#pragma unroll
            for (int c = 0; c < 4; ++c) nxK[c] = ld_early(Kc + size_t(ip * 8 + role * 4 + c) * gst);
            T fb = 0;
#pragma unroll
            for (int c = 0; c < 4; ++c) fb += cK[c] * (xn[c] - cx[c]);
            const T mine = (cu + fb) + alpha * cd;  // u'[role]
            const T other = __shfl_xor_sync(0xffffffffu, mine, 1);
            const T acc = role ? other : mine;
            T s_head, c_head, turn;
            {
                T sn, cs;
                m_sincos(role ? mine : xn[3], &sn, &cs);
                const T os = __shfl_xor_sync(0xffffffffu, sn, 1), oc = __shfl_xor_sync(0xffffffffu, cs, 1);
                step_trig(p_ref, mixed, role ? os : sn, role ? oc : cs, role ? sn : os, role ? cs : oc, &s_head,
                          &c_head, &turn);
            }
            T nx[4];
            step_from_trig(xn, acc, p_dt, p_dtw, p_ref, s_head, c_head, turn, nx);
            // hand the position over first, then this step's global stores
            if (role == 0) {
                pos[i + 1][slot][0] = nx[0];
                pos[i + 1][slot][1] = nx[1];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_addr(&bars[i + 1]));
            if (live) D.Ut[at(Vs, i, role, 2, v)] = mine;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                xn[c] = nx[c];
                if (live && role == 0) D.Xt[at(Vs, i + 1, c, 4, v)] = nx[c];
                cx[c] = nxx[c];
            }
            cu = nxu;
            cd = nxd;
#pragma unroll
            for (int c = 0; c < 4; ++c) cK[c] = nxK[c];
        }
    } else {
        const int M = Pp->wp_len;
        const T* wx = D.wx + Pp->wp_off;
        const T* wy = D.wy + Pp->wp_off;
        int start = 0;
        // the waypoints of the window a scan will look at next are fetched while it works on the current one: the
        // scan of a step is a chain of probes, and with a narrow window every probe would otherwise wait for a new
        // sector of the table (an L2 round trip per probe)
        auto wp_load = [&](int j0, T* ox, T* oy) {
            const int jc = j0 + sub < M ? j0 + sub : M - 1;
            *ox = __ldg(wx + jc);
            *oy = __ldg(wy + jc);
        };
        T cwx, cwy;
        wp_load(0, &cwx, &cwy);
        for (int k = 0; k <= N; ++k) {
            mbar_wait(smem_addr(&bars[k]), parity);
            const T px = pos[k][slot][0], py = pos[k][slot][1];
            int found = -1;
            bool done = !live;
            while (!__all_sync(0xffffffffu, done)) {
                T nwx, nwy;
                wp_load(start + (G - 1), &nwx, &nwy);
                int j = start + sub;
                T dj = wp_dist2(px, py, cwx, cwy);
                T dn = __shfl_down_sync(0xffffffffu, dj, 1, G);
                bool stop = !done && (sub < G - 1) && (j + 1 >= M || !(dn < dj));
                unsigned m = __ballot_sync(0xffffffffu, stop) & grp_mask;
                if (!done) {
                    if (m) {
                        found = start + (__ffs(m) - 1 - grp_shift);
                        done = true;
                    } else {
                        start += G - 1;
                        cwx = nwx;
                        cwy = nwy;
                    }
                }
            }
            if (live && found != start) wp_load(found, &cwx, &cwy);  // the next step starts at the match
            if (live) {
                if (sub == 0) D.ridx_t[size_t(k) * Vs + v] = found;
                start = found;
            }
        }
    }
}

// G = scan lanes per trial: 16 (two blocks per SM) when the whole trial pool fits one wave that way,
// 8 (four blocks per SM, a slower but still hidden scan) beyond that.
template <typename T, int G, bool kLa = false>
__global__ void __launch_bounds__(pipe_threads(G), kLa ? 1 : (G == 16 ? 2 : 4)) k_rollout_match(Dev<T> D, int B) {
    __shared__ T pos[kPipeMaxSteps][kPipeTrials][2];
    __shared__ __align__(8) unsigned long long bars[kPipeMaxSteps];  // bars[k]: positions of step k are in the ring
    const int count = view_count(D, 1, B);
    // (the host sizes the grid from its bound of 20 slots per instance; most rounds use about one: blocks without a
    // trial leave before setting the barriers up — they would otherwise hold an SM's block slot for a few microseconds
    // each, and the look-ahead variant fits one block per SM)
    if (int(blockIdx.x) * kPipeTrials >= count) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int k = threadIdx.x; k <= D.N; k += blockDim.x) mbar_init(smem_addr(&bars[k]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    unsigned parity = 0;
    // (slots are absolute: a look-ahead solve alternates between the two halves of the pool)
    const int v_end = D.pool_base + count;
    for (int base = D.pool_base + blockIdx.x * kPipeTrials; base < v_end; base += gridDim.x * kPipeTrials, parity ^= 1u) {
        rollout_match_group<T, G, kLa>(D, pos, bars, parity, base, v_end, warp == 0, warp - 1, lane);
        __syncthreads();  // the ring is reused by the next group of trials
    }
}

// The line-search verdict of one instance in PH_SEARCH over the cnt > 0 trial slots it evaluated this round, in alpha
// order (iter_step cpp:356-380), followed by solve()'s bookkeeping (end_iteration).
template <typename T>
__device__ __forceinline__ void decide_instance(const Dev<T>& D, int b, int cnt) {
    const size_t Bs = D.Bs;
    const DevParams<T>& P = D.P[D.tmpl[b]];
    const int v0 = D.t_first[b];
    const int a0 = D.aidx[b];
    const T J_cur = D.J_cur[b];
    const GainsAt<T> Gc = gains_at(D, b);
    const T dV0 = Gc.dV[0], dV1 = Gc.dV[Gc.stride];
    bool ended = false;
    T last_J = J_cur;  // cost of the last trial looked at: what iter_step returns when every alpha was rejected (cpp:380)
    // trial costs four at a time (independent loads), then their verdicts in alpha order; the
    // usual case is a single trial
    for (int i0 = 0; i0 < cnt && !ended; i0 += 4) {
    T Jt[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) Jt[q] = i0 + q < cnt ? D.J_t[v0 + i0 + q] : T(0);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int i = i0 + q;
        if (i >= cnt || ended) break;
        const int v = v0 + i, a = a0 + i;
        const T new_J = Jt[q];
        last_J = new_J;
        const T alpha = T(1) / T(1 << a);
        const T actual = J_cur - new_J;
        if (a == 0 && m_fabs(actual) < P.conv_thr) {
            D.wide[b] = 0;
            end_iteration(D, P, b, ST_CONVERGED, a, new_J);  // trial discarded (cpp:358-361)
            ended = true;
        } else {
            const T approx = -(alpha * alpha * dV0 + alpha * dV1);
            if (actual > T(0) && (approx < T(0) || actual / approx > P.accept_thr)) {
                D.commit_src[b] = v;  // x, u <- new (cpp:113-116), copied by the next derivative stage
                D.J_cur[b] = new_J;
                D.rec_valid[b] = 0;
                D.wide[b] = a > 0;
                end_iteration(D, P, b, a == 0 ? ST_RUNNING : ST_SMALL_STEP, a, new_J);
                ended = true;
            }
        }
    }
    }
    if (!ended) {
        if (a0 + cnt >= kNumAlphas) {
            if (P.solve_type == 1 && D.mu) {  // cpp:377-378
                for (int i = 0; i < D.N * D.alm_cols; ++i)
                    D.mu[size_t(i) * Bs + b] = D.mu_next[size_t(i) * Bs + b];
                D.rho[b] = std_min((1 + P.alm_gamma) * D.rho[b], P.max_rho);
                D.rec_valid[b] = 0;
            }
            D.wide[b] = 1;
            end_iteration(D, P, b, ST_FWD_FAIL, -1, last_J);
        } else {
            D.aidx[b] = a0 + cnt;
            D.wide[b] = 1;  // alpha = 1 was rejected: evaluate the remaining alphas together
        }
    }
}

// ---------------------------------------------------------------------------
// K7  the line-search verdict of iter_step (cpp:356-380) over the slots each
//     searching instance evaluated this round, in alpha order, then solve()'s
//     bookkeeping.  The same pass filters the round's work list into the next
//     round's (instances still running, plus finished ones whose last accepted
//     step the derivative stage has yet to commit), keeping instance order: a
//     block handles chunks of 128 list entries, scans its survivors, and chains
//     its count to the preceding chunks' through scan_state (decoupled
//     look-back; one chunk per CTA, handed out by ticket, so a chunk only ever
//     waits on chunks whose CTAs are already running).  The last block to finish
//     publishes progress to the host and re-arms the round counters.
// ---------------------------------------------------------------------------
constexpr unsigned long long kScanAgg = 1ull << 32, kScanIncl = 2ull << 32;
__device__ __forceinline__ unsigned long long scan_word(unsigned epoch, unsigned long long flag, int value) {
    return (static_cast<unsigned long long>(epoch) << 34) | flag | unsigned(value);
}

template <typename T>
__global__ void __launch_bounds__(128) k_decide(Dev<T> D, int B, int par, unsigned epoch) {
    const size_t Bs = D.Bs;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ int s_chunk, s_excl, s_wsum[4], s_active, s_trials;
    const int* list = D.act + size_t(par) * Bs;
    int* next_list = D.act + size_t(par ^ 1) * Bs;
    volatile unsigned long long* state = D.scan_state;
    const int n = D.ctl[CTL_NACT + par];
    const int n_chunks = (n + int(blockDim.x) - 1) / int(blockDim.x);
    int my_active = 0, my_trials = 0;
    // chunks by ticket only when the grid may exceed what is resident at once (a chunk must never wait
    // on one whose block has not started); small grids save the atomic's round trip
    const bool by_ticket = gridDim.x > 2 * 148;
    if (threadIdx.x == 0) s_chunk = by_ticket ? atomicAdd(&D.ctl[CTL_CHUNK], 1) : int(blockIdx.x);
    __syncthreads();
    // One chunk per CTA: the host launches ceil(n_bound / 128) >= n_chunks CTAs.  (A CTA must not take a
    // second chunk: with more CTAs than fit on the GPU it would wait on chunks whose CTAs can only
    // start once it has exited.)  Written as a loop so that the tail of the kernel is reached uniformly.
    for (int chunk = s_chunk; chunk < n_chunks; chunk = n_chunks) {
        const int idx = chunk * int(blockDim.x) + int(threadIdx.x);
        const bool in = idx < n;
        const int b = in ? list[idx] : 0;
        int ph = in ? D.phase[b] : PH_DONE;
        const int cnt = in ? D.t_count[b] : 0;
        if (ph == PH_SEARCH && cnt > 0) {
            my_trials += cnt;
            decide_instance(D, b, cnt);
            ph = D.phase[b];
        }
        my_active += ph != PH_DONE;
        // survivors of this chunk, in list order
        const bool keep = in && (ph != PH_DONE || D.commit_src[b] >= 0);
        const unsigned km = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_wsum[warp] = __popc(km);
        __syncthreads();
        int rank = __popc(km & ((1u << lane) - 1u));
        for (int w = 0; w < warp; ++w) rank += s_wsum[w];
        if (warp == 0) {
            const int total = s_wsum[0] + s_wsum[1] + s_wsum[2] + s_wsum[3];
            int excl = 0;
            if (chunk > 0) {
                if (lane == 0) state[chunk] = scan_word(epoch, kScanAgg, total);
                // look back over the preceding chunks, 32 at a time, nearest first
                for (int look = chunk - 1;; look -= 32) {
                    const int j = look - lane;
                    unsigned long long v = scan_word(epoch, kScanIncl, 0);  // in front of chunk 0
                    if (j >= 0) {
                        do {
                            v = state[j];
                        } while (unsigned(v >> 34) != epoch);
                    }
                    const unsigned incl = __ballot_sync(0xffffffffu, (v & kScanIncl) != 0);
                    const int upto = incl ? __ffs(incl) - 1 : 31;  // nearest chunk holding an inclusive prefix
                    int val = lane <= upto ? int(unsigned(v)) : 0;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
                    excl += val;
                    if (incl) break;
                }
            }
            if (lane == 0) {
                state[chunk] = scan_word(epoch, kScanIncl, excl + total);
                s_excl = excl;
                if (chunk == n_chunks - 1) D.ctl[CTL_NACT + (par ^ 1)] = excl + total;
            }
        }
        __syncthreads();
        if (keep) next_list[s_excl + rank] = b;
        __syncthreads();  // s_wsum / s_excl are reused by the next chunk
    }
    // block-level count, then one atomic per block
    if (threadIdx.x == 0) s_active = s_trials = 0;
    __syncthreads();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        my_active += __shfl_down_sync(0xffffffffu, my_active, o);
        my_trials += __shfl_down_sync(0xffffffffu, my_trials, o);
    }
    if (lane == 0) {
        atomicAdd(&s_active, my_active);
        atomicAdd(&s_trials, my_trials);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // one atomic per block: running instances (24 bits), trials (24 bits) and finished blocks (16 bits) in
        // one 64-bit counter; the block that completes the count has the round's totals in hand
        unsigned long long* tally = reinterpret_cast<unsigned long long*>(&D.ctl[CTL_TALLY]);
        const unsigned long long mine = (1ull << 48) | (static_cast<unsigned long long>(unsigned(s_trials)) << 24) | unsigned(s_active);
        __threadfence();
        const unsigned long long sum = atomicAdd(tally, mine) + mine;
        if (int(sum >> 48) == int(gridDim.x)) {
            __threadfence();
            const int active = int(sum & 0xffffffu);
            const int trials = int((sum >> 24) & 0xffffffu);
            const int n_next = *reinterpret_cast<volatile int*>(&D.ctl[CTL_NACT + (par ^ 1)]);
            const int round = D.ctl[CTL_ROUND] + 1;
            D.ctl[CTL_ROUND] = round;
            *reinterpret_cast<unsigned long long*>(&D.ctl[CTL_TRIALS]) += static_cast<unsigned long long>(trials);
            *tally = 0ull;
            // (look-ahead solves: this round's count is still being read on the other stream; re-arm the next round's)
            D.ctl[D.spec ? CTL_NV2 + ((D.round_id + 1) & 1) : int(CTL_NV)] = 0;
            D.ctl[CTL_CHUNK] = 0;
            // consumed; the verdict kernel of the next round refills it (look-ahead solves: the derivative kernel of
            // this round, on the other stream, may still be reading it — k_adopt clears it)
            if (!D.spec) D.ctl[CTL_NACT + par] = 0;
            // two independent 64-bit words (rounds completed << 32 | instances still running, and
            // rounds completed << 32 | length of the next work list): the host reads each with one
            // load and either may be stale (both counts only ever shrink), so no system-scope fence
            // has to order them
            volatile unsigned long long* hw = reinterpret_cast<volatile unsigned long long*>(D.h_ctl);
            hw[1] = (static_cast<unsigned long long>(unsigned(round)) << 32) | unsigned(n_next);
            hw[0] = (static_cast<unsigned long long>(unsigned(round)) << 32) | unsigned(active);
        }
    }
}

// ---------------------------------------------------------------------------
// Look-ahead rounds (latency-bound batches).  An iteration of the reference is a chain
//     derivatives -> backward pass -> rollout -> cost -> verdict -> derivatives of the accepted trajectory -> ...
// and for small batches every link is a latency chain of its own (~140 us per iteration on B200).  But the
// derivatives and the backward pass of the NEXT iteration depend on the verdict only through WHICH trajectory was
// accepted and the regularisation that follows from it, and 96 % of the iterations of the slowest instances accept
// the full step (alpha = 1, lambda *= decay).  So a round runs, next to the cost and verdict kernels of its trials
// (stream A), backward "jobs" on a second stream (B): derivatives (k_derivs, masked = 2 / 3) and the recursion
// (k_backward_staged, solver = 2 / 3) of
//     * the trials of the running line searches, one job per trial slot, each with the lambda that follows if it is
//       the accepted one (lambda * decay for alpha = 1, lambda otherwise) — records and gains in the slot's own arrays
//       (rec_t, Kg_t, dg_t, dV_t); the alpha = 1 trial always, the others while the work list is short;
//     * the instance's current trajectory — with lambda amplified, for a line search that is evaluating its last
//       alphas (the all-rejected outcome, cpp:375-380), or with the current lambda for an instance that has nothing
//       to roll out this round (first iteration, a verdict no job covered, a failed backward pass) — into the spare
//       copy of the instance's gains.
// k_adopt joins the two streams: where the verdict asks for exactly a backward pass that ran (same trajectory, same
// lambda — bitwise), its gains become the instance's current ones (gsel) and the instance goes straight on to its
// next line search; otherwise it gets a job for what it needs and sits the next round out.  Every instance performs
// the reference's sequence of operations on the same values — the result bits do not depend on the mode — but an
// iteration costs rollout + derivatives + recursion instead of the whole chain.  An accepted trial (with its record
// and gains, if its job was adopted) is copied into the instance's arrays by the next round's derivative kernel
// while the rollouts and the verdict read it from the slot (cur_src, gsel = 2), which is why a look-ahead solve
// alternates between two halves of the trial pool.
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) k_adopt(Dev<T> D, int par_next, int fresh) {
    const size_t Bs = D.Bs;
    const int lane = threadIdx.x & 31;
    const int* list = D.act + size_t(par_next) * Bs;
    const int n = D.ctl[CTL_NACT + par_next];
    const int n_pad = (n + 31) & ~31;  // whole warps: claim_slots is a warp collective
    // every trial of a line search gets a job (and its all-rejected outcome one on the instance) once the work list is
    // short; a long list speculates on the full step only (a job costs a derivative pass and a recursion)
    const bool spec_all = n <= D.spec_all_below;
    if (blockIdx.x == 0 && threadIdx.x == 0) D.ctl[CTL_NACT + (par_next ^ 1)] = 0;  // the list this round consumed
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_pad; idx += gridDim.x * blockDim.x) {
        const bool in = idx < n;
        const int b = in ? list[idx] : 0;
        int want = 0, a0 = 0;
        if (in) {
            const int cs = D.commit_src[b];  // accepted this round (a slot of this round's half), or -1
            int g = D.gsel[b];
            if (g == 2) g = 0;  // the slot adopted a round ago has been copied into copy 0 meanwhile (k_derivs)
            int ph = D.phase[b];
            // (fresh: the first look-ahead round of a solve follows — nothing ran a job, and the job flags of the
            // trial slots are whatever an earlier solve left)
            const int job = (!fresh && D.job_round[b] == D.round_id - 1) ? D.job_src[b] : -2;  // the round that just ended
            D.job_round[b] = D.round_id;
            // The backward pass the verdict asks for is the one over trajectory cs (-1 = the unchanged current one)
            // with lamb[b]: adopted if a job ran exactly that — the accepted slot's, or the instance's own.
            bool hit = false, ok = false;
            if (ph == PH_BACKWARD) {
                const T lamb = D.lamb[b];
                if (job != -2 && job == cs && D.job_lamb[b] == lamb) {
                    // the instance's own job: over the trial that was accepted (cs >= 0), or over the unchanged
                    // current trajectory (cs == -1)
                    hit = true;
                    ok = D.job_ok[b] != 0;
                    g ^= 1;  // the spare copy becomes the current one
                } else if (!fresh && cs >= 0 && D.t_job[cs] && D.jlamb_t[cs] == lamb) {
                    hit = true;
                    ok = D.jok_t[cs] != 0;
                    g = 2;  // gains (and record) stay in the slot's arrays for the coming round
                }
            }
            // (a job of the instance over a trial wrote that trial's record over the current trajectory's)
            if (job >= 0 && !(hit && g != 2)) D.rec_valid[b] = 0;
            D.gsel[b] = g;
            D.cur_src[b] = cs;
            D.commit_src[b] = -1;
            D.t_count[b] = 0;
            after_backward(D, b, hit, ok, ph, &want, &a0);
            ph = D.phase[b];
            // an instance that still needs a backward pass runs it next round (and sits the rollouts out)
            D.job_src[b] = ph == PH_BACKWARD ? -1 : -2;
            if (ph == PH_BACKWARD) D.job_lamb[b] = D.lamb[b];
        }
        claim_slots(D, b, want, a0, lane);  // D.pool_base / pool_cap: the half of the pool the next round uses
        if (in && want > 0) {
            const int cnt = D.t_count[b], v0 = D.t_first[b];
            const DevParams<T>& P = D.P[D.tmpl[b]];
            const T lamb = D.lamb[b];
            // the backward pass that follows if trial i is the accepted one: lambda *= decay after a full step
            // (ST_RUNNING), unchanged after a shortened one (ST_SMALL_STEP) — end_iteration
            // The instance's own job takes the likeliest outcome and costs no copies when adopted (record and gains
            // are written where the instance keeps them): the full step, or — once the round evaluates the last
            // alpha — every step rejected (ST_FWD_FAIL: the current record, recomputed if a job overwrote it, with
            // lambda amplified).  The other trials get jobs of their own slot while the work list is short.
            const bool last = spec_all && cnt > 0 && a0 + cnt >= kNumAlphas;
            for (int i = 0; i < cnt; ++i) {
                const int a = a0 + i;
                D.t_job[v0 + i] = ((a == 0 && last) || (a > 0 && spec_all)) ? 1 : 0;
                D.jlamb_t[v0 + i] = a == 0 ? lamb * P.lamb_decay : lamb;
            }
            if (last) {
                D.job_src[b] = -1;
                D.job_lamb[b] = std_max(P.lamb_amplify, lamb * P.lamb_amplify);
            } else if (a0 == 0 && cnt > 0) {
                D.job_src[b] = v0;
                D.job_lamb[b] = lamb * P.lamb_decay;
            }
        }
    }
}

// End of a look-ahead solve: the current copy of the gains back into the first one (the only one the rest of the
// library knows about); the selectors are cleared by the caller afterwards.  grid.y = rows of K + d + dV.
template <typename T>
__global__ void __launch_bounds__(128) k_gains_home(Dev<T> D, int B) {
    const size_t Bs = D.Bs;
    const int N = D.N;
    const int row = blockIdx.y;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
        if (D.gsel[b] != 1) continue;  // (2: the final commit has just copied the slot's gains into copy 0)
        if (row < N * 8) D.Kg[size_t(row) * Bs + b] = Kg_of(D, 1)[size_t(row) * Bs + b];
        else if (row < N * 10) D.dg[size_t(row - N * 8) * Bs + b] = dg_of(D, 1)[size_t(row - N * 8) * Bs + b];
        else D.dV[size_t(row - N * 10) * Bs + b] = dV_of(D, 1)[size_t(row - N * 10) * Bs + b];
    }
}

// Receding-horizon step (src/motion_planning.cpp:197): ego_state = new_x.row(1); also logs the
// tick's outcome.  hist_x [ticks+1][4][Bs], hist_iters / hist_status [ticks][Bs].
template <typename T>
__global__ void k_advance(Dev<T> D, int B, int tick, T* hist_x, int* hist_iters, int* hist_status) {
    const size_t Bs = D.Bs;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (tick == 0) hist_x[size_t(c) * Bs + b] = D.x0[size_t(c) * Bs + b];
            T v = D.X[at(Bs, 1, c, 4, b)];
            D.x0[size_t(c) * Bs + b] = v;
            hist_x[(size_t(tick + 1) * 4 + c) * Bs + b] = v;
        }
        hist_iters[size_t(tick) * Bs + b] = D.iters[b];
        hist_status[size_t(tick) * Bs + b] = D.status[b];
    }
}

// last_solve_u = u (cpp:144)
template <typename T>
__global__ void __launch_bounds__(128) k_store_last_u(Dev<T> D, int B) {
    const int i = blockIdx.y;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
        D.last_u[at(D.Bs, i, 0, 2, b)] = D.U[at(D.Bs, i, 0, 2, b)];
        D.last_u[at(D.Bs, i, 1, 2, b)] = D.U[at(D.Bs, i, 1, 2, b)];
    }
}

// Obstacle samples arrive as (x, y, yaw); replace the yaw by (sin yaw, cos yaw) in place
// (fields 2 and 3), once per upload.  rows = max_obs * (N+1).
template <typename T>
__global__ void k_obs_sincos(T* obs, int B, size_t Bs) {
    const size_t row = blockIdx.y;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
        T* p = obs + row * 4 * Bs + b;
        T s, c;
        m_sincos(p[2 * Bs], &s, &c);
        p[2 * Bs] = s;
        p[3 * Bs] = c;
    }
}

// sin / cos of the waypoint yaws of all templates, once per cilqr_b200_set_template.
template <typename T>
__global__ void k_wp_sincos(const T* wyaw, T* wsin, T* wcos, int M) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
        T s, c;
        m_sincos(wyaw[i], &s, &c);
        wsin[i] = s;
        wcos[i] = c;
    }
}

__global__ void k_pack_int(const int* __restrict__ src, int* __restrict__ dst, int B, int fill) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) dst[b] = src ? src[b] : fill;
}

// Dense reference layout <-> compact record, for the stage operators.
// dense (host layout, per trajectory): lx [N+1][4], lu [N][2], lxx [N+1][16], luu [N][4], A [N][16], B [N][8].
template <typename T>
__global__ void k_records_from_dense(Dev<T> D, int B, const double* lx, const double* lu, const double* lxx,
                                     const double* luu, const double* A, const double* Bm) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    int k = blockIdx.y;
    if (b >= B) return;
    const int N = D.N;
    const size_t Bs = D.Bs;
    T* rec = rec_at(D, k, b);
    const double* px = lx + (size_t(b) * (N + 1) + k) * 4;
    const double* pxx = lxx + (size_t(b) * (N + 1) + k) * 16;
    for (int c = 0; c < 4; ++c) rec[rf<T>(kRecLx + c)] = T(px[c]);
#ifdef CILQR_PARITY
    for (int e = 0; e < 16; ++e) rec[rf<T>(kRecLxx + e)] = T(pxx[e]);
#else
    int e = 0;
    for (int r = 0; r < 4; ++r)
        for (int c = r; c < 4; ++c, ++e) rec[rf<T>(kRecLxx + e)] = T(pxx[r * 4 + c]);
#endif
    if (k < N) {
        const double* pu = lu + (size_t(b) * N + k) * 2;
        const double* puu = luu + (size_t(b) * N + k) * 4;
        const double* pa = A + (size_t(b) * N + k) * 16;
        const double* pb = Bm + (size_t(b) * N + k) * 8;
        rec[rf<T>(kRecLu + 0)] = T(pu[0]);
        rec[rf<T>(kRecLu + 1)] = T(pu[1]);
        rec[rf<T>(kRecLuu + 0)] = T(puu[0]);
        rec[rf<T>(kRecLuu + 1)] = T(puu[1]);
        rec[rf<T>(kRecLuu + 2)] = T(puu[3]);
        rec[rf<T>(kRecA + 0)] = T(pa[0 * 4 + 2]);
        rec[rf<T>(kRecA + 1)] = T(pa[0 * 4 + 3]);
        rec[rf<T>(kRecA + 2)] = T(pa[1 * 4 + 2]);
        rec[rf<T>(kRecA + 3)] = T(pa[1 * 4 + 3]);
        rec[rf<T>(kRecA + 4)] = T(pa[3 * 4 + 2]);
        rec[rf<T>(kRecB + 0)] = T(pb[0 * 2 + 1]);
        rec[rf<T>(kRecB + 1)] = T(pb[1 * 2 + 1]);
        rec[rf<T>(kRecB + 2)] = T(pb[2 * 2 + 0]);
        rec[rf<T>(kRecB + 3)] = T(pb[3 * 2 + 1]);
    }
}

template <typename T>
__global__ void k_records_to_dense(Dev<T> D, int B, double* lx, double* lu, double* lxx, double* luu, double* A,
                                   double* Bm) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    int k = blockIdx.y;
    if (b >= B) return;
    const int N = D.N;
    const size_t Bs = D.Bs;
    const T* rec = rec_at(D, k, b);
    double* px = lx + (size_t(b) * (N + 1) + k) * 4;
    double* pxx = lxx + (size_t(b) * (N + 1) + k) * 16;
    for (int c = 0; c < 4; ++c) px[c] = double(rec[rf<T>(kRecLx + c)]);
#ifdef CILQR_PARITY
    for (int e = 0; e < 16; ++e) pxx[e] = double(rec[rf<T>(kRecLxx + e)]);
#else
    int e = 0;
    for (int r = 0; r < 4; ++r)
        for (int c = r; c < 4; ++c, ++e) {
            double v = double(rec[rf<T>(kRecLxx + e)]);
            pxx[r * 4 + c] = v;
            pxx[c * 4 + r] = v;
        }
#endif
    if (k < N) {
        double* pu = lu + (size_t(b) * N + k) * 2;
        double* puu = luu + (size_t(b) * N + k) * 4;
        double* pa = A + (size_t(b) * N + k) * 16;
        double* pb = Bm + (size_t(b) * N + k) * 8;
        pu[0] = double(rec[rf<T>(kRecLu + 0)]);
        pu[1] = double(rec[rf<T>(kRecLu + 1)]);
        puu[0] = double(rec[rf<T>(kRecLuu + 0)]);
        puu[1] = puu[2] = double(rec[rf<T>(kRecLuu + 1)]);
        puu[3] = double(rec[rf<T>(kRecLuu + 2)]);
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) pa[r * 4 + c] = (r == c) ? 1.0 : 0.0;
        pa[0 * 4 + 2] = double(rec[rf<T>(kRecA + 0)]);
        pa[0 * 4 + 3] = double(rec[rf<T>(kRecA + 1)]);
        pa[1 * 4 + 2] = double(rec[rf<T>(kRecA + 2)]);
        pa[1 * 4 + 3] = double(rec[rf<T>(kRecA + 3)]);
        pa[3 * 4 + 2] = double(rec[rf<T>(kRecA + 4)]);
        for (int c = 0; c < 8; ++c) pb[c] = 0.0;
        pb[0 * 2 + 1] = double(rec[rf<T>(kRecB + 0)]);
        pb[1 * 2 + 1] = double(rec[rf<T>(kRecB + 1)]);
        pb[2 * 2 + 0] = double(rec[rf<T>(kRecB + 2)]);
        pb[3 * 2 + 1] = double(rec[rf<T>(kRecB + 3)]);
    }
}

// Replicate the records of trajectories [0, B0) over [B0, B) (roofline leg).
template <typename T>
__global__ void k_tile_records(Dev<T> D, int B0, int B) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    int row = blockIdx.y;  // (step, field) flattened
    if (b >= B || b < B0) return;
    const int k = row / kRecFields, c = row % kRecFields;
    rec_at(D, k, b)[rf<T>(c)] = rec_at(D, k, b % B0)[rf<T>(c)];
    // the trajectory itself (the fused flavour of the backward pass reads v, yaw, u) and the template id
    if (c < 4) D.X[at(D.Bs, k, c, 4, b)] = D.X[at(D.Bs, k, c, 4, b % B0)];
    if (c < 2 && k < D.N) D.U[at(D.Bs, k, c, 2, b)] = D.U[at(D.Bs, k, c, 2, b % B0)];
    if (row == 0) D.tmpl[b] = D.tmpl[b % B0];
}

// ---------------------------------------------------------------------------
// Repack (bandwidth-bound batches).  Every array is [row][batch]: once most instances have finished,
// each 8-byte access of a survivor still costs a 32-byte sector (and a scattered 8-byte store a sector
// read-modify-write), so a round over n scattered survivors costs several times a round over a dense
// batch of n.  When the work list has shrunk to half of the slots in use (measured: halving beats
// waiting for a quarter or an eighth, profiles/r01b_findings.txt), the survivors are moved into the
// prefix [0, n): the survivors beyond the prefix trade places with the finished
// instances inside it (disjoint slot pairs, every per-instance array swapped in place), the work list
// becomes 0 .. n-1, and the solve carries on as a dense batch of n.  The same swaps applied again at
// the end of the solve put every instance back into its own slot.
// ---------------------------------------------------------------------------
constexpr int kPlanThreads = 1024;

// One block: the list of round `par` (sorted, n entries) -> swap pairs of `level`, then list := 0 .. n-1.
template <typename T>
__global__ void __launch_bounds__(kPlanThreads) k_plan_repack(Dev<T> D, int par, int level, int* src, int* dst) {
    __shared__ int s_part[kPlanThreads / 32];
    __shared__ int s_inside;
    const size_t Bs = D.Bs;
    int* list = D.act + size_t(par) * Bs;
    int* mark = D.act + size_t(par ^ 1) * Bs;  // the other list is free between rounds
    const int n = D.ctl[CTL_NACT + par];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int s = tid; s < n; s += kPlanThreads) mark[s] = 0;
    __syncthreads();
    int inside = 0;
    for (int i = tid; i < n; i += kPlanThreads) {
        const int e = list[i];
        if (e < n) {
            mark[e] = 1;
            ++inside;
        }
    }
    // entries inside the prefix (the list is sorted: they are its head)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) inside += __shfl_xor_sync(0xffffffffu, inside, o);
    if (lane == 0) s_part[warp] = inside;
    __syncthreads();
    if (tid == 0) {
        int t = 0;
        for (int w = 0; w < kPlanThreads / 32; ++w) t += s_part[w];
        s_inside = t;
    }
    __syncthreads();
    const int n_inside = s_inside;
    // holes of the prefix, in slot order: contiguous chunk per thread, block scan of the chunk counts
    const int chunk = (n + kPlanThreads - 1) / kPlanThreads;
    const int lo = min(tid * chunk, n), hi = min(lo + chunk, n);
    int holes = 0;
    for (int s = lo; s < hi; ++s) holes += mark[s] == 0;
    int incl = holes;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();  // s_part is reused
    if (lane == 31) s_part[warp] = incl;
    __syncthreads();
    int j = incl - holes;
    for (int w = 0; w < warp; ++w) j += s_part[w];
    for (int s = lo; s < hi; ++s)
        if (mark[s] == 0) {
            dst[j] = s;
            src[j] = list[n_inside + j];  // the j-th survivor beyond the prefix
            ++j;
        }
    __syncthreads();  // every read of the old list is done
    for (int i = tid; i < n; i += kPlanThreads) list[i] = i;
    if (tid == 0) D.ctl[CTL_NSWAP + level] = n - n_inside;
}

// base[row][src[j]] <-> base[row][dst[j]] for every pair j and every row of a [rows][stride] array.
template <typename E>
__global__ void __launch_bounds__(128) k_swap_rows(E* base, size_t stride, int rows, const int* __restrict__ src,
                                                   const int* __restrict__ dst, const int* __restrict__ m_ptr) {
    const int m = *m_ptr;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x) {
        const int a = src[j], b = dst[j];
        for (int row = blockIdx.y; row < rows; row += gridDim.y) {
            E* p = base + size_t(row) * stride;
            const E va = p[a], vb = p[b];
            p[a] = vb;
            p[b] = va;
        }
    }
}

// the same for the tiled derivative records
template <typename T>
__global__ void __launch_bounds__(128) k_swap_records(Dev<T> D, const int* __restrict__ src, const int* __restrict__ dst,
                                                      const int* __restrict__ m_ptr) {
    const int m = *m_ptr;
    const int rows = (D.N + 1) * kRecFields;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x) {
        const int a = src[j], b = dst[j];
        for (int row = blockIdx.y; row < rows; row += gridDim.y) {
            const int k = row / kRecFields, c = row % kRecFields;
            T* pa = rec_at(D, k, a) + rf<T>(c);
            T* pb = rec_at(D, k, b) + rf<T>(c);
            const T va = *pa, vb = *pb;
            *pa = vb;
            *pb = va;
        }
    }
}

// ---------------------------------------------------------------------------
// Synthetic workloads generated in place: the device twin of generate_host() in
// toy-example-of-ilqr_b200/scenario.py.  Every floating-point step is an explicit round-to-nearest add / multiply
// (__dadd_rn, __dmul_rn: never contracted into an FMA), floor, min, max or fmod, in the order the numpy code
// performs them, so the arrays agree with the host's bit for bit.
// ---------------------------------------------------------------------------
struct SynthLanes {
    const double* x;
    const double* y;
    const double* yaw;
    const double* lon;
    const double* nx;
    const double* ny;
    const int* off;  // [n_lanes + 1]
};

__device__ __forceinline__ unsigned long long sy_mix64(unsigned long long z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// counter-based uniform [0, 1): splitmix64 finaliser (twice) of (seed, instance id, draw index)
__device__ __forceinline__ double sy_u01(unsigned long long seed, unsigned long long inst, int draw) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (inst * 64ull + (unsigned long long)draw + 1ull);
    z = sy_mix64(sy_mix64(z));
    return double(z >> 11) * (1.0 / 9007199254740992.0);
}
__device__ __forceinline__ double sy_uniform(unsigned long long seed, unsigned long long inst, int draw, double lo, double hi) {
    return __dadd_rn(lo, __dmul_rn(__dsub_rn(hi, lo), sy_u01(seed, inst, draw)));
}
// pose on a lane table at arc length s: linear interpolation of the line's own samples
__device__ __forceinline__ int sy_lane_pose(const SynthLanes& L, int lane, double s, double* px, double* py, double* pyaw) {
    const int o = L.off[lane], M = L.off[lane + 1] - o;
    const double* lon = L.lon + o;
    long long i = (long long)floor(__dmul_rn(__dsub_rn(s, lon[0]), 10.0));
    i = i < 0 ? 0 : (i > M - 2 ? M - 2 : i);
    const double f = __dmul_rn(__dsub_rn(s, lon[i]), 10.0);
    const double *x = L.x + o, *y = L.y + o, *yaw = L.yaw + o;
    *px = __dadd_rn(x[i], __dmul_rn(f, __dsub_rn(x[i + 1], x[i])));
    *py = __dadd_rn(y[i], __dmul_rn(f, __dsub_rn(y[i + 1], y[i])));
    *pyaw = __dadd_rn(yaw[i], __dmul_rn(f, __dsub_rn(yaw[i + 1], yaw[i])));
    return int(i);
}
__device__ __forceinline__ void sy_ego(const cilqr_synth_template_t& st, const SynthLanes& L, unsigned long long seed,
                                       unsigned long long id, double x0[4], double* ref_velo) {
    if (st.ego_kind == 0) {
        x0[0] = sy_uniform(seed, id, 0, -5.0, 5.0);
        x0[1] = sy_uniform(seed, id, 1, -0.6, 0.6);
        x0[2] = sy_uniform(seed, id, 2, 5.0, 10.0);
        x0[3] = sy_uniform(seed, id, 3, -0.05, 0.05);
        *ref_velo = sy_uniform(seed, id, 4, 6.0, 10.0);
    } else {
        const int o = L.off[st.ego_lane], M = L.off[st.ego_lane + 1] - o;
        const double s = fmin(fmax(__dadd_rn(st.ego_s, sy_uniform(seed, id, 0, -5.0, 5.0)), L.lon[o]), L.lon[o + M - 1]);
        const double dl = sy_uniform(seed, id, 1, -0.6, 0.6);
        double px, py, pyaw;
        const int i = sy_lane_pose(L, st.ego_lane, s, &px, &py, &pyaw);
        x0[0] = __dadd_rn(px, __dmul_rn(dl, L.nx[o + i]));
        x0[1] = __dadd_rn(py, __dmul_rn(dl, L.ny[o + i]));
        x0[2] = fmax(__dadd_rn(st.ego_v, sy_uniform(seed, id, 2, -2.0, 2.0)), 0.5);
        x0[3] = __dadd_rn(pyaw, sy_uniform(seed, id, 3, -0.05, 0.05));
        *ref_velo = __dadd_rn(st.target_velocity, sy_uniform(seed, id, 4, -2.0, 2.0));
    }
}

template <typename T>
__global__ void __launch_bounds__(128) k_synth_instances(Dev<T> D, int B, unsigned long long first_id, unsigned long long seed,
                                                         int n_tmpl, const cilqr_synth_template_t* __restrict__ tm, SynthLanes L) {
    const size_t Bs = D.Bs;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
        const unsigned long long id = first_id + (unsigned long long)b;
        const int ti = int(id % (unsigned long long)n_tmpl);
        const cilqr_synth_template_t& st = tm[ti];
        double x0[4], rv;
        sy_ego(st, L, seed, id, x0, &rv);
#pragma unroll
        for (int c = 0; c < 4; ++c) D.x0[size_t(c) * Bs + b] = T(x0[c]);
        D.ref_velo[b] = T(rv);
        D.borders[b] = T(st.borders[0]);
        D.borders[Bs + b] = T(st.borders[1]);
        D.tmpl[b] = ti;
        D.n_obs[b] = st.n_obs;
    }
}

// one thread per (instance, obstacle j, tick k): grid.y = max_obs * (N + 1) rows
template <typename T>
__global__ void __launch_bounds__(128) k_synth_obstacles(Dev<T> D, T* obs, int B, unsigned long long first_id,
                                                         unsigned long long seed, int n_tmpl,
                                                         const cilqr_synth_template_t* __restrict__ tm, SynthLanes L) {
    const size_t Bs = D.Bs;
    const int row = blockIdx.y, j = row / (D.N + 1), k = row % (D.N + 1);
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
        const unsigned long long id = first_id + (unsigned long long)b;
        const cilqr_synth_template_t& st = tm[int(id % (unsigned long long)n_tmpl)];
        double ox = 0, oy = 0, oyaw = 0;
        if (j < st.n_obs) {
            const cilqr_synth_obstacle_t& ob = st.obs[j];
            const double t = __dmul_rn(double(k), st.dt);
            if (ob.kind == 0) {
                const int o = L.off[ob.lane], M = L.off[ob.lane + 1] - o;
                const double lon0 = L.lon[o], lon1 = L.lon[o + M - 1];
                const double ds = sy_uniform(seed, id, ob.draw, -8.0, 8.0);
                const double v = fmax(__dadd_rn(ob.speed, sy_uniform(seed, id, ob.draw + 1, -1.0, 1.0)), 0.5);
                const double s0 = __dadd_rn(ob.start_s, ds), tv = __dmul_rn(t, v);
                const double s = ob.oncoming ? fmin(fmax(__dsub_rn(s0, tv), lon0), lon1) : fmax(fmin(__dadd_rn(s0, tv), lon1), lon0);
                sy_lane_pose(L, ob.lane, s, &ox, &oy, &oyaw);
                if (ob.oncoming) oyaw = fmod(__dadd_rn(oyaw, 3.141592653589793), 6.283185307179586);
            } else {
                ox = sy_uniform(seed, id, ob.draw, ob.x_lo, ob.x_hi);
                if (ob.rel_to_ego) {
                    double x0[4], rv;
                    sy_ego(st, L, seed, id, x0, &rv);
                    ox = __dadd_rn(x0[0], ox);
                }
                const double v = sy_uniform(seed, id, ob.draw_v, ob.v_lo, ob.v_hi);
                oy = (ob.two_lanes && !(sy_u01(seed, id, ob.draw_lane) < 0.5)) ? ob.y1 : ob.y0;
                ox = __dadd_rn(ox, __dmul_rn(ob.direction, __dmul_rn(t, v)));
                oyaw = ob.yaw;
            }
        }
        T* p = obs + (size_t(j) * (D.N + 1) + k) * 4 * Bs + b;
        p[0] = T(ox);
        p[Bs] = T(oy);
        p[2 * Bs] = T(oyaw);
        p[3 * Bs] = T(0);
    }
}

// Writes a buffer larger than L2 so that the next timed launch starts cold.
__global__ void k_flush_l2(float* buf, size_t n) {
    size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    size_t stride = size_t(gridDim.x) * blockDim.x;
    for (; i < n; i += stride) buf[i] = buf[i] * 0.5f + 1.0f;
}

}  // namespace cilqr
