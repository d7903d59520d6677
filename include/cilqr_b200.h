/*
 * cilqr_b200.h — C ABI of libcilqr_b200.so: the batched, B200-native (sm_100a)
 * replacement for the CILQR hot path of PuYuuu/toy-example-of-iLQR.
 *
 * Drop-in boundary.  The reference's only public solver surface is
 *     CILQRSolver::CILQRSolver(const GlobalConfig*)        include/cilqr_solver.hpp:34, src/cilqr_solver.cpp:17-83
 *     CILQRSolver::solve(x0, ref_waypoints, ref_velo,
 *                        obs_preds, road_boaders) -> (u,x)  include/cilqr_solver.hpp:37-41, src/cilqr_solver.cpp:85-153
 * called from src/motion_planning.cpp:178 and :194-196.  Everything below is
 * what a binding for that surface needs: plain pointers and sizes, no C++ or
 * torch types.  The C++ compat class (toy-example-of-ilqr_b200/host/cilqr_solver_compat.hpp)
 * and the ctypes binding (toy-example-of-ilqr_b200/binding.py) sit on top of it;
 * INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *  - every call returns 0 on success and a negative cilqr_error_t otherwise;
 *    cilqr_b200_last_error() returns the message of the calling thread's last
 *    failure.  Nothing throws across this boundary.
 *  - "host layout" = the reference's per-solve matrices, row-major, batch
 *    outermost, always double:  x0 [B][4], u [B][N][2], x [B][N+1][4],
 *    K [B][N][2][4] (rows 2i,2i+1 of the reference's (2N)x4 K), d [B][N][2],
 *    obs [B][max_obs][obs_len][3] = RoutingLine (x,y,yaw) of obstacle j at
 *    tick k (src/utils.cpp:52-58), borders [B][2] = road_boaders.
 *  - a handle is bound to one device and one stream and is not re-entrant
 *    (the reference object is not thread-safe either); use one handle per GPU.
 */
#ifndef CILQR_B200_H
#define CILQR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CILQR_B200_MAX_TEMPLATES 8
#define CILQR_B200_NUM_ALPHAS 20 /* alpha = 1, 1/2, ... > 1e-6  (src/cilqr_solver.cpp:354) */

/* The scalars the reference constructor reads from GlobalConfig
 * (src/cilqr_solver.cpp:18-72).  nx = 4 and nu = 2 are fixed by the model. */
typedef struct cilqr_params_t {
    double dt;                      /* delta_t */
    double w_pos, w_vel, w_yaw;     /* lqr/w_pos (twice), w_vel, w_yaw -> state_weight */
    double w_acc, w_stl;            /* ctrl_weight */
    double obstacle_exp_q1, obstacle_exp_q2, state_exp_q1, state_exp_q2;
    double alm_rho_init, alm_gamma, max_rho, max_mu;
    double init_lamb, lamb_decay, lamb_amplify, max_lamb;
    double convergence_threshold, accept_step_threshold;
    double wheelbase, width, length;
    double velo_max, velo_min, yaw_lim /* read, never used: cpp:67 */, acc_max, acc_min, stl_lim, d_safe;
    int32_t max_iter;
    int32_t solve_type;        /* 0 = "barrier", 1 = "alm"   (lqr/slove_type, cpp:33-41) */
    int32_t reference_point;   /* 0 = "rear_center", 1 = "gravity_center" (cpp:72-76) */
    int32_t use_last_solution; /* lqr/use_last_solution (cpp:32) */
} cilqr_params_t;

typedef enum cilqr_dtype_t { CILQR_F64 = 0, CILQR_F32 = 1 } cilqr_dtype_t;

/* LQRSolveStatus, include/cilqr_solver.hpp:23-29 (same numbering). */
typedef enum cilqr_status_t {
    CILQR_RUNNING = 0,
    CILQR_CONVERGED = 1,
    CILQR_BACKWARD_PASS_FAIL = 2,
    CILQR_FORWARD_PASS_FAIL = 3,
    CILQR_FORWARD_PASS_SMALL_STEP = 4
} cilqr_status_t;

/* How solve() left its loop (src/cilqr_solver.cpp:127-148). */
typedef enum cilqr_exit_t { CILQR_EXIT_MAX_ITER = 0, CILQR_EXIT_CONVERGED = 1, CILQR_EXIT_MAX_LAMB = 2 } cilqr_exit_t;

typedef enum cilqr_error_t {
    CILQR_OK = 0,
    CILQR_ERR_INVALID = -1,   /* bad argument */
    CILQR_ERR_RANGE = -2,     /* obstacle track shorter than N+1: the reference throws std::out_of_range (src/utils.cpp:53-55) */
    CILQR_ERR_CUDA = -3,      /* CUDA runtime failure, message in last_error */
    CILQR_ERR_NO_DEVICE = -4  /* no usable sm_100 device: the product has no CPU fallback */
} cilqr_error_t;

typedef struct cilqr_handle cilqr_handle_t;

const char* cilqr_b200_last_error(void);
const char* cilqr_b200_version(void);

/* Replaces the constructor (cpp:17-83).  Allocates every device buffer for
 * max_batch problems of horizon N with up to max_obs obstacles each; nothing
 * is allocated by later calls.  `params` becomes template 0.  max_batch <= 8388480 per handle. */
int cilqr_b200_create(const cilqr_params_t* params, int device, int max_batch, int N, int max_obs,
                      int dtype, cilqr_handle_t** out);
int cilqr_b200_destroy(cilqr_handle_t* h);

/* Launch stream (a cudaStream_t); NULL = the handle's own stream. */
int cilqr_b200_set_stream(cilqr_handle_t* h, void* cuda_stream);

/* Scenario template t: solver scalars and the reference line's waypoints
 * (ReferenceLine::{x,y,yaw}, include/utils.hpp:44-46; M <= 65535 because the
 * reference indexes them with uint16_t, cpp:291-292).  Either part may be NULL
 * to keep what is there. */
int cilqr_b200_set_template(cilqr_handle_t* h, int tmpl, const cilqr_params_t* params, const double* wx,
                            const double* wy, const double* wyaw, int M);

/* Forget the warm start: every instance behaves as is_first_solve == true (cpp:17, :97-102). */
int cilqr_b200_reset(cilqr_handle_t* h);

/* Replaces solve() (cpp:85-153) for B independent problems, host buffers in,
 * host buffers out; the copies are part of the call.  tmpl (NULL = all 0) and
 * n_obs select template and obstacle count per problem.  Any output may be
 * NULL.  J_out [B][2] = {cost of the initial trajectory (the `J` the reference
 * logs, cpp:104), cost of the returned trajectory}; step_cost_out [B][N+1];
 * K_out / d_out = gains of the last backward pass (cpp:343); status_out =
 * current_solve_status at return; iters_out = iter_step calls made. */
int cilqr_b200_solve_batch(cilqr_handle_t* h, int B, const double* x0, const double* ref_velo,
                           const double* borders, const int32_t* tmpl, const int32_t* n_obs,
                           const double* obs, int obs_len, double* u_out, double* x_out, double* J_out,
                           double* K_out, double* d_out, double* step_cost_out, int32_t* status_out,
                           int32_t* iters_out, int32_t* exit_out);

/* The same in three steps, so that a caller can keep problems resident in HBM
 * and time the solve alone (bench.py `value`). */
int cilqr_b200_upload(cilqr_handle_t* h, int B, const double* x0, const double* ref_velo,
                      const double* borders, const int32_t* tmpl, const int32_t* n_obs, const double* obs,
                      int obs_len);
int cilqr_b200_solve_resident(cilqr_handle_t* h, int B);
int cilqr_b200_download(cilqr_handle_t* h, int B, double* u_out, double* x_out, double* J_out, double* K_out,
                        double* d_out, double* step_cost_out, int32_t* status_out, int32_t* iters_out,
                        int32_t* exit_out);

/* The step after the path: the receding-horizon loop of src/motion_planning.cpp:180-197
 * for B independent scenarios, entirely on the device.  Per tick i = 0 .. ticks-1, with t = 0, dt, 2 dt, ...
 * accumulated in floating point and index = size_t(t / delta_t) exactly as the reference computes it
 * (motion_planning.cpp:180-181; for delta_t = 0.1 the sequence is 0,1,2,3,4,5,5,6,...):
 *     (u, x) = solve(ego_state, ref line, ref_velo, get_sub_routing_lines(tracks, index), borders);
 *     ego_state = x.row(1);
 * with the warm start of lqr/use_last_solution carried between ticks.  tracks [B][max_obs][track_len][3]
 * are the full obstacle tracks (a tick reads samples index .. index+N; a track too short for the last tick is
 * CILQR_ERR_RANGE, the reference's std::out_of_range).  ego_out [B][ticks+1][4] (ego_out[.][0] = x0),
 * iters_out / status_out [B][ticks] may be NULL.  Track and history buffers are (re)allocated here when a
 * longer simulation than any before is requested. */
int cilqr_b200_simulate(cilqr_handle_t* h, int B, const double* x0, const double* ref_velo, const double* borders,
                        const int32_t* tmpl, const int32_t* n_obs, const double* tracks, int track_len, int ticks,
                        double* ego_out, int32_t* iters_out, int32_t* status_out);

/* ---- synthetic workloads (BASELINE.json configs C1..C4), generated in place on the device ----
 * A batch is described by one small descriptor per scenario template (instance id i uses template
 * i % n_tmpl) plus the centre lines the descriptors refer to; every number is a function of
 * (seed, instance id, draw index) through a counter-based generator (splitmix64 finaliser), so any slice
 * [first_id, first_id + B) of a batch can be generated on any GPU without data crossing PCIe.  The arithmetic
 * is spelled out step by step in toy-example-of-ilqr_b200/scenario.py (generate_host), which produces the same
 * arrays with numpy, bit for bit (tests/test_gpu_synth.py). */
#define CILQR_B200_SYNTH_MAX_OBS 16
typedef struct cilqr_synth_obstacle_t {
    int32_t kind;      /* 0: follows lane table `lane` from start_s + U(draw; -8, 8) at max(speed + U(draw+1; -1, 1), 0.5),
                          towards decreasing s with yaw + pi when `oncoming` (src/motion_planning.cpp:149-158);
                          1: x = [ego x0 +] U(draw; x_lo, x_hi) + direction * v t, v = U(draw_v; v_lo, v_hi),
                             y = y0 (or y1 when two_lanes and u01(draw_lane) >= 0.5), constant yaw */
    int32_t lane, oncoming, draw;
    double start_s, speed;
    double x_lo, x_hi, v_lo, v_hi, y0, y1, yaw, direction;
    int32_t rel_to_ego, two_lanes, draw_lane, draw_v;
} cilqr_synth_obstacle_t;
typedef struct cilqr_synth_template_t {
    int32_t ego_kind;  /* 0: x0 = (U0(-5,5), U1(-.6,.6), U2(5,10), U3(-.05,.05)), ref_velo = U4(6,10);
                          1: lane frame of ego_lane at ego_s + U0(-5,5), lateral U1(-.6,.6), v = max(ego_v + U2(-2,2), .5),
                             yaw = lane yaw + U3(-.05,.05), ref_velo = target_velocity + U4(-2,2) */
    int32_t ego_lane, n_obs, reserved;
    double ego_s, ego_v, target_velocity, borders[2], dt;
    cilqr_synth_obstacle_t obs[CILQR_B200_SYNTH_MAX_OBS];
} cilqr_synth_template_t;
/* Centre-line tables: lane l holds samples lane_off[l] .. lane_off[l+1]-1 of x, y, yaw, arc length and the
 * left normal (nx, ny) = (-sin yaw, cos yaw) of each sample (ReferenceLine samples, src/utils.cpp:21-35). */
int cilqr_b200_synth_set_lanes(cilqr_handle_t* h, int n_lanes, const int32_t* lane_off, const double* x,
                               const double* y, const double* yaw, const double* lon, const double* nx,
                               const double* ny);
/* Generates instances [first_id, first_id + B) into the handle (as cilqr_b200_upload would leave them); follow
 * with cilqr_b200_solve_resident.  keep_yaw != 0 leaves the obstacle yaw unconverted for cilqr_b200_synth_download
 * (verification only: such a batch must not be solved). */
int cilqr_b200_synth_generate(cilqr_handle_t* h, int B, uint64_t first_id, uint64_t seed, int n_tmpl,
                              const cilqr_synth_template_t* tmpls, int keep_yaw);
/* The resident problem data back in host layout; obs [B][max_obs][N+1][4] = the four device fields per sample:
 * (x, y, sin yaw, cos yaw), or (x, y, yaw, 0) after a keep_yaw generation.  Any pointer may be NULL. */
int cilqr_b200_synth_download(cilqr_handle_t* h, int B, double* x0, double* ref_velo, double* borders, int32_t* tmpl,
                              int32_t* n_obs, double* obs);

/* Counters of the last solve: total iter_step calls over the batch, line-search
 * trials (forward pass + cost) evaluated, device rounds run, kernels launched,
 * instances per exit reason. */
typedef struct cilqr_counters_t {
    int64_t total_iters;
    int64_t total_trials;
    int32_t rounds;
    int32_t launches;
    int32_t exits[3];
    int32_t reserved;
} cilqr_counters_t;
int cilqr_b200_counters(cilqr_handle_t* h, cilqr_counters_t* out);

/* Execution options; none of them changes a result bit.
 *  CILQR_OPT_WIDE_SEARCH (default 1): instances in a streak of rejected steps evaluate all
 *      remaining alphas of iter_step's line search (cpp:354-372) in one device round and
 *      keep the first that passes; 0 evaluates one alpha per round.
 *  CILQR_OPT_RUN_AHEAD (default 3): device rounds the host may queue beyond the last one
 *      whose active count it has seen.
 *  CILQR_OPT_PREFETCH_BELOW (default 16384, the measured crossover): batches up to this size run the backward pass
 *      with next-step operands prefetched into registers (latency-bound regime); larger
 *      batches use the leaner streaming variant (bandwidth-bound regime).
 *  CILQR_OPT_BENCH_PREFETCH (default -1): which backward kernel cilqr_b200_bench_backward times and
 *      cilqr_b200_stage_backward runs: -1 the one the solver uses at that batch, 0 streaming,
 *      1 register prefetch, 2 staged (below).
 *  CILQR_OPT_PROFILE_STAGES (default 0): record a CUDA event in front of every stage launch of the
 *      following solves (adds a few microseconds per round; not for timed runs).
 *  CILQR_OPT_PIPELINE (default 1): latency-bound batches run forward_pass and the waypoint match
 *      of the new trajectory as one two-stage kernel (the match trails the rollout through a
 *      shared-memory ring); 0 runs them as two kernels; 8 / 16 force the scan window.
 *  CILQR_OPT_STAGED_BACKWARD (default 1): latency-bound batches run the backward pass with one warp
 *      per tile of 32 instances, each step's record brought into a shared-memory ring by bulk
 *      asynchronous copies (cp.async.bulk + mbarrier) three steps ahead of the recursion.
 *  CILQR_OPT_REPACK (default 1): whenever the instances still running are down to half of the slots in
 *      use (batches above 4096), they are moved into a dense prefix of the device arrays and the
 *      solve carries on as a batch of that size; every instance is moved back into its own slot
 *      before the solve returns.  0 disables; a value > 1 sets the smallest batch still repacked.
 *  CILQR_OPT_WIDE_STEP (default 1): in bandwidth-bound rounds an instance in a streak of rejected steps
 *      evaluates 2, 4, 8, 6 more alphas per round instead of all that remain (same decisions, ~18 % fewer
 *      trials; latency-bound rounds keep evaluating all of them at once).
 *  CILQR_OPT_LOOKAHEAD (default 1): look-ahead rounds.  The derivatives and the backward pass of the NEXT iteration
 *      run one device round early, on a second stream next to the cost and verdict kernels of the line search they
 *      depend on — one job per trial ("the backward pass that follows if this trial is the accepted one", cpp:362-366;
 *      96 % of the iterations of the slowest instances accept the full step) plus one for the all-rejected outcome —
 *      and the job the verdict asks for is adopted; an iteration then costs rollout + derivatives + recursion
 *      instead of the whole chain.  1: batches of up to 512 instances (measured 5-9 % faster there, slower above);
 *      n > 1: batches of up to n instances, and larger ones switch to look-ahead rounds once no more than n
 *      instances are still running (n <= 16384; handles created for more than 16384 instances never use them);
 *      0: the stages of an iteration always run one after the other.
 *  CILQR_OPT_FUSED_BACKWARD (default 1): in bandwidth-bound rounds of a barrier-type solve the backward pass computes
 *      the control half of every derivative record (l_u, l_uu, A_k, B_k: cpp:540-610, utils.cpp:285-342) from
 *      (v_k, yaw_k, u_k) itself instead of reading what the derivative stage wrote: 14 of the 28 record fields are
 *      neither written nor read back, and the control half of the derivative stage is not launched.  0: the
 *      derivative stage writes whole records in every round.  Also: CILQR_OPT_BENCH_PREFETCH = 3 selects this
 *      flavour for cilqr_b200_bench_backward / cilqr_b200_stage_backward (which then need the trajectory loaded).
 *  The regime threshold (CILQR_OPT_PREFETCH_BELOW) is applied per round to the number of instances still
 *      running, so a large batch moves to the latency-regime kernels for its stragglers. */
typedef enum cilqr_option_t {
    CILQR_OPT_WIDE_SEARCH = 0,
    CILQR_OPT_RUN_AHEAD = 1,
    CILQR_OPT_PREFETCH_BELOW = 2,
    CILQR_OPT_BENCH_PREFETCH = 3,
    CILQR_OPT_PROFILE_STAGES = 4,
    CILQR_OPT_PIPELINE = 5,
    CILQR_OPT_STAGED_BACKWARD = 6,
    CILQR_OPT_REPACK = 7,
    CILQR_OPT_WIDE_STEP = 8,
    CILQR_OPT_LOOKAHEAD = 9,
    CILQR_OPT_FUSED_BACKWARD = 10
} cilqr_option_t;
int cilqr_b200_set_option(cilqr_handle_t* h, int option, int value);

/* Per-stage device time of the last solve run with CILQR_OPT_PROFILE_STAGES: ms_out [6] and
 * launches_out [6] (may be NULL) for {derivatives, backward pass, forward pass, waypoint match,
 * cost, verdict}, measured between consecutive stage events on the launch stream. */
int cilqr_b200_stage_times(cilqr_handle_t* h, double* ms_out, int32_t* launches_out);

/* Per-iteration decision trace (lockstep tests): record the first `cap` iter_step outcomes of
 * every instance of subsequent solves; get_trace returns status [B][cap] (cilqr_status_t after
 * each iter_step), alpha [B][cap] (index of the accepted / converged alpha, -1 if none) and
 * cost [B][cap] (cost iter_step returned).  Rows beyond iters_out[b] are undefined. */
int cilqr_b200_enable_trace(cilqr_handle_t* h, int cap);
int cilqr_b200_get_trace(cilqr_handle_t* h, int B, int32_t* status, int32_t* alpha, double* cost);

/* ---- stage operators (one per reference function; host layout in and out) ----
 * Used by the parity tests and by bench.py's roofline leg.  Each uploads its
 * inputs, runs exactly the kernel(s) the solve uses for that stage, and
 * downloads the result. */

/* get_init_traj / get_init_traj_increment (cpp:155-197): warm != 0 shifts last_u [B][N][2]. */
int cilqr_b200_stage_init(cilqr_handle_t* h, int B, const double* x0, const int32_t* tmpl, int warm,
                          const double* last_u, double* u_out, double* x_out);
/* get_ref_exact_points (cpp:289-314): indices of the matched waypoints, [B][N+1]. */
int cilqr_b200_stage_ref_match(cilqr_handle_t* h, int B, const double* x, const int32_t* tmpl, int32_t* idx_out);
/* get_total_cost (cpp:199-287): J [B], step_cost [B][N+1].  alm_mu [B][N][8+2*max_obs], alm_rho [B] only in ALM mode. */
int cilqr_b200_stage_cost(cilqr_handle_t* h, int B, const double* u, const double* x, const double* ref_velo,
                          const double* borders, const int32_t* tmpl, const int32_t* n_obs, const double* obs,
                          int obs_len, const double* alm_mu, const double* alm_rho, double* J_out,
                          double* step_cost_out);
/* get_total_cost_derivatives_and_Hessians (cpp:463-690) + get_kinematic_model_derivatives
 * (src/utils.cpp:285-342), dense reference layout out: lx [B][N+1][4], lu [B][N][2],
 * lxx [B][N+1][4][4], luu [B][N][2][2], A [B][N][4][4], Bm [B][N][4][2]. */
int cilqr_b200_stage_derivs(cilqr_handle_t* h, int B, const double* u, const double* x, const double* ref_velo,
                            const double* borders, const int32_t* tmpl, const int32_t* n_obs, const double* obs,
                            int obs_len, const double* alm_mu, const double* alm_rho, double* lx, double* lu,
                            double* lxx, double* luu, double* A, double* Bm, double* alm_mu_next);
/* The Riccati recursion of backward_pass (cpp:391-439) on explicit inputs (dense layout in);
 * d [B][N][2], K [B][N][2][4], dV [B][2], status [B] (RUNNING or BACKWARD_PASS_FAIL). */
int cilqr_b200_stage_backward(cilqr_handle_t* h, int B, const double* lx, const double* lu, const double* lxx,
                              const double* luu, const double* A, const double* Bm, const double* lamb,
                              double* d_out, double* K_out, double* dV_out, int32_t* status_out);
/* forward_pass (cpp:442-461) for one alpha per problem. */
int cilqr_b200_stage_forward(cilqr_handle_t* h, int B, const double* u, const double* x, const double* d,
                             const double* K, const double* alpha, const int32_t* tmpl, double* new_u,
                             double* new_x);

/* Roofline leg: run the backward-pass kernel `reps` times on the derivative
 * records currently resident in the handle (left there by stage_derivs /
 * stage_backward / a solve), timing each launch with CUDA events on the launch
 * stream.  ms_out [reps] per-launch milliseconds; bytes_per_launch = algorithmic
 * bytes of the compact record layout the kernel moves, (38*N + 18) * sizeof(T) * B. */
int cilqr_b200_bench_backward(cilqr_handle_t* h, int B, double lamb, int reps, int flush_l2, float* ms_out,
                              double* bytes_per_launch);
/* Replicate the first B0 resident derivative records up to B (device-side copy), so the
 * roofline leg can run at batch sizes whose inputs were produced from a smaller solve. */
int cilqr_b200_bench_tile_records(cilqr_handle_t* h, int B0, int B);

#ifdef __cplusplus
}
#endif
#endif /* CILQR_B200_H */
