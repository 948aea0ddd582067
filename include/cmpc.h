/*
 * cmpc.h -- C ABI of the batched, B200-native centroidal-MPC condensed-QP solver.
 *
 * Drop-in boundary for ONE path of HuNingHe/Cheeta-MPC: the build + solve that
 * `CentroidalMPC::UpdateMPC` performs per tick (reference CentroidalMPC.cpp:278-370),
 * restated as the condensed QP of SURVEY.md §8(a) and solved for B independent
 * instances per call on one B200.  Plain pointers and sizes only; no C++/torch types.
 *
 * Each entry point cites the reference interface it replaces (file:line, relative to
 * the reference root).  The C++ shim with the reference's class/method names lives in
 * cheeta-mpc_b200/include/CentroidalMPC.h; INTEGRATION.md shows the binding.
 *
 * Layouts (all fp64, instance-major outer dimension, reference packing inside):
 *   state      [B][9+3L]     c(3) v(3) Lm(3) p_i(3)...            CentroidalMPC.cpp:284-291
 *   des_state  [B][9(N+1)]   3 col-major 3x(N+1) blocks pos|vel|am CentroidalMPC.cpp:297-299
 *   des_inputs [B][L(4N+3)]  per leg: contact(N) | des_foot_pos 3x(N+1) col-major  :316-317
 *   forces     [B][L][N][3]  per leg 3xN col-major = the reference's `contact_force_i`
 *                            controller outputs                    CentroidalMPC.cpp:270
 *   H          [B][p][p], g [B][p]  with p = 3LN and U index 3L*j + 3*i + r (step-major)
 *   lam        [B][2][N][L][5] multipliers of the lower (0 <= F f) then upper rows
 *   active     [B][N][L] uint16: bit r (0..4) lower row r active, bit 5+r upper row r
 *                            active, bit 15 = leg pinned (swing, contact <= 0)
 */
#ifndef CMPC_H_
#define CMPC_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CMPC_MAX_LEGS 4
#define CMPC_NUM_WEIGHTS (9 + 9 * CMPC_MAX_LEGS)
#define CMPC_MAX_HORIZON 32

/* per-instance status words (replace the reference's exceptions) */
enum {
  CMPC_STATUS_OK = 0,            /* KKT point, active set verified by the polish step     */
  CMPC_STATUS_OK_IPM = 1,        /* interior-point converged to ipm_tol, polish not accepted */
  CMPC_STATUS_MAX_ITER = 2,      /* iteration cap hit (IPOPT "Maximum_Iterations" analogue) */
  CMPC_STATUS_INVALID_TABLE = 3, /* a horizon step with no stance leg: the reference throws
                                    "mpc table invalid", CentroidalMPC.cpp:328-330         */
  CMPC_STATUS_NUMERICAL = 4      /* non-finite input or factorisation breakdown           */
};

/* return codes of the entry points */
enum {
  CMPC_OK = 0,
  CMPC_ERR_ARG = -1,     /* bad argument (the reference asserts, CentroidalMPC.cpp:24-25) */
  CMPC_ERR_CUDA = -2,    /* CUDA runtime error; text via cmpc_last_error                 */
  CMPC_ERR_STATE = -3,   /* call order (solve before setup) or batch > max_batch         */
  CMPC_ERR_NO_DEVICE = -4
};

/* Constructor arguments of the reference class + solver knobs.
 * Replaces: CentroidalMPC::CentroidalMPC(mass, num_legs, predict_horizon, time_step,
 * weights, mu, ipopt_solver) CentroidalMPC.h:26-27 / CentroidalMPC.cpp:13-32 and the IPOPT
 * option block NonlinearMPC.h:84-100 (tol 1e-8 there). */
typedef struct cmpc_config {
  double mass;
  int32_t num_legs;                 /* 1..CMPC_MAX_LEGS */
  int32_t horizon;                  /* N, 1..CMPC_MAX_HORIZON */
  double dt;
  double mu[CMPC_MAX_LEGS];         /* friction coefficient per leg, > 0 */
  double weights[CMPC_NUM_WEIGHTS]; /* 9+9L used; index map = CentroidalMPC.cpp:208-231 */
  int32_t disc_mode;                /* 0 explicit Euler (reference :90-92), 1 zero-order hold */
  int32_t max_iter;                 /* IPM iteration cap (default 50) */
  double ipm_tol;                   /* scaled residual + gap tolerance (default 1e-9) */
  int32_t polish;                   /* 1 = active-set polish (default), 0 = IPM only */
  int32_t presolve;                 /* 1 (default, needs polish) = first try the unconstrained minimiser
                                       -H^-1 g: one Cholesky of H; if it satisfies every friction /
                                       force-limit row it is the optimum (status OK, iters 0), else
                                       the instance goes through the interior-point iteration */
  int32_t qp_backend;               /* presolve back-end per size class:
                                       0 (default) = condensed dense Cholesky for instances with up to 20 free
                                           leg-steps (n <= 60: trot at horizon 10), stage-wise Riccati sweep over
                                           [x; F_prev] above (stand, horizon 30: O(N 21^3) instead of O((3LN)^3));
                                       1 = dense for every class;  2 = Riccati for every class */
  int32_t reserved;
} cmpc_config;

typedef struct cmpc_stats {
  double kernel_ms;      /* CUDA-event time of the device work of the last call */
  double h2d_ms, d2h_ms; /* copy phases of cmpc_solve_batch (0 for *_device calls) */
  double mean_iters;     /* IPM iterations (= Cholesky factorisations of H + C'SC) */
  int32_t max_iters;
  int32_t n_ok, n_ok_ipm, n_max_iter, n_invalid, n_numerical;
  double max_kkt;        /* max scaled KKT residual over instances with status <= 1 */
  int32_t launches;      /* kernels launched by the call */
  int32_t reserved;
} cmpc_stats;

typedef struct cmpc_handle cmpc_handle;

/* Fill cfg with the reference ctor arguments and default solver knobs.
 * weights has 9+9*num_legs entries, mu has num_legs. */
int cmpc_config_init(cmpc_config* cfg, double mass, int num_legs, int horizon, double dt,
                     const double* weights, const double* mu);

/* Replaces the constructor (CentroidalMPC.cpp:13-32). Validates like its asserts. */
int cmpc_create(const cmpc_config* cfg, cmpc_handle** out);

/* Replaces SetupMPC() (CentroidalMPC.cpp:102-276): one-off allocation of device buffers
 * for up to max_batch instances on CUDA device `device`, upload of constants. No
 * allocation happens afterwards in cmpc_solve_batch[_device], cmpc_solve_batch_sqp, cmpc_foot_plan_batch or the
 * contact-table calls; cmpc_rollout grows its force-log buffer on demand and keeps it; the parity entries
 * (cmpc_build_batch, cmpc_stage_step_batch) and cmpc_measure_fp64_peak use temporaries.  A failed setup releases
 * everything it had allocated, so it can be retried.  Every entry point restores the caller's current device. */
int cmpc_setup(cmpc_handle* h, int max_batch, int device);

/* Replaces NonlinearMPC::UpdateWeights (NonlinearMPC.h:103-105). n = 9+9*num_legs. */
int cmpc_update_weights(cmpc_handle* h, const double* weights, int n);

/* Replaces UpdateMPC(state, des_state, des_inputs) (CentroidalMPC.cpp:278-370) for B
 * instances with HOST buffers: H2D, fused build+solve kernel, D2H.  Optional outputs
 * (iters, kkt, lam, active, stats) may be NULL. */
int cmpc_solve_batch(cmpc_handle* h, int B, const double* state, const double* des_state,
                     const double* des_inputs, double* forces, int32_t* status,
                     int32_t* iters, double* kkt, double* lam, uint16_t* active,
                     cmpc_stats* stats);

/* Same with DEVICE pointers (inputs already resident in HBM, outputs left in HBM);
 * asynchronous on the handle's stream unless stats != NULL (then it synchronises).  The call only enqueues work
 * (kernels -- under stream capture also two memsets --, and -- when a batch has work for both interior-point kernels -- a fork to an auxiliary stream of
 * the handle that is joined again before the call returns control of the stream), so with stats == NULL it can be
 * captured into a CUDA graph on the handle's stream and replayed (tests/test_gpu_parity.py).  The launch plan follows
 * the list counts of the handle's previous call; it changes the time a call takes, never a bit of its results. */
int cmpc_solve_batch_device(cmpc_handle* h, int B, const double* d_state,
                            const double* d_des_state, const double* d_des_inputs,
                            double* d_forces, int32_t* d_status, int32_t* d_iters,
                            double* d_kkt, double* d_lam, uint16_t* d_active,
                            cmpc_stats* stats);

/* The condensed-QP build alone: H = 2(Bqp' L Bqp + K), g, with swing-leg
 * rows/cols pinned to identity.  Host buffers.  For parity tests of the build stage. */
int cmpc_build_batch(cmpc_handle* h, int B, const double* state, const double* des_state,
                     const double* des_inputs, double* H, double* g, int32_t* status);

/* Closed loop (BASELINE config 5): `ticks` MPC ticks on device. Per tick: solve; drive the
 * reference's nonlinear Euler plant (CentroidalMPC.cpp:85-92, true lever arms foot - com) with
 * the first-step forces; rotate the contact table by one step (period = horizon); carry the
 * feet that are in swing under their hips (offsets captured at tick 0); regenerate the
 * references from the new state (constant desired velocity / height / angular momentum taken
 * from node 0 of the inputs).  warm_start != 0: each tick first tries the previous tick's
 * active set (shifted by one step) in the active-set polish and only falls back to the cold
 * interior-point iteration if that guess does not verify.  state / des_state / des_inputs are
 * updated in place (host buffers).  force_log (optional) [ticks][B][3L] first-step forces;
 * iters_sum (optional) [B] factorisations of H + C'SC summed over ticks; status_or (optional)
 * [B] OR of (1 << status) over ticks. */
int cmpc_rollout(cmpc_handle* h, int B, int ticks, int warm_start, double* state,
                 double* des_state, double* des_inputs, double* force_log,
                 int32_t* iters_sum, int32_t* status_or, cmpc_stats* stats);

/* The foot-position half of the reference's outputs (`controller_output.push_back(foot_pos[i])`,
 * CentroidalMPC.cpp:269).  With the lever arms frozen the foot variables decouple from the
 * forces, and their optimum has a closed form per leg and axis:
 *   foot_pos[:,0] = current foot (:166);  foot_pos[:,k+1] = foot_pos[:,k] when 1 - contact_k == 0
 *   (:94), free otherwise;  cost w_p sum_k (foot_pos_k - des_foot_pos_k)^2 (:219-221);  box
 *   step_lb <= foot_pos_k - des_foot_pos_k <= step_ub for k = 1..N, step box +-(0.2, 0.2, 0.1)
 *   (:30-31, :198).  Nodes joined by locked intervals share one value: the current foot for the
 *   group of node 0, else the mean of the desired positions clipped to the intersected boxes.
 * foot_pos [B][L][N+1][3] (each leg 3 x (N+1) column-major like the reference output). Host buffers. */
int cmpc_foot_plan_batch(cmpc_handle* h, int B, const double* state, const double* des_inputs,
                         double* foot_pos);

/* Successive re-linearisation of the frozen lever arms (the model-fidelity number
 * of SURVEY §0).  Iteration 0 is the plain solve (arms = des_foot_pos - des_com_pos).  Every
 * further iteration re-solves with arms = des_foot_pos - c(j), where c is the centre-of-mass
 * path that the previous iteration's forces produce through the reference's NONLINEAR Euler
 * plant (CentroidalMPC.cpp:85-92, lever arm foot - com with com a trajectory variable, :86).
 * At a fixed point the condensed linear model reproduces the reference dynamics exactly.
 * defect [B][sqp_iters + 1] (may be NULL): max_j |x_nonlinear(j) - x_linear(j)|_inf of every
 * iteration's solution, i.e. how far the QP's prediction is from the reference plant.
 * forces / status are those of the last iteration.  Host buffers. */
int cmpc_solve_batch_sqp(cmpc_handle* h, int B, int sqp_iters, const double* state,
                         const double* des_state, const double* des_inputs, double* forces,
                         int32_t* status, double* defect);

/* Parity entry for the stage-wise (Riccati) linear algebra of the interior-point / polish kernel (SURVEY §8 f3;
 * stage interface as in ocs2_sqp/hpipm_catkin/src/HpipmInterface.cpp:166-301).  For every instance it solves
 *     minimise 1/2 d'(H + C' diag(sigma) C) d - rhs' d
 * mode 1: hess [B][N][L][6] = the entries (xx, yy, zz, zx, zy, unused) of 1/2 F_i' diag(sigma) F_i per leg-step;
 * mode 2: sigma = 0 and d restricted per leg-step to the range of the symmetric projector given as
 *         (00, 11, 22, 10, 20, 21) -- the polish's equality-constrained system.
 * d_fused: right-hand side carried by the factor sweep; d_resolve (mode 1 only): the same right-hand side through the
 * stored factors; grad = H rhs + g (rhs taken as a force vector).  All force vectors [B][N][L][3]. Host buffers. */
int cmpc_stage_step_batch(cmpc_handle* h, int B, int mode, const double* state, const double* des_state,
                          const double* des_inputs, const double* hess, const double* rhs, double* d_fused,
                          double* d_resolve, double* grad);

/* Gait template = ModeSequenceTemplate of the vendored OCS2 stack (ocs2_legged_robot/
 * config/command/gait.info; src/gait/ModeSequenceTemplate.cpp:74-87 converts it to a Gait).
 * modes[] are ModeNumber values 0..15 with bits {LF = 8, RF = 4, LH = 2, RH = 1}
 * (include/ocs2_legged_robot/gait/MotionPhaseDefinition.h:47-64,129-132). */
#define CMPC_MAX_GAIT_MODES 8
typedef struct cmpc_gait {
  int32_t num_modes;
  int32_t modes[CMPC_MAX_GAIT_MODES];
  double switching_times[CMPC_MAX_GAIT_MODES + 1]; /* num_modes + 1 ascending times */
} cmpc_gait;

/* Device-side gait -> contact table: for instance b and step j the contact
 * flags at time t0[b] + j*dt are those of gait gait_id[b]:  phase = wrapPhase(t / duration)
 * (Gait.cpp:63-69), mode = modeSequence[upper_bound(eventPhases, phase)] (Gait.cpp:74-88),
 * stance legs = modeNumber2StanceLeg(mode).  OCS2's leg order {LF, RF, LH, RH} is mapped to
 * the reference driver's {lf, rf, rh, lh} (CentoidMPCTest.cpp:43-46).  Only the contact
 * entries of des_inputs are written.  A FLY mode yields a column without stance leg, which
 * the solve flags CMPC_STATUS_INVALID_TABLE like the reference (CentroidalMPC.cpp:328-330).
 * Host buffers; `_device` takes device pointers for gait_id, t0 and des_inputs. */
int cmpc_fill_contact_tables(cmpc_handle* h, int B, const cmpc_gait* gaits, int num_gaits,
                             const int32_t* gait_id, const double* t0, double* des_inputs);
int cmpc_fill_contact_tables_device(cmpc_handle* h, int B, const cmpc_gait* gaits, int num_gaits,
                                    const int32_t* d_gait_id, const double* d_t0, double* d_des_inputs);

/* Gait SWITCH with the reference stack's stance-insertion rule: for instance b the contact flags at times
 * t0[b] + j*dt come from the mode schedule that GaitSchedule produces when template gait_from[b], tiled from t_tile[b]
 * (tileModeSequenceTemplate, ocs2_legged_robot/src/gait/GaitSchedule.cpp:107-137: initial STANCE, then whole template
 * cycles with event times accumulated by repeated addition), is replaced at t_switch[b] by template gait_to[b] through
 * insertModeSequenceTemplate (GaitSchedule.cpp:47-72): events at or after t_switch are erased, the mode active there runs
 * on until t_switch, an intermediate STANCE phase of `stance_time` (phaseTransitionStanceTime) follows unless that mode
 * already is STANCE, then gait_to is tiled from t_switch (+ stance_time).  The mode at time t is
 * modeSequence[lower_bound(eventTimes, t)] (ModeSchedule::modeAtTime of ocs2_core, an un-vendored dependency: a step that
 * lands exactly on an event keeps the earlier mode).  t_switch[b] = NaN or +-inf: no switch.  Leg order and FLY modes as
 * in cmpc_fill_contact_tables.  Host buffers; `_device` takes device pointers for the per-instance arrays. */
int cmpc_fill_contact_tables_switch(cmpc_handle* h, int B, const cmpc_gait* gaits, int num_gaits,
                                    const int32_t* gait_from, const int32_t* gait_to, const double* t_tile,
                                    const double* t_switch, double stance_time, const double* t0, double* des_inputs);
int cmpc_fill_contact_tables_switch_device(cmpc_handle* h, int B, const cmpc_gait* gaits, int num_gaits,
                                           const int32_t* d_gait_from, const int32_t* d_gait_to, const double* d_t_tile,
                                           const double* d_t_switch, double stance_time, const double* d_t0,
                                           double* d_des_inputs);

/* Use an externally owned CUDA stream (cudaStream_t as void*) for all device work: everything a call does is ordered
 * behind what is already on that stream, and whatever is enqueued on it afterwards is ordered behind the call (the
 * library's own copy / auxiliary streams fork from and join back into it by events).  The stream must belong to the
 * handle's device (checked once the handle is set up).  NULL before cmpc_setup means "let the library create its own
 * non-blocking stream"; NULL after cmpc_setup selects the legacy default stream. */
int cmpc_set_stream(cmpc_handle* h, void* cuda_stream);
int cmpc_synchronize(cmpc_handle* h);

/* FP64 DFMA throughput microbenchmark on the handle's device (TFLOP/s): the roofline
 * denominator the driver's MEASURED_PEAKS.json does not carry. */
int cmpc_measure_fp64_peak(cmpc_handle* h, double* tflops);

void cmpc_destroy(cmpc_handle* h);
const char* cmpc_last_error(const cmpc_handle* h);
/* How the last cmpc_solve_batch call moved its data: "zero-copy", "pipelined x2 chunks", "full-duplex x2 chunks",
 * "staged", "progressive", with "(tuning)" / "(tuned)" while / after the library times its routes. */
const char* cmpc_last_route(const cmpc_handle* h);
const char* cmpc_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CMPC_H_ */
