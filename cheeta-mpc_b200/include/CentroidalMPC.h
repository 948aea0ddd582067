/*
 * CentroidalMPC.h -- C++ host-side mirror of the reference class surface over the C ABI.
 *
 * Same class name, constructor argument order and method names as the reference
 *   class CentroidalMPC : public NonlinearMPC           (reference CentroidalMPC.h:15-33)
 *   NonlinearMPC::UpdateWeights                          (reference NonlinearMPC.h:103-105)
 * so that a driver written against the reference (CentoidMPCTest.cpp) compiles against this
 * header with the include line changed.  Differences, all additive:
 *   - UpdateMPC returns the optimal contact forces (the reference returns {} and prints);
 *     order = the reference's controller outputs contact_force_1..L, each 3 x N column-major
 *     (reference CentroidalMPC.cpp:269-273);
 *   - UpdateMPCBatch solves B independent instances in one call;
 *   - Eigen is optional: spans of doubles are always accepted, Eigen::VectorXd overloads
 *     appear when <eigen3/Eigen/Dense> is available;
 *   - errors: constructor argument violations (the reference's asserts, CentroidalMPC.cpp:24-25)
 *     and CUDA failures throw std::runtime_error; an instance whose contact table has a
 *     column without a stance leg throws std::runtime_error("mpc table invalid") from
 *     UpdateMPC exactly like the reference (:329), and is flagged per instance in the batch call.
 * Threading contract = the reference's: one object, one thread at a time.
 */
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/cmpc.h"

#if defined(__has_include)
#if __has_include(<eigen3/Eigen/Dense>)
#include <eigen3/Eigen/Dense>
#define CMPC_HAVE_EIGEN 1
#endif
#endif

/* kept for source compatibility with the reference (NonlinearMPC.h:17-26); ignored: the QP is
 * solved by the library's own batched interior point, not by IPOPT + HSL */
enum class IPOPT_SOLVER : unsigned int { MUMPS = 0, WSMP = 1, PARDISO = 2, MA27 = 3, MA57 = 4, MA77 = 5, MA86 = 6, MA97 = 7 };

/* solver knobs the reference does not have (cmpc_config); none of them changes the optimum returned */
struct CentroidalMPCOptions {
  int disc_mode = 0;     // 0 explicit Euler (the reference), 1 zero-order hold
  int presolve = 1;      // unconstrained minimiser verified first (one Cholesky of H)
  int qp_backend = 0;    // 0 automatic, 1 condensed dense, 2 stage-wise Riccati presolve
  int polish = 1;        // active-set polish after the interior-point iteration
  int max_iter = 50;
  double ipm_tol = 1e-9;
  // Model knob.  0 (default): the convex QP with the lever arms frozen at des_foot_pos - des_com_pos.  k > 0: k
  // re-linearisations of the arms about the centre-of-mass path of the previous solution (cmpc_solve_batch_sqp), which
  // makes the QP's dynamics reproduce the reference's nonlinear Euler plant (CentroidalMPC.cpp:85-92); on the reference
  // driver's fixture the forces then sit 0.012 % from the optimum of the reference's own NLP instead of 0.6 %
  // (tests/golden/make_nlp_pin.py).  Costs one more solve per iteration.
  int sqp_iters = 0;
};

class CentroidalMPC {
 public:
  struct BatchResult {
    std::vector<double> forces;    // [B][L][N][3]
    std::vector<int32_t> status;   // [B], CMPC_STATUS_*
    std::vector<int32_t> iters;    // [B]
    std::vector<double> kkt;       // [B]
    cmpc_stats stats;
  };

  using Options = CentroidalMPCOptions;

  CentroidalMPC() = delete;
  CentroidalMPC(const CentroidalMPC&) = delete;             // reference NonlinearMPC.h:49-51
  CentroidalMPC& operator=(const CentroidalMPC&) = delete;

  CentroidalMPC(double mass, int num_legs, int predict_horizon, double time_step, const double* weights,
                size_t n_weights, const double* mu, size_t n_mu, IPOPT_SOLVER = IPOPT_SOLVER::MA97, int device = 0,
                const Options& opt = Options())
      : num_legs_(num_legs), horizon_(predict_horizon), device_(device), sqp_iters_(opt.sqp_iters) {
    if (!(mass > 0) || num_legs <= 0 || predict_horizon <= 0)
      throw std::runtime_error("CentroidalMPC: mass > 0 && num_legs > 0 && predict_horizon > 0 required");
    if (n_mu != (size_t)num_legs) throw std::runtime_error("CentroidalMPC: mu.size() == num_legs required");
    if (n_weights < (size_t)(9 + 9 * num_legs)) throw std::runtime_error("CentroidalMPC: weights needs 9 + 9*num_legs entries");
    if (cmpc_config_init(&cfg_, mass, num_legs, predict_horizon, time_step, weights, mu) != CMPC_OK)
      throw std::runtime_error("CentroidalMPC: invalid constructor arguments");
    cfg_.disc_mode = opt.disc_mode; cfg_.presolve = opt.presolve; cfg_.qp_backend = opt.qp_backend;
    cfg_.polish = opt.polish; cfg_.max_iter = opt.max_iter; cfg_.ipm_tol = opt.ipm_tol;
    if (cmpc_create(&cfg_, &h_) != CMPC_OK) throw std::runtime_error("CentroidalMPC: invalid constructor arguments");
  }
  CentroidalMPC(double mass, int num_legs, int predict_horizon, double time_step, const std::vector<double>& weights,
                const std::vector<double>& mu, IPOPT_SOLVER s = IPOPT_SOLVER::MA97, int device = 0,
                const Options& opt = Options())
      : CentroidalMPC(mass, num_legs, predict_horizon, time_step, weights.data(), weights.size(), mu.data(), mu.size(), s, device, opt) {}
#ifdef CMPC_HAVE_EIGEN
  CentroidalMPC(double mass, int num_legs, int predict_horizon, double time_step, const Eigen::VectorXd& weights,
                const Eigen::VectorXd& mu, IPOPT_SOLVER s = IPOPT_SOLVER::MA97, int device = 0, const Options& opt = Options())
      : CentroidalMPC(mass, num_legs, predict_horizon, time_step, weights.data(), (size_t)weights.size(), mu.data(), (size_t)mu.size(), s, device, opt) {}
#endif
  ~CentroidalMPC() { cmpc_destroy(h_); }

  /* reference SetupMPC() (CentroidalMPC.cpp:102): one-off device allocation for up to max_batch instances.  Calling it
   * again with a larger batch re-creates the handle (device buffers are sized once per handle). */
  void SetupMPC(int max_batch = 1) {
    if (max_batch_ > 0) {
      if (max_batch <= max_batch_) return;
      cmpc_destroy(h_);
      h_ = nullptr;
      if (cmpc_create(&cfg_, &h_) != CMPC_OK) throw std::runtime_error("CentroidalMPC: invalid constructor arguments");
    }
    check(cmpc_setup(h_, max_batch, device_));
    max_batch_ = max_batch;
  }
  /* reference NonlinearMPC::UpdateWeights */
  void UpdateWeights(const double* w, size_t n) {
    check(cmpc_update_weights(h_, w, (int)n));
    for (size_t i = 0; i < n && i < (size_t)CMPC_NUM_WEIGHTS; ++i) cfg_.weights[i] = w[i];  // survives a handle re-creation
  }
  void UpdateWeights(const std::vector<double>& w) { UpdateWeights(w.data(), w.size()); }

  size_t state_size() const { return 9 + 3 * (size_t)num_legs_; }
  size_t des_state_size() const { return 9 * ((size_t)horizon_ + 1); }
  size_t des_inputs_size() const { return (size_t)num_legs_ * (4 * (size_t)horizon_ + 3); }
  size_t forces_size() const { return 3 * (size_t)num_legs_ * (size_t)horizon_; }

  /* reference UpdateMPC(state, des_state, des_inputs) (CentroidalMPC.cpp:278) */
  std::vector<double> UpdateMPC(const double* state, const double* des_state, const double* des_inputs) {
    BatchResult r = UpdateMPCBatch(1, state, des_state, des_inputs);
    if (r.status[0] == CMPC_STATUS_INVALID_TABLE) throw std::runtime_error("mpc table invalid");
    if (r.status[0] == CMPC_STATUS_NUMERICAL) throw std::runtime_error("CentroidalMPC: numerical failure / non-finite input");
    current_time_ += cfg_.dt;  // reference :368
    return r.forces;
  }
  std::vector<double> UpdateMPC(const std::vector<double>& state, const std::vector<double>& des_state,
                                const std::vector<double>& des_inputs) {
    if (state.size() != state_size() || des_state.size() != des_state_size() || des_inputs.size() != des_inputs_size())
      throw std::runtime_error("CentroidalMPC::UpdateMPC: argument sizes");
    return UpdateMPC(state.data(), des_state.data(), des_inputs.data());
  }
#ifdef CMPC_HAVE_EIGEN
  Eigen::VectorXd UpdateMPC(const Eigen::VectorXd& state, const Eigen::VectorXd& des_state, const Eigen::VectorXd& des_inputs) {
    std::vector<double> f = UpdateMPC(state.data(), des_state.data(), des_inputs.data());
    return Eigen::Map<Eigen::VectorXd>(f.data(), (Eigen::Index)f.size());
  }
#endif

  /* the foot_pos half of the reference's controller outputs (CentroidalMPC.cpp:269): optimal
   * foot positions, per leg 3 x (N+1) column-major, from the decoupled foot sub-problem */
  std::vector<double> FootPlan(const double* state, const double* des_inputs) {
    if (max_batch_ == 0) SetupMPC(1);
    std::vector<double> fp(3 * (size_t)num_legs_ * ((size_t)horizon_ + 1));
    check(cmpc_foot_plan_batch(h_, 1, state, des_inputs, fp.data()));
    return fp;
  }

  /* B instances, instance-major host buffers (layouts in include/cmpc.h) */
  BatchResult UpdateMPCBatch(int B, const double* states, const double* des_states, const double* des_inputs) {
    if (B > max_batch_) SetupMPC(B);   // first call, or a batch larger than any before: (re)allocate
    BatchResult r;
    r.forces.resize((size_t)B * forces_size());
    r.status.resize(B); r.iters.resize(B); r.kkt.resize(B);
    if (sqp_iters_ > 0) {
      r.stats = cmpc_stats();
      check(cmpc_solve_batch_sqp(h_, B, sqp_iters_, states, des_states, des_inputs, r.forces.data(), r.status.data(), nullptr));
    } else {
      check(cmpc_solve_batch(h_, B, states, des_states, des_inputs, r.forces.data(), r.status.data(), r.iters.data(),
                             r.kkt.data(), nullptr, nullptr, &r.stats));
    }
    return r;
  }

  cmpc_handle* handle() { return h_; }
  double current_time() const { return current_time_; }

 private:
  void check(int rc) {
    if (rc != CMPC_OK) throw std::runtime_error(std::string("CentroidalMPC: ") + cmpc_last_error(h_));
  }
  cmpc_config cfg_{};
  cmpc_handle* h_ = nullptr;
  int num_legs_, horizon_, device_, sqp_iters_ = 0, max_batch_ = 0;
  double current_time_ = 0.0;
};
