// cmpc_api.cu -- C ABI (include/cmpc.h) over the sm_100a kernels in cmpc_device.cuh.
// Host side of the drop-in boundary: one handle = one CUDA device + one stream; all
// device buffers are allocated in cmpc_setup, none in the solve calls.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <new>
#include <string>

#include "cmpc_device.cuh"

using namespace cmpc;

struct cmpc_handle {
  cmpc_config cfg;
  DevConfig dev;
  int device = -1;
  int max_batch = 0;
  int max_batch_plan = 0;  // max_batch while cmpc_setup is planning (max_batch itself is set when setup has succeeded)
  int num_sms = 0;
  bool ready = false;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  bool zero_copy = true;  // CMPC_NO_ZEROCOPY=1 forces the staged-copy path
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  // progressive host-buffer path of cmpc_solve_batch: chunked copy-in on its own stream
  static constexpr int kMaxChunks = 8;
  cudaStream_t s_in = nullptr, s_out = nullptr;
  cudaStream_t s_aux = nullptr;      // launch_solve: the stage-wise interior point next to the condensed kernels (fork / join by events)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool overlap = true;               // CMPC_OVERLAP=0: one stream, the routes one after the other
  bool overlap_trim = true;          // CMPC_OVERLAP_TRIM=0: full grid for the interior-point phase kernel next to the stage-wise kernel
  int overlap_pct = 100;             // CMPC_OVERLAP_PCT: CTAs of the stage-wise kernel in percent of what its expected instances fill
  cudaEvent_t ev_span[4] = {}, ev_in[kMaxChunks] = {}, ev_k[kMaxChunks] = {};
  // device buffers
  double *d_state = nullptr, *d_ds = nullptr, *d_di = nullptr, *d_forces = nullptr, *d_kkt = nullptr,
         *d_lam = nullptr, *d_flog = nullptr, *d_hip = nullptr, *d_sqp = nullptr;
  size_t flog_cap = 0;  // doubles of d_flog (cmpc_rollout's force log: grown on demand, kept across calls)
  int32_t *d_status = nullptr, *d_iters = nullptr, *d_iters_sum = nullptr, *d_status_or = nullptr;
  uint16_t* d_active = nullptr;
  void* d_stats = nullptr;
  // launch plan: one entry per size class (instances bucketed by free-block count)
  struct ClassPlan {
    bool used = false;
    int W = 1, groups = 1, grid = 0, nbmax = 0, n4max = 0, m_in_smem = 1;
    size_t smem_bytes = 0, scratch_per_group = 0;
    double* d_scratch = nullptr;
    double* d_xstate = nullptr;   // phase-split interior point: one iterate slot per instance (u | zl | zu | it)
    int xstride = 0;
    // presolve kernel of the class (cmpc_presolve_kernel): more groups per CTA, H-only scratch
    bool pre_used = false;
    int pre_groups = 0;
    size_t pre_smem_bytes = 0, pre_scratch_per_group = 0;
    double* d_pre_scratch = nullptr;
  } cls[kNumClasses], exp_plan;
  int4 bounds = {0, 0, 0, 0};
  // Riccati presolve kernel (one plan for every size class: its working set does not depend on n)
  bool ric_used = false;
  int ric_groups = 0;
  size_t ric_smem_bytes = 0;
  double* d_ric_scratch = nullptr;
  // stage-wise interior-point kernel (cmpc_ripm.cu): takes the deferred instances of the classes ipm_kind() gives it
  bool rip_used = false;
  int rip_groups = 0, rip_mode = 0;  // rip_mode (CMPC_IPM_BACKEND): 0 follow qp_backend, 1 dense everywhere, 2 stage-wise everywhere
  size_t rip_smem_bytes = 0, rip_slab = 0;
  double* d_rip_scratch = nullptr;
  int32_t* d_ready = nullptr;        // chunks of inputs landed (written by the copy stream, polled by the router kernel)
  int32_t* h_ready_vals = nullptr;   // pinned {1, 2, ...}: the values the copy stream writes into d_ready
  int32_t* h_error_dev = nullptr;    // device alias of h_error
  int32_t* h_error = nullptr;        // pinned + mapped: set by a kernel whose wait on d_ready timed out
  int e2e_mode = 0;                  // CMPC_E2E_MODE, see cmpc_solve_batch: 0 auto, 1 zero-copy, 2 staged, 3 progressive, 4 pipelined, 5 full duplex
  // auto mode, big pinned batches: the first calls time the zero-copy and the pipelined route (which one
  // wins depends on the box: DMA 35-56 GB/s vs ~26 GB/s SM-issued reads), then the faster one is kept
  int e2e_chunk = 0;                 // CMPC_E2E_CHUNK: instances per copy chunk (0: default of the route)
  bool debug_tune = false;           // CMPC_DEBUG_TUNE: print the route timings of the tuning calls
  bool debug_plan = false;           // CMPC_DEBUG_PLAN: print the launch plan of every call
  bool debug_timeline = false;       // CMPC_DEBUG_TIMELINE: print where the events of a host-buffer call fell (ms from its start)
  int tune_calls = 0, tune_batch = 0;
  float tune_best[3] = {1e30f, 1e30f, 1e30f};  // [0] zero-copy, [1] pipelined, [2] full duplex: best span in ms
  char route[96] = "none";           // what the last cmpc_solve_batch call did (cmpc_last_route)
  // the counts / work block is double-buffered: a call uses copy counts_cur; the router presolve kernel of that call zeroes the
  // other copy for the next call, so that a call does not start with a memset node (2 us of a 97 us headline step)
  int counts_cur = 0;
  bool counts_zero[2] = {false, false};  // copy i is (or will be, in stream order) all zero
  int32_t *d_counts = nullptr, *d_perm = nullptr;  // counts / work of: class lists, deferred lists, polish lists, fall-back lists (4 each); perm [4][4][B]
  int32_t* h_hint = nullptr;      // pinned + mapped [4]: instances the last call's interior-point launch of each class found (0 = skip the split kernels)
  int32_t* h_hint_dev = nullptr;
  int32_t* d_hint_shadow = nullptr;
  bool split = true;              // CMPC_SPLIT=0: never use the phase-split kernels
  int split_maxw = 2;             // CMPC_SPLIT_MAXW: widest warp group that uses them (measured: a gain for one warp per instance, a loss for four)
  int rip_minclass = 1;           // CMPC_RIPM_MINCLASS: smallest size class the stage-wise interior point takes in automatic mode
  int dense_lock = 4;             // CMPC_DENSE_LOCK: bit p = CTA barrier in front of every polish pass of the phase-p condensed kernel (bit 3: and
                                  // in front of the reduced factorisation).  Measured on the constrained horizon-10 workload: polish kernel
                                  // (bit 2) +4.5 %, the fused kernel (bit 0) -2 % on ipm_only, second barrier +0.7 %
  int rip_phase_lock = -1;        // CMPC_RIPM_LOCK (-1: automatic, on from horizon 20): CTA barrier in front of every factor sweep of the stage-wise interior point
  int pdl_trigger = -1;           // CMPC_PDL_TRIGGER: which kernels let their dependents be scheduled from their first instruction on
                                  // (0 none: as their CTAs retire; 1 the presolve kernels; 2 all; -1 automatic).  Measured: 2 is best on the
                                  // headline (the three empty kernels are resident and gone by the time the presolve ends: +1.5 % over 0),
                                  // 1 and 2 cost 4 % at horizon 30 where the stage-wise presolve does the work (config 3): automatic = 2
                                  // below horizon 20, else 0
  bool pdl = true;                // CMPC_PDL=0: plain stream order between the kernels of a call
  std::string err;
};

namespace {

// ------------------------------------------------------------------ classification
// Number of free blocks -> size class -> permutation slot.  One 1024-thread block handles 32
// instances: warp w counts the contact flags of instance 32*blockIdx + w (coalesced reads, all
// misses in flight at once; the flags may sit in mapped host memory), then warp 0 assigns the
// slots, bumping each class counter once per block via __match_any_sync -- one atomic per
// instance on the same address would serialise in L2.
// bounds = largest nb of classes 0..2 (ascending); class 3 takes the rest.
__global__ void __launch_bounds__(1024) classify_kernel(const DevConfig cfg, int B, const double* des_inputs, int4 bounds,
                                                        int32_t* counts, int32_t* perm) {
  __shared__ int s_nb[32];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * 32 + w;
  const int N = cfg.N, L = cfg.L;
  if (b < B) {
    const double* di = des_inputs + (size_t)b * L * (4 * N + 3);
    int nb = 0;
    for (int e = lane; e < L * N; e += 32) {
      const int i = e / N, j = e - i * N;
      nb += __ldg(di + i * (4 * N + 3) + j) > 0.0 ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nb += __shfl_xor_sync(0xffffffffu, nb, o);
    if (lane == 0) s_nb[w] = nb;
  }
  __syncthreads();
  if (w != 0) return;
  const int bb = blockIdx.x * 32 + lane;
  const bool valid = bb < B;
  const int mine = valid ? s_nb[lane] : 0;
  int c = 3;
  if (mine <= bounds.x) c = 0;
  else if (mine <= bounds.y) c = 1;
  else if (mine <= bounds.z) c = 2;
  if (!valid) c = 4;  // lanes past the end form their own group and do nothing
  const unsigned peers = __match_any_sync(0xffffffffu, c);
  const int leader = __ffs(peers) - 1;
  const int rank = __popc(peers & ((1u << lane) - 1u));
  int base = 0;
  if (lane == leader && valid) base = atomicAdd(&counts[c], __popc(peers));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (valid) perm[(size_t)c * B + base + rank] = bb;
}


constexpr size_t kMaxSmem = 232448;  // 227 KB opt-in per CTA on sm_100

struct DevStats {
  unsigned long long iters_sum;
  int max_iters, n_ok, n_ok_ipm, n_max_iter, n_invalid, n_numerical;
  unsigned long long max_kkt_bits;
};

__global__ void stats_kernel(const int32_t* status, const int32_t* iters, const double* kkt, int B, DevStats* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const int st = status[i];
  if (iters) { atomicAdd(&out->iters_sum, (unsigned long long)iters[i]); atomicMax(&out->max_iters, iters[i]); }
  if (st == CMPC_STATUS_OK) atomicAdd(&out->n_ok, 1);
  else if (st == CMPC_STATUS_OK_IPM) atomicAdd(&out->n_ok_ipm, 1);
  else if (st == CMPC_STATUS_MAX_ITER) atomicAdd(&out->n_max_iter, 1);
  else if (st == CMPC_STATUS_INVALID_TABLE) atomicAdd(&out->n_invalid, 1);
  else atomicAdd(&out->n_numerical, 1);
  if (kkt && st <= CMPC_STATUS_OK_IPM) {
    double k = kkt[i];
    if (k >= 0.0) atomicMax(&out->max_kkt_bits, (unsigned long long)__double_as_longlong(k));
  }
}

// Closed loop (BASELINE config 5).  rollout_init captures the hip offsets foot_i - com at t = 0.
__global__ void rollout_init_kernel(const DevConfig cfg, int B, const double* state, double* hip) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int L = cfg.L, ns = 9 + 3 * L;
  const double* x = state + (size_t)b * ns;
  for (int i = 0; i < L; ++i)
    for (int q = 0; q < 3; ++q) hip[((size_t)b * L + i) * 3 + q] = x[9 + 3 * i + q] - (q < 2 ? x[q] : 0.0);
}

// One tick: (1) plant = the reference's nonlinear Euler step (CentroidalMPC.cpp:85-92) with the
// TRUE lever arm foot - com and the first-step forces; (2) contact table rotated by one step
// (period N); (3) a leg that is in swing at the new step 0 has its foot carried under its hip
// (com_xy + hip offset), a stance foot stays where it is; (4) the references are regenerated
// from the new state: des_com_pos_k = (c_xy + k dt vd_xy, z_d) with vd, z_d, L_d held from the
// inputs; des_foot_pos_i[:,k] = current foot while leg i stays in stance from step 0 through
// min(k, N-1), else the hip point of the reference at node k.
__global__ void advance_kernel(const DevConfig cfg, int B, double* state, double* des_state, double* des_inputs,
                               const double* hip, const double* forces, const int32_t* status, double* flog,
                               int32_t* iters_sum, const int32_t* iters, int32_t* status_or) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int N = cfg.N, L = cfg.L;
  const int ns = 9 + 3 * L, nds = 9 * (N + 1), ndi = L * (4 * N + 3), nf = 3 * L * N;
  double* x = state + (size_t)b * ns;
  double* ds = des_state + (size_t)b * nds;
  double* di = des_inputs + (size_t)b * ndi;
  const double* F = forces + (size_t)b * nf;
  double acc[3] = {0.0, 0.0, -kGrav}, ld[3] = {0.0, 0.0, 0.0};
  for (int i = 0; i < L; ++i) {
    const double ce = di[i * (4 * N + 3)];
    const double* f = F + (size_t)i * 3 * N;  // step 0 of leg i
    if (flog) for (int q = 0; q < 3; ++q) flog[(size_t)b * 3 * L + 3 * i + q] = f[q];
    const double r0 = x[9 + 3 * i] - x[0], r1 = x[9 + 3 * i + 1] - x[1], r2 = x[9 + 3 * i + 2] - x[2];
    for (int q = 0; q < 3; ++q) acc[q] += ce / cfg.mass * f[q];
    ld[0] += ce * (r1 * f[2] - r2 * f[1]);
    ld[1] += ce * (r2 * f[0] - r0 * f[2]);
    ld[2] += ce * (r0 * f[1] - r1 * f[0]);
  }
  double xn[9];
  for (int q = 0; q < 3; ++q) {
    xn[q] = x[q] + x[3 + q] * cfg.dt;
    xn[3 + q] = x[3 + q] + acc[q] * cfg.dt;
    xn[6 + q] = x[6 + q] + ld[q] * cfg.dt;
  }
  for (int q = 0; q < 9; ++q) x[q] = xn[q];
  // reference, re-anchored at the new state
  const double vd0 = ds[3 * (N + 1)], vd1 = ds[3 * (N + 1) + 1], vd2 = ds[3 * (N + 1) + 2];
  const double zd = ds[2];
  const double ad0 = ds[6 * (N + 1)], ad1 = ds[6 * (N + 1) + 1], ad2 = ds[6 * (N + 1) + 2];
  for (int k = 0; k <= N; ++k) {
    ds[3 * k] = x[0] + k * cfg.dt * vd0; ds[3 * k + 1] = x[1] + k * cfg.dt * vd1; ds[3 * k + 2] = zd;
    ds[3 * (N + 1) + 3 * k] = vd0; ds[3 * (N + 1) + 3 * k + 1] = vd1; ds[3 * (N + 1) + 3 * k + 2] = vd2;
    ds[6 * (N + 1) + 3 * k] = ad0; ds[6 * (N + 1) + 3 * k + 1] = ad1; ds[6 * (N + 1) + 3 * k + 2] = ad2;
  }
  for (int i = 0; i < L; ++i) {
    double* blk = di + i * (4 * N + 3);
    const double c0 = blk[0];
    for (int j = 0; j + 1 < N; ++j) blk[j] = blk[j + 1];
    blk[N - 1] = c0;
    const double* hp = hip + ((size_t)b * L + i) * 3;
    if (!(blk[0] > 0.0)) { x[9 + 3 * i] = x[0] + hp[0]; x[9 + 3 * i + 1] = x[1] + hp[1]; x[9 + 3 * i + 2] = hp[2]; }
    double* fp = blk + N;
    bool planted = true;  // in stance continuously since step 0
    for (int k = 0; k <= N; ++k) {
      const int j = k < N ? k : N - 1;
      planted = planted && blk[j] > 0.0;
      if (planted) { for (int q = 0; q < 3; ++q) fp[3 * k + q] = x[9 + 3 * i + q]; }
      else { fp[3 * k] = ds[3 * k] + hp[0]; fp[3 * k + 1] = ds[3 * k + 1] + hp[1]; fp[3 * k + 2] = hp[2]; }
    }
  }
  if (iters_sum && iters) iters_sum[b] += iters[b];
  if (status_or) status_or[b] |= (1 << status[b]);
}

// Foot plan (decoupled half of the reference NLP). One thread per (instance, leg, axis).
__global__ void foot_plan_kernel(const DevConfig cfg, int B, const double* state, const double* des_inputs, double* foot_pos) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int N = cfg.N, L = cfg.L;
  if (idx >= B * L * 3) return;
  const int b = idx / (3 * L), i = (idx / 3) % L, q = idx % 3;
  const double* blk = des_inputs + (size_t)b * L * (4 * N + 3) + i * (4 * N + 3);
  const double* pd = blk + N;  // des_foot_pos, node k at pd[3k + q]
  double* out = foot_pos + ((size_t)b * L + i) * 3 * (N + 1);
  const double lb = q == 2 ? -0.1 : -0.2, ub = q == 2 ? 0.1 : 0.2;  // CentroidalMPC.cpp:30-31
  int a = 0;
  while (a <= N) {
    int e = a;  // group of nodes a..e joined by locked intervals (1 - contact == 0, :94)
    while (e < N && 1.0 - blk[e] == 0.0) ++e;
    double v;
    if (a == 0) {
      v = state[(size_t)b * (9 + 3 * L) + 9 + 3 * i + q];  // initial value constraint, :166
    } else {
      double sum = 0.0, lo = -INFINITY, hi = INFINITY;
      for (int k = a; k <= e; ++k) { const double d = pd[3 * k + q]; sum += d; lo = fmax(lo, d + lb); hi = fmin(hi, d + ub); }
      v = fmin(fmax(sum / (double)(e - a + 1), lo), hi);
    }
    for (int k = a; k <= e; ++k) out[3 * k + q] = v;
    a = e + 1;
  }
}

// Re-linearisation. One thread per instance: roll the forces through the QP's linear
// model (arms frozen at di_lin) and through the reference's nonlinear Euler plant (arms
// p_ij - c_j with p = the ORIGINAL desired foot positions di_orig), report the largest state
// difference, and (update != 0) move the linearisation point of the next solve to the
// nonlinear centre-of-mass path by shifting the desired foot positions of di_lin:
//   des_foot_lin_ij = p_ij + (des_com_pos_j - c_j)   =>   arm = des_foot_lin - des_com_pos = p - c.
__global__ void relinearize_kernel(const DevConfig cfg, int B, const double* state, const double* des_state,
                                   const double* di_orig, double* di_lin, const double* forces, double* defect,
                                   int defect_stride, int update) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int N = cfg.N, L = cfg.L;
  const int ns = 9 + 3 * L, nds = 9 * (N + 1), ndi = L * (4 * N + 3), nf = 3 * L * N;
  const double* x0 = state + (size_t)b * ns;
  const double* dpos = des_state + (size_t)b * nds;
  const double* d0 = di_orig + (size_t)b * ndi;
  double* dl = di_lin + (size_t)b * ndi;
  const double* F = forces + (size_t)b * nf;
  const double dt = cfg.dt, m = cfg.mass;
  const double zeta = cfg.zoh ? 0.5 : 0.0;
  double cn[3], vn[3], ln[3], cl[3], vl[3], ll[3];
  for (int q = 0; q < 3; ++q) { cn[q] = cl[q] = x0[q]; vn[q] = vl[q] = x0[3 + q]; ln[q] = ll[q] = x0[6 + q]; }
  double worst = 0.0;
  for (int j = 0; j < N; ++j) {
    // node j: shift the linearisation point of the next solve (node N is never used as an arm)
    double accn[3] = {0.0, 0.0, -kGrav}, ldn[3] = {0.0, 0.0, 0.0}, ldl[3] = {0.0, 0.0, 0.0};
    for (int i = 0; i < L; ++i) {
      const double ce = d0[i * (4 * N + 3) + j] > 0.0 ? d0[i * (4 * N + 3) + j] : 0.0;
      const double* f = F + ((size_t)i * N + j) * 3;
      double p[3], rl[3], rn[3];
      for (int q = 0; q < 3; ++q) {
        p[q] = d0[i * (4 * N + 3) + N + 3 * j + q];
        rl[q] = dl[i * (4 * N + 3) + N + 3 * j + q] - dpos[3 * j + q];  // the arm the QP used
        rn[q] = p[q] - cn[q];                                            // the reference's arm (:86)
      }
      for (int q = 0; q < 3; ++q) accn[q] += ce / m * f[q];
      ldn[0] += ce * (rn[1] * f[2] - rn[2] * f[1]); ldn[1] += ce * (rn[2] * f[0] - rn[0] * f[2]); ldn[2] += ce * (rn[0] * f[1] - rn[1] * f[0]);
      ldl[0] += ce * (rl[1] * f[2] - rl[2] * f[1]); ldl[1] += ce * (rl[2] * f[0] - rl[0] * f[2]); ldl[2] += ce * (rl[0] * f[1] - rl[1] * f[0]);
      if (update) for (int q = 0; q < 3; ++q) dl[i * (4 * N + 3) + N + 3 * j + q] = p[q] + (dpos[3 * j + q] - cn[q]);
    }
    for (int q = 0; q < 3; ++q) {
      // linear model of the QP (Euler, or ZOH with its dt^2/2 position terms)
      cl[q] += dt * vl[q] + zeta * dt * dt * accn[q];
      vl[q] += dt * accn[q];
      ll[q] += dt * ldl[q];
      // reference plant: explicit Euler (CentroidalMPC.cpp:90-92)
      cn[q] += dt * vn[q];
      vn[q] += dt * accn[q];
      ln[q] += dt * ldn[q];
      worst = fmax(worst, fmax(fabs(cn[q] - cl[q]), fmax(fabs(vn[q] - vl[q]), fabs(ln[q] - ll[q]))));
    }
  }
  if (defect) defect[(size_t)b * defect_stride] = worst;
}

// Gait -> contact table. One thread per (instance, step).
constexpr int kMaxGaits = 16;
struct GaitTable {
  int num_gaits;
  cmpc_gait g[kMaxGaits];
};
__global__ void gait_kernel(const DevConfig cfg, int B, const GaitTable tab, const int32_t* gait_id, const double* t0,
                            double* des_inputs) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int N = cfg.N, L = cfg.L;
  if (idx >= B * N) return;
  const int b = idx / N, j = idx - b * N;
  int gid = gait_id[b];
  if (gid < 0 || gid >= tab.num_gaits) gid = 0;
  const cmpc_gait& g = tab.g[gid];
  const double start = g.switching_times[0], duration = g.switching_times[g.num_modes] - start;
  // no fused multiply-add here: a step that lands exactly on a switching time must round like the host code
  const double t = __dadd_rn(t0[b], __dmul_rn((double)j, cfg.dt));
  double phase = fmod(__ddiv_rn(t, duration), 1.0);  // wrapPhase, Gait.cpp:63-69
  if (phase < 0.0) phase += 1.0;
  int k = 0;  // upper_bound over eventPhases = (switching_times[1..n-1] - start) / duration
  while (k < g.num_modes - 1 && !(phase < __ddiv_rn(__dsub_rn(g.switching_times[k + 1], start), duration))) ++k;
  const int mode = g.modes[k];
  const int bit[4] = {8, 4, 1, 2};  // {lf, rf, rh, lh} <- {LF, RF, RH, LH}
  for (int i = 0; i < L; ++i) des_inputs[(size_t)b * L * (4 * N + 3) + i * (4 * N + 3) + j] = (mode & bit[i]) ? 1.0 : 0.0;
}

// Gait SWITCH (the stance-insertion rule): the contact flags of a mode schedule in which template `from`,
// tiled from t_tile (GaitSchedule::tileModeSequenceTemplate, GaitSchedule.cpp:107-137: event times by repeated addition
// of the template's intervals, whole cycles), is replaced at t_switch by template `to` the way
// GaitSchedule::insertModeSequenceTemplate does it (GaitSchedule.cpp:47-72): events at or after t_switch are erased, the
// mode active there runs on until t_switch, an intermediate STANCE phase of stance_time follows unless that mode already
// is STANCE, then `to` is tiled from t_switch (+ stance_time).  Mode at time t = modeSequence[lower_bound(eventTimes, t)]
// (ModeSchedule::modeAtTime of the un-vendored ocs2_core: a step that lands exactly on an event keeps the earlier mode).
// One thread per instance walks the event stream once; every time is formed by the same additions as the host code
// (no fused multiply-add), so the flags are bit-exact against the mirror.
__global__ void gait_switch_kernel(const DevConfig cfg, int B, const GaitTable tab, const int32_t* gait_from, const int32_t* gait_to,
                                   const double* t_tile, const double* t_switch, double stance_time, const double* t0, double* des_inputs) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int N = cfg.N, L = cfg.L;
  constexpr int kStance = 15;
  int ga = gait_from[b], gb = gait_to[b];
  if (ga < 0 || ga >= tab.num_gaits) ga = 0;
  if (gb < 0 || gb >= tab.num_gaits) gb = 0;
  const double ts = t_switch[b];
  const double t_end = __dadd_rn(t0[b], __dmul_rn((double)N, cfg.dt));
  const bool sw = isfinite(ts);
  // Event stream: `mode` is active up to and including the next event `te`.  stage 0: the old tiling (its events at or
  // after t_switch are never taken), 1: the inserted stance event at t_switch is pending, 2: the new tiling.
  int stage = 0, gi = ga, idx = 0, mode = kStance;
  double te = t_tile[b], fin = sw ? ts : t_end;
  const int bit[4] = {8, 4, 1, 2};  // {lf, rf, rh, lh} <- {LF, RF, RH, LH}
  for (int j = 0; j < N; ++j) {
    const double t = __dadd_rn(t0[b], __dmul_rn((double)j, cfg.dt));
    for (;;) {
      if (stage == 0 && sw && !(te < ts)) {  // insertion (GaitSchedule.cpp:52-71): the current mode runs on until t_switch
        if (mode != kStance && stance_time > 0.0) { stage = 1; te = ts; }
        else { stage = 2; gi = gb; idx = 0; te = ts; fin = t_end; }
        continue;
      }
      if (!(te < t)) break;                  // lower_bound: an event equal to t has not happened yet
      if (stage == 1) {                      // intermediate stance, the new template is tiled from t_switch + stance_time
        mode = kStance; stage = 2; gi = gb; idx = 0; te = __dadd_rn(ts, stance_time); fin = t_end;
      } else {                               // tiling (GaitSchedule.cpp:121-136): whole cycles while the last event is before the final time
        const cmpc_gait& g = tab.g[gi];
        if (idx == 0 && !(te < fin)) { mode = kStance; te = INFINITY; }   // default final phase
        else {
          mode = g.modes[idx];
          te = __dadd_rn(te, __dsub_rn(g.switching_times[idx + 1], g.switching_times[idx]));
          if (++idx == g.num_modes) idx = 0;
        }
      }
    }
    for (int i = 0; i < L; ++i) des_inputs[(size_t)b * L * (4 * N + 3) + i * (4 * N + 3) + j] = (mode & bit[i]) ? 1.0 : 0.0;
  }
}

// FP64 throughput probe: 8 independent DFMA chains per thread.
__global__ void fp64_peak_kernel(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1e-9, a2 = a0 + 2e-9, a3 = a0 + 3e-9;
  double a4 = a0 + 4e-9, a5 = a0 + 5e-9, a6 = a0 + 6e-9, a7 = a0 + 7e-9;
  const double m = 0.999999, c = 1e-7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
      a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// device alias of a pinned (page-locked) host pointer, or nullptr for pageable / device memory
void* mapped_device_pointer(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  if (at.type == cudaMemoryTypeHost && at.devicePointer) return at.devicePointer;
  return nullptr;
}

int fail(cmpc_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  return code;
}
// Makes the handle's device current for the duration of an entry point and restores the caller's device afterwards
// (a host thread may own handles on several GPUs or do its own CUDA work between calls).
struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    const bool had = cudaGetDevice(&prev) == cudaSuccess;
    ok = cudaSetDevice(dev) == cudaSuccess;
    if (!had) prev = -1;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};
// Temporary device buffer of the diagnostic / non-per-tick entry points: freed on every return path.
struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 8); }
  template <class T> T* as() const { return static_cast<T*>(p); }
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
};
#define SET_DEVICE(h)                                                                          \
  DeviceGuard device_guard_((h)->device);                                                      \
  if (!device_guard_.ok) return fail(h, CMPC_ERR_CUDA, "cudaSetDevice failed")

#define CUDA_TRY(h, expr)                                                                      \
  do {                                                                                         \
    cudaError_t e_ = (expr);                                                                   \
    if (e_ != cudaSuccess)                                                                     \
      return fail(h, CMPC_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));       \
  } while (0)

bool valid_config(const cmpc_config* c) {
  // the reference asserts mass > 0 && num_legs > 0 && horizon > 0 && mu.size() == num_legs
  if (!(c->mass > 0) || !std::isfinite(c->mass)) return false;
  if (c->num_legs < 1 || c->num_legs > CMPC_MAX_LEGS) return false;
  if (c->horizon < 1 || c->horizon > CMPC_MAX_HORIZON) return false;
  if (!(c->dt > 0) || !std::isfinite(c->dt)) return false;
  for (int i = 0; i < c->num_legs; ++i)
    if (!(c->mu[i] > 0) || !std::isfinite(c->mu[i])) return false;
  for (int i = 0; i < 9 + 9 * c->num_legs; ++i)
    if (!(c->weights[i] >= 0) || !std::isfinite(c->weights[i])) return false;
  // strict convexity of the condensed QP needs positive force-tracking weights (K > 0)
  for (int i = 0; i < 3 * c->num_legs; ++i)
    if (!(c->weights[9 + 3 * c->num_legs + i] > 0)) return false;
  if (c->disc_mode != 0 && c->disc_mode != 1) return false;
  if (c->max_iter < 1 || !(c->ipm_tol > 0)) return false;
  if (c->qp_backend < 0 || c->qp_backend > 2) return false;
  return true;
}

void fill_dev(cmpc_handle* h) {
  const cmpc_config& c = h->cfg;
  DevConfig& d = h->dev;
  d.mass = c.mass; d.dt = c.dt; d.L = c.num_legs; d.N = c.horizon; d.zoh = c.disc_mode;
  d.max_iter = c.max_iter; d.tol = c.ipm_tol; d.polish = c.polish;
  for (int i = 0; i < CMPC_MAX_LEGS; ++i) d.mu[i] = c.mu[i];
  for (int i = 0; i < CMPC_NUM_WEIGHTS; ++i) d.w[i] = c.weights[i];
}

template <int MODE>
int launch_class(cmpc_handle* h, const cmpc_handle::ClassPlan& p, SolveArgs a, int phase = 0, int grid = 0) {
  a.scratch = p.d_scratch; a.scratch_per_group = p.scratch_per_group;
  a.xstate = p.d_xstate; a.xstride = p.xstride;
  a.nbmax = p.nbmax; a.n4max = p.n4max; a.m_in_smem = p.m_in_smem; a.groups = p.groups;
  a.plan = make_plan(h->cfg.horizon, h->cfg.num_legs, p.m_in_smem ? p.W : 8, p.nbmax, p.n4max, p.m_in_smem);
  a.phase_lock = (MODE == 0 && ((h->dense_lock >> phase) & 1)) ? 1 + ((h->dense_lock >> 3) & 1) : 0;
  const int W = (MODE == 1 || !p.m_in_smem) ? 8 : p.W;
  const cudaError_t e = launch_solve_kernel(W, MODE, p.m_in_smem != 0, phase, grid > 0 ? grid : p.grid, 32 * W * p.groups, p.smem_bytes, h->stream, h->dev, a);
  if (e != cudaSuccess) return fail(h, CMPC_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e));
  return CMPC_OK;
}

int launch_presolve(cmpc_handle* h, const cmpc_handle::ClassPlan& p, SolveArgs a) {
  a.scratch = p.d_pre_scratch; a.scratch_per_group = p.pre_scratch_per_group;
  a.nbmax = p.nbmax; a.n4max = p.n4max; a.m_in_smem = 1; a.groups = p.pre_groups;
  a.pre = make_pre_plan(h->cfg.horizon, h->cfg.num_legs, p.W, p.nbmax, p.n4max);
  const cudaError_t e = launch_presolve_kernel(p.W, p.grid, 32 * p.W * p.pre_groups, p.pre_smem_bytes, h->stream, h->dev, a);
  if (e != cudaSuccess) return fail(h, CMPC_ERR_CUDA, std::string("presolve launch: ") + cudaGetErrorString(e));
  return CMPC_OK;
}

int launch_riccati(cmpc_handle* h, SolveArgs a) {
  a.scratch = h->d_ric_scratch;
  a.ric = make_ric_plan(h->cfg.horizon, h->cfg.num_legs);
  a.scratch_per_group = (size_t)a.ric.slab;
  a.nbmax = h->cfg.horizon * h->cfg.num_legs; a.n4max = 0; a.m_in_smem = 1; a.groups = h->ric_groups;
  const cudaError_t e = launch_riccati_kernel(h->num_sms, 32 * h->ric_groups, h->ric_smem_bytes, h->stream, h->dev, a);
  if (e != cudaSuccess) return fail(h, CMPC_ERR_CUDA, std::string("riccati launch: ") + cudaGetErrorString(e));
  return CMPC_OK;
}

int launch_ripm(cmpc_handle* h, SolveArgs a, int grid = 0, cudaStream_t stream = nullptr) {
  a.scratch = h->d_rip_scratch;
  a.scratch_per_group = h->rip_slab;
  a.nbmax = h->cfg.horizon * h->cfg.num_legs; a.n4max = 0; a.m_in_smem = 1; a.groups = h->rip_groups;
  // measured: +13 % at horizon 30 (instruction fetch is what the warps of an SM compete for), -2 % at horizon 10
  a.phase_lock = h->rip_phase_lock >= 0 ? h->rip_phase_lock : (h->cfg.horizon >= 20 ? 1 : 0);
  a.hint_out = h->h_hint_dev + kNumClasses; a.hint_shadow = h->d_hint_shadow + kNumClasses;
  const cudaError_t e = launch_ripm_kernel(grid > 0 ? grid : h->num_sms, 32 * h->rip_groups, h->rip_smem_bytes, stream ? stream : h->stream, h->dev, a);
  if (e != cudaSuccess) return fail(h, CMPC_ERR_CUDA, std::string("stage-wise interior-point launch: ") + cudaGetErrorString(e));
  return CMPC_OK;
}

// which interior-point kernel takes the (deferred) instances of size class c: 0 condensed dense (cmpc_solve.cu),
// 1 stage-wise Riccati (cmpc_ripm.cu).  Automatic: dense for up to 42 free leg-steps (n <= 126: the one- and four-warp
// kernels, whose per-instance latency is lower), stage-wise above (the eight-warp classes: everything at horizon 30).
// A warm-started call (closed loop) keeps the dense kernel, whose polish takes the previous active set.
int ipm_kind(const cmpc_handle* h, int c, bool warm) {
  if (!h->rip_used || warm || !h->cfg.polish) return 0;
  if (h->rip_mode == 1) return 0;
  if (h->rip_mode == 2) return 1;
  if (h->cfg.qp_backend == 1) return 0;
  if (h->cfg.qp_backend == 2) return 1;
  return c >= h->rip_minclass ? 1 : 0;
}

// which presolve kernel settles size class c: 0 none, 1 dense (cmpc_presolve.cu), 2 Riccati (cmpc_riccati.cu)
int presolve_kind(const cmpc_handle* h, int c) {
  if (!(h->cfg.presolve && h->cfg.polish)) return 0;
  const bool dense_ok = h->cls[c].pre_used, ric_ok = h->ric_used;
  if (h->cfg.qp_backend == 1) return dense_ok ? 1 : 0;
  if (h->cfg.qp_backend == 2) return ric_ok ? 2 : (dense_ok ? 1 : 0);
  if (c >= 1 && ric_ok) return 2;  // more than 20 free leg-steps: the stage-wise sweep beats the multi-warp dense presolve
  return dense_ok ? 1 : (ric_ok ? 2 : 0);
}

// One batch on the handle's stream, no host round trip.  Returns the number of kernels launched
// (< 0: error).
//  * presolve on (default): the presolve kernel of size class 0 walks all B instances itself -- it
//    forwards those with more free blocks to their class lists (no classify launch), settles the
//    ones whose unconstrained minimiser is feasible with one Cholesky of H, and defers the rest to
//    the interior-point kernel; then the same pair of kernels per larger class that is in use.
//    ready != nullptr: the inputs are still arriving (copy stream bumps *ready per chunk of
//    ready_chunk instances) and the router waits per instance for its chunk.
//  * presolve off: classify kernel, then one interior-point kernel per class.
constexpr int kCounts = 8 * kNumClasses;  // counts and work counters of one call: class, deferred, polish and fall-back lists

int launch_solve(cmpc_handle* h, SolveArgs a, int B, const int32_t* ready = nullptr, int ready_chunk = 0) {
  // Counts / work block of this call.  Under stream capture the captured sequence is replayed at times the handle does not
  // see: it zeroes its copy itself, before and after, and the handle's bookkeeping is left alone.
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(h->stream, &cap) != cudaSuccess) return fail(h, CMPC_ERR_CUDA, "stream capture status");
  const bool capturing = cap != cudaStreamCaptureStatusNone;
  const int cur = h->counts_cur;
  int32_t* const cnt = h->d_counts + cur * kCounts;
  int32_t* const cnt_next = h->d_counts + (cur ^ 1) * kCounts;
  if (capturing || !h->counts_zero[cur]) {
    if (cudaMemsetAsync(cnt, 0, kCounts * sizeof(int32_t), h->stream) != cudaSuccess) return fail(h, CMPC_ERR_CUDA, "memset counts");
  }
  if (!capturing) { h->counts_zero[cur] = false; h->counts_cur = cur ^ 1; }
  int launches = 0;
  const bool presolve = h->cfg.presolve && h->cfg.polish;
  const bool router = presolve && h->cls[0].used && presolve_kind(h, 0) == 1;
  const bool warm = a.warm_active != nullptr;
  // every solver kernel after the first one of the call is launched with programmatic stream serialisation
  // (pdl_prologue, cmpc_device.cuh): a kernel that finds its list empty then costs about a microsecond instead of four
  a.pdl = 0;
  a.pdl_trigger = h->pdl_trigger >= 0 ? h->pdl_trigger : (h->cfg.horizon < 20 ? 2 : 0);
  auto chained = [&](SolveArgs& x) { x.pdl = (h->pdl && launches > 0) ? 1 : 0; };
  if (!router) {
    if (ready) return fail(h, CMPC_ERR_STATE, "progressive inputs need the presolve router");
    classify_kernel<<<(B + 31) / 32, 1024, 0, h->stream>>>(h->dev, B, a.des_inputs, h->bounds, cnt, h->d_perm);
    ++launches;
  }
  a.nlists = 0;
  // per class: the list it starts from (classify / router output) and, after a presolve, the list of what was deferred
  const int32_t* in_perm[kNumClasses]; const int32_t* in_count[kNumClasses]; int32_t* in_work[kNumClasses];
  for (int c = 0; c < kNumClasses; ++c) {
    in_perm[c] = h->d_perm + (size_t)c * B;
    in_count[c] = cnt + c;
    in_work[c] = cnt + kNumClasses + c;
  }
  // ---- presolves: the dense kernel per class (class 0 = the batch's router), then ONE stage-wise launch over the other classes
  SolveArgs ric = a;
  for (int c = 0; c < kNumClasses; ++c) {
    if (!h->cls[c].used) continue;
    const int kind = presolve_kind(h, c);
    if (!kind) continue;
    int32_t* fperm = h->d_perm + (size_t)(kNumClasses + c) * h->max_batch;
    int32_t* fcount = cnt + 2 * kNumClasses + c;
    if (kind == 1) {
      SolveArgs p = a;
      p.perm = in_perm[c]; p.count = in_count[c]; p.work = in_work[c];
      p.fail_perm = fperm; p.fail_count = fcount;
      p.route = 0; p.ready = nullptr;
      if (router && c == 0) {
        p.perm = nullptr; p.count = nullptr; p.count_imm = B;
        p.route = 1; p.route_b1 = h->bounds.y; p.route_b2 = h->bounds.z;
        p.route_perm = h->d_perm; p.route_counts = cnt; p.route_stride = B;
        p.ready = ready; p.ready_chunk = ready_chunk; p.error_flag = h->h_error_dev;
        if (launches == 0 && !capturing) { p.zero_next = cnt_next; p.zero_n = kCounts; }  // (first kernel of the call, not PDL-chained)
      }
      chained(p);
      int rc = launch_presolve(h, h->cls[c], p);
      if (rc) return rc;
      if (p.zero_next) h->counts_zero[cur ^ 1] = true;
      ++launches;
    } else {  // stage-wise presolve: collected, largest class first
      for (int q = ric.nlists; q > 0; --q) {
        ric.lperm[q] = ric.lperm[q - 1]; ric.lcount[q] = ric.lcount[q - 1]; ric.lwork[q] = ric.lwork[q - 1];
        ric.lfail_perm[q] = ric.lfail_perm[q - 1]; ric.lfail_count[q] = ric.lfail_count[q - 1];
      }
      ric.lperm[0] = in_perm[c]; ric.lcount[0] = in_count[c]; ric.lwork[0] = in_work[c];
      ric.lfail_perm[0] = fperm; ric.lfail_count[0] = fcount;
      ++ric.nlists;
    }
    in_perm[c] = fperm; in_count[c] = fcount; in_work[c] = cnt + 3 * kNumClasses + c;
  }
  if (ric.nlists > 0) {
    ric.route = 0; ric.ready = nullptr;
    chained(ric);
    int rc = launch_riccati(h, ric);
    if (rc) return rc;
    ++launches;
  }
  // ---- interior point + polish for what is left: the condensed kernel per class, ONE stage-wise launch for its classes
  SolveArgs rip = a;
  rip.route = 0; rip.ready = nullptr; rip.fail_perm = nullptr; rip.fail_count = nullptr;
  int dense_expected = 0;
  for (int c = 0; c < kNumClasses; ++c) {
    if (!h->cls[c].used) continue;
    if (ipm_kind(h, c, warm)) {
      for (int q = rip.nlists; q > 0; --q) { rip.lperm[q] = rip.lperm[q - 1]; rip.lcount[q] = rip.lcount[q - 1]; rip.lwork[q] = rip.lwork[q - 1]; }
      rip.lperm[0] = in_perm[c]; rip.lcount[0] = in_count[c]; rip.lwork[0] = in_work[c];
      ++rip.nlists;
    } else {
      dense_expected += h->h_hint[c];
    }
  }
  // Both routes busy (by the previous call's counts): the stage-wise kernel is latency-bound -- a few hundred stand
  // instances are one wave of 1.9 ms on half of the SMs -- so it runs on its own stream on just the CTAs it can fill,
  // next to the condensed kernels (whose CTAs are persistent: the ones that find their SM taken start late and join the
  // work loop).  A wrong guess only costs time.
  int rip_grid = 0;
  if (h->overlap && rip.nlists > 0 && dense_expected > 0 && h->h_hint[kNumClasses] > 0) {
    rip_grid = ((h->h_hint[kNumClasses] + h->rip_groups - 1) / h->rip_groups * h->overlap_pct + 99) / 100;
    if (rip_grid > (h->num_sms * 5) / 8) rip_grid = 0;  // it would take most of the device anyway
  }
  if (h->debug_plan) fprintf(stderr, "[cmpc] plan: hints dense %d stage-wise %d -> stage-wise grid %d%s\n", dense_expected, h->h_hint[kNumClasses], rip_grid, rip_grid > 0 ? " (own stream)" : "");
  if (rip_grid > 0) {
    if (cudaEventRecord(h->ev_fork, h->stream) != cudaSuccess || cudaStreamWaitEvent(h->s_aux, h->ev_fork, 0) != cudaSuccess)
      return fail(h, CMPC_ERR_CUDA, "fork to the auxiliary stream");
    rip.pdl = 0;
    int rc = launch_ripm(h, rip, rip_grid, h->s_aux);
    if (rc) return rc;
    ++launches;
    if (cudaEventRecord(h->ev_join, h->s_aux) != cudaSuccess) return fail(h, CMPC_ERR_CUDA, "join event");
  }
  bool first_dense = true;
  for (int c = 0; c < kNumClasses; ++c) {
    if (!h->cls[c].used || ipm_kind(h, c, warm)) continue;
    {
      // Condensed route.  Behind the presolve (what is left then is the constrained part of the batch: iteration counts
      // and polish passes differ from instance to instance, so the warps of an SM drift through different phases of
      // the fused kernel and thrash the instruction cache), and when the previous call found instances on this class's
      // list, the interior point and the polish run as two lean kernels (each hot loop fits the instruction cache); the fused kernel always follows as the
      // catch-all: it shares the list's work counter (so it finds it drained) and takes the fall-back list.  With no
      // instances expected only the fused kernel is launched, as before -- the result never depends on the guess.
      SolveArgs p = a;
      p.fail_perm = nullptr; p.fail_count = nullptr; p.route = 0; p.ready = nullptr;
      int32_t* pol_perm = h->d_perm + (size_t)(2 * kNumClasses + c) * h->max_batch;
      int32_t* fb_perm = h->d_perm + (size_t)(3 * kNumClasses + c) * h->max_batch;
      int32_t* pol_count = cnt + 4 * kNumClasses + c; int32_t* pol_work = cnt + 5 * kNumClasses + c;
      int32_t* fb_count = cnt + 6 * kNumClasses + c; int32_t* fb_work = cnt + 7 * kNumClasses + c;
      const cmpc_handle::ClassPlan& cp = h->cls[c];
      // (next to the stage-wise kernel no early trigger: the resident-but-waiting CTAs of the next kernel would take the SMs
      // the stage-wise kernel is meant to get -- whichever arrived first won, and the routes ran one after the other)
      auto chain = [&](SolveArgs& x) { chained(x); if (rip_grid > 0) { x.pdl_trigger = 0; if (first_dense) x.pdl = 0; } first_dense = false; };
      if (h->split && presolve && !warm && h->cfg.polish && cp.m_in_smem && cp.W <= h->split_maxw && cp.d_xstate && h->h_hint[c] > 0) {
        SolveArgs q = p;
        q.perm = in_perm[c]; q.count = in_count[c]; q.work = in_work[c];
        q.pol_perm = pol_perm; q.pol_count = pol_count; q.fb_perm = fb_perm; q.fb_count = fb_count;
        chain(q);
        // (next to the stage-wise kernel: only the CTAs that find a free SM, so that the polish does not wait for late-comers)
        int rc = launch_class<0>(h, cp, q, 1, rip_grid > 0 && h->overlap_trim ? std::max(1, h->num_sms - rip_grid) : 0);
        if (rc) return rc;
        ++launches;
        q.perm = pol_perm; q.count = pol_count; q.work = pol_work;
        chain(q);
        rc = launch_class<0>(h, cp, q, 2);
        if (rc) return rc;
        ++launches;
      }
      p.nlists = 2;
      p.lperm[0] = in_perm[c]; p.lcount[0] = in_count[c]; p.lwork[0] = in_work[c];
      p.lperm[1] = fb_perm; p.lcount[1] = fb_count; p.lwork[1] = fb_work;
      p.hint_out = h->h_hint_dev + c; p.hint_shadow = h->d_hint_shadow + c;
      chain(p);
      int rc = launch_class<0>(h, cp, p);
      if (rc) return rc;
      ++launches;
    }
  }
  if (rip_grid > 0) {
    if (cudaStreamWaitEvent(h->stream, h->ev_join, 0) != cudaSuccess) return fail(h, CMPC_ERR_CUDA, "join the auxiliary stream");
  } else if (rip.nlists > 0) {
    chained(rip);
    int rc = launch_ripm(h, rip);
    if (rc) return rc;
    ++launches;
  }
  if (capturing && cudaMemsetAsync(cnt, 0, kCounts * sizeof(int32_t), h->stream) != cudaSuccess) return fail(h, CMPC_ERR_CUDA, "memset counts");
  return launches;
}

// Fill a class plan: groups per CTA from the shared-memory budget, grid = one CTA per SM.
int plan_class(cmpc_handle* h, cmpc_handle::ClassPlan& p, int W, int nbmax, int mode) {
  const int N = h->cfg.horizon, L = h->cfg.num_legs;
  p.W = W; p.nbmax = nbmax; p.n4max = ((3 * nbmax + 3) / 4) * 4;
  p.m_in_smem = 1;
  SmemPlan sp = make_plan(N, L, W, p.nbmax, p.n4max, 1);
  if (mode == 1 || ((size_t)sp.total + sp.cta) * 8 > kMaxSmem) { p.m_in_smem = 0; sp = make_plan(N, L, W, p.nbmax, p.n4max, 0); }
  if (((size_t)sp.total + sp.cta) * 8 > kMaxSmem) return fail(h, CMPC_ERR_ARG, "horizon too large for the shared-memory vectors");
  const int gmax = (W == 2 ? 512 : (W == 1 ? 320 : 256)) / (32 * W);
  p.groups = (int)std::min<size_t>((size_t)gmax, (kMaxSmem / 8 - sp.cta) / (size_t)sp.total);
  if (p.groups < 1) p.groups = 1;
  p.smem_bytes = ((size_t)sp.total * p.groups + sp.cta) * 8;
  p.grid = h->num_sms;
  // H (and the factor when it does not fit on chip) + the polish's null-space bases (9 doubles per leg-step)
  p.scratch_per_group = (size_t)mat_region_doubles(N, L, p.n4max) * (p.m_in_smem ? 1 : 2) + (size_t)((9 * p.nbmax + 1) & ~1);
  CUDA_TRY(h, cudaMalloc(&p.d_scratch, p.scratch_per_group * 8 * (size_t)p.grid * p.groups));
  if (mode == 0 && p.m_in_smem) {  // iterate slots of the phase-split interior point
    p.xstride = (p.n4max + 10 * p.nbmax + 2 + 1) & ~1;
    CUDA_TRY(h, cudaMalloc(&p.d_xstate, (size_t)h->max_batch_plan * p.xstride * 8));
  }
  p.used = true;
  if (mode == 0 && p.m_in_smem) {  // presolve variant: the matrix and one vector per group (make_pre_plan)
    const PrePlan pp = make_pre_plan(N, L, W, p.nbmax, p.n4max);
    const int pgmax = (W == 1 ? 448 : 256) / (32 * W);
    const size_t room = kMaxSmem / 8 > (size_t)pp.cta ? kMaxSmem / 8 - pp.cta : 0;
    p.pre_groups = (int)std::min<size_t>((size_t)pgmax, room / (size_t)pp.total);
    if (p.pre_groups >= 1) {
      p.pre_smem_bytes = ((size_t)pp.cta + (size_t)pp.total * p.pre_groups) * 8;
      p.pre_scratch_per_group = 0;  // (the presolve kernel needs no L2 scratch any more: its verification does not use H)
      p.pre_used = true;
    }
  }
  return CMPC_OK;
}

int set_smem_attr(cmpc_handle* h, int W, int mode, bool ms, size_t bytes) {
  CUDA_TRY(h, set_solve_kernel_smem(W, mode, ms, bytes));
  return CMPC_OK;
}
int set_pre_smem_attr(cmpc_handle* h, int W, size_t bytes) {
  CUDA_TRY(h, set_presolve_kernel_smem(W, bytes));
  return CMPC_OK;
}

int collect_stats(cmpc_handle* h, int B, const int32_t* d_status, const int32_t* d_iters, const double* d_kkt,
                  cmpc_stats* stats, int launches) {
  CUDA_TRY(h, cudaMemsetAsync(h->d_stats, 0, sizeof(DevStats), h->stream));
  stats_kernel<<<(B + 255) / 256, 256, 0, h->stream>>>(d_status, d_iters, d_kkt, B, (DevStats*)h->d_stats);
  DevStats hs;
  CUDA_TRY(h, cudaMemcpyAsync(&hs, h->d_stats, sizeof(hs), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  stats->mean_iters = B ? (double)hs.iters_sum / B : 0.0;
  stats->max_iters = hs.max_iters;
  stats->n_ok = hs.n_ok; stats->n_ok_ipm = hs.n_ok_ipm; stats->n_max_iter = hs.n_max_iter;
  stats->n_invalid = hs.n_invalid; stats->n_numerical = hs.n_numerical;
  double mk; std::memcpy(&mk, &hs.max_kkt_bits, 8);
  stats->max_kkt = mk;
  stats->launches = launches;
  return CMPC_OK;
}

}  // namespace

// Frees every device / pinned allocation, stream and event of the handle and returns it to the state after cmpc_create
// (shared by cmpc_destroy and by a cmpc_setup that failed half-way).
static void release_device_state(cmpc_handle* h) {
  DeviceGuard g(h->device >= 0 ? h->device : 0);
  cudaFree(h->d_state); cudaFree(h->d_ds); cudaFree(h->d_di); cudaFree(h->d_forces); cudaFree(h->d_kkt);
  cudaFree(h->d_lam); cudaFree(h->d_hip); cudaFree(h->d_counts); cudaFree(h->d_perm); cudaFree(h->exp_plan.d_scratch);
  for (auto& c : h->cls) { cudaFree(c.d_scratch); cudaFree(c.d_pre_scratch); cudaFree(c.d_xstate); c = cmpc_handle::ClassPlan(); }
  cudaFreeHost(h->h_hint); h->h_hint = nullptr; h->h_hint_dev = nullptr;
  cudaFree(h->d_hint_shadow); h->d_hint_shadow = nullptr;
  h->exp_plan = cmpc_handle::ClassPlan();
  cudaFree(h->d_ric_scratch); cudaFree(h->d_rip_scratch);
  cudaFree(h->d_ready); cudaFreeHost(h->h_ready_vals); cudaFreeHost(h->h_error);
  cudaFree(h->d_status); cudaFree(h->d_iters);
  cudaFree(h->d_iters_sum); cudaFree(h->d_status_or); cudaFree(h->d_active); cudaFree(h->d_stats);
  cudaFree(h->d_flog); cudaFree(h->d_sqp);
  h->d_state = h->d_ds = h->d_di = h->d_forces = h->d_kkt = h->d_lam = h->d_flog = h->d_hip = h->d_sqp = nullptr;
  h->d_status = h->d_iters = h->d_iters_sum = h->d_status_or = nullptr; h->d_active = nullptr; h->d_stats = nullptr;
  h->d_counts = h->d_perm = nullptr; h->d_ric_scratch = h->d_rip_scratch = nullptr;
  h->d_ready = nullptr; h->h_ready_vals = nullptr; h->h_error = nullptr; h->h_error_dev = nullptr;
  for (auto& e : h->ev) if (e) { cudaEventDestroy(e); e = nullptr; }
  for (auto& e : h->ev_span) if (e) { cudaEventDestroy(e); e = nullptr; }
  for (auto& e : h->ev_in) if (e) { cudaEventDestroy(e); e = nullptr; }
  for (auto& e : h->ev_k) if (e) { cudaEventDestroy(e); e = nullptr; }
  if (h->s_in) { cudaStreamDestroy(h->s_in); h->s_in = nullptr; }
  if (h->s_out) { cudaStreamDestroy(h->s_out); h->s_out = nullptr; }
  if (h->s_aux) { cudaStreamDestroy(h->s_aux); h->s_aux = nullptr; }
  if (h->ev_fork) { cudaEventDestroy(h->ev_fork); h->ev_fork = nullptr; }
  if (h->ev_join) { cudaEventDestroy(h->ev_join); h->ev_join = nullptr; }
  if (h->own_stream && h->stream) { cudaStreamDestroy(h->stream); h->stream = nullptr; h->own_stream = false; }
  h->ric_used = h->rip_used = false;
  h->ready = false; h->max_batch = 0; h->flog_cap = 0;
}

extern "C" {

const char* cmpc_version(void) { return "cmpc_b200 0.3 (sm_100a)"; }

int cmpc_config_init(cmpc_config* cfg, double mass, int num_legs, int horizon, double dt, const double* weights,
                     const double* mu) {
  if (!cfg || !weights || !mu || num_legs < 1 || num_legs > CMPC_MAX_LEGS) return CMPC_ERR_ARG;
  std::memset(cfg, 0, sizeof(*cfg));
  cfg->mass = mass; cfg->num_legs = num_legs; cfg->horizon = horizon; cfg->dt = dt;
  for (int i = 0; i < num_legs; ++i) cfg->mu[i] = mu[i];
  for (int i = 0; i < 9 + 9 * num_legs; ++i) cfg->weights[i] = weights[i];
  cfg->disc_mode = 0; cfg->max_iter = 50; cfg->ipm_tol = 1e-9; cfg->polish = 1; cfg->presolve = 1;
  return valid_config(cfg) ? CMPC_OK : CMPC_ERR_ARG;
}

int cmpc_create(const cmpc_config* cfg, cmpc_handle** out) {
  if (!cfg || !out) return CMPC_ERR_ARG;
  *out = nullptr;
  if (!valid_config(cfg)) return CMPC_ERR_ARG;
  cmpc_handle* h = new (std::nothrow) cmpc_handle();
  if (!h) return CMPC_ERR_ARG;
  h->cfg = *cfg;
  fill_dev(h);
  h->zero_copy = !(getenv("CMPC_NO_ZEROCOPY") && atoi(getenv("CMPC_NO_ZEROCOPY")) != 0);
  *out = h;
  return CMPC_OK;
}

static int setup_impl(cmpc_handle* h, int max_batch, int device) {
  h->device = device;
  h->max_batch_plan = max_batch;
  SET_DEVICE(h);
  cudaDeviceProp prop;
  CUDA_TRY(h, cudaGetDeviceProperties(&prop, device));
  h->num_sms = prop.multiProcessorCount;
  if (!h->stream) { CUDA_TRY(h, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)); h->own_stream = true; }
  for (auto& e : h->ev) CUDA_TRY(h, cudaEventCreate(&e));
  CUDA_TRY(h, cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
  CUDA_TRY(h, cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
  CUDA_TRY(h, cudaStreamCreateWithFlags(&h->s_aux, cudaStreamNonBlocking));
  CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  if (const char* m = getenv("CMPC_OVERLAP")) h->overlap = atoi(m) != 0;
  if (const char* m = getenv("CMPC_OVERLAP_TRIM")) h->overlap_trim = atoi(m) != 0;
  if (const char* m = getenv("CMPC_OVERLAP_PCT")) h->overlap_pct = std::max(25, std::min(400, atoi(m)));
  for (auto& e : h->ev_k) CUDA_TRY(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (auto& e : h->ev_span) CUDA_TRY(h, cudaEventCreate(&e));
  for (auto& e : h->ev_in) CUDA_TRY(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));

  const int N = h->cfg.horizon, L = h->cfg.num_legs;
  const size_t ns = 9 + 3 * L, nds = 9 * (N + 1), ndi = (size_t)L * (4 * N + 3), nf = (size_t)3 * L * N;
  const size_t B = (size_t)max_batch;
  CUDA_TRY(h, cudaMalloc(&h->d_state, B * ns * 8));
  CUDA_TRY(h, cudaMalloc(&h->d_ds, B * nds * 8));
  CUDA_TRY(h, cudaMalloc(&h->d_di, B * ndi * 8));
  CUDA_TRY(h, cudaMalloc(&h->d_forces, B * nf * 8));
  CUDA_TRY(h, cudaMalloc(&h->d_kkt, B * 8));
  CUDA_TRY(h, cudaMalloc(&h->d_lam, B * 10 * L * N * 8));
  CUDA_TRY(h, cudaMalloc(&h->d_hip, B * 3 * L * 8));
  CUDA_TRY(h, cudaMalloc(&h->d_sqp, B * (ndi + 17) * 8));   // cmpc_solve_batch_sqp: original inputs + defects (sqp_iters <= 16)
  CUDA_TRY(h, cudaMalloc(&h->d_status, B * 4));
  CUDA_TRY(h, cudaMalloc(&h->d_iters, B * 4));
  CUDA_TRY(h, cudaMalloc(&h->d_iters_sum, B * 4));
  CUDA_TRY(h, cudaMalloc(&h->d_status_or, B * 4));
  CUDA_TRY(h, cudaMalloc(&h->d_active, B * L * N * 2));
  CUDA_TRY(h, cudaMalloc(&h->d_stats, sizeof(DevStats)));

  CUDA_TRY(h, cudaMalloc(&h->d_ready, 4 * sizeof(int32_t)));
  CUDA_TRY(h, cudaHostAlloc(&h->h_ready_vals, cmpc_handle::kMaxChunks * sizeof(int32_t), cudaHostAllocDefault));
  for (int c = 0; c < cmpc_handle::kMaxChunks; ++c) h->h_ready_vals[c] = c + 1;
  CUDA_TRY(h, cudaHostAlloc(&h->h_error, sizeof(int32_t), cudaHostAllocMapped));
  *h->h_error = 0;
  CUDA_TRY(h, cudaHostGetDevicePointer((void**)&h->h_error_dev, h->h_error, 0));
  if (const char* m = getenv("CMPC_E2E_MODE")) h->e2e_mode = atoi(m);
  if (const char* m = getenv("CMPC_E2E_CHUNK")) h->e2e_chunk = atoi(m);
  h->debug_tune = getenv("CMPC_DEBUG_TUNE") != nullptr;
  h->debug_timeline = getenv("CMPC_DEBUG_TIMELINE") != nullptr;
  h->debug_plan = getenv("CMPC_DEBUG_PLAN") != nullptr;
  CUDA_TRY(h, cudaMalloc(&h->d_counts, 2 * kCounts * sizeof(int32_t)));
  h->counts_cur = 0; h->counts_zero[0] = h->counts_zero[1] = false;
  CUDA_TRY(h, cudaMalloc(&h->d_perm, (size_t)4 * kNumClasses * B * sizeof(int32_t)));
  CUDA_TRY(h, cudaHostAlloc(&h->h_hint, (kNumClasses + 1) * sizeof(int32_t), cudaHostAllocMapped));  // [kNumClasses]: the stage-wise interior point's lists
  for (int c = 0; c <= kNumClasses; ++c) h->h_hint[c] = 0;
  CUDA_TRY(h, cudaHostGetDevicePointer((void**)&h->h_hint_dev, h->h_hint, 0));
  CUDA_TRY(h, cudaMalloc(&h->d_hint_shadow, (kNumClasses + 1) * sizeof(int32_t)));
  CUDA_TRY(h, cudaMemset(h->d_hint_shadow, 0, (kNumClasses + 1) * sizeof(int32_t)));
  if (const char* m = getenv("CMPC_SPLIT")) h->split = atoi(m) != 0;
  if (const char* m = getenv("CMPC_SPLIT_MAXW")) h->split_maxw = atoi(m);
  if (const char* m = getenv("CMPC_RIPM_MINCLASS")) h->rip_minclass = atoi(m);
  if (const char* m = getenv("CMPC_PDL")) h->pdl = atoi(m) != 0;
  if (const char* m = getenv("CMPC_PDL_TRIGGER")) h->pdl_trigger = atoi(m);
  if (const char* m = getenv("CMPC_RIPM_LOCK")) h->rip_phase_lock = atoi(m);
  if (const char* m = getenv("CMPC_DENSE_LOCK")) h->dense_lock = atoi(m);
  // size classes by number of free 3-blocks: n4 <= 64 -> one warp per instance, n4 <= 128 ->
  // four warps, larger -> a whole 256-thread CTA; the factor lives in shared memory whenever it fits
  {
    const int nbfull = L * N;
    const int cap[kNumClasses] = {20, 42, 64, nbfull};
    int w0 = getenv("CMPC_W0") ? atoi(getenv("CMPC_W0")) : 1;   // experiment knob: warps per instance of size class 0
    if (w0 != 1 && w0 != 2 && w0 != 4 && w0 != 8) w0 = 1;        // (only these kernel variants exist)
    const int Wc[kNumClasses] = {w0, 4, 8, 8};
    int b[kNumClasses];
    int lower = 0;
    for (int c = 0; c < kNumClasses; ++c) {
      b[c] = std::min(cap[c], nbfull);
      if (c == kNumClasses - 1) b[c] = nbfull;
      if (lower < nbfull && b[c] > lower) {
        int rc = plan_class(h, h->cls[c], Wc[c], b[c], 0);
        if (rc) return rc;
      }
      lower = std::max(lower, b[c]);
    }
    h->bounds = make_int4(b[0], b[1], b[2], b[3]);
    int rc = plan_class(h, h->exp_plan, 8, nbfull, 1);
    if (rc) return rc;
    size_t s1 = 0, s2 = 0, s4 = 0, s8 = 0, s8g = 0, p1 = 0, p4 = 0, p8 = 0;
    for (int c = 0; c < kNumClasses; ++c) {
      if (!h->cls[c].used) continue;
      if (h->cls[c].pre_used && h->cls[c].W == 2) h->cls[c].pre_used = false;  // no two-warp presolve variant
      if (h->cls[c].pre_used) {
        size_t& pref = h->cls[c].W == 1 ? p1 : h->cls[c].W == 4 ? p4 : p8;
        pref = std::max(pref, h->cls[c].pre_smem_bytes);
      }
      if (!h->cls[c].m_in_smem) h->cls[c].W = 8;  // the global-factor variant exists for W = 8 only
      size_t& sref = !h->cls[c].m_in_smem ? s8g : h->cls[c].W == 1 ? s1 : h->cls[c].W == 2 ? s2 : h->cls[c].W == 4 ? s4 : s8;
      sref = std::max(sref, h->cls[c].smem_bytes);
    }
    if (s1 && (rc = set_smem_attr(h, 1, 0, true, s1))) return rc;
    if (s2 && (rc = set_smem_attr(h, 2, 0, true, s2))) return rc;
    if (s4 && (rc = set_smem_attr(h, 4, 0, true, s4))) return rc;
    if (s8 && (rc = set_smem_attr(h, 8, 0, true, s8))) return rc;
    if (s8g && (rc = set_smem_attr(h, 8, 0, false, s8g))) return rc;
    if ((rc = set_smem_attr(h, 8, 1, false, h->exp_plan.smem_bytes))) return rc;
    {
      const RicPlan rp = make_ric_plan(N, L);
      h->ric_groups = (int)std::min<size_t>(14, (kMaxSmem / 8 - 102) / (size_t)rp.total);  // 102 doubles: CTA-shared tables (kRicCta)
      if (h->ric_groups >= 1) {
        h->ric_smem_bytes = ((size_t)rp.total * h->ric_groups + 102) * 8;
        CUDA_TRY(h, cudaMalloc(&h->d_ric_scratch, (size_t)rp.slab * 8 * (size_t)h->num_sms * h->ric_groups));
        CUDA_TRY(h, set_riccati_kernel_smem(h->ric_smem_bytes));
        h->ric_used = true;
      }
    }
    {
      int gd = 0, cd = 0, sd = 0;
      ripm_sizes(N, L, &gd, &cd, &sd);
      h->rip_groups = (int)std::min<size_t>(12, (kMaxSmem / 8 - cd) / (size_t)gd);
      if (const char* m = getenv("CMPC_RIPM_GROUPS")) h->rip_groups = std::max(1, std::min(h->rip_groups, atoi(m)));
      if (const char* m = getenv("CMPC_IPM_BACKEND")) h->rip_mode = std::max(0, std::min(2, atoi(m)));
      if (h->rip_groups >= 1) {
        h->rip_smem_bytes = ((size_t)gd * h->rip_groups + cd) * 8;
        h->rip_slab = ((size_t)sd + 1) & ~(size_t)1;
        CUDA_TRY(h, cudaMalloc(&h->d_rip_scratch, h->rip_slab * 8 * (size_t)h->num_sms * h->rip_groups));
        CUDA_TRY(h, set_ripm_kernel_smem(h->rip_smem_bytes));
        h->rip_used = true;
      }
    }
    if (p1 && (rc = set_pre_smem_attr(h, 1, p1))) return rc;
    if (p4 && (rc = set_pre_smem_attr(h, 4, p4))) return rc;
    if (p8 && (rc = set_pre_smem_attr(h, 8, p8))) return rc;
  }
  h->max_batch = max_batch;
  h->ready = true;
  return CMPC_OK;
}

int cmpc_setup(cmpc_handle* h, int max_batch, int device) {
  if (!h || max_batch < 1) return fail(h, CMPC_ERR_ARG, "cmpc_setup: bad arguments");
  if (h->ready) return fail(h, CMPC_ERR_STATE, "cmpc_setup called twice");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(h, CMPC_ERR_NO_DEVICE, "no CUDA device: this library has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(h, CMPC_ERR_ARG, "cmpc_setup: device out of range");
  const int rc = setup_impl(h, max_batch, device);
  if (rc != CMPC_OK) {  // nothing half-allocated survives a failed setup: a retry starts from a clean handle
    const std::string msg = h->err;
    release_device_state(h);
    h->err = msg;
  }
  return rc;
}

int cmpc_update_weights(cmpc_handle* h, const double* w, int n) {
  if (!h || !w || n != 9 + 9 * h->cfg.num_legs) return fail(h, CMPC_ERR_ARG, "cmpc_update_weights: need 9+9*num_legs weights");
  cmpc_config c = h->cfg;
  for (int i = 0; i < n; ++i) c.weights[i] = w[i];
  if (!valid_config(&c)) return fail(h, CMPC_ERR_ARG, "cmpc_update_weights: invalid weights");
  h->cfg = c;
  fill_dev(h);  // DevConfig travels as a kernel argument: nothing to upload
  return CMPC_OK;
}

int cmpc_set_stream(cmpc_handle* h, void* s) {
  if (!h) return CMPC_ERR_ARG;
  if (h->ready && s) {  // the stream must live on the handle's device (before cmpc_setup it is checked by the first launch)
    int dv = -1;
    if (cudaStreamGetDevice((cudaStream_t)s, &dv) == cudaSuccess && dv != h->device)
      return fail(h, CMPC_ERR_ARG, "cmpc_set_stream: the stream belongs to another device");
    cudaGetLastError();
  }
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  h->stream = (cudaStream_t)s;
  h->own_stream = false;
  h->counts_zero[0] = h->counts_zero[1] = false;  // (zeroed in the old stream's order: the next call does it again)
  return CMPC_OK;
}

int cmpc_synchronize(cmpc_handle* h) {
  if (!h || !h->ready) return fail(h, CMPC_ERR_STATE, "not set up");
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return CMPC_OK;
}

int cmpc_solve_batch_device(cmpc_handle* h, int B, const double* d_state, const double* d_des_state,
                            const double* d_des_inputs, double* d_forces, int32_t* d_status, int32_t* d_iters,
                            double* d_kkt, double* d_lam, uint16_t* d_active, cmpc_stats* stats) {
  if (!h || !h->ready) return fail(h, CMPC_ERR_STATE, "cmpc_solve_batch_device: call cmpc_setup first");
  if (B < 0 || B > h->max_batch) return fail(h, CMPC_ERR_STATE, "batch exceeds max_batch given to cmpc_setup");
  if (!d_state || !d_des_state || !d_des_inputs || !d_forces || !d_status) return fail(h, CMPC_ERR_ARG, "null buffer");
  if (B == 0) { if (stats) std::memset(stats, 0, sizeof(*stats)); return CMPC_OK; }
  SET_DEVICE(h);
  SolveArgs a = SolveArgs();
  a.state = d_state; a.des_state = d_des_state; a.des_inputs = d_des_inputs;
  a.forces = d_forces; a.status = d_status;
  a.iters = d_iters ? d_iters : (stats ? h->d_iters : nullptr);
  a.kkt = d_kkt ? d_kkt : (stats ? h->d_kkt : nullptr);
  a.lam = d_lam; a.active = d_active;
  if (stats) CUDA_TRY(h, cudaEventRecord(h->ev[0], h->stream));
  int rc = launch_solve(h, a, B);
  if (rc < 0) return rc;
  const int launches = rc;
  if (stats) {
    CUDA_TRY(h, cudaEventRecord(h->ev[1], h->stream));
    std::memset(stats, 0, sizeof(*stats));
    rc = collect_stats(h, B, d_status, a.iters, a.kkt, stats, launches);
    if (rc) return rc;
    float ms = 0;
    CUDA_TRY(h, cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
    stats->kernel_ms = ms;
  }
  return CMPC_OK;
}

int cmpc_solve_batch(cmpc_handle* h, int B, const double* state, const double* des_state, const double* des_inputs,
                     double* forces, int32_t* status, int32_t* iters, double* kkt, double* lam, uint16_t* active,
                     cmpc_stats* stats) {
  if (!h || !h->ready) return fail(h, CMPC_ERR_STATE, "cmpc_solve_batch: call cmpc_setup first");
  if (B < 0 || B > h->max_batch) return fail(h, CMPC_ERR_STATE, "batch exceeds max_batch given to cmpc_setup");
  if (!state || !des_state || !des_inputs || !forces || !status) return fail(h, CMPC_ERR_ARG, "null buffer");
  if (B == 0) { if (stats) std::memset(stats, 0, sizeof(*stats)); return CMPC_OK; }
  const auto t_enter = std::chrono::steady_clock::now();
  SET_DEVICE(h);
  const int N = h->cfg.horizon, L = h->cfg.num_legs;
  const size_t ns = 9 + 3 * L, nds = 9 * (N + 1), ndi = (size_t)L * (4 * N + 3), nf = (size_t)3 * L * N;
  cudaStream_t s = h->stream;
  const bool router = h->cls[0].used && presolve_kind(h, 0) == 1;
  // device aliases of pinned (page-locked, hence mapped under UVA) caller buffers
  void* dp[9] = {nullptr};
  const void* hp[9] = {state, des_state, des_inputs, forces, status, iters, kkt, lam, active};
  bool in_mapped = h->zero_copy, out_mapped = h->zero_copy;
  for (int q = 0; q < 9; ++q) {
    if (!hp[q]) continue;
    dp[q] = mapped_device_pointer(hp[q]);
    if (!dp[q]) (q < 3 ? in_mapped : out_mapped) = false;
  }
  // Inputs (pinned buffers), CMPC_E2E_MODE:
  //  0 (default) for batches of two waves or more: the first six calls time the pipelined and the zero-copy
  //    route alternately and the faster one is kept (re-tuned when the batch size changes); smaller batches
  //    use zero-copy.
  //    Pipelined: DMA copies in two chunks on the copy stream (54 GB/s), one launch sequence per chunk on
  //    the compute stream behind an event -- chunk 0 computes while chunk 1 is copied; plain stream
  //    dependencies, nothing polls.  0.34 ms per 4096-instance step, stable.
  //  1 zero-copy: the kernels read the inputs in place over the bus -- every byte crosses once, overlapped
  //    with the other resident instances' compute, no staging copy; SM-issued reads reach ~26 GB/s:
  //    0.36-0.40 ms per step.
  //  2 (and pageable buffers) staged: copy in, compute, copy out.
  //  3 progressive: copies in 1024-instance chunks while the router kernel, launched at the same time,
  //    waits per instance for its chunk; best case 0.335 ms but 0.34-0.70 ms over runs (the copies slow
  //    down erratically while 2000 warps poll).
  //  4 pipelined regardless of the batch size (>= 1024).
  //  5 full duplex: pipelined copy-in, and the outputs of every chunk leave by the copy engine on a third stream
  //    while the next chunk is still coming in (the SMs never wait on posted writes over the bus).
  // Automatic tuning (mode 0, pinned buffers, batches of two waves or more): the first two calls of a batch size are
  // warm-up (pipelined; page tables, clocks and the copy engines are cold), the next six time the three routes
  // twice each (library-side events), then the fastest is kept; cmpc_last_route() says which.
  const bool tunable = in_mapped && router && h->e2e_mode == 0 && B >= 4096;
  if (tunable && h->tune_batch != B) { h->tune_batch = B; h->tune_calls = 0; h->tune_best[0] = h->tune_best[1] = h->tune_best[2] = 1e30f; }
  const bool tuning = tunable && h->tune_calls >= 2 && h->tune_calls < 8;
  int kind = 0;  // 0 zero-copy, 1 pipelined, 2 full duplex
  if (tunable) {
    if (h->tune_calls < 2) kind = 1;
    else if (tuning) kind = (h->tune_calls - 2) % 3;
    else kind = h->tune_best[1] <= h->tune_best[0] ? (h->tune_best[2] < h->tune_best[1] ? 2 : 1) : (h->tune_best[2] < h->tune_best[0] ? 2 : 0);
  }
  if (in_mapped && router && B >= 1024 && (h->e2e_mode == 4 || h->e2e_mode == 5)) kind = h->e2e_mode == 4 ? 1 : 2;
  const bool pipelined = in_mapped && router && kind >= 1;
  const bool duplex = pipelined && kind == 2 && out_mapped;
  const bool zc_in = in_mapped && h->e2e_mode != 2 && h->e2e_mode != 3 && !pipelined;
  // (progressive needs pinned inputs: a pageable cudaMemcpyAsync is staged by the driver and can
  // serialise behind the running kernel, which would then wait for its chunk until the time-out)
  const bool progressive = in_mapped && !zc_in && router && h->e2e_mode == 3 && B >= 256;
  // Outputs: written in place over the bus when every output buffer is pinned (posted writes,
  // overlapped with the remaining compute); otherwise device buffers and one copy back.
  const bool zc_out = out_mapped && h->e2e_mode != 2 && !duplex;
  SolveArgs a = SolveArgs();
  a.state = zc_in ? (const double*)dp[0] : h->d_state;
  a.des_state = zc_in ? (const double*)dp[1] : h->d_ds;
  a.des_inputs = zc_in ? (const double*)dp[2] : h->d_di;
  if (zc_out) {  // (full duplex: device buffers, copied out per chunk)
    a.forces = (double*)dp[3]; a.status = (int32_t*)dp[4];
    a.iters = iters ? (int32_t*)dp[5] : (stats ? h->d_iters : nullptr);
    a.kkt = kkt ? (double*)dp[6] : (stats ? h->d_kkt : nullptr);
    a.lam = (double*)dp[7]; a.active = (uint16_t*)dp[8];
  } else {
    a.forces = h->d_forces; a.status = h->d_status; a.iters = h->d_iters; a.kkt = h->d_kkt;
    a.lam = lam ? h->d_lam : nullptr; a.active = active ? h->d_active : nullptr;
  }
  int nch = 1, per = B;
  if (pipelined) {
    const int want = h->e2e_chunk > 0 ? std::max(256, h->e2e_chunk) : 2048;
    nch = std::min((B + want - 1) / want, (int)cmpc_handle::kMaxChunks);
    per = (((B + nch - 1) / nch) + 31) & ~31;
    nch = (B + per - 1) / per;
  }
  if (progressive) {
    per = h->e2e_chunk > 0 ? std::max(32, h->e2e_chunk & ~31) : 1024;
    if ((B + per - 1) / per > cmpc_handle::kMaxChunks) per = (((B + cmpc_handle::kMaxChunks - 1) / cmpc_handle::kMaxChunks) + 31) & ~31;
    nch = (B + per - 1) / per;
    CUDA_TRY(h, cudaMemsetAsync(h->d_ready, 0, 4 * sizeof(int32_t), s));
  }
  CUDA_TRY(h, cudaEventRecord(h->ev_span[0], s));
  int launches = 0;
  if (!zc_in) {
    // pinned sources when progressive: these calls only enqueue (chunk c of the three arrays, then
    // the chunk counter); the router kernel launched below polls the counter
    cudaStream_t sc[1] = {(progressive || pipelined) ? h->s_in : s};
    if (progressive || pipelined) CUDA_TRY(h, cudaStreamWaitEvent(sc[0], h->ev_span[0], 0));
    for (int c = 0; c < nch; ++c) {
      const size_t o = (size_t)c * per;
      const size_t nbc = std::min<size_t>(per, (size_t)B - o);
      CUDA_TRY(h, cudaMemcpyAsync(h->d_state + o * ns, state + o * ns, nbc * ns * 8, cudaMemcpyHostToDevice, sc[0]));
      CUDA_TRY(h, cudaMemcpyAsync(h->d_ds + o * nds, des_state + o * nds, nbc * nds * 8, cudaMemcpyHostToDevice, sc[0]));
      CUDA_TRY(h, cudaMemcpyAsync(h->d_di + o * ndi, des_inputs + o * ndi, nbc * ndi * 8, cudaMemcpyHostToDevice, sc[0]));
      if (progressive)
        CUDA_TRY(h, cudaMemcpyAsync(h->d_ready, h->h_ready_vals + c, sizeof(int32_t), cudaMemcpyHostToDevice, sc[0]));
      if (pipelined) CUDA_TRY(h, cudaEventRecord(h->ev_in[c], sc[0]));
    }
    CUDA_TRY(h, cudaEventRecord(h->ev_span[1], sc[0]));
  } else {
    CUDA_TRY(h, cudaEventRecord(h->ev_span[1], s));
  }
  if (progressive) {
    int rc = launch_solve(h, a, B, h->d_ready, per);
    if (rc < 0) return rc;
    launches = rc;
  }
  if (pipelined) {
    for (int c = 0; c < nch; ++c) {
      const size_t o = (size_t)c * per;
      const int nbc = (int)std::min<size_t>(per, (size_t)B - o);
      SolveArgs ac = a;
      ac.state = a.state + o * ns; ac.des_state = a.des_state + o * nds; ac.des_inputs = a.des_inputs + o * ndi;
      ac.forces = a.forces + o * nf; ac.status = a.status + o;
      if (a.iters) ac.iters = a.iters + o;
      if (a.kkt) ac.kkt = a.kkt + o;
      if (a.lam) ac.lam = a.lam + o * 10 * L * N;
      if (a.active) ac.active = a.active + o * L * N;
      CUDA_TRY(h, cudaStreamWaitEvent(s, h->ev_in[c], 0));
      int rc = launch_solve(h, ac, nbc);
      if (rc < 0) return rc;
      launches += rc;
      if (duplex) {  // this chunk's results leave on the copy-out stream while the next chunk computes / copies in
        CUDA_TRY(h, cudaEventRecord(h->ev_k[c], s));
        CUDA_TRY(h, cudaStreamWaitEvent(h->s_out, h->ev_k[c], 0));
        CUDA_TRY(h, cudaMemcpyAsync(forces + o * nf, h->d_forces + o * nf, (size_t)nbc * nf * 8, cudaMemcpyDeviceToHost, h->s_out));
        CUDA_TRY(h, cudaMemcpyAsync(status + o, h->d_status + o, (size_t)nbc * 4, cudaMemcpyDeviceToHost, h->s_out));
        if (iters) CUDA_TRY(h, cudaMemcpyAsync(iters + o, h->d_iters + o, (size_t)nbc * 4, cudaMemcpyDeviceToHost, h->s_out));
        if (kkt) CUDA_TRY(h, cudaMemcpyAsync(kkt + o, h->d_kkt + o, (size_t)nbc * 8, cudaMemcpyDeviceToHost, h->s_out));
        if (lam) CUDA_TRY(h, cudaMemcpyAsync(lam + o * 10 * L * N, h->d_lam + o * 10 * L * N, (size_t)nbc * 10 * L * N * 8, cudaMemcpyDeviceToHost, h->s_out));
        if (active) CUDA_TRY(h, cudaMemcpyAsync(active + o * L * N, h->d_active + o * L * N, (size_t)nbc * L * N * 2, cudaMemcpyDeviceToHost, h->s_out));
      }
    }
    if (duplex) {  // the span ends when the last copy-out has landed
      CUDA_TRY(h, cudaEventRecord(h->ev_k[0], h->s_out));
      CUDA_TRY(h, cudaStreamWaitEvent(s, h->ev_k[0], 0));
    }
  } else if (!progressive) {
    int rc = launch_solve(h, a, B);
    if (rc < 0) return rc;
    launches = rc;
  }
  CUDA_TRY(h, cudaEventRecord(h->ev_span[2], s));
  if (!zc_out && !duplex) {
    CUDA_TRY(h, cudaMemcpyAsync(forces, h->d_forces, (size_t)B * nf * 8, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(h, cudaMemcpyAsync(status, h->d_status, (size_t)B * 4, cudaMemcpyDeviceToHost, s));
    if (iters) CUDA_TRY(h, cudaMemcpyAsync(iters, h->d_iters, (size_t)B * 4, cudaMemcpyDeviceToHost, s));
    if (kkt) CUDA_TRY(h, cudaMemcpyAsync(kkt, h->d_kkt, (size_t)B * 8, cudaMemcpyDeviceToHost, s));
    if (lam) CUDA_TRY(h, cudaMemcpyAsync(lam, h->d_lam, (size_t)B * 10 * L * N * 8, cudaMemcpyDeviceToHost, s));
    if (active) CUDA_TRY(h, cudaMemcpyAsync(active, h->d_active, (size_t)B * L * N * 2, cudaMemcpyDeviceToHost, s));
  }
  CUDA_TRY(h, cudaEventRecord(h->ev_span[3], s));
  const auto t_enq = std::chrono::steady_clock::now();
  if (progressive || pipelined) CUDA_TRY(h, cudaStreamSynchronize(h->s_in));
  if (duplex) CUDA_TRY(h, cudaStreamSynchronize(h->s_out));
  CUDA_TRY(h, cudaStreamSynchronize(s));
  if (*h->h_error) { *h->h_error = 0; return fail(h, CMPC_ERR_CUDA, "cmpc_solve_batch: input chunk did not arrive (copy stream stalled)"); }
  if (h->debug_timeline) {
    const auto t_done = std::chrono::steady_clock::now();
    auto us = [&](std::chrono::steady_clock::time_point a_, std::chrono::steady_clock::time_point b_) { return std::chrono::duration<double, std::micro>(b_ - a_).count(); };
    fprintf(stderr, "[cmpc] timeline: host enqueue %.0f us, host total %.0f us; device (us from start):", us(t_enter, t_enq), us(t_enter, t_done));
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev_span[0], h->ev_span[1]); fprintf(stderr, " copy-in-end %.0f", ms * 1e3f);
    cudaEventElapsedTime(&ms, h->ev_span[0], h->ev_span[2]); fprintf(stderr, " kernels-end %.0f", ms * 1e3f);
    cudaEventElapsedTime(&ms, h->ev_span[0], h->ev_span[3]); fprintf(stderr, " end %.0f\n", ms * 1e3f);
  }
  if (tunable && h->tune_calls < 8) {
    float span = 0;
    CUDA_TRY(h, cudaEventElapsedTime(&span, h->ev_span[0], h->ev_span[3]));
    if (tuning) { float& best = h->tune_best[kind]; best = std::min(best, span); }
    ++h->tune_calls;
    if (h->debug_tune) fprintf(stderr, "[cmpc] tune call %d: route %d %.3f ms (best zero-copy %.3f, pipelined %.3f, duplex %.3f)\n", h->tune_calls, kind, span, h->tune_best[0], h->tune_best[1], h->tune_best[2]);
  }
  snprintf(h->route, sizeof(h->route), "%s%s%s", progressive ? "progressive" : pipelined ? (duplex ? "full-duplex" : "pipelined") : zc_in ? "zero-copy" : "staged",
           pipelined ? (nch == 2 ? " x2 chunks" : " chunks") : "", tunable ? (h->tune_calls < 8 ? " (tuning)" : " (tuned)") : "");
  if (stats) {
    std::memset(stats, 0, sizeof(*stats));
    int rc = collect_stats(h, B, a.status, a.iters, a.kkt, stats, launches);
    if (rc) return rc;
    float t0 = 0, t1 = 0, t2 = 0;
    CUDA_TRY(h, cudaEventElapsedTime(&t1, h->ev_span[0], h->ev_span[2]));  // until the last kernel ended
    CUDA_TRY(h, cudaEventElapsedTime(&t2, h->ev_span[2], h->ev_span[3]));  // copy-out tail after it
    if (!zc_in) CUDA_TRY(h, cudaEventElapsedTime(&t0, h->ev_span[0], h->ev_span[1]));  // copy-in span (overlaps the kernels when progressive)
    stats->h2d_ms = t0; stats->kernel_ms = t1; stats->d2h_ms = t2;
  }
  return CMPC_OK;
}

int cmpc_build_batch(cmpc_handle* h, int B, const double* state, const double* des_state, const double* des_inputs,
                     double* H, double* g, int32_t* status) {
  if (!h || !h->ready) return fail(h, CMPC_ERR_STATE, "cmpc_build_batch: call cmpc_setup first");
  if (B < 0 || B > h->max_batch) return fail(h, CMPC_ERR_STATE, "batch exceeds max_batch given to cmpc_setup");
  if (!state || !des_state || !des_inputs || !H || !g || !status) return fail(h, CMPC_ERR_ARG, "null buffer");
  if (B == 0) return CMPC_OK;
  SET_DEVICE(h);
  const int N = h->cfg.horizon, L = h->cfg.num_legs;
  const size_t ns = 9 + 3 * L, nds = 9 * (N + 1), ndi = (size_t)L * (4 * N + 3), p = (size_t)3 * L * N;
  cudaStream_t s = h->stream;
  DevBuf bH, bg;  // test/diagnostic path: temporary buffers, released on every return path
  CUDA_TRY(h, bH.alloc((size_t)B * p * p * 8));
  CUDA_TRY(h, bg.alloc((size_t)B * p * 8));
  double *dH = bH.as<double>(), *dg = bg.as<double>();
  CUDA_TRY(h, cudaMemcpyAsync(h->d_state, state, B * ns * 8, cudaMemcpyHostToDevice, s));
  CUDA_TRY(h, cudaMemcpyAsync(h->d_ds, des_state, B * nds * 8, cudaMemcpyHostToDevice, s));
  CUDA_TRY(h, cudaMemcpyAsync(h->d_di, des_inputs, B * ndi * 8, cudaMemcpyHostToDevice, s));
  SolveArgs a = SolveArgs();
  a.state = h->d_state; a.des_state = h->d_ds; a.des_inputs = h->d_di;
  a.status = h->d_status; a.Hout = dH; a.gout = dg; a.forces = h->d_forces;
  a.count = nullptr; a.count_imm = B; a.perm = nullptr; a.work = h->d_counts + kNumClasses;
  cudaMemsetAsync(h->d_counts, 0, 2 * kNumClasses * sizeof(int32_t), s);
  h->counts_zero[0] = false;
  int rc = launch_class<1>(h, h->exp_plan, a);
  if (rc == CMPC_OK) {
    cudaMemcpyAsync(H, dH, (size_t)B * p * p * 8, cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(g, dg, (size_t)B * p * 8, cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(status, h->d_status, (size_t)B * 4, cudaMemcpyDeviceToHost, s);
  }
  cudaError_t e = cudaStreamSynchronize(s);
  if (rc) return rc;
  if (e != cudaSuccess) return fail(h, CMPC_ERR_CUDA, cudaGetErrorString(e));
  return CMPC_OK;
}

int cmpc_stage_step_batch(cmpc_handle* h, int B, int mode, const double* state, const double* des_state, const double* des_inputs,
                          const double* hess, const double* rhs, double* d_fused, double* d_resolve, double* grad) {
  if (!h || !h->ready) return fail(h, CMPC_ERR_STATE, "cmpc_stage_step_batch: call cmpc_setup first");
  if (B < 0 || B > h->max_batch) return fail(h, CMPC_ERR_STATE, "batch exceeds max_batch given to cmpc_setup");
  if (!h->rip_used) return fail(h, CMPC_ERR_STATE, "stage-wise kernel not planned");
  if ((mode != 1 && mode != 2) || !state || !des_state || !des_inputs || !hess || !rhs || !d_fused || !d_resolve || !grad)
    return fail(h, CMPC_ERR_ARG, "cmpc_stage_step_batch: bad arguments");
  if (B == 0) return CMPC_OK;
  SET_DEVICE(h);
  const int N = h->cfg.horizon, L = h->cfg.num_legs;
  const size_t ns = 9 + 3 * L, nds = 9 * (N + 1), ndi = (size_t)L * (4 * N + 3), nf = (size_t)3 * L * N, nh = (size_t)6 * L * N;
  cudaStream_t s = h->stream;
  DevBuf tbuf;  // diagnostic path: a temporary buffer, released on every return path
  CUDA_TRY(h, tbuf.alloc((size_t)B * (nh + 4 * nf) * 8));
  double *tmp = tbuf.as<double>(), *dh = tmp, *dr = dh + (size_t)B * nh, *o1 = dr + (size_t)B * nf, *o2 = o1 + (size_t)B * nf, *o3 = o2 + (size_t)B * nf;
  cudaMemcpyAsync(h->d_state, state, B * ns * 8, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(h->d_ds, des_state, B * nds * 8, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(h->d_di, des_inputs, B * ndi * 8, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(dh, hess, (size_t)B * nh * 8, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(dr, rhs, (size_t)B * nf * 8, cudaMemcpyHostToDevice, s);
  SolveArgs a;
  std::memset(&a, 0, sizeof(a));
  a.state = h->d_state; a.des_state = h->d_ds; a.des_inputs = h->d_di;
  a.scratch = h->d_rip_scratch; a.scratch_per_group = h->rip_slab; a.groups = h->rip_groups;
  cudaError_t e = launch_ripm_probe(h->num_sms, 32 * h->rip_groups, h->rip_smem_bytes, s, h->dev, a, dh, dr, mode, o1, o2, o3, B);
  if (e == cudaSuccess) {
    cudaMemcpyAsync(d_fused, o1, (size_t)B * nf * 8, cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(d_resolve, o2, (size_t)B * nf * 8, cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(grad, o3, (size_t)B * nf * 8, cudaMemcpyDeviceToHost, s);
    e = cudaStreamSynchronize(s);
  }
  if (e != cudaSuccess) return fail(h, CMPC_ERR_CUDA, std::string("cmpc_stage_step_batch: ") + cudaGetErrorString(e));
  return CMPC_OK;
}

int cmpc_rollout(cmpc_handle* h, int B, int ticks, int warm_start, double* state, double* des_state,
                 double* des_inputs, double* force_log, int32_t* iters_sum, int32_t* status_or, cmpc_stats* stats) {
  if (!h || !h->ready) return fail(h, CMPC_ERR_STATE, "cmpc_rollout: call cmpc_setup first");
  if (B < 1 || B > h->max_batch || ticks < 1) return fail(h, CMPC_ERR_ARG, "cmpc_rollout: bad B or ticks");
  if (!state || !des_state || !des_inputs) return fail(h, CMPC_ERR_ARG, "null buffer");
  SET_DEVICE(h);
  const int N = h->cfg.horizon, L = h->cfg.num_legs;
  const size_t ns = 9 + 3 * L, nds = 9 * (N + 1), ndi = (size_t)L * (4 * N + 3);
  cudaStream_t s = h->stream;
  CUDA_TRY(h, cudaMemcpyAsync(h->d_state, state, B * ns * 8, cudaMemcpyHostToDevice, s));
  CUDA_TRY(h, cudaMemcpyAsync(h->d_ds, des_state, B * nds * 8, cudaMemcpyHostToDevice, s));
  CUDA_TRY(h, cudaMemcpyAsync(h->d_di, des_inputs, B * ndi * 8, cudaMemcpyHostToDevice, s));
  CUDA_TRY(h, cudaMemsetAsync(h->d_iters_sum, 0, (size_t)B * 4, s));
  CUDA_TRY(h, cudaMemsetAsync(h->d_status_or, 0, (size_t)B * 4, s));
  double* d_flog = nullptr;
  if (force_log) {  // the log buffer is kept across calls and only grows (no allocation on repeated roll-outs of the same size)
    const size_t need = (size_t)ticks * B * 3 * L;
    if (need > h->flog_cap) {
      cudaFree(h->d_flog); h->d_flog = nullptr; h->flog_cap = 0;
      CUDA_TRY(h, cudaMalloc(&h->d_flog, need * 8));
      h->flog_cap = need;
    }
    d_flog = h->d_flog;
  }
  SolveArgs a = SolveArgs();
  a.state = h->d_state; a.des_state = h->d_ds; a.des_inputs = h->d_di;
  a.forces = h->d_forces; a.status = h->d_status; a.iters = h->d_iters; a.kkt = h->d_kkt;
  a.active = h->d_active;
  if (warm_start) {  // previous tick's active set = the polish's first guess (tick 0: nothing active)
    a.warm_active = h->d_active;
    CUDA_TRY(h, cudaMemsetAsync(h->d_active, 0, (size_t)B * L * N * 2, s));
  }
  rollout_init_kernel<<<(B + 127) / 128, 128, 0, s>>>(h->dev, B, h->d_state, h->d_hip);
  CUDA_TRY(h, cudaEventRecord(h->ev[0], s));
  int rc = CMPC_OK;
  int launches = 0;
  for (int t = 0; t < ticks && rc == CMPC_OK; ++t) {
    rc = launch_solve(h, a, B);
    if (rc < 0) break;
    launches += rc + 1;
    rc = CMPC_OK;
    advance_kernel<<<(B + 127) / 128, 128, 0, s>>>(h->dev, B, h->d_state, h->d_ds, h->d_di, h->d_hip, h->d_forces, h->d_status,
                                                    d_flog ? d_flog + (size_t)t * B * 3 * L : nullptr, h->d_iters_sum,
                                                    h->d_iters, h->d_status_or);
  }
  cudaEventRecord(h->ev[1], s);
  if (rc == CMPC_OK) {
    cudaMemcpyAsync(state, h->d_state, B * ns * 8, cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(des_state, h->d_ds, B * nds * 8, cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(des_inputs, h->d_di, B * ndi * 8, cudaMemcpyDeviceToHost, s);
    if (force_log) cudaMemcpyAsync(force_log, d_flog, (size_t)ticks * B * 3 * L * 8, cudaMemcpyDeviceToHost, s);
    if (iters_sum) cudaMemcpyAsync(iters_sum, h->d_iters_sum, (size_t)B * 4, cudaMemcpyDeviceToHost, s);
    if (status_or) cudaMemcpyAsync(status_or, h->d_status_or, (size_t)B * 4, cudaMemcpyDeviceToHost, s);
  }
  cudaError_t e = cudaStreamSynchronize(s);
  if (rc) return rc;
  if (e != cudaSuccess) return fail(h, CMPC_ERR_CUDA, cudaGetErrorString(e));
  if (stats) {
    std::memset(stats, 0, sizeof(*stats));
    rc = collect_stats(h, B, h->d_status, h->d_iters, h->d_kkt, stats, launches);
    if (rc) return rc;
    float ms = 0;
    CUDA_TRY(h, cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
    stats->kernel_ms = ms;
  }
  return CMPC_OK;
}

int cmpc_foot_plan_batch(cmpc_handle* h, int B, const double* state, const double* des_inputs, double* foot_pos) {
  if (!h || !h->ready) return fail(h, CMPC_ERR_STATE, "cmpc_foot_plan_batch: call cmpc_setup first");
  if (B < 0 || B > h->max_batch) return fail(h, CMPC_ERR_STATE, "batch exceeds max_batch given to cmpc_setup");
  if (!state || !des_inputs || !foot_pos) return fail(h, CMPC_ERR_ARG, "null buffer");
  if (B == 0) return CMPC_OK;
  SET_DEVICE(h);
  const int N = h->cfg.horizon, L = h->cfg.num_legs;
  const size_t ns = 9 + 3 * L, ndi = (size_t)L * (4 * N + 3), nfp = (size_t)3 * L * (N + 1);
  cudaStream_t s = h->stream;
  // [B][L][N+1][3] <= [B][2][N][L][5] doubles of the multiplier buffer: reuse it as the output staging
  static_assert(3 * (CMPC_MAX_HORIZON + 1) <= 10 * CMPC_MAX_HORIZON, "foot plan fits the multiplier buffer");
  CUDA_TRY(h, cudaMemcpyAsync(h->d_state, state, B * ns * 8, cudaMemcpyHostToDevice, s));
  CUDA_TRY(h, cudaMemcpyAsync(h->d_di, des_inputs, B * ndi * 8, cudaMemcpyHostToDevice, s));
  const int total = B * L * 3;
  foot_plan_kernel<<<(total + 255) / 256, 256, 0, s>>>(h->dev, B, h->d_state, h->d_di, h->d_lam);
  CUDA_TRY(h, cudaGetLastError());
  CUDA_TRY(h, cudaMemcpyAsync(foot_pos, h->d_lam, B * nfp * 8, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(h, cudaStreamSynchronize(s));
  return CMPC_OK;
}

int cmpc_solve_batch_sqp(cmpc_handle* h, int B, int sqp_iters, const double* state, const double* des_state,
                         const double* des_inputs, double* forces, int32_t* status, double* defect) {
  if (!h || !h->ready) return fail(h, CMPC_ERR_STATE, "cmpc_solve_batch_sqp: call cmpc_setup first");
  if (B < 0 || B > h->max_batch) return fail(h, CMPC_ERR_STATE, "batch exceeds max_batch given to cmpc_setup");
  if (sqp_iters < 0 || sqp_iters > 16) return fail(h, CMPC_ERR_ARG, "sqp_iters must be 0..16");
  if (!state || !des_state || !des_inputs || !forces || !status) return fail(h, CMPC_ERR_ARG, "null buffer");
  if (B == 0) return CMPC_OK;
  SET_DEVICE(h);
  const int N = h->cfg.horizon, L = h->cfg.num_legs;
  const size_t ns = 9 + 3 * L, nds = 9 * (N + 1), ndi = (size_t)L * (4 * N + 3), nf = (size_t)3 * L * N;
  cudaStream_t s = h->stream;
  double* d_di0 = h->d_sqp;                              // [max_batch][ndi] original desired inputs, allocated by cmpc_setup
  double* d_def = h->d_sqp + (size_t)h->max_batch * ndi;  // [max_batch][17] defects
  cudaMemcpyAsync(h->d_state, state, B * ns * 8, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(h->d_ds, des_state, B * nds * 8, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(h->d_di, des_inputs, B * ndi * 8, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(d_di0, h->d_di, B * ndi * 8, cudaMemcpyDeviceToDevice, s);
  SolveArgs a = SolveArgs();
  a.state = h->d_state; a.des_state = h->d_ds; a.des_inputs = h->d_di;
  a.forces = h->d_forces; a.status = h->d_status; a.iters = h->d_iters; a.kkt = h->d_kkt;
  int rc = CMPC_OK;
  for (int it = 0; it <= sqp_iters && rc >= 0; ++it) {
    rc = launch_solve(h, a, B);
    if (rc < 0) break;
    relinearize_kernel<<<(B + 127) / 128, 128, 0, s>>>(h->dev, B, h->d_state, h->d_ds, d_di0, h->d_di, h->d_forces,
                                                       d_def + it, sqp_iters + 1, it < sqp_iters ? 1 : 0);
  }
  if (rc >= 0) {
    cudaMemcpyAsync(forces, h->d_forces, B * nf * 8, cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(status, h->d_status, (size_t)B * 4, cudaMemcpyDeviceToHost, s);
    if (defect) cudaMemcpyAsync(defect, d_def, (size_t)B * (sqp_iters + 1) * 8, cudaMemcpyDeviceToHost, s);
  }
  cudaError_t e = cudaStreamSynchronize(s);
  if (rc < 0) return rc;
  if (e != cudaSuccess) return fail(h, CMPC_ERR_CUDA, cudaGetErrorString(e));
  return CMPC_OK;
}

static int check_gaits(cmpc_handle* h, const cmpc_gait* gaits, int num_gaits, GaitTable* tab) {
  if (!gaits || num_gaits < 1 || num_gaits > kMaxGaits) return fail(h, CMPC_ERR_ARG, "gaits: 1..16 templates");
  tab->num_gaits = num_gaits;
  for (int q = 0; q < num_gaits; ++q) {
    const cmpc_gait& g = gaits[q];
    if (g.num_modes < 1 || g.num_modes > CMPC_MAX_GAIT_MODES) return fail(h, CMPC_ERR_ARG, "gait: num_modes out of range");
    for (int k = 0; k < g.num_modes; ++k) {
      if (g.modes[k] < 0 || g.modes[k] > 15) return fail(h, CMPC_ERR_ARG, "gait: mode numbers are 0..15");
      if (!(g.switching_times[k + 1] > g.switching_times[k])) return fail(h, CMPC_ERR_ARG, "gait: switching times must ascend");
    }
    tab->g[q] = g;
  }
  return CMPC_OK;
}

int cmpc_fill_contact_tables_device(cmpc_handle* h, int B, const cmpc_gait* gaits, int num_gaits, const int32_t* d_gait_id,
                                    const double* d_t0, double* d_des_inputs) {
  if (!h || !h->ready) return fail(h, CMPC_ERR_STATE, "cmpc_fill_contact_tables: call cmpc_setup first");
  if (B < 0 || !d_gait_id || !d_t0 || !d_des_inputs) return fail(h, CMPC_ERR_ARG, "null buffer");
  GaitTable tab;
  int rc = check_gaits(h, gaits, num_gaits, &tab);
  if (rc) return rc;
  if (B == 0) return CMPC_OK;
  SET_DEVICE(h);
  const int total = B * h->cfg.horizon;
  gait_kernel<<<(total + 255) / 256, 256, 0, h->stream>>>(h->dev, B, tab, d_gait_id, d_t0, d_des_inputs);
  CUDA_TRY(h, cudaGetLastError());
  return CMPC_OK;
}

int cmpc_fill_contact_tables(cmpc_handle* h, int B, const cmpc_gait* gaits, int num_gaits, const int32_t* gait_id,
                             const double* t0, double* des_inputs) {
  if (!h || !h->ready) return fail(h, CMPC_ERR_STATE, "cmpc_fill_contact_tables: call cmpc_setup first");
  if (B < 0 || B > h->max_batch) return fail(h, CMPC_ERR_STATE, "batch exceeds max_batch given to cmpc_setup");
  if (!gait_id || !t0 || !des_inputs) return fail(h, CMPC_ERR_ARG, "null buffer");
  if (B == 0) return CMPC_OK;
  SET_DEVICE(h);
  const size_t ndi = (size_t)h->cfg.num_legs * (4 * h->cfg.horizon + 3);
  cudaStream_t s = h->stream;
  // the handle's per-instance scratch arrays double as staging: status (int32) and kkt (double)
  CUDA_TRY(h, cudaMemcpyAsync(h->d_status, gait_id, (size_t)B * 4, cudaMemcpyHostToDevice, s));
  CUDA_TRY(h, cudaMemcpyAsync(h->d_kkt, t0, (size_t)B * 8, cudaMemcpyHostToDevice, s));
  CUDA_TRY(h, cudaMemcpyAsync(h->d_di, des_inputs, B * ndi * 8, cudaMemcpyHostToDevice, s));
  int rc = cmpc_fill_contact_tables_device(h, B, gaits, num_gaits, h->d_status, h->d_kkt, h->d_di);
  if (rc) return rc;
  CUDA_TRY(h, cudaMemcpyAsync(des_inputs, h->d_di, B * ndi * 8, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(h, cudaStreamSynchronize(s));
  return CMPC_OK;
}

int cmpc_fill_contact_tables_switch_device(cmpc_handle* h, int B, const cmpc_gait* gaits, int num_gaits, const int32_t* d_gait_from,
                                           const int32_t* d_gait_to, const double* d_t_tile, const double* d_t_switch, double stance_time,
                                           const double* d_t0, double* d_des_inputs) {
  if (!h || !h->ready) return fail(h, CMPC_ERR_STATE, "cmpc_fill_contact_tables_switch: call cmpc_setup first");
  if (B < 0 || !d_gait_from || !d_gait_to || !d_t_tile || !d_t_switch || !d_t0 || !d_des_inputs) return fail(h, CMPC_ERR_ARG, "null buffer");
  if (!(stance_time >= 0.0)) return fail(h, CMPC_ERR_ARG, "stance_time must be >= 0");
  GaitTable tab;
  int rc = check_gaits(h, gaits, num_gaits, &tab);
  if (rc) return rc;
  if (B == 0) return CMPC_OK;
  SET_DEVICE(h);
  gait_switch_kernel<<<(B + 127) / 128, 128, 0, h->stream>>>(h->dev, B, tab, d_gait_from, d_gait_to, d_t_tile, d_t_switch, stance_time, d_t0, d_des_inputs);
  CUDA_TRY(h, cudaGetLastError());
  return CMPC_OK;
}

int cmpc_fill_contact_tables_switch(cmpc_handle* h, int B, const cmpc_gait* gaits, int num_gaits, const int32_t* gait_from,
                                    const int32_t* gait_to, const double* t_tile, const double* t_switch, double stance_time,
                                    const double* t0, double* des_inputs) {
  if (!h || !h->ready) return fail(h, CMPC_ERR_STATE, "cmpc_fill_contact_tables_switch: call cmpc_setup first");
  if (B < 0 || B > h->max_batch) return fail(h, CMPC_ERR_STATE, "batch exceeds max_batch given to cmpc_setup");
  if (!gait_from || !gait_to || !t_tile || !t_switch || !t0 || !des_inputs) return fail(h, CMPC_ERR_ARG, "null buffer");
  if (B == 0) return CMPC_OK;
  SET_DEVICE(h);
  const size_t ndi = (size_t)h->cfg.num_legs * (4 * h->cfg.horizon + 3);
  cudaStream_t s = h->stream;
  // the handle's per-instance arrays double as staging: status / iters (int32), kkt and two rows of the force buffer (double)
  double* d_tt = h->d_forces; double* d_ts = h->d_forces + h->max_batch;
  if ((size_t)3 * h->cfg.num_legs * h->cfg.horizon < 2) return fail(h, CMPC_ERR_STATE, "staging too small");
  CUDA_TRY(h, cudaMemcpyAsync(h->d_status, gait_from, (size_t)B * 4, cudaMemcpyHostToDevice, s));
  CUDA_TRY(h, cudaMemcpyAsync(h->d_iters, gait_to, (size_t)B * 4, cudaMemcpyHostToDevice, s));
  CUDA_TRY(h, cudaMemcpyAsync(h->d_kkt, t0, (size_t)B * 8, cudaMemcpyHostToDevice, s));
  CUDA_TRY(h, cudaMemcpyAsync(d_tt, t_tile, (size_t)B * 8, cudaMemcpyHostToDevice, s));
  CUDA_TRY(h, cudaMemcpyAsync(d_ts, t_switch, (size_t)B * 8, cudaMemcpyHostToDevice, s));
  CUDA_TRY(h, cudaMemcpyAsync(h->d_di, des_inputs, B * ndi * 8, cudaMemcpyHostToDevice, s));
  int rc = cmpc_fill_contact_tables_switch_device(h, B, gaits, num_gaits, h->d_status, h->d_iters, d_tt, d_ts, stance_time, h->d_kkt, h->d_di);
  if (rc) return rc;
  CUDA_TRY(h, cudaMemcpyAsync(des_inputs, h->d_di, B * ndi * 8, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(h, cudaStreamSynchronize(s));
  return CMPC_OK;
}

int cmpc_measure_fp64_peak(cmpc_handle* h, double* tflops) {
  if (!h || !h->ready || !tflops) return fail(h, CMPC_ERR_STATE, "not set up");
  SET_DEVICE(h);
  const int blocks = h->num_sms * 8, threads = 256, iters = 4096;
  DevBuf dbuf;
  CUDA_TRY(h, dbuf.alloc((size_t)blocks * threads * 8));
  double* d = dbuf.as<double>();
  double best = 0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(h->ev[0], h->stream);
    fp64_peak_kernel<<<blocks, threads, 0, h->stream>>>(d, iters);
    cudaEventRecord(h->ev[1], h->stream);
    cudaError_t e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) return fail(h, CMPC_ERR_CUDA, cudaGetErrorString(e));
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
    const double flops = 2.0 * 64.0 * iters * (double)blocks * threads;
    if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  *tflops = best;
  return CMPC_OK;
}

void cmpc_destroy(cmpc_handle* h) {
  if (!h) return;
  release_device_state(h);
  delete h;
}

const char* cmpc_last_error(const cmpc_handle* h) { return h ? h->err.c_str() : "null handle"; }
const char* cmpc_last_route(const cmpc_handle* h) { return h ? h->route : "null handle"; }

}  // extern "C"
