// cmpc_device.cuh -- sm_100a device code of the batched centroidal-MPC condensed-QP solver.
//
// One *warp group* (W warps, W = 1, 4 or 8) per MPC instance; several groups per CTA, one
// CTA per SM, persistent over the batch with a device-side work counter.  Groups never use
// a CTA-wide barrier: W = 1 synchronises with __syncwarp and shuffles only, W > 1 with a
// named barrier per group.  Instances are bucketed by their number of free (stance-leg)
// variables so that a trot instance (n = 60 at horizon 10) gets a single warp and ~25 KB
// of shared memory, and 7-8 instances are resident per SM to hide the Cholesky's
// dependency chains behind one another.
//
// Per instance, everything between its 2.3 KB of inputs and 1 KB of outputs stays on chip
// or in an L2-resident scratch slab:
//   * build   : lever arms, A_d^p B_j closed forms ("power stacking" is index arithmetic
//               for the nilpotent centroidal A_c), H = 2(Bqp' L Bqp + K), g by an adjoint
//               sum -- SURVEY §8 a2-a7, reference CentroidalMPC.cpp:85-94,179-232,284-335
//   * solve   : feasible-start Mehrotra primal-dual interior point on the free (stance)
//               variables; per iteration one tiled Cholesky of H + C'SC in shared memory
//   * polish  : per-leg null-space active-set solve with verification/correction passes
// Matrices use the "BC4" layout: lower block triangle of 4x4 tiles, block-column major,
// so a tile is one 128-byte line and the tile Cholesky (POTRF/TRSM/GEMM on 4x4 tiles)
// works on whole tiles held in registers.
#pragma once
#include <utility>
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include <type_traits>

#include "../../include/cmpc.h"

namespace cmpc {

constexpr double kGrav = 9.81;      // CentroidalMPC.cpp:71
constexpr double kFricUb = 5000.0;  // CentroidalMPC.cpp:183
constexpr int kMaxLegs = CMPC_MAX_LEGS;
constexpr int kNumClasses = 4;

struct DevConfig {
  double mass, dt;
  double mu[kMaxLegs];
  double w[CMPC_NUM_WEIGHTS];
  double tol;
  int L, N, zoh, max_iter, polish;
};

struct PrePlan {  // presolve kernel, see make_pre_plan
  int T, cta;
  int x, ce, red, exch, ints, M;  // per-group offsets (doubles)
  int mat, total;
};

struct RicPlan {  // Riccati presolve kernel (cmpc_riccati.cu), see make_ric_plan
  int nz;                                  // 9 + 3L: augmented state [x; F_prev]
  int in, ce, P, p, t, Bb, T1, G, M, m0, xc, ints;  // per-group offsets (doubles)
  int X, F;                                // trajectory x_1..x_N and forces (alias P / T1..M when they fit)
  int total;
  int slab;                                // doubles of L2 scratch per group: N (12 nz + 12) gains
};

struct SmemPlan {
  // offsets in doubles from the start of the group's slab
  int ce, g, u, rd, tv, rhs, du, dua;
  int zl, zu, red, exch, ints, Mm;
  int total;  // doubles, multiple of 2
  int cta;    // doubles of CTA-shared tables in front of the groups: z1[N], z2[N] (build_qp)
};

// One launch = one size class.
struct SolveArgs {
  const double* state;
  const double* des_state;
  const double* des_inputs;
  double* forces;
  int32_t* status;
  int32_t* iters;
  double* kkt;
  double* lam;
  uint16_t* active;
  const uint16_t* warm_active;  // optional: previous tick's active set (closed loop), tried first
  double* Hout;  // build-export mode only
  double* gout;
  double* scratch;           // global (L2-resident) scratch: H per group, and M when it does not fit on chip
  size_t scratch_per_group;  // doubles
  const int32_t* perm;       // instance ids of this class (NULL: identity over [0, count))
  const int32_t* count;      // device pointer to the number of instances of this class (NULL: count_imm)
  int count_imm;
  int32_t* work;             // device work counter (zeroed before the launch)
  int nbmax;                 // most free blocks an instance of this class can have
  int n4max;                 // padded free dimension bound of the class
  int m_in_smem;
  int groups;                // groups per CTA
  SmemPlan plan;             // shared-memory layout of one group, computed on the host
  PrePlan pre;               // presolve kernel: its own shared-memory layout
  RicPlan ric;               // Riccati presolve kernel: its own shared-memory layout
  int32_t* fail_perm;        // presolve kernel only: instances it could not settle, for the IPM kernel
  int32_t* fail_count;
  // presolve kernel of size class 0 as the batch's router (no classify launch): it walks all B
  // instances in order and forwards those with more than nbmax free blocks to their class lists
  int route;                 // 0: perm / count describe this class only
  int route_b1, route_b2;    // largest nb of classes 1 and 2 (class 3 takes the rest)
  int32_t* route_perm;       // [class][route_stride]
  int32_t* route_counts;     // [class]
  int route_stride;
  // progressive input arrival (cmpc_solve_batch, host buffers): instance id / ready_chunk is its
  // copy chunk; *ready = number of chunks whose H2D copy has completed (written by the copy stream)
  const int32_t* ready;
  int ready_chunk;
  int32_t* error_flag;       // set when a wait on *ready times out
  // Several lists in one launch (stage-wise kernels: their code does not depend on the size class).  nlists > 0: the
  // persistent loop drains list 0, then list 1, ... (largest class first: its instances take longest, so they must not be
  // left for a partial wave at the end); perm / count / work / fail_* above are ignored.
  int phase_lock;  // stage-wise interior point: CTA barrier in front of every factor sweep (instruction-cache sharing)
  int pdl_trigger;  // the kernel lets its dependents be scheduled from its first instruction on (1: presolve kernels, 2: all); else as its CTAs retire
  int pdl;  // host side only: launch with programmatic stream serialisation (the kernel before it on the stream is one of ours)
  int nlists;
  const int32_t* lperm[kNumClasses];
  const int32_t* lcount[kNumClasses];
  int32_t* lwork[kNumClasses];
  int32_t* lfail_perm[kNumClasses];
  int32_t* lfail_count[kNumClasses];
  // condensed interior-point kernel split into phases (cmpc_solve.cu): iterate hand-over and lists
  double* xstate;            // [slot][xstride]: u (n4max) | zl (5 nbmax) | zu (5 nbmax) | it
  int xstride;
  int32_t* pol_perm;         // phase 1 -> phase 2: instances whose interior-point iterate is ready for the polish
  int32_t* pol_count;
  int32_t* fb_perm;          // phase 1 / 2 -> fused kernel: everything the split does not finish itself
  int32_t* fb_count;
  int32_t* zero_next;        // router presolve kernel only: the OTHER copy of the counts / work block, zeroed here for the next call
  int zero_n;                //  (the counts are double-buffered so that a call does not start with a memset node; cmpc_api.cu launch_solve)
  int32_t* hint_shadow;      // device copy of *hint_out (the host word is only written when the value changes)
  int32_t* hint_out;         // mapped host word: how many instances this launch found on its lists (next call's launch plan)
};

// ------------------------------------------------------------------ BC4 layout
// Tiles are 16 doubles padded to kTS = 18 (144 B): lanes that walk consecutive tiles with
// 128-bit accesses then fall into distinct 16-byte bank groups (9 t mod 8 = t mod 8), where
// the natural 128 B stride would put all 32 lanes on the same banks.
constexpr int kTS = 18;
__device__ __forceinline__ int blkoff(int bi, int bj, int nblk) {
  return bj * nblk - ((bj * (bj - 1)) >> 1) + (bi - bj);
}
__device__ __forceinline__ int midx(int i, int j, int nblk) {  // requires i>>2 >= j>>2
  return blkoff(i >> 2, j >> 2, nblk) * kTS + ((i & 3) << 2) + (j & 3);
}
__device__ __forceinline__ int sidx(int i, int j, int nblk) {  // symmetric access
  return ((i >> 2) >= (j >> 2)) ? midx(i, j, nblk) : midx(j, i, nblk);
}
__host__ __device__ inline int bc4_tiles(int n) {
  int nblk = (n + 3) >> 2;
  return (nblk * (nblk + 1)) >> 1;
}
__host__ __device__ inline int bc4_doubles(int n) { return bc4_tiles(n) * 18; }
// size of a matrix buffer: it also stages the raw inputs of an instance before the build
__host__ __device__ inline int mat_region_doubles(int N, int L, int n4max) {
  const int nin = (9 + 3 * L) + 9 * (N + 1) + L * (4 * N + 3);
  const int m = bc4_doubles(n4max);
  return ((m > nin ? m : nin) + 1) & ~1;
}

// ------------------------------------------------------------------ group primitives
// Programmatic dependent launch.  Every solver kernel starts with pdl_prologue(): it lets the next kernel on the stream
// be scheduled early (its CTAs become resident as this grid's CTAs retire instead of after a full drain) and waits until
// the kernel before it has completed and its writes are visible.  Both are no-ops for a launch without the attribute.
// A headline step launches three kernels that find their lists empty; back to back they cost 11 us of a 137 us step.
__device__ __forceinline__ void pdl_prologue(bool trigger = true) {
  if (trigger) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
inline cudaError_t launch_ex(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t stream, bool pdl, Args&&... args) {
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3((unsigned)grid); lc.blockDim = dim3((unsigned)block); lc.dynamicSmemBytes = smem; lc.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = at; lc.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&lc, kern, std::forward<Args>(args)...);
}
#endif

template <int W>
struct Group {
  static constexpr int GT = 32 * W;
  int gtid;     // thread index inside the group
  int gid;      // group index inside the CTA
  double* red;  // 3*W doubles of group-private shared scratch (W > 1 only)

  __device__ __forceinline__ void sync() const {
    if constexpr (W == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(gid + 1), "r"(GT) : "memory");
  }
  __device__ __forceinline__ bool all(bool p) const {
    if constexpr (W == 1) {
      __syncwarp();  // votes do not order memory; the callers rely on all() as a barrier
      return __all_sync(0xffffffffu, p);
    } else {
      unsigned r;
      asm volatile(
          "{ .reg .pred p, q; setp.ne.u32 q, %1, 0; bar.red.and.pred p, %2, %3, q; selp.u32 %0, 1, 0, p; }"
          : "=r"(r) : "r"((unsigned)p), "r"(gid + 1), "r"(GT) : "memory");
      return r != 0;
    }
  }
  __device__ __forceinline__ double max(double v) const {
    if constexpr (W == 1) __syncwarp();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if constexpr (W > 1) {
      sync();
      if ((gtid & 31) == 0) red[gtid >> 5] = v;
      sync();
      v = red[0];
#pragma unroll
      for (int k = 1; k < W; ++k) v = fmax(v, red[k]);
    }
    return v;
  }
  __device__ __forceinline__ double sum(double v) const {
    if constexpr (W == 1) __syncwarp();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if constexpr (W > 1) {
      sync();
      if ((gtid & 31) == 0) red[gtid >> 5] = v;
      sync();
      v = red[0];
#pragma unroll
      for (int k = 1; k < W; ++k) v += red[k];
    }
    return v;
  }
  __device__ __forceinline__ void max2_sum(double& a, double& b, double& s) const {
    if constexpr (W == 1) __syncwarp();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a = fmax(a, __shfl_xor_sync(0xffffffffu, a, o));
      b = fmax(b, __shfl_xor_sync(0xffffffffu, b, o));
      s += __shfl_xor_sync(0xffffffffu, s, o);
    }
    if constexpr (W > 1) {
      sync();
      if ((gtid & 31) == 0) { red[gtid >> 5] = a; red[W + (gtid >> 5)] = b; red[2 * W + (gtid >> 5)] = s; }
      sync();
      a = red[0]; b = red[W]; s = red[2 * W];
#pragma unroll
      for (int k = 1; k < W; ++k) { a = fmax(a, red[k]); b = fmax(b, red[W + k]); s += red[2 * W + k]; }
    }
  }
  // broadcast an int from thread 0 of the group
  __device__ __forceinline__ int bcast0(int v, int* slot) const {
    if constexpr (W == 1) {
      return __shfl_sync(0xffffffffu, v, 0);
    } else {
      sync();
      if (gtid == 0) *slot = v;
      sync();
      return *slot;
    }
  }
};

// ------------------------------------------------------------------ 4x4 tile kernels
__device__ __forceinline__ void ld_tile(const double* p, double* r) {
  const double2* p2 = reinterpret_cast<const double2*>(p);
#pragma unroll
  for (int q = 0; q < 8; ++q) { const double2 v = p2[q]; r[2 * q] = v.x; r[2 * q + 1] = v.y; }
}
__device__ __forceinline__ void st_tile(double* p, const double* r) {
  double2* p2 = reinterpret_cast<double2*>(p);
#pragma unroll
  for (int q = 0; q < 8; ++q) p2[q] = make_double2(r[2 * q], r[2 * q + 1]);
}
// 1/x to full double precision (not correctly rounded): MUFU seed + two Newton steps.
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}

// 1/sqrt(x) to full double precision (not correctly rounded): MUFU seed + two Newton steps.
__device__ __forceinline__ double fast_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x * y, y, 1.0);
  y = fma(0.5 * y, e, y);
  e = fma(-x * y, y, 1.0);
  return fma(0.5 * y, e, y);
}

// Cholesky of a 4x4 SPD tile (row-major, lower part read) in "solve form" d[16]:
//   strict lower  : L
//   diagonal      : 1 / l_ii
//   strict upper  : strict lower of L^-1, transposed  (d[4c + r] = Linv(r, c), r > c... see below)
// so that  L^-1 b  and  L^-T b  are four independent short dot products (no substitution
// chain) in the TRSM and in both triangular solves.  d[4*r + c] for r < c holds Linv(c, r).
// Returns false if a pivot is not positive.
__device__ __forceinline__ bool potrf4(const double* a, double* d) {
  const double a00 = a[0], a10 = a[4], a11 = a[5], a20 = a[8], a21 = a[9], a22 = a[10];
  const double a30 = a[12], a31 = a[13], a32 = a[14], a33 = a[15];
  bool ok = a00 > 0.0;
  const double i0 = fast_rsqrt(a00);
  const double l10 = a10 * i0, l20 = a20 * i0, l30 = a30 * i0;
  const double d1 = a11 - l10 * l10;
  ok = ok && d1 > 0.0;
  const double i1 = fast_rsqrt(d1);
  const double l21 = (a21 - l20 * l10) * i1, l31 = (a31 - l30 * l10) * i1;
  const double d2 = a22 - l20 * l20 - l21 * l21;
  ok = ok && d2 > 0.0;
  const double i2 = fast_rsqrt(d2);
  const double l32 = (a32 - l30 * l20 - l31 * l21) * i2;
  const double d3 = a33 - l30 * l30 - l31 * l31 - l32 * l32;
  ok = ok && d3 > 0.0;
  const double i3 = fast_rsqrt(d3);
  // inverse of the unit... of L: Linv(r,c), r > c
  const double v10 = -l10 * i0 * i1;
  const double v21 = -l21 * i1 * i2;
  const double v32 = -l32 * i2 * i3;
  const double v20 = -(l20 * i0 + l21 * v10) * i2;
  const double v31 = -(l31 * i1 + l32 * v21) * i3;
  const double v30 = -(l30 * i0 + l31 * v10 + l32 * v20) * i3;
  d[0] = i0;   d[1] = v10;  d[2] = v20;  d[3] = v30;
  d[4] = l10;  d[5] = i1;   d[6] = v21;  d[7] = v31;
  d[8] = l20;  d[9] = l21;  d[10] = i2;  d[11] = v32;
  d[12] = l30; d[13] = l31; d[14] = l32; d[15] = i3;
  return ok;
}
// y = L^-1 b for a solve-form diagonal tile d
__device__ __forceinline__ void linv4(const double* d, double b0, double b1, double b2, double b3,
                                      double& y0, double& y1, double& y2, double& y3) {
  y0 = d[0] * b0;
  y1 = d[1] * b0 + d[5] * b1;
  y2 = d[2] * b0 + d[6] * b1 + d[10] * b2;
  y3 = d[3] * b0 + d[7] * b1 + d[11] * b2 + d[15] * b3;
}
// x = L^-T y
__device__ __forceinline__ void linvt4(const double* d, double y0, double y1, double y2, double y3,
                                       double& x0, double& x1, double& x2, double& x3) {
  x3 = d[15] * y3;
  x2 = d[10] * y2 + d[11] * y3;
  x1 = d[5] * y1 + d[6] * y2 + d[7] * y3;
  x0 = d[0] * y0 + d[1] * y1 + d[2] * y2 + d[3] * y3;
}

// Rows of a vector are owned by threads: row r = gtid + s * GT, s = 0, 1 (n4 <= 2 GT).
// Fetch the four pivot values x[4kb .. 4kb+3] from their owners' registers.
template <int W>
__device__ __forceinline__ void pivot4(const Group<W>& G, const double (&xr)[2], int kb, double* exch,
                                       double& b0, double& b1, double& b2, double& b3) {
  constexpr int GT = Group<W>::GT;
  const int r0 = 4 * kb;
  if constexpr (W == 1) {  // one warp: straight out of the owners' registers
    const double v = (r0 >= 32) ? xr[1] : xr[0];
    const int l0 = r0 & 31;
    b0 = __shfl_sync(0xffffffffu, v, l0); b1 = __shfl_sync(0xffffffffu, v, l0 + 1);
    b2 = __shfl_sync(0xffffffffu, v, l0 + 2); b3 = __shfl_sync(0xffffffffu, v, l0 + 3);
    return;
  }
  double* buf = exch + ((kb & 1) << 2);  // double-buffered: one barrier per step
  const int q0 = G.gtid - (r0 % GT);
  if (q0 >= 0 && q0 < 4) buf[q0] = (r0 >= GT) ? xr[1] : xr[0];
  G.sync();
  const double2* b2p = reinterpret_cast<const double2*>(buf);
  const double2 u = b2p[0], v = b2p[1];
  b0 = u.x; b1 = u.y; b2 = v.x; b3 = v.y;
}

// (For one- and two-warp groups chol / substitutions / symv / copies are called out of line: the interior-
// point kernel uses each from several places, and inlining them all gave 41 k SASS instructions (660 KB);
// when the resident warps are in different phases of different instances the instruction cache thrashes --
// 8.4 stall cycles per issued instruction waiting for instructions on the tracking-heavy workload, 5.7 with
// the calls.  The 4- and 8-warp variants keep them inline: out of line they spill.)
// D(8x8) += A(8x4) B(4x8) on the FP64 tensor-core path.  Lane l holds A[l >> 2][l & 3], B[l & 3][l >> 2] and
// D[l >> 2][2 (l & 3) + {0, 1}].
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// Tensor-core update of NI vertically stacked 8-row blocks (tile rows r0 + 2i + {0, 1}, r0 = c0) of the tile columns c0
// (and c0 + 1 when `two`) in BC4 layout:  C -= sum_{kk in [k0, k1)} L(rows, kk) L(c0.., kk)'.  The blocks stay in the
// accumulator registers over all k1 - k0 mma steps and are read and written once (see chol_cm_mma in cmpc_presolve.cu for
// the chunk-major twin).  Fragment element of lane l: row (l >> 2) & 3, column l & 3 of tile t + (l >> 4) -- two
// wavefronts per load, no bank conflict at kTS = 18.  A tile row past the end reads the next column's first tile (inside
// the matrix) and only feeds output rows / columns that are not stored.
template <int NI>
__device__ __forceinline__ void mma_update_bc4(double* M, int nblk, int lane, bool two, int c0, int k0, int k1) {
  const int fa = (lane >> 2) & 3, fhi = lane >> 4, ctc = (lane & 3) >> 1;
  const int tcol = c0 + ctc;
  double* cbase = M + (tcol * nblk - ((tcol * (tcol - 1)) >> 1) - tcol) * kTS + 4 * fa + 2 * (lane & 1);  // + kTS (tile row)
  const bool colok = two ? tcol < nblk : ctc == 0;
  double acc[NI][2];
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int trow = c0 + 2 * i + fhi;
    acc[i][0] = 0.0; acc[i][1] = 0.0;
    if (colok && trow < nblk && trow >= tcol) { const double2 c = *reinterpret_cast<const double2*>(cbase + trow * kTS); acc[i][0] = c.x; acc[i][1] = c.y; }
  }
  const double* pb = M + fhi * kTS + 4 * fa + (lane & 3) + (k0 * nblk - ((k0 * (k0 - 1)) >> 1) - k0) * kTS;  // + kTS (tile row): column k0
#pragma unroll 1
  for (int kk = k0; kk < k1; ++kk) {
    const double b = pb[kTS * c0];
    const double bn = -b;
    double a[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) a[i] = (i == 0) ? b : pb[kTS * (c0 + 2 * i)];
#pragma unroll
    for (int i = 0; i < NI; ++i) dmma884(acc[i][0], acc[i][1], a[i], bn);
    pb += kTS * (nblk - kk - 1);
  }
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int trow = c0 + 2 * i + fhi;
    if (colok && trow < nblk && trow >= tcol) *reinterpret_cast<double2*>(cbase + trow * kTS) = make_double2(acc[i][0], acc[i][1]);
  }
}
__device__ __forceinline__ void mma_update_bc4_n(int ni, double* M, int nblk, int lane, bool two, int c0, int k0, int k1) {
  switch (ni) {
    case 1: mma_update_bc4<1>(M, nblk, lane, two, c0, k0, k1); break;
    case 2: mma_update_bc4<2>(M, nblk, lane, two, c0, k0, k1); break;
    case 3: mma_update_bc4<3>(M, nblk, lane, two, c0, k0, k1); break;
    case 4: mma_update_bc4<4>(M, nblk, lane, two, c0, k0, k1); break;
    case 5: mma_update_bc4<5>(M, nblk, lane, two, c0, k0, k1); break;
    case 6: mma_update_bc4<6>(M, nblk, lane, two, c0, k0, k1); break;
    case 7: mma_update_bc4<7>(M, nblk, lane, two, c0, k0, k1); break;
    case 8: mma_update_bc4<8>(M, nblk, lane, two, c0, k0, k1); break;
    default: break;
  }
}

// Tiled right-looking Cholesky in BC4 layout, in place; the whole group calls it.
// tb[t] = bi | bj << 8 for storage tile t.  Diagonal tiles end up in solve form (potrf4).
// If rhs != nullptr the forward substitution  y = L^-1 rhs  is fused into the sweep (rows
// live in registers, the update of step kb rides along with the trailing update) and y
// overwrites rhs.  Returns false (uniformly) on a non-positive pivot.
template <int W>
__device__ __forceinline__ bool chol_bc4_impl(const Group<W> G, double* M, int nblk, const uint16_t* tb, double* rhs, double* exch) {
  constexpr int GT = Group<W>::GT;
  const int gtid = G.gtid;
  const int ntiles = (nblk * (nblk + 1)) >> 1;
  const int n4 = nblk << 2;
  double xr[2] = {0.0, 0.0};
  if (rhs) {
    if (gtid < n4) xr[0] = rhs[gtid];
    if (gtid + GT < n4) xr[1] = rhs[gtid + GT];
  }
  bool ok = true;
  // one warp, n4 <= 64: left-looking over 8-wide block columns, rank-k part on the FP64 tensor cores (mma_update_bc4)
  const bool left = W == 1 && nblk <= 16;
  for (int kb = 0; kb < nblk; ++kb) {
    const int col0 = blkoff(kb, kb, nblk);  // storage index of the diagonal tile of column kb
    const int nrows = nblk - kb;
    if (W == 1 && left && kb > 0) {
      if (kb & 1) mma_update_bc4_n((nrows + 1) >> 1, M, nblk, gtid, false, kb, kb - 1, kb);
      else mma_update_bc4_n((nrows + 1) >> 1, M, nblk, gtid, true, kb, 0, kb);
      G.sync();
    }
    double d[16], a[16];
    ld_tile(M + col0 * kTS, a);             // broadcast loads: every lane factors the same tile
    ok = potrf4(a, d) && ok;
    // TRSM: X = A L^-T, one panel ROW per thread (rows 4(kb+1) .. n4-1)
    for (int r = 4 * (kb + 1) + gtid; r < n4; r += GT) {
      double2* A2 = reinterpret_cast<double2*>(M + (col0 + (r >> 2) - kb) * kTS + ((r & 3) << 2));
      const double2 u = A2[0], v = A2[1];
      double x0, x1, x2, x3;
      linv4(d, u.x, u.y, v.x, v.y, x0, x1, x2, x3);
      A2[0] = make_double2(x0, x1); A2[1] = make_double2(x2, x3);
    }
    double y0 = 0, y1 = 0, y2 = 0, y3 = 0;
    if (rhs) {
      double b0, b1, b2, b3;
      pivot4<W>(G, xr, kb, exch, b0, b1, b2, b3);
      linv4(d, b0, b1, b2, b3, y0, y1, y2, y3);
      if (gtid == 0) {  // y of this step goes straight to the output vector (nobody reads rhs[] during the sweep)
        double2* o2 = reinterpret_cast<double2*>(rhs + 4 * kb);
        o2[0] = make_double2(y0, y1); o2[1] = make_double2(y2, y3);
      }
    }
    ok = G.all(ok);  // also the barrier between the panel and the trailing update
    if (!ok) return false;
    if (gtid == 0) st_tile(M + col0 * kTS, d);
    if (rhs) {
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int r = gtid + s * GT;
        if (r >= 4 * (kb + 1) && r < n4) {
          const double2* A2 = reinterpret_cast<const double2*>(M + (col0 + (r >> 2) - kb) * kTS + ((r & 3) << 2));
          const double2 u = A2[0], v = A2[1];
          xr[s] -= u.x * y0 + u.y * y1 + v.x * y2 + v.y * y3;
        }
      }
    }
    // trailing update: storage tiles of columns kb+1.. are contiguous
    const int t0 = col0 + nrows;
    for (int t = t0 + gtid; t < (left ? 0 : ntiles); t += GT) {
      const int bi = tb[t] & 0xff, bj = tb[t] >> 8;
      double li[16], lj[16], c[16];
      double* C = M + t * kTS;
      ld_tile(M + (col0 + bi - kb) * kTS, li);
      ld_tile(M + (col0 + bj - kb) * kTS, lj);
      ld_tile(C, c);
#pragma unroll
      for (int rr = 0; rr < 4; ++rr)
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          double sacc = c[4 * rr + cc];
#pragma unroll
          for (int k = 0; k < 4; ++k) sacc -= li[4 * rr + k] * lj[4 * cc + k];
          c[4 * rr + cc] = sacc;
        }
      st_tile(C, c);
    }
    G.sync();
  }
  return true;
}
template <int W>
__device__ __noinline__ bool chol_bc4_call(const Group<W> G, double* M, int nblk, const uint16_t* tb, double* rhs, double* exch) { return chol_bc4_impl<W>(G, M, nblk, tb, rhs, exch); }
template <int W>
__device__ __forceinline__ bool chol_bc4(const Group<W> G, double* M, int nblk, const uint16_t* tb, double* rhs, double* exch) {
  if constexpr (W <= 2) return chol_bc4_call<W>(G, M, nblk, tb, rhs, exch); else return chol_bc4_impl<W>(G, M, nblk, tb, rhs, exch);
}

// Forward substitution y = L^-1 x (in place in shared memory), rows in registers.
template <int W>
__device__ __forceinline__ void chol_fwd_bc4_impl(const Group<W> G, const double* M, int nblk, double* x, double* exch) {
  constexpr int GT = Group<W>::GT;
  const int gtid = G.gtid, n4 = nblk << 2;
  double xr[2] = {0.0, 0.0};
  if (gtid < n4) xr[0] = x[gtid];
  if (gtid + GT < n4) xr[1] = x[gtid + GT];
  G.sync();  // everybody holds its rows before step results overwrite x[]
  for (int kb = 0; kb < nblk; ++kb) {
    const int col0 = blkoff(kb, kb, nblk);
    double d[16], b0, b1, b2, b3, y0, y1, y2, y3;
    ld_tile(M + col0 * kTS, d);
    pivot4<W>(G, xr, kb, exch, b0, b1, b2, b3);
    linv4(d, b0, b1, b2, b3, y0, y1, y2, y3);
    if (gtid == 0) {
      double2* o2 = reinterpret_cast<double2*>(x + 4 * kb);
      o2[0] = make_double2(y0, y1); o2[1] = make_double2(y2, y3);
    }
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int r = gtid + s * GT;
      if (r >= 4 * (kb + 1) && r < n4) {
        const double2* A2 = reinterpret_cast<const double2*>(M + (col0 + (r >> 2) - kb) * kTS + ((r & 3) << 2));
        const double2 u = A2[0], v = A2[1];
        xr[s] -= u.x * y0 + u.y * y1 + v.x * y2 + v.y * y3;
      }
    }
  }
  G.sync();
}
template <int W>
__device__ __noinline__ void chol_fwd_bc4_call(const Group<W> G, const double* M, int nblk, double* x, double* exch) { chol_fwd_bc4_impl<W>(G, M, nblk, x, exch); }
template <int W>
__device__ __forceinline__ void chol_fwd_bc4(const Group<W> G, const double* M, int nblk, double* x, double* exch) {
  if constexpr (W <= 2) chol_fwd_bc4_call<W>(G, M, nblk, x, exch); else chol_fwd_bc4_impl<W>(G, M, nblk, x, exch);
}

// Backward substitution x = L^-T y (in place in shared memory), rows in registers.
template <int W>
__device__ __forceinline__ void chol_bwd_bc4_impl(const Group<W> G, const double* M, int nblk, double* x, double* exch) {
  constexpr int GT = Group<W>::GT;
  const int gtid = G.gtid, n4 = nblk << 2;
  double xr[2] = {0.0, 0.0};
  if (gtid < n4) xr[0] = x[gtid];
  if (gtid + GT < n4) xr[1] = x[gtid + GT];
  G.sync();
  for (int kb = nblk - 1; kb >= 0; --kb) {
    double d[16], b0, b1, b2, b3, x0, x1, x2, x3;
    ld_tile(M + blkoff(kb, kb, nblk) * kTS, d);
    pivot4<W>(G, xr, kb, exch, b0, b1, b2, b3);
    linvt4(d, b0, b1, b2, b3, x0, x1, x2, x3);
    if (gtid == 0) {
      double2* o2 = reinterpret_cast<double2*>(x + 4 * kb);
      o2[0] = make_double2(x0, x1); o2[1] = make_double2(x2, x3);
    }
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int r = gtid + s * GT;
      if (r < 4 * kb) {
        const double* A = M + blkoff(kb, r >> 2, nblk) * kTS + (r & 3);  // column r&3 of L(kb, r>>2)
        xr[s] -= A[0] * x0 + A[4] * x1 + A[8] * x2 + A[12] * x3;
      }
    }
  }
  G.sync();
}
template <int W>
__device__ __noinline__ void chol_bwd_bc4_call(const Group<W> G, const double* M, int nblk, double* x, double* exch) { chol_bwd_bc4_impl<W>(G, M, nblk, x, exch); }
template <int W>
__device__ __forceinline__ void chol_bwd_bc4(const Group<W> G, const double* M, int nblk, double* x, double* exch) {
  if constexpr (W <= 2) chol_bwd_bc4_call<W>(G, M, nblk, x, exch); else chol_bwd_bc4_impl<W>(G, M, nblk, x, exch);
}

// y = H x for symmetric H in BC4 layout (diagonal tiles hold both triangles), one row per
// thread (two passes when n4 > GT).  The whole group calls it; ends with a group barrier.
template <int W>
__device__ __forceinline__ void symv_bc4_impl(const Group<W> G, const double* H, int n4, int nblk, const double* x, double* y) {
  constexpr int GT = Group<W>::GT;
  for (int row = G.gtid; row < n4; row += GT) {
    const int bi = row >> 2, ri = row & 3;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    // tiles (bi, bj), bj <= bi : row ri
    for (int bj = 0; bj <= bi; ++bj) {
      const double2* A2 = reinterpret_cast<const double2*>(H + blkoff(bi, bj, nblk) * kTS + (ri << 2));
      const double2* x2 = reinterpret_cast<const double2*>(x + (bj << 2));
      const double2 u = A2[0], v = A2[1], xa = x2[0], xb = x2[1];
      s0 = fma(u.x, xa.x, s0); s1 = fma(u.y, xa.y, s1);
      s2 = fma(v.x, xb.x, s2); s3 = fma(v.y, xb.y, s3);
    }
    // tiles (bj, bi), bj > bi : column ri; consecutive in storage
    const double* A = H + (blkoff(bi, bi, nblk) + 1) * kTS + ri;
    for (int bj = bi + 1; bj < nblk; ++bj, A += kTS) {
      const double2* x2 = reinterpret_cast<const double2*>(x + (bj << 2));
      const double2 xa = x2[0], xb = x2[1];
      s0 = fma(A[0], xa.x, s0); s1 = fma(A[4], xa.y, s1);
      s2 = fma(A[8], xb.x, s2); s3 = fma(A[12], xb.y, s3);
    }
    y[row] = (s0 + s1) + (s2 + s3);
  }
  G.sync();
}
template <int W>
__device__ __noinline__ void symv_bc4_call(const Group<W> G, const double* H, int n4, int nblk, const double* x, double* y) { symv_bc4_impl<W>(G, H, n4, nblk, x, y); }
template <int W>
__device__ __forceinline__ void symv_bc4(const Group<W> G, const double* H, int n4, int nblk, const double* x, double* y) {
  if constexpr (W <= 2) symv_bc4_call<W>(G, H, n4, nblk, x, y); else symv_bc4_impl<W>(G, H, n4, nblk, x, y);
}

// Copy a BC4 matrix, 16 bytes per access (global scratch <-> shared).
template <int W>
__device__ __forceinline__ void copy_mat_impl(const Group<W> G, double* dst, const double* src, int ndoubles) {
  const double2* s2 = reinterpret_cast<const double2*>(src);
  double2* d2 = reinterpret_cast<double2*>(dst);
  if (__isShared(dst)) {
    // global -> shared without a register round trip: every 16-byte piece of the thread is in flight at once (as plain
    // loads the compiler batches four at a time: eight dependent trips to L2 for a 15 KB matrix, every iteration)
    const unsigned da = (unsigned)__cvta_generic_to_shared(d2);
    for (int t = G.gtid; t < (ndoubles >> 1); t += Group<W>::GT)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(da + 16u * (unsigned)t), "l"(s2 + t) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    return;
  }
  for (int t = G.gtid; t < (ndoubles >> 1); t += Group<W>::GT) d2[t] = __ldcg(s2 + t);
}
template <int W>
__device__ __noinline__ void copy_mat_call(const Group<W> G, double* dst, const double* src, int ndoubles) { copy_mat_impl<W>(G, dst, src, ndoubles); }
template <int W>
__device__ __forceinline__ void copy_mat(const Group<W> G, double* dst, const double* src, int ndoubles) {
  if constexpr (W <= 2) copy_mat_call<W>(G, dst, src, ndoubles); else copy_mat_impl<W>(G, dst, src, ndoubles);
}
template <int W>
__device__ __forceinline__ void store_mat_impl(const Group<W> G, double* dst, const double* src, int ndoubles) {
  const double2* s2 = reinterpret_cast<const double2*>(src);
  double2* d2 = reinterpret_cast<double2*>(dst);
  for (int t = G.gtid; t < (ndoubles >> 1); t += Group<W>::GT) __stcg(d2 + t, s2[t]);
}
template <int W>
__device__ __noinline__ void store_mat_call(const Group<W> G, double* dst, const double* src, int ndoubles) { store_mat_impl<W>(G, dst, src, ndoubles); }
template <int W>
__device__ __forceinline__ void store_mat(const Group<W> G, double* dst, const double* src, int ndoubles) {
  if constexpr (W <= 2) store_mat_call<W>(G, dst, src, ndoubles); else store_mat_impl<W>(G, dst, src, ndoubles);
}

// ------------------------------------------------------------------ friction pyramid rows
// y0 = mu fz - fx, y1 = mu fz + fx, y2 = mu fz - fy, y3 = mu fz + fy, y4 = fz
// (F_i of CentroidalMPC.cpp:186-190);  0 <= y <= ub * c  (:199)
__device__ __forceinline__ void row_vec(double mu, int r, double* a) {
  a[0] = (r == 0) ? -1.0 : (r == 1 ? 1.0 : 0.0);
  a[1] = (r == 2) ? -1.0 : (r == 3 ? 1.0 : 0.0);
  a[2] = (r == 4) ? 1.0 : mu;
}
__device__ __forceinline__ void cmul5(double mu, const double* f, double* y) {
  double mf = mu * f[2];
  y[0] = mf - f[0]; y[1] = mf + f[0]; y[2] = mf - f[1]; y[3] = mf + f[1]; y[4] = f[2];
}
__device__ __forceinline__ void ctmul5(double mu, const double* w, double* o) {
  o[0] = w[1] - w[0]; o[1] = w[3] - w[2]; o[2] = mu * (w[0] + w[1] + w[2] + w[3]) + w[4];
}

// Null space of the active rows of one leg-step block (Gram-Schmidt). Returns rank.
static __device__ __noinline__ int block_nullspace(int k, const double (*A)[3], const double* b, double* f0, double (*Z)[3], bool* ok) {
  double Q[3][3];
  int r = 0;
  *ok = true;
  f0[0] = f0[1] = f0[2] = 0.0;
  for (int t = 0; t < k; ++t) {
    double v0 = A[t][0], v1 = A[t][1], v2 = A[t][2];
    double na = sqrt(v0 * v0 + v1 * v1 + v2 * v2);
    for (int s = 0; s < r; ++s) {
      double d = Q[s][0] * A[t][0] + Q[s][1] * A[t][1] + Q[s][2] * A[t][2];
      v0 -= d * Q[s][0]; v1 -= d * Q[s][1]; v2 -= d * Q[s][2];
    }
    double nv = sqrt(v0 * v0 + v1 * v1 + v2 * v2);
    double af0 = A[t][0] * f0[0] + A[t][1] * f0[1] + A[t][2] * f0[2];
    if (r < 3 && nv > 1e-10 * na) {
      Q[r][0] = v0 / nv; Q[r][1] = v1 / nv; Q[r][2] = v2 / nv;
      double aq = A[t][0] * Q[r][0] + A[t][1] * Q[r][1] + A[t][2] * Q[r][2];
      double st = (b[t] - af0) / aq;
      f0[0] += st * Q[r][0]; f0[1] += st * Q[r][1]; f0[2] += st * Q[r][2];
      ++r;
    } else if (fabs(af0 - b[t]) > 1e-9 * (1.0 + fabs(b[t]))) {
      *ok = false;
    }
  }
  if (r == 0) {
    for (int a = 0; a < 3; ++a)
      for (int c = 0; c < 3; ++c) Z[a][c] = (a == c) ? 1.0 : 0.0;
  } else if (r == 1) {
    int m = 0;
    if (fabs(Q[0][1]) < fabs(Q[0][m])) m = 1;
    if (fabs(Q[0][2]) < fabs(Q[0][m])) m = 2;
    double e0 = (m == 0), e1 = (m == 1), e2 = (m == 2);
    double z0 = Q[0][1] * e2 - Q[0][2] * e1, z1 = Q[0][2] * e0 - Q[0][0] * e2, z2 = Q[0][0] * e1 - Q[0][1] * e0;
    double n1 = sqrt(z0 * z0 + z1 * z1 + z2 * z2);
    z0 /= n1; z1 /= n1; z2 /= n1;
    Z[0][0] = z0; Z[0][1] = z1; Z[0][2] = z2;
    Z[1][0] = Q[0][1] * z2 - Q[0][2] * z1; Z[1][1] = Q[0][2] * z0 - Q[0][0] * z2; Z[1][2] = Q[0][0] * z1 - Q[0][1] * z0;
  } else if (r == 2) {
    double z0 = Q[0][1] * Q[1][2] - Q[0][2] * Q[1][1], z1 = Q[0][2] * Q[1][0] - Q[0][0] * Q[1][2],
           z2 = Q[0][0] * Q[1][1] - Q[0][1] * Q[1][0];
    double n1 = sqrt(z0 * z0 + z1 * z1 + z2 * z2);
    Z[0][0] = z0 / n1; Z[0][1] = z1 / n1; Z[0][2] = z2 / n1;
  }
  return r;
}

// least squares S' lam = rb for k (<=3) independent normals via normal equations
static __device__ __noinline__ double small_lsq(int k, const double (*S)[3], const double* rb, double* lam) {
  double G[3][3], y[3];
  for (int a = 0; a < k; ++a) {
    y[a] = S[a][0] * rb[0] + S[a][1] * rb[1] + S[a][2] * rb[2];
    for (int b = 0; b < k; ++b) G[a][b] = S[a][0] * S[b][0] + S[a][1] * S[b][1] + S[a][2] * S[b][2];
  }
  int piv[3] = {0, 1, 2};
  for (int c = 0; c < k; ++c) {
    int m = c;
    for (int r = c + 1; r < k; ++r)
      if (fabs(G[piv[r]][c]) > fabs(G[piv[m]][c])) m = r;
    int t = piv[c]; piv[c] = piv[m]; piv[m] = t;
    double d = G[piv[c]][c];
    if (fabs(d) < 1e-300) return INFINITY;
    for (int r = c + 1; r < k; ++r) {
      double f = G[piv[r]][c] / d;
      for (int cc = c; cc < k; ++cc) G[piv[r]][cc] -= f * G[piv[c]][cc];
      y[piv[r]] -= f * y[piv[c]];
    }
  }
  for (int c = k - 1; c >= 0; --c) {
    double s = y[piv[c]];
    for (int cc = c + 1; cc < k; ++cc) s -= G[piv[c]][cc] * lam[cc];
    lam[c] = s / G[piv[c]][c];
  }
  double res = 0.0;
  for (int a = 0; a < 3; ++a) {
    double s = -rb[a];
    for (int t = 0; t < k; ++t) s += S[t][a] * lam[t];
    res = fmax(res, fabs(s));
  }
  return res;
}

// lam >= 0 with sum lam_t Nrm_t = rb; enumerates independent subsets (degenerate apex).
static __device__ __noinline__ bool block_multipliers(int k, const double (*Nrm)[3], const double* rb, double tol, double* lam) {
  for (int t = 0; t < k; ++t) lam[t] = 0.0;
  if (k == 0) return fmax(fabs(rb[0]), fmax(fabs(rb[1]), fabs(rb[2]))) <= tol;
  double Q[3][3];
  int rank = 0;
  for (int t = 0; t < k && rank < 3; ++t) {
    double v0 = Nrm[t][0], v1 = Nrm[t][1], v2 = Nrm[t][2];
    double na = sqrt(v0 * v0 + v1 * v1 + v2 * v2);
    for (int s = 0; s < rank; ++s) {
      double d = Q[s][0] * Nrm[t][0] + Q[s][1] * Nrm[t][1] + Q[s][2] * Nrm[t][2];
      v0 -= d * Q[s][0]; v1 -= d * Q[s][1]; v2 -= d * Q[s][2];
    }
    double nv = sqrt(v0 * v0 + v1 * v1 + v2 * v2);
    if (nv > 1e-10 * na) { Q[rank][0] = v0 / nv; Q[rank][1] = v1 / nv; Q[rank][2] = v2 / nv; ++rank; }
  }
  bool have_first = false;
  for (int mask = 1; mask < (1 << k); ++mask) {
    if (__popc(mask) != rank) continue;
    double S[3][3], ls[3];
    int idx[3], c = 0;
    for (int t = 0; t < k; ++t)
      if ((mask >> t) & 1) { S[c][0] = Nrm[t][0]; S[c][1] = Nrm[t][1]; S[c][2] = Nrm[t][2]; idx[c++] = t; }
    double res = small_lsq(rank, S, rb, ls);
    if (!isfinite(res)) continue;
    bool okk = res <= tol;
    for (int t = 0; t < rank; ++t)
      if (!(ls[t] >= -tol)) okk = false;
    if (okk || !have_first) {
      for (int t = 0; t < k; ++t) lam[t] = 0.0;
      for (int t = 0; t < rank; ++t) lam[idx[t]] = ls[t];
      have_first = true;
      if (okk) return true;
    }
  }
  return false;
}

// ------------------------------------------------------------------ shared-memory plan (per group)
// Only the multipliers are stored per constraint row: the slacks are recomputed from u (5 flops per
// leg-step), the affine step's row quantities from a copy of the affine direction (dua), and the
// polish's null-space bases live in the group's L2 slab.  Aliases (lifetimes do not overlap):
// eq/qz (build only) in rd..tv; the lever arms (build only) and the polish's candidate point `up`
// in du; the desired fz (build + start point) in rhs.  2886 doubles at nb <= 20, N = 10: 10 groups/SM.
__host__ __device__ inline SmemPlan make_plan(int N, int L, int W, int nbmax, int n4max, int m_in_smem) {
  SmemPlan p;
  const int mmax = 5 * nbmax;
  int o = 0;
  auto take = [&](int cnt) { int r = o; o += (cnt + 1) & ~1; return r; };
  p.ce = take(nbmax);
  p.g = take(n4max); p.u = take(n4max);
  const int nv = (2 * n4max >= 10 * N + 2) ? n4max : (10 * N + 2 + 1) / 2;  // rd+tv also host eq[9N], qz[N]
  p.rd = take(nv); p.tv = take(nv);
  p.rhs = take(n4max); p.du = take(n4max); p.dua = take(n4max);
  p.zl = take(mmax); p.zu = take(mmax);
  p.red = take(W > 1 ? 3 * W : 2);
  p.exch = take(8);
  // bytes: blk_j, blk_i, rk [nbmax each], blk_of [N*L] (int8), actl, actu [mmax each];
  // uint16: off [nbmax], tile table [tiles]; int32 misc[4]
  const int nbytes = 3 * nbmax + N * L + 2 * mmax + 2 + 2 * nbmax + 2 * bc4_tiles(n4max) + 2 + 16;
  p.ints = take((nbytes + 7) / 8);
  o = (o + 1) & ~1;  // 16-byte align tiles
  p.Mm = o;
  if (m_in_smem) o += mat_region_doubles(N, L, n4max);
  p.total = (o + 1) & ~1;
  p.cta = (3 * N + 1) & ~1;
  return p;
}

// ------------------------------------------------------------------ prologue and build (both kernels)
// Shared-memory views of one group used while an instance is unpacked and its QP is built.
struct BuildView {
  double* Mm;   // the matrix buffer; stages the raw inputs [state | des_state | des_inputs] first
  double *ce, *fz, *arm, *eq, *qz, *g;
  const double* qzt;  // CTA-shared table: z-position weight of node k + 1 (an exp() per node and instance otherwise)
  int* misc;    // [0] = nb (clamped to the class bound), [1] = invalid table, [3] = nb unclamped
  uint16_t* tb;
  uint8_t *blk_j, *blk_i;
  int8_t* blk_of;
};

// Inputs -> shared memory, contact table -> free blocks, zero-input rollout errors.  Returns
// whether every input is finite (uniform over the group); ends with a group barrier.
template <int W>
__device__ __forceinline__ bool stage_inputs(const Group<W>& G, const DevConfig& cfg, const SolveArgs& args, int inst,
                                             const BuildView& V) {
  constexpr int GT = Group<W>::GT;
  const int gtid = G.gtid;
  const int N = cfg.N, L = cfg.L;
  const int ns = 9 + 3 * L, nds = 9 * (N + 1), ndi = L * (4 * N + 3);
  const double dt = cfg.dt, mass = cfg.mass;
  // ---- stage the raw inputs once, coalesced (CentroidalMPC.cpp:284-317 copies them blindly;
  // here non-finite values are caught).  L2-only loads (ld.global.cg): each value is read once, and
  // in the progressive host-buffer path the copy engine writes these buffers while the kernel runs.  The buffers may be HBM or mapped pinned host memory
  // (zero-copy end-to-end path): every input byte crosses the bus exactly once.  The staging
  // area is the factor's buffer, which is dead until the build.
  double* g_state = V.Mm;
  double* g_ds = V.Mm + ns;
  double* g_di = V.Mm + ns + nds;
  bool finite = true;
  {
    // The three arrays land back to back in the staging area, so they are read as ONE concatenated stream, eight loads per
    // thread in flight at a time: as three loops the compiler batches within each, and an instance paid five dependent
    // round trips to L2 / HBM (two are left at horizon 10 with one warp).
    const double* s0 = args.state + (size_t)inst * ns;
    const double* s1 = args.des_state + (size_t)inst * nds;
    const double* s2 = args.des_inputs + (size_t)inst * ndi;
    const int ntot = ns + nds + ndi;
    for (int t0 = gtid; t0 < ntot; t0 += 8 * GT) {
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int t = t0 + u * GT;
        v[u] = 0.0;
        if (t < ntot) v[u] = __ldcg(t < ns ? s0 + t : (t < ns + nds ? s1 + (t - ns) : s2 + (t - ns - nds)));
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int t = t0 + u * GT;
        if (t < ntot) { g_state[t] = v[u]; finite = finite && isfinite(v[u]); }
      }
    }
  }
  const double* g_dpos = g_ds;
  const double* g_dvel = g_ds + 3 * (N + 1);
  const double* g_dam = g_ds + 6 * (N + 1);
  finite = G.all(finite);

  // ---- contact table -> free blocks; validity (CentroidalMPC.cpp:328-330).  Lane j of the
  // group's first warp owns step j (N <= 32): column sum, stance count, exclusive prefix by
  // shuffles, then it numbers its own stance legs.
  if (gtid < 32) {
    const int j = gtid;
    double colsum = 0.0;
    int cnt = 0;
    if (j < N)
      for (int i = 0; i < L; ++i) { const double ce = g_di[i * (4 * N + 3) + j]; colsum += ce; cnt += ce > 0.0 ? 1 : 0; }
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (gtid >= o) incl += v; }
    const unsigned bad = __ballot_sync(0xffffffffu, j < N && !(colsum > 0.0));
    const int total = __shfl_sync(0xffffffffu, incl, N - 1);
    if (j < N) {
      int b = incl - cnt;
      const double fzj = (colsum > 0.0) ? mass * kGrav / colsum : 0.0;  // desired fz, :331-333 (one division per step)
      for (int i = 0; i < L; ++i) {
        const double ce = g_di[i * (4 * N + 3) + j];
        if (ce > 0.0 && b < args.nbmax) {
          V.blk_j[b] = j; V.blk_i[b] = i; V.blk_of[j * L + i] = b;
          V.ce[b] = ce;
          V.fz[b] = fzj;
          ++b;
        } else {
          V.blk_of[j * L + i] = -1;
        }
      }
    }
    if (gtid == 0) { V.misc[0] = total < args.nbmax ? total : args.nbmax; V.misc[1] = bad != 0u; V.misc[3] = total; }
  }
  // zero-input rollout (closed form of x_k = A^k x0 + sum A^p d) and e = Q (x - x_ref), nodes 1..N
  for (int k = gtid; k < N; k += GT) {
    const int node = k + 1;
    const double kk = (double)node;
    const double gpos = cfg.zoh ? 0.5 * kk * kk : 0.5 * kk * (kk - 1.0);
    double c[3], v[3];
    for (int a = 0; a < 3; ++a) { c[a] = g_state[a] + kk * dt * g_state[3 + a]; v[a] = g_state[3 + a]; }
    c[2] += gpos * dt * dt * (-kGrav);
    v[2] += kk * dt * (-kGrav);
    const double qz = V.qzt[k];  // omega_{k+1}^2, omega = w2/2 exp(-(k+1)) + w2/2 (:205, :210: inside the square)
    V.qz[k] = qz;
    V.eq[9 * k + 0] = cfg.w[0] * (c[0] - g_dpos[3 * node + 0]);
    V.eq[9 * k + 1] = cfg.w[1] * (c[1] - g_dpos[3 * node + 1]);
    V.eq[9 * k + 2] = qz * (c[2] - g_dpos[3 * node + 2]);
    for (int a = 0; a < 3; ++a) {
      V.eq[9 * k + 3 + a] = cfg.w[3 + a] * (v[a] - g_dvel[3 * node + a]);
      V.eq[9 * k + 6 + a] = cfg.w[6 + a] * (g_state[6 + a] - g_dam[3 * node + a]);
    }
  }
  G.sync();
  return finite;
}

// z-weighted power-stacking sums of the position rows, one table per CTA (they depend on the weights only):
//   z1[j] = sum_{k >= j} (k - j + zeta) qz_k,  z2[j] = sum_{k >= j} (k - j + zeta)^2 qz_k,  qz_k = omega_{k+1}^2
// (CentroidalMPC.cpp:203-210).  Called by the first N threads of the CTA, followed by __syncthreads().
__device__ __forceinline__ void fill_z_tables(const DevConfig& cfg, double* z1, double* z2, int j) {
  const double zeta = cfg.zoh ? 0.5 : 0.0;
  double a1 = 0.0, a2 = 0.0;
  for (int k = j; k < cfg.N; ++k) {
    const double om = (cfg.w[2] * 0.5) * exp(-(double)(k + 1)) + cfg.w[2] * 0.5;
    const double al = (double)(k - j) + zeta;
    a1 += al * om * om; a2 += al * al * om * om;
  }
  z1[j] = a1; z2[j] = a2;
}

// H = 2 (Bqp' L Bqp + K) in BC4 layout into Hb, g into V.g; reads the inputs staged in V.Mm (Hb may
// be V.Mm itself: the staged inputs are consumed before the first tile is written).
// One thread per block pair (b >= b2, hence j >= j2), every lane the same instruction stream.  With
// d = j - j2, cnt = N - j (row blocks k >= j contribute; column (j,i) of Bqp at row block k is
// A_d^{k-j} B_j = [dt^2 (k-j+zeta)(c/m) I; dt (c/m) I; dt c [r]x], Euler zeta = 0, ZOH 1/2):
//   angular   cnt dt^2 c c2 [r]x' diag(w6..8) [r2]x
//   diagonal  (c/m)(c2/m) (dt^4 P_a + cnt dt^2 w[3+a]),  P_a = w[a] (S2 + d S1) for x, y and z2[j] + d z1[j]
//             for z;  S1, S2 = sums of (t + zeta), (t + zeta)^2 over t < cnt
//   same leg  K = W_f + D' W_r D (CentroidalMPC.cpp:223-231)
// Element (gi, gj) of a lower tile lives at R(gi) + C(gj): R = (gi >> 2) kTS + 4 (gi & 3),
// C = colbase(gj >> 2) kTS + (gj & 3).  Far pairs (b2 <= b - 2) lie strictly below the diagonal tiles;
// near pairs (b2 = b, b - 1) may touch one and mirror into it; the two kinds run in separate loops.
template <int W>
__device__ __forceinline__ void build_qp(const Group<W>& G, const DevConfig& cfg, const BuildView& V, double* Hb, int nb,
                                         const double* z1, const double* z2) {
  constexpr int GT = Group<W>::GT;
  const int gtid = G.gtid;
  const int N = cfg.N, L = cfg.L;
  const int ns = 9 + 3 * L, nds = 9 * (N + 1);
  const double dt = cfg.dt, mass = cfg.mass;
  const double zeta = cfg.zoh ? 0.5 : 0.0;
  const int n = 3 * nb, nblk = (n + 3) >> 2, n4 = nblk << 2;
  // lever arms r = des_foot_pos[:, j] - des_com_pos[:, j] (frozen), tile table
  for (int b = gtid; b < nb; b += GT) {
    const int j = V.blk_j[b], i = V.blk_i[b];
    for (int q = 0; q < 3; ++q) V.arm[3 * b + q] = V.Mm[ns + nds + i * (4 * N + 3) + N + 3 * j + q] - V.Mm[ns + 3 * j + q];
  }
  for (int bj = gtid; bj < nblk; bj += GT) {
    const int o = blkoff(bj, bj, nblk);
    for (int bi = bj; bi < nblk; ++bi) V.tb[o + bi - bj] = (uint16_t)(bi | (bj << 8));
  }
    G.sync();  // all reads of the staged inputs (they live in the factor's buffer) are done
    {
      const double dt2 = dt * dt, dt4 = dt2 * dt2;
      const double q0 = cfg.w[6], q1 = cfg.w[7], q2 = cfg.w[8];
      const double im2 = 1.0 / (mass * mass);
      auto Rof = [&](int g) { return (g >> 2) * kTS + ((g & 3) << 2); };
      auto Cof = [&](int g) { const int tj = g >> 2; return (tj * nblk - ((tj * (tj + 1)) >> 1)) * kTS + (g & 3); };
      auto do_pair = [&](int b, int b2, auto far_tag) {
        constexpr bool FAR = decltype(far_tag)::value;
        const int j = V.blk_j[b], i = V.blk_i[b], j2 = V.blk_j[b2], i2 = V.blk_i[b2];
        const double ce = V.ce[b], ce2 = V.ce[b2];
        const double r0 = V.arm[3 * b], r1 = V.arm[3 * b + 1], r2 = V.arm[3 * b + 2];
        const double p0 = V.arm[3 * b2], p1 = V.arm[3 * b2 + 1], p2 = V.arm[3 * b2 + 2];
        const double cnt = (double)(N - j), dd = (double)(j - j2);
        const double s1 = 0.5 * cnt * (cnt - 1.0) + zeta * cnt;
        const double s2 = (cnt - 1.0) * cnt * (2.0 * cnt - 1.0) * (1.0 / 6.0) + zeta * cnt * (cnt - 1.0) + zeta * zeta * cnt;
        const double s0 = s2 + dd * s1, sz = z2[j] + dd * z1[j];
        const double cc = ce * ce2;
        const double sc = cnt * dt2 * cc, cmm = cc * im2;
        double blk[3][3];
        blk[0][0] = sc * (r2 * q1 * p2 + r1 * q2 * p1) + cmm * (dt4 * cfg.w[0] * s0 + cnt * dt2 * cfg.w[3]);
        blk[0][1] = sc * (-r1 * q2 * p0);
        blk[0][2] = sc * (-r2 * q1 * p0);
        blk[1][0] = sc * (-r0 * q2 * p1);
        blk[1][1] = sc * (r2 * q0 * p2 + r0 * q2 * p0) + cmm * (dt4 * cfg.w[1] * s0 + cnt * dt2 * cfg.w[4]);
        blk[1][2] = sc * (-r2 * q0 * p1);
        blk[2][0] = sc * (-r0 * q1 * p2);
        blk[2][1] = sc * (-r1 * q0 * p2);
        blk[2][2] = sc * (r1 * q0 * p1 + r0 * q1 * p0) + cmm * (dt4 * sz + cnt * dt2 * cfg.w[5]);
        if (i == i2 && j - j2 <= 1) {
          const double nn = (j > 0 ? 1.0 : 0.0) + (j + 1 < N ? 1.0 : 0.0);
#pragma unroll
          for (int aa = 0; aa < 3; ++aa) {
            const double wr = cfg.w[9 + 6 * L + 3 * i + aa];
            blk[aa][aa] += (j == j2) ? cfg.w[9 + 3 * L + 3 * i + aa] + nn * wr : -wr;
          }
        }
        const int g0 = 3 * b, h0 = 3 * b2;
        if constexpr (FAR) {
          int R[3], C[3];
#pragma unroll
          for (int aa = 0; aa < 3; ++aa) { R[aa] = Rof(g0 + aa); C[aa] = Cof(h0 + aa); }
#pragma unroll
          for (int aa = 0; aa < 3; ++aa)
#pragma unroll
            for (int bb = 0; bb < 3; ++bb) Hb[R[aa] + C[bb]] = 2.0 * blk[aa][bb];
        } else {
#pragma unroll
          for (int aa = 0; aa < 3; ++aa)
#pragma unroll
            for (int bb = 0; bb < 3; ++bb) {
              const int gi = g0 + aa, gj = h0 + bb;
              const double v = 2.0 * blk[aa][bb];
              if ((gi >> 2) >= (gj >> 2)) Hb[Rof(gi) + Cof(gj)] = v;
              if (b != b2 && (gi >> 2) == (gj >> 2)) Hb[Rof(gj) + Cof(gi)] = v;
            }
        }
      };
      const int nfar = nb >= 3 ? ((nb - 1) * (nb - 2)) >> 1 : 0;
      for (int idx = gtid; idx < nfar; idx += GT) {
        int a = (int)((sqrtf(8.0f * (float)idx + 1.0f) - 1.0f) * 0.5f);
        while (((a + 1) * (a + 2)) >> 1 <= idx) ++a;
        while ((a * (a + 1)) >> 1 > idx) --a;
        do_pair(a + 2, idx - ((a * (a + 1)) >> 1), std::true_type{});
      }
      for (int idx = gtid; idx < 2 * nb - 1; idx += GT) {
        const int b = (idx + 1) >> 1;
        do_pair(b, b - (idx & 1), std::false_type{});
      }
      // padding rows (n .. n4-1): identity
      for (int e = gtid; e < (n4 - n) * n4; e += GT) {
        const int gi = n + e / n4, gj = e % n4;
        if ((gi >> 2) >= (gj >> 2)) Hb[Rof(gi) + Cof(gj)] = gi == gj ? 1.0 : 0.0;
        if ((gi >> 2) == (gj >> 2) && gj < n) Hb[Rof(gj) + Cof(gi)] = 0.0;
      }
    }
    // ---- g = 2 Bqp' L (Aqp x0 + dqp - Xref) - 2 W_f Uref, one thread per block (adjoint sum)
    for (int b = gtid; b < nb; b += GT) {
      const int j = V.blk_j[b], i = V.blk_i[b];
      const double ce = V.ce[b];
      const double r[3] = {V.arm[3 * b], V.arm[3 * b + 1], V.arm[3 * b + 2]};
      double sp[3] = {0, 0, 0}, sv[3] = {0, 0, 0}, sl3[3] = {0, 0, 0};
      for (int k = j; k < N; ++k) {
        const double al = (double)(k - j) + zeta;
        for (int q = 0; q < 3; ++q) {
          sp[q] += al * V.eq[9 * k + q]; sv[q] += V.eq[9 * k + 3 + q]; sl3[q] += V.eq[9 * k + 6 + q];
        }
      }
      const double cm = ce / mass;
      // [r]x' v = v x r
      const double cr[3] = {sl3[1] * r[2] - sl3[2] * r[1], sl3[2] * r[0] - sl3[0] * r[2], sl3[0] * r[1] - sl3[1] * r[0]};
      for (int q = 0; q < 3; ++q) {
        double gq = 2.0 * (cm * (dt * dt * sp[q] + dt * sv[q]) + dt * ce * cr[q]);
        if (q == 2) gq -= 2.0 * cfg.w[9 + 3 * L + 3 * i + 2] * V.fz[b];
        V.g[3 * b + q] = gq;
      }
    }
    if (gtid < n4 - n) V.g[n + gtid] = 0.0;
    G.sync();
}

// ------------------------------------------------------------------ presolve kernel: shared-memory plan
// Chunk-major tile layout of the presolve kernel (cmpc_presolve.cu): tile t of the BC4 ordering is
// eight 16-byte chunks (row a, column pair h -> chunk 2a + h) stored at double2 index  chunk * T + t
// with T odd.  Lanes that walk consecutive tiles, or the four rows of one tile, then fall into
// distinct 16-byte bank groups WITHOUT the padding of the kTS = 18 layout -- 1936 instead of 2160
// doubles at n = 60, which is what lets 14 instances share an SM (two waves for a 4096 batch).
__host__ __device__ inline PrePlan make_pre_plan(int N, int L, int W, int nbmax, int n4max) {
  PrePlan p;
  const int tiles = bc4_tiles(n4max);
  p.T = tiles | 1;
  // CTA-shared tables: z1[N], z2[N] (z-weighted power-stacking sums), s1[N], s2[N] (unit weights), qz[N], wf[3L], wr[3L]
  p.cta = (5 * N + 6 * L + 1) & ~1;
  int o = 0;
  auto take = [&](int cnt) { int r = o; o += (cnt + 1) & ~1; return r; };
  p.x = take(n4max);          // rhs staging, lever arms during the build, then y and the solution
  p.ce = take(nbmax);
  p.red = take(W > 1 ? 3 * W : 0);
  p.exch = take(8);
  // misc int32[4]; tile table uint16[tiles]; blk_j, blk_i uint8[nbmax]; blk_of int8[N L]
  const int nbytes = 16 + 2 * (tiles + (tiles & 1)) + 2 * nbmax + N * L;
  p.ints = take((nbytes + 7) / 8);
  p.M = o;
  // the matrix region also stages [inputs | eq (9N) | qz (N) | fz (nbmax)] before the build
  const int nin = (9 + 3 * L) + 9 * (N + 1) + L * (4 * N + 3);
  const int stage = ((nin + 1) & ~1) + 10 * N + nbmax + 2;
  p.mat = 16 * p.T > stage ? 16 * p.T : stage;
  p.mat = (p.mat + 1) & ~1;
  o += p.mat;
  p.total = o;
  return p;
}

// ------------------------------------------------------------------ Riccati presolve kernel: shared-memory plan
// Per group: the staged inputs stay resident (the backward sweep reads arms, contacts and references
// stage by stage), the 21x21 cost-to-go, the per-stage products, and -- aliased onto those once the
// sweep is done -- the state trajectory and the forces.  Gains go to an L2 slab.
__host__ __device__ inline RicPlan make_ric_plan(int N, int L) {
  RicPlan p;
  const int nz = 9 + 3 * L, nbfull = L * N;
  p.nz = nz;
  int o = 0;
  auto take = [&](int cnt) { int r = o; o += (cnt + 1) & ~1; return r; };
  const int nin = (9 + 3 * L) + 9 * (N + 1) + L * (4 * N + 3);
  p.in = take((nin + 1) & ~1);  // the staged inputs stay resident through all three sweeps
  // (the by-products of the shared prologue -- eq, qz, fz, ce -- are not needed here: they land in the
  // work region P..M, which is initialised afterwards)
  p.ce = 0;
  p.P = take(nz * nz);
  p.Bb = take(nz * 12); p.T1 = take(nz * 12); p.G = take(12 * 13); p.M = take(12 * nz + 12);
  // once the backward sweep is done P and T1..M are dead: the trajectory x_1..x_N and the forces go there
  p.X = (9 * N <= nz * nz) ? p.P : take(9 * N);
  p.F = (3 * L * N <= o - p.T1) ? p.T1 : take(3 * L * N);
  p.p = take(nz); p.t = take(nz); p.m0 = take(12); p.xc = take(2 * nz + 4 * 9 + 12);
  const int nbytes = 16 + 2 * nbfull + nbfull + 64;  // misc, blk_j, blk_i, blk_of, comp tables
  p.ints = take((nbytes + 7) / 8);
  p.total = o;
  p.slab = N * (12 * nz + 12);
  return p;
}

// ------------------------------------------------------------------ kernel launchers (one translation unit per kernel family)
// cmpc_solve.cu: W in {1, 2, 4, 8}; mode 0 = solve, 1 = build-export; ms = factor in shared memory
cudaError_t launch_solve_kernel(int W, int mode, bool ms, int phase, int grid, int block, size_t smem, cudaStream_t stream,
                                const DevConfig& cfg, const SolveArgs& args);
cudaError_t set_solve_kernel_smem(int W, int mode, bool ms, size_t bytes);
// cmpc_presolve.cu: W in {1, 4, 8}
cudaError_t launch_presolve_kernel(int W, int grid, int block, size_t smem, cudaStream_t stream, const DevConfig& cfg,
                                   const SolveArgs& args);
cudaError_t set_presolve_kernel_smem(int W, size_t bytes);
// cmpc_riccati.cu: stage-wise (Riccati) presolve, one warp per instance
cudaError_t launch_riccati_kernel(int grid, int block, size_t smem, cudaStream_t stream, const DevConfig& cfg,
                                  const SolveArgs& args);
cudaError_t set_riccati_kernel_smem(size_t bytes);
// cmpc_ripm.cu: stage-wise (Riccati) interior-point + polish kernel for the instances the presolve defers, one warp per instance
cudaError_t launch_ripm_kernel(int grid, int block, size_t smem, cudaStream_t stream, const DevConfig& cfg,
                               const SolveArgs& args);
cudaError_t set_ripm_kernel_smem(size_t bytes);
void ripm_sizes(int N, int L, int* group_doubles, int* cta_doubles, int* slab_doubles);
cudaError_t launch_ripm_probe(int grid, int block, size_t smem, cudaStream_t stream, const DevConfig& cfg, const SolveArgs& args,
                              const double* hess, const double* rhs, int mode, double* d_fused, double* d_resolve, double* grad, int B);

}  // namespace cmpc
