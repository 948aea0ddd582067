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
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

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

struct SmemPlan {
  // offsets in doubles from the start of the group's slab
  int ce, g, u, rd, tv, rhs, du, dua;
  int zl, zu, red, exch, ints, Mm;
  int total;  // doubles, multiple of 2
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
  int32_t* fail_perm;        // presolve kernel only: instances it could not settle, for the IPM kernel
  int32_t* fail_count;
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
  double* buf = exch + ((kb & 1) << 2);  // double-buffered: one barrier per step
  const int q0 = G.gtid - (r0 % GT);
  if (q0 >= 0 && q0 < 4) buf[q0] = (r0 >= GT) ? xr[1] : xr[0];
  G.sync();
  const double2* b2p = reinterpret_cast<const double2*>(buf);
  const double2 u = b2p[0], v = b2p[1];
  b0 = u.x; b1 = u.y; b2 = v.x; b3 = v.y;
}

// Tiled right-looking Cholesky in BC4 layout, in place; the whole group calls it.
// tb[t] = bi | bj << 8 for storage tile t.  Diagonal tiles end up in solve form (potrf4).
// If rhs != nullptr the forward substitution  y = L^-1 rhs  is fused into the sweep (rows
// live in registers, the update of step kb rides along with the trailing update) and y
// overwrites rhs.  Returns false (uniformly) on a non-positive pivot.
template <int W>
__device__ __forceinline__ bool chol_bc4(const Group<W>& G, double* M, int nblk, const uint16_t* tb, double* rhs, double* exch) {
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
  for (int kb = 0; kb < nblk; ++kb) {
    const int col0 = blkoff(kb, kb, nblk);  // storage index of the diagonal tile of column kb
    const int nrows = nblk - kb;
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
    for (int t = t0 + gtid; t < ntiles; t += GT) {
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

// Forward substitution y = L^-1 x (in place in shared memory), rows in registers.
template <int W>
__device__ __forceinline__ void chol_fwd_bc4(const Group<W>& G, const double* M, int nblk, double* x, double* exch) {
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

// Backward substitution x = L^-T y (in place in shared memory), rows in registers.
template <int W>
__device__ __forceinline__ void chol_bwd_bc4(const Group<W>& G, const double* M, int nblk, double* x, double* exch) {
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

// y = H x for symmetric H in BC4 layout (diagonal tiles hold both triangles), one row per
// thread (two passes when n4 > GT).  The whole group calls it; ends with a group barrier.
template <int W>
__device__ __forceinline__ void symv_bc4(const Group<W>& G, const double* H, int n4, int nblk, const double* x, double* y) {
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

// Copy a BC4 matrix, 16 bytes per access (global scratch <-> shared).
template <int W>
__device__ __forceinline__ void copy_mat(const Group<W>& G, double* dst, const double* src, int ndoubles) {
  const double2* s2 = reinterpret_cast<const double2*>(src);
  double2* d2 = reinterpret_cast<double2*>(dst);
  for (int t = G.gtid; t < (ndoubles >> 1); t += Group<W>::GT) d2[t] = __ldcg(s2 + t);
}
template <int W>
__device__ __forceinline__ void store_mat(const Group<W>& G, double* dst, const double* src, int ndoubles) {
  const double2* s2 = reinterpret_cast<const double2*>(src);
  double2* d2 = reinterpret_cast<double2*>(dst);
  for (int t = G.gtid; t < (ndoubles >> 1); t += Group<W>::GT) __stcg(d2 + t, s2[t]);
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
__device__ __noinline__ int block_nullspace(int k, const double (*A)[3], const double* b, double* f0, double (*Z)[3], bool* ok) {
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
__device__ __noinline__ double small_lsq(int k, const double (*S)[3], const double* rb, double* lam) {
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
__device__ __noinline__ bool block_multipliers(int k, const double (*Nrm)[3], const double* rb, double tol, double* lam) {
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
  return p;
}

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

// ------------------------------------------------------------------ prologue and build (both kernels)
// Shared-memory views of one group used while an instance is unpacked and its QP is built.
struct BuildView {
  double* Mm;   // the matrix buffer; stages the raw inputs [state | des_state | des_inputs] first
  double *ce, *fz, *arm, *eq, *qz, *g;
  int* misc;    // [0] = nb, [1] = invalid table
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
  // here non-finite values are caught).  The buffers may be HBM or mapped pinned host memory
  // (zero-copy end-to-end path): every input byte crosses the bus exactly once.  The staging
  // area is the factor's buffer, which is dead until the build.
  double* g_state = V.Mm;
  double* g_ds = V.Mm + ns;
  double* g_di = V.Mm + ns + nds;
  bool finite = true;
  {
    const double* src = args.state + (size_t)inst * ns;
    for (int t = gtid; t < ns; t += GT) { const double v = __ldg(src + t); g_state[t] = v; finite = finite && isfinite(v); }
    src = args.des_state + (size_t)inst * nds;
    for (int t = gtid; t < nds; t += GT) { const double v = __ldg(src + t); g_ds[t] = v; finite = finite && isfinite(v); }
    src = args.des_inputs + (size_t)inst * ndi;
    for (int t = gtid; t < ndi; t += GT) { const double v = __ldg(src + t); g_di[t] = v; finite = finite && isfinite(v); }
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
      for (int i = 0; i < L; ++i) {
        const double ce = g_di[i * (4 * N + 3) + j];
        if (ce > 0.0 && b < args.nbmax) {
          V.blk_j[b] = j; V.blk_i[b] = i; V.blk_of[j * L + i] = b;
          V.ce[b] = ce;
          V.fz[b] = (colsum > 0.0) ? mass * kGrav / colsum : 0.0;  // desired fz, :331-333
          ++b;
        } else {
          V.blk_of[j * L + i] = -1;
        }
      }
    }
    if (gtid == 0) { V.misc[0] = total < args.nbmax ? total : args.nbmax; V.misc[1] = bad != 0u; }
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
    const double om = (cfg.w[2] * 0.5) * exp(-kk) + cfg.w[2] * 0.5;  // :205
    const double qz = om * om;                                        // :210 (inside the square)
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

// H = 2 (Bqp' L Bqp + K) in BC4 layout into Hb, g into V.g; reads the inputs staged in V.Mm (Hb may
// be V.Mm itself: the staged inputs are consumed before the first tile is written).
template <int W>
__device__ __forceinline__ void build_qp(const Group<W>& G, const DevConfig& cfg, const BuildView& V, double* Hb, int nb) {
  constexpr int GT = Group<W>::GT;
  const int gtid = G.gtid;
  const int N = cfg.N, L = cfg.L;
  const int ns = 9 + 3 * L, nds = 9 * (N + 1);
  const double dt = cfg.dt, mass = cfg.mass;
  const double zeta = cfg.zoh ? 0.5 : 0.0;
  const int n = 3 * nb, nblk = (n + 3) >> 2, n4 = nblk << 2;
  const int matd = ((nblk * (nblk + 1)) >> 1) * kTS;
  // lever arms r = des_foot_pos[:, j] - des_com_pos[:, j] (frozen), tile table
  for (int b = gtid; b < nb; b += GT) {
    const int j = V.blk_j[b], i = V.blk_i[b];
    for (int q = 0; q < 3; ++q) V.arm[3 * b + q] = V.Mm[ns + nds + i * (4 * N + 3) + N + 3 * j + q] - V.Mm[ns + 3 * j + q];
  }
  for (int bj = gtid; bj < nblk; bj += GT) {
    const int o = blkoff(bj, bj, nblk);
    for (int bi = bj; bi < nblk; ++bi) V.tb[o + bi - bj] = (uint16_t)(bi | (bj << 8));
  }
  // ---- H = 2 (Bqp' L Bqp + K) on the free variables, one thread per block pair (b >= b2),
  // assembled on chip (in the factor's buffer) and then streamed to the L2-resident copy.
  // Column (j,i) of Bqp at row block k >= j is A_d^{k-j} B_j =
  //   [ dt^2 (k-j+zeta) (c/m) I ; dt (c/m) I ; dt c [r]x ]   (a2/a3; Euler zeta=0, ZOH 1/2)
    G.sync();  // all reads of the staged inputs (they live in the factor's buffer) are done
    for (int t = gtid; t < matd; t += GT) Hb[t] = 0.0;
    G.sync();
    const int npairs = (nb * (nb + 1)) >> 1;
    const double dt2 = dt * dt, dt4 = dt2 * dt2;
    for (int idx = gtid; idx < npairs; idx += GT) {
      int a = (int)((sqrtf(8.0f * (float)idx + 1.0f) - 1.0f) * 0.5f);
      while (((a + 1) * (a + 2)) >> 1 <= idx) ++a;
      while ((a * (a + 1)) >> 1 > idx) --a;
      const int b2 = idx - ((a * (a + 1)) >> 1), b = a;  // b >= b2  => j >= j2
      const int j = V.blk_j[b], i = V.blk_i[b], j2 = V.blk_j[b2], i2 = V.blk_i[b2];
      const double ce = V.ce[b], ce2 = V.ce[b2];
      const double r0 = V.arm[3 * b], r1 = V.arm[3 * b + 1], r2 = V.arm[3 * b + 2];
      const double p0 = V.arm[3 * b2], p1 = V.arm[3 * b2 + 1], p2 = V.arm[3 * b2 + 2];
      const double cm = ce / mass, cm2 = ce2 / mass;
      // position rows: sum_k alpha_{k-j} alpha_{k-j2} Qp_k ; only the z weight depends on k
      double s0 = 0.0, sz = 0.0;
      for (int k = j; k < N; ++k) {
        const double aa = ((double)(k - j) + zeta) * ((double)(k - j2) + zeta);
        s0 += aa; sz += aa * V.qz[k];
      }
      const double cnt = (double)(N - j);
      // angular rows: dt^2 c c2 [r]x' diag(ql) [r2]x summed over the N-j row blocks below;
      // [r]x = [[0,-rz,ry],[rz,0,-rx],[-ry,rx,0]]
      const double q0 = cfg.w[6], q1 = cfg.w[7], q2 = cfg.w[8];
      const double sc = cnt * dt2 * ce * ce2;
      double blk[3][3];
      blk[0][0] = sc * (r2 * q1 * p2 + r1 * q2 * p1);
      blk[0][1] = sc * (-r1 * q2 * p0);
      blk[0][2] = sc * (-r2 * q1 * p0);
      blk[1][0] = sc * (-r0 * q2 * p1);
      blk[1][1] = sc * (r2 * q0 * p2 + r0 * q2 * p0);
      blk[1][2] = sc * (-r2 * q0 * p1);
      blk[2][0] = sc * (-r0 * q1 * p2);
      blk[2][1] = sc * (-r1 * q0 * p2);
      blk[2][2] = sc * (r1 * q0 * p1 + r0 * q1 * p0);
      const double pos[3] = {cfg.w[0] * s0, cfg.w[1] * s0, sz};
      for (int aa = 0; aa < 3; ++aa) blk[aa][aa] += cm * cm2 * (dt4 * pos[aa] + cnt * dt2 * cfg.w[3 + aa]);
      // K = W_f + D' W_r D (CentroidalMPC.cpp:223-231): same leg, same component
      if (i == i2) {
        for (int aa = 0; aa < 3; ++aa) {
          const double wr = cfg.w[9 + 6 * L + 3 * i + aa];
          if (j == j2) {
            const double nn = (j > 0 ? 1.0 : 0.0) + (j + 1 < N ? 1.0 : 0.0);
            blk[aa][aa] += cfg.w[9 + 3 * L + 3 * i + aa] + nn * wr;
          } else if (j == j2 + 1) {
            blk[aa][aa] -= wr;
          }
        }
      }
      for (int aa = 0; aa < 3; ++aa)
        for (int bb = 0; bb < 3; ++bb) {
          const int gi = 3 * b + aa, gj = 3 * b2 + bb;
          const double v = 2.0 * blk[aa][bb];
          if (b != b2) {
            Hb[midx(gi, gj, nblk)] = v;                              // gi > gj always here
            if ((gi >> 2) == (gj >> 2)) Hb[midx(gj, gi, nblk)] = v;  // same diagonal tile: mirror
          } else if ((gi >> 2) >= (gj >> 2)) {
            // diagonal 3x3 block: all 9 (aa,bb) are visited, so both triangles of a
            // diagonal tile get written; a straddling entry lands in the lower tile only
            Hb[midx(gi, gj, nblk)] = v;
          }
        }
    }
    if (gtid < n4 - n) Hb[midx(n + gtid, n + gtid, nblk)] = 1.0;  // padding rows: identity
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

// ------------------------------------------------------------------ the fused kernel
// MODE 0: solve.  MODE 1: build-export (H, g in the full 3LN layout to global memory).
template <int W, int MODE, bool MS>
__global__ void __launch_bounds__(W == 2 ? 512 : (W == 1 ? 320 : 256)) cmpc_solve_kernel(const DevConfig cfg, const SolveArgs args) {
  extern __shared__ __align__(128) double smem[];
  constexpr int GT = Group<W>::GT;
  const int N = cfg.N, L = cfg.L, nu = 3 * L;
  const int nf = 3 * L * N;
  const int nbfull = L * N, mfull = 5 * nbfull;
  const int nbmax = args.nbmax, mmax = 5 * nbmax;
  const SmemPlan& P = args.plan;
  Group<W> G;
  G.gtid = threadIdx.x % GT;
  G.gid = threadIdx.x / GT;
  const int gtid = G.gtid;
  double* base = smem + (size_t)G.gid * P.total;
  G.red = base + P.red;
  double* s_exch = base + P.exch;
  double* s_ce = base + P.ce;
  double* s_g = base + P.g;
  double* s_u = base + P.u;
  double* s_rd = base + P.rd;
  double* s_f0 = s_rd;   // polish only; rd is recomputed after a rejected polish
  double* s_rhs = base + P.rhs;
  double* s_fz = s_rhs;  // desired fz: build + start point only
  double* s_du = base + P.du;
  double* s_arm = s_du;  // lever arms: build only
  double* s_tv = base + P.tv;
  double* s_up = s_du;   // polish only: candidate point (du is dead there; chol's rhs is tv)
  double* s_eq = s_rd;   // build only: 9N weighted errors then N z-weights, spanning rd..tv
  double* s_qz = s_rd + 9 * N;
  double* s_dua = base + P.dua;  // copy of the affine (predictor) direction
  double* s_zl = base + P.zl;
  double* s_zu = base + P.zu;
  int* s_misc = reinterpret_cast<int*>(base + P.ints);  // [0]=nb, [1]=invalid, [2]=work slot, [3]=nr
  uint16_t* s_off = reinterpret_cast<uint16_t*>(s_misc + 4);
  uint16_t* s_tb = s_off + nbmax + (nbmax & 1);
  uint8_t* s_blk_j = reinterpret_cast<uint8_t*>(s_tb + bc4_tiles(args.n4max) + (bc4_tiles(args.n4max) & 1));
  uint8_t* s_blk_i = s_blk_j + nbmax;
  uint8_t* s_rk = s_blk_i + nbmax;
  int8_t* s_blk_of = reinterpret_cast<int8_t*>(s_rk + nbmax);  // [N*L] free block index or -1 (nb <= 127)
  unsigned char* s_actl = reinterpret_cast<unsigned char*>(s_blk_of + nbfull);
  unsigned char* s_actu = s_actl + mmax;
  const int group_global = blockIdx.x * args.groups + G.gid;
  double* Hm = args.scratch + (size_t)group_global * args.scratch_per_group;
  double* Mm;
  if constexpr (MS) Mm = base + P.Mm; else Mm = Hm + mat_region_doubles(N, L, args.n4max);
  // polish only: null-space bases, 9 doubles per leg-step, behind the matrices in the L2 slab
  double* g_Zt = Hm + (size_t)mat_region_doubles(N, L, args.n4max) * (MS ? 1 : 2);

  const double mass = cfg.mass;
  const int count = args.count ? *args.count : args.count_imm;

  while (true) {
    int slot = 0;
    if (gtid == 0) slot = atomicAdd(args.work, 1);
    slot = G.bcast0(slot, s_misc + 2);
    if (slot >= count) break;
    const int inst = args.perm ? args.perm[slot] : slot;
    BuildView V;
    V.Mm = Mm; V.ce = s_ce; V.fz = s_fz; V.arm = s_arm; V.eq = s_eq; V.qz = s_qz; V.g = s_g;
    V.misc = s_misc; V.tb = s_tb; V.blk_j = s_blk_j; V.blk_i = s_blk_i; V.blk_of = s_blk_of;
    const bool finite = stage_inputs<W>(G, cfg, args, inst, V);
    const int nb = s_misc[0];
    const int n = 3 * nb, nblk = (n + 3) >> 2, n4 = nblk << 2, m = 5 * nb;
    const int ntiles = (nblk * (nblk + 1)) >> 1;
    const int matd = ntiles * kTS;
    const bool invalid = s_misc[1] != 0;

    if (MODE == 0 && (!finite || invalid)) {
      for (int t = gtid; t < nf; t += GT) args.forces[(size_t)inst * nf + t] = 0.0;
      if (args.lam) for (int t = gtid; t < 2 * mfull; t += GT) args.lam[(size_t)inst * 2 * mfull + t] = 0.0;
      if (args.active) for (int t = gtid; t < nbfull; t += GT) args.active[(size_t)inst * nbfull + t] = 0;
      if (gtid == 0) {
        args.status[inst] = !finite ? CMPC_STATUS_NUMERICAL : CMPC_STATUS_INVALID_TABLE;
        if (args.iters) args.iters[inst] = 0;
        if (args.kkt) args.kkt[inst] = 0.0;
      }
      G.sync();
      continue;
    }

    double* Hb;
    if constexpr (MS) Hb = Mm; else Hb = Hm;
    build_qp<W>(G, cfg, V, Hb, nb);
    if constexpr (MS) store_mat<W>(G, Hm, Mm, matd);

    if (MODE == 1) {
      // export H, g in the full 3LN step-major layout with pinned rows/cols = identity
      const int p = nf;
      double* Ho = args.Hout + (size_t)inst * p * p;
      double* go = args.gout + (size_t)inst * p;
      for (int t = gtid; t < p * p; t += GT) {
        const int a = t / p, c = t % p;
        const int ba = s_blk_of[(a / nu) * L + (a % nu) / 3], bc = s_blk_of[(c / nu) * L + (c % nu) / 3];
        double v;
        if (ba < 0 || bc < 0) v = (a == c) ? 1.0 : 0.0;
        else v = Hb[sidx(3 * ba + a % 3, 3 * bc + c % 3, nblk)];
        Ho[t] = v;
      }
      for (int t = gtid; t < p; t += GT) {
        const int ba = s_blk_of[(t / nu) * L + (t % nu) / 3];
        go[t] = ba < 0 ? 0.0 : s_g[3 * ba + t % 3];
      }
      if (gtid == 0) args.status[inst] = !finite ? CMPC_STATUS_NUMERICAL : (invalid ? CMPC_STATUS_INVALID_TABLE : CMPC_STATUS_OK);
      G.sync();
      continue;
    }

    // ---- strictly feasible start f = (0, 0, fz0); centred duals
    for (int b = gtid; b < nb; b += GT) {
      const double mub = cfg.mu[s_blk_i[b]];
      double fz = s_fz[b];
      fz = fmin(fz, 0.5 * mass * kGrav * (double)L * s_ce[b]);
      fz = fmin(fz, 0.5 * kFricUb * s_ce[b] / mub);
      s_u[3 * b] = 0.0; s_u[3 * b + 1] = 0.0; s_u[3 * b + 2] = fz;
    }
    if (gtid < n4 - n) s_u[n + gtid] = 0.0;
    if constexpr (!MS) copy_mat<W>(G, Mm, Hm, matd);
    G.sync();
    // (s_fz aliases s_rhs, s_arm aliases s_du, s_eq aliases s_rd/s_tv: all dead from here on)
    if (gtid < n4 - n) { s_rhs[n + gtid] = 0.0; s_du[n + gtid] = 0.0; s_dua[n + gtid] = 0.0; s_rd[n + gtid] = 0.0; s_tv[n + gtid] = 0.0; }
    symv_bc4<W>(G, Mm, n4, nblk, s_u, s_rd);
    double gmax = 0.0, r0max = 0.0;
    for (int t = gtid; t < n; t += GT) { gmax = fmax(gmax, fabs(s_g[t])); r0max = fmax(r0max, fabs(s_rd[t] + s_g[t])); }
    gmax = G.max(gmax);
    r0max = G.max(r0max);
    const double gs = 1.0 + gmax;
    const double mu0 = fmax(1e-2, r0max);
    for (int b = gtid; b < nb; b += GT) {
      const double mub = cfg.mu[s_blk_i[b]];
      const double ubxy = kFricUb * s_ce[b], ubz = mass * kGrav * (double)L * s_ce[b];  // :183,199
      double y[5];
      cmul5(mub, s_u + 3 * b, y);
      for (int q = 0; q < 5; ++q) {
        const double ub = q < 4 ? ubxy : ubz;
        s_zl[5 * b + q] = mu0 / y[q]; s_zu[5 * b + q] = mu0 / (ub - y[q]);
      }
    }
    G.sync();

    int status = CMPC_STATUS_MAX_ITER, it = 0, npolish = 0;
    bool numerical = false, ipm_ok = false, m_is_h = true, warm_done = false;
    double us = 1.0;
    for (it = 0; it <= cfg.max_iter; ++it) {
      // ---- residuals (M holds a fresh copy of H here)
      if (!m_is_h) { copy_mat<W>(G, Mm, Hm, matd); G.sync(); m_is_h = true; }
      if (it > 0) symv_bc4<W>(G, Mm, n4, nblk, s_u, s_rd);  // it == 0: rd still holds H u0 from the start point
      double rmax = 0.0, umax = 0.0, gap = 0.0;
      for (int b = gtid; b < nb; b += GT) {
        const double ubxy = kFricUb * s_ce[b], ubz = mass * kGrav * (double)L * s_ce[b];
        double w[5], o[3], ys[5];
        cmul5(cfg.mu[s_blk_i[b]], s_u + 3 * b, ys);
        for (int q = 0; q < 5; ++q) {
          const int t = 5 * b + q;
          const double sl = ys[q], su = (q < 4 ? ubxy : ubz) - sl;
          w[q] = s_zl[t] - s_zu[t];
          gap += sl * s_zl[t] + su * s_zu[t];
        }
        ctmul5(cfg.mu[s_blk_i[b]], w, o);
        for (int q = 0; q < 3; ++q) {
          const double rr = s_rd[3 * b + q] + s_g[3 * b + q] - o[q];
          s_rd[3 * b + q] = rr;
          rmax = fmax(rmax, fabs(rr)); umax = fmax(umax, fabs(s_u[3 * b + q]));
        }
      }
      G.max2_sum(rmax, umax, gap);
      const double mu = gap / (2.0 * (double)m);
      us = 1.0 + umax;
      // Convergence. The dual residual has a round-off floor ~ eps * cond(H + C'SC) once the
      // gap is small, so the polish (which verifies the KKT conditions itself) is attempted
      // as soon as the gap is converged and the residual is merely small.
      const bool conv_mu = mu <= cfg.tol * gs * us;
      const bool strict = conv_mu && rmax <= cfg.tol * gs;
      const bool ready = conv_mu && rmax <= 1e4 * cfg.tol * gs;
      ipm_ok = conv_mu && rmax <= 10.0 * cfg.tol * gs;
      // Warm start (closed loop): before the first factorisation, try the previous tick's
      // active set, shifted by one step, as the polish's guess. The polish verifies the KKT
      // conditions, so a wrong guess only costs its correction passes and the IPM runs cold.
      const bool warm_now = cfg.polish && args.warm_active != nullptr && it == 0 && !warm_done;
      if (cfg.polish && ((ready && npolish < 3) || warm_now)) {
        bool any_act = false;
        if (warm_now) {
          warm_done = true;
          const uint16_t* wa = args.warm_active + (size_t)inst * nbfull;
          for (int b = gtid; b < nb; b += GT) {
            const int j = s_blk_j[b], i = s_blk_i[b];
            const int jj = j + 1 < N ? j + 1 : j;
            const unsigned a = wa[jj * L + i];
            for (int q = 0; q < 5; ++q) {
              const bool al = !(a & 0x8000u) && ((a >> q) & 1u), au = !(a & 0x8000u) && ((a >> (5 + q)) & 1u);
              s_actl[5 * b + q] = al; s_actu[5 * b + q] = au;
              any_act = any_act || al || au;
            }
          }
        } else {
          ++npolish;
          for (int b = gtid; b < nb; b += GT) {
            const double ubxy = kFricUb * s_ce[b], ubz = mass * kGrav * (double)L * s_ce[b];
            double ys[5];
            cmul5(cfg.mu[s_blk_i[b]], s_u + 3 * b, ys);
            for (int q = 0; q < 5; ++q) {
              const int t = 5 * b + q;
              const double sl = ys[q], su = (q < 4 ? ubxy : ubz) - sl;
              const bool al = s_zl[t] * us > sl * gs, au = s_zu[t] * us > su * gs;
              s_actl[t] = al; s_actu[t] = au;
              any_act = any_act || al || au;
            }
          }
        }
        bool none_active = G.all(!any_act);
        // ---- active-set polish with correction passes
        bool accepted = false;
        for (int pass = 0; pass < 6 && !accepted; ++pass) {
          int nr, nblk_r;
          if (none_active) {
            // no active row: Z = I, f0 = 0, the reduced system is (H, -g) itself
            nr = n; nblk_r = nblk;
            for (int b = gtid; b < nb; b += GT) { s_rk[b] = 0; s_off[b] = 3 * b; }
            for (int t = gtid; t < n4; t += GT) { s_f0[t] = 0.0; s_tv[t] = t < n ? -s_g[t] : 0.0; }
            if (!m_is_h) copy_mat<W>(G, Mm, Hm, matd);
            G.sync();
          } else {
            bool ok_all = true;
            for (int b = gtid; b < nb; b += GT) {
              const double mub = cfg.mu[s_blk_i[b]];
              const double ubxy = kFricUb * s_ce[b], ubz = mass * kGrav * (double)L * s_ce[b];
              double A[10][3], rhsb[10], Z[3][3], f0[3];
              int k = 0;
              for (int q = 0; q < 5; ++q) if (s_actl[5 * b + q]) { row_vec(mub, q, A[k]); rhsb[k++] = 0.0; }
              for (int q = 0; q < 5; ++q) if (s_actu[5 * b + q]) { row_vec(mub, q, A[k]); rhsb[k++] = q < 4 ? ubxy : ubz; }
              bool okb;
              const int rk = block_nullspace(k, A, rhsb, f0, Z, &okb);
              ok_all = ok_all && okb;
              s_rk[b] = rk;
              for (int q = 0; q < 3; ++q) s_f0[3 * b + q] = f0[q];
              for (int cc = 0; cc < 3 - rk; ++cc)
                for (int q = 0; q < 3; ++q) g_Zt[9 * b + 3 * cc + q] = Z[cc][q];
            }
            if (gtid < n4 - n) s_f0[n + gtid] = 0.0;
            ok_all = G.all(ok_all);
            if (!ok_all) break;
            if (gtid == 0) {
              int o = 0;
              for (int b = 0; b < nb; ++b) { s_off[b] = o; o += 3 - s_rk[b]; }
              s_misc[3] = o;
            }
            if (!m_is_h) copy_mat<W>(G, Mm, Hm, matd);
            G.sync();
            nr = s_misc[3];
            nblk_r = (nr + 3) >> 2;
            const int nr4 = nblk_r << 2;
            // r = H f0 + g   (H read from the shared copy)
            symv_bc4<W>(G, Mm, n4, nblk, s_f0, s_rhs);
            for (int t = gtid; t < n; t += GT) s_rhs[t] += s_g[t];
            G.sync();
            // reduced system Z'HZ t = -Z'(H f0 + g); H is read from the global copy, the
            // reduced matrix is assembled in M
            {
              const int matr = bc4_doubles(nr4);
              for (int t = gtid; t < matr; t += GT) Mm[t] = 0.0;
            }
            for (int b = gtid; b < nb; b += GT)
              for (int cc = 0; cc < 3 - s_rk[b]; ++cc) {
                const double* z = g_Zt + 9 * b + 3 * cc;
                s_tv[s_off[b] + cc] = -(__ldcg(z) * s_rhs[3 * b] + __ldcg(z + 1) * s_rhs[3 * b + 1] + __ldcg(z + 2) * s_rhs[3 * b + 2]);
              }
            G.sync();
            if (gtid < nr4 - nr) { s_tv[nr + gtid] = 0.0; Mm[midx(nr + gtid, nr + gtid, nblk_r)] = 1.0; }
            const int npairs = (nb * (nb + 1)) >> 1;
            for (int idx = gtid; idx < npairs; idx += GT) {
              int a = (int)((sqrtf(8.0f * (float)idx + 1.0f) - 1.0f) * 0.5f);
              while (((a + 1) * (a + 2)) >> 1 <= idx) ++a;
              while ((a * (a + 1)) >> 1 > idx) --a;
              const int b2 = idx - ((a * (a + 1)) >> 1), b = a;
              const int d1 = 3 - s_rk[b], d2 = 3 - s_rk[b2];
              if (d1 == 0 || d2 == 0) continue;
              double Hb3[3][3], Za[3][3], Zb[3][3];
              for (int aa = 0; aa < 3; ++aa)
                for (int bb = 0; bb < 3; ++bb) {
                  Hb3[aa][bb] = __ldcg(Hm + sidx(3 * b + aa, 3 * b2 + bb, nblk));
                  Za[aa][bb] = aa < d1 ? __ldcg(g_Zt + 9 * b + 3 * aa + bb) : 0.0;
                  Zb[aa][bb] = aa < d2 ? __ldcg(g_Zt + 9 * b2 + 3 * aa + bb) : 0.0;
                }
              for (int cc = 0; cc < d1; ++cc)
                for (int c2 = 0; c2 < d2; ++c2) {
                  if (b == b2 && c2 > cc) continue;  // lower part of the diagonal block; mirrored below
                  const double* z = Za[cc];
                  const double* z2 = Zb[c2];
                  double sacc = 0.0;
                  for (int aa = 0; aa < 3; ++aa)
                    for (int bb = 0; bb < 3; ++bb) sacc += z[aa] * Hb3[aa][bb] * z2[bb];
                  const int gi = s_off[b] + cc, gj = s_off[b2] + c2;  // gi >= gj
                  Mm[midx(gi, gj, nblk_r)] = sacc;
                  if ((gi >> 2) == (gj >> 2)) Mm[midx(gj, gi, nblk_r)] = sacc;
                }
            }
            G.sync();
            if (nr > 0) {
              for (int bj = gtid; bj < nblk_r; bj += GT) {
                const int o = blkoff(bj, bj, nblk_r);
                for (int bi = bj; bi < nblk_r; ++bi) s_tb[o + bi - bj] = (uint16_t)(bi | (bj << 8));
              }
              G.sync();
            }
          }
          m_is_h = false;
          bool fact_ok = true;
          if (nr > 0) {
            fact_ok = chol_bc4<W>(G, Mm, nblk_r, s_tb, s_tv, s_exch);  // forward pass fused
            if (fact_ok) chol_bwd_bc4<W>(G, Mm, nblk_r, s_tv, s_exch);
          }
          if (!none_active && nr > 0) {  // restore the tile table of the full system
            for (int bj = gtid; bj < nblk; bj += GT) {
              const int o = blkoff(bj, bj, nblk);
              for (int bi = bj; bi < nblk; ++bi) s_tb[o + bi - bj] = (uint16_t)(bi | (bj << 8));
            }
          }
          if (!fact_ok) break;
          for (int b = gtid; b < nb; b += GT) {
            double f[3] = {s_f0[3 * b], s_f0[3 * b + 1], s_f0[3 * b + 2]};
            if (none_active) {
              for (int q = 0; q < 3; ++q) f[q] = s_tv[3 * b + q];
            } else {
              for (int cc = 0; cc < 3 - s_rk[b]; ++cc) {
                const double tv = s_tv[s_off[b] + cc];
                for (int q = 0; q < 3; ++q) f[q] += __ldcg(g_Zt + 9 * b + 3 * cc + q) * tv;
              }
            }
            for (int q = 0; q < 3; ++q) s_up[3 * b + q] = f[q];
          }
          if (gtid < n4 - n) s_up[n + gtid] = 0.0;
          copy_mat<W>(G, Mm, Hm, matd);
          m_is_h = true;
          G.sync();
          symv_bc4<W>(G, Mm, n4, nblk, s_up, s_rhs);
          // multipliers, verification, correction; when the pass verifies, the same loop runs a
          // second time to commit the multipliers (nothing is stored per row in between)
          bool good = false;
          for (int commit = 0; commit < 2; ++commit) {
            bool okm = true, changed = false, any_act2 = false;
            for (int b = gtid; b < nb; b += GT) {
              const double mub = cfg.mu[s_blk_i[b]];
              const double ubxy = kFricUb * s_ce[b], ubz = mass * kGrav * (double)L * s_ce[b];
              double rb[3], y[5], ll[5] = {0, 0, 0, 0, 0}, lu[5] = {0, 0, 0, 0, 0};
              for (int q = 0; q < 3; ++q) rb[q] = s_rhs[3 * b + q] + s_g[3 * b + q];
              if (none_active) {
                okm = okm && fmax(fabs(rb[0]), fmax(fabs(rb[1]), fabs(rb[2]))) <= 1e-9 * gs;
              } else {
                double Nrm[10][3], lam[10];
                int idx[10], k = 0;
                for (int q = 0; q < 5; ++q) if (s_actl[5 * b + q]) { row_vec(mub, q, Nrm[k]); idx[k++] = q; }
                for (int q = 0; q < 5; ++q) if (s_actu[5 * b + q]) {
                  row_vec(mub, q, Nrm[k]);
                  Nrm[k][0] = -Nrm[k][0]; Nrm[k][1] = -Nrm[k][1]; Nrm[k][2] = -Nrm[k][2];
                  idx[k++] = 5 + q;
                }
                okm = block_multipliers(k, Nrm, rb, 1e-9 * gs, lam) && okm;
                for (int sI = 0; sI < k; ++sI) { if (idx[sI] < 5) ll[idx[sI]] = lam[sI]; else lu[idx[sI] - 5] = lam[sI]; }
              }
              if (commit) {
                for (int q = 0; q < 5; ++q) { s_zl[5 * b + q] = ll[q]; s_zu[5 * b + q] = lu[q]; }
                continue;
              }
              cmul5(mub, s_up + 3 * b, y);
              for (int q = 0; q < 5; ++q) {
                const double ub = q < 4 ? ubxy : ubz;
                const double sl = y[q], su = ub - y[q];
                const bool vl = sl < -1e-9 * us, vu = su < -1e-9 * us;
                const bool nl = ll[q] < -1e-9 * gs, nuu = lu[q] < -1e-9 * gs;
                if (vl || vu || nl || nuu) changed = true;
                const bool al = (s_actl[5 * b + q] || vl) && !nl, au = (s_actu[5 * b + q] || vu) && !nuu;
                s_actl[5 * b + q] = al; s_actu[5 * b + q] = au;
                any_act2 = any_act2 || al || au;
              }
            }
            if (commit) break;
            good = G.all(okm && !changed);
            none_active = G.all(!any_act2);  // unchanged when the pass verified (flags did not move)
            if (!good) break;
          }
          if (good) accepted = true;
        }
        if (accepted) {
          for (int t = gtid; t < n; t += GT) s_u[t] = s_up[t];
          G.sync();
          status = CMPC_STATUS_OK;
          break;
        }
        G.sync();
        // polish not accepted: rd was used as f0 scratch -> recompute the residual
        if (!m_is_h) { copy_mat<W>(G, Mm, Hm, matd); G.sync(); m_is_h = true; }
        symv_bc4<W>(G, Mm, n4, nblk, s_u, s_rd);
        for (int b = gtid; b < nb; b += GT) {
          double w[5], o[3];
          for (int q = 0; q < 5; ++q) w[q] = s_zl[5 * b + q] - s_zu[5 * b + q];
          ctmul5(cfg.mu[s_blk_i[b]], w, o);
          for (int q = 0; q < 3; ++q) s_rd[3 * b + q] += s_g[3 * b + q] - o[q];
        }
        G.sync();
      }
      if (strict && (!cfg.polish || npolish >= 3)) break;
      if (mu <= 1e-8 * cfg.tol * gs * us) break;  // far past convergence: stop before 0/0
      if (it == cfg.max_iter) break;

      // ---- M = H + C' diag(zl/sl + zu/su) C  (only the 3x3 diagonal blocks change), and the
      // affine (predictor) right-hand side  -rd + C'(rcl/sl - rcu/su)  with rc = -s z
      for (int b = gtid; b < nb; b += GT) {
        const double ubxy = kFricUb * s_ce[b], ubz = mass * kGrav * (double)L * s_ce[b];
        double sg[5], tq[5], o[3], ys[5];
        cmul5(cfg.mu[s_blk_i[b]], s_u + 3 * b, ys);
        for (int q = 0; q < 5; ++q) {
          const int t = 5 * b + q;
          const double sl = ys[q], su = (q < 4 ? ubxy : ubz) - sl;
          const double isl = fast_rcp(sl), isu = fast_rcp(su);
          sg[q] = s_zl[t] * isl + s_zu[t] * isu;
          tq[q] = s_zu[t] - s_zl[t];
        }
        const double mb = cfg.mu[s_blk_i[b]], sx = sg[0] + sg[1], sy = sg[2] + sg[3];
        const int g0 = 3 * b, g1 = g0 + 1, g2 = g0 + 2;
        Mm[midx(g0, g0, nblk)] += sx;
        Mm[midx(g1, g1, nblk)] += sy;
        Mm[midx(g2, g2, nblk)] += mb * mb * (sx + sy) + sg[4];
        Mm[midx(g2, g0, nblk)] += mb * (sg[1] - sg[0]);
        Mm[midx(g2, g1, nblk)] += mb * (sg[3] - sg[2]);
        ctmul5(mb, tq, o);
        for (int q = 0; q < 3; ++q) s_du[3 * b + q] = -s_rd[3 * b + q] + o[q];
      }
      m_is_h = false;
      G.sync();
      // factor; the predictor's forward substitution is fused into the sweep
      if (!chol_bc4<W>(G, Mm, nblk, s_tb, s_du, s_exch)) { numerical = true; break; }

      // Per-row step quantities are recomputed where needed instead of stored: with s = slack,
      // z = multiplier, cd = a_r . du:  dz_l = (rc_l - z_l cd) / s_l,  dz_u = (rc_u + z_u cd) / s_u,
      // rc = -s z  (+ sigma mu -/+ cdA dzA in the corrector, A = affine step kept in dua).
      double tmax = 0.0, sigma = 0.0;
      for (int phase = 0; phase < 2; ++phase) {
        // phase 0: affine predictor; phase 1: centred corrector (Mehrotra)
        if (phase) {
          for (int t = gtid; t < n4; t += GT) s_dua[t] = s_du[t];
          G.sync();
          for (int b = gtid; b < nb; b += GT) {
            const double mub = cfg.mu[s_blk_i[b]];
            const double ubxy = kFricUb * s_ce[b], ubz = mass * kGrav * (double)L * s_ce[b];
            double tq[5], o[3], ys[5], ya[5];
            cmul5(mub, s_u + 3 * b, ys);
            cmul5(mub, s_dua + 3 * b, ya);
            for (int q = 0; q < 5; ++q) {
              const int t = 5 * b + q;
              const double sl = ys[q], su = (q < 4 ? ubxy : ubz) - sl, zl = s_zl[t], zu = s_zu[t];
              const double isl = fast_rcp(sl), isu = fast_rcp(su);
              const double dla = (-sl * zl - zl * ya[q]) * isl, dua_ = (-su * zu + zu * ya[q]) * isu;
              const double rcl = -sl * zl + sigma * mu - ya[q] * dla;
              const double rcu = -su * zu + sigma * mu + ya[q] * dua_;
              tq[q] = rcl * isl - rcu * isu;
            }
            ctmul5(mub, tq, o);
            for (int q = 0; q < 3; ++q) s_du[3 * b + q] = -s_rd[3 * b + q] + o[q];
          }
          G.sync();
          chol_fwd_bc4<W>(G, Mm, nblk, s_du, s_exch);
        }
        chol_bwd_bc4<W>(G, Mm, nblk, s_du, s_exch);
        // step to the boundary: alpha_max = 1 / max_i(-ds_i/s_i, -dz_i/z_i)
        double tloc = 0.0;
        for (int b = gtid; b < nb; b += GT) {
          const double mub = cfg.mu[s_blk_i[b]];
          const double ubxy = kFricUb * s_ce[b], ubz = mass * kGrav * (double)L * s_ce[b];
          double ys[5], yd[5], ya[5];
          cmul5(mub, s_u + 3 * b, ys);
          cmul5(mub, s_du + 3 * b, yd);
          if (phase) cmul5(mub, s_dua + 3 * b, ya);
          for (int q = 0; q < 5; ++q) {
            const int t = 5 * b + q;
            const double sl = ys[q], su = (q < 4 ? ubxy : ubz) - sl, zl = s_zl[t], zu = s_zu[t];
            const double isl = fast_rcp(sl), isu = fast_rcp(su);
            double rcl = -sl * zl, rcu = -su * zu;
            if (phase) {
              const double dla = (rcl - zl * ya[q]) * isl, dua_ = (rcu + zu * ya[q]) * isu;
              rcl += sigma * mu - ya[q] * dla; rcu += sigma * mu + ya[q] * dua_;
            }
            const double cd = yd[q];
            const double dl = (rcl - zl * cd) * isl;
            const double du_ = (rcu + zu * cd) * isu;
            tloc = fmax(tloc, fmax(-cd * isl, cd * isu));
            tloc = fmax(tloc, fmax(-dl * fast_rcp(zl), -du_ * fast_rcp(zu)));
          }
        }
        tmax = G.max(tloc);
        if (!phase) {
          const double alpha = tmax > 1.0 ? 1.0 / tmax : 1.0;
          double ga = 0.0;
          for (int b = gtid; b < nb; b += GT) {
            const double mub = cfg.mu[s_blk_i[b]];
            const double ubxy = kFricUb * s_ce[b], ubz = mass * kGrav * (double)L * s_ce[b];
            double ys[5], yd[5];
            cmul5(mub, s_u + 3 * b, ys);
            cmul5(mub, s_du + 3 * b, yd);
            for (int q = 0; q < 5; ++q) {
              const int t = 5 * b + q;
              const double sl = ys[q], su = (q < 4 ? ubxy : ubz) - sl, zl = s_zl[t], zu = s_zu[t];
              const double cd = yd[q];
              const double dl = (-sl * zl - zl * cd) * fast_rcp(sl), du_ = (-su * zu + zu * cd) * fast_rcp(su);
              ga += (sl + alpha * cd) * (zl + alpha * dl) + (su - alpha * cd) * (zu + alpha * du_);
            }
          }
          ga = G.sum(ga);
          const double ratio = ga / gap;
          sigma = ratio * ratio * ratio;
        }
      }
      // fraction to the boundary tau -> 1 as the gap closes (superlinear tail)
      const double tau = fmax(0.995, 1.0 - mu / (gs * us));
      const double alpha = fmin(1.0, tau / fmax(tmax, 1e-300));
      bool fin = true;
      for (int b = gtid; b < nb; b += GT) {
        const double mub = cfg.mu[s_blk_i[b]];
        const double ubxy = kFricUb * s_ce[b], ubz = mass * kGrav * (double)L * s_ce[b];
        double ys[5], yd[5], ya[5];
        cmul5(mub, s_u + 3 * b, ys);       // slacks at the current point (before the update)
        cmul5(mub, s_du + 3 * b, yd);
        cmul5(mub, s_dua + 3 * b, ya);
        for (int q = 0; q < 5; ++q) {
          const int t = 5 * b + q;
          const double sl = ys[q], su = (q < 4 ? ubxy : ubz) - sl, zl = s_zl[t], zu = s_zu[t];
          const double isl = fast_rcp(sl), isu = fast_rcp(su);
          const double dla = (-sl * zl - zl * ya[q]) * isl, dua_ = (-su * zu + zu * ya[q]) * isu;
          const double rcl = -sl * zl + sigma * mu - ya[q] * dla;
          const double rcu = -su * zu + sigma * mu + ya[q] * dua_;
          s_zl[t] = zl + alpha * (rcl - zl * yd[q]) * isl;
          s_zu[t] = zu + alpha * (rcu + zu * yd[q]) * isu;
        }
        for (int q = 0; q < 3; ++q) {
          const double v = s_u[3 * b + q] + alpha * s_du[3 * b + q];
          s_u[3 * b + q] = v; fin = fin && isfinite(v);
        }
      }
      fin = G.all(fin);
      if (!fin) { numerical = true; break; }
    }
    if (numerical) status = CMPC_STATUS_NUMERICAL;
    else if (status != CMPC_STATUS_OK) status = ipm_ok ? CMPC_STATUS_OK_IPM : CMPC_STATUS_MAX_ITER;

    // ---- outputs
    if (!numerical) {
      // scaled KKT residual (same definition as the oracle)
      if (status != CMPC_STATUS_OK) {  // an accepted polish left H u in s_rhs already
        if (!m_is_h) { copy_mat<W>(G, Mm, Hm, matd); G.sync(); m_is_h = true; }
        symv_bc4<W>(G, Mm, n4, nblk, s_u, s_rhs);
      }
      double stat = 0.0, umax = 0.0, prim = 0.0, dual = 0.0, comp = 0.0;
      for (int b = gtid; b < nb; b += GT) {
        const double mub = cfg.mu[s_blk_i[b]];
        const double ubxy = kFricUb * s_ce[b], ubz = mass * kGrav * (double)L * s_ce[b];
        double w[5], o[3], y[5];
        for (int q = 0; q < 5; ++q) w[q] = s_zl[5 * b + q] - s_zu[5 * b + q];
        ctmul5(mub, w, o);
        cmul5(mub, s_u + 3 * b, y);
        for (int q = 0; q < 3; ++q) {
          stat = fmax(stat, fabs(s_rhs[3 * b + q] + s_g[3 * b + q] - o[q]));
          umax = fmax(umax, fabs(s_u[3 * b + q]));
        }
        for (int q = 0; q < 5; ++q) {
          const double ub = q < 4 ? ubxy : ubz;
          const double sl = y[q], su = ub - y[q], zl = s_zl[5 * b + q], zu = s_zu[5 * b + q];
          prim = fmax(prim, fmax(-sl, -su));
          dual = fmax(dual, fmax(-zl, -zu));
          comp = fmax(comp, fmax(fabs(zl * sl), fabs(zu * su)));
        }
      }
      stat = G.max(stat);
      umax = G.max(umax);
      prim = G.max(prim);
      dual = G.max(dual);
      comp = G.max(comp);
      const double usf = 1.0 + umax;
      const double kkt = fmax(fmax(stat / gs, prim / usf), fmax(dual / gs, comp / (gs * usf)));
      // reported active set. Polished: the rows with zero slack at the KKT point (primal
      // definition, unique because the optimum is unique -- the polish's working set can omit
      // redundant rows at the degenerate apex f = 0). Otherwise: the IPM guess.
      for (int b = gtid; b < nb; b += GT) {
        const double ubxy = kFricUb * s_ce[b], ubz = mass * kGrav * (double)L * s_ce[b];
        double ys[5];
        cmul5(cfg.mu[s_blk_i[b]], s_u + 3 * b, ys);
        for (int q = 0; q < 5; ++q) {
          const int t = 5 * b + q;
          const double sl = ys[q], su = (q < 4 ? ubxy : ubz) - sl;
          if (status == CMPC_STATUS_OK) { s_actl[t] = sl <= 1e-9 * usf; s_actu[t] = su <= 1e-9 * usf; }
          else { s_actl[t] = s_zl[t] * usf > sl * gs; s_actu[t] = s_zu[t] * usf > su * gs; }
        }
      }
      G.sync();
      // forces in the reference's per-leg order [L][N][3] (CentroidalMPC.cpp:270)
      for (int t = gtid; t < nf; t += GT) {
        const int i = t / (3 * N), j = (t % (3 * N)) / 3, q = t % 3;
        const int b = s_blk_of[j * L + i];
        args.forces[(size_t)inst * nf + t] = b < 0 ? 0.0 : s_u[3 * b + q];
      }
      if (args.lam) {
        for (int t = gtid; t < 2 * mfull; t += GT) {
          const int side = t / mfull, rem = t % mfull, ji = rem / 5, q = rem % 5;
          const int b = s_blk_of[ji];
          args.lam[(size_t)inst * 2 * mfull + t] = b < 0 ? 0.0 : (side ? s_zu[5 * b + q] : s_zl[5 * b + q]);
        }
      }
      if (args.active) {
        for (int t = gtid; t < nbfull; t += GT) {
          const int b = s_blk_of[t];
          uint16_t a = 0x8000;
          if (b >= 0) {
            a = 0;
            for (int q = 0; q < 5; ++q) a |= (uint16_t)((s_actl[5 * b + q] ? 1 : 0) << q | (s_actu[5 * b + q] ? 1 : 0) << (5 + q));
          }
          args.active[(size_t)inst * nbfull + t] = a;
        }
      }
      if (gtid == 0) {
        args.status[inst] = status;
        if (args.iters) args.iters[inst] = it;
        if (args.kkt) args.kkt[inst] = kkt;
      }
    } else {
      for (int t = gtid; t < nf; t += GT) args.forces[(size_t)inst * nf + t] = 0.0;
      if (args.lam) for (int t = gtid; t < 2 * mfull; t += GT) args.lam[(size_t)inst * 2 * mfull + t] = 0.0;
      if (args.active) for (int t = gtid; t < nbfull; t += GT) args.active[(size_t)inst * nbfull + t] = 0;
      if (gtid == 0) {
        args.status[inst] = status;
        if (args.iters) args.iters[inst] = it;
        if (args.kkt) args.kkt[inst] = 0.0;
      }
    }
    G.sync();
  }
}

// ------------------------------------------------------------------ presolve kernel
// Most ticks of a legged MPC have no friction or force-limit row active (the reference's weights make
// force tracking dominate), and then the optimum is the unconstrained minimiser -H^-1 g.  This kernel
// settles exactly those instances with ONE Cholesky of H: build, factor (forward substitution fused),
// back-substitute, then verify on the original H -- stationarity |H u + g| <= 1e-9 gs and every row
// of 0 <= F f <= ub satisfied to -1e-9 us, the polish's own acceptance test with an empty working
// set -- and write the outputs (status OK, iters 0, multipliers 0).  Anything else (a violated row, a
// warm-start guess with active rows, a failed pivot) is appended to fail_perm and goes through the
// interior-point kernel.  Per group it needs the matrix and three vectors only, so more instances are
// resident per SM than in the IPM kernel.  Shared-memory plan: make_pre_plan (fields of SmemPlan reused:
// u = the solution, rhs = H u; rd..tv host eq/qz during the build).
__host__ __device__ inline SmemPlan make_pre_plan(int N, int L, int W, int nbmax, int n4max) {
  SmemPlan p;
  int o = 0;
  auto take = [&](int cnt) { int r = o; o += (cnt + 1) & ~1; return r; };
  p.ce = take(nbmax);
  p.dua = take(nbmax);      // desired fz
  p.du = take(3 * nbmax);   // lever arms
  p.g = take(n4max);
  const int nv = (2 * n4max >= 10 * N + 2) ? n4max : (10 * N + 2 + 1) / 2;
  p.u = take(nv); p.rhs = take(nv);  // eq[9N], qz[N] live here during the build
  p.rd = p.u; p.tv = p.rhs;
  p.zl = p.zu = 0;
  p.red = take(W > 1 ? 3 * W : 2);
  p.exch = take(8);
  const int nbytes = 16 + 2 * nbmax + 2 * bc4_tiles(n4max) + 2 + 2 * nbmax + N * L;
  p.ints = take((nbytes + 7) / 8);
  o = (o + 1) & ~1;
  p.Mm = o;
  o += mat_region_doubles(N, L, n4max);
  p.total = (o + 1) & ~1;
  return p;
}

template <int W>
__global__ void __launch_bounds__(W == 1 ? 448 : 256) cmpc_presolve_kernel(const DevConfig cfg, const SolveArgs args) {
  extern __shared__ __align__(128) double smem[];
  constexpr int GT = Group<W>::GT;
  const int N = cfg.N, L = cfg.L;
  const int nf = 3 * L * N, nbfull = L * N, mfull = 5 * nbfull;
  const int nbmax = args.nbmax;
  const SmemPlan& P = args.plan;
  Group<W> G;
  G.gtid = threadIdx.x % GT;
  G.gid = threadIdx.x / GT;
  const int gtid = G.gtid;
  double* base = smem + (size_t)G.gid * P.total;
  G.red = base + P.red;
  double* s_exch = base + P.exch;
  double* s_g = base + P.g;
  double* s_x = base + P.u;
  double* s_hx = base + P.rhs;
  int* s_misc = reinterpret_cast<int*>(base + P.ints);  // [0]=nb, [1]=invalid, [2]=work slot
  uint16_t* s_tb = reinterpret_cast<uint16_t*>(s_misc + 4);
  uint8_t* s_blk_j = reinterpret_cast<uint8_t*>(s_tb + bc4_tiles(args.n4max) + (bc4_tiles(args.n4max) & 1));
  uint8_t* s_blk_i = s_blk_j + nbmax;
  int8_t* s_blk_of = reinterpret_cast<int8_t*>(s_blk_i + nbmax);
  double* Mm = base + P.Mm;
  double* Hm = args.scratch + (size_t)(blockIdx.x * args.groups + G.gid) * args.scratch_per_group;
  BuildView V;
  V.Mm = Mm; V.ce = base + P.ce; V.fz = base + P.dua; V.arm = base + P.du; V.eq = s_x; V.qz = s_x + 9 * N; V.g = s_g;
  V.misc = s_misc; V.tb = s_tb; V.blk_j = s_blk_j; V.blk_i = s_blk_i; V.blk_of = s_blk_of;
  const double mass = cfg.mass;
  const int count = args.count ? *args.count : args.count_imm;

  while (true) {
    int slot = 0;
    if (gtid == 0) slot = atomicAdd(args.work, 1);
    slot = G.bcast0(slot, s_misc + 2);
    if (slot >= count) break;
    const int inst = args.perm ? args.perm[slot] : slot;
    bool defer = false;
    if (args.warm_active) {  // a warm-start guess with active rows belongs to the IPM kernel's polish
      bool any = false;
      const uint16_t* wa = args.warm_active + (size_t)inst * nbfull;
      for (int t = gtid; t < nbfull; t += GT) { const unsigned a = wa[t]; any = any || (!(a & 0x8000u) && (a & 0x3ffu)); }
      defer = !G.all(!any);
    }
    bool finite = true;
    int nb = 0, n = 0, nblk = 0, n4 = 0, matd = 0;
    if (!defer) {
      finite = stage_inputs<W>(G, cfg, args, inst, V);
      nb = s_misc[0];
      n = 3 * nb; nblk = (n + 3) >> 2; n4 = nblk << 2;
      matd = ((nblk * (nblk + 1)) >> 1) * kTS;
      const bool invalid = s_misc[1] != 0;
      if (!finite || invalid) {
        for (int t = gtid; t < nf; t += GT) args.forces[(size_t)inst * nf + t] = 0.0;
        if (args.lam) for (int t = gtid; t < 2 * mfull; t += GT) args.lam[(size_t)inst * 2 * mfull + t] = 0.0;
        if (args.active) for (int t = gtid; t < nbfull; t += GT) args.active[(size_t)inst * nbfull + t] = 0;
        if (gtid == 0) {
          args.status[inst] = !finite ? CMPC_STATUS_NUMERICAL : CMPC_STATUS_INVALID_TABLE;
          if (args.iters) args.iters[inst] = 0;
          if (args.kkt) args.kkt[inst] = 0.0;
        }
        G.sync();
        continue;
      }
      build_qp<W>(G, cfg, V, Mm, nb);
      store_mat<W>(G, Hm, Mm, matd);  // the verification below needs H again: keep it in the L2 slab
      for (int t = gtid; t < n4; t += GT) s_x[t] = t < n ? -s_g[t] : 0.0;  // (eq/qz are dead after the build)
      G.sync();
      bool ok = chol_bc4<W>(G, Mm, nblk, s_tb, s_x, s_exch);
      if (ok) {
        chol_bwd_bc4<W>(G, Mm, nblk, s_x, s_exch);
        copy_mat<W>(G, Mm, Hm, matd);
        G.sync();
        symv_bc4<W>(G, Mm, n4, nblk, s_x, s_hx);
      }
      defer = !ok;
    }
    double gs = 1.0, usf = 1.0, stat = 0.0, prim = 0.0;
    if (!defer) {
      double gmax = 0.0, umax = 0.0;
      bool fin = true;
      for (int t = gtid; t < n; t += GT) {
        gmax = fmax(gmax, fabs(s_g[t])); umax = fmax(umax, fabs(s_x[t]));
        stat = fmax(stat, fabs(s_hx[t] + s_g[t]));
        fin = fin && isfinite(s_x[t]) && isfinite(s_hx[t]);
      }
      for (int b = gtid; b < nb; b += GT) {
        const double ce = V.ce[b];
        const double ubxy = kFricUb * ce, ubz = mass * kGrav * (double)L * ce;
        double y[5];
        cmul5(cfg.mu[s_blk_i[b]], s_x + 3 * b, y);
        for (int q = 0; q < 5; ++q) prim = fmax(prim, fmax(-y[q], y[q] - (q < 4 ? ubxy : ubz)));
      }
      gmax = G.max(gmax); umax = G.max(umax); stat = G.max(stat);
      gs = 1.0 + gmax; usf = 1.0 + umax;
      prim = G.max(prim);
      defer = !(stat <= 1e-9 * gs && prim <= 1e-9 * usf);
      defer = !G.all(!defer && fin);  // fmax drops NaNs: a non-finite candidate is caught here
    }
    if (defer) {
      if (gtid == 0) args.fail_perm[atomicAdd(args.fail_count, 1)] = inst;
      G.sync();
      continue;
    }
    // ---- outputs of a verified unconstrained optimum (same conventions as the IPM kernel)
    for (int t = gtid; t < nf; t += GT) {
      const int i = t / (3 * N), j = (t % (3 * N)) / 3, q = t % 3;
      const int b = s_blk_of[j * L + i];
      args.forces[(size_t)inst * nf + t] = b < 0 ? 0.0 : s_x[3 * b + q];
    }
    if (args.lam) for (int t = gtid; t < 2 * mfull; t += GT) args.lam[(size_t)inst * 2 * mfull + t] = 0.0;
    if (args.active) {
      for (int t = gtid; t < nbfull; t += GT) {
        const int b = s_blk_of[t];
        uint16_t a = 0x8000;
        if (b >= 0) {
          const double ce = V.ce[b];
          const double ubxy = kFricUb * ce, ubz = mass * kGrav * (double)L * ce;
          double y[5];
          cmul5(cfg.mu[s_blk_i[b]], s_x + 3 * b, y);
          a = 0;
          for (int q = 0; q < 5; ++q)
            a |= (uint16_t)((y[q] <= 1e-9 * usf ? 1 : 0) << q | (((q < 4 ? ubxy : ubz) - y[q]) <= 1e-9 * usf ? 1 : 0) << (5 + q));
        }
        args.active[(size_t)inst * nbfull + t] = a;
      }
    }
    if (gtid == 0) {
      args.status[inst] = CMPC_STATUS_OK;
      if (args.iters) args.iters[inst] = 0;
      if (args.kkt) args.kkt[inst] = fmax(stat / gs, fmax(prim, 0.0) / usf);
    }
    G.sync();
  }
}

}  // namespace cmpc
