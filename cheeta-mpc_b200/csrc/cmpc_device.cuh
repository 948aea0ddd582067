// cmpc_device.cuh -- sm_100a device code of the batched centroidal-MPC condensed-QP solver.
//
// One CTA per MPC instance (persistent loop over the batch).  Everything an instance
// needs between its 2.3 KB of inputs and its 1 KB of outputs lives on-chip:
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

struct DevConfig {
  double mass, dt;
  double mu[kMaxLegs];
  double w[CMPC_NUM_WEIGHTS];
  double tol;
  int L, N, zoh, max_iter, polish;
};

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
  double* Hout;  // build-export mode only
  double* gout;
  double* scratch;          // global scratch for matrices that do not fit in shared memory
  size_t scratch_per_cta;   // doubles
  int B;
  int h_in_smem, m_in_smem;
  int mat_doubles;          // BC4 size for the launch's worst-case n
};

// ------------------------------------------------------------------ BC4 layout
__device__ __forceinline__ int blkoff(int bi, int bj, int nblk) {
  return bj * nblk - ((bj * (bj - 1)) >> 1) + (bi - bj);
}
__device__ __forceinline__ int midx(int i, int j, int nblk) {  // requires i>>2 >= j>>2
  return (blkoff(i >> 2, j >> 2, nblk) << 4) + ((i & 3) << 2) + (j & 3);
}
__host__ __device__ inline int bc4_doubles(int n) {
  int nblk = (n + 3) >> 2;
  return ((nblk * (nblk + 1)) >> 1) << 4;
}

// ------------------------------------------------------------------ reductions
template <int NT>
__device__ __forceinline__ double block_max(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = red[0];
#pragma unroll
  for (int k = 1; k < NT / 32; ++k) r = fmax(r, red[k]);
  return r;
}
template <int NT>
__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = red[0];
#pragma unroll
  for (int k = 1; k < NT / 32; ++k) r += red[k];
  return r;
}
// two maxima and one sum in a single pass (fewer barriers)
template <int NT>
__device__ __forceinline__ void block_max2_sum(double& a, double& b, double& s, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a = fmax(a, __shfl_xor_sync(0xffffffffu, a, o));
    b = fmax(b, __shfl_xor_sync(0xffffffffu, b, o));
    s += __shfl_xor_sync(0xffffffffu, s, o);
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) {
    red[threadIdx.x >> 5] = a; red[32 + (threadIdx.x >> 5)] = b; red[64 + (threadIdx.x >> 5)] = s;
  }
  __syncthreads();
  a = red[0]; b = red[32]; s = red[64];
#pragma unroll
  for (int k = 1; k < NT / 32; ++k) { a = fmax(a, red[k]); b = fmax(b, red[32 + k]); s += red[64 + k]; }
}

// ------------------------------------------------------------------ 4x4 tile kernels
// Cholesky of a 4x4 SPD tile (row-major, lower part read). Writes l (lower, row-major, upper
// zeroed) and the inverse diagonal. Returns false if a pivot is not positive.
__device__ __forceinline__ bool potrf4(const double* a, double* l, double* dinv) {
  double a00 = a[0], a10 = a[4], a11 = a[5], a20 = a[8], a21 = a[9], a22 = a[10];
  double a30 = a[12], a31 = a[13], a32 = a[14], a33 = a[15];
  bool ok = a00 > 0.0;
  double i0 = rsqrt(a00);
  double l00 = a00 * i0, l10 = a10 * i0, l20 = a20 * i0, l30 = a30 * i0;
  double d1 = a11 - l10 * l10;
  ok = ok && d1 > 0.0;
  double i1 = rsqrt(d1);
  double l11 = d1 * i1, l21 = (a21 - l20 * l10) * i1, l31 = (a31 - l30 * l10) * i1;
  double d2 = a22 - l20 * l20 - l21 * l21;
  ok = ok && d2 > 0.0;
  double i2 = rsqrt(d2);
  double l22 = d2 * i2, l32 = (a32 - l30 * l20 - l31 * l21) * i2;
  double d3 = a33 - l30 * l30 - l31 * l31 - l32 * l32;
  ok = ok && d3 > 0.0;
  double i3 = rsqrt(d3);
  double l33 = d3 * i3;
  l[0] = l00; l[1] = 0; l[2] = 0; l[3] = 0;
  l[4] = l10; l[5] = l11; l[6] = 0; l[7] = 0;
  l[8] = l20; l[9] = l21; l[10] = l22; l[11] = 0;
  l[12] = l30; l[13] = l31; l[14] = l32; l[15] = l33;
  dinv[0] = i0; dinv[1] = i1; dinv[2] = i2; dinv[3] = i3;
  return ok;
}

// Tiled right-looking Cholesky in BC4 layout, in place. All threads must call it.
// Returns false (uniformly) on a non-positive pivot.
template <int NT>
__device__ bool chol_bc4(double* M, int nblk) {
  const int tid = threadIdx.x;
  bool ok = true;
  for (int kb = 0; kb < nblk; ++kb) {
    // POTRF (redundant in every participating thread) + TRSM of the panel below
    // (the factored diagonal tile is stored after the barrier: other threads still read A_kk)
    const int nrows = nblk - kb;
    double l[16], dinv[4];
    if (tid < nrows) {
      const double* Akk = M + ((size_t)blkoff(kb, kb, nblk) << 4);
      ok = potrf4(Akk, l, dinv) && ok;
      for (int bi = kb + tid; bi < nblk; bi += NT) {
        if (bi == kb) continue;
        double* A = M + ((size_t)blkoff(bi, kb, nblk) << 4);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          double x0 = A[4 * r] * dinv[0];
          double x1 = (A[4 * r + 1] - x0 * l[4]) * dinv[1];
          double x2 = (A[4 * r + 2] - x0 * l[8] - x1 * l[9]) * dinv[2];
          double x3 = (A[4 * r + 3] - x0 * l[12] - x1 * l[13] - x2 * l[14]) * dinv[3];
          A[4 * r] = x0; A[4 * r + 1] = x1; A[4 * r + 2] = x2; A[4 * r + 3] = x3;
        }
      }
    }
    ok = __syncthreads_and(ok);
    if (!ok) return false;
    if (tid == 0) {
      double* Akk = M + ((size_t)blkoff(kb, kb, nblk) << 4);
#pragma unroll
      for (int q = 0; q < 16; ++q) Akk[q] = l[q];
    }
    // trailing update: tiles (bi, bj), kb < bj <= bi < nblk
    const int r = nblk - kb - 1;
    const int cnt = (r * (r + 1)) >> 1;
    for (int idx = tid; idx < cnt; idx += NT) {
      int a = (int)((sqrtf(8.0f * (float)idx + 1.0f) - 1.0f) * 0.5f);
      while (((a + 1) * (a + 2)) >> 1 <= idx) ++a;
      while ((a * (a + 1)) >> 1 > idx) --a;
      int c = idx - ((a * (a + 1)) >> 1);
      const int bi = kb + 1 + a, bj = kb + 1 + c;
      const double* Li = M + ((size_t)blkoff(bi, kb, nblk) << 4);
      const double* Lj = M + ((size_t)blkoff(bj, kb, nblk) << 4);
      double* C = M + ((size_t)blkoff(bi, bj, nblk) << 4);
      double li[16], lj[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) { li[q] = Li[q]; lj[q] = Lj[q]; }
#pragma unroll
      for (int rr = 0; rr < 4; ++rr)
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          double s = C[4 * rr + cc];
#pragma unroll
          for (int k = 0; k < 4; ++k) s -= li[4 * rr + k] * lj[4 * cc + k];
          C[4 * rr + cc] = s;
        }
    }
    __syncthreads();
  }
  return true;
}

// Solve L L' x = b (x: n4 doubles in shared memory, overwritten by the solution; tmp: n4
// doubles of scratch).  Forward pass accumulates residuals in x and writes y to tmp; the
// backward pass accumulates in tmp and writes the solution to x -- no element is read and
// written by different threads inside one step, so one barrier per block step suffices.
template <int NT>
__device__ void chol_solve_bc4(const double* M, int nblk, double* x, double* tmp) {
  const int tid = threadIdx.x;
  for (int kb = 0; kb < nblk; ++kb) {
    if (tid < nblk - kb) {
      const double* l = M + ((size_t)blkoff(kb, kb, nblk) << 4);
      const double b0 = x[4 * kb], b1 = x[4 * kb + 1], b2 = x[4 * kb + 2], b3 = x[4 * kb + 3];
      const double y0 = b0 / l[0];
      const double y1 = (b1 - l[4] * y0) / l[5];
      const double y2 = (b2 - l[8] * y0 - l[9] * y1) / l[10];
      const double y3 = (b3 - l[12] * y0 - l[13] * y1 - l[14] * y2) / l[15];
      for (int bi = kb + tid; bi < nblk; bi += NT) {
        if (bi == kb) {
          tmp[4 * kb] = y0; tmp[4 * kb + 1] = y1; tmp[4 * kb + 2] = y2; tmp[4 * kb + 3] = y3;
        } else {
          const double* A = M + ((size_t)blkoff(bi, kb, nblk) << 4);
#pragma unroll
          for (int r = 0; r < 4; ++r)
            x[4 * bi + r] -= A[4 * r] * y0 + A[4 * r + 1] * y1 + A[4 * r + 2] * y2 + A[4 * r + 3] * y3;
        }
      }
    }
    __syncthreads();
  }
  for (int kb = nblk - 1; kb >= 0; --kb) {
    if (tid <= kb) {
      const double* l = M + ((size_t)blkoff(kb, kb, nblk) << 4);
      const double y0 = tmp[4 * kb], y1 = tmp[4 * kb + 1], y2 = tmp[4 * kb + 2], y3 = tmp[4 * kb + 3];
      const double x3 = y3 / l[15];
      const double x2 = (y2 - l[14] * x3) / l[10];
      const double x1 = (y1 - l[9] * x2 - l[13] * x3) / l[5];
      const double x0 = (y0 - l[4] * x1 - l[8] * x2 - l[12] * x3) / l[0];
      for (int bj = kb - tid; bj >= 0; bj -= NT) {
        if (bj == kb) {
          x[4 * kb] = x0; x[4 * kb + 1] = x1; x[4 * kb + 2] = x2; x[4 * kb + 3] = x3;
        } else {
          const double* A = M + ((size_t)blkoff(kb, bj, nblk) << 4);  // L_kj
#pragma unroll
          for (int c = 0; c < 4; ++c)
            tmp[4 * bj + c] -= A[c] * x0 + A[4 + c] * x1 + A[8 + c] * x2 + A[12 + c] * x3;
        }
      }
    }
    __syncthreads();
  }
}

// y = H x for symmetric H in BC4 layout (diagonal tiles hold both triangles).
// tpr lanes cooperate on one row. All threads call; ends with a barrier.
template <int NT>
__device__ void symv_bc4(const double* H, int n4, int nblk, const double* x, double* y) {
  const int tid = threadIdx.x;
  int tpr = 1;
  while (tpr * 2 * n4 <= NT && tpr < 8) tpr *= 2;
  const int rpp = NT / tpr, sub = tid % tpr;
  for (int base = 0; base < n4; base += rpp) {
    const int row = base + tid / tpr;
    double s = 0.0;
    if (row < n4) {
      const int bi = row >> 2, ri = row & 3;
      for (int bj = sub; bj < nblk; bj += tpr) {
        if (bj <= bi) {
          const double* A = H + ((size_t)blkoff(bi, bj, nblk) << 4) + 4 * ri;
          s += A[0] * x[4 * bj] + A[1] * x[4 * bj + 1] + A[2] * x[4 * bj + 2] + A[3] * x[4 * bj + 3];
        } else {
          const double* A = H + ((size_t)blkoff(bj, bi, nblk) << 4) + ri;
          s += A[0] * x[4 * bj] + A[4] * x[4 * bj + 1] + A[8] * x[4 * bj + 2] + A[12] * x[4 * bj + 3];
        }
      }
    }
    for (int o = tpr >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (row < n4 && sub == 0) y[row] = s;
  }
  __syncthreads();
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
__device__ int block_nullspace(int k, const double (*A)[3], const double* b, double* f0, double (*Z)[3], bool* ok) {
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
__device__ double small_lsq(int k, const double (*S)[3], const double* rb, double* lam) {
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
__device__ bool block_multipliers(int k, const double (*Nrm)[3], const double* rb, double tol, double* lam) {
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

// ------------------------------------------------------------------ shared-memory plan
struct SmemPlan {
  // offsets in doubles from the start of dynamic shared memory
  int in_state, in_ds, in_di, eq, qz, mu_b, ubxy, ubz, g, u, rd, rhs, du, f0, up, tv;
  int sl, su, zl, zu, cdu, dzl, dzu, Zt, red, Hm, Mm;
  int ints;  // start of int region (in doubles)
  int total; // doubles
};
__host__ __device__ inline SmemPlan make_plan(int N, int L, int h_in_smem, int m_in_smem) {
  SmemPlan p;
  const int ns = 9 + 3 * L, nds = 9 * (N + 1), ndi = L * (4 * N + 3);
  const int nbmax = L * N, nmax = ((3 * nbmax + 3) >> 2) << 2, mmax = 5 * nbmax;
  int o = 0;
  auto take = [&](int cnt) { int r = o; o += (cnt + 1) & ~1; return r; };
  p.in_state = take(ns); p.in_ds = take(nds); p.in_di = take(ndi);
  p.eq = take(9 * N); p.qz = take(N);
  p.mu_b = take(nbmax); p.ubxy = take(nbmax); p.ubz = take(nbmax);
  p.g = take(nmax); p.u = take(nmax); p.rd = take(nmax); p.rhs = take(nmax); p.du = take(nmax);
  p.f0 = take(nmax); p.up = take(nmax); p.tv = take(nmax);
  p.sl = take(mmax); p.su = take(mmax); p.zl = take(mmax); p.zu = take(mmax);
  p.cdu = take(mmax); p.dzl = take(mmax); p.dzu = take(mmax);
  p.Zt = take(9 * nbmax);
  p.red = take(96);
  // int region: blk_j, blk_i, blk_of[N*L], rk, off (ints) + act flags (bytes)
  p.ints = take((5 * nbmax + 2 * mmax / 4 + 16) / 2 + 8);
  const int mat = bc4_doubles(nmax);
  o = (o + 15) & ~15;  // 128-byte align tiles
  p.Hm = o; if (h_in_smem) o += mat;
  p.Mm = o; if (m_in_smem) o += mat;
  p.total = o;
  return p;
}

// ------------------------------------------------------------------ the fused kernel
// MODE 0: solve.  MODE 1: build-export (H, g in the full 3LN layout to global memory).
template <int NT, int MODE>
__global__ void __launch_bounds__(NT) cmpc_solve_kernel(const DevConfig cfg, const SolveArgs args) {
  extern __shared__ __align__(128) double smem[];
  const int tid = threadIdx.x;
  const int N = cfg.N, L = cfg.L, nu = 3 * L;
  const int ns = 9 + 3 * L, nds = 9 * (N + 1), ndi = L * (4 * N + 3), nf = 3 * L * N;
  const int nbmax = L * N, mmax = 5 * nbmax;
  const SmemPlan P = make_plan(N, L, args.h_in_smem, args.m_in_smem);
  double* s_state = smem + P.in_state;
  double* s_ds = smem + P.in_ds;
  double* s_di = smem + P.in_di;
  double* s_eq = smem + P.eq;
  double* s_qz = smem + P.qz;
  double* s_mu = smem + P.mu_b;
  double* s_ubxy = smem + P.ubxy;
  double* s_ubz = smem + P.ubz;
  double* s_g = smem + P.g;
  double* s_u = smem + P.u;
  double* s_rd = smem + P.rd;
  double* s_rhs = smem + P.rhs;
  double* s_du = smem + P.du;
  double* s_f0 = smem + P.f0;
  double* s_up = smem + P.up;
  double* s_tv = smem + P.tv;
  double* s_sl = smem + P.sl;
  double* s_su = smem + P.su;
  double* s_zl = smem + P.zl;
  double* s_zu = smem + P.zu;
  double* s_cdu = smem + P.cdu;
  double* s_dzl = smem + P.dzl;
  double* s_dzu = smem + P.dzu;
  double* s_Zt = smem + P.Zt;
  double* s_red = smem + P.red;
  int* s_blk_j = reinterpret_cast<int*>(smem + P.ints);
  int* s_blk_i = s_blk_j + nbmax;
  int* s_blk_of = s_blk_i + nbmax;  // [N*L] free block index or -1
  int* s_rk = s_blk_of + nbmax;
  int* s_off = s_rk + nbmax;
  int* s_misc = s_off + nbmax;      // [0]=nb, [1]=invalid, [2]=flag, [3]=nr
  unsigned char* s_actl = reinterpret_cast<unsigned char*>(s_misc + 8);
  unsigned char* s_actu = s_actl + mmax;
  double* Hm = args.h_in_smem ? smem + P.Hm : args.scratch + (size_t)blockIdx.x * args.scratch_per_cta;
  double* Mm = args.m_in_smem ? smem + P.Mm
                              : args.scratch + (size_t)blockIdx.x * args.scratch_per_cta + (args.h_in_smem ? 0 : args.mat_doubles);

  const double* s_dpos = s_ds;
  const double* s_dvel = s_ds + 3 * (N + 1);
  const double* s_dam = s_ds + 6 * (N + 1);
  const double dt = cfg.dt, mass = cfg.mass;
  const double zeta = cfg.zoh ? 0.5 : 0.0;

  for (int inst = blockIdx.x; inst < args.B; inst += gridDim.x) {
    __syncthreads();
    // ---- stage inputs (coalesced), CentroidalMPC.cpp:284-317
    bool finite = true;
    for (int t = tid; t < ns; t += NT) { double v = args.state[(size_t)inst * ns + t]; s_state[t] = v; finite = finite && isfinite(v); }
    for (int t = tid; t < nds; t += NT) { double v = args.des_state[(size_t)inst * nds + t]; s_ds[t] = v; finite = finite && isfinite(v); }
    for (int t = tid; t < ndi; t += NT) { double v = args.des_inputs[(size_t)inst * ndi + t]; s_di[t] = v; finite = finite && isfinite(v); }
    finite = __syncthreads_and(finite);

    // ---- contact table -> free blocks; validity (CentroidalMPC.cpp:328-330)
    if (tid == 0) {
      int nb = 0, invalid = 0;
      for (int j = 0; j < N; ++j) {
        double colsum = 0.0;
        for (int i = 0; i < L; ++i) colsum += s_di[i * (4 * N + 3) + j];
        if (!(colsum > 0.0)) invalid = 1;
        for (int i = 0; i < L; ++i) {
          double ce = s_di[i * (4 * N + 3) + j];
          if (ce > 0.0) {
            s_blk_j[nb] = j; s_blk_i[nb] = i; s_blk_of[j * L + i] = nb;
            s_mu[nb] = cfg.mu[i];
            s_ubxy[nb] = kFricUb * ce;                     // :183,199
            s_ubz[nb] = mass * kGrav * (double)L * ce;
            // desired normal force m g / #stance (:331-333), kept in s_up until g is built
            s_up[nb] = (colsum > 0.0) ? mass * kGrav / colsum : 0.0;
            ++nb;
          } else {
            s_blk_of[j * L + i] = -1;
          }
        }
      }
      s_misc[0] = nb; s_misc[1] = invalid;
      // zero-input rollout x_{k+1} = A x_k + d and weighted error e = Q (x - x_ref), nodes 1..N
      double c[3] = {s_state[0], s_state[1], s_state[2]};
      double v[3] = {s_state[3], s_state[4], s_state[5]};
      double gz = -kGrav;
      for (int k = 0; k < N; ++k) {
        const int node = k + 1;
        for (int a = 0; a < 3; ++a) c[a] += dt * v[a];
        if (cfg.zoh) c[2] += 0.5 * dt * dt * gz;
        v[2] += dt * gz;
        double om = (cfg.w[2] * 0.5) * exp(-(double)node) + cfg.w[2] * 0.5;  // :205
        double qz = om * om;                                                  // :210 (inside the square)
        s_qz[k] = qz;
        s_eq[9 * k + 0] = cfg.w[0] * (c[0] - s_dpos[3 * node + 0]);
        s_eq[9 * k + 1] = cfg.w[1] * (c[1] - s_dpos[3 * node + 1]);
        s_eq[9 * k + 2] = qz * (c[2] - s_dpos[3 * node + 2]);
        for (int a = 0; a < 3; ++a) {
          s_eq[9 * k + 3 + a] = cfg.w[3 + a] * (v[a] - s_dvel[3 * node + a]);
          s_eq[9 * k + 6 + a] = cfg.w[6 + a] * (s_state[6 + a] - s_dam[3 * node + a]);
        }
      }
    }
    __syncthreads();
    const int nb = s_misc[0];
    const int n = 3 * nb, nblk = (n + 3) >> 2, n4 = nblk << 2, m = 5 * nb;
    const bool invalid = s_misc[1] != 0;

    if (!finite || invalid) {
      if (MODE == 0) {
        for (int t = tid; t < nf; t += NT) args.forces[(size_t)inst * nf + t] = 0.0;
        if (args.lam) for (int t = tid; t < 2 * mmax; t += NT) args.lam[(size_t)inst * 2 * mmax + t] = 0.0;
        if (args.active) for (int t = tid; t < nbmax; t += NT) args.active[(size_t)inst * nbmax + t] = 0;
        if (tid == 0) {
          args.status[inst] = !finite ? CMPC_STATUS_NUMERICAL : CMPC_STATUS_INVALID_TABLE;
          if (args.iters) args.iters[inst] = 0;
          if (args.kkt) args.kkt[inst] = 0.0;
        }
        continue;
      }
    }

    // ---- H = 2 (Bqp' L Bqp + K) on the free variables, one thread per block pair (b >= b2).
    // Column (j,i) of Bqp at row block k >= j is A_d^{k-j} B_j =
    //   [ dt^2 (k-j+zeta) (c/m) I ; dt (c/m) I ; dt c [r]x ]   (a2/a3; Euler zeta=0, ZOH 1/2)
    {
      const int mat = bc4_doubles(n4);
      for (int t = tid; t < mat; t += NT) Hm[t] = 0.0;
      __syncthreads();
      const int npairs = (nb * (nb + 1)) >> 1;
      for (int idx = tid; idx < npairs; idx += NT) {
        int a = (int)((sqrtf(8.0f * (float)idx + 1.0f) - 1.0f) * 0.5f);
        while (((a + 1) * (a + 2)) >> 1 <= idx) ++a;
        while ((a * (a + 1)) >> 1 > idx) --a;
        const int b2 = idx - ((a * (a + 1)) >> 1), b = a;  // b >= b2  => j >= j2
        const int j = s_blk_j[b], i = s_blk_i[b], j2 = s_blk_j[b2], i2 = s_blk_i[b2];
        const double ce = s_di[i * (4 * N + 3) + j], ce2 = s_di[i2 * (4 * N + 3) + j2];
        double r[3], r2[3];
        for (int q = 0; q < 3; ++q) {
          r[q] = s_di[i * (4 * N + 3) + N + 3 * j + q] - s_dpos[3 * j + q];
          r2[q] = s_di[i2 * (4 * N + 3) + N + 3 * j2 + q] - s_dpos[3 * j2 + q];
        }
        const double cm = ce / mass, cm2 = ce2 / mass;
        // position rows: sum_k alpha_{k-j} alpha_{k-j2} Qp_k ; only the z weight depends on k
        double s0 = 0.0, sz = 0.0;
        for (int k = j; k < N; ++k) {
          double aa = ((double)(k - j) + zeta) * ((double)(k - j2) + zeta);
          s0 += aa; sz += aa * s_qz[k];
        }
        const double dt2 = dt * dt, dt4 = dt2 * dt2;
        const double cnt = (double)(N - j);
        double blk[3][3];
        // angular rows: dt^2 c c2 [r]x' diag(ql) [r2]x  summed over N-j row blocks
        const double ql0 = cfg.w[6], ql1 = cfg.w[7], ql2 = cfg.w[8];
        // [r]x = [[0,-rz,ry],[rz,0,-rx],[-ry,rx,0]];  ([r]x' Q [r2]x)_{ab} = sum_q [r]x_{qa} ql_q [r2]x_{qb}
        double X[3][3] = {{0.0, -r[2], r[1]}, {r[2], 0.0, -r[0]}, {-r[1], r[0], 0.0}};
        double Y[3][3] = {{0.0, -r2[2], r2[1]}, {r2[2], 0.0, -r2[0]}, {-r2[1], r2[0], 0.0}};
        const double ql[3] = {ql0, ql1, ql2};
        for (int aa = 0; aa < 3; ++aa)
          for (int bb = 0; bb < 3; ++bb) {
            double s = 0.0;
            for (int q = 0; q < 3; ++q) s += X[q][aa] * ql[q] * Y[q][bb];
            blk[aa][bb] = cnt * dt2 * ce * ce2 * s;
          }
        const double pos[3] = {cfg.w[0] * s0, cfg.w[1] * s0, sz};
        for (int aa = 0; aa < 3; ++aa)
          blk[aa][aa] += cm * cm2 * (dt4 * pos[aa] + cnt * dt2 * cfg.w[3 + aa]);
        // K = W_f + D' W_r D (CentroidalMPC.cpp:223-231): same leg, same component
        if (i == i2) {
          for (int aa = 0; aa < 3; ++aa) {
            const double wr = cfg.w[9 + 6 * L + 3 * i + aa];
            if (j == j2) {
              double nn = (j > 0 ? 1.0 : 0.0) + (j + 1 < N ? 1.0 : 0.0);
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
              if ((gi >> 2) >= (gj >> 2)) Hm[midx(gi, gj, nblk)] = v;
              else Hm[midx(gj, gi, nblk)] = v;
              if ((gi >> 2) == (gj >> 2)) Hm[midx(gj, gi, nblk)] = v;  // same diagonal tile: mirror
            } else if ((gi >> 2) >= (gj >> 2)) {
              // diagonal 3x3 block: all 9 (aa,bb) are visited, so both triangles of a
              // diagonal tile get written; a straddling entry lands in the lower tile only
              Hm[midx(gi, gj, nblk)] = v;
            }
          }
      }
      // padding rows (n..n4): identity
      if (tid < n4 - n) Hm[midx(n + tid, n + tid, nblk)] = 1.0;
      // ---- g = 2 Bqp' L (Aqp x0 + dqp - Xref) - 2 W_f Uref, one thread per block (adjoint sum)
      for (int b = tid; b < nb; b += NT) {
        const int j = s_blk_j[b], i = s_blk_i[b];
        const double ce = s_di[i * (4 * N + 3) + j];
        double r[3];
        for (int q = 0; q < 3; ++q) r[q] = s_di[i * (4 * N + 3) + N + 3 * j + q] - s_dpos[3 * j + q];
        double sp[3] = {0, 0, 0}, sv[3] = {0, 0, 0}, sl3[3] = {0, 0, 0};
        for (int k = j; k < N; ++k) {
          const double al = (double)(k - j) + zeta;
          for (int q = 0; q < 3; ++q) {
            sp[q] += al * s_eq[9 * k + q]; sv[q] += s_eq[9 * k + 3 + q]; sl3[q] += s_eq[9 * k + 6 + q];
          }
        }
        const double cm = ce / mass;
        // [r]x' v = v x r
        double cr[3] = {sl3[1] * r[2] - sl3[2] * r[1], sl3[2] * r[0] - sl3[0] * r[2], sl3[0] * r[1] - sl3[1] * r[0]};
        const double fzref = s_up[b];
        for (int q = 0; q < 3; ++q) {
          double gq = 2.0 * (cm * (dt * dt * sp[q] + dt * sv[q]) + dt * ce * cr[q]);
          if (q == 2) gq -= 2.0 * cfg.w[9 + 3 * L + 3 * i + 2] * fzref;
          s_g[3 * b + q] = gq;
        }
      }
      if (tid < n4 - n) s_g[n + tid] = 0.0;
      __syncthreads();
    }

    if (MODE == 1) {
      // export H, g in the full 3LN step-major layout with pinned rows/cols = identity
      const int p = nf;
      double* Ho = args.Hout + (size_t)inst * p * p;
      double* go = args.gout + (size_t)inst * p;
      for (int t = tid; t < p * p; t += NT) {
        const int a = t / p, c = t % p;
        const int ba = s_blk_of[(a / nu) * L + (a % nu) / 3], bc = s_blk_of[(c / nu) * L + (c % nu) / 3];
        double v;
        if (ba < 0 || bc < 0) v = (a == c) ? 1.0 : 0.0;
        else {
          const int gi = 3 * ba + a % 3, gj = 3 * bc + c % 3;
          v = ((gi >> 2) >= (gj >> 2)) ? Hm[midx(gi, gj, nblk)] : Hm[midx(gj, gi, nblk)];
        }
        Ho[t] = v;
      }
      for (int t = tid; t < p; t += NT) {
        const int ba = s_blk_of[(t / nu) * L + (t % nu) / 3];
        go[t] = ba < 0 ? 0.0 : s_g[3 * ba + t % 3];
      }
      if (tid == 0) args.status[inst] = !finite ? CMPC_STATUS_NUMERICAL : (invalid ? CMPC_STATUS_INVALID_TABLE : CMPC_STATUS_OK);
      continue;
    }

    // ---- strictly feasible start f = (0, 0, fz0); centred duals
    for (int b = tid; b < nb; b += NT) {
      double fz = s_up[b];
      fz = fmin(fz, 0.5 * s_ubz[b]);
      fz = fmin(fz, 0.5 * s_ubxy[b] / s_mu[b]);
      s_u[3 * b] = 0.0; s_u[3 * b + 1] = 0.0; s_u[3 * b + 2] = fz;
    }
    if (tid < n4 - n) { s_u[n + tid] = 0.0; s_rhs[n + tid] = 0.0; s_du[n + tid] = 0.0; s_f0[n + tid] = 0.0; s_tv[n + tid] = 0.0; }
    __syncthreads();
    symv_bc4<NT>(Hm, n4, nblk, s_u, s_rd);
    double gmax = 0.0, r0max = 0.0;
    for (int t = tid; t < n; t += NT) { gmax = fmax(gmax, fabs(s_g[t])); r0max = fmax(r0max, fabs(s_rd[t] + s_g[t])); }
    gmax = block_max<NT>(gmax, s_red);
    r0max = block_max<NT>(r0max, s_red);
    const double gs = 1.0 + gmax;
    const double mu0 = fmax(1e-2, r0max);
    for (int b = tid; b < nb; b += NT) {
      double y[5];
      cmul5(s_mu[b], s_u + 3 * b, y);
      for (int q = 0; q < 5; ++q) {
        const double ub = q < 4 ? s_ubxy[b] : s_ubz[b];
        s_sl[5 * b + q] = y[q]; s_su[5 * b + q] = ub - y[q];
        s_zl[5 * b + q] = mu0 / y[q]; s_zu[5 * b + q] = mu0 / (ub - y[q]);
      }
    }
    __syncthreads();

    int status = CMPC_STATUS_MAX_ITER, it = 0, npolish = 0;
    bool numerical = false, ipm_ok = false;
    double us = 1.0;
    for (it = 0; it <= cfg.max_iter; ++it) {
      // ---- residuals
      symv_bc4<NT>(Hm, n4, nblk, s_u, s_rd);
      double rmax = 0.0, umax = 0.0, gap = 0.0;
      for (int b = tid; b < nb; b += NT) {
        double w[5], o[3];
        for (int q = 0; q < 5; ++q) {
          w[q] = s_zl[5 * b + q] - s_zu[5 * b + q];
          gap += s_sl[5 * b + q] * s_zl[5 * b + q] + s_su[5 * b + q] * s_zu[5 * b + q];
        }
        ctmul5(s_mu[b], w, o);
        for (int q = 0; q < 3; ++q) {
          const double rr = s_rd[3 * b + q] + s_g[3 * b + q] - o[q];
          s_rd[3 * b + q] = rr;
          rmax = fmax(rmax, fabs(rr)); umax = fmax(umax, fabs(s_u[3 * b + q]));
        }
      }
      block_max2_sum<NT>(rmax, umax, gap, s_red);
      const double mu = gap / (2.0 * (double)m);
      us = 1.0 + umax;
      // Convergence. The dual residual has a round-off floor ~ eps * cond(H + C'SC) once the
      // gap is small, so the polish (which verifies the KKT conditions itself) is attempted
      // as soon as the gap is converged and the residual is merely small.
      const bool conv_mu = mu <= cfg.tol * gs * us;
      const bool strict = conv_mu && rmax <= cfg.tol * gs;
      const bool ready = conv_mu && rmax <= 1e4 * cfg.tol * gs;
      ipm_ok = conv_mu && rmax <= 10.0 * cfg.tol * gs;
      if (cfg.polish && ready && npolish < 3) {
        ++npolish;
        for (int t = tid; t < m; t += NT) {
          s_actl[t] = s_zl[t] * us > s_sl[t] * gs;
          s_actu[t] = s_zu[t] * us > s_su[t] * gs;
        }
        __syncthreads();
        // ---- active-set polish with correction passes
        bool accepted = false;
        for (int pass = 0; pass < 6 && !accepted; ++pass) {
          bool ok_all = true;
          for (int b = tid; b < nb; b += NT) {
            double A[10][3], rhsb[10], Z[3][3], f0[3];
            int k = 0;
            for (int q = 0; q < 5; ++q) if (s_actl[5 * b + q]) { row_vec(s_mu[b], q, A[k]); rhsb[k++] = 0.0; }
            for (int q = 0; q < 5; ++q) if (s_actu[5 * b + q]) { row_vec(s_mu[b], q, A[k]); rhsb[k++] = q < 4 ? s_ubxy[b] : s_ubz[b]; }
            bool okb;
            const int rk = block_nullspace(k, A, rhsb, f0, Z, &okb);
            ok_all = ok_all && okb;
            s_rk[b] = rk;
            for (int q = 0; q < 3; ++q) s_f0[3 * b + q] = f0[q];
            for (int cc = 0; cc < 3 - rk; ++cc)
              for (int q = 0; q < 3; ++q) s_Zt[9 * b + 3 * cc + q] = Z[cc][q];
          }
          ok_all = __syncthreads_and(ok_all);
          if (!ok_all) break;
          if (tid == 0) {
            int o = 0;
            for (int b = 0; b < nb; ++b) { s_off[b] = o; o += 3 - s_rk[b]; }
            s_misc[3] = o;
          }
          __syncthreads();
          const int nr = s_misc[3];
          const int nblk_r = (nr + 3) >> 2, nr4 = nblk_r << 2;
          // r = H f0 + g
          symv_bc4<NT>(Hm, n4, nblk, s_f0, s_rhs);
          for (int t = tid; t < n; t += NT) s_rhs[t] += s_g[t];
          {
            const int matr = bc4_doubles(nr4);
            for (int t = tid; t < matr; t += NT) Mm[t] = 0.0;
          }
          __syncthreads();
          // reduced system Z'HZ t = -Z'(H f0 + g)
          for (int b = tid; b < nb; b += NT) {
            for (int cc = 0; cc < 3 - s_rk[b]; ++cc) {
              const double* z = s_Zt + 9 * b + 3 * cc;
              s_tv[s_off[b] + cc] = -(z[0] * s_rhs[3 * b] + z[1] * s_rhs[3 * b + 1] + z[2] * s_rhs[3 * b + 2]);
            }
          }
          if (tid < nr4 - nr) { s_tv[nr + tid] = 0.0; Mm[midx(nr + tid, nr + tid, nblk_r)] = 1.0; }
          {
            const int npairs = (nb * (nb + 1)) >> 1;
            for (int idx = tid; idx < npairs; idx += NT) {
              int a = (int)((sqrtf(8.0f * (float)idx + 1.0f) - 1.0f) * 0.5f);
              while (((a + 1) * (a + 2)) >> 1 <= idx) ++a;
              while ((a * (a + 1)) >> 1 > idx) --a;
              const int b2 = idx - ((a * (a + 1)) >> 1), b = a;
              const int d1 = 3 - s_rk[b], d2 = 3 - s_rk[b2];
              if (d1 == 0 || d2 == 0) continue;
              double Hb[3][3];
              for (int aa = 0; aa < 3; ++aa)
                for (int bb = 0; bb < 3; ++bb) {
                  const int gi = 3 * b + aa, gj = 3 * b2 + bb;
                  Hb[aa][bb] = ((gi >> 2) >= (gj >> 2)) ? Hm[midx(gi, gj, nblk)] : Hm[midx(gj, gi, nblk)];
                }
              for (int cc = 0; cc < d1; ++cc)
                for (int c2 = 0; c2 < d2; ++c2) {
                  const double* z = s_Zt + 9 * b + 3 * cc;
                  const double* z2 = s_Zt + 9 * b2 + 3 * c2;
                  double s = 0.0;
                  for (int aa = 0; aa < 3; ++aa)
                    for (int bb = 0; bb < 3; ++bb) s += z[aa] * Hb[aa][bb] * z2[bb];
                  const int gi = s_off[b] + cc, gj = s_off[b2] + c2;
                  if (b == b2 && c2 > cc) continue;  // lower part of the diagonal block; mirrored below
                  if ((gi >> 2) >= (gj >> 2)) Mm[midx(gi, gj, nblk_r)] = s;
                  else Mm[midx(gj, gi, nblk_r)] = s;
                  if ((gi >> 2) == (gj >> 2)) Mm[midx(gj, gi, nblk_r)] = s;
                }
            }
          }
          __syncthreads();
          bool fact_ok = true;
          if (nr > 0) {
            fact_ok = chol_bc4<NT>(Mm, nblk_r);
            if (fact_ok) chol_solve_bc4<NT>(Mm, nblk_r, s_tv, s_du);
          }
          if (!fact_ok) break;
          for (int b = tid; b < nb; b += NT) {
            double f[3] = {s_f0[3 * b], s_f0[3 * b + 1], s_f0[3 * b + 2]};
            for (int cc = 0; cc < 3 - s_rk[b]; ++cc) {
              const double tv = s_tv[s_off[b] + cc];
              for (int q = 0; q < 3; ++q) f[q] += s_Zt[9 * b + 3 * cc + q] * tv;
            }
            for (int q = 0; q < 3; ++q) s_up[3 * b + q] = f[q];
          }
          if (tid < n4 - n) s_up[n + tid] = 0.0;
          __syncthreads();
          symv_bc4<NT>(Hm, n4, nblk, s_up, s_rhs);
          // multipliers, verification, correction
          bool okm = true, changed = false;
          for (int b = tid; b < nb; b += NT) {
            double Nrm[10][3], lam[10], rb[3], y[5];
            int idx[10], k = 0;
            for (int q = 0; q < 5; ++q) if (s_actl[5 * b + q]) { row_vec(s_mu[b], q, Nrm[k]); idx[k++] = q; }
            for (int q = 0; q < 5; ++q) if (s_actu[5 * b + q]) {
              row_vec(s_mu[b], q, Nrm[k]);
              Nrm[k][0] = -Nrm[k][0]; Nrm[k][1] = -Nrm[k][1]; Nrm[k][2] = -Nrm[k][2];
              idx[k++] = 5 + q;
            }
            for (int q = 0; q < 3; ++q) rb[q] = s_rhs[3 * b + q] + s_g[3 * b + q];
            okm = block_multipliers(k, Nrm, rb, 1e-9 * gs, lam) && okm;
            double ll[5] = {0, 0, 0, 0, 0}, lu[5] = {0, 0, 0, 0, 0};
            for (int s = 0; s < k; ++s) { if (idx[s] < 5) ll[idx[s]] = lam[s]; else lu[idx[s] - 5] = lam[s]; }
            cmul5(s_mu[b], s_up + 3 * b, y);
            for (int q = 0; q < 5; ++q) {
              const double ub = q < 4 ? s_ubxy[b] : s_ubz[b];
              const double sl = y[q], su = ub - y[q];
              const bool vl = sl < -1e-9 * us, vu = su < -1e-9 * us;
              const bool nl = ll[q] < -1e-9 * gs, nuu = lu[q] < -1e-9 * gs;
              if (vl || vu || nl || nuu) changed = true;
              s_actl[5 * b + q] = (s_actl[5 * b + q] || vl) && !nl;
              s_actu[5 * b + q] = (s_actu[5 * b + q] || vu) && !nuu;
              s_cdu[5 * b + q] = sl;   // candidate slacks / multipliers, committed on accept
              s_dzl[5 * b + q] = ll[q]; s_dzu[5 * b + q] = lu[q];
            }
          }
          const bool good = __syncthreads_and(okm && !changed);
          if (good) accepted = true;
        }
        if (accepted) {
          for (int t = tid; t < n; t += NT) s_u[t] = s_up[t];
          for (int b = tid; b < nb; b += NT)
            for (int q = 0; q < 5; ++q) {
              const double ub = q < 4 ? s_ubxy[b] : s_ubz[b];
              s_sl[5 * b + q] = s_cdu[5 * b + q]; s_su[5 * b + q] = ub - s_cdu[5 * b + q];
              s_zl[5 * b + q] = s_dzl[5 * b + q]; s_zu[5 * b + q] = s_dzu[5 * b + q];
            }
          __syncthreads();
          status = CMPC_STATUS_OK;
          break;
        }
        __syncthreads();
      }
      if (strict && (!cfg.polish || npolish >= 3)) break;
      if (mu <= 1e-8 * cfg.tol * gs * us) break;  // far past convergence: stop before 0/0
      if (it == cfg.max_iter) break;

      // ---- M = H + C' diag(zl/sl + zu/su) C  (only the 3x3 diagonal blocks change)
      {
        const int mat = bc4_doubles(n4);
        for (int t = tid; t < mat; t += NT) Mm[t] = Hm[t];
        __syncthreads();
        for (int b = tid; b < nb; b += NT) {
          double sg[5];
          for (int q = 0; q < 5; ++q) sg[q] = s_zl[5 * b + q] / s_sl[5 * b + q] + s_zu[5 * b + q] / s_su[5 * b + q];
          const double mb = s_mu[b], sx = sg[0] + sg[1], sy = sg[2] + sg[3];
          const int g0 = 3 * b, g1 = g0 + 1, g2 = g0 + 2;
          Mm[midx(g0, g0, nblk)] += sx;
          Mm[midx(g1, g1, nblk)] += sy;
          Mm[midx(g2, g2, nblk)] += mb * mb * (sx + sy) + sg[4];
          Mm[midx(g2, g0, nblk)] += mb * (sg[1] - sg[0]);
          Mm[midx(g2, g1, nblk)] += mb * (sg[3] - sg[2]);
        }
        __syncthreads();
      }
      if (!chol_bc4<NT>(Mm, nblk)) { numerical = true; break; }

      double alpha = 1.0, sigma = 0.0;
      for (int phase = 0; phase < 2; ++phase) {
        // phase 0: affine predictor; phase 1: centred corrector (Mehrotra)
        for (int b = tid; b < nb; b += NT) {
          double tq[5], o[3];
          for (int q = 0; q < 5; ++q) {
            const int t = 5 * b + q;
            double rcl = -s_sl[t] * s_zl[t], rcu = -s_su[t] * s_zu[t];
            if (phase) { rcl += sigma * mu - s_cdu[t] * s_dzl[t]; rcu += sigma * mu + s_cdu[t] * s_dzu[t]; }
            tq[q] = rcl / s_sl[t] - rcu / s_su[t];
          }
          ctmul5(s_mu[b], tq, o);
          for (int q = 0; q < 3; ++q) s_du[3 * b + q] = -s_rd[3 * b + q] + o[q];
        }
        __syncthreads();
        chol_solve_bc4<NT>(Mm, nblk, s_du, s_rhs);
        double amin = 1.0, ga = 0.0;
        for (int b = tid; b < nb; b += NT) {
          double y[5];
          cmul5(s_mu[b], s_du + 3 * b, y);
          for (int q = 0; q < 5; ++q) {
            const int t = 5 * b + q;
            double rcl = -s_sl[t] * s_zl[t], rcu = -s_su[t] * s_zu[t];
            if (phase) { rcl += sigma * mu - s_cdu[t] * s_dzl[t]; rcu += sigma * mu + s_cdu[t] * s_dzu[t]; }
            const double cd = y[q];
            const double dl = (rcl - s_zl[t] * cd) / s_sl[t];
            const double du_ = (rcu + s_zu[t] * cd) / s_su[t];
            if (cd < 0.0) amin = fmin(amin, -s_sl[t] / cd);
            if (cd > 0.0) amin = fmin(amin, s_su[t] / cd);
            if (dl < 0.0) amin = fmin(amin, -s_zl[t] / dl);
            if (du_ < 0.0) amin = fmin(amin, -s_zu[t] / du_);
            s_cdu[t] = cd; s_dzl[t] = dl; s_dzu[t] = du_;
          }
        }
        alpha = -block_max<NT>(-amin, s_red);
        if (!phase) {
          for (int t = tid; t < m; t += NT)
            ga += (s_sl[t] + alpha * s_cdu[t]) * (s_zl[t] + alpha * s_dzl[t]) +
                  (s_su[t] - alpha * s_cdu[t]) * (s_zu[t] + alpha * s_dzu[t]);
          ga = block_sum<NT>(ga, s_red);
          const double ratio = ga / gap;
          sigma = ratio * ratio * ratio;
        }
      }
      alpha = fmin(1.0, 0.995 * alpha);
      bool fin = true;
      for (int t = tid; t < n; t += NT) { const double v = s_u[t] + alpha * s_du[t]; s_u[t] = v; fin = fin && isfinite(v); }
      __syncthreads();
      for (int b = tid; b < nb; b += NT) {
        double y[5];
        cmul5(s_mu[b], s_u + 3 * b, y);
        for (int q = 0; q < 5; ++q) {
          const int t = 5 * b + q;
          const double ub = q < 4 ? s_ubxy[b] : s_ubz[b];
          s_zl[t] += alpha * s_dzl[t]; s_zu[t] += alpha * s_dzu[t];
          s_sl[t] = y[q]; s_su[t] = ub - y[q];
        }
      }
      fin = __syncthreads_and(fin);
      if (!fin) { numerical = true; break; }
    }
    if (numerical) status = CMPC_STATUS_NUMERICAL;
    else if (status != CMPC_STATUS_OK) status = ipm_ok ? CMPC_STATUS_OK_IPM : CMPC_STATUS_MAX_ITER;

    // ---- outputs
    if (!numerical) {
      // scaled KKT residual (same definition as the oracle)
      symv_bc4<NT>(Hm, n4, nblk, s_u, s_rhs);
      double stat = 0.0, umax = 0.0, prim = 0.0, dual = 0.0, comp = 0.0;
      for (int b = tid; b < nb; b += NT) {
        double w[5], o[3], y[5];
        for (int q = 0; q < 5; ++q) w[q] = s_zl[5 * b + q] - s_zu[5 * b + q];
        ctmul5(s_mu[b], w, o);
        cmul5(s_mu[b], s_u + 3 * b, y);
        for (int q = 0; q < 3; ++q) {
          stat = fmax(stat, fabs(s_rhs[3 * b + q] + s_g[3 * b + q] - o[q]));
          umax = fmax(umax, fabs(s_u[3 * b + q]));
        }
        for (int q = 0; q < 5; ++q) {
          const double ub = q < 4 ? s_ubxy[b] : s_ubz[b];
          const double sl = y[q], su = ub - y[q], zl = s_zl[5 * b + q], zu = s_zu[5 * b + q];
          prim = fmax(prim, fmax(-sl, -su));
          dual = fmax(dual, fmax(-zl, -zu));
          comp = fmax(comp, fmax(fabs(zl * sl), fabs(zu * su)));
        }
      }
      stat = block_max<NT>(stat, s_red);
      umax = block_max<NT>(umax, s_red);
      prim = block_max<NT>(prim, s_red);
      dual = block_max<NT>(dual, s_red);
      comp = block_max<NT>(comp, s_red);
      const double usf = 1.0 + umax;
      const double kkt = fmax(fmax(stat / gs, prim / usf), fmax(dual / gs, comp / (gs * usf)));
      if (status != CMPC_STATUS_OK) {
        for (int t = tid; t < m; t += NT) {
          s_actl[t] = s_zl[t] * usf > s_sl[t] * gs;
          s_actu[t] = s_zu[t] * usf > s_su[t] * gs;
        }
      }
      __syncthreads();
      // forces in the reference's per-leg order [L][N][3] (CentroidalMPC.cpp:270)
      for (int t = tid; t < nf; t += NT) {
        const int i = t / (3 * N), j = (t % (3 * N)) / 3, q = t % 3;
        const int b = s_blk_of[j * L + i];
        args.forces[(size_t)inst * nf + t] = b < 0 ? 0.0 : s_u[3 * b + q];
      }
      if (args.lam) {
        for (int t = tid; t < 2 * mmax; t += NT) {
          const int side = t / mmax, rem = t % mmax, ji = rem / 5, q = rem % 5;
          const int b = s_blk_of[ji];
          args.lam[(size_t)inst * 2 * mmax + t] = b < 0 ? 0.0 : (side ? s_zu[5 * b + q] : s_zl[5 * b + q]);
        }
      }
      if (args.active) {
        for (int t = tid; t < nbmax; t += NT) {
          const int b = s_blk_of[t];
          uint16_t a = 0x8000;
          if (b >= 0) {
            a = 0;
            for (int q = 0; q < 5; ++q) a |= (uint16_t)((s_actl[5 * b + q] ? 1 : 0) << q | (s_actu[5 * b + q] ? 1 : 0) << (5 + q));
          }
          args.active[(size_t)inst * nbmax + t] = a;
        }
      }
      if (tid == 0) {
        args.status[inst] = status;
        if (args.iters) args.iters[inst] = it;
        if (args.kkt) args.kkt[inst] = kkt;
      }
    } else {
      for (int t = tid; t < nf; t += NT) args.forces[(size_t)inst * nf + t] = 0.0;
      if (args.lam) for (int t = tid; t < 2 * mmax; t += NT) args.lam[(size_t)inst * 2 * mmax + t] = 0.0;
      if (args.active) for (int t = tid; t < nbmax; t += NT) args.active[(size_t)inst * nbmax + t] = 0;
      if (tid == 0) {
        args.status[inst] = status;
        if (args.iters) args.iters[inst] = it;
        if (args.kkt) args.kkt[inst] = 0.0;
      }
    }
  }
}

}  // namespace cmpc
