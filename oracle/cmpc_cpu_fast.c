/*
 * cmpc_cpu_fast.c -- the CPU BASELINE port of the centroidal-MPC condensed-QP path.
 *
 * TEST / BENCH INFRASTRUCTURE ONLY (bench.py's cpu_baseline and --impl reference legs, tests/).  The product
 * (libcmpc_b200.so) never links or calls it and has no CPU fallback.
 *
 * Same model, same algorithm and same conventions as oracle/cmpc_oracle.c (presolve, feasible-start Mehrotra
 * interior point, verified active-set polish; every reference citation there applies here), but written the way
 * a CPU implementation meant to be fast would be, so that "the reference CPU path timed beside the GPU" is not a
 * straw man:
 *   - the QP is built directly on the FREE (stance-leg) variables from closed forms of A_d^p B_j (no dense
 *     B_qp, no triple-loop H): O(n^2) instead of O(q p^2);
 *   - no allocation inside a solve: one workspace per worker thread, sized once;
 *   - a persistent pthread pool with dynamic chunking instead of a thread spawn per call;
 *   - built with -O3 -march=native (oracle/Makefile).
 * cmpc_oracle.c stays the slow, explicit CHECKER; this file is checked against it (tests/test_oracle.py::
 * test_fast_cpu_port_matches_oracle: forces <= 1e-9 relative, status / iterations / active set identical).
 */
#include "cmpc_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define GRAV 9.81       /* CentroidalMPC.cpp:71 */
#define FRIC_UB 5000.0  /* CentroidalMPC.cpp:183 */
#define MAXNB (CMPC_MAX_LEGS * CMPC_MAX_HORIZON)
#define MAXN (3 * MAXNB)

typedef struct {
  int n, nb, N, L;
  int blk_j[MAXNB], blk_i[MAXNB];
  double mu[MAXNB], ce[MAXNB], fz[MAXNB], arm[3 * MAXNB], ub[5 * MAXNB];
  double z1[CMPC_MAX_HORIZON], z2[CMPC_MAX_HORIZON], eq[9 * CMPC_MAX_HORIZON];
  double *H, *M;                 /* n x n each (leading dimension n) */
  double g[MAXN], u[MAXN], rd[MAXN], rhs[MAXN], du[MAXN], f0[MAXN], up[MAXN], r[MAXN], t[MAXN], Zt[9 * MAXNB];
  double sl[5 * MAXNB], su[5 * MAXNB], zl[5 * MAXNB], zu[5 * MAXNB], cdu[5 * MAXNB], dzl[5 * MAXNB], dzu[5 * MAXNB];
  double rcl[5 * MAXNB], rcu[5 * MAXNB], tmp[5 * MAXNB], y[5 * MAXNB], ll[5 * MAXNB], lu[5 * MAXNB];
  unsigned char actl[5 * MAXNB], actu[5 * MAXNB];
  int rk[MAXNB], off[MAXNB];
} fast_ws;

static fast_ws* ws_new(void) {
  fast_ws* w = calloc(1, sizeof(fast_ws));
  w->H = malloc((size_t)MAXN * MAXN * 8);
  w->M = malloc((size_t)MAXN * MAXN * 8);
  return w;
}
static void ws_free(fast_ws* w) { if (w) { free(w->H); free(w->M); free(w); } }

static int all_finite(const double* v, int n) {
  for (int i = 0; i < n; ++i) if (!isfinite(v[i])) return 0;
  return 1;
}
static double maxabs(const double* v, int n) {
  double m = 0;
  for (int i = 0; i < n; ++i) { double a = fabs(v[i]); if (a > m) m = a; }
  return m;
}

/* Compact build (SURVEY §8 a2-a7 on the free variables): H = 2 (Bqp' L Bqp + K), g, from the closed forms
 * A_d^p B_j = [dt^2 (p + zeta)(c/m) I; dt (c/m) I; dt c [r]x]  (zeta = 0 Euler, 1/2 ZOH; CentroidalMPC.cpp:85-92).
 * Returns 1 if a step has no stance leg (:328-330). */
static int build_compact(const cmpc_config* c, const double* state, const double* des_state, const double* des_inputs, fast_ws* w) {
  const int N = c->horizon, L = c->num_legs;
  const double dt = c->dt, mass = c->mass, zeta = c->disc_mode ? 0.5 : 0.0;
  const double* wt = c->weights;
  const double *dpos = des_state, *dvel = des_state + 3 * (N + 1), *dam = des_state + 6 * (N + 1);
  int nb = 0, invalid = 0;
  for (int j = 0; j < N; ++j) {
    double colsum = 0;
    for (int i = 0; i < L; ++i) colsum += des_inputs[i * (4 * N + 3) + j];
    if (!(colsum > 0)) invalid = 1;
    for (int i = 0; i < L; ++i) {
      const double ce = des_inputs[i * (4 * N + 3) + j];
      if (!(ce > 0)) continue;
      w->blk_j[nb] = j; w->blk_i[nb] = i; w->mu[nb] = c->mu[i]; w->ce[nb] = ce;
      w->fz[nb] = colsum > 0 ? mass * GRAV / colsum : 0.0;                       /* :331-333 */
      for (int q = 0; q < 3; ++q) w->arm[3 * nb + q] = des_inputs[i * (4 * N + 3) + N + 3 * j + q] - dpos[3 * j + q];
      for (int r = 0; r < 4; ++r) w->ub[5 * nb + r] = FRIC_UB * ce;             /* :183,199 */
      w->ub[5 * nb + 4] = mass * GRAV * L * ce;
      ++nb;
    }
  }
  w->nb = nb; w->n = 3 * nb; w->N = N; w->L = L;
  if (invalid) return 1;
  const int n = w->n;
  /* zero-input roll-out errors e_k = Q_k (x_k - xref_k), z-weighted power-stacking sums (:203-210) */
  for (int k = 0; k < N; ++k) {
    const int node = k + 1;
    const double kk = node, gpos = zeta > 0 ? 0.5 * kk * kk : 0.5 * kk * (kk - 1.0);
    const double om = (wt[2] / 2) * exp(-kk) + wt[2] / 2, qz = om * om;
    double cpos[3], v[3];
    for (int a = 0; a < 3; ++a) { cpos[a] = state[a] + kk * dt * state[3 + a]; v[a] = state[3 + a]; }
    cpos[2] += gpos * dt * dt * (-GRAV); v[2] += kk * dt * (-GRAV);
    w->eq[9 * k + 0] = wt[0] * (cpos[0] - dpos[3 * node]);
    w->eq[9 * k + 1] = wt[1] * (cpos[1] - dpos[3 * node + 1]);
    w->eq[9 * k + 2] = qz * (cpos[2] - dpos[3 * node + 2]);
    for (int a = 0; a < 3; ++a) {
      w->eq[9 * k + 3 + a] = wt[3 + a] * (v[a] - dvel[3 * node + a]);
      w->eq[9 * k + 6 + a] = wt[6 + a] * (state[6 + a] - dam[3 * node + a]);
    }
  }
  for (int j = 0; j < N; ++j) {
    double a1 = 0, a2 = 0;
    for (int k = j; k < N; ++k) {
      const double om = (wt[2] / 2) * exp(-(double)(k + 1)) + wt[2] / 2, al = (double)(k - j) + zeta;
      a1 += al * om * om; a2 += al * al * om * om;
    }
    w->z1[j] = a1; w->z2[j] = a2;
  }
  const double dt2 = dt * dt, dt4 = dt2 * dt2, im2 = 1.0 / (mass * mass);
  const double q0 = wt[6], q1 = wt[7], q2 = wt[8];
  for (int b = 0; b < nb; ++b) {
    const int j = w->blk_j[b], i = w->blk_i[b];
    const double* r = w->arm + 3 * b;
    const double cnt = N - j;
    const double s1 = 0.5 * cnt * (cnt - 1.0) + zeta * cnt;
    const double s2 = (cnt - 1.0) * cnt * (2.0 * cnt - 1.0) / 6.0 + zeta * cnt * (cnt - 1.0) + zeta * zeta * cnt;
    for (int b2 = 0; b2 <= b; ++b2) {
      const int j2 = w->blk_j[b2], i2 = w->blk_i[b2];
      const double* p = w->arm + 3 * b2;
      const double dd = j - j2, s0 = s2 + dd * s1, sz = w->z2[j] + dd * w->z1[j];
      const double cc = w->ce[b] * w->ce[b2], sc = cnt * dt2 * cc, cmm = cc * im2;
      double blk[3][3];
      blk[0][0] = sc * (r[2] * q1 * p[2] + r[1] * q2 * p[1]) + cmm * (dt4 * wt[0] * s0 + cnt * dt2 * wt[3]);
      blk[0][1] = sc * (-r[1] * q2 * p[0]);
      blk[0][2] = sc * (-r[2] * q1 * p[0]);
      blk[1][0] = sc * (-r[0] * q2 * p[1]);
      blk[1][1] = sc * (r[2] * q0 * p[2] + r[0] * q2 * p[0]) + cmm * (dt4 * wt[1] * s0 + cnt * dt2 * wt[4]);
      blk[1][2] = sc * (-r[2] * q0 * p[1]);
      blk[2][0] = sc * (-r[0] * q1 * p[2]);
      blk[2][1] = sc * (-r[1] * q0 * p[2]);
      blk[2][2] = sc * (r[1] * q0 * p[1] + r[0] * q1 * p[0]) + cmm * (dt4 * sz + cnt * dt2 * wt[5]);
      if (i == i2 && j - j2 <= 1) {                                              /* K = W_f + D' W_r D (:223-231) */
        const double nn = (j > 0 ? 1.0 : 0.0) + (j + 1 < N ? 1.0 : 0.0);
        for (int a = 0; a < 3; ++a) {
          const double wr = wt[9 + 6 * L + 3 * i + a];
          blk[a][a] += (j == j2) ? wt[9 + 3 * L + 3 * i + a] + nn * wr : -wr;
        }
      }
      for (int a = 0; a < 3; ++a)
        for (int a2 = 0; a2 < 3; ++a2) {
          const double v = 2.0 * blk[a][a2];
          w->H[(size_t)(3 * b + a) * n + 3 * b2 + a2] = v;
          w->H[(size_t)(3 * b2 + a2) * n + 3 * b + a] = v;
        }
    }
    /* g = 2 Bqp' L (Aqp x0 + dqp - Xref) - 2 W_f Uref by an adjoint sum */
    double sp[3] = {0, 0, 0}, sv[3] = {0, 0, 0}, sl3[3] = {0, 0, 0};
    for (int k = j; k < N; ++k) {
      const double al = (double)(k - j) + zeta;
      for (int q = 0; q < 3; ++q) { sp[q] += al * w->eq[9 * k + q]; sv[q] += w->eq[9 * k + 3 + q]; sl3[q] += w->eq[9 * k + 6 + q]; }
    }
    const double ce = w->ce[b], cm = ce / mass;
    const double cr[3] = {sl3[1] * r[2] - sl3[2] * r[1], sl3[2] * r[0] - sl3[0] * r[2], sl3[0] * r[1] - sl3[1] * r[0]};
    for (int q = 0; q < 3; ++q) {
      double gq = 2.0 * (cm * (dt2 * sp[q] + dt * sv[q]) + dt * ce * cr[q]);
      if (q == 2) gq -= 2.0 * wt[9 + 3 * L + 3 * i + 2] * w->fz[b];
      w->g[3 * b + q] = gq;
    }
  }
  return 0;
}

/* ---------------------------------------------------------------- dense helpers (row-major, leading dimension n) */
static int cholesky(double* A, int n) {
  for (int j = 0; j < n; ++j) {
    double* rj = A + (size_t)j * n;
    double d = rj[j];
    for (int k = 0; k < j; ++k) d -= rj[k] * rj[k];
    if (!(d > 0) || !isfinite(d)) return -1;
    d = sqrt(d);
    rj[j] = d;
    const double id = 1.0 / d;
    for (int i = j + 1; i < n; ++i) {
      double* ri = A + (size_t)i * n;
      double s = ri[j];
      for (int k = 0; k < j; ++k) s -= ri[k] * rj[k];
      ri[j] = s * id;
    }
  }
  return 0;
}
static void chol_solve(const double* Lc, int n, double* x) {
  for (int i = 0; i < n; ++i) {
    const double* ri = Lc + (size_t)i * n;
    double s = x[i];
    for (int k = 0; k < i; ++k) s -= ri[k] * x[k];
    x[i] = s / ri[i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = x[i];
    for (int k = i + 1; k < n; ++k) s -= Lc[(size_t)k * n + i] * x[k];
    x[i] = s / Lc[(size_t)i * n + i];
  }
}
static void symv(const double* H, int n, const double* x, double* y) {
  for (int i = 0; i < n; ++i) {
    const double* ri = H + (size_t)i * n;
    double s = 0;
    for (int k = 0; k < n; ++k) s += ri[k] * x[k];
    y[i] = s;
  }
}

/* friction pyramid rows (CentroidalMPC.cpp:186-190) */
static void row_vec(double mu, int r, double a[3]) {
  a[0] = a[1] = 0; a[2] = mu;
  if (r == 0) a[0] = -1; else if (r == 1) a[0] = 1; else if (r == 2) a[1] = -1;
  else if (r == 3) a[1] = 1; else a[2] = 1;
}
static void Cmul(const double* mu, int nb, const double* u, double* y) {
  for (int b = 0; b < nb; ++b) {
    const double* f = u + 3 * b; double* o = y + 5 * b; const double mf = mu[b] * f[2];
    o[0] = mf - f[0]; o[1] = mf + f[0]; o[2] = mf - f[1]; o[3] = mf + f[1]; o[4] = f[2];
  }
}
static void CTmul_add(const double* mu, int nb, const double* w, double sgn, double* out) {
  for (int b = 0; b < nb; ++b) {
    const double* v = w + 5 * b; double* o = out + 3 * b;
    o[0] += sgn * (v[1] - v[0]); o[1] += sgn * (v[3] - v[2]);
    o[2] += sgn * (mu[b] * (v[0] + v[1] + v[2] + v[3]) + v[4]);
  }
}

static int block_nullspace(int k, double A[][3], const double* b, double f0[3], double Z[3][3], int* ok) {
  double Q[3][3]; int r = 0; *ok = 1;
  f0[0] = f0[1] = f0[2] = 0;
  for (int t = 0; t < k; ++t) {
    double v[3] = {A[t][0], A[t][1], A[t][2]};
    const double na = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    for (int s = 0; s < r; ++s) {
      const double d = Q[s][0] * A[t][0] + Q[s][1] * A[t][1] + Q[s][2] * A[t][2];
      for (int a = 0; a < 3; ++a) v[a] -= d * Q[s][a];
    }
    const double nv = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    const double af0 = A[t][0] * f0[0] + A[t][1] * f0[1] + A[t][2] * f0[2];
    if (r < 3 && nv > 1e-10 * na) {
      for (int a = 0; a < 3; ++a) Q[r][a] = v[a] / nv;
      const double aq = A[t][0] * Q[r][0] + A[t][1] * Q[r][1] + A[t][2] * Q[r][2];
      const double st = (b[t] - af0) / aq;
      for (int a = 0; a < 3; ++a) f0[a] += st * Q[r][a];
      ++r;
    } else if (fabs(af0 - b[t]) > 1e-9 * (1 + fabs(b[t]))) {
      *ok = 0;
    }
  }
  if (r == 0) {
    for (int a = 0; a < 3; ++a) for (int c = 0; c < 3; ++c) Z[a][c] = (a == c);
  } else if (r == 1) {
    int m = 0;
    if (fabs(Q[0][1]) < fabs(Q[0][m])) m = 1;
    if (fabs(Q[0][2]) < fabs(Q[0][m])) m = 2;
    double e[3] = {0, 0, 0}; e[m] = 1;
    double z1[3] = {Q[0][1] * e[2] - Q[0][2] * e[1], Q[0][2] * e[0] - Q[0][0] * e[2], Q[0][0] * e[1] - Q[0][1] * e[0]};
    const double n1 = sqrt(z1[0] * z1[0] + z1[1] * z1[1] + z1[2] * z1[2]);
    for (int a = 0; a < 3; ++a) z1[a] /= n1;
    double z2[3] = {Q[0][1] * z1[2] - Q[0][2] * z1[1], Q[0][2] * z1[0] - Q[0][0] * z1[2], Q[0][0] * z1[1] - Q[0][1] * z1[0]};
    for (int a = 0; a < 3; ++a) { Z[0][a] = z1[a]; Z[1][a] = z2[a]; }
  } else if (r == 2) {
    double z[3] = {Q[0][1] * Q[1][2] - Q[0][2] * Q[1][1], Q[0][2] * Q[1][0] - Q[0][0] * Q[1][2], Q[0][0] * Q[1][1] - Q[0][1] * Q[1][0]};
    const double n1 = sqrt(z[0] * z[0] + z[1] * z[1] + z[2] * z[2]);
    for (int a = 0; a < 3; ++a) Z[0][a] = z[a] / n1;
  }
  return r;
}
static double small_lsq(int k, double S[][3], const double rb[3], double* lam) {
  double G[3][3], y[3];
  for (int a = 0; a < k; ++a) {
    y[a] = S[a][0] * rb[0] + S[a][1] * rb[1] + S[a][2] * rb[2];
    for (int b = 0; b < k; ++b) G[a][b] = S[a][0] * S[b][0] + S[a][1] * S[b][1] + S[a][2] * S[b][2];
  }
  int piv[3] = {0, 1, 2};
  for (int c = 0; c < k; ++c) {
    int m = c;
    for (int r = c + 1; r < k; ++r) if (fabs(G[piv[r]][c]) > fabs(G[piv[m]][c])) m = r;
    int t = piv[c]; piv[c] = piv[m]; piv[m] = t;
    const double d = G[piv[c]][c];
    if (fabs(d) < 1e-300) return INFINITY;
    for (int r = c + 1; r < k; ++r) {
      const double f = G[piv[r]][c] / d;
      for (int cc = c; cc < k; ++cc) G[piv[r]][cc] -= f * G[piv[c]][cc];
      y[piv[r]] -= f * y[piv[c]];
    }
  }
  for (int c = k - 1; c >= 0; --c) {
    double s = y[piv[c]];
    for (int cc = c + 1; cc < k; ++cc) s -= G[piv[c]][cc] * lam[cc];
    lam[c] = s / G[piv[c]][c];
  }
  double res = 0;
  for (int a = 0; a < 3; ++a) {
    double s = -rb[a];
    for (int t = 0; t < k; ++t) s += S[t][a] * lam[t];
    if (fabs(s) > res) res = fabs(s);
  }
  return res;
}
static int block_multipliers(int k, double Nrm[][3], const double rb[3], double tol, double* lam) {
  for (int t = 0; t < k; ++t) lam[t] = 0;
  if (k == 0) return maxabs(rb, 3) <= tol;
  double Q[3][3]; int rank = 0;
  for (int t = 0; t < k && rank < 3; ++t) {
    double v[3] = {Nrm[t][0], Nrm[t][1], Nrm[t][2]};
    const double na = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    for (int s = 0; s < rank; ++s) {
      const double d = Q[s][0] * Nrm[t][0] + Q[s][1] * Nrm[t][1] + Q[s][2] * Nrm[t][2];
      for (int a = 0; a < 3; ++a) v[a] -= d * Q[s][a];
    }
    const double nv = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (nv > 1e-10 * na) { for (int a = 0; a < 3; ++a) Q[rank][a] = v[a] / nv; ++rank; }
  }
  int have_first = 0;
  for (int mask = 1; mask < (1 << k); ++mask) {
    if (__builtin_popcount(mask) != rank) continue;
    double S[3][3], ls[3]; int idx[3], c = 0;
    for (int t = 0; t < k; ++t) if (mask >> t & 1) { memcpy(S[c], Nrm[t], 24); idx[c++] = t; }
    const double res = small_lsq(rank, S, rb, ls);
    if (!isfinite(res)) continue;
    int okk = res <= tol;
    for (int t = 0; t < rank; ++t) if (!(ls[t] >= -tol)) okk = 0;
    if (okk || !have_first) {
      for (int t = 0; t < k; ++t) lam[t] = 0;
      for (int t = 0; t < rank; ++t) lam[idx[t]] = ls[t];
      have_first = 1;
      if (okk) return 1;
    }
  }
  return 0;
}

/* Active-set polish (same passes as cmpc_oracle.c::polish); the empty working set skips the null-space algebra. */
static int polish(fast_ws* w, double gs, double us, double* u, double* zl, double* zu, int max_pass) {
  const int n = w->n, nb = w->nb, m = 5 * nb;
  int accepted = 0;
  for (int pass = 0; pass < max_pass && !accepted; ++pass) {
    int ok_all = 1, nr = 0, any = 0;
    for (int q = 0; q < m; ++q) any |= w->actl[q] | w->actu[q];
    if (!any) {
      for (int a = 0; a < n; ++a) { w->f0[a] = 0; w->t[a] = -w->g[a]; }
      memcpy(w->M, w->H, (size_t)n * n * 8);
      if (cholesky(w->M, n)) break;
      chol_solve(w->M, n, w->t);
      memcpy(w->up, w->t, n * 8);
    } else {
      for (int b = 0; b < nb; ++b) {
        double A[10][3], rhs[10], Z[3][3]; int k = 0, ok;
        for (int q = 0; q < 5; ++q) if (w->actl[5 * b + q]) { row_vec(w->mu[b], q, A[k]); rhs[k++] = 0; }
        for (int q = 0; q < 5; ++q) if (w->actu[5 * b + q]) { row_vec(w->mu[b], q, A[k]); rhs[k++] = w->ub[5 * b + q]; }
        w->rk[b] = block_nullspace(k, A, rhs, w->f0 + 3 * b, Z, &ok);
        ok_all &= ok;
        memcpy(w->Zt + 9 * b, Z, 72);
        w->off[b] = nr; nr += 3 - w->rk[b];
      }
      if (!ok_all) break;
      symv(w->H, n, w->f0, w->r);
      for (int a = 0; a < n; ++a) w->r[a] += w->g[a];
      for (int b = 0; b < nb; ++b)
        for (int cc = 0; cc < 3 - w->rk[b]; ++cc) {
          const double* z = w->Zt + 9 * b + 3 * cc;
          w->t[w->off[b] + cc] = -(z[0] * w->r[3 * b] + z[1] * w->r[3 * b + 1] + z[2] * w->r[3 * b + 2]);
          for (int b2 = 0; b2 <= b; ++b2)
            for (int c2 = 0; c2 < 3 - w->rk[b2]; ++c2) {
              const double* z2 = w->Zt + 9 * b2 + 3 * c2; double s = 0;
              for (int a = 0; a < 3; ++a)
                for (int a2 = 0; a2 < 3; ++a2) s += z[a] * w->H[(size_t)(3 * b + a) * n + 3 * b2 + a2] * z2[a2];
              w->M[(size_t)(w->off[b] + cc) * nr + w->off[b2] + c2] = s;
              w->M[(size_t)(w->off[b2] + c2) * nr + w->off[b] + cc] = s;
            }
        }
      memcpy(w->up, w->f0, n * 8);
      if (nr > 0) {
        if (cholesky(w->M, nr)) break;
        chol_solve(w->M, nr, w->t);
        for (int b = 0; b < nb; ++b)
          for (int cc = 0; cc < 3 - w->rk[b]; ++cc)
            for (int a = 0; a < 3; ++a) w->up[3 * b + a] += w->Zt[9 * b + 3 * cc + a] * w->t[w->off[b] + cc];
      }
    }
    symv(w->H, n, w->up, w->r);
    for (int a = 0; a < n; ++a) w->r[a] += w->g[a];
    Cmul(w->mu, nb, w->up, w->y);
    if (!(us > 0)) us = 1 + maxabs(w->up, n);
    int okm = 1, changed = 0;
    for (int b = 0; b < nb; ++b) {
      double Nrm[10][3], lam[10]; int idx[10], k = 0;
      for (int q = 0; q < 5; ++q) if (w->actl[5 * b + q]) { row_vec(w->mu[b], q, Nrm[k]); idx[k++] = q; }
      for (int q = 0; q < 5; ++q) if (w->actu[5 * b + q]) {
        row_vec(w->mu[b], q, Nrm[k]);
        for (int a = 0; a < 3; ++a) Nrm[k][a] = -Nrm[k][a];
        idx[k++] = 5 + q;
      }
      okm &= block_multipliers(k, Nrm, w->r + 3 * b, 1e-9 * gs, lam);
      for (int q = 0; q < 5; ++q) w->ll[5 * b + q] = w->lu[5 * b + q] = 0;
      for (int s = 0; s < k; ++s) { if (idx[s] < 5) w->ll[5 * b + idx[s]] = lam[s]; else w->lu[5 * b + idx[s] - 5] = lam[s]; }
    }
    for (int q = 0; q < m; ++q) {
      const double sl = w->y[q], su = w->ub[q] - w->y[q];
      const int vl = sl < -1e-9 * us, vu = su < -1e-9 * us;
      const int nl = w->ll[q] < -1e-9 * gs, nu_ = w->lu[q] < -1e-9 * gs;
      if (vl || vu || nl || nu_) changed = 1;
      w->actl[q] = (w->actl[q] | vl) & !nl;
      w->actu[q] = (w->actu[q] | vu) & !nu_;
    }
    if (!changed && okm) accepted = 1;
  }
  if (accepted) { memcpy(u, w->up, n * 8); memcpy(zl, w->ll, m * 8); memcpy(zu, w->lu, m * 8); }
  return accepted;
}

static void solve_one(const cmpc_config* c, const double* state, const double* des_state, const double* des_inputs,
                      double* forces, int32_t* status_out, int32_t* iters_out, double* kkt_out, uint16_t* active_out, fast_ws* w) {
  const int N = c->horizon, L = c->num_legs;
  if (N < 1 || L < 1 || N > CMPC_MAX_HORIZON || L > CMPC_MAX_LEGS) { *status_out = CMPC_STATUS_NUMERICAL; return; }
  const size_t p = (size_t)3 * L * N;
  memset(forces, 0, p * 8);
  if (active_out) for (int t = 0; t < L * N; ++t) active_out[t] = 0;
  if (iters_out) *iters_out = 0;
  if (kkt_out) *kkt_out = 0;
  if (!all_finite(state, 9 + 3 * L) || !all_finite(des_state, 9 * (N + 1)) || !all_finite(des_inputs, L * (4 * N + 3))) {
    *status_out = CMPC_STATUS_NUMERICAL; return;
  }
  if (build_compact(c, state, des_state, des_inputs, w)) { *status_out = CMPC_STATUS_INVALID_TABLE; return; }
  const int n = w->n, nb = w->nb, m = 5 * nb;
  double *u = w->u, *rd = w->rd, *rhs = w->rhs, *du = w->du, *sl = w->sl, *su = w->su, *zl = w->zl, *zu = w->zu;
  int status = CMPC_STATUS_MAX_ITER, it = 0;
  memset(u, 0, (size_t)n * 8);
  for (int b = 0; b < nb; ++b) {
    double fz = w->fz[b];
    if (fz > 0.5 * w->ub[5 * b + 4]) fz = 0.5 * w->ub[5 * b + 4];
    if (fz > 0.5 * w->ub[5 * b] / w->mu[b]) fz = 0.5 * w->ub[5 * b] / w->mu[b];
    u[3 * b + 2] = fz;
  }
  Cmul(w->mu, nb, u, sl);
  for (int t = 0; t < m; ++t) su[t] = w->ub[t] - sl[t];
  const double gs = 1 + maxabs(w->g, n);
  symv(w->H, n, u, rd);
  for (int a = 0; a < n; ++a) rd[a] += w->g[a];
  double mu0 = maxabs(rd, n); if (mu0 < 1e-2) mu0 = 1e-2;
  for (int t = 0; t < m; ++t) { zl[t] = mu0 / sl[t]; zu[t] = mu0 / su[t]; }
  memset(w->actl, 0, (size_t)m); memset(w->actu, 0, (size_t)m);
  int npolish = 0, numerical = 0, ipm_ok = 0;
  if (c->polish && c->presolve && polish(w, gs, 0.0, u, zl, zu, 1)) status = CMPC_STATUS_OK;
  for (it = 0; status != CMPC_STATUS_OK && it <= c->max_iter; ++it) {
    symv(w->H, n, u, rd);
    for (int a = 0; a < n; ++a) rd[a] += w->g[a];
    CTmul_add(w->mu, nb, zl, -1.0, rd);
    CTmul_add(w->mu, nb, zu, +1.0, rd);
    double gap = 0;
    for (int t = 0; t < m; ++t) gap += sl[t] * zl[t] + su[t] * zu[t];
    const double mu = gap / (2.0 * m), us = 1 + maxabs(u, n), rmax = maxabs(rd, n);
    const int conv_mu = mu <= c->ipm_tol * gs * us;
    const int strict = conv_mu && rmax <= c->ipm_tol * gs;
    const int ready = conv_mu && rmax <= 1e4 * c->ipm_tol * gs;
    ipm_ok = conv_mu && rmax <= 10 * c->ipm_tol * gs;
    if (c->polish && ready && npolish < 3) {
      ++npolish;
      for (int t = 0; t < m; ++t) { w->actl[t] = zl[t] * us > sl[t] * gs; w->actu[t] = zu[t] * us > su[t] * gs; }
      if (polish(w, gs, us, u, zl, zu, 6)) { status = CMPC_STATUS_OK; break; }
    }
    if (strict && (!c->polish || npolish >= 3)) break;
    if (mu <= 1e-8 * c->ipm_tol * gs * us) break;
    if (it == c->max_iter) break;
    memcpy(w->M, w->H, (size_t)n * n * 8);
    for (int b = 0; b < nb; ++b) {
      double sg[5];
      for (int q = 0; q < 5; ++q) sg[q] = zl[5 * b + q] / sl[5 * b + q] + zu[5 * b + q] / su[5 * b + q];
      const double mb = w->mu[b], sx = sg[0] + sg[1], sy = sg[2] + sg[3];
      double* d = w->M + (size_t)(3 * b) * n + 3 * b;
      d[0] += sx; d[n + 1] += sy; d[2 * n + 2] += mb * mb * (sx + sy) + sg[4];
      d[2] += mb * (sg[1] - sg[0]); d[2 * n] += mb * (sg[1] - sg[0]);
      d[n + 2] += mb * (sg[3] - sg[2]); d[2 * n + 1] += mb * (sg[3] - sg[2]);
    }
    if (cholesky(w->M, n)) { numerical = 1; break; }
    double alpha = 1, sigma = 0;
    for (int phase = 0; phase < 2; ++phase) {
      for (int t = 0; t < m; ++t) {
        w->rcl[t] = -sl[t] * zl[t]; w->rcu[t] = -su[t] * zu[t];
        if (phase) { w->rcl[t] += sigma * mu - w->cdu[t] * w->dzl[t]; w->rcu[t] += sigma * mu + w->cdu[t] * w->dzu[t]; }
        w->tmp[t] = w->rcl[t] / sl[t] - w->rcu[t] / su[t];
      }
      for (int a = 0; a < n; ++a) rhs[a] = -rd[a];
      CTmul_add(w->mu, nb, w->tmp, +1.0, rhs);
      memcpy(du, rhs, n * 8);
      chol_solve(w->M, n, du);
      Cmul(w->mu, nb, du, w->cdu);
      alpha = 1;
      for (int t = 0; t < m; ++t) {
        w->dzl[t] = (w->rcl[t] - zl[t] * w->cdu[t]) / sl[t];
        w->dzu[t] = (w->rcu[t] + zu[t] * w->cdu[t]) / su[t];
        if (w->cdu[t] < 0 && -sl[t] / w->cdu[t] < alpha) alpha = -sl[t] / w->cdu[t];
        if (w->cdu[t] > 0 && su[t] / w->cdu[t] < alpha) alpha = su[t] / w->cdu[t];
        if (w->dzl[t] < 0 && -zl[t] / w->dzl[t] < alpha) alpha = -zl[t] / w->dzl[t];
        if (w->dzu[t] < 0 && -zu[t] / w->dzu[t] < alpha) alpha = -zu[t] / w->dzu[t];
      }
      if (!phase) {
        double ga = 0;
        for (int t = 0; t < m; ++t)
          ga += (sl[t] + alpha * w->cdu[t]) * (zl[t] + alpha * w->dzl[t]) + (su[t] - alpha * w->cdu[t]) * (zu[t] + alpha * w->dzu[t]);
        const double ratio = ga / gap;
        sigma = ratio * ratio * ratio;
      }
    }
    { double tau = 1.0 - mu / (gs * us); if (tau < 0.995) tau = 0.995; alpha *= tau; }
    if (alpha > 1) alpha = 1;
    for (int a = 0; a < n; ++a) u[a] += alpha * du[a];
    for (int t = 0; t < m; ++t) { zl[t] += alpha * w->dzl[t]; zu[t] += alpha * w->dzu[t]; }
    Cmul(w->mu, nb, u, sl);
    for (int t = 0; t < m; ++t) su[t] = w->ub[t] - sl[t];
    if (!all_finite(u, n)) { numerical = 1; break; }
  }
  if (numerical) status = CMPC_STATUS_NUMERICAL;
  else if (status != CMPC_STATUS_OK) status = ipm_ok ? CMPC_STATUS_OK_IPM : CMPC_STATUS_MAX_ITER;
  if (!numerical) {
    /* scaled KKT residual, reported active set: as cmpc_oracle.c */
    symv(w->H, n, u, w->r);
    for (int a = 0; a < n; ++a) w->r[a] += w->g[a];
    CTmul_add(w->mu, nb, zl, -1.0, w->r);
    CTmul_add(w->mu, nb, zu, +1.0, w->r);
    Cmul(w->mu, nb, u, sl);
    const double us = 1 + maxabs(u, n);
    double stat = maxabs(w->r, n) / gs, prim = 0, dual = 0, comp = 0;
    for (int t = 0; t < m; ++t) {
      su[t] = w->ub[t] - sl[t];
      if (-sl[t] > prim) prim = -sl[t];
      if (-su[t] > prim) prim = -su[t];
      if (-zl[t] > dual) dual = -zl[t];
      if (-zu[t] > dual) dual = -zu[t];
      if (fabs(zl[t] * sl[t]) > comp) comp = fabs(zl[t] * sl[t]);
      if (fabs(zu[t] * su[t]) > comp) comp = fabs(zu[t] * su[t]);
    }
    prim /= us; dual /= gs; comp /= gs * us;
    double k = stat; if (prim > k) k = prim; if (dual > k) k = dual; if (comp > k) k = comp;
    if (kkt_out) *kkt_out = k;
    for (int b = 0; b < nb; ++b) {
      const int j = w->blk_j[b], i = w->blk_i[b];
      for (int r = 0; r < 3; ++r) forces[(size_t)i * 3 * N + 3 * j + r] = u[3 * b + r];
      if (active_out) {
        uint16_t a = 0;
        for (int q = 0; q < 5; ++q) {
          const int t = 5 * b + q;
          int al, au;
          if (status == CMPC_STATUS_OK) { al = sl[t] <= 1e-9 * us; au = su[t] <= 1e-9 * us; }
          else { al = zl[t] * us > sl[t] * gs; au = zu[t] * us > su[t] * gs; }
          a |= (uint16_t)((al ? 1 : 0) << q | (au ? 1 : 0) << (5 + q));
        }
        active_out[j * L + i] = a;
      }
    }
    if (active_out)
      for (int j = 0; j < N; ++j)
        for (int i = 0; i < L; ++i) if (!(des_inputs[i * (4 * N + 3) + j] > 0)) active_out[j * L + i] = 0x8000;
  }
  *status_out = status;
  if (iters_out) *iters_out = it;
}

/* ---------------------------------------------------------------- persistent thread pool */
typedef struct {
  const cmpc_config* c; int B; const double *st, *ds, *di;
  double* forces; int32_t *status, *iters; double* kkt; uint16_t* active;
} fast_job;

static struct {
  pthread_mutex_t mu; pthread_cond_t go, done;
  pthread_t th[256]; fast_ws* ws[256];
  int nthreads, generation, running, next, shutdown, spawn_gen;
  fast_job job;
} P = {PTHREAD_MUTEX_INITIALIZER, PTHREAD_COND_INITIALIZER, PTHREAD_COND_INITIALIZER, {0}, {0}, 0, 0, 0, 0, 0, 0, {0}};

static void run_chunks(fast_ws* w) {
  const fast_job* J = &P.job;
  const cmpc_config* c = J->c;
  const int N = c->horizon, L = c->num_legs;
  const size_t ns = 9 + 3 * L, nd = 9 * (N + 1), ni = (size_t)L * (4 * N + 3), nf = (size_t)3 * L * N;
  for (;;) {
    const int b0 = __atomic_fetch_add(&P.next, 16, __ATOMIC_RELAXED);
    if (b0 >= J->B) break;
    const int b1 = b0 + 16 < J->B ? b0 + 16 : J->B;
    for (int b = b0; b < b1; ++b)
      solve_one(c, J->st + b * ns, J->ds + b * nd, J->di + b * ni, J->forces + b * nf, J->status + b,
                J->iters ? J->iters + b : NULL, J->kkt ? J->kkt + b : NULL, J->active ? J->active + (size_t)b * L * N : NULL, w);
  }
}
static void* pool_worker(void* arg) {
  const int id = (int)(size_t)arg;
  pthread_mutex_lock(&P.mu);
  int seen = P.spawn_gen;  /* the generation at creation time: the job published right after it is this worker's first */
  for (;;) {
    while (P.generation == seen && !P.shutdown) pthread_cond_wait(&P.go, &P.mu);
    if (P.shutdown) break;
    seen = P.generation;
    pthread_mutex_unlock(&P.mu);
    run_chunks(P.ws[id]);
    pthread_mutex_lock(&P.mu);
    if (--P.running == 0) pthread_cond_signal(&P.done);
  }
  pthread_mutex_unlock(&P.mu);
  return NULL;
}

/* Worker threads are created on the first call (or when nthreads changes) and then re-used. */
int cmpc_fast_solve_batch(const cmpc_config* c, int B, const double* state, const double* des_state, const double* des_inputs,
                          double* forces, int32_t* status, int32_t* iters, double* kkt, uint16_t* active, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  if (c->horizon > CMPC_MAX_HORIZON || c->num_legs > CMPC_MAX_LEGS) return -1;
  pthread_mutex_lock(&P.mu);
  if (P.nthreads != nthreads) {
    if (P.nthreads > 1) {  /* retire the old pool */
      P.shutdown = 1;
      pthread_cond_broadcast(&P.go);
      pthread_mutex_unlock(&P.mu);
      for (int t = 1; t < P.nthreads; ++t) pthread_join(P.th[t], NULL);
      pthread_mutex_lock(&P.mu);
      P.shutdown = 0;
    }
    for (int t = 0; t < 256; ++t) if (t >= nthreads && P.ws[t]) { ws_free(P.ws[t]); P.ws[t] = NULL; }
    for (int t = 0; t < nthreads; ++t) if (!P.ws[t]) P.ws[t] = ws_new();
    P.nthreads = nthreads;
    P.spawn_gen = P.generation;
    for (int t = 1; t < nthreads; ++t) pthread_create(&P.th[t], NULL, pool_worker, (void*)(size_t)t);
  }
  P.job = (fast_job){c, B, state, des_state, des_inputs, forces, status, iters, kkt, active};
  P.next = 0;
  P.running = nthreads - 1;
  ++P.generation;
  pthread_cond_broadcast(&P.go);
  pthread_mutex_unlock(&P.mu);
  run_chunks(P.ws[0]);  /* the calling thread works too */
  pthread_mutex_lock(&P.mu);
  while (P.running > 0) pthread_cond_wait(&P.done, &P.mu);
  pthread_mutex_unlock(&P.mu);
  return 0;
}
