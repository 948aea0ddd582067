/*
 * cmpc_oracle.c -- CPU fp64 restatement of the centroidal-MPC condensed-QP path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this.  The product (libcmpc_b200.so)
 * never links or calls it and has no CPU fallback.
 *
 * PARITY UNPINNED: the reference's arithmetic lives in un-vendored CasADi + IPOPT + HSL
 * (CMakeLists.txt:7 links bare `casadi`, no version pin), none present here; the
 * reference returns `{}` from UpdateMPC (CentroidalMPC.cpp:369) and ships no expected
 * outputs (CentoidMPCTest.cpp:113-115).  This file restates the published model line by
 * line (citations below) with lever arms frozen at the reference trajectory, which turns
 * the reference NLP (CentroidalMPC.cpp:102-276) into the convex QP
 *     min 1/2 U'HU + g'U   s.t.  0 <= F_i f_ij <= ub * c_ij
 * It is pinned instead by oracle/numpy_mirror.py (independent dense build + independent
 * KKT evaluation + textbook active-set solver) and by tests/golden.
 *
 * Deliberately dense and explicit: A_d/B_j are formed as matrices, A_qp by repeated
 * multiplication ("power stacking"), B_qp block by block, H by a triple loop -- so that
 * the CUDA path's closed forms and recursions are checked against different arithmetic.
 */
#include "cmpc_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define GRAV 9.81       /* CentroidalMPC.cpp:71 */
#define FRIC_UB 5000.0  /* CentroidalMPC.cpp:183 */
#define NX 9

typedef struct {
  int N, L, nu, p, q;
  double x0[NX];
  double feet[3 * CMPC_MAX_LEGS];
  const double *dpos, *dvel, *dam;               /* 3x(N+1) col-major each */
  const double* contact[CMPC_MAX_LEGS];          /* N each */
  const double* dfoot[CMPC_MAX_LEGS];            /* 3x(N+1) col-major each */
  int invalid;
} orc_inputs;

static void unpack(const cmpc_config* c, const double* state, const double* des_state,
                   const double* des_inputs, orc_inputs* in) {
  /* CentroidalMPC.cpp:284-317 */
  int N = c->horizon, L = c->num_legs;
  in->N = N; in->L = L; in->nu = 3 * L; in->p = 3 * L * N; in->q = NX * N;
  memcpy(in->x0, state, NX * sizeof(double));
  memcpy(in->feet, state + NX, 3 * L * sizeof(double));
  in->dpos = des_state;
  in->dvel = des_state + 3 * (N + 1);
  in->dam = des_state + 6 * (N + 1);
  for (int i = 0; i < L; ++i) {
    in->contact[i] = des_inputs + i * (4 * N + 3);
    in->dfoot[i] = des_inputs + i * (4 * N + 3) + N;
  }
  in->invalid = 0;
  for (int j = 0; j < N; ++j) { /* CentroidalMPC.cpp:328-330 */
    double s = 0;
    for (int i = 0; i < L; ++i) s += in->contact[i][j];
    if (!(s > 0)) in->invalid = 1;
  }
}

static int all_finite(const double* v, int n) {
  for (int i = 0; i < n; ++i) if (!isfinite(v[i])) return 0;
  return 1;
}

/* A_d (9x9), B_j (9 x nu), d (9) for interval j. CentroidalMPC.cpp:85-92 with the lever
 * arm r_i = des_foot_pos_i[:,j] - des_com_pos[:,j] frozen. */
static void discretize(const cmpc_config* c, const orc_inputs* in, int j, double* Ad,
                       double* Bj, double* dj) {
  int nu = in->nu;
  double dt = c->dt, m = c->mass;
  double Ac[NX * NX] = {0}, Bc[NX * 3 * CMPC_MAX_LEGS] = {0}, dc[NX] = {0};
  for (int r = 0; r < 3; ++r) Ac[r * NX + 3 + r] = 1.0;
  dc[5] = -GRAV;
  for (int i = 0; i < in->L; ++i) {
    double ce = in->contact[i][j] > 0 ? in->contact[i][j] : 0.0;
    double r[3];
    for (int a = 0; a < 3; ++a) r[a] = in->dfoot[i][3 * j + a] - in->dpos[3 * j + a];
    for (int a = 0; a < 3; ++a) Bc[(3 + a) * nu + 3 * i + a] = ce / m;
    /* c * [r]x */
    Bc[(6 + 0) * nu + 3 * i + 1] = -ce * r[2];
    Bc[(6 + 0) * nu + 3 * i + 2] = ce * r[1];
    Bc[(6 + 1) * nu + 3 * i + 0] = ce * r[2];
    Bc[(6 + 1) * nu + 3 * i + 2] = -ce * r[0];
    Bc[(6 + 2) * nu + 3 * i + 0] = -ce * r[1];
    Bc[(6 + 2) * nu + 3 * i + 1] = ce * r[0];
  }
  /* Euler: A = I + dt Ac, B = dt Bc, d = dt dc.
   * ZOH:   A = I + dt Ac (Ac^2 = 0), B = dt Bc + dt^2/2 Ac Bc, d = dt dc + dt^2/2 Ac dc. */
  for (int a = 0; a < NX * NX; ++a) Ad[a] = dt * Ac[a];
  for (int a = 0; a < NX; ++a) Ad[a * NX + a] += 1.0;
  for (int a = 0; a < NX * nu; ++a) Bj[a] = dt * Bc[a];
  for (int a = 0; a < NX; ++a) dj[a] = dt * dc[a];
  if (c->disc_mode == 1) {
    for (int a = 0; a < NX; ++a) {
      for (int k = 0; k < NX; ++k) {
        double f = 0.5 * dt * dt * Ac[a * NX + k];
        if (f == 0.0) continue;
        for (int b = 0; b < nu; ++b) Bj[a * nu + b] += f * Bc[k * nu + b];
        dj[a] += f * dc[k];
      }
    }
  }
}

static void matmul(const double* A, const double* B, double* C, int m, int k, int n) {
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) {
      double s = 0;
      for (int t = 0; t < k; ++t) s += A[i * k + t] * B[t * n + j];
      C[i * n + j] = s;
    }
}

typedef struct {
  int N, L, nu, p, q;
  double *Aqp, *Bqp, *dqp, *Lw, *Xref, *K, *Uref, *wf, *H, *g, *xfree;
  unsigned char* pinned;
  int invalid;
} orc_qp;

static void qp_free(orc_qp* Q) {
  free(Q->Aqp); free(Q->Bqp); free(Q->dqp); free(Q->Lw); free(Q->Xref); free(Q->K);
  free(Q->Uref); free(Q->wf); free(Q->H); free(Q->g); free(Q->xfree); free(Q->pinned);
}

/* SURVEY §8 a2-a7. H, g in the full 3LN layout, pinned rows/cols masked to identity. */
static void build_qp(const cmpc_config* c, const orc_inputs* in, orc_qp* Q) {
  int N = in->N, L = in->L, nu = in->nu, p = in->p, q = in->q;
  const double* w = c->weights;
  Q->N = N; Q->L = L; Q->nu = nu; Q->p = p; Q->q = q; Q->invalid = in->invalid;
  Q->Aqp = calloc((size_t)q * NX, 8); Q->Bqp = calloc((size_t)q * p, 8);
  Q->dqp = calloc(q, 8); Q->Lw = calloc(q, 8); Q->Xref = calloc(q, 8);
  Q->K = calloc((size_t)p * p, 8); Q->Uref = calloc(p, 8); Q->wf = calloc(p, 8);
  Q->H = calloc((size_t)p * p, 8); Q->g = calloc(p, 8); Q->xfree = calloc(q, 8);
  Q->pinned = calloc(p, 1);

  double Ad[NX * NX];
  double* Bs = calloc((size_t)N * NX * nu, 8);
  double* ds = calloc((size_t)N * NX, 8);
  for (int j = 0; j < N; ++j) discretize(c, in, j, Ad, Bs + (size_t)j * NX * nu, ds + j * NX);
  /* power stacking: Apow[k] = Ad^k */
  double* Apow = calloc((size_t)(N + 1) * NX * NX, 8);
  for (int a = 0; a < NX; ++a) Apow[a * NX + a] = 1.0;
  for (int k = 1; k <= N; ++k) matmul(Ad, Apow + (size_t)(k - 1) * NX * NX, Apow + (size_t)k * NX * NX, NX, NX, NX);
  double blk[NX * 3 * CMPC_MAX_LEGS], v9[NX];
  for (int k = 0; k < N; ++k) {
    memcpy(Q->Aqp + (size_t)k * NX * NX, Apow + (size_t)(k + 1) * NX * NX, NX * NX * 8);
    for (int j = 0; j <= k; ++j) {
      matmul(Apow + (size_t)(k - j) * NX * NX, Bs + (size_t)j * NX * nu, blk, NX, NX, nu);
      for (int a = 0; a < NX; ++a)
        memcpy(Q->Bqp + (size_t)(NX * k + a) * p + nu * j, blk + a * nu, nu * 8);
      matmul(Apow + (size_t)(k - j) * NX * NX, ds + j * NX, v9, NX, NX, 1);
      for (int a = 0; a < NX; ++a) Q->dqp[NX * k + a] += v9[a];
    }
  }
  /* cost weights CentroidalMPC.cpp:203-216: omega_k multiplies the z error INSIDE the square */
  for (int k = 0; k < N; ++k) {
    int node = k + 1;
    double om = (w[2] / 2) * exp(-(double)node) + w[2] / 2;
    double* l = Q->Lw + NX * k;
    l[0] = w[0]; l[1] = w[1]; l[2] = om * om;
    for (int a = 3; a < NX; ++a) l[a] = w[a];
    for (int a = 0; a < 3; ++a) {
      Q->Xref[NX * k + a] = in->dpos[3 * node + a];
      Q->Xref[NX * k + 3 + a] = in->dvel[3 * node + a];
      Q->Xref[NX * k + 6 + a] = in->dam[3 * node + a];
    }
  }
  /* force tracking (CentroidalMPC.cpp:223-225), desired fz = m g / #stance (:331-333) */
  for (int j = 0; j < N; ++j) {
    double colsum = 0;
    for (int i = 0; i < L; ++i) colsum += in->contact[i][j];
    for (int i = 0; i < L; ++i) {
      int stance = in->contact[i][j] > 0;
      for (int r = 0; r < 3; ++r) {
        int a = nu * j + 3 * i + r;
        Q->wf[a] = w[9 + 3 * L + 3 * i + r];
        Q->pinned[a] = !stance;
        Q->K[(size_t)a * p + a] += Q->wf[a];
      }
      if (stance && !in->invalid) Q->Uref[nu * j + 3 * i + 2] = c->mass * GRAV / colsum;
    }
  }
  /* force-rate term D' W_r D (CentroidalMPC.cpp:227-231) */
  for (int j = 0; j + 1 < N; ++j)
    for (int a = 0; a < nu; ++a) {
      double wr = w[9 + 6 * L + a];
      int u0 = nu * j + a, u1 = nu * (j + 1) + a;
      Q->K[(size_t)u0 * p + u0] += wr; Q->K[(size_t)u1 * p + u1] += wr;
      Q->K[(size_t)u0 * p + u1] -= wr; Q->K[(size_t)u1 * p + u0] -= wr;
    }
  /* H = 2 (Bqp' L Bqp + K): lower-block-triangular Bqp, pinned columns are zero */
  for (int a = 0; a < p; ++a) {
    if (Q->pinned[a]) continue;
    int ja = a / nu;
    for (int b = 0; b <= a; ++b) {
      if (Q->pinned[b]) continue;
      double s = 0;
      for (int r = NX * ja; r < q; ++r) s += Q->Bqp[(size_t)r * p + a] * Q->Lw[r] * Q->Bqp[(size_t)r * p + b];
      s = 2.0 * (s + Q->K[(size_t)a * p + b]);
      Q->H[(size_t)a * p + b] = s; Q->H[(size_t)b * p + a] = s;
    }
  }
  /* g = 2 Bqp' L (Aqp x0 + dqp - Xref) - 2 W_f Uref */
  for (int r = 0; r < q; ++r) {
    double s = Q->dqp[r];
    for (int a = 0; a < NX; ++a) s += Q->Aqp[(size_t)r * NX + a] * in->x0[a];
    Q->xfree[r] = s;
  }
  for (int a = 0; a < p; ++a) {
    if (Q->pinned[a]) { Q->H[(size_t)a * p + a] = 1.0; continue; }
    double s = 0;
    for (int r = NX * (a / nu); r < q; ++r) s += Q->Bqp[(size_t)r * p + a] * Q->Lw[r] * (Q->xfree[r] - Q->Xref[r]);
    Q->g[a] = 2.0 * s - 2.0 * Q->wf[a] * Q->Uref[a];
  }
  free(Bs); free(ds); free(Apow);
}

int cmpc_oracle_build(const cmpc_config* c, const double* state, const double* des_state,
                      const double* des_inputs, double* H, double* g, int32_t* status) {
  orc_inputs in; orc_qp Q;
  unpack(c, state, des_state, des_inputs, &in);
  build_qp(c, &in, &Q);
  memcpy(H, Q.H, (size_t)Q.p * Q.p * 8);
  memcpy(g, Q.g, (size_t)Q.p * 8);
  if (status) *status = in.invalid ? CMPC_STATUS_INVALID_TABLE : CMPC_STATUS_OK;
  qp_free(&Q);
  return 0;
}

/* ---------------------------------------------------------------- dense helpers */
static int cholesky(double* A, int n) { /* in place, lower, row-major; 0 ok */
  for (int j = 0; j < n; ++j) {
    double d = A[(size_t)j * n + j];
    for (int k = 0; k < j; ++k) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
    if (!(d > 0) || !isfinite(d)) return -1;
    d = sqrt(d);
    A[(size_t)j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double s = A[(size_t)i * n + j];
      const double *ri = A + (size_t)i * n, *rj = A + (size_t)j * n;
      for (int k = 0; k < j; ++k) s -= ri[k] * rj[k];
      A[(size_t)i * n + j] = s / d;
    }
  }
  return 0;
}
static void chol_solve(const double* Lc, int n, double* x) {
  for (int i = 0; i < n; ++i) {
    double s = x[i];
    for (int k = 0; k < i; ++k) s -= Lc[(size_t)i * n + k] * x[k];
    x[i] = s / Lc[(size_t)i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = x[i];
    for (int k = i + 1; k < n; ++k) s -= Lc[(size_t)k * n + i] * x[k];
    x[i] = s / Lc[(size_t)i * n + i];
  }
}
static void symv(const double* H, int n, const double* x, double* y) {
  for (int i = 0; i < n; ++i) {
    double s = 0;
    for (int k = 0; k < n; ++k) s += H[(size_t)i * n + k] * x[k];
    y[i] = s;
  }
}
static double maxabs(const double* v, int n) {
  double m = 0;
  for (int i = 0; i < n; ++i) if (fabs(v[i]) > m) m = fabs(v[i]);
  return m;
}

/* ---------------------------------------------------------------- friction pyramid
 * rows of F_i (CentroidalMPC.cpp:186-190): y0 = mu fz - fx, y1 = mu fz + fx,
 * y2 = mu fz - fy, y3 = mu fz + fy, y4 = fz;  0 <= y <= ub c (:199). */
static void row_vec(double mu, int r, double a[3]) {
  a[0] = a[1] = 0; a[2] = mu;
  if (r == 0) a[0] = -1; else if (r == 1) a[0] = 1; else if (r == 2) a[1] = -1;
  else if (r == 3) a[1] = 1; else a[2] = 1;
}
static void Cmul(const double* mu, int nb, const double* u, double* y) {
  for (int b = 0; b < nb; ++b) {
    const double* f = u + 3 * b; double* o = y + 5 * b; double mf = mu[b] * f[2];
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

/* Null space of the active rows of one block by Gram-Schmidt. Returns rank; f0 satisfies
 * the active equalities, Z (3 x (3-rank), column-major in Z[col][3]). ok=0 if the active
 * equalities are inconsistent. */
static int block_nullspace(int k, double A[][3], const double* b, double f0[3], double Z[3][3], int* ok) {
  double Q[3][3]; int r = 0; *ok = 1;
  f0[0] = f0[1] = f0[2] = 0;
  for (int t = 0; t < k; ++t) {
    double v[3] = {A[t][0], A[t][1], A[t][2]};
    double na = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    for (int s = 0; s < r; ++s) {
      double d = Q[s][0] * A[t][0] + Q[s][1] * A[t][1] + Q[s][2] * A[t][2];
      for (int a = 0; a < 3; ++a) v[a] -= d * Q[s][a];
    }
    double nv = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    double af0 = A[t][0] * f0[0] + A[t][1] * f0[1] + A[t][2] * f0[2];
    if (r < 3 && nv > 1e-10 * na) {
      for (int a = 0; a < 3; ++a) Q[r][a] = v[a] / nv;
      double aq = A[t][0] * Q[r][0] + A[t][1] * Q[r][1] + A[t][2] * Q[r][2];
      double st = (b[t] - af0) / aq;
      for (int a = 0; a < 3; ++a) f0[a] += st * Q[r][a];
      ++r;
    } else if (fabs(af0 - b[t]) > 1e-9 * (1 + fabs(b[t]))) {
      *ok = 0;
    }
  }
  if (r == 0) {
    for (int a = 0; a < 3; ++a) for (int c = 0; c < 3; ++c) Z[a][c] = (a == c);
  } else if (r == 1) {
    int m = 0; /* axis least aligned with q */
    if (fabs(Q[0][1]) < fabs(Q[0][m])) m = 1;
    if (fabs(Q[0][2]) < fabs(Q[0][m])) m = 2;
    double e[3] = {0, 0, 0}; e[m] = 1;
    double z1[3] = {Q[0][1] * e[2] - Q[0][2] * e[1], Q[0][2] * e[0] - Q[0][0] * e[2], Q[0][0] * e[1] - Q[0][1] * e[0]};
    double n1 = sqrt(z1[0] * z1[0] + z1[1] * z1[1] + z1[2] * z1[2]);
    for (int a = 0; a < 3; ++a) z1[a] /= n1;
    double z2[3] = {Q[0][1] * z1[2] - Q[0][2] * z1[1], Q[0][2] * z1[0] - Q[0][0] * z1[2], Q[0][0] * z1[1] - Q[0][1] * z1[0]};
    for (int a = 0; a < 3; ++a) { Z[0][a] = z1[a]; Z[1][a] = z2[a]; }
  } else if (r == 2) {
    double z[3] = {Q[0][1] * Q[1][2] - Q[0][2] * Q[1][1], Q[0][2] * Q[1][0] - Q[0][0] * Q[1][2], Q[0][0] * Q[1][1] - Q[0][1] * Q[1][0]};
    double n1 = sqrt(z[0] * z[0] + z[1] * z[1] + z[2] * z[2]);
    for (int a = 0; a < 3; ++a) Z[0][a] = z[a] / n1;
  }
  return r;
}

/* Solve S' lam = rb for the k (<=3) independent signed normals S (k x 3) by the normal
 * equations; returns residual inf-norm. */
static double small_lsq(int k, double S[][3], const double rb[3], double* lam) {
  double G[3][3], y[3];
  for (int a = 0; a < k; ++a) {
    y[a] = S[a][0] * rb[0] + S[a][1] * rb[1] + S[a][2] * rb[2];
    for (int b = 0; b < k; ++b) G[a][b] = S[a][0] * S[b][0] + S[a][1] * S[b][1] + S[a][2] * S[b][2];
  }
  /* Gaussian elimination with partial pivoting on k x k */
  int piv[3] = {0, 1, 2};
  for (int c = 0; c < k; ++c) {
    int m = c;
    for (int r = c + 1; r < k; ++r) if (fabs(G[piv[r]][c]) > fabs(G[piv[m]][c])) m = r;
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
  double res = 0;
  for (int a = 0; a < 3; ++a) {
    double s = -rb[a];
    for (int t = 0; t < k; ++t) s += S[t][a] * lam[t];
    if (fabs(s) > res) res = fabs(s);
  }
  return res;
}

/* lam >= 0 with sum lam_t Nrm_t = rb over the k active signed normals; enumerates
 * independent subsets of size rank in a fixed order (degenerate vertices: the apex). */
static int block_multipliers(int k, double Nrm[][3], const double rb[3], double tol, double* lam) {
  for (int t = 0; t < k; ++t) lam[t] = 0;
  if (k == 0) return maxabs(rb, 3) <= tol;
  /* rank by Gram-Schmidt */
  double Q[3][3]; int rank = 0;
  for (int t = 0; t < k && rank < 3; ++t) {
    double v[3] = {Nrm[t][0], Nrm[t][1], Nrm[t][2]};
    double na = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    for (int s = 0; s < rank; ++s) {
      double d = Q[s][0] * Nrm[t][0] + Q[s][1] * Nrm[t][1] + Q[s][2] * Nrm[t][2];
      for (int a = 0; a < 3; ++a) v[a] -= d * Q[s][a];
    }
    double nv = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (nv > 1e-10 * na) { for (int a = 0; a < 3; ++a) Q[rank][a] = v[a] / nv; ++rank; }
  }
  int have_first = 0;
  for (int mask = 1; mask < (1 << k); ++mask) {
    if (__builtin_popcount(mask) != rank) continue;
    double S[3][3]; int idx[3], c = 0;
    for (int t = 0; t < k; ++t) if (mask >> t & 1) { memcpy(S[c], Nrm[t], 24); idx[c++] = t; }
    /* independence check: Gram determinant via elimination inside small_lsq */
    double ls[3];
    double res = small_lsq(rank, S, rb, ls);
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

/* ---------------------------------------------------------------- the solve */
typedef struct {
  int n, nb;
  int* free_idx;   /* n: full-layout index of each free variable */
  int* blk_j; int* blk_i;
  double *mu, *ub; /* nb, 5 nb */
  double *H, *g;   /* compact n x n, n */
} orc_compact;

static void compact(const cmpc_config* c, const orc_inputs* in, const orc_qp* Q, orc_compact* P) {
  int p = Q->p, n = 0;
  P->free_idx = malloc(p * sizeof(int));
  for (int a = 0; a < p; ++a) if (!Q->pinned[a]) P->free_idx[n++] = a;
  P->n = n; P->nb = n / 3;
  P->blk_j = malloc((P->nb + 1) * sizeof(int)); P->blk_i = malloc((P->nb + 1) * sizeof(int));
  P->mu = malloc((P->nb + 1) * 8); P->ub = malloc((5 * P->nb + 1) * 8);
  P->H = malloc(((size_t)n * n + 1) * 8); P->g = malloc((n + 1) * 8);
  for (int b = 0; b < P->nb; ++b) {
    int a = P->free_idx[3 * b], j = a / Q->nu, i = (a % Q->nu) / 3;
    P->blk_j[b] = j; P->blk_i[b] = i; P->mu[b] = c->mu[i];
    double ce = in->contact[i][j];
    for (int r = 0; r < 4; ++r) P->ub[5 * b + r] = FRIC_UB * ce;       /* CentroidalMPC.cpp:183,199 */
    P->ub[5 * b + 4] = c->mass * GRAV * c->num_legs * ce;
  }
  for (int a = 0; a < n; ++a) {
    P->g[a] = Q->g[P->free_idx[a]];
    for (int b = 0; b < n; ++b) P->H[(size_t)a * n + b] = Q->H[(size_t)P->free_idx[a] * p + P->free_idx[b]];
  }
}
static void compact_free(orc_compact* P) {
  free(P->free_idx); free(P->blk_j); free(P->blk_i); free(P->mu); free(P->ub); free(P->H); free(P->g);
}

static double kkt_scaled(const orc_compact* P, const double* u, const double* zl, const double* zu) {
  int n = P->n, nb = P->nb, m = 5 * nb;
  double* r = malloc((n + 1) * 8); double* y = malloc((m + 1) * 8);
  symv(P->H, n, u, r);
  for (int a = 0; a < n; ++a) r[a] += P->g[a];
  CTmul_add(P->mu, nb, zl, -1.0, r);
  CTmul_add(P->mu, nb, zu, +1.0, r);
  Cmul(P->mu, nb, u, y);
  double gs = 1 + maxabs(P->g, n), us = 1 + maxabs(u, n);
  double stat = maxabs(r, n) / gs, prim = 0, dual = 0, comp = 0;
  for (int t = 0; t < m; ++t) {
    double sl = y[t], su = P->ub[t] - y[t];
    if (-sl > prim) prim = -sl;
    if (-su > prim) prim = -su;
    if (-zl[t] > dual) dual = -zl[t];
    if (-zu[t] > dual) dual = -zu[t];
    if (fabs(zl[t] * sl) > comp) comp = fabs(zl[t] * sl);
    if (fabs(zu[t] * su) > comp) comp = fabs(zu[t] * su);
  }
  prim /= us; dual /= gs; comp /= gs * us;
  free(r); free(y);
  double k = stat;
  if (prim > k) k = prim;
  if (dual > k) k = dual;
  if (comp > k) k = comp;
  return k;
}

/* Active-set polish: equality-constrained solve in the per-block null space, multipliers,
 * verification, and up to 6 correction passes.  actl/actu: 5 nb flags in/out.
 * Returns 1 and overwrites u, zl, zu when a verified KKT point is found. */
static int polish(const orc_compact* P, double gs, double us, unsigned char* actl,
                  unsigned char* actu, double* u, double* zl, double* zu, int max_pass) {
  int n = P->n, nb = P->nb, m = 5 * nb;
  double* f0 = malloc((n + 1) * 8); double* Zt = malloc((9 * nb + 1) * 8); /* [b][col][3] */
  int* rk = malloc((nb + 1) * sizeof(int)); int* off = malloc((nb + 1) * sizeof(int));
  double* up = malloc((n + 1) * 8); double* r = malloc((n + 1) * 8);
  double* y = malloc((m + 1) * 8); double* ll = malloc((m + 1) * 8); double* lu = malloc((m + 1) * 8);
  double* Mr = malloc(((size_t)n * n + 1) * 8); double* t = malloc((n + 1) * 8);
  int accepted = 0;
  for (int pass = 0; pass < max_pass && !accepted; ++pass) {
    int ok_all = 1, nr = 0;
    for (int b = 0; b < nb; ++b) {
      double A[10][3], rhs[10]; int k = 0;
      for (int q = 0; q < 5; ++q) if (actl[5 * b + q]) { row_vec(P->mu[b], q, A[k]); rhs[k++] = 0; }
      for (int q = 0; q < 5; ++q) if (actu[5 * b + q]) { row_vec(P->mu[b], q, A[k]); rhs[k++] = P->ub[5 * b + q]; }
      double Z[3][3]; int ok;
      rk[b] = block_nullspace(k, A, rhs, f0 + 3 * b, Z, &ok);
      ok_all &= ok;
      memcpy(Zt + 9 * b, Z, 72);
      off[b] = nr; nr += 3 - rk[b];
    }
    if (!ok_all) break;
    /* reduced system Z'HZ t = -Z'(g + H f0) */
    symv(P->H, n, f0, r);
    for (int a = 0; a < n; ++a) r[a] += P->g[a];
    for (int b = 0; b < nb; ++b)
      for (int cc = 0; cc < 3 - rk[b]; ++cc) {
        const double* z = Zt + 9 * b + 3 * cc;
        t[off[b] + cc] = -(z[0] * r[3 * b] + z[1] * r[3 * b + 1] + z[2] * r[3 * b + 2]);
        for (int b2 = 0; b2 <= b; ++b2)
          for (int c2 = 0; c2 < 3 - rk[b2]; ++c2) {
            const double* z2 = Zt + 9 * b2 + 3 * c2; double s = 0;
            for (int a = 0; a < 3; ++a)
              for (int a2 = 0; a2 < 3; ++a2) s += z[a] * P->H[(size_t)(3 * b + a) * n + 3 * b2 + a2] * z2[a2];
            Mr[(size_t)(off[b] + cc) * nr + off[b2] + c2] = s;
            Mr[(size_t)(off[b2] + c2) * nr + off[b] + cc] = s;
          }
      }
    memcpy(up, f0, n * 8);
    if (nr > 0) {
      if (cholesky(Mr, nr)) break;
      chol_solve(Mr, nr, t);
      for (int b = 0; b < nb; ++b)
        for (int cc = 0; cc < 3 - rk[b]; ++cc)
          for (int a = 0; a < 3; ++a) up[3 * b + a] += Zt[9 * b + 3 * cc + a] * t[off[b] + cc];
    }
    symv(P->H, n, up, r);
    for (int a = 0; a < n; ++a) r[a] += P->g[a];
    Cmul(P->mu, nb, up, y);
    if (!(us > 0)) us = 1 + maxabs(up, n); /* presolve: scale of the candidate itself */
    int okm = 1, changed = 0;
    for (int b = 0; b < nb; ++b) {
      double Nrm[10][3], lam[10]; int idx[10], k = 0;
      for (int q = 0; q < 5; ++q) if (actl[5 * b + q]) { row_vec(P->mu[b], q, Nrm[k]); idx[k++] = q; }
      for (int q = 0; q < 5; ++q) if (actu[5 * b + q]) {
        row_vec(P->mu[b], q, Nrm[k]);
        for (int a = 0; a < 3; ++a) Nrm[k][a] = -Nrm[k][a];
        idx[k++] = 5 + q;
      }
      okm &= block_multipliers(k, Nrm, r + 3 * b, 1e-9 * gs, lam);
      for (int q = 0; q < 5; ++q) ll[5 * b + q] = lu[5 * b + q] = 0;
      for (int s = 0; s < k; ++s) { if (idx[s] < 5) ll[5 * b + idx[s]] = lam[s]; else lu[5 * b + idx[s] - 5] = lam[s]; }
    }
    for (int q = 0; q < m; ++q) {
      double sl = y[q], su = P->ub[q] - y[q];
      int vl = sl < -1e-9 * us, vu = su < -1e-9 * us;
      int nl = ll[q] < -1e-9 * gs, nu_ = lu[q] < -1e-9 * gs;
      if (vl || vu || nl || nu_) changed = 1;
      actl[q] = (actl[q] | vl) & !nl;
      actu[q] = (actu[q] | vu) & !nu_;
    }
    if (!changed && okm) accepted = 1;
  }
  if (accepted) { memcpy(u, up, n * 8); memcpy(zl, ll, m * 8); memcpy(zu, lu, m * 8); }
  free(f0); free(Zt); free(rk); free(off); free(up); free(r); free(y); free(ll); free(lu); free(Mr); free(t);
  return accepted;
}

int cmpc_oracle_solve(const cmpc_config* c, const double* state, const double* des_state,
                      const double* des_inputs, double* forces, int32_t* status_out,
                      int32_t* iters_out, double* kkt_out, double* lam_out, uint16_t* active_out) {
  orc_inputs in; orc_qp Q; orc_compact P;
  int N = c->horizon, L = c->num_legs, p = 3 * L * N;
  unpack(c, state, des_state, des_inputs, &in);
  int status = CMPC_STATUS_MAX_ITER, it = 0;
  double kkt = 0;
  memset(forces, 0, p * 8);
  if (lam_out) memset(lam_out, 0, (size_t)2 * 5 * L * N * 8);
  if (active_out) for (int t = 0; t < L * N; ++t) active_out[t] = 0;
  if (!all_finite(state, 9 + 3 * L) || !all_finite(des_state, 9 * (N + 1)) || !all_finite(des_inputs, L * (4 * N + 3))) {
    *status_out = CMPC_STATUS_NUMERICAL; if (iters_out) *iters_out = 0; if (kkt_out) *kkt_out = 0; return 0;
  }
  if (in.invalid) { /* mirrors the throw at CentroidalMPC.cpp:328-330: outputs zeroed */
    *status_out = CMPC_STATUS_INVALID_TABLE; if (iters_out) *iters_out = 0; if (kkt_out) *kkt_out = 0; return 0;
  }
  build_qp(c, &in, &Q);
  compact(c, &in, &Q, &P);
  int n = P.n, nb = P.nb, m = 5 * nb;
  double *u = calloc(n + 1, 8), *rd = malloc((n + 1) * 8), *rhs = malloc((n + 1) * 8), *du = malloc((n + 1) * 8);
  double *M = malloc(((size_t)n * n + 1) * 8);
  double *sl = malloc((m + 1) * 8), *su = malloc((m + 1) * 8), *zl = malloc((m + 1) * 8), *zu = malloc((m + 1) * 8);
  double *cdu = malloc((m + 1) * 8), *dzl = malloc((m + 1) * 8), *dzu = malloc((m + 1) * 8);
  double *rcl = malloc((m + 1) * 8), *rcu = malloc((m + 1) * 8), *tmp = malloc((m + 1) * 8);
  unsigned char *actl = calloc(m + 1, 1), *actu = calloc(m + 1, 1);

  /* strictly feasible start f = (0, 0, fz0) */
  for (int b = 0; b < nb; ++b) {
    double fz = Q.Uref[P.free_idx[3 * b + 2]];
    if (fz > 0.5 * P.ub[5 * b + 4]) fz = 0.5 * P.ub[5 * b + 4];
    if (fz > 0.5 * P.ub[5 * b] / P.mu[b]) fz = 0.5 * P.ub[5 * b] / P.mu[b];
    u[3 * b + 2] = fz;
  }
  Cmul(P.mu, nb, u, sl);
  for (int t = 0; t < m; ++t) su[t] = P.ub[t] - sl[t];
  double gs = 1 + maxabs(P.g, n);
  symv(P.H, n, u, rd);
  for (int a = 0; a < n; ++a) rd[a] += P.g[a];
  double mu0 = maxabs(rd, n); if (mu0 < 1e-2) mu0 = 1e-2;
  for (int t = 0; t < m; ++t) { zl[t] = mu0 / sl[t]; zu[t] = mu0 / su[t]; }
  int npolish = 0, numerical = 0, ipm_ok = 0;
  /* Presolve: the unconstrained minimiser -H^-1 g is the optimum whenever it is feasible (no row
   * active).  One factorisation of H, verified like any polish; otherwise the IPM runs cold. */
  if (c->polish && c->presolve && polish(&P, gs, 0.0, actl, actu, u, zl, zu, 1)) status = CMPC_STATUS_OK;
  for (it = 0; status != CMPC_STATUS_OK && it <= c->max_iter; ++it) {
    symv(P.H, n, u, rd);
    for (int a = 0; a < n; ++a) rd[a] += P.g[a];
    CTmul_add(P.mu, nb, zl, -1.0, rd);
    CTmul_add(P.mu, nb, zu, +1.0, rd);
    double gap = 0;
    for (int t = 0; t < m; ++t) gap += sl[t] * zl[t] + su[t] * zu[t];
    double mu = gap / (2.0 * m);
    double us = 1 + maxabs(u, n);
    /* Convergence. The dual residual has a round-off floor ~ eps * cond(H + C'SC) once the
     * gap is small, so the polish (which verifies the KKT conditions itself) is attempted
     * as soon as the gap is converged and the residual is merely small. */
    double rmax = maxabs(rd, n);
    int conv_mu = mu <= c->ipm_tol * gs * us;
    int strict = conv_mu && rmax <= c->ipm_tol * gs;
    int ready = conv_mu && rmax <= 1e4 * c->ipm_tol * gs;
    ipm_ok = conv_mu && rmax <= 10 * c->ipm_tol * gs;
    if (c->polish && ready && npolish < 3) {
      ++npolish;
      for (int t = 0; t < m; ++t) { actl[t] = zl[t] * us > sl[t] * gs; actu[t] = zu[t] * us > su[t] * gs; }
      if (polish(&P, gs, us, actl, actu, u, zl, zu, 6)) { status = CMPC_STATUS_OK; break; }
    }
    if (strict && (!c->polish || npolish >= 3)) break;
    if (mu <= 1e-8 * c->ipm_tol * gs * us) break; /* far past convergence: stop before 0/0 */
    if (it == c->max_iter) break;
    /* M = H + C' diag(zl/sl + zu/su) C : only the 3x3 diagonal blocks change */
    memcpy(M, P.H, (size_t)n * n * 8);
    for (int b = 0; b < nb; ++b) {
      double sg[5];
      for (int q = 0; q < 5; ++q) sg[q] = zl[5 * b + q] / sl[5 * b + q] + zu[5 * b + q] / su[5 * b + q];
      double mb = P.mu[b], sx = sg[0] + sg[1], sy = sg[2] + sg[3];
      double* d = M + (size_t)(3 * b) * n + 3 * b;
      d[0] += sx; d[n + 1] += sy; d[2 * n + 2] += mb * mb * (sx + sy) + sg[4];
      d[2] += mb * (sg[1] - sg[0]); d[2 * n] += mb * (sg[1] - sg[0]);
      d[n + 2] += mb * (sg[3] - sg[2]); d[2 * n + 1] += mb * (sg[3] - sg[2]);
    }
    if (cholesky(M, n)) { numerical = 1; break; }
    double alpha = 1, sigma = 0;
    for (int phase = 0; phase < 2; ++phase) {
      /* phase 0: affine predictor; phase 1: centred corrector (Mehrotra) */
      for (int t = 0; t < m; ++t) {
        rcl[t] = -sl[t] * zl[t]; rcu[t] = -su[t] * zu[t];
        if (phase) { rcl[t] += sigma * mu - cdu[t] * dzl[t]; rcu[t] += sigma * mu + cdu[t] * dzu[t]; }
        tmp[t] = rcl[t] / sl[t] - rcu[t] / su[t];
      }
      for (int a = 0; a < n; ++a) rhs[a] = -rd[a];
      CTmul_add(P.mu, nb, tmp, +1.0, rhs);
      memcpy(du, rhs, n * 8);
      chol_solve(M, n, du);
      Cmul(P.mu, nb, du, cdu);
      alpha = 1;
      for (int t = 0; t < m; ++t) {
        dzl[t] = (rcl[t] - zl[t] * cdu[t]) / sl[t];
        dzu[t] = (rcu[t] + zu[t] * cdu[t]) / su[t];
        if (cdu[t] < 0 && -sl[t] / cdu[t] < alpha) alpha = -sl[t] / cdu[t];
        if (cdu[t] > 0 && su[t] / cdu[t] < alpha) alpha = su[t] / cdu[t];
        if (dzl[t] < 0 && -zl[t] / dzl[t] < alpha) alpha = -zl[t] / dzl[t];
        if (dzu[t] < 0 && -zu[t] / dzu[t] < alpha) alpha = -zu[t] / dzu[t];
      }
      if (!phase) {
        double ga = 0;
        for (int t = 0; t < m; ++t)
          ga += (sl[t] + alpha * cdu[t]) * (zl[t] + alpha * dzl[t]) + (su[t] - alpha * cdu[t]) * (zu[t] + alpha * dzu[t]);
        double ratio = ga / gap;
        sigma = ratio * ratio * ratio;
      }
    }
    /* fraction to the boundary tau -> 1 as the gap closes (superlinear tail) */
    { double tau = 1.0 - mu / (gs * us); if (tau < 0.995) tau = 0.995; alpha *= tau; }
    if (alpha > 1) alpha = 1;
    for (int a = 0; a < n; ++a) u[a] += alpha * du[a];
    for (int t = 0; t < m; ++t) { zl[t] += alpha * dzl[t]; zu[t] += alpha * dzu[t]; }
    Cmul(P.mu, nb, u, sl);
    for (int t = 0; t < m; ++t) su[t] = P.ub[t] - sl[t];
    if (!all_finite(u, n)) { numerical = 1; break; }
  }
  if (numerical) status = CMPC_STATUS_NUMERICAL;
  else if (status != CMPC_STATUS_OK) status = ipm_ok ? CMPC_STATUS_OK_IPM : CMPC_STATUS_MAX_ITER;
  if (status <= CMPC_STATUS_MAX_ITER && !numerical) {
    kkt = kkt_scaled(&P, u, zl, zu);
    {
      /* reported active set. Polished: the rows with zero slack at the KKT point (primal
       * definition, unique because the optimum is unique -- the polish's working set can
       * omit redundant rows at the degenerate apex f = 0). Otherwise: the IPM guess. */
      double us = 1 + maxabs(u, n);
      Cmul(P.mu, nb, u, sl);
      for (int t = 0; t < m; ++t) {
        su[t] = P.ub[t] - sl[t];
        if (status == CMPC_STATUS_OK) { actl[t] = sl[t] <= 1e-9 * us; actu[t] = su[t] <= 1e-9 * us; }
        else { actl[t] = zl[t] * us > sl[t] * gs; actu[t] = zu[t] * us > su[t] * gs; }
      }
    }
    for (int b = 0; b < nb; ++b) {
      int j = P.blk_j[b], i = P.blk_i[b];
      for (int r = 0; r < 3; ++r) forces[(size_t)i * 3 * N + 3 * j + r] = u[3 * b + r];
      if (lam_out)
        for (int q = 0; q < 5; ++q) {
          lam_out[((size_t)j * L + i) * 5 + q] = zl[5 * b + q];
          lam_out[(size_t)5 * L * N + ((size_t)j * L + i) * 5 + q] = zu[5 * b + q];
        }
      if (active_out) {
        uint16_t a = 0;
        for (int q = 0; q < 5; ++q) a |= (uint16_t)((actl[5 * b + q] ? 1 : 0) << q | (actu[5 * b + q] ? 1 : 0) << (5 + q));
        active_out[j * L + i] = a;
      }
    }
    if (active_out)
      for (int j = 0; j < N; ++j)
        for (int i = 0; i < L; ++i) if (!(in.contact[i][j] > 0)) active_out[j * L + i] = 0x8000;
  } else {
    memset(forces, 0, p * 8);
  }
  *status_out = status;
  if (iters_out) *iters_out = it;
  if (kkt_out) *kkt_out = kkt;
  free(u); free(rd); free(rhs); free(du); free(M); free(sl); free(su); free(zl); free(zu);
  free(cdu); free(dzl); free(dzu); free(rcl); free(rcu); free(tmp); free(actl); free(actu);
  compact_free(&P); qp_free(&Q);
  return 0;
}

/* ---------------------------------------------------------------- batch driver */
typedef struct {
  const cmpc_config* c; int B; const double *st, *ds, *di;
  double* forces; int32_t *status, *iters; double* kkt; double* lam; uint16_t* active;
  int tid, nthreads;
} orc_job;

static void* worker(void* arg) {
  orc_job* J = arg;
  const cmpc_config* c = J->c;
  int N = c->horizon, L = c->num_legs;
  size_t ns = 9 + 3 * L, nd = 9 * (N + 1), ni = (size_t)L * (4 * N + 3), nf = (size_t)3 * L * N;
  for (int b = J->tid; b < J->B; b += J->nthreads) {
    int32_t st, it; double k;
    cmpc_oracle_solve(c, J->st + b * ns, J->ds + b * nd, J->di + b * ni, J->forces + b * nf, &st, &it, &k,
                      J->lam ? J->lam + (size_t)b * 10 * L * N : NULL, J->active ? J->active + (size_t)b * L * N : NULL);
    J->status[b] = st;
    if (J->iters) J->iters[b] = it;
    if (J->kkt) J->kkt[b] = k;
  }
  return NULL;
}

int cmpc_oracle_solve_batch(const cmpc_config* c, int B, const double* state, const double* des_state,
                            const double* des_inputs, double* forces, int32_t* status, int32_t* iters,
                            double* kkt, double* lam, uint16_t* active, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256]; orc_job jobs[256];
  for (int t = 0; t < nthreads; ++t) {
    jobs[t] = (orc_job){c, B, state, des_state, des_inputs, forces, status, iters, kkt, lam, active, t, nthreads};
    if (nthreads == 1) worker(&jobs[t]); else pthread_create(&th[t], NULL, worker, &jobs[t]);
  }
  if (nthreads > 1) for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  return 0;
}

/* The reference plant, verbatim (CentroidalMPC.cpp:85-92): explicit Euler with the TRUE
 * lever arm foot - com.  forces [L][3]. */
void cmpc_oracle_plant_step(const cmpc_config* c, const double* x, const double* feet,
                            const double* contact, const double* forces, double* xn) {
  double acc[3] = {0, 0, -GRAV}, ld[3] = {0, 0, 0};
  for (int i = 0; i < c->num_legs; ++i) {
    const double* f = forces + 3 * i; double ce = contact[i];
    double r[3] = {feet[3 * i] - x[0], feet[3 * i + 1] - x[1], feet[3 * i + 2] - x[2]};
    for (int a = 0; a < 3; ++a) acc[a] += ce / c->mass * f[a];
    ld[0] += ce * (r[1] * f[2] - r[2] * f[1]);
    ld[1] += ce * (r[2] * f[0] - r[0] * f[2]);
    ld[2] += ce * (r[0] * f[1] - r[1] * f[0]);
  }
  for (int a = 0; a < 3; ++a) {
    xn[a] = x[a] + x[3 + a] * c->dt;
    xn[3 + a] = x[3 + a] + acc[a] * c->dt;
    xn[6 + a] = x[6 + a] + ld[a] * c->dt;
  }
}
