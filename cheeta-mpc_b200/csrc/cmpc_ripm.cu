// cmpc_ripm.cu -- stage-wise (Riccati) interior-point + active-set polish kernel (sm_100a).  SURVEY §8 f3.
//
// The same feasible-start Mehrotra primal-dual iteration and the same verified active-set polish as the
// condensed kernel (cmpc_solve.cu) and the oracle, but every linear system is solved in the un-condensed
// optimal-control form of the QP -- the stage-wise (A, B, Q, R, S) view of the reference's HPIPM adapter
// (ocs2_sqp/hpipm_catkin/src/HpipmInterface.cpp:166-301) -- instead of through a Cholesky factor of the
// 3LN x 3LN matrix H + C'SC:
//     minimise 1/2 d'(H + C' diag(sigma) C) d - rhs' d          (one Newton step, or one polish pass)
// is a homogeneous LQR problem over the deviation state z_k = [xi_k; d_{k-1}] (9 + 12 entries; the force-rate
// term of CentroidalMPC.cpp:227-231 couples consecutive inputs), because C' diag(sigma) C only touches the
// 3 x 3 input-cost block of each leg-step.  One backward sweep factors it (a 12 x 12 Cholesky per stage),
// every further right-hand side costs a vector sweep: O(N 21^3) per iteration instead of O((3LN)^3), so a
// horizon-30 instance with active friction rows no longer pays dense n = 144..360 factorisations.
//   * Forces keep the FULL step-major layout [N][L][3]; swing legs are masked per stage (uniform branches at
//     leg granularity), so no index compaction exists anywhere and every register index is static.
//   * The 21 x 21 cost-to-go lives in registers, one row per lane; products with the structured input
//     matrix [dt^2 zeta c/m I; dt c/m I; dt c [r]x; e] are closed forms (15 flops per row and leg).
//   * The polish's equality-constrained solve uses the same sweep: with P_i the orthogonal projector onto
//     the null space of leg-step i's working rows, G <- P G P + (I - P), M <- P M solves the reduced system
//     Z'HZ t = -Z'(H f0 + g) without renumbering anything.
//   * Stage factors (Y = L^-1 M, y0, L) go to an L2-resident slab and are re-used by the corrector.
// One warp per instance, persistent CTAs, device work counter.  Results equal the condensed route's (the
// Newton directions are the same vectors up to round-off): tests/test_gpu_parity.py::test_ripm_*.
#include "cmpc_device.cuh"

namespace cmpc {

namespace {

constexpr int kNF = 12;                       // input slots of one stage: 4 legs x 3 (legs >= L are never in stance)
constexpr int kNZ = 9 + kNF;                  // augmented deviation state [xi; d_prev]
constexpr int kYS = 22;                       // row stride of Y: 21 columns + y0
constexpr int kFac = kNF * kYS + kNF * kNF;   // doubles per stage in the slab: Y [12][22], L [12][12]
constexpr unsigned kFull = 0xffffffffu;

struct Rip {
  const double* in;  // staged inputs [state | des_state | des_inputs]
  int N, L, nf, ns, nds, nfN, nbfull;
  double *u, *du, *dua, *rd;  // force-space vectors, full layout
  double *Rs;                 // 6 per leg-step: 1/2 C'SC entries (interior point) or the projector (polish)
  double *W, *G, *Y, *Lm;     // stage work area (contiguous); X (9N) aliases it between sweeps
  double *m0, *ps, *zs, *us, *tab;
  double *zl, *zu, *fac;      // slab (global, L2-resident)
  uint16_t* act;
  const double* qz;           // CTA-shared: omega_node^2, node 0..N
};

// tab: arm [4][3] at 0, coef (bp, bv, s) [4][3] at 12, Rl (xx, yy, zz, zx, zy, -) [4][6] at 24, Pi (00, 11, 22, 10, 20, 21)
// [4][6] at 48, fz at 72
constexpr int kTabArm = 0, kTabCoef = 12, kTabR = 24, kTabPi = 48, kTabFz = 72, kTabSize = 74;

__device__ __forceinline__ double contact_at(const Rip& R, int i, int k) { return R.in[R.ns + R.nds + i * (4 * R.N + 3) + k]; }
__device__ __forceinline__ double xref_at(const Rip& R, int node, int r) {
  return R.in[R.ns + (r / 3) * 3 * (R.N + 1) + 3 * node + (r % 3)];
}
__device__ __forceinline__ double qdiag_at(const DevConfig& cfg, const double* qz, int node, int r) {
  return r != 2 ? cfg.w[r] : qz[node];  // omega inside the square (CentroidalMPC.cpp:205-210)
}

// Per-stage tables; returns the stance mask (bit i: leg i has contact > 0 at step k).
// mode 0: dynamics only; 1: + input Hessian blocks Wf + rate Wr + 1/2 C'SC (Rs); 2: + Wf + rate Wr and the projectors (Rs).
__device__ __forceinline__ unsigned stage_tables(const Rip& R, const DevConfig& cfg, int k, int lane, int mode) {
  __syncwarp();
  bool st = false;
  double ce = 0.0;
  if (lane < 4 && lane < R.L) ce = contact_at(R, lane, k);
  st = lane < 4 && ce > 0.0;
  double colsum = (lane < 4) ? ce : 0.0;
  colsum += __shfl_xor_sync(kFull, colsum, 1);
  colsum += __shfl_xor_sync(kFull, colsum, 2);
  if (lane == 0) R.tab[kTabFz] = cfg.mass * kGrav / colsum;  // desired fz of the stance legs (:331-333)
  if (st) {
    const int i = lane, N = R.N, L = R.L;
    const double* foot = R.in + R.ns + R.nds + i * (4 * N + 3) + N + 3 * k;
    const double* com = R.in + R.ns + 3 * k;
    const double dt = cfg.dt, cm = ce / cfg.mass;
    R.tab[kTabArm + 3 * i] = foot[0] - com[0];  // frozen lever arm
    R.tab[kTabArm + 3 * i + 1] = foot[1] - com[1];
    R.tab[kTabArm + 3 * i + 2] = foot[2] - com[2];
    R.tab[kTabCoef + 3 * i] = (cfg.zoh ? 0.5 : 0.0) * dt * dt * cm;
    R.tab[kTabCoef + 3 * i + 1] = dt * cm;
    R.tab[kTabCoef + 3 * i + 2] = dt * ce;
    if (mode >= 1) {
      const double rate = k >= 1 ? 1.0 : 0.0;
      const double* rs = R.Rs + 6 * (k * L + i);
      double d0 = cfg.w[9 + 3 * L + 3 * i] + rate * cfg.w[9 + 6 * L + 3 * i];
      double d1 = cfg.w[9 + 3 * L + 3 * i + 1] + rate * cfg.w[9 + 6 * L + 3 * i + 1];
      double d2 = cfg.w[9 + 3 * L + 3 * i + 2] + rate * cfg.w[9 + 6 * L + 3 * i + 2];
      double zx = 0.0, zy = 0.0;
      if (mode == 1) { d0 += rs[0]; d1 += rs[1]; d2 += rs[2]; zx = rs[3]; zy = rs[4]; }
      double* rl = R.tab + kTabR + 6 * i;
      rl[0] = d0; rl[1] = d1; rl[2] = d2; rl[3] = zx; rl[4] = zy;
      if (mode == 2) {
        double* pi = R.tab + kTabPi + 6 * i;
#pragma unroll
        for (int q = 0; q < 6; ++q) pi[q] = rs[q];
      }
    }
  }
  const unsigned mask = __ballot_sync(kFull, st);
  __syncwarp();
  return mask;
}

// (B' v)[3i + q] without the selection row:  bp v[q] + bv v[3 + q] + s (v[6:9] x arm)[q]
__device__ __forceinline__ void bt_mul(const double* tab, int i, const double* v, double& g0, double& g1, double& g2) {
  const double a0 = tab[kTabArm + 3 * i], a1 = tab[kTabArm + 3 * i + 1], a2 = tab[kTabArm + 3 * i + 2];
  const double bp = tab[kTabCoef + 3 * i], bv = tab[kTabCoef + 3 * i + 1], s = tab[kTabCoef + 3 * i + 2];
  g0 = bp * v[0] + bv * v[3] + s * (v[7] * a2 - v[8] * a1);
  g1 = bp * v[1] + bv * v[4] + s * (v[8] * a0 - v[6] * a2);
  g2 = bp * v[2] + bv * v[5] + s * (v[6] * a1 - v[7] * a0);
}
// symmetric 3 x 3 projector (00, 11, 22, 10, 20, 21) times a vector
__device__ __forceinline__ void proj3(const double* p, double& x, double& y, double& z) {
  const double nx = p[0] * x + p[3] * y + p[4] * z;
  const double ny = p[3] * x + p[1] * y + p[5] * z;
  const double nz = p[4] * x + p[5] * y + p[2] * z;
  x = nx; y = ny; z = nz;
}

// (B u)[r] for state row r < 9 with the stage inputs u (12 slots, zeros on swing legs)
__device__ __forceinline__ double b_mul_row(const double* tab, unsigned mask, const double* u, int r) {
  const int grp = r / 3, q = r - 3 * grp;
  double acc = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (!((mask >> i) & 1u)) continue;
    const double u0 = u[3 * i], u1 = u[3 * i + 1], u2 = u[3 * i + 2];
    const double a0 = tab[kTabArm + 3 * i], a1 = tab[kTabArm + 3 * i + 1], a2 = tab[kTabArm + 3 * i + 2];
    const double uq = q == 0 ? u0 : (q == 1 ? u1 : u2);
    // (arm x u)[q]
    const double cx = q == 0 ? a1 * u2 - a2 * u1 : (q == 1 ? a2 * u0 - a0 * u2 : a0 * u1 - a1 * u0);
    acc += grp == 0 ? tab[kTabCoef + 3 * i] * uq : (grp == 1 ? tab[kTabCoef + 3 * i + 1] * uq : tab[kTabCoef + 3 * i + 2] * cx);
  }
  return acc;
}

// Backward sweep: factor the stage systems of  1/2 d'(H + C'SC) d - rhs'd  (mode 1) or of the projected polish
// system (mode 2); the right-hand side in R.du rides along (y0 = column 21 of Y).  Factors -> slab.
// Returns false (uniformly) on a non-positive pivot.
__device__ __noinline__ bool lqr_factor(const Rip& R, const DevConfig& cfg, int lane, int mode) {
  const int N = R.N, L = R.L, nf = R.nf;
  const double dt = cfg.dt;
  double P[kNZ];
#pragma unroll
  for (int c = 0; c < kNZ; ++c) P[c] = 0.0;
  {
    const double qd = lane < 9 ? qdiag_at(cfg, R.qz, N, lane) : 0.0;  // V_N = xi' Q_N xi
#pragma unroll
    for (int c = 0; c < 9; ++c) if (lane == c) P[c] = qd;
  }
  double pv = 0.0;
  bool ok = true;
  for (int k = N - 1; k >= 0; --k) {
    const unsigned mask = stage_tables(R, cfg, k, lane, mode);
    const double rate = k >= 1 ? 1.0 : 0.0;
    const double* tab = R.tab;
    // ---- T1 = P Bbar, row `lane`
    if (lane < kNZ) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        double w0 = 0.0, w1 = 0.0, w2 = 0.0;
        if ((mask >> i) & 1u) {
          bt_mul(tab, i, P, w0, w1, w2);
          w0 += P[9 + 3 * i]; w1 += P[10 + 3 * i]; w2 += P[11 + 3 * i];
        }
        R.W[lane * kNF + 3 * i] = w0; R.W[lane * kNF + 3 * i + 1] = w1; R.W[lane * kNF + 3 * i + 2] = w2;
      }
      R.ps[lane] = pv;
    }
    __syncwarp();
    // ---- m0 = Bbar' p - rhs / 2 (slot `lane`), G = Bbar' T1 + R (pairs leg x column)
    if (lane < kNF) {
      const int i = lane / 3, q = lane - 3 * i;
      double mv = 0.0;
      if ((mask >> i) & 1u) {
        double g0, g1, g2;
        bt_mul(tab, i, R.ps, g0, g1, g2);
        mv = (q == 0 ? g0 : (q == 1 ? g1 : g2)) + R.ps[9 + lane] - 0.5 * R.du[k * nf + lane];
      }
      R.m0[lane] = mv;
    }
    for (int e = lane; e < 4 * kNF; e += 32) {
      const int i = e / kNF, c2 = e - kNF * i, i2 = c2 / 3;
      if (((mask >> i) & 1u) && ((mask >> i2) & 1u)) {
        double v[9];
#pragma unroll
        for (int s = 0; s < 9; ++s) v[s] = R.W[s * kNF + c2];
        double g0, g1, g2;
        bt_mul(tab, i, v, g0, g1, g2);
        g0 += R.W[(9 + 3 * i) * kNF + c2]; g1 += R.W[(10 + 3 * i) * kNF + c2]; g2 += R.W[(11 + 3 * i) * kNF + c2];
        if (i2 == i) {
          const double* rl = tab + kTabR + 6 * i;
          const int cq = c2 - 3 * i;
          g0 += cq == 0 ? rl[0] : (cq == 2 ? rl[3] : 0.0);
          g1 += cq == 1 ? rl[1] : (cq == 2 ? rl[4] : 0.0);
          g2 += cq == 0 ? rl[3] : (cq == 1 ? rl[4] : rl[2]);
        }
        R.G[(3 * i) * kNF + c2] = g0; R.G[(3 * i + 1) * kNF + c2] = g1; R.G[(3 * i + 2) * kNF + c2] = g2;
      }
    }
    __syncwarp();
    if (mode == 2) {
      // G <- Pi G Pi + (I - Pi): rows (lane = column), then columns (lane = row)
      if (lane < kNF && ((mask >> (lane / 3)) & 1u)) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (!((mask >> i) & 1u)) continue;
          double x = R.G[(3 * i) * kNF + lane], y = R.G[(3 * i + 1) * kNF + lane], z = R.G[(3 * i + 2) * kNF + lane];
          proj3(tab + kTabPi + 6 * i, x, y, z);
          R.G[(3 * i) * kNF + lane] = x; R.G[(3 * i + 1) * kNF + lane] = y; R.G[(3 * i + 2) * kNF + lane] = z;
        }
      }
      __syncwarp();
      if (lane < kNF && ((mask >> (lane / 3)) & 1u)) {
        const int il = lane / 3, ql = lane - 3 * il;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (!((mask >> i) & 1u)) continue;
          double x = R.G[lane * kNF + 3 * i], y = R.G[lane * kNF + 3 * i + 1], z = R.G[lane * kNF + 3 * i + 2];
          const double* pi = tab + kTabPi + 6 * i;
          proj3(pi, x, y, z);
          if (i == il) {  // + (I - Pi), row ql
            x += (ql == 0 ? 1.0 : 0.0) - (ql == 0 ? pi[0] : (ql == 1 ? pi[3] : pi[4]));
            y += (ql == 1 ? 1.0 : 0.0) - (ql == 0 ? pi[3] : (ql == 1 ? pi[1] : pi[5]));
            z += (ql == 2 ? 1.0 : 0.0) - (ql == 0 ? pi[4] : (ql == 1 ? pi[5] : pi[2]));
          }
          R.G[lane * kNF + 3 * i] = x; R.G[lane * kNF + 3 * i + 1] = y; R.G[lane * kNF + 3 * i + 2] = z;
        }
      }
      __syncwarp();
    }
    // ---- Cholesky of G, row `lane` in registers; columns are broadcast through Lm.  Swing legs are skipped.
    {
      double g[kNF];
      const bool mine = lane < kNF && ((mask >> (lane / 3)) & 1u);
#pragma unroll
      for (int b = 0; b < kNF; ++b) g[b] = mine ? R.G[lane * kNF + b] : 0.0;
#pragma unroll
      for (int c = 0; c < kNF; ++c) {
        if (!((mask >> (c / 3)) & 1u)) continue;
        const double d = __shfl_sync(kFull, g[c], c);
        ok = ok && d > 0.0;
        const double inv = fast_rsqrt(d);
        const double l = g[c] * inv;
        if (lane == c) R.Lm[c * kNF + c] = inv;
        else if (lane > c && lane < kNF) R.Lm[lane * kNF + c] = mine ? l : 0.0;
        __syncwarp();
#pragma unroll
        for (int c2 = c + 1; c2 < kNF; ++c2) {
          if (!((mask >> (c2 / 3)) & 1u)) continue;
          g[c2] = fma(-l, R.Lm[c2 * kNF + c], g[c2]);
        }
      }
    }
    if (!ok) break;  // uniform: d is a broadcast value
    // ---- Y = L^-1 [M | m0], lane = column (column 21 = m0); M = T1(0:9)' Abar - rate [0, Wr]
    double y[kNF];
    {
      const int col = lane;
#pragma unroll
      for (int ia = 0; ia < 4; ++ia) {
        if (!((mask >> ia) & 1u)) { y[3 * ia] = 0.0; y[3 * ia + 1] = 0.0; y[3 * ia + 2] = 0.0; continue; }
        double mf[3];
#pragma unroll
        for (int qa = 0; qa < 3; ++qa) {
          const int c = 3 * ia + qa;
          double v = 0.0;
          if (col < 9) v = R.W[col * kNF + c];
          if (col >= 3 && col < 6) v = fma(dt, R.W[(col - 3) * kNF + c], v);
          if (col == 9 + c) v = (c < nf) ? -rate * cfg.w[9 + 6 * L + c] : 0.0;
          if (col == kNZ) v = R.m0[c];
          mf[qa] = v;
        }
        if (mode == 2) proj3(tab + kTabPi + 6 * ia, mf[0], mf[1], mf[2]);
#pragma unroll
        for (int qa = 0; qa < 3; ++qa) {
          const int a = 3 * ia + qa;
          double acc = mf[qa];
#pragma unroll
          for (int ib = 0; ib <= ia; ++ib) {
            if (!((mask >> ib) & 1u)) continue;
#pragma unroll
            for (int qb = 0; qb < 3; ++qb) {
              const int b = 3 * ib + qb;
              if (b < a) acc = fma(-R.Lm[a * kNF + b], y[b], acc);
            }
          }
          y[a] = acc * R.Lm[a * kNF + a];
        }
      }
      double* fk = R.fac + (size_t)k * kFac;
      if (col <= kNZ) {
#pragma unroll
        for (int a = 0; a < kNF; ++a) { R.Y[a * kYS + col] = y[a]; __stcg(fk + a * kYS + col, y[a]); }
      }
      for (int e = lane; e < kNF * kNF; e += 32) __stcg(fk + kNF * kYS + e, R.Lm[e]);
    }
    __syncwarp();
    if (k >= 1) {
      // ---- P <- blkdiag(Q_k, Wr) + Abar' P Abar - Y'Y (row `lane`);  p <- Abar' p - Y' y0
      P[3] = fma(dt, P[0], P[3]); P[4] = fma(dt, P[1], P[4]); P[5] = fma(dt, P[2], P[5]);
#pragma unroll
      for (int c = 0; c < 9; ++c) {
        const double up = __shfl_sync(kFull, P[c], (lane + 29) & 31);  // row lane - 3
        if (lane >= 3 && lane < 6) P[c] = fma(dt, up, P[c]);
        if (lane >= 9) P[c] = 0.0;
      }
#pragma unroll
      for (int c = 9; c < kNZ; ++c) P[c] = 0.0;
      const double dg = lane < 9 ? qdiag_at(cfg, R.qz, k, lane) : ((lane - 9 < nf) ? cfg.w[9 + 6 * L + ((lane - 9) < nf ? lane - 9 : 0)] : 0.0);
#pragma unroll
      for (int c = 0; c < kNZ; ++c) if (lane == c) P[c] += dg;
      double pn = 0.0;
      if (lane < 9) pn = R.ps[lane] + ((lane >= 3 && lane < 6) ? dt * R.ps[lane - 3] : 0.0);
#pragma unroll
      for (int ia = 0; ia < 4; ++ia) {
        if (!((mask >> ia) & 1u)) continue;
#pragma unroll
        for (int qa = 0; qa < 3; ++qa) {
          const int a = 3 * ia + qa;
          const double ya = y[a];
          const double2* Y2 = reinterpret_cast<const double2*>(R.Y + a * kYS);
#pragma unroll
          for (int c = 0; c < 10; ++c) {
            const double2 yy = Y2[c];
            P[2 * c] = fma(-ya, yy.x, P[2 * c]); P[2 * c + 1] = fma(-ya, yy.y, P[2 * c + 1]);
          }
          const double2 yl = Y2[10];  // column 20 and y0
          P[20] = fma(-ya, yl.x, P[20]);
          pn = fma(-ya, yl.y, pn);
        }
      }
      pv = pn;
    }
  }
  return ok;
}

// Backward vector sweep for a new right-hand side (R.du) with the stored factors: y0 of every stage -> slab.
__device__ __noinline__ void lqr_backsolve(const Rip& R, const DevConfig& cfg, int lane) {
  const int N = R.N, nf = R.nf;
  const double dt = cfg.dt;
  double pv = 0.0;
  for (int k = N - 1; k >= 0; --k) {
    const unsigned mask = stage_tables(R, cfg, k, lane, 0);
    const double* tab = R.tab;
    double* fk = R.fac + (size_t)k * kFac;
    const bool mine = lane < kNF && ((mask >> (lane / 3)) & 1u);
    // row `lane` of L and column `lane` of Y (issued early: L2 latency overlaps the m0 computation)
    double Lr[kNF], Yc[kNF];
#pragma unroll
    for (int b = 0; b < kNF; ++b) Lr[b] = (mine && b <= lane) ? __ldcg(fk + kNF * kYS + lane * kNF + b) : 0.0;
#pragma unroll
    for (int a = 0; a < kNF; ++a) Yc[a] = (k >= 1 && lane < kNZ) ? __ldcg(fk + a * kYS + lane) : 0.0;
    if (lane < kNZ) R.ps[lane] = pv;
    __syncwarp();
    double acc = 0.0;
    if (mine) {
      const int i = lane / 3, q = lane - 3 * i;
      double g0, g1, g2;
      bt_mul(tab, i, R.ps, g0, g1, g2);
      acc = (q == 0 ? g0 : (q == 1 ? g1 : g2)) + R.ps[9 + lane] - 0.5 * R.du[k * nf + lane];
    }
    // y0 = L^-1 m0: column-oriented substitution, one broadcast per column
    double inv = 1.0;
#pragma unroll
    for (int b = 0; b < kNF; ++b) if (lane == b) inv = Lr[b];
    double y0 = 0.0;
#pragma unroll
    for (int b = 0; b < kNF; ++b) {
      if (!((mask >> (b / 3)) & 1u)) continue;
      const double yb = __shfl_sync(kFull, acc * inv, b);
      if (lane > b) acc = fma(-Lr[b], yb, acc);
      if (lane == b) y0 = yb;
    }
    if (lane < kNF) { __stcg(fk + lane * kYS + kNZ, y0); R.m0[lane] = y0; }
    __syncwarp();
    if (k >= 1) {
      double pn = 0.0;
      if (lane < 9) pn = R.ps[lane] + ((lane >= 3 && lane < 6) ? dt * R.ps[lane - 3] : 0.0);
#pragma unroll
      for (int a = 0; a < kNF; ++a) {
        if (!((mask >> (a / 3)) & 1u)) continue;
        pn = fma(-Yc[a], R.m0[a], pn);
      }
      pv = lane < kNZ ? pn : 0.0;
    }
  }
}

// Forward sweep: t_k = -L^-T (Y z_k + y0), d_k -> dst (full layout), z_{k+1} = [A xi + B d; d].
__device__ __noinline__ void lqr_forward(const Rip& R, const DevConfig& cfg, int lane, double* dst) {
  const int N = R.N, nf = R.nf;
  const double dt = cfg.dt;
  double zr = 0.0;
  for (int k = 0; k < N; ++k) {
    const unsigned mask = stage_tables(R, cfg, k, lane, 0);
    const double* tab = R.tab;
    const double* fk = R.fac + (size_t)k * kFac;
    const bool mine = lane < kNF && ((mask >> (lane / 3)) & 1u);
    double Lc[kNF];  // column `lane` of L (rows below the diagonal), 1 / l on the diagonal
#pragma unroll
    for (int b = 0; b < kNF; ++b) Lc[b] = (mine && b >= lane) ? __ldcg(fk + kNF * kYS + b * kNF + lane) : 0.0;
    double acc = 0.0;
    if (lane < kNZ) R.zs[lane] = zr;
    __syncwarp();
    if (mine) {
      const double2* row = reinterpret_cast<const double2*>(fk + lane * kYS);
      const double2* z2 = reinterpret_cast<const double2*>(R.zs);
#pragma unroll
      for (int c = 0; c < 10; ++c) { const double2 yy = __ldcg(row + c), zz = z2[c]; acc = fma(yy.x, zz.x, acc); acc = fma(yy.y, zz.y, acc); }
      const double2 yl = __ldcg(row + 10);
      acc = fma(yl.x, R.zs[20], acc) + yl.y;
    }
    double inv = 1.0;
#pragma unroll
    for (int b = 0; b < kNF; ++b) if (lane == b) inv = Lc[b];
    double uo = 0.0;
#pragma unroll
    for (int b = kNF - 1; b >= 0; --b) {
      if (!((mask >> (b / 3)) & 1u)) continue;
      const double ub = __shfl_sync(kFull, acc * inv, b);
      if (lane < b) acc = fma(-Lc[b], ub, acc);
      if (lane == b) uo = -ub;
    }
    if (lane < kNF) R.us[lane] = uo;
    if (lane < nf) dst[k * nf + lane] = uo;
    __syncwarp();
    double zn = 0.0;
    if (lane < 9) zn = R.zs[lane] + (lane < 3 ? dt * R.zs[lane + 3] : 0.0) + b_mul_row(tab, mask, R.us, lane);
    else if (lane < kNZ) zn = R.us[lane - 9];
    zr = zn;
  }
}

// dst = H src + g on the stance entries (0 elsewhere) by one roll-out and one adjoint sweep over the problem
// data (CentroidalMPC.cpp:85-92 dynamics with frozen arms, :203-232 cost) -- independent of the factors.
// X (9N doubles) aliases the stage work area.
__device__ __noinline__ void stage_gradient(const Rip& R, const DevConfig& cfg, int lane, const double* src, double* dst) {
  const int N = R.N, nf = R.nf, L = R.L;
  const double dt = cfg.dt;
  const double zeta = cfg.zoh ? 0.5 : 0.0;
  const double dpz = zeta * dt * dt * (-kGrav), dvz = dt * (-kGrav);
  double* X = R.W;
  double xr = lane < 9 ? R.in[lane] : 0.0;
  for (int k = 0; k < N; ++k) {
    const unsigned mask = stage_tables(R, cfg, k, lane, 0);
    if (lane < 9) R.zs[lane] = xr;
    if (lane < kNF) R.us[lane] = lane < nf ? src[k * nf + lane] : 0.0;
    __syncwarp();
    if (lane < 9) {
      xr = R.zs[lane] + (lane < 3 ? dt * R.zs[lane + 3] : 0.0) + b_mul_row(R.tab, mask, R.us, lane) + (lane == 2 ? dpz : (lane == 5 ? dvz : 0.0));
      X[9 * k + lane] = xr;
    }
  }
  double lam = 0.0;
  for (int k = N - 1; k >= 0; --k) {
    const unsigned mask = stage_tables(R, cfg, k, lane, 0);
    if (lane < 9) R.zs[lane] = lam;
    __syncwarp();
    if (lane < 9) {
      lam = 2.0 * qdiag_at(cfg, R.qz, k + 1, lane) * (X[9 * k + lane] - xref_at(R, k + 1, lane)) + R.zs[lane] +
            ((lane >= 3 && lane < 6) ? dt * R.zs[lane - 3] : 0.0);
      R.ps[lane] = lam;
    }
    __syncwarp();
    if (lane < nf) {
      const int i = lane / 3, q = lane - 3 * i;
      double gv = 0.0;
      if ((mask >> i) & 1u) {
        const double wf = cfg.w[9 + 3 * L + lane], wr = cfg.w[9 + 6 * L + lane];
        const double f = src[k * nf + lane], fr = (q == 2) ? R.tab[kTabFz] : 0.0;
        double g0, g1, g2;
        bt_mul(R.tab, i, R.ps, g0, g1, g2);
        gv = 2.0 * wf * (f - fr) + (q == 0 ? g0 : (q == 1 ? g1 : g2));
        if (k >= 1) gv += 2.0 * wr * (f - src[(k - 1) * nf + lane]);
        if (k + 1 < N) gv -= 2.0 * wr * (src[(k + 1) * nf + lane] - f);
      }
      dst[k * nf + lane] = gv;
    }
  }
  __syncwarp();
}

}  // namespace

// shared-memory plan of one group (doubles), see the kernel
__host__ __device__ inline void ripm_plan(int N, int L, int* off /*[16]*/, int* total, int* slab) {
  const int nfN = 3 * L * N, nbfull = L * N;
  const int nin = (9 + 3 * L) + 9 * (N + 1) + L * (4 * N + 3);
  int o = 0;
  auto take = [&](int cnt) { int r = o; o += (cnt + 1) & ~1; return r; };
  off[0] = take(nin);
  off[1] = take(nfN); off[2] = take(nfN); off[3] = take(nfN); off[4] = take(nfN);  // u, du, dua, rd
  off[5] = take(6 * nbfull);                                                        // Rs
  int work = kNZ * kNF + kNF * kNF + kNF * kYS + kNF * kNF;                         // W, G, Y, Lm (contiguous)
  if (work < 9 * N) work = 9 * N;
  off[6] = take(work);
  off[7] = take(kNF); off[8] = take(kNZ + 1); off[9] = take(kNZ + 1); off[10] = take(kNF); off[11] = take(kTabSize);  // m0, ps, zs, us, tab
  off[12] = take((2 * nbfull + 16 + 7) / 8);                                        // act (uint16), misc
  *total = o;
  *slab = N * kFac + 2 * 5 * nbfull;
}

__global__ void __launch_bounds__(384, 1) cmpc_ripm_kernel(const __grid_constant__ DevConfig cfg, const __grid_constant__ SolveArgs args) {
  extern __shared__ __align__(128) double smem[];
  const int N = cfg.N, L = cfg.L;
  const int nf = 3 * L, nfN = nf * N, nbfull = L * N, mfull = 5 * nbfull;
  const int ns = 9 + 3 * L, nds = 9 * (N + 1), ndi = L * (4 * N + 3);
  const int lane = threadIdx.x & 31, gid = threadIdx.x >> 5;
  int off[16], total, slabsz;
  ripm_plan(N, L, off, &total, &slabsz);
  const int ctad = (N + 2) & ~1;  // CTA-shared: omega^2 of nodes 0..N
  double* c_qz = smem;
  double* base = smem + ctad + (size_t)gid * total;
  double* slab = args.scratch + (size_t)(blockIdx.x * args.groups + gid) * args.scratch_per_group;
  Rip R;
  R.in = base + off[0]; R.N = N; R.L = L; R.nf = nf; R.ns = ns; R.nds = nds; R.nfN = nfN; R.nbfull = nbfull;
  R.u = base + off[1]; R.du = base + off[2]; R.dua = base + off[3]; R.rd = base + off[4];
  R.Rs = base + off[5];
  R.W = base + off[6]; R.G = R.W + kNZ * kNF; R.Y = R.G + kNF * kNF; R.Lm = R.Y + kNF * kYS;
  R.m0 = base + off[7]; R.ps = base + off[8]; R.zs = base + off[9]; R.us = base + off[10]; R.tab = base + off[11];
  R.act = reinterpret_cast<uint16_t*>(base + off[12]);
  R.fac = slab; R.zl = slab + (size_t)N * kFac; R.zu = R.zl + mfull;
  R.qz = c_qz;
  double* s_in = base + off[0];
  const double mass = cfg.mass;
  const int count = args.count ? *args.count : args.count_imm;
  if (count <= 0) return;
  for (int e = threadIdx.x; e <= N; e += blockDim.x) {
    const double om = (cfg.w[2] * 0.5) * exp(-(double)e) + cfg.w[2] * 0.5;  // CentroidalMPC.cpp:205
    c_qz[e] = om * om;
  }
  __syncthreads();

  while (true) {
    int slot = 0;
    if (lane == 0) slot = atomicAdd(args.work, 1);
    slot = __shfl_sync(kFull, slot, 0);
    if (slot >= count) break;
    const int inst = args.perm ? args.perm[slot] : slot;
    // ---- stage the raw inputs once (coalesced, L2-only loads), validate (CentroidalMPC.cpp:284-330)
    bool finite = true;
    {
      const double* src = args.state + (size_t)inst * ns;
      for (int t = lane; t < ns; t += 32) { const double v = __ldcg(src + t); s_in[t] = v; finite = finite && isfinite(v); }
      src = args.des_state + (size_t)inst * nds;
      for (int t = lane; t < nds; t += 32) { const double v = __ldcg(src + t); s_in[ns + t] = v; finite = finite && isfinite(v); }
      src = args.des_inputs + (size_t)inst * ndi;
      for (int t = lane; t < ndi; t += 32) { const double v = __ldcg(src + t); s_in[ns + nds + t] = v; finite = finite && isfinite(v); }
    }
    __syncwarp();
    finite = __all_sync(kFull, finite);
    bool invalid = false;
    int nb = 0;
    {
      double colsum = 0.0;
      if (lane < N)
        for (int i = 0; i < L; ++i) { const double ce = contact_at(R, i, lane); colsum += ce; nb += ce > 0.0 ? 1 : 0; }
      invalid = __any_sync(kFull, lane < N && !(colsum > 0.0));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) nb += __shfl_xor_sync(kFull, nb, o);
    }
    if (!finite || invalid) {
      for (int t = lane; t < nfN; t += 32) args.forces[(size_t)inst * nfN + t] = 0.0;
      if (args.lam) for (int t = lane; t < 2 * mfull; t += 32) args.lam[(size_t)inst * 2 * mfull + t] = 0.0;
      if (args.active) for (int t = lane; t < nbfull; t += 32) args.active[(size_t)inst * nbfull + t] = 0;
      if (lane == 0) {
        args.status[inst] = !finite ? CMPC_STATUS_NUMERICAL : CMPC_STATUS_INVALID_TABLE;
        if (args.iters) args.iters[inst] = 0;
        if (args.kkt) args.kkt[inst] = 0.0;
      }
      __syncwarp();
      continue;
    }
    const int m = 5 * nb;

    // ---- g (gradient at 0) for the scale gs, strictly feasible start f = (0, 0, fz0), centred duals
    for (int t = lane; t < nfN; t += 32) { R.dua[t] = 0.0; R.du[t] = 0.0; }
    __syncwarp();
    stage_gradient(R, cfg, lane, R.dua, R.rd);
    double gmax = 0.0;
    for (int t = lane; t < nfN; t += 32) gmax = fmax(gmax, fabs(R.rd[t]));
#pragma unroll 1
    for (int tb = lane; tb < nbfull; tb += 32) {
      const int k = tb / L, i = tb - k * L;
      const double ce = contact_at(R, i, k);
      double fz = 0.0;
      if (ce > 0.0) {
        double colsum = 0.0;
        for (int i2 = 0; i2 < L; ++i2) colsum += contact_at(R, i2, k);
        fz = mass * kGrav / colsum;
        fz = fmin(fz, 0.5 * mass * kGrav * (double)L * ce);
        fz = fmin(fz, 0.5 * kFricUb * ce / cfg.mu[i]);
      }
      R.u[3 * tb] = 0.0; R.u[3 * tb + 1] = 0.0; R.u[3 * tb + 2] = fz;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) gmax = fmax(gmax, __shfl_xor_sync(kFull, gmax, o));
    __syncwarp();
    const double gs = 1.0 + gmax;
    stage_gradient(R, cfg, lane, R.u, R.rd);  // H u0 + g
    double r0max = 0.0;
    for (int t = lane; t < nfN; t += 32) r0max = fmax(r0max, fabs(R.rd[t]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r0max = fmax(r0max, __shfl_xor_sync(kFull, r0max, o));
    const double mu0 = fmax(1e-2, r0max);
#pragma unroll 1
    for (int tb = lane; tb < nbfull; tb += 32) {
      const int k = tb / L, i = tb - k * L;
      const double ce = contact_at(R, i, k);
      double y[5] = {1.0, 1.0, 1.0, 1.0, 1.0};
      const double ubxy = kFricUb * ce, ubz = mass * kGrav * (double)L * ce;  // :183,199
      if (ce > 0.0) cmul5(cfg.mu[i], R.u + 3 * tb, y);
      for (int q = 0; q < 5; ++q) {
        const double ub = q < 4 ? ubxy : ubz;
        R.zl[5 * tb + q] = ce > 0.0 ? mu0 / y[q] : 0.0;
        R.zu[5 * tb + q] = ce > 0.0 ? mu0 / (ub - y[q]) : 0.0;
      }
    }
    __syncwarp();

    int status = CMPC_STATUS_MAX_ITER, it = 0, npolish = 0;
    bool numerical = false, ipm_ok = false, grad_fresh = true;  // rd holds H u + g of the current u
    double us = 1.0;
#pragma unroll 1
    for (it = 0; it <= cfg.max_iter; ++it) {
      if (!grad_fresh) stage_gradient(R, cfg, lane, R.u, R.rd);
      grad_fresh = false;
      double rmax = 0.0, umax = 0.0, gap = 0.0;
#pragma unroll 1
      for (int tb = lane; tb < nbfull; tb += 32) {
        const int k = tb / L, i = tb - k * L;
        const double ce = contact_at(R, i, k);
        if (!(ce > 0.0)) continue;
        const double ubxy = kFricUb * ce, ubz = mass * kGrav * (double)L * ce;
        double w[5], o[3], ys[5];
        cmul5(cfg.mu[i], R.u + 3 * tb, ys);
        for (int q = 0; q < 5; ++q) {
          const int t = 5 * tb + q;
          const double sl = ys[q], su = (q < 4 ? ubxy : ubz) - sl, zl = R.zl[t], zu = R.zu[t];
          w[q] = zl - zu;
          gap += sl * zl + su * zu;
        }
        ctmul5(cfg.mu[i], w, o);
        for (int q = 0; q < 3; ++q) {
          const double rr = R.rd[3 * tb + q] - o[q];
          R.rd[3 * tb + q] = rr;
          rmax = fmax(rmax, fabs(rr)); umax = fmax(umax, fabs(R.u[3 * tb + q]));
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        rmax = fmax(rmax, __shfl_xor_sync(kFull, rmax, o));
        umax = fmax(umax, __shfl_xor_sync(kFull, umax, o));
        gap += __shfl_xor_sync(kFull, gap, o);
      }
      __syncwarp();
      const double mu = gap / (2.0 * (double)m);
      us = 1.0 + umax;
      const bool conv_mu = mu <= cfg.tol * gs * us;
      const bool strict = conv_mu && rmax <= cfg.tol * gs;
      const bool ready = conv_mu && rmax <= 1e4 * cfg.tol * gs;
      ipm_ok = conv_mu && rmax <= 10.0 * cfg.tol * gs;
      if (cfg.polish && ready && npolish < 3) {
        ++npolish;
        // ---- active-set guess from the interior-point iterate
#pragma unroll 1
        for (int tb = lane; tb < nbfull; tb += 32) {
          const int k = tb / L, i = tb - k * L;
          const double ce = contact_at(R, i, k);
          uint16_t a = 0;
          if (ce > 0.0) {
            const double ubxy = kFricUb * ce, ubz = mass * kGrav * (double)L * ce;
            double ys[5];
            cmul5(cfg.mu[i], R.u + 3 * tb, ys);
            for (int q = 0; q < 5; ++q) {
              const int t = 5 * tb + q;
              const double sl = ys[q], su = (q < 4 ? ubxy : ubz) - sl;
              if (R.zl[t] * us > sl * gs) a |= (uint16_t)(1u << q);
              if (R.zu[t] * us > su * gs) a |= (uint16_t)(1u << (5 + q));
            }
          }
          R.act[tb] = a;
        }
        __syncwarp();
        bool accepted = false;
#pragma unroll 1
        for (int pass = 0; pass < 6 && !accepted; ++pass) {
          // per leg-step: particular solution f0 (-> rd) and the projector onto the null space of the working rows (-> Rs)
          bool ok_all = true;
#pragma unroll 1
          for (int tb = lane; tb < nbfull; tb += 32) {
            const int k = tb / L, i = tb - k * L;
            const double ce = contact_at(R, i, k);
            double f0[3] = {0.0, 0.0, 0.0};
            if (ce > 0.0) {
              const double mub = cfg.mu[i];
              const double ubxy = kFricUb * ce, ubz = mass * kGrav * (double)L * ce;
              const unsigned a = R.act[tb];
              double A[10][3], rhsb[10], Z[3][3];
              int kk = 0;
              for (int q = 0; q < 5; ++q) if ((a >> q) & 1u) { row_vec(mub, q, A[kk]); rhsb[kk++] = 0.0; }
              for (int q = 0; q < 5; ++q) if ((a >> (5 + q)) & 1u) { row_vec(mub, q, A[kk]); rhsb[kk++] = q < 4 ? ubxy : ubz; }
              bool okb;
              const int rk = block_nullspace(kk, A, rhsb, f0, Z, &okb);
              ok_all = ok_all && okb;
              double p00 = 0.0, p11 = 0.0, p22 = 0.0, p10 = 0.0, p20 = 0.0, p21 = 0.0;
              for (int cc = 0; cc < 3 - rk; ++cc) {
                p00 += Z[cc][0] * Z[cc][0]; p11 += Z[cc][1] * Z[cc][1]; p22 += Z[cc][2] * Z[cc][2];
                p10 += Z[cc][1] * Z[cc][0]; p20 += Z[cc][2] * Z[cc][0]; p21 += Z[cc][2] * Z[cc][1];
              }
              double* rs = R.Rs + 6 * tb;
              rs[0] = p00; rs[1] = p11; rs[2] = p22; rs[3] = p10; rs[4] = p20; rs[5] = p21;
            }
            R.rd[3 * tb] = f0[0]; R.rd[3 * tb + 1] = f0[1]; R.rd[3 * tb + 2] = f0[2];
          }
          __syncwarp();
          ok_all = __all_sync(kFull, ok_all);
          if (!ok_all) break;
          stage_gradient(R, cfg, lane, R.rd, R.dua);  // H f0 + g
          for (int t = lane; t < nfN; t += 32) R.du[t] = -R.dua[t];
          __syncwarp();
          if (!lqr_factor(R, cfg, lane, 2)) break;
          lqr_forward(R, cfg, lane, R.du);
          __syncwarp();
          for (int t = lane; t < nfN; t += 32) R.du[t] += R.rd[t];  // candidate point up = f0 + Z t
          __syncwarp();
          stage_gradient(R, cfg, lane, R.du, R.dua);  // H up + g
          // multipliers, verification, correction; when the pass verifies the loop runs once more to commit
          bool good = false;
#pragma unroll 1
          for (int commit = 0; commit < 2; ++commit) {
            bool okm = true, changed = false;
#pragma unroll 1
            for (int tb = lane; tb < nbfull; tb += 32) {
              const int k = tb / L, i = tb - k * L;
              const double ce = contact_at(R, i, k);
              if (!(ce > 0.0)) continue;
              const double mub = cfg.mu[i];
              const double ubxy = kFricUb * ce, ubz = mass * kGrav * (double)L * ce;
              const unsigned a = R.act[tb];
              double rb[3], y[5], ll[5] = {0, 0, 0, 0, 0}, lu[5] = {0, 0, 0, 0, 0};
              for (int q = 0; q < 3; ++q) rb[q] = R.dua[3 * tb + q];
              if ((a & 0x3ffu) == 0u) {
                okm = okm && fmax(fabs(rb[0]), fmax(fabs(rb[1]), fabs(rb[2]))) <= 1e-9 * gs;
              } else {
                double Nrm[10][3], lam[10];
                int idx[10], kk = 0;
                for (int q = 0; q < 5; ++q) if ((a >> q) & 1u) { row_vec(mub, q, Nrm[kk]); idx[kk++] = q; }
                for (int q = 0; q < 5; ++q) if ((a >> (5 + q)) & 1u) {
                  row_vec(mub, q, Nrm[kk]);
                  Nrm[kk][0] = -Nrm[kk][0]; Nrm[kk][1] = -Nrm[kk][1]; Nrm[kk][2] = -Nrm[kk][2];
                  idx[kk++] = 5 + q;
                }
                okm = block_multipliers(kk, Nrm, rb, 1e-9 * gs, lam) && okm;
                for (int sI = 0; sI < kk; ++sI) { if (idx[sI] < 5) ll[idx[sI]] = lam[sI]; else lu[idx[sI] - 5] = lam[sI]; }
              }
              if (commit) {
                for (int q = 0; q < 5; ++q) { R.zl[5 * tb + q] = ll[q]; R.zu[5 * tb + q] = lu[q]; }
                continue;
              }
              cmul5(mub, R.du + 3 * tb, y);
              unsigned an = 0;
              for (int q = 0; q < 5; ++q) {
                const double ub = q < 4 ? ubxy : ubz;
                const double sl = y[q], su = ub - y[q];
                const bool vl = sl < -1e-9 * us, vu = su < -1e-9 * us;
                const bool nl = ll[q] < -1e-9 * gs, nuu = lu[q] < -1e-9 * gs;
                if (vl || vu || nl || nuu) changed = true;
                const bool al = (((a >> q) & 1u) || vl) && !nl, au = (((a >> (5 + q)) & 1u) || vu) && !nuu;
                an |= (al ? 1u : 0u) << q | (au ? 1u : 0u) << (5 + q);
              }
              R.act[tb] = (uint16_t)an;
            }
            __syncwarp();
            if (commit) break;
            good = __all_sync(kFull, okm && !changed);
            if (!good) break;
          }
          if (good) accepted = true;
        }
        if (accepted) {
          for (int t = lane; t < nfN; t += 32) { R.u[t] = R.du[t]; R.rd[t] = R.dua[t]; }  // rd = H u + g at the KKT point
          __syncwarp();
          status = CMPC_STATUS_OK;
          break;
        }
        __syncwarp();
        // polish not accepted: rd was the f0 scratch -> recompute the dual residual
        stage_gradient(R, cfg, lane, R.u, R.rd);
#pragma unroll 1
        for (int tb = lane; tb < nbfull; tb += 32) {
          const int k = tb / L, i = tb - k * L;
          if (!(contact_at(R, i, k) > 0.0)) continue;
          double w[5], o[3];
          for (int q = 0; q < 5; ++q) w[q] = R.zl[5 * tb + q] - R.zu[5 * tb + q];
          ctmul5(cfg.mu[i], w, o);
          for (int q = 0; q < 3; ++q) R.rd[3 * tb + q] -= o[q];
        }
        __syncwarp();
      }
      if (strict && (!cfg.polish || npolish >= 3)) break;
      if (mu <= 1e-8 * cfg.tol * gs * us) break;  // far past convergence: stop before 0/0
      if (it == cfg.max_iter) break;

      // ---- 1/2 C' diag(zl/sl + zu/su) C per leg-step -> Rs;  affine right-hand side -rd + C'(zu - zl) -> du
#pragma unroll 1
      for (int tb = lane; tb < nbfull; tb += 32) {
        const int k = tb / L, i = tb - k * L;
        const double ce = contact_at(R, i, k);
        if (!(ce > 0.0)) { R.du[3 * tb] = 0.0; R.du[3 * tb + 1] = 0.0; R.du[3 * tb + 2] = 0.0; continue; }
        const double ubxy = kFricUb * ce, ubz = mass * kGrav * (double)L * ce;
        const double mb = cfg.mu[i];
        double sg[5], tq[5], o[3], ys[5];
        cmul5(mb, R.u + 3 * tb, ys);
        for (int q = 0; q < 5; ++q) {
          const int t = 5 * tb + q;
          const double sl = ys[q], su = (q < 4 ? ubxy : ubz) - sl, zl = R.zl[t], zu = R.zu[t];
          sg[q] = zl * fast_rcp(sl) + zu * fast_rcp(su);
          tq[q] = zu - zl;
        }
        const double sx = sg[0] + sg[1], sy = sg[2] + sg[3];
        double* rs = R.Rs + 6 * tb;
        rs[0] = 0.5 * sx; rs[1] = 0.5 * sy; rs[2] = 0.5 * (mb * mb * (sx + sy) + sg[4]);
        rs[3] = 0.5 * mb * (sg[1] - sg[0]); rs[4] = 0.5 * mb * (sg[3] - sg[2]);
        ctmul5(mb, tq, o);
        for (int q = 0; q < 3; ++q) R.du[3 * tb + q] = -R.rd[3 * tb + q] + o[q];
      }
      __syncwarp();
      if (!lqr_factor(R, cfg, lane, 1)) { numerical = true; break; }
      lqr_forward(R, cfg, lane, R.du);
      __syncwarp();

      double tmax = 0.0, sigma = 0.0;
#pragma unroll 1
      for (int phase = 0; phase < 2; ++phase) {
        // phase 0: affine predictor (solved above); phase 1: centred corrector (Mehrotra)
        if (phase) {
          for (int t = lane; t < nfN; t += 32) R.dua[t] = R.du[t];
          __syncwarp();
#pragma unroll 1
          for (int tb = lane; tb < nbfull; tb += 32) {
            const int k = tb / L, i = tb - k * L;
            const double ce = contact_at(R, i, k);
            if (!(ce > 0.0)) continue;  // du = dua = 0 there
            const double mub = cfg.mu[i];
            const double ubxy = kFricUb * ce, ubz = mass * kGrav * (double)L * ce;
            double tq[5], o[3], ys[5], ya[5];
            cmul5(mub, R.u + 3 * tb, ys);
            cmul5(mub, R.dua + 3 * tb, ya);
            for (int q = 0; q < 5; ++q) {
              const int t = 5 * tb + q;
              const double sl = ys[q], su = (q < 4 ? ubxy : ubz) - sl, zl = R.zl[t], zu = R.zu[t];
              const double isl = fast_rcp(sl), isu = fast_rcp(su);
              const double dla = (-sl * zl - zl * ya[q]) * isl, dua_ = (-su * zu + zu * ya[q]) * isu;
              const double rcl = -sl * zl + sigma * mu - ya[q] * dla;
              const double rcu = -su * zu + sigma * mu + ya[q] * dua_;
              tq[q] = rcl * isl - rcu * isu;
            }
            ctmul5(mub, tq, o);
            for (int q = 0; q < 3; ++q) R.du[3 * tb + q] = -R.rd[3 * tb + q] + o[q];
          }
          __syncwarp();
          lqr_backsolve(R, cfg, lane);
          lqr_forward(R, cfg, lane, R.du);
          __syncwarp();
        }
        // step to the boundary: alpha_max = 1 / max_i(-ds_i/s_i, -dz_i/z_i)
        double tloc = 0.0;
#pragma unroll 1
        for (int tb = lane; tb < nbfull; tb += 32) {
          const int k = tb / L, i = tb - k * L;
          const double ce = contact_at(R, i, k);
          if (!(ce > 0.0)) continue;
          const double mub = cfg.mu[i];
          const double ubxy = kFricUb * ce, ubz = mass * kGrav * (double)L * ce;
          double ys[5], yd[5], ya[5] = {0, 0, 0, 0, 0};
          cmul5(mub, R.u + 3 * tb, ys);
          cmul5(mub, R.du + 3 * tb, yd);
          if (phase) cmul5(mub, R.dua + 3 * tb, ya);
          for (int q = 0; q < 5; ++q) {
            const int t = 5 * tb + q;
            const double sl = ys[q], su = (q < 4 ? ubxy : ubz) - sl, zl = R.zl[t], zu = R.zu[t];
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
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tloc = fmax(tloc, __shfl_xor_sync(kFull, tloc, o));
        tmax = tloc;
        if (!phase) {
          const double alpha = tmax > 1.0 ? 1.0 / tmax : 1.0;
          double ga = 0.0;
#pragma unroll 1
          for (int tb = lane; tb < nbfull; tb += 32) {
            const int k = tb / L, i = tb - k * L;
            const double ce = contact_at(R, i, k);
            if (!(ce > 0.0)) continue;
            const double mub = cfg.mu[i];
            const double ubxy = kFricUb * ce, ubz = mass * kGrav * (double)L * ce;
            double ys[5], yd[5];
            cmul5(mub, R.u + 3 * tb, ys);
            cmul5(mub, R.du + 3 * tb, yd);
            for (int q = 0; q < 5; ++q) {
              const int t = 5 * tb + q;
              const double sl = ys[q], su = (q < 4 ? ubxy : ubz) - sl, zl = R.zl[t], zu = R.zu[t];
              const double cd = yd[q];
              const double dl = (-sl * zl - zl * cd) * fast_rcp(sl), du_ = (-su * zu + zu * cd) * fast_rcp(su);
              ga += (sl + alpha * cd) * (zl + alpha * dl) + (su - alpha * cd) * (zu + alpha * du_);
            }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) ga += __shfl_xor_sync(kFull, ga, o);
          const double ratio = ga / gap;
          sigma = ratio * ratio * ratio;
        }
      }
      // fraction to the boundary tau -> 1 as the gap closes (superlinear tail)
      const double tau = fmax(0.995, 1.0 - mu / (gs * us));
      const double alpha = fmin(1.0, tau / fmax(tmax, 1e-300));
      bool fin = true;
#pragma unroll 1
      for (int tb = lane; tb < nbfull; tb += 32) {
        const int k = tb / L, i = tb - k * L;
        const double ce = contact_at(R, i, k);
        if (!(ce > 0.0)) continue;
        const double mub = cfg.mu[i];
        const double ubxy = kFricUb * ce, ubz = mass * kGrav * (double)L * ce;
        double ys[5], yd[5], ya[5];
        cmul5(mub, R.u + 3 * tb, ys);  // slacks at the current point (before the update)
        cmul5(mub, R.du + 3 * tb, yd);
        cmul5(mub, R.dua + 3 * tb, ya);
        for (int q = 0; q < 5; ++q) {
          const int t = 5 * tb + q;
          const double sl = ys[q], su = (q < 4 ? ubxy : ubz) - sl, zl = R.zl[t], zu = R.zu[t];
          const double isl = fast_rcp(sl), isu = fast_rcp(su);
          const double dla = (-sl * zl - zl * ya[q]) * isl, dua_ = (-su * zu + zu * ya[q]) * isu;
          const double rcl = -sl * zl + sigma * mu - ya[q] * dla;
          const double rcu = -su * zu + sigma * mu + ya[q] * dua_;
          R.zl[t] = zl + alpha * (rcl - zl * yd[q]) * isl;
          R.zu[t] = zu + alpha * (rcu + zu * yd[q]) * isu;
        }
        for (int q = 0; q < 3; ++q) {
          const double v = R.u[3 * tb + q] + alpha * R.du[3 * tb + q];
          R.u[3 * tb + q] = v; fin = fin && isfinite(v);
        }
      }
      __syncwarp();
      fin = __all_sync(kFull, fin);
      if (!fin) { numerical = true; break; }
    }
    if (numerical) status = CMPC_STATUS_NUMERICAL;
    else if (status != CMPC_STATUS_OK) status = ipm_ok ? CMPC_STATUS_OK_IPM : CMPC_STATUS_MAX_ITER;

    // ---- outputs
    if (!numerical) {
      if (status != CMPC_STATUS_OK) stage_gradient(R, cfg, lane, R.u, R.rd);  // an accepted polish left H u + g in rd
      double stat = 0.0, umax = 0.0, prim = 0.0, dual = 0.0, comp = 0.0;
#pragma unroll 1
      for (int tb = lane; tb < nbfull; tb += 32) {
        const int k = tb / L, i = tb - k * L;
        const double ce = contact_at(R, i, k);
        if (!(ce > 0.0)) continue;
        const double mub = cfg.mu[i];
        const double ubxy = kFricUb * ce, ubz = mass * kGrav * (double)L * ce;
        double w[5], o[3], y[5];
        for (int q = 0; q < 5; ++q) w[q] = R.zl[5 * tb + q] - R.zu[5 * tb + q];
        ctmul5(mub, w, o);
        cmul5(mub, R.u + 3 * tb, y);
        for (int q = 0; q < 3; ++q) {
          stat = fmax(stat, fabs(R.rd[3 * tb + q] - o[q]));
          umax = fmax(umax, fabs(R.u[3 * tb + q]));
        }
        for (int q = 0; q < 5; ++q) {
          const double ub = q < 4 ? ubxy : ubz;
          const double sl = y[q], su = ub - y[q], zl = R.zl[5 * tb + q], zu = R.zu[5 * tb + q];
          prim = fmax(prim, fmax(-sl, -su));
          dual = fmax(dual, fmax(-zl, -zu));
          comp = fmax(comp, fmax(fabs(zl * sl), fabs(zu * su)));
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        stat = fmax(stat, __shfl_xor_sync(kFull, stat, o)); umax = fmax(umax, __shfl_xor_sync(kFull, umax, o));
        prim = fmax(prim, __shfl_xor_sync(kFull, prim, o)); dual = fmax(dual, __shfl_xor_sync(kFull, dual, o));
        comp = fmax(comp, __shfl_xor_sync(kFull, comp, o));
      }
      const double usf = 1.0 + umax;
      const double kkt = fmax(fmax(stat / gs, prim / usf), fmax(dual / gs, comp / (gs * usf)));
      // reported active set: polished -> rows with zero slack at the KKT point; otherwise the interior-point guess
#pragma unroll 1
      for (int tb = lane; tb < nbfull; tb += 32) {
        const int k = tb / L, i = tb - k * L;
        const double ce = contact_at(R, i, k);
        uint16_t a = 0x8000;
        if (ce > 0.0) {
          const double ubxy = kFricUb * ce, ubz = mass * kGrav * (double)L * ce;
          double ys[5];
          cmul5(cfg.mu[i], R.u + 3 * tb, ys);
          a = 0;
          for (int q = 0; q < 5; ++q) {
            const int t = 5 * tb + q;
            const double sl = ys[q], su = (q < 4 ? ubxy : ubz) - sl;
            bool al, au;
            if (status == CMPC_STATUS_OK) { al = sl <= 1e-9 * usf; au = su <= 1e-9 * usf; }
            else { al = R.zl[t] * usf > sl * gs; au = R.zu[t] * usf > su * gs; }
            a |= (uint16_t)((al ? 1 : 0) << q | (au ? 1 : 0) << (5 + q));
          }
        }
        R.act[tb] = a;
      }
      __syncwarp();
      // forces in the reference's per-leg order [L][N][3] (CentroidalMPC.cpp:270)
      for (int t = lane; t < nfN; t += 32) {
        const int i = t / (3 * N), rem = t - i * 3 * N, j = rem / 3, q = rem - 3 * j;
        args.forces[(size_t)inst * nfN + t] = R.u[j * nf + 3 * i + q];
      }
      if (args.lam) {
        for (int t = lane; t < mfull; t += 32) {
          args.lam[(size_t)inst * 2 * mfull + t] = R.zl[t];
          args.lam[(size_t)inst * 2 * mfull + mfull + t] = R.zu[t];
        }
      }
      if (args.active) for (int t = lane; t < nbfull; t += 32) args.active[(size_t)inst * nbfull + t] = R.act[t];
      if (lane == 0) {
        args.status[inst] = status;
        if (args.iters) args.iters[inst] = it;
        if (args.kkt) args.kkt[inst] = kkt;
      }
    } else {
      for (int t = lane; t < nfN; t += 32) args.forces[(size_t)inst * nfN + t] = 0.0;
      if (args.lam) for (int t = lane; t < 2 * mfull; t += 32) args.lam[(size_t)inst * 2 * mfull + t] = 0.0;
      if (args.active) for (int t = lane; t < nbfull; t += 32) args.active[(size_t)inst * nbfull + t] = 0;
      if (lane == 0) {
        args.status[inst] = status;
        if (args.iters) args.iters[inst] = it;
        if (args.kkt) args.kkt[inst] = 0.0;
      }
    }
    __syncwarp();
  }
}


// Diagnostic / parity entry (cmpc_stage_step_batch): the stage-wise linear algebra alone.  For every instance:
// d_fused = the solution of the stage system with the right-hand side riding along the factor sweep,
// d_resolve = the same right-hand side through the stored factors (corrector path), grad = H rhs + g
// (roll-out + adjoint sweep with rhs taken as a force vector).  hess: 6 doubles per leg-step in the Rs layout
// (mode 1: 1/2 C'SC entries xx, yy, zz, zx, zy, -; mode 2: projector 00, 11, 22, 10, 20, 21).
__global__ void __launch_bounds__(384, 1) cmpc_ripm_probe_kernel(const __grid_constant__ DevConfig cfg, const __grid_constant__ SolveArgs args,
                                                                 const double* hess, const double* rhs, int mode, double* d_fused,
                                                                 double* d_resolve, double* grad, int B) {
  extern __shared__ __align__(128) double smem[];
  const int N = cfg.N, L = cfg.L;
  const int nf = 3 * L, nfN = nf * N, nbfull = L * N, mfull = 5 * nbfull;
  const int ns = 9 + 3 * L, nds = 9 * (N + 1), ndi = L * (4 * N + 3);
  const int lane = threadIdx.x & 31, gid = threadIdx.x >> 5;
  int off[16], total, slabsz;
  ripm_plan(N, L, off, &total, &slabsz);
  const int ctad = (N + 2) & ~1;
  double* c_qz = smem;
  double* base = smem + ctad + (size_t)gid * total;
  double* slab = args.scratch + (size_t)(blockIdx.x * args.groups + gid) * args.scratch_per_group;
  Rip R;
  R.in = base + off[0]; R.N = N; R.L = L; R.nf = nf; R.ns = ns; R.nds = nds; R.nfN = nfN; R.nbfull = nbfull;
  R.u = base + off[1]; R.du = base + off[2]; R.dua = base + off[3]; R.rd = base + off[4];
  R.Rs = base + off[5];
  R.W = base + off[6]; R.G = R.W + kNZ * kNF; R.Y = R.G + kNF * kNF; R.Lm = R.Y + kNF * kYS;
  R.m0 = base + off[7]; R.ps = base + off[8]; R.zs = base + off[9]; R.us = base + off[10]; R.tab = base + off[11];
  R.act = reinterpret_cast<uint16_t*>(base + off[12]);
  R.fac = slab; R.zl = slab + (size_t)N * kFac; R.zu = R.zl + mfull;
  R.qz = c_qz;
  double* s_in = base + off[0];
  for (int e = threadIdx.x; e <= N; e += blockDim.x) {
    const double om = (cfg.w[2] * 0.5) * exp(-(double)e) + cfg.w[2] * 0.5;
    c_qz[e] = om * om;
  }
  __syncthreads();
  for (int inst = blockIdx.x * args.groups + gid; inst < B; inst += gridDim.x * args.groups) {
    for (int t = lane; t < ns; t += 32) s_in[t] = args.state[(size_t)inst * ns + t];
    for (int t = lane; t < nds; t += 32) s_in[ns + t] = args.des_state[(size_t)inst * nds + t];
    for (int t = lane; t < ndi; t += 32) s_in[ns + nds + t] = args.des_inputs[(size_t)inst * ndi + t];
    for (int t = lane; t < 6 * nbfull; t += 32) R.Rs[t] = hess[(size_t)inst * 6 * nbfull + t];
    for (int t = lane; t < nfN; t += 32) { R.du[t] = rhs[(size_t)inst * nfN + t]; R.u[t] = R.du[t]; }
    __syncwarp();
    const bool ok = lqr_factor(R, cfg, lane, mode);
    lqr_forward(R, cfg, lane, R.du);
    __syncwarp();
    for (int t = lane; t < nfN; t += 32) { d_fused[(size_t)inst * nfN + t] = ok ? R.du[t] : nan(""); R.du[t] = R.u[t]; }
    __syncwarp();
    if (ok) {
      lqr_backsolve(R, cfg, lane);
      lqr_forward(R, cfg, lane, R.du);
      __syncwarp();
    }
    for (int t = lane; t < nfN; t += 32) d_resolve[(size_t)inst * nfN + t] = ok ? R.du[t] : nan("");
    __syncwarp();
    stage_gradient(R, cfg, lane, R.u, R.rd);
    for (int t = lane; t < nfN; t += 32) grad[(size_t)inst * nfN + t] = R.rd[t];
    __syncwarp();
  }
}

cudaError_t launch_ripm_probe(int grid, int block, size_t smem, cudaStream_t stream, const DevConfig& cfg, const SolveArgs& args,
                              const double* hess, const double* rhs, int mode, double* d_fused, double* d_resolve, double* grad, int B) {
  cudaError_t e = cudaFuncSetAttribute(cmpc_ripm_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  cmpc_ripm_probe_kernel<<<grid, block, smem, stream>>>(cfg, args, hess, rhs, mode, d_fused, d_resolve, grad, B);
  return cudaGetLastError();
}

void ripm_sizes(int N, int L, int* group_doubles, int* cta_doubles, int* slab_doubles) {
  int off[16];
  ripm_plan(N, L, off, group_doubles, slab_doubles);
  *cta_doubles = (N + 2) & ~1;
}

cudaError_t launch_ripm_kernel(int grid, int block, size_t smem, cudaStream_t stream, const DevConfig& cfg, const SolveArgs& args) {
  cmpc_ripm_kernel<<<grid, block, smem, stream>>>(cfg, args);
  return cudaGetLastError();
}

cudaError_t set_ripm_kernel_smem(size_t bytes) {
  return cudaFuncSetAttribute(cmpc_ripm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

}  // namespace cmpc
