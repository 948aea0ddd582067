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

extern __shared__ __align__(128) double smem[];

namespace {

constexpr int kNF = 12;                       // input slots of one stage: 4 legs x 3 (legs >= L are never in stance)
constexpr int kNZ = 9 + kNF;                  // augmented deviation state [xi; d_prev]
constexpr int kYS = 22;                       // row stride of Y: 21 columns + y0
constexpr int kOffL = kNF * kYS;              // stage factors in the slab: Y | y0 [12][22], strictly lower L [12][12] (zeros elsewhere,
constexpr int kOffD = kOffL + kNF * kNF;      //   also on swing-leg rows / columns: the vector sweeps run unmasked), 1 / l_aa [12]
constexpr int kFac = kOffD + kNF;
constexpr unsigned kFull = 0xffffffffu;
// The polish's first working set: rows with kGuess z / gs > s / us.  The condensed kernel and the oracle use 1 (a row is
// guessed active when its scaled multiplier exceeds its scaled slack); here the guess is generous, because a weakly
// active row that is missed costs a whole correction pass (two gradient sweeps and a factor sweep) while a row guessed
// in vain only shows up with a negative multiplier in the same pass that would have been needed anyway.  The accepted
// point is the verified KKT point either way (unique optimum); measured on the tracking-heavy horizon-30 workload the
// polish takes 1.5 instead of 2.2 passes per instance, results identical.
constexpr double kGuess = 100.0;

// CTA-shared header (constants every sweep needs; read with broadcast LDS instead of through the kernel parameters,
// which a separately compiled device function can only reach with generic loads)
constexpr int kHdrQz = 0;    // omega_node^2, node 0..32 (CentroidalMPC.cpp:205-210)
constexpr int kHdrW = 34;    // state weights w[0..9), then wf[12], wr[12] (zero for legs >= L)
constexpr int kHdrSc = 68;   // dt, zeta dt^2 / m, dt / m, mass
constexpr int kHdr = 72;

// Shared-memory layout of one group (offsets in doubles from smem[0]); every device function recomputes it from
// (N, L, group base) -- a dozen integer operations per call -- so that all accesses are LDS/STS.
struct Lay {
  int st;    // state (9 + 3L) | des_state (9 (N + 1))
  int tab;   // [N][4][4]: lever arm (3), contact (1) per leg slot
  int fz;    // [N] desired fz of the stance legs (:331-333)
  int u, du, rd;  // force-space vectors, full layout [N][L][3] (the copy of the affine direction lives in the slab)
  int rsb;   // [2][24]: input-Hessian data of the current / next stage (cp.async from the slab)
  int W, G, Y, Lm;     // stage work area (contiguous); X (9N) aliases it between sweeps
  int m0, ps, zs, us, act, total;
};
__host__ __device__ __forceinline__ Lay make_lay(int N, int L, int gb) {
  const int nfN = (3 * L * N + 1) & ~1, nbfull = L * N;
  (void)nbfull;
  Lay y;
  int o = gb;
  auto take = [&](int cnt) { int r = o; o += (cnt + 1) & ~1; return r; };
  y.st = take(9 + 3 * L + 9 * (N + 1));
  y.tab = take(16 * N);
  y.fz = take(N);
  y.u = take(nfN); y.du = take(nfN); y.rd = take(nfN);
  y.rsb = take(48);
  int work = 2 * kFac;  // W, G, Y, Lm of the factor sweep (804) <= the vector sweeps' double buffer of stage factors
  if (work < 9 * N) work = 9 * N;
  y.W = take(work); y.G = y.W + kNZ * kNF; y.Y = y.G + kNF * kNF; y.Lm = y.Y + kNF * kYS;
  y.m0 = take(kNF); y.ps = take(kNZ + 1); y.zs = take(kNZ + 1); y.us = take(kNF);
  y.act = take((2 * nbfull + 7) / 8);
  y.total = o - gb;
  return y;
}

// leg-step index -> stage: a shift for four legs (a division by a run-time value is ~20 instructions, and every per-leg-step
// loop of the iteration has one)
__device__ __forceinline__ int div_legs(int tb, int L) { return L == 4 ? tb >> 2 : tb / L; }
struct Leg { double a0, a1, a2, bp, bv, s; };
__device__ __forceinline__ Leg leg_at(const double* tk, int i, double kp, double kv, double ks) {
  const double2 p = *reinterpret_cast<const double2*>(tk + 4 * i), q = *reinterpret_cast<const double2*>(tk + 4 * i + 2);
  Leg g;
  g.a0 = p.x; g.a1 = p.y; g.a2 = q.x;
  g.bp = kp * q.y; g.bv = kv * q.y; g.s = ks * q.y;
  return g;
}
// reg += v on lane `c` only, as one predicated add: written as a plain `if (lane == c) P[c] += v` chain the compiler sees a
// dynamically indexed array and demotes the whole register-resident row to local memory
__device__ __forceinline__ void add_on_lane(double& reg, double v, int lane, int c) {
  asm("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %2, %3;\n\t@p add.f64 %0, %0, %1;\n\t}" : "+d"(reg) : "d"(v), "r"(lane), "r"(c));
}
// (B' v)[3i + q] without the selection row:  bp v[q] + bv v[3 + q] + s (v[6:9] x arm)[q]
__device__ __forceinline__ void bt_mul(const Leg& g, const double* v, double& g0, double& g1, double& g2) {
  g0 = g.bp * v[0] + g.bv * v[3] + g.s * (v[7] * g.a2 - v[8] * g.a1);
  g1 = g.bp * v[1] + g.bv * v[4] + g.s * (v[8] * g.a0 - v[6] * g.a2);
  g2 = g.bp * v[2] + g.bv * v[5] + g.s * (v[6] * g.a1 - v[7] * g.a0);
}
// symmetric 3 x 3 projector (00, 11, 22, 10, 20, 21) times a vector
__device__ __forceinline__ void proj3(const double* p, double& x, double& y, double& z) {
  const double nx = p[0] * x + p[3] * y + p[4] * z;
  const double ny = p[3] * x + p[1] * y + p[5] * z;
  const double nz = p[4] * x + p[5] * y + p[2] * z;
  x = nx; y = ny; z = nz;
}
// (B u)[r] for state row r < 9 with the stage inputs u (12 slots, zeros on swing legs): row r = 3 grp + q is
//   grp 0: sum_i kp c_i u_i[q];  grp 1: sum_i kv c_i u_i[q];  grp 2: sum_i ks c_i (arm_i x u_i)[q]
// -- per lane a coefficient for the direct term and one for the cross term, and the cyclic successors of q.
struct BRow { int q, q1, q2; double cd, cx; };
__device__ __forceinline__ BRow make_brow(int r, double kp, double kv, double ks) {
  BRow b;
  const int grp = r < 9 ? r / 3 : 0;
  b.q = r < 9 ? r - 3 * grp : 0; b.q1 = b.q == 2 ? 0 : b.q + 1; b.q2 = b.q1 == 2 ? 0 : b.q1 + 1;
  b.cd = grp == 0 ? kp : (grp == 1 ? kv : 0.0); b.cx = grp == 2 ? ks : 0.0;
  return b;
}
__device__ __forceinline__ double b_mul_row(const double* tk, const double* u, const BRow& b) {
  double acc = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {  // swing legs: contact 0 and u = 0
    const double ce = tk[4 * i + 3];
    const double t = b.cd * u[3 * i + b.q] + b.cx * (tk[4 * i + b.q1] * u[3 * i + b.q2] - tk[4 * i + b.q2] * u[3 * i + b.q1]);
    acc = fma(ce > 0.0 ? ce : 0.0, t, acc);
  }
  return acc;
}
__device__ __forceinline__ unsigned stance_mask(const double* tk, int lane) {
  return __ballot_sync(kFull, lane < 4 && tk[4 * (lane & 3) + 3] > 0.0);
}

// Backward sweep: factor the stage systems of  1/2 d'(H + C'SC) d - rhs'd  (mode 1) or of the projected polish
// system (mode 2); the right-hand side in du rides along (y0 = column 21 of Y).  Factors -> slab.
// Returns false (uniformly) on a non-positive pivot.
// Stage factors travel from the slab into a shared-memory double buffer (the stage work area, idle during the vector
// sweeps) with 16-byte asynchronous copies issued one stage ahead: L2 latency stays off the dependent chain and no
// registers are spent on staging.
__device__ __forceinline__ void fac_prefetch(double* buf, const double* fk, int lane) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(buf);
#pragma unroll
  for (int c = 0; c < (kFac / 2 + 31) / 32; ++c) {
    const int e = lane + 32 * c;
    if (e < kFac / 2) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa + 16u * e), "l"(fk + 2 * e) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void fac_wait() {
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncwarp();
}

// Rsg (slab): 6 doubles per leg-step -- 1/2 C'SC entries (xx, yy, zz, zx, zy, -) in mode 1, the projector (00, 11, 22, 10, 20, 21) in mode 2.
__device__ __forceinline__ void rs_prefetch(double* buf, const double* src, int L, int lane) {
  if (lane < 3 * L) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(buf);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa + 16u * lane), "l"(src + 2 * lane) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __noinline__ bool lqr_factor(int N, int L, int gb, int lane, int mode, double* fac, const double* Rsg) {
  const Lay y_ = make_lay(N, L, gb);
  double* const W = smem + y_.W; double* const G = smem + y_.G; double* const Y = smem + y_.Y; double* const Lm = smem + y_.Lm;
  double* const m0 = smem + y_.m0; double* const ps = smem + y_.ps;
  const double* const du = smem + y_.du;
  const double* const hw = smem + kHdrW;
  const double dt = smem[kHdrSc], kp = smem[kHdrSc + 1], kv = smem[kHdrSc + 2];
  const int nf = 3 * L;
  double P[kNZ];
#pragma unroll
  for (int c = 0; c < kNZ; ++c) P[c] = 0.0;
  {
    const double qd = lane < 9 ? (lane == 2 ? smem[kHdrQz + N] : hw[lane]) : 0.0;  // V_N = xi' Q_N xi
#pragma unroll
    for (int c = 0; c < 9; ++c) add_on_lane(P[c], qd, lane, c);
  }
  double pv = 0.0;
  bool ok = true;
  __syncwarp();
  rs_prefetch(smem + y_.rsb + ((N - 1) & 1) * 24, Rsg + (size_t)6 * (N - 1) * L, L, lane);
  fac_wait();
#pragma unroll 1
  for (int k = N - 1; k >= 0; --k) {
    const double* tk = smem + y_.tab + 16 * k;
    const double* Rs = smem + y_.rsb + (k & 1) * 24;  // this stage's block: 6 per leg
    const unsigned mask = stance_mask(tk, lane);
    const double rate = k >= 1 ? 1.0 : 0.0;
    if (k >= 1) rs_prefetch(smem + y_.rsb + ((k - 1) & 1) * 24, Rsg + (size_t)6 * (k - 1) * L, L, lane);
    // ---- T1 = P Bbar, row `lane`
    if (lane < kNZ) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        double w0 = 0.0, w1 = 0.0, w2 = 0.0;
        if ((mask >> i) & 1u) {
          const Leg g = leg_at(tk, i, kp, kv, dt);
          bt_mul(g, P, w0, w1, w2);
          w0 += P[9 + 3 * i]; w1 += P[10 + 3 * i]; w2 += P[11 + 3 * i];
        }
        W[lane * kNF + 3 * i] = w0; W[lane * kNF + 3 * i + 1] = w1; W[lane * kNF + 3 * i + 2] = w2;
      }
      ps[lane] = pv;
    }
    __syncwarp();
    // ---- m0 = Bbar' p - rhs / 2 (slot `lane`), G = Bbar' T1 + R (pairs leg x column)
    if (lane < kNF) {
      const int i = lane / 3, q = lane - 3 * i;
      double mv = 0.0;
      if ((mask >> i) & 1u) {
        double v[9];
#pragma unroll
        for (int s = 0; s < 9; ++s) v[s] = ps[s];
        double g0, g1, g2;
        bt_mul(leg_at(tk, i, kp, kv, dt), v, g0, g1, g2);
        mv = (q == 0 ? g0 : (q == 1 ? g1 : g2)) + ps[9 + lane] - 0.5 * du[k * nf + lane];
      }
      m0[lane] = mv;
    }
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const int e = lane + 32 * pass;
      const int i = e / kNF, c2 = e - kNF * i, i2 = c2 / 3;
      if (e < 4 * kNF && ((mask >> i) & 1u) && ((mask >> i2) & 1u)) {
        double v[9];
#pragma unroll
        for (int s = 0; s < 9; ++s) v[s] = W[s * kNF + c2];
        double g0, g1, g2;
        bt_mul(leg_at(tk, i, kp, kv, dt), v, g0, g1, g2);
        g0 += W[(9 + 3 * i) * kNF + c2]; g1 += W[(10 + 3 * i) * kNF + c2]; g2 += W[(11 + 3 * i) * kNF + c2];
        if (i2 == i) {
          // input Hessian block of the leg: Wf + rate Wr (+ 1/2 C'SC in mode 1)
          const int cq = c2 - 3 * i;
          const double* rs = Rs + 6 * i;
          const double dq = hw[9 + c2] + rate * hw[21 + c2] + (mode == 1 ? rs[cq] : 0.0);
          const double zx = mode == 1 ? rs[3] : 0.0, zy = mode == 1 ? rs[4] : 0.0;
          g0 += cq == 0 ? dq : (cq == 2 ? zx : 0.0);
          g1 += cq == 1 ? dq : (cq == 2 ? zy : 0.0);
          g2 += cq == 0 ? zx : (cq == 1 ? zy : dq);
        }
        G[(3 * i) * kNF + c2] = g0; G[(3 * i + 1) * kNF + c2] = g1; G[(3 * i + 2) * kNF + c2] = g2;
      }
    }
    __syncwarp();
    if (mode == 2) {
      // G <- Pi G Pi + (I - Pi): rows (lane = column), then columns (lane = row)
      if (lane < kNF && ((mask >> (lane / 3)) & 1u)) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (!((mask >> i) & 1u)) continue;
          double x = G[(3 * i) * kNF + lane], yy = G[(3 * i + 1) * kNF + lane], z = G[(3 * i + 2) * kNF + lane];
          proj3(Rs + 6 * i, x, yy, z);
          G[(3 * i) * kNF + lane] = x; G[(3 * i + 1) * kNF + lane] = yy; G[(3 * i + 2) * kNF + lane] = z;
        }
      }
      __syncwarp();
      if (lane < kNF && ((mask >> (lane / 3)) & 1u)) {
        const int il = lane / 3, ql = lane - 3 * il;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (!((mask >> i) & 1u)) continue;
          double x = G[lane * kNF + 3 * i], yy = G[lane * kNF + 3 * i + 1], z = G[lane * kNF + 3 * i + 2];
          const double* pi = Rs + 6 * i;
          proj3(pi, x, yy, z);
          if (i == il) {  // + (I - Pi), row ql
            x += (ql == 0 ? 1.0 : 0.0) - (ql == 0 ? pi[0] : (ql == 1 ? pi[3] : pi[4]));
            yy += (ql == 1 ? 1.0 : 0.0) - (ql == 0 ? pi[3] : (ql == 1 ? pi[1] : pi[5]));
            z += (ql == 2 ? 1.0 : 0.0) - (ql == 0 ? pi[4] : (ql == 1 ? pi[5] : pi[2]));
          }
          G[lane * kNF + 3 * i] = x; G[lane * kNF + 3 * i + 1] = yy; G[lane * kNF + 3 * i + 2] = z;
        }
      }
      __syncwarp();
    }
    // ---- Cholesky of G, row `lane` in registers; columns are broadcast through Lm.  Swing legs are skipped.
    {
      double g[kNF];
      const bool mine = lane < kNF && ((mask >> (lane / 3)) & 1u);
      const int rl = lane < kNF ? lane : 0;
#pragma unroll
      for (int b = 0; b < kNF; b += 2) {
        const double2 t = *reinterpret_cast<const double2*>(G + rl * kNF + b);
        g[b] = mine ? t.x : 0.0; g[b + 1] = mine ? t.y : 0.0;
      }
#pragma unroll
      for (int c = 0; c < kNF; ++c) {
        if (!((mask >> (c / 3)) & 1u)) continue;
        const double d = __shfl_sync(kFull, g[c], c);
        ok = ok && d > 0.0;
        const double inv = fast_rsqrt(d);
        const double l = g[c] * inv;
        if (lane < kNF && lane >= c) Lm[lane * kNF + c] = lane == c ? inv : l;
        __syncwarp();
#pragma unroll
        for (int c2 = c + 1; c2 < kNF; ++c2) {
          if (!((mask >> (c2 / 3)) & 1u)) continue;
          g[c2] = fma(-l, Lm[c2 * kNF + c], g[c2]);
        }
      }
    }
    if (!ok) { fac_wait(); break; }  // uniform: d is a broadcast value
    // ---- Y = L^-1 [M | m0], lane = column (column 21 = m0); M = T1(0:9)' Abar - rate [0, Wr]
    double y[kNF];
    {
      const int col = lane;
      const int cw = col < 9 ? col : 0, cw3 = (col >= 3 && col < 6) ? col - 3 : 0;
      const double f1 = col < 9 ? 1.0 : 0.0, f3 = (col >= 3 && col < 6) ? dt : 0.0;
#pragma unroll
      for (int ia = 0; ia < 4; ++ia) {
        if (!((mask >> ia) & 1u)) { y[3 * ia] = 0.0; y[3 * ia + 1] = 0.0; y[3 * ia + 2] = 0.0; continue; }
        double mf[3];
#pragma unroll
        for (int qa = 0; qa < 3; ++qa) {
          const int c = 3 * ia + qa;
          double v = f1 * W[cw * kNF + c] + f3 * W[cw3 * kNF + c];
          if (col == 9 + c) v = -rate * hw[21 + c];
          if (col == kNZ) v = m0[c];
          mf[qa] = v;
        }
        if (mode == 2) proj3(Rs + 6 * ia, mf[0], mf[1], mf[2]);
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
              if (b < a) acc = fma(-Lm[a * kNF + b], y[b], acc);
            }
          }
          y[a] = acc * Lm[a * kNF + a];
        }
      }
      double* fk = fac + (size_t)k * kFac;
      if (col <= kNZ) {
#pragma unroll
        for (int a = 0; a < kNF; ++a) { Y[a * kYS + col] = y[a]; __stcg(fk + a * kYS + col, y[a]); }
      }
      // L to the slab, one row per lane: strictly lower part with zeros for swing legs, 1 / l_aa (1 for a swing leg) behind it
      if (lane < kNF) {
        const bool ra = (mask >> (lane / 3)) & 1u;
        const double2* row = reinterpret_cast<const double2*>(Lm + lane * kNF);
        double2* dst = reinterpret_cast<double2*>(fk + kOffL + lane * kNF);
        double dinv = 1.0;
#pragma unroll
        for (int c = 0; c < kNF; c += 2) {
          double2 v = row[c >> 1];
          if (c == (lane & ~1)) dinv = (lane & 1) ? v.y : v.x;
          v.x = (ra && c < lane && ((mask >> (c / 3)) & 1u)) ? v.x : 0.0;
          v.y = (ra && c + 1 < lane && ((mask >> ((c + 1) / 3)) & 1u)) ? v.y : 0.0;
          __stcg(dst + (c >> 1), v);
        }
        __stcg(fk + kOffL + kNF * kNF + lane, ra ? dinv : 1.0);
      }
    }
    __syncwarp();
    if (k >= 1) {
      // ---- P <- blkdiag(Q_k, Wr) + Abar' P Abar - Y'Y (row `lane`);  p <- Abar' p - Y' y0
      P[3] = fma(dt, P[0], P[3]); P[4] = fma(dt, P[1], P[4]); P[5] = fma(dt, P[2], P[5]);
      const double rowk = lane < 9 ? 1.0 : 0.0, rowa = (lane >= 3 && lane < 6) ? dt : 0.0;
#pragma unroll
      for (int c = 0; c < 9; ++c) {
        const double up = __shfl_sync(kFull, P[c], (lane + 29) & 31);  // row lane - 3
        P[c] = rowk * fma(rowa, up, P[c]);
      }
#pragma unroll
      for (int c = 9; c < kNZ; ++c) P[c] = 0.0;
      const double dg = lane < 9 ? (lane == 2 ? smem[kHdrQz + k] : hw[lane]) : (lane < kNZ ? hw[21 + lane - 9] : 0.0);
#pragma unroll
      for (int c = 0; c < kNZ; ++c) add_on_lane(P[c], dg, lane, c);
      double pn = 0.0;
      if (lane < 9) pn = ps[lane] + ((lane >= 3 && lane < 6) ? dt * ps[lane - 3] : 0.0);
      const int cl = lane < kNZ ? lane : 0;
#pragma unroll 1
      for (int ia = 0; ia < 4; ++ia) {  // one leg (three rows of Y) per trip: the stance test once per leg
        if (!((mask >> ia) & 1u)) continue;
        const double* Ya = Y + 3 * ia * kYS;
#pragma unroll
        for (int qa = 0; qa < 3; ++qa) {
          const double ya = Ya[qa * kYS + cl];
          const double2* Y2 = reinterpret_cast<const double2*>(Ya + qa * kYS);
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
    fac_wait();
  }
  return ok;
}

// Backward vector sweep for a new right-hand side (du) with the stored factors: y0 of every stage -> slab.
__device__ __noinline__ void lqr_backsolve(int N, int L, int gb, int lane, double* fac) {
  const Lay y_ = make_lay(N, L, gb);
  double* const m0 = smem + y_.m0; double* const ps = smem + y_.ps;
  double* const buf = smem + y_.W;
  const double* const du = smem + y_.du;
  const double dt = smem[kHdrSc], kp = smem[kHdrSc + 1], kv = smem[kHdrSc + 2];
  const int nf = 3 * L;
  double pv = 0.0;
  fac_prefetch(buf + ((N - 1) & 1) * kFac, fac + (size_t)(N - 1) * kFac, lane);
  fac_wait();
#pragma unroll 1
  for (int k = N - 1; k >= 0; --k) {
    const double* tk = smem + y_.tab + 16 * k;
    const double* fb = buf + (k & 1) * kFac;
    const unsigned mask = stance_mask(tk, lane);
    if (k >= 1) fac_prefetch(buf + ((k - 1) & 1) * kFac, fac + (size_t)(k - 1) * kFac, lane);
    if (lane < kNZ) ps[lane] = pv;
    __syncwarp();
    const bool mine = lane < kNF && ((mask >> (lane / 3)) & 1u);
    const int rl = lane < kNF ? lane : 0;
    // row `lane` of the strictly lower L, 1 / l_aa (lanes >= 12 compute on row 0 and are ignored)
    double Lx[kNF];
#pragma unroll
    for (int b = 0; b < kNF; b += 2) {
      const double2 t = *reinterpret_cast<const double2*>(fb + kOffL + rl * kNF + b);
      Lx[b] = t.x; Lx[b + 1] = t.y;
    }
    double inv = fb[kOffD + rl];
    double acc = 0.0;
    if (mine) {
      const int i = lane / 3, q = lane - 3 * i;
      double v[9];
#pragma unroll
      for (int s = 0; s < 9; ++s) v[s] = ps[s];
      double g0, g1, g2;
      bt_mul(leg_at(tk, i, kp, kv, dt), v, g0, g1, g2);
      acc = (q == 0 ? g0 : (q == 1 ? g1 : g2)) + ps[9 + lane] - 0.5 * du[k * nf + lane];
    }
    // y0 = L^-1 m0: column-oriented substitution, one broadcast per column (swing slots carry zeros)
#pragma unroll
    for (int b = 0; b < kNF; ++b) {
      const double yb = __shfl_sync(kFull, acc * inv, b);
      acc = fma(-Lx[b], yb, acc);
    }
    const double y0 = mine ? acc * inv : 0.0;
    if (lane < kNF) { __stcg(fac + (size_t)k * kFac + lane * kYS + kNZ, y0); m0[lane] = y0; }
    __syncwarp();
    if (k >= 1) {
      double pn = 0.0;
      if (lane < 9) pn = ps[lane] + ((lane >= 3 && lane < 6) ? dt * ps[lane - 3] : 0.0);
      const int cl = lane < kNZ ? lane : 0;
#pragma unroll
      for (int a = 0; a < kNF; ++a) {
        if (!((mask >> (a / 3)) & 1u)) continue;
        pn = fma(-fb[a * kYS + cl], m0[a], pn);
      }
      pv = lane < kNZ ? pn : 0.0;
    }
    fac_wait();
  }
}

// Forward sweep: t_k = -L^-T (Y z_k + y0), d_k -> dst (full layout, shared memory), z_{k+1} = [A xi + B d; d].
__device__ __noinline__ void lqr_forward(int N, int L, int gb, int lane, const double* fac, int dst_off) {
  const Lay y_ = make_lay(N, L, gb);
  double* const zs = smem + y_.zs; double* const us = smem + y_.us;
  double* const buf = smem + y_.W;
  double* const dst = smem + dst_off;
  const double dt = smem[kHdrSc], kp = smem[kHdrSc + 1], kv = smem[kHdrSc + 2];
  const int nf = 3 * L;
  double zr = 0.0;
  const BRow br = make_brow(lane, kp, kv, dt);
  __syncwarp();
  fac_prefetch(buf, fac, lane);
  fac_wait();
#pragma unroll 1
  for (int k = 0; k < N; ++k) {
    const double* tk = smem + y_.tab + 16 * k;
    const double* fb = buf + (k & 1) * kFac;
    const unsigned mask = stance_mask(tk, lane);
    if (k + 1 < N) fac_prefetch(buf + ((k + 1) & 1) * kFac, fac + (size_t)(k + 1) * kFac, lane);
    const bool mine = lane < kNF && ((mask >> (lane / 3)) & 1u);
    const int rl = lane < kNF ? lane : 0;
    if (lane < kNZ) zs[lane] = zr;
    __syncwarp();
    // column `lane` of the strictly lower L, 1 / l_aa
    double Lx[kNF];
#pragma unroll
    for (int b = 0; b < kNF; ++b) Lx[b] = fb[kOffL + b * kNF + rl];
    const double inv = fb[kOffD + rl];
    double acc = 0.0;
    {
      const double2* row = reinterpret_cast<const double2*>(fb + rl * kYS);
      const double2* z2 = reinterpret_cast<const double2*>(zs);
#pragma unroll
      for (int c = 0; c < 10; ++c) { const double2 yy = row[c], zz = z2[c]; acc = fma(yy.x, zz.x, acc); acc = fma(yy.y, zz.y, acc); }
      const double2 yl = row[10];
      acc = fma(yl.x, zs[20], acc) + yl.y;
      if (!mine) acc = 0.0;
    }
#pragma unroll
    for (int b = kNF - 1; b >= 0; --b) {
      const double ub = __shfl_sync(kFull, acc * inv, b);
      acc = fma(-Lx[b], ub, acc);
    }
    const double uo = mine ? -(acc * inv) : 0.0;
    if (lane < kNF) us[lane] = uo;
    if (lane < nf) dst[k * nf + lane] = uo;
    __syncwarp();
    double zn = 0.0;
    if (lane < 9) zn = zs[lane] + (lane < 3 ? dt * zs[lane + 3] : 0.0) + b_mul_row(tk, us, br);
    else if (lane < kNZ) zn = us[lane - 9];
    zr = zn;
    fac_wait();
  }
}

// dst = H src + g on the stance entries (0 elsewhere) by one roll-out and one adjoint sweep over the problem
// data (CentroidalMPC.cpp:85-92 dynamics with frozen arms, :203-232 cost) -- independent of the factors.
// X (9N doubles) aliases the stage work area.  src: shared-memory offset; dst: shared or global.
__device__ __noinline__ void stage_gradient(int N, int L, int gb, int lane, int src_off, double* dst) {
  const Lay y_ = make_lay(N, L, gb);
  double* const X = smem + y_.W; double* const zs = smem + y_.zs; double* const us = smem + y_.us; double* const ps = smem + y_.ps;
  const double* const src = smem + src_off;
  const double* const st = smem + y_.st; const double* const hw = smem + kHdrW;
  const double dt = smem[kHdrSc], kp = smem[kHdrSc + 1], kv = smem[kHdrSc + 2];
  const int nf = 3 * L, ns = 9 + 3 * L;
  const double dpz = -kGrav * kp * smem[kHdrSc + 3], dvz = dt * (-kGrav);  // kp m = zeta dt^2
  double xr = lane < 9 ? st[lane] : 0.0;
  const BRow br = make_brow(lane, kp, kv, dt);
#pragma unroll 1
  for (int k = 0; k < N; ++k) {
    const double* tk = smem + y_.tab + 16 * k;
    __syncwarp();
    if (lane < 9) zs[lane] = xr;
    if (lane < kNF) us[lane] = lane < nf ? src[k * nf + lane] : 0.0;
    __syncwarp();
    if (lane < 9) {
      xr = zs[lane] + (lane < 3 ? dt * zs[lane + 3] : 0.0) + b_mul_row(tk, us, br) + (lane == 2 ? dpz : (lane == 5 ? dvz : 0.0));
      X[9 * k + lane] = xr;
    }
  }
  double lam = 0.0;
  const int rg = lane < 9 ? lane / 3 : 0, rq = lane < 9 ? lane - 3 * rg : 0;
#pragma unroll 1
  for (int k = N - 1; k >= 0; --k) {
    const double* tk = smem + y_.tab + 16 * k;
    const unsigned mask = stance_mask(tk, lane);
    __syncwarp();
    if (lane < 9) zs[lane] = lam;
    __syncwarp();
    if (lane < 9) {
      const double qd = lane == 2 ? smem[kHdrQz + k + 1] : hw[lane];
      const double xref = st[ns + rg * 3 * (N + 1) + 3 * (k + 1) + rq];
      lam = 2.0 * qd * (X[9 * k + lane] - xref) + zs[lane] + ((lane >= 3 && lane < 6) ? dt * zs[lane - 3] : 0.0);
      ps[lane] = lam;
    }
    __syncwarp();
    if (lane < nf) {
      const int i = lane / 3, q = lane - 3 * i;
      double gv = 0.0;
      if ((mask >> i) & 1u) {
        const double wf = hw[9 + lane], wr = hw[21 + lane];
        const double f = src[k * nf + lane], fr = (q == 2) ? smem[y_.fz + k] : 0.0;
        double v[9];
#pragma unroll
        for (int s = 0; s < 9; ++s) v[s] = ps[s];
        double g0, g1, g2;
        bt_mul(leg_at(tk, i, kp, kv, dt), v, g0, g1, g2);
        gv = 2.0 * wf * (f - fr) + (q == 0 ? g0 : (q == 1 ? g1 : g2));
        if (k >= 1) gv += 2.0 * wr * (f - src[(k - 1) * nf + lane]);
        if (k + 1 < N) gv -= 2.0 * wr * (src[(k + 1) * nf + lane] - f);
      }
      dst[k * nf + lane] = gv;
    }
  }
  __syncwarp();
}

// CTA header + per-group staging shared by the solve and the probe kernel.  Returns finite (uniform).
__device__ __forceinline__ void fill_header(const DevConfig& cfg) {
  const int N = cfg.N, L = cfg.L;
  for (int e = threadIdx.x; e <= N; e += blockDim.x) {
    const double om = (cfg.w[2] * 0.5) * exp(-(double)e) + cfg.w[2] * 0.5;  // CentroidalMPC.cpp:205
    smem[kHdrQz + e] = om * om;
  }
  for (int e = threadIdx.x; e < 33; e += blockDim.x) {
    double v = 0.0;
    if (e < 9) v = cfg.w[e];
    else if (e < 21) v = (e - 9 < 3 * L) ? cfg.w[9 + 3 * L + e - 9] : 0.0;
    else v = (e - 21 < 3 * L) ? cfg.w[9 + 6 * L + e - 21] : 0.0;
    smem[kHdrW + e] = v;
  }
  if (threadIdx.x == 0) {
    smem[kHdrSc] = cfg.dt; smem[kHdrSc + 1] = (cfg.zoh ? 0.5 : 0.0) * cfg.dt * cfg.dt / cfg.mass;
    smem[kHdrSc + 2] = cfg.dt / cfg.mass; smem[kHdrSc + 3] = cfg.mass;
  }
  __syncthreads();
}
// Inputs -> shared memory (coalesced, L2-only loads): state | des_state stay resident, des_inputs is turned into the
// per-stage table (lever arm, contact) and the desired fz.  invalid = a step without stance leg (:328-330).
__device__ __forceinline__ bool stage_instance(const DevConfig& cfg, const SolveArgs& args, const Lay& y_, int inst, int lane, bool& invalid, int& nb) {
  const int N = cfg.N, L = cfg.L;
  const int ns = 9 + 3 * L, nds = 9 * (N + 1), ndi = L * (4 * N + 3);
  double* const st = smem + y_.st; double* const di = smem + y_.u;  // des_inputs: staged in the (still dead) vector region
  bool finite = true;
  {
    const double* src = args.state + (size_t)inst * ns;
    for (int t = lane; t < ns; t += 32) { const double v = __ldcg(src + t); st[t] = v; finite = finite && isfinite(v); }
    src = args.des_state + (size_t)inst * nds;
    for (int t = lane; t < nds; t += 32) { const double v = __ldcg(src + t); st[ns + t] = v; finite = finite && isfinite(v); }
    src = args.des_inputs + (size_t)inst * ndi;
    for (int t = lane; t < ndi; t += 32) { const double v = __ldcg(src + t); di[t] = v; finite = finite && isfinite(v); }
  }
  __syncwarp();
  finite = __all_sync(kFull, finite);
  double* const tab = smem + y_.tab;
  for (int e = lane; e < 4 * N; e += 32) {
    const int k = e >> 2, i = e & 3;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, ce = 0.0;
    if (i < L) {
      ce = di[i * (4 * N + 3) + k];
      const double* foot = di + i * (4 * N + 3) + N + 3 * k;
      const double* com = st + ns + 3 * k;
      a0 = foot[0] - com[0]; a1 = foot[1] - com[1]; a2 = foot[2] - com[2];  // frozen lever arm
    }
    tab[4 * e] = a0; tab[4 * e + 1] = a1; tab[4 * e + 2] = a2; tab[4 * e + 3] = ce;
  }
  nb = 0;
  double colsum = 0.0;
  if (lane < N)
    for (int i = 0; i < L; ++i) { const double ce = di[i * (4 * N + 3) + lane]; colsum += ce; nb += ce > 0.0 ? 1 : 0; }
  if (lane < N) smem[y_.fz + lane] = cfg.mass * kGrav / colsum;
  invalid = __any_sync(kFull, lane < N && !(colsum > 0.0));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nb += __shfl_xor_sync(kFull, nb, o);
  __syncwarp();
  return finite;
}

}  // namespace

__global__ void __launch_bounds__(384, 1) cmpc_ripm_kernel(const __grid_constant__ DevConfig cfg, const __grid_constant__ SolveArgs args) {
  pdl_prologue(args.pdl_trigger > 1);
  const int N = cfg.N, L = cfg.L;
  const int nf = 3 * L, nfN = nf * N, nbfull = L * N, mfull = 5 * nbfull;
  const int lane = threadIdx.x & 31, gid = threadIdx.x >> 5;
  const Lay y0_ = make_lay(N, L, 0);
  const int gb = kHdr + gid * y0_.total;
  const Lay y_ = make_lay(N, L, gb);
  double* const slab = args.scratch + (size_t)(blockIdx.x * args.groups + gid) * args.scratch_per_group;
  double* const fac = slab;
  double* const g_zl = slab + (size_t)N * kFac;
  double* const g_zu = g_zl + mfull;
  double* const g_dua = g_zu + mfull;                       // copy of the affine direction; the polish's gradients
  double* const g_Rs = g_dua + ((nfN + 1) & ~1);            // input-Hessian data per leg-step (lqr_factor stages it)
  double* const s_u = smem + y_.u; double* const s_du = smem + y_.du; double* const s_rd = smem + y_.rd;
  const double* const s_tab = smem + y_.tab;
  uint16_t* const s_act = reinterpret_cast<uint16_t*>(smem + y_.act);
  const double mass = cfg.mass;
  // work lists: one (perm / count / work) or several drained in order (SolveArgs::nlists)
  const int nl = args.nlists > 0 ? args.nlists : 1;
  int total = 0;
  for (int q = 0; q < nl; ++q) {
    const int c = args.nlists > 0 ? *args.lcount[q] : (args.count ? *args.count : args.count_imm);
    total += c > 0 ? c : 0;
  }
  // launch planning: the host reads this word before the next call (written only when it changes)
  if (args.hint_out && blockIdx.x == 0 && threadIdx.x == 0 && *args.hint_shadow != total) { *args.hint_shadow = total; *args.hint_out = total; }
  if (total <= 0) return;  // nothing to do (uniform over the grid)
  fill_header(cfg);
  int li = 0;

  while (true) {
    const int count = args.nlists > 0 ? *args.lcount[li] : (args.count ? *args.count : args.count_imm);
    int slot = 0;
    if (lane == 0) slot = atomicAdd(args.nlists > 0 ? args.lwork[li] : args.work, 1);
    slot = __shfl_sync(kFull, slot, 0);
    if (slot >= count) {
      if (++li >= nl) break;
      continue;
    }
    const int32_t* perm = args.nlists > 0 ? args.lperm[li] : args.perm;
    const int inst = perm ? perm[slot] : slot;
    bool invalid = false;
    int nb = 0;
    const bool finite = stage_instance(cfg, args, y_, inst, lane, invalid, nb);
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
    for (int t = lane; t < nfN; t += 32) { g_dua[t] = 0.0; s_du[t] = 0.0; }
    __syncwarp();
    stage_gradient(N, L, gb, lane, y_.du, s_rd);
    double gmax = 0.0;
    for (int t = lane; t < nfN; t += 32) gmax = fmax(gmax, fabs(s_rd[t]));
#pragma unroll 1
    for (int tb = lane; tb < nbfull; tb += 32) {
      const int k = div_legs(tb, L), i = tb - k * L;
      const double ce = s_tab[16 * k + 4 * i + 3];
      double fz = 0.0;
      if (ce > 0.0) {
        fz = smem[y_.fz + k];
        fz = fmin(fz, 0.5 * mass * kGrav * (double)L * ce);
        fz = fmin(fz, 0.5 * kFricUb * ce / cfg.mu[i]);
      }
      s_u[3 * tb] = 0.0; s_u[3 * tb + 1] = 0.0; s_u[3 * tb + 2] = fz;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) gmax = fmax(gmax, __shfl_xor_sync(kFull, gmax, o));
    __syncwarp();
    const double gs = 1.0 + gmax;
    stage_gradient(N, L, gb, lane, y_.u, s_rd);  // H u0 + g
    double r0max = 0.0;
    for (int t = lane; t < nfN; t += 32) r0max = fmax(r0max, fabs(s_rd[t]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r0max = fmax(r0max, __shfl_xor_sync(kFull, r0max, o));
    const double mu0 = fmax(1e-2, r0max);
#pragma unroll 1
    for (int tb = lane; tb < nbfull; tb += 32) {
      const int k = div_legs(tb, L), i = tb - k * L;
      const double ce = s_tab[16 * k + 4 * i + 3];
      double y[5] = {1.0, 1.0, 1.0, 1.0, 1.0};
      const double ubxy = kFricUb * ce, ubz = mass * kGrav * (double)L * ce;  // :183,199
      if (ce > 0.0) cmul5(cfg.mu[i], s_u + 3 * tb, y);
      for (int q = 0; q < 5; ++q) {
        const double ub = q < 4 ? ubxy : ubz;
        g_zl[5 * tb + q] = ce > 0.0 ? mu0 / y[q] : 0.0;
        g_zu[5 * tb + q] = ce > 0.0 ? mu0 / (ub - y[q]) : 0.0;
      }
    }
    __syncwarp();

    int status = CMPC_STATUS_MAX_ITER, it = 0, npolish = 0;
    bool numerical = false, ipm_ok = false;
    bool grad_fresh = true, grad_exact = true;  // rd holds H u + g of the current u / it comes from a sweep, not from an update
    double us = 1.0;
#pragma unroll 1
    for (it = 0; it <= cfg.max_iter; ++it) {
      if (!grad_fresh) { stage_gradient(N, L, gb, lane, y_.u, s_rd); grad_exact = true; }
      grad_fresh = false;
      double rmax = 0.0, umax = 0.0, gap = 0.0, mu = 0.0;
      bool conv_mu = false;
      // H u + g is carried from iteration to iteration by the update  H (u + a du) + g = (H u + g) + a H du  with
      // H du = rhs - C'SC du read off the Newton system (no sweep).  Every decision that ends the iteration (polish,
      // termination) is taken on a residual from a fresh roll-out + adjoint sweep: once the gap is converged the
      // gradient is recomputed and the test repeated.
#pragma unroll 1
      for (int attempt = 0; attempt < 2; ++attempt) {
      rmax = 0.0; umax = 0.0; gap = 0.0;
#pragma unroll 1
      for (int tb = lane; tb < nbfull; tb += 32) {
        const int k = div_legs(tb, L), i = tb - k * L;
        const double ce = s_tab[16 * k + 4 * i + 3];
        if (!(ce > 0.0)) continue;
        const double ubxy = kFricUb * ce, ubz = mass * kGrav * (double)L * ce;
        double w[5], o[3], ys[5];
        cmul5(cfg.mu[i], s_u + 3 * tb, ys);
        for (int q = 0; q < 5; ++q) {
          const int t = 5 * tb + q;
          const double sl = ys[q], su = (q < 4 ? ubxy : ubz) - sl, zl = g_zl[t], zu = g_zu[t];
          w[q] = zl - zu;
          gap += sl * zl + su * zu;
        }
        ctmul5(cfg.mu[i], w, o);
        for (int q = 0; q < 3; ++q) {
          const double rr = s_rd[3 * tb + q] - o[q];
          s_rd[3 * tb + q] = rr;
          rmax = fmax(rmax, fabs(rr)); umax = fmax(umax, fabs(s_u[3 * tb + q]));
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        rmax = fmax(rmax, __shfl_xor_sync(kFull, rmax, o));
        umax = fmax(umax, __shfl_xor_sync(kFull, umax, o));
        gap += __shfl_xor_sync(kFull, gap, o);
      }
      __syncwarp();
      mu = gap / (2.0 * (double)m);
      us = 1.0 + umax;
      conv_mu = mu <= cfg.tol * gs * us;
      if (grad_exact || !conv_mu) break;
      stage_gradient(N, L, gb, lane, y_.u, s_rd);
      grad_exact = true;
      }
      const bool strict = conv_mu && rmax <= cfg.tol * gs;
      const bool ready = conv_mu && rmax <= 1e4 * cfg.tol * gs;
      ipm_ok = conv_mu && rmax <= 10.0 * cfg.tol * gs;
      if (cfg.polish && ready && npolish < 3) {
        ++npolish;
        // ---- active-set guess from the interior-point iterate
#pragma unroll 1
        for (int tb = lane; tb < nbfull; tb += 32) {
          const int k = div_legs(tb, L), i = tb - k * L;
          const double ce = s_tab[16 * k + 4 * i + 3];
          uint16_t a = 0;
          if (ce > 0.0) {
            const double ubxy = kFricUb * ce, ubz = mass * kGrav * (double)L * ce;
            double ys[5];
            cmul5(cfg.mu[i], s_u + 3 * tb, ys);
            for (int q = 0; q < 5; ++q) {
              const int t = 5 * tb + q;
              const double sl = ys[q], su = (q < 4 ? ubxy : ubz) - sl;
              if (kGuess * g_zl[t] * us > sl * gs) a |= (uint16_t)(1u << q);
              if (kGuess * g_zu[t] * us > su * gs) a |= (uint16_t)(1u << (5 + q));
            }
          }
          s_act[tb] = a;
        }
        __syncwarp();
        bool accepted = false;
#pragma unroll 1
        for (int pass = 0; pass < 6 && !accepted; ++pass) {
          // per leg-step: particular solution f0 (-> rd) and the projector onto the null space of the working rows (-> Rs)
          bool ok_all = true;
#pragma unroll 1
          for (int tb = lane; tb < nbfull; tb += 32) {
            const int k = div_legs(tb, L), i = tb - k * L;
            const double ce = s_tab[16 * k + 4 * i + 3];
            double f0[3] = {0.0, 0.0, 0.0};
            if (ce > 0.0) {
              const double mub = cfg.mu[i];
              const double ubxy = kFricUb * ce, ubz = mass * kGrav * (double)L * ce;
              const unsigned a = s_act[tb];
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
              double* rs = g_Rs + 6 * tb;
              rs[0] = p00; rs[1] = p11; rs[2] = p22; rs[3] = p10; rs[4] = p20; rs[5] = p21;
            }
            s_rd[3 * tb] = f0[0]; s_rd[3 * tb + 1] = f0[1]; s_rd[3 * tb + 2] = f0[2];
          }
          __syncwarp();
          ok_all = __all_sync(kFull, ok_all);
          if (!ok_all) break;
          stage_gradient(N, L, gb, lane, y_.rd, g_dua);  // H f0 + g
          for (int t = lane; t < nfN; t += 32) s_du[t] = -g_dua[t];
          __syncwarp();
          if (args.phase_lock) __syncthreads_and(0);
          if (!lqr_factor(N, L, gb, lane, 2, fac, g_Rs)) break;
          lqr_forward(N, L, gb, lane, fac, y_.du);
          for (int t = lane; t < nfN; t += 32) s_du[t] += s_rd[t];  // candidate point up = f0 + Z t
          __syncwarp();
          stage_gradient(N, L, gb, lane, y_.du, g_dua);  // H up + g
          // multipliers, verification, correction; when the pass verifies the loop runs once more to commit
          bool good = false;
#pragma unroll 1
          for (int commit = 0; commit < 2; ++commit) {
            bool okm = true, changed = false;
#pragma unroll 1
            for (int tb = lane; tb < nbfull; tb += 32) {
              const int k = div_legs(tb, L), i = tb - k * L;
              const double ce = s_tab[16 * k + 4 * i + 3];
              if (!(ce > 0.0)) continue;
              const double mub = cfg.mu[i];
              const double ubxy = kFricUb * ce, ubz = mass * kGrav * (double)L * ce;
              const unsigned a = s_act[tb];
              double rb[3], y[5], ll[5] = {0, 0, 0, 0, 0}, lu[5] = {0, 0, 0, 0, 0};
              for (int q = 0; q < 3; ++q) rb[q] = g_dua[3 * tb + q];
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
                for (int q = 0; q < 5; ++q) { g_zl[5 * tb + q] = ll[q]; g_zu[5 * tb + q] = lu[q]; }
                continue;
              }
              cmul5(mub, s_du + 3 * tb, y);
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
              s_act[tb] = (uint16_t)an;
            }
            __syncwarp();
            if (commit) break;
            good = __all_sync(kFull, okm && !changed);
            if (!good) break;
          }
          if (good) accepted = true;
        }
        if (accepted) {
          for (int t = lane; t < nfN; t += 32) { s_u[t] = s_du[t]; s_rd[t] = g_dua[t]; }  // rd = H u + g at the KKT point
          __syncwarp();
          status = CMPC_STATUS_OK;
          break;
        }
        __syncwarp();
        // polish not accepted: rd was the f0 scratch -> recompute the dual residual
        stage_gradient(N, L, gb, lane, y_.u, s_rd);
#pragma unroll 1
        for (int tb = lane; tb < nbfull; tb += 32) {
          const int k = div_legs(tb, L), i = tb - k * L;
          if (!(s_tab[16 * k + 4 * i + 3] > 0.0)) continue;
          double w[5], o[3];
          for (int q = 0; q < 5; ++q) w[q] = g_zl[5 * tb + q] - g_zu[5 * tb + q];
          ctmul5(cfg.mu[i], w, o);
          for (int q = 0; q < 3; ++q) s_rd[3 * tb + q] -= o[q];
        }
        __syncwarp();
      }
      if (strict && (!cfg.polish || npolish >= 3)) break;
      if (mu <= 1e-8 * cfg.tol * gs * us) break;  // far past convergence: stop before 0/0
      if (it == cfg.max_iter) break;

      // ---- 1/2 C' diag(zl/sl + zu/su) C per leg-step -> Rs;  affine right-hand side -rd + C'(zu - zl) -> du
#pragma unroll 1
      for (int tb = lane; tb < nbfull; tb += 32) {
        const int k = div_legs(tb, L), i = tb - k * L;
        const double ce = s_tab[16 * k + 4 * i + 3];
        if (!(ce > 0.0)) { s_du[3 * tb] = 0.0; s_du[3 * tb + 1] = 0.0; s_du[3 * tb + 2] = 0.0; continue; }
        const double ubxy = kFricUb * ce, ubz = mass * kGrav * (double)L * ce;
        const double mb = cfg.mu[i];
        double sg[5], tq[5], o[3], ys[5];
        cmul5(mb, s_u + 3 * tb, ys);
        for (int q = 0; q < 5; ++q) {
          const int t = 5 * tb + q;
          const double sl = ys[q], su = (q < 4 ? ubxy : ubz) - sl, zl = g_zl[t], zu = g_zu[t];
          sg[q] = zl * fast_rcp(sl) + zu * fast_rcp(su);
          tq[q] = zu - zl;
        }
        const double sx = sg[0] + sg[1], sy = sg[2] + sg[3];
        double* rs = g_Rs + 6 * tb;
        rs[0] = 0.5 * sx; rs[1] = 0.5 * sy; rs[2] = 0.5 * (mb * mb * (sx + sy) + sg[4]);
        rs[3] = 0.5 * mb * (sg[1] - sg[0]); rs[4] = 0.5 * mb * (sg[3] - sg[2]);
        ctmul5(mb, tq, o);
        for (int q = 0; q < 3; ++q) s_du[3 * tb + q] = -s_rd[3 * tb + q] + o[q];
      }
      __syncwarp();
      if (args.phase_lock) __syncthreads_and(0);
      if (!lqr_factor(N, L, gb, lane, 1, fac, g_Rs)) { numerical = true; break; }

      double tmax = 0.0, sigma = 0.0;
#pragma unroll 1
      for (int phase = 0; phase < 2; ++phase) {
        // phase 0: affine predictor (its right-hand side rode along the factor sweep); phase 1: centred corrector (Mehrotra)
        if (phase) {
          for (int t = lane; t < nfN; t += 32) g_dua[t] = s_du[t];
          __syncwarp();
#pragma unroll 1
          for (int tb = lane; tb < nbfull; tb += 32) {
            const int k = div_legs(tb, L), i = tb - k * L;
            const double ce = s_tab[16 * k + 4 * i + 3];
            if (!(ce > 0.0)) continue;  // du = dua = 0 there
            const double mub = cfg.mu[i];
            const double ubxy = kFricUb * ce, ubz = mass * kGrav * (double)L * ce;
            double tq[5], o[3], ys[5], ya[5];
            cmul5(mub, s_u + 3 * tb, ys);
            cmul5(mub, g_dua + 3 * tb, ya);
            for (int q = 0; q < 5; ++q) {
              const int t = 5 * tb + q;
              const double sl = ys[q], su = (q < 4 ? ubxy : ubz) - sl, zl = g_zl[t], zu = g_zu[t];
              const double isl = fast_rcp(sl), isu = fast_rcp(su);
              const double dla = (-sl * zl - zl * ya[q]) * isl, dua_ = (-su * zu + zu * ya[q]) * isu;
              const double rcl = -sl * zl + sigma * mu - ya[q] * dla;
              const double rcu = -su * zu + sigma * mu + ya[q] * dua_;
              tq[q] = rcl * isl - rcu * isu;
            }
            ctmul5(mub, tq, o);
            for (int q = 0; q < 3; ++q) s_du[3 * tb + q] = -s_rd[3 * tb + q] + o[q];
          }
          __syncwarp();
          lqr_backsolve(N, L, gb, lane, fac);
        }
        lqr_forward(N, L, gb, lane, fac, y_.du);
        // step to the boundary: alpha_max = 1 / max_i(-ds_i/s_i, -dz_i/z_i)
        double tloc = 0.0;
#pragma unroll 1
        for (int tb = lane; tb < nbfull; tb += 32) {
          const int k = div_legs(tb, L), i = tb - k * L;
          const double ce = s_tab[16 * k + 4 * i + 3];
          if (!(ce > 0.0)) continue;
          const double mub = cfg.mu[i];
          const double ubxy = kFricUb * ce, ubz = mass * kGrav * (double)L * ce;
          double ys[5], yd[5], ya[5] = {0, 0, 0, 0, 0};
          cmul5(mub, s_u + 3 * tb, ys);
          cmul5(mub, s_du + 3 * tb, yd);
          if (phase) cmul5(mub, g_dua + 3 * tb, ya);
          for (int q = 0; q < 5; ++q) {
            const int t = 5 * tb + q;
            const double sl = ys[q], su = (q < 4 ? ubxy : ubz) - sl, zl = g_zl[t], zu = g_zu[t];
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
            const int k = div_legs(tb, L), i = tb - k * L;
            const double ce = s_tab[16 * k + 4 * i + 3];
            if (!(ce > 0.0)) continue;
            const double mub = cfg.mu[i];
            const double ubxy = kFricUb * ce, ubz = mass * kGrav * (double)L * ce;
            double ys[5], yd[5];
            cmul5(mub, s_u + 3 * tb, ys);
            cmul5(mub, s_du + 3 * tb, yd);
            for (int q = 0; q < 5; ++q) {
              const int t = 5 * tb + q;
              const double sl = ys[q], su = (q < 4 ? ubxy : ubz) - sl, zl = g_zl[t], zu = g_zu[t];
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
        const int k = div_legs(tb, L), i = tb - k * L;
        const double ce = s_tab[16 * k + 4 * i + 3];
        if (!(ce > 0.0)) continue;
        const double mub = cfg.mu[i];
        const double ubxy = kFricUb * ce, ubz = mass * kGrav * (double)L * ce;
        double ys[5], yd[5], ya[5];
        cmul5(mub, s_u + 3 * tb, ys);  // slacks at the current point (before the update)
        cmul5(mub, s_du + 3 * tb, yd);
        cmul5(mub, g_dua + 3 * tb, ya);
        double wo[5], hq[5], o1[3], o2[3];
        for (int q = 0; q < 5; ++q) {
          const int t = 5 * tb + q;
          const double sl = ys[q], su = (q < 4 ? ubxy : ubz) - sl, zl = g_zl[t], zu = g_zu[t];
          const double isl = fast_rcp(sl), isu = fast_rcp(su);
          const double dla = (-sl * zl - zl * ya[q]) * isl, dua_ = (-su * zu + zu * ya[q]) * isu;
          const double rcl = -sl * zl + sigma * mu - ya[q] * dla;
          const double rcu = -su * zu + sigma * mu + ya[q] * dua_;
          wo[q] = zl - zu;                                                   // C'(zl - zu): dual residual -> gradient
          hq[q] = (rcl * isl - rcu * isu) - (zl * isl + zu * isu) * yd[q];   // corrector right-hand side rows minus S C du
          g_zl[t] = zl + alpha * (rcl - zl * yd[q]) * isl;
          g_zu[t] = zu + alpha * (rcu + zu * yd[q]) * isu;
        }
        ctmul5(mub, wo, o1);
        ctmul5(mub, hq, o2);
        for (int q = 0; q < 3; ++q) {
          const double v = s_u[3 * tb + q] + alpha * s_du[3 * tb + q];
          s_u[3 * tb + q] = v; fin = fin && isfinite(v);
          const double rdv = s_rd[3 * tb + q];                              // (H + C'SC) du = -rd + C' rows  =>  H du = -rd + o2
          s_rd[3 * tb + q] = rdv + o1[q] + alpha * (o2[q] - rdv);            // H u_new + g
        }
      }
      __syncwarp();
      fin = __all_sync(kFull, fin);
      if (!fin) { numerical = true; break; }
      grad_fresh = true; grad_exact = false;
    }
    if (numerical) status = CMPC_STATUS_NUMERICAL;
    else if (status != CMPC_STATUS_OK) status = ipm_ok ? CMPC_STATUS_OK_IPM : CMPC_STATUS_MAX_ITER;

    // ---- outputs
    if (!numerical) {
      if (status != CMPC_STATUS_OK) stage_gradient(N, L, gb, lane, y_.u, s_rd);  // an accepted polish left H u + g in rd
      double stat = 0.0, umax = 0.0, prim = 0.0, dual = 0.0, comp = 0.0;
#pragma unroll 1
      for (int tb = lane; tb < nbfull; tb += 32) {
        const int k = div_legs(tb, L), i = tb - k * L;
        const double ce = s_tab[16 * k + 4 * i + 3];
        if (!(ce > 0.0)) continue;
        const double mub = cfg.mu[i];
        const double ubxy = kFricUb * ce, ubz = mass * kGrav * (double)L * ce;
        double w[5], o[3], y[5];
        for (int q = 0; q < 5; ++q) w[q] = g_zl[5 * tb + q] - g_zu[5 * tb + q];
        ctmul5(mub, w, o);
        cmul5(mub, s_u + 3 * tb, y);
        for (int q = 0; q < 3; ++q) {
          stat = fmax(stat, fabs(s_rd[3 * tb + q] - o[q]));
          umax = fmax(umax, fabs(s_u[3 * tb + q]));
        }
        for (int q = 0; q < 5; ++q) {
          const double ub = q < 4 ? ubxy : ubz;
          const double sl = y[q], su = ub - y[q], zl = g_zl[5 * tb + q], zu = g_zu[5 * tb + q];
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
        const int k = div_legs(tb, L), i = tb - k * L;
        const double ce = s_tab[16 * k + 4 * i + 3];
        uint16_t a = 0x8000;
        if (ce > 0.0) {
          const double ubxy = kFricUb * ce, ubz = mass * kGrav * (double)L * ce;
          double ys[5];
          cmul5(cfg.mu[i], s_u + 3 * tb, ys);
          a = 0;
          for (int q = 0; q < 5; ++q) {
            const int t = 5 * tb + q;
            const double sl = ys[q], su = (q < 4 ? ubxy : ubz) - sl;
            bool al, au;
            if (status == CMPC_STATUS_OK) { al = sl <= 1e-9 * usf; au = su <= 1e-9 * usf; }
            else { al = g_zl[t] * usf > sl * gs; au = g_zu[t] * usf > su * gs; }
            a |= (uint16_t)((al ? 1 : 0) << q | (au ? 1 : 0) << (5 + q));
          }
        }
        s_act[tb] = a;
      }
      __syncwarp();
      // forces in the reference's per-leg order [L][N][3] (CentroidalMPC.cpp:270)
      for (int t = lane; t < nfN; t += 32) {
        const int i = t / (3 * N), rem = t - i * 3 * N, j = rem / 3, q = rem - 3 * j;
        args.forces[(size_t)inst * nfN + t] = s_u[j * nf + 3 * i + q];
      }
      if (args.lam) {
        for (int t = lane; t < mfull; t += 32) {
          args.lam[(size_t)inst * 2 * mfull + t] = g_zl[t];
          args.lam[(size_t)inst * 2 * mfull + mfull + t] = g_zu[t];
        }
      }
      if (args.active) for (int t = lane; t < nbfull; t += 32) args.active[(size_t)inst * nbfull + t] = s_act[t];
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
  // phase lock: the warps of the CTA enter the factor sweep (half of all instructions) together, so that they share its
  // instruction fetches; a warp that has run out of work keeps the others' barriers company until everybody is done
  if (args.phase_lock) while (!__syncthreads_and(1)) {}
}

// Diagnostic / parity entry (cmpc_stage_step_batch): the stage-wise linear algebra alone.  For every instance:
// d_fused = the solution of the stage system with the right-hand side riding along the factor sweep,
// d_resolve = the same right-hand side through the stored factors (corrector path), grad = H rhs + g
// (roll-out + adjoint sweep with rhs taken as a force vector).  hess: 6 doubles per leg-step in the Rs layout
// (mode 1: 1/2 C'SC entries xx, yy, zz, zx, zy, -; mode 2: projector 00, 11, 22, 10, 20, 21).
__global__ void __launch_bounds__(384, 1) cmpc_ripm_probe_kernel(const __grid_constant__ DevConfig cfg, const __grid_constant__ SolveArgs args,
                                                                 const double* hess, const double* rhs, int mode, double* d_fused,
                                                                 double* d_resolve, double* grad, int B) {
  const int N = cfg.N, L = cfg.L;
  const int nf = 3 * L, nfN = nf * N, nbfull = L * N;
  const int lane = threadIdx.x & 31, gid = threadIdx.x >> 5;
  const Lay y0_ = make_lay(N, L, 0);
  const int gb = kHdr + gid * y0_.total;
  const Lay y_ = make_lay(N, L, gb);
  double* const fac = args.scratch + (size_t)(blockIdx.x * args.groups + gid) * args.scratch_per_group;
  double* const s_u = smem + y_.u; double* const s_du = smem + y_.du; double* const s_rd = smem + y_.rd;
  double* const g_Rs = fac + (size_t)N * kFac;  // (the solve kernel keeps its multipliers here)
  fill_header(cfg);
  for (int inst = blockIdx.x * args.groups + gid; inst < B; inst += gridDim.x * args.groups) {
    bool invalid; int nb;
    stage_instance(cfg, args, y_, inst, lane, invalid, nb);
    for (int t = lane; t < 6 * nbfull; t += 32) g_Rs[t] = hess[(size_t)inst * 6 * nbfull + t];
    for (int t = lane; t < nfN; t += 32) { s_du[t] = rhs[(size_t)inst * nfN + t]; s_u[t] = s_du[t]; }
    __syncwarp();
    const bool ok = lqr_factor(N, L, gb, lane, mode, fac, g_Rs);
    lqr_forward(N, L, gb, lane, fac, y_.du);
    for (int t = lane; t < nfN; t += 32) { d_fused[(size_t)inst * nfN + t] = ok ? s_du[t] : nan(""); s_du[t] = s_u[t]; }
    __syncwarp();
    if (ok) {
      lqr_backsolve(N, L, gb, lane, fac);
      lqr_forward(N, L, gb, lane, fac, y_.du);
    }
    for (int t = lane; t < nfN; t += 32) { d_resolve[(size_t)inst * nfN + t] = ok ? s_du[t] : nan(""); }
    __syncwarp();
    stage_gradient(N, L, gb, lane, y_.u, s_rd);
    for (int t = lane; t < nfN; t += 32) grad[(size_t)inst * nfN + t] = s_rd[t];
    __syncwarp();
  }
}

cudaError_t launch_ripm_probe(int grid, int block, size_t smem_bytes, cudaStream_t stream, const DevConfig& cfg, const SolveArgs& args,
                              const double* hess, const double* rhs, int mode, double* d_fused, double* d_resolve, double* grad, int B) {
  cudaError_t e = cudaFuncSetAttribute(cmpc_ripm_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  if (e != cudaSuccess) return e;
  cmpc_ripm_probe_kernel<<<grid, block, smem_bytes, stream>>>(cfg, args, hess, rhs, mode, d_fused, d_resolve, grad, B);
  return cudaGetLastError();
}

void ripm_sizes(int N, int L, int* group_doubles, int* cta_doubles, int* slab_doubles) {
  const Lay y = make_lay(N, L, 0);
  *group_doubles = y.total;
  *cta_doubles = kHdr;
  *slab_doubles = N * kFac + 2 * 5 * L * N + ((3 * L * N + 1) & ~1) + 6 * L * N;
}

cudaError_t launch_ripm_kernel(int grid, int block, size_t smem_bytes, cudaStream_t stream, const DevConfig& cfg, const SolveArgs& args) {
  return launch_ex(cmpc_ripm_kernel, grid, block, smem_bytes, stream, args.pdl != 0, cfg, args);
}

cudaError_t set_ripm_kernel_smem(size_t bytes) {
  return cudaFuncSetAttribute(cmpc_ripm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

}  // namespace cmpc
