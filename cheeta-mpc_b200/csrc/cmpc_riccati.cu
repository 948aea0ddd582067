// cmpc_riccati.cu -- stage-wise (Riccati) presolve kernel (sm_100a) and its launcher.  SURVEY §8 f3.
//
// The same unconstrained minimiser the dense presolve kernel computes, but from the un-condensed
// optimal-control form of the QP -- the stage-wise (A, B, b, Q, R, q, r) view of the reference's
// HPIPM adapter (ocs2_sqp/hpipm_catkin/src/HpipmInterface.cpp:166-301) -- instead of the condensed
// 3LN x 3LN Hessian:
//   x_{j+1} = A x_j + Bf_j F_j + d,                        F_j in R^{3L}, swing-leg entries pinned to 0
//   cost    = sum_k (x_k - xr_k)' Q_k (x_k - xr_k) + sum_j (F_j - Fr_j)' Wf (F_j - Fr_j)
//             + sum_j (F_{j+1} - F_j)' Wr (F_{j+1} - F_j)                    (CentroidalMPC.cpp:203-232)
// The force-rate term couples consecutive inputs, so the sweep runs on the augmented state
// z_k = [x_k; F_{k-1}] (9 + 3L = 21 entries): a backward Riccati recursion with one m x m Cholesky per
// stage (m <= 12 free inputs), gains parked in an L2 slab, then a forward roll-out.  Cost
// O(N (9 + 3L)^3) instead of O((3LN)^3): at horizon 30 / stand (n = 360) 1 MFLOP instead of 15.5.
// Verification is independent of the recursion: stationarity H U + g is evaluated by a roll-out and
// an adjoint (costate) sweep over the problem data, the constraint rows as in the dense kernels.
// Verified -> outputs (status OK, iters 0); anything else is deferred to the interior-point kernel.
// One warp per instance.
#include "cmpc_device.cuh"

namespace cmpc {

namespace {

constexpr int kMu = 3 * kMaxLegs;  // most free inputs of one stage
constexpr int kGld = 13;           // leading dimension of the stage Hessian G
constexpr int kRicCta = 10 + 58 + 34;  // doubles of CTA-shared tables (78 bytes + 231 uint16 decode entries + z weights of nodes 0..32)

struct RicView {
  const double* in;  // staged inputs [state | des_state | des_inputs]
  int N, L, ns, nds, nz, nf;
  double *P, *p, *t, *Bb, *T1, *G, *M, *m0;
  uint8_t *cmp;      // cmp[a] = component 3i + q of the a-th free input of the current stage
  int8_t* inv;       // inverse map, -1 for pinned components
};

__device__ __forceinline__ double contact_of(const RicView& R, int i, int k) { return R.in[R.ns + R.nds + i * (4 * R.N + 3) + k]; }

// Free components of stage k (warp-uniform result m), B-bar = [Bf S; S] (nz x m), column a in Bb[.][a].
__device__ __forceinline__ int stage_setup(const RicView& R, const DevConfig& cfg, int k, int lane) {
  const int nf = R.nf;
  const double zeta = cfg.zoh ? 0.5 : 0.0;
  const bool st = lane < nf && contact_of(R, lane / 3, k) > 0.0;
  const unsigned mask = __ballot_sync(0xffffffffu, st);
  const int m = __popc(mask);
  if (lane < nf) R.inv[lane] = st ? (int8_t)__popc(mask & ((1u << lane) - 1u)) : (int8_t)-1;
  if (st) R.cmp[__popc(mask & ((1u << lane) - 1u))] = (uint8_t)lane;
  __syncwarp();
  if (lane < m) {  // only the nine state rows of B-bar are ever read: the selection rows are applied through cmp[]
    const int a = lane, c = R.cmp[a], i = c / 3, q = c - 3 * i;
    const double ce = contact_of(R, i, k), cm = ce * fast_rcp(cfg.mass), dt = cfg.dt;  // (a division is ~30 dependent instructions)
    const double* foot = R.in + R.ns + R.nds + i * (4 * R.N + 3) + R.N + 3 * k;
    const double* com = R.in + R.ns + 3 * k;
    const double r0 = foot[0] - com[0], r1 = foot[1] - com[1], r2 = foot[2] - com[2];  // frozen lever arm
    const double bp = zeta * dt * dt * cm, bv = dt * cm;
#pragma unroll
    for (int x = 0; x < 3; ++x) { R.Bb[x * kMu + a] = x == q ? bp : 0.0; R.Bb[(3 + x) * kMu + a] = x == q ? bv : 0.0; }
    // dt c [r]x e_q = dt c (r x e_q)
    const double s = dt * ce;
    const double v0 = q == 0 ? 0.0 : (q == 1 ? -r2 : r1);
    const double v1 = q == 0 ? r2 : (q == 1 ? 0.0 : -r0);
    const double v2 = q == 0 ? -r1 : (q == 1 ? r0 : 0.0);
    R.Bb[6 * kMu + a] = s * v0; R.Bb[7 * kMu + a] = s * v1; R.Bb[8 * kMu + a] = s * v2;
  }
  __syncwarp();
  return m;
}

// state weights of node `node` (1..N): diag(w0, w1, omega^2, w3..w8), omega inside the square (:205-210)
__device__ __forceinline__ double qdiag(const DevConfig& cfg, const double* qz, int node, int r) {
  return r != 2 ? cfg.w[r] : qz[node];  // qz[node] = ((w2 / 2) e^-node + w2 / 2)^2, tabulated once per CTA
}
// reference of node `node`, entry r of [c; v; L]
__device__ __forceinline__ double xref(const RicView& R, int node, int r) {
  return R.in[R.ns + (r / 3) * 3 * (R.N + 1) + 3 * node + (r % 3)];
}
// (A' v)[r] for A = [[I, dt I, 0], [0, I, 0], [0, 0, I]]
__device__ __forceinline__ double At_mul(const double* v, int r, double dt) { return (r >= 3 && r < 6) ? dt * v[r - 3] + v[r] : v[r]; }
// (A v)[r]
__device__ __forceinline__ double A_mul(const double* v, int r, double dt) { return r < 3 ? v[r] + dt * v[r + 3] : v[r]; }

// Triangular solves of one column with the stage factor Lf (strict lower = l, diagonal = 1 / l_cc); the
// number of free inputs MT is compiled in (3, 6, 9 or 12: whole legs), so no iteration is predicated.
template <int MT>
__device__ __forceinline__ void tri_fwd(const double* Lf, double* M, double* m0, int nz, int col, double (&y)[kMu]) {
#pragma unroll
  for (int a = 0; a < MT; ++a) {
    double acc = col < nz ? M[a * nz + col] : m0[a];
#pragma unroll
    for (int b = 0; b < a; ++b) acc -= Lf[a * kGld + b] * y[b];
    y[a] = acc * Lf[a * kGld + a];
  }
#pragma unroll
  for (int a = MT; a < kMu; ++a) y[a] = 0.0;
#pragma unroll
  for (int a = 0; a < MT; ++a) { if (col < nz) M[a * nz + col] = y[a]; else m0[a] = y[a]; }
}
template <int MT>
__device__ __forceinline__ void tri_bwd(const double* Lf, double (&y)[kMu]) {
#pragma unroll
  for (int a = MT - 1; a >= 0; --a) {
    double acc = y[a];
#pragma unroll
    for (int b = a + 1; b < MT; ++b) acc -= Lf[b * kGld + a] * y[b];
    y[a] = acc * Lf[a * kGld + a];
  }
}

}  // namespace

__global__ void __launch_bounds__(448) cmpc_riccati_kernel(const DevConfig cfg, const SolveArgs args) {
  extern __shared__ __align__(128) double smem[];
  pdl_prologue(args.pdl_trigger > 1);
  constexpr int W = 1, GT = 32;
  const int N = cfg.N, L = cfg.L;
  const int nf = 3 * L, nz = 9 + nf, nfN = nf * N, nbfull = L * N, mfull = 5 * nbfull;
  const int ns = 9 + 3 * L, nds = 9 * (N + 1), ndi = L * (4 * N + 3);
  const RicPlan& P = args.ric;
  Group<W> G;
  G.gtid = threadIdx.x % GT;
  G.gid = threadIdx.x / GT;
  G.red = nullptr;
  const int lane = G.gtid;
  // CTA-shared: pair e of a lower triangle stored row by row -> (row << 4 | column), 78 pairs for 12 x 12
  uint8_t* c_tri = reinterpret_cast<uint8_t*>(smem);
  uint16_t* c_trz = reinterpret_cast<uint16_t*>(smem + 10);  // the same for the nz x nz cost-to-go: (row << 8 | column), 231 pairs at nz = 21
  double* c_qz = smem + 68;
  double* base = smem + kRicCta + (size_t)G.gid * P.total;
  int* s_misc = reinterpret_cast<int*>(base + P.ints);  // [0]=nb, [1]=invalid, [2]=work slot, [3]=nb unclamped
  uint8_t* s_blk_j = reinterpret_cast<uint8_t*>(s_misc + 4);
  uint8_t* s_blk_i = s_blk_j + nbfull;
  int8_t* s_blk_of = reinterpret_cast<int8_t*>(s_blk_i + nbfull);
  BuildView V;
  V.Mm = base + P.in;
  V.qzt = c_qz + 1;  // the table is indexed by node
  V.eq = base + P.P; V.qz = V.eq + 9 * N; V.fz = V.qz + N + (N & 1); V.ce = V.fz + nbfull;  // scratch of the shared prologue, dead after it
  V.arm = nullptr; V.g = nullptr; V.tb = nullptr;
  V.misc = s_misc; V.blk_j = s_blk_j; V.blk_i = s_blk_i; V.blk_of = s_blk_of;
  RicView R;
  R.in = V.Mm; R.N = N; R.L = L; R.ns = ns; R.nds = nds; R.nz = nz; R.nf = nf;
  R.P = base + P.P; R.p = base + P.p; R.t = base + P.t; R.Bb = base + P.Bb; R.T1 = base + P.T1;
  R.G = base + P.G; R.M = base + P.M; R.m0 = base + P.m0;
  R.cmp = reinterpret_cast<uint8_t*>(s_blk_of + nbfull);
  R.inv = reinterpret_cast<int8_t*>(R.cmp + 16);
  double* s_xc = base + P.xc;       // current state (9) and previous force (nf): z = [x; Fp]
  double* s_lam = s_xc + nz;        // costates: of the solution [0..9), of the zero-input roll-out [9..18)
  double* s_u = s_lam + 18;         // stage inputs (m)
  double* s_X = base + P.X;         // x_1 .. x_N after the backward sweep
  double* s_F = base + P.F;         // forces, step-major [N][nf]
  double* slab = args.scratch + (size_t)(blockIdx.x * args.groups + G.gid) * args.scratch_per_group;
  const int kstride = kMu * nz + kMu;  // gains of one stage: K' [nz][12], then kff [12]
  const double dt = cfg.dt, mass = cfg.mass;
  const double zeta = cfg.zoh ? 0.5 : 0.0;
  const double dpz = zeta * dt * dt * (-kGrav), dvz = dt * (-kGrav);  // affine term d = [dpz e_z; dvz e_z; 0]
  const int nl = args.nlists > 0 ? args.nlists : 1;
  int total = 0;
  for (int q = 0; q < nl; ++q) {
    const int c = args.nlists > 0 ? *args.lcount[q] : (args.count ? *args.count : args.count_imm);
    total += c > 0 ? c : 0;
  }
  if (total <= 0) return;  // empty lists (uniform over the grid): nothing to set up
  int li = 0;
  for (int e = threadIdx.x; e <= N; e += blockDim.x) {
    const double om = (cfg.w[2] * 0.5) * exp(-(double)e) + cfg.w[2] * 0.5;  // CentroidalMPC.cpp:205
    c_qz[e] = om * om;
  }
  for (int e = threadIdx.x; e < 78; e += blockDim.x) {
    int rr = 0;
    while (((rr + 1) * (rr + 2)) >> 1 <= e) ++rr;
    c_tri[e] = (uint8_t)((rr << 4) | (e - ((rr * (rr + 1)) >> 1)));
  }
  for (int e = threadIdx.x; e < ((nz * (nz + 1)) >> 1); e += blockDim.x) {
    int rr = 0;
    while (((rr + 1) * (rr + 2)) >> 1 <= e) ++rr;
    c_trz[e] = (uint16_t)((rr << 8) | (e - ((rr * (rr + 1)) >> 1)));
  }
  __syncthreads();

  while (true) {
    const int count = args.nlists > 0 ? *args.lcount[li] : (args.count ? *args.count : args.count_imm);
    int slot = 0;
    if (lane == 0) slot = atomicAdd(args.nlists > 0 ? args.lwork[li] : args.work, 1);
    slot = G.bcast0(slot, s_misc + 2);
    if (slot >= count) {
      if (++li >= nl) break;
      continue;
    }
    const int32_t* perm = args.nlists > 0 ? args.lperm[li] : args.perm;
    const int inst = perm ? perm[slot] : slot;
    const bool finite = stage_inputs<W>(G, cfg, args, inst, V);
    const bool invalid = s_misc[1] != 0;
    if (!finite || invalid) {
      for (int t = lane; t < nfN; t += GT) args.forces[(size_t)inst * nfN + t] = 0.0;
      if (args.lam) for (int t = lane; t < 2 * mfull; t += GT) args.lam[(size_t)inst * 2 * mfull + t] = 0.0;
      if (args.active) for (int t = lane; t < nbfull; t += GT) args.active[(size_t)inst * nbfull + t] = 0;
      if (lane == 0) {
        args.status[inst] = !finite ? CMPC_STATUS_NUMERICAL : CMPC_STATUS_INVALID_TABLE;
        if (args.iters) args.iters[inst] = 0;
        if (args.kkt) args.kkt[inst] = 0.0;
      }
      __syncwarp();
      continue;
    }
    bool defer = false;
    if (args.warm_active) {  // a warm-start guess with active rows belongs to the IPM kernel's polish
      bool any = false;
      const uint16_t* wa = args.warm_active + (size_t)inst * nbfull;
      for (int t = lane; t < nbfull; t += GT) { const unsigned a = wa[t]; any = any || (!(a & 0x8000u) && (a & 0x3ffu)); }
      defer = !G.all(!any);
    }

    if (!defer) {
      // ---- backward sweep.  V_N = (x - xr_N)' Q_N (x - xr_N)
      for (int e = lane; e < nz * nz; e += GT) R.P[e] = 0.0;
      __syncwarp();
      if (lane < 9) { const double qd = qdiag(cfg, c_qz, N, lane); R.P[lane * nz + lane] = qd; R.p[lane] = -qd * xref(R, N, lane); }
      else if (lane < nz) R.p[lane] = 0.0;
      __syncwarp();
      bool ok = true;
      for (int k = N - 1; k >= 0; --k) {
        const int m = stage_setup(R, cfg, k, lane);
        const double rate = k >= 1 ? 1.0 : 0.0;
        double colsum = 0.0;
        for (int i = 0; i < L; ++i) colsum += contact_of(R, i, k);
        const double fz = mass * kGrav * fast_rcp(colsum);  // desired fz of the stance legs (:331-333)
        // t = P dbar + p
        if (lane < nz) R.t[lane] = R.P[lane * nz + 2] * dpz + R.P[lane * nz + 5] * dvz + R.p[lane];
        // T1 = P Bbar : rows of Bbar below the state block are a selection
        if (lane < nz) {  // lane = row r: the state part of the P row stays in registers, Bbar entries are broadcast loads
          const int r = lane;
          double pr[9];
#pragma unroll
          for (int s = 0; s < 9; ++s) pr[s] = R.P[r * nz + s];
          for (int a = 0; a < m; ++a) {
            double acc = R.P[r * nz + 9 + R.cmp[a]];
#pragma unroll
            for (int s = 0; s < 9; ++s) acc = fma(pr[s], R.Bb[s * kMu + a], acc);
            R.T1[r * kMu + a] = acc;
          }
        }
        __syncwarp();
        // G = Bbar' T1 + S'(Wf + rate Wr) S  (lower triangle), M = T1' Abar - rate [0, S' Wr], m0
        for (int e = lane; e < kMu * kMu; e += GT) {
          const int a = e / kMu, b = e - a * kMu;
          if (a < m && b <= a) {
            const int c = R.cmp[a];
            double acc = R.T1[(9 + c) * kMu + b];
#pragma unroll
            for (int s = 0; s < 9; ++s) acc = fma(R.Bb[s * kMu + a], R.T1[s * kMu + b], acc);
            if (a == b) acc += cfg.w[9 + 3 * L + c] + rate * cfg.w[9 + 6 * L + c];
            R.G[a * kGld + b] = acc;
          }
        }
        if (lane < m) {  // row a of M = (T1 column a)' Abar, minus the rate coupling to the previous force
          const int a = lane, c = R.cmp[a];
          double* Ma = R.M + a * nz;
#pragma unroll
          for (int x = 0; x < 3; ++x) {
            const double tp = R.T1[x * kMu + a];
            Ma[x] = tp; Ma[3 + x] = dt * tp + R.T1[(3 + x) * kMu + a]; Ma[6 + x] = R.T1[(6 + x) * kMu + a];
          }
          for (int col = 9; col < nz; ++col) Ma[col] = 0.0;
          Ma[9 + c] = -rate * cfg.w[9 + 6 * L + c];
        }
        if (lane < m) {
          const int a = lane, c = R.cmp[a];
          double acc = R.t[9 + c];
#pragma unroll
          for (int s = 0; s < 9; ++s) acc = fma(R.Bb[s * kMu + a], R.t[s], acc);
          if (c % 3 == 2) acc -= cfg.w[9 + 3 * L + c] * fz;
          R.m0[a] = acc;
        }
        __syncwarp();
        // Cholesky of G (m x m).  The factor goes to Lf (the T1 buffer, dead by now): strict lower = l_rc,
        // diagonal = 1 / l_cc; the trailing update works from the UNSCALED column (times 1 / d), so scaling and
        // update of a column need no barrier between them -- one barrier per column.
        double* Lf = R.T1;
        for (int c = 0; c < m; ++c) {
          const double d = R.G[c * kGld + c];
          ok = ok && d > 0.0;
          const double inv = fast_rsqrt(d), inv2 = inv * inv;
          if (lane > c && lane < m) Lf[lane * kGld + c] = R.G[lane * kGld + c] * inv;
          if (lane == c) Lf[c * kGld + c] = inv;
          const int mt = m - c - 1, npair = (mt * (mt + 1)) >> 1;
          for (int e = lane; e < npair; e += GT) {  // one (row, column) pair of the trailing lower triangle per lane and pass
            const int rr = c_tri[e] >> 4, cc = c_tri[e] & 15;
            const int r = c + 1 + rr, c2 = c + 1 + cc;
            R.G[r * kGld + c2] -= R.G[r * kGld + c] * R.G[c2 * kGld + c] * inv2;
          }
          __syncwarp();
        }
        ok = __all_sync(0xffffffffu, ok);
        if (!ok) break;
        // Y = L^-1 [M | m0]: lane = column (nz columns of M, column nz = m0), rows in registers
        double y[kMu];
        const int col = lane;
        if (col <= nz) {
          switch (m) {
            case 3: tri_fwd<3>(Lf, R.M, R.m0, nz, col, y); break;
            case 6: tri_fwd<6>(Lf, R.M, R.m0, nz, col, y); break;
            case 9: tri_fwd<9>(Lf, R.M, R.m0, nz, col, y); break;
            default: tri_fwd<12>(Lf, R.M, R.m0, nz, col, y); break;
          }
        }
        __syncwarp();
        if (k >= 1) {
          // P <- blkdiag(Q_k, Wr) + Abar' P Abar - Y'Y ;  p <- [-Q_k xr_k; 0] + Abar' t - Y' y0
          for (int e = lane; e < 27; e += GT) { const int r = e / 3, c3 = e - 3 * r; R.P[r * nz + 3 + c3] += dt * R.P[r * nz + c3]; }  // X A
          __syncwarp();
          for (int e = lane; e < 27; e += GT) { const int c = e / 3, r3 = e - 3 * c; R.P[(3 + r3) * nz + c] += dt * R.P[r3 * nz + c]; }  // A'(X A)
          __syncwarp();
          {
            // P <- base - Y'Y in 2 x 2 blocks of the lower triangle (block pair e -> (br, bc) from c_tri): four
            // accumulators per four loads; m is a multiple of 3 (whole legs), so the a-loop needs no predicate
            const int nzb = (nz + 1) >> 1, nblkp = (nzb * (nzb + 1)) >> 1;
            for (int e = lane; e < nblkp; e += GT) {
              const int br = c_tri[e] >> 4, bc = c_tri[e] & 15;
              const int r0 = 2 * br, c0 = 2 * bc;
              const int r1 = r0 + 1 < nz ? r0 + 1 : r0, c1 = c0 + 1 < nz ? c0 + 1 : c0;  // (clamped: duplicates are not stored)
              double a00 = 0.0, a01 = 0.0, a10 = 0.0, a11 = 0.0;
              for (int a = 0; a < m; a += 3) {
#pragma unroll
                for (int x = 0; x < 3; ++x) {
                  const double* Ya = R.M + (a + x) * nz;
                  const double yr0 = Ya[r0], yr1 = Ya[r1], yc0 = Ya[c0], yc1 = Ya[c1];
                  a00 = fma(yr0, yc0, a00); a01 = fma(yr0, yc1, a01); a10 = fma(yr1, yc0, a10); a11 = fma(yr1, yc1, a11);
                }
              }
              auto put = [&](int r, int c, double yy) {  // entry (r, c), c <= r, and its mirror
                double v = (r < 9) ? R.P[r * nz + c] : 0.0;  // only the state block of Abar' P Abar is non-zero
                if (r == c) v += r < 9 ? qdiag(cfg, c_qz, k, r) : cfg.w[9 + 6 * L + r - 9];
                v -= yy;
                R.P[r * nz + c] = v; R.P[c * nz + r] = v;
              };
              put(r0, c0, a00);
              if (r0 + 1 < nz) {
                put(r0 + 1, c0, a10);
                if (c0 + 1 < nz && c0 + 1 <= r0 + 1) put(r0 + 1, c0 + 1, a11);
              }
              if (c0 + 1 < nz && c0 + 1 <= r0) put(r0, c0 + 1, a01);  // (upper entry of a diagonal block: written by its mirror)
            }
          }
          double pn = 0.0;
          if (lane < nz) {
            pn = lane < 9 ? At_mul(R.t, lane, dt) - qdiag(cfg, c_qz, k, lane) * xref(R, k, lane) : 0.0;
            for (int a = 0; a < m; a += 3)
              pn -= R.M[a * nz + lane] * R.m0[a] + R.M[(a + 1) * nz + lane] * R.m0[a + 1] + R.M[(a + 2) * nz + lane] * R.m0[a + 2];
          }
          __syncwarp();
          if (lane < nz) R.p[lane] = pn;
        }
        // gains K = L^-T Y, kff = L^-T y0 -> L2 slab, transposed (K'[col][a]) for the forward sweep
        if (col <= nz) {
          switch (m) {
            case 3: tri_bwd<3>(Lf, y); break;
            case 6: tri_bwd<6>(Lf, y); break;
            case 9: tri_bwd<9>(Lf, y); break;
            default: tri_bwd<12>(Lf, y); break;
          }
          double* dst = slab + (size_t)k * kstride + col * kMu;
#pragma unroll
          for (int a = 0; a < kMu; ++a) __stcg(dst + a, a < m ? y[a] : 0.0);
        }
        __syncwarp();
      }
      defer = !ok;
    }

    double gs = 1.0, usf = 1.0, stat = 0.0, prim = 0.0;
    if (!defer) {
      // ---- forward roll-out  u_k = -K_k [x_k; F_{k-1}] - kff_k
      if (lane < 9) s_xc[lane] = R.in[lane];
      else if (lane < nz) s_xc[lane] = 0.0;
      for (int e = lane; e < nfN; e += GT) s_F[e] = 0.0;
      __syncwarp();
      for (int k = 0; k < N; ++k) {
        const int m = stage_setup(R, cfg, k, lane);
        const double* gk = slab + (size_t)k * kstride;
        if (lane < m) {
          double acc = __ldcg(gk + nz * kMu + lane);
          for (int c = 0; c < nz; ++c) acc = fma(__ldcg(gk + c * kMu + lane), s_xc[c], acc);
          s_u[lane] = -acc;
          s_F[k * nf + R.cmp[lane]] = -acc;
        }
        __syncwarp();
        double xn = 0.0;
        if (lane < 9) {
          xn = A_mul(s_xc, lane, dt) + (lane == 2 ? dpz : (lane == 5 ? dvz : 0.0));
          for (int a = 0; a < m; ++a) xn = fma(R.Bb[lane * kMu + a], s_u[a], xn);
        }
        __syncwarp();
        if (lane < 9) { s_xc[lane] = xn; s_X[9 * k + lane] = xn; }
        else if (lane < nz) s_xc[lane] = s_F[k * nf + lane - 9];
        __syncwarp();
      }
      // ---- verification: (H U + g) and g on the free entries by adjoint sweeps, constraint rows
      double gmax = 0.0, umax = 0.0;
      bool fin = true;
      if (lane < 18) s_lam[lane] = 0.0;
      __syncwarp();
      for (int k = N - 1; k >= 0; --k) {
        const int m = stage_setup(R, cfg, k, lane);
        double l1 = 0.0, l2 = 0.0;
        if (lane < 9) {
          l1 = 2.0 * qdiag(cfg, c_qz, k + 1, lane) * (s_X[9 * k + lane] - xref(R, k + 1, lane)) + At_mul(s_lam, lane, dt);
          // zero-input state of node k + 1 in closed form (x = A^node x0 + sum A^p d), for g = gradient at U = 0
          const double kk = (double)(k + 1);
          double xz = R.in[lane];
          if (lane < 3) xz += kk * dt * R.in[3 + lane];
          if (lane == 2) xz += (cfg.zoh ? 0.5 * kk * kk : 0.5 * kk * (kk - 1.0)) * dt * dt * (-kGrav);
          if (lane == 5) xz += kk * dt * (-kGrav);
          l2 = 2.0 * qdiag(cfg, c_qz, k + 1, lane) * (xz - xref(R, k + 1, lane)) + At_mul(s_lam + 9, lane, dt);
        }
        __syncwarp();
        if (lane < 9) { s_lam[lane] = l1; s_lam[9 + lane] = l2; }
        __syncwarp();
        double colsum = 0.0;
        for (int i = 0; i < L; ++i) colsum += contact_of(R, i, k);
        const double fz = mass * kGrav * fast_rcp(colsum);
        if (lane < m) {
          const int a = lane, c = R.cmp[a];
          const double wf = cfg.w[9 + 3 * L + c], wr = cfg.w[9 + 6 * L + c];
          const double f = s_F[k * nf + c], fr = (c % 3 == 2) ? fz : 0.0;
          double g1 = 2.0 * wf * (f - fr), g0 = -2.0 * wf * fr;
#pragma unroll
          for (int s = 0; s < 9; ++s) { g1 = fma(R.Bb[s * kMu + a], s_lam[s], g1); g0 = fma(R.Bb[s * kMu + a], s_lam[9 + s], g0); }
          if (k >= 1) g1 += 2.0 * wr * (f - s_F[(k - 1) * nf + c]);
          if (k + 1 < N) g1 -= 2.0 * wr * (s_F[(k + 1) * nf + c] - f);
          stat = fmax(stat, fabs(g1)); gmax = fmax(gmax, fabs(g0)); umax = fmax(umax, fabs(f));
          fin = fin && isfinite(f) && isfinite(g1);
        }
        if (lane < L && contact_of(R, lane, k) > 0.0) {
          const double ce = contact_of(R, lane, k);
          const double ubxy = kFricUb * ce, ubz = mass * kGrav * (double)L * ce;
          double yv[5];
          cmul5(cfg.mu[lane], s_F + k * nf + 3 * lane, yv);
          for (int q = 0; q < 5; ++q) prim = fmax(prim, fmax(-yv[q], yv[q] - (q < 4 ? ubxy : ubz)));
        }
        __syncwarp();
      }
      gmax = G.max(gmax); umax = G.max(umax); stat = G.max(stat); prim = G.max(prim);
      gs = 1.0 + gmax; usf = 1.0 + umax;
      defer = !(stat <= 1e-9 * gs && prim <= 1e-9 * usf);
      defer = !G.all(!defer && fin);
    }
    if (defer) {
      if (lane == 0) {
        if (args.nlists > 0) args.lfail_perm[li][atomicAdd(args.lfail_count[li], 1)] = inst;
        else args.fail_perm[atomicAdd(args.fail_count, 1)] = inst;
      }
      __syncwarp();
      continue;
    }
    // ---- outputs (same conventions as the dense kernels): forces [L][N][3] (CentroidalMPC.cpp:270)
    for (int t = lane; t < nfN; t += GT) {
      const int i = t / (3 * N), rem = t - i * 3 * N, j = rem / 3, q = rem - 3 * j;
      args.forces[(size_t)inst * nfN + t] = s_F[j * nf + 3 * i + q];
    }
    if (args.lam) for (int t = lane; t < 2 * mfull; t += GT) args.lam[(size_t)inst * 2 * mfull + t] = 0.0;
    if (args.active) {
      for (int t = lane; t < nbfull; t += GT) {
        const int j = t / L, i = t - j * L;
        const double ce = contact_of(R, i, j);
        uint16_t a = 0x8000;
        if (ce > 0.0) {
          const double ubxy = kFricUb * ce, ubz = mass * kGrav * (double)L * ce;
          double yv[5];
          cmul5(cfg.mu[i], s_F + j * nf + 3 * i, yv);
          a = 0;
          for (int q = 0; q < 5; ++q)
            a |= (uint16_t)((yv[q] <= 1e-9 * usf ? 1 : 0) << q | (((q < 4 ? ubxy : ubz) - yv[q]) <= 1e-9 * usf ? 1 : 0) << (5 + q));
        }
        args.active[(size_t)inst * nbfull + t] = a;
      }
    }
    if (lane == 0) {
      args.status[inst] = CMPC_STATUS_OK;
      if (args.iters) args.iters[inst] = 0;
      if (args.kkt) args.kkt[inst] = fmax(stat * fast_rcp(gs), fmax(prim, 0.0) * fast_rcp(usf));
    }
    __syncwarp();
  }
}

cudaError_t launch_riccati_kernel(int grid, int block, size_t smem, cudaStream_t stream, const DevConfig& cfg,
                                  const SolveArgs& args) {
  return launch_ex(cmpc_riccati_kernel, grid, block, smem, stream, args.pdl != 0, cfg, args);
}

cudaError_t set_riccati_kernel_smem(size_t bytes) {
  return cudaFuncSetAttribute(cmpc_riccati_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

}  // namespace cmpc
