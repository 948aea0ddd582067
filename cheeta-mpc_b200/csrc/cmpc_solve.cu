// cmpc_solve.cu -- the fused build + interior-point + polish kernel (sm_100a) and its launcher.
#include "cmpc_device.cuh"

namespace cmpc {

// ------------------------------------------------------------------ the fused kernel
// MODE 0: solve.  MODE 1: build-export (H, g in the full 3LN layout to global memory).
// PHASE splits the solve so that each kernel's hot loop fits the instruction cache (the fused kernel is 30 k SASS
// instructions; with the resident warps of an SM in different phases of different instances more than half of its stall
// cycles on the constrained workload were instruction fetch):
//   0  fused: build + interior point + polish + outputs (every path; also the catch-all for what the split hands back)
//   1  interior point only: build, start point, Mehrotra iterations until the polish would be attempted; the iterate
//      (u, zl, zu, gs, it) goes to a state slot and the instance to the polish list.  Any other way out of the loop
//      (iteration cap, numerical trouble) hands the instance to the fall-back list, which the fused kernel solves from
//      scratch -- deterministically the same result.
//   2  polish only: build, load the iterate, the verified active-set polish, outputs; a rejected polish (rare: the
//      fused kernel would go back to the interior point) hands the instance to the fall-back list.
template <int W, int MODE, bool MS, int PHASE>
__global__ void __launch_bounds__(W == 2 ? 512 : (W == 1 ? 320 : 256), 1) cmpc_solve_kernel(const DevConfig cfg, const SolveArgs args) {
  extern __shared__ __align__(128) double smem[];
  pdl_prologue(args.pdl_trigger > 1);
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
  double* c_z1 = smem;      // CTA-shared z-weighted sums for the build (fill_z_tables)
  double* c_z2 = smem + N;
  double* c_qz = smem + 2 * N;  // z-position weight per node (stage_inputs)
  double* base = smem + P.cta + (size_t)G.gid * P.total;
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
  // work lists: one (perm / count / work) or several drained in order (SolveArgs::nlists)
  const int nl = args.nlists > 0 ? args.nlists : 1;
  int total = 0;
  for (int q = 0; q < nl; ++q) {
    const int c = args.nlists > 0 ? *args.lcount[q] : (args.count ? *args.count : args.count_imm);
    total += c > 0 ? c : 0;
  }
  // for the host's next launch plan: a word in mapped host memory, written only when the value changes (a device-side
  // shadow tells): in steady state no launch touches the bus for it
  if (args.hint_out && blockIdx.x == 0 && threadIdx.x == 0 && *args.hint_shadow != total) { *args.hint_shadow = total; *args.hint_out = total; }
  if (total <= 0) return;  // empty lists (uniform over the grid): nothing to set up
  if ((int)threadIdx.x < N) {
    fill_z_tables(cfg, c_z1, c_z2, threadIdx.x);
    const double om = (cfg.w[2] * 0.5) * exp(-(double)(threadIdx.x + 1)) + cfg.w[2] * 0.5;  // node k + 1, CentroidalMPC.cpp:205
    c_qz[threadIdx.x] = om * om;
  }
  __syncthreads();
  int li = 0;

  while (true) {
    const int count = args.nlists > 0 ? *args.lcount[li] : (args.count ? *args.count : args.count_imm);
    int slot = 0;
    if (gtid == 0) slot = atomicAdd(args.nlists > 0 ? args.lwork[li] : args.work, 1);
    slot = G.bcast0(slot, s_misc + 2);
    if (slot >= count) {
      if (++li >= nl) break;
      continue;
    }
    const int32_t* lperm = args.nlists > 0 ? args.lperm[li] : args.perm;
    const int inst = lperm ? lperm[slot] : slot;
    BuildView V;
    V.Mm = Mm; V.ce = s_ce; V.fz = s_fz; V.arm = s_arm; V.eq = s_eq; V.qz = s_qz; V.g = s_g; V.qzt = c_qz;
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
    build_qp<W>(G, cfg, V, Hb, nb, c_z1, c_z2);
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

    // ---- strictly feasible start f = (0, 0, fz0); centred duals  (phase 2: the interior-point kernel's iterate)
    const double* xs = nullptr;
    if constexpr (PHASE == 2) xs = args.xstate + (size_t)slot * args.xstride;
    for (int b = gtid; b < nb; b += GT) {
      if constexpr (PHASE == 2) {
        s_u[3 * b] = __ldcg(xs + 3 * b); s_u[3 * b + 1] = __ldcg(xs + 3 * b + 1); s_u[3 * b + 2] = __ldcg(xs + 3 * b + 2);
      } else {
        const double mub = cfg.mu[s_blk_i[b]];
        double fz = s_fz[b];
        fz = fmin(fz, 0.5 * mass * kGrav * (double)L * s_ce[b]);
        fz = fmin(fz, 0.5 * kFricUb * s_ce[b] / mub);
        s_u[3 * b] = 0.0; s_u[3 * b + 1] = 0.0; s_u[3 * b + 2] = fz;
      }
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
      if constexpr (PHASE == 2) {
        for (int q = 0; q < 5; ++q) {
          s_zl[5 * b + q] = __ldcg(xs + args.n4max + 5 * b + q); s_zu[5 * b + q] = __ldcg(xs + args.n4max + mmax + 5 * b + q);
        }
      } else {
        const double mub = cfg.mu[s_blk_i[b]];
        const double ubxy = kFricUb * s_ce[b], ubz = mass * kGrav * (double)L * s_ce[b];  // :183,199
        double y[5];
        cmul5(mub, s_u + 3 * b, y);
        for (int q = 0; q < 5; ++q) {
          const double ub = q < 4 ? ubxy : ubz;
          s_zl[5 * b + q] = mu0 / y[q]; s_zu[5 * b + q] = mu0 / (ub - y[q]);
        }
      }
    }
    G.sync();

    int status = CMPC_STATUS_MAX_ITER, it = 0, npolish = 0;
    bool numerical = false, ipm_ok = false, m_is_h = true, warm_done = false, handed_over = false;
    double us = 1.0;
    if constexpr (PHASE == 2) it = (int)__ldcg(xs + args.n4max + 2 * mmax);
    const int it_first = it, it_last = PHASE == 2 ? it : cfg.max_iter;
    for (it = it_first; it <= it_last; ++it) {
      // ---- residuals (M holds a fresh copy of H here)
      if (!m_is_h) { copy_mat<W>(G, Mm, Hm, matd); G.sync(); m_is_h = true; }
      if (it > 0 || PHASE == 2) symv_bc4<W>(G, Mm, n4, nblk, s_u, s_rd);  // fused, it == 0: rd still holds H u0 from the start point
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
      const bool ready = PHASE == 2 || (conv_mu && rmax <= 1e4 * cfg.tol * gs);  // (phase 2 exists because phase 1 saw `ready`)
      ipm_ok = conv_mu && rmax <= 10.0 * cfg.tol * gs;
      // Warm start (closed loop): before the first factorisation, try the previous tick's
      // active set, shifted by one step, as the polish's guess. The polish verifies the KKT
      // conditions, so a wrong guess only costs its correction passes and the IPM runs cold.
      const bool warm_now = cfg.polish && args.warm_active != nullptr && it == 0 && !warm_done;
      if constexpr (PHASE == 1) {
        if (ready) {
          // hand the iterate to the polish kernel: state slot = position in its list
          int ps = 0;
          if (gtid == 0) ps = atomicAdd(args.pol_count, 1);
          ps = G.bcast0(ps, s_misc + 2);
          double* xo = args.xstate + (size_t)ps * args.xstride;
          for (int t = gtid; t < n4; t += GT) __stcg(xo + t, s_u[t]);
          for (int t = gtid; t < m; t += GT) { __stcg(xo + args.n4max + t, s_zl[t]); __stcg(xo + args.n4max + mmax + t, s_zu[t]); }
          if (gtid == 0) { __stcg(xo + args.n4max + 2 * mmax, (double)it); args.pol_perm[ps] = inst; }
          handed_over = true;
          break;
        }
      }
      if constexpr (PHASE != 1)
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
          if (args.phase_lock) __syncthreads_and(0);  // the groups of the CTA enter a polish pass together (shared instruction fetch)
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
          if (args.phase_lock > 1) __syncthreads_and(0);
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
                int k = 0;
                for (int q = 0; q < 5; ++q) if (s_actl[5 * b + q]) { row_vec(mub, q, Nrm[k]); ++k; }
                for (int q = 0; q < 5; ++q) if (s_actu[5 * b + q]) {
                  row_vec(mub, q, Nrm[k]);
                  Nrm[k][0] = -Nrm[k][0]; Nrm[k][1] = -Nrm[k][1]; Nrm[k][2] = -Nrm[k][2];
                  ++k;
                }
                okm = block_multipliers(k, Nrm, rb, 1e-9 * gs, lam) && okm;
                // scatter back in the order the rows were collected (statically indexed targets: a dynamically indexed
                // register array becomes a select chain per element -- 1.7 k instructions of this kernel were that)
                int kk = 0;
                for (int q = 0; q < 5; ++q) if (s_actl[5 * b + q]) ll[q] = lam[kk++];
                for (int q = 0; q < 5; ++q) if (s_actu[5 * b + q]) lu[q] = lam[kk++];
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
        if constexpr (PHASE == 2) break;  // not accepted: the fused kernel takes the instance from scratch (fall-back list)
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
      if constexpr (PHASE != 2) {

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
      }  // PHASE != 2
    }
    if constexpr (PHASE != 0) {
      // split kernels: phase 1 never writes outputs; phase 2 only for an accepted polish.  Everything else goes to the
      // fall-back list of the fused kernel.
      if ((PHASE == 1 && !handed_over) || (PHASE == 2 && status != CMPC_STATUS_OK)) {
        if (gtid == 0) args.fb_perm[atomicAdd(args.fb_count, 1)] = inst;
      }
      if (PHASE == 1 || status != CMPC_STATUS_OK) { G.sync(); continue; }
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
  if (args.phase_lock) while (!__syncthreads_and(1)) {}  // keep the other groups' barriers company until everybody is done
}


namespace {
template <int W, int MODE, bool MS, int PHASE>
cudaError_t launch_t(int grid, int block, size_t smem, cudaStream_t stream, const DevConfig& cfg, const SolveArgs& args) {
  return launch_ex(cmpc_solve_kernel<W, MODE, MS, PHASE>, grid, block, smem, stream, args.pdl != 0, cfg, args);
}
template <int W, int MODE, bool MS, int PHASE>
cudaError_t attr_t(size_t bytes) {
  return cudaFuncSetAttribute(cmpc_solve_kernel<W, MODE, MS, PHASE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}
template <int W>
cudaError_t launch_w(int phase, int grid, int block, size_t smem, cudaStream_t stream, const DevConfig& cfg, const SolveArgs& args) {
  switch (phase) {
    case 0: return launch_t<W, 0, true, 0>(grid, block, smem, stream, cfg, args);
    case 1: return launch_t<W, 0, true, 1>(grid, block, smem, stream, cfg, args);
    case 2: return launch_t<W, 0, true, 2>(grid, block, smem, stream, cfg, args);
  }
  return cudaErrorInvalidValue;
}
template <int W>
cudaError_t attr_w(size_t bytes) {
  cudaError_t e = attr_t<W, 0, true, 0>(bytes);
  if (e == cudaSuccess) e = attr_t<W, 0, true, 1>(bytes);
  if (e == cudaSuccess) e = attr_t<W, 0, true, 2>(bytes);
  return e;
}
}  // namespace

// phase: 0 fused, 1 interior point only, 2 polish only (the split exists for the shared-memory-factor variants)
cudaError_t launch_solve_kernel(int W, int mode, bool ms, int phase, int grid, int block, size_t smem, cudaStream_t stream,
                                const DevConfig& cfg, const SolveArgs& args) {
  if (mode == 1) return launch_t<8, 1, false, 0>(grid, block, smem, stream, cfg, args);
  if (!ms) return launch_t<8, 0, false, 0>(grid, block, smem, stream, cfg, args);
  switch (W) {
    case 1: return launch_w<1>(phase, grid, block, smem, stream, cfg, args);
    case 2: return launch_w<2>(phase, grid, block, smem, stream, cfg, args);
    case 4: return launch_w<4>(phase, grid, block, smem, stream, cfg, args);
    case 8: return launch_w<8>(phase, grid, block, smem, stream, cfg, args);
  }
  return cudaErrorInvalidValue;
}

cudaError_t set_solve_kernel_smem(int W, int mode, bool ms, size_t bytes) {
  if (mode == 1) return attr_t<8, 1, false, 0>(bytes);
  if (!ms) return attr_t<8, 0, false, 0>(bytes);
  switch (W) {
    case 1: return attr_w<1>(bytes);
    case 2: return attr_w<2>(bytes);
    case 4: return attr_w<4>(bytes);
    case 8: return attr_w<8>(bytes);
  }
  return cudaErrorInvalidValue;
}

}  // namespace cmpc
