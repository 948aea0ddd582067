// cmpc_presolve.cu -- the presolve kernel (sm_100a) and its launcher.
//
// Most ticks of a legged MPC have no friction or force-limit row active (the reference's weights make
// force tracking dominate, CentoidMPCTest.cpp:19-33), and then the optimum of the condensed QP is the
// unconstrained minimiser -H^-1 g.  This kernel settles exactly those instances with ONE Cholesky of
// H: build, factor (forward substitution fused), back-substitute, then verify on the problem itself --
// stationarity |H u + g| <= 1e-9 gs with H u from a roll-out / adjoint sweep over the dynamics and weights (not
// from the matrix that was factored) and every row of 0 <= F f <= ub satisfied to -1e-9 us, i.e. the
// polish's own acceptance test with an empty working set -- and write the outputs (status OK, iters
// 0, multipliers 0).  Anything else (a violated row, a warm-start guess with active rows, a failed
// pivot) is appended to fail_perm and goes through the interior-point kernel (cmpc_solve.cu).
//
// Per group (one warp at n <= 60) it keeps in shared memory only the matrix in the chunk-major tile
// layout (cmpc_device.cuh, make_pre_plan), ONE n-vector and a few tables; right-hand side, gradient
// and residual live in registers (one or two rows per lane).  2066 doubles per instance at n = 60:
// 14 instances per SM, so a 4096-instance batch runs in two waves.
#include <type_traits>

#include "cmpc_device.cuh"

namespace cmpc {

namespace {

// ------------------------------------------------------------------ chunk-major tiles
template <int TT>
struct CM {
  int Trt;
  __device__ __forceinline__ int T() const { return TT ? TT : Trt; }
  // whole tile t -> 16 registers (row-major)
  __device__ __forceinline__ void ld(const double2* M2, int t, double* r) const {
#pragma unroll
    for (int q = 0; q < 8; ++q) { const double2 v = M2[q * T() + t]; r[2 * q] = v.x; r[2 * q + 1] = v.y; }
  }
  __device__ __forceinline__ void st(double2* M2, int t, const double* r) const {
#pragma unroll
    for (int q = 0; q < 8; ++q) M2[q * T() + t] = make_double2(r[2 * q], r[2 * q + 1]);
  }
  // row a of tile t: two chunks
  __device__ __forceinline__ const double2* row(const double2* M2, int t, int a) const { return M2 + (2 * a) * T() + t; }
  __device__ __forceinline__ double2* row(double2* M2, int t, int a) const { return M2 + (2 * a) * T() + t; }
  // element (a, c) of tile t as a scalar address
  __device__ __forceinline__ const double* elem(const double2* M2, int t, int a, int c) const {
    return reinterpret_cast<const double*>(M2 + (2 * a + (c >> 1)) * T() + t) + (c & 1);
  }
};

// Tiled right-looking Cholesky, in place, chunk-major layout; the forward substitution of the
// right-hand side held in registers (xr: rows gtid and gtid + GT) is fused into the sweep and its
// result y is written to the shared vector x.  Diagonal tiles end up in solve form (potrf4).
template <int W, int TT>
__device__ __forceinline__ bool chol_cm(const Group<W>& G, const CM<TT>& cm, double2* M2, int nblk, const uint16_t* tb,
                                        double (&xr)[2], double* x, double* exch) {
  constexpr int GT = Group<W>::GT;
  const int gtid = G.gtid;
  const int ntiles = (nblk * (nblk + 1)) >> 1;
  const int n4 = nblk << 2;
  const int T = cm.T();
  bool ok = true;
  int col0 = 0;  // storage index of the diagonal tile of column kb
  for (int kb = 0; kb < nblk; ++kb) {
    const int nrows = nblk - kb;
    double d[16], a[16];
    cm.ld(M2, col0, a);  // broadcast loads: every lane factors the same tile
    ok = potrf4(a, d) && ok;
    // TRSM: X = A L^-T, one panel ROW per thread (rows 4(kb+1) .. n4-1)
    for (int r = 4 * (kb + 1) + gtid; r < n4; r += GT) {
      double2* p = cm.row(M2, col0 + (r >> 2) - kb, r & 3);
      const double2 u = p[0], v = p[T];
      double x0, x1, x2, x3;
      linv4(d, u.x, u.y, v.x, v.y, x0, x1, x2, x3);
      p[0] = make_double2(x0, x1); p[T] = make_double2(x2, x3);
    }
    double b0, b1, b2, b3, y0, y1, y2, y3;
    pivot4<W>(G, xr, kb, exch, b0, b1, b2, b3);
    linv4(d, b0, b1, b2, b3, y0, y1, y2, y3);
    if (gtid == 0) {
      double2* o2 = reinterpret_cast<double2*>(x + 4 * kb);
      o2[0] = make_double2(y0, y1); o2[1] = make_double2(y2, y3);
    }
    ok = G.all(ok);  // also the barrier between the panel and the trailing update
    if (!ok) return false;
    if (gtid == 0) cm.st(M2, col0, d);
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int r = gtid + s * GT;
      if (r >= 4 * (kb + 1) && r < n4) {
        const double2* p = cm.row(M2, col0 + (r >> 2) - kb, r & 3);
        const double2 u = p[0], v = p[T];
        xr[s] -= u.x * y0 + u.y * y1 + v.x * y2 + v.y * y3;
      }
    }
    // trailing update: storage tiles of columns kb+1.. are contiguous
    const int t0 = col0 + nrows;
    for (int t = t0 + gtid; t < ntiles; t += GT) {
      const int bi = tb[t] & 0xff, bj = tb[t] >> 8;
      double li[16], lj[16], c[16];
      cm.ld(M2, col0 + bi - kb, li);
      cm.ld(M2, col0 + bj - kb, lj);
      cm.ld(M2, t, c);
#pragma unroll
      for (int rr = 0; rr < 4; ++rr)
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          double sacc = c[4 * rr + cc];
#pragma unroll
          for (int k = 0; k < 4; ++k) sacc -= li[4 * rr + k] * lj[4 * cc + k];
          c[4 * rr + cc] = sacc;
        }
      cm.st(M2, t, c);
    }
    G.sync();
    col0 = t0;
  }
  return true;
}

// Tensor-core update of NI vertically stacked 8-row blocks (tile rows r0 + 2i + {0, 1}) of the tile columns c0 (and
// c0 + 1 when TWO):  C -= sum_{kk in [k0, k1)} L(rows, kk) L(c0.., kk)'.  The blocks stay in the accumulator
// registers (two doubles per lane) over all k1 - k0 mma steps and are read and written once.  A / B fragments come
// straight out of the chunk-major tiles: lane l reads element (row (l >> 2) & 3, column l & 3) of tile t + (l >> 4);
// with T = 1 (mod 8) the 32 addresses fall into distinct banks.  A tile row past the end of the matrix reads
// whatever follows in the chunk (T > number of tiles keeps it inside the matrix region) and only feeds output rows /
// columns that are not stored.
template <int NI, int TT>
__device__ __forceinline__ void mma_update(const CM<TT>& cm, double2* M2, int nblk, int lane, bool TWO, int r0, int c0, int k0, int k1) {
  const int T = cm.T();
  const int fa = (lane >> 2) & 3, fhi = lane >> 4, ctc = (lane & 3) >> 1;
  // C fragment: row lane >> 2 of the 8, columns 2 (lane & 3) + {0, 1}: tile (row fhi, column ctc), chunk 2 fa + (lane & 1)
  const int tcol = c0 + ctc;
  const int cbase = (2 * fa + (lane & 1)) * T + tcol * nblk - ((tcol * (tcol - 1)) >> 1) - tcol;  // + tile row
  const bool colok = TWO ? tcol < nblk : ctc == 0;
  double acc[NI][2];
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int trow = r0 + 2 * i + fhi;
    acc[i][0] = 0.0; acc[i][1] = 0.0;
    if (colok && trow < nblk && trow >= tcol) { const double2 c = M2[cbase + trow]; acc[i][0] = c.x; acc[i][1] = c.y; }
  }
  // fragment element inside the tile pair (t, t + 1)
  const double* pb = reinterpret_cast<const double*>(M2) + 2 * ((2 * fa + ctc) * T + fhi) + (lane & 1)
                     + 2 * (k0 * nblk - ((k0 * (k0 - 1)) >> 1) - k0);  // + 2 (tile row): column k0
#pragma unroll 1
  for (int kk = k0; kk < k1; ++kk) {
    const double b = pb[2 * c0];
    const double bn = -b;
    double a[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) a[i] = (i == 0) ? b : pb[2 * (r0 + 2 * i)];  // r0 == c0: the first block row is the B block
#pragma unroll
    for (int i = 0; i < NI; ++i) dmma884(acc[i][0], acc[i][1], a[i], bn);
    pb += 2 * (nblk - kk - 1);
  }
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int trow = r0 + 2 * i + fhi;
    if (colok && trow < nblk && trow >= tcol) M2[cbase + trow] = make_double2(acc[i][0], acc[i][1]);
  }
}

template <int TT>
__device__ __forceinline__ void mma_update_n(int ni, const CM<TT>& cm, double2* M2, int nblk, int lane, bool TWO, int r0, int c0, int k0, int k1) {
  switch (ni) {
    case 1: mma_update<1, TT>(cm, M2, nblk, lane, TWO, r0, c0, k0, k1); break;
    case 2: mma_update<2, TT>(cm, M2, nblk, lane, TWO, r0, c0, k0, k1); break;
    case 3: mma_update<3, TT>(cm, M2, nblk, lane, TWO, r0, c0, k0, k1); break;
    case 4: mma_update<4, TT>(cm, M2, nblk, lane, TWO, r0, c0, k0, k1); break;
    case 5: mma_update<5, TT>(cm, M2, nblk, lane, TWO, r0, c0, k0, k1); break;
    case 6: mma_update<6, TT>(cm, M2, nblk, lane, TWO, r0, c0, k0, k1); break;
    case 7: mma_update<7, TT>(cm, M2, nblk, lane, TWO, r0, c0, k0, k1); break;
    case 8: mma_update<8, TT>(cm, M2, nblk, lane, TWO, r0, c0, k0, k1); break;
    default: break;
  }
}

// One-warp variant of chol_cm (n4 <= 64): LEFT-looking over 8-wide block columns, the rank-k part on the FP64 tensor
// cores.  Before the tile columns 2J, 2J + 1 are factored, the block column gets its whole update from the columns
// to its left in one go (mma_update) -- the right-looking sweep re-reads and re-writes every trailing tile at every
// step, which was 59 % of the kernel's shared-memory wavefronts.  Inside the pair of tile columns the odd column
// gets its rank-4 update from the even one the same way.  Forward substitution fused as in chol_cm.
template <int TT>
__device__ __forceinline__ bool chol_cm_mma(const Group<1>& G, const CM<TT>& cm, double2* M2, int nblk, double (&xr)[2],
                                            double* x, double* exch) {
  const int lane = G.gtid;
  const int n4 = nblk << 2;
  const int T = cm.T();
  bool ok = true;
  int col0 = 0;  // storage index of the diagonal tile of column kb
  for (int kb = 0; kb < nblk; ++kb) {
    const int nrows = nblk - kb;
    if (kb > 0) {
      mma_update_n<TT>((nrows + 1) >> 1, cm, M2, nblk, lane, !(kb & 1), kb, kb, (kb & 1) ? kb - 1 : 0, kb);
      G.sync();
    }
    double d[16], a[16];
    cm.ld(M2, col0, a);  // broadcast loads: every lane factors the same tile
    ok = potrf4(a, d) && ok;
    // TRSM: X = A L^-T, one panel ROW per thread (rows 4(kb+1) .. n4-1)
    for (int r = 4 * (kb + 1) + lane; r < n4; r += 32) {
      double2* p = cm.row(M2, col0 + (r >> 2) - kb, r & 3);
      const double2 u = p[0], v = p[T];
      double x0, x1, x2, x3;
      linv4(d, u.x, u.y, v.x, v.y, x0, x1, x2, x3);
      p[0] = make_double2(x0, x1); p[T] = make_double2(x2, x3);
    }
    double b0, b1, b2, b3, y0, y1, y2, y3;
    pivot4<1>(G, xr, kb, exch, b0, b1, b2, b3);
    linv4(d, b0, b1, b2, b3, y0, y1, y2, y3);
    if (lane == 0) {
      double2* o2 = reinterpret_cast<double2*>(x + 4 * kb);
      o2[0] = make_double2(y0, y1); o2[1] = make_double2(y2, y3);
    }
    ok = G.all(ok);  // also the barrier between the panel and what reads it
    if (!ok) return false;
    if (lane == 0) cm.st(M2, col0, d);
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int r = lane + s * 32;
      if (r >= 4 * (kb + 1) && r < n4) {
        const double2* p = cm.row(M2, col0 + (r >> 2) - kb, r & 3);
        const double2 u = p[0], v = p[T];
        xr[s] -= u.x * y0 + u.y * y1 + v.x * y2 + v.y * y3;
      }
    }
    G.sync();
    col0 += nrows;
  }
  return true;
}

// Backward substitution x = L^-T y, y read from the shared vector x, result written back to it.
template <int W, int TT>
__device__ __forceinline__ void bwd_cm(const Group<W>& G, const CM<TT>& cm, const double2* M2, int nblk, double* x, double* exch) {
  constexpr int GT = Group<W>::GT;
  const int gtid = G.gtid, n4 = nblk << 2;
  double xr[2] = {0.0, 0.0};
  int pre[2];  // tile (kb, r >> 2) = pre + kb
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int r = gtid + s * GT;
    if (r < n4) xr[s] = x[r];
    const int bj = r >> 2;
    pre[s] = bj * nblk - ((bj * (bj - 1)) >> 1) - bj;
  }
  G.sync();
  int diag = ((nblk * (nblk + 1)) >> 1) - 1;  // diagonal tile of the last column
  for (int kb = nblk - 1; kb >= 0; --kb) {
    double d[16], b0, b1, b2, b3, x0, x1, x2, x3;
    cm.ld(M2, diag, d);
    pivot4<W>(G, xr, kb, exch, b0, b1, b2, b3);
    linvt4(d, b0, b1, b2, b3, x0, x1, x2, x3);
    if (gtid == 0) {
      double2* o2 = reinterpret_cast<double2*>(x + 4 * kb);
      o2[0] = make_double2(x0, x1); o2[1] = make_double2(x2, x3);
    }
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int r = gtid + s * GT;
      if (r < 4 * kb) {  // column r & 3 of L(kb, r >> 2)
        const int t = pre[s] + kb, c = r & 3;
        xr[s] -= *cm.elem(M2, t, 0, c) * x0 + *cm.elem(M2, t, 1, c) * x1 + *cm.elem(M2, t, 2, c) * x2 + *cm.elem(M2, t, 3, c) * x3;
      }
    }
    diag -= nblk - kb + 1;  // diagonal tile of column kb - 1
  }
  G.sync();
}

__device__ __forceinline__ int ld_acquire(const int32_t* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

constexpr int kOutMap = 4;  // force-output slots per thread whose (step, leg, component) is precomputed

}  // namespace

template <int W, int TT>
__global__ void __launch_bounds__(W == 1 ? 448 : 256) cmpc_presolve_kernel(const DevConfig cfg, const SolveArgs args) {
  extern __shared__ __align__(128) double smem[];
  pdl_prologue(args.pdl_trigger != 0);
  constexpr int GT = Group<W>::GT;
  const int N = cfg.N, L = cfg.L;
  const int nf = 3 * L * N, nbfull = L * N, mfull = 5 * nbfull;
  const int nbmax = args.nbmax;
  const PrePlan& P = args.pre;
  CM<TT> cm;
  cm.Trt = P.T;
  Group<W> G;
  G.gtid = threadIdx.x % GT;
  G.gid = threadIdx.x / GT;
  const int gtid = G.gtid;
  // CTA-shared tables
  double* c_z1 = smem;            // z1[j] = sum_{k >= j} (k - j + zeta) qz_k
  double* c_z2 = c_z1 + N;        // z2[j] = sum_{k >= j} (k - j + zeta)^2 qz_k
  double* c_s1 = c_z2 + N;        // s1[j], s2[j]: the same sums with unit weights (x and y)
  double* c_s2 = c_s1 + N;
  double* c_qz = c_s2 + N;        // qz[k]: z-position weight of node k + 1 (:205)
  double* c_wf = c_qz + N;        // force-tracking weights  (CentroidalMPC.cpp:223-225)
  double* c_wr = c_wf + 3 * L;    // force-rate weights      (:227-231)
  double* base = smem + P.cta + (size_t)G.gid * P.total;
  G.red = base + P.red;
  double* s_exch = base + P.exch;
  double* s_x = base + P.x;
  double* s_ce = base + P.ce;
  int* s_misc = reinterpret_cast<int*>(base + P.ints);  // [0]=nb, [1]=invalid, [2]=work slot
  uint16_t* s_tb = reinterpret_cast<uint16_t*>(s_misc + 4);
  const int tiles_max = bc4_tiles(args.n4max);
  uint8_t* s_blk_j = reinterpret_cast<uint8_t*>(s_tb + tiles_max + (tiles_max & 1));
  uint8_t* s_blk_i = s_blk_j + nbmax;
  int8_t* s_blk_of = reinterpret_cast<int8_t*>(s_blk_i + nbmax);
  double* Mm = base + P.M;
  double2* M2 = reinterpret_cast<double2*>(Mm);
  const int ns = 9 + 3 * L, nds = 9 * (N + 1), ndi = L * (4 * N + 3);
  const int nin = (ns + nds + ndi + 1) & ~1;
  BuildView V;
  V.Mm = Mm; V.ce = s_ce; V.eq = Mm + nin; V.qz = V.eq + 9 * N; V.fz = V.qz + N; V.arm = s_x; V.g = s_x;
  V.qzt = c_qz;
  V.misc = s_misc; V.tb = s_tb; V.blk_j = s_blk_j; V.blk_i = s_blk_i; V.blk_of = s_blk_of;
  const double mass = cfg.mass, dt = cfg.dt;
  const double imass = 1.0 / mass;  // (once per thread: every per-instance use multiplies)
  const double zeta = cfg.zoh ? 0.5 : 0.0;
  // (first kernel of a call, launched without PDL: every kernel of the call before this one has completed, so its counts
  // block -- the one the NEXT call will use -- is free)
  if (args.zero_next && blockIdx.x == 0 && (int)threadIdx.x < args.zero_n) args.zero_next[threadIdx.x] = 0;
  const int count = args.count ? *args.count : args.count_imm;
  if (count <= 0) return;  // empty list (uniform over the grid): nothing to set up

  // forces leave in the reference's per-leg order [L][N][3] (CentroidalMPC.cpp:270): output t reads
  // component q of free block blk_of[j L + i]; the map is the same for every instance
  int omap[kOutMap];
#pragma unroll
  for (int k = 0; k < kOutMap; ++k) {
    const int t = gtid + k * GT;
    const int i = t / (3 * N), rem = t - i * 3 * N, j = rem / 3;
    omap[k] = ((j * L + i) << 2) | (rem - 3 * j);
  }
  if ((int)threadIdx.x < N) {
    const int j = threadIdx.x;
    double z1 = 0.0, z2 = 0.0;
    for (int k = j; k < N; ++k) {
      const double om = (cfg.w[2] * 0.5) * exp(-(double)(k + 1)) + cfg.w[2] * 0.5;  // node k + 1, :205
      const double al = (double)(k - j) + zeta;
      z1 += al * om * om; z2 += al * al * om * om;
    }
    c_z1[j] = z1; c_z2[j] = z2;
    { const double om = (cfg.w[2] * 0.5) * exp(-(double)(j + 1)) + cfg.w[2] * 0.5; c_qz[j] = om * om; }
    const double cnt = (double)(N - j);
    c_s1[j] = 0.5 * cnt * (cnt - 1.0) + zeta * cnt;
    c_s2[j] = (cnt - 1.0) * cnt * (2.0 * cnt - 1.0) * (1.0 / 6.0) + zeta * cnt * (cnt - 1.0) + zeta * zeta * cnt;
  }
  if ((int)threadIdx.x < 3 * L) {
    c_wf[threadIdx.x] = cfg.w[9 + 3 * L + threadIdx.x];
    c_wr[threadIdx.x] = cfg.w[9 + 6 * L + threadIdx.x];
  }
  __syncthreads();

  while (true) {
    int slot = 0;
    if (gtid == 0) slot = atomicAdd(args.work, 1);
    slot = G.bcast0(slot, s_misc + 2);
    if (slot >= count) break;
    const int inst = args.perm ? args.perm[slot] : slot;
    if (args.ready) {
      // progressive arrival: the copy stream bumps *ready after each chunk of inputs has landed
      const int chunk = inst / args.ready_chunk;
      if (gtid == 0) {
        const long long t0 = clock64();
        while (ld_acquire(args.ready) <= chunk) {
          __nanosleep(1000);
          if (clock64() - t0 > 6000000000LL) { atomicExch(args.error_flag, 1); break; }  // ~3 s: never hang the GPU
        }
      }
      G.sync();
    }
    bool defer = false;
    int nb = 0, n = 0, nblk = 0, n4 = 0;
    double xr[2] = {0.0, 0.0};
    double arm[3] = {0.0, 0.0, 0.0}, gb[3] = {0.0, 0.0, 0.0};  // lever arm and gradient g of block gtid (the verification needs them again)
    {
      const bool finite = stage_inputs<W>(G, cfg, args, inst, V);
      nb = s_misc[0];
      n = 3 * nb; nblk = (n + 3) >> 2; n4 = nblk << 2;
      const bool invalid = s_misc[1] != 0;
      if (args.route && s_misc[3] > nbmax) {  // belongs to a larger size class: forward it
        if (gtid == 0) {
          const int tot = s_misc[3];
          const int c = tot <= args.route_b1 ? 1 : (tot <= args.route_b2 ? 2 : 3);
          args.route_perm[(size_t)c * args.route_stride + atomicAdd(args.route_counts + c, 1)] = inst;
        }
        G.sync();
        continue;
      }
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
      if (args.warm_active) {  // a warm-start guess with active rows belongs to the IPM kernel's polish
        bool any = false;
        const uint16_t* wa = args.warm_active + (size_t)inst * nbfull;
        for (int t = gtid; t < nbfull; t += GT) { const unsigned a = wa[t]; any = any || (!(a & 0x8000u) && (a & 0x3ffu)); }
        defer = !G.all(!any);
      }
    }
    if (!defer) {
      // ---- g = 2 Bqp' L (Aqp x0 + dqp - Xref) - 2 W_f Uref: thread b owns block b (nb <= GT), adjoint
      // sums over the staged errors; the lever arm stays in registers until the staged inputs are dead
      // suffix sums of the staged errors, in place: eq[9k + q] <- sum_{k' >= k} eq[9k' + q] for the velocity and
      // angular-momentum rows, sum_{k' >= k} (k' - k + zeta) eq[9k' + q] for the position rows (lane q, serial in k)
      if (gtid < 9) {
        double S = 0.0, Pw = 0.0;
        for (int k = N - 1; k >= 0; --k) {
          const double e = V.eq[9 * k + gtid];
          Pw += S + zeta * e;
          S += e;
          V.eq[9 * k + gtid] = gtid < 3 ? Pw : S;
        }
      }
      G.sync();
      if (gtid < nb) {
        const int b = gtid, j = s_blk_j[b], i = s_blk_i[b];
        const double ce = s_ce[b];
        for (int q = 0; q < 3; ++q) arm[q] = Mm[ns + nds + i * (4 * N + 3) + N + 3 * j + q] - Mm[ns + 3 * j + q];
        const double* e9 = V.eq + 9 * j;
        const double sl3[3] = {e9[6], e9[7], e9[8]};
        const double cmass = ce * imass;
        const double cr[3] = {sl3[1] * arm[2] - sl3[2] * arm[1], sl3[2] * arm[0] - sl3[0] * arm[2], sl3[0] * arm[1] - sl3[1] * arm[0]};
        for (int q = 0; q < 3; ++q) {
          double gq = 2.0 * (cmass * (dt * dt * e9[q] + dt * e9[3 + q]) + dt * ce * cr[q]);
          if (q == 2) gq -= 2.0 * c_wf[3 * i + 2] * V.fz[b];
          s_x[3 * b + q] = gq;
          gb[q] = gq;
        }
      }
      if (gtid < n4 - n) s_x[n + gtid] = 0.0;
      if constexpr (W > 1) {  // tile table of the right-looking sweep (the one-warp sweep does not use it)
        for (int bj = gtid; bj < nblk; bj += GT) {
          const int o = blkoff(bj, bj, nblk);
          for (int bi = bj; bi < nblk; ++bi) s_tb[o + bi - bj] = (uint16_t)(bi | (bj << 8));
        }
      }
      G.sync();
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int r = gtid + s * GT;
        if (r < n4) xr[s] = -s_x[r];
      }
      G.sync();
      if (gtid < nb) { s_x[3 * gtid] = arm[0]; s_x[3 * gtid + 1] = arm[1]; s_x[3 * gtid + 2] = arm[2]; }
      G.sync();  // every read of the staged inputs is done: the tiles may overwrite them

      // ---- H = 2 (Bqp' L Bqp + K), one thread per block pair (b >= b2, hence j >= j2), every lane the
      // same instruction stream.  With d = j - j2, cnt = N - j (row blocks k >= j contribute):
      //   angular   cnt dt^2 c c2 [r]x' diag(w6..8) [r2]x
      //   diagonal  (c/m)(c2/m) (dt^4 P_a + cnt dt^2 w[3+a]),  P_a = w[a] (S2 + d S1) for x, y and
      //             z2[j] + d z1[j] for z;  S1, S2 = sums of (t + zeta), (t + zeta)^2 over t < cnt
      //   same leg  K = W_f + D' W_r D (CentroidalMPC.cpp:223-231)
      // Element (gi, gj) lives at double index R(gi) + C(gj) of the chunk-major layout:
      //   R = 4 (gi & 3) T + 2 (gi >> 2),  C = 2 T ((gj & 3) >> 1) + 2 colbase(gj >> 2) + (gj & 1).
      {
        const int T = cm.T();
        const double dt2 = dt * dt, dt4 = dt2 * dt2;
        const double q0 = cfg.w[6], q1 = cfg.w[7], q2 = cfg.w[8];
        const double im2 = 1.0 / (mass * mass);
        auto Rof = [&](int g) { return 4 * (g & 3) * T + 2 * (g >> 2); };
        auto Cof = [&](int g) { const int tj = g >> 2; return 2 * T * ((g & 3) >> 1) + 2 * (tj * nblk - ((tj * (tj + 1)) >> 1)) + (g & 1); };
        // One block pair.  FAR (b2 <= b - 2): all nine elements lie strictly below the diagonal tiles,
        // separable addressing.  Near pairs (b2 = b or b - 1) may touch a diagonal tile: only its lower triangle is
        // ever read (potrf4; the tensor-core update carries the upper one along without looking at it).
        // One loop over all pairs, far ones first (one copy of the block arithmetic: the kernel's working set of
        // instructions matters more than the one round in which a warp runs both store paths).
        auto do_pair = [&](int b, int b2, bool FAR) {
          const int j = s_blk_j[b], i = s_blk_i[b], j2 = s_blk_j[b2], i2 = s_blk_i[b2];
          const double ce = s_ce[b], ce2 = s_ce[b2];
          const double r0 = s_x[3 * b], r1 = s_x[3 * b + 1], r2 = s_x[3 * b + 2];
          const double p0 = s_x[3 * b2], p1 = s_x[3 * b2 + 1], p2 = s_x[3 * b2 + 2];
          const double cnt = (double)(N - j), dd = (double)(j - j2);
          const double s0 = c_s2[j] + dd * c_s1[j], sz = c_z2[j] + dd * c_z1[j];
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
              const double wr = c_wr[3 * i + aa];
              blk[aa][aa] += (j == j2) ? c_wf[3 * i + aa] + nn * wr : -wr;
            }
          }
          const int g0 = 3 * b, h0 = 3 * b2;
          if (FAR) {
            int R[3], C[3];
#pragma unroll
            for (int aa = 0; aa < 3; ++aa) { R[aa] = Rof(g0 + aa); C[aa] = Cof(h0 + aa); }
#pragma unroll
            for (int aa = 0; aa < 3; ++aa)
#pragma unroll
              for (int bb = 0; bb < 3; ++bb) Mm[R[aa] + C[bb]] = 2.0 * blk[aa][bb];
          } else {
#pragma unroll
            for (int aa = 0; aa < 3; ++aa)
#pragma unroll
              for (int bb = 0; bb < 3; ++bb) {
                const int gi = g0 + aa, gj = h0 + bb;
                const double v = 2.0 * blk[aa][bb];
                if ((gi >> 2) >= (gj >> 2)) Mm[Rof(gi) + Cof(gj)] = v;
              }
          }
        };
        const int nfar = nb >= 3 ? ((nb - 1) * (nb - 2)) >> 1 : 0;
        {  // far pair idx = a (a + 1) / 2 + rem, 0 <= rem <= a: block row a + 2, block column rem (decoded incrementally);
           // then the 2 nb - 1 near pairs
          int a = 0, rem = gtid;
          const int npairs = nfar + 2 * nb - 1;
          for (int idx = gtid; idx < npairs; idx += GT) {
            int b, b2;
            const bool far = idx < nfar;
            if (far) {
              while (rem > a) { rem -= a + 1; ++a; }
              b = a + 2; b2 = rem;
              rem += GT;
            } else {
              const int e = idx - nfar;
              b = (e + 1) >> 1; b2 = b - (e & 1);
            }
            do_pair(b, b2, far);
          }
        }
        // padding rows (n .. n4-1): identity
        for (int e = gtid; e < (n4 - n) * n4; e += GT) {
          const int gi = n + e / n4, gj = e % n4;
          if ((gi >> 2) >= (gj >> 2)) Mm[Rof(gi) + Cof(gj)] = gi == gj ? 1.0 : 0.0;
        }
      }
      G.sync();
      bool ok;
      if constexpr (W == 1) ok = chol_cm_mma<TT>(G, cm, M2, nblk, xr, s_x, s_exch);
      else ok = chol_cm<W, TT>(G, cm, M2, nblk, s_tb, xr, s_x, s_exch);
      if (ok) {
        bwd_cm<W, TT>(G, cm, M2, nblk, s_x, s_exch);
      }
      defer = !ok;
    }
    double gs = 1.0, usf = 1.0, stat = 0.0, prim = 0.0;
    if (!defer) {
      // Stationarity H u + g on the problem itself, not on the factored matrix: H u = 2 (B' L B u + K u) by one roll-out of
      // the forces through the dynamics (CentroidalMPC.cpp:85-92, frozen arms), the stage weights (:203-221), and the
      // adjoint sums the gradient g was built with -- O(N) work on data that is still on chip, where a product with H
      // needs H a second time (it used to be parked in L2 across the factorisation: 13 % of the kernel's time with the
      // copy back).  Scratch: the matrix region, dead after the back substitution.
      double* vC = Mm;                // [nb][6]  wrench of block b: ce u, ce arm x u
      double* vW = vC + 6 * nbmax;    // [N][6]   wrench of step k
      double* vE = vW + 6 * N;        // [N][9]   weighted state deviation at node k + 1, then its suffix sums
      double ub[3] = {0.0, 0.0, 0.0};
      if (gtid < nb) {
        const int b = gtid;
        const double ce = s_ce[b];
        for (int q = 0; q < 3; ++q) ub[q] = s_x[3 * b + q];
        vC[6 * b] = ce * ub[0]; vC[6 * b + 1] = ce * ub[1]; vC[6 * b + 2] = ce * ub[2];
        vC[6 * b + 3] = ce * (arm[1] * ub[2] - arm[2] * ub[1]);
        vC[6 * b + 4] = ce * (arm[2] * ub[0] - arm[0] * ub[2]);
        vC[6 * b + 5] = ce * (arm[0] * ub[1] - arm[1] * ub[0]);
      }
      G.sync();
      if (gtid < N) {
        double a6[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        for (int i = 0; i < L; ++i) {
          const int b = s_blk_of[gtid * L + i];
          if (b >= 0)
            for (int q = 0; q < 6; ++q) a6[q] += vC[6 * b + q];
        }
        for (int q = 0; q < 6; ++q) vW[6 * gtid + q] = a6[q];
      }
      G.sync();
      if (gtid < 9) {  // one state row per lane: roll-out, weights, suffix sums (position rows carry their own velocity)
        const int grp = gtid / 3, q = gtid - 3 * grp;
        const double kb = grp == 2 ? dt : dt * imass, zc = zeta * dt * dt * imass;
        const double wr9 = cfg.w[gtid];
        const double* src = vW + (grp == 2 ? 3 + q : q);
        double a = 0.0, bq = 0.0;
        for (int k = 0; k < N; ++k) {
          const double X = src[6 * k];
          a = fma(dt, bq, a) + zc * X;
          bq = fma(kb, X, bq);
          vE[9 * k + gtid] = (gtid == 2 ? c_qz[k] : wr9) * (grp == 0 ? a : bq);
        }
        double S = 0.0, Pw = 0.0;
        for (int k = N - 1; k >= 0; --k) {
          const double e = vE[9 * k + gtid];
          Pw += S + zeta * e; S += e;
          vE[9 * k + gtid] = grp == 0 ? Pw : S;
        }
      }
      G.sync();
      double gmax = 0.0, umax = 0.0;
      bool fin = true;
      if (gtid < nb) {
        const int b = gtid, j = s_blk_j[b], i = s_blk_i[b];
        const double ce = s_ce[b], cmass = ce * imass;
        const double* e9 = vE + 9 * j;
        const double sl3[3] = {e9[6], e9[7], e9[8]};
        const double cr[3] = {sl3[1] * arm[2] - sl3[2] * arm[1], sl3[2] * arm[0] - sl3[0] * arm[2], sl3[0] * arm[1] - sl3[1] * arm[0]};
        const int bp = j > 0 ? s_blk_of[(j - 1) * L + i] : -1, bn = j + 1 < N ? s_blk_of[(j + 1) * L + i] : -1;
        const double nn = (j > 0 ? 1.0 : 0.0) + (j + 1 < N ? 1.0 : 0.0);
        for (int q = 0; q < 3; ++q) {
          const double up = bp >= 0 ? s_x[3 * bp + q] : 0.0, un = bn >= 0 ? s_x[3 * bn + q] : 0.0;
          const double ku = c_wf[3 * i + q] * ub[q] + c_wr[3 * i + q] * (nn * ub[q] - up - un);
          const double hu = 2.0 * (cmass * (dt * dt * e9[q] + dt * e9[3 + q]) + dt * ce * cr[q] + ku);
          gmax = fmax(gmax, fabs(gb[q])); umax = fmax(umax, fabs(ub[q]));
          stat = fmax(stat, fabs(hu + gb[q]));
          fin = fin && isfinite(ub[q]) && isfinite(hu);
        }
      }
      if (gtid < nb) {
        const int b = gtid;
        const double ce = s_ce[b];
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
    {
      double* fo = args.forces + (size_t)inst * nf;
#pragma unroll
      for (int k = 0; k < kOutMap; ++k) {
        const int t = gtid + k * GT;
        if (t < nf) { const int b = s_blk_of[omap[k] >> 2]; fo[t] = b < 0 ? 0.0 : s_x[3 * b + (omap[k] & 3)]; }
      }
      for (int t = gtid + kOutMap * GT; t < nf; t += GT) {
        const int i = t / (3 * N), j = (t % (3 * N)) / 3, q = t % 3;
        const int b = s_blk_of[j * L + i];
        fo[t] = b < 0 ? 0.0 : s_x[3 * b + q];
      }
    }
    if (args.lam) for (int t = gtid; t < 2 * mfull; t += GT) args.lam[(size_t)inst * 2 * mfull + t] = 0.0;
    if (args.active) {
      for (int t = gtid; t < nbfull; t += GT) {
        const int b = s_blk_of[t];
        uint16_t a = 0x8000;
        if (b >= 0) {
          const double ce = s_ce[b];
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
      if (args.kkt) args.kkt[inst] = fmax(stat * fast_rcp(gs), fmax(prim, 0.0) * fast_rcp(usf));
    }
    G.sync();
  }
}

namespace {
template <int W, int TT>
cudaError_t launch_t(int grid, int block, size_t smem, cudaStream_t stream, const DevConfig& cfg, const SolveArgs& args) {
  return launch_ex(cmpc_presolve_kernel<W, TT>, grid, block, smem, stream, args.pdl != 0, cfg, args);
}
}  // namespace

// The headline class (n <= 60, one warp per instance) has 120 tiles: T = 121 is compiled in so that
// chunk addresses are immediates; every other class takes T from the plan.
cudaError_t launch_presolve_kernel(int W, int grid, int block, size_t smem, cudaStream_t stream, const DevConfig& cfg,
                                   const SolveArgs& args) {
  switch (W) {
    case 1: return args.pre.T == 121 ? launch_t<1, 121>(grid, block, smem, stream, cfg, args) : launch_t<1, 0>(grid, block, smem, stream, cfg, args);
    case 4: return launch_t<4, 0>(grid, block, smem, stream, cfg, args);
    case 8: return launch_t<8, 0>(grid, block, smem, stream, cfg, args);
  }
  return cudaErrorInvalidValue;
}

cudaError_t set_presolve_kernel_smem(int W, size_t bytes) {
  cudaError_t e = cudaErrorInvalidValue;
  switch (W) {
    case 1:
      e = cudaFuncSetAttribute(cmpc_presolve_kernel<1, 121>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(cmpc_presolve_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
      break;
    case 4: e = cudaFuncSetAttribute(cmpc_presolve_kernel<4, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes); break;
    case 8: e = cudaFuncSetAttribute(cmpc_presolve_kernel<8, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes); break;
  }
  return e;
}

}  // namespace cmpc
