// fwd_fast.cu -- forward solve for dense A with polyhedral cones (zero + nonneg rows): the same operator
// splitting as fwd.cu (same formulas, same termination rules; SURVEY.md 8a F3-F6, the work diffcp/SCS do
// at src/cvxpylayers/interfaces/diffcp_if.py:365,369), re-laid out around what the shared-memory
// micro-benchmarks showed (tools/microbench.cu, profiles/README.md): the per-iteration products with A were
// bound by shared-memory bandwidth and shuffle throughput, not by FP64 issue.
//
//   * A LIVES IN REGISTERS: thread (R, C) of the 512 keeps the 4 x 10 tile A[4R..4R+3, 10C..10C+9] of the
//     equilibrated matrix for the whole solve (80 of its 128 registers; m n <= 20480).  A x is 40 FMAs per
//     thread + a 10-term partial sum per row through shared memory, A' y is 40 FMAs + a (m/4)-term partial
//     sum per column; neither touches the 160 KB of A again and neither uses a shuffle.
//   * K^{-1} IS EXPLICIT: after the packed Cholesky + inverse (common.cuh) the symmetric n x n inverse
//     Linv' Linv is formed once into the shared memory that staged A, so the linear solve of an iteration is
//     ONE dense product (2 x 10 tiles, same partial-sum scheme) instead of two packed triangular ones.
//   * Ruiz passes run on the register tiles too (row / column maxima through the same partial buffers);
//     P sits in shared memory as a packed symmetric matrix during equilibration and K formation.
//   * The four dot products of the tau root are reduced only over the warps that own an output.
// Everything else (metric, tau root, over-relaxation, adaptive checks, certificates, adaptive scale with
// on-chip re-factorisation, write-back) is the algorithm of fwd.cu verbatim.
#include "common.cuh"

namespace {

constexpr int FT = 512;          // threads per CTA
constexpr int TR = 4, TC = 10;   // register tile of A

struct FastGeom { int CT, RTu, npad, mpad, KR, npk, XD; };

__host__ __device__ inline bool fwdf_geom(int n, int m, FastGeom &g) {
  g.CT = (n + TC - 1) / TC; g.RTu = (m + TR - 1) / TR; g.npad = g.CT * TC; g.mpad = g.RTu * TR;
  if ((long long)g.CT * g.RTu > FT || n > FT || m > FT) return false;
  g.KR = (((n + 1) / 2) * g.CT <= FT) ? 2 : 4;
  if (((n + g.KR - 1) / g.KR) * g.CT > FT) return false;
  g.npk = (n * (n + 1) / 2 + 1) & ~1;
  const int rowsR = g.mpad > n + 3 ? g.mpad : n + 3;
  const long long iter = (long long)n * g.npad + (long long)g.RTu * g.npad + (long long)rowsR * g.CT;   // Kinv | column partials | row partials
  const long long stage = (long long)m * n + 2;                                                          // A as delivered by TMA / staged for K
  long long xd = iter > stage ? iter : stage;
  if (xd < FT) xd = FT;                                                                                  // P_mul scratch
  g.XD = (int)((xd + 1) & ~1LL);
  return true;
}
__host__ __device__ inline size_t fwdf_smem_doubles(int n, int m) {
  FastGeom g;
  if (!fwdf_geom(n, m, g)) return (size_t)1 << 40;
  return 4 + (size_t)g.XD + g.npk + 9 * (size_t)g.npad + 7 * (size_t)g.mpad + 8 * 32;
}

struct FSmem {   // few base pointers, the vectors are addressed as base + k * stride (registers are for the A tile)
  uint64_t *bar; int *ibuf;
  double *X, *Li, *vx, *vy, *red;
  int npad, mpad, oXC, oXR;
  __device__ __forceinline__ double *wx() const { return vx; }
  __device__ __forceinline__ double *ux() const { return vx + npad; }
  __device__ __forceinline__ double *utx() const { return vx + 2 * npad; }
  __device__ __forceinline__ double *gx() const { return vx + 3 * npad; }
  __device__ __forceinline__ double *ch() const { return vx + 4 * npad; }
  __device__ __forceinline__ double *En() const { return vx + 5 * npad; }
  __device__ __forceinline__ double *tn() const { return vx + 6 * npad; }
  __device__ __forceinline__ double *tn2() const { return vx + 7 * npad; }
  __device__ __forceinline__ double *tn3() const { return vx + 8 * npad; }
  __device__ __forceinline__ double *wy() const { return vy; }
  __device__ __forceinline__ double *uy() const { return vy + mpad; }
  __device__ __forceinline__ double *uty() const { return vy + 2 * mpad; }
  __device__ __forceinline__ double *gy() const { return vy + 3 * mpad; }
  __device__ __forceinline__ double *bh() const { return vy + 4 * mpad; }
  __device__ __forceinline__ double *Dm() const { return vy + 5 * mpad; }
  __device__ __forceinline__ double *tm() const { return vy + 6 * mpad; }
  __device__ __forceinline__ double *Kinv() const { return X; }    // views into X during the iterations
  __device__ __forceinline__ double *XC() const { return X + oXC; }
  __device__ __forceinline__ double *XR() const { return X + oXR; }
};

__device__ __forceinline__ double inv_ry_f(int z, int i, double scale) { return i < z ? BC_ZERO_CONE_FACTOR * scale : scale; }
__device__ __forceinline__ double dmax(double a, double b) { return a > b ? a : b; }

// ---------------------------------------------------------------- register-tile products
// out_j = sum_i A_ij y_i.  Every active thread folds its 4 rows into 10 column partials, thread j < n adds
// the RTu partials of its column.  ep(j, value) runs on thread j.  One barrier inside, none at the end.
template <class Epi>
__device__ __forceinline__ void rt_cols(const double (&ar)[TR][TC], const FastGeom &g, bool act, int R, int C, const double *y,
                                        double *XC, int n, Epi ep) {
  if (act) {
    const double2 y01 = *reinterpret_cast<const double2 *>(y + TR * R), y23 = *reinterpret_cast<const double2 *>(y + TR * R + 2);
    double2 *dst = reinterpret_cast<double2 *>(XC + R * g.npad + TC * C);
#pragma unroll
    for (int c = 0; c < TC; c += 2) {
      const double q0 = fma(ar[3][c], y23.y, fma(ar[2][c], y23.x, fma(ar[1][c], y01.y, ar[0][c] * y01.x)));
      const double q1 = fma(ar[3][c + 1], y23.y, fma(ar[2][c + 1], y23.x, fma(ar[1][c + 1], y01.y, ar[0][c + 1] * y01.x)));
      dst[c >> 1] = make_double2(q0, q1);
    }
  }
  __syncthreads();
  const int t = threadIdx.x;
  if (t < n) {
    const double *p = XC + t;
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    int Rr = 0;
    for (; Rr + 3 < g.RTu; Rr += 4) {
      s0 += p[Rr * g.npad]; s1 += p[(Rr + 1) * g.npad]; s2 += p[(Rr + 2) * g.npad]; s3 += p[(Rr + 3) * g.npad];
    }
    for (; Rr < g.RTu; Rr++) s0 += p[Rr * g.npad];
    ep(t, (s0 + s1) + (s2 + s3));
  }
}
// out_i = sum_j A_ij x_j.  ep(i, value) runs on thread i < m.  One barrier inside, none at the end.
template <class Epi>
__device__ __forceinline__ void rt_rows(const double (&ar)[TR][TC], const FastGeom &g, bool act, int R, int C, const double *x,
                                        double *XR, int m, Epi ep) {
  if (act) {
    double xv[TC];
#pragma unroll
    for (int c = 0; c < TC; c += 2) { const double2 v = *reinterpret_cast<const double2 *>(x + TC * C + c); xv[c] = v.x; xv[c + 1] = v.y; }
#pragma unroll
    for (int r = 0; r < TR; r++) {
      double s0 = 0, s1 = 0;
#pragma unroll
      for (int c = 0; c < TC; c += 2) { s0 = fma(ar[r][c], xv[c], s0); s1 = fma(ar[r][c + 1], xv[c + 1], s1); }
      XR[(TR * R + r) * g.CT + C] = s0 + s1;
    }
  }
  __syncthreads();
  const int t = threadIdx.x;
  if (t < m) {
    const double *p = XR + t * g.CT;
    double s0 = 0, s1 = 0;
    int c = 0;
    for (; c + 1 < g.CT; c += 2) { s0 += p[c]; s1 += p[c + 1]; }
    if (c < g.CT) s0 += p[c];
    ep(t, s0 + s1);
  }
}
// out_i = sum_j Kinv_ij x_j for the symmetric n x n inverse stored with row stride npad (KR x 10 tiles read from
// shared memory).  ep(i, value) runs on thread i < n.  One barrier inside, none at the end.
template <int KR, class Epi>
__device__ __forceinline__ void kinv_rows(const double *Kinv, const FastGeom &g, int n, const double *x, double *XR, Epi ep) {
  const int t = threadIdx.x, R = t / g.CT, C = t - R * g.CT;
  if (KR * R < n) {
    double xv[TC];
#pragma unroll
    for (int c = 0; c < TC; c += 2) { const double2 v = *reinterpret_cast<const double2 *>(x + TC * C + c); xv[c] = v.x; xv[c + 1] = v.y; }
#pragma unroll
    for (int r = 0; r < KR; r++) {
      const int i = KR * R + r;
      if (i < n) {
        const double2 *row = reinterpret_cast<const double2 *>(Kinv + i * g.npad + TC * C);
        double s0 = 0, s1 = 0;
#pragma unroll
        for (int c = 0; c < TC; c += 2) { const double2 q = row[c >> 1]; s0 = fma(q.x, xv[c], s0); s1 = fma(q.y, xv[c + 1], s1); }
        XR[i * g.CT + C] = s0 + s1;
      }
    }
  }
  __syncthreads();
  if (t < n) {
    const double *p = XR + t * g.CT;
    double s0 = 0, s1 = 0;
    int c = 0;
    for (; c + 1 < g.CT; c += 2) { s0 += p[c]; s1 += p[c + 1]; }
    if (c < g.CT) s0 += p[c];
    ep(t, s0 + s1);
  }
}
template <class Epi>
__device__ __forceinline__ void kinv_mul(const double *Kinv, const FastGeom &g, int n, const double *x, double *XR, Epi ep) {
  if (g.KR == 2) kinv_rows<2>(Kinv, g, n, x, XR, ep); else kinv_rows<4>(Kinv, g, n, x, XR, ep);
}

// Sum of four per-thread values over the block when only the first `nwc` warps hold non-zero terms.
// Two barriers; every thread ends with the same bits.
__device__ __forceinline__ void reduce4_lead(double (&v)[4], double *red, int nwc) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp < nwc) {
    const double k = butterfly4(v[0], v[1], v[2], v[3], lane);   // lane 8 q holds the warp sum of value order[q]
    if ((lane & 7) == 0) red[(((lane >> 4) & 1) * 2 + ((lane >> 3) & 1)) * 16 + warp] = k;
  }
  __syncthreads();
  // second stage: lane = 16 * (value >> 1) ... keep it simple: lanes 0..15 sum values 0/1, all lanes read two partials
  {
    const int w = lane & 15;
    double a0 = w < nwc ? red[(lane >> 4) * 16 + w] : 0.0;          // value 0 (lanes 0-15) / value 1 (lanes 16-31)
    double a1 = w < nwc ? red[(2 + (lane >> 4)) * 16 + w] : 0.0;    // value 2 / value 3
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) { a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o); }
    v[0] = __shfl_sync(0xffffffffu, a0, 0); v[1] = __shfl_sync(0xffffffffu, a0, 16);
    v[2] = __shfl_sync(0xffffffffu, a1, 0); v[3] = __shfl_sync(0xffffffffu, a1, 16);
  }
  __syncthreads();   // red may be rewritten by the next reduction
}

// K = rho_x I + scale * sum_i w_i a_i a_i' (+ P^ already sitting in K as unscaled packed P when Psm) for the
// staged, equilibrated A (row-major m x n in shared memory), packed lower.
__device__ void form_K(const double *Av, int m, int n, int z, double scale, double rho_x, double *K, bool haveP, const double *En) {
  const int T = blockDim.x, t = threadIdx.x;
  if ((n & 1) == 0) {
    const int nb = n >> 1, ntile = (nb * (nb + 1)) >> 1;
    for (int e = t; e < ntile; e += T) {
      int J = (int)((sqrtf(8.0f * e + 1.0f) - 1.0f) * 0.5f);
      while (((J + 1) * (J + 2)) >> 1 <= e) J++;
      while ((J * (J + 1)) >> 1 > e) J--;
      const int Kb = e - ((J * (J + 1)) >> 1);
      const double2 *pj = reinterpret_cast<const double2 *>(Av) + J, *pk = reinterpret_cast<const double2 *>(Av) + Kb;
      double z00 = 0, z01 = 0, z10 = 0, z11 = 0, s00 = 0, s01 = 0, s10 = 0, s11 = 0;
      int i = 0;
      for (; i < z; i++) { const double2 u = pj[i * nb], v = pk[i * nb]; z00 = fma(u.x, v.x, z00); z01 = fma(u.x, v.y, z01); z10 = fma(u.y, v.x, z10); z11 = fma(u.y, v.y, z11); }
      for (; i < m; i++) { const double2 u = pj[i * nb], v = pk[i * nb]; s00 = fma(u.x, v.x, s00); s01 = fma(u.x, v.y, s01); s10 = fma(u.y, v.x, s10); s11 = fma(u.y, v.y, s11); }
      const int j0 = 2 * J, k0 = 2 * Kb;
      const int e00 = ((j0 * (j0 + 1)) >> 1) + k0, e10 = (((j0 + 1) * (j0 + 2)) >> 1) + k0;
      double v00 = (z00 * BC_ZERO_CONE_FACTOR + s00) * scale + (j0 == k0 ? rho_x : 0.0);
      double v01 = (z01 * BC_ZERO_CONE_FACTOR + s01) * scale;
      double v10 = (z10 * BC_ZERO_CONE_FACTOR + s10) * scale;
      double v11 = (z11 * BC_ZERO_CONE_FACTOR + s11) * scale + (j0 == k0 ? rho_x : 0.0);
      if (haveP) {
        v00 += K[e00] * En[k0] * En[j0];
        if (k0 + 1 <= j0) v01 += K[e00 + 1] * En[k0 + 1] * En[j0];
        v10 += K[e10] * En[k0] * En[j0 + 1];
        v11 += K[e10 + 1] * En[k0 + 1] * En[j0 + 1];
      }
      K[e00] = v00;
      if (k0 + 1 <= j0) K[e00 + 1] = v01;
      K[e10] = v10; K[e10 + 1] = v11;
    }
  } else {
    const int npk = n * (n + 1) / 2;
    for (int e = t; e < npk; e += T) {
      int j = (int)((sqrtf(8.0f * e + 1.0f) - 1.0f) * 0.5f);
      while ((j + 1) * (j + 2) / 2 <= e) j++;
      while (j * (j + 1) / 2 > e) j--;
      const int k = e - j * (j + 1) / 2;
      double acc0 = 0, acc1 = 0;
      const double *cj = Av + j, *ck = Av + k;
      int i = 0;
      for (; i < z; i++) acc0 = fma(cj[i * n], ck[i * n], acc0);
      for (; i < m; i++) acc1 = fma(cj[i * n], ck[i * n], acc1);
      double v = (acc0 * BC_ZERO_CONE_FACTOR + acc1) * scale + (j == k ? rho_x : 0.0);
      if (haveP) v += K[e] * En[k] * En[j];
      K[e] = v;
    }
  }
  __syncthreads();
}

// Kinv = X' X for the packed lower-triangular X = L^{-1}: full symmetric n x n with row stride npad.
// 2 x 2 tiles of the lower triangle; both triangles are written.
__device__ void form_Kinv(const double *Xp, int n, int npad, double *Kinv) {
  const int T = blockDim.x, t = threadIdx.x;
  const int nb = (n + 1) >> 1, ntile = (nb * (nb + 1)) >> 1;
  for (int e = t; e < ntile; e += T) {
    int I = (int)((sqrtf(8.0f * e + 1.0f) - 1.0f) * 0.5f);
    while (((I + 1) * (I + 2)) >> 1 <= e) I++;
    while ((I * (I + 1)) >> 1 > e) I--;
    const int J = e - ((I * (I + 1)) >> 1);
    const int i0 = 2 * I, j0 = 2 * J;
    double c00 = 0, c01 = 0, c10 = 0, c11 = 0;
    {   // k = i0: X[k][i0 + 1] is above the diagonal
      const double *row = Xp + ((i0 * (i0 + 1)) >> 1);
      const double a0 = row[i0], b0 = row[j0], b1 = (j0 + 1 <= i0) ? row[j0 + 1] : 0.0;
      c00 = a0 * b0; c01 = a0 * b1;
    }
    for (int k = i0 + 1; k < n; k++) {
      const double *row = Xp + ((k * (k + 1)) >> 1);
      const double a0 = row[i0], a1 = row[i0 + 1], b0 = row[j0], b1 = row[j0 + 1];
      c00 = fma(a0, b0, c00); c01 = fma(a0, b1, c01); c10 = fma(a1, b0, c10); c11 = fma(a1, b1, c11);
    }
    const bool i1 = i0 + 1 < n, j1 = j0 + 1 < n;
    Kinv[i0 * npad + j0] = c00; Kinv[j0 * npad + i0] = c00;
    if (j1) { Kinv[i0 * npad + j0 + 1] = c01; Kinv[(j0 + 1) * npad + i0] = c01; }
    if (i1) { Kinv[(i0 + 1) * npad + j0] = c10; Kinv[j0 * npad + i0 + 1] = c10; }
    if (i1 && j1) { Kinv[(i0 + 1) * npad + j0 + 1] = c11; Kinv[(j0 + 1) * npad + i0 + 1] = c11; }
  }
  // padding columns [n, npad) must stay finite: the tile products multiply them by zeros of the vectors
  for (int k = t; k < n * (npad - n); k += T) { const int i = k / (npad - n), c = n + k % (npad - n); Kinv[i * npad + c] = 0.0; }
  __syncthreads();
}

__device__ __forceinline__ void carve_fast(FSmem &M, double *base, const FastGeom &g, int n) {
  double *q = base;
  M.bar = (uint64_t *)q; q += 2;
  M.ibuf = (int *)q; q += 2;
  M.X = q; q += g.XD;
  M.Li = q; q += g.npk;
  M.vx = q; q += 9 * g.npad;
  M.vy = q; q += 7 * g.mpad;
  M.red = q;
  M.npad = g.npad; M.mpad = g.mpad;
  M.oXC = n * g.npad; M.oXR = M.oXC + g.RTu * g.npad;
}

}  // namespace

__global__ void __launch_bounds__(FT, 1) fwd_fast_kernel(const __grid_constant__ FwdArgs a) {
  extern __shared__ __align__(16) double smem[];
  const DevStruct &S = a.S;
  const int n = S.n, m = S.m, t = threadIdx.x, z = S.z, zl = S.z + S.l;
  const bcone_settings &st = a.st;
  FastGeom g;
  fwdf_geom(n, m, g);
  FSmem M;
  carve_fast(M, smem, g, n);
  if (t == 0) { mbar_init(M.bar, 1); fence_mbar_init(); }
  // vectors: the padding entries are read by the tile products (against zero matrix entries) and must stay finite
  for (int k = t; k < 9 * g.npad + 7 * g.mpad; k += FT) M.vx[k] = 0.0;
  __syncthreads();
  uint32_t tma_phase = 0;
  const ColPlan plN = make_colplan(n, n);
  const double rho_x = st.rho_x, alpha = st.alpha, dtau = BC_TAU_FACTOR;
  const int R = t / g.CT, C = t - R * g.CT;
  const bool act = R < g.RTu;
  const int nwc = (max(m, n) + 31) >> 5;   // warps owning an output of the products
  const bool p_tma = (S.nnzP % 2 == 0) && (((uintptr_t)a.P_vals & 15) == 0) && ((size_t)S.nnzP * 8 < (1u << 20));
  double ar[TR][TC];

  for (;;) {
    if (t == 0) M.ibuf[0] = atomicAdd(a.counter, 1);
    __syncthreads();
    const int inst = M.ibuf[0];
    if (inst >= a.B) break;
    const double *Ag = a.A_vals + (size_t)inst * S.nnzA;
    const double *Pg = (a.P_vals && S.nnzP > 0) ? a.P_vals + (size_t)inst * S.nnzP : nullptr;
    const double *bg = a.b + (size_t)inst * m, *cg = a.c + (size_t)inst * n;
    PhaseTimer pt; pt.start(a.prof);
    PhaseTimer pi; pi.start(a.prof);

    // ---- stage the instance ----
    if (a.use_tma) {
      if (t == 0) {
        fence_proxy_async();
        mbar_expect_tx(M.bar, (uint32_t)(S.nnzA * sizeof(double)));
        tma_bulk_g2s(M.X, Ag, (uint32_t)(S.nnzA * sizeof(double)), M.bar);
      }
    } else {
      for (int k = t; k < S.nnzA; k += FT) M.X[k] = Ag[k];
    }
    double nb0 = 0, nc0 = 0;
    if (t < m) { const double v = bg[t]; M.bh()[t] = v; M.Dm()[t] = 1.0; nb0 = fabs(v); }
    if (t < n) { const double v = cg[t]; M.ch()[t] = v; M.En()[t] = 1.0; nc0 = fabs(v); }
    // P as a packed symmetric matrix (lower, row j at j(j+1)/2) in the factor's buffer
    auto scatter_P = [&]() {
      if (!Pg) return;
      if (!S.p_dense) { for (int e = t; e < g.npk; e += FT) M.Li[e] = 0.0; __syncthreads(); }
      for (int k = t; k < S.nnzP; k += FT) {
        const int i = __ldg(S.P_rowof + k), j = __ldg(S.P_indices + k);   // j >= i
        M.Li[((j * (j + 1)) >> 1) + i] = Pg[k];
      }
    };
    scatter_P();
    if (a.use_tma) { mbar_wait(M.bar, tma_phase); tma_phase ^= 1; }
    __syncthreads();
    // ---- register tiles ----
#pragma unroll
    for (int r = 0; r < TR; r++)
#pragma unroll
      for (int c = 0; c < TC; c++) {
        const int i = TR * R + r, j = TC * C + c;
        ar[r][c] = (act && i < m && j < n) ? M.X[i * n + j] : 0.0;
      }
    __syncthreads();   // X is free: partial buffers of the Ruiz passes
    pt.stamp(0);

    // ---- Ruiz equilibration: A^ = D A E, P^ = E P E ----
    if (st.normalize) {
      for (int pass = 0; pass < st.ruiz_passes; pass++) {
        pi.skip();
        if (act) {
          double e[TC], d[TR], rowp[TR] = {0, 0, 0, 0};
#pragma unroll
          for (int c = 0; c < TC; c += 2) { const double2 v = *reinterpret_cast<const double2 *>(M.En() + TC * C + c); e[c] = v.x; e[c + 1] = v.y; }
#pragma unroll
          for (int r = 0; r < TR; r += 2) { const double2 v = *reinterpret_cast<const double2 *>(M.Dm() + TR * R + r); d[r] = v.x; d[r + 1] = v.y; }
          double2 *dst = reinterpret_cast<double2 *>(M.XC() + R * g.npad + TC * C);
#pragma unroll
          for (int c = 0; c < TC; c += 2) {
            double c0 = 0, c1 = 0;
#pragma unroll
            for (int r = 0; r < TR; r++) {
              const double v0 = fabs(ar[r][c]) * e[c] * d[r], v1 = fabs(ar[r][c + 1]) * e[c + 1] * d[r];
              c0 = dmax(c0, v0); c1 = dmax(c1, v1);
              rowp[r] = dmax(rowp[r], dmax(v0, v1));
            }
            dst[c >> 1] = make_double2(c0, c1);
          }
#pragma unroll
          for (int r = 0; r < TR; r++) M.XR()[(TR * R + r) * g.CT + C] = rowp[r];
        }
        __syncthreads();
        if (t < m) { const double *p = M.XR() + t * g.CT; double r = 0; for (int c = 0; c < g.CT; c++) r = dmax(r, p[c]); M.tm()[t] = r; }
        {
          const int j = t - (FT - 128);   // the upper warps take the column maxima
          if (j >= 0 && j < n) { const double *p = M.XC() + j; double r = 0; for (int Rr = 0; Rr < g.RTu; Rr++) r = dmax(r, p[Rr * g.npad]); M.tn()[j] = r; }
          if (n > 128 && t < n && t >= 128) { const double *p = M.XC() + t; double r = 0; for (int Rr = 0; Rr < g.RTu; Rr++) r = dmax(r, p[Rr * g.npad]); M.tn()[t] = r; }
        }
        __syncthreads();
        pi.stamp(16);
        if (Pg) {   // column maxima of |P^|: four lanes per index over the packed symmetric matrix
          const int j = t >> 2, q = t & 3;
          double mx = 0;
          if (j < n) {
            const double ej = M.En()[j];
            for (int i = q; i < n; i += 4) {
              const int lo = min(i, j), hi = max(i, j);
              const double p = M.Li[((hi * (hi + 1)) >> 1) + lo];
              const double elo = i < j ? M.En()[i] : ej, ehi = i < j ? ej : M.En()[i];
              mx = dmax(mx, fabs(p * elo * ehi));
            }
          }
          mx = dmax(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
          mx = dmax(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
          if (q == 0 && j < n) M.tn()[j] = dmax(M.tn()[j], mx);
          __syncthreads();
        }
        pi.stamp(17);
        if (t < m) { const double r = M.tm()[t]; M.Dm()[t] *= fmin(fmax(r < 1e-8 ? 1.0 : rsqrt(r), BC_EQ_MIN), BC_EQ_MAX); }
        if (t < n) { const double r = M.tn()[t]; M.En()[t] *= fmin(fmax(r < 1e-8 ? 1.0 : rsqrt(r), BC_EQ_MIN), BC_EQ_MAX); }
        __syncthreads();
        pi.stamp(18);
      }
      if (st.ruiz_passes > 0 && act) {   // A^ = D A E on the tiles
        double e[TC], d[TR];
#pragma unroll
        for (int c = 0; c < TC; c += 2) { const double2 v = *reinterpret_cast<const double2 *>(M.En() + TC * C + c); e[c] = v.x; e[c + 1] = v.y; }
#pragma unroll
        for (int r = 0; r < TR; r += 2) { const double2 v = *reinterpret_cast<const double2 *>(M.Dm() + TR * R + r); d[r] = v.x; d[r + 1] = v.y; }
#pragma unroll
        for (int r = 0; r < TR; r++)
#pragma unroll
          for (int c = 0; c < TC; c++) ar[r][c] *= d[r] * e[c];
      }
    }
    double sigma;
    {
      double v[4] = {nb0, nc0, 0, 0};
      if (t < m) { const double q = M.Dm()[t] * M.bh()[t]; M.bh()[t] = q; v[2] = fabs(q); }
      if (t < n) { const double q = M.En()[t] * M.ch()[t]; M.ch()[t] = q; v[3] = fabs(q); }
      block_reduce<4, true>(v, M.red);
      nb0 = v[0]; nc0 = v[1];
      sigma = fmax(v[2], v[3]);
      sigma = (!st.normalize || sigma < 1e-6) ? 1.0 : 1.0 / sigma;
      if (t < m) M.bh()[t] *= sigma;
      if (t < n) M.ch()[t] *= sigma;
      __syncthreads();
    }
    pt.stamp(1);

    double scale = st.scale, gRg = 0, ry_z = 0, ry_l = 0;
    int status = BCONE_INACCURATE, it = 0;
    if (t < n) { M.wx()[t] = 0; M.ux()[t] = 0; M.utx()[t] = 0; }
    if (t < m) { M.wy()[t] = 0; M.uy()[t] = 0; M.uty()[t] = 0; }
    double w_tau = 1.0, u_tau = 0.0, ut_tau = 0.0;
    double sum_log = 0, rp = nan(""), rd = nan(""), gap = nan("");
    int n_log = 0, last_up = 0;
    int next_check = st.check_interval < 10 ? st.check_interval : 10, prev_it = 0;
    double prev_lr = 0;
    bool refactor = true, first = true;

    for (it = 1; it <= st.max_iters; it++) {
      if (refactor) {
        // Factorisation at the current scale (the one place it is written, so the tiles stay in registers):
        // stage A^ from the tiles -> K -> Cholesky -> Linv -> Kinv; then g = (R_z + M)^{-1} h and g'Rg.
        PhaseTimer pf; pf.start(a.prof);
        if (!first) { scatter_P(); }   // the factor's buffer held P in CSR order for the checks
        if (act) {
#pragma unroll
          for (int r = 0; r < TR; r++) {
            const int i = TR * R + r;
            if (i < m) {
#pragma unroll
              for (int c = 0; c < TC; c++) { const int j = TC * C + c; if (j < n) M.X[i * n + j] = ar[r][c]; }
            }
          }
        }
        __syncthreads();
        form_K(M.X, m, n, z, scale, rho_x, M.Li, Pg != nullptr, M.En());
        pf.stamp(19);
        const bool okf = chol_inv_packed(M.Li, n, M.red);
        if (!okf) { status = BCONE_FAILED; if (first) it = 0; break; }
        pf.stamp(21);
        // the tiles come back from the staged copy: nothing has to stay live across the factorisation
#pragma unroll
        for (int r = 0; r < TR; r++)
#pragma unroll
          for (int c = 0; c < TC; c++) {
            const int i = TR * R + r, j = TC * C + c;
            ar[r][c] = (act && i < m && j < n) ? M.X[i * n + j] : 0.0;
          }
        __syncthreads();
        form_Kinv(M.Li, n, g.npad, M.Kinv());
        if (Pg) {   // the factor's buffer now carries P (CSR order) for the termination checks
          if (p_tma) {
            if (t == 0) {
              fence_proxy_async();
              mbar_expect_tx(M.bar, (uint32_t)(S.nnzP * sizeof(double)));
              tma_bulk_g2s(M.Li, Pg, (uint32_t)(S.nnzP * sizeof(double)), M.bar);
            }
          } else {
            for (int k = t; k < S.nnzP; k += FT) M.Li[k] = Pg[k];
          }
        }
        pf.stamp(20);
        ry_z = 1.0 / (BC_ZERO_CONE_FACTOR * scale); ry_l = 1.0 / scale;
        if (t < m) M.tm()[t] = M.bh()[t] * inv_ry_f(z, t, scale);
        __syncthreads();
        rt_cols(ar, g, act, R, C, M.tm(), M.XC(), n, [&](int j, double v) { M.tn()[j] = M.ch()[j] - v; });
        __syncthreads();
        kinv_mul(M.Kinv(), g, n, M.tn(), M.XR(), [&](int j, double v) { M.gx()[j] = v; });
        __syncthreads();
        double acc[1] = {0};
        rt_rows(ar, g, act, R, C, M.gx(), M.XR(), m, [&](int i, double v) {
          const double iry = inv_ry_f(z, i, scale), gi = (M.bh()[i] + v) * iry;
          M.gy()[i] = gi; acc[0] = fma((1.0 / iry) * gi, gi, acc[0]); });
        if (t < n) acc[0] = fma(rho_x * M.gx()[t], M.gx()[t], acc[0]);
        block_reduce<1, false>(acc, M.red);
        gRg = acc[0];
        if (Pg && p_tma) { mbar_wait(M.bar, tma_phase); tma_phase ^= 1; }
        __syncthreads();
        pf.stamp(22);
        pt.stamp(2);
        refactor = false; first = false;
      }
      double d4[4] = {0, 0, 0, 0};   // mu'g, p'Rg, p'Rp, p'mu (R-weighted)
      auto dots = [&](double r, double pk, double wk, double gk) {
        d4[0] = fma(r * wk, gk, d4[0]); d4[1] = fma(r * pk, gk, d4[1]);
        d4[2] = fma(r * pk, pk, d4[2]); d4[3] = fma(r * pk, wk, d4[3]);
      };
      pi.skip();
      rt_cols(ar, g, act, R, C, M.wy(), M.XC(), n, [&](int j, double v) { M.tn()[j] = rho_x * M.wx()[j] - v; });
      __syncthreads();
      pi.stamp(23);
      double px = 0, py = 0;
      kinv_mul(M.Kinv(), g, n, M.tn(), M.XR(), [&](int j, double v) { px = v; M.utx()[j] = v; dots(rho_x, v, M.wx()[j], M.gx()[j]); });
      __syncthreads();
      pi.stamp(24);
      rt_rows(ar, g, act, R, C, M.utx(), M.XR(), m, [&](int i, double v) {
        const bool zr = i < z;
        const double iry = zr ? BC_ZERO_CONE_FACTOR * scale : scale, wk = M.wy()[i];
        py = wk + v * iry;
        dots(zr ? ry_z : ry_l, py, wk, M.gy()[i]); });
      pi.stamp(26);
      reduce4_lead(d4, M.red, nwc);
      pi.stamp(27);
      const double qa = dtau + gRg, qb = d4[0] - 2.0 * d4[1] - dtau * w_tau, qc = d4[2] - d4[3];
      double disc = qb * qb - 4.0 * qa * qc;
      if (disc < 0) disc = 0;
      const double tau_t = (-qb + sqrt(disc)) / (2.0 * qa);
      const bool check = st.adaptive_check ? (it >= next_check || it == st.max_iters) : ((it % st.check_interval == 0) || it == st.max_iters);
      const bool fused = !check;
      if (t < n) {
        const double utk = px - tau_t * M.gx()[t], wk = M.wx()[t], uk = 2.0 * utk - wk;
        M.utx()[t] = utk; M.ux()[t] = uk;
        if (fused) M.wx()[t] = wk + alpha * (uk - utk);
      }
      if (t < m) {
        const double utk = py - tau_t * M.gy()[t], wk = M.wy()[t];
        double uk = 2.0 * utk - wk;
        if (t >= z && t < zl) uk = fmax(uk, 0.0);
        M.uty()[t] = utk; M.uy()[t] = uk;
        if (fused) M.wy()[t] = wk + alpha * (uk - utk);
      }
      ut_tau = tau_t; u_tau = fmax(2.0 * tau_t - w_tau, 0.0);
      if (fused) w_tau += alpha * (u_tau - ut_tau);
      __syncthreads();
      pi.stamp(28);
      pt.stamp(3);
      if (check) {
        // ---- termination quantities on the un-normalised data ----
        const double tau = u_tau;
        double ax = 0, aty = 0, pxu = 0;
        rt_rows(ar, g, act, R, C, M.ux(), M.XR(), m, [&](int i, double v) { ax = v; });
        rt_cols(ar, g, act, R, C, M.uy(), M.XC(), n, [&](int j, double v) { aty = v; });
        if (t < n) { M.tn2()[t] = 0.0; M.tn3()[t] = M.En()[t] * M.ux()[t]; }
        __syncthreads();
        if (Pg) {  // P^ u_x = E (P (E u_x)); scratch: the column-partial buffer
          P_mul(S, M.Li, M.tn3(), M.XC(), [&](int j, double v) { M.tn2()[j] += v; }, plN);
          if (t < n) pxu = M.tn2()[t] * M.En()[t];
        }
        double sm[3] = {0, 0, 0};   // xPx_u, ctx_u, bty_u
        double mx[7] = {0, 0, 0, 0, 0, 0, 0};  // rp, nAx, nS, nAxs, rd, nPx, nATy
        if (t < m) {
          const double rsk = (M.uy()[t] - (2.0 * M.uty()[t] - M.wy()[t])) / inv_ry_f(z, t, scale);
          const double sc = 1.0 / (M.Dm()[t] * sigma);
          mx[0] = fabs(ax + rsk - M.bh()[t] * tau) * sc;
          mx[1] = fabs(ax) * sc; mx[2] = fabs(rsk) * sc;
          mx[3] = fabs(ax + rsk) * sc;
          sm[2] = M.bh()[t] * M.uy()[t];
        }
        if (t < n) {
          const double sc = 1.0 / (M.En()[t] * sigma);
          mx[4] = fabs(pxu + aty + M.ch()[t] * tau) * sc;
          mx[5] = fabs(pxu) * sc; mx[6] = fabs(aty) * sc;
          sm[0] = M.ux()[t] * pxu; sm[1] = M.ch()[t] * M.ux()[t];
        }
        block_reduce<3, false>(sm, M.red);
        block_reduce<7, true>(mx, M.red);
        const double s2 = sigma * sigma;
        bool done = false;
        if (tau > 1e-12) {
          const double itau = 1.0 / tau;
          const double xPx = sm[0] * itau * itau / s2, ctx = sm[1] * itau / s2, bty = sm[2] * itau / s2;
          rp = mx[0] * itau; rd = mx[4] * itau; gap = fabs(xPx + ctx + bty);
          const double np_ = fmax(fmax(mx[1] * itau, mx[2] * itau), nb0);
          const double nd_ = fmax(fmax(mx[5] * itau, mx[6] * itau), nc0);
          const double tp = st.eps_abs + st.eps_rel * np_, td = st.eps_abs + st.eps_rel * nd_;
          const double tg = st.eps_abs + st.eps_rel * fmax(fmax(fabs(xPx), fabs(ctx)), fabs(bty));
          if (rp <= tp && rd <= td && gap <= tg) { status = BCONE_SOLVED; done = true; }
          else if (st.adaptive_check) {
            const double lr = log(fmax(fmax(rp / tp, rd / td), gap / tg));
            int step = st.check_interval;
            if (prev_it > 0 && lr < prev_lr) { const double need = lr * (it - prev_it) / (prev_lr - lr); step = (int)ceil(0.9 * need) + 1; }
            step = max(3, min(step, st.check_interval));
            prev_it = it; prev_lr = lr; next_check = it + step;
          }
          if (!done && st.adaptive_scale) {
            const double relp = rp / fmax(np_, 1e-18), reld = rd / fmax(nd_, 1e-18);
            if (relp > 0 && reld > 0) { sum_log += log(relp) - log(reld); n_log++; }
          }
        }
        if (st.adaptive_check && next_check <= it) next_check = it + st.check_interval;
        if (!done) {
          const double bty_c = sm[2] / s2, ctx_c = sm[1] / s2;
          if (bty_c < 0 && mx[6] / (-bty_c) <= st.eps_infeas) { status = BCONE_INFEASIBLE; done = true; }
          else if (ctx_c < 0 && fmax(mx[5], mx[3]) / (-ctx_c) <= st.eps_infeas) { status = BCONE_UNBOUNDED; done = true; }
        }
        if (done) break;
        if (st.adaptive_scale && n_log > 0 && it - last_up >= BC_RESCALE_MIN_ITERS) {
          const double fac = sqrt(exp(sum_log / n_log));
          if (fac > 3.1622776601683795 || fac < 0.31622776601683794) {
            const double ns = fmin(fmax(scale * fac, BC_MIN_SCALE), BC_MAX_SCALE);
            if (ns != scale) {
              // keep R (w + u - 2 u~) invariant across the metric change (y block only)
              const double ratio = ns / scale;  // r_old / r_new
              if (t < m) M.wy()[t] = ratio * (M.wy()[t] + M.uy()[t] - 2.0 * M.uty()[t]) + 2.0 * M.uty()[t] - M.uy()[t];
              scale = ns;
              refactor = true;
              sum_log = 0; n_log = 0; last_up = it;
            }
          }
        }
        pt.stamp(4);
        if (it < st.max_iters) {  // (the last iterate keeps w so that s = R(u - t) is recoverable)
          if (t < n) M.wx()[t] += alpha * (M.ux()[t] - M.utx()[t]);
          if (t < m) M.wy()[t] += alpha * (M.uy()[t] - M.uty()[t]);
          w_tau += alpha * (u_tau - ut_tau);
          __syncthreads();
        }
      }
    }
    if (it > st.max_iters) it = st.max_iters;
    pt.stamp(4);
    // ---- write back ----
    {
      double *xo = a.x + (size_t)inst * n, *yo = a.y + (size_t)inst * m, *so = a.s + (size_t)inst * m;
      if (status == BCONE_SOLVED || status == BCONE_INACCURATE) {
        double tau = u_tau;
        if (!(tau > 1e-12)) tau = 1e-12;
        const double k0 = 1.0 / (sigma * tau);
        if (t < n) xo[t] = M.En()[t] * M.ux()[t] * k0;
        if (t < m) {
          const double rsk = (M.uy()[t] - (2.0 * M.uty()[t] - M.wy()[t])) / inv_ry_f(z, t, scale);
          yo[t] = M.Dm()[t] * M.uy()[t] * k0;
          so[t] = rsk * k0 / M.Dm()[t];
        }
      } else {
        const double qn = nan("");
        if (t < n) xo[t] = qn;
        if (t < m) { yo[t] = qn; so[t] = qn; }
      }
      if (t == 0) {
        a.status[inst] = status; a.iters[inst] = it;
        if (a.resid) { a.resid[inst * 3 + 0] = rp; a.resid[inst * 3 + 1] = rd; a.resid[inst * 3 + 2] = gap; }
      }
    }
    __syncthreads();
  }
}

// ----------------------------------------------------------------------------- host launcher
extern "C" size_t bc_fwdf_smem_bytes(int n, int m) { return fwdf_smem_doubles(n, m) * sizeof(double); }
extern "C" int bc_fwdf_threads(void) { return FT; }
// Eligibility beyond "dense A, polyhedral cones, direct mode" (checked by the caller): the tile grid has to
// cover the matrix with at least half of the threads busy.
extern "C" int bc_fwdf_eligible(int n, int m) {
  FastGeom g;
  if (!fwdf_geom(n, m, g)) return 0;
  return g.CT * g.RTu >= FT / 2;
}
extern "C" cudaError_t bc_fwdf_configure(size_t smem) {
  return cudaFuncSetAttribute(fwd_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}
extern "C" cudaError_t bc_fwdf_occupancy(size_t smem, int *ctas_per_sm) {
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, fwd_fast_kernel, FT, smem);
}
extern "C" cudaError_t bc_fwdf_launch(const FwdArgs *a, int grid, size_t smem, cudaStream_t stream) {
  fwd_fast_kernel<<<grid, FT, smem, stream>>>(*a);
  return cudaGetLastError();
}
