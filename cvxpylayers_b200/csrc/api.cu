// api.cu -- the C ABI of libbcone.so (declared in include/bcone.h): handle management,
// structure upload, launch geometry, and the five entry points the reference-side binding
// calls.  No torch types, no exceptions across the boundary.
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <cstdlib>
#include "common.cuh"

// ---- kernels / launchers implemented in fwd.cu, bwd.cu, pack.cu ----
extern "C" {
size_t bc_fwd_smem_bytes(int n, int m, int nnzA, int threads, int max_psd, int indirect, int ns, int nexp);
size_t bc_fwd_ws_doubles(int n, int m, int with_factor);
cudaError_t bc_fwd_configure(int dense, int indirect, size_t smem, int small_cta);
cudaError_t bc_fwd_occupancy(int dense, int indirect, int threads, size_t smem, int *ctas, int small_cta);
cudaError_t bc_fwd_launch(const FwdArgs *a, int indirect, int grid, int threads, size_t smem, cudaStream_t st, int small_cta);
size_t bc_fwdf_smem_bytes(int n, int m);
int bc_fwdf_threads(void);
size_t bc_fwdf_cache_doubles(int n, int m);
int bc_fwdf_eligible(int n, int m);
cudaError_t bc_fwdf_configure(int n, int m, size_t smem);
cudaError_t bc_fwdf_occupancy(int n, int m, size_t smem, int *ctas);
cudaError_t bc_fwdf_launch(const FwdArgs *a, int grid, size_t smem, cudaStream_t st);
size_t bc_bwd_ws_doubles(int n, int m, int npoly);
size_t bc_bwd_smem_bytes(int n, int m, int npoly, int nnzA, int nnzP_smem, int threads, int max_psd, int psd_total, int nexp, int vec_global);
cudaError_t bc_bwd_configure(int dense, size_t smem, int small_cta);
cudaError_t bc_bwd_occupancy(int dense, int threads, size_t smem, int *ctas, int small_cta);
cudaError_t bc_bwd_launch(const BwdArgs *a, int grid, int threads, size_t smem, cudaStream_t st, int small_cta);
size_t bc_bwdf_smem_bytes(int n, int m, int nnzA, int nnzP, int threads);
cudaError_t bc_bwdf_configure(int n, size_t smem);
cudaError_t bc_bwdf_occupancy(int n, int threads, size_t smem, int *ctas);
cudaError_t bc_bwdf_launch(const BwdArgs *a, int grid, int threads, size_t smem, cudaStream_t st);
size_t bc_bwdb_smem_bytes(int n, int m, int threads);
cudaError_t bc_bwdb_configure(size_t smem);
cudaError_t bc_bwdb_launch(const BwdArgs *a, int grid, int threads, size_t smem, cudaStream_t st);
cudaError_t bc_b2e(const double *in, double *out, int K, int B, int ldo, int roff, const int *smap, const int *dmap, double sign, long long ldb, cudaStream_t st);
cudaError_t bc_e2b(const double *in, double *out, int K, int B, int ldi, int roff, const int *smap, const int *dmap, double sign, long long ldb, cudaStream_t st);
cudaError_t bc_rows_from_param(const double *param, long long stride, const int *map, int K, int B, int op, double *rows, cudaStream_t st);
cudaError_t bc_param_from_rows(const double *grows, const double *param, long long stride, const int *map, int K, int B, int op, double *gparam, cudaStream_t st);
cudaError_t bc_gather_cols(const double *in, long long ld, const int *map, const double *scale, int K, int B, int op, double *out, cudaStream_t st);
cudaError_t bc_scatter_cols(const double *gout, const double *out, long long ld, const int *map, const double *scale, int K, int B, int op, double *gin, cudaStream_t st);
cudaError_t bc_p2e(const double *p, const int *rptr, const int *cols, const double *vals, double *out, int K, int B, int ldo, int roff,
                   const int *smap, const int *dmap, double sign, cudaStream_t st);
cudaError_t bc_e2p(const double *in, const int *rptr, const int *cols, const double *vals, double *dp, int K, int B, int ldi, int roff,
                   const int *smap, const int *dmap, double sign, int skip, cudaStream_t st);
}

namespace {
struct Handle {
  DevStruct S{};
  int device = 0, max_batch = 0, num_sms = 0;
  std::vector<void *> allocs;
  static constexpr int RING = 16;  // concurrent calls on different streams each get their own work-queue counters
  int *counters = nullptr;  // RING x {fwd queue, bwd queue, fail count, fallback queue}
  int slot = 0;
  int nnz_aug = 0, nb = 0;
  int *d_gather = nullptr, *d_bidx = nullptr;
  int *d_gatherP = nullptr; int nnzP_b = -1;   // boundary rows of P_eval feeding the engine's upper-triangular slots (bcone_set_boundary_quad)
  // parameter -> matrix maps (bcone_set_param_maps): CSR [rows x P1] per boundary tensor, device copies
  struct PMap { int *ptr = nullptr, *col = nullptr; double *val = nullptr; int rows = 0; };
  PMap pmA, pmq, pmP;
  int P1 = 0;
  int fwd_threads = 0, bwd_threads = 0, fwd_ctas = 0, bwd_ctas = 0;
  size_t fwd_smem = 0, bwd_smem = 0;
  int tma_ok = 0, psd_total = 0, p_in_smem = 0;
  int fwd_indirect = 0, bwd_vec_global = 0;   // large instances: CG instead of Cholesky, vectors in a global slab
  // <= 256-thread instances have a second build of the generic kernels for four resident CTAs per SM (64 registers).  It wins
  // when the batch exceeds what the 128-register build keeps resident (more CTAs overlap each other's stalls: exp-cone workload
  // +24 %) and loses when every instance is resident anyway and only its own latency counts (SDP at B = 256: -25 %), so the
  // choice is made per launch from the batch size.  BCONE_SMALL_CTA=0 disables, =2 forces.
  int fwd_small = 0, bwd_small = 0, fwd_ctas_small = 0, bwd_ctas_small = 0, small_mode = 1;
  int fwd_factor_global = 0;                  // in between: values on chip, vectors + packed Cholesky factor in the slab (direct solve from L2 / HBM)
  size_t fwd_ws_stride = 0, bwd_ws_stride = 0;
  // Per-stream scratch slabs (one per CTA of the grid): launches on different streams may overlap, launches on one
  // stream cannot, so the stream is the unit of ownership.  Allocated on first use.
  struct StreamWs { cudaStream_t s; double *fwd = nullptr, *bwd = nullptr, *aa = nullptr, *park = nullptr; size_t aa_cap = 0; };
  std::vector<StreamWs> sws;
  int block_bwd = 0, blk_threads = 0; size_t blk_smem = 0;   // KKT-block preconditioned backward (lsqr_precond = 2)
  int *fail_list[RING] = {nullptr}; int fail_cap[RING] = {0};
  int fast_fwd = 0;  // dense A, polyhedral cones, direct mode: register-tiled forward (fwd_fast.cu)
  int fast_bwd = 0;  // dense A, polyhedral cones, dense-or-no P: fused single-pass backward (bwd_fast.cu)
  long long launches = 0;
  int last_block_slot = -1;   // ring slot of the last block-preconditioned vjp (its fallback counter is read by bcone_fallback_count)
  unsigned long long *prof = nullptr;   // device [32] phase cycle counters (bcone_set_profile)
  std::string err;
};
thread_local std::string g_create_err;

template <class T>
T *upload(Handle *h, const std::vector<T> &v) {
  if (v.empty()) return nullptr;
  void *p = nullptr;
  if (cudaMalloc(&p, v.size() * sizeof(T)) != cudaSuccess) return nullptr;
  h->allocs.push_back(p);
  cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
  return (T *)p;
}
int fail(Handle *h, int code, const std::string &msg) {
  if (h) h->err = msg; else g_create_err = msg;
  return code;
}
int cuda_fail(Handle *h, cudaError_t e, const char *where) {
  return fail(h, BCONE_ECUDA, std::string(where) + ": " + cudaGetErrorString(e));
}
Handle::StreamWs *stream_ws(Handle *h, cudaStream_t s) {
  for (auto &w : h->sws) if (w.s == s) return &w;
  Handle::StreamWs w; w.s = s;
  h->sws.push_back(w);
  return &h->sws.back();
}
// slab of `doubles` per CTA for `ctas` CTAs; *cap tracks the current size in doubles (0: fixed-size slab)
bool ensure_slab(Handle *h, double **p, size_t *cap, size_t doubles) {
  if (*p && (!cap || *cap >= doubles)) return true;
  double *q = nullptr;
  if (cudaMalloc((void **)&q, doubles * sizeof(double)) != cudaSuccess) return false;
  h->allocs.push_back(q);   // (an outgrown slab stays alive until destroy: a kernel in flight may still use it)
  *p = q;
  if (cap) *cap = doubles;
  return true;
}
}  // namespace

extern "C" void bcone_default_settings(bcone_settings *st) {
  st->eps_abs = 1e-4; st->eps_rel = 1e-4; st->eps_infeas = 1e-7;
  st->alpha = 1.5; st->rho_x = 1e-6; st->scale = 0.1;
  st->lsqr_atol = 1e-8; st->lsqr_btol = 1e-8; st->lsqr_conlim = 1e8;
  st->max_iters = 100000; st->normalize = 1; st->adaptive_scale = 1; st->check_interval = 25;
  st->ruiz_passes = 10; st->lsqr_iter_lim = -1; st->lsqr_precond = 0; st->adaptive_check = 0;
  st->acceleration_lookback = 10; st->acceleration_interval = 10;   // SCS defaults
}

extern "C" const char *bcone_last_error(void *handle) {
  return handle ? ((Handle *)handle)->err.c_str() : g_create_err.c_str();
}

extern "C" int bcone_create(const bcone_desc *d, void **out) {
  if (!d || !out) return fail(nullptr, BCONE_EINVAL, "null argument");
  *out = nullptr;
  if (d->n <= 0 || d->m < 0 || d->nnzA < 0 || !d->A_indptr || (d->nnzA > 0 && !d->A_indices))
    return fail(nullptr, BCONE_EINVAL, "bad dimensions / missing A structure");
  if (d->ep < 0 || d->ed < 0) return fail(nullptr, BCONE_EINVAL, "negative cone count");
  const int n = d->n, m = d->m;
  long long rows = d->z + d->l;
  int max_psd = 0, psd_total = 0;
  for (int i = 0; i < d->nq; i++) rows += d->q[i];
  for (int i = 0; i < d->ns; i++) { int k = d->s[i]; rows += (long long)k * (k + 1) / 2; max_psd = std::max(max_psd, k); psd_total += k * k + k; }
  const int exp_start = (int)rows;
  rows += 3LL * (d->ep + d->ed);
  if (rows != m) return fail(nullptr, BCONE_EINVAL, "cone dimensions do not add up to m");
  if (d->A_indptr[0] != 0 || d->A_indptr[m] != d->nnzA) return fail(nullptr, BCONE_EINVAL, "A_indptr inconsistent with nnzA");
  if (cudaSetDevice(d->device) != cudaSuccess) return fail(nullptr, BCONE_ECUDA, "cudaSetDevice failed (no CUDA device?)");
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, d->device) != cudaSuccess) return fail(nullptr, BCONE_ECUDA, "cudaGetDeviceProperties failed");

  Handle *h = new Handle();
  h->device = d->device; h->max_batch = d->max_batch; h->num_sms = prop.multiProcessorCount;
  h->psd_total = psd_total;
  DevStruct &S = h->S;
  S.n = n; S.m = m; S.nnzA = d->nnzA; S.nnzP = d->P_indptr ? d->nnzP : 0;
  S.z = d->z; S.l = d->l; S.nq = d->nq; S.ns = d->ns; S.max_psd = max_psd;
  S.ep = d->ep; S.ed = d->ed; S.exp_start = exp_start;
  // --- host-side structure analysis ---
  std::vector<int> indptr(d->A_indptr, d->A_indptr + m + 1), indices(d->A_indices, d->A_indices + d->nnzA);
  std::vector<int> rowof(d->nnzA), colptr(n + 1, 0), rowidx(d->nnzA), perm(d->nnzA);
  bool dense = (long long)d->nnzA == (long long)m * n && d->nnzA > 0;
  for (int i = 0; i < m; i++) {
    if (indptr[i + 1] < indptr[i]) { delete h; return fail(nullptr, BCONE_EINVAL, "A_indptr not monotone"); }
    for (int k = indptr[i]; k < indptr[i + 1]; k++) {
      int j = indices[k];
      if (j < 0 || j >= n) { delete h; return fail(nullptr, BCONE_EINVAL, "A column index out of range"); }
      rowof[k] = i; colptr[j + 1]++;
      if (dense && j != k - indptr[i]) dense = false;
    }
    if (dense && indptr[i + 1] - indptr[i] != n) dense = false;
  }
  for (int j = 0; j < n; j++) colptr[j + 1] += colptr[j];
  {
    std::vector<int> fill(colptr.begin(), colptr.end() - 1);
    for (int k = 0; k < d->nnzA; k++) { int p = fill[indices[k]]++; rowidx[p] = rowof[k]; perm[p] = k; }
  }
  S.dense = dense ? 1 : 0;
  std::vector<int> ctype, cstart, csize, corder;
  {
    int off = d->z + d->l;
    for (int i = 0; i < d->nq; i++) { ctype.push_back(BC_CSOC); cstart.push_back(off); csize.push_back(d->q[i]); corder.push_back(0); off += d->q[i]; }
    for (int i = 0; i < d->ns; i++) { int k = d->s[i], sz = k * (k + 1) / 2; ctype.push_back(BC_CPSD); cstart.push_back(off); csize.push_back(sz); corder.push_back(k); off += sz; }
  }
  S.ncones = (int)ctype.size();
  S.A_indptr = upload(h, indptr); S.A_indices = upload(h, indices); S.A_rowof = upload(h, rowof);
  S.At_colptr = upload(h, colptr); S.At_rowidx = upload(h, rowidx); S.At_perm = upload(h, perm);
  S.cone_type = upload(h, ctype); S.cone_start = upload(h, cstart); S.cone_size = upload(h, csize); S.cone_order = upload(h, corder);
  if (S.nnzP > 0) {
    std::vector<int> pptr(d->P_indptr, d->P_indptr + n + 1), pidx(d->P_indices, d->P_indices + S.nnzP), prow(S.nnzP);
    for (int i = 0; i < n; i++) for (int k = pptr[i]; k < pptr[i + 1]; k++) {
      if (pidx[k] < i || pidx[k] >= n) { bcone_destroy(h); return fail(nullptr, BCONE_EINVAL, "P must be upper triangular CSR"); }
      prow[k] = i;
    }
    S.P_indptr = upload(h, pptr); S.P_indices = upload(h, pidx); S.P_rowof = upload(h, prow);
    // CSC view of the upper triangle + dense-pattern detection (row-major full upper triangle)
    std::vector<int> pc(n + 1, 0), pr(S.nnzP), pp(S.nnzP);
    bool pd = (long long)S.nnzP == (long long)n * (n + 1) / 2;
    for (int i = 0; i < n; i++) {
      if (pd && pptr[i + 1] - pptr[i] != n - i) pd = false;
      for (int k = pptr[i]; k < pptr[i + 1]; k++) { pc[pidx[k] + 1]++; if (pd && pidx[k] != i + (k - pptr[i])) pd = false; }
    }
    for (int j = 0; j < n; j++) pc[j + 1] += pc[j];
    { std::vector<int> fill(pc.begin(), pc.end() - 1); for (int k = 0; k < S.nnzP; k++) { int p = fill[pidx[k]]++; pr[p] = prow[k]; pp[p] = k; } }
    S.Pt_colptr = upload(h, pc); S.Pt_rowidx = upload(h, pr); S.Pt_perm = upload(h, pp);
    S.p_dense = pd ? 1 : 0;
  }
  if (cudaMalloc((void **)&h->counters, Handle::RING * 4 * sizeof(int)) != cudaSuccess) { bcone_destroy(h); return fail(nullptr, BCONE_ENOMEM, "cudaMalloc counters"); }
  h->allocs.push_back(h->counters);

  // --- launch geometry: threads by problem size, shared memory must hold the whole instance ---
  const size_t smem_cap = prop.sharedMemPerBlockOptin;
  int threads = d->nnzA >= 8192 ? 512 : (d->nnzA >= 1024 ? 256 : 128);
  while (threads < 512 && threads < n) threads *= 2;  // transposed products want one lane per column
  const int npoly = d->z + d->l;
  // DIRECT (everything on chip) if the instance fits; else values on chip with the vectors and the packed Cholesky factor
  // in a per-CTA slab of global memory (two triangular products per iteration read it from L2: n <= 512, i.e. <= 1 MB
  // per CTA); else INDIRECT (conjugate gradients, SCS's "indirect" mode).  BCONE_FWD_MODE=indirect forces the last one.
  const char *fm = getenv("BCONE_FWD_MODE");
  const bool force_indirect = fm && std::string(fm) == "indirect";
  auto pick_fwd = [&]() -> bool {
    for (int ind = 0; ind <= 1; ind++)
      for (int tt = threads; tt >= 64; tt /= 2) {
        size_t sm = bc_fwd_smem_bytes(n, m, d->nnzA, tt, max_psd, ind, d->ns, d->ep + d->ed);
        if (sm <= smem_cap) {
          h->fwd_threads = tt; h->fwd_smem = sm; h->fwd_indirect = ind;
          // (n <= 512: one thread per column in the transposed triangular product; measured on the sparse LP with n = 1000 the
          //  slab mode streams 8 MB of factor per iteration and CTA from HBM and loses to conjugate gradients, 12.6 s vs 5.8 s per
          //  512-batch, while at n = 101 it wins 37x)
          if (ind && !force_indirect && n <= 512) { h->fwd_indirect = 0; h->fwd_factor_global = 1; }
          return true;
        }
      }
    return false;
  };
  auto pick_bwd = [&]() -> bool {   // prefer P staged in shared memory, then vectors on chip, then vectors in L2
    for (int vg = 0; vg <= 1; vg++)
      for (int psm = (S.nnzP > 0 ? 1 : 0); psm >= 0; psm--)
        for (int tt = threads; tt >= 64; tt /= 2) {
          size_t sm = bc_bwd_smem_bytes(n, m, npoly, d->nnzA, psm ? S.nnzP : 0, tt, max_psd, psd_total, d->ep + d->ed, vg);
          if (sm <= smem_cap) { h->bwd_threads = tt; h->bwd_smem = sm; h->p_in_smem = psm; h->bwd_vec_global = vg; return true; }
          if (psm) break;  // do not trade threads for P residency
        }
    return false;
  };
  // fast backward path: same launch geometry fields, different kernel
  if (S.dense && S.ncones == 0 && d->ep + d->ed == 0 && n <= 128 && (n % 2) == 0 && (S.nnzP == 0 || S.p_dense)) {
    for (int tt = threads; tt >= 64; tt /= 2) {
      size_t sm = bc_bwdf_smem_bytes(n, m, d->nnzA, S.nnzP, tt);
      if (sm <= smem_cap) { h->fast_bwd = 1; h->bwd_threads = tt; h->bwd_smem = sm; h->p_in_smem = S.nnzP > 0; break; }
    }
  }
  if (h->fast_bwd && S.nnzP > 0 && 6 * (n + m + 1) >= 8 * n + 72) {   // block-preconditioned variant for strongly convex QPs
    for (int tt = threads; tt >= 128; tt /= 2) {
      size_t sm = bc_bwdb_smem_bytes(n, m, tt);
      if (sm <= smem_cap) { h->block_bwd = 1; h->blk_threads = tt; h->blk_smem = sm; break; }
    }
  }
  if (!pick_fwd() || (!h->fast_bwd && !pick_bwd())) {
    char buf[256];
    snprintf(buf, sizeof buf, "instance does not fit the shared-memory-resident engine (fwd %zu B / bwd %zu B needed, %zu B per CTA available)",
             bc_fwd_smem_bytes(n, m, d->nnzA, 64, max_psd, 1, d->ns, d->ep + d->ed), bc_bwd_smem_bytes(n, m, npoly, d->nnzA, 0, 64, max_psd, psd_total, d->ep + d->ed, 1), smem_cap);
    bcone_destroy(h);
    return fail(nullptr, BCONE_EUNSUPPORTED, buf);
  }
  cudaError_t e;
  // register-tiled forward (fwd_fast.cu) when the structure allows it; BCONE_NO_FAST_FWD=1 keeps the generic kernel
  if (S.dense && S.ncones == 0 && d->ep + d->ed == 0 && !h->fwd_indirect && !h->fwd_factor_global && bc_fwdf_eligible(n, m) &&
      bc_fwdf_smem_bytes(n, m) <= smem_cap && !(getenv("BCONE_NO_FAST_FWD") && atoi(getenv("BCONE_NO_FAST_FWD")))) {
    if (bc_fwdf_configure(n, m, bc_fwdf_smem_bytes(n, m)) == cudaSuccess) {
      h->fast_fwd = 1; h->fwd_threads = bc_fwdf_threads(); h->fwd_smem = bc_fwdf_smem_bytes(n, m);
    }
  }
  {
    const char *sc = getenv("BCONE_SMALL_CTA");
    h->small_mode = sc ? atoi(sc) : 1;
    const bool allow = h->small_mode != 0;
    h->fwd_small = allow && !h->fast_fwd && !h->fwd_indirect && h->fwd_threads <= 256 && h->fwd_smem <= 56 * 1024;
    h->bwd_small = allow && !h->fast_bwd && h->bwd_threads <= 256 && h->bwd_smem <= 56 * 1024;
  }
  if (h->fwd_small && bc_fwd_configure(S.dense, 0, h->fwd_smem, 1) != cudaSuccess) h->fwd_small = 0;
  if (h->bwd_small && bc_bwd_configure(S.dense, h->bwd_smem, 1) != cudaSuccess) h->bwd_small = 0;
  if ((e = bc_fwd_configure(S.dense, h->fwd_indirect, h->fast_fwd ? bc_fwd_smem_bytes(n, m, d->nnzA, 64, max_psd, h->fwd_indirect, d->ns, d->ep + d->ed) : h->fwd_smem, 0)) != cudaSuccess ||
      (e = (h->fast_bwd ? bc_bwdf_configure(n, h->bwd_smem) : bc_bwd_configure(S.dense, h->bwd_smem, 0))) != cudaSuccess) {
    std::string msg = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e);
    bcone_destroy(h);
    return fail(nullptr, BCONE_ECUDA, msg);
  }
  if (h->block_bwd && (e = bc_bwdb_configure(h->blk_smem)) != cudaSuccess) h->block_bwd = 0;
  if (h->fast_fwd) bc_fwdf_occupancy(n, m, h->fwd_smem, &h->fwd_ctas);
  else bc_fwd_occupancy(S.dense, h->fwd_indirect, h->fwd_threads, h->fwd_smem, &h->fwd_ctas, 0);
  if (h->fast_bwd) bc_bwdf_occupancy(n, h->bwd_threads, h->bwd_smem, &h->bwd_ctas);
  else bc_bwd_occupancy(S.dense, h->bwd_threads, h->bwd_smem, &h->bwd_ctas, 0);
  if (h->fwd_ctas < 1) h->fwd_ctas = 1;
  if (h->bwd_ctas < 1) h->bwd_ctas = 1;
  if (h->fwd_small) { bc_fwd_occupancy(S.dense, 0, h->fwd_threads, h->fwd_smem, &h->fwd_ctas_small, 1); if (h->fwd_ctas_small <= h->fwd_ctas) h->fwd_small = 0; }
  if (h->bwd_small) { bc_bwd_occupancy(S.dense, h->bwd_threads, h->bwd_smem, &h->bwd_ctas_small, 1); if (h->bwd_ctas_small <= h->bwd_ctas) h->bwd_small = 0; }
  if (h->fwd_indirect || h->fwd_factor_global) h->fwd_ws_stride = bc_fwd_ws_doubles(n, m, h->fwd_factor_global);
  if (h->bwd_vec_global && !h->fast_bwd) h->bwd_ws_stride = bc_bwd_ws_doubles(n, m, npoly);
  h->tma_ok = (d->nnzA > 0 && (d->nnzA % 2) == 0 && (size_t)d->nnzA * 8 < (1u << 20)) ? 1 : 0;
  *out = h;
  return BCONE_OK;
}

extern "C" void bcone_destroy(void *handle) {
  if (!handle) return;
  Handle *h = (Handle *)handle;
  cudaSetDevice(h->device);
  for (void *p : h->allocs) cudaFree(p);
  delete h;
}

extern "C" int bcone_set_boundary(void *handle, int32_t nnz_aug, const int32_t *gather, int32_t nb, const int32_t *b_idx) {
  Handle *h = (Handle *)handle;
  if (!h) return BCONE_EINVAL;
  if (nnz_aug != h->S.nnzA + nb || (h->S.nnzA > 0 && !gather) || (nb > 0 && !b_idx)) return fail(h, BCONE_EINVAL, "set_boundary: inconsistent sizes");
  for (int k = 0; k < h->S.nnzA; k++) if (gather[k] < 0 || gather[k] >= h->S.nnzA) return fail(h, BCONE_EINVAL, "set_boundary: gather out of range");
  for (int r = 0; r < nb; r++) if (b_idx[r] < 0 || b_idx[r] >= h->S.m) return fail(h, BCONE_EINVAL, "set_boundary: b_idx out of range");
  cudaSetDevice(h->device);
  h->nnz_aug = nnz_aug; h->nb = nb;
  h->d_gather = upload(h, std::vector<int>(gather, gather + h->S.nnzA));
  h->d_bidx = upload(h, std::vector<int>(b_idx, b_idx + nb));
  return BCONE_OK;
}

extern "C" int bcone_set_boundary_quad(void *handle, int32_t nnzP_boundary, const int32_t *gatherP) {
  Handle *h = (Handle *)handle;
  if (!h) return BCONE_EINVAL;
  if (h->S.nnzP > 0 && (!gatherP || nnzP_boundary < 1)) return fail(h, BCONE_EINVAL, "set_boundary_P: missing gather map");
  for (int k = 0; k < h->S.nnzP; k++) if (gatherP[k] < 0 || gatherP[k] >= nnzP_boundary) return fail(h, BCONE_EINVAL, "set_boundary_P: gather out of range");
  cudaSetDevice(h->device);
  h->nnzP_b = nnzP_boundary;
  h->d_gatherP = h->S.nnzP > 0 ? upload(h, std::vector<int>(gatherP, gatherP + h->S.nnzP)) : nullptr;
  return BCONE_OK;
}

#define CK(call, where) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(h, e_, where); } while (0)

// ---- parameter -> matrix affine map fused into ingest / emit ----
extern "C" int bcone_set_param_maps(void *handle, int32_t P1, const int32_t *A_ptr, const int32_t *A_col, const double *A_val,
                                    const int32_t *q_ptr, const int32_t *q_col, const double *q_val,
                                    const int32_t *P_ptr, const int32_t *P_col, const double *P_val) {
  Handle *h = (Handle *)handle;
  if (!h || P1 <= 0 || !A_ptr || !q_ptr) return fail(h, BCONE_EINVAL, "set_param_maps: null argument");
  if (h->nnz_aug == 0 && h->S.nnzA + h->nb != 0) return fail(h, BCONE_EINVAL, "set_param_maps: call bcone_set_boundary first");
  const DevStruct &S = h->S;
  const int rowsA = h->nnz_aug, rowsq = S.n + 1, rowsP = (P_ptr && S.nnzP > 0) ? (h->nnzP_b > 0 ? h->nnzP_b : S.nnzP) : 0;
  const int32_t *ptrs[3] = {A_ptr, q_ptr, P_ptr}, *colsv[3] = {A_col, q_col, P_col};
  const double *valsv[3] = {A_val, q_val, P_val};
  const int rows[3] = {rowsA, rowsq, rowsP};
  std::vector<int> count(P1, 0);
  for (int w = 0; w < 3; w++) {
    if (!rows[w]) continue;
    if (ptrs[w][0] != 0) return fail(h, BCONE_EINVAL, "set_param_maps: row pointer must start at 0");
    for (int r = 0; r < rows[w]; r++) if (ptrs[w][r + 1] < ptrs[w][r]) return fail(h, BCONE_EINVAL, "set_param_maps: row pointer not monotone");
    const int nz = ptrs[w][rows[w]];
    if (nz > 0 && (!colsv[w] || !valsv[w])) return fail(h, BCONE_EINVAL, "set_param_maps: missing column / value array");
    for (int e = 0; e < nz; e++) {
      if (colsv[w][e] < 0 || colsv[w][e] >= P1) return fail(h, BCONE_EINVAL, "set_param_maps: parameter index out of range");
      count[colsv[w][e]]++;
    }
  }
  cudaSetDevice(h->device);
  Handle::PMap *dst[3] = {&h->pmA, &h->pmq, &h->pmP};
  for (int w = 0; w < 3; w++) {
    *dst[w] = Handle::PMap();
    if (!rows[w]) continue;
    const int nz = ptrs[w][rows[w]];
    std::vector<int> ptr(ptrs[w], ptrs[w] + rows[w] + 1), col(std::max(nz, 1), 0);
    std::vector<double> val(std::max(nz, 1), 0.0);
    for (int e = 0; e < nz; e++) { col[e] = colsv[w][e] | (count[colsv[w][e]] == 1 ? 0x40000000 : 0); val[e] = valsv[w][e]; }   // bit 30: exclusive column
    dst[w]->ptr = upload(h, ptr); dst[w]->col = upload(h, col); dst[w]->val = upload(h, val); dst[w]->rows = rows[w];
    if (!dst[w]->ptr || !dst[w]->col || !dst[w]->val) return fail(h, BCONE_ENOMEM, "set_param_maps: cudaMalloc");
  }
  h->P1 = P1;
  return BCONE_OK;
}

extern "C" int bcone_ingest_params(void *handle, int32_t B, const double *p_stack, double *A_vals, double *P_vals, double *b, double *c, void *stream) {
  Handle *h = (Handle *)handle;
  if (!h || B <= 0 || !p_stack || !A_vals || !b || !c) return fail(h, BCONE_EINVAL, "ingest_params: null argument");
  if (h->P1 <= 0) return fail(h, BCONE_EINVAL, "ingest_params: call bcone_set_param_maps first");
  cudaStream_t st = (cudaStream_t)stream;
  const DevStruct &S = h->S;
  CK(bc_p2e(p_stack, h->pmA.ptr, h->pmA.col, h->pmA.val, A_vals, S.nnzA, B, S.nnzA, 0, h->d_gather, nullptr, -1.0, st), "ingest_params A");
  CK(cudaMemsetAsync(b, 0, (size_t)B * S.m * sizeof(double), st), "ingest_params b memset");
  CK(bc_p2e(p_stack, h->pmA.ptr, h->pmA.col, h->pmA.val, b, h->nb, B, S.m, S.nnzA, nullptr, h->d_bidx, 1.0, st), "ingest_params b");
  CK(bc_p2e(p_stack, h->pmq.ptr, h->pmq.col, h->pmq.val, c, S.n, B, S.n, 0, nullptr, nullptr, 1.0, st), "ingest_params c");
  h->launches += 3;
  if (P_vals && S.nnzP > 0) {
    if (!h->pmP.rows) return fail(h, BCONE_EINVAL, "ingest_params: structure has P but no parameter map for it");
    CK(bc_p2e(p_stack, h->pmP.ptr, h->pmP.col, h->pmP.val, P_vals, S.nnzP, B, S.nnzP, 0, h->d_gatherP, nullptr, 1.0, st), "ingest_params P");
    h->launches++;
  }
  return BCONE_OK;
}

extern "C" int bcone_emit_params(void *handle, int32_t B, const double *dA_vals, const double *dP_vals, const double *db, const double *dc,
                                 double *dp_stack, void *stream) {
  Handle *h = (Handle *)handle;
  if (!h || B <= 0 || !dA_vals || !db || !dc || !dp_stack) return fail(h, BCONE_EINVAL, "emit_params: null argument");
  if (h->P1 <= 0) return fail(h, BCONE_EINVAL, "emit_params: call bcone_set_param_maps first");
  cudaStream_t st = (cudaStream_t)stream;
  const DevStruct &S = h->S;
  const int skip = h->P1 - 1;   // the constant-1 row of p_stack is not a parameter
  CK(cudaMemsetAsync(dp_stack, 0, (size_t)h->P1 * B * sizeof(double), st), "emit_params memset");
  CK(bc_e2p(dA_vals, h->pmA.ptr, h->pmA.col, h->pmA.val, dp_stack, S.nnzA, B, S.nnzA, 0, nullptr, h->d_gather, -1.0, skip, st), "emit_params dA");
  CK(bc_e2p(db, h->pmA.ptr, h->pmA.col, h->pmA.val, dp_stack, h->nb, B, S.m, S.nnzA, h->d_bidx, nullptr, 1.0, skip, st), "emit_params db");
  CK(bc_e2p(dc, h->pmq.ptr, h->pmq.col, h->pmq.val, dp_stack, S.n, B, S.n, 0, nullptr, nullptr, 1.0, skip, st), "emit_params dc");
  h->launches += 3;
  if (dP_vals && S.nnzP > 0 && h->pmP.rows) {
    CK(bc_e2p(dP_vals, h->pmP.ptr, h->pmP.col, h->pmP.val, dp_stack, S.nnzP, B, S.nnzP, 0, nullptr, h->d_gatherP, 1.0, skip, st), "emit_params dP");
    h->launches++;
  }
  return BCONE_OK;
}

extern "C" int bcone_ingest_pitched(void *handle, int32_t B, int64_t ldb, const double *A_eval, const double *q_eval, const double *P_eval,
                                    double *A_vals, double *P_vals, double *b, double *c, void *stream) {
  Handle *h = (Handle *)handle;
  if (!h || B <= 0 || ldb < B || !A_eval || !q_eval || !A_vals || !b || !c) return fail(h, BCONE_EINVAL, "ingest: null argument / bad pitch");
  if (h->nnz_aug == 0 && h->S.nnzA + h->nb != 0) return fail(h, BCONE_EINVAL, "ingest: call bcone_set_boundary first");
  cudaStream_t st = (cudaStream_t)stream;
  const DevStruct &S = h->S;
  CK(bc_b2e(A_eval, A_vals, S.nnzA, B, S.nnzA, 0, h->d_gather, nullptr, -1.0, ldb, st), "ingest A");
  CK(cudaMemsetAsync(b, 0, (size_t)B * S.m * sizeof(double), st), "ingest b memset");
  CK(bc_b2e(A_eval, b, h->nb, B, S.m, S.nnzA, nullptr, h->d_bidx, 1.0, ldb, st), "ingest b");
  CK(bc_b2e(q_eval, c, S.n, B, S.n, 0, nullptr, nullptr, 1.0, ldb, st), "ingest c");
  h->launches += 3;
  if (P_eval && P_vals && S.nnzP > 0) { CK(bc_b2e(P_eval, P_vals, S.nnzP, B, S.nnzP, 0, h->d_gatherP, nullptr, 1.0, ldb, st), "ingest P"); h->launches++; }
  return BCONE_OK;
}
extern "C" int bcone_ingest(void *handle, int32_t B, const double *A_eval, const double *q_eval, const double *P_eval,
                            double *A_vals, double *P_vals, double *b, double *c, void *stream) {
  return bcone_ingest_pitched(handle, B, B, A_eval, q_eval, P_eval, A_vals, P_vals, b, c, stream);
}

extern "C" int bcone_emit_pitched(void *handle, int32_t B, int64_t ldb, const double *dA_vals, const double *dP_vals, const double *db,
                                  const double *dc, double *dA_eval, double *dq_eval, double *dP_eval, void *stream) {
  Handle *h = (Handle *)handle;
  if (!h || B <= 0 || ldb < B || !dA_vals || !db || !dc || !dA_eval || !dq_eval) return fail(h, BCONE_EINVAL, "emit: null argument / bad pitch");
  cudaStream_t st = (cudaStream_t)stream;
  const DevStruct &S = h->S;
  CK(bc_e2b(dA_vals, dA_eval, S.nnzA, B, S.nnzA, 0, nullptr, h->d_gather, -1.0, ldb, st), "emit dA");
  CK(bc_e2b(db, dA_eval, h->nb, B, S.m, S.nnzA, h->d_bidx, nullptr, 1.0, ldb, st), "emit db");
  CK(bc_e2b(dc, dq_eval, S.n, B, S.n, 0, nullptr, nullptr, 1.0, ldb, st), "emit dc");
  CK(cudaMemset2DAsync(dq_eval + (size_t)S.n * ldb, (size_t)ldb * sizeof(double), 0, (size_t)B * sizeof(double), 1, st), "emit dq tail");
  h->launches += 3;
  if (dP_vals && dP_eval && S.nnzP > 0) {
    // boundary rows without an engine slot (the lower triangle of a full symmetric pattern) get a zero gradient:
    // the engine reads the upper triangle only, so that is the derivative of what was computed
    if (h->d_gatherP && h->nnzP_b != S.nnzP)
      CK(cudaMemset2DAsync(dP_eval, (size_t)ldb * sizeof(double), 0, (size_t)B * sizeof(double), (size_t)h->nnzP_b, st), "emit dP memset");
    CK(bc_e2b(dP_vals, dP_eval, S.nnzP, B, S.nnzP, 0, nullptr, h->d_gatherP, 1.0, ldb, st), "emit dP"); h->launches++;
  }
  return BCONE_OK;
}
extern "C" int bcone_emit(void *handle, int32_t B, const double *dA_vals, const double *dP_vals, const double *db,
                          const double *dc, double *dA_eval, double *dq_eval, double *dP_eval, void *stream) {
  return bcone_emit_pitched(handle, B, B, dA_vals, dP_vals, db, dc, dA_eval, dq_eval, dP_eval, stream);
}

// ---- layer prologue / epilogue: index maps instead of the reference's reshape / permute / cat chains ----
// All pointers are DEVICE pointers (maps: int32, scales: double); no handle, no state.  op: 0 identity, 1 exp, 2 log.
#define CKG(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { g_create_err = std::string(__func__) + ": " + cudaGetErrorString(e_); return BCONE_ECUDA; } } while (0)
extern "C" int bcone_rows_from_param(const double *param, int64_t stride, const int32_t *map, int32_t K, int32_t B, int32_t op, double *rows, void *stream) {
  if (!param || !rows || K < 0 || B < 0) return BCONE_EINVAL;
  CKG(bc_rows_from_param(param, stride, map, K, B, op, rows, (cudaStream_t)stream));
  return BCONE_OK;
}
extern "C" int bcone_param_from_rows(const double *grows, const double *param, int64_t stride, const int32_t *map, int32_t K, int32_t B, int32_t op,
                                     double *gparam, void *stream) {
  if (!grows || !gparam || K < 0 || B < 0 || (op == 2 && !param)) return BCONE_EINVAL;
  CKG(bc_param_from_rows(grows, param, stride, map, K, B, op, gparam, (cudaStream_t)stream));
  return BCONE_OK;
}
extern "C" int bcone_gather_cols(const double *in, int64_t ld, const int32_t *map, const double *scale, int32_t K, int32_t B, int32_t op, double *out, void *stream) {
  if (!in || !out || !map || K < 0 || B < 0) return BCONE_EINVAL;
  CKG(bc_gather_cols(in, ld, map, scale, K, B, op, out, (cudaStream_t)stream));
  return BCONE_OK;
}
extern "C" int bcone_scatter_cols(const double *gout, const double *out, int64_t ld, const int32_t *map, const double *scale, int32_t K, int32_t B, int32_t op,
                                  double *gin, void *stream) {
  if (!gout || !gin || !map || K < 0 || B < 0 || (op == 1 && !out)) return BCONE_EINVAL;
  CKG(bc_scatter_cols(gout, out, ld, map, scale, K, B, op, gin, (cudaStream_t)stream));
  return BCONE_OK;
}

// ---- peer exchange: a buffer on one GPU that every rank of the node can write (CUDA IPC over NVLink) -----------------------
extern "C" int bcone_peer_alloc(int32_t device, int64_t bytes, void **ptr, void *ipc_handle64) {
  if (!ptr || !ipc_handle64 || bytes <= 0) return BCONE_EINVAL;
  if (cudaSetDevice(device) != cudaSuccess) return BCONE_ECUDA;
  void *p = nullptr;
  cudaError_t e = cudaMalloc(&p, (size_t)bytes);
  if (e != cudaSuccess) { g_create_err = std::string("bcone_peer_alloc: ") + cudaGetErrorString(e); return BCONE_ENOMEM; }
  cudaIpcMemHandle_t hnd;
  e = cudaIpcGetMemHandle(&hnd, p);
  if (e != cudaSuccess) { cudaFree(p); g_create_err = std::string("bcone_peer_alloc (ipc handle): ") + cudaGetErrorString(e); return BCONE_ECUDA; }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
  memcpy(ipc_handle64, &hnd, 64);
  *ptr = p;
  return BCONE_OK;
}
extern "C" int bcone_peer_open(int32_t device, const void *ipc_handle64, void **ptr) {
  if (!ptr || !ipc_handle64) return BCONE_EINVAL;
  if (cudaSetDevice(device) != cudaSuccess) return BCONE_ECUDA;
  cudaIpcMemHandle_t hnd;
  memcpy(&hnd, ipc_handle64, 64);
  cudaError_t e = cudaIpcOpenMemHandle(ptr, hnd, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) { g_create_err = std::string("bcone_peer_open: ") + cudaGetErrorString(e); cudaGetLastError(); return BCONE_ECUDA; }
  return BCONE_OK;
}
extern "C" int bcone_peer_close(void *ptr) { return cudaIpcCloseMemHandle(ptr) == cudaSuccess ? BCONE_OK : BCONE_ECUDA; }
extern "C" int bcone_peer_free(void *ptr) { return cudaFree(ptr) == cudaSuccess ? BCONE_OK : BCONE_ECUDA; }
// dst / src: any device pointers of this process' address space (local or peer-mapped); pitches and width in bytes.
// height = 1 is a plain copy.  Runs on the copy engines: no SM is taken from a solve that is in flight.
extern "C" int bcone_copy2d_async(void *dst, int64_t dpitch, const void *src, int64_t spitch, int64_t width, int64_t height, void *stream) {
  cudaError_t e = height <= 1 ? cudaMemcpyAsync(dst, src, (size_t)width, cudaMemcpyDefault, (cudaStream_t)stream)
                              : cudaMemcpy2DAsync(dst, (size_t)dpitch, src, (size_t)spitch, (size_t)width, (size_t)height, cudaMemcpyDefault, (cudaStream_t)stream);
  if (e != cudaSuccess) { g_create_err = std::string("bcone_copy2d_async: ") + cudaGetErrorString(e); return BCONE_ECUDA; }
  return BCONE_OK;
}

extern "C" int bcone_solve_cached(void *handle, int32_t B, const double *A_vals, const double *P_vals, const double *b, const double *c,
                                  const double *x0, const double *y0, const double *s0, double *x, double *y, double *s, int32_t *status,
                                  int32_t *iters, double *resid, void *cache, int32_t reuse, const bcone_settings *stg, void *stream);
extern "C" int bcone_solve_warm(void *handle, int32_t B, const double *A_vals, const double *P_vals, const double *b, const double *c,
                                const double *x0, const double *y0, const double *s0, double *x, double *y, double *s, int32_t *status,
                                int32_t *iters, double *resid, const bcone_settings *stg, void *stream) {
  return bcone_solve_cached(handle, B, A_vals, P_vals, b, c, x0, y0, s0, x, y, s, status, iters, resid, nullptr, 0, stg, stream);
}
extern "C" size_t bcone_cache_bytes(void *handle, int32_t B) {
  Handle *h = (Handle *)handle;
  if (!h || B <= 0 || !h->fast_fwd) return 0;
  return (size_t)B * bc_fwdf_cache_doubles(h->S.n, h->S.m) * sizeof(double);
}
extern "C" int bcone_solve(void *handle, int32_t B, const double *A_vals, const double *P_vals, const double *b,
                           const double *c, double *x, double *y, double *s, int32_t *status, int32_t *iters,
                           double *resid, const bcone_settings *stg, void *stream) {
  return bcone_solve_warm(handle, B, A_vals, P_vals, b, c, nullptr, nullptr, nullptr, x, y, s, status, iters, resid, stg, stream);
}
extern "C" int bcone_solve_cached(void *handle, int32_t B, const double *A_vals, const double *P_vals, const double *b, const double *c,
                                  const double *x0, const double *y0, const double *s0, double *x, double *y, double *s, int32_t *status,
                                  int32_t *iters, double *resid, void *cache, int32_t reuse, const bcone_settings *stg, void *stream) {
  Handle *h = (Handle *)handle;
  if (h && cache && !h->fast_fwd) return fail(h, BCONE_EINVAL, "solve: this structure has no cached set-up path (bcone_cache_bytes() is 0)");
  if (h && cache && ((uintptr_t)cache & 15)) return fail(h, BCONE_EINVAL, "solve: cache must be 16-byte aligned");
  if (h && ((x0 != nullptr) != (y0 != nullptr) || (x0 != nullptr) != (s0 != nullptr))) return fail(h, BCONE_EINVAL, "solve: warm start needs x0, y0 and s0 together");
  if (!h || B <= 0 || !A_vals || !b || !c || !x || !y || !s || !status || !iters || !stg) return fail(h, BCONE_EINVAL, "solve: null argument");
  if (h->S.nnzP > 0 && !P_vals) return fail(h, BCONE_EINVAL, "solve: structure has P but P_vals is NULL");
  if (stg->check_interval <= 0 || stg->max_iters <= 0) return fail(h, BCONE_EINVAL, "solve: check_interval and max_iters must be positive");
  if (stg->acceleration_lookback > BC_AA_MAXMEM || stg->acceleration_lookback < -BC_AA_MAXMEM)
    return fail(h, BCONE_EINVAL, "solve: |acceleration_lookback| must be <= 16");
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(h->device), "solve set device");
  FwdArgs a;
  a.S = h->S; a.B = B; a.A_vals = A_vals; a.P_vals = h->S.nnzP > 0 ? P_vals : nullptr; a.b = b; a.c = c;
  a.x = x; a.y = y; a.s = s; a.status = status; a.iters = iters; a.resid = resid; a.st = *stg;
  a.x0 = x0; a.y0 = y0; a.s0 = s0;
  a.cache = (double *)cache; a.cache_stride = h->fast_fwd ? (long long)bc_fwdf_cache_doubles(h->S.n, h->S.m) : 0; a.cache_reuse = cache && reuse;
  int *ctr = h->counters + 4 * (h->slot++ % Handle::RING);
  a.counter = ctr; a.use_tma = h->tma_ok && (((uintptr_t)A_vals & 15) == 0);
  // the 4-CTA/SM build only when the batch does not fit the resident capacity of the 128-register build
  const int use_small = h->fwd_small && (h->small_mode == 2 || B > h->num_sms * h->fwd_ctas);
  const int ctas = use_small ? h->fwd_ctas_small : h->fwd_ctas;
  const int grid = std::min(B, h->num_sms * ctas);
  const size_t max_grid = (size_t)h->num_sms * std::max(h->fwd_ctas, h->fwd_ctas_small);
  Handle::StreamWs *sw = stream_ws(h, st);
  a.ws = nullptr; a.ws_stride = (long long)h->fwd_ws_stride; a.prof = h->prof;
  if (h->fwd_indirect || h->fwd_factor_global) {
    if (!ensure_slab(h, &sw->fwd, nullptr, h->fwd_ws_stride * max_grid)) return fail(h, BCONE_ENOMEM, "cudaMalloc forward workspace");
    a.ws = sw->fwd;
  }
  a.aa_ws = nullptr; a.aa_stride = 0;
  if (stg->acceleration_lookback != 0 && stg->max_iters > 1) {
    const int mem = std::abs(stg->acceleration_lookback);
    a.aa_stride = (long long)((aa_ws_doubles(h->S.n + h->S.m + 1, mem) + 1) & ~(size_t)1);
    if (!ensure_slab(h, &sw->aa, &sw->aa_cap, (size_t)a.aa_stride * max_grid)) return fail(h, BCONE_ENOMEM, "cudaMalloc acceleration workspace");
    a.aa_ws = sw->aa;
  }
  a.park = nullptr;
  if (h->fast_fwd) {   // where the register tile waits while a termination check or an acceleration event runs (128 KB per CTA, L2)
    if (!ensure_slab(h, &sw->park, nullptr, (size_t)32 * 512 * max_grid)) return fail(h, BCONE_ENOMEM, "cudaMalloc tile parking slab");
    a.park = sw->park;
  }
  CK(cudaMemsetAsync(ctr, 0, sizeof(int), st), "solve counter");
  if (h->fast_fwd) CK(bc_fwdf_launch(&a, grid, h->fwd_smem, st), "solve launch (fast)");
  else CK(bc_fwd_launch(&a, h->fwd_indirect, grid, h->fwd_threads, h->fwd_smem, st, use_small), "solve launch");
  h->launches++;
  return BCONE_OK;
}

extern "C" int bcone_vjp(void *handle, int32_t B, const double *A_vals, const double *P_vals, const double *b,
                         const double *c, const double *x, const double *y, const double *s, const double *dx,
                         const double *dy, double *dA_vals, double *dP_vals, double *db, double *dc,
                         int32_t *lsqr_iters, const bcone_settings *stg, void *stream) {
  Handle *h = (Handle *)handle;
  if (!h || B <= 0 || !A_vals || !b || !c || !x || !y || !s || !dx || !dy || !dA_vals || !db || !dc || !stg)
    return fail(h, BCONE_EINVAL, "vjp: null argument");
  if (h->S.nnzP > 0 && !P_vals) return fail(h, BCONE_EINVAL, "vjp: structure has P but P_vals is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  BwdArgs a;
  a.S = h->S; a.B = B; a.A_vals = A_vals; a.P_vals = h->S.nnzP > 0 ? P_vals : nullptr; a.b = b; a.c = c;
  a.x = x; a.y = y; a.s = s; a.dx = dx; a.dy = dy; a.dA = dA_vals; a.dP = dP_vals; a.db = db; a.dc = dc;
  const int slot = h->slot++ % Handle::RING;
  int *ctr = h->counters + 4 * slot;
  a.lsqr_iters = lsqr_iters; a.st = *stg; a.counter = ctr + 1;
  a.use_tma = h->tma_ok && (((uintptr_t)A_vals & 15) == 0); a.psd_total = h->psd_total; a.p_in_smem = h->p_in_smem;
  CK(cudaSetDevice(h->device), "vjp set device");
  a.ws = nullptr; a.ws_stride = (long long)h->bwd_ws_stride;
  if (h->bwd_vec_global && !h->fast_bwd) {
    Handle::StreamWs *sw = stream_ws(h, st);
    if (!ensure_slab(h, &sw->bwd, nullptr, h->bwd_ws_stride * (size_t)h->num_sms * std::max(h->bwd_ctas, h->bwd_ctas_small))) return fail(h, BCONE_ENOMEM, "cudaMalloc backward workspace");
    a.ws = sw->bwd;
  }
  a.inst_list = nullptr; a.B_dev = nullptr; a.fail_list = nullptr; a.fail_count = nullptr; a.prof = h->prof;
  if (h->block_bwd && stg->lsqr_precond == 2) {
    // pass 1: block-preconditioned solve; pass 2: equilibrated LSQR on the instances it rejected
    if (h->fail_cap[slot] < B) {
      int *p = nullptr;
      CK(cudaMalloc((void **)&p, (size_t)B * sizeof(int)), "vjp fail list");
      h->allocs.push_back(p); h->fail_list[slot] = p; h->fail_cap[slot] = B;
    }
    CK(cudaMemsetAsync(ctr + 1, 0, 3 * sizeof(int), st), "vjp counters");
    h->last_block_slot = slot;
    a.fail_list = h->fail_list[slot]; a.fail_count = ctr + 2;
    CK(bc_bwdb_launch(&a, std::min(B, h->num_sms), h->blk_threads, h->blk_smem, st), "vjp launch (block)");
    BwdArgs f = a;
    f.st.lsqr_precond = 1; f.counter = ctr + 3; f.inst_list = h->fail_list[slot]; f.B_dev = ctr + 2;
    f.fail_list = nullptr; f.fail_count = nullptr;
    CK(bc_bwdf_launch(&f, std::min(B, h->num_sms * h->bwd_ctas), h->bwd_threads, h->bwd_smem, st), "vjp launch (fallback)");
    h->launches += 2;
    return BCONE_OK;
  }
  if (a.st.lsqr_precond == 2) a.st.lsqr_precond = 1;   // block factorisation not available for this structure
  CK(cudaMemsetAsync(ctr + 1, 0, sizeof(int), st), "vjp counter");
  const int use_small = h->bwd_small && (h->small_mode == 2 || B > h->num_sms * h->bwd_ctas);
  const int grid = std::min(B, h->num_sms * (use_small ? h->bwd_ctas_small : h->bwd_ctas));
  if (h->fast_bwd) CK(bc_bwdf_launch(&a, grid, h->bwd_threads, h->bwd_smem, st), "vjp launch (fast)");
  else CK(bc_bwd_launch(&a, grid, h->bwd_threads, h->bwd_smem, st, use_small), "vjp launch");
  h->launches++;
  return BCONE_OK;
}

// Strided host<->device copy on the caller's stream (cudaMemcpy2DAsync): lets the reference-facing
// call move a batch slice of the [rows, B] boundary tensors without a host-side repack.
extern "C" int bcone_memcpy2d(void *dst, int64_t dpitch, const void *src, int64_t spitch, int64_t width, int64_t height,
                              int32_t to_device, void *stream) {
  cudaError_t e = cudaMemcpy2DAsync(dst, (size_t)dpitch, src, (size_t)spitch, (size_t)width, (size_t)height,
                                    to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, (cudaStream_t)stream);
  if (e != cudaSuccess) { g_create_err = std::string("bcone_memcpy2d: ") + cudaGetErrorString(e); return BCONE_ECUDA; }
  return BCONE_OK;
}

// Debug: enable (on != 0) or read-and-reset the per-phase cycle counters of the forward / block-backward
// kernels.  out (HOST, 16 x uint64) may be NULL.  Phases: 0 load, 1 equilibration, 2 factorisation + g,
// 3 iterations, 4 checks | 8 load, 9 P factor, 10 W, 11 S, 12 S factor, 13 q + LSQR, 14 solve + write.
extern "C" int bcone_set_profile(void *handle, int32_t on, uint64_t *out) {
  Handle *h = (Handle *)handle;
  if (!h) return BCONE_EINVAL;
  cudaSetDevice(h->device);
  if (out && h->prof) { cudaDeviceSynchronize(); cudaMemcpy(out, h->prof, 32 * sizeof(uint64_t), cudaMemcpyDeviceToHost); cudaMemset(h->prof, 0, 32 * sizeof(uint64_t)); }
  if (on && !h->prof) { if (cudaMalloc((void **)&h->prof, 32 * sizeof(uint64_t)) != cudaSuccess) return BCONE_ENOMEM; h->allocs.push_back(h->prof); cudaMemset(h->prof, 0, 32 * sizeof(uint64_t)); }
  if (!on) h->prof = nullptr;
  return BCONE_OK;
}

// Instances of the last block-preconditioned bcone_vjp (lsqr_precond = 2) that the block factorisation rejected and the
// equilibrated LSQR solved instead.  Synchronises the device.  -1: no such call yet.
extern "C" int bcone_fallback_count(void *handle, int32_t *out) {
  Handle *h = (Handle *)handle;
  if (!h || !out) return BCONE_EINVAL;
  *out = -1;
  if (h->last_block_slot < 0) return BCONE_OK;
  cudaSetDevice(h->device);
  if (cudaDeviceSynchronize() != cudaSuccess) return BCONE_ECUDA;
  int v = 0;
  if (cudaMemcpy(&v, h->counters + 4 * h->last_block_slot + 2, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return BCONE_ECUDA;
  *out = v;
  return BCONE_OK;
}

extern "C" int64_t bcone_launch_count(void *handle) { return handle ? ((Handle *)handle)->launches : 0; }

extern "C" int bcone_path_info(void *handle, int32_t *fwd_path, int32_t *bwd_path) {
  Handle *h = (Handle *)handle;
  if (!h) return BCONE_EINVAL;
  if (fwd_path) *fwd_path = h->fast_fwd ? 2 : (h->fwd_indirect ? 1 : (h->fwd_factor_global ? 3 : 0));
  if (bwd_path) *bwd_path = h->block_bwd ? 2 : (h->fast_bwd ? 1 : 0);
  return BCONE_OK;
}

extern "C" int bcone_kernel_info(void *handle, int32_t *ft, int32_t *fs, int32_t *fc, int32_t *bt, int32_t *bs, int32_t *bcx) {
  Handle *h = (Handle *)handle;
  if (!h) return BCONE_EINVAL;
  if (ft) *ft = h->fwd_threads; if (fs) *fs = (int32_t)h->fwd_smem; if (fc) *fc = h->fwd_ctas;
  if (bt) *bt = h->bwd_threads; if (bs) *bs = (int32_t)h->bwd_smem; if (bcx) *bcx = h->bwd_ctas;
  return BCONE_OK;
}
