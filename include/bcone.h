/*
 * bcone.h -- C ABI of the B200-native batched cone-program solve-and-differentiate
 * engine (libbcone.so).  Plain pointers and sizes only; every data pointer is the
 * CALLER'S DEVICE MEMORY unless stated, every call is asynchronous on the given
 * CUDA stream, every function returns 0 on success and a negative code on error
 * (message via bcone_last_error); nothing throws across this boundary.
 *
 * What each entry point replaces in the reference (cvxpy/cvxpylayers @ f0b1c15):
 *   bcone_create   <- DIFFCP_ctx.__init__            src/cvxpylayers/interfaces/diffcp_if.py:105-120
 *                     (captures the sparsity structure + cone dims once per layer;
 *                      CSR twin: MOREAU_ctx.__init__ moreau_if.py:181-222)
 *   bcone_ingest   <- _build_diffcp_matrices         diffcp_if.py:46-70   (per-instance Python loop
 *                     turning the [nnz,B] boundary tensors into solver data A=-A_cvx, b, c)
 *   bcone_solve    <- diffcp.solve_and_derivative_batch / solve_only_batch
 *                                                    diffcp_if.py:365, :369 (SCS forward solve)
 *   bcone_vjp      <- the adjoint closure adj_batch  diffcp_if.py:86      (diffcp adjoint_derivative)
 *   bcone_emit     <- _compute_gradients re-packing  diffcp_if.py:88-94 + stacks :396-397
 *                     (dA_eval = [-dA ; db[b_idx]], dq_eval = [dc ; 0])
 *
 * Solver form (SCS / diffcp convention):  min 1/2 x'Px + c'x  s.t.  Ax + s = b, s in K,
 * dual y in K*.  K = zero(z) x nonneg(l) x SOC(q[0..nq)) x PSD(s[0..ns)) x exp(ep) x exp*(ed) in that
 * row order (exp triples (x,y,z): y e^{x/y} <= z).
 * All floating point data is fp64.
 */
#ifndef BCONE_H
#define BCONE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int32_t n, m, nnzA, nnzP;
  const int32_t *A_indptr, *A_indices; /* HOST pointers, CSR of A (m+1 / nnzA); copied        */
  const int32_t *P_indptr, *P_indices; /* HOST pointers, CSR upper triangle of P, or NULL     */
  int32_t z, l, nq, ns, ep, ed;        /* cone spec in SCS row order z, l, q, s, ep, ed        */
  const int32_t *q, *s;                /* HOST: SOC sizes [nq], PSD orders [ns]               */
  int32_t device;                      /* CUDA device ordinal                                 */
  int32_t max_batch;                   /* workspace is sized for this many instances          */
} bcone_desc;

typedef struct {
  double eps_abs, eps_rel, eps_infeas; /* SCS termination (defaults 1e-4, 1e-4, 1e-7)         */
  double alpha, rho_x, scale;          /* over-relaxation 1.5, 1e-6, 0.1                      */
  double lsqr_atol, lsqr_btol, lsqr_conlim; /* 1e-8, 1e-8, 1e8 (diffcp / SciPy LSQR rules)    */
  int32_t max_iters, normalize, adaptive_scale, check_interval;
  int32_t ruiz_passes, lsqr_iter_lim;  /* lsqr_iter_lim < 0 -> 2N like diffcp                 */
  int32_t lsqr_precond;                /* 0 plain LSQR (reference semantics), 1 diagonally equilibrated,
                                          2 KKT-block preconditioned where applicable (else 1) */
  int32_t adaptive_check;              /* 0: test termination every check_interval iterations (SCS);
                                          1: place the checks by log-linear extrapolation (<= check_interval apart) */
  int32_t acceleration_lookback;       /* SCS: Anderson acceleration window; 10 (type-I), < 0 type-II, 0 off
                                          (what the reference's tests pass, tests/test_torch.py:401-405); |.| <= 16 */
  int32_t acceleration_interval;       /* SCS: accelerate every this many iterations (10)                    */
} bcone_settings;

enum { BCONE_SOLVED = 1, BCONE_INACCURATE = 2, BCONE_UNBOUNDED = -1, BCONE_INFEASIBLE = -2, BCONE_FAILED = -4 };
enum { BCONE_OK = 0, BCONE_EINVAL = -1, BCONE_ECUDA = -2, BCONE_ENOMEM = -3, BCONE_EUNSUPPORTED = -4 };

void bcone_default_settings(bcone_settings *st);

int bcone_create(const bcone_desc *desc, void **handle);
void bcone_destroy(void *handle);
const char *bcone_last_error(void *handle); /* handle may be NULL: last create() error */

/* Boundary layout -> engine layout.  A_eval[nnz_aug, B], q_eval[n+1, B], P_eval[nnzP, B] are the
 * reference's batch-contiguous value matrices; gather[k] (HOST int32 [nnzA], given once at
 * bcone_set_boundary) is the row of A_eval feeding CSR slot k.  Outputs: A_vals[B,nnzA] = -A_eval,
 * b[B,m] (zeros off b_idx), c[B,n], P_vals[B,nnzP]. */
int bcone_set_boundary(void *handle, int32_t nnz_aug, const int32_t *gather, int32_t nb, const int32_t *b_idx);
/* Optional: P_eval[nnzP_boundary, B] rows in the reference's order for ANY symmetric pattern cvxpy emits (upper, lower or
 * full); gatherP[k] (HOST int32 [nnzP]) is the row feeding the engine's upper-triangular CSR slot k.  Rows that feed no
 * slot (the mirror entries of a full pattern) receive a zero gradient in bcone_emit.  Default: identity. */
int bcone_set_boundary_quad(void *handle, int32_t nnzP_boundary, const int32_t *gatherP);
int bcone_ingest(void *handle, int32_t B, const double *A_eval, const double *q_eval, const double *P_eval,
                 double *A_vals, double *P_vals, double *b, double *c, void *cuda_stream);
/* Engine gradients -> boundary layout: dA_eval[nnz_aug,B] = [-dA (boundary order) ; db[b_idx]],
 * dq_eval[n+1,B] = [dc ; 0], dP_eval[nnzP,B]. */
int bcone_emit(void *handle, int32_t B, const double *dA_vals, const double *dP_vals, const double *db,
               const double *dc, double *dA_eval, double *dq_eval, double *dP_eval, void *cuda_stream);

/* Same as bcone_ingest / bcone_emit on a COLUMN SLICE of the boundary tensors: the pointers address column lo of tensors whose
 * rows are ldb doubles apart (ldb = the full batch), B = hi - lo instances.  A sharded or pipelined caller reads its slice of
 * A_eval in place and writes its slice of dA_eval straight into the full gradient tensor (no staging copy, no concatenation). */
int bcone_ingest_pitched(void *handle, int32_t B, int64_t ldb, const double *A_eval, const double *q_eval, const double *P_eval,
                         double *A_vals, double *P_vals, double *b, double *c, void *cuda_stream);
int bcone_emit_pitched(void *handle, int32_t B, int64_t ldb, const double *dA_vals, const double *dP_vals, const double *db,
                       const double *dc, double *dA_eval, double *dq_eval, double *dP_eval, void *cuda_stream);

/* Peer exchange for a batch sharded over the GPUs of one node (one exchange of solutions + gradients).  The
 * rank that owns the autograd graph allocates the destination with bcone_peer_alloc and passes the 64-byte CUDA IPC handle
 * to the other ranks (any host channel: torch.distributed, MPI, a pipe); they map it with bcone_peer_open and push their
 * shard with bcone_copy2d_async -- peer-to-peer over NVLink on the copy engines, chunk by chunk behind the solve, so the
 * transfer takes no SM and hides behind the next chunk's kernels.  bcone_copy2d_async also serves local strided copies. */
int bcone_peer_alloc(int32_t device, int64_t bytes, void **ptr, void *ipc_handle64);
int bcone_peer_open(int32_t device, const void *ipc_handle64, void **ptr);
int bcone_peer_close(void *ptr);
int bcone_peer_free(void *ptr);
int bcone_copy2d_async(void *dst, int64_t dpitch, const void *src, int64_t spitch, int64_t width, int64_t height, void *cuda_stream);

/* Parameter -> matrix affine map fused into the load stage (replaces the reference's sparse x dense products around
 * the solver interface: forward  A_eval = A_param @ p_stack etc. at src/cvxpylayers/torch/cvxpylayer.py:443-451,
 * transposes at :33-37).  The three maps are HOST CSR matrices [rows x P1] (P1 = total parameter size + 1; the last row
 * of p_stack is the constant 1, torch/cvxpylayer.py:84-141), rows in BOUNDARY order: A map nnz_aug rows ([A_cvx values ; b
 * entries], exactly the rows of A_eval), q map n + 1 rows, P map nnzP rows or NULL.  Call after bcone_set_boundary.
 * bcone_ingest_params: p_stack[P1, B] (device, batch axis contiguous) -> A_vals[B,nnzA] = -A_eval, b, c, P_vals without
 *   materialising A_eval.  bcone_emit_params: engine gradients -> dp_stack[P1, B] = (the three maps)' applied to
 *   [-dA ; db[b_idx]], [dc ; 0], dP; the row of the constant is left 0.  Only parameters and parameter gradients have
 *   to cross PCIe (or NVLink, for a sharded batch) on this path. */
int bcone_set_param_maps(void *handle, int32_t P1, const int32_t *A_ptr, const int32_t *A_col, const double *A_val,
                         const int32_t *q_ptr, const int32_t *q_col, const double *q_val,
                         const int32_t *P_ptr, const int32_t *P_col, const double *P_val);
int bcone_ingest_params(void *handle, int32_t B, const double *p_stack, double *A_vals, double *P_vals, double *b,
                        double *c, void *cuda_stream);
int bcone_emit_params(void *handle, int32_t B, const double *dA_vals, const double *dP_vals, const double *db,
                      const double *dc, double *dp_stack, void *cuda_stream);

/* Layer prologue / epilogue on the device: what the reference does per call with expand / permute / reshape /
 * cat / transpose chains (_flatten_and_batch_params, src/cvxpylayers/torch/cvxpylayer.py:84-141) and with slices, Fortran
 * reshapes and a symmetric scatter (_recover_results, :225-282) are index maps, one launch each.  DEVICE pointers throughout
 * (maps int32, scales fp64), no handle.  op: 0 identity, 1 exp (GP variables), 2 log (GP parameters).
 *   rows_from_param: rows[k, b] = f(param[b * stride + map[k]]), k < K -- the K rows of p_stack[.., B] that one parameter owns
 *                    (stride 0 broadcasts an unbatched parameter; map = Fortran-order flattening, NULL = identity)
 *   param_from_rows: the adjoint (batch-sum for stride 0, x 1/p for log); gparam must be zeroed by the caller when stride = 0
 *   gather_cols    : out[b, k] = f(scale[k] * in[b * ld + map[k]]) -- one requested variable out of primal[B, n] / dual[B, m]
 *   scatter_cols   : the adjoint, accumulated into gin (zeroed by the caller) */
int bcone_rows_from_param(const double *param, int64_t stride, const int32_t *map, int32_t K, int32_t B, int32_t op, double *rows,
                          void *cuda_stream);
int bcone_param_from_rows(const double *grows, const double *param, int64_t stride, const int32_t *map, int32_t K, int32_t B,
                          int32_t op, double *gparam, void *cuda_stream);
int bcone_gather_cols(const double *in, int64_t ld, const int32_t *map, const double *scale, int32_t K, int32_t B, int32_t op,
                      double *out, void *cuda_stream);
int bcone_scatter_cols(const double *gout, const double *out, int64_t ld, const int32_t *map, const double *scale, int32_t K,
                       int32_t B, int32_t op, double *gin, void *cuda_stream);

/* Forward: instance-contiguous inputs A_vals[B,nnzA] (CSR order), P_vals[B,nnzP] or NULL, b[B,m], c[B,n];
 * outputs x[B,n], y[B,m], s[B,m], status[B], iters[B] (int32), resid[B,3] or NULL. */
int bcone_solve(void *handle, int32_t B, const double *A_vals, const double *P_vals, const double *b,
                const double *c, double *x, double *y, double *s, int32_t *status, int32_t *iters,
                double *resid, const bcone_settings *st, void *cuda_stream);

/* Warm-started forward (the reference has the API for one backend only, torch/cvxpylayer.py:464-487,
 * interfaces/moreau_if.py:237-256): x0[B,n], y0[B,m], s0[B,m] = a solution of a nearby problem (typically the previous call of
 * a training loop); the operator splitting starts at the fixed point that solution would be.  All three NULL = bcone_solve. */
int bcone_solve_warm(void *handle, int32_t B, const double *A_vals, const double *P_vals, const double *b, const double *c,
                     const double *x0, const double *y0, const double *s0, double *x, double *y, double *s,
                     int32_t *status, int32_t *iters, double *resid, const bcone_settings *st, void *cuda_stream);

/* Forward with a cached set-up (second half; the reference's template is the one-time `setup()` of
 * interfaces/moreau_if.py:237-256, `PA_is_constant`): the equilibration (D, E) and the factorisation (K^-1 at its final scale)
 * of every instance are kept in `cache` -- caller-owned device memory of bcone_cache_bytes(handle, B) bytes, 16-byte aligned.
 *   reuse = 0: solve as bcone_solve_warm and write the set-up;
 *   reuse = 1: the caller states that A_vals and P_vals are the ones of the call that wrote `cache` (same B, same order; b and c
 *              are free to change): the kernel skips the Ruiz passes, the formation of K, its Cholesky factorisation and
 *              inverse, and starts at the cached scale.  A record that was never completed (or was written with another
 *              rho_x) is rebuilt in place, so reuse = 1 on a fresh zero-filled buffer is safe.
 * An adaptive re-scaling inside a solve re-factorises as usual and refreshes the record.  Only the register-tiled dense kernel
 * (dense A, zero + nonneg rows, direct mode) has this path: bcone_cache_bytes returns 0 for every other structure, and
 * bcone_solve_cached then requires cache = NULL (it is bcone_solve_warm). */
size_t bcone_cache_bytes(void *handle, int32_t B);
int bcone_solve_cached(void *handle, int32_t B, const double *A_vals, const double *P_vals, const double *b, const double *c,
                       const double *x0, const double *y0, const double *s0, double *x, double *y, double *s,
                       int32_t *status, int32_t *iters, double *resid, void *cache, int32_t reuse,
                       const bcone_settings *st, void *cuda_stream);

/* Backward (stateless): adjoint of the solution map at (x,y,s) applied to (dx,dy), ds = 0.
 * Outputs dA_vals[B,nnzA] (every structural entry), dP_vals[B,nnzP] or NULL, db[B,m], dc[B,n],
 * lsqr_iters[B] or NULL. */
int bcone_vjp(void *handle, int32_t B, const double *A_vals, const double *P_vals, const double *b,
              const double *c, const double *x, const double *y, const double *s, const double *dx,
              const double *dy, double *dA_vals, double *dP_vals, double *db, double *dc,
              int32_t *lsqr_iters, const bcone_settings *st, void *cuda_stream);

/* Pitched host<->device copy on the caller's stream (bytes): moves a batch slice [rows, lo:hi] of a
 * boundary tensor directly between pinned host memory and a contiguous device chunk. */
int bcone_memcpy2d(void *dst, int64_t dpitch, const void *src, int64_t spitch, int64_t width, int64_t height,
                   int32_t to_device, void *cuda_stream);

/* Debug: per-phase cycle counters of the forward / block-backward kernels (see csrc/api.cu). */
int bcone_set_profile(void *handle, int32_t on, uint64_t *out16);

/* Introspection for benchmarks/tests: kernel launches issued by this handle so far, and the
 * launch geometry chosen for the forward / backward kernels. */
int64_t bcone_launch_count(void *handle);
/* instances of the last lsqr_precond = 2 bcone_vjp that fell back from the block factorisation to the equilibrated LSQR
 * (device synchronising; -1 before the first such call) */
int bcone_fallback_count(void *handle, int32_t *out);
int bcone_kernel_info(void *handle, int32_t *fwd_threads, int32_t *fwd_smem, int32_t *fwd_ctas_per_sm,
                      int32_t *bwd_threads, int32_t *bwd_smem, int32_t *bwd_ctas_per_sm);
/* Which kernels the structure selected.  fwd_path: 0 generic on-chip Cholesky (fwd.cu), 1 generic indirect (CG),
 * 2 register-tiled dense/polyhedral (fwd_fast.cu), 3 generic with the values on chip and the Cholesky factor + vectors in a
 * per-CTA slab of global memory (instances between the two; BCONE_FWD_MODE=indirect in the environment forces 1 instead).  bwd_path: 0 generic LSQR (bwd.cu), 1 fused single-pass LSQR
 * (bwd_fast.cu), 2 KKT-block preconditioned (bwd_block.cu, used when lsqr_precond = 2; falls back to 1 per instance). */
int bcone_path_info(void *handle, int32_t *fwd_path, int32_t *bwd_path);

#ifdef __cplusplus
}
#endif
#endif
