/*
 * cagp_b200.h - C ABI of the B200-native CaGP hot path (libcagp_b200.so).
 *
 * Every entry point is stream-ordered and non-blocking: no device allocation, no host
 * synchronisation, no host read of device scalars.  The caller (XLA, torch, ...) owns every
 * buffer, including scratch.  Kernel hyper-parameters are passed as DEVICE pointers because
 * they are traced values under jit in the reference (operators/lazy_kernel.py:37-57 stores the
 * GPJax kernel object; its lengthscale / variance are JAX arrays).
 *
 * Return value: 0 on success, otherwise a cagp_status code; cagp_last_error_string() gives a
 * thread-local description.  Nothing throws across this boundary.
 *
 * Reference citations are relative to /root/reference/src/cagpjax.
 *
 * Data layout conventions
 *   X       row-major (n, d) inputs exactly as the reference holds them (cagp.py:91).
 *   xs      "scaled SoA" copy, shape (d, ld): xs[t*ld + i] = (X[i*d + t] - origin_t) * c_family / lengthscale_t.
 *           c_family folds the kernel's exponent constant and log2(e) so that the kernels
 *           evaluate 2^(-q) (RBF) or 2^(-sqrt q) (Matern) on q = sum_t (xs1_t - xs2_t)^2.
 *   Mt      the projected kernel matrix M' = K'(X1, X2) S stored k-major: Mt[b*ld + i] = M'[i, b]
 *           (K' is the unit-variance kernel; the caller multiplies by `variance`).
 *   blocks  BlockDiagonalSparse layout (operators/block_diagonal_sparse.py:49-52): block_size =
 *           n / k, the last block absorbs n % k.
 */
#ifndef CAGP_B200_H_
#define CAGP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* cagp_stream_t; /* cudaStream_t */

enum cagp_status {
  CAGP_OK = 0,
  CAGP_ERR_INVALID = 1,   /* bad argument (shape, enum, null pointer) */
  CAGP_ERR_UNSUPPORTED = 2,
  CAGP_ERR_CUDA = 3,      /* a CUDA runtime call / launch failed */
  CAGP_ERR_CUBLAS = 4,
  CAGP_ERR_WORKSPACE = 5  /* workspace too small */
};

enum cagp_family { CAGP_RBF = 0, CAGP_MATERN12 = 1, CAGP_MATERN32 = 2, CAGP_MATERN52 = 3 };
enum cagp_dtype { CAGP_F32 = 0, CAGP_F64 = 1 };

/* Stationary kernel descriptor; replaces the GPJax kernel object held by LazyKernel
 * (operators/lazy_kernel.py:37-57).  `variance` is not used by the kernels (unit variance). */
typedef struct cagp_kernel_desc {
  int32_t family;            /* cagp_family */
  int32_t dtype;             /* cagp_dtype of every floating-point buffer in the call */
  int32_t n_lengthscale;     /* 1 (isotropic) or d (ARD) */
  int32_t reserved;
  const void* lengthscale;   /* device pointer, n_lengthscale elements */
  const void* bbox;          /* optional (may be NULL): device pointer to the bounding box of ALL scaled points
                                of the call, {min_0..min_{D-1}, max_0..max_{D-1}} with D = cagp_padded_dim(d)
                                (cagp_bounding_box).  Lets the kernels prove launch-wide (a) that no pair can reach
                                the exp2 clamp (3 integer instructions less per evaluation) and (b), for fp64 RBF
                                with every |x|^2 < 300, that the factored form K' = 2^(-|a|^2) 2^(-|b|^2) 2^(2 a.b)
                                is accurate (D instead of 2 D fp64 slots per distance); results agree with the
                                general path to ~3e-14 relative. */
} cagp_kernel_desc;

const char* cagp_last_error_string(void);
int cagp_version(void);

/* Coordinate dimension the kernels are instantiated for: d is zero-padded up to this value
 * (1, 2, 3, 4, 8, 16, 32); returns -1 if d > 32.  xs buffers have this many rows. */
int cagp_padded_dim(int32_t d);

/* xs = scale_and_transpose(X - origin).  Replaces `x / lengthscale` inside the GPJax kernel call made
 * at operators/lazy_kernel.py:83-85.  X (n, d) row-major; xs (cagp_padded_dim(d), ld), ld >= n; padding
 * rows are zeroed.  `origin` (device pointer, d elements, or NULL) is subtracted before scaling: stationary
 * kernels only see differences, and removing a common offset keeps those differences accurate for data
 * far from zero.  Both operands of a product must use the same origin. */
int cagp_scale_inputs(const cagp_kernel_desc* kd, const void* X, int64_t n, int32_t d,
                      const void* origin, void* xs, int64_t ld, cagp_stream_t stream);

/* Bounding box of scaled inputs xs (dpad, ld): bbox[t] = min_i xs[t][i], bbox[dpad + t] = max_i xs[t][i]. */
int cagp_bounding_box(int32_t dtype, const void* xs, int64_t n, int64_t ld, int32_t dpad, void* bbox,
                      cagp_stream_t stream);

/* Mt = (K'(X1, X2) S)^T for a BlockDiagonalSparse S with non-zeros `s` (n) and k blocks.
 * Replaces LazyKernel._matmat applied to to_dense(BlockDiagonalSparse)
 * (operators/lazy_kernel.py:79-90 x operators/block_diagonal_sparse.py:48-70) without
 * materialising K or the dense S.  xs1 (d, ld1) has m points, xs2 (d, ld2) has n points.
 * Mt (k, ldm), ldm >= m. */
int cagp_ks_blocksparse_fwd(const cagp_kernel_desc* kd, const void* xs1, int64_t m, int64_t ld1,
                            const void* xs2, int64_t n, int64_t ld2, int32_t d, const void* s,
                            int32_t k, void* Mt, int64_t ldm, cagp_stream_t stream);

/* Reverse mode of cagp_ks_blocksparse_fwd with respect to the lengthscale(s) and s, given
 * Gt[b*ldg + i] = dL/dM'[i, b] (k, ldg):
 *   s_bar[j]   = sum_i K'(x1_i, x2_j) Gt[blk(j)][i]
 *   ls_bar[t]  = sum_ij Gt[blk(j)][i] s_j dK'(x1_i, x2_j)/d lengthscale_t
 * This is what JAX reverse mode produces through lax.map + checkpoint in the reference
 * (operators/lazy_kernel.py:88-89; the reference has no custom VJP).  K' is recomputed, never
 * stored.  `ws` is scratch of cagp_ks_bwd_ws_bytes() bytes.  s_bar (n), ls_bar (n_lengthscale)
 * are overwritten. */
size_t cagp_ks_bwd_ws_bytes(const cagp_kernel_desc* kd, int64_t m, int64_t n, int32_t d, int32_t k);
int cagp_ks_blocksparse_bwd(const cagp_kernel_desc* kd, const void* xs1, int64_t m, int64_t ld1,
                            const void* xs2, int64_t n, int64_t ld2, int32_t d, const void* s,
                            int32_t k, const void* Gt, int64_t ldg, void* s_bar, void* ls_bar,
                            void* ws, size_t ws_bytes, cagp_stream_t stream);

/* Symmetric variants for the Gram case X1 == X2 (kernel.gram, computations/lazy.py:88-106): each
 * K'(x_i, x_j) is evaluated once per unordered pair of action blocks {a, b} and feeds both
 * M'[i, b] and M'[j, a].  Tiles are (a, (a + off) mod k), off = 0 .. k/2; [a_begin, a_end) selects the
 * action blocks a whose tiles this call computes (all blocks: 0, k; a rank's share otherwise).
 * Mt (k, ldm) with ldm >= n holds ALL n columns: the call writes Mt[b][i in a] and Mt[a][j in b] for
 * its tiles and leaves the other entries untouched (they belong to other ranks' tiles). */
int cagp_ks_blocksparse_fwd_sym(const cagp_kernel_desc* kd, const void* xs, int64_t n, int64_t ld,
                                int32_t d, const void* s, int32_t k, int32_t a_begin, int32_t a_end,
                                void* Mt, int64_t ldm, cagp_stream_t stream);

/* Peer-memory variant of the symmetric forward for one-process-per-GPU runs (NVLink / NVSwitch): the
 * compute and the exchange are ONE kernel.  Rank r owns the rows [row_begin[r], row_begin[r+1]) (aligned
 * to action blocks) and a panel (k, ldm) for them; `panels` lists all ranks' panels as pointers valid on
 * THIS device (peer-mapped, e.g. symmetric memory).  The tiles of blocks [a_begin, a_end) are computed
 * here; their row sums are stored into panels[me], their column sums directly into the panel of the rank
 * that owns the column block (remote stores), which replaces the all-to-all that otherwise follows
 * cagp_ks_blocksparse_fwd_sym.  Panels whose column sums are accumulated atomically (blocks split over
 * several CTAs, i.e. row k-1) must be zeroed by their owners, and a cross-rank barrier must separate
 * zeroing, this call and the first use of the panels.  `panels` / `row_begin` are HOST arrays. */
int cagp_ks_blocksparse_fwd_sym_peer(const cagp_kernel_desc* kd, const void* xs, int64_t n, int64_t ld,
                                     int32_t d, const void* s, int32_t k, int32_t a_begin, int32_t a_end,
                                     void* const* panels, const int64_t* row_begin, int32_t world,
                                     int32_t me, int64_t ldm, cagp_stream_t stream);

/* Introspection (host only, no launch): 1 if the symmetric forward (backward != 0: backward) call with this
 * descriptor and shape also launches the float64-RBF fast-path kernel (csrc/ks_symx.cuh), which does the work
 * whenever the bounding box in `kd` proves every scaled |x|^2 < 900 (decided on the device; below 300 without, above
 * with an exponent clamp); else 0.  Forward: isotropic or ARD lengthscales; backward: isotropic only.  Tests and
 * bench.py use it to name the kernel instance they exercise. */
int cagp_ks_sym_fast_path(const cagp_kernel_desc* kd, int64_t n, int32_t d, int32_t k, int32_t backward);

/* Host-only: which rows of a panel the symmetric forward ACCUMULATES into (atomics, because a row block is split over
 * several CTAs) and which the caller of the peer variant must therefore zero before the cross-rank barrier:
 * 0 = none, 1 = row k-1 only (the overhang block), 2 = the whole panel. */
int cagp_ks_fwd_sym_zero_rows(const cagp_kernel_desc* kd, int64_t n, int32_t d, int32_t k);

/* Reverse mode of the symmetric forward for the same tiles, given the FULL cotangent Gt (k, ldg >= n):
 * s_bar (n) and ls_bar (n_lengthscale) receive this call's tiles' contributions (s_bar is zeroed
 * first and accumulated with atomics: sums agree to rounding, not bitwise, between runs). */
size_t cagp_ks_bwd_sym_ws_bytes(const cagp_kernel_desc* kd, int64_t n, int32_t d, int32_t k,
                                int32_t a_begin, int32_t a_end);
int cagp_ks_blocksparse_bwd_sym(const cagp_kernel_desc* kd, const void* xs, int64_t n, int64_t ld,
                                int32_t d, const void* s, int32_t k, int32_t a_begin, int32_t a_end,
                                const void* Gt, int64_t ldg, void* s_bar, void* ls_bar, void* ws,
                                size_t ws_bytes, cagp_stream_t stream);

/* A'[a][b] = sum_{i in block a} s_i M'[i, b]  (k, k) row-major and c'[b] = sum_i M'[i, b] yt_i,
 * for the m rows held by the caller: global row index of local row i is row_offset + i (row
 * sharding); n_total is the global n that defines the block layout.  Replaces
 * BlockDiagonalSparse._rmatmat (operators/block_diagonal_sparse.py:72-95) applied to K S and the
 * contraction (K S)^T (y - mean) used by predict (models/cagp.py:160-163). */
int cagp_project_rows(int32_t dtype, const void* Mt, int64_t ldm, int64_t m, int64_t row_offset,
                      int64_t n_total, int32_t k, const void* s, const void* yt, void* A, void* c,
                      cagp_stream_t stream);

/* B' = M'^T M' (k, k), full symmetric matrix.  Replaces the lazy inv_congruence_transform / cola.diag chain
 * (solvers/cholesky.py:57-63, distributions.py:53-56) whose trace only needs this k x k Gram matrix.
 * float64: a hand-written fp64 tensor-core kernel (mma.sync.m8n8k4.f64) that computes only the lower-triangular
 * 128 x 128 tiles (n k^2 flops), splits the m dimension over CTAs and reduces the partial tiles from `ws`
 * (cagp_gram_ws_bytes() bytes, caller-owned) in a fixed order; float32: cuBLAS Sgemm, ws may be NULL. */
size_t cagp_gram_ws_bytes(int32_t dtype, int64_t m, int32_t k);
int cagp_gram_mtm(int32_t dtype, const void* Mt, int64_t ldm, int64_t m, int32_t k, void* B, void* ws,
                  size_t ws_bytes, cagp_stream_t stream);

/* Gt = W Mt + outer terms:  Gt[b][i] = sum_c W[b][c] Mt[c][i] + Abar[blk(i)][b] s_i + cbar[b] yt_i.
 * W (k, k) row-major (the caller passes Bbar + Bbar^T).  Cotangent of M' through A', B', c'.  float64: one fp64
 * tensor-core kernel with the outer terms as its epilogue; float32: prefill kernel + cuBLAS Sgemm. */
int cagp_assemble_cotangent(int32_t dtype, const void* Mt, int64_t ldm, int64_t m, int64_t row_offset,
                            int64_t n_total, int32_t k, const void* W, const void* Abar,
                            const void* cbar, const void* s, const void* yt, void* Gt, int64_t ldg,
                            cagp_stream_t stream);

/* Peer-memory variant of cagp_assemble_cotangent for one-process-per-GPU runs (float64): compute and exchange are ONE
 * kernel.  Every rank h holds a full-width buffer buffers[h] (k, ld >= n_total) indexed by GLOBAL row; this rank's panel
 * (global rows [row_offset, row_offset + m)) is stored into its own buffer and row b additionally - remote st.global
 * over NVLink / NVSwitch, from the GEMM epilogue - into the buffer of the rank that owns action block b
 * (block_begin[h] <= b < block_begin[h+1]; `buffers`, `block_begin` are HOST arrays of world resp. world + 1 entries,
 * pointers valid on THIS device, e.g. symmetric memory).  Afterwards buffers[me] holds what
 * cagp_ks_blocksparse_bwd_sym needs for this rank's tiles (G'[i in own rows, :] and G'[:, a in own blocks]), which
 * replaces the all-to-all of the cotangent.  A cross-rank barrier must separate the previous use of the buffers, this
 * call, and the backward kernel. */
int cagp_assemble_cotangent_peer(int32_t dtype, const void* Mt, int64_t ldm, int64_t m, int64_t row_offset,
                                 int64_t n_total, int32_t k, const void* W, const void* Abar, const void* cbar,
                                 const void* s, const void* yt, void* const* buffers, const int32_t* block_begin,
                                 int32_t world, int32_t me, int64_t ld, cagp_stream_t stream);

/* Direct (non-kernel) cotangents of the projections for the caller's m rows:
 *   s_dir[i] = sum_b Abar[blk(i)][b] M'[i, b],   yt_bar[i] = sum_b cbar[b] M'[i, b]. */
int cagp_project_rows_bwd(int32_t dtype, const void* Mt, int64_t ldm, int64_t m, int64_t row_offset,
                          int64_t n_total, int32_t k, const void* Abar, const void* cbar, void* s_dir,
                          void* yt_bar, cagp_stream_t stream);

/* Predictive moments at the m points behind Mt = (K'(X*, X) S)^T (models/cagp.py:160-167,
 * distributions.py:48-56):
 *   mean[i] = mean_const + variance * sum_b M'[i, b] w[b]
 *   var[i]  = variance * kdiag - variance^2 * sum_{b,c} M'[i, b] Pinv[b][c] M'[i, c]
 * Pinv (k, k) row-major symmetric; `scalars` is a device array {variance, mean_const}.
 * `ws`: cagp_predict_ws_bytes() bytes of scratch (float64: 2 x ceil(k/128) x m partial sums - the product Pinv Mt is
 * contracted inside the tensor-core kernel and never written; float32: a (k, m) panel for cuBLAS Sgemm). */
size_t cagp_predict_ws_bytes(int32_t dtype, int64_t m, int32_t k);
int cagp_predict_moments(int32_t dtype, const void* Mt, int64_t ldm, int64_t m, int32_t k,
                         const void* w, const void* Pinv, const void* scalars, void* mean, void* var,
                         void* ws, size_t ws_bytes, cagp_stream_t stream);

/* Dense right-hand side: Y = K'(X1, X2) V with V (n, p) row-major (ldv >= p), Y (m, p) row-major
 * (ldy >= p).  This is LazyKernel._matmat for a dense operand (operators/lazy_kernel.py:79-90):
 * matvecs (p = 1) and dense action matrices (cola.ops.Dense actions through the congruence fallback
 * linalg/congruence.py:15-17).  X @ K (lazy_kernel.py:92-103) is the same call with
 * xs1 / xs2 swapped, since K'(x, y) = K'(y, x). */
int cagp_ks_dense_fwd(const cagp_kernel_desc* kd, const void* xs1, int64_t m, int64_t ld1,
                      const void* xs2, int64_t n, int64_t ld2, int32_t d, const void* V, int32_t p,
                      int64_t ldv, void* Y, int64_t ldy, cagp_stream_t stream);

/* Lengthscale part of the reverse mode of cagp_ks_dense_fwd given the cotangent Ybar (m, p):
 *   ls_bar[t] = sum_ij (Ybar_i . V_j) dK'(x1_i, x2_j) / d lengthscale_t.
 * Both dense operands are passed TRANSPOSED (point index contiguous): Vt (p, ldv >= n) and
 * Ybt (p, ldyb >= m), so that tile loads are coalesced.
 * (The other cotangent, V_bar = K'^T Ybar, is cagp_ks_dense_fwd with xs1 / xs2 swapped.) */
size_t cagp_ks_dense_bwd_ls_ws_bytes(const cagp_kernel_desc* kd, int64_t m, int64_t n, int32_t d);
int cagp_ks_dense_bwd_ls(const cagp_kernel_desc* kd, const void* xs1, int64_t m, int64_t ld1,
                         const void* xs2, int64_t n, int64_t ld2, int32_t d, const void* Vt, int64_t ldv,
                         const void* Ybt, int64_t ldyb, int32_t p, void* ls_bar, void* ws,
                         size_t ws_bytes, cagp_stream_t stream);

/* Materialised unit-variance cross-covariance C[i*ldc + j] = K'(x1_i, x2_j), (m, ldc >= n) row-major.  Replaces
 * kernel.cross_covariance(train_inputs, pseudo_inputs) as called by PseudoInputPolicy.to_actions
 * (policies/pseudoinput.py:68-74): the result is the dense action matrix fed to cagp_ks_dense_fwd. */
int cagp_cross_cov_fwd(const cagp_kernel_desc* kd, const void* xs1, int64_t m, int64_t ld1, const void* xs2,
                       int64_t n, int64_t ld2, int32_t d, void* C, int64_t ldc, cagp_stream_t stream);

/* Reverse mode of cagp_cross_cov_fwd given Cbar (m, ldc), with respect to the x2 points (trainable
 * pseudo-inputs) and the lengthscale(s):
 *   x2s_bar[t*n + j] = sum_i Cbar_ij dK'(x1_i, x2_j) / d xs2[t][j]   (cagp_padded_dim(d), n), cotangent of the SCALED
 *                      coordinates: the caller multiplies row t by c_family / lengthscale_t to get dL/dX2[j][t];
 *   ls_bar[t]        = sum_ij Cbar_ij dK'(x1_i, x2_j) / d lengthscale_t   (through both operands).
 * Deterministic (per-chunk partials in `ws`, cagp_cross_cov_bwd_ws_bytes() bytes, then one reduction). */
size_t cagp_cross_cov_bwd_ws_bytes(const cagp_kernel_desc* kd, int64_t m, int64_t n, int32_t d);
int cagp_cross_cov_bwd(const cagp_kernel_desc* kd, const void* xs1, int64_t m, int64_t ld1, const void* xs2,
                       int64_t n, int64_t ld2, int32_t d, const void* Cbar, int64_t ldc, void* x2s_bar,
                       void* ls_bar, void* ws, size_t ws_bytes, cagp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CAGP_B200_H_ */
