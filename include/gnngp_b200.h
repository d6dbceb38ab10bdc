/*
 * gnngp_b200.h — C ABI of the B200 (sm_100a) GNNGP hot path.
 *
 * The reference (niuzehao/gnn-gp) has no FFI: its hot path is Python calling ATen
 * (cuSPARSE / cuBLAS / cuSOLVER / TensorIterator kernels).  Each entry point below replaces the
 * ATen call(s) at the cited reference lines (paths relative to /root/reference); the Python
 * surface that binds them (gnn-gp_b200/{compute,predict,gnngp}.py) keeps the reference's function
 * names and signatures.  INTEGRATION.md shows the ctypes stub a maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless named host_*; the library never allocates or frees
 *     caller-visible device memory; scratch comes in through (workspace, workspace_bytes) where needed
 *   - matrices are fp32 row-major with an explicit leading dimension (elements, not bytes)
 *   - one call = kernels enqueued on `stream` (a cudaStream_t), no host synchronisation
 *   - return 0 on success, a negative gnngp_status otherwise; gnngp_last_error() has the text
 *   - there is no CPU fallback: without a CUDA device every compute call fails with GNNGP_ERR_CUDA
 */
#ifndef GNNGP_B200_H_
#define GNNGP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *gnngp_stream_t; /* cudaStream_t */

enum gnngp_status {
    GNNGP_OK = 0,
    GNNGP_ERR_ARG = -1,       /* bad size / null pointer / misaligned buffer */
    GNNGP_ERR_CUDA = -2,      /* a CUDA runtime / driver call failed */
    GNNGP_ERR_WORKSPACE = -3, /* workspace too small */
    GNNGP_ERR_UNSUPPORTED = -4
};

/* ---- library ---------------------------------------------------------------------------- */
int gnngp_abi_version(void);
const char *gnngp_last_error(void);
/* SM count and compute capability of the current device (host out-pointers). */
int gnngp_device_info(int *host_sm_count, int *host_cc_major, int *host_cc_minor);
/* Dense-contraction engine: 0 = fp32 CUDA-core tiles, 1 = tcgen05 3xTF32 with chunked accumulation
 * (TMA + TMEM; 128x256 tiles, two epilogue warpgroups), 2 = auto (tcgen05 when the shape fills the tile
 * grid).  Returns the previous setting. */
int gnngp_set_gemm_engine(int engine);
/* Number of kernels this library has launched since load (for bench.py's gpu_launches). */
int64_t gnngp_launch_count(void);
/* Number of launches of the tcgen05 / TMEM / TMA contraction kernel since load. */
int64_t gnngp_tc_launch_count(void);

/* ---- adjacency: COO -> CSR ---------------------------------------------------------------
 * Replaces the implicit coalesce ATen runs on the sparse operand of every `A @ X`
 * (src/compute.py:142,145,...; A built uncoalesced at src/datasets.py:328).  Entries are sorted by
 * (row, col) with a stable radix sort; duplicates stay adjacent and are summed by the SpMM.
 * rowptr: int64[n_rows+1], colidx: int32[nnz], vals: float[nnz]. */
size_t gnngp_csr_from_coo_workspace_bytes(int64_t nnz, int64_t n_rows);
int gnngp_csr_from_coo(const int64_t *row, const int64_t *col, const float *val, int64_t nnz,
                       int64_t n_rows, int64_t n_cols, int64_t *rowptr, int32_t *colidx, float *vals,
                       void *workspace, size_t workspace_bytes, gnngp_stream_t stream);

/* ---- K1: CSR SpMM  C[:, 0:k] = alpha * A @ B ----------------------------------------------
 * Replaces `sigma_w * A @ Q` (src/compute.py:142,145,164,169,186,205,209) and `A @ (A @ C0).T`
 * (src/compute.py:132,135,151,156,176,195,198).  B: [n_cols x k] ld ldb, C: [n_rows x k] ld ldc.
 * Default (tile_cols = 0 and a workspace of gnngp_spmm_workspace_bytes(n_cols, k)): B is copied once
 * into 64-column panels in the workspace and gathered from there.  tile_cols = 32/64/128 (or a null
 * workspace) gathers straight from the row-major B with that tile width.
 * nnz_per_warp: nonzeros per warp work item (0 = default: 512, or 1024 on graphs with >= 2^25 nonzeros).  Rows that straddle a work item are
 * combined with fp32 atomics, so those rows of C are zero-filled first (the call does that itself). */
size_t gnngp_spmm_workspace_bytes(int64_t n_cols, int64_t k);
/* Tuning flags of the SpMM kernels (results do not depend on them); returns the previous setting.  Default 7.
 * 1: rows written by a single work item leave with one 16-byte streaming store per lane (when C's rows are
 *    16-byte aligned) instead of four 4-byte stores;
 * 2: only rows that are accumulated with atomics (cut by a work-item border) or empty are zero-filled,
 *    instead of the whole C;
 * 4: graphs with fewer than 64 nonzeros per row on average take the short-row kernel of the row-major path
 *    (gathers issued four nonzeros at a time across row boundaries, 128-column tiles). */
int gnngp_spmm_set_flags(int flags);
int gnngp_spmm_csr_f32(const int64_t *rowptr, const int32_t *colidx, const float *vals,
                       int64_t n_rows, int64_t n_cols, int64_t nnz, const float *B, int64_t ldb,
                       int64_t k, float alpha, float *C, int64_t ldc, int tile_cols, int nnz_per_warp,
                       void *workspace, size_t workspace_bytes, gnngp_stream_t stream);

/* Row-sharded mode: the same product with the operand already laid out as 64-column panels
 * ([ceil(k/64)][n_cols][64] floats, gnngp_spmm_workspace_bytes(n_cols, k) bytes), and the kernel that builds
 * those panels by pulling each rank's [rows_per_rank x ld] row block out of peer memory over NVLink
 * (peer_ptrs_dev: device array of `world` device pointers — the gnngp_exchange_* buffers below, mapped by CUDA IPC).
 * Together they replace "NCCL all-gather of the factor rows, then A_r @ X" (SURVEY.md section 8e). */
int gnngp_panelize_peers_f32(const void *peer_ptrs_dev, int world, int64_t rows_per_rank, int64_t ld, int64_t n,
                             int64_t k, float *panels, gnngp_stream_t stream);
/* The panel copy on its own (local operand B: [n x k] ld ldb, any alignment), for callers that run several
 * products — e.g. the row blocks of a graph that is still being uploaded — against one operand. */
int gnngp_panelize_f32(const float *B, int64_t ldb, int64_t n, int64_t k, float *panels, gnngp_stream_t stream);
int gnngp_spmm_panels_f32(const int64_t *rowptr, const int32_t *colidx, const float *vals, int64_t n_rows,
                          int64_t n_cols, int64_t nnz, const float *panels, int64_t k, float alpha, float *C,
                          int64_t ldc, int nnz_per_warp, gnngp_stream_t stream);

/* ---- dense contractions  D = A * B^T  (A: [M x K] ld lda, B: [N x K] ld ldb) ----------------
 * Both operands K-contiguous ("NT").  The 3xTF32 engine needs a workspace for the hi/lo operand
 * split: query it with gnngp_gemm_workspace_bytes(M, N, K). */
size_t gnngp_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K);

/* plain: C = alpha * A B^T (+ beta * C).  Replaces `X @ X.T` (src/compute.py:25), `Q @ Q[mask].T`,
 * `K[~mask] @ V` (src/compute.py:16), `Q[mb].T @ Q[mb]` (src/predict.py:57), `v.T @ (...)`
 * (src/predict.py:30,60). */
int gnngp_gemm_nt_f32(const float *A, int64_t lda, const float *B, int64_t ldb, float *C, int64_t ldc,
                      int64_t M, int64_t N, int64_t K, float alpha, float beta,
                      void *workspace, size_t workspace_bytes, gnngp_stream_t stream);

/* K5: symmetric product C = A A^T (A: [n x K]); only the tiles on or below the diagonal are guaranteed to
 * be written (zero C first).  Replaces `Q[mb].T @ Q[mb]` (src/predict.py:57), whose only consumer is
 * torch.linalg.eigh (lower triangle). */
int gnngp_syrk_f32(const float *A, int64_t lda, float *C, int64_t ldc, int64_t n, int64_t K, void *workspace,
                   size_t workspace_bytes, gnngp_stream_t stream);

/* K2: landmark Gram with the ReLU arc-cosine expectation fused into the epilogue.
 * E[i,j] = J1(clamp((Q_i . L_j) / s_i / t_j)) * s_i * t_j,  J1(c) = (sin th + (pi - th) cos th)/(2 pi),
 * th = acos c.  Replaces the Gram + 12 elementwise passes of src/compute.py:124-126.
 * Q: [n x m] rows, L: [n_land x m] landmark rows, s: [n] row norms, t: [n_land] landmark norms. */
int gnngp_gram_arccos_f32(const float *Q, int64_t ldq, int64_t n, const float *L, int64_t ldl,
                          int64_t n_land, int64_t m, const float *s, const float *t, float *E,
                          int64_t lde, void *workspace, size_t workspace_bytes, gnngp_stream_t stream);

/* Initial kernels between the rows of X ([n x f]) and of R ([r x f]) — src/compute.py:24-35, 46-64.
 * kind: 0 linear, 1 rbf exp(-gamma |x-y|_2^2), 2 sigmoid tanh(gamma x.y + c), 3 polynomial (x.y + c)^d,
 * 4 arccos 1 - acos(clamp(x.y))/pi (inputs already normalised by the caller).
 * xx / rr: squared row norms (rbf only, else may be null). */
int gnngp_initial_kernel_f32(int kind, const float *X, int64_t ldx, int64_t n, const float *R, int64_t ldr,
                             int64_t r, int64_t f, const float *xx, const float *rr, float gamma, float c,
                             float d, float *C0, int64_t ldc, void *workspace, size_t workspace_bytes,
                             gnngp_stream_t stream);
/* laplacian exp(-gamma |x-y|_1): not a contraction, its own kernel (src/compute.py:28-29, 51-53). */
int gnngp_laplacian_kernel_f32(const float *X, int64_t ldx, int64_t n, const float *R, int64_t ldr, int64_t r,
                               int64_t f, float gamma, float *C0, int64_t ldc, gnngp_stream_t stream);

/* K6: posterior means for all nuggets in one contraction, scattered into fitted[G][n][C]:
 * fitted[g][i][c] = sum_j Q[i][j] * Bt[g*C + c][j].  Replaces the 101-iteration loop of
 * src/predict.py:31-32 / 61-62.  Bt: [(G*C) x m] ld ldb. */
int gnngp_nugget_sweep_f32(const float *Q, int64_t ldq, int64_t n, int64_t m, const float *Bt, int64_t ldb,
                           int64_t G, int64_t C, float *fitted, void *workspace, size_t workspace_bytes,
                           gnngp_stream_t stream);
/* S[(g*C + c)][j] = temp[j][c] / (w[j] + eps[g] * d[0])  — the `temp/(w+eps*d)` of src/predict.py:32,62
 * for every nugget; d is a device scalar.  temp: [m x C] ld ldt, S: [(G*C) x m] ld lds. */
int gnngp_nugget_coeff_f32(const float *temp, int64_t ldt, const float *w, const float *eps, const float *d,
                           int64_t m, int64_t C, int64_t G, float *S, int64_t lds, gnngp_stream_t stream);

/* Batched Cholesky substitution in fp64: X_b = (L_b L_b^T)^-1 B_b for b < batch, all in ONE launch.
 * L: lower factors, row-major, leading dimension ldl, l_stride elements between factors (the strict upper
 * triangle is not read); B: [m x nrhs] right-hand sides, b_stride between batches (0 = the same B for every
 * factor); X: [m x nrhs] solutions, ldx (a multiple of 8, >= nrhs rounded up to 8; rows 16-byte aligned; the
 * padding columns are scratch), x_stride.  X may not alias B.  Used by the nugget-parallel fit of the
 * sharded mode: per nugget it is the coefficient block `v @ (temp / (w + eps*d))` of src/predict.py:61-62,
 * i.e. (Q_b^T Q_b + eps d I)^-1 Q_b^T y_b, computed from a Cholesky factor instead of the replicated eigh. */
int gnngp_potrs_batched_f64(const double *L, int64_t ldl, int64_t l_stride, const double *B, int64_t ldb,
                            int64_t b_stride, double *X, int64_t ldx, int64_t x_stride, int64_t m, int64_t nrhs,
                            int64_t batch, gnngp_stream_t stream);

/* ---- row / column utilities -------------------------------------------------------------- */
/* out[i] = sum_j Q[i][j]^2 (take_sqrt = 0) or its square root (1) — src/compute.py:124, src/predict.py:58. */
int gnngp_row_sumsq_f32(const float *Q, int64_t ldq, int64_t n, int64_t m, int take_sqrt, float *out,
                        gnngp_stream_t stream);
/* out[0] = mean(v[0:n]) or, stride > 1, the mean of v[0], v[stride], ... (diagonal: stride = ld + 1). */
int gnngp_mean_f32(const float *v, int64_t n, int64_t stride, float *out, gnngp_stream_t stream);
/* M[i][col] = value for every row — the bias column `Q[:,-1] = sigma_b` (src/compute.py:142,145). */
int gnngp_fill_col_f32(float *M, int64_t ld, int64_t n, float value, gnngp_stream_t stream);
/* dst[r][:] = src[idx[r]][:]  — boolean-mask row gathers `Q[mask]`, `K[mask]` (src/compute.py:11,125). */
int gnngp_gather_rows_f32(const float *src, int64_t lds, const int64_t *idx, int64_t n_idx, int64_t ncols,
                          float *dst, int64_t ldd, gnngp_stream_t stream);
/* dst[idx[r]][:] = src[r][:] — `Q[mask] = ...` (src/compute.py:15). */
int gnngp_scatter_rows_f32(const float *src, int64_t lds, const int64_t *idx, int64_t n_idx, int64_t ncols,
                           float *dst, int64_t ldd, gnngp_stream_t stream);
/* dst[c][r] = src[idx[r]][c] (idx may be null = identity): `Q[mb].T`, `(A @ C0).T`, `K[:,mb]`. */
int gnngp_gather_rows_transposed_f32(const float *src, int64_t lds, const int64_t *idx, int64_t n_idx,
                                     int64_t ncols, float *dst, int64_t ldd, gnngp_stream_t stream);
/* dst[i][j] = src[i][idx[j]] — column gather `K[mb][:,mb]`, `K[:,mb]` (src/predict.py:27,32). */
int gnngp_gather_cols_f32(const float *src, int64_t lds, int64_t n, const int64_t *idx, int64_t n_idx,
                          float *dst, int64_t ldd, gnngp_stream_t stream);
/* out = a*x + b*y + c elementwise over an [n x m] block (y may be null); in-place allowed. */
int gnngp_affine_f32(float *out, int64_t ldo, const float *x, int64_t ldx, const float *y, int64_t ldy,
                     int64_t n, int64_t m, float a, float b, float c, gnngp_stream_t stream);

/* ---- Nystrom square root pieces (src/compute.py:6-17) ------------------------------------ */
/* root[j] = sqrt(max(D[j],0)); inv[j] = root[j] / max(D[j], 1e-4 * D[n-1])  (D ascending). */
int gnngp_nystrom_spectrum_f32(const float *D, int64_t n, float *root, float *inv, gnngp_stream_t stream);
/* out[i][j] = V[i][j] * d[j]   (transpose = 0)   or   out[j][i] = V[i][j] * d[j]   (transpose = 1). */
int gnngp_scale_cols_f32(const float *V, int64_t ldv, int64_t n, int64_t m, const float *d, int transpose,
                         float *out, int64_t ldo, gnngp_stream_t stream);

/* ---- exact path: g(K) on a dense kernel (src/compute.py:110-117) --------------------------
 * E[i][j] = J1(clamp(K[i][j] / s_i / s_j)) s_i s_j with s = sqrt(diag K); E may alias K. */
int gnngp_arccos_dense_f32(const float *K, int64_t ldk, int64_t n, float *E, int64_t lde, float *s_scratch,
                           gnngp_stream_t stream);

/* ---- targets and metrics (src/predict.py:23-25, 53-55, 66-90) ----------------------------- */
/* Yt[c][r] = (y[idx[r]] == c) — transposed one-hot of the training labels. */
int gnngp_onehot_t_f32(const int64_t *y, const int64_t *idx, int64_t n_idx, int64_t C, float *Yt, int64_t ldy,
                       gnngp_stream_t stream);
/* K7 classification: counts[g][s] = #{ i : membership[i] has bit s, argmax_c fitted[g][i][c] == y[i] }.
 * membership: uint8 bit set per node (bit s = node belongs to subset s, up to 8 subsets). counts int32 [G x 8],
 * zeroed by the call. */
int gnngp_argmax_count(const float *fitted, const int64_t *y, const uint8_t *membership, int64_t G, int64_t n,
                       int64_t C, int32_t *counts, gnngp_stream_t stream);
/* K7 regression: sse[g][s] = sum over subset s of sum_k (fitted[g][i][k] - y[i][k])^2 (double [G x 8]). */
int gnngp_sqerr_sum(const float *fitted, const float *y, const uint8_t *membership, int64_t G, int64_t n,
                    int64_t k, double *sse, gnngp_stream_t stream);

/* ---- peer exchange buffers for the fused all-gather + panel layout (sharded mode) ---------------------------
 * Plain device allocations that other ranks of the same node map through CUDA IPC, so that gnngp_panelize_peers_f32
 * can read every rank's row block over NVLink with ordinary loads.  `alloc` returns the device pointer (0 on failure)
 * and writes the 64-byte IPC handle to handle_out (host memory); `open` maps a peer's handle and returns the peer
 * pointer valid in this process; `copy2d` publishes a [rows x cols] fp32 block into such a buffer (stream-ordered). */
int64_t gnngp_exchange_alloc(size_t bytes, void *handle_out);
int64_t gnngp_exchange_open(const void *handle);
int gnngp_exchange_close(int64_t peer_ptr);
int gnngp_exchange_free(int64_t ptr);
int gnngp_copy2d_f32(const float *src, int64_t lds, int64_t rows, int64_t cols, int64_t dst_ptr, int64_t ldd,
                     gnngp_stream_t stream);

/* ---- K9: posterior solve for all nuggets by band reduction (csrc/bandfit.cu) ------------------------------------
 * Replaces `w, v = torch.linalg.eigh(Q[mb].T @ Q[mb])` and the per-nugget `v @ (temp / (w + eps*d))` of
 * src/predict.py:57-62.  gram: fp32 [m x m], LOWER triangle valid (gnngp_syrk_f32's output); rhs_t: fp32 [C x m] =
 * (Q_b^T y_b)^T; shifts: fp64 [G] on the device (eps_g * d); coeff_t: fp32 [(G*C) x m], row g*C + c =
 * ((gram + shifts[g] I)^-1 rhs)[:, c] — the B^T operand of gnngp_nugget_sweep_f32.  info (device int32): number of
 * shifts whose banded Cholesky met a non-positive pivot (the caller then takes the eigen route).  fp64 throughout:
 * successive band reduction to bandwidth 128 with block reflectors in GEMM form, one banded Cholesky + two
 * substitutions per shift.  C <= 48.  No host synchronisation. */
size_t gnngp_band_fit_workspace_bytes(int64_t m, int64_t C, int64_t G);
int gnngp_band_fit_f64(const float *gram, int64_t ldg, int64_t m, const float *rhs_t, int64_t ldr, int64_t C,
                       const double *shifts, int64_t G, float *coeff_t, int64_t ldo, int32_t *info, void *workspace,
                       size_t workspace_bytes, gnngp_stream_t stream);

/* ---- K10: fp64 dense operators of the Krylov Nystrom root (csrc/dense64.cu, gnn-gp_b200/krylov_root.py) ---------
 * Together they replace `D, V = torch.linalg.eigh(K[mask])` of src/compute.py:11 for large landmark blocks: the root
 * only needs the Cholesky factor of the block and the eigenpairs ABOVE the clamp D[-1]*1e-4 (src/compute.py:13), which
 * a block Lanczos process made of these calls delivers (host logic in krylov_root.py).  All matrices fp64 row-major.
 *   sym_f64_from_f32     a[n x n] = symmetric copy of the LOWER triangle of the fp32 block g
 *   dgemm_f64            C[m x n] = alpha A B + beta C on the fp64 tensor path; A(i,p) = a[i*a_rs + p*a_cs],
 *                        B(p,j) = b[p*b_rs + j*b_cs] (one unit stride per operand: plain or transposed views);
 *                        lower != 0 computes only the 128 x 64 tiles that touch the lower triangle (symmetric updates)
 *   potrf_f64            in-place blocked Cholesky factorisation a = L L^T (lower triangle read, strict upper triangle
 *                        zeroed on return); info (device int32) is incremented once per block that met a non-positive pivot
 *   potrf_inv_block_f64  one block of nb <= 128 rows: L in place and, when linv != NULL, L^-1 (full nb x nb, zeros above)
 *   residual_sumsq_f64   out[j] = sum_i (r[i][j] - theta[j] v[i][j])^2  (Ritz residuals |M y - theta y|^2)
 *   transpose_scale_f64  out[j][i] = alpha a[i][j];  scale_cols_f64  out[i][j] = v[i][j] d[j];  f64_to_f32  dst = (float)src
 *   qr_rotate_f64        x[rows x n] <- x Q_H with Q_H the orthogonal factor of the Householder QR of z [n x kp] (kp a
 *                        multiple of 128, z overwritten): the leading columns of x Q_H are +-(x z_j) for orthonormal z_j,
 *                        the others x times an orthonormal basis of the complement (csrc/bandfit.cu panel kernels)
 *   band_solve_f64       (band(h) + shifts[g] I) x = rhs for every shift: h symmetric [dim x dim], the band |i-j| <= 128 of
 *                        its lower triangle is used (projected block-tridiagonal matrices); rhs_t [C x dim] vectors as rows,
 *                        sol_t [(G*C) x dim]; one CTA per shift (the band fit's banded Cholesky kernel); C <= 48
 * No host synchronisation in any of them. */
size_t gnngp_band_solve_workspace_bytes(int64_t dim, int64_t C, int64_t G);
int gnngp_band_solve_f64(const double *h, int64_t ldh, int64_t dim, const double *rhs_t, int64_t ldr, int64_t C,
                         const double *shifts, int64_t G, double *sol_t, int64_t lds, int32_t *info, void *workspace,
                         size_t workspace_bytes, gnngp_stream_t stream);
size_t gnngp_qr_rotate_workspace_bytes(int64_t rows, int64_t n, int64_t kp);
int gnngp_qr_rotate_f64(double *x, int64_t ldx, int64_t rows, double *z, int64_t ldz, int64_t n, int64_t kp,
                        void *workspace, size_t workspace_bytes, gnngp_stream_t stream);
int gnngp_sym_f64_from_f32(const float *g, int64_t ldg, int64_t n, double *a, int64_t lda, gnngp_stream_t stream);
int gnngp_dgemm_f64(int64_t m, int64_t n, int64_t k, const double *a, int64_t a_rs, int64_t a_cs, const double *b,
                    int64_t b_rs, int64_t b_cs, double alpha, double beta, double *c, int64_t ldc, int lower,
                    gnngp_stream_t stream);
size_t gnngp_potrf_f64_workspace_bytes(int64_t n);
int gnngp_potrf_f64(double *a, int64_t lda, int64_t n, int32_t *info, void *workspace, size_t workspace_bytes,
                    gnngp_stream_t stream);
int gnngp_potrf_inv_block_f64(double *a, int64_t lda, int64_t nb, double *linv, int64_t ldi, int32_t *info,
                              gnngp_stream_t stream);
int gnngp_residual_sumsq_f64(const double *r, int64_t ldr, const double *v, int64_t ldv, const double *theta, int64_t n,
                             int64_t k, double *out, gnngp_stream_t stream);
int gnngp_transpose_scale_f64(const double *a, int64_t lda, int64_t rows, int64_t cols, double alpha, double *out,
                              int64_t ldo, gnngp_stream_t stream);
int gnngp_scale_cols_f64(const double *v, int64_t ldv, int64_t rows, int64_t cols, const double *d, double *out,
                         int64_t ldo, gnngp_stream_t stream);
int gnngp_f64_to_f32(const double *src, int64_t lds, int64_t rows, int64_t cols, float *dst, int64_t ldd,
                     gnngp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GNNGP_B200_H_ */
