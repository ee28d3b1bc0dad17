/* cpsd_b200 -- C ABI of the B200 (sm_100a) kernels behind the cross-validated
 * align -> reduce -> decode loop of coganlab/cross_patient_speech_decoding.
 *
 * The reference is pure Python: its "interface" for this path is the sklearn-style class
 * surface (AlignCCA / AlignMCCA / JointPCA / DimRedReshape / NoCenterPCA / crossPtDecoder_*),
 * mirrored in Python by the package cross_patient_speech_decoding_b200.  Those classes and
 * the batched engine call ONLY the functions below (through ctypes, see _lib.py); each
 * entry cites the reference code whose arithmetic it replaces (paths relative to
 * aligned_decoding/ in the reference tree).
 *
 * Conventions: plain C, device pointers unless named *_host, row-major fp32, explicit
 * leading dimensions and batch strides (in elements), per-problem sizes optionally read
 * from device int arrays (`*_dev`, may be NULL -> the `*_fixed` scalar applies), every
 * function is asynchronous on `stream` and returns 0 on success (CPSD_OK) or a non-zero
 * status with a message retrievable through cpsd_last_error().  No global state except
 * the error string (thread local) and a launch counter.
 */
#ifndef CPSD_B200_H
#define CPSD_B200_H

#include <cuda_runtime_api.h>
#include "../cross_patient_speech_decoding_b200/csrc/descs.h"

#ifdef __cplusplus
extern "C" {
#endif

#define CPSD_OK 0
#define CPSD_ERR_INVALID 1
#define CPSD_ERR_CUDA 2
#define CPSD_ERR_UNSUPPORTED 3

/* ---- library ------------------------------------------------------------------- */
int cpsd_version(void);
int cpsd_device_arch(void);                 /* major*10+minor of the current device */
const char* cpsd_last_error(void);
long long cpsd_launch_count(void);          /* kernels launched by this library so far */
void cpsd_reset_launch_count(void);
int cpsd_desc_sizes(int* out_host);         /* sizeof of the 7 descriptor records */

/* ---- streaming stages ------------------------------------------------------------
 * cnd_avg / extract_group_conditions (alignment/alignment_utils.py:42-61, 12-39): mean of
 * the trials of each class, trials kept in [trial][time][channel] layout. */
int cpsd_class_mean(const cpsd_class_mean_desc* descs, int nprob, int nslot_max, int TC,
                    cudaStream_t stream);
/* column means over selected row segments (centering of AlignCCA.py:259-260, PCA mean_,
 * mvlearn MCCA means_) */
int cpsd_colsum(const cpsd_colsum_desc* descs, int nprob, int p_max, cudaStream_t stream);
int cpsd_center_rows(float* Z, int ld, long long strideZ, const float* mu, int ldmu,
                     const int* nrows_dev, int nrows_fixed, int nrows_max, int ncols, int nprob,
                     cudaStream_t stream);
int cpsd_copy_rows(const float* src, int lds, long long strideS, float* dst, int ldd,
                   long long strideD, const int* r0_dev, int r0_fixed, int nrows, int ncols,
                   int src_rows, int nprob, cudaStream_t stream);
/* out[p] = base + sign * sum of the listed per-trial scatter matrices (fp64): the train-set
 * Gram of AlignMCCA.n_components_var (AlignMCCA.py:156-174) as all-trials Gram minus the
 * held-out trials' Grams */
int cpsd_sum_mats_f64(const double* base, const double* mats, long long mat_stride,
                      const int* list_ptr, const int* list, double sign, double* out,
                      long long out_stride, int elems, int nprob, cudaStream_t stream);
/* per-trial column sums (fp64) and the centred covariance of a trial subset from per-trial
 * statistics, cov = (G - s s^T / n) / (n - 1), mu = s_mu / n (s_mu NULL: s; normalize 0: the centred scatter without the 1 / (n - 1)): sklearn PCA's covariance of a fold's
 * train trials (decoders/cross_pt_decoders.py:234-241 -> PCA.fit) without re-reading the trials */
int cpsd_trial_colsum_f64(const float* X, int n_trials, int T, int C, int ldx, double* sums, int lds,
                          cudaStream_t stream);
int cpsd_cov_from_sums(double* G, int ldg, long long strideG, const double* s, int lds,
                       const int* nrows_dev, int C, float* mu, int ldmu, const double* s_mu,
                       int normalize, int nprob, cudaStream_t stream);
/* electrode subsampling of resident trials: dst = src[:, idx] (the channel lists of
 * processing_utils/grid_subsampling.py:8-61 and poisson_disk_sampling.py:9-77) and the block
 * means of spatial_avg_data (processing_utils/spatial_avg_subsampling.py:74-96) */
int cpsd_gather_channels(const float* src, int lds, const int* idx, int nidx, float* dst, int ldd,
                         long long nrows, cudaStream_t stream);
/* zero the columns j >= k_dev[p / group] of every problem's (rows x cols) matrix: JointPCA read-in
 * matrices when n_components is a variance fraction (alignment/JointPCA.py:199 with the script's
 * n_comp = 0.9, scripts/aligned_decode_svm_ncv.py:186-190) */
int cpsd_mask_cols(float* M, int ld, long long stride, int rows, int cols, const int* k_dev, int group,
                   int nprob, cudaStream_t stream);
/* dst[i] = src[idx[i]] over the trial axis (rows of TC = time x channels floats): the random
 * trial subsets of scripts/aligned_decode_cross_patient_subsample.py:303-312 on resident data */
int cpsd_gather_trials(const float* src, long long TC, const int* idx, int n, float* dst,
                       cudaStream_t stream);
/* alignment/metrics.py:41-68 pt_corr: Pearson r of every condition's flattened (time x feature)
 * block, a / b: (nrows x len) fp64 */
int cpsd_pearson_rows(const double* a, const double* b, int nrows, long long len, double* r,
                      cudaStream_t stream);
int cpsd_region_mean_f64(const double* data, int ntrials, int nelec, int T, const int* reg_ptr,
                         const int* reg_elec, int nreg, double* out, cudaStream_t stream);
/* ingest: the reference keeps trials as float64 (pickled numpy); cast once on the device */
int cpsd_cast_f64_f32(const double* src, float* dst, long long n, cudaStream_t stream);
int cpsd_permute_cols(const float* src, int lds, long long strideS, const int* perm, int ld_perm,
                      float* dst, int ldd, long long strideD, int nrows, int ncols, int nprob,
                      cudaStream_t stream);

/* ---- covariance / projection GEMMs -------------------------------------------------
 * X^T X, X^T Y over trial- or class-selected rows: sklearn PCA covariance_eigh reached from
 * decoders/cross_pt_decoders.py:234-241; AlignMCCA.n_components_var (alignment/AlignMCCA.py:
 * 156-174); scatter blocks of CCA_align (alignment/AlignCCA.py:235-285) and of the MCCA
 * GEVP (AlignMCCA.py:140-154). */
int cpsd_gram_tn(const cpsd_gram_tn_desc* descs, int nprob, int p_max, int q_max,
                 cudaStream_t stream);
/* same with fp64 accumulation; every record's `out` points to doubles (ldo in doubles).  Feeds
 * the variance thresholds and eigen-solvers of the PCA stages. */
int cpsd_gram_tn_f64(const cpsd_gram_tn_desc* descs, int nprob, int p_max, int q_max,
                     cudaStream_t stream);
/* fp64 accumulation, every problem's row segments dealt to nsplit CTAs per output tile (few
 * problems x many rows); the partial tiles are staged in part_ws
 * (cpsd_gram_tn_split_ws_elems doubles) and added in a fixed order (bit-reproducible). */
long long cpsd_gram_tn_split_ws_elems(int nprob, int p_max, int ldo_max, int nsplit);
int cpsd_gram_tn_f64_split(const cpsd_gram_tn_desc* descs, int nprob, int p_max, int q_max,
                           int ldo_max, int nsplit, double* part_ws, cudaStream_t stream);
/* (X - mu) W: PCA.transform, AlignCCA.transform (AlignCCA.py:93), MCCA transform_view
 * (AlignMCCA.py:110,125), JointPCA.transform (JointPCA.py:132,149); output rows land
 * directly in the pooled trials x (time*latent) matrix (cross_pt_decoders.py:260-270). */
int cpsd_proj_nn(const cpsd_proj_desc* descs, int nprob, int nseg_max, int seg_len, int q_max,
                 cudaStream_t stream);
/* the same projection for all folds of a batch on the tensor cores (tcgen05 kind::tf32,
 * 3xTF32, TMA): a persistent CTA keeps one 128-row tile of a patient in shared memory and
 * streams the loadings of every fold past it.  Needs the tf32 hi/lo split of every patient
 * (cpsd_split_tf32, once), tensor maps (cpsd_tmap_encode_f32: host-encoded, copied to the
 * device by the caller; X maps with box_rows = 128, L^T maps with box_rows = 32) and the
 * per-problem L^T hi/lo + mu L prepared by cpsd_proj_tc_prep.  Any channel count <= 256 (the
 * hi / lo arrays get a row stride padded to a multiple of 4 floats by cpsd_split_tf32_2d; TMA
 * zero-fills the box beyond the last channel; more than 128 channels run as two resident panels,
 * the second adding to the first's output) and any latent size <= 128 (chunks of 32 columns;
 * the L^T arrays hold ceil(Q / 32) blocks of 32 rows x ltc columns per problem, ltc = 128 or
 * 256 >= the widest patient). */
int cpsd_split_tf32(const float* src, float* hi, float* lo, long long n, cudaStream_t stream);
int cpsd_tmap_encode_f32(void* map_out_host, const float* base, long long rows, int cols,
                         long long ld, int box_rows);
int cpsd_split_tf32_2d(const float* src, long long lds, long long rows, int cols, float* hi,
                       float* lo, int ldd, cudaStream_t stream);
int cpsd_proj_tc_prep(const float* L, int ldl, long long strideL, const float* mu_base,
                      const int* slot, int ld_mu, const int* cdim, int Q, int ltc, float* LtHi,
                      float* LtLo, float* muL, int nprob, cudaStream_t stream);
int cpsd_proj_tc(const void* xmaps_dev, const void* ltmaps_dev, int P, int B, int T, int Q, int ltc,
                 const int* n_trials_host, const int* n_chan_host, int n_max, const int* dst_row,
                 const float* muL, float* Y, long long strideY, int num_sms, cudaStream_t stream);
/* the same for several REPLICAS of the patient set in one launch (independent jobs with the same
 * shapes whose folds share a batch, cv_align_decode_stream): xmaps / n_trials / n_chan hold
 * n_rep * P entries (replica-major); replica r is multiplied with the folds
 * [fold_beg_host[r], fold_beg_host[r] + fold_cnt_host[r]) of the batch */
int cpsd_proj_tc_rep(const void* xmaps_dev, const void* ltmaps_dev, int P, int n_rep, int B, int T,
                     int Q, int ltc, const int* n_trials_host, const int* n_chan_host,
                     const int* fold_beg_host, const int* fold_cnt_host, int n_max,
                     const int* dst_row, const float* muL, float* Y, long long strideY, int num_sms,
                     cudaStream_t stream);
/* A B^T over the long feature axis: Gram of the pooled matrix for the decoder-stage PCA
 * (decomposition/DimRedReshape.py:47-49 -> sklearn PCA full SVD). */
int cpsd_gram_nt(const cpsd_gram_nt_desc* descs, int nprob, int m_max, int n_max,
                 cudaStream_t stream);
/* same product on the tensor cores: TMA -> 128B-swizzled smem -> tcgen05.mma.kind::tf32 with
 * the accumulator in TMEM, 3xTF32 split (hi*hi + hi*lo + lo*hi), symmetric problems only.
 * descs_host is a HOST array (tensor maps are encoded on the host); split_ws holds the
 * hi/lo copies (2 * sum m*lda floats), map_ws / stage_host (pinned) hold the tensor maps,
 * both cpsd_gram_nt_tc_ws_bytes(nprob) bytes. */
int cpsd_gram_nt_tc(const cpsd_gram_nt_desc* descs_host, int nprob, int m_max, int n_max,
                    float* split_ws, long long split_ws_elems, void* map_ws, void* stage_host,
                    cudaStream_t stream);
/* Gram of the centred rows (row p of mu subtracted inside the hi/lo split) */
int cpsd_gram_nt_tc_centered(const cpsd_gram_nt_desc* descs_host, int nprob, int m_max, int n_max,
                             float* split_ws, long long split_ws_elems, void* map_ws,
                             void* stage_host, const float* mu, int ldmu, cudaStream_t stream);
/* profiling hook: CUDA event (cudaEvent_t) recorded between the hi/lo split and the MMA kernel of
 * the following cpsd_gram_nt_tc* calls; NULL switches it off */
int cpsd_gram_nt_tc_probe(void* event);
int cpsd_gram_nt_tc_ws_bytes(int nprob);

/* ---- small solvers ---------------------------------------------------------------
 * symmetric eigen-decomposition, n <= 128, one CTA per problem, shared-memory Jacobi
 * (replaces LAPACK syevd/gesdd behind PCA, n_components_var and the MCCA GEVP). */
int cpsd_eig_sym_small(const float* A, int lda, long long strideA, const int* n_dev, int n_fixed,
                       int nprob, float* evals, int ld_e, float* evecs, int ldv, long long strideV,
                       int max_sweeps, float tol, int* sweeps_out, cudaStream_t stream);
/* same solver iterating the matrix in fp64 (A holds doubles) with fp32 eigenvectors: rotation
 * angles stay accurate for near-degenerate (noise-level) eigenvalues */
int cpsd_eig_sym_small_f64(const double* A, int lda, long long strideA, const int* n_dev,
                           int n_fixed, int nprob, float* evals, int ld_e, float* evecs, int ldv,
                           long long strideV, int max_sweeps, float tol, int* sweeps_out,
                           cudaStream_t stream);
/* the same solver on a subset (sel) of a batch, results to slot out_idx[problem], eigenvector
 * accumulator started from V0[v0_idx[problem]] (the caller rotates A into that basis first):
 * the per-fold scatter matrices of a cross-validation (cross_pt_decoders.py:234-241,
 * AlignMCCA.py:140-174 refit them from scratch every fold) are nearly diagonal in the
 * eigenbasis of their all-trials version, which halves the number of sweeps */
int cpsd_eig_sym_small_f64_warm(const double* A, int lda, long long strideA, const int* n_dev,
                                int n_fixed, const int* sel, int nsel, const int* out_idx,
                                float* evals, int ld_e, float* evecs, int ldv, long long strideV,
                                const float* V0, int ldv0, long long strideV0, const int* v0_idx,
                                int max_sweeps, float tol, int* sweeps_out, cudaStream_t stream);
/* C[ci[p]] = alpha * op(A[ai[p]]) * B[bi[p]] + beta * D[di[p]] in fp64 (null list: p; D may be
 * NULL): the basis changes V0^T A V0 of the warm-started solves */
int cpsd_dgemm_batched(int transA, int m, int n, int k, double alpha, const double* A, int lda,
                       long long strideA, const int* a_idx, const double* B, int ldb,
                       long long strideB, const int* b_idx, double beta, const double* D, int ldd,
                       long long strideD, const int* d_idx, double* C, int ldc, long long strideC,
                       const int* c_idx, int nprob, cudaStream_t stream);
int cpsd_cast_f32_f64_idx(const float* src, long long strideS, const int* idx, double* dst,
                          long long strideD, long long elems, int nprob, cudaStream_t stream);
/* block two-sided Jacobi for n > 128 (n_pad multiple of 128): pooled-Gram PCA, large GEVPs.
 * Every round's 128x128 tile rotations are kept in a rotation log (Rlog,
 * cpsd_bj_rlog_elems() floats); cpsd_bj_eigvecs replays them on the identity columns of the
 * leading k eigenvalues only, instead of accumulating all n eigenvectors every round. */
int cpsd_bj_schedule(int n_pad, int* pairs_host);
long long cpsd_bj_rlog_elems(int n_pad, int nprob, int max_sweeps);
int cpsd_eig_sym_block(float* K, int ld, long long stride, int n_pad, const int* n_dev,
                       int n_fixed, int nprob, const int* pairs_dev, float* Rlog, float* fwork,
                       int* iwork, float* evals, int* perm, int ld_e, int max_sweeps, float tol,
                       float* RTbuf /* nprob*(n_pad/128)*128*128 floats: tcgen05 tile update;
                                       NULL: fp32 SIMT update */,
                       cudaStream_t stream);
int cpsd_bj_eigvecs(const float* Rlog, int n_pad, int nprob, const int* pairs_dev,
                    const int* iwork, const int* perm, int ld_perm, const int* k_dev, int k_fixed,
                    int k_launch, float* E, int lde, long long strideE, int max_sweeps,
                    cudaStream_t stream);
/* component counts: sklearn PCA float n_components (mode 0), AlignMCCA.n_components_var
 * (mode 1, AlignMCCA.py:174), NoCenterPCA (mode 2, NoCenterPCA.py:101-103), integer (mode 3) */
int cpsd_select_k(const float* evals, int ld_e, const int* n_dev, int n_fixed, float thr, int mode,
                  int kmin, int kmax, int* k_out, int k_stride, int nprob, cudaStream_t stream);
/* same selection from the leading n eigenvalues only, total variance (trace) given */
int cpsd_select_k_total(const float* evals, int ld_e, const int* n_dev, int n_fixed,
                        const float* total_dev, float thr, int mode, int kmin, int kmax, int* k_out,
                        int k_stride, int nprob, cudaStream_t stream);
/* Leading m (<= 128) eigen-pairs of symmetric PSD matrices by block subspace iteration
 * (Y = K Q, Cholesky QR) + Rayleigh-Ritz: the decoder-stage PCA with a float n_components
 * (decomposition/DimRedReshape.py:47-49 -> sklearn PCA) keeps only the components that explain
 * the requested variance, and the total variance is the trace.  K's padding (rows / columns
 * >= n) is zeroed, K is otherwise preserved.  ws: cpsd_eig_topk_ws_elems() floats; the
 * eigenvectors land at ws + cpsd_eig_topk_voff() as (n_pad x m) per problem, row stride m,
 * problem stride 2*n_pad*m.  resid[prob][j] = ||K v_j - theta_j v_j||; status bit 0 = the
 * block lost rank.  init = 0 continues from the Ritz vectors of the previous call.
 * f64_gram != 0 accumulates the Gram of every Cholesky-QR step in fp64: needed when the leading
 * spectrum spans more than ~3e3 (an fp32 Gram of K Q is then no longer positive definite and
 * status reports it); callers try 0 first. */
long long cpsd_eig_topk_ws_elems(int n_pad, int m, int nprob);
long long cpsd_eig_topk_voff(int n_pad, int m, int nprob);
int cpsd_eig_sym_topk(float* K, int ld, long long stride, int n_pad, const int* n_dev, int n_fixed,
                      int nprob, int m, int iters, int init, float* ws, float* evals, int ld_e,
                      float* total, float* resid, int* status, int eig_sweeps, float eig_tol,
                      int f64_gram, cudaStream_t stream);
/* same solver with Y = K Q on the tensor cores (tcgen05 kind::tf32 + TMA): single-pass TF32
 * for the first tf32_iters iterations of a fresh start (the iteration is self-correcting),
 * 3xTF32 afterwards.  tc_ws: cpsd_topk_tc_ws_elems() floats; map_dev: cpsd_topk_tc_map_bytes()
 * bytes (64-byte aligned) filled by cpsd_topk_tc_encode (host-encoded tensor maps, staged in
 * pinned stage_host) whenever K or tc_ws move.  Needs m = 128 and densely packed K. */
long long cpsd_topk_tc_ws_elems(int n_pad, int nprob);
int cpsd_topk_tc_map_bytes(int nprob);
int cpsd_topk_tc_encode(const float* K, int ld, long long stride, int n_pad, int nprob, float* tc_ws,
                        void* map_dev, void* stage_host, cudaStream_t stream);
int cpsd_topk_tc_split_k(const float* K, int ld, long long stride, int n_pad, int nprob,
                         float* tc_ws, cudaStream_t stream);
int cpsd_topk_tc_kq(const float* Q, long long strideQ, float* Y, long long strideY, int n_pad,
                    int nprob, int terms, float* tc_ws, const void* map_dev, cudaStream_t stream);
int cpsd_eig_sym_topk_tc(float* K, int ld, long long stride, int n_pad, const int* n_dev,
                         int n_fixed, int nprob, int m, int iters, int init, float* ws, float* evals,
                         int ld_e, float* total, float* resid, int* status, int eig_sweeps,
                         float eig_tol, float* tc_ws, const void* map_dev, int tf32_iters,
                         int f64_gram, cudaStream_t stream);
/* building blocks of the above, exported for the kernel-level parity tests:
 * C = alpha op(A) B (batched, element strides; trans_a: A stored K x M), and the inverse of
 * the upper Cholesky factor of an m x m (m <= 128) Gram (fp64 in shared memory). */
int cpsd_sgemm_batched(int trans_a, int M, int N, int K, float alpha, const float* A, int lda,
                       long long strideA, const float* B, int ldb, long long strideB, float* C,
                       int ldc, long long strideC, int nprob, cudaStream_t stream);
int cpsd_chol_inv(const float* S, int lds, long long strideS, int m, float* Rinv, int ldr,
                  long long strideR, int* status, int nprob, cudaStream_t stream);
/* SPD solve S W = B (fp64 Cholesky in shared memory, m <= 128; B is used as workspace): the
 * least-squares read-in matrices pinv(X_p) @ latent of get_joint_PCA_transforms
 * (alignment/JointPCA.py:203-206) from X_p^T X_p and X_p^T latent */
int cpsd_chol_solve_f64(const double* S, int lds, long long strideS, int m, double* B, int ldb,
                        long long strideB, int q, float* W, int ldw, long long strideW,
                        int* status, int nprob, cudaStream_t stream);
/* any m: the factor lives in ws (nprob * m * (m + 1) doubles) instead of shared memory */
int cpsd_chol_solve_f64_ws(const double* S, int lds, long long strideS, int m, double* B, int ldb,
                           long long strideB, int q, float* W, int ldw, long long strideW,
                           int* status, double* ws, int nprob, cudaStream_t stream);
/* PCA components with sklearn's svd_flip sign convention, zero padded to dmax columns */
int cpsd_pca_basis(const float* evecs, int ldv, long long strideV, const int* k_dev,
                   const int* cdim, int c_fixed, int dmax, float* W, int ldw, int Cmax, int nprob,
                   cudaStream_t stream);
/* CCA_align in Gram form: Cholesky whitening + one-sided Jacobi SVD (AlignCCA.py:235-285),
 * and the b->a map M_b pinv(M_a) of AlignCCA.transform (AlignCCA.py:93) */
int cpsd_cca_solve(const cpsd_cca_desc* descs, int nprob, int dmax, cudaStream_t stream);
/* the same in fp64 (the default of the engine and of AlignCCA): the records' Saa / Sbb / Sab
 * point to DOUBLES (row stride lds in doubles, e.g. from cpsd_gram_tn_f64); Cholesky factors,
 * whitened cross-scatter, its SVD and every back-substitution in fp64; dmax <= 256 (operands
 * in shared memory while they fit, else in the L2-resident workspace `ws` of
 * cpsd_cca_solve_f64_ws_elems doubles); outputs fp32 as above; info[1] = 1 flags a scatter
 * matrix whose smallest Cholesky pivot fell below rank_tol (rank-deficient latents: the
 * reference truncates to matrix_rank at AlignCCA.py:263-265, here the caller is told) */
long long cpsd_cca_solve_f64_ws_elems(int nprob, int dmax);
int cpsd_cca_solve_f64(const cpsd_cca_desc* descs, int nprob, int dmax, double* ws,
                       cudaStream_t stream);
/* symmetric eigen-decomposition in fp64 for 128 < n <= n_cap <= 256 (one-sided Jacobi on an
 * L2-resident workspace of cpsd_eig_sym_f64_ws_elems doubles): channel covariances / scatter
 * matrices of patients with more than 128 electrodes (sklearn PCA reached from
 * decoders/cross_pt_decoders.py:234-241; AlignMCCA.n_components_var AlignMCCA.py:156-174).
 * Positive semi-definite input; eigenvalues descending (fp32), eigenvectors as sorted columns
 * (fp32; may be NULL) */
long long cpsd_eig_sym_f64_ws_elems(int nprob, int n_cap);
int cpsd_eig_sym_f64(const double* A, int lda, long long strideA, const int* n_dev, int n_fixed,
                     int nprob, float* evals, int ld_e, float* evecs, int ldv, long long strideV,
                     int max_sweeps, double* ws, int n_cap, int* sweeps_out, cudaStream_t stream);

/* ---- MCCA assembly (mvlearn.embed.MCCA as called at AlignMCCA.py:152-153) ---------- */
int cpsd_mcca_mask(const float* evecs, int ldv, long long strideV, const float* evals, int ld_e,
                   const int* rank, const int* cdim, int R, int Cmax, float* Vr, float* d2,
                   int* r_eff, int nprob, cudaStream_t stream);
/* same with an indirection: problem p reads the eigen-pairs of slot src_idx[p] (fold-invariant
 * view statistics of the cross patients are solved once and shared by the folds) */
int cpsd_mcca_mask_idx(const float* evecs, int ldv, long long strideV, const float* evals, int ld_e,
                       const int* rank, const int* cdim, const int* src_idx, int R, int Cmax,
                       float* Vr, float* d2, int* r_eff, int nprob, cudaStream_t stream);
int cpsd_mcca_build(const float* G, int ldg, long long strideG, const int* r_eff, int P, int R,
                    float reg, float* M, int ldm, long long strideM, int* n_out, int* cidx,
                    float* dh, int n_comp, int* status, int nfold, cudaStream_t stream);
/* the cross-scatter in two parts: per-fold target rows [G_tt | G_tx] (R x P*R) and the cached
 * fold-invariant cross x cross block of slot xslot[f] */
int cpsd_mcca_build_split(const float* Gt, int ldg, long long strideG, const float* Gxx, int ldx,
                          long long strideX, const int* xslot, const int* r_eff, int P, int R,
                          float reg, float* M, int ldm, long long strideM, int* n_out, int* cidx,
                          float* dh, int n_comp, int* status, int nfold, cudaStream_t stream);
int cpsd_mcca_loadings(const float* Vr, const float* U, int ldu, long long strideU,
                       const int* perm, int ld_perm, const int* r_eff, const float* dh, int P,
                       int R, int Cmax, int n_comp, float* L, int ldl, int nfold,
                       cudaStream_t stream);

/* ---- JointPCA assembly (alignment/JointPCA.py:165-211), batched over folds ----------------
 * G: fp64 Gram of the channel-concatenated class averages (blocks u <= v), s: its column sums.
 * joint_cov: sklearn-PCA covariance of the concatenation (fp32, zero padded);  joint_rhs:
 * X_p^T (M - 1 mean^T) V_k per (fold, patient), the right-hand sides of the least-squares
 * read-in matrices pinv(X_p) @ latent (solved by cpsd_chol_solve_f64 on G's diagonal blocks). */
int cpsd_joint_cov(const double* G, const float* s, const int* nrows, int n, float* cov, int n_pad,
                   int nfold, cudaStream_t stream);
int cpsd_joint_rhs(const double* G, const float* s, const int* nrows, int n, const float* V, int ldv,
                   long long strideV, const int* coff, int P, int C_max, int k, double* rhs, int ldr,
                   long long strideR, int nfold, cudaStream_t stream);

/* ---- decoder stage -----------------------------------------------------------------
 * PCA scores of the pooled train / test trials from the Gram eigen-pairs */
int cpsd_scores_train(const float* V, int ldv, long long strideV, const float* evals,
                      const int* perm, int ld_e, const int* k_dev, const int* n_dev, int n_fixed,
                      int n_max, float* St, int lds, long long strideS, int kcap, int nfold,
                      cudaStream_t stream);
int cpsd_scores_test(const float* Kte, int ldk, long long strideK, const float* V, int ldv,
                     long long strideV, const float* evals, const int* perm, int ld_e,
                     const int* k_dev, const int* n_dev, int n_fixed, int n_te, float* Ste,
                     int ldt, long long strideT, int kcap, int nfold, cudaStream_t stream);
/* one-vs-rest linear SVM: dual coordinate descent (liblinear solve_l2r_l1l2_svc, L2 loss)
 * + Newton polish on the same objective; the decoder injected at
 * decoders/cross_pt_decoders.py:46-59 */
int cpsd_svm_fit_ovr(const cpsd_svm_desc* descs, int ntask, int k_max, int n_max,
                     cudaStream_t stream);
/* same; dcd_epochs_max = the largest dcd_epochs of any task (0: Newton only, smaller footprint) */
int cpsd_svm_fit_ovr_ex(const cpsd_svm_desc* descs, int ntask, int k_max, int n_max,
                        int dcd_epochs_max, cudaStream_t stream);
int cpsd_svm_predict_ovr(const float* Xt, int ldx, long long strideX, const double* W, int ldw,
                         long long strideW, const int* k_dev, int k_fixed, const int* n_te,
                         int n_te_max, const int* classes, int ncls, int* yhat, double* dec,
                         int nfold, cudaStream_t stream);

/* kernel C-SVC with one-vs-one voting (libsvm's algorithm): the reference scripts' literal
 * decoders SVC(kernel='rbf', class_weight='balanced') (scripts/aligned_decode_svm_ncv.py:313-317)
 * and SVC(kernel='linear') inside BaggingClassifier (scripts/aligned_decode_svm.py:262-263).
 * St: feature-major pool scores (k x n, leading dimension lds) per fold; kernel 0 = linear,
 * 1 = rbf; gamma <= 0 means 'scale' = 1 / (k * var(X)); the per-fold gamma is written to
 * gamma_out; y: pool labels (ldy per fold), classes: sorted label values; K: (n x n) float kernel
 * matrix per fold in CLASS-SORTED sample order (perm / cls_off / sqn: [fold][ldk] ints,
 * [fold][ncls+1] ints, [fold][ldk] doubles, written here), so that a class pair's block is two
 * contiguous row / column ranges. */
int cpsd_svc_kernel_matrix(const float* St, int lds, long long strideS, const int* k_dev,
                           int k_fixed, const int* n_dev, int n_fixed, int n_max, const int* y, int ldy,
                           const int* classes, int ncls, int kernel, double gamma, double* gamma_out,
                           int* perm, int* cls_off, double* sqn, float* K, int ldk, long long strideK,
                           int nfold, cudaStream_t stream);
/* SMO per (fold, class pair), one warp each, pairs in libsvm's order, on the class-sorted kernel
 * matrix cpsd_svc_kernel_matrix left in K (perm: [fold][ldk] sorted position -> pool sample,
 * cls_off: [fold][ncls+1]); balanced != 0: C_c = C * n / (n_classes * count_c) (sklearn
 * class_weight='balanced'); eps = libsvm's tol; coef: [fold][ncls-1][ldc] in libsvm's sv_coef
 * layout indexed by pool sample; rho: [fold][npair]; info: [fold][npair][2] = iterations,
 * status (0 converged, 1 iteration cap, 2 class absent, 3 pair larger than m_max). */
int cpsd_svc_fit_ovo(const float* K, int ldk, long long strideK, const int* perm, const int* cls_off,
                     const int* n_dev, int n_fixed, int ncls, double C, int balanced, double eps,
                     int max_iter, double* coef, int ldc, double* rho, int* info, int m_max, int nfold,
                     cudaStream_t stream);
/* votes of all pair decisions, first maximum wins (libsvm svm_predict_values); dec (optional):
 * [fold][n_te_max][npair] */
int cpsd_svc_predict_ovo(const float* St, int lds, long long strideS, const float* Ste, int ldt,
                         long long strideT, const int* k_dev, int k_fixed, const int* n_dev,
                         int n_fixed, int n_max, const int* n_te, int n_te_max, const int* y,
                         int ldy, const int* classes, int ncls, int kernel, const double* gamma,
                         const double* coef, int ldc, const double* rho, int* yhat, double* dec,
                         int k_max, int nfold, cudaStream_t stream);
/* Bagging around the C-SVC (BaggingClassifier(estimator=SVC(kernel='linear'), n_estimators=10),
 * scripts/aligned_decode_svm.py:262-265): cpsd_bag_gather makes n_est bootstrap problems per fold
 * (resampled training scores and labels from the index table idx[fold * n_est + e][t], the fold's
 * test scores and sizes copied) that run through cpsd_svc_kernel_matrix / fit / predict as
 * nfold * n_est folds; cpsd_bag_vote takes the majority vote of the estimators' labels (first
 * maximum in class order, numpy argmax over BaggingClassifier.predict_proba's vote counts) */
int cpsd_bag_gather(const float* St, int lds, long long strideS, const float* Ste, int ldt,
                    long long strideT, const int* y, int ldy, const int* idx, int ldi,
                    const int* k_dev, const int* n_dev, const int* nte_dev, int n_est, int kb,
                    float* St_b, int lds_b, float* Ste_b, int ldt_b, int* y_b, int* k_b, int* n_b,
                    int* nte_b, int nfold, cudaStream_t stream);
int cpsd_bag_vote(const int* yhat_b, int n_est, const int* classes, int ncls, const int* nte_dev,
                  int n_te_max, int* yhat, int nfold, cudaStream_t stream);

/* fused per-trial predict of a fitted cross-patient decoder (crossPtDecoder.predict,
 * decoders/cross_pt_decoders.py:70-71,444: aligner.transform(X, idx=0) -> DimRedReshape/PCA
 * transform -> one-vs-rest linear decision), one CTA per trial.  X: (n, T, C) fp64 device;
 * mu: (C) or NULL; A: (C x Q); pmean: (T*Q); P: (T*Q x k2) = components_^T; W: (ncls x (k2+1)),
 * intercept last (ncls = 1: binary, classes holds 2 labels); dec optional (n x ncls).  Each
 * trial is split into nsplit time slices (one CTA each; the last to finish reduces in slice
 * order and decides); ws_part: n * nsplit * roundup(k2, 32) doubles, ws_count: n ints, zero
 * before the first call.  scores_out (optional): PCA scores, feature-major (k2 x ld_scores) fp32;
 * with W = NULL the kernel stops there (input of cpsd_svc_predict_ovo for the C-SVC decoders). */
int cpsd_predict_fused(const double* X, int n, int T, int C, const float* mu, const float* A, int Q,
                       const float* pmean, const float* P, int k2, const double* W,
                       const int* classes, int ncls, int* yhat, double* dec, int nsplit,
                       double* ws_part, int* ws_count, float* scores_out, int ld_scores,
                       cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CPSD_B200_H */
