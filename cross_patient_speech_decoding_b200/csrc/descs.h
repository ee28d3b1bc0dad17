/* Descriptor records shared by host (ctypes / numpy structured arrays) and device.
 * Every batched kernel takes a device array of these; one record = one problem
 * (one (fold, patient) pair, one fold, ...).  All pointers are device pointers.
 * Rows of a logical operand are addressed as `nseg` segments of `seg_len` contiguous
 * rows each (a segment = one trial or one class average of T time bins), so trial /
 * class selections per CV fold never move data.  Keep in sync with _lib.py dtypes. */
#ifndef CPSD_DESCS_H
#define CPSD_DESCS_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  const float* A;      /* rows x p operand, row stride lda */
  const float* B;      /* rows x q operand, row stride ldb (may equal A) */
  const int* segA;     /* [nseg] first row of each segment in A */
  const int* segB;     /* [nseg] first row of each segment in B */
  const float* muA;    /* optional [p] vector subtracted from every A row */
  const float* muB;    /* optional [q] */
  float* out;          /* p x q, row stride ldo:  out = alpha * sum_r (a_r-muA)^T (b_r-muB) */
  int nseg, seg_len, p, q, lda, ldb, ldo, sym;
  float alpha;
  int pad_;
} cpsd_gram_tn_desc;

typedef struct {
  const float* A;      /* rows x p */
  const int* segA;
  float* out;          /* [p]  out = alpha * column sums */
  int nseg, seg_len, p, lda;
  float alpha;
  int pad_;
} cpsd_colsum_desc;

typedef struct {
  const float* X;       /* rows x C input, row stride ldx */
  const int* seg_src;   /* [nseg] first input row of each segment */
  const int* seg_dst;   /* [nseg] first output row of each segment */
  const float* mu;      /* optional [C] subtracted from every input row */
  const float* W;       /* C x q, row stride ldw */
  float* Y;             /* output rows x q, row stride ldy */
  int nseg, seg_len, C, q, ldx, ldw, ldy, pad_;
} cpsd_proj_desc;

typedef struct {
  const float* A;       /* m x k, row stride lda (k contiguous) */
  const float* B;       /* n x k, row stride ldb */
  float* out;           /* m x n, row stride ldo: out = alpha * A B^T */
  int m, n, k, lda, ldb, ldo, sym;
  float alpha;
} cpsd_gram_nt_desc;

typedef struct {
  const float* X;        /* (N, T, C) trials, contiguous */
  const int* member_ptr; /* [nslot+1] CSR offsets into members */
  const int* members;    /* trial ids */
  float* out;            /* (nslot, T, C) class means */
  int nslot, TC, pad0_, pad1_;
} cpsd_class_mean_desc;

typedef struct {
  const float* St;       /* k x n feature-major training scores, row stride lds */
  const int* y;          /* [n] integer labels */
  const int* k_dev;      /* optional device scalar with the feature count */
  double* w;             /* [kmax+1] output weights (bias last) */
  int* info;             /* [4] newton iterations, cg iterations, dcd epochs, status */
  int n, k, lds, cls;
  double C;
  double tol_dcd;        /* liblinear eps: stop DCD when PGmax - PGmin <= tol_dcd */
  double tol_newton;     /* stop Newton when |grad|_inf <= tol_newton * max(1, |grad(0)|_inf) */
  int max_newton, dcd_epochs;
} cpsd_svm_desc;

typedef struct {
  const float* Saa;      /* da x da scatter of the centred target latents, row stride lds */
  const float* Sbb;      /* db x db */
  const float* Sab;      /* da x db */
  const int* da_dev;     /* optional device scalars overriding da / db */
  const int* db_dev;
  float* Ma;             /* dmax x dmax (row stride ldm): M_a in the leading da x d block */
  float* Mb;             /* dmax x dmax: M_b in the leading db x d block */
  float* G;              /* dmax x dmax (row stride ldg): b->a map M_b pinv(M_a), db x da */
  float* rho;            /* [dmax] canonical correlations, descending, clamped to [0,1] */
  int* info;             /* [4] d, rank warning, svd sweeps, reserved */
  int da, db, lds, ldm, ldg, mode;
  float rank_tol;
  int pad_;
} cpsd_cca_desc;

#ifdef __cplusplus
}
#endif
#endif
