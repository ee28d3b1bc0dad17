// fp64 dense solvers for the alignment stage, one CTA per problem:
//
//  k_cca_solve_f64 : pairwise CCA in Gram form (reference: alignment/AlignCCA.py:235-285
//                    CCA_align, and the b->a map M_b pinv(M_a) of AlignCCA.py:93).  The
//                    Gram form squares the condition number of the latents, so the Cholesky
//                    factors, the whitened cross-scatter K = Ra^-T Sab Rb^-1 = Qa^T Qb, its
//                    SVD and every back-substitution run in fp64 (the fp32 solver moved the
//                    b->a map by up to 5e-3 on noisy latents with ~100 dimensions).
//  k_eig_sym_f64   : symmetric eigen-decomposition for 128 < n <= 256 in fp64 (channel
//                    covariances / scatters of patients with more than 128 electrodes:
//                    sklearn PCA via decoders/cross_pt_decoders.py:234-241, the spectrum of
//                    AlignMCCA.n_components_var AlignMCCA.py:156-174, the per-view bases of
//                    mvlearn's MCCA).  n <= 128 stays with the shared-memory tile solver.
//
// Both use the same one-sided (Hestenes) Jacobi iteration on ROWS of a matrix (a "column" of
// the mathematical operand is stored as a contiguous row, so one warp owns a row pair, reads it
// coalesced / conflict-free, keeps it in registers and writes it back once).  Matrices live in
// shared memory when the pair (W, V) fits, else in a global (L2-resident) workspace: the code is
// written against generic pointers.
#include "common.cuh"
#include "descs.h"

namespace {

#define S64_NT 512
#define S64_MAXE 8      // a row of <= 256 doubles = 8 per lane
#define S64_NMAX 256

// pairs (i, j), i + j = s (mod m), m even: step s has m/2 (s odd) or m/2 - 1 (s even) pairs
__device__ __forceinline__ void s64_pair(int s, int k, int m, int& i, int& j) {
  const int h = (s + 1) >> 1;
  const int e = (s & 1) ? 0 : 1;
  i = (h + e + k) % m;
  j = (h - 1 - k) % m;
  if (j < 0) j += m;
}

// Rotates row pairs of Wt (n rows of length m) until all rows are mutually orthogonal; the
// same rotations go to the rows of Vt (n rows of length nv; may be null).  Returns the number
// of sweeps.  All threads of the CTA must call it.
__device__ int s64_jacobi_rows(double* Wt, double* Vt, int n, int m, int nv, int ldw, int ldv,
                               int* flag, int max_sweeps) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int me = (n + 1) & ~1;
  int sw = 0;
  if (n < 2) return 0;
  for (; sw < max_sweeps; ++sw) {
    if (threadIdx.x == 0) *flag = 0;
    __syncthreads();
    for (int s = 0; s < me; ++s) {
      const int npairs = (s & 1) ? (me >> 1) : (me >> 1) - 1;
      for (int k = wid; k < npairs; k += nw) {
        int p, q;
        s64_pair(s, k, me, p, q);
        if (p >= n || q >= n) continue;
        double* xp = Wt + (long long)p * ldw;
        double* yq = Wt + (long long)q * ldw;
        double x[S64_MAXE], y[S64_MAXE];
        double al = 0.0, be = 0.0, ga = 0.0;
#pragma unroll
        for (int u = 0; u < S64_MAXE; ++u) {
          const int r = lane + 32 * u;
          x[u] = (r < m) ? xp[r] : 0.0;
          y[u] = (r < m) ? yq[r] : 0.0;
          al = fma(x[u], x[u], al);
          be = fma(y[u], y[u], be);
          ga = fma(x[u], y[u], ga);
        }
        al = warp_sum(al);
        be = warp_sum(be);
        ga = warp_sum(ga);
        const double lim = sqrt(al) * sqrt(be);
        if (fabs(ga) > 4e-16 * lim && fabs(ga) > 1e-290) {
          if (fabs(ga) > 1e-13 * lim && lane == 0) *flag = 1;
          const double zeta = (be - al) / (2.0 * ga);
          const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          const double c = 1.0 / sqrt(1.0 + t * t);
          const double sn = t * c;
#pragma unroll
          for (int u = 0; u < S64_MAXE; ++u) {
            const int r = lane + 32 * u;
            if (r < m) {
              xp[r] = c * x[u] - sn * y[u];
              yq[r] = sn * x[u] + c * y[u];
            }
          }
          if (Vt) {
            double* vp = Vt + (long long)p * ldv;
            double* vq = Vt + (long long)q * ldv;
#pragma unroll
            for (int u = 0; u < S64_MAXE; ++u) {
              const int r = lane + 32 * u;
              if (r < nv) {
                const double a = vp[r], b = vq[r];
                vp[r] = c * a - sn * b;
                vq[r] = sn * a + c * b;
              }
            }
          }
        }
      }
      __syncthreads();
    }
    const int f = *flag;
    __syncthreads();
    if (!f) { ++sw; break; }
  }
  return sw;
}

// In-place upper Cholesky A = R^T R of the leading n x n block (row stride ld); the strict
// lower part is zeroed.  *minpiv_out (thread 0) = smallest pivot / largest diagonal entry.
__device__ void s64_chol_upper(double* A, int n, int ld, double* minpiv_out) {
  double mp = 1e300, dmax = 0.0;
  for (int i = 0; i < n; ++i) dmax = fmax(dmax, A[i * ld + i]);
  const double floor_ = fmax(dmax, 1e-280) * 1e-24;
  for (int k = 0; k < n; ++k) {
    __syncthreads();
    const double akk = A[k * ld + k];
    const double piv = sqrt(fmax(akk, floor_));
    mp = fmin(mp, akk / fmax(dmax, 1e-280));
    __syncthreads();
    for (int j = k + threadIdx.x; j < n; j += blockDim.x)
      A[k * ld + j] = (j == k) ? piv : A[k * ld + j] / piv;
    __syncthreads();
    const int m = n - k - 1;
    for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
      const int i = k + 1 + e / m, j = k + 1 + e % m;
      if (j >= i) A[i * ld + j] -= A[k * ld + i] * A[k * ld + j];
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    const int i = e / n, j = e % n;
    if (j < i) A[i * ld + j] = 0.0;
  }
  if (threadIdx.x == 0 && minpiv_out) *minpiv_out = mp;
  __syncthreads();
}

// R^T Y = B in place (R upper n x n, B n x m), one thread per column of B
__device__ void s64_solve_rt_cols(const double* R, int ldr, double* B, int ldb, int n, int m) {
  for (int c = threadIdx.x; c < m; c += blockDim.x) {
    for (int i = 0; i < n; ++i) {
      double v = B[i * ldb + c];
      for (int k = 0; k < i; ++k) v = fma(-R[k * ldr + i], B[k * ldb + c], v);
      B[i * ldb + c] = v / R[i * ldr + i];
    }
  }
  __syncthreads();
}

// R X = B in place (R upper n x n, B n x m), one thread per column of B
__device__ void s64_solve_r_cols(const double* R, int ldr, double* B, int ldb, int n, int m) {
  for (int c = threadIdx.x; c < m; c += blockDim.x) {
    for (int i = n - 1; i >= 0; --i) {
      double v = B[i * ldb + c];
      for (int k = i + 1; k < n; ++k) v = fma(-R[i * ldr + k], B[k * ldb + c], v);
      B[i * ldb + c] = v / R[i * ldr + i];
    }
  }
  __syncthreads();
}

// B <- B R^-1 (R upper n x n, B m x n), one thread per row of B
__device__ void s64_solve_right_r(const double* R, int ldr, double* B, int ldb, int m, int n) {
  for (int r = threadIdx.x; r < m; r += blockDim.x) {
    for (int j = 0; j < n; ++j) {
      double v = B[r * ldb + j];
      for (int k = 0; k < j; ++k) v = fma(-B[r * ldb + k], R[k * ldr + j], v);
      B[r * ldb + j] = v / R[j * ldr + j];
    }
  }
  __syncthreads();
}

// every row x of Xt (nrows rows of length n) <- R^-1 x  (R upper n x n), one thread per row
__device__ void s64_solve_r_rows(const double* R, int ldr, double* Xt, int ldx, int n, int nrows) {
  for (int j = threadIdx.x; j < nrows; j += blockDim.x) {
    double* x = Xt + (long long)j * ldx;
    for (int i = n - 1; i >= 0; --i) {
      double v = x[i];
      for (int k = i + 1; k < n; ++k) v = fma(-R[i * ldr + k], x[k], v);
      x[i] = v / R[i * ldr + i];
    }
  }
  __syncthreads();
}

// squared norms of the first n rows of Wt (length m), one warp per row
__device__ void s64_row_norms(const double* Wt, int ldw, int n, int m, double* out) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int j = wid; j < n; j += nw) {
    double a = 0.0;
    for (int r = lane; r < m; r += 32) a = fma(Wt[(long long)j * ldw + r], Wt[(long long)j * ldw + r], a);
    a = warp_sum(a);
    if (lane == 0) out[j] = sqrt(a);
  }
  __syncthreads();
}

// descending rank of vals[0..n) (ties by index)
__device__ void s64_rank_desc(const double* vals, int n, int* rank) {
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    int r = 0;
    const double vj = vals[j];
    for (int u = 0; u < n; ++u) r += (vals[u] > vj) || (vals[u] == vj && u < j);
    rank[j] = r;
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------
// Pairwise CCA.  descs: cpsd_cca_desc whose Saa / Sbb / Sab point to DOUBLES (row stride
// lds in doubles); outputs fp32 as in the fp32 entry.  gws: per-CTA global workspace of
// 4 * dmax * (dmax + 1) doubles; with use_smem the two SVD operands sit in shared memory.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(S64_NT)
k_cca_solve_f64(const cpsd_cca_desc* __restrict__ descs, int nprob, int dmax, double* __restrict__ gws,
                int use_smem) {
  extern __shared__ __align__(16) unsigned char smraw[];
  __shared__ double sig[S64_NMAX];
  __shared__ int rank[S64_NMAX];
  __shared__ double minpiv[2];
  __shared__ int flag;
  const int ld = dmax + 1;
  const long long msz = (long long)dmax * ld;
  double* B0 = gws + (long long)blockIdx.x * 4 * msz;
  double* B1 = B0 + msz;
  double* B2 = use_smem ? reinterpret_cast<double*>(smraw) : B1 + msz;
  double* B3 = B2 + msz;

  for (int prob = blockIdx.x; prob < nprob; prob += gridDim.x) {
    const cpsd_cca_desc t = descs[prob];
    const double* Saa = reinterpret_cast<const double*>(t.Saa);
    const double* Sbb = reinterpret_cast<const double*>(t.Sbb);
    const double* Sab = reinterpret_cast<const double*>(t.Sab);
    int da = t.da_dev ? t.da_dev[0] : t.da;
    int db = t.db_dev ? t.db_dev[0] : t.db;
    da = max(0, min(da, dmax));
    db = max(0, min(db, dmax));
    const int d = min(da, db);
    __syncthreads();     // previous problem's readers of the workspace are done
    if (d == 0) {
      if (threadIdx.x == 0 && t.info) { t.info[0] = 0; t.info[1] = 2; t.info[2] = 0; t.info[3] = 0; }
      for (int e = threadIdx.x; e < dmax * dmax; e += blockDim.x) {
        const int i = e / dmax, j = e % dmax;
        t.Ma[(long long)i * t.ldm + j] = 0.f;
        t.Mb[(long long)i * t.ldm + j] = 0.f;
        t.G[(long long)i * t.ldg + j] = 0.f;
      }
      for (int j = threadIdx.x; j < dmax; j += blockDim.x) t.rho[j] = 0.f;
      continue;
    }
    for (int e = threadIdx.x; e < da * da; e += blockDim.x) {
      const int i = e / da, j = e % da;
      B0[i * ld + j] = 0.5 * (Saa[(long long)i * t.lds + j] + Saa[(long long)j * t.lds + i]);
    }
    for (int e = threadIdx.x; e < db * db; e += blockDim.x) {
      const int i = e / db, j = e % db;
      B1[i * ld + j] = 0.5 * (Sbb[(long long)i * t.lds + j] + Sbb[(long long)j * t.lds + i]);
    }
    for (int e = threadIdx.x; e < da * db; e += blockDim.x) {
      const int i = e / db, j = e % db;
      B2[i * ld + j] = Sab[(long long)i * t.lds + j];
    }
    __syncthreads();
    s64_chol_upper(B0, da, ld, &minpiv[0]);
    s64_chol_upper(B1, db, ld, &minpiv[1]);
    // K = Ra^-T Sab Rb^-1 (da x db) = Qa^T Qb of the reference's thin QR factors
    s64_solve_rt_cols(B0, ld, B2, ld, da, db);
    s64_solve_right_r(B1, ld, B2, ld, da, db);

    // SVD by row rotations.  da >= db: rows of Wt = columns of K (-> B3 = K^T), Vt = B2;
    // da < db: rows of Wt = columns of K^T = rows of K (B2 as it is), Vt = B3.
    const bool tall = da >= db;
    const int m = tall ? da : db, n = d;
    double* Wt = tall ? B3 : B2;
    double* Vt = tall ? B2 : B3;
    if (tall) {
      for (int e = threadIdx.x; e < da * db; e += blockDim.x) {
        const int i = e / db, j = e % db;
        B3[j * ld + i] = B2[i * ld + j];
      }
      __syncthreads();
    }
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
      const int i = e / n, j = e % n;
      Vt[i * ld + j] = (i == j) ? 1.0 : 0.0;
    }
    __syncthreads();
    const int sweeps = s64_jacobi_rows(Wt, Vt, n, m, n, ld, ld, &flag, 40);
    s64_row_norms(Wt, ld, n, m, sig);
    s64_rank_desc(sig, n, rank);
    for (int e = threadIdx.x; e < n * m; e += blockDim.x) {
      const int j = e / m, r = e % m;
      Wt[j * ld + r] = (sig[j] > 0.0) ? Wt[j * ld + r] / sig[j] : 0.0;
    }
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
      double s = sig[j];
      s = (s < 0.0) ? 0.0 : s;        // clamp as AlignCCA.py:282-283
      s = (s >= 1.0) ? 1.0 : s;
      t.rho[rank[j]] = (float)s;
    }
    for (int j = n + threadIdx.x; j < dmax; j += blockDim.x) t.rho[j] = 0.f;
    __syncthreads();
    // a-side vectors (d rows of length da) are in B3, b-side (d rows of length db) in B2
    double* UaT = B3;
    double* VbT = B2;
    // M_b = Rb^-1 V_d
    s64_solve_r_rows(B1, ld, VbT, ld, db, d);
    for (int e = threadIdx.x; e < db * d; e += blockDim.x) {
      const int i = e / d, j = e % d;
      t.Mb[(long long)i * t.ldm + rank[j]] = (float)VbT[j * ld + i];
    }
    __syncthreads();
    double* Pm;     // pinv(M_a), d x da, rows in Jacobi order
    if (d == da) {
      // square M_a: pinv(M_a) = U^T Ra
      Pm = B1;
      for (int e = threadIdx.x; e < d * da; e += blockDim.x) {
        const int j = e / da, c = e % da;
        double a = 0.0;
        for (int k = 0; k <= c; ++k) a = fma(UaT[j * ld + k], B0[k * ld + c], a);
        Pm[j * ld + c] = a;
      }
      __syncthreads();
      s64_solve_r_rows(B0, ld, UaT, ld, da, d);       // rows of UaT <- M_a^T
    } else {
      // thin M_a (da x d, d < da): pinv = (M_a^T M_a)^-1 M_a^T
      s64_solve_r_rows(B0, ld, UaT, ld, da, d);
      for (int e = threadIdx.x; e < d * d; e += blockDim.x) {
        const int i = e / d, j = e % d;
        double a = 0.0;
        for (int k = 0; k < da; ++k) a = fma(UaT[i * ld + k], UaT[j * ld + k], a);
        B1[i * ld + j] = a;
      }
      __syncthreads();
      s64_chol_upper(B1, d, ld, nullptr);
      for (int e = threadIdx.x; e < d * da; e += blockDim.x) {
        const int i = e / da, c = e % da;
        B0[i * ld + c] = UaT[i * ld + c];
      }
      __syncthreads();
      s64_solve_rt_cols(B1, ld, B0, ld, d, da);
      s64_solve_r_cols(B1, ld, B0, ld, d, da);
      Pm = B0;
    }
    for (int e = threadIdx.x; e < da * d; e += blockDim.x) {
      const int i = e / d, j = e % d;
      t.Ma[(long long)i * t.ldm + rank[j]] = (float)UaT[j * ld + i];
    }
    // G (db x da) = M_b pinv(M_a); zero-filled padding so projections can use dmax columns
    for (int e = threadIdx.x; e < dmax * dmax; e += blockDim.x) {
      const int i = e / dmax, c = e % dmax;
      double a = 0.0;
      if (i < db && c < da)
        for (int k = 0; k < d; ++k) a = fma(VbT[k * ld + i], Pm[k * ld + c], a);
      t.G[(long long)i * t.ldg + c] = (float)a;
      if (c >= d || i >= da) t.Ma[(long long)i * t.ldm + c] = 0.f;
      if (c >= d || i >= db) t.Mb[(long long)i * t.ldm + c] = 0.f;
    }
    if (threadIdx.x == 0 && t.info) {
      t.info[0] = d;
      t.info[1] = (fmin(minpiv[0], minpiv[1]) < (double)t.rank_tol) ? 1 : 0;   // rank deficiency
      t.info[2] = sweeps;
      t.info[3] = 0;
    }
  }
}

// ---------------------------------------------------------------------------------------
// Symmetric eigen-decomposition in fp64, n <= 256, by row rotations of W = A (rows of the
// symmetric A are its columns): on exit row j of W is lambda_j v_j and row j of V is v_j.
// Eigenvalues descending (fp32 out), eigenvectors as sorted columns (fp32 out, optional).
// ws: per CTA 2 * n_cap * n_cap doubles.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(S64_NT)
k_eig_sym_f64(const double* __restrict__ A, int lda, long long strideA, const int* __restrict__ n_dev,
              int n_fixed, int nprob, float* __restrict__ evals, int ld_e, float* __restrict__ evecs,
              int ldv, long long strideV, int max_sweeps, double* __restrict__ ws, int n_cap,
              int* __restrict__ sweeps_out) {
  __shared__ double lam[S64_NMAX];
  __shared__ int rank[S64_NMAX];
  __shared__ int flag;
  const int ld = n_cap;
  double* Wt = ws + (long long)blockIdx.x * 2 * n_cap * n_cap;
  double* Vt = Wt + (long long)n_cap * n_cap;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int prob = blockIdx.x; prob < nprob; prob += gridDim.x) {
    int n = n_dev ? n_dev[prob] : n_fixed;
    n = max(0, min(n, n_cap));
    const double* Ag = A + (long long)prob * strideA;
    __syncthreads();
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
      const int r = e / n, c = e % n;
      Wt[r * ld + c] = 0.5 * (Ag[(long long)r * lda + c] + Ag[(long long)c * lda + r]);
      Vt[r * ld + c] = (r == c) ? 1.0 : 0.0;
    }
    __syncthreads();
    const int sw = s64_jacobi_rows(Wt, Vt, n, n, n, ld, ld, &flag, max_sweeps);
    if (sweeps_out && threadIdx.x == 0) sweeps_out[prob] = sw;
    // lambda_j = <w_j, v_j> (keeps the sign of slightly negative eigenvalues)
    for (int j = wid; j < n; j += nw) {
      double a = 0.0;
      for (int r = lane; r < n; r += 32) a = fma(Wt[j * ld + r], Vt[j * ld + r], a);
      a = warp_sum(a);
      if (lane == 0) lam[j] = a;
    }
    __syncthreads();
    s64_rank_desc(lam, n, rank);
    float* ev = evals + (long long)prob * ld_e;
    for (int i = threadIdx.x; i < ld_e; i += blockDim.x)
      if (i >= n) ev[i] = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) ev[rank[i]] = (float)lam[i];
    if (evecs) {
      float* Vg = evecs + (long long)prob * strideV;
      for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
        const int c = e / n, r = e % n;           // row c of Vt = eigenvector c, element r
        Vg[(long long)r * ldv + rank[c]] = (float)Vt[c * ld + r];
      }
    }
  }
}

int s64_grid(int nprob) { return nprob < 296 ? nprob : 296; }

// ---------------------------------------------------------------------------------------
// Batched fp64 GEMM on index lists: C[ci[p]] = alpha * op(A[ai[p]]) * B[bi[p]] + beta * D[di[p]]
// (null list: p).  64 x 64 output tile per CTA, 4 x 4 per thread, k in slabs of 16.  Used for the
// basis changes of the warm-started eigen-solves (V0^T A V0, fold-invariant bases, a few hundred
// 128^3 products per batch) and the fp64 re-orthonormalisation of those bases.
// ---------------------------------------------------------------------------------------
#define DG_T 64
#define DG_K 16
struct DgemmArgs {
  const double* A; int lda; long long sA; const int* ai;
  const double* B; int ldb; long long sB; const int* bi;
  const double* D; int ldd; long long sD; const int* di;
  double* C; int ldc; long long sC; const int* ci;
  int m, n, k, transA;
  double alpha, beta;
};

__global__ void __launch_bounds__(256)
k_dgemm_batched(DgemmArgs a) {
  __shared__ double As[DG_K][DG_T + 1];
  __shared__ double Bs[DG_K][DG_T];
  const int p = blockIdx.z;
  const double* A = a.A + (long long)(a.ai ? a.ai[p] : p) * a.sA;
  const double* B = a.B + (long long)(a.bi ? a.bi[p] : p) * a.sB;
  double* C = a.C + (long long)(a.ci ? a.ci[p] : p) * a.sC;
  const int r0 = blockIdx.y * DG_T, c0 = blockIdx.x * DG_T;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double acc[4][4] = {};
  for (int k0 = 0; k0 < a.k; k0 += DG_K) {
    for (int e = threadIdx.x; e < DG_K * DG_T; e += 256) {
      int kk, r;
      if (a.transA) { kk = e / DG_T; r = e - kk * DG_T; }     // A is k x m: rows contiguous in r
      else { r = e / DG_K; kk = e - r * DG_K; }               // A is m x k: rows contiguous in kk
      const int gr = r0 + r, gk = k0 + kk;
      double v = 0.0;
      if (gr < a.m && gk < a.k) v = a.transA ? A[(long long)gk * a.lda + gr] : A[(long long)gr * a.lda + gk];
      As[kk][r] = v;
    }
    for (int e = threadIdx.x; e < DG_K * DG_T; e += 256) {
      const int kk = e / DG_T, c = e - kk * DG_T;
      const int gk = k0 + kk, gc = c0 + c;
      Bs[kk][c] = (gk < a.k && gc < a.n) ? B[(long long)gk * a.ldb + gc] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < DG_K; ++kk) {
      double av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[kk][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const double* D = a.D ? a.D + (long long)(a.di ? a.di[p] : p) * a.sD : nullptr;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gr = r0 + ty + 16 * i;
    if (gr >= a.m) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gc = c0 + tx + 16 * j;
      if (gc >= a.n) continue;
      double v = a.alpha * acc[i][j];
      if (D) v += a.beta * D[(long long)gr * a.ldd + gc];
      C[(long long)gr * a.ldc + gc] = v;
    }
  }
}

__global__ void __launch_bounds__(256)
k_cast_f32_f64_idx(const float* __restrict__ src, long long strideS, const int* __restrict__ idx,
                   double* __restrict__ dst, long long strideD, long long elems) {
  const int p = blockIdx.y;
  const float* s = src + (long long)(idx ? idx[p] : p) * strideS;
  double* d = dst + (long long)p * strideD;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < elems;
       e += (long long)gridDim.x * blockDim.x)
    d[e] = (double)s[e];
}

}  // namespace

extern "C" int cpsd_dgemm_batched(int transA, int m, int n, int k, double alpha, const double* A,
                                  int lda, long long strideA, const int* a_idx, const double* B,
                                  int ldb, long long strideB, const int* b_idx, double beta,
                                  const double* D, int ldd, long long strideD, const int* d_idx,
                                  double* C, int ldc, long long strideC, const int* c_idx, int nprob,
                                  cudaStream_t stream) {
  CPSD_CHECK_ARG(m >= 0 && n >= 0 && k >= 0 && nprob >= 0, "dgemm_batched: bad dims");
  CPSD_CHECK_ARG(lda > 0 && ldb >= n && ldc >= n, "dgemm_batched: bad leading dimension");
  CPSD_CHECK_ARG(nprob <= 65535, "dgemm_batched: more than 65535 problems");
  if (m == 0 || n == 0 || nprob == 0) return CPSD_OK;
  DgemmArgs a{A, lda, strideA, a_idx, B, ldb, strideB, b_idx, D, ldd, strideD, d_idx,
              C, ldc, strideC, c_idx, m, n, k, transA, alpha, beta};
  k_dgemm_batched<<<dim3((n + DG_T - 1) / DG_T, (m + DG_T - 1) / DG_T, nprob), 256, 0, stream>>>(a);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

extern "C" int cpsd_cast_f32_f64_idx(const float* src, long long strideS, const int* idx, double* dst,
                                     long long strideD, long long elems, int nprob,
                                     cudaStream_t stream) {
  CPSD_CHECK_ARG(elems >= 0 && nprob >= 0 && nprob <= 65535, "cast_f32_f64_idx: bad dims");
  if (elems == 0 || nprob == 0) return CPSD_OK;
  long long bx = (elems + 255) / 256;
  if (bx > 64) bx = 64;
  k_cast_f32_f64_idx<<<dim3((int)bx, nprob), 256, 0, stream>>>(src, strideS, idx, dst, strideD, elems);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

extern "C" long long cpsd_cca_solve_f64_ws_elems(int nprob, int dmax) {
  return (long long)s64_grid(nprob > 0 ? nprob : 1) * 4 * dmax * (dmax + 1);
}

extern "C" int cpsd_cca_solve_f64(const cpsd_cca_desc* descs_dev, int nprob, int dmax, double* ws,
                                  cudaStream_t stream) {
  CPSD_CHECK_ARG(nprob >= 0 && dmax > 0, "cca_solve_f64: bad dims");
  if (dmax > S64_NMAX) {
    cpsd_set_error("cca_solve_f64: latent dimension above 256");
    return CPSD_ERR_UNSUPPORTED;
  }
  CPSD_CHECK_ARG(ws != nullptr, "cca_solve_f64: workspace is NULL");
  if (nprob == 0) return CPSD_OK;
  const size_t need = (size_t)2 * dmax * (dmax + 1) * sizeof(double);
  const int use_smem = need <= 216 * 1024;
  const size_t smem = use_smem ? need : 0;
  CPSD_CUDA(cudaFuncSetAttribute(k_cca_solve_f64, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(216 * 1024)));
  k_cca_solve_f64<<<s64_grid(nprob), S64_NT, smem, stream>>>(descs_dev, nprob, dmax, ws, use_smem);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

extern "C" long long cpsd_eig_sym_f64_ws_elems(int nprob, int n_cap) {
  return (long long)s64_grid(nprob > 0 ? nprob : 1) * 2 * n_cap * n_cap;
}

extern "C" int cpsd_eig_sym_f64(const double* A, int lda, long long strideA, const int* n_dev,
                                int n_fixed, int nprob, float* evals, int ld_e, float* evecs, int ldv,
                                long long strideV, int max_sweeps, double* ws, int n_cap,
                                int* sweeps_out, cudaStream_t stream) {
  CPSD_CHECK_ARG(nprob >= 0 && lda >= 0 && ld_e >= 0, "eig_sym_f64: bad dims");
  CPSD_CHECK_ARG(n_cap > 0 && n_cap <= S64_NMAX && n_fixed <= n_cap, "eig_sym_f64: n must be in 1..256");
  CPSD_CHECK_ARG(ws != nullptr, "eig_sym_f64: workspace is NULL");
  if (nprob == 0) return CPSD_OK;
  k_eig_sym_f64<<<s64_grid(nprob), S64_NT, 0, stream>>>(A, lda, strideA, n_dev, n_fixed, nprob, evals,
                                                        ld_e, evecs, ldv, strideV, max_sweeps, ws,
                                                        n_cap, sweeps_out);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}
