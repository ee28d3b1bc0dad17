// Pairwise CCA alignment in Gram form, one CTA per (fold, cross-patient) problem, all
// factors resident in shared memory.
//
// Reference: alignment/AlignCCA.py:235-285 (CCA_align): centre, thin QR of both latent
// matrices, SVD of Qa^T Qb, M = pinv(R) {U,V}[:, :d], clamp; AlignCCA.py:93 uses
// G = M_b pinv(M_a) to map patient B into patient A's latent space.
// Here: with the scatter matrices Saa = La^T La, Sbb, Sab of the
// centred latents, Cholesky Saa = Ra^T Ra (Ra equals the QR factor up to row signs),
// K = Ra^-T Sab Rb^-1 = Qa^T Qb, one-sided (Hestenes) Jacobi SVD of K, back-substitution.
#include "common.cuh"
#include "descs.h"

namespace {

#define CC_NT 256

// In-place upper Cholesky A = R^T R of the leading n x n block (row stride ld).
// Returns the smallest pivot ratio seen through *minpiv (thread 0 writes).
__device__ void chol_upper(float* A, int n, int ld, float* minpiv_out) {
  float mp = 1e30f;
  float dmax = 0.f;
  for (int i = 0; i < n; ++i) dmax = fmaxf(dmax, A[i * ld + i]);
  for (int k = 0; k < n; ++k) {
    __syncthreads();
    const float akk = A[k * ld + k];
    const float piv = sqrtf(fmaxf(akk, 1e-30f));
    mp = fminf(mp, akk / fmaxf(dmax, 1e-30f));
    __syncthreads();
    // scale row k
    for (int j = k + threadIdx.x; j < n; j += CC_NT) A[k * ld + j] = (j == k) ? piv : A[k * ld + j] / piv;
    __syncthreads();
    // trailing update A[i][j] -= R[k][i] R[k][j], i,j > k (upper part only)
    const int m = n - k - 1;
    for (int e = threadIdx.x; e < m * m; e += CC_NT) {
      const int i = k + 1 + e / m, j = k + 1 + e % m;
      if (j >= i) A[i * ld + j] -= A[k * ld + i] * A[k * ld + j];
    }
  }
  __syncthreads();
  // zero the strict lower part
  for (int e = threadIdx.x; e < n * n; e += CC_NT) {
    const int i = e / n, j = e % n;
    if (j < i) A[i * ld + j] = 0.f;
  }
  if (threadIdx.x == 0 && minpiv_out) *minpiv_out = mp;
  __syncthreads();
}

// Solve R^T Y = B in place (R upper n x n, B n x m): forward substitution, one thread per column.
__device__ void solve_rt(const float* R, int ldr, float* B, int ldb, int n, int m) {
  for (int c = threadIdx.x; c < m; c += CC_NT) {
    for (int i = 0; i < n; ++i) {
      float v = B[i * ldb + c];
      for (int k = 0; k < i; ++k) v = fmaf(-R[k * ldr + i], B[k * ldb + c], v);
      B[i * ldb + c] = v / R[i * ldr + i];
    }
  }
  __syncthreads();
}

// Solve R X = B in place (R upper n x n, B n x m): back substitution, one thread per column.
__device__ void solve_r(const float* R, int ldr, float* B, int ldb, int n, int m) {
  for (int c = threadIdx.x; c < m; c += CC_NT) {
    for (int i = n - 1; i >= 0; --i) {
      float v = B[i * ldb + c];
      for (int k = i + 1; k < n; ++k) v = fmaf(-R[i * ldr + k], B[k * ldb + c], v);
      B[i * ldb + c] = v / R[i * ldr + i];
    }
  }
  __syncthreads();
}

// B <- B R^-1 (R upper n x n, B m x n): row-wise forward substitution, one thread per row.
__device__ void solve_right_r(const float* R, int ldr, float* B, int ldb, int m, int n) {
  for (int r = threadIdx.x; r < m; r += CC_NT) {
    for (int j = 0; j < n; ++j) {
      float v = B[r * ldb + j];
      for (int k = 0; k < j; ++k) v = fmaf(-B[r * ldb + k], R[k * ldr + j], v);
      B[r * ldb + j] = v / R[j * ldr + j];
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void mod_pair_c(int s, int k, int m, int& i, int& j) {
  const int h = (s + 1) >> 1;
  const int e = (s & 1) ? 0 : 1;
  i = (h + e + k) % m;
  j = h - 1 - k;
  j %= m;
  if (j < 0) j += m;
}

// One-sided Jacobi SVD of W (m x n, m >= n, row stride ld): on exit the columns of W are
// u_j * sigma_j, V (n x n) holds the right vectors.  Returns sweeps used.
__device__ int svd_onesided(float* W, float* V, int m, int n, int ld, int* flag) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = CC_NT >> 5;
  for (int e = threadIdx.x; e < n * n; e += CC_NT) V[(e / n) * ld + (e % n)] = (e / n == e % n) ? 1.f : 0.f;
  __syncthreads();
  const int me = (n + 1) & ~1;  // even number of slots; slot n (if any) is a dummy column
  int sw = 0;
  for (; sw < 30; ++sw) {
    if (threadIdx.x == 0) *flag = 0;
    __syncthreads();
    for (int s = 0; s < me; ++s) {
      const int npairs = (s & 1) ? (me >> 1) : (me >> 1) - 1;
      for (int k = wid; k < npairs; k += nw) {
        int p, q;
        mod_pair_c(s, k, me, p, q);
        if (p >= n || q >= n) continue;
        float al = 0.f, be = 0.f, ga = 0.f;
        for (int r = lane; r < m; r += 32) {
          const float x = W[r * ld + p], y = W[r * ld + q];
          al = fmaf(x, x, al); be = fmaf(y, y, be); ga = fmaf(x, y, ga);
        }
        al = warp_sum(al); be = warp_sum(be); ga = warp_sum(ga);
        if (fabsf(ga) > 1e-7f * sqrtf(al * be) && fabsf(ga) > 1e-37f) {
          if (fabsf(ga) > 3e-7f * sqrtf(al * be) && lane == 0) *flag = 1;
          // fp64 evaluation, rounded once: keeps c^2 + s^2 = 1 unbiased (see jacobi.cu)
          const double zeta = ((double)be - (double)al) / (2.0 * (double)ga);
          const double td = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          const double cd = 1.0 / sqrt(1.0 + td * td);
          const float c = (float)cd, sn = (float)(td * cd);
          for (int r = lane; r < m; r += 32) {
            const float x = W[r * ld + p], y = W[r * ld + q];
            W[r * ld + p] = c * x - sn * y;
            W[r * ld + q] = sn * x + c * y;
          }
          for (int r = lane; r < n; r += 32) {
            const float x = V[r * ld + p], y = V[r * ld + q];
            V[r * ld + p] = c * x - sn * y;
            V[r * ld + q] = sn * x + c * y;
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

__global__ void __launch_bounds__(CC_NT)
k_cca_solve(const cpsd_cca_desc* __restrict__ descs, int dmax) {
  extern __shared__ float sm[];
  const cpsd_cca_desc t = descs[blockIdx.x];
  const int ld = dmax + 1;
  float* B0 = sm;
  float* B1 = B0 + dmax * ld;
  float* B2 = B1 + dmax * ld;
  float* B3 = B2 + dmax * ld;
  float* sig = B3 + dmax * ld;           // [dmax]
  int* rank = reinterpret_cast<int*>(sig + dmax);  // [dmax]
  __shared__ float minpiv[2];
  __shared__ int flag;

  int da = t.da_dev ? t.da_dev[0] : t.da;
  int db = t.db_dev ? t.db_dev[0] : t.db;
  da = max(0, min(da, dmax));
  db = max(0, min(db, dmax));
  const int d = min(da, db);
  if (d == 0) {
    if (threadIdx.x == 0 && t.info) { t.info[0] = 0; t.info[1] = 2; t.info[2] = 0; t.info[3] = 0; }
    return;
  }
  // load: B0 = Saa, B1 = Sbb, B2 = Sab (da x db)
  for (int e = threadIdx.x; e < da * da; e += CC_NT) {
    const int i = e / da, j = e % da;
    B0[i * ld + j] = 0.5f * (t.Saa[(long long)i * t.lds + j] + t.Saa[(long long)j * t.lds + i]);
  }
  for (int e = threadIdx.x; e < db * db; e += CC_NT) {
    const int i = e / db, j = e % db;
    B1[i * ld + j] = 0.5f * (t.Sbb[(long long)i * t.lds + j] + t.Sbb[(long long)j * t.lds + i]);
  }
  for (int e = threadIdx.x; e < da * db; e += CC_NT) {
    const int i = e / db, j = e % db;
    B2[i * ld + j] = t.Sab[(long long)i * t.lds + j];
  }
  __syncthreads();
  chol_upper(B0, da, ld, &minpiv[0]);
  chol_upper(B1, db, ld, &minpiv[1]);
  // K = Ra^-T Sab Rb^-1
  solve_rt(B0, ld, B2, ld, da, db);
  solve_right_r(B1, ld, B2, ld, da, db);

  // SVD of K.  If da < db work on K^T so the rotated matrix has m >= n columns.
  const bool tr = da < db;
  const int m = tr ? db : da, n = tr ? da : db;   // n == d
  if (tr) {
    // transpose B2 (da x db) into B3 (db x da), then swap roles
    for (int e = threadIdx.x; e < da * db; e += CC_NT) {
      const int i = e / db, j = e % db;
      B3[j * ld + i] = B2[i * ld + j];
    }
    __syncthreads();
  }
  float* W = tr ? B3 : B2;
  float* Vr = tr ? B2 : B3;
  const int sweeps = svd_onesided(W, Vr, m, n, ld, &flag);
  // singular values + ordering
  for (int j = threadIdx.x; j < n; j += CC_NT) {
    float a = 0.f;
    for (int r = 0; r < m; ++r) a = fmaf(W[r * ld + j], W[r * ld + j], a);
    sig[j] = sqrtf(a);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < n; j += CC_NT) {
    int r = 0;
    for (int u = 0; u < n; ++u) r += (sig[u] > sig[j]) || (sig[u] == sig[j] && u < j);
    rank[j] = r;
  }
  __syncthreads();
  // normalise left vectors, write rho (clamped like AlignCCA.py:282-283)
  for (int e = threadIdx.x; e < m * n; e += CC_NT) {
    const int r = e / n, j = e % n;
    W[r * ld + j] = (sig[j] > 0.f) ? W[r * ld + j] / sig[j] : 0.f;
  }
  for (int j = threadIdx.x; j < n; j += CC_NT) {
    float s = sig[j];
    s = (s < 0.f) ? 0.f : s;
    s = (s >= 1.f) ? 1.f : s;
    t.rho[rank[j]] = s;
  }
  for (int j = n + threadIdx.x; j < dmax; j += CC_NT) t.rho[j] = 0.f;
  __syncthreads();
  // U_d (da x d) and V_d (db x d), columns still in Jacobi order (rank[] sorts them)
  float* U = tr ? Vr : W;    // da x d
  float* V = tr ? W : Vr;    // db x d
  // M_b = Rb^-1 V_d   (in place), M_a = Ra^-1 U_d needs U kept for the square shortcut
  solve_r(B1, ld, V, ld, db, d);
  for (int e = threadIdx.x; e < db * d; e += CC_NT) {
    const int i = e / d, j = e % d;
    t.Mb[(long long)i * t.ldm + rank[j]] = V[i * ld + j];
  }
  __syncthreads();
  if (d == da) {
    // pinv(M_a) = U^T Ra  -> B1 (d x da);  G = M_b (U^T Ra)
    for (int e = threadIdx.x; e < d * da; e += CC_NT) {
      const int i = e / da, j = e % da;
      float a = 0.f;
      for (int k = 0; k <= j && k < da; ++k) a = fmaf(U[k * ld + i], B0[k * ld + j], a);
      B1[i * ld + j] = a;
    }
    __syncthreads();
    solve_r(B0, ld, U, ld, da, d);     // U <- M_a
  } else {
    // thin case: pinv(M_a) = (M_a^T M_a)^-1 M_a^T
    solve_r(B0, ld, U, ld, da, d);     // U <- M_a (da x d)
    for (int e = threadIdx.x; e < d * d; e += CC_NT) {
      const int i = e / d, j = e % d;
      float a = 0.f;
      for (int k = 0; k < da; ++k) a = fmaf(U[k * ld + i], U[k * ld + j], a);
      B1[i * ld + j] = a;
    }
    __syncthreads();
    chol_upper(B1, d, ld, nullptr);    // N = C^T C
    // B0 <- M_a^T (d x da), then solve C^T C Z = M_a^T
    for (int e = threadIdx.x; e < d * da; e += CC_NT) {
      const int i = e / da, j = e % da;
      B0[i * ld + j] = U[j * ld + i];
    }
    __syncthreads();
    solve_rt(B1, ld, B0, ld, d, da);
    solve_r(B1, ld, B0, ld, d, da);
    for (int e = threadIdx.x; e < d * da; e += CC_NT) {
      const int i = e / da, j = e % da;
      B1[i * ld + j] = B0[i * ld + j];
    }
    __syncthreads();
  }
  for (int e = threadIdx.x; e < da * d; e += CC_NT) {
    const int i = e / d, j = e % d;
    t.Ma[(long long)i * t.ldm + rank[j]] = U[i * ld + j];
  }
  // G (db x da) = M_b pinv(M_a); zero-fill the padding so projections can use dmax columns
  for (int e = threadIdx.x; e < dmax * dmax; e += CC_NT) {
    const int i = e / dmax, j = e % dmax;
    float a = 0.f;
    if (i < db && j < da)
      for (int k = 0; k < d; ++k) a = fmaf(V[i * ld + k], B1[k * ld + j], a);
    t.G[(long long)i * t.ldg + j] = a;
  }
  // zero-fill M padding
  for (int e = threadIdx.x; e < dmax * dmax; e += CC_NT) {
    const int i = e / dmax, j = e % dmax;
    if (j >= d || i >= da) t.Ma[(long long)i * t.ldm + j] = 0.f;
    if (j >= d || i >= db) t.Mb[(long long)i * t.ldm + j] = 0.f;
  }
  if (threadIdx.x == 0 && t.info) {
    t.info[0] = d;
    t.info[1] = (fminf(minpiv[0], minpiv[1]) < t.rank_tol) ? 1 : 0;   // rank-deficiency warning
    t.info[2] = sweeps;
    t.info[3] = 0;
  }
}

// W (C x dmax) <- leading k sorted eigenvectors with sklearn's svd_flip(u_based_decision=
// False) sign (largest-|entry| of each component positive; sklearn _pca.py:640), zero beyond k.
__global__ void __launch_bounds__(128)
k_pca_basis(const float* __restrict__ evecs, int ldv, long long strideV,
            const int* __restrict__ k_dev, const int* __restrict__ cdim, int c_fixed, int dmax,
            float* __restrict__ W, int ldw, int Cmax) {
  const int p = blockIdx.x;
  const int C = cdim ? cdim[p] : c_fixed;
  int k = k_dev[p];
  if (k > dmax) k = dmax;
  if (k > C) k = C;
  const float* V = evecs + (long long)p * strideV;
  float* Wp = W + (long long)p * Cmax * ldw;
  __shared__ float sgn[256];
  for (int j = threadIdx.x; j < dmax; j += blockDim.x) {
    float best = -1.f, s = 1.f;
    if (j < k) {
      for (int c = 0; c < C; ++c) {
        const float v = V[(long long)c * ldv + j];
        if (fabsf(v) > best) { best = fabsf(v); s = (v < 0.f) ? -1.f : 1.f; }
      }
    }
    if (j < 256) sgn[j] = s;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < Cmax * dmax; e += blockDim.x) {
    const int c = e / dmax, j = e % dmax;
    Wp[(long long)c * ldw + j] = (c < C && j < k) ? V[(long long)c * ldv + j] * sgn[j] : 0.f;
  }
}

}  // namespace

extern "C" int cpsd_cca_solve(const cpsd_cca_desc* descs_dev, int nprob, int dmax,
                              cudaStream_t stream) {
  CPSD_CHECK_ARG(nprob >= 0 && dmax > 0, "cca_solve: bad dims");
  if (nprob == 0) return CPSD_OK;
  const size_t smem = (size_t)4 * dmax * (dmax + 1) * sizeof(float) + 2 * dmax * 4 + 16;
  if (smem > 227 * 1024) {
    cpsd_set_error("cca_solve: latent dimension too large for the shared-memory solver (max 116)");
    return CPSD_ERR_UNSUPPORTED;
  }
  CPSD_CUDA(cudaFuncSetAttribute(k_cca_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_cca_solve<<<nprob, CC_NT, smem, stream>>>(descs_dev, dmax);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

extern "C" int cpsd_pca_basis(const float* evecs, int ldv, long long strideV, const int* k_dev,
                              const int* cdim, int c_fixed, int dmax, float* W, int ldw, int Cmax,
                              int nprob, cudaStream_t stream) {
  CPSD_CHECK_ARG(nprob >= 0 && dmax > 0 && dmax <= 256, "pca_basis: dmax must be in 1..256");
  if (nprob == 0) return CPSD_OK;
  k_pca_basis<<<nprob, 128, 0, stream>>>(evecs, ldv, strideV, k_dev, cdim, c_fixed, dmax, W, ldw,
                                         Cmax);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}
