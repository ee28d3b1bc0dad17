// Leading eigen-pairs of the pooled Gram by block subspace iteration.
//
// The decoder-stage PCA (decomposition/DimRedReshape.py:47-49 -> sklearn PCA with a float
// n_components) keeps only the components that explain `decoder_var` of the variance: on the
// headline workload 65 of 1144.  The total variance is the trace of the Gram, so the full
// spectrum is never needed: a 128-wide orthonormal block Q is driven to the dominant
// invariant subspace by  Y = K Q,  Q <- Y R^{-1}  (Cholesky QR, R from the fp64 Cholesky of
// Y^T Y), followed by one Rayleigh-Ritz step  H = Q^T K Q = W Theta W^T,  V = Q W.  The
// per-pair residuals ||K v - theta v|| come out of the same product ([Q;Y] W), so the host
// can accept the block, ask for more iterations or fall back to the full block-Jacobi
// solver (jacobi.cu) when the requested variance is not reached inside the block.
//
//  k_sgemm_*      : batched strided fp32 GEMM  C = alpha * op(A) B  (128x64 CTA tile, 8x4 per
//                   thread, double-buffered smem, register prefetch).  TA=1 reads A as K x M
//                   (both operands stream along their contiguous axis: the fast path used for
//                   K Q, Y^T Y and Q^T Y); TA=0 reads A as M x K.
//  k_chol_inv     : one CTA per problem, fp64 Cholesky of an m x m (m <= 128) Gram in shared
//                   memory and the inverse of its triangular factor.
//  k_topk_init    : zero the padding of K, trace, pseudo-random start block.
//  k_topk_resid   : residual norms of the Ritz pairs.
#include "common.cuh"

extern "C" int cpsd_eig_sym_small(const float* A, int lda, long long strideA, const int* n_dev,
                                  int n_fixed, int nprob, float* evals, int ld_e, float* evecs,
                                  int ldv, long long strideV, int max_sweeps, float tol,
                                  int* sweeps_out, cudaStream_t stream);

namespace {

#define SG_BM 128
#define SG_BN 64
#define SG_BK 16

// C (M x N, ldc) = alpha * op(A) * B (K x N, ldb);  TA: A is K x M (lda), else M x K (lda).
template <int TA>
__global__ void __launch_bounds__(256, 2)
k_sgemm(const float* __restrict__ A, int lda, long long strideA, const float* __restrict__ B,
        int ldb, long long strideB, float* __restrict__ C, int ldc, long long strideC, int M, int N,
        int K, float alpha) {
  __shared__ __align__(16) float As[2][SG_BK][SG_BM];
  __shared__ __align__(16) float Bs[2][SG_BK][SG_BN];
  const int prob = blockIdx.z;
  const float* Ag = A + (long long)prob * strideA;
  const float* Bg = B + (long long)prob * strideB;
  float* Cg = C + (long long)prob * strideC;
  const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const bool vecA = ((lda & 3) == 0) && ((((uintptr_t)Ag) & 15) == 0);
  const bool vecB = ((ldb & 3) == 0) && ((((uintptr_t)Bg) & 15) == 0);

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb;
  auto load_tiles = [&](int k0) {
    if (TA) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int e = tid + 256 * u;
        const int k = e >> 5, mm = (e & 31) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k0 + k < K) {
          const float* src = Ag + (long long)(k0 + k) * lda + m0 + mm;
          if (vecA && m0 + mm + 3 < M) {
            v = *reinterpret_cast<const float4*>(src);
          } else {
            if (m0 + mm + 0 < M) v.x = src[0];
            if (m0 + mm + 1 < M) v.y = src[1];
            if (m0 + mm + 2 < M) v.z = src[2];
            if (m0 + mm + 3 < M) v.w = src[3];
          }
        }
        ra[u] = v;
      }
    } else {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int e = tid + 256 * u;
        const int r = e >> 2, kq = (e & 3) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m0 + r < M) {
          const float* src = Ag + (long long)(m0 + r) * lda + k0 + kq;
          if (vecA && k0 + kq + 3 < K) {
            v = *reinterpret_cast<const float4*>(src);
          } else {
            if (k0 + kq + 0 < K) v.x = src[0];
            if (k0 + kq + 1 < K) v.y = src[1];
            if (k0 + kq + 2 < K) v.z = src[2];
            if (k0 + kq + 3 < K) v.w = src[3];
          }
        }
        ra[u] = v;
      }
    }
    {
      const int k = tid >> 4, nn = (tid & 15) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k0 + k < K) {
        const float* src = Bg + (long long)(k0 + k) * ldb + n0 + nn;
        if (vecB && n0 + nn + 3 < N) {
          v = *reinterpret_cast<const float4*>(src);
        } else {
          if (n0 + nn + 0 < N) v.x = src[0];
          if (n0 + nn + 1 < N) v.y = src[1];
          if (n0 + nn + 2 < N) v.z = src[2];
          if (n0 + nn + 3 < N) v.w = src[3];
        }
      }
      rb = v;
    }
  };
  auto store_tiles = [&](int buf) {
    if (TA) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int e = tid + 256 * u;
        const int k = e >> 5, mm = (e & 31) * 4;
        *reinterpret_cast<float4*>(&As[buf][k][mm]) = ra[u];
      }
    } else {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int e = tid + 256 * u;
        const int r = e >> 2, kq = (e & 3) * 4;
        As[buf][kq + 0][r] = ra[u].x;
        As[buf][kq + 1][r] = ra[u].y;
        As[buf][kq + 2][r] = ra[u].z;
        As[buf][kq + 3][r] = ra[u].w;
      }
    }
    const int k = tid >> 4, nn = (tid & 15) * 4;
    *reinterpret_cast<float4*>(&Bs[buf][k][nn]) = rb;
  };

  const int nk = (K + SG_BK - 1) / SG_BK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles((kt + 1) * SG_BK);
#pragma unroll
    for (int kk = 0; kk < SG_BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }

  const bool vecC = ((ldc & 3) == 0) && ((((uintptr_t)Cg) & 15) == 0) && (n0 + tx * 4 + 3 < N);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gi = m0 + ty * 8 + i;
    if (gi >= M) continue;
    float* dst = Cg + (long long)gi * ldc + n0 + tx * 4;
    if (vecC) {
      *reinterpret_cast<float4*>(dst) = make_float4(alpha * acc[i][0], alpha * acc[i][1],
                                                    alpha * acc[i][2], alpha * acc[i][3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n0 + tx * 4 + j < N) dst[j] = alpha * acc[i][j];
    }
  }
}

// ---------------------------------------------------------------------------------------
// S (m x m, fp32, symmetric, lds / strideS) -> Rinv (m x m fp32 upper triangular, ldr /
// strideR) with S = R^T R.  fp64 in shared memory.  status[prob] |= 1 when a pivot is not
// positive (rank-deficient block); the pivot is then replaced so the factor stays finite.
// ---------------------------------------------------------------------------------------
#define CH_LD 129
// S (m x m, fp64, lower 32x32 tiles only) = Y^T Y with fp64 accumulation; Y: n x m row-major.
// The Gram of the Cholesky-QR step: the block K Q has the condition number of the leading
// spectrum (1e4 and more for low-rank-plus-noise data), which an fp32 Gram squares past 1/eps.
__global__ void __launch_bounds__(256)
k_gram_cols_f64(const float* __restrict__ Y, int ldy, long long strideY, int n, int m,
                double* __restrict__ S, long long strideS) {
  __shared__ float a[32][33], b[32][33];
  int ti = 0, rem = blockIdx.x;           // lower-triangular tile index -> (ti, tj), tj <= ti
  while (rem > ti) { rem -= ti + 1; ++ti; }
  const int tj = rem;
  const float* Yg = Y + (long long)blockIdx.y * strideY;
  double* Sg = S + (long long)blockIdx.y * strideS;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int i0 = ti * 32, j0 = tj * 32;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int r0 = 0; r0 < n; r0 += 32) {
    for (int r = ty; r < 32; r += 8) {
      const bool ok = r0 + r < n;
      a[r][tx] = (ok && i0 + tx < m) ? Yg[(long long)(r0 + r) * ldy + i0 + tx] : 0.f;
      b[r][tx] = (ok && j0 + tx < m) ? Yg[(long long)(r0 + r) * ldy + j0 + tx] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < 32; ++r) {
      const double bv = (double)b[r][tx];
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[e] = fma((double)a[r][ty + 8 * e], bv, acc[e]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int i = i0 + ty + 8 * e, j = j0 + tx;
    if (i < m && j < m) Sg[(long long)i * m + j] = acc[e];
  }
}

#define CH_NT 1024
// TI = float: symmetric fp32 Gram (both triangles averaged); TI = double: lower triangle of an
// fp64 Gram (k_gram_cols_f64)
template <typename TI>
__global__ void __launch_bounds__(CH_NT)
k_chol_inv(const TI* __restrict__ S, int lds, long long strideS, int m,
           float* __restrict__ Rinv, int ldr, long long strideR, int* __restrict__ status) {
  // P: strict lower triangle + diagonal = Cholesky factor L; afterwards the strict upper
  // triangle receives X^T, X = L^{-1} (so P's upper triangle is R^{-1} off the diagonal)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* P = reinterpret_cast<double*>(smem_raw);          // m x CH_LD
  __shared__ double xd[128];                                // diagonal of X = 1 / diag(L)
  __shared__ double s_piv, s_inv;
  __shared__ int s_bad;
  const int prob = blockIdx.x;
  const TI* Sg = S + (long long)prob * strideS;
  const int tid = threadIdx.x;
  const int tx = tid & 31, ty = tid >> 5;
  if (tid == 0) s_bad = 0;
  for (int e = tid; e < m * m; e += CH_NT) {
    const int r = e / m, c = e - r * m;
    double v = 0.0;
    if (c <= r) {
      if (sizeof(TI) == 8) v = (double)Sg[(long long)r * lds + c];
      else v = 0.5 * ((double)Sg[(long long)r * lds + c] + (double)Sg[(long long)c * lds + r]);
    }
    P[r * CH_LD + c] = v;
  }
  __syncthreads();
  double dmax = 0.0;
  for (int i = 0; i < m; ++i) dmax = fmax(dmax, P[i * CH_LD + i]);
  const double floor_piv = dmax * 1e-13 + 1e-290;

  for (int j = 0; j < m; ++j) {
    if (tid == 0) {
      double d = P[j * CH_LD + j];
      if (!(d > floor_piv)) {
        s_bad = 1;
        d = floor_piv;
      }
      s_piv = sqrt(d);
      s_inv = 1.0 / s_piv;       // one fp64 division per column, not one per thread (32 warps
    }                            // queueing on the fp64 pipe for the same quotient)
    __syncthreads();
    const double ljj = s_piv;
    const double inv = s_inv;
    for (int i = j + tid; i < m; i += CH_NT) P[i * CH_LD + j] = (i == j) ? ljj : P[i * CH_LD + j] * inv;
    __syncthreads();
    // trailing update of the lower triangle: L[i][k] -= L[i][j] L[k][j], j < k <= i < m
    for (int i = j + 1 + ty; i < m; i += 32) {
      const double lij = P[i * CH_LD + j];
      for (int k = j + 1 + tx; k <= i; k += 32) P[i * CH_LD + k] -= lij * P[k * CH_LD + j];
    }
    __syncthreads();
  }
  // X = L^{-1} (lower), right-looking: row k is final once scaled by 1/l_kk, then it is
  // eliminated from every row below.  X[i][c] (c < i) lives at P[c][i].
  for (int k = tid; k < m; k += CH_NT) xd[k] = 1.0 / P[k * CH_LD + k];   // all reciprocal pivots at once
  __syncthreads();
  for (int k = 0; k < m; ++k) {
    const double inv = xd[k];
    for (int c = tid; c < k; c += CH_NT) P[c * CH_LD + k] *= inv;
    __syncthreads();
    for (int i = k + 1 + ty; i < m; i += 32) {
      const double lik = P[i * CH_LD + k];
      for (int c = tx; c <= k; c += 32) {
        const double xkc = (c == k) ? inv : P[c * CH_LD + k];
        P[c * CH_LD + i] -= lik * xkc;
      }
    }
    __syncthreads();
  }
  // Rinv = (L^T)^{-1} = X^T (upper)
  float* Rg = Rinv + (long long)prob * strideR;
  for (int e = tid; e < m * m; e += CH_NT) {
    const int r = e / m, c = e - r * m;
    Rg[(long long)r * ldr + c] = (r < c) ? (float)P[r * CH_LD + c] : (r == c ? (float)xd[r] : 0.f);
  }
  if (tid == 0 && s_bad && status) atomicOr(&status[prob], 1);
}

// ---------------------------------------------------------------------------------------
// SPD solve  S W = B  (S m x m fp64, B m x q fp64, both row-major) by Cholesky in shared
// memory: W = S^{-1} B written as fp32 (ldw).  Least-squares read-in matrices of JointPCA
// (alignment/JointPCA.py:203-206: pinv(X_p) @ latent = (X_p^T X_p)^{-1} X_p^T latent for a
// full-column-rank X_p).  B is used as workspace.  status |= 1 on a non-positive pivot.
__global__ void __launch_bounds__(CH_NT)
k_chol_solve_f64(const double* __restrict__ S, int lds, long long strideS, int m,
                 double* __restrict__ B, int ldb, long long strideB, int q,
                 float* __restrict__ W, int ldw, long long strideW, int* __restrict__ status,
                 double* __restrict__ ws) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // lower factor, m rows of ldp doubles: shared memory for m <= 128, else this problem's
  // slice of the caller's workspace (L2-resident; patients with more than 128 electrodes)
  const int ldp = ws ? m + 1 : CH_LD;
  double* P = ws ? ws + (long long)blockIdx.x * m * (m + 1) : reinterpret_cast<double*>(smem_raw);
  __shared__ double s_piv;
  __shared__ int s_bad;
  const int prob = blockIdx.x;
  const double* Sg = S + (long long)prob * strideS;
  double* Bg = B + (long long)prob * strideB;
  float* Wg = W + (long long)prob * strideW;
  const int tid = threadIdx.x;
  const int tx = tid & 31, ty = tid >> 5;
  if (tid == 0) s_bad = 0;
  for (int e = tid; e < m * m; e += CH_NT) {
    const int r = e / m, c = e - r * m;
    P[r * ldp + c] = (c <= r) ? 0.5 * (Sg[(long long)r * lds + c] + Sg[(long long)c * lds + r]) : 0.0;
  }
  __syncthreads();
  double dmax = 0.0;
  for (int i = 0; i < m; ++i) dmax = fmax(dmax, P[i * ldp + i]);
  const double floor_piv = dmax * 1e-14 + 1e-290;
  for (int j = 0; j < m; ++j) {
    if (tid == 0) {
      double d = P[j * ldp + j];
      if (!(d > floor_piv)) {
        s_bad = 1;
        d = floor_piv;
      }
      s_piv = sqrt(d);
    }
    __syncthreads();
    const double ljj = s_piv;
    const double inv = 1.0 / ljj;
    for (int i = j + tid; i < m; i += CH_NT) P[i * ldp + j] = (i == j) ? ljj : P[i * ldp + j] * inv;
    __syncthreads();
    for (int i = j + 1 + ty; i < m; i += 32) {
      const double lij = P[i * ldp + j];
      for (int k = j + 1 + tx; k <= i; k += 32) P[i * ldp + k] -= lij * P[k * ldp + j];
    }
    __syncthreads();
  }
  // one thread per right-hand side: forward then back substitution, in place in B
  for (int c = tid; c < q; c += CH_NT) {
    for (int i = 0; i < m; ++i) {
      double acc = Bg[(long long)i * ldb + c];
      for (int k = 0; k < i; ++k) acc -= P[i * ldp + k] * Bg[(long long)k * ldb + c];
      Bg[(long long)i * ldb + c] = acc / P[i * ldp + i];
    }
    for (int i = m - 1; i >= 0; --i) {
      double acc = Bg[(long long)i * ldb + c];
      for (int k = i + 1; k < m; ++k) acc -= P[k * ldp + i] * Bg[(long long)k * ldb + c];
      acc /= P[i * ldp + i];
      Bg[(long long)i * ldb + c] = acc;
      Wg[(long long)i * ldw + c] = (float)acc;
    }
  }
  if (tid == 0 && s_bad && status) atomicOr(&status[prob], 1);
}

// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float hash_uniform(unsigned int x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return (float)(x >> 8) * (1.0f / 8388608.0f) - 1.0f;    // [-1, 1)
}

// grid (16, nprob): zero rows / columns >= n of K; CTA 0 also writes the trace and status = 0.
// Q (n_pad x m, row stride m): pseudo-random start block (identical for every problem), rows
// >= n zero.
__global__ void __launch_bounds__(256)
k_topk_init(float* __restrict__ K, int ld, long long stride, int n_pad, const int* __restrict__ n_dev,
            int n_fixed, float* __restrict__ Q, long long strideQ, int m, float* __restrict__ total,
            int* __restrict__ status, int init_q) {
  __shared__ double red[33];
  const int prob = blockIdx.y;
  int n = n_dev ? n_dev[prob] : n_fixed;
  n = max(0, min(n, n_pad));
  float* Kg = K + (long long)prob * stride;
  const int rows_per = (n_pad + gridDim.x - 1) / gridDim.x;
  const int r0 = blockIdx.x * rows_per, r1 = min(n_pad, r0 + rows_per);
  for (int r = r0; r < r1; ++r) {
    if (r >= n) {
      for (int c = threadIdx.x; c < n_pad; c += 256) Kg[(long long)r * ld + c] = 0.f;
    } else {
      for (int c = n + threadIdx.x; c < n_pad; c += 256) Kg[(long long)r * ld + c] = 0.f;
    }
  }
  if (init_q) {
    float* Qg = Q + (long long)prob * strideQ;
    for (int r = r0; r < r1; ++r)
      for (int c = threadIdx.x; c < m; c += 256)
        Qg[(long long)r * m + c] = (r < n) ? hash_uniform((unsigned)(r * 131 + c) * 2654435761u + 12345u)
                                           : 0.f;
  }
  if (blockIdx.x == 0) {
    double t = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) t += (double)Kg[(long long)i * ld + i];
    t = block_sum(t, red);
    if (threadIdx.x == 0) {
      total[prob] = (float)t;
      if (init_q) status[prob] = 0;
    }
  }
}

// H <- (H + H^T)/2 in place (m x m, row stride m)
__global__ void __launch_bounds__(256)
k_symmetrize(float* __restrict__ H, int m, long long stride) {
  float* Hg = H + (long long)blockIdx.x * stride;
  for (int e = threadIdx.x; e < m * m; e += 256) {
    const int r = e / m, c = e - r * m;
    if (r < c) {
      const float v = 0.5f * (Hg[r * m + c] + Hg[c * m + r]);
      Hg[r * m + c] = v;
      Hg[c * m + r] = v;
    }
  }
}

// VV (2*n_pad x m): rows [0, n_pad) = V, rows [n_pad, 2 n_pad) = K V.
// resid[prob][j] = || K v_j - theta_j v_j ||_2
__global__ void __launch_bounds__(256)
k_topk_resid(const float* __restrict__ VV, long long strideVV, int n_pad, int m,
             const float* __restrict__ evals, int ld_e, float* __restrict__ resid) {
  __shared__ float part[256];
  const int prob = blockIdx.x;
  const float* V = VV + (long long)prob * strideVV;
  const float* KV = V + (long long)n_pad * m;
  // 256 threads = (256 / m_lanes) row phases x m columns; m <= 128 -> >= 2 phases
  const int j = threadIdx.x % m, ph = threadIdx.x / m;
  const int nph = 256 / m;
  float acc = 0.f;
  if (ph < nph) {
    const float th = evals[(long long)prob * ld_e + j];
    for (int i = ph; i < n_pad; i += nph) {
      const float d = KV[(long long)i * m + j] - th * V[(long long)i * m + j];
      acc = fmaf(d, d, acc);
    }
  }
  part[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < m) {
    float t = 0.f;
    for (int p = 0; p < nph; ++p) t += part[p * m + threadIdx.x];
    resid[(long long)prob * m + threadIdx.x] = sqrtf(t);
  }
}

template <int TA>
int launch_sgemm(const float* A, int lda, long long sA, const float* B, int ldb, long long sB,
                 float* C, int ldc, long long sC, int M, int N, int K, float alpha, int nprob,
                 cudaStream_t stream) {
  dim3 grid((N + SG_BN - 1) / SG_BN, (M + SG_BM - 1) / SG_BM, nprob);
  k_sgemm<TA><<<grid, 256, 0, stream>>>(A, lda, sA, B, ldb, sB, C, ldc, sC, M, N, K, alpha);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

}  // namespace

#define CPSD_TRY(call)            \
  do {                            \
    int st__ = (call);            \
    if (st__ != CPSD_OK) return st__; \
  } while (0)

// C = alpha * op(A) B, batched with element strides.  trans_a != 0: A is stored K x M.
extern "C" int cpsd_sgemm_batched(int trans_a, int M, int N, int K, float alpha, const float* A,
                                  int lda, long long strideA, const float* B, int ldb,
                                  long long strideB, float* C, int ldc, long long strideC, int nprob,
                                  cudaStream_t stream) {
  CPSD_CHECK_ARG(M > 0 && N > 0 && K > 0 && nprob >= 0, "sgemm_batched: bad dims");
  CPSD_CHECK_ARG(nprob <= 65535, "sgemm_batched: nprob > 65535");
  if (nprob == 0) return CPSD_OK;
  if (trans_a) return launch_sgemm<1>(A, lda, strideA, B, ldb, strideB, C, ldc, strideC, M, N, K, alpha, nprob, stream);
  return launch_sgemm<0>(A, lda, strideA, B, ldb, strideB, C, ldc, strideC, M, N, K, alpha, nprob, stream);
}

extern "C" int cpsd_chol_inv(const float* S, int lds, long long strideS, int m, float* Rinv, int ldr,
                             long long strideR, int* status, int nprob, cudaStream_t stream) {
  CPSD_CHECK_ARG(m > 0 && m <= 128, "chol_inv: m must be in 1..128");
  if (nprob == 0) return CPSD_OK;
  const size_t smem = 128 * CH_LD * sizeof(double);
  CPSD_CUDA(cudaFuncSetAttribute(k_chol_inv<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_chol_inv<float><<<nprob, CH_NT, smem, stream>>>(S, lds, strideS, m, Rinv, ldr, strideR, status);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

// Rinv = chol(Y^T Y)^{-T} with the Gram accumulated in fp64 (S64: nprob * m * m doubles of
// workspace): the orthonormalisation step of the subspace iteration.
static int chol_qr_factor(const float* Y, int ldy, long long strideY, int n, int m, double* S64,
                          float* Rinv, int ldr, long long strideR, int* status, int nprob,
                          cudaStream_t stream) {
  const int nt = (m + 31) / 32;
  k_gram_cols_f64<<<dim3(nt * (nt + 1) / 2, nprob), 256, 0, stream>>>(Y, ldy, strideY, n, m, S64,
                                                                      (long long)m * m);
  CPSD_LAUNCH_CHECK();
  const size_t smem = 128 * CH_LD * sizeof(double);
  CPSD_CUDA(cudaFuncSetAttribute(k_chol_inv<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_chol_inv<double><<<nprob, CH_NT, smem, stream>>>(S64, m, (long long)m * m, m, Rinv, ldr, strideR, status);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

extern "C" int cpsd_chol_solve_f64(const double* S, int lds, long long strideS, int m, double* B,
                                   int ldb, long long strideB, int q, float* W, int ldw,
                                   long long strideW, int* status, int nprob, cudaStream_t stream) {
  CPSD_CHECK_ARG(m > 0 && m <= 128 && q > 0, "chol_solve_f64: m must be in 1..128 (wider: cpsd_chol_solve_f64_ws)");
  if (nprob == 0) return CPSD_OK;
  const size_t smem = 128 * CH_LD * sizeof(double);
  CPSD_CUDA(cudaFuncSetAttribute(k_chol_solve_f64, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
  k_chol_solve_f64<<<nprob, CH_NT, smem, stream>>>(S, lds, strideS, m, B, ldb, strideB, q, W, ldw,
                                                   strideW, status, nullptr);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

// Any m: the Cholesky factor lives in the caller's workspace (nprob * m * (m + 1) doubles)
// instead of shared memory -- the read-in matrices of patients with more than 128 electrodes.
extern "C" int cpsd_chol_solve_f64_ws(const double* S, int lds, long long strideS, int m, double* B,
                                      int ldb, long long strideB, int q, float* W, int ldw,
                                      long long strideW, int* status, double* ws, int nprob,
                                      cudaStream_t stream) {
  CPSD_CHECK_ARG(m > 0 && q > 0 && ws != nullptr, "chol_solve_f64_ws: bad dims / workspace");
  if (nprob == 0) return CPSD_OK;
  k_chol_solve_f64<<<nprob, CH_NT, 0, stream>>>(S, lds, strideS, m, B, ldb, strideB, q, W, ldw,
                                                strideW, status, ws);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

// workspace floats: per problem  Q,Y (2 n_pad m)  +  VV (2 n_pad m)  +  S, Rinv, H, W (4 m m)
extern "C" long long cpsd_eig_topk_ws_elems(int n_pad, int m, int nprob) {
  return (long long)nprob * (4LL * n_pad * m + 5LL * m * m);     // S holds doubles
}

// Leading m (<= 128) eigen-pairs of nprob symmetric PSD matrices K (n_pad x n_pad, leading
// n x n block used, the padding is zeroed) by `iters` steps of block subspace iteration +
// Rayleigh-Ritz.
//   init != 0 : start from the pseudo-random block;  init == 0 : continue from the Ritz vectors
//               a previous call left in the workspace (more iterations on the same K).
// Outputs: evals[prob][0..m) descending (entries m..ld_e zeroed), V = ws + voff (n_pad x m per
// problem, row stride m, problem stride 2*n_pad*m; use cpsd_eig_topk_voff()), total[prob] =
// trace, resid[prob][j] = ||K v_j - theta_j v_j||, status[prob] bit 0 = rank-deficient block.
extern "C" long long cpsd_eig_topk_voff(int n_pad, int m, int nprob) {
  return (long long)nprob * 2LL * n_pad * m;
}

extern "C" int cpsd_topk_tc_split_k(const float* K, int ld, long long stride, int n_pad, int nprob,
                                    float* tc_ws, cudaStream_t stream);
extern "C" int cpsd_topk_tc_kq(const float* Q, long long strideQ, float* Y, long long strideY, int n_pad,
                               int nprob, int terms, float* tc_ws, const void* map_dev,
                               cudaStream_t stream);

static int eig_sym_topk_impl(float* K, int ld, long long stride, int n_pad, const int* n_dev,
                             int n_fixed, int nprob, int m, int iters, int init, float* ws,
                             float* evals, int ld_e, float* total, float* resid, int* status,
                             int eig_sweeps, float eig_tol, float* tc_ws, const void* map_dev,
                             int tf32_iters, int f64_gram, cudaStream_t stream) {
  CPSD_CHECK_ARG(n_pad > 0 && ld >= n_pad, "eig_sym_topk: bad dims");
  CPSD_CHECK_ARG(m > 0 && m <= 128 && (m % 4) == 0 && m <= n_pad, "eig_sym_topk: m must be a multiple of 4 in 4..128");
  CPSD_CHECK_ARG(ld_e >= m && iters >= 0, "eig_sym_topk: bad ld_e / iters");
  if (nprob == 0) return CPSD_OK;
  const long long sQY = 2LL * n_pad * m, sMM = (long long)m * m;
  float* QY = ws;                               // [prob][Q | Y]
  float* VV = QY + (long long)nprob * sQY;      // [prob][V | KV]
  double* S = reinterpret_cast<double*>(VV + (long long)nprob * sQY);   // fp64 Gram (8-byte aligned:
  float* Rinv = reinterpret_cast<float*>(S) + 2LL * nprob * sMM;        //  every offset is even)
  float* H = Rinv + (long long)nprob * sMM;
  float* W = H + (long long)nprob * sMM;
  float* Q = QY;
  float* Y = QY + (long long)n_pad * m;

  k_topk_init<<<dim3(16, nprob), 256, 0, stream>>>(K, ld, stride, n_pad, n_dev, n_fixed, Q, sQY, m,
                                                   total, status, init);
  CPSD_LAUNCH_CHECK();
  const bool tc = tc_ws != nullptr;
  if (tc) {
    CPSD_CHECK_ARG(m == 128 && n_pad % 128 == 0 && ld == n_pad && stride == (long long)n_pad * n_pad,
                   "eig_sym_topk_tc: needs m = 128 and densely packed K with n_pad % 128 == 0");
    CPSD_TRY(cpsd_topk_tc_split_k(K, ld, stride, n_pad, nprob, tc_ws, stream));
  }
  // Y = K Q: tensor cores (single-pass TF32 while the block is still far from converged -- the
  // iteration is self-correcting -- then 3xTF32) or the fp32 SIMT GEMM
  auto kq = [&](int terms) -> int {
    if (tc) return cpsd_topk_tc_kq(Q, sQY, Y, sQY, n_pad, nprob, terms, tc_ws, map_dev, stream);
    return launch_sgemm<1>(K, ld, stride, Q, m, sQY, Y, m, sQY, n_pad, m, n_pad, 1.f, nprob, stream);
  };
  if (!init) {
    // resume from the Ritz vectors (orthonormal up to rounding)
    CPSD_CUDA(cudaMemcpy2DAsync(Q, sQY * sizeof(float), VV, sQY * sizeof(float),
                                (size_t)n_pad * m * sizeof(float), nprob, cudaMemcpyDeviceToDevice,
                                stream));
  }
  auto orth = [&](const float* src) -> int {   // Q <- src * chol(src^T src)^{-1}
    if (f64_gram) return chol_qr_factor(src, m, sQY, n_pad, m, S, Rinv, m, sMM, status, nprob, stream);
    float* S32 = reinterpret_cast<float*>(S);
    CPSD_TRY(launch_sgemm<1>(src, m, sQY, src, m, sQY, S32, m, sMM, m, m, n_pad, 1.f, nprob, stream));
    CPSD_TRY(cpsd_chol_inv(S32, m, sMM, m, Rinv, m, sMM, status, nprob, stream));
    return CPSD_OK;
  };
  if (init) {
    // orthonormalise the start block: Y <- Q, Q <- Y Rinv
    CPSD_CUDA(cudaMemcpy2DAsync(Y, sQY * sizeof(float), Q, sQY * sizeof(float),
                                (size_t)n_pad * m * sizeof(float), nprob, cudaMemcpyDeviceToDevice,
                                stream));
    CPSD_TRY(orth(Y));
    CPSD_TRY(launch_sgemm<0>(Y, m, sQY, Rinv, m, sMM, Q, m, sQY, n_pad, m, m, 1.f, nprob, stream));
  }
  for (int it = 0; it < iters; ++it) {
    CPSD_TRY(kq((init && it < tf32_iters) ? 1 : 3));
    CPSD_TRY(orth(Y));
    CPSD_TRY(launch_sgemm<0>(Y, m, sQY, Rinv, m, sMM, Q, m, sQY, n_pad, m, m, 1.f, nprob, stream));
  }
  // second orthonormalisation pass (Cholesky QR loses eps * cond^2), then Rayleigh-Ritz
  CPSD_CUDA(cudaMemcpy2DAsync(Y, sQY * sizeof(float), Q, sQY * sizeof(float),
                              (size_t)n_pad * m * sizeof(float), nprob, cudaMemcpyDeviceToDevice,
                              stream));
  CPSD_TRY(orth(Y));
  CPSD_TRY(launch_sgemm<0>(Y, m, sQY, Rinv, m, sMM, Q, m, sQY, n_pad, m, m, 1.f, nprob, stream));
  CPSD_TRY(kq(3));
  CPSD_TRY(launch_sgemm<1>(Q, m, sQY, Y, m, sQY, H, m, sMM, m, m, n_pad, 1.f, nprob, stream));
  k_symmetrize<<<nprob, 256, 0, stream>>>(H, m, sMM);
  CPSD_LAUNCH_CHECK();
  CPSD_TRY(cpsd_eig_sym_small(H, m, sMM, nullptr, m, nprob, evals, ld_e, W, m, sMM, eig_sweeps,
                              eig_tol, nullptr, stream));
  // [V ; K V] = [Q ; Y] W
  CPSD_TRY(launch_sgemm<0>(QY, m, sQY, W, m, sMM, VV, m, sQY, 2 * n_pad, m, m, 1.f, nprob, stream));
  k_topk_resid<<<nprob, 256, 0, stream>>>(VV, sQY, n_pad, m, evals, ld_e, resid);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

extern "C" int cpsd_eig_sym_topk(float* K, int ld, long long stride, int n_pad, const int* n_dev,
                                 int n_fixed, int nprob, int m, int iters, int init, float* ws,
                                 float* evals, int ld_e, float* total, float* resid, int* status,
                                 int eig_sweeps, float eig_tol, int f64_gram, cudaStream_t stream) {
  return eig_sym_topk_impl(K, ld, stride, n_pad, n_dev, n_fixed, nprob, m, iters, init, ws, evals, ld_e,
                           total, resid, status, eig_sweeps, eig_tol, nullptr, nullptr, 0, f64_gram,
                           stream);
}

// f64_gram != 0: the Gram of every Cholesky-QR step is accumulated in fp64 (k_gram_cols_f64) --
// needed when the leading spectrum spans more than ~3e3 (low-rank signal over a noise floor),
// where an fp32 Gram of K Q is no longer positive definite; costs ~0.5 ms per 138 problems and
// step, so callers try the fp32 Gram first and switch when status reports a failed pivot.
// Same solver with K Q on the tensor cores (tcgen05, tc_gram.cu): tc_ws = cpsd_topk_tc_ws_elems()
// floats, map_dev = the tensor maps written by cpsd_topk_tc_encode for this K / tc_ws; the
// first tf32_iters iterations of a fresh start use single-pass TF32, the rest 3xTF32.
extern "C" int cpsd_eig_sym_topk_tc(float* K, int ld, long long stride, int n_pad, const int* n_dev,
                                    int n_fixed, int nprob, int m, int iters, int init, float* ws,
                                    float* evals, int ld_e, float* total, float* resid, int* status,
                                    int eig_sweeps, float eig_tol, float* tc_ws, const void* map_dev,
                                    int tf32_iters, int f64_gram, cudaStream_t stream) {
  CPSD_CHECK_ARG(tc_ws != nullptr && map_dev != nullptr, "eig_sym_topk_tc: tc_ws / map_dev is NULL");
  return eig_sym_topk_impl(K, ld, stride, n_pad, n_dev, n_fixed, nprob, m, iters, init, ws, evals, ld_e,
                           total, resid, status, eig_sweeps, eig_tol, tc_ws, map_dev, tf32_iters,
                           f64_gram, stream);
}
