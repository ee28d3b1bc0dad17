// Streaming (HBM-bound) stages and the small glue kernels between the solvers.
//
//  k_class_mean     condition averaging: mean over the trials of each class
//                   (reference alignment/alignment_utils.py:42-61 cnd_avg, :12-39
//                   extract_group_conditions).  Trials stay in their native
//                   [trial][time][channel] layout; each thread owns a float4 of the
//                   (time, channel) plane and walks the class's member list.
//  k_center_rows    subtract a column-mean vector from a row block in place.
//  k_copy_rows      strided batched 2-D copy.
//  k_mcca_*         assemble / back-transform the MCCA generalised eigenproblem
//                   (mvlearn _i_mcca 'gevp' form used through AlignMCCA.py:152-153).
//  k_scores_*       PCA scores of train / test trials from the Gram eigen-pairs.
#include "common.cuh"
#include "descs.h"

namespace {

__global__ void __launch_bounds__(256)
k_class_mean(const cpsd_class_mean_desc* __restrict__ descs) {
  const cpsd_class_mean_desc d = descs[blockIdx.z];
  const int slot = blockIdx.y;
  if (slot >= d.nslot) return;
  const int b = d.member_ptr[slot], e = d.member_ptr[slot + 1];
  const float inv = (e > b) ? 1.f / (float)(e - b) : 0.f;
  float* o = d.out + (long long)slot * d.TC;
  if ((d.TC & 3) == 0 && ((((uintptr_t)d.X) | ((uintptr_t)d.out)) & 15) == 0) {
    const int n4 = d.TC >> 2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int m = b; m < e; ++m) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(d.X + (long long)d.members[m] * d.TC) + i);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
      reinterpret_cast<float4*>(o)[i] = acc;
    }
  } else {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.TC; i += gridDim.x * blockDim.x) {
      float acc = 0.f;
      for (int m = b; m < e; ++m) acc += d.X[(long long)d.members[m] * d.TC + i];
      o[i] = acc * inv;
    }
  }
}

// Z[p][r][c] -= mu[p][c]   r < nrows[p]
__global__ void __launch_bounds__(256)
k_center_rows(float* __restrict__ Z, int ld, long long strideZ, const float* __restrict__ mu,
              int ldmu, const int* __restrict__ nrows_dev, int nrows_fixed, int ncols) {
  const int p = blockIdx.z;
  const int nrows = nrows_dev ? nrows_dev[p] : nrows_fixed;
  float* Zp = Z + (long long)p * strideZ;
  const float* mp = mu + (long long)p * ldmu;
  for (int r = blockIdx.y; r < nrows; r += gridDim.y) {
    float* row = Zp + (long long)r * ld;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ncols; c += gridDim.x * blockDim.x)
      row[c] -= mp[c];
  }
}

// dst[p][r][c] = src[p][r0[p] + r][c]
__global__ void __launch_bounds__(256)
k_copy_rows(const float* __restrict__ src, int lds, long long strideS, float* __restrict__ dst,
            int ldd, long long strideD, const int* __restrict__ r0_dev, int r0_fixed, int nrows,
            int ncols, int src_rows) {
  const int p = blockIdx.z;
  const int r0 = r0_dev ? r0_dev[p] : r0_fixed;
  for (int r = blockIdx.y; r < nrows; r += gridDim.y) {
    const bool in = (r0 + r) < src_rows;       // rows past the source block read as zero
    const float* s = src + (long long)p * strideS + (long long)(r0 + r) * lds;
    float* d = dst + (long long)p * strideD + (long long)r * ldd;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ncols; c += gridDim.x * blockDim.x)
      d[c] = in ? s[c] : 0.f;
  }
}

// fp64 host data (the reference's native dtype) -> fp32 device layout
__global__ void __launch_bounds__(256)
k_cast_f64_f32(const double* __restrict__ src, float* __restrict__ dst, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = (float)src[i];
}

// dst[p][r][j] = src[p][r][perm[p][j]]   j < ncols, r < nrows
__global__ void __launch_bounds__(256)
k_permute_cols(const float* __restrict__ src, int lds, long long strideS,
               const int* __restrict__ perm, int ld_perm, float* __restrict__ dst, int ldd,
               long long strideD, int nrows, int ncols) {
  const int p = blockIdx.z;
  for (int r = blockIdx.y; r < nrows; r += gridDim.y) {
    const float* s = src + (long long)p * strideS + (long long)r * lds;
    float* d = dst + (long long)p * strideD + (long long)r * ldd;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < ncols; j += gridDim.x * blockDim.x)
      d[j] = s[perm[(long long)p * ld_perm + j]];
  }
}

// ------------------------------------------------------------------------------- MCCA
// Per (fold, view): keep the leading r = min(R, rank, n_valid) principal directions of the
// centred condition averages.  Vr (C x R, zero beyond r), d2 [R] squared singular values.
__global__ void __launch_bounds__(128)
k_mcca_mask(const float* __restrict__ evecs, int ldv, long long strideV,
            const float* __restrict__ evals, int ld_e, const int* __restrict__ rank,
            const int* __restrict__ cdim, const int* __restrict__ src_idx, int R, int Cmax,
            float* __restrict__ Vr, float* __restrict__ d2, int* __restrict__ r_eff) {
  const int p = blockIdx.x;  // fold * P + view
  const int sp = src_idx ? src_idx[p] : p;   // where this problem's eigen-pairs live
  const int C = cdim[p];
  int r = rank ? rank[p] : R;
  if (r > R) r = R;
  if (r > C) r = C;
  if (r < 0) r = 0;
  const float* V = evecs + (long long)sp * strideV;
  float* o = Vr + (long long)p * Cmax * R;
  for (int e = threadIdx.x; e < Cmax * R; e += blockDim.x) {
    const int c = e / R, j = e - c * R;
    o[e] = (c < C && j < r) ? V[(long long)c * ldv + j] : 0.f;
  }
  for (int j = threadIdx.x; j < R; j += blockDim.x)
    d2[(long long)p * R + j] = (j < r) ? fmaxf(evals[(long long)sp * ld_e + j], 0.f) : 0.f;
  if (threadIdx.x == 0) r_eff[p] = r;
}

// Per fold: compact the P*R reduced coordinates to the n = sum_v r_v valid ones and write
// the whitened SUMCOR matrix  M = Dh^-1/2 LHS Dh^-1/2  (unit diagonal, zero inside a view,
// scaled cross-scatter between views).  reg < 0 means regs=None.
// Split source (Gx != NULL): G holds only the target's R rows [G_tt | G_tx] per fold and the
// cross x cross block comes from slot xslot[f] of Gx (fold-invariant, cached).
struct GzView {
  const float* Gf;
  const float* Gx;
  int ldg, ldx, R;
  __device__ __forceinline__ float at(int ia, int ib) const {
    if (!Gx) return Gf[(long long)ia * ldg + ib];
    if (ia < R) return Gf[(long long)ia * ldg + ib];
    if (ib < R) return Gf[(long long)ib * ldg + ia];
    return Gx[(long long)(ia - R) * ldx + (ib - R)];
  }
};

__global__ void __launch_bounds__(256)
k_mcca_build(const float* __restrict__ G, int ldg, long long strideG, const float* __restrict__ Gxx,
             int ldx, long long strideX, const int* __restrict__ xslot,
             const int* __restrict__ r_eff, int P, int R, float reg, float* __restrict__ M, int ldm,
             long long strideM, int* __restrict__ n_out, int* __restrict__ cidx,
             float* __restrict__ dh, int n_comp, int* __restrict__ status) {
  extern __shared__ int sh[];
  int* cmap = sh;                   // compact -> padded index
  float* dhs = reinterpret_cast<float*>(sh + P * R);
  __shared__ int n_s;
  const int f = blockIdx.x;
  GzView Gv;
  Gv.Gf = G + (long long)f * strideG;
  Gv.Gx = Gxx ? Gxx + (long long)xslot[f] * strideX : nullptr;
  Gv.ldg = ldg; Gv.ldx = ldx; Gv.R = R;
  if (threadIdx.x == 0) {
    int n = 0;
    for (int v = 0; v < P; ++v) {
      const int r = r_eff[f * P + v];
      for (int j = 0; j < r; ++j) cmap[n++] = v * R + j;
    }
    n_s = n;
    n_out[f] = n;
    if (status && n < n_comp) status[f] = 1;
  }
  __syncthreads();
  const int n = n_s;
  for (int a = threadIdx.x; a < P * R; a += blockDim.x) {
    float d = 0.f;
    if (a < n) {
      const int ia = cmap[a];
      const float gaa = Gv.at(ia, ia);
      d = (reg >= 0.f) ? (1.f - reg) * gaa + reg : gaa;
      dhs[a] = d;
    }
    cidx[(long long)f * P * R + a] = (a < n) ? cmap[a] : -1;
    dh[(long long)f * P * R + a] = d;
  }
  __syncthreads();
  float* Mf = M + (long long)f * strideM;
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    const int a = e / n, b = e - a * n;
    const int ia = cmap[a], ib = cmap[b];
    float v;
    if (a == b) v = 1.f;
    else if (ia / R == ib / R) v = 0.f;
    else v = Gv.at(ia, ib) * rsqrtf(dhs[a] * dhs[b]);
    Mf[(long long)a * ldm + b] = v;
  }
}

// loadings[f][v] (Cmax x n_comp) = Vr[f][v] (Cmax x R) @ ( Dh^-1/2 u )[rows of view v]
// u: eigenvectors of M; column q is U[row][perm ? perm[q] : q].
__global__ void __launch_bounds__(128)
k_mcca_loadings(const float* __restrict__ Vr, const float* __restrict__ U, int ldu,
                long long strideU, const int* __restrict__ perm, int ld_perm,
                const int* __restrict__ r_eff, const float* __restrict__ dh, int P, int R,
                int Cmax, int n_comp, float* __restrict__ L, int ldl) {
  const int p = blockIdx.x;  // fold * P + view
  const int f = p / P, v = p - f * P;
  int off = 0;
  for (int u = 0; u < v; ++u) off += r_eff[f * P + u];
  const int r = r_eff[p];
  const float* Uf = U + (long long)f * strideU;
  const float* Vp = Vr + (long long)p * Cmax * R;
  const float* dhf = dh + (long long)f * P * R;
  float* Lp = L + (long long)p * Cmax * ldl;
  for (int e = threadIdx.x; e < Cmax * n_comp; e += blockDim.x) {
    const int c = e / n_comp, q = e - c * n_comp;
    const int col = perm ? perm[(long long)f * ld_perm + q] : q;
    float acc = 0.f;
    for (int j = 0; j < r; ++j)
      acc = fmaf(Vp[c * R + j], Uf[(long long)(off + j) * ldu + col] * rsqrtf(dhf[off + j]), acc);
    Lp[(long long)c * ldl + q] = acc;
  }
}

// ------------------------------------------------------------------------- PCA scores
// St[f][j][i] = V[f][i][perm[j]] * sqrt(max(lambda_j, 0))     j < k[f], i < n[f]
__global__ void __launch_bounds__(256)
k_scores_train(const float* __restrict__ V, int ldv, long long strideV,
               const float* __restrict__ evals, const int* __restrict__ perm, int ld_e,
               const int* __restrict__ k_dev, const int* __restrict__ n_dev, int n_fixed,
               float* __restrict__ St, int lds, long long strideS, int kcap) {
  const int f = blockIdx.z;
  const int k = min(k_dev[f], kcap);
  const int n = n_dev ? n_dev[f] : n_fixed;
  const float* Vf = V + (long long)f * strideV;
  float* Sf = St + (long long)f * strideS;
  for (int j = blockIdx.y; j < k; j += gridDim.y) {
    const float sc = sqrtf(fmaxf(evals[(long long)f * ld_e + j], 0.f));
    const int col = perm ? perm[(long long)f * ld_e + j] : j;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
      Sf[(long long)j * lds + i] = Vf[(long long)i * ldv + col] * sc;
  }
}

// Ste[f][j][t] = ( sum_i Kte[f][t][i] V[f][i][perm[j]] ) / sqrt(lambda_j)
__global__ void __launch_bounds__(128)
k_scores_test(const float* __restrict__ Kte, int ldk, long long strideK,
              const float* __restrict__ V, int ldv, long long strideV,
              const float* __restrict__ evals, const int* __restrict__ perm, int ld_e,
              const int* __restrict__ k_dev, const int* __restrict__ n_dev, int n_fixed,
              int n_te, float* __restrict__ Ste, int ldt, long long strideT, int kcap) {
  const int f = blockIdx.z;
  const int k = min(k_dev[f], kcap);
  const int n = n_dev ? n_dev[f] : n_fixed;
  const int t = blockIdx.y;
  if (t >= n_te) return;
  const float* Vf = V + (long long)f * strideV;
  const float* kr = Kte + (long long)f * strideK + (long long)t * ldk;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < k; j += gridDim.x * blockDim.x) {
    const int col = perm ? perm[(long long)f * ld_e + j] : j;
    float acc = 0.f;
    for (int i = 0; i < n; ++i) acc = fmaf(kr[i], Vf[(long long)i * ldv + col], acc);
    const float lam = evals[(long long)f * ld_e + j];
    Ste[(long long)f * strideT + (long long)j * ldt + t] = (lam > 0.f) ? acc * rsqrtf(lam) : 0.f;
  }
}

// out[p] = base + sign * sum_{j in list[list_ptr[p] .. list_ptr[p+1])} mats[list[j]]   (fp64)
__global__ void __launch_bounds__(256)
k_sum_mats_f64(const double* __restrict__ base, const double* __restrict__ mats, long long mat_stride,
               const int* __restrict__ list_ptr, const int* __restrict__ list, double sign,
               double* __restrict__ out, long long out_stride, int elems) {
  const int p = blockIdx.y;
  const int j0 = list_ptr[p], j1 = list_ptr[p + 1];
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < elems; e += gridDim.x * blockDim.x) {
    double acc = 0.0;
    for (int j = j0; j < j1; ++j) acc += mats[(long long)list[j] * mat_stride + e];
    out[(long long)p * out_stride + e] = (base ? base[e] : 0.0) + sign * acc;
  }
}

// dst[r][j] = src[r][idx[j]]   (electrode subsampling: keep a subset of the channels)
__global__ void __launch_bounds__(256)
k_gather_channels(const float* __restrict__ src, int lds, const int* __restrict__ idx, int nidx,
                  float* __restrict__ dst, int ldd, long long nrows) {
  const long long total = nrows * nidx;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / nidx;
    const int j = (int)(e - r * nidx);
    dst[r * ldd + j] = src[r * lds + idx[j]];
  }
}

// out[trial][t][reg] = mean over the electrodes e of region reg of data[trial][e][t]
// (data: trials x electrodes x time, time contiguous; fp64 like the reference's np.mean)
__global__ void __launch_bounds__(128)
k_region_mean_f64(const double* __restrict__ data, int nelec, int T, const int* __restrict__ reg_ptr,
                  const int* __restrict__ reg_elec, int nreg, double* __restrict__ out) {
  const int trial = blockIdx.y, reg = blockIdx.x;
  const int e0 = reg_ptr[reg], e1 = reg_ptr[reg + 1];
  const double* d = data + (long long)trial * nelec * T;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    double acc = 0.0;
    for (int e = e0; e < e1; ++e) acc += d[(long long)reg_elec[e] * T + t];
    out[((long long)trial * T + t) * nreg + reg] = (e1 > e0) ? acc / (double)(e1 - e0) : 0.0;
  }
}

// r[c] = Pearson correlation of rows a[c][:], b[c][:] (fp64, two-pass like scipy.stats.pearsonr:
// centre, then normalise) -- alignment/metrics.py:41-68 pt_corr, one CTA per condition.
__global__ void __launch_bounds__(256)
k_pearson_rows(const double* __restrict__ a, const double* __restrict__ b, long long len,
               double* __restrict__ r) {
  __shared__ double red[40];
  const double* x = a + (long long)blockIdx.x * len;
  const double* y = b + (long long)blockIdx.x * len;
  double sx = 0.0, sy = 0.0;
  for (long long i = threadIdx.x; i < len; i += blockDim.x) { sx += x[i]; sy += y[i]; }
  const double mx = block_sum(sx, red) / (double)len;
  const double my = block_sum(sy, red) / (double)len;
  double sxx = 0.0, syy = 0.0, sxy = 0.0;
  for (long long i = threadIdx.x; i < len; i += blockDim.x) {
    const double dx = x[i] - mx, dy = y[i] - my;
    sxx = fma(dx, dx, sxx);
    syy = fma(dy, dy, syy);
    sxy = fma(dx, dy, sxy);
  }
  sxx = block_sum(sxx, red);
  syy = block_sum(syy, red);
  sxy = block_sum(sxy, red);
  if (threadIdx.x == 0) {
    double v = sxy / (sqrt(sxx) * sqrt(syy));
    r[blockIdx.x] = fmax(-1.0, fmin(1.0, v));      // scipy clips rounding overshoot the same way
  }
}

// ---- JointPCA (alignment/JointPCA.py:165-211) --------------------------------------------
// G (n x n fp64): Gram of the channel-concatenated class averages, only the blocks u <= v are
// stored (diagonal blocks completely); s (n, fp32): column sums; nrows: rows of the matrix.
__device__ __forceinline__ double gsym(const double* G, int n, int i, int j) {
  return (i <= j) ? G[(long long)i * n + j] : G[(long long)j * n + i];
}

// cov[i][j] = (G_ij - nrows m_i m_j) / (nrows - 1), m = s / nrows: sklearn PCA's covariance of
// the concatenated matrix, written as fp32 (n_pad x n_pad, zero padded)
__global__ void __launch_bounds__(256)
k_joint_cov(const double* __restrict__ G, const float* __restrict__ s, const int* __restrict__ nrows,
            int n, float* __restrict__ cov, int n_pad) {
  const int f = blockIdx.y;
  const double* Gf = G + (long long)f * n * n;
  const float* sf = s + (long long)f * n;
  float* cf = cov + (long long)f * n_pad * n_pad;
  const double nr = (double)nrows[f];
  const long long total = (long long)n_pad * n_pad;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e / n_pad), j = (int)(e - (long long)i * n_pad);
    float v = 0.f;
    if (i < n && j < n)
      v = (float)((gsym(Gf, n, i, j) - (double)sf[i] * (double)sf[j] / nr) / (nr - 1.0));
    cf[e] = v;
  }
}

// rhs[f][view] (C x k, fp64, row stride ldr) = X_p^T (M - 1 mean^T) V_k
//   = sum_c (G[off + r][c] - s[off + r] s[c] / nrows) V[c][j]       grid (C_max, B * P)
__global__ void __launch_bounds__(128)
k_joint_rhs(const double* __restrict__ G, const float* __restrict__ s, const int* __restrict__ nrows,
            int n, const float* __restrict__ V, int ldv, long long strideV, const int* __restrict__ coff,
            int P, int k, double* __restrict__ rhs, int ldr, long long strideR) {
  const int fp = blockIdx.y, f = fp / P, p = fp - f * P;
  const int C = coff[p + 1] - coff[p], r = blockIdx.x;
  if (r >= C) return;
  const int j = threadIdx.x;
  const int gi = coff[p] + r;
  const double* Gf = G + (long long)f * n * n;
  const float* sf = s + (long long)f * n;
  const float* Vf = V + (long long)f * strideV;
  const double si = (double)sf[gi] / (double)nrows[f];
  double acc = 0.0;
  if (j < k)
    for (int c = 0; c < n; ++c)
      acc = fma(gsym(Gf, n, gi, c) - si * (double)sf[c], (double)Vf[(long long)c * ldv + j], acc);
  if (j < k) rhs[(long long)fp * strideR + (long long)r * ldr + j] = acc;
}

}  // namespace

extern "C" int cpsd_joint_cov(const double* G, const float* s, const int* nrows, int n, float* cov,
                              int n_pad, int nfold, cudaStream_t stream) {
  CPSD_CHECK_ARG(n > 0 && n_pad >= n && nfold >= 0, "joint_cov: bad dims");
  if (nfold == 0) return CPSD_OK;
  k_joint_cov<<<dim3(148, nfold), 256, 0, stream>>>(G, s, nrows, n, cov, n_pad);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

extern "C" int cpsd_joint_rhs(const double* G, const float* s, const int* nrows, int n, const float* V,
                              int ldv, long long strideV, const int* coff, int P, int C_max, int k,
                              double* rhs, int ldr, long long strideR, int nfold, cudaStream_t stream) {
  CPSD_CHECK_ARG(n > 0 && P > 0 && k > 0 && k <= 128 && C_max > 0, "joint_rhs: k must be in 1..128");
  if (nfold == 0) return CPSD_OK;
  k_joint_rhs<<<dim3(C_max, nfold * P), 128, 0, stream>>>(G, s, nrows, n, V, ldv, strideV, coff, P, k, rhs,
                                                         ldr, strideR);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

// Electrode subsampling on resident trials (scripts/aligned_decode_{grid,pitch}_subsample.py
// index the channel axis with the lists made by processing_utils/{grid_subsampling,
// poisson_disk_sampling}.py): dst (nrows x nidx) = src[:, idx].
extern "C" int cpsd_gather_channels(const float* src, int lds, const int* idx, int nidx, float* dst,
                                    int ldd, long long nrows, cudaStream_t stream) {
  CPSD_CHECK_ARG(nidx > 0 && nrows >= 0 && lds > 0 && ldd >= nidx, "gather_channels: bad dims");
  if (nrows == 0) return CPSD_OK;
  long long nb = (nrows * nidx + 255) / 256;
  if (nb > 148 * 32) nb = 148 * 32;
  k_gather_channels<<<(int)nb, 256, 0, stream>>>(src, lds, idx, nidx, dst, ldd, nrows);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

// Per-trial column sums in fp64 (sums[t][c] = sum over the T rows of trial t of X[t*T + r][c]) and
// the covariance of a trial subset from per-trial statistics: with G = sum of the subset's
// per-trial Grams X_t^T X_t and s = sum of its column sums over n = (#trials * T) rows,
//     cov = (G - s s^T / n) / (n - 1),   mu = s_mu / n
// (normalize == 0: the centred scatter G - s s^T / n itself, as the MCCA view solves take it)
// (s_mu = s unless G and s were moved to another basis: the means stay in channel coordinates)
// -- sklearn PCA's centred covariance (decoders/cross_pt_decoders.py:234-241 -> PCA.fit) of every
// CV fold's train trials without another pass over the data: the per-trial statistics are
// computed once per target, a fold is a list of trials (cpsd_sum_mats_f64).
namespace {
__global__ void __launch_bounds__(256)
k_trial_colsum_f64(const float* __restrict__ X, int T, int C, int ldx, double* __restrict__ sums, int lds) {
  const int t = blockIdx.x;
  const float* Xt = X + (long long)t * T * ldx;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double a = 0.0;
    for (int r = 0; r < T; ++r) a += (double)Xt[(long long)r * ldx + c];
    sums[(long long)t * lds + c] = a;
  }
}

__global__ void __launch_bounds__(256)
k_cov_from_sums(double* __restrict__ G, int ldg, long long strideG, const double* __restrict__ s, int lds,
                const int* __restrict__ nrows, int C, float* __restrict__ mu, int ldmu,
                const double* __restrict__ s_mu, int normalize) {
  const int p = blockIdx.x;
  const double n = (double)nrows[p];
  double* Gp = G + (long long)p * strideG;
  const double* sp = s + (long long)p * lds;
  for (int e = threadIdx.x; e < C * C; e += blockDim.x) {
    const int i = e / C, j = e - i * C;
    const double v = Gp[(long long)i * ldg + j] - sp[i] * sp[j] / n;
    Gp[(long long)i * ldg + j] = normalize ? v / (n - 1.0) : v;
  }
  if (mu) {
    const double* sm = s_mu ? s_mu + (long long)p * lds : sp;
    for (int c = threadIdx.x; c < C; c += blockDim.x) mu[(long long)p * ldmu + c] = (float)(sm[c] / n);
  }
}
}  // namespace

extern "C" int cpsd_trial_colsum_f64(const float* X, int n_trials, int T, int C, int ldx, double* sums,
                                     int lds, cudaStream_t stream) {
  CPSD_CHECK_ARG(n_trials >= 0 && T > 0 && C > 0 && ldx >= C && lds >= C, "trial_colsum_f64: bad dims");
  if (n_trials == 0) return CPSD_OK;
  k_trial_colsum_f64<<<n_trials, 256, 0, stream>>>(X, T, C, ldx, sums, lds);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

extern "C" int cpsd_cov_from_sums(double* G, int ldg, long long strideG, const double* s, int lds,
                                  const int* nrows_dev, int C, float* mu, int ldmu,
                                  const double* s_mu, int normalize, int nprob, cudaStream_t stream) {
  CPSD_CHECK_ARG(nprob >= 0 && C > 0 && ldg >= C && lds >= C && nrows_dev != nullptr, "cov_from_sums: bad dims");
  if (nprob == 0) return CPSD_OK;
  k_cov_from_sums<<<nprob, 256, 0, stream>>>(G, ldg, strideG, s, lds, nrows_dev, C, mu, ldmu, s_mu, normalize);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

// Zeroes the columns j >= k[p / group] of problem p's (rows x cols) matrix: the read-in matrices
// of a JointPCA whose component count is a variance fraction (JointPCA(n_components=0.9),
// scripts/aligned_decode_svm_ncv.py:186-190) are computed at a fixed width and cut per fold.
namespace {
__global__ void k_mask_cols(float* __restrict__ M, int ld, long long stride, int rows, int cols,
                            const int* __restrict__ k_dev, int group) {
  const int p = blockIdx.x;
  const int k = k_dev[p / group];
  float* Mp = M + (long long)p * stride;
  for (int e = threadIdx.x; e < rows * cols; e += blockDim.x) {
    const int r = e / cols, j = e - r * cols;
    if (j >= k) Mp[(long long)r * ld + j] = 0.f;
  }
}
}  // namespace

extern "C" int cpsd_mask_cols(float* M, int ld, long long stride, int rows, int cols, const int* k_dev,
                              int group, int nprob, cudaStream_t stream) {
  CPSD_CHECK_ARG(nprob >= 0 && rows > 0 && cols > 0 && ld >= cols && group > 0 && k_dev != nullptr,
                 "mask_cols: bad dims");
  if (nprob == 0) return CPSD_OK;
  k_mask_cols<<<nprob, 256, 0, stream>>>(M, ld, stride, rows, cols, k_dev, group);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

// Trial subsampling on resident trials (scripts/aligned_decode_cross_patient_subsample.py:
// 303-312 index the trial axis with np.random.choice): dst[i] = src[idx[i]], rows of TC floats.
namespace {
__global__ void __launch_bounds__(256)
k_gather_trials(const float* __restrict__ src, long long TC, const int* __restrict__ idx,
                float* __restrict__ dst, int vec) {
  const long long row = blockIdx.y;
  const float* s = src + (long long)idx[row] * TC;
  float* d = dst + row * TC;
  if (vec) {
    const float4* s4 = reinterpret_cast<const float4*>(s);
    float4* d4 = reinterpret_cast<float4*>(d);
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < TC / 4;
         e += (long long)gridDim.x * blockDim.x)
      d4[e] = s4[e];
  } else {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < TC;
         e += (long long)gridDim.x * blockDim.x)
      d[e] = s[e];
  }
}
}  // namespace

extern "C" int cpsd_gather_trials(const float* src, long long TC, const int* idx, int n, float* dst,
                                  cudaStream_t stream) {
  CPSD_CHECK_ARG(TC > 0 && n >= 0 && n <= 65535, "gather_trials: bad dims");
  if (n == 0) return CPSD_OK;
  const int vec = (TC % 4 == 0) && ((((uintptr_t)src) & 15) == 0) && ((((uintptr_t)dst) & 15) == 0);
  long long per = vec ? TC / 4 : TC;
  int bx = (int)((per + 255) / 256);
  if (bx > 32) bx = 32;
  k_gather_trials<<<dim3(bx, n), 256, 0, stream>>>(src, TC, idx, dst, vec);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

// spatial_avg_data (processing_utils/spatial_avg_subsampling.py:74-96): block means of the
// electrode grid.  data (trials x nelec x T) fp64, regions as CSR (reg_ptr, reg_elec = flat
// electrode index x * grid_y + y), out (trials x T x nreg) fp64.
extern "C" int cpsd_region_mean_f64(const double* data, int ntrials, int nelec, int T,
                                    const int* reg_ptr, const int* reg_elec, int nreg, double* out,
                                    cudaStream_t stream) {
  CPSD_CHECK_ARG(ntrials >= 0 && nelec > 0 && T > 0 && nreg > 0, "region_mean_f64: bad dims");
  CPSD_CHECK_ARG(ntrials <= 65535, "region_mean_f64: too many trials");
  if (ntrials == 0) return CPSD_OK;
  k_region_mean_f64<<<dim3(nreg, ntrials), 128, 0, stream>>>(data, nelec, T, reg_ptr, reg_elec, nreg,
                                                             out);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

extern "C" int cpsd_pearson_rows(const double* a, const double* b, int nrows, long long len, double* r,
                                 cudaStream_t stream) {
  CPSD_CHECK_ARG(nrows >= 0 && len >= 2, "pearson_rows: need at least 2 samples per row");
  if (nrows == 0) return CPSD_OK;
  k_pearson_rows<<<nrows, 256, 0, stream>>>(a, b, len, r);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

// Scatter matrix of a trial subset from per-trial scatter matrices: the uncentred Gram of the
// target's train trials (AlignMCCA.n_components_var, AlignMCCA.py:156-174) is the all-trials
// Gram minus the Grams of the few held-out trials.
extern "C" int cpsd_sum_mats_f64(const double* base, const double* mats, long long mat_stride,
                                 const int* list_ptr, const int* list, double sign, double* out,
                                 long long out_stride, int elems, int nprob, cudaStream_t stream) {
  CPSD_CHECK_ARG(nprob >= 0 && elems > 0, "sum_mats_f64: bad dims");
  if (nprob == 0) return CPSD_OK;
  CPSD_CHECK_ARG(nprob <= 65535, "sum_mats_f64: nprob > 65535");
  int bx = (elems + 255) / 256;
  if (bx > 16) bx = 16;
  k_sum_mats_f64<<<dim3(bx, nprob), 256, 0, stream>>>(base, mats, mat_stride, list_ptr, list, sign, out,
                                                      out_stride, elems);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

extern "C" int cpsd_class_mean(const cpsd_class_mean_desc* descs_dev, int nprob, int nslot_max,
                               int TC, cudaStream_t stream) {
  CPSD_CHECK_ARG(nprob >= 0 && nslot_max >= 0 && TC > 0, "class_mean: bad dims");
  if (nprob == 0 || nslot_max == 0) return CPSD_OK;
  CPSD_CHECK_ARG(nslot_max <= 65535 && nprob <= 65535, "class_mean: grid too large");
  int bx = (TC / 4 + 255) / 256;
  if (bx > 32) bx = 32;
  if (bx < 1) bx = 1;
  k_class_mean<<<dim3(bx, nslot_max, nprob), 256, 0, stream>>>(descs_dev);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

extern "C" int cpsd_center_rows(float* Z, int ld, long long strideZ, const float* mu, int ldmu,
                                const int* nrows_dev, int nrows_fixed, int nrows_max, int ncols,
                                int nprob, cudaStream_t stream) {
  CPSD_CHECK_ARG(nprob >= 0 && ncols > 0 && nrows_max >= 0, "center_rows: bad dims");
  if (nprob == 0 || nrows_max == 0) return CPSD_OK;
  int bx = (ncols + 255) / 256;
  if (bx > 8) bx = 8;
  int by = nrows_max < 256 ? nrows_max : 256;
  k_center_rows<<<dim3(bx, by, nprob), 256, 0, stream>>>(Z, ld, strideZ, mu, ldmu, nrows_dev,
                                                        nrows_fixed, ncols);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

extern "C" int cpsd_copy_rows(const float* src, int lds, long long strideS, float* dst, int ldd,
                              long long strideD, const int* r0_dev, int r0_fixed, int nrows,
                              int ncols, int src_rows, int nprob, cudaStream_t stream) {
  CPSD_CHECK_ARG(nprob >= 0 && ncols > 0 && nrows >= 0, "copy_rows: bad dims");
  if (nprob == 0 || nrows == 0) return CPSD_OK;
  int bx = (ncols + 255) / 256;
  if (bx > 8) bx = 8;
  int by = nrows < 256 ? nrows : 256;
  k_copy_rows<<<dim3(bx, by, nprob), 256, 0, stream>>>(src, lds, strideS, dst, ldd, strideD, r0_dev,
                                                      r0_fixed, nrows, ncols, src_rows);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

extern "C" int cpsd_cast_f64_f32(const double* src, float* dst, long long n, cudaStream_t stream) {
  CPSD_CHECK_ARG(n >= 0, "cast_f64_f32: bad size");
  if (n == 0) return CPSD_OK;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_cast_f64_f32<<<(int)blocks, 256, 0, stream>>>(src, dst, n);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

extern "C" int cpsd_permute_cols(const float* src, int lds, long long strideS, const int* perm,
                                 int ld_perm, float* dst, int ldd, long long strideD, int nrows,
                                 int ncols, int nprob, cudaStream_t stream) {
  CPSD_CHECK_ARG(nprob >= 0 && ncols > 0 && nrows >= 0, "permute_cols: bad dims");
  if (nprob == 0 || nrows == 0) return CPSD_OK;
  int bx = (ncols + 255) / 256;
  if (bx > 8) bx = 8;
  int by = nrows < 256 ? nrows : 256;
  k_permute_cols<<<dim3(bx, by, nprob), 256, 0, stream>>>(src, lds, strideS, perm, ld_perm, dst, ldd,
                                                         strideD, nrows, ncols);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

extern "C" int cpsd_mcca_mask(const float* evecs, int ldv, long long strideV, const float* evals,
                              int ld_e, const int* rank, const int* cdim, int R, int Cmax,
                              float* Vr, float* d2, int* r_eff, int nprob, cudaStream_t stream) {
  CPSD_CHECK_ARG(nprob >= 0 && R > 0 && Cmax > 0, "mcca_mask: bad dims");
  if (nprob == 0) return CPSD_OK;
  k_mcca_mask<<<nprob, 128, 0, stream>>>(evecs, ldv, strideV, evals, ld_e, rank, cdim, nullptr, R,
                                         Cmax, Vr, d2, r_eff);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

// Same, with the eigen-pairs of problem p read from slot src_idx[p] of evecs / evals: the
// view statistics of the cross patients do not depend on the fold, so many (fold, view)
// problems share one solved slot.
extern "C" int cpsd_mcca_mask_idx(const float* evecs, int ldv, long long strideV, const float* evals,
                                  int ld_e, const int* rank, const int* cdim, const int* src_idx,
                                  int R, int Cmax, float* Vr, float* d2, int* r_eff, int nprob,
                                  cudaStream_t stream) {
  CPSD_CHECK_ARG(nprob >= 0 && R > 0 && Cmax > 0, "mcca_mask_idx: bad dims");
  if (nprob == 0) return CPSD_OK;
  k_mcca_mask<<<nprob, 128, 0, stream>>>(evecs, ldv, strideV, evals, ld_e, rank, cdim, src_idx, R,
                                         Cmax, Vr, d2, r_eff);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

extern "C" int cpsd_mcca_build(const float* G, int ldg, long long strideG, const int* r_eff, int P,
                               int R, float reg, float* M, int ldm, long long strideM, int* n_out,
                               int* cidx, float* dh, int n_comp, int* status, int nfold,
                               cudaStream_t stream) {
  CPSD_CHECK_ARG(nfold >= 0 && P > 0 && R > 0, "mcca_build: bad dims");
  CPSD_CHECK_ARG(ldm >= P * R || ldm >= 1, "mcca_build: bad ldm");
  if (nfold == 0) return CPSD_OK;
  const size_t smem = (size_t)P * R * (sizeof(int) + sizeof(float));
  k_mcca_build<<<nfold, 256, smem, stream>>>(G, ldg, strideG, nullptr, 0, 0, nullptr, r_eff, P, R, reg,
                                             M, ldm, strideM, n_out, cidx, dh, n_comp, status);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

// Same with the cross-scatter in two parts: Gt (per fold, R x P*R: the target's rows
// [G_tt | G_tx]) and Gxx (slot xslot[f], (P-1)R x (P-1)R: the cross patients' block, which does
// not depend on the fold and is computed once per shared class set).
extern "C" int cpsd_mcca_build_split(const float* Gt, int ldg, long long strideG, const float* Gxx,
                                     int ldx, long long strideX, const int* xslot, const int* r_eff,
                                     int P, int R, float reg, float* M, int ldm, long long strideM,
                                     int* n_out, int* cidx, float* dh, int n_comp, int* status,
                                     int nfold, cudaStream_t stream) {
  CPSD_CHECK_ARG(nfold >= 0 && P > 0 && R > 0, "mcca_build_split: bad dims");
  CPSD_CHECK_ARG(Gxx != nullptr && xslot != nullptr, "mcca_build_split: Gxx / xslot is NULL");
  if (nfold == 0) return CPSD_OK;
  const size_t smem = (size_t)P * R * (sizeof(int) + sizeof(float));
  k_mcca_build<<<nfold, 256, smem, stream>>>(Gt, ldg, strideG, Gxx, ldx, strideX, xslot, r_eff, P, R,
                                             reg, M, ldm, strideM, n_out, cidx, dh, n_comp, status);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

extern "C" int cpsd_mcca_loadings(const float* Vr, const float* U, int ldu, long long strideU,
                                  const int* perm, int ld_perm, const int* r_eff, const float* dh,
                                  int P, int R, int Cmax, int n_comp, float* L, int ldl, int nfold,
                                  cudaStream_t stream) {
  CPSD_CHECK_ARG(nfold >= 0 && P > 0 && R > 0 && n_comp > 0, "mcca_loadings: bad dims");
  if (nfold == 0) return CPSD_OK;
  k_mcca_loadings<<<nfold * P, 128, 0, stream>>>(Vr, U, ldu, strideU, perm, ld_perm, r_eff, dh, P, R,
                                                 Cmax, n_comp, L, ldl);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

extern "C" int cpsd_scores_train(const float* V, int ldv, long long strideV, const float* evals,
                                 const int* perm, int ld_e, const int* k_dev, const int* n_dev,
                                 int n_fixed, int n_max, float* St, int lds, long long strideS,
                                 int kcap, int nfold, cudaStream_t stream) {
  CPSD_CHECK_ARG(nfold >= 0 && kcap > 0 && n_max > 0, "scores_train: bad dims");
  if (nfold == 0) return CPSD_OK;
  int by = kcap < 128 ? kcap : 128;
  k_scores_train<<<dim3((n_max + 255) / 256, by, nfold), 256, 0, stream>>>(
      V, ldv, strideV, evals, perm, ld_e, k_dev, n_dev, n_fixed, St, lds, strideS, kcap);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

extern "C" int cpsd_scores_test(const float* Kte, int ldk, long long strideK, const float* V, int ldv,
                                long long strideV, const float* evals, const int* perm, int ld_e,
                                const int* k_dev, const int* n_dev, int n_fixed, int n_te,
                                float* Ste, int ldt, long long strideT, int kcap, int nfold,
                                cudaStream_t stream) {
  CPSD_CHECK_ARG(nfold >= 0 && kcap > 0 && n_te >= 0, "scores_test: bad dims");
  if (nfold == 0 || n_te == 0) return CPSD_OK;
  int bx = (kcap + 127) / 128;
  if (bx > 16) bx = 16;
  k_scores_test<<<dim3(bx, n_te, nfold), 128, 0, stream>>>(Kte, ldk, strideK, V, ldv, strideV, evals,
                                                           perm, ld_e, k_dev, n_dev, n_fixed, n_te,
                                                           Ste, ldt, strideT, kcap);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}
