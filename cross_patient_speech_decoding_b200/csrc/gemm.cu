// Batched covariance / projection GEMMs (fp32 SIMT tiles, fp32 accumulate).
//
//  k_gram_tn  : out = alpha * sum_rows (a_r - muA)^T (b_r - muB)   tall-skinny "TN" Gram.
//               Per-patient channel covariances for PCA (reference: sklearn PCA reached
//               from decoders/cross_pt_decoders.py:234-241), the uncentered spectrum of
//               AlignMCCA.n_components_var (AlignMCCA.py:156-174), the condition-average
//               scatter / cross-scatter blocks of CCA (AlignCCA.py:235-285, Gram form) and
//               MCCA (AlignMCCA.py:140-154).
//  k_proj_nn  : Y = (X - mu) W   projection of trials x time rows onto latent directions
//               (PCA.transform, AlignCCA.transform AlignCCA.py:93, MCCA transform_view
//               AlignMCCA.py:110, JointPCA.transform JointPCA.py:132), written straight into
//               the pooled trials x (time*latent) layout.
//  k_gram_nt  : out = alpha * A B^T over the long feature axis (pooled Gram of the
//               decoder-stage PCA, DimRedReshape.py:47-49).
//  k_colsum   : column means over segment rows.
#include "common.cuh"
#include "descs.h"

namespace {

// ------------------------------------------------------------------------------ gram_tn
#define GT_TILE 64
#define GT_RK 16
// nsplit > 1: the segments are dealt round-robin to nsplit CTAs per output tile (blockIdx.x =
// tile + tiles_x * split); every CTA writes its partial tile to part[split][prob] and
// k_gram_tn_reduce adds the partials in a fixed order (deterministic, no atomics) -- for
// batches with few problems and many rows (the per-fold scatters: 3 tiles x 10 400 rows).
template <typename AccT>
__global__ void __launch_bounds__(256)
k_gram_tn(const cpsd_gram_tn_desc* __restrict__ descs, int tiles_x, int nsplit, AccT* __restrict__ part,
          long long part_stride) {
  const cpsd_gram_tn_desc d = descs[blockIdx.z];
  const int bx = blockIdx.x % tiles_x, split = blockIdx.x / tiles_x;
  const int i0 = blockIdx.y * GT_TILE, j0 = bx * GT_TILE;
  if (i0 >= d.p || j0 >= d.q) return;
  if ((d.sym & 1) && blockIdx.y > bx) return;
  __shared__ __align__(16) float As[GT_RK][GT_TILE];
  __shared__ __align__(16) float Bs[GT_RK][GT_TILE];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  AccT acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = (AccT)0;

  // this thread's load slots: 4 elements of each operand per chunk
  const int lc = threadIdx.x & 63;   // column inside tile
  const int lr = threadIdx.x >> 6;   // 0..3, rows lr, lr+4, lr+8, lr+12
  const int ca = i0 + lc, cb = j0 + lc;
  const bool va = ca < d.p, vb = cb < d.q;
  const float ma = (va && d.muA) ? d.muA[ca] : 0.f;
  const float mb = (vb && d.muB) ? d.muB[cb] : 0.f;

  for (int s = split; s < d.nseg; s += nsplit) {
    const long long ra = d.segA[s], rb = d.segB ? d.segB[s] : ra;
    for (int r0 = 0; r0 < d.seg_len; r0 += GT_RK) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = lr + 4 * u;
        const bool vr = (r0 + r) < d.seg_len;
        float a = 0.f, b = 0.f;
        if (vr && va) a = d.A[(ra + r0 + r) * d.lda + ca] - ma;
        if (vr && vb) b = d.B[(rb + r0 + r) * d.ldb + cb] - mb;
        As[r][lc] = a;
        Bs[r][lc] = b;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < GT_RK; ++kk) {
        const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        const AccT a[4] = {(AccT)a4.x, (AccT)a4.y, (AccT)a4.z, (AccT)a4.w};
        const AccT b[4] = {(AccT)b4.x, (AccT)b4.y, (AccT)b4.z, (AccT)b4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
  AccT* out = reinterpret_cast<AccT*>(d.out);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gi = i0 + ty * 4 + i;
    if (gi >= d.p) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gj = j0 + tx * 4 + j;
      if (gj >= d.q) continue;
      const AccT v = (AccT)d.alpha * acc[i][j];
      if (nsplit > 1) {
        AccT* po = part + ((long long)split * gridDim.z + blockIdx.z) * part_stride;
        po[(long long)gi * d.ldo + gj] = v;
        if ((d.sym & 1) && blockIdx.y != bx) po[(long long)gj * d.ldo + gi] = v;
      } else {
        out[(long long)gi * d.ldo + gj] = v;
        if ((d.sym & 1) && blockIdx.y != bx) out[(long long)gj * d.ldo + gi] = v;
      }
    }
  }
}

// out[prob] = sum over splits (ascending) of part[split][prob]
__global__ void __launch_bounds__(256)
k_gram_tn_reduce(const cpsd_gram_tn_desc* __restrict__ descs, const double* __restrict__ part,
                 long long part_stride, int nsplit) {
  const cpsd_gram_tn_desc d = descs[blockIdx.y];
  double* out = reinterpret_cast<double*>(d.out);
  const int total = d.p * d.q;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int i = e / d.q, j = e - i * d.q;
    double acc = 0.0;
    for (int s = 0; s < nsplit; ++s)
      acc += part[((long long)s * gridDim.y + blockIdx.y) * part_stride + (long long)i * d.ldo + j];
    out[(long long)i * d.ldo + j] = acc;
  }
}

// ------------------------------------------------------------------------------- colsum
// grid (ceil(p/32), nprob), 256 threads: 32 columns x 8 row phases
__global__ void __launch_bounds__(256)
k_colsum(const cpsd_colsum_desc* __restrict__ descs) {
  const cpsd_colsum_desc d = descs[blockIdx.y];
  const int lc = threadIdx.x & 31;
  const int c = blockIdx.x * 32 + lc;
  const int ph = threadIdx.x >> 5;
  __shared__ double part[8][32];
  double acc = 0.0;
  if (c < d.p) {
    for (int s = 0; s < d.nseg; ++s) {
      const long long r0 = d.segA[s];
      float loc = 0.f;
      for (int r = ph; r < d.seg_len; r += 8) loc += d.A[(r0 + r) * d.lda + c];
      acc += (double)loc;
    }
  }
  part[ph][lc] = acc;
  __syncthreads();
  if (ph == 0 && c < d.p) {
    double t = 0.0;
#pragma unroll
    for (int u = 0; u < 8; ++u) t += part[u][lc];
    d.out[c] = (float)(t * (double)d.alpha);
  }
}

// ------------------------------------------------------------------------------ proj_nn
#define PJ_ROWS 64
#define PJ_CK 32
template <int QT>
__global__ void __launch_bounds__(256)
k_proj_nn(const cpsd_proj_desc* __restrict__ descs, int blocks_per_seg) {
  const cpsd_proj_desc d = descs[blockIdx.y];
  const int seg = blockIdx.x / blocks_per_seg, blk = blockIdx.x - seg * blocks_per_seg;
  if (seg >= d.nseg) return;
  const int t0 = blk * PJ_ROWS;
  if (t0 >= d.seg_len) return;
  __shared__ float Xs[PJ_ROWS][PJ_CK + 1];
  __shared__ float Ws[PJ_CK][16 * QT];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const long long src0 = (long long)d.seg_src[seg] + t0;
  const long long dst0 = (long long)d.seg_dst[seg] + t0;
  const int nrow = min(PJ_ROWS, d.seg_len - t0);
  float acc[4][QT];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int u = 0; u < QT; ++u) acc[i][u] = 0.f;

  for (int c0 = 0; c0 < d.C; c0 += PJ_CK) {
    // X tile: 64 rows x 32 channels
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int e = threadIdx.x + 256 * u;
      const int r = e >> 5, c = e & 31;
      float v = 0.f;
      if (r < nrow && c0 + c < d.C) {
        v = d.X[(src0 + r) * d.ldx + c0 + c];
        if (d.mu) v -= d.mu[c0 + c];
      }
      Xs[r][c] = v;
    }
    for (int e = threadIdx.x; e < PJ_CK * 16 * QT; e += 256) {
      const int c = e / (16 * QT), j = e - c * (16 * QT);
      float v = 0.f;
      if (c0 + c < d.C && j < d.q) v = d.W[(long long)(c0 + c) * d.ldw + j];
      Ws[c][j] = v;
    }
    __syncthreads();
#pragma unroll 8
    for (int cc = 0; cc < PJ_CK; ++cc) {
      float a[4], b[QT];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = Xs[ty * 4 + i][cc];
#pragma unroll
      for (int u = 0; u < QT; ++u) b[u] = Ws[cc][tx + 16 * u];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int u = 0; u < QT; ++u) acc[i][u] = fmaf(a[i], b[u], acc[i][u]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = ty * 4 + i;
    if (r >= nrow) continue;
#pragma unroll
    for (int u = 0; u < QT; ++u) {
      const int j = tx + 16 * u;
      if (j < d.q) d.Y[(dst0 + r) * d.ldy + j] = acc[i][u];
    }
  }
}

// ------------------------------------------------------------------------------ gram_nt
#define NT_TILE 128
#define NT_BK 16
__global__ void __launch_bounds__(256)
k_gram_nt(const cpsd_gram_nt_desc* __restrict__ descs) {
  const cpsd_gram_nt_desc d = descs[blockIdx.z];
  const int i0 = blockIdx.y * NT_TILE, j0 = blockIdx.x * NT_TILE;
  if (i0 >= d.m || j0 >= d.n) return;
  if (d.sym && blockIdx.y > blockIdx.x) return;
  __shared__ float As[NT_TILE][NT_BK + 1];
  __shared__ float Bs[NT_TILE][NT_BK + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  const bool vec_ok = ((d.lda & 3) == 0) && ((d.ldb & 3) == 0) && ((d.k & 3) == 0) &&
                      ((((uintptr_t)d.A) & 15) == 0) && ((((uintptr_t)d.B) & 15) == 0);

  for (int k0 = 0; k0 < d.k; k0 += NT_BK) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int e = threadIdx.x + 256 * u;
      const int r = e >> 2, kq = (e & 3) * 4;
      float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
      if (i0 + r < d.m) {
        const float* src = d.A + (long long)(i0 + r) * d.lda + k0 + kq;
        if (vec_ok && k0 + kq + 3 < d.k) {
          const float4 v = *reinterpret_cast<const float4*>(src);
          a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (k0 + kq + c < d.k) a[c] = src[c];
        }
      }
      if (j0 + r < d.n) {
        const float* src = d.B + (long long)(j0 + r) * d.ldb + k0 + kq;
        if (vec_ok && k0 + kq + 3 < d.k) {
          const float4 v = *reinterpret_cast<const float4*>(src);
          b[0] = v.x; b[1] = v.y; b[2] = v.z; b[3] = v.w;
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (k0 + kq + c < d.k) b[c] = src[c];
        }
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        As[r][kq + c] = a[c];
        Bs[r][kq + c] = b[c];
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < NT_BK; ++kk) {
      float a[8], b[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = As[ty + 16 * i][kk];
#pragma unroll
      for (int j = 0; j < 8; ++j) b[j] = Bs[tx + 16 * j][kk];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gi = i0 + ty + 16 * i;
    if (gi >= d.m) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int gj = j0 + tx + 16 * j;
      if (gj >= d.n) continue;
      const float v = d.alpha * acc[i][j];
      d.out[(long long)gi * d.ldo + gj] = v;
      if (d.sym && blockIdx.y != blockIdx.x && gj < d.m && gi < d.n)
        d.out[(long long)gj * d.ldo + gi] = v;
    }
  }
}

}  // namespace

extern "C" int cpsd_gram_tn(const cpsd_gram_tn_desc* descs_dev, int nprob, int p_max, int q_max,
                            cudaStream_t stream) {
  CPSD_CHECK_ARG(nprob >= 0 && p_max > 0 && q_max > 0, "gram_tn: bad dims");
  if (nprob == 0) return CPSD_OK;
  CPSD_CHECK_ARG(nprob <= 65535, "gram_tn: nprob > 65535");
  dim3 grid((q_max + GT_TILE - 1) / GT_TILE, (p_max + GT_TILE - 1) / GT_TILE, nprob);
  k_gram_tn<float><<<grid, 256, 0, stream>>>(descs_dev, grid.x, 1, nullptr, 0);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

// Same contract with fp64 accumulation; every record's `out` points to DOUBLES (row stride
// ldo in doubles).  Used for the scatter matrices that feed variance thresholds and the
// eigen-solvers (noise-level PCA directions are gap-sensitive).
extern "C" int cpsd_gram_tn_f64(const cpsd_gram_tn_desc* descs_dev, int nprob, int p_max,
                                int q_max, cudaStream_t stream) {
  CPSD_CHECK_ARG(nprob >= 0 && p_max > 0 && q_max > 0, "gram_tn_f64: bad dims");
  if (nprob == 0) return CPSD_OK;
  CPSD_CHECK_ARG(nprob <= 65535, "gram_tn_f64: nprob > 65535");
  dim3 grid((q_max + GT_TILE - 1) / GT_TILE, (p_max + GT_TILE - 1) / GT_TILE, nprob);
  k_gram_tn<double><<<grid, 256, 0, stream>>>(descs_dev, grid.x, 1, nullptr, 0);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

// fp64 accumulation with the segments of every problem split over nsplit CTAs per tile.  The
// partial tiles go to part_ws (nsplit * nprob * ldo_max * p_max doubles, see
// cpsd_gram_tn_split_ws_elems) and are added in a fixed order: bit-reproducible.
extern "C" long long cpsd_gram_tn_split_ws_elems(int nprob, int p_max, int ldo_max, int nsplit) {
  return (long long)nsplit * nprob * ldo_max * p_max;
}

extern "C" int cpsd_gram_tn_f64_split(const cpsd_gram_tn_desc* descs_dev, int nprob, int p_max,
                                      int q_max, int ldo_max, int nsplit, double* part_ws,
                                      cudaStream_t stream) {
  CPSD_CHECK_ARG(nprob >= 0 && p_max > 0 && q_max > 0 && ldo_max >= q_max, "gram_tn_f64_split: bad dims");
  CPSD_CHECK_ARG(nsplit >= 1 && nsplit <= 64, "gram_tn_f64_split: nsplit must be in 1..64");
  CPSD_CHECK_ARG(part_ws != nullptr || nsplit == 1, "gram_tn_f64_split: part_ws is NULL");
  if (nprob == 0) return CPSD_OK;
  CPSD_CHECK_ARG(nprob <= 65535, "gram_tn_f64_split: nprob > 65535");
  const int tx = (q_max + GT_TILE - 1) / GT_TILE;
  dim3 grid(tx * nsplit, (p_max + GT_TILE - 1) / GT_TILE, nprob);
  const long long pstride = (long long)ldo_max * p_max;
  k_gram_tn<double><<<grid, 256, 0, stream>>>(descs_dev, tx, nsplit, part_ws, pstride);
  CPSD_LAUNCH_CHECK();
  if (nsplit > 1) {
    int bx = (p_max * q_max + 255) / 256;
    if (bx > 32) bx = 32;
    k_gram_tn_reduce<<<dim3(bx, nprob), 256, 0, stream>>>(descs_dev, part_ws, pstride, nsplit);
    CPSD_LAUNCH_CHECK();
  }
  return CPSD_OK;
}

extern "C" int cpsd_colsum(const cpsd_colsum_desc* descs_dev, int nprob, int p_max,
                           cudaStream_t stream) {
  CPSD_CHECK_ARG(nprob >= 0 && p_max > 0, "colsum: bad dims");
  if (nprob == 0) return CPSD_OK;
  CPSD_CHECK_ARG(nprob <= 65535, "colsum: nprob > 65535");
  dim3 grid((p_max + 31) / 32, nprob);
  k_colsum<<<grid, 256, 0, stream>>>(descs_dev);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

extern "C" int cpsd_proj_nn(const cpsd_proj_desc* descs_dev, int nprob, int nseg_max, int seg_len,
                            int q_max, cudaStream_t stream) {
  CPSD_CHECK_ARG(nprob >= 0 && nseg_max >= 0 && seg_len > 0, "proj_nn: bad dims");
  CPSD_CHECK_ARG(q_max > 0 && q_max <= 256, "proj_nn: q must be in 1..256");
  if (nprob == 0 || nseg_max == 0) return CPSD_OK;
  CPSD_CHECK_ARG(nprob <= 65535, "proj_nn: nprob > 65535");
  const int bps = (seg_len + PJ_ROWS - 1) / PJ_ROWS;
  dim3 grid(nseg_max * bps, nprob);
  if (q_max <= 32)
    k_proj_nn<2><<<grid, 256, 0, stream>>>(descs_dev, bps);
  else if (q_max <= 64)
    k_proj_nn<4><<<grid, 256, 0, stream>>>(descs_dev, bps);
  else if (q_max <= 128)
    k_proj_nn<8><<<grid, 256, 0, stream>>>(descs_dev, bps);
  else
    k_proj_nn<16><<<grid, 256, 0, stream>>>(descs_dev, bps);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

extern "C" int cpsd_gram_nt(const cpsd_gram_nt_desc* descs_dev, int nprob, int m_max, int n_max,
                            cudaStream_t stream) {
  CPSD_CHECK_ARG(nprob >= 0 && m_max > 0 && n_max > 0, "gram_nt: bad dims");
  if (nprob == 0) return CPSD_OK;
  CPSD_CHECK_ARG(nprob <= 65535, "gram_nt: nprob > 65535");
  dim3 grid((n_max + NT_TILE - 1) / NT_TILE, (m_max + NT_TILE - 1) / NT_TILE, nprob);
  k_gram_nt<<<grid, 256, 0, stream>>>(descs_dev);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}
