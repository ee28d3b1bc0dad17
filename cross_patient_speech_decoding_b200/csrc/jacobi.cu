// Symmetric eigen-solvers for the alignment / PCA stages.
//
//  * k_eig_tile   : one CTA per problem, n <= 128, matrix + eigenvectors resident in
//                   shared memory, parallel two-sided Jacobi (modulus ordering, fused
//                   2x2-block updates, 2 barriers per step).  Replaces the LAPACK calls
//                   behind the reference's per-patient PCA (sklearn covariance_eigh,
//                   reached from decoders/cross_pt_decoders.py:234-241), the singular
//                   values of AlignMCCA.n_components_var (alignment/AlignMCCA.py:156-174)
//                   and the MCCA generalised eigenproblem (AlignMCCA.py:152-153).
//  * k_bj_*       : block two-sided Jacobi for n > 128 (the decoder-stage PCA of the
//                   pooled trials x (time*latent) matrix, decomposition/DimRedReshape.py:
//                   47-49, done here on the n_pool x n_pool Gram).  64-wide blocks are
//                   paired round-robin; each pair's 128x128 diagonal tile is rotated in
//                   shared memory by the same tile routine, and the accumulated rotation
//                   is applied to the rest of the matrix and to the eigenvector matrix by
//                   tile GEMMs.
#include "common.cuh"

#define TS 128          // tile size
#define BS 64           // block size of the block-Jacobi
#define NT 256          // threads per CTA

namespace {

struct StepBuf {
  int pi[TS / 2];
  int pj[TS / 2];
  double c[TS / 2];
  double s[TS / 2];
  double d[TS / 2];   // t * a_pq (diagonal update)
};

// pair (i, j) of slot k at step s, modulus ordering over m (even) indices:
// all pairs with i + j == s (mod m).  Odd s has m/2 pairs, even s has m/2 - 1.
__device__ __forceinline__ void mod_pair(int s, int k, int m, int& i, int& j) {
  const int h = (s + 1) >> 1;
  const int e = (s & 1) ? 0 : 1;
  i = (h + e + k) % m;
  j = h - 1 - k;
  j %= m;
  if (j < 0) j += m;
}

// Rotation parameters of every pair of the step.  They are always evaluated in fp64 and
// rounded once: fp32 evaluation of c = 1/sqrt(1+t^2), s = t c has a systematic bias in
// c^2 + s^2 (+5e-8 per rotation, measured) that accumulates over ~1e4 rotations per column
// into 1e-4 norm drift of the eigenvectors.
//   As : m x m symmetric (row stride TS), npairs pairs stored in sb by the caller; only the
//   first nrot slots rotate (the rest carry idle indices with an identity).
// Returns (per calling thread) the largest |a_pq| it looked at.
template <typename T>
__device__ __forceinline__ float tile_rotations(T* As, StepBuf* sb, int npairs, int nrot,
                                                float skip_thr) {
  float seen = 0.f;
  const int k = threadIdx.x;
  if (k < npairs) {
    const int i = sb->pi[k], j = sb->pj[k];
    const double app = (double)As[i * TS + i], aqq = (double)As[j * TS + j];
    const double apq = (double)As[i * TS + j];
    double c = 1.0, s = 0.0, d = 0.0;
    const float aa = (k < nrot) ? fabsf((float)apq) : 0.f;
    seen = aa;
    if (aa > skip_thr) {
      const double tau = (aqq - app) / (2.0 * apq);
      const double t = copysign(1.0, tau) / (fabs(tau) + sqrt(1.0 + tau * tau));
      if (t == t && fabs(t) <= 1.0) {  // guards inf/nan
        c = 1.0 / sqrt(1.0 + t * t);
        s = t * c;
        d = t * apq;
      }
    }
    sb->c[k] = c;
    sb->s[k] = s;
    sb->d[k] = d;
  }
  return seen;
}

template <typename T>
__device__ __forceinline__ void tile_apply(T* As, float* Vs, StepBuf* sb, int npairs, int vrows) {
  const int nblk = npairs * npairs;
  for (int b = threadIdx.x; b < nblk; b += NT) {
    const int k = b / npairs, l = b - k * npairs;
    const int ik = sb->pi[k], jk = sb->pj[k];
    const T ck = (T)sb->c[k], sk = (T)sb->s[k];
    if (k == l) {
      if (sk != (T)0) {
        const T d = (T)sb->d[k];
        As[ik * TS + ik] -= d;
        As[jk * TS + jk] += d;
        As[ik * TS + jk] = (T)0;
        As[jk * TS + ik] = (T)0;
      }
      continue;
    }
    const int il = sb->pi[l], jl = sb->pj[l];
    const T cl = (T)sb->c[l], sl = (T)sb->s[l];
    if (sk == (T)0 && sl == (T)0) continue;
    const T app = As[ik * TS + il], apq = As[ik * TS + jl];
    const T aqp = As[jk * TS + il], aqq = As[jk * TS + jl];
    // columns (il, jl) rotated by (cl, sl)
    const T tpp = cl * app - sl * apq, tpq = sl * app + cl * apq;
    const T tqp = cl * aqp - sl * aqq, tqq = sl * aqp + cl * aqq;
    // rows (ik, jk) rotated by (ck, sk)
    As[ik * TS + il] = ck * tpp - sk * tqp;
    As[jk * TS + il] = sk * tpp + ck * tqp;
    As[ik * TS + jl] = ck * tpq - sk * tqq;
    As[jk * TS + jl] = sk * tpq + ck * tqq;
  }
  const int nv = vrows * npairs;
  for (int b = threadIdx.x; b < nv; b += NT) {
    const int r = b / npairs, l = b - r * npairs;
    const float cl = (float)sb->c[l], sl = (float)sb->s[l];
    if (sl == 0.f) continue;
    const int il = sb->pi[l], jl = sb->pj[l];
    const float vp = Vs[r * TS + il], vq = Vs[r * TS + jl];
    Vs[r * TS + il] = cl * vp - sl * vq;
    Vs[r * TS + jl] = sl * vp + cl * vq;
  }
}

// Full sweep (all m(m-1)/2 pairs, modulus ordering).  Returns block-wide max |a_pq| seen
// (valid in all threads).  `redmax` is one shared int.
template <typename T>
__device__ float tile_sweep_full(T* As, float* Vs, StepBuf* sb, int m, int vrows, float skip_thr,
                                 int* redmax) {
  float seen = 0.f;
  const int npairs = m >> 1;
  for (int s = 0; s < m; ++s) {
    // even steps leave two indices (s/2 and s/2 + m/2) unpaired: they ride along in the last
    // slot with an identity rotation so that their rows / columns still get updated
    const int nrot = (s & 1) ? npairs : npairs - 1;
    if ((int)threadIdx.x < npairs) {
      int i, j;
      if ((int)threadIdx.x < nrot) {
        mod_pair(s, threadIdx.x, m, i, j);
      } else {
        i = s >> 1;
        j = (i + npairs) % m;
      }
      sb->pi[threadIdx.x] = i;
      sb->pj[threadIdx.x] = j;
    }
    __syncthreads();
    seen = fmaxf(seen, tile_rotations<T>(As, sb, npairs, nrot, skip_thr));
    __syncthreads();
    tile_apply<T>(As, Vs, sb, npairs, vrows);
    __syncthreads();
  }
  if (threadIdx.x == 0) *redmax = 0;
  __syncthreads();
  seen = warp_max(seen);
  if ((threadIdx.x & 31) == 0) atomicMax(redmax, __float_as_int(seen));
  __syncthreads();
  return __int_as_float(*redmax);
}

// Cross sweep for a 128-tile made of two 64-blocks: pairs (i, 64 + (i+s)%64) only.
__device__ float tile_sweep_cross(float* As, float* Vs, StepBuf* sb, float skip_thr,
                                  int* redmax) {
  float seen = 0.f;
  for (int s = 0; s < BS; ++s) {
    if (threadIdx.x < BS) {
      sb->pi[threadIdx.x] = threadIdx.x;
      sb->pj[threadIdx.x] = BS + ((threadIdx.x + s) & (BS - 1));
    }
    __syncthreads();
    seen = fmaxf(seen, tile_rotations<float>(As, sb, BS, BS, skip_thr));
    __syncthreads();
    tile_apply<float>(As, Vs, sb, BS, TS);
    __syncthreads();
  }
  if (threadIdx.x == 0) *redmax = 0;
  __syncthreads();
  seen = warp_max(seen);
  if ((threadIdx.x & 31) == 0) atomicMax(redmax, __float_as_int(seen));
  __syncthreads();
  return __int_as_float(*redmax);
}

template <typename T>
__device__ float tile_diag_absmax(const T* As, int m, int* redmax) {
  if (threadIdx.x == 0) *redmax = 0;
  __syncthreads();
  float v = 0.f;
  for (int i = threadIdx.x; i < m; i += NT) v = fmaxf(v, fabsf((float)As[i * TS + i]));
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) atomicMax(redmax, __float_as_int(v));
  __syncthreads();
  const float r = __int_as_float(*redmax);
  __syncthreads();
  return r;
}

// ---------------------------------------------------------------------------------------
// n <= 128 eigen-decomposition, one CTA per problem.  T = float: everything fp32.
// T = double: the matrix iterates in fp64 (so rotation angles are accurate independently of
// eigenvalue gaps) while the eigenvectors accumulate in fp32 -- their rounding noise is a
// gap-independent ~1e-6 perturbation of the basis.
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(NT)
k_eig_tile(const T* __restrict__ A, int lda, long long strideA, const int* __restrict__ n_dev,
           int n_fixed, float* __restrict__ evals, int ld_e, float* __restrict__ evecs, int ldv,
           long long strideV, int max_sweeps, float tol, int* __restrict__ sweeps_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* As = reinterpret_cast<T*>(smem_raw);
  float* Vs = reinterpret_cast<float*>(As + TS * TS);
  StepBuf* sb = reinterpret_cast<StepBuf*>(Vs + TS * TS);
  int* redmax = reinterpret_cast<int*>(sb + 1);
  int* rank = redmax + 1;  // TS ints

  const int prob = blockIdx.x;
  int n = n_dev ? n_dev[prob] : n_fixed;
  if (n > TS) n = TS;
  if (n < 0) n = 0;
  const int m = (n + 1) & ~1;
  const T* Ag = A + (long long)prob * strideA;

  for (int e = threadIdx.x; e < TS * TS; e += NT) {
    const int r = e >> 7, c = e & (TS - 1);
    T v = (T)0;
    if (r < n && c < n) v = Ag[(long long)r * lda + c];
    As[e] = v;
    Vs[e] = (r == c) ? 1.f : 0.f;
  }
  __syncthreads();
  // symmetrise (inputs are Grams computed tile-wise; removes last-ulp asymmetry)
  for (int e = threadIdx.x; e < TS * TS; e += NT) {
    const int r = e >> 7, c = e & (TS - 1);
    if (r < c) {
      const T v = (T)0.5 * (As[r * TS + c] + As[c * TS + r]);
      As[r * TS + c] = v;
      As[c * TS + r] = v;
    }
  }
  __syncthreads();

  int sw = 0;
  if (m >= 2) {
    for (; sw < max_sweeps; ++sw) {
      const float dmax = tile_diag_absmax<T>(As, m, redmax);
      const float skip = (sizeof(T) == 8 ? 1e-15f : 1e-9f) * dmax + 1e-37f;
      const float off = tile_sweep_full<T>(As, Vs, sb, m, m, skip, redmax);
      if (off <= tol * dmax) { ++sw; break; }
    }
  }
  if (sweeps_out && threadIdx.x == 0) sweeps_out[prob] = sw;

  // sort descending by counting
  for (int i = threadIdx.x; i < n; i += NT) {
    const T li = As[i * TS + i];
    int r = 0;
    for (int j = 0; j < n; ++j) {
      const T lj = As[j * TS + j];
      r += (lj > li) || (lj == li && j < i);
    }
    rank[i] = r;
  }
  __syncthreads();
  float* ev = evals + (long long)prob * ld_e;
  for (int i = threadIdx.x; i < ld_e; i += NT) {
    if (i >= n) ev[i] = 0.f;
  }
  for (int i = threadIdx.x; i < n; i += NT) ev[rank[i]] = (float)As[i * TS + i];
  if (evecs) {
    float* Vg = evecs + (long long)prob * strideV;
    for (int e = threadIdx.x; e < n * n; e += NT) {
      const int r = e / n, c = e - r * n;
      Vg[(long long)r * ldv + rank[c]] = Vs[r * TS + c];
    }
  }
}

// ---------------------------------------------------------------------------------------
// Block Jacobi: inner rotation of one pair's diagonal tile.
// grid (npairs, nprob).  pairs: [npairs][2] block indices of this round.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int tile_gidx(int r, int I, int J) {
  return (r < BS) ? (I * BS + r) : (J * BS + r - BS);
}

__global__ void __launch_bounds__(NT)
k_bj_inner(float* __restrict__ K, int ldk, long long strideK, const int* __restrict__ pairs,
           float* __restrict__ Rbuf, const float* __restrict__ scale, float* __restrict__ conv,
           const int* __restrict__ done, int full_mode) {
  extern __shared__ float smem[];
  float* As = smem;
  float* Vs = smem + TS * TS;
  StepBuf* sb = reinterpret_cast<StepBuf*>(Vs + TS * TS);
  int* redmax = reinterpret_cast<int*>(sb + 1);

  const int pr = blockIdx.x, prob = blockIdx.y, npairs = gridDim.x;
  if (done && done[prob]) return;
  const int I = pairs[2 * pr], J = pairs[2 * pr + 1];
  float* Kg = K + (long long)prob * strideK;
  for (int e = threadIdx.x; e < TS * TS; e += NT) {
    const int r = e >> 7, c = e & (TS - 1);
    As[e] = Kg[(long long)tile_gidx(r, I, J) * ldk + tile_gidx(c, I, J)];
    Vs[e] = (r == c) ? 1.f : 0.f;
  }
  __syncthreads();
  const float sc = scale[prob];
  const float skip = 1e-9f * sc + 1e-37f;
  float off;
  if (full_mode)
    off = tile_sweep_full<float>(As, Vs, sb, TS, TS, skip, redmax);
  else
    off = tile_sweep_cross(As, Vs, sb, skip, redmax);
  if (threadIdx.x == 0 && sc > 0.f)
    atomicMax(reinterpret_cast<int*>(conv + prob), __float_as_int(off / sc));
  for (int e = threadIdx.x; e < TS * TS; e += NT) {
    const int r = e >> 7, c = e & (TS - 1);
    Kg[(long long)tile_gidx(r, I, J) * ldk + tile_gidx(c, I, J)] = As[e];
  }
  float* Rg = Rbuf + ((long long)prob * npairs + pr) * (TS * TS);
  for (int e = threadIdx.x; e < TS * TS; e += NT) Rg[e] = Vs[e];
}

// ---------------------------------------------------------------------------------------
// Block Jacobi: apply the round's rotations to the off-diagonal tiles of K (p < q, both
// mirror images written) and to the column panels of the eigenvector matrix V.
// grid (n_ktasks + n_vtasks, nprob)
// ---------------------------------------------------------------------------------------
#define LDX 129
__global__ void __launch_bounds__(NT)
k_bj_update(float* __restrict__ K, float* __restrict__ V, int ld, long long stride, int n_pad,
            const int* __restrict__ pairs, int npairs, const float* __restrict__ Rbuf,
            const int* __restrict__ done) {
  extern __shared__ float smem[];
  float* Xs = smem;                 // [128][129]
  float* Rs = smem + TS * LDX;      // [128][128]
  const int prob = blockIdx.y;
  if (done && done[prob]) return;
  const int n_ktasks = npairs * (npairs - 1) / 2;
  int task = blockIdx.x;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float* Kg = K + (long long)prob * stride;
  float* Vg = V + (long long)prob * stride;
  const float* Rp_all = Rbuf + (long long)prob * npairs * (TS * TS);

  if (task < n_ktasks) {
    // decode (p, q), p < q
    int p = 0, rem = task;
    while (rem >= npairs - 1 - p) { rem -= npairs - 1 - p; ++p; }
    const int q = p + 1 + rem;
    const int Ip = pairs[2 * p], Jp = pairs[2 * p + 1];
    const int Iq = pairs[2 * q], Jq = pairs[2 * q + 1];
    const float* Rq = Rp_all + (long long)q * (TS * TS);
    const float* Rp = Rp_all + (long long)p * (TS * TS);
    for (int e = threadIdx.x; e < TS * TS; e += NT) {
      const int r = e >> 7, c = e & (TS - 1);
      Xs[r * LDX + c] = Kg[(long long)tile_gidx(r, Ip, Jp) * ld + tile_gidx(c, Iq, Jq)];
      Rs[e] = Rq[e];
    }
    __syncthreads();
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (int kk = 0; kk < TS; ++kk) {
      float a[8], b[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = Xs[(ty + 16 * i) * LDX + kk];
#pragma unroll
      for (int j = 0; j < 8; ++j) b[j] = Rs[kk * TS + tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) Xs[(ty + 16 * i) * LDX + tx + 16 * j] = acc[i][j];
    for (int e = threadIdx.x; e < TS * TS; e += NT) Rs[e] = Rp[e];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (int kk = 0; kk < TS; ++kk) {
      float a[8], b[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = Rs[kk * TS + ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 8; ++j) b[j] = Xs[kk * LDX + tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int gr = tile_gidx(ty + 16 * i, Ip, Jp);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int gc = tile_gidx(tx + 16 * j, Iq, Jq);
        Kg[(long long)gr * ld + gc] = acc[i][j];
        Kg[(long long)gc * ld + gr] = acc[i][j];
      }
    }
  } else {
    task -= n_ktasks;
    const int rb = task / npairs, q = task - rb * npairs;
    const int Iq = pairs[2 * q], Jq = pairs[2 * q + 1];
    const float* Rq = Rp_all + (long long)q * (TS * TS);
    const int row0 = rb * TS;
    for (int e = threadIdx.x; e < TS * TS; e += NT) {
      const int r = e >> 7, c = e & (TS - 1);
      Xs[r * LDX + c] = Vg[(long long)(row0 + r) * ld + tile_gidx(c, Iq, Jq)];
      Rs[e] = Rq[e];
    }
    __syncthreads();
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (int kk = 0; kk < TS; ++kk) {
      float a[8], b[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = Xs[(ty + 16 * i) * LDX + kk];
#pragma unroll
      for (int j = 0; j < 8; ++j) b[j] = Rs[kk * TS + tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int gr = row0 + ty + 16 * i;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        Vg[(long long)gr * ld + tile_gidx(tx + 16 * j, Iq, Jq)] = acc[i][j];
    }
  }
}

// per-sweep convergence bookkeeping: done[p] |= conv[p] <= tol ; conv[p] = 0
__global__ void k_bj_check(float* conv, int* done, int* sweeps, float tol, int nprob) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nprob) return;
  if (!done[p]) {
    sweeps[p] += 1;
    if (conv[p] <= tol) done[p] = 1;
  }
  conv[p] = 0.f;
}

// prepare: V = I, scale = max |diag|, flags reset; also symmetrise K and zero the padding
__global__ void __launch_bounds__(NT)
k_bj_prepare(float* __restrict__ K, float* __restrict__ V, int ld, long long stride, int n_pad,
             const int* __restrict__ n_dev, int n_fixed, float* __restrict__ scale,
             float* __restrict__ conv, int* __restrict__ done, int* __restrict__ sweeps) {
  __shared__ int redmax;
  const int prob = blockIdx.y;
  const int n = n_dev ? n_dev[prob] : n_fixed;
  float* Kg = K + (long long)prob * stride;
  float* Vg = V + (long long)prob * stride;
  // each CTA of grid.x handles a slab of rows
  const int rows_per = (n_pad + gridDim.x - 1) / gridDim.x;
  const int r0 = blockIdx.x * rows_per, r1 = min(n_pad, r0 + rows_per);
  for (int r = r0; r < r1; ++r) {
    for (int c = threadIdx.x; c < n_pad; c += NT) {
      Vg[(long long)r * ld + c] = (r == c) ? 1.f : 0.f;
      if (r >= n || c >= n) Kg[(long long)r * ld + c] = 0.f;
    }
  }
  if (blockIdx.x == 0) {
    if (threadIdx.x == 0) redmax = 0;
    __syncthreads();
    float v = 0.f;
    for (int i = threadIdx.x; i < n; i += NT) v = fmaxf(v, fabsf(Kg[(long long)i * ld + i]));
    v = warp_max(v);
    if ((threadIdx.x & 31) == 0) atomicMax(&redmax, __float_as_int(v));
    __syncthreads();
    if (threadIdx.x == 0) {
      scale[prob] = __int_as_float(redmax);
      conv[prob] = 0.f;
      done[prob] = (n <= 1 || __int_as_float(redmax) == 0.f) ? 1 : 0;
      sweeps[prob] = 0;
    }
  }
}

// sorted eigenvalues (descending) + permutation of V's columns
__global__ void __launch_bounds__(NT)
k_bj_extract(const float* __restrict__ K, int ld, long long stride, const int* __restrict__ n_dev,
             int n_fixed, float* __restrict__ evals, int* __restrict__ perm, int ld_e) {
  extern __shared__ float dg[];
  const int prob = blockIdx.x;
  const int n = n_dev ? n_dev[prob] : n_fixed;
  const float* Kg = K + (long long)prob * stride;
  for (int i = threadIdx.x; i < n; i += NT) dg[i] = Kg[(long long)i * ld + i];
  __syncthreads();
  for (int i = threadIdx.x; i < ld_e; i += NT) {
    if (i >= n) {
      evals[(long long)prob * ld_e + i] = 0.f;
      perm[(long long)prob * ld_e + i] = i;
    }
  }
  for (int i = threadIdx.x; i < n; i += NT) {
    const float li = dg[i];
    int r = 0;
    for (int j = 0; j < n; ++j) {
      const float lj = dg[j];
      r += (lj > li) || (lj == li && j < i);
    }
    evals[(long long)prob * ld_e + r] = li;
    perm[(long long)prob * ld_e + r] = i;
  }
}

// Number of components from a descending spectrum.
//  mode 0: sklearn PCA float n_components   k = #{cumsum(ratio) <= thr} + 1   (_pca.py:652-666)
//  mode 1: AlignMCCA.n_components_var       k = argmax(cumsum(ratio) > thr)   (AlignMCCA.py:174)
//  mode 2: NoCenterPCA float                k = argmax(cumsum(ratio) >= thr)+1 (NoCenterPCA.py:101-103)
//  mode 3: integer request                  k = (int)thr
// Negative eigenvalues are clipped to 0 (sklearn covariance_eigh does the same); result is
// clamped to [kmin, min(kmax, n)].
__global__ void k_select_k(const float* __restrict__ evals, int ld_e, const int* __restrict__ n_dev,
                           int n_fixed, float thr, int mode, int kmin, int kmax,
                           int* __restrict__ k_out, int k_stride, int nprob) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nprob) return;
  const int n = n_dev ? n_dev[p] : n_fixed;
  const float* ev = evals + (long long)p * ld_e;
  int k;
  if (mode == 3) {
    k = (int)thr;
  } else {
    double tot = 0.0;
    for (int i = 0; i < n; ++i) tot += fmax((double)ev[i], 0.0);
    double cum = 0.0;
    k = -1;
    int cnt_le = 0;
    for (int i = 0; i < n; ++i) {
      cum += fmax((double)ev[i], 0.0);
      const double ratio = (tot > 0.0) ? cum / tot : 0.0;
      if (mode == 0) {
        if (ratio <= (double)thr) ++cnt_le;
      } else if (mode == 1) {
        if (k < 0 && ratio > (double)thr) k = i;
      } else {
        if (k < 0 && ratio >= (double)thr) k = i + 1;
      }
    }
    if (mode == 0) k = cnt_le + 1;
    if (mode == 1 && k < 0) k = 0;
    if (mode == 2 && k < 0) k = 1;
  }
  const int hi = min(kmax, n);
  if (k > hi) k = hi;
  if (k < kmin) k = kmin;
  k_out[(long long)p * k_stride] = k;
}

}  // namespace

// =======================================================================================
// C ABI
// =======================================================================================
static size_t tile_smem_bytes(size_t elem = sizeof(float)) {
  return TS * TS * (elem + sizeof(float)) + sizeof(StepBuf) + (1 + TS) * sizeof(int) + 16;
}

extern "C" int cpsd_eig_sym_small(const float* A, int lda, long long strideA, const int* n_dev,
                                  int n_fixed, int nprob, float* evals, int ld_e, float* evecs,
                                  int ldv, long long strideV, int max_sweeps, float tol,
                                  int* sweeps_out, cudaStream_t stream) {
  CPSD_CHECK_ARG(nprob >= 0 && lda >= 0 && ld_e >= 0, "eig_sym_small: bad dims");
  CPSD_CHECK_ARG(n_fixed <= TS, "eig_sym_small: n > 128 (use cpsd_eig_sym_block)");
  if (nprob == 0) return CPSD_OK;
  const size_t smem = tile_smem_bytes();
  CPSD_CUDA(cudaFuncSetAttribute(k_eig_tile<float>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
  k_eig_tile<float><<<nprob, NT, smem, stream>>>(A, lda, strideA, n_dev, n_fixed, evals, ld_e, evecs,
                                                 ldv, strideV, max_sweeps, tol, sweeps_out);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

// fp64 matrix / fp32 eigenvectors variant (A holds doubles, lda / strideA in doubles)
extern "C" int cpsd_eig_sym_small_f64(const double* A, int lda, long long strideA, const int* n_dev,
                                      int n_fixed, int nprob, float* evals, int ld_e, float* evecs,
                                      int ldv, long long strideV, int max_sweeps, float tol,
                                      int* sweeps_out, cudaStream_t stream) {
  CPSD_CHECK_ARG(nprob >= 0 && lda >= 0 && ld_e >= 0, "eig_sym_small_f64: bad dims");
  CPSD_CHECK_ARG(n_fixed <= TS, "eig_sym_small_f64: n > 128");
  if (nprob == 0) return CPSD_OK;
  const size_t smem = tile_smem_bytes(sizeof(double));
  CPSD_CUDA(cudaFuncSetAttribute(k_eig_tile<double>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
  k_eig_tile<double><<<nprob, NT, smem, stream>>>(A, lda, strideA, n_dev, n_fixed, evals, ld_e,
                                                  evecs, ldv, strideV, max_sweeps, tol, sweeps_out);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

// Workspace (device, caller-allocated):
//   Rbuf  : nprob * (n_pad/128) * 128*128 floats
//   fwork : 2 * nprob floats (scale, conv)
//   iwork : 2 * nprob ints (done, sweeps)
//   pairs : (nb-1) * (nb/2) * 2 ints, round-robin schedule for nb = n_pad/64 blocks, as
//           written by cpsd_bj_schedule().
extern "C" int cpsd_bj_schedule(int n_pad, int* pairs_host) {
  CPSD_CHECK_ARG(n_pad > 0 && n_pad % TS == 0, "bj_schedule: n_pad must be a multiple of 128");
  const int nb = n_pad / BS;
  int idx[1024];
  CPSD_CHECK_ARG(nb <= 1024, "bj_schedule: n_pad too large");
  for (int i = 0; i < nb; ++i) idx[i] = i;
  int o = 0;
  for (int r = 0; r < nb - 1; ++r) {
    for (int i = 0; i < nb / 2; ++i) {
      int a = idx[i], b = idx[nb - 1 - i];
      if (a > b) { int t = a; a = b; b = t; }
      pairs_host[o++] = a;
      pairs_host[o++] = b;
    }
    // rotate all but the first
    int last = idx[nb - 1];
    for (int i = nb - 1; i > 1; --i) idx[i] = idx[i - 1];
    idx[1] = last;
  }
  return CPSD_OK;
}

extern "C" int cpsd_eig_sym_block(float* K, float* V, int ld, long long stride, int n_pad,
                                  const int* n_dev, int n_fixed, int nprob, const int* pairs_dev,
                                  float* Rbuf, float* fwork, int* iwork, float* evals, int* perm,
                                  int ld_e, int max_sweeps, float tol, cudaStream_t stream) {
  CPSD_CHECK_ARG(n_pad > 0 && n_pad % TS == 0, "eig_sym_block: n_pad must be a multiple of 128");
  CPSD_CHECK_ARG(ld >= n_pad && ld_e >= n_pad, "eig_sym_block: ld < n_pad");
  if (nprob == 0) return CPSD_OK;
  const int nb = n_pad / BS, npairs = nb / 2, nrounds = nb - 1;
  float* scale = fwork;
  float* conv = fwork + nprob;
  int* done = iwork;
  int* sweeps = iwork + nprob;
  const size_t smem_in = tile_smem_bytes();
  const size_t smem_up = (TS * LDX + TS * TS) * sizeof(float);
  CPSD_CUDA(cudaFuncSetAttribute(k_bj_inner, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_in));
  CPSD_CUDA(cudaFuncSetAttribute(k_bj_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_up));
  k_bj_prepare<<<dim3(16, nprob), NT, 0, stream>>>(K, V, ld, stride, n_pad, n_dev, n_fixed, scale,
                                                   conv, done, sweeps);
  CPSD_LAUNCH_CHECK();
  const int n_ktasks = npairs * (npairs - 1) / 2;
  const int n_vtasks = (n_pad / TS) * npairs;
  for (int sw = 0; sw < max_sweeps; ++sw) {
    for (int r = 0; r < nrounds; ++r) {
      const int* pr = pairs_dev + (size_t)r * npairs * 2;
      k_bj_inner<<<dim3(npairs, nprob), NT, smem_in, stream>>>(K, ld, stride, pr, Rbuf, scale, conv,
                                                               done, r == 0 ? 1 : 0);
      CPSD_LAUNCH_CHECK();
      k_bj_update<<<dim3(n_ktasks + n_vtasks, nprob), NT, smem_up, stream>>>(
          K, V, ld, stride, n_pad, pr, npairs, Rbuf, done);
      CPSD_LAUNCH_CHECK();
    }
    k_bj_check<<<(nprob + 127) / 128, 128, 0, stream>>>(conv, done, sweeps, fmaxf(tol, 2e-6f), nprob);
    CPSD_LAUNCH_CHECK();
  }
  k_bj_extract<<<nprob, NT, n_pad * sizeof(float), stream>>>(K, ld, stride, n_dev, n_fixed, evals,
                                                             perm, ld_e);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

extern "C" int cpsd_select_k(const float* evals, int ld_e, const int* n_dev, int n_fixed, float thr,
                             int mode, int kmin, int kmax, int* k_out, int k_stride, int nprob,
                             cudaStream_t stream) {
  CPSD_CHECK_ARG(mode >= 0 && mode <= 3, "select_k: bad mode");
  if (nprob == 0) return CPSD_OK;
  k_select_k<<<(nprob + 63) / 64, 64, 0, stream>>>(evals, ld_e, n_dev, n_fixed, thr, mode, kmin, kmax,
                                                   k_out, k_stride, nprob);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}
