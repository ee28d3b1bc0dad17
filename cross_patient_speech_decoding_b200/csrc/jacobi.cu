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
#include "bj_tc.cuh"

#define TS 128          // tile size
#define BS 64           // block size of the block-Jacobi
#define NT 256          // threads per CTA (block-Jacobi kernels)
// threads per CTA of the n <= 128 tile solver: one problem per SM and the Jacobi step is bound
// by shared-memory latency, so it wants every warp slot of the SM
#ifndef ET_NT
#define ET_NT 1024
#endif

namespace {

struct StepBuf {
  int pi[TS / 2];
  int pj[TS / 2];
  double c[TS / 2];
  double s[TS / 2];
  double d[TS / 2];   // t * a_pq (diagonal update)
  float cf[TS / 2];   // fp32 copies (eigenvector update, fp32 matrix update): converting per
  float sf[TS / 2];   // use would put an F2F on every element update
  float df[TS / 2];
};

// optional arguments of the n <= 128 tile solver (all null = plain batched solve)
struct EigWarm {
  const int* sel;        // problem of CTA b (null: b)
  const int* out_idx;    // output slot of a problem (null: the problem index)
  const float* V0;       // initial eigenvector accumulators (null: identity)
  int ldv0;
  long long strideV0;
  const int* v0_idx;     // which V0 a problem starts from (< 0: identity; null: V0[0])
  int zero_pad;          // also write zeros to the eigenvector block outside n x n (ldv >= 128)
};

template <typename T> struct RotView;
template <> struct RotView<double> {
  static __device__ __forceinline__ double c(const StepBuf* sb, int k) { return sb->c[k]; }
  static __device__ __forceinline__ double s(const StepBuf* sb, int k) { return sb->s[k]; }
  static __device__ __forceinline__ double d(const StepBuf* sb, int k) { return sb->d[k]; }
};
template <> struct RotView<float> {
  static __device__ __forceinline__ float c(const StepBuf* sb, int k) { return sb->cf[k]; }
  static __device__ __forceinline__ float s(const StepBuf* sb, int k) { return sb->sf[k]; }
  static __device__ __forceinline__ float d(const StepBuf* sb, int k) { return sb->df[k]; }
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
    sb->cf[k] = (float)c;
    sb->sf[k] = (float)s;
    sb->df[k] = (float)d;
  }
  return seen;
}

// NP > 0: compile-time pair count (power of two) -- the block-Jacobi tiles; NP == 0: runtime.
template <typename T, int NP>
__device__ __forceinline__ void tile_apply(T* As, float* Vs, StepBuf* sb, int npairs_rt, int vrows) {
  const int npairs = NP > 0 ? NP : npairs_rt;
  const int nblk = npairs * npairs;
  for (int b = threadIdx.x; b < nblk; b += blockDim.x) {
    const int k = NP > 0 ? b / NP : b / npairs;
    const int l = b - k * npairs;
    const int ik = sb->pi[k], jk = sb->pj[k];
    const T ck = RotView<T>::c(sb, k), sk = RotView<T>::s(sb, k);
    if (k == l) {
      if (sk != (T)0) {
        const T d = RotView<T>::d(sb, k);
        As[ik * TS + ik] -= d;
        As[jk * TS + jk] += d;
        As[ik * TS + jk] = (T)0;
        As[jk * TS + ik] = (T)0;
      }
      continue;
    }
    const int il = sb->pi[l], jl = sb->pj[l];
    const T cl = RotView<T>::c(sb, l), sl = RotView<T>::s(sb, l);
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
  for (int b = threadIdx.x; b < nv; b += blockDim.x) {
    const int r = NP > 0 ? b / NP : b / npairs;
    const int l = b - r * npairs;
    const float cl = sb->cf[l], sl = sb->sf[l];
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
    if (m == TS) tile_apply<T, TS / 2>(As, Vs, sb, npairs, vrows);
    else tile_apply<T, 0>(As, Vs, sb, npairs, vrows);
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
    tile_apply<float, BS>(As, Vs, sb, BS, TS);
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
  for (int i = threadIdx.x; i < m; i += blockDim.x) v = fmaxf(v, fabsf((float)As[i * TS + i]));
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) atomicMax(redmax, __float_as_int(v));
  __syncthreads();
  const float r = __int_as_float(*redmax);
  __syncthreads();
  return r;
}

// ---------------------------------------------------------------------------------------
// Symmetric-storage variants for the n <= 128 tile solver: only the upper triangle (with the
// diagonal) of the matrix lives in shared memory, in rows of LDS = TS + 1 elements, so that
// element (i, j) = As[min * LDS + max] sits in bank (i + j) mod 32 whichever way it is
// addressed -- a warp whose lanes walk consecutive column indices never conflicts, row-wise
// or column-wise.  A Jacobi step then touches every matrix element once (2x2 blocks with
// k <= l only) instead of twice, which halves the shared-memory traffic that bounds the solver.
// ---------------------------------------------------------------------------------------
#define LDS_SYM (TS + 1)
__device__ __forceinline__ int sym_idx(int i, int j) {
  return (i < j) ? i * LDS_SYM + j : j * LDS_SYM + i;
}

// Pair indices of step s AND their rotations in one phase (one barrier less per step).
// t = sign(x) y / (|x| + sqrt(x^2 + y^2)) with x = a_qq - a_pp, y = 2 a_pq is the smaller root
// of t^2 + 2 tau t - 1 = 0 (tau = x / y) written without forming tau: one sqrt, one division
// and one rsqrt per rotation, all in fp64 (see tile_rotations for why not fp32).
template <typename T>
__device__ __forceinline__ float tile_rotations_sym(T* As, StepBuf* sb, int s, int m, int npairs,
                                                    int nrot, float skip_thr) {
  float seen = 0.f;
  const int k = threadIdx.x;
  if (k < npairs) {
    int i, j;
    if (k < nrot) {
      mod_pair(s, k, m, i, j);
    } else {
      // even steps leave two indices unpaired: they ride along with an identity rotation
      i = s >> 1;
      j = (i + npairs) % m;
    }
    sb->pi[k] = i;
    sb->pj[k] = j;
    const double app = (double)As[i * LDS_SYM + i], aqq = (double)As[j * LDS_SYM + j];
    const double apq = (double)As[sym_idx(i, j)];
    double c = 1.0, sn = 0.0, d = 0.0;
    const float aa = (k < nrot) ? fabsf((float)apq) : 0.f;
    seen = aa;
    if (aa > skip_thr) {
      const double x = aqq - app, y = 2.0 * apq;
      const double t = (x >= 0.0 ? y : -y) / (fabs(x) + sqrt(fma(x, x, y * y)));
      if (t == t && fabs(t) <= 1.0) {  // guards inf/nan
        c = rsqrt(fma(t, t, 1.0));
        sn = t * c;
        d = t * apq;
      }
    }
    sb->c[k] = c;
    sb->s[k] = sn;
    sb->d[k] = d;
    sb->cf[k] = (float)c;
    sb->sf[k] = (float)sn;
    sb->df[k] = (float)d;
  }
  return seen;
}

// One 2x2 block: rows (ik, jk) x columns (il, jl), k < l.
template <typename T>
__device__ __forceinline__ void sym_block(T* As, const StepBuf* sb, int k, int l) {
  const T sk = RotView<T>::s(sb, k), sl = RotView<T>::s(sb, l);
  if (sk == (T)0 && sl == (T)0) return;
  const T ck = RotView<T>::c(sb, k), cl = RotView<T>::c(sb, l);
  const int ik = sb->pi[k], jk = sb->pj[k], il = sb->pi[l], jl = sb->pj[l];
  const int xpp = sym_idx(ik, il), xpq = sym_idx(ik, jl);
  const int xqp = sym_idx(jk, il), xqq = sym_idx(jk, jl);
  const T app = As[xpp], apq = As[xpq], aqp = As[xqp], aqq = As[xqq];
  // columns (il, jl) rotated by (cl, sl)
  const T tpp = cl * app - sl * apq, tpq = sl * app + cl * apq;
  const T tqp = cl * aqp - sl * aqq, tqq = sl * aqp + cl * aqq;
  // rows (ik, jk) rotated by (ck, sk)
  As[xpp] = ck * tpp - sk * tqp;
  As[xqp] = sk * tpp + ck * tqp;
  As[xpq] = ck * tpq - sk * tqq;
  As[xqq] = sk * tpq + ck * tqq;
}

template <typename T, int NP>
__device__ __forceinline__ void tile_apply_sym_A(T* As, StepBuf* sb, int npairs_rt) {
  const int npairs = NP > 0 ? NP : npairs_rt;
  // diagonal 2x2 blocks (the rotated pairs themselves)
  for (int k = threadIdx.x; k < npairs; k += blockDim.x) {
    if (RotView<T>::s(sb, k) != (T)0) {
      const int ik = sb->pi[k], jk = sb->pj[k];
      const T d = RotView<T>::d(sb, k);
      As[ik * LDS_SYM + ik] -= d;
      As[jk * LDS_SYM + jk] += d;
      As[sym_idx(ik, jk)] = (T)0;
    }
  }
  if (NP == 64) {
    // strict upper triangle of the 64 x 64 block grid, folded so that every lane has work:
    // slot (rp, c), c < 64, holds row rp (63 - rp entries) followed by row 62 - rp (rp + 1
    // entries); row 31 (32 entries) has a slot line of its own.  32 x 64 = 2048 slots.
    for (int b = threadIdx.x; b < 32 * 64; b += blockDim.x) {
      const int rp = b >> 6, c = b & 63;
      int k, l;
      if (rp < 31) {
        if (c < 63 - rp) { k = rp; l = rp + 1 + c; }
        else { k = 62 - rp; l = c; }                 // (63 - rp) + (c - (63 - rp))
      } else {
        if (c >= 32) continue;
        k = 31; l = 32 + c;
      }
      sym_block<T>(As, sb, k, l);
    }
  } else {
    const int nblk = npairs * npairs;
    for (int b = threadIdx.x; b < nblk; b += blockDim.x) {
      const int k = b / npairs;
      const int l = b - k * npairs;
      if (l > k) sym_block<T>(As, sb, k, l);
    }
  }
}

// Eigenvector update of one step by the threads t0 .. blockDim.x-1 (the first t0 threads are
// busy with the next step's rotations, see tile_sweep_full_sym).
template <int NP>
__device__ __forceinline__ void tile_apply_sym_V(float* Vs, const StepBuf* sb, int npairs_rt,
                                                 int vrows, int t0) {
  const int npairs = NP > 0 ? NP : npairs_rt;
  const int nv = vrows * npairs;
  if ((int)threadIdx.x < t0) return;
  for (int b = threadIdx.x - t0; b < nv; b += blockDim.x - t0) {
    const int r = NP > 0 ? b / NP : b / npairs;
    const int l = b - r * npairs;
    const float cl = sb->cf[l], sl = sb->sf[l];
    if (sl == 0.f) continue;
    const int il = sb->pi[l], jl = sb->pj[l];
    const float vp = Vs[r * TS + il], vq = Vs[r * TS + jl];
    Vs[r * TS + il] = cl * vp - sl * vq;
    Vs[r * TS + jl] = sl * vp + cl * vq;
  }
}

// One sweep.  Per step: [matrix update with rot(s)] | barrier | [rot(s+1) on the first 64
// threads  ||  eigenvector update with rot(s) on the others] | barrier -- the fp64
// sqrt / division / rsqrt chain of the next step's rotations hides behind the eigenvector
// update of the current one (two StepBufs alternate).
template <typename T>
__device__ float tile_sweep_full_sym(T* As, float* Vs, StepBuf* sb2, int m, int vrows, float skip_thr,
                                     int* redmax) {
  const int npairs = m >> 1;
  const int t0 = (vrows > 0 && (int)blockDim.x >= 4 * BS) ? BS : 0;   // threads reserved for rot(s+1)
  int cur = 0;
  float seen = tile_rotations_sym<T>(As, sb2, 0, m, npairs, npairs - 1, skip_thr);
  __syncthreads();
  for (int s = 0; s < m; ++s) {
    StepBuf* sb = sb2 + cur;
    if (m == TS) tile_apply_sym_A<T, TS / 2>(As, sb, npairs);
    else tile_apply_sym_A<T, 0>(As, sb, npairs);
    __syncthreads();
    if (s + 1 < m) {
      const int nrot = ((s + 1) & 1) ? npairs : npairs - 1;
      seen = fmaxf(seen, tile_rotations_sym<T>(As, sb2 + (cur ^ 1), s + 1, m, npairs, nrot, skip_thr));
    }
    if (vrows > 0) {
      if (m == TS) tile_apply_sym_V<TS / 2>(Vs, sb, npairs, vrows, t0);
      else tile_apply_sym_V<0>(Vs, sb, npairs, vrows, t0);
    }
    __syncthreads();
    cur ^= 1;
  }
  if (threadIdx.x == 0) *redmax = 0;
  __syncthreads();
  seen = warp_max(seen);
  if ((threadIdx.x & 31) == 0) atomicMax(redmax, __float_as_int(seen));
  __syncthreads();
  return __int_as_float(*redmax);
}

template <typename T>
__device__ float tile_diag_absmax_sym(const T* As, int m, int* redmax) {
  if (threadIdx.x == 0) *redmax = 0;
  __syncthreads();
  float v = 0.f;
  for (int i = threadIdx.x; i < m; i += blockDim.x) v = fmaxf(v, fabsf((float)As[i * LDS_SYM + i]));
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
__global__ void __launch_bounds__(ET_NT, 1)
k_eig_tile(const T* __restrict__ A, int lda, long long strideA, const int* __restrict__ n_dev,
           int n_fixed, float* __restrict__ evals, int ld_e, float* __restrict__ evecs, int ldv,
           long long strideV, int max_sweeps, float tol, int* __restrict__ sweeps_out,
           EigWarm warm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* As = reinterpret_cast<T*>(smem_raw);                  // upper triangle, rows of LDS_SYM
  float* Vs = reinterpret_cast<float*>(As + TS * LDS_SYM);
  StepBuf* sb = reinterpret_cast<StepBuf*>(Vs + TS * TS);   // two of them (alternating steps)
  int* redmax = reinterpret_cast<int*>(sb + 2);
  int* rank = redmax + 1;  // TS ints

  // warm.sel: CTA b works on problem sel[b]; warm.out_idx: its results go to slot out_idx[prob];
  // warm.V0 / v0_idx: the eigenvector accumulator starts from V0[v0_idx[prob]] (the caller has
  // rotated A into that basis, A' = V0^T A V0, so that A' is nearly diagonal and the sweeps
  // converge in about half as many passes) instead of the identity
  const int prob = warm.sel ? warm.sel[blockIdx.x] : blockIdx.x;
  const int oslot = warm.out_idx ? warm.out_idx[prob] : prob;
  int n = n_dev ? n_dev[prob] : n_fixed;
  if (n > TS) n = TS;
  if (n < 0) n = 0;
  const int m = (n + 1) & ~1;
  const T* Ag = A + (long long)prob * strideA;
  const float* V0 = nullptr;
  if (warm.V0 && evecs) {
    const int vi = warm.v0_idx ? warm.v0_idx[prob] : 0;
    if (vi >= 0) V0 = warm.V0 + (long long)vi * warm.strideV0;
  }

  // upper triangle of the symmetrised input (inputs are Grams computed tile-wise; the average
  // removes last-ulp asymmetry), identity eigenvectors
  for (int e = threadIdx.x; e < TS * TS; e += ET_NT) {
    const int r = e >> 7, c = e & (TS - 1);
    if (r <= c) {
      T v = (T)0;
      if (c < n) v = (T)0.5 * (Ag[(long long)r * lda + c] + Ag[(long long)c * lda + r]);
      As[r * LDS_SYM + c] = v;
    }
    float v0 = (r == c) ? 1.f : 0.f;
    if (V0 && r < n && c < n) v0 = V0[(long long)r * warm.ldv0 + c];
    Vs[e] = v0;
  }
  __syncthreads();

  int sw = 0;
  if (m >= 2) {
    for (; sw < max_sweeps; ++sw) {
      const float dmax = tile_diag_absmax_sym<T>(As, m, redmax);
      const float skip = (sizeof(T) == 8 ? 1e-15f : 1e-9f) * dmax + 1e-37f;
      const float off = tile_sweep_full_sym<T>(As, Vs, sb, m, evecs ? m : 0, skip, redmax);
      if (off <= tol * dmax) { ++sw; break; }
    }
  }
  if (sweeps_out && threadIdx.x == 0) sweeps_out[prob] = sw;

  // sort descending by counting
  for (int i = threadIdx.x; i < n; i += ET_NT) {
    const T li = As[i * LDS_SYM + i];
    int r = 0;
    for (int j = 0; j < n; ++j) {
      const T lj = As[j * LDS_SYM + j];
      r += (lj > li) || (lj == li && j < i);
    }
    rank[i] = r;
  }
  __syncthreads();
  float* ev = evals + (long long)oslot * ld_e;
  for (int i = threadIdx.x; i < ld_e; i += ET_NT) {
    if (i >= n) ev[i] = 0.f;
  }
  for (int i = threadIdx.x; i < n; i += ET_NT) ev[rank[i]] = (float)As[i * LDS_SYM + i];
  if (evecs) {
    float* Vg = evecs + (long long)oslot * strideV;
    for (int e = threadIdx.x; e < n * n; e += ET_NT) {
      const int r = e / n, c = e - r * n;
      Vg[(long long)r * ldv + rank[c]] = Vs[r * TS + c];
    }
    if (warm.zero_pad && n < TS) {
      for (int e = threadIdx.x; e < TS * TS; e += ET_NT) {
        const int r = e >> 7, c = e & (TS - 1);
        if (r >= n || c >= n) Vg[(long long)r * ldv + c] = 0.f;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// Block Jacobi: inner rotation of one pair's diagonal tile.
// grid (npairs, nprob).  pairs: [npairs][2] block indices of this round.  The accumulated
// 128x128 rotation of the tile goes into the rotation log:
//   Rlog[((prob * total_rounds + round_idx) * npairs + pair)][128*128]
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int tile_gidx(int r, int I, int J) {
  return (r < BS) ? (I * BS + r) : (J * BS + r - BS);
}

__global__ void __launch_bounds__(NT)
k_bj_inner(float* __restrict__ K, int ldk, long long strideK, const int* __restrict__ pairs,
           float* __restrict__ Rlog, int total_rounds, int round_idx,
           const float* __restrict__ scale, float* __restrict__ conv,
           const int* __restrict__ done, int full_mode, float* __restrict__ RTbuf) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* As = reinterpret_cast<float*>(smem_raw);
  float* Vs = As + TS * TS;
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
  float* Rg = Rlog + (((long long)prob * total_rounds + round_idx) * npairs + pr) * (TS * TS);
  for (int e = threadIdx.x; e < TS * TS; e += NT) Rg[e] = Vs[e];
  if (RTbuf) {   // transposed copy for the tensor-core update (K-major operand)
    float* RTg = RTbuf + ((long long)prob * npairs + pr) * (TS * TS);
    for (int e = threadIdx.x; e < TS * TS; e += NT) RTg[e] = Vs[(e & (TS - 1)) * TS + (e >> 7)];
  }
}

// ---------------------------------------------------------------------------------------
// Within-block sweep ("W" round, once per sweep): one CTA per 64-block rotates all pairs
// inside its 64x64 diagonal block (64 steps x 32 pairs, 64 KB of shared memory -> 3 CTAs/SM)
// and logs the rotation as the block-diagonal 128x128 matrix diag(R_I, R_J) of the round-0
// pair (I, J) the block belongs to.  K itself is updated by the tile-update kernel, which
// in W rounds also covers the diagonal pair tiles.  grid (nb, nprob).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT)
k_bj_within(const float* __restrict__ K, int ldk, long long strideK, const int* __restrict__ pairs,
            int npairs, float* __restrict__ Rlog, int total_rounds, int round_idx,
            const float* __restrict__ scale, float* __restrict__ conv,
            const int* __restrict__ done, float* __restrict__ RTbuf) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* As = reinterpret_cast<float*>(smem_raw);     // [64][TS]
  float* Vs = As + BS * TS;                           // [64][TS]
  StepBuf* sb = reinterpret_cast<StepBuf*>(Vs + BS * TS);
  int* redmax = reinterpret_cast<int*>(sb + 1);
  const int blk = blockIdx.x, prob = blockIdx.y;
  if (done && done[prob]) return;
  int pr = 0, half = 0;
  for (int u = 0; u < npairs; ++u) {
    if (pairs[2 * u] == blk) { pr = u; half = 0; }
    if (pairs[2 * u + 1] == blk) { pr = u; half = 1; }
  }
  const float* Kg = K + (long long)prob * strideK;
  for (int e = threadIdx.x; e < BS * BS; e += NT) {
    const int r = e >> 6, c = e & (BS - 1);
    As[r * TS + c] = Kg[(long long)(blk * BS + r) * ldk + blk * BS + c];
    Vs[r * TS + c] = (r == c) ? 1.f : 0.f;
  }
  __syncthreads();
  const float sc = scale[prob];
  const float off = tile_sweep_full<float>(As, Vs, sb, BS, BS, 1e-9f * sc + 1e-37f, redmax);
  if (threadIdx.x == 0 && sc > 0.f)
    atomicMax(reinterpret_cast<int*>(conv + prob), __float_as_int(off / sc));
  // rows [half*64, half*64+64) of the pair's block-diagonal rotation (and of its transpose)
  float* Rg = Rlog + (((long long)prob * total_rounds + round_idx) * npairs + pr) * (TS * TS);
  float* RTg = RTbuf ? RTbuf + ((long long)prob * npairs + pr) * (TS * TS) : nullptr;
  for (int e = threadIdx.x; e < BS * TS; e += NT) {
    const int r = e >> 7, c = e & (TS - 1);
    const bool in = (c >> 6) == half;
    const int cc = c & (BS - 1);
    Rg[(half * BS + r) * TS + c] = in ? Vs[r * TS + cc] : 0.f;
    if (RTg) RTg[(half * BS + r) * TS + c] = in ? Vs[cc * TS + r] : 0.f;
  }
}

// ---------------------------------------------------------------------------------------
// Cross-sweep variant of the inner kernel (16 of every 17 rounds): only the 64x64 pairs
// (i in block I, j in block J) rotate, in 64 steps of 64 disjoint pairs
// (l, 64 + (l+s)%64).  The tile stays in shared memory (64 KB -> 2 CTAs/SM at 128 registers), but the
// accumulated rotation R lives in REGISTERS: warp w owns rows 16w..16w+15, lane x holds
// R[r][x], R[r][x+32] and the two J-block columns currently paired with them; the J columns
// ride a 64-slot ring that advances one lane per step (2 shuffles), so the partner of
// column l is always in the same lane.  Rotation parameters are evaluated in fp32 here; the
// resulting O(1e-6) column-norm drift of R is removed by normalising R's columns before
// they are logged.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 2)
k_bj_inner_cross(float* __restrict__ K, int ldk, long long strideK, const int* __restrict__ pairs,
                 float* __restrict__ Rlog, int total_rounds, int round_idx,
                 const float* __restrict__ scale, float* __restrict__ conv,
                 const int* __restrict__ done, float* __restrict__ RTbuf) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* As = reinterpret_cast<float*>(smem_raw);          // [128][128]
  float2* cs = reinterpret_cast<float2*>(As + TS * TS);    // [64] (c, s) of the step
  float* dd = reinterpret_cast<float*>(cs + BS);           // [64] t * a_pq
  float* nrm = dd + BS;                                    // [128] column norms^2 of R
  float* nrm_part = nrm + TS;                              // [8][128] per-warp partial norms
  int* redmax = reinterpret_cast<int*>(nrm_part + (NT / 32) * TS);

  const int pr = blockIdx.x, prob = blockIdx.y, npairs = gridDim.x;
  if (done && done[prob]) return;
  const int I = pairs[2 * pr], J = pairs[2 * pr + 1];
  float* Kg = K + (long long)prob * strideK;
  for (int e = threadIdx.x; e < TS * TS; e += NT) {
    const int r = e >> 7, c = e & (TS - 1);
    As[e] = Kg[(long long)tile_gidx(r, I, J) * ldk + tile_gidx(c, I, J)];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float vI0[16], vI1[16], vJ0[16], vJ1[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int r = warp * 16 + i;
    vI0[i] = (r == lane) ? 1.f : 0.f;
    vI1[i] = (r == lane + 32) ? 1.f : 0.f;
    vJ0[i] = (r == lane + 64) ? 1.f : 0.f;
    vJ1[i] = (r == lane + 96) ? 1.f : 0.f;
  }
  const float sc = scale[prob];
  const float skip = 1e-9f * sc + 1e-37f;
  float seen = 0.f;
  const int l = threadIdx.x & 63, kq = threadIdx.x >> 6;
  __syncthreads();

  for (int s = 0; s < BS; ++s) {
    if (threadIdx.x < BS) {
      const int i = threadIdx.x, j = BS + ((i + s) & (BS - 1));
      const float app = As[i * TS + i], aqq = As[j * TS + j], apq = As[i * TS + j];
      float c = 1.f, sn = 0.f, d = 0.f;
      const float aa = fabsf(apq);
      seen = fmaxf(seen, aa);
      if (aa > skip) {
        const float tau = (aqq - app) / (2.f * apq);
        const float t = copysignf(1.f, tau) / (fabsf(tau) + sqrtf(1.f + tau * tau));
        if (t == t && fabsf(t) <= 1.f) {
          c = 1.f / sqrtf(1.f + t * t);
          sn = t * c;
          d = t * apq;
        }
      }
      cs[i] = make_float2(c, sn);
      dd[i] = d;
    }
    __syncthreads();
    // ---- tile update: thread owns column pair l, walks row pairs k = kq, kq+4, ...
    {
      const float2 csl = cs[l];
      const int jl = BS + ((l + s) & (BS - 1));
#pragma unroll 4
      for (int i = 0; i < 16; ++i) {
        const int k = kq + 4 * i;
        const float2 csk = cs[k];
        const int jk = BS + ((k + s) & (BS - 1));
        if (k == l) {
          if (csk.y != 0.f) {
            const float d = dd[k];
            As[k * TS + k] -= d;
            As[jk * TS + jk] += d;
            As[k * TS + jk] = 0.f;
            As[jk * TS + k] = 0.f;
          }
        } else if (csk.y != 0.f || csl.y != 0.f) {
          const float app = As[k * TS + l], apq = As[k * TS + jl];
          const float aqp = As[jk * TS + l], aqq = As[jk * TS + jl];
          const float tpp = csl.x * app - csl.y * apq, tpq = csl.y * app + csl.x * apq;
          const float tqp = csl.x * aqp - csl.y * aqq, tqq = csl.y * aqp + csl.x * aqq;
          As[k * TS + l] = csk.x * tpp - csk.y * tqp;
          As[jk * TS + l] = csk.y * tpp + csk.x * tqp;
          As[k * TS + jl] = csk.x * tpq - csk.y * tqq;
          As[jk * TS + jl] = csk.y * tpq + csk.x * tqq;
        }
      }
    }
    // ---- R update in registers: column lane (+32) pairs with ring slot lane (+32)
    {
      const float2 c0 = cs[lane], c1 = cs[lane + 32];
      const int src = (lane + 1) & 31;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float p0 = vI0[i], q0 = vJ0[i], p1 = vI1[i], q1 = vJ1[i];
        vI0[i] = c0.x * p0 - c0.y * q0;
        const float nq0 = c0.y * p0 + c0.x * q0;
        vI1[i] = c1.x * p1 - c1.y * q1;
        const float nq1 = c1.y * p1 + c1.x * q1;
        const float a = __shfl_sync(0xffffffffu, nq0, src);
        const float b = __shfl_sync(0xffffffffu, nq1, src);
        vJ0[i] = (lane == 31) ? b : a;
        vJ1[i] = (lane == 31) ? a : b;
      }
    }
    __syncthreads();
  }
  // convergence measure
  if (threadIdx.x == 0) *redmax = 0;
  __syncthreads();
  seen = warp_max(seen);
  if (lane == 0) atomicMax(redmax, __float_as_int(seen));
  // column norms of R (rows are spread over the 8 warps): per-warp partials, summed in a fixed
  // order (shared-memory float atomics would make the last bit depend on the warp schedule)
  __syncthreads();
  {
    float n0 = 0.f, n1 = 0.f, n2 = 0.f, n3 = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      n0 = fmaf(vI0[i], vI0[i], n0);
      n1 = fmaf(vI1[i], vI1[i], n1);
      n2 = fmaf(vJ0[i], vJ0[i], n2);
      n3 = fmaf(vJ1[i], vJ1[i], n3);
    }
    float* pw = nrm_part + warp * TS;
    pw[lane] = n0;
    pw[lane + 32] = n1;
    pw[lane + 64] = n2;
    pw[lane + 96] = n3;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < TS; c += NT) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) t += nrm_part[w * TS + c];
    nrm[c] = t;
  }
  __syncthreads();
  if (threadIdx.x == 0 && sc > 0.f)
    atomicMax(reinterpret_cast<int*>(conv + prob), __float_as_int(__int_as_float(*redmax) / sc));
  for (int e = threadIdx.x; e < TS * TS; e += NT) {
    const int r = e >> 7, c = e & (TS - 1);
    Kg[(long long)tile_gidx(r, I, J) * ldk + tile_gidx(c, I, J)] = As[e];
  }
  float* Rg = Rlog + (((long long)prob * total_rounds + round_idx) * npairs + pr) * (TS * TS);
  const float s0 = rsqrtf(nrm[lane]), s1 = rsqrtf(nrm[lane + 32]);
  const float s2 = rsqrtf(nrm[lane + 64]), s3 = rsqrtf(nrm[lane + 96]);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float* row = Rg + (warp * 16 + i) * TS;
    row[lane] = vI0[i] * s0;
    row[lane + 32] = vI1[i] * s1;
    row[lane + 64] = vJ0[i] * s2;
    row[lane + 96] = vJ1[i] * s3;
  }
  if (RTbuf) {
    // transposed copy through the (now free) tile buffer, XOR-swizzled so that both the row
    // writes and the column reads are bank-conflict free
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int r = warp * 16 + i, x = r & 31;
      As[r * TS + (lane ^ x)] = vI0[i] * s0;
      As[r * TS + 32 + (lane ^ x)] = vI1[i] * s1;
      As[r * TS + 64 + (lane ^ x)] = vJ0[i] * s2;
      As[r * TS + 96 + (lane ^ x)] = vJ1[i] * s3;
    }
    __syncthreads();
    float* RTg = RTbuf + ((long long)prob * npairs + pr) * (TS * TS);
    for (int e = threadIdx.x; e < TS * TS; e += NT) {
      const int c = e >> 7, r = e & (TS - 1);
      RTg[e] = As[r * TS + (c & ~31) + ((c & 31) ^ (r & 31))];
    }
  }
}

// ---------------------------------------------------------------------------------------
// Block Jacobi: apply the round's rotations to the off-diagonal tiles of K (p < q, both
// mirror images written):  K[p,q] <- R_p^T K[p,q] R_q.   grid (npairs(npairs-1)/2, nprob)
// The eigenvector matrix is NOT touched here: rotations are replayed on the wanted columns
// afterwards (k_bj_backapply), which costs k/n of the classical accumulate-every-round.
// ---------------------------------------------------------------------------------------
#define LDX 129
__global__ void __launch_bounds__(NT)
k_bj_update(float* __restrict__ K, int ld, long long stride, const int* __restrict__ pairs,
            int npairs, const float* __restrict__ Rlog, int total_rounds, int round_idx,
            const int* __restrict__ done) {
  extern __shared__ float smem[];
  float* Xs = smem;                 // [128][129]
  float* Rs = smem + TS * LDX;      // [128][128]
  const int prob = blockIdx.y;
  if (done && done[prob]) return;
  const int task = blockIdx.x;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float* Kg = K + (long long)prob * stride;
  const float* Rp_all = Rlog + ((long long)prob * total_rounds + round_idx) * npairs * (TS * TS);

  // decode (p, q), p < q; tasks beyond the strict upper triangle are the diagonal tiles (W rounds)
  const int n_off = npairs * (npairs - 1) / 2;
  int p = 0, q;
  if (task < n_off) {
    int rem = task;
    while (rem >= npairs - 1 - p) { rem -= npairs - 1 - p; ++p; }
    q = p + 1 + rem;
  } else {
    p = q = task - n_off;
  }
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
      if (p != q) Kg[(long long)gc * ld + gr] = acc[i][j];
    }
  }
}

// tensor-core version of k_bj_update (see bj_tc.cuh); RTbuf holds this round's transposed
// rotations, one 128x128 block per (problem, pair)
__global__ void __launch_bounds__(NT, 1)
k_bj_update_tc(float* __restrict__ K, int ld, long long stride, const int* __restrict__ pairs,
               int npairs, const float* __restrict__ RTbuf, const int* __restrict__ done) {
  extern __shared__ __align__(1024) unsigned char smem_tc[];
  const int prob = blockIdx.y;
  if (done && done[prob]) return;
  const int n_off = npairs * (npairs - 1) / 2;
  int p = 0, q;
  if ((int)blockIdx.x < n_off) {
    int rem = blockIdx.x;
    while (rem >= npairs - 1 - p) { rem -= npairs - 1 - p; ++p; }
    q = p + 1 + rem;
  } else {
    p = q = blockIdx.x - n_off;          // diagonal pair tile (W rounds only)
  }
  const float* RT = RTbuf + (long long)prob * npairs * (TS * TS);
  bjtc::update_tile_tc(K + (long long)prob * stride, ld, RT + (long long)p * (TS * TS),
                       RT + (long long)q * (TS * TS), pairs[2 * p], pairs[2 * p + 1],
                       pairs[2 * q], pairs[2 * q + 1], p != q, smem_tc,
                       [](int r, int I, int J) { return tile_gidx(r, I, J); });
}

// ---------------------------------------------------------------------------------------
// Eigenvectors on demand.  V = R_1 R_2 ... R_L (one block-diagonal rotation per round), so
// the wanted columns are  E = R_1 ( R_2 ( ... ( R_L  I[:, sel] ))).
// k_bj_einit writes I[:, sel] (sel = perm[0..k)), k_bj_backapply applies one round:
//   E[rows of pair p, 64-column chunk] <- R_p E[rows of pair p, chunk]
// grid (npairs * nchunks, nprob); chunks beyond k[prob] and rounds beyond the problem's own
// sweep count exit immediately.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT)
k_bj_einit(float* __restrict__ E, int lde, long long strideE, int n_pad,
           const int* __restrict__ perm, int ld_perm, const int* __restrict__ k_dev, int k_fixed) {
  const int prob = blockIdx.y;
  int k = k_dev ? k_dev[prob] : k_fixed;
  if (k > lde) k = lde;
  float* Ep = E + (long long)prob * strideE;
  const int* pp = perm + (long long)prob * ld_perm;
  for (int r = blockIdx.x; r < n_pad; r += gridDim.x)
    for (int j = threadIdx.x; j < k; j += NT) Ep[(long long)r * lde + j] = (pp[j] == r) ? 1.f : 0.f;
}

#define EC 32
__global__ void __launch_bounds__(NT)
k_bj_backapply(float* __restrict__ E, int lde, long long strideE, const int* __restrict__ pairs,
               int npairs, const float* __restrict__ Rlog, int total_rounds, int round_idx,
               int nrounds, const int* __restrict__ sweeps, const int* __restrict__ k_dev,
               int k_fixed) {
  extern __shared__ float smem[];
  float* Rs = smem;               // [128][129]
  float* Es = smem + TS * LDX;    // [128][EC]
  const int prob = blockIdx.y;
  if (round_idx >= sweeps[prob] * nrounds) return;
  int k = k_dev ? k_dev[prob] : k_fixed;
  if (k > lde) k = lde;
  const int pr = blockIdx.x;
  const int I = pairs[2 * pr], J = pairs[2 * pr + 1];
  float* Ep = E + (long long)prob * strideE;
  const float* Rg = Rlog + (((long long)prob * total_rounds + round_idx) * npairs + pr) * (TS * TS);
  for (int e = threadIdx.x; e < TS * TS; e += NT) Rs[(e >> 7) * LDX + (e & (TS - 1))] = Rg[e];
  // 128 x 32 outputs per chunk, 256 threads: rows ty + 16 i (i < 8), cols tx + 16 j (j < 2)
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  for (int c0 = 0; c0 < k; c0 += EC) {
    __syncthreads();
    for (int e = threadIdx.x; e < TS * EC; e += NT) {
      const int r = e / EC, c = e - r * EC;
      Es[e] = (c0 + c < k) ? Ep[(long long)tile_gidx(r, I, J) * lde + c0 + c] : 0.f;
    }
    __syncthreads();
    float acc[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; }
#pragma unroll 4
    for (int kk = 0; kk < TS; ++kk) {
      float a[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = Rs[(ty + 16 * i) * LDX + kk];
      const float b0 = Es[kk * EC + tx], b1 = Es[kk * EC + tx + 16];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc[i][0] = fmaf(a[i], b0, acc[i][0]);
        acc[i][1] = fmaf(a[i], b1, acc[i][1]);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int gr = tile_gidx(ty + 16 * i, I, J);
      if (c0 + tx < k) Ep[(long long)gr * lde + c0 + tx] = acc[i][0];
      if (c0 + tx + 16 < k) Ep[(long long)gr * lde + c0 + tx + 16] = acc[i][1];
    }
  }
}

// per-sweep convergence bookkeeping: done[p] |= conv[p] <= tol ; conv[p] = 0
__global__ void k_bj_check(float* conv, int* done, int* sweeps, float* hist, int hist_len,
                           float tol, int nprob) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nprob) return;
  if (!done[p]) {
    if (hist && sweeps[p] < hist_len) hist[(long long)p * hist_len + sweeps[p]] = conv[p];
    sweeps[p] += 1;
    if (conv[p] <= tol) done[p] = 1;
  }
  conv[p] = 0.f;
}

// prepare: V = I, scale = max |diag|, flags reset; also symmetrise K and zero the padding
__global__ void __launch_bounds__(NT)
k_bj_prepare(float* __restrict__ K, int ld, long long stride, int n_pad,
             const int* __restrict__ n_dev, int n_fixed, float* __restrict__ scale,
             float* __restrict__ conv, int* __restrict__ done, int* __restrict__ sweeps) {
  __shared__ int redmax;
  const int prob = blockIdx.y;
  const int n = n_dev ? n_dev[prob] : n_fixed;
  float* Kg = K + (long long)prob * stride;
  // each CTA of grid.x handles a slab of rows
  const int rows_per = (n_pad + gridDim.x - 1) / gridDim.x;
  const int r0 = blockIdx.x * rows_per, r1 = min(n_pad, r0 + rows_per);
  for (int r = r0; r < r1; ++r) {
    for (int c = threadIdx.x; c < n_pad; c += NT) {
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
                           int n_fixed, const float* __restrict__ tot_dev, float thr, int mode,
                           int kmin, int kmax, int* __restrict__ k_out, int k_stride, int nprob) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nprob) return;
  const int n = n_dev ? n_dev[p] : n_fixed;
  const float* ev = evals + (long long)p * ld_e;
  int k;
  if (mode == 3) {
    k = (int)thr;
  } else {
    double tot = 0.0;
    if (tot_dev) tot = (double)tot_dev[p];   // only the leading n values are known: total given
    else for (int i = 0; i < n; ++i) tot += fmax((double)ev[i], 0.0);
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
  return TS * LDS_SYM * elem + TS * TS * sizeof(float) + 2 * sizeof(StepBuf) + (1 + TS) * sizeof(int) + 16;
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
  k_eig_tile<float><<<nprob, ET_NT, smem, stream>>>(A, lda, strideA, n_dev, n_fixed, evals, ld_e, evecs,
                                                 ldv, strideV, max_sweeps, tol, sweeps_out, EigWarm{});
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
  k_eig_tile<double><<<nprob, ET_NT, smem, stream>>>(A, lda, strideA, n_dev, n_fixed, evals, ld_e,
                                                  evecs, ldv, strideV, max_sweeps, tol, sweeps_out,
                                                  EigWarm{});
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

// Same solver on a selected subset of a batch, optionally warm-started: CTA b solves problem
// sel[b] (sel null: b) of A, writes eigenvalues / sorted eigenvectors to slot out_idx[problem]
// (null: the problem index) and starts its eigenvector accumulator from V0[v0_idx[problem]]
// (entry < 0: identity); the eigenvector block is zero outside n x n.  The caller passes A already rotated into that basis (V0^T A V0, see
// cpsd_dgemm_batched): the per-fold scatter matrices of a cross-validation differ from their
// all-trials version by a few percent, so in its eigenbasis they are nearly diagonal.
extern "C" int cpsd_eig_sym_small_f64_warm(const double* A, int lda, long long strideA, const int* n_dev,
                                           int n_fixed, const int* sel, int nsel, const int* out_idx,
                                           float* evals, int ld_e, float* evecs, int ldv,
                                           long long strideV, const float* V0, int ldv0,
                                           long long strideV0, const int* v0_idx, int max_sweeps,
                                           float tol, int* sweeps_out, cudaStream_t stream) {
  CPSD_CHECK_ARG(nsel >= 0 && lda >= 0 && ld_e >= 0, "eig_sym_small_f64_warm: bad dims");
  CPSD_CHECK_ARG(n_fixed <= TS, "eig_sym_small_f64_warm: n > 128");
  CPSD_CHECK_ARG(V0 == nullptr || ldv0 > 0, "eig_sym_small_f64_warm: bad V0 stride");
  CPSD_CHECK_ARG(evecs == nullptr || ldv >= TS, "eig_sym_small_f64_warm: ldv < 128");
  if (nsel == 0) return CPSD_OK;
  const size_t smem = tile_smem_bytes(sizeof(double));
  CPSD_CUDA(cudaFuncSetAttribute(k_eig_tile<double>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
  EigWarm w{sel, out_idx, V0, ldv0, strideV0, v0_idx, 1};
  k_eig_tile<double><<<nsel, ET_NT, smem, stream>>>(A, lda, strideA, n_dev, n_fixed, evals, ld_e,
                                                 evecs, ldv, strideV, max_sweeps, tol, sweeps_out, w);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

// Block Jacobi eigen-solver for n_pad = multiple of 128 (> 128).  K is destroyed.
// Workspace (device, caller-allocated):
//   Rlog  : nprob * max_sweeps * nb * (nb/2) * 128*128 floats, nb = n_pad/64 (rotation log:
//           per sweep one within-block round + nb-1 cross rounds)
//   fwork : (2 + 16) * nprob floats (scale, conv, per-sweep convergence history)
//   iwork : 2 * nprob ints (done, sweeps)
//   pairs : (nb-1) * (nb/2) * 2 ints, round-robin schedule as written by cpsd_bj_schedule().
// Outputs: evals (descending) and perm (position of each eigenvalue on the diagonal); the
// eigenvectors of any leading subset follow from cpsd_bj_eigvecs().
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

#define BJ_HIST 16

extern "C" long long cpsd_bj_rlog_elems(int n_pad, int nprob, int max_sweeps) {
  const long long nb = n_pad / BS;
  return (long long)nprob * max_sweeps * nb * (nb / 2) * (TS * TS);
}

extern "C" int cpsd_eig_sym_block(float* K, int ld, long long stride, int n_pad, const int* n_dev,
                                  int n_fixed, int nprob, const int* pairs_dev, float* Rlog,
                                  float* fwork, int* iwork, float* evals, int* perm, int ld_e,
                                  int max_sweeps, float tol, float* RTbuf, cudaStream_t stream) {
  CPSD_CHECK_ARG(n_pad > 0 && n_pad % TS == 0, "eig_sym_block: n_pad must be a multiple of 128");
  CPSD_CHECK_ARG(ld >= n_pad && ld_e >= n_pad, "eig_sym_block: ld < n_pad");
  if (nprob == 0) return CPSD_OK;
  const int nb = n_pad / BS, npairs = nb / 2, nrounds = nb - 1;
  const int rps = nrounds + 1;                 // rounds per sweep: 1 within-block + nrounds cross
  const int total_rounds = max_sweeps * rps;
  float* scale = fwork;
  float* conv = fwork + nprob;
  float* hist = fwork + 2 * nprob;
  int* done = iwork;
  int* sweeps = iwork + nprob;
  const size_t smem_w = 2 * BS * TS * sizeof(float) + sizeof(StepBuf) + 32;
  const size_t smem_up = (TS * LDX + TS * TS) * sizeof(float);
  const size_t smem_cross = TS * TS * sizeof(float) + BS * 12 + TS * 4 + (NT / 32) * TS * 4 + 16;
  CPSD_CUDA(cudaFuncSetAttribute(k_bj_inner_cross, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem_cross));
  CPSD_CUDA(cudaFuncSetAttribute(k_bj_within, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w));
  CPSD_CUDA(cudaFuncSetAttribute(k_bj_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_up));
  CPSD_CUDA(cudaFuncSetAttribute(k_bj_update_tc, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 bjtc::SMEM_BYTES));
  k_bj_prepare<<<dim3(16, nprob), NT, 0, stream>>>(K, ld, stride, n_pad, n_dev, n_fixed, scale, conv,
                                                   done, sweeps);
  CPSD_LAUNCH_CHECK();
  const int n_off = npairs * (npairs - 1) / 2;
  for (int sw = 0; sw < max_sweeps; ++sw) {
    for (int rr = 0; rr < rps; ++rr) {
      // rr == 0: within-block round on the pairing of cross round 0; rr >= 1: cross round rr-1
      const int* pr = pairs_dev + (size_t)(rr == 0 ? 0 : rr - 1) * npairs * 2;
      const int ridx = sw * rps + rr;
      if (rr == 0)
        k_bj_within<<<dim3(nb, nprob), NT, smem_w, stream>>>(K, ld, stride, pr, npairs, Rlog,
                                                             total_rounds, ridx, scale, conv, done,
                                                             RTbuf);
      else
        k_bj_inner_cross<<<dim3(npairs, nprob), NT, smem_cross, stream>>>(
            K, ld, stride, pr, Rlog, total_rounds, ridx, scale, conv, done, RTbuf);
      CPSD_LAUNCH_CHECK();
      const int ntask = n_off + (rr == 0 ? npairs : 0);   // W rounds also rotate the diagonal tiles
      if (ntask > 0 && RTbuf) {
        k_bj_update_tc<<<dim3(ntask, nprob), NT, bjtc::SMEM_BYTES, stream>>>(K, ld, stride, pr,
                                                                            npairs, RTbuf, done);
        CPSD_LAUNCH_CHECK();
      } else if (ntask > 0) {
        k_bj_update<<<dim3(ntask, nprob), NT, smem_up, stream>>>(K, ld, stride, pr, npairs, Rlog,
                                                                 total_rounds, ridx, done);
        CPSD_LAUNCH_CHECK();
      }
    }
    k_bj_check<<<(nprob + 127) / 128, 128, 0, stream>>>(conv, done, sweeps, hist, BJ_HIST,
                                                        fmaxf(tol, 2e-6f), nprob);
    CPSD_LAUNCH_CHECK();
  }
  k_bj_extract<<<nprob, NT, n_pad * sizeof(float), stream>>>(K, ld, stride, n_dev, n_fixed, evals,
                                                             perm, ld_e);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

// Leading k eigenvectors (columns sorted like evals) from the rotation log of
// cpsd_eig_sym_block: E (n_pad x lde, only the first k[prob] columns are written).
// k_launch bounds the columns the grid covers (>= every k[prob]).
extern "C" int cpsd_bj_eigvecs(const float* Rlog, int n_pad, int nprob, const int* pairs_dev,
                               const int* iwork, const int* perm, int ld_perm, const int* k_dev,
                               int k_fixed, int k_launch, float* E, int lde, long long strideE,
                               int max_sweeps, cudaStream_t stream) {
  CPSD_CHECK_ARG(n_pad > 0 && n_pad % TS == 0, "bj_eigvecs: n_pad must be a multiple of 128");
  CPSD_CHECK_ARG(k_launch > 0 && k_launch <= lde, "bj_eigvecs: bad k_launch");
  if (nprob == 0) return CPSD_OK;
  const int nb = n_pad / BS, npairs = nb / 2, nrounds = nb - 1;
  const int rps = nrounds + 1;
  const int total_rounds = max_sweeps * rps;
  const int* sweeps = iwork + nprob;
  k_bj_einit<<<dim3(n_pad < 256 ? n_pad : 256, nprob), NT, 0, stream>>>(E, lde, strideE, n_pad, perm,
                                                                        ld_perm, k_dev, k_fixed);
  CPSD_LAUNCH_CHECK();
  const size_t smem = (TS * LDX + TS * EC) * sizeof(float);
  CPSD_CUDA(cudaFuncSetAttribute(k_bj_backapply, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int ridx = total_rounds - 1; ridx >= 0; --ridx) {
    const int rr = ridx % rps;
    const int* pr = pairs_dev + (size_t)(rr == 0 ? 0 : rr - 1) * npairs * 2;
    k_bj_backapply<<<dim3(npairs, nprob), NT, smem, stream>>>(
        E, lde, strideE, pr, npairs, Rlog, total_rounds, ridx, rps, sweeps, k_dev, k_fixed);
    CPSD_LAUNCH_CHECK();
  }
  return CPSD_OK;
}

extern "C" int cpsd_select_k(const float* evals, int ld_e, const int* n_dev, int n_fixed, float thr,
                             int mode, int kmin, int kmax, int* k_out, int k_stride, int nprob,
                             cudaStream_t stream) {
  CPSD_CHECK_ARG(mode >= 0 && mode <= 3, "select_k: bad mode");
  if (nprob == 0) return CPSD_OK;
  k_select_k<<<(nprob + 63) / 64, 64, 0, stream>>>(evals, ld_e, n_dev, n_fixed, nullptr, thr, mode,
                                                   kmin, kmax, k_out, k_stride, nprob);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

// Same selection when only the leading n eigenvalues are known and the total variance (the
// trace) is given per problem (cpsd_eig_sym_topk).
extern "C" int cpsd_select_k_total(const float* evals, int ld_e, const int* n_dev, int n_fixed,
                                   const float* total_dev, float thr, int mode, int kmin, int kmax,
                                   int* k_out, int k_stride, int nprob, cudaStream_t stream) {
  CPSD_CHECK_ARG(mode >= 0 && mode <= 3, "select_k_total: bad mode");
  CPSD_CHECK_ARG(total_dev != nullptr, "select_k_total: total_dev is NULL");
  if (nprob == 0) return CPSD_OK;
  k_select_k<<<(nprob + 63) / 64, 64, 0, stream>>>(evals, ld_e, n_dev, n_fixed, total_dev, thr, mode,
                                                   kmin, kmax, k_out, k_stride, nprob);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}
