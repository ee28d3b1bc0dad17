// Kernel C-SVC with one-vs-one voting: the reference scripts' literal decoder
// `SVC(kernel='rbf', class_weight='balanced')` (aligned_decoding/scripts/aligned_decode_svm_ncv.py:313-317,
// aligned_decode_grid_subsample.py:261) and the `SVC(kernel='linear')` inside its
// `BaggingClassifier` (aligned_decode_svm.py:262-263).  sklearn hands both to libsvm; this file
// restates libsvm's published algorithm (Fan, Chen, Lin 2005: SMO with second-order working-set
// selection, float kernel values, double gradients, eps = tol stopping rule, first-maximum vote)
// as batched kernels:
//   k_svc_gamma      gamma='scale' = 1 / (n_features * var(X)) per fold
//   k_svc_perm       class-sorted order of the pool + squared norms
//   k_svc_kmat       (n x n) kernel matrix per fold in class-sorted order (fp64 accumulation,
//                    stored as float like libsvm's Qfloat cache; upper tiles, mirrored)
//   k_svc_smo        one WARP per (fold, class pair) dual problem: members, alpha and the
//                    gradient live in shared memory, no block barrier inside the SMO loop
//   k_svc_predict    one CTA per (fold, held-out trial): kernel row, all pair decisions, vote
// No shrinking (it changes the iteration path, not the optimum the eps rule accepts).
#include "common.cuh"

namespace {

constexpr int SVC_WARPS = 2;       // tasks per CTA of k_svc_smo
constexpr double SVC_TAU = 1e-12;
constexpr double SVC_INF = 1.0e300;

__global__ void k_svc_gamma(const float* __restrict__ St, int lds, long long strideS,
                            const int* __restrict__ k_dev, int k_fixed,
                            const int* __restrict__ n_dev, int n_fixed, int kernel, double gamma_in,
                            double* __restrict__ gamma_out) {
  __shared__ double red[40];
  const int f = blockIdx.x;
  const int k = k_dev ? k_dev[f] : k_fixed;
  const int n = n_dev ? n_dev[f] : n_fixed;
  double g = gamma_in;
  if (kernel == 1 && gamma_in <= 0.0) {       // 'scale'
    const float* X = St + (long long)f * strideS;
    double s = 0.0;
    for (int j = 0; j < k; ++j)
      for (int t = threadIdx.x; t < n; t += blockDim.x) s += (double)X[(long long)j * lds + t];
    const double cnt = (double)k * (double)n;
    const double mean = block_sum(s, red) / fmax(cnt, 1.0);
    double v = 0.0;
    for (int j = 0; j < k; ++j)
      for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const double d = (double)X[(long long)j * lds + t] - mean;
        v += d * d;
      }
    const double var = block_sum(v, red) / fmax(cnt, 1.0);
    g = (var > 0.0 && k > 0) ? 1.0 / ((double)k * var) : 1.0;
  }
  if (threadIdx.x == 0) gamma_out[f] = g;
}

// One CTA per fold: class-sorted order of the pool (stable counting sort by class, warp 0) and
// the squared norms of the samples (all threads).  perm[pos] = pool index of the sample at
// sorted position pos; cls_off[c] .. cls_off[c+1] = positions of class c.
__global__ void k_svc_perm(const float* __restrict__ St, int lds, long long strideS,
                           const int* __restrict__ k_dev, int k_fixed,
                           const int* __restrict__ n_dev, int n_fixed, const int* __restrict__ y,
                           int ldy, const int* __restrict__ classes, int ncls,
                           int* __restrict__ perm, int ldp, int* __restrict__ cls_off,
                           double* __restrict__ sqn) {
  const int f = blockIdx.x;
  const int k = k_dev ? k_dev[f] : k_fixed;
  const int n = n_dev ? n_dev[f] : n_fixed;
  const float* X = St + (long long)f * strideS;
  const int* yf = y + (long long)f * ldy;
  for (int t = threadIdx.x; t < n; t += blockDim.x) {
    double v = 0.0;
    for (int j = 0; j < k; ++j) {
      const double x = (double)X[(long long)j * lds + t];
      v = fma(x, x, v);
    }
    sqn[(long long)f * ldp + t] = v;
  }
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    int pos = 0;
    for (int c = 0; c < ncls; ++c) {
      if (lane == 0) cls_off[(long long)f * (ncls + 1) + c] = pos;
      const int want = classes[c];
      for (int t0 = 0; t0 < n; t0 += 32) {
        const int t = t0 + lane;
        const bool hit = t < n && yf[t] == want;
        const unsigned msk = __ballot_sync(0xffffffffu, hit);
        if (hit) perm[(long long)f * ldp + pos + __popc(msk & ((1u << lane) - 1u))] = t;
        pos += __popc(msk);
      }
    }
    if (lane == 0) cls_off[(long long)f * (ncls + 1) + ncls] = pos;
  }
}

// Kernel matrix in class-sorted order, K'[p][q] = k(x_perm[p], x_perm[q]): 32x32 tiles of the
// upper triangle (mirrored on write), 256 threads x 4 entries, features staged 32 at a time,
// libsvm's form exp(-gamma (|a|^2 + |b|^2 - 2 a.b)) with the dot product accumulated in fp64.
__global__ void k_svc_kmat(const float* __restrict__ St, int lds, long long strideS,
                           const int* __restrict__ k_dev, int k_fixed,
                           const int* __restrict__ n_dev, int n_fixed, int kernel,
                           const double* __restrict__ gamma, const int* __restrict__ perm, int ldp,
                           const double* __restrict__ sqn, float* __restrict__ K, int ldk,
                           long long strideK) {
  __shared__ float a[32][33], b[32][33];
  __shared__ int pa[32], pb[32];
  const int f = blockIdx.z;
  const int k = k_dev ? k_dev[f] : k_fixed;
  const int n = n_dev ? n_dev[f] : n_fixed;
  const int i0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  if (t0 < i0 || i0 >= n || t0 >= n) return;
  const float* X = St + (long long)f * strideS;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // ty 0..7
  if (threadIdx.x < 32) pa[tx] = i0 + tx < n ? perm[(long long)f * ldp + i0 + tx] : -1;
  else if (threadIdx.x < 64) pb[tx] = t0 + tx < n ? perm[(long long)f * ldp + t0 + tx] : -1;
  __syncthreads();
  const int ia = pa[tx], ib = pb[tx];
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int j0 = 0; j0 < k; j0 += 32) {
    for (int r = ty; r < 32; r += 8) {
      const int j = j0 + r;
      a[r][tx] = (j < k && ia >= 0) ? X[(long long)j * lds + ia] : 0.f;
      b[r][tx] = (j < k && ib >= 0) ? X[(long long)j * lds + ib] : 0.f;
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
  const double g = gamma[f];
  float* Kf = K + (long long)f * strideK;
  const double sb = ib >= 0 ? sqn[(long long)f * ldp + ib] : 0.0;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int r = ty + 8 * e;
    const int i = i0 + r, t = t0 + tx;
    float v = 0.f;
    if (i < n && t < n) {
      v = kernel == 1 ? (float)exp(-g * (sqn[(long long)f * ldp + pa[r]] + sb - 2.0 * acc[e]))
                      : (float)acc[e];
      Kf[(long long)i * ldk + t] = v;
    }
    a[r][tx] = v;                       // staged for the mirrored tile
  }
  if (t0 == i0) return;
  __syncthreads();
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int r = ty + 8 * e;           // row of the mirrored tile = column of this one
    const int i = t0 + r, t = i0 + tx;
    if (i < n && t < n) Kf[(long long)i * ldk + t] = a[tx][r];
  }
}

struct Pick { double v; int i; };
// larger value wins, ties go to the larger index (libsvm scans upwards with >=)
__device__ __forceinline__ Pick warp_argmax_last(Pick p) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double v = __shfl_xor_sync(0xffffffffu, p.v, o);
    const int i = __shfl_xor_sync(0xffffffffu, p.i, o);
    if (v > p.v || (v == p.v && i > p.i)) { p.v = v; p.i = i; }
  }
  return p;
}
__device__ __forceinline__ double warp_maxd(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_mind(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// pair p -> (a, b), a < b, in libsvm's order (a outer, b inner)
__device__ __forceinline__ void pair_of(int p, int ncls, int& a, int& b) {
  a = 0;
  int rem = p;
  while (rem >= ncls - 1 - a) { rem -= ncls - 1 - a; ++a; }
  b = a + 1 + rem;
}

// Dual problem of one class pair, one warp.  Members: class a (y=+1) then class b (y=-1), both
// in pool order = two contiguous ranges of the class-sorted kernel matrix, so a kernel row is
// two coalesced segments.  Shared memory per warp: alpha[m_max], G[m_max] (double), qi[m_max]
// (row i of the pair's kernel block, kept from the selection pass for the gradient update),
// qd[m_max] (diagonal).
constexpr int SVC_U = 4;                // independent loads in flight per lane
__global__ void __launch_bounds__(32 * SVC_WARPS)
k_svc_smo(const float* __restrict__ K, int ldk, long long strideK, const int* __restrict__ perm,
          int ldp, const int* __restrict__ cls_off, const int* __restrict__ n_dev, int n_fixed,
          int ncls, double Cpar, int balanced, double eps, int max_iter, double* __restrict__ coef,
          int ldc, double* __restrict__ rho, int* __restrict__ info, int m_max, int ntask) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int task = blockIdx.x * SVC_WARPS + w;
  if (task >= ntask) return;
  const int npair = ncls * (ncls - 1) / 2;
  const int f = task / npair, p = task - f * npair;
  int ca, cb;
  pair_of(p, ncls, ca, cb);
  const size_t per_warp = (size_t)m_max * (8 + 8 + 4 + 4) + 16;
  unsigned char* base = smem_raw + (size_t)w * ((per_warp + 15) & ~(size_t)15);
  double* alpha = (double*)base;
  double* G = alpha + m_max;
  float* qi = (float*)(G + m_max);
  float* qd = qi + m_max;

  const int n = n_dev ? n_dev[f] : n_fixed;
  const int* off = cls_off + (long long)f * (ncls + 1);
  const int offA = off[ca], ma = off[ca + 1] - offA;
  const int offB = off[cb], mb = off[cb + 1] - offB;
  const int m = ma + mb;
  const float* Kf = K + (long long)f * strideK;
  int* inf = info + 2 * (long long)task;
  if (ma == 0 || mb == 0) {            // class absent from this pool: no such pair in libsvm
    if (lane == 0) { rho[task] = 0.0; inf[0] = 0; inf[1] = 2; }
    return;
  }
  if (m > m_max) {
    if (lane == 0) { rho[task] = 0.0; inf[0] = 0; inf[1] = 3; }
    return;
  }
  double Cp = Cpar, Cn = Cpar;
  if (balanced) {                       // n_samples / (n_classes_present * count_c), sklearn
    int present = 0;
    for (int c = 0; c < ncls; ++c) present += off[c + 1] > off[c];
    Cp = Cpar * ((double)n / ((double)present * (double)ma));
    Cn = Cpar * ((double)n / ((double)present * (double)mb));
  }
  // sorted position of local member t, and its label sign
#define SVC_POS(t) ((t) < ma ? offA + (t) : offB + (t) - ma)
#define SVC_Y(t) ((t) < ma ? 1 : -1)
  for (int t = lane; t < m; t += 32) {
    alpha[t] = 0.0;
    G[t] = -1.0;
    const int ps = SVC_POS(t);
    qd[t] = Kf[(long long)ps * ldk + ps];
  }
  __syncwarp();

  int it = 0, status = 1;
  while (it < max_iter) {
    // --- i: maximal violator of the "up" set
    Pick pi{-SVC_INF, -1};
    for (int t = lane; t < m; t += 32) {
      const double a = alpha[t], g = G[t];
      if (t < ma) { if (a < Cp && -g >= pi.v) { pi.v = -g; pi.i = t; } }
      else        { if (a > 0.0 && g >= pi.v) { pi.v = g; pi.i = t; } }
    }
    pi = warp_argmax_last(pi);
    const int i = pi.i;
    const double Gmax = pi.v;
    if (i < 0) { status = 0; break; }
    const float* Ki = Kf + (long long)SVC_POS(i) * ldk;
    const int yi = SVC_Y(i);
    const double qdi = (double)qd[i];
    // --- j: second-order choice in the "low" set (row i is gathered once and kept in qi)
    Pick pj{-SVC_INF, -1};               // maximise -obj  (obj <= min, ties to the larger index)
    double Gmax2 = -SVC_INF;
    for (int t0 = lane; t0 < m; t0 += 32 * SVC_U) {
      float kv[SVC_U];
#pragma unroll
      for (int u = 0; u < SVC_U; ++u) {
        const int t = t0 + 32 * u;
        kv[u] = t < m ? Ki[SVC_POS(t)] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < SVC_U; ++u) {
        const int t = t0 + 32 * u;
        if (t >= m) break;
        qi[t] = kv[u];
        const double a = alpha[t], g = G[t];
        double gd;
        bool ok;
        if (t < ma) { ok = a > 0.0; gd = Gmax + g; if (ok) Gmax2 = fmax(Gmax2, g); }
        else        { ok = a < Cn;  gd = Gmax - g; if (ok) Gmax2 = fmax(Gmax2, -g); }
        if (ok && gd > 0.0) {
          // QD_i + QD_t - 2 y_i y_t Q_it with Q_it = y_i y_t K_it
          double quad = qdi + (double)qd[t] - 2.0 * (double)kv[u];
          if (!(quad > 0.0)) quad = SVC_TAU;
          const double nobj = (gd * gd) / quad;
          if (nobj >= pj.v) { pj.v = nobj; pj.i = t; }
        }
      }
    }
    pj = warp_argmax_last(pj);
    Gmax2 = warp_maxd(Gmax2);
    const int j = pj.i;
    if (Gmax + Gmax2 < eps || j < 0) { status = 0; break; }
    ++it;
    __syncwarp();                        // qi complete
    const float* Kj = Kf + (long long)SVC_POS(j) * ldk;
    const int yj = SVC_Y(j);
    const double Ci = yi > 0 ? Cp : Cn, Cj = yj > 0 ? Cp : Cn;
    const double ai0 = alpha[i], aj0 = alpha[j], Gi = G[i], Gj = G[j];
    const double Qij = (double)((float)(yi * yj) * qi[j]);
    double ai = ai0, aj = aj0;
    if (yi != yj) {
      double quad = qdi + (double)qd[j] + 2.0 * Qij;
      if (!(quad > 0.0)) quad = SVC_TAU;
      const double delta = (-Gi - Gj) / quad;
      const double diff = ai - aj;
      ai += delta;
      aj += delta;
      if (diff > 0.0) { if (aj < 0.0) { aj = 0.0; ai = diff; } }
      else            { if (ai < 0.0) { ai = 0.0; aj = -diff; } }
      if (diff > Ci - Cj) { if (ai > Ci) { ai = Ci; aj = Ci - diff; } }
      else                { if (aj > Cj) { aj = Cj; ai = Cj + diff; } }
    } else {
      double quad = qdi + (double)qd[j] - 2.0 * Qij;
      if (!(quad > 0.0)) quad = SVC_TAU;
      const double delta = (Gi - Gj) / quad;
      const double sum = ai + aj;
      ai -= delta;
      aj += delta;
      if (sum > Ci) { if (ai > Ci) { ai = Ci; aj = sum - Ci; } }
      else          { if (aj < 0.0) { aj = 0.0; ai = sum; } }
      if (sum > Cj) { if (aj > Cj) { aj = Cj; ai = sum - Cj; } }
      else          { if (ai < 0.0) { ai = 0.0; aj = sum; } }
    }
    const double dai = ai - ai0, daj = aj - aj0;
    __syncwarp();                        // every lane has read alpha/G of i and j
    for (int t0 = lane; t0 < m; t0 += 32 * SVC_U) {
      float kv[SVC_U];
#pragma unroll
      for (int u = 0; u < SVC_U; ++u) {
        const int t = t0 + 32 * u;
        kv[u] = t < m ? Kj[SVC_POS(t)] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < SVC_U; ++u) {
        const int t = t0 + 32 * u;
        if (t >= m) break;
        const int yt = SVC_Y(t);
        const float qit = (float)(yi * yt) * qi[t];
        const float qjt = (float)(yj * yt) * kv[u];
        G[t] += (double)qit * dai + (double)qjt * daj;
      }
    }
    if (lane == 0) { alpha[i] = ai; alpha[j] = aj; }
    __syncwarp();
  }
  // rho: mean of y G over the free variables, else the midpoint of the bounds
  double ub = SVC_INF, lbv = -SVC_INF, sum_free = 0.0;
  int nfree = 0;
  for (int t = lane; t < m; t += 32) {
    const double a = alpha[t];
    const int yt = SVC_Y(t);
    const double yG = (double)yt * G[t];
    const double Ct = yt > 0 ? Cp : Cn;
    if (a >= Ct)      { if (yt < 0) ub = fmin(ub, yG); else lbv = fmax(lbv, yG); }
    else if (a <= 0.0) { if (yt > 0) ub = fmin(ub, yG); else lbv = fmax(lbv, yG); }
    else { ++nfree; sum_free += yG; }
  }
  nfree = __reduce_add_sync(0xffffffffu, nfree);
  sum_free = warp_sum(sum_free);
  ub = warp_mind(ub);
  lbv = warp_maxd(lbv);
  const double r = nfree > 0 ? sum_free / (double)nfree : 0.5 * (ub + lbv);
  // coefficients in libsvm's sv_coef layout, indexed by pool sample: a class-a member stores
  // the pair under slot b-1, a class-b member under slot a
  double* cf = coef + (long long)f * (ncls - 1) * ldc;
  const int* pf = perm + (long long)f * ldp;
  for (int t = lane; t < m; t += 32) {
    const int slot = t < ma ? cb - 1 : ca;
    cf[(long long)slot * ldc + pf[SVC_POS(t)]] = alpha[t] * (double)SVC_Y(t);
  }
  if (lane == 0) { rho[task] = r; inf[0] = it; inf[1] = status; }
#undef SVC_POS
#undef SVC_Y
}

// One CTA per (fold, held-out trial).  dec (optional): [fold][n_te_max][npair] pair decisions.
__global__ void k_svc_predict(const float* __restrict__ St, int lds, long long strideS,
                              const float* __restrict__ Ste, int ldt, long long strideT,
                              const int* __restrict__ k_dev, int k_fixed,
                              const int* __restrict__ n_dev, int n_fixed,
                              const int* __restrict__ n_te, int n_te_max, const int* __restrict__ y,
                              int ldy, const int* __restrict__ classes, int ncls, int kernel,
                              const double* __restrict__ gamma, const double* __restrict__ coef,
                              int ldc, const double* __restrict__ rho, int* __restrict__ yhat,
                              double* __restrict__ dec) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int votes[64], counts[64];
  __shared__ float z[1024];
  const int f = blockIdx.y, s = blockIdx.x;
  const int nt = n_te ? n_te[f] : n_te_max;
  if (s >= nt) {
    if (threadIdx.x == 0) yhat[(long long)f * n_te_max + s] = -1;
    return;
  }
  const int k = k_dev ? k_dev[f] : k_fixed;
  const int n = n_dev ? n_dev[f] : n_fixed;
  double* kv = (double*)smem_raw;
  const float* X = St + (long long)f * strideS;
  const float* Z = Ste + (long long)f * strideT;
  const int* yf = y + (long long)f * ldy;
  for (int j = threadIdx.x; j < k; j += blockDim.x) z[j] = Z[(long long)j * ldt + s];
  if (threadIdx.x < 64) { votes[threadIdx.x] = 0; counts[threadIdx.x] = 0; }
  __syncthreads();
  const double g = gamma[f];
  for (int t = threadIdx.x; t < n; t += blockDim.x) {
    double acc = 0.0;
    for (int j = 0; j < k; ++j) {
      const double xv = (double)X[(long long)j * lds + t], zv = (double)z[j];
      if (kernel == 1) { const double d = xv - zv; acc = fma(d, d, acc); }
      else acc = fma(xv, zv, acc);
    }
    kv[t] = kernel == 1 ? exp(-g * acc) : acc;
    const int lab = yf[t];
    for (int c = 0; c < ncls; ++c)
      if (classes[c] == lab) atomicAdd(&counts[c], 1);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int npair = ncls * (ncls - 1) / 2;
  const double* cf = coef + (long long)f * (ncls - 1) * ldc;
  for (int p = w; p < npair; p += nw) {
    int ca, cb;
    pair_of(p, ncls, ca, cb);
    if (counts[ca] == 0 || counts[cb] == 0) {
      if (dec && lane == 0) dec[((long long)f * n_te_max + s) * npair + p] = 0.0;
      continue;
    }
    const int la = classes[ca], lb = classes[cb];
    double sum = 0.0;
    for (int t = lane; t < n; t += 32) {
      const int lab = yf[t];
      if (lab == la) sum = fma(cf[(long long)(cb - 1) * ldc + t], kv[t], sum);
      else if (lab == lb) sum = fma(cf[(long long)ca * ldc + t], kv[t], sum);
    }
    sum = warp_sum(sum) - rho[(long long)f * npair + p];
    if (lane == 0) {
      if (dec) dec[((long long)f * n_te_max + s) * npair + p] = sum;
      atomicAdd(&votes[sum > 0.0 ? ca : cb], 1);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int best = -1, bv = -1;
    for (int c = 0; c < ncls; ++c)
      if (counts[c] > 0 && votes[c] > bv) { bv = votes[c]; best = c; }
    yhat[(long long)f * n_te_max + s] = best >= 0 ? classes[best] : -1;
  }
}

static size_t smo_smem(int m_max) {
  const size_t per_warp = (size_t)m_max * (8 + 8 + 4 + 4) + 16;
  return SVC_WARPS * ((per_warp + 15) & ~(size_t)15);
}

}  // namespace

extern "C" int cpsd_svc_kernel_matrix(const float* St, int lds, long long strideS, const int* k_dev,
                                      int k_fixed, const int* n_dev, int n_fixed, int n_max, const int* y,
                                      int ldy, const int* classes, int ncls, int kernel, double gamma,
                                      double* gamma_out, int* perm, int* cls_off, double* sqn, float* K,
                                      int ldk, long long strideK, int nfold, cudaStream_t stream) {
  CPSD_CHECK_ARG(nfold >= 0 && n_max > 0 && ldk >= n_max && (kernel == 0 || kernel == 1) && ncls >= 2,
                 "svc_kernel_matrix: bad dims");
  if (nfold == 0) return CPSD_OK;
  k_svc_gamma<<<nfold, 256, 0, stream>>>(St, lds, strideS, k_dev, k_fixed, n_dev, n_fixed, kernel, gamma,
                                         gamma_out);
  CPSD_LAUNCH_CHECK();
  k_svc_perm<<<nfold, 256, 0, stream>>>(St, lds, strideS, k_dev, k_fixed, n_dev, n_fixed, y, ldy, classes,
                                        ncls, perm, ldk, cls_off, sqn);
  CPSD_LAUNCH_CHECK();
  const int tiles = (n_max + 31) / 32;
  k_svc_kmat<<<dim3(tiles, tiles, nfold), 256, 0, stream>>>(St, lds, strideS, k_dev, k_fixed, n_dev, n_fixed,
                                                             kernel, gamma_out, perm, ldk, sqn, K, ldk,
                                                             strideK);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

extern "C" int cpsd_svc_fit_ovo(const float* K, int ldk, long long strideK, const int* perm,
                                const int* cls_off, const int* n_dev, int n_fixed, int ncls, double C,
                                int balanced, double eps, int max_iter, double* coef, int ldc,
                                double* rho, int* info, int m_max, int nfold, cudaStream_t stream) {
  CPSD_CHECK_ARG(nfold >= 0 && ncls >= 2 && ncls <= 64 && m_max > 0 && C > 0 && eps > 0 && max_iter > 0,
                 "svc_fit_ovo: bad arguments");
  if (nfold == 0) return CPSD_OK;
  const size_t smem = smo_smem(m_max);
  CPSD_CHECK_ARG(smem <= 227 * 1024, "svc_fit_ovo: pair size exceeds the shared-memory budget");
  CPSD_CUDA(cudaFuncSetAttribute(k_svc_smo, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CPSD_CUDA(cudaMemsetAsync(coef, 0, sizeof(double) * (size_t)nfold * (ncls - 1) * ldc, stream));
  const int ntask = nfold * (ncls * (ncls - 1) / 2);
  k_svc_smo<<<(ntask + SVC_WARPS - 1) / SVC_WARPS, 32 * SVC_WARPS, smem, stream>>>(
      K, ldk, strideK, perm, ldk, cls_off, n_dev, n_fixed, ncls, C, balanced, eps, max_iter, coef, ldc,
      rho, info, m_max, ntask);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

extern "C" int cpsd_svc_predict_ovo(const float* St, int lds, long long strideS, const float* Ste, int ldt,
                                    long long strideT, const int* k_dev, int k_fixed, const int* n_dev,
                                    int n_fixed, int n_max, const int* n_te, int n_te_max, const int* y,
                                    int ldy, const int* classes, int ncls, int kernel,
                                    const double* gamma, const double* coef, int ldc, const double* rho,
                                    int* yhat, double* dec, int k_max, int nfold, cudaStream_t stream) {
  CPSD_CHECK_ARG(nfold >= 0 && ncls >= 2 && ncls <= 64 && n_te_max > 0 && n_max > 0 && k_max <= 1024,
                 "svc_predict_ovo: bad arguments (at most 64 classes, 1024 features)");
  if (nfold == 0) return CPSD_OK;
  const size_t smem = sizeof(double) * (size_t)n_max;
  CPSD_CHECK_ARG(smem <= 200 * 1024, "svc_predict_ovo: pool too large");
  CPSD_CUDA(cudaFuncSetAttribute(k_svc_predict, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_svc_predict<<<dim3(n_te_max, nfold), 256, smem, stream>>>(St, lds, strideS, Ste, ldt, strideT, k_dev,
                                                              k_fixed, n_dev, n_fixed, n_te, n_te_max, y,
                                                              ldy, classes, ncls, kernel, gamma, coef, ldc,
                                                              rho, yhat, dec);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

// ---------------------------------------------------------------------------------------
// Bagging (sklearn BaggingClassifier(estimator=SVC(kernel='linear'), n_estimators=10), the
// decoder of the reference's scripts/aligned_decode_svm.py:262-265): every fold becomes
// n_est bootstrap problems that run as extra folds of the C-SVC kernels above.
//   k_bag_gather: problem p = fold * n_est + e gets the resampled training scores
//                 St_b[p][j][t] = St[fold][j][idx[p][t]], labels y_b[p][t] = y[fold][idx[p][t]],
//                 a copy of the fold's test scores and of its sizes.
//   k_bag_vote  : label = classes[argmax_c #{e : estimator e voted c}] (first maximum, as
//                 numpy argmax over BaggingClassifier.predict_proba's vote counts).
// ---------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256)
k_bag_gather(const float* __restrict__ St, int lds, long long strideS, const float* __restrict__ Ste,
             int ldt, long long strideT, const int* __restrict__ y, int ldy,
             const int* __restrict__ idx, int ldi, const int* __restrict__ k_dev,
             const int* __restrict__ n_dev, const int* __restrict__ nte_dev, int n_est, int kb,
             float* __restrict__ St_b, int lds_b, float* __restrict__ Ste_b, int ldt_b,
             int* __restrict__ y_b, int* __restrict__ k_b, int* __restrict__ n_b,
             int* __restrict__ nte_b) {
  const int p = blockIdx.x, f = p / n_est;
  const int k = min(k_dev[f], kb), n = n_dev[f], nte = nte_dev[f];
  const int* ix = idx + (long long)p * ldi;
  const float* S = St + (long long)f * strideS;
  float* Sb = St_b + (long long)p * kb * lds_b;
  for (int e = threadIdx.x; e < kb * lds_b; e += blockDim.x) {
    const int j = e / lds_b, t = e - j * lds_b;
    Sb[e] = (j < k && t < n) ? S[(long long)j * lds + ix[t]] : 0.f;
  }
  const int* yf = y + (long long)f * ldy;
  for (int t = threadIdx.x; t < lds_b; t += blockDim.x) y_b[(long long)p * lds_b + t] = (t < n) ? yf[ix[t]] : 0;
  const float* Z = Ste + (long long)f * strideT;
  float* Zb = Ste_b + (long long)p * kb * ldt_b;
  for (int e = threadIdx.x; e < kb * ldt_b; e += blockDim.x) {
    const int j = e / ldt_b, s = e - j * ldt_b;
    Zb[e] = (j < k && s < nte) ? Z[(long long)j * ldt + s] : 0.f;
  }
  if (threadIdx.x == 0) { k_b[p] = k; n_b[p] = n; nte_b[p] = nte; }
}

__global__ void k_bag_vote(const int* __restrict__ yhat_b, int n_est, const int* __restrict__ classes,
                           int ncls, const int* __restrict__ nte_dev, int n_te_max,
                           int* __restrict__ yhat) {
  const int f = blockIdx.x;
  const int nte = nte_dev[f];
  for (int s = threadIdx.x; s < n_te_max; s += blockDim.x) {
    if (s >= nte) { yhat[(long long)f * n_te_max + s] = -1; continue; }
    int best = -1, bv = -1;
    for (int c = 0; c < ncls; ++c) {
      int v = 0;
      for (int e = 0; e < n_est; ++e)
        v += (yhat_b[((long long)f * n_est + e) * n_te_max + s] == classes[c]);
      if (v > bv) { bv = v; best = c; }
    }
    yhat[(long long)f * n_te_max + s] = classes[best];
  }
}
}  // namespace

extern "C" int cpsd_bag_gather(const float* St, int lds, long long strideS, const float* Ste, int ldt,
                               long long strideT, const int* y, int ldy, const int* idx, int ldi,
                               const int* k_dev, const int* n_dev, const int* nte_dev, int n_est, int kb,
                               float* St_b, int lds_b, float* Ste_b, int ldt_b, int* y_b, int* k_b,
                               int* n_b, int* nte_b, int nfold, cudaStream_t stream) {
  CPSD_CHECK_ARG(nfold >= 0 && n_est > 0 && kb > 0 && lds_b > 0 && ldt_b > 0 && ldi >= 1,
                 "bag_gather: bad dims");
  if (nfold == 0) return CPSD_OK;
  k_bag_gather<<<nfold * n_est, 256, 0, stream>>>(St, lds, strideS, Ste, ldt, strideT, y, ldy, idx, ldi,
                                                  k_dev, n_dev, nte_dev, n_est, kb, St_b, lds_b, Ste_b,
                                                  ldt_b, y_b, k_b, n_b, nte_b);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

extern "C" int cpsd_bag_vote(const int* yhat_b, int n_est, const int* classes, int ncls,
                             const int* nte_dev, int n_te_max, int* yhat, int nfold,
                             cudaStream_t stream) {
  CPSD_CHECK_ARG(nfold >= 0 && n_est > 0 && ncls > 0 && n_te_max > 0, "bag_vote: bad dims");
  if (nfold == 0) return CPSD_OK;
  k_bag_vote<<<nfold, 128, 0, stream>>>(yhat_b, n_est, classes, ncls, nte_dev, n_te_max, yhat);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}
