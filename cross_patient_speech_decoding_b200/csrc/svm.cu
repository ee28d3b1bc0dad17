// Batched one-vs-rest linear SVM (L2-regularised, squared hinge, regularised bias) --
// the objective liblinear solves for sklearn's LinearSVC(loss='squared_hinge',
// fit_intercept=True, intercept_scaling=1), which is the decoder injected into the
// reference's crossPtDecoder classes (decoders/cross_pt_decoders.py:46-59) for this path:
//
//      min_w  1/2 |w|^2 + C sum_i max(0, 1 - y_i w.x~_i)^2 ,   x~_i = [x_i, 1]
//
// One CTA per (fold, class) sub-problem, all arithmetic in fp64, features read from a
// feature-major fp32 score matrix (k x n) so both X v and X^T u stream coalesced.
//
//  phase 1: dual coordinate descent (Hsieh et al. 2008 Alg. 3, the update rule of
//           liblinear's solve_l2r_l1l2_svc for the L2 loss: D_ii = 1/(2C), U = inf) with a
//           fresh random permutation per epoch and liblinear's stopping rule
//           PGmax - PGmin <= eps.
//  phase 2: finite Newton (generalised Hessian I + 2C X_I^T X_I on the active set I,
//           Jacobi-preconditioned CG, exact line search on the piecewise quadratic) on
//           the same objective, started from the DCD iterate w = sum_i alpha_i y_i x~_i.
//           The dual is so ill-conditioned on unscaled PCA scores that coordinate descent
//           alone does not reach the optimum in 1e6 epochs (measured on the reference's
//           own LinearSVC); phase 2 guarantees the unique optimum.
#include "common.cuh"
#include "descs.h"

namespace {

#define SV_NT 256
#ifndef SV_MINB
#define SV_MINB 5
#endif
#define SV_INEXACT_ITS 12

struct SvmSmem {
  double* w; double* g; double* d; double* r; double* zz; double* p; double* hp; double* dg;
  double* z; double* q; double* alpha; double* qd;
  float* yv; int* perm; double* red;
};

// dcd == 0: no task of the launch runs the dual-CD phase, so its per-sample arrays (alpha, QD,
// permutation: 20 bytes per sample) are not carved -- 32 KB instead of 50 KB per task at
// n = 1152, which is what lets five tasks share an SM.
__device__ __forceinline__ SvmSmem carve(unsigned char* base, int kp_max, int n_max, int dcd) {
  SvmSmem s;
  double* dp = reinterpret_cast<double*>(base);
  s.red = dp; dp += 40;
  s.w = dp; dp += kp_max; s.g = dp; dp += kp_max; s.d = dp; dp += kp_max; s.r = dp; dp += kp_max;
  s.zz = dp; dp += kp_max; s.p = dp; dp += kp_max; s.hp = dp; dp += kp_max; s.dg = dp; dp += kp_max;
  s.z = dp; dp += n_max; s.q = dp; dp += n_max;
  s.alpha = dp; s.qd = dp;
  if (dcd) { dp += n_max; s.qd = dp; dp += n_max; }
  s.yv = reinterpret_cast<float*>(dp);
  s.perm = reinterpret_cast<int*>(s.yv + n_max);
  return s;
}

// out_i = sum_j St[j][i] v[j] + v[k]   (thread per sample)
__device__ __forceinline__ void xmul(const float* __restrict__ St, int lds, int n, int k,
                                     const double* v, double* out) {
  for (int i = threadIdx.x; i < n; i += SV_NT) {
    double a = v[k];
    for (int j = 0; j < k; ++j) a = fma((double)St[(long long)j * lds + i], v[j], a);
    out[i] = a;
  }
}

// out_j = sum_i St[j][i] u[i]  (warp per feature), out_k = sum_i u[i]
__device__ __forceinline__ void xtmul(const float* __restrict__ St, int lds, int n, int k,
                                      const double* u, double* out) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = SV_NT >> 5;
  for (int j = wid; j <= k; j += nw) {
    double a = 0.0;
    if (j < k) {
      const float* row = St + (long long)j * lds;
      for (int i = lane; i < n; i += 32) a = fma((double)row[i], u[i], a);
    } else {
      for (int i = lane; i < n; i += 32) a += u[i];
    }
    a = warp_sum(a);
    if (lane == 0) out[j] = a;
  }
}

__device__ __forceinline__ double bdot(const double* a, const double* b, int n, double* red) {
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += SV_NT) s = fma(a[i], b[i], s);
  return block_sum(s, red);
}

__global__ void __launch_bounds__(SV_NT, SV_MINB)
k_svm_fit(const cpsd_svm_desc* __restrict__ descs, int kp_max, int n_max, int dcd) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cpsd_svm_desc t = descs[blockIdx.x];
  if (!dcd) t.dcd_epochs = 0;                 // (the host checked: no task asked for the phase)
  SvmSmem s = carve(smem_raw, kp_max, n_max, dcd);
  const int n = min(t.n, n_max);
  int k = t.k_dev ? t.k_dev[0] : t.k;
  if (k > kp_max - 1) k = kp_max - 1;
  if (k < 0) k = 0;
  const int kp = k + 1;
  const float* St = t.St;
  const int lds = t.lds;
  const double C = t.C;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;

  for (int i = threadIdx.x; i < n; i += SV_NT) {
    s.yv[i] = (t.y[i] == t.cls) ? 1.f : -1.f;
    if (dcd) s.alpha[i] = 0.0;
  }
  for (int j = threadIdx.x; j < kp; j += SV_NT) s.w[j] = 0.0;
  __syncthreads();

  int dcd_done = 0, dcd_conv = 0, status = 0;
  // ----------------------------------------------------------------- phase 1: dual CD
  if (t.dcd_epochs > 0) {
    // QD_i = |x~_i|^2 + 1/(2C)
    for (int i = threadIdx.x; i < n; i += SV_NT) {
      double a = 1.0;
      for (int j = 0; j < k; ++j) {
        const double x = (double)St[(long long)j * lds + i];
        a = fma(x, x, a);
      }
      s.qd[i] = a + 0.5 / C;
      s.perm[i] = i;
    }
    __syncthreads();
    unsigned int rng = 0x9E3779B9u ^ (unsigned int)(blockIdx.x * 2654435761u);
    if (rng == 0u) rng = 1u;
    const double diag = 0.5 / C;
    for (int ep = 0; ep < t.dcd_epochs; ++ep) {
      if (wid == 0) {
        if (lane == 0) {
          for (int i = 0; i < n; ++i) {  // Fisher-Yates like liblinear's swap loop
            rng ^= rng << 13; rng ^= rng >> 17; rng ^= rng << 5;
            const int j = i + (int)(rng % (unsigned int)(n - i));
            const int tmp = s.perm[i]; s.perm[i] = s.perm[j]; s.perm[j] = tmp;
          }
        }
        __syncwarp();
        double pgmax = -1e300, pgmin = 1e300;
        for (int it = 0; it < n; ++it) {
          const int i = s.perm[it];
          const double yi = (double)s.yv[i];
          double dot = 0.0;
          for (int j = lane; j < k; j += 32) dot = fma((double)St[(long long)j * lds + i], s.w[j], dot);
          dot = warp_sum(dot) + s.w[k];
          const double ai = s.alpha[i];
          const double G = yi * dot - 1.0 + diag * ai;
          double PG = G;
          if (ai == 0.0 && G > 0.0) PG = 0.0;   // projected gradient at the lower bound
          pgmax = fmax(pgmax, PG);
          pgmin = fmin(pgmin, PG);
          if (fabs(PG) > 1e-12) {
            const double an = fmax(ai - G / s.qd[i], 0.0);
            const double dlt = (an - ai) * yi;
            __syncwarp();
            for (int j = lane; j < k; j += 32) s.w[j] = fma(dlt, (double)St[(long long)j * lds + i], s.w[j]);
            if (lane == 0) { s.w[k] += dlt; s.alpha[i] = an; }
            __syncwarp();
          }
        }
        if (lane == 0) s.red[36] = pgmax - pgmin;
      }
      __syncthreads();
      dcd_done = ep + 1;
      const double gap = s.red[36];
      __syncthreads();
      if (gap <= t.tol_dcd) { dcd_conv = 1; break; }
    }
  }

  // ----------------------------------------------------------------- phase 2: Newton
  int newton_its = 0, cg_total = 0;
  double g0 = -1.0;
  for (int it = 0; it <= t.max_newton; ++it) {
    // margins z_i = y_i x~_i.w ; u_i = y_i max(0, 1 - z_i)
    xmul(St, lds, n, k, s.w, s.z);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += SV_NT) {
      const double zi = (double)s.yv[i] * s.z[i];
      s.z[i] = zi;
      s.q[i] = (zi < 1.0) ? (double)s.yv[i] * (1.0 - zi) : 0.0;
    }
    __syncthreads();
    xtmul(St, lds, n, k, s.q, s.g);
    __syncthreads();
    double gmax = 0.0;
    for (int j = threadIdx.x; j < kp; j += SV_NT) {
      const double gj = s.w[j] - 2.0 * C * s.g[j];
      s.g[j] = gj;
      gmax = fmax(gmax, fabs(gj));
    }
    // block max via sum trick is wrong; do a proper max reduce through red[]
    {
      double m = gmax;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
      __syncthreads();
      if (lane == 0) s.red[wid] = m;
      __syncthreads();
      if (threadIdx.x == 0) {
        double mm = 0.0;
        for (int u = 0; u < SV_NT / 32; ++u) mm = fmax(mm, s.red[u]);
        s.red[37] = mm;
      }
      __syncthreads();
      gmax = s.red[37];
      __syncthreads();
    }
    if (t.max_newton == 0) { status = (t.dcd_epochs > 0 && !dcd_conv) ? 1 : 0; break; }
    if (g0 < 0.0) {
      // reference scale: gradient at w = 0 is -2C X^T y, bounded below by 1
      g0 = fmax(1.0, gmax);
      if (t.dcd_epochs > 0) {
        // recompute the w=0 gradient scale so the stopping rule does not depend on phase 1
        for (int i = threadIdx.x; i < n; i += SV_NT) s.q[i] = (double)s.yv[i];
        __syncthreads();
        xtmul(St, lds, n, k, s.q, s.hp);
        __syncthreads();
        double m = 0.0;
        for (int j = 0; j < kp; ++j) m = fmax(m, fabs(2.0 * C * s.hp[j]));
        g0 = fmax(1.0, m);
        __syncthreads();
      }
    }
    if (gmax <= t.tol_newton * g0) { status = 0; break; }
    if (it == t.max_newton) { status = 1; break; }
    ++newton_its;

    // Jacobi preconditioner on the active set
    {
      const int nw = SV_NT >> 5;
      for (int j = wid; j <= k; j += nw) {
        double a = 0.0;
        if (j < k) {
          const float* row = St + (long long)j * lds;
          for (int i = lane; i < n; i += 32)
            if (s.z[i] < 1.0) { const double x = (double)row[i]; a = fma(x, x, a); }
        } else {
          for (int i = lane; i < n; i += 32) if (s.z[i] < 1.0) a += 1.0;
        }
        a = warp_sum(a);
        if (lane == 0) s.dg[j] = 1.0 + 2.0 * C * a;
      }
    }
    for (int j = threadIdx.x; j < kp; j += SV_NT) { s.d[j] = 0.0; s.r[j] = -s.g[j]; }
    __syncthreads();
    for (int j = threadIdx.x; j < kp; j += SV_NT) { s.zz[j] = s.r[j] / s.dg[j]; s.p[j] = s.zz[j]; }
    __syncthreads();
    double rz = bdot(s.r, s.zz, kp, s.red);
    const double r0 = sqrt(bdot(s.r, s.r, kp, s.red));
    const int cg_max = 4 * kp + 20;
    // inexact Newton: the linear system is solved to a relative residual that tightens with
    // the gradient (eta = min(0.1, |g| / |g0|), quadratic forcing), down to 1e-10
    // -- for the first SV_INEXACT_ITS steps.  On separable, ill-conditioned pools (noisy latents,
    // score variances spanning 1e2) the active set keeps changing by one or two samples per step
    // and inexact steps zig-zag without ever reaching the optimum (measured: |g|/|g0| stuck at
    // 1e-3 after 60 steps; liblinear's own trust-region Newton stalls on the same problems);
    // the finite Newton method terminates only with exact steps, so from then on the system is
    // solved to 1e-10 (60-110 steps, ~1e4 CG iterations on those problems).
    const double eta = (it < SV_INEXACT_ITS) ? fmax(1e-10, fmin(0.1, gmax / g0)) : 1e-10;
    for (int cg = 0; cg < cg_max; ++cg) {
      xmul(St, lds, n, k, s.p, s.q);
      __syncthreads();
      for (int i = threadIdx.x; i < n; i += SV_NT) if (!(s.z[i] < 1.0)) s.q[i] = 0.0;
      __syncthreads();
      xtmul(St, lds, n, k, s.q, s.hp);
      __syncthreads();
      for (int j = threadIdx.x; j < kp; j += SV_NT) s.hp[j] = s.p[j] + 2.0 * C * s.hp[j];
      __syncthreads();
      const double php = bdot(s.p, s.hp, kp, s.red);
      const double al = rz / php;
      for (int j = threadIdx.x; j < kp; j += SV_NT) {
        s.d[j] = fma(al, s.p[j], s.d[j]);
        s.r[j] = fma(-al, s.hp[j], s.r[j]);
      }
      __syncthreads();
      ++cg_total;
      const double rn = sqrt(bdot(s.r, s.r, kp, s.red));
      if (rn <= eta * r0) break;
      for (int j = threadIdx.x; j < kp; j += SV_NT) s.zz[j] = s.r[j] / s.dg[j];
      __syncthreads();
      const double rzn = bdot(s.r, s.zz, kp, s.red);
      const double beta = rzn / rz;
      rz = rzn;
      for (int j = threadIdx.x; j < kp; j += SV_NT) s.p[j] = fma(beta, s.p[j], s.zz[j]);
      __syncthreads();
    }
    // exact line search along d: phi'(t) = w.d + t d.d - 2C sum_{m_i(t)>0} m_i(t) q_i,
    // m_i(t) = 1 - z_i - t q_i, q_i = y_i x~_i.d   (Newton on the piecewise-linear phi')
    xmul(St, lds, n, k, s.d, s.q);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += SV_NT) s.q[i] *= (double)s.yv[i];
    __syncthreads();
    const double wd = bdot(s.w, s.d, kp, s.red);
    const double dd = bdot(s.d, s.d, kp, s.red);
    const double gd = fabs(bdot(s.g, s.d, kp, s.red)) + 1e-300;   // |phi'(0)|
    double tstep = 1.0, lo = 0.0, hi = -1.0;                      // hi < 0: no upper bracket yet
    for (int ls = 0; ls < 80; ++ls) {
      double a1 = 0.0, a2 = 0.0;
      for (int i = threadIdx.x; i < n; i += SV_NT) {
        const double qi = s.q[i];
        const double mi = 1.0 - s.z[i] - tstep * qi;
        if (mi > 0.0) { a1 = fma(mi, qi, a1); a2 = fma(qi, qi, a2); }
      }
      a1 = block_sum(a1, s.red);
      a2 = block_sum(a2, s.red);
      const double dphi = wd + tstep * dd - 2.0 * C * a1;
      const double ddphi = dd + 2.0 * C * a2;
      if (fabs(dphi) <= 1e-13 * gd) break;
      if (dphi < 0.0) lo = tstep; else hi = tstep;
      double tn = tstep - dphi / ddphi;
      if (!(tn > lo) || (hi > 0.0 && !(tn < hi))) tn = (hi > 0.0) ? 0.5 * (lo + hi) : 2.0 * tstep;
      if (fabs(tn - tstep) <= 1e-15 * fabs(tstep)) { tstep = tn; break; }
      tstep = tn;
    }
    for (int j = threadIdx.x; j < kp; j += SV_NT) s.w[j] = fma(tstep, s.d[j], s.w[j]);
    __syncthreads();
  }

  for (int j = threadIdx.x; j < kp; j += SV_NT) t.w[j] = s.w[j];
  for (int j = kp + threadIdx.x; j < kp_max; j += SV_NT) t.w[j] = 0.0;
  if (threadIdx.x == 0 && t.info) {
    t.info[0] = newton_its;
    t.info[1] = cg_total;
    t.info[2] = dcd_done;
    t.info[3] = status;
  }
}

// One-vs-rest scoring: label = classes[argmax_c w_c . [x, 1]] (first maximum wins, as
// numpy argmax in sklearn LinearClassifierMixin.predict).
// Xt: feature-major test scores (k x n_te) per fold; W: (ncls x ldw) per fold, bias at k.
__global__ void k_svm_predict(const float* __restrict__ Xt, int ldx, long long strideX,
                              const double* __restrict__ W, int ldw, long long strideW,
                              const int* __restrict__ k_dev, int k_fixed,
                              const int* __restrict__ n_te, int n_te_max,
                              const int* __restrict__ classes, int ncls, int* __restrict__ yhat,
                              double* __restrict__ dec, int nfold) {
  const int f = blockIdx.x;
  if (f >= nfold) return;
  const int k = k_dev ? k_dev[f] : k_fixed;
  const int nt = n_te ? n_te[f] : n_te_max;
  const float* X = Xt + (long long)f * strideX;
  const double* Wf = W + (long long)f * strideW;
  for (int t = threadIdx.x; t < n_te_max; t += blockDim.x) {
    if (t >= nt) { yhat[(long long)f * n_te_max + t] = -1; continue; }
    int best = 0;
    double bestv = -1e300;
    for (int c = 0; c < ncls; ++c) {
      const double* w = Wf + (long long)c * ldw;
      double a = w[k];
      for (int j = 0; j < k; ++j) a = fma((double)X[(long long)j * ldx + t], w[j], a);
      if (dec) dec[((long long)f * n_te_max + t) * ncls + c] = a;
      if (a > bestv) { bestv = a; best = c; }
    }
    yhat[(long long)f * n_te_max + t] = classes[best];
  }
}

}  // namespace

static size_t svm_smem_bytes(int kp_max, int n_max, int dcd) {
  return (40 + 8 * (size_t)kp_max + (dcd ? 4 : 2) * (size_t)n_max) * sizeof(double) +
         (dcd ? 2 : 1) * (size_t)n_max * 4 + 16;
}

// dcd_epochs_max: the largest dcd_epochs of any task of the launch (0: Newton only -- the
// dual-CD work arrays are left out of shared memory and more tasks share an SM).
extern "C" int cpsd_svm_fit_ovr_ex(const cpsd_svm_desc* descs_dev, int ntask, int k_max, int n_max,
                                   int dcd_epochs_max, cudaStream_t stream) {
  CPSD_CHECK_ARG(ntask >= 0 && k_max >= 0 && n_max > 0, "svm_fit_ovr: bad dims");
  if (ntask == 0) return CPSD_OK;
  const int dcd = dcd_epochs_max > 0;
  const size_t smem = svm_smem_bytes(k_max + 1, n_max, dcd);
  CPSD_CHECK_ARG(smem <= 227 * 1024, "svm_fit_ovr: n_max/k_max exceed the shared-memory budget");
  CPSD_CUDA(cudaFuncSetAttribute(k_svm_fit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_svm_fit<<<ntask, SV_NT, smem, stream>>>(descs_dev, k_max + 1, n_max, dcd);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

extern "C" int cpsd_svm_fit_ovr(const cpsd_svm_desc* descs_dev, int ntask, int k_max, int n_max,
                                cudaStream_t stream) {
  return cpsd_svm_fit_ovr_ex(descs_dev, ntask, k_max, n_max, 1, stream);
}

extern "C" int cpsd_svm_predict_ovr(const float* Xt, int ldx, long long strideX, const double* W,
                                    int ldw, long long strideW, const int* k_dev, int k_fixed,
                                    const int* n_te, int n_te_max, const int* classes, int ncls,
                                    int* yhat, double* dec, int nfold, cudaStream_t stream) {
  CPSD_CHECK_ARG(nfold >= 0 && ncls > 0 && n_te_max > 0, "svm_predict_ovr: bad dims");
  if (nfold == 0) return CPSD_OK;
  k_svm_predict<<<nfold, 128, 0, stream>>>(Xt, ldx, strideX, W, ldw, strideW, k_dev, k_fixed, n_te,
                                           n_te_max, classes, ncls, yhat, dec, nfold);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}
