// Fused per-trial predict of a fitted cross-patient decoder (BASELINE config 5: "per-trial
// aligned projection + SVM decode latency at batch 1 and batch 256"): what
// crossPtDecoder.predict does through three objects -- aligner.transform(X, idx=0)
// (decoders/cross_pt_decoders.py:444, AlignMCCA.py:110: (X - mu) L), the flatten + PCA transform
// of DimRedReshape (decomposition/DimRedReshape.py:52-65) and the one-vs-rest linear decision
// (sklearn LinearClassifierMixin.predict) -- as ONE kernel, one CTA per trial, no intermediate
// leaving the SM:
//   1. the trial (T x C, host float64) is centred and staged in shared memory as fp32
//   2. z = flatten((X - mu) A)            (T*Q values, shared memory)
//   3. s = (z - m) P                      (F x k2 matrix P streamed from L2, fp64 accumulation)
//   4. label = classes[argmax_c  w_c . [s, 1]]
#include "common.cuh"

namespace {

constexpr int PF_NT = 256;

// grid (n trials, nsplit time slices): a CTA handles the rows [t0, t1) of its trial, adds its
// partial scores to the workspace, and the last CTA of a trial to arrive (ticket counter) sums
// the partials in slice order (deterministic) and decides.
__global__ void __launch_bounds__(PF_NT)
k_predict_fused(const double* __restrict__ X, int T, int C, const float* __restrict__ mu,
                const float* __restrict__ A, int Q, const float* __restrict__ pmean,
                const float* __restrict__ P, int k2, const double* __restrict__ W,
                const int* __restrict__ classes, int ncls, int* __restrict__ yhat,
                double* __restrict__ dec, int rows_per, double* __restrict__ ws_part,
                int* __restrict__ ws_count, float* __restrict__ scores_out, int ld_scores) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_last;
  const int trial = blockIdx.x, slice = blockIdx.y, nsplit = gridDim.y;
  const int t0 = slice * rows_per;
  const int t1 = min(T, t0 + rows_per);
  const int nt = max(t1 - t0, 0);
  float* xs = reinterpret_cast<float*>(smem_raw);                     // rows_per x C, centred
  float* z = xs + (((size_t)rows_per * C + 1) & ~(size_t)1);          // rows_per x Q
  double* part = reinterpret_cast<double*>(z + (((size_t)rows_per * Q + 1) & ~(size_t)1));  // [nw][k2p]
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = PF_NT >> 5;
  const double* Xt = X + ((long long)trial * T + t0) * C;
  for (int e = threadIdx.x; e < nt * C; e += PF_NT) {
    const int c = e % C;
    xs[e] = (float)(Xt[e] - (double)(mu ? mu[c] : 0.f));
  }
  __syncthreads();
  const int Fs = nt * Q;
  const long long f0 = (long long)t0 * Q;
  for (int e = threadIdx.x; e < Fs; e += PF_NT) {
    const int t = e / Q, q = e - t * Q;
    const float* xr = xs + (size_t)t * C;
    float acc = 0.f;
    for (int c = 0; c < C; ++c) acc = fmaf(xr[c], A[(long long)c * Q + q], acc);
    z[e] = acc - pmean[f0 + e];
  }
  __syncthreads();
  // s_j = sum_f z_f P[f][j]: warps split f, lanes split j (coalesced rows of P)
  const int k2p = (k2 + 31) & ~31;
  for (int j0 = 0; j0 < k2p; j0 += 32) {
    const int j = j0 + lane;
    double acc = 0.0;
    if (j < k2)
      for (int f = wid; f < Fs; f += nw)
        acc = fma((double)z[f], (double)P[(f0 + f) * k2 + j], acc);
    part[wid * k2p + j] = acc;
  }
  __syncthreads();
  double* gp = ws_part + ((long long)trial * nsplit + slice) * k2p;
  for (int j = threadIdx.x; j < k2; j += PF_NT) {
    double a = 0.0;
    for (int w = 0; w < nw; ++w) a += part[w * k2p + j];
    gp[j] = a;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&ws_count[trial], 1) == nsplit - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double* s = part;                                         // k2 scores
  const volatile double* gall = ws_part + (long long)trial * nsplit * k2p;
  for (int j = threadIdx.x; j < k2; j += PF_NT) {
    double a = 0.0;
    for (int u = 0; u < nsplit; ++u) a += gall[(long long)u * k2p + j];
    s[j] = a;
    if (scores_out) scores_out[(long long)j * ld_scores + trial] = (float)a;   // feature-major
  }
  __syncthreads();
  if (W == nullptr) {                                       // scores only (kernel-SVM decoders)
    if (threadIdx.x == 0) ws_count[trial] = 0;
    return;
  }
  double* dv = s + k2p;                                     // ncls decisions
  for (int c = wid; c < ncls; c += nw) {
    const double* w = W + (long long)c * (k2 + 1);
    double a = 0.0;
    for (int j = lane; j < k2; j += 32) a = fma(s[j], w[j], a);
    a = warp_sum(a);
    if (lane == 0) {
      dv[c] = a + w[k2];
      if (dec) dec[(long long)trial * ncls + c] = dv[c];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int best = 0;
    double bv = dv[0];
    for (int c = 1; c < ncls; ++c)
      if (dv[c] > bv) { bv = dv[c]; best = c; }
    if (ncls == 1) best = dv[0] > 0.0 ? 1 : 0;               // sklearn's binary case (one row)
    yhat[trial] = classes[best];
    ws_count[trial] = 0;                                     // ready for the next call
  }
}

}  // namespace

// X: (n, T, C) fp64 on the device; mu (C) or NULL; A: (C x Q); pmean: (T*Q); P: (T*Q x k2) =
// components_^T; W: (ncls x (k2+1)) with the intercept last (ncls = 1: binary, classes has 2
// entries); yhat: n labels; dec (optional): (n x ncls) decision values.  nsplit time slices per
// trial (1..T; more slices = lower latency for few trials); ws_part: n * nsplit * roundup(k2, 32)
// doubles; ws_count: n ints, zero before the first call (the kernel leaves them zero).
// scores_out (optional): the PCA scores, feature-major (k2 x ld_scores) fp32 -- with W = NULL the
// kernel stops there (input of cpsd_svc_predict_ovo for the kernel-SVM decoders).
extern "C" int cpsd_predict_fused(const double* X, int n, int T, int C, const float* mu, const float* A,
                                  int Q, const float* pmean, const float* P, int k2, const double* W,
                                  const int* classes, int ncls, int* yhat, double* dec, int nsplit,
                                  double* ws_part, int* ws_count, float* scores_out, int ld_scores,
                                  cudaStream_t stream) {
  CPSD_CHECK_ARG(n >= 0 && T > 0 && C > 0 && Q > 0 && k2 > 0 && ncls > 0, "predict_fused: bad dims");
  CPSD_CHECK_ARG(W != nullptr || scores_out != nullptr, "predict_fused: neither W nor scores_out");
  CPSD_CHECK_ARG(scores_out == nullptr || ld_scores >= n, "predict_fused: ld_scores < n");
  CPSD_CHECK_ARG(nsplit >= 1 && nsplit <= T && nsplit <= 65535 && ws_part && ws_count,
                 "predict_fused: bad nsplit / workspace");
  if (n == 0) return CPSD_OK;
  const int rows_per = (T + nsplit - 1) / nsplit;
  const size_t k2p = (size_t)((k2 + 31) & ~31);
  const size_t smem = sizeof(float) * ((((size_t)rows_per * C + 1) & ~(size_t)1) +
                                       (((size_t)rows_per * Q + 1) & ~(size_t)1)) +
                      sizeof(double) * ((PF_NT / 32) * k2p + k2p + (size_t)ncls + 2);
  CPSD_CHECK_ARG(smem <= 227 * 1024, "predict_fused: slice (rows x C) + latent (rows x Q) exceed shared memory");
  CPSD_CUDA(cudaFuncSetAttribute(k_predict_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_predict_fused<<<dim3(n, nsplit), PF_NT, smem, stream>>>(X, T, C, mu, A, Q, pmean, P, k2, W, classes, ncls,
                                                            yhat, dec, rows_per, ws_part, ws_count,
                                                            scores_out, ld_scores);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}
