// Projection of every trial of every patient into the pooled (trial x time*latent) matrix of
// every fold on the 5th-generation tensor cores (sm_100a only):
//
//      Z[f][dst(f, v, trial)][t][:] = (X_v[trial][t][:] - mu_{f,v}) L_{f,v}
//
// (MCCA transform_view, alignment/AlignMCCA.py:110,125; PCA.transform / AlignCCA.transform
// AlignCCA.py:93; pooling decoders/cross_pt_decoders.py:260-270).  The trials of a patient
// are the same for every fold -- only the C x Q loadings change -- so the work is organised
// around the X tile: a persistent CTA loads one 128-row tile of X_v (all channels, tf32 hi/lo
// split precomputed once per patient) into shared memory with TMA and then streams the
// loadings of a whole group of folds past it:
//
//   warp 0   : TMA producer  (X tile once per item; L^T hi/lo of each fold into a 2-stage ring)
//   warp 1   : tcgen05.mma.kind::tf32 issuer, M=128 x N=32 x K=8, 3xTF32
//              (hi*hi + hi*lo + lo*hi), accumulator in TMEM, two accumulator stages
//   warps 2-5: epilogue: tcgen05.ld -> subtract mu L -> smem staging -> coalesced stores to
//              the fold's pooled matrix (per-trial destination rows from a table)
//
// HBM traffic per batch: X (hi+lo) once + the pooled matrices once; the loadings and the
// repeated X tiles come from L2.
#include "tc_common.cuh"
#include <cuda.h>

namespace {
using namespace tc;

constexpr int PT_BM = 128;                 // rows per tile
constexpr int PT_N = 32;                   // latent columns (padded)
constexpr int PT_BK = 32;                  // fp32 per 128-byte swizzle row
constexpr int PT_KB = 4;                   // k-blocks: channels <= 128
constexpr int PT_A_TILE = PT_BM * PT_BK * 4;          // 16 KB
constexpr int PT_B_TILE = PT_N * PT_BK * 4;           // 4 KB
constexpr int PT_A_BYTES = PT_KB * 2 * PT_A_TILE;     // 128 KB
constexpr int PT_B_STAGE = PT_KB * 2 * PT_B_TILE;     // 32 KB
constexpr int PT_B_STAGES = 2;
constexpr int PT_LDS = 33;                 // staging row stride (floats)
constexpr int PT_STAGING = PT_BM * PT_LDS * 4;
constexpr int PT_THREADS = 192;
constexpr int PT_MAXP = 16;
constexpr uint32_t PT_TMEM_COLS = 64;      // two 32-column accumulator stages

struct ProjTcParams {
  int P, B, T, Q;
  int n_max;             // trials per patient in the destination table (row stride)
  int fg, ngroups;       // folds per group, groups per tile
  int ntile_total;
  int tile_prefix[PT_MAXP + 1];
  int nrows[PT_MAXP];    // N_v * T
  int kblocks[PT_MAXP];  // ceil(C_v / 32)
  long long strideY;     // floats between the pooled matrices of consecutive folds
};

__global__ void __launch_bounds__(PT_THREADS, 1)
k_proj_tc(const CUtensorMap* __restrict__ xmaps, const CUtensorMap* __restrict__ ltmaps,
          const ProjTcParams prm, const int* __restrict__ dst_row, const float* __restrict__ muL,
          float* __restrict__ Y) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + PT_A_BYTES;
  float* stg = reinterpret_cast<float*>(sB + PT_B_STAGES * PT_B_STAGE);
  int* row_off = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(stg) + PT_STAGING);  // [128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(row_off + 2 * PT_BM);
  uint64_t* a_full = bars;            // 1
  uint64_t* a_empty = bars + 1;       // 1
  uint64_t* b_full = bars + 2;        // 2
  uint64_t* b_empty = bars + 4;       // 2
  uint64_t* t_full = bars + 6;        // 2
  uint64_t* t_empty = bars + 8;       // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nitems = prm.ntile_total * prm.ngroups;

  if (warp == 0 && lane == 0) {
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
      mbar_init(&t_full[s], 1);
      mbar_init(&t_empty[s], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(PT_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // item -> (patient v, tile inside the patient, fold group)
  auto decode = [&](int item, int& v, int& tile, int& f0, int& f1) {
    const int tg = item / prm.ngroups, g = item - tg * prm.ngroups;
    v = 0;
    while (v + 1 < prm.P && tg >= prm.tile_prefix[v + 1]) ++v;
    tile = tg - prm.tile_prefix[v];
    f0 = g * prm.fg;
    f1 = min(prm.B, f0 + prm.fg);
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t it_a = 0, it_b = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it_a) {
        int v, tile, f0, f1;
        decode(item, v, tile, f0, f1);
        const int kbn = prm.kblocks[v];
        const CUtensorMap* mh = xmaps + 2 * v;
        const CUtensorMap* ml = mh + 1;
        mbar_wait(a_empty, (it_a & 1u) ^ 1u);
        mbar_expect_tx(a_full, (uint32_t)(kbn * 2 * PT_A_TILE));
        for (int kb = 0; kb < kbn; ++kb) {
          tma_load_2d(sA + (kb * 2) * PT_A_TILE, mh, kb * PT_BK, tile * PT_BM, a_full);
          tma_load_2d(sA + (kb * 2 + 1) * PT_A_TILE, ml, kb * PT_BK, tile * PT_BM, a_full);
        }
        for (int f = f0; f < f1; ++f, ++it_b) {
          const int s = it_b & 1;
          mbar_wait(&b_empty[s], ((it_b >> 1) & 1u) ^ 1u);
          mbar_expect_tx(&b_full[s], (uint32_t)(kbn * 2 * PT_B_TILE));
          uint8_t* st = sB + s * PT_B_STAGE;
          const int brow = (f * prm.P + v) * PT_N;
          for (int kb = 0; kb < kbn; ++kb) {
            tma_load_2d(st + (kb * 2) * PT_B_TILE, ltmaps, kb * PT_BK, brow, &b_full[s]);
            tma_load_2d(st + (kb * 2 + 1) * PT_B_TILE, ltmaps + 1, kb * PT_BK, brow, &b_full[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = idesc_tf32(PT_BM, PT_N);
      uint32_t it_a = 0, it_b = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it_a) {
        int v, tile, f0, f1;
        decode(item, v, tile, f0, f1);
        const int kbn = prm.kblocks[v];
        mbar_wait(a_full, it_a & 1u);
        for (int f = f0; f < f1; ++f, ++it_b) {
          const int s = it_b & 1;
          const uint32_t ph = (it_b >> 1) & 1u;
          mbar_wait(&b_full[s], ph);
          mbar_wait(&t_empty[s], ph ^ 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t tacc = tmem_base + (uint32_t)(s * PT_N);
          const uint32_t sa = smem_u32(sA);
          const uint32_t sb = smem_u32(sB + s * PT_B_STAGE);
          for (int kb = 0; kb < kbn; ++kb) {
            const uint64_t a_hi = make_smem_desc(sa + (kb * 2) * PT_A_TILE);
            const uint64_t a_lo = make_smem_desc(sa + (kb * 2 + 1) * PT_A_TILE);
            const uint64_t b_hi = make_smem_desc(sb + (kb * 2) * PT_B_TILE);
            const uint64_t b_lo = make_smem_desc(sb + (kb * 2 + 1) * PT_B_TILE);
#pragma unroll
            for (int k = 0; k < PT_BK / 8; ++k) {
              const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);
#ifdef PT_EXPERIMENT_1TERM
              umma_tf32(tacc, a_hi + adv, b_hi + adv, idesc, (kb != 0 || k != 0) ? 1u : 0u);
#else
              umma_tf32(tacc, a_lo + adv, b_hi + adv, idesc, (kb != 0 || k != 0) ? 1u : 0u);
              umma_tf32(tacc, a_hi + adv, b_lo + adv, idesc, 1u);
              umma_tf32(tacc, a_hi + adv, b_hi + adv, idesc, 1u);
#endif
            }
          }
          umma_commit(&b_empty[s]);
          umma_commit(&t_full[s]);
        }
        umma_commit(a_empty);          // X tile free once every fold's MMAs retired
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue
    const int quad = warp & 3;
    const int r_own = quad * 32 + lane;        // TMEM lane = tile row owned by this thread
    const int Q = prm.Q;
    // every epilogue warp stages and writes its own 32 rows (no cross-warp barrier): copy-out
    // pattern of this lane = element e = lane + 32 i of the warp's (32 x Q) block
    uint32_t rj[PT_N];
#pragma unroll
    for (int i = 0; i < PT_N; ++i) {
      // the block has 32 Q elements = Q per lane; slots i >= Q repeat an earlier element of the
      // same lane (a benign duplicate store) so that the copy-out loop needs no validity test
      const int e = lane + 32 * (i % Q);
      const int r = e / Q, j = e - r * Q;
      rj[i] = (uint32_t)((quad * 32 + r) << 8) | (uint32_t)j;
    }
    uint32_t it_b = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
      int v, tile, f0, f1;
      decode(item, v, tile, f0, f1);
      // this thread's row of the tile: trial and time bin (item-invariant)
      const int R = tile * PT_BM + r_own;
      int my_tr = -1, my_t = 0;
      if (R < prm.nrows[v]) {
        my_tr = R / prm.T;
        my_t = R - my_tr * prm.T;
      }
      const int* drow = dst_row + ((long long)f0 * prm.P + v) * prm.n_max;
      int d_next = (my_tr >= 0) ? drow[my_tr] : -1;
      // mu L of the fold: lane j keeps column j, broadcast by shuffle in the epilogue
      float ml_next = muL[(long long)(f0 * prm.P + v) * PT_N + lane];
      for (int f = f0; f < f1; ++f, ++it_b) {
        const int s = it_b & 1;
        const int d = d_next;
        const float ml_mine = ml_next;
        if (f + 1 < f1) {        // destination / mu L of the next fold: in flight during this one
          drow += (long long)prm.P * prm.n_max;
          d_next = (my_tr >= 0) ? drow[my_tr] : -1;
          ml_next = muL[(long long)((f + 1) * prm.P + v) * PT_N + lane];
        }
        mbar_wait(&t_full[s], (it_b >> 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t vv[32];
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(s * PT_N), vv);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_empty[s]);
#pragma unroll
        for (int j = 0; j < PT_N; ++j)
          stg[r_own * PT_LDS + j] = __uint_as_float(vv[j]) - __shfl_sync(0xffffffffu, ml_mine, j);
        row_off[r_own] = (d >= 0) ? (d * prm.T + my_t) * Q : -1;
        __syncwarp();
        float* Yf = Y + (long long)f * prm.strideY;
        // all shared-memory reads first (independent), then the predicated global stores
        int offs[PT_N];
        float vals[PT_N];
#pragma unroll
        for (int i = 0; i < PT_N; ++i) {
          const int r = (int)(rj[i] >> 8), j = (int)(rj[i] & 255u);
          const int off = row_off[r];
          offs[i] = (off >= 0) ? off + j : -1;
          vals[i] = stg[r * PT_LDS + j];
        }
#pragma unroll
        for (int i = 0; i < PT_N; ++i)
          if (offs[i] >= 0) Yf[offs[i]] = vals[i];
        __syncwarp();
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(PT_TMEM_COLS)
                 : "memory");
  }
}

// Per problem p = fold * P + view: L^T split into tf32 hi / lo (32 x 128, zero padded) and the
// row vector mu L.  mu of problem p lives at mu_base + slot[p] * ld_mu.
__global__ void __launch_bounds__(128)
k_proj_tc_prep(const float* __restrict__ L, int ldl, long long strideL, const float* __restrict__ mu_base,
               const int* __restrict__ slot, int ld_mu, const int* __restrict__ cdim, int Q,
               float* __restrict__ LtHi, float* __restrict__ LtLo, float* __restrict__ muL) {
  const int p = blockIdx.x;
  const int C = min(cdim[p], PT_KB * PT_BK);
  const float* Lp = L + (long long)p * strideL;
  const float* mu = mu_base ? mu_base + (long long)(slot ? slot[p] : p) * ld_mu : nullptr;
  float* hi = LtHi + (long long)p * PT_N * (PT_KB * PT_BK);
  float* lo = LtLo + (long long)p * PT_N * (PT_KB * PT_BK);
  for (int e = threadIdx.x; e < PT_N * PT_KB * PT_BK; e += blockDim.x) {
    const int j = e / (PT_KB * PT_BK), c = e - j * (PT_KB * PT_BK);
    float x = 0.f;
    if (j < Q && c < C) x = Lp[(long long)c * ldl + j];
    float h, l;
    split_tf32(x, h, l);
    hi[e] = h;
    lo[e] = l;
  }
  if (threadIdx.x < PT_N) {
    const int j = threadIdx.x;
    double a = 0.0;
    if (mu && j < Q)
      for (int c = 0; c < C; ++c) a = fma((double)mu[c], (double)Lp[(long long)c * ldl + j], a);
    muL[(long long)p * PT_N + j] = (float)a;
  }
}

__global__ void __launch_bounds__(256)
k_split_flat(const float* __restrict__ src, float* __restrict__ hi, float* __restrict__ lo,
             long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float h, l;
    split_tf32(src[i], h, l);
    hi[i] = h;
    lo[i] = l;
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                             CUtensorMapFloatOOBfill);

EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(p);
  }
  return fn;
}

}  // namespace

// 128-byte tensor map (HOST output) of a row-major fp32 matrix (rows x cols, row stride ld
// floats, ld % 4 == 0) with a (box_rows x 32-float) box and 128-byte swizzle.
extern "C" int cpsd_tmap_encode_f32(void* map_out_host, const float* base, long long rows, int cols,
                                    long long ld, int box_rows) {
  CPSD_CHECK_ARG(rows > 0 && cols > 0 && ld >= cols && (ld & 3) == 0, "tmap_encode: bad dims (ld % 4)");
  CPSD_CHECK_ARG(box_rows > 0 && box_rows <= 256, "tmap_encode: bad box");
  CPSD_CHECK_ARG((((uintptr_t)base) & 15) == 0, "tmap_encode: base must be 16-byte aligned");
  EncodeFn enc = get_encode();
  if (!enc) {
    cpsd_set_error("tmap_encode: cuTensorMapEncodeTiled entry point unavailable");
    return CPSD_ERR_CUDA;
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {PT_BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(reinterpret_cast<CUtensorMap*>(map_out_host), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<float*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    cpsd_set_error("tmap_encode: cuTensorMapEncodeTiled failed");
    return CPSD_ERR_CUDA;
  }
  return CPSD_OK;
}

// x -> (hi, lo) tf32 split of a whole array (hi exactly representable in tf32, lo = x - hi)
extern "C" int cpsd_split_tf32(const float* src, float* hi, float* lo, long long n,
                               cudaStream_t stream) {
  CPSD_CHECK_ARG(n >= 0, "split_tf32: bad n");
  if (n == 0) return CPSD_OK;
  long long nb = (n + 255) / 256;
  if (nb > 148 * 16) nb = 148 * 16;
  k_split_flat<<<(int)nb, 256, 0, stream>>>(src, hi, lo, n);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

extern "C" int cpsd_proj_tc_prep(const float* L, int ldl, long long strideL, const float* mu_base,
                                 const int* slot, int ld_mu, const int* cdim, int Q, float* LtHi,
                                 float* LtLo, float* muL, int nprob, cudaStream_t stream) {
  CPSD_CHECK_ARG(nprob >= 0 && Q > 0 && Q <= PT_N, "proj_tc_prep: Q must be in 1..32");
  if (nprob == 0) return CPSD_OK;
  k_proj_tc_prep<<<nprob, 128, 0, stream>>>(L, ldl, strideL, mu_base, slot, ld_mu, cdim, Q, LtHi, LtLo,
                                            muL);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

// xmaps_dev: 2*P tensor maps (hi, lo per patient; box 128 rows), ltmaps_dev: 2 maps over the
// LtHi / LtLo arrays (box 32 rows).  n_trials[v], n_chan[v]: patient shapes (channels <= 128,
// multiple of 4).  dst_row: [B][P][n_max] destination trial row in fold f's pooled matrix (or
// -1).  Y: pooled matrices, fold stride strideY floats, row stride Q*T (trial) / Q (time bin).
extern "C" int cpsd_proj_tc(const void* xmaps_dev, const void* ltmaps_dev, int P, int B, int T, int Q,
                            const int* n_trials_host, const int* n_chan_host, int n_max,
                            const int* dst_row, const float* muL, float* Y, long long strideY,
                            int num_sms, cudaStream_t stream) {
  CPSD_CHECK_ARG(P > 0 && P <= PT_MAXP && B > 0 && T > 0, "proj_tc: bad dims");
  CPSD_CHECK_ARG(Q > 0 && Q <= PT_N, "proj_tc: Q must be in 1..32");
  ProjTcParams prm;
  prm.P = P; prm.B = B; prm.T = T; prm.Q = Q; prm.n_max = n_max; prm.strideY = strideY;
  int tot = 0;
  for (int v = 0; v < P; ++v) {
    CPSD_CHECK_ARG(n_chan_host[v] > 0 && n_chan_host[v] <= PT_KB * PT_BK && (n_chan_host[v] & 3) == 0,
                   "proj_tc: channels must be a multiple of 4 and <= 128");
    CPSD_CHECK_ARG(n_trials_host[v] <= n_max, "proj_tc: n_trials > n_max");
    prm.tile_prefix[v] = tot;
    prm.nrows[v] = n_trials_host[v] * T;
    prm.kblocks[v] = (n_chan_host[v] + PT_BK - 1) / PT_BK;
    tot += (prm.nrows[v] + PT_BM - 1) / PT_BM;
  }
  for (int v = P; v < PT_MAXP; ++v) { prm.nrows[v] = 0; prm.kblocks[v] = 0; }
  for (int v = P; v <= PT_MAXP; ++v) prm.tile_prefix[v] = tot;
  prm.ntile_total = tot;
  if (tot == 0) return CPSD_OK;
  if (num_sms <= 0) num_sms = 148;
  // fold groups: enough items to balance the persistent grid, few enough to amortise the X tile
  int ngroups = 1;
  while (ngroups * 2 <= B && (long long)tot * ngroups < 8LL * num_sms) ngroups *= 2;
  prm.fg = (B + ngroups - 1) / ngroups;
  prm.ngroups = (B + prm.fg - 1) / prm.fg;
  const long long nitems = (long long)tot * prm.ngroups;
  const int grid = (int)(nitems < num_sms ? nitems : num_sms);
  const size_t smem = 1024 + PT_A_BYTES + PT_B_STAGES * PT_B_STAGE + PT_STAGING + 2 * PT_BM * 4 + 128;
  CPSD_CUDA(cudaFuncSetAttribute(k_proj_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_proj_tc<<<grid, PT_THREADS, smem, stream>>>(reinterpret_cast<const CUtensorMap*>(xmaps_dev),
                                                reinterpret_cast<const CUtensorMap*>(ltmaps_dev), prm,
                                                dst_row, muL, Y);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}
