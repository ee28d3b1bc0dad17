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
constexpr int PT_N = 32;                   // latent columns per MMA (wider latents run as chunks of 32)
constexpr int PT_BK = 32;                  // fp32 per 128-byte swizzle row
constexpr int PT_KB = 4;                   // k-blocks per resident X panel: 128 channels
constexpr int PT_MAXC = 256;               // channels per patient: two panels, the second adds to Y
constexpr int PT_MAXQ = 128;               // latent columns: four chunks
constexpr int PT_A_TILE = PT_BM * PT_BK * 4;          // 16 KB
constexpr int PT_B_TILE = PT_N * PT_BK * 4;           // 4 KB
constexpr int PT_A_BYTES = PT_KB * 2 * PT_A_TILE;     // 128 KB
constexpr int PT_B_STAGE = PT_KB * 2 * PT_B_TILE;     // 32 KB
constexpr int PT_B_STAGES = 2;
constexpr int PT_LDS = 33;                 // staging row stride (floats)
constexpr int PT_STAGING = PT_BM * PT_LDS * 4;
constexpr int PT_THREADS = 192;
constexpr int PT_MAXP = 128;               // (replica, patient) pairs of one launch
constexpr uint32_t PT_TMEM_COLS = 64;      // two 32-column accumulator stages

struct ProjTcParams {
  int P, B, T, Q;        // P = patients per replica (row stride of the L / destination tables)
  int PV;                // virtual patients = replicas x P: virtual patient vv = replica * P + v is
                         // multiplied with the loadings of the folds [fbeg[vv], fbeg[vv] + fcnt[vv])
  int fbeg[PT_MAXP];
  int fcnt[PT_MAXP];
  int nq;                // latent chunks of 32 columns (Q <= 32: 1)
  int ltc;               // columns of the L^T arrays (128 or 256)
  int n_max;             // trials per patient in the destination table (row stride)
  int fg, ngroups;       // folds per group, groups per tile
  int ntile_total;
  int tile_prefix[PT_MAXP + 1];
  int nrows[PT_MAXP];    // N_v * T
  int kblocks[PT_MAXP];  // ceil(C_v / 32), up to 8
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
  // vv: virtual patient (indexes the X maps / shapes); v = vv % P indexes the per-fold tables
  auto decode = [&](int item, int& vv, int& v, int& tile, int& f0, int& f1) {
    const int tg = item / prm.ngroups, g = item - tg * prm.ngroups;
    vv = 0;
    while (vv + 1 < prm.PV && tg >= prm.tile_prefix[vv + 1]) ++vv;
    v = vv % prm.P;
    tile = tg - prm.tile_prefix[vv];
    f0 = prm.fbeg[vv] + g * prm.fg;
    f1 = min(prm.fbeg[vv] + prm.fcnt[vv], f0 + prm.fg);     // empty when this group lies past the range
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t it_a = 0, it_b = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        int vv, v, tile, f0, f1;
        decode(item, vv, v, tile, f0, f1);
        if (f0 >= f1) continue;
        const CUtensorMap* mh = xmaps + 2 * vv;
        const CUtensorMap* ml = mh + 1;
        // channel panels of <= 128 channels: the X panel stays in shared memory while the
        // loadings of every (fold, latent chunk) of the group stream past it
        for (int kb0 = 0; kb0 < prm.kblocks[vv]; kb0 += PT_KB, ++it_a) {
          const int kbn = min(PT_KB, prm.kblocks[vv] - kb0);
          mbar_wait(a_empty, (it_a & 1u) ^ 1u);
          mbar_expect_tx(a_full, (uint32_t)(kbn * 2 * PT_A_TILE));
          for (int kb = 0; kb < kbn; ++kb) {
            tma_load_2d(sA + (kb * 2) * PT_A_TILE, mh, (kb0 + kb) * PT_BK, tile * PT_BM, a_full);
            tma_load_2d(sA + (kb * 2 + 1) * PT_A_TILE, ml, (kb0 + kb) * PT_BK, tile * PT_BM, a_full);
          }
          for (int vf = f0 * prm.nq; vf < f1 * prm.nq; ++vf, ++it_b) {
            const int f = vf / prm.nq, qc = vf - f * prm.nq;
            const int s = it_b & 1;
            mbar_wait(&b_empty[s], ((it_b >> 1) & 1u) ^ 1u);
            mbar_expect_tx(&b_full[s], (uint32_t)(kbn * 2 * PT_B_TILE));
            uint8_t* st = sB + s * PT_B_STAGE;
            const int brow = ((f * prm.P + v) * prm.nq + qc) * PT_N;
            for (int kb = 0; kb < kbn; ++kb) {
              tma_load_2d(st + (kb * 2) * PT_B_TILE, ltmaps, (kb0 + kb) * PT_BK, brow, &b_full[s]);
              tma_load_2d(st + (kb * 2 + 1) * PT_B_TILE, ltmaps + 1, (kb0 + kb) * PT_BK, brow, &b_full[s]);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // (whole warp, converged; one elected lane issues: see umma_tf32_w)
    {
      const uint32_t idesc = idesc_tf32(PT_BM, PT_N);
      uint32_t it_a = 0, it_b = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        int vv, v, tile, f0, f1;
        decode(item, vv, v, tile, f0, f1);
        if (f0 >= f1) continue;
        for (int kb0 = 0; kb0 < prm.kblocks[vv]; kb0 += PT_KB, ++it_a) {
        const int kbn = min(PT_KB, prm.kblocks[vv] - kb0);
        mbar_wait(a_full, it_a & 1u);
        for (int vf = f0 * prm.nq; vf < f1 * prm.nq; ++vf, ++it_b) {
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
              umma_tf32_w(tacc, a_hi + adv, b_hi + adv, idesc, (kb != 0 || k != 0) ? 1u : 0u);
#else
              umma_tf32_w(tacc, a_lo + adv, b_hi + adv, idesc, (kb != 0 || k != 0) ? 1u : 0u);
              umma_tf32_w(tacc, a_hi + adv, b_lo + adv, idesc, 1u);
              umma_tf32_w(tacc, a_hi + adv, b_hi + adv, idesc, 1u);
#endif
            }
          }
          umma_commit_w(&b_empty[s]);
          umma_commit_w(&t_full[s]);
        }
        umma_commit_w(a_empty);          // X panel free once every fold's MMAs retired
        }
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
      const int Qc = Q < PT_N ? Q : PT_N;
      const int e = lane + 32 * (i % Qc);
      const int r = e / Qc, j = e - r * Qc;
      rj[i] = (uint32_t)((quad * 32 + r) << 8) | (uint32_t)j;
    }
    // staging index of the same elements (fast copy-out below)
    int sidx[PT_N];
#pragma unroll
    for (int i = 0; i < PT_N; ++i) sidx[i] = (int)(rj[i] >> 8) * PT_LDS + (int)(rj[i] & 255u);
    const int Qc_ = Q < PT_N ? Q : PT_N;
    uint32_t it_b = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
      int vv, v, tile, f0, f1;
      decode(item, vv, v, tile, f0, f1);
      if (f0 >= f1) continue;
      // this thread's row of the tile: trial and time bin (item-invariant)
      const int R = tile * PT_BM + r_own;
      int my_tr = -1, my_t = 0;
      if (R < prm.nrows[vv]) {
        my_tr = R / prm.T;
        my_t = R - my_tr * prm.T;
      }
      for (int kb0 = 0; kb0 < prm.kblocks[vv]; kb0 += PT_KB) {
      const bool first = kb0 == 0;       // later channel panels add to what the first one wrote
      for (int vf = f0 * prm.nq; vf < f1 * prm.nq; ++vf, ++it_b) {
        const int f = vf / prm.nq, qc = vf - f * prm.nq;
        const int s = it_b & 1;
        const int d = (my_tr >= 0) ? dst_row[((long long)f * prm.P + v) * prm.n_max + my_tr] : -1;
        // mu L of the (fold, chunk): 32 floats, the same for every lane (broadcast 16-byte loads,
        // issued before the wait for the accumulator)
        float4 ml4[PT_N / 4];
        {
          const float4* mp = reinterpret_cast<const float4*>(
              muL + ((long long)(f * prm.P + v) * prm.nq + qc) * PT_N);
#pragma unroll
          for (int j = 0; j < PT_N / 4; ++j) ml4[j] = first ? __ldg(mp + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        mbar_wait(&t_full[s], (it_b >> 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t vv[32];
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(s * PT_N), vv);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_empty[s]);
#pragma unroll
        for (int j = 0; j < PT_N / 4; ++j) {
          stg[r_own * PT_LDS + 4 * j + 0] = __uint_as_float(vv[4 * j + 0]) - ml4[j].x;
          stg[r_own * PT_LDS + 4 * j + 1] = __uint_as_float(vv[4 * j + 1]) - ml4[j].y;
          stg[r_own * PT_LDS + 4 * j + 2] = __uint_as_float(vv[4 * j + 2]) - ml4[j].z;
          stg[r_own * PT_LDS + 4 * j + 3] = __uint_as_float(vv[4 * j + 3]) - ml4[j].w;
        }
        const int my_off = (d >= 0) ? (d * prm.T + my_t) * Q : -1;
        row_off[r_own] = my_off;
        // fast copy-out: the warp's 32 rows are consecutive time bins of ONE kept trial, i.e. one
        // contiguous block of 32 Q floats that starts at lane 0's row -- element e = lane + 32 i
        // goes to base + e, no per-element row lookup or predicate (5 of 6 blocks with T = 200)
        const int base0 = __shfl_sync(0xffffffffu, my_off, 0);
        const bool contiguous = __all_sync(0xffffffffu, my_off >= 0 && my_off == base0 + lane * Q);
        __syncwarp();
        float* Yf = Y + (long long)f * prm.strideY;
        if (prm.nq == 1 && first && contiguous) {
          float* Yb = Yf + base0 + lane;
#pragma unroll
          for (int i = 0; i < PT_N; ++i)
            if (i < Qc_) Yb[32 * i] = stg[sidx[i]];
        } else if (prm.nq == 1) {
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
          if (first) {
#pragma unroll
            for (int i = 0; i < PT_N; ++i)
              if (offs[i] >= 0) Yf[offs[i]] = vals[i];
          } else {
            // (no duplicate slots here: an element is added exactly once)
#pragma unroll
            for (int i = 0; i < PT_N; ++i)
              if (i < Q && offs[i] >= 0) Yf[offs[i]] += vals[i];
          }
        } else {
          // wide latents: this chunk is 32 consecutive floats (128 bytes) of every output row
          const int col = qc * PT_N + lane;
          if (col < Q) {
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
              const int off = row_off[quad * 32 + r];
              if (off >= 0) {
                const float val = stg[(quad * 32 + r) * PT_LDS + lane];
                if (first) Yf[off + col] = val; else Yf[off + col] += val;
              }
            }
          }
        }
        __syncwarp();
      }
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

// Per problem p = fold * P + view and latent chunk qc: 32 rows of L^T (latent columns qc*32 ..
// qc*32+31) split into tf32 hi / lo, each row ltc floats (channels, zero padded), and the row
// vector mu L of the chunk.  mu of problem p lives at mu_base + slot[p] * ld_mu.
__global__ void __launch_bounds__(128)
k_proj_tc_prep(const float* __restrict__ L, int ldl, long long strideL, const float* __restrict__ mu_base,
               const int* __restrict__ slot, int ld_mu, const int* __restrict__ cdim, int Q, int nq,
               int ltc, float* __restrict__ LtHi, float* __restrict__ LtLo, float* __restrict__ muL) {
  const int p = blockIdx.x, qc = blockIdx.y;
  const int C = min(cdim[p], ltc);
  const float* Lp = L + (long long)p * strideL;
  const float* mu = mu_base ? mu_base + (long long)(slot ? slot[p] : p) * ld_mu : nullptr;
  const long long blk = (long long)p * nq + qc;
  float* hi = LtHi + blk * PT_N * ltc;
  float* lo = LtLo + blk * PT_N * ltc;
  for (int e = threadIdx.x; e < PT_N * ltc; e += blockDim.x) {
    const int j = e / ltc, c = e - j * ltc;
    const int col = qc * PT_N + j;
    float x = 0.f;
    if (col < Q && c < C) x = Lp[(long long)c * ldl + col];
    float h, l;
    split_tf32(x, h, l);
    hi[e] = h;
    lo[e] = l;
  }
  if (threadIdx.x < PT_N) {
    const int col = qc * PT_N + threadIdx.x;
    double a = 0.0;
    if (mu && col < Q)
      for (int c = 0; c < C; ++c) a = fma((double)mu[c], (double)Lp[(long long)c * ldl + col], a);
    muL[blk * PT_N + threadIdx.x] = (float)a;
  }
}

// hi / lo split of a (rows x cols) matrix (row stride lds) into arrays with row stride ldd >= cols
// (ldd a multiple of 4 floats so that TMA can address rows; the padding columns are zeroed)
__global__ void __launch_bounds__(256)
k_split_2d(const float* __restrict__ src, long long lds, long long rows, int cols,
           float* __restrict__ hi, float* __restrict__ lo, int ldd) {
  const long long total = rows * ldd;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const long long r = e / ldd;
    const int c = (int)(e - r * ldd);
    float h = 0.f, l = 0.f;
    if (c < cols) split_tf32(src[r * lds + c], h, l);
    hi[e] = h;
    lo[e] = l;
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

extern "C" int cpsd_split_tf32_2d(const float* src, long long lds, long long rows, int cols, float* hi,
                                  float* lo, int ldd, cudaStream_t stream) {
  CPSD_CHECK_ARG(rows >= 0 && cols > 0 && lds >= cols && ldd >= cols && (ldd & 3) == 0,
                 "split_tf32_2d: bad dims (ldd % 4)");
  if (rows == 0) return CPSD_OK;
  long long nb = (rows * ldd + 255) / 256;
  if (nb > 148 * 16) nb = 148 * 16;
  k_split_2d<<<(int)nb, 256, 0, stream>>>(src, lds, rows, cols, hi, lo, ldd);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

extern "C" int cpsd_proj_tc_prep(const float* L, int ldl, long long strideL, const float* mu_base,
                                 const int* slot, int ld_mu, const int* cdim, int Q, int ltc,
                                 float* LtHi, float* LtLo, float* muL, int nprob,
                                 cudaStream_t stream) {
  CPSD_CHECK_ARG(nprob >= 0 && Q > 0 && Q <= PT_MAXQ, "proj_tc_prep: Q must be in 1..128");
  CPSD_CHECK_ARG(ltc == 128 || ltc == 256, "proj_tc_prep: ltc must be 128 or 256");
  if (nprob == 0) return CPSD_OK;
  const int nq = (Q + PT_N - 1) / PT_N;
  k_proj_tc_prep<<<dim3(nprob, nq), 128, 0, stream>>>(L, ldl, strideL, mu_base, slot, ld_mu, cdim, Q, nq,
                                                      ltc, LtHi, LtLo, muL);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

// xmaps_dev: 2*P tensor maps (hi, lo per patient; box 128 rows), ltmaps_dev: 2 maps over the
// LtHi / LtLo arrays (box 32 rows, ltc columns).  n_trials[v], n_chan[v]: patient shapes
// (channels <= 256; the hi / lo arrays may have a padded row stride, cpsd_split_tf32_2d).
// dst_row: [B][P][n_max] destination trial row in fold f's pooled matrix (or -1).  Y: pooled
// matrices, fold stride strideY floats, row stride Q*T (trial) / Q (time bin); Q <= 128.
extern "C" int cpsd_proj_tc_rep(const void* xmaps_dev, const void* ltmaps_dev, int P, int n_rep, int B,
                                int T, int Q, int ltc, const int* n_trials_host, const int* n_chan_host,
                                const int* fold_beg_host, const int* fold_cnt_host, int n_max,
                                const int* dst_row, const float* muL, float* Y, long long strideY,
                                int num_sms, cudaStream_t stream) {
  CPSD_CHECK_ARG(P > 0 && n_rep > 0 && P * n_rep <= PT_MAXP && B > 0 && T > 0,
                 "proj_tc: bad dims (patients x replicas <= 128)");
  CPSD_CHECK_ARG(Q > 0 && Q <= PT_MAXQ, "proj_tc: Q must be in 1..128");
  CPSD_CHECK_ARG(ltc == 128 || ltc == 256, "proj_tc: ltc must be 128 or 256");
  ProjTcParams prm;
  prm.P = P; prm.B = B; prm.T = T; prm.Q = Q; prm.n_max = n_max; prm.strideY = strideY;
  prm.PV = P * n_rep;
  prm.nq = (Q + PT_N - 1) / PT_N;
  prm.ltc = ltc;
  int tot = 0, fmax = 0;
  for (int vv = 0; vv < prm.PV; ++vv) {
    CPSD_CHECK_ARG(n_chan_host[vv] > 0 && n_chan_host[vv] <= ltc && n_chan_host[vv] <= PT_MAXC,
                   "proj_tc: channels must be <= 256 (and <= ltc)");
    CPSD_CHECK_ARG(n_trials_host[vv] <= n_max, "proj_tc: n_trials > n_max");
    const int r = vv / P;
    prm.fbeg[vv] = fold_beg_host ? fold_beg_host[r] : 0;
    prm.fcnt[vv] = fold_cnt_host ? fold_cnt_host[r] : B;
    CPSD_CHECK_ARG(prm.fbeg[vv] >= 0 && prm.fcnt[vv] >= 0 && prm.fbeg[vv] + prm.fcnt[vv] <= B,
                   "proj_tc: fold range outside the batch");
    if (prm.fcnt[vv] > fmax) fmax = prm.fcnt[vv];
    prm.tile_prefix[vv] = tot;
    prm.nrows[vv] = n_trials_host[vv] * T;
    prm.kblocks[vv] = (n_chan_host[vv] + PT_BK - 1) / PT_BK;
    tot += (prm.nrows[vv] + PT_BM - 1) / PT_BM;
  }
  for (int vv = prm.PV; vv < PT_MAXP; ++vv) { prm.nrows[vv] = 0; prm.kblocks[vv] = 0; prm.fbeg[vv] = 0; prm.fcnt[vv] = 0; }
  for (int vv = prm.PV; vv <= PT_MAXP; ++vv) prm.tile_prefix[vv] = tot;
  prm.ntile_total = tot;
  if (tot == 0 || fmax == 0) return CPSD_OK;
  if (num_sms <= 0) num_sms = 148;
  // fold groups: enough items to balance the persistent grid, few enough to amortise the X tile
  int ngroups = 1;
  while (ngroups * 2 <= fmax && (long long)tot * ngroups < 8LL * num_sms) ngroups *= 2;
  prm.fg = (fmax + ngroups - 1) / ngroups;
  prm.ngroups = (fmax + prm.fg - 1) / prm.fg;
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

extern "C" int cpsd_proj_tc(const void* xmaps_dev, const void* ltmaps_dev, int P, int B, int T, int Q,
                            int ltc, const int* n_trials_host, const int* n_chan_host, int n_max,
                            const int* dst_row, const float* muL, float* Y, long long strideY,
                            int num_sms, cudaStream_t stream) {
  return cpsd_proj_tc_rep(xmaps_dev, ltmaps_dev, P, 1, B, T, Q, ltc, n_trials_host, n_chan_host, nullptr,
                          nullptr, n_max, dst_row, muL, Y, strideY, num_sms, stream);
}
