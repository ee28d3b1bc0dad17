// Pooled Gram  K = Z Z^T  on the 5th-generation tensor cores (sm_100a only).
//
// Replaces the fp32 SIMT k_gram_nt for the decoder-stage PCA (reference:
// decomposition/DimRedReshape.py:47-49 -> sklearn PCA full SVD of the pooled
// trials x (time*latent) matrix; here the n_pool x n_pool Gram over the long axis).
//
// Precision: 3xTF32.  Z is split once into hi = tf32(Z) (low 13 mantissa bits cleared) and
// lo = Z - hi (exact in fp32); each 128x128 output tile accumulates
//        hi_i hi_j^T + hi_i lo_j^T + lo_i hi_j^T
// in fp32 in TMEM, i.e. every product term except lo*lo (relative 2^-20) -- the same error
// class as an fp32 FMA chain, which is what the PCA variance threshold needs
// (plain TF32 would perturb the spectrum at 1e-3 relative).
//
// Structure (one CTA per upper-triangular output tile, 192 threads):
//   warp 0   : TMA producer   cp.async.bulk.tensor.2d -> 128B-swizzled smem ring (3 stages
//              of {A_hi, A_lo, B_hi, B_lo}, 128 rows x 32 fp32 each)
//   warp 1   : TMEM allocator + tcgen05.mma.kind::tf32 issuer, elected lane of the converged warp (M=128, N=128,
//              K=8), tcgen05.commit releases smem stages / signals the epilogue
//   warps 2-5: epilogue  tcgen05.ld (32 lanes x 32 columns per warp) of each partial
//              accumulator -> fp32 register accumulators -> global, plus the mirrored tile
#include "tc_common.cuh"
#include "descs.h"
#include <cuda.h>

namespace {
using namespace tc;

constexpr int BM = 128, BN = 128, BK = 32;      // BK fp32 = 128 bytes = one swizzle row
constexpr int STAGES = 3;
constexpr int TILE_BYTES = BM * BK * 4;          // 16 KB
constexpr int STAGE_BYTES = 4 * TILE_BYTES;      // A_hi, A_lo, B_hi, B_lo
constexpr int TC_THREADS = 192;
constexpr uint32_t TMEM_COLS = 256;          // two 128-column accumulator stages

struct TcProb {
  float* out;
  const float* src;     // row-major source (n rows, lda floats apart)
  const float* mu;      // optional [k] vector subtracted from every row before the split
  float* hi;
  float* lo;
  int n, ldo, lda, pad_;
};

// The tensor core's fp32 accumulator truncates on every accumulate (measured: relative error
// growing linearly with K, 1e-4 at K=6000).  So TMEM only accumulates CHUNK k-blocks
// (48 MMAs); the epilogue warps drain each partial tile into round-to-nearest fp32 register
// accumulators while the MMA warp fills the other TMEM stage.
#ifndef GRAM_EXP_CHUNK
#define GRAM_EXP_CHUNK 4
#endif
constexpr int CHUNK = GRAM_EXP_CHUNK;
constexpr int ACC_STAGES = 2;

__global__ void __launch_bounds__(TC_THREADS, 1)
k_gram_tc(const CUtensorMap* __restrict__ maps, const TcProb* __restrict__ probs, int k_total) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int prob = blockIdx.z;
  const int tm = blockIdx.y, tn = blockIdx.x;
  if (tm > tn) return;
  const TcProb pr = probs[prob];
  if (tm * BM >= pr.n || tn * BN >= pr.n) return;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;            // [ACC_STAGES]
  uint64_t* tmem_empty = tmem_full + ACC_STAGES;   // [ACC_STAGES]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + ACC_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef GRAM_EXP_NOB
  const bool diag = true;
#else
  const bool diag = (tm == tn);
#endif
  const CUtensorMap* map_hi = maps + 2 * prob;
  const CUtensorMap* map_lo = maps + 2 * prob + 1;
  const int nkb = (k_total + BK - 1) / BK;
  const int nchunk = (nkb + CHUNK - 1) / CHUNK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(map_lo) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < ACC_STAGES; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 4);              // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* st = smem + stage * STAGE_BYTES;
        mbar_expect_tx(&full[stage], diag ? 2 * TILE_BYTES : 4 * TILE_BYTES);
        tma_load_2d(st, map_hi, kb * BK, tm * BM, &full[stage]);
        tma_load_2d(st + TILE_BYTES, map_lo, kb * BK, tm * BM, &full[stage]);
        if (!diag) {
          tma_load_2d(st + 2 * TILE_BYTES, map_hi, kb * BK, tn * BN, &full[stage]);
          tma_load_2d(st + 3 * TILE_BYTES, map_lo, kb * BK, tn * BN, &full[stage]);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // (the whole warp runs the loop converged; one elected lane issues, see umma_tf32_w)
    {
      // instruction descriptor: D=f32, A=B=tf32, both K-major, N=128, M=128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)(BM >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      for (int c = 0; c < nchunk; ++c) {
        const int as = c & 1;
        const uint32_t aphase = (uint32_t)(c >> 1) & 1u;
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tacc = tmem_base + (uint32_t)(as * BN);
        const int kb_end = min(nkb, (c + 1) * CHUNK);
        for (int kb = c * CHUNK; kb < kb_end; ++kb) {
          mbar_wait(&full[stage], phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t a_hi = make_smem_desc(sa);
          const uint64_t a_lo = make_smem_desc(sa + TILE_BYTES);
          const uint64_t b_hi = diag ? a_hi : make_smem_desc(sa + 2 * TILE_BYTES);
          const uint64_t b_lo = diag ? a_lo : make_smem_desc(sa + 3 * TILE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);   // 32 bytes per K=8 step
            umma_tf32_w(tacc, a_lo + adv, b_hi + adv, idesc, (kb != c * CHUNK || k != 0) ? 1u : 0u);
            umma_tf32_w(tacc, a_hi + adv, b_lo + adv, idesc, 1u);
            umma_tf32_w(tacc, a_hi + adv, b_hi + adv, idesc, 1u);
          }
          umma_commit_w(&empty[stage]);        // frees the smem stage when the MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_w(&tmem_full[as]);         // partial accumulator complete
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue
    const int quad = warp & 3;               // TMEM lane quadrant this warp may read
    const int row = tm * BM + quad * 32 + lane;
    float acc[BN];
#pragma unroll
    for (int j = 0; j < BN; ++j) acc[j] = 0.f;
    for (int c = 0; c < nchunk; ++c) {
      const int as = c & 1;
      const uint32_t aphase = (uint32_t)(c >> 1) & 1u;
      mbar_wait(&tmem_full[as], aphase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * BN + c0);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, "
            "[%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
              "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]),
              "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
              "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
              "=r"(v[31])
            : "r"(taddr)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
#ifdef GRAM_EXP_NOEPI
        if (c == 0)
#endif
        for (int j = 0; j < 32; ++j) acc[c0 + j] += __uint_as_float(v[j]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
    }
#pragma unroll
    for (int j = 0; j < BN; ++j) {
      const int col = tn * BN + j;
      // diagonal tiles keep their upper triangle and mirror it, so the result is exactly
      // symmetric (the two triangles accumulate the hi/lo cross terms in a different order)
      if (row < pr.n && col < pr.n && (!diag || col >= row)) {
        pr.out[(long long)row * pr.ldo + col] = acc[j];
        if (col != row) pr.out[(long long)col * pr.ldo + row] = acc[j];
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(TMEM_COLS)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------------------
// General batched NT product on the tensor cores:  C_p (M x 128) = A_p (M x K) B_p^T,
// B_p (128 x K), both operands K-major (row-major with K contiguous).  One CTA per 128-row
// tile of A.  terms == 1: single-pass TF32 straight from the fp32 operands (the tensor core
// drops the low 13 mantissa bits) -- used where the product only has to be directionally
// right (early subspace iterations, which are self-correcting); terms == 3: 3xTF32 on
// pre-split hi/lo operands, drained every `chunk` k-blocks into fp32 registers (fp32-class).
// maps: per problem 5 tensor maps {A, A_hi, A_lo, B_hi (or B itself for terms == 1), B_lo}.
// ---------------------------------------------------------------------------------------
constexpr int GT_MAPS = 5;
constexpr int GT_MAX_STAGES = 6;

__global__ void __launch_bounds__(TC_THREADS, 1)
k_gemm_tc_nt(const CUtensorMap* __restrict__ maps, float* __restrict__ C, int ldc, long long strideC,
             int M, int k_total, int terms, int chunk) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int prob = blockIdx.z;
  const int tm = blockIdx.x;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  const int stage_bytes = (terms == 1 ? 2 : 4) * TILE_BYTES;
  const int nstages = terms == 1 ? GT_MAX_STAGES : 3;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + GT_MAX_STAGES * 2 * TILE_BYTES);
  uint64_t* empty = full + GT_MAX_STAGES;
  uint64_t* tmem_full = empty + GT_MAX_STAGES;     // [ACC_STAGES]
  uint64_t* tmem_empty = tmem_full + ACC_STAGES;   // [ACC_STAGES]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + ACC_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const CUtensorMap* pm = maps + (long long)prob * GT_MAPS;
  const int nkb = (k_total + BK - 1) / BK;
  const int nchunk = (nkb + chunk - 1) / chunk;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < GT_MAX_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < ACC_STAGES; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* st = smem + stage * stage_bytes;
        mbar_expect_tx(&full[stage], (uint32_t)stage_bytes);
        if (terms == 1) {
          tma_load_2d(st, pm + 0, kb * BK, tm * BM, &full[stage]);
          tma_load_2d(st + TILE_BYTES, pm + 3, kb * BK, 0, &full[stage]);
        } else {
          tma_load_2d(st, pm + 1, kb * BK, tm * BM, &full[stage]);
          tma_load_2d(st + TILE_BYTES, pm + 2, kb * BK, tm * BM, &full[stage]);
          tma_load_2d(st + 2 * TILE_BYTES, pm + 3, kb * BK, 0, &full[stage]);
          tma_load_2d(st + 3 * TILE_BYTES, pm + 4, kb * BK, 0, &full[stage]);
        }
        if (++stage == nstages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    {   // whole warp, converged; one elected lane issues (umma_tf32_w)
      const uint32_t idesc = idesc_tf32(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      for (int c = 0; c < nchunk; ++c) {
        const int as = c & 1;
        const uint32_t aphase = (uint32_t)(c >> 1) & 1u;
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tacc = tmem_base + (uint32_t)(as * BN);
        const int kb_end = min(nkb, (c + 1) * chunk);
        for (int kb = c * chunk; kb < kb_end; ++kb) {
          mbar_wait(&full[stage], phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_u32(smem + stage * stage_bytes);
          if (terms == 1) {
            const uint64_t a = make_smem_desc(sa), b = make_smem_desc(sa + TILE_BYTES);
#pragma unroll
            for (int k = 0; k < BK / 8; ++k) {
              const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);
              umma_tf32_w(tacc, a + adv, b + adv, idesc, (kb != c * chunk || k != 0) ? 1u : 0u);
            }
          } else {
            const uint64_t a_hi = make_smem_desc(sa), a_lo = make_smem_desc(sa + TILE_BYTES);
            const uint64_t b_hi = make_smem_desc(sa + 2 * TILE_BYTES);
            const uint64_t b_lo = make_smem_desc(sa + 3 * TILE_BYTES);
#pragma unroll
            for (int k = 0; k < BK / 8; ++k) {
              const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);
              umma_tf32_w(tacc, a_lo + adv, b_hi + adv, idesc, (kb != c * chunk || k != 0) ? 1u : 0u);
              umma_tf32_w(tacc, a_hi + adv, b_lo + adv, idesc, 1u);
              umma_tf32_w(tacc, a_hi + adv, b_hi + adv, idesc, 1u);
            }
          }
          umma_commit_w(&empty[stage]);
          if (++stage == nstages) { stage = 0; phase ^= 1; }
        }
        umma_commit_w(&tmem_full[as]);
      }
    }
  } else {
    const int quad = warp & 3;
    const int row = tm * BM + quad * 32 + lane;
    float acc[BN];
#pragma unroll
    for (int j = 0; j < BN; ++j) acc[j] = 0.f;
    for (int c = 0; c < nchunk; ++c) {
      const int as = c & 1;
      const uint32_t aphase = (uint32_t)(c >> 1) & 1u;
      mbar_wait(&tmem_full[as], aphase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * BN + c0), v);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[c0 + j] += __uint_as_float(v[j]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
    }
    if (row < M) {
      float4* dst = reinterpret_cast<float4*>(C + (long long)prob * strideC + (long long)row * ldc);
#pragma unroll
      for (int j = 0; j < BN; j += 4) dst[j >> 2] = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(TMEM_COLS)
                 : "memory");
  }
}

// src (rows x cols, ld_src) -> dst_hi / dst_lo (cols x rows, ld_dst): transpose + tf32 split
// (lo may be NULL: plain transpose).  grid (ceil(cols/32), ceil(rows/32), nprob), 32 x 8 threads
__global__ void __launch_bounds__(256)
k_transpose_split(const float* __restrict__ src, int ld_src, long long stride_src,
                  float* __restrict__ hi, float* __restrict__ lo, int ld_dst, long long stride_dst,
                  int rows, int cols) {
  __shared__ float tile[32][33];
  const float* s = src + (long long)blockIdx.z * stride_src;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < rows && c < cols) ? s[(long long)r * ld_src + c] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;       // output row = source column
    if (c < cols && r < rows) {
      const float x = tile[tx][i];
      const long long o = (long long)blockIdx.z * stride_dst + (long long)c * ld_dst + r;
      if (lo) {
        float h, l;
        split_tf32(x, h, l);
        hi[o] = h;
        lo[o] = l;
      } else {
        hi[o] = x;
      }
    }
  }
}

__global__ void __launch_bounds__(256)
k_split_flat2(const float* __restrict__ src, float* __restrict__ hi, float* __restrict__ lo,
              long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    const float4 x = *reinterpret_cast<const float4*>(src + i);
    float4 h, l;
    split_tf32(x.x, h.x, l.x);
    split_tf32(x.y, h.y, l.y);
    split_tf32(x.z, h.z, l.z);
    split_tf32(x.w, h.w, l.w);
    *reinterpret_cast<float4*>(hi + i) = h;
    *reinterpret_cast<float4*>(lo + i) = l;
  }
}

// all problems in one launch: grid (column chunks, row phases, problem); optional centring
// (x - mu[col]) fused into the split, so the centred matrix is never written in fp32
__global__ void __launch_bounds__(256)
k_split_tf32_batched(const TcProb* __restrict__ probs, int ncols) {
  const TcProb pr = probs[blockIdx.z];
  const int nvec = ncols >> 2;
  for (int r = blockIdx.y; r < pr.n; r += gridDim.y) {
    const float4* s = reinterpret_cast<const float4*>(pr.src + (long long)r * pr.lda);
    float4* h = reinterpret_cast<float4*>(pr.hi + (long long)r * pr.lda);
    float4* l = reinterpret_cast<float4*>(pr.lo + (long long)r * pr.lda);
    const float4* m = reinterpret_cast<const float4*>(pr.mu);
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nvec; c += gridDim.x * blockDim.x) {
      float4 x = s[c];
      if (m) {
        const float4 mm = m[c];
        x.x -= mm.x; x.y -= mm.y; x.z -= mm.z; x.w -= mm.w;
      }
      float4 xh, xl;
      split_tf32(x.x, xh.x, xl.x);
      split_tf32(x.y, xh.y, xl.y);
      split_tf32(x.z, xh.z, xl.z);
      split_tf32(x.w, xh.w, xl.w);
      h[c] = xh;
      l[c] = xl;
    }
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

// Profiling hook: an event recorded between the hi/lo split and the MMA kernel of the next
// cpsd_gram_nt_tc* calls (NULL: off), so that the two kernels can be timed separately with CUDA
// events outside a profiler.  Not thread-safe: set by the one profiling caller only.
static cudaEvent_t g_probe_event = nullptr;
extern "C" int cpsd_gram_nt_tc_probe(void* event) {
  g_probe_event = reinterpret_cast<cudaEvent_t>(event);
  return CPSD_OK;
}

// Symmetric Gram of row-major fp32 matrices on the tensor cores.
//   descs_host : HOST array of records (A == B, sym = 1, k % 4 == 0, lda % 4 == 0 required)
//   split_ws   : device workspace, >= 2 * sum_p m_p * lda_p floats (hi / lo copies)
//   map_ws     : device workspace, >= nprob * (2 * 128 + 16) bytes, 64-byte aligned
//   stage_host : pinned host staging of the same size as map_ws
static int gram_nt_tc_impl(const cpsd_gram_nt_desc* descs_host, int nprob, int m_max, int n_max,
                           float* split_ws, long long split_ws_elems, void* map_ws,
                           void* stage_host, const float* mu, int ldmu, cudaStream_t stream) {
  CPSD_CHECK_ARG(nprob >= 0 && m_max > 0 && n_max == m_max, "gram_nt_tc: bad dims");
  if (nprob == 0) return CPSD_OK;
  CPSD_CHECK_ARG(nprob <= 65535, "gram_nt_tc: nprob > 65535");
  EncodeFn enc = get_encode();
  if (!enc) {
    cpsd_set_error("gram_nt_tc: cuTensorMapEncodeTiled entry point unavailable");
    return CPSD_ERR_CUDA;
  }
  CUtensorMap* maps_h = reinterpret_cast<CUtensorMap*>(stage_host);
  TcProb* probs_h = reinterpret_cast<TcProb*>(reinterpret_cast<uint8_t*>(stage_host) +
                                              (size_t)nprob * 2 * sizeof(CUtensorMap));
  long long used = 0;
  int k_total = descs_host[0].k;
  int rows_max = 0;
  for (int p = 0; p < nprob; ++p) {
    const cpsd_gram_nt_desc& d = descs_host[p];
    CPSD_CHECK_ARG(d.A == d.B && d.sym == 1 && d.m == d.n, "gram_nt_tc: symmetric problems only");
    CPSD_CHECK_ARG((d.k & 3) == 0 && (d.lda & 3) == 0 && d.k == k_total,
                   "gram_nt_tc: k and lda must be multiples of 4 and k uniform over the batch");
    CPSD_CHECK_ARG((((uintptr_t)d.A) & 15) == 0, "gram_nt_tc: operand must be 16-byte aligned");
    CPSD_CHECK_ARG(d.alpha == 1.0f, "gram_nt_tc: alpha must be 1");
    const long long elems = (long long)d.m * d.lda;
    CPSD_CHECK_ARG(used + 2 * elems <= split_ws_elems, "gram_nt_tc: split workspace too small");
    float* hi = split_ws + used;
    float* lo = hi + elems;
    used += 2 * elems;
    const cuuint64_t gdim[2] = {(cuuint64_t)d.k, (cuuint64_t)d.m};
    const cuuint64_t gstr[1] = {(cuuint64_t)d.lda * 4};
    const cuuint32_t box[2] = {BK, BM};
    const cuuint32_t estr[2] = {1, 1};
    float* srcs[2] = {hi, lo};
    for (int u = 0; u < 2; ++u) {
      CUresult r = enc(&maps_h[2 * p + u], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, srcs[u], gdim, gstr,
                       box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        cpsd_set_error("gram_nt_tc: cuTensorMapEncodeTiled failed");
        return CPSD_ERR_CUDA;
      }
    }
    probs_h[p].out = d.out;
    probs_h[p].src = d.A;
    probs_h[p].mu = mu ? mu + (long long)p * ldmu : nullptr;
    probs_h[p].hi = hi;
    probs_h[p].lo = lo;
    probs_h[p].n = d.m;
    probs_h[p].ldo = d.ldo;
    probs_h[p].lda = d.lda;
    probs_h[p].pad_ = 0;
    if (d.m > rows_max) rows_max = d.m;
  }
  const size_t bytes = (size_t)nprob * (2 * sizeof(CUtensorMap) + sizeof(TcProb));
  CPSD_CUDA(cudaMemcpyAsync(map_ws, stage_host, bytes, cudaMemcpyHostToDevice, stream));
  const CUtensorMap* maps_d = reinterpret_cast<const CUtensorMap*>(map_ws);
  const TcProb* probs_d = reinterpret_cast<const TcProb*>(reinterpret_cast<uint8_t*>(map_ws) +
                                                          (size_t)nprob * 2 * sizeof(CUtensorMap));
  {
    int bx = (k_total / 4 + 255) / 256;
    if (bx > 4) bx = 4;
    int by = rows_max < 64 ? rows_max : 64;
    k_split_tf32_batched<<<dim3(bx, by, nprob), 256, 0, stream>>>(probs_d, k_total);
    CPSD_LAUNCH_CHECK();
    if (g_probe_event) CPSD_CUDA(cudaEventRecord(g_probe_event, stream));
  }
  const size_t smem = (size_t)STAGES * STAGE_BYTES + 1024 + 128;
  CPSD_CUDA(cudaFuncSetAttribute(k_gram_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int tiles = (m_max + BM - 1) / BM;
  k_gram_tc<<<dim3(tiles, tiles, nprob), TC_THREADS, smem, stream>>>(maps_d, probs_d, k_total);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

extern "C" int cpsd_gram_nt_tc(const cpsd_gram_nt_desc* descs_host, int nprob, int m_max, int n_max,
                               float* split_ws, long long split_ws_elems, void* map_ws,
                               void* stage_host, cudaStream_t stream) {
  return gram_nt_tc_impl(descs_host, nprob, m_max, n_max, split_ws, split_ws_elems, map_ws, stage_host,
                         nullptr, 0, stream);
}

// Gram of the CENTRED rows: mu (nprob x ldmu) row p is subtracted from every row of problem p
// inside the hi/lo split (sklearn PCA centring, _pca.py, fused: the centred matrix is never
// written in fp32).
extern "C" int cpsd_gram_nt_tc_centered(const cpsd_gram_nt_desc* descs_host, int nprob, int m_max,
                                        int n_max, float* split_ws, long long split_ws_elems,
                                        void* map_ws, void* stage_host, const float* mu, int ldmu,
                                        cudaStream_t stream) {
  CPSD_CHECK_ARG(mu != nullptr && (ldmu & 3) == 0 && (((uintptr_t)mu) & 15) == 0,
                 "gram_nt_tc_centered: mu must be 16-byte aligned with ldmu % 4 == 0");
  return gram_nt_tc_impl(descs_host, nprob, m_max, n_max, split_ws, split_ws_elems, map_ws, stage_host,
                         mu, ldmu, stream);
}

// ---- tensor-core K Q of the top-k subspace iteration (subspace.cu) -------------------------
// Layout of tc_ws (floats): K_hi, K_lo (nprob * n_pad^2 each), Qt_hi, Qt_lo (nprob * 128 * n_pad
// each).  maps: nprob * 5 tensor maps (device, 64-byte aligned), encoded on the host into
// stage_host (pinned, same size) by cpsd_topk_tc_encode whenever K / tc_ws move.
extern "C" long long cpsd_topk_tc_ws_elems(int n_pad, int nprob) {
  return (long long)nprob * (2LL * n_pad * n_pad + 2LL * 128 * n_pad);
}
extern "C" int cpsd_topk_tc_map_bytes(int nprob) { return nprob * GT_MAPS * (int)sizeof(CUtensorMap) + 64; }

extern "C" int cpsd_topk_tc_encode(const float* K, int ld, long long stride, int n_pad, int nprob,
                                   float* tc_ws, void* map_dev, void* stage_host, cudaStream_t stream) {
  CPSD_CHECK_ARG(n_pad > 0 && n_pad % 128 == 0 && ld % 4 == 0, "topk_tc_encode: bad dims");
  EncodeFn enc = get_encode();
  if (!enc) {
    cpsd_set_error("topk_tc_encode: cuTensorMapEncodeTiled entry point unavailable");
    return CPSD_ERR_CUDA;
  }
  CUtensorMap* mh = reinterpret_cast<CUtensorMap*>(stage_host);
  const long long nn = (long long)n_pad * n_pad, qn = 128LL * n_pad;
  float* Khi = tc_ws;
  float* Klo = Khi + (long long)nprob * nn;
  float* Qhi = Klo + (long long)nprob * nn;
  float* Qlo = Qhi + (long long)nprob * qn;
  const cuuint32_t box[2] = {BK, BM};
  const cuuint32_t estr[2] = {1, 1};
  for (int p = 0; p < nprob; ++p) {
    float* bases[GT_MAPS] = {const_cast<float*>(K) + (long long)p * stride, Khi + p * nn, Klo + p * nn,
                             Qhi + p * qn, Qlo + p * qn};
    const cuuint64_t rows[GT_MAPS] = {(cuuint64_t)n_pad, (cuuint64_t)n_pad, (cuuint64_t)n_pad, 128, 128};
    const cuuint64_t lds[GT_MAPS] = {(cuuint64_t)ld, (cuuint64_t)n_pad, (cuuint64_t)n_pad,
                                     (cuuint64_t)n_pad, (cuuint64_t)n_pad};
    for (int u = 0; u < GT_MAPS; ++u) {
      const cuuint64_t gdim[2] = {(cuuint64_t)n_pad, rows[u]};
      const cuuint64_t gstr[1] = {lds[u] * 4};
      CUresult r = enc(&mh[p * GT_MAPS + u], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, bases[u], gdim, gstr,
                       box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        cpsd_set_error("topk_tc_encode: cuTensorMapEncodeTiled failed");
        return CPSD_ERR_CUDA;
      }
    }
  }
  CPSD_CUDA(cudaMemcpyAsync(map_dev, stage_host, (size_t)nprob * GT_MAPS * sizeof(CUtensorMap),
                            cudaMemcpyHostToDevice, stream));
  return CPSD_OK;
}

// K -> K_hi / K_lo (once per eigen-solve, after the padding has been zeroed)
extern "C" int cpsd_topk_tc_split_k(const float* K, int ld, long long stride, int n_pad, int nprob,
                                    float* tc_ws, cudaStream_t stream) {
  CPSD_CHECK_ARG(ld == n_pad && stride == (long long)n_pad * n_pad,
                 "topk_tc_split_k: K must be densely packed (ld = n_pad)");
  const long long n = (long long)nprob * n_pad * n_pad;
  float* Khi = tc_ws;
  float* Klo = Khi + n;
  long long nb = (n / 4 + 255) / 256;
  if (nb > 148 * 16) nb = 148 * 16;
  k_split_flat2<<<(int)nb, 256, 0, stream>>>(K, Khi, Klo, n);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

// Y (n_pad x 128 per problem, row stride 128, problem stride strideY) = K Q with Q given as
// (n_pad x 128, row stride 128, problem stride strideQ): transposes (and for terms == 3 splits)
// Q into the workspace, then runs the tensor-core product.
extern "C" int cpsd_topk_tc_kq(const float* Q, long long strideQ, float* Y, long long strideY, int n_pad,
                               int nprob, int terms, float* tc_ws, const void* map_dev,
                               cudaStream_t stream) {
  CPSD_CHECK_ARG(terms == 1 || terms == 3, "topk_tc_kq: terms must be 1 or 3");
  CPSD_CHECK_ARG(n_pad > 0 && n_pad % 128 == 0 && nprob > 0 && nprob <= 65535, "topk_tc_kq: bad dims");
  const long long nn = (long long)n_pad * n_pad, qn = 128LL * n_pad;
  float* Qhi = tc_ws + 2LL * nprob * nn;
  float* Qlo = Qhi + (long long)nprob * qn;
  k_transpose_split<<<dim3(4, n_pad / 32, nprob), 256, 0, stream>>>(
      Q, 128, strideQ, Qhi, terms == 3 ? Qlo : nullptr, n_pad, qn, n_pad, 128);
  CPSD_LAUNCH_CHECK();
  const size_t smem = (size_t)GT_MAX_STAGES * 2 * TILE_BYTES + 1024 + 256;
  CPSD_CUDA(cudaFuncSetAttribute(k_gemm_tc_nt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_gemm_tc_nt<<<dim3(n_pad / BM, 1, nprob), TC_THREADS, smem, stream>>>(
      reinterpret_cast<const CUtensorMap*>(map_dev), Y, 128, strideY, n_pad, n_pad, terms,
      terms == 1 ? (n_pad / BK) : CHUNK);
  CPSD_LAUNCH_CHECK();
  return CPSD_OK;
}

// bytes of map_ws / stage_host needed for nprob problems
extern "C" int cpsd_gram_nt_tc_ws_bytes(int nprob) {
  return nprob * (int)(2 * sizeof(CUtensorMap) + sizeof(TcProb)) + 64;
}
