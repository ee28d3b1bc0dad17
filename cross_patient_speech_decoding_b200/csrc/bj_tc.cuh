// Block-Jacobi off-diagonal tile update on the 5th-generation tensor cores:
//     K[p,q]  <-  R_p^T  K[p,q]  R_q            (128 x 128 tiles, 3xTF32, fp32 accumulate)
// One CTA per tile pair p < q.  Both products run as tcgen05.mma.kind::tf32 with K-major,
// 128B-swizzled shared-memory operands that the CTA's threads build themselves (global ->
// hi/lo split -> swizzled st.shared), because the tile is a gather of four 64x64 blocks:
//   GEMM1  Y = X R_q      A = X (row-major = K-major), B = R_q^T (the inner kernels also
//                         store every rotation transposed)
//   Y: TMEM -> registers -> hi/lo split -> written TRANSPOSED into shared memory, which is
//                         exactly the K-major B operand of the second product
//   GEMM2  Z = R_p^T Y    A = R_p^T, B = Y^T
//   Z: TMEM -> registers -> global, plus the mirrored tile K[q,p] = Z^T.
// Included by jacobi.cu (shares tile_gidx / TS / BS).
#pragma once
#include "tc_common.cuh"

namespace bjtc {
using namespace tc;

constexpr int KB = 32;                       // fp32 per k-block (= 128 B swizzle row)
constexpr int OP_BYTES = 128 * KB * 4;       // one operand k-block: 128 rows x 128 B
constexpr int NKB = 128 / KB;                // 4 k-blocks per product
constexpr int SMEM_BYTES = 8 * OP_BYTES + 4 * OP_BYTES + 1024 + 128;

// 16 consecutive fp32 of row r (columns c0 .. c0+15 of the k-block) -> hi / lo operand tiles
__device__ __forceinline__ void store_split16(uint8_t* hi_tile, uint8_t* lo_tile, int r, int c0,
                                              const float4* v) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int unit = ((c0 >> 2) + i) ^ (r & 7);
    const int off = r * 128 + unit * 16;
    float4 h, l;
    split_tf32(v[i].x, h.x, l.x);
    split_tf32(v[i].y, h.y, l.y);
    split_tf32(v[i].z, h.z, l.z);
    split_tf32(v[i].w, h.w, l.w);
    *reinterpret_cast<float4*>(hi_tile + off) = h;
    *reinterpret_cast<float4*>(lo_tile + off) = l;
  }
}

__device__ __forceinline__ void issue_kblock(uint32_t tmem, uint32_t a_hi, uint32_t a_lo,
                                             uint32_t b_hi, uint32_t b_lo, uint32_t idesc,
                                             bool first) {
  const uint64_t dah = make_smem_desc(a_hi), dal = make_smem_desc(a_lo);
  const uint64_t dbh = make_smem_desc(b_hi), dbl = make_smem_desc(b_lo);
#pragma unroll
  for (int k = 0; k < KB / 8; ++k) {
    const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);
    umma_tf32(tmem, dal + adv, dbh + adv, idesc, (first && k == 0) ? 0u : 1u);
    umma_tf32(tmem, dah + adv, dbl + adv, idesc, 1u);
    umma_tf32(tmem, dah + adv, dbh + adv, idesc, 1u);
  }
}

template <typename GIDX>
__device__ __forceinline__ void update_tile_tc(float* __restrict__ Kg, int ld,
                                               const float* __restrict__ RTp,
                                               const float* __restrict__ RTq, int Ip, int Jp,
                                               int Iq, int Jq, bool mirror, uint8_t* smem_raw,
                                               GIDX gidx) {
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* regA = smem;                      // 128 KB: GEMM1 stages, later Y^T hi | lo
  uint8_t* regB = smem + 8 * OP_BYTES;       // 64 KB: GEMM2 A stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 12 * OP_BYTES);
  uint64_t* done1 = bars;                    // [2]
  uint64_t* acc1 = bars + 2;
  uint64_t* done2 = bars + 3;                // [2]
  uint64_t* acc2 = bars + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < 6; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(128u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  const uint32_t idesc = idesc_tf32(128, 128);
  const int r = tid >> 1, c0 = (tid & 1) * 16;       // this thread's row / half of a k-block
  const long long grow = (long long)gidx(r, Ip, Jp) * ld;

  // ---------------------------------------------------------------- GEMM1: Y = X R_q
  for (int kb = 0; kb < NKB; ++kb) {
    const int st = kb & 1;
    if (kb >= 2) mbar_wait(&done1[st], 0);
    uint8_t* sbase = regA + st * 4 * OP_BYTES;
    const int gcol = (kb < 2 ? Iq * 64 + kb * 32 : Jq * 64 + (kb - 2) * 32) + c0;
    float4 x[4], w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      x[i] = *reinterpret_cast<const float4*>(Kg + grow + gcol + 4 * i);
      w[i] = *reinterpret_cast<const float4*>(RTq + r * 128 + kb * 32 + c0 + 4 * i);
    }
    store_split16(sbase, sbase + OP_BYTES, r, c0, x);
    store_split16(sbase + 2 * OP_BYTES, sbase + 3 * OP_BYTES, r, c0, w);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t s0 = smem_u32(sbase);
      issue_kblock(tmem, s0, s0 + OP_BYTES, s0 + 2 * OP_BYTES, s0 + 3 * OP_BYTES, idesc, kb == 0);
      umma_commit(&done1[st]);
      if (kb == NKB - 1) umma_commit(acc1);
    }
  }
  // ---------------------------------------------------------------- Y -> Y^T (hi | lo)
  mbar_wait(acc1, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int quad = warp & 3, chalf = warp >> 2;
  {
    uint8_t* yhi = regA + quad * OP_BYTES;             // k-block = this warp's 32 rows of Y
    uint8_t* ylo = regA + 4 * OP_BYTES + quad * OP_BYTES;
#pragma unroll 1
    for (int cc = 0; cc < 64; cc += 32) {
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(chalf * 64 + cc), v);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int n = chalf * 64 + cc + j;
        float h, l;
        split_tf32(__uint_as_float(v[j]), h, l);
        const int off = n * 128 + (((lane >> 2) ^ (n & 7)) << 4) + (lane & 3) * 4;
        *reinterpret_cast<float*>(yhi + off) = h;
        *reinterpret_cast<float*>(ylo + off) = l;
      }
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  // ---------------------------------------------------------------- GEMM2: Z = R_p^T Y
  for (int kb = 0; kb < NKB; ++kb) {
    const int st = kb & 1;
    if (kb >= 2) mbar_wait(&done2[st], 0);
    uint8_t* sbase = regB + st * 2 * OP_BYTES;
    float4 w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      w[i] = *reinterpret_cast<const float4*>(RTp + r * 128 + kb * 32 + c0 + 4 * i);
    store_split16(sbase, sbase + OP_BYTES, r, c0, w);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t s0 = smem_u32(sbase);
      const uint32_t y0 = smem_u32(regA + kb * OP_BYTES);
      issue_kblock(tmem, s0, s0 + OP_BYTES, y0, y0 + 4 * OP_BYTES, idesc, kb == 0);
      umma_commit(&done2[st]);
      if (kb == NKB - 1) umma_commit(acc2);
    }
  }
  // ---------------------------------------------------------------- Z -> K[p,q], K[q,p]
  mbar_wait(acc2, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    const int m = quad * 32 + lane;
    const int gr = gidx(m, Ip, Jp);
#pragma unroll 1
    for (int cc = 0; cc < 64; cc += 32) {
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(chalf * 64 + cc), v);
      const int gc0 = gidx(chalf * 64 + cc, Iq, Jq);   // 32 consecutive global columns
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(Kg + (long long)gr * ld + gc0 + j) =
            make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                        __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
      if (mirror) {
#pragma unroll
        for (int j = 0; j < 32; ++j) Kg[(long long)(gc0 + j) * ld + gr] = __uint_as_float(v[j]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u)
                 : "memory");
  }
}

}  // namespace bjtc
