// gemm_tc.cuh — prefill linear layers on the 5th-generation tensor cores (tcgen05 / UMMA), sm_100a.
//
//   C[M,N] = epilogue( A[M,K] · W[N,K]^T )      A, W, C in D (bf16 | f16), fp32 accumulation in TMEM
//
// Replaces, for seq > 1, the reference's `linear_forward` (backends/mod.rs:206-241 -> candle matmul -> cuBLAS
// gemm) for the four projections of a block (attention.rs:162-164,354; mlp.rs:22,30) with the epilogues the
// reference runs as separate tensor ops fused in: +bias (->D), +residual (transformer.rs:123,131),
// silu(gate)*up on the row-interleaved Wgu (mlp.rs:22-28).  Rounding points as everywhere: the fp32
// accumulator is rounded to D before any epilogue op.
//
// Structure (one persistent CTA per SM, 256 threads, warp-specialised):
//   warp 0   TMA producer: cp.async.bulk.tensor.2d of a 128x64 A tile and a 128x64 W tile per k-block into a
//            6-stage shared-memory ring, 128-byte swizzle, completion on full[] mbarriers
//   warp 1   MMA issuer: one thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=128, K=16) x4 per
//            k-block from shared-memory descriptors into one of two TMEM accumulators (128 lanes x 128
//            fp32 columns each); tcgen05.commit releases the smem stage / publishes the accumulator
//   warp 2   TMEM allocator (256 columns)
//   warps 4-7  epilogue: tcgen05.ld 32x32b.x32 (thread = accumulator row), fused epilogue, 16-byte global
//            stores; overlaps with the MMAs of the next tile through the second accumulator
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace cake {

// Tile: 128 x BN with BN in {128, 256}.  A 128x128 tile needs 32 KB of operands per 2.1 MFLOP k-block = 64 flop/B, i.e.
// ~240 GB/s per SM at the tensor peak — more than one SM pulls out of L2, which is what capped round 1 at ~55 % tensor-pipe
// activity.  128x256 (all 512 TMEM columns: two 128x256 fp32 accumulators) moves 48 KB per 4.2 MFLOP = 87 flop/B.
constexpr int TC_BM = 128, TC_BK = 64, TC_THREADS = 256;
constexpr int TC_BN = 128;                     // granularity the host checks N against (BN = 256 when N % 256 == 0)
template <int BN> struct TcCfg {
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int STAGE_BYTES = (TC_BM + BN) * TC_BK * 2;  // 32 KB | 48 KB
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};
constexpr int TC_SMEM_BYTES = TcCfg<256>::SMEM_BYTES > TcCfg<128>::SMEM_BYTES ? TcCfg<256>::SMEM_BYTES : TcCfg<128>::SMEM_BYTES;
enum { TCE_PLAIN = 0, TCE_RESIDUAL = 1, TCE_SWIGLU = 2 };

struct TcParams {
  const void *bias;      // [N] D or nullptr (TCE_PLAIN)
  const void *residual;  // [M,N] D (TCE_RESIDUAL)
  void *C;               // [M,N] D  ([M,N/2] for TCE_SWIGLU)
  int M, N, K;
  int act;  // TCE_SWIGLU: 0 silu, 1 gelu_tanh
};

// Tile rasterisation: groups of TC_GM m-tiles (16 x 128 rows = 16 MB of A at K=4096) are walked n-major, so the A
// group stays in the 126 MB L2 while each W tile streams once per group.  With plain m-fastest order a 1 GB A
// (bs=32 x 4k) is re-read for every W tile column: measured 815 TFLOP/s, HBM-bound (profiles/README.md).
constexpr int TC_GM = 16;
__device__ __forceinline__ void tc_tile(int t, int tiles_m, int tiles_n, int bn, int &m0, int &n0) {
  const int per_group = TC_GM * tiles_n;
  const int mg = t / per_group, r = t % per_group;
  const int gm = min(TC_GM, tiles_m - mg * TC_GM);
  m0 = (mg * TC_GM + r % gm) * TC_BM;
  n0 = (r / gm) * bn;
}

// ---- PTX wrappers ------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// the same load delivered to the same shared-memory offset of every CTA in cta_mask (and signalling each one's mbarrier)
__device__ __forceinline__ void tma_load_2d_mc(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint64_t *bar, uint16_t cta_mask) {  // arrive on the barrier at this offset in every CTA of the mask
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_c, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_c),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor, K-major operand, 128-byte swizzle, tile rows of 64 D elements (128 B):
// start address >> 4 | LBO = 1 (unused for swizzled K-major) | SBO = 1024 B (8 rows x 128 B) >> 4 |
// version 1 (sm_100) | layout type 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t tc_smem_desc(const void *tile) {
  const uint64_t addr = (uint64_t)(smem_u32(tile) & 0x3ffff) >> 4;
  return addr | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor, kind::f16: D = f32, A/B = bf16 (1) or f16 (0), both K-major, N >> 3, M >> 4
template <typename T> __device__ __forceinline__ uint32_t tc_idesc(int bn);
template <> __device__ __forceinline__ uint32_t tc_idesc<__nv_bfloat16>(int bn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}
template <> __device__ __forceinline__ uint32_t tc_idesc<__half>(int bn) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}

// MC (BN == 256 only, launched as clusters of 2 CTAs, M % 256 == 0): the two CTAs of a cluster work on vertically adjacent
// 128-row tiles of the SAME 256-column panel and share its W tile — each loads one 128-row half of it and multicasts the
// half into both CTAs' shared memory, so a CTA pulls 16 KB (A) + 16 KB (its W half) instead of 48 KB per k-block out of L2:
// the operand traffic of a 2-CTA MMA without giving up the independent cta_group::1 accumulators.  A stage may be refilled
// only when BOTH CTAs have consumed it (the peer's multicast lands in our buffer): every MMA commit arrives on the empty
// barrier of both CTAs (count 2).
template <typename T, int EPI, int BN, bool MC = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const TcParams p) {
  static_assert(!MC || BN == 256, "multicast variant is built for 128 x 256 tiles");
  extern __shared__ unsigned char smem_raw_tc[];
  unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw_tc) + 1023) & ~(uintptr_t)1023);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + TcCfg<BN>::STAGES * TcCfg<BN>::STAGE_BYTES);
  uint64_t *empty = full + TcCfg<BN>::STAGES;
  uint64_t *acc_full = empty + TcCfg<BN>::STAGES;   // [2]
  uint64_t *acc_empty = acc_full + 2;       // [2]
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (p.M + TC_BM - 1) / TC_BM, tiles_n = p.N / BN;
  const int kblocks = p.K / TC_BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TcCfg<BN>::STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], MC ? 2 : 1); }
    for (int a = 0; a < 2; a++) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 128); }
    mbar_fence_init();
  }
  if (warp == 2) {  // TMEM: 2 x BN columns = two 128 x BN fp32 accumulators
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(2 * BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (MC) cluster_sync_all();  // the peer's barriers exist before anything is multicast to it
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t tmem_base = *tmem_ptr;
  // work distribution: non-MC: one tile per CTA and round; MC: one (256-row, 256-column) unit per cluster and round
  const uint32_t crank = MC ? cluster_rank() : 0u;
  const int first = MC ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, stride = MC ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int units_m = MC ? tiles_m / 2 : tiles_m, n_units = units_m * tiles_n;
  auto unit_tile = [&](int u, int &m0, int &n0) {
    tc_tile(u, units_m, tiles_n, BN, m0, n0);   // m0 in units of TC_BM rows of the unit grid
    if (MC) m0 = m0 * 2 + (int)crank * TC_BM;
  };

  if (warp == 0) {
    // ===================== TMA producer ==================================================================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
      int s = 0;
      uint32_t ph = 0;
      for (int t = first; t < n_units; t += stride) {
        int m0, n0;
        unit_tile(t, m0, n0);
        for (int kb = 0; kb < kblocks; kb++) {
          mbar_wait(&empty[s], ph ^ 1u);
          unsigned char *st = smem + (size_t)s * TcCfg<BN>::STAGE_BYTES;
          mbar_arrive_expect_tx(&full[s], TcCfg<BN>::STAGE_BYTES);
          tma_load_2d(st, &map_a, kb * TC_BK, m0, &full[s]);
          if (MC)  // map_b has a 128-row box here: our half of the W tile, delivered to both CTAs
            tma_load_2d_mc(st + TC_BM * TC_BK * 2 + crank * (128 * TC_BK * 2), &map_b, kb * TC_BK, n0 + (int)crank * 128, &full[s], (uint16_t)3);
          else
            tma_load_2d(st + TC_BM * TC_BK * 2, &map_b, kb * TC_BK, n0, &full[s]);
          if (++s == TcCfg<BN>::STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =======================================================
    if (lane == 0) {
      const uint32_t idesc = tc_idesc<T>(BN);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int t = first; t < n_units; t += stride, it++) {
        const int a = it & 1;
        const uint32_t aph = (uint32_t)(it >> 1) & 1u;
        mbar_wait(&acc_empty[a], aph ^ 1u);  // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_c = tmem_base + (uint32_t)(a * BN);
        for (int kb = 0; kb < kblocks; kb++) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const unsigned char *st = smem + (size_t)s * TcCfg<BN>::STAGE_BYTES;
          const uint64_t da = tc_smem_desc(st), db = tc_smem_desc(st + TC_BM * TC_BK * 2);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; k++)  // +32 bytes (2 x 16 B units) per K=16 slab inside the swizzle atom
            tc_mma_f16(tmem_c, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
          if (MC) tc_commit_mc(&empty[s], (uint16_t)3);  // ... in BOTH CTAs: the peer multicasts into this stage too
          else tc_commit(&empty[s]);  // arrives when the MMAs above have read the stage
          if (++s == TcCfg<BN>::STAGES) { s = 0; ph ^= 1u; }
        }
        tc_commit(&acc_full[a]);  // accumulator complete
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue warps (TMEM lanes 32*(warp%4) .. +32) ================================
    const int wq = warp & 3;
    const T *bias = reinterpret_cast<const T *>(p.bias);
    const T *res = reinterpret_cast<const T *>(p.residual);
    T *C = reinterpret_cast<T *>(p.C);
    int it = 0;
    for (int t = first; t < n_units; t += stride, it++) {
      const int a = it & 1;
      const uint32_t aph = (uint32_t)(it >> 1) & 1u;
      int m0, n0;
      unit_tile(t, m0, n0);
      const int row = m0 + wq * 32 + lane;
      mbar_wait(&acc_full[a], aph);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tc_ld32(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(a * BN + c0), r);
        if (row < p.M) {
          if (EPI == TCE_SWIGLU) {
            uint32_t o[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
              float v[2];
#pragma unroll
              for (int h = 0; h < 2; h++) {
                const float g = rnd<T>(__uint_as_float(r[4 * j + 2 * h])), u = rnd<T>(__uint_as_float(r[4 * j + 2 * h + 1]));
                const float sl = gate_act<T>(g, p.act);
                v[h] = sl * u;
              }
              o[j] = pack2<T>(v[0], v[1]);
            }
            uint4 *dst = reinterpret_cast<uint4 *>(C + (size_t)row * (p.N / 2) + (n0 + c0) / 2);
            dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
            dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
          } else {
            uint32_t o[16];
            const size_t off = (size_t)row * p.N + n0 + c0;
            uint32_t rr[16];  // the residual row segment as four 16-byte loads (scalar 2-byte loads made this epilogue longer
                              // than a K=4096 main loop: ncu r02, o_proj 68 % tensor-pipe active against 88-91 % elsewhere)
            if (EPI == TCE_RESIDUAL) {
              const uint4 *r4 = reinterpret_cast<const uint4 *>(res + off);
#pragma unroll
              for (int j = 0; j < 4; j++) {
                const uint4 v = r4[j];
                rr[4 * j] = v.x; rr[4 * j + 1] = v.y; rr[4 * j + 2] = v.z; rr[4 * j + 3] = v.w;
              }
            }
#pragma unroll
            for (int j = 0; j < 16; j++) {
              float v0 = rnd<T>(__uint_as_float(r[2 * j])), v1 = rnd<T>(__uint_as_float(r[2 * j + 1]));
              if (EPI == TCE_PLAIN && bias) {
                v0 = rnd<T>(v0 + DT<T>::to_f(bias[n0 + c0 + 2 * j]));
                v1 = rnd<T>(v1 + DT<T>::to_f(bias[n0 + c0 + 2 * j + 1]));
              }
              if (EPI == TCE_RESIDUAL) {
                const float2 rv = DT<T>::unpack2(rr[j]);
                v0 += rv.x;
                v1 += rv.y;
              }
              o[j] = pack2<T>(v0, v1);
            }
            uint4 *dst = reinterpret_cast<uint4 *>(C + off);
#pragma unroll
            for (int j = 0; j < 4; j++) dst[j] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[a]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (MC) cluster_sync_all();  // nobody leaves while the peer may still signal our barriers
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN) : "memory");
  }
}

}  // namespace cake
