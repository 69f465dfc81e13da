// attn_prefill_tc.cuh — causal GQA flash attention for prefill (seq > 1, head_dim 128) on the 5th-generation tensor
// cores (tcgen05 / UMMA + TMEM), sm_100a.
//
// Same operator and the same rounding points as attn_prefill_mma_kernel (attn_prefill.cuh), which restates
// cake-core/src/models/common/attention.rs:299-349 (scores = q k^T * 1/sqrt(hd) in fp32, causal mask, fp32 softmax,
// probabilities x V, one rounding to D at the end): S in fp32 from D operands, P kept to ~16 mantissa bits as
// P_hi + P_lo (two D operands), O accumulated in fp32, rounded once.  What changes is who multiplies: mma.sync tops out
// near 465 TFLOP/s on this part (ncu r02: HMMA pipe 50 % active at 231 TFLOP/s), a quarter of what tcgen05 delivers.
//
// One CTA = one 128-query tile of one head; two CTAs per SM (2 x 256 TMEM columns, 2 x 97 KB shared memory), so one CTA's
// softmax overlaps the other's MMAs.  192 threads:
//   warps 0-3  softmax: thread = query row (TMEM lane).  tcgen05.ld of the 128 x 64 S tile, scale + mask + running max /
//              sum entirely in-thread (no shuffles), P_hi / P_lo packed to D and stored with tcgen05.st over the S tile they
//              came from (the MMA takes its A operand from TMEM), lazy rescale of O in TMEM (only when the row max grew by
//              more than tau, P stays <= e^tau; the final O / l is the same quotient), epilogue O / l -> D -> global
//   warp 4     TMA producer: Q tile once, then 64-key K and V tiles through two-stage rings (3-D maps: keys past the
//              visible length and query rows past S are zero-filled by the TMA unit, nothing uninitialised reaches an MMA)
//   warp 5     TMEM allocator + MMA issuer (one thread): S = Q K^T (8 x M128 N64 K16, operands from shared memory) into one
//              of two S buffers — S of tile j+1 is issued BEFORE P V of tile j, so the softmax of j+1 runs under it —
//              and O += P_hi V + P_lo V (8 x M128 N128 K16, P from TMEM, V consumed straight from its [key][hd] cache
//              layout as an MN-major B operand)
#pragma once
#include "gemm_tc.cuh"

namespace cake {

constexpr int FT_BM = 128, FT_BN = 64, FT_HD = 128, FT_THREADS = 192;
constexpr int FT_Q_BYTES = FT_BM * FT_HD * 2;   // 32 KB: two 128-row x 64-dim boxes
constexpr int FT_KV_BYTES = FT_BN * FT_HD * 2;  // 16 KB: two 64-key x 64-dim boxes
constexpr int FT_SMEM_BYTES = FT_Q_BYTES + 4 * FT_KV_BYTES + 1024 /*align*/ + 128 /*barriers*/;  // Q + 2 K stages + 2 V stages
constexpr int FT_TMEM_COLS = 256;               // S/P buffers: columns 0..63 and 64..127 (P_hi | P_lo packed over S), O: 128..255

__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] x B[smem]: the A operand (M = lane, K packed two D values per 32-bit column) comes from tensor memory
__device__ __forceinline__ void tc_mma_f16_ts(uint32_t tmem_c, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_c),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ float ft_ex2(float x) {  // MUFU.EX2: 2^x, 2^-inf = 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// MN-major B operand (V: rows = keys = K dimension, 64 contiguous head dims = one 128-byte swizzle row):
// LBO = distance between the two 64-dim halves, SBO = 1024 B between groups of 8 keys
__device__ __forceinline__ uint64_t ft_desc_mn(const void *tile, uint32_t lbo_bytes) {
  const uint64_t addr = (uint64_t)(smem_u32(tile) & 0x3ffff) >> 4;
  return addr | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

template <typename T>
__global__ void __launch_bounds__(FT_THREADS, 2)
attn_prefill_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                       const __grid_constant__ CUtensorMap map_v, T *__restrict__ y, int S, int n_heads, int n_kv, int pos0,
                       float scale_log2, float tau_log2) {
  extern __shared__ unsigned char ft_smem_raw[];
  unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(ft_smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char *q_s = smem;
  unsigned char *k_s = q_s + FT_Q_BYTES;       // [2] stages
  unsigned char *v_s = k_s + 2 * FT_KV_BYTES;  // [2] stages
  uint64_t *bars = reinterpret_cast<uint64_t *>(v_s + 2 * FT_KV_BYTES);
  uint64_t *q_full = bars, *k_full = bars + 1, *k_empty = bars + 3, *v_full = bars + 5, *v_empty = bars + 7, *s_full = bars + 9,
           *p_full = bars + 11, *o_full = bars + 12;
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = gridDim.x - 1 - blockIdx.x;  // heavy (late) query tiles first
  const int h = blockIdx.y, b = blockIdx.z, kvh = h / (n_heads / n_kv);
  const int m0 = mt * FT_BM;
  const int kv_end = pos0 + min(S, m0 + FT_BM);  // keys visible to the last query of this tile
  const int n_tiles = (kv_end + FT_BN - 1) / FT_BN;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 13; i++) mbar_init(&bars[i], i == 11 ? 128u : 1u);
    mbar_fence_init();
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(FT_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_o = tmem_base + 128u;  // S/P buffer i at columns 64 i

  if (warp == 4) {
    // ===================== TMA producer ====================================================================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v) : "memory");
      mbar_arrive_expect_tx(q_full, FT_Q_BYTES);
      tma_load_3d(q_s, &map_q, h * FT_HD, m0, b, q_full);
      tma_load_3d(q_s + FT_Q_BYTES / 2, &map_q, h * FT_HD + 64, m0, b, q_full);
      const int slab = b * n_kv + kvh;
      for (int j = 0; j < n_tiles; j++) {
        const int st = j & 1;
        const uint32_t eph = (uint32_t)((j >> 1) - 1) & 1u;  // the stage's previous use (tile j-2) has been consumed
        unsigned char *kd = k_s + st * FT_KV_BYTES, *vd = v_s + st * FT_KV_BYTES;
        if (j >= 2) mbar_wait(&k_empty[st], eph);
        mbar_arrive_expect_tx(&k_full[st], FT_KV_BYTES);
        tma_load_3d(kd, &map_k, 0, j * FT_BN, slab, &k_full[st]);
        tma_load_3d(kd + FT_KV_BYTES / 2, &map_k, 64, j * FT_BN, slab, &k_full[st]);
        if (j >= 2) mbar_wait(&v_empty[st], eph);
        mbar_arrive_expect_tx(&v_full[st], FT_KV_BYTES);
        tma_load_3d(vd, &map_v, 0, j * FT_BN, slab, &v_full[st]);
        tma_load_3d(vd + FT_KV_BYTES / 2, &map_v, 64, j * FT_BN, slab, &v_full[st]);
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer (one thread) =========================================================
    if (lane == 0) {
      const uint32_t idesc_s = tc_idesc<T>(FT_BN);                   // M128 N64, A and B K-major
      const uint32_t idesc_o = tc_idesc<T>(FT_HD) | (1u << 16);      // M128 N128, A from TMEM, B (V) MN-major
      auto issue_s = [&](int j) {  // S_j = Q K_j^T into S buffer j&1
        const int st = j & 1;
        mbar_wait(&k_full[st], (uint32_t)(j >> 1) & 1u);
        tc_fence_after();
        const unsigned char *kt = k_s + st * FT_KV_BYTES;
#pragma unroll
        for (int k = 0; k < FT_HD / 16; k++) {  // head dims 16k..16k+15: box k/4, 32-byte slab k%4 inside the swizzle row
          const uint64_t da = tc_smem_desc(q_s + (k >> 2) * (FT_Q_BYTES / 2)) + (uint64_t)((k & 3) * 2);
          const uint64_t db = tc_smem_desc(kt + (k >> 2) * (FT_KV_BYTES / 2)) + (uint64_t)((k & 3) * 2);
          tc_mma_f16(tmem_base + (uint32_t)(st * 64), da, db, idesc_s, k ? 1u : 0u);
        }
        tc_commit(&k_empty[st]);
        tc_commit(&s_full[st]);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < n_tiles; j++) {
        const int st = j & 1;
        // S of the next tile first: its buffer held P of tile j-1, whose P V is already queued ahead of it in the pipe
        if (j + 1 < n_tiles) issue_s(j + 1);
        mbar_wait(p_full, (uint32_t)j & 1u);  // P_j in TMEM (over S_j), O rescaled
        mbar_wait(&v_full[st], (uint32_t)(j >> 1) & 1u);
        tc_fence_after();
        const unsigned char *vt = v_s + st * FT_KV_BYTES;
#pragma unroll
        for (int half = 0; half < 2; half++)
#pragma unroll
          for (int ks = 0; ks < FT_BN / 16; ks++) {  // keys 16ks..16ks+15 = 8 packed columns of P_hi (0..31) or P_lo (32..63)
            const uint32_t ta = tmem_base + (uint32_t)(st * 64 + half * 32 + ks * 8);
            const uint64_t db = ft_desc_mn(vt + ks * 2048, FT_KV_BYTES / 2);
            tc_mma_f16_ts(tmem_o, ta, db, idesc_o, (j | half | ks) ? 1u : 0u);
          }
        tc_commit(&v_empty[st]);
        tc_commit(o_full);
      }
    }
  } else {
    // ===================== softmax / correction / epilogue (thread = query row) =============================
    const int row = warp * 32 + lane;
    const int qp = pos0 + m0 + row;  // absolute position of this row's query
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f;  // m_run: reference max of the raw scores (unscaled)
    for (int j = 0; j < n_tiles; j++) {
      const uint32_t tmem_sp = tmem_base + lane_addr + (uint32_t)((j & 1) * 64);
      mbar_wait(&s_full[j & 1], (uint32_t)(j >> 1) & 1u);
      tc_fence_after();
      float sv[FT_BN];
      {
        uint32_t r[32];
        tc_ld32(tmem_sp, r);
#pragma unroll
        for (int i = 0; i < 32; i++) sv[i] = __uint_as_float(r[i]);
        tc_ld32(tmem_sp + 32u, r);
#pragma unroll
        for (int i = 0; i < 32; i++) sv[32 + i] = __uint_as_float(r[i]);
      }
      // causal mask (attention.rs:314-341) — only tiles reaching past the first row's position need the compare
      float mx = -INFINITY;
      if (j * FT_BN + FT_BN - 1 > pos0 + m0) {
#pragma unroll
        for (int i = 0; i < FT_BN; i++) {
          if (j * FT_BN + i > qp) sv[i] = -INFINITY;
          mx = fmaxf(mx, sv[i]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < FT_BN; i++) mx = fmaxf(mx, sv[i]);
      }
      // running max with lazy update: keep the old reference while the new max exceeds it by at most tau
      float fac = 1.f;
      bool resc = false;
      if (j == 0) {
        m_run = mx;
      } else if ((mx - m_run) * scale_log2 > tau_log2) {
        fac = ft_ex2((m_run - mx) * scale_log2);
        m_run = mx;
        resc = true;
      }
      // p = exp((s - m) / sqrt(hd)) = 2^(s c - m c), c = log2(e) / sqrt(hd): one FFMA + one MUFU.EX2 per score; 2^-inf = 0 masks
      const float mc = m_run * scale_log2;
      float rs = 0.f;
#pragma unroll
      for (int i = 0; i < FT_BN; i++) {
        const float p = ft_ex2(fmaf(sv[i], scale_log2, -mc));
        sv[i] = p;
        rs += p;
      }
      l_run = l_run * fac + rs;
      if (j > 0) {
        mbar_wait(o_full, (uint32_t)(j - 1) & 1u);  // P V of the previous tile is complete: O stable
        tc_fence_after();
        if (__any_sync(0xffffffffu, resc)) {
#pragma unroll 1
          for (int c0 = 0; c0 < FT_HD; c0 += 32) {
            uint32_t r[32];
            tc_ld32(tmem_o + lane_addr + (uint32_t)c0, r);
#pragma unroll
            for (int i = 0; i < 32; i++) r[i] = __float_as_uint(__uint_as_float(r[i]) * fac);
            tc_st32(tmem_o + lane_addr + (uint32_t)c0, r);
          }
        }
      }
      // P = P_hi + P_lo, both in D, packed two keys per column over the S tile: columns 0..31 P_hi, 32..63 P_lo
      {
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int i = 0; i < 32; i++) {
          const float p0 = sv[2 * i], p1 = sv[2 * i + 1];
          hi[i] = pack2x<T>(p0, p1);
          const float2 hf = DT<T>::unpack2(hi[i]);
          lo[i] = pack2x<T>(p0 - hf.x, p1 - hf.y);
        }
        tc_st32(tmem_sp, hi);
        tc_st32(tmem_sp + 32u, lo);
      }
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(p_full);
    }
    // ---- epilogue: O / l, one rounding to D (attention.rs:346), store (b, s, h, hd) ---------------------------
    mbar_wait(o_full, (uint32_t)(n_tiles - 1) & 1u);
    tc_fence_after();
    const int t = m0 + row;
    const float inv = 1.0f / l_run;
    T *dst = y + ((size_t)(b * S + t) * n_heads + h) * FT_HD;
#pragma unroll 1
    for (int c0 = 0; c0 < FT_HD; c0 += 32) {
      uint32_t r[32];
      tc_ld32(tmem_o + lane_addr + (uint32_t)c0, r);
      if (t < S) {
        uint32_t o[16];
#pragma unroll
        for (int i = 0; i < 16; i++) o[i] = pack2x<T>(__uint_as_float(r[2 * i]) * inv, __uint_as_float(r[2 * i + 1]) * inv);
        uint4 *d4 = reinterpret_cast<uint4 *>(dst + c0);
#pragma unroll
        for (int i = 0; i < 4; i++) d4[i] = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(FT_TMEM_COLS) : "memory");
  }
}

}  // namespace cake
