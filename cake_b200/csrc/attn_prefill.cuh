// attn_prefill.cuh — causal grouped-query attention for seq > 1 (prompt ingestion) on the tensor cores.
//
// Replaces the reference's generic branch (attention.rs:300-346): to_dtype(F32) x3, repeat_kv, q@k^T / sqrt(hd),
// masked_fill(j-(T-S) > i, -inf), softmax, @v, to_dtype(D) — ~8 launches and an S x T score matrix in HBM — with one
// flash-style kernel: S never leaves registers.  Numerics stay those of the f32 branch:
//   * q,k are D values, so q.k products are exact in fp32; mma.sync (D inputs, fp32 accumulate) == f32 matmul up
//     to summation order;
//   * softmax in fp32 (max-subtracted, online rescaling; __expf = MUFU ex2, ~2 ulp fp32);
//   * P (fp32) is split into two D halves, P = P_hi + P_lo, and both are multiplied with V on the tensor cores:
//     ~16 mantissa bits, far below the final rounding to D (an FA2-style single bf16 P would lose 8 bits);
//   * one rounding of the result to D.
// Layout: grid (ceil(S/64), n_heads, batch), 4 warps x 16 query rows; K/V tiles of 64 positions stream from the
// cache through a 2-stage cp.async pipeline into XOR-swizzled shared memory; ldmatrix feeds the fragments.
// A kv head's tile is shared by the G query heads through L2.  Queries sit at absolute positions pos0+t and see
// cache rows [0, pos0+t].
#pragma once
#include "common.cuh"

namespace cake {

constexpr int FA_BM = 64, FA_BN = 64, FA_THREADS = 128;

template <typename T> struct FaMma;
template <> struct FaMma<__nv_bfloat16> {
  static __device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
};
template <> struct FaMma<__half> {
  static __device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
};
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, int src_bytes) {  // src_bytes 0 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// byte offset of 16-byte chunk `c` of row `r` in a [rows][HD] tile of D, XOR-swizzled so that ldmatrix's 8 rows
// (stride HD*2 bytes, a multiple of 128) land in 8 different bank groups
template <int HD> __device__ __forceinline__ uint32_t fa_off(int r, int c) { return (uint32_t)(r * HD * 2 + ((c ^ (r & 7)) << 4)); }

// Occupancy (ncu r02: 189 registers + 80 KB -> 2 CTAs = 2 warps per sub-partition, HMMA pipe 50 % active, issue-bound on
// dependent MMA chains): the Q tile is only needed until its fragments sit in registers, so it is staged in the second V
// buffer (64 KB per CTA instead of 80), and the register budget is capped for 3 CTAs per SM.
template <typename T, int HD>
__global__ void __launch_bounds__(FA_THREADS, (HD == 128) ? 3 : 4)
attn_prefill_mma_kernel(const T *__restrict__ qkv, const T *__restrict__ kcache, const T *__restrict__ vcache, T *__restrict__ y,
                        int S, int n_heads, int n_kv, int cap, int pos0, float scale) {
  constexpr int CH = HD / 8;        // 16-byte chunks per row
  constexpr int KS = HD / 16;       // k-steps of the QK^T contraction
  constexpr int NT = FA_BN / 8;     // 8-key n-tiles per KV tile
  constexpr int DT = HD / 8;        // 8-dim n-tiles of the output
  constexpr int TILE_B = FA_BN * HD * 2;
  extern __shared__ __align__(128) unsigned char fa_smem[];
  unsigned char *k_s = fa_smem;                   // [2][64][HD]
  unsigned char *v_s = k_s + 2 * TILE_B;          // [2][64][HD]
  unsigned char *q_s = v_s + TILE_B;              // [64][HD]: aliases V buffer 1, which is first filled after the Q fragments are read
  pdl_launch_dependents();
  pdl_wait();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = gridDim.x - 1 - blockIdx.x;  // heavy (late) query tiles first
  const int h = blockIdx.y, b = blockIdx.z, G = n_heads / n_kv;
  const int m0 = mt * FA_BM, nh_all = n_heads + 2 * n_kv;
  const T *kc = kcache + ((size_t)b * n_kv + h / G) * cap * HD;
  const T *vc = vcache + ((size_t)b * n_kv + h / G) * cap * HD;
  const int kv_end = pos0 + min(S, m0 + FA_BM);  // keys visible to the last query of this tile
  const int n_tiles = (kv_end + FA_BN - 1) / FA_BN;

  auto load_kv = [&](int j, int buf) {
    for (int i = threadIdx.x; i < FA_BN * CH; i += FA_THREADS) {
      const int r = i / CH, c = i % CH, key = j * FA_BN + r;
      const bool ok = key < kv_end;
      const size_t g = (size_t)(ok ? key : 0) * HD + c * 8;
      cp_async16(smem_u32(k_s + buf * TILE_B) + fa_off<HD>(r, c), kc + g, ok ? 16 : 0);
      cp_async16(smem_u32(v_s + buf * TILE_B) + fa_off<HD>(r, c), vc + g, ok ? 16 : 0);
    }
  };
  // Q tile (rotated q lives in the qkv buffer, rope_append_kernel) + first KV tile
  for (int i = threadIdx.x; i < FA_BM * CH; i += FA_THREADS) {
    const int r = i / CH, c = i % CH, t = m0 + r;
    const bool ok = t < S;
    const T *src = qkv + ((size_t)(b * S + (ok ? t : 0)) * nh_all + h) * HD + c * 8;
    cp_async16(smem_u32(q_s) + fa_off<HD>(r, c), src, ok ? 16 : 0);
  }
  load_kv(0, 0);
  cp_async_commit();

  float o[DT][4];
#pragma unroll
  for (int i = 0; i < DT; i++) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  uint32_t qf[KS][4];
  const int g = lane >> 2, c2 = (lane & 3) * 2;
  const int qpos0 = pos0 + m0 + warp * 16 + g;  // absolute position of this thread's first row (second: +8)
  const int mi = lane >> 3, lr = lane & 7;

  // Q fragments first (they live in V buffer 1), then the pipeline may overwrite that buffer
  cp_async_wait<0>();
  __syncthreads();
#pragma unroll
  for (int ks = 0; ks < KS; ks++)
    ldsm_x4(qf[ks], smem_u32(q_s) + fa_off<HD>(warp * 16 + (mi & 1) * 8 + lr, ks * 2 + (mi >> 1)));
  __syncthreads();
  for (int j = 0; j < n_tiles; j++) {
    const int buf = j & 1;
    if (j + 1 < n_tiles) load_kv(j + 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    // ---- S = Q K^T (16 x 64 per warp), fp32 ------------------------------------------------------------
    float s[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; nt++) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
    const uint32_t kb_addr = smem_u32(k_s + buf * TILE_B), vb_addr = smem_u32(v_s + buf * TILE_B);
#pragma unroll
    for (int ks = 0; ks < KS; ks++) {
#pragma unroll
      for (int np = 0; np < NT / 2; np++) {
        uint32_t kf[4];
        ldsm_x4(kf, kb_addr + fa_off<HD>(np * 16 + (mi >> 1) * 8 + lr, ks * 2 + (mi & 1)));
        FaMma<T>::mma(s[2 * np], qf[ks], kf[0], kf[1]);
        FaMma<T>::mma(s[2 * np + 1], qf[ks], kf[2], kf[3]);
      }
    }
    // ---- scale, causal mask (attention.rs:314-341), online softmax -------------------------------------
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < NT; nt++)
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const int key = j * FA_BN + nt * 8 + c2 + (e & 1);
        const int qp = qpos0 + (e >> 1) * 8;
        const float v = (key <= qp) ? s[nt][e] * scale : -INFINITY;
        s[nt][e] = v;
        mx[e >> 1] = fmaxf(mx[e >> 1], v);
      }
    float fac[2], rs[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; r++) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float mn = fmaxf(m_run[r], mx[r]);
      fac[r] = (m_run[r] == -INFINITY) ? 0.f : __expf(m_run[r] - mn);
      m_run[r] = mn;
    }
#pragma unroll
    for (int nt = 0; nt < NT; nt++)
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const float mn = m_run[e >> 1];
        // MUFU ex2-based exp (~2 ulp in fp32, far below the D rounding of the result): S*T/2 of these per head
        const float p = (mn == -INFINITY) ? 0.f : __expf(s[nt][e] - mn);
        s[nt][e] = p;
        rs[e >> 1] += p;
      }
#pragma unroll
    for (int r = 0; r < 2; r++) {
      rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 1);
      rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 2);
      l_run[r] = l_run[r] * fac[r] + rs[r];
    }
#pragma unroll
    for (int dt = 0; dt < DT; dt++) {
      o[dt][0] *= fac[0]; o[dt][1] *= fac[0]; o[dt][2] *= fac[1]; o[dt][3] *= fac[1];
    }
    // ---- O += P V with P = P_hi + P_lo (both in D) -----------------------------------------------------
#pragma unroll
    for (int kb = 0; kb < FA_BN / 16; kb++) {
      uint32_t ph[4], pl[4];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const float p0 = s[2 * kb + (q >> 1)][(q & 1) * 2], p1 = s[2 * kb + (q >> 1)][(q & 1) * 2 + 1];
        const float h0 = rnd<T>(p0), h1 = rnd<T>(p1);
        ph[q] = pack2<T>(h0, h1);
        pl[q] = pack2<T>(p0 - h0, p1 - h1);
      }
#pragma unroll
      for (int dp = 0; dp < DT / 2; dp++) {
        uint32_t vf[4];
        ldsm_x4_trans(vf, vb_addr + fa_off<HD>(kb * 16 + (mi & 1) * 8 + lr, dp * 2 + (mi >> 1)));
        FaMma<T>::mma(o[2 * dp], ph, vf[0], vf[1]);
        FaMma<T>::mma(o[2 * dp], pl, vf[0], vf[1]);
        FaMma<T>::mma(o[2 * dp + 1], ph, vf[2], vf[3]);
        FaMma<T>::mma(o[2 * dp + 1], pl, vf[2], vf[3]);
      }
    }
    __syncthreads();  // everyone is done with buffer `buf` before the next iteration's prefetch overwrites it
  }
  cp_async_wait<0>();
  // ---- normalise, round once to D (attention.rs:346), store (b, s, h, hd) ---------------------------------
#pragma unroll
  for (int r = 0; r < 2; r++) {
    const int t = m0 + warp * 16 + g + r * 8;
    if (t < S) {
      const float inv = 1.0f / l_run[r];
      T *dst = y + ((size_t)(b * S + t) * n_heads + h) * HD;
#pragma unroll
      for (int dt = 0; dt < DT; dt++)
        *reinterpret_cast<uint32_t *>(dst + dt * 8 + c2) = pack2<T>(o[dt][2 * r] * inv, o[dt][2 * r + 1] * inv);
    }
  }
}

}  // namespace cake
