// attn_decode.cuh — one decoded token of grouped-query attention with the KV-cache append fused in.
//
// Replaces, per layer and token, the reference's: narrow q/k/v (attention.rs:166-174), QK-norm
// (:202-215), RoPE on q and k (:242-253, backends/mod.rs:444-482), Tensor::cat KV append — a full
// re-copy of the layer's cache every token (cache.rs:195-196) — to_dtype(F32) + repeat_kv + matmul +
// softmax + matmul (attention.rs:300-346).  Here the cache is pre-allocated (n_kv, max_seq, hd) in D and
// the new K/V row is written in place (algorithmic 2*n_kv*hd*es bytes instead of 2*L*n_kv*hd*es*2).
//
// Flash-decoding layout: grid = (NSPLIT, n_kv); a CTA scans a contiguous slice of the sequence for ONE
// kv head and all G = n_h/n_kv query heads that share it, so every K/V byte is read from HBM once
// (not G times as with repeat_kv).  16-byte coalesced loads, HD/8 lanes per cached row, fp32
// scores/softmax/PV exactly as the reference's f32 branch, warp-shuffle reductions.  The split
// partials (m, l, acc) are merged by the last CTA to finish (atomic ticket), in split order, so the
// result is deterministic.
#pragma once
#include "common.cuh"

namespace cake {

constexpr int ATTN_THREADS = 128;
constexpr int ATTN_TILE = 128;  // cached positions staged in shared memory per pass
constexpr int ATTN_MAX_G = 8;
constexpr int ATTN_MAX_SPLIT = 32;

struct AttnDecodeArgs {
  const void *qkv;      // [(n_h + 2 n_kv) * hd] D: fused projection output of this token
  void *kcache;         // [n_kv, cap, hd] D (this layer)
  void *vcache;
  const void *cos_t;    // [max_seq, rot/2] D
  const void *sin_t;
  const void *q_norm;   // [hd] D or nullptr
  const void *k_norm;
  void *y;              // [n_h * hd] D
  float *ws_ml;         // [n_h, nsplit, 2]
  float *ws_acc;        // [n_h, nsplit, hd]
  unsigned *counters;   // [n_kv]
  const int *d_pos;     // device-resident position of this token (== cache length before append)
  int n_heads, n_kv, cap, rot, nsplit;
  float eps, scale;
  int window;  // cache.rs:173-205 sliding window (0 = full context)
};

__host__ __device__ inline size_t attn_smem_bytes(int hd, int es) {
  return (size_t)2 * ATTN_TILE * hd * es   // K and V tiles
         + (size_t)ATTN_MAX_G * hd * 4     // q (f32)
         + (size_t)ATTN_MAX_G * ATTN_TILE * 4  // scores / probabilities
         + (size_t)ATTN_MAX_G * hd * 4 * ((hd < 128) ? (128 / hd) : 1) * ((hd < 128) ? 1 : 0)  // PV cross-group reduce
         + 256;
}

// RMSNorm (optional, attention.rs:202-215) + rotate-half RoPE (backends/mod.rs:444-482) of one head
// vector held as f32 in shared memory (values stay D-representable), by one warp, in place.
template <typename T, int HD>
__device__ __forceinline__ void norm_rope_inplace(float *buf, const T *norm_w, float eps, const T *cosr, const T *sinr,
                                                  int rot, int lane) {
  if (norm_w) {
    float ss = 0.f;
    for (int d = lane; d < HD; d += 32) ss += buf[d] * buf[d];
    ss = warp_sum(ss);
    const float inv = 1.0f / sqrtf(ss / (float)HD + eps);
    for (int d = lane; d < HD; d += 32) buf[d] = rnd<T>(buf[d] * inv * DT<T>::to_f(norm_w[d]));
    __syncwarp();
  }
  const int half = rot / 2;
  for (int i = lane; i < half; i += 32) {
    const float c = DT<T>::to_f(cosr[i]), s = DT<T>::to_f(sinr[i]);
    const float x1 = buf[i], x2 = buf[i + half];
    // per-op rounding in D (half-crate / __nv_bfloat16 operator semantics)
    buf[i] = rnd<T>(rnd<T>(x1 * c) - rnd<T>(x2 * s));
    buf[i + half] = rnd<T>(rnd<T>(x2 * c) + rnd<T>(x1 * s));
  }
  __syncwarp();
}
template <typename T, int HD>
__device__ __forceinline__ void norm_rope_head(const T *src, float *dst_smem, const T *norm_w, float eps,
                                               const T *cosr, const T *sinr, int rot, int lane) {
  for (int d = lane; d < HD; d += 32) dst_smem[d] = DT<T>::to_f(src[d]);
  __syncwarp();
  norm_rope_inplace<T, HD>(dst_smem, norm_w, eps, cosr, sinr, rot, lane);
}

template <typename T, int HD>
__global__ void __launch_bounds__(ATTN_THREADS) attn_decode_kernel(const AttnDecodeArgs a) {
  constexpr int LPR = HD / 8;          // lanes per cached row (16 B each)
  constexpr int RPW = 32 / LPR;        // rows per warp per iteration
  constexpr int NW = ATTN_THREADS / 32;
  constexpr int NPG = (HD < ATTN_THREADS) ? ATTN_THREADS / HD : 1;  // position groups in the PV phase
  constexpr int DPT = (HD > ATTN_THREADS) ? HD / ATTN_THREADS : 1;  // dims per thread in the PV phase
  constexpr int es = sizeof(T);
  const int G = a.n_heads / a.n_kv;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  T *Ks = reinterpret_cast<T *>(smem_raw);
  T *Vs = Ks + (size_t)ATTN_TILE * HD;
  float *q_s = reinterpret_cast<float *>(Vs + (size_t)ATTN_TILE * HD);  // [G][HD]
  float *sc = q_s + ATTN_MAX_G * HD;                                    // [G][TILE]
  float *pvred = sc + ATTN_MAX_G * ATTN_TILE;                           // [NPG][G][HD] when NPG > 1
  __shared__ float m_run[ATTN_MAX_G], l_run[ATTN_MAX_G], fac[ATTN_MAX_G];
  __shared__ float wgt[ATTN_MAX_G][ATTN_MAX_SPLIT];
  __shared__ uint64_t bar;
  __shared__ int is_last;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x, kvh = blockIdx.y;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  pdl_launch_dependents();
  pdl_wait();

  const int apos = *a.d_pos;  // absolute position (RoPE); the cache row indices below are relative to the window start
  const int ws = (a.window > 0 && apos + 1 > a.window) ? apos + 1 - a.window : 0;  // cache.rs:173-205
  const int pos = apos - ws;
  const int Tn = pos + 1;  // visible KV length after the append
  int per = (Tn + a.nsplit - 1) / a.nsplit;
  per = (per + 7) & ~7;
  const int s0 = min(Tn, split * per), s1 = min(Tn, s0 + per);
  const bool owner = (pos >= s0 && pos < s1);

  const T *qkv = reinterpret_cast<const T *>(a.qkv);
  const T *cosr = reinterpret_cast<const T *>(a.cos_t) + (size_t)apos * (a.rot / 2);
  const T *sinr = reinterpret_cast<const T *>(a.sin_t) + (size_t)apos * (a.rot / 2);
  T *kc = reinterpret_cast<T *>(a.kcache) + ((size_t)kvh * a.cap + ws) * HD;
  T *vc = reinterpret_cast<T *>(a.vcache) + ((size_t)kvh * a.cap + ws) * HD;
  __syncthreads();  // barrier init visible

  // ---- TMA: stage the first tile of cached K/V rows (everything except the row being appended) ------
  uint32_t bar_phase = 0;
  auto issue_tile = [&](int t0) {  // thread 0 only
    const int t1 = min(s1, t0 + ATTN_TILE);
    const int nold = min(t1, pos) - t0;  // rows that already live in the cache
    if (nold > 0) {
      const uint32_t bytes = (uint32_t)nold * HD * es;
      const uint64_t pol = policy_evict_last();
      mbar_arrive_expect_tx(&bar, 2 * bytes);
      bulk_g2s(Ks, kc + (size_t)t0 * HD, bytes, &bar, pol);
      bulk_g2s(Vs, vc + (size_t)t0 * HD, bytes, &bar, pol);
    } else {
      mbar_arrive(&bar);
    }
  };
  if (threadIdx.x == 0 && s0 < s1) issue_tile(s0);

  // ---- q (all CTAs) and the new k,v row (owner CTA): norm + RoPE, write the cache row in place ------
  for (int g = warp; g < G; g += NW)
    norm_rope_head<T, HD>(qkv + (size_t)(kvh * G + g) * HD, q_s + g * HD, reinterpret_cast<const T *>(a.q_norm), a.eps,
                          cosr, sinr, a.rot, lane);
  if (owner) {
    const int slot = (pos - s0) % ATTN_TILE;  // the appended row is the last row of the owner's last tile
    if (warp == NW - 1) {
      float *tmp = sc;  // scores buffer is free until the first tile is processed
      norm_rope_head<T, HD>(qkv + (size_t)(a.n_heads + kvh) * HD, tmp, reinterpret_cast<const T *>(a.k_norm), a.eps, cosr,
                            sinr, a.rot, lane);
      for (int d = lane; d < HD; d += 32) {
        const T kv = DT<T>::from_f(tmp[d]);
        kc[(size_t)pos * HD + d] = kv;
        if (pos - s0 < ATTN_TILE) Ks[(size_t)slot * HD + d] = kv;
      }
    } else if (warp == NW - 2) {
      const T *vsrc = qkv + (size_t)(a.n_heads + a.n_kv + kvh) * HD;
      for (int d = lane; d < HD; d += 32) {
        const T vv = vsrc[d];
        vc[(size_t)pos * HD + d] = vv;
        if (pos - s0 < ATTN_TILE) Vs[(size_t)slot * HD + d] = vv;
      }
    }
  }
  if (threadIdx.x < ATTN_MAX_G) { m_run[threadIdx.x] = -INFINITY; l_run[threadIdx.x] = 0.f; }
  __syncthreads();

  const int grp = lane / LPR, gl = lane % LPR;  // row group within the warp / lane within the row
  float qreg[ATTN_MAX_G][8];
#pragma unroll
  for (int g = 0; g < ATTN_MAX_G; g++)
#pragma unroll
    for (int i = 0; i < 8; i++) qreg[g][i] = (g < G) ? q_s[g * HD + gl * 8 + i] : 0.f;
  const int pv_d = threadIdx.x % HD, pv_g = threadIdx.x / HD;  // PV phase: dim / position group
  float acc[ATTN_MAX_G][DPT];
#pragma unroll
  for (int g = 0; g < ATTN_MAX_G; g++)
#pragma unroll
    for (int i = 0; i < DPT; i++) acc[g][i] = 0.f;

  for (int t0 = s0; t0 < s1; t0 += ATTN_TILE) {
    const int tn = min(ATTN_TILE, s1 - t0);
    if (t0 > s0) {  // later tiles of a long range: restage (single buffer)
      if (threadIdx.x == 0) issue_tile(t0);
      if (owner && t0 + ATTN_TILE > pos) {  // the appended row lands in this tile: take it from global (own write)
        const int slot = pos - t0;
        for (int d = threadIdx.x; d < HD; d += ATTN_THREADS) {
          Ks[(size_t)slot * HD + d] = kc[(size_t)pos * HD + d];
          Vs[(size_t)slot * HD + d] = vc[(size_t)pos * HD + d];
        }
      }
    }
    mbar_wait(&bar, bar_phase);
    bar_phase ^= 1u;
    __syncthreads();
    // ---- scores: s[g][p] = (q_g . k_p) * scale, f32 ---------------------------------------------
    for (int pb = warp * RPW; pb < tn; pb += NW * RPW) {  // warp-uniform trip count (shuffles below)
      const int p = pb + grp;
      const bool valid = p < tn;
      float kf[8];
      uint4 kraw = make_uint4(0u, 0u, 0u, 0u);
      if (valid) kraw = *reinterpret_cast<const uint4 *>(Ks + (size_t)p * HD + gl * 8);
      unpack8<T>(kraw, kf);
#pragma unroll
      for (int g = 0; g < ATTN_MAX_G; g++) {
        if (g < G) {
          float s = 0.f;
#pragma unroll
          for (int i = 0; i < 8; i++) s = fmaf(qreg[g][i], kf[i], s);
#pragma unroll
          for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          if (gl == 0 && valid) sc[g * ATTN_TILE + p] = s * a.scale;
        }
      }
    }
    __syncthreads();
    // ---- online softmax bookkeeping, one warp per head -------------------------------------------
    for (int g = warp; g < G; g += NW) {
      float mx = -INFINITY;
      for (int p = lane; p < tn; p += 32) mx = fmaxf(mx, sc[g * ATTN_TILE + p]);
      mx = warp_max(mx);
      const float m_new = fmaxf(m_run[g], mx);
      float sum = 0.f;
      for (int p = lane; p < tn; p += 32) {
        const float e = expf(sc[g * ATTN_TILE + p] - m_new);
        sc[g * ATTN_TILE + p] = e;
        sum += e;
      }
      sum = warp_sum(sum);
      if (lane == 0) {
        const float f = (m_run[g] == -INFINITY) ? 0.f : expf(m_run[g] - m_new);
        fac[g] = f;
        l_run[g] = l_run[g] * f + sum;
        m_run[g] = m_new;
      }
    }
    __syncthreads();
    // ---- PV: thread owns dim(s) pv_d of every head; position groups split the tile when HD < 128 ----
#pragma unroll
    for (int g = 0; g < ATTN_MAX_G; g++)
      if (g < G) {
        const float f = fac[g];
#pragma unroll
        for (int i = 0; i < DPT; i++) acc[g][i] *= f;
      }
#pragma unroll 4
    for (int p = pv_g; p < tn; p += NPG) {
      float vf[DPT];
#pragma unroll
      for (int i = 0; i < DPT; i++) vf[i] = DT<T>::to_f(Vs[(size_t)p * HD + pv_d + i * ATTN_THREADS]);
#pragma unroll
      for (int g = 0; g < ATTN_MAX_G; g++)
        if (g < G) {
          const float e = sc[g * ATTN_TILE + p];
#pragma unroll
          for (int i = 0; i < DPT; i++) acc[g][i] = fmaf(e, vf[i], acc[g][i]);
        }
    }
    __syncthreads();
  }

  // ---- write this split's partial (m, l, acc) --------------------------------------------------------
  if (NPG > 1) {
#pragma unroll
    for (int g = 0; g < ATTN_MAX_G; g++)
      if (g < G) pvred[((size_t)pv_g * ATTN_MAX_G + g) * HD + pv_d] = acc[g][0];
    __syncthreads();
    for (int i = threadIdx.x; i < G * HD; i += ATTN_THREADS) {
      const int g = i / HD, d = i % HD;
      float s = 0.f;
      for (int r = 0; r < NPG; r++) s += pvred[((size_t)r * ATTN_MAX_G + g) * HD + d];
      a.ws_acc[((size_t)(kvh * G + g) * a.nsplit + split) * HD + d] = s;
    }
  } else {
#pragma unroll
    for (int g = 0; g < ATTN_MAX_G; g++)
      if (g < G) {
#pragma unroll
        for (int i = 0; i < DPT; i++)
          a.ws_acc[((size_t)(kvh * G + g) * a.nsplit + split) * HD + pv_d + i * ATTN_THREADS] = acc[g][i];
      }
  }
  if (threadIdx.x < G) {
    a.ws_ml[((size_t)(kvh * G + threadIdx.x) * a.nsplit + split) * 2 + 0] = m_run[threadIdx.x];
    a.ws_ml[((size_t)(kvh * G + threadIdx.x) * a.nsplit + split) * 2 + 1] = l_run[threadIdx.x];
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned ticket = atomicAdd(&a.counters[kvh], 1u);
    is_last = (ticket == (unsigned)a.nsplit - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // ---- last CTA of this kv head: merge the splits in order, round once to D (attention.rs:346) -------
  for (int g = warp; g < G; g += NW) {  // one warp per head: lane s owns split s
    const int h = kvh * G + g;
    float m = -INFINITY, l = 0.f;
    if (lane < a.nsplit) {
      m = __ldcg(a.ws_ml + ((size_t)h * a.nsplit + lane) * 2);
      l = __ldcg(a.ws_ml + ((size_t)h * a.nsplit + lane) * 2 + 1);
    }
    const float M = warp_max(m);
    const float w = (m == -INFINITY) ? 0.f : expf(m - M);
    const float L = warp_sum(w * l);
    wgt[g][lane] = w / L;
  }
  __syncthreads();
  T *y = reinterpret_cast<T *>(a.y);
  for (int i = threadIdx.x; i < G * HD; i += ATTN_THREADS) {
    const int g = i / HD, d = i % HD, h = kvh * G + g;
    const float *src = a.ws_acc + (size_t)h * a.nsplit * HD + d;
    float o = 0.f;
#pragma unroll 8
    for (int s = 0; s < a.nsplit; s++) o = fmaf(wgt[g][s], __ldcg(src + (size_t)s * HD), o);
    y[(size_t)h * HD + d] = DT<T>::from_f(o);
  }
  if (threadIdx.x == 0) a.counters[kvh] = 0;  // re-arm for the next launch / graph replay
}

}  // namespace cake
