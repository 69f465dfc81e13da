// prefill.cuh — the seq>1 / batch>1 side of the block forward (prompt ingestion), plus the small
// bookkeeping kernels.  Same rounding points as the decode path.
//
// The GEMM here is the generic CUDA-core tile kernel used for correctness bring-up of the prefill
// path; the tcgen05 (UMMA + TMEM + TMA) GEMM replaces it for large M (see gemm_tc.cuh once present).
#pragma once
#include "common.cuh"

namespace cake {

// ---- RMSNorm over rows (backends/mod.rs:244-246): one CTA per row ---------------------------------
template <typename T>
__global__ void __launch_bounds__(256) rmsnorm_rows_kernel(const T *__restrict__ x, const T *__restrict__ w,
                                                           T *__restrict__ out, int n, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[8];
  const T *xr = x + (size_t)blockIdx.x * n;
  T *orow = out + (size_t)blockIdx.x * n;
  float ss = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float v = DT<T>::to_f(xr[i]);
    ss += v * v;
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); i++) tot += red[i];
  const float inv = 1.0f / sqrtf(tot / (float)n + eps);
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    orow[i] = DT<T>::from_f(DT<T>::to_f(xr[i]) * inv * DT<T>::to_f(w[i]));
}

// ---- the sibling block structures (olmo2 / gemma3 block.rs) -----------------------------------------------------------
// out = res + rms_norm(x) * w: the post-attention / post-feedforward norm followed by the residual add (olmo2/block.rs:77-90,
// gemma3/block.rs:120-133); two roundings to D, as the reference's rms_norm and `+` produce.  One CTA per row.
template <typename T>
__global__ void __launch_bounds__(256) rmsnorm_residual_rows_kernel(const T *__restrict__ x, const T *__restrict__ w,
                                                                    const T *__restrict__ res, T *__restrict__ out, int n, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[8];
  const T *xr = x + (size_t)blockIdx.x * n;
  const T *rr = res + (size_t)blockIdx.x * n;
  T *orow = out + (size_t)blockIdx.x * n;
  float ss = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = DT<T>::to_f(xr[i]);
    ss += v * v;
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); i++) tot += red[i];
  const float inv = 1.0f / sqrtf(tot / (float)n + eps);
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    orow[i] = DT<T>::from_f(rnd<T>(DT<T>::to_f(xr[i]) * inv * DT<T>::to_f(w[i])) + DT<T>::to_f(rr[i]));
}

// In-place RmsNorm of an n-wide segment of every row of a wider matrix (row stride `stride` elements): the OLMo2 QK-norm over
// the whole q (or k) projection inside the fused qkv buffer, attention.rs:176-192.  One CTA per row.
template <typename T>
__global__ void __launch_bounds__(256) rmsnorm_strided_kernel(T *__restrict__ x, const T *__restrict__ w, int stride, int n, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[8];
  T *xr = x + (size_t)blockIdx.x * stride;
  float ss = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = DT<T>::to_f(xr[i]);
    ss += v * v;
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); i++) tot += red[i];
  const float inv = 1.0f / sqrtf(tot / (float)n + eps);
  for (int i = threadIdx.x; i < n; i += blockDim.x) xr[i] = DT<T>::from_f(DT<T>::to_f(xr[i]) * inv * DT<T>::to_f(w[i]));
}

// The same operator at HBM speed for the row lengths the served models have (n = 256 NV): one WARP per row, the row
// read once with NV 16-byte loads per lane (all in flight together) and kept in registers between the sum of squares
// and the scaling; 8 rows per CTA.  The one-CTA-per-row kernel above (2-byte accesses, row read twice) ran at 1.85 TB/s
// on bs=32 x 4096 (ncu r02 launch list: 1.16 ms per call, 2.3 ms of a 51 ms layer).
template <typename T, int NV>
__global__ void __launch_bounds__(256) rmsnorm_rows_vec_kernel(const T *__restrict__ x, const T *__restrict__ w,
                                                               T *__restrict__ out, int rows, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int n = NV * 256;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const uint4 *xr = reinterpret_cast<const uint4 *>(x + (size_t)row * n);
  const uint4 *wr = reinterpret_cast<const uint4 *>(w);
  uint4 *orow = reinterpret_cast<uint4 *>(out + (size_t)row * n);
  uint4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; i++) v[i] = __ldcg(xr + i * 32 + lane);  // streamed once: L2 only
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NV; i++) {
    float f[8];
    unpack8<T>(v[i], f);
#pragma unroll
    for (int e = 0; e < 8; e++) ss += f[e] * f[e];
  }
  ss = warp_sum(ss);
  const float inv = 1.0f / sqrtf(ss / (float)n + eps);
#pragma unroll
  for (int i = 0; i < NV; i++) {
    float f[8], g[8];
    unpack8<T>(v[i], f);
    unpack8<T>(wr[i * 32 + lane], g);
    uint4 o;
    o.x = pack2x<T>(f[0] * inv * g[0], f[1] * inv * g[1]);
    o.y = pack2x<T>(f[2] * inv * g[2], f[3] * inv * g[3]);
    o.z = pack2x<T>(f[4] * inv * g[4], f[5] * inv * g[5]);
    o.w = pack2x<T>(f[6] * inv * g[6], f[7] * inv * g[7]);
    orow[i * 32 + lane] = o;
  }
}

// ---- C[M,N] = A[M,K] · W[N,K]^T, fp32 accumulate, ->D, then (+bias ->D) or (+residual ->D) --------
constexpr int GV0_BM = 64, GV0_BN = 64, GV0_BK = 16;
template <typename T>
__global__ void __launch_bounds__(256) gemm_v0_kernel(const T *__restrict__ A, const T *__restrict__ W,
                                                      const T *__restrict__ bias, const T *__restrict__ residual,
                                                      T *__restrict__ C, int M, int N, int K) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float As[GV0_BK][GV0_BM + 4];
  __shared__ float Ws[GV0_BK][GV0_BN + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * GV0_BM, n0 = blockIdx.x * GV0_BN;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += GV0_BK) {
    for (int i = threadIdx.x; i < GV0_BM * GV0_BK; i += 256) {
      const int r = i / GV0_BK, c = i % GV0_BK;
      As[c][r] = (m0 + r < M && k0 + c < K) ? DT<T>::to_f(A[(size_t)(m0 + r) * K + k0 + c]) : 0.f;
      Ws[c][r] = (n0 + r < N && k0 + c < K) ? DT<T>::to_f(W[(size_t)(n0 + r) * K + k0 + c]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GV0_BK; k++) {
      float av[4], wv[4];
#pragma unroll
      for (int i = 0; i < 4; i++) { av[i] = As[k][ty * 4 + i]; wv[i] = Ws[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < M && n < N) {
        float v = rnd<T>(acc[i][j]);
        if (bias) v = rnd<T>(v + DT<T>::to_f(bias[n]));
        if (residual) v = v + DT<T>::to_f(residual[(size_t)m * N + n]);
        C[(size_t)m * N + n] = DT<T>::from_f(v);
      }
    }
}

// ---- silu(gate)*up on the row-interleaved gate_up output: gu[m][2i]=gate_i, gu[m][2i+1]=up_i -------
template <typename T>
__global__ void swiglu_rows_kernel(const T *__restrict__ gu, T *__restrict__ out, size_t total /* M*I */, int act) {
  pdl_launch_dependents();
  pdl_wait();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const float g = DT<T>::to_f(gu[2 * i]), u = DT<T>::to_f(gu[2 * i + 1]);
    const float s = gate_act<T>(g, act);
    out[i] = DT<T>::from_f(s * u);
  }
}

// ---- QK-norm + RoPE + KV append for S positions (attention.rs:202-253, cache.rs:184-210) -----------
// qkv: [B, S, (n_h + 2 n_kv) * hd].  q is normalised/rotated in place; k,v go to the cache rows
// [pos0 + t].  One warp per (b, t, head) with head in [0, n_h + 2 n_kv).
template <typename T>
__global__ void __launch_bounds__(128) rope_append_kernel(T *__restrict__ qkv, T *__restrict__ kcache,
                                                          T *__restrict__ vcache, const T *__restrict__ cos_t,
                                                          const T *__restrict__ sin_t, const T *__restrict__ q_norm,
                                                          const T *__restrict__ k_norm, int B, int S, int n_heads,
                                                          int n_kv, int hd, int rot, int cap, int pos0, float eps, int rope_on) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float sm[];  // 4 warps * hd
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nh_all = n_heads + 2 * n_kv;
  const long item = (long)blockIdx.x * 4 + warp;
  if (item >= (long)B * S * nh_all) return;
  const int head = (int)(item % nh_all);
  const int t = (int)((item / nh_all) % S);
  const int b = (int)(item / ((long)nh_all * S));
  T *src = qkv + ((size_t)(b * S + t) * nh_all + head) * hd;
  float *buf = sm + warp * hd;
  const int pos = pos0 + t;
  if (head >= n_heads + n_kv) {  // v: straight copy into the cache
    const int kvh = head - n_heads - n_kv;
    T *dst = vcache + (((size_t)b * n_kv + kvh) * cap + pos) * hd;
    for (int d = lane; d < hd; d += 32) dst[d] = src[d];
    return;
  }
  const bool is_k = head >= n_heads;
  const T *nw = is_k ? k_norm : q_norm;
  for (int d = lane; d < hd; d += 32) buf[d] = DT<T>::to_f(src[d]);
  __syncwarp();
  if (nw) {
    float ss = 0.f;
    for (int d = lane; d < hd; d += 32) ss += buf[d] * buf[d];
    ss = warp_sum(ss);
    const float inv = 1.0f / sqrtf(ss / (float)hd + eps);
    for (int d = lane; d < hd; d += 32) buf[d] = rnd<T>(buf[d] * inv * DT<T>::to_f(nw[d]));
    __syncwarp();
  }
  const int half = rot / 2;
  const T *cr = cos_t + (size_t)pos * half, *sr = sin_t + (size_t)pos * half;
  for (int i = lane; i < half && rope_on; i += 32) {  // rope_on == 0: this layer does not rotate (attention.rs:242-253)
    const float c = DT<T>::to_f(cr[i]), s = DT<T>::to_f(sr[i]);
    const float x1 = buf[i], x2 = buf[i + half];
    buf[i] = rnd<T>(rnd<T>(x1 * c) - rnd<T>(x2 * s));
    buf[i + half] = rnd<T>(rnd<T>(x2 * c) + rnd<T>(x1 * s));
  }
  __syncwarp();
  T *dst = is_k ? kcache + (((size_t)b * n_kv + (head - n_heads)) * cap + pos) * hd : src;
  for (int d = lane; d < hd; d += 32) dst[d] = DT<T>::from_f(buf[d]);
}

// The same operator for head_dim 128 with full rotary (every served family but the partial-rotary ones): one warp per
// TOKEN walks its n_h + 2 n_kv heads four at a time — lane l holds elements {2l, 2l+1} and {64+2l, 65+2l} of a head, i.e.
// both partners of its two rotation pairs, so nothing goes through shared memory; every access is a coalesced 128-byte
// warp transaction and eight of them are in flight per lane; cos/sin are read once per token instead of once per head.
// Identical arithmetic and rounding points (QK-norm -> D, products -> D, sum -> D).  The one-warp-per-head kernel above ran
// at 1.4 TB/s on bs=32 x 4096 (ncu r02 launch list: 2.35 ms per layer).
template <typename T>
__global__ void __launch_bounds__(128) rope_append_vec_kernel(T *__restrict__ qkv, T *__restrict__ kcache, T *__restrict__ vcache,
                                                              const T *__restrict__ cos_t, const T *__restrict__ sin_t,
                                                              const T *__restrict__ q_norm, const T *__restrict__ k_norm, int B,
                                                              int S, int n_heads, int n_kv, int cap, int pos0, float eps, int rope_on) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int HD = 128, HC = 4;
  const int lane = threadIdx.x & 31;
  const long tok = (long)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (tok >= (long)B * S) return;
  const int t = (int)(tok % S), b = (int)(tok / S);
  const int nh_all = n_heads + 2 * n_kv, pos = pos0 + t;
  uint32_t *row = reinterpret_cast<uint32_t *>(qkv + (size_t)tok * nh_all * HD);  // 64 words per head
  float2 c = DT<T>::unpack2(reinterpret_cast<const uint32_t *>(cos_t + (size_t)pos * (HD / 2))[lane]);
  float2 sn = DT<T>::unpack2(reinterpret_cast<const uint32_t *>(sin_t + (size_t)pos * (HD / 2))[lane]);
  if (!rope_on) { c = make_float2(1.f, 1.f); sn = make_float2(0.f, 0.f); }  // no rotation on this layer: x*1 - y*0 = x exactly
  float2 qw0 = {1.f, 1.f}, qw1 = {1.f, 1.f}, kw0 = {1.f, 1.f}, kw1 = {1.f, 1.f};
  if (q_norm) { qw0 = DT<T>::unpack2(reinterpret_cast<const uint32_t *>(q_norm)[lane]); qw1 = DT<T>::unpack2(reinterpret_cast<const uint32_t *>(q_norm)[32 + lane]); }
  if (k_norm) { kw0 = DT<T>::unpack2(reinterpret_cast<const uint32_t *>(k_norm)[lane]); kw1 = DT<T>::unpack2(reinterpret_cast<const uint32_t *>(k_norm)[32 + lane]); }
  for (int h0 = 0; h0 < nh_all; h0 += HC) {
    uint32_t a0[HC], a1[HC];
#pragma unroll
    for (int i = 0; i < HC; i++)
      if (h0 + i < nh_all) {
        a0[i] = row[(h0 + i) * 64 + lane];
        a1[i] = row[(h0 + i) * 64 + 32 + lane];
      }
#pragma unroll
    for (int i = 0; i < HC; i++) {
      const int head = h0 + i;
      if (head >= nh_all) break;
      if (head >= n_heads + n_kv) {  // v: straight copy into the cache
        uint32_t *dst = reinterpret_cast<uint32_t *>(vcache + (((size_t)b * n_kv + (head - n_heads - n_kv)) * cap + pos) * HD);
        dst[lane] = a0[i];
        dst[32 + lane] = a1[i];
        continue;
      }
      const bool is_k = head >= n_heads;
      float2 x0 = DT<T>::unpack2(a0[i]), x1 = DT<T>::unpack2(a1[i]);  // x0 = elements 2l, 2l+1; x1 = 64+2l, 65+2l
      if (is_k ? (k_norm != nullptr) : (q_norm != nullptr)) {
        const float ss = warp_sum(x0.x * x0.x + x0.y * x0.y + x1.x * x1.x + x1.y * x1.y);
        const float inv = 1.0f / sqrtf(ss / (float)HD + eps);
        const float2 w0 = is_k ? kw0 : qw0, w1 = is_k ? kw1 : qw1;
        x0.x = rnd<T>(x0.x * inv * w0.x); x0.y = rnd<T>(x0.y * inv * w0.y);
        x1.x = rnd<T>(x1.x * inv * w1.x); x1.y = rnd<T>(x1.y * inv * w1.y);
      }
      const float r0x = rnd<T>(x0.x * c.x) - rnd<T>(x1.x * sn.x), r0y = rnd<T>(x0.y * c.y) - rnd<T>(x1.y * sn.y);
      const float r1x = rnd<T>(x1.x * c.x) + rnd<T>(x0.x * sn.x), r1y = rnd<T>(x1.y * c.y) + rnd<T>(x0.y * sn.y);
      uint32_t *dst = is_k ? reinterpret_cast<uint32_t *>(kcache + (((size_t)b * n_kv + (head - n_heads)) * cap + pos) * HD)
                           : row + head * 64;
      dst[lane] = pack2x<T>(r0x, r0y);
      dst[32 + lane] = pack2x<T>(r1x, r1y);
    }
  }
}

// ---- causal attention for S query positions against the cache (attention.rs:300-346) ---------------
// One warp per (b, head, t).  f32 throughout, online softmax, result ->D.  Query t (absolute position
// pos0+t) sees cache rows [0, pos0+t]  (mask j-(T-S) > i, attention.rs:314-341).
template <typename T>
__global__ void __launch_bounds__(128) attn_prefill_v0_kernel(const T *__restrict__ qkv, const T *__restrict__ kcache,
                                                              const T *__restrict__ vcache, T *__restrict__ y, int B,
                                                              int S, int n_heads, int n_kv, int hd, int cap, int pos0,
                                                              float scale) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long item = (long)blockIdx.x * 4 + warp;
  if (item >= (long)B * S * n_heads) return;
  const int h = (int)(item % n_heads);
  const int t = (int)((item / n_heads) % S);
  const int b = (int)(item / ((long)n_heads * S));
  const int nh_all = n_heads + 2 * n_kv, G = n_heads / n_kv;
  const T *q = qkv + ((size_t)(b * S + t) * nh_all + h) * hd;
  const T *kc = kcache + ((size_t)b * n_kv + h / G) * cap * hd;
  const T *vc = vcache + ((size_t)b * n_kv + h / G) * cap * hd;
  constexpr int MAXD = 8;  // hd <= 256
  float qf[MAXD], acc[MAXD];
  const int nd = (hd + 31) / 32;
#pragma unroll
  for (int i = 0; i < MAXD; i++) {
    const int d = lane + 32 * i;
    qf[i] = (i < nd && d < hd) ? DT<T>::to_f(q[d]) : 0.f;
    acc[i] = 0.f;
  }
  float m = -INFINITY, l = 0.f;
  const int kv_len = pos0 + t + 1;
  for (int p = 0; p < kv_len; p++) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXD; i++) {
      const int d = lane + 32 * i;
      if (i < nd && d < hd) s = fmaf(qf[i], DT<T>::to_f(kc[(size_t)p * hd + d]), s);
    }
    s = warp_sum(s) * scale;
    const float mn = fmaxf(m, s);
    const float f = (m == -INFINITY) ? 0.f : expf(m - mn);
    const float e = expf(s - mn);
    l = l * f + e;
#pragma unroll
    for (int i = 0; i < MAXD; i++) {
      const int d = lane + 32 * i;
      if (i < nd && d < hd) acc[i] = acc[i] * f + e * DT<T>::to_f(vc[(size_t)p * hd + d]);
    }
    m = mn;
  }
  const float inv = 1.0f / l;
  T *yo = y + ((size_t)(b * S + t) * n_heads + h) * hd;
#pragma unroll
  for (int i = 0; i < MAXD; i++) {
    const int d = lane + 32 * i;
    if (i < nd && d < hd) yo[d] = DT<T>::from_f(acc[i] * inv);
  }
}

// ---- embedding gather (backends/mod.rs:513-528) ---------------------------------------------------
template <typename T>
__global__ void embed_kernel(const T *__restrict__ E, const uint32_t *__restrict__ ids, T *__restrict__ x, int n_tok,
                             int H, int vocab, float scale /* text_model.rs:274-276 embed_scale; 0 = none */) {
  pdl_launch_dependents();
  pdl_wait();
  const int t = blockIdx.x;
  if (t >= n_tok) return;
  uint32_t id = ids[t];
  if (id >= (uint32_t)vocab) id = 0;
  const uint4 *src = reinterpret_cast<const uint4 *>(E + (size_t)id * H);
  uint4 *dst = reinterpret_cast<uint4 *>(x + (size_t)t * H);
  if (scale == 0.f) {
    for (int i = threadIdx.x; i < H * (int)sizeof(T) / 16; i += blockDim.x) dst[i] = src[i];
  } else {  // `x * scale` in D: the scalar is rounded to D first (candle affine on a half tensor), one rounding of the product
    const float sc = rnd<T>(scale);
    for (int i = threadIdx.x; i < H * (int)sizeof(T) / 16; i += blockDim.x) {
      float f[8];
      unpack8<T>(src[i], f);
      uint4 o;
      o.x = pack2<T>(f[0] * sc, f[1] * sc); o.y = pack2<T>(f[2] * sc, f[3] * sc);
      o.z = pack2<T>(f[4] * sc, f[5] * sc); o.w = pack2<T>(f[6] * sc, f[7] * sc);
      dst[i] = o;
    }
  }
}

// ---- decode-loop bookkeeping ----------------------------------------------------------------------
__global__ void set_int_kernel(int *p, int v) { *p = v; }
__global__ void set_u32_kernel(uint32_t *p, uint32_t v) { *p = v; }
__global__ void advance_kernel(int *pos, int *step) {
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x == 0) { *pos += 1; *step += 1; }
}

// ---- argmax over logits already in memory (first maximum wins), single CTA ------------------------
template <typename T>
__global__ void __launch_bounds__(1024) argmax_kernel(const T *__restrict__ logits, int V, uint32_t *out) {
  __shared__ float bv[32];
  __shared__ int bi[32];
  float best = -INFINITY;
  int idx = 0x7fffffff;
  for (int i = threadIdx.x; i < V; i += blockDim.x) {
    const float v = DT<T>::to_f(logits[i]);
    if (v > best) { best = v; idx = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > best || (ov == best && oi < idx)) { best = ov; idx = oi; }
  }
  if ((threadIdx.x & 31) == 0) { bv[threadIdx.x >> 5] = best; bi[threadIdx.x >> 5] = idx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); w++)
      if (bv[w] > best || (bv[w] == best && bi[w] < idx)) { best = bv[w]; idx = bi[w]; }
    *out = (uint32_t)idx;
  }
}

// ---- repeat penalty (text_model.rs:60-99) on device, arithmetic in D; tokens are pre-deduplicated --
template <typename T>
__global__ void repeat_penalty_kernel(T *logits, int V, float penalty, const uint32_t *toks, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t t = toks[i];
  if (t >= (uint32_t)V) return;
  const float recip = rnd<T>(1.0f / penalty), pen = rnd<T>(penalty);
  const float sel = DT<T>::to_f(logits[t]);
  const float mult = sel >= 0.f ? recip : pen;
  const float penalized = rnd<T>(sel * mult);
  const float delta = rnd<T>(penalized - sel);
  logits[t] = DT<T>::from_f(sel + delta);
}

// ---- synthetic cache fill (bench only) ------------------------------------------------------------
template <typename T>
__global__ void fill_synth_kernel(T *p, size_t n, uint32_t seed, float amp) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u ^ seed;
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    p[i] = DT<T>::from_f(((float)(h & 0xffff) / 32768.0f - 1.0f) * amp);
  }
}

}  // namespace cake
