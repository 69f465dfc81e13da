// sample.cuh — device-side token samplers: the tail of TextModelBase::next_token
// (text_model.rs:102-118 create_logits_processor, :429-460 the sampling step) without shipping the logits to the
// host: one CTA reads the (vocab,) logits of the last position from L2/HBM and writes 4 bytes.
//
// The arithmetic behind `LogitsProcessor::sample` lives in candle-transformers 0.9.2 (generation/mod.rs) and
// candle-nn 0.9.2 (sampling.rs), crates that are not vendored under /root/reference; restated from their published
// algorithm (oracle/sampling.py holds the CPU restatement the tests compare with):
//   all kinds        logits -> f32
//   GumbelSoftmax    argmax(logits + T * g),  g = -log(-log(u)),  u ~ U(1e-7, 0.999)      (candle_nn::sampling)
//   All / TopK / TopP / TopKThenTopP:
//                    prs = softmax(logits * (1/T))  (max-subtracted, f32), then
//     All            multinomial(prs)                         (rand WeightedIndex: first i with cumsum_i > x*total)
//     TopK{k}        the k most probable tokens, multinomial over them
//     TopP{p}        sort descending, keep tokens while the running sum before them is < p, multinomial over the
//                    kept probabilities in VOCABULARY order (the others zeroed)       (p <= 0 or >= 1: All)
//     TopKThenTopP   top-k, then the nucleus rule inside those k
// What cannot be pinned: (1) the random stream — candle draws from rand::StdRng / the device RNG, here a counter-based
// Philox4x32-10 keyed by (seed, step); tests supply the uniforms instead (`noise`); (2) candle's TopK builds its
// candidate list with select_nth_unstable, whose ORDER is unspecified, so which token a given uniform selects is
// implementation-defined there; here the candidates are ordered by probability (descending, ties by lower id).
// Everything is deterministic: reductions and scans use fixed trees, never floating-point atomics.
#pragma once
#include "common.cuh"

namespace cake {

enum SampleKind { SAMPLE_ARGMAX = 0, SAMPLE_ALL = 1, SAMPLE_TOPK = 2, SAMPLE_TOPP = 3, SAMPLE_TOPK_TOPP = 4, SAMPLE_GUMBEL = 5 };
constexpr int SAMPLE_THREADS = 1024;
constexpr int SAMPLE_MAX_K = 1024;   // candidates sorted in shared memory (top-k requests above this are refused by the host)

struct SampleArgs {
  int kind, top_k;
  float temperature, top_p;
  unsigned long long seed;
};

// ---- Philox4x32-10 (counter-based; Salmon et al. 2011) ------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const unsigned long long p0 = (unsigned long long)0xD2511F53u * ctr.x, p1 = (unsigned long long)0xCD9E8D57u * ctr.z;
    ctr = make_uint4((uint32_t)(p1 >> 32) ^ ctr.y ^ key.x, (uint32_t)p1, (uint32_t)(p0 >> 32) ^ ctr.w ^ key.y, (uint32_t)p0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}
__device__ __forceinline__ float u01(uint32_t r) { return (float)(r >> 8) * (1.0f / 16777216.0f); }  // [0, 1)
// uniform for vocabulary entry v (Gumbel) / the single draw of a step (v = 0xffffffff)
__device__ __forceinline__ float sample_uniform(const SampleArgs &s, unsigned step, unsigned v) {
  const uint4 r = philox4x32(make_uint4(v >> 2, step, 0x5a4d504cu, 0u), make_uint2((uint32_t)s.seed, (uint32_t)(s.seed >> 32)));
  const uint32_t w = (v & 3) == 0 ? r.x : (v & 3) == 1 ? r.y : (v & 3) == 2 ? r.z : r.w;
  return u01(w);
}

// ---- block-wide helpers over SAMPLE_THREADS threads, fixed reduction order --------------------------------------
__device__ __forceinline__ float block_sum(float v, float *sh) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = (threadIdx.x < 32) ? sh[threadIdx.x] : 0.f;
  if (threadIdx.x < 32) t = warp_sum(t);
  if (threadIdx.x == 0) sh[32] = t;
  __syncthreads();
  const float r = sh[32];
  __syncthreads();
  return r;
}
__device__ __forceinline__ float block_max(float v, float *sh) {
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = (threadIdx.x < 32) ? sh[threadIdx.x] : -INFINITY;
  if (threadIdx.x < 32) t = warp_max(t);
  if (threadIdx.x == 0) sh[32] = t;
  __syncthreads();
  const float r = sh[32];
  __syncthreads();
  return r;
}
__device__ __forceinline__ int block_sum_i(int v, int *sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  int t = (threadIdx.x < 32) ? sh[threadIdx.x] : 0;
  if (threadIdx.x < 32) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  if (threadIdx.x == 0) sh[32] = t;
  __syncthreads();
  const int r = sh[32];
  __syncthreads();
  return r;
}
__device__ __forceinline__ int block_min_i(int v, int *sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  int t = (threadIdx.x < 32) ? sh[threadIdx.x] : 0x7fffffff;
  if (threadIdx.x < 32) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t = min(t, __shfl_xor_sync(0xffffffffu, t, o));
  }
  if (threadIdx.x == 0) sh[32] = t;
  __syncthreads();
  const int r = sh[32];
  __syncthreads();
  return r;
}

// Multinomial draw over w[0..n) in index order (w >= 0): the first i whose inclusive prefix sum exceeds u * total
// (rand::distributions::WeightedIndex).  Thread t owns the contiguous chunk [t*C, (t+1)*C); chunk sums are combined by a
// fixed tree, so the result is deterministic.  `keep(i)` masks entries out (TopP).
template <typename Keep>
__device__ __forceinline__ int block_multinomial(const float *w, int n, float u, Keep keep, float *shf, int *shi) {
  const int C = (n + SAMPLE_THREADS - 1) / SAMPLE_THREADS;
  const int i0 = min(n, (int)threadIdx.x * C), i1 = min(n, i0 + C);
  float loc = 0.f;
  for (int i = i0; i < i1; i++) loc += keep(i) ? w[i] : 0.f;
  // exclusive scan of the 1024 chunk sums: warp scan, then a scan of the 32 warp totals
  float inc = loc;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(0xffffffffu, inc, o);
    if ((int)(threadIdx.x & 31) >= o) inc += t;
  }
  if ((threadIdx.x & 31) == 31) shf[threadIdx.x >> 5] = inc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float wt = shf[threadIdx.x], winc = wt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_up_sync(0xffffffffu, winc, o);
      if ((int)threadIdx.x >= o) winc += t;
    }
    shf[32 + threadIdx.x] = winc - wt;  // exclusive prefix of warp totals
    if (threadIdx.x == 31) shf[64] = winc;  // grand total
  }
  __syncthreads();
  const float total = shf[64];
  float run = shf[32 + (threadIdx.x >> 5)] + (inc - loc);
  const float target = u * total;
  int found = 0x7fffffff;
  for (int i = i0; i < i1; i++) {
    if (!keep(i)) continue;
    run += w[i];
    if (run > target) { found = i; break; }
  }
  __syncthreads();
  int r = block_min_i(found, shi);
  if (r == 0x7fffffff) {  // u*total rounded up to the total itself: the last kept entry (WeightedIndex can return len-1)
    int last = -1;
    for (int i = i1 - 1; i >= i0; i--) if (keep(i) && w[i] > 0.f) { last = i; break; }
    r = -block_min_i(last < 0 ? 0x7fffffff : -last, shi);
    if (r == -0x7fffffff) r = 0;
  }
  return r;
}

// `noise`: nullable; tests pass the uniforms explicitly — vocab floats in [0,1) for Gumbel, one float otherwise.
template <typename T>
__global__ void __launch_bounds__(SAMPLE_THREADS) sample_kernel(const T *__restrict__ logits, int V, SampleArgs s, float *__restrict__ P,
                                                                 const float *__restrict__ noise, const int *d_step, int step_bias,
                                                                 uint32_t *token_out, uint32_t *ring, int ring_cap) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float shf[80];
  __shared__ int shi[40];
  __shared__ float c_p[SAMPLE_MAX_K];
  __shared__ int c_i[SAMPLE_MAX_K];
  __shared__ int n_cand, chosen;
  const int tid = threadIdx.x;
  const unsigned step = (unsigned)(d_step ? *d_step + step_bias : 0);
  auto finish = [&](int tok) {
    if (tid == 0) {
      *token_out = (uint32_t)tok;
      if (ring) ring[step % (unsigned)ring_cap] = (uint32_t)tok;
    }
  };

  if (s.kind == SAMPLE_ARGMAX || s.temperature <= 0.f) {  // Sampling::ArgMax: first maximum wins
    float best = -INFINITY;
    int idx = 0x7fffffff;
    for (int i = tid; i < V; i += SAMPLE_THREADS) {
      const float v = DT<T>::to_f(logits[i]);
      if (v > best) { best = v; idx = i; }
    }
    const float bm = block_max(best, shf);
    finish(block_min_i(best == bm ? idx : 0x7fffffff, shi));
    return;
  }
  if (s.kind == SAMPLE_GUMBEL) {
    // candle_nn::sampling::gumbel_softmax: minus_g = log(-log(u)), u in [1e-7, 0.999); T == 1: logits - minus_g, else
    // logits + minus_g * (-T); argmax
    float best = -INFINITY;
    int idx = 0x7fffffff;
    for (int i = tid; i < V; i += SAMPLE_THREADS) {
      const float r = noise ? noise[i] : sample_uniform(s, step, (unsigned)i);
      const float u = 1e-7f + r * (0.999f - 1e-7f);
      const float minus_g = logf(-logf(u));
      const float l = DT<T>::to_f(logits[i]);
      const float v = (s.temperature == 1.0f) ? l - minus_g : l + minus_g * (-s.temperature);
      if (v > best || (v == best && i < idx)) { best = v; idx = i; }
    }
    const float bm = block_max(best, shf);
    finish(block_min_i(best == bm ? idx : 0x7fffffff, shi));
    return;
  }

  // ---- prs = softmax(logits * (1/T)) in f32 ------------------------------------------------------------------
  const float inv_t = (float)(1.0 / (double)s.temperature);
  float mx = -INFINITY;
  for (int i = tid; i < V; i += SAMPLE_THREADS) mx = fmaxf(mx, DT<T>::to_f(logits[i]) * inv_t);
  mx = block_max(mx, shf);
  float sum = 0.f;
  for (int i = tid; i < V; i += SAMPLE_THREADS) {
    const float e = expf(DT<T>::to_f(logits[i]) * inv_t - mx);
    P[i] = e;
    sum += e;
  }
  sum = block_sum(sum, shf);
  for (int i = tid; i < V; i += SAMPLE_THREADS) P[i] = P[i] / sum;
  __syncthreads();
  const float u = noise ? noise[0] : sample_uniform(s, step, 0xffffffffu);

  const bool want_k = (s.kind == SAMPLE_TOPK || s.kind == SAMPLE_TOPK_TOPP) && s.top_k < V;
  const bool want_p = (s.kind == SAMPLE_TOPP || s.kind == SAMPLE_TOPK_TOPP) && s.top_p > 0.f && s.top_p < 1.f;
  if (!want_k && !want_p) {  // All (also TopK with k >= vocab, TopP with p outside (0,1))
    finish(block_multinomial(P, V, u, [](int) { return true; }, shf, shi));
    return;
  }

  // ---- the K most probable entries: bisection on the (monotone) bit pattern of the k-th largest probability -----
  // pure TopP first looks at the SAMPLE_MAX_K most probable tokens: the nucleus almost always lies within them
  const int K = want_k ? s.top_k : SAMPLE_MAX_K;
  uint32_t lo = 0u, hi = 0x7f800000u;  // count(P >= lo) >= K always holds for lo = 0
  while (hi - lo > 1u) {
    const uint32_t mid = lo + (hi - lo) / 2u;
    int cnt = 0;
    for (int i = tid; i < V; i += SAMPLE_THREADS) cnt += (__float_as_uint(P[i]) >= mid);
    cnt = block_sum_i(cnt, shi);
    if (cnt >= K) lo = mid; else hi = mid;
  }
  // lo = bit pattern of the K-th largest probability: take everything above it, then ties in index order up to K
  {
    int cnt = 0;
    for (int i = tid; i < V; i += SAMPLE_THREADS) cnt += (__float_as_uint(P[i]) > lo);
    const int n_gt = block_sum_i(cnt, shi);
    if (tid == 0) n_cand = 0;
    __syncthreads();
    // deterministic gather: thread t owns a contiguous chunk; exclusive offsets by two integer scans (gt, then ties)
    const int C = (V + SAMPLE_THREADS - 1) / SAMPLE_THREADS;
    const int i0 = min(V, tid * C), i1 = min(V, i0 + C);
    int my_gt = 0, my_eq = 0;
    for (int i = i0; i < i1; i++) {
      const uint32_t b = __float_as_uint(P[i]);
      my_gt += (b > lo);
      my_eq += (b == lo);
    }
    auto excl_scan = [&](int v) -> int {  // exclusive prefix over threads
      int inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if ((tid & 31) >= o) inc += t;
      }
      if ((tid & 31) == 31) shi[tid >> 5] = inc;
      __syncthreads();
      if (tid < 32) {
        int wt = shi[tid], winc = wt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, winc, o);
          if (tid >= o) winc += t;
        }
        shi[tid] = winc - wt;
      }
      __syncthreads();
      const int r = shi[tid >> 5] + inc - v;
      __syncthreads();
      return r;
    };
    int off_gt = excl_scan(my_gt), off_eq = excl_scan(my_eq);
    for (int i = i0; i < i1; i++) {
      const uint32_t b = __float_as_uint(P[i]);
      if (b > lo) { c_p[off_gt] = P[i]; c_i[off_gt] = i; off_gt++; }
      else if (b == lo) {
        const int slot = n_gt + off_eq;
        if (slot < K) { c_p[slot] = P[i]; c_i[slot] = i; }
        off_eq++;
      }
    }
    __syncthreads();
  }
  // ---- sort the K candidates: probability descending, ties by lower id (bitonic, shared memory) ---------------
  int n = 1;
  while (n < K) n <<= 1;
  for (int i = K + tid; i < n; i += SAMPLE_THREADS) { c_p[i] = -1.f; c_i[i] = 0x7fffffff; }
  __syncthreads();
  for (int k2 = 2; k2 <= n; k2 <<= 1)
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < n; i += SAMPLE_THREADS) {
        const int l = i ^ j;
        if (l > i) {
          const bool up = ((i & k2) == 0);
          const bool a_first = (c_p[i] > c_p[l]) || (c_p[i] == c_p[l] && c_i[i] < c_i[l]);  // i should precede l
          if (a_first != up) {
            const float tp = c_p[i]; c_p[i] = c_p[l]; c_p[l] = tp;
            const int ti = c_i[i]; c_i[i] = c_i[l]; c_i[l] = ti;
          }
        }
      }
      __syncthreads();
    }
  // ---- the rest is a short sequential walk over at most 1024 sorted candidates ---------------------------------
  if (s.kind == SAMPLE_TOPP) {
    // nucleus inside the top-SAMPLE_MAX_K: keep while the running sum BEFORE the token is < p
    if (tid == 0) {
      float cum = 0.f;
      int m = 0;
      for (; m < K; m++) {
        if (cum >= s.top_p) break;
        cum += c_p[m];
      }
      n_cand = (m == K && cum < s.top_p) ? -1 : m;   // -1: the nucleus is larger than the candidate list
      chosen = m > 0 ? __float_as_int(c_p[m - 1]) : 0;
    }
    __syncthreads();
    if (n_cand >= 0) {
      // kept set = sorted candidates [0, n_cand): in vocabulary order that is {P > tau} plus the first ties at tau
      const uint32_t tau = (uint32_t)chosen;
      int ties_kept = 0;
      if (tid == 0) {
        for (int m = 0; m < n_cand; m++) ties_kept += (__float_as_uint(c_p[m]) == tau);
        shi[39] = ties_kept;
      }
      __syncthreads();
      ties_kept = shi[39];
      __syncthreads();
      // rank of a tie = number of equal entries with a lower id: ties are few; count them on the fly
      finish(block_multinomial(P, V, u, [&](int i) {
        const uint32_t b = __float_as_uint(P[i]);
        if (b > tau) return true;
        if (b != tau) return false;
        int rank = 0;
        for (int m = 0; m < K; m++) rank += (__float_as_uint(c_p[m]) == tau && c_i[m] < i);
        return rank < ties_kept;
      }, shf, shi));
      return;
    }
    // flat distribution: the nucleus exceeds 1024 tokens.  Bisection on the bit pattern of the probability at which the
    // descending running sum crosses p (mass reductions by fixed trees), then the same vocabulary-order draw.
    uint32_t a = 0u, b = 0x7f800000u;  // mass(P >= a) >= p holds for a = 0
    while (b - a > 1u) {
      const uint32_t mid = a + (b - a) / 2u;
      float m = 0.f;
      for (int i = tid; i < V; i += SAMPLE_THREADS) m += (__float_as_uint(P[i]) >= mid) ? P[i] : 0.f;
      m = block_sum(m, shf);
      if (m >= s.top_p) a = mid; else b = mid;
    }
    float m_gt = 0.f;
    for (int i = tid; i < V; i += SAMPLE_THREADS) m_gt += (__float_as_uint(P[i]) > a) ? P[i] : 0.f;
    m_gt = block_sum(m_gt, shf);
    const float tauf = __uint_as_float(a);
    int need = 0;  // ties at tau kept: until the running sum before a tie reaches p
    { float cum = m_gt; while (cum < s.top_p && need < V) { cum += tauf; need++; } }
    // ties in index order: keep the first `need` (prefix count of ties by chunks)
    const int C = (V + SAMPLE_THREADS - 1) / SAMPLE_THREADS;
    const int i0 = min(V, tid * C), i1 = min(V, i0 + C);
    int my_eq = 0;
    for (int i = i0; i < i1; i++) my_eq += (__float_as_uint(P[i]) == a);
    int inc = my_eq;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if ((tid & 31) >= o) inc += t;
    }
    if ((tid & 31) == 31) shi[tid >> 5] = inc;
    __syncthreads();
    if (tid < 32) {
      int wt = shi[tid], winc = wt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, winc, o);
        if (tid >= o) winc += t;
      }
      shi[tid] = winc - wt;
    }
    __syncthreads();
    int eq_before = shi[tid >> 5] + inc - my_eq;
    __syncthreads();
    // zero the dropped entries in place (P is scratch), then draw
    for (int i = i0; i < i1; i++) {
      const uint32_t bb = __float_as_uint(P[i]);
      if (bb < a) P[i] = 0.f;
      else if (bb == a) { if (eq_before >= need) P[i] = 0.f; eq_before++; }
    }
    __syncthreads();
    finish(block_multinomial(P, V, u, [](int) { return true; }, shf, shi));
    return;
  }
  // TopK / TopKThenTopP: draw over the sorted candidates (our defined order: probability descending)
  if (tid == 0) {
    int m = K;
    if (want_p) {  // nucleus rule inside the k candidates (sample_topk_topp: sort, clip, multinomial)
      float cum = 0.f;
      for (m = 0; m < K; m++) {
        if (cum >= s.top_p) break;
        cum += c_p[m];
      }
    }
    float total = 0.f;
    for (int j = 0; j < m; j++) total += c_p[j];
    const float target = u * total;
    float run = 0.f;
    int pick = m - 1;
    for (int j = 0; j < m; j++) {
      run += c_p[j];
      if (run > target) { pick = j; break; }
    }
    chosen = c_i[pick < 0 ? 0 : pick];
  }
  __syncthreads();
  finish(chosen);
}

}  // namespace cake
