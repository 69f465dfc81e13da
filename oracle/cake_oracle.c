/*
 * cake_oracle.c — CPU restatement of cake's sharded transformer-block forward path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under cake_b200/ (the product) may include, link,
 * import or execute this file.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and there only as the checker / CPU arm.
 *
 * What it restates (all citations relative to /root/reference, commit be522af):
 *   text_model.rs:266-368   TextModelBase::forward  (embed -> blocks -> ln_f -> last -> lm_head)
 *   transformer.rs:103-135  Transformer::forward    (pre-norm block, two residual adds)
 *   attention.rs:152-357    CausalSelfAttention::forward, generic f32 branch :300-346
 *                           (the branch `cake run --cpu` and non-flash CUDA builds execute)
 *   cache.rs:31-114,184-210 Cache::new (RoPE tables, llama3 scaling), process_kv_inner
 *   mlp.rs:21-31            MLP::forward (fused gate_up, silu*mul, down)
 *   backends/mod.rs:206-246 linear_forward / rms_norm ; :407-433 softmax ; :444-482 rope ;
 *                           :513-528 embedding ; cpu/mod.rs:87-89 silu_mul
 *   text_model.rs:60-99     apply_repeat_penalty_gpu ; :102-118 ArgMax when temperature<=0
 *
 * The arithmetic underneath those call sites lives in candle 0.9.2 (crates.io, Cargo.lock:817-933),
 * which is NOT vendored under /root/reference and cannot be built here (no Rust toolchain).
 * PARITY PINNING: op level only.  The reference's own tests hold known answers for
 * rms_norm, causal mask, cos(pos 0), silu, cache growth (tests/unit_tests/test_backend_ops.rs:87-98,
 * :204-237; test_cache.rs:25-96; test_backends.rs:29-38) — tests/test_oracle_kat.py checks this file
 * against every one of them.  For whole-block / logits outputs the reference holds NO expected
 * tensors (test_blocks.rs:877-933 assert shapes only): block-level parity is UNPINNED in the
 * reference; we cross-validate this file against HuggingFace transformers (independent fp32
 * implementation of the same architectures) via committed fixtures in tests/golden/.
 *
 * dtype boundaries mirrored (D = model dtype): every tensor the reference materialises in D is
 * rounded to D here (values are carried in float but are always D-representable):
 *   rms_norm: f32 sum of squares, x*rsqrt*w in f32, one rounding           
 *   linear:   D x D products (exact in f32), f32 accumulate, one rounding; bias add rounds again
 *   rope:     per-op rounding in D (half-crate / __nv_bfloat16 operator semantics)
 *   attention core: f32 throughout, one rounding of the result (attention.rs:301-346)
 *   residual adds in D; silu -> D then mul -> D (cpu/mod.rs:87-89); logits D -> f32 -> argmax
 */
#include <immintrin.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { ORA_BF16 = 0, ORA_F16 = 1, ORA_F32 = 2 };

typedef struct ora_config {
  int hidden, inter, n_heads, n_kv_heads, head_dim, n_layers, vocab, max_seq;
  float rms_eps, rope_theta, partial_rotary;
  int qkv_bias, qk_norm, tie_embeddings;
  int rope_llama3;
  float rope_factor, rope_low, rope_high;
  int rope_orig_max;
  int dtype;     /* ORA_* */
  int sliding_window; /* cache.rs:173-205 (0 = none) */
  int use_gelu_mlp;   /* mlp.rs:25-26 */
  float embed_scale;  /* text_model.rs:274-276 (0 = none) */
  int pre_reshape_qk_norm; /* attention.rs:176-192 (OLMo2): QK-norm over the whole q / k projection, before the head reshape */
  int silu_mode; /* 0: silu in f32 then round (cpu/mod.rs:87-89 via candle_nn::ops::silu)
                    1: per-op D arithmetic x/(1+exp(-x))*y (cuda ops.cu:105-109) — tolerance probe */
} ora_config;

typedef struct ora_layer {
  const void *q, *k, *v, *o, *gate, *up, *down, *ln1, *ln2;
  const void *q_bias, *k_bias, *v_bias; /* nullable */
  const void *q_norm, *k_norm;          /* nullable */
  /* the sibling block structures (models/{olmo2,gemma3,exaone4}/block.rs): ln1 / ln2 may be NULL (OLMo2 has no input
   * norms); post_attn / post_ffn, when set, are RmsNorms applied to the attention / MLP output BEFORE the residual add */
  const void *post_attn, *post_ffn;     /* nullable */
  int window;   /* this layer's sliding window (attention.rs load_custom): -1 = the config's, 0 = none (global layer) */
  int no_rope;  /* 1: this layer does not rotate q / k (attention.rs:242-253; exaone4 global, gemma3 local layers) */
} ora_layer;

typedef struct ora_model {
  ora_config cfg;
  ora_layer *layers;
  const void *embed, *ln_f, *lm_head;
  float *cos_t, *sin_t; /* (max_seq, rot/2), rounded to D */
  int rot;
} ora_model;

typedef struct ora_cache {
  int n_layers, cap, n_kv, hd;
  int *len;
  float **k, **v; /* per layer: (n_kv, cap, hd), D-representable values */
} ora_cache;

/* ---------------------------------------------------------------- rounding */
static inline float bf16_bits_to_f32(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static inline uint16_t f32_to_bf16_bits(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40); /* quiet NaN */
  uint32_t r = 0x7fffu + ((u >> 16) & 1u); /* round to nearest even */
  return (uint16_t)((u + r) >> 16);
}
static inline float rnd(float f, int dt) {
  if (dt == ORA_BF16) return bf16_bits_to_f32(f32_to_bf16_bits(f));
  if (dt == ORA_F16) return (float)(_Float16)f;
  return f;
}
static inline float wld(const void *w, size_t i, int dt) {
  if (dt == ORA_BF16) return bf16_bits_to_f32(((const uint16_t *)w)[i]);
  if (dt == ORA_F16) return (float)((const _Float16 *)w)[i];
  return ((const float *)w)[i];
}
float ora_round(float f, int dt) { return rnd(f, dt); }
static size_t esize(int dt) { return dt == ORA_F32 ? 4 : 2; }

/* ---------------------------------------------------------------- dot / linear */
/* 16 independent partial sums so gcc vectorises without reassociating a single chain.
 * Products of two D values are exact in f32 for bf16 (8x8 bit) and f16 (11x11 bit). */
#if defined(__AVX2__) && defined(__FMA__)
/* bf16 row . f32 vector with 2x8 fp32 partial sums.  The products are exact in fp32 (x holds D values), so
 * fused multiply-add equals multiply-then-add here; only the summation order differs from the scalar loop. */
static float dot_row_bf16_avx2(const uint16_t *w, const float *x, int K) {
  __m256 a0 = _mm256_setzero_ps(), a1 = _mm256_setzero_ps();
  int k = 0;
  for (; k + 16 <= K; k += 16) {
    const __m256i raw = _mm256_loadu_si256((const __m256i *)(w + k));
    const __m256i lo = _mm256_slli_epi32(_mm256_cvtepu16_epi32(_mm256_castsi256_si128(raw)), 16);
    const __m256i hi = _mm256_slli_epi32(_mm256_cvtepu16_epi32(_mm256_extracti128_si256(raw, 1)), 16);
    a0 = _mm256_fmadd_ps(_mm256_castsi256_ps(lo), _mm256_loadu_ps(x + k), a0);
    a1 = _mm256_fmadd_ps(_mm256_castsi256_ps(hi), _mm256_loadu_ps(x + k + 8), a1);
  }
  float t[8];
  _mm256_storeu_ps(t, _mm256_add_ps(a0, a1));
  float s = ((t[0] + t[1]) + (t[2] + t[3])) + ((t[4] + t[5]) + (t[6] + t[7]));
  for (; k < K; k++) s += bf16_bits_to_f32(w[k]) * x[k];
  return s;
}
#endif

/* Summation-order probe (tolerance measurement, not a behaviour switch of the restatement): the reference fixes
 * "f32 accumulate" but not the ORDER — candle's CPU gemm, cuBLAS and any GPU kernel each add the K exact products in
 * a different order, and after 32 layers of bf16 roundings that order is visible in the logits.
 *   0  default: 2x8 SIMD partial sums (bf16, AVX2) / 16 scalar partial sums
 *   1  16 scalar partial sums for every dtype (a second, equally valid f32 order)
 *   2  f64 accumulation, one rounding to f32 (the order-free limit every f32 order scatters around)
 * tests compare logits across these modes to state how far two CORRECT implementations may differ. */
static int g_dot_mode = 0;
void ora_set_dot_mode(int m) { g_dot_mode = m; }
int ora_dot_mode(void) { return g_dot_mode; }

static float dot_row(const void *W, size_t off, const float *x, int K, int dt) {
  if (g_dot_mode == 2) {
    double s = 0.0;
    for (int k = 0; k < K; k++) s += (double)wld(W, off + (size_t)k, dt) * (double)x[k];
    return (float)s;
  }
#if defined(__AVX2__) && defined(__FMA__)
  if (dt == ORA_BF16 && g_dot_mode == 0) return dot_row_bf16_avx2((const uint16_t *)W + off, x, K);
#endif
  float acc[16];
  for (int j = 0; j < 16; j++) acc[j] = 0.f;
  int k = 0;
  if (dt == ORA_BF16) {
    const uint16_t *w = (const uint16_t *)W + off;
    for (; k + 16 <= K; k += 16)
      for (int j = 0; j < 16; j++) acc[j] += bf16_bits_to_f32(w[k + j]) * x[k + j];
    for (; k < K; k++) acc[k & 15] += bf16_bits_to_f32(w[k]) * x[k];
  } else if (dt == ORA_F16) {
    const _Float16 *w = (const _Float16 *)W + off;
    for (; k + 16 <= K; k += 16)
      for (int j = 0; j < 16; j++) acc[j] += (float)w[k + j] * x[k + j];
    for (; k < K; k++) acc[k & 15] += (float)w[k] * x[k];
  } else {
    const float *w = (const float *)W + off;
    for (; k + 16 <= K; k += 16)
      for (int j = 0; j < 16; j++) acc[j] += w[k + j] * x[k + j];
    for (; k < K; k++) acc[k & 15] += w[k] * x[k];
  }
  float s = 0.f;
  for (int j = 0; j < 16; j++) s += acc[j];
  return s;
}

/* backends/mod.rs:206-241 linear_forward: out = x @ W^T (+ bias).  W is HF [N,K] row-major in D.
 * x: (S,K) float holding D values; out: (S,N), leading dimension ldo. */
void ora_linear(const float *x, int S, int K, const void *W, int N, const void *bias, float *out,
                int ldo, int dt) {
#pragma omp parallel for schedule(static)
  for (int n = 0; n < N; n++) {
    float b = bias ? wld(bias, (size_t)n, dt) : 0.f;
    for (int s = 0; s < S; s++) {
      float v = rnd(dot_row(W, (size_t)n * K, x + (size_t)s * K, K, dt), dt);
      if (bias) v = rnd(v + b, dt); /* broadcast_add in D */
      out[(size_t)s * ldo + n] = v;
    }
  }
}

/* backends/mod.rs:244-246 -> candle_nn::ops::rms_norm: x*w/sqrt(mean(x^2)+eps); f32 sum. */
void ora_rms_norm(const float *x, int rows, int n, const void *w, float eps, float *out, int dt) {
  for (int r = 0; r < rows; r++) {
    const float *xr = x + (size_t)r * n;
    float ss = 0.f;
    for (int i = 0; i < n; i++) ss += xr[i] * xr[i];
    float inv = 1.0f / sqrtf(ss / (float)n + eps);
    for (int i = 0; i < n; i++) out[(size_t)r * n + i] = rnd(xr[i] * inv * wld(w, (size_t)i, dt), dt);
  }
}

/* cache.rs:43-96 — theta_i = 1/base^(i/rot) for i = 0,2,4..; llama3 scaling :49-80;
 * ang = pos*theta (f32); cos/sin -> D.  Tables are (max_seq, rot/2). */
void ora_rope_tables(const ora_config *c, float *cos_t, float *sin_t) {
  int rot = (int)((float)c->head_dim * c->partial_rotary);
  int half = rot / 2;
  float *theta = (float *)malloc(sizeof(float) * (size_t)(half > 0 ? half : 1));
  for (int j = 0; j < half; j++) theta[j] = 1.0f / powf(c->rope_theta, (float)(2 * j) / (float)rot);
  if (c->rope_llama3 && c->rope_orig_max > 0) {
    float old = (float)c->rope_orig_max;
    float low_wl = old / c->rope_low, high_wl = old / c->rope_high;
    for (int j = 0; j < half; j++) {
      float wl = 2.0f * 3.14159265358979323846f / theta[j];
      if (wl < high_wl) {
      } else if (wl > low_wl) {
        theta[j] /= c->rope_factor;
      } else {
        float smooth = (old / wl - c->rope_low) / (c->rope_high - c->rope_low);
        theta[j] = (1.0f - smooth) * (theta[j] / c->rope_factor) + smooth * theta[j];
      }
    }
  }
  for (int p = 0; p < c->max_seq; p++)
    for (int j = 0; j < half; j++) {
      float ang = (float)p * theta[j];
      cos_t[(size_t)p * half + j] = rnd(cosf(ang), c->dtype);
      sin_t[(size_t)p * half + j] = rnd(sinf(ang), c->dtype);
    }
  free(theta);
}

/* backends/mod.rs:444-482 (+ attention.rs:59-67 partial rotary): non-interleaved rotate-half on
 * one head vector at absolute position pos.  Per-op rounding in D. */
static void rope_vec(float *x, int hd, int rot, const float *cos_t, const float *sin_t, int pos, int dt) {
  int half = rot / 2;
  const float *c = cos_t + (size_t)pos * half, *s = sin_t + (size_t)pos * half;
  for (int i = 0; i < half; i++) {
    float x1 = x[i], x2 = x[i + half];
    x[i] = rnd(rnd(x1 * c[i], dt) - rnd(x2 * s[i], dt), dt);
    x[i + half] = rnd(rnd(x2 * c[i], dt) + rnd(x1 * s[i], dt), dt);
  }
  (void)hd;
}
void ora_rope(float *x, int n_vec, int hd, int rot, const float *cos_t, const float *sin_t, int pos, int dt) {
  for (int v = 0; v < n_vec; v++) rope_vec(x + (size_t)v * hd, hd, rot, cos_t, sin_t, pos, dt);
}

/* cache.rs:150-160 — mask[i][j] = j > i (u8). */
void ora_causal_mask(int seq, uint8_t *out) {
  for (int i = 0; i < seq; i++)
    for (int j = 0; j < seq; j++) out[i * seq + j] = (uint8_t)(j > i);
}

/* cpu/mod.rs:87-89: silu(gate) * up.  mode 0: silu in f32, round, mul, round.
 * mode 1 (ops.cu:105-109): every op in D. */
float ora_silu_mul(float g, float u, int dt, int mode) {
  if (mode == 0 || dt == ORA_F32) {
    float s = rnd(g / (1.0f + expf(-g)), dt);
    return rnd(s * u, dt);
  }
  float e = rnd(expf(-g), dt);
  float d = rnd(1.0f + e, dt);
  float q = rnd(g / d, dt);
  return rnd(q * u, dt);
}
/* mlp.rs:25-26 use_gelu_mlp: `backend.gelu(gate) * up`; candle's `gelu` is the tanh approximation
 * 0.5 x (1 + tanh(sqrt(2/pi) x (1 + 0.044715 x^2))).  Evaluated in f32 and rounded to D, then the product rounded to D
 * (assumption: candle's half-precision gelu may round per op; unverifiable here, same status as the rms_norm note). */
float ora_gelu_mul(float g, float u, int dt) {
  const float a = rnd(0.5f * g * (1.0f + tanhf(0.7978845608028654f * g * (1.0f + 0.044715f * g * g))), dt);
  return rnd(a * u, dt);
}

/* ---------------------------------------------------------------- model / cache */
ora_model *ora_model_create(const ora_config *c) {
  ora_model *m = (ora_model *)calloc(1, sizeof(ora_model));
  m->cfg = *c;
  m->layers = (ora_layer *)calloc((size_t)c->n_layers, sizeof(ora_layer));
  m->rot = (int)((float)c->head_dim * c->partial_rotary);
  size_t n = (size_t)c->max_seq * (size_t)(m->rot / 2);
  m->cos_t = (float *)malloc(sizeof(float) * (n ? n : 1));
  m->sin_t = (float *)malloc(sizeof(float) * (n ? n : 1));
  ora_rope_tables(c, m->cos_t, m->sin_t);
  return m;
}
void ora_model_set_layer(ora_model *m, int i, const ora_layer *l) { m->layers[i] = *l; }
void ora_model_set_head(ora_model *m, const void *embed, const void *ln_f, const void *lm_head) {
  m->embed = embed;
  m->ln_f = ln_f;
  m->lm_head = lm_head ? lm_head : embed; /* text_model.rs:164-167 tie_word_embeddings */
}
void ora_model_free(ora_model *m) {
  if (!m) return;
  free(m->layers);
  free(m->cos_t);
  free(m->sin_t);
  free(m);
}
const float *ora_model_cos(const ora_model *m) { return m->cos_t; }
const float *ora_model_sin(const ora_model *m) { return m->sin_t; }

ora_cache *ora_cache_create(const ora_config *c, int cap) {
  ora_cache *k = (ora_cache *)calloc(1, sizeof(ora_cache));
  k->n_layers = c->n_layers;
  k->cap = cap;
  k->n_kv = c->n_kv_heads;
  k->hd = c->head_dim;
  k->len = (int *)calloc((size_t)c->n_layers, sizeof(int));
  k->k = (float **)calloc((size_t)c->n_layers, sizeof(float *));
  k->v = (float **)calloc((size_t)c->n_layers, sizeof(float *));
  return k;
}
/* cache.rs:247-253 clear() */
void ora_cache_clear(ora_cache *k) {
  for (int i = 0; i < k->n_layers; i++) k->len[i] = 0;
}
void ora_cache_free(ora_cache *k) {
  if (!k) return;
  for (int i = 0; i < k->n_layers; i++) {
    free(k->k[i]);
    free(k->v[i]);
  }
  free(k->k);
  free(k->v);
  free(k->len);
  free(k);
}
int ora_cache_len(const ora_cache *k, int layer) { return k->len[layer]; }
/* direct fill (used to synthesise a pre-filled cache for the CPU timing leg / parity at depth) */
float *ora_cache_k(ora_cache *k, int layer) {
  size_t n = (size_t)k->n_kv * k->cap * k->hd;
  if (!k->k[layer]) {
    k->k[layer] = (float *)calloc(n, sizeof(float));
    k->v[layer] = (float *)calloc(n, sizeof(float));
  }
  return k->k[layer];
}
float *ora_cache_v(ora_cache *k, int layer) {
  ora_cache_k(k, layer);
  return k->v[layer];
}
void ora_cache_set_len(ora_cache *k, int layer, int len) { k->len[layer] = len; }

/* ---------------------------------------------------------------- one block */
/* transformer.rs:103-135 with attention.rs:152-357 (generic f32 branch) and mlp.rs:21-31.
 * x,out: (S,H).  Returns 0, or <0 on error (cache overflow). */
int ora_block_forward(const ora_model *m, int layer, ora_cache *kc, const float *x, int S, int index_pos,
                      float *out) {
  const ora_config *c = &m->cfg;
  const ora_layer *L = &m->layers[layer];
  const int H = c->hidden, I = c->inter, nh = c->n_heads, nkv = c->n_kv_heads, hd = c->head_dim;
  const int dt = c->dtype, g = nh / nkv, sq = nh * hd, skv = nkv * hd;
  const int P0 = kc->len[layer], T = P0 + S;
  if (T > kc->cap || index_pos + S > c->max_seq) return -1;
  float *kk = ora_cache_k(kc, layer), *vv = ora_cache_v(kc, layer);

  float *h1 = (float *)malloc(sizeof(float) * (size_t)S * H);
  float *q = (float *)malloc(sizeof(float) * (size_t)S * sq);
  float *kn = (float *)malloc(sizeof(float) * (size_t)S * skv);
  float *vn = (float *)malloc(sizeof(float) * (size_t)S * skv);
  float *y = (float *)malloc(sizeof(float) * (size_t)S * sq);
  float *x1 = (float *)malloc(sizeof(float) * (size_t)S * H);
  float *gu = (float *)malloc(sizeof(float) * (size_t)S * 2 * I);
  float *mm = (float *)malloc(sizeof(float) * (size_t)S * I);

  /* transformer.rs:112 rms_1 (olmo2/block.rs:70-76: no input norm, the attention reads x itself) */
  if (L->ln1) ora_rms_norm(x, S, H, L->ln1, c->rms_eps, h1, dt);
  else memcpy(h1, x, sizeof(float) * (size_t)S * H);
  /* attention.rs:162-174 fused qkv linear == three row-stacked linears (cat dim 0, :109-113) */
  ora_linear(h1, S, H, L->q, sq, L->q_bias, q, sq, dt);
  ora_linear(h1, S, H, L->k, skv, L->k_bias, kn, skv, dt);
  ora_linear(h1, S, H, L->v, skv, L->v_bias, vn, skv, dt);
  /* attention.rs:176-192 OLMo2: QK-norm over the full projection (norm dim = size_q / size_kv), before the reshape */
  if (c->qk_norm && c->pre_reshape_qk_norm && L->q_norm && L->k_norm) {
    ora_rms_norm(q, S, sq, L->q_norm, c->rms_eps, q, dt);
    ora_rms_norm(kn, S, skv, L->k_norm, c->rms_eps, kn, dt);
  }
  /* attention.rs:202-215 per-head QK-norm over head_dim (Qwen3) */
  if (c->qk_norm && !c->pre_reshape_qk_norm && L->q_norm && L->k_norm) {
    ora_rms_norm(q, S * nh, hd, L->q_norm, c->rms_eps, q, dt);
    ora_rms_norm(kn, S * nkv, hd, L->k_norm, c->rms_eps, kn, dt);
  }
  /* attention.rs:242-253 RoPE at absolute positions index_pos + t; v is not rotated */
  for (int t = 0; t < S && !L->no_rope; t++) {
    ora_rope(q + (size_t)t * sq, nh, hd, m->rot, m->cos_t, m->sin_t, index_pos + t, dt);
    ora_rope(kn + (size_t)t * skv, nkv, hd, m->rot, m->cos_t, m->sin_t, index_pos + t, dt);
  }
  /* cache.rs:184-210 append along seq: layout (n_kv, cap, hd) */
  for (int t = 0; t < S; t++)
    for (int h = 0; h < nkv; h++) {
      memcpy(kk + ((size_t)h * kc->cap + P0 + t) * hd, kn + (size_t)t * skv + (size_t)h * hd, sizeof(float) * hd);
      memcpy(vv + ((size_t)h * kc->cap + P0 + t) * hd, vn + (size_t)t * skv + (size_t)h * hd, sizeof(float) * hd);
    }
  kc->len[layer] = T;

  /* attention.rs:300-346: f32; repeat_kv maps q-head j -> kv-head j/g (:359-364);
   * att = q k^T * (1/sqrt(hd)) (candle `Tensor / f64` is affine(1/rhs, 0) evaluated in f32);
   * mask j-(T-S) > i when S>1 (:314-341); softmax max-subtracted, *1/sum (backends/mod.rs:407-433);
   * att @ v; to_dtype(in_dtype) (:346). */
  const float scale = (float)(1.0 / sqrt((double)hd));
  /* cache.rs:173-205 process_kv_windowed: when a cache entry already exists, cat(cache, new) is cut to its last
   * limit = min(window, max_seq_len) rows; the FIRST call stores (and attends over) everything (test_cache.rs:99-124).
   * The rows stay in place here; ws is the first row that survives the cut. */
  int ws = 0;
  const int window = L->window >= 0 ? L->window : c->sliding_window;  /* per-layer override: load_custom(…, sliding_window, …) */
  if (window > 0 && P0 > 0) {
    const int limit = window < c->max_seq ? window : c->max_seq;
    if (T > limit) ws = T - limit;
  }
#pragma omp parallel for collapse(2) schedule(static)
  for (int t = 0; t < S; t++)
    for (int j = 0; j < nh; j++) {
      float *att = (float *)malloc(sizeof(float) * (size_t)T);
      const float *qv = q + (size_t)t * sq + (size_t)j * hd;
      const float *kh = kk + (size_t)(j / g) * kc->cap * hd;
      const float *vh = vv + (size_t)(j / g) * kc->cap * hd;
      float mx = -INFINITY;
      for (int p = 0; p < ws; p++) att[p] = 0.f;
      for (int p = ws; p < T; p++) {
        float s = 0.f;
        for (int d = 0; d < hd; d++) s += qv[d] * kh[(size_t)p * hd + d];
        s *= scale;
        if (S > 1 && (p - (T - S)) > t) s = -INFINITY;
        att[p] = s;
        if (s > mx) mx = s;
      }
      float sum = 0.f;
      for (int p = ws; p < T; p++) {
        float e = expf(att[p] - mx);
        att[p] = e;
        sum += e;
      }
      float inv = 1.0f / sum;
      float *yo = y + (size_t)t * sq + (size_t)j * hd;
      for (int d = 0; d < hd; d++) yo[d] = 0.f;
      for (int p = ws; p < T; p++) {
        float a = att[p] * inv;
        for (int d = 0; d < hd; d++) yo[d] += a * vh[(size_t)p * hd + d];
      }
      for (int d = 0; d < hd; d++) yo[d] = rnd(yo[d], dt);
      free(att);
    }
  /* attention.rs:354 o_proj; transformer.rs:123 residual (D add) */
  ora_linear(y, S, sq, L->o, H, NULL, x1, H, dt);
  /* olmo2/block.rs:77-79, gemma3/block.rs:120-122: post-attention RmsNorm of the attention output, then the residual */
  if (L->post_attn) ora_rms_norm(x1, S, H, L->post_attn, c->rms_eps, x1, dt);
  for (size_t i = 0; i < (size_t)S * H; i++) x1[i] = rnd(x1[i] + x[i], dt);
  /* transformer.rs:129 rms_2; mlp.rs:22-30 */
  if (L->ln2) ora_rms_norm(x1, S, H, L->ln2, c->rms_eps, h1, dt);
  else memcpy(h1, x1, sizeof(float) * (size_t)S * H);
  ora_linear(h1, S, H, L->gate, I, NULL, gu, 2 * I, dt);
  ora_linear(h1, S, H, L->up, I, NULL, gu + I, 2 * I, dt);
  for (int t = 0; t < S; t++)
    for (int i = 0; i < I; i++)
      mm[(size_t)t * I + i] = c->use_gelu_mlp ? ora_gelu_mul(gu[(size_t)t * 2 * I + i], gu[(size_t)t * 2 * I + I + i], dt)
                                              : ora_silu_mul(gu[(size_t)t * 2 * I + i], gu[(size_t)t * 2 * I + I + i], dt, c->silu_mode);
  ora_linear(mm, S, I, L->down, H, NULL, out, H, dt);
  if (L->post_ffn) ora_rms_norm(out, S, H, L->post_ffn, c->rms_eps, out, dt);  /* olmo2/block.rs:84-86, gemma3/block.rs:130-132 */
  /* transformer.rs:131 mlp residual */
  for (size_t i = 0; i < (size_t)S * H; i++) out[i] = rnd(out[i] + x1[i], dt);

  free(h1); free(q); free(kn); free(vn); free(y); free(x1); free(gu); free(mm);
  return 0;
}

/* ---------------------------------------------------------------- head / tail */
/* backends/mod.rs:513-528 embedding == index_select rows (text_model.rs:271) */
void ora_embed(const ora_model *m, const uint32_t *ids, int S, float *x) {
  int H = m->cfg.hidden;
  for (int t = 0; t < S; t++)
    for (int i = 0; i < H; i++) x[(size_t)t * H + i] = wld(m->embed, (size_t)ids[t] * H + i, m->cfg.dtype);
  /* text_model.rs:274-276 `x * scale as f64`: candle's scalar multiply is affine(mul, 0) in the tensor's dtype — the
   * scalar is converted to D, the product rounded once (assumption about candle internals, stated in the header) */
  if (m->cfg.embed_scale != 0.f) {
    const float sc = rnd(m->cfg.embed_scale, m->cfg.dtype);
    for (size_t i = 0; i < (size_t)S * H; i++) x[i] = rnd(x[i] * sc, m->cfg.dtype);
  }
}
/* text_model.rs:336-352: ln_f, last position, lm_head -> logits (V) in D */
void ora_logits(const ora_model *m, const float *x, int S, float *logits) {
  int H = m->cfg.hidden, dt = m->cfg.dtype;
  float *hf = (float *)malloc(sizeof(float) * (size_t)H);
  ora_rms_norm(x + (size_t)(S - 1) * H, 1, H, m->ln_f, m->cfg.rms_eps, hf, dt);
  ora_linear(hf, 1, H, m->lm_head, m->cfg.vocab, NULL, logits, m->cfg.vocab, dt);
  free(hf);
}
/* text_model.rs:104-105 Sampling::ArgMax over f32(logits); first maximum wins. */
uint32_t ora_argmax(const float *logits, int V) {
  int best = 0;
  for (int i = 1; i < V; i++)
    if (logits[i] > logits[best]) best = i;
  return (uint32_t)best;
}
/* text_model.rs:60-99 apply_repeat_penalty_gpu, arithmetic in D. ctx tokens are de-duplicated. */
void ora_repeat_penalty(float *logits, int V, float penalty, const uint32_t *ctx, int n, int dt) {
  float recip = rnd(1.0f / penalty, dt), pen = rnd(penalty, dt);
  for (int i = 0; i < n; i++) {
    int dup = 0;
    for (int j = 0; j < i; j++)
      if (ctx[j] == ctx[i]) dup = 1;
    if (dup || ctx[i] >= (uint32_t)V) continue;
    float sel = logits[ctx[i]];
    float mult = sel >= 0.f ? recip : pen;
    float penalized = rnd(sel * mult, dt);
    float delta = rnd(penalized - sel, dt);
    logits[ctx[i]] = rnd(sel + delta, dt);
  }
}
/* text_model.rs:266-368 whole forward: ids (S) at index_pos -> logits (V).
 * layers [l0,l1) only (a shard) when x_io != NULL: then the embed/head are skipped and x_io is
 * transformed in place (Worker semantics, worker.rs:442-479). */
int ora_forward(const ora_model *m, ora_cache *kc, const uint32_t *ids, int S, int index_pos, float *logits) {
  int H = m->cfg.hidden;
  float *a = (float *)malloc(sizeof(float) * (size_t)S * H), *b = (float *)malloc(sizeof(float) * (size_t)S * H);
  ora_embed(m, ids, S, a);
  for (int l = 0; l < m->cfg.n_layers; l++) {
    int rc = ora_block_forward(m, l, kc, a, S, index_pos, b);
    if (rc) { free(a); free(b); return rc; }
    float *t = a; a = b; b = t;
  }
  ora_logits(m, a, S, logits);
  free(a); free(b);
  return 0;
}
int ora_forward_layers(const ora_model *m, ora_cache *kc, int l0, int l1, float *x_io, int S, int index_pos) {
  int H = m->cfg.hidden;
  float *b = (float *)malloc(sizeof(float) * (size_t)S * H);
  for (int l = l0; l < l1; l++) {
    int rc = ora_block_forward(m, l, kc, x_io, S, index_pos, b);
    if (rc) { free(b); return rc; }
    memcpy(x_io, b, sizeof(float) * (size_t)S * H);
  }
  free(b);
  return 0;
}
void ora_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
int ora_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
size_t ora_esize(int dt) { return esize(dt); }
