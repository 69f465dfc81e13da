"""Builds a stand-in for libcake_b200.so (CPU, TEST-ONLY — it lives under tests/ and the product never loads it):

* ``build(dir)``: recording mode.  Every entry point of include/cake_b200.h exists; `ctx_create` / `head_load` /
  `block_load` log what they were given (FNV-1a of the bytes behind every weight pointer, sizes derived from the
  config), everything else fails with a status.  Checks the compiled host's loading path without a GPU.
* ``build(dir, oracle=True)``: emulation mode.  The same ABI computed by the CPU oracle (oracle/cake_oracle.c): blocks,
  per-session caches with the library's `index_pos == cache length` rule, embed / logits / repeat-penalty, the host-
  stepped decode loop.  "Device" pointers are plain host memory.  With it the compiled host side (cake_run's
  TextModelBase / Master loop, cake_worker's sessions and forward_ops) runs end to end on the CPU and must reproduce
  the oracle's tokens / activations — the host logic is what is under test, the arithmetic is the checker's.

    LD_LIBRARY_PATH=<dir of the fake> FAKE_B200_LOG=log.txt cake_b200/host/cake_run <model_dir> --prompt-ids 1
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

HAND = r'''
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "cake_b200.h"

struct cake_b200_ctx { cake_b200_config cfg; };
struct cake_b200_block { int layer; };
struct cake_b200_cache { int dummy; };
static cake_b200_ctx g_ctx;
static FILE *logf(void) { const char *p = getenv("FAKE_B200_LOG"); return p ? fopen(p, "a") : NULL; }
static unsigned long long fnv(const void *p, size_t n) {
  if (!p) return 0ull;
  unsigned long long h = 1469598103934665603ull;
  const unsigned char *b = (const unsigned char *)p;
  for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
  return h;
}
const char *cake_b200_last_error(void) { return "fake libcake_b200: no device"; }
const char *cake_b200_version(void) { return "cake_b200 fake"; }
int cake_b200_ctx_create(int device, const cake_b200_config *cfg, cake_b200_ctx **out) {
  g_ctx.cfg = *cfg;
  FILE *f = logf();
  if (f) {
    fprintf(f, "ctx %d %d %d %d %d %d %d %d %d %g %g %g %d %d %d %d\n", device, cfg->hidden, cfg->inter, cfg->n_heads, cfg->n_kv_heads,
            cfg->head_dim, cfg->n_layers, cfg->vocab, cfg->max_seq, (double)cfg->rms_eps, (double)cfg->rope_theta,
            (double)cfg->partial_rotary, cfg->qkv_bias, cfg->qk_norm, cfg->tie_embeddings, cfg->dtype);
    fclose(f);
  }
  *out = &g_ctx;
  return 0;
}
void cake_b200_ctx_destroy(cake_b200_ctx *c) { (void)c; }
int cake_b200_head_load(cake_b200_ctx *c, const void *embed, const void *ln_f, const void *lm_head) {
  const size_t es = 2, H = c->cfg.hidden, V = c->cfg.vocab;
  FILE *f = logf();
  if (f) { fprintf(f, "head %016llx %016llx %016llx\n", fnv(embed, V * H * es), fnv(ln_f, H * es), fnv(lm_head, V * H * es)); fclose(f); }
  return 0;
}
int cake_b200_block_load(cake_b200_ctx *c, int layer, const void *q, const void *k, const void *v, const void *o, const void *gate,
                         const void *up, const void *down, const void *ln1, const void *ln2, const void *qb, const void *kb,
                         const void *vb, const void *qn, const void *kn, cake_b200_block **out) {
  const size_t es = 2, H = c->cfg.hidden, I = c->cfg.inter, hd = c->cfg.head_dim, sq = (size_t)c->cfg.n_heads * hd, skv = (size_t)c->cfg.n_kv_heads * hd;
  FILE *f = logf();
  if (f) {
    fprintf(f, "block %d %016llx %016llx %016llx %016llx %016llx %016llx %016llx %016llx %016llx %016llx %016llx %016llx %016llx %016llx\n", layer,
            fnv(q, sq * H * es), fnv(k, skv * H * es), fnv(v, skv * H * es), fnv(o, H * sq * es), fnv(gate, I * H * es), fnv(up, I * H * es),
            fnv(down, H * I * es), fnv(ln1, H * es), fnv(ln2, H * es), fnv(qb, sq * es), fnv(kb, skv * es), fnv(vb, skv * es),
            fnv(qn, hd * es), fnv(kn, hd * es));
    fclose(f);
  }
  cake_b200_block *b = (cake_b200_block *)malloc(sizeof *b);
  b->layer = layer;
  *out = b;
  return 0;
}
void cake_b200_block_free(cake_b200_block *b) { free(b); }
int cake_b200_cache_create(cake_b200_ctx *c, int batch, int max_seq, cake_b200_cache **out) {
  (void)c; (void)batch; (void)max_seq;
  static cake_b200_cache k;
  *out = &k;
  return 0;
}
void cake_b200_cache_free(cake_b200_cache *k) { (void)k; }
'''
ORACLE = r'''
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "cake_b200.h"

/* oracle/cake_oracle.c API (test infrastructure) */
typedef struct ora_config { int hidden, inter, n_heads, n_kv_heads, head_dim, n_layers, vocab, max_seq; float rms_eps, rope_theta, partial_rotary;
  int qkv_bias, qk_norm, tie_embeddings; int rope_llama3; float rope_factor, rope_low, rope_high; int rope_orig_max; int dtype; int sliding_window, use_gelu_mlp; float embed_scale; int pre_reshape_qk_norm; int silu_mode; } ora_config;
typedef struct ora_layer { const void *q, *k, *v, *o, *gate, *up, *down, *ln1, *ln2, *q_bias, *k_bias, *v_bias, *q_norm, *k_norm, *post_attn, *post_ffn; int window, no_rope; } ora_layer;
typedef struct ora_model ora_model;
typedef struct ora_cache ora_cache;
ora_model *ora_model_create(const ora_config *);
void ora_model_set_layer(ora_model *, int, const ora_layer *);
void ora_model_set_head(ora_model *, const void *, const void *, const void *);
void ora_model_free(ora_model *);
ora_cache *ora_cache_create(const ora_config *, int cap);
void ora_cache_clear(ora_cache *);
void ora_cache_free(ora_cache *);
int ora_cache_len(const ora_cache *, int);
int ora_block_forward(const ora_model *, int layer, ora_cache *, const float *x, int S, int index_pos, float *out);
void ora_embed(const ora_model *, const uint32_t *ids, int S, float *x);
void ora_logits(const ora_model *, const float *x, int S, float *logits);
uint32_t ora_argmax(const float *logits, int V);
void ora_repeat_penalty(float *logits, int V, float penalty, const uint32_t *ctx, int n, int dt);
float ora_round(float f, int dt);

struct cake_b200_ctx { cake_b200_config cfg; ora_config oc; ora_model *m; int n_blocks_dec; int dec_idx[512]; cake_b200_cache *dec_cache; int pos; float *last_logits;
  uint32_t cur_token; uint32_t ring[4096]; unsigned long long steps; };
struct cake_b200_block { int layer; cake_b200_ctx *c; ora_layer l; };
struct cake_b200_cache { cake_b200_ctx *c; ora_cache *k; ora_cache **rows; int batch; int cap; };  /* k == rows[0] */
static __thread char g_err[256] = "";
static int fail(int code, const char *msg) { snprintf(g_err, sizeof g_err, "%s", msg); return code; }
const char *cake_b200_last_error(void) { return g_err; }
const char *cake_b200_version(void) { return "cake_b200 fake (oracle emulation)"; }

/* D bits <-> f32 (values carried by the oracle are exactly representable in D) */
static float h2f(uint16_t h) { uint32_t s = (h >> 15) & 1, e = (h >> 10) & 31, m = h & 1023, u;
  if (e == 0) { if (!m) u = s << 31; else { e = 113; while (!(m & 1024)) { m <<= 1; e--; } u = (s << 31) | (e << 23) | ((m & 1023) << 13); } }
  else if (e == 31) u = (s << 31) | 0x7f800000u | (m << 13); else u = (s << 31) | ((e + 112) << 23) | (m << 13);
  float f; memcpy(&f, &u, 4); return f; }
static uint16_t f2h(float f) { uint32_t u; memcpy(&u, &f, 4); uint32_t s = (u >> 16) & 0x8000, e = (u >> 23) & 255, m = u & 0x7fffff;
  if (e >= 143) return (uint16_t)(s | 0x7c00 | ((e == 255 && m) ? 0x200 : 0)); if (e <= 102) return (uint16_t)s;
  if (e <= 112) { m |= 0x800000; return (uint16_t)(s | (m >> (126 - e))); } return (uint16_t)(s | ((e - 112) << 10) | (m >> 13)); }
static void to_f32(const cake_b200_ctx *c, const void *src, float *dst, size_t n) { const uint16_t *p = (const uint16_t *)src;
  for (size_t i = 0; i < n; i++) { if (c->cfg.dtype == CAKE_B200_BF16) { uint32_t u = (uint32_t)p[i] << 16; memcpy(&dst[i], &u, 4); } else dst[i] = h2f(p[i]); } }
static void from_f32(const cake_b200_ctx *c, const float *src, void *dst, size_t n) { uint16_t *p = (uint16_t *)dst;
  for (size_t i = 0; i < n; i++) { float r = ora_round(src[i], c->cfg.dtype); if (c->cfg.dtype == CAKE_B200_BF16) { uint32_t u; memcpy(&u, &r, 4); p[i] = (uint16_t)(u >> 16); } else p[i] = f2h(r); } }

static const void *own_copy(const void *p, size_t n) { void *q = malloc(n); memcpy(q, p, n); return q; }  /* leaked: test processes are short-lived */

int cake_b200_ctx_create(int device, const cake_b200_config *cfg, cake_b200_ctx **out) {
  if (!cfg || !out) return fail(CAKE_B200_EINVAL, "null argument");
  cake_b200_ctx *c = (cake_b200_ctx *)calloc(1, sizeof *c);
  c->cfg = *cfg;
  memcpy(&c->oc, cfg, sizeof *cfg);
  c->oc.silu_mode = 0;
  c->m = ora_model_create(&c->oc);
  *out = c; (void)device;
  return 0;
}
void cake_b200_ctx_destroy(cake_b200_ctx *c) { if (c) { ora_model_free(c->m); free(c->last_logits); free(c); } }
int cake_b200_sync(cake_b200_ctx *c) { return c ? 0 : fail(CAKE_B200_EINVAL, "null argument"); }
int cake_b200_dev_alloc(cake_b200_ctx *c, size_t bytes, void **out) { if (!c || !out) return fail(CAKE_B200_EINVAL, "null argument"); *out = malloc(bytes ? bytes : 16); return 0; }
int cake_b200_dev_free(cake_b200_ctx *c, void *p) { (void)c; free(p); return 0; }
int cake_b200_head_load(cake_b200_ctx *c, const void *embed, const void *ln_f, const void *lm_head) {
  if (!c || !embed || !ln_f) return fail(CAKE_B200_EINVAL, "null argument");
  ln_f = own_copy(ln_f, (size_t)c->cfg.hidden * 2);   /* norm vectors may be load-time temporaries (residual_rms_norm): the real library copies everything */
  ora_model_set_head(c->m, embed, ln_f, lm_head ? lm_head : embed);
  return 0;
}
int cake_b200_block_load(cake_b200_ctx *c, int layer, const void *q, const void *k, const void *v, const void *o, const void *gate,
                         const void *up, const void *down, const void *ln1, const void *ln2, const void *qb, const void *kb,
                         const void *vb, const void *qn, const void *kn, cake_b200_block **out) {
  if (!c || !q || !k || !v || !o || !gate || !up || !down || !out) return fail(CAKE_B200_EINVAL, "null weight pointer");
  if (layer < 0 || layer >= c->cfg.n_layers) return fail(CAKE_B200_EINVAL, "layer out of range");
  const size_t sq_ = (size_t)c->cfg.n_heads * c->cfg.head_dim, skv_ = (size_t)c->cfg.n_kv_heads * c->cfg.head_dim;
  if (ln1) ln1 = own_copy(ln1, (size_t)c->cfg.hidden * 2);
  if (ln2) ln2 = own_copy(ln2, (size_t)c->cfg.hidden * 2);
  if (qn) qn = own_copy(qn, (c->cfg.pre_reshape_qk_norm ? sq_ : (size_t)c->cfg.head_dim) * 2);
  if (kn) kn = own_copy(kn, (c->cfg.pre_reshape_qk_norm ? skv_ : (size_t)c->cfg.head_dim) * 2);
  ora_layer l = {q, k, v, o, gate, up, down, ln1, ln2, qb, kb, vb, qn, kn, NULL, NULL, -1, 0};
  ora_model_set_layer(c->m, layer, &l);
  cake_b200_block *b = (cake_b200_block *)malloc(sizeof *b);
  b->layer = layer; b->c = c; b->l = l;
  *out = b;
  return 0;
}
int cake_b200_block_set_variant(cake_b200_block *b, const cake_b200_block_variant *v) {
  if (!b || !v) return fail(CAKE_B200_EINVAL, "null argument");
  if (v->sliding_window < -1) return fail(CAKE_B200_EINVAL, "bad sliding_window");
  if (v->post_attention_norm) b->l.post_attn = own_copy(v->post_attention_norm, (size_t)b->c->cfg.hidden * 2);
  if (v->post_feedforward_norm) b->l.post_ffn = own_copy(v->post_feedforward_norm, (size_t)b->c->cfg.hidden * 2);
  b->l.window = v->sliding_window;
  b->l.no_rope = v->use_rope ? 0 : 1;
  ora_model_set_layer(b->c->m, b->layer, &b->l);
  return 0;
}
void cake_b200_block_free(cake_b200_block *b) { free(b); }
int cake_b200_block_layer(const cake_b200_block *b) { return b ? b->layer : -1; }
int cake_b200_cache_create(cake_b200_ctx *c, int batch, int max_seq, cake_b200_cache **out) {
  if (!c || !out || batch < 1 || max_seq < 1) return fail(CAKE_B200_EINVAL, "bad cache arguments");
  if (max_seq > c->cfg.max_seq) return fail(CAKE_B200_EINVAL, "cache max_seq exceeds config max_seq");
  cake_b200_cache *k = (cake_b200_cache *)calloc(1, sizeof *k);
  k->c = c; k->batch = batch; k->cap = max_seq;
  k->rows = (ora_cache **)calloc((size_t)batch, sizeof *k->rows);
  for (int b = 0; b < batch; b++) k->rows[b] = ora_cache_create(&c->oc, max_seq);
  k->k = k->rows[0];
  *out = k;
  return 0;
}
int cake_b200_cache_clear(cake_b200_cache *k) { if (!k) return fail(CAKE_B200_EINVAL, "null argument"); for (int b = 0; b < k->batch; b++) ora_cache_clear(k->rows[b]); return 0; }
void cake_b200_cache_free(cake_b200_cache *k) { if (k) { for (int b = 0; b < k->batch; b++) ora_cache_free(k->rows[b]); free(k->rows); free(k); } }
float *ora_cache_k(ora_cache *, int);
float *ora_cache_v(ora_cache *, int);
int cake_b200_cache_read(cake_b200_cache *k, int idx, int which, void *out_host, size_t bytes) {
  if (!k || !out_host) return fail(CAKE_B200_EINVAL, "null argument");
  if (idx < 0 || idx >= k->c->cfg.n_layers) return fail(CAKE_B200_EINVAL, "layer has no cache");
  const int len = ora_cache_len(k->k, idx), hd = k->c->cfg.head_dim, nkv = k->c->cfg.n_kv_heads;
  if (bytes < (size_t)k->batch * nkv * len * hd * 2) return fail(CAKE_B200_EINVAL, "cache_read needs more bytes");
  uint16_t *dst = (uint16_t *)out_host;
  for (int b = 0; b < k->batch; b++) {
    const float *src = which ? ora_cache_v(k->rows[b], idx) : ora_cache_k(k->rows[b], idx);   /* (n_kv, cap, hd) */
    for (int h = 0; h < nkv; h++) { from_f32(k->c, src + (size_t)h * k->cap * hd, dst, (size_t)len * hd); dst += (size_t)len * hd; }
  }
  return 0;
}
int cake_b200_cache_len(const cake_b200_cache *k, int i) { return (!k || i < 0 || i >= k->c->cfg.n_layers) ? -1 : ora_cache_len(k->k, i); }

static int forward_f32(cake_b200_ctx *c, cake_b200_block *const *blocks, const int *idx, int n, cake_b200_cache *kc, float *x, int seq, int pos, int row) {
  ora_cache *oc = kc->rows[row];
  if (pos + seq > kc->cap) return fail(CAKE_B200_ESTATE, "index_pos + seq exceeds cache capacity");
  for (int i = 0; i < n; i++) {
    if (idx[i] < 0 || idx[i] >= c->cfg.n_layers) return fail(CAKE_B200_EINVAL, "block_idx out of range");
    if (ora_cache_len(oc, idx[i]) != pos) { snprintf(g_err, sizeof g_err, "block %d: index_pos %d != cache length %d", idx[i], pos, ora_cache_len(oc, idx[i])); return CAKE_B200_ESTATE; }
  }
  float *tmp = (float *)malloc((size_t)seq * c->cfg.hidden * 4);
  for (int i = 0; i < n; i++) {
    if (ora_block_forward(c->m, blocks[i]->layer, oc, x, seq, pos, tmp)) { free(tmp); return fail(CAKE_B200_ESTATE, "oracle block_forward failed"); }
    memcpy(x, tmp, (size_t)seq * c->cfg.hidden * 4);
  }
  free(tmp);
  return 0;
}
int cake_b200_forward_batch(cake_b200_ctx *c, cake_b200_block *const *blocks, const int *idx, int n, cake_b200_cache *kc, const void *x, void *y,
                            int batch, int seq, int pos) {
  if (!c || !blocks || !idx || !kc || !x || !y || n < 1) return fail(CAKE_B200_EINVAL, "null/empty argument");
  if (batch != kc->batch) { snprintf(g_err, sizeof g_err, "batch %d != cache batch %d", batch, kc->batch); return CAKE_B200_EINVAL; }
  if (seq < 1 || pos < 0) return fail(CAKE_B200_ESTATE, "bad position");
  for (int i = 0; i < n; i++) {
    const int w = blocks[i]->l.window >= 0 ? blocks[i]->l.window : c->cfg.sliding_window;
    if (w > 0 && pos > 0 && seq > w) return fail(CAKE_B200_EINVAL, "a chunk on a non-empty cache exceeds the sliding window");
  }
  const size_t ne = (size_t)seq * c->cfg.hidden;
  float *f = (float *)malloc(ne * 4);
  int rc = 0;
  for (int i = 0; i < n && !rc; i++)   /* validate every row's position before anything is appended (all rows advance together) */
    for (int b = 0; b < batch; b++)
      if (ora_cache_len(kc->rows[b], idx[i]) != pos) { snprintf(g_err, sizeof g_err, "block %d: index_pos %d != cache length %d", idx[i], pos, ora_cache_len(kc->rows[b], idx[i])); rc = CAKE_B200_ESTATE; break; }
  for (int b = 0; b < batch && !rc; b++) {   /* every sequence of the batch against its own cache */
    to_f32(c, (const uint16_t *)x + (size_t)b * ne, f, ne);
    rc = forward_f32(c, blocks, idx, n, kc, f, seq, pos, b);
    if (!rc) from_f32(c, f, (uint16_t *)y + (size_t)b * ne, ne);
  }
  free(f);
  return rc;
}
int cake_b200_forward_batch_host(cake_b200_ctx *c, cake_b200_block *const *blocks, const int *idx, int n, cake_b200_cache *kc, const void *x, void *y,
                                 int batch, int seq, int pos) { return cake_b200_forward_batch(c, blocks, idx, n, kc, x, y, batch, seq, pos); }
int cake_b200_embed(cake_b200_ctx *c, const uint32_t *ids, int batch, int seq, void *x) {
  if (!c || !ids || !x || batch != 1) return fail(CAKE_B200_EINVAL, "bad embed arguments");
  for (int i = 0; i < seq; i++) if (ids[i] >= (uint32_t)c->cfg.vocab) return fail(CAKE_B200_EINVAL, "token id out of range");
  const size_t ne = (size_t)seq * c->cfg.hidden;
  float *f = (float *)malloc(ne * 4);
  ora_embed(c->m, ids, seq, f);
  from_f32(c, f, x, ne);
  free(f);
  return 0;
}
static void logits_f32(cake_b200_ctx *c, const void *x, int seq, float *lg) {
  const size_t ne = (size_t)seq * c->cfg.hidden;
  float *f = (float *)malloc(ne * 4);
  to_f32(c, x, f, ne);
  ora_logits(c->m, f, seq, lg);
  free(f);
}
int cake_b200_logits(cake_b200_ctx *c, const void *x, int batch, int seq, void *logits_dev, uint32_t *argmax_host) {
  if (!c || !x || batch != 1) return fail(CAKE_B200_EINVAL, "bad logits arguments");
  float *lg = (float *)malloc((size_t)c->cfg.vocab * 4);
  logits_f32(c, x, seq, lg);
  if (logits_dev) from_f32(c, lg, logits_dev, (size_t)c->cfg.vocab);
  if (argmax_host) *argmax_host = ora_argmax(lg, c->cfg.vocab);
  free(lg);
  return 0;
}
int cake_b200_repeat_penalty_argmax(cake_b200_ctx *c, void *logits_dev, float penalty, const uint32_t *toks, int n, uint32_t *argmax_host) {
  if (!c || !logits_dev || !argmax_host) return fail(CAKE_B200_EINVAL, "null argument");
  float *lg = (float *)malloc((size_t)c->cfg.vocab * 4);
  to_f32(c, logits_dev, lg, (size_t)c->cfg.vocab);
  if (penalty != 1.0f) ora_repeat_penalty(lg, c->cfg.vocab, penalty, toks, n, c->cfg.dtype);
  from_f32(c, lg, logits_dev, (size_t)c->cfg.vocab);
  *argmax_host = ora_argmax(lg, c->cfg.vocab);
  free(lg);
  return 0;
}
/* host-stepped decode loop, world == 1 */
static cake_b200_block g_dec_blocks[512];
int cake_b200_decode_build(cake_b200_ctx *c, cake_b200_block *const *blocks, const int *idx, int n, cake_b200_cache *kc, int rank, int world) {
  if (!c || !blocks || !idx || !kc || n < 1 || n > 512) return fail(CAKE_B200_EINVAL, "bad decode_build arguments");
  for (int i = 0; i < n; i++)
    if (!blocks[i]->l.ln1 || !blocks[i]->l.ln2 || blocks[i]->l.post_attn || blocks[i]->l.post_ffn || blocks[i]->l.window >= 0 || blocks[i]->l.no_rope || c->cfg.pre_reshape_qk_norm)
      return fail(CAKE_B200_EINVAL, "sibling block structure: the decode graph covers the standard block only; step it with cake_b200_forward_batch");
  if (world != 1 || rank != 0) return fail(CAKE_B200_EINVAL, "the emulation handles world == 1");
  c->n_blocks_dec = n; c->dec_cache = kc;
  for (int i = 0; i < n; i++) { c->dec_idx[i] = idx[i]; g_dec_blocks[i] = *blocks[i]; }
  return 0;
}
int cake_b200_decode_begin(cake_b200_ctx *c, uint32_t first_token, int pos) { if (!c || !c->n_blocks_dec) return fail(CAKE_B200_ESTATE, "decode_build has not been called"); c->pos = pos; c->cur_token = first_token; return 0; }
int cake_b200_decode_step_host(cake_b200_ctx *c, uint32_t token_in, uint32_t *token_out) {
  if (!c || !c->n_blocks_dec || !token_out) return fail(CAKE_B200_ESTATE, "decode_build has not been called");
  if (token_in >= (uint32_t)c->cfg.vocab) token_in = 0;
  float *x = (float *)malloc((size_t)c->cfg.hidden * 4);
  ora_embed(c->m, &token_in, 1, x);
  cake_b200_block *bl[512];
  for (int i = 0; i < c->n_blocks_dec; i++) bl[i] = &g_dec_blocks[i];
  int rc = forward_f32(c, bl, c->dec_idx, c->n_blocks_dec, c->dec_cache, x, 1, c->pos, 0);
  if (!rc) {
    if (!c->last_logits) c->last_logits = (float *)malloc((size_t)c->cfg.vocab * 4);
    ora_logits(c->m, x, 1, c->last_logits);
    *token_out = ora_argmax(c->last_logits, c->cfg.vocab);
    c->pos++;
    c->cur_token = *token_out;
    c->ring[c->steps++ % 4096] = *token_out;
  }
  free(x);
  return rc;
}
/* the device-resident loop: every step feeds the previous step's token back */
int cake_b200_decode_run(cake_b200_ctx *c, int n_steps) {
  if (!c || !c->n_blocks_dec) return fail(CAKE_B200_ESTATE, "decode_build has not been called");
  for (int i = 0; i < n_steps; i++) { uint32_t t; int rc = cake_b200_decode_step_host(c, c->cur_token, &t); if (rc) return rc; }
  return 0;
}
int cake_b200_decode_tokens(cake_b200_ctx *c, uint32_t *out, int n) {
  if (!c || !out || n < 0 || (unsigned long long)n > c->steps || n > 4096) return fail(CAKE_B200_EINVAL, "bad decode_tokens arguments");
  for (int i = 0; i < n; i++) out[i] = c->ring[(c->steps - n + i) % 4096];
  return 0;
}
int cake_b200_decode_logits(cake_b200_ctx *c, void *logits_host, size_t bytes) {
  if (!c || !logits_host || !c->last_logits || bytes < (size_t)c->cfg.vocab * 2) return fail(CAKE_B200_EINVAL, "bad decode_logits arguments");
  from_f32(c, c->last_logits, logits_host, (size_t)c->cfg.vocab);
  return 0;
}
'''

DONE_ORACLE = {"cake_b200_last_error", "cake_b200_version", "cake_b200_ctx_create", "cake_b200_ctx_destroy", "cake_b200_sync", "cake_b200_dev_alloc",
               "cake_b200_dev_free", "cake_b200_head_load", "cake_b200_block_load", "cake_b200_block_set_variant", "cake_b200_block_free", "cake_b200_block_layer",
               "cake_b200_cache_create", "cake_b200_cache_clear", "cake_b200_cache_free", "cake_b200_cache_len", "cake_b200_cache_read", "cake_b200_forward_batch",
               "cake_b200_forward_batch_host", "cake_b200_embed", "cake_b200_logits", "cake_b200_repeat_penalty_argmax", "cake_b200_decode_build",
               "cake_b200_decode_begin", "cake_b200_decode_step_host", "cake_b200_decode_run", "cake_b200_decode_tokens", "cake_b200_decode_logits"}

DONE = {"cake_b200_last_error", "cake_b200_version", "cake_b200_ctx_create", "cake_b200_ctx_destroy", "cake_b200_head_load",
        "cake_b200_block_load", "cake_b200_block_free", "cake_b200_cache_create", "cake_b200_cache_free"}


def build(out_dir: str, oracle: bool = False) -> str:
    hand, done = (ORACLE, DONE_ORACLE) if oracle else (HAND, DONE)
    hdr = open(os.path.join(ROOT, "include", "cake_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    stubs = []
    for m in re.finditer(r"^\s*([A-Za-z_][A-Za-z0-9_ \*]*?)\b(cake_b200_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S | re.M):
        ret, name, args = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        if name in done:
            continue
        body = "return;" if ret == "void" else ("return 0;" if "*" in ret else "return -1;")
        # unnamed parameters are legal in prototypes but not in definitions: name them
        params = []
        for i, a in enumerate(x.strip() for x in args.split(",")):
            if a == "void" or not a:
                params.append(a)
            elif re.search(r"[A-Za-z_][A-Za-z0-9_]*$", a) and not re.search(r"(\*|\b(int|float|size_t|uint32_t|uint64_t|void|cake_b200_[a-z]+))\s*$", a):
                params.append(a)
            else:
                params.append(f"{a} a{i}")
        stubs.append(f"{ret} {name}({', '.join(params)}) {{ {body} }}")
    os.makedirs(out_dir, exist_ok=True)
    src = os.path.join(out_dir, "fake_b200.c")
    with open(src, "w") as f:
        f.write(hand + "\n" + "\n".join(stubs) + "\n")
    so = os.path.join(out_dir, "libcake_b200.so")
    # -Bsymbolic: calls between the stand-in's own entry points must not be interposed by a real libcake_b200.so that the
    # same process loaded earlier with RTLD_GLOBAL (pytest runs everything in one interpreter)
    cmd = ["/usr/bin/gcc", "-shared", "-fPIC", "-O1", "-w", "-Wl,-Bsymbolic", "-I", os.path.join(ROOT, "include"), "-o", so, src]
    if oracle:
        sys.path.insert(0, ROOT)
        from oracle import oracle as O
        ora_so = O.build()
        cmd += ["-L", os.path.dirname(ora_so), "-lcake_oracle", "-Wl,-rpath," + os.path.dirname(ora_so)]
    subprocess.check_call(cmd)
    return so


if __name__ == "__main__":
    print(build(sys.argv[1] if len(sys.argv) > 1 else ".", oracle="--oracle" in sys.argv))
