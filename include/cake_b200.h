/*
 * cake_b200.h — C ABI of libcake_b200.so: the B200-native (sm_100a) implementation of cake's
 * sharded transformer-block forward path.
 *
 * This is the drop-in boundary.  A Rust `B200Transformer: Forwarder` (see INTEGRATION.md and
 * rust/b200_transformer.rs) binds exactly these entry points; nothing here mentions torch or candle
 * types — only opaque handles, plain pointers and sizes.  Conventions follow the reference's two
 * existing FFI surfaces (cake-core/src/backends/rocm/ffi.rs:12-51 function table with int status
 * returns; cake-mobile's cake_mobile_c.h explicit free functions):
 *   - every function returns CAKE_B200_OK (0) or a negative CAKE_B200_E*; the message of the last
 *     failure on the calling thread is available from cake_b200_last_error();
 *   - all handles are thread-compatible: any thread may call (the device is bound inside every call, as
 *     worker.rs:412-418 has to), but calls on ONE ctx share its stream and scratch buffers, so concurrent callers
 *     must serialise them (cake_worker / wire.py hold one lock per ctx); a cake_b200_block is immutable after load and
 *     may be shared by several caches/sessions (worker.rs:60-75 shares blocks across connections, each with its own
 *     Cache); a cake_b200_cache belongs to one session; null handles are reported as CAKE_B200_EINVAL, never
 *     dereferenced;
 *   - "_dev" pointers are device memory on the ctx's GPU, everything is enqueued on the ctx's
 *     stream and is asynchronous unless the name ends in _host or the doc says it synchronises.
 *
 * Reference interface each entry replaces (paths relative to /root/reference/cake-core/src):
 *   cake_b200_ctx_create      Context::from_args device/dtype/config part        cake/mod.rs:114-392
 *   cake_b200_block_load      Forwarder::load -> Transformer::load               cake/mod.rs:513-515, models/common/transformer.rs:79-101,
 *                             (qkv / gate_up fusion done here)                   attention.rs:76-149, mlp.rs:34-59
 *   cake_b200_cache_create    Cache::new / Cache::as_new                         models/common/cache.rs:31-114,241-245
 *   cake_b200_cache_clear     Cache::clear / Message::Goodbye handling           cache.rs:247-253, sharding/worker.rs:364-371
 *   cake_b200_forward_batch   Forwarder::forward / forward_mut / forward_batch   cake/mod.rs:517-543, transformer.rs:103-135
 *   cake_b200_head_load       TextModelBase::load embed/ln_f/lm_head             models/common/text_model.rs:159-193
 *   cake_b200_embed           backend.embedding (index_select)                   text_model.rs:271, backends/mod.rs:513-528
 *   cake_b200_logits          ln_f -> last position -> lm_head (-> ArgMax)       text_model.rs:336-352,102-118
 *   cake_b200_comm_* / send / recv   Client::forward_batch <-> Worker loop tensor hand-off,
 *                             replaced on this path by NCCL p2p over NVLink      sharding/client.rs:79-115,165-174, worker.rs:358-575
 *   cake_b200_decode_*        the per-token hot loop                              sharding/master.rs:131-155, text_model.rs:397-495
 */
#ifndef CAKE_B200_H
#define CAKE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CAKE_B200_OK 0
#define CAKE_B200_EINVAL (-1)  /* bad argument / unsupported configuration */
#define CAKE_B200_ECUDA (-2)   /* CUDA runtime error */
#define CAKE_B200_ENCCL (-3)   /* NCCL error or libnccl not loadable */
#define CAKE_B200_ESTATE (-4)  /* call sequence error (e.g. index_pos != cache length) */
#define CAKE_B200_ENOMEM (-5)

/* model dtype D (cake/mod.rs:75 --dtype): weights, activations and KV cache are stored in D */
#define CAKE_B200_BF16 0
#define CAKE_B200_F16 1

typedef struct cake_b200_ctx cake_b200_ctx;     /* per (process, device): stream, workspaces, graphs, NCCL comm, head weights */
typedef struct cake_b200_block cake_b200_block; /* immutable weights of one transformer layer */
typedef struct cake_b200_cache cake_b200_cache; /* KV cache of one session (== reference Cache); RoPE tables live in the ctx */

/* models/common/config.rs:87-150, the fields the dense Llama-family block reads */
typedef struct cake_b200_config {
  int hidden, inter, n_heads, n_kv_heads, head_dim, n_layers, vocab, max_seq;
  float rms_eps, rope_theta, partial_rotary;
  int qkv_bias, qk_norm, tie_embeddings;
  int rope_llama3; /* cache.rs:49-80 */
  float rope_factor, rope_low, rope_high;
  int rope_orig_max;
  int dtype; /* CAKE_B200_BF16 | CAKE_B200_F16 */
  int sliding_window; /* config.rs:121 + cache.rs:173-205: attention sees the last `sliding_window` positions (0 = full
                         context).  The cache stays append-only in place; a call at index_pos P with s tokens attends from
                         position max(0, P + s - window) when P > 0 and over everything when P == 0 — exactly what the
                         reference's "cat, then keep the last `window`" leaves visible (first call stores everything,
                         test_cache.rs:99-124). */
  int use_gelu_mlp;   /* mlp.rs:25-26: gelu_tanh(gate) * up instead of silu(gate) * up */
  float embed_scale;  /* text_model.rs:274-276: x = embedding * scale (0 = none) */
  int pre_reshape_qk_norm; /* config.rs:116, attention.rs:176-192 (OLMo2): QK-norm over the whole q / k projection (weights
                              [n_heads*head_dim] / [n_kv_heads*head_dim]) before the head reshape, instead of per head */
} cake_b200_config;

const char *cake_b200_last_error(void); /* thread-local; valid until the next call on this thread */
const char *cake_b200_version(void);

/* ---- context ---------------------------------------------------------------------------------- */
int cake_b200_ctx_create(int device, const cake_b200_config *cfg, cake_b200_ctx **out);
void cake_b200_ctx_destroy(cake_b200_ctx *);
int cake_b200_sync(cake_b200_ctx *);              /* cudaStreamSynchronize(ctx stream) */
void *cake_b200_stream(cake_b200_ctx *);          /* the cudaStream_t everything is enqueued on */
int cake_b200_launch_count(cake_b200_ctx *, uint64_t *kernels); /* kernels of this library launched so far (graph replays included) */

/* Device buffers for hosts that have no allocator of their own (the C++ host, tests); synchronous. */
int cake_b200_dev_alloc(cake_b200_ctx *, size_t bytes, void **out);
int cake_b200_dev_free(cake_b200_ctx *, void *p);

/* ---- blocks ----------------------------------------------------------------------------------- */
/* Pointers are HF-layout [out,in] row-major tensors in dtype D, in host OR device memory (copied
 * with cudaMemcpyDefault).  bias / norm pointers may be NULL when the config does not use them.
 * Fuses q,k,v -> Wqkv and gate,up -> Wgu (row-interleaved) into the device layout. */
int cake_b200_block_load(cake_b200_ctx *, int layer_idx, const void *q, const void *k, const void *v, const void *o,
                         const void *gate, const void *up, const void *down, const void *ln1, const void *ln2,
                         const void *q_bias, const void *k_bias, const void *v_bias, const void *q_norm,
                         const void *k_norm, cake_b200_block **out);
void cake_b200_block_free(cake_b200_block *);
int cake_b200_block_layer(const cake_b200_block *); /* global layer index given at load */

/* The sibling block structures — models/olmo2/block.rs:62-90, models/gemma3/block.rs:60-135, models/exaone4/block.rs:50-110 —
 * are the same attention and MLP with a different norm placement and a per-layer attention mode
 * (CausalSelfAttention::load_custom(vb, cfg, use_qk_norm, sliding_window, use_rope), attention.rs:76-149):
 *   - cake_b200_block_load accepts ln1 == NULL and/or ln2 == NULL: no pre-attention / pre-MLP norm (OLMo2);
 *   - post_attention_norm / post_feedforward_norm, when non-NULL ([hidden], D, host or device): RmsNorm of the attention /
 *     MLP output BEFORE the residual add (OLMo2, Gemma3);
 *   - sliding_window: -1 = the config's, 0 = none (a global layer), > 0 = this layer's window (cache.rs:173-205);
 *   - use_rope: 0 = q and k are not rotated on this layer (EXAONE4 global layers, Gemma3 local layers; attention.rs:242-253).
 * Part of loading: call it right after cake_b200_block_load, before the block is used.  A block with any of these set
 * runs its single-token steps through the batched path as well (cake_b200_forward_batch works for every shape);
 * cake_b200_decode_build refuses such blocks with CAKE_B200_EINVAL — hosts step them with cake_b200_forward_batch. */
typedef struct cake_b200_block_variant {
  int sliding_window;
  int use_rope;
  const void *post_attention_norm;
  const void *post_feedforward_norm;
} cake_b200_block_variant;
int cake_b200_block_set_variant(cake_b200_block *, const cake_b200_block_variant *);

/* ---- cache ------------------------------------------------------------------------------------ */
int cake_b200_cache_create(cake_b200_ctx *, int batch, int max_seq, cake_b200_cache **out);
int cake_b200_cache_clear(cake_b200_cache *);
void cake_b200_cache_free(cake_b200_cache *);
int cake_b200_cache_len(const cake_b200_cache *, int block_idx); /* tokens stored for that layer, or <0 */
/* Debug / parity access: copies layer `block_idx`'s K or V, positions [0,len), to host as
 * (batch, n_kv, len, head_dim) in dtype D.  Synchronises. */
int cake_b200_cache_read(cake_b200_cache *, int block_idx, int which /*0=K,1=V*/, void *out_host, size_t bytes);
/* Fill positions [0,len) of every layer this process owns with a deterministic pseudo-random pattern
 * (bench only: lets the decode benchmark start at a given context length without a long prefill). */
int cake_b200_cache_fill_synthetic(cake_b200_cache *, const int *block_idx, int n_blocks, int len, uint32_t seed);

/* ---- the block forward ------------------------------------------------------------------------ */
/* == Forwarder::forward_batch over `n_blocks` consecutive local layers (n_blocks == 1 == forward).
 * x_dev, y_dev: (batch, seq, hidden) in D; may alias.  index_pos: absolute position of x[:,0]
 * and it must equal the cache length of each block (CAKE_B200_ESTATE otherwise — the reference
 * only ever calls it that way, text_model.rs:401-412,454). */
int cake_b200_forward_batch(cake_b200_ctx *, cake_b200_block *const *blocks, const int *block_idx, int n_blocks,
                            cake_b200_cache *, const void *x_dev, void *y_dev, int batch, int seq, int index_pos);
/* Same with HOST buffers: H2D copy of x, forward, D2H copy of y, synchronises (the e2e path). */
int cake_b200_forward_batch_host(cake_b200_ctx *, cake_b200_block *const *blocks, const int *block_idx, int n_blocks,
                                 cake_b200_cache *, const void *x_host, void *y_host, int batch, int seq,
                                 int index_pos);

/* ---- head / tail (master role) ---------------------------------------------------------------- */
int cake_b200_head_load(cake_b200_ctx *, const void *embed, const void *ln_f, const void *lm_head /* NULL if tied */);
int cake_b200_embed(cake_b200_ctx *, const uint32_t *ids_host, int batch, int seq, void *x_dev);
/* ln_f on the last position of each batch row, lm_head, logits (batch, vocab) in D written to
 * logits_dev (nullable); if argmax_host != NULL also computes the greedy token per row
 * (first maximum wins), copies it to argmax_host and synchronises. */
int cake_b200_logits(cake_b200_ctx *, const void *x_dev, int batch, int seq, void *logits_dev, uint32_t *argmax_host);
/* text_model.rs:60-99 repeat penalty on device over logits_dev (vocab) in D, then greedy token. */
int cake_b200_repeat_penalty_argmax(cake_b200_ctx *, void *logits_dev, float penalty, const uint32_t *ctx_tokens_host,
                                    int n_tokens, uint32_t *argmax_host);

/* ---- activation hand-off between shards (NCCL p2p over NVLink, in-stream) ---------------------- */
int cake_b200_comm_unique_id(void *out128); /* ncclGetUniqueId; 128 bytes */
int cake_b200_comm_init(cake_b200_ctx *, const void *unique_id128, int rank, int world);
int cake_b200_send(cake_b200_ctx *, const void *x_dev, size_t bytes, int peer);
int cake_b200_recv(cake_b200_ctx *, void *x_dev, size_t bytes, int peer);

/* Optional, after comm_init: hand-off fused into the decode kernels over NVLink peer memory instead of NCCL kernels.
 * Every rank exports a 64-byte IPC handle of its inbox and imports the handle of rank (r+1) % world; decode_build then
 * makes the last kernel of a shard write the next shard's inbox directly and release it (red.release.sys), and the
 * first kernel of a shard acquire its own inbox.  Without these calls the decode graph uses ncclSend/ncclRecv. */
int cake_b200_ring_export(cake_b200_ctx *, void *handle64);
int cake_b200_ring_import(cake_b200_ctx *, const void *next_rank_handle64);

/* ---- the decode hot loop (master.rs:131-155), one CUDA graph per shard ------------------------- */
/* Builds the per-token step for this rank's contiguous layer range (batch 1, seq 1):
 *   rank 0          : embed(token) -> its blocks -> [send -> ... -> recv from last rank] -> ln_f/lm_head/argmax -> token
 *   ranks 1..world-1: recv -> its blocks -> send to (rank+1) % world
 * Position and token are device-resident, so `n` steps replay back-to-back without host work.
 * world == 1 needs no communicator.  Greedy only (temperature <= 0 and repeat_penalty == 1). */
int cake_b200_decode_build(cake_b200_ctx *, cake_b200_block *const *blocks, const int *block_idx, int n_blocks,
                           cake_b200_cache *, int rank, int world);
/* first_token: token fed at the first step (ignored on ranks > 0); index_pos: its position. */
int cake_b200_decode_begin(cake_b200_ctx *, uint32_t first_token, int index_pos);
int cake_b200_decode_run(cake_b200_ctx *, int n_steps);                    /* async: enqueue n graph replays */
int cake_b200_decode_tokens(cake_b200_ctx *, uint32_t *out_host, int n);   /* last n generated tokens; synchronises (rank 0) */
/* One step driven from the host like TextModelBase::next_token: H2D of the token id, one replay,
 * D2H of the sampled token, synchronise.  (rank 0; other ranks call cake_b200_decode_run(ctx,1).) */
int cake_b200_decode_step_host(cake_b200_ctx *, uint32_t token_in, uint32_t *token_out);
int cake_b200_decode_logits(cake_b200_ctx *, void *logits_host, size_t bytes); /* logits (vocab) of the last step, D; synchronises */

/* ---- sampling on the device (text_model.rs:102-118 create_logits_processor, :429-460) ------------------------------
 * The reference picks a candle_transformers::generation::Sampling from (temperature, top_k, top_p): temperature <= 0
 * -> ArgMax; else (None, None) -> GumbelSoftmax, (k, None) -> TopK, (None, p) -> TopP, (k, p) -> TopKThenTopP ("All" =
 * plain multinomial is what TopK/TopP degrade to for k >= vocab / p outside (0,1)).  Only the 4-byte token leaves the GPU.
 * kind: 0 ArgMax, 1 All, 2 TopK, 3 TopP, 4 TopKThenTopP, 5 GumbelSoftmax.  top_k <= 1024. */
typedef struct cake_b200_sampling {
  int kind, top_k;
  float temperature, top_p;
  uint64_t seed; /* LogitsProcessor::from_sampling(seed, ..): keys a counter-based Philox stream (candle's own RNG stream
                    cannot be reproduced; see csrc/sample.cuh) */
} cake_b200_sampling;
/* next_token's tail in one call: repeat penalty over ctx_tokens_host (text_model.rs:60-99; skipped when penalty == 1 or
 * n_tokens == 0, applied IN PLACE to logits_dev), then one draw.  `step` selects the random numbers of this draw;
 * noise_host (nullable) supplies the uniforms in [0,1) instead — vocab floats for GumbelSoftmax, one float otherwise
 * (how the parity tests pin the arithmetic).  Synchronises. */
int cake_b200_sample(cake_b200_ctx *, void *logits_dev, const cake_b200_sampling *, float repeat_penalty,
                     const uint32_t *ctx_tokens_host, int n_tokens, uint64_t step, const float *noise_host,
                     uint32_t *token_host);
/* Sampler of the graph-captured decode loop (call before cake_b200_decode_build; NULL or kind 0 = greedy, the default):
 * a sampler kernel runs behind the decode kernel of rank 0, reads the logits it left in HBM/L2 and overwrites the greedy
 * token; the draw of step s uses (seed, s).  No repeat penalty inside the graph. */
int cake_b200_decode_set_sampling(cake_b200_ctx *, const cake_b200_sampling *);

/* ---- measurement aid (bench.py roofline leg) ---------------------------------------------------- */
/* Times ONE of the decode kernels of the given blocks in isolation with CUDA events on the ctx stream:
 * `reps` rounds over the blocks, back to back (which: 0 qkv GEMV, 1 o_proj GEMV, 2 gate_up GEMV,
 * 3 down GEMV, 4 attention at the cache's current length).  Different blocks hold distinct weights, so
 * with >= 2 blocks whose matrices exceed L2 no launch is served from cache.  Returns the mean
 * milliseconds per launch.  Synchronises.  Outputs go to scratch buffers; the cache is not advanced. */
int cake_b200_bench_kernel(cake_b200_ctx *, cake_b200_block *const *blocks, const int *block_idx, int n_blocks,
                           cake_b200_cache *, int which, int reps, float *ms_per_launch);

/* Weight upload statistics of this ctx (utils/mod.rs:255-384 is the reference's loader): block_load / head_load stage host
 * tensors through two pinned 64 MB buffers (filled by a few host threads) and copy them with cudaMemcpy(2D)Async on a
 * copy stream straight into the fused qkv / row-interleaved gate_up layouts; compute entry points wait for the uploads
 * issued so far.  bytes = host bytes uploaded, seconds = time spent staging + enqueueing them.  Synchronises the copy
 * stream. */
int cake_b200_load_stats(cake_b200_ctx *, double *bytes, double *seconds);

/* Per-step timeline of this rank's decode kernels for the last `n_steps` steps since cake_b200_decode_begin
 * (n_steps <= 2048): 8 u64 per step = {entry, input acquired, exit, 0} of the layer launch followed by the same for
 * rank 0's head-only launch when sharded (zeros otherwise); %globaltimer nanoseconds of CTA 0.  "input acquired"
 * is when the hidden state of the previous shard arrived (== entry on a single GPU).  Replaces eyeballing the
 * reference's per-hop log lines (worker.rs:505-530 ops/s + read/write stats).  Synchronises. */
int cake_b200_decode_trace(cake_b200_ctx *, uint64_t *out_host, int n_steps);

#ifdef __cplusplus
}
#endif
#endif /* CAKE_B200_H */
