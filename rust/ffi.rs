//! Raw bindings to `include/cake_b200.h` (libcake_b200.so).  Written against the header; there is no Rust
//! toolchain in the authoring environment, so this file is reviewed, not compiled there.
//! Conventions follow cake's own FFI precedent `cake-core/src/backends/rocm/ffi.rs` (status-code returns,
//! opaque handles, explicit free functions).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)] pub struct cake_b200_ctx { _p: [u8; 0] }
#[repr(C)] pub struct cake_b200_block { _p: [u8; 0] }
#[repr(C)] pub struct cake_b200_cache { _p: [u8; 0] }

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct cake_b200_config {
    pub hidden: c_int, pub inter: c_int, pub n_heads: c_int, pub n_kv_heads: c_int, pub head_dim: c_int,
    pub n_layers: c_int, pub vocab: c_int, pub max_seq: c_int,
    pub rms_eps: f32, pub rope_theta: f32, pub partial_rotary: f32,
    pub qkv_bias: c_int, pub qk_norm: c_int, pub tie_embeddings: c_int,
    pub rope_llama3: c_int, pub rope_factor: f32, pub rope_low: f32, pub rope_high: f32, pub rope_orig_max: c_int,
    pub dtype: c_int, // 0 = bf16, 1 = f16
    pub sliding_window: c_int, pub use_gelu_mlp: c_int, pub embed_scale: f32,
    pub pre_reshape_qk_norm: c_int, // config.rs:116 (OLMo2)
}

#[link(name = "cake_b200")]
extern "C" {
    pub fn cake_b200_last_error() -> *const c_char;
    pub fn cake_b200_ctx_create(device: c_int, cfg: *const cake_b200_config, out: *mut *mut cake_b200_ctx) -> c_int;
    pub fn cake_b200_ctx_destroy(ctx: *mut cake_b200_ctx);
    pub fn cake_b200_sync(ctx: *mut cake_b200_ctx) -> c_int;
    pub fn cake_b200_block_load(ctx: *mut cake_b200_ctx, layer_idx: c_int,
        q: *const c_void, k: *const c_void, v: *const c_void, o: *const c_void,
        gate: *const c_void, up: *const c_void, down: *const c_void, ln1: *const c_void, ln2: *const c_void,
        q_bias: *const c_void, k_bias: *const c_void, v_bias: *const c_void,
        q_norm: *const c_void, k_norm: *const c_void, out: *mut *mut cake_b200_block) -> c_int;
    pub fn cake_b200_block_free(b: *mut cake_b200_block);
    pub fn cake_b200_cache_create(ctx: *mut cake_b200_ctx, batch: c_int, max_seq: c_int, out: *mut *mut cake_b200_cache) -> c_int;
    pub fn cake_b200_cache_clear(c: *mut cake_b200_cache) -> c_int;
    pub fn cake_b200_cache_free(c: *mut cake_b200_cache);
    pub fn cake_b200_forward_batch(ctx: *mut cake_b200_ctx, blocks: *const *mut cake_b200_block, block_idx: *const c_int,
        n_blocks: c_int, cache: *mut cake_b200_cache, x_dev: *const c_void, y_dev: *mut c_void,
        batch: c_int, seq: c_int, index_pos: c_int) -> c_int;
    pub fn cake_b200_forward_batch_host(ctx: *mut cake_b200_ctx, blocks: *const *mut cake_b200_block, block_idx: *const c_int,
        n_blocks: c_int, cache: *mut cake_b200_cache, x_host: *const c_void, y_host: *mut c_void,
        batch: c_int, seq: c_int, index_pos: c_int) -> c_int;
    pub fn cake_b200_head_load(ctx: *mut cake_b200_ctx, embed: *const c_void, ln_f: *const c_void, lm_head: *const c_void) -> c_int;
    pub fn cake_b200_embed(ctx: *mut cake_b200_ctx, ids_host: *const u32, batch: c_int, seq: c_int, x_dev: *mut c_void) -> c_int;
    pub fn cake_b200_logits(ctx: *mut cake_b200_ctx, x_dev: *const c_void, batch: c_int, seq: c_int,
        logits_dev: *mut c_void, argmax_host: *mut u32) -> c_int;
    pub fn cake_b200_comm_unique_id(out128: *mut c_void) -> c_int;
    pub fn cake_b200_comm_init(ctx: *mut cake_b200_ctx, unique_id128: *const c_void, rank: c_int, world: c_int) -> c_int;
    pub fn cake_b200_send(ctx: *mut cake_b200_ctx, x_dev: *const c_void, bytes: usize, peer: c_int) -> c_int;
    pub fn cake_b200_recv(ctx: *mut cake_b200_ctx, x_dev: *mut c_void, bytes: usize, peer: c_int) -> c_int;
    pub fn cake_b200_decode_build(ctx: *mut cake_b200_ctx, blocks: *const *mut cake_b200_block, block_idx: *const c_int,
        n_blocks: c_int, cache: *mut cake_b200_cache, rank: c_int, world: c_int) -> c_int;
    pub fn cake_b200_decode_begin(ctx: *mut cake_b200_ctx, first_token: u32, index_pos: c_int) -> c_int;
    pub fn cake_b200_decode_run(ctx: *mut cake_b200_ctx, n_steps: c_int) -> c_int;
    pub fn cake_b200_decode_tokens(ctx: *mut cake_b200_ctx, out_host: *mut u32, n: c_int) -> c_int;
    pub fn cake_b200_decode_step_host(ctx: *mut cake_b200_ctx, token_in: u32, token_out: *mut u32) -> c_int;
    pub fn cake_b200_decode_logits(ctx: *mut cake_b200_ctx, logits_host: *mut c_void, bytes: usize) -> c_int;
    pub fn cake_b200_repeat_penalty_argmax(ctx: *mut cake_b200_ctx, logits_dev: *mut c_void, penalty: f32,
        ctx_tokens_host: *const u32, n_tokens: c_int, argmax_host: *mut u32) -> c_int;
    // fused NVLink hand-off between the shards of one box (exchange the 64-byte handles out of band)
    pub fn cake_b200_ring_export(ctx: *mut cake_b200_ctx, handle64: *mut c_void) -> c_int;
    pub fn cake_b200_ring_import(ctx: *mut cake_b200_ctx, next_rank_handle64: *const c_void) -> c_int;
    // introspection / utilities
    pub fn cake_b200_version() -> *const c_char;
    pub fn cake_b200_stream(ctx: *mut cake_b200_ctx) -> *mut c_void;
    pub fn cake_b200_launch_count(ctx: *mut cake_b200_ctx, kernels: *mut u64) -> c_int;
    pub fn cake_b200_dev_alloc(ctx: *mut cake_b200_ctx, bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn cake_b200_dev_free(ctx: *mut cake_b200_ctx, p: *mut c_void) -> c_int;
    pub fn cake_b200_block_layer(b: *const cake_b200_block) -> c_int;
    pub fn cake_b200_cache_len(c: *const cake_b200_cache, block_idx: c_int) -> c_int;
    pub fn cake_b200_cache_read(c: *mut cake_b200_cache, block_idx: c_int, which: c_int, out_host: *mut c_void, bytes: usize) -> c_int;
    pub fn cake_b200_cache_fill_synthetic(c: *mut cake_b200_cache, block_idx: *const c_int, n_blocks: c_int, len: c_int, seed: u32) -> c_int;
    pub fn cake_b200_bench_kernel(ctx: *mut cake_b200_ctx, blocks: *const *mut cake_b200_block, block_idx: *const c_int,
        n_blocks: c_int, cache: *mut cake_b200_cache, which: c_int, reps: c_int, ms_per_launch: *mut f32) -> c_int;
    pub fn cake_b200_decode_trace(ctx: *mut cake_b200_ctx, out_host: *mut u64, n_steps: c_int) -> c_int;
    pub fn cake_b200_sample(ctx: *mut cake_b200_ctx, logits_dev: *mut c_void, sampling: *const cake_b200_sampling, repeat_penalty: f32,
        ctx_tokens_host: *const u32, n_tokens: c_int, step: u64, noise_host: *const f32, token_host: *mut u32) -> c_int;
    pub fn cake_b200_decode_set_sampling(ctx: *mut cake_b200_ctx, sampling: *const cake_b200_sampling) -> c_int;
    pub fn cake_b200_load_stats(ctx: *mut cake_b200_ctx, bytes: *mut f64, seconds: *mut f64) -> c_int;
    pub fn cake_b200_block_set_variant(block: *mut cake_b200_block, variant: *const cake_b200_block_variant) -> c_int;
}

/// `cake_b200_block_variant` (include/cake_b200.h): norm placement and per-layer attention mode of the OLMo2 / Gemma3 /
/// EXAONE4 blocks (models/{olmo2,gemma3,exaone4}/block.rs).
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct cake_b200_block_variant {
    pub sliding_window: c_int, // -1 = the config's, 0 = none, > 0 = this layer's
    pub use_rope: c_int,
    pub post_attention_norm: *const c_void,
    pub post_feedforward_norm: *const c_void,
}

/// `cake_b200_sampling` (include/cake_b200.h): the Sampling enum of candle_transformers::generation flattened.
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct cake_b200_sampling {
    pub kind: c_int,
    pub top_k: c_int,
    pub temperature: f32,
    pub top_p: f32,
    pub seed: u64,
}

/// `anyhow!(cake_b200_last_error())` — the library stringifies failures with context the same way
/// `transformer.rs:113-132` does.
pub fn last_error() -> String {
    unsafe { std::ffi::CStr::from_ptr(cake_b200_last_error()).to_string_lossy().into_owned() }
}
