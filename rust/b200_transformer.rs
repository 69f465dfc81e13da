//! `B200Transformer: Forwarder` — the drop-in for `models::common::Transformer` on a B200.
//! Select it with `type Shardable = B200Transformer;` in `models/llama3/llama.rs:52` and
//! `models/qwen3/model.rs:24`.  Written against `include/cake_b200.h`; not compiled in the authoring
//! environment (no cargo/rustc).  Tensor names are the ones `Transformer::load` reads
//! (`transformer.rs:79-101`, `attention.rs:96-129`, `mlp.rs:38-50`).
use anyhow::{anyhow, Result};
use async_trait::async_trait;
use candle_core::{DType, Device, Storage, Tensor};

use crate::cake::{Context, Forwarder};
use super::ffi;

/// Per-process library state: one `cake_b200_ctx` per CUDA device and one `cake_b200_cache` per cake
/// `Cache` (i.e. per session / worker connection, `worker.rs:60-75`).  Stored in `Context` next to `cache`.
pub struct B200State {
    pub ctx: *mut ffi::cake_b200_ctx,
    pub cache: *mut ffi::cake_b200_cache,
}
unsafe impl Send for B200State {}
unsafe impl Sync for B200State {}

#[derive(Debug)]
pub struct B200Transformer {
    name: String,
    layer: usize,
    handle: *mut ffi::cake_b200_block,
}
unsafe impl Send for B200Transformer {}
unsafe impl Sync for B200Transformer {}

impl std::fmt::Display for B200Transformer {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        write!(f, "{} (local, b200)", &self.name)
    }
}

fn dev_ptr(t: &Tensor) -> Result<*const std::ffi::c_void> {
    // candle: Tensor -> CudaStorage -> device pointer (x is contiguous (b, s, H) in the model dtype)
    let (storage, layout) = t.storage_and_layout();
    match &*storage {
        Storage::Cuda(s) => {
            let off = layout.start_offset() * t.dtype().size_in_bytes();
            Ok((s.as_cuda_slice_ptr()? as usize + off) as *const _)
        }
        _ => Err(anyhow!("B200Transformer needs a CUDA tensor")),
    }
}

fn host_ptr(vb: &candle_nn::VarBuilder, name: &str, shape: &[usize]) -> Result<(Tensor, *const std::ffi::c_void)> {
    // mmapped safetensors tensor on the CPU device in the model dtype; the library copies from host memory
    let t = vb.get(shape, name)?.contiguous()?;
    let p = t.storage_and_layout().0.as_host_ptr()?;
    Ok((t, p))
}

#[async_trait]
impl Forwarder for B200Transformer {
    fn load(name: String, ctx: &Context) -> Result<Box<Self>> {
        let cfg = ctx.config.as_ref().expect("No config specified");
        let vb = ctx.var_builder.as_ref().expect("No var_builder specified").pp(&name);
        let st = ctx.b200.as_ref().expect("B200 state not initialised");
        let h = cfg.hidden_size;
        let hd = cfg.head_dim.unwrap_or(h / cfg.num_attention_heads);
        let (sq, skv, i) = (hd * cfg.num_attention_heads, hd * cfg.num_key_value_heads, cfg.intermediate_size);
        let a = vb.pp("self_attn");
        let m = vb.pp("mlp");
        let (_k0, q) = host_ptr(&a, "q_proj.weight", &[sq, h])?;
        let (_k1, k) = host_ptr(&a, "k_proj.weight", &[skv, h])?;
        let (_k2, v) = host_ptr(&a, "v_proj.weight", &[skv, h])?;
        let (_k3, o) = host_ptr(&a, "o_proj.weight", &[h, sq])?;
        let (_k4, g) = host_ptr(&m, "gate_proj.weight", &[i, h])?;
        let (_k5, u) = host_ptr(&m, "up_proj.weight", &[i, h])?;
        let (_k6, d) = host_ptr(&m, "down_proj.weight", &[h, i])?;
        let (_k7, l1) = host_ptr(&vb, "input_layernorm.weight", &[h])?;
        let (_k8, l2) = host_ptr(&vb, "post_attention_layernorm.weight", &[h])?;
        let null = std::ptr::null();
        let (mut qb, mut kb, mut vbias, mut qn, mut kn) = (null, null, null, null, null);
        let mut keep = vec![];
        if cfg.use_qkv_bias {
            for (dst, n, s) in [(&mut qb, "q_proj.bias", sq), (&mut kb, "k_proj.bias", skv), (&mut vbias, "v_proj.bias", skv)] {
                let (t, p) = host_ptr(&a, n, &[s])?; *dst = p; keep.push(t);
            }
        }
        if cfg.use_qk_norm {
            for (dst, n) in [(&mut qn, "q_norm.weight"), (&mut kn, "k_norm.weight")] {
                let (t, p) = host_ptr(&a, n, &[hd])?; *dst = p; keep.push(t);
            }
        }
        let layer: usize = name.rsplit('.').next().unwrap().parse()?;
        let mut handle = std::ptr::null_mut();
        let rc = unsafe { ffi::cake_b200_block_load(st.ctx, layer as i32, q, k, v, o, g, u, d, l1, l2, qb, kb, vbias, qn, kn, &mut handle) };
        if rc != 0 { return Err(anyhow!("{}: {}", name, ffi::last_error())); }
        Ok(Box::new(Self { name, layer, handle }))
    }

    async fn forward(&self, x: &Tensor, index_pos: usize, block_idx: usize, ctx: &mut Context) -> Result<Tensor> {
        let (b, s, _h) = x.dims3()?;
        let st = ctx.b200.as_ref().expect("B200 state not initialised");
        let x = x.contiguous()?;
        let y = x.zeros_like()?; // output buffer owned by candle, filled by the library
        let blocks = [self.handle];
        let idx = [block_idx as i32];
        let rc = unsafe { ffi::cake_b200_forward_batch(st.ctx, blocks.as_ptr(), idx.as_ptr(), 1, st.cache,
            dev_ptr(&x)?, dev_ptr(&y)? as *mut _, b as i32, s as i32, index_pos as i32) };
        if rc != 0 { return Err(anyhow!("attention/mlp: {}", ffi::last_error())); }
        Ok(y)
    }

    async fn forward_mut(&mut self, x: &Tensor, index_pos: usize, block_idx: usize, ctx: &mut Context) -> Result<Tensor> {
        self.forward(x, index_pos, block_idx, ctx).await
    }

    fn layer_name(&self) -> &str { &self.name }
}

impl Drop for B200Transformer {
    fn drop(&mut self) { unsafe { ffi::cake_b200_block_free(self.handle) } }
}
