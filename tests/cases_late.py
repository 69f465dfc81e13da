"""Bodies of the cases in tests/test_zz_gpu_batch.py, parameterised by how a context is made, so that the same code runs
on a GPU (real library, tolerance in ulps of D) and on the CPU over the emulated library (tolerance 0: the emulation IS
the oracle, what is checked there is the test's own indexing and the host plumbing)."""
import pytest

from oracle import oracle as O
from tests.util import checkpoint, max_ulp_err, mean_ulp_err, medium_config, rand_x, to_np


def batched_prefill_then_decode(make_ctx, dtype, block_tol, mean_tol, kv_tol, **cfg_kw):
    from cake_b200.capi import CakeB200Error
    from cake_b200.model import B200Transformer, Cache
    cfg = medium_config(**cfg_kw)
    sd = checkpoint(cfg, dtype, seed=31)
    B, S = 3, 9
    ctx = make_ctx(cfg, sd, dtype, 64)
    try:
        ctx.cache = Cache(ctx, batch=B, max_seq=64)
        blk = B200Transformer.load(cfg.layer_name(1), ctx)
        x = rand_x((B, S + 2, cfg.hidden_size), dtype, seed=8)
        om = O.OracleModel(cfg, sd, dtype, max_seq=64)
        caches = [om.new_cache() for _ in range(B)]
        y = blk.forward(ctx.to_device(x[:, :S].contiguous()), 0, 1, ctx)          # prefill (B, S, H) at position 0
        ctx.sync()
        assert tuple(y.shape) == (B, S, cfg.hidden_size)
        for b in range(B):
            ref = om.block_forward(1, x[b, :S].float().numpy(), 0, caches[b])
            e = max_ulp_err(to_np(y[b]), ref, dtype)
            assert e <= block_tol and mean_ulp_err(to_np(y[b]), ref, dtype) <= mean_tol, f"prefill seq {b}: {e} ulp"
        for t in (S, S + 1):                                                        # then (B, 1, H) steps
            y = blk.forward(ctx.to_device(x[:, t:t + 1].contiguous()), t, 1, ctx)
            ctx.sync()
            for b in range(B):
                ref = om.block_forward(1, x[b, t:t + 1].float().numpy(), t, caches[b])
                e = max_ulp_err(to_np(y[b]), ref, dtype)
                assert e <= block_tol, f"decode @{t} seq {b}: {e} ulp"
        assert ctx.cache.len(1) == S + 2
        k, v = ctx.cache.kv(1)                                                      # (B, n_kv, len, hd)
        for b in range(B):
            ko, vo = caches[b].kv(1)
            assert max_ulp_err(to_np(k[b]), ko[:, :S + 2], dtype) <= kv_tol
            assert max_ulp_err(to_np(v[b]), vo[:, :S + 2], dtype) <= kv_tol
        # a batch that does not match the cache's batch is an error, not a crash
        with pytest.raises(CakeB200Error, match="batch"):
            blk.forward(ctx.to_device(x[:1, :1].contiguous()), S + 2, 1, ctx)
    finally:
        ctx.close()


def phi_style_block(make_ctx, dtype, block_tol, mean_tol, kv_tol, **cfg_kw):
    from cake_b200.model import B200Transformer
    cfg = medium_config(partial_rotary_factor=0.75, fused_qkv_proj=True, fused_gate_up_proj=True, **cfg_kw)
    sd = checkpoint(cfg, dtype, seed=41)
    om = O.OracleModel(cfg, sd, dtype, max_seq=64)
    oc = om.new_cache()
    ctx = make_ctx(cfg, sd, dtype, 64)
    try:
        blk = B200Transformer.load(cfg.layer_name(1), ctx)
        x = rand_x((1, 12, cfg.hidden_size), dtype, seed=6)
        ref = om.block_forward(1, x[0, :9].float().numpy(), 0, oc)
        y = blk.forward(ctx.to_device(x[:, :9].contiguous()), 0, 1, ctx)         # prefill kernels
        ctx.sync()
        e = max_ulp_err(to_np(y[0]), ref, dtype)
        assert e <= block_tol and mean_ulp_err(to_np(y[0]), ref, dtype) <= mean_tol, f"prefill: {e} ulp"
        for t in range(9, 12):                                                     # decode megakernel
            ref = om.block_forward(1, x[0, t:t + 1].float().numpy(), t, oc)
            y = blk.forward(ctx.to_device(x[:, t:t + 1].contiguous()), t, 1, ctx)
            ctx.sync()
            e = max_ulp_err(to_np(y[0]), ref, dtype)
            assert e <= block_tol, f"decode @{t}: {e} ulp"
        k, _ = ctx.cache.kv(1)
        ko, _ = oc.kv(1)
        assert max_ulp_err(to_np(k[0]), ko[:, :12], dtype) <= kv_tol              # rotated part and pass-through part of K
    finally:
        ctx.close()
