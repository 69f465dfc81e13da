"""Parity at BASELINE.json's full layer dimensions (Llama-3-8B: H=4096, I=14336, 32/8 heads of 128), where every
kernel runs with its production geometry (32 KB TMA stages, 16-warp megakernel slices, 128x128 UMMA tiles, 64x64
flash tiles): one block against the oracle, plus size-independent properties at depth."""
import numpy as np
import pytest
import torch

from cake_b200.config import llama3_8b
from cake_b200.synth import make_layer
from oracle import oracle as O
from tests.util import max_ulp_err, mean_ulp_err, rand_x, to_np

pytestmark = pytest.mark.gpu


def _ctx(cfg, sd, max_seq):
    from cake_b200.model import Context
    return Context(cfg, sd, "bf16", device=0, max_seq=max_seq)


def test_llama3_8b_block_matches_oracle_prefill_and_decode():
    from cake_b200.model import B200Transformer
    cfg = llama3_8b(max_seq=256)
    cfg.num_hidden_layers = 1
    sd = make_layer(cfg, 0, "bf16", seed=77)
    om = O.OracleModel(cfg, sd, "bf16", max_seq=256)
    oc = om.new_cache(256)
    ctx = _ctx(cfg, sd, 256)
    blk = B200Transformer.load(cfg.layer_name(0), ctx)
    x = rand_x((1, 52, cfg.hidden_size), "bf16", seed=6)
    y_ref = om.block_forward(0, x[0, :48].float().numpy(), 0, oc)      # prefill: tcgen05 GEMMs + flash attention
    y = blk.forward(ctx.to_device(x[:, :48]), 0, 0, ctx)
    ctx.sync()
    e = max_ulp_err(to_np(y[0]), y_ref, "bf16")
    print(f"L8 block prefill(48): max {e:.2f} ulp, mean {mean_ulp_err(to_np(y[0]), y_ref, 'bf16'):.3f} ulp")
    assert e <= 4.0 and mean_ulp_err(to_np(y[0]), y_ref, "bf16") <= 0.25
    for t in range(48, 52):                                             # decode: persistent megakernel
        y_ref = om.block_forward(0, x[0, t:t + 1].float().numpy(), t, oc)
        y = blk.forward(ctx.to_device(x[:, t:t + 1]), t, 0, ctx)
        ctx.sync()
        e = max_ulp_err(to_np(y[0]), y_ref, "bf16")
        assert e <= 4.0, f"decode @{t}: {e} ulp"
    k, v = ctx.cache.kv(0)
    ko, vo = oc.kv(0)
    assert max_ulp_err(to_np(k[0]), ko[:, :52], "bf16") <= 2.0
    assert max_ulp_err(to_np(v[0]), vo[:, :52], "bf16") <= 2.0
    ctx.close()


def test_llama3_8b_decode_equals_prefill_at_depth():
    """Size-independent property: the hidden state of position t computed by token-by-token decode (megakernel,
    flash-decoding over 600+ cached rows, several K/V tiles per split) equals the one computed by a one-shot
    causal prefill (tcgen05 GEMM + flash prefill) — two disjoint kernel paths, same arithmetic."""
    from cake_b200.model import B200Transformer
    cfg = llama3_8b(max_seq=1024)
    cfg.num_hidden_layers = 2
    sd = {}
    for i in range(2):
        sd.update(make_layer(cfg, i, "bf16", seed=5))
    ctx = _ctx(cfg, sd, 1024)
    blks = [B200Transformer.load(cfg.layer_name(i), ctx) for i in range(2)]
    batch = lambda pos: [(b.layer_name(), pos, i) for i, b in enumerate(blks)]
    x = ctx.to_device(rand_x((1, 640, cfg.hidden_size), "bf16", seed=9))
    full = blks[0].forward_batch(x, batch(0), ctx, blocks=blks)         # one-shot prefill of 640
    ctx.sync()
    full = full.cpu()
    ctx.cache.clear()
    blks[0].forward_batch(x[:, :632], batch(0), ctx, blocks=blks)       # prefill 632, then decode 8
    for t in range(632, 640):
        y = blks[0].forward_batch(x[:, t:t + 1], batch(t), ctx, blocks=blks)
        ctx.sync()
        e = max_ulp_err(to_np(y[0, 0]), to_np(full[0, t]), "bf16")
        assert e <= 4.0, f"position {t}: {e} ulp"
    # determinism of the whole path: a second identical run is bit-equal
    ctx.cache.clear()
    again = blks[0].forward_batch(x, batch(0), ctx, blocks=blks)
    ctx.sync()
    assert torch.equal(again.cpu(), full)
    ctx.close()
