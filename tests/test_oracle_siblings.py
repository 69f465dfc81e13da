"""Properties of the sibling block structures in the oracle that no HF fixture can pin (the reference's Gemma3 local layers
run without RoPE and its EXAONE4 block is pre-norm — transformers does both differently, tests/golden/make_golden.py):
restated from models/{gemma3,exaone4}/block.rs and checked here through what the code implies."""
import numpy as np
import pytest

from cake_b200.synth import residual_deltas
from oracle import oracle as O
from tests.util import checkpoint, medium_config, rand_x

KINDS = {
    "gemma3": dict(block_kind="gemma3", use_qk_norm=True, residual_rms_norm=True, use_gelu_mlp=True, tie_word_embeddings=True,
                   sliding_window=5, global_layers=[False, True, False, True]),
    "exaone4": dict(block_kind="exaone4", use_qk_norm=True, sliding_window=5, global_layers=[False, True, False, True]),
}


def _model(kind, dtype="bf16"):
    cfg = medium_config(num_hidden_layers=4, **KINDS[kind])
    sd = checkpoint(cfg, dtype, seed=21)
    if cfg.residual_rms_norm:
        sd = residual_deltas(sd)
    return cfg, sd, O.OracleModel(cfg, sd, dtype, max_seq=64)


@pytest.mark.parametrize("kind", list(KINDS))
def test_layers_without_rope_do_not_see_positions(kind):
    """attention.rs:242-253: `use_rope == false` skips apply_rotary_emb for q and k — such a layer has no other source of
    position, so the same tokens give the same output wherever they sit; a rotating layer must differ.
    gemma3/block.rs:62-66: LOCAL layers are the ones without RoPE; exaone4/block.rs:52-58: GLOBAL layers are."""
    cfg, _, om = _model(kind)
    x = rand_x((1, 4, cfg.hidden_size), "bf16", seed=22)[0].float().numpy()
    for layer in range(4):
        a = om.block_forward(layer, x, 0, om.new_cache())
        b = om.block_forward(layer, x, 7, om.new_cache())
        no_rope = cfg.layer_variant(layer, 64)["no_rope"]
        assert no_rope == ((not cfg.global_layers[layer]) if kind == "gemma3" else cfg.global_layers[layer])
        if no_rope:
            np.testing.assert_array_equal(a, b)
        else:
            assert np.abs(a - b).max() > 0


def test_exaone4_local_layer_is_the_standard_block_with_that_window():
    """exaone4/block.rs:96-110 is transformer.rs:103-135 (pre-norms) with load_custom(…, Some(window), use_rope = true) on
    local layers: bit-equal to the Llama-structured block of a config whose global sliding_window is that window."""
    cfg, sd, om = _model("exaone4")
    plain = medium_config(num_hidden_layers=4, use_qk_norm=True, sliding_window=5)
    pm = O.OracleModel(plain, sd, "bf16", max_seq=64)
    x = rand_x((1, 20, cfg.hidden_size), "bf16", seed=23)[0].float().numpy()
    ca, cb, pos = om.new_cache(), pm.new_cache(), 0
    for n in (4, 3, 1, 1, 2) + (1,) * 9:   # past the window: the trim is active from the second call on
        np.testing.assert_array_equal(om.block_forward(0, x[pos:pos + n], pos, ca), pm.block_forward(0, x[pos:pos + n], pos, cb))
        pos += n


@pytest.mark.parametrize("kind", list(KINDS))
def test_global_layers_attend_over_everything_and_local_ones_over_the_window(kind):
    """load_custom(…, sliding_window = None) on global layers, Some(w) on local ones (cache.rs:173-205): after the window
    has filled, changing a token that fell out of it must not change a local layer's output and must change a global one's."""
    cfg, _, om = _model(kind)
    x = rand_x((1, 12, cfg.hidden_size), "bf16", seed=24)[0].float().numpy()
    x2 = x.copy()
    x2[1] += 0.5                               # row 1 is outside the last-5 window of position 11
    for layer in range(4):
        outs = []
        for xs in (x, x2):
            c = om.new_cache()
            om.block_forward(layer, xs[:4], 0, c)
            for t in range(4, 11):
                om.block_forward(layer, xs[t:t + 1], t, c)
            outs.append(om.block_forward(layer, xs[11:12], 11, c))
        if cfg.global_layers[layer]:
            assert np.abs(outs[0] - outs[1]).max() > 0
        else:
            np.testing.assert_array_equal(outs[0], outs[1])
