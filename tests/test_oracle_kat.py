"""Pins the CPU oracle against every known-answer test the reference holds for this path
(SURVEY.md §8c).  Citations are to /root/reference/cake-core/tests/unit_tests/."""
import math

import numpy as np
import pytest
import torch

from cake_b200.config import reference_test_config
from cake_b200.synth import make_checkpoint
from oracle import oracle as O


def test_rms_norm_unit_weight_normalizes():
    # test_backend_ops.rs:87-98: rms_norm([3,4], w=1, eps=1e-8) == [3,4]/3.5355 (+-1e-4)
    out = O.rms_norm(np.array([[3.0, 4.0]], np.float32), torch.ones(2), 1e-8, "f32")
    rms = math.sqrt((9 + 16) / 2)
    assert abs(out[0, 0] - 3 / rms) < 1e-4 and abs(out[0, 1] - 4 / rms) < 1e-4


def test_causal_mask_rows():
    # test_backend_ops.rs:215-226 and test_cache.rs:56-74: mask[i][j] = j > i
    m = O.causal_mask(4)
    assert m[0].tolist() == [0, 1, 1, 1]
    assert m[1].tolist() == [0, 0, 1, 1]
    assert m[3].tolist() == [0, 0, 0, 0]


def test_silu_known_values():
    # test_backends.rs:29-38 silu(1)*1 ~= 0.731 (+-0.01); test_backend_ops: silu(0)=0
    assert abs(O.silu_mul(1.0, 1.0, "f32") - 0.731) < 0.01
    assert O.silu_mul(0.0, 5.0, "f32") == 0.0
    # both rounding conventions agree to 2 bf16 ulp on a sweep (tolerance probe)
    for g in np.linspace(-6, 6, 97):
        a, b = O.silu_mul(float(g), 1.5, "bf16", 0), O.silu_mul(float(g), 1.5, "bf16", 1)
        assert abs(a - b) <= 2 ** -6 * max(1.0, abs(a))


def test_cos_sin_at_pos0_and_determinism():
    # test_cache.rs:25-34 cos(pos 0) == 1 (+-1e-5); :37-46 determinism
    cfg = reference_test_config()
    sd = make_checkpoint(cfg, "f32", seed=1)
    m = O.OracleModel(cfg, sd, "f32")
    cos, sin = m.rope_tables()
    assert cos.shape == (cfg.max_seq_len, cfg.hd // 2)
    assert np.all(np.abs(cos[0] - 1.0) < 1e-5) and np.all(sin[0] == 0.0)
    m2 = O.OracleModel(cfg, sd, "f32")
    assert np.array_equal(m2.rope_tables()[0], cos)
    # theta_i = base^(-2i/rot) (cache.rs:43-46)
    i = 3
    assert abs(cos[5, i] - math.cos(5 * 10000.0 ** (-2 * i / cfg.hd))) < 1e-5


def test_kv_cache_accumulation_and_clear():
    # test_cache.rs:77-96 (4 then 1 -> 5) and clear (test_cache.rs:127-150)
    cfg = reference_test_config()
    m = O.OracleModel(cfg, make_checkpoint(cfg, "f32", seed=1), "f32")
    c = m.new_cache()
    x = np.random.default_rng(0).uniform(-0.1, 0.1, (4, cfg.hidden_size)).astype(np.float32)
    m.block_forward(0, x, 0, c)
    assert c.len(0) == 4
    m.block_forward(0, x[:1], 4, c)
    assert c.len(0) == 5 and c.len(1) == 0
    c.clear()
    assert c.len(0) == 0


@pytest.mark.parametrize("kw", [dict(), dict(use_qk_norm=True), dict(use_qkv_bias=True)])
def test_block_scenarios_shape_and_determinism(kw):
    # test_blocks.rs:877-933: prefill (1,4,64); generation (1,1,64); prefill -> generate @4; qk-norm.
    # test_attention.rs determinism: bit-equal across two loads.
    cfg = reference_test_config(**kw)
    sd = make_checkpoint(cfg, "f32", seed=2, std=0.1)
    x = np.random.default_rng(1).uniform(-0.1, 0.1, (5, cfg.hidden_size)).astype(np.float32)
    outs = []
    for _ in range(2):
        m = O.OracleModel(cfg, sd, "f32")
        c = m.new_cache()
        y = m.block_forward(0, x[:4], 0, c)
        assert y.shape == (4, cfg.hidden_size) and np.isfinite(y).all() and not np.allclose(y, x[:4])
        y1 = m.block_forward(0, x[4:5], 4, c)
        assert y1.shape == (1, cfg.hidden_size)
        outs.append((y, y1))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


def test_causality_prefill_equals_incremental():
    # test_backends.rs:96-124 (causal attention masks the future): outputs at position t must not
    # depend on later tokens, so one-shot prefill == token-by-token decode.
    cfg = reference_test_config()
    m = O.OracleModel(cfg, make_checkpoint(cfg, "f32", seed=4, std=0.1), "f32")
    x = np.random.default_rng(2).uniform(-0.5, 0.5, (6, cfg.hidden_size)).astype(np.float32)
    c = m.new_cache()
    full = m.block_forward(1, x, 0, c)
    c2 = m.new_cache()
    inc = np.concatenate([m.block_forward(1, x[t:t + 1], t, c2) for t in range(6)])
    np.testing.assert_allclose(full, inc, rtol=0, atol=2e-6)


def test_repeat_penalty_rule():
    # text_model.rs:60-99 and its inline tests :543-683: >=0 -> /penalty, <0 -> *penalty, dedup.
    lg = np.array([2.0, -2.0, 0.0, 4.0], np.float32)
    out = O.repeat_penalty(lg, 2.0, [0, 1, 0, 2], "f32")
    assert out.tolist() == [1.0, -4.0, 0.0, 4.0]


def test_argmax_first_max_wins():
    assert O.argmax(np.array([1.0, 7.0, 7.0, 3.0], np.float32)) == 1


def test_dtype_rounding_points_bf16():
    # every tensor the reference materialises in D is D-representable in the oracle
    cfg = reference_test_config()
    sd = make_checkpoint(cfg, "bf16", seed=5, std=0.1)
    m = O.OracleModel(cfg, sd, "bf16")
    x = O.round_to(np.random.default_rng(3).uniform(-1, 1, (3, cfg.hidden_size)), "bf16")
    c = m.new_cache()
    y = m.block_forward(0, x, 0, c)
    assert np.array_equal(y, O.round_to(y, "bf16"))
    k, v = c.kv(0)
    assert np.array_equal(k[:, :3], O.round_to(k[:, :3], "bf16"))
    lg = m.logits(y)
    assert np.array_equal(lg, O.round_to(lg, "bf16"))
