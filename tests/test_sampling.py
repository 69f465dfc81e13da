"""Device-side samplers: text_model.rs:102-118 picks a Sampling from (temperature, top_k, top_p) and
:429-460 draws from the (optionally repeat-penalised) logits.  The CUDA samplers (csrc/sample.cuh, through
cake_b200_sample / the decode graph) are compared with the CPU restatement oracle/sampling.py with the SAME uniforms
supplied to both sides — candle's own random stream cannot be reproduced (see the module headers)."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import sampling as S


def test_kind_selection_follows_create_logits_processor():
    from cake_b200.model import sampling_from_args
    assert S.kind_from_args(0.0, None, None) == S.ARGMAX and sampling_from_args(0.0, 5, 0.9, 1).kind == 0
    assert S.kind_from_args(0.8, None, None) == S.GUMBEL and sampling_from_args(0.8, None, None, 1).kind == 5
    assert S.kind_from_args(0.8, 40, None) == S.TOPK and sampling_from_args(0.8, 40, None, 1).kind == 2
    assert S.kind_from_args(0.8, None, 0.9) == S.TOPP and sampling_from_args(0.8, None, 0.9, 1).kind == 3
    assert S.kind_from_args(0.8, 40, 0.9) == S.TOPK_TOPP and sampling_from_args(0.8, 40, 0.9, 1).kind == 4


def test_restatement_draws_follow_the_distribution():
    rng = np.random.default_rng(3)
    l = (rng.standard_normal(12) * 1.5).astype(np.float32)
    p = S.probs(l, 0.9)
    n = 40000
    for kind, kw, expect in (
            (S.ALL, {}, p),
            (S.TOPK, {"top_k": 4}, None),
            (S.TOPP, {"top_p": 0.7}, None)):
        c = np.zeros(12)
        for u in rng.random(n):
            c[S.sample(l, kind, 0.9, noise=[u], **kw)] += 1
        f = c / n
        if expect is None:
            order = np.argsort(-p, kind="stable")
            keep = order[:4] if kind == S.TOPK else order[:int(np.searchsorted(np.cumsum(p[order]), 0.7) + 1)]
            expect = np.zeros(12)
            expect[keep] = p[keep] / p[keep].sum()
        assert np.abs(f - expect).max() < 0.012
    # Gumbel-max: argmax(l + T g) is a draw from softmax(l / T)
    c = np.zeros(12)
    for _ in range(n):
        c[S.sample(l, S.GUMBEL, 0.9, noise=rng.random(12))] += 1
    assert np.abs(c / n - p).max() < 0.012


def test_topp_ties_and_degenerate_cases():
    l = np.zeros(8, np.float32)          # all equal: ties resolved by id (stable sort)
    assert S.sample(l, S.TOPP, 1.0, top_p=0.3, noise=[0.99]) == 2        # 3 tokens of 1/8 cover 0.3 -> ids 0,1,2
    assert S.sample(l, S.TOPK, 1.0, top_k=8, noise=[0.999]) == 7        # k >= vocab: plain multinomial
    assert S.sample(l, S.TOPP, 1.0, top_p=1.5, noise=[0.0]) == 0        # p outside (0,1): plain multinomial
    assert S.sample(np.array([0, 5, 1], np.float32), S.ALL, 0.0, noise=[0.3]) == 1   # temperature <= 0: argmax


# ------------------------------------------------------------------------------------------------ GPU
def _close_draw(lf, kind, T, k, p, got, want, slack=1e-3):
    """Two correct samplers may return different tokens for the same uniform only because they add the float32
    probabilities in a different order (the reference sums 150 k of them sequentially): accept a draw whose position in the
    sampling order lies within `slack` of probability mass of the restatement's."""
    pr = S.probs(lf, T).astype(np.float64)
    V = pr.size
    srt = np.argsort(-pr, kind="stable")
    if kind in (S.TOPK, S.TOPK_TOPP) and k < V:
        order = srt[:k]
        if kind == S.TOPK_TOPP and 0 < p < 1:
            cum = np.cumsum(pr[order])
            order = order[:int(np.searchsorted(cum, p - slack) + 2)]   # nucleus inside the k, boundary token allowed
        w = pr[order]
    elif kind == S.TOPP and 0 < p < 1:
        cum = np.cumsum(pr[srt])
        kept = srt[:int(np.searchsorted(cum, p + slack) + 2)]
        mask = np.zeros(V, bool)
        mask[kept] = True
        order = np.nonzero(mask)[0]                                    # vocabulary order
        w = pr[order]
    else:
        order, w = np.arange(V), pr
    pos = {int(t): i for i, t in enumerate(order)}
    if got not in pos or want not in pos:
        return False
    c = np.cumsum(w) / w.sum()
    a, b = sorted((pos[got], pos[want]))
    return (c[b - 1] - c[a]) <= slack if b > a + 1 else True           # mass strictly between the two draws


def _ctx(vocab):
    from cake_b200.config import Config
    from cake_b200.model import Context
    cfg = Config(hidden_size=64, intermediate_size=128, vocab_size=vocab, num_hidden_layers=1, num_attention_heads=4,
                 num_key_value_heads=2, max_seq_len=64)
    return Context(cfg, {}, "bf16", device=0)


def _gpu_sample(ctx, logits_bf16, kind, temperature, top_k, top_p, noise, step=0, seed=1, penalty=1.0, pen_tokens=()):
    from cake_b200.capi import CSampling, byref, c_uint32, check, lib, ptr
    s = CSampling(kind, top_k, temperature, top_p, seed)
    d = ctx.to_device(logits_bf16.clone())
    out = c_uint32()
    nz = None if noise is None else np.ascontiguousarray(noise, np.float32).ctypes.data_as(ctypes.POINTER(ctypes.c_float))
    arr = (c_uint32 * max(1, len(pen_tokens)))(*pen_tokens)
    check(lib().cake_b200_sample(ctx.h, ptr(d), byref(s), penalty, arr, len(pen_tokens), step, nz, byref(out)))
    return int(out.value), d


@pytest.mark.gpu
@pytest.mark.parametrize("vocab", [1000, 151936])
def test_gpu_samplers_equal_the_restatement_with_supplied_uniforms(vocab):
    rng = np.random.default_rng(vocab)
    ctx = _ctx(vocab)
    try:
        for trial in range(4):
            scale = (0.5, 2.0, 4.0, 8.0)[trial]                       # flat ... peaked distributions
            l = torch.from_numpy((rng.standard_normal(vocab) * scale).astype(np.float32)).to(torch.bfloat16)
            lf = l.float().numpy()
            cases = [(S.ARGMAX, 0.0, 0, 0.0), (S.ALL, 0.9, 0, 0.0), (S.TOPK, 0.7, 1, 0.0), (S.TOPK, 0.7, 40, 0.0), (S.TOPK, 1.3, 1024, 0.0),
                     (S.TOPP, 0.8, 0, 0.9), (S.TOPP, 1.0, 0, 0.5), (S.TOPP, 1.2, 0, 0.999), (S.TOPK_TOPP, 0.8, 50, 0.9), (S.TOPK_TOPP, 1.0, 8, 0.3)]
            for kind, T, k, p in cases:
                for u in (0.0, 0.2113, 0.5, 0.8731, 0.99999):
                    want = S.sample(lf, kind, T, top_k=k, top_p=p, noise=[u])
                    got, _ = _gpu_sample(ctx, l, kind, T, k, p, [u])
                    if got != want:
                        assert _close_draw(lf, kind, T, k, p, got, want), (kind, T, k, p, u, got, want)
            noise = rng.random(vocab).astype(np.float32)
            for T in (1.0, 0.7):
                want = S.sample(lf, S.GUMBEL, T, noise=noise)
                got, _ = _gpu_sample(ctx, l, S.GUMBEL, T, 0, 0.0, noise)
                assert got == want, ("gumbel", T)
    finally:
        ctx.close()


@pytest.mark.gpu
def test_gpu_repeat_penalty_composes_in_front_of_the_sampler():
    from oracle import oracle as O
    vocab = 2048
    rng = np.random.default_rng(5)
    ctx = _ctx(vocab)
    try:
        l = torch.from_numpy((rng.standard_normal(vocab) * 3).astype(np.float32)).to(torch.bfloat16)
        top = int(np.argmax(l.float().numpy()))
        pen_tokens = [top, 7, top, 9]
        ref = O.repeat_penalty(l.float().numpy(), 1.7, pen_tokens, "bf16")
        want = S.sample(ref, S.TOPK, 0.8, top_k=20, noise=[0.4])
        got, d = _gpu_sample(ctx, l, S.TOPK, 0.8, 20, 0.0, [0.4], penalty=1.7, pen_tokens=pen_tokens)
        assert got == want
        assert np.array_equal(d.float().cpu().numpy(), ref)          # penalised in place, bit-exact (D arithmetic)
    finally:
        ctx.close()


@pytest.mark.gpu
def test_gpu_philox_stream_is_a_proper_sampler():
    """Without supplied uniforms the draws come from Philox(seed, step): frequencies over many steps follow the
    distribution, different seeds give different streams, the same (seed, step) is reproducible."""
    vocab = 1000
    ctx = _ctx(vocab)
    try:
        l = torch.full((vocab,), -30.0).to(torch.bfloat16)
        l[:6] = torch.tensor([2.0, 1.0, 0.0, -1.0, 1.5, 0.5]).to(torch.bfloat16)
        p = S.probs(l.float().numpy(), 1.0)
        n = 3000
        for kind in (S.ALL, S.GUMBEL):
            c = np.zeros(vocab)
            for step in range(n):
                c[_gpu_sample(ctx, l, kind, 1.0, 0, 0.0, None, step=step, seed=11)[0]] += 1
            assert np.abs(c / n - p).max() < 0.03, kind
        a = [_gpu_sample(ctx, l, S.ALL, 1.0, 0, 0.0, None, step=s, seed=11)[0] for s in range(40)]
        b = [_gpu_sample(ctx, l, S.ALL, 1.0, 0, 0.0, None, step=s, seed=12)[0] for s in range(40)]
        assert a == [_gpu_sample(ctx, l, S.ALL, 1.0, 0, 0.0, None, step=s, seed=11)[0] for s in range(40)] and a != b
    finally:
        ctx.close()


@pytest.mark.gpu
def test_gpu_decode_graph_sampler_equals_cake_b200_sample_on_the_same_logits():
    """The sampler kernel captured behind the decode kernel draws with (seed, step): every token of the graph loop must be
    the token cake_b200_sample draws from the logits of that very step with the same counter, and the sampled (not the
    greedy) token must be the one fed back."""
    from cake_b200.capi import byref, c_uint32, check, lib
    from cake_b200.model import Context, TextModelBase
    from tests.util import checkpoint, medium_config
    cfg = medium_config(num_hidden_layers=2)
    sd = checkpoint(cfg, "bf16", seed=31)
    ctx = Context(cfg, sd, "bf16", device=0, max_seq=64)
    try:
        m = TextModelBase.load(ctx, temperature=1.5, top_k=30, top_p=None, seed=77)
        m.prepare_prompt([5, 9, 300, 17])
        t0 = m.next_token(0).id
        m.decode_build()
        check(lib().cake_b200_decode_begin(ctx.h, t0, m.index_pos))
        cur, greedy_differs = t0, 0
        lg = torch.empty(cfg.vocab_size, dtype=torch.bfloat16)
        for i in range(8):
            nxt = c_uint32()
            check(lib().cake_b200_decode_step_host(ctx.h, cur, byref(nxt)))
            check(lib().cake_b200_decode_logits(ctx.h, lg.data_ptr(), lg.numel() * 2))
            d = ctx.to_device(lg)
            o = c_uint32()
            check(lib().cake_b200_sample(ctx.h, d.data_ptr(), byref(m.sampling), 1.0, (c_uint32 * 1)(), 0, i, None, byref(o)))
            assert int(nxt.value) == int(o.value), (i, nxt.value, o.value)
            greedy_differs += int(nxt.value) != int(torch.argmax(lg.float()))
            cur = int(nxt.value)
        assert greedy_differs > 0   # at temperature 1.5 over 30 candidates some draw is not the argmax
        toks = (c_uint32 * 8)()
        check(lib().cake_b200_decode_tokens(ctx.h, toks, 8))
        assert int(toks[7]) == cur    # the ring holds the sampled tokens
    finally:
        ctx.close()
