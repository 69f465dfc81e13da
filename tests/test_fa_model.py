"""Executable model (numpy, f32) of the arithmetic of attn_prefill_tc_kernel (cake_b200/csrc/attn_prefill_tc.cuh): 64-key
tiles, running reference max with the LAZY update (kept while the new max exceeds it by at most tau), p = 2^(s c - m c),
P split into two D operands (P_hi + P_lo) accumulated in f32, one rounding of O / l at the end — against the reference
formulation the oracle restates (attention.rs:300-346: f32 scores, max-subtracted softmax, P V, one rounding).  It pins the
numerical claims the kernel's header makes without needing a GPU: the lazy rule changes nothing beyond f32 rounding, the
two-operand P keeps the result within one ulp of D, and a single D operand (FlashAttention-2's choice) would not."""
import numpy as np
import pytest
import torch

BN = 64


def _rnd(a, dt):
    return torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(dt).float().numpy()


def _kernel_model(q, k, v, dt, tau, split=True):
    """q: (T, hd) rows at positions 0..T-1 (causal), k / v: (T, hd); all values representable in D."""
    T, hd = q.shape
    c = np.float32(1.0 / np.sqrt(hd) * 1.4426950408889634)
    tau2 = np.float32(tau * 1.4426950408889634)
    out = np.empty((T, hd), np.float32)
    for i in range(T):
        m = l = None
        o = np.zeros(hd, np.float32)
        for j0 in range(0, i + 1, BN):
            keys = np.arange(j0, min(j0 + BN, T))
            s = (k[keys] @ q[i]).astype(np.float32)
            s[keys > i] = -np.inf
            mx = s.max()
            if m is None:
                m, l = mx, np.float32(0)
            elif (mx - m) * c > tau2:          # lazy: rescale only when the reference max is more than tau behind
                fac = np.exp2((m - mx) * c).astype(np.float32)
                o *= fac
                l = l * fac
                m = mx
            p = np.exp2(s * c - m * c).astype(np.float32)
            l = np.float32(l + p.sum(dtype=np.float32))
            hi = _rnd(p, dt)
            o += hi @ v[keys]
            if split:
                o += _rnd(p - hi, dt) @ v[keys]
        out[i] = o / l
    return _rnd(out, dt)


def _reference(q, k, v, dt):
    T, hd = q.shape
    s = (q @ k.T).astype(np.float32) * np.float32(1.0 / np.sqrt(hd))
    s[np.triu_indices(T, 1)] = -np.inf
    p = np.exp(s - s.max(axis=1, keepdims=True))
    return _rnd((p / p.sum(axis=1, keepdims=True)) @ v, dt)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_tcgen05_attention_arithmetic_stays_within_one_ulp_of_the_reference_formulation(dtype):
    rng = np.random.default_rng(7)
    T, hd = 300, 128
    q, k, v = (_rnd(rng.standard_normal((T, hd)) * s, dtype) for s in (1.5, 1.5, 1.0))
    ref = _reference(q, k, v, dtype)
    ulp = (2.0 ** -7 if dtype == torch.bfloat16 else 2.0 ** -10) * 2.0 ** np.floor(np.log2(np.maximum(np.abs(ref), 2.0 ** -6)))
    lazy, eager = _kernel_model(q, k, v, dtype, tau=5.545), _kernel_model(q, k, v, dtype, tau=0.0)
    assert (np.abs(lazy - ref) / ulp).max() <= 1.0 and (np.abs(eager - ref) / ulp).max() <= 1.0
    # the lazy rule only moves f32 roundings: after the rounding to D at most a sliver of elements differs, by one ulp
    diff = lazy != eager
    assert diff.mean() < 0.02 and (np.abs(lazy - eager) / ulp).max() <= 1.0
    # what the second operand buys: with P rounded once to D the error is visibly larger (and exceeds 1 ulp in bf16)
    single = _kernel_model(q, k, v, dtype, tau=5.545, split=False)
    assert np.abs(single - ref).mean() > 3 * np.abs(lazy - ref).mean()
    if dtype == torch.bfloat16:
        assert (np.abs(single - ref) / ulp).max() > 1.0
