"""CPU restatement of the reference's token sampling (TEST INFRASTRUCTURE ONLY — see cake_oracle.c's header).

text_model.rs:102-118 maps (temperature, top_k, top_p) to a candle_transformers::generation::Sampling and
text_model.rs:429-460 calls LogitsProcessor::sample on the (optionally repeat-penalised) logits.  The algorithms live
in candle-transformers 0.9.2 (generation/mod.rs) and candle-nn 0.9.2 (sampling.rs) — crates.io dependencies that are
not vendored under /root/reference (Cargo.lock:817-933); restated here from their published source, in float32 with the
same operation order:
  sample_argmax            first maximum
  sample_gumbel_softmax    u ~ U(1e-7, 0.999); minus_g = log(-log(u)); T == 1: argmax(l - minus_g) else argmax(l + minus_g * (-T))
  prs(T)                   softmax_last_dim(l * (1/T))  (max-subtracted)
  sample_multinomial       rand::distributions::WeightedIndex: first i whose inclusive prefix sum exceeds x, x ~ U(0, total)
  sample_topp              stable sort descending; zero every token whose preceding running sum is already >= p; multinomial
  sample_topk              the k most probable, multinomial over them
  sample_topk_topp         top-k, then the nucleus rule inside the k, multinomial
PARITY UNPINNED in two places, by construction: the random stream (candle: rand::StdRng / device RNG) is replaced by
uniforms the caller supplies, and candle's TopK candidate ORDER comes from select_nth_unstable (unspecified) — here, and in
csrc/sample.cuh, candidates are ordered by probability descending, ties by lower id."""
from __future__ import annotations

import numpy as np

ARGMAX, ALL, TOPK, TOPP, TOPK_TOPP, GUMBEL = range(6)


def kind_from_args(temperature: float, top_k, top_p) -> int:
    """text_model.rs:102-118."""
    if temperature <= 0.0:
        return ARGMAX
    if top_k is None and top_p is None:
        return GUMBEL
    if top_k is not None and top_p is None:
        return TOPK
    if top_k is None:
        return TOPP
    return TOPK_TOPP


def _multinomial(w: np.ndarray, u: float) -> int:
    w = w.astype(np.float32)
    cum = np.cumsum(w, dtype=np.float32)          # sequential float32 prefix sums, as WeightedIndex builds them
    target = np.float32(u) * cum[-1]
    i = int(np.searchsorted(cum, target, side="right"))   # first i with cum[i] > target
    return min(i, len(w) - 1)


def probs(logits: np.ndarray, temperature: float) -> np.ndarray:
    x = logits.astype(np.float32) * np.float32(1.0 / float(temperature))
    e = np.exp(x - x.max(), dtype=np.float32)
    return (e / e.sum(dtype=np.float32)).astype(np.float32)


def _order(p: np.ndarray) -> np.ndarray:
    """probability descending, ties by lower id (a stable sort on -p)."""
    return np.argsort(-p, kind="stable")


def sample(logits: np.ndarray, kind: int, temperature: float = 1.0, top_k: int = 0, top_p: float = 0.0, noise=None) -> int:
    l = np.asarray(logits, dtype=np.float32)
    V = l.size
    if kind == ARGMAX or temperature <= 0.0:
        return int(np.argmax(l))
    if kind == GUMBEL:
        u = np.float32(1e-7) + np.asarray(noise, np.float32) * (np.float32(0.999) - np.float32(1e-7))
        minus_g = np.log(-np.log(u, dtype=np.float32), dtype=np.float32)
        v = l - minus_g if temperature == 1.0 else l + minus_g * np.float32(-temperature)
        return int(np.argmax(v))
    p = probs(l, temperature)
    u = float(np.asarray(noise, np.float32).reshape(-1)[0])
    want_k = kind in (TOPK, TOPK_TOPP) and top_k < V
    want_p = kind in (TOPP, TOPK_TOPP) and 0.0 < top_p < 1.0
    if not want_k and not want_p:
        return _multinomial(p, u)
    order = _order(p)
    if not want_k:  # TopP over the vocabulary: the kept probabilities are drawn in VOCABULARY order
        kept = p.copy()
        cum = np.float32(0.0)
        for idx in order:
            if cum >= np.float32(top_p):
                kept[idx] = 0.0
            else:
                cum = np.float32(cum + p[idx])
        return _multinomial(kept, u)
    cand = order[:top_k]
    w = p[cand].copy()
    if want_p:
        cum = np.float32(0.0)
        m = 0
        for m in range(len(w) + 1):
            if m == len(w) or cum >= np.float32(top_p):
                break
            cum = np.float32(cum + w[m])
        w = w[:m]
        cand = cand[:m]
    return int(cand[_multinomial(w, u)])
