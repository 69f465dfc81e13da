"""Two cases written after round 1's GPU budget was spent, hence kept in their own file, last in collection order:

* batch > 1 through the Forwarder (BASELINE.json configs[4] is a bs=32 prefill; `cake run` itself never builds b > 1,
  text_model.rs:418-420, so the reference's contract here is the Forwarder's `x: (b, s, H)`): every sequence of the
  batch must equal the oracle run on that sequence alone with its own cache;
* the Phi-3/4 shape of the block: rotary on the first 3/4 of every head only (`rot < head_dim`), qkv_proj / gate_up_proj
  stored pre-fused.

The bodies live in tests/cases_late.py and are also run on the CPU over the emulated library
(tests/test_host_logic_cpu.py), so the test code itself is known to be right; on the GPU they are NOT YET RUN."""
import pytest

from tests import cases_late

pytestmark = pytest.mark.gpu


def _gpu_ctx(cfg, sd, dtype, max_seq):
    from cake_b200.model import Context
    return Context(cfg, sd, dtype, device=0, max_seq=max_seq)


@pytest.mark.parametrize("dtype", ["bf16", "f16"])
def test_batched_prefill_then_batched_decode_matches_oracle_per_sequence(dtype):
    cases_late.batched_prefill_then_decode(_gpu_ctx, dtype, block_tol=3.0, mean_tol=0.25, kv_tol=2.0)


def test_phi_style_block_partial_rotary_and_prefused_weights():
    cases_late.phi_style_block(_gpu_ctx, "bf16", block_tol=3.0, mean_tol=0.25, kv_tol=2.0)
