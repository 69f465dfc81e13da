"""The Python host logic above the C ABI — B200Transformer / TextModelBase / Master and the TCP split with WireRemote —
run end to end on the CPU over the oracle-backed emulation of the library (tests/fake_b200, tests/cpu_ctx.py).  What is
checked is the host side: block walk and grouping (text_model.rs:284-332), positions and cache handling, the sampling
rule incl. repeat penalty (:435-460), goodbye / reset, and the remote-layer path; token ids must equal the oracle's own."""
import numpy as np
import pytest

from cake_b200.model import B200Transformer, Master, TextModelBase
from oracle import oracle as O
from tests.cpu_ctx import CpuContext, use_emulation
from tests.util import checkpoint, medium_config

KW = dict(num_hidden_layers=4, hidden_size=128, intermediate_size=256, vocab_size=256, num_attention_heads=4,
          num_key_value_heads=2, head_dim=32)


def _oracle_tokens(cfg, sd, prompt, n, penalty=1.0, last_n=128):
    om = O.OracleModel(cfg, sd, "bf16", max_seq=64)
    cache, ids, pos, out = om.new_cache(), list(prompt), 0, []
    for _ in range(n):
        lg = om.forward(ids, pos, cache)
        pos += len(ids)
        if penalty != 1.0:
            lg = O.repeat_penalty(O.round_to(lg, "bf16"), penalty, out[-last_n:], "bf16")
        out.append(O.argmax(lg))
        ids = [out[-1]]
    return out


@pytest.mark.parametrize("flavour,penalty", [("llama", 1.0), ("llama", 1.25), ("qwen3_tied", 1.0), ("qwen2_bias", 1.0), ("phi_fused", 1.0)])
def test_master_generate_text_equals_the_oracle(tmp_path, monkeypatch, flavour, penalty):
    use_emulation(monkeypatch, tmp_path)
    extra = {"llama": {}, "qwen3_tied": dict(use_qk_norm=True, tie_word_embeddings=True), "qwen2_bias": dict(use_qkv_bias=True),
             "phi_fused": dict(fused_qkv_proj=True, fused_gate_up_proj=True, partial_rotary_factor=0.5)}[flavour]
    cfg = medium_config(**KW, **extra)
    sd = checkpoint(cfg, "bf16", seed=33, peaked=not cfg.tie_word_embeddings)
    prompt = np.random.default_rng(4).integers(0, cfg.vocab_size - 1, 8).tolist()
    ctx = CpuContext(cfg, sd, "bf16", max_seq=64)
    try:
        model = TextModelBase.load(ctx, repeat_penalty=penalty)
        master = Master(model)
        res = master.generate_text(prompt, 10)
        assert res["tokens"] == _oracle_tokens(cfg, sd, prompt, 10, penalty) and res["generated"] == 10
        assert [ctx.cache.len(i) for i in range(4)] == [8 + 9] * 4          # prompt + 9 fed-back tokens
        # a second prompt on the same model: prepare_prompt clears the cache (text_model.rs:371-395)
        res2 = master.generate_text(prompt[:3], 4)
        assert res2["tokens"] == _oracle_tokens(cfg, sd, prompt[:3], 4, penalty)
        model.reset()
        assert ctx.cache.len(0) == 0 and model.index_pos == 0
        # the host-stepped decode loop (what the graph replays on a GPU) continues a prefilled prompt identically
        if penalty == 1.0:
            model.prepare_prompt(prompt)
            first = model.next_token(0).id
            model.decode_build()
            toks = [first]
            from cake_b200.capi import byref, c_uint32, check, lib
            check(lib().cake_b200_decode_begin(ctx.h, first, model.index_pos))
            for _ in range(5):
                nxt = c_uint32()
                check(lib().cake_b200_decode_step_host(ctx.h, toks[-1], byref(nxt)))
                toks.append(int(nxt.value))
            assert toks == _oracle_tokens(cfg, sd, prompt, 6)
            # and the device-resident loop API (decode_begin / decode_run / decode_tokens) behind TextModelBase.decode_greedy
            model.prepare_prompt(prompt)
            first = model.next_token(0).id
            model.decode_build()
            assert [first] + model.decode_greedy(first, 5) == _oracle_tokens(cfg, sd, prompt, 6)
            # prepare_prompt does not reset `generated` (text_model.rs:371-395; only reset() does, :497-502): 1 + (1 + 5)
            assert model.index_pos == len(prompt) + 5 and model.generated == 7
    finally:
        ctx.close()


def test_eos_stops_generation(tmp_path, monkeypatch):
    use_emulation(monkeypatch, tmp_path)
    cfg = medium_config(**KW)
    sd = checkpoint(cfg, "bf16", seed=33, peaked=True)
    prompt = [5, 9, 200]
    want = _oracle_tokens(cfg, sd, prompt, 8)
    cfg.eos_token_id = [want[3]]
    first_hit = want.index(want[3])
    ctx = CpuContext(cfg, sd, "bf16", max_seq=64)
    try:
        res = Master(TextModelBase.load(ctx)).generate_text(prompt, 8)
        assert res["tokens"] == want[:first_hit]          # master.rs:145-148: the end-of-stream token is not emitted
    finally:
        ctx.close()


def test_master_with_tcp_worker_split_equals_the_oracle(tmp_path, monkeypatch):
    """The deployment shape of tests/test_wire.py::test_gpu_master_with_tcp_worker_generates_the_same_tokens, on the CPU:
    master keeps layers 0-1, a WireWorker(B200Backend) serves layers 2-3 from the shards of a checkpoint on disk, one
    connection per remote layer (text_model.rs:211-227)."""
    from cake_b200.loader import open_model, save_checkpoint
    from cake_b200.parallel import parse_topology
    from cake_b200.wire import B200Backend, WireClient, WireRemote, WireWorker
    use_emulation(monkeypatch, tmp_path)
    cfg = medium_config(**KW)
    sd = checkpoint(cfg, "bf16", seed=23, peaked=True)
    model_dir = tmp_path / "model"
    save_checkpoint(str(model_dir), cfg, sd, shard_bytes=200_000)
    prompt = np.random.default_rng(9).integers(0, cfg.vocab_size - 1, 7).tolist()
    topo = parse_topology({"w1": {"host": "set below", "layers": [f"{cfg.model_prefix}.layers.2-3"]}})
    wcfg, wvb = open_model(str(model_dir), topo, worker="w1")
    wctx = CpuContext(wcfg, wvb, "bf16", max_seq=64)
    w = mctx = None
    clients = []
    try:
        w = WireWorker(B200Backend(wctx, {n: B200Transformer.load(n, wctx) for n in topo["w1"]["layers"]})).start()
        topo["w1"]["host"] = w.address
        mcfg, mvb = open_model(str(model_dir), topo)
        mctx = CpuContext(mcfg, mvb, "bf16", max_seq=64, topology=topo)

        def make_remote(owner, name, c):
            clients.append(WireClient(topo[owner]["host"], name, timeout=30))
            return WireRemote(clients[-1], name, c)

        model = TextModelBase.load(mctx, make_remote=make_remote)
        assert [b.ident() for b in model.blocks] == ["local", "local", w.address, w.address]
        got = Master(model).generate_text(prompt, 10)["tokens"]
        model.goodbye()
        assert got == _oracle_tokens(cfg, sd, prompt, 10) and len(clients) == 2 and w.connections == 2
    finally:
        for c in clients:
            c.close()
        if w is not None:
            w.stop()
        if mctx is not None:
            mctx.close()
        wctx.close()


SMALL = dict(hidden_size=128, intermediate_size=256, vocab_size=256, num_attention_heads=4, num_key_value_heads=2, head_dim=32)


@pytest.mark.parametrize("dtype", ["bf16", "f16"])
def test_late_gpu_case_batched_runs_on_the_emulation(tmp_path, monkeypatch, dtype):
    """tests/test_zz_gpu_batch.py's batched case, same code, emulated library: exact (the emulation is the oracle), which
    shows the test's own indexing / shapes / error expectation are right before it meets a GPU."""
    from tests import cases_late
    use_emulation(monkeypatch, tmp_path)
    cases_late.batched_prefill_then_decode(lambda c, s, d, m: CpuContext(c, s, d, max_seq=m), dtype, 0.0, 0.0, 0.0, **SMALL)


def test_late_gpu_case_phi_runs_on_the_emulation(tmp_path, monkeypatch):
    from tests import cases_late
    use_emulation(monkeypatch, tmp_path)
    cases_late.phi_style_block(lambda c, s, d, m: CpuContext(c, s, d, max_seq=m), "bf16", 0.0, 0.0, 0.0, **SMALL)
