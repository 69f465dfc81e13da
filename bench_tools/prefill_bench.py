"""Prefill timing of the block forward on Llama-3-8B shapes (profiling aid; the headline metric is decode).
   python bench_tools/prefill_bench.py [n_layers] [batch] [seq]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cake_b200.config import llama3_8b
from cake_b200.model import B200Transformer, Cache, Context
from cake_b200.synth import make_layer

nl = int(sys.argv[1]) if len(sys.argv) > 1 else 4
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
S = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
cfg = llama3_8b(max_seq=S + 8); cfg.num_hidden_layers = nl
sd = {}
for i in range(nl): sd.update(make_layer(cfg, i, "bf16", device="cuda"))
ctx = Context(cfg, sd, "bf16", max_seq=S + 8)
ctx.cache = Cache(ctx, B, S + 8)
blks = [B200Transformer.load(cfg.layer_name(i), ctx) for i in range(nl)]
x = ctx.to_device(torch.randn(B, S, cfg.hidden_size) * 0.5)
batch = [(b.layer_name(), 0, i) for i, b in enumerate(blks)]
H, I = cfg.hidden_size, cfg.intermediate_size
flop_lin = 2.0 * B * S * (cfg.size_q + 2 * cfg.size_kv + cfg.size_q + 3 * I) * H
flop_att = 2.0 * 2 * B * cfg.num_attention_heads * cfg.hd * S * S / 2
for rep in range(3):
    ctx.cache.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.sync()
    e0.record(ctx.torch_stream)
    y = blks[0].forward_batch(x, batch, ctx, blocks=blks)
    e1.record(ctx.torch_stream)
    ctx.sync()
    ms = e0.elapsed_time(e1) / nl
    print(f"rep {rep}: {ms:.3f} ms/layer  B={B} S={S}  linear {flop_lin / ms / 1e9:.1f} TFLOP/s-equivalent (linears only), "
          f"{B * S / (ms * cfg.num_hidden_layers if False else ms) :.0f} tok/s/layer", flush=True)
print("finite:", bool(torch.isfinite(y.float()).all()))
tot = flop_lin + flop_att
print(f"B={B} S={S}: {ms:.3f} ms/layer -> {32 * ms / 1e3:.3f} s for 32 layers = {B * S / (32 * ms / 1e3):.0f} prefill tok/s ; "
      f"{tot / ms / 1e9:.0f} TFLOP/s (linears {flop_lin / 1e12:.2f} + causal attention {flop_att / 1e12:.3f} TFLOP per layer)")
