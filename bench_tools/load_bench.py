"""Weight-load throughput: a Llama-3-8B-shaped checkpoint of N layers is written to /dev/shm as sharded
safetensors, read back through cake_b200.loader (mmap) and uploaded through cake_b200_block_load's pinned double-buffered
copy pipe into the fused device layouts.   python bench_tools/load_bench.py [n_layers=8]"""
import ctypes, os, shutil, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cake_b200.capi import byref, check, lib
from cake_b200.config import llama3_8b
from cake_b200.loader import open_model, save_checkpoint
from cake_b200.model import B200Transformer, Context
from cake_b200.synth import make_layer

nl = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = llama3_8b(max_seq=256); cfg.num_hidden_layers = nl; cfg.vocab_size = 1024
d = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
try:
    sd = {}
    for i in range(nl):
        sd.update({k: v.cpu() for k, v in make_layer(cfg, i, "bf16", device="cuda").items()})
    g = torch.Generator().manual_seed(1)
    sd["model.embed_tokens.weight"] = torch.randn(cfg.vocab_size, cfg.hidden_size, generator=g).to(torch.bfloat16)
    sd["model.norm.weight"] = torch.ones(cfg.hidden_size, dtype=torch.bfloat16)
    sd["lm_head.weight"] = sd["model.embed_tokens.weight"].clone()
    save_checkpoint(d, cfg, sd, shard_bytes=2_000_000_000)
    del sd
    got, vb = open_model(d)
    ctx = Context(got, vb, "bf16", device=0, max_seq=256)
    t0 = time.perf_counter()
    blocks = [B200Transformer.load(got.layer_name(i), ctx) for i in range(nl)]
    ctx.sync()
    wall = time.perf_counter() - t0
    b, s = ctypes.c_double(), ctypes.c_double()
    check(lib().cake_b200_load_stats(ctx.h, byref(b), byref(s)))
    print(f"{nl} Llama-3-8B layers from mmapped safetensors (page cache): {b.value / 1e9:.2f} GB in {wall:.2f} s wall = "
          f"{b.value / 1e9 / wall:.2f} GB/s (staging + enqueue time inside the library {s.value:.2f} s)")
    ctx.close()
finally:
    shutil.rmtree(d, ignore_errors=True)
