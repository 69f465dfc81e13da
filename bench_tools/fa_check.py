"""A/B of the two prefill attention kernels (tcgen05 vs mma.sync) through one Llama-3-8B-width block.
   python bench_tools/fa_check.py            # driver: runs both modes in subprocesses and compares the block outputs
   python bench_tools/fa_check.py run <out>  # worker: writes outputs for the mode in CAKE_B200_FA"""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

CASES = [(1, 128, 0), (1, 200, 0), (2, 384, 0), (1, 1000, 0), (1, 300, 212), (2, 2048, 0)]  # (batch, seq, chunked-prefill split)

def worker(out):
    import torch
    from cake_b200.config import llama3_8b
    from cake_b200.model import B200Transformer, Cache, Context
    from cake_b200.synth import make_layer
    cfg = llama3_8b(max_seq=2100); cfg.num_hidden_layers = 1
    sd = make_layer(cfg, 0, "bf16", device="cuda")
    ctx = Context(cfg, sd, "bf16", max_seq=2100)
    blk = B200Transformer.load(cfg.layer_name(0), ctx)
    res = {}
    for ci, (B, S, split) in enumerate(CASES):
        ctx.cache = Cache(ctx, B, 2100)
        g = torch.Generator().manual_seed(ci)
        x = ctx.to_device(torch.randn(B, S, cfg.hidden_size, generator=g) * 0.5)
        if split:
            y0 = blk.forward(x[:, :split].contiguous(), 0, 0, ctx)
            y1 = blk.forward(x[:, split:].contiguous(), split, 0, ctx)
            y = torch.cat([y0, y1], dim=1)
        else:
            y = blk.forward(x, 0, 0, ctx)
        ctx.sync()
        res[f"c{ci}"] = y.float().cpu().numpy()
    np.savez(out, **res)

if len(sys.argv) > 1 and sys.argv[1] == "run":
    worker(sys.argv[2]); sys.exit(0)
outs = {}
for mode, tau in (("mma", None), ("tc", "0"), ("tc", "5.545")):
    env = dict(os.environ, CAKE_B200_FA=mode)
    if tau is not None: env["CAKE_B200_FA_TAU"] = tau
    path = f"/tmp/fa_{mode}_{tau}.npz"
    r = subprocess.run([sys.executable, "-X", "faulthandler", __file__, "run", path], env=env, timeout=240, capture_output=True, text=True)
    if r.returncode: print(mode, tau, "FAILED rc", r.returncode, r.stderr[-1500:]); continue
    outs[(mode, tau)] = np.load(path)
ref = outs.get(("mma", None))
for key, o in outs.items():
    if key[0] == "mma" or ref is None: continue
    for ci, (B, S, split) in enumerate(CASES):
        a, b = ref[f"c{ci}"], o[f"c{ci}"]
        d = np.abs(a - b); ulp = d / np.maximum(np.abs(a) * 2.0 ** -8, 1e-30)
        bad = int((d > 0.05).sum())
        print(f"tc tau={key[1]} vs mma  B={B} S={S} split={split}: finite={np.isfinite(b).all()} max|d|={d.max():.4g} mean|d|={d.mean():.3g} "
              f"mismatch(>0.05)={bad}" + (f" first bad idx {np.argwhere(d > 0.05)[0].tolist()}" if bad else ""))
